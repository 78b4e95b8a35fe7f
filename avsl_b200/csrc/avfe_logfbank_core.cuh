// Log mel-filterbank (python_speech_features.logfbank) building blocks, compiled for the device
// (avfe_logfbank.cu) and for the host (tests/hostcheck: TEST-ONLY harness, not a product path).
//
// Frames are 400 samples every 160 (no centring, rectangular window), zero-padded to a 512-point
// transform.  Two real frames ride in the real / imaginary parts of one complex 512-point FFT, and
// ONE WARP owns one FFT: 512 = 16 x 32, two rounds of register-resident 16-point DFTs (radix 4 x 4)
// with a single shared-memory exchange between them.  With n = 32a + l and k = r + 16s':
//   stage 1  lane l          Y[l][r] = sum_a z[32a + l] W16^(ar)  (a <= 12: the rest is the zero
//                            padding), times W512^(lr)                              -> S[r][l]
//   stage 2  lane l = r+16q  the 32-point DFT over the lanes of stage 1 for row r, split by the
//                            parity q of s' = 2s + q (decimation in frequency):
//                            u[m] = (S[r][m] + (-1)^q S[r][m+16]) W32^(mq),  X[l + 32s] = DFT16(u)[s]
// so lane l ends up with bins l, l + 32, ..., l + 480 in registers: consecutive lanes hold
// consecutive bins.  The untangling of the two real frames needs X[512 - k] next to X[k]
// (k <= 256), which is x[15 - s] of lane 32 - l: only the upper halves x[8..15] cross lanes,
// through shared memory (U).  The layouts (S rows 33 float2 apart, U[j][lane]) make every 64-bit
// access bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AVFE_HD __host__ __device__ __forceinline__
#else
#define AVFE_HD inline
#endif

namespace avfe {
namespace fbk {

constexpr int kNfft = 512;
constexpr int kFrame = 400;          // winlen 0.025 s at 16 kHz
constexpr int kHop = 160;            // winstep 0.01 s
constexpr int kBins = kNfft / 2 + 1; // 257
constexpr int kSRow = 33;            // float2 stride of one r-row of S (banks 2r, 2r + 1 for a fixed column)
constexpr int kSFloat2 = 16 * kSRow; // 528 float2 per FFT
// A warp's slot: S (1056 floats); once stage 2 has loaded it, the upper-half exchange U (8 x 32
// float2) lives in its tail and the pair's two power rows in its head.  Slots are 1066 floats apart
// and the two rows 261, so that tile frame f = 2w + r starts at bank 5f mod 32: the 16 frames of a
// tile read the same bin from 16 different banks.
constexpr int kSlotFloats = 1066;
constexpr int kPRow = 261;
constexpr int kUOffset = 544;        // floats; U = slot + 544 .. slot + 1056
constexpr float kPreemph = 0.97f;

// Complex add / subtract / scale as ONE packed instruction each on sm_100a (FADD2 / FFMA2: two
// float32 lanes per 64-bit register pair), with the roundings of the scalar form the host build uses.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
AVFE_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
AVFE_HD float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
AVFE_HD float2 caxpy(float s, float2 a, float2 y) { return __ffma2_rn(make_float2(s, s), a, y); }   // s * a + y
#else
AVFE_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AVFE_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AVFE_HD float2 caxpy(float s, float2 a, float2 y) { return make_float2(fmaf(s, a.x, y.x), fmaf(s, a.y, y.y)); }
#endif
AVFE_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward 4-point DFT (W4 = -i), in place
AVFE_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
  x3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
}

// forward 16-point DFT, natural order in and out: n = 4 n1 + n2, k = k1 + 4 k2;
// X[k1 + 4 k2] = sum_n2 W4^(n2 k2) W16^(n2 k1) sum_n1 x[4 n1 + n2] W4^(n1 k1)
AVFE_HD void dft16(float2 (&x)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);     // x[4 k1 + n2] = A[n2][k1]
  // twiddles W16^(n2 k1), W16 = exp(-2 pi i / 16)
  x[5] = cmul(x[5], make_float2(c1, -s1));                                           // W^1
  x[6] = make_float2(h * (x[6].x + x[6].y), h * (x[6].y - x[6].x));                  // W^2 = (1 - i) / sqrt 2
  x[7] = cmul(x[7], make_float2(s1, -c1));                                           // W^3
  x[9] = make_float2(h * (x[9].x + x[9].y), h * (x[9].y - x[9].x));                  // W^2
  x[10] = make_float2(x[10].y, -x[10].x);                                            // W^4 = -i
  x[11] = make_float2(h * (x[11].y - x[11].x), -h * (x[11].x + x[11].y));            // W^6 = -(1 + i) / sqrt 2
  x[13] = cmul(x[13], make_float2(s1, -c1));                                         // W^3
  x[14] = make_float2(h * (x[14].y - x[14].x), -h * (x[14].x + x[14].y));            // W^6
  x[15] = cmul(x[15], make_float2(-c1, s1));                                         // W^9 = -W^1
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);   // x[4 k1 + k2] = X[k1 + 4 k2]
  // transpose to natural order (register renaming once unrolled)
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = k1 + 1; k2 < 4; ++k2) {
      const float2 t = x[4 * k1 + k2];
      x[4 * k1 + k2] = x[4 * k2 + k1];
      x[4 * k2 + k1] = t;
    }
}

// preemphasis(signal, 0.97)[n] in float32, as numpy evaluates signal[1:] - 0.97 * signal[:-1] on a
// float32 signal; samples at or beyond the clip length are the zero padding framesig appends
AVFE_HD float preemph_sample(const float* clip, int64_t len, int64_t n) {
  if (n < 0 || n >= len) return 0.0f;
  if (n == 0) return clip[0];
#if defined(__CUDA_ARCH__)
  return __fsub_rn(clip[n], __fmul_rn(kPreemph, clip[n - 1]));
#else
  const volatile float p = kPreemph * clip[n - 1];
  return clip[n] - p;
#endif
}

// twiddle table: tw1[r * 32 + l] = W512^(l r) (r = 0..15)
AVFE_HD int tw1_index(int r, int l) { return (l * r) & (kNfft - 1); }        // exponent of W512

// stage 1, lane l.  ya / yb: the two frames' 400 pre-emphasised samples
AVFE_HD void fft_stage1(int l, const float* ya, const float* yb, const float2* tw1, float2* S) {
  float2 x[16];
#pragma unroll
  for (int a = 0; a < 12; ++a) x[a] = make_float2(ya[32 * a + l], yb[32 * a + l]);
  x[12] = (l < kFrame - 384) ? make_float2(ya[384 + l], yb[384 + l]) : make_float2(0.0f, 0.0f);
  x[13] = x[14] = x[15] = make_float2(0.0f, 0.0f);
  dft16(x);
  S[l] = x[0];
#pragma unroll
  for (int r = 1; r < 16; ++r) S[r * kSRow + l] = cmul(x[r], tw1[r * 32 + l]);
}

// stage 2, lane l = r + 16 q: afterwards x[s] = X[l + 32 s].  The twiddles W32^(m q) are selected
// immediates applied under a predicate (q = 0: none) instead of a table: no shared-memory traffic.
AVFE_HD void fft_stage2(int l, const float2* S, float2 (&x)[16]) {
  // W32^m = exp(-2 pi i m / 32), m = 1..15
  constexpr float kC[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                            0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                            0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                            -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  constexpr float kS[16] = {0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                            -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f,
                            -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254524f,
                            -0.70710678118654752f, -0.55557023301960218f, -0.38268343236508977f, -0.19509032201612825f};
  const int r = l & 15, q = l >> 4;
  const float2* row = S + r * kSRow;
  const float sgn = q ? -1.0f : 1.0f;
#pragma unroll
  for (int m = 0; m < 16; ++m) x[m] = caxpy(sgn, row[m + 16], row[m]);
  if (q) {                                                           // predicated multiplies by immediates
#pragma unroll
    for (int m = 1; m < 16; ++m) x[m] = cmul(x[m], make_float2(kC[m], kS[m]));
  }
  dft16(x);
}

// upper halves to U[j][lane] (j = s - 8)
AVFE_HD void fft_upper_store(int l, const float2 (&x)[16], float2* U) {
#pragma unroll
  for (int j = 0; j < 8; ++j) U[j * 32 + l] = x[8 + j];
}

// power spectra 1/512 |X|^2 of the two real frames: lane l owns bins l + 32 s, s = 0..7, lane 0 also 256
AVFE_HD void fft_power(int l, const float2 (&x)[16], const float2* U, float* Pa, float* Pb) {
  const float scale = 0.25f / (float)kNfft;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const float2 z = x[s];
    float2 zm;
    if (l == 0) zm = (s == 0) ? x[0] : U[(8 - s) * 32];             // X[512 - 32 s] = own x[16 - s]
    else zm = U[(7 - s) * 32 + 32 - l];                             // x[15 - s] of lane 32 - l
    const float ar = z.x + zm.x, ai = z.y - zm.y, br = z.y + zm.y, bi = zm.x - z.x;
    Pa[l + 32 * s] = scale * (ar * ar + ai * ai);
    Pb[l + 32 * s] = scale * (br * br + bi * bi);
  }
  if (l == 0) {                                                      // Nyquist bin: its own mirror
    const float ar = x[8].x + x[8].x, br = x[8].y + x[8].y;
    Pa[kNfft / 2] = scale * (ar * ar);
    Pb[kNfft / 2] = scale * (br * br);
  } else if (l < 5) {                                                // slots the quad-padded filter taps may read (zero weights)
    Pa[kNfft / 2 + l] = 0.0f;
    Pb[kNfft / 2 + l] = 0.0f;
  }
}

// feat = fb_row . pspec over [lo, hi); zero -> float64 eps (numpy.finfo(float).eps); natural log
AVFE_HD float log_energy(float acc) {
  if (acc == 0.0f) acc = 2.220446049250313e-16f;
#if defined(__CUDA_ARCH__)
  return __log2f(acc) * 0.69314718055994531f;   // MUFU.LG2: abs error ~2e-6 over the range of the energies
#else
  return logf(acc);
#endif
}
AVFE_HD float log_fbank(const float* P, const float* w, int lo, int hi) {
  float acc = 0.0f;
#pragma unroll 4
  for (int k = lo; k < hi; ++k) acc = fmaf(w[k], P[k], acc);
  return log_energy(acc);
}

}  // namespace fbk
}  // namespace avfe
