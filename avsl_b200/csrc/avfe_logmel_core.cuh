// Log-mel building blocks, compiled for the device (avfe_logmel.cu) and for the host
// (tests/hostcheck: TEST-ONLY harness that steps these codelets thread by thread on the CPU
// to check the index algebra against the oracle; not a product path).
//
// STFT of one tile = 32 consecutive frames of one clip = 16 complex 400-point FFTs
// (frames 2g and 2g+1 ride in the real and imaginary parts of FFT g).  400 = 20 x 20
// Cooley-Tukey; each 20-point DFT is a register-resident 4 x 5 prime-factor (Good-Thomas)
// codelet with no internal twiddles.
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
namespace lm {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kTileFrames = 32;
constexpr int kPairs = kTileFrames / 2;                          // 16 complex FFTs per tile
constexpr int kThreads = kPairs * 20;                            // 320: one thread per (FFT, column)
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kNfft;   // 5360
constexpr int kZRow = 21;            // float2 row stride of the 20x20 exchange (20 + 1 pad)
// float2 per FFT: 840 words = 8 (mod 32), so a warp's tail lanes (next FFT, columns 0..11) land
// on banks 8..31 while its lanes 16..19 use banks 0..7: conflict-free 64-bit exchanges
constexpr int kZPair = 20 * kZRow;
// After the spectra have been consumed the whole Z storage is free; the 32 power rows are
// written there with an odd stride so that lane = frame reads hit 32 different banks.
constexpr int kPStride = 203;        // >= 201 bins + the 3 padded reads of the quad-packed mel filters
// The staged audio span is skewed by 20 floats per 320 samples: FFT g+1 reads exactly 320
// samples after FFT g, i.e. the same banks; the skew moves its lanes (columns 0..11) to banks
// 20..31, next to lanes 0..19 of FFT g.
constexpr int kTileSkew = 20;
constexpr int kTileFloats = kTileSamples + kTileSkew * (kTileSamples / 320 + 1);   // 5700
constexpr int kBinsPerThread = 11;   // thread (g, j) untangles bins k = j + 20*m, m = 0..10

// Complex add / subtract / scale as ONE packed instruction each on sm_100a (FADD2 / FFMA2 /
// FMUL2: two float32 lanes per 64-bit register pair, Blackwell-only), with the same roundings as
// the scalar form, which the host build (tests/hostcheck) uses.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
AVFE_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
AVFE_HD float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }   // a - b, exact negation
AVFE_HD float2 cscale(float s, float2 a) { return __fmul2_rn(make_float2(s, s), a); }             // s * a
AVFE_HD float2 caxpy(float s, float2 a, float2 y) { return __ffma2_rn(make_float2(s, s), a, y); } // s * a + y
#else
AVFE_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AVFE_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AVFE_HD float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
AVFE_HD float2 caxpy(float s, float2 a, float2 y) { return make_float2(fmaf(s, a.x, y.x), fmaf(s, a.y, y.y)); }
#endif
AVFE_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward 4-point DFT (W = -i)
AVFE_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  x1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i*t3
  x3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i*t3
}

// forward 5-point DFT
AVFE_HD void dft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 m1 = caxpy(c2, t2, caxpy(c1, t1, x0));
  const float2 m2 = caxpy(c1, t2, caxpy(c2, t1, x0));
  const float2 n1 = caxpy(s2, t4, cscale(s1, t3));
  const float2 n2 = caxpy(-s1, t4, cscale(s2, t3));
  x0 = cadd(cadd(x0, t1), t2);
  x1 = make_float2(m1.x + n1.y, m1.y - n1.x);   // m1 - i*n1
  x4 = make_float2(m1.x - n1.y, m1.y + n1.x);   // m1 + i*n1
  x2 = make_float2(m2.x + n2.y, m2.y - n2.x);
  x3 = make_float2(m2.x - n2.y, m2.y + n2.x);
}

// In-place forward 20-point DFT, natural order in and out.
// Good-Thomas: n = (5*n1 + 4*n2) mod 20, k = (5*k1 + 16*k2) mod 20.
AVFE_HD void dft20(float2 (&x)[20]) {
  float2 a[4][5];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1)
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) a[n1][n2] = x[(5 * n1 + 4 * n2) % 20];
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) dft4(a[0][n2], a[1][n2], a[2][n2], a[3][n2]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft5(a[k1][0], a[k1][1], a[k1][2], a[k1][3], a[k1][4]);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) x[(5 * k1 + 16 * k2) % 20] = a[k1][k2];
}

// Source sample for padded-signal position p (p = 0 is 200 samples before the clip start):
// torch.stft(center=True, pad_mode="reflect") over the clip zero-extended to Lp = L + padding.
AVFE_HD float padded_sample(const float* clip, int64_t L, int64_t Lp, int64_t p) {
  int64_t j = p - kNfft / 2;
  if (j < 0) j = -j;
  if (j >= Lp) j = 2 * (Lp - 1) - j;
  if (j < 0 || j >= L) return 0.0f;   // zero padding (or beyond the last computed frame)
  return clip[j];
}

// Stage 1, thread (g, j): column j of FFT g.  z[m] = hann[j+20m] * (xa + i*xb)[j+20m];
// 20-point DFT over m; times W400^(j*k1); stored at Z[g][k1][j].
// tw[k1*20 + j] = exp(-2*pi*i*j*k1/400).
// physical position of span sample s in the skewed tile
AVFE_HD int tile_pos(int s) { return s + kTileSkew * (s / 320); }

// Stage 1 comes in two halves so that the audio span can share its storage with Z (which is what
// lets three CTAs fit on an SM): stage1_load puts the thread's 2 x 20 windowed samples in registers;
// after a CTA barrier (every thread has its samples) stage1_dft transforms and overwrites the span.
AVFE_HD void stage1_load(int g, int j, const float* tile, const float* hann, float2 (&x)[20]) {
  // frame 2g starts at span sample 320g, frame 2g+1 160 later; within the thread's 20 reads the
  // skew block index is g plus one or two carries that depend on (j + 20m) only
  const float* base = tile + 320 * g + kTileSkew * g + j;
#pragma unroll
  for (int m = 0; m < 20; ++m) {
    const float w = hann[j + 20 * m];
    const int oa = 20 * m, ob = 160 + 20 * m;            // offsets inside the pair's 320-blocks
    const int ca = (oa + j >= 320) ? kTileSkew : 0;      // oa + j <= 399
    const int cb = (ob + j >= 320) ? kTileSkew : 0;      // ob + j <= 559
    x[m] = make_float2(w * base[oa + ca], w * base[ob + cb]);
  }
}

AVFE_HD void stage1_dft(int g, int j, float2 (&x)[20], const float2* tw, float2* Z) {
  dft20(x);
  float2* z = Z + g * kZPair + j;
  z[0] = x[0];
#pragma unroll
  for (int k1 = 1; k1 < 20; ++k1) z[k1 * kZRow] = cmul(x[k1], tw[k1 * 20 + j]);
}

AVFE_HD void stage1(int g, int j, const float* tile, const float* hann, const float2* tw,
                    float2* Z) {
  float2 x[20];
  stage1_load(g, j, tile, hann, x);
  stage1_dft(g, j, x, tw, Z);
}

// Stage 2, thread (g, k1): 20-point DFT over j of row k1, in place:
// slot [g][k1][k2] then holds bin k = k1 + 20*k2.
AVFE_HD void stage2(int g, int k1, float2* Z) {
  float2 x[20];
  float2* z = Z + g * kZPair + k1 * kZRow;
#pragma unroll
  for (int j = 0; j < 20; ++j) x[j] = z[j];
  dft20(x);
#pragma unroll
  for (int k2 = 0; k2 < 20; ++k2) z[k2] = x[k2];
}

// Stage 2 that keeps its result in registers: after the DFT thread (g, k1) holds bins k1 + 20*k2 in
// x[k2].  The untangling of bin k needs bin 400-k, which is x[19-k2] of thread (g, 20-k1): only the
// upper half x[10..19] is ever read by a partner, so only that half goes back to shared memory (into
// the thread's own row: nobody else touches it during this stage, hence no barrier in front).
AVFE_HD void stage2_keep(int g, int k1, float2* Z, float2 (&x)[20]) {
  float2* z = Z + g * kZPair + k1 * kZRow;
#pragma unroll
  for (int j = 0; j < 20; ++j) x[j] = z[j];
  dft20(x);
#pragma unroll
  for (int k2 = 10; k2 < 20; ++k2) z[k2] = x[k2];
}

// Untangle from registers (after a barrier behind stage2_keep): own bins k = j + 20*m from x[m], the
// mirrors from the partner's row; thread j = 0 (bins 20*m) is its own partner.  Same arithmetic as
// split_load.
AVFE_HD void split_from_regs(int g, int j, const float2* Z, const float2 (&x)[20], float (&pa)[kBinsPerThread],
                             float (&pb)[kBinsPerThread]) {
  const float2* zp = Z + g * kZPair + (20 - j) * kZRow;      // partner row (j > 0)
#pragma unroll
  for (int m = 0; m < kBinsPerThread; ++m) {
    if (m == 10 && j != 0) { pa[m] = 0.0f; pb[m] = 0.0f; continue; }
    const float2 u = x[m];
    float2 v;
    if (j == 0) v = (m == 0) ? x[0] : x[20 - m];
    else v = zp[19 - m];
    const float ar = u.x + v.x, ai = u.y - v.y;
    const float br = u.y + v.y, bi = v.x - u.x;
    pa[m] = 0.25f * (ar * ar + ai * ai);
    pb[m] = 0.25f * (br * br + bi * bi);
  }
}

// Float offset of the power row of tile frame f (0..31) inside the Z storage.
AVFE_HD int prow_offset(int f) { return f * kPStride; }

// Untangle the two real frames packed in FFT g.  Thread (g, j) owns bins k = j + 20*m
// (m = 0..10; m = 10 exists only for j = 0, k = 200).  Bin k sits in slot j*21 + m; its mirror
// 400-k in slot (20-j)*21 + (19-m) for j > 0 and slot 20-m for j = 0 (slot 0 for k = 0).
// Xa = (Z[k] + conj(Z[400-k]))/2, Xb = (Z[k] - conj(Z[400-k]))/(2i); powers kept in registers
// (phase A) so that phase B can overwrite the Z storage with the power rows after a barrier.
AVFE_HD void split_load(int g, int j, const float2* Z, float (&pa)[kBinsPerThread],
                        float (&pb)[kBinsPerThread]) {
  const float2* z = Z + g * kZPair;
#pragma unroll
  for (int m = 0; m < kBinsPerThread; ++m) {
    if (m == 10 && j != 0) { pa[m] = 0.0f; pb[m] = 0.0f; continue; }
    const int ms = (j == 0) ? ((m == 0) ? 0 : 20 - m) : (20 - j) * kZRow + (19 - m);
    const float2 u = z[j * kZRow + m];
    const float2 v = z[ms];
    const float ar = u.x + v.x, ai = u.y - v.y;
    const float br = u.y + v.y, bi = v.x - u.x;
    pa[m] = 0.25f * (ar * ar + ai * ai);
    pb[m] = 0.25f * (br * br + bi * bi);
  }
}

AVFE_HD void split_store(int g, int j, float* P, const float (&pa)[kBinsPerThread],
                         const float (&pb)[kBinsPerThread]) {
  float* ra = P + prow_offset(2 * g) + j;
  float* rb = P + prow_offset(2 * g + 1) + j;
#pragma unroll
  for (int m = 0; m < kBinsPerThread; ++m) {
    if (m == 10 && j != 0) continue;
    ra[20 * m] = pa[m];
    rb[20 * m] = pb[m];
  }
  // the quad-packed mel filters may read up to 3 floats past bin 200 (with zero weights): keep
  // those slots finite (0 * NaN would poison the dot product)
  if (j == 1) { ra[200] = 0.0f; ra[201] = 0.0f; rb[200] = 0.0f; rb[201] = 0.0f; }   // bins 201, 202
  if (j == 2 && g == kPairs - 1) rb[kPStride - 2] = 0.0f;                          // one past the last row
}

AVFE_HD float log10_floor(float acc) {
  acc = fmaxf(acc, 1e-10f);
#if defined(__CUDA_ARCH__)
  return __log2f(acc) * 0.30102999566398120f;   // MUFU.LG2: abs error ~1e-6 in log10 units
#else
  return log2f(acc) * 0.30102999566398120f;
#endif
}

// log10(max(mel, 1e-10)) for one (frame, filter): dot product over the filter's support.
// Prow points at the first bin of the support; w holds the filter's weights zero-padded to a
// multiple of four and 16-byte aligned (n4 = number of quads), so there is no tail loop.  The
// power rows are long enough (433 floats) for the padded reads.
AVFE_HD float mel_log10_quads(const float* Prow, const float4* w, int n4) {
  float a0 = 0.0f, a1 = 0.0f;
  // slaney filters need 1..4 quads: one guarded, fully unrolled block of four per trip
  for (int q0 = 0; q0 < n4; q0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u;
      if (q < n4) {
        const float4 c = w[q];
        a0 = fmaf(c.x, Prow[4 * q], a0);
        a1 = fmaf(c.y, Prow[4 * q + 1], a1);
        a0 = fmaf(c.z, Prow[4 * q + 2], a0);
        a1 = fmaf(c.w, Prow[4 * q + 3], a1);
      }
    }
  }
  return log10_floor(a0 + a1);
}

// generic form (weights anywhere, any length) for filterbanks too dense to be packed
AVFE_HD float mel_log10(const float* Prow, const float* w, int n) {
  float acc = 0.0f;
  for (int k = 0; k < n; ++k) acc = fmaf(w[k], Prow[k], acc);
  return log10_floor(acc);
}

// order-preserving float <-> int key for atomicMax
AVFE_HD int float_key(float f) {
  union { float f; int i; } u;
  u.f = f;
  return u.i >= 0 ? u.i : (u.i ^ 0x7fffffff);
}
AVFE_HD float key_float(int k) {
  union { float f; int i; } u;
  u.i = k >= 0 ? k : (k ^ 0x7fffffff);
  return u.f;
}

}  // namespace lm
}  // namespace avfe
