// Log mel-filterbank (python_speech_features.logfbank) building blocks, compiled for the device
// (avfe_logfbank.cu) and for the host (tests/hostcheck: TEST-ONLY harness, not a product path).
//
// Frames are 400 samples every 160 (no centring, rectangular window), zero-padded to a 512-point
// transform.  Two real frames ride in the real / imaginary parts of one complex 512-point FFT,
// 512 = 8 x 8 x 8: three rounds of register-resident 8-point DFTs, 64 threads per FFT, two
// shared-memory exchanges.  With n = 64a + m, k = r + 8s (and inside the 64-point step
// m = 8b + c, s = u + 8v):
//   step 1  thread m       Y[m][r] = sum_a x[64a+m] W8^(ar),  times W512^(mr)      -> S[r][m]
//   step 2  thread (r,c)   T[u]    = sum_b S[r][8b+c] W8^(bu), times W512^(8cu)    -> S'[u][c][r]
//   step 3  thread (u,r)   X[r + 8u + 64v] = sum_c S'[u][c][r] W8^(cv)              -> C[k]
// (layouts chosen so that every 64-bit exchange is bank-conflict free and the spectrum is written
// in natural order by consecutive lanes)
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
constexpr int kFftThreads = 64;
constexpr int kRow = 72;             // float2 stride of one r-row in S (8 x 9 for step 2's layout)
constexpr int kSFloat2 = 8 * kRow;   // 576 float2 per FFT
constexpr int kPStride = 264;        // floats per power row (257 + pad)
constexpr float kPreemph = 0.97f;

AVFE_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AVFE_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AVFE_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
AVFE_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 8-point DFT, natural order in and out (radix-2 decimation in frequency)
AVFE_HD void dft8(float2 (&x)[8]) {
  const float h = 0.70710678118654752f;
  const float2 a0 = cadd(x[0], x[4]), a4 = csub(x[0], x[4]);
  const float2 a1 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]);
  const float2 a2 = cadd(x[2], x[6]), a6 = csub(x[2], x[6]);
  const float2 a3 = cadd(x[3], x[7]), a7 = csub(x[3], x[7]);
  // even outputs: 4-point DFT of a0..a3
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = mul_mi(csub(a1, a3));
  x[0] = cadd(b0, b1); x[4] = csub(b0, b1); x[2] = cadd(b2, b3); x[6] = csub(b2, b3);
  // odd outputs: 4-point DFT of a4, a5 W8, a6 W8^2, a7 W8^3   (W8 = (1 - i) / sqrt 2)
  const float2 c1 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));
  const float2 c2 = mul_mi(a6);
  const float2 c3 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));
  const float2 d0 = cadd(a4, c2), d2 = csub(a4, c2), d1 = cadd(c1, c3), d3 = mul_mi(csub(c1, c3));
  x[1] = cadd(d0, d1); x[5] = csub(d0, d1); x[3] = cadd(d2, d3); x[7] = csub(d2, d3);
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

// step 1.  ya / yb: the two frames' 400 pre-emphasised samples; tw[j] = exp(-2 pi i j / 512)
AVFE_HD void step1(int m, const float* ya, const float* yb, const float2* tw, float2* S) {
  float2 x[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int n = 64 * a + m;
    x[a] = (n < kFrame) ? make_float2(ya[n], yb[n]) : make_float2(0.0f, 0.0f);
  }
  dft8(x);
  S[m] = x[0];
#pragma unroll
  for (int r = 1; r < 8; ++r) S[r * kRow + m] = cmul(x[r], tw[(m * r) & (kNfft - 1)]);
}

// step 2 in two halves (all loads, barrier, all stores: the exchange is in place)
AVFE_HD void step2_load(int t, const float2* S, float2 (&x)[8]) {
  const int r = t >> 3, c = t & 7;
#pragma unroll
  for (int b = 0; b < 8; ++b) x[b] = S[r * kRow + 8 * b + c];
}
AVFE_HD void step2_store(int t, const float2* tw, float2 (&x)[8], float2* S) {
  const int r = t >> 3, c = t & 7;
  dft8(x);
  S[9 * c + r] = x[0];
#pragma unroll
  for (int u = 1; u < 8; ++u) S[u * kRow + 9 * c + r] = cmul(x[u], tw[(8 * c * u) & (kNfft - 1)]);
}

// step 3: spectrum in natural order
AVFE_HD void step3(int t, const float2* S, float2* C) {
  const int r = t & 7, u = t >> 3;                   // consecutive lanes write consecutive bins
  float2 x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = S[u * kRow + 9 * c + r];
  dft8(x);
#pragma unroll
  for (int v = 0; v < 8; ++v) C[t + 64 * v] = x[v];
}

// power spectra 1/512 |X|^2 of the two real frames packed in C: thread t owns bins t + 64 j
AVFE_HD void power_rows(int t, const float2* C, float* Pa, float* Pb) {
  const float scale = 0.25f / (float)kNfft;
#pragma unroll
  for (int jj = 0; jj < 5; ++jj) {
    const int k = t + 64 * jj;
    if (k > kNfft / 2) break;
    const float2 z = C[k], zm = C[(kNfft - k) & (kNfft - 1)];
    const float ar = z.x + zm.x, ai = z.y - zm.y, br = z.y + zm.y, bi = zm.x - z.x;
    Pa[k] = scale * (ar * ar + ai * ai);
    Pb[k] = scale * (br * br + bi * bi);
  }
}

// log of one filterbank energy: feat = fb_row . pspec, zero -> float64 eps (numpy.finfo(float).eps)
AVFE_HD float log_fbank(const float* P, const float* w, int lo, int hi) {
  float acc = 0.0f;
#pragma unroll 4
  for (int k = lo; k < hi; ++k) acc = fmaf(w[k], P[k], acc);
  if (acc == 0.0f) acc = 2.220446049250313e-16f;
  return logf(acc);
}

}  // namespace fbk
}  // namespace avfe
