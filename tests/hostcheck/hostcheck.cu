// TEST-ONLY host harness.  Steps the __host__ __device__ codelets of libavfe
// (avfe_logmel_core.cuh, avfe_lip_math.cuh) thread by thread on the CPU so the index algebra
// (Good-Thomas maps, 20x20 exchange layout, two-frames-per-FFT untangling, reflect padding,
// similarity fit, bilinear blend order) can be checked against the oracle in the no-GPU test
// tier.  It is never loaded by avsl_b200 and is not a CPU fallback: the product path only
// calls libavfe.so kernels.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "avfe_lip_math.cuh"
#include "avfe_logfbank_core.cuh"
#include "avfe_logmel_core.cuh"
#include "avfe_noise_core.cuh"

using namespace avfe;

extern "C" {

// raw log10-mel of one tile (32 frames starting at t0) -> out[n_mels][32]
void hc_logmel_tile(const float* clip, int64_t L, int64_t Lp, int64_t t0, int n_mels,
                    const float* fb, float* out) {
  using namespace lm;
  std::vector<float> hann(kNfft);
  std::vector<float2> tw(kNfft);
  for (int i = 0; i < kNfft; ++i) hann[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / kNfft));
  for (int k1 = 0; k1 < 20; ++k1)
    for (int j = 0; j < 20; ++j) {
      const double a = -2.0 * M_PI * ((j * k1) % kNfft) / kNfft;
      tw[k1 * 20 + j] = make_float2((float)cos(a), (float)sin(a));
    }
  std::vector<float> buf(kTileFloats);
  std::vector<float2> Z(kPairs * kZPair);
  std::vector<float> pa(kThreads * kBinsPerThread), pb(kThreads * kBinsPerThread);
  for (int i = 0; i < kTileSamples; ++i) buf[tile_pos(i)] = padded_sample(clip, L, Lp, t0 * kHop + i);
  for (int tid = 0; tid < kThreads; ++tid) stage1(tid / 20, tid % 20, buf.data(), hann.data(), tw.data(), Z.data());
  // stage 2 keeps its bins in "registers" (one array per thread) and stores only the upper half;
  // barrier; untangle from registers + the partner's row; barrier; power rows overwrite the Z slab
  std::vector<float2> regs(kThreads * 20);
  for (int tid = 0; tid < kThreads; ++tid)
    stage2_keep(tid / 20, tid % 20, Z.data(), *reinterpret_cast<float2(*)[20]>(&regs[tid * 20]));
  for (int tid = 0; tid < kThreads; ++tid)
    split_from_regs(tid / 20, tid % 20, Z.data(), *reinterpret_cast<const float2(*)[20]>(&regs[tid * 20]),
                    *reinterpret_cast<float(*)[kBinsPerThread]>(&pa[tid * kBinsPerThread]),
                    *reinterpret_cast<float(*)[kBinsPerThread]>(&pb[tid * kBinsPerThread]));
  float* P = reinterpret_cast<float*>(Z.data());
  for (int tid = 0; tid < kThreads; ++tid)
    split_store(tid / 20, tid % 20, P, *reinterpret_cast<float(*)[kBinsPerThread]>(&pa[tid * kBinsPerThread]),
                *reinterpret_cast<float(*)[kBinsPerThread]>(&pb[tid * kBinsPerThread]));
  for (int m = 0; m < n_mels; ++m) {
    int lo = kBins, hi = 0;
    for (int k = 0; k < kBins; ++k)
      if (fb[m * kBins + k] != 0.0f) { lo = lo < k ? lo : k; hi = k + 1; }
    if (lo >= hi) { lo = 0; hi = 0; }
    {
      // same packed form the kernel uses: weights zero-padded to quads
      const int n4 = (hi - lo + 3) >> 2;
      alignas(16) float wq[208] = {0};
      for (int k = lo; k < hi; ++k) wq[k - lo] = fb[m * kBins + k];
      for (int f = 0; f < kTileFrames; ++f)
        out[m * kTileFrames + f] = mel_log10_quads(P + prow_offset(f) + lo, reinterpret_cast<const float4*>(wq), n4);
    }
  }
}

// raw log filterbank energies of frames fa, fb (two frames of one clip) -> out[2][nfilt]: the warp's
// 16 x 32 transform stepped lane by lane
void hc_logfbank_pair(const float* clip, int64_t len, int64_t fa, int64_t fb, int nfilt, const float* fbank,
                      float* out) {
  using namespace fbk;
  std::vector<float2> tw1(16 * 32);
  auto w512 = [](int e) { return make_float2((float)cos(-2.0 * M_PI * e / kNfft), (float)sin(-2.0 * M_PI * e / kNfft)); };
  for (int r = 0; r < 16; ++r)
    for (int l = 0; l < 32; ++l) tw1[r * 32 + l] = w512(tw1_index(r, l));
  std::vector<float> ya(kFrame), yb(kFrame);
  for (int n = 0; n < kFrame; ++n) {
    ya[n] = preemph_sample(clip, len, fa * kHop + n);
    yb[n] = preemph_sample(clip, len, fb * kHop + n);
  }
  std::vector<float> slot(kSlotFloats);
  float2* S = reinterpret_cast<float2*>(slot.data());
  float2* U = reinterpret_cast<float2*>(slot.data() + kUOffset);
  for (int l = 0; l < 32; ++l) fft_stage1(l, ya.data(), yb.data(), tw1.data(), S);
  std::vector<float2> regs(32 * 16);
  for (int l = 0; l < 32; ++l) fft_stage2(l, S, *reinterpret_cast<float2(*)[16]>(&regs[l * 16]));
  for (int l = 0; l < 32; ++l) fft_upper_store(l, *reinterpret_cast<float2(*)[16]>(&regs[l * 16]), U);
  for (int l = 0; l < 32; ++l)
    fft_power(l, *reinterpret_cast<float2(*)[16]>(&regs[l * 16]), U, slot.data(), slot.data() + kPRow);
  for (int f = 0; f < 2; ++f)
    for (int m = 0; m < nfilt; ++m) {
      int lo = kBins, hi = 0;
      for (int k = 0; k < kBins; ++k)
        if (fbank[m * kBins + k] != 0.0f) { lo = lo < k ? lo : k; hi = k + 1; }
      if (lo >= hi) { lo = 0; hi = 0; }
      out[f * nfilt + m] = log_fbank(slot.data() + f * kPRow, fbank + m * kBins, lo, hi);
    }
}

void hc_dft16(const float* in_ri, float* out_ri) {
  float2 x[16];
  for (int i = 0; i < 16; ++i) x[i] = make_float2(in_ri[2 * i], in_ri[2 * i + 1]);
  fbk::dft16(x);
  for (int i = 0; i < 16; ++i) { out_ri[2 * i] = x[i].x; out_ri[2 * i + 1] = x[i].y; }
}

void hc_dft20(const float* in_ri, float* out_ri) {
  float2 x[20];
  for (int i = 0; i < 20; ++i) x[i] = make_float2(in_ri[2 * i], in_ri[2 * i + 1]);
  lm::dft20(x);
  for (int i = 0; i < 20; ++i) { out_ri[2 * i] = x[i].x; out_ri[2 * i + 1] = x[i].y; }
}

int hc_float_key(float f) { return lm::float_key(f); }
float hc_key_float(int k) { return lm::key_float(k); }

void hc_similarity_fit(const double* src, const double* dst, int n, double* fwd6, double* inv6) {
  similarity_fit(reinterpret_cast<const double(*)[2]>(src), reinterpret_cast<const double(*)[2]>(dst), n, fwd6);
  affine_inverse(fwd6, inv6);
}

void hc_crop_origin(double cx, double cy, int half, int std_size, int* rc) {
  crop_origin(cx, cy, half, half, std_size, std_size, &rc[0], &rc[1]);
}

// warp window [r0, r0+h) x [c0, c0+w) of the std frame with inverse matrix rows inv6
void hc_warp_window(const uint8_t* gray, int H, int W, const double* inv6, int r0, int c0, int h,
                    int w, uint8_t* out) {
  double lut[256];
  for (int k = 0; k < 256; ++k) lut[k] = f64div((double)k, 255.0);
  auto tap = [&](int r, int c) -> double { return lut[gray[(size_t)r * W + c]]; };
  for (int pr = 0; pr < h; ++pr)
    for (int pc = 0; pc < w; ++pc) {
      const double tfr = (double)(r0 + pr), tfc = (double)(c0 + pc);
      const double sc = f64add(f64add(f64mul(inv6[0], tfc), f64mul(inv6[1], tfr)), inv6[2]);
      const double sr = f64add(f64add(f64mul(inv6[3], tfc), f64mul(inv6[4], tfr)), inv6[5]);
      out[pr * w + pc] = bilinear_u8(sr, sc, H, W, tap);
    }
}

void hc_gray(const uint8_t* bgr, int64_t n, uint8_t* gray) {
  for (int64_t i = 0; i < n; ++i)
    gray[i] = (uint8_t)gray_from_bgr(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}

// The sum of squares of x[(i) % period], i < n, evaluated the way noise_leaf_kernel +
// noise_combine_kernel do it: heap-addressed tree, 8 "lanes" per leaf joined by xor steps 1, 2, 4,
// a sequential tail, children added bottom-up.  Returns the root; *leaves counts the leaf nodes.
float hc_noise_tree_sumsq(const float* x, uint32_t n, uint32_t period, int64_t max_len, int* leaves) {
  const int depth = tree_depth(max_len);
  const uint32_t slots = 2u << depth;
  std::vector<float> heap(slots, NAN);
  *leaves = 0;
  for (uint32_t pos = 0; pos < n; pos += 64) {          // one 8-lane group per multiple of 64
    uint32_t off, len;
    const uint32_t k = locate_piece(pos, n, off, len);
    if (pos - off >= 64u) continue;                       // the piece belongs to the group before
    if (k >= slots || locate_node(k, n, off, len) != 1 || !isnan(heap[k])) return NAN;
    ++*leaves;
    float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t body = len & ~7u;
    auto sq = [&](uint32_t i) { const float v = x[(off + i) % period]; return v * v; };
    if (len >= 8)
      for (int j = 0; j < 8; ++j) {
        r[j] = sq(j);
        for (uint32_t i = 8; i < body; i += 8) r[j] = r[j] + sq(i + j);
      }
    for (int step = 1; step < 8; step <<= 1) {
      float t[8];
      for (int j = 0; j < 8; ++j) t[j] = r[j] + r[j ^ step];
      memcpy(r, t, sizeof(r));
    }
    float res = len < 8 ? 0.f : r[0];
    for (uint32_t i = (len < 8 ? 0u : body); i < len; ++i) res = res + sq(i);
    heap[k] = res;
  }
  for (int d = depth - 1; d >= 0; --d)
    for (uint32_t k = 1u << d; k < (2u << d); ++k) {
      uint32_t off, len;
      if (locate_node(k, n, off, len) == 2) heap[k] = heap[2 * k] + heap[2 * k + 1];
    }
  return heap[1];
}

}  // extern "C"
