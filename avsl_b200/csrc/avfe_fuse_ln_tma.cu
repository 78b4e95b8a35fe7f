// Fusion + transpose + LayerNorm with the tile fill done by tensor-map TMA (UTMALDG) instead of
// per-thread loads -- avsl/modules/av_hubert_encoder.py:315-330, same arithmetic as fuse_ln_kernel
// (avfe_fuse_ln.cu), which stays the path for feature maps whose rows are not 16-byte aligned.
//
// fuse_ln_kernel is bound by the SM's load/store pipe: every element is loaded from global memory in
// 32-byte row pieces, stored to shared memory, loaded again and stored to global memory.  Here a
// [256 channels x 16 bytes] box per TMA request lands the [C', TT] tile in shared memory (128-byte
// L2 promotion: the neighbouring tiles, read by other CTAs at the same time, hit), so the LSU only
// sees the transposed drain: two passes over the tile (moments, then normalise + store), each one
// 128-bit shared load per 16-byte channel row (consecutive lanes, consecutive rows: conflict-free)
// and, in the second pass, full-line stores of out[b, t, :].  16-byte rows keep a tile at 32 KB, so
// five CTAs share an SM and their TMA latency overlaps (32-byte rows, three CTAs: 69 % of the HBM
// peak instead of 7x %, profiles/r02).
// Contract: unit time stride, row pitch a multiple of 16 bytes (T = 750 in a [B, C, 752] allocation),
// C a multiple of 256.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "avfe_common.cuh"

namespace avfe {
namespace flt {

// tuning knobs (profiles/fuseln_sweep.sh): resident CTAs per SM the two kernels are compiled for
#ifndef AVFE_FLT_TILE_CTAS
#define AVFE_FLT_TILE_CTAS 4
#endif
#ifndef AVFE_FLT_RING_CTAS
#define AVFE_FLT_RING_CTAS 3
#endif
constexpr int kThreads = 256;
constexpr int kRowBytes = 16;                  // one channel row of a tile: 4 floats or 8 halves (six CTAs per SM)
constexpr int kBoxRows = 256;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  const uint8_t* mask;
  const float* gamma;
  const float* beta;
  void* out;
  float wa, wv, eps;
  int C, Cout, T, tiles_per_sample;
};

// the TT values of channel row r of one modality's tile: one 128-bit load; consecutive lanes read
// consecutive rows, i.e. a contiguous 512 bytes per warp (no swizzle needed)
template <typename T, int TT>
__device__ __forceinline__ void load_row(const uint8_t* tile, int r, float (&x)[TT]) {
  const uint4 q = *reinterpret_cast<const uint4*>(tile + r * kRowBytes);
  const T* a = reinterpret_cast<const T*>(&q);
#pragma unroll
  for (int i = 0; i < TT; ++i) x[i] = to_f32<T>(a[i]);
}

// fused value of channel c' (rounded to T like the forward's staging)
template <typename T, int MODE, int TT>
__device__ __forceinline__ void fused_row(const uint8_t* tile_a, const uint8_t* tile_v, int c, int C, bool has_a,
                                          bool has_v, float wa, float wv, float (&x)[TT]) {
  if (MODE == AVFE_FUSE_CONCAT) {
    const bool first = c < C;
    if (first ? has_a : has_v) load_row<T, TT>(first ? tile_a : tile_v, first ? c : c - C, x);
    else {
#pragma unroll
      for (int i = 0; i < TT; ++i) x[i] = 0.0f;
    }
    return;
  }
  float xa[TT], xv[TT];
  if (has_a) load_row<T, TT>(tile_a, c, xa);
  if (has_v) load_row<T, TT>(tile_v, c, xv);
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    const float p = has_a ? xa[i] : 0.0f, q = has_v ? xv[i] : 0.0f;
    const float s = (MODE == AVFE_FUSE_SUM) ? __fadd_rn(p, q) : __fadd_rn(__fmul_rn(wa, p), __fmul_rn(wv, q));
    x[i] = to_f32<T>(from_f32<T>(s));
  }
}

// The 16-bit types: one tile per CTA, four CTAs per SM (their tiles carry half the bytes for the same
// arithmetic: they need the warps more than a second stage -- 59 % of the HBM peak like this, 48 % with the
// two-stage ring below at three CTAs per SM).  CPL = channels per lane: 2 (a lane stores a pair = 4 bytes).
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, AVFE_FLT_TILE_CTAS)
fuse_ln_tma_tile_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_v, const Args a) {
  constexpr int TT = kRowBytes / (int)sizeof(T);
  constexpr int CPL = (sizeof(T) == 4) ? 1 : 2;
  extern __shared__ __align__(1024) uint8_t flt_smem[];
  uint8_t* tiles = flt_smem + ((1024u - (smem_u32(flt_smem) & 1023u)) & 1023u);
  uint8_t* tile_a = tiles;
  uint8_t* tile_v = tiles + (size_t)a.C * kRowBytes;
  float* red = reinterpret_cast<float*>(tiles + 2 * (size_t)a.C * kRowBytes);   // [8 warps][2 TT], mean[TT], rstd[TT], K[TT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 8 * 2 * TT + 3 * TT + (TT & 1));   // one per 256 fused channels
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t b = blockIdx.x / a.tiles_per_sample;
  const int t0 = (int)(blockIdx.x % a.tiles_per_sample) * TT;
  const int nt = min(TT, a.T - t0);
  unsigned m = 3u;
  if (a.mask != nullptr) m = (a.mask[2 * b] ? 1u : 0u) | (a.mask[2 * b + 1] ? 2u : 0u);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;

  // One barrier per block of 256 fused channels, so that the moments pass can start on the first
  // boxes while the later ones are still in flight.  concat: block q is one box (of fa or fv);
  // add / weighted sum: block q is box q of both maps.
  const int n_blocks = a.Cout / kBoxRows;
  if (tid == 0) {
    for (int q = 0; q < n_blocks; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + q)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int q = 0; q < n_blocks; ++q) {
      const bool is_v = (MODE == AVFE_FUSE_CONCAT) && q * kBoxRows >= a.C;
      const int c0 = is_v ? q * kBoxRows - a.C : q * kBoxRows;
      const bool ld_a = (MODE == AVFE_FUSE_CONCAT) ? (!is_v && has_a) : has_a;
      const bool ld_v = (MODE == AVFE_FUSE_CONCAT) ? (is_v && has_v) : has_v;
      const unsigned bytes = (unsigned)((ld_a ? 1 : 0) + (ld_v ? 1 : 0)) * kBoxRows * kRowBytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + q)), "r"(bytes) : "memory");
      if (ld_a)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile_a + (size_t)c0 * kRowBytes)), "l"(&map_a), "r"(smem_u32(bar + q)), "r"(t0), "r"(c0), "r"((int)b) : "memory");
      if (ld_v)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile_v + (size_t)c0 * kRowBytes)), "l"(&map_v), "r"(smem_u32(bar + q)), "r"(t0), "r"(c0), "r"((int)b) : "memory");
    }
  }
  // LayerNorm weight / bias of this thread's channels: fetched while the tile is in flight
  constexpr int kIter = 8;                              // channel groups per thread (C' <= 2048 * CPL)
  float g[kIter][CPL], be[kIter][CPL];
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int c = (tid + k * kThreads) * CPL;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      g[k][u] = (a.gamma && c + u < a.Cout) ? a.gamma[c + u] : 1.0f;
      be[k][u] = (a.beta && c + u < a.Cout) ? a.beta[c + u] : 0.0f;
    }
  }
  __syncthreads();                                      // barriers initialised before anybody polls them
  auto wait_block = [&](int q) {
    asm volatile(
        "{\n.reg .pred p;\nFLT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra FLT_WAIT;\n}" ::"r"(smem_u32(bar + q)), "r"(0) : "memory");
  };

  // ---- pass 1: shifted moments per time step (K = fused channel 0), this thread's channels
  float K[TT], s1[TT], s2[TT];
  wait_block(0);
  fused_row<T, MODE, TT>(tile_a, tile_v, 0, a.C, has_a, has_v, a.wa, a.wv, K);
#pragma unroll
  for (int i = 0; i < TT; ++i) { s1[i] = 0.0f; s2[i] = 0.0f; }
  for (int c = tid * CPL; c < a.Cout; c += kThreads * CPL) {
    // this round's channels are blocks [c0, c0 + CPL) for every thread of the CTA
    const int q0 = (c - tid * CPL) / kBoxRows;
#pragma unroll
    for (int u = 0; u < CPL; ++u)
      if (q0 + u < n_blocks) wait_block(q0 + u);
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      float x[TT];
      fused_row<T, MODE, TT>(tile_a, tile_v, c + u, a.C, has_a, has_v, a.wa, a.wv, x);
#pragma unroll
      for (int i = 0; i < TT; i += 2) {                 // packed fp32x2 (FADD2 / FFMA2), the scalar roundings
        const float2 d = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(-K[i], -K[i + 1]));
        const float2 t1 = __fadd2_rn(make_float2(s1[i], s1[i + 1]), d);
        const float2 t2 = __ffma2_rn(d, d, make_float2(s2[i], s2[i + 1]));
        s1[i] = t1.x; s1[i + 1] = t1.y; s2[i] = t2.x; s2[i + 1] = t2.y;
      }
    }
  }
  // warp reduction by halving: after step k a lane keeps 2 TT / 2^k of the (s1 | s2) vector
  float v[2 * TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) { v[i] = s1[i]; v[TT + i] = s2[i]; }
  int idx = 0;                                          // first element index this lane still owns
  constexpr int kFold = 32 / (2 * TT);                  // lanes that end up with the same element
#pragma unroll
  for (int half = TT, o = 16; half >= 1; half >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;                 // upper lanes keep the upper half of the vector
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[half + i];
      const float recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[i] = (upper ? v[half + i] : v[i]) + recv;
    }
    idx += upper ? half : 0;
  }
#pragma unroll
  for (int o = kFold / 2; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  if ((lane & (kFold - 1)) == 0) red[wid * 2 * TT + idx] = v[0];
  float* mean_s = red + 8 * 2 * TT;
  float* rstd_s = mean_s + TT;
  float* kref_s = rstd_s + TT;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < TT; ++i) kref_s[i] = K[i];
  }
  __syncthreads();
  if (tid < TT) {
    float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t1 += red[w * 2 * TT + tid]; t2 += red[w * 2 * TT + TT + tid]; }
    const float n = (float)a.Cout;
    const float md = t1 / n;
    mean_s[tid] = kref_s[tid] + md;
    rstd_s[tid] = rsqrtf(fmaxf((t2 - t1 * md) / n, 0.0f) + a.eps);
  }
  __syncthreads();

  // ---- pass 2: normalise and store; a warp writes 32 * CPL consecutive channels of one time step
  float mu[TT], rs[TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) { mu[i] = mean_s[i]; rs[i] = rstd_s[i]; }
  T* out = static_cast<T*>(a.out) + (b * (int64_t)a.T + t0) * a.Cout;
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int c = (tid + k * kThreads) * CPL;
    if (c >= a.Cout) break;
    float x[CPL][TT];
#pragma unroll
    for (int u = 0; u < CPL; ++u) fused_row<T, MODE, TT>(tile_a, tile_v, c + u, a.C, has_a, has_v, a.wa, a.wv, x[u]);
    // (x - mean) * rstd as packed fp32x2 over pairs of time steps, then the affine per channel
#pragma unroll
    for (int u = 0; u < CPL; ++u)
#pragma unroll
      for (int i = 0; i < TT; i += 2) {
        const float2 h = __fmul2_rn(__fadd2_rn(make_float2(x[u][i], x[u][i + 1]), make_float2(-mu[i], -mu[i + 1])),
                                    make_float2(rs[i], rs[i + 1]));
        x[u][i] = fmaf(h.x, g[k][u], be[k][u]);
        x[u][i + 1] = fmaf(h.y, g[k][u], be[k][u]);
      }
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      if (i < nt) {
        T* o = out + (int64_t)i * a.Cout + c;
        if (CPL == 1) {
          o[0] = from_f32<T>(x[0][i]);
        } else {
          T pair[2];
          pair[0] = from_f32<T>(x[0][i]);
          pair[1] = from_f32<T>(x[CPL - 1][i]);
          *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(pair);
        }
      }
    }
  }
}

// CPL = channels per lane: 1 for float (a warp stores 32 floats = one line), 2 for the 16-bit types
// (a lane stores a pair = 4 bytes).
// Persistent CTAs with a two-deep tile ring: the TMA boxes of tile i+1 are requested before tile i is
// touched, so a CTA hides its own TMA latency (before: one tile per CTA, latency hidden only by the
// co-resident CTAs), and the barriers, the LayerNorm weights (32 registers) and the launch of 6,000-12,000
// CTAs are paid once per CTA instead of once per 32 KB tile.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, AVFE_FLT_RING_CTAS)
fuse_ln_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_v, const Args a,
                   const int n_tiles) {
  constexpr int TT = kRowBytes / (int)sizeof(T);
  constexpr int CPL = (sizeof(T) == 4) ? 1 : 2;
  extern __shared__ __align__(1024) uint8_t flt_smem[];
  uint8_t* tiles = flt_smem + ((1024u - (smem_u32(flt_smem) & 1023u)) & 1023u);
  const size_t stage_bytes = 2 * (size_t)a.C * kRowBytes;                         // fa tile | fv tile
  float* red = reinterpret_cast<float*>(tiles + 2 * stage_bytes);                 // [8 warps][2 TT], mean[TT], rstd[TT], K[TT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 8 * 2 * TT + 3 * TT + (TT & 1));   // [2 stages][16]: one per 256 fused channels
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_blocks = a.Cout / kBoxRows;

  auto sample_flags = [&](int64_t b) -> unsigned {
    return a.mask == nullptr ? 3u : ((a.mask[2 * b] ? 1u : 0u) | (a.mask[2 * b + 1] ? 2u : 0u));
  };
  // One barrier per block of 256 fused channels, so that the moments pass can start on the first
  // boxes while the later ones are still in flight.  concat: block q is one box (of fa or fv);
  // add / weighted sum: block q is box q of both maps.
  auto issue = [&](int tile, int st) {                  // thread 0 only
    const int64_t b = tile / a.tiles_per_sample;
    const int t0 = (tile % a.tiles_per_sample) * TT;
    const unsigned m = sample_flags(b);
    const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
    uint8_t* tile_a = tiles + st * stage_bytes;
    uint8_t* tile_v = tile_a + (size_t)a.C * kRowBytes;
    // the stage was read through the generic proxy one iteration ago: order those reads before the async writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int q = 0; q < n_blocks; ++q) {
      const bool is_v = (MODE == AVFE_FUSE_CONCAT) && q * kBoxRows >= a.C;
      const int c0 = is_v ? q * kBoxRows - a.C : q * kBoxRows;
      const bool ld_a = (MODE == AVFE_FUSE_CONCAT) ? (!is_v && has_a) : has_a;
      const bool ld_v = (MODE == AVFE_FUSE_CONCAT) ? (is_v && has_v) : has_v;
      const unsigned bytes = (unsigned)((ld_a ? 1 : 0) + (ld_v ? 1 : 0)) * kBoxRows * kRowBytes;
      uint64_t* bq = bar + st * 16 + q;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bq)), "r"(bytes) : "memory");
      if (ld_a)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile_a + (size_t)c0 * kRowBytes)), "l"(&map_a), "r"(smem_u32(bq)), "r"(t0), "r"(c0), "r"((int)b) : "memory");
      if (ld_v)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile_v + (size_t)c0 * kRowBytes)), "l"(&map_v), "r"(smem_u32(bq)), "r"(t0), "r"(c0), "r"((int)b) : "memory");
    }
  };
  if (tid == 0) {
    for (int q = 0; q < 2 * 16; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + q)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < n_tiles) issue((int)blockIdx.x, 0);
  }
  // LayerNorm weight / bias of this thread's channels: fetched once, while the first tile is in flight
  constexpr int kIter = 8;                              // channel groups per thread (C' <= 2048 * CPL)
  float g[kIter][CPL], be[kIter][CPL];
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int c = (tid + k * kThreads) * CPL;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      g[k][u] = (a.gamma && c + u < a.Cout) ? a.gamma[c + u] : 1.0f;
      be[k][u] = (a.beta && c + u < a.Cout) ? a.beta[c + u] : 0.0f;
    }
  }
  __syncthreads();                                      // barriers initialised before anybody polls them

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
  const int st = it & 1;
  const unsigned parity = (unsigned)(it >> 1) & 1u;
  if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + (int)gridDim.x, st ^ 1);   // the other stage was drained one iteration ago
  uint8_t* tile_a = tiles + st * stage_bytes;
  uint8_t* tile_v = tile_a + (size_t)a.C * kRowBytes;
  const int64_t b = tile / a.tiles_per_sample;
  const int t0 = (tile % a.tiles_per_sample) * TT;
  const int nt = min(TT, a.T - t0);
  const unsigned m = sample_flags(b);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
  auto wait_block = [&](int q) {
    asm volatile(
        "{\n.reg .pred p;\nFLT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra FLT_WAIT;\n}" ::"r"(smem_u32(bar + st * 16 + q)), "r"(parity) : "memory");
  };

  // ---- pass 1: shifted moments per time step (K = fused channel 0), this thread's channels
  float K[TT], s1[TT], s2[TT];
  wait_block(0);
  fused_row<T, MODE, TT>(tile_a, tile_v, 0, a.C, has_a, has_v, a.wa, a.wv, K);
#pragma unroll
  for (int i = 0; i < TT; ++i) { s1[i] = 0.0f; s2[i] = 0.0f; }
  for (int c = tid * CPL; c < a.Cout; c += kThreads * CPL) {
    // this round's channels are blocks [c0, c0 + CPL) for every thread of the CTA
    const int q0 = (c - tid * CPL) / kBoxRows;
#pragma unroll
    for (int u = 0; u < CPL; ++u)
      if (q0 + u < n_blocks) wait_block(q0 + u);
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      float x[TT];
      fused_row<T, MODE, TT>(tile_a, tile_v, c + u, a.C, has_a, has_v, a.wa, a.wv, x);
#pragma unroll
      for (int i = 0; i < TT; i += 2) {                 // packed fp32x2 (FADD2 / FFMA2), the scalar roundings
        const float2 d = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(-K[i], -K[i + 1]));
        const float2 t1 = __fadd2_rn(make_float2(s1[i], s1[i + 1]), d);
        const float2 t2 = __ffma2_rn(d, d, make_float2(s2[i], s2[i + 1]));
        s1[i] = t1.x; s1[i + 1] = t1.y; s2[i] = t2.x; s2[i + 1] = t2.y;
      }
    }
  }
  // warp reduction by halving: after step k a lane keeps 2 TT / 2^k of the (s1 | s2) vector
  float v[2 * TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) { v[i] = s1[i]; v[TT + i] = s2[i]; }
  int idx = 0;                                          // first element index this lane still owns
  constexpr int kFold = 32 / (2 * TT);                  // lanes that end up with the same element
#pragma unroll
  for (int half = TT, o = 16; half >= 1; half >>= 1, o >>= 1) {
    const bool upper = (lane & o) != 0;                 // upper lanes keep the upper half of the vector
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[half + i];
      const float recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[i] = (upper ? v[half + i] : v[i]) + recv;
    }
    idx += upper ? half : 0;
  }
#pragma unroll
  for (int o = kFold / 2; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  if ((lane & (kFold - 1)) == 0) red[wid * 2 * TT + idx] = v[0];
  float* mean_s = red + 8 * 2 * TT;
  float* rstd_s = mean_s + TT;
  float* kref_s = rstd_s + TT;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < TT; ++i) kref_s[i] = K[i];
  }
  __syncthreads();
  if (tid < TT) {
    float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t1 += red[w * 2 * TT + tid]; t2 += red[w * 2 * TT + TT + tid]; }
    const float n = (float)a.Cout;
    const float md = t1 / n;
    mean_s[tid] = kref_s[tid] + md;
    rstd_s[tid] = rsqrtf(fmaxf((t2 - t1 * md) / n, 0.0f) + a.eps);
  }
  __syncthreads();

  // ---- pass 2: normalise and store; a warp writes 32 * CPL consecutive channels of one time step
  float mu[TT], rs[TT];
#pragma unroll
  for (int i = 0; i < TT; ++i) { mu[i] = mean_s[i]; rs[i] = rstd_s[i]; }
  T* out = static_cast<T*>(a.out) + (b * (int64_t)a.T + t0) * a.Cout;
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int c = (tid + k * kThreads) * CPL;
    if (c >= a.Cout) break;
    float x[CPL][TT];
#pragma unroll
    for (int u = 0; u < CPL; ++u) fused_row<T, MODE, TT>(tile_a, tile_v, c + u, a.C, has_a, has_v, a.wa, a.wv, x[u]);
    // (x - mean) * rstd as packed fp32x2 over pairs of time steps, then the affine per channel
#pragma unroll
    for (int u = 0; u < CPL; ++u)
#pragma unroll
      for (int i = 0; i < TT; i += 2) {
        const float2 h = __fmul2_rn(__fadd2_rn(make_float2(x[u][i], x[u][i + 1]), make_float2(-mu[i], -mu[i + 1])),
                                    make_float2(rs[i], rs[i + 1]));
        x[u][i] = fmaf(h.x, g[k][u], be[k][u]);
        x[u][i + 1] = fmaf(h.y, g[k][u], be[k][u]);
      }
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      if (i < nt) {
        T* o = out + (int64_t)i * a.Cout + c;
        if (CPL == 1) {
          o[0] = from_f32<T>(x[0][i]);
        } else {
          T pair[2];
          pair[0] = from_f32<T>(x[0][i]);
          pair[1] = from_f32<T>(x[CPL - 1][i]);
          *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(pair);
        }
      }
    }
  }
  __syncthreads();                                      // the stage (and red / mean / rstd) may be reused
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <typename T, int MODE>
static int launch(const CUtensorMap& ma, const CUtensorMap& mv, const Args& a, int64_t B, cudaStream_t s) {
  constexpr int TT = kRowBytes / (int)sizeof(T);
  constexpr bool kRing = sizeof(T) == 4;                 // float: persistent CTAs with a two-stage ring
  const size_t tile_bytes = 2 * (size_t)a.C * kRowBytes;
  const size_t smem = (kRing ? 2 : 1) * tile_bytes + (8 * 2 * TT + 3 * TT + 2) * sizeof(float) + 2 * 16 * 8 + 16 + 1024;
  const int64_t n_tiles = B * a.tiles_per_sample;
  if (n_tiles > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  if constexpr (kRing) {
    if (cudaFuncSetAttribute(fuse_ln_tma_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
    int resident = 0;                                    // one resident wave of CTAs striding over the tiles
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fuse_ln_tma_kernel<T, MODE>, kThreads, smem) != cudaSuccess || resident < 1) {
      cudaGetLastError();
      resident = 1;
    }
    int64_t ctas = (int64_t)resident * kNumSMs;
    if (ctas > n_tiles) ctas = n_tiles;
    fuse_ln_tma_kernel<T, MODE><<<(unsigned)ctas, kThreads, smem, s>>>(ma, mv, a, (int)n_tiles);
  } else {
    if (cudaFuncSetAttribute(fuse_ln_tma_tile_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
    fuse_ln_tma_tile_kernel<T, MODE><<<(unsigned)n_tiles, kThreads, smem, s>>>(ma, mv, a);
  }
  count_launch();
  return check_launch();
}

template <typename T>
static int pick_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mv, const Args& a, int64_t B, cudaStream_t s) {
  switch (mode) {
    case AVFE_FUSE_CONCAT: return launch<T, AVFE_FUSE_CONCAT>(ma, mv, a, B, s);
    case AVFE_FUSE_SUM: return launch<T, AVFE_FUSE_SUM>(ma, mv, a, B, s);
    case AVFE_FUSE_WSUM: return launch<T, AVFE_FUSE_WSUM>(ma, mv, a, B, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}

}  // namespace flt
}  // namespace avfe

using namespace avfe;

// 1 if avfe_fuse_layernorm_pitched can take the TMA path for this layout
extern "C" int avfe_fuse_layernorm_tma_ok(int dtype, int64_t C, int64_t T, int64_t t_pitch) {
  const int64_t esz = (dtype == AVFE_F32) ? 4 : 2;
  return (dtype == AVFE_F32 || dtype == AVFE_F16 || dtype == AVFE_BF16) && C > 0 && (C % flt::kBoxRows) == 0 &&
         t_pitch >= T && (t_pitch * esz) % 16 == 0 && 2 * C <= 2048 * (esz == 4 ? 1 : 2);
}

extern "C" int avfe_fuse_layernorm_pitched(const void* fa, const void* fv, const uint8_t* mask, int mode, float w_a,
                                           float w_v, int dtype, int64_t B, int64_t C, int64_t T, int64_t t_pitch,
                                           const float* gamma, const float* beta, float eps, void* out,
                                           avfe_stream_t stream) {
  if (B < 0 || C < 0 || T < 0 || !(eps >= 0.0f)) return AVFE_ERR_INVALID_ARG;
  if (mode != AVFE_FUSE_CONCAT && mode != AVFE_FUSE_SUM && mode != AVFE_FUSE_WSUM) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || C == 0 || T == 0) return AVFE_OK;
  if (!fa || !fv || !out) return AVFE_ERR_INVALID_ARG;
  if (!avfe_fuse_layernorm_tma_ok(dtype, C, T, t_pitch)) {
    if (t_pitch == T) return avfe_fuse_layernorm(fa, fv, mask, mode, w_a, w_v, dtype, B, C, T, gamma, beta, eps, out, stream);
    return AVFE_ERR_UNSUPPORTED;                        // padded rows that TMA cannot address
  }
  if (!aligned16(fa) || !aligned16(fv) || (reinterpret_cast<uintptr_t>(out) & 3u)) return AVFE_ERR_ALIGNMENT;
  if (B > 0x7fffffffLL || T > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  flt::EncodeTiledFn enc = flt::encode_fn();
  if (!enc) return AVFE_ERR_CUDA;
  const int esz = (dtype == AVFE_F32) ? 4 : 2;
  const CUtensorMapDataType dt = (dtype == AVFE_F32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : (dtype == AVFE_F16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ma, mv;
  const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[2] = {(cuuint64_t)t_pitch * esz, (cuuint64_t)C * t_pitch * esz};
  const cuuint32_t box[3] = {(cuuint32_t)(flt::kRowBytes / esz), (cuuint32_t)flt::kBoxRows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&ma, dt, 3, const_cast<void*>(fa), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&mv, dt, 3, const_cast<void*>(fv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return AVFE_ERR_CUDA;
  flt::Args a;
  a.mask = mask; a.gamma = gamma; a.beta = beta; a.out = out; a.wa = w_a; a.wv = w_v; a.eps = eps;
  a.C = (int)C; a.Cout = (int)(mode == AVFE_FUSE_CONCAT ? 2 * C : C); a.T = (int)T;
  const int TT = flt::kRowBytes / esz;
  a.tiles_per_sample = (int)((T + TT - 1) / TT);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case AVFE_F32: return flt::pick_mode<float>(mode, ma, mv, a, B, s);
    case AVFE_F16: return flt::pick_mode<__half>(mode, ma, mv, a, B, s);
    default: return flt::pick_mode<__nv_bfloat16>(mode, ma, mv, a, B, s);
  }
}
