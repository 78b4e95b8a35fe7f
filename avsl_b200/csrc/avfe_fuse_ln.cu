// Fusion + transpose + LayerNorm in one pass: the tail of AVHuBERTEncoderWrapper.forward,
// avsl/modules/av_hubert_encoder.py:315-330
//     features = cat([fa, fv], 1) | fa + fv            (:315-326, missing modality zero-filled)
//     features = features.transpose(1, 2)              (:329)
//     features = self.layer_norm(features)             (:330; LayerNorm = nn.LayerNorm evaluated in
//                                                       float32 and cast back, av_hubert_layers.py:438-440)
// [B, C, T] x 2  ->  [B, T, C'] with C' = 2C (concat) or C (sum / weighted sum).
//
// HBM-bound: the unfused chain writes and re-reads the fused tensor twice (cat, transpose copy,
// LayerNorm); here every input byte is read once and every output byte written once.  A CTA owns
// a tile of TT consecutive time steps of one sample, all C' channels: rows are read along T
// (TT elements = 32 or 64 contiguous bytes per channel row), parked in shared memory in the input
// dtype (row stride padded to an odd number of words so that both the row-wise fill and the
// column-wise drain are bank-conflict free) and drained channel-contiguous: a warp writes 32
// consecutive channels of one time step, i.e. full 128-byte lines.  The kernel is bound by the
// load/store pipe of the SM (every element crosses shared memory), so the moments are NOT taken
// from the tile: each thread accumulates sum(x - K) and sum((x - K)^2) of the values it fills,
// K = the time step's first channel (shifted moments: no cancellation), and the partials meet
// through a few shuffles and one shared-memory round.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "avfe_common.cuh"

namespace avfe {
namespace fln {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

constexpr int kThreads = 512;

// elements of padding per tile row: the row stride in 32-bit words must be odd
template <typename T> struct Pad { static constexpr int value = 4 / sizeof(T); };

template <typename T, int TT>
__host__ __device__ constexpr int row_stride() { return TT + Pad<T>::value; }

struct Args {
  const void* fa;
  const void* fv;
  const uint8_t* mask;      // [B,2] or nullptr
  const float* gamma;       // [C'] or nullptr (= 1)
  const float* beta;        // [C'] or nullptr (= 0)
  void* out;                // [B, T, C']
  float wa, wv, eps;
  int64_t B;
  int C, Cout, T, tiles_per_sample;
};

// VEC consecutive time steps (VEC * sizeof(T) = 8 bytes for float, 4 for the half types) as one load
template <typename T, int VEC> struct Pack;
template <> struct Pack<float, 2> { typedef float2 type; };
template <> struct Pack<float, 1> { typedef float type; };
template <> struct Pack<__half, 2> { typedef __half2 type; };
template <> struct Pack<__half, 1> { typedef __half type; };
template <> struct Pack<__nv_bfloat16, 2> { typedef __nv_bfloat162 type; };
template <> struct Pack<__nv_bfloat16, 1> { typedef __nv_bfloat16 type; };

// A tile touches 32 or 64 bytes of every channel row; the rest of the 128-byte line belongs to the
// neighbouring tiles, which other CTAs are reading at about the same time: ask L2 to fetch the
// whole line (ld.global.L2::128B) so that DRAM sees full-line bursts and the neighbours hit.
__device__ __forceinline__ float2 ld_line(const float2* p) {
  float2 v;
  asm volatile("ld.global.L2::128B.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_line(const float* p) {
  float v;
  asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_line_u32(const void* p) {
  unsigned v;
  asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned short ld_line_u16(const void* p) {
  unsigned short v;
  asm volatile("ld.global.L2::128B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ __half2 ld_line(const __half2* p) { const unsigned v = ld_line_u32(p); return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ __half ld_line(const __half* p) { const unsigned short v = ld_line_u16(p); return *reinterpret_cast<const __half*>(&v); }
__device__ __forceinline__ __nv_bfloat162 ld_line(const __nv_bfloat162* p) { const unsigned v = ld_line_u32(p); return *reinterpret_cast<const __nv_bfloat162*>(&v); }
__device__ __forceinline__ __nv_bfloat16 ld_line(const __nv_bfloat16* p) { const unsigned short v = ld_line_u16(p); return *reinterpret_cast<const __nv_bfloat16*>(&v); }

template <typename T, int MODE>
__device__ __forceinline__ T combine1(T x, T y, float wa, float wv) {
  const float a = to_f32<T>(x), v = to_f32<T>(y);
  return from_f32<T>((MODE == AVFE_FUSE_SUM) ? __fadd_rn(a, v) : __fadd_rn(__fmul_rn(wa, a), __fmul_rn(wv, v)));
}

// One CTA = one tile: time steps [t0, t0 + TT) of sample b, all Cout channels.
template <typename T, int MODE, int TT, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
fuse_ln_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char fln_smem[];
  typedef typename Pack<T, VEC>::type P;
  constexpr int RS = row_stride<T, TT>();
  T* tile = reinterpret_cast<T*>(fln_smem);                       // [Cout][RS]
  float* stat = reinterpret_cast<float*>(fln_smem + (((size_t)a.Cout * RS * sizeof(T) + 15) & ~(size_t)15));
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t b = blockIdx.x / a.tiles_per_sample;
  const int t0 = (int)(blockIdx.x % a.tiles_per_sample) * TT;
  const int nt = min(TT, a.T - t0);
  unsigned m = 3u;
  if (a.mask != nullptr) m = (a.mask[2 * b] ? 1u : 0u) | (a.mask[2 * b + 1] ? 2u : 0u);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
  const T* fa = static_cast<const T*>(a.fa) + b * (int64_t)a.C * a.T + t0;
  const T* fv = static_cast<const T*>(a.fv) + b * (int64_t)a.C * a.T + t0;
  const T zero = from_f32<T>(0.0f);

  // ---- fill: thread = (row group, pack of VEC time steps); TT / VEC lanes cover one row's
  // contiguous bytes, U rows per thread in flight before anything is stored
  constexpr int kLanesPerRow = TT / VEC;
  constexpr int kRowsPerPass = kThreads / kLanesPerRow;
  const int tp = (tid % kLanesPerRow) * VEC, r0 = tid / kLanesPerRow;
  constexpr int U = 8;
  // with VEC == 2 (T even) a pack never straddles the end of the sample: tp + 1 < nt iff tp < nt
  auto load = [&](const T* row, bool live) -> P {
    if (live) return *reinterpret_cast<const P*>(row + tp);
    P z;
    T* zp = reinterpret_cast<T*>(&z);
#pragma unroll
    for (int i = 0; i < VEC; ++i) zp[i] = zero;
    return z;
  };
  // fused value of (channel cc, time steps tp .. tp + VEC - 1) from its loaded packs
  auto fused = [&](const P& xa, const P& xv, int i) -> T {
    const T* pa = reinterpret_cast<const T*>(&xa);
    const T* pv = reinterpret_cast<const T*>(&xv);
    return (MODE == AVFE_FUSE_CONCAT) ? pa[i] : combine1<T, MODE>(pa[i], pv[i], a.wa, a.wv);
  };
  float K[VEC], s1[VEC], s2[VEC];
  {
    const bool live = tp < nt;
    const P ka = load(fa, live && has_a);                // fused channel 0: fa row 0 (+ fv row 0 unless concat)
    const P kv = (MODE == AVFE_FUSE_CONCAT) ? ka : load(fv, live && has_v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { K[i] = to_f32<T>(fused(ka, kv, i)); s1[i] = 0.0f; s2[i] = 0.0f; }
  }
  // Row pointers advance incrementally (one 64-bit add per U rows).  "concat" is two plain
  // copies: fused channels [0, C) from fa, [C, 2C) from fv.
  const bool t_live = tp < nt;
  const int64_t pass_stride = (int64_t)kRowsPerPass * a.T;
  auto put = [&](T* dst, const P& xa, const P& xv) {
    T f[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      f[i] = fused(xa, xv, i);
      const float d = to_f32<T>(f[i]) - K[i];
      s1[i] += d;
      s2[i] = fmaf(d, d, s2[i]);
    }
    if (VEC == 2 && sizeof(T) == 2) {                  // both time steps in one 32-bit store
      *reinterpret_cast<P*>(dst) = *reinterpret_cast<const P*>(f);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) dst[i] = f[i];
    }
  };
  P zpack;
  {
    T* zp = reinterpret_cast<T*>(&zpack);
#pragma unroll
    for (int i = 0; i < VEC; ++i) zp[i] = zero;
  }
  if (MODE == AVFE_FUSE_CONCAT) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const bool on = t_live && (half ? has_v : has_a);
      const T* src = (half ? fv : fa) + (int64_t)r0 * a.T + tp;
      T* dst = tile + (half * a.C + r0) * RS + tp;
      for (int c = r0; c < a.C; c += kRowsPerPass * U, src += U * pass_stride, dst += U * kRowsPerPass * RS) {
        P v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          v[u] = (on && c + u * kRowsPerPass < a.C) ? ld_line(reinterpret_cast<const P*>(src + u * pass_stride)) : zpack;
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (c + u * kRowsPerPass < a.C) put(dst + u * kRowsPerPass * RS, v[u], v[u]);
      }
    }
  } else {
    const T* pa = fa + (int64_t)r0 * a.T + tp;
    const T* pv = fv + (int64_t)r0 * a.T + tp;
    T* dst = tile + r0 * RS + tp;
    for (int c = r0; c < a.C; c += kRowsPerPass * U, pa += U * pass_stride, pv += U * pass_stride,
             dst += U * kRowsPerPass * RS) {
      P va[U], vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool on = t_live && c + u * kRowsPerPass < a.C;
        va[u] = (on && has_a) ? ld_line(reinterpret_cast<const P*>(pa + u * pass_stride)) : zpack;
        vv[u] = (on && has_v) ? ld_line(reinterpret_cast<const P*>(pv + u * pass_stride)) : zpack;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c + u * kRowsPerPass < a.C) put(dst + u * kRowsPerPass * RS, va[u], vv[u]);
    }
  }
  __syncthreads();

  // ---- moments per time step: lanes holding the same time steps are kLanesPerRow apart
  constexpr int kWarps = kThreads / 32;
  float* mean_s = stat;
  float* rstd_s = stat + TT;
  float* kref_s = stat + 2 * TT;
  float* part = stat + 3 * TT;                                    // [2][TT][kWarps]
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
#pragma unroll
    for (int o = 16; o >= kLanesPerRow; o >>= 1) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
    }
    if (lane < kLanesPerRow) {
      part[(tp + i) * kWarps + wid] = s1[i];
      part[(TT + tp + i) * kWarps + wid] = s2[i];
    }
    if (tid < kLanesPerRow) kref_s[tp + i] = K[i];
  }
  __syncthreads();
  if (tid < TT) {
    float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) { t1 += part[tid * kWarps + k]; t2 += part[(TT + tid) * kWarps + k]; }
    const float n = (float)a.Cout;
    const float md = t1 / n;                                       // mean - K
    mean_s[tid] = kref_s[tid] + md;
    rstd_s[tid] = rsqrtf(fmaxf((t2 - t1 * md) / n, 0.0f) + a.eps);
  }
  __syncthreads();

  // ---- drain: a warp takes 32 consecutive channels (their gamma / beta in registers) through
  // all time steps of the tile: column reads of the tile are conflict-free (odd row stride) and
  // every store instruction writes one full 128-byte line of out[b, t, :].  mean / rstd of four
  // time steps come in as one 128-bit broadcast load each.
  T* out = static_cast<T*>(a.out) + (b * (int64_t)a.T + t0) * a.Cout;
  const float4* mean4 = reinterpret_cast<const float4*>(mean_s);
  const float4* rstd4 = reinterpret_cast<const float4*>(rstd_s);
  for (int c = wid * 32 + lane; c < a.Cout; c += kThreads) {
    const float g = a.gamma ? a.gamma[c] : 1.0f, be = a.beta ? a.beta[c] : 0.0f;
    const T* col = tile + c * RS;
    T* o = out + c;
#pragma unroll
    for (int q = 0; q < TT / 4; ++q) {
      const float4 mu = mean4[q], rs = rstd4[q];
      const float m[4] = {mu.x, mu.y, mu.z, mu.w}, r[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = 4 * q + i;
        if (t < nt) *o = from_f32<T>(fmaf((to_f32<T>(col[t]) - m[i]) * r[i], g, be));
        o += a.Cout;
      }
    }
  }
}

template <typename T, int TT>
static size_t smem_bytes(int Cout) {
  const size_t tile = ((size_t)Cout * row_stride<T, TT>() * sizeof(T) + 15) & ~(size_t)15;
  return tile + (size_t)(3 * TT + 2 * TT * (kThreads / 32)) * sizeof(float);
}

template <typename T, int MODE, int TT, int VEC>
static int launch_vec(const Args& a, cudaStream_t s) {
  const size_t smem = smem_bytes<T, TT>(a.Cout);
  if (smem > 227 * 1024) return AVFE_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(fuse_ln_kernel<T, MODE, TT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  const int64_t ctas = a.B * a.tiles_per_sample;
  if (ctas > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  fuse_ln_kernel<T, MODE, TT, VEC><<<(unsigned)ctas, kThreads, smem, s>>>(a);
  count_launch();
  return check_launch();
}

// two time steps per load when every row keeps that alignment (T even, pointers aligned)
template <typename T, int MODE, int TT>
static int launch(const Args& a, cudaStream_t s) {
  const uintptr_t al = 2 * sizeof(T) - 1;
  const bool vec2 = (a.T % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.fa) & al) == 0) &&
                    ((reinterpret_cast<uintptr_t>(a.fv) & al) == 0);
  return vec2 ? launch_vec<T, MODE, TT, 2>(a, s) : launch_vec<T, MODE, TT, 1>(a, s);
}

// TT = time steps per tile: 64 contiguous bytes per channel row when three tiles fit an SM, else 32
template <typename T, int MODE>
static int pick_tile(Args a, cudaStream_t s) {
  constexpr int TT_WIDE = 64 / sizeof(T), TT_NARROW = 32 / sizeof(T);
  if (smem_bytes<T, TT_WIDE>(a.Cout) <= 75 * 1024) {
    a.tiles_per_sample = (a.T + TT_WIDE - 1) / TT_WIDE;
    return launch<T, MODE, TT_WIDE>(a, s);
  }
  a.tiles_per_sample = (a.T + TT_NARROW - 1) / TT_NARROW;
  return launch<T, MODE, TT_NARROW>(a, s);
}

template <typename T>
static int pick_mode(int mode, const Args& a, cudaStream_t s) {
  switch (mode) {
    case AVFE_FUSE_CONCAT: return pick_tile<T, AVFE_FUSE_CONCAT>(a, s);
    case AVFE_FUSE_SUM:    return pick_tile<T, AVFE_FUSE_SUM>(a, s);
    case AVFE_FUSE_WSUM:   return pick_tile<T, AVFE_FUSE_WSUM>(a, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}

}  // namespace fln
}  // namespace avfe

extern "C" int avfe_fuse_layernorm(const void* fa, const void* fv, const uint8_t* mask, int mode,
                                   float w_a, float w_v, int dtype, int64_t B, int64_t C, int64_t T,
                                   const float* gamma, const float* beta, float eps, void* out,
                                   avfe_stream_t stream) {
  using namespace avfe;
  if (B < 0 || C < 0 || T < 0 || !(eps >= 0.0f)) return AVFE_ERR_INVALID_ARG;
  if (mode != AVFE_FUSE_CONCAT && mode != AVFE_FUSE_SUM && mode != AVFE_FUSE_WSUM) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || C == 0 || T == 0) return AVFE_OK;
  if (!fa || !fv || !out) return AVFE_ERR_INVALID_ARG;
  if (C > (1 << 20) || T > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  fln::Args a;
  a.fa = fa; a.fv = fv; a.mask = mask; a.gamma = gamma; a.beta = beta; a.out = out;
  a.wa = w_a; a.wv = w_v; a.eps = eps; a.B = B; a.C = (int)C;
  a.Cout = (int)(mode == AVFE_FUSE_CONCAT ? 2 * C : C);
  a.T = (int)T; a.tiles_per_sample = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case AVFE_F32:  return fln::pick_mode<float>(mode, a, s);
    case AVFE_F16:  return fln::pick_mode<__half>(mode, a, s);
    case AVFE_BF16: return fln::pick_mode<__nv_bfloat16>(mode, a, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}
