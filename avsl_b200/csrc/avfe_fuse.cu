// Modality fusion (concat / sum / weighted sum) under a per-sample modality-dropout mask.
// Replaces avsl/modules/av_hubert_encoder.py:292-298,315-326 of the reference.
//
// Pure streaming op: HBM-bound, no reuse, so no shared memory and no tensor cores.  Each CTA
// owns a contiguous run of 16-byte vectors of ONE sample; the sample's two mask bytes are
// fetched by lane 0 of every warp and broadcast with a warp shuffle, so the masked-out
// modality is never read (its bytes do not count against the roofline).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "avfe_common.cuh"

namespace avfe {

template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// out = a (+) v on one 16-byte vector.  Explicit _rn intrinsics: no FMA contraction, so the
// result is the two-rounding value torch produces on the CPU.
template <typename T, int MODE>
__device__ __forceinline__ uint4 combine(const uint4& a, const uint4& v, float wa, float wv) {
  constexpr int N = Vec16<T>::N;
  uint4 r;
  const T* pa = reinterpret_cast<const T*>(&a);
  const T* pv = reinterpret_cast<const T*>(&v);
  T* pr = reinterpret_cast<T*>(&r);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float x = to_f32<T>(pa[i]), y = to_f32<T>(pv[i]);
    float s = (MODE == AVFE_FUSE_SUM) ? __fadd_rn(x, y)
                                      : __fadd_rn(__fmul_rn(wa, x), __fmul_rn(wv, y));
    pr[i] = from_f32<T>(s);
  }
  return r;
}

__device__ __forceinline__ unsigned sample_mask(const uint8_t* mask, int64_t b) {
  unsigned m = 3u;
  if (mask != nullptr) {
    unsigned v = 0u;
    if ((threadIdx.x & 31) == 0) v = (mask[2 * b] ? 1u : 0u) | (mask[2 * b + 1] ? 2u : 0u);
    m = __shfl_sync(0xffffffffu, v, 0);
  }
  return m;
}

constexpr int kFuseThreads = 256;
constexpr int kFuseUnroll = 4;

template <typename T, int MODE>
__global__ void __launch_bounds__(kFuseThreads)
fuse_vec_kernel(const uint4* __restrict__ fa, const uint4* __restrict__ fv,
                const uint8_t* __restrict__ mask, float wa, float wv, int64_t nvec,
                uint4* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const unsigned m = sample_mask(mask, b);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
  const uint4* pa = fa + b * nvec;
  const uint4* pv = fv + b * nvec;
  const int64_t base = (int64_t)blockIdx.x * (kFuseThreads * kFuseUnroll) + threadIdx.x;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);

  uint4 ra[kFuseUnroll], rv[kFuseUnroll];
#pragma unroll
  for (int j = 0; j < kFuseUnroll; ++j) {  // all loads first: 2*UNROLL 128-bit requests in flight
    const int64_t i = base + (int64_t)j * kFuseThreads;
    ra[j] = (has_a && i < nvec) ? ldg_stream(pa + i) : zero;
    rv[j] = (has_v && i < nvec) ? ldg_stream(pv + i) : zero;
  }
  if (MODE == AVFE_FUSE_CONCAT) {
    uint4* oa = out + b * 2 * nvec;
    uint4* ov = oa + nvec;
#pragma unroll
    for (int j = 0; j < kFuseUnroll; ++j) {
      const int64_t i = base + (int64_t)j * kFuseThreads;
      if (i < nvec) {
        stg_stream(oa + i, ra[j]);
        stg_stream(ov + i, rv[j]);
      }
    }
  } else {
    uint4* o = out + b * nvec;
#pragma unroll
    for (int j = 0; j < kFuseUnroll; ++j) {
      const int64_t i = base + (int64_t)j * kFuseThreads;
      if (i < nvec) stg_stream(o + i, combine<T, MODE>(ra[j], rv[j], wa, wv));
    }
  }
}

// Element-wise path for shapes/pointers that are not 16-byte friendly (still on the GPU).
template <typename T, int MODE>
__global__ void __launch_bounds__(kFuseThreads)
fuse_scalar_kernel(const T* __restrict__ fa, const T* __restrict__ fv,
                   const uint8_t* __restrict__ mask, float wa, float wv, int64_t n,
                   T* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const unsigned m = sample_mask(mask, b);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
  const T zero = from_f32<T>(0.0f);
  for (int64_t i = (int64_t)blockIdx.x * kFuseThreads + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * kFuseThreads) {
    const T a = has_a ? fa[b * n + i] : zero;
    const T v = has_v ? fv[b * n + i] : zero;
    if (MODE == AVFE_FUSE_CONCAT) {
      out[b * 2 * n + i] = a;
      out[b * 2 * n + n + i] = v;
    } else {
      const float x = to_f32<T>(a), y = to_f32<T>(v);
      const float s = (MODE == AVFE_FUSE_SUM) ? __fadd_rn(x, y)
                                              : __fadd_rn(__fmul_rn(wa, x), __fmul_rn(wv, y));
      out[b * n + i] = from_f32<T>(s);
    }
  }
}

template <typename T, int MODE>
static int launch_fuse(const void* fa, const void* fv, const uint8_t* mask, float wa, float wv,
                       int64_t B, int64_t n, void* out, cudaStream_t s) {
  const bool vec_ok = ((n * (int64_t)sizeof(T)) % 16 == 0) && aligned16(fa) && aligned16(fv) &&
                      aligned16(out);
  if (vec_ok) {
    const int64_t nvec = n * (int64_t)sizeof(T) / 16;
    const int64_t per_cta = kFuseThreads * kFuseUnroll;
    dim3 grid((unsigned)((nvec + per_cta - 1) / per_cta), (unsigned)B);
    fuse_vec_kernel<T, MODE><<<grid, kFuseThreads, 0, s>>>(
        static_cast<const uint4*>(fa), static_cast<const uint4*>(fv), mask, wa, wv, nvec,
        static_cast<uint4*>(out));
  } else {
    int64_t gx = (n + kFuseThreads - 1) / kFuseThreads;
    if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
    dim3 grid((unsigned)gx, (unsigned)B);
    fuse_scalar_kernel<T, MODE><<<grid, kFuseThreads, 0, s>>>(
        static_cast<const T*>(fa), static_cast<const T*>(fv), mask, wa, wv, n,
        static_cast<T*>(out));
  }
  count_launch();
  return check_launch();
}

template <typename T>
static int dispatch_mode(int mode, const void* fa, const void* fv, const uint8_t* mask, float wa,
                         float wv, int64_t B, int64_t n, void* out, cudaStream_t s) {
  switch (mode) {
    case AVFE_FUSE_CONCAT: return launch_fuse<T, AVFE_FUSE_CONCAT>(fa, fv, mask, wa, wv, B, n, out, s);
    case AVFE_FUSE_SUM:    return launch_fuse<T, AVFE_FUSE_SUM>(fa, fv, mask, wa, wv, B, n, out, s);
    case AVFE_FUSE_WSUM:   return launch_fuse<T, AVFE_FUSE_WSUM>(fa, fv, mask, wa, wv, B, n, out, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}

}  // namespace avfe

extern "C" int avfe_fuse(const void* fa, const void* fv, const uint8_t* mask, int mode, float w_a,
                         float w_v, int dtype, int64_t B, int64_t C, int64_t T, void* out,
                         avfe_stream_t stream) {
  if (B < 0 || C < 0 || T < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || C == 0 || T == 0) return AVFE_OK;
  if (!fa || !fv || !out) return AVFE_ERR_INVALID_ARG;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;  // gridDim.y
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = C * T;
  switch (dtype) {
    case AVFE_F32:  return avfe::dispatch_mode<float>(mode, fa, fv, mask, w_a, w_v, B, n, out, s);
    case AVFE_F16:  return avfe::dispatch_mode<__half>(mode, fa, fv, mask, w_a, w_v, B, n, out, s);
    case AVFE_BF16: return avfe::dispatch_mode<__nv_bfloat16>(mode, fa, fv, mask, w_a, w_v, B, n, out, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}
