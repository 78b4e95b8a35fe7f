// Backward of the fusion block, so that the drop-in can sit in the reference's TRAINING forward
// (AVHuBERTEncoderWrapper.forward runs under self.training with modality dropout,
// avsl/modules/av_hubert_encoder.py:292-330): gradients must reach the audio / video feature
// extractors and the LayerNorm parameters.
//
//   fuse_bwd_kernel        d(cat | add | weighted sum)/d(fa, fv) under the modality mask: a masked-out
//                          modality was zero-filled, so its gradient is zero
//   fuse_ln_bwd_kernel     backward of fusion + transpose(1,2) + LayerNorm in one kernel: reads
//                          grad_out [B,T,C'] and the forward inputs [B,C,T] x 2, recomputes the
//                          moments (same shifted sums as the forward), writes grad_fa / grad_fv in
//                          the inputs' [B,C,T] layout and per-CTA partial sums of d(gamma), d(beta)
//   fuse_ln_bwd_reduce_kernel  sums the partials in a fixed order (deterministic)
//
// LayerNorm backward for one (b, t):  xh = (x - mean) * rstd,  g' = g * gamma,
//   dx = rstd * (g' - mean_c(g') - xh * mean_c(g' * xh)),  dgamma[c] = sum_{b,t} g * xh,  dbeta[c] = sum g.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "avfe_common.cuh"

namespace avfe {
namespace fbw {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------ plain fusion
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
fuse_bwd_kernel(const T* __restrict__ gout, const uint8_t* __restrict__ mask, float wa, float wv, int64_t n,
                T* __restrict__ gfa, T* __restrict__ gfv) {
  const int64_t b = blockIdx.y;
  unsigned m = 3u;
  if (mask != nullptr) m = (mask[2 * b] ? 1u : 0u) | (mask[2 * b + 1] ? 2u : 0u);
  const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
  const T zero = from_f32<T>(0.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T ga, gv;
    if (MODE == AVFE_FUSE_CONCAT) {
      ga = has_a ? gout[b * 2 * n + i] : zero;
      gv = has_v ? gout[b * 2 * n + n + i] : zero;
    } else {
      const T g = gout[b * n + i];
      if (MODE == AVFE_FUSE_SUM) {
        ga = has_a ? g : zero;
        gv = has_v ? g : zero;
      } else {
        ga = has_a ? from_f32<T>(__fmul_rn(wa, to_f32<T>(g))) : zero;
        gv = has_v ? from_f32<T>(__fmul_rn(wv, to_f32<T>(g))) : zero;
      }
    }
    gfa[b * n + i] = ga;
    gfv[b * n + i] = gv;
  }
}

template <typename T>
static int launch_fuse_bwd(int mode, const void* gout, const uint8_t* mask, float wa, float wv, int64_t B, int64_t n,
                           void* gfa, void* gfv, cudaStream_t s) {
  int64_t gx = (n + 255) / 256;
  if (gx > 8 * kNumSMs) gx = 8 * kNumSMs;
  dim3 grid((unsigned)gx, (unsigned)B);
  const T* g = static_cast<const T*>(gout);
  T* a = static_cast<T*>(gfa);
  T* v = static_cast<T*>(gfv);
  switch (mode) {
    case AVFE_FUSE_CONCAT: fuse_bwd_kernel<T, AVFE_FUSE_CONCAT><<<grid, 256, 0, s>>>(g, mask, wa, wv, n, a, v); break;
    case AVFE_FUSE_SUM: fuse_bwd_kernel<T, AVFE_FUSE_SUM><<<grid, 256, 0, s>>>(g, mask, wa, wv, n, a, v); break;
    case AVFE_FUSE_WSUM: fuse_bwd_kernel<T, AVFE_FUSE_WSUM><<<grid, 256, 0, s>>>(g, mask, wa, wv, n, a, v); break;
    default: return AVFE_ERR_INVALID_ARG;
  }
  count_launch();
  return check_launch();
}

// ------------------------------------------------------------------ fusion + transpose + LayerNorm
constexpr int kTT = 32;        // time steps per tile: lane = time step
constexpr int kWarps = 8;
constexpr int kCC = 256;       // channels of grad_out staged per round
constexpr int kGS = kCC + 1;   // row stride of the staged tile (odd: conflict-free both ways)

struct LnBwdArgs {
  const void* fa;
  const void* fv;
  const uint8_t* mask;
  const float* gamma;       // [C'] or nullptr (= 1)
  const void* gout;         // [B, T, C']
  void* gfa;                // [B, C, T]
  void* gfv;
  float* partial;           // [grid][2][C']
  float wa, wv, eps;
  int64_t B;
  int C, Cout, T, tiles_per_sample;
  int64_t n_tiles;
};

// the fused (pre-LayerNorm) value of channel c at time t0 + lane, as the forward kernel stores it
template <typename T, int MODE>
__device__ __forceinline__ float fused_value(const T* __restrict__ fa, const T* __restrict__ fv, int c, int C, int Tn,
                                            bool has_a, bool has_v, float wa, float wv, bool ok) {
  if (!ok) return 0.0f;
  if (MODE == AVFE_FUSE_CONCAT) {
    if (c < C) return has_a ? to_f32<T>(fa[(int64_t)c * Tn]) : 0.0f;
    return has_v ? to_f32<T>(fv[(int64_t)(c - C) * Tn]) : 0.0f;
  }
  const float a = has_a ? to_f32<T>(fa[(int64_t)c * Tn]) : 0.0f;
  const float v = has_v ? to_f32<T>(fv[(int64_t)c * Tn]) : 0.0f;
  const float s = (MODE == AVFE_FUSE_SUM) ? __fadd_rn(a, v) : __fadd_rn(__fmul_rn(wa, a), __fmul_rn(wv, v));
  return to_f32<T>(from_f32<T>(s));      // the forward rounds the fused value to T before normalising
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kWarps * 32, 2)
fuse_ln_bwd_kernel(const LnBwdArgs a) {
  extern __shared__ __align__(16) unsigned char fbw_smem[];
  float* gt = reinterpret_cast<float*>(fbw_smem);           // [kTT][kGS] staged grad_out
  float* dgam = gt + kTT * kGS;                             // [Cout]
  float* dbet = dgam + a.Cout;                              // [Cout]
  float* red = dbet + a.Cout;                               // [kWarps][4][32]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int c = tid; c < 2 * a.Cout; c += kWarps * 32) dgam[c] = 0.0f;
  __syncthreads();
  const float inv_c = 1.0f / (float)a.Cout;

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t b = tile / a.tiles_per_sample;
    const int t0 = (int)(tile % a.tiles_per_sample) * kTT;
    const bool ok = t0 + lane < a.T;
    unsigned m = 3u;
    if (a.mask != nullptr) m = (a.mask[2 * b] ? 1u : 0u) | (a.mask[2 * b + 1] ? 2u : 0u);
    const bool has_a = (m & 1u) != 0, has_v = (m & 2u) != 0;
    const T* fa = static_cast<const T*>(a.fa) + b * (int64_t)a.C * a.T + t0 + lane;
    const T* fv = static_cast<const T*>(a.fv) + b * (int64_t)a.C * a.T + t0 + lane;
    const T* go = static_cast<const T*>(a.gout) + (b * (int64_t)a.T + t0) * a.Cout;
    const float K = fused_value<T, MODE>(fa, fv, 0, a.C, a.T, has_a, has_v, a.wa, a.wv, ok);

    float mean = 0.f, rstd = 0.f, c1 = 0.f, c2 = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      float sx = 0.f, sxx = 0.f, a1 = 0.f, a2 = 0.f;
      for (int c0 = 0; c0 < a.Cout; c0 += kCC) {
        const int nc = min(kCC, a.Cout - c0);
        __syncthreads();                                     // the previous round's readers are done
        for (int tt = wid; tt < kTT; tt += kWarps) {
          const bool row_ok = t0 + tt < a.T;
          for (int c = lane; c < nc; c += 32)
            gt[tt * kGS + c] = row_ok ? to_f32<T>(go[(int64_t)tt * a.Cout + c0 + c]) : 0.0f;
        }
        __syncthreads();
        for (int cl = wid; cl < nc; cl += kWarps) {
          const int c = c0 + cl;
          const float x = fused_value<T, MODE>(fa, fv, c, a.C, a.T, has_a, has_v, a.wa, a.wv, ok);
          const float g = gt[lane * kGS + cl];
          const float gg = g * (a.gamma ? a.gamma[c] : 1.0f);
          if (pass == 0) {
            const float d = x - K;
            sx += d; sxx += d * d; a1 += gg; a2 += gg * d;
          } else {
            const float xh = (x - mean) * rstd;
            const float dx = rstd * (gg - c1 - xh * c2);
            // route to the inputs
            if (ok) {
              T* gfa = static_cast<T*>(a.gfa) + b * (int64_t)a.C * a.T + t0 + lane;
              T* gfv = static_cast<T*>(a.gfv) + b * (int64_t)a.C * a.T + t0 + lane;
              if (MODE == AVFE_FUSE_CONCAT) {
                if (c < a.C) gfa[(int64_t)c * a.T] = from_f32<T>(has_a ? dx : 0.0f);
                else gfv[(int64_t)(c - a.C) * a.T] = from_f32<T>(has_v ? dx : 0.0f);
              } else {
                const float sa = (MODE == AVFE_FUSE_WSUM) ? a.wa : 1.0f, sv = (MODE == AVFE_FUSE_WSUM) ? a.wv : 1.0f;
                gfa[(int64_t)c * a.T] = from_f32<T>(has_a ? sa * dx : 0.0f);
                gfv[(int64_t)c * a.T] = from_f32<T>(has_v ? sv * dx : 0.0f);
              }
            }
            float pg = ok ? g * xh : 0.0f, pb = ok ? g : 0.0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              pg += __shfl_xor_sync(0xffffffffu, pg, o);
              pb += __shfl_xor_sync(0xffffffffu, pb, o);
            }
            if (lane == 0) { dgam[c] += pg; dbet[c] += pb; }   // channel c belongs to this warp only
          }
        }
      }
      if (pass == 0) {
        red[(wid * 4 + 0) * 32 + lane] = sx;
        red[(wid * 4 + 1) * 32 + lane] = sxx;
        red[(wid * 4 + 2) * 32 + lane] = a1;
        red[(wid * 4 + 3) * 32 + lane] = a2;
        __syncthreads();
        sx = sxx = a1 = a2 = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          sx += red[(w * 4 + 0) * 32 + lane];
          sxx += red[(w * 4 + 1) * 32 + lane];
          a1 += red[(w * 4 + 2) * 32 + lane];
          a2 += red[(w * 4 + 3) * 32 + lane];
        }
        const float dm = sx * inv_c;                         // mean - K
        mean = K + dm;
        const float var = fmaxf(sxx * inv_c - dm * dm, 0.0f);
        rstd = rsqrtf(var + a.eps);
        c1 = a1 * inv_c;
        c2 = rstd * (a2 - dm * a1) * inv_c;
      }
    }
  }
  __syncthreads();
  float* part = a.partial + (int64_t)blockIdx.x * 2 * a.Cout;
  for (int c = tid; c < 2 * a.Cout; c += kWarps * 32) part[c] = dgam[c];
}

__global__ void __launch_bounds__(256)
fuse_ln_bwd_reduce_kernel(const float* __restrict__ partial, int n_part, int Cout, float* __restrict__ dgamma,
                          float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 2 * Cout) return;
  float s = 0.0f;
  for (int p = 0; p < n_part; ++p) s += partial[(int64_t)p * 2 * Cout + c];
  if (c < Cout) { if (dgamma) dgamma[c] = s; }
  else if (dbeta) dbeta[c - Cout] = s;
}

static int ln_bwd_grid(int64_t n_tiles) {
  int64_t g = 2 * (int64_t)kNumSMs;
  return (int)(n_tiles < g ? n_tiles : g);
}

static size_t ln_bwd_smem(int Cout) { return sizeof(float) * ((size_t)kTT * kGS + 2 * (size_t)Cout + kWarps * 4 * 32); }

template <typename T, int MODE>
static int launch_ln_bwd(const LnBwdArgs& a, int grid, cudaStream_t s) {
  const size_t smem = ln_bwd_smem(a.Cout);
  if (smem > 200 * 1024) return AVFE_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(fuse_ln_bwd_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return AVFE_ERR_CUDA;
  fuse_ln_bwd_kernel<T, MODE><<<grid, kWarps * 32, smem, s>>>(a);
  count_launch();
  return check_launch();
}

template <typename T>
static int dispatch_ln_bwd(int mode, const LnBwdArgs& a, int grid, cudaStream_t s) {
  switch (mode) {
    case AVFE_FUSE_CONCAT: return launch_ln_bwd<T, AVFE_FUSE_CONCAT>(a, grid, s);
    case AVFE_FUSE_SUM: return launch_ln_bwd<T, AVFE_FUSE_SUM>(a, grid, s);
    case AVFE_FUSE_WSUM: return launch_ln_bwd<T, AVFE_FUSE_WSUM>(a, grid, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}

}  // namespace fbw
}  // namespace avfe

using namespace avfe;

extern "C" int avfe_fuse_backward(const void* grad_out, const uint8_t* mask, int mode, float w_a, float w_v,
                                  int dtype, int64_t B, int64_t C, int64_t T, void* grad_fa, void* grad_fv,
                                  avfe_stream_t stream) {
  if (B < 0 || C < 0 || T < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || C == 0 || T == 0) return AVFE_OK;
  if (!grad_out || !grad_fa || !grad_fv) return AVFE_ERR_INVALID_ARG;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = C * T;
  switch (dtype) {
    case AVFE_F32: return fbw::launch_fuse_bwd<float>(mode, grad_out, mask, w_a, w_v, B, n, grad_fa, grad_fv, s);
    case AVFE_F16: return fbw::launch_fuse_bwd<__half>(mode, grad_out, mask, w_a, w_v, B, n, grad_fa, grad_fv, s);
    case AVFE_BF16: return fbw::launch_fuse_bwd<__nv_bfloat16>(mode, grad_out, mask, w_a, w_v, B, n, grad_fa, grad_fv, s);
    default: return AVFE_ERR_INVALID_ARG;
  }
}

extern "C" size_t avfe_fuse_layernorm_backward_workspace_bytes(int64_t B, int64_t C, int64_t T, int mode) {
  if (B <= 0 || C <= 0 || T <= 0) return 16;
  const int64_t Cout = (mode == AVFE_FUSE_CONCAT) ? 2 * C : C;
  const int64_t n_tiles = B * ((T + fbw::kTT - 1) / fbw::kTT);
  return (size_t)fbw::ln_bwd_grid(n_tiles) * 2 * (size_t)Cout * sizeof(float);
}

extern "C" int avfe_fuse_layernorm_backward(const void* fa, const void* fv, const uint8_t* mask, int mode,
                                            float w_a, float w_v, int dtype, int64_t B, int64_t C, int64_t T,
                                            const float* gamma, float eps, const void* grad_out, void* grad_fa,
                                            void* grad_fv, float* grad_gamma, float* grad_beta, void* workspace,
                                            size_t workspace_bytes, avfe_stream_t stream) {
  if (B < 0 || C < 0 || T < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || C == 0 || T == 0) return AVFE_OK;
  if (!fa || !fv || !grad_out || !grad_fa || !grad_fv) return AVFE_ERR_INVALID_ARG;
  if (mode != AVFE_FUSE_CONCAT && mode != AVFE_FUSE_SUM && mode != AVFE_FUSE_WSUM) return AVFE_ERR_INVALID_ARG;
  if (C > (1 << 20) || T > (1 << 30)) return AVFE_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < avfe_fuse_layernorm_backward_workspace_bytes(B, C, T, mode) || !aligned16(workspace))
    return AVFE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  fbw::LnBwdArgs a;
  a.fa = fa; a.fv = fv; a.mask = mask; a.gamma = gamma; a.gout = grad_out; a.gfa = grad_fa; a.gfv = grad_fv;
  a.partial = static_cast<float*>(workspace);
  a.wa = w_a; a.wv = w_v; a.eps = eps; a.B = B; a.C = (int)C; a.T = (int)T;
  a.Cout = (mode == AVFE_FUSE_CONCAT) ? 2 * (int)C : (int)C;
  a.tiles_per_sample = (int)((T + fbw::kTT - 1) / fbw::kTT);
  a.n_tiles = B * a.tiles_per_sample;
  const int grid = fbw::ln_bwd_grid(a.n_tiles);
  int rc;
  switch (dtype) {
    case AVFE_F32: rc = fbw::dispatch_ln_bwd<float>(mode, a, grid, s); break;
    case AVFE_F16: rc = fbw::dispatch_ln_bwd<__half>(mode, a, grid, s); break;
    case AVFE_BF16: rc = fbw::dispatch_ln_bwd<__nv_bfloat16>(mode, a, grid, s); break;
    default: return AVFE_ERR_INVALID_ARG;
  }
  if (rc != AVFE_OK) return rc;
  if (grad_gamma || grad_beta) {
    fbw::fuse_ln_bwd_reduce_kernel<<<(2 * a.Cout + 255) / 256, 256, 0, s>>>(a.partial, grid, a.Cout, grad_gamma, grad_beta);
    count_launch();
    rc = check_launch();
  }
  return rc;
}
