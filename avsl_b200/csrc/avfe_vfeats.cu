// V8 as its call site runs it: load_video_feats_from_decord_reader, utils/hf_video_utils.py:103-138
// (called from safe_load_video_feats_from_hf_object, avsl/whisper_flamingo_ft_ami.py:279-286, with
// what decord returns: uint8 [T,H,W,3] RGB).
//
//   3 channels : gray = np.dot(frames[..., :3], [0.2989, 0.5870, 0.1140])   float64      (:105)
//                if gray.max() > 1.0 (over the WHOLE stack): float32(gray) / 255.0        (:116-117)
//                else the float64 values stay as they are (an all-dark video)
//   1 channel  : uint8 -> float32 / 255.0                                                 (:114-115)
//   H, W >= crop: centre crop [(H-crop)//2 : +crop]                                       (:120-125)
//   otherwise  : cv2.resize(frame, (crop, crop)) -- INTER_LINEAR on the float frame       (:126-132)
//   (x - mean) / std in the array's dtype, trailing channel axis, .astype(float32)        (:135-138, ft_ami:286)
//
// The float64 dot product is evaluated here in plain left-to-right order; numpy hands it to BLAS,
// whose order / FMA use is build-dependent -- but the float32 cast of it is the same for every
// one of the 2^24 RGB triples whichever order is used (checked exhaustively, tests/), so the
// bright branch is bit-exact.  cv2.resize is restated from OpenCV's own generic path (resize.cpp:
// HResizeLinear / VResizeLinear<float>, separately rounded products and sums); x86 wheels of cv2
// dispatch to IPP, whose float32 results differ from that by a few 1e-6 (tolerance in the tests).
#include "avfe_common.cuh"

namespace avfe {

__device__ __forceinline__ double rgb_dot(const uint8_t* __restrict__ p) {
  return __dadd_rn(__dadd_rn(__dmul_rn((double)p[0], 0.2989), __dmul_rn((double)p[1], 0.5870)),
                   __dmul_rn((double)p[2], 0.1140));
}

// flag |= any(gray > 1.0); CTAs stop as soon as somebody has raised it
__global__ void __launch_bounds__(256)
vfeats_bright_kernel(const uint8_t* __restrict__ rgb, int64_t n_px, int* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n_px; base += stride) {   // warp-uniform trip count
    if (*reinterpret_cast<volatile int*>(flag)) return;
    const int64_t i = base + threadIdx.x;
    const bool hit = i < n_px && rgb_dot(rgb + 3 * i) > 1.0;
    if (__any_sync(0xffffffffu, hit)) {
      if ((threadIdx.x & 31) == 0) atomicOr(flag, 1);
      return;
    }
  }
}

// one source sample in the dtype the reference holds at that point: T = float (after /255) or
// double (all-dark 3-channel stack)
template <typename T, int CH>
__device__ __forceinline__ T vfeats_sample(const uint8_t* __restrict__ frame, int W, int r, int c) {
  if (CH == 1) return (T)__fdiv_rn((float)frame[(int64_t)r * W + c], 255.0f);
  const double g = rgb_dot(frame + ((int64_t)r * W + c) * 3);
  if (sizeof(T) == 8) return (T)g;
  return (T)__fdiv_rn(__double2float_rn(g), 255.0f);
}

template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// cv2.resize coordinate of output index d along an axis of length n_src -> n_dst (resize.cpp):
// fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s
__device__ __forceinline__ void resize_coord(int d, int n_src, int n_dst, int& s, float& f) {
  const double inv_scale = __ddiv_rn((double)n_dst, (double)n_src);
  const double scale = __ddiv_rn(1.0, inv_scale);
  f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
  s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
}

template <typename T, int CH, bool RESIZE>
__global__ void __launch_bounds__(256)
vfeats_kernel(const uint8_t* __restrict__ frames, int64_t N, int H, int W, int crop, double mean, double stdv,
              const int* __restrict__ bright, int want_bright, float* __restrict__ out) {
  if (CH == 3 && (*bright != 0) != (want_bright != 0)) return;   // the other instantiation does this stack
  const int sh = (H - crop) / 2, sw = (W - crop) / 2;
  const int64_t per = (int64_t)crop * crop, total = N * per;
  const T m = (T)mean, sd = (T)stdv;        // float32 array op python-float -> float32 (NumPy weak scalars)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / per;
    const int r = (int)((i % per) / crop), c = (int)(i % crop);
    const uint8_t* frame = frames + f * (int64_t)H * W * CH;
    T x;
    if (!RESIZE) {
      x = vfeats_sample<T, CH>(frame, W, sh + r, sw + c);
    } else {
      int sx, sy;
      float fx, fy;
      resize_coord(c, W, crop, sx, fx);
      resize_coord(r, H, crop, sy, fy);
      if (sx < 0) { fx = 0.f; sx = 0; }
      if (sx >= W - 1) { fx = 0.f; sx = W - 1; }
      const int x1 = min(sx + 1, W - 1);
      const int r0 = min(max(sy, 0), H - 1), r1 = min(max(sy + 1, 0), H - 1);
      const T a0 = (T)__fsub_rn(1.f, fx), a1 = (T)fx, b0 = (T)__fsub_rn(1.f, fy), b1 = (T)fy;
      const T h0 = add_rn(mul_rn(vfeats_sample<T, CH>(frame, W, r0, sx), a0), mul_rn(vfeats_sample<T, CH>(frame, W, r0, x1), a1));
      const T h1 = add_rn(mul_rn(vfeats_sample<T, CH>(frame, W, r1, sx), a0), mul_rn(vfeats_sample<T, CH>(frame, W, r1, x1), a1));
      x = add_rn(mul_rn(h0, b0), mul_rn(h1, b1));
    }
    if (sizeof(T) == 8) out[i] = __double2float_rn(__ddiv_rn(__dsub_rn((double)x, (double)m), (double)sd));
    else out[i] = __fdiv_rn(__fsub_rn((float)x, (float)m), (float)sd);
  }
}

}  // namespace avfe

using namespace avfe;

extern "C" size_t avfe_video_feats_workspace_bytes(void) { return 16; }

extern "C" int avfe_video_feats(const uint8_t* frames, int channels, int64_t N, int H, int W, int crop,
                                double mean, double std, float* out, void* workspace,
                                size_t workspace_bytes, avfe_stream_t stream) {
  if (N < 0 || H <= 0 || W <= 0 || crop <= 0 || (channels != 1 && channels != 3)) return AVFE_ERR_INVALID_ARG;
  if (N == 0) return AVFE_OK;
  if (!frames || !out) return AVFE_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = N * (int64_t)crop * crop;
  int64_t ctas = (total + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  const bool resize = crop > H || crop > W;            // start_h < 0 or start_w < 0 (:123,126)
  if (channels == 1) {
    if (resize) vfeats_kernel<float, 1, true><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, nullptr, 1, out);
    else vfeats_kernel<float, 1, false><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, nullptr, 1, out);
    count_launch();
    return check_launch();
  }
  if (!workspace || workspace_bytes < avfe_video_feats_workspace_bytes() || !aligned16(workspace)) return AVFE_ERR_WORKSPACE;
  int* flag = static_cast<int*>(workspace);
  if (cudaMemsetAsync(flag, 0, 16, s) != cudaSuccess) return AVFE_ERR_CUDA;
  const int64_t n_px = N * (int64_t)H * W;
  int64_t fctas = (n_px + 255) / 256;
  if (fctas > (int64_t)kNumSMs * 8) fctas = (int64_t)kNumSMs * 8;
  vfeats_bright_kernel<<<(unsigned)fctas, 256, 0, s>>>(frames, n_px, flag);
  // both branches are enqueued; the one the flag does not select returns at once (no host sync)
  if (resize) {
    vfeats_kernel<float, 3, true><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, flag, 1, out);
    vfeats_kernel<double, 3, true><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, flag, 0, out);
  } else {
    vfeats_kernel<float, 3, false><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, flag, 1, out);
    vfeats_kernel<double, 3, false><<<(unsigned)ctas, 256, 0, s>>>(frames, N, H, W, crop, mean, std, flag, 0, out);
  }
  count_launch(3);
  return check_launch();
}
