// SpecAugment frequency / time masks applied in place to a log-mel batch on the GPU.  Replaces the
// upstream spec_augment call of AmiVideoHFDataset.__getitem__, avsl/whisper_flamingo_ft_ami.py:216-224
// (policies "ls-double" / "ls-basic"; the mask positions are drawn on the host, see
// avsl_b200/audio.py::spec_augment_bands).  A band is a rectangle [f0, f1) x [t0, t1) of one
// clip's [n_mels, n_frames] matrix; only the masked elements are written.
#include "avfe_common.cuh"

namespace avfe {

constexpr int kMaskSplit = 8;    // CTAs per rectangle

__global__ void __launch_bounds__(256)
spec_mask_kernel(float* __restrict__ mel, int n_mels, int64_t n_frames, const int32_t* __restrict__ bands,
                 int n_bands, float fill) {
  const int64_t b = blockIdx.y;
  const int32_t* r = bands + (b * n_bands + blockIdx.x) * 4;
  const int f0 = max(r[0], 0), f1 = min(r[1], n_mels);
  const int64_t t0 = max(r[2], 0), t1 = min((int64_t)r[3], n_frames);
  if (f1 <= f0 || t1 <= t0) return;
  float* base = mel + b * (int64_t)n_mels * n_frames;
  // rows of the rectangle are dealt to the kMaskSplit CTAs; a row is written left to right
  const int64_t w = t1 - t0;
  if (w >= 256) {
    for (int f = f0 + blockIdx.z; f < f1; f += kMaskSplit) {
      float* row = base + (int64_t)f * n_frames + t0;
      for (int64_t t = threadIdx.x; t < w; t += 256) row[t] = fill;
    }
  } else {                                                          // narrow (time) band: a warp per row
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int f = f0 + blockIdx.z * 8 + wid; f < f1; f += kMaskSplit * 8) {
      float* row = base + (int64_t)f * n_frames + t0;
      for (int64_t t = lane; t < w; t += 32) row[t] = fill;
    }
  }
}

// SpecAugment's first step, time warping (Park et al. 2019, sec. 2): a point `center` on the time
// axis of a clip's [n_mels, tau] spectrogram is moved to `warped`; the two sides are stretched /
// squeezed linearly.  Output frame t takes its value from source position
//     s(t) = t * center / warped                                   (t <  warped)
//          = center + (t - warped) * (tau - center) / (tau - warped)   (t >= warped)
// by linear interpolation between the neighbouring source frames (float32, products rounded
// separately).  Frames at and beyond tau (the padding) are copied.  The upstream sampler /
// interpolant is un-vendored: this definition is the build's, the warp points are an input.
__global__ void __launch_bounds__(256)
spec_time_warp_kernel(const float* __restrict__ in, int n_mels, int64_t n_frames, const int32_t* __restrict__ warp,
                      float* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const int tau = min(max(warp[3 * b], 0), (int)min(n_frames, (int64_t)0x7fffffff));
  const int center = warp[3 * b + 1], warped = warp[3 * b + 2];
  const bool active = tau >= 2 && center > 0 && center < tau && warped > 0 && warped < tau;
  const float a0 = active ? __fdiv_rn((float)center, (float)warped) : 1.0f;
  const float a1 = active ? __fdiv_rn((float)(tau - center), (float)(tau - warped)) : 1.0f;
  const int64_t total = (int64_t)n_mels * n_frames;
  const float* src = in + b * total;
  float* dst = out + b * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / n_frames;
    const int t = (int)(i - f * n_frames);
    float v;
    if (!active || t >= tau) v = src[i];
    else {
      const float s = (t < warped) ? __fmul_rn((float)t, a0) : __fadd_rn((float)center, __fmul_rn((float)(t - warped), a1));
      int s0 = (int)floorf(s);
      s0 = min(max(s0, 0), tau - 1);
      const int s1 = min(s0 + 1, tau - 1);
      const float w = fminf(fmaxf(__fsub_rn(s, (float)s0), 0.0f), 1.0f);
      const float* row = src + f * n_frames;
      v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, w), row[s0]), __fmul_rn(w, row[s1]));
    }
    dst[i] = v;
  }
}

}  // namespace avfe

extern "C" int avfe_spec_time_warp_f32(const float* mel, int64_t B, int n_mels, int64_t n_frames, const int32_t* warp,
                                       float* out, avfe_stream_t stream) {
  using namespace avfe;
  if (B < 0 || n_mels < 0 || n_frames < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || n_mels == 0 || n_frames == 0) return AVFE_OK;
  if (!mel || !warp || !out || mel == out) return AVFE_ERR_INVALID_ARG;
  if (B > 65535 || n_frames > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  int64_t gx = ((int64_t)n_mels * n_frames + 255) / 256;
  if (gx > 2 * kNumSMs) gx = 2 * kNumSMs;
  dim3 grid((unsigned)gx, (unsigned)B);
  spec_time_warp_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mel, n_mels, n_frames, warp, out);
  count_launch();
  return check_launch();
}

extern "C" int avfe_spec_mask_f32(float* mel, int64_t B, int n_mels, int64_t n_frames, const int32_t* bands,
                                  int n_bands, float fill, avfe_stream_t stream) {
  using namespace avfe;
  if (B < 0 || n_mels < 0 || n_frames < 0 || n_bands < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || n_bands == 0 || n_mels == 0 || n_frames == 0) return AVFE_OK;
  if (!mel || !bands) return AVFE_ERR_INVALID_ARG;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  dim3 grid((unsigned)n_bands, (unsigned)B, kMaskSplit);
  spec_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mel, n_mels, n_frames, bands, n_bands, fill);
  count_launch();
  return check_launch();
}
