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

}  // namespace avfe

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
