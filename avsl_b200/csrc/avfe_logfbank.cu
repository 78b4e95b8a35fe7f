// AV-HuBERT audio features: python_speech_features.logfbank(audio, samplerate=16000) (26 log
// mel-filterbank energies per 25 ms / 10 ms frame), frame stacking and per-row normalisation.
// Replaces extract_logfbank_features + audio_to_tensor, preprocess/audio_process.py:152-197
// (same code in utils/data_loading.py:181-201); called from process_audio_for_av_hubert (:199-236)
// and process_audio_dual_encoder (:267-299).
//
//   logfbank_kernel   one CTA = 8 consecutive frames of one clip: pre-emphasis (float32, as numpy
//                     evaluates it) while staging the audio span, four complex 512-point FFTs (two
//                     real frames each, 8 x 8 x 8, 64 threads per FFT), power spectrum / 512,
//                     26 triangular filters, log, stack `stack` frames per row, (x - mean) / (std + 1e-5)
//
// Frames, spectra and filterbank energies never touch HBM: traffic is the audio read plus the
// [rows, 26 * stack] output.
#include "avfe_common.cuh"
#include "avfe_logfbank_core.cuh"

namespace avfe {
namespace fbk {

#include "avfe_logfbank_tables.inc"   // kTw512Re / kTw512Im: exp(-2 pi i j / 512), float64 -> float32

constexpr int kTileFrames = 8;
constexpr int kFfts = kTileFrames / 2;
constexpr int kThreads = kFfts * kFftThreads;                 // 256
constexpr int kSpan = (kTileFrames - 1) * kHop + kFrame;       // 1520 samples per tile
constexpr int kMaxFilt = 64;

struct Smem {
  float y[kSpan];                       // pre-emphasised samples of the tile (zero past the clip)
  float2 tw[kNfft];
  float2 S[kFfts][kSFloat2];            // exchange storage; the power rows overwrite it
  float2 C[kFfts][kNfft];               // spectra
  float feat[kTileFrames][kMaxFilt];    // log energies, frame-major = the stacked row layout
  int lo[kMaxFilt], hi[kMaxFilt];       // filter supports
  float stat[kTileFrames][2];
};

// number of frames framesig makes of `len` samples (python_speech_features.sigproc.framesig)
__host__ __device__ inline int64_t num_frames(int64_t len) {
  return len <= kFrame ? 1 : 1 + (len - kFrame + kHop - 1) / kHop;
}

// filter supports [first, last + 1) of the nonzero weights, once per call: one warp per filter
__global__ void __launch_bounds__(256)
logfbank_prep_kernel(const float* __restrict__ fb, int nfilt, int* __restrict__ supports) {
  const int lane = threadIdx.x & 31;
  for (int m = threadIdx.x >> 5; m < nfilt; m += blockDim.x >> 5) {
    int lo = kBins, hi = 0;
    for (int k = lane; k < kBins; k += 32)
      if (fb[(size_t)m * kBins + k] != 0.0f) { lo = min(lo, k); hi = max(hi, k + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { supports[2 * m] = lo < hi ? lo : 0; supports[2 * m + 1] = lo < hi ? hi : 0; }
  }
}

// grid = (tile slots, clips): a CTA walks the tiles blockIdx.x, + gridDim.x, ... of its clip, so
// that the twiddle table and the filter supports are set up once per CTA, not once per 8 frames
__global__ void __launch_bounds__(kThreads)
logfbank_kernel(const float* __restrict__ audio, const int64_t* __restrict__ offsets,
                const int64_t* __restrict__ row_offsets, const float* __restrict__ fb, int nfilt,
                const int* __restrict__ supports, int stack, int normalize, float* __restrict__ out) {
  __shared__ Smem sm;
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int64_t beg = offsets[b], len = offsets[b + 1] - beg;
  const int64_t nfr = num_frames(len);
  const int64_t rows = (nfr + stack - 1) / stack;               // stacked rows (zero-padded tail)
  if ((int64_t)blockIdx.x * kTileFrames >= rows * stack) return;   // CTA-uniform
  const float* clip = audio + beg;
  for (int i = tid; i < kNfft; i += kThreads) sm.tw[i] = make_float2(kTw512Re[i], kTw512Im[i]);
  if (tid < nfilt) { sm.lo[tid] = supports[2 * tid]; sm.hi[tid] = supports[2 * tid + 1]; }

  for (int64_t f0 = (int64_t)blockIdx.x * kTileFrames; f0 < rows * stack; f0 += (int64_t)gridDim.x * kTileFrames) {
  for (int i = tid; i < kSpan; i += kThreads) sm.y[i] = preemph_sample(clip, len, f0 * kHop + i);
  __syncthreads();

  const int g = tid / kFftThreads, t = tid % kFftThreads;       // FFT g: tile frames 2g, 2g + 1
  float2* S = sm.S[g];
  float2* C = sm.C[g];
  step1(t, sm.y + (2 * g) * kHop, sm.y + (2 * g + 1) * kHop, sm.tw, S);
  __syncthreads();
  float2 x[8];
  step2_load(t, S, x);
  __syncthreads();
  step2_store(t, sm.tw, x, S);
  __syncthreads();
  step3(t, S, C);
  __syncthreads();
  float* P = reinterpret_cast<float*>(S);                       // two power rows per FFT
  power_rows(t, C, P, P + kPStride);
  __syncthreads();

  // log filterbank energies: thread = (frame, filter); frames past the clip's last one are the
  // zero rows extract_logfbank_features appends before stacking
  for (int i = tid; i < kTileFrames * nfilt; i += kThreads) {
    const int f = i / nfilt, m = i - f * nfilt;
    const float* Pf = reinterpret_cast<const float*>(sm.S[f >> 1]) + (f & 1) * kPStride;
    sm.feat[f][m] = (f0 + f < nfr) ? log_fbank(Pf, fb + (size_t)m * kBins, sm.lo[m], sm.hi[m]) : 0.0f;
  }
  __syncthreads();

  // rows of `stack` consecutive frames; audio_to_tensor: (x - mean) / (std + 1e-5), population std
  const int rows_here = kTileFrames / stack, width = stack * nfilt;
  if (normalize) {
    const int lane = tid & 31, wid = tid >> 5;
    for (int r = wid; r < rows_here; r += kThreads / 32) {
      float s = 0.0f;
      for (int i = lane; i < width; i += 32) s += sm.feat[r * stack + i / nfilt][i % nfilt];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / (float)width;
      float v = 0.0f;
      for (int i = lane; i < width; i += 32) {
        const float d = sm.feat[r * stack + i / nfilt][i % nfilt] - mean;
        v = fmaf(d, d, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) { sm.stat[r][0] = mean; sm.stat[r][1] = sqrtf(v / (float)width) + 1e-5f; }
    }
    __syncthreads();
  }
  const int64_t row0 = f0 / stack;
  float* o = out + (row_offsets[b] + row0) * width;
  for (int i = tid; i < rows_here * width; i += kThreads) {
    const int r = i / width, c = i - r * width;
    if (row0 + r >= rows) continue;
    float v = sm.feat[r * stack + c / nfilt][c % nfilt];
    if (normalize) v = __fdiv_rn(v - sm.stat[r][0], sm.stat[r][1]);
    o[i] = v;
  }
  __syncthreads();                                              // staging buffers are reused
  }
}

}  // namespace fbk
}  // namespace avfe

using namespace avfe;

extern "C" int64_t avfe_logfbank_num_frames(int64_t n_samples) {
  return n_samples < 0 ? 0 : fbk::num_frames(n_samples);
}

extern "C" size_t avfe_logfbank_workspace_bytes(void) { return 2 * fbk::kMaxFilt * sizeof(int); }

extern "C" int avfe_logfbank_f32(const float* audio, const int64_t* offsets, const int64_t* row_offsets,
                                 int64_t B, int64_t max_samples, const float* fbank, int nfilt,
                                 int stack, int normalize, float* out, void* workspace,
                                 size_t workspace_bytes, avfe_stream_t stream) {
  if (B < 0 || nfilt <= 0 || stack <= 0 || max_samples < 0) return AVFE_ERR_INVALID_ARG;
  if (nfilt > fbk::kMaxFilt || (fbk::kTileFrames % stack) != 0) return AVFE_ERR_UNSUPPORTED;
  if (B == 0) return AVFE_OK;
  if (!audio || !offsets || !row_offsets || !fbank || !out) return AVFE_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < avfe_logfbank_workspace_bytes()) return AVFE_ERR_WORKSPACE;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  const int64_t nfr = fbk::num_frames(max_samples);
  const int64_t padded = (nfr + stack - 1) / stack * stack;
  const int64_t tiles = (padded + fbk::kTileFrames - 1) / fbk::kTileFrames;
  if (tiles > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* supports = static_cast<int*>(workspace);
  fbk::logfbank_prep_kernel<<<1, 256, 0, s>>>(fbank, nfilt, supports);
  // about 4 resident CTAs per SM in total, each walking several tiles of its clip
  int64_t slots = (4 * (int64_t)kNumSMs + B - 1) / B;
  if (slots > tiles) slots = tiles;
  if (slots < 1) slots = 1;
  dim3 grid((unsigned)slots, (unsigned)B);
  fbk::logfbank_kernel<<<grid, fbk::kThreads, 0, s>>>(audio, offsets, row_offsets, fbank, nfilt, supports, stack,
                                                      normalize, out);
  count_launch(2);
  return check_launch();
}
