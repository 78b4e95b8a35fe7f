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
constexpr int kMaxFilt = 40;             // filters per bank (python_speech_features: 26; fbank-40 fits)

constexpr int kMaxTaps = 640;             // packed nonzero filter weights (26 HTK filters: 2 x 257 at most)

constexpr int kPRow = 268;            // floats between the power rows of consecutive tile frames (= 12 banks)

struct Smem {
  float2 tw[kNfft];
  float2 S[kFfts][kSFloat2];            // exchange storage; the power rows overwrite it
  float2 C[kFfts][kNfft];               // spectra; the tile's pre-emphasised samples (kSpan floats, zero past
                                        // the clip) live here until step 1 has consumed them
  float feat[kTileFrames * kMaxFilt];   // log energies [frame][nfilt] packed = the stacked row layout
  float wts[kMaxTaps];                  // filter weights, supports back to back
  int lo[kMaxFilt], hi[kMaxFilt], woff[kMaxFilt];   // filter supports and weight offsets
  float stat[kTileFrames][2];
};

// workspace: supports [kMaxFilt][2], weight offsets [kMaxFilt], total, packed weights [kMaxTaps]
struct FilterPack {
  int support[kMaxFilt][2];
  int woff[kMaxFilt];
  int total, pad[3];
  float wts[kMaxTaps];
};

// number of frames framesig makes of `len` samples (python_speech_features.sigproc.framesig)
__host__ __device__ inline int64_t num_frames(int64_t len) {
  return len <= kFrame ? 1 : 1 + (len - kFrame + kHop - 1) / kHop;
}

// filter supports [first, last + 1) of the nonzero weights and the weights packed back to back,
// once per call: one warp per filter, then a serial prefix over <= 64 lengths
__global__ void __launch_bounds__(256)
logfbank_prep_kernel(const float* __restrict__ fb, int nfilt, FilterPack* __restrict__ pack) {
  __shared__ int s_lo[kMaxFilt], s_hi[kMaxFilt], s_off[kMaxFilt + 1];
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
    if (lane == 0) { s_lo[m] = lo < hi ? lo : 0; s_hi[m] = lo < hi ? hi : 0; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int m = 0; m < nfilt; ++m) { s_off[m] = acc; acc += s_hi[m] - s_lo[m]; }
    s_off[nfilt] = acc;
    pack->total = acc;
  }
  __syncthreads();
  const bool fits = s_off[nfilt] <= kMaxTaps;
  for (int m = threadIdx.x >> 5; m < nfilt; m += blockDim.x >> 5) {
    if (lane == 0) { pack->support[m][0] = s_lo[m]; pack->support[m][1] = s_hi[m]; pack->woff[m] = s_off[m]; }
    if (fits)
      for (int k = s_lo[m] + lane; k < s_hi[m]; k += 32) pack->wts[s_off[m] + k - s_lo[m]] = fb[(size_t)m * kBins + k];
  }
}

// One wave of CTAs; the tiles of the whole (ragged) batch form one list that the CTAs stride over,
// so that the twiddle table and the filter supports are set up once per CTA and a long clip does not
// leave the CTAs of the short ones idle.  The list is virtual: clip b's tiles start at unit
// V_b = row_offsets[b] / rows_per_tile + b (monotone, and V_{b+1} - V_b >= the clip's tile count), a
// unit that falls into the slack between two clips is skipped; the clip of a unit is found by a
// binary search over row_offsets.
__global__ void __launch_bounds__(kThreads)
logfbank_kernel(const float* __restrict__ audio, const int64_t* __restrict__ offsets,
                const int64_t* __restrict__ row_offsets, int64_t B, const float* __restrict__ fb, int nfilt,
                const FilterPack* __restrict__ pack, int stack, int normalize, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fbk_smem[];
  Smem& sm = *reinterpret_cast<Smem*>(fbk_smem);
  const int tid = threadIdx.x;
  for (int i = tid; i < kNfft; i += kThreads) sm.tw[i] = make_float2(kTw512Re[i], kTw512Im[i]);
  if (tid < nfilt) { sm.lo[tid] = pack->support[tid][0]; sm.hi[tid] = pack->support[tid][1]; sm.woff[tid] = pack->woff[tid]; }
  const bool packed = pack->total <= kMaxTaps;
  if (packed)
    for (int i = tid; i < pack->total; i += kThreads) sm.wts[i] = pack->wts[i];
  // tile-invariant work split of the filterbank stage: item = (frame, filter)
  constexpr int kItems = (kTileFrames * kMaxFilt + kThreads - 1) / kThreads;   // 2
  int it_f[kItems], it_m[kItems];
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const int i = tid + q * kThreads;
    // the 8 frames of one filter sit in adjacent lanes: a warp holds 4 neighbouring filters, whose
    // supports have similar lengths (a warp runs as long as its widest filter), and the 8 power rows
    // are kPRow = 268 floats apart, i.e. 12 banks: same-bin reads of the 8 frames do not collide
    it_f[q] = (i < kTileFrames * nfilt) ? i % kTileFrames : -1;
    it_m[q] = (i < kTileFrames * nfilt) ? i / kTileFrames : 0;
  }
  const int rows_here = kTileFrames / stack, width = stack * nfilt;
  float* tile_y = reinterpret_cast<float*>(sm.C);               // dead between power_rows and the next step 3
  static_assert(kSpan * sizeof(float) <= sizeof(sm.C), "the tile's samples must fit in the spectrum storage");

  const int rpt_shift = 31 - __clz(rows_here);                   // rows per tile is 1, 2, 4 or 8 (8 % stack == 0)
  const int stack_shift = 31 - __clz(stack);                     // ... and so is stack: shifts, not 64-bit divisions per tile
  const int64_t units = (row_offsets[B] >> rpt_shift) + B;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
  int64_t b = 0;
  for (int64_t hi = B - 1; b < hi;) {                           // largest b with V_b <= u (CTA-uniform)
    const int64_t mid = (b + hi + 1) >> 1;
    if ((row_offsets[mid] >> rpt_shift) + mid <= u) b = mid; else hi = mid - 1;
  }
  const int64_t beg = offsets[b], len = offsets[b + 1] - beg;
  const int64_t nfr = num_frames(len);
  const int64_t rows = (nfr + stack - 1) >> stack_shift;        // stacked rows (zero-padded tail)
  const int64_t f0 = (u - ((row_offsets[b] >> rpt_shift) + b)) * kTileFrames;
  if (len <= 0 || f0 < 0 || f0 >= (rows << stack_shift)) continue;   // slack unit
  const float* clip = audio + beg;
  // raw samples of the span (one coalesced load each, plus the sample before the span), then the
  // pre-emphasis y[n] = x[n] - 0.97 x[n-1] from registers; beyond the clip: framesig's zeros
  {
    const int64_t s0 = f0 * kHop;
    constexpr int kPer = (kSpan + kThreads - 1) / kThreads;       // 6
    float cur[kPer], prev[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int64_t n = s0 + tid + q * kThreads;
      cur[q] = (n < len) ? clip[n] : 0.0f;
      prev[q] = (n >= 1 && n < len) ? clip[n - 1] : 0.0f;       // a neighbouring lane's line: L1 hit
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = tid + q * kThreads;
      const int64_t n = s0 + i;
      if (i < kSpan) tile_y[i] = (n == 0 || n >= len) ? cur[q] : __fsub_rn(cur[q], __fmul_rn(kPreemph, prev[q]));
    }
  }
  __syncthreads();

  const int g = tid / kFftThreads, t = tid % kFftThreads;       // FFT g: tile frames 2g, 2g + 1
  float2* S = sm.S[g];
  float2* C = sm.C[g];
  step1(t, tile_y + (2 * g) * kHop, tile_y + (2 * g + 1) * kHop, sm.tw, S);
  __syncthreads();
  float2 x[8];
  step2_load(t, S, x);
  __syncthreads();
  step2_store(t, sm.tw, x, S);
  __syncthreads();
  step3(t, S, C);
  __syncthreads();
  float* Pall = reinterpret_cast<float*>(sm.S);                 // the tile's power rows overwrite the exchange storage
  power_rows(t, C, Pall + (2 * g) * kPRow, Pall + (2 * g + 1) * kPRow);
  __syncthreads();

  // log filterbank energies: item = (frame, filter); frames past the clip's last one are the
  // zero rows extract_logfbank_features appends before stacking
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const int f = it_f[q], m = it_m[q];
    if (f < 0) continue;
    const float* Pf = Pall + f * kPRow;
    float v = 0.0f;
    if (f0 + f < nfr) {
      const int lo = sm.lo[m], hi = sm.hi[m];
      v = packed ? log_fbank(Pf + lo, sm.wts + sm.woff[m], 0, hi - lo) : log_fbank(Pf, fb + (size_t)m * kBins, lo, hi);
    }
    sm.feat[f * nfilt + m] = v;
  }
  __syncthreads();

  // rows of `stack` consecutive frames (contiguous in feat); audio_to_tensor:
  // (x - mean) / (std + 1e-5), population std
  if (normalize) {
    const int lane = tid & 31, wid = tid >> 5;
    for (int r = wid; r < rows_here; r += kThreads / 32) {
      const float* row = sm.feat + r * width;
      float sum = 0.0f;
      for (int i = lane; i < width; i += 32) sum += row[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / (float)width;
      float v = 0.0f;
      for (int i = lane; i < width; i += 32) {
        const float d = row[i] - mean;
        v = fmaf(d, d, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) { sm.stat[r][0] = mean; sm.stat[r][1] = sqrtf(v / (float)width) + 1e-5f; }
    }
    __syncthreads();
  }
  const int64_t row0 = f0 >> stack_shift;
  float* o = out + (row_offsets[b] + row0) * width;
  for (int r = 0; r < rows_here; ++r) {
    if (row0 + r >= rows) break;
    const float mean = normalize ? sm.stat[r][0] : 0.0f, sd = normalize ? sm.stat[r][1] : 1.0f;
    for (int c = tid; c < width; c += kThreads) {
      const float v = sm.feat[r * width + c];
      o[r * width + c] = normalize ? __fdiv_rn(v - mean, sd) : v;
    }
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

extern "C" size_t avfe_logfbank_workspace_bytes(void) { return sizeof(fbk::FilterPack); }

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
  if (!aligned16(workspace)) return AVFE_ERR_WORKSPACE;
  fbk::FilterPack* pack = static_cast<fbk::FilterPack*>(workspace);
  fbk::logfbank_prep_kernel<<<1, 256, 0, s>>>(fbank, nfilt, pack);
  if (cudaFuncSetAttribute(fbk::logfbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(fbk::Smem)) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  // one resident wave (5 CTAs per SM: shared memory), each CTA striding over the batch's tile list
  int resident = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, fbk::logfbank_kernel, fbk::kThreads, sizeof(fbk::Smem)) !=
          cudaSuccess || resident < 1) {
    cudaGetLastError();
    resident = 1;
  }
  int64_t ctas = (int64_t)resident * kNumSMs;
  if (ctas > tiles * B) ctas = tiles * B;
  if (ctas < 1) ctas = 1;
  fbk::logfbank_kernel<<<(unsigned)ctas, fbk::kThreads, sizeof(fbk::Smem), s>>>(audio, offsets, row_offsets, B, fbank, nfilt,
                                                                              pack, stack, normalize, out);
  count_launch(2);
  return check_launch();
}
