// AV-HuBERT audio features: python_speech_features.logfbank(audio, samplerate=16000) (26 log
// mel-filterbank energies per 25 ms / 10 ms frame), frame stacking and per-row normalisation.
// Replaces extract_logfbank_features + audio_to_tensor, preprocess/audio_process.py:152-197
// (same code in utils/data_loading.py:181-201); called from process_audio_for_av_hubert (:199-236)
// and process_audio_dual_encoder (:267-299).
//
//   logfbank_kernel   one CTA pass = 16 consecutive frames of one clip: pre-emphasis (float32, as
//                     numpy evaluates it) while staging the audio span; every warp transforms one
//                     pair of frames (complex 512-point FFT = 16 x 32, register-resident 16-point
//                     codelets, one shared-memory exchange, avfe_logfbank_core.cuh) and leaves the
//                     pair's power rows in its own exchange slot; filters are dealt to the warps by
//                     weight (prep kernel), a warp's lanes = 16 frames x the two halves of a filter's
//                     support; log, stack `stack` frames per row, (x - mean) / (std + 1e-5)
//
// Frames, spectra and filterbank energies never touch HBM: traffic is the audio read plus the
// [rows, 26 * stack] output.
#include "avfe_common.cuh"
#include "avfe_logfbank_core.cuh"

namespace avfe {
namespace fbk {

#include "avfe_logfbank_tables.inc"   // kTw512Re / kTw512Im: exp(-2 pi i j / 512), float64 -> float32

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;                          // 256
constexpr int kTileFrames = 2 * kWarps;                        // 16: one pair of frames per warp
constexpr int kSpan = (kTileFrames - 1) * kHop + kFrame;       // 2800 samples per tile
constexpr int kMaxFilt = 40;             // filters per bank (python_speech_features: 26; fbank-40 fits)
constexpr int kMaxTaps = 768;            // packed filter weights, each half-support zero-padded to quads (26 HTK filters: 572)

struct Smem {
  float slots[kWarps * kSlotFloats];     // per warp: S exchange, then U (tail) and the pair's power rows (head)
  float y[kSpan];                        // the tile's pre-emphasised samples, zero past the clip
  float2 tw1[16 * 32];                   // W512^(l r), lane-contiguous rows
  float feat[kTileFrames * kMaxFilt];    // log energies [frame][nfilt] packed = the stacked row layout
  __align__(16) float wts[kMaxTaps];     // filter weights: per filter the two half-supports, each zero-padded to quads
  int4 rec[kMaxFilt];                    // first bin, weight offset, quads of the first / second half
  int lo[kMaxFilt], hi[kMaxFilt];        // filter supports (dense filterbanks: weights stay in global memory)
  unsigned char wf[kWarps][kMaxFilt];    // the filters of each warp ...
  int wf_cnt[kWarps];                    // ... and how many
};

// workspace: supports, weight offsets, the warps' filter lists, packed weights
struct FilterPack {
  int support[kMaxFilt][2];
  int4 rec[kMaxFilt];                    // first bin, weight offset, quads of the first / second half
  int total, pad[3];                     // total = padded weights; > kMaxTaps: not packed
  int wf_cnt[kWarps];
  unsigned char wf[kWarps][kMaxFilt];
  float wts[kMaxTaps];
};

// number of frames framesig makes of `len` samples (python_speech_features.sigproc.framesig)
__host__ __device__ inline int64_t num_frames(int64_t len) {
  return len <= kFrame ? 1 : 1 + (len - kFrame + kHop - 1) / kHop;
}

// filter supports [first, last + 1) of the nonzero weights and the weights packed back to back,
// once per call: one warp per filter, then a serial prefix over <= 40 lengths and the deal of the
// filters to the tile kernel's warps (longest support first, always to the least loaded warp)
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
  // half-supports in quads: the first half takes ceil(n / 2) taps rounded up to a quad
  auto q0_of = [](int n) { return ((n + 1) / 2 + 3) >> 2; };
  auto q1_of = [&](int n) { const int rest = n - 4 * q0_of(n); return rest > 0 ? (rest + 3) >> 2 : 0; };
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int m = 0; m < nfilt; ++m) { const int n = s_hi[m] - s_lo[m]; s_off[m] = acc; acc += 4 * (q0_of(n) + q1_of(n)); }
    s_off[nfilt] = acc;
    pack->total = acc;
    // longest-processing-time deal: cost of a filter = the quads of its longer half + a constant
    int load[kWarps], cnt[kWarps];
    unsigned long long done = 0ull;
    for (int w = 0; w < kWarps; ++w) { load[w] = 0; cnt[w] = 0; }
    for (int i = 0; i < nfilt; ++i) {
      int best = -1, best_len = -1;
      for (int m = 0; m < nfilt; ++m)
        if (!((done >> m) & 1ull) && s_hi[m] - s_lo[m] > best_len) { best = m; best_len = s_hi[m] - s_lo[m]; }
      done |= 1ull << best;
      int w = 0;
      for (int v = 1; v < kWarps; ++v) if (load[v] < load[w]) w = v;
      pack->wf[w][cnt[w]++] = (unsigned char)best;
      load[w] += 9 * q0_of(best_len) + 12;
    }
    for (int w = 0; w < kWarps; ++w) pack->wf_cnt[w] = cnt[w];
  }
  __syncthreads();
  const bool fits = s_off[nfilt] <= kMaxTaps;
  for (int m = threadIdx.x >> 5; m < nfilt; m += blockDim.x >> 5) {
    const int n = s_hi[m] - s_lo[m], q0 = q0_of(n), q1 = q1_of(n);
    if (lane == 0) {
      pack->support[m][0] = s_lo[m];
      pack->support[m][1] = s_hi[m];
      pack->rec[m] = make_int4(s_lo[m], s_off[m], q0, q1);
    }
    if (fits)
      for (int k = lane; k < 4 * (q0 + q1); k += 32) pack->wts[s_off[m] + k] = (k < n) ? fb[(size_t)m * kBins + s_lo[m] + k] : 0.0f;
  }
}

// One resident wave of CTAs; the tiles of the whole (ragged) batch form one list, of which every CTA
// takes a contiguous run (consecutive tiles of a clip overlap by 240 samples, and the clip of the
// next tile is the same or one of the next few).  The list is virtual: clip b's tiles start at unit
// V_b = row_offsets[b] / rows_per_tile + b (monotone, and V_{b+1} - V_b >= the clip's tile count); a
// unit that falls into the slack between two clips is skipped.
__global__ void __launch_bounds__(kThreads, 4)
logfbank_kernel(const float* __restrict__ audio, const int64_t* __restrict__ offsets,
                const int64_t* __restrict__ row_offsets, int64_t B, const float* __restrict__ fb, int nfilt,
                const FilterPack* __restrict__ pack, int stack, int normalize, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fbk_smem[];
  Smem& sm = *reinterpret_cast<Smem*>(fbk_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 16 * 32; i += kThreads) {
    const int e = tw1_index(i >> 5, i & 31);
    sm.tw1[i] = make_float2(kTw512Re[e], kTw512Im[e]);
  }
  if (tid < nfilt) { sm.lo[tid] = pack->support[tid][0]; sm.hi[tid] = pack->support[tid][1]; sm.rec[tid] = pack->rec[tid]; }
  if (tid < kWarps) sm.wf_cnt[tid] = pack->wf_cnt[tid];
  for (int i = tid; i < kWarps * kMaxFilt; i += kThreads) sm.wf[i / kMaxFilt][i % kMaxFilt] = pack->wf[i / kMaxFilt][i % kMaxFilt];
  const bool packed = pack->total <= kMaxTaps;
  if (packed)
    for (int i = tid; i < pack->total; i += kThreads) sm.wts[i] = pack->wts[i];

  const int rows_here = kTileFrames / stack, width = stack * nfilt;
  const int rpt_shift = 31 - __clz(rows_here);                   // rows per tile and stack are powers of two:
  const int stack_shift = 31 - __clz(stack);                     // shifts, not 64-bit divisions per tile
  const int64_t units = (row_offsets[B] >> rpt_shift) + B;
  const int64_t per = (units + gridDim.x - 1) / gridDim.x;
  const int64_t u_begin = (int64_t)blockIdx.x * per, u_end = min(units, u_begin + per);
  auto first_unit = [&](int64_t c) -> int64_t { return (row_offsets[c] >> rpt_shift) + c; };
  int64_t b = 0;
  for (int64_t hi = B - 1; b < hi;) {                           // largest b with V_b <= u_begin, once per CTA
    const int64_t mid = (b + hi + 1) >> 1;
    if (first_unit(mid) <= u_begin) b = mid; else hi = mid - 1;
  }
  float* slot = sm.slots + wid * kSlotFloats;                    // this warp's exchange slot
  __syncthreads();

  for (int64_t u = u_begin; u < u_end; ++u) {
    while (b + 1 < B && first_unit(b + 1) <= u) ++b;             // CTA-uniform; the run is contiguous: a step or two
    const int64_t beg = offsets[b], len = offsets[b + 1] - beg;
    const int64_t nfr = num_frames(len);
    const int64_t rows = (nfr + stack - 1) >> stack_shift;      // stacked rows (zero-padded tail)
    const int64_t f0 = (u - first_unit(b)) * kTileFrames;
    if (len <= 0 || f0 >= (rows << stack_shift)) continue;       // slack unit
    const float* clip = audio + beg;
    // raw samples of the span (one coalesced load each, plus the sample before), then the
    // pre-emphasis y[n] = x[n] - 0.97 x[n-1] from registers; beyond the clip: framesig's zeros
    {
      const int64_t s0 = f0 * kHop;
      const int64_t left = len - s0;                             // samples of the clip from the span's start on
      const float* src = clip + s0;
      constexpr int kPer = (kSpan + kThreads - 1) / kThreads;    // 11
      if (s0 > 0 && left >= kSpan) {                             // interior tile (CTA-uniform): nothing to test per sample
        const float* mine = src + tid;
#pragma unroll
        for (int q0 = 0; q0 < kPer; q0 += 4) {                   // four samples per thread in flight
          float cur[4], prev[4];
#pragma unroll
          for (int q = q0; q < q0 + 4 && q < kPer; ++q)
            if (q < kPer - 1 || tid + q * kThreads < kSpan) {
              cur[q - q0] = mine[q * kThreads];
              prev[q - q0] = mine[q * kThreads - 1];             // a neighbouring lane's line: L1 hit
            }
#pragma unroll
          for (int q = q0; q < q0 + 4 && q < kPer; ++q)
            if (q < kPer - 1 || tid + q * kThreads < kSpan)
              sm.y[tid + q * kThreads] = __fsub_rn(cur[q - q0], __fmul_rn(kPreemph, prev[q - q0]));
        }
      } else {
#pragma unroll 1
        for (int i = tid; i < kSpan; i += kThreads) {
          const bool in = i < left;                              // also false for every i when left <= 0
          const float cur = in ? src[i] : 0.0f;
          const float prev = (in && (i > 0 || s0 > 0)) ? src[i - 1] : 0.0f;
          sm.y[i] = __fsub_rn(cur, __fmul_rn(kPreemph, prev));   // clip[0] - 0.97 * 0 = clip[0]; 0 - 0 past the end
        }
      }
    }
    __syncthreads();                                             // span visible (and the previous tile's rows are out)

    // ---- this warp's pair of frames: FFT, untangle, power rows into the head of its own slot ----
    {
      float2* S = reinterpret_cast<float2*>(slot);
      float2* U = reinterpret_cast<float2*>(slot + kUOffset);
      fft_stage1(lane, sm.y + (2 * wid) * kHop, sm.y + (2 * wid + 1) * kHop, sm.tw1, S);
      __syncwarp();
      float2 x[16];
      fft_stage2(lane, S, x);
      __syncwarp();                                              // every lane has read S: U may overwrite its tail
      fft_upper_store(lane, x, U);
      __syncwarp();
      fft_power(lane, x, U, slot, slot + kPRow);                 // rows and U do not overlap
    }
    __syncthreads();                                             // all 16 power rows written; the span is dead

    // ---- log filterbank energies: warp = some filters (dealt by weight), lane = (frame, half of the
    // support); frames past the clip's last one are the zero rows extract_logfbank_features appends ----
    {
      const int f = lane & 15, h = lane >> 4;
      const float* Pf = sm.slots + (f >> 1) * kSlotFloats + (f & 1) * kPRow;
      const bool live = f0 + f < nfr;
      const int cnt = sm.wf_cnt[wid];
      if (packed) {
        // per filter: one record, then quads of (LDS.128 weights, 4 LDS powers, 4 FFMA); the padded
        // taps carry zero weights and read finite slots (bins 257..260 of a row are zeroed)
        for (int i = 0; i < cnt; ++i) {
          const int m = sm.wf[wid][i];
          const int4 rc = sm.rec[m];                             // first bin, weight offset, quads of either half
          const int skip = h ? 4 * rc.z : 0, quads = h ? rc.w : rc.z;
          const float4* w4 = reinterpret_cast<const float4*>(sm.wts + rc.y + skip);
          const float* p = Pf + rc.x + skip;
          float2 a01 = make_float2(0.0f, 0.0f);              // two accumulator chains as one packed fp32x2 chain (FFMA2)
#pragma unroll 2
          for (int q = 0; q < quads; ++q) {
            const float4 c = w4[q];
            a01 = __ffma2_rn(make_float2(c.x, c.y), make_float2(p[4 * q], p[4 * q + 1]), a01);
            a01 = __ffma2_rn(make_float2(c.z, c.w), make_float2(p[4 * q + 2], p[4 * q + 3]), a01);
          }
          float acc = a01.x + a01.y;
          acc += __shfl_xor_sync(0xffffffffu, acc, 16);
          if (h == 0) sm.feat[f * nfilt + m] = live ? log_energy(acc) : 0.0f;
        }
      } else {                                                   // dense filterbank: weights from global memory
        for (int i = 0; i < cnt; ++i) {
          const int m = sm.wf[wid][i];
          const int lo = sm.lo[m], n = sm.hi[m] - lo, n0 = (n + 1) >> 1;
          const float* w = fb + (size_t)m * kBins + lo;
          const float* p = Pf + lo;
          float acc = 0.0f;
          for (int k = h ? n0 : 0; k < (h ? n : n0); ++k) acc = fmaf(__ldg(w + k), p[k], acc);
          acc += __shfl_xor_sync(0xffffffffu, acc, 16);
          if (h == 0) sm.feat[f * nfilt + m] = live ? log_energy(acc) : 0.0f;
        }
      }
    }
    __syncthreads();                                             // feat complete; the power rows are dead

    // ---- rows of `stack` consecutive frames (contiguous in feat); audio_to_tensor:
    // (x - mean) / (std + 1e-5), population std; one warp per row, which also writes it out ----
    {
      const int64_t row0 = f0 >> stack_shift;
      float* o = out + (row_offsets[b] + row0) * width;
      for (int r = wid; r < rows_here; r += kWarps) {
        if (row0 + r >= rows) break;
        const float* row = sm.feat + r * width;
        float mean = 0.0f, sd = 1.0f;
        if (normalize) {
          float sum = 0.0f;
          for (int i = lane; i < width; i += 32) sum += row[i];
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
          mean = sum / (float)width;
          float v = 0.0f;
          for (int i = lane; i < width; i += 32) {
            const float d = row[i] - mean;
            v = fmaf(d, d, v);
          }
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
          sd = __frcp_rn(sqrtf(v / (float)width) + 1e-5f);      // one reciprocal per row: 1 ulp from the division, bar 1e-3
        }
        for (int c = lane; c < width; c += 32) o[r * width + c] = normalize ? (row[c] - mean) * sd : row[c];
      }
    }
    // no barrier here: the next tile's span store touches only y (dead since the second barrier), and
    // its first barrier orders this tile's feat reads before the next filter stage's writes
  }
}

}  // namespace fbk
}  // namespace avfe

using namespace avfe;

extern "C" int64_t avfe_logfbank_num_frames(int64_t n_samples) {
  return n_samples < 0 ? 0 : fbk::num_frames(n_samples);
}

extern "C" size_t avfe_logfbank_workspace_bytes(void) { return sizeof(fbk::FilterPack); }

static int logfbank_check(int64_t B, int64_t max_samples, int nfilt, int stack) {
  if (B < 0 || nfilt <= 0 || stack <= 0 || max_samples < 0) return AVFE_ERR_INVALID_ARG;
  if (nfilt > fbk::kMaxFilt || (fbk::kTileFrames % stack) != 0 || (stack & (stack - 1)) != 0) return AVFE_ERR_UNSUPPORTED;
  return AVFE_OK;
}

extern "C" int avfe_logfbank_prepare(const float* fbank, int nfilt, void* workspace, size_t workspace_bytes,
                                     avfe_stream_t stream) {
  if (nfilt <= 0) return AVFE_ERR_INVALID_ARG;
  if (nfilt > fbk::kMaxFilt) return AVFE_ERR_UNSUPPORTED;
  if (!fbank) return AVFE_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < avfe_logfbank_workspace_bytes() || !aligned16(workspace)) return AVFE_ERR_WORKSPACE;
  fbk::logfbank_prep_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(fbank, nfilt, static_cast<fbk::FilterPack*>(workspace));
  count_launch();
  return check_launch();
}

extern "C" int avfe_logfbank_prepared_f32(const float* audio, const int64_t* offsets, const int64_t* row_offsets,
                                          int64_t B, int64_t max_samples, const float* fbank, int nfilt,
                                          int stack, int normalize, float* out, const void* workspace,
                                          size_t workspace_bytes, avfe_stream_t stream) {
  const int rc = logfbank_check(B, max_samples, nfilt, stack);
  if (rc != AVFE_OK) return rc;
  if (B == 0) return AVFE_OK;
  if (!audio || !offsets || !row_offsets || !fbank || !out) return AVFE_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < avfe_logfbank_workspace_bytes() || !aligned16(workspace)) return AVFE_ERR_WORKSPACE;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  const int64_t nfr = fbk::num_frames(max_samples);
  const int64_t padded = (nfr + stack - 1) / stack * stack;
  const int64_t tiles = (padded + fbk::kTileFrames - 1) / fbk::kTileFrames;
  if (tiles > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const fbk::FilterPack* pack = static_cast<const fbk::FilterPack*>(workspace);
  if (cudaFuncSetAttribute(fbk::logfbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(fbk::Smem)) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  // one resident wave (4 CTAs per SM), each CTA taking a contiguous run of the batch's tile list
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
  count_launch();
  return check_launch();
}

extern "C" int avfe_logfbank_f32(const float* audio, const int64_t* offsets, const int64_t* row_offsets,
                                 int64_t B, int64_t max_samples, const float* fbank, int nfilt,
                                 int stack, int normalize, float* out, void* workspace,
                                 size_t workspace_bytes, avfe_stream_t stream) {
  int rc = logfbank_check(B, max_samples, nfilt, stack);
  if (rc != AVFE_OK) return rc;
  if (B == 0) return AVFE_OK;
  if (!audio || !offsets || !row_offsets || !fbank || !out) return AVFE_ERR_INVALID_ARG;
  rc = avfe_logfbank_prepare(fbank, nfilt, workspace, workspace_bytes, stream);
  if (rc != AVFE_OK) return rc;
  return avfe_logfbank_prepared_f32(audio, offsets, row_offsets, B, max_samples, fbank, nfilt, stack, normalize, out,
                                    workspace, workspace_bytes, stream);
}
