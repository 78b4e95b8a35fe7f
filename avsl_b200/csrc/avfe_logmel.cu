// Whisper log-mel spectrogram, batched.  Replaces whisper.log_mel_spectrogram as called at
// avsl/whisper_flamingo_ft_ami.py:209-213 (HF twin: avsl/whisper_ft.py:347-350), plus
// whisper.pad_or_trim (:209-210) and the peak normalisation of
// preprocess/audio_process.py:312-317.
//
//   logmel_prep_kernel      per-clip max keys <- -inf, mel filter supports (first/last nonzero)
//   logmel_tile_kernel      persistent CTAs; one tile = 32 frames: reflect-padded audio span
//                           staged once in shared memory, Hann window, 16 complex 400-point
//                           shared-memory FFTs (two real frames each), power, sparse mel
//                           projection, log10, raw store + per-clip atomic max
//   logmel_finalize_kernel  max(x, clip_max - 8), (x + 4) / 4 in place (output still in L2)
//
// Frames, spectra and powers never touch HBM: traffic is the audio read plus the output.
#include <limits.h>

#include "avfe_common.cuh"
#include "avfe_logmel_core.cuh"

namespace avfe {
namespace lm {

// hann[400] and tw[400] (tw[k1*20+j] = exp(-2*pi*i*j*k1/400)), generated in float64 by
// avsl_b200/build.py and rounded to float32.
#include "avfe_logmel_tables.inc"

constexpr int kMaxPacked = 1024;   // packed filter weights kept in shared memory (slaney: ~400-470)
constexpr int kMaxMels = 128;

// Filterbank in sparse form, built by logmel_prep_kernel in the workspace.
struct MelPack {
  int4 rec[kMaxMels];      // per filter: first bin, support length, offset of its weights in wts, quads
  int total;               // sum of quad-padded support lengths; > kMaxPacked => weights stay in global fb
  int pad[3];
  float wts[kMaxPacked];
};

struct Smem {
  float2 Z[kPairs * kZPair];            // 53,760 B  20x20 exchange / spectra, then the 32 power rows
  float audio[2][kTileFloats];          // 45,600 B  double-buffered reflect-padded audio span (skewed)
  float2 tw[kNfft];                     //  3,200 B
  float hann[kNfft];                    //  1,600 B
  __align__(16) float wts[kMaxPacked];  //  4,096 B  zero-padded to quads
  int4 rec[kMaxMels];                   //  2,048 B
  int red[16];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// block 0: sparse form of the filterbank; blocks 1..: per-clip max keys <- -inf
__global__ void __launch_bounds__(1024)
logmel_prep_kernel(const float* __restrict__ fb, int n_mels, int64_t B, int* __restrict__ clip_max,
                   MelPack* __restrict__ pack) {
  if (blockIdx.x > 0) {
    const int64_t i = (int64_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x;
    if (i < B) clip_max[i] = INT_MIN;
    return;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int m = wid; m < n_mels; m += 32) {                 // one warp per filter
    const float* row = fb + (size_t)m * kBins;
    int lo = kBins, hi = 0;
    for (int k = lane; k < kBins; k += 32)
      if (row[k] != 0.0f) { lo = min(lo, k); hi = max(hi, k + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) pack->rec[m] = (lo < hi) ? make_int4(lo, hi - lo, 0, (hi - lo + 3) >> 2) : make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int m = 0; m < n_mels; ++m) { pack->rec[m].z = acc; acc += 4 * pack->rec[m].w; }   // quads
    pack->total = acc;
  }
  __syncthreads();
  if (pack->total <= kMaxPacked) {
    for (int m = wid; m < n_mels; m += 32) {
      const int4 rc = pack->rec[m];
      for (int k = lane; k < 4 * rc.w; k += 32)
        pack->wts[rc.z + k] = (k < rc.y) ? fb[(size_t)m * kBins + rc.x + k] : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2)
logmel_tile_kernel(const float* __restrict__ audio, int64_t B, int64_t L, int64_t Lp,
                   int64_t n_frames, int n_mels, const float* __restrict__ fb,
                   const MelPack* __restrict__ pack, float* __restrict__ out,
                   int* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool packed = pack->total <= kMaxPacked;
  for (int i = tid; i < kNfft; i += kThreads) {
    sm.tw[i] = make_float2(kTwRe[i], kTwIm[i]);
    sm.hann[i] = kHann[i];
  }
  for (int i = tid; i < n_mels; i += kThreads) sm.rec[i] = pack->rec[i];
  if (packed)
    for (int i = tid; i < pack->total; i += kThreads) sm.wts[i] = pack->wts[i];
  // tiles are numbered clip-major; this CTA takes tiles blockIdx.x, +gridDim.x, ...; the
  // (clip, tile-in-clip) pair is advanced incrementally in 32-bit arithmetic
  const int tiles_per_clip = (int)((n_frames + kTileFrames - 1) / kTileFrames);
  const int g = tid / 20, j = tid % 20;
  const bool base_ok = ((reinterpret_cast<uintptr_t>(audio) & 15u) == 0) && (L % 4 == 0);
  const float floor_v = log10_floor(0.0f);               // value of every all-zero frame
  const int step_b = (int)(gridDim.x / (unsigned)tiles_per_clip);
  const int step_t = (int)(gridDim.x % (unsigned)tiles_per_clip);

  // Interior tiles (no reflection, no zero padding) are fetched with 16-byte cp.async one tile
  // ahead; clip-edge tiles are assembled sample by sample when their turn comes.
  auto is_fast = [&](int tt) -> bool {
    const int64_t p0 = (int64_t)tt * (kTileFrames * kHop);
    return base_ok && p0 >= kNfft / 2 && p0 - kNfft / 2 + kTileSamples <= L;
  };
  auto prefetch = [&](int64_t b, int tt, int buf) {
    if (b < B && is_fast(tt)) {
      const float* src = audio + b * L + ((int64_t)tt * (kTileFrames * kHop) - kNfft / 2);
      for (int i = tid; i < kTileSamples / 4; i += kThreads)
        cp_async16(&sm.audio[buf][tile_pos(4 * i)], src + 4 * i);     // skew is a multiple of 4 floats
    }
    cp_async_commit();
  };

  int64_t b = blockIdx.x / (unsigned)tiles_per_clip;
  int tt = (int)(blockIdx.x % (unsigned)tiles_per_clip);
  int cur = 0;
  prefetch(b, tt, 0);
  for (; b < B; cur ^= 1) {
    // next tile of this CTA
    int64_t bn = b + step_b;
    int tn = tt + step_t;
    if (tn >= tiles_per_clip) { tn -= tiles_per_clip; ++bn; }
    prefetch(bn, tn, cur ^ 1);
    cp_async_wait<1>();                                  // this tile's group has landed
    const int64_t t0 = (int64_t)tt * kTileFrames;
    float* au = sm.audio[cur];
    bool nz = false;
    if (is_fast(tt)) {
      for (int i = tid; i < kTileSamples / 4; i += kThreads) {   // the chunks this thread fetched
        const float4 q = *reinterpret_cast<const float4*>(&au[tile_pos(4 * i)]);
        nz |= (q.x != 0.0f) | (q.y != 0.0f) | (q.z != 0.0f) | (q.w != 0.0f);
      }
    } else {
      const float* clip = audio + b * L;
      const int64_t p0 = t0 * kHop;
      for (int i = tid; i < kTileSamples; i += kThreads) {
        const float q = padded_sample(clip, L, Lp, p0 + i);
        au[tile_pos(i)] = q;
        nz |= (q != 0.0f);
      }
    }
    // barrier: tile visible to all, previous tile's power rows no longer needed
    const int any = __syncthreads_or(nz ? 1 : 0);
    const int64_t t = t0 + lane;
    const bool live = t < n_frames;
    float vmax = -INFINITY;
    float* orow = out + (b * n_mels + wid) * n_frames + t;       // filter `wid`, this lane's frame
    const int64_t ostep = (int64_t)(kThreads / 32) * n_frames;
    if (!any) {
      // all 32 frames are digital silence: |X|^2 = 0 -> mel = 0 -> log10(1e-10); skip the FFTs
      if (live) {
        for (int m = wid; m < n_mels; m += kThreads / 32, orow += ostep) *orow = floor_v;
        vmax = floor_v;
      }
    } else {
      // ---- 16 complex FFT-400: columns, twiddle, rows ----
      stage1(g, j, au, sm.hann, sm.tw, sm.Z);
      __syncthreads();
      stage2(g, j, sm.Z);
      __syncthreads();
      // ---- untangle the two real frames of each FFT; power rows overwrite the Z storage ----
      float pa[kBinsPerThread], pb[kBinsPerThread];
      split_load(g, j, sm.Z, pa, pb);
      __syncthreads();
      float* P = reinterpret_cast<float*>(sm.Z);
      split_store(g, j, P, pa, pb);
      __syncthreads();
      // ---- sparse mel projection + log10: warp = filter, lane = frame ----
      const float* prow = P + prow_offset(lane);
      for (int m = wid; m < n_mels; m += kThreads / 32, orow += ostep) {
        const int4 rc = sm.rec[m];                       // first bin, length, weight offset, quads
        const float v = packed ? mel_log10_quads(prow + rc.x, reinterpret_cast<const float4*>(sm.wts + rc.z), rc.w)
                               : mel_log10(prow + rc.x, fb + (size_t)m * kBins + rc.x, rc.y);
        if (live) {
          *orow = v;
          vmax = fmaxf(vmax, v);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) sm.red[wid] = float_key(vmax);
    __syncthreads();
    if (tid == 0) {
      int k = sm.red[0];
      for (int w = 1; w < kThreads / 32; ++w) k = max(k, sm.red[w]);
      atomicMax(clip_max + b, k);
    }
    b = bn; tt = tn;
  }
  cp_async_wait<0>();
}

__global__ void __launch_bounds__(256)
logmel_finalize_kernel(float* __restrict__ out, const int* __restrict__ clip_max,
                       int64_t per_clip, int64_t total) {
  if ((per_clip & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    float4* o4 = reinterpret_cast<float4*>(out);
    const int64_t n4 = total >> 2, per4 = per_clip >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float floor_v = __fsub_rn(key_float(clip_max[i / per4]), 8.0f);
      float4 x = o4[i];
      x.x = __fmul_rn(__fadd_rn(fmaxf(x.x, floor_v), 4.0f), 0.25f);
      x.y = __fmul_rn(__fadd_rn(fmaxf(x.y, floor_v), 4.0f), 0.25f);
      x.z = __fmul_rn(__fadd_rn(fmaxf(x.z, floor_v), 4.0f), 0.25f);
      x.w = __fmul_rn(__fadd_rn(fmaxf(x.w, floor_v), 4.0f), 0.25f);
      o4[i] = x;
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float floor_v = __fsub_rn(key_float(clip_max[i / per_clip]), 8.0f);
    const float x = fmaxf(out[i], floor_v);
    out[i] = __fmul_rn(__fadd_rn(x, 4.0f), 0.25f);
  }
}

__global__ void __launch_bounds__(256)
pad_or_trim_kernel(const float* __restrict__ in, int64_t B, int64_t L_in, int64_t L_out,
                   float* __restrict__ out) {
  const int64_t total = B * L_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L_out, k = i % L_out;
    out[i] = (k < L_in) ? in[b * L_in + k] : 0.0f;
  }
}

// ragged variant: clip b = in[offsets[b] : offsets[b+1]]
__global__ void __launch_bounds__(256)
pad_or_trim_ragged_kernel(const float* __restrict__ in, const int64_t* __restrict__ offsets,
                          int64_t L_out, float* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const int64_t beg = offsets[b], len = offsets[b + 1] - beg;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < L_out;
       k += (int64_t)gridDim.x * blockDim.x)
    out[b * L_out + k] = (k < len) ? in[beg + k] : 0.0f;
}

// per-clip min/max -> scratch[2b], scratch[2b+1] as ordered int keys
__global__ void __launch_bounds__(256)
peak_minmax_kernel(const float* __restrict__ audio, int64_t L, int* __restrict__ keys) {
  const int64_t b = blockIdx.y;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = audio[b * L + i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(keys + 2 * b, float_key(mn));
    atomicMax(keys + 2 * b + 1, float_key(mx));
  }
}

__global__ void peak_init_kernel(int* __restrict__ keys, int64_t B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { keys[2 * i] = INT_MAX; keys[2 * i + 1] = INT_MIN; }
}

__global__ void __launch_bounds__(256)
peak_scale_kernel(const float* audio, int64_t L, const int* __restrict__ keys, float* out) {
  const int64_t b = blockIdx.y;
  const float mn = key_float(keys[2 * b]), mx = key_float(keys[2 * b + 1]);
  const bool scale = (mx > 1.0f) || (mn < -1.0f);
  const float d = fmaxf(fabsf(mx), fabsf(mn));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = audio[b * L + i];
    out[b * L + i] = scale ? __fdiv_rn(v, d) : v;
  }
}

}  // namespace lm
}  // namespace avfe

using namespace avfe;

extern "C" size_t avfe_logmel_workspace_bytes(int64_t B, int64_t L, int64_t padding, int n_mels) {
  (void)L; (void)padding;
  if (B < 0 || n_mels < 0) return 0;
  (void)n_mels;
  return (((size_t)B * sizeof(int) + 15) & ~(size_t)15) + sizeof(lm::MelPack) + 64;
}

extern "C" int avfe_logmel_f32(const float* audio, int64_t B, int64_t L, int64_t padding,
                               int n_mels, const float* mel_filters, float* out, void* workspace,
                               size_t workspace_bytes, avfe_stream_t stream) {
  if (B < 0 || L < 0 || padding < 0 || n_mels <= 0) return AVFE_ERR_INVALID_ARG;
  if (n_mels > 128) return AVFE_ERR_UNSUPPORTED;
  const int64_t Lp = L + padding;
  const int64_t n_frames = Lp / lm::kHop;
  if (B == 0 || n_frames == 0) return AVFE_OK;
  if (Lp <= lm::kNfft / 2) return AVFE_ERR_INVALID_ARG;   // reflect pad needs pad < length
  if (!audio || !mel_filters || !out) return AVFE_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < avfe_logmel_workspace_bytes(B, L, padding, n_mels) ||
      !aligned16(workspace))
    return AVFE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* clip_max = static_cast<int*>(workspace);
  lm::MelPack* pack = reinterpret_cast<lm::MelPack*>(static_cast<char*>(workspace) +
                                                     (((size_t)B * sizeof(int) + 15) & ~(size_t)15));

  lm::logmel_prep_kernel<<<(unsigned)(1 + (B + 1023) / 1024), 1024, 0, s>>>(mel_filters, n_mels, B,
                                                                           clip_max, pack);
  count_launch();

  // per-device attribute: set on every call (cheap) so multi-device processes stay correct
  if (cudaFuncSetAttribute(lm::logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(lm::Smem)) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  const int64_t n_tiles = B * ((n_frames + lm::kTileFrames - 1) / lm::kTileFrames);
  int64_t ctas = n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs;   // 2 resident CTAs per SM
  lm::logmel_tile_kernel<<<(unsigned)ctas, lm::kThreads, sizeof(lm::Smem), s>>>(
      audio, B, L, Lp, n_frames, n_mels, mel_filters, pack, out, clip_max);
  count_launch();

  const int64_t per_clip = (int64_t)n_mels * n_frames, total = B * per_clip;
  int64_t fin = (total / 4 + 255) / 256 + 1;
  if (fin > (int64_t)kNumSMs * 16) fin = (int64_t)kNumSMs * 16;
  lm::logmel_finalize_kernel<<<(unsigned)fin, 256, 0, s>>>(out, clip_max, per_clip, total);
  count_launch();
  return check_launch();
}

extern "C" int avfe_pad_or_trim_f32(const float* in, int64_t B, int64_t L_in, int64_t L_out,
                                    float* out, avfe_stream_t stream) {
  if (B < 0 || L_in < 0 || L_out < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || L_out == 0) return AVFE_OK;
  if (!out || (!in && L_in > 0)) return AVFE_ERR_INVALID_ARG;
  const int64_t total = B * L_out;
  int64_t ctas = (total + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  lm::pad_or_trim_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, B, L_in, L_out, out);
  count_launch();
  return check_launch();
}

extern "C" int avfe_pad_or_trim_ragged_f32(const float* in, const int64_t* offsets, int64_t B,
                                           int64_t L_out, float* out, avfe_stream_t stream) {
  if (B < 0 || L_out < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || L_out == 0) return AVFE_OK;
  if (!in || !offsets || !out) return AVFE_ERR_INVALID_ARG;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  int64_t gx = (L_out + 256 * 4 - 1) / (256 * 4);
  if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
  dim3 grid((unsigned)gx, (unsigned)B);
  lm::pad_or_trim_ragged_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, offsets,
                                                                                    L_out, out);
  count_launch();
  return check_launch();
}

extern "C" int avfe_peak_normalize_f32(const float* audio, int64_t B, int64_t L, float* out,
                                       float* scratch, avfe_stream_t stream) {
  if (B < 0 || L < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || L == 0) return AVFE_OK;
  if (!audio || !out || !scratch) return AVFE_ERR_INVALID_ARG;
  if (B > 65535) return AVFE_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* keys = reinterpret_cast<int*>(scratch);
  lm::peak_init_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(keys, B);
  int64_t gx = (L + 256 * 8 - 1) / (256 * 8);
  if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
  dim3 grid((unsigned)gx, (unsigned)B);
  lm::peak_minmax_kernel<<<grid, 256, 0, s>>>(audio, L, keys);
  lm::peak_scale_kernel<<<grid, 256, 0, s>>>(audio, L, keys, out);
  count_launch(3);
  return check_launch();
}
