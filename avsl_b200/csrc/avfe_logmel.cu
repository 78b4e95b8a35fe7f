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
//   logmel_finalize_tiles_kernel  the clamp max(x, clip_max - 8), only for the tiles that need it
//
// Frames, spectra and powers never touch HBM: traffic is the audio read plus the output.
#include <limits.h>

#include <type_traits>

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
  int max_quads;           // longest support in quads
  int balanced;            // 1 if assign[] is valid (every filter <= 4 quads, n_mels <= 128)
  int pad[1];
  // mel work split for the tile kernel: thread i computes filter (a & 0xff) for (a >> 16) frames
  // starting at tile frame ((a >> 8) & 0xff); threads are grouped by quad count so warps do not
  // diverge, and every thread gets ~12-16 (quad, frame) units
  unsigned assign[kThreads];
  float wts[kMaxPacked];
};

// 65.9 KB per CTA: three CTAs (30 warps) per SM.  The Z storage is used three times per tile: first
// as the landing zone of the reflect-padded audio span (5,700 skewed floats), then -- once every
// thread holds its windowed samples in registers -- as the 20x20 exchange, finally for the 32 power
// rows (first 26 KB) and the staged outputs (behind them).
constexpr int kStageOutOffset = ((kTileFrames * kPStride + 3) / 4) * 4;      // floats; 16-byte aligned
static_assert(kTileFloats * 4 <= kPairs * kZPair * 8, "audio span fits the Z storage");
// staged outputs [n_mels][kSoStride]: lanes that hold DIFFERENT filters write the same frame column,
// so the row stride must be odd (with 32 they all hit one bank: 5.0 M of the 13.3 M excess
// shared-memory wavefronts of a 64 x 30 s launch, profiles/r02/logmel_r2a_*).
constexpr int kSoStride = kTileFrames + 1;
static_assert((kStageOutOffset + kMaxMels * kSoStride) * 4 <= kPairs * kZPair * 8, "power rows + staged outputs fit the Z storage");
static_assert(kTileFloats <= kStageOutOffset, "the next tile's audio never lands on outputs that are still being written");
struct Smem {
  float2 Z[kPairs * kZPair];            // 53,760 B  audio span / 20x20 exchange / power rows + staged outputs
  float2 tw[kNfft];                     //  3,200 B
  float hann[kNfft];                    //  1,600 B
  __align__(16) float wts[kMaxPacked];  //  4,096 B  zero-padded to quads
  int4 rec[kMaxMels];                   //  2,048 B
  unsigned assign[kThreads];            //  1,280 B
  unsigned long long bar;               // completion of the bulk copies of the audio span
  int red[14];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
// mbarrier + 1-D bulk copy (TMA, UBLKCP): the aligned interior tiles' audio span is fetched as 17
// contiguous pieces of 320 samples, each landing at its skewed position -- no per-thread copy
// instructions and none of the bank conflicts the 16-byte cp.async form has on the skewed span
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nLM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra LM_WAIT;\n}" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// block 0: sparse form of the filterbank; blocks 1..: per-clip max keys <- -inf
__global__ void __launch_bounds__(1024)
logmel_prep_kernel(const float* __restrict__ fb, int n_mels, int64_t B, int* __restrict__ clip_max,
                   MelPack* __restrict__ pack) {
  (void)B; (void)clip_max;
  if (blockIdx.x > 0) return;
  // filter supports, weight offsets and the balanced mel work table, all in parallel over the
  // filters (a serial thread would crawl through dependent shared-memory round trips)
  __shared__ int s_lo[kMaxMels], s_len[kMaxMels], s_q[kMaxMels], s_woff[kMaxMels];
  __shared__ int s_tier_total[4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 4) s_tier_total[tid] = 0;
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
    if (lane == 0) {
      const int len = lo < hi ? hi - lo : 0;
      s_lo[m] = lo < hi ? lo : 0; s_len[m] = len; s_q[m] = (len + 3) >> 2;
    }
  }
  __syncthreads();
  // threads per filter under tier r: ~16 (quad, frame) units per thread first, coarser tiers if
  // the CTA's 320 threads do not suffice
  auto tm_of = [](int q, int tier) -> int {
    switch (tier) {
      case 0:  return q <= 1 ? 2 : (q == 2 ? 4 : 8);
      case 1:  return q <= 1 ? 2 : (q == 4 ? 8 : 4);
      case 2:  return q <= 1 ? 1 : (q == 2 ? 2 : 4);
      default: return 1;
    }
  };
  int total = 0, mq = 0;
  if (tid < n_mels) {
    int woff = 0;
    for (int m = 0; m < n_mels; ++m) {                     // every thread scans all filters: n_mels <= 128
      if (m < tid) woff += 4 * s_q[m];
      total += 4 * s_q[m];
      mq = max(mq, s_q[m]);
    }
    s_woff[tid] = woff;
#pragma unroll
    for (int r = 0; r < 4; ++r) atomicAdd(&s_tier_total[r], tm_of(s_q[tid], r));
  }
  __syncthreads();
  int tier = 3;
  for (int r = 3; r >= 0; --r) if (s_tier_total[r] <= kThreads) tier = r;
  const bool ok = (n_mels <= kMaxMels) && (n_mels <= kThreads);
  if (tid < n_mels) {
    // max quads over all filters (every thread computed it) decides whether the table is usable
    const bool balanced = ok && mq <= 4;
    const int q = s_q[tid], tmm = tm_of(q, tier);
    int off = 0;                                           // threads before this filter: larger quad counts first
    for (int m = 0; m < n_mels; ++m) {
      const int qm = s_q[m];
      if (qm > q || (qm == q && m < tid)) off += tm_of(qm, tier);
    }
    const int nf = kTileFrames / tmm;
    if (balanced)
      for (int i = 0; i < tmm; ++i)
        pack->assign[off + i] = (unsigned)tid | ((unsigned)(i * nf) << 8) | ((unsigned)nf << 16);
    pack->rec[tid] = make_int4(s_lo[tid], s_len[tid], s_woff[tid], q);
    if (tid == 0) { pack->total = total; pack->max_quads = mq; pack->balanced = balanced ? 1 : 0; }
  }
  // unused tail of the table
  for (int i = s_tier_total[tier] + tid; i < kThreads; i += blockDim.x) pack->assign[i] = 0u;
  // packed, quad-padded weights
  int all = 0;
  for (int m = 0; m < n_mels; ++m) all += 4 * s_q[m];
  if (all <= kMaxPacked) {
    for (int m = wid; m < n_mels; m += 32) {
      const int off = s_woff[m], len = s_len[m], lo = s_lo[m];
      for (int k = lane; k < 4 * s_q[m]; k += 32)
        pack->wts[off + k] = (k < len) ? fb[(size_t)m * kBins + lo + k] : 0.0f;
    }
  }
}

// clip b = audio[b*L : (b+1)*L] (dense batch) or audio[offsets[b] : offsets[b+1]] cut to Lp
// samples (ragged batch: pad_or_trim fused in; the zero padding is never materialised)
__device__ __forceinline__ const float* clip_of(const float* audio, const int64_t* offsets, int64_t b,
                                                int64_t L, int64_t Lp, int64_t& len) {
  if (offsets != nullptr) {
    const int64_t o = offsets[b];
    len = min(offsets[b + 1] - o, Lp);
    return audio + o;
  }
  len = L;
  return audio + b * L;
}

// Tile classes.  kSilent: every sample the tile touches (reflection included) lies in the zero
// padding, known from the clip length alone - nothing is read.  kFast16 / kFast4: interior tile
// (no reflection, no padding) fetched one tile ahead with 16- or 4-byte cp.async.  kEdge:
// assembled sample by sample when its turn comes.
enum { kSilent = 0, kFast16 = 1, kFast4 = 2, kEdge = 3 };
__device__ __forceinline__ int classify_tile(const float* clip, int64_t len, int64_t Lp, int tt) {
  const int64_t lo = (int64_t)tt * (kTileFrames * kHop) - kNfft / 2, hi = lo + kTileSamples;
  if (lo >= len && (hi <= Lp || 2 * (Lp - 1) - (hi - 1) >= len)) return kSilent;
  if (lo >= 0 && hi <= len)
    return ((reinterpret_cast<uintptr_t>(clip + lo) & 15u) == 0) ? kFast16 : kFast4;
  return kEdge;
}

// One warp per clip: list the tiles that have to be computed (everything that is not silent
// by length), flag the others for logmel_finalize_kernel, and start the clip maximum at the
// silent value when the clip has silent tiles (else at "-inf").
__global__ void __launch_bounds__(256)
logmel_live_kernel(const float* __restrict__ audio, const int64_t* __restrict__ offsets, int64_t B,
                   int64_t L, int64_t Lp, int tiles_per_clip, int* __restrict__ clip_max,
                   uint8_t* __restrict__ silent, int* __restrict__ live_count,
                   int2* __restrict__ live_list, int* __restrict__ tile_min) {
  // programmatic dependent launch: the tile kernel behind this one loads its tables meanwhile and
  // waits (griddepcontrol.wait) only in front of its first read of the live list
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  int64_t len;
  const float* clip = clip_of(audio, offsets, b, L, Lp, len);
  int n_live = 0;
  for (int t0 = 0; t0 < tiles_per_clip; t0 += 32) {
    const int tt = t0 + lane;
    const bool live = tt < tiles_per_clip && classify_tile(clip, len, Lp, tt) != kSilent;
    n_live += __popc(__ballot_sync(0xffffffffu, live));
  }
  int base = 0;
  if (lane == 0 && n_live) base = atomicAdd(live_count, n_live);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int t0 = 0; t0 < tiles_per_clip; t0 += 32) {
    const int tt = t0 + lane;
    const bool live = tt < tiles_per_clip && classify_tile(clip, len, Lp, tt) != kSilent;
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if (tt < tiles_per_clip) {
      silent[b * tiles_per_clip + tt] = live ? 0 : 1;
      tile_min[b * tiles_per_clip + tt] = 0x7f7fffff;          // key of FLT_MAX: the tile kernel lowers it
    }
    if (live) live_list[base + __popc(m & ((1u << lane) - 1u))] = make_int2((int)b, tt);
    base += __popc(m);
  }
  // 0x80808080 as an ordered key is about -3.4e38: below every log10 value
  if (lane == 0) clip_max[b] = (n_live < tiles_per_clip) ? float_key(log10_floor(0.0f)) : (int)0x80808080;
}

__device__ __forceinline__ float normalised(float x);

__global__ void __launch_bounds__(kThreads, 3)
logmel_tile_kernel(const float* __restrict__ audio, const int64_t* __restrict__ offsets, int64_t B,
                   int64_t L, int64_t Lp, int64_t n_frames, int n_mels, const float* __restrict__ fb,
                   const MelPack* __restrict__ pack, float* __restrict__ out,
                   int* __restrict__ clip_max, uint8_t* __restrict__ silent,
                   const int* __restrict__ live_count, const int2* __restrict__ live_list,
                   int* __restrict__ tile_min, int tiles_per_clip) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool packed = pack->total <= kMaxPacked;
  for (int i = tid; i < kNfft; i += kThreads) {
    sm.tw[i] = make_float2(kTwRe[i], kTwIm[i]);
    sm.hann[i] = kHann[i];
  }
  for (int i = tid; i < n_mels; i += kThreads) sm.rec[i] = pack->rec[i];
  sm.assign[tid] = pack->assign[tid];
  if (packed)
    for (int i = tid; i < pack->total; i += kThreads) sm.wts[i] = pack->wts[i];
  // tiles are numbered clip-major; this CTA takes tiles blockIdx.x, +gridDim.x, ...; the
  // (clip, tile-in-clip) pair is advanced incrementally in 32-bit arithmetic
  const int g = tid / 20, j = tid % 20;
  const bool mel_fast = packed && pack->balanced != 0;
  asm volatile("griddepcontrol.wait;" ::: "memory");      // logmel_live_kernel (launched in front of this kernel) has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finalize grid may be set up; it waits for this one to complete
  const int n_live = *live_count;                        // tiles to compute (logmel_live_kernel)

  // (clip, tile) records are read from the live list one tile ahead
  const int2 none = make_int2(-1, 0);
  const int stride = (int)gridDim.x;
  auto list_at = [&](int item) -> int2 { return item < n_live ? live_list[item] : none; };
  float* au = reinterpret_cast<float*>(sm.Z);            // the audio span shares the Z storage
  float* P = reinterpret_cast<float*>(sm.Z);             // so do the power rows ...
  float* stage_out = P + kStageOutOffset;                // ... and the staged outputs [n_mels][32]

  // (clip, tile) record, clip pointer / length and tile class are looked up ONE TILE AHEAD, so that
  // their dependent global loads overlap the previous tile's arithmetic and the span's copy can be
  // issued the moment a tile starts
  struct TileRef { const float* clip; int64_t len; int cls; };
  auto locate = [&](int2 rec) -> TileRef {
    TileRef r{nullptr, 0, kSilent};
    if (rec.x >= 0) {
      r.clip = clip_of(audio, offsets, rec.x, L, Lp, r.len);
      r.cls = classify_tile(r.clip, r.len, Lp, rec.y);
    }
    return r;
  };
  int2 it = list_at((int)blockIdx.x);
  TileRef ref = locate(it);
  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned bar_phase = 0;
  __syncthreads();                                       // tables and barrier visible
  for (int item = (int)blockIdx.x; item < n_live; item += stride) {
    const int64_t b = it.x;
    const int tt = it.y;
    const int64_t t0 = (int64_t)tt * kTileFrames;
    // ---- the tile's audio span.  No barrier is needed in front of it: what it overwrites (the
    // power rows) was last read before the barrier that follows the mel projection, and the
    // staged outputs the slower warps may still be writing out lie behind the span. ----
    {
      const float* src = ref.clip + ((int64_t)tt * (kTileFrames * kHop) - kNfft / 2);
      if (ref.cls == kFast16) {
        if (wid == 0) {
          // the span overwrites power rows written through the generic proxy: order the two proxies
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (lane == 0) mbar_expect_tx(&sm.bar, kTileSamples * 4);
          __syncwarp();
          constexpr int kPieces = (kTileSamples + 319) / 320;              // 17: sixteen of 320 samples + one of 240
          if (lane < kPieces)
            bulk_g2s(au + lane * (320 + kTileSkew), src + 320 * lane,
                     (unsigned)(4 * (lane < kPieces - 1 ? 320 : kTileSamples - 320 * (kPieces - 1))), &sm.bar);
          mbar_wait(&sm.bar, bar_phase);                 // one warp polls; the CTA barrier below releases the rest
        }
        bar_phase ^= 1u;
      } else if (ref.cls == kFast4) {
        for (int i = tid; i < kTileSamples; i += kThreads) cp_async4(&au[tile_pos(i)], src + i);
      } else {                                           // clip edges: reflection / zero padding, sample by sample
        const int64_t p0 = t0 * kHop;
        for (int i = tid; i < kTileSamples; i += kThreads) au[tile_pos(i)] = padded_sample(ref.clip, ref.len, Lp, p0 + i);
      }
      cp_async_commit();
      cp_async_wait<0>();
    }
    const int2 it_next = list_at(item + stride);         // next tile's lookups ride under this tile's work
    const TileRef ref_next = locate(it_next);
    __syncthreads();                                     // span visible to all
    float vmax = -INFINITY, vmin = INFINITY;
    {
      // ---- 16 complex FFT-400: columns, twiddle, rows ----
      {
        float2 x[20];
        stage1_load(g, j, au, sm.hann, x);
        __syncthreads();                                 // every thread holds its samples: Z may overwrite the span
        stage1_dft(g, j, x, sm.tw, sm.Z);
      }
      __syncthreads();
      // ---- rows; untangle the two real frames of each FFT straight from the registers (the
      // partner's half comes through shared memory); power rows then overwrite the Z storage ----
      float pa[kBinsPerThread], pb[kBinsPerThread];
      {
        float2 x[20];
        stage2_keep(g, j, sm.Z, x);
        __syncthreads();
        split_from_regs(g, j, sm.Z, x, pa, pb);
      }
      __syncthreads();
      split_store(g, j, P, pa, pb);
      __syncthreads();
      // ---- sparse mel projection + log10 ----
      if (mel_fast) {
        // thread = (filter m, run of nf consecutive frames) from the balanced table: the filter's
        // <= 16 quad-padded weights sit in registers, so a (filter, frame) pair costs 4 LDS + 4
        // FFMA per quad.  Results are staged in the (now free) audio buffer and written out row
        // by row so that the global stores are coalesced.
        const unsigned a = sm.assign[tid];
        const int nf = (int)(a >> 16);
        if (nf > 0) {
          const int m = (int)(a & 0xffu), f0 = (int)((a >> 8) & 0xffu);
          const int4 rc = sm.rec[m];                     // first bin, length, weight offset, quads
          const float4* wq = reinterpret_cast<const float4*>(sm.wts + rc.z);
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 w0 = rc.w > 0 ? wq[0] : z4, w1 = rc.w > 1 ? wq[1] : z4;
          const float4 w2 = rc.w > 2 ? wq[2] : z4, w3 = rc.w > 3 ? wq[3] : z4;
          const float* prow = P + prow_offset(f0) + rc.x;
          float* so = stage_out + m * kSoStride + f0;
          // one loop per support length (uniform within a warp: threads are sorted by quads),
          // two frames per trip so that the second frame's loads overlap the first frame's FMAs
          auto dot = [&](const float* pr, int quads) -> float {
            float a0 = 0.0f, a1 = 0.0f;
            a0 = fmaf(w0.x, pr[0], a0); a1 = fmaf(w0.y, pr[1], a1);
            a0 = fmaf(w0.z, pr[2], a0); a1 = fmaf(w0.w, pr[3], a1);
            if (quads > 1) {
              a0 = fmaf(w1.x, pr[4], a0); a1 = fmaf(w1.y, pr[5], a1);
              a0 = fmaf(w1.z, pr[6], a0); a1 = fmaf(w1.w, pr[7], a1);
            }
            if (quads > 2) {
              a0 = fmaf(w2.x, pr[8], a0); a1 = fmaf(w2.y, pr[9], a1);
              a0 = fmaf(w2.z, pr[10], a0); a1 = fmaf(w2.w, pr[11], a1);
            }
            if (quads > 3) {
              a0 = fmaf(w3.x, pr[12], a0); a1 = fmaf(w3.y, pr[13], a1);
              a0 = fmaf(w3.z, pr[14], a0); a1 = fmaf(w3.w, pr[15], a1);
            }
            return a0 + a1;
          };
          auto run = [&](auto quads_c) {
            constexpr int Q = decltype(quads_c)::value;
            int i = 0;
            for (; i + 2 <= nf; i += 2, prow += 2 * kPStride) {   // nf is even for every tier but the coarsest
              const float s0 = dot(prow, Q), s1 = dot(prow + kPStride, Q);
              so[i] = log10_floor(s0);
              so[i + 1] = log10_floor(s1);
            }
            if (i < nf) so[i] = log10_floor(dot(prow, Q));
          };
          switch (rc.w) {
            case 1:  run(std::integral_constant<int, 1>{}); break;
            case 2:  run(std::integral_constant<int, 2>{}); break;
            case 3:  run(std::integral_constant<int, 3>{}); break;
            default: run(std::integral_constant<int, 4>{}); break;
          }
        }
        __syncthreads();
        const int64_t t = t0 + lane;
        if (t < n_frames) {
          float* orow = out + (b * n_mels + wid) * n_frames + t;
          for (int m = wid; m < n_mels; m += kThreads / 32, orow += (int64_t)(kThreads / 32) * n_frames) {
            const float v = stage_out[m * kSoStride + lane];
            *orow = normalised(v);
            vmax = fmaxf(vmax, v);
            vmin = fminf(vmin, v);
          }
        }
      } else {
        // generic filterbanks: warp = filter, lane = frame
        const int64_t t = t0 + lane;
        const bool live = t < n_frames;
        const float* prow = P + prow_offset(lane);
        float* orow = out + (b * n_mels + wid) * n_frames + t;
        for (int m = wid; m < n_mels; m += kThreads / 32, orow += (int64_t)(kThreads / 32) * n_frames) {
          const int4 rc = sm.rec[m];
          const float v = packed ? mel_log10_quads(prow + rc.x, reinterpret_cast<const float4*>(sm.wts + rc.z), rc.w)
                                 : mel_log10(prow + rc.x, fb + (size_t)m * kBins + rc.x, rc.y);
          if (live) {
            *orow = normalised(v);
            vmax = fmaxf(vmax, v);
            vmin = fminf(vmin, v);
          }
        }
        __syncthreads();                                 // the power rows are read to the end of this path
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if (lane == 0) {
      atomicMax(clip_max + b, float_key(vmax));
      atomicMin(tile_min + b * tiles_per_clip + tt, float_key(vmin));
    }
    it = it_next;
    ref = ref_next;
  }
}

// The tile kernel stores y = (x + 4) / 4 right away; what is left is the clamp max(x, clip_max - 8),
// known only when the whole clip has been seen.  Rounding is monotonic, so
// (max(x, f) + 4) / 4 == max(y, (f + 4) / 4) bit for bit: the clamp is an elementwise maximum with a
// per-clip constant, and a tile whose smallest raw value is >= f (tile_min, kept by the tile kernel)
// needs no work at all -- it is neither read nor written again (a 64 x 30 s batch of noise: every
// tile).  Silent tiles (flagged by logmel_live_kernel) were never written: their raw value is the
// constant log10(1e-10), stored here without a read.
// One CTA per tile: a tile that needs nothing costs one flag load and an exit.
__device__ __forceinline__ float normalised(float x) { return __fmul_rn(__fadd_rn(x, 4.0f), 0.25f); }

__global__ void __launch_bounds__(128)
logmel_finalize_tiles_kernel(float* __restrict__ out, const int* __restrict__ clip_max,
                             const uint8_t* __restrict__ silent, const int* __restrict__ tile_min,
                             int64_t n_frames, int n_mels, int tiles_per_clip, int vec_ok) {
  const int64_t tile = blockIdx.x, b = tile / tiles_per_clip;
  const int tt = (int)(tile - b * tiles_per_clip);
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the tile kernel has completed (programmatic dependent launch)
  const float floor_v = __fsub_rn(key_float(clip_max[b]), 8.0f);
  const bool sil = silent[tile] != 0;
  if (!sil && !(key_float(tile_min[tile]) < floor_v)) return;      // the common case: nothing to clamp
  const float floor_y = normalised(floor_v);
  const float sil_y = normalised(fmaxf(log10_floor(0.0f), floor_v));
  const int64_t t0 = (int64_t)tt * kTileFrames;
  const int nt = (int)min((int64_t)kTileFrames, n_frames - t0);
  float* base = out + b * n_mels * n_frames + t0;
  if (vec_ok && nt == kTileFrames) {
    for (int idx = threadIdx.x; idx < n_mels * (kTileFrames / 4); idx += blockDim.x) {
      float4* p = reinterpret_cast<float4*>(base + (int64_t)(idx >> 3) * n_frames) + (idx & 7);
      if (sil) *p = make_float4(sil_y, sil_y, sil_y, sil_y);
      else {
        float4 v = *p;
        v.x = fmaxf(v.x, floor_y); v.y = fmaxf(v.y, floor_y); v.z = fmaxf(v.z, floor_y); v.w = fmaxf(v.w, floor_y);
        *p = v;
      }
    }
  } else {
    for (int idx = threadIdx.x; idx < n_mels * kTileFrames; idx += blockDim.x) {
      const int f = idx & (kTileFrames - 1);
      if (f >= nt) continue;
      float* p = base + (int64_t)(idx / kTileFrames) * n_frames + f;
      *p = sil ? sil_y : fmaxf(*p, floor_y);
    }
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
  if (L < 0 || padding < 0) return 0;
  const size_t tiles = (size_t)B * (size_t)(((L + padding) / lm::kHop + lm::kTileFrames - 1) / lm::kTileFrames);
  // clip maxima | MelPack | silent flags | live-tile count | live-tile list | per-tile minima
  return (((size_t)B * sizeof(int) + 15) & ~(size_t)15) + sizeof(lm::MelPack) + ((tiles + 15) & ~(size_t)15) + 16 +
         tiles * sizeof(int2) + tiles * sizeof(int) + 64;
}

extern "C" size_t avfe_logmel_pack_bytes(void) { return sizeof(lm::MelPack); }

extern "C" int avfe_logmel_prepare(const float* mel_filters, int n_mels, void* pack,
                                   avfe_stream_t stream) {
  if (n_mels <= 0) return AVFE_ERR_INVALID_ARG;
  if (n_mels > 128) return AVFE_ERR_UNSUPPORTED;
  if (!mel_filters || !pack || !aligned16(pack)) return AVFE_ERR_INVALID_ARG;
  lm::logmel_prep_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      mel_filters, n_mels, 0, nullptr, static_cast<lm::MelPack*>(pack));
  count_launch();
  return check_launch();
}

// shared tail of the two log-mel entry points: clip maxima <- -inf, tile kernel, finalize
static int logmel_run(const float* audio, const int64_t* offsets, int64_t B, int64_t L, int64_t padding, int n_mels,
                      const float* mel_filters, const lm::MelPack* pack, float* out,
                      int* clip_max, uint8_t* silent, cudaStream_t s) {
  const int64_t Lp = L + padding;
  const int64_t n_frames = Lp / lm::kHop;
  const int64_t tiles_per_clip = (n_frames + lm::kTileFrames - 1) / lm::kTileFrames;
  const int64_t n_tiles = B * tiles_per_clip;
  if (n_tiles > INT_MAX) return AVFE_ERR_UNSUPPORTED;
  // live-tile count and list follow the silent flags in the workspace
  int* live_count = reinterpret_cast<int*>(silent + (((size_t)n_tiles + 15) & ~(size_t)15));
  int2* live_list = reinterpret_cast<int2*>(live_count + 4);
  int* tile_min = reinterpret_cast<int*>(live_list + n_tiles);
  if (cudaMemsetAsync(live_count, 0, 16, s) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  lm::logmel_live_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(
      audio, offsets, B, L, Lp, (int)tiles_per_clip, clip_max, silent, live_count, live_list, tile_min);
  count_launch();
  // per-device attribute: set on every call (cheap) so multi-device processes stay correct
  if (cudaFuncSetAttribute(lm::logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(lm::Smem)) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  int64_t ctas = n_tiles < 3 * kNumSMs ? n_tiles : 3 * kNumSMs;   // 3 resident CTAs per SM
  {
    // programmatic stream serialization: the tile kernel's prologue (twiddles, window, filter tables into
    // shared memory) runs under logmel_live_kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(lm::kThreads);
    cfg.dynamicSmemBytes = sizeof(lm::Smem);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int64_t offsets_B = B;
    const int tpc = (int)tiles_per_clip;
    if (cudaLaunchKernelEx(&cfg, lm::logmel_tile_kernel, audio, offsets, offsets_B, L, Lp, n_frames, n_mels, mel_filters,
                           (const lm::MelPack*)pack, out, clip_max, silent, (const int*)live_count,
                           (const int2*)live_list, tile_min, tpc) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
  }
  count_launch();
  const int vec_ok = ((n_frames & 3) == 0 && aligned16(out)) ? 1 : 0;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_tiles);
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int tpc = (int)tiles_per_clip;
    if (cudaLaunchKernelEx(&cfg, lm::logmel_finalize_tiles_kernel, out, (const int*)clip_max, (const uint8_t*)silent,
                           (const int*)tile_min, n_frames, n_mels, tpc, vec_ok) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
  }
  count_launch();
  return check_launch();
}

static int logmel_check(const float* audio, int64_t B, int64_t L, int64_t padding, int n_mels,
                        const float* out, const void* workspace, size_t workspace_bytes) {
  if (B < 0 || L < 0 || padding < 0 || n_mels <= 0) return AVFE_ERR_INVALID_ARG;
  if (n_mels > 128) return AVFE_ERR_UNSUPPORTED;
  const int64_t Lp = L + padding;
  if (B == 0 || Lp / lm::kHop == 0) return 1;            // nothing to do
  if (Lp <= lm::kNfft / 2) return AVFE_ERR_INVALID_ARG;  // reflect pad needs pad < length
  if (!audio || !out) return AVFE_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < avfe_logmel_workspace_bytes(B, L, padding, n_mels) ||
      !aligned16(workspace))
    return AVFE_ERR_WORKSPACE;
  return AVFE_OK;
}

extern "C" int avfe_logmel_f32(const float* audio, int64_t B, int64_t L, int64_t padding,
                               int n_mels, const float* mel_filters, float* out, void* workspace,
                               size_t workspace_bytes, avfe_stream_t stream) {
  const int rc = logmel_check(audio, B, L, padding, n_mels, out, workspace, workspace_bytes);
  if (rc != AVFE_OK) return rc > 0 ? AVFE_OK : rc;
  if (!mel_filters) return AVFE_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* clip_max = static_cast<int*>(workspace);
  lm::MelPack* pack = reinterpret_cast<lm::MelPack*>(static_cast<char*>(workspace) +
                                                     (((size_t)B * sizeof(int) + 15) & ~(size_t)15));
  uint8_t* silent = reinterpret_cast<uint8_t*>(pack) + sizeof(lm::MelPack);
  lm::logmel_prep_kernel<<<1, 1024, 0, s>>>(mel_filters, n_mels, 0, nullptr, pack);
  count_launch();
  return logmel_run(audio, nullptr, B, L, padding, n_mels, mel_filters, pack, out, clip_max, silent, s);
}

extern "C" int avfe_logmel_prepared_f32(const float* audio, int64_t B, int64_t L, int64_t padding,
                                        int n_mels, const float* mel_filters, const void* pack,
                                        float* out, void* workspace, size_t workspace_bytes,
                                        avfe_stream_t stream) {
  const int rc = logmel_check(audio, B, L, padding, n_mels, out, workspace, workspace_bytes);
  if (rc != AVFE_OK) return rc > 0 ? AVFE_OK : rc;
  if (!mel_filters || !pack || !aligned16(pack)) return AVFE_ERR_INVALID_ARG;
  int* clip_max = static_cast<int*>(workspace);
  uint8_t* silent = reinterpret_cast<uint8_t*>(static_cast<char*>(workspace) +
                                               (((size_t)B * sizeof(int) + 15) & ~(size_t)15) + sizeof(lm::MelPack));
  return logmel_run(audio, nullptr, B, L, padding, n_mels, mel_filters, static_cast<const lm::MelPack*>(pack), out,
                    clip_max, silent, static_cast<cudaStream_t>(stream));
}

extern "C" int avfe_logmel_ragged_f32(const float* audio, const int64_t* offsets, int64_t B,
                                      int64_t length, int n_mels, const float* mel_filters,
                                      const void* pack, float* out, void* workspace,
                                      size_t workspace_bytes, avfe_stream_t stream) {
  const int rc = logmel_check(audio, B, length, 0, n_mels, out, workspace, workspace_bytes);
  if (rc != AVFE_OK) return rc > 0 ? AVFE_OK : rc;
  if (!offsets || !mel_filters) return AVFE_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* clip_max = static_cast<int*>(workspace);
  lm::MelPack* own = reinterpret_cast<lm::MelPack*>(static_cast<char*>(workspace) +
                                                    (((size_t)B * sizeof(int) + 15) & ~(size_t)15));
  uint8_t* silent = reinterpret_cast<uint8_t*>(own) + sizeof(lm::MelPack);
  const lm::MelPack* use = static_cast<const lm::MelPack*>(pack);
  if (use == nullptr) {                                   // no prepared pack: analyse the filters now
    lm::logmel_prep_kernel<<<1, 1024, 0, s>>>(mel_filters, n_mels, 0, nullptr, own);
    count_launch();
    use = own;
  } else if (!aligned16(pack)) {
    return AVFE_ERR_INVALID_ARG;
  }
  // L = 0: clips come from `offsets`; the padded length is `length`
  return logmel_run(audio, offsets, B, 0, length, n_mels, mel_filters, use, out, clip_max, silent, s);
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
