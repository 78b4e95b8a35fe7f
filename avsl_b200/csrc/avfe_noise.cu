// SNR noise mixing on the GPU, bit-exact with the reference's numpy code: add_noise(clean_wav,
// noise_wav, snr), preprocess/audio_process.py:110-150 (called by process_audio_for_av_hubert,
// :222-224, in front of the logfbank features).
//
//   clean_rms = sqrt(mean(clean^2));  noise'[i] = noise[i mod Ln], i < Lc;  noise_rms likewise
//   mixed = clean + noise' * ((clean_rms / 10^(snr/20)) / noise_rms)
//   if it leaves the int16 range: mixed *= 32767 / max  (or -32768 / min);   int16 by truncation
//
// Every step is float32 in the reference, so the only thing that decides the last bit of the gain
// is the ORDER of the two sums of squares.  numpy adds pairwise: the array is halved (at n/2
// rounded down to a multiple of 8) until a piece has at most 128 elements, a piece is summed in 8
// interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail.
// The kernels below rebuild exactly that tree:
//
//   noise_leaf_kernel     the tree is addressed like a heap (root 1, children 2k / 2k+1).  A piece
//                         has 64..128 elements, so the lane standing on the first multiple of
//                         64 inside it owns it: every lane of a warp walks down from the root
//                         for its own multiple of 64 (one walk serves the clean signal and the
//                         repeated noise: same length, same tree), then the warp's 8-lane groups
//                         go through the owners four at a time -- one numpy accumulator per
//                         lane, all loads ahead of the additions, three xor-shuffles (the same
//                         parenthesisation), the sequential tail handed over by shuffle.
//   noise_combine_kernel  one CTA per clip copies the clip's two heaps to shared memory, derives
//                         the node lengths top-down, adds the children level by level,
//                         bottom-up, then takes the two rms values and the gain and arms the
//                         clip's max / min.
//   noise_mix_kernel<0>   mixed = clean + noise' * gain with a rounded product and a rounded sum
//                         (no FMA) on groups of four samples, chunks walked in column order (the
//                         repeats of a short noise clip back to back), provisional int16 / float
//                         output, per-clip max / min through order-preserving integer atomics.
//   noise_mix_kernel<1>   the rescale pass: exits at once unless the clip left the int16 range;
//                         otherwise recomputes the mix, applies the reduction rate and rewrites
//                         the clip.
//
// Algorithmic bytes per sample: 4 (clean) + 4 (noise, at most Ln of them) + 2 (int16 out).  The four
// kernels above read both waveforms twice (sum of squares, then the mix); they are the fallback for
// clips whose tree is too deep for a shared-memory heap.  The default path is noise_cluster_kernel
// (further down): one launch per batch, a thread-block cluster per clip, partial sums exchanged
// through distributed shared memory, the mix re-reading its stretch from L2 -- the waveforms cross HBM
// once.
#include "avfe_common.cuh"
#include "avfe_noise_core.cuh"

namespace avfe {

constexpr int kLeafWarps = 4;        // warps per CTA of the leaf kernel (2048 samples each)
constexpr int kMixChunk = 2048;     // samples per CTA and step of the mix / rescale kernels
constexpr size_t kCombineSmemMax = 160 * 1024;   // both heaps and the node lengths of a clip up to ~57 s (depth 13)

struct NoiseClip {        // per-clip scalars in the workspace
  float gain;
  uint32_t max_key, min_key;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ int16_t to_i16(float v) {
  // C cast as x86 executes it (cvttss2si, low 16 bits); NaN / out of int32 range -> 0
  if (!(fabsf(v) < 2147483648.f)) return 0;
  return (int16_t)__float2int_rz(v);
}

struct NoiseArgs {
  const float* clean;
  const int64_t* clean_offsets;
  const float* noise;
  const int64_t* noise_offsets;
  const float* snr_ratio;
  float* heap;            // [B][2][heap_slots]
  NoiseClip* clips;       // [B]
  uint32_t heap_slots;    // 2^(depth+1)
  int depth;
};

// One piece (<= 128 elements from `src`, contiguous) summed the numpy way by 8 lanes: lane j owns
// accumulator j.  All loads are issued before the first addition (the chain of additions is short;
// loads waiting behind additions were the cost), the (< 8) leftover elements are fetched by lanes
// 0..6 and handed over by shuffle.  WRAP: the piece crosses the end of a repeated noise clip, every
// index goes through `mod period`.  Returns the piece's sum in every lane of the group.
template <bool WRAP>
__device__ __forceinline__ void piece_loads(const float* __restrict__ src, uint32_t start, uint32_t len, uint32_t period,
                                            int j, float (&v)[kLeafMax / 8], float& tail) {
  const uint32_t body = len & ~7u;
  const float* __restrict__ mine = src + start + j;        // one address per lane, constant offsets from it
  if (!WRAP && len >= 64u) {                                // every piece of a tree deeper than its root has >= 64 elements
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = mine[8 * t];
#pragma unroll
    for (int t = 8; t < kLeafMax / 8; ++t) {
      v[t] = 0.f;
      if ((uint32_t)(8 * t) < body) v[t] = mine[8 * t];
    }
  } else {
#pragma unroll
    for (int t = 0; t < kLeafMax / 8; ++t) {
      v[t] = 0.f;
      if ((uint32_t)(8 * t) < body) v[t] = WRAP ? src[(start + (uint32_t)(8 * t + j)) % period] : mine[8 * t];
    }
  }
  tail = 0.f;
  if (body + (uint32_t)j < len) tail = WRAP ? src[(start + body + (uint32_t)j) % period] : mine[body];
}

template <bool TAIL = true>
__device__ __forceinline__ float piece_sum(const float (&v)[kLeafMax / 8], float tail) {
  float r = __fmul_rn(v[0], v[0]);
#pragma unroll
  for (int t = 1; t < kLeafMax / 8; ++t) r = __fadd_rn(r, __fmul_rn(v[t], v[t]));   // r >= +0: an absent element adds +0, exact
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  if (TAIL) {                                                // leftover (< 8) elements, added one by one
    tail = __fmul_rn(tail, tail);
#pragma unroll
    for (int u = 0; u < 7; ++u) r = __fadd_rn(r, __shfl_sync(0xffffffffu, tail, u, 8));
  }
  return r;
}

// A warp covers 32 consecutive multiples of 64 samples of one clip.  Phase 1: every lane walks the
// tree down to the piece that holds its own multiple (the clean signal and the repeated noise have
// the same length, hence the same tree: one walk serves both).  Phase 2: eight rounds; in a round
// each 8-lane group takes over one lane's piece -- if that lane stands on the piece's FIRST
// multiple of 64 (a piece has 64..128 elements, so one or two multiples) -- and sums it for both
// signals.
__global__ void __launch_bounds__(kLeafWarps * 32)
noise_leaf_kernel(NoiseArgs a) {
  const int64_t b = blockIdx.y;
  const int lane = threadIdx.x & 31, j = lane & 7, grp = lane >> 3;
  const int64_t c0 = a.clean_offsets[b];
  const uint32_t n = (uint32_t)(a.clean_offsets[b + 1] - c0);
  const int64_t z0 = a.noise_offsets[b];
  const uint32_t period = (uint32_t)(a.noise_offsets[b + 1] - z0);
  const float* __restrict__ clean = a.clean + c0;
  const float* __restrict__ noise = a.noise + z0;
  float* heap = a.heap + b * 2 * (int64_t)a.heap_slots;
  // a warp per 2048-sample span of the clip (the loop runs once with the grid the library launches)
  for (uint64_t warp = blockIdx.x * (uint32_t)kLeafWarps + (threadIdx.x >> 5); warp * 2048u < n; warp += gridDim.x * (uint32_t)kLeafWarps) {
    const uint32_t pos = ((uint32_t)warp * 32u + (uint32_t)lane) * 64u;
    uint32_t k = 0, off = 0, len = 0, p0 = 0;
    if (pos < n) {
      k = locate_piece(pos, n, off, len);
      if (pos - off >= 64u) k = 0;                       // the piece belongs to the lane before
      if (period > 0) p0 = off % period;
    }
    // only about 17 of the 32 lanes own a piece (a piece holds one or two multiples of 64): the
    // groups walk the owners, four per round, instead of all 32 lanes
    const uint32_t owners = __ballot_sync(0xffffffffu, k != 0);
    const int n_own = __popc(owners);
#pragma unroll 1
    for (int r = 0; 4 * r < n_own; ++r) {
      const int nth = 4 * r + grp;
      const int cand = nth < n_own ? (int)__fns(owners, 0, nth + 1) : lane;     // an idle group reads its own lane, kk := 0
      const uint32_t k_any = __shfl_sync(0xffffffffu, k, cand);
      const uint32_t kk = nth < n_own ? k_any : 0u;
      const uint32_t o = __shfl_sync(0xffffffffu, off, cand);
      const uint32_t l_any = __shfl_sync(0xffffffffu, len, cand);
      const uint32_t l = kk ? l_any : 0u;                                    // l = 0: every load predicated off
      const uint32_t q = __shfl_sync(0xffffffffu, p0, cand);
      float vc[kLeafMax / 8], vz[kLeafMax / 8], tc, tz;
      piece_loads<false>(clean, o, l, 0u, j, vc, tc);
      const uint32_t lz = period > 0 ? l : 0u;
      if (q + lz <= period) {
        piece_loads<false>(noise, q, lz, period, j, vz, tz);
      } else {
        piece_loads<true>(noise, q, lz, period, j, vz, tz);
      }
      const float sc = piece_sum(vc, tc);
      const float sz = piece_sum(vz, tz);
      if (kk != 0 && kk < a.heap_slots && j < 2) heap[(int64_t)j * a.heap_slots + kk] = j ? sz : sc;   // kk >= slots: a clip longer than the max_len the caller promised
    }
  }
}

// One CTA per clip.  Shared memory: the lengths of the tree's nodes (top-down, so that no node has
// to be located by a walk) and both heaps; the children are added level by level, bottom-up.
template <bool SMEM>
__global__ void __launch_bounds__(1024)
noise_combine_kernel(NoiseArgs a) {
  extern __shared__ float heap_s[];
  const int64_t b = blockIdx.x;
  const uint32_t n = (uint32_t)(a.clean_offsets[b + 1] - a.clean_offsets[b]);
  const uint32_t period = (uint32_t)(a.noise_offsets[b + 1] - a.noise_offsets[b]);
  float* heap_g = a.heap + b * 2 * (int64_t)a.heap_slots;
  float* heap = heap_g;
  if (n > (uint32_t)kLeafMax) {
    if (SMEM) {
      heap = heap_s;
      uint32_t* len_s = reinterpret_cast<uint32_t*>(heap_s + 2 * a.heap_slots);   // nodes above the last level
      const float4* g4 = reinterpret_cast<const float4*>(heap_g);      // 256-byte aligned, slots % 4 == 0
      float4* s4 = reinterpret_cast<float4*>(heap_s);
      const uint32_t quads = a.heap_slots / 2;
      for (uint32_t t = threadIdx.x; t < quads; t += 8 * blockDim.x) {
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t + u * blockDim.x < quads) q[u] = g4[t + u * blockDim.x];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t + u * blockDim.x < quads) s4[t + u * blockDim.x] = q[u];
      }
      if (threadIdx.x == 0) len_s[1] = n;
      __syncthreads();
      for (int d = 0; d + 1 < a.depth; ++d) {             // lengths of the levels 1 .. depth-1
        for (uint32_t k = (1u << d) + threadIdx.x; k < (2u << d); k += blockDim.x) {
          const uint32_t len = len_s[k];
          const uint32_t half = len > (uint32_t)kLeafMax ? (len >> 1) & ~7u : 0u;
          len_s[2 * k] = half;
          len_s[2 * k + 1] = len > (uint32_t)kLeafMax ? len - half : 0u;
        }
        __syncthreads();
      }
      for (int d = a.depth - 1; d >= 0; --d) {
        for (uint32_t k = (1u << d) + threadIdx.x; k < (2u << d); k += blockDim.x)
          if (len_s[k] > (uint32_t)kLeafMax) {
            heap_s[k] = __fadd_rn(heap_s[2 * k], heap_s[2 * k + 1]);
            heap_s[a.heap_slots + k] = __fadd_rn(heap_s[a.heap_slots + 2 * k], heap_s[a.heap_slots + 2 * k + 1]);
          }
        __syncthreads();
      }
    } else {
      for (int d = a.depth - 1; d >= 0; --d) {
        const uint32_t first = 1u << d;
        for (uint32_t t = threadIdx.x; t < 2 * first; t += blockDim.x) {
          const uint32_t k = first + (t >> 1);
          float* h = heap_g + (t & 1u) * (int64_t)a.heap_slots;
          uint32_t off, len;
          if (locate_node(k, n, off, len) == 2) h[k] = __fadd_rn(h[2 * k], h[2 * k + 1]);
        }
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) {
    NoiseClip c;
    c.gain = 0.f;
    c.max_key = 0u;
    c.min_key = 0xffffffffu;
    c.pad = 0u;
    if (n > 0 && period > 0) {
      const float clean_ms = __double2float_rn(__ddiv_rn((double)heap[1], (double)n));
      const float noise_ms = __double2float_rn(__ddiv_rn((double)heap[a.heap_slots + 1], (double)n));
      const float clean_rms = __fsqrt_rn(clean_ms), noise_rms = __fsqrt_rn(noise_ms);
      c.gain = __fdiv_rn(__fdiv_rn(clean_rms, a.snr_ratio[b]), noise_rms);
    }
    a.clips[b] = c;
  }
}

struct MixArgs {
  NoiseArgs n;
  int16_t* out_i16;
  float* out_f32;
};

// Samples are handled in groups of four on ABSOLUTE 16-byte boundaries of the packed buffers (a
// clip may start anywhere).  A CTA strides over the clip's chunks of kMixChunk samples with the
// next chunk's loads in flight while the current one is mixed.  A chunk that lies wholly inside
// the clip (CTA-uniform test) takes the lean path: one float4 of clean signal and one 8-byte int16
// (16-byte float) store per group, the noise position carried along incrementally instead of a
// `mod period` per group.  The chunks at the two ends of the clip go sample by sample.
template <bool RESCALE>
__global__ void __launch_bounds__(256)
noise_mix_kernel(MixArgs m) {
  const NoiseArgs& a = m.n;
  const int64_t b = blockIdx.y;
  // behind noise_cluster_kernel the rescale pass is launched programmatically (its grid is set up while
  // the clusters run); a no-op under a plain launch
  if (RESCALE) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t c0 = a.clean_offsets[b];
  const int64_t c1 = a.clean_offsets[b + 1];
  const int64_t g0 = c0 & ~(int64_t)3;                    // first group of the clip (may start before it)
  const int64_t total = (c1 - g0 + kMixChunk - 1) / kMixChunk;      // chunks of this clip
  const int64_t z0 = a.noise_offsets[b];
  const uint32_t period = (uint32_t)(a.noise_offsets[b + 1] - z0);
  // Chunk order.  A repeated noise clip is read once from HBM if its repeats are mixed back to back:
  // the clip's chunks form `cols` columns (chunk c belongs to column c mod cols, cols = the noise
  // period in chunks), a CTA takes the columns blockIdx.x, + gridDim.x, ... and walks each column
  // down (c, c + cols, c + 2 cols, ...: the same stretch of noise, give or take a chunk's rounding).
  // Without repetition (or with a period so short that L2 keeps it anyway) every column is one chunk.
  int64_t cols = ((int64_t)period + kMixChunk / 2) / kMixChunk;
  if (period == 0 || cols < (int64_t)gridDim.x || cols >= total) cols = total;
  if ((int64_t)blockIdx.x >= cols) return;
  const NoiseClip clip = a.clips[b];
  float rate = 1.f;
  if (RESCALE) {
    const float hi = key_float(clip.max_key), lo = key_float(clip.min_key);
    if (!(hi > 32767.f || lo < -32768.f)) return;
    rate = (hi >= fabsf(lo)) ? __fdiv_rn(32767.f, hi) : __fdiv_rn(-32768.f, lo);
  }
  const float* __restrict__ noise = a.noise + z0;
  const bool noise_vec = ((reinterpret_cast<uintptr_t>(noise) & 15u) == 0);
  float vmax = -INFINITY, vmin = INFINITY;
  constexpr int kRounds = kMixChunk / 1024;               // groups per thread and chunk
  // noise position of this thread's first group of a chunk, kept in [0, period): it advances by
  // 1024 per round, by `cols` chunks down a column and by gridDim.x chunks to the next column, each
  // step reduced mod period once
  const uint32_t per = period ? period : 1u;
  const uint32_t step_round = 1024u % per;
  const uint32_t step_down = (uint32_t)((cols * kMixChunk) % per);
  const uint32_t step_col = (uint32_t)(((int64_t)gridDim.x * kMixChunk) % per);
  struct Walk {
    int64_t col, c;          // column, chunk
    uint32_t pos_col, pos;   // noise position at the head of the column / at chunk c
  };
  auto advance = [&](Walk& w) {
    w.c += cols;
    w.pos += step_down;
    if (w.pos >= per) w.pos -= per;
    if (w.c >= total) {
      w.col += gridDim.x;
      w.c = w.col;
      w.pos_col += step_col;
      if (w.pos_col >= per) w.pos_col -= per;
      w.pos = w.pos_col;
    }
  };
  Walk cur;
  cur.col = cur.c = blockIdx.x;
  {
    const int64_t i0 = g0 + (int64_t)blockIdx.x * kMixChunk + 4 * (int)threadIdx.x - c0;      // >= -3
    cur.pos_col = cur.pos = (uint32_t)(((i0 % (int64_t)per) + per) % per);
  }

  auto fetch = [&](const Walk& w, float4 (&x)[kRounds], float4 (&z)[kRounds]) {
    const int64_t base = g0 + w.c * kMixChunk;
    const bool inner = base >= c0 && base + kMixChunk <= c1;
    uint32_t p = w.pos;
    if (inner) {
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int64_t g = base + r * 1024 + 4 * (int)threadIdx.x;
        x[r] = *reinterpret_cast<const float4*>(a.clean + g);
        z[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (period > 0) {
          if (noise_vec && (p & 3u) == 0 && p + 4 <= period) {
            z[r] = *reinterpret_cast<const float4*>(noise + p);
          } else if (p + 4 <= period) {
            z[r] = make_float4(noise[p], noise[p + 1], noise[p + 2], noise[p + 3]);
          } else {
            z[r] = make_float4(noise[p], noise[(p + 1) % period], noise[(p + 2) % period], noise[(p + 3) % period]);
          }
        }
        p += step_round;
        if (p >= per) p -= per;
      }
    } else {
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int64_t g = base + r * 1024 + 4 * (int)threadIdx.x;
        x[r] = z[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        float* xs = reinterpret_cast<float*>(&x[r]);
        float* zs = reinterpret_cast<float*>(&z[r]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (g + q >= c0 && g + q < c1) {
            xs[q] = a.clean[g + q];
            if (period > 0) zs[q] = noise[(uint32_t)(g + q - c0) % period];
          }
      }
    }
  };

  auto mix4 = [&](const float4& x, const float4& z, int16_t (&q16)[4], float (&v)[4]) {
    const float xs[4] = {x.x, x.y, x.z, x.w};
    const float zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = xs[q];
      if (period > 0) v[q] = __fadd_rn(v[q], __fmul_rn(zs[q], clip.gain));
      if (RESCALE) v[q] = __fmul_rn(v[q], rate);
      q16[q] = to_i16(v[q]);
    }
  };

  auto process = [&](const Walk& w, const float4 (&x)[kRounds], const float4 (&z)[kRounds]) {
    const int64_t base = g0 + w.c * kMixChunk;
    const bool inner = base >= c0 && base + kMixChunk <= c1;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int64_t g = base + r * 1024 + 4 * (int)threadIdx.x;
      int16_t q16[4];
      float v[4];
      mix4(x[r], z[r], q16, v);
      if (inner) {
        if (!RESCALE) {
          vmax = fmaxf(fmaxf(vmax, fmaxf(v[0], v[1])), fmaxf(v[2], v[3]));
          vmin = fminf(fminf(vmin, fminf(v[0], v[1])), fminf(v[2], v[3]));
        }
        if (m.out_i16) {
          const uint32_t lo = (uint16_t)q16[0] | ((uint32_t)(uint16_t)q16[1] << 16);
          const uint32_t hi = (uint16_t)q16[2] | ((uint32_t)(uint16_t)q16[3] << 16);
          *reinterpret_cast<uint2*>(m.out_i16 + g) = make_uint2(lo, hi);
        }
        if (m.out_f32)
          *reinterpret_cast<float4*>(m.out_f32 + g) = make_float4((float)q16[0], (float)q16[1], (float)q16[2], (float)q16[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (g + q >= c0 && g + q < c1) {
            if (!RESCALE) {
              vmax = fmaxf(vmax, v[q]);
              vmin = fminf(vmin, v[q]);
            }
            if (m.out_i16) m.out_i16[g + q] = q16[q];
            if (m.out_f32) m.out_f32[g + q] = (float)q16[q];
          }
      }
    }
  };

  float4 xa[kRounds], za[kRounds], xb[kRounds], zb[kRounds];
  fetch(cur, xa, za);
  for (;;) {
    Walk nxt = cur;
    advance(nxt);
    bool more = nxt.col < cols;
    if (more) fetch(nxt, xb, zb);
    process(cur, xa, za);
    if (!more) break;
    cur = nxt;
    advance(nxt);
    more = nxt.col < cols;
    if (more) fetch(nxt, xa, za);
    process(cur, xb, zb);
    if (!more) break;
    cur = nxt;
  }
  if (!RESCALE) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
    }
    if ((threadIdx.x & 31) == 0 && vmax >= vmin) {
      atomicMax(&a.clips[b].max_key, float_key(vmax));
      atomicMin(&a.clips[b].min_key, float_key(vmin));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same pipeline in ONE launch per batch: a thread-block cluster owns a clip, so that the clip's
// two waveforms cross HBM once (the mix re-reads them from L2 a few microseconds after the sums).
//
// CTA `rank` of a CS-CTA cluster owns the subtree under heap node CS + rank of numpy's pairwise
// tree (depth log2 CS), i.e. one contiguous CS-th of the clip:
//   1. leaf sums of its pieces straight into a LOCAL heap in shared memory (same walk as
//      noise_leaf_kernel, started at the subtree's root),
//   2. the subtree is added up in shared memory (node lengths top-down, children bottom-up),
//   3. the CS pairs of partial sums are exchanged through distributed shared memory (every CTA stores
//      its pair into every CTA's table), one cluster barrier, and every CTA adds the top log2 CS
//      levels in numpy's order and derives the gain -- redundantly, with the same roundings,
//   4. the CTA mixes its own stretch (still in L2), provisional int16 / float output, per-clip max / min
//      through the same order-preserving atomics; the rescale pass stays noise_mix_kernel<true>.
// A clip too short to have CS * 256 samples is handled by CTA 0 alone from the root (its tree may
// end above depth log2 CS).  The exchange table is double-buffered: a CTA that is two clips ahead
// has passed a barrier every other CTA has arrived at, hence nobody still reads the older entry.
constexpr int kClusterThreads = 512;                       // two CTAs per SM (of different clips: their phases overlap)
constexpr int kClusterMaxLocalDepth = 12;                    // 8192 slots per signal: 80 KB of shared memory

struct ClusterSmem {
  float2 xch[2][8];                                          // partial sums (clean, noise) of every rank
  float gain;
  uint32_t pad[3];
  float heap[1];                                             // [2][slots] sums, then [slots / 2] node lengths
};

// the leaf piece under (off, len) that holds element pos: local heap index (root = 1), offset, length
__device__ __forceinline__ uint32_t locate_piece_from(uint32_t pos, uint32_t& off, uint32_t& len) {
  uint32_t k = 1;
  while (len > (uint32_t)kLeafMax) {
    const uint32_t half = (len >> 1) & ~7u;
    if (pos < off + half) {
      len = half;
      k = 2 * k;
    } else {
      off += half;
      len -= half;
      k = 2 * k + 1;
    }
  }
  return k;
}

template <int CS>
__global__ void __launch_bounds__(kClusterThreads, 2)
noise_cluster_kernel(MixArgs m, int local_depth) {
  extern __shared__ __align__(16) unsigned char cl_raw[];
  ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>(cl_raw);
  const NoiseArgs& a = m.n;
  const uint32_t slots = 2u << local_depth;
  float* hc = sm.heap;
  float* hz = sm.heap + slots;
  uint32_t* len_s = reinterpret_cast<uint32_t*>(sm.heap + 2 * slots);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the rescale grid may be set up; it waits for this one
  unsigned rank = 0;
  if (CS > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // "I am running": waited for just before the first remote store (a CTA's shared memory must not be
    // written before the CTA has started)
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  }
  const int64_t b = blockIdx.x / CS;
  const int64_t c0 = a.clean_offsets[b];
  const uint32_t n = (uint32_t)(a.clean_offsets[b + 1] - c0);
  const int64_t z0 = a.noise_offsets[b];
  const uint32_t period = (uint32_t)(a.noise_offsets[b + 1] - z0);
  const float* __restrict__ clean = a.clean + c0;
  const float* __restrict__ noise = a.noise + z0;

  // this CTA's subtree: node CS + rank when the tree is full down to that depth, else the root for rank 0
  const bool split = CS > 1 && n >= (uint32_t)(256 * CS);
  uint32_t off_r = 0, len_r = (split || rank == 0) ? n : 0u;
  if (split) {
#pragma unroll
    for (int d = 0; (1 << d) < CS; ++d) {                    // walk the bits of `rank` from the top
      const uint32_t half = (len_r >> 1) & ~7u;
      constexpr int kLog = (CS == 8) ? 3 : (CS == 4) ? 2 : 1;
      if ((rank >> (kLog - 1 - d)) & 1u) { off_r += half; len_r -= half; } else { len_r = half; }
    }
  }
  if (rank == 0 && tid == 0) {                               // armed before the barrier: every CTA's atomics come after it
    a.clips[b].max_key = 0u;
    a.clips[b].min_key = 0xffffffffu;
  }

  // ---- 1. leaf sums of the pieces of [off_r, off_r + len_r) ----
  if (len_r > 0) {
    const int j = lane & 7, grp = lane >> 3;
    const uint32_t first = (off_r + 63u) >> 6, end = off_r + len_r;     // multiples of 64 inside the stretch
    const uint32_t count = ((end - 1u) >> 6) >= first ? ((end - 1u) >> 6) - first + 1u : 0u;
    for (uint32_t blk = (uint32_t)wid; blk * 32u < count; blk += kClusterThreads / 32) {
      const uint32_t idx = blk * 32u + (uint32_t)lane;
      const uint32_t pos = (first + idx) << 6;
      uint32_t k = 0, off = off_r, len = len_r, p0 = 0;
      if (idx < count) {
        k = locate_piece_from(pos, off, len);
        if (pos - off >= 64u) k = 0;                         // the piece belongs to the lane before
        if (period > 0) p0 = off % period;
      }
      uint32_t owners = __ballot_sync(0xffffffffu, k != 0);
      // pieces whose length is not a multiple of 8 exist only on the right edge of the tree: the
      // seven-shuffle tail is skipped for the rounds (almost all) that have none
      const uint32_t tails = __ballot_sync(0xffffffffu, k != 0 && (len & 7u) != 0);
#pragma unroll 1
      while (owners != 0u) {
        // the four lowest owners, one per 8-lane group
        uint32_t mleft = owners;
        int cand = lane;
        bool has = false;
#pragma unroll
        for (int gsel = 0; gsel < 4; ++gsel) {
          const int pos = __ffs((int)mleft) - 1;             // -1 when the set is exhausted
          if (gsel == grp) { has = pos >= 0; cand = has ? pos : lane; }
          mleft &= mleft - 1u;
        }
        const uint32_t round_set = owners & ~mleft;
        owners = mleft;
        const uint32_t k_any = __shfl_sync(0xffffffffu, k, cand);
        const uint32_t kk = has ? k_any : 0u;
        const uint32_t o = __shfl_sync(0xffffffffu, off, cand);
        const uint32_t l_any = __shfl_sync(0xffffffffu, len, cand);
        const uint32_t l = kk ? l_any : 0u;
        const uint32_t q = __shfl_sync(0xffffffffu, p0, cand);
        float vc[kLeafMax / 8], vz[kLeafMax / 8], tc, tz;
        piece_loads<false>(clean, o, l, 0u, j, vc, tc);
        const uint32_t lz = period > 0 ? l : 0u;
        if (q + lz <= period) {
          piece_loads<false>(noise, q, lz, period, j, vz, tz);
        } else {
          piece_loads<true>(noise, q, lz, period, j, vz, tz);
        }
        float sc, sz;
        if (round_set & tails) {                             // warp-uniform
          sc = piece_sum<true>(vc, tc);
          sz = piece_sum<true>(vz, tz);
        } else {
          sc = piece_sum<false>(vc, tc);
          sz = piece_sum<false>(vz, tz);
        }
        if (kk != 0 && kk < slots && j < 2) (j ? hz : hc)[kk] = j ? sz : sc;
      }
    }
  }
  __syncthreads();

  // ---- 2. the subtree, in shared memory ----
  if (len_r > (uint32_t)kLeafMax) {
    if (tid == 0) len_s[1] = len_r;
    __syncthreads();
    for (int d = 0; d + 1 < local_depth; ++d) {
      for (uint32_t k = (1u << d) + tid; k < (2u << d); k += kClusterThreads) {
        const uint32_t len = len_s[k];
        const uint32_t half = len > (uint32_t)kLeafMax ? (len >> 1) & ~7u : 0u;
        len_s[2 * k] = half;
        len_s[2 * k + 1] = len > (uint32_t)kLeafMax ? len - half : 0u;
      }
      __syncthreads();
    }
    for (int d = local_depth - 1; d >= 0; --d) {
      for (uint32_t k = (1u << d) + tid; k < (2u << d); k += kClusterThreads)
        if (len_s[k] > (uint32_t)kLeafMax) {
          hc[k] = __fadd_rn(hc[2 * k], hc[2 * k + 1]);
          hz[k] = __fadd_rn(hz[2 * k], hz[2 * k + 1]);
        }
      __syncthreads();
    }
  }

  // ---- 3. exchange, top levels, gain ----
  const int par = (int)(b & 1);       // one clip per cluster and launch: the table is written once; parity kept for striding callers
  if (CS > 1) {
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");      // every CTA of the cluster has started
    if (tid < CS) {
      const float2 mine = make_float2(len_r > 0 ? hc[1] : 0.f, len_r > 0 ? hz[1] : 0.f);
      const unsigned local = (unsigned)__cvta_generic_to_shared(&sm.xch[par][rank]);
      unsigned remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((unsigned)tid));
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(mine.x), "f"(mine.y) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    if (tid == 0) sm.xch[par][0] = make_float2(hc[1], hz[1]);
    __syncthreads();
  }
  if (tid == 0) {
    float sc, sz;
    if (split) {
      float pc[8], pz[8];
#pragma unroll
      for (int r = 0; r < CS; ++r) { pc[r] = sm.xch[par][r].x; pz[r] = sm.xch[par][r].y; }
#pragma unroll
      for (int w = CS; w > 1; w >>= 1)
#pragma unroll
        for (int r = 0; r < w / 2; ++r) {
          pc[r] = __fadd_rn(pc[2 * r], pc[2 * r + 1]);
          pz[r] = __fadd_rn(pz[2 * r], pz[2 * r + 1]);
        }
      sc = pc[0];
      sz = pz[0];
    } else {
      sc = sm.xch[par][0].x;
      sz = sm.xch[par][0].y;
    }
    float gain = 0.f;
    if (n > 0 && period > 0) {
      const float clean_ms = __double2float_rn(__ddiv_rn((double)sc, (double)n));
      const float noise_ms = __double2float_rn(__ddiv_rn((double)sz, (double)n));
      const float clean_rms = __fsqrt_rn(clean_ms), noise_rms = __fsqrt_rn(noise_ms);
      gain = __fdiv_rn(__fdiv_rn(clean_rms, a.snr_ratio[b]), noise_rms);
    }
    sm.gain = gain;
    if (rank == 0) a.clips[b].gain = gain;
  }
  __syncthreads();
  if (len_r == 0) return;
  const float gain = sm.gain;

  // ---- 4. the mix of [off_r, off_r + len_r): full groups of four samples on absolute 16-byte boundaries
  // (32-bit indices from the first group, nothing to test per sample), the <= 3 samples in front of
  // and behind them one thread each ----
  const int64_t lo = c0 + off_r, hi = lo + len_r;            // absolute sample range in the packed buffer
  const int64_t a_lo = min((lo + 3) & ~(int64_t)3, hi), a_hi = max(hi & ~(int64_t)3, a_lo);
  const uint32_t nb = (uint32_t)((a_hi - a_lo) >> 2);         // full groups
  const bool noise_vec = ((reinterpret_cast<uintptr_t>(noise) & 15u) == 0);
  const uint32_t per = period ? period : 1u;
  float vmax = -INFINITY, vmin = INFINITY;
  auto mix1 = [&](float x, float z) -> float { return period > 0 ? __fadd_rn(x, __fmul_rn(z, gain)) : x; };
  {                                                           // edges: samples [lo, a_lo) and [a_hi, hi)
    const int64_t s = tid < 3 ? lo + tid : a_hi + (tid - 32);
    const bool mine = tid < 3 ? s < a_lo : (tid >= 32 && tid < 35 && s < hi);
    if (mine) {
      const float v = mix1(a.clean[s], period > 0 ? noise[(uint32_t)(s - c0) % per] : 0.f);
      vmax = v;
      vmin = v;
      const int16_t q = to_i16(v);
      if (m.out_i16) m.out_i16[s] = q;
      if (m.out_f32) m.out_f32[s] = (float)q;
    }
  }
  const float4* __restrict__ xb = reinterpret_cast<const float4*>(a.clean + a_lo);
  const uint32_t step = (uint32_t)((4u * kClusterThreads) % per);
  uint32_t p = (uint32_t)((uint64_t)(a_lo - c0 + 4 * (int64_t)tid) % per);   // noise position of this thread's group
  constexpr int kU = 4;                                       // groups in flight per thread
  for (uint32_t g0 = (uint32_t)tid; g0 < nb; g0 += kU * kClusterThreads) {
    float4 x[kU], z[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t g = g0 + u * kClusterThreads;
      x[u] = z[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < nb) {
        x[u] = xb[g];
        if (period > 0) {
          if (noise_vec && (p & 3u) == 0 && p + 4 <= period) {
            z[u] = *reinterpret_cast<const float4*>(noise + p);
          } else if (p + 4 <= period) {
            z[u] = make_float4(noise[p], noise[p + 1], noise[p + 2], noise[p + 3]);
          } else {
            z[u] = make_float4(noise[p], noise[(p + 1) % period], noise[(p + 2) % period], noise[(p + 3) % period]);
          }
        }
      }
      p += step;
      if (p >= per) p -= per;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t g = g0 + u * kClusterThreads;
      if (g >= nb) break;
      const float v0 = mix1(x[u].x, z[u].x), v1 = mix1(x[u].y, z[u].y), v2 = mix1(x[u].z, z[u].z), v3 = mix1(x[u].w, z[u].w);
      vmax = fmaxf(fmaxf(vmax, fmaxf(v0, v1)), fmaxf(v2, v3));
      vmin = fminf(fminf(vmin, fminf(v0, v1)), fminf(v2, v3));
      const int16_t q0 = to_i16(v0), q1 = to_i16(v1), q2 = to_i16(v2), q3 = to_i16(v3);
      const int64_t s = a_lo + 4 * (int64_t)g;
      if (m.out_i16) {
        const uint32_t w0 = (uint16_t)q0 | ((uint32_t)(uint16_t)q1 << 16);
        const uint32_t w1 = (uint16_t)q2 | ((uint32_t)(uint16_t)q3 << 16);
        *reinterpret_cast<uint2*>(m.out_i16 + s) = make_uint2(w0, w1);
      }
      if (m.out_f32) *reinterpret_cast<float4*>(m.out_f32 + s) = make_float4((float)q0, (float)q1, (float)q2, (float)q3);
    }
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, sft));
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, sft));
  }
  if (lane == 0 && vmax >= vmin) {
    atomicMax(&a.clips[b].max_key, float_key(vmax));
    atomicMin(&a.clips[b].min_key, float_key(vmin));
  }
}

template <int CS>
static int launch_noise_cluster(const MixArgs& m, int64_t B, int local_depth, cudaStream_t s) {
  const size_t slots = (size_t)2 << local_depth;
  const size_t smem = offsetof(ClusterSmem, heap) + (2 * slots + slots / 2) * sizeof(float);
  if (cudaFuncSetAttribute(noise_cluster_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(kClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, noise_cluster_kernel<CS>, m, local_depth) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  return AVFE_OK;
}

inline size_t noise_heap_bytes(int64_t B, int depth) {
  return (size_t)B * 2 * ((size_t)2 << depth) * sizeof(float);
}

}  // namespace avfe

extern "C" size_t avfe_add_noise_workspace_bytes(int64_t B, int64_t max_len) {
  using namespace avfe;
  if (B <= 0 || max_len <= 0) return 256;
  const int depth = tree_depth(max_len);
  if (depth > kMaxDepth) return 0;
  return 256 + noise_heap_bytes(B, depth) + (size_t)B * sizeof(NoiseClip);
}

extern "C" int avfe_add_noise(const float* clean, const int64_t* clean_offsets, const float* noise,
                              const int64_t* noise_offsets, const float* snr_ratio, int64_t B,
                              int64_t max_len, int16_t* out_i16, float* out_f32, void* workspace,
                              size_t workspace_bytes, avfe_stream_t stream) {
  using namespace avfe;
  if (B < 0 || max_len < 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || max_len == 0) return AVFE_OK;
  if (!clean || !clean_offsets || !noise || !noise_offsets || !snr_ratio || !workspace) return AVFE_ERR_INVALID_ARG;
  if (!out_i16 && !out_f32) return AVFE_ERR_INVALID_ARG;
  if (B > 65535 || max_len >= ((int64_t)1 << 31)) return AVFE_ERR_UNSUPPORTED;
  if (!aligned16(clean) || (out_i16 && (reinterpret_cast<uintptr_t>(out_i16) & 7u)) || (out_f32 && !aligned16(out_f32)))
    return AVFE_ERR_INVALID_ARG;                       // packed buffers: 16-byte aligned (int16 output: 8)
  const int depth = tree_depth(max_len);
  if (depth > kMaxDepth) return AVFE_ERR_UNSUPPORTED;
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  const size_t need = (p - reinterpret_cast<uintptr_t>(workspace)) + noise_heap_bytes(B, depth) + (size_t)B * sizeof(NoiseClip);
  if (workspace_bytes < need) return AVFE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MixArgs m;
  m.n.clean = clean;
  m.n.clean_offsets = clean_offsets;
  m.n.noise = noise;
  m.n.noise_offsets = noise_offsets;
  m.n.snr_ratio = snr_ratio;
  m.n.heap = reinterpret_cast<float*>(p);
  m.n.clips = reinterpret_cast<NoiseClip*>(p + noise_heap_bytes(B, depth));
  m.n.heap_slots = 2u << depth;
  m.n.depth = depth;
  m.out_i16 = out_i16;
  m.out_f32 = out_f32;
  const unsigned chunks_rs = (unsigned)((max_len + 3 + kMixChunk - 1) / kMixChunk);
  // one launch per batch, a cluster per clip (the waveforms cross HBM once); cluster width by clip length
  {
    const int cs = max_len >= 131072 ? 8 : max_len >= 32768 ? 4 : max_len >= 8192 ? 2 : 1;
    const int log_cs = cs == 8 ? 3 : cs == 4 ? 2 : cs == 2 ? 1 : 0;
    // local heap depth: the subtree of a split clip, or a whole clip shorter than 256 * cs samples
    int local_depth = depth - log_cs;
    const int short_depth = tree_depth((int64_t)256 * cs);
    if (local_depth < short_depth) local_depth = short_depth;
    if (local_depth < 1) local_depth = 1;
    if (local_depth <= kClusterMaxLocalDepth && B * cs <= 0x7fffffffLL) {
      int rc = cs == 8 ? launch_noise_cluster<8>(m, B, local_depth, s)
             : cs == 4 ? launch_noise_cluster<4>(m, B, local_depth, s)
             : cs == 2 ? launch_noise_cluster<2>(m, B, local_depth, s)
                       : launch_noise_cluster<1>(m, B, local_depth, s);
      if (rc != AVFE_OK) return rc;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(chunks_rs < 16u ? chunks_rs : 16u, (unsigned)B);
      cfg.blockDim = dim3(256);
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, noise_mix_kernel<true>, m) != cudaSuccess) {
        cudaGetLastError();
        return AVFE_ERR_CUDA;
      }
      count_launch(2);
      return check_launch();
    }
  }
  // clips too long for the cluster kernel's shared-memory heap: leaf sums, combine, mix as separate launches
  // small CTAs (4 warps = 8192 samples) scheduled by the hardware: 39 us on 64 x 30 s, against 43.5 us
  // with 8-warp CTAs (3.2 waves) and 49-61 us with one resident wave of striding CTAs (the warps' rounds
  // are serial, so static striding leaves the tail uneven)
  const unsigned leaf_ctas = (unsigned)((max_len + kLeafWarps * 2048 - 1) / (kLeafWarps * 2048));
  noise_leaf_kernel<<<dim3(leaf_ctas, (unsigned)B), kLeafWarps * 32, 0, s>>>(m.n);
  const size_t heap_smem = (2 * (size_t)m.n.heap_slots + m.n.heap_slots / 2) * sizeof(float);     // heaps + node lengths
  if (heap_smem <= kCombineSmemMax) {
    if (heap_smem > 48 * 1024 &&
        cudaFuncSetAttribute(noise_combine_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kCombineSmemMax) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
    noise_combine_kernel<true><<<(unsigned)B, 1024, heap_smem, s>>>(m.n);
  } else {
    noise_combine_kernel<false><<<(unsigned)B, 1024, 0, s>>>(m.n);
  }
  const unsigned chunks = (unsigned)((max_len + 3 + kMixChunk - 1) / kMixChunk);   // + 3: a clip may start 3 past a group boundary
  // one wave of CTAs, each striding over its clip's chunks
  int resident = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, noise_mix_kernel<false>, 256, 0) != cudaSuccess || resident < 1) {
    cudaGetLastError();
    resident = 1;
  }
  unsigned per_clip = (unsigned)((int64_t)resident * kNumSMs / B);
  if (per_clip < 1) per_clip = 1;
  if (per_clip > chunks) per_clip = chunks;
  noise_mix_kernel<false><<<dim3(per_clip, (unsigned)B), 256, 0, s>>>(m);
  // the rescale pass is a no-op for a clip that stayed inside int16: a few CTAs per clip, striding
  noise_mix_kernel<true><<<dim3(chunks < 16u ? chunks : 16u, (unsigned)B), 256, 0, s>>>(m);
  count_launch(4);
  return check_launch();
}
