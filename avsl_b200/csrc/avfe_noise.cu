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
// Algorithmic bytes per sample: 4 (clean) + 4 (noise, at most Ln of them) + 2 (int16 out); the
// design reads both waveforms twice (sum of squares, then the mix), each pass touching every byte
// once, which bounds it near half of the HBM roofline (measured: 33 %).
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
#pragma unroll
  for (int t = 0; t < kLeafMax / 8; ++t) {
    v[t] = 0.f;
    if ((uint32_t)(8 * t) < body) v[t] = WRAP ? src[(start + (uint32_t)(8 * t + j)) % period] : mine[8 * t];
  }
  tail = 0.f;
  if (body + (uint32_t)j < len) tail = WRAP ? src[(start + body + (uint32_t)j) % period] : mine[body];
}

__device__ __forceinline__ float piece_sum(const float (&v)[kLeafMax / 8], float tail) {
  float r = __fmul_rn(v[0], v[0]);
#pragma unroll
  for (int t = 1; t < kLeafMax / 8; ++t) r = __fadd_rn(r, __fmul_rn(v[t], v[t]));   // r >= +0: an absent element adds +0, exact
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
  r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
  tail = __fmul_rn(tail, tail);
#pragma unroll
  for (int u = 0; u < 7; ++u) r = __fadd_rn(r, __shfl_sync(0xffffffffu, tail, u, 8));
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
