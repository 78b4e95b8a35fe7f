// post_extract_proj fused behind the fusion block: the one dense contraction on the path.
// Reference: avsl/modules/av_hubert_encoder.py:315-334
//     features = cat([fa, fv], 1)                  [B, 2C, T]     (missing modality zero-filled)
//     features = features.transpose(1, 2)          [B, T, 2C]
//     features = self.layer_norm(features)         LayerNorm(2C), float32 arithmetic
//     features = self.post_extract_proj(features)  nn.Linear(2C -> D), fp16 / bf16 under `precision: 16`
// [B, C, T] x 2 (fp16 / bf16)  ->  [B, T, D], without materialising the fused, transposed or
// normalised tensors.
//
// LayerNorm is folded around the GEMM so that the tensor cores consume the RAW features:
//     out[m, n] = sum_k ((x[m,k] - mu_m) * r_m * g_k + b_k) * W[n,k] + bias[n]
//               = r_m * (acc[m,n] - mu_m * s_n) + c_n,      acc = x @ W'^T,  W'[n,k] = g_k * W[n,k]
//     s_n = sum_k W'[n,k],  c_n = sum_k b_k * W[n,k] + bias[n]        (pep_fold_kernel, once per weights)
//     mu_m, r_m: per (b, t) over the 2C channels                      (pep_stats_kernel)
// x is stored [B, C, T], i.e. M-major for the GEMM (consecutive t are contiguous): the A operand is
// fed to tcgen05.mma as an MN-major tile, straight from a 3-D tensor-map TMA box (no transpose pass).
// That needs the time rows to keep 16-byte alignment: the row pitch (elements between consecutive
// channels) must be a multiple of 8 -- T = 750 lives in a [B, C, 752] allocation.
//
// pep_gemm_kernel: persistent, one CTA per SM, 128 x 256 output tiles, BLOCK_K = 64, 4-stage
// TMA -> shared ring (48 KB per stage), accumulators in TMEM (2 x 256 columns: the epilogue of
// tile i overlaps the MMAs of tile i+1).  Warp roles: 0 = TMA producer, 1 = MMA issuer (one
// elected lane, tcgen05.mma.cta_group::1.kind::f16, M 128 x N 256 x K 16), 2 = TMEM allocator,
// 4..7 = epilogue (tcgen05.ld 32x32b, LayerNorm fold, cast, store).  K blocks of a masked-out
// modality are skipped by producer and issuer alike (zero-fill contributes nothing).
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "avfe_common.cuh"

namespace avfe {
namespace pep {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4;
constexpr int kABytes = kBM * kBK * 2;            // 16 KB: two 64-wide MN atoms of 64 K rows x 128 B
constexpr int kBBytes = kBN * kBK * 2;            // 32 KB: 256 rows x 128 B
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 256;
constexpr int kTmemCols = 512;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nPEP_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra PEP_WAIT;\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (tcgen05): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D f32, A/B f16 (0) or bf16 (1), A MN-major, B K-major, N 256, M 128
__host__ __device__ constexpr uint32_t make_idesc(int ab_format) {
  return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | (1u << 15) | (0u << 16) |
         ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

struct GemmArgs {
  const float2* stats;      // [B*T] (mean, rstd)
  const float* s;           // [D]  sum_k W'[n,k]
  const float* c;           // [D]  sum_k beta_k W[n,k] + bias[n]
  const uint8_t* mask;      // [B,2] or nullptr
  void* out;                // [B, T, D] f16 / bf16
  int B, C, T, D;
  int tiles_per_sample, n_tiles_n, n_tiles;
  int bf16;
  int experiment;
};

struct Smem {
  uint64_t full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ unsigned sample_mask(const uint8_t* mask, int b) {
  return mask ? ((mask[2 * b] ? 1u : 0u) | (mask[2 * b + 1] ? 2u : 0u)) : 3u;
}

__global__ void __launch_bounds__(kThreads, 1)
pep_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_v,
                const __grid_constant__ CUtensorMap map_w, const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t pep_smem[];
  // 128-byte swizzle atoms repeat every 1024 bytes: align the tiles by hand (1 KB of slack is allocated)
  uint8_t* tiles = pep_smem + ((1024u - (smem_u32(pep_smem) & 1023u)) & 1023u);   // kStages x (A 16 KB | B 32 KB)
  float* sc = reinterpret_cast<float*>(tiles + kStages * kStageBytes);     // s[D] then c[D]
  Smem& sm = *reinterpret_cast<Smem*>(tiles + kStages * kStageBytes + 2 * (size_t)g.D * sizeof(float));
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < g.D; i += kThreads) { sc[i] = g.s[i]; sc[g.D + i] = g.c[i]; }
  if (wid == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (wid == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const int kb_per_mod = g.C / kBK;                                       // K blocks per modality

  if (wid == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int nt = tile % g.n_tiles_n, mb = tile / g.n_tiles_n;
        const int b = mb / g.tiles_per_sample, t0 = (mb % g.tiles_per_sample) * kBM;
        const unsigned m = sample_mask(g.mask, b);
        for (int kb = 0; kb < 2 * kb_per_mod; ++kb) {
          const int mod = kb / kb_per_mod;
          if (!((m >> mod) & 1u)) continue;
          mbar_wait(&sm.empty[stage], phase ^ 1u);
          uint8_t* a_dst = tiles + stage * kStageBytes;
          uint8_t* b_dst = a_dst + kABytes;
          mbar_expect_tx(&sm.full[stage], kStageBytes);
          const CUtensorMap* map = mod ? &map_v : &map_a;
          const int c0 = (kb - mod * kb_per_mod) * kBK;
          tma_load_3d(a_dst, map, &sm.full[stage], t0, c0, b);                  // t0 .. t0+63
          tma_load_3d(a_dst + kABytes / 2, map, &sm.full[stage], t0 + 64, c0, b);   // t0+64 .. t0+127
          tma_load_2d(b_dst, &map_w, &sm.full[stage], kb * kBK, nt * kBN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (wid == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(g.bf16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int mb = tile / g.n_tiles_n;
        const int b = mb / g.tiles_per_sample;
        const unsigned m = sample_mask(g.mask, b);
        mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1u);                         // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < 2 * kb_per_mod; ++kb) {
          if (!((m >> (kb / kb_per_mod)) & 1u)) continue;
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(tiles + stage * kStageBytes), b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // A (MN-major, 128B swizzle): 16 K rows = 2 groups of 8 x 128 B; atoms 8 KB apart
            const uint64_t adesc = make_desc(a_base + k * 2048, kABytes / 2, 1024);
            // B (K-major, 128B swizzle): 16 K elements = 32 B inside the 128-B row; 8-row groups 1 KB apart
            const uint64_t bdesc = make_desc(b_base + k * 32, 0, 1024);
            tc_mma_f16(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(&sm.empty[stage]);                                          // stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&sm.tmem_full[acc]);                                          // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (wid >= 4) {
    // ===================== epilogue: TMEM -> LayerNorm fold -> global =====================
    const int ew = wid - 4;                                                     // TMEM lanes 32*ew .. 32*ew+31
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
      const int nt = tile % g.n_tiles_n, mb = tile / g.n_tiles_n;
      const int b = mb / g.tiles_per_sample, t0 = (mb % g.tiles_per_sample) * kBM;
      const int t = t0 + ew * 32 + lane;
      const bool ok = t < g.T;
      float2 st = make_float2(0.f, 0.f);
      if (ok) st = g.stats[(int64_t)b * g.T + t];
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kBN);
      uint16_t* orow = static_cast<uint16_t*>(g.out) + ((int64_t)b * g.T + t) * g.D + nt * kBN;
      for (int ch = 0; ch < kBN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        if (ok) {
          const float* s = sc + nt * kBN + ch * 32;
          const float* c = s + g.D;
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float y0 = fmaf(st.y, __uint_as_float(v[2 * j]) - st.x * s[2 * j], c[2 * j]);
            const float y1 = fmaf(st.y, __uint_as_float(v[2 * j + 1]) - st.x * s[2 * j + 1], c[2 * j + 1]);
            if (g.bf16) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
              packed[j] = *reinterpret_cast<const uint32_t*>(&h);
            } else {
              const __half2 h = __floats2half2_rn(y0, y1);
              packed[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Two CTAs of one cluster (same TPC) share a 256 x 256 tile: CTA r stages the A rows of its own 128
// time steps and HALF of the W' rows (128 of the 256 output channels); one tcgen05.mma.cta_group::2
// (M 256, N 256, K 16), issued by the leader CTA, reads both CTAs' shared memory and accumulates
// into both CTAs' TMEM.  Per CTA and K block 32 KB are staged instead of 48 KB: the single-CTA
// kernel pulls 14 TB/s out of L2, which is what bounds it.
constexpr int kStages2 = 6;
constexpr int kB2Bytes = (kBN / 2) * kBK * 2;       // 16 KB: this CTA's 128 W' rows
constexpr int kStage2Bytes = kABytes + kB2Bytes;    // 32 KB

struct Smem2 {
  uint64_t full[kStages2], empty[kStages2], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data into this CTA's shared memory, completion on the LEADER's barrier
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc2_commit_both(uint64_t* bar) {      // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc2(int ab_format) {      // M 256 across the pair
  return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | (1u << 15) | (0u << 16) |
         ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
}

// g.tiles_per_sample here counts PAIR tiles (256 time steps) per sample; map_w's box is 64 x 128.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pep_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ CUtensorMap map_w, const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t pep_smem[];
  uint8_t* tiles = pep_smem + ((1024u - (smem_u32(pep_smem) & 1023u)) & 1023u);   // same offset in both CTAs
  float* sc = reinterpret_cast<float*>(tiles + kStages2 * kStage2Bytes);
  Smem2& sm = *reinterpret_cast<Smem2*>(tiles + kStages2 * kStage2Bytes + 2 * (size_t)g.D * sizeof(float));
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  for (int i = threadIdx.x; i < g.D; i += kThreads) { sc[i] = g.s[i]; sc[g.D + i] = g.c[i]; }
  if (wid == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (wid == 1 && lane == 0) {
    for (int s = 0; s < kStages2; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full[a], 1); mbar_init(&sm.tmem_empty[a], 8); }   // 4 warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const int kb_per_mod = g.C / kBK;

  if (wid == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < g.n_tiles; tile += n_pairs) {
        const int nt = tile % g.n_tiles_n, mb = tile / g.n_tiles_n;
        const int b = mb / g.tiles_per_sample, t0 = (mb % g.tiles_per_sample) * (2 * kBM) + (int)rank * kBM;
        const unsigned m = sample_mask(g.mask, b);
        for (int kb = 0; kb < 2 * kb_per_mod; ++kb) {
          const int mod = kb / kb_per_mod;
          if (!((m >> mod) & 1u)) continue;
          mbar_wait(&sm.empty[stage], phase ^ 1u);      // the pair's MMAs have retired this stage (both CTAs are told)
          uint8_t* a_dst = tiles + stage * kStage2Bytes;
          uint8_t* b_dst = a_dst + kABytes;
          const uint32_t bar = map_to_rank(&sm.full[stage], 0);
          if (leader) mbar_expect_tx(&sm.full[stage], 2 * kStage2Bytes);     // both CTAs' bytes land on the leader's barrier
          const CUtensorMap* map = mod ? &map_v : &map_a;
          const int c0 = (kb - mod * kb_per_mod) * kBK;
          tma2_load_3d(a_dst, map, bar, t0, c0, b);
          tma2_load_3d(a_dst + kABytes / 2, map, bar, t0 + 64, c0, b);
          tma2_load_2d(b_dst, &map_w, bar, kb * kBK, nt * kBN + (int)rank * (kBN / 2));
          if (++stage == kStages2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (wid == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc2(g.bf16) & ~((uint32_t)g.experiment << 15);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < g.n_tiles; tile += n_pairs) {
        const int mb = tile / g.n_tiles_n;
        const int b = mb / g.tiles_per_sample;
        const unsigned m = sample_mask(g.mask, b);
        mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1u);          // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < 2 * kb_per_mod; ++kb) {
          if (!((m >> (kb / kb_per_mod)) & 1u)) continue;
          mbar_wait(&sm.full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(tiles + stage * kStage2Bytes), b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = make_desc(a_base + k * 2048, kABytes / 2, 1024);
            const uint64_t bdesc = make_desc(b_base + k * 32, 0, 1024);
            tc2_mma_f16(d_tmem, adesc, bdesc, idesc, accumulate);
            accumulate = 1;
          }
          tc2_commit_both(&sm.empty[stage]);
          if (++stage == kStages2) { stage = 0; phase ^= 1u; }
        }
        tc2_commit_both(&sm.tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (wid >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int ew = wid - 4;
    int acc = 0; uint32_t acc_phase = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the moments kernel in front of this launch has completed
    for (int tile = pair; tile < g.n_tiles; tile += n_pairs) {
      const int nt = tile % g.n_tiles_n, mb = tile / g.n_tiles_n;
      const int b = mb / g.tiles_per_sample, t0 = (mb % g.tiles_per_sample) * (2 * kBM) + (int)rank * kBM;
      const int t = t0 + ew * 32 + lane;
      const bool ok = t < g.T;
      float2 st = make_float2(0.f, 0.f);
      if (ok) st = g.stats[(int64_t)b * g.T + t];
      mbar_wait(&sm.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * kBN);
      uint16_t* orow = static_cast<uint16_t*>(g.out) + ((int64_t)b * g.T + t) * g.D + nt * kBN;
      for (int ch = 0; ch < kBN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        if (ok) {
          const float* s = sc + nt * kBN + ch * 32;
          const float* c = s + g.D;
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float y0 = fmaf(st.y, __uint_as_float(v[2 * j]) - st.x * s[2 * j], c[2 * j]);
            const float y1 = fmaf(st.y, __uint_as_float(v[2 * j + 1]) - st.x * s[2 * j + 1], c[2 * j + 1]);
            if (g.bf16) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
              packed[j] = *reinterpret_cast<const uint32_t*>(&h);
            } else {
              const __half2 h = __floats2half2_rn(y0, y1);
              packed[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_rank(&sm.tmem_empty[acc], 0));   // the issuer lives in the leader CTA
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody leaves (or frees TMEM) while the peer may still signal it
  if (wid == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ LayerNorm moments per (b, t)
// [B, C, T] x 2 (row pitch `pitch` elements) -> (mean, rstd) over the 2C fused channels.  A CTA owns
// 64 consecutive time steps of one sample; a warp reads 4 channel rows x 128 bytes per instruction
// (lane = 8 time steps, 16-byte loads), shifted moments (K = first channel's value) as in fuse_ln_kernel.
template <typename T>
__device__ __forceinline__ float cvt(T v);
template <> __device__ __forceinline__ float cvt<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float cvt<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256, 3)
pep_stats_kernel(const T* __restrict__ fa, const T* __restrict__ fv, const uint8_t* __restrict__ mask, int C, int Tn,
                 int64_t pitch, float eps, float2* __restrict__ stats) {
  __shared__ float red[8][2][64];
  // programmatic dependent launch: the GEMM behind this kernel needs the moments only in its epilogue, so
  // its CTAs may take the SMs this grid's last wave leaves free (they wait with griddepcontrol.wait)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tl = (lane & 7) * 8;                       // this lane's 8 time steps inside the 64
  const int rsub = lane >> 3;                          // which of the warp's 4 channel rows
  const unsigned m = sample_mask(mask, b);
  float sum[8], sq[8], K[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sum[j] = 0.f; sq[j] = 0.f; K[j] = 0.f; }
  const bool full = t0 + tl + 8 <= Tn;
  auto load8 = [&](const T* base, int c, float (&x)[8]) {
    const T* p = base + ((int64_t)b * C + c) * pitch + t0 + tl;
    if (full) {
      const uint4 q = *reinterpret_cast<const uint4*>(p);
      const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = cvt<T>(e[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (t0 + tl + j < Tn) ? cvt<T>(p[j]) : 0.f;
    }
  };
  // the shift: channel 0 of the fused tensor (audio channel 0, or 0 when audio is masked out)
  if (m & 1u) load8(fa, 0, K);
  for (int mod = 0; mod < 2; ++mod) {
    if (!((m >> mod) & 1u)) {                           // zero-filled modality: C zeros per time step
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float n = (float)((C - (wid * 4 + rsub) + 31) / 32);   // channels this lane would have visited
        sum[j] -= n * K[j]; sq[j] += n * K[j] * K[j];
      }
      continue;
    }
    const T* base = mod ? fv : fa;
    // eight independent 16-byte loads per lane in flight (the loop is latency-bound otherwise)
    int c = wid * 4 + rsub;
    for (; c + 7 * 32 < C; c += 8 * 32) {
      if (full) {                                        // the loads stay packed (32 registers instead of 64) until they are used
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = *reinterpret_cast<const uint4*>(base + ((int64_t)b * C + c + 32 * u) * pitch + t0 + tl);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const T* e = reinterpret_cast<const T*>(&q[u]);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {               // packed fp32x2 (FADD2 / FFMA2), the scalar roundings
            const float2 d = __fadd2_rn(make_float2(cvt<T>(e[j]), cvt<T>(e[j + 1])), make_float2(-K[j], -K[j + 1]));
            const float2 t1 = __fadd2_rn(make_float2(sum[j], sum[j + 1]), d);
            const float2 t2 = __ffma2_rn(d, d, make_float2(sq[j], sq[j + 1]));
            sum[j] = t1.x; sum[j + 1] = t1.y; sq[j] = t2.x; sq[j + 1] = t2.y;
          }
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
          float x[8];
          load8(base, c + 32 * u, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = x[j] - K[j]; sum[j] += d; sq[j] += d * d; }
        }
      }
    }
    for (; c < C; c += 32) {
      float x[8];
      load8(base, c, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = x[j] - K[j]; sum[j] += d; sq[j] += d * d; }
    }
  }
  // lanes l, l^8, l^16, l^24 hold the same time steps
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], 8);  sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 8);
    sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], 16); sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[wid][0][tl + j] = sum[j]; red[wid][1][tl + j] = sq[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 64 && t0 + threadIdx.x < Tn) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { s += red[w][0][threadIdx.x]; q += red[w][1][threadIdx.x]; }
    // K of this time step
    float k0 = 0.f;
    if (m & 1u) k0 = cvt<T>(fa[((int64_t)b * C) * pitch + t0 + threadIdx.x]);
    const float inv = 1.0f / (float)(2 * C);
    const float dm = s * inv;
    const float var = fmaxf(q * inv - dm * dm, 0.f);
    stats[(int64_t)b * Tn + t0 + threadIdx.x] = make_float2(k0 + dm, rsqrtf(var + eps));
  }
}

// ------------------------------------------------------------------ weight folding (once per weights)
// W [D, K] (f32 master or f16/bf16), gamma/beta [K] f32, bias [D] f32 ->
// W' [D, K] in the GEMM dtype, s [D] = sum_k f32(W'[n,k]), c [D] = sum_k beta_k * W[n,k] + bias[n]
template <typename TW, typename TO>
__global__ void __launch_bounds__(256)
pep_fold_kernel(const TW* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ bias, int K, TO* __restrict__ Wf, float* __restrict__ s, float* __restrict__ c) {
  __shared__ float red[2][8];
  const int n = blockIdx.x;
  float ss = 0.f, cc = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float w = (float)W[(int64_t)n * K + k];
    const TO wf = (TO)(w * (gamma ? gamma[k] : 1.0f));
    Wf[(int64_t)n * K + k] = wf;
    ss += (float)wf;
    cc += (beta ? beta[k] : 0.0f) * w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); cc += __shfl_xor_sync(0xffffffffu, cc, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = cc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b2 = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b2 += red[1][w]; }
    s[n] = a;
    c[n] = b2 + (bias ? bias[n] : 0.0f);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace pep
}  // namespace avfe

using namespace avfe;

extern "C" size_t avfe_proj_fold_bytes(int64_t D, int64_t K) {
  if (D <= 0 || K <= 0) return 0;
  return (size_t)D * K * 2 + 2 * (size_t)D * sizeof(float);       // W' (16-bit) | s | c
}

extern "C" int avfe_proj_fold(const void* W, int w_dtype, const float* gamma, const float* beta, const float* bias,
                              int64_t D, int64_t K, int dtype, void* folded, avfe_stream_t stream) {
  if (D <= 0 || K <= 0 || !W || !folded) return AVFE_ERR_INVALID_ARG;
  if (dtype != AVFE_F16 && dtype != AVFE_BF16) return AVFE_ERR_UNSUPPORTED;
  if (!aligned16(folded)) return AVFE_ERR_ALIGNMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* s = reinterpret_cast<float*>(static_cast<char*>(folded) + (size_t)D * K * 2);
  float* c = s + D;
#define AVFE_FOLD(TW, TO) pep::pep_fold_kernel<TW, TO><<<(unsigned)D, 256, 0, st>>>(static_cast<const TW*>(W), gamma, beta, bias, (int)K, static_cast<TO*>(folded), s, c)
  if (dtype == AVFE_F16) {
    if (w_dtype == AVFE_F32) AVFE_FOLD(float, __half);
    else if (w_dtype == AVFE_F16) AVFE_FOLD(__half, __half);
    else return AVFE_ERR_UNSUPPORTED;
  } else {
    if (w_dtype == AVFE_F32) AVFE_FOLD(float, __nv_bfloat16);
    else if (w_dtype == AVFE_BF16) AVFE_FOLD(__nv_bfloat16, __nv_bfloat16);
    else return AVFE_ERR_UNSUPPORTED;
  }
#undef AVFE_FOLD
  count_launch();
  return check_launch();
}

extern "C" size_t avfe_fuse_ln_proj_workspace_bytes(int64_t B, int64_t T) {
  if (B <= 0 || T <= 0) return 16;
  return (size_t)B * T * sizeof(float2);
}

extern "C" int avfe_fuse_ln_proj(const void* fa, const void* fv, const uint8_t* mask, int dtype, int64_t B, int64_t C,
                                 int64_t T, int64_t t_pitch, const void* folded, int64_t D, float eps, void* out,
                                 void* workspace, size_t workspace_bytes, avfe_stream_t stream) {
  if (B < 0 || C <= 0 || T < 0 || D <= 0) return AVFE_ERR_INVALID_ARG;
  if (B == 0 || T == 0) return AVFE_OK;
  if (!fa || !fv || !folded || !out) return AVFE_ERR_INVALID_ARG;
  if (dtype != AVFE_F16 && dtype != AVFE_BF16) return AVFE_ERR_UNSUPPORTED;
  // tensor-map TMA: 16-byte aligned rows; tile shapes: C a multiple of BLOCK_K, D of BLOCK_N
  if (t_pitch < T || (t_pitch % 8) != 0 || (C % pep::kBK) != 0 || (D % pep::kBN) != 0 || D > 4096) return AVFE_ERR_UNSUPPORTED;
  if (!aligned16(fa) || !aligned16(fv) || !aligned16(folded) || !aligned16(out)) return AVFE_ERR_ALIGNMENT;
  if (!workspace || workspace_bytes < avfe_fuse_ln_proj_workspace_bytes(B, T) || !aligned16(workspace)) return AVFE_ERR_WORKSPACE;
  if (B > 65535 || T > (1 << 24)) return AVFE_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pep::EncodeTiledFn enc = pep::encode_fn();
  if (!enc) return AVFE_ERR_CUDA;

  const int64_t K = 2 * C;
  const CUtensorMapDataType dt = (dtype == AVFE_F16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap map_a, map_v, map_w;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)t_pitch * 2, (cuuint64_t)C * t_pitch * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)pep::kBK, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map_a, dt, 3, const_cast<void*>(fa), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AVFE_ERR_CUDA;
    if (enc(&map_v, dt, 3, const_cast<void*>(fv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AVFE_ERR_CUDA;
  }
  // CTA pairs (cta_group::2) unless AVFE_PEP_SINGLE_CTA is set (kept for A/B measurements)
  static const bool single_cta = (getenv("AVFE_PEP_SINGLE_CTA") != nullptr);
  {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)D};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)pep::kBK, (cuuint32_t)(single_cta ? pep::kBN : pep::kBN / 2)};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map_w, dt, 2, const_cast<void*>(folded), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return AVFE_ERR_CUDA;
  }

  float2* stats = static_cast<float2*>(workspace);
  dim3 sgrid((unsigned)((T + 63) / 64), (unsigned)B);
  if (dtype == AVFE_F16)
    pep::pep_stats_kernel<__half><<<sgrid, 256, 0, st>>>(static_cast<const __half*>(fa), static_cast<const __half*>(fv), mask,
                                                         (int)C, (int)T, t_pitch, eps, stats);
  else
    pep::pep_stats_kernel<__nv_bfloat16><<<sgrid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(fa),
                                                                static_cast<const __nv_bfloat16*>(fv), mask, (int)C, (int)T,
                                                                t_pitch, eps, stats);
  count_launch();

  pep::GemmArgs g;
  g.stats = stats;
  g.s = reinterpret_cast<const float*>(static_cast<const char*>(folded) + (size_t)D * K * 2);
  g.c = g.s + D;
  g.mask = mask; g.out = out; g.B = (int)B; g.C = (int)C; g.T = (int)T; g.D = (int)D;
  g.n_tiles_n = (int)(D / pep::kBN);
  g.bf16 = (dtype == AVFE_BF16) ? 1 : 0;
  g.experiment = getenv("AVFE_PEP_EXPERIMENT") ? 1 : 0;
  if (single_cta) {
    g.tiles_per_sample = (int)((T + pep::kBM - 1) / pep::kBM);
    g.n_tiles = (int)B * g.tiles_per_sample * g.n_tiles_n;
    const size_t smem = (size_t)pep::kStages * pep::kStageBytes + 2 * (size_t)D * sizeof(float) + sizeof(pep::Smem) + 1024;
    if (cudaFuncSetAttribute(pep::pep_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
    const int grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
    pep::pep_gemm_kernel<<<grid, pep::kThreads, smem, st>>>(map_a, map_v, map_w, g);
  } else {
    g.tiles_per_sample = (int)((T + 2 * pep::kBM - 1) / (2 * pep::kBM));          // pair tiles: 256 time steps
    g.n_tiles = (int)B * g.tiles_per_sample * g.n_tiles_n;
    const size_t smem = (size_t)pep::kStages2 * pep::kStage2Bytes + 2 * (size_t)D * sizeof(float) + sizeof(pep::Smem2) + 1024;
    if (cudaFuncSetAttribute(pep::pep_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
    const int pairs = g.n_tiles < kNumSMs / 2 ? g.n_tiles : kNumSMs / 2;
    // launched with programmatic stream serialization: producer / issuer warps start while the moments
    // kernel drains, the epilogue warps wait for it (griddepcontrol.wait)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(pep::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, pep::pep_gemm2_kernel, map_a, map_v, map_w, g) != cudaSuccess) {   // __cluster_dims__(2,1,1)
      cudaGetLastError();
      return AVFE_ERR_CUDA;
    }
  }
  count_launch();
  return check_launch();
}
