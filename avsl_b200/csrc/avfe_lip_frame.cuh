// lip_frame_kernel: gray conversion, transform fit and ROI warp of a batch in ONE persistent
// launch in which every BGR byte is read from HBM exactly once.  One 1024-thread CTA per SM owns
// whole frames (f = blockIdx.x + k * gridDim.x); three warp roles, one CTA-wide barrier at start-up:
//
//   tform warps  (2) window-smoothed similarity fit, cut_patch origin and source footprint of
//                the CTA's frames, a few frames ahead of everybody else (tform_frame, shared
//                with tform_kernel), published through a 4-deep descriptor ring
//   stream warps (8 for an 88-px window, 9 for 96) pull the frame through shared memory in
//                1024-px chunks with one bulk async copy (TMA, cp.async.bulk + mbarrier, L2
//                evict-first hint) per chunk and a 4- / 3-deep ring per warp, convert BGR->gray
//                (integer dp2a), store the gray frame with 16-byte stores, and drop the gray
//                pixels that lie inside the frame's ROI footprint into one of two
//                shared-memory footprint tiles
//   blend warps  (22 / 21) float64 bilinear blend of the previous frame's ROI from that tile in
//                skimage's operation order: u8 ROI and/or normalised f32 centre crop
//
// Hand-over: mbarriers (ring FULL per stage, descriptor FULL/EMPTY, tile EMPTY) and one hardware
// named barrier per tile slot for tile FULL, so that the waiting blend warps do not poll;
// the FP64 work of frame k hides under the memory time of frame k+1.
// Included by avfe_lip.cu only (after avfe_lip_queue.cuh, whose gray and blend helpers it uses).
#pragma once

namespace avfe {

// tuning knobs (profiles/lip_sweep.sh rebuilds the library with other values: -DAVFE_LIP_...)
#ifndef AVFE_LIP_PHASES_88
#define AVFE_LIP_PHASES_88 8
#endif
#ifndef AVFE_LIP_PHASES_96
#define AVFE_LIP_PHASES_96 7
#endif
#ifndef AVFE_LIP_RING_88
#define AVFE_LIP_RING_88 4
#endif
#ifndef AVFE_LIP_RING_96
#define AVFE_LIP_RING_96 3
#endif
#ifndef AVFE_LIP_SLOTS_88
#define AVFE_LIP_SLOTS_88 2
#endif
#ifndef AVFE_LIP_SLOTS_96
#define AVFE_LIP_SLOTS_96 2
#endif
#ifndef AVFE_LIP_EVICT_FIRST
#define AVFE_LIP_EVICT_FIRST 1
#endif
#ifndef AVFE_LIP_DESC_RING
#define AVFE_LIP_DESC_RING 4
#endif
constexpr int kTformRoleWarps = 2;
constexpr int kDescRing = AVFE_LIP_DESC_RING;
constexpr int kFrameTilePx = 16384;         // staged footprint capacity per slot (one byte per pixel)

template <int SPAN>
struct FrameRoles {
  static constexpr int kSide = SPAN;
  // thread = (column, row phase).  88-px window: 8 phases of 11 rows (704 threads = 22 blend warps, 8 stream
  // warps); 96-px window: 7 phases of 14 rows (672 threads = 21 blend warps, 9 stream warps).  The split,
  // the ring depth and the number of footprint slots were swept on one box (profiles/lip_sweep.sh,
  // profiles/r02/lip_sweep.txt; 16,500 frames of 224 x 224 / 8,000 of 352 x 288), phases/ring/slots:
  //   88 px: 8/6/3 (round 1) 0.661 / 0.581 ms; 7/4/3 0.654 / 0.565; 7/3/3 0.635 / 0.546; 7/3/2 0.626 / 0.540;
  //          8/4/3 0.634 / 0.555; 8/4/2 0.618 / 0.546 (chosen); 8/3/2 0.648; 8/5/2 0.633; 7/2/3 0.704; 9/6/2 0.733
  //   96 px: 6/4/3 (round 1) 0.714 ms; 8/4/3 0.764; 7/4/3 0.693; 7/3/3 0.682; 7/3/2 0.680 (chosen); 6/3/2 0.695
  // About 90-96 KB of bulk copies in flight per SM is the optimum: 144 KB (round 1) is 6 % slower -- deeper
  // read queues only get in the way of the gray / feature write streams -- and 60-72 KB starves the stream.
  static constexpr int kPhases = (SPAN == 96) ? AVFE_LIP_PHASES_96 : AVFE_LIP_PHASES_88;
  static constexpr int kBlendActive = kPhases * SPAN;                 // 704 or 672 threads with pixels
  static constexpr int kBlendWarps = (kBlendActive + 31) / 32;        // 22 or 21
  static constexpr int kRowsMax = (SPAN + kPhases - 1) / kPhases;     // 11 or 14
  // bulk copies in flight per stream warp (x 3 KB): 8 x 4 = 96 KB / 9 x 3 = 81 KB per SM
  static constexpr int kRing = (SPAN == 96) ? AVFE_LIP_RING_96 : AVFE_LIP_RING_88;
  // footprint tiles: the stream may run this many frames ahead of the blend (what fits)
  static constexpr int kSlots = (SPAN == 96) ? AVFE_LIP_SLOTS_96 : AVFE_LIP_SLOTS_88;
  static constexpr int kBlendThreads = kBlendWarps * 32;
  static constexpr int kStreamWarps = 32 - kBlendWarps - kTformRoleWarps;   // 8 or 9
  static constexpr int kStreamFirst = kBlendWarps;                    // warp index of the first stream warp
  static constexpr int kTformFirst = 32 - kTformRoleWarps;            // highest warp ids: scheduled first
  static constexpr int kHandoverThreads = 32 * (kBlendWarps + kStreamWarps);   // tile FULL barriers
};

struct FrameJob {
  LipJob lip;                // frames, sizes, outputs (xf / counter unused)
  TformArgs tf;
  int chunks_per_frame;      // ceil(H*W / 1024)
  int groups_per_frame;      // H*W / 16
  int row_groups;            // W / 16
  unsigned row_magic;        // exact g / row_groups for g < 2^22 (umulhi)
};

template <int SPAN>
struct FrameSmem {
  double lut255[256 * 16];                    // k / 255.0, 16 interleaved copies (see FusedSmem)
  float lutn[256];                            // ((k/255) - mean) / std in float32
  unsigned long long ring_full[FrameRoles<SPAN>::kStreamWarps][FrameRoles<SPAN>::kRing];
  unsigned long long desc_full[kDescRing], desc_empty[kDescRing];
  unsigned long long tile_empty[FrameRoles<SPAN>::kSlots];
  FrameXform desc[kDescRing];
  int64_t dst[kDescRing];                     // f32 output slot of the frame (collation), < 0: dropped
  __align__(16) uint8_t tile[FrameRoles<SPAN>::kSlots][kFrameTilePx];    // gray footprint, one byte per pixel
  __align__(16) uint4 ring[FrameRoles<SPAN>::kStreamWarps][FrameRoles<SPAN>::kRing][kChunkVec];
};

// ---------------------------------------------------------------- mbarrier / bulk-copy PTX
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AVFE_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra AVFE_DONE;\n"
      "bra AVFE_WAIT;\n"
      "AVFE_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)                       // suspend-time hint (ns); the wait is re-armed if it expires
      : "memory");
}
// the same for a warp that is far ahead of its consumers (the tform warps): one probe, then sleep --
// a polling warp costs the stream warps issue slots (10 % of the kernel's instructions were SYNCS /
// BRA / NANOSLEEP of such loops, profiles/r01/lip_frame_kernel_final4_by_line.txt)
__device__ __forceinline__ void mbar_wait_idle(void* bar, unsigned parity, unsigned sleep_ns) {
  for (;;) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    __nanosleep(sleep_ns);
  }
}
// global -> shared bulk copy (TMA, 1-D); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, void* bar,
                                          unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// frames are read exactly once: ask L2 to evict them first
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long p;
#if AVFE_LIP_EVICT_FIRST
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
#else
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
#endif
  return p;
}

// tile FULL: hardware barrier 1 + slot with a compile-time id (a register id would make ptxas
// reserve all 16 barriers); stream warps arrive, blend warps sync
template <bool SYNC>
__device__ __forceinline__ void tile_bar(int slot, int count) {
  if (slot == 0)      { if (SYNC) bar_sync_id<1>(count); else bar_arrive_id<1>(count); }
  else if (slot == 1) { if (SYNC) bar_sync_id<2>(count); else bar_arrive_id<2>(count); }
  else                { if (SYNC) bar_sync_id<3>(count); else bar_arrive_id<3>(count); }
}

// ---------------------------------------------------------------- tform warps
template <int SPAN>
__device__ __forceinline__ void frame_tform_run(const FrameJob& j, FrameSmem<SPAN>& sm, int tw, int lane,
                                                int nk) {
  for (int k = tw; k < nk; k += kTformRoleWarps) {
    const int64_t f = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
    int64_t dst;
    const FrameXform x = tform_frame(j.tf, f, lane, dst);
    const int slot = k % kDescRing, use = k / kDescRing;
    if (use > 0) mbar_wait_idle(&sm.desc_empty[slot], (unsigned)(use - 1) & 1u, 1000u);   // every reader is done with it; a frame takes ~6 us
    if (lane == 0) {
      sm.desc[slot] = x;
      sm.dst[slot] = dst;
      mbar_arrive(&sm.desc_full[slot]);                  // release: the record is visible to the waiters
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- stream warps
template <int SPAN>
__device__ __forceinline__ void frame_stream_run(const FrameJob& j, FrameSmem<SPAN>& sm, int sw, int lane,
                                                 int nk) {
  using R = FrameRoles<SPAN>;
  constexpr int S = R::kStreamWarps;
  const LipJob& L = j.lip;
  const int cpf = j.chunks_per_frame, gpf = j.groups_per_frame;
  const int64_t frame_bytes = (int64_t)gpf * 48, frame_px = (int64_t)gpf * 16;
  uint4(*ring)[kChunkVec] = sm.ring[sw];
  unsigned long long* full = sm.ring_full[sw];

  // issue side: (ik, ic) = next chunk to request and where it lives; the source pointer is
  // advanced incrementally (S chunks ahead inside a frame, to this warp's first chunk of the
  // CTA's next frame at a frame change)
  int ik = (sw < cpf) ? 0 : nk, ic = sw, istage = 0;     // a frame may have fewer chunks than stream warps
  const uint8_t* isrc = L.frames + (int64_t)blockIdx.x * frame_bytes + (int64_t)sw * (kChunkVec * 16);
  const int64_t next_frame = (int64_t)gridDim.x * frame_bytes;
  const unsigned last_bytes = (unsigned)(gpf - (cpf - 1) * 64) * 48u;   // the frame's last chunk may be short
  const bool full_chunks = (gpf & 63) == 0;
  const unsigned long long l2_policy = l2_evict_first_policy();
  auto issue = [&]() {
    if (ik < nk) {
      if (lane == 0) {
        const unsigned bytes = (ic == cpf - 1) ? last_bytes : (unsigned)(kChunkVec * 16);
        mbar_arrive_expect_tx(&full[istage], bytes);
        bulk_load(ring[istage], isrc, bytes, &full[istage], l2_policy);
      }
      ic += S;
      isrc += S * (kChunkVec * 16);
      if (ic >= cpf) { isrc += next_frame - (int64_t)(ic - sw) * (kChunkVec * 16); ic = sw; ++ik; }
    }
    if (++istage == R::kRing) istage = 0;
  };
#pragma unroll 1
  for (int i = 0; i < R::kRing - 1; ++i) issue();

  int stage = 0;
  unsigned phase = 0;                                    // parity of the current pass over the ring
  for (int k = 0; k < nk; ++k) {
    const int64_t f = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
    // this frame's footprint box (the tform warps are ahead) and its tile slot
    const int dslot = k % kDescRing, slot = k % R::kSlots;
    mbar_wait(&sm.desc_full[dslot], (unsigned)(k / kDescRing) & 1u);
    const unsigned box_lo = sm.desc[dslot].box_lo, box_hi = sm.desc[dslot].box_hi;
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.desc_empty[dslot]);
    const bool staged = ((box_lo >> 27) & 1u) != 0;
    const int br0 = (int)(box_lo & 0x1fffu), bcg = (int)((box_lo >> 13) & 0x1fffu) >> 4;
    const int brows = staged ? (int)(box_hi & 0x1fffu) : 0, bpg = (int)((box_hi >> 13) & 0x1fffu) >> 4;
    if (k >= R::kSlots) mbar_wait(&sm.tile_empty[slot], (unsigned)(k / R::kSlots - 1) & 1u);   // blend warps left the slot
    uint4* tile = reinterpret_cast<uint4*>(sm.tile[slot]);
    uint4* gout = reinterpret_cast<uint4*>(L.gray_out + f * frame_px);
    // chunks that hold rows of the footprint box: [c_lo, c_hi] (warp-uniform, once per frame), so
    // that the 68 % of the chunks outside it pay two compares instead of the per-lane box test
    int c_lo = 1, c_hi = 0;
    if (brows > 0) {
      c_lo = (br0 * j.row_groups) >> 6;
      c_hi = ((br0 + brows) * j.row_groups - 1) >> 6;
    }
    for (int c = sw; c < cpf; c += S) {
      mbar_wait(&full[stage], phase);                    // chunk landed
      const uint4* s = ring[stage];
      // lane owns 48 contiguous bytes (16 px) of each of the chunk's two halves
      const uint4 p0 = s[3 * lane], p1 = s[3 * lane + 1], p2 = s[3 * lane + 2];
      const uint4 q0 = s[96 + 3 * lane], q1 = s[96 + 3 * lane + 1], q2 = s[96 + 3 * lane + 2];
      __syncwarp();                                      // every lane has read: the stage may be refilled
      issue();                                           // into the stage consumed one iteration ago
      const uint4 ga = gray16_dp2a(p0, p1, p2), gb = gray16_dp2a(q0, q1, q2);
      const int g0 = c * 64 + lane, g1 = g0 + 32;        // 16-px group indices inside the frame
      if (full_chunks || g0 < gpf) gout[g0] = ga;
      if (full_chunks || g1 < gpf) gout[g1] = gb;
      if (c >= c_lo && c <= c_hi) {
        // groups never straddle rows (W % 16 == 0); deposit those inside the footprint box
        const int ra = (int)__umulhi((unsigned)g0, j.row_magic), ca = g0 - ra * j.row_groups;
        const int rb = (int)__umulhi((unsigned)g1, j.row_magic), cb = g1 - rb * j.row_groups;
        if ((unsigned)(ra - br0) < (unsigned)brows && (unsigned)(ca - bcg) < (unsigned)bpg && g0 < gpf) {
          tile[(ra - br0) * bpg + (ca - bcg)] = ga;
        }
        if ((unsigned)(rb - br0) < (unsigned)brows && (unsigned)(cb - bcg) < (unsigned)bpg && g1 < gpf) {
          tile[(rb - br0) * bpg + (cb - bcg)] = gb;
        }
      }
      if (++stage == R::kRing) { stage = 0; phase ^= 1u; }
    }
    // this warp's part of the footprint is in: hardware barrier 1 + slot, which the blend warps
    // wait on without polling (stream warps only arrive)
    tile_bar<false>(slot, R::kHandoverThreads);
  }
}

// ---------------------------------------------------------------- blend warps
template <int SPAN>
__device__ __forceinline__ void frame_blend_item(const LipJob& j, int64_t f, int64_t slot_f32, const FrameXform& x,
                                                 const uint8_t* tile, const double* lut255, const float* lutn,
                                                 int tid) {
  using R = FrameRoles<SPAN>;
  const Footprint fp = unpack_footprint(x);
  const double* lut = lut255 + (tid & 15);              // this lane's copy of the k/255 table
  const int off = (j.roi - j.crop) / 2;
  const int lo = j.lip_u8 ? 0 : off;
  constexpr int S = SPAN;
  uint8_t* out_u8 = j.lip_u8 ? j.lip_u8 + f * (int64_t)j.roi * j.roi : nullptr;
  float* out_f32 = (j.lip_f32 && slot_f32 >= 0) ? j.lip_f32 + slot_f32 * (int64_t)j.crop * j.crop : nullptr;
  if (out_u8 == nullptr && out_f32 == nullptr) return;  // frame dropped by the collation trim
  auto emit = [&](int r, int c, uint32_t v) {
    if (SPAN == 88) {                                   // centre crop only: window == f32 output
      out_f32[r * 88 + c] = lutn[v];
      return;
    }
    const int pr = lo + r, pc = lo + c;
    if (out_u8) out_u8[pr * j.roi + pc] = (uint8_t)v;
    if (out_f32) {
      const int cr = pr - off, cc = pc - off;
      if ((unsigned)cr < (unsigned)j.crop && (unsigned)cc < (unsigned)j.crop)
        out_f32[cr * j.crop + cc] = lutn[v];
    }
  };
  if (x.r0 < 0) {                                       // clip without any detection: zero ROI
    for (int idx = tid; idx < S * S; idx += R::kBlendThreads) emit(idx / S, idx % S, 0u);
    return;
  }
  // x_ = (M0*c + M1*r) + M2: two roundings per product as in skimage's _transform_affine; the
  // integer coordinates are exact in float64, so tr + kPhases is the next row of this thread exactly
  const double m0 = x.inv[0], m1 = x.inv[1], m2 = x.inv[2], m3 = x.inv[3], m4 = x.inv[4], m5 = x.inv[5];
  const int br0 = fp.r0, bc0 = fp.c0, pitch = fp.pitch;
  if (tid >= R::kBlendActive) return;                   // idle lanes of the last blend warp
  const int rr = tid / S, c = tid - rr * S;             // thread = (column, row phase)
  const double tc = (double)(x.c0 + lo + c);
  const double cx = f64mul(m0, tc), cy = f64mul(m3, tc);
  double tr = (double)(x.r0 + lo + rr);
  if (fp.interior && fp.staged) {
#pragma unroll
    for (int i = 0; i < R::kRowsMax; ++i) {
      if (S % R::kPhases != 0 && rr + R::kPhases * i >= S) break;     // 88 = 14 * 6 + 4
      const double sc = f64add(f64add(cx, f64mul(m1, tr)), m2);
      const double sr = f64add(f64add(cy, f64mul(m4, tr)), m5);
      emit(rr + R::kPhases * i, c, bilinear_interior(sr, sc, tile, pitch, br0, bc0, lut));
      tr = f64add(tr, (double)R::kPhases);
    }
    return;
  }
  // ROI hanging over the frame border, or a footprint too large to stage: bounds-checked taps
  const int H = j.H, W = j.W;
  const uint8_t* img = j.frames + f * (int64_t)H * W * 3;
  const int brows = fp.staged ? fp.rows : 0;
  auto tap = [&](int r, int cc) -> double {
    const int tr_ = r - br0, tc_ = cc - bc0;
    if ((unsigned)tr_ < (unsigned)brows && (unsigned)tc_ < (unsigned)pitch) return lut_at(lut, tile_off(tile[tr_ * pitch + tc_]));
    const uint8_t* p = img + ((int64_t)r * W + cc) * 3;                 // not staged: global tap
    return lut[16 * gray_from_bgr(__ldg(p), __ldg(p + 1), __ldg(p + 2))];
  };
  for (int i = 0; i < R::kRowsMax; ++i) {
    if (rr + R::kPhases * i >= S) break;
    const double sc = f64add(f64add(cx, f64mul(m1, tr)), m2);
    const double sr = f64add(f64add(cy, f64mul(m4, tr)), m5);
    emit(rr + R::kPhases * i, c, (uint32_t)bilinear_u8(sr, sc, H, W, tap));
    tr = f64add(tr, (double)R::kPhases);
  }
}

template <int SPAN>
__device__ __forceinline__ void frame_blend_run(const FrameJob& j, FrameSmem<SPAN>& sm, int tid, int nk) {
  const int lane = tid & 31;
  for (int k = 0; k < nk; ++k) {
    const int64_t f = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
    const int dslot = k % kDescRing, slot = k % FrameRoles<SPAN>::kSlots;
    mbar_wait(&sm.desc_full[dslot], (unsigned)(k / kDescRing) & 1u);
    const FrameXform x = sm.desc[dslot];
    const int64_t slot_f32 = sm.dst[dslot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.desc_empty[dslot]);
    tile_bar<true>(slot, FrameRoles<SPAN>::kHandoverThreads);   // the whole frame has been streamed
    frame_blend_item<SPAN>(j.lip, f, slot_f32, x, sm.tile[slot], sm.lut255, sm.lutn, tid);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.tile_empty[slot]);
  }
}

template <int SPAN>
__global__ void __launch_bounds__(1024, 1)
lip_frame_kernel(const FrameJob j) {
  using R = FrameRoles<SPAN>;
  extern __shared__ __align__(16) unsigned char frame_smem_raw[];
  FrameSmem<SPAN>& sm = *reinterpret_cast<FrameSmem<SPAN>*>(frame_smem_raw);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  // frames of this CTA: blockIdx.x, + gridDim.x, ...
  const int64_t N = j.lip.N;
  const int nk = (int)((N - blockIdx.x + gridDim.x - 1) / gridDim.x);
  if (tid == 0) {
    for (int w = 0; w < R::kStreamWarps; ++w)
      for (int s = 0; s < R::kRing; ++s) mbar_init(&sm.ring_full[w][s], 1u);
    for (int s = 0; s < kDescRing; ++s) {
      mbar_init(&sm.desc_full[s], 1u);
      mbar_init(&sm.desc_empty[s], (unsigned)(R::kStreamWarps + R::kBlendWarps));
    }
    for (int s = 0; s < R::kSlots; ++s) {
      mbar_init(&sm.tile_empty[s], (unsigned)R::kBlendWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < 256; k += 1024) {
    const double q = f64div((double)k, 255.0);
#pragma unroll
    for (int c = 0; c < 16; ++c) sm.lut255[k * 16 + c] = q;
    sm.lutn[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)k, 255.0f), j.lip.mean), j.lip.stdv);
  }
  __syncthreads();                                       // the only CTA-wide barrier
  if (wid >= R::kTformFirst) {
    frame_tform_run<SPAN>(j, sm, wid - R::kTformFirst, lane, nk);
  } else if (wid >= R::kStreamFirst) {
    frame_stream_run<SPAN>(j, sm, wid - R::kStreamFirst, lane, nk);
  } else {
    frame_blend_run<SPAN>(j, sm, tid, nk);
  }
}

}  // namespace avfe
