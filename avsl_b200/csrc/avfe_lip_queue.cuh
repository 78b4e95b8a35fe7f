// lip_fused_kernel: the gray conversion and the ROI warp of a batch in ONE persistent launch,
// one CTA per SM, three warp roles that never wait on a CTA-wide barrier:
//
//   blend warps    (22 for an 88-px window, 24 for 96) float64 bilinear blend of one frame's ROI
//                  in skimage's operation order, u8 ROI + normalised f32 centre crop; thread =
//                  (column, row phase), rows fully unrolled; issue / FP64-pipe bound
//   producer warps (NSLOT: 2 beside the stream warps, 6 without them) pull frames from an atomic
//                  work queue and stage them ahead, one slot per producer warp: frame descriptor
//                  + source footprint fetched with 16-byte cp.async.  Six footprints in flight per
//                  SM are what keeps PCIe busy when `frames` is mapped pinned HOST memory (the
//                  zero-copy e2e mode: only the footprints cross the link).  The blend warps turn the raw BGR footprint into gray
//                  in shared memory themselves (704 threads, two quads each - so there is no
//                  dependency on the stream warps) and release the slot right after that
//   stream warps   (the remaining 8 / 6) BGR->gray over the flat pixel stream in 1024-px chunks,
//                  statically strided, each warp with its own 4-stage cp.async ring (9 KB in
//                  flight per warp while a chunk is converted); HBM-bound, integer dp2a
//
// Producer and blend warps hand slots over with named barriers (FULL/EMPTY per slot); the stream
// warps synchronise with nobody.  The FP64 work hides under the stream's memory time.
// Included by avfe_lip.cu only.
#pragma once

namespace avfe {

constexpr int kStreamSlots = 2;             // footprint slots when the stream warps use the ring space
constexpr int kQueueSlots = 6;              // ... and when they do not (the ring space holds 4 more slots)
constexpr int kMaxSlots = 6;
constexpr int kRingStages = 4;              // 3 chunks (9 KB) per stream warp in flight
constexpr int kChunkVec = 192;              // one stream chunk = 2 groups = 1024 px = 192 uint4 in
constexpr int kMaxStreamWarps = 8;
constexpr unsigned kItemDone = 0xffffffffu;

// warp roles for a compile-time window side (0 = run-time side, laid out like 88) and slot count
template <int SPAN, int NSLOT = kStreamSlots>
struct Roles {
  static constexpr int kSide = SPAN ? SPAN : 88;
  static constexpr int kSlots = NSLOT;
  static constexpr int kProducerWarps = NSLOT;                        // one producer warp per slot
  static constexpr int kProducerThreads = kProducerWarps * 32;
  static constexpr int kBlendWarps = 8 * kSide / 32;                  // 22 or 24
  static constexpr int kBlendThreads = kBlendWarps * 32;
  static constexpr int kStreamWarps = 32 - kBlendWarps - kProducerWarps;   // 8 or 6 (NSLOT == 2 only)
  static constexpr int kHandoverThreads = kBlendThreads + kProducerThreads;   // blend + producer threads
  static constexpr int kSlotThreads = kBlendThreads + 32;   // a slot's barriers: blend warps + its producer warp
};
// named barriers: FULL[slot] = 2 + slot, EMPTY[slot] = 8 + slot, 14 = blend warps only
constexpr int kBarFull = 2, kBarEmpty = 8, kBarBlend = 14;

struct LipJob {
  const uint8_t* frames;   // [N,H,W,channels]
  int channels, H, W;
  int64_t N;
  const FrameXform* xf;
  const int64_t* dst_slot; // f32 output slot of each frame (collation); nullptr: slot = frame index
  int roi, crop;
  float mean, stdv;
  uint8_t* gray_out;       // stream group output (nullptr: no stream group)
  uint8_t* lip_u8;         // nullable
  float* lip_f32;          // nullable
  unsigned* counter;       // work queue of the producers (zeroed by tform_kernel)
  int64_t ngroups;         // full 512-px groups in the flat pixel stream
  int stage_align;         // cp.async width for footprints (16 or 4); 0 = footprints not staged
  int host_frames;         // `frames` is mapped pinned HOST memory: footprints are pulled across PCIe
};

struct ItemDesc {
  FrameXform x;                       // matrix rows, crop origin, packed footprint
  unsigned item;                      // frame index, or kItemDone
  unsigned pad[3];
};

struct BlendBuf {                     // private to the blend warps, double-buffered per item
  double colx[kMaxRoi], coly[kMaxRoi], rowx[kMaxRoi], rowy[kMaxRoi];   // hoisted M0*c, M3*c, M1*r, M4*r
  uint16_t tile[kTilePx];             // gray footprint, stored as 128*k (byte offset of lut255 row k)
};

struct FusedSmem {
  // k / 255.0 (img_as_float), 16 interleaved copies: entry (k, c) at [k*16 + c].  A lane reads
  // copy (lane & 15), so the 16 lanes of a 64-bit shared-load phase always hit 16 different bank
  // pairs whatever their k: the tap lookups are bank-conflict free.
  double lut255[256 * 16];
  float lutn[256];                    // ((k/255) - mean) / std in float32
  unsigned pad[4];
  ItemDesc desc[kMaxSlots];           // slot s: written by its producer, read by the blend warps
  BlendBuf buf[2];
  // slot s: cp.async landing zone, footprint bytes as in the frame.  raw[0..1] are followed by the
  // stream warps' ring; without stream warps that space is slots 2..5 (raw_slot()).
  uint4 raw[kStreamSlots][kTilePx * 3 / 16];
  uint4 ring[kMaxStreamWarps][kRingStages][kChunkVec];
};
static_assert(sizeof(uint4) * kMaxStreamWarps * kRingStages * kChunkVec >=
              sizeof(uint4) * (kQueueSlots - kStreamSlots) * (kTilePx * 3 / 16), "ring space holds the extra slots");
__device__ __forceinline__ uint4* raw_slot(FusedSmem& sm, int s) { return &sm.raw[0][0] + (size_t)s * (kTilePx * 3 / 16); }

__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_sync_id(int count) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive_id(int count) {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- gray, dp2a form
// 2*Y = 2*(3735 B + 19235 G + 9798 R) + 32768, so that Y = byte 2 of the accumulator: two dp2a
// per pixel (16-bit coefficients against the byte pair where the pixel's bytes sit) and no
// byte extraction.  K2(a, b) packs the doubled coefficients of (lower byte, upper byte).
#define AVFE_K2(a, b) ((uint32_t)(2 * (a)) | ((uint32_t)(2 * (b)) << 16))
__device__ __forceinline__ uint32_t gray_acc(uint32_t w, uint32_t wn, int o) {
  switch (o) {                       // o = byte offset of the pixel inside w (compile-time)
    case 0:  return __dp2a_hi(AVFE_K2(9798, 0), w, __dp2a_lo(AVFE_K2(3735, 19235), w, 32768u));
    case 1:  return __dp2a_hi(AVFE_K2(19235, 9798), w, __dp2a_lo(AVFE_K2(0, 3735), w, 32768u));
    case 2:  return __dp2a_lo(AVFE_K2(9798, 0), wn, __dp2a_hi(AVFE_K2(3735, 19235), w, 32768u));
    default: return __dp2a_lo(AVFE_K2(19235, 9798), wn, __dp2a_hi(AVFE_K2(0, 3735), w, 32768u));
  }
}
// four consecutive pixels from three words -> four accumulators (Y in byte 2 of each)
__device__ __forceinline__ void gray_acc4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&a)[4]) {
  a[0] = gray_acc(w0, w1, 0);
  a[1] = gray_acc(w0, w1, 3);
  a[2] = gray_acc(w1, w2, 2);
  a[3] = gray_acc(w2, 0u, 1);
}
__device__ __forceinline__ uint32_t pack_y4(const uint32_t (&a)[4]) {
  return __byte_perm(__byte_perm(a[0], a[1], 0x0062), __byte_perm(a[2], a[3], 0x0062), 0x5410);
}
__device__ __forceinline__ uint4 gray16_dp2a(const uint4& q0, const uint4& q1, const uint4& q2) {
  uint32_t a[4];
  uint4 r;
  gray_acc4(q0.x, q0.y, q0.z, a); r.x = pack_y4(a);
  gray_acc4(q0.w, q1.x, q1.y, a); r.y = pack_y4(a);
  gray_acc4(q1.z, q1.w, q2.x, a); r.z = pack_y4(a);
  gray_acc4(q2.y, q2.z, q2.w, a); r.w = pack_y4(a);
  return r;
}

// ---------------------------------------------------------------- stream warps
__device__ __forceinline__ void stream_run(const LipJob& j, FusedSmem& sm, int tid, int n_warps) {
  const int lane = tid & 31, wid = tid >> 5;
  const uint4* src = reinterpret_cast<const uint4*>(j.frames);
  uint4* dst = reinterpret_cast<uint4*>(j.gray_out);
  uint4(*ring)[kChunkVec] = sm.ring[wid];
  const int64_t nchunks = j.ngroups / 2;                       // 2 groups (1024 px) per chunk
  const int64_t stride = (int64_t)gridDim.x * n_warps;
  const int64_t first = (int64_t)blockIdx.x * n_warps + wid;
  auto issue = [&](int64_t c, int stage) {
    if (c < nchunks) {
      const uint4* p = src + c * kChunkVec;
#pragma unroll
      for (int i = 0; i < kChunkVec / 32; ++i) cp_async16(&ring[stage][lane + 32 * i], p + lane + 32 * i);
    }
    cp_async_commit_group();
  };
  issue(first, 0);
  issue(first + stride, 1);
  issue(first + 2 * stride, 2);
  int stage = 0;
  for (int64_t c = first; c < nchunks; c += stride) {
    int nxt = stage + 3; if (nxt >= kRingStages) nxt -= kRingStages;
    issue(c + 3 * stride, nxt);                                // keeps three chunks in flight
    cp_async_wait_group<3>();                                  // chunk c has landed (this lane's part)
    __syncwarp();
    const uint4* s = ring[stage];
    // lane owns 48 contiguous bytes (16 px) of each of the chunk's two groups
    const uint4 p0 = s[3 * lane], p1 = s[3 * lane + 1], p2 = s[3 * lane + 2];
    const uint4 q0 = s[96 + 3 * lane], q1 = s[96 + 3 * lane + 1], q2 = s[96 + 3 * lane + 2];
    __syncwarp();                                              // slot may be refilled from now on
    stg_stream(dst + c * 64 + lane, gray16_dp2a(p0, p1, p2));
    stg_stream(dst + c * 64 + 32 + lane, gray16_dp2a(q0, q1, q2));
    if (++stage == kRingStages) stage = 0;
  }
  cp_async_wait_group<0>();
  // pixels past the last full chunk (whole batch, not per frame)
  const int64_t npx = j.N * (int64_t)j.H * j.W;
  const int nthreads = n_warps * 32;
  for (int64_t i = nchunks * 1024 + (int64_t)blockIdx.x * nthreads + tid; i < npx;
       i += (int64_t)gridDim.x * nthreads)
    j.gray_out[i] = (uint8_t)gray_from_bgr(j.frames[3 * i], j.frames[3 * i + 1], j.frames[3 * i + 2]);
}

// ---------------------------------------------------------------- producer warps
// issue the cp.async copies of one footprint into raw (bytes exactly as in the frame)
__device__ __forceinline__ void prefetch_footprint(const LipJob& j, int64_t f, const Footprint& fp,
                                                   uint4* raw, int tid) {
  if (fp.staged) {
    const int C = j.channels;
    const int64_t row_bytes = (int64_t)j.W * C;
    const uint8_t* base = j.frames + (f * (int64_t)j.H * j.W + (int64_t)fp.r0 * j.W + fp.c0) * C;
    const int lane = tid & 31;                           // one producer warp copies one footprint
    if (j.stage_align == 16 && j.host_frames) {
      // Host-resident frames: 128-bit loads through registers, eight per lane in flight (4 KB per
      // warp, 24 KB per SM).  Measured on the benchmark batch (16,441 frames): 7.2 ms against
      // 14.4 ms with the cp.async loop below, which puts a whole footprint's ~36 requests per lane
      // in flight at once; 7.2 ms is the link's rate for 240..288-byte row pieces
      // (profiles/r02/pcie_probe.txt: 46.5 GB/s against 55.6 GB/s for the copy engine).
      const int vpr = fp.pitch * C / 16;
      const int total = fp.rows * vpr;
      const unsigned magic = (0xFFFFFFFFu / (unsigned)vpr) + 1u;
      for (int b0 = 0; b0 < total; b0 += 256) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = b0 + 32 * u + lane;
          if (idx < total) {
            const int r = (int)__umulhi((unsigned)idx, magic), w = idx - r * vpr;
            v[u] = ldg_stream(reinterpret_cast<const uint4*>(base + r * row_bytes + 16 * w));
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = b0 + 32 * u + lane;
          if (idx < total) raw[idx] = v[u];
        }
      }
    } else if (j.stage_align == 16) {
      const int vpr = fp.pitch * C / 16;                 // 16-byte chunks per footprint row
      const int total = fp.rows * vpr;
      const unsigned magic = (0xFFFFFFFFu / (unsigned)vpr) + 1u;       // exact idx / vpr (idx < 2^16)
      for (int idx = lane; idx < total; idx += 32) {
        const int r = (int)__umulhi((unsigned)idx, magic), v = idx - r * vpr;
        cp_async16(raw + idx, base + r * row_bytes + 16 * v);
      }
    } else {
      const int wpr = fp.pitch * C / 4;
      const int total = fp.rows * wpr;
      const unsigned magic = (0xFFFFFFFFu / (unsigned)wpr) + 1u;
      uint32_t* raw32 = reinterpret_cast<uint32_t*>(raw);
      for (int idx = lane; idx < total; idx += 32) {
        const int r = (int)__umulhi((unsigned)idx, magic), w = idx - r * wpr;
        cp_async4(raw32 + idx, base + r * row_bytes + 4 * w);
      }
    }
  }
  cp_async_commit_group();
}

// raw footprint bytes -> tile (8 * gray), four pixels per step
__device__ __forceinline__ void convert_footprint(const LipJob& j, const Footprint& fp, const uint4* raw4,
                                                  uint16_t* tile, int tid, int nthreads) {
  if (!fp.staged) return;
  const uint32_t* raw = reinterpret_cast<const uint32_t*>(raw4);
  const int quads = fp.rows * fp.pitch / 4;
  uint2* t2 = reinterpret_cast<uint2*>(tile);
  if (j.channels == 3) {
    for (int q = tid; q < quads; q += nthreads) {
      uint32_t a[4];
      gray_acc4(raw[3 * q], raw[3 * q + 1], raw[3 * q + 2], a);
      // Y = byte 2 of a[k]; store 128*Y as u16: (a >> 9) & 0x7f80
      t2[q] = make_uint2(((a[0] >> 9) & 0x7f80u) | (((a[1] >> 9) & 0x7f80u) << 16),
                         ((a[2] >> 9) & 0x7f80u) | (((a[3] >> 9) & 0x7f80u) << 16));
    }
  } else {
    for (int q = tid; q < quads; q += nthreads) {
      const uint32_t w = raw[q];
      t2[q] = make_uint2(((w & 0xffu) << 7) | (((w >> 8) & 0xffu) << 23),
                         (((w >> 16) & 0xffu) << 7) | (((w >> 24) & 0xffu) << 23));
    }
  }
}

// Each of the two producer warps owns one slot and runs its own chain, independently of the
// other: draw a frame from the queue, fetch its descriptor, wait until the blend warps have
// released the slot, copy the footprint, publish.  While the blend warps work on one slot the
// other warp's copy is in flight.
template <int SPAN, int NSLOT>
__device__ __forceinline__ void producer_run(const LipJob& j, FusedSmem& sm, int tid) {
  using R = Roles<SPAN, NSLOT>;
  const unsigned total = (unsigned)j.N;
  const int lane = tid & 31, s = tid >> 5;              // slot == producer warp index
  for (unsigned n = 0;; ++n) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(j.counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    FrameXform x;
    if (t < total) x = j.xf[t];
    if (n >= 1) bar_sync(kBarEmpty + s, R::kSlotThreads);   // blend warps are done with this slot's previous item
    if (t >= total) {
      if (lane == 0) sm.desc[s].item = kItemDone;
      __syncwarp();
      bar_arrive(kBarFull + s, R::kSlotThreads);
      break;
    }
    prefetch_footprint(j, (int64_t)t, unpack_footprint(x), raw_slot(sm, s), lane);
    if (lane == 0) { sm.desc[s].x = x; sm.desc[s].item = t; }
    cp_async_wait_group<0>();
    __syncwarp();
    bar_arrive(kBarFull + s, R::kSlotThreads);          // slot FULL
  }
}

// ---------------------------------------------------------------- blend warps
__device__ __forceinline__ double lut_at(const double* lut, uint32_t off8) {   // off8 = 128 * k
  return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(lut) + off8);
}
// tile element -> byte offset of its lut255 row: the generic kernel stores 128 * k as u16, the frame
// kernel the gray level itself as u8 (twice the pixels in the same shared memory, one shift per tap)
__device__ __forceinline__ uint32_t tile_off(uint16_t v) { return v; }
__device__ __forceinline__ uint32_t tile_off(uint8_t v) { return (uint32_t)v << 7; }

// One output pixel whose four taps are known to lie inside the staged tile (interior ROI):
// no bounds tests, tap offsets by increment.  Same roundings as bilinear_u8.
template <typename T>
__device__ __forceinline__ uint32_t bilinear_interior(double r, double c, const T* tile, int pitch,
                                                      int br0, int bc0, const double* lut) {
  const double fr = floor(r), fc = floor(c);
  const double dr = f64sub(r, fr), dc = f64sub(c, fc);
  const T* p = tile + ((int)fr - br0) * pitch + ((int)fc - bc0);
  const int oc = (dc != 0.0) ? 1 : 0;                 // ceil(c) - floor(c)
  const int orow = (dr != 0.0) ? pitch : 0;           // (ceil(r) - floor(r)) * pitch
  const double tl = lut_at(lut, tile_off(p[0])), tr = lut_at(lut, tile_off(p[oc]));
  const double bl = lut_at(lut, tile_off(p[orow])), br = lut_at(lut, tile_off(p[orow + oc]));
  const double omc = f64sub(1.0, dc), omr = f64sub(1.0, dr);
  const double top = f64add(f64mul(omc, tl), f64mul(dc, tr));
  const double bot = f64add(f64mul(omc, bl), f64mul(dc, br));
  const double v = f64add(f64mul(omr, top), f64mul(dr, bot));
  return (uint32_t)(int)f64mul(v, 255.0);
}

template <int SPAN, int NSLOT>
__device__ __forceinline__ void blend_item(const LipJob& j, int64_t f, const FrameXform& x,
                                           const Footprint& fp, const BlendBuf& slot,
                                           const FusedSmem& sm, int tid) {
  using R = Roles<SPAN, NSLOT>;
  const double* lut = sm.lut255 + (tid & 15);           // this lane's copy of the k/255 table
  const int off = (j.roi - j.crop) / 2;
  const int lo = j.lip_u8 ? 0 : off;
  const int span = SPAN ? SPAN : (j.lip_u8 ? j.roi : j.crop);
  const int H = j.H, W = j.W;
  uint8_t* out_u8 = j.lip_u8 ? j.lip_u8 + f * (int64_t)j.roi * j.roi : nullptr;
  const int64_t slot_f32 = j.dst_slot ? j.dst_slot[f] : f;          // < 0: dropped by the collation trim
  float* out_f32 = (j.lip_f32 && slot_f32 >= 0) ? j.lip_f32 + slot_f32 * (int64_t)j.crop * j.crop : nullptr;
  if (out_u8 == nullptr && out_f32 == nullptr) return;
  const int npix = span * span;
  auto emit = [&](int r, int c, uint32_t v) {
    if (SPAN == 88) {                                   // centre crop only: window == f32 output
      out_f32[r * 88 + c] = sm.lutn[v];
      return;
    }
    const int pr = lo + r, pc = lo + c;
    if (out_u8) out_u8[pr * j.roi + pc] = (uint8_t)v;
    if (out_f32) {
      const int cr = pr - off, cc = pc - off;
      if ((unsigned)cr < (unsigned)j.crop && (unsigned)cc < (unsigned)j.crop)
        out_f32[cr * j.crop + cc] = sm.lutn[v];
    }
  };
  if (x.r0 < 0) {                                       // clip without any detection: zero ROI
    for (int idx = tid; idx < npix; idx += R::kBlendThreads) emit(idx / span, idx % span, 0u);
    return;
  }
  const double m2 = x.inv[2], m5 = x.inv[5];
  const int br0 = fp.r0, bc0 = fp.c0, pitch = fp.pitch;
  if (fp.interior && SPAN != 0) {
    // common case.  Thread = (column c, row phase): the column products stay in registers, the
    // row products are warp-wide broadcasts, rows advance by 8 with no index arithmetic, and
    // the fully unrolled rows give the scheduler independent chains to interleave.
    constexpr int S = R::kSide;
    const int rr = tid / S, c = tid - rr * S;           // every blend thread owns one column phase
    const double cx = slot.colx[c], cy = slot.coly[c];
#pragma unroll
    for (int i = 0; i < S / 8; ++i) {
      const int r = rr + 8 * i;
      const double sc = f64add(f64add(cx, slot.rowx[r]), m2);
      const double sr = f64add(f64add(cy, slot.rowy[r]), m5);
      emit(r, c, bilinear_interior(sr, sc, slot.tile, pitch, br0, bc0, lut));
    }
    return;
  }
  const bool bgr = (j.channels == 3);
  const uint8_t* img = j.frames + f * (int64_t)H * W * (bgr ? 3 : 1);
  const int brows = fp.staged ? fp.rows : 0;
  auto tap = [&](int r, int c) -> double {
    const int rr = r - br0, cc = c - bc0;
    if ((unsigned)rr < (unsigned)brows && (unsigned)cc < (unsigned)pitch)
      return lut_at(lut, slot.tile[rr * pitch + cc]);
    const uint8_t* p = img + ((int64_t)r * W + c) * (bgr ? 3 : 1);       // not staged: global tap
    return lut[16 * (bgr ? gray_from_bgr(__ldg(p), __ldg(p + 1), __ldg(p + 2)) : (uint32_t)__ldg(p))];
  };
  for (int idx = tid; idx < npix; idx += R::kBlendThreads) {
    const int r = idx / span, c = idx - r * span;
    const double sc = f64add(f64add(slot.colx[c], slot.rowx[r]), m2);
    const double sr = f64add(f64add(slot.coly[c], slot.rowy[r]), m5);
    const uint32_t v = fp.interior ? bilinear_interior(sr, sc, slot.tile, pitch, br0, bc0, lut)
                                   : (uint32_t)bilinear_u8(sr, sc, H, W, tap);
    emit(r, c, v);
  }
}

template <int SPAN, int NSLOT>
__device__ __forceinline__ void blend_run(const LipJob& j, FusedSmem& sm, int tid) {
  using R = Roles<SPAN, NSLOT>;
  for (int k = tid; k < 256; k += R::kBlendThreads) {
    const double q = f64div((double)k, 255.0);
#pragma unroll
    for (int c = 0; c < 16; ++c) sm.lut255[k * 16 + c] = q;
    sm.lutn[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)k, 255.0f), j.mean), j.stdv);
  }
  const int off = (j.roi - j.crop) / 2;
  const int lo = j.lip_u8 ? 0 : off;
  const int span = SPAN ? SPAN : (j.lip_u8 ? j.roi : j.crop);
  unsigned alive = (1u << NSLOT) - 1u;                  // a slot's producer publishes kItemDone once
  unsigned nb = 0;                                      // items blended so far: picks the BlendBuf
  for (int s = 0; alive != 0u; s = (s + 1 == NSLOT) ? 0 : s + 1) {
    if (!((alive >> s) & 1u)) continue;
    bar_sync(kBarFull + s, R::kSlotThreads);            // slot FULL: descriptor + raw footprint landed
    const unsigned item = sm.desc[s].item;
    if (item == kItemDone) {
      alive &= ~(1u << s);
      continue;
    }
    const FrameXform x = sm.desc[s].x;
    const Footprint fp = unpack_footprint(x);
    BlendBuf& buf = sm.buf[nb & 1u];
    ++nb;
    convert_footprint(j, fp, raw_slot(sm, s), buf.tile, tid, R::kBlendThreads);
    if (x.r0 >= 0) {
      // hoisted products: x_ = (M0*c + M1*r) + M2 is evaluated as (colx[c] + rowx[r]) + M2 with
      // the same two roundings per product as skimage's _transform_affine
      if (tid < span) {
        const double tc = (double)(x.c0 + lo + tid);
        buf.colx[tid] = f64mul(x.inv[0], tc);
        buf.coly[tid] = f64mul(x.inv[3], tc);
      } else if (tid >= 128 && tid < 128 + span) {
        const double tr = (double)(x.r0 + lo + tid - 128);
        buf.rowx[tid - 128] = f64mul(x.inv[1], tr);
        buf.rowy[tid - 128] = f64mul(x.inv[4], tr);
      }
    }
    bar_sync(kBarBlend, R::kBlendThreads);              // tile + tables ready; raw/desc no longer read
    bar_arrive(kBarEmpty + s, R::kSlotThreads);         // slot EMPTY: its producer may refill it
    blend_item<SPAN, NSLOT>(j, (int64_t)item, x, fp, buf, sm, tid);
  }
}

// STREAM: blend + 2 producer + stream warps (1024 threads); otherwise blend + 6 producer warps.
template <bool STREAM, int SPAN>
__global__ void __launch_bounds__(STREAM ? 1024 : Roles<SPAN, kQueueSlots>::kHandoverThreads, 1)
lip_fused_kernel(const LipJob j) {
  constexpr int NSLOT = STREAM ? kStreamSlots : kQueueSlots;
  using R = Roles<SPAN, NSLOT>;
  extern __shared__ __align__(16) unsigned char fused_smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fused_smem_raw);
  const int tid = threadIdx.x;
  if (tid < R::kBlendThreads) {
    blend_run<SPAN, NSLOT>(j, sm, tid);
  } else if (tid < R::kHandoverThreads) {
    producer_run<SPAN, NSLOT>(j, sm, tid - R::kBlendThreads);
  } else if (STREAM) {
    stream_run(j, sm, tid - R::kHandoverThreads, R::kStreamWarps);
  }
}

}  // namespace avfe
