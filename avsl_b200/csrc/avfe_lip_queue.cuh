// lip_fused_kernel: the gray conversion and the ROI warp of a batch in ONE persistent launch,
// warp-specialised so that both kinds of work are resident on every SM all the time:
//
//   stream warps   (8 per CTA) BGR->gray over the flat pixel stream, statically strided, six
//                  128-bit loads in flight per lane; HBM-bound, integer dp2a
//   compute warps  (8 per CTA) one frame's ROI per work-queue item: the source footprint of the
//                  NEXT item is fetched with cp.async while the current one is blended (gray
//                  computed from the BGR footprint, so there is no dependency on the stream
//                  warps), float64 bilinear blend in skimage's operation order, u8 ROI +
//                  normalised f32 centre crop; FP64-pipe / issue-bound
//
// The two groups never synchronise with each other (compute warps use named barrier 2); the
// FP64 work hides under the memory time of the stream warps.  Included by avfe_lip.cu only.
#pragma once

namespace avfe {

constexpr int kGroupThreads = 256;
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kTilePx = 9216;               // staged footprint capacity (e.g. 96 x 96 source px)
constexpr int kMaxRoi = 128;

struct LipJob {
  const uint8_t* frames;   // [N,H,W,channels]
  int channels, H, W;
  int64_t N;
  const FrameXform* xf;
  int roi, crop;
  float mean, stdv;
  uint8_t* gray_out;       // stream group output (nullptr: no stream group)
  uint8_t* lip_u8;         // nullable
  float* lip_f32;          // nullable
  unsigned* counter;       // work queue of the compute group (zeroed by tform_kernel)
  int64_t ngroups;         // full 512-px groups in the flat pixel stream
};

struct FusedSmem {
  double lut255[256];                 // k / 255.0 (img_as_float)
  double colx[kMaxRoi], coly[kMaxRoi], rowx[kMaxRoi], rowy[kMaxRoi];
  float lutn[256];                    // ((k/255) - mean) / std in float32
  unsigned item[2];
  unsigned pad[2];
  uint32_t raw[2][kTilePx * 3 / 4];   // cp.async landing zone: footprint bytes as fetched
  uint16_t tile[kTilePx];             // gray footprint, stored as 8*k (byte offset into lut255)
  uint4 slab[kGroupWarps][2][96];     // stream group: two 512-px groups per warp
};

__device__ __forceinline__ void group_barrier(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kGroupThreads) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- gray, dp2a form
// 2*Y = 2*(3735 B + 19235 G + 9798 R) + 32768, so that Y = byte 2 of the accumulator: two dp2a
// per pixel (16-bit coefficients against the byte pair where the pixel's bytes sit) and no
// byte extraction.  K(a, b) packs the doubled coefficients of (lower byte, upper byte).
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

// ---------------------------------------------------------------- stream group
__device__ __forceinline__ void stream_group_run(const LipJob& j, FusedSmem& sm, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  const uint4* src = reinterpret_cast<const uint4*>(j.frames);
  uint4* dst = reinterpret_cast<uint4*>(j.gray_out);
  uint4* s0 = sm.slab[wid][0];
  uint4* s1 = sm.slab[wid][1];
  const int64_t stride = (int64_t)gridDim.x * kGroupWarps;
  int64_t g = (int64_t)blockIdx.x * kGroupWarps + wid;
  // two groups (1024 px, 3 KB) per iteration: six independent 128-bit loads per lane, staged
  // through the warp's slab so that every lane then owns 48 contiguous bytes (16 px)
  for (; g + stride < j.ngroups; g += 2 * stride) {
    const uint4* a = src + g * 96;
    const uint4* b = src + (g + stride) * 96;
    const uint4 a0 = ldg_stream(a + lane), a1 = ldg_stream(a + lane + 32), a2 = ldg_stream(a + lane + 64);
    const uint4 b0 = ldg_stream(b + lane), b1 = ldg_stream(b + lane + 32), b2 = ldg_stream(b + lane + 64);
    s0[lane] = a0; s0[lane + 32] = a1; s0[lane + 64] = a2;
    s1[lane] = b0; s1[lane + 32] = b1; s1[lane + 64] = b2;
    __syncwarp();
    const uint4 p0 = s0[3 * lane], p1 = s0[3 * lane + 1], p2 = s0[3 * lane + 2];
    const uint4 q0 = s1[3 * lane], q1 = s1[3 * lane + 1], q2 = s1[3 * lane + 2];
    __syncwarp();
    stg_stream(dst + g * 32 + lane, gray16_dp2a(p0, p1, p2));
    stg_stream(dst + (g + stride) * 32 + lane, gray16_dp2a(q0, q1, q2));
  }
  if (g < j.ngroups) {
    const uint4* a = src + g * 96;
    const uint4 a0 = ldg_stream(a + lane), a1 = ldg_stream(a + lane + 32), a2 = ldg_stream(a + lane + 64);
    s0[lane] = a0; s0[lane + 32] = a1; s0[lane + 64] = a2;
    __syncwarp();
    const uint4 p0 = s0[3 * lane], p1 = s0[3 * lane + 1], p2 = s0[3 * lane + 2];
    stg_stream(dst + g * 32 + lane, gray16_dp2a(p0, p1, p2));
  }
  // pixels past the last full group (whole batch, not per frame)
  const int64_t npx = j.N * (int64_t)j.H * j.W;
  for (int64_t i = j.ngroups * 512 + (int64_t)blockIdx.x * kGroupThreads + tid; i < npx;
       i += (int64_t)gridDim.x * kGroupThreads)
    j.gray_out[i] = (uint8_t)gray_from_bgr(j.frames[3 * i], j.frames[3 * i + 1], j.frames[3 * i + 2]);
}

// ---------------------------------------------------------------- compute group
// Source footprint of one ROI window, frame-clipped and 4-pixel aligned.
struct Footprint {
  int r0, c0, rows, pitch;   // staged box (pitch % 4 == 0); rows == 0: nothing staged
  bool interior;             // every tap of every output pixel lies inside the frame
  bool staged;               // the box is in shared memory (else taps come from global memory)
};

__device__ __forceinline__ Footprint footprint_of(const FrameXform& x, int lo, int span, int H, int W,
                                                  bool can_stage) {
  Footprint fp{0, 0, 0, 0, false, false};
  if (x.r0 < 0) return fp;
  // an affine map takes its extrema at the window corners
  double rmin = 1e300, rmax = -1e300, cmin = 1e300, cmax = -1e300;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double tr = (double)(x.r0 + lo + ((k & 1) ? span - 1 : 0));
    const double tc = (double)(x.c0 + lo + ((k & 2) ? span - 1 : 0));
    const double sc = x.inv[0] * tc + x.inv[1] * tr + x.inv[2];
    const double sr = x.inv[3] * tc + x.inv[4] * tr + x.inv[5];
    rmin = fmin(rmin, sr); rmax = fmax(rmax, sr);
    cmin = fmin(cmin, sc); cmax = fmax(cmax, sc);
  }
  const bool finite = (rmin == rmin) && (cmin == cmin) && fabs(rmin) < 1e9 && fabs(rmax) < 1e9 &&
                      fabs(cmin) < 1e9 && fabs(cmax) < 1e9;
  if (!finite) return fp;
  // one pixel of slack each side (covers the rounding of the hoisted evaluation)
  const double fr0 = floor(rmin) - 1.0, fr1 = ceil(rmax) + 1.0;
  const double fc0 = floor(cmin) - 1.0, fc1 = ceil(cmax) + 1.0;
  fp.interior = fr0 >= 0.0 && fc0 >= 0.0 && fr1 <= (double)(H - 1) && fc1 <= (double)(W - 1);
  const int r0 = (int)fmax(fr0, 0.0), c0 = ((int)fmax(fc0, 0.0)) & ~3;
  const int rows = (int)fmin(fr1, (double)(H - 1)) - r0 + 1;
  const int cols = (int)fmin(fc1, (double)(W - 1)) - c0 + 1;
  if (rows <= 0 || cols <= 0) { fp.interior = false; return fp; }   // window entirely off-frame
  const int pitch = min((cols + 3) & ~3, W - c0);                    // W % 4 == 0 when can_stage
  fp.r0 = r0; fp.c0 = c0; fp.rows = rows; fp.pitch = pitch;
  fp.staged = can_stage && (rows * pitch <= kTilePx);
  if (!fp.staged) fp.interior = false;
  return fp;
}

// issue the cp.async copies of one footprint into raw[buf] (bytes exactly as in the frame)
__device__ __forceinline__ void prefetch_footprint(const LipJob& j, int64_t f, const Footprint& fp,
                                                   uint32_t* raw, int tid) {
  if (fp.staged) {
    const int C = j.channels;
    const int wpr = fp.pitch * C / 4;                    // words per footprint row
    const int total = fp.rows * wpr;
    const unsigned magic = (0xFFFFFFFFu / (unsigned)wpr) + 1u;   // exact idx / wpr (idx < 2^16)
    const uint8_t* base = j.frames + (f * (int64_t)j.H * j.W + (int64_t)fp.r0 * j.W + fp.c0) * C;
    const int64_t row_bytes = (int64_t)j.W * C;
    for (int idx = tid; idx < total; idx += kGroupThreads) {
      const int r = (int)__umulhi((unsigned)idx, magic), w = idx - r * wpr;
      cp_async4(raw + idx, base + r * row_bytes + 4 * w);
    }
  }
  cp_async_commit_group();
}

// raw footprint bytes -> tile (8 * gray), four pixels per step
__device__ __forceinline__ void convert_footprint(const LipJob& j, const Footprint& fp, const uint32_t* raw,
                                                  uint16_t* tile, int tid) {
  if (!fp.staged) return;
  const int quads = fp.rows * fp.pitch / 4;
  uint2* t2 = reinterpret_cast<uint2*>(tile);
  if (j.channels == 3) {
    for (int q = tid; q < quads; q += kGroupThreads) {
      uint32_t a[4];
      gray_acc4(raw[3 * q], raw[3 * q + 1], raw[3 * q + 2], a);
      // Y = byte 2 of a[k]; store 8*Y as u16: (a >> 13) & 0x7f8
      t2[q] = make_uint2(((a[0] >> 13) & 0x7f8u) | (((a[1] >> 13) & 0x7f8u) << 16),
                         ((a[2] >> 13) & 0x7f8u) | (((a[3] >> 13) & 0x7f8u) << 16));
    }
  } else {
    for (int q = tid; q < quads; q += kGroupThreads) {
      const uint32_t w = raw[q];
      t2[q] = make_uint2(((w & 0xffu) << 3) | (((w >> 8) & 0xffu) << 19),
                         (((w >> 16) & 0xffu) << 3) | (((w >> 24) & 0xffu) << 19));
    }
  }
}

__device__ __forceinline__ double lut_at(const double* lut, uint32_t off8) {
  return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(lut) + off8);
}

// One output pixel whose four taps are known to lie inside the staged tile (interior ROI):
// no bounds tests, tap offsets by increment.  Same roundings as bilinear_u8.
__device__ __forceinline__ uint32_t bilinear_interior(double r, double c, const uint16_t* tile, int pitch,
                                                      int br0, int bc0, const double* lut) {
  const double fr = floor(r), fc = floor(c);
  const double dr = f64sub(r, fr), dc = f64sub(c, fc);
  const uint16_t* p = tile + ((int)fr - br0) * pitch + ((int)fc - bc0);
  const int oc = (dc != 0.0) ? 1 : 0;                 // ceil(c) - floor(c)
  const int orow = (dr != 0.0) ? pitch : 0;           // (ceil(r) - floor(r)) * pitch
  const double tl = lut_at(lut, p[0]), tr = lut_at(lut, p[oc]);
  const double bl = lut_at(lut, p[orow]), br = lut_at(lut, p[orow + oc]);
  const double omc = f64sub(1.0, dc), omr = f64sub(1.0, dr);
  const double top = f64add(f64mul(omc, tl), f64mul(dc, tr));
  const double bot = f64add(f64mul(omc, bl), f64mul(dc, br));
  const double v = f64add(f64mul(omr, top), f64mul(dr, bot));
  return (uint32_t)(int)f64mul(v, 255.0);
}

// SPAN = side of the evaluated window (96 when the u8 ROI is wanted, 88 for the centre crop
// only, 0 = run-time value).  The footprint has already been converted into sm.tile.
template <int SPAN>
__device__ __forceinline__ void blend_item(const LipJob& j, int64_t f, const FrameXform& x,
                                           const Footprint& fp, FusedSmem& sm, int tid) {
  const int off = (j.roi - j.crop) / 2;
  const int lo = j.lip_u8 ? 0 : off;
  const int span = SPAN ? SPAN : (j.lip_u8 ? j.roi : j.crop);
  const int H = j.H, W = j.W;
  uint8_t* out_u8 = j.lip_u8 ? j.lip_u8 + f * (int64_t)j.roi * j.roi : nullptr;
  float* out_f32 = j.lip_f32 ? j.lip_f32 + f * (int64_t)j.crop * j.crop : nullptr;
  const int npix = span * span;
  if (x.r0 < 0) {                                      // clip without any detection: zero ROI
    for (int idx = tid; idx < npix; idx += kGroupThreads) {
      const int pr = lo + idx / span, pc = lo + idx % span;
      if (out_u8) out_u8[pr * j.roi + pc] = 0;
      const int cr = pr - off, cc = pc - off;
      if (out_f32 && cr >= 0 && cr < j.crop && cc >= 0 && cc < j.crop) out_f32[cr * j.crop + cc] = sm.lutn[0];
    }
    return;
  }
  const bool bgr = (j.channels == 3);
  const uint8_t* img = j.frames + f * (int64_t)H * W * (bgr ? 3 : 1);
  const int br0 = fp.r0, bc0 = fp.c0, brows = fp.staged ? fp.rows : 0, pitch = fp.pitch;
  auto tap = [&](int r, int c) -> double {
    const int rr = r - br0, cc = c - bc0;
    if ((unsigned)rr < (unsigned)brows && (unsigned)cc < (unsigned)pitch)
      return lut_at(sm.lut255, sm.tile[rr * pitch + cc]);
    const uint8_t* p = img + ((int64_t)r * W + c) * (bgr ? 3 : 1);       // not staged: global tap
    return sm.lut255[bgr ? gray_from_bgr(__ldg(p), __ldg(p + 1), __ldg(p + 2)) : (uint32_t)__ldg(p)];
  };
  const double m2 = x.inv[2], m5 = x.inv[5];
#pragma unroll 2
  for (int idx = tid; idx < npix; idx += kGroupThreads) {
    const int r = idx / span, c = idx - r * span;        // constant divisor when SPAN != 0
    const double sc = f64add(f64add(sm.colx[c], sm.rowx[r]), m2);
    const double sr = f64add(f64add(sm.coly[c], sm.rowy[r]), m5);
    const uint32_t v = fp.interior ? bilinear_interior(sr, sc, sm.tile, pitch, br0, bc0, sm.lut255)
                                   : (uint32_t)bilinear_u8(sr, sc, H, W, tap);
    const int pr = lo + r, pc = lo + c;
    if (out_u8) out_u8[pr * j.roi + pc] = (uint8_t)v;
    if (out_f32) {
      const int cr = pr - off, cc = pc - off;
      if ((unsigned)cr < (unsigned)j.crop && (unsigned)cc < (unsigned)j.crop)
        out_f32[cr * j.crop + cc] = sm.lutn[v];
    }
  }
}

__device__ __forceinline__ void compute_group_run(const LipJob& j, FusedSmem& sm, int tid) {
  const unsigned total = (unsigned)j.N;
  const int off = (j.roi - j.crop) / 2;
  const int lo = j.lip_u8 ? 0 : off;
  const int span = j.lip_u8 ? j.roi : j.crop;
  // 4-byte cp.async staging needs 4-aligned footprint rows
  const bool can_stage = (j.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(j.frames) & 3u) == 0);
  if (tid == 0) { sm.item[0] = atomicAdd(j.counter, 1u); sm.item[1] = atomicAdd(j.counter, 1u); }
  group_barrier(2);
  unsigned t = sm.item[0], tn = sm.item[1];
  group_barrier(2);
  FrameXform x;
  Footprint fp{0, 0, 0, 0, false, false};
  int cur = 0;
  if (t < total) {
    x = j.xf[t];
    fp = footprint_of(x, lo, span, j.H, j.W, can_stage);     // every thread computes the same box
    prefetch_footprint(j, (int64_t)t, fp, sm.raw[0], tid);
  }
  while (t < total) {
    if (tid == 0) sm.item[cur] = atomicAdd(j.counter, 1u);   // item after next, read at the loop end
    cp_async_wait_all();
    group_barrier(2);                                        // raw[cur] complete for all threads
    convert_footprint(j, fp, sm.raw[cur], sm.tile, tid);
    if (x.r0 >= 0) {
      // hoisted products: x_ = (M0*c + M1*r) + M2 is evaluated as (colx[c] + rowx[r]) + M2 with
      // the same two roundings per product as skimage's _transform_affine
      if (tid < span) {
        const double tc = (double)(x.c0 + lo + tid);
        sm.colx[tid] = f64mul(x.inv[0], tc);
        sm.coly[tid] = f64mul(x.inv[3], tc);
      } else if (tid >= 128 && tid < 128 + span) {
        const double tr = (double)(x.r0 + lo + tid - 128);
        sm.rowx[tid - 128] = f64mul(x.inv[1], tr);
        sm.rowy[tid - 128] = f64mul(x.inv[4], tr);
      }
    }
    group_barrier(2);                                        // tile + tables ready
    if (tn < total) {                                        // next footprint streams in meanwhile
      const FrameXform xn = j.xf[tn];
      const Footprint fpn = footprint_of(xn, lo, span, j.H, j.W, can_stage);
      prefetch_footprint(j, (int64_t)tn, fpn, sm.raw[cur ^ 1], tid);
    }
    if (j.lip_u8 != nullptr && j.roi == 96) blend_item<96>(j, (int64_t)t, x, fp, sm, tid);
    else if (j.lip_u8 == nullptr && j.crop == 88) blend_item<88>(j, (int64_t)t, x, fp, sm, tid);
    else blend_item<0>(j, (int64_t)t, x, fp, sm, tid);
    group_barrier(2);              // item done: tile and tables are free, sm.item[cur] is visible
    const unsigned tnn = sm.item[cur];
    t = tn; tn = tnn;
    cur ^= 1;
    if (t < total) {               // recomputed rather than kept live across the blend (registers)
      x = j.xf[t];
      fp = footprint_of(x, lo, span, j.H, j.W, can_stage);
    }
  }
  cp_async_wait_all();
}

// STREAM: CTA = 8 compute warps + 8 stream warps; otherwise 8 compute warps only.
template <bool STREAM>
__global__ void __launch_bounds__(STREAM ? 2 * kGroupThreads : kGroupThreads, 2)
lip_fused_kernel(const LipJob j) {
  extern __shared__ __align__(16) unsigned char fused_smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fused_smem_raw);
  const int tid = threadIdx.x;
  if (tid < kGroupThreads) {
    for (int k = tid; k < 256; k += kGroupThreads) {
      sm.lut255[k] = f64div((double)k, 255.0);
      sm.lutn[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)k, 255.0f), j.mean), j.stdv);
    }
    compute_group_run(j, sm, tid);     // starts with a group barrier: LUTs are visible
  } else if (STREAM) {
    stream_group_run(j, sm, tid - kGroupThreads);
  }
}

}  // namespace avfe
