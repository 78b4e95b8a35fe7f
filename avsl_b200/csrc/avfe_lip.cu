// Lip-ROI path: BGR->gray, landmark interpolation, window-smoothed similarity fit, bilinear
// warp restricted to the cut_patch window, centre crop and normalisation.
// Reference: preprocess/video_process.py:201-214,369-475; utils/lips_cropping.py:41-163;
// utils/hf_video_utils.py:113-138.
//
// Kernels
//   lip_frame_kernel  (avfe_lip_frame.cuh) the standard case in ONE launch: persistent CTAs own
//                     whole frames; tform warps fit the transforms (tform_frame: V2 on the fly, V3,
//                     V4 fit, V6, V7), stream warps pull the BGR frame through shared memory with
//                     TMA bulk copies, write the gray frame and deposit the ROI's source footprint,
//                     blend warps do the float64 bilinear blend, crop and normalisation
//   tform_kernel +    the generic path (gray input, rows not 16-byte aligned, other ROI sizes, no
//   lip_fused_kernel  gray output): one warp per frame for the transforms, then a warp-specialised
//                     (avfe_lip_queue.cuh) work-queue kernel that stages each footprint itself
//   gray_vec_kernel   the gray conversion alone (avfe_bgr2gray_u8)
//   small single-purpose kernels for the per-function entry points (warp_full, cut_patch, ...)
#include <stddef.h>

#include "avfe_common.cuh"
#include "avfe_lip_math.cuh"

namespace avfe {

// ------------------------------------------------------------------ landmarks with on-the-fly fill
struct LmView {
  const double* lm;       // [N,68,2]
  const uint8_t* valid;   // [N] or nullptr
  int64_t beg, end;       // the clip's frame range
};

// previous / next valid frame of the clip (p < beg or q >= end when there is none)
__device__ __forceinline__ void lm_neighbours(const LmView& v, int64_t f, int64_t& p, int64_t& q) {
  p = f; q = f;
  if (v.valid == nullptr) return;
  while (p >= v.beg && !v.valid[p]) --p;
  while (q < v.end && !v.valid[q]) ++q;
}

// element e (0..135) of frame f after landmarks_interpolate (utils/lips_cropping.py:41-89)
__device__ __forceinline__ double lm_value(const LmView& v, int64_t f, int64_t p, int64_t q, int e) {
  if (p == f) return v.lm[f * 136 + e];
  const bool has_p = p >= v.beg, has_q = q < v.end;
  if (!has_p && !has_q) return nan("");          // no detection in the whole clip
  if (!has_p) return v.lm[q * 136 + e];          // leading frames replicate the first detection
  if (!has_q) return v.lm[p * 136 + e];          // trailing frames replicate the last detection
  // start + idx/float(stop-start) * delta   (utils/lips_cropping.py:54-57)
  const double s = v.lm[p * 136 + e], t = v.lm[q * 136 + e];
  const double w = f64div((double)(f - p), (double)(q - p));
  return f64add(s, f64mul(w, f64sub(t, s)));
}

__device__ __forceinline__ int64_t find_clip(const int64_t* __restrict__ clip_offsets, int64_t n_clips,
                                             int64_t f) {
  int64_t lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (clip_offsets[mid] <= f) lo = mid; else hi = mid;
  }
  return lo;
}

// lm_value in two phases so that a warp can put all its loads in flight before any arithmetic:
// lm_taps issues the (at most two) loads, lm_blend applies landmarks_interpolate's rule.
struct LmTaps { double s, t; };
__device__ __forceinline__ LmTaps lm_taps(const LmView& v, int64_t f, int64_t p, int64_t q, int e) {
  const bool has_p = p >= v.beg, has_q = q < v.end;
  int64_t a = f, b = f;                               // p == f (a detection) or no detection at all
  if (p != f && (has_p || has_q)) {
    a = has_p ? p : q;                                // one-sided: replicate the nearest detection
    b = has_q ? q : p;
  }
  LmTaps r;
  r.s = v.lm[a * 136 + e];
  r.t = (b != a) ? v.lm[b * 136 + e] : r.s;
  return r;
}
__device__ __forceinline__ double lm_blend(const LmView& v, int64_t f, int64_t p, int64_t q, const LmTaps& x) {
  if (p == f) return x.s;
  const bool has_p = p >= v.beg, has_q = q < v.end;
  if (!has_p && !has_q) return nan("");               // no detection in the whole clip
  if (!has_p || !has_q) return x.s;
  // start + idx/float(stop-start) * delta   (utils/lips_cropping.py:54-57)
  const double w = f64div((double)(f - p), (double)(q - p));
  return f64add(x.s, f64mul(w, f64sub(x.t, x.s)));
}

// find_clip by a whole warp: binary search down to <= 1024 candidates, then every lane counts
// the boundaries <= f among its share (independent loads) and the counts are summed
__device__ __forceinline__ int64_t find_clip_warp(const int64_t* __restrict__ clip_offsets, int64_t n_clips,
                                                  int64_t f, int lane) {
  int64_t lo = 0, hi = n_clips;
  while (hi - lo > 1024) {
    const int64_t mid = (lo + hi) >> 1;
    if (clip_offsets[mid] <= f) lo = mid; else hi = mid;
  }
  int cnt = 0;
  for (int64_t i = lo + 1 + lane; i < hi; i += 32) cnt += (clip_offsets[i] <= f) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  return lo + cnt;
}

// V2 alone (avfe_landmarks_interpolate): one CTA per frame
__global__ void __launch_bounds__(160)
lm_fill_kernel(const double* __restrict__ lm, const uint8_t* __restrict__ valid,
               const int64_t* __restrict__ clip_offsets, int64_t n_clips, int64_t N,
               double* __restrict__ out) {
  const int64_t f = blockIdx.x;
  if (f >= N) return;
  const int64_t c = find_clip(clip_offsets, n_clips, f);
  LmView v{lm, valid, clip_offsets[c], clip_offsets[c + 1]};
  int64_t p, q;
  lm_neighbours(v, f, p, q);
  const int t = threadIdx.x;
  if (t < kNumLandmarks * 2) out[f * 136 + t] = lm_value(v, f, p, q, t);
}

// ------------------------------------------------------------------ V3+V4(fit)+V6+V7
// per-frame record written to the workspace by tform_kernel and consumed by the compute warps
struct FrameXform {
  double inv[6];     // rows 0,1 of tform.inverse.params
  int32_t r0, c0;    // cut_patch origin in the std frame (or -1,-1)
  uint32_t box_lo;   // source footprint: r0 | c0 << 13 | interior << 26 | staged << 27
  uint32_t box_hi;   //                   rows | pitch << 13
};

constexpr int kTilePx = 8192;   // staged footprint capacity of the generic kernel (e.g. 80 x 96 source pixels + alignment)
constexpr int kMaxRoi = 128;

// Source footprint of one ROI window, frame-clipped and aligned for cp.async staging.
struct Footprint {
  int r0, c0, rows, pitch;   // staged box; rows == 0: nothing staged
  bool interior;             // every tap of every output pixel lies inside the frame and the box
  bool staged;               // the box goes through shared memory (else taps come from global)
};

__device__ __forceinline__ Footprint unpack_footprint(const FrameXform& x) {
  Footprint fp;
  fp.r0 = (int)(x.box_lo & 0x1fffu);
  fp.c0 = (int)((x.box_lo >> 13) & 0x1fffu);
  fp.interior = ((x.box_lo >> 26) & 1u) != 0;
  fp.staged = ((x.box_lo >> 27) & 1u) != 0;
  fp.rows = (int)(x.box_hi & 0x1fffu);
  fp.pitch = (int)((x.box_hi >> 13) & 0x1fffu);
  return fp;
}

// align = pixel alignment of the box's first column and pitch (16 or 4; 0 = do not stage);
// cap = pixels the consumer's shared-memory tile holds
__device__ __forceinline__ void pack_footprint(FrameXform& x, int lo, int span, int H, int W, int align, int cap) {
  x.box_lo = 0u; x.box_hi = 0u;
  if (x.r0 < 0 || align == 0 || H > 8191 || W > 8191) return;
  // an affine map takes its extrema at the window corners
  double rmin = 1e300, rmax = -1e300, cmin = 1e300, cmax = -1e300;
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
  if (!finite) return;
  // one pixel of slack each side (covers the rounding of the hoisted evaluation)
  const double fr0 = floor(rmin) - 1.0, fr1 = ceil(rmax) + 1.0;
  const double fc0 = floor(cmin) - 1.0, fc1 = ceil(cmax) + 1.0;
  bool interior = fr0 >= 0.0 && fc0 >= 0.0 && fr1 <= (double)(H - 1) && fc1 <= (double)(W - 1);
  const int r0 = (int)fmax(fr0, 0.0), c0 = ((int)fmax(fc0, 0.0)) & ~(align - 1);
  int rows = (int)fmin(fr1, (double)(H - 1)) - r0 + 1;
  const int cols = (int)fmin(fc1, (double)(W - 1)) - c0 + 1;
  if (rows <= 0 || cols <= 0) return;                       // window entirely off-frame
  const int pitch = min((cols + align - 1) & ~(align - 1), W - c0);   // W % align == 0
  if (rows * pitch > cap) {
    // too large for the tile: stage the rows that fit; the taps below them come from global memory
    // (bounds-checked path), which is still most of the traffic saved
    rows = cap / pitch;
    interior = false;
    if (rows <= 0) return;
  }
  x.box_lo = (uint32_t)r0 | ((uint32_t)c0 << 13) | (interior ? (1u << 26) : 0u) | (1u << 27);
  x.box_hi = (uint32_t)rows | ((uint32_t)pitch << 13);
}

constexpr int kTformWarps = 4;

// everything tform_frame needs (the arguments of tform_kernel)
struct TformArgs {
  const double* lm;
  const uint8_t* valid;
  const int64_t* clip_offsets;
  int64_t n_clips;
  const double* mean_face;
  const double* tforms_in;
  int std_size, roi, window, fp_lo, fp_span, H, W, fp_align, fp_cap;
  int32_t* crop_rc;
  double* tforms_out;
  // collation (avfe_lip_roi_collate): frame t of clip c goes to slot c*T_pad + t of the padded
  // [n_clips, T_pad] video batch when t < min(keep[c], T_pad) and is dropped otherwise;
  // T_pad == 0: packed output, frame f goes to slot f
  const int64_t* keep;
  int64_t T_pad;
};

// Window-smoothed similarity fit + cut_patch origin + footprint of frame f, computed by one warp
// (V2 on the fly, V3, V4 fit, V6, V7).  The record is returned in every lane; lane 0 also writes
// the optional crop_rc / tforms outputs.
__device__ __forceinline__ FrameXform tform_frame(const TformArgs& a, int64_t f, int lane, int64_t& dst) {
  const unsigned full = 0xffffffffu;
  const int64_t c = find_clip_warp(a.clip_offsets, a.n_clips, f, lane);
  LmView v{a.lm, a.valid, a.clip_offsets[c], a.clip_offsets[c + 1]};
  const int64_t T = v.end - v.beg;
  dst = f;
  if (a.T_pad > 0) {
    int64_t kept = a.keep != nullptr ? a.keep[c] : T;
    if (kept > a.T_pad) kept = a.T_pad;
    const int64_t t = f - v.beg;
    dst = (t < kept) ? c * a.T_pad + t : -1;
  }
  // margin = min(T, 12); frame i <= T-margin is fitted on mean(lm[i:i+margin]); later frames
  // reuse the transform of frame T-margin (preprocess/video_process.py:369-370,417-427,455-464)
  const int margin = (int)(T < a.window ? T : a.window);   // window <= 31 is enforced by the caller
  int64_t i0 = f - v.beg;
  if (i0 > T - margin) i0 = T - margin;
  const int64_t w0 = v.beg + i0;
  // lanes 0..margin-1 look up the neighbours of window frame `lane`, lane 31 those of frame f
  const int64_t mine = (lane < margin) ? (w0 + lane) : f;
  int64_t p, q;
  lm_neighbours(v, mine, p, q);

  double fwd[6], inv[6];
  if (a.tforms_in != nullptr) {
    const double* ti = a.tforms_in + f * 18;
#pragma unroll
    for (int k = 0; k < 6; ++k) { fwd[k] = ti[k]; inv[k] = ti[9 + k]; }
  } else {
    // np.mean(axis=0) of the 5 stable points (33,36,39,42,45) adds the window rows in order.
    // slot = one coordinate of one point.  Loads first: lane (slot, grp) fetches window frames
    // grp, grp+3, grp+6, grp+9, so a 12-frame window costs one round of loads instead of twelve
    // dependent ones; then lanes sum their slot's values in row order through shuffles.
    const int grp = lane / 10, slot = lane - 10 * grp;  // lanes 30, 31 (grp 3) fetch nothing
    const int e = (33 + 3 * (slot >> 1)) * 2 + (slot & 1);
    double acc = 0.0;
    if (margin <= 12) {
      LmTaps taps[4];
      int64_t pj[4], qj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = 3 * u + grp;                      // <= 12
        pj[u] = __shfl_sync(full, p, jj);
        qj[u] = __shfl_sync(full, q, jj);
        const bool mine_u = grp < 3 && jj < margin;
        if (!mine_u) { pj[u] = f; qj[u] = f; }           // harmless in-range address
        taps[u] = lm_taps(v, mine_u ? w0 + jj : f, pj[u], qj[u], e);
      }
      double vals[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = 3 * u + grp;
        vals[u] = (grp < 3 && jj < margin) ? lm_blend(v, w0 + jj, pj[u], qj[u], taps[u]) : 0.0;
      }
#pragma unroll
      for (int jj = 0; jj < 12; ++jj) {
        const double t = __shfl_sync(full, vals[jj / 3], slot + 10 * (jj % 3));
        if (jj < margin) acc = f64add(acc, t);
      }
    } else {
      for (int j = 0; j < margin; ++j) {
        const int64_t pw = __shfl_sync(full, p, j), qw = __shfl_sync(full, q, j);
        acc = f64add(acc, lm_value(v, w0 + j, pw, qw, e));
      }
    }
    const double mean = f64div(acc, (double)margin);
    double src[kNumStable][2], dst[kNumStable][2];
#pragma unroll
    for (int k = 0; k < kNumStable; ++k) {
      src[k][0] = __shfl_sync(full, mean, 2 * k);
      src[k][1] = __shfl_sync(full, mean, 2 * k + 1);
      dst[k][0] = a.mean_face[(33 + 3 * k) * 2 + 0];
      dst[k][1] = a.mean_face[(33 + 3 * k) * 2 + 1];
    }
    similarity_fit(src, dst, kNumStable, fwd);          // every lane computes the same fit
    affine_inverse(fwd, inv);
  }
  // trans(cur_landmarks)[48:68] -> mean -> cut_patch origin; lane k < 20 transforms point 48+k
  const int64_t pf = __shfl_sync(full, p, 31), qf = __shfl_sync(full, q, 31);
  const int k = 48 + (lane % 20);
  const LmTaps xt = lm_taps(v, f, pf, qf, 2 * k), yt = lm_taps(v, f, pf, qf, 2 * k + 1);
  const double x = lm_blend(v, f, pf, qf, xt), y = lm_blend(v, f, pf, qf, yt);
  const double tx = f64add(f64add(f64mul(x, fwd[0]), f64mul(y, fwd[1])), fwd[2]);
  const double ty = f64add(f64add(f64mul(x, fwd[3]), f64mul(y, fwd[4])), fwd[5]);
  double cx = 0.0, cy = 0.0;
#pragma unroll
  for (int j = 0; j < 20; ++j) {
    cx = f64add(cx, __shfl_sync(full, tx, j));
    cy = f64add(cy, __shfl_sync(full, ty, j));
  }
  cx = f64div(cx, 20.0);
  cy = f64div(cy, 20.0);
  int r0, c0;
  crop_origin(cx, cy, a.roi / 2, a.roi / 2, a.std_size, a.std_size, &r0, &c0);
  FrameXform o;
#pragma unroll
  for (int j = 0; j < 6; ++j) o.inv[j] = inv[j];
  o.r0 = r0; o.c0 = c0;
  pack_footprint(o, a.fp_lo, a.fp_span, a.H, a.W, a.fp_align, a.fp_cap);
  if (lane == 0) {
    if (a.crop_rc != nullptr) { a.crop_rc[2 * f] = r0; a.crop_rc[2 * f + 1] = c0; }
    if (a.tforms_out != nullptr) {
      double* to = a.tforms_out + f * 18;
#pragma unroll
      for (int j = 0; j < 6; ++j) { to[j] = fwd[j]; to[9 + j] = inv[j]; }
      to[6] = 0.0; to[7] = 0.0; to[8] = 1.0;
      to[15] = 0.0; to[16] = 0.0; to[17] = 1.0;
    }
  }
  return o;
}

__global__ void __launch_bounds__(kTformWarps * 32)
tform_kernel(const TformArgs a, int64_t N, FrameXform* __restrict__ xf, int64_t* __restrict__ dst_slot,
             unsigned* __restrict__ queue_counter) {
  if (queue_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *queue_counter = 0u;
  const int lane = threadIdx.x & 31;
  const int64_t f = (int64_t)blockIdx.x * kTformWarps + (threadIdx.x >> 5);
  if (f >= N) return;                                  // warp-uniform
  int64_t dst;
  const FrameXform o = tform_frame(a, f, lane, dst);
  if (lane == 0) { xf[f] = o; dst_slot[f] = dst; }
}

// ------------------------------------------------------------------ V1: BGR -> gray
// 16 pixels from 48 packed bytes held in 12 words.  PRMT aligns each BGR triple, two dp4a
// evaluate the 15-bit dot product as (hi<<8)+lo with coefficient bytes
// 3735 = 14*256+151, 19235 = 75*256+35, 9798 = 38*256+70.
__device__ __forceinline__ uint32_t gray_dp4a(uint32_t bgrx) {
  const uint32_t lo = __dp4a(bgrx, 0x00462397u, 16384u);
  const uint32_t hi = __dp4a(bgrx, 0x00264B0Eu, 0u);
  return (hi * 256u + lo) >> 15;
}

__device__ __forceinline__ uint4 gray16(const uint32_t (&w)[12]) {
  uint32_t y[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int k = 3 * p, wi = k >> 2, o = k & 3;
    const uint32_t sel = (uint32_t)(o | ((o + 1) << 4) | ((o + 2) << 8) | (o << 12));
    const uint32_t hiw = (wi + 1 < 12) ? w[wi + 1] : 0u;
    y[p] = gray_dp4a(__byte_perm(w[wi], hiw, sel));
  }
  uint4 r;
  r.x = y[0] | (y[1] << 8) | (y[2] << 16) | (y[3] << 24);
  r.y = y[4] | (y[5] << 8) | (y[6] << 16) | (y[7] << 24);
  r.z = y[8] | (y[9] << 8) | (y[10] << 16) | (y[11] << 24);
  r.w = y[12] | (y[13] << 8) | (y[14] << 16) | (y[15] << 24);
  return r;
}

constexpr int kGrayThreads = 256;
constexpr int kGrayWarps = kGrayThreads / 32;

// One warp converts one group of 512 px: 96 coalesced 16-byte loads (3 per lane) into the
// warp's 1536-byte shared slab, then every lane reads back its own 48 contiguous bytes
// (stride 48 B = 12 banks: conflict-free for 128-bit accesses) and stores 16 gray bytes.
__device__ __forceinline__ void gray_group(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                           uint4* slab, int lane) {
  const uint4 a = ldg_stream(src + lane);
  const uint4 b = ldg_stream(src + lane + 32);
  const uint4 c = ldg_stream(src + lane + 64);
  slab[lane] = a;
  slab[lane + 32] = b;
  slab[lane + 64] = c;
  __syncwarp();
  uint32_t w[12];
  const uint4 q0 = slab[3 * lane], q1 = slab[3 * lane + 1], q2 = slab[3 * lane + 2];
  __syncwarp();
  w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
  w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
  w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
  stg_stream(dst + lane, gray16(w));
}

__global__ void __launch_bounds__(kGrayThreads)
gray_vec_kernel(const uint4* __restrict__ bgr, int64_t ngroups /* of 512 px */,
                uint4* __restrict__ gray) {
  __shared__ uint4 slab[kGrayWarps][96];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * kGrayWarps + wid;
  const int64_t nwarps = (int64_t)gridDim.x * kGrayWarps;
  for (int64_t g = warp_global; g < ngroups; g += nwarps)
    gray_group(bgr + g * 96, gray + g * 32, slab[wid], lane);
}

// tail / unaligned path: one pixel per thread
__global__ void __launch_bounds__(256)
gray_scalar_kernel(const uint8_t* __restrict__ bgr, int64_t first_px, int64_t npx,
                   uint8_t* __restrict__ gray) {
  for (int64_t i = first_px + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx;
       i += (int64_t)gridDim.x * blockDim.x) {
    gray[i] = (uint8_t)gray_from_bgr(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
  }
}

static int launch_gray(const uint8_t* bgr, int64_t npx, uint8_t* gray, cudaStream_t s) {
  if (npx == 0) return AVFE_OK;
  int64_t done = 0;
  if (aligned16(bgr) && aligned16(gray) && npx >= 512) {
    const int64_t ngroups = npx / 512;
    int64_t ctas = (ngroups + kGrayWarps - 1) / kGrayWarps;
    const int64_t cap = (int64_t)kNumSMs * 8 * 4;   // 8 resident CTAs/SM, a few waves each
    if (ctas > cap) ctas = cap;
    gray_vec_kernel<<<(unsigned)ctas, kGrayThreads, 0, s>>>(
        reinterpret_cast<const uint4*>(bgr), ngroups, reinterpret_cast<uint4*>(gray));
    count_launch();
    done = ngroups * 512;
  }
  if (done < npx) {
    int64_t rem = npx - done;
    int64_t ctas = (rem + 255) / 256;
    if (ctas > (int64_t)kNumSMs * 8) ctas = (int64_t)kNumSMs * 8;
    gray_scalar_kernel<<<(unsigned)ctas, 256, 0, s>>>(bgr, done, npx, gray);
    count_launch();
  }
  return check_launch();
}

}  // namespace avfe

#include "avfe_lip_queue.cuh"   // lip_fused_kernel: stream warps + compute warps
#include "avfe_lip_frame.cuh"   // lip_frame_kernel: frame-owner CTAs, every BGR byte read once

namespace avfe {

// ------------------------------------------------------------------ single-purpose kernels
__device__ __forceinline__ void fill_lut255(double* lut255) {
  for (int k = threadIdx.x; k < 256; k += blockDim.x) lut255[k] = f64div((double)k, 255.0);
}

// Full-frame warp with a caller-supplied inverse matrix (affine or projective).
__global__ void __launch_bounds__(256)
warp_full_kernel(const uint8_t* __restrict__ gray, int H, int W, const double* __restrict__ M,
                 int out_h, int out_w, uint8_t* __restrict__ out) {
  __shared__ double lut255[256];
  fill_lut255(lut255);
  __syncthreads();
  const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5];
  const double m6 = M[6], m7 = M[7], m8 = M[8];
  const bool affine = (m6 == 0.0) && (m7 == 0.0) && (m8 == 1.0);
  auto tap = [&](int r, int c) -> double { return lut255[__ldg(gray + (size_t)r * W + c)]; };
  const int64_t total = (int64_t)out_h * out_w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const double tfr = (double)(idx / out_w), tfc = (double)(idx % out_w);
    double sc = f64add(f64add(f64mul(m0, tfc), f64mul(m1, tfr)), m2);
    double sr = f64add(f64add(f64mul(m3, tfc), f64mul(m4, tfr)), m5);
    if (!affine) {                          // _transform_projective
      const double z = f64add(f64add(f64mul(m6, tfc), f64mul(m7, tfr)), m8);
      sc = f64div(sc, z);
      sr = f64div(sr, z);
    }
    out[idx] = bilinear_u8(sr, sc, H, W, tap);
  }
}

// V8 on an existing ROI stack: /255, centre crop, (x-mean)/std, all in float32.
__global__ void __launch_bounds__(256)
video_feats_kernel(const uint8_t* __restrict__ roi, int64_t N, int Hin, int Win, int crop,
                   float mean, float stdv, float* __restrict__ out) {
  __shared__ float lutn[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x)
    lutn[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)k, 255.0f), mean), stdv);
  __syncthreads();
  const int sh = (Hin - crop) / 2, sw = (Win - crop) / 2;
  const int64_t per = (int64_t)crop * crop, total = N * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / per;
    const int r = (int)((i % per) / crop), c = (int)(i % crop);
    out[i] = lutn[roi[(f * Hin + sh + r) * Win + sw + c]];
  }
}

// single similarity fit (estimate_transform('similarity', src, dst)) -> fwd 3x3 + inverse 3x3
__global__ void similarity_fit_kernel(const double* __restrict__ src, const double* __restrict__ dst,
                                      int n, double* __restrict__ out18) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double fwd[6], inv[6];
  similarity_fit(reinterpret_cast<const double(*)[2]>(src),
                 reinterpret_cast<const double(*)[2]>(dst), n, fwd);
  affine_inverse(fwd, inv);
  for (int k = 0; k < 6; ++k) { out18[k] = fwd[k]; out18[9 + k] = inv[k]; }
  out18[6] = 0.0; out18[7] = 0.0; out18[8] = 1.0;
  out18[15] = 0.0; out18[16] = 0.0; out18[17] = 1.0;
}

// cut_patch (utils/lips_cropping.py:127-163) on one 2-D uint8 image
__global__ void __launch_bounds__(256)
cut_patch_kernel(const uint8_t* __restrict__ img, int H, int W, const double* __restrict__ lm, int n,
                 int half_h, int half_w, uint8_t* __restrict__ out, int32_t* __restrict__ rc) {
  __shared__ int s_rc[2];
  if (threadIdx.x == 0) {
    double cx = 0.0, cy = 0.0;
    for (int k = 0; k < n; ++k) { cx = f64add(cx, lm[2 * k]); cy = f64add(cy, lm[2 * k + 1]); }
    cx = f64div(cx, (double)n);
    cy = f64div(cy, (double)n);
    crop_origin(cx, cy, half_h, half_w, H, W, &s_rc[0], &s_rc[1]);
    if (rc != nullptr && blockIdx.x == 0) { rc[0] = s_rc[0]; rc[1] = s_rc[1]; }
  }
  __syncthreads();
  const int r0 = s_rc[0], c0 = s_rc[1], oh = 2 * half_h, ow = 2 * half_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < oh * ow; i += gridDim.x * blockDim.x) {
    const int r = r0 + i / ow, c = c0 + i % ow;
    // numpy slicing semantics: rows/cols outside the image simply do not exist; the clamps in
    // crop_origin keep the window inside whenever the image is at least patch-sized
    out[i] = (r0 >= 0 && r >= 0 && r < H && c >= 0 && c < W) ? img[(size_t)r * W + c] : 0;
  }
}

// Padded part of a collated video batch: frames [kept_c, T_pad) of every clip are zero (the
// upstream collator's np.pad) and flagged in padding_mask (1 = padding).
__global__ void __launch_bounds__(256)
collate_tail_kernel(const int64_t* __restrict__ clip_offsets, const int64_t* __restrict__ keep, int64_t T_pad,
                    int frame_floats, float* __restrict__ video, uint8_t* __restrict__ padding_mask) {
  const int64_t c = blockIdx.y;
  int64_t kept = clip_offsets[c + 1] - clip_offsets[c];
  if (keep != nullptr && keep[c] < kept) kept = keep[c];
  if (kept > T_pad) kept = T_pad;
  if (kept < 0) kept = 0;
  if (padding_mask != nullptr)
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T_pad; t += (int64_t)gridDim.x * blockDim.x)
      padding_mask[c * T_pad + t] = t >= kept ? 1 : 0;
  if (video == nullptr) return;
  float* base = video + (c * T_pad + kept) * frame_floats;
  const int64_t n = (T_pad - kept) * frame_floats;
  if ((frame_floats & 3) == 0 && (reinterpret_cast<uintptr_t>(video) & 15u) == 0) {
    float4* b4 = reinterpret_cast<float4*>(base);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (int64_t)gridDim.x * blockDim.x)
      b4[i] = z;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      base[i] = 0.0f;
  }
}

}  // namespace avfe

// ====================================================================== C ABI
using namespace avfe;

template <bool STREAM, int SPAN>
static int launch_fused(const LipJob& j, cudaStream_t s) {
  // without stream warps the ring space holds footprint slots 2..5
  const int smem = STREAM ? (int)sizeof(FusedSmem)
                          : (int)(offsetof(FusedSmem, raw) + sizeof(uint4) * kQueueSlots * (kTilePx * 3 / 16));
  if (cudaFuncSetAttribute(lip_fused_kernel<STREAM, SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           smem) != cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  const int threads = STREAM ? 1024 : Roles<SPAN, kQueueSlots>::kHandoverThreads;
  const int64_t ctas = j.N < kNumSMs ? j.N : kNumSMs;                  // one persistent CTA per SM
  lip_fused_kernel<STREAM, SPAN><<<(unsigned)(STREAM ? kNumSMs : ctas), threads, smem, s>>>(j);
  return AVFE_OK;
}

template <int SPAN>
static int launch_frame(const FrameJob& j, cudaStream_t s) {
  const int smem = (int)sizeof(FrameSmem<SPAN>);
  if (cudaFuncSetAttribute(lip_frame_kernel<SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return AVFE_ERR_CUDA;
  }
  const int64_t ctas = j.lip.N < kNumSMs ? j.lip.N : kNumSMs;          // one persistent CTA per SM
  lip_frame_kernel<SPAN><<<(unsigned)ctas, 1024, smem, s>>>(j);
  return AVFE_OK;
}

extern "C" int avfe_bgr2gray_u8(const uint8_t* bgr, int64_t N, int H, int W, uint8_t* gray,
                                avfe_stream_t stream) {
  if (N < 0 || H < 0 || W < 0) return AVFE_ERR_INVALID_ARG;
  const int64_t npx = N * (int64_t)H * W;
  if (npx == 0) return AVFE_OK;
  if (!bgr || !gray) return AVFE_ERR_INVALID_ARG;
  return launch_gray(bgr, npx, gray, static_cast<cudaStream_t>(stream));
}

extern "C" int avfe_warp_affine_u8(const uint8_t* gray, int H, int W, const double* inv_matrix,
                                   int out_h, int out_w, uint8_t* out, avfe_stream_t stream) {
  if (H <= 0 || W <= 0 || out_h < 0 || out_w < 0) return AVFE_ERR_INVALID_ARG;
  if (out_h == 0 || out_w == 0) return AVFE_OK;
  if (!gray || !inv_matrix || !out) return AVFE_ERR_INVALID_ARG;
  const int64_t total = (int64_t)out_h * out_w;
  int64_t ctas = (total + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 8) ctas = (int64_t)kNumSMs * 8;
  warp_full_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gray, H, W, inv_matrix, out_h, out_w, out);
  count_launch();
  return check_launch();
}

extern "C" size_t avfe_lip_workspace_bytes(int64_t N) {
  if (N < 0) return 0;
  // work-queue counter + one FrameXform and one output slot index per frame
  return (size_t)N * (sizeof(FrameXform) + sizeof(int64_t)) + 256;
}

static int lip_roi_impl(const uint8_t* frames, int channels, int64_t N, int H, int W,
                        const int64_t* clip_offsets, int64_t n_clips,
                        const double* landmarks, const uint8_t* lm_valid,
                        const double* mean_face, const double* tforms_in,
                        int std_size, int roi, int crop, int window, float mean,
                        float std, uint8_t* gray_out, uint8_t* lip_u8, float* lip_f32,
                        int32_t* crop_rc, double* tforms, const int64_t* keep, int64_t T_pad,
                        void* workspace, size_t workspace_bytes, avfe_stream_t stream) {
  if (N < 0 || n_clips < 0 || H <= 0 || W <= 0) return AVFE_ERR_INVALID_ARG;
  if (channels != 1 && channels != 3) return AVFE_ERR_INVALID_ARG;
  if (roi <= 0 || (roi & 1) || crop <= 0 || crop > roi || ((roi - crop) & 1) || window <= 0 ||
      std_size < roi)
    return AVFE_ERR_INVALID_ARG;
  if (roi > kMaxRoi || window > 31) return AVFE_ERR_UNSUPPORTED;
  if (N == 0 || n_clips == 0) return AVFE_OK;
  if (!frames || !clip_offsets || !landmarks || !mean_face) return AVFE_ERR_INVALID_ARG;
  if (gray_out && channels != 3) return AVFE_ERR_INVALID_ARG;
  if (N > 0x0fffffffLL) return AVFE_ERR_UNSUPPORTED;
  // the one pointer that may be mapped pinned host memory (zero-copy mode, avfe.h "Conventions")
  bool frames_on_host = false;
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, frames) == cudaSuccess) frames_on_host = (pa.type == cudaMemoryTypeHost);
    else cudaGetLastError();
  }
  if (frames_on_host && gray_out) return AVFE_ERR_INVALID_ARG;     // gray needs every pixel on the device
  if (!workspace || workspace_bytes < avfe_lip_workspace_bytes(N) || !aligned16(workspace))
    return AVFE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  unsigned* counter = static_cast<unsigned*>(workspace);
  FrameXform* xf = reinterpret_cast<FrameXform*>(static_cast<char*>(workspace) + 256);
  int64_t* dst_slot = reinterpret_cast<int64_t*>(xf + N);
  // footprint staging: 16-byte cp.async when frame rows keep that alignment, else 4-byte, else none
  const int C = channels;
  const uintptr_t fbase = reinterpret_cast<uintptr_t>(frames);
  int fp_align = 0;
  if (W % 16 == 0 && (fbase & 15u) == 0) fp_align = 16;
  else if (W % 4 == 0 && (fbase & 3u) == 0) fp_align = 4;
  (void)C;
  const int fp_lo = lip_u8 ? 0 : (roi - crop) / 2, fp_span = lip_u8 ? roi : crop;
  TformArgs ta;
  ta.lm = landmarks; ta.valid = lm_valid; ta.clip_offsets = clip_offsets; ta.n_clips = n_clips;
  ta.mean_face = mean_face; ta.tforms_in = tforms_in; ta.std_size = std_size; ta.roi = roi;
  ta.window = window; ta.fp_lo = fp_lo; ta.fp_span = fp_span; ta.H = H; ta.W = W; ta.fp_align = fp_align;
  ta.fp_cap = kTilePx;
  ta.crop_rc = crop_rc; ta.tforms_out = tforms; ta.keep = keep; ta.T_pad = T_pad;

  // window side known at compile time for the two standard configurations
  const int span_sel = (lip_u8 != nullptr && roi == 96) ? 96
                     : (lip_u8 == nullptr && lip_f32 != nullptr && crop == 88) ? 88 : 0;
  // Standard case: BGR frames whose rows keep 16-byte alignment, gray frames wanted.  One launch
  // in which each CTA owns whole frames (fit + gray + footprint + blend), every byte read once.
  if (span_sel != 0 && channels == 3 && fp_align == 16 && gray_out != nullptr && aligned16(gray_out) &&
      W >= 32 && W <= 8191 && H <= 8191 && (int64_t)H * W / 16 < (1 << 22)) {
    FrameJob fj;
    LipJob& j = fj.lip;
    j.frames = frames; j.channels = channels; j.H = H; j.W = W; j.N = N; j.xf = nullptr; j.dst_slot = nullptr;
    j.roi = roi; j.crop = crop; j.mean = mean; j.stdv = std;
    j.gray_out = gray_out; j.lip_u8 = lip_u8; j.lip_f32 = lip_f32; j.counter = nullptr;
    j.ngroups = 0; j.stage_align = fp_align; j.host_frames = 0;
    fj.tf = ta;
    fj.tf.fp_cap = kFrameTilePx;                            // u8 tiles: twice the generic kernel's capacity
    fj.groups_per_frame = (int)((int64_t)H * W / 16);
    fj.chunks_per_frame = (fj.groups_per_frame + 63) / 64;
    fj.row_groups = W / 16;
    fj.row_magic = (unsigned)(0x100000000ULL / (unsigned)fj.row_groups) + 1u;
    const int rc = (span_sel == 96) ? launch_frame<96>(fj, s) : launch_frame<88>(fj, s);
    if (rc != AVFE_OK) return rc;
    count_launch();
    return check_launch();
  }

  tform_kernel<<<(unsigned)((N + kTformWarps - 1) / kTformWarps), kTformWarps * 32, 0, s>>>(ta, N, xf, dst_slot, counter);
  count_launch();

  const int64_t npx = (int64_t)H * W;
  const bool want_roi = (lip_u8 != nullptr) || (lip_f32 != nullptr);
  // the gray conversion rides along in the fused launch when the pixel stream is 16-byte
  // aligned; otherwise (or when no ROI is wanted) the flat streaming kernel does it
  const bool fuse_gray = want_roi && gray_out != nullptr && aligned16(frames) && aligned16(gray_out) &&
                         N * npx >= 1024;
  if (gray_out != nullptr && !fuse_gray) {
    int rc = launch_gray(frames, N * npx, gray_out, s);
    if (rc != AVFE_OK) return rc;
  }
  if (want_roi) {
    LipJob j;
    j.frames = frames; j.channels = channels; j.H = H; j.W = W; j.N = N; j.xf = xf; j.dst_slot = dst_slot;
    j.roi = roi; j.crop = crop; j.mean = mean; j.stdv = std;
    j.gray_out = fuse_gray ? gray_out : nullptr;
    j.lip_u8 = lip_u8; j.lip_f32 = lip_f32; j.counter = counter;
    j.ngroups = (N * npx) / 512;
    j.stage_align = fp_align;
    j.host_frames = frames_on_host ? 1 : 0;
    int rc = AVFE_OK;
    if (fuse_gray) {
      if (span_sel == 96) rc = launch_fused<true, 96>(j, s);
      else if (span_sel == 88) rc = launch_fused<true, 88>(j, s);
      else rc = launch_fused<true, 0>(j, s);
    } else {
      if (span_sel == 96) rc = launch_fused<false, 96>(j, s);
      else if (span_sel == 88) rc = launch_fused<false, 88>(j, s);
      else rc = launch_fused<false, 0>(j, s);
    }
    if (rc != AVFE_OK) return rc;
    count_launch();
  }
  return check_launch();
}

extern "C" int avfe_lip_roi_batch(const uint8_t* frames, int channels, int64_t N, int H, int W,
                                  const int64_t* clip_offsets, int64_t n_clips,
                                  const double* landmarks, const uint8_t* lm_valid,
                                  const double* mean_face, const double* tforms_in,
                                  int std_size, int roi, int crop, int window, float mean,
                                  float std, uint8_t* gray_out, uint8_t* lip_u8, float* lip_f32,
                                  int32_t* crop_rc, double* tforms, void* workspace,
                                  size_t workspace_bytes, avfe_stream_t stream) {
  return lip_roi_impl(frames, channels, N, H, W, clip_offsets, n_clips, landmarks, lm_valid, mean_face,
                      tforms_in, std_size, roi, crop, window, mean, std, gray_out, lip_u8, lip_f32, crop_rc,
                      tforms, nullptr, 0, workspace, workspace_bytes, stream);
}

extern "C" int avfe_lip_roi_collate(const uint8_t* frames, int channels, int64_t N, int H, int W,
                                    const int64_t* clip_offsets, int64_t n_clips,
                                    const double* landmarks, const uint8_t* lm_valid,
                                    const double* mean_face, const double* tforms_in,
                                    int std_size, int roi, int crop, int window, float mean,
                                    float std, const int64_t* keep_frames, int64_t T_pad,
                                    uint8_t* gray_out, float* video, uint8_t* padding_mask,
                                    void* workspace, size_t workspace_bytes, avfe_stream_t stream) {
  if (T_pad <= 0 || n_clips < 0 || crop <= 0) return AVFE_ERR_INVALID_ARG;
  if (n_clips == 0) return AVFE_OK;
  if (!video || !clip_offsets) return AVFE_ERR_INVALID_ARG;
  if (n_clips > 65535) return AVFE_ERR_UNSUPPORTED;
  if (N > 0) {
    const int rc = lip_roi_impl(frames, channels, N, H, W, clip_offsets, n_clips, landmarks, lm_valid,
                                mean_face, tforms_in, std_size, roi, crop, window, mean, std, gray_out,
                                nullptr, video, nullptr, nullptr, keep_frames, T_pad, workspace,
                                workspace_bytes, stream);
    if (rc != AVFE_OK) return rc;
  }
  int64_t gx = (T_pad * (int64_t)crop * crop / 4 + 255) / 256;
  if (gx > 8) gx = 8;                                     // the padded part is a few frames per clip
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)n_clips);
  collate_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(clip_offsets, keep_frames, T_pad,
                                                                       crop * crop, video, padding_mask);
  count_launch();
  return check_launch();
}

extern "C" int avfe_video_feats_u8(const uint8_t* roi_u8, int64_t N, int Hin, int Win, int crop,
                                   float mean, float std, float* out, avfe_stream_t stream) {
  if (N < 0 || Hin <= 0 || Win <= 0 || crop <= 0 || crop > Hin || crop > Win)
    return AVFE_ERR_INVALID_ARG;
  if (N == 0) return AVFE_OK;
  if (!roi_u8 || !out) return AVFE_ERR_INVALID_ARG;
  const int64_t total = N * (int64_t)crop * crop;
  int64_t ctas = (total + 255) / 256;
  if (ctas > (int64_t)kNumSMs * 16) ctas = (int64_t)kNumSMs * 16;
  video_feats_kernel<<<(unsigned)ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      roi_u8, N, Hin, Win, crop, mean, std, out);
  count_launch();
  return check_launch();
}

extern "C" int avfe_landmarks_interpolate(const double* landmarks, const uint8_t* lm_valid,
                                          const int64_t* clip_offsets, int64_t n_clips, int64_t N,
                                          double* out, avfe_stream_t stream) {
  if (N < 0 || n_clips < 0) return AVFE_ERR_INVALID_ARG;
  if (N == 0 || n_clips == 0) return AVFE_OK;
  if (!landmarks || !lm_valid || !clip_offsets || !out) return AVFE_ERR_INVALID_ARG;
  if (N > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  lm_fill_kernel<<<(unsigned)N, 160, 0, static_cast<cudaStream_t>(stream)>>>(
      landmarks, lm_valid, clip_offsets, n_clips, N, out);
  count_launch();
  return check_launch();
}

extern "C" int avfe_similarity_fit(const double* src, const double* dst, int n, double* out18,
                                   avfe_stream_t stream) {
  if (n <= 0 || !src || !dst || !out18) return AVFE_ERR_INVALID_ARG;
  similarity_fit_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, out18);
  count_launch();
  return check_launch();
}

extern "C" int avfe_cut_patch_u8(const uint8_t* img, int H, int W, const double* landmarks, int n,
                                 int half_h, int half_w, uint8_t* out, int32_t* rc,
                                 avfe_stream_t stream) {
  if (H <= 0 || W <= 0 || n <= 0 || half_h <= 0 || half_w <= 0) return AVFE_ERR_INVALID_ARG;
  if (!img || !landmarks || !out) return AVFE_ERR_INVALID_ARG;
  int ctas = (4 * half_h * half_w + 255) / 256;
  if (ctas > kNumSMs) ctas = kNumSMs;
  cut_patch_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, H, W, landmarks, n,
                                                                       half_h, half_w, out, rc);
  count_launch();
  return check_launch();
}
