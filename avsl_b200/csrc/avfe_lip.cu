// Lip-ROI path: BGR->gray, landmark interpolation, window-smoothed similarity fit, bilinear
// warp restricted to the cut_patch window, centre crop and normalisation.
// Reference: preprocess/video_process.py:201-214,369-475; utils/lips_cropping.py:41-163;
// utils/hf_video_utils.py:113-138.
//
// Kernels (all HBM-bound integer/byte work except the float64 blend):
//   lm_fill_kernel   one CTA per frame; fills failed detections (V2)
//   tform_kernel     one thread per frame; window mean, similarity fit, inverse, crop origin
//   gray_vec_kernel  streaming BGR->gray, 16 px per thread, coalesced 128-bit loads staged
//                    through warp-private shared memory so each thread owns 48 contiguous bytes
//   warp_kernel      one CTA per frame; float64 bilinear taps + u8 ROI + normalised f32 crop
#include "avfe_common.cuh"
#include "avfe_lip_math.cuh"

namespace avfe {

// ------------------------------------------------------------------ V2: landmark fill
__global__ void __launch_bounds__(160)
lm_fill_kernel(const double* __restrict__ lm, const uint8_t* __restrict__ valid,
               const int64_t* __restrict__ clip_offsets, int64_t n_clips, int64_t N,
               double* __restrict__ out) {
  const int64_t f = blockIdx.x;
  if (f >= N) return;
  // locate the clip: binary search over the (small) offsets array
  int64_t lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (clip_offsets[mid] <= f) lo = mid; else hi = mid;
  }
  const int64_t beg = clip_offsets[lo], end = clip_offsets[lo + 1];
  int64_t p = f, q = f;                       // previous / next valid frame of the clip
  while (p >= beg && !valid[p]) --p;
  while (q < end && !valid[q]) ++q;
  const int t = threadIdx.x;
  if (t >= kNumLandmarks * 2) return;
  double v;
  if (p == f) {
    v = lm[f * 136 + t];
  } else if (p < beg && q >= end) {
    v = nan("");                              // no detection in the whole clip
  } else if (p < beg) {
    v = lm[q * 136 + t];                      // leading frames replicate the first detection
  } else if (q >= end) {
    v = lm[p * 136 + t];                      // trailing frames replicate the last detection
  } else {
    // start + idx/float(stop-start) * delta   (utils/lips_cropping.py:54-57)
    const double s = lm[p * 136 + t], e = lm[q * 136 + t];
    const double w = f64div((double)(f - p), (double)(q - p));
    v = f64add(s, f64mul(w, f64sub(e, s)));
  }
  out[f * 136 + t] = v;
}

// ------------------------------------------------------------------ V3+V4(fit)+V6+V7
// per-frame record written to the workspace and consumed by warp_kernel
struct FrameXform {
  double inv[6];   // rows 0,1 of tform.inverse.params
  int32_t r0, c0;  // cut_patch origin in the std frame (or -1,-1)
  int32_t pad[2];
};

__global__ void __launch_bounds__(128)
tform_kernel(const double* __restrict__ lm, const int64_t* __restrict__ clip_offsets,
             int64_t n_clips, int64_t N, const double* __restrict__ mean_face,
             const double* __restrict__ tforms_in, int std_size, int roi, int window,
             FrameXform* __restrict__ xf, int32_t* __restrict__ crop_rc,
             double* __restrict__ tforms_out) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  int64_t lo = 0, hi = n_clips;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (clip_offsets[mid] <= f) lo = mid; else hi = mid;
  }
  const int64_t beg = clip_offsets[lo];
  const int64_t T = clip_offsets[lo + 1] - beg;
  const int stable[kNumStable] = {33, 36, 39, 42, 45};

  double fwd[6], inv[6];
  if (tforms_in != nullptr) {
    const double* ti = tforms_in + f * 18;
    fwd[0] = ti[0]; fwd[1] = ti[1]; fwd[2] = ti[2]; fwd[3] = ti[3]; fwd[4] = ti[4]; fwd[5] = ti[5];
    inv[0] = ti[9]; inv[1] = ti[10]; inv[2] = ti[11]; inv[3] = ti[12]; inv[4] = ti[13]; inv[5] = ti[14];
  } else {
    // margin = min(T, 12); frame i <= T-margin is fitted on mean(lm[i:i+margin]); later
    // frames reuse the transform of frame T-margin (preprocess/video_process.py:369-370,
    // 417-427,455-464).
    const int64_t margin = T < window ? T : window;
    int64_t i = f - beg;
    if (i > T - margin) i = T - margin;
    const double* base = lm + (beg + i) * 136;
    double src[kNumStable][2], dst[kNumStable][2];
#pragma unroll
    for (int k = 0; k < kNumStable; ++k) {
      double ax = 0.0, ay = 0.0;
      for (int64_t j = 0; j < margin; ++j) {       // np.mean(axis=0): sequential row adds
        ax = f64add(ax, base[j * 136 + stable[k] * 2 + 0]);
        ay = f64add(ay, base[j * 136 + stable[k] * 2 + 1]);
      }
      src[k][0] = f64div(ax, (double)margin);
      src[k][1] = f64div(ay, (double)margin);
      dst[k][0] = mean_face[stable[k] * 2 + 0];
      dst[k][1] = mean_face[stable[k] * 2 + 1];
    }
    similarity_fit(src, dst, kNumStable, fwd);
    affine_inverse(fwd, inv);
  }
  // trans(cur_landmarks)[48:68] -> mean -> cut_patch origin
  const double* cur = lm + f * 136;
  double cx = 0.0, cy = 0.0;
  for (int k = 48; k < 68; ++k) {
    const double x = cur[2 * k], y = cur[2 * k + 1];
    cx = f64add(cx, f64add(f64add(f64mul(x, fwd[0]), f64mul(y, fwd[1])), fwd[2]));
    cy = f64add(cy, f64add(f64add(f64mul(x, fwd[3]), f64mul(y, fwd[4])), fwd[5]));
  }
  cx = f64div(cx, 20.0);
  cy = f64div(cy, 20.0);
  int r0, c0;
  crop_origin(cx, cy, roi / 2, roi / 2, std_size, std_size, &r0, &c0);
  FrameXform o;
#pragma unroll
  for (int k = 0; k < 6; ++k) o.inv[k] = inv[k];
  o.r0 = r0; o.c0 = c0; o.pad[0] = 0; o.pad[1] = 0;
  xf[f] = o;
  if (crop_rc != nullptr) { crop_rc[2 * f] = r0; crop_rc[2 * f + 1] = c0; }
  if (tforms_out != nullptr) {
    double* to = tforms_out + f * 18;
    to[0] = fwd[0]; to[1] = fwd[1]; to[2] = fwd[2]; to[3] = fwd[3]; to[4] = fwd[4]; to[5] = fwd[5];
    to[6] = 0.0; to[7] = 0.0; to[8] = 1.0;
    to[9] = inv[0]; to[10] = inv[1]; to[11] = inv[2]; to[12] = inv[3]; to[13] = inv[4]; to[14] = inv[5];
    to[15] = 0.0; to[16] = 0.0; to[17] = 1.0;
  }
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

// Each warp iteration converts 512 px: 96 coalesced 16-byte loads (3 per lane) into the
// warp's 1536-byte shared slab, then every lane reads back its own 48 contiguous bytes
// (stride 48 B = 12 banks: conflict-free for 128-bit accesses) and stores 16 gray bytes.
__global__ void __launch_bounds__(kGrayThreads)
gray_vec_kernel(const uint4* __restrict__ bgr, int64_t ngroups /* of 512 px */,
                uint4* __restrict__ gray) {
  __shared__ uint4 slab[kGrayWarps][96];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * kGrayWarps + wid;
  const int64_t nwarps = (int64_t)gridDim.x * kGrayWarps;
  for (int64_t g = warp_global; g < ngroups; g += nwarps) {
    const uint4* src = bgr + g * 96;
    const uint4 a = ldg_stream(src + lane);
    const uint4 b = ldg_stream(src + lane + 32);
    const uint4 c = ldg_stream(src + lane + 64);
    slab[wid][lane] = a;
    slab[wid][lane + 32] = b;
    slab[wid][lane + 64] = c;
    __syncwarp();
    uint32_t w[12];
    const uint4 q0 = slab[wid][3 * lane], q1 = slab[wid][3 * lane + 1], q2 = slab[wid][3 * lane + 2];
    __syncwarp();
    w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
    w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
    w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
    stg_stream(gray + g * 32 + lane, gray16(w));
  }
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

// ------------------------------------------------------------------ V4/V5 warp + V7 + V8
constexpr int kWarpThreads = 256;

__device__ __forceinline__ void fill_luts(double* lut255, float* lutn, float mean, float stdv) {
  for (int k = threadIdx.x; k < 256; k += blockDim.x) {
    lut255[k] = f64div((double)k, 255.0);                                   // img_as_float
    if (lutn) lutn[k] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)k, 255.0f), mean), stdv);
  }
}

// One CTA per frame.  SRC_BGR: taps are converted from the BGR frame on the fly (used when
// the caller does not want the gray frames materialised); otherwise taps come from the gray
// frame (just written by gray_vec_kernel, typically still in L2).
template <bool SRC_BGR>
__global__ void __launch_bounds__(kWarpThreads)
warp_kernel(const uint8_t* __restrict__ src, int H, int W, const FrameXform* __restrict__ xf,
            int roi, int crop, float mean, float stdv, uint8_t* __restrict__ lip_u8,
            float* __restrict__ lip_f32) {
  __shared__ double lut255[256];
  __shared__ float lutn[256];
  fill_luts(lut255, lutn, mean, stdv);
  const int64_t f = blockIdx.x;
  const FrameXform x = xf[f];
  __syncthreads();
  const int off = (roi - crop) / 2;
  // evaluate the whole ROI only when the u8 ROI is wanted, else just the centre crop
  const int lo = lip_u8 ? 0 : off, span = lip_u8 ? roi : crop;
  const uint8_t* img = src + (size_t)f * H * W * (SRC_BGR ? 3 : 1);
  auto tap = [&](int r, int c) -> uint32_t {
    if (SRC_BGR) {
      const uint8_t* p = img + ((size_t)r * W + c) * 3;
      return gray_from_bgr(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
    return __ldg(img + (size_t)r * W + c);
  };
  for (int idx = threadIdx.x; idx < span * span; idx += kWarpThreads) {
    const int pr = lo + idx / span, pc = lo + idx % span;   // position inside the ROI
    uint8_t v = 0;
    if (x.r0 >= 0) {
      const double tfr = (double)(x.r0 + pr), tfc = (double)(x.c0 + pc);
      // _transform_affine: x_ = M0*x + M1*y + M2 ; y_ = M3*x + M4*y + M5
      const double sc = f64add(f64add(f64mul(x.inv[0], tfc), f64mul(x.inv[1], tfr)), x.inv[2]);
      const double sr = f64add(f64add(f64mul(x.inv[3], tfc), f64mul(x.inv[4], tfr)), x.inv[5]);
      v = bilinear_u8(sr, sc, H, W, lut255, tap);
    }
    if (lip_u8) lip_u8[(size_t)f * roi * roi + (size_t)pr * roi + pc] = v;
    if (lip_f32) {
      const int cr = pr - off, cc = pc - off;
      if (cr >= 0 && cr < crop && cc >= 0 && cc < crop)
        lip_f32[(size_t)f * crop * crop + (size_t)cr * crop + cc] = lutn[v];
    }
  }
}

// Full-frame warp with a caller-supplied inverse matrix (affine or projective).
__global__ void __launch_bounds__(256)
warp_full_kernel(const uint8_t* __restrict__ gray, int H, int W, const double* __restrict__ M,
                 int out_h, int out_w, uint8_t* __restrict__ out) {
  __shared__ double lut255[256];
  fill_luts(lut255, nullptr, 0.f, 1.f);
  __syncthreads();
  const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5];
  const double m6 = M[6], m7 = M[7], m8 = M[8];
  const bool affine = (m6 == 0.0) && (m7 == 0.0) && (m8 == 1.0);
  auto tap = [&](int r, int c) -> uint32_t { return __ldg(gray + (size_t)r * W + c); };
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
    out[idx] = bilinear_u8(sr, sc, H, W, lut255, tap);
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

}  // namespace avfe

// ====================================================================== C ABI
using namespace avfe;

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
  // filled landmarks [N,68,2] f64 + one FrameXform per frame
  return (size_t)N * 136 * sizeof(double) + (size_t)N * sizeof(FrameXform) + 64;
}

extern "C" int avfe_lip_roi_batch(const uint8_t* frames, int channels, int64_t N, int H, int W,
                                  const int64_t* clip_offsets, int64_t n_clips,
                                  const double* landmarks, const uint8_t* lm_valid,
                                  const double* mean_face, const double* tforms_in,
                                  int std_size, int roi, int crop, int window, float mean,
                                  float std, uint8_t* gray_out, uint8_t* lip_u8, float* lip_f32,
                                  int32_t* crop_rc, double* tforms, void* workspace,
                                  size_t workspace_bytes, avfe_stream_t stream) {
  if (N < 0 || n_clips < 0 || H <= 0 || W <= 0) return AVFE_ERR_INVALID_ARG;
  if (channels != 1 && channels != 3) return AVFE_ERR_INVALID_ARG;
  if (roi <= 0 || (roi & 1) || crop <= 0 || crop > roi || ((roi - crop) & 1) || window <= 0 ||
      std_size < roi)
    return AVFE_ERR_INVALID_ARG;
  if (N == 0 || n_clips == 0) return AVFE_OK;
  if (!frames || !clip_offsets || !landmarks || !mean_face) return AVFE_ERR_INVALID_ARG;
  if (gray_out && channels != 3) return AVFE_ERR_INVALID_ARG;
  if (N > 0x7fffffffLL) return AVFE_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < avfe_lip_workspace_bytes(N) || !aligned16(workspace))
    return AVFE_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  double* lm_filled = static_cast<double*>(workspace);
  FrameXform* xf = reinterpret_cast<FrameXform*>(lm_filled + (size_t)N * 136);
  const double* lm = landmarks;
  if (lm_valid != nullptr) {
    lm_fill_kernel<<<(unsigned)N, 160, 0, s>>>(landmarks, lm_valid, clip_offsets, n_clips, N,
                                               lm_filled);
    count_launch();
    lm = lm_filled;
  }
  tform_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(lm, clip_offsets, n_clips, N, mean_face,
                                                          tforms_in, std_size, roi, window, xf,
                                                          crop_rc, tforms);
  count_launch();
  if (gray_out != nullptr) {
    int rc = launch_gray(frames, N * (int64_t)H * W, gray_out, s);
    if (rc != AVFE_OK) return rc;
  }
  if (lip_u8 != nullptr || lip_f32 != nullptr) {
    if (channels == 3 && gray_out == nullptr) {
      warp_kernel<true><<<(unsigned)N, kWarpThreads, 0, s>>>(frames, H, W, xf, roi, crop, mean, std,
                                                            lip_u8, lip_f32);
    } else {
      const uint8_t* g = (channels == 3) ? gray_out : frames;
      warp_kernel<false><<<(unsigned)N, kWarpThreads, 0, s>>>(g, H, W, xf, roi, crop, mean, std,
                                                             lip_u8, lip_f32);
    }
    count_launch();
  }
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
