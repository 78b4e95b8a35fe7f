// Scalar building blocks of the lip-ROI path, written once and compiled for the device
// (kernels in avfe_lip.cu) and for the host (tests/hostcheck, a TEST-ONLY harness that checks
// these codelets against the oracle without a GPU; it is not a product CPU path).
//
// All float64 arithmetic goes through dmul/dadd/dsub/ddiv so that nvcc cannot contract a
// multiply and an add into an FMA: the oracle (numpy / skimage's C) rounds each operation
// separately, and bit-exact uint8 output needs the same roundings.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AVFE_HD __host__ __device__ __forceinline__
#else
#define AVFE_HD inline
#endif

namespace avfe {

#if defined(__CUDA_ARCH__)
AVFE_HD double f64mul(double a, double b) { return __dmul_rn(a, b); }
AVFE_HD double f64add(double a, double b) { return __dadd_rn(a, b); }
AVFE_HD double f64sub(double a, double b) { return __dsub_rn(a, b); }
AVFE_HD double f64div(double a, double b) { return __ddiv_rn(a, b); }
#else
// host build is compiled with -ffp-contract=off
AVFE_HD double f64mul(double a, double b) { return a * b; }
AVFE_HD double f64add(double a, double b) { return a + b; }
AVFE_HD double f64sub(double a, double b) { return a - b; }
AVFE_HD double f64div(double a, double b) { return a / b; }
#endif

// cv2.cvtColor(BGR2GRAY), 15-bit fixed point (preprocess/video_process.py:214).
AVFE_HD uint32_t gray_from_bgr(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

constexpr int kNumLandmarks = 68;
constexpr int kNumStable = 5;

// Least-squares similarity (rotation + isotropic scale + translation) mapping src -> dst.
// This is skimage's _umeyama for dim == 2 in closed form: with A = dst_c^T src_c / n,
// E = A00 + A11, Hh = A10 - A01, the SVD-based rotation is [[E,-Hh],[Hh,E]] / hypot(E,Hh) and
// (S . d) = hypot(E,Hh) for either sign of det(A), so scale * R = [[E,-Hh],[Hh,E]] / var(src).
// fwd = {a, -b, tx, b, a, ty} (first two rows of the 3x3).  All-equal src (rank 0) -> NaN.
AVFE_HD void similarity_fit(const double (*src)[2], const double (*dst)[2], int n, double fwd[6]) {
  double sx = 0.0, sy = 0.0, dx = 0.0, dy = 0.0;
  for (int i = 0; i < n; ++i) {
    sx = f64add(sx, src[i][0]); sy = f64add(sy, src[i][1]);
    dx = f64add(dx, dst[i][0]); dy = f64add(dy, dst[i][1]);
  }
  const double dn = (double)n;
  sx = f64div(sx, dn); sy = f64div(sy, dn); dx = f64div(dx, dn); dy = f64div(dy, dn);
  double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0, vx = 0.0, vy = 0.0;
  for (int i = 0; i < n; ++i) {
    const double px = f64sub(src[i][0], sx), py = f64sub(src[i][1], sy);
    const double qx = f64sub(dst[i][0], dx), qy = f64sub(dst[i][1], dy);
    a00 = f64add(a00, f64mul(qx, px)); a01 = f64add(a01, f64mul(qx, py));
    a10 = f64add(a10, f64mul(qy, px)); a11 = f64add(a11, f64mul(qy, py));
    vx = f64add(vx, f64mul(px, px)); vy = f64add(vy, f64mul(py, py));
  }
  a00 = f64div(a00, dn); a01 = f64div(a01, dn); a10 = f64div(a10, dn); a11 = f64div(a11, dn);
  const double var = f64add(f64div(vx, dn), f64div(vy, dn));
  const double E = f64add(a00, a11), Hh = f64sub(a10, a01);
  double a = f64div(E, var), b = f64div(Hh, var);
  if (E == 0.0 && Hh == 0.0) { a = nan(""); b = nan(""); }  // rank 0: skimage returns nan * T
  fwd[0] = a; fwd[1] = -b; fwd[3] = b; fwd[4] = a;
  fwd[2] = f64sub(dx, f64sub(f64mul(a, sx), f64mul(b, sy)));
  fwd[5] = f64sub(dy, f64add(f64mul(b, sx), f64mul(a, sy)));
}

// Inverse of a 2-D affine map {m0,m1,m2; m3,m4,m5; 0,0,1}.
AVFE_HD void affine_inverse(const double f[6], double inv[6]) {
  const double det = f64sub(f64mul(f[0], f[4]), f64mul(f[1], f[3]));
  const double i00 = f64div(f[4], det), i01 = f64div(-f[1], det);
  const double i10 = f64div(-f[3], det), i11 = f64div(f[0], det);
  inv[0] = i00; inv[1] = i01; inv[3] = i10; inv[4] = i11;
  inv[2] = -f64add(f64mul(i00, f[2]), f64mul(i01, f[5]));
  inv[5] = -f64add(f64mul(i10, f[2]), f64mul(i11, f[5]));
}

// cut_patch's centre handling (utils/lips_cropping.py:141-162): clamp the centre so the patch
// fits (the reference's raise branches are unreachable after the clamps), then Python round()
// (half-to-even == rint) of the centre minus the half size.
AVFE_HD void crop_origin(double cx, double cy, int half_h, int half_w, int img_h, int img_w,
                         int* r0, int* c0) {
  if (!(isfinite(cx) && isfinite(cy))) { *r0 = -1; *c0 = -1; return; }
  if (cy - half_h < 0) cy = half_h;
  if (cx - half_w < 0) cx = half_w;
  if (cy + half_h > img_h) cy = img_h - half_h;
  if (cx + half_w > img_w) cx = img_w - half_w;
  *r0 = (int)(rint(cy) - (double)half_h);
  *c0 = (int)(rint(cx) - (double)half_w);
}

// One output pixel of skimage _warp_fast (order 1, mode 'constant', cval 0) on an
// img_as_float'ed uint8 image, then (*255).astype(uint8).
// `tap(r, c)` returns the source pixel as float64 k/255.0 (only called for in-frame taps).
template <typename TapFn>
AVFE_HD uint8_t bilinear_u8(double r, double c, int H, int W, TapFn tap) {
  const double fr = floor(r), fc = floor(c);
  // far outside the image every tap is cval; also keeps the int conversions in range
  if (!(fr >= -2.0 && fr <= (double)H + 1.0 && fc >= -2.0 && fc <= (double)W + 1.0)) return 0;
  const int minr = (int)fr, minc = (int)fc;
  const int maxr = (int)ceil(r), maxc = (int)ceil(c);
  const double dr = f64sub(r, fr), dc = f64sub(c, fc);
  const bool r0ok = (minr >= 0) && (minr < H), r1ok = (maxr >= 0) && (maxr < H);
  const bool c0ok = (minc >= 0) && (minc < W), c1ok = (maxc >= 0) && (maxc < W);
  const double tl = (r0ok && c0ok) ? tap(minr, minc) : 0.0;
  const double tr = (r0ok && c1ok) ? tap(minr, maxc) : 0.0;
  const double bl = (r1ok && c0ok) ? tap(maxr, minc) : 0.0;
  const double br = (r1ok && c1ok) ? tap(maxr, maxc) : 0.0;
  const double omc = f64sub(1.0, dc), omr = f64sub(1.0, dr);
  const double top = f64add(f64mul(omc, tl), f64mul(dc, tr));
  const double bot = f64add(f64mul(omc, bl), f64mul(dc, br));
  const double v = f64add(f64mul(omr, top), f64mul(dr, bot));
  // _clip_warp_output clips to [min(img.min(),0), max(img.max(),0)]; for a convex blend of
  // values k/255 this cannot change trunc(v*255) ("clip is a no-op"), so it is
  // not evaluated.
  return (uint8_t)(int)f64mul(v, 255.0);
}

}  // namespace avfe
