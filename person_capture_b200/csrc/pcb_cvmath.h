// Exact integer / fixed-point arithmetic of the OpenCV uint8 kernels on the path, written once
// as host+device inline functions: the CUDA kernels (preproc.cu, align.cu) call them per output
// pixel, and tests/hostsim compiles the same header with g++ to pin the arithmetic against the
// real cv2 calls on the CPU (test infrastructure only -- the product never runs this on the host).
//
// What each function reproduces (OpenCV 4.x sources; reference call sites in brackets):
//   pcb_lin_coef / pcb_resize_linear_px  cv::resize INTER_LINEAR 8U  [InsightFace SCRFD.detect letterbox;
//                                         face_embedder.py:1285,1472,2264]
//   pcb_resize_area_*                     cv::resize INTER_AREA 8U    [gui_app.py:1505-1507; face_embedder.py:1472]
//   pcb_invert_affine / pcb_warp_px       cv::warpAffine INTER_LINEAR|BORDER_REFLECT [face_embedder.py:1473,1630]
//   pcb_lmeds_similarity                  cv::estimateAffinePartial2D(method=LMEDS)  [face_embedder.py:1466]
//   pcb_gray                              cv::cvtColor BGR2GRAY       [face_embedder.py:1275]
// Floating-point steps use explicit non-fused operations so results do not depend on FMA
// contraction (OpenCV's x86 baseline build does not fuse).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PCB_HD __host__ __device__ __forceinline__
#else
#define PCB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PCB_FMUL(a, b) __fmul_rn((a), (b))
#define PCB_FADD(a, b) __fadd_rn((a), (b))
#define PCB_FSUB(a, b) __fsub_rn((a), (b))
#define PCB_DMUL(a, b) __dmul_rn((a), (b))
#define PCB_DADD(a, b) __dadd_rn((a), (b))
#define PCB_DSUB(a, b) __dsub_rn((a), (b))
#define PCB_DDIV(a, b) __ddiv_rn((a), (b))
#else
// host build: compile with -ffp-contract=off
#define PCB_FMUL(a, b) ((float)(a) * (float)(b))
#define PCB_FADD(a, b) ((float)(a) + (float)(b))
#define PCB_FSUB(a, b) ((float)(a) - (float)(b))
#define PCB_DMUL(a, b) ((double)(a) * (double)(b))
#define PCB_DADD(a, b) ((double)(a) + (double)(b))
#define PCB_DSUB(a, b) ((double)(a) - (double)(b))
#define PCB_DDIV(a, b) ((double)(a) / (double)(b))
#endif

PCB_HD int pcb_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
PCB_HD int pcb_iminf(int a, int b) { return a < b ? a : b; }
PCB_HD int pcb_imaxf(int a, int b) { return a > b ? a : b; }
PCB_HD int pcb_cvround_f(float v) { return (int)lrintf(v); }    // cvRound: round half to even
PCB_HD int pcb_cvround_d(double v) { return (int)lrint(v); }
PCB_HD uint8_t pcb_sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ---------------------------------------------------------------------------------------
// Source view: a frame seen through cv2.rotate(deg) and cv2.copyMakeBorder(pad, REPLICATE),
// evaluated by index arithmetic instead of materialising the rotated / padded image
// (face_embedder.py:2165-2169, 2292-2294, 2394).
// ---------------------------------------------------------------------------------------
struct PcbView {
  const uint8_t* base;   // frame [h][w][3]
 int h, w;              // stored frame dims
  int ld;                // row pitch in pixels (>= w; a crop view keeps the frame's pitch)
  int rot;               // 0 | 90 | 180 | 270 (cv2.ROTATE_90_CLOCKWISE = 90, COUNTERCLOCKWISE = 270)
  int pad;               // replicate border
  int vh, vw;            // dims of the virtual image = rotated dims + 2*pad
};

PCB_HD PcbView pcb_make_view(const uint8_t* base, int h, int w, int rot, int pad) {
  PcbView v;
  v.base = base; v.h = h; v.w = w; v.ld = w; v.rot = rot; v.pad = pad;
  const int rh = (rot == 90 || rot == 270) ? w : h;
  const int rw = (rot == 90 || rot == 270) ? h : w;
  v.vh = rh + 2 * pad;
  v.vw = rw + 2 * pad;
  return v;
}

PCB_HD const uint8_t* pcb_view_px(const PcbView& v, int y, int x) {
  const int rh = v.vh - 2 * v.pad, rw = v.vw - 2 * v.pad;
  const int ry = pcb_clampi(y - v.pad, 0, rh - 1), rx = pcb_clampi(x - v.pad, 0, rw - 1);
  int sy, sx;
  if (v.rot == 90) { sy = v.h - 1 - rx; sx = ry; }
  else if (v.rot == 180) { sy = v.h - 1 - ry; sx = v.w - 1 - rx; }
  else if (v.rot == 270) { sy = rx; sx = v.w - 1 - ry; }
  else { sy = ry; sx = rx; }
  return v.base + ((long long)sy * v.ld + sx) * 3;
}

// ---------------------------------------------------------------------------------------
// INTER_LINEAR (8U fixed point: 11-bit coefficients, two-stage shift)
// ---------------------------------------------------------------------------------------
struct PcbLinCoef { int s; int a0, a1; };

// `inv` is OpenCV's inv_scale: dst/src when the caller gave dsize, fx/fy itself when it gave factors.
PCB_HD double pcb_inv_scale(int src, int dst) { return PCB_DDIV((double)dst, (double)src); }

PCB_HD PcbLinCoef pcb_lin_coef(int d, int src, int dst, bool horizontal, bool area_mode, double inv) {
  const double scale = PCB_DDIV(1.0, inv);
  int s;
  float f;
  if (area_mode) {   // INTER_AREA requested on an up-scaled axis: bilinear with area-style taps
    s = (int)floor(PCB_DMUL((double)d, scale));
    f = (float)PCB_DSUB((double)(d + 1), PCB_DMUL((double)(s + 1), inv));
    f = f <= 0.f ? 0.f : PCB_FSUB(f, floorf(f));
  } else {
    f = (float)PCB_DSUB(PCB_DMUL(PCB_DADD((double)d, 0.5), scale), 0.5);
    s = (int)floorf(f);
    f = PCB_FSUB(f, (float)s);
  }
  if (horizontal) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  PcbLinCoef c;
  c.s = s;
  c.a0 = pcb_cvround_f(PCB_FMUL(PCB_FSUB(1.f, f), 2048.f));
  c.a1 = pcb_cvround_f(PCB_FMUL(f, 2048.f));
  return c;
}

// one output pixel of cv::resize(view, (dw, dh), INTER_LINEAR) (or the INTER_AREA-on-upscale variant)
PCB_HD void pcb_resize_linear_px(const PcbView& v, int dy, int dx, int dh, int dw, bool area_mode, double inv_x, double inv_y,
                                 uint8_t out[3]) {
  const PcbLinCoef cx = pcb_lin_coef(dx, v.vw, dw, true, area_mode, inv_x);
  const PcbLinCoef cy = pcb_lin_coef(dy, v.vh, dh, false, area_mode, inv_y);
  const int x0 = cx.s, x1 = pcb_iminf(cx.s + 1, v.vw - 1);
  const int y0 = pcb_clampi(cy.s, 0, v.vh - 1), y1 = pcb_clampi(cy.s + 1, 0, v.vh - 1);
  const uint8_t* p00 = pcb_view_px(v, y0, x0);
  const uint8_t* p01 = pcb_view_px(v, y0, x1);
  const uint8_t* p10 = pcb_view_px(v, y1, x0);
  const uint8_t* p11 = pcb_view_px(v, y1, x1);
  for (int c = 0; c < 3; ++c) {
    const int h0 = p00[c] * cx.a0 + p01[c] * cx.a1;
    const int h1 = p10[c] * cx.a0 + p11[c] * cx.a1;
    const int r = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
    out[c] = pcb_sat_u8(r);
  }
}

// exact integer-ratio INTER_AREA (resizeAreaFast_): (2,2) rounds half up, others round half to even
PCB_HD void pcb_resize_area_int_px(const PcbView& v, int dy, int dx, int isy, int isx, uint8_t out[3]) {
  int s[3] = {0, 0, 0};
  for (int yy = 0; yy < isy; ++yy)
    for (int xx = 0; xx < isx; ++xx) {
      const uint8_t* p = pcb_view_px(v, dy * isy + yy, dx * isx + xx);
      s[0] += p[0]; s[1] += p[1]; s[2] += p[2];
    }
  if (isx == 2 && isy == 2) {
    for (int c = 0; c < 3; ++c) out[c] = (uint8_t)((s[c] + 2) >> 2);
  } else {
    const float sc = 1.f / (float)(isx * isy);
    for (int c = 0; c < 3; ++c) out[c] = pcb_sat_u8(pcb_cvround_f(PCB_FMUL((float)s[c], sc)));
  }
}

// fractional INTER_AREA (resizeArea_): float32 tap table, horizontal then vertical accumulation in
// OpenCV's order.  Taps of destination index d: [s1-1 partial] [s1..s2) full [s2 partial].
struct PcbAreaTaps { int s1, s2; float wl, wm, wr; bool has_l, has_r; };

PCB_HD PcbAreaTaps pcb_area_taps(int d, int src, int dst, double inv) {
  const double scale = PCB_DDIV(1.0, inv);
  const double fs1 = PCB_DMUL((double)d, scale);
  const double fs2 = PCB_DADD(fs1, scale);
  const double rem = PCB_DSUB((double)src, fs1);
  const double cell = scale < rem ? scale : rem;
  int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
  s2 = pcb_iminf(s2, src - 1);
  s1 = pcb_iminf(s1, s2);
  PcbAreaTaps t;
  t.s1 = s1; t.s2 = s2;
  t.has_l = PCB_DSUB((double)s1, fs1) > 1e-3;
  t.wl = (float)PCB_DDIV(PCB_DSUB((double)s1, fs1), cell);
  t.wm = (float)PCB_DDIV(1.0, cell);
  const double r = PCB_DSUB(fs2, (double)s2);
  t.has_r = r > 1e-3;
  double rr = r < 1.0 ? r : 1.0;
  rr = rr < cell ? rr : cell;
  t.wr = (float)PCB_DDIV(rr, cell);
  return t;
}

PCB_HD void pcb_area_hrow(const PcbView& v, int sy, const PcbAreaTaps& tx, float buf[3]) {
  buf[0] = buf[1] = buf[2] = 0.f;
  if (tx.has_l) {
    const uint8_t* p = pcb_view_px(v, sy, tx.s1 - 1);
    for (int c = 0; c < 3; ++c) buf[c] = PCB_FADD(buf[c], PCB_FMUL((float)p[c], tx.wl));
  }
  for (int sx = tx.s1; sx < tx.s2; ++sx) {
    const uint8_t* p = pcb_view_px(v, sy, sx);
    for (int c = 0; c < 3; ++c) buf[c] = PCB_FADD(buf[c], PCB_FMUL((float)p[c], tx.wm));
  }
  if (tx.has_r) {
    const uint8_t* p = pcb_view_px(v, sy, tx.s2);
    for (int c = 0; c < 3; ++c) buf[c] = PCB_FADD(buf[c], PCB_FMUL((float)p[c], tx.wr));
  }
}

PCB_HD void pcb_resize_area_frac_px(const PcbView& v, int dy, int dx, int dh, int dw, double inv_x, double inv_y, uint8_t out[3]) {
  const PcbAreaTaps tx = pcb_area_taps(dx, v.vw, dw, inv_x);
  const PcbAreaTaps ty = pcb_area_taps(dy, v.vh, dh, inv_y);
  float sum[3] = {0.f, 0.f, 0.f};
  float buf[3];
  if (ty.has_l) {
    pcb_area_hrow(v, ty.s1 - 1, tx, buf);
    for (int c = 0; c < 3; ++c) sum[c] = PCB_FADD(sum[c], PCB_FMUL(ty.wl, buf[c]));
  }
  for (int sy = ty.s1; sy < ty.s2; ++sy) {
    pcb_area_hrow(v, sy, tx, buf);
    for (int c = 0; c < 3; ++c) sum[c] = PCB_FADD(sum[c], PCB_FMUL(ty.wm, buf[c]));
  }
  if (ty.has_r) {
    pcb_area_hrow(v, ty.s2, tx, buf);
    for (int c = 0; c < 3; ++c) sum[c] = PCB_FADD(sum[c], PCB_FMUL(ty.wr, buf[c]));
  }
  for (int c = 0; c < 3; ++c) out[c] = pcb_sat_u8(pcb_cvround_f(sum[c]));
}

// Dispatch of cv::resize for 8UC3, as OpenCV selects the kernel.
enum { PCB_RS_LINEAR = 0, PCB_RS_AREA_INT = 1, PCB_RS_AREA_FRAC = 2, PCB_RS_LINEAR_AREAMODE = 3 };
struct PcbResizePlan { int mode, isx, isy; double inv_x, inv_y; };

// fx, fy > 0: cv2.resize(src, None, fx=, fy=) form (inv_scale = the factor itself); else dsize form.
PCB_HD PcbResizePlan pcb_resize_plan(int sh, int sw, int dh, int dw, bool inter_area, double fx = 0.0, double fy = 0.0) {
  PcbResizePlan p;
  p.inv_x = fx > 0.0 ? fx : pcb_inv_scale(sw, dw);
  p.inv_y = fy > 0.0 ? fy : pcb_inv_scale(sh, dh);
  const double scx = PCB_DDIV(1.0, p.inv_x);
  const double scy = PCB_DDIV(1.0, p.inv_y);
  const int isx = pcb_cvround_d(scx), isy = pcb_cvround_d(scy);
  const bool fast = fabs(scx - isx) < 2.220446049250313e-16 && fabs(scy - isy) < 2.220446049250313e-16;
  p.isx = isx; p.isy = isy;
  if (!inter_area && fast && isx == 2 && isy == 2) inter_area = true;   // LINEAR with exact 2x == AREA
  if (!inter_area) { p.mode = PCB_RS_LINEAR; return p; }
  if (scx >= 1.0 && scy >= 1.0) { p.mode = fast ? PCB_RS_AREA_INT : PCB_RS_AREA_FRAC; return p; }
  p.mode = PCB_RS_LINEAR_AREAMODE;
  return p;
}

PCB_HD void pcb_resize_px(const PcbView& v, const PcbResizePlan& p, int dy, int dx, int dh, int dw, uint8_t out[3]) {
  if (p.mode == PCB_RS_AREA_INT) pcb_resize_area_int_px(v, dy, dx, p.isy, p.isx, out);
  else if (p.mode == PCB_RS_AREA_FRAC) pcb_resize_area_frac_px(v, dy, dx, dh, dw, p.inv_x, p.inv_y, out);
  else pcb_resize_linear_px(v, dy, dx, dh, dw, p.mode == PCB_RS_LINEAR_AREAMODE, p.inv_x, p.inv_y, out);
}

// ---------------------------------------------------------------------------------------
// warpAffine: 1/32-pixel fixed-point coordinates, 15-bit bilinear weights, BORDER_REFLECT
// ---------------------------------------------------------------------------------------
struct PcbWarp { double m00, m01, m02, m10, m11, m12; };

PCB_HD PcbWarp pcb_invert_affine(const double M[6]) {
  double D = PCB_DSUB(PCB_DMUL(M[0], M[4]), PCB_DMUL(M[1], M[3]));
  D = D != 0.0 ? PCB_DDIV(1.0, D) : 0.0;
  PcbWarp w;
  w.m00 = PCB_DMUL(M[4], D);
  w.m11 = PCB_DMUL(M[0], D);
  w.m01 = PCB_DMUL(M[1], -D);
  w.m10 = PCB_DMUL(M[3], -D);
  w.m02 = PCB_DSUB(PCB_DMUL(-w.m00, M[2]), PCB_DMUL(w.m01, M[5]));
  w.m12 = PCB_DSUB(PCB_DMUL(-w.m10, M[2]), PCB_DMUL(w.m11, M[5]));
  return w;
}

PCB_HD int pcb_reflect(int p, int n) {   // cv::borderInterpolate(p, n, BORDER_REFLECT)
  if ((unsigned)p < (unsigned)n) return p;
  if (n == 1) return 0;
  do {
    if (p < 0) p = -p - 1;
    else p = n - 1 - (p - n);
  } while ((unsigned)p >= (unsigned)n);
  return p;
}

// src: [h][w][3] with `row_stride` bytes between rows (a crop view of a frame)
PCB_HD void pcb_warp_px(const uint8_t* src, long long row_stride, int h, int w, const PcbWarp& c, int y, int x, uint8_t out[3]) {
  const int adelta = pcb_cvround_d(PCB_DMUL(PCB_DMUL(c.m00, (double)x), 1024.0));
  const int bdelta = pcb_cvround_d(PCB_DMUL(PCB_DMUL(c.m10, (double)x), 1024.0));
  const int X0 = pcb_cvround_d(PCB_DMUL(PCB_DADD(PCB_DMUL(c.m01, (double)y), c.m02), 1024.0)) + 16;
  const int Y0 = pcb_cvround_d(PCB_DMUL(PCB_DADD(PCB_DMUL(c.m11, (double)y), c.m12), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = pcb_clampi(X >> 5, -32768, 32767), sy = pcb_clampi(Y >> 5, -32768, 32767);
  const int fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  const int x0 = pcb_reflect(sx, w), x1 = pcb_reflect(sx + 1, w);
  const int y0 = pcb_reflect(sy, h), y1 = pcb_reflect(sy + 1, h);
  const uint8_t* r0 = src + (long long)y0 * row_stride;
  const uint8_t* r1 = src + (long long)y1 * row_stride;
  for (int ch = 0; ch < 3; ++ch) {
    const int v = r0[x0 * 3 + ch] * w00 + r0[x1 * 3 + ch] * w01 + r1[x0 * 3 + ch] * w10 + r1[x1 * 3 + ch] * w11;
    out[ch] = pcb_sat_u8((v + 16384) >> 15);
  }
}

PCB_HD uint8_t pcb_gray(const uint8_t* bgr) {   // cv2 4.13: 15-bit BY/GY/RY coefficients
  return (uint8_t)((bgr[0] * 3735 + bgr[1] * 19235 + bgr[2] * 9798 + 16384) >> 15);
}

// ---------------------------------------------------------------------------------------
// estimateAffinePartial2D(src, dst, method=LMEDS) for `count` (5 or 3) points.
// RNG(2^64-1) is re-seeded per call, 13 two-point hypotheses, median of float32 squared
// errors, sigma rule for inliers, then the least-squares similarity over the inliers
// (OpenCV refines with <=10 LM steps to the same optimum; |dM| ~ 1e-13).
// Returns false when OpenCV would return an empty matrix.
// ---------------------------------------------------------------------------------------
PCB_HD uint32_t pcb_rng_next(uint64_t& st) {
  st = (uint64_t)(uint32_t)st * 4164903690ULL + (st >> 32);
  return (uint32_t)st;
}

PCB_HD void pcb_errors(const double Md[6], const float* src, const float* dst, int count, float* err) {
  const float F0 = (float)Md[0], F1 = (float)Md[1], F2 = (float)Md[2], F3 = (float)Md[3], F4 = (float)Md[4], F5 = (float)Md[5];
  for (int i = 0; i < count; ++i) {
    const float fx = src[2 * i], fy = src[2 * i + 1];
    const float a = PCB_FSUB(PCB_FADD(PCB_FADD(PCB_FMUL(F0, fx), PCB_FMUL(F1, fy)), F2), dst[2 * i]);
    const float b = PCB_FSUB(PCB_FADD(PCB_FADD(PCB_FMUL(F3, fx), PCB_FMUL(F4, fy)), F5), dst[2 * i + 1]);
    err[i] = PCB_FADD(PCB_FMUL(a, a), PCB_FMUL(b, b));
  }
}

PCB_HD bool pcb_lmeds_similarity(const float* src, const float* dst, int count, double M[6]) {
  uint64_t st = 0xFFFFFFFFFFFFFFFFULL;
  double best[6] = {0, 0, 0, 0, 0, 0};
  double best_med = 1.7976931348623157e308;
  bool have = false;
  for (int it = 0; it < 13; ++it) {
    const int i0 = (int)(pcb_rng_next(st) % (uint32_t)count);
    int i1;
    do { i1 = (int)(pcb_rng_next(st) % (uint32_t)count); } while (i1 == i0);
    const double x1 = src[2 * i0], y1 = src[2 * i0 + 1], x2 = src[2 * i1], y2 = src[2 * i1 + 1];
    const double X1 = dst[2 * i0], Y1 = dst[2 * i0 + 1], X2 = dst[2 * i1], Y2 = dst[2 * i1 + 1];
    const double dx = x1 - x2, dy = y1 - y2, DX = X1 - X2, DY = Y1 - Y2;
    const double d = 1.0 / (dx * dx + dy * dy);
    const double S0 = d * (DX * dx + DY * dy);
    const double S1 = d * (DY * dx - DX * dy);
    const double cr = x1 * y2 - x2 * y1;
    const double S2 = d * (DY * cr - (X1 * y2 - X2 * y1) * dy - (X1 * x2 - X2 * x1) * dx);
    const double S3 = d * (-DX * cr - (Y1 * x2 - Y2 * x1) * dx - (Y1 * y2 - Y2 * y1) * dy);
    const double Mh[6] = {S0, -S1, S2, S1, S0, S3};
    float err[5];
    pcb_errors(Mh, src, dst, count, err);
    float e[5];
    for (int i = 0; i < count; ++i) e[i] = err[i];
    for (int i = 1; i < count; ++i) {   // insertion sort
      float key = e[i];
      int j = i - 1;
      while (j >= 0 && e[j] > key) { e[j + 1] = e[j]; --j; }
      e[j + 1] = key;
    }
    const double med = (count & 1) ? (double)e[count / 2] : ((double)e[count / 2 - 1] + (double)e[count / 2]) * 0.5;
    if (med < best_med) {
      best_med = med;
      for (int k = 0; k < 6; ++k) best[k] = Mh[k];
      have = true;
    }
  }
  if (!have) return false;
  double sigma = 2.5 * 1.4826 * (1.0 + 5.0 / (double)(count - 2)) * sqrt(best_med);
  if (!(sigma > 0.001)) sigma = 0.001;
  const float thr = (float)(sigma * sigma);
  float err[5];
  pcb_errors(best, src, dst, count, err);
  int n = 0;
  double ms0 = 0, ms1 = 0, md0 = 0, md1 = 0;
  bool in[5];
  for (int i = 0; i < count; ++i) {
    in[i] = err[i] <= thr;
    if (in[i]) { ++n; ms0 += src[2 * i]; ms1 += src[2 * i + 1]; md0 += dst[2 * i]; md1 += dst[2 * i + 1]; }
  }
  if (n < 2) return false;
  ms0 /= n; ms1 /= n; md0 /= n; md1 /= n;
  double den = 0, na = 0, nb = 0;
  for (int i = 0; i < count; ++i) {
    if (!in[i]) continue;
    const double sx = src[2 * i] - ms0, sy = src[2 * i + 1] - ms1, tx = dst[2 * i] - md0, ty = dst[2 * i + 1] - md1;
    den += sx * sx + sy * sy;
    na += sx * tx + sy * ty;
    nb += sx * ty - sy * tx;
  }
  const double a = na / den, b = nb / den;
  M[0] = a; M[1] = -b; M[2] = md0 - (a * ms0 - b * ms1);
  M[3] = b; M[4] = a;  M[5] = md1 - (b * ms0 + a * ms1);
  for (int k = 0; k < 6; ++k) if (!isfinite(M[k])) return false;
  return true;
}

// _canon_5pts (face_embedder.py:1430-1463): stable sort by y, eyes/mouth by x, validity checks.
PCB_HD bool pcb_canon_5pts(const float* pts, float* out) {
  for (int i = 0; i < 10; ++i) if (!isfinite(pts[i])) return false;
  int ord[5] = {0, 1, 2, 3, 4};
  for (int i = 1; i < 5; ++i) {
    int k = ord[i], j = i - 1;
    while (j >= 0 && pts[2 * ord[j] + 1] > pts[2 * k + 1]) { ord[j + 1] = ord[j]; --j; }
    ord[j + 1] = k;
  }
  int e0 = ord[0], e1 = ord[1], nose = ord[2], m0 = ord[3], m1 = ord[4];
  if (pts[2 * e0] > pts[2 * e1]) { int t = e0; e0 = e1; e1 = t; }
  if (pts[2 * m0] > pts[2 * m1]) { int t = m0; m0 = m1; m1 = t; }
  if (!(pts[2 * e0] < pts[2 * e1] && pts[2 * m0] < pts[2 * m1])) return false;
  const float upper_eye = fmaxf(pts[2 * e0 + 1], pts[2 * e1 + 1]);
  const float lower_mouth = fminf(pts[2 * m0 + 1], pts[2 * m1 + 1]);
  if (!(pts[2 * nose + 1] > upper_eye && pts[2 * nose + 1] < lower_mouth)) return false;
  const int sel[5] = {e0, e1, nose, m0, m1};
  for (int i = 0; i < 5; ++i) { out[2 * i] = pts[2 * sel[i]]; out[2 * i + 1] = pts[2 * sel[i] + 1]; }
  return true;
}
