// TEST INFRASTRUCTURE ONLY.  Compiles person_capture_b200/csrc/pcb_cvmath.h (the arithmetic the
// CUDA kernels execute per pixel) for the host so that tests can pin it against real cv2 calls
// without a GPU.  Never loaded by the product package.
#include "../../person_capture_b200/csrc/pcb_cvmath.h"

extern "C" {

void hs_resize(const uint8_t* src, int h, int w, int rot, int pad, uint8_t* dst, int dh, int dw, int inter_area) {
  PcbView v = pcb_make_view(src, h, w, rot, pad);
  PcbResizePlan p = pcb_resize_plan(v.vh, v.vw, dh, dw, inter_area != 0);
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) pcb_resize_px(v, p, y, x, dh, dw, dst + ((long long)y * dw + x) * 3);
}

void hs_resize_factor(const uint8_t* src, int h, int w, uint8_t* dst, double fx, double fy, int inter_area) {
  PcbView v = pcb_make_view(src, h, w, 0, 0);
  const int dh = pcb_cvround_d((double)h * fy), dw = pcb_cvround_d((double)w * fx);
  PcbResizePlan p = pcb_resize_plan(h, w, dh, dw, inter_area != 0, fx, fy);
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) pcb_resize_px(v, p, y, x, dh, dw, dst + ((long long)y * dw + x) * 3);
}

void hs_view(const uint8_t* src, int h, int w, int rot, int pad, uint8_t* dst) {
  PcbView v = pcb_make_view(src, h, w, rot, pad);
  for (int y = 0; y < v.vh; ++y)
    for (int x = 0; x < v.vw; ++x) {
      const uint8_t* p = pcb_view_px(v, y, x);
      for (int c = 0; c < 3; ++c) dst[((long long)y * v.vw + x) * 3 + c] = p[c];
    }
}

void hs_warp(const uint8_t* src, int h, int w, long long row_stride, const double* M, uint8_t* dst, int dh, int dw) {
  PcbWarp c = pcb_invert_affine(M);
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) pcb_warp_px(src, row_stride, h, w, c, y, x, dst + ((long long)y * dw + x) * 3);
}

int hs_lmeds(const float* src, const float* dst, int count, double* M) { return pcb_lmeds_similarity(src, dst, count, M) ? 1 : 0; }
int hs_canon(const float* pts, float* out) { return pcb_canon_5pts(pts, out) ? 1 : 0; }
void hs_gray(const uint8_t* bgr, int n, uint8_t* out) { for (int i = 0; i < n; ++i) out[i] = pcb_gray(bgr + 3 * i); }
}
