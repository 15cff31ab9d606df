// K0 / K1 and the ArcFace input kernel (HBM-bound byte work; no tensor cores).
//   K0  pcb_resize_area / pcb_resize_linear : cv2.resize on uint8 BGR frames, bit-exact
//       (gui_app.py:1505-1507 INTER_AREA pre-scan downscale; face_embedder.py:2263-2264 scale TTA).
//   K1  letterbox_kernel : InsightFace SCRFD.detect letterbox (aspect-preserving bilinear resize
//       anchored top-left, zero pad) + SCRFD.forward blobFromImage ((x-127.5)/128, BGR->RGB),
//       fused with the stem's im2col: it emits the fp16 "stem patch" tensor
//       [n][S/2+2][S/2+2][32] whose 27 channels are the 3x3 stride-2 neighbourhood of the
//       normalised image, so the 3->C stem convolution runs on the tensor cores as a 1x1 GEMM.
//       Optional rotation / replicate border are index maps (face_embedder.py:2165-2169, 2394).
//   chip_patch_kernel : _arcface_preprocess (face_embedder.py:1281-1288) + optional cv2.flip
//       (face_embedder.py:1297-1298) fused with the ArcFace stem im2col (stride 1).
#include "pcb_common.cuh"
#include "pcb_cvmath.h"

namespace {

// ------------------------------------------------------------------ K0 generic resize
__global__ void resize_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int h, int w, int dh, int dw,
                                 PcbResizePlan plan) {
  const long long total = (long long)n * dh * dw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % dw);
    const int y = (int)((idx / dw) % dh);
    const int img = (int)(idx / ((long long)dw * dh));
    PcbView v = pcb_make_view(src + (size_t)img * h * w * 3, h, w, 0, 0);
    uint8_t o[3];
    pcb_resize_px(v, plan, y, x, dh, dw, o);
    uint8_t* d = dst + idx * 3;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
  }
}

// exact 2x INTER_AREA, vectorised: one thread = 16 output pixels (2 x 96 B in, 48 B out)
__global__ void area2x_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n, int h, int w, int dh, int dw) {
  const int groups = dw / 16;
  const long long total = (long long)n * dh * groups;
  const int src_row16 = w * 3 / 16;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const int y = (int)((idx / groups) % dh);
    const int img = (int)(idx / ((long long)groups * dh));
    const uint4* r0 = src + ((size_t)img * h + 2 * y) * src_row16 + g * 6;
    const uint4* r1 = r0 + src_row16;
    uint4 a[6], b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { a[i] = __ldg(r0 + i); b[i] = __ldg(r1 + i); }
    const uint8_t* pa = (const uint8_t*)a;
    const uint8_t* pb = (const uint8_t*)b;
    uint4 o[3];
    uint8_t* po = (uint8_t*)o;
#pragma unroll
    for (int px = 0; px < 16; ++px)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        po[px * 3 + c] = (uint8_t)((pa[px * 6 + c] + pa[px * 6 + 3 + c] + pb[px * 6 + c] + pb[px * 6 + 3 + c] + 2) >> 2);
    uint4* d = dst + ((size_t)img * dh + y) * (dw * 3 / 16) + g * 3;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
  }
}

// ------------------------------------------------------------------ K1 letterbox
struct LinTab { int s; int a0; int a1; int pad_; };

__global__ void lin_table_kernel(LinTab* tab, int src, int dst, int horizontal) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dst) return;
  PcbLinCoef c = pcb_lin_coef(d, src, dst, horizontal != 0, false, pcb_inv_scale(src, dst));
  tab[d].s = c.s; tab[d].a0 = c.a0; tab[d].a1 = c.a1; tab[d].pad_ = 0;
}

struct LetterboxParams {
  const uint8_t* frames;
  int n, h, w, rot, pad;
  int S, new_h, new_w;
  int mode;                 // PCB_RS_LINEAR (tables) or any other plan (per-pixel generic)
  PcbResizePlan plan;
  const LinTab* xtab;
  const LinTab* ytab;
  __half* out;              // [n][S/2+2][S/2+2][32]
  uint8_t* det_img;         // optional [n][S][S][3]
};

constexpr int kLbTile = 16;                 // patch pixels per block edge
constexpr int kLbDet = 2 * kLbTile + 1;     // det pixels needed per edge (33)

// kPlain: no rotation, no replicate border, bilinear tables (the upright pass of every frame): source pixels are addressed directly
template <bool kPlain>
__global__ void __launch_bounds__(256) letterbox_kernel(const LetterboxParams p) {
  __shared__ __align__(4) uint8_t tile[kLbDet][kLbDet][4];   // [.][.][3] = 1 when the det pixel is inside [0,S)
  // (v - 127.5) / 128 as fp16, tabulated per byte value (exact: the float expression has 9 significant bits); entry 256 = the
  // zero of taps outside the detector image.  Replaces four conversions / multiplies per tap, 27 taps per output pixel.
  __shared__ __half lut[257];
  lut[threadIdx.x] = __float2half_rn(((float)threadIdx.x - 127.5f) * (1.0f / 128.0f));
  if (threadIdx.x == 0) lut[256] = __float2half_rn(0.f);
  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kLbTile, ox0 = blockIdx.x * kLbTile;
  const int dy0 = 2 * oy0 - 1, dx0 = 2 * ox0 - 1;
  const int half = p.S / 2;
  if ((dy0 >= p.new_h || dx0 >= p.new_w) && dy0 >= 0 && dx0 >= 0 && dy0 + kLbDet <= p.S && dx0 + kLbDet <= p.S && !p.det_img) {
    // the block lies in the zero padding of the letterbox (44 % of a 16:9 frame's square): every tap is the normalised zero pixel
    const int ly = threadIdx.x / kLbTile, lx = threadIdx.x % kLbTile;
    const __half z = __float2half_rn((0.f - 127.5f) * (1.0f / 128.0f)), o0 = __float2half_rn(0.f);
    __align__(16) __half vals[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) vals[j] = j < 27 ? z : o0;
    __half* o = p.out + pcb_prow(img, oy0 + ly, ox0 + lx, half, half) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) ((uint4*)o)[j] = ((const uint4*)vals)[j];
    return;
  }
  PcbView v = pcb_make_view(p.frames + (size_t)img * p.h * p.w * 3, p.h, p.w, p.rot, p.pad);
  const uint8_t* frame = p.frames + (size_t)img * p.h * p.w * 3;
  for (int i = threadIdx.x; i < kLbDet * kLbDet; i += blockDim.x) {
    const int ty = i / kLbDet, tx = i - ty * kLbDet;
    const int y = dy0 + ty, x = dx0 + tx;
    uint8_t o[3] = {0, 0, 0};
    uint8_t inside = 0;
    if (y >= 0 && y < p.S && x >= 0 && x < p.S) {
      inside = 1;
      if (y < p.new_h && x < p.new_w) {
        if (kPlain || p.mode == PCB_RS_LINEAR) {
          const LinTab cx = p.xtab[x], cy = p.ytab[y];
          const int x1 = pcb_iminf(cx.s + 1, v.vw - 1);
          const int y0 = pcb_clampi(cy.s, 0, v.vh - 1), y1 = pcb_clampi(cy.s + 1, 0, v.vh - 1);
          const uint8_t *p00, *p01, *p10, *p11;
          if (kPlain) {
            const int xs = pcb_clampi(cx.s, 0, p.w - 1);          // what pcb_view_px does for an unrotated, unpadded view
            const uint8_t* r0 = frame + (size_t)y0 * p.w * 3;
            const uint8_t* r1 = frame + (size_t)y1 * p.w * 3;
            const int xe = pcb_clampi(x1, 0, p.w - 1);
            p00 = r0 + xs * 3; p01 = r0 + xe * 3; p10 = r1 + xs * 3; p11 = r1 + xe * 3;
          } else {
            p00 = pcb_view_px(v, y0, cx.s);
            p01 = pcb_view_px(v, y0, x1);
            p10 = pcb_view_px(v, y1, cx.s);
            p11 = pcb_view_px(v, y1, x1);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int h0 = p00[c] * cx.a0 + p01[c] * cx.a1;
            const int h1 = p10[c] * cx.a0 + p11[c] * cx.a1;
            o[c] = pcb_sat_u8((((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2);
          }
        } else {
          pcb_resize_px(v, p.plan, y, x, p.new_h, p.new_w, o);
        }
      }
      if (p.det_img && ty >= 1 && tx >= 1) {   // each det pixel is owned by exactly one block
        uint8_t* d = p.det_img + (((size_t)img * p.S + y) * p.S + x) * 3;
        d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
      }
    }
    *(uint32_t*)tile[ty][tx] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)inside << 24);
  }
  __syncthreads();
  const int ly = threadIdx.x / kLbTile, lx = threadIdx.x % kLbTile;
  const int oy = oy0 + ly, ox = ox0 + lx;
  if (oy >= half || ox >= half) return;
  __align__(16) __half vals[32];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const uint32_t t = *(const uint32_t*)tile[2 * ly + ky][2 * lx + kx];      // B | G << 8 | R << 16 | inside << 24
      const bool in = (t >> 24) != 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // blobFromImage(swapRB): channel c of the blob is BGR channel 2-c; (v - 127.5) / 128 is exact in fp16
        const uint32_t v = (t >> (8 * (2 - c))) & 0xffu;
        vals[(ky * 3 + kx) * 3 + c] = lut[in ? v : 256u];
      }
    }
#pragma unroll
  for (int j = 27; j < 32; ++j) vals[j] = __float2half_rn(0.f);
  __half* o = p.out + pcb_prow(img, oy, ox, half, half) * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) ((uint4*)o)[j] = ((const uint4*)vals)[j];
}

// ------------------------------------------------------------------ ArcFace input
// chips [f][112][112][3] BGR -> patch tensor [f(*2)][114][114][32].  with_flip: 0 = chips only, 1 = chips then (images
// f..2f-1) their mirror images, 2 = mirror images only (flip-TTA computed lazily for faces that turn out to need it).
__global__ void __launch_bounds__(256) chip_patch_kernel(const uint8_t* __restrict__ chips, __half* __restrict__ out, int f, int with_flip) {
  // rgb.astype(float32) / 127.5 - 1.0 (float32 ops), then fp16 storage: a function of the byte alone, so it is tabulated once per
  // block with the exact operations (the IEEE division was ~15 instructions x 27 taps per pixel; ncu: 213 us per 504 chips at 26 %
  // of HBM peak, instruction bound)
  __shared__ __half lut[256];
  lut[threadIdx.x] = __float2half_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 127.5f), 1.0f));
  __syncthreads();
  const int total_imgs = with_flip == 1 ? 2 * f : f;
  const long long total = (long long)total_imgs * PCB_CHIP * PCB_CHIP;
  const __half zero = __float2half_rn(0.f);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % PCB_CHIP);
    const int y = (int)((idx / PCB_CHIP) % PCB_CHIP);
    const int img = (int)(idx / (PCB_CHIP * PCB_CHIP));
    const bool flip = with_flip == 2 || img >= f;
    const uint8_t* chip = chips + (size_t)(img >= f ? img - f : img) * PCB_CHIP * PCB_CHIP * 3;
    __align__(16) __half vals[32];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool in = yy >= 0 && yy < PCB_CHIP && xx >= 0 && xx < PCB_CHIP;
        const int sx = flip ? (PCB_CHIP - 1 - xx) : xx;
        const uint8_t* px = chip + ((size_t)(in ? yy : 0) * PCB_CHIP + (in ? sx : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) vals[(ky * 3 + kx) * 3 + c] = in ? lut[px[2 - c]] : zero;      // BGR -> RGB
      }
#pragma unroll
    for (int j = 27; j < 32; ++j) vals[j] = zero;
    __half* o = out + pcb_prow(img, y, x, PCB_CHIP, PCB_CHIP) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) ((uint4*)o)[j] = ((const uint4*)vals)[j];
  }
}

inline int grid_for(long long total, pcb_ctx* c, int per_block = 256) {
  long long b = (total + per_block - 1) / per_block;
  long long cap = (long long)c->num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

int resize_common(pcb_ctx* c, const uint8_t* src, int n, int h, int w, uint8_t* dst, int nh, int nw, bool area, double fx = 0.0,
                  double fy = 0.0) {
  PCB_ENTER(c);
  if (!src || !dst || n <= 0 || h <= 0 || w <= 0 || nh <= 0 || nw <= 0) return pcb_fail(c, PCB_ERR_ARG, "resize: bad arguments");
  PcbResizePlan plan = pcb_resize_plan(h, w, nh, nw, area, fx, fy);
  if (plan.mode == PCB_RS_AREA_INT && plan.isx == 2 && plan.isy == 2 && (nw % 16) == 0 && ((uintptr_t)src % 16) == 0 &&
      ((uintptr_t)dst % 16) == 0 && w == 2 * nw && h == 2 * nh) {
    const long long total = (long long)n * nh * (nw / 16);
    area2x_kernel<<<grid_for(total, c), 256, 0, c->stream>>>((const uint4*)src, (uint4*)dst, n, h, w, nh, nw);
    PCB_LAUNCH_CHECK(c, "area2x_kernel");
    return PCB_OK;
  }
  const long long total = (long long)n * nh * nw;
  resize_u8_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(src, dst, n, h, w, nh, nw, plan);
  PCB_LAUNCH_CHECK(c, "resize_u8_kernel");
  return PCB_OK;
}

}  // namespace

extern "C" int pcb_resize_area(pcb_ctx* c, const uint8_t* src, int n, int h, int w, uint8_t* dst, int nh, int nw) {
  return resize_common(c, src, n, h, w, dst, nh, nw, true);
}
extern "C" int pcb_resize_linear(pcb_ctx* c, const uint8_t* src, int n, int h, int w, uint8_t* dst, int nh, int nw) {
  return resize_common(c, src, n, h, w, dst, nh, nw, false);
}
extern "C" int pcb_resize_factor(pcb_ctx* c, const uint8_t* src, int n, int h, int w, uint8_t* dst, double fx, double fy, int inter_area) {
  // cv2.resize(img, None, fx=, fy=): dsize = (cvRound(w*fx), cvRound(h*fy)), inv_scale = the factor itself
  if (!(fx > 0.0) || !(fy > 0.0)) return pcb_fail(c, PCB_ERR_ARG, "resize_factor: bad factors");
  return resize_common(c, src, n, h, w, dst, pcb_cvround_d((double)h * fy), pcb_cvround_d((double)w * fx), inter_area != 0, fx, fy);
}

// letterbox geometry exactly as InsightFace SCRFD.detect computes it (python float arithmetic)
void pcb_letterbox_geometry(int vh, int vw, int S, int* new_h, int* new_w, double* det_scale) {
  const double im_ratio = (double)vh / (double)vw;
  const double model_ratio = 1.0;
  if (im_ratio > model_ratio) {
    *new_h = S;
    *new_w = (int)((double)S / im_ratio);
  } else {
    *new_w = S;
    *new_h = (int)((double)S * im_ratio);
  }
  *det_scale = (double)(*new_h) / (double)vh;
}


int pcb_letterbox_impl(pcb_ctx* c, const uint8_t* frames, int n, int h, int w, int S, int rot, int pad, __half* out,
                       uint8_t* det_img, double* det_scale_out) {
  if (!frames || !out || n <= 0 || S % 32 != 0 || S < 32) return pcb_fail(c, PCB_ERR_ARG, "letterbox: bad arguments");
  if (!(rot == 0 || rot == 90 || rot == 180 || rot == 270) || pad < 0) return pcb_fail(c, PCB_ERR_ARG, "letterbox: bad rot/pad");
  PcbView v = pcb_make_view(frames, h, w, rot, pad);
  LetterboxParams p{};
  p.frames = frames; p.n = n; p.h = h; p.w = w; p.rot = rot; p.pad = pad; p.S = S;
  double ds;
  pcb_letterbox_geometry(v.vh, v.vw, S, &p.new_h, &p.new_w, &ds);
  if (det_scale_out) *det_scale_out = ds;
  if (p.new_h < 1 || p.new_w < 1) return pcb_fail(c, PCB_ERR_ARG, "letterbox: degenerate aspect ratio");
  p.plan = pcb_resize_plan(v.vh, v.vw, p.new_h, p.new_w, false);
  p.mode = p.plan.mode;
  if (p.mode == PCB_RS_LINEAR) {
    std::vector<int> key = {v.vh, v.vw, p.new_h, p.new_w};
    auto& cache = c->lin_tabs;          // owned by the context: freed (with every other allocation) by pcb_destroy
    auto it = cache.find(key);
    if (it == cache.end()) {
      void* ex = pcb_dev_alloc(c, sizeof(LinTab) * p.new_w, false);
      void* ey = pcb_dev_alloc(c, sizeof(LinTab) * p.new_h, false);
      if (!ex || !ey) return pcb_fail(c, PCB_ERR_CUDA, "letterbox: table alloc");
      lin_table_kernel<<<(p.new_w + 127) / 128, 128, 0, c->stream>>>((LinTab*)ex, v.vw, p.new_w, 1);
      PCB_LAUNCH_CHECK(c, "lin_table_kernel");
      lin_table_kernel<<<(p.new_h + 127) / 128, 128, 0, c->stream>>>((LinTab*)ey, v.vh, p.new_h, 0);
      PCB_LAUNCH_CHECK(c, "lin_table_kernel");
      it = cache.emplace(key, std::make_pair(ex, ey)).first;
    }
    p.xtab = (const LinTab*)it->second.first;
    p.ytab = (const LinTab*)it->second.second;
  }
  p.out = out;
  p.det_img = det_img;
  const int half = S / 2;
  dim3 grid((half + kLbTile - 1) / kLbTile, (half + kLbTile - 1) / kLbTile, n);
  if (p.mode == PCB_RS_LINEAR && rot == 0 && pad == 0) letterbox_kernel<true><<<grid, 256, 0, c->stream>>>(p);
  else letterbox_kernel<false><<<grid, 256, 0, c->stream>>>(p);
  PCB_LAUNCH_CHECK(c, "letterbox_kernel");
  return PCB_OK;
}

int pcb_chip_patch_impl(pcb_ctx* c, const uint8_t* chips, int f, int with_flip, __half* out) {
  const long long total = (long long)(with_flip == 1 ? 2 * f : f) * PCB_CHIP * PCB_CHIP;
  chip_patch_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(chips, out, f, with_flip);
  PCB_LAUNCH_CHECK(c, "chip_patch_kernel");
  return PCB_OK;
}
