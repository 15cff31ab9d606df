// K3: anchor decode + score filter + sort + greedy NMS + the reference's per-pass bookkeeping.
// Replaces InsightFace SCRFD.forward / SCRFD.detect post-processing / SCRFD.nms (third-party,
// restated in oracle/scrfd_detect.py; thr 0.4, "+1" areas, keep when ovr <= thr) and
// FaceEmbedder._extract_with_scrfd_raw::_accumulate + the min-size filter
// (person_capture/face_embedder.py:2214-2245, 2317-2322).
//
// Head maps are the fp32 P-layout outputs of the SCRFD graph (fp32 so that box / landmark distances of
// 8..16 stride units are not quantised to 2^-7), 32 channels per pixel:
//   [cls a0, cls a1 | reg a0 (l,t,r,b), reg a1 | kps a0 (10), kps a1 (10) | 2 pad].
// All box arithmetic is float32 with explicit non-fused operations in numpy's order, so keep
// sets are bit-identical to the oracle given identical head maps.  Ties in score are broken by
// anchor row (level 8 rows first, then 16, 32; row = (y*w + x)*2 + a), the rule the oracle pins.
#include "pcb_common.cuh"

namespace {

constexpr int kCandCap = 8192;   // candidates (score >= thr) per frame before NMS

struct Cand {
  float score;
  int order;        // global anchor row
  float box[4];
  float kps[10];
};

struct DecodeParams {
  const float* head[3];
  int hw[3];              // feature map edge per level (S/8, S/16, S/32)
  int row_off[3];         // first anchor row of each level
  float reg_scale[3];
  int n, S;
  float thr, det_scale;
  Cand* cands;            // [n][kCandCap]
  int* cand_count;        // [n]
  int* err;
};

__global__ void decode_kernel(const DecodeParams p) {
  const int img = blockIdx.y;
  const int per_img = p.row_off[2] + p.hw[2] * p.hw[2] * 2;
  // conservative float pre-filter in logit space; the exact test is on the float32 sigmoid below
  const float lthr = logf(p.thr / (1.f - p.thr)) - 0.01f;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < per_img; row += gridDim.x * blockDim.x) {
    const int lvl = row >= p.row_off[2] ? 2 : (row >= p.row_off[1] ? 1 : 0);
    const int r = row - p.row_off[lvl];
    const int a = r & 1;
    const int loc = r >> 1;
    const int w = p.hw[lvl];
    const int y = loc / w, x = loc - y * w;
    const float* px = p.head[lvl] + pcb_prow(img, y, x, w, w) * 32;
    const float logit = px[a];
    if (!(logit >= lthr)) continue;
    // torch.sigmoid on float32: evaluate in double and round once
    const float score = (float)(1.0 / (1.0 + exp(-(double)logit)));
    if (!(score >= p.thr)) continue;
    const int slot = atomicAdd(p.cand_count + img, 1);
    if (slot >= kCandCap) { atomicCAS(p.err, 0, 201); continue; }
    Cand c;
    c.score = score;
    c.order = row;
    const float stride = (float)(8 << lvl);
    const float cx = __fmul_rn((float)x, stride), cy = __fmul_rn((float)y, stride);
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = __fmul_rn(__fmul_rn(px[2 + 4 * a + j], p.reg_scale[lvl]), stride);
    c.box[0] = __fdiv_rn(__fsub_rn(cx, d[0]), p.det_scale);
    c.box[1] = __fdiv_rn(__fsub_rn(cy, d[1]), p.det_scale);
    c.box[2] = __fdiv_rn(__fadd_rn(cx, d[2]), p.det_scale);
    c.box[3] = __fdiv_rn(__fadd_rn(cy, d[3]), p.det_scale);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const float k = __fmul_rn(px[10 + 10 * a + j], stride);
      c.kps[j] = __fdiv_rn(__fadd_rn((j & 1) ? cy : cx, k), p.det_scale);
    }
    p.cands[(size_t)img * kCandCap + slot] = c;
  }
}

struct NmsParams {
  const Cand* cands;
  const int* cand_count;
  int n, max_det;
  int rot, pad, fix_mode;
  float fix_scale_inv;
  int H0, W0, min_box;
  float* det;        // [n][max_det][5]
  float* kps;        // [n][max_det][10]
  int* raw_count;
  int* acc_box;
  float* acc_kps;
  float* acc_score;
  int* acc_count;
  int* acc_unfiltered;
  int* err;
};

__device__ __forceinline__ void unrotate_i(int xr, int yr, int deg, int W0, int H0, int& xo, int& yo) {
  if (deg == 90) { xo = yr; yo = H0 - 1 - xr; }
  else if (deg == 180) { xo = W0 - 1 - xr; yo = H0 - 1 - yr; }
  else if (deg == 270) { xo = W0 - 1 - yr; yo = xr; }
  else { xo = xr; yo = yr; }
}
__device__ __forceinline__ void unrotate_d(double xr, double yr, int deg, int W0, int H0, double& xo, double& yo) {
  if (deg == 90) { xo = yr; yo = (double)(H0 - 1) - xr; }
  else if (deg == 180) { xo = (double)(W0 - 1) - xr; yo = (double)(H0 - 1) - yr; }
  else if (deg == 270) { xo = (double)(W0 - 1) - yr; yo = xr; }
  else { xo = xr; yo = yr; }
}

__global__ void __launch_bounds__(256) nms_kernel(const NmsParams p) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* keys = (unsigned long long*)nms_smem;                 // [kCandCap]
  unsigned short* slot_of = (unsigned short*)(keys + kCandCap);             // [kCandCap]
  unsigned char* dead = (unsigned char*)(slot_of + kCandCap);               // [kCandCap]
  int* kept_list = (int*)(dead + kCandCap);                                 // [1024]
  __shared__ int kept_n;
  const int img = blockIdx.x;
  int P = p.cand_count[img];
  if (P > kCandCap) P = kCandCap;
  const Cand* cands = p.cands + (size_t)img * kCandCap;
  int P2 = 1;
  while (P2 < P) P2 <<= 1;
  for (int i = threadIdx.x; i < P2; i += blockDim.x) {
    if (i < P) {
      // ascending sort of (~score_bits, order): descending score, then ascending anchor row
      const unsigned sb = __float_as_uint(cands[i].score);
      keys[i] = ((unsigned long long)(~sb) << 32) | (unsigned)cands[i].order;
    } else {
      keys[i] = ~0ull;
    }
    dead[i] = 0;
  }
  __syncthreads();
  // candidate slots are permuted alongside the keys
  for (int i = threadIdx.x; i < P2; i += blockDim.x) slot_of[i] = (unsigned short)i;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const unsigned long long a = keys[i], b = keys[ixj];
          if ((a > b) == up) {
            keys[i] = b; keys[ixj] = a;
            const unsigned short t = slot_of[i]; slot_of[i] = slot_of[ixj]; slot_of[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) kept_n = 0;
  __syncthreads();
  // greedy NMS in sorted order (SCRFD.nms): float32, "+1" extents, suppress when ovr > 0.4
  for (int i = 0; i < P; ++i) {
    if (dead[i]) continue;            // uniform: dead[] is only written between barriers
    const Cand& ci = cands[slot_of[i]];
    if (threadIdx.x == 0) {
      if (kept_n < 1024) kept_list[kept_n] = slot_of[i];
      kept_n++;
    }
    const float ix1 = ci.box[0], iy1 = ci.box[1], ix2 = ci.box[2], iy2 = ci.box[3];
    const float iarea = __fmul_rn(__fadd_rn(__fsub_rn(ix2, ix1), 1.f), __fadd_rn(__fsub_rn(iy2, iy1), 1.f));
    for (int j = i + 1 + threadIdx.x; j < P; j += blockDim.x) {
      if (dead[j]) continue;
      const Cand& cj = cands[slot_of[j]];
      const float xx1 = fmaxf(ix1, cj.box[0]), yy1 = fmaxf(iy1, cj.box[1]);
      const float xx2 = fminf(ix2, cj.box[2]), yy2 = fminf(iy2, cj.box[3]);
      const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
      const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
      const float inter = __fmul_rn(w, h);
      const float jarea = __fmul_rn(__fadd_rn(__fsub_rn(cj.box[2], cj.box[0]), 1.f), __fadd_rn(__fsub_rn(cj.box[3], cj.box[1]), 1.f));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
      if (!(ovr <= 0.4f)) dead[j] = 1;
    }
    __syncthreads();
  }
  __syncthreads();
  const int K = kept_n;
  if (K > p.max_det || K > 1024) {
    if (threadIdx.x == 0) {
      atomicCAS(p.err, 0, 202);
      p.raw_count[img] = 0;
      p.acc_count[img] = 0;
      if (p.acc_unfiltered) p.acc_unfiltered[img] = 0;
    }
    return;
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const Cand& c = cands[kept_list[i]];
    float* d = p.det + ((size_t)img * p.max_det + i) * 5;
    d[0] = c.box[0]; d[1] = c.box[1]; d[2] = c.box[2]; d[3] = c.box[3]; d[4] = c.score;
    float* k = p.kps + ((size_t)img * p.max_det + i) * 10;
#pragma unroll
    for (int j = 0; j < 10; ++j) k[j] = c.kps[j];
  }
  if (threadIdx.x == 0) {
    p.raw_count[img] = K;
    // _accumulate + min-size filter, order preserving (serial: K is small)
    int m = 0, m_all = 0;
    for (int i = 0; i < K; ++i) {
      const Cand& c = cands[kept_list[i]];
      float bb[4] = {c.box[0], c.box[1], c.box[2], c.box[3]};
      float kp[10];
      for (int j = 0; j < 10; ++j) kp[j] = c.kps[j];
      if (p.fix_mode == PCB_FIX_SCALE) {
        for (int j = 0; j < 4; ++j) bb[j] = __fmul_rn(bb[j], p.fix_scale_inv);
        for (int j = 0; j < 10; ++j) kp[j] = __fmul_rn(kp[j], p.fix_scale_inv);
      } else if (p.fix_mode == PCB_FIX_PADPROBE) {
        const float fp = (float)p.pad;
        for (int j = 0; j < 4; ++j) bb[j] = __fsub_rn(bb[j], fp);
        // python-float clamps stored back into the float32 array (face_embedder.py:2305-2314)
        double b0 = fmax(0.0, fmin((double)(p.W0 - 1), (double)bb[0])); bb[0] = (float)b0;
        double b1 = fmax(0.0, fmin((double)(p.H0 - 1), (double)bb[1])); bb[1] = (float)b1;
        double b2 = fmax((double)bb[0] + 1.0, fmin((double)p.W0, (double)bb[2])); bb[2] = (float)b2;
        double b3 = fmax((double)bb[1] + 1.0, fmin((double)p.H0, (double)bb[3])); bb[3] = (float)b3;
        for (int j = 0; j < 10; ++j) {
          const float lim = (float)((j & 1) ? (p.H0 - 1) : (p.W0 - 1));
          kp[j] = fminf(fmaxf(__fsub_rn(kp[j], fp), 0.f), lim);
        }
      } else if (p.fix_mode == PCB_FIX_UNPAD) {
        const float fp = (float)p.pad;
        for (int j = 0; j < 4; ++j) bb[j] = __fsub_rn(bb[j], fp);
        for (int j = 0; j < 10; ++j) kp[j] = __fsub_rn(kp[j], fp);
      }
      const int x1 = (int)bb[0], y1 = (int)bb[1], x2 = (int)bb[2], y2 = (int)bb[3];   // int(): toward zero
      int ax, ay, bx, by;
      unrotate_i(x1, y1, p.rot, p.W0, p.H0, ax, ay);
      unrotate_i(x2, y2, p.rot, p.W0, p.H0, bx, by);
      int xa1 = min(ax, bx), ya1 = min(ay, by), xa2 = max(ax, bx), ya2 = max(ay, by);
      xa1 = max(0, min(p.W0 - 1, xa1));
      ya1 = max(0, min(p.H0 - 1, ya1));
      xa2 = max(xa1 + 1, min(p.W0, xa2));
      ya2 = max(ya1 + 1, min(p.H0, ya2));
      if (xa2 - xa1 <= 2 || ya2 - ya1 <= 2) continue;
      ++m_all;
      if (xa2 - xa1 < p.min_box || ya2 - ya1 < p.min_box) continue;
      int* ob = p.acc_box + ((size_t)img * p.max_det + m) * 4;
      ob[0] = xa1; ob[1] = ya1; ob[2] = xa2; ob[3] = ya2;
      float* ok = p.acc_kps + ((size_t)img * p.max_det + m) * 10;
      for (int j = 0; j < 5; ++j) {
        double ox, oy;
        unrotate_d((double)kp[2 * j], (double)kp[2 * j + 1], p.rot, p.W0, p.H0, ox, oy);
        ok[2 * j] = (float)(ox - (double)xa1);
        ok[2 * j + 1] = (float)(oy - (double)ya1);
      }
      p.acc_score[(size_t)img * p.max_det + m] = c.score;
      ++m;
    }
    p.acc_count[img] = m;
    if (p.acc_unfiltered) p.acc_unfiltered[img] = m_all;
  }
}

}  // namespace

// scratch for candidates lives in the context (grown on demand)
static int ensure_scratch(pcb_ctx* c, size_t bytes) {
  if (c->scratch_bytes >= bytes) return PCB_OK;
  void* p = pcb_dev_alloc(c, bytes, false);
  if (!p) return pcb_fail(c, PCB_ERR_CUDA, "scratch alloc failed");
  if (c->scratch) {
    cudaStreamSynchronize(c->stream);
    pcb_dev_free(c, c->scratch);
  }
  c->scratch = p;
  c->scratch_bytes = bytes;
  return PCB_OK;
}

int pcb_decode_nms_impl(pcb_ctx* c, const float* h8, const float* h16, const float* h32, const float* reg_scale3,
                        const pcb_detect_args* a, float det_scale) {
  if (!a || a->n <= 0 || a->max_det <= 0 || a->max_det > 1024) return pcb_fail(c, PCB_ERR_ARG, "decode_nms: bad arguments");
  const size_t cand_bytes = (size_t)a->n * kCandCap * sizeof(Cand);
  const size_t need = cand_bytes + (size_t)a->n * sizeof(int) + 256;
  int rc = ensure_scratch(c, need);
  if (rc) return rc;
  Cand* cands = (Cand*)c->scratch;
  int* counts = (int*)((uint8_t*)c->scratch + ((cand_bytes + 255) / 256) * 256);
  PCB_CUDA(c, cudaMemsetAsync(counts, 0, (size_t)a->n * sizeof(int), c->stream));
  DecodeParams dp{};
  dp.head[0] = h8; dp.head[1] = h16; dp.head[2] = h32;
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    dp.hw[l] = a->S / (8 << l);
    dp.row_off[l] = off;
    off += dp.hw[l] * dp.hw[l] * 2;
    dp.reg_scale[l] = reg_scale3 ? reg_scale3[l] : 1.f;
  }
  dp.n = a->n; dp.S = a->S; dp.thr = a->det_thresh; dp.det_scale = det_scale;
  dp.cands = cands; dp.cand_count = counts; dp.err = c->d_err;
  dim3 grid((off + 255) / 256, a->n);
  if (grid.x > 64) grid.x = 64;
  decode_kernel<<<grid, 256, 0, c->stream>>>(dp);
  PCB_LAUNCH_CHECK(c, "decode_kernel");
  NmsParams np{};
  np.cands = cands; np.cand_count = counts; np.n = a->n; np.max_det = a->max_det;
  np.rot = a->rot_deg; np.pad = a->pad_replicate; np.fix_mode = a->fix_mode; np.fix_scale_inv = a->fix_scale_inv;
  np.H0 = a->orig_h; np.W0 = a->orig_w; np.min_box = a->min_box_px;
  np.det = a->det_dev; np.kps = a->kps_dev; np.raw_count = a->raw_count_dev;
  np.acc_box = a->acc_box_dev; np.acc_kps = a->acc_kps_dev; np.acc_score = a->acc_score_dev; np.acc_count = a->acc_count_dev; np.acc_unfiltered = a->acc_unfiltered_dev;
  np.err = c->d_err;
  const size_t nms_smem_bytes = (size_t)kCandCap * (8 + 2 + 1) + 1024 * 4;
  static unsigned long long attr_devs = 0;     // the opt-in is per device
  if (pcb_attr_needed(&attr_devs, c->device))
    PCB_CUDA(c, cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes));
  nms_kernel<<<a->n, 256, nms_smem_bytes, c->stream>>>(np);
  PCB_LAUNCH_CHECK(c, "nms_kernel");
  return PCB_OK;
}
