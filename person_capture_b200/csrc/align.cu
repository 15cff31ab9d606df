// K4: per-face alignment to 112x112 + quality.  Replaces, from
// person_capture/face_embedder.py::_extract_with_scrfd_raw:
//   :2439-2443  cross-pass suppression (sort by (score, area) desc, keep when IoU < 0.45; _iou :2484-2494)
//   :2447-2452  integer crop of the frame
//   :2454-2460  _canon_5pts (:1430-1463) -> _align_by_5pts (:1465-1473: LMedS similarity + warpAffine
//               INTER_LINEAR/BORDER_REFLECT on the crop) | _upright_by_eye_roll (:1571-1647) | cv2.resize
//   :2461       _face_quality (:1274-1276: variance of the 4-neighbour Laplacian of BGR2GRAY)
// All pixel arithmetic is OpenCV's fixed point (pcb_cvmath.h), so chips are bit-identical to cv2
// whenever the similarity matrix agrees (LMedS inlier sets are exact; the final least-squares
// fit differs from OpenCV's LM refinement by ~1e-13).
#include "pcb_common.cuh"
#include "pcb_cvmath.h"

namespace {

__constant__ float kArcDst[10] = {38.2946f, 51.6963f, 73.5318f, 51.5014f, 56.0252f, 71.7366f, 41.5493f, 92.3655f, 70.7299f, 92.2041f};

struct FacePlan {
  int frame, kind;
  int box[4];
  int rw, rh;            // rotated-crop dims (kinds 1, 2)
  long long roll_off;    // byte offset of the rotated crop in the roll scratch
  double M[6];           // final similarity (kinds 0, 1)
  double Mrot[6];        // eye-roll rotation (kinds 1, 2)
};

struct SelectParams {
  int n, h, w, max_det, max_faces;
  const int* acc_box; const float* acc_kps; const float* acc_score; const int* acc_count;
  int* kept_idx;         // [n][max_det] scratch
  int* face_count;       // [n]
  int* face_off;         // [n]
  int* face_total;       // [1]
  int* err;
};

__device__ __forceinline__ double iou_int(const int* a, const int* b) {
  const int x1 = max(a[0], b[0]), y1 = max(a[1], b[1]), x2 = min(a[2], b[2]), y2 = min(a[3], b[3]);
  const long long inter = (long long)max(0, x2 - x1) * max(0, y2 - y1);
  const long long aa = (long long)max(0, a[2] - a[0]) * max(0, a[3] - a[1]);
  const long long ab = (long long)max(0, b[2] - b[0]) * max(0, b[3] - b[1]);
  const long long den = aa + ab - inter;
  return den > 0 ? (double)inter / (double)den : 0.0;
}

__global__ void select_kernel(const SelectParams p) {
  extern __shared__ int sm_rank[];   // [max_det] sorted order
  const int img = blockIdx.x;
  const int m = p.acc_count[img];
  const int* boxes = p.acc_box + (size_t)img * p.max_det * 4;
  const float* scores = p.acc_score + (size_t)img * p.max_det;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const float si = scores[i];
    const long long ai = (long long)(boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      if (j == i) continue;
      const float sj = scores[j];
      const long long aj = (long long)(boxes[4 * j + 2] - boxes[4 * j]) * (boxes[4 * j + 3] - boxes[4 * j + 1]);
      const bool greater = sj > si || (sj == si && aj > ai);
      const bool equal = sj == si && aj == ai;
      if (greater || (equal && j < i)) ++rank;   // sorted(..., reverse=True) is stable
    }
    sm_rank[rank] = i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int* kept = p.kept_idx + (size_t)img * p.max_det;
    int k = 0;
    for (int r = 0; r < m; ++r) {
      const int i = sm_rank[r];
      bool ok = true;
      for (int q = 0; q < k && ok; ++q) ok = iou_int(boxes + 4 * i, boxes + 4 * kept[q]) < 0.45;
      if (ok) kept[k++] = i;
    }
    p.face_count[img] = k;
  }
}

__global__ void scan_kernel(const SelectParams p) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < p.n; ++i) { p.face_off[i] = tot; tot += p.face_count[i]; }
    if (tot > p.max_faces) { atomicCAS(p.err, 0, 301); tot = p.max_faces; }
    p.face_total[0] = tot;
  }
}

struct PlanParams {
  SelectParams s;
  FacePlan* plans;
  int* face_frame; int* face_box; int* face_kind;
  unsigned long long* roll_used;   // bump allocator (bytes)
  unsigned long long roll_cap;
};

__device__ void rotation_matrix(double cx, double cy, double angle_deg, double scale, double M[6]) {
  const double a = angle_deg * 3.1415926535897932384626433832795 / 180.0;
  const double alpha = cos(a) * scale, beta = sin(a) * scale;
  M[0] = alpha; M[1] = beta; M[2] = (1.0 - alpha) * cx - beta * cy;
  M[3] = -beta; M[4] = alpha; M[5] = beta * cx + (1.0 - alpha) * cy;
}

__device__ bool align_matrix(const float* canon, double M[6]) {
  float dst[10];
  for (int i = 0; i < 10; ++i) dst[i] = kArcDst[i];
  if (pcb_lmeds_similarity(canon, dst, 5, M)) return true;
  return pcb_lmeds_similarity(canon, dst, 3, M);
}

__global__ void plan_kernel(const PlanParams p) {
  const int img = blockIdx.x;
  const int k = p.s.face_count[img];
  const int off = p.s.face_off[img];
  for (int f = threadIdx.x; f < k; f += blockDim.x) {
    const int gi = off + f;
    if (gi >= p.s.max_faces) continue;
    const int i = p.s.kept_idx[(size_t)img * p.s.max_det + f];
    const int* b = p.s.acc_box + ((size_t)img * p.s.max_det + i) * 4;
    const float* kp = p.s.acc_kps + ((size_t)img * p.s.max_det + i) * 10;
    FacePlan pl;
    pl.frame = img;
    const int W0 = p.s.w, H0 = p.s.h;
    const int xi1 = max(0, min(W0 - 1, b[0])), yi1 = max(0, min(H0 - 1, b[1]));
    const int xi2 = max(xi1 + 1, min(W0, b[2])), yi2 = max(yi1 + 1, min(H0, b[3]));
    pl.box[0] = xi1; pl.box[1] = yi1; pl.box[2] = xi2; pl.box[3] = yi2;
    const int cw = xi2 - xi1, ch = yi2 - yi1;
    pl.rw = cw; pl.rh = ch; pl.roll_off = 0;
    for (int j = 0; j < 6; ++j) { pl.M[j] = 0; pl.Mrot[j] = 0; }
    float pts[10], canon[10];
    for (int j = 0; j < 10; ++j) pts[j] = kp[j];
    int kind = 3;
    if (pcb_canon_5pts(pts, canon)) {
      kind = align_matrix(canon, pl.M) ? 0 : 3;
    } else {
      // _upright_by_eye_roll
      bool finite = true;
      for (int j = 0; j < 10; ++j) finite = finite && isfinite(pts[j]);
      if (finite) {
        float c[10];
        for (int j = 0; j < 5; ++j) {
          c[2 * j] = fminf(fmaxf(pts[2 * j], 0.f), (float)max(0, cw - 1));
          c[2 * j + 1] = fminf(fmaxf(pts[2 * j + 1], 0.f), (float)max(0, ch - 1));
        }
        float vx = __fsub_rn(c[2], c[0]), vy = __fsub_rn(c[3], c[1]);
        bool have = hypotf(vx, vy) >= 1e-3f;
        if (!have) {
          vx = __fsub_rn(c[8], c[6]); vy = __fsub_rn(c[9], c[7]);
          have = hypotf(vx, vy) >= 1e-3f;
        }
        if (have) {
          double angle = atan2((double)vy, (double)vx) * (180.0 / 3.1415926535897932384626433832795);
          if (angle < -90.0) angle += 180.0;
          else if (angle > 90.0) angle -= 180.0;
          if (fabs(angle) >= 8.0) {
            if (angle > 80.0) angle = 90.0;
            else if (angle < -80.0) angle = -90.0;
            const int side = max(ch, cw);
            const double scale = side <= 256 ? 1.0 : 256.0 / (double)side;
            rotation_matrix((double)((float)(cw / 2.0)), (double)((float)(ch / 2.0)), -angle, scale, pl.Mrot);
            float pr[10], canon2[10];
            for (int j = 0; j < 5; ++j) {
              pr[2 * j] = (float)(pl.Mrot[0] * (double)pts[2 * j] + pl.Mrot[1] * (double)pts[2 * j + 1] + pl.Mrot[2]);
              pr[2 * j + 1] = (float)(pl.Mrot[3] * (double)pts[2 * j] + pl.Mrot[4] * (double)pts[2 * j + 1] + pl.Mrot[5]);
            }
            kind = 2;
            if (pcb_canon_5pts(pr, canon2) && align_matrix(canon2, pl.M)) kind = 1;
            const unsigned long long bytes = ((unsigned long long)cw * ch * 3 + 255ull) & ~255ull;
            const unsigned long long o = atomicAdd(p.roll_used, bytes);
            if (o + bytes > p.roll_cap) { atomicCAS(p.s.err, 0, 302); kind = 3; }
            else pl.roll_off = (long long)o;
          }
        }
      }
    }
    pl.kind = kind;
    p.plans[gi] = pl;
    p.face_frame[gi] = img;
    p.face_kind[gi] = kind;
    for (int j = 0; j < 4; ++j) p.face_box[(size_t)gi * 4 + j] = pl.box[j];
  }
}

struct RenderParams {
  const uint8_t* frames; int h, w;
  const FacePlan* plans; const int* face_total;
  uint8_t* roll; uint8_t* chips; double* quality;
};

__global__ void __launch_bounds__(256) roll_kernel(const RenderParams p) {
  const int gi = blockIdx.x;
  if (gi >= p.face_total[0]) return;
  const FacePlan pl = p.plans[gi];
  if (pl.kind != 1 && pl.kind != 2) return;
  const int cw = pl.rw, ch = pl.rh;
  const uint8_t* crop = p.frames + ((size_t)pl.frame * p.h * p.w + (size_t)pl.box[1] * p.w + pl.box[0]) * 3;
  const PcbWarp wc = pcb_invert_affine(pl.Mrot);
  uint8_t* dst = p.roll + pl.roll_off;
  for (int i = threadIdx.x; i < cw * ch; i += blockDim.x) {
    const int y = i / cw, x = i - y * cw;
    uint8_t o[3];
    pcb_warp_px(crop, (long long)p.w * 3, ch, cw, wc, y, x, o);
    dst[(size_t)i * 3] = o[0]; dst[(size_t)i * 3 + 1] = o[1]; dst[(size_t)i * 3 + 2] = o[2];
  }
}

__global__ void __launch_bounds__(256) chip_kernel(const RenderParams p) {
  __shared__ uint8_t gray[PCB_CHIP * PCB_CHIP];
  __shared__ long long red1[8], red2[8];
  const int gi = blockIdx.x;
  if (gi >= p.face_total[0]) return;
  const FacePlan pl = p.plans[gi];
  const int cw = pl.box[2] - pl.box[0], ch = pl.box[3] - pl.box[1];
  const uint8_t* crop = p.frames + ((size_t)pl.frame * p.h * p.w + (size_t)pl.box[1] * p.w + pl.box[0]) * 3;
  const bool rolled = pl.kind == 1 || pl.kind == 2;
  const uint8_t* src = rolled ? p.roll + pl.roll_off : crop;
  const int sw = cw, sh = ch;
  const long long pitch_px = rolled ? cw : p.w;
  uint8_t* chip = p.chips + (size_t)gi * PCB_CHIP * PCB_CHIP * 3;
  PcbWarp wc;
  PcbView v;
  PcbResizePlan rp;
  if (pl.kind <= 1) {
    wc = pcb_invert_affine(pl.M);
  } else {
    v = pcb_make_view(src, sh, sw, 0, 0);
    v.ld = (int)pitch_px;
    rp = pcb_resize_plan(sh, sw, PCB_CHIP, PCB_CHIP, max(sh, sw) > PCB_CHIP);
  }
  for (int i = threadIdx.x; i < PCB_CHIP * PCB_CHIP; i += blockDim.x) {
    const int y = i / PCB_CHIP, x = i - y * PCB_CHIP;
    uint8_t o[3];
    if (pl.kind <= 1) pcb_warp_px(src, pitch_px * 3, sh, sw, wc, y, x, o);
    else pcb_resize_px(v, rp, y, x, PCB_CHIP, PCB_CHIP, o);
    chip[i * 3] = o[0]; chip[i * 3 + 1] = o[1]; chip[i * 3 + 2] = o[2];
    gray[i] = pcb_gray(o);
  }
  __syncthreads();
  // 4-neighbour Laplacian (ksize=1) with BORDER_REFLECT_101; exact integer sums
  long long s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < PCB_CHIP * PCB_CHIP; i += blockDim.x) {
    const int y = i / PCB_CHIP, x = i - y * PCB_CHIP;
    const int yu = y == 0 ? 1 : y - 1, yd = y == PCB_CHIP - 1 ? PCB_CHIP - 2 : y + 1;
    const int xl = x == 0 ? 1 : x - 1, xr = x == PCB_CHIP - 1 ? PCB_CHIP - 2 : x + 1;
    const int l = (int)gray[yu * PCB_CHIP + x] + gray[yd * PCB_CHIP + x] + gray[y * PCB_CHIP + xl] + gray[y * PCB_CHIP + xr] - 4 * (int)gray[i];
    s1 += l;
    s2 += (long long)l * l;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { red1[threadIdx.x >> 5] = s1; red2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t1 = 0, t2 = 0;
    for (int i = 0; i < 8; ++i) { t1 += red1[i]; t2 += red2[i]; }
    const double n = (double)(PCB_CHIP * PCB_CHIP);
    // population variance: (n*sum(x^2) - sum(x)^2) / n^2, numerator exact in int64
    p.quality[gi] = (double)((long long)(PCB_CHIP * PCB_CHIP) * t2 - t1 * t1) / (n * n);
  }
}

}  // namespace

extern "C" int pcb_align(pcb_ctx* c, const pcb_align_args* a) {
  if (!a || a->n <= 0 || a->max_det <= 0 || a->max_faces <= 0) return pcb_fail(c, PCB_ERR_ARG, "align: bad arguments");
  PCB_ENTER(c);
  // context-owned scratch (freed with the context): kept_idx [n][max_det], face_off [n], plans [max_faces], roll allocator + buffer
  size_t* sz = c->align_sz;
  void** bf = c->align_buf;
  const size_t want[4] = {(size_t)a->n * a->max_det * sizeof(int), (size_t)a->n * sizeof(int), (size_t)a->max_faces * sizeof(FacePlan),
                          (size_t)256 << 20};
  for (int i = 0; i < 4; ++i) {
    if (sz[i] < want[i]) {
      void* nb = pcb_dev_alloc(c, want[i], false);
      if (!nb) return pcb_fail(c, PCB_ERR_CUDA, "align: scratch alloc failed");
      if (bf[i]) {                        // kernels of earlier calls may still read the old buffer
        PCB_CUDA(c, cudaStreamSynchronize(c->stream));
        pcb_dev_free(c, bf[i]);
      }
      bf[i] = nb;
      sz[i] = want[i];
    }
  }
  if (!bf[4]) {
    bf[4] = pcb_dev_alloc(c, 256, true);
    if (!bf[4]) return pcb_fail(c, PCB_ERR_CUDA, "align: scratch alloc failed");
  }
  PCB_CUDA(c, cudaMemsetAsync(bf[4], 0, 8, c->stream));
  SelectParams sp{};
  sp.n = a->n; sp.h = a->h; sp.w = a->w; sp.max_det = a->max_det; sp.max_faces = a->max_faces;
  sp.acc_box = a->acc_box_dev; sp.acc_kps = a->acc_kps_dev; sp.acc_score = a->acc_score_dev; sp.acc_count = a->acc_count_dev;
  sp.kept_idx = (int*)bf[0]; sp.face_count = a->face_count_dev; sp.face_off = (int*)bf[1]; sp.face_total = a->face_total_dev;
  sp.err = c->d_err;
  select_kernel<<<a->n, 128, a->max_det * sizeof(int), c->stream>>>(sp);
  PCB_LAUNCH_CHECK(c, "select_kernel");
  scan_kernel<<<1, 32, 0, c->stream>>>(sp);
  PCB_LAUNCH_CHECK(c, "scan_kernel");
  PlanParams pp{};
  pp.s = sp; pp.plans = (FacePlan*)bf[2];
  pp.face_frame = a->face_frame_dev; pp.face_box = a->face_box_dev; pp.face_kind = a->face_kind_dev;
  pp.roll_used = (unsigned long long*)bf[4]; pp.roll_cap = want[3];
  plan_kernel<<<a->n, 32, 0, c->stream>>>(pp);
  PCB_LAUNCH_CHECK(c, "plan_kernel");
  RenderParams rp{};
  rp.frames = a->frames_dev; rp.h = a->h; rp.w = a->w; rp.plans = (const FacePlan*)bf[2]; rp.face_total = a->face_total_dev;
  rp.roll = (uint8_t*)bf[3]; rp.chips = a->chips_dev; rp.quality = a->quality_dev;
  roll_kernel<<<a->max_faces, 256, 0, c->stream>>>(rp);
  PCB_LAUNCH_CHECK(c, "roll_kernel");
  chip_kernel<<<a->max_faces, 256, 0, c->stream>>>(rp);
  PCB_LAUNCH_CHECK(c, "chip_kernel");
  return PCB_OK;
}
