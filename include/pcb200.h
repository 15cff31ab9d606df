/* libpcb200 -- C ABI of the B200-native identity hot path (SCRFD detect -> 5-point align ->
 * ArcFace embed -> bank distance) that replaces the ONNX Runtime / TensorRT / InsightFace / cv2
 * calls inside the reference's FaceEmbedder (person_capture/face_embedder.py) and the
 * per-face distance in Processor._fd_min (person_capture/gui_app.py:660-674).
 *
 * The reference has no native interface for this path (its boundary is the Python object
 * FaceEmbedder, face_embedder.py:376-2508); this header follows the one native-binding
 * precedent in the reference (hdr_preview/pc_hdr_vulkan.h:19-46 bound by ctypes in
 * person_capture/hdr_preview.py:19-102): extern "C", opaque context, create/.../destroy --
 * plus int return codes and pcb_last_error(), because this path has no CPU fallback and a
 * failure must never be read as "no face".
 *
 * Conventions: every pointer named *_dev is a CUDA device pointer on the context's device
 * (e.g. torch.Tensor.data_ptr()); *_host is host memory.  Calls are asynchronous and ordered
 * on the context's stream; results are valid after pcb_sync() (or any call documented as
 * synchronous).  All functions return 0 on success, non-zero on failure.
 */
#ifndef PCB200_H
#define PCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcb_ctx pcb_ctx;

#define PCB_FEAT_DIM 512
#define PCB_CHIP 112

/* ---- context ------------------------------------------------------------------------- */
/* One context per GPU per thread.  `cuda_stream` may be NULL (the context creates its own). */
pcb_ctx* pcb_create(int device, void* cuda_stream);
void pcb_destroy(pcb_ctx* ctx);
const char* pcb_last_error(pcb_ctx* ctx);
/* Waits for the stream and checks the device-side error word (watchdogs, capacity overflow). */
int pcb_sync(pcb_ctx* ctx);
/* 0 = tcgen05/TMEM implicit GEMM (default), 1 = CUDA-core validation kernel (tests only),
 * 2 = first tcgen05 formulation (one TMA load per tap; A/B baseline for profiles). */
int pcb_set_conv_impl(pcb_ctx* ctx, int impl);
/* Geometry of the activation layout device buffers such as pcb_letterbox's patch tensor use: [n][h + pad][w + pad][c], image
 * pixel (y, x) at [y + pad_lo][x + pad_lo], every other element zero (default build: pad_lo = 1, pad = 2, a ring of zeros). */
void pcb_layout_pad(int* pad_lo, int* pad);
/* Number of kernels this library launched on the context since the last reset. */
long long pcb_launch_count(pcb_ctx* ctx);
void pcb_reset_launch_count(pcb_ctx* ctx);

/* Profiling for bench.py: when on, every convolution launch is bracketed by CUDA events on the
 * context's stream and its algorithmic FLOPs (2*out_px*cout*cin*taps from the layer shape) are
 * accumulated.  pcb_profile_read synchronises and returns the totals since the last reset. */
int pcb_set_profile(pcb_ctx* ctx, int on);
int pcb_profile_read(pcb_ctx* ctx, double* conv_ms, double* conv_flops, long long* conv_launches, int reset);

/* ---- graphs (replaces ort.InferenceSession construction, face_embedder.py:1102-1107, 891-915) */
enum { PCB_OP_CONV = 1, PCB_OP_AFFINE = 2, PCB_OP_MAXPOOL3S2 = 3, PCB_OP_AVGPOOL2 = 4,
       PCB_OP_UPSAMPLE_ADD = 5, PCB_OP_ADD = 6, PCB_OP_AFFINE_FLATTEN = 7, PCB_OP_FC = 8 };
enum { PCB_ACT_NONE = 0, PCB_ACT_RELU = 1, PCB_ACT_PRELU = 2 };
enum { PCB_MODEL_SCRFD = 0, PCB_MODEL_ARCFACE = 1 };

enum { PCB_OPF_OUT_F32 = 1 };   /* pcb_op.flags: primary output kept in fp32 (residual stream) */

typedef struct pcb_op {
  int32_t kind;             /* PCB_OP_* */
  int32_t in0, in1, out;    /* tensor ids; tensor 0 is the graph input; in1 = residual/second addend or -1 */
  int32_t cin, cout, k, stride, act;
  int32_t out2;             /* CONV only: second fp16 output = scale2 * y + bias2 of the activated value y, or -1
                               (fuses the BatchNorm that precedes the next block's first conv, iResNet bn1) */
  int32_t flags;            /* PCB_OPF_* */
  int64_t w_off;            /* byte offset in blob: fp16 [cout][cin][k][k] (FC: [cout][cin]); -1 if none */
  int64_t scale_off;        /* fp32 [cout] */
  int64_t bias_off;         /* fp32 [cout] */
  int64_t slope_off;        /* fp32 [cout] (PReLU) or -1 */
  int64_t scale2_off;       /* fp32 [cout] for out2, or -1 */
  int64_t bias2_off;
} pcb_op;

/* Loads a graph into `slot`.  Tensor 0 is the stem-patch input produced by the pre-processing
 * kernels: 27 channels = the 3x3 neighbourhood (ky,kx,rgb) of the normalised image, so the
 * first op must be the stem convolution given as k=3 over cin=3 (stride 2 for SCRFD, 1 for
 * ArcFace); the library executes it as a 1x1 GEMM over the patch tensor.
 * `outputs` lists the tensor ids the graph produces (SCRFD: head maps of strides 8,16,32). */
int pcb_model_load(pcb_ctx* ctx, int slot, const pcb_op* ops, int n_ops, int n_tensors,
                   const void* blob_host, size_t blob_bytes, const int32_t* outputs, int n_outputs,
                   const float* reg_scale3 /* SCRFD per-level bbox scale or NULL */);
/* Debug/parity helper (synchronous): copy tensor `tid` of the last run of `slot` to host as
 * fp32 NCHW [n][c][h][w]; returns dims through n,c,h,w.  out_host may be NULL to query dims. */
int pcb_model_get_tensor(pcb_ctx* ctx, int slot, int tid, float* out_host, int* n, int* c, int* h, int* w);

/* ---- K0: pre-scan downscale (replaces cv2.resize(..., INTER_AREA), gui_app.py:1505-1507) --- */
/* uint8 BGR HWC [n][h][w][3] -> [n][nh][nw][3]; bit-exact with OpenCV's INTER_AREA. */
int pcb_resize_area(pcb_ctx* ctx, const uint8_t* src_dev, int n, int h, int w,
                    uint8_t* dst_dev, int nh, int nw);
/* uint8 bilinear (cv2.resize default INTER_LINEAR; scale TTA at face_embedder.py:2263-2264). */
int pcb_resize_linear(pcb_ctx* ctx, const uint8_t* src_dev, int n, int h, int w,
                      uint8_t* dst_dev, int nh, int nw);
/* cv2.resize(img, None, fx=, fy=, interpolation=INTER_AREA|INTER_LINEAR): dst is
 * [n][cvRound(h*fy)][cvRound(w*fx)][3] and the sampling scale is the factor itself (not dst/src). */
int pcb_resize_factor(pcb_ctx* ctx, const uint8_t* src_dev, int n, int h, int w, uint8_t* dst_dev,
                      double fx, double fy, int inter_area);

/* ---- K1+K2+K3: one SCRFD pass (replaces scrfd.detect(img, input_size=(S,S)) with
 *      scrfd.det_thresh = thresh, face_embedder.py:2176-2187; InsightFace SCRFD.detect/forward/nms)
 *      followed by the reference's per-pass bookkeeping `_accumulate` + min-size filter
 *      (face_embedder.py:2214-2245, 2317-2322). ------------------------------------------------ */
enum { PCB_FIX_NONE = 0, PCB_FIX_SCALE = 1, PCB_FIX_PADPROBE = 2, PCB_FIX_UNPAD = 3 };

typedef struct pcb_detect_args {
  const uint8_t* frames_dev;   /* [n][h][w][3] BGR uint8 */
  int32_t n, h, w;
  int32_t S;                   /* square detector input, multiple of 32 */
  float det_thresh;
  int32_t rot_deg;             /* 0|90|180|270: detect on cv2.rotate(frame) without materialising it */
  int32_t pad_replicate;       /* replicate border of this many px around the (rotated) frame */
  int32_t fix_mode;            /* PCB_FIX_*: transform applied to dets before _accumulate */
  float fix_scale_inv;         /* PCB_FIX_SCALE: boxes/kps *= fix_scale_inv (frame given is pre-scaled) */
  int32_t orig_h, orig_w;      /* H0, W0 of the frame faces are reported in (== h,w unless SCALE) */
  int32_t min_box_px;          /* scrfd_min_box_px (8) */
  int32_t max_det;             /* capacity per frame of the output arrays */
  /* outputs (device) */
  float* det_dev;              /* [n][max_det][5]  x1,y1,x2,y2,score in (rotated/padded) image px, NMS order */
  float* kps_dev;              /* [n][max_det][10] */
  int32_t* raw_count_dev;      /* [n] detections after NMS (len(bboxes) in the reference) */
  int32_t* acc_box_dev;        /* [n][max_det][4] int boxes in original-frame px after _accumulate + min-size */
  float* acc_kps_dev;          /* [n][max_det][10] crop-local landmarks */
  float* acc_score_dev;        /* [n][max_det] */
  int32_t* acc_count_dev;      /* [n] */
  int32_t* acc_unfiltered_dev; /* [n] optional (may be NULL): entries that survived _accumulate before the
                                  min-size filter -- the reference breaks its scale-TTA loop on this list */
} pcb_detect_args;
int pcb_detect(pcb_ctx* ctx, const pcb_detect_args* a);

/* K1 alone: letterbox-resize + normalise + stem patch tensor.  out_dev is fp16
 * [n][S/2+pad][S/2+pad][32] (P-layout, see pcb_layout_pad). det_img_dev (optional, may be NULL) receives
 * the uint8 letterboxed image [n][S][S][3] for parity tests. */
int pcb_letterbox(pcb_ctx* ctx, const uint8_t* frames_dev, int n, int h, int w, int S, int rot_deg,
                  int pad_replicate, void* out_dev, uint8_t* det_img_dev);
/* K3 alone on explicit head maps (P-layout [n][S/s+pad][S/s+pad][32] for s=8,16,32; see pcb_layout_pad). */
int pcb_decode_nms(pcb_ctx* ctx, const void* head8_dev, const void* head16_dev, const void* head32_dev,
                   const float* reg_scale3_host, const pcb_detect_args* a, float det_scale);

/* ---- K4: cross-pass suppression + alignment + quality (replaces face_embedder.py:2439-2463:
 *      _iou suppression, crop, _canon_5pts, _align_by_5pts / _upright_by_eye_roll / resize,
 *      _face_quality) ----------------------------------------------------------------------- */
typedef struct pcb_align_args {
  const uint8_t* frames_dev;   /* [n][h][w][3] */
  int32_t n, h, w;
  int32_t max_det;
  const int32_t* acc_box_dev; const float* acc_kps_dev; const float* acc_score_dev; const int32_t* acc_count_dev;
  int32_t max_faces;           /* capacity of the outputs below */
  /* outputs (device) */
  int32_t* face_count_dev;     /* [n] faces kept per frame */
  int32_t* face_total_dev;     /* [1] */
  int32_t* face_frame_dev;     /* [max_faces] frame index of each face (faces are grouped by frame, kept order) */
  int32_t* face_box_dev;       /* [max_faces][4] xi1,yi1,xi2,yi2 */
  int32_t* face_kind_dev;      /* [max_faces] 0 align, 1 eye-roll+align, 2 eye-roll+resize, 3 resize */
  uint8_t* chips_dev;          /* [max_faces][112][112][3] BGR */
  double* quality_dev;         /* [max_faces] */
} pcb_align_args;
int pcb_align(pcb_ctx* ctx, const pcb_align_args* a);

/* ---- ArcFace (replaces _arcface_preprocess + arc_sess.run, face_embedder.py:1281-1288, 1369) -- */
/* chips uint8 [f][112][112][3] BGR -> raw embeddings e(x) [f][512] and, if emb_flip_dev != NULL,
 * e(flip x) [f][512] (cv2.flip(chip, 1), face_embedder.py:1297-1298).  emb_dev may be NULL when only
 * e(flip x) is wanted (the reference computes the flip pass only while a span is active, :1295). */
int pcb_embed(pcb_ctx* ctx, const uint8_t* chips_dev, int f, float* emb_dev, float* emb_flip_dev);

/* ---- K5: bank distance (replaces face_embedder.py:1383-1389 + Processor._fd_min,
 *      gui_app.py:660-674) -------------------------------------------------------------------- */
int pcb_set_bank(pcb_ctx* ctx, const float* bank_host /* [rows][512] */, int rows);
/* feat = normalise(emb + (use_flip[i] ? emb_flip : 0)); sim = max_j(bank_j . feat) in float32, so the
 * caller forms fd = 1.0 - sim exactly as _fd_min does; sim = -8 (fd = 9.0, the reference's sentinel)
 * when the bank is empty.  emb_flip_dev NULL: no flip; use_flip_dev NULL with emb_flip_dev set: flip
 * everywhere.  feat_dev / argmax_dev may be NULL. */
int pcb_match(pcb_ctx* ctx, const float* emb_dev, const float* emb_flip_dev, const uint8_t* use_flip_dev,
              int f, float* feat_dev, float* sim_dev, int32_t* argmax_dev);

/* ---- live reference bank (replaces Processor._stream_ref_bank_update, gui_app.py:922-986, for the replay below) ----
 * Host object: rows are unit fp32 vectors.  pcb_bank_offer applies the reference's streaming rule -- skip a zero vector,
 * "dup" when max cosine >= dedup, append while rows < cap, else replace the worst row when
 * s_new > s_worst + margin with s = wa*(1-fd_anchor) + wd*(1-nn_sim) (+ wq*min(max(q,0),1000)/300 for the candidate). */
typedef struct pcb_bank pcb_bank;
typedef struct pcb_bank_cfg {
  int32_t cap;                 /* prescan_bank_max */
  double dedup, margin;        /* prescan_diversity_dedup_cos, prescan_replace_margin */
  double wa, wd, wq;           /* prescan_weights */
} pcb_bank_cfg;
enum { PCB_BANK_SKIP = 0, PCB_BANK_ADDED = 1, PCB_BANK_DUP = 2, PCB_BANK_REPLACED = 3 };
pcb_bank* pcb_bank_create(const pcb_bank_cfg* cfg, const float* rows_host /* [n][512] unit rows or NULL */, int n);
void pcb_bank_destroy(pcb_bank* b);
/* -> PCB_BANK_*; *slot_out (may be NULL) = the row that was written for ADDED / REPLACED */
int pcb_bank_offer(pcb_bank* b, const float* vec /* [512], any norm */, double quality, int32_t* slot_out);
int pcb_bank_rows(const pcb_bank* b);
long long pcb_bank_version(const pcb_bank* b);      /* number of ADDED + REPLACED so far */
const float* pcb_bank_data(const pcb_bank* b);      /* [rows][512], valid until the next offer */

/* ---- live distances: the cosine of every face-table row to the live bank, kept current on the GPU while the replay
 *      changes the bank (the per-face Processor._fd_min calls of gui_app.py:1512-1549, batched).  Arithmetic per
 *      (row, bank row) pair is pcb_match's, so distances equal what pcb_match returns for the same bank. ---- */
/* feats_dev: [rows][512] unit features (pcb_match's feat output); copied in normalised form, so the caller may free it. */
int pcb_live_begin(pcb_ctx* ctx, const float* feats_dev, int rows);
/* Brings sim[r] = max_j bank_j . feat_r up to date for r >= row_lo of each of the `segments` equal parts of the table
 * (the replay keeps [plain; flip] halves and only ever needs rows of samples it has not passed yet) and copies them to
 * the host: synchronous.  changed_slot >= 0: only that bank row differs from the previous call (one dot product per
 * face row; rows whose argmax was that slot are recomputed in full); -1: the whole bank is new.
 * *sim_host_out: pinned [rows] float, entries outside the refreshed range are stale.  Empty bank: sim = -8 (fd 9.0). */
int pcb_live_refresh(pcb_ctx* ctx, const float* bank_host, int bank_rows, int changed_slot, int row_lo, int segments,
                     const float** sim_host_out);

/* ---- host replay of the pre-scan state machine (replaces the per-sample body of Processor._prescan,
 *      gui_app.py:1468-1655, for samples whose detections / features were precomputed in batches) -------- */
#define PCB_REPLAY_META 10
/* meta[s] = { up_start (-1: no upright face), up_count, hits90, hits270, heavy_raw90, heavy_raw270,
 *             heavy90_start (-1: none), heavy90_count, heavy270_start, heavy270_count }; rows index the face table */
typedef struct pcb_replay_cfg {
  double enter, exit_thr, fd_add, quality_min;
  int64_t total_frames, pad, min_len, exit_cool;
  int32_t stride, cooldown, fd9_skip, fd9_grace, fd9_period;
} pcb_replay_cfg;
typedef struct pcb_replay_state {      /* FaceEmbedder counters the reference advances per extract() call */
  int64_t frame_idx, last_face_idx;
  int32_t no_face_streak, rot_cycle, prescan_rr, trk_active;
} pcb_replay_state;
/* refresh: the bank changed (changed_slot >= 0: that row only; -1: everything, e.g. after new flip features) -- the callee
 * rewrites fd_plain / fd_flip for rows >= row_lo in place.  Used when ctx is NULL (CPU tests, custom matchers); with a
 * context the library refreshes through pcb_live_refresh itself and the callback may be NULL.
 * need_flip: flip-TTA features of sample s (and a look-ahead the callee chooses) are missing; the callee computes them,
 * sets flip_ready and the host features, and -- with a context -- calls pcb_live_begin again.
 * Both return 0 on success; a negative value aborts the replay (pcb_replay then returns PCB_REPLAY_ABORTED = 5). */
typedef int (*pcb_replay_refresh_cb)(void* user, const float* bank_rows, int n_rows, int changed_slot, int row_lo);
typedef int (*pcb_replay_flip_cb)(void* user, int sample);
#define PCB_REPLAY_ABORTED 5
typedef struct pcb_replay_io {
  const int32_t* meta; const int64_t* frame_idx; int32_t n_samples;
  int32_t n_rows;                          /* face-table rows */
  const double* quality; const int64_t* area;
  const uint8_t* flip_ready;               /* [n_rows] or NULL: all flip features present */
  const float* feat_plain; const float* feat_flip;   /* host [n_rows][512]: what a bank offer appends (flip while a span is active);
                                                         both NULL (ctx required): rows are read from the device tables below */
  double* fd_plain; double* fd_flip;       /* [n_rows] host distances, kept current by the refresh */
  pcb_replay_refresh_cb refresh; pcb_replay_flip_cb need_flip; void* user;
  /* outputs */
  double* best_out; uint8_t* skip_out; uint8_t* active_out; int32_t* nfaces_out;
  int64_t* spans_out; int32_t max_spans; int32_t* n_spans_out;
  int64_t* refreshes_out;                  /* may be NULL: distance refreshes performed */
  const float* feat_plain_dev; const float* feat_flip_dev;   /* device [n_rows][512], used when the host tables are NULL: a 2 KB
                                               * read per offer that is not a certain duplicate (tens per pre-scan) instead of
                                               * a device->host copy of every feature */
} pcb_replay_io;
/* ctx may be NULL (no GPU: distances come from io->refresh).  With ctx, pcb_live_begin must have been called on the
 * [plain; flip] table (2 * n_rows rows) and fd_plain / fd_flip are filled by the library before the first sample. */
int pcb_replay(pcb_ctx* ctx, const pcb_replay_cfg* cfg, pcb_bank* bank, const pcb_replay_io* io, pcb_replay_state* st);

#ifdef __cplusplus
}
#endif
#endif /* PCB200_H */
