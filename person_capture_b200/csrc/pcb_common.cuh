// Shared declarations for libpcb200 (sm_100a).  Internal header; the public C ABI is
// include/pcb200.h.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/pcb200.h"

#define PCB_OK 0
#define PCB_ERR_CUDA 1
#define PCB_ERR_ARG 2
#define PCB_ERR_KERNEL 3   // device-side watchdog / overflow flag raised
#define PCB_ERR_STATE 4

// ---------------------------------------------------------------------------------------
// Activation layout ("P-layout"): fp16 [N][H+kPad][W+kPad][Cp], Cp = channels rounded up to 8, image pixel (y, x) at
// [y+kPadLo][x+kPadLo], every other element zero.  A 3x3/pad-1 convolution over it is nine row-shifted views of the same
// [rows, Cp] matrix (rows = N*(H+kPad)*(W+kPad)): tap (dy,dx) of row r reads row r + dy*(W+kPad) + dx, which is what
// the TMA-fed implicit GEMM consumes.  Kernels only ever write image pixels, so the padding stays zero after the
// one-time memset at allocation.
//   kPadLo/kPad = 1/2 (default): a one-pixel ring of zeros around every image.
//   kPadLo/kPad = 0/1 (-DPCB_PAD_LO=0 -DPCB_PAD=1): ONE trailing zero column per image row and ONE trailing zero row per
//     image -- x = -1 is then the pad column that ends the previous row, y = -1 the pad row that ends the previous image,
//     rows before the tensor's first are TMA out-of-bounds zero fill.  It cuts the GEMM rows spent on padding from
//     (H+2)(W+2)/(HW) to (H+1)(W+1)/(HW) (1.31 -> 1.15 at 14x14, 1.65 -> 1.31 at 7x7) and passes every parity test, but
//     measured on B200 it is SLOWER overall (651-684 vs 732-761 TFLOP/s on the layer profile, same process, alternating
//     libraries): 7x7 and 16x16 layers gain 13-20 %, the 14x14 stage gains nothing at 444 images (391 pair tiles still
//     take 6 waves on 74 CTA pairs) and every wide-map layer loses 12-30 % in its MMA phase (tensor pipe 37 % active
//     instead of 49 % on the 112x112 64->64 layer, ncu) although tools/mma_rate.cu shows the tensor pipe itself is
//     indifferent to the row pitch -- an interaction not understood yet, so the ring stays the product layout.
// ---------------------------------------------------------------------------------------
#ifndef PCB_PAD_LO
#define PCB_PAD_LO 1
#define PCB_PAD 2
#endif
constexpr int kPadLo = PCB_PAD_LO;   // pad elements before the first row / column of an image
constexpr int kPad = PCB_PAD;        // pad elements per dimension in total
__host__ __device__ __forceinline__ long long pcb_prow(int img, int y, int x, int h, int w) {
  return ((long long)img * (h + kPad) + (y + kPadLo)) * (w + kPad) + (x + kPadLo);
}
// The layout is a property of the TENSOR (round 2): small maps of the ArcFace graph use the trailing-pad form (pad_lo 0, pad 1
// -- 225 instead of 256 GEMM rows per 14x14 image, 64 instead of 81 per 7x7 image), everything else the ring (pad_lo 1, pad 2).
// Round 1 measured the trailing pad as a GLOBAL switch: small maps +13-20 %, wide maps -12...-30 %, so it stayed off; the
// stride-2 convolutions that cross from 28x28 to 14x14 write the other layout for free (they address output rows themselves).
__host__ __device__ __forceinline__ long long pcb_prow_l(int img, int y, int x, int h, int w, int pad_lo, int pad) {
  return ((long long)img * (h + pad) + (y + pad_lo)) * (w + pad) + (x + pad_lo);
}
struct PTensor {
  __half* data = nullptr;
  int n = 0, h = 0, w = 0, c = 0, cp = 0;  // logical dims, cp = padded channel stride
  int pad_lo = kPadLo, pad = kPad;         // spatial padding of THIS tensor (see above)
  bool dense = false;                      // dense: [n][cp] rows without spatial padding (FC input)
  bool f32 = false;                        // elements are float (iResNet residual stream) instead of __half
  size_t rows() const { return dense ? (size_t)n : (size_t)n * (h + pad) * (w + pad); }
  size_t bytes() const { return rows() * cp * (f32 ? sizeof(float) : sizeof(__half)); }
};

static inline int pcb_round_up(int x, int m) { return (x + m - 1) / m * m; }

struct ConvWeights {
  // packed for the implicit GEMM: [npad][taps][cin_w] fp16 (zero padded), scale/bias/slope [npad] fp32
  __half* w = nullptr;
  float* scale = nullptr;
  float* bias = nullptr;
  float* slope = nullptr;
  int cin = 0, cout = 0, k = 0, taps = 0;
  int cin_w = 0;   // per-tap K extent in the packed weight (multiple of 64)
  int npad = 0;    // rows in the packed weight (multiple of n_tile)
  int rows_alloc = 0;  // rows actually allocated (npad rounded up to 128, zero filled)
  int n_tile = 0;  // UMMA N used for this layer
};

struct ConvArgs {
  const PTensor* in;
  const PTensor* out;       // fp16 P-layout output (or nullptr when out_f32 is used)
  float* out_f32;           // dense fp32 output [rows][out_f32_stride] (FC)
  int out_f32_stride;
  const PTensor* residual;  // optional, same geometry as out (fp16 or fp32)
  const PTensor* out2;      // optional second fp16 output: scale2 * y + bias2 (same geometry as out)
  const float* scale2;
  const float* bias2;
  const ConvWeights* w;
  int stride;               // 1 or 2 (spatial); ignored for dense
  int act;                  // 0 none, 1 relu, 2 prelu
};

struct pcb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;
  std::string last_error;
  int* d_err = nullptr;       // device error word (watchdog / overflow)
  int* h_err = nullptr;       // pinned mirror
  int conv_impl = 0;          // 0 = tcgen05 product (tc2), 1 = CUDA-core validation, 2 = tcgen05 baseline (tc), 3 = transposed experiment (tc3)
  long long launches = 0;     // kernels launched since the last pcb_reset_counters
  std::vector<void*> allocs;  // everything cudaMalloc'ed through the context
  struct Model* models[4] = {nullptr, nullptr, nullptr, nullptr};
  float* bank = nullptr;
  int bank_rows = 0, bank_cap = 0;
  float* bank_stage = nullptr;   // pinned staging for asynchronous bank uploads
  cudaEvent_t bank_ev = nullptr; // completion of the last bank upload
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // per-context scratch that used to live in function-static maps keyed by the context pointer (a recreated context
  // could inherit freed device pointers): K4 align buffers and the letterbox coefficient tables
  size_t align_sz[5] = {0, 0, 0, 0, 0};
  void* align_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  std::map<std::vector<int>, std::pair<void*, void*>> lin_tabs;
  // live distance table of the pre-scan replay (match.cu): normalised face rows, their best cosine / argmax against
  // the live bank on the device, and the pinned host mirror pcb_live_refresh hands to the replay
  float* match_v = nullptr;          // normalised face vectors of the tiled large-bank matcher
  size_t match_v_bytes = 0;
  float* live_v = nullptr;
  float* live_sim = nullptr;
  int* live_arg = nullptr;
  float* live_sim_host = nullptr;
  float* live_row_stage = nullptr;   // pinned [512]: the one bank row that changed
  int live_rows = 0, live_cap = 0;
  int live_bank_rows = 0;            // bank rows the device-side sims are current for (-1: never refreshed)
  // small ArcFace calls (lock-face ROI path: 1-2 faces per frame) replay a captured CUDA graph of the ~110 launches
  struct EmbedGraph { cudaGraphExec_t exec = nullptr; unsigned long long gen = 0; bool warmed = false; };
  std::map<int, EmbedGraph> embed_graphs;     // key = mode * 64 + faces
  uint8_t* embed_stage = nullptr;             // [8][112][112][3] fixed input of those graphs
  // profiling (bench.py roofline): CUDA events around every conv launch + algorithmic FLOPs
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<double> ev_flops;      // per recorded launch
  std::vector<std::string> ev_desc;  // per recorded launch: layer shape (PCB_PROFILE_DUMP csv)
  double prof_ms = 0.0, prof_flops = 0.0;
  long long prof_launches = 0;
};

// records an event pair around a conv launch when profiling is on
struct PcbConvTimer {
  pcb_ctx* c;
  cudaEvent_t e1 = nullptr;
  PcbConvTimer(pcb_ctx* ctx, double flops, const char* desc = nullptr);
  ~PcbConvTimer();
};

int pcb_fail(pcb_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess);
// Every extern "C" entry point that touches the device starts with this: a process may hold contexts on several GPUs
// (Engine(device=k) is public API) and the calling thread's current device is whatever the last call left behind.
#define PCB_ENTER(ctx)                                                           \
  do {                                                                           \
    if (!(ctx)) return PCB_ERR_ARG;                                              \
    cudaError_t _e0 = cudaSetDevice((ctx)->device);                              \
    if (_e0 != cudaSuccess) return pcb_fail((ctx), PCB_ERR_CUDA, "cudaSetDevice", _e0); \
  } while (0)
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember it per (function, device)
static inline bool pcb_attr_needed(unsigned long long* mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}
#define PCB_CUDA(ctx, call)                                                      \
  do {                                                                           \
    cudaError_t _e = (call);                                                     \
    if (_e != cudaSuccess) return pcb_fail((ctx), PCB_ERR_CUDA, #call, _e);      \
  } while (0)
#define PCB_LAUNCH_CHECK(ctx, name)                                              \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) return pcb_fail((ctx), PCB_ERR_CUDA, name, _e);       \
    (ctx)->launches++;                                                           \
  } while (0)

void* pcb_dev_alloc(pcb_ctx* c, size_t bytes, bool zero);
void pcb_dev_free(pcb_ctx* c, void* p);   // cudaFree + forget; the caller orders it after the last use on the stream

// conv_tc.cu / conv_simple.cu
int pcb_conv_tc2(pcb_ctx* c, const ConvArgs& a);   // product kernel: pixels on UMMA M, couts on N, operand reuse in shared memory
int pcb_conv_tc(pcb_ctx* c, const ConvArgs& a);    // first formulation, kept as the A/B baseline (impl 2)
int pcb_conv_simple(pcb_ctx* c, const ConvArgs& a);
// ops.cu
int pcb_op_affine(pcb_ctx* c, const PTensor& in, const PTensor& out, const float* scale, const float* bias);
int pcb_op_maxpool3s2(pcb_ctx* c, const PTensor& in, const PTensor& out);
int pcb_op_avgpool2(pcb_ctx* c, const PTensor& in, const PTensor& out);
int pcb_op_upsample_add(pcb_ctx* c, const PTensor& big, const PTensor& small, const PTensor& out);
int pcb_op_add(pcb_ctx* c, const PTensor& a, const PTensor& b, const PTensor& out);
int pcb_op_affine_flatten(pcb_ctx* c, const PTensor& in, const PTensor& out_dense, const float* scale, const float* bias);
