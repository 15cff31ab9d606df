// K5: fused (sum flip pair) + L2-normalise + bank cosine + max/argmax.
// Replaces person_capture/face_embedder.py:1383-1389 (f = e(x) [+ e(flip x)]; f /= max(|f|, 1e-6))
// and Processor._fd_min, person_capture/gui_app.py:660-674 (vec / max(|vec|, 1e-6); 1 - max(bank @ vec)).
// HBM/L2-bound on the bank read (B*512*4 bytes per face block); no tensor cores: 2*F*B*512 FLOP
// is negligible next to the convolutions.
#include <string.h>

#include "pcb_common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) match_kernel(const float* __restrict__ emb, const float* __restrict__ emb_flip,
                                                    const uint8_t* __restrict__ use_flip, int f, const float* __restrict__ bank,
                                                    int rows, float* __restrict__ feat_out, float* __restrict__ sim_out,
                                                    int* __restrict__ arg_out) {
  __shared__ __align__(16) float v[PCB_FEAT_DIM];
  __shared__ float red[8];
  __shared__ float wmax[8];
  __shared__ int warg[8];
  const int face = blockIdx.x;
  if (face >= f) return;
  const bool fl = emb_flip != nullptr && (use_flip == nullptr || use_flip[face] != 0);
  float x0 = emb[(size_t)face * PCB_FEAT_DIM + threadIdx.x];
  float x1 = emb[(size_t)face * PCB_FEAT_DIM + 256 + threadIdx.x];
  if (fl) {
    x0 += emb_flip[(size_t)face * PCB_FEAT_DIM + threadIdx.x];
    x1 += emb_flip[(size_t)face * PCB_FEAT_DIM + 256 + threadIdx.x];
  }
  float nrm = sqrtf(block_sum(x0 * x0 + x1 * x1, red));
  nrm = fmaxf(nrm, 1e-6f);
  x0 /= nrm; x1 /= nrm;
  if (feat_out) {
    feat_out[(size_t)face * PCB_FEAT_DIM + threadIdx.x] = x0;
    feat_out[(size_t)face * PCB_FEAT_DIM + 256 + threadIdx.x] = x1;
  }
  // _fd_min renormalises the (already unit) feature
  float n2 = sqrtf(block_sum(x0 * x0 + x1 * x1, red));
  n2 = fmaxf(n2, 1e-6f);
  v[threadIdx.x] = x0 / n2;
  v[256 + threadIdx.x] = x1 / n2;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = -3.0e38f;
  int barg = -1;
  const float4* v4 = (const float4*)v;
  for (int r = warp; r < rows; r += 8) {
    const float4* b4 = (const float4*)(bank + (size_t)r * PCB_FEAT_DIM);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = __ldg(b4 + lane + 32 * j);
      const float4 q = v4[lane + 32 * j];
      acc += a.x * q.x + a.y * q.y + a.z * q.z + a.w * q.w;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (acc > best) { best = acc; barg = r; }   // first occurrence wins within a warp (rows ascending)
  }
  if (lane == 0) { wmax[warp] = best; warg[warp] = barg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -3.0e38f;
    int a = -1;
    for (int i = 0; i < 8; ++i)
      if (warg[i] >= 0 && (wmax[i] > m || (wmax[i] == m && warg[i] < a))) { m = wmax[i]; a = warg[i]; }
    // empty bank: sim = -8 so that fd = 1 - sim = 9.0, the reference's sentinel (gui_app.py:662-673)
    sim_out[face] = a >= 0 ? m : -8.0f;
    if (arg_out) arg_out[face] = a;
  }
}

}  // namespace

extern "C" int pcb_set_bank(pcb_ctx* c, const float* bank_host, int rows) {
  if (rows < 0 || (rows > 0 && !bank_host)) return pcb_fail(c, PCB_ERR_ARG, "set_bank: bad arguments");
  if (rows > c->bank_cap) {
    int cap = rows < 64 ? 64 : rows;
    float* nb = (float*)pcb_dev_alloc(c, (size_t)cap * PCB_FEAT_DIM * sizeof(float), false);
    if (!nb) return pcb_fail(c, PCB_ERR_CUDA, "set_bank: alloc failed");
    if (c->bank_ev) PCB_CUDA(c, cudaEventSynchronize(c->bank_ev));
    if (c->bank_stage) cudaFreeHost(c->bank_stage);
    c->bank_stage = nullptr;
    PCB_CUDA(c, cudaMallocHost((void**)&c->bank_stage, (size_t)cap * PCB_FEAT_DIM * sizeof(float)));
    if (!c->bank_ev) PCB_CUDA(c, cudaEventCreateWithFlags(&c->bank_ev, cudaEventDisableTiming));
    c->bank = nb;
    c->bank_cap = cap;
  }
  if (rows > 0) {
    // the live bank changes every few samples during a replay: stage through pinned memory so the upload is asynchronous
    // (the caller may reuse its buffer immediately) and only wait for the PREVIOUS upload before overwriting the stage
    PCB_CUDA(c, cudaEventSynchronize(c->bank_ev));
    memcpy(c->bank_stage, bank_host, (size_t)rows * PCB_FEAT_DIM * sizeof(float));
    PCB_CUDA(c, cudaMemcpyAsync(c->bank, c->bank_stage, (size_t)rows * PCB_FEAT_DIM * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PCB_CUDA(c, cudaEventRecord(c->bank_ev, c->stream));
  }
  c->bank_rows = rows;
  return PCB_OK;
}

extern "C" int pcb_match(pcb_ctx* c, const float* emb_dev, const float* emb_flip_dev, const uint8_t* use_flip_dev, int f,
                         float* feat_dev, float* sim_dev, int32_t* argmax_dev) {
  if (f < 0 || (f > 0 && (!emb_dev || !sim_dev))) return pcb_fail(c, PCB_ERR_ARG, "match: bad arguments");
  if (f == 0) return PCB_OK;
  match_kernel<<<f, 256, 0, c->stream>>>(emb_dev, emb_flip_dev, use_flip_dev, f, c->bank, c->bank_rows, feat_dev, sim_dev, argmax_dev);
  PCB_LAUNCH_CHECK(c, "match_kernel");
  return PCB_OK;
}
