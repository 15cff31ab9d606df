// K5: fused (sum flip pair) + L2-normalise + bank cosine + max/argmax, and the live distance table of the pre-scan replay.
// Replaces person_capture/face_embedder.py:1383-1389 (f = e(x) [+ e(flip x)]; f /= max(|f|, 1e-6))
// and Processor._fd_min, person_capture/gui_app.py:660-674 (vec / max(|vec|, 1e-6); 1 - max(bank @ vec)).
// HBM/L2-bound on the bank read (B*512*4 bytes per face block); no tensor cores: 2*F*B*512 FLOP
// is negligible next to the convolutions.
//
// One arithmetic for every path: a (face row, bank row) cosine is always computed by ONE warp with `row_dot` below
// (explicit fma chain per lane, xor-butterfly sum), so pcb_match, the full live refresh and the one-row incremental
// refresh give bit-identical similarities -- max(old, dot(new row)) IS the full recomputation.
#include <stdlib.h>
#include <string.h>

#include "pcb_common.cuh"

namespace {

constexpr int kD = PCB_FEAT_DIM;

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// cosine of one bank row and one (normalised) face row: lane l owns elements 4*(l + 32*j) .. +3, j = 0..3
__device__ __forceinline__ float row_dot(const float4* __restrict__ b4, const float4* v4, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 a = __ldg(b4 + lane + 32 * j);
    const float4 q = v4[lane + 32 * j];
    acc = __fmaf_rn(a.x, q.x, acc);
    acc = __fmaf_rn(a.y, q.y, acc);
    acc = __fmaf_rn(a.z, q.z, acc);
    acc = __fmaf_rn(a.w, q.w, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// 256 threads: x0/x1 are this thread's two elements of the face vector.  Normalises as the reference does twice
// (_arcface_encode, then _fd_min on the already-unit feature) and leaves the result in v[512] (shared).
__device__ __forceinline__ void normalise_twice(float x0, float x1, float* v, float* red, float* feat_out_row) {
  float nrm = sqrtf(block_sum(x0 * x0 + x1 * x1, red));
  nrm = fmaxf(nrm, 1e-6f);
  x0 /= nrm; x1 /= nrm;
  if (feat_out_row) {
    feat_out_row[threadIdx.x] = x0;
    feat_out_row[256 + threadIdx.x] = x1;
  }
  float n2 = sqrtf(block_sum(x0 * x0 + x1 * x1, red));
  n2 = fmaxf(n2, 1e-6f);
  v[threadIdx.x] = x0 / n2;
  v[256 + threadIdx.x] = x1 / n2;
  __syncthreads();
}

// best bank row for the face vector in shared memory: 8 warps take rows r = warp, warp + 8, ...; first occurrence wins
__device__ __forceinline__ void bank_scan(const float* v, const float* __restrict__ bank, int rows, float* wmax, int* warg,
                                          float* sim_out, int* arg_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = -3.0e38f;
  int barg = -1;
  const float4* v4 = (const float4*)v;
  for (int r = warp; r < rows; r += 8) {
    const float acc = row_dot((const float4*)(bank + (size_t)r * kD), v4, lane);
    if (acc > best) { best = acc; barg = r; }   // rows ascending within a warp
  }
  if (lane == 0) { wmax[warp] = best; warg[warp] = barg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -3.0e38f;
    int a = -1;
    for (int i = 0; i < 8; ++i)
      if (warg[i] >= 0 && (wmax[i] > m || (wmax[i] == m && warg[i] < a))) { m = wmax[i]; a = warg[i]; }
    // empty bank: sim = -8 so that fd = 1 - sim = 9.0, the reference's sentinel (gui_app.py:662-673)
    *sim_out = a >= 0 ? m : -8.0f;
    if (arg_out) *arg_out = a;
  }
}

__global__ void __launch_bounds__(256) match_kernel(const float* __restrict__ emb, const float* __restrict__ emb_flip,
                                                    const uint8_t* __restrict__ use_flip, int f, const float* __restrict__ bank,
                                                    int rows, float* __restrict__ feat_out, float* __restrict__ sim_out,
                                                    int* __restrict__ arg_out) {
  __shared__ __align__(16) float v[kD];
  __shared__ float red[8];
  __shared__ float wmax[8];
  __shared__ int warg[8];
  const int face = blockIdx.x;
  if (face >= f) return;
  const bool fl = emb_flip != nullptr && (use_flip == nullptr || use_flip[face] != 0);
  float x0 = emb[(size_t)face * kD + threadIdx.x];
  float x1 = emb[(size_t)face * kD + 256 + threadIdx.x];
  if (fl) {
    x0 += emb_flip[(size_t)face * kD + threadIdx.x];
    x1 += emb_flip[(size_t)face * kD + 256 + threadIdx.x];
  }
  normalise_twice(x0, x1, v, red, feat_out ? feat_out + (size_t)face * kD : nullptr);
  bank_scan(v, bank, rows, wmax, warg, sim_out + face, arg_out ? arg_out + face : nullptr);
}

// Few faces x large bank (the lock-face ROI site against a 10 000-row bank: one face per frame): match_kernel would scan the
// whole bank with ONE block.  Here block (face, seg) scans rows [seg * seg_rows, ...) with the same row_dot arithmetic and
// leaves its (max, first row) in part[face][seg]; match_split_reduce_kernel picks the maximum with the smallest row index,
// which is match_kernel's rule -- results are bit-identical.
__global__ void __launch_bounds__(256) match_split_kernel(const float* __restrict__ emb, const float* __restrict__ emb_flip,
                                                          const uint8_t* __restrict__ use_flip, int f, const float* __restrict__ bank,
                                                          int rows, int seg_rows, float* __restrict__ feat_out,
                                                          float* __restrict__ part_sim, int* __restrict__ part_arg) {
  __shared__ __align__(16) float v[kD];
  __shared__ float red[8];
  __shared__ float wmax[8];
  __shared__ int warg[8];
  const int face = blockIdx.x, seg = blockIdx.y;
  const bool fl = emb_flip != nullptr && (use_flip == nullptr || use_flip[face] != 0);
  float x0 = emb[(size_t)face * kD + threadIdx.x];
  float x1 = emb[(size_t)face * kD + 256 + threadIdx.x];
  if (fl) {
    x0 += emb_flip[(size_t)face * kD + threadIdx.x];
    x1 += emb_flip[(size_t)face * kD + 256 + threadIdx.x];
  }
  normalise_twice(x0, x1, v, red, (feat_out && seg == 0) ? feat_out + (size_t)face * kD : nullptr);
  const int r0 = seg * seg_rows;
  const int n = rows - r0 < seg_rows ? rows - r0 : seg_rows;
  float sim;
  int arg;
  bank_scan(v, bank + (size_t)r0 * kD, n, wmax, warg, &sim, &arg);
  if (threadIdx.x == 0) {
    part_sim[face * gridDim.y + seg] = sim;
    part_arg[face * gridDim.y + seg] = arg >= 0 ? arg + r0 : -1;
  }
}

__global__ void match_split_reduce_kernel(const float* __restrict__ part_sim, const int* __restrict__ part_arg, int segs,
                                          float* __restrict__ sim_out, int* __restrict__ arg_out) {
  const int face = blockIdx.x;
  float m = -3.0e38f;
  int a = -1;
  for (int s = 0; s < segs; ++s) {                       // segments ascend in row index: strict > keeps the first occurrence
    const int as = part_arg[face * segs + s];
    const float ms = part_sim[face * segs + s];
    if (as >= 0 && ms > m) { m = ms; a = as; }
  }
  sim_out[face] = a >= 0 ? m : -8.0f;
  if (arg_out) arg_out[face] = a;
}

// live table: V[r] = the vector match_kernel would hold in shared memory for feature row r
__global__ void __launch_bounds__(256) live_prepare_kernel(const float* __restrict__ feats, float* __restrict__ V, int rows) {
  __shared__ __align__(16) float v[kD];
  __shared__ float red[8];
  const int r = blockIdx.x;
  if (r >= rows) return;
  normalise_twice(feats[(size_t)r * kD + threadIdx.x], feats[(size_t)r * kD + 256 + threadIdx.x], v, red, nullptr);
  V[(size_t)r * kD + threadIdx.x] = v[threadIdx.x];
  V[(size_t)r * kD + 256 + threadIdx.x] = v[256 + threadIdx.x];
}

// rows [row_lo, seg) of each of `segments` parts of V; block b handles row index b of that compacted range
__device__ __forceinline__ int live_row(int b, int row_lo, int seg) {
  const int per = seg - row_lo;
  return (b / per) * seg + row_lo + (b % per);
}

__global__ void __launch_bounds__(256) live_full_kernel(const float* __restrict__ V, const float* __restrict__ bank, int bank_rows,
                                                        int row_lo, int seg, float* __restrict__ sim, int* __restrict__ arg) {
  __shared__ __align__(16) float v[kD];
  __shared__ float wmax[8];
  __shared__ int warg[8];
  const int r = live_row(blockIdx.x, row_lo, seg);
  v[threadIdx.x] = V[(size_t)r * kD + threadIdx.x];
  v[256 + threadIdx.x] = V[(size_t)r * kD + 256 + threadIdx.x];
  __syncthreads();
  bank_scan(v, bank, bank_rows, wmax, warg, sim + r, arg + r);
}

// one bank row (slot k) changed: sim = max(sim, dot(row k)) -- unless k WAS the argmax (replaced row), then rescan.
// One warp per face row.
__global__ void __launch_bounds__(256) live_update_kernel(const float* __restrict__ V, const float* __restrict__ bank, int bank_rows,
                                                          int k, int row_lo, int seg, int n_work, float* __restrict__ sim,
                                                          int* __restrict__ arg) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= n_work) return;
  const int r = live_row(w, row_lo, seg);
  const float4* v4 = (const float4*)(V + (size_t)r * kD);
  const int a0 = arg[r];
  float m;
  int a;
  if (a0 == k || a0 < 0) {
    m = -3.0e38f;
    a = -1;
    for (int j = 0; j < bank_rows; ++j) {
      const float d = row_dot((const float4*)(bank + (size_t)j * kD), v4, lane);
      if (d > m) { m = d; a = j; }
    }
  } else {
    m = sim[r];
    a = a0;
    const float d = row_dot((const float4*)(bank + (size_t)k * kD), v4, lane);
    if (d > m || (d == m && k < a)) { m = d; a = k; }
  }
  if (lane == 0) {
    sim[r] = a >= 0 ? m : -8.0f;
    arg[r] = a;
  }
}

// ---- large banks (BASELINE config 4: 10 000 rows x thousands of faces): the per-face kernel above re-reads the whole bank from L2
// for every face (82 GB for 4 096 faces).  Here a block owns 32 faces (4 per warp, kept in registers in row_dot's lane layout)
// and streams the bank through shared memory once (16 rows per stage, double buffered with cp.async): 32x less L2 traffic,
// 64 FMA per 4 LDS.128.  Every (face, bank row) cosine is computed with row_dot's fma chain and its butterfly tree (the four
// partial sums of a warp are reduced by a transposing butterfly: same pairs, same order), so results are bit-identical to
// match_kernel's, including the first-occurrence argmax.
constexpr int kGemmFaces = 32, kGemmRows = 16;
constexpr size_t kGemmSmem = 2 * kGemmRows * PCB_FEAT_DIM * sizeof(float);

__global__ void __launch_bounds__(256) match_prep_kernel(const float* __restrict__ emb, const float* __restrict__ emb_flip,
                                                         const uint8_t* __restrict__ use_flip, int f, float* __restrict__ feat_out,
                                                         float* __restrict__ V) {
  __shared__ __align__(16) float v[kD];
  __shared__ float red[8];
  const int face = blockIdx.x;
  if (face >= f) return;
  const bool fl = emb_flip != nullptr && (use_flip == nullptr || use_flip[face] != 0);
  float x0 = emb[(size_t)face * kD + threadIdx.x];
  float x1 = emb[(size_t)face * kD + 256 + threadIdx.x];
  if (fl) {
    x0 += emb_flip[(size_t)face * kD + threadIdx.x];
    x1 += emb_flip[(size_t)face * kD + 256 + threadIdx.x];
  }
  normalise_twice(x0, x1, v, red, feat_out ? feat_out + (size_t)face * kD : nullptr);
  V[(size_t)face * kD + threadIdx.x] = v[threadIdx.x];
  V[(size_t)face * kD + 256 + threadIdx.x] = v[256 + threadIdx.x];
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;     // src-size 0: zero fill (rows past the bank's end)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(256) match_gemm_kernel(const float* __restrict__ V, int f, const float* __restrict__ bank, int rows,
                                                         float* __restrict__ sim_out, int* __restrict__ arg_out) {
  extern __shared__ __align__(16) float stage_dyn[];           // [2][kGemmRows][kD] = 64 KB (opt-in dynamic shared memory)
  float (*stage)[kGemmRows][kD] = (float (*)[kGemmRows][kD])stage_dyn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int face0 = blockIdx.x * kGemmFaces + warp * 4;
  // my 4 faces, 16 elements each: float4 index lane + 32 j (row_dot's layout)
  float4 q[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int face = face0 + a < f ? face0 + a : f - 1;       // clamp: surplus lanes compute a duplicate that is not stored
    const float4* v4 = (const float4*)(V + (size_t)face * kD);
#pragma unroll
    for (int j = 0; j < 4; ++j) q[a][j] = v4[lane + 32 * j];
  }
  const int my = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);   // the face whose total this lane ends up holding
  float best = -3.0e38f;
  int barg = -1;
  const int n_stages = (rows + kGemmRows - 1) / kGemmRows;
  auto load_stage = [&](int st) {
    const int r0 = st * kGemmRows;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&stage[st & 1][0][0]);
    for (int i = threadIdx.x; i < kGemmRows * (kD / 4); i += 256) {
      const int rr = i / (kD / 4), c4 = i - rr * (kD / 4);
      const bool in = r0 + rr < rows;
      cp_async16(dst0 + (uint32_t)(i * 16), bank + (size_t)(in ? r0 + rr : 0) * kD + c4 * 4, in);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_stage(0);
  for (int st = 0; st < n_stages; ++st) {
    if (st + 1 < n_stages) {
      load_stage(st + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int r0 = st * kGemmRows;
    const int nr = rows - r0 < kGemmRows ? rows - r0 : kGemmRows;
    for (int rr = 0; rr < nr; ++rr) {
      const float4* b4 = (const float4*)&stage[st & 1][rr][0];
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = b4[lane + 32 * j];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k] = __fmaf_rn(a.x, q[k][j].x, acc[k]);
          acc[k] = __fmaf_rn(a.y, q[k][j].y, acc[k]);
          acc[k] = __fmaf_rn(a.z, q[k][j].z, acc[k]);
          acc[k] = __fmaf_rn(a.w, q[k][j].w, acc[k]);
        }
      }
      // transposing butterfly: xor 16 leaves two faces per lane, xor 8 one; then the plain butterfly (same pairs as row_dot)
      const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
      const float s0 = hi16 ? acc[0] : acc[2], s1 = hi16 ? acc[1] : acc[3];      // what the partner keeps
      float k0 = hi16 ? acc[2] : acc[0], k1 = hi16 ? acc[3] : acc[1];            // what I keep
      k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
      k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      const float s = hi8 ? k0 : k1;
      float t = hi8 ? k1 : k0;
      t += __shfl_xor_sync(0xffffffffu, s, 8);
      t += __shfl_xor_sync(0xffffffffu, t, 4);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      if (t > best) { best = t; barg = r0 + rr; }       // rows ascending: first occurrence wins
    }
    __syncthreads();      // the stage is free for the load after next
  }
  if ((lane & 7) == 0 && face0 + my < f) {
    sim_out[face0 + my] = barg >= 0 ? best : -8.0f;
    if (arg_out) arg_out[face0 + my] = barg;
  }
}

// device bank with room for `rows` rows (geometric growth; the old buffer is released once the stream has drained)
int bank_reserve(pcb_ctx* c, int rows) {
  if (rows <= c->bank_cap) return PCB_OK;
  int cap = c->bank_cap < 64 ? 64 : c->bank_cap;
  while (cap < rows) cap *= 2;
  float* nb = (float*)pcb_dev_alloc(c, (size_t)cap * kD * sizeof(float), false);
  if (!nb) return pcb_fail(c, PCB_ERR_CUDA, "set_bank: alloc failed");
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->bank_rows > 0 && c->bank)
    PCB_CUDA(c, cudaMemcpyAsync(nb, c->bank, (size_t)c->bank_rows * kD * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  pcb_dev_free(c, c->bank);
  if (c->bank_stage) cudaFreeHost(c->bank_stage);
  c->bank_stage = nullptr;
  PCB_CUDA(c, cudaMallocHost((void**)&c->bank_stage, (size_t)cap * kD * sizeof(float)));
  if (!c->bank_ev) PCB_CUDA(c, cudaEventCreateWithFlags(&c->bank_ev, cudaEventDisableTiming));
  c->bank = nb;
  c->bank_cap = cap;
  return PCB_OK;
}

}  // namespace

extern "C" int pcb_set_bank(pcb_ctx* c, const float* bank_host, int rows) {
  PCB_ENTER(c);
  if (rows < 0 || (rows > 0 && !bank_host)) return pcb_fail(c, PCB_ERR_ARG, "set_bank: bad arguments");
  int rc = bank_reserve(c, rows);
  if (rc) return rc;
  if (rows > 0) {
    // stage through pinned memory so the upload is asynchronous (the caller may reuse its buffer immediately); only
    // the PREVIOUS upload is waited for before the stage is overwritten
    PCB_CUDA(c, cudaEventSynchronize(c->bank_ev));
    memcpy(c->bank_stage, bank_host, (size_t)rows * kD * sizeof(float));
    PCB_CUDA(c, cudaMemcpyAsync(c->bank, c->bank_stage, (size_t)rows * kD * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PCB_CUDA(c, cudaEventRecord(c->bank_ev, c->stream));
  }
  c->bank_rows = rows;
  c->live_bank_rows = -1;      // the live table's similarities no longer describe the device bank
  return PCB_OK;
}

extern "C" int pcb_match(pcb_ctx* c, const float* emb_dev, const float* emb_flip_dev, const uint8_t* use_flip_dev, int f,
                         float* feat_dev, float* sim_dev, int32_t* argmax_dev) {
  PCB_ENTER(c);
  if (f < 0 || (f > 0 && (!emb_dev || !sim_dev))) return pcb_fail(c, PCB_ERR_ARG, "match: bad arguments");
  if (f == 0) return PCB_OK;
  // large bank x many faces: the tiled kernel (bank streamed once per 32 faces); PCB_MATCH_GEMM_ROWS moves / disables the switch
  const char* env = getenv("PCB_MATCH_GEMM_ROWS");
  const int min_rows = env ? atoi(env) : 1024;
  if (min_rows > 0 && c->bank_rows >= min_rows && f >= 2 * kGemmFaces) {
    const size_t need = (size_t)f * kD * sizeof(float);
    if (c->match_v_bytes < need) {
      void* nb = pcb_dev_alloc(c, need, false);
      if (!nb) return pcb_fail(c, PCB_ERR_CUDA, "match: scratch alloc failed");
      if (c->match_v) { PCB_CUDA(c, cudaStreamSynchronize(c->stream)); pcb_dev_free(c, c->match_v); }
      c->match_v = (float*)nb;
      c->match_v_bytes = need;
    }
    match_prep_kernel<<<f, 256, 0, c->stream>>>(emb_dev, emb_flip_dev, use_flip_dev, f, feat_dev, c->match_v);
    PCB_LAUNCH_CHECK(c, "match_prep_kernel");
    static unsigned long long attr_devs = 0;
    if (pcb_attr_needed(&attr_devs, c->device))
      PCB_CUDA(c, cudaFuncSetAttribute(match_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    match_gemm_kernel<<<(f + kGemmFaces - 1) / kGemmFaces, 256, kGemmSmem, c->stream>>>(c->match_v, f, c->bank, c->bank_rows, sim_dev, argmax_dev);
    PCB_LAUNCH_CHECK(c, "match_gemm_kernel");
    return PCB_OK;
  }
  if (min_rows > 0 && c->bank_rows >= min_rows) {
    // few faces: split the bank over enough blocks to fill the GPU (>= 128 rows per block)
    int segs = (2 * c->num_sms + f - 1) / f;
    const int max_segs = (c->bank_rows + 127) / 128;
    if (segs > max_segs) segs = max_segs;
    if (segs >= 2) {
      const int seg_rows = (c->bank_rows + segs - 1) / segs;
      segs = (c->bank_rows + seg_rows - 1) / seg_rows;
      const size_t need = (size_t)f * segs * 2 * sizeof(float);
      if (c->match_v_bytes < need) {
        void* nb = pcb_dev_alloc(c, need < 65536 ? 65536 : need, false);
        if (!nb) return pcb_fail(c, PCB_ERR_CUDA, "match: scratch alloc failed");
        if (c->match_v) { PCB_CUDA(c, cudaStreamSynchronize(c->stream)); pcb_dev_free(c, c->match_v); }
        c->match_v = (float*)nb;
        c->match_v_bytes = need < 65536 ? 65536 : need;
      }
      float* part_sim = c->match_v;
      int* part_arg = (int*)(c->match_v + (size_t)f * segs);
      match_split_kernel<<<dim3(f, segs), 256, 0, c->stream>>>(emb_dev, emb_flip_dev, use_flip_dev, f, c->bank, c->bank_rows, seg_rows,
                                                              feat_dev, part_sim, part_arg);
      PCB_LAUNCH_CHECK(c, "match_split_kernel");
      match_split_reduce_kernel<<<f, 1, 0, c->stream>>>(part_sim, part_arg, segs, sim_dev, argmax_dev);
      PCB_LAUNCH_CHECK(c, "match_split_reduce_kernel");
      return PCB_OK;
    }
  }
  match_kernel<<<f, 256, 0, c->stream>>>(emb_dev, emb_flip_dev, use_flip_dev, f, c->bank, c->bank_rows, feat_dev, sim_dev, argmax_dev);
  PCB_LAUNCH_CHECK(c, "match_kernel");
  return PCB_OK;
}

// replay.cu: the feature row a bank offer appends, read from the device table on demand (a handful per pre-scan: the replay
// skips certain duplicates without looking at the vector)
int pcb_fetch_row(pcb_ctx* c, const float* row_dev, float* out_host) {
  PCB_ENTER(c);
  if (!c->live_row_stage) PCB_CUDA(c, cudaMallocHost((void**)&c->live_row_stage, kD * sizeof(float)));
  // live_row_stage: every user (this, pcb_live_refresh's bank row upload) synchronises the stream before returning
  PCB_CUDA(c, cudaMemcpyAsync(c->live_row_stage, row_dev, kD * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(out_host, c->live_row_stage, kD * sizeof(float));
  return PCB_OK;
}

extern "C" int pcb_live_begin(pcb_ctx* c, const float* feats_dev, int rows) {
  PCB_ENTER(c);
  if (rows < 0 || (rows > 0 && !feats_dev)) return pcb_fail(c, PCB_ERR_ARG, "live_begin: bad arguments");
  if (rows > c->live_cap) {
    int cap = c->live_cap < 1024 ? 1024 : c->live_cap;
    while (cap < rows) cap *= 2;
    PCB_CUDA(c, cudaStreamSynchronize(c->stream));
    pcb_dev_free(c, c->live_v); pcb_dev_free(c, c->live_sim); pcb_dev_free(c, c->live_arg);
    if (c->live_sim_host) cudaFreeHost(c->live_sim_host);
    c->live_sim_host = nullptr;
    c->live_cap = 0;
    c->live_v = (float*)pcb_dev_alloc(c, (size_t)cap * kD * sizeof(float), false);
    c->live_sim = (float*)pcb_dev_alloc(c, (size_t)cap * sizeof(float), false);
    c->live_arg = (int*)pcb_dev_alloc(c, (size_t)cap * sizeof(int), false);
    if (!c->live_v || !c->live_sim || !c->live_arg) return pcb_fail(c, PCB_ERR_CUDA, "live_begin: alloc failed");
    PCB_CUDA(c, cudaMallocHost((void**)&c->live_sim_host, (size_t)cap * sizeof(float)));
    c->live_cap = cap;
  }
  if (!c->live_row_stage) PCB_CUDA(c, cudaMallocHost((void**)&c->live_row_stage, kD * sizeof(float)));
  c->live_rows = rows;
  c->live_bank_rows = -1;
  if (rows > 0) {
    live_prepare_kernel<<<rows, 256, 0, c->stream>>>(feats_dev, c->live_v, rows);
    PCB_LAUNCH_CHECK(c, "live_prepare_kernel");
  }
  return PCB_OK;
}

extern "C" int pcb_live_refresh(pcb_ctx* c, const float* bank_host, int bank_rows, int changed_slot, int row_lo, int segments,
                                const float** sim_host_out) {
  PCB_ENTER(c);
  if (bank_rows < 0 || (bank_rows > 0 && !bank_host) || segments < 1 || c->live_rows % segments || changed_slot >= bank_rows)
    return pcb_fail(c, PCB_ERR_ARG, "live_refresh: bad arguments");
  if (sim_host_out) *sim_host_out = c->live_sim_host;
  const int seg = c->live_rows / segments;
  if (row_lo < 0) row_lo = 0;
  if (seg == 0 || row_lo >= seg) return PCB_OK;
  const int n_work = (seg - row_lo) * segments;
  // the incremental form needs the previous state to describe exactly the bank minus this row
  const bool incremental = changed_slot >= 0 && c->live_bank_rows >= 0 &&
                           (c->live_bank_rows == bank_rows || (c->live_bank_rows == bank_rows - 1 && changed_slot == bank_rows - 1));
  if (incremental) {
    int rc = bank_reserve(c, bank_rows);
    if (rc) return rc;
    // one 2 KB row; the previous refresh ended with a stream synchronisation, so the stage is free
    memcpy(c->live_row_stage, bank_host + (size_t)changed_slot * kD, kD * sizeof(float));
    PCB_CUDA(c, cudaMemcpyAsync(c->bank + (size_t)changed_slot * kD, c->live_row_stage, kD * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->bank_rows = bank_rows;
    live_update_kernel<<<(n_work + 7) / 8, 256, 0, c->stream>>>(c->live_v, c->bank, bank_rows, changed_slot, row_lo, seg, n_work,
                                                                 c->live_sim, c->live_arg);
    PCB_LAUNCH_CHECK(c, "live_update_kernel");
  } else {
    int rc = pcb_set_bank(c, bank_host, bank_rows);
    if (rc) return rc;
    live_full_kernel<<<n_work, 256, 0, c->stream>>>(c->live_v, c->bank, bank_rows, row_lo, seg, c->live_sim, c->live_arg);
    PCB_LAUNCH_CHECK(c, "live_full_kernel");
  }
  for (int s = 0; s < segments; ++s)
    PCB_CUDA(c, cudaMemcpyAsync(c->live_sim_host + (size_t)s * seg + row_lo, c->live_sim + (size_t)s * seg + row_lo,
                                (size_t)(seg - row_lo) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->live_bank_rows = bank_rows;
  return PCB_OK;
}
