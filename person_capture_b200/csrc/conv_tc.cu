// K2: implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged by TMA).  Replaces the ONNX Runtime / TensorRT execution of the SCRFD
// and ArcFace graphs (reference: person_capture/face_embedder.py:1102-1107, 1341, 1369).
//
// Formulation.  Activations live in the P-layout (pcb_common.cuh): a [rows, Cp] fp16 matrix
// whose rows enumerate the zero-ringed pixels of the whole batch.  For a 3x3/pad-1 conv, tap
// (dy,dx) of output row p reads input row p + dy*(W+2) + dx, so each tap is the SAME 2-D
// tensor map at a row offset: no im2col buffer, no halo logic, and TMA's out-of-bounds
// zero-fill covers the first/last tile and ragged channel counts.  One CTA tile is
//   D[128 rows, n_tile] += A_tap,kc[128, 64] * W_tap,kc[n_tile, 64]^T   over taps x k-chunks,
// with A/W tiles in 128B-swizzled K-major shared memory, D (fp32) double-buffered in TMEM.
// Stride-2 convs are evaluated on the input grid and the epilogue keeps the even pixels.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles):
//   warp 0  TMA producer        warp 1  MMA issuer (+TMEM alloc)       warps 2-5  epilogue
// Epilogue: tcgen05.ld 16 columns at a time -> y = acc*scale+bias (+residual) -> ReLU/PReLU ->
// fp16 -> 16-byte stores of the interior pixels only (the zero ring is never written).
//
// Every mbarrier wait is bounded (watchdog): on timeout the kernel raises the context's error
// word and drains instead of hanging the GPU.
#include <stdio.h>

#include "pcb_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kKC = 64;                 // K elements per pipeline stage (one 128B swizzle atom)
constexpr int kABytes = kBlockM * kKC * 2;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;    // TMEM columns between the two accumulator buffers
constexpr int kMaxStages = 8;

struct ConvTcParams {
  int rows;        // input rows (N*(H+2)*(W+2)) or dense rows
  int hp, wp;      // input padded dims; 0 for dense
  int taps;        // 9 or 1
  int kchunks;     // ceil(cin_eff / 64)
  int cin_w;       // packed weight K extent per tap
  int n_tile, n_tiles, m_tiles;
  int stride;      // 1 | 2
  int hp_out, wp_out;
  int pad_lo, pad_hi, pad_lo_out;   // input padding before / after the pixels, output padding before (per-tensor layout)
  int out_cp;      // channel stride of the fp16 output
  int out_c_store; // channels to store (<= out_cp, multiple of 8)
  int act;
  int dense;
  int stages;
  const float* scale;
  const float* bias;
  const float* slope;
  const __half* residual;
  const float* residual_f32;   // residual stream kept in fp32 (same P-layout)
  int res_cp;
  __half* out;
  float* out_s32;              // primary output as fp32 P-layout (residual stream)
  __half* out2;                // second output: fp16(scale2 * y + bias2)
  int out2_cp;
  const float* scale2;
  const float* bias2;
  float* out_f32;
  int out_f32_stride;
  int out_f32_cols;
  int* err;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait.  Returns false (and raises *err) if the barrier did not flip in ~0.3 s or if
// another role already raised the error word.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)err != 0) return false;
      if (clock64() - t0 > 600000000LL) {
        atomicCAS(err, 0, code);
        return false;
      }
    }
  }
  return true;
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t lo = (uint64_t)((saddr >> 4) & 0x3fff);             // start address, LBO = 0
  uint64_t hi = (uint64_t)(1024 >> 4)                           // SBO
                | (1ull << 14)                                  // descriptor version 1 (sm_100)
                | (2ull << 29);                                 // layout type: SWIZZLE_128B
  return lo | (hi << 32);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB | B n_tile*128 B], then barriers
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.n_tile * kKC * 2;
  const int stage_bytes = kABytes + b_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int ksteps = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
        const int m0 = mt * kBlockM, n0 = nt * p.n_tile;
        for (int t = 0; ok && t < p.taps; ++t) {
          const int shift = (p.taps == 9) ? ((t / 3 - 1) * p.wp + (t % 3 - 1)) : 0;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 101)) { ok = false; break; }
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
            tma_load_2d(&tmA, &full_bar[stage], sa, kc * kKC, m0 + shift);
            tma_load_2d(&tmB, &full_bar[stage], sa + kABytes, t * p.cin_w + kc * kKC, n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4)                          // D format: F32
                             | (0u << 7) | (0u << 10)           // A, B format: F16
                             | ((uint32_t)(p.n_tile >> 3) << 17)
                             | ((uint32_t)(kBlockM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; ok && tile < total_tiles; tile += gridDim.x) {
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err, 102)) { ok = false; break; }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (!mbar_wait(&full_bar[stage], phase, p.err, 103)) { ok = false; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = make_desc_sw128(sa);
          const uint64_t db = make_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kKC / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom
            umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles, nt = tile - mt * p.n_tiles;
      const int n0 = nt * p.n_tile;
      const long long prow = (long long)mt * kBlockM + row_in_tile;
      bool valid = prow < p.rows;
      long long orow = prow;
      if (!p.dense) {
        const int plane = p.hp * p.wp;
        const int img = (int)(prow / plane);
        const int rem = (int)(prow - (long long)img * plane);
        const int y = rem / p.wp, x = rem - y * p.wp;
        valid = valid && y >= p.pad_lo && y <= p.hp - 1 - p.pad_hi && x >= p.pad_lo && x <= p.wp - 1 - p.pad_hi;
        if (p.stride == 2) {
          valid = valid && (((y - p.pad_lo) | (x - p.pad_lo)) & 1) == 0;
          orow = ((long long)img * p.hp_out + ((y - p.pad_lo) >> 1) + p.pad_lo_out) * p.wp_out + ((x - p.pad_lo) >> 1) + p.pad_lo_out;
        }
      }
      if (ok) ok = mbar_wait(&tfull_bar[acc], acc_phase, p.err, 104);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride;
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t v[16];
        __syncwarp();   // lanes diverge on `valid` below; tcgen05.ld is .sync.aligned
        tmem_ld16(t_row + c0, v);
        const int ch = n0 + c0;
        if (!valid) continue;
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = fmaf(__uint_as_float(v[j]), __ldg(p.scale + ch + j), __ldg(p.bias + ch + j));
        if (p.out_f32) {
          float* o = p.out_f32 + orow * p.out_f32_stride + ch;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (ch + j < p.out_f32_cols) *(float4*)(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          continue;
        }
        if (p.residual_f32) {
          const float* r = p.residual_f32 + orow * p.res_cp + ch;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (ch + q4 * 4 < p.out_c_store) {
              const float4 rv = *(const float4*)(r + q4 * 4);
              y[q4 * 4] += rv.x; y[q4 * 4 + 1] += rv.y; y[q4 * 4 + 2] += rv.z; y[q4 * 4 + 3] += rv.w;
            }
          }
        } else if (p.residual) {
          const __half* r = p.residual + orow * p.res_cp + ch;
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (ch + h8 * 8 < p.out_c_store) {
              uint4 rv = *(const uint4*)(r + h8 * 8);
              const __half2* rh = (const __half2*)&rv;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 f = __half22float2(rh[j]);
                y[h8 * 8 + 2 * j] += f.x;
                y[h8 * 8 + 2 * j + 1] += f.y;
              }
            }
          }
        }
        if (p.act == PCB_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
        } else if (p.act == PCB_ACT_PRELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = y[j] >= 0.f ? y[j] : y[j] * __ldg(p.slope + ch + j);
        }
        if (p.out_s32) {
          float* o = p.out_s32 + orow * p.out_cp + ch;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (ch + q4 * 4 < p.out_c_store) *(float4*)(o + q4 * 4) = make_float4(y[q4 * 4], y[q4 * 4 + 1], y[q4 * 4 + 2], y[q4 * 4 + 3]);
        } else {
          __half* o = p.out + orow * p.out_cp + ch;
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (ch + h8 * 8 < p.out_c_store) {
              uint4 pk;
              __half2* ph = (__half2*)&pk;
#pragma unroll
              for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(y[h8 * 8 + 2 * j], y[h8 * 8 + 2 * j + 1]);
              *(uint4*)(o + h8 * 8) = pk;
            }
          }
        }
        if (p.out2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaf(y[j], __ldg(p.scale2 + ch + j), __ldg(p.bias2 + ch + j));
          __half* o2 = p.out2 + orow * p.out2_cp + ch;
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (ch + h8 * 8 < p.out_c_store) {
              uint4 pk;
              __half2* ph = (__half2*)&pk;
#pragma unroll
              for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(y[h8 * 8 + 2 * j], y[h8 * 8 + 2 * j + 1]);
              *(uint4*)(o2 + h8 * 8) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp16 map over a row-major [rows][cols] matrix with a {64, box_rows} box, 128B swizzle.
bool make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKC, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

int pcb_conv_tc(pcb_ctx* c, const ConvArgs& a) {
  const PTensor& in = *a.in;
  const ConvWeights& w = *a.w;
  ConvTcParams p{};
  p.rows = (int)in.rows();
  p.dense = in.dense ? 1 : 0;
  p.hp = in.dense ? 0 : in.h + in.pad;
  p.wp = in.dense ? 0 : in.w + in.pad;
  p.pad_lo = in.pad_lo;
  p.pad_hi = in.pad - in.pad_lo;
  p.pad_lo_out = in.pad_lo;
  p.taps = w.taps;
  p.cin_w = w.cin_w;
  p.kchunks = w.cin_w / kKC;
  p.n_tile = w.n_tile;
  p.n_tiles = w.npad / w.n_tile;
  p.m_tiles = (p.rows + kBlockM - 1) / kBlockM;
  p.stride = in.dense ? 1 : a.stride;
  p.act = a.act;
  p.scale = w.scale;
  p.bias = w.bias;
  p.slope = w.slope;
  p.err = c->d_err;
  if (a.out_f32) {
    p.out_f32 = a.out_f32;
    p.out_f32_stride = a.out_f32_stride;
    p.out_f32_cols = w.cout;
  } else {
    const PTensor& out = *a.out;
    if (out.f32) p.out_s32 = (float*)out.data;
    else p.out = out.data;
    p.out_cp = out.cp;
    p.out_c_store = out.cp;
    p.hp_out = out.h + out.pad;
    p.wp_out = out.w + out.pad;
    p.pad_lo_out = out.pad_lo;
    if (a.stride == 1 && (out.pad != in.pad || out.pad_lo != in.pad_lo)) return pcb_fail(c, PCB_ERR_ARG, "conv_tc: stride-1 layers keep the layout");
    if (out.cp > w.npad) return pcb_fail(c, PCB_ERR_ARG, "conv_tc: output channels exceed packed weight rows");
    if (a.residual) {
      if (a.residual->f32) p.residual_f32 = (const float*)a.residual->data;
      else p.residual = a.residual->data;
      p.res_cp = a.residual->cp;
    }
    if (a.out2) {
      if (a.out2->cp != out.cp || a.out2->f32) return pcb_fail(c, PCB_ERR_ARG, "conv_tc: out2 geometry");
      p.out2 = a.out2->data;
      p.out2_cp = a.out2->cp;
      p.scale2 = a.scale2;
      p.bias2 = a.bias2;
    }
  }
  if (w.n_tile % 16 || w.n_tile > 256 || w.n_tile < 16) return pcb_fail(c, PCB_ERR_ARG, "conv_tc: bad n_tile");
  const int b_bytes = w.n_tile * kKC * 2;
  const int stage_bytes = kABytes + b_bytes;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return pcb_fail(c, PCB_ERR_ARG, "conv_tc: tile too large");
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align*/ + (2 * kMaxStages + 4) * 8 + 16;

  CUtensorMap tmA, tmB;
  if (!make_map_2d(&tmA, in.data, (uint64_t)p.rows, (uint64_t)in.cp, (uint64_t)in.cp, kBlockM))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed");
  if (!make_map_2d(&tmB, w.w, (uint64_t)w.npad, (uint64_t)w.taps * w.cin_w, (uint64_t)w.taps * w.cin_w, (uint32_t)w.n_tile))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed");

  static unsigned long long attr_devs = 0;     // the opt-in is per device
  if (pcb_attr_needed(&attr_devs, c->device))
    PCB_CUDA(c, cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < c->num_sms ? total : c->num_sms;
  // algorithmic FLOPs of this layer: 2 * output pixels * cout * cin * taps (real, unpadded extents)
  double out_px = in.dense ? (double)in.n : (double)in.n * (in.h / p.stride) * (in.w / p.stride);
  const double k_real = (w.taps == 1 && w.cin == 3) ? 27.0 : (double)w.cin * w.taps;
  char desc[160];
  desc[0] = 0;
  if (c->profile)
    snprintf(desc, sizeof desc, "n=%d,h=%d,w=%d,cin=%d,cout=%d,taps=%d,stride=%d,ntile=%d,tiles=%d,grid=%d,res=%d,f32out=%d,out2=%d", in.n, in.h,
             in.w, w.cin, w.cout, w.taps, p.stride, p.n_tile, total, grid, a.residual ? (a.residual->f32 ? 2 : 1) : 0,
             (a.out && a.out->f32) ? 1 : 0, a.out2 ? 1 : 0);
  PcbConvTimer timer(c, 2.0 * out_px * (double)w.cout * k_real, desc);
  conv_tc_kernel<<<grid, kThreads, smem, c->stream>>>(tmA, tmB, p);
  PCB_LAUNCH_CHECK(c, "conv_tc_kernel");
  return PCB_OK;
}
