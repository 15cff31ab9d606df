// K2 experiment (conv impl 3; the product kernel is conv_tc2.cu): implicit GEMM on the 5th-gen tensor cores with the
// OUTPUT CHANNELS on the UMMA M dimension and 256 PIXELS on the N dimension:
//     D[128 couts, 256 pixels] += W_tap,kc[128, 64] * X_tap,kc[256, 64]^T      (fp16 in, fp32 in TMEM)
// Replaces the ONNX Runtime / TensorRT execution of the SCRFD and ArcFace graphs (reference:
// person_capture/face_embedder.py:1102-1107, 1341, 1369).
//
// Why this orientation (measured on B200 with the per-role stall counters of conv_tc2.cu, PCB_CONV_DEBUG):
// a tcgen05.mma with both operands in shared memory costs ~140-165 cycles whatever N is (the 128-row A
// operand is read at ~32 B/cycle), so the pixels-on-M kernels (conv_tc.cu, conv_tc2.cu) run the 64- and
// 128-channel layers at 1/4 and 1/2 of the rate of the 256-channel ones.  With couts on M every
// instruction is a full 128x256x16, and the accumulator is 256 TMEM columns, so it is always
// double-buffered (MMA of tile i+1 overlaps the epilogue of tile i).
//
// Operand staging (as conv_tc2.cu): activations live in the P-layout ([rows, Cp] fp16, zero ring), one
// halo tile of 256 + 2*(W+3) rows per 64-channel chunk is loaded ONCE by TMA and the nine taps are the
// same shared-memory tile at row offsets (UMMA descriptor start address; the 128B swizzle is a function of
// absolute smem address bits, so no base-offset is needed).  Weight tiles [128 couts, 64] stream through
// their own ring (or stay resident when all taps*kchunks tiles fit and there is one cout tile).
// MMAs on all-zero K slices (cin not a multiple of 64) are skipped in units of 16 channels.
//
// Warp roles (352 threads, 1 CTA/SM, persistent over tiles):
//   warp 0  X producer (TMA)    warp 1  MMA issuer (+TMEM alloc)    warp 2  W producer (TMA)
//   warps 3-10  epilogue: TMEM lane quarter = warp % 4 (32 couts), pixel half = (warp - 3) / 4.
// Epilogue: a thread owns ONE output channel (scale/bias/slope are registers) and reads 32 pixels per
// tcgen05.ld; pixel -> output-row mapping (ring / stride-2 / tail masking) comes from a 256-entry table
// the epilogue warps build per tile; loads/stores are 2 bytes per lane = 64 contiguous bytes per warp.
//
// Every mbarrier wait is bounded (watchdog): on timeout the kernel raises the context's error word
// and drains instead of hanging the GPU.
#include <stdio.h>
#include <stdlib.h>

#include "pcb_common.cuh"

namespace {

constexpr int kPix = 256;                      // pixels (UMMA N) per tile
constexpr int kMTile = 128;                    // couts (UMMA M) per tile
constexpr int kKC = 64;                        // K elements per stage (one 128B swizzle atom)
constexpr int kWBytes = kMTile * kKC * 2;      // 16 KB weight tile
constexpr int kThreads = 352;
constexpr int kEpiWarp0 = 3;
constexpr int kEpiWarps = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxX = 3;
constexpr int kMaxW = 18;
constexpr int kMaxXLoads = 12;
constexpr int kSmemBudget = 224 * 1024;

struct XLoad {
  int row_rel;    // first global row of the box relative to the tile's first pixel row
  int smem_off;   // byte offset inside the X stage
  int map2;       // 0: 128-row box map, 1: small box map
};

struct Conv3Params {
  int rows;        // input rows N*(H+2)*(W+2)
  int hp, wp;      // input padded dims
  int taps;        // 9 or 1
  int kchunks;     // ceil(cin_eff / 64)
  int kinstr_last; // 16-channel MMA steps in the last chunk (1..4)
  int cin_w;       // packed weight K extent per tap
  int m_tiles, p_tiles;
  int x_stages, x_stage_bytes, x_tx_bytes;
  int w_stages, w_resident;
  int w_slot_bytes;   // 16 KB, or 8 KB when cout <= 64 (only 64 weight rows are loaded; the other accumulator rows are never stored)
  int n_xloads;
  XLoad xloads[kMaxXLoads];
  int tap_off[9];  // byte offset of tap t's first row inside the X stage
  int stride;      // 1 | 2
  int hp_out, wp_out;
  int out_cp;      // channel stride of the P-layout output
  int out_c_store; // channels to store
  int act;
  const float* scale;
  const float* bias;
  const float* slope;
  const __half* residual;
  int res_cp;
  __half* out;
  float* out_s32;              // primary output as fp32 P-layout (SCRFD head maps)
  __half* out2;                // second output: fp16(scale2 * y + bias2)
  int out2_cp;
  const float* scale2;
  const float* bias2;
  int* err;
  unsigned long long* dbg;     // optional [16]: block 0 writes clocks / wait cycles (PCB_CONV_DEBUG)
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
// Bounded wait: false (and *err raised) if the barrier did not flip in ~0.3 s or another role failed.
// `acc` accumulates the cycles spent blocked (debug counters of block 0).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  uint32_t spins = 0;
  bool ok = true;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)err != 0) { ok = false; break; }
      if (clock64() - t0 > 600000000LL) {
        atomicCAS(err, 0, code);
        ok = false;
        break;
      }
    }
  }
  acc += clock64() - t0;
  return ok;
}
// One lane of a converged warp (the canonical single-issuer idiom: control flow stays warp-uniform, so the
// operands of UTCHMMA / UTMALDG / SYNCS live in uniform registers instead of being broadcast lane by lane).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart, base-offset 0
// (the swizzle is applied to absolute shared-memory address bits, see conv_tc2.cu).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t lo = (uint64_t)((saddr >> 4) & 0x3fff);             // start address, LBO = 0
  uint64_t hi = (uint64_t)(1024 >> 4)                           // SBO
                | (1ull << 14)                                  // descriptor version 1 (sm_100)
                | (2ull << 29);                                 // layout type: SWIZZLE_128B
  return lo | (hi << 32);
}

// output row of input P-row `prow` (or -1: ring / odd pixel of a stride-2 conv / beyond the tensor)
__device__ __forceinline__ int out_row(const Conv3Params& p, long long prow) {
  if (prow >= p.rows) return -1;
  const int plane = p.hp * p.wp;
  const int img = (int)(prow / plane);
  const int rem = (int)(prow - (long long)img * plane);
  const int y = rem / p.wp, x = rem - y * p.wp;
  if (y < 1 || y > p.hp - 2 || x < 1 || x > p.wp - 2) return -1;
  if (p.stride == 2) {
    if ((((y - 1) | (x - 1)) & 1) != 0) return -1;
    return (img * p.hp_out + ((y - 1) >> 1) + 1) * p.wp_out + ((x - 1) >> 1) + 1;
  }
  return (int)prow;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ Conv3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_x = smem;
  uint8_t* smem_w = smem + (size_t)p.x_stages * p.x_stage_bytes;
  int* row_tab = (int*)(smem_w + (size_t)p.w_stages * p.w_slot_bytes + kWBytes);   // [8 warps][128] pixel -> output row
  uint64_t* bars = (uint64_t*)(row_tab + kEpiWarps * 128);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + kMaxX;
  uint64_t* w_full = x_empty + kMaxX;
  uint64_t* w_empty = w_full + kMaxW;
  uint64_t* tfull_bar = w_empty + kMaxW;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.p_tiles;
  const int ksteps = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.x_stages; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < p.w_stages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
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
    // ===================== X producer (activation halo tiles) =====================
    {
      if (elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX2) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      long long w_wait = 0;
      const long long c0 = clock64();
      const unsigned long long n0s = globaltimer_ns();
      for (int tile = blockIdx.x; ok && tile < total_tiles; tile += gridDim.x) {
        const int pt = tile / p.m_tiles;
        const int p0 = pt * kPix;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          ok = __all_sync(0xffffffffu, mbar_wait(&x_empty[stage], phase ^ 1, p.err, 101, w_wait));
          if (!ok) break;
          if (elect_one()) {
            const uint32_t sx = smem_u32(smem_x + (size_t)stage * p.x_stage_bytes);
            mbar_expect_tx(&x_full[stage], (uint32_t)p.x_tx_bytes);
            for (int l = 0; l < p.n_xloads; ++l) {
              const XLoad ld = p.xloads[l];
              tma_load_2d(ld.map2 ? &tmX2 : &tmX, &x_full[stage], sx + ld.smem_off, kc * kKC, p0 + ld.row_rel);
            }
          }
          __syncwarp();
          if (++stage == p.x_stages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) {
        p.dbg[0] = (unsigned long long)c0;
        p.dbg[1] = n0s;
        p.dbg[4] = (unsigned long long)w_wait;
      }
    }
  } else if (warp == 2) {
    // ===================== W producer (weight tiles) =====================
    {
      if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      long long w_wait = 0;
      if (p.w_resident) {
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kc = ks / p.taps, t = ks - kc * p.taps;
            mbar_expect_tx(&w_full[ks], (uint32_t)p.w_slot_bytes);
            tma_load_2d(&tmW, &w_full[ks], smem_u32(smem_w + (size_t)ks * p.w_slot_bytes), t * p.cin_w + kc * kKC, 0);
          }
        }
        __syncwarp();
      } else {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < total_tiles; tile += gridDim.x) {
          const int pt = tile / p.m_tiles, mi = tile - pt * p.m_tiles;
          const int m0 = mi * kMTile;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kc = ks / p.taps, t = ks - kc * p.taps;
            ok = __all_sync(0xffffffffu, mbar_wait(&w_empty[stage], phase ^ 1, p.err, 105, w_wait));
            if (!ok) break;
            if (elect_one()) {
              mbar_expect_tx(&w_full[stage], (uint32_t)p.w_slot_bytes);
              tma_load_2d(&tmW, &w_full[stage], smem_u32(smem_w + (size_t)stage * p.w_slot_bytes), t * p.cin_w + kc * kKC, m0);
            }
            __syncwarp();
            if (++stage == p.w_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[5] = (unsigned long long)w_wait;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    {
      const uint32_t idesc = (1u << 4)                          // D format: F32
                             | (0u << 7) | (0u << 10)           // A, B format: F16
                             | ((uint32_t)(kPix >> 3) << 17)    // N = 256 pixels
                             | ((uint32_t)(kMTile >> 4) << 24); // M = 128 couts
      int x_stage = 0, w_stage = 0;
      uint32_t x_phase = 0, w_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ok = true;
      bool w_loaded = false;
      long long wt_x = 0, wt_w = 0, wt_t = 0;
      for (int tile = blockIdx.x; ok && tile < total_tiles; tile += gridDim.x) {
        ok = __all_sync(0xffffffffu, mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err, 102, wt_t));
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kPix);
        for (int kc = 0; ok && kc < p.kchunks; ++kc) {
          ok = __all_sync(0xffffffffu, mbar_wait(&x_full[x_stage], x_phase, p.err, 103, wt_x));
          if (!ok) break;
          const uint32_t sx = smem_u32(smem_x + (size_t)x_stage * p.x_stage_bytes);
          const int kinstr = (kc == p.kchunks - 1) ? p.kinstr_last : kKC / 16;
          for (int t = 0; t < p.taps; ++t) {
            const int slot = p.w_resident ? kc * p.taps + t : w_stage;
            if (!p.w_resident || !w_loaded) {
              ok = __all_sync(0xffffffffu, mbar_wait(&w_full[slot], p.w_resident ? 0u : w_phase, p.err, 106, wt_w));
              if (!ok) break;
            }
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = make_desc_sw128(smem_u32(smem_w + (size_t)slot * p.w_slot_bytes));
              const uint64_t db = make_desc_sw128(sx + (uint32_t)p.tap_off[t]);
              // advance 16 elements (32 bytes) along K inside the swizzle atom
              umma_f16(d_tmem, da, db, idesc, (kc | t) != 0 ? 1u : 0u);
              if (kinstr > 1) umma_f16(d_tmem, da + 2, db + 2, idesc, 1u);
              if (kinstr > 2) umma_f16(d_tmem, da + 4, db + 4, idesc, 1u);
              if (kinstr > 3) umma_f16(d_tmem, da + 6, db + 6, idesc, 1u);
              if (!p.w_resident) umma_commit(&w_empty[w_stage]);
            }
            __syncwarp();
            if (!p.w_resident) {
              if (++w_stage == p.w_stages) { w_stage = 0; w_phase ^= 1; }
            }
          }
          if (!ok) break;
          if (elect_one()) umma_commit(&x_empty[x_stage]);
          __syncwarp();
          if (++x_stage == p.x_stages) { x_stage = 0; x_phase ^= 1; }
        }
        if (!ok) break;
        w_loaded = true;
        if (elect_one()) umma_commit(&tfull_bar[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) {
        p.dbg[6] = (unsigned long long)wt_x;
        p.dbg[7] = (unsigned long long)wt_w;
        p.dbg[8] = (unsigned long long)wt_t;
      }
    }
  } else {
    // ===================== epilogue (warps 3..10) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;     // pixel half: columns [half*128, half*128+128)
    int* tab = row_tab + (warp - kEpiWarp0) * 128;   // this warp's pixel -> output-row table (128 pixels)
    int acc = 0;
    uint32_t acc_phase = 0;
    bool ok = true;
    long long w_full_wait = 0, t_epi = 0, t_ld = 0;
    int cur_mi = -1;
    float sc = 0.f, bi = 0.f, sl = 0.f, sc2 = 0.f, bi2 = 0.f;
    int ch = 0;
    bool ch_ok = false;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int pt = tile / p.m_tiles, mi = tile - pt * p.m_tiles;
      if (mi != cur_mi) {
        cur_mi = mi;
        ch = mi * kMTile + q * 32 + lane;
        ch_ok = ch < p.out_c_store;
        sc = p.scale[ch];
        bi = p.bias[ch];
        sl = p.slope ? p.slope[ch] : 0.f;
        sc2 = p.scale2 ? p.scale2[ch] : 0.f;
        bi2 = p.bias2 ? p.bias2[ch] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) tab[e * 32 + lane] = out_row(p, (long long)pt * kPix + half * 128 + e * 32 + lane);
      __syncwarp();
      if (ok) ok = mbar_wait(&tfull_bar[acc], acc_phase, p.err, 104, w_full_wait);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const long long te0 = clock64();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kPix + half * 128);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        const long long tl0 = clock64();
        tmem_ld32(t_row + c * 32, v);
        t_ld += clock64() - tl0;
        int orow[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int4 o4 = *(const int4*)(tab + c * 32 + g * 4);
          orow[g * 4 + 0] = o4.x;
          orow[g * 4 + 1] = o4.y;
          orow[g * 4 + 2] = o4.z;
          orow[g * 4 + 3] = o4.w;
        }
        if (ch_ok) {
          float r[32];
          if (p.residual) {
            const __half* rc = p.residual + ch;
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = orow[j] >= 0 ? __half2float(rc[(long long)orow[j] * p.res_cp]) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (orow[j] < 0) continue;     // uniform across the warp: every lane looks at the same pixel
            float y = fmaf(__uint_as_float(v[j]), sc, bi);
            if (p.residual) y += r[j];
            if (p.act == PCB_ACT_RELU) y = fmaxf(y, 0.f);
            else if (p.act == PCB_ACT_PRELU) y = y >= 0.f ? y : y * sl;
            if (p.out_s32) p.out_s32[(long long)orow[j] * p.out_cp + ch] = y;
            else p.out[(long long)orow[j] * p.out_cp + ch] = __float2half_rn(y);
            if (p.out2) p.out2[(long long)orow[j] * p.out2_cp + ch] = __float2half_rn(fmaf(y, sc2, bi2));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      t_epi += clock64() - te0;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.dbg && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0) {
      p.dbg[9] = (unsigned long long)w_full_wait;
      p.dbg[10] = (unsigned long long)t_epi;
      p.dbg[11] = (unsigned long long)t_ld;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    p.dbg[2] = (unsigned long long)clock64();
    p.dbg[3] = globaltimer_ns();
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

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// X-stage plan for a 256-pixel tile: one contiguous halo range when 2*(W+3) extra rows fit two small
// boxes, else three row bands (dy = -1, 0, +1).  Returns false when neither fits the box limits.
bool plan_x(Conv3Params& p, int* x2_rows) {
  p.n_xloads = 0;
  *x2_rows = 8;
  auto add = [&](int row_rel, int off, int map2) {
    if (p.n_xloads >= kMaxXLoads) return false;
    p.xloads[p.n_xloads++] = XLoad{row_rel, off, map2};
    return true;
  };
  const int main_boxes = kPix / 128;
  if (p.taps == 1) {
    for (int j = 0; j < main_boxes; ++j) add(j * 128, j * 128 * 128, 0);
    p.tap_off[0] = 0;
    p.x_stage_bytes = kPix * 128;
    p.x_tx_bytes = kPix * 128;
    return true;
  }
  const int halo = p.wp + 1;
  const int extra = pcb_round_up(2 * halo, 16);          // split into one or two small boxes
  const int n_small = extra <= 256 ? 1 : 2;
  const long long merged_rows = (long long)kPix + extra;
  const long long banded_rows = 3LL * (kPix + 8);
  if (extra <= 512 && merged_rows <= banded_rows) {
    for (int j = 0; j < main_boxes; ++j) add(-halo + j * 128, j * 128 * 128, 0);
    const int small = extra / n_small;
    for (int s = 0; s < n_small; ++s) add(-halo + kPix + s * small, (kPix + s * small) * 128, 1);
    *x2_rows = small;
    for (int t = 0; t < 9; ++t) p.tap_off[t] = (halo + (t / 3 - 1) * p.wp + (t % 3 - 1)) * 128;
    p.x_tx_bytes = (int)merged_rows * 128;
    p.x_stage_bytes = pcb_round_up(p.x_tx_bytes, 1024);
    return true;
  }
  const int band_bytes = pcb_round_up((kPix + 8) * 128, 1024);
  for (int b = 0; b < 3; ++b) {
    for (int j = 0; j < main_boxes; ++j)
      if (!add((b - 1) * p.wp - 1 + j * 128, b * band_bytes + j * 128 * 128, 0)) return false;
    if (!add((b - 1) * p.wp - 1 + kPix, b * band_bytes + kPix * 128, 1)) return false;
  }
  for (int t = 0; t < 9; ++t) p.tap_off[t] = (t / 3) * band_bytes + (t % 3) * 128;
  p.x_tx_bytes = (int)banded_rows * 128;
  p.x_stage_bytes = 3 * band_bytes;
  return true;
}

}  // namespace

int pcb_conv_tc3(pcb_ctx* c, const ConvArgs& a) {
  const PTensor& in = *a.in;
  const ConvWeights& w = *a.w;
  // dense (FC), fp32 residual streams and fp32 dense outputs stay on the pixels-on-M kernel
  if (in.dense || a.out_f32 || (a.residual && a.residual->f32) || env_int("PCB_CONV_NO_TC3", 0)) return pcb_conv_tc2(c, a);
  Conv3Params p{};
  p.rows = (int)in.rows();
  p.hp = in.h + 2;
  p.wp = in.w + 2;
  p.taps = w.taps;
  p.cin_w = w.cin_w;
  const int cin_eff = (w.taps == 1 && w.cin == 3) ? 27 : w.cin;   // stem: 27 patch channels
  p.kchunks = (cin_eff + kKC - 1) / kKC;
  p.kinstr_last = (cin_eff - (p.kchunks - 1) * kKC + 15) / 16;
  p.m_tiles = w.rows_alloc / kMTile;
  p.p_tiles = (p.rows + kPix - 1) / kPix;
  p.stride = a.stride;
  p.act = a.act;
  p.scale = w.scale;
  p.bias = w.bias;
  p.slope = w.slope;
  p.err = c->d_err;
  const PTensor& out = *a.out;
  if (out.f32) p.out_s32 = (float*)out.data;
  else p.out = out.data;
  p.out_cp = out.cp;
  p.out_c_store = out.cp;
  p.hp_out = out.h + 2;
  p.wp_out = out.w + 2;
  if (out.cp > w.rows_alloc) return pcb_fail(c, PCB_ERR_ARG, "conv_tc3: output channels exceed packed weight rows");
  if ((long long)out.rows() >= 0x7fffffffLL) return pcb_conv_tc2(c, a);
  if (a.residual) {
    p.residual = a.residual->data;
    p.res_cp = a.residual->cp;
  }
  if (a.out2) {
    if (a.out2->cp != out.cp || a.out2->f32) return pcb_fail(c, PCB_ERR_ARG, "conv_tc3: out2 geometry");
    p.out2 = a.out2->data;
    p.out2_cp = a.out2->cp;
    p.scale2 = a.scale2;
    p.bias2 = a.bias2;
  }
  int x2_rows = 8;
  if (!plan_x(p, &x2_rows)) return pcb_conv_tc2(c, a);
  if (env_int("PCB_TAP_ALIGN", 0)) for (int t = 0; t < 9; ++t) p.tap_off[t] &= ~1023;   // timing experiment only (wrong sums)
  const int ksteps = p.taps * p.kchunks;
  const int w_box_rows = out.cp <= 64 ? 64 : kMTile;
  p.w_slot_bytes = w_box_rows * kKC * 2;
  const int wsb = p.w_slot_bytes;
  // kWBytes of slack after the last slot: with 8 KB slots the UMMA still reads 128 rows from the slot start
  const int fixed = kWBytes + kEpiWarps * 128 * 4 + (2 * kMaxX + 2 * kMaxW + 4) * 8 + 16 + 1024;
  const int room = kSmemBudget - fixed;
  // with a single chunk per tile one X stage per tile is consumed: two stages double-buffer across tiles
  p.x_stages = 2;
  if (2 * p.x_stage_bytes + 3 * wsb > room) return pcb_conv_tc2(c, a);   // very wide maps
  p.w_resident = 0;
  if (p.m_tiles == 1 && ksteps <= kMaxW && 2 * p.x_stage_bytes + ksteps * wsb <= room && !env_int("PCB_CONV_NO_RESIDENT", 0)) {
    p.w_resident = 1;
    p.w_stages = ksteps;
    p.x_stages = (room - ksteps * wsb) / p.x_stage_bytes;
  } else {
    p.w_stages = (room - 2 * p.x_stage_bytes) / wsb;
    if (p.w_stages > 9 && (room - 3 * p.x_stage_bytes) / wsb >= 6) {
      p.x_stages = 3;
      p.w_stages = (room - 3 * p.x_stage_bytes) / wsb;
    }
    if (p.w_stages > kMaxW) p.w_stages = kMaxW;
  }
  if (p.x_stages > kMaxX) p.x_stages = kMaxX;
  const size_t smem = (size_t)p.x_stages * p.x_stage_bytes + (size_t)p.w_stages * wsb + fixed;

  CUtensorMap tmX, tmX2, tmW;
  if (!make_map_2d(&tmX, in.data, (uint64_t)p.rows, (uint64_t)in.cp, (uint64_t)in.cp, 128))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed");
  if (!make_map_2d(&tmX2, in.data, (uint64_t)p.rows, (uint64_t)in.cp, (uint64_t)in.cp, (uint32_t)x2_rows))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(X2) failed");
  if (!make_map_2d(&tmW, w.w, (uint64_t)w.rows_alloc, (uint64_t)w.taps * w.cin_w, (uint64_t)w.taps * w.cin_w, (uint32_t)w_box_rows))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed");

  static bool attr_set = false;
  if (!attr_set) {
    PCB_CUDA(c, cudaFuncSetAttribute(conv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int total = p.m_tiles * p.p_tiles;
  const int grid = total < c->num_sms ? total : c->num_sms;
  // algorithmic FLOPs of this layer: 2 * output pixels * cout * cin * taps (real, unpadded extents)
  const double out_px = (double)in.n * (in.h / p.stride) * (in.w / p.stride);
  const double k_real = (w.taps == 1 && w.cin == 3) ? 27.0 : (double)w.cin * w.taps;
  char desc[220];
  desc[0] = 0;
  if (c->profile)
    snprintf(desc, sizeof desc, "tc3 n=%d,h=%d,w=%d,cin=%d,cout=%d,taps=%d,stride=%d,mtiles=%d,tiles=%d,grid=%d,res=%d,f32out=%d,out2=%d,xst=%d,wst=%d,wres=%d",
             in.n, in.h, in.w, w.cin, w.cout, w.taps, p.stride, p.m_tiles, total, grid, a.residual ? 1 : 0, out.f32 ? 1 : 0,
             a.out2 ? 1 : 0, p.x_stages, p.w_stages, p.w_resident);
  static const int debug = env_int("PCB_CONV_DEBUG", 0);
  static unsigned long long* dbg_dev = nullptr;
  if (debug && c->profile) {
    if (!dbg_dev) dbg_dev = (unsigned long long*)pcb_dev_alloc(c, 16 * sizeof(unsigned long long), true);
    p.dbg = dbg_dev;
  }
  {
    PcbConvTimer timer(c, 2.0 * out_px * (double)w.cout * k_real, desc);
    conv_tc3_kernel<<<grid, kThreads, smem, c->stream>>>(tmX, tmX2, tmW, p);
  }
  PCB_LAUNCH_CHECK(c, "conv_tc3_kernel");
  if (p.dbg) {
    // debug only: serialises the stream.  SM MHz seen by block 0, its cycles, and the fraction of them each role
    // of block 0 spent blocked (X/W producer on empty slots, MMA on X/W full and TMEM empty, epilogue on TMEM full)
    unsigned long long h[16];
    cudaMemcpyAsync(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    const double cyc = (double)(h[2] - h[0]), ns = (double)(h[3] - h[1]);
    fprintf(stderr, "CONVDBG %s | mhz=%.0f cyc=%.0f prodA_wait=%.2f prodB_wait=%.2f mma_waitA=%.2f mma_waitB=%.2f mma_waitT=%.2f epi_wait=%.2f epi_busy=%.2f epi_ldtm=%.2f\n",
            desc, ns > 0 ? cyc / ns * 1e3 : 0.0, cyc, h[4] / cyc, h[5] / cyc, h[6] / cyc, h[7] / cyc, h[8] / cyc, h[9] / cyc, h[10] / cyc, h[11] / cyc);
  }
  return PCB_OK;
}
