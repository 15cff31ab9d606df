// K2 (product kernel): implicit-GEMM convolution on the 5th-gen tensor cores with operand reuse in
// shared memory.  Replaces the ONNX Runtime / TensorRT execution of the SCRFD and ArcFace graphs
// (reference: person_capture/face_embedder.py:1102-1107, 1341, 1369).
//
// Why a second kernel: the first formulation (conv_tc.cu, kept as the A/B baseline, conv impl 2)
// issues one TMA load of the activation tile per tap, i.e. it reads A nine times and the weight
// matrix once per 128 output rows.  ncu shows it pinned at the L2->SM bandwidth limit (~12 TB/s) with
// the tensor pipe half idle.  This kernel cuts L2 traffic ~2.6x:
//   * A halo tile: for a tile of M = MT*128 consecutive P-layout rows the rows
//     [m0 - (W+3), m0 + M + (W+3)) of one 64-channel chunk are loaded ONCE; tap (dy,dx) is the same
//     shared-memory tile at a row offset of (dy+1)*(W+2) + dx + 1, expressed through the UMMA
//     descriptor start address alone (the swizzle is a function of absolute smem address bits).  For wide
//     maps (2*(W+3) > 256 rows) three row bands (dy = -1,0,1) are loaded instead (3x, not 9x).
//   * MT = 2: two 128-row accumulators share every weight stage, halving weight traffic per FLOP.
//   * small layers (taps*kchunks weight tiles fit in the ring and there is one N tile) keep the whole
//     weight matrix resident in shared memory for the lifetime of the CTA.
// Pipelines: A ring (2-3 stages) and B ring (up to 18 stages) with separate producer warps, TMEM
// accumulators double-buffered when MT*N <= 256 columns.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over tiles):
//   warp 0  A producer (TMA)    warp 1  MMA issuer 0 (+TMEM alloc)    warp 2  B producer (TMA)    warp 3  MMA issuer 1
//   warps 4-11  epilogue: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4.
// Registers: the kernel is built for 168 per thread (384 threads); warps 0-3 (one warpgroup) give theirs back with
// setmaxnreg.dec 72 and the two epilogue warpgroups take 216 each with setmaxnreg.inc.  At 168 the epilogue spilled loop
// invariants to local memory, and with the L1 squeezed to ~28 KB by the operand rings every reload was a long-scoreboard
// stall (ncu: 22 % of the epilogue's samples on the instruction after an LDL).
// Two issuers: measured with tools/mma_rate.cu, ONE thread sustains a tcgen05.mma every ~65-85 cycles (59 + 0.2 N), so
// instructions with N <= 128 (32-64 tensor-pipe cycles) leave the pipe half idle however the loop is written; two warps
// issuing to disjoint accumulators reach the nominal rate (N = 128: 64.0 cycles, 8190 FLOP/clk/SM).  The issuers split a
// K step by accumulator: MT = 2 -> one 128-row sub-tile each; MT = 1 -> one half of the N columns each (two N/2-wide
// instructions on the same A rows).  Empty / TMEM-full barriers then take one commit per issuer.
// Epilogue per 32 columns: tcgen05.ld.x32 -> y = acc*scale+bias (+residual) -> ReLU/PReLU -> fp16
// (or fp32 head maps) -> 16-byte stores of interior pixels; optional second output scale2*y+bias2
// (the next block's BatchNorm).  Per-channel vectors live in shared memory; the residual rows of a
// tile are prefetched into registers before the accumulator is waited for.
//
// Kernel variants (template parameter kVar, so the role loops carry no per-variant branches):
//   0 halo     the formulation above, one CTA per tile, tcgen05.mma.cta_group::1 (M = 128)
//   1 strided  stride-2 convolutions on the output grid, one 4-D strided TMA load per tap
//   2 pair     halo formulation on a 2-CTA cluster (one TPC): the CTAs own adjacent M tiles and HALF of every
//              weight stage each; the leader issues tcgen05.mma.cta_group::2 (M = 256, 128 rows per SM), so an SM
//              reads only half of B from its own shared memory per MMA and TMA writes half the weight bytes into
//              it -- the shared-memory port, not the tensor pipe, paces the cta_group::1 kernel on the wide layers.
//
// Every mbarrier wait is bounded (watchdog): on timeout the kernel raises the context's error word
// and drains instead of hanging the GPU.
#include <stdio.h>
#include <stdlib.h>

#include "pcb_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kKC = 64;                       // K elements per weight stage (one 128B swizzle atom)
constexpr int kThreads = 384;
constexpr int kIss1Warp = 3;                  // second MMA issuer warp
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxA = 6;
constexpr int kMaxB = 18;
constexpr int kMaxALoads = 12;
constexpr int kSmemBudget = 224 * 1024;
constexpr int kPairMinN = 128;               // default policy: pair mode for layers with n_tile >= this and >= 2 K chunks (PCB_CONV_PAIR=2 overrides)

struct ALoad {
  int row_rel;    // first global row of the box relative to m0
  int smem_off;   // byte offset inside the A stage
  int map2;       // 0: 128-row box map, 1: small box map
};

struct Conv2Params {
  int rows;        // input rows (N*(H+2)*(W+2)) or dense rows
  int hp, wp;      // input padded dims; 0 for dense
  int taps;        // 9 or 1
  int kchunks;     // ceil(cin_eff / kc)
  int kinstr_last; // 16-channel MMA steps in the last chunk (1..kc/16)
  int kc;          // K elements per stage: 64 (128-byte swizzled rows) or 32 (64-byte rows, layers with <= 32 input channels)
  int row_bytes;   // shared-memory row pitch of an operand tile: 2 * kc
  int sub_bytes;   // one 128-row A sub-tile: 128 * row_bytes
  int cin_w;       // packed weight K extent per tap
  int n_tile, n_tiles, m_tiles;   // m_tiles counts tiles of mt*128 rows
  int mt;          // 128-row sub-tiles per CTA tile
  int sub_cols;    // TMEM columns between sub-tile accumulators (n_tile rounded up to 32)
  int acc_bufs;    // 1 | 2
  int a_stages, a_stage_bytes, a_tx_bytes;
  int b_stages, b_bytes, b_resident;   // b_bytes: bytes of one weight stage in THIS CTA's shared memory
  int b_rows;      // weight rows per stage in this CTA (n_tile, or n_tile/2 in pair mode)
  int n_iss;       // MMA issuer warps in use (1 | 2)
  int split_n;     // 1: the issuers split the N columns (MT = 1); 0: they split the sub-tiles (MT = 2) or there is one issuer
  int b_nloads;    // TMA loads per weight stage and CTA (2 in pair + split_n mode: this CTA's quarter of each column half)
  int b_load_rows; // weight rows per load (= box rows of the weight map)
  int n_aloads;
  ALoad aloads[kMaxALoads];
  int tap_off[9];  // byte offset of tap t's first row inside the A stage
  // strided mode (stride-2 convolutions): the tile is nb images x by output rows x bx output pixels (<= 128 GEMM rows),
  // each tap is its own TMA load through a 4-D map with element strides {1,2,2,1} (only the even input pixels of the tap
  // are fetched), so the GEMM runs on the OUTPUT grid instead of computing all input positions and dropping 3 of 4
  int strided;
  int tma_store;   // 1: fp16 outputs leave through TMA stores of the staged 32-row x 32-channel chunk (stride-1 spatial layers)
  int epi_row_split;   // 1: the two epilogue warps of a TMEM lane quarter take one 128-row sub-tile each (all columns); 0: one column half each
  int pair;        // 1: 2-CTA cluster, cta_group::2 (m_tiles counts PAIR tiles of 2*mt*128 rows; b_bytes is this CTA's half stage)
  int s_bx, s_by, s_nb;        // output pixels / output rows / images per tile
  int s_tx, s_ty;              // tiles per output row / per image column of rows
  int s_n;                     // images
  int s_ho, s_wo;              // output dims
  uint32_t fd_plane_mul, fd_wp_mul;   // fast_div multipliers for hp*wp and wp
  int fd_plane_shift, fd_wp_shift;
  int stride;      // 1 | 2
  int hp_out, wp_out;
  int pad_lo, pad_hi;   // input tensor: padding before / after the pixels of a row and of an image (per-tensor layout)
  int pad_lo_out;       // output tensor: padding before the pixels
  int out_cp;      // channel stride of the P-layout output
  int out_c_store; // channels to store (<= out_cp, multiple of 8)
  int act;
  int dense;
  int vec_n;       // shared-memory entries of each per-channel vector (npad rounded up to 32)
  int vec_real;    // entries present in global memory (npad)
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
  float* out_f32;              // dense fp32 output (FC)
  int out_f32_stride;
  int out_f32_cols;
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
// (A suspend-time hint on try_wait was tried -- 20 us, to park the eight epilogue warps and the producers instead of letting them
// come back every ~500 cycles -- and measured 5-8 % SLOWER on the layer profile: the wake-up is late.  Plain try_wait stays.)
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

// mbar_wait that also accumulates the cycles spent waiting (debug instrumentation of block 0).  The first try_wait is
// timed too: mbarrier.try_wait may suspend the thread inside the instruction before it reports success.
__device__ __forceinline__ bool mbar_wait_t(uint64_t* bar, uint32_t parity, int* err, int code, long long& acc, bool timed) {
  if (!timed) return mbar_wait(bar, parity, err, code);
  const long long t0 = clock64();
  const bool ok = mbar_wait(bar, parity, err, code);
  acc += clock64() - t0;
  return ok;
}
// Address-based forms for the MMA issuers' loop (shared-memory addresses precomputed as 32-bit values).
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_addr(uint32_t bar, uint32_t parity, int* err, int code, long long& acc, bool timed) {
  if (!timed && mbar_try_wait_addr(bar, parity)) return true;
  const long long t0 = clock64();      // debug runs time the first try too: try_wait may suspend inside the instruction
  uint32_t spins = 0;
  bool ok = true;
  while (!mbar_try_wait_addr(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*(volatile int*)err != 0) { ok = false; break; }
      if (clock64() - t0 > 600000000LL) {
        atomicCAS(err, 0, code);
        ok = false;
        break;
      }
    }
  }
  if (timed) acc += clock64() - t0;
  return ok;
}
// One lane of a converged warp (the canonical single-issuer idiom: control flow stays warp-uniform, so the
// operands of UTCHMMA / UTMALDG / SYNCS live in uniform registers instead of being broadcast lane by lane --
// with `if (lane == 0)` around the loop every tcgen05.mma was wrapped in an ELECT/R2UR/BRA.U.ANY loop and
// issue, not the tensor pipe, set the pace at ~150 cycles per instruction).
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

// Address of `bar` in the shared memory of the cluster's rank-0 CTA (the pair's MMA leader).
__device__ __forceinline__ uint32_t leader_addr(const void* bar) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_u32(bar)));
  return r;
}
// kPair: the destination is this CTA's shared memory, the mbarrier (a shared::cluster address) is the leader's.
template <bool kPair>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  if (kPair)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// TMA store of a staged chunk (shared -> global, bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA).  Relaxed: the hand-over is of TMEM columns,
// ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; no generic-memory data rides on it, and the release
// form cost a MEMBAR + ERRBAR per tile (5 % of the epilogue's stall samples in ncu).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <bool kPair>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  if (kPair)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
}
// kPair: the arrival is multicast to the barrier at this offset in BOTH CTAs of the pair (each frees its own stage /
// wakes its own epilogue) once the pair's MMAs issued so far have completed.
template <bool kPair>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (kPair)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  if (kPair)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptors (built inline in the issuer loop): K-major, 128B-swizzled operand tile, rows of 128 bytes,
// 8-row groups 1024 bytes apart (64-byte rows / SWIZZLE_64B / 512 bytes for layers with <= 32 input channels).  The start
// address may sit on any row (tap shifts).  Measured on B200 (tools/conv_check.py): the tensor core applies the swizzle to
// ABSOLUTE shared-memory address bits (the same rule TMA writes with), so a row-shifted start needs base-offset 0; setting
// base-offset = (addr >> 7) & 7 double-counts the phase and gives wrong sums (that experiment: PCB_DESC_MODE=0 up to commit 9c41a29).

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// n / d for 0 <= n < 2^31 with the multiplier the host precomputed (fd_mul = floor(2^32 * (2^s - d) / d) + 1, s = ceil(log2 d))
__device__ __forceinline__ int fast_div(int n, uint32_t mul, int shift) {
  return (int)(((uint64_t)__umulhi((uint32_t)n, mul) + (uint32_t)n) >> shift);
}

struct RowInfo {
  bool valid;
  long long orow;
};

template <bool kStrided>
__device__ __forceinline__ RowInfo row_info(const Conv2Params& p, long long prow) {
  RowInfo r;
  if (kStrided) {
    // prow = tile * 128 + row in tile; rows of a tile enumerate (image, output row, output pixel) of the strided box
    const int tile = (int)(prow >> 7), rr = (int)(prow & 127);
    const int tx = tile % p.s_tx;
    const int ty = (tile / p.s_tx) % p.s_ty;
    const int ig = tile / (p.s_tx * p.s_ty);
    const int per_img = p.s_by * p.s_bx;
    const int io = rr / per_img, rem = rr - io * per_img;
    const int j = rem / p.s_bx, i = rem - j * p.s_bx;
    const int img = ig * p.s_nb + io, oy = ty * p.s_by + j, ox = tx * p.s_bx + i;
    r.valid = io < p.s_nb && img < p.s_n && oy < p.s_ho && ox < p.s_wo;
    r.orow = ((long long)img * p.hp_out + oy + p.pad_lo_out) * p.wp_out + ox + p.pad_lo_out;
    return r;
  }
  r.valid = prow < p.rows;
  r.orow = prow;
  if (!p.dense) {
    const int plane = p.hp * p.wp;
    const int img = fast_div((int)prow, p.fd_plane_mul, p.fd_plane_shift);
    const int rem = (int)prow - img * plane;
    const int y = fast_div(rem, p.fd_wp_mul, p.fd_wp_shift), x = rem - y * p.wp;
    r.valid = r.valid && y >= p.pad_lo && y <= p.hp - 1 - p.pad_hi && x >= p.pad_lo && x <= p.wp - 1 - p.pad_hi;
    if (p.stride == 2) {
      r.valid = r.valid && (((y - p.pad_lo) | (x - p.pad_lo)) & 1) == 0;
      r.orow = ((long long)img * p.hp_out + ((y - p.pad_lo) >> 1) + p.pad_lo_out) * p.wp_out + ((x - p.pad_lo) >> 1) + p.pad_lo_out;
    }
  }
  return r;
}

__device__ __forceinline__ void unpack_h8(const uint4& v, float* f) {
  const __half2* h = (const __half2*)&v;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_h8(const float* f) {
  uint4 v;
  __half2* h = (__half2*)&v;
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

constexpr int kVarHalo = 0, kVarStrided = 1, kVarPair = 2;

// kAct: PCB_ACT_*; kRes: residual add; kOut2: second (affine) output; kOutMode: 0 fp16 P-layout, 1 fp32 P-layout, 2 dense fp32;
// kVar: kVarHalo | kVarStrided | kVarPair (launched as 2-CTA clusters)
template <int kAct, bool kRes, bool kOut2, int kOutMode, int kVar>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ Conv2Params p) {
  constexpr bool kPair = kVar == kVarPair;
  constexpr bool kStrided = kVar == kVarStrided;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)p.a_stages * p.a_stage_bytes;
  float* vec = (float*)(smem_b + (size_t)p.b_stages * p.b_bytes);   // scale | bias | slope | scale2 | bias2
  // [8 warps][2 KB] epilogue staging buffers, 1 KB aligned: the TMA store's 64B swizzle is a function of address bits 7-8
  uint8_t* stage_buf = (uint8_t*)(((uintptr_t)(vec + 5 * p.vec_n) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(stage_buf + kEpiWarps * 2048);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMaxA;
  uint64_t* b_full = a_empty + kMaxA;
  uint64_t* b_empty = b_full + kMaxB;
  uint64_t* tfull_bar = b_empty + kMaxB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  // pair mode: clusters are consecutive CTA pairs along x; rank 0 leads.  Both CTAs walk the same PAIR tiles; inside a
  // pair tile CTA r owns M tile 2*mt_idx + r.  The full barriers that count are the leader's (both CTAs' TMA loads
  // complete on them, the leader alone expects the bytes); the empty / TMEM-full barriers are per CTA (multicast commit);
  // the TMEM-empty barrier that counts is the leader's (both CTAs' epilogue warps arrive on it).
  const int cta_rank = kPair ? (int)(blockIdx.x & 1) : 0;
  const bool leader = cta_rank == 0;
  const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int tile_rows = p.mt * kBlockM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], p.n_iss);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], p.n_iss);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], p.n_iss);
      mbar_init(&tempty_bar[s], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (kPair) {
      // the same warp of both CTAs allocates; the columns are the same in both halves of the pair
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // per-channel vectors -> shared memory (epilogue reads them as broadcast float4)
  for (int i = threadIdx.x; i < p.vec_n; i += kThreads) {
    const bool in = i < p.vec_real;
    vec[i] = in ? p.scale[i] : 0.f;
    vec[p.vec_n + i] = in ? p.bias[i] : 0.f;
    vec[2 * p.vec_n + i] = (in && p.slope) ? p.slope[i] : 0.f;
    vec[3 * p.vec_n + i] = (in && p.scale2) ? p.scale2[i] : 0.f;
    vec[4 * p.vec_n + i] = (in && p.bias2) ? p.bias2[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // the leader's barriers exist before the peer's loads / arrivals target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch: everything above touches only this kernel's own state and constants (barriers, TMEM,
  // per-channel vectors), so it may run while the previous kernel of the stream drains its last tiles.  From here on the
  // warps read activations / residuals and write outputs: they wait for the previous grid to complete (and its writes to
  // be visible).  The weight producer (warp 2) does not wait -- weights are constants -- so the B ring is already full when
  // the first activation tile lands.  The trigger for OUR dependent goes out right away: this grid is fully resident
  // (persistent, <= 1 CTA per SM), so the next kernel's CTAs can only take SMs our CTAs have left.  Both instructions are
  // no-ops when the launch carries no programmatic attribute (PCB_CONV_PDL=0).
  if (warp != 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // (setmaxnreg sits at the top of each role branch: ptxas bounds the registers of the code a setmaxnreg dominates)
  if (warp < kEpiWarp0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");     // one instruction for the whole warpgroup (warps 0-3)
  if (warp == 0) {
    // ===================== A producer (whole warp loops, one elected lane issues) =====================
    {
      if (elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      long long w_empty = 0;
      const long long c0 = clock64();
      const unsigned long long n0s = globaltimer_ns();
      for (int tile = tile0; ok && tile < total_tiles; tile += tile_step) {
        const int mt_idx = tile / p.n_tiles;
        if (kStrided) {
          // tile -> (image group, output row block, output column block); one 4-D strided load per (chunk, tap)
          const int tx = mt_idx % p.s_tx;
          const int ty = (mt_idx / p.s_tx) % p.s_ty;
          const int ig = mt_idx / (p.s_tx * p.s_ty);
          for (int kc = 0; ok && kc < p.kchunks; ++kc) {
            for (int t = 0; t < p.taps; ++t) {
              ok = __all_sync(0xffffffffu, mbar_wait_t(&a_empty[stage], phase ^ 1, p.err, 101, w_empty, p.dbg != nullptr));
              if (!ok) break;
              if (elect_one()) {
                const int ky = p.taps == 9 ? t / 3 : 1, kx = p.taps == 9 ? t % 3 : 1;
                mbar_expect_tx(&a_full[stage], (uint32_t)p.a_tx_bytes);
                tma_load_4d(&tmA, &a_full[stage], smem_u32(smem_a + (size_t)stage * p.a_stage_bytes), kc * p.kc,
                            2 * tx * p.s_bx + kx - 1 + p.pad_lo, 2 * ty * p.s_by + ky - 1 + p.pad_lo, ig * p.s_nb);   // -1: left / top pad = out-of-bounds zero fill
              }
              __syncwarp();
              if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          const int m0 = (kPair ? 2 * mt_idx + cta_rank : mt_idx) * tile_rows;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            ok = __all_sync(0xffffffffu, mbar_wait_t(&a_empty[stage], phase ^ 1, p.err, 101, w_empty, p.dbg != nullptr));
            if (!ok) break;
            if (elect_one()) {
              const uint32_t sa = smem_u32(smem_a + (size_t)stage * p.a_stage_bytes);
              const uint32_t bar = kPair ? leader_addr(&a_full[stage]) : smem_u32(&a_full[stage]);
              if (leader) mbar_expect_tx(&a_full[stage], (uint32_t)(kPair ? 2 * p.a_tx_bytes : p.a_tx_bytes));
              for (int l = 0; l < p.n_aloads; ++l) {
                const ALoad ld = p.aloads[l];
                tma_load_2d<kPair>(ld.map2 ? &tmA2 : &tmA, bar, sa + ld.smem_off, kc * p.kc, m0 + ld.row_rel);
              }
            }
            __syncwarp();
            if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) {
        p.dbg[0] = (unsigned long long)c0;
        p.dbg[1] = n0s;
        p.dbg[4] = (unsigned long long)w_empty;
      }
    }
  } else if (warp == 2) {
    // ===================== B producer =====================
    {
      if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      const int ksteps = p.taps * p.kchunks;
      const uint32_t b_tx = (uint32_t)(kPair ? 2 * p.b_bytes : p.b_bytes);
      // weight rows of load h of a stage: h * n_tile/2 + rank * b_load_rows (one load: this CTA's rows of the stage; two loads,
      // pair + split_n: this CTA's quarter of each issuer's column half), packed back to back in the stage
      const int n_off = cta_rank * p.b_load_rows;
      const uint32_t load_bytes = (uint32_t)(p.b_load_rows * p.row_bytes);
      long long w_empty = 0;
      if (p.b_resident) {
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kc = ks / p.taps, t = ks - kc * p.taps;
            if (leader) mbar_expect_tx(&b_full[ks], b_tx);
            for (int h = 0; h < p.b_nloads; ++h)
              tma_load_2d<kPair>(&tmB, kPair ? leader_addr(&b_full[ks]) : smem_u32(&b_full[ks]),
                                 smem_u32(smem_b + (size_t)ks * p.b_bytes) + h * load_bytes, t * p.cin_w + kc * p.kc, n_off + h * (p.n_tile / 2));
          }
        }
        __syncwarp();
      } else {
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = tile0; ok && tile < total_tiles; tile += tile_step) {
          const int mt_idx = tile / p.n_tiles, nt = tile - mt_idx * p.n_tiles;
          const int n0 = nt * p.n_tile + n_off;
          for (int ks = 0; ks < ksteps; ++ks) {
            const int kc = ks / p.taps, t = ks - kc * p.taps;
            ok = __all_sync(0xffffffffu, mbar_wait_t(&b_empty[stage], phase ^ 1, p.err, 105, w_empty, p.dbg != nullptr));
            if (!ok) break;
            if (elect_one()) {
              if (leader) mbar_expect_tx(&b_full[stage], b_tx);
              for (int h = 0; h < p.b_nloads; ++h)
                tma_load_2d<kPair>(&tmB, kPair ? leader_addr(&b_full[stage]) : smem_u32(&b_full[stage]),
                                   smem_u32(smem_b + (size_t)stage * p.b_bytes) + h * load_bytes, t * p.cin_w + kc * p.kc, n0 + h * (p.n_tile / 2));
            }
            __syncwarp();
            if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[5] = (unsigned long long)w_empty;
    }
  } else if (warp == 1 || warp == kIss1Warp) {
    // ===================== MMA issuers (whole warp loops, one elected lane issues; pair: the leader CTA only) =====================
    const int iss = warp == 1 ? 0 : 1;
    if (leader && iss < p.n_iss) {
      // The loop below is what paces the tensor pipe (ncu: the issuer warps spent ~220 cycles per instruction where the
      // pipe needs 64-128), so everything loop-invariant lives in registers: shared-memory addresses as 32-bit values
      // computed once (re-deriving them from generic pointers cost an S2UR/ULEA/ULOP3 chain per use), barrier addresses as
      // base + 8 * index, descriptors as a constant high word plus (address >> 4), parameters copied out of the constant
      // bank (the asm memory clobbers made the compiler reload them every tap).
      const int taps = p.taps, kchunks = p.kchunks, mt = p.mt, n_a = p.a_stages, n_b = p.b_stages;
      const bool b_res = p.b_resident != 0, sw64 = p.kc == 32;
      const int kfull = p.kc / 16, klast = p.kinstr_last;
      const int n_instr = p.split_n ? p.n_tile / 2 : p.n_tile;
      const uint32_t idesc = (1u << 4)                          // D format: F32
                             | (0u << 7) | (0u << 10)           // A, B format: F16
                             | ((uint32_t)(n_instr >> 3) << 17)
                             | ((uint32_t)((kPair ? 2 * kBlockM : kBlockM) >> 4) << 24);
      // my share of every K step: sub-tiles j0, j0 + jstep, ... and the weight rows / accumulator columns from b_off / d_off
      const int j0 = p.split_n ? 0 : iss, jstep = p.split_n ? 1 : p.n_iss;
      const uint32_t d_off = p.split_n ? (uint32_t)(iss * (p.n_tile / 2)) : 0u;
      // operand addresses in 16-byte units (what the descriptor's start-address field holds); the & 0x3fff drops the cluster
      // rank bits of a shared::cluster-form address once, every later offset is a plain add (all offsets stay < 256 KB)
      const uint32_t a_base = (smem_u32(smem_a) >> 4) & 0x3fffu, a_bytes = (uint32_t)p.a_stage_bytes >> 4;
      const uint32_t b_base = ((smem_u32(smem_b) + (p.split_n ? (uint32_t)(iss * (p.b_rows / 2) * p.row_bytes) : 0u)) >> 4) & 0x3fffu,
                     b_bytes = (uint32_t)p.b_bytes >> 4;
      const uint32_t sub_bytes = (uint32_t)p.sub_bytes >> 4, sub_cols = (uint32_t)p.sub_cols;
      const uint32_t bar_a_full = smem_u32(a_full), bar_a_empty = smem_u32(a_empty), bar_b_full = smem_u32(b_full),
                     bar_b_empty = smem_u32(b_empty), bar_tfull = smem_u32(tfull_bar), bar_tempty = smem_u32(tempty_bar);
      // descriptor = constant high word (SBO, version, swizzle mode) | start address >> 4 (LBO 0, base offset 0)
      const uint64_t desc_hi = (uint64_t)((uint32_t)((sw64 ? 512 : 1024) >> 4) | (1u << 14) | ((sw64 ? 4u : 2u) << 29)) << 32;
      const bool timed = p.dbg != nullptr;
      uint32_t a_stage = 0, b_stage = 0, a_phase = 0, b_phase = 0, acc = 0, acc_phase = 0;
      uint32_t sa = a_base;                // address of A stage a_stage
      bool ok = true;
      bool b_loaded = false;
      long long w_a = 0, w_b = 0, w_t = 0;
      for (int tile = tile0; ok && tile < total_tiles; tile += tile_step) {
        ok = __all_sync(0xffffffffu, mbar_wait_addr(bar_tempty + 8u * acc, acc_phase ^ 1, p.err, 102, w_t, timed));
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u + d_off;
        uint32_t sb = b_base;              // resident weights: slot kc * taps + t, walked in order
        for (int kc = 0; ok && kc < kchunks; ++kc) {
          if (!kStrided) {
            ok = __all_sync(0xffffffffu, mbar_wait_addr(bar_a_full + 8u * a_stage, a_phase, p.err, 103, w_a, timed));
            if (!ok) break;
          }
          const int kinstr = (kc == kchunks - 1) ? klast : kfull;
          if (!kStrided && b_res && b_loaded) {
            // resident weights, already loaded: nothing to wait for inside the K chunk, so all taps go out in ONE elected block
            // (per-tap elect / vote / reconvergence bookkeeping is what limits the N <= 64 layers: their instructions need
            // only 16-32 tensor-pipe cycles each)
            if (elect_one()) {
              uint32_t sbt = sb;
              for (int t = 0; t < taps; ++t) {
                const uint64_t db = desc_hi | (uint64_t)sbt;
                const uint32_t first = (kc | t) != 0 ? 1u : 0u;
                const uint32_t a_tap = sa + ((uint32_t)p.tap_off[t] >> 4);
                for (int j = j0; j < mt; j += jstep) {
                  const uint64_t da = desc_hi | (uint64_t)(a_tap + (uint32_t)j * sub_bytes);
                  const uint32_t d = d_tmem + (uint32_t)j * sub_cols;
                  if (kinstr == 4) {
                    umma_f16<kPair>(d, da, db, idesc, first);
                    umma_f16<kPair>(d, da + 2, db + 2, idesc, 1u);
                    umma_f16<kPair>(d, da + 4, db + 4, idesc, 1u);
                    umma_f16<kPair>(d, da + 6, db + 6, idesc, 1u);
                  } else {
                    umma_f16<kPair>(d, da, db, idesc, first);
                    if (kinstr > 1) umma_f16<kPair>(d, da + 2, db + 2, idesc, 1u);
                    if (kinstr > 2) umma_f16<kPair>(d, da + 4, db + 4, idesc, 1u);
                  }
                }
                sbt += b_bytes;
              }
              umma_commit_addr<kPair>(bar_a_empty + 8u * a_stage);
            }
            __syncwarp();
            sb += (uint32_t)taps * b_bytes;
            sa += a_bytes;
            if (++a_stage == (uint32_t)n_a) { a_stage = 0; a_phase ^= 1; sa = a_base; }
            continue;
          }
          for (int t = 0; t < taps; ++t) {
            if (kStrided) {
              ok = __all_sync(0xffffffffu, mbar_wait_addr(bar_a_full + 8u * a_stage, a_phase, p.err, 103, w_a, timed));
              if (!ok) break;
            }
            if (!b_res) sb = b_base + b_stage * b_bytes;
            if (!b_res || !b_loaded) {
              const uint32_t slot = b_res ? (uint32_t)(kc * taps + t) : b_stage;
              ok = __all_sync(0xffffffffu, mbar_wait_addr(bar_b_full + 8u * slot, b_res ? 0u : b_phase, p.err, 106, w_b, timed));
              if (!ok) break;
            }
            tc_fence_after();
            if (elect_one()) {
              const uint64_t db = desc_hi | (uint64_t)sb;
              const uint32_t first = (kc | t) != 0 ? 1u : 0u;
              const uint32_t a_tap = sa + (kStrided ? 0u : (uint32_t)p.tap_off[t] >> 4);
              for (int j = j0; j < mt; j += jstep) {
                const uint64_t da = desc_hi | (uint64_t)(a_tap + (uint32_t)j * sub_bytes);
                const uint32_t d = d_tmem + (uint32_t)j * sub_cols;
                // advance 16 elements (32 bytes) along K inside the swizzle atom; all-zero K slices are skipped
                if (kinstr == 4) {
                  umma_f16<kPair>(d, da, db, idesc, first);
                  umma_f16<kPair>(d, da + 2, db + 2, idesc, 1u);
                  umma_f16<kPair>(d, da + 4, db + 4, idesc, 1u);
                  umma_f16<kPair>(d, da + 6, db + 6, idesc, 1u);
                } else {
                  umma_f16<kPair>(d, da, db, idesc, first);
                  if (kinstr > 1) umma_f16<kPair>(d, da + 2, db + 2, idesc, 1u);
                  if (kinstr > 2) umma_f16<kPair>(d, da + 4, db + 4, idesc, 1u);
                }
              }
              if (!b_res) umma_commit_addr<kPair>(bar_b_empty + 8u * b_stage);
              if (kStrided) umma_commit_addr<kPair>(bar_a_empty + 8u * a_stage);
            }
            __syncwarp();
            if (b_res) {
              sb += b_bytes;
            } else if (++b_stage == (uint32_t)n_b) {
              b_stage = 0;
              b_phase ^= 1;
            }
            if (kStrided) {
              sa += a_bytes;
              if (++a_stage == (uint32_t)n_a) { a_stage = 0; a_phase ^= 1; sa = a_base; }
            }
          }
          if (!ok) break;
          if (!kStrided) {
            if (elect_one()) umma_commit_addr<kPair>(bar_a_empty + 8u * a_stage);
            __syncwarp();
            sa += a_bytes;
            if (++a_stage == (uint32_t)n_a) { a_stage = 0; a_phase ^= 1; sa = a_base; }
          }
        }
        if (!ok) break;
        b_loaded = true;
        if (elect_one()) umma_commit_addr<kPair>(bar_tfull + 8u * acc);
        __syncwarp();
        if (p.acc_bufs == 2) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0 && iss == 0) {
        p.dbg[6] = (unsigned long long)w_a;
        p.dbg[7] = (unsigned long long)w_b;
        p.dbg[8] = (unsigned long long)w_t;
      }
    }
  }
  } else {
    // ===================== epilogue (warps 4..11) =====================
    // Work split: TMEM lane quarter q = warp % 4 (hardware rule), h = the other bit.  Layers with n_tile <= 64 and two
    // sub-tiles give sub-tile h to the warp (all columns); otherwise h selects a column half.  A warp's columns are walked
    // in chunks of 64 (one tcgen05.ld.x64) with a trailing 32-column chunk where the span needs it; everything per chunk
    // that does not depend on the data (tile decode, row decode, accumulator hand-shake) is hoisted to the tile / sub-tile
    // level: ncu showed ~240 of the ~340 instructions per 32-column chunk of the previous flat loop were such overhead, and
    // with two epilogue warps per scheduler the layers with N <= 128 ran at the latency of that instruction stream.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;
    const int row_in_tile = q * 32 + lane;
    const bool row_split = p.epi_row_split != 0;
    const int split = ((p.n_tile / 2 + 31) / 32) * 32;
    const int col_lo = row_split ? 0 : (half ? split : 0);
    const int col_hi = row_split ? p.n_tile : (half ? p.n_tile : (split < p.n_tile ? split : p.n_tile));
    const int span = col_hi > col_lo ? col_hi - col_lo : 0;
    const int n64 = span >> 6;                                  // 64-column chunks per sub-tile
    const int nch = n64 + (((span & 63) + 31) >> 5);            // + one 32-column chunk for the rest
    const int j_lo = row_split ? half : 0, j_n = row_split ? 1 : p.mt;
    // per-channel vectors, as shared-memory byte addresses (explicit ld.shared: through a generic pointer these were LD.E)
    const uint32_t v_scale = smem_u32(vec), v_bias = v_scale + 4u * p.vec_n, v_slope = v_scale + 8u * p.vec_n,
                   v_scale2 = v_scale + 12u * p.vec_n, v_bias2 = v_scale + 16u * p.vec_n;
    // this warp's 2 KB transpose buffer: 32 rows x 64 B (32 fp16 channels), 16-byte pieces XOR-swizzled by
    // (row >> 1) & 3 so that both the row-wise writes and the 4-lanes-per-row read-back are conflict-free
    uint8_t* stg = stage_buf + (warp - kEpiWarp0) * 2048;
    const uint32_t stg_a = smem_u32(stg);
    const uint32_t stg_w = stg_a + (uint32_t)(lane * 64);                  // my row, as writer
    const int w_sw = (lane >> 1) & 3;
    const int rb_piece = lane & 3;                                          // as reader: piece rb_piece of row i*8 + lane/4
    const uint32_t tempty_l0 = kPair ? leader_addr(&tempty_bar[0]) : 0u, tempty_l1 = kPair ? leader_addr(&tempty_bar[1]) : 0u;
    const bool use_tma_store = !kStrided && p.tma_store != 0;
    const int n_tiles = p.n_tiles, n_tile = p.n_tile, sub_cols = p.sub_cols;
    const int out_c_store = p.out_c_store;
    const int step_q = tile_step / n_tiles, step_r = tile_step - step_q * n_tiles;      // tile -> (m tile, n tile) without divisions
    int acc = 0;
    uint32_t acc_phase = 0;
    bool ok = true;
    long long w_full = 0, t_epi = 0, t_ld = 0;
    const bool timed = p.dbg != nullptr;
    const int my_tiles = tile0 < total_tiles ? (total_tiles - tile0 + tile_step - 1) / tile_step : 0;
    auto tile_done = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(acc ? tempty_l1 : tempty_l0);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (p.acc_bufs == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    };
    if (nch == 0) {
      // no columns for this warp: it only takes part in the accumulator hand-shake
      for (int ti = 0; ti < my_tiles; ++ti) {
        ok = __all_sync(0xffffffffu, mbar_wait(&tfull_bar[acc], acc_phase, p.err, 104));
        if (!ok) break;
        tile_done();
      }
    } else {
      // position in this warp's chunk sequence: (tile, sub-tile, chunk), with the tile decoded incrementally
      struct Pos {
        int j, c, mt_idx, nt;
      };
      auto advance = [&](Pos& s) {
        if (++s.c == nch) {
          s.c = 0;
          if (++s.j == j_n) {
            s.j = 0;
            s.mt_idx += step_q;
            s.nt += step_r;
            if (s.nt >= n_tiles) { s.nt -= n_tiles; ++s.mt_idx; }
          }
        }
      };
      auto first_row = [&](const Pos& s) {
        return (long long)(kPair ? 2 * s.mt_idx + cta_rank : s.mt_idx) * tile_rows + (j_lo + s.j) * kBlockM + row_in_tile;
      };
      auto chunk_col = [&](int c) { return col_lo + (c < n64 ? c * 64 : n64 * 64 + (c - n64) * 32); };
      const int G = my_tiles * j_n * nch;
      const int t0_m = tile0 / n_tiles;
      // residual rows are fetched TWO chunks ahead of their use (64-column chunks: the same bytes in flight as the four
      // 32-column chunks of the previous loop) into two register sets, static slot through the 2x unrolled loop
      uint4 R[2][8];
      Pos pf{0, 0, t0_m, tile0 - t0_m * n_tiles};
      int pf_g = 0;
      RowInfo pf_row;
      pf_row.valid = false;
      pf_row.orow = 0;
      auto res_load = [&](uint4 (&dst)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = make_uint4(0, 0, 0, 0);
        if (pf_g >= G) return;
        if (pf.c == 0) pf_row = row_info<kStrided>(p, first_row(pf));
        const int ch = pf.nt * n_tile + chunk_col(pf.c);
        const int wpieces = pf.c < n64 ? 8 : 4;
        if (pf_row.valid) {
          const __half* src = p.residual + pf_row.orow * p.res_cp + ch;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < wpieces && ch + e * 8 < out_c_store) dst[e] = *(const uint4*)(src + e * 8);
        }
        ++pf_g;
        advance(pf);
      };
      if (kRes) {
        res_load(R[0]);
        res_load(R[1]);
      }
      Pos cur{0, 0, t0_m, tile0 - t0_m * n_tiles};
      int n0 = 0;
      long long row0 = 0;
      RowInfo ri;
      ri.valid = false;
      ri.orow = 0;
      // rows this lane writes back after the transpose (per-lane store path): i*8 + lane/4, i = 0..3, as element offsets of piece rb_piece
      long long woff[4] = {0, 0, 0, 0}, woff2[4] = {0, 0, 0, 0};
      bool wvalid[4] = {false, false, false, false};
      int prow32 = 0;
      long long te0 = 0;

      // one 32-column piece of a chunk: y = acc*scale+bias (+residual) -> activation -> store (+ second output)
      auto piece = [&](const uint32_t* v, const uint4* res4, int ch) {
        float y[32];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 s4 = ld_shared_f4(v_scale + 4u * (uint32_t)(ch + e * 4));
          const float4 b4 = ld_shared_f4(v_bias + 4u * (uint32_t)(ch + e * 4));
          y[e * 4 + 0] = fmaf(__uint_as_float(v[e * 4 + 0]), s4.x, b4.x);
          y[e * 4 + 1] = fmaf(__uint_as_float(v[e * 4 + 1]), s4.y, b4.y);
          y[e * 4 + 2] = fmaf(__uint_as_float(v[e * 4 + 2]), s4.z, b4.z);
          y[e * 4 + 3] = fmaf(__uint_as_float(v[e * 4 + 3]), s4.w, b4.w);
        }
        if (kOutMode == 2) {
          if (ri.valid) {
            float* o = p.out_f32 + ri.orow * p.out_f32_stride + ch;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (ch + e * 4 < p.out_f32_cols) *(float4*)(o + e * 4) = make_float4(y[e * 4], y[e * 4 + 1], y[e * 4 + 2], y[e * 4 + 3]);
          }
          return;
        }
        if (kRes) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f[8];
            unpack_h8(res4[e], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) y[e * 8 + k] += f[k];
          }
        }
        if (kAct == PCB_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 32; ++e) y[e] = fmaxf(y[e], 0.f);
        } else if (kAct == PCB_ACT_PRELU) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 a4 = ld_shared_f4(v_slope + 4u * (uint32_t)(ch + e * 4));
            y[e * 4 + 0] = y[e * 4 + 0] >= 0.f ? y[e * 4 + 0] : y[e * 4 + 0] * a4.x;
            y[e * 4 + 1] = y[e * 4 + 1] >= 0.f ? y[e * 4 + 1] : y[e * 4 + 1] * a4.y;
            y[e * 4 + 2] = y[e * 4 + 2] >= 0.f ? y[e * 4 + 2] : y[e * 4 + 2] * a4.z;
            y[e * 4 + 3] = y[e * 4 + 3] >= 0.f ? y[e * 4 + 3] : y[e * 4 + 3] * a4.w;
          }
        }
        if (kOutMode == 1) {
          if (ri.valid) {
            float* o = p.out_s32 + ri.orow * p.out_cp + ch;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (ch + e * 4 < out_c_store) *(float4*)(o + e * 4) = make_float4(y[e * 4], y[e * 4 + 1], y[e * 4 + 2], y[e * 4 + 3]);
          }
          return;
        }
        auto second = [&]() {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 s4 = ld_shared_f4(v_scale2 + 4u * (uint32_t)(ch + e * 4));
            const float4 b4 = ld_shared_f4(v_bias2 + 4u * (uint32_t)(ch + e * 4));
            y[e * 4 + 0] = fmaf(y[e * 4 + 0], s4.x, b4.x);
            y[e * 4 + 1] = fmaf(y[e * 4 + 1], s4.y, b4.y);
            y[e * 4 + 2] = fmaf(y[e * 4 + 2], s4.z, b4.z);
            y[e * 4 + 3] = fmaf(y[e * 4 + 3], s4.w, b4.w);
          }
        };
        if (use_tma_store) {
          // The piece (32 consecutive P-rows x 32 channels) is staged in the 64B-swizzled layout TMA expects and leaves
          // through ONE bulk tensor store per output: no read-back, no per-row address arithmetic, asynchronous.  Rows
          // that are not image pixels are staged as zeros, so the padding of the output stays zero; rows past the end of
          // the tensor and channels past its width are clipped by the tensor map.
          uint4 pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = ri.valid ? pack_h8(y + e * 8) : make_uint4(0, 0, 0, 0);
          if (lane == 0) tma_store_wait_read();       // the previous store has finished reading the staging buffer
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 4; ++e) st_shared_v4(stg_w + (uint32_t)(((e ^ w_sw) & 3) * 16), pk[e]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmO, stg_a, ch, prow32);
          if (kOut2) {
            second();
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[e] = ri.valid ? pack_h8(y + e * 8) : make_uint4(0, 0, 0, 0);
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 4; ++e) st_shared_v4(stg_w + (uint32_t)(((e ^ w_sw) & 3) * 16), pk[e]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(&tmO2, stg_a, ch, prow32);
          }
        } else {
          // transpose through shared memory so that 4 lanes write one row's 64 contiguous bytes (full sectors)
          const bool piece_ok = ch + rb_piece * 8 < out_c_store;
          __syncwarp();      // the previous piece's read-back is done before the buffer is rewritten
#pragma unroll
          for (int e = 0; e < 4; ++e) st_shared_v4(stg_w + (uint32_t)(((e ^ w_sw) & 3) * 16), pack_h8(y + e * 8));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            const uint4 o4 = ld_shared_v4(stg_a + (uint32_t)(r * 64 + ((rb_piece ^ (r >> 1)) & 3) * 16));
            if (wvalid[i] && piece_ok) *(uint4*)(p.out + woff[i] + ch) = o4;
          }
          if (kOut2) {
            second();
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 4; ++e) st_shared_v4(stg_w + (uint32_t)(((e ^ w_sw) & 3) * 16), pack_h8(y + e * 8));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2);
              const uint4 o4 = ld_shared_v4(stg_a + (uint32_t)(r * 64 + ((rb_piece ^ (r >> 1)) & 3) * 16));
              if (wvalid[i] && piece_ok) *(uint4*)(p.out2 + woff2[i] + ch) = o4;
            }
          }
        }
      };

      for (int g0 = 0; ok && g0 < G; g0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (g0 + u >= G) break;
          if (cur.j == 0 && cur.c == 0) {
            // first chunk of a tile: wait for its accumulator
            n0 = cur.nt * n_tile;
            ok = __all_sync(0xffffffffu, mbar_wait_t(&tfull_bar[acc], acc_phase, p.err, 104, w_full, timed));
            if (!ok) break;
            tc_fence_after();
          }
          if (timed) te0 = clock64();
          if (cur.c == 0) {
            row0 = first_row(cur);
            ri = row_info<kStrided>(p, row0);
            prow32 = (int)(row0 - lane);                      // first P-row of this warp's 32 rows (TMA store coordinate)
            if (kOutMode == 0 && !use_tma_store) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int src = i * 8 + (lane >> 2);
                const long long wr = __shfl_sync(0xffffffffu, ri.orow, src);
                woff[i] = wr * p.out_cp + rb_piece * 8;
                if (kOut2) woff2[i] = wr * p.out2_cp + rb_piece * 8;
                wvalid[i] = __shfl_sync(0xffffffffu, ri.valid ? 1 : 0, src) != 0;
              }
            }
          }
          const bool wide = cur.c < n64;
          const int col = chunk_col(cur.c);
          const bool last = cur.j == j_n - 1 && cur.c == nch - 1;
          const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + (j_lo + cur.j) * sub_cols + col);
          const int ch = n0 + col;
          __syncwarp();   // tcgen05.ld is .sync.aligned
          long long tl0 = 0;
          if (timed) tl0 = clock64();
          if (wide) {
            uint32_t v[64];
            tmem_ld64(t_addr, v);
            if (timed) t_ld += clock64() - tl0;
            if (last) tile_done();          // the accumulator is in registers: hand it back before the arithmetic
            piece(v, R[u], ch);
            piece(v + 32, R[u] + 4, ch + 32);
          } else {
            uint32_t v[32];
            tmem_ld32(t_addr, v);
            if (timed) t_ld += clock64() - tl0;
            if (last) tile_done();
            piece(v, R[u], ch);
          }
          if (kRes) res_load(R[u]);     // this slot's next use: two chunks from now
          advance(cur);
          if (timed) t_epi += clock64() - te0;
        }
      }
    }
    if (lane == 0) tma_store_wait_all();     // every bulk store of this thread has been written before the CTA exits
    if (p.dbg && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0) {
      p.dbg[9] = (unsigned long long)w_full;
      p.dbg[10] = (unsigned long long)t_epi;
      p.dbg[11] = (unsigned long long)t_ld;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // no CTA leaves while its peer can still commit to / arrive on its barriers or read its operands
  if (warp == 1) {
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
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
bool make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows,
                 int kc = kKC) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 4-D fp16 map over the P-layout [n][hp][wp][cp] for stride-2 taps: box {kc channels, 2*bx, 2*by, nb} traversed with
// element strides {1, 2, 2, 1} -> nb * by * bx rows of kc channels in shared memory (128B / 64B swizzle as the 2-D maps).
bool make_map_4d_s2(CUtensorMap* m, const void* base, int n, int hp, int wp, int cp, int kc, int bx, int by, int nb) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)cp, (cuuint64_t)wp, (cuuint64_t)hp, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)cp * 2, (cuuint64_t)wp * cp * 2, (cuuint64_t)hp * wp * cp * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(2 * bx), (cuuint32_t)(2 * by), (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 2, 2, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

// Fills the A-stage plan (loads, tap offsets, stage bytes) for `mt` sub-tiles.  Returns false if no
// plan exists (box limits).
bool plan_a(Conv2Params& p, int mt, int* a2_rows) {
  p.mt = mt;
  p.n_aloads = 0;
  *a2_rows = 8;
  auto add = [&](int row_rel, int off, int map2) {
    if (p.n_aloads >= kMaxALoads) return false;
    p.aloads[p.n_aloads++] = ALoad{row_rel, off, map2};
    return true;
  };
  if (p.taps == 1) {
    for (int j = 0; j < mt; ++j) add(j * kBlockM, j * p.sub_bytes, 0);
    p.tap_off[0] = 0;
    p.a_stage_bytes = mt * p.sub_bytes;
    p.a_tx_bytes = mt * p.sub_bytes;
    return true;
  }
  const int halo = p.wp + 1;
  // merged: one contiguous halo range = mt 128-row boxes + the 2*halo extra rows as one or two small boxes (<= 256 rows each)
  const int merged_extra = pcb_round_up(2 * halo, 16);
  const int n_small = merged_extra <= 256 ? 1 : 2;
  const long long merged_rows = (long long)mt * kBlockM + merged_extra;
  const long long banded_rows = 3LL * (mt * kBlockM + 8);
  if (merged_extra <= 512 && merged_rows <= banded_rows) {
    for (int j = 0; j < mt; ++j) add(-halo + j * kBlockM, j * p.sub_bytes, 0);
    const int small = merged_extra / n_small;
    for (int sidx = 0; sidx < n_small; ++sidx) add(-halo + mt * kBlockM + sidx * small, (mt * kBlockM + sidx * small) * p.row_bytes, 1);
    *a2_rows = small;
    for (int t = 0; t < 9; ++t) p.tap_off[t] = (halo + (t / 3 - 1) * p.wp + (t % 3 - 1)) * p.row_bytes;
    p.a_tx_bytes = (int)merged_rows * p.row_bytes;
    p.a_stage_bytes = pcb_round_up(p.a_tx_bytes, 1024);
    return true;
  }
  const int band_bytes = pcb_round_up((mt * kBlockM + 8) * p.row_bytes, 1024);
  for (int b = 0; b < 3; ++b) {
    for (int j = 0; j < mt; ++j)
      if (!add((b - 1) * p.wp - 1 + j * kBlockM, b * band_bytes + j * p.sub_bytes, 0)) return false;
    if (!add((b - 1) * p.wp - 1 + mt * kBlockM, b * band_bytes + mt * p.sub_bytes, 1)) return false;
  }
  for (int t = 0; t < 9; ++t) p.tap_off[t] = (t / 3) * band_bytes + (t % 3) * p.row_bytes;
  p.a_tx_bytes = (int)banded_rows * p.row_bytes;
  p.a_stage_bytes = 3 * band_bytes;
  return true;
}

}  // namespace

int pcb_conv_tc2(pcb_ctx* c, const ConvArgs& a) {
  const PTensor& in = *a.in;
  const ConvWeights& w = *a.w;
  if (a.residual && a.residual->f32) return pcb_conv_tc(c, a);   // fp32 residual stream: baseline kernel only
  Conv2Params p{};
  p.rows = (int)in.rows();
  p.dense = in.dense ? 1 : 0;
  p.hp = in.dense ? 0 : in.h + in.pad;
  p.wp = in.dense ? 0 : in.w + in.pad;
  p.pad_lo = in.pad_lo;
  p.pad_hi = in.pad - in.pad_lo;
  p.pad_lo_out = in.pad_lo;
  p.taps = w.taps;
  p.cin_w = w.cin_w;
  {
    const int cin_eff = in.dense ? w.cin : (w.taps == 1 && w.cin == 3) ? 27 : w.cin;   // stem: 27 patch channels
    // <= 32 input channels (SCRFD stems, patch tensors): 64-byte operand rows halve the shared memory per stage, which
    // buys the deeper A pipeline those wide-map layers need (one K chunk per tile: a stage is only free per finished tile)
    p.kc = (!in.dense && cin_eff <= 32 && !env_int("PCB_CONV_NO_NARROW", 0)) ? 32 : kKC;
    p.row_bytes = 2 * p.kc;
    p.sub_bytes = kBlockM * p.row_bytes;
    p.kchunks = (cin_eff + p.kc - 1) / p.kc;
    p.kinstr_last = (cin_eff - (p.kchunks - 1) * p.kc + 15) / 16;
  }
  p.n_tile = w.n_tile;
  p.n_tiles = w.npad / w.n_tile;
  p.vec_n = pcb_round_up(w.npad, 32);
  p.vec_real = w.npad;
  p.stride = in.dense ? 1 : a.stride;
  p.act = a.act;
  p.scale = w.scale;
  p.bias = w.bias;
  p.slope = w.slope;
  p.err = c->d_err;
  {
    auto fd = [](int d, uint32_t* mul, int* shift) {
      int sft = 0;
      while ((1LL << sft) < d) ++sft;
      *mul = (uint32_t)((((1ULL << sft) - (unsigned long long)d) << 32) / (unsigned long long)d + 1ULL);
      *shift = sft;
    };
    fd(p.hp * p.wp > 0 ? p.hp * p.wp : 1, &p.fd_plane_mul, &p.fd_plane_shift);
    fd(p.wp > 0 ? p.wp : 1, &p.fd_wp_mul, &p.fd_wp_shift);
  }
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
    // stride-1 layers address output rows as input rows: same layout on both sides (and on the residual / second output)
    if (!in.dense && a.stride == 1 && (out.pad != in.pad || out.pad_lo != in.pad_lo))
      return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: a stride-1 layer cannot change the activation layout");
    if ((a.residual && (a.residual->pad != out.pad || a.residual->pad_lo != out.pad_lo)) ||
        (a.out2 && (a.out2->pad != out.pad || a.out2->pad_lo != out.pad_lo)))
      return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: residual / second output must share the output's layout");
    if (out.cp > w.npad) return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: output channels exceed packed weight rows");
    if (a.residual) {
      p.residual = a.residual->data;
      p.res_cp = a.residual->cp;
    }
    if (a.out2) {
      if (a.out2->cp != out.cp || a.out2->f32) return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: out2 geometry");
      p.out2 = a.out2->data;
      p.out2_cp = a.out2->cp;
      p.scale2 = a.scale2;
      p.bias2 = a.bias2;
    }
  }
  if (w.n_tile % 16 || w.n_tile > 256 || w.n_tile < 16) return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: bad n_tile");
  p.sub_cols = pcb_round_up(w.n_tile, 32);
  const int ksteps = p.taps * p.kchunks;
  const int fixed = 5 * p.vec_n * 4 + 1024 + kEpiWarps * 2048 + (2 * kMaxA + 2 * kMaxB + 4) * 8 + 16 + 1024;
  const int room = kSmemBudget - fixed;

  // stride-2 convolutions: GEMM on the output grid, one strided 4-D TMA load per tap (see Conv2Params::strided)
  const bool strided = !in.dense && a.stride == 2 && a.out && !env_int("PCB_CONV_NO_STRIDED", 0);
  if (strided) {
    const int Ho = a.out->h, Wo = a.out->w;
    int bx = 1;
    for (int d = 1; d <= Wo && d <= 128; ++d) if (Wo % d == 0) bx = d;
    int by = 1;
    for (int d = 1; d <= Ho && d * bx <= 128; ++d) if (Ho % d == 0) by = d;
    int nb = 1;
    if (bx == Wo && by == Ho) nb = 128 / (bx * by);
    if (nb > in.n) nb = in.n;
    if (nb < 1) nb = 1;
    p.strided = 1;
    p.s_bx = bx; p.s_by = by; p.s_nb = nb;
    p.s_tx = Wo / bx; p.s_ty = Ho / by;
    p.s_n = in.n; p.s_ho = Ho; p.s_wo = Wo;
    p.mt = 1;
    p.n_aloads = 0;
    p.a_stage_bytes = kBlockM * p.row_bytes;
    p.a_tx_bytes = nb * by * bx * p.row_bytes;
    p.m_tiles = ((in.n + nb - 1) / nb) * p.s_tx * p.s_ty;
    p.acc_bufs = 2;
  }

  // Pair mode (2-CTA cluster, cta_group::2): PCB_CONV_PAIR = 0 never, 1 (default) by the policy below, 2 wherever it is legal.
  // Policy from per-layer A/B runs on B200 (tools/profile_layers.py with PCB_CONV_PAIR=0/2, profiles/): see DESIGN.md 3.1.
  static const int pair_mode = env_int("PCB_CONV_PAIR", 1);
  const long long m_tiles128 = ((long long)p.rows + kBlockM - 1) / kBlockM;
  const bool pair_legal = !strided && (c->num_sms % 2 == 0) && m_tiles128 >= 4;
  const bool pair_wanted = pair_mode == 2 || (pair_mode == 1 && w.n_tile >= kPairMinN && p.kchunks >= 2 && m_tiles128 * p.n_tiles >= 2 * c->num_sms);
  p.pair = (pair_legal && pair_wanted) ? 1 : 0;
  p.b_rows = p.pair ? w.n_tile / 2 : w.n_tile;
  p.b_bytes = p.b_rows * p.row_bytes;
  const int units = p.pair ? c->num_sms / 2 : c->num_sms;     // tiles in flight: one per CTA, or one per CTA pair

  // choose MT (1 or 2): fewer L2 bytes per FLOP at MT=2, but half as many tiles to spread over the SMs
  const int force_mt = env_int("PCB_CONV_MT", 0);
  int best_mt = 0;
  double best_cost = 0.0;
  int a2_rows = 8;
  for (int mt = 1; mt <= 2 && !strided; ++mt) {
    if (force_mt && mt != force_mt) continue;
    if (mt * p.sub_cols > (int)kTmemCols) continue;
    Conv2Params q = p;
    int a2 = 8;
    if (!plan_a(q, mt, &a2)) continue;
    const int min_b = ksteps < 3 ? ksteps : 3;
    if (2 * q.a_stage_bytes + min_b * q.b_bytes + fixed > kSmemBudget) continue;
    const long long unit_rows = (long long)mt * kBlockM * (p.pair ? 2 : 1);
    const long long m_tiles = ((long long)p.rows + unit_rows - 1) / unit_rows;
    const long long tiles = m_tiles * p.n_tiles;
    const long long waves = (tiles + units - 1) / units;
    // relative time per tile, measured on B200 (tools/profile_layers.py, PCB_CONV_MT): MT=2 shares each weight stage
    // between two accumulators and is ~1.4x faster per row while TMEM can still be double-buffered (2*N <= 256);
    // for N > 128 it would be single-buffered and the exposed epilogue costs more than the saved weight traffic
    const double cost = (double)waves * mt * (mt == 1 ? 1.0 : (2 * p.sub_cols > 256 ? 1.05 : 0.72));
    if (!best_mt || cost < best_cost) { best_mt = mt; best_cost = cost; }
  }
  if (!strided) {
    if (!best_mt) return pcb_conv_tc(c, a);   // no shared-memory plan (very wide maps): baseline kernel
    plan_a(p, best_mt, &a2_rows);
    if (env_int("PCB_TAP_ALIGN", 0)) for (int t = 0; t < 9; ++t) p.tap_off[t] &= ~1023;   // timing experiment only (wrong sums)
    const long long unit_rows = (long long)p.mt * kBlockM * (p.pair ? 2 : 1);
    p.m_tiles = (int)(((long long)p.rows + unit_rows - 1) / unit_rows);
    p.acc_bufs = (p.mt * p.sub_cols <= 256) ? 2 : 1;
  }
  // two MMA issuers where the K step can be split by accumulator (PCB_CONV_ISS=1 forces one)
  static const int iss_mode = env_int("PCB_CONV_ISS", 2);
  // MT = 1 layers split the N extent between the two issuers, pair layers excepted: there each half-N instruction re-reads the
  // whole A operand of both CTAs from shared memory for 64 tensor cycles of work, and one issuer keeps the pipe fed
  // (same-box A/B, gpurun r2r: 14x14 / 7x7 / 28x28->256 layers +3 % without the split, the stride-2 layers -3 %)
  static const int split_pair = env_int("PCB_CONV_SPLIT_PAIR", 0);
  p.n_iss = 1;
  p.split_n = 0;
  if (iss_mode >= 2) {
    if (p.mt == 2) p.n_iss = 2;
    else if (w.n_tile % 32 == 0 && (!p.pair || split_pair)) { p.n_iss = 2; p.split_n = 1; }
  }
  // epilogue split: narrow layers with two sub-tiles give each warp of a lane quarter its own sub-tile (one 64-column chunk
  // per tile, and no idle warps when n_tile <= 32); everything else splits the columns
  p.epi_row_split = (p.mt == 2 && w.n_tile <= 64 && !env_int("PCB_EPI_NO_ROWSPLIT", 0)) ? 1 : 0;
  p.b_nloads = (p.pair && p.split_n) ? 2 : 1;
  p.b_load_rows = p.b_rows / p.b_nloads;
  p.b_resident = 0;
  if (p.n_tiles == 1 && ksteps <= kMaxB && 2 * p.a_stage_bytes + ksteps * p.b_bytes <= room && !env_int("PCB_CONV_NO_RESIDENT", 0)) {
    p.b_resident = 1;
    p.b_stages = ksteps;
    p.a_stages = (room - ksteps * p.b_bytes) / p.a_stage_bytes;
  } else {
    // two A stages, the rest to B; a third A stage when B already has >= 6
    p.a_stages = 2;
    p.b_stages = (room - 2 * p.a_stage_bytes) / p.b_bytes;
    if (p.b_stages > 8 && (room - 3 * p.a_stage_bytes) / p.b_bytes >= 6) {
      p.a_stages = 3;
      p.b_stages = (room - 3 * p.a_stage_bytes) / p.b_bytes;
    }
    if (p.b_stages > kMaxB) p.b_stages = kMaxB;
  }
  if (strided && !p.b_resident) {
    // every tap is its own A stage: split the room evenly between the A and B rings
    p.a_stages = (room / 2) / p.a_stage_bytes;
    p.b_stages = (room - p.a_stages * p.a_stage_bytes) / p.b_bytes;
    if (p.b_stages > kMaxB) p.b_stages = kMaxB;
  }
  if (p.a_stages > kMaxA) p.a_stages = kMaxA;
  const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * p.b_bytes + fixed;

  CUtensorMap tmA, tmA2, tmB, tmO, tmO2;
  if (strided) {
    if (!make_map_4d_s2(&tmA, in.data, in.n, in.h + in.pad, in.w + in.pad, in.cp, p.kc, p.s_bx, p.s_by, p.s_nb))
      return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(A strided) failed");
    tmA2 = tmA;
  } else {
    if (!make_map_2d(&tmA, in.data, (uint64_t)p.rows, (uint64_t)in.cp, (uint64_t)in.cp, kBlockM, p.kc))
      return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed");
    if (!make_map_2d(&tmA2, in.data, (uint64_t)p.rows, (uint64_t)in.cp, (uint64_t)in.cp, (uint32_t)a2_rows, p.kc))
      return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(A2) failed");
  }
  if (!make_map_2d(&tmB, w.w, (uint64_t)w.npad, (uint64_t)w.taps * w.cin_w, (uint64_t)w.taps * w.cin_w, (uint32_t)p.b_load_rows, p.kc))
    return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed");

  // TMA stores for the fp16 outputs of stride-1 spatial layers (output P-rows = input P-rows).  PCB_CONV_TMA_STORE = 0 per-lane
  // stores everywhere, 1 (default) TMA stores on maps at least 24 pixels wide, 2 wherever legal.  Measured (same library, same box):
  // the epilogue-bound layers gain 3-10 % (28x28 +res+out2 -6.5 %, stem -10 %), the 14x14 / 7x7 layers LOSE ~2 %: there the
  // padding rows the store also writes (as zeros) are 23-40 % of the tile and those layers are bound by TMA weight loads.
  static const int tma_store_mode = env_int("PCB_CONV_TMA_STORE", 1);
  p.tma_store = (tma_store_mode && (tma_store_mode == 2 || in.w >= 24) && !strided && !in.dense && p.stride == 1 && p.out &&
                 !p.out_s32 && !p.out_f32) ? 1 : 0;
  tmO = tmB;
  tmO2 = tmB;
  if (p.tma_store) {
    if (!make_map_2d(&tmO, p.out, (uint64_t)a.out->rows(), (uint64_t)p.out_cp, (uint64_t)p.out_cp, 32, 32))
      return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(out) failed");
    if (p.out2 && !make_map_2d(&tmO2, p.out2, (uint64_t)a.out2->rows(), (uint64_t)p.out2_cp, (uint64_t)p.out2_cp, 32, 32))
      return pcb_fail(c, PCB_ERR_CUDA, "cuTensorMapEncodeTiled(out2) failed");
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = p.pair ? 2 * (total < units ? total : units) : (total < c->num_sms ? total : c->num_sms);
  // algorithmic FLOPs of this layer: 2 * output pixels * cout * cin * taps (real, unpadded extents)
  double out_px = in.dense ? (double)in.n : (double)in.n * (in.h / p.stride) * (in.w / p.stride);
  const double k_real = (w.taps == 1 && w.cin == 3) ? 27.0 : (double)w.cin * w.taps;
  char desc[200];
  desc[0] = 0;
  if (c->profile)
    snprintf(desc, sizeof desc, "n=%d,h=%d,w=%d,cin=%d,cout=%d,taps=%d,stride=%d,ntile=%d,mt=%d,tiles=%d,grid=%d,res=%d,f32out=%d,out2=%d,ast=%d,bst=%d,bres=%d,sbox=%dx%dx%d,pair=%d,iss=%d",
             in.n, in.h, in.w, w.cin, w.cout, w.taps, p.stride, p.n_tile, p.mt, total, grid, a.residual ? 1 : 0,
             (a.out && a.out->f32) ? 1 : 0, a.out2 ? 1 : 0, p.a_stages, p.b_stages, p.b_resident, p.strided ? p.s_nb : 0,
             p.strided ? p.s_by : 0, p.strided ? p.s_bx : 0, p.pair, p.n_iss + p.split_n * 10);
  static const int debug = env_int("PCB_CONV_DEBUG", 0);
  static unsigned long long* dbg_dev = nullptr;    // debug builds of the stall counters assume one context
  if (debug && c->profile) {
    if (!dbg_dev) dbg_dev = (unsigned long long*)pcb_dev_alloc(c, 16 * sizeof(unsigned long long), true);
    p.dbg = dbg_dev;
  }
  {
    PcbConvTimer timer(c, 2.0 * out_px * (double)w.cout * k_real, desc);
    const int mode = p.out_f32 ? 2 : (p.out_s32 ? 1 : 0);
    if (mode != 0 && (p.residual || p.out2)) return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: fp32 outputs take no residual / second output");
    const int key = mode == 0 ? (p.act * 4 + (p.residual ? 2 : 0) + (p.out2 ? 1 : 0)) : (100 + mode * 4 + p.act);
    cudaError_t le = cudaErrorInvalidValue;
    // programmatic dependent launch (see the kernel's prologue); PCB_CONV_PDL=0 launches without the attribute.  Always off while
    // the per-launch profile or the stall counters are on (an event between two launches serialises them anyway).
    static const int pdl_mode = env_int("PCB_CONV_PDL", 1);
    const bool pdl = pdl_mode != 0 && !c->profile && !p.dbg;
#define PCB_TC2_LAUNCH(ACT, RES, OUT2, MODE, VAR)                                                                       \
  {                                                                                                                    \
    static unsigned long long attr_devs = 0; /* the opt-in is per device */                                            \
    if (pcb_attr_needed(&attr_devs, c->device))                                                                        \
      cudaFuncSetAttribute(conv_tc2_kernel<ACT, RES, OUT2, MODE, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    cudaLaunchConfig_t cfg = {};                                                                                       \
    cfg.gridDim = dim3(grid);                                                                                          \
    cfg.blockDim = dim3(kThreads);                                                                                     \
    cfg.dynamicSmemBytes = smem;                                                                                       \
    cfg.stream = c->stream;                                                                                            \
    cudaLaunchAttribute at[2];                                                                                         \
    int n_at = 0;                                                                                                      \
    if (VAR == kVarPair) {                                                                                             \
      at[n_at].id = cudaLaunchAttributeClusterDimension;                                                               \
      at[n_at].val.clusterDim.x = 2;                                                                                   \
      at[n_at].val.clusterDim.y = 1;                                                                                   \
      at[n_at].val.clusterDim.z = 1;                                                                                   \
      ++n_at;                                                                                                          \
    }                                                                                                                  \
    if (pdl) {                                                                                                         \
      at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                \
      at[n_at].val.programmaticStreamSerializationAllowed = 1;                                                         \
      ++n_at;                                                                                                          \
    }                                                                                                                  \
    cfg.attrs = at;                                                                                                    \
    cfg.numAttrs = n_at;                                                                                               \
    le = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<ACT, RES, OUT2, MODE, VAR>, tmA, tmA2, tmB, tmO, tmO2, p);            \
  }
#define PCB_TC2_CASE(KEY, ACT, RES, OUT2, MODE)                                                                         \
  case KEY:                                                                                                            \
    if (p.pair) PCB_TC2_LAUNCH(ACT, RES, OUT2, MODE, kVarPair)                                                         \
    else if (p.strided) PCB_TC2_LAUNCH(ACT, RES, OUT2, MODE, kVarStrided)                                              \
    else PCB_TC2_LAUNCH(ACT, RES, OUT2, MODE, kVarHalo)                                                                \
    break;
    switch (key) {
      PCB_TC2_CASE(0, 0, false, false, 0)
      PCB_TC2_CASE(1, 0, false, true, 0)
      PCB_TC2_CASE(2, 0, true, false, 0)
      PCB_TC2_CASE(3, 0, true, true, 0)
      PCB_TC2_CASE(4, 1, false, false, 0)
      PCB_TC2_CASE(5, 1, false, true, 0)
      PCB_TC2_CASE(6, 1, true, false, 0)
      PCB_TC2_CASE(7, 1, true, true, 0)
      PCB_TC2_CASE(8, 2, false, false, 0)
      PCB_TC2_CASE(9, 2, false, true, 0)
      PCB_TC2_CASE(10, 2, true, false, 0)
      PCB_TC2_CASE(11, 2, true, true, 0)
      PCB_TC2_CASE(104, 0, false, false, 1)
      PCB_TC2_CASE(105, 1, false, false, 1)
      PCB_TC2_CASE(106, 2, false, false, 1)
      PCB_TC2_CASE(108, 0, false, false, 2)
      default: break;
    }
#undef PCB_TC2_CASE
#undef PCB_TC2_LAUNCH
    if (le != cudaSuccess) return pcb_fail(c, PCB_ERR_ARG, "conv_tc2: unsupported epilogue combination");
  }
  PCB_LAUNCH_CHECK(c, "conv_tc2_kernel");
  if (p.dbg) {
    // debug only: serialises the stream.  Columns: SM MHz seen by block 0, its cycles, and the cycles each role
    // of block 0 spent blocked (A/B producer on empty slots, MMA on A/B full and on TMEM empty, epilogue on TMEM full)
    unsigned long long h[16];
    cudaMemcpyAsync(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    const double cyc = (double)(h[2] - h[0]), ns = (double)(h[3] - h[1]);
    fprintf(stderr, "CONVDBG %s | mhz=%.0f cyc=%.0f prodA_wait=%.2f prodB_wait=%.2f mma_waitA=%.2f mma_waitB=%.2f mma_waitT=%.2f epi_wait=%.2f epi_busy=%.2f epi_ldtm=%.2f\n",
            desc, ns > 0 ? cyc / ns * 1e3 : 0.0, cyc, h[4] / cyc, h[5] / cyc, h[6] / cyc, h[7] / cyc, h[8] / cyc, h[9] / cyc, h[10] / cyc, h[11] / cyc);
  }
  return PCB_OK;
}
