// Micro-benchmark: issue rate of tcgen05.mma.kind::f16 (SS mode, K = 16 per instruction) on B200 as a function of N, of
// cta_group (1: M = 128 per CTA; 2: M = 256 per CTA pair), of the A start row (aligned to the 8-row swizzle group or not) and of
// how many distinct B stages are cycled.  Operands are static shared-memory tiles (no TMA in the loop), so the figure is the
// tensor pipe's own pace: cycles per instruction and the implied FLOP/clk/SM.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t lo = (uint64_t)((saddr >> 4) & 0x3fff);
  uint64_t hi = (uint64_t)(1024 >> 4) | (1ull << 14) | (2ull << 29);
  return lo | (hi << 32);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <bool kPair>
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  if (kPair)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}

struct Args {
  int n, iters, a_shift_rows, b_slots, mt;   // mt: accumulators cycled (sub-tiles sharing a B slot)
  int issuers;                                // 1 | 2 issuing warps (accumulators split between them)
  int wp;                                     // > 0: A start rows follow the nine tap offsets (wp + 1 + dy*wp + dx) of a halo tile with this row pitch
  int order;                                  // 0: four K steps per accumulator, then the next accumulator; 1: accumulators innermost
  unsigned long long* out;                    // [grid] cycles
};

template <bool kPair>
__global__ void __launch_bounds__(128, 1) rate_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  // fill: small fp16 values
  __half* h = (__half*)smem;
  for (int i = threadIdx.x; i < 96 * 1024; i += blockDim.x) h[i] = __float2half((float)((i * 37) % 17 - 8) * 0.0625f);
  const int warp = threadIdx.x >> 5;
  const int rank = kPair ? (blockIdx.x & 1) : 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(a.issuers));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kPair) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp < a.issuers && rank == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.n >> 3) << 17) | ((uint32_t)((kPair ? 256 : 128) >> 4) << 24);
    // A: 64 KB region from offset 0 (start shifted by a_shift_rows * 128 B); B slots of 32 KB from offset 96 KB
    const uint32_t sa = smem_u32(smem) + (uint32_t)a.a_shift_rows * 128u;
    const uint32_t sb = smem_u32(smem) + 96u * 1024u;
    const long long t0 = clock64();
    for (int it = 0; it < a.iters; ++it) {
      const uint64_t db = make_desc(sb + (uint32_t)(it % a.b_slots) * 32768u);
      if (!elect_one()) continue;
      if (a.wp > 0) {
        // the product kernel's inner loop: tap t of a halo tile, issuer w owns sub-tile w (128 rows further down), resident weights
        const int t = it % 9;
        const uint32_t off = (uint32_t)(a.wp + 1 + (t / 3 - 1) * a.wp + (t % 3 - 1)) * 128u;
        const uint64_t da = make_desc(smem_u32(smem) + off + (uint32_t)uwarp * 16384u);
        const uint64_t db = make_desc(sb + (uint32_t)t * 8192u);
        const uint32_t d = tmem + (uint32_t)(uwarp * a.n);
        umma<kPair>(d, da, db, idesc, it ? 1u : 0u);
        umma<kPair>(d, da + 2, db + 2, idesc, 1u);
        umma<kPair>(d, da + 4, db + 4, idesc, 1u);
        umma<kPair>(d, da + 6, db + 6, idesc, 1u);
        continue;
      }
      if (a.order == 2) {
        // straight line: descriptors precomputed, 16 instructions per iteration on 4 accumulators x 4 K steps
        const uint64_t da0 = make_desc(sa), db0 = make_desc(sb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = tmem + (uint32_t)((j % a.mt) * a.n);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<kPair>(d, da0 + (uint64_t)(j * 1024 + 2 * k), db0 + 2 * k, idesc, 1u);
        }
      } else if (a.order == 0) {
        for (int j = uwarp; j < a.mt; j += a.issuers) {
          const uint64_t da = make_desc(sa + (uint32_t)j * 16384u + (uint32_t)(it % 9) * 128u * (a.a_shift_rows ? 1u : 0u));
          const uint32_t d = tmem + (uint32_t)(j * a.n);
          umma<kPair>(d, da, db, idesc, it ? 1u : 0u);
          umma<kPair>(d, da + 2, db + 2, idesc, 1u);
          umma<kPair>(d, da + 4, db + 4, idesc, 1u);
          umma<kPair>(d, da + 6, db + 6, idesc, 1u);
        }
      } else {
        for (int k = 0; k < 4; ++k) {
          for (int j = 0; j < a.mt; ++j) {
            const uint64_t da = make_desc(sa + (uint32_t)j * 16384u + (uint32_t)(it % 9) * 128u * (a.a_shift_rows ? 1u : 0u));
            const uint32_t d = tmem + (uint32_t)(j * a.n);
            umma<kPair>(d, da + 2 * k, db + 2 * k, idesc, (it | k) ? 1u : 0u);
          }
        }
      }
    }
    if (elect_one()) {
    if (kPair)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    if (uwarp == 0) a.out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kPair) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(unsigned long long));
  const size_t smem = 200 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("%5s %5s %4s %6s %6s %6s | %10s %12s %10s\n", "pair", "N", "mt", "order", "issuers", "grid", "cyc/mma", "FLOP/clk/SM", "nominal");
  printf("halo-tap sweep: N, wp -> cycles per MMA per issuer (2 issuers, one sub-tile each)\n");
  for (int n : {64, 128}) {
    for (int wp : {113, 114, 112, 120, 57, 58, 56, 64, 29, 30, 32, 15, 16, 129, 130, 128}) {
      Args a{n, 3600, 0, 1, 2, 2, wp, 0, d_out};
      cudaMemset(d_out, 0, 148 * sizeof(unsigned long long));
      for (int rep = 0; rep < 2; ++rep) rate_kernel<false><<<148, 128, smem>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148];
      cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
      double cyc = 0; int cnt = 0;
      for (int i = 0; i < 148; ++i) if (h[i]) { cyc += (double)h[i]; ++cnt; }
      printf("  N=%3d wp=%3d (wp mod 8 = %d): %.1f\n", n, wp, wp % 8, cyc / cnt / (a.iters * 4.0));
    }
  }
  for (int grid : {148}) {
    for (int pair = 0; pair < 1; ++pair) {
      for (int n : {32, 64, 96, 128, 224, 256}) {
        for (int mt : {1, 2, 4}) {
          if (mt * n > 512) continue;
          for (int cfg = 0; cfg < 4; ++cfg) {
            const int order = cfg == 3 ? 2 : 0, issuers = cfg == 1 ? 2 : 1;
            if (cfg == 1 && mt == 1) continue;
            if (cfg == 2) continue;
            {
              const int shift = 17, bslots = 3;
              Args a{n, 4000, shift, bslots, mt, issuers, 0, order, d_out};
              cudaMemset(d_out, 0, 148 * sizeof(unsigned long long));
              for (int rep = 0; rep < 2; ++rep) {
                if (pair) {
                  cudaLaunchConfig_t cfg = {};
                  cfg.gridDim = dim3(grid);
                  cfg.blockDim = dim3(128);
                  cfg.dynamicSmemBytes = smem;
                  cudaLaunchAttribute at[1];
                  at[0].id = cudaLaunchAttributeClusterDimension;
                  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                  cfg.attrs = at; cfg.numAttrs = 1;
                  cudaLaunchKernelEx(&cfg, rate_kernel<true>, a);
                } else {
                  rate_kernel<false><<<grid, 128, smem>>>(a);
                }
              }
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
              unsigned long long h[148];
              cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
              double cyc = 0; int cnt = 0;
              for (int i = 0; i < grid; ++i) if (h[i]) { cyc += (double)h[i]; ++cnt; }
              cyc /= cnt;
              const double per = cyc / (a.iters * 4.0 * (order == 2 ? 4 : mt));
              const double flop_per_sm = 2.0 * 128 * n * 16;   // per SM per instruction (pair: each SM does 128 x N)
              printf("%5d %5d %4d %6d %6d %6d | %10.1f %12.0f %10.1f\n", pair, n, mt, order, issuers, grid, per, flop_per_sm / per, n / 2.0);
            }
          }
        }
      }
    }
  }
  return 0;
}
