// Micro-benchmark #4: separates OPERAND effects from ISSUE effects.  The issue loop is fixed (one thread,
// rolled, descriptors read from a small shared-memory table), only the table changes:
//   table 0  bench2-like: A tiles 4 KB apart, K halves 64 KB apart; B slabs 4608 B apart
//   table 1  sweep-like:  A = 3 channel chunks x 3 row shifts inside one staged column (plane pitch 2176 B),
//                         B = 9 consecutive weight slabs
//   table 2  sweep-like A, but the three row shifts are 0 (same tile three times)
//   table 3  sweep-like A with plane pitch 2560 B and row shift 16 (d = 16)
//   table 4  bench2-like A, sweep-like B (9 slabs)
// and the number of issuing warps (each walks the same table with its own offset, all into ONE accumulator).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench4 tools/umma_bench4.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > 2000000000ll) __trap();
}
__device__ __forceinline__ void umma_lohi(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Cfg { int table; int issuers; int iters; int N; int d_slide; int d_alt; int stage_rot; int unroll9; };   // d_alt: consecutive MMAs of a thread alternate between D and D + d_alt columns

__global__ void __launch_bounds__(192, 1) bench4_kernel(const Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  __shared__ uint2 tab[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x < 9) {
    const int k = threadIdx.x, kc = k / 3, dh = k % 3;
    uint32_t a_addr, a_lbo, b_addr;
    const uint32_t b_lbo = (uint32_t)c.N * 16;
    const uint32_t b_sweep = sbase + 1024 + (uint32_t)k * (uint32_t)c.N * 32;
    const uint32_t b_b2 = sbase + 128 * 1024 + (uint32_t)(k % 4) * (uint32_t)c.N * 32;
    switch (c.table) {
      case 0: a_addr = sbase + 1024 + (uint32_t)(k % 8) * 4096; a_lbo = 65536; b_addr = b_b2; break;
      case 1: a_addr = sbase + 48 * 1024 + (uint32_t)kc * 2 * 2176 + (uint32_t)dh * 16; a_lbo = 2176; b_addr = b_sweep; break;
      case 2: a_addr = sbase + 48 * 1024 + (uint32_t)kc * 2 * 2176; a_lbo = 2176; b_addr = b_sweep; break;
      case 3: a_addr = sbase + 48 * 1024 + (uint32_t)kc * 2 * 2560 + (uint32_t)dh * 256; a_lbo = 2560; b_addr = b_sweep; break;
      case 4: a_addr = sbase + 48 * 1024 + (uint32_t)(k % 8) * 4096; a_lbo = 65536; b_addr = b_sweep; break;
      default: a_addr = sbase + 48 * 1024 + (uint32_t)(k % 8) * 4096; a_lbo = 65536; b_addr = b_b2; break;   // 5: A region moved only
    }
    tab[k] = make_uint2(((a_addr >> 4) & 0x3FFF) | (((a_lbo >> 4) & 0x3FFF) << 16), ((b_addr >> 4) & 0x3FFF) | (((b_lbo >> 4) & 0x3FFF) << 16));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
  if (warp >= 1 && warp <= c.issuers && lane == 0) {
    const int me = warp - 1;
    const uint32_t hi = (128u >> 4) | (1u << 14);
    int k = me * 3;
    uint32_t d = tmem;
    int g = 0;
    const long long t0 = clock64();
    uint32_t a_add = 0;
    int stg = 0;
    if (c.unroll9) {
      // straight-line burst of 9 MMAs per step (table read once per step into registers)
      for (int i = 0; i < c.iters; i += 9) {
        uint2 t[9];
#pragma unroll
        for (int u = 0; u < 9; ++u) t[u] = tab[u];
#pragma unroll
        for (int u = 0; u < 9; ++u) {
          if (c.unroll9 == 2 && (u % 3) != me) continue;    // by-dh split
          umma_lohi(d, t[u].x + a_add, t[u].y, hi, idesc, 1u);
        }
        if (c.stage_rot) { if (++stg == c.stage_rot) { stg = 0; a_add = 0; } else a_add += 816; }
      }
    } else {
      for (int i = 0; i < c.iters; ++i) {
        const uint2 t = tab[k];
        umma_lohi(d + ((i & 1) ? (uint32_t)c.d_alt : 0u), t.x + a_add, t.y, hi, idesc, 1u);
        if (++k == 9) {
          k = 0;
          if (c.d_slide) { d += 48; if (++g == 7) { g = 0; d = tmem; } }
          if (c.stage_rot) { if (++stg == c.stage_rot) { stg = 0; a_add = 0; } else a_add += 816; }   // 13056 B per stage
        }
      }
    }
    umma_commit(smem_u32(&bars[me]));
    mbar_wait(smem_u32(&bars[me]), 0);
    cycles[blockIdx.x * 4 + me] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, 148 * 4 * sizeof(long long));
  cudaFuncSetAttribute(bench4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int IT = 9 * 400;
  printf("%5s %7s %4s %5s | %12s %12s\n", "table", "issuers", "N", "slide", "cyc/MMA/iss", "cyc/MMA");
  for (int N : {144})
    for (int table : {1})
      for (int iss : {1, 3})
        for (int var = 0; var < 6; ++var) {
          const int slide = 0, alt = 0;
          const int rot = (var == 1 || var == 3 || var == 5) ? 8 : 0;
          const int un = var / 2;   // 0 rolled, 1 unrolled burst (every issuer all 9), 2 unrolled by-dh split
          if (un == 2 && iss == 1) continue;
          Cfg c{table, iss, IT, N, slide, alt, rot, un};
          for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_cycles, 0, 148 * 4 * sizeof(long long));
            bench4_kernel<<<148, 192, 200 * 1024>>>(c, d_cycles);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("table %d failed: %s\n", table, cudaGetErrorString(e)); return 1; }
          }
          std::vector<long long> h(148 * 4);
          cudaMemcpy(h.data(), d_cycles, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
          double mx = 0;
          for (int b = 0; b < 148; ++b) for (int i = 0; i < iss; ++i) mx = mx > h[b * 4 + i] ? mx : (double)h[b * 4 + i];
          const double per_mma_pipe = (un == 2) ? mx / IT : mx / IT / iss;   // un==2: IT counts all MMAs of the step across issuers
          printf("%5d %7d %4d %5d | %12.1f %12.1f   stage_rot %d unroll %d\n", table, iss, N, slide, mx / IT, per_mma_pipe, rot, un);
        }
  return 0;
}
