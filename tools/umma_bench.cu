// Micro-benchmark: steady-state cost of one tcgen05.mma (cta_group::1, kind::f16, M=128, K=16)
// as a function of N, operand layout, accumulator rotation and the number of issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_bench tools/umma_bench.cu
// The numbers decide how conv_tc.cu feeds the tensor pipe (see profiles/r1_umma_bench.md).
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

struct Cfg {
  int N;          // UMMA N
  int layout;     // 0 = SWIZZLE_NONE (planar-8 slabs, K halves `lbo` apart), 2 = SWIZZLE_128B rows
  int lbo;        // bytes between K halves for layout 0
  int n_acc;      // accumulators rotated through by each issuer
  int issuers;    // issuing warps (1..4)
  int a_stride;   // bytes added to the A start address per MMA (0 = same tile every time)
  int iters;      // MMAs per issuer
  int a_off;      // constant byte offset of the A start address (16-byte granular misalignment)
  int b_rot;      // rotate the B operand over this many 1536-byte slabs (0/1 = fixed)
  int ld_warps;   // extra warps that stream tcgen05.ld from TMEM while the MMAs run (epilogue emulation)
  int st_warps;   // extra warps that stream st.shared.v4 into a spare smem region (TMA-write emulation)
};

__global__ void __launch_bounds__(672, 1) umma_bench_kernel(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int s_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    s_done = 0;
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
  long long t0 = 0, t1 = 0;
  if (warp >= 1 && warp <= c.issuers && lane == 0) {
    const int me = warp - 1;
    const uint32_t a_base = smem_u32(smem) + 1024;
    const uint32_t b_base = smem_u32(smem) + 128 * 1024;
    const int cols_per_issuer = 512 / c.issuers;
    // tight issue loop: no divisions, descriptors advanced incrementally (the issue cost must stay below the
    // pipe cost for the measurement to show the pipe)
    const uint32_t a_lo0 = ((a_base + me * 2048 + c.a_off) >> 4) & 0x3FFF;
    const uint32_t a_hi_lbo = ((uint32_t)(c.lbo >> 4) & 0x3FFF) << 16;
    const uint32_t b_lo0 = ((b_base >> 4) & 0x3FFF) | ((((uint32_t)c.N * 16) >> 4) << 16);
    const uint32_t hi = (c.layout == 0) ? ((128u >> 4) | (1u << 14)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
    const uint32_t a_step = (uint32_t)c.a_stride >> 4;
    int ai = 0, bi = 0, di = 0;
    uint32_t a_lo = a_lo0, b_lo = b_lo0, d = tmem + me * cols_per_issuer;
    const int b_rot = c.b_rot > 1 ? c.b_rot : 1;
    t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint64_t ad = ((uint64_t)hi << 32) | (a_lo | a_hi_lbo);
      const uint64_t bd = ((uint64_t)hi << 32) | b_lo;
      umma(d, ad, bd, idesc, i >= c.n_acc ? 1u : 0u);
      a_lo += a_step; if (++ai == 32) { ai = 0; a_lo = a_lo0; }
      b_lo += 96; if (++bi == b_rot) { bi = 0; b_lo = b_lo0; }
      d += c.N; if (++di == c.n_acc) { di = 0; d = tmem + me * cols_per_issuer; }
    }
    umma_commit(smem_u32(&bars[me]));
    mbar_wait(smem_u32(&bars[me]), 0);
    t1 = clock64();
    cycles[blockIdx.x * 4 + me] = t1 - t0;
    atomicAdd((int*)&s_done, 1);
  } else if (warp >= 5 && warp < 5 + c.ld_warps) {
    // epilogue emulation: keep reading 16 TMEM columns of this warp's lane quarter
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 496;
    uint32_t acc = 0;
    while (s_done < c.issuers) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] ^ v[15];
    }
    if (acc == 0x12345678u) cycles[0] = 1;
  } else if (warp >= 17 && warp < 17 + c.st_warps) {
    // TMA-write emulation: stream 16-byte stores into a spare 24 KB region
    uint4* dst = reinterpret_cast<uint4*>(smem + 172 * 1024);
    int k = 0;
    while (s_done < c.issuers) {
      dst[(k * 32 + lane) % 1536] = make_uint4(k, k, k, k);
      ++k;
    }
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
  cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<Cfg> cfgs;
  const int IT = 4000;
  //            N  lay lbo    nacc iss a_stride it  a_off b_rot ld st
  for (int iss : {1, 2, 3, 4}) cfgs.push_back({48, 0, 11904, 2, iss, 2048, IT, 0, 0, 0, 0});
  cfgs.push_back({48, 0, 11904, 2, 3, 2048, IT, 16, 0, 0, 0});    // misaligned A
  cfgs.push_back({48, 0, 11904, 2, 3, 2048, IT, 0, 27, 0, 0});    // B rotates over 27 weight slabs
  cfgs.push_back({48, 0, 11904, 2, 3, 656, IT, 16, 27, 0, 0});    // both, row-pitch stride
  cfgs.push_back({48, 0, 11904, 2, 3, 656, IT, 16, 27, 12, 0});   // + 12 warps streaming tcgen05.ld
  cfgs.push_back({48, 0, 11904, 2, 3, 656, IT, 16, 27, 0, 2});    // + 2 warps streaming st.shared
  cfgs.push_back({48, 0, 11904, 2, 3, 656, IT, 16, 27, 12, 2});   // both
  cfgs.push_back({48, 0, 11904, 1, 3, 656, IT, 16, 27, 12, 2});   // both, one accumulator per issuer
  cfgs.push_back({48, 0, 11904, 1, 1, 656, IT, 16, 27, 0, 0});    // single issuer
  cfgs.push_back({240, 0, 11904, 1, 1, 2048, IT, 0, 0, 0, 0});    // N = 240 reference
  printf("%4s %6s %6s %5s %7s %8s %5s %5s %3s %3s | %12s %14s %10s\n", "N", "layout", "lbo", "n_acc", "issuers", "a_stride", "a_off", "b_rot", "ld", "st", "cyc/MMA/iss",
         "cyc/MMA total", "ideal N/2");
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_cycles, 0, 148 * 4 * sizeof(long long));
      umma_bench_kernel<<<148, 672, 200 * 1024>>>(c, d_cycles);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("config N=%d failed: %s\n", c.N, cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(148 * 4);
    cudaMemcpy(h.data(), d_cycles, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) for (int i = 0; i < c.issuers; ++i) mx = mx > h[b * 4 + i] ? mx : (double)h[b * 4 + i];
    printf("%4d %6d %6d %5d %7d %8d %5d %5d %3d %3d | %12.1f %14.1f %10.1f\n", c.N, c.layout, c.lbo, c.n_acc, c.issuers, c.a_stride, c.a_off, c.b_rot, c.ld_warps, c.st_warps,
           mx / c.iters, mx / c.iters / c.issuers, c.N / 2.0);
  }
  return 0;
}
