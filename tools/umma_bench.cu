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
};

__global__ void __launch_bounds__(160, 1) umma_bench_kernel(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
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
  long long t0 = 0, t1 = 0;
  if (warp >= 1 && warp <= c.issuers && lane == 0) {
    const int me = warp - 1;
    const uint32_t a_base = smem_u32(smem) + 1024;
    const uint32_t b_base = smem_u32(smem) + 160 * 1024;
    const int cols_per_issuer = 512 / c.issuers;
    t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint32_t a_addr = a_base + (uint32_t)((i % 32) * c.a_stride) + me * 2048 + c.a_off;
      uint64_t ad, bd;
      if (c.layout == 0) { ad = desc(a_addr, c.lbo, 128, 0); bd = desc(b_base, c.N * 16, 128, 0); }
      else { ad = desc(a_addr, 16, 1024, 2); bd = desc(b_base, 16, 1024, 2); }
      const uint32_t d = tmem + me * cols_per_issuer + (i % c.n_acc) * c.N;
      umma(d, ad, bd, idesc, i >= c.n_acc ? 1u : 0u);
    }
    umma_commit(smem_u32(&bars[me]));
    mbar_wait(smem_u32(&bars[me]), 0);
    t1 = clock64();
    cycles[blockIdx.x * 4 + me] = t1 - t0;
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
  for (int N : {48, 96, 144, 192, 240}) cfgs.push_back({N, 0, 11904, 512 / N > 5 ? 5 : 512 / N, 1, 2048, IT});
  for (int N : {48, 96, 240}) cfgs.push_back({N, 2, 0, 512 / N > 5 ? 5 : 512 / N, 1, 4096, IT});   // 128B swizzle
  cfgs.push_back({48, 0, 128, 5, 1, 2048, IT});        // K halves adjacent-ish (different core-matrix order)
  cfgs.push_back({48, 0, 11904, 1, 1, 2048, IT});      // one accumulator: dependent chain
  cfgs.push_back({48, 0, 11904, 5, 1, 0, IT});         // same A tile every time
  cfgs.push_back({48, 0, 11904, 2, 2, 2048, IT});      // 2 issuers
  cfgs.push_back({48, 0, 11904, 2, 4, 2048, IT});      // 4 issuers
  cfgs.push_back({48, 0, 11904, 1, 4, 2048, IT});      // 4 issuers, 1 accumulator each
  cfgs.push_back({96, 0, 11904, 1, 4, 2048, IT});
  cfgs.push_back({240, 0, 11904, 1, 2, 2048, IT});
  cfgs.push_back({48, 0, 11904, 2, 4, 2048, IT, 16});     // A start misaligned by 16 B (dw = +-1 at d = 1)
  cfgs.push_back({48, 0, 11904, 2, 4, 2048, IT, 64});     // misaligned by 64 B (d = 4)
  cfgs.push_back({48, 0, 11904, 2, 4, 656, IT, 0});       // row pitch 41 positions
  cfgs.push_back({48, 0, 11904, 2, 4, 2048, IT, 128});    // aligned shift (d = 8)
  cfgs.push_back({48, 0, 11968, 2, 4, 2048, IT, 0});      // LBO not a multiple of 128
  cfgs.push_back({48, 0, 12032, 2, 4, 2048, IT, 0});      // LBO = 94 * 128
  cfgs.push_back({48, 2, 0, 2, 4, 4096, IT, 0});          // 128B swizzle, 4 issuers
  cfgs.push_back({32, 0, 11904, 2, 4, 2048, IT, 0});
  cfgs.push_back({64, 0, 11904, 2, 4, 2048, IT, 0});
  printf("%4s %6s %6s %5s %7s %8s %5s | %12s %14s %10s\n", "N", "layout", "lbo", "n_acc", "issuers", "a_stride", "a_off", "cyc/MMA/iss",
         "cyc/MMA total", "ideal N/2");
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_cycles, 0, 148 * 4 * sizeof(long long));
      umma_bench_kernel<<<148, 160, 200 * 1024>>>(c, d_cycles);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("config N=%d failed: %s\n", c.N, cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(148 * 4);
    cudaMemcpy(h.data(), d_cycles, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) for (int i = 0; i < c.issuers; ++i) mx = mx > h[b * 4 + i] ? mx : (double)h[b * 4 + i];
    printf("%4d %6d %6d %5d %7d %8d %5d | %12.1f %14.1f %10.1f\n", c.N, c.layout, c.lbo, c.n_acc, c.issuers, c.a_stride, c.a_off,
           mx / c.iters, mx / c.iters / c.issuers, c.N / 2.0);
  }
  return 0;
}
