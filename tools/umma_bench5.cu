// Micro-benchmark #5 (groundwork for the next staging layout): does a K-major SWIZZLE_32B operand whose descriptor
// START ADDRESS is shifted by an arbitrary number of 32-byte rows still read the rows it should?
//
// The column-sweep kernel takes its height taps as row offsets of the A descriptor into one staged column.  Today a
// column is six 8-channel planes ([plane][row][16 B], SWIZZLE_NONE) = six bulk-copy requests per step.  With 16
// channels per row ([row][32 B], SWIZZLE_32B) a column would be three requests -- if the tensor core applies the
// 32-byte swizzle (16-byte chunk index ^= address bit 7) to ABSOLUTE shared-memory addresses, so that a start
// address in the middle of an 8-row atom keeps working.  This program answers that numerically:
//   A[r][k]  (r = 0 .. 159 rows, k = 0 .. 15) is written in the swizzled form at smem offset
//            r*32 + (((k/8) ^ ((r>>2)&1)) * 16) + (k%8)*2      (atom = 8 rows x 32 B = 256 B, buffer 1024-aligned)
//   B[n][k]  (n = 0 .. 15) in the canonical SWIZZLE_NONE K-major form the kernel already uses
//   D[i][n] = sum_k A[r0 + i][k] * B[n][k]  for row shifts r0 = 0 .. 17, one 128 x 16 x 16 MMA each,
// compared exactly (small integers) with the CPU.  Also tried: the descriptor's base-offset field = (start>>7)&7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench5 tools/umma_bench5.cu
#include <cuda_bf16.h>
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int kRows = 160, kN = 16, kShifts = 18;

__global__ void __launch_bounds__(128, 1) bench5_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, int use_base_offset,
                                                        float* D /* [kShifts][128][kN] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* sA = smem;                 // kRows * 32 B, swizzled
  unsigned char* sB = smem + 8192;          // canonical: [K half][n][16 B]
  for (int i = threadIdx.x; i < kRows * 16; i += blockDim.x) {
    const int r = i / 16, k = i % 16;
    const int off = r * 32 + (((k >> 3) ^ ((r >> 2) & 1)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[r * 16 + k];
  }
  for (int i = threadIdx.x; i < kN * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 3) * (kN * 16) + n * 16 + (k & 7) * 2) = B[n * 16 + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // instruction descriptor: D = f32, A = B = bf16, K-major both, N = 16, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | (8u << 24);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kShifts; ++s) {
      const uint32_t a_addr = smem_u32(sA) + (uint32_t)s * 32u;
      // A: SWIZZLE_32B (layout type 6 in bits 61-63), SBO = 256 B (8 rows), LBO unused (1), version 1 in bits 46-47
      uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
                    ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
      if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      // B: SWIZZLE_NONE, LBO = kN*16 B between K halves, SBO = 128 B between 8-column groups
      const uint32_t b_addr = smem_u32(sB);
      const uint64_t db = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)((kN * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
                          ((uint64_t)1 << 46);
      umma(tmem + (uint32_t)(s * kN), da, db, idesc, 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = 0; s < kShifts; ++s) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * kN)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < kN; ++n) D[((size_t)s * 128 + warp * 32 + lane) * kN + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  std::vector<__nv_bfloat16> hA(kRows * 16), hB(kN * 16);
  std::vector<float> fA(kRows * 16), fB(kN * 16);
  for (int r = 0; r < kRows; ++r)
    for (int k = 0; k < 16; ++k) { fA[r * 16 + k] = (float)(((r * 7 + k * 3) % 13) - 6); hA[r * 16 + k] = __float2bfloat16(fA[r * 16 + k]); }
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < 16; ++k) { fB[n * 16 + k] = (float)(((n * 5 + k * 11) % 9) - 4); hB[n * 16 + k] = __float2bfloat16(fB[n * 16 + k]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, (size_t)kShifts * 128 * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int ubo = 0; ubo < 2; ++ubo) {
    cudaMemset(dD, 0, (size_t)kShifts * 128 * kN * 4);
    bench5_kernel<<<1, 128, 16384>>>(dA, dB, ubo, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("base_offset=%d: kernel failed: %s\n", ubo, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD((size_t)kShifts * 128 * kN);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    printf("descriptor base-offset field %s:\n", ubo ? "= (start >> 7) & 7" : "left 0");
    for (int s = 0; s < kShifts; ++s) {
      int bad = 0, first_bad = -1;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < kN; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 16; ++k) ref += fA[(s + i) * 16 + k] * fB[n * 16 + k];
          if (hD[((size_t)s * 128 + i) * kN + n] != ref) { ++bad; if (first_bad < 0) first_bad = i; }
        }
      printf("  row shift %2d (start +%3d B): %s", s, s * 32, bad ? "WRONG" : "exact");
      if (bad) printf("  (%d of %d elements, first bad row %d)", bad, 128 * kN, first_bad);
      printf("\n");
    }
  }
  return 0;
}
