// Micro-benchmark #2 (column-sweep design questions), M=128, K=16, kind::f16, cta_group::1:
//   (1) single-thread issue interval vs N (48 / 96 / 144) with a rolled and an 8x unrolled issue loop;
//   (2) two issuer warps accumulating into the SAME TMEM columns: timing and exactness (integer-valued
//       operands make every fp32 sum exact, so any lost update shows up as a mismatch);
//   (3) N=144 with three stacked 48-row weight blocks (LBO = 2304).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_bench2 tools/umma_bench2.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
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

struct Cfg {
  int N;        // UMMA N (multiple of 16)
  int issuers;  // 1 or 2
  int same_acc; // issuers write the same accumulator columns
  int unroll;   // 1 = rolled loop, 8 = unrolled by 8 with independent descriptor registers
  int iters;    // MMAs per issuer (multiple of 8)
  int n_a;      // A tiles rotated through (each 4 KB)
  int n_b;      // B slabs rotated through
  int a_lbo;    // bytes between the K halves of A (0 = 64 KB)
  int d_group;  // > 0: the accumulator address advances by d_step columns every d_group MMAs (sliding window), wrapping below 512 - N
  int d_step;
  int commit_every;  // > 0: an extra tcgen05.commit every this many MMAs
};

constexpr int A_OFF = 1024, B_OFF = 128 * 1024;

template <int UNROLL>
__device__ __forceinline__ void issue_loop(const Cfg& c, uint32_t tmem_d, uint32_t a_lo0, uint32_t b_lo0, uint32_t hi,
                                           uint32_t idesc, int first, int me, uint32_t bar_extra) {
  const uint32_t b_step = ((uint32_t)c.N * 32) >> 4;   // one slab = 2 K-halves x N rows x 16 B
  if (UNROLL == 1) {
    int ai = me % c.n_a, bi = me % c.n_b;
    uint32_t d = tmem_d;
    int gi = 0, ci = 0;
    for (int i = 0; i < c.iters; ++i) {
      umma_lohi(d, a_lo0 + ai * 256, b_lo0 + bi * b_step, hi, idesc, (i > 0 || !first) ? 1u : 0u);
      if (++ai == c.n_a) ai = 0;
      if (++bi == c.n_b) bi = 0;
      if (c.d_group > 0 && ++gi == c.d_group) { gi = 0; d += c.d_step; if ((d & 0xFFFF) + c.N > 512) d = tmem_d; }
      if (c.commit_every > 0 && ++ci == c.commit_every) { ci = 0; umma_commit(bar_extra); }
    }
  } else {
    // 8 descriptors live in registers; n_a, n_b must divide 8 or be >= 8 (rotation restarts every 8)
    uint32_t al[8], bl[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { al[u] = a_lo0 + ((u + me) % c.n_a) * 256; bl[u] = b_lo0 + ((u + me) % c.n_b) * b_step; }
    for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_lohi(tmem_d, al[u], bl[u], hi, idesc, (i + u > 0 || !first) ? 1u : 0u);
    }
  }
}

__global__ void __launch_bounds__(192, 1) bench2_kernel(Cfg c, long long* cycles, float* dout) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // integer-valued bf16 operands: A in {0,1,2,3}, B in {-1,0,1,2}
  __nv_bfloat16* sm16 = reinterpret_cast<__nv_bfloat16*>(smem);
  for (int i = threadIdx.x; i < 200 * 1024 / 2; i += blockDim.x) {
    const uint32_t h = (uint32_t)i * 2654435761u;
    const int v = (int)((h >> 13) & 3);
    sm16[i] = __float2bfloat16((i * 2 < B_OFF) ? (float)v : (float)(v - 1));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
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
    const uint32_t a_lo0 = (((smem_u32(smem) + A_OFF) >> 4) & 0x3FFF) | ((((uint32_t)(c.a_lbo ? c.a_lbo : 64 * 1024)) >> 4) << 16);   // default LBO 64 KB: K halves far apart
    const uint32_t b_lo0 = (((smem_u32(smem) + B_OFF) >> 4) & 0x3FFF) | ((((uint32_t)c.N * 16) >> 4) << 16);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t d = tmem + (c.same_acc ? 0 : me * 256);
    // when both issuers share the accumulator, issuer 1 must not clear it: issuer 0 clears, then a barrier
    if (c.same_acc && c.issuers == 2) {
      if (me == 0) { umma_lohi(d, a_lo0, b_lo0, hi, idesc, 0u); umma_commit(smem_u32(&bars[2])); }
      mbar_wait(smem_u32(&bars[2]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int first = (c.same_acc && c.issuers == 2) ? 0 : 1;
    const long long t0 = clock64();
    if (c.unroll == 1) issue_loop<1>(c, d, a_lo0, b_lo0, hi, idesc, first, me, smem_u32(&bars[3]));
    else issue_loop<8>(c, d, a_lo0, b_lo0, hi, idesc, first, me, smem_u32(&bars[3]));
    umma_commit(smem_u32(&bars[me]));
    mbar_wait(smem_u32(&bars[me]), 0);
    const long long t1 = clock64();
    cycles[blockIdx.x * 2 + me] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // dump accumulator columns [0, N) of issuer 0 (CTA 0 only)
  if (blockIdx.x == 0 && warp >= 2 && warp < 6) {
    const int q = warp & 3;
    for (int col = 0; col < c.N; col += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((uint32_t)(q * 32) << 16) + col) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int e = 0; e < 16; ++e) dout[(q * 32 + lane) * 256 + col + e] = __uint_as_float(v[e]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// host model of what the accumulator must hold: sum over all issued MMAs of A_tile x B_slab
static float bf16_val(size_t i) {
  const uint32_t h = (uint32_t)i * 2654435761u;
  const int v = (int)((h >> 13) & 3);
  return (i * 2 < (size_t)B_OFF) ? (float)v : (float)(v - 1);
}
static double expect(const Cfg& c, int row, int col) {
  // A element (row m, k): byte A_OFF + tile*4096 + (m/8)*128 + (m%8)*16 + (k%8)*2 + (k/8)*65536
  // B element (n, k):     byte B_OFF + slab*(N*32) + (n/8)*128 + (n%8)*16 + (k%8)*2 + (k/8)*N*16
  double s = 0;
  std::vector<long long> cnt((size_t)c.n_a * c.n_b, 0);
  auto add_seq = [&](int me, int iters) {
    for (int i = 0; i < iters; ++i) {
      int ai, bi;
      if (c.unroll == 1) { ai = (me + i) % c.n_a; bi = (me + i) % c.n_b; }
      else { ai = ((i % 8) + me) % c.n_a; bi = ((i % 8) + me) % c.n_b; }
      cnt[(size_t)ai * c.n_b + bi]++;
    }
  };
  add_seq(0, c.iters);
  if (c.same_acc && c.issuers == 2) { add_seq(1, c.iters); cnt[0]++; }
  for (int ai = 0; ai < c.n_a; ++ai)
    for (int bi = 0; bi < c.n_b; ++bi) {
      if (!cnt[(size_t)ai * c.n_b + bi]) continue;
      double dot = 0;
      for (int k = 0; k < 16; ++k) {
        const size_t ab = (size_t)A_OFF + (size_t)ai * 4096 + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2 + (size_t)(k / 8) * 65536;
        const size_t bb = (size_t)B_OFF + (size_t)bi * c.N * 32 + (col / 8) * 128 + (col % 8) * 16 + (k % 8) * 2 + (size_t)(k / 8) * c.N * 16;
        dot += (double)bf16_val(ab / 2) * (double)bf16_val(bb / 2);
      }
      s += dot * (double)cnt[(size_t)ai * c.n_b + bi];
    }
  return s;
}

int main() {
  long long* d_cycles;
  float* d_out;
  cudaMalloc(&d_cycles, 148 * 2 * sizeof(long long));
  cudaMalloc(&d_out, 128 * 256 * sizeof(float));
  cudaFuncSetAttribute(bench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<Cfg> cfgs;
  const int IT = 4000;
  //              N  iss same unroll it n_a n_b
  for (int N : {48, 96, 144, 192, 240})
    for (int un : {1, 8}) cfgs.push_back({N, 1, 0, un, IT, 8, 4, 0, 0, 0, 0});
  for (int N : {48, 96, 144})
    for (int un : {1, 8}) {
      cfgs.push_back({N, 2, 0, un, IT, 8, 4, 0, 0, 0, 0});
      cfgs.push_back({N, 2, 1, un, IT, 8, 4, 0, 0, 0, 0});
    }
  // column-sweep questions (timing only; the accumulator check does not model these)
  cfgs.clear();
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 0, 0, 0});        // reference
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 2080, 0, 0, 0});     // A LBO = 130 rows (not a multiple of 128 B)
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 2176, 0, 0, 0});     // A LBO = 136 rows
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 9, 48, 0});       // accumulator window slides by 48 columns every 9 MMAs
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 1, 48, 0});       // ... every MMA
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 9, 64, 0});       // slides by 64 columns
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 0, 0, 9});        // a commit every 9 MMAs
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 0, 0, 0, 3});        // a commit every 3 MMAs
  cfgs.push_back({144, 1, 0, 1, IT, 8, 4, 2080, 9, 48, 3});    // everything
  cfgs.push_back({48, 1, 0, 1, IT, 8, 4, 0, 1, 48, 0});        // N=48 hopping blocks
  printf("%4s %3s %4s %6s | %12s %13s %9s | %s\n", "N", "iss", "same", "unroll", "cyc/MMA/iss", "cyc/MMA total", "ideal N/2", "accumulator check");
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_cycles, 0, 148 * 2 * sizeof(long long));
      bench2_kernel<<<148, 192, 200 * 1024>>>(c, d_cycles, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("config N=%d failed: %s\n", c.N, cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> h(148 * 2);
    std::vector<float> o(128 * 256);
    cudaMemcpy(h.data(), d_cycles, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(o.data(), d_out, o.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b) for (int i = 0; i < c.issuers; ++i) mx = mx > h[b * 2 + i] ? mx : (double)h[b * 2 + i];
    int bad = 0; double worst = 0;
    for (int r = 0; r < 128; r += 7)
      for (int col = 0; col < c.N; col += 5) {
        const double e = expect(c, r, col), g = o[r * 256 + col];
        if (e != g) { ++bad; if (fabs(e - g) > worst) worst = fabs(e - g); }
      }
    printf("%4d %3d %4d %6d | %12.1f %13.1f %9.1f | %s (mismatches %d, worst %.0f) lbo %d slide %d/%d commit %d\n", c.N, c.issuers, c.same_acc, c.unroll,
           mx / c.iters, mx / c.iters / c.issuers, c.N / 2.0, bad ? "MISMATCH" : "exact", bad, worst, c.a_lbo, c.d_group, c.d_step, c.commit_every);
  }
  return 0;
}
