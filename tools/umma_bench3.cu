// Micro-benchmark #3: the MMA schedule of the column-sweep kernel (resnet_sweep.cuh) in isolation.
// One step = 9 MMAs (3 channel chunks x 3 height taps) of M=128, N=144, K=16 on a sliding window of three
// 48-column accumulator blocks in a ring of 10.  The issue loop is the kernel's: whole warp walks the
// schedule, operands live in uniform registers, one elected lane issues.  Feature bits:
//   1  the window slides by one block per step (otherwise every step hits blocks 0..2)
//   2  first MMA of a step split into N=96 (accumulate) + N=48 (overwrite) for the fresh block
//   4  ring wrap: windows that would cross block 10 are issued as two pieces
//   8  commits: one per step on a "stage" barrier + one on an "accumulator" barrier
//  16  other warps stream 16-byte shared-memory stores (TMA fill emulation, ~13 KB per step)
//  32  other warps stream tcgen05.ld of a drained block (epilogue emulation)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench3 tools/umma_bench3.cu
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

struct Cfg { int feat; int steps; int box_rows; int d; int lbo_override; int kc_stride16; };

constexpr int CP = 48, NB = 10;

__global__ void __launch_bounds__(512, 1) bench3_kernel(const Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[32];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int s_done;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    s_done = 0;
    for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bars[i]), 1);
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
  const int n_iss = (c.feat & (64 | 256)) ? 3 : 1;
  const bool by_step = (c.feat & 256) != 0;   // issuer m issues ALL MMAs of steps st % 3 == m
  if (warp >= 1 && warp <= n_iss) {
    const int me = warp - 1;
    const bool leader = (c.feat & 128) ? (lane == 0) : elect_one();
    const uint32_t sbase = smem_u32(smem);
    constexpr uint32_t hi = (128u >> 4) | (1u << 14);
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    constexpr uint32_t idesc_blk = (uint32_t)(CP >> 3) << 17;
    constexpr uint32_t W_LBO = 3 * CP * 16, W_SLAB = 2 * W_LBO, blk16 = CP;
    const uint32_t plane16 = (uint32_t)c.box_rows;
    const uint32_t a_lbo = ((c.lbo_override ? (uint32_t)c.lbo_override >> 4 : plane16) & 0x3FFFu) << 16;
    const uint32_t kcs = c.kc_stride16 ? (uint32_t)c.kc_stride16 : 2 * plane16;
    constexpr uint32_t b_lbo = ((W_LBO >> 4) & 0x3FFFu) << 16;
    // feat 1024: a zero the compiler cannot see through keeps the operands in ordinary registers (R2UR issue path)
    const uint32_t nz = (c.feat & 1024) ? *reinterpret_cast<volatile uint32_t*>(&bars[31]) >> 31 : 0u;
    const uint32_t w16 = ((sbase + 1024) >> 4) + nz;            // weights: 41.5 KB from +1 KB
    const uint32_t ring16 = ((sbase + 48 * 1024) >> 4) + nz;    // 8 stages of 6 planes x box_rows x 16 B from +48 KB
    const uint32_t slot16 = (6 * plane16 * 16 + 127) / 128 * 8;
    int stage = 0, pos = 0;
    const long long t0 = clock64();
    for (int st = 0; st < c.steps; ++st) {
      const int n = 3;
      const int p0 = (c.feat & 1) ? pos : 0;
      const int wrap_at = ((c.feat & 4) && NB - p0 < n) ? NB - p0 : n;
      const int fresh_from = (c.feat & 2) ? n - 1 : n;
      const uint32_t a16 = ring16 + (uint32_t)stage * slot16;
      const uint32_t d1 = tmem + (uint32_t)(((c.feat & 4) || p0 + n <= NB ? p0 : 0) * CP), id1 = idesc0 + (uint32_t)wrap_at * idesc_blk;
      const bool two = wrap_at < n;
      const uint32_t d2 = tmem, bo2 = (uint32_t)wrap_at * blk16, id2 = idesc0 + (uint32_t)(n - wrap_at) * idesc_blk;
      if (c.feat & 512) {
        // all operands of the step computed and pinned in registers BEFORE the burst: no register that an
        // in-flight MMA still reads is rewritten between two MMAs
        uint32_t al[9], bl[9], bl2[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int kc = k / 3, dh = k % 3;
          al[k] = ((a16 + (uint32_t)kc * kcs + (uint32_t)(dh * c.d)) & 0x3FFFu) | a_lbo;
          const uint32_t b16 = w16 + (uint32_t)((k * W_SLAB) >> 4);
          bl[k] = (b16 & 0x3FFFu) | b_lbo;
          bl2[k] = ((b16 + bo2) & 0x3FFFu) | b_lbo;
          asm volatile("" : "+r"(al[k]), "+r"(bl[k]), "+r"(bl2[k]));
        }
        if (leader && (!by_step || st % 3 == me)) {
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            if (n_iss == 3 && !by_step && k % 3 != me) continue;
            umma_lohi(d1, al[k], bl[k], hi, id1, 1u);
            if (two) umma_lohi(d2, al[k], bl2[k], hi, id2, 1u);
          }
          if (c.feat & 8) {
            umma_commit(smem_u32(&bars[stage]));
            umma_commit(smem_u32(&bars[8 + pos]));
          }
        }
      } else if (leader && (!by_step || st % 3 == me)) {
#pragma unroll
        for (int kc = 0; kc < 3; ++kc) {
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            if (n_iss == 3 && !by_step && dh != me) continue;
            const uint32_t a_lo = ((a16 + (uint32_t)kc * kcs + (uint32_t)(dh * c.d)) & 0x3FFFu) | a_lbo;
            const uint32_t b16 = w16 + (uint32_t)(((kc * 3 + dh) * W_SLAB) >> 4);
            if (kc == 0 && dh == 0 && n_iss == 1 && !by_step) {
              int s0 = 0;
              while (s0 < n) {
                int e = n;
                if (wrap_at > s0 && wrap_at < e) e = wrap_at;
                if (fresh_from > s0 && fresh_from < e) e = fresh_from;
                int q = ((c.feat & 4) || p0 + n <= NB ? p0 : 0) + s0;
                if (q >= NB) q -= NB;
                umma_lohi(tmem + (uint32_t)(q * CP), a_lo, ((b16 + (uint32_t)s0 * blk16) & 0x3FFFu) | b_lbo, hi,
                          idesc0 + (uint32_t)(e - s0) * idesc_blk, (s0 >= fresh_from || st == 0) ? 0u : 1u);
                s0 = e;
              }
            } else if (c.feat & 2048) {
              umma_lohi(d1, a_lo, (b16 & 0x3FFFu) | b_lbo, hi, id1, 1u);   // no predicated-off twin
            } else {
              umma_lohi(d1, a_lo, (b16 & 0x3FFFu) | b_lbo, hi, id1, 1u);
              if (two) umma_lohi(d2, a_lo, ((b16 + bo2) & 0x3FFFu) | b_lbo, hi, id2, 1u);
            }
          }
        }
        if (c.feat & 8) {
          umma_commit(smem_u32(&bars[stage]));
          umma_commit(smem_u32(&bars[8 + pos]));   // (barrier counts are 1: phases just flip more often with 3 issuers)
        }
      }
      __syncwarp();
      if (++stage == (c.lbo_override ? 2 : 8)) stage = 0;
      if (++pos == NB) pos = 0;
    }
    if (leader) umma_commit(smem_u32(&bars[28 + me]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[28 + me]), 0);
    const long long t1 = clock64();
    if (lane == 0 && me == 0) { cycles[blockIdx.x] = t1 - t0; s_done = 1; }
  } else if (warp >= 2 && warp < 4 && (c.feat & 16)) {
    // TMA fill emulation: ~13 KB of 16-byte stores per ~650 cycles, into a spare region (+176 KB)
    uint4* dst = reinterpret_cast<uint4*>(smem + 176 * 1024);
    int k = 0;
    while (!s_done) {
      for (int u = 0; u < 13; ++u) dst[((k + u) * 32 + lane) % 1280] = make_uint4(k, k, k, k);   // 13 x 512 B per warp
      k += 13;
      const long long t = clock64();
      while (clock64() - t < 600) {}
    }
  } else if (warp >= 4 && (c.feat & 32)) {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 480 + 16 * ((warp - 4) % 2);
    uint32_t acc = 0;
    while (!s_done) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] ^ v[15];
      const long long t = clock64();
      while (clock64() - t < 500) {}
    }
    if (acc == 0x12345678u) cycles[200] = 1;
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
  cudaMalloc(&d_cycles, 256 * sizeof(long long));
  cudaFuncSetAttribute(bench3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ST = 2000;
  printf("%5s %8s %3s | %10s %10s  (ideal 648 cycles per step = 9 x 72)\n", "feat", "box_rows", "d", "cyc/step", "cyc/MMA~");
  struct V { int feat, br, d, lbo, kcs; };
  const V vs[] = {{2048, 136, 1, 0, 0}, {2048 + 1, 136, 1, 0, 0}, {2048 + 1 + 8, 136, 1, 0, 0}, {2048 + 128, 136, 1, 0, 0}, {2048 + 256 + 1 + 8, 136, 1, 0, 0}, {2048 + 64 + 1 + 8, 136, 1, 0, 0},
                  {1024 + 128, 136, 1, 0, 0}, {1024 + 128 + 64, 136, 1, 0, 0}, {1024 + 128 + 64 + 1 + 4, 136, 1, 0, 0}, {1024 + 128 + 64 + 1 + 4 + 8, 136, 1, 0, 0},
                  {1024 + 128 + 256, 136, 1, 0, 0}, {1024 + 128 + 256 + 1 + 4 + 8, 136, 1, 0, 0}, {128 + 64, 136, 1, 0, 0}, {64, 136, 1, 0, 0},
                  {0, 136, 1, 0, 0}, {512, 136, 1, 0, 0}, {512 + 1, 136, 1, 0, 0}, {512 + 1 + 4, 136, 1, 0, 0}, {512 + 1 + 4 + 8, 136, 1, 0, 0},
                  {512 + 64 + 1 + 4 + 8, 136, 1, 0, 0}, {512 + 256 + 1 + 4 + 8, 136, 1, 0, 0},
                  {512 + 1 + 4 + 8 + 16 + 32, 136, 1, 0, 0}, {512 + 256 + 1 + 4 + 8 + 16 + 32, 160, 16, 0, 0}};
  for (const V& v : vs) {
      const int feat = v.feat, br = v.br;
      Cfg c{feat, ST, br, v.d, v.lbo, v.kcs};
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d_cycles, 0, 256 * sizeof(long long));
        bench3_kernel<<<148, 512, 200 * 1024>>>(c, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("feat %d failed: %s\n", feat, cudaGetErrorString(e)); return 1; }
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d_cycles, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int b = 0; b < 148; ++b) mx = mx > h[b] ? mx : (double)h[b];
      printf("%5d %8d %3d | %10.1f %10.1f   lbo %d kc_stride16 %d\n", feat, br, c.d, mx / ST, mx / ST / 9, v.lbo, v.kcs);
    }
  return 0;
}
