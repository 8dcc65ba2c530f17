// bf16 tensor-core ResNet path for sm_100a: tcgen05.mma (UMMA) implicit-GEMM 3x3 dilated
// convolution with TMA-staged halo tiles, TMEM accumulators and a fused
// ReLU / residual / BatchNorm epilogue (/root/reference/model/resnet.py:48-55), plus the
// conv_0 and mean+linear kernels for the same activation layout.
//
// Activation layout ("planar-8"): [B][NP][H][W][8] bf16, NP = CP/8 planes of 8 channels,
// CP = channels padded to a multiple of 16 (45 -> 48).  One plane row is W*16 contiguous bytes,
// so a TMA box {8 ch, Wp, rows} lands in shared memory as a dense array of 16-byte "positions":
// exactly the UMMA K-major SWIZZLE_NONE canonical layout (core matrix = 8 positions x 16 B,
// SBO = 128 B), in which a pixel shift is nothing but a start-address offset.
//
// Implicit GEMM per tap (dh, dw) and 16-channel chunk kc:
//   D[128 positions x CP] += A[128 positions x 16] (shifted view of the staged tile)
//                          * B[16 x CP]            (weights, resident in shared memory)
// Positions are flat indices p = r * Wp + c into the zero-padded tile rows (Wp = W + d: the d
// left-pad columns of row r+1 double as the right padding of row r; TMA out-of-bounds fill
// supplies every zero).  Outputs at pad positions are computed and discarded.
//
// Warp roles (192 threads, 1 CTA per SM, persistent over tiles):
//   warp 0     TMA producer: one stage = one 16-channel chunk of the haloed input tile
//   warp 1     MMA issuer (one elected lane) + TMEM allocation
//   warps 2-5  epilogue: tcgen05.ld -> ReLU -> +skip -> BN -> bf16 planar-8 stores
// TMEM holds two accumulator buffers (up to 5 M-tiles x CP columns each) so the epilogue of
// tile i overlaps the MMAs of tile i+1.
#include "tc.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace kws {

constexpr int kTcIssuers = 3;      // MMA-issuing warps (M-tiles dealt round-robin)
// warp 0 TMA, warps 1-3 MMA, then 4*NKC epilogue warps: warp e owns TMEM lane quarter (warp % 4) and the
// 16-channel group e / 4 (two planar-8 planes) of EVERY M-tile
__host__ __device__ constexpr int tc_epi_warps(int NKC) { return 4 * NKC; }
__host__ __device__ constexpr int tc_threads(int NKC) { return 32 * (1 + kTcIssuers + tc_epi_warps(NKC)); }
constexpr int kTcMaxMt = 8;       // upper bound of M-tiles per tile (min(8, kAccCols / CP) at run time)
constexpr int kAccCols = 256;     // TMEM columns per accumulator buffer
constexpr int kTcMaxLanes = 4;
constexpr int kMaxStages = 8;
// ---------------------------------------------------------------------------------------------
struct TcGeom {
  int H, W, d, dpad, Wp, R, tiles_per_utt;
  int Hpad;                            // rows per plane in memory (rows >= H are kept zero)
  int phase, chunks_per_phase;         // phase tiling: a tile = rows p, p+d, p+2d, ... (R of them) of phase p
  int rows_box, n_boxes, box_stride;   // bytes between row-block boxes inside a slab
  int h_start[3];                      // first input row of box bx relative to the tile's h0
  int tap_off[3];                      // byte offset inside a slab of tap row dh = -1, 0, +1
  int slab_bytes, stage_bytes, n_stages, side_taps;
  int smem_w_off, smem_ring_off, smem_total;
};

struct TcConvParams {
  const __nv_bfloat16* wpack;  // [9][NKC][2][CP][8]; input-channel axis pre-multiplied by the previous layer's BN scale
  const float* kconst;         // [CP] per-channel epilogue constant: (mean of the skip layer, even layers) - mean of this layer
  const __nv_bfloat16* skip;   // even layers: centred activation z of layer i-2 (conv_0 output for i = 2); may alias y
  __nv_bfloat16* y;            // centred activation z = x - mean (the 1/sigma factor lives in the next layer's weights)
  float* pool_sum;             // [B][CP] per-utterance sums of z over H*W (last layer), or nullptr
  int B, total_tiles;
  TcGeom g;
};

// HAS_PREV: even layer (adds and rewrites the skip tensor).  DO_POOL: last layer (accumulates the
// global mean instead of storing the activation).
template <int NKC, bool HAS_PREV, bool DO_POOL>
__global__ void __launch_bounds__(tc_threads(NKC), 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcConvParams p) {
  constexpr int CP = 16 * NKC;       // padded channels = UMMA N
  constexpr int NP = 2 * NKC;        // 8-channel planes
  constexpr int W_HALF = CP * 16;    // bytes of one [CP][8] weight half-slab
  constexpr int W_BYTES = 9 * NKC * 2 * W_HALF;
  constexpr int MAXMT = (kAccCols / CP) < kTcMaxMt ? (kAccCols / CP) : kTcMaxMt;
  constexpr int MAXU = (MAXMT + kTcIssuers - 1) / kTcIssuers;   // M-tiles per issuer warp
  extern __shared__ __align__(1024) unsigned char smem[];
  const TcGeom& g = p.g;

  // ---- shared memory carve-up
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);   // full[8], empty[8], tfull[2], tempty[2], wfull
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * 24);
  float* s_kconst = reinterpret_cast<float*>(smem + 256);         // [CP]
  unsigned char* s_w = smem + g.smem_w_off;
  unsigned char* s_ring = smem + g.smem_ring_off;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t wfull_bar = bar0 + 8u * (2 * kMaxStages + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < g.n_stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kTcIssuers); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), kTcIssuers); mbar_init(tempty_bar(a), tc_epi_warps(NKC)); }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
    // weights: one asynchronous bulk copy global -> shared, completion on wfull_bar
    mbar_expect_tx(wfull_bar, W_BYTES);
    bulk_load(smem_u32(s_w), p.wpack, W_BYTES, wfull_bar);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < CP; i += tc_threads(NKC)) s_kconst[i] = p.kconst[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, weights) overlapped the tail of the
  // previous layer's kernel; the activations it wrote are only touched after this wait.  The next
  // layer's kernel may be scheduled as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tiles_per_utt = g.tiles_per_utt;
  // Each CTA takes a contiguous range of tiles: consecutive tiles of one utterance stay on one SM (L2/TLB
  // locality, and the fused pooling flushes once per utterance instead of once per tile).
  const int t_begin = (int)(((int64_t)blockIdx.x * p.total_tiles) / gridDim.x);
  const int t_end = (int)(((int64_t)(blockIdx.x + 1) * p.total_tiles) / gridDim.x);
  // tile index inside an utterance -> (phase p, first row r0 in phase-row units, number of rows).
  // Row r of the tile is image row (r0 + r) * hstep + p, hstep = d in phase mode and 1 otherwise.
  auto tile_decode = [&](int tix, int& ph, int& r0, int& rows) {
    if (g.phase) {
      ph = tix / g.chunks_per_phase;
      r0 = (tix - ph * g.chunks_per_phase) * g.R;
      rows = min(g.R, (g.H - ph + g.d - 1) / g.d - r0);
    } else {
      ph = 0;
      r0 = tix * g.R;
      rows = min(g.R, g.H - r0);
    }
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(2 * g.n_boxes * g.rows_box * g.Wp * 16);
      for (int t = t_begin; t < t_end; ++t) {
        const int b = t / tiles_per_utt, tix = t - b * tiles_per_utt;
        int ph, r0, rows;
        tile_decode(tix, ph, r0, rows);
        if (rows <= 0) continue;
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sbase = smem_u32(s_ring + (size_t)stage * g.stage_bytes);
          for (int half = 0; half < 2; ++half)
            for (int bx = 0; bx < g.n_boxes; ++bx)
              tma_load_5d(sbase + half * g.slab_bytes + bx * g.box_stride, &tmap, full_bar(stage), 0, -g.dpad,
                          r0 + g.h_start[bx], ph, b * NP + 2 * kc + half);
          if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp <= kTcIssuers) {
    // ================================ MMA issuers (3 warps) ================================
    // A 128 x CP x 16 MMA lasts ~CP/2 tensor-pipe cycles, less than one thread needs to set up and
    // issue it, so the M-tiles of a tile are dealt round-robin to kTcIssuers warps (disjoint TMEM
    // accumulators, hence no ordering between them); every issuer commits to the same barriers.
    // All operands are warp-uniform.  A descriptor = {lo: addr>>4 | (LBO>>4)<<16, hi: SBO>>4 | version};
    // per MMA only the 14-bit address field of `lo` changes, by a precomputed 16-byte-unit offset.
    const int me = warp - 1;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    constexpr uint32_t idesc = umma_idesc(128, CP);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_fields = ((uint32_t)(g.slab_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t b_lo_base = ((smem_u32(s_w) >> 4) & 0x3FFFu) | (((uint32_t)(W_HALF >> 4) & 0x3FFFu) << 16);
    int tap16[9];   // start offset of tap (dh,dw) inside a stage, in 16-byte units (may be negative)
#pragma unroll
    for (int dh = 0; dh < 3; ++dh)
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) tap16[dh * 3 + dw] = (g.tap_off[dh] >> 4) + (dw - 1) * g.d;
    const bool side = g.side_taps != 0;
    const int first_tap = side ? 0 : 1;
    const bool leader = elect_one();
    mbar_wait(wfull_bar, 0);   // weights have landed
    for (int t = t_begin; t < t_end; ++t) {
      int ph, r0, rows;
      tile_decode(t % tiles_per_utt, ph, r0, rows);
      if (rows <= 0) continue;
      const int n_mt = (rows * g.Wp + 127) >> 7;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + acc * kAccCols + me * CP;
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (leader) {
          const uint32_t a_lo_stage =
              (((smem_u32(s_ring + (size_t)stage * g.stage_bytes) >> 4) + me * 128) & 0x3FFFu) | a_lo_fields;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if ((tap % 3) != 1 && !side) continue;
            const uint32_t a_lo = a_lo_stage + (uint32_t)tap16[tap];
            const uint32_t b_lo = b_lo_base + (uint32_t)(((tap * NKC + kc) * 2 * W_HALF) >> 4);
            if (kc == 0 && tap <= 1) {
              // the first tap of a tile overwrites the accumulators (tap 0, or tap 1 when d >= W)
              const uint32_t accum = (tap == first_tap) ? 0u : 1u;
#pragma unroll
              for (int u = 0; u < MAXU; ++u)
                if (me + u * kTcIssuers < n_mt)
                  umma_f16_lohi_rt(d_base + u * kTcIssuers * CP, a_lo + u * kTcIssuers * 128, b_lo, desc_hi, idesc, accum);
            } else {
#pragma unroll
              for (int u = 0; u < MAXU; ++u)
                if (me + u * kTcIssuers < n_mt)
                  umma_f16_lohi<true>(d_base + u * kTcIssuers * CP, a_lo + u * kTcIssuers * 128, b_lo, desc_hi, idesc);
            }
          }
          umma_commit(empty_bar(stage));                       // stage reusable once these MMAs retire
          if (kc == NKC - 1) umma_commit(tfull_bar(acc));      // accumulators complete
        }
        __syncwarp();
        if (++stage == g.n_stages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================================ epilogue (4*NKC warps) ================================
    // Warp e reads TMEM lane quarter q = warp % 4 (fixed by hardware) and the 16 accumulator columns of
    // channel group j = e / 4, for every M-tile of the tile: z = ReLU(acc) (+ skip) + kconst, two 16-byte
    // planar-8 stores per position.  The 16 per-channel constants live in registers.
    const int q = warp & 3;
    const int j = (warp - (1 + kTcIssuers)) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int64_t plane_stride = (int64_t)g.Hpad * g.W;   // in 16-byte (8-channel) units
    const int hstep = g.phase ? g.d : 1;
    const uint4* skip_in = reinterpret_cast<const uint4*>(p.skip) + (int64_t)(2 * j) * plane_stride;
    uint4* y_out = reinterpret_cast<uint4*>(p.y) + (int64_t)(2 * j) * plane_stride;
    float kc_reg[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) kc_reg[c] = s_kconst[16 * j + c];
    float psum[DO_POOL ? 16 : 1];   // this thread's share of sum_{h,w} z of the current utterance (resnet.py:57-58)
    int pool_b = -1;
    auto pool_flush = [&]() {
      if constexpr (DO_POOL) {
        if (pool_b >= 0) {   // warp-reduce the 32 positions, one atomic per channel
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float sum = psum[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == c) atomicAdd(p.pool_sum + (int64_t)pool_b * CP + 16 * j + c, sum);
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) psum[c] = 0.f;
      }
    };
    pool_flush();
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / tiles_per_utt, tix = t - b * tiles_per_utt;
      if (DO_POOL && b != pool_b) { pool_flush(); pool_b = b; }
      int ph, r0, rows;
      tile_decode(tix, ph, r0, rows);
      if (rows <= 0) continue;
      const int n_mt = (rows * g.Wp + 127) >> 7;
      const int64_t utt_base = ((int64_t)b * NP) * plane_stride + (int64_t)(r0 * hstep + ph) * g.W;
      // position of this thread in M-tile `mt`: valid flag and offset (16-byte units) inside a plane
      auto locate = [&](int mt, bool& valid) -> int64_t {
        const int pos = mt * 128 + q * 32 + lane;
        const int r = pos / g.Wp;
        const int w = pos - r * g.Wp - g.dpad;
        valid = (w >= 0) && (r < rows) && (mt < n_mt);
        return utt_base + (int64_t)(r * hstep) * g.W + w;
      };
      // The skip tensor is fetched one M-tile ahead; the first fetch is issued BEFORE waiting for
      // the accumulators, so its latency hides behind the MMAs of this tile.
      uint4 pv_next[2];
      if constexpr (HAS_PREV) {
        bool v0;
        const int64_t b0 = locate(0, v0);
        if (v0) { pv_next[0] = skip_in[b0]; pv_next[1] = skip_in[b0 + plane_stride]; }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int mt = 0; mt < n_mt; ++mt) {
        bool valid;
        const int64_t base = locate(mt, valid);
        uint4 pv[2];
        if constexpr (HAS_PREV) {
          pv[0] = pv_next[0]; pv[1] = pv_next[1];
          bool v2;
          const int64_t b2 = locate(mt + 1, v2);
          if (v2) { pv_next[0] = skip_in[b2]; pv_next[1] = skip_in[b2 + plane_stride]; }
        }
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols + mt * CP + 16 * j, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = fmaxf(__uint_as_float(v[8 * hf + e]), 0.f) + kc_reg[8 * hf + e];
            if constexpr (HAS_PREV) {
              const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&pv[hf]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(pb[e]);
                x[2 * e] += f.x;
                x[2 * e + 1] += f.y;
              }
            }
            if constexpr (DO_POOL) {
#pragma unroll
              for (int e = 0; e < 8; ++e) psum[8 * hf + e] += x[e];
            } else {
              uint4 yo;
              __nv_bfloat162* yb = reinterpret_cast<__nv_bfloat162*>(&yo);
#pragma unroll
              for (int e = 0; e < 4; ++e) yb[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
              y_out[base + hf * plane_stride] = yo;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    pool_flush();
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace kws
#include "resnet_fused.cuh"
#include "resnet_sweep.cuh"
namespace kws {

// ---------------------------------------------------------------------------------------------
// conv_0 (1 -> C, 3x3, pad 1) + ReLU + AvgPool -> planar-8 bf16 (resnet.py:40-44).
// One thread per output pixel; channels in groups of 8 -> one 16-byte store per plane.
__global__ void __launch_bounds__(256)
conv0_p8_kernel(const float* __restrict__ feat, const float* __restrict__ w0, __nv_bfloat16* __restrict__ out,
                float* __restrict__ pool_sum, int T, int F, int C, int NP, int ph, int pw, int Ho, int Wo,
                int rows_per_tile, int Hpad) {
  if (pool_sum != nullptr && blockIdx.x == 0 && threadIdx.x < NP * 8)
    pool_sum[(int64_t)blockIdx.y * NP * 8 + threadIdx.x] = 0.f;
  extern __shared__ __align__(16) float smem_f[];
  const int in_rows = rows_per_tile * ph + 2, in_cols = F + 2;
  float* s_in = smem_f;
  float* s_w = smem_f + round_up(in_rows * in_cols, 4);   // [NP*8][12], zero for pad channels
  const int64_t b = blockIdx.y;
  const int ho0 = blockIdx.x * rows_per_tile;
  const float* src = feat + b * (int64_t)T * F;
  for (int i = threadIdx.x; i < in_rows * in_cols; i += blockDim.x) {
    const int r = i / in_cols, c = i - r * in_cols;
    const int h = ho0 * ph - 1 + r, w = c - 1;
    s_in[i] = (h >= 0 && h < T && w >= 0 && w < F) ? __ldg(src + (int64_t)h * F + w) : 0.f;
  }
  for (int i = threadIdx.x; i < NP * 8 * 12; i += blockDim.x) {
    const int c = i / 12, k = i - c * 12;
    s_w[i] = (k < 9 && c < C) ? __ldg(w0 + c * 9 + k) : 0.f;
  }
  __syncthreads();
  const int r = threadIdx.x / Wo, wo = threadIdx.x - r * Wo;
  const int ho = ho0 + r;
  if (r >= rows_per_tile || ho >= Ho) return;
  const float inv = 1.f / (float)(ph * pw);
  const int64_t plane_stride = (int64_t)Hpad * Wo;
  uint4* dst = reinterpret_cast<uint4*>(out) + (b * NP) * plane_stride + (int64_t)ho * Wo + wo;
  for (int pl = 0; pl < NP; ++pl) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j) {
        const float* q = s_in + (r * ph + i) * in_cols + wo * pw + j;
        float xin[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int e = 0; e < 3; ++e) xin[a * 3 + e] = q[a * in_cols + e];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float* wc = s_w + (pl * 8 + e) * 12;
          const float4 wa = *reinterpret_cast<const float4*>(wc);
          const float4 wb = *reinterpret_cast<const float4*>(wc + 4);
          float v = xin[0] * wa.x;
          v = fmaf(xin[1], wa.y, v); v = fmaf(xin[2], wa.z, v); v = fmaf(xin[3], wa.w, v);
          v = fmaf(xin[4], wb.x, v); v = fmaf(xin[5], wb.y, v); v = fmaf(xin[6], wb.z, v);
          v = fmaf(xin[7], wb.w, v); v = fmaf(xin[8], wc[8], v);
          acc[e] += fmaxf(v, 0.f);
        }
      }
    uint4 o;
    __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(acc[2 * e] * inv, acc[2 * e + 1] * inv);
    dst[pl * plane_stride] = o;
  }
}

// conv_0 without pooling (res15): each thread computes 4 consecutive pixels x all channels, so
// every weight read from shared memory feeds 4 FMAs; stores are 64 contiguous bytes per plane.
constexpr int kC0Px = 4;
__global__ void __launch_bounds__(256)
conv0_p8_w4_kernel(const float* __restrict__ feat, const float* __restrict__ w0, __nv_bfloat16* __restrict__ out,
                   float* __restrict__ pool_sum, int T, int F, int C, int NP, int groups_per_row,
                   int rows_per_tile, int Hpad) {
  extern __shared__ __align__(16) float smem_f[];
  const int in_rows = rows_per_tile + 2, in_cols = groups_per_row * kC0Px + 2;
  float* s_in = smem_f;
  float* s_w = smem_f + round_up(in_rows * in_cols, 4);   // [NP*8][12]
  const int64_t b = blockIdx.y;
  const int h0 = blockIdx.x * rows_per_tile;
  if (pool_sum != nullptr && blockIdx.x == 0 && threadIdx.x < NP * 8) pool_sum[b * NP * 8 + threadIdx.x] = 0.f;
  const float* src = feat + b * (int64_t)T * F;
  for (int i = threadIdx.x; i < in_rows * in_cols; i += blockDim.x) {
    const int r = i / in_cols, c = i - r * in_cols;
    const int h = h0 - 1 + r, w = c - 1;
    s_in[i] = (h >= 0 && h < T && w >= 0 && w < F) ? __ldg(src + (int64_t)h * F + w) : 0.f;
  }
  for (int i = threadIdx.x; i < NP * 8 * 12; i += blockDim.x) {
    const int c = i / 12, k = i - c * 12;
    s_w[i] = (k < 9 && c < C) ? __ldg(w0 + c * 9 + k) : 0.f;
  }
  __syncthreads();
  const int r = threadIdx.x / groups_per_row, gx = threadIdx.x - r * groups_per_row;
  const int h = h0 + r;
  if (r >= rows_per_tile || h >= T) return;
  const int w0px = gx * kC0Px;
  float pch[3][kC0Px + 2];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int e = 0; e < kC0Px + 2; ++e) pch[a][e] = s_in[(r + a) * in_cols + w0px + e];
  const int64_t plane_stride = (int64_t)Hpad * F;
  uint4* dst = reinterpret_cast<uint4*>(out) + (b * NP) * plane_stride + (int64_t)h * F + w0px;
  for (int pl = 0; pl < NP; ++pl) {
    float acc[kC0Px][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float* wc = s_w + (pl * 8 + e) * 12;
      const float4 wa = *reinterpret_cast<const float4*>(wc);
      const float4 wb = *reinterpret_cast<const float4*>(wc + 4);
      const float w8 = wc[8];
#pragma unroll
      for (int px = 0; px < kC0Px; ++px) {
        float v = pch[0][px] * wa.x;
        v = fmaf(pch[0][px + 1], wa.y, v); v = fmaf(pch[0][px + 2], wa.z, v);
        v = fmaf(pch[1][px], wa.w, v); v = fmaf(pch[1][px + 1], wb.x, v); v = fmaf(pch[1][px + 2], wb.y, v);
        v = fmaf(pch[2][px], wb.z, v); v = fmaf(pch[2][px + 1], wb.w, v); v = fmaf(pch[2][px + 2], w8, v);
        acc[px][e] = fmaxf(v, 0.f);
      }
    }
#pragma unroll
    for (int px = 0; px < kC0Px; ++px) {
      if (w0px + px < F) {
        uint4 o;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(acc[px][2 * e], acc[px][2 * e + 1]);
        dst[pl * plane_stride + px] = o;
      }
    }
  }
}

// Linear on the pooled sums produced by the last convolution's epilogue (resnet.py:57-59).
__global__ void __launch_bounds__(256)
tail_pool_kernel(const float* __restrict__ pool_sum, const float* __restrict__ scale, const float* __restrict__ out_w,
                 const float* __restrict__ out_b, float* __restrict__ logits, int64_t B, int C, int CP, int HW,
                 int n_labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_labels) return;
  const int64_t b = i / n_labels;
  const int l = (int)(i - b * n_labels);
  const float inv = 1.f / (float)HW;
  float v = __ldg(out_b + l);
  // pool_sum holds sums of z = x - mean; BatchNorm output mean = z_mean / sigma (resnet.py:55-58)
  for (int c = 0; c < C; ++c) v = fmaf(pool_sum[b * CP + c] * inv * __ldg(scale + c), __ldg(out_w + l * C + c), v);
  logits[i] = v;
}

// mean over H*W of planar-8 bf16 + Linear (resnet.py:57-59).  One CTA per utterance.
__global__ void __launch_bounds__(256)
tail_p8_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ out_w, const float* __restrict__ out_b,
               float* __restrict__ logits, int C, int NP, int HW, int n_labels) {
  __shared__ float s_part[8][64];
  __shared__ float s_mean[64];
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pl = 0; pl < NP; ++pl) {
    const uint4* src = reinterpret_cast<const uint4*>(y) + (b * NP + pl) * (int64_t)HW;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const uint4 v = __ldg(src + i);
      const __nv_bfloat162* vb = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(vb[e]);
        s[2 * e] += f.x;
        s[2 * e + 1] += f.y;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
      if (lane == 0) s_part[warp][pl * 8 + e] = s[e];
    }
  }
  __syncthreads();
  if (threadIdx.x < NP * 8) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += s_part[wv][threadIdx.x];
    s_mean[threadIdx.x] = t / (float)HW;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < n_labels; l += 256) {
    float v = __ldg(out_b + l);
    for (int c = 0; c < C; ++c) v = fmaf(s_mean[c], __ldg(out_w + l * C + c), v);
    logits[b * n_labels + l] = v;
  }
}

// torch [C][C][3][3] fp32 -> [9][NKC][2][CP][8] bf16 (tap, 16-ch chunk, K half, cout, 8 cin)
// `in_scale` (nullable): BN scale 1/sigma of the PREVIOUS layer, folded into the input-channel axis
// (exact: the convolution is linear and zero padding stays zero).
// `out_lo` (nullable): the residual v - float(bf16(v)) in the same layout (the lo weight set of the split-bf16 mode).
__global__ void pack_conv3x3_tc_kernel(const float* __restrict__ w, const float* __restrict__ in_scale,
                                       __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int C, int NKC) {
  const int CP = 16 * NKC;
  const int total = 9 * NKC * 2 * CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int t = i >> 3;
    const int co = t % CP; t /= CP;
    const int half = t & 1; t >>= 1;
    const int kc = t % NKC;
    const int tap = t / NKC;
    const int ci = kc * 16 + half * 8 + e;
    float v = (co < C && ci < C) ? w[((int64_t)co * C + ci) * 9 + tap] : 0.f;
    if (in_scale != nullptr && ci < C) v *= in_scale[ci];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = hi;
    if (out_lo != nullptr) out_lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// torch [C][C][3][3] fp32 -> [NKC][3 dh][2 K halves][3 blocks][CP][8] bf16 for the column-sweep kernel
// (resnet_sweep.cuh): block k of a (16-channel chunk, height tap) slab holds the width tap dw = 2 - k, so that
// one N = 3*CP MMA on input column w feeds output columns w-d, w, w+d.  `in_scale` as above.
__global__ void pack_conv3x3_sw_kernel(const float* __restrict__ w, const float* __restrict__ in_scale,
                                       __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int C, int NKC) {
  const int CP = 16 * NKC;
  const int total = NKC * 3 * 2 * 3 * CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int t = i >> 3;
    const int co = t % CP; t /= CP;
    const int blk = t % 3; t /= 3;
    const int half = t & 1; t >>= 1;
    const int dh = t % 3;
    const int kc = t / 3;
    const int ci = kc * 16 + half * 8 + e;
    const int tap = dh * 3 + (2 - blk);
    float v = (co < C && ci < C) ? w[((int64_t)co * C + ci) * 9 + tap] : 0.f;
    if (in_scale != nullptr && ci < C) v *= in_scale[ci];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = hi;
    if (out_lo != nullptr) out_lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// conv_0 weights [C][1][3][3] fp32 -> one-chunk sweep slab set [3 dh][2 K halves][3 blocks][CP][8] bf16: the staged
// "activation" of the conv_0 pseudo-layer carries the bf16 high and low parts of the feature in channels 0 and 1, so
// the weight sits in both positions (k = 0, 1 of K half 0) and everything else is zero.  split != 0 (bf16x3): channel 2
// carries the feature's high part again and k = 2 holds the weight's residual w - float(bf16(w)).
__global__ void pack_conv0_sw_kernel(const float* __restrict__ w0, __nv_bfloat16* __restrict__ out, int C, int CP, int split) {
  const int total = 3 * 2 * 3 * CP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int t = i >> 3;
    const int co = t % CP; t /= CP;
    const int blk = t % 3; t /= 3;
    const int half = t & 1; t >>= 1;
    const int dh = t;
    const int tap = dh * 3 + (2 - blk);
    const float w = (half == 0 && e < 3 && co < C) ? w0[co * 9 + tap] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    out[i] = e < 2 ? hi : (split && e == 2) ? __float2bfloat16_rn(w - __bfloat162float(hi)) : __float2bfloat16_rn(0.f);
  }
}

// Per-layer epilogue constants.  The activation stored by layer i is z_i = x_i - mean_i (the 1/sigma
// factor is folded into layer i+1's weights).  The skip tensor x_{i-2} of an even layer is not stored:
// it is z_{i-2} + mean_{i-2}, so z_i = ReLU(conv) + z_{i-2} + (mean_{i-2} - mean_i)   (resnet.py:49-55).
// `skip_mean` is nullptr for odd layers and for layer 2 (whose skip is the un-normalised conv_0 output).
__global__ void pad_bn_kernel(const float* __restrict__ scale, const float* __restrict__ mean,
                              const float* __restrict__ skip_mean, float* __restrict__ scale_p,
                              float* __restrict__ kconst_p, int C, int CP) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < CP) {
    scale_p[c] = c < C ? scale[c] : 0.f;
    kconst_p[c] = c < C ? (skip_mean != nullptr ? skip_mean[c] : 0.f) - mean[c] : 0.f;
  }
}

// rows [H, Hpad) of every plane are read by the phase-tiled TMA boxes and must be zero
__global__ void zero_pad_rows_kernel(uint4* __restrict__ buf, int64_t planes, int H, int Hpad, int W) {
  const int64_t per_plane = (int64_t)(Hpad - H) * W;
  const int64_t total = planes * per_plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pl = i / per_plane;
    buf[pl * (int64_t)Hpad * W + (int64_t)H * W + (i - pl * per_plane)] = make_uint4(0, 0, 0, 0);
  }
}

// ---------------------------------------------------------------------------------------------
// host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

struct TcResNet {
  kws_resnet_config cfg{};
  int NKC = 0, CP = 0, NP = 0;
  bool supported = false;
  int n_sms = 148;
  void* blob = nullptr;
  // per layer; every packed weight array is [hi set][lo set] (the lo set = residuals, read by the split-bf16 mode only)
  std::vector<__nv_bfloat16*> wpack;        // position-major layout
  std::vector<__nv_bfloat16*> wpack_sw;     // column-sweep layout (resnet_sweep.cuh)
  std::vector<float*> scale_p, shift_p;     // per layer, padded to CP: BN 1/sigma and the epilogue constant
  float* conv0_w = nullptr;                 // [C][9]
  __nv_bfloat16* conv0_wb = nullptr;        // conv_0 as a one-chunk sweep slab set (pack_conv0_sw_kernel)
  __nv_bfloat16* conv0_wb3 = nullptr;       // same for the split-bf16 mode
  float* out_w = nullptr;
  float* out_b = nullptr;
  std::map<std::tuple<const void*, int64_t, int, int, int>, CUtensorMap> maps;
  // whole-network persistent kernel (resnet_fused.cuh): device copies of the layer table and tensor maps,
  // rebuilt when the input shape or the workspace changes
  bool fused_enabled = true;
  void* fused_dev = nullptr;   // [TcLayerDesc x n_layers][pad][CUtensorMap x n_layers]
  std::tuple<int, int, const void*, int> fused_key{-1, -1, nullptr, 0};
  TcFusedParams fused_prm{};
  int fused_smem = 0;
  // column-sweep whole-network kernel (resnet_sweep.cuh), preferred when the map is tall enough
  bool sweep_enabled = true;
  bool sweep_k32 = true;       // HONK2_TC_SWEEP_K32=0: single-strip maps use the planar layout of the multi-strip maps too
  int l2_policy = 1, wait_polls = 48, sweep_max_stages = 0, sweep_min_pct = 60;
  bool sweep_discard = true;
  bool sweep_pack = true;      // HONK2_TC_SWEEP_PACK=0: short / pooled maps stay on the position-major kernel
  // chunk pipelining: consecutive chunks run on `lanes` internal streams so that the prologue / tail of one
  // chunk's layer kernels overlaps the steady state of another's (each lane has its own activation buffers)
  int lanes = 1;
  cudaStream_t lane_stream[kTcMaxLanes] = {};
  cudaEvent_t ev_start = nullptr, ev_done[kTcMaxLanes] = {};
};

static int tc_max_mt(int CP) {
  // M-tiles per tile, bounded by the TMEM accumulator buffer (HONK2_TC_MAXMT lowers it, for experiments:
  // measured on B200, larger tiles win even though 5 M-tiles do not divide evenly over 3 issuer warps)
  static const int env = [] { const char* e = std::getenv("HONK2_TC_MAXMT"); return e ? std::atoi(e) : 0; }();
  const int cap = std::min(kTcMaxMt, kAccCols / CP);
  return env > 0 ? std::max(1, std::min(cap, env)) : cap;
}

// Phase tiling pays when the dilation is large: a tile of R consecutive rows needs R + 2d (or 3R) input
// rows, a tile of R rows spaced d apart needs R + 2.
static bool tc_use_phase(int H, int d) { return d >= 8 && d <= 32 && 2 * d <= H; }

// Rows per plane in memory: H rounded up to the largest phase-tiled dilation of the network.
static int tc_hpad(const kws_resnet_config& c, int H) {
  int m = 1;
  for (int i = 1; i <= c.n_layers; ++i) {
    const int d = c.use_dilation ? (1 << ((i - 1) / 3)) : 1;
    if (tc_use_phase(H, d)) m = std::max(m, d);
  }
  return round_up(H, m);
}

// Tile geometry of one layer launch.  Returns false if the layer cannot be tiled.
static bool tc_geom(int NKC, int H, int Hpad, int W, int d, TcGeom* g) {
  const int CP = 16 * NKC;
  g->H = H; g->W = W; g->d = d; g->Hpad = Hpad;
  g->side_taps = d < W ? 1 : 0;
  g->dpad = g->side_taps ? d : 0;
  g->Wp = W + g->dpad;
  if (g->Wp > 256) return false;
  const int max_pos = tc_max_mt(CP) * 128;
  const int w_bytes = 9 * NKC * 2 * CP * 16;
  g->smem_w_off = 256 + round_up(2 * CP * 4, 128);
  g->smem_ring_off = round_up(g->smem_w_off + w_bytes, 1024);
  const int budget = 227 * 1024 - g->smem_ring_off - 4096;
  g->phase = (tc_use_phase(H, d) && Hpad % d == 0) ? 1 : 0;
  g->chunks_per_phase = 1;
  if (g->phase) {
    // rows of one phase: p, p+d, ...; in phase-row units the convolution has dilation 1 in h
    const int rows_max = ceil_div(H, d);
    const int Rmax = std::min(rows_max, max_pos / g->Wp);
    if (Rmax < 1) return false;
    int best_R = 0, best_mt = 1 << 30;
    for (int R = Rmax; R >= 1; --R) {   // fewest issued M-tiles over all phases
      int mt = 0;
      for (int ph = 0; ph < d; ++ph) {
        const int n = ceil_div(H - ph, d);
        mt += (n / R) * ceil_div(R * g->Wp, 128) + ceil_div((n % R) * g->Wp, 128);
      }
      if (mt < best_mt) { best_mt = mt; best_R = R; }
    }
    const int R = best_R;
    g->R = R;
    g->chunks_per_phase = ceil_div(rows_max, R);
    g->tiles_per_utt = d * g->chunks_per_phase;
    g->n_boxes = 1;
    g->rows_box = R + 3;   // one halo row above and below + the spill row
    g->box_stride = round_up(g->rows_box * g->Wp * 16, 128);
    g->slab_bytes = g->box_stride;
    g->stage_bytes = 2 * g->slab_bytes;
    for (int k = 0; k < 3; ++k) {
      g->h_start[k] = -1;
      g->tap_off[k] = k * g->Wp * 16;
    }
    g->n_stages = std::min(kMaxStages, budget / g->stage_bytes);
    if (g->n_stages < 2) return false;
  } else {
    const int Rmax = std::min(H, max_pos / g->Wp);
    if (Rmax < 1) return false;
    // pick the rows-per-tile that issues the fewest 128-position M-tiles per utterance among those whose
    // staged input fits shared memory with >= 2 stages
    int best_R = 0, best_mt = 1 << 30, best_stages = 0;
    for (int R = Rmax; R >= 1; --R) {
      const int full = H / R, rem = H - full * R;
      const int mt = full * ceil_div(R * g->Wp, 128) + ceil_div(rem * g->Wp, 128);
      if (mt >= best_mt) continue;
      const bool dense = d <= R;
      const int rows_box = dense ? R + 2 * d + 1 : R + 1;
      if (rows_box > 256) continue;
      const int slab = (dense ? 1 : 3) * round_up(rows_box * g->Wp * 16, 128);
      if (slab >= (1 << 18)) continue;
      const int stages = std::min(kMaxStages, budget / (2 * slab));
      if (stages < 2) continue;
      best_R = R; best_mt = mt; best_stages = stages;
    }
    if (best_R == 0) return false;
    const int R = best_R;
    const bool dense = d <= R;
    g->R = R;
    g->tiles_per_utt = ceil_div(H, R);
    g->n_boxes = dense ? 1 : 3;
    g->rows_box = dense ? R + 2 * d + 1 : R + 1;
    g->box_stride = round_up(g->rows_box * g->Wp * 16, 128);
    g->slab_bytes = g->n_boxes * g->box_stride;
    g->stage_bytes = 2 * g->slab_bytes;
    for (int k = 0; k < 3; ++k) {
      g->h_start[k] = dense ? -d : (k - 1) * d;
      g->tap_off[k] = dense ? k * d * g->Wp * 16 : k * g->box_stride;
    }
    g->n_stages = best_stages;
  }
  g->smem_total = g->smem_ring_off + g->n_stages * g->stage_bytes + 4096;
  if (g->smem_total < 120 * 1024) g->smem_total = 120 * 1024;   // one CTA per SM (it owns all 512 TMEM columns)
  return true;
}

int tc_resnet_create(const kws_resnet_config& cfg, TcResNet** out) {
  TcResNet* p = new TcResNet();
  p->cfg = cfg;
  const int C = cfg.n_maps;
  p->NKC = ceil_div(C, 16);
  p->CP = 16 * p->NKC;
  p->NP = 2 * p->NKC;
  p->supported = p->NKC >= 1 && p->NKC <= 4 && get_encode_fn() != nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&p->n_sms, cudaDevAttrMultiProcessorCount, dev);
  if (p->supported) {
    const int n = cfg.n_layers;
    // one weight set is a multiple of 256 bytes (CP is a multiple of 16), so [hi set][lo set] is contiguous
    const size_t w_bytes = 2 * (size_t)9 * p->NKC * 2 * p->CP * 16;
    const size_t v_bytes = round_up<size_t>(p->CP * sizeof(float), 256);
    const size_t c0_bytes = round_up<size_t>((size_t)3 * 2 * 3 * p->CP * 16, 256);
    const size_t total = n * (2 * w_bytes + 2 * v_bytes) + 2 * c0_bytes + round_up<size_t>(C * 9 * 4, 256) +
                         round_up<size_t>((size_t)cfg.n_labels * C * 4, 256) + round_up<size_t>(cfg.n_labels * 4, 256);
    if (cudaMalloc(&p->blob, total) != cudaSuccess) {
      set_error("tc_resnet_create: cudaMalloc(%zu) failed", total);
      delete p;
      return KWS_ERR_CUDA;
    }
    char* b = static_cast<char*>(p->blob);
    for (int i = 0; i < n; ++i) {
      p->wpack.push_back(reinterpret_cast<__nv_bfloat16*>(b)); b += w_bytes;
      p->wpack_sw.push_back(reinterpret_cast<__nv_bfloat16*>(b)); b += w_bytes;
      p->scale_p.push_back(reinterpret_cast<float*>(b)); b += v_bytes;
      p->shift_p.push_back(reinterpret_cast<float*>(b)); b += v_bytes;
    }
    p->conv0_wb = reinterpret_cast<__nv_bfloat16*>(b); b += c0_bytes;
    p->conv0_wb3 = reinterpret_cast<__nv_bfloat16*>(b); b += c0_bytes;
    p->conv0_w = reinterpret_cast<float*>(b); b += round_up<size_t>(C * 9 * 4, 256);
    p->out_w = reinterpret_cast<float*>(b); b += round_up<size_t>((size_t)cfg.n_labels * C * 4, 256);
    p->out_b = reinterpret_cast<float*>(b);
    // every knob is read HERE, once per model handle (DESIGN.md section 4 lists them)
    auto env_int = [](const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; };
    p->fused_enabled = env_int("HONK2_TC_FUSED", 1) != 0;
    p->sweep_enabled = env_int("HONK2_TC_SWEEP", 1) != 0;
    p->sweep_k32 = env_int("HONK2_TC_SWEEP_K32", 1) != 0;
    p->sweep_discard = env_int("HONK2_TC_SWEEP_DISCARD", 1) != 0;
    p->sweep_pack = env_int("HONK2_TC_SWEEP_PACK", 1) != 0;
    p->sweep_max_stages = env_int("HONK2_TC_SWEEP_STAGES", kSwMaxStages);
    p->sweep_min_pct = env_int("HONK2_TC_SWEEP_MINPCT", 60);
    p->l2_policy = env_int("HONK2_TC_L2POLICY", 1);
    p->wait_polls = env_int("HONK2_TC_WAIT_POLLS", 48);
    p->lanes = std::max(1, std::min(kTcMaxLanes, env_int("HONK2_TC_LANES", 2)));
    bool ok = cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) == cudaSuccess;
    for (int l = 0; l < p->lanes && ok; ++l)
      ok = cudaStreamCreateWithFlags(&p->lane_stream[l], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->ev_done[l], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      set_error("tc_resnet_create: could not create the chunk-pipelining streams");
      tc_resnet_destroy(p);
      return KWS_ERR_CUDA;
    }
  }
  *out = p;
  return KWS_OK;
}

void tc_resnet_destroy(TcResNet* p) {
  if (!p) return;
  if (p->blob) cudaFree(p->blob);
  if (p->fused_dev) cudaFree(p->fused_dev);
  for (int l = 0; l < kTcMaxLanes; ++l) {
    if (p->lane_stream[l]) cudaStreamDestroy(p->lane_stream[l]);
    if (p->ev_done[l]) cudaEventDestroy(p->ev_done[l]);
  }
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  delete p;
}

int tc_resnet_set_weights(TcResNet* p, const kws_resnet_weights& w, float* const* bn_scale, float* const* bn_shift,
                          cudaStream_t st) {
  if (!p->supported) return KWS_OK;
  const int C = p->cfg.n_maps, n = p->cfg.n_layers, L = p->cfg.n_labels;
  for (int i = 0; i < n; ++i) {
    const int set_elems = 9 * p->NKC * 2 * p->CP * 8;   // one weight set; the lo set follows it
    pack_conv3x3_tc_kernel<<<ceil_div(set_elems, 256), 256, 0, st>>>(
        w.conv_w[i], i > 0 ? bn_scale[i - 1] : nullptr, p->wpack[i], p->wpack[i] + set_elems, C, p->NKC);
    KWS_CUDA(cudaGetLastError());
    pack_conv3x3_sw_kernel<<<ceil_div(set_elems, 256), 256, 0, st>>>(
        w.conv_w[i], i > 0 ? bn_scale[i - 1] : nullptr, p->wpack_sw[i], p->wpack_sw[i] + set_elems, C, p->NKC);
    KWS_CUDA(cudaGetLastError());
    // layer number i+1 is even and > 2  <=>  its skip comes from a normalised layer (i-1 in 0-based terms)
    const float* skip_mean = ((i + 1) % 2 == 0 && i >= 3) ? w.bn_mean[i - 2] : nullptr;
    pad_bn_kernel<<<1, 64, 0, st>>>(bn_scale[i], w.bn_mean[i], skip_mean, p->scale_p[i], p->shift_p[i], C, p->CP);
    KWS_CUDA(cudaGetLastError());
  }
  KWS_CUDA(cudaMemcpyAsync(p->conv0_w, w.conv0_w, sizeof(float) * C * 9, cudaMemcpyDeviceToDevice, st));
  pack_conv0_sw_kernel<<<ceil_div(3 * 2 * 3 * p->CP * 8, 256), 256, 0, st>>>(w.conv0_w, p->conv0_wb, C, p->CP, 0);
  KWS_CUDA(cudaGetLastError());
  pack_conv0_sw_kernel<<<ceil_div(3 * 2 * 3 * p->CP * 8, 256), 256, 0, st>>>(w.conv0_w, p->conv0_wb3, C, p->CP, 1);
  KWS_CUDA(cudaGetLastError());
  KWS_CUDA(cudaMemcpyAsync(p->out_w, w.out_w, sizeof(float) * L * C, cudaMemcpyDeviceToDevice, st));
  KWS_CUDA(cudaMemcpyAsync(p->out_b, w.out_b, sizeof(float) * L, cudaMemcpyDeviceToDevice, st));
  return KWS_OK;
}

static void tc_map_hw(const kws_resnet_config& c, int T, int F, int* H, int* W) {
  const int ph = c.pool_h > 0 ? c.pool_h : 1, pw = c.pool_w > 0 ? c.pool_w : 1;
  *H = T / ph;
  *W = F / pw;
}

static int64_t tc_chunk(const TcResNet* p, int64_t B, int H, int W, int chunk) {
  const int64_t per = (int64_t)p->NP * H * W * 16;
  int64_t c = chunk > 0 ? chunk : (256ll << 20) / (2 * std::max<int64_t>(per, 1));
  if (chunk <= 0) {
    // whole number of CTA waves for the common 7-tile geometry is not knowable here; keep it simple
    if (c < 16) c = 16;
    if (c > 2048) c = 2048;
  }
  if (c > B) c = B;
  if (c < 1) c = 1;
  return c;
}

static bool tc_layers_ok(const TcResNet* p, int H, int W) {
  TcGeom g;
  const int Hpad = tc_hpad(p->cfg, H);
  for (int i = 1; i <= p->cfg.n_layers; ++i) {
    const int d = p->cfg.use_dilation ? (1 << ((i - 1) / 3)) : 1;
    if (!tc_geom(p->NKC, H, Hpad, W, d, &g)) return false;
  }
  return true;
}

static size_t tc_lane_bytes(const TcResNet* p, int64_t chunk, int H, int W, size_t* buf_out) {
  const size_t buf = round_up<size_t>((size_t)chunk * p->NP * tc_hpad(p->cfg, H) * W * 16, 1024);
  if (buf_out) *buf_out = buf;
  return 2 * buf + round_up<size_t>((size_t)chunk * p->CP * 4, 1024);   // two activation buffers + pooled sums
}

static int tc_encode_map(const void* base, int64_t planes, int H, int W, const TcGeom& g, CUtensorMap* out);

static int tc_get_map(TcResNet* p, const void* base, int64_t planes, int H, int W, const TcGeom& g, CUtensorMap** out) {
  auto key = std::make_tuple(base, planes, H, W, (g.d * 1024 + g.rows_box) * 2 + g.phase);
  auto it = p->maps.find(key);
  if (it == p->maps.end()) {
    if (p->maps.size() > 256) p->maps.clear();
    CUtensorMap m;
    KWS_TRY(tc_encode_map(base, planes, H, W, g, &m));
    it = p->maps.emplace(key, m).first;
  }
  *out = &it->second;
  return KWS_OK;
}

static int tc_encode_map(const void* base, int64_t planes, int H, int W, const TcGeom& g, CUtensorMap* out) {
  {
    {
    CUtensorMap& m = *out;
    // dims: 8 channels, W, rows (of one phase), phases, planes.  Plain tiling is the 1-phase case.
    const cuuint64_t row_stride = (cuuint64_t)W * 16, plane_stride = (cuuint64_t)g.Hpad * W * 16;
    cuuint64_t dims[5], strides[4];
    dims[0] = 8; dims[1] = (cuuint64_t)W; dims[4] = (cuuint64_t)planes;
    strides[0] = 16; strides[3] = plane_stride;
    if (g.phase) {
      dims[2] = (cuuint64_t)(g.Hpad / g.d); dims[3] = (cuuint64_t)g.d;
      strides[1] = row_stride * g.d; strides[2] = row_stride;
    } else {
      dims[2] = (cuuint64_t)H; dims[3] = 1;
      strides[1] = row_stride; strides[2] = plane_stride;
    }
    const cuuint32_t box[5] = {8, (cuuint32_t)g.Wp, (cuuint32_t)g.rows_box, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed (%d) for W=%d H=%d planes=%lld box=%dx%d phase=%d", (int)r, W, H,
                (long long)planes, g.Wp, g.rows_box, g.phase);
      return KWS_ERR_CUDA;
    }
    }
  }
  return KWS_OK;
}

// ---- whole-network persistent kernel: geometry, shared-memory plan, device tables ----------------
struct TcFusedPlan {
  bool ok = false, split = false;
  int Hpad = 0, n_slots = 0, smem_total = 0, n_stages = 0;
  int w_off[2] = {0, 0}, ring_off = 0, slot_bytes = 0;
  std::vector<TcGeom> geoms;
};

static TcFusedPlan tc_fused_plan(const TcResNet* p, int H, int W, bool split) {
  TcFusedPlan f;
  const kws_resnet_config& c = p->cfg;
  if ((!p->fused_enabled && !split) || c.n_layers < 1 || c.n_layers > kFusedMaxLayers || c.n_labels > 4096) return f;
  f.Hpad = tc_hpad(c, H);
  f.n_slots = p->n_sms;
  f.split = split;
  const int w_bytes = (split ? 2 : 1) * 9 * p->NKC * 2 * p->CP * 16;
  f.w_off[0] = 6144;   // control block: barriers, constants, layer descriptors, conv_0 weights, pooled sums
  f.w_off[1] = split ? f.w_off[0] : f.w_off[0] + round_up(w_bytes, 128);   // (split: one buffer for both weight sets)
  f.ring_off = round_up(f.w_off[1] + w_bytes, 1024);
  for (int i = 1; i <= c.n_layers; ++i) {
    TcGeom g;
    const int d = c.use_dilation ? (1 << ((i - 1) / 3)) : 1;
    // (tc_geom sizes its tiles for the layer-per-launch kernel's shared memory; the check below is this kernel's)
    if (!tc_geom(p->NKC, H, f.Hpad, W, d, &g)) return f;
    f.slot_bytes = std::max(f.slot_bytes, round_up(g.stage_bytes, 128));
    f.geoms.push_back(g);
  }
  // as many ring slots as fit (large dilations stage three row-blocks per tile: fewer, larger slots)
  f.n_stages = std::min(kFusedStages, (227 * 1024 - 4096 - f.ring_off) / std::max(f.slot_bytes, 1));
  f.smem_total = f.ring_off + f.n_stages * f.slot_bytes + 4096;
  f.ok = f.n_stages >= 2;
  return f;
}

static size_t tc_fused_ws_bytes(const TcResNet* p, const TcFusedPlan& f, int W, size_t* buf_out) {
  const size_t buf = round_up<size_t>((size_t)f.n_slots * (f.split ? 2 : 1) * p->NP * f.Hpad * W * 16, 1024);
  if (buf_out) *buf_out = buf;
  return 2 * buf;
}

template <int NKC, bool SPLIT>
static int tc_launch_fused_v(const TcFusedParams& prm, int grid, int smem, cudaStream_t st) {
  KWS_CUDA(cudaFuncSetAttribute(resnet_tc_fused_kernel<NKC, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  resnet_tc_fused_kernel<NKC, SPLIT><<<grid, tc_threads(NKC), smem, st>>>(prm);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}
template <int NKC>
static int tc_launch_fused(const TcFusedParams& prm, bool split, int grid, int smem, cudaStream_t st) {
  return split ? tc_launch_fused_v<NKC, true>(prm, grid, smem, st) : tc_launch_fused_v<NKC, false>(prm, grid, smem, st);
}

static int tc_fused_forward(TcResNet* p, const TcFusedPlan& f, const float* feat, int64_t B, int T, int F, int H, int W,
                            float* logits, void* ws, LaunchProfiler* prof, cudaStream_t st) {
  const kws_resnet_config& c = p->cfg;
  size_t buf = 0;
  tc_fused_ws_bytes(p, f, W, &buf);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws));
  __nv_bfloat16* Q = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + buf);
  const auto key = std::make_tuple(T, F, (const void*)ws, f.split ? 1 : 0);
  if (p->fused_key != key) {
    // (re)build the device tables: layer descriptors + one input tensor map per layer
    const int n = c.n_layers;
    const size_t maps_off = round_up<size_t>(sizeof(TcLayerDesc) * n, 64);
    const size_t total = maps_off + sizeof(CUtensorMap) * n;
    std::vector<unsigned char> host(total, 0);
    for (int i = 1; i <= n; ++i) {
      TcLayerDesc L{};
      L.g = f.geoms[i - 1];
      L.wpack = p->wpack[i - 1];
      L.kconst = p->shift_p[i - 1];
      L.has_skip = (i % 2 == 0) ? 1 : 0;
      L.in_buf = L.has_skip ? 1 : 0;
      L.last = (i == n) ? 1 : 0;
      memcpy(host.data() + sizeof(TcLayerDesc) * (i - 1), &L, sizeof(L));
      CUtensorMap m;
      KWS_TRY(tc_encode_map(L.in_buf ? Q : P, (int64_t)f.n_slots * (f.split ? 2 : 1) * p->NP, H, W, L.g, &m));
      memcpy(host.data() + maps_off + sizeof(CUtensorMap) * (i - 1), &m, sizeof(m));
    }
    if (p->fused_dev) { KWS_CUDA(cudaStreamSynchronize(st)); KWS_CUDA(cudaFree(p->fused_dev)); p->fused_dev = nullptr; }
    KWS_CUDA(cudaMalloc(&p->fused_dev, total));
    KWS_CUDA(cudaMemcpyAsync(p->fused_dev, host.data(), total, cudaMemcpyHostToDevice, st));
    KWS_CUDA(cudaStreamSynchronize(st));   // `host` goes out of scope; this happens once per shape
    TcFusedParams& q = p->fused_prm;
    q = TcFusedParams{};
    q.layers = reinterpret_cast<const TcLayerDesc*>(p->fused_dev);
    q.maps = reinterpret_cast<const CUtensorMap*>(static_cast<char*>(p->fused_dev) + maps_off);
    q.conv0_w = p->conv0_w;
    q.last_scale = p->scale_p[n - 1];
    q.out_w = p->out_w;
    q.out_b = p->out_b;
    q.P = P; q.Q = Q;
    q.n_layers = n; q.C = c.n_maps; q.n_labels = c.n_labels; q.T = T; q.F = F;
    q.ph = c.pool_h > 0 ? c.pool_h : 1; q.pw = c.pool_w > 0 ? c.pool_w : 1;
    q.H = H; q.W = W; q.Hpad = f.Hpad;
    q.smem_w_off[0] = f.w_off[0]; q.smem_w_off[1] = f.w_off[1];
    q.smem_ring_off = f.ring_off; q.ring_slot_bytes = f.slot_bytes; q.n_stages = f.n_stages;
    q.l2_policy = p->l2_policy;
    p->fused_smem = f.smem_total;
    p->fused_key = key;
  }
  if (prof) prof->tick(1, st);
  if (f.Hpad > H) {
    for (__nv_bfloat16* bp : {P, Q}) {
      zero_pad_rows_kernel<<<256, 256, 0, st>>>(reinterpret_cast<uint4*>(bp), (int64_t)f.n_slots * (f.split ? 2 : 1) * p->NP, H, f.Hpad, W);
      KWS_CHECK_LAUNCH();
    }
  }
  if (prof) prof->tick(0, st);   // the whole network is one launch: it IS the dominant kernel
  TcFusedParams prm = p->fused_prm;
  prm.feat = feat;
  prm.logits = logits;
  prm.B = B;
  const int grid = (int)std::min<int64_t>(f.n_slots, B);
  static const bool dbg_on = [] { const char* e = std::getenv("HONK2_TC_DEBUG"); return e && std::atoi(e) != 0; }();
  static long long* dbg_buf = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
    prm.debug = dbg_buf;
  }
  struct DbgPrint {
    long long* buf; cudaStream_t st;
    ~DbgPrint() {
      if (!buf) return;
      long long h[16];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
      const double tot = (double)(h[0] + h[1] + h[2] + h[3] + h[4] + h[5]);
      fprintf(stderr, "[fused dbg] issuer 0 of CTA 0, %lld utterances, cycles: conv_0 wait %.1f%%, layer barrier (drain) %.1f%%, "
              "accumulator-free wait %.1f%%, TMA-data wait %.1f%%, issuing MMAs %.1f%% (total %.0f, %.0f per utterance)\n",
              h[6], 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, tot, tot / (double)h[6]);
    }
  } dbg_print{dbg_on ? dbg_buf : nullptr, st};
  switch (p->NKC) {
    case 1: return tc_launch_fused<1>(prm, f.split, grid, p->fused_smem, st);
    case 2: return tc_launch_fused<2>(prm, f.split, grid, p->fused_smem, st);
    case 3: return tc_launch_fused<3>(prm, f.split, grid, p->fused_smem, st);
    default: return tc_launch_fused<4>(prm, f.split, grid, p->fused_smem, st);
  }
}

// ---- column-sweep whole-network kernel (resnet_sweep.cuh): plan, launch ------------------------------
struct TcSweepPlan {
  bool ok = false;
  int n_slots = 0, n_strips = 0, smem_total = 0, n_stages = 0;
  int c0w_off = 0, w_off[2] = {0, 0}, skip_off = 0, ring_off = 0, slot_bytes = 0, slack_bytes = 0;
  int dmax = 1, chunk_rows = 0;
  bool k32 = false, split = false;
  // packed strips (short maps): pack_n utterances stacked in one strip of Hst rows, conv_0 (+ pool) by the pre-pass kernel
  int pack_n = 1, pack_pitch = 0, Hst = 0, pool_off = 0;
  bool packed = false;
};

// (PH, PW) pooling windows the pre-pass kernel of the packed-strip mode is instantiated for
static bool tc_pack_pool_ok(int ph, int pw) { return (ph == 1 && pw == 1) || (ph == 2 && pw == 2) || (ph == 4 && pw == 3); }

static TcSweepPlan tc_sweep_plan(const TcResNet* p, int H, int W, bool split) {
  TcSweepPlan f;
  const int H_map = H;
  (void)H_map;
  const kws_resnet_config& c = p->cfg;
  if (!p->sweep_enabled || c.n_layers < 1 || c.n_layers > kFusedMaxLayers || c.n_labels > 4096) return f;
  if (W < 1 || W > kSwMaxW || H < 1) return f;
  if (p->NKC > 3) return f;   // 64 maps: 640 threads leave 96 registers, the epilogue spills (position-major kernel instead)
  if ((1 + c.n_layers) * p->CP * 4 > kSwKcBytes) return f;   // per-layer epilogue constants live in shared memory
  int dmax = 1;
  for (int i = 1; i <= c.n_layers; ++i) {
    const int d = c.use_dilation ? (1 << ((i - 1) / 3)) : 1;
    if (d > 64) return f;   // (a staged column of 128 + 2d rows per plane would no longer leave room for three stages)
    dmax = std::max(dmax, d);
  }
  const int ph = c.pool_h > 0 ? c.pool_h : 1, pw = c.pool_w > 0 ? c.pool_w : 1;
  const bool pooled = ph > 1 || pw > 1;
  // The 128 lanes of an MMA are 128 rows of one column.  Short maps (res8 / res26 after pooling, short clips) are
  // PACKED: several utterances share a strip, stacked along the rows with dmax zero rows between them, and conv_0
  // (+ ReLU + AvgPool) is computed by a CUDA-core pre-pass (the sweep's own conv_0 pseudo-layer works on the unpooled map).
  const bool is_short = H * 100 < ceil_div(H, 128) * 128 * p->sweep_min_pct;
  if (pooled || is_short) {
    const int pitch = H + dmax;
    const int n = (128 + dmax) / pitch;
    if (!p->sweep_pack || n < 1 || !tc_pack_pool_ok(ph, pw) || !(p->sweep_k32 || split)) return f;
    f.packed = true;
    f.pack_n = n;
    f.pack_pitch = pitch;
    f.Hst = n * pitch - dmax;
    if (f.Hst * 100 < 128 * p->sweep_min_pct && n == 1) return f;   // a single short map gains nothing here
  }
  if (f.packed) H = f.Hst;   // from here on the "map" is the stack
  f.n_strips = ceil_div(H, 128);
  f.n_slots = p->n_sms;
  f.dmax = dmax;
  f.split = split;
  // 16 channels per row: [K chunk][w][h][32 B] in HBM, swizzle-32B operand in shared memory, one bulk copy per chunk
  // and step (single-strip maps; HONK2_TC_SWEEP_K32=0 keeps them on the planar layout of the multi-strip maps)
  f.k32 = f.n_strips == 1 && (p->sweep_k32 || split);
  if (split && !f.k32) return f;   // the split-bf16 mode is built on the 32-byte-row layout
  const int set_bytes = 9 * p->NKC * 2 * p->CP * 16;           // one weight set
  const int w_bytes = (split ? 2 : 1) * set_bytes;
  f.c0w_off = kSwCtrlBytes;
  f.w_off[0] = f.c0w_off + round_up(3 * 2 * 3 * p->CP * 16, 128);
  f.w_off[1] = split ? f.w_off[0] : f.w_off[0] + round_up(w_bytes, 128);   // (split: one buffer, see the kernel)
  f.skip_off = round_up(f.w_off[1] + w_bytes, 1024);
  f.pool_off = f.skip_off + (split ? 0 : sw_groups(p->NKC) * p->NP * 2048);   // one skip slot per epilogue warp group
  f.ring_off = f.pool_off + (f.packed ? round_up(f.pack_n * sw_epi_warps(p->NKC) * p->CP * 4, 1024) : 0);   // packed: pooled sums per stacked utterance
  // rows per chunk of a staged column.  The split mode stages twice the chunks: its chunks are cut down to the rows
  // the map's own lanes read (H + 2 dmax); the lanes past the map then read into the next chunk / the slack.
  const int full_rows = (128 + 2 * dmax + 7) & ~7;
  f.chunk_rows = (split && f.n_strips == 1) ? std::min(full_rows, round_up(H + 2 * dmax, 8)) : full_rows;
  f.slot_bytes = (split ? 2 : 1) * p->NP * f.chunk_rows * 16;
  f.slack_bytes = (full_rows - f.chunk_rows) * 32;
  const int max_stages = std::min(kSwMaxStages, p->sweep_max_stages > 0 ? p->sweep_max_stages : kSwMaxStages);
  f.n_stages = std::min(max_stages, (227 * 1024 - f.slack_bytes - f.ring_off) / f.slot_bytes);
  if (f.n_stages < 3) return f;
  f.smem_total = f.ring_off + f.n_stages * f.slot_bytes + f.slack_bytes;
  f.ok = true;
  return f;
}

// packed strips: utterances per pre-pass + sweep launch pair (a multiple of the group size)
static int64_t tc_pack_chunk(const TcSweepPlan& f, int64_t B, int chunk_cfg) {
  int64_t c = chunk_cfg > 0 ? chunk_cfg : 4096;
  c = round_up<int64_t>(c, f.pack_n);
  return std::min<int64_t>(c, round_up<int64_t>(std::max<int64_t>(B, 1), f.pack_n));
}

static size_t tc_sweep_ws_bytes(const TcResNet* p, const TcSweepPlan& f, int H, int W, int64_t B, int chunk_cfg,
                                size_t* buf_out, size_t* ext_group_out = nullptr) {
  const int Hs = f.packed ? f.Hst : H;
  const size_t buf = round_up<size_t>((size_t)f.n_slots * (f.split ? 2 : 1) * p->NP * Hs * W * 16, 1024);
  if (buf_out) *buf_out = buf;
  size_t ext = 0;
  if (f.packed) {
    const size_t per_group = (size_t)(f.split ? 2 : 1) * p->NP * Hs * W * 16;
    if (ext_group_out) *ext_group_out = per_group;
    ext = round_up<size_t>((size_t)(tc_pack_chunk(f, B, chunk_cfg) / f.pack_n) * per_group, 1024);
  }
  return 2 * buf + ext;
}

// conv_0 (1 -> C, 3x3, pad 1) + ReLU + AvgPool(PH, PW) (resnet.py:40-44) for the packed-strip mode of the sweep kernel:
// writes the bf16 activations of a group of `pack_n` stacked utterances exactly as the sweep kernel's epilogue would
// ([K chunk][w][stacked row][16 ch = 32 B], halves swapped where ((dmax + row) >> 2) & 1, zero rows between the
// utterances; split: the lo parts in chunks NKC .. 2 NKC-1).  One thread = one (column, stacked row) position of a group;
// a group's W * Hst positions are dealt to blocks of 128 threads column by column (rows fastest: 32-byte stores coalesce).
template <int PH, int PW>
__global__ void __launch_bounds__(128)
conv0_pool_pack_kernel(const float* __restrict__ feat, const float* __restrict__ w0, uint4* __restrict__ ext, int64_t B_utt,
                       int T, int F, int C, int NKC, int W, int Hs, int pitch, int pack_n, int Hst, int dmax, int split,
                       int64_t ext_stride) {
  __shared__ __align__(16) float s_w[64 * 12];
  const int CP = 16 * NKC;
  for (int i = threadIdx.x; i < CP * 12; i += blockDim.x) {
    const int c = i / 12, k = i - c * 12;
    s_w[i] = (k < 9 && c < C) ? w0[c * 9 + k] : 0.f;
  }
  __syncthreads();
  const int tiles = (W * Hst + 127) >> 7;
  const int64_t g = blockIdx.x / tiles;
  const int item = (blockIdx.x - (int)(g * tiles)) * 128 + threadIdx.x;
  if (item >= W * Hst) return;
  const int wo = item / Hst, row = item - wo * Hst;
  const int u = row / pitch, ho = row - u * pitch;
  const int64_t b = g * pack_n + u;
  const bool live = ho < Hs && u < pack_n && b < B_utt;
  float patch[PH + 2][PW + 2];
#pragma unroll
  for (int a = 0; a < PH + 2; ++a)
#pragma unroll
    for (int e = 0; e < PW + 2; ++e) {
      const int hh = ho * PH - 1 + a, ww = wo * PW - 1 + e;
      patch[a][e] = (live && hh >= 0 && hh < T && ww >= 0 && ww < F) ? __ldg(feat + (b * T + hh) * (int64_t)F + ww) : 0.f;
    }
  constexpr float inv = 1.f / (float)(PH * PW);
  const int swl = ((dmax + row) >> 2) & 1;
  // (a warp whose 32 items are all gap rows / past the batch only stores zeros; the pad channels C .. CP-1 are zeros too)
  const bool warp_live = __any_sync(0xffffffffu, live);
  for (int kc = 0; kc < NKC; ++kc) {
    float out[16];
#pragma unroll
    for (int cl = 0; cl < 16; ++cl) {
      if (!warp_live || 16 * kc + cl >= C) { out[cl] = 0.f; continue; }
      const float* wc = s_w + (16 * kc + cl) * 12;
      const float4 wa = *reinterpret_cast<const float4*>(wc);
      const float4 wb = *reinterpret_cast<const float4*>(wc + 4);
      const float w8 = wc[8];
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < PH; ++i)
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          float v = patch[i][j] * wa.x;
          v = fmaf(patch[i][j + 1], wa.y, v); v = fmaf(patch[i][j + 2], wa.z, v);
          v = fmaf(patch[i + 1][j], wa.w, v); v = fmaf(patch[i + 1][j + 1], wb.x, v); v = fmaf(patch[i + 1][j + 2], wb.y, v);
          v = fmaf(patch[i + 2][j], wb.z, v); v = fmaf(patch[i + 2][j + 1], wb.w, v); v = fmaf(patch[i + 2][j + 2], w8, v);
          acc += fmaxf(v, 0.f);
        }
      out[cl] = live ? acc * inv : 0.f;
    }
    uint4 hi[2], lo[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&hi[hf]);
      __nv_bfloat162* lb = reinterpret_cast<__nv_bfloat162*>(&lo[hf]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = out[8 * hf + 2 * e], c2 = out[8 * hf + 2 * e + 1];
        hb[e] = __floats2bfloat162_rn(a, c2);
        const float2 fh = __bfloat1622float2(hb[e]);
        lb[e] = __floats2bfloat162_rn(a - fh.x, c2 - fh.y);
      }
    }
    uint4* dst = ext + g * ext_stride + ((int64_t)(kc * W + wo) * Hst + row) * 2;
    dst[swl] = hi[0];
    dst[swl ^ 1] = hi[1];
    if (split) {
      uint4* dl = ext + g * ext_stride + ((int64_t)((NKC + kc) * W + wo) * Hst + row) * 2;
      dl[swl] = lo[0];
      dl[swl ^ 1] = lo[1];
    }
  }
}

static int tc_launch_conv0_pack(int ph, int pw, const float* feat, const float* w0, uint4* ext, int64_t B_utt, int64_t groups,
                                int T, int F, int C, int NKC, int W, int Hs, const TcSweepPlan& f, int64_t ext_stride,
                                cudaStream_t st) {
  const int64_t blocks = groups * ceil_div(W * f.Hst, 128);
  KWS_REQUIRE(blocks < (int64_t)2147483647, "conv_0 pre-pass: batch too large for one launch");
#define KWS_C0P(PH, PW)                                                                                                  \
  conv0_pool_pack_kernel<PH, PW><<<(unsigned)blocks, 128, 0, st>>>(feat, w0, ext, B_utt, T, F, C, NKC, W, Hs, f.pack_pitch,    \
                                                                   f.pack_n, f.Hst, f.dmax, f.split ? 1 : 0, ext_stride)
  if (ph == 1 && pw == 1) KWS_C0P(1, 1);
  else if (ph == 2 && pw == 2) KWS_C0P(2, 2);
  else if (ph == 4 && pw == 3) KWS_C0P(4, 3);
  else { set_error("conv_0 pre-pass: pooling %dx%d is not instantiated", ph, pw); return KWS_ERR_UNSUPPORTED; }
#undef KWS_C0P
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

template <int NKC, bool DBG, bool K32, bool SPLIT, bool PACK>
static int tc_launch_sweep_v(const SwParams& prm, int grid, int smem, cudaStream_t st) {
  KWS_CUDA(cudaFuncSetAttribute(resnet_tc_sweep_kernel<NKC, DBG, K32, SPLIT, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  resnet_tc_sweep_kernel<NKC, DBG, K32, SPLIT, PACK><<<grid, sw_threads(NKC), smem, st>>>(prm);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

template <int NKC>
static int tc_launch_sweep(const SwParams& prm, bool split, int grid, int smem, cudaStream_t st) {
  if (prm.ext_in != nullptr)   // packed strips (always the 32-byte-row layout; no cycle-accounting build)
    return split ? tc_launch_sweep_v<NKC, false, true, true, true>(prm, grid, smem, st)
                 : tc_launch_sweep_v<NKC, false, true, false, true>(prm, grid, smem, st);
  if (split) return tc_launch_sweep_v<NKC, false, true, true, false>(prm, grid, smem, st);
  if (prm.debug != nullptr)
    return prm.k32 ? tc_launch_sweep_v<NKC, true, true, false, false>(prm, grid, smem, st) : tc_launch_sweep_v<NKC, true, false, false, false>(prm, grid, smem, st);
  return prm.k32 ? tc_launch_sweep_v<NKC, false, true, false, false>(prm, grid, smem, st) : tc_launch_sweep_v<NKC, false, false, false, false>(prm, grid, smem, st);
}

static int tc_sweep_launch(TcResNet* p, const TcSweepPlan& f, const float* feat, int64_t B, int64_t B_utt, int T, int F, int H,
                           int W, float* logits, void* ws, size_t buf, const uint4* ext, int64_t ext_stride, int sub_h,
                           LaunchProfiler* prof, cudaStream_t st);

static int tc_sweep_forward(TcResNet* p, const TcSweepPlan& f, const float* feat, int64_t B, int T, int F, int H, int W,
                            float* logits, void* ws, int chunk_cfg, LaunchProfiler* prof, cudaStream_t st) {
  const kws_resnet_config& c = p->cfg;
  size_t buf = 0, ext_group = 0;
  tc_sweep_ws_bytes(p, f, H, W, B, chunk_cfg, &buf, &ext_group);
  const int n = c.n_layers;
  if (f.packed) {
    // packed strips: conv_0 + pool by the pre-pass kernel, then the sweep over groups of pack_n stacked utterances,
    // one launch pair per sub-batch.  The rows between stacked utterances are never stored: zero both buffers once.
    KWS_CUDA(cudaMemsetAsync(ws, 0, 2 * buf, st));
    const int64_t chunk = tc_pack_chunk(f, B, chunk_cfg);
    const int ph = c.pool_h > 0 ? c.pool_h : 1, pw = c.pool_w > 0 ? c.pool_w : 1;
    uint4* ext = reinterpret_cast<uint4*>(static_cast<char*>(ws) + 2 * buf);
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
      const int64_t nb = std::min(chunk, B - b0);
      const int64_t groups = ceil_div<int64_t>(nb, f.pack_n);
      if (prof) prof->tick(1, st);
      KWS_TRY(tc_launch_conv0_pack(ph, pw, feat + b0 * (int64_t)T * F, p->conv0_w, ext, nb, groups, T, F, c.n_maps, p->NKC, W,
                                   H, f, (int64_t)(ext_group / 16), st));
      KWS_TRY(tc_sweep_launch(p, f, nullptr, groups, nb, T, F, f.Hst, W, logits + b0 * c.n_labels, ws, buf, ext,
                              (int64_t)(ext_group / 16), H, prof, st));
    }
    return KWS_OK;
  }
  return tc_sweep_launch(p, f, feat, B, B, T, F, H, W, logits, ws, buf, nullptr, 0, H, prof, st);
}

// one launch of the sweep kernel: B units (utterances, or groups of stacked utterances when ext != nullptr)
static int tc_sweep_launch(TcResNet* p, const TcSweepPlan& f, const float* feat, int64_t B, int64_t B_utt, int T, int F, int H,
                           int W, float* logits, void* ws, size_t buf, const uint4* ext, int64_t ext_stride, int sub_h,
                           LaunchProfiler* prof, cudaStream_t st) {
  const kws_resnet_config& c = p->cfg;
  const int n = c.n_layers;
  // (no device-side tables: every per-layer quantity is derived from these parameters inside the kernel, so a change
  // of batch size, shape or workspace costs nothing)
  SwParams prm{};
  prm.wpack0 = reinterpret_cast<const unsigned char*>(p->wpack_sw[0]);
  prm.kconst0 = reinterpret_cast<const unsigned char*>(p->shift_p[0]);
  prm.layer_stride = n > 1 ? (int64_t)(reinterpret_cast<const unsigned char*>(p->wpack_sw[1]) - prm.wpack0) : 0;
  prm.use_dilation = c.use_dilation ? 1 : 0;
  prm.conv0_wb = reinterpret_cast<const unsigned char*>(f.split ? p->conv0_wb3 : p->conv0_wb);
  prm.last_scale = p->scale_p[n - 1];
  prm.out_w = p->out_w;
  prm.out_b = p->out_b;
  prm.P = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws));
  prm.Q = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + buf);
  prm.n_layers = n; prm.C = c.n_maps; prm.n_labels = c.n_labels; prm.T = T; prm.F = F;
  prm.H = H; prm.W = W; prm.n_strips = f.n_strips;
  prm.smem_c0w_off = f.c0w_off;
  prm.smem_skip_off = f.skip_off;
  prm.dmax = f.dmax;
  prm.chunk_rows = f.chunk_rows;
  prm.ring_slack_bytes = f.slack_bytes;
  prm.k32 = f.k32 ? 1 : 0;
  prm.discard_q = p->sweep_discard && f.n_strips == 1 && !f.split && !f.packed;   // (packed: the zero rows inside a column must survive)
  prm.ext_in = ext; prm.ext_stride = ext_stride;
  prm.pack_n = f.packed ? f.pack_n : 1; prm.pack_h = sub_h; prm.pack_pitch = f.packed ? f.pack_pitch : H;
  prm.smem_pool_off = f.pool_off;
  prm.B_utt = B_utt;
  prm.smem_w_off[0] = f.w_off[0]; prm.smem_w_off[1] = f.w_off[1];
  prm.smem_ring_off = f.ring_off; prm.ring_slot_bytes = f.slot_bytes; prm.n_stages = f.n_stages;
  prm.l2_policy = p->l2_policy;
  prm.wait_polls = p->wait_polls;
  prm.feat = feat;
  prm.logits = logits;
  prm.B = B;
  if (prof) prof->tick(0, st);   // the whole network is one launch: it IS the dominant kernel
  const int grid = (int)std::min<int64_t>(f.n_slots, B);
  static const bool dbg_on = [] { const char* e = std::getenv("HONK2_TC_DEBUG"); return e && std::atoi(e) != 0; }();
  static long long* dbg_buf = nullptr;
  if (dbg_on && !f.split && !f.packed) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(long long));
    prm.debug = dbg_buf;
    const char* e = std::getenv("HONK2_TC_DIAG");
    prm.diag = e ? std::atoi(e) : 0;
  }
  static const bool trace_on = [] { const char* e = std::getenv("HONK2_TC_TRACE"); return e && std::atoi(e) != 0; }();
  static long long* trace_buf = nullptr;
  if (prm.debug != nullptr && trace_on) {
    if (!trace_buf) { cudaMalloc(&trace_buf, 8 * kSwTraceLen * sizeof(long long)); }
    cudaMemsetAsync(trace_buf, 0, 8 * kSwTraceLen * sizeof(long long), st);
    prm.trace = trace_buf;
  }
  struct DbgPrint {
    long long* buf; cudaStream_t st; long long* trace;
    ~DbgPrint() {
      if (!buf) return;
      long long h[16];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
      if (trace) {   // event timestamps of CTA 0 -> HONK2_TC_TRACE_FILE (one row per event kind)
        std::vector<long long> t(8 * kSwTraceLen);
        cudaMemcpy(t.data(), trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const char* fn = std::getenv("HONK2_TC_TRACE_FILE");
        if (FILE* f = fopen(fn ? fn : "sweep_trace.csv", "w")) {
          fprintf(f, "idx,producer_copy,issuer_start,issuer_after_tempty,issuer_after_full,issuer_after_issue,epi0_tfull,epi0_done,epi7_done\n");
          for (int i = 0; i < kSwTraceLen; ++i) {
            fprintf(f, "%d", i);
            for (int r = 0; r < 8; ++r) fprintf(f, ",%lld", t[(size_t)r * kSwTraceLen + i]);
            fprintf(f, "\n");
          }
          fclose(f);
        }
      }
      const double ptot = (double)(h[13] + h[14] + h[15] + h[7]);
      fprintf(stderr, "[sweep dbg] producer of CTA 0: waiting for the previous layer's column %.1f%%, for a free stage %.1f%%, building "
              "conv_0 columns %.1f%%, issuing copies / weights / bookkeeping %.1f%% (total %.0f)\n",
              100 * h[13] / ptot, 100 * h[14] / ptot, 100 * h[15] / ptot, 100 * h[7] / ptot, ptot);
      h[7] = 0;
      const double tot = (double)(h[0] + h[1] + h[2] + h[3] + h[5] + h[6] + h[7]);
      fprintf(stderr, "[sweep dbg] issuer 0 of CTA 0, %lld utterances, cycles: weights wait %.1f%%, accumulator-free wait %.1f%%, "
              "TMA-data wait %.1f%%, utterance boundary (tail + conv_0 + first load) %.1f%%, issuing own steps %.1f%%, first loop top after an own step %.1f%%, counting along %.1f%% "
              "(total %.0f, %.0f per utterance)\n",
              h[4], 100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[5] / tot, 100 * h[3] / tot, 100 * h[7] / tot, 100 * h[6] / tot, tot,
              tot / (double)std::max(1ll, h[4]));
      const double et = (double)(h[8] + h[9] + h[10] + h[11] + h[12]);
      fprintf(stderr, "[sweep dbg] epilogue warp 0 of CTA 0: waiting for accumulators %.1f%%, TMEM load + re-zero %.1f%%, publishing the "
              "previous column %.1f%%, math + stores (+ guard, tail) %.1f%%, conv_0 %.1f%% (total %.0f)\n",
              100 * h[8] / et, 100 * h[9] / et, 100 * h[10] / et, 100 * h[11] / et, 100 * h[12] / et, et);
    }
  } dbg_print{prm.debug, st, prm.trace};
  switch (p->NKC) {
    case 1: return tc_launch_sweep<1>(prm, f.split, grid, f.smem_total, st);
    case 2: return tc_launch_sweep<2>(prm, f.split, grid, f.smem_total, st);
    default: return tc_launch_sweep<3>(prm, f.split, grid, f.smem_total, st);
  }
}

size_t tc_resnet_workspace_bytes(const TcResNet* p, int64_t B, int T, int F, int chunk, bool split) {
  if (!p || !p->supported) return 0;
  int H, W;
  tc_map_hw(p->cfg, T, F, &H, &W);
  if (H < 1 || W < 1 || W > 256 || !tc_layers_ok(p, H, W)) return 0;
  const TcSweepPlan sp = tc_sweep_plan(p, H, W, split);
  const TcFusedPlan f = tc_fused_plan(p, H, W, split);
  if (split) {   // the split-bf16 mode exists in the two whole-network kernels only
    if (sp.ok) return tc_sweep_ws_bytes(p, sp, H, W, B, chunk, nullptr);
    return f.ok ? tc_fused_ws_bytes(p, f, W, nullptr) : 0;
  }
  const int64_t c = tc_chunk(p, B, H, W, chunk);
  const size_t layered = p->lanes * tc_lane_bytes(p, c, H, W, nullptr);
  // the layer-per-launch path stays available (profiling, shapes the fused kernel cannot stage)
  size_t need = f.ok ? std::max(layered, tc_fused_ws_bytes(p, f, W, nullptr)) : layered;
  if (sp.ok) need = std::max(need, tc_sweep_ws_bytes(p, sp, H, W, B, chunk, nullptr));
  return need;
}

const char* tc_resnet_kernel_path(const TcResNet* p, int T, int F, bool split) {
  if (!p || !p->supported) return "unsupported";
  int H, W;
  tc_map_hw(p->cfg, T, F, &H, &W);
  if (H < 1 || W < 1 || W > 256 || !tc_layers_ok(p, H, W)) return "unsupported";
  if (tc_sweep_plan(p, H, W, split).ok) return "resnet_tc_sweep_kernel";
  if (tc_fused_plan(p, H, W, split).ok) return "resnet_tc_fused_kernel";
  return split ? "unsupported" : "conv3x3_tc_kernel";
}

template <int NKC, bool HAS_PREV, bool DO_POOL>
static int tc_launch_conv3(const CUtensorMap& map, const TcConvParams& prm, int grid, cudaStream_t st) {
  KWS_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<NKC, HAS_PREV, DO_POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                prm.g.smem_total));
  static const bool use_pdl = [] { const char* e = std::getenv("HONK2_TC_PDL"); return e == nullptr || std::atoi(e) != 0; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc_threads(NKC));
  cfg.dynamicSmemBytes = prm.g.smem_total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  KWS_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NKC, HAS_PREV, DO_POOL>, map, prm));
  ++g_launches;
  return KWS_OK;
}

template <int NKC>
static int tc_launch_conv(const CUtensorMap& map, const TcConvParams& prm, int grid, cudaStream_t st) {
  const bool prev = prm.skip != nullptr, pool = prm.pool_sum != nullptr;
  if (prev && pool) return tc_launch_conv3<NKC, true, true>(map, prm, grid, st);
  if (prev) return tc_launch_conv3<NKC, true, false>(map, prm, grid, st);
  if (pool) return tc_launch_conv3<NKC, false, true>(map, prm, grid, st);
  return tc_launch_conv3<NKC, false, false>(map, prm, grid, st);
}

int tc_resnet_forward(TcResNet* p, const float* feat, int64_t B, int T, int F, float* logits, void* ws,
                      size_t ws_bytes, int chunk_cfg, bool split, LaunchProfiler* prof, cudaStream_t st) {
  if (!p->supported) {
    set_error("bf16 tensor-core path supports 1..64 feature maps (got %d) and needs cuTensorMapEncodeTiled",
              p->cfg.n_maps);
    return KWS_ERR_UNSUPPORTED;
  }
  const kws_resnet_config& c = p->cfg;
  int H, W;
  tc_map_hw(c, T, F, &H, &W);
  KWS_REQUIRE(H >= 1 && W >= 1, "ResNet: input %dx%d is smaller than the pooling window", T, F);
  const size_t need = tc_resnet_workspace_bytes(p, B, T, F, chunk_cfg, split);
  if (need == 0) {
    set_error("%s tensor-core path cannot tile a %dx%d map", split ? "split-bf16" : "bf16", H, W);
    return KWS_ERR_UNSUPPORTED;
  }
  if (ws == nullptr || ws_bytes < need) {
    set_error("ResNet bf16 forward needs %zu bytes of workspace, got %zu", need, ws_bytes);
    return KWS_ERR_WORKSPACE;
  }
  {
    // whole-network persistent kernel unless per-launch profiling was requested
    static const bool prof_layered = [] { const char* e = std::getenv("HONK2_TC_PROFILE_LAYERED"); return e && std::atoi(e) != 0; }();
    const bool layered = prof && prof->enabled && prof_layered && !split;
    const TcSweepPlan sp = tc_sweep_plan(p, H, W, split);
    if (sp.ok && !layered) {
      return tc_sweep_forward(p, sp, feat, B, T, F, H, W, logits, ws, chunk_cfg, prof, st);
    }
    const TcFusedPlan f = tc_fused_plan(p, H, W, split);
    if (f.ok && !layered) {
      return tc_fused_forward(p, f, feat, B, T, F, H, W, logits, ws, prof, st);
    }
    if (split) {
      set_error("split-bf16 tensor-core path: no whole-network kernel can stage a %dx%d map", H, W);
      return KWS_ERR_UNSUPPORTED;
    }
  }
  const int64_t chunk = tc_chunk(p, B, H, W, chunk_cfg);
  const int Hpad = tc_hpad(c, H);
  size_t buf = 0;
  const size_t lane_bytes = tc_lane_bytes(p, chunk, H, W, &buf);
  // per-launch profiling needs one ordered stream; a single chunk has nothing to overlap with
  const int lanes = (prof && prof->enabled) || B <= chunk ? 1 : p->lanes;
  cudaStream_t caller = st;
  if (lanes > 1) {
    KWS_CUDA(cudaEventRecord(p->ev_start, caller));
    for (int l = 0; l < lanes; ++l) KWS_CUDA(cudaStreamWaitEvent(p->lane_stream[l], p->ev_start, 0));
  }
  const bool fuse_pool = c.n_layers >= 1;
  const int ph = c.pool_h > 0 ? c.pool_h : 1, pw = c.pool_w > 0 ? c.pool_w : 1;
  KWS_REQUIRE(W <= 256, "conv_0: map width %d exceeds 256", W);
  const bool fast0 = (ph == 1 && pw == 1);
  const int groups0 = ceil_div(W, kC0Px);
  const int rows0f = std::max(1, std::min(H, 256 / groups0));
  const size_t smem0f = sizeof(float) * (round_up((rows0f + 2) * (groups0 * kC0Px + 2), 4) + p->NP * 8 * 12);
  const int rows0 = std::max(1, std::min(H, 256 / W));
  const size_t smem0 = sizeof(float) * (round_up((rows0 * ph + 2) * (F + 2), 4) + p->NP * 8 * 12);
  KWS_REQUIRE(smem0 <= 48 * 1024, "conv_0: tile needs %zu bytes of shared memory", smem0);

  int64_t chunk_idx = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, ++chunk_idx) {
    const int64_t nb = std::min(chunk, B - b0);
    const int lane = (int)(chunk_idx % lanes);
    if (lanes > 1) st = p->lane_stream[lane];
    char* lws = static_cast<char*>(ws) + (size_t)lane * lane_bytes;
    // P: conv_0 output and the even layers' activations (updated in place, it doubles as the skip tensor);
    // Q: the odd layers' activations
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(lws);
    __nv_bfloat16* Q = reinterpret_cast<__nv_bfloat16*>(lws + buf);
    float* pool = reinterpret_cast<float*>(lws + 2 * buf);
    if (Hpad > H && chunk_idx < lanes) {   // first use of this lane's buffers in this call
      for (__nv_bfloat16* bp : {P, Q}) {
        zero_pad_rows_kernel<<<256, 256, 0, st>>>(reinterpret_cast<uint4*>(bp), chunk * p->NP, H, Hpad, W);
        KWS_CHECK_LAUNCH();
      }
    }
    if (prof) prof->tick(1, st);
    if (fast0 && smem0f <= 48 * 1024)
      conv0_p8_w4_kernel<<<dim3(ceil_div(H, rows0f), (unsigned)nb), 256, smem0f, st>>>(
          feat + b0 * (int64_t)T * F, p->conv0_w, P, fuse_pool ? pool : nullptr, T, F, c.n_maps, p->NP, groups0,
          rows0f, Hpad);
    else
      conv0_p8_kernel<<<dim3(ceil_div(H, rows0), (unsigned)nb), 256, smem0, st>>>(
          feat + b0 * (int64_t)T * F, p->conv0_w, P, fuse_pool ? pool : nullptr, T, F, c.n_maps, p->NP, ph, pw, H,
          W, rows0, Hpad);
    KWS_CHECK_LAUNCH();
    for (int i = 1; i <= c.n_layers; ++i) {
      TcConvParams prm;
      const int d = c.use_dilation ? (1 << ((i - 1) / 3)) : 1;
      KWS_REQUIRE(tc_geom(p->NKC, H, Hpad, W, d, &prm.g), "bf16 conv: cannot tile H=%d W=%d d=%d", H, W, d);
      const bool even = (i % 2 == 0), last = (i == c.n_layers);
      const __nv_bfloat16* x = even ? Q : P;     // input: output of the previous layer
      CUtensorMap* map = nullptr;
      KWS_TRY(tc_get_map(p, x, nb * p->NP, H, W, prm.g, &map));
      prm.wpack = p->wpack[i - 1];
      prm.kconst = p->shift_p[i - 1];
      prm.skip = even ? P : nullptr;             // resnet.py:51-53, reconstructed from z_{i-2}
      prm.y = last ? nullptr : (even ? P : Q);   // even layers overwrite their skip tensor element-wise
      prm.pool_sum = last ? pool : nullptr;
      prm.B = (int)nb;
      prm.total_tiles = (int)nb * prm.g.tiles_per_utt;
      const int grid = std::min(p->n_sms, prm.total_tiles);
      if (prof) prof->tick(0, st);
      switch (p->NKC) {
        case 1: KWS_TRY(tc_launch_conv<1>(*map, prm, grid, st)); break;
        case 2: KWS_TRY(tc_launch_conv<2>(*map, prm, grid, st)); break;
        case 3: KWS_TRY(tc_launch_conv<3>(*map, prm, grid, st)); break;
        default: KWS_TRY(tc_launch_conv<4>(*map, prm, grid, st)); break;
      }
    }
    if (prof) prof->tick(1, st);
    if (fuse_pool)
      tail_pool_kernel<<<(unsigned)ceil_div<int64_t>(nb * c.n_labels, 256), 256, 0, st>>>(
          pool, p->scale_p[c.n_layers - 1], p->out_w, p->out_b, logits + b0 * c.n_labels, nb, c.n_maps, p->CP, H * W,
          c.n_labels);
    else
      tail_p8_kernel<<<(unsigned)nb, 256, 0, st>>>(P, p->out_w, p->out_b, logits + b0 * c.n_labels, c.n_maps, p->NP,
                                                   H * W, c.n_labels);   // n_layers == 0: Hpad == H
    KWS_CHECK_LAUNCH();
  }
  if (lanes > 1) {
    for (int l = 0; l < lanes; ++l) {
      KWS_CUDA(cudaEventRecord(p->ev_done[l], p->lane_stream[l]));
      KWS_CUDA(cudaStreamWaitEvent(caller, p->ev_done[l], 0));
    }
  }
  return KWS_OK;
}

}  // namespace kws
