// fp32 CUDA-core kernels of the ResNet path: the reference arithmetic (parity mode).  Activations are planar
// [B][C][H][W] float32, the same order torch uses, so every intermediate can be compared with the reference module.
//
//   conv0_f32_kernel         conv_0 + ReLU + AvgPool          /root/reference/model/resnet.py:40-44
//   conv3x3_f32_row_kernel   conv_i + ReLU + skip + BatchNorm /root/reference/model/resnet.py:48-55: persistent, the layer's weights
//                            resident in shared memory, 1 row x 8 columns x Q maps per thread (maps whose width is a multiple of 8)
//   conv3x3_f32_res_kernel   the same, 8 rows x 1 column x Q maps per thread (the pooled maps: 13 and 20 columns)
//   conv3x3_f32_kernel       the same, one 8-row tile per CTA (weights that do not fit in shared memory; HONK2_F32_RESIDENT=0)
//   tail_f32_kernel          mean over H*W + Linear           /root/reference/model/resnet.py:57-59
#include <algorithm>
#include <cstdlib>
#include "kernels.cuh"

namespace kws {

// =============================================================================================
// conv_0: one thread per (pooled) output pixel, loops over the C output maps.
// PH/PWD > 0: pooling window known at compile time, the (PH+2)x(PWD+2) input patch lives in
// registers.  PH == 0: runtime window, patch read from shared memory.
constexpr int kConv0Threads = 256;

template <int PH, int PWD>
__global__ void __launch_bounds__(kConv0Threads)
conv0_f32_kernel(const float* __restrict__ feat, const float* __restrict__ w0, float* __restrict__ out,
                 int T, int F, int C, int ph_rt, int pw_rt, int Ho, int Wo, int rows_per_tile) {
  extern __shared__ __align__(16) float smem[];
  const int ph = PH > 0 ? PH : ph_rt, pw = PWD > 0 ? PWD : pw_rt;
  const int in_rows = rows_per_tile * ph + 2;
  const int in_cols = F + 2;
  float* s_in = smem;                                   // [in_rows][in_cols]
  float* s_w = smem + round_up(in_rows * in_cols, 4);   // [C][12]

  const int64_t b = blockIdx.y;
  const int ho0 = blockIdx.x * rows_per_tile;
  const int h_in0 = ho0 * ph - 1;
  const float* src = feat + b * (int64_t)T * F;
  for (int i = threadIdx.x; i < in_rows * in_cols; i += kConv0Threads) {
    const int r = i / in_cols, c = i - r * in_cols;
    const int h = h_in0 + r, w = c - 1;
    s_in[i] = (h >= 0 && h < T && w >= 0 && w < F) ? __ldg(src + (int64_t)h * F + w) : 0.f;
  }
  for (int i = threadIdx.x; i < C * 12; i += kConv0Threads) {
    const int c = i / 12, k = i - c * 12;
    s_w[i] = k < 9 ? __ldg(w0 + c * 9 + k) : 0.f;
  }
  __syncthreads();

  const int r = threadIdx.x / Wo, wo = threadIdx.x - r * Wo;
  const int ho = ho0 + r;
  if (r >= rows_per_tile || ho >= Ho) return;
  const float inv = 1.f / (float)(ph * pw);
  float* dst = out + ((b * C) * (int64_t)Ho + ho) * Wo + wo;
  const int64_t cstride = (int64_t)Ho * Wo;

  if constexpr (PH > 0) {
    float p[PH + 2][PWD + 2];
#pragma unroll
    for (int i = 0; i < PH + 2; ++i)
#pragma unroll
      for (int j = 0; j < PWD + 2; ++j) p[i][j] = s_in[(r * PH + i) * in_cols + wo * PWD + j];
    for (int c = 0; c < C; ++c) {
      const float4 wa = *reinterpret_cast<const float4*>(s_w + c * 12);
      const float4 wb = *reinterpret_cast<const float4*>(s_w + c * 12 + 4);
      const float w8 = s_w[c * 12 + 8];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < PH; ++i)
#pragma unroll
        for (int j = 0; j < PWD; ++j) {
          float v = p[i][j] * wa.x;
          v = fmaf(p[i][j + 1], wa.y, v);
          v = fmaf(p[i][j + 2], wa.z, v);
          v = fmaf(p[i + 1][j], wa.w, v);
          v = fmaf(p[i + 1][j + 1], wb.x, v);
          v = fmaf(p[i + 1][j + 2], wb.y, v);
          v = fmaf(p[i + 2][j], wb.z, v);
          v = fmaf(p[i + 2][j + 1], wb.w, v);
          v = fmaf(p[i + 2][j + 2], w8, v);
          sum += fmaxf(v, 0.f);
        }
      dst[c * cstride] = (PH * PWD == 1) ? sum : sum * inv;
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const float* wc = s_w + c * 12;
      float sum = 0.f;
      for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j) {
          const float* q = s_in + (r * ph + i) * in_cols + wo * pw + j;
          float v = 0.f;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int e = 0; e < 3; ++e) v = fmaf(q[a * in_cols + e], wc[a * 3 + e], v);
          sum += fmaxf(v, 0.f);
        }
      dst[c * cstride] = sum * inv;
    }
  }
}

int launch_conv0_f32(const float* feat, const float* w0, float* out, int64_t B, int T, int F, int C,
                     int ph, int pw, cudaStream_t st) {
  KWS_REQUIRE(ph >= 1 && pw >= 1, "conv_0: bad pool %dx%d", ph, pw);
  const int Ho = T / ph, Wo = F / pw;
  KWS_REQUIRE(Ho >= 1 && Wo >= 1, "conv_0: input %dx%d smaller than pool %dx%d", T, F, ph, pw);
  KWS_REQUIRE(Wo <= kConv0Threads, "conv_0: pooled width %d exceeds %d", Wo, kConv0Threads);
  KWS_REQUIRE(B <= 65535, "conv_0: chunk too large");
  const int rows = max(1, min(Ho, kConv0Threads / Wo));
  const int tiles = ceil_div(Ho, rows);
  const size_t smem = sizeof(float) * (round_up((rows * ph + 2) * (F + 2), 4) + C * 12);
  KWS_REQUIRE(smem <= 200 * 1024, "conv_0: tile needs %zu bytes of shared memory", smem);
  dim3 grid(tiles, (unsigned)B);
#define KWS_LAUNCH_CONV0(PH_, PW_)                                                                \
  do {                                                                                            \
    if (smem > 48 * 1024)                                                                         \
      KWS_CUDA(cudaFuncSetAttribute(conv0_f32_kernel<PH_, PW_>,                                   \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    conv0_f32_kernel<PH_, PW_><<<grid, kConv0Threads, smem, st>>>(feat, w0, out, T, F, C, ph, pw, \
                                                                  Ho, Wo, rows);                  \
  } while (0)
  if (ph == 1 && pw == 1) KWS_LAUNCH_CONV0(1, 1);
  else if (ph == 4 && pw == 3) KWS_LAUNCH_CONV0(4, 3);
  else if (ph == 2 && pw == 2) KWS_LAUNCH_CONV0(2, 2);
  else KWS_LAUNCH_CONV0(0, 0);
#undef KWS_LAUNCH_CONV0
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

// =============================================================================================
// conv_i: direct 3x3 dilated convolution, register blocked.
//   CTA    = one utterance x RG*8 output rows x all W columns x all C output maps
//   thread = (cout group cg of Q maps, row group rg of 8 rows, column w): 8 x Q accumulators
//   smem   = for CK input maps at a time: the three row sets {h-d, h, h+d} of the tile, each row
//            zero padded by min(d, W-ish) columns on both sides; and the CK x 9 x C weight slab.
// Lanes run along w, so every shared-memory read of the input is conflict free and every
// global store is a contiguous row segment.
constexpr int kPH = 8;     // output rows per thread
// weight pitch per map group: Q rounded up to whole float4s
__host__ __device__ constexpr int qp_of(int Q) { return (Q + 3) & ~3; }

struct Conv3x3Geom {
  int Q, CG, RG, CK, wpad, row_stride, threads, tiles_h;
  size_t smem;
};

// Resident-weight kernels: 4 warps on one of the SM's four sub-partitions (16 K registers each) leave 128 registers per
// thread, 3 warps 168; the narrow thread tiles (Q <= 9: at most 72 accumulators) fit the former.
__host__ __device__ constexpr int res_max_threads(int Q) { return Q <= 9 ? 512 : 384; }

int conv3x3_f32_q(int C) {
  // Output maps per thread.  Score = (real maps / padded maps) x (how evenly the warps of a persistent CTA of the
  // resident-weight row kernel fill the SM's four schedulers on a 40-column map: 8 rows x 5 column blocks x CG threads
  // per unit).  45 maps: Q = 8 (48 padded maps, two 240-thread units = 15 warps) beats Q = 9 (45 maps, two 200-thread
  // units = 12.5 warps: one scheduler carries 4 warps, the others 3, and every barrier waits for it).
  int best = 9;
  double best_score = -1.0;
  for (int q = 12; q >= 8; --q) {
    const int cg = ceil_div(C, q);
    const int unit = cg * 40;
    int ns = res_max_threads(q) / unit;
    if (ns < 1) ns = 1;
    if (ns > 8) ns = 8;
    const int warps = ceil_div(ns * unit, 32);
    const double util = (double)(ns * unit) / (double)(ceil_div(warps, 4) * 128);
    const double score = (double)C / (cg * q) * util;
    if (score > best_score + 1e-9) { best_score = score; best = q; }
  }
  return best;
}

static bool conv3x3_geom(int C, int H, int W, int d, Conv3x3Geom* g) {
  g->Q = conv3x3_f32_q(C);
  g->CG = ceil_div(C, g->Q);
  if (W * g->CG > 256) return false;
  const int max_rg = max(1, 256 / (W * g->CG));
  const int need_rg = ceil_div(H, kPH);
  g->RG = min(max_rg, need_rg);
  // spread rows evenly over the tiles that are needed anyway
  g->tiles_h = ceil_div(need_rg, g->RG);
  g->RG = ceil_div(need_rg, g->tiles_h);
  g->wpad = d < W ? d : 0;
  g->row_stride = W + 2 * g->wpad;
  g->threads = round_up(W * g->CG * g->RG, 32);
  const size_t per_ch = sizeof(float) * (3 * g->RG * kPH * g->row_stride + 9 * g->CG * qp_of(g->Q));
  const size_t budget = 46 * 1024;
  int ck = (int)(budget / per_ch);
  if (ck < 1) ck = 1;
  if (ck > C) ck = C;
  // prefer an even split of C
  const int chunks = ceil_div(C, ck);
  g->CK = ceil_div(C, chunks);
  g->smem = per_ch * g->CK;
  return g->smem <= 200 * 1024;
}

template <int Q>
__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(Conv3x3F32 a, Conv3x3Geom g) {
  extern __shared__ __align__(16) float smem[];
  const int Hr = g.RG * kPH;
  float* s_in = smem;                                   // [CK][3][Hr][row_stride]
  float* s_w = smem + g.CK * 3 * Hr * g.row_stride;     // [CK][9][CG*QP]
  const int C = a.C, H = a.H, W = a.W, d = a.d;
  const int64_t b = blockIdx.y;
  const int h0 = blockIdx.x * Hr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  // thread -> (cg, rg, w)
  const int per_cg = g.RG * W;
  const int cg = tid / per_cg;
  const int rem = tid - cg * per_cg;
  const int rg = rem / W;
  const int w = rem - rg * W;
  const bool active = cg < g.CG;

  float acc[kPH][Q];
#pragma unroll
  for (int j = 0; j < kPH; ++j)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[j][q] = 0.f;

  const float* xb = a.x + b * (int64_t)C * H * W;
  constexpr int QP = qp_of(Q);
  const int w_slab = 9 * g.CG * QP;
  const bool side_taps = g.wpad > 0;  // d >= W: the dw != 0 taps only ever read padding

  for (int c0 = 0; c0 < C; c0 += g.CK) {
    // ---- stage inputs: rows (ci, dh, r)
    const int n_rows = g.CK * 3 * Hr;
    for (int row = warp; row < n_rows; row += nwarps) {
      const int ci = row / (3 * Hr);
      const int r2 = row - ci * 3 * Hr;
      const int dh = r2 / Hr;
      const int r = r2 - dh * Hr;
      const int h = h0 + r + (dh - 1) * d;
      const bool ok = (c0 + ci < C) && h >= 0 && h < H;
      const float* srow = xb + ((int64_t)(c0 + ci) * H + (ok ? h : 0)) * W;
      float* drow = s_in + row * g.row_stride;
      for (int col = lane; col < g.row_stride; col += 32) {
        const int ww = col - g.wpad;
        drow[col] = (ok && ww >= 0 && ww < W) ? __ldg(srow + ww) : 0.f;
      }
    }
    // ---- stage weights: contiguous slab [CK][9][CG*QP]
    {
      const int n4 = g.CK * w_slab / 4;
      const int valid4 = max(0, min(g.CK, C - c0)) * w_slab / 4;
      const float4* src = reinterpret_cast<const float4*>(a.wt + (int64_t)c0 * w_slab);
      float4* dst = reinterpret_cast<float4*>(s_w);
      for (int i = tid; i < n4; i += blockDim.x)
        dst[i] = i < valid4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    if (active) {
      for (int ci = 0; ci < g.CK; ++ci) {
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const float* in_row = s_in + ((ci * 3 + dh) * Hr + rg * kPH) * g.row_stride + g.wpad + w;
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            if (dw != 1 && !side_taps) continue;
            const float* ip = in_row + (dw - 1) * d;
            float xv[kPH];
#pragma unroll
            for (int j = 0; j < kPH; ++j) xv[j] = ip[j * g.row_stride];
            const float* wp = s_w + ((ci * 9 + dh * 3 + dw) * g.CG + cg) * QP;
            float wv[QP];
#pragma unroll
            for (int q4 = 0; q4 < QP / 4; ++q4) {
              const float4 t = *reinterpret_cast<const float4*>(wp + 4 * q4);
              wv[4 * q4] = t.x; wv[4 * q4 + 1] = t.y; wv[4 * q4 + 2] = t.z; wv[4 * q4 + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < kPH; ++j)
#pragma unroll
              for (int q = 0; q < Q; ++q) acc[j][q] = fmaf(xv[j], wv[q], acc[j][q]);
          }
        }
      }
    }
    __syncthreads();
  }

  if (!active) return;
  // ---- epilogue: ReLU, residual (pre-BN skip), BatchNorm (resnet.py:49-55)
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int co = cg * Q + q;
    if (co >= C) break;
    const float sc = __ldg(a.bn_scale + co), sh = __ldg(a.bn_shift + co);
#pragma unroll
    for (int j = 0; j < kPH; ++j) {
      const int h = h0 + rg * kPH + j;
      if (h >= H) break;
      const int64_t off = ((b * C + co) * (int64_t)H + h) * W + w;
      float v = fmaxf(acc[j][q], 0.f);
      if (a.prev_in) {
        v += a.prev_in[off];
        a.prev_out[off] = v;
      }
      a.y[off] = fmaf(v, sc, sh);
    }
  }
}


// =============================================================================================
// conv_i, resident-weight form (the default wherever it fits): same thread tile (8 rows x Q maps, lanes along w), but
//   * the CTA is persistent (one per SM) and keeps the WHOLE layer's packed weights in shared memory -- they are read
//     from global memory once per CTA instead of once per 8-row tile;
//   * a CTA works on NS independent 8-row "units" (utterance, row group) at a time -- units are numbered through the
//     whole sub-batch, so the last row group of one utterance and the first of the next share a CTA and no lane idles
//     on the H = 101 -> 13 x 8 split;
//   * the input rows of the next CK input maps arrive by cp.async (zero fill for rows / maps / units outside the
//     problem = the reference's zero padding, resnet.py:22-24) into the other half of a double buffer while the
//     FMAs of the current chunk run, continuously across units: the old kernel staged with dependent scalar loads
//     between two __syncthreads and spent most of its time waiting for them (ncu: FMA pipe 20 %).
// The left / right padding columns of every staged row are zeroed once and never written.

struct ConvResGeom {
  int Q, CG, NS, CK, n_chunks, wpad, row_stride, threads, sub_threads, units_per_utt, vec, seg_per_row;
  int w_floats, sub_floats, buf_floats;   // packed weights; one unit's chunk [CK][3][8][row_stride]; NS of them
  int64_t n_units, n_items;
  uint32_t m_seg, m_24, m_ck, m_upu;      // magic multipliers: n / d == __umulhi(n, m) for the ranges used here
  size_t smem;
};

static uint32_t magic_div(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t m) { return m == 0u ? n : __umulhi(n, m); }

static bool conv3x3_res_geom(int64_t B, int C, int H, int W, int d, ConvResGeom* g) {
  g->Q = conv3x3_f32_q(C);
  g->CG = ceil_div(C, g->Q);
  g->sub_threads = g->CG * W;
  const int max_threads = res_max_threads(g->Q);
  if (g->sub_threads > max_threads) return false;
  g->units_per_utt = ceil_div(H, kPH);
  g->n_units = B * (int64_t)g->units_per_utt;
  if (g->n_units >= (1ll << 24)) return false;          // (fast_div exactness: n * d < 2^32)
  g->wpad = d < W ? round_up(d, 4) : 0;
  if (W == 40 && d <= 16) g->wpad = 16;   // the 40-mel maps: one compile-time row pitch (72 floats) for every dilation
  g->row_stride = W + 2 * g->wpad;
  g->vec = (W % 4 == 0) ? 1 : 0;
  g->seg_per_row = g->vec ? W / 4 : W;
  g->w_floats = C * 9 * g->CG * qp_of(g->Q);
  const size_t cap = 226 * 1024;
  int ns = max_threads / g->sub_threads;
  if (ns > 8) ns = 8;
  if ((int64_t)ns > g->n_units) ns = (int)g->n_units;
  for (; ns >= 1; --ns) {
    for (int ck = 8; ck >= 2; --ck) {
      if (ck > C && ck > 2) continue;
      const size_t need = sizeof(float) * ((size_t)g->w_floats + 2ull * ns * ck * 3 * kPH * g->row_stride);
      if (need <= cap) {
        g->NS = ns;
        // even split of C over the chunks
        const int chunks = ceil_div(C, ck);
        g->CK = ceil_div(C, chunks);
        g->n_chunks = chunks;
        g->sub_floats = g->CK * 3 * kPH * g->row_stride;
        g->buf_floats = ns * g->sub_floats;
        g->smem = sizeof(float) * ((size_t)g->w_floats + 2ull * g->buf_floats);
        g->threads = round_up(ns * g->sub_threads, 32);
        g->n_items = ceil_div(g->n_units, (int64_t)ns);
        g->m_seg = magic_div((uint32_t)g->seg_per_row);
        g->m_24 = magic_div(3 * kPH);
        g->m_ck = magic_div((uint32_t)g->CK);
        g->m_upu = magic_div((uint32_t)g->units_per_utt);
        // fast_div is exact while n * d < 2^32: pieces per chunk and unit numbers are far below that
        return (int64_t)ns * g->CK * 3 * kPH * g->seg_per_row < (1 << 20);
      }
    }
  }
  return false;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// STRIDE > 0: the staged rows' pitch is known at compile time (row offsets become immediates of the loads; with a
// runtime pitch every shared-memory load of the inner loop carried its own address instruction)
// (13 warps = 4 on one of the SM's four sub-partitions of 16 K registers: 128 registers per thread; 12 warps: 168)
template <int Q, int STRIDE>
__global__ void __launch_bounds__(res_max_threads(Q), 1)
conv3x3_f32_res_kernel(const Conv3x3F32 a, const ConvResGeom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int QP = qp_of(Q);
  float* s_w = smem;                      // [C][9][CG*QP]
  float* s_in = smem + g.w_floats;        // [2][NS][CK][3 dh][8 rows][row_stride]
  const int C = a.C, H = a.H, W = a.W, d = a.d;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int stride = STRIDE > 0 ? STRIDE : g.row_stride;

  // ---- once per CTA: weights, zeroed staging buffers
  {
    const float4* src = reinterpret_cast<const float4*>(a.wt);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < g.w_floats / 4; i += nthr) dst[i] = __ldg(src + i);
    float4* z = reinterpret_cast<float4*>(s_in);
    for (int i = tid; i < g.buf_floats / 2; i += nthr) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  // ---- this thread's output tile: unit `sub` of the item, maps cg*Q .. cg*Q+Q-1, column w, 8 rows
  const int sub = tid / g.sub_threads;
  const int rem = tid - sub * g.sub_threads;
  const int cg = rem / W;
  const int w = rem - cg * W;
  const bool active = sub < g.NS;
  const bool side_taps = g.wpad > 0;   // d >= W: the dw != 0 taps only ever read padding

  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_my = first < g.n_items ? (g.n_items - first + step - 1) / step : 0;
  const int n_chunks = g.n_chunks;
  const int n_pieces = g.NS * g.CK * 3 * kPH * g.seg_per_row;
  const uint32_t s_in_u32 = (uint32_t)__cvta_generic_to_shared(s_in);

  // stage chunk `chunk` of item `item` into buffer `buf` (asynchronously)
  auto stage = [&](int64_t item, int chunk, int buf) {
    const int c0 = chunk * g.CK;
    const uint32_t unit0 = (uint32_t)(item * g.NS);
    for (int p = tid; p < n_pieces; p += nthr) {
      const uint32_t rowid = fast_div((uint32_t)p, g.m_seg);
      const uint32_t seg = (uint32_t)p - rowid * (uint32_t)g.seg_per_row;
      const uint32_t sc = fast_div(rowid, g.m_24);
      const uint32_t r24 = rowid - sc * (3u * kPH);
      const uint32_t su = fast_div(sc, g.m_ck);
      const uint32_t ci = sc - su * (uint32_t)g.CK;
      const uint32_t unit = unit0 + su;
      const uint32_t b = fast_div(unit, g.m_upu);
      const int t = (int)(unit - b * (uint32_t)g.units_per_utt);
      const int dh = (int)(r24 >> 3), r = (int)(r24 & 7u);
      const int h = t * kPH + r + (dh - 1) * d;
      const int c = c0 + (int)ci;
      const bool ok = (int64_t)unit < g.n_units && c < C && h >= 0 && h < H;
      const uint32_t dst = s_in_u32 + 4u * ((uint32_t)buf * (uint32_t)g.buf_floats + sc * (uint32_t)(3 * kPH * stride) +
                                            r24 * (uint32_t)stride + (uint32_t)g.wpad + seg * (g.vec ? 4u : 1u));
      const float* src = ok ? a.x + (((int64_t)b * C + c) * H + h) * (int64_t)W + seg * (g.vec ? 4 : 1) : a.x;
      if (g.vec) cp_async16(dst, src, ok ? 16u : 0u);
      else cp_async4(dst, src, ok ? 4u : 0u);
    }
    cp_async_commit();
  };

  float acc[kPH][Q];
#pragma unroll
  for (int j = 0; j < kPH; ++j)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[j][q] = 0.f;

  if (n_my > 0) stage(first, 0, 0);
  int64_t it_s = 0; int ch_s = 1;            // next (item index, chunk) to stage
  if (ch_s == n_chunks) { ch_s = 0; ++it_s; }
  int buf = 0;
  for (int64_t it = 0; it < n_my; ++it) {
    const int64_t item = first + it * step;
    for (int ch = 0; ch < n_chunks; ++ch, buf ^= 1) {
      if (it_s < n_my) {
        stage(first + it_s * step, ch_s, buf ^ 1);
        if (++ch_s == n_chunks) { ch_s = 0; ++it_s; }
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (active) {
        const float* in_base = s_in + buf * g.buf_floats + sub * g.sub_floats + g.wpad + w;
        const float* w_base = s_w + ((size_t)(ch * g.CK) * 9 * g.CG + cg) * QP;
        const int ck_n = min(g.CK, C - ch * g.CK);
        for (int ci = 0; ci < ck_n; ++ci) {
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            const float* in_row = in_base + ((ci * 3 + dh) * kPH) * stride;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
              if (dw != 1 && !side_taps) continue;
              const float* ip = in_row + (dw - 1) * d;
              float xv[kPH];
#pragma unroll
              for (int j = 0; j < kPH; ++j) xv[j] = ip[j * stride];
              const float* wp = w_base + (size_t)((ci * 9 + dh * 3 + dw) * g.CG) * QP;
              float wv[QP];
#pragma unroll
              for (int q4 = 0; q4 < QP / 4; ++q4) {
                const float4 t4 = *reinterpret_cast<const float4*>(wp + 4 * q4);
                wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
              }
#pragma unroll
              for (int j = 0; j < kPH; ++j)
#pragma unroll
                for (int q = 0; q < Q; ++q) acc[j][q] = fmaf(xv[j], wv[q], acc[j][q]);
            }
          }
        }
      }
      __syncthreads();   // buffer `buf` is refilled by the next iteration's stage()
    }
    // ---- epilogue of this item: ReLU, residual (pre-BN skip), BatchNorm (resnet.py:49-55)
    if (active) {
      const int64_t unit = item * g.NS + sub;
      if (unit < g.n_units) {
        const int64_t b = unit / g.units_per_utt;
        const int h0 = (int)(unit - b * g.units_per_utt) * kPH;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int co = cg * Q + q;
          if (co < C) {
            const float sc = __ldg(a.bn_scale + co), sh = __ldg(a.bn_shift + co);
            const int64_t off0 = ((b * C + co) * (int64_t)H + h0) * W + w;
            float v[kPH];
#pragma unroll
            for (int j = 0; j < kPH; ++j) v[j] = fmaxf(acc[j][q], 0.f);
            if (a.prev_in) {
#pragma unroll
              for (int j = 0; j < kPH; ++j) if (h0 + j < H) v[j] += a.prev_in[off0 + (int64_t)j * W];
#pragma unroll
              for (int j = 0; j < kPH; ++j) if (h0 + j < H) a.prev_out[off0 + (int64_t)j * W] = v[j];
            }
#pragma unroll
            for (int j = 0; j < kPH; ++j) if (h0 + j < H) a.y[off0 + (int64_t)j * W] = fmaf(v[j], sc, sh);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kPH; ++j)
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[j][q] = 0.f;
  }
}

// =============================================================================================
// conv_i, resident weights, ROW tile (maps whose width is a multiple of 8: the 40-mel maps of every unpooled net):
// thread = one output row x 8 consecutive columns x Q maps; a warp = 8 rows x 4 column blocks.  Everything else is the
// resident-weight kernel above (persistent CTA, weights read once, NS units per CTA, cp.async double buffer).  Why a
// second tile: with lanes along w, a tap costs 8 scalar shared-memory loads + 3 weight loads per 72 FMAs, and the
// shared-memory pipe (ncu: LSU wavefronts) runs at ~80 % of the FMA pipe's time -- the two cannot overlap perfectly, so
// the kernel stalled at 38 % of the FMA peak.  With 8 consecutive columns per thread the three width taps of a row
// come out of ONE 16-float register window (dilation 1, 2, 4: four 16-byte loads per input row instead of 24 scalar
// ones; dilation 8, 16: two 16-byte loads per tap), i.e. 13-15 load instructions per 216 FMAs instead of 33.
// The staged row pitch S is a multiple of 4 floats with S/4 odd: the 8 rows of a quarter warp then hit 8 different
// 16-byte bank groups.  Rows are staged as the union h0-d .. h0+7+d (d <= 8) or as three sets of 8 (d > 8).
struct ConvRowGeom {
  int Q, CG, NS, CK, n_chunks, pad, S, threads, sub_threads, units_per_utt, seg_per_row, WB;
  int rs, rows_st;                        // row step between height taps in the staged tile, staged rows per map
  int w_floats, sub_floats, buf_floats;
  int64_t n_units, n_items;
  uint32_t m_seg, m_rows, m_ck, m_upu;
  size_t smem;
};

static bool conv3x3_row_geom(int64_t B, int C, int H, int W, int d, ConvRowGeom* g) {
  if (W % 8 != 0) return false;
  if (!(d == 1 || d == 2 || d == 4 || d % 4 == 0)) return false;
  g->Q = conv3x3_f32_q(C);
  g->CG = ceil_div(C, g->Q);
  g->WB = W / 8;
  g->sub_threads = g->CG * g->WB * 8;
  const int max_threads = res_max_threads(g->Q);
  if (g->sub_threads > max_threads) return false;
  g->units_per_utt = ceil_div(H, kPH);
  g->n_units = B * (int64_t)g->units_per_utt;
  if (g->n_units >= (1ll << 24)) return false;
  g->pad = d < W ? round_up(d, 4) : 0;     // (the register window of the small dilations reads 4 floats either side)
  if (g->pad == 0 && d < W) return false;
  int S = W + 2 * g->pad;
  while ((S / 4) % 2 == 0) S += 4;
  g->S = S;
  g->rs = d < 8 ? d : 8;
  g->rows_st = kPH + 2 * g->rs;
  g->seg_per_row = W / 4;
  g->w_floats = C * 9 * g->CG * qp_of(g->Q);
  const size_t cap = 226 * 1024;
  int ns = max_threads / g->sub_threads;
  if (ns > 8) ns = 8;
  if ((int64_t)ns > g->n_units) ns = (int)g->n_units;
  for (; ns >= 1; --ns) {
    for (int ck = 8; ck >= 2; --ck) {
      if (ck > C && ck > 2) continue;
      const size_t need = sizeof(float) * ((size_t)g->w_floats + 2ull * ns * ck * g->rows_st * S);
      if (need <= cap) {
        g->NS = ns;
        const int chunks = ceil_div(C, ck);
        g->CK = ceil_div(C, chunks);
        g->n_chunks = chunks;
        g->sub_floats = g->CK * g->rows_st * S;
        g->buf_floats = ns * g->sub_floats;
        g->smem = sizeof(float) * ((size_t)g->w_floats + 2ull * g->buf_floats);
        g->threads = round_up(ns * g->sub_threads, 32);
        g->n_items = ceil_div(g->n_units, (int64_t)ns);
        g->m_seg = magic_div((uint32_t)g->seg_per_row);
        g->m_rows = magic_div((uint32_t)g->rows_st);
        g->m_ck = magic_div((uint32_t)g->CK);
        g->m_upu = magic_div((uint32_t)g->units_per_utt);
        return (int64_t)ns * g->CK * g->rows_st * g->seg_per_row < (1 << 20);
      }
    }
  }
  return false;
}

// DM = 1, 2, 4: the dilation (register-window form); DM = 0: dilation a multiple of 4 (>= 8) or >= W (centre tap only)
template <int Q, int DM>
__global__ void __launch_bounds__(res_max_threads(Q), 1)
conv3x3_f32_row_kernel(const Conv3x3F32 a, const ConvRowGeom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int QP = qp_of(Q);
  float* s_w = smem;                      // [C][9][CG*QP]
  float* s_in = smem + g.w_floats;        // [2][NS][CK][rows_st][S]
  const int C = a.C, H = a.H, W = a.W, d = DM > 0 ? DM : a.d;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int S = g.S;
  {
    const float4* src = reinterpret_cast<const float4*>(a.wt);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < g.w_floats / 4; i += nthr) dst[i] = __ldg(src + i);
    float4* z = reinterpret_cast<float4*>(s_in);
    for (int i = tid; i < g.buf_floats / 2; i += nthr) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  // thread -> (unit of the item, map group, column block, row)
  const int sub = tid / g.sub_threads;
  const int rem = tid - sub * g.sub_threads;
  const int cg = rem / (g.WB * 8);
  const int rem2 = rem - cg * (g.WB * 8);
  const int wbi = rem2 >> 3, r = rem2 & 7;
  const bool active = sub < g.NS;
  const bool side_taps = g.pad > 0;

  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_my = first < g.n_items ? (g.n_items - first + step - 1) / step : 0;
  const int n_chunks = g.n_chunks;
  const int n_pieces = g.NS * g.CK * g.rows_st * g.seg_per_row;
  const uint32_t s_in_u32 = (uint32_t)__cvta_generic_to_shared(s_in);

  auto stage = [&](int64_t item, int chunk, int buf) {
    const int c0 = chunk * g.CK;
    const uint32_t unit0 = (uint32_t)(item * g.NS);
    for (int p = tid; p < n_pieces; p += nthr) {
      const uint32_t rowid = fast_div((uint32_t)p, g.m_seg);
      const uint32_t seg = (uint32_t)p - rowid * (uint32_t)g.seg_per_row;
      const uint32_t sc = fast_div(rowid, g.m_rows);
      const int i = (int)(rowid - sc * (uint32_t)g.rows_st);      // staged row
      const uint32_t su = fast_div(sc, g.m_ck);
      const uint32_t ci = sc - su * (uint32_t)g.CK;
      const uint32_t unit = unit0 + su;
      const uint32_t b = fast_div(unit, g.m_upu);
      const int h0 = (int)(unit - b * (uint32_t)g.units_per_utt) * kPH;
      // union of the three taps' rows (d <= 8), or three sets of 8 rows (d > 8)
      const int h = d <= 8 ? h0 - d + i : h0 + (i & 7) + ((i >> 3) - 1) * d;
      const int c = c0 + (int)ci;
      const bool ok = (int64_t)unit < g.n_units && c < C && h >= 0 && h < H;
      const uint32_t dst = s_in_u32 + 4u * ((uint32_t)buf * (uint32_t)g.buf_floats + rowid * (uint32_t)S + (uint32_t)g.pad + seg * 4u);
      const float* src = ok ? a.x + (((int64_t)b * C + c) * H + h) * (int64_t)W + seg * 4 : a.x;
      cp_async16(dst, src, ok ? 16u : 0u);
    }
    cp_async_commit();
  };

  float acc[8][Q];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[j][q] = 0.f;

  if (n_my > 0) stage(first, 0, 0);
  int64_t it_s = 0; int ch_s = 1;
  if (ch_s == n_chunks) { ch_s = 0; ++it_s; }
  int buf = 0;
  const int tap_rows = g.rs * S;   // floats between the rows of two consecutive height taps
  for (int64_t it = 0; it < n_my; ++it) {
    const int64_t item = first + it * step;
    for (int ch = 0; ch < n_chunks; ++ch, buf ^= 1) {
      if (it_s < n_my) {
        stage(first + it_s * step, ch_s, buf ^ 1);
        if (++ch_s == n_chunks) { ch_s = 0; ++it_s; }
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (active) {
        const float* in_base = s_in + buf * g.buf_floats + sub * g.sub_floats + r * S + g.pad + wbi * 8;
        const float* w_base = s_w + ((size_t)(ch * g.CK) * 9 * g.CG + cg) * QP;
        const int ck_n = min(g.CK, C - ch * g.CK);
        for (int ci = 0; ci < ck_n; ++ci) {
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            const float* row = in_base + ci * (g.rows_st * S) + dh * tap_rows;
            const float* wrow = w_base + (size_t)((ci * 9 + dh * 3) * g.CG) * QP;
            if constexpr (DM > 0) {
              float xw[16];
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const float4 t4 = *reinterpret_cast<const float4*>(row - 4 + 4 * v);
                xw[4 * v] = t4.x; xw[4 * v + 1] = t4.y; xw[4 * v + 2] = t4.z; xw[4 * v + 3] = t4.w;
              }
#pragma unroll
              for (int dw = 0; dw < 3; ++dw) {
                const float* wp = wrow + (size_t)(dw * g.CG) * QP;
                float wv[QP];
#pragma unroll
                for (int q4 = 0; q4 < QP / 4; ++q4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(wp + 4 * q4);
                  wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                  for (int q = 0; q < Q; ++q) acc[j][q] = fmaf(xw[4 + j + (dw - 1) * DM], wv[q], acc[j][q]);
              }
            } else {
#pragma unroll
              for (int dw = 0; dw < 3; ++dw) {
                if (dw != 1 && !side_taps) continue;
                const float* ip = row + (dw - 1) * d;
                const float4 x0 = *reinterpret_cast<const float4*>(ip);
                const float4 x1 = *reinterpret_cast<const float4*>(ip + 4);
                const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const float* wp = wrow + (size_t)(dw * g.CG) * QP;
                float wv[QP];
#pragma unroll
                for (int q4 = 0; q4 < QP / 4; ++q4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(wp + 4 * q4);
                  wv[4 * q4] = t4.x; wv[4 * q4 + 1] = t4.y; wv[4 * q4 + 2] = t4.z; wv[4 * q4 + 3] = t4.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                  for (int q = 0; q < Q; ++q) acc[j][q] = fmaf(xv[j], wv[q], acc[j][q]);
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- epilogue of this item: ReLU, residual (pre-BN skip), BatchNorm (resnet.py:49-55); 8 columns = two 16-byte stores
    if (active) {
      const int64_t unit = item * g.NS + sub;
      if (unit < g.n_units) {
        const int64_t b = unit / g.units_per_utt;
        const int h = (int)(unit - b * g.units_per_utt) * kPH + r;
        if (h < H) {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int co = cg * Q + q;
            if (co < C) {
              const float sc = __ldg(a.bn_scale + co), sh = __ldg(a.bn_shift + co);
              const int64_t off = ((b * C + co) * (int64_t)H + h) * W + wbi * 8;
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[j][q], 0.f);
              if (a.prev_in) {
                const float4 p0 = *reinterpret_cast<const float4*>(a.prev_in + off);
                const float4 p1 = *reinterpret_cast<const float4*>(a.prev_in + off + 4);
                v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
                v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
                *reinterpret_cast<float4*>(a.prev_out + off) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(a.prev_out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
              *reinterpret_cast<float4*>(a.y + off) =
                  make_float4(fmaf(v[0], sc, sh), fmaf(v[1], sc, sh), fmaf(v[2], sc, sh), fmaf(v[3], sc, sh));
              *reinterpret_cast<float4*>(a.y + off + 4) =
                  make_float4(fmaf(v[4], sc, sh), fmaf(v[5], sc, sh), fmaf(v[6], sc, sh), fmaf(v[7], sc, sh));
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[j][q] = 0.f;
  }
}

template <int Q>
static int launch_row_q(const Conv3x3F32& a, const ConvRowGeom& g, unsigned grid, cudaStream_t st) {
#define KWS_LAUNCH_ROW(DM_)                                                                                     \
  do {                                                                                                          \
    KWS_CUDA(cudaFuncSetAttribute(conv3x3_f32_row_kernel<Q, DM_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                  226 * 1024));                                                                 \
    conv3x3_f32_row_kernel<Q, DM_><<<grid, g.threads, g.smem, st>>>(a, g);                                      \
  } while (0)
  if (a.d == 1) KWS_LAUNCH_ROW(1);
  else if (a.d == 2) KWS_LAUNCH_ROW(2);
  else if (a.d == 4) KWS_LAUNCH_ROW(4);
  else KWS_LAUNCH_ROW(0);
#undef KWS_LAUNCH_ROW
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

int launch_conv3x3_f32(const Conv3x3F32& a, cudaStream_t st) {
  KWS_REQUIRE(a.C >= 1 && a.H >= 1 && a.W >= 1 && a.d >= 1, "conv3x3: bad shape");
  KWS_REQUIRE((a.prev_in == nullptr) == (a.prev_out == nullptr), "conv3x3: prev_in/prev_out mismatch");
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y) |
                           reinterpret_cast<uintptr_t>(a.prev_in) | reinterpret_cast<uintptr_t>(a.prev_out)) & 15) == 0;
  ConvRowGeom wg;
  if (a.resident == 1 && aligned16 && conv3x3_row_geom(a.B, a.C, a.H, a.W, a.d, &wg)) {
    const unsigned grid = (unsigned)std::min<int64_t>(wg.n_items, kNumSMs);
    switch (wg.Q) {
      case 8: return launch_row_q<8>(a, wg, grid, st);
      case 9: return launch_row_q<9>(a, wg, grid, st);
      case 10: return launch_row_q<10>(a, wg, grid, st);
      case 11: return launch_row_q<11>(a, wg, grid, st);
      case 12: return launch_row_q<12>(a, wg, grid, st);
      default:
        set_error("conv3x3: no kernel for Q=%d", wg.Q);
        return KWS_ERR_INVALID;
    }
  }
  ConvResGeom rg;
  if (a.resident && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && conv3x3_res_geom(a.B, a.C, a.H, a.W, a.d, &rg)) {
    const unsigned grid = (unsigned)std::min<int64_t>(rg.n_items, kNumSMs);
#define KWS_LAUNCH_C3R(Q_)                                                                        \
  case Q_:                                                                                        \
    if (rg.row_stride == 72) {                                                                    \
      KWS_CUDA(cudaFuncSetAttribute(conv3x3_f32_res_kernel<Q_, 72>,                               \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));    \
      conv3x3_f32_res_kernel<Q_, 72><<<grid, rg.threads, rg.smem, st>>>(a, rg);                   \
    } else {                                                                                      \
      KWS_CUDA(cudaFuncSetAttribute(conv3x3_f32_res_kernel<Q_, 0>,                                \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));    \
      conv3x3_f32_res_kernel<Q_, 0><<<grid, rg.threads, rg.smem, st>>>(a, rg);                    \
    }                                                                                             \
    break;
    switch (rg.Q) {
      KWS_LAUNCH_C3R(8)
      KWS_LAUNCH_C3R(9)
      KWS_LAUNCH_C3R(10)
      KWS_LAUNCH_C3R(11)
      KWS_LAUNCH_C3R(12)
      default:
        set_error("conv3x3: no kernel for Q=%d", rg.Q);
        return KWS_ERR_INVALID;
    }
#undef KWS_LAUNCH_C3R
    KWS_CHECK_LAUNCH();
    return KWS_OK;
  }
  Conv3x3Geom g;
  KWS_REQUIRE(conv3x3_geom(a.C, a.H, a.W, a.d, &g),
              "conv3x3 fp32: unsupported geometry C=%d H=%d W=%d d=%d", a.C, a.H, a.W, a.d);
  KWS_REQUIRE(a.B <= 65535, "conv3x3: chunk too large");
  dim3 grid(g.tiles_h, (unsigned)a.B);
#define KWS_LAUNCH_C3(Q_)                                                                        \
  case Q_:                                                                                       \
    KWS_CUDA(cudaFuncSetAttribute(conv3x3_f32_kernel<Q_>,                                        \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));     \
    conv3x3_f32_kernel<Q_><<<grid, g.threads, g.smem, st>>>(a, g);                               \
    break;
  switch (g.Q) {
    KWS_LAUNCH_C3(8)
    KWS_LAUNCH_C3(9)
    KWS_LAUNCH_C3(10)
    KWS_LAUNCH_C3(11)
    KWS_LAUNCH_C3(12)
    default:
      set_error("conv3x3: no kernel for Q=%d", g.Q);
      return KWS_ERR_INVALID;
  }
#undef KWS_LAUNCH_C3
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

// =============================================================================================
// tail: one CTA per utterance; a warp reduces one map at a time, then n_labels dot products.
constexpr int kTailThreads = 256;

__global__ void __launch_bounds__(kTailThreads)
tail_f32_kernel(const float* __restrict__ y, const float* __restrict__ out_w,
                const float* __restrict__ out_b, float* __restrict__ logits, int C, int HW,
                int n_labels) {
  extern __shared__ float s_mean[];  // [C]
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* yb = y + b * (int64_t)C * HW;
  const bool vec = (HW & 3) == 0;
  for (int c = warp; c < C; c += kTailThreads / 32) {
    const float* p = yb + (int64_t)c * HW;
    float s = 0.f;
    if (vec) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      for (int i = lane; i < HW / 4; i += 32) {
        const float4 v = __ldg(p4 + i);
        s += (v.x + v.y) + (v.z + v.w);
      }
    } else {
      for (int i = lane; i < HW; i += 32) s += __ldg(p + i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_mean[c] = s / (float)HW;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < n_labels; l += kTailThreads) {
    float v = __ldg(out_b + l);
    for (int c = 0; c < C; ++c) v = fmaf(s_mean[c], __ldg(out_w + l * C + c), v);
    logits[b * n_labels + l] = v;
  }
}

int launch_tail_f32(const float* y, const float* out_w, const float* out_b, float* logits, int64_t B,
                    int C, int HW, int n_labels, cudaStream_t st) {
  KWS_REQUIRE(C * sizeof(float) <= 48 * 1024, "tail: too many maps (%d)", C);
  tail_f32_kernel<<<(unsigned)B, kTailThreads, C * sizeof(float), st>>>(y, out_w, out_b, logits, C, HW,
                                                                      n_labels);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

}  // namespace kws
