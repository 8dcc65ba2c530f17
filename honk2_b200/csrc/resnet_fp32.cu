// fp32 CUDA-core kernels of the ResNet path (parity mode, and the final path for the 19-map
// "-narrow" nets and the 1-channel conv_0).  Activations are planar [B][C][H][W] float32, the
// same order torch uses, so every intermediate can be compared with the reference module.
//
//   conv0_f32_kernel    conv_0 + ReLU + AvgPool          /root/reference/model/resnet.py:40-44
//   conv3x3_f32_kernel  conv_i + ReLU + skip + BatchNorm /root/reference/model/resnet.py:48-55
//   tail_f32_kernel     mean over H*W + Linear           /root/reference/model/resnet.py:57-59
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
constexpr int kQP = 12;    // padded Q (three float4 per tap)

struct Conv3x3Geom {
  int Q, CG, RG, CK, wpad, row_stride, threads, tiles_h;
  size_t smem;
};

int conv3x3_f32_q(int C) {
  // fewest padded maps; ties -> larger Q
  int best = 9, waste = 1 << 30;
  for (int q = 12; q >= 8; --q) {
    const int w = ceil_div(C, q) * q - C;
    if (w < waste) { waste = w; best = q; }
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
  const size_t per_ch = sizeof(float) * (3 * g->RG * kPH * g->row_stride + 9 * g->CG * kQP);
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
  float* s_w = smem + g.CK * 3 * Hr * g.row_stride;     // [CK][9][CG*12]
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
  const int w_slab = 9 * g.CG * kQP;
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
    // ---- stage weights: contiguous slab [CK][9][CG*12]
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
            const float* wp = s_w + ((ci * 9 + dh * 3 + dw) * g.CG + cg) * kQP;
            float wv[kQP];
#pragma unroll
            for (int q4 = 0; q4 < kQP / 4; ++q4) {
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

int launch_conv3x3_f32(const Conv3x3F32& a, cudaStream_t st) {
  Conv3x3Geom g;
  KWS_REQUIRE(a.C >= 1 && a.H >= 1 && a.W >= 1 && a.d >= 1, "conv3x3: bad shape");
  KWS_REQUIRE(conv3x3_geom(a.C, a.H, a.W, a.d, &g),
              "conv3x3 fp32: unsupported geometry C=%d H=%d W=%d d=%d", a.C, a.H, a.W, a.d);
  KWS_REQUIRE(a.B <= 65535, "conv3x3: chunk too large");
  KWS_REQUIRE((a.prev_in == nullptr) == (a.prev_out == nullptr), "conv3x3: prev_in/prev_out mismatch");
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
