// fp32 CUDA-core kernels for the CNN family (cnn-trad-fpool3 and the other nine members).
// Activations are planar [B][C][H][W] float32 like torch, so `x.view(B, -1)` (cnn.py:93) is the
// buffer itself.
//
//   conv_gen_f32_kernel  Conv2d(stride, no padding, bias) + ReLU   /root/reference/model/cnn.py:82-83,87-88
//   maxpool_f32_kernel   MaxPool2d(stride = kernel, floor)         /root/reference/model/cnn.py:85,91
//   linear_f32_kernel    Linear                                    /root/reference/model/cnn.py:95-106
#include "kernels.cuh"

namespace kws {

// =============================================================================================
// Direct convolution.  thread = (cout group of 8, row group of 4 output rows, output column);
// lanes run along the output column.  Input maps are staged CK at a time in shared memory
// (full input rows), weights are read through the read-only path as two float4 per tap --
// every lane of a warp that shares a cout group reads the same address.
constexpr int kCQ = 8;   // output maps per thread
constexpr int kCR = 4;   // output rows per thread

struct ConvGenGeom {
  int Ho, Wo, CG, CGB, RG, WT, CK, in_rows, threads, tiles_h, tiles_w, zsplit, CoutPad;
  size_t smem;
};

static bool conv_gen_geom(const ConvGenF32& a, ConvGenGeom* g) {
  g->Ho = (a.H - a.KH) / a.SH + 1;
  g->Wo = (a.W - a.KW) / a.SW + 1;
  if (a.H < a.KH || a.W < a.KW || g->Ho < 1 || g->Wo < 1) return false;
  g->CG = ceil_div(a.Cout, kCQ);
  g->CoutPad = g->CG * kCQ;
  g->WT = g->Wo <= 64 ? g->Wo : 32;
  g->tiles_w = ceil_div(g->Wo, g->WT);
  const int max_cgb = max(1, 256 / g->WT);
  g->zsplit = ceil_div(g->CG, max_cgb);
  g->CGB = ceil_div(g->CG, g->zsplit);
  const int need_rg = ceil_div(g->Ho, kCR);
  g->RG = min(max(1, 256 / (g->WT * g->CGB)), need_rg);
  g->tiles_h = ceil_div(need_rg, g->RG);
  g->RG = ceil_div(need_rg, g->tiles_h);
  g->threads = round_up(g->WT * g->CGB * g->RG, 32);
  g->in_rows = (g->RG * kCR - 1) * a.SH + a.KH;
  const size_t per_ch = sizeof(float) * g->in_rows * a.W;
  int ck = (int)((40 * 1024) / per_ch);
  if (ck < 1) ck = 1;
  if (ck > a.Cin) ck = a.Cin;
  g->CK = ceil_div(a.Cin, ceil_div(a.Cin, ck));
  g->smem = per_ch * g->CK;
  return g->smem <= 200 * 1024;
}

__global__ void __launch_bounds__(256)
conv_gen_f32_kernel(ConvGenF32 a, ConvGenGeom g) {
  extern __shared__ __align__(16) float s_in[];  // [CK][in_rows][W]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int tile_h = blockIdx.x / g.tiles_w, tile_w = blockIdx.x - tile_h * g.tiles_w;
  const int ho0 = tile_h * g.RG * kCR;
  const int wo0 = tile_w * g.WT;
  const int cg0 = blockIdx.z * g.CGB;

  const int per_cg = g.RG * g.WT;
  const int cgl = tid / per_cg;
  const int rem = tid - cgl * per_cg;
  const int rg = rem / g.WT;
  const int wo = wo0 + (rem - rg * g.WT);
  const int cg = cg0 + cgl;
  const bool active = cgl < g.CGB && cg < g.CG && wo < g.Wo;

  float acc[kCR][kCQ];
#pragma unroll
  for (int j = 0; j < kCR; ++j)
#pragma unroll
    for (int q = 0; q < kCQ; ++q) acc[j][q] = 0.f;

  const float* xb = a.x + b * (int64_t)a.Cin * a.H * a.W;
  const int h_in0 = ho0 * a.SH;
  const int plane = g.in_rows * a.W;

  for (int c0 = 0; c0 < a.Cin; c0 += g.CK) {
    const int n = g.CK * plane;
    for (int i = tid; i < n; i += blockDim.x) {
      const int ci = i / plane;
      const int r2 = i - ci * plane;
      const int r = r2 / a.W, w = r2 - r * a.W;
      const int h = h_in0 + r;
      s_in[i] = (c0 + ci < a.Cin && h < a.H) ? __ldg(xb + ((int64_t)(c0 + ci) * a.H + h) * a.W + w) : 0.f;
    }
    __syncthreads();
    if (active) {
      const int ck = min(g.CK, a.Cin - c0);
      for (int ci = 0; ci < ck; ++ci) {
        const float* wbase = a.wt + (int64_t)(c0 + ci) * a.KH * a.KW * g.CoutPad + cg * kCQ;
        const float* ibase = s_in + ci * plane + (rg * kCR * a.SH) * a.W + wo * a.SW;
        for (int kh = 0; kh < a.KH; ++kh) {
          for (int kw = 0; kw < a.KW; ++kw) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wbase + (kh * a.KW + kw) * g.CoutPad));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wbase + (kh * a.KW + kw) * g.CoutPad) + 1);
            const float wv[kCQ] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float xv[kCR];
#pragma unroll
            for (int j = 0; j < kCR; ++j) xv[j] = ibase[(j * a.SH + kh) * a.W + kw];
#pragma unroll
            for (int j = 0; j < kCR; ++j)
#pragma unroll
              for (int q = 0; q < kCQ; ++q) acc[j][q] = fmaf(xv[j], wv[q], acc[j][q]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!active) return;
#pragma unroll
  for (int q = 0; q < kCQ; ++q) {
    const int co = cg * kCQ + q;
    if (co >= a.Cout) break;
    const float bias = __ldg(a.bias + co);
#pragma unroll
    for (int j = 0; j < kCR; ++j) {
      const int ho = ho0 + rg * kCR + j;
      if (ho >= g.Ho) break;
      a.y[((b * a.Cout + co) * (int64_t)g.Ho + ho) * g.Wo + wo] = fmaxf(acc[j][q] + bias, 0.f);
    }
  }
}

int launch_conv_gen_f32(const ConvGenF32& a, cudaStream_t st) {
  ConvGenGeom g;
  KWS_REQUIRE(a.Cin >= 1 && a.Cout >= 1 && a.KH >= 1 && a.KW >= 1 && a.SH >= 1 && a.SW >= 1,
              "conv: bad shape");
  KWS_REQUIRE(conv_gen_geom(a, &g), "conv fp32: unsupported geometry Cin=%d %dx%d k=%dx%d s=%dx%d", a.Cin,
              a.H, a.W, a.KH, a.KW, a.SH, a.SW);
  KWS_REQUIRE(a.B <= 65535 && g.zsplit <= 65535, "conv: chunk too large");
  KWS_CUDA(cudaFuncSetAttribute(conv_gen_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                200 * 1024));
  dim3 grid(g.tiles_h * g.tiles_w, (unsigned)a.B, g.zsplit);
  conv_gen_f32_kernel<<<grid, g.threads, g.smem, st>>>(a, g);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

// =============================================================================================
__global__ void __launch_bounds__(256)
maxpool_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int H, int W,
                   int Ho, int Wo, int kh, int kw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wo = (int)(i % Wo);
  const int64_t t = i / Wo;
  const int ho = (int)(t % Ho);
  const int64_t plane = t / Ho;
  const float* p = x + (plane * H + (int64_t)ho * kh) * W + wo * kw;
  float m = p[0];
  for (int a = 0; a < kh; ++a)
    for (int e = 0; e < kw; ++e) {
      const float v = p[a * W + e];
      m = (v > m || v != v) ? v : m;  // NaN propagates like torch
    }
  y[i] = m;
}

int launch_maxpool_f32(const float* x, float* y, int64_t planes, int H, int W, int kh, int kw,
                       cudaStream_t st) {
  KWS_REQUIRE(kh >= 1 && kw >= 1 && H >= kh && W >= kw, "maxpool: window %dx%d on %dx%d", kh, kw, H, W);
  const int Ho = H / kh, Wo = W / kw;
  const int64_t total = planes * Ho * Wo;
  if (total == 0) return KWS_OK;
  const int64_t blocks = ceil_div<int64_t>(total, 256);
  KWS_REQUIRE(blocks < 2147483647LL, "maxpool: too many elements");
  maxpool_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, total, H, W, Ho, Wo, kh, kw);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

// =============================================================================================
// Linear as a tiled SGEMM: CTA tile 64 (rows of x) x 32 (output features), BK = 16, 256 threads,
// each thread a 4 x 2 register tile.  K is summed in order, one fmaf chain per output.
constexpr int kLBM = 64, kLBN = 32, kLBK = 16;

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int64_t M, int N, int K) {
  __shared__ float xs[kLBK][kLBM + 4];
  __shared__ float ws[kLBK][kLBN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * kLBM;
  const int n0 = blockIdx.y * kLBN;
  const int tm = tid >> 4, tn = tid & 15;  // 16 x 16 threads: rows tm*4.., cols tn*2..
  float acc[4][2] = {};
  // loader mapping: x tile 64x16 -> 1024 elements, 4 per thread; w tile 32x16 -> 512, 2 per thread
  for (int k0 = 0; k0 < K; k0 += kLBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int64_t m = m0 + r;
      xs[kk][r] = (m < M && k0 + kk < K) ? __ldg(x + m * K + k0 + kk) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = tid + e * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int n = n0 + r;
      ws[kk][r] = (n < N && k0 + kk < K) ? __ldg(w + (int64_t)n * K + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kLBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&xs[kk][tm * 4]);
      const float2 bv = *reinterpret_cast<const float2*>(&ws[kk][tn * 2]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(a4[i], bv.x, acc[i][0]);
        acc[i][1] = fmaf(a4[i], bv.y, acc[i][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm * 4 + i;
    if (m >= M) break;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tn * 2 + j;
      if (n < N) y[m * N + n] = acc[i][j] + __ldg(bias + n);
    }
  }
}

int launch_linear_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int N,
                      int K, cudaStream_t st) {
  KWS_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear: bad shape");
  if (M == 0) return KWS_OK;
  dim3 grid((unsigned)ceil_div<int64_t>(M, kLBM), ceil_div(N, kLBN));
  linear_f32_kernel<<<grid, 256, 0, st>>>(x, w, bias, y, M, N, K);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

}  // namespace kws
