// fp32 CUDA-core kernels for the CNN family (cnn-trad-fpool3 and the other nine members).
// Activations are planar [B][C][H][W] float32 like torch, so `x.view(B, -1)` (cnn.py:93) is the
// buffer itself.
//
//   conv_row_f32_kernel  Conv2d(stride 1, 4 or 8 kernel columns, bias) + ReLU: persistent row-tile kernel (cnn-trad-fpool3's layers)
//   conv_gen_f32_kernel  Conv2d(stride, no padding, bias) + ReLU   /root/reference/model/cnn.py:82-83,87-88 (every other shape)
//   maxpool_f32_kernel   MaxPool2d(stride = kernel, floor)         /root/reference/model/cnn.py:85,91
//   linear_f32_kernel    Linear (split-K + linear_reduce_f32_kernel for long reductions)   /root/reference/model/cnn.py:95-106
#include <algorithm>
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

// =============================================================================================
// Stride-1 convolutions with 4 or 8 kernel columns (both convolutions of cnn-trad-fpool3, cnn.py:82-88): the scheme of
// the ResNet path's fp32 row kernel (resnet_fp32.cu).  Persistent CTA; thread = one output row x 8 output columns x 8
// output maps, so the KW width taps of an input row come out of one register window of 8 + KW - 1 floats (16-byte
// shared-memory loads, row pitch S with S/4 odd: conflict free) and a tap costs two 16-byte weight loads per 64 FMAs;
// a CTA works on NS "units" (utterance, 8 output rows) numbered through the sub-batch; the input rows -- and, when the
// layer's weights do not fit in shared memory (conv_1: 655 KB), the weights -- of the next CK input maps arrive by
// cp.async into the other half of a double buffer while the current chunk is computed.
constexpr int kRowThreadsMax = 512;

struct ConvRowGenGeom {
  int Ho, Wo, WB, CG, CoutPad, NS, CK, n_chunks, RI, S, NW, threads, sub_threads, units_per_utt, seg_per_row, vec;
  int w_resident, w_ci_floats, w_floats, sub_floats, in_floats, buf_floats;
  int64_t n_units, n_items;
  uint32_t m_seg, m_ri, m_ck, m_upu;
  size_t smem;
};

static uint32_t magic_div_u32(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ uint32_t fast_div_u32(uint32_t n, uint32_t m) { return m == 0u ? n : __umulhi(n, m); }

static bool conv_row_gen_geom(const ConvGenF32& a, ConvRowGenGeom* g) {
  if (a.SH != 1 || a.SW != 1 || (a.KW != 4 && a.KW != 8)) return false;
  if (a.H < a.KH || a.W < a.KW) return false;
  g->Ho = a.H - a.KH + 1;
  g->Wo = a.W - a.KW + 1;
  g->WB = ceil_div(g->Wo, 8);
  g->CG = ceil_div(a.Cout, kCQ);
  g->CoutPad = g->CG * kCQ;
  g->sub_threads = 8 * g->WB * g->CG;
  if (g->sub_threads > kRowThreadsMax) return false;
  g->units_per_utt = ceil_div(g->Ho, 8);
  g->n_units = a.B * (int64_t)g->units_per_utt;
  if (g->n_units >= (1ll << 24) || g->n_units < 1) return false;
  g->RI = 8 + a.KH - 1;
  int S = round_up(std::max(g->WB * 8 + a.KW - 1, a.W), 4);
  while ((S / 4) % 2 == 0) S += 4;
  g->S = S;
  g->NW = ceil_div(8 + a.KW - 1, 4);
  g->vec = (a.W % 4 == 0) ? 1 : 0;
  g->seg_per_row = g->vec ? a.W / 4 : a.W;
  g->w_ci_floats = a.KH * a.KW * g->CoutPad;
  const size_t cap = 224 * 1024;
  const size_t w_all = (size_t)a.Cin * g->w_ci_floats * sizeof(float);
  g->w_resident = w_all <= 96 * 1024 ? 1 : 0;
  int ns = kRowThreadsMax / g->sub_threads;
  if (ns > 8) ns = 8;
  if ((int64_t)ns > g->n_units) ns = (int)g->n_units;
  for (; ns >= 1; --ns) {
    for (int ck = 8; ck >= 1; --ck) {
      if (ck > a.Cin) continue;
      const size_t in_b = sizeof(float) * (size_t)ns * ck * g->RI * S;
      const size_t w_b = g->w_resident ? 0 : sizeof(float) * (size_t)ck * g->w_ci_floats;
      const size_t need = (g->w_resident ? w_all : 0) + 2 * (in_b + w_b);
      if (need > cap) continue;
      g->NS = ns;
      const int chunks = ceil_div(a.Cin, ck);
      g->CK = ceil_div(a.Cin, chunks);
      g->n_chunks = chunks;
      g->sub_floats = g->CK * g->RI * S;
      g->in_floats = ns * g->sub_floats;
      g->w_floats = g->w_resident ? a.Cin * g->w_ci_floats : g->CK * g->w_ci_floats;
      g->buf_floats = g->in_floats + (g->w_resident ? 0 : g->w_floats);
      g->smem = sizeof(float) * ((size_t)(g->w_resident ? g->w_floats : 0) + 2ull * g->buf_floats);
      g->threads = round_up(ns * g->sub_threads, 32);
      g->n_items = ceil_div(g->n_units, (int64_t)ns);
      g->m_seg = magic_div_u32((uint32_t)g->seg_per_row);
      g->m_ri = magic_div_u32((uint32_t)g->RI);
      g->m_ck = magic_div_u32((uint32_t)g->CK);
      g->m_upu = magic_div_u32((uint32_t)g->units_per_utt);
      return (int64_t)ns * g->CK * g->RI * g->seg_per_row < (1 << 20);
    }
  }
  return false;
}

__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cpa4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int KW>
__global__ void __launch_bounds__(kRowThreadsMax, 1)
conv_row_f32_kernel(const ConvGenF32 a, const ConvRowGenGeom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NWIN = (8 + KW - 1 + 3) / 4;          // 16-byte loads per input row window
  // layout: [resident weights][buffer 0: inputs | chunk weights][buffer 1: ...]
  float* s_wres = smem;
  float* s_buf = smem + (g.w_resident ? g.w_floats : 0);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int S = g.S;
  {
    if (g.w_resident) {
      const float4* src = reinterpret_cast<const float4*>(a.wt);
      float4* dst = reinterpret_cast<float4*>(s_wres);
      for (int i = tid; i < g.w_floats / 4; i += nthr) dst[i] = __ldg(src + i);
    }
    // the columns past the input width stay zero (cp.async only ever writes [0, W) of a staged row)
    float4* z = reinterpret_cast<float4*>(s_buf);
    for (int i = tid; i < 2 * g.buf_floats / 4; i += nthr) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const int sub = tid / g.sub_threads;
  const int rem = tid - sub * g.sub_threads;
  const int cg = rem / (g.WB * 8);
  const int rem2 = rem - cg * (g.WB * 8);
  const int wbi = rem2 >> 3, r = rem2 & 7;
  const bool active = sub < g.NS;

  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t n_my = first < g.n_items ? (g.n_items - first + step - 1) / step : 0;
  const int n_chunks = g.n_chunks;
  const int n_pieces = g.NS * g.CK * g.RI * g.seg_per_row;
  const uint32_t s_buf_u32 = (uint32_t)__cvta_generic_to_shared(s_buf);
  const int64_t in_utt = (int64_t)a.Cin * a.H * a.W;

  auto stage = [&](int64_t item, int chunk, int buf) {
    const int c0 = chunk * g.CK;
    const uint32_t unit0 = (uint32_t)(item * g.NS);
    const uint32_t base = s_buf_u32 + 4u * (uint32_t)buf * (uint32_t)g.buf_floats;
    for (int p = tid; p < n_pieces; p += nthr) {
      const uint32_t rowid = fast_div_u32((uint32_t)p, g.m_seg);
      const uint32_t seg = (uint32_t)p - rowid * (uint32_t)g.seg_per_row;
      const uint32_t sc = fast_div_u32(rowid, g.m_ri);
      const int i = (int)(rowid - sc * (uint32_t)g.RI);
      const uint32_t su = fast_div_u32(sc, g.m_ck);
      const uint32_t ci = sc - su * (uint32_t)g.CK;
      const uint32_t unit = unit0 + su;
      const uint32_t b = fast_div_u32(unit, g.m_upu);
      const int h = (int)(unit - b * (uint32_t)g.units_per_utt) * 8 + i;     // input row = output row + kh (stride 1)
      const int c = c0 + (int)ci;
      const bool ok = (int64_t)unit < g.n_units && c < a.Cin && h < a.H;
      const uint32_t dst = base + 4u * (rowid * (uint32_t)S + seg * (g.vec ? 4u : 1u));
      const float* src = ok ? a.x + (int64_t)b * in_utt + ((int64_t)c * a.H + h) * a.W + seg * (g.vec ? 4 : 1) : a.x;
      if (g.vec) cpa16(dst, src, ok ? 16u : 0u);
      else cpa4(dst, src, ok ? 4u : 0u);
    }
    if (!g.w_resident) {
      // this chunk's weights: CK x [KH][KW][CoutPad] contiguous floats (zero fill for the maps past Cin)
      const uint32_t wdst = base + 4u * (uint32_t)g.in_floats;
      const int n4 = g.w_floats / 4;
      const int valid4 = max(0, min(g.CK, a.Cin - c0)) * g.w_ci_floats / 4;
      const float4* src = reinterpret_cast<const float4*>(a.wt + (int64_t)c0 * g.w_ci_floats);
      for (int i = tid; i < n4; i += nthr) cpa16(wdst + 16u * (uint32_t)i, i < valid4 ? (const void*)(src + i) : (const void*)a.wt, i < valid4 ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][kCQ];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int q = 0; q < kCQ; ++q) acc[j][q] = 0.f;

  if (n_my > 0) stage(first, 0, 0);
  int64_t it_s = 0; int ch_s = 1;
  if (ch_s == n_chunks) { ch_s = 0; ++it_s; }
  int buf = 0;
  const int w_tap = g.CoutPad;                 // floats between consecutive taps of one input map
  for (int64_t it = 0; it < n_my; ++it) {
    const int64_t item = first + it * step;
    for (int ch = 0; ch < n_chunks; ++ch, buf ^= 1) {
      if (it_s < n_my) {
        stage(first + it_s * step, ch_s, buf ^ 1);
        if (++ch_s == n_chunks) { ch_s = 0; ++it_s; }
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      if (active) {
        const float* bufp = s_buf + buf * g.buf_floats;
        const float* in_base = bufp + sub * g.sub_floats + r * S + wbi * 8;
        const float* w_base = (g.w_resident ? s_wres + (size_t)(ch * g.CK) * g.w_ci_floats : bufp + g.in_floats) + cg * kCQ;
        const int ck_n = min(g.CK, a.Cin - ch * g.CK);
        for (int ci = 0; ci < ck_n; ++ci) {
          const float* in_ci = in_base + ci * (g.RI * S);
          const float* w_ci = w_base + (size_t)ci * g.w_ci_floats;
#pragma unroll 2
          for (int kh = 0; kh < a.KH; ++kh) {
            float xw[4 * NWIN];
#pragma unroll
            for (int v = 0; v < NWIN; ++v) {
              const float4 t4 = *reinterpret_cast<const float4*>(in_ci + kh * S + 4 * v);
              xw[4 * v] = t4.x; xw[4 * v + 1] = t4.y; xw[4 * v + 2] = t4.z; xw[4 * v + 3] = t4.w;
            }
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
              const float* wp = w_ci + (kh * KW + kw) * w_tap;
              const float4 w0 = *reinterpret_cast<const float4*>(wp);
              const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
              const float wv[kCQ] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < kCQ; ++q) acc[j][q] = fmaf(xw[j + kw], wv[q], acc[j][q]);
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- epilogue: bias + ReLU (cnn.py:82-83, 87-88)
    if (active) {
      const int64_t unit = item * g.NS + sub;
      if (unit < g.n_units) {
        const int64_t b = unit / g.units_per_utt;
        const int ho = (int)(unit - b * g.units_per_utt) * 8 + r;
        if (ho < g.Ho) {
#pragma unroll
          for (int q = 0; q < kCQ; ++q) {
            const int co = cg * kCQ + q;
            if (co < a.Cout) {
              const float bias = __ldg(a.bias + co);
              float* dst = a.y + ((b * a.Cout + co) * (int64_t)g.Ho + ho) * g.Wo + wbi * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (wbi * 8 + j < g.Wo) dst[j] = fmaxf(acc[j][q] + bias, 0.f);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < kCQ; ++q) acc[j][q] = 0.f;
  }
}

int launch_conv_gen_f32(const ConvGenF32& a, cudaStream_t st) {
  KWS_REQUIRE(a.Cin >= 1 && a.Cout >= 1 && a.KH >= 1 && a.KW >= 1 && a.SH >= 1 && a.SW >= 1,
              "conv: bad shape");
  ConvRowGenGeom rg;
  if (a.row_kernel && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && conv_row_gen_geom(a, &rg)) {
    const unsigned grid = (unsigned)std::min<int64_t>(rg.n_items, kNumSMs);
    if (a.KW == 4) {
      KWS_CUDA(cudaFuncSetAttribute(conv_row_f32_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      conv_row_f32_kernel<4><<<grid, rg.threads, rg.smem, st>>>(a, rg);
    } else {
      KWS_CUDA(cudaFuncSetAttribute(conv_row_f32_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      conv_row_f32_kernel<8><<<grid, rg.threads, rg.smem, st>>>(a, rg);
    }
    KWS_CHECK_LAUNCH();
    return KWS_OK;
  }
  ConvGenGeom g;
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
// Split-K (gridDim.z > 1): CTA z sums k in [z * kps, (z + 1) * kps) into its own [M][N] slice of `y` (no bias);
// linear_reduce_f32_kernel then adds the slices in slice order -- the first Linear of the CNN family has M = the
// sub-batch, N = 32 and K = 37376, i.e. four CTAs walking 2336 K steps each without it.
constexpr int kLBM = 64, kLBN = 32, kLBK = 16;

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int64_t M, int N, int K, int kps) {
  __shared__ float xs[kLBK][kLBM + 4];
  __shared__ float ws[kLBK][kLBN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * kLBM;
  const int n0 = blockIdx.y * kLBN;
  const int kb = blockIdx.z * kps, ke = min(K, kb + kps);
  y += (int64_t)blockIdx.z * M * N;
  const int tm = tid >> 4, tn = tid & 15;  // 16 x 16 threads: rows tm*4.., cols tn*2..
  float acc[4][2] = {};
  // loader mapping: x tile 64x16 -> 1024 elements, 4 per thread; w tile 32x16 -> 512, 2 per thread
  for (int k0 = kb; k0 < ke; k0 += kLBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int64_t m = m0 + r;
      xs[kk][r] = (m < M && k0 + kk < ke) ? __ldg(x + m * K + k0 + kk) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = tid + e * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int n = n0 + r;
      ws[kk][r] = (n < N && k0 + kk < ke) ? __ldg(w + (int64_t)n * K + k0 + kk) : 0.f;
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
      if (n < N) y[m * N + n] = acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
    }
  }
}

__global__ void __launch_bounds__(256)
linear_reduce_f32_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ y,
                         int64_t MN, int N, int splits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float v = part[i];
  for (int s = 1; s < splits; ++s) v += part[(int64_t)s * MN + i];   // fixed order: reproducible
  y[i] = v + __ldg(bias + (int)(i % N));
}

int launch_linear_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int N,
                      int K, float* scratch, size_t scratch_floats, cudaStream_t st) {
  KWS_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear: bad shape");
  if (M == 0) return KWS_OK;
  const unsigned gm = (unsigned)ceil_div<int64_t>(M, kLBM), gn = (unsigned)ceil_div(N, kLBN);
  // Split a long reduction into slices of 1024: the slice count depends on K alone, never on M, so that a row's sum is
  // added in the same order whatever the batch or sub-batch size (utterances stay bit-independent of their batch).
  constexpr int kps = 1024;
  const int splits = K >= 4 * kps ? ceil_div(K, kps) : 1;
  if (splits <= 1 || scratch == nullptr || scratch_floats < (size_t)splits * (size_t)(M * N)) {
    linear_f32_kernel<<<dim3(gm, gn, 1), 256, 0, st>>>(x, w, bias, y, M, N, K, K);
    KWS_CHECK_LAUNCH();
    return KWS_OK;
  }
  linear_f32_kernel<<<dim3(gm, gn, (unsigned)splits), 256, 0, st>>>(x, w, nullptr, scratch, M, N, K, kps);
  KWS_CHECK_LAUNCH();
  linear_reduce_f32_kernel<<<(unsigned)ceil_div<int64_t>(M * N, 256), 256, 0, st>>>(scratch, bias, y, M * N, N, splits);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

}  // namespace kws
