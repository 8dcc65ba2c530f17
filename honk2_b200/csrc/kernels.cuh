// Internal launch interface between the model orchestration (model.cu) and the kernel files.
// All pointers are device pointers; every launcher returns a KWS_* status.
#pragma once
#include "common.cuh"

namespace kws {

// ---- fp32 CUDA-core path (resnet_fp32.cu); activations are planar [B][C][H][W] float ----------

// conv_0 (1 -> C, 3x3, pad 1, no bias) + ReLU + AvgPool(ph,pw) (resnet.py:40-44).
// feat [B][T][F] -> out [B][C][Ho][Wo] with Ho = T/ph, Wo = F/pw (floor).  w0: [C][9].
int launch_conv0_f32(const float* feat, const float* w0, float* out, int64_t B, int T, int F, int C,
                     int ph, int pw, cudaStream_t st);

// conv_i (C -> C, 3x3, dilation = padding = d, no bias) + ReLU + optional residual + BN
// (resnet.py:48-55).  wt is the packed layout made by pack_conv3x3_f32 (model.cu).
struct Conv3x3F32 {
  const float* x;        // [B][C][H][W] input (post-BN output of the previous layer)
  const float* wt;       // [C][9][CG*QP] zero padded (QP = Q rounded up to a multiple of 4)
  const float* prev_in;  // pre-BN skip tensor to add (even layers) or nullptr
  float* prev_out;       // where the new skip tensor goes (even layers; may alias prev_in)
  float* y;              // [B][C][H][W] BN output
  const float* bn_scale; // [C] 1/sqrt(var+eps)
  const float* bn_shift; // [C] -mean*scale
  int64_t B;
  int C, H, W, d;
  int resident = 1;      // 0: never the resident-weight persistent kernel (HONK2_F32_RESIDENT=0, read per model handle)
};
int launch_conv3x3_f32(const Conv3x3F32& a, cudaStream_t st);
// Q (output channels per thread) the conv3x3 kernel uses for C maps, and the padded row width.
int conv3x3_f32_q(int C);

// global mean over H*W + Linear(C -> n_labels) (resnet.py:57-59).  y [B][C][HW].
int launch_tail_f32(const float* y, const float* out_w, const float* out_b, float* logits, int64_t B,
                    int C, int HW, int n_labels, cudaStream_t st);

// ---- generic fp32 layers for the CNN family (cnn_fp32.cu) ------------------------------------

// Conv2d(Cin -> Cout, (KH,KW), stride (SH,SW), no padding, bias) + ReLU (cnn.py:82-83, :87-88).
// x [B][Cin][H][W] -> y [B][Cout][Ho][Wo]; wt packed [Cin][KH][KW][CoutPad] (CoutPad = 8*ceil(Cout/8)).
struct ConvGenF32 {
  const float* x;
  const float* wt;
  const float* bias;  // [Cout]
  float* y;
  int64_t B;
  int Cin, H, W, Cout, KH, KW, SH, SW;
  int row_kernel = 1;   // 0: never the persistent row-tile kernel (HONK2_F32_RESIDENT=0, read per model handle)
};
int launch_conv_gen_f32(const ConvGenF32& a, cudaStream_t st);

// MaxPool2d((kh,kw)), stride = kernel, floor (cnn.py:85,91).  x [B*C][H][W] -> y [B*C][H/kh][W/kw].
int launch_maxpool_f32(const float* x, float* y, int64_t planes, int H, int W, int kh, int kw,
                       cudaStream_t st);

// Linear: y[M][N] = x[M][K] . w[N][K]^T + bias[N] (cnn.py:95-106).  `scratch` (optional, `scratch_floats` floats, must not
// overlap x or y): room for split-K partial sums when the reduction is long and the output small.
int launch_linear_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int N,
                      int K, float* scratch, size_t scratch_floats, cudaStream_t st);

// ---- bf16 tensor-core path (conv_tc.cu); activations are [B][H][W][48|...] bf16 --------------
struct TcPlan;  // opaque, owned by the model

}  // namespace kws
