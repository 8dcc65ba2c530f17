// Fused MFCC front-end for sm_100a: reflect-pad framing, periodic Hann window, 480-point real
// STFT, power, Slaney mel filterbank, ln, and the reference's degenerate "DCT" (x2), one kernel.
//
// Replaces AudioProcessor.compute_mfccs (/root/reference/utils/audio_processor.py:18-30) for a
// whole batch (the collate loop of data_loader/audio_data_loader.py:26-29).
//
// Algorithm (fp32 throughout): the 480-point real DFT of a windowed frame is computed as a
// 240-point complex FFT of z[m] = x[2m] + i x[2m+1], factored 240 = 16 x 15 (Cooley-Tukey):
// 15 lanes each run a radix-2^4 FFT in registers, twiddle by W240^(n2 k1), exchange through
// shared memory, 16 lanes each run a 15-point DFT (conjugate-pair form, constant twiddles),
// then the real-FFT split recovers bins 0..239, |X|^2, sparse triangular mel weights, 2*ln.
// A warp works on two frames at a time (one per half-warp); a CTA stages the 2880 samples its
// 16 frames cover once (coalesced loads, reflect indexing at the clip edges).
#include "common.cuh"
#include <cmath>
#include <vector>
#include <cstring>

namespace kws {

constexpr int kNfft = 480;
constexpr int kHop = 160;
constexpr int kNz = 240;            // complex FFT length
constexpr int kScStride = kNz + 8;   // per-frame scratch pitch (float2): 496 words = 16 banks off, so the two half-warps (two
                                     // frames) of a warp hit disjoint bank halves in the 32-bit phases (power / mel)
constexpr int kTwA = 16 * 16;        // stage-A twiddles, [k1][lane] (lane-contiguous: no bank conflicts)
constexpr int kFramesPerCta = 16;
constexpr int kMfccThreads = 256;
constexpr int kStageSamples = (kFramesPerCta - 1) * kHop + kNfft;  // 2880
constexpr int kEdgeStageSamples = kFramesPerCta * kNfft;            // 7680: every frame slot stages its own 480 samples
constexpr int kMaxMels = 64;

struct FrontendTables {
  const float* window;    // [480]
  const float2* tw240;    // [16][16] stage-A twiddles W240^(l k1) = (cos, -sin)(2 pi l k1 / 240) at [k1 * 16 + l]
  const float2* tw480;    // [240] (sin, cos)(2 pi k / 480)
  const int* mel_lo;      // [n_mels] first bin
  const int* mel_cnt;     // [n_mels] number of bins
  const int* mel_off;     // [n_mels] offset into mel_w
  const float* mel_w;     // [nnz]
  int n_mels, nnz, nb16;  // nb16: number of 16-bin groups to evaluate (bins 0..16*nb16-1)
};

struct Frontend {
  int sr, n_mels, n_fft, hop;
  float f_min, f_max;
  void* dev_blob = nullptr;
  FrontendTables t{};
  size_t smem_bytes = 0;         // batch kernel
  size_t smem_bytes_edges = 0;   // window-edge kernel of the streaming front-end
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// radix-2 decimation-in-frequency FFT of 16 complex values held in registers; on return
// x[i] holds X[bitrev4(i)].
__device__ __forceinline__ void fft16_dif(float2 (&x)[16]) {
  constexpr float c16[8] = {1.f, 0.923879533f, 0.707106781f, 0.382683432f,
                            0.f, -0.382683432f, -0.707106781f, -0.923879533f};
  constexpr float s16[8] = {0.f, 0.382683432f, 0.707106781f, 0.923879533f,
                            1.f, 0.923879533f, 0.707106781f, 0.382683432f};
#pragma unroll
  for (int len = 16; len >= 2; len >>= 1) {
    const int half = len >> 1;
    const int step = 16 / len;
#pragma unroll
    for (int base = 0; base < 16; base += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        float2 u = x[base + j], v = x[base + j + half];
        x[base + j] = make_float2(u.x + v.x, u.y + v.y);
        float2 t = make_float2(u.x - v.x, u.y - v.y);
        const int tw = j * step;  // W16^tw = (c, -s)
        if (tw == 0) {
          x[base + j + half] = t;
        } else if (tw == 4) {
          x[base + j + half] = make_float2(t.y, -t.x);
        } else {
          x[base + j + half] = make_float2(t.x * c16[tw] + t.y * s16[tw], t.y * c16[tw] - t.x * s16[tw]);
        }
      }
    }
  }
}

__host__ __device__ constexpr int bitrev4(int i) {
  return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
}

// 15-point DFT, Z[k] = sum_n y[n] exp(-2 pi i n k / 15), conjugate-pair form.
__device__ __forceinline__ void dft15(const float2 (&y)[15], float2 (&z)[15]) {
  constexpr float c15[15] = {1.f, 0.913545458f, 0.669130606f, 0.309016994f, -0.104528463f,
                             -0.5f, -0.809016994f, -0.978147601f, -0.978147601f, -0.809016994f,
                             -0.5f, -0.104528463f, 0.309016994f, 0.669130606f, 0.913545458f};
  constexpr float s15[15] = {0.f, 0.406736643f, 0.743144825f, 0.951056516f, 0.994521895f,
                             0.866025404f, 0.587785252f, 0.207911691f, -0.207911691f, -0.587785252f,
                             -0.866025404f, -0.994521895f, -0.951056516f, -0.743144825f, -0.406736643f};
  float2 a[8], b[8];
#pragma unroll
  for (int j = 1; j <= 7; ++j) {
    a[j] = make_float2(y[j].x + y[15 - j].x, y[j].y + y[15 - j].y);
    b[j] = make_float2(y[j].x - y[15 - j].x, y[j].y - y[15 - j].y);
  }
  float2 s0 = y[0];
#pragma unroll
  for (int j = 1; j <= 7; ++j) { s0.x += a[j].x; s0.y += a[j].y; }
  z[0] = s0;
#pragma unroll
  for (int k = 1; k <= 7; ++k) {
    float2 A = y[0], Bv = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 1; j <= 7; ++j) {
      const int idx = (j * k) % 15;
      A.x = fmaf(c15[idx], a[j].x, A.x);
      A.y = fmaf(c15[idx], a[j].y, A.y);
      Bv.x = fmaf(s15[idx], b[j].x, Bv.x);
      Bv.y = fmaf(s15[idx], b[j].y, Bv.y);
    }
    // Z[k] = A - i B ; Z[15-k] = A + i B
    z[k] = make_float2(A.x + Bv.y, A.y - Bv.x);
    z[15 - k] = make_float2(A.x - Bv.y, A.y + Bv.x);
  }
}

// EDGES = false: a CTA computes 16 consecutive frames of one clip; clip b starts at wav + b * wav_stride (wav_stride = N
// for a dense [B, N] batch, = the shift for overlapping windows of a stream).
// EDGES = true (streaming front-end): a CTA computes the four frames of four windows that touch the reflect padding
// (t = 0, 1, T-2, T-1); every other frame of a window is shared with the stream-level frame table.
//
// SMP = float: waveforms as the reference's loader hands them over (librosa float32 in [-1, 1)).  SMP = int16_t: the 16-bit
// PCM samples as they sit in the wav files (dataset/gsc_dataset.py:169 loads them through librosa, which returns exactly
// s / 32768): converted while staging, bit-identical features, half the bytes from the host and from HBM.
__device__ __forceinline__ float smp_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float smp_ld(const int16_t* p) { return (float)__ldg(p) * (1.f / 32768.f); }

template <bool EDGES, typename SMP>
__global__ void __launch_bounds__(kMfccThreads)
mfcc_kernel(FrontendTables tb, const SMP* __restrict__ wav, int64_t wav_stride, int64_t B, int N, int T,
            int tiles_per_utt, float* __restrict__ feat) {
  constexpr int kStage = EDGES ? kEdgeStageSamples : kStageSamples;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* scratch = reinterpret_cast<float2*>(smem_raw);                 // [16][248]
  float2* s_tw240 = scratch + kFramesPerCta * kScStride;                 // [16][16]
  float2* s_tw480 = s_tw240 + kTwA;                                      // [240]
  float* s_wave = reinterpret_cast<float*>(s_tw480 + kNz);               // [2880] ([16][480] for EDGES)
  float* s_win = s_wave + kStage;                                        // [480]
  int* s_lo = reinterpret_cast<int*>(s_win + kNfft);                     // [64]
  int* s_cnt = s_lo + kMaxMels;
  int* s_off = s_cnt + kMaxMels;
  float* s_w = reinterpret_cast<float*>(s_off + kMaxMels);               // [nnz]

  const int tid = threadIdx.x;
  const int64_t b = EDGES ? (int64_t)blockIdx.x * 4 : blockIdx.x / tiles_per_utt;   // (EDGES: first of four windows)
  const int t0 = EDGES ? 0 : (blockIdx.x % tiles_per_utt) * kFramesPerCta;
  const SMP* w = wav + b * wav_stride;

  if constexpr (EDGES) {
    // frame slot s = 4 * (window - b) + e, e -> frame t = 0, 1, T-2, T-1: samples [160 t - 240, +480) of its window
    for (int i = tid; i < kEdgeStageSamples; i += kMfccThreads) {
      const int slot = i / kNfft, n = i - slot * kNfft;
      const int wi = slot >> 2, e = slot & 3;
      const int te = e < 2 ? e : T - 4 + e;
      int j = te * kHop - kNfft / 2 + n;
      if (j < 0) j = -j;
      if (j >= N) j = 2 * (N - 1) - j;
      j = max(0, min(j, N - 1));
      s_wave[i] = b + wi < B ? smp_ld(w + (int64_t)wi * wav_stride + j) : 0.f;
    }
  } else {
    // stage: samples [160*t0 - 240, +2880) with librosa 'reflect' padding at the clip edges; only the frames that exist
    // (the last tile of a clip usually holds fewer than 16).  Interior runs are copied 16 bytes at a time.
    const int j0 = t0 * kHop - kNfft / 2;
    const int n_fr = min(kFramesPerCta, T - t0);
    const int n_stage = (n_fr - 1) * kHop + kNfft;
    const bool vec = ((reinterpret_cast<uintptr_t>(w) & 15) == 0);   // (j0 is a multiple of 16 samples)
    constexpr int V = 16 / (int)sizeof(SMP);   // samples per 16-byte load (n_stage is a multiple of 8)
    if (vec) {
      for (int iv = tid; iv < n_stage / V; iv += kMfccThreads) {
        const int i = iv * V, j = j0 + i;
        float e[V];
        if (j >= 0 && j + V - 1 < N) {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(w + j));
          if constexpr (sizeof(SMP) == 4) {
            e[0] = __uint_as_float(raw.x); e[1] = __uint_as_float(raw.y); e[2] = __uint_as_float(raw.z); e[3] = __uint_as_float(raw.w);
          } else {
            const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              e[2 * q] = (float)(int16_t)(r[q] & 0xFFFFu) * (1.f / 32768.f);
              e[2 * q + 1] = (float)(int16_t)(r[q] >> 16) * (1.f / 32768.f);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < V; ++q) {
            int jj = j + q;
            if (jj < 0) jj = -jj;
            if (jj >= N) jj = 2 * (N - 1) - jj;
            jj = max(0, min(jj, N - 1));  // only reachable for samples of masked frames
            e[q] = smp_ld(w + jj);
          }
        }
#pragma unroll
        for (int q = 0; q < V; q += 4)
          *reinterpret_cast<float4*>(s_wave + i + q) = make_float4(e[q], e[q + 1], e[q + 2], e[q + 3]);
      }
    } else {
      for (int i = tid; i < n_stage; i += kMfccThreads) {
        int j = j0 + i;
        if (j < 0) j = -j;
        if (j >= N) j = 2 * (N - 1) - j;
        j = max(0, min(j, N - 1));
        s_wave[i] = smp_ld(w + j);
      }
    }
  }
  for (int i = tid; i < kNfft; i += kMfccThreads) s_win[i] = tb.window[i];
  for (int i = tid; i < kNz; i += kMfccThreads) s_tw480[i] = tb.tw480[i];
  for (int i = tid; i < kTwA; i += kMfccThreads) s_tw240[i] = tb.tw240[i];
  for (int i = tid; i < tb.n_mels; i += kMfccThreads) {
    s_lo[i] = tb.mel_lo[i]; s_cnt[i] = tb.mel_cnt[i]; s_off[i] = tb.mel_off[i];
  }
  for (int i = tid; i < tb.nnz; i += kMfccThreads) s_w[i] = tb.mel_w[i];
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int half = lane >> 4, l = lane & 15;
  const int t_local = 2 * warp + half;
  const int t = EDGES ? ((t_local & 3) < 2 ? (t_local & 3) : T - 4 + (t_local & 3)) : t0 + t_local;
  const int64_t b_out = EDGES ? b + (t_local >> 2) : b;
  // (whole warps without a frame -- the tail of a clip's last tile -- are done: only __syncwarp from here on)
  if (!EDGES && t0 + 2 * warp >= T) return;
  float2* sc = scratch + t_local * kScStride;
  const float* fr = s_wave + (EDGES ? kNfft : kHop) * t_local;

  // ---- stage A: 15 radix-16 FFTs over n1 (lane = n2), twiddle by W240^(n2 k1)
  if (l < 15) {
    float2 x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int m = 15 * n1 + l;
      float2 s = *reinterpret_cast<const float2*>(fr + 2 * m);
      float2 wv = *reinterpret_cast<const float2*>(s_win + 2 * m);
      x[n1] = make_float2(s.x * wv.x, s.y * wv.y);
    }
    fft16_dif(x);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k1 = bitrev4(i);
      sc[k1 * 15 + l] = cmul(x[i], s_tw240[k1 * 16 + l]);
    }
  }
  __syncwarp();

  // ---- stage B: 16 DFT-15 over n2 (lane = k1): Z[k1 + 16 k2] stays in this lane's registers
  float2 z[15];
  {
    float2 y[15];
#pragma unroll
    for (int n2 = 0; n2 < 15; ++n2) y[n2] = sc[l * 15 + n2];
    dft15(y, z);
  }

  // ---- real-FFT split + power: X[k] = E + (-sin, -cos)(theta_k) * O for k = l + 16 j.  Z[k] is this lane's z[j];
  // its mirror Z[240 - k] = Z[(16 - l) + 16 (14 - j)] is lane 16 - l's z[14 - j] (one shuffle inside the half-warp),
  // or, for lane 0, its own z[(15 - j) mod 15]: no trip through shared memory.
  float p[15];
  const int mirror = (16 - l) & 15;
#pragma unroll
  for (int j = 0; j < 15; ++j) {
    p[j] = 0.f;
    if (j < tb.nb16) {
      const int k = l + 16 * j;
      const float2 zk = z[j];
      float2 zm;
      zm.x = __shfl_sync(0xffffffffu, z[14 - j].x, mirror, 16);
      zm.y = __shfl_sync(0xffffffffu, z[14 - j].y, mirror, 16);
      if (l == 0) zm = z[(15 - j) % 15];
      const float2 tw = s_tw480[k];  // (sin, cos)
      const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);
      const float orr = 0.5f * (zk.x - zm.x), oi = 0.5f * (zk.y + zm.y);
      const float xr = er - tw.x * orr + tw.y * oi;
      const float xi = ei - tw.x * oi - tw.y * orr;
      p[j] = xr * xr + xi * xi;
    }
  }
  __syncwarp();
  float* P = reinterpret_cast<float*>(sc);
#pragma unroll
  for (int j = 0; j < 15; ++j)
    if (j < tb.nb16) P[l + 16 * j] = p[j];
  __syncwarp();

  // ---- sparse mel filterbank, ln, x2 (reference "DCT" of a length-1 axis)
  if (t < T && b_out < B) {
    float* out = feat + (b_out * (int64_t)T + t) * tb.n_mels;
    for (int m = l; m < tb.n_mels; m += 16) {
      const int lo = s_lo[m], cnt = s_cnt[m], off = s_off[m];
      float acc = 0.f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(s_w[off + i], P[lo + i], acc);
      out[m] = acc > 0.f ? 2.f * logf(acc) : 2.f * acc;
    }
  }
}

// Streaming front-end: interior frames of window k (t = 2 .. T-3) are rows (k * m + t) of the stream-level frame
// table S (m = shift / hop).  Pure copy, 16 bytes per thread when the row length allows.
template <typename V>
__global__ void __launch_bounds__(256)
mfcc_stream_gather_kernel(const V* __restrict__ S, int64_t K, int T, int m, int row_v, V* __restrict__ feat) {
  const int64_t per_win = (int64_t)(T - 4) * row_v;
  const int64_t total = K * per_win;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / per_win;
    const int64_t r = i - k * per_win;   // (t - 2) * row_v + c
    feat[k * (int64_t)T * row_v + 2 * row_v + r] = __ldg(S + (k * m + 2) * (int64_t)row_v + r);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: table construction (double precision, float32 rounding where numpy rounds)

static double hz_to_mel_slaney(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz_slaney(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

}  // namespace kws

using namespace kws;

struct kws_frontend : kws::Frontend {};

extern "C" int kws_frontend_create(int sr, int n_mels, float f_min, float f_max, int n_fft, int hop,
                                   kws_frontend_t** out) {
  KWS_REQUIRE(out != nullptr, "kws_frontend_create: out is null");
  *out = nullptr;
  KWS_REQUIRE(n_fft == kNfft && hop == kHop,
              "kws_frontend_create: only n_fft=480 / hop=160 is built (got %d / %d)", n_fft, hop);
  KWS_REQUIRE(n_mels >= 1 && n_mels <= kMaxMels, "kws_frontend_create: n_mels must be in [1,%d]", kMaxMels);
  KWS_REQUIRE(sr > 0 && f_min >= 0 && f_max > f_min && f_max <= sr / 2.0f,
              "kws_frontend_create: need 0 <= f_min < f_max <= sr/2");
  int major = 0, minor = 0, sms = 0;
  KWS_TRY(kws_device_info(&major, &minor, &sms));

  const int n_bins = 1 + n_fft / 2;
  // librosa.filters.mel(htk=False, norm='slaney'), float32 storage like numpy
  std::vector<double> mel_f(n_mels + 2);
  const double m_lo = hz_to_mel_slaney(f_min), m_hi = hz_to_mel_slaney(f_max);
  for (int i = 0; i < n_mels + 2; ++i)
    mel_f[i] = mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  std::vector<int> lo(n_mels), cnt(n_mels), off(n_mels);
  std::vector<float> wts;
  int max_bin = 0;
  for (int i = 0; i < n_mels; ++i) {
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    std::vector<float> row(n_bins);
    for (int k = 0; k < n_bins; ++k) {
      const double f = (double)sr / 2 * k / (n_bins - 1);
      const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
      const double upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
      const float tri = (float)std::fmax(0.0, std::fmin(lower, upper));
      row[k] = (float)((double)tri * enorm);
      if (row[k] != 0.f) { if (first < 0) first = k; last = k; }
    }
    off[i] = (int)wts.size();
    if (first < 0) { lo[i] = 0; cnt[i] = 0; continue; }
    lo[i] = first; cnt[i] = last - first + 1;
    for (int k = first; k <= last; ++k) wts.push_back(row[k]);
    if (last > max_bin) max_bin = last;
  }
  KWS_REQUIRE(max_bin < kNz, "kws_frontend_create: filterbank touches the Nyquist bin (unsupported)");
  const int nb16 = max_bin / 16 + 1;

  const double PI = 3.14159265358979323846;
  std::vector<float> window(kNfft);
  for (int n = 0; n < kNfft; ++n) window[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / kNfft));
  std::vector<float2> tw240(kTwA), tw480(kNz);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int l = 0; l < 16; ++l) {
      const int j = (l * k1) % kNz;
      tw240[k1 * 16 + l] = make_float2((float)std::cos(2.0 * PI * j / kNz), (float)(-std::sin(2.0 * PI * j / kNz)));
    }
  for (int j = 0; j < kNz; ++j)
    tw480[j] = make_float2((float)std::sin(2.0 * PI * j / kNfft), (float)std::cos(2.0 * PI * j / kNfft));

  // one device blob
  const size_t nnz = wts.size();
  size_t o_win = 0, o_tw240 = o_win + sizeof(float) * kNfft, o_tw480 = o_tw240 + sizeof(float2) * kTwA,
         o_lo = o_tw480 + sizeof(float2) * kNz, o_cnt = o_lo + sizeof(int) * n_mels,
         o_off = o_cnt + sizeof(int) * n_mels, o_w = o_off + sizeof(int) * n_mels,
         total = o_w + sizeof(float) * (nnz + 1);
  std::vector<unsigned char> host(total);
  memcpy(host.data() + o_win, window.data(), sizeof(float) * kNfft);
  memcpy(host.data() + o_tw240, tw240.data(), sizeof(float2) * kTwA);
  memcpy(host.data() + o_tw480, tw480.data(), sizeof(float2) * kNz);
  memcpy(host.data() + o_lo, lo.data(), sizeof(int) * n_mels);
  memcpy(host.data() + o_cnt, cnt.data(), sizeof(int) * n_mels);
  memcpy(host.data() + o_off, off.data(), sizeof(int) * n_mels);
  if (nnz) memcpy(host.data() + o_w, wts.data(), sizeof(float) * nnz);

  kws_frontend* fe = new kws_frontend();
  fe->sr = sr; fe->n_mels = n_mels; fe->n_fft = n_fft; fe->hop = hop; fe->f_min = f_min; fe->f_max = f_max;
  cudaError_t e = cudaMalloc(&fe->dev_blob, total);
  if (e == cudaSuccess) e = cudaMemcpy(fe->dev_blob, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("kws_frontend_create: table upload failed: %s", cudaGetErrorString(e));
    if (fe->dev_blob) cudaFree(fe->dev_blob);
    delete fe;
    return KWS_ERR_CUDA;
  }
  unsigned char* d = static_cast<unsigned char*>(fe->dev_blob);
  fe->t.window = reinterpret_cast<const float*>(d + o_win);
  fe->t.tw240 = reinterpret_cast<const float2*>(d + o_tw240);
  fe->t.tw480 = reinterpret_cast<const float2*>(d + o_tw480);
  fe->t.mel_lo = reinterpret_cast<const int*>(d + o_lo);
  fe->t.mel_cnt = reinterpret_cast<const int*>(d + o_cnt);
  fe->t.mel_off = reinterpret_cast<const int*>(d + o_off);
  fe->t.mel_w = reinterpret_cast<const float*>(d + o_w);
  fe->t.n_mels = n_mels; fe->t.nnz = (int)nnz; fe->t.nb16 = nb16;
  fe->smem_bytes = sizeof(float2) * (kFramesPerCta * kScStride + kTwA + kNz) +
                   sizeof(float) * (kStageSamples + kNfft) + sizeof(int) * 3 * kMaxMels +
                   sizeof(float) * (nnz + 1);
  fe->smem_bytes_edges = fe->smem_bytes + sizeof(float) * (kEdgeStageSamples - kStageSamples);
  e = cudaFuncSetAttribute(mfcc_kernel<false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mfcc_kernel<false, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mfcc_kernel<true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->smem_bytes_edges);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mfcc_kernel<true, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->smem_bytes_edges);
  if (e != cudaSuccess) {
    set_error("kws_frontend_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    cudaFree(fe->dev_blob);
    delete fe;
    return KWS_ERR_CUDA;
  }
  *out = fe;
  return KWS_OK;
}

extern "C" void kws_frontend_destroy(kws_frontend_t* fe) {
  if (!fe) return;
  if (fe->dev_blob) cudaFree(fe->dev_blob);
  delete fe;
}

extern "C" int kws_frontend_n_frames(const kws_frontend_t* fe, int n_samples) {
  if (!fe || n_samples < 0) return 0;
  return 1 + n_samples / fe->hop;
}

extern "C" int kws_frontend_n_mels(const kws_frontend_t* fe) { return fe ? fe->n_mels : 0; }

template <typename SMP>
static int mfcc_forward_any(const char* who, const kws_frontend_t* fe, const SMP* wav, int64_t B, int n_samples, float* feat,
                            void* stream) {
  KWS_REQUIRE(fe != nullptr, "%s: frontend is null", who);
  KWS_REQUIRE(B >= 0, "%s: negative batch", who);
  // librosa's reflect padding of n_fft/2 needs more than n_fft/2 samples
  KWS_REQUIRE(n_samples > kNfft / 2, "%s: need more than %d samples per clip (got %d)", who, kNfft / 2, n_samples);
  if (B == 0) return KWS_OK;
  KWS_REQUIRE(wav != nullptr && feat != nullptr, "%s: null buffer", who);
  const int T = 1 + n_samples / kHop;
  const int tiles = ceil_div(T, kFramesPerCta);
  const int64_t blocks = B * tiles;
  KWS_REQUIRE(blocks < (int64_t)2147483647, "%s: batch too large for one launch", who);
  mfcc_kernel<false, SMP><<<(unsigned)blocks, kMfccThreads, fe->smem_bytes, as_stream(stream)>>>(
      fe->t, wav, (int64_t)n_samples, B, n_samples, T, tiles, feat);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

extern "C" int kws_mfcc_forward(const kws_frontend_t* fe, const float* wav, int64_t B, int n_samples,
                                float* feat, void* stream) {
  return mfcc_forward_any<float>("kws_mfcc_forward", fe, wav, B, n_samples, feat, stream);
}

extern "C" int kws_mfcc_forward_pcm16(const kws_frontend_t* fe, const int16_t* wav, int64_t B, int n_samples,
                                      float* feat, void* stream) {
  return mfcc_forward_any<int16_t>("kws_mfcc_forward_pcm16", fe, wav, B, n_samples, feat, stream);
}

extern "C" size_t kws_mfcc_stream_scratch_bytes(const kws_frontend_t* fe, int64_t n_windows, int window, int shift) {
  if (!fe || n_windows < 1 || window < 1 || shift < 1) return 0;
  const int T = 1 + window / kHop;
  if (shift % kHop != 0 || T < 5) return 0;   // no frame is shared: computed straight into the output
  const int64_t span = (n_windows - 1) * (int64_t)shift + window;
  return (size_t)(1 + span / kHop) * fe->n_mels * sizeof(float);
}

template <typename SMP>
static int mfcc_stream_forward_any(const kws_frontend_t* fe, const SMP* wav, int64_t n_windows, int window,
                                   int shift, float* feat, void* scratch, size_t scratch_bytes, void* stream) {
  KWS_REQUIRE(fe != nullptr, "kws_mfcc_stream_forward: frontend is null");
  KWS_REQUIRE(n_windows >= 0, "kws_mfcc_stream_forward: negative window count");
  KWS_REQUIRE(window > kNfft / 2, "kws_mfcc_stream_forward: need more than %d samples per window (got %d)",
              kNfft / 2, window);
  KWS_REQUIRE(shift >= 1, "kws_mfcc_stream_forward: shift must be positive (got %d)", shift);
  if (n_windows == 0) return KWS_OK;
  KWS_REQUIRE(wav != nullptr && feat != nullptr, "kws_mfcc_stream_forward: null buffer");
  cudaStream_t st = as_stream(stream);
  const int T = 1 + window / kHop;
  const int tiles = ceil_div(T, kFramesPerCta);
  const int64_t span = (n_windows - 1) * (int64_t)shift + window;   // samples the windows cover
  if (shift % kHop != 0 || T < 5) {
    // windows do not share frames (or have no interior frame): every window in full, read in place from the stream
    const int64_t blocks = n_windows * tiles;
    KWS_REQUIRE(blocks < (int64_t)2147483647, "kws_mfcc_stream_forward: too many windows for one launch");
    mfcc_kernel<false, SMP><<<(unsigned)blocks, kMfccThreads, fe->smem_bytes, st>>>(
        fe->t, wav, (int64_t)shift, n_windows, window, T, tiles, feat);
    KWS_CHECK_LAUNCH();
    return KWS_OK;
  }
  KWS_REQUIRE(span < (int64_t)2147483647, "kws_mfcc_stream_forward: the windows cover %lld samples; split the call",
              (long long)span);
  const size_t need = kws_mfcc_stream_scratch_bytes(fe, n_windows, window, shift);
  KWS_REQUIRE(scratch != nullptr && scratch_bytes >= need,
              "kws_mfcc_stream_forward: needs %zu bytes of scratch, got %zu", need, scratch_bytes);
  // 1) the frame table of the whole span as ONE clip: rows 2 .. J-3 do not touch its reflect padding
  float* S = static_cast<float*>(scratch);
  const int J = 1 + (int)(span / kHop);
  mfcc_kernel<false, SMP><<<(unsigned)ceil_div(J, kFramesPerCta), kMfccThreads, fe->smem_bytes, st>>>(
      fe->t, wav, span, (int64_t)1, (int)span, J, ceil_div(J, kFramesPerCta), S);
  KWS_CHECK_LAUNCH();
  // 2) the four frames per window that do (reflect padding at the WINDOW's edges, audio_processor.py:19-26 per window)
  mfcc_kernel<true, SMP><<<(unsigned)ceil_div<int64_t>(n_windows, 4), kMfccThreads, fe->smem_bytes_edges, st>>>(
      fe->t, wav, (int64_t)shift, n_windows, window, T, 1, feat);
  KWS_CHECK_LAUNCH();
  // 3) interior frames: copies of table rows
  const int m = shift / kHop;
  const bool vec = fe->n_mels % 4 == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;
  const int row_v = vec ? fe->n_mels / 4 : fe->n_mels;
  const int64_t total = n_windows * (int64_t)(T - 4) * row_v;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(total, 256), 148 * 16);
  if (vec)
    mfcc_stream_gather_kernel<float4><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(S), n_windows, T, m, row_v,
                                                            reinterpret_cast<float4*>(feat));
  else
    mfcc_stream_gather_kernel<float><<<grid, 256, 0, st>>>(S, n_windows, T, m, row_v, feat);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

extern "C" int kws_mfcc_stream_forward(const kws_frontend_t* fe, const float* wav, int64_t n_windows, int window,
                                       int shift, float* feat, void* scratch, size_t scratch_bytes, void* stream) {
  return mfcc_stream_forward_any<float>(fe, wav, n_windows, window, shift, feat, scratch, scratch_bytes, stream);
}

extern "C" int kws_mfcc_stream_forward_pcm16(const kws_frontend_t* fe, const int16_t* wav, int64_t n_windows, int window,
                                             int shift, float* feat, void* scratch, size_t scratch_bytes, void* stream) {
  return mfcc_stream_forward_any<int16_t>(fe, wav, n_windows, window, shift, feat, scratch, scratch_bytes, stream);
}
