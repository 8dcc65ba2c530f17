// Library-wide plumbing of the honk2_b200 C ABI: error text, device probe, and the on-device
// accuracy counter that replaces Acc.accumulate (/root/reference/metric/acc.py:14-24).
#include "common.cuh"

namespace kws {

static thread_local char g_error[1024] = "";
thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// One thread per utterance: first-max argmax (torch.argmax tie rule), warp-aggregated count.
__global__ void __launch_bounds__(256)
acc_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t B,
           int n_labels, unsigned long long* __restrict__ counts, int64_t* __restrict__ pred) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (b < B) {
    const float* row = logits + b * n_labels;
    float best = row[0];
    int arg = 0;
    for (int j = 1; j < n_labels; ++j) {
      const float v = row[j];
      if (v > best || (v != v && best == best)) { best = v; arg = j; }  // NaN wins like torch
    }
    if (pred) pred[b] = arg;
    hit = (target[b] == (int64_t)arg);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, hit);
  __shared__ int s_hits;
  if (threadIdx.x == 0) s_hits = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&s_hits, __popc(ballot));
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    const int64_t n = min((int64_t)blockDim.x, B - first);
    if (s_hits) atomicAdd(&counts[0], (unsigned long long)s_hits);
    atomicAdd(&counts[1], (unsigned long long)n);
  }
}

// Evaluation statistics of one batch in one pass (run/test.py:28-33): per utterance the first-max argmax, then
//   counts[0] += hit, counts[1] += 1                                   (metric/acc.py:14-24)
//   class_counts[2 t] += 1, class_counts[2 t + 1] += hit  (t = target)  (metric/per_class_acc.py:14-45)
//   loss_sum += logsumexp(row) - row[t]                                (loss_function.py:7-9, CrossEntropyLoss, summed)
// Block-aggregated in shared memory, one set of global atomics per block; targets outside [0, n_labels) count as
// misses and contribute no loss term.
__global__ void __launch_bounds__(256)
eval_stats_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int64_t B, int n_labels,
                  unsigned long long* __restrict__ counts, unsigned long long* __restrict__ class_counts,
                  double* __restrict__ loss_sum, int64_t* __restrict__ pred) {
  extern __shared__ unsigned int s_cls[];   // [2 * n_labels] when class_counts != nullptr
  __shared__ int s_hits;
  __shared__ double s_loss[8];
  if (class_counts != nullptr)
    for (int i = threadIdx.x; i < 2 * n_labels; i += blockDim.x) s_cls[i] = 0u;
  if (threadIdx.x == 0) s_hits = 0;
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  double loss = 0.0;
  if (b < B) {
    const float* row = logits + b * n_labels;
    float best = row[0];
    int arg = 0;
    for (int j = 1; j < n_labels; ++j) {
      const float v = row[j];
      if (v > best || (v != v && best == best)) { best = v; arg = j; }  // NaN wins like torch
    }
    if (pred) pred[b] = arg;
    const int64_t t = target[b];
    hit = (t == (int64_t)arg);
    if (t >= 0 && t < n_labels) {
      if (class_counts != nullptr) {
        atomicAdd(&s_cls[2 * t], 1u);
        if (hit) atomicAdd(&s_cls[2 * t + 1], 1u);
      }
      if (loss_sum != nullptr) {
        float sum = 0.f;
        for (int j = 0; j < n_labels; ++j) sum += expf(row[j] - best);
        loss = (double)(logf(sum) + best - row[t]);
      }
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(&s_hits, __popc(ballot));
  if (loss_sum != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = loss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    const int64_t n = min((int64_t)blockDim.x, B - first);
    if (counts != nullptr) {
      if (s_hits) atomicAdd(&counts[0], (unsigned long long)s_hits);
      atomicAdd(&counts[1], (unsigned long long)n);
    }
    if (loss_sum != nullptr) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += s_loss[w];
      atomicAdd(loss_sum, tot);
    }
  }
  if (class_counts != nullptr)
    for (int i = threadIdx.x; i < 2 * n_labels; i += blockDim.x)
      if (s_cls[i]) atomicAdd(&class_counts[i], (unsigned long long)s_cls[i]);
}

}  // namespace kws

using namespace kws;

extern "C" int kws_abi_version(void) { return KWS_ABI_VERSION; }

extern "C" const char* kws_last_error(void) { return kws::g_error; }

extern "C" int kws_device_info(int* cc_major, int* cc_minor, int* n_sms) {
  int dev = 0;
  KWS_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0, sms = 0;
  KWS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  KWS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  KWS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  if (n_sms) *n_sms = sms;
  if (major != 10) {
    set_error("honk2_b200 is built for sm_100a only; device %d is compute capability %d.%d "
              "(there is no fallback path)", dev, major, minor);
    return KWS_ERR_UNSUPPORTED;
  }
  return KWS_OK;
}

extern "C" int kws_acc_accumulate(const float* logits, const int64_t* target, int64_t B,
                                  int n_labels, int64_t* counts, int64_t* pred, void* stream) {
  KWS_REQUIRE(B >= 0 && n_labels >= 1, "kws_acc_accumulate: bad shape B=%lld n_labels=%d",
              (long long)B, n_labels);
  KWS_REQUIRE(counts != nullptr, "kws_acc_accumulate: counts is null");
  if (B == 0) return KWS_OK;
  KWS_REQUIRE(logits != nullptr && target != nullptr, "kws_acc_accumulate: null buffer");
  const int64_t blocks = ceil_div<int64_t>(B, 256);
  acc_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      logits, target, B, n_labels, reinterpret_cast<unsigned long long*>(counts), pred);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}

extern "C" int kws_eval_accumulate(const float* logits, const int64_t* target, int64_t B, int n_labels,
                                   int64_t* counts, int64_t* class_counts, double* loss_sum, int64_t* pred,
                                   void* stream) {
  KWS_REQUIRE(B >= 0 && n_labels >= 1 && n_labels <= 4096, "kws_eval_accumulate: bad shape B=%lld n_labels=%d",
              (long long)B, n_labels);
  if (B == 0) return KWS_OK;
  KWS_REQUIRE(logits != nullptr && target != nullptr, "kws_eval_accumulate: null buffer");
  const int64_t blocks = ceil_div<int64_t>(B, 256);
  const size_t smem = class_counts != nullptr ? sizeof(unsigned int) * 2 * n_labels : 0;
  eval_stats_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(
      logits, target, B, n_labels, reinterpret_cast<unsigned long long*>(counts),
      reinterpret_cast<unsigned long long*>(class_counts), loss_sum, pred);
  KWS_CHECK_LAUNCH();
  return KWS_OK;
}
