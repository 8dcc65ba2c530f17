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
