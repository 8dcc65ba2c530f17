// placeholder until the tcgen05 path lands (replaced in the next commit)
#include "tc.cuh"
namespace kws {
struct TcResNet { kws_resnet_config cfg; };
int tc_resnet_create(const kws_resnet_config& cfg, TcResNet** out) { *out = new TcResNet{cfg}; return KWS_OK; }
void tc_resnet_destroy(TcResNet* p) { delete p; }
int tc_resnet_set_weights(TcResNet*, const kws_resnet_weights&, float* const*, float* const*, cudaStream_t) { return KWS_OK; }
size_t tc_resnet_workspace_bytes(const TcResNet*, int64_t, int, int, int) { return 0; }
int tc_resnet_forward(TcResNet*, const float*, int64_t, int, int, float*, void*, size_t, int, LaunchProfiler*, cudaStream_t) {
  set_error("bf16 tensor-core path is not built in this revision");
  return KWS_ERR_UNSUPPORTED;
}
}  // namespace kws
