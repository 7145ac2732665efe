// neural_af_tc.cu -- placeholder, replaced by the tcgen05 kernel.
#include "common.cuh"
#include "policy.cuh"
namespace mbo {
int tc_pack_weights(mbo_policy* p, cudaStream_t) { p->tc_ready = 0; return MBO_OK; }
int tc_forward(const mbo_policy*, int, int, int, const int32_t*, const mbo_candidates*, const float*, const float*,
               const float*, float*, void*, size_t, cudaStream_t) {
  set_error("tensor-core NeuralAF path not built yet");
  return MBO_ERR_UNSUPPORTED;
}
}
