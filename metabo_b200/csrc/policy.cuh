// policy.cuh -- the opaque mbo_policy handle (device copies of one MLP's parameters).
#pragma once
#include "common.cuh"

struct mbo_policy {
  int n_layers = 0;
  int widths[MBO_MAX_LAYERS + 1] = {0};
  int activation = MBO_ACT_RELU;
  float* d_Wt[MBO_MAX_LAYERS] = {nullptr};   // [in][out] fp32, for the CUDA-core kernel
  float* d_b[MBO_MAX_LAYERS] = {nullptr};    // [out] fp32
  float* d_W_raw = nullptr;                  // all layers, torch layout [out][in], concatenated
  size_t raw_off[MBO_MAX_LAYERS] = {0};
  uint64_t alloc_sig = 0;
  // tensor-core pack (neural_af_tc.cu): UMMA-canonical K-chunked tiles, see that file
  void* d_tc = nullptr;
  size_t tc_bytes = 0;
  int tc_ready = 0;       // 1 when the architecture is F->H->H->H->H->1 ReLU with H <= 208
  int tc_f16 = 0;         // 1 when additionally H <= 206 (MBO_MLP_FP16 folds the bias into two spare K columns)
  int tc_H = 0, tc_F = 0;
};

namespace mbo {
// neural_af_tc.cu: the tcgen05 policy forward on raw weights, used by the PPO update (ppo_update.cu)
size_t tc_pack_bytes();
int tc_pack_raw_bf16x3(const float* const* W, const float* const* b, int F, int H, void* pack, cudaStream_t s);
int tc_forward_dense_bf16x3(const void* pack, int E, int M, int F_state, int F, const int32_t* cols, const float* states,
                            float* logits, int* err_flag, float* const* h_out, uint32_t* const* bits_out, int H,
                            cudaStream_t s);
}  // namespace mbo
