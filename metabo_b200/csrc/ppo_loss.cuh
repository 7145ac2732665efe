// ppo_loss.cuh -- per-episode PPO loss terms and dlogits (policies.py:150-161, ppo.py:148-168), shared by the tensor-core
// update path (ppo_update.cu) and the CUDA-core path for small / tanh nets (ppo_update_simt.cu).
#pragma once
#include "common.cuh"

namespace mbo {
namespace upd {

// ---- loss terms and dlogits, one CTA per episode (policies.py:150-161, ppo.py:148-168) -----------------------------
static __device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += s_red[i];      // same order in every thread
  return t;
}
static __global__ void __launch_bounds__(256) upd_loss_kernel(int M, const float* __restrict__ logits,
                                                        const int32_t* __restrict__ actions, const float* __restrict__ advs,
                                                        const float* __restrict__ logp_old, float eps, float ent_coeff,
                                                        float inv_n, float* __restrict__ dl, float* __restrict__ terms) {
  __shared__ double s_red[8];
  __shared__ float s_max[8];
  const int e = blockIdx.x, tid = threadIdx.x;
  const float* l = logits + (size_t)e * M;
  float mx = -INFINITY;
  for (int i = tid; i < M; i += 256) mx = fmaxf(mx, l[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) s_max[tid >> 5] = mx;
  __syncthreads();
  mx = s_max[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_max[i]);
  double se = 0.0;
  for (int i = tid; i < M; i += 256) se += (double)expf(l[i] - mx);
  se = block_sum(se, s_red);
  const float lse = mx + (float)log(se);
  double pe = 0.0;                                  // sum p log p
  for (int i = tid; i < M; i += 256) { const float lp = l[i] - lse; pe += (double)(expf(lp) * lp); }
  pe = block_sum(pe, s_red);
  const float ent = (float)(-pe);
  int act = actions[e];
  act = act < 0 ? 0 : (act >= M ? M - 1 : act);
  const float lpa = l[act] - lse;
  const float adv = advs ? advs[e] : 0.f;
  const float ratio = expf(lpa - (logp_old ? logp_old[e] : lpa));
  const float clipped = fminf(fmaxf(ratio, 1.f - eps), 1.f + eps);
  const float s1 = ratio * adv, s2 = clipped * adv;
  // d(-min(s1, s2) / n) / d ratio: torch.min splits a tie evenly and clamp passes the gradient inside its range, so
  // the unclipped and the tie case both give -adv / n; when the clipped branch is the smaller one the gradient is 0
  const float g_lp = (s1 <= s2 ? -adv * inv_n : 0.f) * ratio;
  const float g_ent = -ent_coeff * inv_n;
  for (int i = tid; i < M; i += 256) {
    const float lp = l[i] - lse, p = expf(lp);
    dl[(size_t)e * M + i] = g_lp * ((i == act ? 1.f : 0.f) - p) - g_ent * p * (lp + ent);
  }
  if (tid == 0) {
    terms[e * 4 + 0] = lpa;
    terms[e * 4 + 1] = ent;
    terms[e * 4 + 2] = -fminf(s1, s2) * inv_n;
    terms[e * 4 + 3] = ratio;
  }
}

}  // namespace upd
}  // namespace mbo
