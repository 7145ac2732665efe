// reduce.cu -- per-episode reductions over the M logits (one CTA per episode).
//
//   argmax        NeuralAF.act, deterministic branch      metabo/policies/policies.py:141-142
//   top-k         np.argsort(-af)[:k] in get_af_maxima    metabo/environment/metabo_gym.py:709,719
//   sample        Categorical(logits=...).sample()        metabo/policies/policies.py:144-146
//   logp/entropy  Categorical.log_prob / .entropy         metabo/policies/policies.py:157-159
//
// Tie rule: the lowest index among equal values wins (torch.argmax's rule; np.argsort's order on
// ties is unspecified in the reference, any deterministic rule is admissible).
// The sampler is Gumbel-max over a counter-based generator keyed (seed, episode, step,
// candidate): no state, reproducible under any sharding of episodes over GPUs, and restated
// bit-for-bit in oracle/sampling_oracle.py for the parity tests.
#include "common.cuh"

namespace mbo {

constexpr int RT = 256;

struct KV {
  float v;
  int i;
};
// a "precedes" b in descending-value / ascending-index order
__device__ __forceinline__ bool precedes(const KV& a, const KV& b) {
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}
__device__ __forceinline__ KV warp_best(KV x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    KV y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    if (precedes(y, x)) x = y;
  }
  return x;
}
__device__ KV block_best(KV x, KV* sh) {
  x = warp_best(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    KV y = l < RT / 32 ? sh[l] : KV{-INFINITY, 0x7fffffff};
    y = warp_best(y);
    if (l == 0) sh[0] = y;
  }
  __syncthreads();
  return sh[0];
}
__device__ double block_sum(double x, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    double y = l < RT / 32 ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
    if (l == 0) sh[0] = y;
  }
  __syncthreads();
  return sh[0];
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ull;
__device__ __forceinline__ float gumbel_noise(uint64_t key, int m) {
  uint64_t r = mix64(key + GOLD * (uint64_t)(m + 1));
  // 23-bit uniform strictly inside (0,1): (k + 0.5) * 2^-23 is exact in fp32 for every k < 2^23 (a 24-bit k + 0.5 is
  // not representable from 2^23 on, and k = 2^24 - 1 rounded to u = 1 -> noise = +inf)
  float u = ((float)(r >> 41) + 0.5f) * (1.0f / 8388608.0f);
  return -logf(-logf(u));
}

// k == 1 doubles as argmax.  noise != 0: Gumbel-max sampling (k must be 1).
__global__ void __launch_bounds__(RT)
topk_kernel(int M, int k, const float* __restrict__ logits, int noise, uint64_t seed, uint64_t step,
            const uint64_t* __restrict__ step_dev, uint64_t episode_offset, int32_t* __restrict__ idx,
            float* __restrict__ val) {
  __shared__ KV sh[RT / 32];
  const int b = blockIdx.x;
  const float* lg = logits + (size_t)b * M;
  uint64_t key = 0;
  if (noise) key = mix64(mix64(seed + GOLD * (episode_offset + (uint64_t)b + 1)) ^ (step + (step_dev ? *step_dev : 0ull) + GOLD));
  KV prev{INFINITY, -1};
  for (int r = 0; r < k; ++r) {
    KV best{-INFINITY, 0x7fffffff};
    for (int m = threadIdx.x; m < M; m += RT) {
      float v = lg[m];
      if (noise) v += gumbel_noise(key, m);
      KV c{v, m};
      if (v == v && precedes(prev, c) && precedes(c, best)) best = c;
    }
    best = block_best(best, sh);
    if (threadIdx.x == 0) {
      int bi = best.i == 0x7fffffff ? 0 : best.i;   // all-NaN row: fall back to index 0
      idx[(size_t)b * k + r] = bi;
      if (val) val[(size_t)b * k + r] = noise ? lg[bi] : best.v;
    }
    prev = best;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RT)
logp_entropy_kernel(int M, const float* __restrict__ logits, const int32_t* __restrict__ actions,
                    float* __restrict__ logp, float* __restrict__ entropy) {
  __shared__ KV sh[RT / 32];
  __shared__ double shd[RT / 32];
  const int b = blockIdx.x;
  const float* lg = logits + (size_t)b * M;
  KV best{-INFINITY, 0x7fffffff};
  for (int m = threadIdx.x; m < M; m += RT) {
    KV c{lg[m], m};
    if (precedes(c, best)) best = c;
  }
  best = block_best(best, sh);
  const float mx = best.v;
  double s = 0.0, w = 0.0;
  for (int m = threadIdx.x; m < M; m += RT) {
    float d = lg[m] - mx;
    float e = expf(d);
    s += (double)e;
    w += (double)(e * d);
  }
  s = block_sum(s, shd);
  w = block_sum(w, shd);
  if (threadIdx.x == 0) {
    double ls = log(s);
    if (entropy) entropy[b] = (float)(ls - w / s);
    if (logp) {
      int a = actions[b];
      a = a < 0 ? 0 : (a >= M ? M - 1 : a);
      logp[b] = (float)((double)(lg[a] - mx) - ls);
    }
  }
}

// Closed-form acquisition functions on the GP posterior (the reference's baselines):
//   EI   metabo/policies/policies.py:236-267   (m - inc) Phi(z) + s phi(z),  z = (m - inc) / s,   0 where s == 0
//   PI   metabo/policies/policies.py:270-307   Phi((m - (inc + xi)) / s),                         0 where s == 0
//   UCB  metabo/policies/policies.py:181-233   m + kappa s;  kappa fixed, or GP-UCB:
//        kappa_t = sqrt(2 log((t + 1)^(D/2 + 2) pi^2 / (3 delta)))
// The reference evaluates them on the float32 state: differences and z are formed in fp32 and widened, pdf / cdf are
// fp64 (scipy.stats.norm); the same here, result stored as fp32 next to the NeuralAF logits so that the top-k /
// argmax kernels serve both.
__global__ void __launch_bounds__(256)
closed_form_af_kernel(int kind, size_t total, int M, const float* __restrict__ mean, const float* __restrict__ std_,
                      const float* __restrict__ epi, double p0, double p1, int D, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / (size_t)M);
  const float m = mean[i], s = std_[i];
  const float inc = epi[b * 4 + 0];
  double r;
  if (kind == MBO_AF_UCB) {
    double kappa = p0;
    if (p1 > 0.0) {                                  // GP-UCB schedule (delta = p1)
      const double t = (double)epi[b * 4 + 2] + 1.0;
      kappa = sqrt(2.0 * log(pow(t, 0.5 * D + 2.0) * 9.869604401089358 / (3.0 * p1)));
    }
    r = (double)m + kappa * (double)s;
  } else if (s == 0.0f) {
    r = 0.0;
  } else if (kind == MBO_AF_EI) {
    const float d = m - inc;
    const double z = (double)(d / s);
    const double pdf = 0.3989422804014327 * exp(-0.5 * z * z);
    const double cdf = 0.5 * erfc(-z * 0.7071067811865476);
    r = (double)d * cdf + (double)s * pdf;
  } else {                                           // PI, xi = p0
    const float d = m - (float)((double)inc + p0);
    const double z = (double)(d / s);
    r = 0.5 * erfc(-z * 0.7071067811865476);
  }
  out[i] = (float)r;
}

}  // namespace mbo

using namespace mbo;

extern "C" int mbo_af_closed_form(int kind, int B, int M, const float* mean, const float* std_, const float* epi,
                                  double p0, double p1, int D, float* af, void* stream) {
  MBO_CHECK_ARG(kind == MBO_AF_EI || kind == MBO_AF_UCB || kind == MBO_AF_PI, "mbo_af_closed_form: unknown kind %d", kind);
  MBO_CHECK_ARG(B > 0 && M > 0 && mean && std_ && epi && af, "mbo_af_closed_form: bad argument");
  MBO_CHECK_ARG(!(kind == MBO_AF_UCB && p1 > 0.0) || D > 0, "mbo_af_closed_form: GP-UCB needs D");
  const size_t total = (size_t)B * M;
  closed_form_af_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, total, M, mean, std_, epi,
                                                                                           p0, p1, D, af);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_logits_argmax(int B, int M, const float* logits, int32_t* action, float* max_logit, void* stream) {
  MBO_CHECK_ARG(B > 0 && M > 0 && logits && action, "mbo_logits_argmax: bad argument");
  topk_kernel<<<B, RT, 0, (cudaStream_t)stream>>>(M, 1, logits, 0, 0, 0, nullptr, 0, action, max_logit);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_logits_topk(int B, int M, int k, const float* logits, int32_t* idx, float* val, void* stream) {
  MBO_CHECK_ARG(B > 0 && M > 0 && logits && idx, "mbo_logits_topk: bad argument");
  MBO_CHECK_ARG(k >= 1 && k <= 8 && k <= M, "mbo_logits_topk: k must be in 1..min(8, M)");
  topk_kernel<<<B, RT, 0, (cudaStream_t)stream>>>(M, k, logits, 0, 0, 0, nullptr, 0, idx, val);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_logits_sample(int B, int M, const float* logits, uint64_t seed, uint64_t step,
                                 const uint64_t* step_dev, uint64_t episode_offset, int32_t* action, void* stream) {
  MBO_CHECK_ARG(B > 0 && M > 0 && logits && action, "mbo_logits_sample: bad argument");
  topk_kernel<<<B, RT, 0, (cudaStream_t)stream>>>(M, 1, logits, 1, seed, step, step_dev, episode_offset, action, nullptr);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_logits_logp_entropy(int B, int M, const float* logits, const int32_t* actions, float* logp,
                                       float* entropy, void* stream) {
  MBO_CHECK_ARG(B > 0 && M > 0 && logits, "mbo_logits_logp_entropy: bad argument");
  MBO_CHECK_ARG(!logp || actions, "mbo_logits_logp_entropy: logp needs actions");
  logp_entropy_kernel<<<B, RT, 0, (cudaStream_t)stream>>>(M, logits, actions, logp, entropy);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}
