// env.cu -- device-resident environment step of the analytic objective families.
//
// One launch replaces the ~40 elementwise torch kernels a BO step of `VecMetaBO` used to enqueue between two
// AF-optimisation rounds (at the reference's batch of 40 episodes a rollout step is pure launch latency):
//   metabo/environment/objectives.py:26-80   Branin (normalised, negated, translated / scaled variant)
//   metabo/environment/objectives.py:165-219 Hartmann-3 (likewise)
//   metabo/environment/metabo_gym.py:195-217 step: evaluate f(x), append (x, y) to the episode's data
//   metabo/environment/metabo_gym.py:679-704 reward from the simple regret (none / neg_linear / neg_log10)
//   metabo/environment/metabo_gym.py:723-730 incumbent, :642-671 the per-episode state scalars of the NEXT state
// Arithmetic: fp64, one rounding per operation in the order of the torch expressions of objectives.py (explicit
// __dmul_rn / __dadd_rn: no fused multiply-adds), so that the values agree with the torch path to the last bits.
#include "common.cuh"

namespace mbo {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double dclip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// objectives.py:26-50 -- x in the unit square, maximise
__device__ double branin_unit(double u0, double u1) {
  const double PI = 3.141592653589793;
  const double b = 5.1 / (4 * PI * PI), c = 5 / PI, r = 6.0, s = 10.0, t = 1 / (8 * PI);
  const double x1 = dsub(dmul(u0, 15.0), 5.0);
  const double x2 = dmul(u1, 15.0);
  const double inner = dsub(dadd(dsub(x2, dmul(b, dmul(x1, x1))), dmul(c, x1)), r);
  const double f = dadd(dadd(dmul(1.0, dmul(inner, inner)), dmul(s * (1 - t), cos(x1))), s);
  return -dmul(1 / 51.44, dsub(f, 54.44));
}

// objectives.py:165-196
__device__ double hartmann3_unit(double u0, double u1, double u2) {
  const double A[4][3] = {{3.0, 10, 30}, {0.1, 10, 35}, {3.0, 10, 30}, {0.1, 10, 35}};
  const double P[4][3] = {{1e-4 * 3689, 1e-4 * 1170, 1e-4 * 2673}, {1e-4 * 4699, 1e-4 * 4387, 1e-4 * 7470},
                          {1e-4 * 1091, 1e-4 * 8732, 1e-4 * 5547}, {1e-4 * 381, 1e-4 * 5743, 1e-4 * 8828}};
  const double al[4] = {1.0, 1.2, 3.0, 3.2};
  const double u[3] = {u0, u1, u2};
  double f = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double e = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double d = dsub(u[j], P[i][j]);
      const double term = dmul(A[i][j], dmul(d, d));
      e = j == 0 ? term : dadd(e, term);
    }
    const double term = dmul(al[i], exp(-e));
    f = i == 0 ? term : dadd(f, term);
  }
  f = -f;
  return -dmul(1 / 0.95, dsub(f, -0.93));
}

struct EnvStepArgs {
  int family, B, T, D, t_idx, reward_kind;
  const double *x, *trans, *scale, *y_max, *y_min;
  double *X, *Y, *best, *reward;
  int32_t* n;
  float* epi;
  float t_next, T_budget, T_training;
};

__global__ void env_step_kernel(EnvStepArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const double* x = a.x + (size_t)b * a.D;
  const double* tr = a.trans + (size_t)b * a.D;
  double y;
  if (a.family == MBO_ENV_BRANIN) {
    // objectives.py:71-80: translations clipped to the range that keeps the optimum inside the domain
    const double t0 = dclip(tr[0], -0.12, 0.87), t1 = dclip(tr[1], -0.81, 0.18);
    y = dmul(a.scale[b], branin_unit(dsub(x[0], t0), dsub(x[1], t1)));
  } else {
    const double t0 = dclip(tr[0], -0.11, 0.88), t1 = dclip(tr[1], -0.55, 0.44), t2 = dclip(tr[2], -0.85, 0.14);
    y = dmul(a.scale[b], hartmann3_unit(dsub(x[0], t0), dsub(x[1], t1), dsub(x[2], t2)));
  }
  // metabo_gym.py:206-208 add_data
  for (int d = 0; d < a.D; ++d) a.X[((size_t)b * a.T + a.t_idx) * a.D + d] = x[d];
  a.Y[(size_t)b * a.T + a.t_idx] = y;
  a.n[b] = a.t_idx + 1;
  const double best = fmax(a.best[b], y);
  a.best[b] = best;
  // metabo_gym.py:679-704 reward: simple regret min(y_max - Y), optionally transformed
  const double regret = dsub(a.y_max[b], best);
  double rew = regret;
  if (a.reward_kind == MBO_REWARD_NEG_LINEAR) rew = -regret;
  else if (a.reward_kind == MBO_REWARD_NEG_LOG10) rew = -log10(fmax(regret, 1e-20));
  a.reward[b] = rew;
  if (a.epi) {
    // metabo_gym.py:723-730 incumbent (y_min for an empty GP), :642-671 timestep_perc / timestep / budget
    const double inc = isinf(best) ? a.y_min[b] : best;
    float* e = a.epi + (size_t)b * 4;
    e[0] = (float)inc;
    e[1] = a.t_next / a.T_budget;
    e[2] = a.T_training >= 0.f ? fminf(a.t_next, a.T_training) : a.t_next;
    e[3] = a.T_training >= 0.f ? a.T_training : a.T_budget;
  }
}

// get_state (metabo_gym.py:624-677) materialised: state[b, m, :] = the F selected columns of
// {mean, std, x_0 .. x_{D-1}, incumbent, timestep_perc, timestep, budget} as fp32, one thread per (episode, candidate).
struct StatePackArgs {
  mbo_candidates cand;
  int B, F;
  int src[16];
  const float *mean, *std_, *epi;
  float* state;
};

__global__ void state_pack_kernel(StatePackArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int M = a.cand.M;
  if (i >= (size_t)a.B * M) return;
  const int b = (int)(i / M), m = (int)(i - (size_t)b * M);
  double x[MBO_MAX_D];
  load_candidate_compact(a.cand, b, m, x);
  float v[MBO_FEAT_BUDGET + 1];
  v[MBO_FEAT_MEAN] = a.mean[i];
  v[MBO_FEAT_STD] = a.std_[i];
#pragma unroll
  for (int d = 0; d < MBO_MAX_D; ++d) v[MBO_FEAT_X0 + d] = (float)x[d];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[MBO_FEAT_INCUMBENT + e] = a.epi[b * 4 + e];
  float* out = a.state + i * a.F;
  for (int f = 0; f < a.F; ++f) out[f] = v[a.src[f]];
}

}  // namespace mbo

using namespace mbo;

extern "C" int mbo_state_pack(int B, const mbo_candidates* cand, int F, const int32_t* feat_src, const float* mean,
                              const float* std_, const float* epi, float* state, void* stream) {
  MBO_CHECK_ARG(B > 0 && cand && feat_src && mean && std_ && epi && state, "mbo_state_pack: bad argument");
  MBO_CHECK_ARG(F >= 1 && F <= 16, "mbo_state_pack: 1 <= F <= 16 state columns");
  int rc = check_candidates(cand);
  if (rc) return rc;
  StatePackArgs a;
  a.cand = *cand; a.B = B; a.F = F; a.mean = mean; a.std_ = std_; a.epi = epi; a.state = state;
  for (int f = 0; f < 16; ++f) {
    a.src[f] = f < F ? feat_src[f] : 0;
    MBO_CHECK_ARG(a.src[f] >= 0 && a.src[f] <= MBO_FEAT_BUDGET && (a.src[f] < MBO_FEAT_X0 || a.src[f] >= MBO_FEAT_INCUMBENT ||
                                                                     a.src[f] - MBO_FEAT_X0 < cand->D),
                  "mbo_state_pack: bad feature code %d", a.src[f]);
  }
  const size_t tot = (size_t)B * cand->M;
  state_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_env_step(int family, int B, int T, int D, int t_idx, const double* x, const double* trans,
                            const double* scale, const double* y_max, const double* y_min, double* X, double* Y,
                            int32_t* n, double* best, int reward_kind, double* reward, float* epi, float t_next,
                            float T_budget, float T_training, void* stream) {
  MBO_CHECK_ARG(family == MBO_ENV_BRANIN || family == MBO_ENV_HARTMANN3, "mbo_env_step: unknown objective family %d", family);
  MBO_CHECK_ARG(D == (family == MBO_ENV_BRANIN ? 2 : 3), "mbo_env_step: family %d has no %d-dimensional domain", family, D);
  MBO_CHECK_ARG(B > 0 && T > 0 && t_idx >= 0 && t_idx < T, "mbo_env_step: need B > 0 and 0 <= t_idx < T");
  MBO_CHECK_ARG(reward_kind >= MBO_REWARD_NONE && reward_kind <= MBO_REWARD_NEG_LOG10, "mbo_env_step: unknown reward transformation");
  MBO_CHECK_ARG(x && trans && scale && y_max && y_min && X && Y && n && best && reward, "mbo_env_step: null pointer");
  EnvStepArgs a;
  a.family = family; a.B = B; a.T = T; a.D = D; a.t_idx = t_idx; a.reward_kind = reward_kind;
  a.x = x; a.trans = trans; a.scale = scale; a.y_max = y_max; a.y_min = y_min;
  a.X = X; a.Y = Y; a.best = best; a.reward = reward; a.n = n; a.epi = epi;
  a.t_next = t_next; a.T_budget = T_budget; a.T_training = T_training;
  env_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}
