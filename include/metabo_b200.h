/*
 * metabo_b200.h -- C ABI of libmetabo_b200.so (sm_100a only).
 *
 * Drop-in boundary for MetaBO's data-parallel hot path: per BO step and per episode,
 * GP posterior (mean, std) at M candidates -> features -> NeuralAF MLP -> argmax /
 * categorical over the M logits.  The reference (boschresearch/MetaBO) has no FFI of its
 * own (pure Python); each entry point below names the reference function it replaces.
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every array argument is a DEVICE pointer into caller-owned memory (torch tensors);
 *     nothing is allocated inside, workspaces are sized by the *_bytes queries;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no host sync);
 *   - return value: 0 = ok, negative = error (mbo_last_error() gives a thread-local string);
 *     no C++ exception crosses the boundary;
 *   - episode b of B is an independent unit (one BO episode); candidates are per episode.
 */
#ifndef METABO_B200_H
#define METABO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBO_OK 0
#define MBO_ERR_ARG (-1)
#define MBO_ERR_CUDA (-2)
#define MBO_ERR_UNSUPPORTED (-3)

/* GP kernels of metabo_gym.py:554-564 (GPy.kern.RBF / Matern32 / Matern52, ARD) */
#define MBO_KERNEL_RBF 0
#define MBO_KERNEL_MATERN32 1
#define MBO_KERNEL_MATERN52 2

/* arithmetic of the posterior kernel */
#define MBO_GP_FP64 0 /* validation mode: everything fp64 */
#define MBO_GP_FP32 1 /* kernel rows, solve and dots in fp32 (fit is always fp64) */

/* feature sources, one per policy-input column (metabo_gym.py:630-671 column order) */
#define MBO_FEAT_MEAN 0
#define MBO_FEAT_STD 1
#define MBO_FEAT_X0 2 /* MBO_FEAT_X0 + d, d < 8 */
#define MBO_FEAT_INCUMBENT 10
#define MBO_FEAT_TPERC 11
#define MBO_FEAT_TIMESTEP 12
#define MBO_FEAT_BUDGET 13
#define MBO_MAX_FEATURES 16
#define MBO_MAX_D 8
#define MBO_MAX_LAYERS 8

/* activation of the hidden layers (policies.py:69-74) */
#define MBO_ACT_RELU 0
#define MBO_ACT_TANH 1

/* closed-form acquisition functions (mbo_af_closed_form) */
#define MBO_AF_EI 0
#define MBO_AF_UCB 1
#define MBO_AF_PI 2

/* MLP arithmetic */
#define MBO_MLP_FP32 0   /* CUDA-core fp32 FMA (validation mode, any architecture) */
#define MBO_MLP_TF32 1   /* tcgen05 kind::tf32, fp32 accumulate in TMEM */
#define MBO_MLP_BF16X3 2 /* tcgen05 kind::f16 bf16 hi/lo split, 3 MMAs per product */
#define MBO_MLP_FP16 3   /* tcgen05 kind::f16 fp16: tf32's 11-bit significand at twice the rate, biases inside the
                            MMAs (H <= 206); activations beyond fp16's range (65504 after the per-layer power-of-two
                            renormalisation) are clamped and set bit 1 of the workspace's error word */

/* Candidate set of one call.  Candidate m of episode b is
 *   m <  n_head                : head[b, m, :]                    (per-episode points)
 *   m >= n_head, local == 0    : tail[m - n_head, :]              (shared by all episodes)
 *   m >= n_head, local != 0    : tail[s, :] * (hi - lo) + lo,  s = (m-n_head) % n_tail,
 *                                cube = (m-n_head) / n_tail, lo/hi = cubes[b, cube, {0,1}, :]
 * which covers metabo_gym.py's three sets: multistart grid (tail only), local Sobol grids
 * (:711-716, tail = Sobol table, cubes from util.py:45-51) and xi_t (:619, head = af maxima,
 * tail = multistart grid), plus per-episode streaming candidates (head only).
 * Coordinates are fp64 (as in the reference) unless head_is_f32 / tail_is_f32 is set. */
typedef struct {
  int M;              /* candidates per episode */
  int D;              /* input dimension, <= MBO_MAX_D */
  int n_head;         /* per-episode leading points */
  int n_tail;         /* rows of the shared table */
  int local;          /* 1: tail rows are mapped into per-episode cubes */
  int n_cubes;        /* cubes per episode when local (M = n_head + n_cubes * n_tail) */
  int head_is_f32;
  int tail_is_f32;
  const void* head;   /* [B, n_head, D] */
  const void* tail;   /* [n_tail, D] */
  const double* cubes; /* [B, n_cubes, 2, D] (lo row, hi row), local only */
} mbo_candidates;

const char* mbo_last_error(void);
int mbo_version(void);

/* ---------------------------------------------------------------------------------------
 * GP fit  (replaces MetaBO.update_gp -> GPy set_XY, metabo_gym.py:610-613, and the kernel /
 * prior-mean setup of reset_gp :549-593).  Full fp64 refit of B episodes:
 *   Ky = K(X,X) + (noise + 1e-8) I,  L = chol(Ky) (GPy jitchol retry ladder on failure),
 *   Linv = L^-1,  beta = Linv (Y - m),  m = mean(Y) if use_prior_mean else 0.
 *   n[b] may be 0 (empty GP, metabo_gym.py:736-738) and differ per episode, n[b] <= nmax.
 * gp_state: [B, mbo_gp_state_doubles(nmax, D)] fp64, opaque to the caller.
 * status:   [B] int32, 0 ok, k>0 = number of jitter retries used, -1 = not positive definite.
 * ------------------------------------------------------------------------------------- */
size_t mbo_gp_state_doubles(int nmax, int D);
int mbo_gp_fit(int B, int nmax, int D, int kernel_id, const int32_t* n,
               const double* X /*[B,nmax,D]*/, const double* Y /*[B,nmax]*/,
               const double* lengthscale /*[B,D]*/, const double* variance /*[B]*/,
               const double* noise /*[B]*/, int use_prior_mean,
               double* gp_state, int32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------
 * GP posterior  (replaces MetaBO.eval_gp, metabo_gym.py:732-745 -> GPy predict_noiseless):
 *   mean = k*^T Ky^-1 (Y-m) + m,  var = max(k** - |Linv k*|^2, 1e-15),  std = sqrt(var);
 *   empty GP: mean = 0, std = sqrt(kernel_variance).
 * Outputs mean/std [B, M]: fp64 when out_is_f32 == 0, else fp32 (the cast get_state does).
 * n_active_max: upper bound of n[b] over the batch (<= nmax; sizes shared memory, pass nmax
 * if unknown).  The kernel traps nothing: an episode with n[b] > n_active_max is an error of
 * the caller and is reported through the status word of the launch check only in debug builds.
 * ------------------------------------------------------------------------------------- */
int mbo_gp_posterior(int B, int nmax, int n_active_max, int kernel_id, int precision,
                     const double* gp_state, const mbo_candidates* cand, void* mean, void* std,
                     int out_is_f32, void* stream);

/* ---------------------------------------------------------------------------------------
 * NeuralAF weights  (NeuralAF.policy_net / value_net, policies.py:87-97; mlp.py:25-50).
 * The handle owns device copies of the weights in the layouts the kernels want (fp32 for the
 * CUDA-core path, TF32 / bf16 hi+lo UMMA-canonical tiles for the tcgen05 path).
 *   widths[0..n_layers]: layer sizes d_in, h_1, ..., d_out(=1);  W_l: [widths[l+1], widths[l]]
 *   row-major (torch nn.Linear.weight), b_l: [widths[l+1]];  host or device pointers.
 * ------------------------------------------------------------------------------------- */
typedef struct mbo_policy mbo_policy;
int mbo_policy_create(mbo_policy** out);
void mbo_policy_destroy(mbo_policy* p);
int mbo_policy_set_weights(mbo_policy* p, int n_layers, const int32_t* widths, int activation,
                           const float* const* W, const float* const* b, int on_device, void* stream);

/* ---------------------------------------------------------------------------------------
 * NeuralAF evaluation  (replaces get_state feature packing metabo_gym.py:624-677 +
 * NeuralAF.forward policies.py:104-125 / MLP.forward mlp.py:52-61 for the policy net).
 *   feat_src[F]   : MBO_FEAT_* code of every policy-input column (masked columns omitted)
 *   epi[B,4] fp32 : per-episode {incumbent, t/T, timestep, budget} (get_state :642-671)
 *   mean, std     : [B, M] fp32 from mbo_gp_posterior
 *   logits [B, M] fp32.
 * mode MBO_MLP_FP32 supports any architecture; the tensor-core modes need a
 * F->H->H->H->H->1 ReLU net with F <= 15 and H <= 208 (the shipped 4x200 policies).
 * ------------------------------------------------------------------------------------- */
int mbo_af_logits(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src,
                  const mbo_candidates* cand, const float* mean, const float* std,
                  const float* epi, float* logits, void* workspace, size_t workspace_bytes,
                  void* stream);
size_t mbo_af_workspace_bytes(int mode, int B, int M);

/* NeuralAF evaluation with the top-k selection of get_af_maxima (metabo_gym.py:709,719) fused into the
 * tensor-core kernel's epilogue: every epilogue warp emits its k best (logit, index) pairs, a merge kernel
 * reduces them per episode -> topk_idx / topk_val [B,k] (value descending, ties by lowest index, k <= 8).
 * `logits` may be null: then the M logits of an episode never reach HBM. */
int mbo_af_topk(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src,
                const mbo_candidates* cand, const float* mean, const float* std, const float* epi, int k,
                int32_t* topk_idx, float* topk_val, float* logits, void* workspace, size_t workspace_bytes,
                void* stream);

/* Same network on an already materialised state tensor (NeuralAF.forward on arbitrary input,
 * policies.py:104-115; the PPO update re-evaluates stored states, ppo.py:148):
 *   states [B, M, F_state] fp32, cols[F] = state column feeding policy input i (the boolean
 *   feature mask of policies.py:109-113 expressed as indices) -> logits [B, M] fp32. */
int mbo_af_logits_dense(const mbo_policy* p, int mode, int B, int M, int F_state, int F,
                        const int32_t* cols, const float* states, float* logits, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Fused form of mbo_gp_posterior + mbo_af_logits for n_active_max <= 32 observations (the T = 30
 * configs): dedicated fp64 warps of the tensor-core kernel compute the GP posterior of the NEXT
 * candidate tile while the tensor pipe runs the MLP of the current one; mean/std never touch HBM
 * unless mean_out/std_out ([B, M] fp32, may both be null) are requested for the PPO state buffer.
 * Returns MBO_ERR_UNSUPPORTED for larger n or unsupported architectures (use the two-call path). */
int mbo_af_eval_fused(const mbo_policy* p, int mode, int B, int nmax, int n_active_max, int kernel_id,
                      const double* gp_state, int F, const int32_t* feat_src, const mbo_candidates* cand,
                      const float* epi, float* logits, float* mean_out, float* std_out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* value network on (t, T) of row 0 (policies.py:117-121): tT [B,2] fp32 -> values [B] */
int mbo_value_net(const mbo_policy* value_net, int B, const float* tT, float* values, void* stream);

/* ---------------------------------------------------------------------------------------
 * Reductions over the M logits of each episode
 *   argmax         NeuralAF.act deterministic branch (policies.py:141-142), first-max ties
 *   topk           np.argsort(-af)[:k] of get_af_maxima (metabo_gym.py:709,719), k <= 8
 *   sample         Categorical(logits).sample() (policies.py:144-146) by Gumbel-max with a
 *                  counter-based generator keyed (seed, episode, step, candidate)
 *   logp/entropy   Categorical.log_prob / entropy (policies.py:157-159)
 * ------------------------------------------------------------------------------------- */
/* Closed-form acquisition functions of the reference's baselines on the GP posterior (policies.py:181-307):
 * kind MBO_AF_EI, MBO_AF_PI (p0 = xi), MBO_AF_UCB (p0 = kappa; p1 > 0 selects the GP-UCB schedule with delta = p1 and
 * input dimension D).  mean, std [B,M] fp32 from mbo_gp_posterior, epi [B,4] as in mbo_af_logits (incumbent, -,
 * timestep, -).  af [B,M] fp32: feed it to mbo_logits_argmax / mbo_logits_topk like NeuralAF logits. */
int mbo_af_closed_form(int kind, int B, int M, const float* mean, const float* std, const float* epi, double p0,
                       double p1, int D, float* af, void* stream);

int mbo_logits_argmax(int B, int M, const float* logits, int32_t* action, float* max_logit, void* stream);
int mbo_logits_topk(int B, int M, int k, const float* logits, int32_t* idx /*[B,k]*/,
                    float* val /*[B,k]*/, void* stream);
int mbo_logits_sample(int B, int M, const float* logits, uint64_t seed, uint64_t step,
                      const uint64_t* step_dev /* optional device counter added to `step` (CUDA-graph replays) */,
                      uint64_t episode_offset /* global index of row 0 */, int32_t* action, void* stream);
int mbo_logits_logp_entropy(int B, int M, const float* logits, const int32_t* actions,
                            float* logp, float* entropy, void* stream);

/* gather candidate coordinates by index: x_out[b, j, :] = candidate idx[b, j] (fp64), the
 * convert_idx_to_x / af_maxima gathers of metabo_gym.py:709,719,838-842 */
int mbo_candidates_gather(int B, const mbo_candidates* cand, int k, const int32_t* idx,
                          double* x_out /*[B,k,D]*/, void* stream);

/* cubes around start points (util.py:45-51 get_cube_around, domain = unit cube):
 * cubes[b,j,0,:] = max(x - diam/2, 0), cubes[b,j,1,:] = min(x + diam/2, 1) */
int mbo_cubes_around(int B, int k, int D, const double* x /*[B,k,D]*/, double diam,
                     double* cubes /*[B,k,2,D]*/, void* stream);

/* ---- device-resident environment step ----------------------------------------------
 * One BO step of B lock-step episodes on the analytic objective families: replaces, for every lane b,
 *   y = s_b f(x_b - clip(t_b))            metabo/environment/objectives.py:26-80 (Branin), :165-219 (Hartmann-3)
 *   X[b, t_idx] = x_b, Y[b, t_idx] = y    metabo_gym.py:195-217 (step -> add_data), n[b] = t_idx + 1
 *   best[b] = max(best[b], y); reward[b]  metabo_gym.py:679-704 (simple regret y_max - best, transformed)
 *   epi[b] = {incumbent, t_next / T_budget, timestep, budget} of the NEXT state (metabo_gym.py:642-671, 723-730;
 *            T_training >= 0: timestep = min(t_next, T_training), budget = T_training); epi may be NULL
 * fp64, one rounding per operation in the order of the reference's expressions.  trans [B, D], scale / y_max /
 * y_min / best / reward [B] fp64, x [B, D] fp64, X [B, T, D], Y [B, T], n [B] int32, epi [B, 4] fp32. */
#define MBO_ENV_BRANIN 0
#define MBO_ENV_HARTMANN3 1
#define MBO_REWARD_NONE 0
#define MBO_REWARD_NEG_LINEAR 1
#define MBO_REWARD_NEG_LOG10 2
int mbo_env_step(int family, int B, int T, int D, int t_idx, const double* x, const double* trans,
                 const double* scale, const double* y_max, const double* y_min, double* X, double* Y,
                 int32_t* n, double* best, int reward_kind, double* reward, float* epi, float t_next,
                 float T_budget, float T_training, void* stream);

/* get_state (metabo_gym.py:624-677) materialised for the PPO buffer: state[b, m, f] = column feat_src[f] (MBO_FEAT_*,
 * any of the ten columns, HOST array of F <= 16 codes) of {mean, std, x, incumbent, timestep_perc, timestep, budget} as
 * fp32; mean / std [B, M] fp32, epi [B, 4] fp32, state [B, M, F] fp32. */
int mbo_state_pack(int B, const mbo_candidates* cand, int F, const int32_t* feat_src, const float* mean,
                   const float* std_, const float* epi, float* state /*[B,M,F]*/, void* stream);

/* ---- PPO policy update ------------------------------------------------------------
 * The policy-net half of one PPO optimiser step on one (shard of a) minibatch: replaces the autograd
 * forward + backward of ppo.py:129-178 (clipped surrogate + entropy bonus), policies.py:150-161
 * (Categorical(logits).log_prob / entropy) and mlp.py:25-61 for the F -> H -> H -> H -> H -> 1 ReLU
 * policy net (H <= 207, F <= 8).  Forward: one launch of the tcgen05 policy kernel in its fp32-class
 * bf16x3 mode (activations + ReLU words stored for the backward); backward: TF32 tensor-core GEMMs with fp32
 * accumulation; deterministic (fixed-order split-K reduction, no atomics).  The call enqueues on `stream` and
 * on one library-internal stream forked from it and joined before the call returns (legal inside a stream
 * capture: the forks become parallel graph branches); the FIRST call of a process creates that stream and
 * must not be made under capture.
 *   states      [E, M, F_state] fp32 (the PPO buffer's states); feat_col[F] = state column of policy
 *               feature f (policies.py:109-114 feature mask), HOST array
 *   W, b        HOST arrays of 5 device pointers, torch layout W_l [out][in], b_l [out]
 *   actions     [E] int32; advs, logp_old [E] fp32 (advantages already normalised, ppo.py:141-144)
 *   inv_n_global  1 / (global minibatch size): every loss term is divided by it (ppo.py:148-168 mean)
 *   dW, db      HOST arrays of 5 device pointers, gradients of
 *               loss_ppo + ent_coeff * loss_ent  w.r.t. W_l, b_l (overwritten, not accumulated);
 *               dW == NULL: forward only (advs / logp_old / db may be NULL): terms[:, 0:2] = log-prob and
 *               entropy under these weights -- the frozen old policy's logprobs_old of ppo.py:146-147, in the
 *               same arithmetic as the update itself (ratio == 1 exactly while theta == theta_old)
 *   terms       [E, 4] fp32: logp(action), entropy, -min(ratio A, clip(ratio) A) / n, ratio
 *   workspace   mbo_ppo_workspace_bytes(E, M) bytes, 256-byte aligned, zero-initialised ONCE by the
 *               caller; its first int32 is a sticky error word (bit 0: a pipeline wait timed out)
 * mbo_ppo_debug_offsets: byte offsets inside the workspace of h1..h4, dZ1..dZ4 ([E*M, 208] fp32 each),
 * logits, dlogits ([E*M] fp32), terms, error word -- for the per-stage parity tests. */
size_t mbo_ppo_workspace_bytes(int E, int M);
int mbo_ppo_debug_offsets(int E, int M, size_t* off /*[12]*/);
int mbo_ppo_policy_grad(int E, int M, int F_state, int F, const int32_t* feat_col, int H,
                        const float* states, const float* const* W, const float* const* b,
                        const int32_t* actions, const float* advs, const float* logp_old, float epsilon,
                        float ent_coeff, float inv_n_global, float* const* dW, float* const* db,
                        float* terms, void* workspace, size_t workspace_bytes, void* stream);

/* Value-net half of the optimiser step: forward + backward of value_coeff * mean((vpreds - tdlamrets)^2)
 * (ppo.py:150-156) for the 2 -> Hv -> Hv -> Hv -> Hv -> 1 ReLU value net of policies.py:87-92, whose inputs are
 * states[e, 0, t_col] and states[e, 0, T_col] (policies.py:116-121); fp32 CUDA cores.
 *   states       device pointer to the minibatch states, state_stride = floats between consecutive transitions (M * F)
 *   coeff        value_coeff / (global minibatch size)
 *   dW, db       gradients (overwritten); values [E] = vpreds; vterms [E] = (vpreds - tdlamrets)^2
 *   workspace    mbo_ppo_value_workspace_bytes(E, Hv) bytes */
size_t mbo_ppo_value_workspace_bytes(int E, int Hv);
int mbo_ppo_value_grad(int E, size_t state_stride, int t_col, int T_col, int Hv, const float* states,
                       const float* const* W, const float* const* b, const float* tdlamrets, float coeff,
                       float* const* dW, float* const* db, float* values, float* vterms, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Minibatch gather + advantage normalisation of one optimiser step (ppo.py:129-144; batchrecorder.py:318-333 yields the
 * shuffled minibatches): transitions sel[0..E) (device int64 indices into the flattened batch) are gathered from the
 * batch arrays; advantages become (adv - mean) / std with the biased std of the minibatch (left alone if std is 0 / NaN).
 * row_floats = M * F floats per state.  logp_old_all may be NULL (zeros are written).  adv_stats (device, 3 doubles
 * {sum, sum of squares, count}, may be NULL): moments of the GLOBAL minibatch when it is sharded over ranks (NULL: the
 * moments of these E transitions). */
int mbo_ppo_prepare_minibatch(int E, size_t row_floats, const int64_t* sel, const float* states_all,
                              const int32_t* actions_all, const float* advs_all, const float* tdlamrets_all,
                              const float* logp_old_all, int normalize_advs, const double* adv_stats, float* states,
                              int32_t* actions, float* advs, float* tdlamrets, float* logp_old, void* stream);
/* out[j] = {sum, sum of squares, count} (doubles) of advs_all[sel[bounds[j] .. bounds[j+1])] for the n_steps local
 * minibatch shards of one epoch (bounds: device int32 [n_steps + 1]); the ranks sum the table with mbo_comm_allreduce. */
int mbo_ppo_adv_stats(int n_steps, const int32_t* bounds, const int64_t* sel, const float* advs_all, double* out,
                      void* stream);
/* out4 = {loss_ppo, loss_value, loss_ent, loss} (ppo.py:158-168) from mbo_ppo_policy_grad's terms [E,4] and
 * mbo_ppo_value_grad's vterms [E] (NULL: no value net). */
int mbo_ppo_loss_sums(int E, const float* terms, const float* vterms, float inv_n_global, float value_coeff,
                      float ent_coeff, float* out4, void* stream);

/* Adam step (ppo.py:93, torch.optim.Adam defaults: no weight decay, no amsgrad) over n <= 32 parameter tensors in
 * one launch: m = lerp(m, g, 1 - beta1); v = beta2 v + (1 - beta2) g^2; p -= lr / (1 - beta1^t) * m /
 * (sqrt(v) / sqrt(1 - beta2^t) + eps).  p, g, m, v: HOST arrays of n device pointers (fp32), sizes: HOST array of
 * element counts.  state: HOST array of n device pointers, each to 3 x 8 bytes {int64 step t, double, double},
 * zero-initialised by the caller (one step count per tensor, as torch keeps it); the call advances t on the device,
 * so a captured CUDA graph of the step can be replayed. */
int mbo_adam_step(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                  const int64_t* sizes, int64_t* const* state, float lr, float beta1, float beta2, float eps,
                  void* stream);

/* PPO optimiser step for every network shape the tensor-core path (mbo_ppo_policy_grad / mbo_ppo_value_grad) does not
 * cover: any MLP of mlp.py:25-61 with n_layers <= 8 linear layers of width <= 512, ReLU or tanh (policies.py:69-74),
 * d_out = 1 -- e.g. train_metabo_gps.py:62-94's 4 -> 20 -> 20 -> 1 policy and 2 -> 20 -> 20 -> 1 value net -- in fp32 on
 * CUDA cores (csrc/ppo_update_simt.cu).  widths: HOST array of n_layers + 1 ints {inputs, hidden..., 1}; W / b / dW /
 * db: HOST arrays of n_layers device pointers in torch layouts.  Same semantics, terms and scaling as the calls
 * above (dW == NULL: forward only).  Workspace: mbo_ppo_simt_workspace_bytes(rows = E * M or E, n_layers, widths, E). */
size_t mbo_ppo_simt_workspace_bytes(long long rows, int n_layers, const int32_t* widths, int E);
int mbo_ppo_policy_grad_simt(int E, int M, int F_state, int n_layers, const int32_t* widths, const int32_t* feat_col,
                             int act, const float* states, const float* const* W, const float* const* b,
                             const int32_t* actions, const float* advs, const float* logp_old, float epsilon,
                             float ent_coeff, float inv_n_global, float* const* dW, float* const* db, float* terms,
                             void* workspace, size_t workspace_bytes, void* stream);
int mbo_ppo_value_grad_simt(int E, size_t state_stride, int t_col, int T_col, int n_layers, const int32_t* widths,
                            int act, const float* states, const float* const* W, const float* const* b,
                            const float* tdlamrets, float coeff, float* const* dW, float* const* db, float* values,
                            float* vterms, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Peer-memory exchange for the data-parallel PPO update (one process per GPU; the reference's learner is a single
 * process -- ppo.py:170-172 `loss.backward(); optimizer.step()` -- so the flat gradient all-reduce is the only collective
 * of the path).  An `mbo_comm` owns one cudaMalloc'ed region per rank, exported as a CUDA IPC handle; the
 * ranks exchange the MBO_COMM_HANDLE_BYTES-byte handles on the host (e.g. torch.distributed.all_gather_object) and map
 * each other's regions over NVLink / NVSwitch.  All calls below except create / export / connect / destroy / status
 * only enqueue kernels (capturable in CUDA graphs); every rank must enqueue the same sequence of calls.  Sums are formed
 * in rank order 0..world-1 on every rank: results are bit-identical across ranks.  Waits on peers are bounded (10 s): a
 * time-out sets a sticky error bit (later waits return at once) that mbo_comm_status reports; nothing hangs. */
typedef struct mbo_comm mbo_comm;
#define MBO_COMM_HANDLE_BYTES 64
#define MBO_COMM_MAX_RANKS 16
#define MBO_COMM_F32 0
#define MBO_COMM_F64 1
/* slot_bytes: largest payload of one call (e.g. 4 x number of parameters).  Allocates 2 slots + 2.3 KB of flags. */
int mbo_comm_create(int rank, int world, size_t slot_bytes, mbo_comm** out);
int mbo_comm_export(mbo_comm* c, void* handle_out /* HOST, MBO_COMM_HANDLE_BYTES */);
int mbo_comm_connect(mbo_comm* c, const void* handles /* HOST, world x MBO_COMM_HANDLE_BYTES, rank order */);
void mbo_comm_destroy(mbo_comm* c);
/* dst[i] = sum over ranks of src[i] (count elements of fp32 / fp64, device pointers; src == dst allowed). */
int mbo_comm_allreduce(mbo_comm* c, int dtype, const void* src, void* dst, size_t count, void* stream);
/* The optimiser step of ppo.py:170-172 under data parallelism as ONE kernel: all-reduce (sum) of the flat fp32
 * gradient `grad` (the concatenation of the n tensors' gradients in the order of p / m / v / sizes / state, as written
 * by mbo_ppo_policy_grad / mbo_ppo_value_grad into views of one buffer) pulled from every rank's slot over NVLink,
 * consumed on the fly by the Adam update of mbo_adam_step (same arguments, same arithmetic).  grad_out (optional,
 * device, count floats): the reduced gradient.  Weights stay replicated: every rank applies the identical update. */
int mbo_comm_allreduce_adam(mbo_comm* c, const float* grad, float* grad_out, int n, float* const* p, float* const* m,
                            float* const* v, const int64_t* sizes, int64_t* const* state, float lr, float beta1,
                            float beta2, float eps, void* stream);
/* Synchronises `stream`; *epoch_out = completed calls, *err_out = sticky error bits (bit 0: a peer wait timed out),
 * cleared by the call. */
int mbo_comm_status(mbo_comm* c, unsigned long long* epoch_out, unsigned int* err_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* METABO_B200_H */
