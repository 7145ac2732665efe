// ppo_update_simt.cu -- PPO optimiser step (ppo.py:129-178) for the networks the tensor-core path does not cover:
// any MLP of mlp.py:25-61 (1..7 hidden layers of any width <= 512, ReLU or tanh, d_out = 1), in fp32 on CUDA cores --
// the 4 -> 20 -> 20 -> 1 policy and the 2 -> 20 -> 20 -> 1 value net of train_metabo_gps.py:62-94 (C4), tanh policies
// (policies.py:69-74).  With it every architecture NeuralAF can build is updated by the library's own kernels.
//
// Layout: activations and gradients are FEATURE-MAJOR, a_l[i][r] (r = row = (transition, candidate) or transition), so
// that with one thread per row every load / store is coalesced and every weight read is a warp-wide broadcast:
//   forward  z_l[o][r] = b_l[o] + sum_i W_l[o][i] a_{l-1}[i][r],  a_l = act(z_l);  output = z_{L-1}[0][r]
//   backward dA_{l-1}[i][r] = sum_o dZ_l[o][r] W_l[o][i],  dZ_{l-1} = dA_{l-1} * act'(a_{l-1})
//   wgrad    dW_l[o][i] = sum_r dZ_l[o][r] a_{l-1}[i][r] -- a dot product of two contiguous rows: one CTA per (row chunk,
//            o), partial sums per chunk, summed over the chunks in a fixed order (bitwise reproducible, no atomics).
// Layer 0 reads its inputs straight from the state tensor (row stride + column list: the policy's feature mask
// policies.py:109-114, or the value net's (t, T) columns of candidate 0, :116-121).
// FLOPs are negligible for the nets this path serves (1 000 per row for C4's policy); it is bound by launch count
// (~4 L + 3 launches) and runs inside the captured epoch graph like the tensor-core path.
#include "common.cuh"
#include "ppo_loss.cuh"

namespace mbo {
namespace upds {

constexpr int MAXL = MBO_MAX_LAYERS;       // linear layers
constexpr int MAXW = 512;
constexpr int THREADS = 256;
constexpr int CHUNK_ROWS = 2048;           // rows per wgrad partial

struct SNet {
  const float* W[MAXL];
  const float* b[MAXL];
  float* dW[MAXL];
  float* db[MAXL];
  int width[MAXL + 1];       // width[0] = inputs of layer 0, width[l + 1] = outputs of layer l (last: 1)
  int col[MBO_MAX_FEATURES];
  int L, act;
  long long in_stride;       // floats between the inputs of consecutive rows in `states`
};
struct SWs {
  float* a[MAXL];            // a[l]: output of layer l (post-activation), [width[l + 1]][Rp]; a[L - 1] = the output [Rp]
  float* dz[2];              // ping-pong gradient buffers [Hmax][Rp]
  float* part;               // [n_chunk][P] wgrad partials
  float* terms;              // [E][4] (policy) / unused
  long long Rp;
  int n_chunk, P;
  size_t bytes;
};

static size_t carve(long long R, int L, const int* width, int E, unsigned char* base, SWs* ws) {
  const long long Rp = (R + 31) & ~31LL;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return base ? base + o : nullptr; };
  int hmax = 1, P = 0;
  for (int l = 0; l < L; ++l) { hmax = width[l + 1] > hmax ? width[l + 1] : hmax; P += (width[l] + 1) * width[l + 1]; }
  SWs w{};
  for (int l = 0; l < L; ++l) w.a[l] = (float*)take((size_t)width[l + 1] * Rp * 4);
  w.dz[0] = (float*)take((size_t)hmax * Rp * 4);
  w.dz[1] = (float*)take((size_t)hmax * Rp * 4);
  w.n_chunk = (int)((R + CHUNK_ROWS - 1) / CHUNK_ROWS);
  w.P = P;
  w.part = (float*)take((size_t)w.n_chunk * P * 4);
  w.terms = (float*)take((size_t)(E > 0 ? E : 1) * 16);
  w.Rp = Rp;
  w.bytes = off;
  if (ws) *ws = w;
  return off;
}

__device__ __forceinline__ float act_f(float z, int act) { return act == MBO_ACT_RELU ? fmaxf(z, 0.f) : tanhf(z); }
__device__ __forceinline__ float dact_f(float a, int act) { return act == MBO_ACT_RELU ? (a > 0.f ? 1.f : 0.f) : 1.f - a * a; }

// one thread per row, four output columns at a time (every input value feeds four FMAs)
__global__ void __launch_bounds__(THREADS) fwd_kernel(SNet net, int l, const float* __restrict__ states, long long R, long long Rp,
                                                      const float* __restrict__ a_in, float* __restrict__ a_out) {
  const long long r = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (r >= R) return;
  const int Kin = net.width[l], Hout = net.width[l + 1];
  const bool last = l == net.L - 1;
  const float* __restrict__ W = net.W[l];
  const float* __restrict__ b = net.b[l];
  float x0[MBO_MAX_FEATURES];
  if (l == 0) {
#pragma unroll
    for (int i = 0; i < MBO_MAX_FEATURES; ++i) x0[i] = i < Kin ? states[r * net.in_stride + net.col[i]] : 0.f;
  }
  for (int o0 = 0; o0 < Hout; o0 += 4) {
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = o0 + j < Hout ? __ldg(b + o0 + j) : 0.f;
    if (l == 0) {
#pragma unroll
      for (int i = 0; i < MBO_MAX_FEATURES; ++i) {
        if (i < Kin) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (o0 + j < Hout) acc[j] = fmaf(x0[i], __ldg(W + (size_t)(o0 + j) * Kin + i), acc[j]);
        }
      }
    } else {
      for (int i = 0; i < Kin; ++i) {
        const float x = a_in[(size_t)i * Rp + r];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o0 + j < Hout) acc[j] = fmaf(x, __ldg(W + (size_t)(o0 + j) * Kin + i), acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (o0 + j < Hout) a_out[(size_t)(o0 + j) * Rp + r] = last ? acc[j] : act_f(acc[j], net.act);
  }
}

// dZ_{l-1}[i][r] = (sum_o dZ_l[o][r] W_l[o][i]) * act'(a_{l-1}[i][r]), l >= 1
__global__ void __launch_bounds__(THREADS) bwd_kernel(SNet net, int l, long long R, long long Rp, const float* __restrict__ dz_in,
                                                      const float* __restrict__ a_prev, float* __restrict__ dz_out) {
  const long long r = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (r >= R) return;
  const int Kin = net.width[l], Hout = net.width[l + 1];
  const float* __restrict__ W = net.W[l];
  for (int i0 = 0; i0 < Kin; i0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int o = 0; o < Hout; ++o) {
      const float d = dz_in[(size_t)o * Rp + r];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j < Kin) acc[j] = fmaf(d, __ldg(W + (size_t)o * Kin + i0 + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < Kin) dz_out[(size_t)(i0 + j) * Rp + r] = acc[j] * dact_f(a_prev[(size_t)(i0 + j) * Rp + r], net.act);
  }
}

// grid (n_chunk, Hout): partial dW_l[o][0..Kin) and db_l[o] over the rows of one chunk
__global__ void __launch_bounds__(THREADS) wgrad_kernel(SNet net, int l, const float* __restrict__ states, long long R, long long Rp,
                                                        const float* __restrict__ dz, const float* __restrict__ a_prev,
                                                        float* __restrict__ part, int P, int p_off) {
  __shared__ double s_red[8];
  const int o = blockIdx.y, Kin = net.width[l];
  const long long r0 = (long long)blockIdx.x * CHUNK_ROWS, r1 = min(R, r0 + CHUNK_ROWS);
  constexpr int PER = CHUNK_ROWS / THREADS;
  float d[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const long long r = r0 + threadIdx.x + (long long)k * THREADS;
    d[k] = r < r1 ? dz[(size_t)o * Rp + r] : 0.f;
  }
  float* out = part + (size_t)blockIdx.x * P + p_off + (size_t)o * (Kin + 1);
  for (int i = 0; i <= Kin; ++i) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const long long r = r0 + threadIdx.x + (long long)k * THREADS;
      if (r < r1) {
        const float x = i == Kin ? 1.f : (l == 0 ? states[r * net.in_stride + net.col[i]] : a_prev[(size_t)i * Rp + r]);
        s = fmaf(d[k], x, s);
      }
    }
    const double t = mbo::upd::block_sum((double)s, s_red);
    if (threadIdx.x == 0) out[i] = (float)t;
  }
}

// sums the chunk partials in a fixed order and scatters into the torch layouts (dW [out][in], db [out])
__global__ void __launch_bounds__(THREADS) reduce_kernel(SNet net, const float* __restrict__ part, int n_chunk, int P) {
  const int p = blockIdx.x * THREADS + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int c = 0; c < n_chunk; ++c) s += part[(size_t)c * P + p];
  int q = p;
  for (int l = 0; l < net.L; ++l) {
    const int Kin = net.width[l], n = (Kin + 1) * net.width[l + 1];
    if (q < n) {
      const int o = q / (Kin + 1), i = q - o * (Kin + 1);
      if (i == Kin) net.db[l][o] = s; else net.dW[l][(size_t)o * Kin + i] = s;
      return;
    }
    q -= n;
  }
}

// value loss (ppo.py:150-156): vterms = (v - ret)^2, dz = 2 coeff (v - ret)
__global__ void vloss_kernel(int E, const float* __restrict__ v, const float* __restrict__ ret, float coeff, float* __restrict__ dz,
                             float* __restrict__ values, float* __restrict__ vterms) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float d = v[e] - ret[e];
  if (dz) dz[e] = 2.f * coeff * d;
  if (values) values[e] = v[e];
  if (vterms) vterms[e] = d * d;
}

static int fill_net(SNet& net, int L, const int32_t* widths, int act, const float* const* W, const float* const* b, float* const* dW,
                    float* const* db, const char* who) {
  MBO_CHECK_ARG(L >= 1 && L <= MAXL && widths, "%s: 1..%d linear layers", who, MAXL);
  MBO_CHECK_ARG(act == MBO_ACT_RELU || act == MBO_ACT_TANH, "%s: unknown activation", who);
  MBO_CHECK_ARG(widths[0] >= 1 && widths[0] <= MBO_MAX_FEATURES, "%s: 1..%d input features", who, MBO_MAX_FEATURES);
  MBO_CHECK_ARG(widths[L] == 1, "%s: d_out must be 1", who);
  for (int l = 0; l <= L; ++l) {
    MBO_CHECK_ARG(widths[l] >= 1 && widths[l] <= MAXW, "%s: layer width %d outside 1..%d", who, widths[l], MAXW);
    net.width[l] = widths[l];
  }
  for (int l = 0; l < L; ++l) {
    MBO_CHECK_ARG(W && b && W[l] && b[l], "%s: null layer pointer", who);
    MBO_CHECK_ARG(!dW || (db && dW[l] && db[l]), "%s: null gradient pointer", who);
    net.W[l] = W[l]; net.b[l] = b[l]; net.dW[l] = dW ? dW[l] : nullptr; net.db[l] = dW ? db[l] : nullptr;
  }
  net.L = L; net.act = act;
  return MBO_OK;
}

static void run_forward(const SNet& net, const SWs& ws, const float* states, long long R, cudaStream_t st) {
  const unsigned grid = (unsigned)((R + THREADS - 1) / THREADS);
  for (int l = 0; l < net.L; ++l)
    fwd_kernel<<<grid, THREADS, 0, st>>>(net, l, states, R, ws.Rp, l ? ws.a[l - 1] : nullptr, ws.a[l]);
}
// dz[0] holds dZ of the last layer ([1][Rp]) on entry
static void run_backward(const SNet& net, const SWs& ws, const float* states, long long R, cudaStream_t st) {
  const unsigned grid = (unsigned)((R + THREADS - 1) / THREADS);
  int cur = 0, p_off = 0;
  int offs[MAXL];
  for (int l = 0; l < net.L; ++l) { offs[l] = p_off; p_off += (net.width[l] + 1) * net.width[l + 1]; }
  for (int l = net.L - 1; l >= 0; --l) {
    wgrad_kernel<<<dim3(ws.n_chunk, net.width[l + 1]), THREADS, 0, st>>>(net, l, states, R, ws.Rp, ws.dz[cur], l ? ws.a[l - 1] : nullptr,
                                                                         ws.part, ws.P, offs[l]);
    if (l > 0) {
      bwd_kernel<<<grid, THREADS, 0, st>>>(net, l, R, ws.Rp, ws.dz[cur], ws.a[l - 1], ws.dz[cur ^ 1]);
      cur ^= 1;
    }
  }
  reduce_kernel<<<(ws.P + THREADS - 1) / THREADS, THREADS, 0, st>>>(net, ws.part, ws.n_chunk, ws.P);
}

}  // namespace upds
}  // namespace mbo

using namespace mbo;
using namespace mbo::upds;

extern "C" size_t mbo_ppo_simt_workspace_bytes(long long rows, int n_layers, const int32_t* widths, int E) {
  if (rows <= 0 || n_layers < 1 || n_layers > MAXL || !widths) return 0;
  int w[MAXL + 1];
  for (int l = 0; l <= n_layers; ++l) { if (widths[l] < 1 || widths[l] > MAXW) return 0; w[l] = widths[l]; }
  return carve(rows, n_layers, w, E, nullptr, nullptr);
}

extern "C" int mbo_ppo_policy_grad_simt(int E, int M, int F_state, int n_layers, const int32_t* widths, const int32_t* feat_col, int act,
                                        const float* states, const float* const* W, const float* const* b, const int32_t* actions,
                                        const float* advs, const float* logp_old, float epsilon, float ent_coeff, float inv_n_global,
                                        float* const* dW, float* const* db, float* terms, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  MBO_CHECK_ARG(E > 0 && M > 0, "mbo_ppo_policy_grad_simt: E, M must be > 0");
  MBO_CHECK_ARG((size_t)E * M < (1u << 30), "mbo_ppo_policy_grad_simt: too many rows");
  SNet net{};
  const int rc = fill_net(net, n_layers, widths, act, W, b, dW, db, "mbo_ppo_policy_grad_simt");
  if (rc != MBO_OK) return rc;
  const bool fwd_only = dW == nullptr;
  MBO_CHECK_ARG(states && actions && terms && workspace && feat_col, "mbo_ppo_policy_grad_simt: null pointer");
  MBO_CHECK_ARG(fwd_only || (advs && logp_old), "mbo_ppo_policy_grad_simt: null pointer");
  MBO_CHECK_ARG(F_state >= widths[0], "mbo_ppo_policy_grad_simt: more policy features than state columns");
  for (int f = 0; f < MBO_MAX_FEATURES; ++f) {
    net.col[f] = f < widths[0] ? feat_col[f] : 0;
    MBO_CHECK_ARG(net.col[f] >= 0 && net.col[f] < F_state, "mbo_ppo_policy_grad_simt: feature column out of range");
  }
  net.in_stride = F_state;
  const long long R = (long long)E * M;
  MBO_CHECK_ARG(workspace_bytes >= mbo_ppo_simt_workspace_bytes(R, n_layers, widths, E), "mbo_ppo_policy_grad_simt: workspace too small");
  MBO_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "mbo_ppo_policy_grad_simt: workspace must be 256-byte aligned");
  SWs ws;
  carve(R, n_layers, net.width, E, (unsigned char*)workspace, &ws);
  cudaStream_t st = (cudaStream_t)stream;
  run_forward(net, ws, states, R, st);
  // logits = a[L - 1] [R] (row = transition * M + candidate): per-episode log-softmax terms and dlogits into dz[0]
  mbo::upd::upd_loss_kernel<<<E, 256, 0, st>>>(M, ws.a[n_layers - 1], actions, advs, logp_old, epsilon, ent_coeff, inv_n_global,
                                               ws.dz[0], ws.terms);
  if (!fwd_only) run_backward(net, ws, states, R, st);
  MBO_CUDA(cudaMemcpyAsync(terms, ws.terms, (size_t)E * 16, cudaMemcpyDeviceToDevice, st));
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_ppo_value_grad_simt(int E, size_t state_stride, int t_col, int T_col, int n_layers, const int32_t* widths, int act,
                                       const float* states, const float* const* W, const float* const* b, const float* tdlamrets,
                                       float coeff, float* const* dW, float* const* db, float* values, float* vterms, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  MBO_CHECK_ARG(E > 0 && state_stride > 0, "mbo_ppo_value_grad_simt: E, state_stride must be > 0");
  SNet net{};
  const int rc = fill_net(net, n_layers, widths, act, W, b, dW, db, "mbo_ppo_value_grad_simt");
  if (rc != MBO_OK) return rc;
  MBO_CHECK_ARG(widths[0] == 2, "mbo_ppo_value_grad_simt: the value net has 2 inputs (t, T)");
  MBO_CHECK_ARG(states && workspace && (values || vterms), "mbo_ppo_value_grad_simt: null pointer");
  MBO_CHECK_ARG(!dW || tdlamrets, "mbo_ppo_value_grad_simt: tdlamrets needed for the gradient");
  MBO_CHECK_ARG(t_col >= 0 && T_col >= 0 && (size_t)t_col < state_stride && (size_t)T_col < state_stride,
                "mbo_ppo_value_grad_simt: t / T column out of range");
  net.col[0] = t_col; net.col[1] = T_col;
  net.in_stride = (long long)state_stride;
  MBO_CHECK_ARG(workspace_bytes >= mbo_ppo_simt_workspace_bytes(E, n_layers, widths, E), "mbo_ppo_value_grad_simt: workspace too small");
  MBO_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "mbo_ppo_value_grad_simt: workspace must be 256-byte aligned");
  SWs ws;
  carve(E, n_layers, net.width, E, (unsigned char*)workspace, &ws);
  cudaStream_t st = (cudaStream_t)stream;
  run_forward(net, ws, states, E, st);
  vloss_kernel<<<(E + 255) / 256, 256, 0, st>>>(E, ws.a[n_layers - 1], tdlamrets ? tdlamrets : ws.a[n_layers - 1], coeff,
                                                dW ? ws.dz[0] : nullptr, values, tdlamrets ? vterms : nullptr);
  if (dW) run_backward(net, ws, states, E, st);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}
