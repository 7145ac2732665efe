// mlp_simt.cu -- CUDA-core fp32 NeuralAF evaluation (validation mode, any architecture).
//
// Replaces get_state feature packing (metabo/environment/metabo_gym.py:624-677) followed by
// NeuralAF.forward (metabo/policies/policies.py:104-125) -> MLP.forward (mlp.py:52-61).
// fp32 FMA with sequential accumulation over the input index: agrees with torch-CPU fp32 to
// ~1e-6 relative (summation order differs).  The throughput path is neural_af_tc.cu.
#include "common.cuh"
#include "policy.cuh"
#include <stdlib.h>

namespace mbo {

constexpr int ST = 256;   // threads
constexpr int TILE = 32;  // rows (candidates) per CTA

struct SimtNet {
  int n_layers;
  int widths[MBO_MAX_LAYERS + 1];
  int activation;
  const float* Wt[MBO_MAX_LAYERS];
  const float* b[MBO_MAX_LAYERS];
};

struct FeatSpec {
  int F;
  int src[MBO_MAX_FEATURES];   // MBO_FEAT_* codes, or state-column indices in dense mode
  const float* dense;          // [B, M, dense_F] materialised states, or nullptr
  int dense_F;
};

__device__ __forceinline__ float act_fn(int a, float v) { return a == MBO_ACT_RELU ? fmaxf(v, 0.f) : tanhf(v); }

// rows: either (episode, candidate) pairs with features assembled on the fly (direct == nullptr)
// or rows of a dense [R, F] fp32 matrix (direct != nullptr; the value net path)
__global__ void __launch_bounds__(ST)
mlp_simt_kernel(SimtNet net, FeatSpec fs, mbo_candidates cand, int tiles_per_episode,
                const float* __restrict__ mean, const float* __restrict__ std_, const float* __restrict__ epi,
                const float* __restrict__ direct, int R, int hmax, float* __restrict__ out) {
  extern __shared__ float smf[];
  float* bufA = smf;                 // [hmax][TILE]
  float* bufB = smf + hmax * TILE;
  const int tid = threadIdx.x;
  int b = 0, row0;
  int rows_valid;
  if (direct) {
    row0 = blockIdx.x * TILE;
    rows_valid = min(TILE, R - row0);
  } else {
    b = blockIdx.x / tiles_per_episode;
    row0 = (blockIdx.x - b * tiles_per_episode) * TILE;
    rows_valid = min(TILE, cand.M - row0);
  }
  const int F = net.widths[0];
  // ---- layer-0 input ----
  if (direct) {
    for (int e = tid; e < F * TILE; e += ST) {
      int f = e / TILE, c = e - f * TILE;
      bufA[f * TILE + c] = c < rows_valid ? direct[(size_t)(row0 + c) * F + f] : 0.f;
    }
  } else {
    if (tid < TILE) {
      int c = tid;
      double x[MBO_MAX_D];
#pragma unroll
      for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
      float mu = 0.f, sd = 0.f;
      if (c < rows_valid && !fs.dense) {
        load_candidate(cand, b, row0 + c, x);
        mu = mean[(size_t)b * cand.M + row0 + c];
        sd = std_[(size_t)b * cand.M + row0 + c];
      }
      for (int f = 0; f < F; ++f) {
        int s = fs.src[f];
        float v;
        if (fs.dense) v = c < rows_valid ? fs.dense[((size_t)b * cand.M + row0 + c) * fs.dense_F + s] : 0.f;
        else if (s == MBO_FEAT_MEAN) v = mu;
        else if (s == MBO_FEAT_STD) v = sd;
        else if (s >= MBO_FEAT_X0 && s < MBO_FEAT_X0 + MBO_MAX_D) {
          v = 0.f;
#pragma unroll
          for (int d = 0; d < MBO_MAX_D; ++d) if (s - MBO_FEAT_X0 == d) v = (float)x[d];
        } else v = epi[b * 4 + (s - MBO_FEAT_INCUMBENT)];
        bufA[f * TILE + c] = c < rows_valid ? v : 0.f;
      }
    }
  }
  __syncthreads();
  float* in = bufA;
  float* outb = bufB;
  for (int l = 0; l < net.n_layers; ++l) {
    const int K = net.widths[l], N = net.widths[l + 1];
    const float* __restrict__ Wt = net.Wt[l];
    const float* __restrict__ bias = net.b[l];
    const bool last = (l == net.n_layers - 1);
    if (N >= 8) {
      for (int o = tid; o < N; o += ST) {
        float acc[TILE];
        const float b0 = bias[o];
#pragma unroll
        for (int c = 0; c < TILE; ++c) acc[c] = b0;
        for (int k = 0; k < K; ++k) {
          const float w = Wt[(size_t)k * N + o];
          const float4* a4 = (const float4*)(in + k * TILE);
#pragma unroll
          for (int c4 = 0; c4 < TILE / 4; ++c4) {
            float4 a = a4[c4];
            acc[c4 * 4 + 0] = fmaf(w, a.x, acc[c4 * 4 + 0]);
            acc[c4 * 4 + 1] = fmaf(w, a.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(w, a.z, acc[c4 * 4 + 2]);
            acc[c4 * 4 + 3] = fmaf(w, a.w, acc[c4 * 4 + 3]);
          }
        }
        if (last) {
#pragma unroll
          for (int c = 0; c < TILE; ++c) outb[o * TILE + c] = acc[c];
        } else {
#pragma unroll
          for (int c = 0; c < TILE; ++c) outb[o * TILE + c] = act_fn(net.activation, acc[c]);
        }
      }
    } else {
      // narrow output (the 1-unit head): 8 lanes per row split K
      for (int o = 0; o < N; ++o) {
        int c = tid >> 3, part = tid & 7;
        float s = 0.f;
        for (int k = part; k < K; k += 8) s = fmaf(Wt[(size_t)k * N + o], in[k * TILE + c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (part == 0) {
          float v = s + bias[o];
          outb[o * TILE + c] = last ? v : act_fn(net.activation, v);
        }
      }
    }
    __syncthreads();
    float* t = in; in = outb; outb = t;
  }
  // `in` now holds the output [N_out][TILE]; N_out == 1 for every MetaBO net
  const int NO = net.widths[net.n_layers];
  for (int e = tid; e < NO * TILE; e += ST) {
    int o = e / TILE, c = e - o * TILE;
    if (c < rows_valid) {
      size_t r = direct ? (size_t)(row0 + c) : (size_t)b * cand.M + row0 + c;
      out[r * NO + o] = in[o * TILE + c];
    }
  }
}

// Dense rows, few of them (the value net: [B, 2] inputs, policies.py:117-121).  mlp_simt_kernel streams every weight
// from L2 inside its k loop; with 32 rows per CTA a batch of 60-1024 rows is 2-32 CTAs, each paying ~800 dependent
// L2 latencies (130-150 us).  Here a CTA handles VR rows and stages one layer's weights in shared memory with
// coalesced 16-byte loads before using them; the arithmetic (bias first, k ascending, one fma per term) and therefore
// the result is the same as mlp_simt_kernel's.
constexpr int VR = 8;
__global__ void __launch_bounds__(ST)
mlp_rows_kernel(SimtNet net, const float* __restrict__ direct, int R, int hmax, int wmax, float* __restrict__ out) {
  extern __shared__ __align__(16) float smv[];
  float* sW = smv;                      // [K][N] of the current layer
  float* bufA = smv + wmax;             // [hmax][VR]
  float* bufB = bufA + hmax * VR;
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * VR;
  const int rows_valid = min(VR, R - row0);
  const int F = net.widths[0];
  for (int e = tid; e < F * VR; e += ST) {
    const int f = e / VR, c = e - f * VR;
    bufA[f * VR + c] = c < rows_valid ? direct[(size_t)(row0 + c) * F + f] : 0.f;
  }
  float* in = bufA;
  float* outb = bufB;
  for (int l = 0; l < net.n_layers; ++l) {
    const int K = net.widths[l], N = net.widths[l + 1];
    const float* __restrict__ Wt = net.Wt[l];
    const int cnt = K * N;
    __syncthreads();                    // previous layer done with sW / wrote `in`
    if ((cnt & 3) == 0 && (((size_t)Wt) & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(Wt);
      float4* dst = reinterpret_cast<float4*>(sW);
      for (int e = tid; e < cnt / 4; e += ST) dst[e] = src[e];
    } else {
      for (int e = tid; e < cnt; e += ST) sW[e] = Wt[e];
    }
    __syncthreads();
    const bool last = (l == net.n_layers - 1);
    for (int o = tid; o < N; o += ST) {
      float acc[VR];
      const float b0 = net.b[l][o];
#pragma unroll
      for (int c = 0; c < VR; ++c) acc[c] = b0;
      for (int k = 0; k < K; ++k) {
        const float w = sW[k * N + o];
        const float4* a4 = reinterpret_cast<const float4*>(in + k * VR);
#pragma unroll
        for (int c4 = 0; c4 < VR / 4; ++c4) {
          const float4 a = a4[c4];
          acc[c4 * 4 + 0] = fmaf(w, a.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(w, a.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(w, a.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(w, a.w, acc[c4 * 4 + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < VR; ++c) outb[o * VR + c] = last ? acc[c] : act_fn(net.activation, acc[c]);
    }
    float* t = in; in = outb; outb = t;
  }
  __syncthreads();
  const int NO = net.widths[net.n_layers];
  for (int e = tid; e < NO * VR; e += ST) {
    const int o = e / VR, c = e - o * VR;
    if (c < rows_valid) out[(size_t)(row0 + c) * NO + o] = in[o * VR + c];
  }
}


// ---- narrow nets (every width <= 32: train_metabo_gps.py's 4 -> 20 -> 20 -> 1 policy, policies.py:69-74) ----------------
// mlp_simt_kernel maps a thread to an OUTPUT unit: with 20-wide layers 20 of its 256 threads work (measured: 885 us per
// 2 M candidates, 37 % of a C4 rollout step).  Here a thread owns a candidate row, its activations live in registers, and the
// weights -- staged once per CTA, rows padded to a multiple of four outputs -- are warp-broadcast LDS.128: four fma per
// load.  Same arithmetic per output as mlp_simt_kernel's wide branch (bias first, k ascending, one fma per term).
constexpr int NW_MAX = 32;          // widest layer / most policy features
constexpr int NW_THREADS = 256;

__global__ void __launch_bounds__(NW_THREADS)
mlp_narrow_kernel(SimtNet net, FeatSpec fs, mbo_candidates cand, const float* __restrict__ mean, const float* __restrict__ std_,
                  const float* __restrict__ epi, const float* __restrict__ direct, long long R, float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];      // per layer: W^T [K][N4] then bias [N4]
  const int tid = threadIdx.x;
  {
    int off = 0;
    for (int l = 0; l < net.n_layers; ++l) {
      const int K = net.widths[l], N = net.widths[l + 1], N4 = (N + 3) & ~3;
      for (int e = tid; e < K * N4; e += NW_THREADS) {
        const int k = e / N4, o = e - k * N4;
        sw[off + e] = o < N ? net.Wt[l][(size_t)k * N + o] : 0.f;
      }
      for (int o = tid; o < N4; o += NW_THREADS) sw[off + K * N4 + o] = o < N ? net.b[l][o] : 0.f;
      off += (K + 1) * N4;
    }
  }
  __syncthreads();
  const long long row = (long long)blockIdx.x * NW_THREADS + tid;
  if (row >= R) return;
  const int F = net.widths[0];
  float h[NW_MAX];
#pragma unroll
  for (int i = 0; i < NW_MAX; ++i) h[i] = 0.f;
  if (direct) {
#pragma unroll
    for (int f = 0; f < NW_MAX; ++f) if (f < F) h[f] = direct[(size_t)row * F + f];
  } else {
    const int b = (int)(row / cand.M), m = (int)(row - (long long)b * cand.M);
    double x[MBO_MAX_D];
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
    float mu = 0.f, sd = 0.f;
    if (!fs.dense) {
      load_candidate(cand, b, m, x);
      mu = mean[row];
      sd = std_[row];
    }
#pragma unroll
    for (int f = 0; f < NW_MAX; ++f) {
      if (f < F) {
        const int sidx = fs.src[f];
        float v;
        if (fs.dense) v = fs.dense[(size_t)row * fs.dense_F + sidx];
        else if (sidx == MBO_FEAT_MEAN) v = mu;
        else if (sidx == MBO_FEAT_STD) v = sd;
        else if (sidx >= MBO_FEAT_X0 && sidx < MBO_FEAT_X0 + MBO_MAX_D) {
          v = 0.f;
#pragma unroll
          for (int d = 0; d < MBO_MAX_D; ++d) if (sidx - MBO_FEAT_X0 == d) v = (float)x[d];
        } else v = epi[b * 4 + (sidx - MBO_FEAT_INCUMBENT)];
        h[f] = v;
      }
    }
  }
  int off = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    const int K = net.widths[l], N = net.widths[l + 1], N4 = (N + 3) & ~3;
    const float4* W4 = reinterpret_cast<const float4*>(sw + off);
    const float4* B4 = reinterpret_cast<const float4*>(sw + off + K * N4);
    const bool last = l == net.n_layers - 1;
    float o_[NW_MAX];
#pragma unroll
    for (int o4 = 0; o4 < NW_MAX / 4; ++o4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * o4 < N) {
        acc = B4[o4];
#pragma unroll
        for (int k = 0; k < NW_MAX; ++k) {
          if (k < K) {
            const float4 w = W4[k * (N4 >> 2) + o4];
            acc.x = fmaf(w.x, h[k], acc.x); acc.y = fmaf(w.y, h[k], acc.y);
            acc.z = fmaf(w.z, h[k], acc.z); acc.w = fmaf(w.w, h[k], acc.w);
          }
        }
        if (!last) {
          acc.x = act_fn(net.activation, acc.x); acc.y = act_fn(net.activation, acc.y);
          acc.z = act_fn(net.activation, acc.z); acc.w = act_fn(net.activation, acc.w);
        }
      }
      o_[4 * o4 + 0] = acc.x; o_[4 * o4 + 1] = acc.y; o_[4 * o4 + 2] = acc.z; o_[4 * o4 + 3] = acc.w;
    }
#pragma unroll
    for (int i = 0; i < NW_MAX; ++i) h[i] = i < N ? o_[i] : 0.f;
    off += (K + 1) * N4;
  }
  const int NO = net.widths[net.n_layers];
#pragma unroll
  for (int o = 0; o < NW_MAX; ++o) if (o < NO) out[(size_t)row * NO + o] = h[o];
}

int simt_forward(const mbo_policy* p, int B, int F, const int32_t* feat_src, const mbo_candidates* cand,
                 const float* mean, const float* std_, const float* epi, const float* direct, int R,
                 float* out, cudaStream_t s, const float* dense = nullptr, int dense_F = 0) {
  SimtNet net;
  net.n_layers = p->n_layers;
  net.activation = p->activation;
  int hmax = 0;
  for (int l = 0; l <= p->n_layers; ++l) { net.widths[l] = p->widths[l]; hmax = max(hmax, p->widths[l]); }
  for (int l = 0; l < p->n_layers; ++l) { net.Wt[l] = p->d_Wt[l]; net.b[l] = p->d_b[l]; }
  FeatSpec fs;
  fs.F = F;
  fs.dense = dense;
  fs.dense_F = dense_F;
  mbo_candidates c0;
  memset(&c0, 0, sizeof(c0));
  int grid, tiles = 1;
  // narrow nets with many rows: thread = row (mlp_narrow_kernel)
  bool narrow = hmax <= NW_MAX && getenv("MBO_MLP_NO_NARROW") == nullptr;
  const long long rows = direct ? (long long)R : (long long)B * cand->M;
  if (narrow && rows >= 4096) {
    size_t smem_w = 0;
    for (int l = 0; l < p->n_layers; ++l) smem_w += sizeof(float) * (size_t)(p->widths[l] + 1) * ((p->widths[l + 1] + 3) & ~3);
    if (!direct) {
      for (int f = 0; f < F; ++f) fs.src[f] = feat_src[f];
      c0 = *cand;
    }
    mlp_narrow_kernel<<<(unsigned)((rows + NW_THREADS - 1) / NW_THREADS), NW_THREADS, smem_w, s>>>(net, fs, c0, mean, std_, epi, direct, rows, out);
    MBO_LAUNCH_CHECK();
    return MBO_OK;
  }
  if (direct) {
    // few dense rows: weights staged in shared memory when the widest layer fits
    int wmax = 0;
    for (int l = 0; l < p->n_layers; ++l) wmax = max(wmax, p->widths[l] * p->widths[l + 1]);
    wmax = (wmax + 3) & ~3;
    const size_t smem_rows = sizeof(float) * ((size_t)wmax + 2 * (size_t)hmax * VR);
    if (R <= 2048 && smem_rows <= 200 * 1024) {
      MBO_CUDA(cudaFuncSetAttribute(mlp_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
      mlp_rows_kernel<<<(R + VR - 1) / VR, ST, smem_rows, s>>>(net, direct, R, hmax, wmax, out);
      MBO_LAUNCH_CHECK();
      return MBO_OK;
    }
    grid = (R + TILE - 1) / TILE;
  } else {
    for (int f = 0; f < F; ++f) fs.src[f] = feat_src[f];
    c0 = *cand;
    tiles = (cand->M + TILE - 1) / TILE;
    grid = B * tiles;
  }
  size_t smem = sizeof(float) * 2 * hmax * TILE;
  MBO_CUDA(cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp_simt_kernel<<<grid, ST, smem, s>>>(net, fs, c0, tiles, mean, std_, epi, direct, R, hmax, out);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

int tc_forward(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src, const mbo_candidates* cand,
               const float* mean, const float* std_, const float* epi, float* logits, void* ws, size_t ws_bytes,
               cudaStream_t s, const float* dense = nullptr, int dense_F = 0, int topk = 0, int32_t* topk_idx = nullptr,
               float* topk_val = nullptr);   // neural_af_tc.cu
size_t tc_workspace_bytes(int B, int M);

}  // namespace mbo

using namespace mbo;

extern "C" size_t mbo_af_workspace_bytes(int mode, int B, int M) {
  if (mode == MBO_MLP_FP32) return 256;
  return tc_workspace_bytes(B, M);     // error flag + per-warp top-k partial lists of the fused epilogue
}

static int check_af_args(const char* who, const mbo_policy* p, int B, int F, const int32_t* feat_src,
                         const mbo_candidates* cand, const float* mean, const float* std_, const float* epi) {
  MBO_CHECK_ARG(p && p->n_layers > 0, "%s: policy has no weights", who);
  MBO_CHECK_ARG(B > 0 && cand && mean && std_ && epi && feat_src, "%s: null pointer", who);
  MBO_CHECK_ARG(F == p->widths[0] && F <= MBO_MAX_FEATURES, "%s: F=%d does not match the policy input width %d", who, F, p->widths[0]);
  MBO_CHECK_ARG(p->widths[p->n_layers] == 1, "%s: policy head must have one output", who);
  int rc = check_candidates(cand);
  if (rc) return rc;
  for (int f = 0; f < F; ++f) {
    int s = feat_src[f];
    bool ok = s == MBO_FEAT_MEAN || s == MBO_FEAT_STD || (s >= MBO_FEAT_X0 && s < MBO_FEAT_X0 + cand->D) ||
              (s >= MBO_FEAT_INCUMBENT && s <= MBO_FEAT_BUDGET);
    MBO_CHECK_ARG(ok, "%s: bad feature source %d at column %d", who, s, f);
  }
  return MBO_OK;
}

extern "C" int mbo_af_topk(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src,
                           const mbo_candidates* cand, const float* mean, const float* std_, const float* epi, int k,
                           int32_t* topk_idx, float* topk_val, float* logits, void* workspace, size_t workspace_bytes,
                           void* stream) {
  int rc = check_af_args("mbo_af_topk", p, B, F, feat_src, cand, mean, std_, epi);
  if (rc) return rc;
  MBO_CHECK_ARG(mode == MBO_MLP_TF32 || mode == MBO_MLP_BF16X3 || mode == MBO_MLP_FP16, "mbo_af_topk: tensor-core modes only (use mbo_af_logits + mbo_logits_topk)");
  MBO_CHECK_ARG(topk_idx && k >= 1, "mbo_af_topk: bad k / null output");
  return tc_forward(p, mode, B, F, feat_src, cand, mean, std_, epi, logits, workspace, workspace_bytes,
                    (cudaStream_t)stream, nullptr, 0, k, topk_idx, topk_val);
}

extern "C" int mbo_af_logits(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src,
                             const mbo_candidates* cand, const float* mean, const float* std_,
                             const float* epi, float* logits, void* workspace, size_t workspace_bytes,
                             void* stream) {
  MBO_CHECK_ARG(p && p->n_layers > 0, "mbo_af_logits: policy has no weights");
  MBO_CHECK_ARG(B > 0 && cand && mean && std_ && epi && logits && feat_src, "mbo_af_logits: null pointer");
  MBO_CHECK_ARG(F == p->widths[0] && F <= MBO_MAX_FEATURES, "mbo_af_logits: F=%d does not match the policy input width %d", F, p->widths[0]);
  MBO_CHECK_ARG(p->widths[p->n_layers] == 1, "mbo_af_logits: policy head must have one output");
  int rc = check_candidates(cand);
  if (rc) return rc;
  for (int f = 0; f < F; ++f) {
    int s = feat_src[f];
    bool ok = s == MBO_FEAT_MEAN || s == MBO_FEAT_STD || (s >= MBO_FEAT_X0 && s < MBO_FEAT_X0 + cand->D) ||
              (s >= MBO_FEAT_INCUMBENT && s <= MBO_FEAT_BUDGET);
    MBO_CHECK_ARG(ok, "mbo_af_logits: bad feature source %d at column %d", s, f);
  }
  if (mode == MBO_MLP_FP32)
    return simt_forward(p, B, F, feat_src, cand, mean, std_, epi, nullptr, 0, logits, (cudaStream_t)stream);
  if (mode == MBO_MLP_TF32 || mode == MBO_MLP_BF16X3 || mode == MBO_MLP_FP16)
    return tc_forward(p, mode, B, F, feat_src, cand, mean, std_, epi, logits, workspace, workspace_bytes,
                      (cudaStream_t)stream);
  set_error("mbo_af_logits: unknown mode %d", mode);
  return MBO_ERR_ARG;
}

extern "C" int mbo_af_logits_dense(const mbo_policy* p, int mode, int B, int M, int F_state, int F,
                                   const int32_t* cols, const float* states, float* logits, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  MBO_CHECK_ARG(p && p->n_layers > 0, "mbo_af_logits_dense: policy has no weights");
  MBO_CHECK_ARG(B > 0 && M > 0 && states && logits && cols, "mbo_af_logits_dense: bad argument");
  MBO_CHECK_ARG(F == p->widths[0] && F <= MBO_MAX_FEATURES, "mbo_af_logits_dense: F=%d does not match the policy input width %d", F, p->widths[0]);
  MBO_CHECK_ARG(p->widths[p->n_layers] == 1, "mbo_af_logits_dense: policy head must have one output");
  for (int f = 0; f < F; ++f) MBO_CHECK_ARG(cols[f] >= 0 && cols[f] < F_state, "mbo_af_logits_dense: column index out of range");
  mbo_candidates c;
  memset(&c, 0, sizeof(c));
  c.M = M;
  c.D = 1;
  if (mode == MBO_MLP_FP32)
    return simt_forward(p, B, F, cols, &c, nullptr, nullptr, nullptr, nullptr, 0, logits, (cudaStream_t)stream, states, F_state);
  if (mode == MBO_MLP_TF32 || mode == MBO_MLP_BF16X3 || mode == MBO_MLP_FP16)
    return tc_forward(p, mode, B, F, cols, &c, nullptr, nullptr, nullptr, logits, workspace, workspace_bytes,
                      (cudaStream_t)stream, states, F_state);
  set_error("mbo_af_logits_dense: unknown mode %d", mode);
  return MBO_ERR_ARG;
}

extern "C" int mbo_value_net(const mbo_policy* p, int B, const float* tT, float* values, void* stream) {
  MBO_CHECK_ARG(p && p->n_layers > 0 && B > 0 && tT && values, "mbo_value_net: bad argument");
  MBO_CHECK_ARG(p->widths[p->n_layers] == 1, "mbo_value_net: head must have one output");
  return simt_forward(p, B, p->widths[0], nullptr, nullptr, nullptr, nullptr, nullptr, tT, B, values,
                      (cudaStream_t)stream);
}
