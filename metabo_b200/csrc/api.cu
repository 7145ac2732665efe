// api.cu -- error reporting, policy handle, candidate helpers of the C ABI.
#include <stdarg.h>
#include <vector>
#include "common.cuh"
#include "policy.cuh"

namespace mbo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_candidates(const mbo_candidates* c) {
  MBO_CHECK_ARG(c->M > 0, "candidates: M must be > 0");
  MBO_CHECK_ARG(c->D > 0 && c->D <= MBO_MAX_D, "candidates: D must be in 1..%d", MBO_MAX_D);
  MBO_CHECK_ARG(c->n_head >= 0 && c->n_tail >= 0, "candidates: negative part size");
  if (c->local) {
    MBO_CHECK_ARG(c->n_cubes > 0 && c->cubes, "candidates: local set needs cubes");
    MBO_CHECK_ARG(c->M == c->n_head + c->n_cubes * c->n_tail, "candidates: M != n_head + n_cubes * n_tail");
  } else {
    MBO_CHECK_ARG(c->M == c->n_head + c->n_tail, "candidates: M != n_head + n_tail");
  }
  MBO_CHECK_ARG(c->n_head == 0 || c->head, "candidates: head pointer is null");
  MBO_CHECK_ARG(c->n_tail == 0 || c->tail, "candidates: tail pointer is null");
  return MBO_OK;
}

__global__ void gather_kernel(int B, mbo_candidates cand, int k, const int32_t* __restrict__ idx,
                              double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * k) return;
  int b = i / k;
  int m = idx[i];
  double x[MBO_MAX_D];
#pragma unroll
  for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
  m = m < 0 ? 0 : (m >= cand.M ? cand.M - 1 : m);
  load_candidate(cand, b, m, x);
#pragma unroll
  for (int d = 0; d < MBO_MAX_D; ++d)
    if (d < cand.D) out[(size_t)i * cand.D + d] = x[d];
}

__global__ void cubes_kernel(int total, int D, const double* __restrict__ x, double half, double* __restrict__ cubes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // over (b, j, d)
  if (i >= total * D) return;
  int p = i / D, d = i - p * D;
  double v = x[i];
  cubes[(size_t)p * 2 * D + d] = fmax(v - half, 0.0);
  cubes[(size_t)p * 2 * D + D + d] = fmin(v + half, 1.0);
}

}  // namespace mbo

using namespace mbo;

extern "C" const char* mbo_last_error(void) { return g_err; }
extern "C" int mbo_version(void) { return 100; }

extern "C" int mbo_policy_create(mbo_policy** out) {
  MBO_CHECK_ARG(out, "mbo_policy_create: null out");
  *out = new (std::nothrow) mbo_policy();
  if (!*out) { set_error("mbo_policy_create: out of host memory"); return MBO_ERR_CUDA; }
  return MBO_OK;
}

static void policy_free(mbo_policy* p) {
  for (int l = 0; l < MBO_MAX_LAYERS; ++l) {
    if (p->d_Wt[l]) cudaFree(p->d_Wt[l]);
    if (p->d_b[l]) cudaFree(p->d_b[l]);
    p->d_Wt[l] = p->d_b[l] = nullptr;
  }
  if (p->d_W_raw) cudaFree(p->d_W_raw);
  if (p->d_tc) cudaFree(p->d_tc);
  p->d_W_raw = nullptr;
  p->d_tc = nullptr;
  p->tc_ready = 0;
  p->alloc_sig = 0;
}

extern "C" void mbo_policy_destroy(mbo_policy* p) {
  if (!p) return;
  policy_free(p);
  delete p;
}

namespace mbo {
__global__ void transpose_kernel(int out, int in, const float* __restrict__ W, float* __restrict__ Wt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out * in) return;
  int o = i / in, k = i - o * in;
  Wt[(size_t)k * out + o] = W[i];
}
int tc_pack_weights(mbo_policy* p, cudaStream_t s);   // neural_af_tc.cu
}  // namespace mbo

extern "C" int mbo_policy_set_weights(mbo_policy* p, int n_layers, const int32_t* widths, int activation,
                                      const float* const* W, const float* const* b, int on_device,
                                      void* stream) {
  MBO_CHECK_ARG(p && widths && W && b, "mbo_policy_set_weights: null pointer");
  MBO_CHECK_ARG(n_layers >= 1 && n_layers <= MBO_MAX_LAYERS, "mbo_policy_set_weights: 1..%d layers", MBO_MAX_LAYERS);
  MBO_CHECK_ARG(activation == MBO_ACT_RELU || activation == MBO_ACT_TANH, "mbo_policy_set_weights: bad activation");
  for (int l = 0; l <= n_layers; ++l)
    MBO_CHECK_ARG(widths[l] >= 1 && widths[l] <= 1024, "mbo_policy_set_weights: width %d out of range", widths[l]);
  cudaStream_t s = (cudaStream_t)stream;
  // (re)allocate only when the architecture changes: weight refreshes after optimiser steps are copies
  uint64_t sig = 1469598103934665603ull;
  for (int l = 0; l <= n_layers; ++l) sig = (sig ^ (uint64_t)widths[l]) * 1099511628211ull;
  sig = (sig ^ (uint64_t)n_layers) * 1099511628211ull;
  size_t raw_floats = 0;
  for (int l = 0; l < n_layers; ++l) raw_floats += (size_t)widths[l] * widths[l + 1];
  if (sig != p->alloc_sig) {
    policy_free(p);
    for (int l = 0; l < n_layers; ++l) {
      MBO_CUDA(cudaMalloc(&p->d_Wt[l], sizeof(float) * widths[l] * widths[l + 1]));
      MBO_CUDA(cudaMalloc(&p->d_b[l], sizeof(float) * widths[l + 1]));
    }
    MBO_CUDA(cudaMalloc(&p->d_W_raw, sizeof(float) * raw_floats));
    p->alloc_sig = sig;
  }
  p->n_layers = n_layers;
  p->activation = activation;
  for (int l = 0; l <= n_layers; ++l) p->widths[l] = widths[l];
  size_t off = 0;
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int l = 0; l < n_layers; ++l) {
    int in = widths[l], out = widths[l + 1];
    float* raw = p->d_W_raw + off;
    p->raw_off[l] = off;
    MBO_CUDA(cudaMemcpyAsync(raw, W[l], sizeof(float) * in * out, kind, s));
    MBO_CUDA(cudaMemcpyAsync(p->d_b[l], b[l], sizeof(float) * out, kind, s));
    int tot = in * out;
    transpose_kernel<<<(tot + 255) / 256, 256, 0, s>>>(out, in, raw, p->d_Wt[l]);
    MBO_LAUNCH_CHECK();
    off += (size_t)in * out;
  }
  p->tc_ready = 0;
  return tc_pack_weights(p, s);
}

extern "C" int mbo_candidates_gather(int B, const mbo_candidates* cand, int k, const int32_t* idx,
                                     double* x_out, void* stream) {
  MBO_CHECK_ARG(B > 0 && k > 0 && cand && idx && x_out, "mbo_candidates_gather: bad argument");
  int rc = check_candidates(cand);
  if (rc) return rc;
  int tot = B * k;
  gather_kernel<<<(tot + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, *cand, k, idx, x_out);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_cubes_around(int B, int k, int D, const double* x, double diam, double* cubes, void* stream) {
  MBO_CHECK_ARG(B > 0 && k > 0 && D > 0 && D <= MBO_MAX_D && x && cubes, "mbo_cubes_around: bad argument");
  int tot = B * k * D;
  cubes_kernel<<<(tot + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B * k, D, x, 0.5 * diam, cubes);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}
