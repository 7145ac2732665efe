// common.cuh -- shared helpers for libmetabo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/metabo_b200.h"

namespace mbo {

void set_error(const char* fmt, ...);

#define MBO_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      mbo::set_error(__VA_ARGS__);          \
      return MBO_ERR_ARG;                   \
    }                                       \
  } while (0)

#define MBO_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      mbo::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MBO_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define MBO_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      mbo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MBO_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// ---- gp_state blob layout (doubles), see gp.cu ------------------------------------------------
constexpr int GP_HDR = 16;  // [0] n, [1] prior mean m, [2] variance, [3] status, [4..4+D) lengthscale
__host__ __device__ inline int gp_np(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int gp_linv_doubles(int np_) { int nb = np_ >> 3; return 32 * nb * (nb + 1); }
__host__ __device__ inline int gp_linv_block_off(int ib) { int nb = ib >> 3; return 32 * nb * (nb + 1); }
__host__ __device__ inline int gp_off_xs() { return GP_HDR; }
__host__ __device__ inline int gp_off_beta(int np_, int D) { return GP_HDR + np_ * D; }
__host__ __device__ inline int gp_off_linv(int np_, int D) { return GP_HDR + np_ * D + np_; }
__host__ __device__ inline int gp_state_doubles(int nmax, int D) {
  int np_ = gp_np(nmax);
  return GP_HDR + np_ * D + np_ + gp_linv_doubles(np_);
}

// candidate m of episode b -> coordinates (fp64), shared by every kernel that touches candidates
__device__ __forceinline__ void load_candidate(const mbo_candidates& c, int b, int m, double* x) {
  const int D = c.D;
  if (m < c.n_head) {
    size_t o = ((size_t)b * c.n_head + m) * D;
    if (c.head_is_f32) {
      const float* h = (const float*)c.head;
#pragma unroll
      for (int d = 0; d < MBO_MAX_D; ++d) if (d < D) x[d] = (double)h[o + d];
    } else {
      const double* h = (const double*)c.head;
#pragma unroll
      for (int d = 0; d < MBO_MAX_D; ++d) if (d < D) x[d] = h[o + d];
    }
    return;
  }
  int r = m - c.n_head;
  int cube = 0;
  if (c.local) { cube = r / c.n_tail; r -= cube * c.n_tail; }
  size_t o = (size_t)r * D;
  if (c.tail_is_f32) {
    const float* t = (const float*)c.tail;
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d) if (d < D) x[d] = (double)t[o + d];
  } else {
    const double* t = (const double*)c.tail;
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d) if (d < D) x[d] = t[o + d];
  }
  if (c.local) {
    const double* lo = c.cubes + ((size_t)b * c.n_cubes + cube) * 2 * D;
    const double* hi = lo + D;
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d)
      if (d < D) x[d] = __dadd_rn(__dmul_rn(x[d], hi[d] - lo[d]), lo[d]);  // util.py:35-37: X * ptp + lo, two roundings
  }
}

int check_candidates(const mbo_candidates* c);

}  // namespace mbo
