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

// Same arithmetic as load_candidate in a fraction of the instructions (one code path for head / tail and f32 / f64
// sources instead of four unrolled ones); all loads are independent, so one memory round trip covers them.
__device__ __forceinline__ void load_candidate_compact(const mbo_candidates& c, int b, int m, double* x) {
  const int D = c.D;
  const bool head = m < c.n_head;
  int r = m - c.n_head;
  int cube = 0;
  if (!head && c.local) { cube = r / c.n_tail; r -= cube * c.n_tail; }
  const size_t o = head ? ((size_t)b * c.n_head + m) * D : (size_t)r * D;
  const void* src = head ? c.head : c.tail;
  const bool f32 = head ? c.head_is_f32 != 0 : c.tail_is_f32 != 0;
  const bool local = !head && c.local;
  const double* lo = c.cubes + ((size_t)b * c.n_cubes + cube) * 2 * D;
  double l[MBO_MAX_D], h[MBO_MAX_D];
#pragma unroll
  for (int d = 0; d < MBO_MAX_D; ++d) {
    x[d] = 0.0; l[d] = 0.0; h[d] = 0.0;
    if (d < D) {
      x[d] = f32 ? (double)((const float*)src)[o + d] : ((const double*)src)[o + d];
      if (local) { l[d] = lo[d]; h[d] = lo[D + d]; }
    }
  }
  if (local) {
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d) x[d] = __dadd_rn(__dmul_rn(x[d], h[d] - l[d]), l[d]);   // util.py:35-37, two roundings
  }
}

// stationary kernels of GPy (Stationary.K_of_r): RBF, Matern-3/2, Matern-5/2 on the scaled squared distance
template <typename T> __device__ __forceinline__ T t_exp(T x);
// exp(x) for x <= 0 in fp64, branch-free (interleaves under unrolling; CUDA's exp() carries special-case
// branches): x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18 relative), 2^k by
// exponent arithmetic.  Max error ~1 ulp; arguments below -708 flush to exp(-708) ~ 3e-308 (kernel value 0
// for every practical purpose).
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, -708.0);
  const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer in the low word
  const double t = fma(x, 1.4426950408889634074, MAGIC);
  const int k = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -6.93147180369123816490e-01, x);      // ln2 high part
  r = fma(kf, -1.90821492927058770002e-10, r);             // ln2 low part
  double p = 1.6059043836821613e-10;                       // 1/13!
  p = fma(p, r, 2.08767569878681e-09);                     // 1/12!
  p = fma(p, r, 2.505210838544172e-08);                    // 1/11!
  p = fma(p, r, 2.755731922398589e-07);                    // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);                   // 1/9!
  p = fma(p, r, 2.48015873015873e-05);                     // 1/8!
  p = fma(p, r, 1.984126984126984e-04);                    // 1/7!
  p = fma(p, r, 1.388888888888889e-03);                    // 1/6!
  p = fma(p, r, 8.333333333333333e-03);                    // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp_nonpos(x); }

// Four exp_nonpos evaluations with the four dependent chains interleaved step by step: one fp64 warp has to
// cover the DFMA latency with instruction-level parallelism (the compiler does not interleave unrolled copies
// of the scalar routine by itself).  Bitwise identical to exp_nonpos per element.
template <int NV>
__device__ __forceinline__ void exp_nonpos_xn(const double* xin, double* out) {
  const double MAGIC = 6755399441055744.0;
  double x[NV], kf[NV], r[NV], p[NV];
  int k[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) x[u] = fmax(xin[u], -708.0);
#pragma unroll
  for (int u = 0; u < NV; ++u) { const double t = fma(x[u], 1.4426950408889634074, MAGIC); k[u] = __double2loint(t); kf[u] = t - MAGIC; }
#pragma unroll
  for (int u = 0; u < NV; ++u) r[u] = fma(kf[u], -6.93147180369123816490e-01, x[u]);
#pragma unroll
  for (int u = 0; u < NV; ++u) r[u] = fma(kf[u], -1.90821492927058770002e-10, r[u]);
  const double c[13] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                        2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                        4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
#pragma unroll
  for (int u = 0; u < NV; ++u) p[u] = 1.6059043836821613e-10;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
#pragma unroll
    for (int u = 0; u < NV; ++u) p[u] = fma(p[u], r[u], c[i]);
  }
#pragma unroll
  for (int u = 0; u < NV; ++u) out[u] = __hiloint2double(__double2hiint(p[u]) + (k[u] << 20), __double2loint(p[u]));
}

__device__ __forceinline__ void exp_nonpos_x4(const double* xin, double* out) { exp_nonpos_xn<4>(xin, out); }

// k(r2) for NV scaled squared distances at once (fp64), same values as kernel_of_r2<double>
template <int NV>
__device__ __forceinline__ void kernel_of_r2_xn(int kid, double var, const double* r2, double* out) {
  double arg[NV], pre[NV], e[NV];
  if (kid == MBO_KERNEL_RBF) {
#pragma unroll
    for (int u = 0; u < NV; ++u) { arg[u] = -0.5 * r2[u]; pre[u] = var; }
  } else if (kid == MBO_KERNEL_MATERN32) {
    const double s3 = 1.7320508075688772935;
#pragma unroll
    for (int u = 0; u < NV; ++u) { const double r = sqrt(r2[u]); arg[u] = -s3 * r; pre[u] = var * (1.0 + s3 * r); }
  } else {
    const double s5 = 2.2360679774997896964;
#pragma unroll
    for (int u = 0; u < NV; ++u) { const double r = sqrt(r2[u]); arg[u] = -s5 * r; pre[u] = var * (1.0 + s5 * r + (5.0 / 3.0) * r2[u]); }
  }
  exp_nonpos_xn<NV>(arg, e);
#pragma unroll
  for (int u = 0; u < NV; ++u) out[u] = pre[u] * e[u];
}
template <> __device__ __forceinline__ float t_exp<float>(float x) { return __expf(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }

template <typename T>
__device__ __forceinline__ T kernel_of_r2(int kid, T var, T r2) {
  if (kid == MBO_KERNEL_RBF) return var * t_exp<T>(T(-0.5) * r2);
  T r = t_sqrt<T>(r2);
  if (kid == MBO_KERNEL_MATERN32) {
    const T s3 = T(1.7320508075688772935);
    return var * (T(1) + s3 * r) * t_exp<T>(-s3 * r);
  }
  const T s5 = T(2.2360679774997896964);
  return var * (T(1) + s5 * r + T(5.0 / 3.0) * r2) * t_exp<T>(-s5 * r);
}

__device__ __forceinline__ void kernel_of_r2_x4(int kid, double var, const double* r2, double* out) { kernel_of_r2_xn<4>(kid, var, r2, out); }

int check_candidates(const mbo_candidates* c);

}  // namespace mbo
