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
// exp(x) for x <= 0 in fp64, branch-free (CUDA's exp() carries special-case branches and its unrolled copies are not
// interleaved by the compiler): x = (64 e + j) ln2/64 + r with |r| <= ln2/128, exp(x) = 2^e * 2^(j/64) * exp(r);
// 2^(j/64) from a 64-entry table, exp(r) by a degree-5 Taylor polynomial (truncation 3.5e-17 relative), 2^e by
// exponent arithmetic: 10 fp64 operations instead of the 18 of a table-free degree-13 version.  Error <= ~2 ulp;
// arguments below -708 flush to exp(-708) ~ 3e-308 (kernel value 0 for every practical purpose).
__device__ static const double MBO_EXP2_TAB[64] = {
    1, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.1023825833078409, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.2021567314527031, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.2553807570246911, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.3396675240533029,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.5590044002378369, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.6457554781539649, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.7186192981224779, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.9784560263879509};
constexpr double MBO_EXP_INV = 92.332482616893657;      // 64 / ln2
constexpr double MBO_EXP_HI = 0.010830424696223417;       // ln2 / 64, upper 36 bits (k * HI is exact)
constexpr double MBO_EXP_LO = 2.5728046223276691e-14;       // ln2 / 64 - HI
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, -708.0);
  const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer in the low word
  const double t = fma(x, MBO_EXP_INV, MAGIC);
  const int k = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -MBO_EXP_HI, x);
  r = fma(kf, -MBO_EXP_LO, r);
  double p = 8.333333333333333e-03;                        // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= MBO_EXP2_TAB[k & 63];
  return __hiloint2double(__double2hiint(p) + ((k >> 6) << 20), __double2loint(p));
}
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp_nonpos(x); }

// NV exp_nonpos evaluations with the NV dependent chains interleaved step by step: one fp64 warp has to cover the
// DFMA latency with instruction-level parallelism.  Bitwise identical to exp_nonpos per element.  `tab` may point
// to a shared-memory copy of MBO_EXP2_TAB.
// TS: stride (in doubles) between consecutive table entries (32: the per-lane replicated shared-memory copy of the tensor kernel)
template <int NV, int TS = 1>
__device__ __forceinline__ void exp_nonpos_xn(const double* xin, double* out, const double* tab = MBO_EXP2_TAB) {
  const double MAGIC = 6755399441055744.0;
  double x[NV], kf[NV], r[NV], p[NV], tv[NV];
  int k[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) x[u] = fmax(xin[u], -708.0);
#pragma unroll
  for (int u = 0; u < NV; ++u) { const double t = fma(x[u], MBO_EXP_INV, MAGIC); k[u] = __double2loint(t); kf[u] = t - MAGIC; }
#pragma unroll
  for (int u = 0; u < NV; ++u) tv[u] = tab[(k[u] & 63) * TS];
#pragma unroll
  for (int u = 0; u < NV; ++u) r[u] = fma(kf[u], -MBO_EXP_HI, x[u]);
#pragma unroll
  for (int u = 0; u < NV; ++u) r[u] = fma(kf[u], -MBO_EXP_LO, r[u]);
  const double c[5] = {4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
#pragma unroll
  for (int u = 0; u < NV; ++u) p[u] = 8.333333333333333e-03;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int u = 0; u < NV; ++u) p[u] = fma(p[u], r[u], c[i]);
  }
#pragma unroll
  for (int u = 0; u < NV; ++u) p[u] *= tv[u];
#pragma unroll
  for (int u = 0; u < NV; ++u) out[u] = __hiloint2double(__double2hiint(p[u]) + ((k[u] >> 6) << 20), __double2loint(p[u]));
}

__device__ __forceinline__ void exp_nonpos_x4(const double* xin, double* out) { exp_nonpos_xn<4>(xin, out); }

// k(r2) for NV scaled squared distances at once (fp64), same values as kernel_of_r2<double>
template <int NV, int TS = 1>
__device__ __forceinline__ void kernel_of_r2_xn(int kid, double var, const double* r2, double* out,
                                                const double* tab = MBO_EXP2_TAB) {
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
  exp_nonpos_xn<NV, TS>(arg, e, tab);
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

__device__ __forceinline__ void kernel_of_r2_x4(int kid, double var, const double* r2, double* out,
                                                const double* tab = MBO_EXP2_TAB) { kernel_of_r2_xn<4>(kid, var, r2, out, tab); }

int check_candidates(const mbo_candidates* c);

}  // namespace mbo
