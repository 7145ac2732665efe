// gp.cu -- batched GP fit (K0) and GP posterior (K1) for B independent BO episodes.
//
// Replaces, per episode, what the reference does through GPy on the CPU:
//   fit        MetaBO.update_gp -> gp.set_XY            metabo/environment/metabo_gym.py:610-613
//   posterior  MetaBO.eval_gp  -> gp.predict_noiseless  metabo/environment/metabo_gym.py:732-745
// Model (GPy 1.9.8 semantics): Ky = K(X,X) + (noise + 1e-8) I,
// var clipped at 1e-15, empty GP -> mean 0 / var = kernel variance.
//
// Numerics differ from GPy's *route* on purpose: distances are formed directly (no Gram trick)
// and the variance uses v = L^-1 k* (|v|^2) instead of an explicit Ky^-1 quadratic form; on the
// ill-conditioned states (duplicate observations, cond ~4e9) this is ~5 orders of magnitude more
// accurate than the GPy route (see tests/test_gp_gpu.py and DESIGN.md "GP precision").
//
// gp_state blob per episode (fp64):  header[16] | Xs[NP][D] (X / lengthscale) | beta[NP] |
// LinvB: for every 8-row block ib, (ib+8) x 8 doubles, element (j, r) = Linv[ib + r][j]
// (zero above the diagonal / beyond n) so the posterior kernel reads 8 rows of one column with
// two broadcast LDS.128.
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace mbo {

// ------------------------------------------------------------------------------------------------
// K0: fit.  One CTA per episode.
// ------------------------------------------------------------------------------------------------
constexpr int FIT_THREADS = 128;

__global__ void __launch_bounds__(FIT_THREADS)
gp_fit_kernel(int nmax, int D, int kid, const int32_t* __restrict__ n_arr, const double* __restrict__ X,
              const double* __restrict__ Y, const double* __restrict__ lengthscale,
              const double* __restrict__ variance, const double* __restrict__ noise, int use_prior_mean,
              double* __restrict__ gp_state, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(n_arr[b], nmax);
  const int NP = gp_np(nmax);
  const int lda = NP + 1;
  double* A = sm;                    // [NP][lda]
  double* Xs = A + (size_t)NP * lda; // [NP][D]
  double* yv = Xs + NP * D;          // [NP]
  double* red = yv + NP;             // [FIT_THREADS]
  __shared__ int s_fail;
  __shared__ double s_m;

  const size_t stride = gp_state_doubles(nmax, D);
  double* st = gp_state + (size_t)b * stride;
  const double var = variance[b];

  // zero the whole blob first (padding rows/cols must be exactly 0 for the posterior kernel)
  for (int i = tid; i < (int)stride; i += FIT_THREADS) st[i] = 0.0;
  __syncthreads();
  if (tid == 0) {
    st[0] = (double)n;
    st[2] = var;
    for (int d = 0; d < D; ++d) st[4 + d] = lengthscale[b * D + d];
  }
  if (n == 0) {
    if (tid == 0) { st[1] = 0.0; st[3] = 0.0; status[b] = 0; }
    return;
  }
  // scaled inputs and targets
  for (int i = tid; i < n * D; i += FIT_THREADS) {
    int d = i % D;
    Xs[i] = X[(size_t)b * nmax * D + i] / lengthscale[b * D + d];
  }
  double part = 0.0;
  for (int i = tid; i < n; i += FIT_THREADS) { double y = Y[(size_t)b * nmax + i]; yv[i] = y; part += y; }
  red[tid] = part;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < FIT_THREADS; ++i) s += red[i];
    s_m = use_prior_mean ? s / (double)n : 0.0;   // metabo_gym.py:573 np.mean(self.Y)
  }
  __syncthreads();
  const double m = s_m;
  for (int i = tid; i < n; i += FIT_THREADS) yv[i] -= m;

  const double diag0 = var + noise[b] + 1e-8;      // k(x,x) + sigma_n^2 + GPy's 1e-8
  double jitter = 0.0;
  int tries = 0;
  while (true) {
    // ---- build the lower triangle of Ky ----
    for (int e = tid; e < n * n; e += FIT_THREADS) {
      int i = e / n, j = e - i * n;
      if (j > i) continue;
      double v;
      if (i == j) {
        v = diag0 + jitter;
      } else {
        double r2 = 0.0;
        for (int d = 0; d < D; ++d) { double t = Xs[i * D + d] - Xs[j * D + d]; r2 += t * t; }
        v = kernel_of_r2<double>(kid, var, r2);
      }
      A[i * lda + j] = v;
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    // ---- right-looking Cholesky, lower, in place ----
    for (int k = 0; k < n; ++k) {
      if (tid == 0) {
        double akk = A[k * lda + k];
        if (!(akk > 0.0)) { s_fail = 1; akk = 1.0; }
        A[k * lda + k] = sqrt(akk);
      }
      __syncthreads();
      if (s_fail) break;
      const double inv = 1.0 / A[k * lda + k];
      for (int i = k + 1 + tid; i < n; i += FIT_THREADS) A[i * lda + k] *= inv;
      __syncthreads();
      // trailing update: A[i][j] -= A[i][k] * A[j][k], k < j <= i
      const int rem = n - k - 1;
      for (int e = tid; e < rem * rem; e += FIT_THREADS) {
        int ii = e / rem, jj = e - ii * rem;
        if (jj > ii) continue;
        int i = k + 1 + ii, j = k + 1 + jj;
        A[i * lda + j] -= A[i * lda + k] * A[j * lda + k];
      }
      __syncthreads();
    }
    __syncthreads();
    if (!s_fail) break;
    // GPy jitchol: jitter = mean(diag) * 1e-6, x10 per retry, 5 tries
    ++tries;
    if (tries > 5) break;
    jitter = (tries == 1) ? diag0 * 1e-6 : jitter * 10.0;
    __syncthreads();
  }
  if (s_fail) {
    if (tid == 0) { status[b] = -1; st[3] = -1.0; st[0] = 0.0; }  // degrade to the empty-GP prediction
    return;
  }
  // ---- in-place inverse of the lower-triangular factor, row by row ----
  // Linv[i][c] = -(sum_{j=c}^{i-1} L[i][j] Linv[j][c]) / L[i][i],  Linv[i][i] = 1 / L[i][i]
  for (int i = 0; i < n; ++i) {   // n <= nmax <= FIT_THREADS: thread c owns column c
    double val = 0.0;
    if (tid <= i) {
      const double dinv = 1.0 / A[i * lda + i];
      if (tid == i) {
        val = dinv;
      } else {
        double s = 0.0;
        for (int j = tid; j < i; ++j) s += A[i * lda + j] * A[j * lda + tid];
        val = -s * dinv;
      }
    }
    __syncthreads();
    if (tid <= i) A[i * lda + tid] = val;
    __syncthreads();
  }
  // ---- beta = Linv (y - m) ----
  for (int i = tid; i < n; i += FIT_THREADS) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s += A[i * lda + j] * yv[j];
    st[gp_off_beta(NP, D) + i] = s;
  }
  for (int i = tid; i < n * D; i += FIT_THREADS) st[gp_off_xs() + i] = Xs[i];
  // ---- blocked Linv ----
  double* LB = st + gp_off_linv(NP, D);
  const int npn = gp_np(n);
  for (int ib = 0; ib < npn; ib += 8) {
    double* blk = LB + gp_linv_block_off(ib);
    const int cols = ib + 8;
    for (int e = tid; e < cols * 8; e += FIT_THREADS) {
      int j = e >> 3, r = e & 7;
      int i = ib + r;
      blk[e] = (i < n && j <= i) ? A[i * lda + j] : 0.0;
    }
  }
  if (tid == 0) { st[1] = m; st[3] = (double)tries; status[b] = tries; }
}

// ------------------------------------------------------------------------------------------------
// K1: posterior.  CTA = 128 threads, tile = 128*TC candidates of one episode.
//   phase 1: episode blob -> smem
//   phase 2: thread computes k*_j for its TC candidates into its private smem column
//   phase 3: v = Linv k* in 8-row register blocks (2 broadcast LDS.128 per column feed 8*TC FMAs),
//            mean += v_i beta_i, q += v_i^2
// ------------------------------------------------------------------------------------------------
constexpr int POST_THREADS = 128;

// DT: compile-time input dimension (0 = run time, predicated to MBO_MAX_D).  The distance loops are a third of
// the kernel's issued instructions when they are unrolled to 8 dimensions and predicated down to 3.
// KID: compile-time kernel family (the three K_of_r bodies in one loop cost register moves at every merge point).
template <typename T, int TC, int DT, int KID>
__global__ void __launch_bounds__(POST_THREADS)
gp_posterior_kernel(int nmax, int kid_rt, const double* __restrict__ gp_state, mbo_candidates cand,
                    int tiles_per_episode, int ctas_per_episode, void* __restrict__ mean_out,
                    void* __restrict__ std_out, int out_is_f32) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int kid = KID;
  (void)kid_rt;
  const int D = DT ? DT : cand.D;
  constexpr int DU = DT ? DT : MBO_MAX_D;       // unroll bound of the per-dimension loops
  // a CTA loads the episode blob once and walks tiles cta, cta + ctas_per_episode, ... of that episode
  const int b = blockIdx.x / ctas_per_episode;
  const int cta = blockIdx.x - b * ctas_per_episode;
  const int tid = threadIdx.x;
  const size_t stride = gp_state_doubles(nmax, D);
  const double* st = gp_state + (size_t)b * stride;
  const int NPmax = gp_np(nmax);
  const int n = (int)st[0];
  const int NP = gp_np(n);
  const double m = st[1];
  const double var = st[2];

  T* sXs = (T*)smraw;                         // [NP][D]
  T* sBeta = sXs + NP * D;                    // [NP]
  T* sL = sBeta + NP;                         // blocked Linv, 16B aligned (NP*D + NP is even... padded below)
  {
    size_t off = (size_t)(NP * D + NP);
    off = (off + 3) & ~(size_t)3;
    sL = sXs + off;
  }
  const int lcount = gp_linv_doubles(NP);
  T* sK = sL + lcount;                        // [NP][POST_THREADS*TC]
  __shared__ double s_exp2[64];               // shared-memory copy of the exp table (lane-varying index)
  if (tid < 64) s_exp2[tid] = MBO_EXP2_TAB[tid];
  for (int i = tid; i < NP * D; i += POST_THREADS) sXs[i] = (T)st[gp_off_xs() + i];
  for (int i = tid; i < NP; i += POST_THREADS) sBeta[i] = (T)st[gp_off_beta(NPmax, D) + i];
  for (int i = tid; i < lcount; i += POST_THREADS) sL[i] = (T)st[gp_off_linv(NPmax, D) + i];
  __syncthreads();

  for (int tile = cta; tile < tiles_per_episode; tile += ctas_per_episode) {
  const int m0 = tile * POST_THREADS * TC;
  T xs[TC][DU];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    int mi = m0 + tid * TC + c;
    double x[MBO_MAX_D];
#pragma unroll
    for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
    if (mi < cand.M) load_candidate(cand, b, mi, x);
    // same operation as the fit (X / lengthscale) so a candidate equal to an observation has r == 0
#pragma unroll
    for (int d = 0; d < DU; ++d) xs[c][d] = d < D ? (T)(x[d] / st[4 + d]) : T(0);
  }
  constexpr int KS = POST_THREADS * TC;
  // phase 2
  if (sizeof(T) == 8) {
    // fp64: four kernel values at a time with interleaved exp chains (see exp_nonpos_x4)
#pragma unroll
    for (int c = 0; c < TC; ++c) {
      for (int j0 = 0; j0 < n; j0 += 4) {
        double r2v[4], kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = min(j0 + u, n - 1);
          double r2 = 0.0;
#pragma unroll
          for (int d = 0; d < DU; ++d)
            if (DT || d < D) { const double t = (double)xs[c][d] - (double)sXs[jj * D + d]; r2 = fma(t, t, r2); }
          r2v[u] = r2;
        }
        kernel_of_r2_x4(kid, var, r2v, kv, s_exp2);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < n) sK[(j0 + u) * KS + tid * TC + c] = (T)kv[u];
      }
    }
  } else {
    for (int j = 0; j < n; ++j) {
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        T r2 = T(0);
#pragma unroll
        for (int d = 0; d < DU; ++d)
          if (DT || d < D) { T t = xs[c][d] - sXs[j * D + d]; r2 += t * t; }
        sK[j * KS + tid * TC + c] = kernel_of_r2<T>(kid, (T)var, r2);
      }
    }
  }
  for (int j = n; j < NP; ++j) {
#pragma unroll
    for (int c = 0; c < TC; ++c) sK[j * KS + tid * TC + c] = T(0);
  }
  // phase 3 (each thread only reads the smem column it wrote: no barrier needed)
  T q[TC], mu[TC];
#pragma unroll
  for (int c = 0; c < TC; ++c) { q[c] = T(0); mu[c] = T(0); }
  for (int ib = 0; ib < NP; ib += 8) {
    T acc[8][TC];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < TC; ++c) acc[r][c] = T(0);
    const T* blk = sL + gp_linv_block_off(ib);
    const int cols = ib + 8;
#pragma unroll 4
    for (int j = 0; j < cols; ++j) {
      T l[8];
      if (sizeof(T) == 8) {
        const double2* p = (const double2*)(blk + j * 8);
        double2 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
        l[0] = (T)a0.x; l[1] = (T)a0.y; l[2] = (T)a1.x; l[3] = (T)a1.y;
        l[4] = (T)a2.x; l[5] = (T)a2.y; l[6] = (T)a3.x; l[7] = (T)a3.y;
      } else {
        const float4* p = (const float4*)(blk + j * 8);
        float4 a0 = p[0], a1 = p[1];
        l[0] = (T)a0.x; l[1] = (T)a0.y; l[2] = (T)a0.z; l[3] = (T)a0.w;
        l[4] = (T)a1.x; l[5] = (T)a1.y; l[6] = (T)a1.z; l[7] = (T)a1.w;
      }
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        T kv = sK[j * KS + tid * TC + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][c] = fma(l[r], kv, acc[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      T be = sBeta[ib + r];
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        q[c] = fma(acc[r][c], acc[r][c], q[c]);
        mu[c] = fma(acc[r][c], be, mu[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    int mi = m0 + tid * TC + c;
    if (mi >= cand.M) continue;
    double mean_v, var_v;
    if (n == 0) {
      mean_v = 0.0;               // metabo_gym.py:737: zero, not the prior mean
      var_v = var;
    } else {
      mean_v = (double)mu[c] + m;
      var_v = var - (double)q[c];
      var_v = var_v < 1e-15 ? 1e-15 : var_v;    // GPy clip
    }
    double sd = sqrt(var_v);
    size_t o = (size_t)b * cand.M + mi;
    if (out_is_f32) { ((float*)mean_out)[o] = (float)mean_v; ((float*)std_out)[o] = (float)sd; }
    else { ((double*)mean_out)[o] = mean_v; ((double*)std_out)[o] = sd; }
  }
  }   // tile loop (each thread only touches its own shared-memory column: no barrier between tiles)
}

// ------------------------------------------------------------------------------------------------
// K1m: fp64 posterior with the blocked triangular mat-vec v = Linv k* on fp64 TENSOR instructions
// (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).  A warp works on 8 candidates per pass:
//   A fragment (8 x 4, row)  = Linv[8 rb + g][4 kb + t]  straight from the blob's 8-row blocks (conflict-free LDS.64),
//   B fragment (4 x 8, col)  = k*[4 kb + t] of candidate g -- the lane that needs the value is the lane that computes
//                              it (obs j = 4 kb + t, candidate g; g = lane >> 2, t = lane & 3), so k* is a per-lane
//                              spill area [NKB][32] in shared memory: no cross-lane traffic, no barrier,
//   C fragment (8 x 8)       = v[8 rb + g] of candidates 2t, 2t + 1; mean += v beta, q += v^2 per lane, summed over the
//                              8 row lanes (g) with three shuffle steps.
// Measured on B200 (tools/micro/dmma_bw.cu): DMMA and DFMA share the fp64 pipe (36.7 vs ~37 TFLOP/s, no gain from
// mixing) -- the win is issue slots and latency tolerance: 20 DMMA + 40 LDS.64 per pass replace 640 DFMA + 240 LDS, the
// per-thread k* column ([NP][128 TC] doubles, 65-106 KB per CTA) shrinks to 8-27 KB, so 8 CTAs (n = 30) or 16 warps
// (n = 100) are resident instead of 3 CTAs / 4 warps.  Same values as gp_posterior_kernel<double> up to the order of
// the fp64 additions (~1e-15 relative; tests/test_gp_gpu.py).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NRB_T > 0: n <= 8 NRB_T observations, every loop over row / k blocks unrolled and k* kept in REGISTERS (2 NRB_T doubles
// per lane); NRB_T == 0: any n, dynamic loops, k* in a per-lane shared-memory spill area.
// Shared-memory traffic is what bounded both this kernel's first version and the CUDA-core kernel (LSU data-pipe
// wavefronts at 80 % of peak in ncu): so (i) Linv is re-laid-out into mma fragment order while the blob is copied in
// (tile (rb, kb) = 32 consecutive doubles in lane order: every A load is a conflict-free LDS.64), (ii) the 64-entry
// exp table is replicated per lane ([64][32] doubles: a lane-varying index costs the minimum two wavefronts instead
// of ~8), (iii) candidates are prepared once per lane (load + X / lengthscale) and handed to the pass that needs
// them by shuffles, (iv) the per-candidate epilogue (clip, sqrt, store) runs once per lane after the four passes.
template <int DT, int KID, int NRB_T, int TSV = 32>
__global__ void __launch_bounds__(NRB_T ? 128 : 256, NRB_T ? (NRB_T <= 4 ? 8 : 4) : 1)      // NRB_T <= 4: 64 registers (no spills) -> 8 CTAs / SM; 6, 8 (n <= 48, 64): 128
gp_posterior_mma_kernel(int nmax, const double* __restrict__ gp_state, mbo_candidates cand, int tiles_per_episode,
                        int ctas_per_episode, void* __restrict__ mean_out, void* __restrict__ std_out, int out_is_f32) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int D = DT ? DT : cand.D;
  constexpr int DU = DT ? DT : MBO_MAX_D;
  const int b = blockIdx.x / ctas_per_episode;
  const int cta = blockIdx.x - b * ctas_per_episode;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const size_t stride = gp_state_doubles(nmax, D);
  const double* st = gp_state + (size_t)b * stride;
  const int NPmax = gp_np(nmax);
  const int n = (int)st[0];
  // NRB_T > 0: the unrolled loops cover 8 NRB_T rows whatever this episode's n is; the blob is zero beyond n (the fit
  // clears it), so the extra rows / k blocks of a shorter episode contribute exact zeros
  const int NP = NRB_T ? 8 * NRB_T : gp_np(n);
  const int NKB = NP >> 2, NRB = NRB_T ? NRB_T : (NP >> 3);
  const double m = st[1];
  const double var = st[2];

  // TS copies of the exp table: 32 = one per lane (16 KB), 8 = 4 KB (a CTA then needs 12 KB at n = 30 and fits beside the
  // resident tensor-core MLP CTA of another stream); the generic kernel (large blobs) keeps a single 512-byte copy
  constexpr int TS = NRB_T ? TSV : 1;
  double* sTab = (double*)smraw;              // [64][TS] exp table (TS = 32: one copy per lane)
  double* sXs = sTab + 64 * TS;               // [NP][D]
  double* sBeta = sXs + NP * D;               // [NP]
  double* sA = sBeta + NP;                    // Linv in mma A-fragment order
  const int lcount = gp_linv_doubles(NP);
  double* sOut = sA + lcount + (size_t)warp * 64;                              // [2][32] q / mu of the warp's 32 candidates
  double* sK = sA + lcount + (size_t)nwarps * 64 + (size_t)warp * NKB * 32;    // NRB_T == 0: this warp's k* spill area [NKB][32]
  for (int i = tid; i < 64 * TS; i += blockDim.x) sTab[i] = MBO_EXP2_TAB[i / TS];
  for (int i = tid; i < NP * D; i += blockDim.x) sXs[i] = st[gp_off_xs() + i];
  for (int i = tid; i < NP; i += blockDim.x) sBeta[i] = st[gp_off_beta(NPmax, D) + i];
  {
    const double* lb = st + gp_off_linv(NPmax, D);
    for (int rb = 0; rb < (NP >> 3); ++rb) {
      const int off = gp_linv_block_off(8 * rb), cnt = (8 * rb + 8) * 8;
      for (int e = tid; e < cnt; e += blockDim.x) {          // blob: (j, r) -> j * 8 + r;  fragment order: tile j / 4, lane 4 r + j % 4
        const int j = e >> 3, r = e & 7;
        sA[off + (j >> 2) * 32 + r * 4 + (j & 3)] = lb[off + e];
      }
    }
  }
  double ls[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) ls[d] = d < D ? st[4 + d] : 1.0;
  __syncthreads();
  const double* tab = sTab + (lane & (TS - 1));    // entry j of this lane's copy: tab[j * TS]

  const int TILE = nwarps * 32;
  for (int tile = cta; tile < tiles_per_episode; tile += ctas_per_episode) {
    const int w0 = tile * TILE + warp * 32;
    // this lane's own candidate, scaled like the fit scales the observations (a candidate equal to an observation has r == 0)
    double own[DU];
    {
      double x[MBO_MAX_D];
#pragma unroll
      for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
      if (w0 + lane < cand.M) load_candidate_compact(cand, b, w0 + lane, x);
#pragma unroll
      for (int d = 0; d < DU; ++d) own[d] = d < D ? x[d] / ls[d] : 0.0;
    }
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      double xs[DU];
#pragma unroll
      for (int d = 0; d < DU; ++d) xs[d] = __shfl_sync(0xffffffffu, own[d], pass * 8 + g);
      // ---- k*: lane (g, t) evaluates observation j = 4 kb + t for candidate g ----
      double K[NRB_T ? 2 * NRB_T : 1];
      constexpr int NKB_U = NRB_T ? 2 * NRB_T : 0;
      if (NRB_T) {
        constexpr int GX = (NKB_U % 4 == 0) ? 4 : 2;           // kernel values per step (interleaved exp chains)
#pragma unroll
        for (int kb0 = 0; kb0 < NKB_U; kb0 += GX) {
          double r2v[GX], kv[GX];
#pragma unroll
          for (int u = 0; u < GX; ++u) {
            const int jj = max(min(4 * (kb0 + u) + t, n - 1), 0);
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < DU; ++d)
              if (DT || d < D) { const double dd = xs[d] - sXs[jj * D + d]; r2 = fma(dd, dd, r2); }
            r2v[u] = r2;
          }
          kernel_of_r2_xn<GX, TS>(KID, var, r2v, kv, tab);
#pragma unroll
          for (int u = 0; u < GX; ++u) K[kb0 + u] = (4 * (kb0 + u) + t < n) ? kv[u] : 0.0;
        }
      } else {
        for (int kb0 = 0; kb0 < NKB; kb0 += 2) {
          double r2v[2], kv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int jj = max(min(4 * (kb0 + u) + t, n - 1), 0);
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < DU; ++d)
              if (DT || d < D) { const double dd = xs[d] - sXs[jj * D + d]; r2 = fma(dd, dd, r2); }
            r2v[u] = r2;
          }
          kernel_of_r2_xn<2, TS>(KID, var, r2v, kv, tab);
#pragma unroll
          for (int u = 0; u < 2; ++u) sK[(kb0 + u) * 32 + lane] = (4 * (kb0 + u) + t < n) ? kv[u] : 0.0;     // NKB is even
        }
      }
      // ---- v = Linv k* by row blocks (DMMA), mean / q accumulation ----
      double q0 = 0.0, q1 = 0.0, mu0 = 0.0, mu1 = 0.0;
      if (NRB_T) {
#pragma unroll
        for (int rb = 0; rb < (NRB_T ? NRB_T : 1); ++rb) {
          double c0 = 0.0, c1 = 0.0;
          const double* blk = sA + 32 * rb * (rb + 1) + lane;      // gp_linv_block_off(8 rb)
#pragma unroll
          for (int kb = 0; kb < 2 * rb + 2; ++kb) dmma_884(c0, c1, blk[kb * 32], K[kb]);
          const double be = sBeta[8 * rb + g];
          q0 = fma(c0, c0, q0); q1 = fma(c1, c1, q1);
          mu0 = fma(c0, be, mu0); mu1 = fma(c1, be, mu1);
        }
      } else {
        for (int rb = 0; rb < NRB; ++rb) {
          double c0 = 0.0, c1 = 0.0;
          const double* blk = sA + gp_linv_block_off(8 * rb) + lane;
          const int nk = 2 * rb + 2;
#pragma unroll 2
          for (int kb = 0; kb < nk; ++kb) dmma_884(c0, c1, blk[kb * 32], sK[kb * 32 + lane]);
          const double be = sBeta[8 * rb + g];
          q0 = fma(c0, c0, q0); q1 = fma(c1, c1, q1);
          mu0 = fma(c0, be, mu0); mu1 = fma(c1, be, mu1);
        }
      }
      // ---- sum over the 8 row lanes (g): exchange-and-halve butterfly, 4 double shuffles instead of 12 ----
      const bool hi4 = lane & 16, hi3 = lane & 8;
      const double r0 = __shfl_xor_sync(0xffffffffu, hi4 ? q0 : mu0, 16), r1 = __shfl_xor_sync(0xffffffffu, hi4 ? q1 : mu1, 16);
      const double a0 = (hi4 ? mu0 : q0) + r0, a1 = (hi4 ? mu1 : q1) + r1;       // bit 4 set: the mu pair, else the q pair
      const double r2x = __shfl_xor_sync(0xffffffffu, hi3 ? a0 : a1, 8);
      double a = (hi3 ? a1 : a0) + r2x;                                            // bit 3 set: candidate 2t + 1, else 2t
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if (!(lane & 4)) sOut[(hi4 ? 32 : 0) + pass * 8 + 2 * t + (hi3 ? 1 : 0)] = a;
    }
    __syncwarp();
    const int mo = w0 + lane;
    if (mo < cand.M) {
      double mean_v, var_v;
      if (n == 0) { mean_v = 0.0; var_v = var; }               // metabo_gym.py:737: zero, not the prior mean
      else {
        mean_v = sOut[32 + lane] + m;
        var_v = var - sOut[lane];
        var_v = var_v < 1e-15 ? 1e-15 : var_v;                // GPy clip
      }
      const double sd = sqrt(var_v);
      const size_t o = (size_t)b * cand.M + mo;
      if (out_is_f32) { ((float*)mean_out)[o] = (float)mean_v; ((float*)std_out)[o] = (float)sd; }
      else { ((double*)mean_out)[o] = mean_v; ((double*)std_out)[o] = sd; }
    }
    __syncwarp();
  }
}

static size_t posterior_mma_smem_bytes(int ns, int D, int warps, bool k_in_smem, int ts) {
  const int NP = gp_np(ns);
  return sizeof(double) * ((size_t)64 * (k_in_smem ? 1 : ts) + (size_t)NP * D + NP + gp_linv_doubles(NP) + (size_t)warps * 64 +
                           (k_in_smem ? (size_t)warps * (NP >> 2) * 32 : 0));
}

template <int DT, int KID>
static int launch_posterior_mma_dk(int B, int nmax, int ns, const double* gp_state, const mbo_candidates& cand, void* mean, void* std_,
                                   int out_is_f32, cudaStream_t s) {
  const int NP = gp_np(ns);
  const int nrb = NP >> 3;
  // small blobs: 4-warp CTAs (many resident CTAs); large blobs (n ~ 100: 50 KB): 8 warps share one copy
  int warps = sizeof(double) * ((size_t)NP * cand.D + NP + gp_linv_doubles(NP)) > 20 * 1024 ? 8 : 4;
  // k* in registers: n <= 32 exactly (NRB_T = 1..4), n <= 48 / 64 on the NRB_T = 6 / 8 instantiations (a shorter episode's
  // extra row / k blocks are zeros) as long as the blob of this call has that many rows
  const int npmax_rb = gp_np(nmax) >> 3;
  const char* ts_env = getenv("MBO_GP_TS");
  const int ts = ts_env && atoi(ts_env) == 32 ? 32 : 8;      // 8 measured as fast as 32 (n = 30: 1.71 vs 1.73 ms) at a quarter of the table
  int nrb_t = nrb <= 4 ? nrb : (nrb <= 6 && npmax_rb >= 6 ? 6 : (nrb <= 8 && npmax_rb >= 8 ? 8 : 0));
  if (nrb_t > 4 && (ts == 32 || getenv("MBO_GP_NO_WIDE_REG"))) nrb_t = 0;      // (A/B: the shared-memory k* kernel)
  const bool reg_k = nrb_t >= 1;
  if (reg_k) warps = 4;                                        // the register kernels are 4-warp CTAs
  const size_t smem = posterior_mma_smem_bytes(reg_k ? 8 * nrb_t : ns, cand.D, warps, !reg_k, ts);
  auto kern = gp_posterior_mma_kernel<DT, KID, 0>;
  if (ts == 32) {
    if (nrb == 1) kern = gp_posterior_mma_kernel<DT, KID, 1, 32>;
    else if (nrb == 2) kern = gp_posterior_mma_kernel<DT, KID, 2, 32>;
    else if (nrb == 3) kern = gp_posterior_mma_kernel<DT, KID, 3, 32>;
    else if (nrb == 4) kern = gp_posterior_mma_kernel<DT, KID, 4, 32>;
  } else {
    if (nrb == 1) kern = gp_posterior_mma_kernel<DT, KID, 1, 8>;
    else if (nrb == 2) kern = gp_posterior_mma_kernel<DT, KID, 2, 8>;
    else if (nrb == 3) kern = gp_posterior_mma_kernel<DT, KID, 3, 8>;
    else if (nrb == 4) kern = gp_posterior_mma_kernel<DT, KID, 4, 8>;
    else if (nrb_t == 6) kern = gp_posterior_mma_kernel<DT, KID, 6, 8>;
    else if (nrb_t == 8) kern = gp_posterior_mma_kernel<DT, KID, 8, 8>;
  }
  MBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // same shared-memory carve-out as the tensor-core MLP kernel (maximum): an SM has ONE L1 / shared split at a time, so
  // a kernel that prefers another split cannot start on an SM where an MLP CTA of another stream is resident
  MBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  const int TILE = warps * 32;
  const int tiles = (cand.M + TILE - 1) / TILE;
  int ctas = tiles;
  if ((long long)B * tiles > 4096) { ctas = (tiles + 3) / 4; if (ctas < 1) ctas = 1; }
  kern<<<B * ctas, warps * 32, smem, s>>>(nmax, gp_state, cand, tiles, ctas, mean, std_, out_is_f32);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

template <int KID>
static int launch_posterior_mma_k(int B, int nmax, int ns, const double* gp_state, const mbo_candidates& cand, void* mean, void* std_,
                                  int out_is_f32, cudaStream_t s) {
  // the dimensions of the shipped experiments (Branin 2, Hartmann-3 / GP samples 3) and of the sweep (4, 6) get their
  // own instantiation; everything else runs the generic one (distance loops predicated to MBO_MAX_D)
  if (cand.D == 2) return launch_posterior_mma_dk<2, KID>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  if (cand.D == 3) return launch_posterior_mma_dk<3, KID>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  if (cand.D == 4) return launch_posterior_mma_dk<4, KID>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  if (cand.D == 6) return launch_posterior_mma_dk<6, KID>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  return launch_posterior_mma_dk<0, KID>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
}

static int launch_posterior_mma(int B, int nmax, int ns, int kid, const double* gp_state, const mbo_candidates& cand,
                                void* mean, void* std_, int out_is_f32, cudaStream_t s) {
  if (kid == MBO_KERNEL_RBF) return launch_posterior_mma_k<MBO_KERNEL_RBF>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  if (kid == MBO_KERNEL_MATERN32) return launch_posterior_mma_k<MBO_KERNEL_MATERN32>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
  return launch_posterior_mma_k<MBO_KERNEL_MATERN52>(B, nmax, ns, gp_state, cand, mean, std_, out_is_f32, s);
}

template <typename T, int TC>
static size_t posterior_smem_bytes(int nmax, int D) {
  int NP = gp_np(nmax);
  size_t off = (size_t)(NP * D + NP);
  off = (off + 3) & ~(size_t)3;
  return sizeof(T) * (off + gp_linv_doubles(NP) + (size_t)NP * POST_THREADS * TC);
}

template <typename T, int TC>
static int launch_posterior(int B, int nmax, int ns, int kid, const double* gp_state, const mbo_candidates& cand,
                            void* mean, void* std_, int out_is_f32, cudaStream_t s) {
  size_t smem = posterior_smem_bytes<T, TC>(ns, cand.D);
  // the dimensions of the shipped experiments get their own instantiation (fp64 only), everything else the generic one
  auto pick = [&](auto dt) {
    constexpr int DTV = decltype(dt)::value;
    return kid == MBO_KERNEL_RBF ? gp_posterior_kernel<T, TC, DTV, MBO_KERNEL_RBF>
           : kid == MBO_KERNEL_MATERN32 ? gp_posterior_kernel<T, TC, DTV, MBO_KERNEL_MATERN32>
                                        : gp_posterior_kernel<T, TC, DTV, MBO_KERNEL_MATERN52>;
  };
  auto kern = pick(std::integral_constant<int, 0>{});
  if (sizeof(T) == 8) {
    if (cand.D == 2) kern = pick(std::integral_constant<int, 2>{});
    else if (cand.D == 3) kern = pick(std::integral_constant<int, 3>{});
    else if (cand.D == 4) kern = pick(std::integral_constant<int, 4>{});
    else if (cand.D == 6) kern = pick(std::integral_constant<int, 6>{});
  }
  MBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int tiles = (cand.M + POST_THREADS * TC - 1) / (POST_THREADS * TC);
  // enough CTAs to fill the machine several times over, few enough that the blob load is amortised over ~4 tiles
  int ctas = tiles;
  if ((long long)B * tiles > 4096) { ctas = (tiles + 3) / 4; if (ctas < 1) ctas = 1; }
  kern<<<B * ctas, POST_THREADS, smem, s>>>(nmax, kid, gp_state, cand, tiles, ctas, mean, std_, out_is_f32);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

}  // namespace mbo

using namespace mbo;

extern "C" size_t mbo_gp_state_doubles(int nmax, int D) { return (size_t)gp_state_doubles(nmax, D); }

extern "C" int mbo_gp_fit(int B, int nmax, int D, int kernel_id, const int32_t* n, const double* X,
                          const double* Y, const double* lengthscale, const double* variance,
                          const double* noise, int use_prior_mean, double* gp_state, int32_t* status,
                          void* stream) {
  MBO_CHECK_ARG(B > 0 && nmax > 0 && nmax <= 128, "mbo_gp_fit: need B > 0 and 0 < nmax <= 128 (got %d, %d)", B, nmax);
  MBO_CHECK_ARG(D > 0 && D <= MBO_MAX_D, "mbo_gp_fit: D must be in 1..%d", MBO_MAX_D);
  MBO_CHECK_ARG(kernel_id >= 0 && kernel_id <= 2, "mbo_gp_fit: unknown kernel id %d", kernel_id);
  MBO_CHECK_ARG(n && X && Y && lengthscale && variance && noise && gp_state && status, "mbo_gp_fit: null pointer");
  int NP = gp_np(nmax);
  size_t smem = sizeof(double) * ((size_t)NP * (NP + 1) + (size_t)NP * D + NP + FIT_THREADS);
  MBO_CUDA(cudaFuncSetAttribute(gp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gp_fit_kernel<<<B, FIT_THREADS, smem, (cudaStream_t)stream>>>(nmax, D, kernel_id, n, X, Y, lengthscale, variance,
                                                               noise, use_prior_mean, gp_state, status);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_gp_posterior(int B, int nmax, int n_active_max, int kernel_id, int precision,
                                const double* gp_state, const mbo_candidates* cand, void* mean, void* std_,
                                int out_is_f32, void* stream) {
  MBO_CHECK_ARG(B > 0 && nmax > 0 && nmax <= 128, "mbo_gp_posterior: need B > 0 and 0 < nmax <= 128");
  MBO_CHECK_ARG(kernel_id >= 0 && kernel_id <= 2, "mbo_gp_posterior: unknown kernel id %d", kernel_id);
  MBO_CHECK_ARG(gp_state && cand && mean && std_, "mbo_gp_posterior: null pointer");
  MBO_CHECK_ARG(n_active_max >= 0 && n_active_max <= nmax, "mbo_gp_posterior: n_active_max must be in 0..nmax");
  const int ns = n_active_max > 0 ? n_active_max : 1;   // sizes shared memory only
  int rc = check_candidates(cand);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (precision == MBO_GP_FP64) {
    const char* simt_env = getenv("MBO_GP_SIMT");            // 1: the CUDA-core (DFMA) kernel, kept as the A/B and parity reference
    const int simt = simt_env ? atoi(simt_env) : 0;
    if (!simt) return launch_posterior_mma(B, nmax, ns, kernel_id, gp_state, *cand, mean, std_, out_is_f32, s);
    static const int force_tc1 = getenv("MBO_GP_TC1") ? 1 : 0;     // tuning switch
    if (!force_tc1 && posterior_smem_bytes<double, 2>(ns, cand->D) <= 100 * 1024)
      return launch_posterior<double, 2>(B, nmax, ns, kernel_id, gp_state, *cand, mean, std_, out_is_f32, s);
    return launch_posterior<double, 1>(B, nmax, ns, kernel_id, gp_state, *cand, mean, std_, out_is_f32, s);
  } else if (precision == MBO_GP_FP32) {
    if (posterior_smem_bytes<float, 2>(ns, cand->D) <= 100 * 1024)
      return launch_posterior<float, 2>(B, nmax, ns, kernel_id, gp_state, *cand, mean, std_, out_is_f32, s);
    return launch_posterior<float, 1>(B, nmax, ns, kernel_id, gp_state, *cand, mean, std_, out_is_f32, s);
  }
  set_error("mbo_gp_posterior: unknown precision %d", precision);
  return MBO_ERR_ARG;
}
