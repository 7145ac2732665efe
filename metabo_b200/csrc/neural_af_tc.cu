// neural_af_tc.cu -- fused NeuralAF evaluation on tcgen05 tensor cores (sm_100a).
//
// Replaces get_state feature packing (metabo/environment/metabo_gym.py:624-677) followed by
// NeuralAF.forward -> MLP.forward (metabo/policies/policies.py:104-125, mlp.py:52-61) for the
// shipped F -> H -> H -> H -> H -> 1 ReLU policies (H = 200): features and hidden activations
// never touch HBM; one CTA owns a tile of 128 candidates of one episode.
//
// Data flow per tile (TMEM columns: R0 = [0,208), R1 = [208,416), S = [416,448)):
//   features (fp32)  --split hi/lo tf32-->  S                         (feature warps, tcgen05.st)
//   L0: D(R0)  = S * W0^T     3xTF32 (hi*hi + lo*hi + hi*lo), bias folded in as a ones column
//   E0: R0 <- act(relu(D))    in place, converted to the MMA operand format (fp16 mode: by the feature warps)
//   L1: D(R1)  = A(R0) * W1^T ; E1 in place on R1 ; L2: D(R0) = A(R1) * W2^T ; E2 ; L3: D(R1) = A(R0) * W3^T
//   E3: logit = sum_j relu(D_j + b3_j) * w4_j + b4                    (CUDA cores, fp32)
// The A operand of every MMA comes from TMEM (tcgen05.mma "TS" form), the B operand (weights, K-major, no-swizzle
// canonical layout) from shared memory: tf32 / bf16x3 stream it from L2 through a ring filled with cp.async.bulk (TMA
// engine, 1-D) and released by tcgen05.commit; the fp16 mode keeps most of it resident (see TcSmem).  The epilogue of
// layer l hands the next layer its K range chunk by chunk (mbarriers), so the MMAs of layer l+1 start while the
// epilogue of layer l is still converting.
//
// Arithmetic modes (hidden layers): MBO_MLP_TF32   kind::tf32, operands rounded to tf32 (rna);
//                                   MBO_MLP_BF16X3 kind::f16 bf16, a = a_hi + a_lo, w = w_hi + w_lo,
//                                                  a*w ~ a_hi*w_hi + a_lo*w_hi + a_hi*w_lo;
//                                   MBO_MLP_FP16   kind::f16 fp16: same 11-bit significand as tf32 at twice the tensor
//                                                  rate.  Biases ride in the MMA (a ones column that the weights
//                                                  propagate from layer to layer), so an epilogue element costs half
//                                                  an instruction (cvt.rn.satfinite.relu.f16x2.f32); layers with an
//                                                  unusual weight scale are renormalised by a power of two at pack
//                                                  time (exact: ReLU is positively homogeneous); an activation that
//                                                  reaches the fp16 range limit raises bit 1 of the error flag.
// Warp roles (448 threads): warps 0-7 epilogue (TMEM lanes 32*(w&3).., chunk parity w>>2), warp 8 TMA
// producer + TMEM allocation, warp 9 MMA issuer, warps 10-13 feature warps: they gather mean/std/candidates of the
// tiles ahead from global memory and write the layer-0 operand into TMEM, so that this latency (~4 000 cycles per
// tile when the epilogue warps did it) is off the tensor pipe's critical path.  Every mbarrier wait is bounded
// (globaltimer) and raises a device-side error flag instead of hanging.
#include "common.cuh"
#include "policy.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "tmem_ldst.cuh"

#ifdef MBO_TC_TRACE
// debug build only (make EXTRA=-DMBO_TC_TRACE): clock64 stamps of pipeline events of CTA 0, tiles T0..T0+3
__device__ long long g_tc_trace[8192];
#ifndef MBO_TC_TRACE_CTA_B
#define MBO_TC_TRACE_CTA_B 82
#endif
#define TC_STAT_BEGIN() const long long _ts = clock64()
#define TC_STAT_END(acc) (acc) += clock64() - _ts
#ifndef MBO_TC_TRACE_T0
#define MBO_TC_TRACE_T0 200
#endif
#define TC_TRACE_T(tt, slot) do { if ((blockIdx.x == 0 || blockIdx.x == MBO_TC_TRACE_CTA_B) && (tt) >= MBO_TC_TRACE_T0 && (tt) < MBO_TC_TRACE_T0 + 4 && lane == 0) g_tc_trace[(blockIdx.x ? 2048 : 0) + ((tt) - MBO_TC_TRACE_T0) * 512 + (slot)] = clock64(); } while (0)
#define TC_TRACE(slot) TC_TRACE_T(t, slot)
#else
#define TC_TRACE(slot) do { } while (0)
#define TC_TRACE_T(tt, slot) do { } while (0)
#define TC_STAT_BEGIN() do { } while (0)
#define TC_STAT_END(acc) do { } while (0)
#endif

namespace mbo {

constexpr int TC_N = 208;            // padded hidden width (UMMA N, multiple of 16)
constexpr int TC_M = 128;            // candidates per tile (UMMA M)
constexpr int TC_THREADS = 448;        // 8 epilogue warps + TMA warp + MMA warp + 4 feature warps
constexpr int WARP_TMA = 8, WARP_MMA = 9, WARP_GP0 = 10;   // warps 10-13: feature warps, or the fused GP posterior (FUSE)
constexpr int TC_THREADS_FUSED = 448;
constexpr int GP_NP = 32;                 // fused GP: n <= 32 observations
constexpr int GPB_XS = 16, GPB_BETA = GPB_XS + GP_NP * MBO_MAX_D, GPB_LINV = GPB_BETA + GP_NP;
constexpr int GPB_DOUBLES = GPB_LINV + 640;   // header | Xs[32][8] | beta[32] | LinvB(4 blocks)
constexpr int TC_K0 = 16;            // layer-0 K capacity (features + ones column <= 16); 8 are used when F <= 7
constexpr int CHUNK_ROW_BYTES = TC_N * 16;   // one 16-byte K chunk for all 208 rows = 3328 B
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_R0 = 0, COL_R1 = 208, COL_S = 416;

template <int MODE> struct ModeCfg;
template <> struct ModeCfg<MBO_MLP_TF32> {
  static constexpr int K = 200;                // 25 slices of 8
  static constexpr int SLICE_K = 8;
  static constexpr int SLICES = 25;
  static constexpr int SLICES_PER_STAGE = 5;
  static constexpr int STAGES_PER_LAYER = 5;
  static constexpr int STAGE_BYTES = SLICES_PER_STAGE * 2 * CHUNK_ROW_BYTES;   // 33 280
  static constexpr int NSTAGE = 6;
  static constexpr int NSTAGE_FUSED = 4;       // fused GP needs 42 KB of shared memory for its scratch
  static constexpr int CHUNK_COLS = 40;        // accumulator columns converted per hand-off
  static constexpr int NCHUNK = 5;
};
template <> struct ModeCfg<MBO_MLP_BF16X3> {
  static constexpr int K = 208;                // 13 slices of 16
  static constexpr int SLICE_K = 16;
  static constexpr int SLICES = 13;
  static constexpr int SLICES_PER_STAGE = 1;
  static constexpr int STAGES_PER_LAYER = 13;
  static constexpr int STAGE_BYTES = 2 * 2 * CHUNK_ROW_BYTES;                  // hi + lo: 13 312
  static constexpr int NSTAGE = 14;
  static constexpr int NSTAGE_FUSED = 11;
  static constexpr int CHUNK_COLS = 32;
  static constexpr int NCHUNK = 7;             // 6 x 32 + 1 x 16
};

template <> struct ModeCfg<MBO_MLP_FP16> {
  static constexpr int K = 208;                // 13 slices of 16; K columns H, H+1 carry the bias (hi, lo)
  static constexpr int SLICE_K = 16;
  static constexpr int SLICES = 13;
  static constexpr int SLICES_PER_STAGE = 2;
  static constexpr int STAGES_PER_LAYER = 7;   // the last stage holds one slice
  static constexpr int STAGE_BYTES = 2 * 2 * CHUNK_ROW_BYTES;                  // 13 312
  static constexpr int NSTAGE = 14;            // two layers resident
  static constexpr int NSTAGE_FUSED = 11;
  static constexpr int CHUNK_COLS = 32;
  static constexpr int NCHUNK = 7;             // 6 x 32 + 1 x 16
};

constexpr int W0_BYTES = 2 * (TC_K0 / 4) * CHUNK_ROW_BYTES;   // hi + lo: 13 312
constexpr int TAIL_FLOATS = 5 * TC_N + 4;                    // zeros(b0 is folded into W0) b1 b2 b3 | w4 | b4

// packed-weight buffer layout (bytes):
//   [W0][tf32 hidden x3][bf16 hidden x3][tail floats] | fp16 mode: [W0h][fp16 hidden x3][tail floats][scale exponents]
struct PackLayout {
  size_t off_w0, off_tf32, off_bf16, off_tail, off_w0h, off_f16, off_tailh, off_scale, off_ssq, total;
};
__host__ __device__ inline PackLayout pack_layout() {
  PackLayout p;
  p.off_w0 = 0;
  p.off_tf32 = W0_BYTES;
  p.off_bf16 = p.off_tf32 + (size_t)3 * ModeCfg<MBO_MLP_TF32>::STAGES_PER_LAYER * ModeCfg<MBO_MLP_TF32>::STAGE_BYTES;
  p.off_tail = p.off_bf16 + (size_t)3 * ModeCfg<MBO_MLP_BF16X3>::STAGES_PER_LAYER * ModeCfg<MBO_MLP_BF16X3>::STAGE_BYTES;
  p.off_w0h = p.off_tail + sizeof(float) * TAIL_FLOATS;
  p.off_f16 = p.off_w0h + W0_BYTES;
  p.off_tailh = p.off_f16 + (size_t)3 * ModeCfg<MBO_MLP_FP16>::STAGES_PER_LAYER * ModeCfg<MBO_MLP_FP16>::STAGE_BYTES;
  p.off_scale = p.off_tailh + sizeof(float) * TAIL_FLOATS;
  p.off_ssq = p.off_scale + 32;          // tc_scale_kernel scratch: 8 sums of squares (fp64) + a block counter
  p.total = p.off_ssq + 96;
  return p;
}

// ------------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t WAIT_TIMEOUT_NS = 2000000000ull;   // 2 s: far beyond any legitimate wait
// Bounded wait.  On a time-out (a pipeline bug) the CTA-wide abort flag is latched and from then on every wait of
// every role returns immediately: the roles run through their remaining loop iterations on garbage, the kernel ends,
// and the host sees bit 0 of the error word.  The slow path is out of line and the call sites carry no result
// handling: they are inlined ~45 times, and the kernel's instruction footprint is what the instruction caches (and
// the GPC-level fetch path behind them) see.
__device__ __forceinline__ int lds_i32(uint32_t addr) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t abort_addr) {
  if (lds_i32(abort_addr)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (lds_i32(abort_addr)) return;
      if (globaltimer_ns() - t0 > WAIT_TIMEOUT_NS) {
        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(abort_addr), "r"(1) : "memory");
        return;
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t abort_addr) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, abort_addr);
}
// Two waits whose ~90-cycle try_wait latencies overlap (both probes are issued before either result is used).
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b,
                                           uint32_t abort_addr) {
  uint32_t ok_a, ok_b;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "selp.u32 %1, 1, 0, q;\n\t}"
      : "=r"(ok_a), "=r"(ok_b) : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b) : "memory");
  if (!ok_a) mbar_wait_slow(bar_a, par_a, abort_addr);
  if (!ok_b) mbar_wait_slow(bar_b, par_b, abort_addr);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_bf16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, no-swizzle canonical layout: LBO = byte distance between the two 16-byte K chunks of one
// MMA, SBO = byte distance between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(CHUNK_ROW_BYTES >> 4) << 16) |
         ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) {          // start address + LBO (low word)
  return ((saddr & 0x3FFFFu) >> 4) | ((uint32_t)(CHUNK_ROW_BYTES >> 4) << 16);
}
__device__ __forceinline__ uint64_t make_desc64(uint32_t lo) {          // high word: SBO = 128 B, version 1
  constexpr uint32_t HI = (uint32_t)(128 >> 4) | (1u << 14);
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(HI));
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// cute::UMMA::InstrDescriptor: c_format f32, a/b format, K-major both, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_format) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
constexpr uint32_t IDESC_TF32 = make_idesc(2), IDESC_BF16 = make_idesc(1), IDESC_F16 = make_idesc(0);

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* W[5];
  const float* b[5];
  int F, H;
};

// fp16 mode: per-layer power-of-two renormalisation.  The RMS magnitude of layer l's output is estimated from the
// weights alone: est_0 = 1, est_{l+1}^2 = g_l^2 est_l^2 + rms(b_l)^2 with g_l the RMS row norm of W_l.  The
// cumulative exponent E_l = -rint(log2 est_{l+1}) (0 while |E_l| < 6, so the shipped policies run unscaled) brings
// the activations handed to layer l+1 to O(1); layer l is packed as 2^(E_l - E_{l-1}) W_l with bias 2^E_l b_l and
// the head as 2^-E_3 w4 -- the same function in exact arithmetic because relu(c x) = c relu(x) for c > 0.
// exps[0..3] = per-layer exponents, exps[4] = sticky "a packed value left fp16's range" flag (set by the pack).
// Eight CTAs, one per sum (q < 4: sum W_q^2, q >= 4: sum b_{q-4}^2); the last CTA to finish derives the exponents
// (one CTA doing the eight sums in sequence took 220 us per weight sync).  `scratch`: 8 doubles + a counter that the
// last CTA resets, zero-initialised once when the pack buffer is allocated.
__global__ void tc_scale_kernel(PackArgs a, int* __restrict__ exps, double* __restrict__ scratch) {
  __shared__ double sh[256];
  __shared__ bool last;
  const int F = a.F, H = a.H;
  const int q = blockIdx.x, l = q & 3;
  {
    const int cnt = q < 4 ? H * (l == 0 ? F : H) : H;
    const float* src = q < 4 ? a.W[l] : a.b[l];
    double acc = 0.0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) { const double w = src[i]; acc += w * w; }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
  }
  unsigned int* counter = (unsigned int*)(scratch + 8);
  if (threadIdx.x == 0) {
    scratch[q] = sh[0];
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    __threadfence();
    volatile double* ssq = scratch;
    *counter = 0u;
    double est2 = 1.0;
    int Eprev = 0;
    for (int l2 = 0; l2 < 4; ++l2) {
      est2 = ssq[l2] / H * est2 + ssq[4 + l2] / H;
      int E = Eprev;
      if (est2 > 0.0 && est2 < 1e300) {       // zero, inf or NaN parameters: keep the previous scale
        E = -(int)rint(0.5 * log2(est2));
        if (E > -6 && E < 6) E = 0;
        E = E > Eprev + 20 ? Eprev + 20 : (E < Eprev - 20 ? Eprev - 20 : E);
        E = E > 60 ? 60 : (E < -60 ? -60 : E);
      }
      exps[l2] = E - Eprev;
      Eprev = E;
    }
    exps[4] = 0;
  }
}

__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {   // lo -> bits 0..15
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_bits_to_float(uint32_t h) {
  return __half2float(__ushort_as_half((unsigned short)h));
}

// bf16x3_only: pack just what the bf16x3 mode reads (layer 0, the bf16 hi / lo hidden stages, the fp32 tail) -- the PPO
// update re-packs the current weights on every optimiser step and needs neither the tf32 nor the fp16 images
__global__ void tc_pack_kernel(PackArgs a, unsigned char* __restrict__ out, int bf16x3_only) {
  const PackLayout pl = pack_layout();
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit word index
  const size_t byte = w * 4;
  if (byte >= pl.total) return;
  if (bf16x3_only && ((byte >= pl.off_tf32 && byte < pl.off_bf16) || byte >= pl.off_w0h)) return;
  uint32_t* dst = (uint32_t*)(out + byte);
  const int F = a.F, H = a.H;
  if (byte < pl.off_tf32) {
    // W0: [part][chunk][n][4 floats]; k = chunk*4 + e; column F carries the bias
    size_t i = byte / 4;
    int e = i % 4; i /= 4;
    int n = i % TC_N; i /= TC_N;
    int chunk = i % (TC_K0 / 4); i /= (TC_K0 / 4);
    int part = (int)i;
    int k = chunk * 4 + e;
    float v = 0.f;
    if (n < H) v = k < F ? a.W[0][n * F + k] : (k == F ? a.b[0][n] : 0.f);
    float hi = tf32_rna(v);
    float lo = tf32_rna(v - hi);
    *dst = __float_as_uint(part == 0 ? hi : lo);
  } else if (byte < pl.off_bf16) {
    using C = ModeCfg<MBO_MLP_TF32>;
    size_t i = (byte - pl.off_tf32) / 4;
    int e = i % 4; i /= 4;
    int n = i % TC_N; i /= TC_N;
    int chunk = i % 2; i /= 2;
    int slice = i % C::SLICES_PER_STAGE; i /= C::SLICES_PER_STAGE;
    int stage = i % C::STAGES_PER_LAYER; i /= C::STAGES_PER_LAYER;
    int layer = (int)i + 1;
    int k = (stage * C::SLICES_PER_STAGE + slice) * C::SLICE_K + chunk * 4 + e;
    float v = (n < H && k < H) ? a.W[layer][n * H + k] : 0.f;
    *dst = __float_as_uint(tf32_rna(v));
  } else if (byte < pl.off_tail) {
    using C = ModeCfg<MBO_MLP_BF16X3>;
    size_t i = (byte - pl.off_bf16) / 4;
    int pair = i % 4; i /= 4;
    int n = i % TC_N; i /= TC_N;
    int chunk = i % 2; i /= 2;
    int part = i % 2; i /= 2;
    int stage = i % C::STAGES_PER_LAYER; i /= C::STAGES_PER_LAYER;
    int layer = (int)i + 1;
    int k = stage * C::SLICE_K + chunk * 8 + pair * 2;
    float v0 = (n < H && k < H) ? a.W[layer][n * H + k] : 0.f;
    float v1 = (n < H && k + 1 < H) ? a.W[layer][n * H + k + 1] : 0.f;
    __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    __nv_bfloat16 o0 = h0, o1 = h1;
    if (part == 1) {
      o0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
      o1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
    }
    *dst = (uint32_t)__bfloat16_as_ushort(o0) | ((uint32_t)__bfloat16_as_ushort(o1) << 16);
  } else if (byte < pl.off_w0h) {
    size_t i = (byte - pl.off_tail) / 4;
    float v = 0.f;
    if (i < (size_t)TC_N) {
      v = 0.f;
    } else if (i < (size_t)4 * TC_N) {
      int l = (int)(i / TC_N), n = (int)(i % TC_N);
      v = n < H ? a.b[l][n] : 0.f;
    } else if (i < (size_t)5 * TC_N) {
      int n = (int)(i - 4 * TC_N);
      v = n < H ? a.W[4][n] : 0.f;
    } else if (i == (size_t)5 * TC_N) {
      v = a.b[4][0];
    }
    *dst = __float_as_uint(v);
  } else if (byte < pl.off_scale) {
    // ---- fp16 mode (needs H + 2 <= 208): scaled weights, bias and ones columns inside the MMAs ----
    const int* ex = (const int*)(out + pl.off_scale);          // written by tc_scale_kernel (stream order)
    if (byte < pl.off_f16) {
      // W0h: as W0, scaled by 2^e0; rows H, H+1 read the features' ones column -> D0[:, H] = D0[:, H+1] = 1
      size_t i = (byte - pl.off_w0h) / 4;
      int e = i % 4; i /= 4;
      int n = i % TC_N; i /= TC_N;
      int chunk = i % (TC_K0 / 4); i /= (TC_K0 / 4);
      int part = (int)i;
      int k = chunk * 4 + e;
      float v = 0.f;
      if (n < H) v = ldexpf(k < F ? a.W[0][n * F + k] : (k == F ? a.b[0][n] : 0.f), ex[0]);
      else if (n < H + 2 && k == F) v = 1.0f;
      float hi = tf32_rna(v);
      float lo = tf32_rna(v - hi);
      *dst = __float_as_uint(part == 0 ? hi : lo);
    } else if (byte < pl.off_tailh) {
      using C = ModeCfg<MBO_MLP_FP16>;
      size_t i = (byte - pl.off_f16) / 4;
      int pair = i % 4; i /= 4;
      int n = i % TC_N; i /= TC_N;
      int chunk = i % 2; i /= 2;
      int slice = i % C::SLICES_PER_STAGE; i /= C::SLICES_PER_STAGE;
      int stage = i % C::STAGES_PER_LAYER; i /= C::STAGES_PER_LAYER;
      int layer = (int)i + 1;
      int k = (stage * C::SLICES_PER_STAGE + slice) * C::SLICE_K + chunk * 8 + pair * 2;
      int cum = 0;
      for (int l = 0; l <= layer; ++l) cum += ex[l];
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int kk = k + u;
        float x = 0.f;
        if (n < H) {
          if (kk < H) x = ldexpf(a.W[layer][n * H + kk], ex[layer]);
          else if (kk == H || kk == H + 1) {
            const float bs = ldexpf(a.b[layer][n], cum);
            const float bh = f16_bits_to_float(pack_f16x2_sat(bs, 0.f) & 0xffffu);
            x = kk == H ? bs : bs - bh;          // hi is rounded by the final conversion (and range-checked there)
          }
        } else if (n < H + 2 && kk == H) x = 1.0f;           // keeps the ones columns alive in the next layer
        v[u] = x;
      }
      if (!(fabsf(v[0]) <= 65504.f) || !(fabsf(v[1]) <= 65504.f)) atomicOr((int*)(out + pl.off_scale) + 4, 1);
      *dst = pack_f16x2_sat(v[0], v[1]);
    } else {
      // tail of the fp16 mode: biases live in the MMAs; head weights undo the cumulative scale
      size_t i = (byte - pl.off_tailh) / 4;
      float v = 0.f;
      if (i >= (size_t)4 * TC_N && i < (size_t)5 * TC_N) {
        int n = (int)(i - 4 * TC_N);
        v = n < H ? ldexpf(a.W[4][n], -(ex[0] + ex[1] + ex[2] + ex[3])) : 0.f;
      } else if (i == (size_t)5 * TC_N) {
        v = a.b[4][0];
      }
      *dst = __float_as_uint(v);
    }
  }
}

int tc_pack_weights(mbo_policy* p, cudaStream_t s) {
  p->tc_ready = 0;
  if (p->n_layers != 5 || p->activation != MBO_ACT_RELU) return MBO_OK;
  const int F = p->widths[0], H = p->widths[1];
  if (F + 1 > TC_K0 || H > TC_N || H < 8 || p->widths[5] != 1) return MBO_OK;   // F <= 15
  for (int l = 1; l <= 4; ++l)
    if (p->widths[l] != H) return MBO_OK;
  const PackLayout pl = pack_layout();
  if (!p->d_tc || p->tc_bytes != pl.total) {
    if (p->d_tc) cudaFree(p->d_tc);
    p->d_tc = nullptr;
    MBO_CUDA(cudaMalloc(&p->d_tc, pl.total));
    MBO_CUDA(cudaMemsetAsync((unsigned char*)p->d_tc + pl.off_ssq, 0, 96, s));     // tc_scale_kernel's block counter
    p->tc_bytes = pl.total;
  }
  PackArgs a;
  for (int l = 0; l < 5; ++l) { a.W[l] = p->d_W_raw + p->raw_off[l]; a.b[l] = p->d_b[l]; }
  a.F = F; a.H = H;
  size_t words = pl.total / 4;
  tc_scale_kernel<<<8, 256, 0, s>>>(a, (int*)((unsigned char*)p->d_tc + pl.off_scale), (double*)((unsigned char*)p->d_tc + pl.off_ssq));
  MBO_LAUNCH_CHECK();
  tc_pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(a, (unsigned char*)p->d_tc, 0);
  MBO_LAUNCH_CHECK();
  p->tc_ready = 1;
  p->tc_f16 = H + 2 <= TC_N;     // two spare K columns for the bias
  p->tc_F = F; p->tc_H = H;
  return MBO_OK;
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
struct TcFeat {
  int F;
  int k0_slices;       // 1: F + 1 <= 8 (one K=8 slice for layer 0), 2: F + 1 <= 16
  int src[TC_K0];      // MBO_FEAT_* codes, or state-column indices in dense mode
  const float* dense;  // [B, M, dense_F] materialised states, or nullptr
  int dense_F;
  int standard;        // 1: src = [MEAN, STD, X0 .. X(D-1), per-episode scalars ...] -- get_state's own column order
                       // (metabo_gym.py:624-677) with any of the scalar columns; branch-free gather
};

// MBO_MLP_FP16 (not FUSE) keeps most weight stages resident in shared memory: all of layer 2 and the first NRES1
// stages of layers 1 and 3.  The remaining stages of layers 1 and 3 (chunks NRES1..6) time-share one set of slots:
// when layer 1 is complete (an epilogue thread sees d_full) the slots are refilled with layer 3's stages, when layer
// 3 is complete with the next tile's layer-1 stages -- 120 KB instead of 260 KB per tile from L2, issued from code
// that is hot anyway.  The chip-wide L2 -> SM rate (~6 300 B/clk, 43 B/clk per SM) is what bounds a kernel that
// re-streams all weights for every 128-candidate tile (tf32 mode: 520 KB per tile = 12 200 clk); and a dedicated
// producer warp whose loop body runs once per ~900 cycles pays an instruction-cache miss on every pass.
constexpr int STG_LD = 36;   // STORE staging tile: 32 fp32 columns per row + 4 of padding (float4 rows stay conflict-free)
template <int MODE, bool FUSE = false, int K0S = 2, bool STORE = false>
struct TcSmem {
  using C = ModeCfg<MODE>;
  static constexpr bool RES = MODE == MBO_MLP_FP16 && !FUSE;
  static constexpr int NRES1 = K0S == 1 ? 2 : 1;                      // resident leading stages of layers 1 and 3
  static constexpr int NSLOT = C::STAGES_PER_LAYER - NRES1;           // time-shared slots (the last one is half a stage)
  static constexpr int NA = NSLOT - 3;                                // slots [0, NA) arrive with the first refill copy, the rest with the second
  static constexpr int NS = RES ? NSLOT : ((FUSE || STORE) ? C::NSTAGE_FUSED : C::NSTAGE);   // STORE trades three stages for its staging tiles
  static constexpr size_t SLOT_BYTES = RES ? (size_t)(NSLOT - 1) * C::STAGE_BYTES + C::STAGE_BYTES / 2 : 0;
  static constexpr size_t L2_BYTES = (size_t)6 * C::STAGE_BYTES + C::STAGE_BYTES / 2;
  static constexpr size_t RES_BYTES = RES ? (size_t)2 * NRES1 * C::STAGE_BYTES + L2_BYTES : 0;   // [L1 | L2 | L3]
  static constexpr size_t W0_SMEM = RES ? (size_t)K0S * (W0_BYTES / 2) : (size_t)W0_BYTES;   // RES: only the slices in use
  static constexpr size_t off_ring = 0;
  static constexpr size_t off_res = off_ring + (RES ? SLOT_BYTES : (size_t)NS * C::STAGE_BYTES);
  static constexpr size_t off_w0 = off_res + RES_BYTES;
  static constexpr size_t off_tail = off_w0 + W0_SMEM;
  static constexpr size_t off_bars = off_tail + sizeof(float) * TAIL_FLOATS;
  static constexpr int n_bars = 2 * NS + C::NCHUNK + 2 + 3 + 4 + 1;   // ring, a_ready, d_full[2], feat_full, s_free, w0_full, gp_full[2], gp_empty[2], r1_free
  static constexpr size_t off_misc = off_bars + ((8 * n_bars + 15) & ~(size_t)15);   // keep what follows 16-byte aligned
  static constexpr size_t off_part = off_misc + 16;
  static constexpr size_t off_gp = off_part + sizeof(float) * TC_M;             // fused GP: episode blob + mu/sd hand-off
  static constexpr size_t off_musd = off_gp + (FUSE ? sizeof(double) * GPB_DOUBLES : 0);
  static constexpr size_t off_ks = off_musd + (FUSE ? sizeof(float) * 4 * TC_M : 0);        // k* scratch [32][128] fp64
  static constexpr size_t off_stg = off_ks + (FUSE ? sizeof(double) * GP_NP * TC_M : 0);   // STORE: [8 warps][32 rows][STG_LD] fp32
  static constexpr size_t total = off_stg + (STORE ? sizeof(float) * 8 * 32 * STG_LD : 0);
};

// per-mode conversion of one accumulator chunk into the next layer's A operand (in registers)
template <int MODE> struct ChunkOps;
template <> struct ChunkOps<MBO_MLP_TF32> {
  // 40 columns: relu(acc + bias) rounded to tf32, stored in place
  static __device__ __forceinline__ void load(uint32_t taddr, uint32_t* v, int) {
    tmem_ld32(taddr, v);
    tmem_ld8(taddr + 32, v + 32);
  }
  static __device__ __forceinline__ void convert_store(uint32_t taddr, uint32_t* v, const float* bias, int) {
    const float4* b4 = reinterpret_cast<const float4*>(bias);      // shared memory, 16-byte aligned
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const float4 b = b4[j];
      // relu(acc + bias), then round-to-nearest to tf32 by adding half an ulp: the tensor core ignores
      // the low 13 mantissa bits of a kind::tf32 operand, so no masking is needed
      v[4 * j + 0] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 0]) + b.x, 0.f)) + 0x1000u;
      v[4 * j + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 1]) + b.y, 0.f)) + 0x1000u;
      v[4 * j + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 2]) + b.z, 0.f)) + 0x1000u;
      v[4 * j + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 3]) + b.w, 0.f)) + 0x1000u;
    }
    tmem_st32(taddr, v);
    tmem_st8(taddr + 32, v + 32);
  }
};
template <> struct ChunkOps<MBO_MLP_BF16X3> {
  // 32 columns (16 for the last chunk) -> [hi: bf16 pairs | lo: bf16 pairs] in place
  static __device__ __forceinline__ void load(uint32_t taddr, uint32_t* v, int ncols) {
    if (ncols == 32) tmem_ld32(taddr, v);
    else tmem_ld16(taddr, v);
  }
  static __device__ __forceinline__ void convert_store(uint32_t taddr, uint32_t* v, const float* bias, int ncols) {
    uint32_t ph[16], plo[16];
    const float2* b2 = reinterpret_cast<const float2*>(bias);      // shared memory, 8-byte aligned
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a0 = 0.f, a1 = 0.f;
      if (2 * j < ncols) {
        const float2 b = b2[j];
        a0 = fmaxf(__uint_as_float(v[2 * j]) + b.x, 0.f);
        a1 = fmaxf(__uint_as_float(v[2 * j + 1]) + b.y, 0.f);
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);      // .x (low half) = element 2j
      const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
      const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a0 - h0, a1 - h1);
      ph[j] = hb;
      plo[j] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    if (ncols == 32) {
      tmem_st16(taddr, ph);
      tmem_st16(taddr + 16, plo);
    } else {   // last chunk: 16 K values -> 8 columns hi at +0, 8 columns lo at +8 (stays inside the region)
      tmem_st8(taddr, ph);
      tmem_st8(taddr + 8, plo);
    }
  }
};

template <> struct ChunkOps<MBO_MLP_FP16> {
  // 32 accumulator columns -> 16 columns of fp16 pairs at the start of the chunk's own columns; the bias is already
  // in the accumulator.  The last chunk has 16 real columns: it is processed like the others (one code path, no
  // run-time choice between load/store shapes) -- the 16 extra columns it reads belong to the neighbouring TMEM
  // region, and the 8 extra columns it writes (K = 208..223) are never an MMA operand.  `sat` tracks the largest
  // converted bit pattern of the real columns (values are >= 0, so unsigned 16-bit compares order them) for the
  // range check at the end of the kernel.
  static __device__ __forceinline__ void load(uint32_t taddr, uint32_t* v, int) { tmem_ld32(taddr, v); }
  static __device__ __forceinline__ void convert_store(uint32_t taddr, uint32_t* v, int ncols, uint32_t& sat) {
    uint32_t ph[16];
    const uint32_t tail_mask = ncols == 32 ? 0xffffffffu : 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;"
          : "=r"(ph[j]) : "f"(__uint_as_float(v[2 * j + 1])), "f"(__uint_as_float(v[2 * j])));
      sat = __vmaxu2(sat, j < 8 ? ph[j] : (ph[j] & tail_mask));
    }
    tmem_st16(taddr, ph);
  }
};

struct GpFuse {
  const double* gp_state;   // [B, gp_state_doubles(nmax, D)]
  int nmax, kid;
  float* mean_out;          // optional [B, M] copies of the fused posterior (PPO state buffer), may be null
  float* std_out;
};

// fused top-k epilogue: every epilogue warp (32 candidate rows) emits its k best (logit, index) pairs into
// part_val / part_idx [(episode * tiles_per_episode + tile) * 4 + quad][8]; topk_merge_kernel reduces them per episode
struct TcTopk {
  int k;                 // 0 = off, else 1..8
  float* part_val;
  int32_t* part_idx;
};

// STORE (PPO update forward, ppo.py:146-156): besides the logits the kernel writes what the backward needs -- the four
// post-ReLU activations as plain fp32 rows [rows][208] (row = episode * M + candidate; column H is a ones column: the
// wgrad GEMMs take their bias gradient from it) and the ReLU decisions of h1..h3 as bit words [rows][8].
struct TcStore {
  float* h[4];
  uint32_t* bits[3];
  int H;
};

template <int MODE, bool FUSE, int K0S, bool STORE = false>     // K0S: K = 8 slices of layer 0 (1: F + 1 <= 8, 2: F + 1 <= 16)
// Register budget: the SM's register file is four 16 K sub-partition files and this CTA puts 4 of its 14 warps on two of
// them, so above 104 registers no warp of another kernel can co-reside (tools/micro/coresidency.cu).  Co-residency with
// the fp64 GP kernel was measured at 104 registers and does not pay (both kernels live on the shared-memory data path,
// profiles/r02_notes.md section 6), so the kernel takes the registers the head epilogue's double-buffered loads want.
__global__ void __launch_bounds__(FUSE ? TC_THREADS_FUSED : TC_THREADS, 1)
neural_af_tc_kernel(const unsigned char* __restrict__ pack, TcFeat fs, mbo_candidates cand, int B,
                    int tiles_per_episode, const float* __restrict__ mean, const float* __restrict__ std_,
                    const float* __restrict__ epi, float* __restrict__ logits, int* __restrict__ err_flag, int debug_noload,
                    GpFuse gp, TcTopk tk, TcStore st) {
  static_assert(!STORE || (MODE == MBO_MLP_BF16X3 && !FUSE), "STORE: bf16x3 (fp32-class) forward of the PPO update only");
  using C = ModeCfg<MODE>;
  using SM = TcSmem<MODE, FUSE, K0S, STORE>;
  constexpr int NTHREADS = FUSE ? TC_THREADS_FUSED : TC_THREADS;
  constexpr int NS = SM::NS;                      // ring depth
  constexpr bool RES = SM::RES;                   // fp16: even weight stages resident, odd ones streamed
  // fp16 (not FUSE): the feature warps own layer 0's epilogue.  At a tile boundary the epilogue warps are busy with
  // the previous tile's head layer for ~1 500 cycles; layer 1 no longer waits for them.  (The tf32 / bf16x3 chunk
  // conversions are register-heavy -- a second instance of them spills -- and those modes are bound by weight
  // streaming rather than by this hand-off.)
  constexpr bool E0_BY_FEAT = MODE == MBO_MLP_FP16 && !FUSE;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackLayout pl = pack_layout();

  unsigned char* ring = smem + SM::off_ring;
  float* s_tail = (float*)(smem + SM::off_tail);
  uint64_t* bars = (uint64_t*)(smem + SM::off_bars);
  uint32_t* s_tmem = (uint32_t*)(smem + SM::off_misc);
  volatile int* s_abort_p = (volatile int*)(smem + SM::off_misc + 4);
  const uint32_t s_abort = smem_u32(smem + SM::off_misc + 4);     // shared-window address: cheap to pass to the slow wait path
  float* s_part = (float*)(smem + SM::off_part);          // [128] partial head sums of the upper column half

  const uint32_t bar0 = smem_u32(bars);
  auto BAR_WFULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto BAR_WEMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(NS + s); };
  auto BAR_AREADY = [&](int c) { return bar0 + 8u * (uint32_t)(2 * NS + c); };
  auto BAR_DFULL = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NS + C::NCHUNK + i); };
  const uint32_t BAR_FEAT = bar0 + 8u * (uint32_t)(2 * NS + C::NCHUNK + 2);
  const uint32_t BAR_SFREE = BAR_FEAT + 8u;
  const uint32_t BAR_W0 = BAR_FEAT + 16u;
  auto BAR_GPFULL = [&](int i) { return BAR_FEAT + 24u + 8u * (uint32_t)i; };
  auto BAR_GPEMPTY = [&](int i) { return BAR_FEAT + 40u + 8u * (uint32_t)i; };
  const uint32_t BAR_R1FREE = BAR_FEAT + 56u;       // the head epilogue has read the previous tile's D3 out of R1
  float* s_mu = (float*)(smem + SM::off_musd);            // [2][128]
  float* s_sd = s_mu + 2 * TC_M;                          // [2][128]

#ifdef MBO_TC_TRACE
  const long long trace_c0 = clock64();
  const uint64_t trace_g0 = globaltimer_ns();
#endif
  const int total_tiles = B * tiles_per_episode;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup ----
  if (tid == 0) {
    *s_abort_p = 0;
    for (int s = 0; s < NS; ++s) { mbar_init(BAR_WFULL(s), 1); mbar_init(BAR_WEMPTY(s), 1); }
    for (int c = 0; c < C::NCHUNK; ++c) mbar_init(BAR_AREADY(c), 4);     // one elected arrive per epilogue warp
    mbar_init(BAR_DFULL(0), 1);
    mbar_init(BAR_DFULL(1), 1);
    mbar_init(BAR_FEAT, 4);
    mbar_init(BAR_SFREE, 1);
    mbar_init(BAR_W0, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR_GPFULL(i), 128); mbar_init(BAR_GPEMPTY(i), 128); }
    mbar_init(BAR_R1FREE, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr bool F16 = MODE == MBO_MLP_FP16;
  for (int i = tid; i < TAIL_FLOATS; i += NTHREADS)
    s_tail[i] = ((const float*)(pack + (F16 ? pl.off_tailh : pl.off_tail)))[i];
  if (warp == WARP_TMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // RES: (re)fill the time-shared weight slots with stages NRES1..6 of hidden layer `li` (0: layer 1, 2: layer 3).
  // Called by ONE thread, after it has observed the completion of every MMA that read the slots' old contents.
  auto refill_slots = [&](int li) {
    if constexpr (RES) {
      // two bulk copies (issuing one costs the thread ~160 cycles): slots [0, NA) on barrier 0, [NA, NSLOT) on barrier NA
      const unsigned char* src = pack + pl.off_f16 + (size_t)(li * C::STAGES_PER_LAYER + SM::NRES1) * C::STAGE_BYTES;
      const uint32_t dst = smem_u32(ring);
      constexpr uint32_t BYTES_A = (uint32_t)SM::NA * C::STAGE_BYTES;
      constexpr uint32_t BYTES_B = (uint32_t)SM::SLOT_BYTES - BYTES_A;
      mbar_expect_tx(BAR_WFULL(0), BYTES_A);
      bulk_g2s(dst, src, BYTES_A, BAR_WFULL(0));
      mbar_expect_tx(BAR_WFULL(SM::NA), BYTES_B);
      bulk_g2s(dst + BYTES_A, src + BYTES_A, BYTES_B, BAR_WFULL(SM::NA));
    }
  };

  if (warp == WARP_TMA) {
    // ================= TMA producer (warp-uniform control flow, one elected lane issues) =================
    const bool leader = elect_one();
    if (my_tiles > 0) {
      const unsigned char* wsrc = pack + (MODE == MBO_MLP_TF32 ? pl.off_tf32 : (F16 ? pl.off_f16 : pl.off_bf16));
      if (leader) {
        if constexpr (RES) {
          // layer-0 weights (only the K slices in use, hi and lo part) + the resident stages, one barrier
          mbar_expect_tx(BAR_W0, (uint32_t)(SM::W0_SMEM + SM::RES_BYTES));
          constexpr uint32_t PART = (uint32_t)(SM::W0_SMEM / 2);
          bulk_g2s(smem_u32(smem + SM::off_w0), pack + pl.off_w0h, PART, BAR_W0);
          bulk_g2s(smem_u32(smem + SM::off_w0) + PART, pack + pl.off_w0h + W0_BYTES / 2, PART, BAR_W0);
          const uint32_t res = smem_u32(smem + SM::off_res);
          for (int j = 0; j < SM::NRES1; ++j) {
            bulk_g2s(res + j * C::STAGE_BYTES, wsrc + (size_t)j * C::STAGE_BYTES, C::STAGE_BYTES, BAR_W0);
            bulk_g2s(res + (uint32_t)(SM::NRES1 * C::STAGE_BYTES + SM::L2_BYTES) + j * C::STAGE_BYTES,
                     wsrc + (size_t)(2 * C::STAGES_PER_LAYER + j) * C::STAGE_BYTES, C::STAGE_BYTES, BAR_W0);
          }
          for (int j = 0; j < C::STAGES_PER_LAYER; ++j)
            bulk_g2s(res + (SM::NRES1 + j) * C::STAGE_BYTES, wsrc + (size_t)(C::STAGES_PER_LAYER + j) * C::STAGE_BYTES,
                     j < C::STAGES_PER_LAYER - 1 ? C::STAGE_BYTES : C::STAGE_BYTES / 2, BAR_W0);
          refill_slots(0);                            // the first tile's layer-1 stages
        } else {
          mbar_expect_tx(BAR_W0, W0_BYTES);
          bulk_g2s(smem_u32(smem + SM::off_w0), pack + (F16 ? pl.off_w0h : pl.off_w0), W0_BYTES, BAR_W0);
        }
      }
      const uint32_t ring_addr = smem_u32(ring);
      uint32_t stage = 0, phase = 0;
#pragma unroll 1
      for (int t = 0; t < (RES ? 0 : my_tiles); ++t) {      // RES: no streaming loop, the epilogue refills the slots
#pragma unroll 1
        for (int ls = 0; ls < 3 * C::STAGES_PER_LAYER; ++ls) {
          mbar_wait(BAR_WEMPTY(stage), phase ^ 1u, s_abort);
          if (leader) {
            if (debug_noload && (t > 0 || ls >= NS)) {
              mbar_arrive(BAR_WFULL(stage));    // timing experiment only: pretend the stage is resident
            } else {
              // fp16 mode: the last stage of a layer holds one slice
              const uint32_t bytes = (F16 && (ls % C::STAGES_PER_LAYER) == C::STAGES_PER_LAYER - 1) ? C::STAGE_BYTES / 2
                                                                                                  : C::STAGE_BYTES;
              mbar_expect_tx(BAR_WFULL(stage), bytes);
              bulk_g2s(ring_addr + stage * C::STAGE_BYTES, wsrc + (size_t)ls * C::STAGE_BYTES, bytes, BAR_WFULL(stage));
            }
          }
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer (warp-uniform control flow, one elected lane issues) =================
    // Everything that feeds a tcgen05.mma is a base register plus a compile-time constant: the single
    // issuing thread must spend far fewer cycles per MMA than the ~104 the tensor pipe needs for it.
    const bool leader = elect_one();
    if (my_tiles > 0) {
      mbar_wait(BAR_W0, 0, s_abort);
      uint32_t stage = 0, wphase = 0, aphase = 0, fphase = 0;
      long long st_feat = 0, st_a = 0, st_w = 0;
      (void)st_feat; (void)st_a; (void)st_w;
      const uint32_t w0_lo = desc_lo(smem_u32(smem + SM::off_w0));
      const uint32_t ring_lo = desc_lo(smem_u32(ring));
      const uint32_t res_lo = desc_lo(smem_u32(smem + SM::off_res));
      (void)res_lo;
      constexpr uint32_t STAGE_D = C::STAGE_BYTES >> 4;            // descriptor units (16 B)
      constexpr uint32_t SLICE_D = (2 * CHUNK_ROW_BYTES) >> 4;     // one K slice = two 16-byte chunks
      for (int t = 0; t < my_tiles; ++t) {
        // ---- L0: D(R0) = S * W0^T, 3xTF32 ----
        {
          TC_STAT_BEGIN();
          mbar_wait(BAR_FEAT, fphase, s_abort);
          TC_STAT_END(st_feat);
        }
        fphase ^= 1u;
        tc_fence_after();
        TC_TRACE(0);
        if (leader) {
          const uint32_t a_hi = tmem + COL_S, a_lo = tmem + COL_S + 16;
          constexpr uint32_t W0_PART_D = (uint32_t)(RES ? SM::W0_SMEM / 2 : (TC_K0 / 4) * CHUNK_ROW_BYTES) >> 4;   // hi -> lo part
#pragma unroll
          for (int sl = 0; sl < K0S; ++sl) {
            const uint32_t wl = w0_lo + sl * SLICE_D;
            mma_ts_tf32(tmem + COL_R0, a_hi + sl * 8, make_desc64(wl), IDESC_TF32, sl ? 1u : 0u);
            mma_ts_tf32(tmem + COL_R0, a_lo + sl * 8, make_desc64(wl), IDESC_TF32, 1);
            mma_ts_tf32(tmem + COL_R0, a_hi + sl * 8, make_desc64(wl + W0_PART_D), IDESC_TF32, 1);
          }
          tc_commit(BAR_SFREE);
          if (!E0_BY_FEAT) tc_commit(BAR_DFULL(0));
        }
        TC_TRACE(1);
        if (RES && t > 0) {
          // The next use of the shared slots is this tile's layer 1, ~2 000 cycles away: refill them as soon as the
          // previous tile's layer 3 is complete.  The issuing thread is idle until layer 0's epilogue anyway.
          mbar_wait(BAR_DFULL(1), 1u, s_abort);
          if (leader) refill_slots(0);
        }
        // ---- L1..L3 (rolled: three copies of the chunk loop are instruction-cache footprint for nothing) ----
#pragma unroll 1
        for (int l = 1; l <= 3; ++l) {
          const uint32_t tA = tmem + ((l & 1) ? COL_R0 : COL_R1);
          const uint32_t tD = tmem + ((l & 1) ? COL_R1 : COL_R0);
          // RES: resident weights of this layer ([L1 | L2 | L3] region)
          const uint32_t res_l = res_lo + (uint32_t)((l == 1 ? 0 : (l == 2 ? SM::NRES1 * C::STAGE_BYTES
                                                                           : SM::NRES1 * C::STAGE_BYTES + SM::L2_BYTES)) >> 4);
          (void)res_l;
          if (l == 1 && t > 0) {
            // L1 overwrites R1: the head epilogue of the previous tile must have its accumulators in registers.
            // (Layer 0's first chunks come from the feature warps, so a_ready alone no longer implies this.)
            mbar_wait(BAR_R1FREE, (uint32_t)(t - 1) & 1u, s_abort);
          }
          if constexpr (F16 && !RES) {
            // The ring holds two whole layers, so this layer's 7 stages have normally been resident for a long
            // time: check them all now, while the epilogue is still converting the first chunk.  A ready
            // mbarrier wait still costs ~80 cycles of latency, and with two 104-cycle MMAs per chunk the issuing
            // thread cannot afford two of them per chunk.
            TC_STAT_BEGIN();
            uint32_t st = stage, ph = wphase;
#pragma unroll
            for (int i = 0; i < C::STAGES_PER_LAYER; ++i) {
              mbar_wait(BAR_WFULL(st), ph, s_abort);
              if (++st == NS) { st = 0; ph ^= 1u; }
            }
            TC_STAT_END(st_w);
          }
#pragma unroll
          for (int c = 0; c < C::NCHUNK; ++c) {
            {
              TC_STAT_BEGIN();
              if (RES && c >= SM::NRES1 && l != 2)      // this chunk's weights sit in a shared slot: probe both barriers at once
                mbar_wait2(BAR_AREADY(c), aphase, BAR_WFULL(c - SM::NRES1 < SM::NA ? 0 : SM::NA), l == 1 ? 0u : 1u, s_abort);
              else
                mbar_wait(BAR_AREADY(c), aphase, s_abort);
              TC_STAT_END(st_a);
            }
            TC_TRACE(10 * l + c);
            if (MODE == MBO_MLP_TF32) {
              // chunk c == ring stage: 5 slices of K=8
              {
                TC_STAT_BEGIN();
                mbar_wait(BAR_WFULL(stage), wphase, s_abort);
                TC_STAT_END(st_w);
              }
              tc_fence_after();
              if (leader) {
                const uint32_t lo = ring_lo + stage * STAGE_D;
#pragma unroll
                for (int sc = 0; sc < 5; ++sc)
                  mma_ts_tf32(tD, tA + (c * 5 + sc) * 8, make_desc64(lo + sc * SLICE_D), IDESC_TF32, (c | sc) ? 1u : 0u);
                tc_commit(BAR_WEMPTY(stage));
              }
              if (++stage == NS) { stage = 0; wphase ^= 1u; }
            } else if (RES) {
              // fp16 pairs of 32 K values in the chunk's first 16 columns = 2 slices (1 in the last chunk).  Layer 2
              // and the first NRES1 chunks of layers 1 / 3 read resident weights, the other chunks the shared slots
              // (layer 1's fill is the slot barrier's even phase, layer 3's the odd one).
              const bool slot = c >= SM::NRES1 && l != 2;
              tc_fence_after();
              if (leader) {
                const uint32_t lo = slot ? ring_lo + (uint32_t)(c - SM::NRES1) * STAGE_D : res_l + (uint32_t)c * STAGE_D;
                mma_ts_bf16(tD, tA + c * 32, make_desc64(lo), IDESC_F16, c ? 1u : 0u);
                if (c != C::NCHUNK - 1) mma_ts_bf16(tD, tA + c * 32 + 8, make_desc64(lo + SLICE_D), IDESC_F16, 1u);
              }
            } else if (F16) {
              // chunk c == ring stage: fp16 pairs of 32 K values in the chunk's first 16 columns = 2 slices (1 in the last)
              tc_fence_after();
              if (leader) {
                const uint32_t lo = ring_lo + stage * STAGE_D;
                mma_ts_bf16(tD, tA + c * 32, make_desc64(lo), IDESC_F16, c ? 1u : 0u);
                if (c != C::NCHUNK - 1) mma_ts_bf16(tD, tA + c * 32 + 8, make_desc64(lo + SLICE_D), IDESC_F16, 1u);
                tc_commit(BAR_WEMPTY(stage));
              }
              if (++stage == NS) { stage = 0; wphase ^= 1u; }
            } else {
              // chunk c holds [hi(16 cols) | lo(16 cols)] of 32 K values = 2 slices (1 in the last chunk); slice == stage
              constexpr int LASTC = C::NCHUNK - 1;
#pragma unroll
              for (int sc = 0; sc < 2; ++sc) {
                if (c == LASTC && sc == 1) break;
                mbar_wait(BAR_WFULL(stage), wphase, s_abort);
                tc_fence_after();
                if (leader) {
                  const uint32_t lo = ring_lo + stage * STAGE_D;
                  const uint32_t a_hi = tA + c * 32 + sc * 8;
                  const uint32_t a_lo = a_hi + (c == LASTC ? 8 : 16);
                  mma_ts_bf16(tD, a_hi, make_desc64(lo), IDESC_BF16, (c | sc) ? 1u : 0u);
                  mma_ts_bf16(tD, a_lo, make_desc64(lo), IDESC_BF16, 1u);
                  mma_ts_bf16(tD, a_hi, make_desc64(lo + SLICE_D), IDESC_BF16, 1u);
                  tc_commit(BAR_WEMPTY(stage));
                }
                if (++stage == NS) { stage = 0; wphase ^= 1u; }
              }
            }
          }
          aphase ^= 1u;
          if (leader) tc_commit(BAR_DFULL(l & 1));
          TC_TRACE(10 * l + 8);
        }
      }
#ifdef MBO_TC_TRACE
      if (lane == 0 && blockIdx.x < 148) {
        g_tc_trace[5120 + blockIdx.x * 8 + 0] = st_feat;
        g_tc_trace[5120 + blockIdx.x * 8 + 1] = st_a;
        g_tc_trace[5120 + blockIdx.x * 8 + 2] = st_w;
      }
#endif
    }
  } else if (warp < WARP_TMA || !FUSE) {
    // ================= epilogue warps (0-7) and feature warps (10-13, !FUSE) =================
    // epilogue warp w: TMEM lanes 32*(w&3)..+31 (thread = candidate row), column half h = w>>2:
    // half h converts chunks h, h+2, ...; FUSE: half 1 also builds the next tile's features (the GP warps feed it)
    const bool feat_warp = warp >= WARP_GP0;
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;                       // 0..127
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t dph = 0;                                // bit i = phase of d_full[i]
    uint32_t sphase = 0;
    uint32_t sat = 0u;                               // fp16 mode: largest activation bit pattern seen
    long long st_d = 0, st_build = 0;
    (void)st_d; (void)st_build;
    const float* s_w4 = s_tail + 4 * TC_N;          // s_tail = [zeros | b1 | b2 | b3 | w4 | b4]

    // Feature vector of this thread's row for tile `t_idx` (fp32, ones column appended).  It is gathered one tile
    // AHEAD into registers (`pf`) so that the global-memory latency of mean/std/candidates never sits between the
    // last hidden-layer epilogue and the head epilogue (half-0 warps wait for half-1 at the pair barrier there).
    constexpr int KF = 8 * K0S;                        // feature slots (incl. the ones column)
    // All loops here are rolled and `f` is indexed at run time (local memory): this code runs once per tile on warps
    // with plenty of slack, and its size is what matters (the unrolled version was 1 400 of the kernel's 4 700
    // instructions and streamed through the instruction cache on every tile).
    auto gather_features = [&](int t_idx, float* f) {
      const int tile = (int)blockIdx.x + t_idx * (int)gridDim.x;
      const int b = tile / tiles_per_episode;
      const int m = (tile - b * tiles_per_episode) * TC_M + row;
#pragma unroll 1
      for (int i = 0; i < KF; ++i) f[i] = 0.f;
      if (t_idx >= my_tiles || m >= cand.M) return;
      if (fs.dense) {
        const float* srow = fs.dense + ((size_t)b * cand.M + m) * fs.dense_F;
#pragma unroll 1
        for (int i = 0; i < fs.F; ++i) f[i] = srow[fs.src[i]];
      } else if (fs.standard && !FUSE) {
        // get_state's own column order: predicated, branch-free, registers only (this is a serial chain on one warp)
        double x[MBO_MAX_D];
        load_candidate_compact(cand, b, m, x);
        const float mu = mean[(size_t)b * cand.M + m], sd = std_[(size_t)b * cand.M + m];
        const int nx = 2 + cand.D;
#pragma unroll
        for (int i = 0; i < KF - 1; ++i) {
          float v = 0.f;
          if (i == 0) v = mu;
          else if (i == 1) v = sd;
          else {
            if (i - 2 < MBO_MAX_D && i < nx) v = (float)x[i - 2 < MBO_MAX_D ? i - 2 : 0];
            if (i >= nx && i < fs.F) v = epi[b * 4 + (fs.src[i] - MBO_FEAT_INCUMBENT)];
          }
          f[i] = v;
        }
        {
          // mean / std are streamed from HBM exactly once: pull the next tile's lines into L2 now
          const int tile_n = tile + (int)gridDim.x;
          const int bn = tile_n / tiles_per_episode;
          const int mn = (tile_n - bn * tiles_per_episode) * TC_M + row;
          if (bn < B && mn < cand.M && (lane & 7) == 0) {       // one prefetch per 32-byte sector
            asm volatile("prefetch.global.L2 [%0];" ::"l"(mean + (size_t)bn * cand.M + mn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(std_ + (size_t)bn * cand.M + mn));
          }
        }
      } else {
        // every load is issued before the first use (one memory round trip), then the policy's columns are picked
        // from a table indexed by the MBO_FEAT_* code
        double x[MBO_MAX_D];
        load_candidate_compact(cand, b, m, x);
        float mu = 0.f, sd = 0.f;
        if (!FUSE) {
          mu = mean[(size_t)b * cand.M + m];
          sd = std_[(size_t)b * cand.M + m];
          // mean / std are streamed from HBM exactly once: pull the next tile's lines into L2 now, so that the next
          // gather pays an L2 hit instead of a DRAM round trip (the gather is a serial chain on four warps and must
          // stay well below a tile time)
          const int tile_n = tile + (int)gridDim.x;
          const int bn = tile_n / tiles_per_episode;
          const int mn = (tile_n - bn * tiles_per_episode) * TC_M + row;
          if (bn < B && mn < cand.M && (lane & 7) == 0) {       // one prefetch per 32-byte sector
            asm volatile("prefetch.global.L2 [%0];" ::"l"(mean + (size_t)bn * cand.M + mn));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(std_ + (size_t)bn * cand.M + mn));
          }
        }
        const float4 e4 = make_float4(epi[b * 4 + 0], epi[b * 4 + 1], epi[b * 4 + 2], epi[b * 4 + 3]);
        float sv[MBO_FEAT_BUDGET + 1];
        sv[MBO_FEAT_MEAN] = mu;
        sv[MBO_FEAT_STD] = sd;
#pragma unroll
        for (int d = 0; d < MBO_MAX_D; ++d) sv[MBO_FEAT_X0 + d] = (float)x[d];
        sv[MBO_FEAT_INCUMBENT] = e4.x; sv[MBO_FEAT_INCUMBENT + 1] = e4.y;
        sv[MBO_FEAT_INCUMBENT + 2] = e4.z; sv[MBO_FEAT_INCUMBENT + 3] = e4.w;
#pragma unroll 1
        for (int i = 0; i < fs.F; ++i) f[i] = sv[fs.src[i]];
      }
      f[fs.F] = 1.0f;    // bias column
    };
    float pf[KF];                                   // prefetched features of the next tile to be built

    auto build_features = [&](int t_idx) {
      if (FUSE) {   // posterior of this tile from the GP warps
        mbar_wait(BAR_GPFULL(t_idx & 1), (uint32_t)(t_idx >> 1) & 1u, s_abort);
        const int tile = (int)blockIdx.x + t_idx * (int)gridDim.x;
        const int b = tile / tiles_per_episode;
        const int m = (tile - b * tiles_per_episode) * TC_M + row;
        if (m < cand.M && !fs.dense) {
#pragma unroll
          for (int i = 0; i < KF - 1; ++i) {
            if (i < fs.F && fs.src[i] == MBO_FEAT_MEAN) pf[i] = s_mu[(t_idx & 1) * TC_M + row];
            if (i < fs.F && fs.src[i] == MBO_FEAT_STD) pf[i] = s_sd[(t_idx & 1) * TC_M + row];
          }
        }
      }
      uint32_t hi[KF], lo[KF];
#pragma unroll
      for (int i = 0; i < KF; ++i) {
        float h = tf32_rna(pf[i]);
        hi[i] = __float_as_uint(h);
        lo[i] = __float_as_uint(tf32_rna(pf[i] - h));
      }
      if (FUSE) mbar_arrive(BAR_GPEMPTY(t_idx & 1));   // mu/sd of this buffer consumed
      if (K0S == 1) {
        tmem_st8(tlane + COL_S, hi);
        tmem_st8(tlane + COL_S + 16, lo);
      } else {
        tmem_st16(tlane + COL_S, hi);
        tmem_st16(tlane + COL_S + 16, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_FEAT);
      if (FUSE) gather_features(t_idx + 1, pf);
    };

    // relu(D + b) -> next layer's A operand, in place, for chunks c0, c0 + step, ... < cend of the region at `col`
    // (loads double-buffered, loop rolled over pairs of chunks); every finished chunk is handed to the MMA warp
    // through its own mbarrier.
    constexpr int MAXI = (C::NCHUNK + 1) / 2;
    // STORE: a finished block of <= 32 activation columns goes through this warp's staging tile so that global memory
    // sees whole 128-byte row segments (thread = row on the TMEM side, 8 lanes = one row segment on the HBM side; a
    // thread storing its own row would touch 32 different lines per instruction: measured 150 instead of 53 us per call)
    float* const stg = (float*)(smem + SM::off_stg) + (STORE ? warp * 32 * STG_LD : 0);
    float* st_base = nullptr;           // first row of this warp's 32 rows in the activation being stored
    uint32_t* st_bits = nullptr;        // this thread's row of ReLU bit words (null: padding row / no words for this layer)
    int st_nrows = 0;                   // rows of this warp that exist (ragged last tile of an episode)
    auto st_tile = [&](int t, float* h, uint32_t* bits) {
      const int tile_s = (int)blockIdx.x + t * (int)gridDim.x;
      const int b_s = tile_s / tiles_per_episode;
      const int m0 = (tile_s - b_s * tiles_per_episode) * TC_M + quad * 32;
      const size_t r0 = (size_t)b_s * cand.M + m0;
      st_nrows = min(max(cand.M - m0, 0), 32);
      st_base = h + r0 * TC_N;
      st_bits = (bits && lane < st_nrows) ? bits + (r0 + lane) * 8 : nullptr;
    };
    auto st_block = [&](auto&& get, int n4, int gcol) {      // get(j) = columns gcol + 4 j .. of this thread's row
      float4* srow = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < n4) {
          float4 v = get(j);
          const int col0 = gcol + 4 * j;
          const int d = st.H - col0;                                            // the ones column
          v.x = d == 0 ? 1.0f : v.x; v.y = d == 1 ? 1.0f : v.y; v.z = d == 2 ? 1.0f : v.z; v.w = d == 3 ? 1.0f : v.w;
          srow[j] = v;
        }
      }
      __syncwarp();
      const int ch = lane & 7, rr = lane >> 3;
      if (ch < n4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rr;
          if (r < st_nrows)
            *reinterpret_cast<float4*>(st_base + (size_t)r * TC_N + gcol + 4 * ch) = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * ch);
        }
      }
      __syncwarp();
    };
    auto convert_chunks = [&](auto rolled, uint32_t col, const float* bias, int c0, int step, int cend, int tt, int trace_slot) {
      uint32_t va[C::CHUNK_COLS], vb[C::CHUNK_COLS];
      auto ncols_of = [&](int c) { return (MODE == MBO_MLP_TF32) ? 40 : (c == C::NCHUNK - 1 ? 16 : 32); };
      auto one = [&](int c, uint32_t* cur, uint32_t* nxt, int i) {
        tmem_wait_ld();
        if (c + step < cend) ChunkOps<MODE>::load(tlane + col + (c + step) * C::CHUNK_COLS, nxt, ncols_of(c + step));
        if constexpr (F16) ChunkOps<MODE>::convert_store(tlane + col + c * C::CHUNK_COLS, cur, ncols_of(c), sat);
        else ChunkOps<MODE>::convert_store(tlane + col + c * C::CHUNK_COLS, cur, bias + c * C::CHUNK_COLS, ncols_of(c));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR_AREADY(c));
        if (quad == 0) TC_TRACE_T(tt, trace_slot + i);
        if constexpr (STORE) {
          // after the hand-off (off the MMA's critical path): relu(D + b) of this chunk as fp32 + its ReLU bit word
          const int nc = ncols_of(c);
          const float* bb = bias + c * C::CHUNK_COLS;
          uint32_t word = 0u;
          st_block([&](int j) {
            float4 o;
            o.x = fmaxf(__uint_as_float(cur[4 * j + 0]) + bb[4 * j + 0], 0.f);
            o.y = fmaxf(__uint_as_float(cur[4 * j + 1]) + bb[4 * j + 1], 0.f);
            o.z = fmaxf(__uint_as_float(cur[4 * j + 2]) + bb[4 * j + 2], 0.f);
            o.w = fmaxf(__uint_as_float(cur[4 * j + 3]) + bb[4 * j + 3], 0.f);
            word |= (o.x > 0.f ? 1u : 0u) << (4 * j) | (o.y > 0.f ? 1u : 0u) << (4 * j + 1) |
                    (o.z > 0.f ? 1u : 0u) << (4 * j + 2) | (o.w > 0.f ? 1u : 0u) << (4 * j + 3);
            return o;
          }, nc / 4, c * C::CHUNK_COLS);
          if (st_bits) st_bits[c] = word;
        }
      };
      if (c0 < cend) ChunkOps<MODE>::load(tlane + col + c0 * C::CHUNK_COLS, va, ncols_of(c0));
      if constexpr (decltype(rolled)::value) {
#pragma unroll 1
        for (int i = 0, c = c0; c < cend; i += 2, c += 2 * step) {
          one(c, va, vb, i);
          if (c + step < cend) one(c + step, vb, va, i + 1);
        }
      } else {       // at most MAXI chunks, unrolled (the epilogue warps' copy: its first chunk is on the critical path)
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
          const int c = c0 + step * i;
          if (c < cend) one(c, (i & 1) ? vb : va, (i & 1) ? va : vb, i);
        }
      }
    };

    if (feat_warp) {
      // ---- feature warps: S(t) may be overwritten once L0(t-1) has been committed (s_free) ----
#pragma unroll 1
      for (int t = -2; t < my_tiles; ++t) {
        if (t >= 0) {
          // L0(t) committed: D0 is complete in R0 and S may be overwritten
          mbar_wait(BAR_SFREE, sphase, s_abort);
          sphase ^= 1u;
          tc_fence_after();
          if (quad == 0) TC_TRACE(300);
          if (E0_BY_FEAT) convert_chunks(std::true_type{}, COL_R0, s_tail, 0, 1, C::NCHUNK, t, 310);
        }
        {
          TC_STAT_BEGIN();
          if (t >= -1 && t + 1 < my_tiles) build_features(t + 1);
          if (t >= 0 && quad == 0) TC_TRACE(302);
          gather_features(t + 2, pf);                  // loads land while this warp waits for the next s_free
          TC_STAT_END(st_build);
        }
        if (t >= 0 && quad == 0) TC_TRACE(301);
      }
    }
    if (FUSE && my_tiles > 0 && half == 1) { gather_features(0, pf); build_features(0); }
    for (int t = 0; t < my_tiles && !feat_warp; ++t) {
      // ---- E0..E2: relu(D + b) -> next layer's A operand, in place ----
#pragma unroll 1
      for (int l = E0_BY_FEAT ? 1 : 0; l < 3; ++l) {
        const uint32_t col = (l & 1) ? COL_R1 : COL_R0;
        {
          TC_STAT_BEGIN();
          mbar_wait(BAR_DFULL(l & 1), (dph >> (l & 1)) & 1u, s_abort);
          TC_STAT_END(st_d);
        }
        dph ^= 1u << (l & 1);
        tc_fence_after();
        if (quad == 0) TC_TRACE(100 + 100 * half + 20 * l);
        const float* bias = s_tail + l * TC_N;
        if constexpr (STORE) st_tile(t, st.h[l], st.bits[l]);
        convert_chunks(std::false_type{}, col, bias, half, 2, C::NCHUNK, t, 100 + 100 * half + 20 * l + 1);
        // Layer 1 is complete (this warp saw d_full): the shared slots take layer 3's stages.  Issuing a bulk copy
        // costs the thread ~160 cycles, so this sits after the chunk conversions, on the half with fewer chunks;
        // the copies land ~2 000 cycles before layer 3 needs them.
        if (RES && l == 1 && tid == 4 * 32) refill_slots(2);
        if (FUSE && l == 2 && half == 1) {
          if (t + 1 < my_tiles) {
            // features of the next tile while L3 runs
            mbar_wait(BAR_SFREE, sphase, s_abort);
            tc_fence_after();
            if (quad == 0) TC_TRACE(100 + 100 * half + 80);
            build_features(t + 1);
            if (quad == 0) TC_TRACE(100 + 100 * half + 81);
          }
          sphase ^= 1u;
        }
      }
      // ---- E3: logit = relu(D + b3) . w4 + b4, columns split between the two halves ----
      {
        TC_STAT_BEGIN();
        mbar_wait(BAR_DFULL(1), (dph >> 1) & 1u, s_abort);
        TC_STAT_END(st_d);
      }
      dph ^= 2u;
      tc_fence_after();
      if (quad == 0) TC_TRACE(100 + 100 * half + 70);
      float acc = 0.f, acc2 = 0.f;       // two independent chains
      {
        // 104 columns per thread as three 32-column blocks and one 8-column block (rolled: this code runs once per
        // tile, and its size matters more than the ~100 cycles of load latency the rolling exposes)
        const int cb = half * 104;
        const float* b3 = s_tail + 3 * TC_N + cb;
        const float* w4 = s_w4 + cb;
        uint32_t v0[32], v1[32];
        if constexpr (STORE) st_tile(t, st.h[3], nullptr);
        auto dot = [&](const uint32_t* v, int j0, int n4) {     // columns j0 .. j0 + 4 n4 of this thread's range
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < n4) {
              const float4 w = reinterpret_cast<const float4*>(w4 + j0)[j];
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (!F16) b = reinterpret_cast<const float4*>(b3 + j0)[j];   // fp16 mode: b3 is in the accumulator
              const float x0 = fmaxf(__uint_as_float(v[4 * j + 0]) + b.x, 0.f), x1 = fmaxf(__uint_as_float(v[4 * j + 1]) + b.y, 0.f);
              const float x2 = fmaxf(__uint_as_float(v[4 * j + 2]) + b.z, 0.f), x3 = fmaxf(__uint_as_float(v[4 * j + 3]) + b.w, 0.f);
              acc = fmaf(x0, w.x, acc);
              acc2 = fmaf(x1, w.y, acc2);
              acc = fmaf(x2, w.z, acc);
              acc2 = fmaf(x3, w.w, acc2);
            }
          }
          if constexpr (STORE) {
            st_block([&](int j) {
              float4 b = reinterpret_cast<const float4*>(b3 + j0)[j];
              return make_float4(fmaxf(__uint_as_float(v[4 * j + 0]) + b.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + b.y, 0.f),
                                 fmaxf(__uint_as_float(v[4 * j + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + b.w, 0.f));
            }, n4, cb + j0);
          }
        };
        // R1 is what the next tile's layer 1 waits for (r1_free): get D3 into registers with two load round trips instead
        // of four -- two 32-column loads in flight, the third / fourth issued while the first two blocks are reduced
        tmem_ld32(tlane + COL_R1 + cb, v0);
        tmem_ld32(tlane + COL_R1 + cb + 32, v1);
        tmem_wait_ld();
        dot(v0, 0, 8);
        tmem_ld32(tlane + COL_R1 + cb + 64, v0);
        dot(v1, 32, 8);
        tmem_ld8(tlane + COL_R1 + cb + 96, v1);
        tmem_wait_ld();
        tc_fence_before();
        if (lane == 0) mbar_arrive(BAR_R1FREE);     // D3 is in registers: layer 1 of the next tile may overwrite R1
        dot(v0, 64, 8);
        dot(v1, 96, 2);
        acc += acc2;
      }
      tc_fence_before();   // order this tile's TMEM reads before the next hand-offs
      if (quad == 0) TC_TRACE(100 + 100 * half + 71);
      float lgt = 0.f;
      int tile_m = 0x7fffffff, tile_id = 0;
      if (half == 1) s_part[row] = acc;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      if (half == 0) {
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int b = tile / tiles_per_episode;
        const int m = (tile - b * tiles_per_episode) * TC_M + row;
        lgt = acc + s_part[row] + s_tail[5 * TC_N];
        tile_m = m;
        tile_id = tile;
        if (logits && m < cand.M) logits[(size_t)b * cand.M + m] = lgt;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // s_part reusable
      if (quad == 0) TC_TRACE(100 + 100 * half + 72);
      if (tk.k && half == 0) {
        // warp top-k over this warp's 32 rows (value descending, index ascending; NaN / padding never win), off the
        // pair barrier's critical path; two REDUX per round on an order-preserving integer key
        const bool valid = tile_m < cand.M && lgt == lgt;
        const uint32_t u = __float_as_uint(lgt + 0.0f);                                 // -0 -> +0 (compare equal as floats)
        uint32_t key = valid ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0u;     // 0 < every valid key
        int mi = valid ? tile_m : 0x7fffffff;
        uint32_t outk = 0u;
        int outm = 0x7fffffff;
#pragma unroll 1
        for (int r = 0; r < tk.k; ++r) {
          const uint32_t bk = __reduce_max_sync(0xffffffffu, key);
          const int bm = __reduce_min_sync(0xffffffffu, key == bk ? mi : 0x7fffffff);
          if (lane == r) { outk = bk; outm = bm; }
          if (mi == bm) { key = 0u; mi = 0x7fffffff; }
        }
        if (lane < tk.k) {
          const size_t p = ((size_t)tile_id * 4 + quad) * 8 + lane;
          const uint32_t uo = (outk & 0x80000000u) ? (outk & 0x7fffffffu) : ~outk;
          tk.part_val[p] = outk ? __uint_as_float(uo) : -INFINITY;
          tk.part_idx[p] = outk ? outm : 0x7fffffff;
        }
      }
    }
#ifdef MBO_TC_TRACE
    if (lane == 0 && blockIdx.x < 148 && warp == 0) g_tc_trace[5120 + blockIdx.x * 8 + 3] = st_d;
    if (lane == 0 && blockIdx.x < 148 && warp == WARP_GP0) g_tc_trace[5120 + blockIdx.x * 8 + 4] = st_build;
#endif
    // fp16 mode: an activation at the range limit (65504, or NaN) was clamped somewhere -> bit 1 of the flag
    if (F16 && ((sat & 0xffffu) >= 0x7bffu || (sat >> 16) >= 0x7bffu)) atomicOr(err_flag, 2);
  } else if (FUSE) {
    // ================= fused GP posterior warps (fp64): thread = candidate row of the NEXT tile =================
    // mean = v.beta + m, var = max(k** - |v|^2, 1e-15), v = Linv k*  (same arithmetic as gp_posterior_kernel<double>),
    // k* in registers (n <= 32, loops fully unrolled), episode blob in shared memory.
    const int g = tid - WARP_GP0 * 32;
    double* blob = (double*)(smem + SM::off_gp);
    double* ks = (double*)(smem + SM::off_ks);
    const int D = cand.D;
    const int NPmax = gp_np(gp.nmax);
    const size_t gstride = gp_state_doubles(gp.nmax, D);
    int cur_b = -1;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(BAR_GPEMPTY(buf), ((uint32_t)(t >> 1) & 1u) ^ 1u, s_abort);
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int b = tile / tiles_per_episode;
      const int m = (tile - b * tiles_per_episode) * TC_M + g;
      if (b != cur_b) {
        asm volatile("bar.sync 5, 128;" ::: "memory");           // everyone done with the previous episode's blob
        const double* st = gp.gp_state + (size_t)b * gstride;
        const int n_ = (int)st[0];
        const int np_ = gp_np(n_);
        for (int i = g; i < 16; i += 128) blob[i] = st[i];
        for (int i = g; i < np_ * D; i += 128) blob[GPB_XS + i] = st[gp_off_xs() + i];
        for (int i = g; i < GP_NP; i += 128) blob[GPB_BETA + i] = i < np_ ? st[gp_off_beta(NPmax, D) + i] : 0.0;
        const int lc = gp_linv_doubles(np_);
        for (int i = g; i < lc; i += 128) blob[GPB_LINV + i] = st[gp_off_linv(NPmax, D) + i];
        asm volatile("bar.sync 5, 128;" ::: "memory");
        cur_b = b;
      }
      const int n = (int)blob[0];
      const double mprior = blob[1], var = blob[2];
      double mean_v = 0.0, var_v = var;
      if (n > 0 && m < cand.M) {
        double x[MBO_MAX_D];
#pragma unroll
        for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
        load_candidate(cand, b, m, x);
#pragma unroll
        for (int d = 0; d < MBO_MAX_D; ++d) x[d] = d < D ? x[d] / blob[4 + d] : 0.0;
        // k* into this thread's private shared-memory column (rolled loops: the fused role must stay small
        // enough for the instruction cache; a fully unrolled register version thrashed it)
        double* kcol = ks + g;
        for (int j0 = 0; j0 < n; j0 += 8) {          // 8 interleaved exp chains per step
          double r2v[8], kv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int jj = min(j0 + u, n - 1);
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < MBO_MAX_D; ++d)
              if (d < D) { const double tt = x[d] - blob[GPB_XS + jj * D + d]; r2 = fma(tt, tt, r2); }
            r2v[u] = r2;
          }
          kernel_of_r2_xn<8>(gp.kid, var, r2v, kv);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (j0 + u < n) kcol[(j0 + u) * TC_M] = kv[u];
        }
        for (int j = n; j < gp_np(n); ++j) kcol[j * TC_M] = 0.0;
        double q = 0.0, mu = 0.0;
        // v = Linv k*, two 8-row blocks per pass (16 independent DFMA chains); LinvB rows beyond n are zero
        for (int ib = 0; ib < n; ib += 16) {
          double acc0[8], acc1[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { acc0[r] = 0.0; acc1[r] = 0.0; }
          const double* blk0 = blob + GPB_LINV + gp_linv_block_off(ib);
          const double* blk1 = blob + GPB_LINV + gp_linv_block_off(ib + 8);
          const bool two = ib + 8 < gp_np(n);
#pragma unroll 2
          for (int j = 0; j < ib + 8; ++j) {
            const double kj = kcol[j * TC_M];
            const double2* p = (const double2*)(blk0 + j * 8);
            const double2 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
            acc0[0] = fma(a0.x, kj, acc0[0]); acc0[1] = fma(a0.y, kj, acc0[1]);
            acc0[2] = fma(a1.x, kj, acc0[2]); acc0[3] = fma(a1.y, kj, acc0[3]);
            acc0[4] = fma(a2.x, kj, acc0[4]); acc0[5] = fma(a2.y, kj, acc0[5]);
            acc0[6] = fma(a3.x, kj, acc0[6]); acc0[7] = fma(a3.y, kj, acc0[7]);
            if (two) {
              const double2* pp = (const double2*)(blk1 + j * 8);
              const double2 b0 = pp[0], b1 = pp[1], b2 = pp[2], b3 = pp[3];
              acc1[0] = fma(b0.x, kj, acc1[0]); acc1[1] = fma(b0.y, kj, acc1[1]);
              acc1[2] = fma(b1.x, kj, acc1[2]); acc1[3] = fma(b1.y, kj, acc1[3]);
              acc1[4] = fma(b2.x, kj, acc1[4]); acc1[5] = fma(b2.y, kj, acc1[5]);
              acc1[6] = fma(b3.x, kj, acc1[6]); acc1[7] = fma(b3.y, kj, acc1[7]);
            }
          }
          if (two) {
#pragma unroll 2
            for (int j = ib + 8; j < ib + 16; ++j) {
              const double kj = kcol[j * TC_M];
              const double2* pp = (const double2*)(blk1 + j * 8);
              const double2 b0 = pp[0], b1 = pp[1], b2 = pp[2], b3 = pp[3];
              acc1[0] = fma(b0.x, kj, acc1[0]); acc1[1] = fma(b0.y, kj, acc1[1]);
              acc1[2] = fma(b1.x, kj, acc1[2]); acc1[3] = fma(b1.y, kj, acc1[3]);
              acc1[4] = fma(b2.x, kj, acc1[4]); acc1[5] = fma(b2.y, kj, acc1[5]);
              acc1[6] = fma(b3.x, kj, acc1[6]); acc1[7] = fma(b3.y, kj, acc1[7]);
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            q = fma(acc0[r], acc0[r], q);
            mu = fma(acc0[r], blob[GPB_BETA + ib + r], mu);
          }
          if (two) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              q = fma(acc1[r], acc1[r], q);
              mu = fma(acc1[r], blob[GPB_BETA + ib + 8 + r], mu);
            }
          }
        }
        mean_v = mu + mprior;
        var_v = var - q;
        var_v = var_v < 1e-15 ? 1e-15 : var_v;
      }
      const float fm = (float)mean_v, fsd = (float)sqrt(var_v);
      s_mu[buf * TC_M + g] = fm;
      s_sd[buf * TC_M + g] = fsd;
      if (gp.mean_out && m < cand.M) {
        gp.mean_out[(size_t)b * cand.M + m] = fm;
        gp.std_out[(size_t)b * cand.M + m] = fsd;
      }
      mbar_arrive(BAR_GPFULL(buf));      // release: orders the shared-memory stores above
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#ifdef MBO_TC_TRACE
  if (tid == 0 && blockIdx.x < 148) {     // per-CTA totals: cycles, nanoseconds, tiles
    g_tc_trace[4096 + blockIdx.x * 3 + 0] = clock64() - trace_c0;
    g_tc_trace[4096 + blockIdx.x * 3 + 1] = (long long)(globaltimer_ns() - trace_g0);
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tc_trace[4096 + blockIdx.x * 3 + 2] = my_tiles + ((long long)smid << 20);
  }
#endif
  if (tid == 0 && *s_abort_p) atomicOr(err_flag, 1);
  if (F16 && tid == 0 && ((const int*)(pack + pl.off_scale))[4]) atomicOr(err_flag, 2);   // a packed weight/bias was clamped
  if (warp == WARP_TMA) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// per episode: merge the 4 * tiles_per_episode warp partial lists into the episode's top-k
__global__ void __launch_bounds__(128)
topk_merge_kernel(int n_lists, int k, const float* __restrict__ part_val, const int32_t* __restrict__ part_idx,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ float sv[4];
  __shared__ int si[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* pv = part_val + (size_t)b * n_lists * 8;
  const int32_t* pi = part_idx + (size_t)b * n_lists * 8;
  float prev_v = INFINITY;
  int prev_i = -1;
  for (int r = 0; r < k; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int e = tid; e < n_lists * k; e += 128) {
      const int l = e / k, j = e - l * k;
      const float v = pv[l * 8 + j];
      const int i = pi[l * 8 + j];
      const bool after_prev = (v < prev_v) || (v == prev_v && i > prev_i);
      if (i != 0x7fffffff && after_prev && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { sv[w] = bv; si[w] = bi; }
    __syncthreads();
    bv = sv[0]; bi = si[0];
#pragma unroll
    for (int q = 1; q < 4; ++q)
      if (sv[q] > bv || (sv[q] == bv && si[q] < bi)) { bv = sv[q]; bi = si[q]; }
    if (tid == 0) {
      out_idx[(size_t)b * k + r] = bi == 0x7fffffff ? 0 : bi;
      if (out_val) out_val[(size_t)b * k + r] = bv;
    }
    prev_v = bv;
    prev_i = bi;
  }
}

static int g_num_sms = 0;

// [MEAN, STD, X0 .. X(D-1), per-episode scalar codes ...]: the column order get_state produces
static bool tc_standard_layout(int F, const int32_t* src, int D) {
  if (F < 2 + D || src[0] != MBO_FEAT_MEAN || src[1] != MBO_FEAT_STD) return false;
  for (int d = 0; d < D; ++d)
    if (src[2 + d] != MBO_FEAT_X0 + d) return false;
  for (int i = 2 + D; i < F; ++i)
    if (src[i] < MBO_FEAT_INCUMBENT || src[i] > MBO_FEAT_BUDGET) return false;
  return true;
}

template <int MODE, bool FUSE, int K0S, bool STORE = false>
static int launch_tc_pack(const unsigned char* pack, int B, const TcFeat& fs, const mbo_candidates& cand, const float* mean,
                          const float* std_, const float* epi, float* logits, int* err_flag, cudaStream_t s,
                          const GpFuse& gp, const TcTopk& tk, const TcStore& st);

template <int MODE, bool FUSE, int K0S>
static int launch_tc_k(const mbo_policy* p, int B, const TcFeat& fs, const mbo_candidates& cand, const float* mean,
                     const float* std_, const float* epi, float* logits, int* err_flag, cudaStream_t s,
                     const GpFuse& gp, const TcTopk& tk) {
  TcStore st;
  memset(&st, 0, sizeof(st));
  return launch_tc_pack<MODE, FUSE, K0S, false>((const unsigned char*)p->d_tc, B, fs, cand, mean, std_, epi, logits, err_flag, s, gp, tk, st);
}

template <int MODE, bool FUSE, int K0S, bool STORE>
static int launch_tc_pack(const unsigned char* pack, int B, const TcFeat& fs, const mbo_candidates& cand, const float* mean,
                          const float* std_, const float* epi, float* logits, int* err_flag, cudaStream_t s,
                          const GpFuse& gp, const TcTopk& tk, const TcStore& st) {
  using SM = TcSmem<MODE, FUSE, K0S, STORE>;
  auto kern = neural_af_tc_kernel<MODE, FUSE, K0S, STORE>;
  MBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::total));
  if (g_num_sms == 0) {
    int dev = 0;
    MBO_CUDA(cudaGetDevice(&dev));
    MBO_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_per_episode = (cand.M + TC_M - 1) / TC_M;
  const int total = B * tiles_per_episode;
  const int grid = total < g_num_sms ? total : g_num_sms;
  kern<<<grid, FUSE ? TC_THREADS_FUSED : TC_THREADS, SM::total, s>>>(
      pack, fs, cand, B, tiles_per_episode, mean, std_, epi, logits, err_flag,
      getenv("MBO_TC_DEBUG_NOLOAD") ? 1 : 0, gp, tk, st);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

template <int MODE, bool FUSE>
static int launch_tc(const mbo_policy* p, int B, const TcFeat& fs, const mbo_candidates& cand, const float* mean,
                     const float* std_, const float* epi, float* logits, int* err_flag, cudaStream_t s,
                     const GpFuse& gp, const TcTopk& tk) {
  if (fs.k0_slices == 1) return launch_tc_k<MODE, FUSE, 1>(p, B, fs, cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  return launch_tc_k<MODE, FUSE, 2>(p, B, fs, cand, mean, std_, epi, logits, err_flag, s, gp, tk);
}

int tc_forward(const mbo_policy* p, int mode, int B, int F, const int32_t* feat_src, const mbo_candidates* cand,
               const float* mean, const float* std_, const float* epi, float* logits, void* ws, size_t ws_bytes,
               cudaStream_t s, const float* dense, int dense_F, int topk, int32_t* topk_idx, float* topk_val) {
  if (!p->tc_ready) {
    set_error("tensor-core NeuralAF path needs a F->H->H->H->H->1 ReLU policy with F <= %d and H <= %d", TC_K0 - 1, TC_N);
    return MBO_ERR_UNSUPPORTED;
  }
  MBO_CHECK_ARG(ws && ws_bytes >= sizeof(int), "mbo_af_logits: tensor-core modes need a workspace of mbo_af_workspace_bytes()");
  TcFeat fs;
  fs.F = F;
  fs.dense = dense;
  fs.dense_F = dense_F;
  fs.k0_slices = (F + 1 <= 8) ? 1 : 2;
  for (int i = 0; i < TC_K0; ++i) fs.src[i] = i < F ? feat_src[i] : 0;
  fs.standard = !fs.dense && tc_standard_layout(F, feat_src, cand->D);
  int* err_flag = (int*)ws;
  GpFuse gp;
  memset(&gp, 0, sizeof(gp));
  TcTopk tk;
  memset(&tk, 0, sizeof(tk));
  const int tiles = (cand->M + TC_M - 1) / TC_M;
  const size_t n_part = (size_t)B * tiles * 4 * 8;
  if (topk) {
    MBO_CHECK_ARG(topk >= 1 && topk <= 8 && topk <= cand->M && topk_idx, "mbo_af_topk: k must be in 1..min(8, M)");
    MBO_CHECK_ARG(ws_bytes >= 256 + n_part * 8, "mbo_af_topk: workspace smaller than mbo_af_workspace_bytes()");
    tk.k = topk;
    tk.part_val = (float*)((char*)ws + 256);
    tk.part_idx = (int32_t*)((char*)ws + 256 + n_part * 4);
  }
  int rc;
  if (mode == MBO_MLP_FP16 && !p->tc_f16) {
    set_error("MBO_MLP_FP16 needs H <= %d (two spare K columns carry the bias); use MBO_MLP_TF32", TC_N - 2);
    return MBO_ERR_UNSUPPORTED;
  }
  if (mode == MBO_MLP_TF32)
    rc = launch_tc<MBO_MLP_TF32, false>(p, B, fs, *cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  else if (mode == MBO_MLP_FP16)
    rc = launch_tc<MBO_MLP_FP16, false>(p, B, fs, *cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  else
    rc = launch_tc<MBO_MLP_BF16X3, false>(p, B, fs, *cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  if (rc || !topk) return rc;
  topk_merge_kernel<<<B, 128, 0, s>>>(tiles * 4, topk, tk.part_val, tk.part_idx, topk_idx, topk_val);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

// ---- entry points of the PPO update (ppo_update.cu): the current raw weights, packed for the bf16x3 mode on every
// optimiser step, and the policy forward on materialised states with optional activation / ReLU-bit stores ----------
size_t tc_pack_bytes() { return pack_layout().total; }

int tc_pack_raw_bf16x3(const float* const* W, const float* const* b, int F, int H, void* pack, cudaStream_t s) {
  MBO_CHECK_ARG(F + 1 <= TC_K0 && H >= 8 && H <= TC_N, "tc_pack_raw_bf16x3: unsupported shape F=%d H=%d", F, H);
  PackArgs a;
  for (int l = 0; l < 5; ++l) { a.W[l] = W[l]; a.b[l] = b[l]; }
  a.F = F; a.H = H;
  const size_t words = pack_layout().total / 4;
  tc_pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(a, (unsigned char*)pack, 1);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

int tc_forward_dense_bf16x3(const void* pack, int E, int M, int F_state, int F, const int32_t* cols, const float* states,
                            float* logits, int* err_flag, float* const* h_out, uint32_t* const* bits_out, int H,
                            cudaStream_t s) {
  TcFeat fs;
  fs.F = F;
  fs.dense = states;
  fs.dense_F = F_state;
  fs.k0_slices = (F + 1 <= 8) ? 1 : 2;
  for (int i = 0; i < TC_K0; ++i) fs.src[i] = i < F ? cols[i] : 0;
  fs.standard = 0;
  mbo_candidates c;
  memset(&c, 0, sizeof(c));
  c.M = M;
  c.D = 1;
  GpFuse gp;
  memset(&gp, 0, sizeof(gp));
  TcTopk tk;
  memset(&tk, 0, sizeof(tk));
  TcStore st;
  memset(&st, 0, sizeof(st));
  const bool store = h_out != nullptr;
  if (store) {
    for (int l = 0; l < 4; ++l) st.h[l] = h_out[l];
    for (int l = 0; l < 3; ++l) st.bits[l] = bits_out[l];
    st.H = H;
  }
  const unsigned char* pk = (const unsigned char*)pack;
  if (store) {
    if (fs.k0_slices == 1) return launch_tc_pack<MBO_MLP_BF16X3, false, 1, true>(pk, E, fs, c, nullptr, nullptr, nullptr, logits, err_flag, s, gp, tk, st);
    return launch_tc_pack<MBO_MLP_BF16X3, false, 2, true>(pk, E, fs, c, nullptr, nullptr, nullptr, logits, err_flag, s, gp, tk, st);
  }
  if (fs.k0_slices == 1) return launch_tc_pack<MBO_MLP_BF16X3, false, 1, false>(pk, E, fs, c, nullptr, nullptr, nullptr, logits, err_flag, s, gp, tk, st);
  return launch_tc_pack<MBO_MLP_BF16X3, false, 2, false>(pk, E, fs, c, nullptr, nullptr, nullptr, logits, err_flag, s, gp, tk, st);
}

size_t tc_workspace_bytes(int B, int M) {
  const int tiles = (M + TC_M - 1) / TC_M;
  return 256 + (size_t)B * tiles * 4 * 8 * 8;
}

}  // namespace mbo

using namespace mbo;

extern "C" int mbo_af_eval_fused(const mbo_policy* p, int mode, int B, int nmax, int n_active_max, int kernel_id,
                                 const double* gp_state, int F, const int32_t* feat_src, const mbo_candidates* cand,
                                 const float* epi, float* logits, float* mean_out, float* std_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  MBO_CHECK_ARG(p && gp_state && feat_src && cand && epi && logits, "mbo_af_eval_fused: null pointer");
  MBO_CHECK_ARG(mode == MBO_MLP_TF32 || mode == MBO_MLP_BF16X3 || mode == MBO_MLP_FP16, "mbo_af_eval_fused: tensor-core modes only");
  if (!p->tc_ready) {
    set_error("mbo_af_eval_fused needs a F->H->H->H->H->1 ReLU policy with F <= %d and H <= %d", TC_K0 - 1, TC_N);
    return MBO_ERR_UNSUPPORTED;
  }
  if (n_active_max > GP_NP) {
    set_error("mbo_af_eval_fused supports n <= %d observations per episode (got %d); use mbo_gp_posterior + mbo_af_logits",
              GP_NP, n_active_max);
    return MBO_ERR_UNSUPPORTED;
  }
  MBO_CHECK_ARG(B > 0 && nmax > 0 && nmax <= 128 && n_active_max >= 0, "mbo_af_eval_fused: bad sizes");
  MBO_CHECK_ARG(kernel_id >= 0 && kernel_id <= 2, "mbo_af_eval_fused: unknown kernel id %d", kernel_id);
  MBO_CHECK_ARG(F == p->widths[0], "mbo_af_eval_fused: F=%d does not match the policy input width %d", F, p->widths[0]);
  MBO_CHECK_ARG((mean_out == nullptr) == (std_out == nullptr), "mbo_af_eval_fused: mean_out/std_out must both be set or null");
  MBO_CHECK_ARG(workspace && workspace_bytes >= sizeof(int), "mbo_af_eval_fused: workspace too small");
  int rc = check_candidates(cand);
  if (rc) return rc;
  for (int f = 0; f < F; ++f) {
    int s = feat_src[f];
    bool ok = s == MBO_FEAT_MEAN || s == MBO_FEAT_STD || (s >= MBO_FEAT_X0 && s < MBO_FEAT_X0 + cand->D) ||
              (s >= MBO_FEAT_INCUMBENT && s <= MBO_FEAT_BUDGET);
    MBO_CHECK_ARG(ok, "mbo_af_eval_fused: bad feature source %d at column %d", s, f);
  }
  TcFeat fs;
  fs.F = F;
  fs.dense = nullptr;
  fs.dense_F = 0;
  fs.k0_slices = (F + 1 <= 8) ? 1 : 2;
  for (int i = 0; i < TC_K0; ++i) fs.src[i] = i < F ? feat_src[i] : 0;
  fs.standard = tc_standard_layout(F, feat_src, cand->D);
  GpFuse gp;
  gp.gp_state = gp_state;
  gp.nmax = nmax;
  gp.kid = kernel_id;
  gp.mean_out = mean_out;
  gp.std_out = std_out;
  cudaStream_t s = (cudaStream_t)stream;
  TcTopk tk;
  memset(&tk, 0, sizeof(tk));
  if (mode == MBO_MLP_FP16 && !p->tc_f16) {
    set_error("MBO_MLP_FP16 needs H <= %d (two spare K columns carry the bias); use MBO_MLP_TF32", TC_N - 2);
    return MBO_ERR_UNSUPPORTED;
  }
  if (mode == MBO_MLP_TF32)
    return launch_tc<MBO_MLP_TF32, true>(p, B, fs, *cand, nullptr, nullptr, epi, logits, (int*)workspace, s, gp, tk);
  if (mode == MBO_MLP_FP16)
    return launch_tc<MBO_MLP_FP16, true>(p, B, fs, *cand, nullptr, nullptr, epi, logits, (int*)workspace, s, gp, tk);
  return launch_tc<MBO_MLP_BF16X3, true>(p, B, fs, *cand, nullptr, nullptr, epi, logits, (int*)workspace, s, gp, tk);
}

#ifdef MBO_TC_TRACE
extern "C" int mbo_debug_tc_trace(long long* host_out, int n) {
  MBO_CUDA(cudaDeviceSynchronize());
  MBO_CUDA(cudaMemcpyFromSymbol(host_out, g_tc_trace, sizeof(long long) * (n < 8192 ? n : 8192)));
  return MBO_OK;
}
#endif
