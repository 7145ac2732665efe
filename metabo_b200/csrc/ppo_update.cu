// ppo_update.cu -- the policy-net half of one PPO optimiser step, hand-written for sm_100a.
//
// Reference: ppo.py:129-178 (clipped surrogate + entropy bonus on one minibatch), policies.py:150-161
// (Categorical(logits): log_prob(action), entropy) and the autograd backward of mlp.py:25-61 for the
// F -> H -> H -> H -> H -> 1 ReLU policy net.  One call = forward with saved activations, loss terms,
// dlogits, dgrad chain, wgrad of every layer.  `rows` R = E episodes x M candidates (60 x 966 in C3).
//
// Data layout in HBM (workspace): activations H1..H4 are row-major [R, 208] plain fp32, stored once; the forward GEMM's
// loader threads split them in shared memory into tf32 hi + lo (3xTF32).  Pre-activation gradients dZ1..dZ4 are
// row-major [R, 208] fp32 rounded to nearest tf32 by their producers (the tensor core's truncation is then exact).
// Column H of every H_l is a ones column (bias rides inside the GEMMs: forward Wb = [W | b], wgrad column H = db),
// columns above are zero.  ReLU decisions of h2, h3 travel as bit words [R][8] from the forward to the dgrad epilogue.
//
// Kernels
//   upd_pack_kernel      W1..W3 -> UMMA-canonical K-major images [W | b] (forward) and W^T (dgrad)
//   upd_layer0_kernel    h1 = relu(x W0^T + b0)                              CUDA cores (K = F <= 8)
//   upd_ts_kernel<0>     h_{l+1} = relu([h_l,1] Wb^T) (+ head logit)         tcgen05 kind::tf32 x 3 (hi/lo), A from TMEM, B from smem
//   upd_gemm_kernel<0>   the same with A split in shared memory (SS form)    kept as a bitwise-equal variant
//   upd_loss_kernel      per episode: log-softmax, logp(a), entropy, clipped surrogate, dlogits
//   upd_dz4_kernel       dZ4 = dlogit (x) w4 . [h4 > 0]  (+ partial sums of dw4, db4)
//   upd_gemm_kernel<1> / upd_dgrad_persist_kernel   dZ_l = (dZ_{l+1} W_l) . [h_l > 0]   tcgen05 kind::tf32
//   upd_wgrad_kernel     dWb_l = dZ_{l+1}^T [h_l,1]  split-K over CTAs        tcgen05 kind::tf32, both operands MN-major
//   upd_wgrad0_kernel    dW0 = (dH1 . [h1 > 0])^T x, db0                      CUDA cores
//   upd_reduce_kernel    fixed-order sum of the split-K partials -> dW, db    (deterministic, no atomics)
//
// Operand staging: the activation / gradient tiles are row-major fp32 in HBM; 16-byte cp.async copies place each
// chunk where the no-swizzle UMMA canonical layout wants it (K-major for forward / dgrad, MN-major for wgrad where
// K = rows), completion is signalled with cp.async.mbarrier.arrive.noinc, the weight images arrive by one
// cp.async.bulk per stage.  All GEMMs here are HBM-bound (50 FLOP/B at TF32), not tensor-bound.
// Also here: the value net (v_* kernels, fp32 CUDA cores), Adam, minibatch gather / advantage normalisation, loss sums.
#include <stdlib.h>
#include "common.cuh"
#include "ppo_loss.cuh"
#include "policy.cuh"
#include "tmem_ldst.cuh"

namespace mbo {
namespace upd {

constexpr int HP = 208;             // padded width: UMMA N, row pitch of H / dZ (floats)
constexpr int TM = 128;             // rows per output tile (UMMA M)
// The GEMM pipelines are latency-bound (a stage is refilled only after its MMAs completed), so stages are small and
// many: what counts is the number of stages in flight, 3 of 4 here.
constexpr int KT = 16;              // K floats per pipeline stage of the dgrad GEMM (two K = 8 MMA steps)
constexpr int NKT = 13;             // 208 / 16
constexpr int KT3 = 8;              // forward (3xTF32: hi and lo parts of both operands share a stage; one K = 8 step)
constexpr int NKT3 = 26;            // 208 / 8
constexpr int WG_KT = 32;           // K rows per stage of the wgrad GEMM
constexpr int A_STAGE = TM * KT * 4;          // 8 192 (dgrad: 16 K values; forward: 8 K values x {hi, lo})
constexpr int B_STAGE = HP * KT * 4;          // 13 312
constexpr int IMG_FLOATS = 7 * HP * 32;       // 46 592 floats per weight image (the 32-K-per-stage dgrad image is the largest)
constexpr int STG_PITCH = 36;                 // floats per row of the epilogue's per-warp staging tile (32 + 4: conflict-free)
constexpr int GEMM_MB = 1;          // 128-row blocks per CTA.  2 (shared weight stages, 3 pipeline stages, 1 CTA/SM) measured SLOWER
                                    // (fwd 71 -> 81 us, dgrad 61 -> 94 us): two co-resident CTAs hide each other's epilogue
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;   // warps 0-3: A loaders then epilogue; warp 4: MMA issue; warp 5: weight stages
constexpr int GEMM_A_STAGE = GEMM_MB * A_STAGE;
constexpr int GEMM_SMEM = GEMM_STAGES * (GEMM_A_STAGE + B_STAGE) + 256;
constexpr int FWD_LAND = 6;         // forward: landing slots of the raw fp32 A chunks (5 K tiles of loads in flight per thread)
constexpr int FWD_LAND_BYTES = TM * KT3 * 4;                                // 4 096
constexpr int FWD_SMEM = GEMM_SMEM + FWD_LAND * FWD_LAND_BYTES;
constexpr int WG_STAGES = 3;
constexpr int WG_A_STAGE = 8 * 4096;          // 8 blocks of 32 columns (256 = two M blocks) x 32 K rows x 128 B
constexpr int WG_B_STAGE = 7 * 4096;          // 7 blocks (224 >= 208 columns)
constexpr int WG_SMEM = WG_STAGES * (WG_A_STAGE + WG_B_STAGE) + 256 + 1024;   // + alignment slack
constexpr int WG_SPLIT = 49;                  // 49 x 3 layers = 147 CTAs, one wave
constexpr int THIN_THREADS = 208;             // 52 four-column groups x 4 row lanes
constexpr int THIN_CTAS = 592;             // dz4 / layer 0: 4 per SM (latency-bound streams: 33 -> 22 us for dz4)
constexpr int WG0_CTAS = 296;              // layer-0 wgrad: its per-CTA partial (208 x 9 floats) makes more CTAs a loss
constexpr int MAXF = 8;

// ---- PTX wrappers (own namespace: neural_af_tc.cu has same-named helpers with other constants) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// round to nearest tf32 (ties away from zero, = cvt.rna.tf32.f32 for finite values) with two full-rate integer
// instructions: cvt.rna runs on the conversion unit and made the forward loader's hi / lo split the slowest link of the
// pipeline (~270 cycles per 16-byte chunk, tools/upd_trace.py)
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
// Bounded wait: after 2 s (a pipeline bug) the CTA-wide abort flag is latched, every later wait returns at once,
// the kernel runs out on garbage and the host finds bit 0 of the error word set.  Nothing hangs.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (*abort_flag) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*abort_flag) return;
      if (globaltimer_ns() - t0 > 2000000000ull) { *abort_flag = 1; return; }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, abort_flag);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// whole-range L2 prefetch: DRAM sees one sequential burst instead of the pipeline's strided 32-byte pieces
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// 16-byte asynchronous copy, zero-filled when `valid` is false (src must still be a mapped address)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the mbarrier gets one (counted) arrival when every cp.async this thread has issued so far is complete
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, both operands through shared-memory descriptors
__device__ __forceinline__ void mma_ss_tf32(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (128 rows x 8 tf32 columns) read from tensor memory
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// cute::UMMA::SmemDescriptor, no swizzle, version 1.  K-major: LBO = distance of the two 16-byte K chunks of one
// MMA, SBO = distance of 8-row groups.  MN-major: SBO = distance of 16-byte chunks along M/N, LBO = distance of
// 8-row K groups (mma_sm100_desc.hpp make_umma_desc<Major::MN>, INTERLEAVE).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D fp32, A/B tf32, M = 128, N = 208; bits 15/16 = A/B MN-major
constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_MN = IDESC_K | (1u << 15) | (1u << 16);

// ---- arguments ---------------------------------------------------------------------------------------------------
struct Net {
  const float* W[5];
  const float* b[5];
  float* dW[5];
  float* db[5];
  int F, F_state, H;
  int col[MAXF];     // state column of policy feature f
};

struct Ws {          // workspace carve-up (device pointers)
  int* err;
  float* img_fwd[3];  // hi parts of [W | b], KT3 stages
  float* img_fwl[3];  // lo parts
  float* img_dgr[3];  // W^T, KT stages
  float* img_dgp[3];  // W^T, 32 K values per stage (persistent dgrad)
  float* Hs[4];      // h1..h4, plain fp32
  uint32_t* hbits[3]; // ReLU decisions of h1, h2, h3: [R][8] words, bit q of word c = (h[row][32 c + q] > 0)
  float* dZ[4];      // dZ1..dZ4
  float* logits;
  float* dl;
  float* terms;      // [E,4]: logp(a), entropy, -min(surr1, surr2) / n, ratio
  float* part_big;   // [WG_SPLIT][3][HP][HP]
  float* part_head;  // [THIN_CTAS][HP]
  float* part_l0;    // [THIN_CTAS][HP][MAXF + 1]
  void* tcpack;      // the current weights packed for the tcgen05 policy kernel (bf16x3 images), re-packed every call
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static size_t carve(int E, int M, unsigned char* base, Ws* w) {
  const size_t R = (size_t)E * M;
  const size_t Rp = align_up(R, TM);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return base ? base + o : nullptr; };
  Ws t;
  t.err = (int*)take(256);
  for (int l = 0; l < 3; ++l) t.img_fwd[l] = (float*)take((size_t)IMG_FLOATS * 4);
  for (int l = 0; l < 3; ++l) t.img_fwl[l] = (float*)take((size_t)IMG_FLOATS * 4);
  for (int l = 0; l < 3; ++l) t.img_dgr[l] = (float*)take((size_t)IMG_FLOATS * 4);
  for (int l = 0; l < 3; ++l) t.img_dgp[l] = (float*)take((size_t)IMG_FLOATS * 4);
  for (int l = 0; l < 4; ++l) t.Hs[l] = (float*)take(Rp * HP * 4);
  for (int l = 0; l < 3; ++l) t.hbits[l] = (uint32_t*)take(Rp * 8 * 4);
  for (int l = 0; l < 4; ++l) t.dZ[l] = (float*)take(Rp * HP * 4);
  t.logits = (float*)take(Rp * 4);
  t.dl = (float*)take(Rp * 4);
  t.terms = (float*)take((size_t)E * 4 * 4);
  t.part_big = (float*)take((size_t)WG_SPLIT * 3 * HP * HP * 4);
  t.part_head = (float*)take((size_t)THIN_CTAS * HP * 4);
  t.part_l0 = (float*)take((size_t)THIN_CTAS * HP * (MAXF + 1) * 4);
  t.tcpack = (void*)take(tc_pack_bytes());
  if (w) *w = t;
  return off;
}

// ---- weight images -----------------------------------------------------------------------------------------------
// K-major canonical, one stage per KT (dgrad) / KT3 (forward) K values: [kt][16-byte K chunk c][n (208)][4 floats]
// (4 chunks per dgrad stage, 2 per forward stage).
// Forward: [W | b] split into tf32 hi and lo parts (3xTF32: hi*hi + lo*hi + hi*lo, fp32-level logits); dgrad: W^T.
__global__ void upd_pack_kernel(Net net, Ws ws, int kinds) {      // kinds: bit k set = image kind k is needed by this call
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * IMG_FLOATS) return;
  const int img = idx / IMG_FLOATS, o = idx - img * IMG_FLOATS;
  const int kind = img / 3, l = img - kind * 3;     // kind 0: forward hi / lo, 1: dgrad (KT), 2: dgrad (32 per stage); layer l + 1
  if (!((kinds >> kind) & 1)) return;
  const int kt_len = kind == 0 ? KT3 : (kind == 1 ? KT : 32);
  const int kt = o / (HP * kt_len), r = o - kt * (HP * kt_len);
  const int c = r / (HP * 4), r2 = r - c * (HP * 4);
  const int n = r2 >> 2, e = r2 & 3;
  const int k = kt * kt_len + c * 4 + e;
  const int H = net.H;
  const float* W = net.W[l + 1];
  float v = 0.f;
  if (kind == 0) {
    if (k >= HP) return;
    if (n < H) v = k < H ? W[(size_t)n * H + k] : (k == H ? net.b[l + 1][n] : 0.f);     // [W | b]
    const float hi = tf32_rn(v);
    ws.img_fwd[l][o] = hi;
    ws.img_fwl[l][o] = tf32_rn(v - hi);
  } else {
    if (kind == 1 && k >= HP) return;
    if (n < H && k < H) v = W[(size_t)k * H + n];                                        // W^T: N = in, K = out
    (kind == 1 ? ws.img_dgr[l] : ws.img_dgp[l])[o] = tf32_rn(v);
  }
}

// ---- layer 0 on CUDA cores ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THIN_THREADS) upd_layer0_kernel(Net net, const float* __restrict__ states, int R,
                                                                   float* __restrict__ H1) {
  const int cg = threadIdx.x % 52, rsub = threadIdx.x / 52;
  const int F = net.F, H = net.H;
  float w[4][MAXF], bias[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = cg * 4 + e;
    bias[e] = j < H ? net.b[0][j] : 0.f;
#pragma unroll
    for (int f = 0; f < MAXF; ++f) w[e][f] = (j < H && f < F) ? net.W[0][(size_t)j * F + f] : 0.f;
  }
  const int rpc = (R + gridDim.x - 1) / gridDim.x;
  const int r_end = min(R, (int)(blockIdx.x + 1) * rpc);
  for (int r = blockIdx.x * rpc + rsub; r < r_end; r += 4) {
    float x[MAXF];
#pragma unroll
    for (int f = 0; f < MAXF; ++f) x[f] = f < F ? __ldg(states + (size_t)r * net.F_state + net.col[f]) : 0.f;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // torch's Linear on the CPU accumulates the dot product first and adds the bias last
      float acc = 0.f;
#pragma unroll
      for (int f = 0; f < MAXF; ++f) acc = fmaf(x[f], w[e][f], acc);
      const int j = cg * 4 + e;
      acc = fmaxf(acc + bias[e], 0.f);
      o[e] = j < H ? acc : (j == H ? 1.f : 0.f);
    }
    *(float4*)(H1 + (size_t)r * HP + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

#ifdef MBO_UPD_TRACE
// debug build (tools/build_trace_lib.sh): clock64 stamps of one CTA's pipeline roles, read back by tools/upd_trace.py
__device__ long long g_upd_trace[4096];
#define UPD_TRACE(role, idx) do { if (blockIdx.x == gridDim.x / 3) g_upd_trace[(role) * 512 + (idx)] = clock64(); } while (0)
#else
#define UPD_TRACE(role, idx) do { } while (0)
#endif

// ---- forward / dgrad GEMM -----------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A;        // [R, HP] rows of this GEMM (h_l for the forward, dZ_{l+1} for the dgrad), plain fp32
  const float* Bimg;     // weight image (forward: hi parts)
  const float* Bimg_lo;  // forward: lo parts
  float* Out;            // [R, HP]: h_{l+1} (fp32, unrounded) or dZ_l (rounded to tf32)
  const uint32_t* bits_in;   // dgrad: ReLU decisions of h_l (nullptr: no mask, the consumer applies it)
  uint32_t* bits_out;        // forward: ReLU decisions of h_{l+1} for the dgrad, or nullptr
  uint32_t* bits_a;          // TS forward: ReLU decisions of its INPUT h_l, built by the producer threads (thread = row), or nullptr
  const float* w4;       // forward of the last hidden layer: head weights [H], else nullptr
  const float* b4;
  float* logits;         // [R]
  int R, H;
  int dbg;               // timing experiments (MBO_UPD_DBG): 1 no epilogue, 2 no MMAs, 4 no A loads, 8 no weight loads
};

// EPI 0: forward, 3xTF32 (stage = 16 K values x {hi, lo} of both operands, 3 MMAs per K = 8 step), epilogue ReLU +
//        hi / lo split + ones column (+ head logit from the unrounded fp32 activations).
// EPI 1: dgrad, single TF32 (stage = 32 K values), epilogue multiplies by the ReLU mask of h_l.
// The epilogue moves every 32-column chunk through a per-warp staging tile in shared memory (the pipeline stages are
// idle by then) so that global memory sees whole 128-byte lines: thread = row from TMEM, lane = 16-byte chunk towards HBM.
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) upd_gemm_kernel(GemmArgs a, int* __restrict__ err) {
  constexpr bool X3 = EPI == 0;
  constexpr int NKT_ = X3 ? NKT3 : NKT;
  constexpr int KT_ = X3 ? KT3 : KT;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + GEMM_STAGES * GEMM_A_STAGE;
  uint64_t* bars = (uint64_t*)(smem + GEMM_STAGES * (GEMM_A_STAGE + B_STAGE));
  unsigned char* sLand = smem + GEMM_STAGES * (GEMM_A_STAGE + B_STAGE) + 256;       // forward only
  uint32_t* s_tmem = (uint32_t*)(bars + 2 * GEMM_STAGES + 2);
  volatile int* s_abort = (volatile int*)(s_tmem + 1);
  __shared__ float s_w4[HP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * (TM * GEMM_MB);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (GEMM_STAGES + s); };
  const uint32_t ACCF = bar0 + 8u * (2 * GEMM_STAGES);

  if (tid == 0) {
    *s_abort = 0;
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(FULL(s), (X3 ? 64 : 128) + 1); mbar_init(EMPTY(s), 1); }
    mbar_init(ACCF, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (EPI == 0 && a.w4) for (int i = tid; i < HP; i += GEMM_THREADS) s_w4[i] = i < a.H ? a.w4[i] : 0.f;
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u * GEMM_MB) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // ---- A loaders: lane -> (16-byte chunk quarter cq, row-in-group r8): the 8 lanes of one cq write 128 contiguous
    // bytes of one core-matrix column (conflict-free) and read 64 B per row ----
    const int r8 = lane & 7, cq = lane >> 3;
    if (X3) {
      // ---- forward, activations stored once as plain fp32: every thread copies 2 chunks (16 B) per K tile into a ring
      // of FWD_LAND landing slots, 5 tiles ahead of their use (the ring is private per thread: a slot is reused only after
      // this thread has split its own chunks out of it, so no barrier is involved and the loads run ahead of the MMAs),
      // then splits exactly those chunks into tf32 hi and lo inside the operand stage -- only its own cp.async groups
      // have to be complete.  (With the loads issued into the 4 operand stages, 2 tiles ahead, the kernel waited for
      // DRAM on every tile: 34 of its 54 us.) ----
      // The split + fence.proxy.async + arrive chain of one K tile takes a thread ~700 cycles (tools/upd_trace.py), more
      // than the 3 MMAs of the tile: the four loader warps therefore work as two groups on alternate K tiles (64
      // threads x 4 chunks each), so that two tiles are in that chain at any time.
      constexpr int D = 2;                                               // own tiles in flight beyond the current one (6 slots / 2 groups - 1)
      const int grp = warp >> 1;                                         // K tiles kt = grp, grp + 2, ...
      const int j0 = (warp & 1) * 8 + (cq >> 1);                          // 8-row group of chunk i = 0 (i: + 2 i)
      const uint32_t coff = (cq & 1) * (TM * 16) + r8 * 16;
      auto issue = [&](int kt) {
        if (kt < NKT_) {
          const uint32_t land = smem_u32(sLand + (kt % FWD_LAND) * FWD_LAND_BYTES);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = j0 + i * 2;
            const int grow = row0 + j * 8 + r8;
            const bool ok = grow < a.R;
            const float* src = a.A + (size_t)(ok ? grow : 0) * HP + kt * KT_ + (cq & 1) * 4;
            cp_async16(land + coff + j * 128, src, ok);
          }
        }
        cp_async_commit();                                                // one group per own K tile, also past the end
      };
      for (int n = 0; n < D; ++n) issue(grp + 2 * n);
      for (int kt = grp; kt < NKT_; kt += 2) {
        const int s = kt % GEMM_STAGES, u = kt / GEMM_STAGES;
        if (tid == 0) UPD_TRACE(0, 4 * kt + 0);
        cp_async_wait<D - 1>();                                           // tile kt has landed (this thread's chunks)
        if (tid == 0) UPD_TRACE(0, 4 * kt + 1);
        if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
        if (tid == 0) UPD_TRACE(0, 4 * kt + 2);
        const uint32_t land = smem_u32(sLand + (kt % FWD_LAND) * FWD_LAND_BYTES);
        const uint32_t opnd = smem_u32(sA + s * GEMM_A_STAGE);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = coff + (j0 + i * 2) * 128;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(land + off) : "memory");
          const float4 hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
          const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);    // exact; the MMA drops bits below 2^-22 |v|
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(opnd + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(opnd + off + TM * KT3 * 4), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
        }
        if (tid == 0) UPD_TRACE(4, 2 * kt + 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0) UPD_TRACE(4, 2 * kt + 1);
        mbar_arrive(FULL(s));
        if (tid == 0) UPD_TRACE(0, 4 * kt + 3);
        issue(kt + 2 * D);                                                // into the slot of tile kt - 2: this group's previous tile
      }
      cp_async_wait<0>();
    } else
    for (int kt = 0; kt < NKT_; ++kt) {
      const int s = kt % GEMM_STAGES, u = kt / GEMM_STAGES;
      if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
#pragma unroll
      for (int mb = 0; mb < GEMM_MB; ++mb) {
        const uint32_t dstA = smem_u32(sA + s * GEMM_A_STAGE + mb * A_STAGE);
        const int brow0 = row0 + mb * TM;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = warp * 4 + i;
          const int grow = brow0 + j * 8 + r8;               // j = 8-row group
          const bool ok = grow < a.R;
          if (a.dbg & 4) continue;
          const float* src = a.A + (size_t)(ok ? grow : 0) * HP + kt * KT_ + cq * 4;
          cp_async16(dstA + cq * (TM * 16) + j * 128 + r8 * 16, src, ok);
        }
      }
      cp_async_arrive(FULL(s));
    }
    // ---- epilogue ----
    uint32_t mbits[GEMM_MB][8];
    if (EPI == 1) {                        // the row's ReLU words: fetched while the last MMAs are still running
#pragma unroll
      for (int mb = 0; mb < GEMM_MB; ++mb) {
        const int row = row0 + mb * TM + warp * 32 + lane;
        uint4 w0 = make_uint4(~0u, ~0u, ~0u, ~0u), w1 = w0;
        if (a.bits_in && row < a.R) {
          w0 = *(const uint4*)(a.bits_in + (size_t)row * 8);
          w1 = *(const uint4*)(a.bits_in + (size_t)row * 8 + 4);
        }
        mbits[mb][0] = w0.x; mbits[mb][1] = w0.y; mbits[mb][2] = w0.z; mbits[mb][3] = w0.w;
        mbits[mb][4] = w1.x; mbits[mb][5] = w1.y; mbits[mb][6] = w1.z; mbits[mb][7] = w1.w;
      }
    }
    if (tid == 0) UPD_TRACE(3, 0);
    mbar_wait(ACCF, 0, s_abort);
    tc_fence_after();
    if (tid == 0) UPD_TRACE(3, 1);
    float* stg = (float*)smem + warp * (32 * STG_PITCH);            // per-warp staging tile [32 rows][36 floats]
    const int cr = lane >> 3, cc = lane & 7;                         // coalesced side: 4 rows x 8 chunks per instruction
#pragma unroll 1
    for (int mb = 0; mb < GEMM_MB; ++mb) {
    const int wrow0 = row0 + mb * TM + warp * 32;
    if (wrow0 >= a.R || (a.dbg & 1)) break;                                         // warp-uniform: nothing to store for these rows
    const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + mb * 256;
    float head = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < HP; c0 += 32) {
      uint32_t v[32];
      const int nc = HP - c0 >= 32 ? 32 : 16;
      if (nc == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
      tmem_wait_ld();
      float x[32];
      if (EPI == 0) {
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          x[q] = fmaxf(__uint_as_float(v[q]), 0.f);
          if (q < nc) {
            if (a.w4) head = fmaf(x[q], s_w4[c0 + q], head);
            word |= (x[q] > 0.f ? 1u : 0u) << q;
          }
        }
        if (a.bits_out && wrow0 + lane < a.R) a.bits_out[(size_t)(wrow0 + lane) * 8 + (c0 >> 5)] = word;
      } else {
        const uint32_t word = mbits[mb][c0 >> 5];
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = (q < nc && ((word >> q) & 1u)) ? tf32_rn(__uint_as_float(v[q])) : 0.f;
      }
      {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q * 4 < nc) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float t = x[q * 4 + e];
              if (EPI == 0 && c0 + q * 4 + e == a.H) t = 1.f;          // ones column; columns above H are 0 (zero weight rows)
              o[e] = t;
            }
            *(float4*)(stg + lane * STG_PITCH + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
        float* outp = a.Out;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + cr, grow = wrow0 + r;
          if (grow < a.R && cc * 4 < nc)
            *(float4*)(outp + (size_t)grow * HP + c0 + cc * 4) = *(const float4*)(stg + r * STG_PITCH + cc * 4);
        }
        __syncwarp();
      }
    }
    if (EPI == 0 && a.w4 && wrow0 + lane < a.R) a.logits[wrow0 + lane] = head + a.b4[0];
    }
    if (tid == 0) UPD_TRACE(3, 2);
    tc_fence_before();
  } else if (warp == 4) {
    // ---- MMA issue ----
    if (lane == 0) {
      for (int kt = 0; kt < NKT_; ++kt) {
        const int s = kt % GEMM_STAGES, u = kt / GEMM_STAGES;
        UPD_TRACE(1, 3 * kt + 0);
        mbar_wait(FULL(s), u & 1, s_abort);
        tc_fence_after();
        UPD_TRACE(1, 3 * kt + 1);
        const uint32_t aS0 = smem_u32(sA + s * GEMM_A_STAGE), bS = smem_u32(sB + s * B_STAGE);
        if (a.dbg & 2) {
        } else if (X3) {
          const uint64_t bh = make_desc(bS, HP * 16, 128);
          const uint64_t bl = make_desc(bS + HP * KT3 * 4, HP * 16, 128);
#pragma unroll
          for (int mb = 0; mb < GEMM_MB; ++mb) {
            const uint32_t aS = aS0 + mb * A_STAGE;
            const uint64_t ah = make_desc(aS, TM * 16, 128);
            const uint64_t al = make_desc(aS + TM * KT3 * 4, TM * 16, 128);
            mma_ss_tf32(tmem + mb * 256, al, bh, IDESC_K, kt != 0);        // small terms first
            mma_ss_tf32(tmem + mb * 256, ah, bl, IDESC_K, 1);
            mma_ss_tf32(tmem + mb * 256, ah, bh, IDESC_K, 1);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t bd = make_desc(bS + ks * 2 * (HP * 16), HP * 16, 128);
#pragma unroll
            for (int mb = 0; mb < GEMM_MB; ++mb) {
              const uint64_t ad = make_desc(aS0 + mb * A_STAGE + ks * 2 * (TM * 16), TM * 16, 128);
              mma_ss_tf32(tmem + mb * 256, ad, bd, IDESC_K, (kt | ks) != 0);
            }
          }
        }
        tc_commit(EMPTY(s));
        UPD_TRACE(1, 3 * kt + 2);
      }
      tc_commit(ACCF);
    }
    __syncwarp();
  } else {
    // ---- the CTA's A tile is one contiguous block of rows (pitch = row length): pull it into L2 in 16-row pieces.
    // The pipeline reads 32 (forward) or 64 bytes per row and stage, i.e. every DRAM page of the tile 26 / 13 times
    // over; without this the kernel ran at ~1.3 TB/s of DRAM reads ----
    {
      const int r_lo = row0 + lane * 16 * GEMM_MB;
      const int nrow = min(16 * GEMM_MB, a.R - r_lo);
      if (lane < 8 && nrow > 0) bulk_prefetch_l2(a.A + (size_t)r_lo * HP, (uint32_t)nrow * HP * 4);
    }
    // ---- weight stages: bulk copies of the packed images ----
    if (lane == 0) {
      for (int kt = 0; kt < NKT_; ++kt) {
        const int s = kt % GEMM_STAGES, u = kt / GEMM_STAGES;
        UPD_TRACE(2, 2 * kt + 0);
        if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
        if (a.dbg & 8) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(FULL(s)) : "memory");
          continue;
        }
        mbar_expect_tx(FULL(s), B_STAGE);
        if (X3) {
          bulk_g2s(smem_u32(sB + s * B_STAGE), a.Bimg + (size_t)kt * (HP * KT3), HP * KT3 * 4, FULL(s));
          bulk_g2s(smem_u32(sB + s * B_STAGE) + HP * KT3 * 4, a.Bimg_lo + (size_t)kt * (HP * KT3), HP * KT3 * 4, FULL(s));
        } else {
          bulk_g2s(smem_u32(sB + s * B_STAGE), a.Bimg + (size_t)kt * (HP * KT), B_STAGE, FULL(s));
        }
        UPD_TRACE(2, 2 * kt + 1);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0 && *s_abort) atomicOr(err, 1);
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u * GEMM_MB) : "memory");
  }
}

// ---- forward GEMM with the A operand in tensor memory (TS form) ---------------------------------------------------------
// The SS-form forward is bound by the shared-memory pipe (its three MMAs per K tile read A hi twice, A lo, B hi twice,
// B lo, and the loaders write the split operands): here the activations never touch shared memory.  Thread = row: each
// loader thread reads its row straight from global memory (one 128-byte line = 4 K tiles at a time, one line ahead in
// registers), splits in registers into tf32 hi / lo and writes them with tcgen05.st into a 3-slot ring of 16 TMEM columns
// next to the 208 accumulator columns (256-column allocation, 2 CTAs per SM).  Only the weight stages (8 x 13 KB, bulk
// copies) live in shared memory.  Warps 0-3: A producers, then epilogue; warp 4: MMA issue; warp 5: weight stages.
constexpr int TS_BSTAGES = 8;
constexpr int TS_ASLOTS = 3;
constexpr int TS_ACOL = HP;                              // first TMEM column of the A ring
constexpr int TS_SMEM = TS_BSTAGES * B_STAGE + 512;      // 107 008: two CTAs per SM; the epilogue's staging tile aliases the idle weight stages

// EPI 0: forward (3xTF32, 26 K tiles of 8: hi and lo halves of a 16-column slot); EPI 1: dgrad (13 K tiles of 16: dZ is
// already rounded to tf32, one 16-column slot = two K = 8 MMAs), epilogue masked by the ReLU bit words.
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) upd_ts_kernel(GemmArgs a, int* __restrict__ err) {
  constexpr int NKT_ = EPI == 0 ? NKT3 : NKT;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sB = smem;
  float* s_stg = (float*)smem;                            // used after ACCF: every weight stage has been consumed by then
  uint64_t* bars = (uint64_t*)(smem + TS_BSTAGES * B_STAGE);
  uint32_t* s_tmem = (uint32_t*)(bars + 2 * TS_BSTAGES + 2 * TS_ASLOTS + 2);
  volatile int* s_abort = (volatile int*)(s_tmem + 1);
  __shared__ float s_w4[HP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TM;
  const uint32_t bar0 = smem_u32(bars);
  auto BFULL = [&](int s) { return bar0 + 8u * s; };
  auto BEMPTY = [&](int s) { return bar0 + 8u * (TS_BSTAGES + s); };
  auto AFULL = [&](int s) { return bar0 + 8u * (2 * TS_BSTAGES + s); };
  auto AEMPTY = [&](int s) { return bar0 + 8u * (2 * TS_BSTAGES + TS_ASLOTS + s); };
  const uint32_t ACCF = bar0 + 8u * (2 * TS_BSTAGES + 2 * TS_ASLOTS);

  if (tid == 0) {
    *s_abort = 0;
    for (int s = 0; s < TS_BSTAGES; ++s) { mbar_init(BFULL(s), 1); mbar_init(BEMPTY(s), 1); }
    for (int s = 0; s < TS_ASLOTS; ++s) { mbar_init(AFULL(s), 128); mbar_init(AEMPTY(s), 1); }
    mbar_init(ACCF, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (EPI == 0 && a.w4) for (int i = tid; i < HP; i += GEMM_THREADS) s_w4[i] = i < a.H ? a.w4[i] : 0.f;
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // ---- A producers: thread = row (TMEM lane) ----
    const int row = row0 + warp * 32 + lane;
    const bool ok = row < a.R;
    const float4* src = (const float4*)(a.A + (size_t)(ok ? row : 0) * HP);
    const uint32_t tA = tmem + ((uint32_t)(warp * 32) << 16) + TS_ACOL;
    constexpr int NLINE = (HP + 31) / 32;                 // 7 lines of 32 columns (the last one half)
    float4 cur[8], nxt[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cur[q] = ok ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int ln = 0; ln < NLINE; ++ln) {
      const int nq = ln + 1 < NLINE ? (HP - (ln + 1) * 32 >= 32 ? 8 : (HP - (ln + 1) * 32) / 4) : 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) nxt[q] = (ok && q < nq) ? src[(ln + 1) * 8 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == 0 && a.bits_a && ok) {                                   // ReLU word of this 32-column line of the input
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          word |= (cur[q].x > 0.f ? 1u : 0u) << (4 * q) | (cur[q].y > 0.f ? 1u : 0u) << (4 * q + 1) |
                  (cur[q].z > 0.f ? 1u : 0u) << (4 * q + 2) | (cur[q].w > 0.f ? 1u : 0u) << (4 * q + 3);
        }
        if (ln * 32 <= a.H && a.H < ln * 32 + 32) word &= ~(1u << (a.H - ln * 32));      // the ones column is not a unit
        a.bits_a[(size_t)row * 8 + ln] = word;
      }
      if (EPI == 0) {
        const int ntile = ln + 1 < NLINE ? 4 : (HP - ln * 32) / 8;        // K tiles in this line (4, last line 2)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < ntile) {
            const int kt = ln * 4 + t, slot = kt % TS_ASLOTS, u = kt / TS_ASLOTS;
            const float v[8] = {cur[2 * t].x, cur[2 * t].y, cur[2 * t].z, cur[2 * t].w,
                                cur[2 * t + 1].x, cur[2 * t + 1].y, cur[2 * t + 1].z, cur[2 * t + 1].w};
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float h = tf32_rn(v[e]);
              hi[e] = __float_as_uint(h);
              lo[e] = __float_as_uint(v[e] - h);             // exact; the MMA drops bits below 2^-22 |v|
            }
            if (u > 0) { mbar_wait(AEMPTY(slot), (u - 1) & 1, s_abort); tc_fence_after(); }
            tmem_st8(tA + slot * 16, hi);
            tmem_st8(tA + slot * 16 + 8, lo);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(AFULL(slot));
          }
        }
      } else {
        const int ntile = ln + 1 < NLINE ? 2 : (HP - ln * 32) / 16;       // K tiles of 16 in this line (2, last line 1)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < ntile) {
            const int kt = ln * 2 + t, slot = kt % TS_ASLOTS, u = kt / TS_ASLOTS;
            uint32_t w[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              w[4 * q + 0] = __float_as_uint(cur[4 * t + q].x); w[4 * q + 1] = __float_as_uint(cur[4 * t + q].y);
              w[4 * q + 2] = __float_as_uint(cur[4 * t + q].z); w[4 * q + 3] = __float_as_uint(cur[4 * t + q].w);
            }
            if (u > 0) { mbar_wait(AEMPTY(slot), (u - 1) & 1, s_abort); tc_fence_after(); }
            tmem_st16(tA + slot * 16, w);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(AFULL(slot));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
    }
    // ---- epilogue: (forward) ReLU, ones column, head logit, ReLU bit words / (dgrad) ReLU mask from the bit words;
    // whole 128-byte lines through a staging tile ----
    uint32_t mbits[8];
    if (EPI == 1) {
      uint4 w0 = make_uint4(~0u, ~0u, ~0u, ~0u), w1 = w0;
      if (a.bits_in && ok) {
        w0 = *(const uint4*)(a.bits_in + (size_t)row * 8);
        w1 = *(const uint4*)(a.bits_in + (size_t)row * 8 + 4);
      }
      mbits[0] = w0.x; mbits[1] = w0.y; mbits[2] = w0.z; mbits[3] = w0.w;
      mbits[4] = w1.x; mbits[5] = w1.y; mbits[6] = w1.z; mbits[7] = w1.w;
    }
    mbar_wait(ACCF, 0, s_abort);
    tc_fence_after();
    float* stg = s_stg + warp * (32 * STG_PITCH);
    const int cr = lane >> 3, cc = lane & 7;
    const int wrow0 = row0 + warp * 32;
    if (wrow0 < a.R) {
      const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16);
      float head = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < HP; c0 += 32) {
        uint32_t v[32];
        const int nc = HP - c0 >= 32 ? 32 : 16;
        if (nc == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
        tmem_wait_ld();
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q * 4 < nc) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = c0 + q * 4 + e;
              if (EPI == 0) {
                const float x = fmaxf(__uint_as_float(v[q * 4 + e]), 0.f);
                if (a.w4) head = fmaf(x, s_w4[col], head);
                word |= (x > 0.f ? 1u : 0u) << (q * 4 + e);
                o[e] = col == a.H ? 1.f : x;                  // ones column; columns above H are 0 (zero weight rows)
              } else {
                o[e] = ((mbits[c0 >> 5] >> (q * 4 + e)) & 1u) ? tf32_rn(__uint_as_float(v[q * 4 + e])) : 0.f;
              }
            }
            *(float4*)(stg + lane * STG_PITCH + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        if (EPI == 0 && a.bits_out && wrow0 + lane < a.R) a.bits_out[(size_t)(wrow0 + lane) * 8 + (c0 >> 5)] = word;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + cr, grow = wrow0 + r;
          if (grow < a.R && cc * 4 < nc)
            *(float4*)(a.Out + (size_t)grow * HP + c0 + cc * 4) = *(const float4*)(stg + r * STG_PITCH + cc * 4);
        }
        __syncwarp();
      }
      if (EPI == 0 && a.w4 && wrow0 + lane < a.R) a.logits[wrow0 + lane] = head + a.b4[0];
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ---- MMA issue: per K tile lo*hi + hi*lo + hi*hi ----
    if (lane == 0) {
      for (int kt = 0; kt < NKT_; ++kt) {
        const int sb = kt % TS_BSTAGES, ub = kt / TS_BSTAGES, sa = kt % TS_ASLOTS, ua = kt / TS_ASLOTS;
        mbar_wait(BFULL(sb), ub & 1, s_abort);
        mbar_wait(AFULL(sa), ua & 1, s_abort);
        tc_fence_after();
        const uint32_t bS = smem_u32(sB + sb * B_STAGE);
        const uint64_t b0 = make_desc(bS, HP * 16, 128);                       // forward: hi image; dgrad: K 0..7
        const uint64_t b1 = make_desc(bS + HP * KT3 * 4, HP * 16, 128);        // forward: lo image; dgrad: K 8..15
        const uint32_t a0 = tmem + TS_ACOL + sa * 16, a1 = a0 + 8;             // forward: hi / lo; dgrad: K 0..7 / 8..15
        if (EPI == 0) {
          mma_ts_tf32(tmem, a1, b0, IDESC_K, kt != 0);       // lo * hi: small terms first
          mma_ts_tf32(tmem, a0, b1, IDESC_K, 1);
          mma_ts_tf32(tmem, a0, b0, IDESC_K, 1);
        } else {
          mma_ts_tf32(tmem, a0, b0, IDESC_K, kt != 0);
          mma_ts_tf32(tmem, a1, b1, IDESC_K, 1);
        }
        tc_commit(BEMPTY(sb));
        tc_commit(AEMPTY(sa));
      }
      tc_commit(ACCF);
    }
    __syncwarp();
  } else {
    // ---- weight stages: hi and lo image of one K tile per stage ----
    if (lane == 0) {
      for (int kt = 0; kt < NKT_; ++kt) {
        const int s = kt % TS_BSTAGES, u = kt / TS_BSTAGES;
        if (u > 0) mbar_wait(BEMPTY(s), (u - 1) & 1, s_abort);
        mbar_expect_tx(BFULL(s), B_STAGE);
        if (EPI == 0) {
          bulk_g2s(smem_u32(sB + s * B_STAGE), a.Bimg + (size_t)kt * (HP * KT3), HP * KT3 * 4, BFULL(s));
          bulk_g2s(smem_u32(sB + s * B_STAGE) + HP * KT3 * 4, a.Bimg_lo + (size_t)kt * (HP * KT3), HP * KT3 * 4, BFULL(s));
        } else {
          bulk_g2s(smem_u32(sB + s * B_STAGE), a.Bimg + (size_t)kt * (HP * KT), B_STAGE, BFULL(s));
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0 && *s_abort) atomicOr(err, 1);
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// ---- persistent dgrad GEMM -----------------------------------------------------------------------------------------
// One CTA per SM over a contiguous range of rows (R / grid, rounded up to 8), 128-row tiles inside it; two accumulators
// in TMEM (2 x 256 columns) so that the epilogue warps drain tile i while the loaders / MMA warp already run tile i + 1.
// Warps 0-3: epilogue, 4-7: A loaders, 8: MMA issue, 9: weight stages.  4 stages of 32 K values: every row gives a whole
// 128-byte line per stage (lane -> (row of 4, chunk of 8)); the 16-byte chunk planes are padded by 16 B so that those
// cp.async stores stay conflict-free.  Same MMA sequence per tile as upd_gemm_kernel<1>: bitwise identical results.
constexpr int PG_STAGES = 4;
constexpr int PG_KT = 32, PG_NKT = 7;
constexpr int PG_LBO_A = TM * 16 + 16;
constexpr int PG_A_STAGE = 8 * PG_LBO_A;                 // 16 512
constexpr int PG_B_STAGE = HP * PG_KT * 4;               // 26 624
constexpr int PG_THREADS = 320;
constexpr int PG_STG_BYTES = 4 * 32 * STG_PITCH * 4;     // 18 432
constexpr int PG_SMEM = PG_STAGES * (PG_A_STAGE + PG_B_STAGE) + PG_STG_BYTES + 256;

__global__ void __launch_bounds__(PG_THREADS, 1) upd_dgrad_persist_kernel(GemmArgs a, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + PG_STAGES * PG_A_STAGE;
  float* s_stg = (float*)(smem + PG_STAGES * (PG_A_STAGE + PG_B_STAGE));
  uint64_t* bars = (uint64_t*)(smem + PG_STAGES * (PG_A_STAGE + PG_B_STAGE) + PG_STG_BYTES);
  uint32_t* s_tmem = (uint32_t*)(bars + 2 * PG_STAGES + 4);
  volatile int* s_abort = (volatile int*)(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (PG_STAGES + s); };
  auto ACCF = [&](int i) { return bar0 + 8u * (2 * PG_STAGES + i); };
  auto ACCE = [&](int i) { return bar0 + 8u * (2 * PG_STAGES + 2 + i); };
  // this CTA's rows
  const int rpc = (((a.R + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;
  const int r_begin = min(a.R, (int)blockIdx.x * rpc), r_end = min(a.R, r_begin + rpc);
  const int n_tiles = (r_end - r_begin + TM - 1) / TM;

  if (tid == 0) {
    *s_abort = 0;
    for (int s = 0; s < PG_STAGES; ++s) { mbar_init(FULL(s), 128 + 1); mbar_init(EMPTY(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(ACCF(i), 1); mbar_init(ACCE(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // ---- epilogue warps ----
    float* stg = s_stg + warp * (32 * STG_PITCH);
    const int cr = lane >> 3, cc = lane & 7;
    for (int t = 0; t < n_tiles; ++t) {
      const int slot = t & 1;
      const int wrow0 = r_begin + t * TM + warp * 32;
      uint32_t mbits[8];
      {                                                      // the row's ReLU words, fetched before the accumulator is ready
        const int row = wrow0 + lane;
        uint4 w0 = make_uint4(~0u, ~0u, ~0u, ~0u), w1 = w0;
        if (a.bits_in && row < r_end) {
          w0 = *(const uint4*)(a.bits_in + (size_t)row * 8);
          w1 = *(const uint4*)(a.bits_in + (size_t)row * 8 + 4);
        }
        mbits[0] = w0.x; mbits[1] = w0.y; mbits[2] = w0.z; mbits[3] = w0.w;
        mbits[4] = w1.x; mbits[5] = w1.y; mbits[6] = w1.z; mbits[7] = w1.w;
      }
      mbar_wait(ACCF(slot), (t >> 1) & 1, s_abort);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + slot * 256;
      const bool live = wrow0 < r_end;                       // warp-uniform
#pragma unroll 1
      for (int c0 = 0; c0 < HP; c0 += 32) {
        uint32_t v[32];
        const int nc = HP - c0 >= 32 ? 32 : 16;
        if (nc == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
        tmem_wait_ld();
        if (c0 + 32 >= HP) {                                 // the accumulator is in registers: hand the slot back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACCE(slot));
        }
        if (!live || (a.dbg & 1)) continue;
        const uint32_t word = mbits[c0 >> 5];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q * 4 < nc) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = ((word >> (q * 4 + e)) & 1u) ? tf32_rn(__uint_as_float(v[q * 4 + e])) : 0.f;
            *(float4*)(stg + lane * STG_PITCH + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + cr, grow = wrow0 + r;
          if (grow < r_end && cc * 4 < nc)
            *(float4*)(a.Out + (size_t)grow * HP + c0 + cc * 4) = *(const float4*)(stg + r * STG_PITCH + cc * 4);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ---- A loaders: one instruction = 4 whole 128-byte lines ----
    const int lw = warp - 4;
    const int lr = lane >> 3, c = lane & 7;
    int g = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int row0 = r_begin + t * TM;
      for (int kt = 0; kt < PG_NKT; ++kt, ++g) {
        const int s = g % PG_STAGES, u = g / PG_STAGES;
        if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
        const uint32_t dstA = smem_u32(sA + s * PG_A_STAGE);
        if (!(a.dbg & 4)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lw * 8 + i) * 4 + lr;
            const int grow = row0 + row;
            const bool ok = grow < r_end && kt * PG_KT + c * 4 < HP;
            const float* src = a.A + (size_t)(ok ? grow : 0) * HP + (ok ? kt * PG_KT + c * 4 : 0);
            cp_async16(dstA + c * PG_LBO_A + (row >> 3) * 128 + (row & 7) * 16, src, ok);
          }
        }
        cp_async_arrive(FULL(s));
      }
    }
  } else if (warp == 8) {
    // ---- MMA issue ----
    if (lane == 0) {
      int g = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int slot = t & 1;
        if (t >= 2) { mbar_wait(ACCE(slot), ((t >> 1) - 1) & 1, s_abort); tc_fence_after(); }
        const uint32_t d = tmem + slot * 256;
        for (int kt = 0; kt < PG_NKT; ++kt, ++g) {
          const int s = g % PG_STAGES, u = g / PG_STAGES;
          mbar_wait(FULL(s), u & 1, s_abort);
          tc_fence_after();
          const uint32_t aS = smem_u32(sA + s * PG_A_STAGE), bS = smem_u32(sB + s * PG_B_STAGE);
          const int nks = (a.dbg & 2) ? 0 : (kt == PG_NKT - 1 ? 2 : 4);       // the last stage holds K 192..207
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t bd = make_desc(bS + ks * 2 * (HP * 16), HP * 16, 128);
            const uint64_t ad = make_desc(aS + ks * 2 * PG_LBO_A, PG_LBO_A, 128);
            mma_ss_tf32(d, ad, bd, IDESC_K, (kt | ks) != 0);
          }
          tc_commit(EMPTY(s));
        }
        tc_commit(ACCF(slot));
      }
    }
    __syncwarp();
  } else {
    // ---- weight stages ----
    if (lane == 0) {
      int g = 0;
      for (int t = 0; t < n_tiles; ++t) {
        for (int kt = 0; kt < PG_NKT; ++kt, ++g) {
          const int s = g % PG_STAGES, u = g / PG_STAGES;
          if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
          if (a.dbg & 8) { mbar_arrive(FULL(s)); continue; }
          mbar_expect_tx(FULL(s), PG_B_STAGE);
          bulk_g2s(smem_u32(sB + s * PG_B_STAGE), a.Bimg + (size_t)kt * (HP * PG_KT), PG_B_STAGE, FULL(s));
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && *s_abort) atomicOr(err, 1);
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// ---- fused dgrad chain: dZ4 -> dZ3 -> dZ2 -> dZ1 in ONE persistent kernel ----------------------------------------------
// The three dgrad GEMMs of a 128-row tile run back to back in one CTA: the masked, tf32-rounded output of GEMM i leaves
// through the staging tile to HBM (the wgrad needs it) AND is written by the same epilogue threads straight into the
// NEXT GEMM's A operand in shared memory (the whole K extent of the tile, 7 K stages in the K-major canonical layout), so
// dZ3 and dZ2 are never read back from HBM: 2 of the 3 A streams (96 MB each at C3) and two launches disappear, and the
// next GEMM's MMAs start on K stage 0 while the epilogue is still converting stage 1.  Same MMA sequence, same masks and
// roundings per tile as upd_dgrad_persist_kernel: bitwise identical results (tested).
// Warps 0-7: epilogue (the accumulator's column blocks alternate between warps 0-3 and 4-7: the kernel is bound by the
// epilogue's latency chain, not by the tensor pipe or HBM); warps 4-7 also load dZ4; 8: MMA issue; 9: weight stages (3 images
// per tile through a 3-stage ring).
struct ChainArgs {
  const float* A;              // dZ4 [R, HP]
  const float* Bimg[3];        // W^T images of hidden layers 3, 2, 1 (32 K values per stage)
  float* Out[3];               // dZ3, dZ2, dZ1
  const uint32_t* bits[3];     // ReLU words of h3, h2, h1
  int R, H;
};
constexpr int CH_BSTAGES = 3;
constexpr int CH_A_BYTES = (PG_NKT - 1) * PG_A_STAGE + 4 * PG_LBO_A;        // the last K stage holds K 192..207: 4 chunk planes
constexpr int CH_STG_BYTES = 8 * 32 * STG_PITCH * 4;                         // one staging tile per epilogue warp (8)
constexpr int CH_SMEM = CH_A_BYTES + CH_BSTAGES * PG_B_STAGE + CH_STG_BYTES + 256;      // 224 320

__global__ void __launch_bounds__(PG_THREADS, 1) upd_dgrad_chain_kernel(ChainArgs a, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                                   // [7 K stages] the tile's A operand (dZ4, then dZ3, then dZ2)
  unsigned char* sB = smem + CH_A_BYTES;
  float* s_stg = (float*)(sB + CH_BSTAGES * PG_B_STAGE);
  uint64_t* bars = (uint64_t*)((unsigned char*)s_stg + CH_STG_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto AL = [&](int k) { return bar0 + 8u * k; };                      // K stage k of dZ4 has landed (128 loader threads)
  auto AE = [&](int k) { return bar0 + 8u * (PG_NKT + k); };           // K stage k written by the epilogue (4 warps)
  auto BFULL = [&](int s) { return bar0 + 8u * (2 * PG_NKT + s); };
  auto BEMPTY = [&](int s) { return bar0 + 8u * (2 * PG_NKT + CH_BSTAGES + s); };
  auto ACCF = [&](int i) { return bar0 + 8u * (2 * PG_NKT + 2 * CH_BSTAGES + i); };
  auto ACCE = [&](int i) { return bar0 + 8u * (2 * PG_NKT + 2 * CH_BSTAGES + 2 + i); };
  const uint32_t AFREE = bar0 + 8u * (2 * PG_NKT + 2 * CH_BSTAGES + 4);   // the tile's last GEMM has read the A operand
  uint32_t* s_tmem = (uint32_t*)(bars + 2 * PG_NKT + 2 * CH_BSTAGES + 5);
  volatile int* s_abort = (volatile int*)(s_tmem + 1);
  const int rpc = (((a.R + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;
  const int r_begin = min(a.R, (int)blockIdx.x * rpc), r_end = min(a.R, r_begin + rpc);
  const int n_tiles = (r_end - r_begin + TM - 1) / TM;

  if (tid == 0) {
    *s_abort = 0;
    for (int k = 0; k < PG_NKT; ++k) { mbar_init(AL(k), 128); mbar_init(AE(k), 4); }
    for (int s = 0; s < CH_BSTAGES; ++s) { mbar_init(BFULL(s), 1); mbar_init(BEMPTY(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(ACCF(i), 1); mbar_init(ACCE(i), 8); }
    mbar_init(AFREE, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 8) {
    // ---- warps 0-7: epilogue (TMEM lane quarter warp & 3; warps 0-3 take the column blocks 0, 2, 4, 6, warps 4-7 the blocks
    // 1, 3, 5); warps 4-7 are also the loaders of dZ4 ----
    const int quad = warp & 3, half = warp >> 2;
    float* stg = s_stg + warp * (32 * STG_PITCH);
    const int cr = lane >> 3, cc = lane & 7;
    const int trow = quad * 32 + lane;                                  // this thread's row of the tile
    const uint32_t a_row = smem_u32(sA) + (trow >> 3) * 128 + (trow & 7) * 16;
    // GEMM g = 3 t + i produces dZ_{3-i}
    auto epilogue = [&](int t, int i) {
      const int g = 3 * t + i, slot = g & 1;
      const int wrow0 = r_begin + t * TM + quad * 32;
      const bool live = wrow0 < r_end;                       // warp-uniform
      uint32_t mbits[8];
      {                                                      // the row's ReLU words, fetched before the accumulator is ready
        const int row = wrow0 + lane;
        uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;      // rows beyond the range: zeros into the next A operand
        if (row < r_end) {
          w0 = *(const uint4*)(a.bits[i] + (size_t)row * 8);
          w1 = *(const uint4*)(a.bits[i] + (size_t)row * 8 + 4);
        }
        mbits[0] = w0.x; mbits[1] = w0.y; mbits[2] = w0.z; mbits[3] = w0.w;
        mbits[4] = w1.x; mbits[5] = w1.y; mbits[6] = w1.z; mbits[7] = w1.w;
      }
      mbar_wait(ACCF(slot), (g >> 1) & 1, s_abort);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + slot * 256;
      float* out = a.Out[i];
#pragma unroll 1
      for (int c0 = 32 * half; c0 < HP; c0 += 64) {
        uint32_t v[32];
        const int nc = HP - c0 >= 32 ? 32 : 16;
        if (nc == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
        tmem_wait_ld();
        if (c0 + 64 >= HP) {                                 // this warp's last block is in registers: hand the slot back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACCE(slot));
        }
        const uint32_t word = mbits[c0 >> 5];
        const uint32_t a_stage = a_row + (c0 >> 5) * PG_A_STAGE;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q * 4 < nc) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = ((word >> (q * 4 + e)) & 1u) ? tf32_rn(__uint_as_float(v[q * 4 + e])) : 0.f;
            *(float4*)(stg + lane * STG_PITCH + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
            if (i < 2)                                       // K chunk q of stage c0 / 32 of the next GEMM's A operand
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_stage + q * PG_LBO_A), "f"(o[0]), "f"(o[1]),
                           "f"(o[2]), "f"(o[3]) : "memory");
          }
        }
        if (i < 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // read by the tensor core (async proxy)
        __syncwarp();
        if (i < 2 && lane == 0) mbar_arrive(AE(c0 >> 5));
        if (live) {
#pragma unroll
          for (int r8 = 0; r8 < 8; ++r8) {
            const int r = r8 * 4 + cr, grow = wrow0 + r;
            if (grow < r_end && cc * 4 < nc)
              *(float4*)(out + (size_t)grow * HP + c0 + cc * 4) = *(const float4*)(stg + r * STG_PITCH + cc * 4);
          }
        }
        __syncwarp();
      }
    };
    // dZ4 of tile t: the whole K extent, one barrier per K stage
    auto load_tile = [&](int t) {
      const int lr = lane >> 3, c = lane & 7;
      const int row0 = r_begin + t * TM;
      for (int kt = 0; kt < PG_NKT; ++kt) {
        const uint32_t dstA = smem_u32(sA + kt * PG_A_STAGE);
        if (kt < PG_NKT - 1 || c < 4) {                      // the last stage has 4 chunk planes
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (quad * 8 + i) * 4 + lr;
            const int grow = row0 + row;
            const bool ok = grow < r_end;
            const float* src = a.A + (size_t)(ok ? grow : 0) * HP + kt * PG_KT + c * 4;
            cp_async16(dstA + c * PG_LBO_A + (row >> 3) * 128 + (row & 7) * 16, src, ok);
          }
        }
        cp_async_arrive(AL(kt));
      }
    };
    if (half == 1 && n_tiles > 0) load_tile(0);
    for (int t = 0; t < n_tiles; ++t) {
      epilogue(t, 0);
      epilogue(t, 1);
      if (half == 1 && t + 1 < n_tiles) {
        mbar_wait(AFREE, t & 1, s_abort);                    // this tile's last GEMM is done with sA
        load_tile(t + 1);
      }
      epilogue(t, 2);
    }
  } else if (warp == 8) {
    // ---- MMA issue ----
    if (lane == 0) {
      int g = 0, gb = 0;
      for (int t = 0; t < n_tiles; ++t) {
        for (int i = 0; i < 3; ++i, ++g) {
          const int slot = g & 1;
          if (g >= 2) { mbar_wait(ACCE(slot), ((g >> 1) - 1) & 1, s_abort); tc_fence_after(); }
          const uint32_t d = tmem + slot * 256;
          for (int kt = 0; kt < PG_NKT; ++kt, ++gb) {
            const int s = gb % CH_BSTAGES, u = gb / CH_BSTAGES;
            // A stage kt: the loaders' (GEMM 0 of the tile, once per tile) or the previous epilogue's (twice per tile)
            if (i == 0) mbar_wait(AL(kt), t & 1, s_abort);
            else mbar_wait(AE(kt), (i - 1) & 1, s_abort);
            mbar_wait(BFULL(s), u & 1, s_abort);
            tc_fence_after();
            const uint32_t aS = smem_u32(sA + kt * PG_A_STAGE), bS = smem_u32(sB + s * PG_B_STAGE);
            const int nks = kt == PG_NKT - 1 ? 2 : 4;       // the last stage holds K 192..207
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t bd = make_desc(bS + ks * 2 * (HP * 16), HP * 16, 128);
              const uint64_t ad = make_desc(aS + ks * 2 * PG_LBO_A, PG_LBO_A, 128);
              mma_ss_tf32(d, ad, bd, IDESC_K, (kt | ks) != 0);
            }
            tc_commit(BEMPTY(s));
          }
          tc_commit(ACCF(slot));
          if (i == 2) tc_commit(AFREE);
        }
      }
    }
    __syncwarp();
  } else {
    // ---- weight stages: the three W^T images per tile ----
    if (lane == 0) {
      int gb = 0;
      for (int t = 0; t < n_tiles; ++t) {
        for (int i = 0; i < 3; ++i) {
          for (int kt = 0; kt < PG_NKT; ++kt, ++gb) {
            const int s = gb % CH_BSTAGES, u = gb / CH_BSTAGES;
            if (u > 0) mbar_wait(BEMPTY(s), (u - 1) & 1, s_abort);
            mbar_expect_tx(BFULL(s), PG_B_STAGE);
            bulk_g2s(smem_u32(sB + s * PG_B_STAGE), a.Bimg[i] + (size_t)kt * (HP * PG_KT), PG_B_STAGE, BFULL(s));
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && *s_abort) atomicOr(err, 1);
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- dZ4 and the head's weight gradient -------------------------------------------------------------------------------
__global__ void __launch_bounds__(THIN_THREADS) upd_dz4_kernel(int R, int H, const float* __restrict__ H4,
                                                                const float* __restrict__ dl, const float* __restrict__ w4,
                                                                float* __restrict__ dZ4, float* __restrict__ part_head) {
  __shared__ float s_acc[4][HP];
  const int cg = threadIdx.x % 52, rsub = threadIdx.x / 52;
  float w[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = cg * 4 + e < H ? w4[cg * 4 + e] : 0.f;
  const int rpc = (R + gridDim.x - 1) / gridDim.x;
  const int r_end = min(R, (int)(blockIdx.x + 1) * rpc);
  constexpr int U = 4;                                   // rows in flight per thread
  for (int r0 = blockIdx.x * rpc + rsub; r0 < r_end; r0 += 4 * U) {
    float4 h[U];
    float d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + 4 * u;
      h[u] = r < r_end ? *(const float4*)(H4 + (size_t)r * HP + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      d[u] = r < r_end ? __ldg(dl + r) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + 4 * u;
      const float hv[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[e] = fmaf(d[u], hv[e], acc[e]);                     // column H is the ones column: db4
        o[e] = (hv[e] > 0.f && cg * 4 + e < H) ? tf32_rn(d[u] * w[e]) : 0.f;
      }
      if (r < r_end) *(float4*)(dZ4 + (size_t)r * HP + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) s_acc[rsub][cg * 4 + e] = acc[e];
  __syncthreads();
  const int j = threadIdx.x;
  part_head[(size_t)blockIdx.x * HP + j] = (s_acc[0][j] + s_acc[1][j]) + (s_acc[2][j] + s_acc[3][j]);
}

// ---- layer-0 weight gradient: dW0[o][f] = sum_r dZ1[r][o] x[r][f], db0[o] = sum_r dZ1[r][o] ----------------------------
__global__ void __launch_bounds__(THIN_THREADS) upd_wgrad0_kernel(Net net, const float* __restrict__ states, int R,
                                                                   const float* __restrict__ dH1, const float* __restrict__ H1,
                                                                   float* __restrict__ part_l0) {
  __shared__ float s_acc[3][HP * (MAXF + 1)];
  const int cg = threadIdx.x % 52, rsub = threadIdx.x / 52;
  const int F = net.F;
  float acc[4][MAXF + 1];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int f = 0; f <= MAXF; ++f) acc[e][f] = 0.f;
  const int rpc = (R + gridDim.x - 1) / gridDim.x;
  const int r_end = min(R, (int)(blockIdx.x + 1) * rpc);
  constexpr int U = 4;                                   // rows in flight per thread
  for (int r0 = blockIdx.x * rpc + rsub; r0 < r_end; r0 += 4 * U) {
    float4 d[U];
    float x[U][MAXF + 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + 4 * u;
      const bool ok = r < r_end;
      d[u] = ok ? *(const float4*)(dH1 + (size_t)r * HP + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (H1) {                                    // SS-forward variant: the dgrad stored dH1, dZ1 = dH1 . [h1 > 0] here
        const float4 h = ok ? *(const float4*)(H1 + (size_t)r * HP + cg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[u].x = h.x > 0.f ? d[u].x : 0.f;
        d[u].y = h.y > 0.f ? d[u].y : 0.f;
        d[u].z = h.z > 0.f ? d[u].z : 0.f;
        d[u].w = h.w > 0.f ? d[u].w : 0.f;
      }
#pragma unroll
      for (int f = 0; f < MAXF; ++f) x[u][f] = (ok && f < F) ? __ldg(states + (size_t)r * net.F_state + net.col[f]) : 0.f;
      x[u][MAXF] = 1.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int f = 0; f <= MAXF; ++f) acc[e][f] = fmaf(dv[e], x[u][f], acc[e][f]);
    }
  }
  // fixed-order reduction over the 4 row lanes: lanes 1..3 park their sums, lane 0 adds them in order
  if (rsub > 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int f = 0; f <= MAXF; ++f) s_acc[rsub - 1][(cg * 4 + e) * (MAXF + 1) + f] = acc[e][f];
  }
  __syncthreads();
  if (rsub == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int f = 0; f <= MAXF; ++f) {
        const int i = (cg * 4 + e) * (MAXF + 1) + f;
        part_l0[(size_t)blockIdx.x * HP * (MAXF + 1) + i] = ((acc[e][f] + s_acc[0][i]) + s_acc[1][i]) + s_acc[2][i];
      }
  }
}

// ---- wgrad of the three H x H layers: dWb_l[o][i] = sum_r dZ_{l+1}[r][o] * [h_l, 1][r][i] -------------------------------
// K = rows: both operands are MN-major, and for tf32 the only MN-major shared-memory layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom; a no-swizzle MN-major tf32 descriptor reads zeros):
// blocks of 32 columns x 32 rows, row pitch 128 B, atoms of 4 rows (512 B) whose 32-byte chunks are XOR-permuted by
// the row index (address bits 5-6 ^= bits 7-8).  LBO = distance of 32-column blocks (4096 B), SBO = distance of 4-row
// atoms (512 B); one K = 8 MMA reads two atoms, the next K step starts 1024 B further.
struct WgradArgs {
  const float* dZ[3];    // dZ_{l+1}, l = 1..3
  const float* Hin[3];   // h_l
  float* part;           // [WG_SPLIT][3][HP][HP]
  int R, H;
};

__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}

__global__ void __launch_bounds__(GEMM_THREADS) upd_wgrad_kernel(WgradArgs a, int* __restrict__ err) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // the swizzle works on address bits
  unsigned char* sA = smem;
  unsigned char* sB = smem + WG_STAGES * WG_A_STAGE;
  uint64_t* bars = (uint64_t*)(smem + WG_STAGES * (WG_A_STAGE + WG_B_STAGE));
  uint32_t* s_tmem = (uint32_t*)(bars + 8);
  volatile int* s_abort = (volatile int*)(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, li = blockIdx.y;
  const int nk_total = (a.R + WG_KT - 1) / WG_KT;
  const int per = (nk_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int kt0 = split * per, kt1 = min(nk_total, kt0 + per);
  const int nk = max(0, kt1 - kt0);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (WG_STAGES + s); };
  const uint32_t ACCF = bar0 + 8u * (2 * WG_STAGES);

  if (tid == 0) {
    *s_abort = 0;
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(FULL(s), 128); mbar_init(EMPTY(s), 1); }
    mbar_init(ACCF, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // column block 7 (columns 224..255) of the A operand is never loaded: zero it once
  for (int s = 0; s < WG_STAGES; ++s)
    for (int i = tid; i < 4096 / 16; i += GEMM_THREADS)
      *(float4*)(sA + s * WG_A_STAGE + 7 * 4096 + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // ---- loaders: per operand 7 column blocks x 8 four-row atoms per stage; one warp instruction = one atom of one
    // block: lane -> (16-byte chunk jj of the 128-byte row, row rr of the atom): full 128-byte lines from global ----
    const int jj = lane & 7, rr = lane >> 3;
    const float* gA = a.dZ[li];
    const float* gB = a.Hin[li];
    const uint32_t sw_off = (uint32_t)((((jj >> 1) ^ rr) << 5) | ((jj & 1) << 4)) + rr * 128;   // swizzled offset inside the atom
    for (int it = 0; it < nk; ++it) {
      const int s = it % WG_STAGES, u = it / WG_STAGES;
      if (u > 0) mbar_wait(EMPTY(s), (u - 1) & 1, s_abort);
      const uint32_t dA = smem_u32(sA + s * WG_A_STAGE), dB = smem_u32(sB + s * WG_B_STAGE);
      const int krow0 = (kt0 + it) * WG_KT;
#pragma unroll 4
      for (int i = 0; i < 28; ++i) {
        int q = warp * 28 + i;                     // 0..111: operand (2) x column block (7) x atom (8)
        const bool isB = q >= 56;
        q -= isB ? 56 : 0;
        const int blk = q >> 3, atom = q & 7;
        const int col = blk * 32 + jj * 4;
        const int grow = krow0 + atom * 4 + rr;
        const bool ok = grow < a.R && col < HP;
        const float* src = (isB ? gB : gA) + (size_t)(ok ? grow : 0) * HP + (ok ? col : 0);
        cp_async16((isB ? dB : dA) + blk * 4096 + atom * 512 + sw_off, src, ok);
      }
      cp_async_arrive(FULL(s));
    }
    // ---- epilogue: two M blocks of 128 output rows o; thread = o; row-contiguous store of the partial ----
    if (nk > 0) { mbar_wait(ACCF, 0, s_abort); tc_fence_after(); }
    float* part = a.part + ((size_t)split * 3 + li) * HP * HP;
#pragma unroll 1
    for (int mb = 0; mb < 2; ++mb) {
      const int o = mb * 128 + warp * 32 + lane;
      const bool ok = o < a.H;
      const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + mb * 256;
#pragma unroll 1
      for (int c0 = 0; c0 < HP; c0 += 32) {
        uint32_t v[32];
        const int nc = HP - c0 >= 32 ? 32 : 16;
        if (nk > 0) {
          if (nc == 32) tmem_ld32(tbase + c0, v); else tmem_ld16(tbase + c0, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
        if (ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q * 4 < nc)
              *(uint4*)(part + (size_t)o * HP + c0 + q * 4) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % WG_STAGES, u = it / WG_STAGES;
        mbar_wait(FULL(s), u & 1, s_abort);
        tc_fence_after();
        const uint32_t aS = smem_u32(sA + s * WG_A_STAGE), bS = smem_u32(sB + s * WG_B_STAGE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = make_desc_mn32(bS + ks * 1024);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const uint64_t ad = make_desc_mn32(aS + mb * (4 * 4096) + ks * 1024);
            mma_ss_tf32(tmem + mb * 256, ad, bd, IDESC_MN, (it | ks) != 0);
          }
        }
        tc_commit(EMPTY(s));
      }
      if (nk > 0) tc_commit(ACCF);
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0 && *s_abort) atomicOr(err, 1);
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- fixed-order reduction of every partial ------------------------------------------------------------------------------
// blocks [0, nblkA): the three H x (H + 1) layers, one thread per output, 49 split-K partials;
// then one WARP per output of the head (H + 1) and of layer 0 (H x (F + 1)): lanes stride over the n_thin CTA partials,
// xor-tree combine (a fixed order: bitwise reproducible).
__global__ void __launch_bounds__(256) upd_reduce_kernel(Net net, Ws ws, int n_head, int n_l0, int nblkA) {
  const int H = net.H, F = net.F;
  if ((int)blockIdx.x < nblkA) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= 3 * H * (H + 1)) return;
    const int l = idx / (H * (H + 1)), r = idx - l * H * (H + 1);
    const int o = r / (H + 1), i = r - o * (H + 1);
    const float* p = ws.part_big + ((size_t)l * HP + o) * HP + i;
    float s = 0.f;
#pragma unroll 7
    for (int sp = 0; sp < WG_SPLIT; ++sp) s += p[(size_t)sp * 3 * HP * HP];
    if (i < H) net.dW[l + 1][(size_t)o * H + i] = s; else net.db[l + 1][o] = s;
    return;
  }
  const int lane = threadIdx.x & 31;
  int idx = ((int)blockIdx.x - nblkA) * 8 + (threadIdx.x >> 5);
  const int nB = H + 1, nC = H * (F + 1);
  if (idx >= nB + nC) return;
  const float* p;
  size_t stride;
  float* out;
  int n_thin = n_head;
  if (idx < nB) {
    p = ws.part_head + idx; stride = HP;
    out = idx < H ? net.dW[4] + idx : net.db[4];
  } else {
    idx -= nB;
    n_thin = n_l0;
    const int o = idx / (F + 1), f = idx - o * (F + 1);
    p = ws.part_l0 + (size_t)o * (MAXF + 1) + (f < F ? f : MAXF); stride = (size_t)HP * (MAXF + 1);
    out = f < F ? net.dW[0] + (size_t)o * F + f : net.db[0] + o;
  }
  float s = 0.f;
  for (int c = lane; c < n_thin; c += 32) s += p[(size_t)c * stride];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) *out = s;
}

// ---- Adam (torch.optim.Adam defaults of ppo.py:93) over every parameter tensor in one launch -------------------------
constexpr int ADAM_MAX_TENSORS = 32;
struct AdamArgs {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long end[ADAM_MAX_TENSORS];     // exclusive prefix end of tensor t in the flat index space
  long long* st[ADAM_MAX_TENSORS];     // per-tensor state {step, 1 - beta1^step, 1 - beta2^step}
  int n;
};
// per-tensor state = {step (int64), 1 - beta1^step (double), 1 - beta2^step (double)} on the device (torch keeps one step
// per parameter: a tensor without a gradient does not advance); a replayed CUDA graph advances it
__global__ void adam_tick_kernel(AdamArgs a, double beta1, double beta2) {
  const int i = threadIdx.x;
  if (i >= a.n) return;
  long long* state = a.st[i];
  const long long t = state[0] + 1;
  state[0] = t;
  ((double*)state)[1] = 1.0 - pow(beta1, (double)t);
  ((double*)state)[2] = 1.0 - pow(beta2, (double)t);
}
__global__ void __launch_bounds__(256) adam_step_kernel(AdamArgs a, float lr, float beta1, float beta2, float eps) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.end[a.n - 1]) return;
  int t = 0;
  while (i >= a.end[t]) ++t;
  const long long j = i - (t ? a.end[t - 1] : 0);
  const double* state = (const double*)a.st[t];
  const float bc1 = (float)state[1], bc2s = (float)sqrt(state[2]);
  const float g = a.g[t][j];
  const float m = a.m[t][j] + (g - a.m[t][j]) * (1.f - beta1);        // exp_avg.lerp_(grad, 1 - beta1)
  const float v = a.v[t][j] * beta2 + (1.f - beta2) * g * g;
  a.m[t][j] = m;
  a.v[t][j] = v;
  a.p[t][j] -= (lr / bc1) * m / (sqrtf(v) / bc2s + eps);
}

static bool g_attr_done = false;
static int g_num_sms = 148;
// second stream of a policy-gradient call: the CUDA-core layer-0 wgrad runs beside the tensor-core wgrad of the hidden
// layers (fork / join by events; inside a stream capture the two become parallel branches of the graph).  Created on the
// first call, which must not be a captured one (PPO runs its first optimiser steps eagerly).
static cudaStream_t g_side = nullptr;
static cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr, g_ev_fork0 = nullptr, g_ev_pack = nullptr;
static int set_attrs() {
  if (g_attr_done) return MBO_OK;
  MBO_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
  MBO_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
  MBO_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
  MBO_CUDA(cudaEventCreateWithFlags(&g_ev_fork0, cudaEventDisableTiming));
  MBO_CUDA(cudaEventCreateWithFlags(&g_ev_pack, cudaEventDisableTiming));
  MBO_CUDA(cudaFuncSetAttribute(upd_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_dgrad_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_dgrad_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_ts_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
  MBO_CUDA(cudaFuncSetAttribute(upd_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); }
  g_attr_done = true;
  return MBO_OK;
}

}  // namespace upd
}  // namespace mbo

using namespace mbo;
using namespace mbo::upd;

extern "C" size_t mbo_ppo_workspace_bytes(int E, int M) {
  if (E <= 0 || M <= 0) return 0;
  return carve(E, M, nullptr, nullptr);
}

extern "C" int mbo_ppo_debug_offsets(int E, int M, size_t* off /*[12]*/) {
  MBO_CHECK_ARG(E > 0 && M > 0 && off, "mbo_ppo_debug_offsets: bad arguments");
  Ws w;
  unsigned char* base = (unsigned char*)4096;     // any non-null base: only differences are reported
  carve(E, M, base, &w);
  for (int l = 0; l < 4; ++l) off[l] = (unsigned char*)w.Hs[l] - base;
  for (int l = 0; l < 4; ++l) off[4 + l] = (unsigned char*)w.dZ[l] - base;
  off[8] = (unsigned char*)w.logits - base;
  off[9] = (unsigned char*)w.dl - base;
  off[10] = (unsigned char*)w.terms - base;
  off[11] = (unsigned char*)w.part_big - base;
  return MBO_OK;
}

extern "C" int mbo_ppo_policy_grad(int E, int M, int F_state, int F, const int32_t* feat_col, int H,
                                   const float* states, const float* const* W, const float* const* b,
                                   const int32_t* actions, const float* advs, const float* logp_old, float epsilon,
                                   float ent_coeff, float inv_n_global, float* const* dW, float* const* db,
                                   float* terms, void* workspace, size_t workspace_bytes, void* stream) {
  MBO_CHECK_ARG(E > 0 && M > 0, "mbo_ppo_policy_grad: E, M must be > 0");
  MBO_CHECK_ARG((size_t)E * M < (1u << 30), "mbo_ppo_policy_grad: too many rows");
  MBO_CHECK_ARG(F >= 1 && F <= MAXF && F_state >= F, "mbo_ppo_policy_grad: 1 <= F <= %d policy features of F_state", MAXF);
  MBO_CHECK_ARG(H >= 8 && H <= HP - 1, "mbo_ppo_policy_grad: hidden width must be in 8..%d", HP - 1);
  const bool fwd_only = dW == nullptr;      // log-probs / entropies under these weights only (the frozen old policy)
  MBO_CHECK_ARG(states && W && b && actions && terms && workspace && feat_col, "mbo_ppo_policy_grad: null pointer");
  MBO_CHECK_ARG(fwd_only || (advs && logp_old && db), "mbo_ppo_policy_grad: null pointer");
  MBO_CHECK_ARG(workspace_bytes >= mbo_ppo_workspace_bytes(E, M), "mbo_ppo_policy_grad: workspace too small");
  MBO_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "mbo_ppo_policy_grad: workspace must be 256-byte aligned");
  if (set_attrs() != MBO_OK) return MBO_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  Net net;
  for (int l = 0; l < 5; ++l) {
    MBO_CHECK_ARG(W[l] && b[l] && (fwd_only || (dW[l] && db[l])), "mbo_ppo_policy_grad: null layer pointer");
    net.W[l] = W[l]; net.b[l] = b[l]; net.dW[l] = fwd_only ? nullptr : dW[l]; net.db[l] = fwd_only ? nullptr : db[l];
  }
  net.F = F; net.F_state = F_state; net.H = H;
  for (int f = 0; f < MAXF; ++f) {
    net.col[f] = f < F ? feat_col[f] : 0;
    MBO_CHECK_ARG(net.col[f] >= 0 && net.col[f] < F_state, "mbo_ppo_policy_grad: feature column out of range");
  }
  Ws ws;
  carve(E, M, (unsigned char*)workspace, &ws);
  const int R = E * M;
  const int tiles = (R + TM * GEMM_MB - 1) / (TM * GEMM_MB);
  const char* dbg_env = getenv("MBO_UPD_DBG");
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  // which GEMM variant runs (both are kept and tested against each other): measured on the C3 minibatch, the persistent
  // double-buffered kernel wins for the dgrad (33.5 vs 38 us) and loses for the 3xTF32 forward (81 vs 75 us)
  const char* l0_env = getenv("MBO_L0_CTAS");
  const char* w0_env = getenv("MBO_WG0_CTAS");
  const int l0_ctas = l0_env ? atoi(l0_env) : 4 * THIN_CTAS;
  const int wg0_ctas = w0_env ? (atoi(w0_env) < THIN_CTAS ? atoi(w0_env) : THIN_CTAS) : WG0_CTAS;
  // which dgrad kernel runs (both are kept and tested bitwise against each other): on the C3 minibatch the persistent
  // double-buffered kernel takes 33.5 us, the one-tile-per-CTA kernel 38 us
  const char* ts_env = getenv("MBO_UPD_FWD_TS");
  const bool fwd_ts = ts_env ? atoi(ts_env) != 0 : true;      // A operand in TMEM (default) or in shared memory (SS form)
  const char* pd_env = getenv("MBO_UPD_PERSIST_DGRAD");
  const bool persist = pd_env ? atoi(pd_env) != 0 : true;
  const char* dt_env = getenv("MBO_UPD_DGRAD_TS");
  const bool dgrad_ts = dt_env ? atoi(dt_env) != 0 : false;
  const int pg_grid = g_num_sms < tiles ? g_num_sms : tiles;

  // Forward (policies.py:104-125 on the minibatch).  Default: ONE launch of the tcgen05 policy kernel of the rollouts
  // (neural_af_tc.cu, bf16x3 = fp32-class arithmetic, all four layers with the activations resident in tensor memory)
  // on the current weights, which also writes the fp32 activations h1..h4 and the ReLU bit words the backward reads --
  // instead of layer 0 + three GEMM launches whose activations made an HBM round trip each.  MBO_UPD_FWD_K2=0 selects
  // the GEMM-per-layer forward (3xTF32), kept as the A/B reference.
  const char* k2_env = getenv("MBO_UPD_FWD_K2");
  const bool fwd_k2 = k2_env ? atoi(k2_env) != 0 : true;
  const char* fk_env = getenv("MBO_UPD_FORK");
  const bool fork = fk_env ? atoi(fk_env) != 0 : true;
  // weight images of the backward GEMMs (W^T, tf32): needed by the first dgrad only -- packed on the second stream while
  // the forward runs; the GEMM-per-layer forward also needs the [W | b] hi / lo images, up front
  const int pack_kinds = (fwd_k2 ? 0 : 1) | (fwd_only ? 0 : ((persist && !dgrad_ts) ? 4 : 2));
  const bool pack_aside = fork && fwd_k2 && !fwd_only;
  if (pack_aside) {
    MBO_CUDA(cudaEventRecord(g_ev_fork0, st));
    MBO_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork0, 0));
    upd_pack_kernel<<<(9 * IMG_FLOATS + 255) / 256, 256, 0, g_side>>>(net, ws, pack_kinds);
    MBO_CUDA(cudaEventRecord(g_ev_pack, g_side));
  } else if (pack_kinds) {
    upd_pack_kernel<<<(9 * IMG_FLOATS + 255) / 256, 256, 0, st>>>(net, ws, pack_kinds);
  }
  if (fwd_k2) {
    int rc = tc_pack_raw_bf16x3(W, b, F, H, ws.tcpack, st);
    if (rc != MBO_OK) return rc;
    rc = tc_forward_dense_bf16x3(ws.tcpack, E, M, F_state, F, feat_col, states, ws.logits, ws.err, fwd_only ? nullptr : ws.Hs,
                                 ws.hbits, H, st);
    if (rc != MBO_OK) return rc;
  } else {
  upd_layer0_kernel<<<l0_ctas, THIN_THREADS, 0, st>>>(net, states, R, ws.Hs[0]);
  for (int l = 0; l < 3; ++l) {
    GemmArgs g{};
    g.A = ws.Hs[l]; g.Bimg = ws.img_fwd[l]; g.Bimg_lo = ws.img_fwl[l]; g.Out = ws.Hs[l + 1];
    g.bits_out = l < 2 ? ws.hbits[l + 1] : nullptr; g.bits_a = (fwd_ts && l == 0) ? ws.hbits[0] : nullptr;
    g.R = R; g.H = H; g.dbg = dbg;
    if (l == 2) { g.w4 = W[4]; g.b4 = b[4]; g.logits = ws.logits; }
    if (fwd_ts) upd_ts_kernel<0><<<tiles, GEMM_THREADS, TS_SMEM, st>>>(g, ws.err);
    else upd_gemm_kernel<0><<<tiles, GEMM_THREADS, FWD_SMEM, st>>>(g, ws.err);
  }
  }
  upd_loss_kernel<<<E, 256, 0, st>>>(M, ws.logits, actions, advs, logp_old, epsilon, ent_coeff, inv_n_global, ws.dl,
                                      ws.terms);
  if (fwd_only) {
    MBO_CUDA(cudaMemcpyAsync(terms, ws.terms, (size_t)E * 16, cudaMemcpyDeviceToDevice, st));
    MBO_LAUNCH_CHECK();
    return MBO_OK;
  }
  upd_dz4_kernel<<<THIN_CTAS, THIN_THREADS, 0, st>>>(R, H, ws.Hs[3], ws.dl, W[4], ws.dZ[3], ws.part_head);
  if (pack_aside) MBO_CUDA(cudaStreamWaitEvent(st, g_ev_pack, 0));
  // Default: the three dgrad GEMMs as ONE persistent kernel (dZ3 / dZ2 go from the epilogue straight into the next GEMM's
  // shared-memory operand); MBO_UPD_DGRAD_CHAIN=0: one launch per layer (kept: bitwise equal, the A/B reference)
  const char* ch_env = getenv("MBO_UPD_DGRAD_CHAIN");
  const bool chain = (ch_env ? atoi(ch_env) != 0 : true) && persist && !dgrad_ts && (fwd_ts || fwd_k2);
  if (chain) {
    ChainArgs c{};
    c.A = ws.dZ[3];
    for (int i = 0; i < 3; ++i) { c.Bimg[i] = ws.img_dgp[2 - i]; c.Out[i] = ws.dZ[2 - i]; c.bits[i] = ws.hbits[2 - i]; }
    c.R = R; c.H = H;
    upd_dgrad_chain_kernel<<<pg_grid, PG_THREADS, CH_SMEM, st>>>(c, ws.err);
  }
  for (int l = chain ? -1 : 2; l >= 0; --l) {      // dZ_{l+1} = (dZ_{l+2} W_{l+1}) . [h_{l+1} > 0]
    GemmArgs g{};
    g.A = ws.dZ[l + 1]; g.Bimg = (persist && !dgrad_ts) ? ws.img_dgp[l] : ws.img_dgr[l]; g.Out = ws.dZ[l]; g.bits_in = (l > 0 || fwd_ts || fwd_k2) ? ws.hbits[l] : nullptr;   // SS forward: no h1 words, l = 0 stores dH1, masked in wgrad0
    g.R = R; g.H = H; g.dbg = dbg;
    if (dgrad_ts) upd_ts_kernel<1><<<tiles, GEMM_THREADS, TS_SMEM, st>>>(g, ws.err);
    else if (persist) upd_dgrad_persist_kernel<<<pg_grid, PG_THREADS, PG_SMEM, st>>>(g, ws.err);
    else upd_gemm_kernel<1><<<tiles, GEMM_THREADS, GEMM_SMEM, st>>>(g, ws.err);
  }
  WgradArgs wa{};
  for (int l = 0; l < 3; ++l) { wa.dZ[l] = ws.dZ[l + 1]; wa.Hin[l] = ws.Hs[l]; }
  wa.part = ws.part_big; wa.R = R; wa.H = H;

  cudaStream_t st0 = st;                    // stream of the layer-0 wgrad
  if (fork) {
    MBO_CUDA(cudaEventRecord(g_ev_fork, st));
    MBO_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork, 0));
    st0 = g_side;
  }
  upd_wgrad_kernel<<<dim3(WG_SPLIT, 3), GEMM_THREADS, WG_SMEM, st>>>(wa, ws.err);
  upd_wgrad0_kernel<<<wg0_ctas, THIN_THREADS, 0, st0>>>(net, states, R, ws.dZ[0], (fwd_ts || fwd_k2) ? nullptr : ws.Hs[0], ws.part_l0);
  if (fork) {
    MBO_CUDA(cudaEventRecord(g_ev_join, g_side));
    MBO_CUDA(cudaStreamWaitEvent(st, g_ev_join, 0));
  }
  const int nblkA = (3 * H * (H + 1) + 255) / 256, nblkT = ((H + 1) + H * (F + 1) + 7) / 8;
  upd_reduce_kernel<<<nblkA + nblkT, 256, 0, st>>>(net, ws, THIN_CTAS, wg0_ctas, nblkA);
  MBO_CUDA(cudaMemcpyAsync(terms, ws.terms, (size_t)E * 16, cudaMemcpyDeviceToDevice, st));
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_adam_step(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                             const int64_t* sizes, int64_t* const* state, float lr, float beta1, float beta2, float eps,
                             void* stream) {
  MBO_CHECK_ARG(n >= 1 && n <= ADAM_MAX_TENSORS, "mbo_adam_step: 1..%d tensors per call", ADAM_MAX_TENSORS);
  MBO_CHECK_ARG(p && g && m && v && sizes && state, "mbo_adam_step: null pointer");
  AdamArgs a;
  long long tot = 0;
  for (int t = 0; t < n; ++t) {
    MBO_CHECK_ARG(p[t] && g[t] && m[t] && v[t] && state[t] && sizes[t] > 0, "mbo_adam_step: null / empty tensor %d", t);
    a.p[t] = p[t]; a.g[t] = g[t]; a.m[t] = m[t]; a.v[t] = v[t]; a.st[t] = (long long*)state[t];
    tot += sizes[t];
    a.end[t] = tot;
  }
  a.n = n;
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, ADAM_MAX_TENSORS, 0, st>>>(a, (double)beta1, (double)beta2);
  adam_step_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, lr, beta1, beta2, eps);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

// ======================================================================================================================
// Value-net half of the optimiser step (ppo.py:150-156: loss_value = mean((vpreds - tdlamrets)^2), policies.py:116-121:
// value net on states[:, 0, [t_idx, T_idx]]).  E rows x (2 -> Hv -> Hv -> Hv -> Hv -> 1): 40 MFLOP, CUDA cores, fp32,
// 15 small launches (forward x5, loss, wgrad x5, dgrad x4) instead of ~45 torch kernels.
// ======================================================================================================================
namespace mbo {
namespace upd {

constexpr int V_MAXH = 256;

// out[e][j] = act(b[j] + sum_k in[e][k] W[j][k]).  One warp per (4 rows, 1 output column): lanes split K (coalesced
// reads of the W row and of the activation rows), xor-tree sum.  Layer 0 reads its two inputs from the state tensor
// (row 0 of every episode's state matrix, columns c0 / c1).
__global__ void __launch_bounds__(256) v_fwd_kernel(int E, int Kin, int Hout, const float* __restrict__ in, size_t in_stride,
                                                     int c0, int c1, const float* __restrict__ W, const float* __restrict__ b,
                                                     float* __restrict__ out, int relu) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_eg = (E + 3) >> 2;
  if (warp >= n_eg * Hout) return;
  const int j = warp / n_eg, e0 = (warp - j * n_eg) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 >= 0) {
    if (lane < 2) {
      const float w = W[(size_t)j * 2 + lane];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (e0 + r < E) acc[r] = in[(size_t)(e0 + r) * in_stride + (lane ? c1 : c0)] * w;
    }
  } else {
    for (int k = lane; k < Kin; k += 32) {
      const float w = W[(size_t)j * Kin + k];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (e0 + r < E) acc[r] = fmaf(in[(size_t)(e0 + r) * in_stride + k], w, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  if (lane < 4 && e0 + lane < E) {
    const float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + b[j];
    out[(size_t)(e0 + lane) * Hout + j] = relu ? fmaxf(v, 0.f) : v;
  }
}

// dv[e] = 2 (v[e] - ret[e]) * coeff (coeff = value_coeff / n_global), vterms[e] = (v - ret)^2
__global__ void v_loss_kernel(int E, const float* __restrict__ v, const float* __restrict__ ret, float coeff,
                              float* __restrict__ dv, float* __restrict__ values, float* __restrict__ vterms) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float d = v[e] - ret[e];
  dv[e] = 2.f * d * coeff;
  values[e] = v[e];
  vterms[e] = d * d;
}

// dIn[e][i] = [h[e][i] > 0] * sum_o dZ[e][o] W[o][i].  CTA = (32 input columns, 8 rows): the W column block and the 8
// dZ rows are staged in shared memory with every thread loading (the first version walked W with one L2 round trip per
// 8 outputs and took 11 us); warp = row, lane = column.
__global__ void __launch_bounds__(256) v_dgrad_kernel(int E, int Kin, int Hout, const float* __restrict__ dZ,
                                                       const float* __restrict__ W, const float* __restrict__ h,
                                                       float* __restrict__ dIn) {
  __shared__ float Ws[V_MAXH][32];
  __shared__ float Zs[8][V_MAXH];
  const int i0 = blockIdx.x * 32, e0 = blockIdx.y * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < Hout * 32; t += 256) {
    const int o = t >> 5, ii = t & 31;
    Ws[o][ii] = i0 + ii < Kin ? W[(size_t)o * Kin + i0 + ii] : 0.f;
  }
  for (int t = threadIdx.x; t < 8 * Hout; t += 256) {
    const int r = t / Hout, o = t - r * Hout;
    Zs[r][o] = e0 + r < E ? dZ[(size_t)(e0 + r) * Hout + o] : 0.f;
  }
  __syncthreads();
  const int e = e0 + warp, i = i0 + lane;
  if (e >= E || i >= Kin) return;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  int o = 0;
  for (; o + 3 < Hout; o += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = fmaf(Zs[warp][o + u], Ws[o + u][lane], a[u]);
  }
  for (; o < Hout; ++o) a[0] = fmaf(Zs[warp][o], Ws[o][lane], a[0]);
  dIn[(size_t)e * Kin + i] = h[(size_t)e * Kin + i] > 0.f ? (a[0] + a[1]) + (a[2] + a[3]) : 0.f;
}

// dW[o][i] = sum_e dZ[e][o] in[e][i], db[o] = sum_e dZ[e][o]; one thread per (o, i), four independent chains over e
__global__ void __launch_bounds__(256) v_wgrad_kernel(int E, int Kin, int Hout, const float* __restrict__ dZ,
                                                       const float* __restrict__ in, size_t in_stride, int c0, int c1,
                                                       float* __restrict__ dW, float* __restrict__ db) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= Hout * Kin) return;
  const int o = idx / Kin, i = idx - o * Kin;
  const int col = c0 >= 0 ? (i == 0 ? c0 : c1) : i;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
  int e = 0;
  for (; e + 3 < E; e += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d = dZ[(size_t)(e + u) * Hout + o];
      acc[u] = fmaf(d, in[(size_t)(e + u) * in_stride + col], acc[u]);
      accb[u] += d;
    }
  }
  for (; e < E; ++e) {
    const float d = dZ[(size_t)e * Hout + o];
    acc[0] = fmaf(d, in[(size_t)e * in_stride + col], acc[0]);
    accb[0] += d;
  }
  dW[idx] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  if (i == 0) db[o] = (accb[0] + accb[1]) + (accb[2] + accb[3]);
}

}  // namespace upd
}  // namespace mbo

extern "C" size_t mbo_ppo_value_workspace_bytes(int E, int Hv) {
  if (E <= 0 || Hv <= 0) return 0;
  return (size_t)6 * E * Hv * 4 + 2 * (size_t)E * 4 + 1024;
}

extern "C" int mbo_ppo_value_grad(int E, size_t state_stride, int t_col, int T_col, int Hv, const float* states,
                                  const float* const* W, const float* const* b, const float* tdlamrets, float coeff,
                                  float* const* dW, float* const* db, float* values, float* vterms, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  MBO_CHECK_ARG(E > 0 && Hv >= 1 && Hv <= V_MAXH, "mbo_ppo_value_grad: E > 0, hidden width 1..%d", V_MAXH);
  MBO_CHECK_ARG(states && W && b && tdlamrets && dW && db && values && vterms && workspace, "mbo_ppo_value_grad: null pointer");
  MBO_CHECK_ARG(t_col >= 0 && T_col >= 0 && (size_t)t_col < state_stride && (size_t)T_col < state_stride,
                "mbo_ppo_value_grad: t / T column outside the state row");
  MBO_CHECK_ARG(workspace_bytes >= mbo_ppo_value_workspace_bytes(E, Hv), "mbo_ppo_value_grad: workspace too small");
  for (int l = 0; l < 5; ++l) MBO_CHECK_ARG(W[l] && b[l] && dW[l] && db[l], "mbo_ppo_value_grad: null layer pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float* base = (float*)workspace;
  float* h[4];
  for (int l = 0; l < 4; ++l) h[l] = base + (size_t)l * E * Hv;
  float* dz[2] = {base + (size_t)4 * E * Hv, base + (size_t)5 * E * Hv};
  float* v = base + (size_t)6 * E * Hv;
  float* dv = v + E;
  using namespace mbo::upd;
  const int n_eg4 = (E + 3) / 4;
  const int gF = (n_eg4 * Hv + 7) / 8, gF1 = (n_eg4 + 7) / 8;             // 8 warps per CTA, one warp per (4 rows, column)
  const dim3 gD((Hv + 31) / 32, (E + 7) / 8);                             // CTA = (32 columns, 8 rows)
  v_fwd_kernel<<<gF, 256, 0, st>>>(E, 2, Hv, states, state_stride, t_col, T_col, W[0], b[0], h[0], 1);
  for (int l = 1; l < 4; ++l) v_fwd_kernel<<<gF, 256, 0, st>>>(E, Hv, Hv, h[l - 1], (size_t)Hv, -1, -1, W[l], b[l], h[l], 1);
  v_fwd_kernel<<<gF1, 256, 0, st>>>(E, Hv, 1, h[3], (size_t)Hv, -1, -1, W[4], b[4], v, 0);
  v_loss_kernel<<<(E + 127) / 128, 128, 0, st>>>(E, v, tdlamrets, coeff, dv, values, vterms);
  // head: dW4 = dv^T h4, dz(h4) = dv (x) w4 . [h4 > 0]
  v_wgrad_kernel<<<(Hv + 255) / 256, 256, 0, st>>>(E, Hv, 1, dv, h[3], (size_t)Hv, -1, -1, dW[4], db[4]);
  v_dgrad_kernel<<<gD, 256, 0, st>>>(E, Hv, 1, dv, W[4], h[3], dz[0]);
  int cur = 0;
  for (int l = 3; l >= 1; --l) {
    v_wgrad_kernel<<<(Hv * Hv + 255) / 256, 256, 0, st>>>(E, Hv, Hv, dz[cur], h[l - 1], (size_t)Hv, -1, -1, dW[l], db[l]);
    v_dgrad_kernel<<<gD, 256, 0, st>>>(E, Hv, Hv, dz[cur], W[l], h[l - 1], dz[cur ^ 1]);
    cur ^= 1;
  }
  v_wgrad_kernel<<<(Hv * 2 + 255) / 256, 256, 0, st>>>(E, 2, Hv, dz[cur], states, state_stride, t_col, T_col, dW[0], db[0]);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

// ======================================================================================================================
// Minibatch gather + advantage normalisation (ppo.py:129-144) and the logged loss sums (ppo.py:158-176) as two / one
// launches instead of ~30 torch kernels per optimiser step.
// ======================================================================================================================
namespace mbo {
namespace upd {

// block 0: scalars of the E selected transitions + (adv - mean) / std with the biased std in double (ppo.py:141-144: left
// alone if std is 0 or NaN); blocks 1..: state rows
__global__ void __launch_bounds__(256) prep_kernel(int E, size_t row_floats, const long long* __restrict__ sel,
                                                    const float* __restrict__ states_all, const int32_t* __restrict__ actions_all,
                                                    const float* __restrict__ advs_all, const float* __restrict__ ret_all,
                                                    const float* __restrict__ lpo_all, int normalize,
                                                    const double* __restrict__ adv_stats,
                                                    float* __restrict__ states, int32_t* __restrict__ actions,
                                                    float* __restrict__ advs, float* __restrict__ ret, float* __restrict__ lpo) {
  if (blockIdx.x == 0) {
    __shared__ double s_red[8];
    double s = 0.0, s2 = 0.0, cnt = (double)E;
    if (adv_stats) {        // data-parallel update: {sum, sum of squares, count} over the GLOBAL minibatch (all ranks)
      s = adv_stats[0]; s2 = adv_stats[1]; cnt = adv_stats[2];
    } else {
      for (int e = threadIdx.x; e < E; e += 256) { const double a = (double)advs_all[sel[e]]; s += a; s2 += a * a; }
      s = block_sum(s, s_red);
      s2 = block_sum(s2, s_red);
    }
    const double mean = s / cnt;
    const double var = fmax(s2 / cnt - mean * mean, 0.0);
    const double sd = sqrt(var);
    const bool ok = normalize && sd > 0.0 && !isnan(sd);
    for (int e = threadIdx.x; e < E; e += 256) {
      const long long i = sel[e];
      const float a = advs_all[i];
      advs[e] = ok ? (float)(((double)a - mean) / sd) : a;
      actions[e] = actions_all[i];
      ret[e] = ret_all[i];
      lpo[e] = lpo_all ? lpo_all[i] : 0.f;
    }
    return;
  }
  const int per = (E + (int)gridDim.x - 2) / ((int)gridDim.x - 1);
  const int e0 = ((int)blockIdx.x - 1) * per, e1 = min(E, e0 + per);
  for (int e = e0; e < e1; ++e) {
    const float* src = states_all + (size_t)sel[e] * row_floats;
    float* dst = states + (size_t)e * row_floats;
    if ((row_floats & 3) == 0) {
      for (size_t i = threadIdx.x; i < row_floats / 4; i += 256) ((float4*)dst)[i] = ((const float4*)src)[i];
    } else {
      for (size_t i = threadIdx.x; i < row_floats; i += 256) dst[i] = src[i];
    }
  }
}

// local {sum, sum of squares, count} of the advantages of every minibatch of an epoch (one CTA per minibatch): the
// ranks all-reduce the [n_steps, 3] table once per epoch, prep_kernel then normalises with the global moments
__global__ void __launch_bounds__(256) adv_stats_kernel(const int32_t* __restrict__ bounds, const long long* __restrict__ sel,
                                                         const float* __restrict__ advs_all, double* __restrict__ out) {
  __shared__ double s_red[8];
  const int a = bounds[blockIdx.x], b = bounds[blockIdx.x + 1];
  double s = 0.0, s2 = 0.0;
  for (int e = a + threadIdx.x; e < b; e += 256) { const double x = (double)advs_all[sel[e]]; s += x; s2 += x * x; }
  s = block_sum(s, s_red);
  s2 = block_sum(s2, s_red);
  if (threadIdx.x == 0) { out[blockIdx.x * 3] = s; out[blockIdx.x * 3 + 1] = s2; out[blockIdx.x * 3 + 2] = (double)(b - a); }
}

// out = {loss_ppo, loss_value, loss_ent, loss} of ppo.py:158-168 from the per-transition terms
__global__ void __launch_bounds__(256) loss_sums_kernel(int E, const float* __restrict__ terms, const float* __restrict__ vterms,
                                                         float inv_n, float value_coeff, float ent_coeff, float* __restrict__ out) {
  __shared__ double s_red[8];
  double p = 0.0, h = 0.0, v = 0.0;
  for (int e = threadIdx.x; e < E; e += 256) {
    p += (double)terms[e * 4 + 2];
    h += (double)terms[e * 4 + 1];
    v += vterms ? (double)vterms[e] : 0.0;
  }
  p = block_sum(p, s_red);
  h = block_sum(h, s_red);
  v = block_sum(v, s_red);
  if (threadIdx.x == 0) {
    const float l_ppo = (float)p, l_val = (float)(v * inv_n), l_ent = (float)(-h * inv_n);
    out[0] = l_ppo; out[1] = l_val; out[2] = l_ent;
    out[3] = l_ppo + value_coeff * l_val + ent_coeff * l_ent;
  }
}

}  // namespace upd
}  // namespace mbo

extern "C" int mbo_ppo_prepare_minibatch(int E, size_t row_floats, const int64_t* sel, const float* states_all,
                                         const int32_t* actions_all, const float* advs_all, const float* tdlamrets_all,
                                         const float* logp_old_all, int normalize_advs, const double* adv_stats,
                                         float* states, int32_t* actions, float* advs, float* tdlamrets, float* logp_old,
                                         void* stream) {
  MBO_CHECK_ARG(E > 0 && row_floats > 0, "mbo_ppo_prepare_minibatch: E, row_floats must be > 0");
  MBO_CHECK_ARG(sel && states_all && actions_all && advs_all && tdlamrets_all && states && actions && advs && tdlamrets &&
                logp_old, "mbo_ppo_prepare_minibatch: null pointer");
  const int nb = 1 + (E < 296 ? E : 296);
  mbo::upd::prep_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(E, row_floats, (const long long*)sel, states_all, actions_all,
                                                              advs_all, tdlamrets_all, logp_old_all, normalize_advs, adv_stats,
                                                              states, actions, advs, tdlamrets, logp_old);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_ppo_adv_stats(int n_steps, const int32_t* bounds, const int64_t* sel, const float* advs_all, double* out,
                                 void* stream) {
  MBO_CHECK_ARG(n_steps > 0 && bounds && sel && advs_all && out, "mbo_ppo_adv_stats: bad arguments");
  mbo::upd::adv_stats_kernel<<<n_steps, 256, 0, (cudaStream_t)stream>>>(bounds, (const long long*)sel, advs_all, out);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_ppo_loss_sums(int E, const float* terms, const float* vterms, float inv_n_global, float value_coeff,
                                 float ent_coeff, float* out4, void* stream) {
  MBO_CHECK_ARG(E > 0 && terms && out4, "mbo_ppo_loss_sums: bad arguments");
  mbo::upd::loss_sums_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(E, terms, vterms, inv_n_global, value_coeff, ent_coeff, out4);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

#ifdef MBO_UPD_TRACE
extern "C" int mbo_upd_trace_read(long long* host_out /*[4096]*/) {
  return cudaMemcpyFromSymbol(host_out, mbo::upd::g_upd_trace, sizeof(long long) * 4096) == cudaSuccess ? 0 : -2;
}
#endif
