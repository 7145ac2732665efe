// neural_af_tc.cu -- fused NeuralAF evaluation on tcgen05 tensor cores (sm_100a).
//
// Replaces get_state feature packing (metabo/environment/metabo_gym.py:624-677) followed by
// NeuralAF.forward -> MLP.forward (metabo/policies/policies.py:104-125, mlp.py:52-61) for the
// shipped F -> H -> H -> H -> H -> 1 ReLU policies (H = 200): features and hidden activations
// never touch HBM; one CTA owns a tile of 128 candidates of one episode.
//
// Data flow per tile (TMEM columns: R0 = [0,208), R1 = [208,416), S = [416,448)):
//   features (fp32)  --split hi/lo tf32-->  S                         (epilogue warps, tcgen05.st)
//   L0: D(R0)  = S * W0^T     3xTF32 (hi*hi + lo*hi + hi*lo), bias folded in as a ones column
//   E0: R0 <- act(relu(D))    in place, converted to the MMA operand format
//   L1: D(R1)  = A(R0) * W1^T ; E1 in place on R1 ; L2: D(R0) = A(R1) * W2^T ; E2 ; L3: D(R1) = A(R0) * W3^T
//   E3: logit = sum_j relu(D_j + b3_j) * w4_j + b4                    (CUDA cores, fp32)
// The A operand of every MMA comes from TMEM (tcgen05.mma "TS" form), the B operand (weights,
// K-major, no-swizzle canonical layout) streams from L2 through a shared-memory ring filled with
// cp.async.bulk (TMA engine, 1-D) and released by tcgen05.commit.  The epilogue of layer l hands
// the next layer its K range chunk by chunk (mbarriers), so the MMAs of layer l+1 start while the
// epilogue of layer l is still converting.
//
// Arithmetic modes (hidden layers): MBO_MLP_TF32   kind::tf32, operands rounded to tf32 (rna);
//                                   MBO_MLP_BF16X3 kind::f16 bf16, a = a_hi + a_lo, w = w_hi + w_lo,
//                                                  a*w ~ a_hi*w_hi + a_lo*w_hi + a_hi*w_lo.
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lanes 32*(w&3).., chunk parity w>>2), warp 8 TMA
// producer + TMEM allocation, warp 9 MMA issuer.  Every mbarrier wait is bounded (globaltimer) and raises a
// device-side error flag instead of hanging.
#include "common.cuh"
#include "policy.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>
#include "tmem_ldst.cuh"

namespace mbo {

constexpr int TC_N = 208;            // padded hidden width (UMMA N, multiple of 16)
constexpr int TC_M = 128;            // candidates per tile (UMMA M)
constexpr int TC_THREADS = 320;        // 8 epilogue warps + TMA warp + MMA warp
constexpr int WARP_TMA = 8, WARP_MMA = 9, WARP_GP0 = 10;   // warps 10-13: fused GP posterior (FUSE only)
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

constexpr int W0_BYTES = 2 * (TC_K0 / 4) * CHUNK_ROW_BYTES;   // hi + lo: 13 312
constexpr int TAIL_FLOATS = 5 * TC_N + 4;                    // zeros(b0 is folded into W0) b1 b2 b3 | w4 | b4

// packed-weight buffer layout (bytes): [W0][tf32 hidden x3][bf16 hidden x3][tail floats]
struct PackLayout {
  size_t off_w0, off_tf32, off_bf16, off_tail, total;
};
__host__ __device__ inline PackLayout pack_layout() {
  PackLayout p;
  p.off_w0 = 0;
  p.off_tf32 = W0_BYTES;
  p.off_bf16 = p.off_tf32 + (size_t)3 * ModeCfg<MBO_MLP_TF32>::STAGES_PER_LAYER * ModeCfg<MBO_MLP_TF32>::STAGE_BYTES;
  p.off_tail = p.off_bf16 + (size_t)3 * ModeCfg<MBO_MLP_BF16X3>::STAGES_PER_LAYER * ModeCfg<MBO_MLP_BF16X3>::STAGE_BYTES;
  p.total = p.off_tail + sizeof(float) * TAIL_FLOATS;
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
// bounded wait; returns false (and latches *abort) on timeout or when another role already aborted
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (*abort_flag) return false;
      if (globaltimer_ns() - t0 > WAIT_TIMEOUT_NS) { *abort_flag = 1; return false; }
    }
  }
  return true;
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
constexpr uint32_t IDESC_TF32 = make_idesc(2), IDESC_BF16 = make_idesc(1);

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

__global__ void tc_pack_kernel(PackArgs a, unsigned char* __restrict__ out) {
  const PackLayout pl = pack_layout();
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit word index
  const size_t byte = w * 4;
  if (byte >= pl.total) return;
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
  } else {
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
    p->tc_bytes = pl.total;
  }
  PackArgs a;
  for (int l = 0; l < 5; ++l) { a.W[l] = p->d_W_raw + p->raw_off[l]; a.b[l] = p->d_b[l]; }
  a.F = F; a.H = H;
  size_t words = pl.total / 4;
  tc_pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(a, (unsigned char*)p->d_tc);
  MBO_LAUNCH_CHECK();
  p->tc_ready = 1;
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
};

template <int MODE, bool FUSE = false>
struct TcSmem {
  using C = ModeCfg<MODE>;
  static constexpr int NS = FUSE ? C::NSTAGE_FUSED : C::NSTAGE;
  static constexpr size_t off_ring = 0;
  static constexpr size_t off_w0 = off_ring + (size_t)NS * C::STAGE_BYTES;
  static constexpr size_t off_tail = off_w0 + W0_BYTES;
  static constexpr size_t off_bars = off_tail + sizeof(float) * TAIL_FLOATS;
  static constexpr int n_bars = 2 * NS + C::NCHUNK + 2 + 3 + 4;   // ring, a_ready, d_full[2], feat_full, s_free, w0_full, gp_full[2], gp_empty[2]
  static constexpr size_t off_misc = off_bars + 8 * n_bars;
  static constexpr size_t off_part = off_misc + 16;
  static constexpr size_t off_gp = off_part + sizeof(float) * TC_M;             // fused GP: episode blob + mu/sd hand-off
  static constexpr size_t off_musd = off_gp + (FUSE ? sizeof(double) * GPB_DOUBLES : 0);
  static constexpr size_t off_ks = off_musd + (FUSE ? sizeof(float) * 4 * TC_M : 0);        // k* scratch [32][128] fp64
  static constexpr size_t total = off_ks + (FUSE ? sizeof(double) * GP_NP * TC_M : 0);
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

template <int MODE, bool FUSE, int K0S>     // K0S: K = 8 slices of layer 0 (1: F + 1 <= 8, 2: F + 1 <= 16)
__global__ void __launch_bounds__(FUSE ? TC_THREADS_FUSED : TC_THREADS, 1)
neural_af_tc_kernel(const unsigned char* __restrict__ pack, TcFeat fs, mbo_candidates cand, int B,
                    int tiles_per_episode, const float* __restrict__ mean, const float* __restrict__ std_,
                    const float* __restrict__ epi, float* __restrict__ logits, int* __restrict__ err_flag, int debug_noload,
                    GpFuse gp, TcTopk tk) {
  using C = ModeCfg<MODE>;
  using SM = TcSmem<MODE, FUSE>;
  constexpr int NTHREADS = FUSE ? TC_THREADS_FUSED : TC_THREADS;
  constexpr int NS = SM::NS;                      // ring depth
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PackLayout pl = pack_layout();

  unsigned char* ring = smem + SM::off_ring;
  float* s_tail = (float*)(smem + SM::off_tail);
  uint64_t* bars = (uint64_t*)(smem + SM::off_bars);
  uint32_t* s_tmem = (uint32_t*)(smem + SM::off_misc);
  volatile int* s_abort = (volatile int*)(smem + SM::off_misc + 4);
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
  float* s_mu = (float*)(smem + SM::off_musd);            // [2][128]
  float* s_sd = s_mu + 2 * TC_M;                          // [2][128]

  const int total_tiles = B * tiles_per_episode;
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup ----
  if (tid == 0) {
    *s_abort = 0;
    for (int s = 0; s < NS; ++s) { mbar_init(BAR_WFULL(s), 1); mbar_init(BAR_WEMPTY(s), 1); }
    for (int c = 0; c < C::NCHUNK; ++c) mbar_init(BAR_AREADY(c), 4);     // one elected arrive per epilogue warp
    mbar_init(BAR_DFULL(0), 1);
    mbar_init(BAR_DFULL(1), 1);
    mbar_init(BAR_FEAT, 4);
    mbar_init(BAR_SFREE, 1);
    mbar_init(BAR_W0, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR_GPFULL(i), 128); mbar_init(BAR_GPEMPTY(i), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < TAIL_FLOATS; i += NTHREADS) s_tail[i] = ((const float*)(pack + pl.off_tail))[i];
  if (warp == WARP_TMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == WARP_TMA) {
    // ================= TMA producer (warp-uniform control flow, one elected lane issues) =================
    const bool leader = elect_one();
    if (my_tiles > 0) {
      if (leader) {
        mbar_expect_tx(BAR_W0, W0_BYTES);
        bulk_g2s(smem_u32(smem + SM::off_w0), pack + pl.off_w0, W0_BYTES, BAR_W0);
      }
      const unsigned char* wsrc = pack + (MODE == MBO_MLP_TF32 ? pl.off_tf32 : pl.off_bf16);
      const uint32_t ring_addr = smem_u32(ring);
      uint32_t stage = 0, phase = 0;
      bool ok = true;
      for (int t = 0; t < my_tiles && ok; ++t) {
        for (int ls = 0; ls < 3 * C::STAGES_PER_LAYER; ++ls) {
          if (!mbar_wait(BAR_WEMPTY(stage), phase ^ 1u, s_abort)) { ok = false; break; }
          if (leader) {
            if (debug_noload && (t > 0 || ls >= NS)) {
              mbar_arrive(BAR_WFULL(stage));    // timing experiment only: pretend the stage is resident
            } else {
              mbar_expect_tx(BAR_WFULL(stage), C::STAGE_BYTES);
              bulk_g2s(ring_addr + stage * C::STAGE_BYTES, wsrc + (size_t)ls * C::STAGE_BYTES, C::STAGE_BYTES,
                       BAR_WFULL(stage));
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
      bool ok = mbar_wait(BAR_W0, 0, s_abort);
      uint32_t stage = 0, wphase = 0, aphase = 0, fphase = 0;
      const uint32_t w0_lo = desc_lo(smem_u32(smem + SM::off_w0));
      const uint32_t ring_lo = desc_lo(smem_u32(ring));
      constexpr uint32_t STAGE_D = C::STAGE_BYTES >> 4;            // descriptor units (16 B)
      constexpr uint32_t SLICE_D = (2 * CHUNK_ROW_BYTES) >> 4;     // one K slice = two 16-byte chunks
      for (int t = 0; t < my_tiles && ok; ++t) {
        // ---- L0: D(R0) = S * W0^T, 3xTF32 ----
        if (!mbar_wait(BAR_FEAT, fphase, s_abort)) { ok = false; break; }
        fphase ^= 1u;
        tc_fence_after();
        if (leader) {
          const uint32_t a_hi = tmem + COL_S, a_lo = tmem + COL_S + 16;
          constexpr uint32_t W0_PART_D = (uint32_t)((TC_K0 / 4) * CHUNK_ROW_BYTES) >> 4;     // hi -> lo part
#pragma unroll
          for (int sl = 0; sl < K0S; ++sl) {
            const uint32_t wl = w0_lo + sl * SLICE_D;
            mma_ts_tf32(tmem + COL_R0, a_hi + sl * 8, make_desc64(wl), IDESC_TF32, sl ? 1u : 0u);
            mma_ts_tf32(tmem + COL_R0, a_lo + sl * 8, make_desc64(wl), IDESC_TF32, 1);
            mma_ts_tf32(tmem + COL_R0, a_hi + sl * 8, make_desc64(wl + W0_PART_D), IDESC_TF32, 1);
          }
          tc_commit(BAR_SFREE);
          tc_commit(BAR_DFULL(0));
        }
        // ---- L1..L3 ----
#pragma unroll
        for (int l = 1; l <= 3; ++l) {
          if (!ok) break;
          const uint32_t tA = tmem + ((l & 1) ? COL_R0 : COL_R1);
          const uint32_t tD = tmem + ((l & 1) ? COL_R1 : COL_R0);
#pragma unroll
          for (int c = 0; c < C::NCHUNK; ++c) {
            if (!ok) break;
            if (!mbar_wait(BAR_AREADY(c), aphase, s_abort)) { ok = false; break; }
            if (MODE == MBO_MLP_TF32) {
              // chunk c == ring stage: 5 slices of K=8
              if (!mbar_wait(BAR_WFULL(stage), wphase, s_abort)) { ok = false; break; }
              tc_fence_after();
              if (leader) {
                const uint32_t lo = ring_lo + stage * STAGE_D;
#pragma unroll
                for (int sc = 0; sc < 5; ++sc)
                  mma_ts_tf32(tD, tA + (c * 5 + sc) * 8, make_desc64(lo + sc * SLICE_D), IDESC_TF32, (c | sc) ? 1u : 0u);
                tc_commit(BAR_WEMPTY(stage));
              }
              if (++stage == NS) { stage = 0; wphase ^= 1u; }
            } else {
              // chunk c holds [hi(16 cols) | lo(16 cols)] of 32 K values = 2 slices (1 in the last chunk); slice == stage
              constexpr int LASTC = C::NCHUNK - 1;
#pragma unroll
              for (int sc = 0; sc < 2; ++sc) {
                if (c == LASTC && sc == 1) break;
                if (!mbar_wait(BAR_WFULL(stage), wphase, s_abort)) { ok = false; break; }
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
        }
      }
    }
  } else if (warp < WARP_TMA) {
    // ================= epilogue warps =================
    // warp w: TMEM lanes 32*(w&3)..+31 (thread = candidate row), column half h = w>>2:
    // half h converts chunks h, h+2, ...; half 1 also builds the next tile's features
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;                       // 0..127
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t dphase[2] = {0, 0};
    uint32_t sphase = 0;
    bool ok = true;
    const float* s_w4 = s_tail + 4 * TC_N;          // s_tail = [zeros | b1 | b2 | b3 | w4 | b4]

    // Feature vector of this thread's row for tile `t_idx` (fp32, ones column appended).  It is gathered one tile
    // AHEAD into registers (`pf`) so that the global-memory latency of mean/std/candidates never sits between the
    // last hidden-layer epilogue and the head epilogue (half-0 warps wait for half-1 at the pair barrier there).
    constexpr int KF = 8 * K0S;                        // feature slots (incl. the ones column)
    auto gather_features = [&](int t_idx, float* f) {
      const int tile = (int)blockIdx.x + t_idx * (int)gridDim.x;
      const int b = tile / tiles_per_episode;
      const int m = (tile - b * tiles_per_episode) * TC_M + row;
#pragma unroll
      for (int i = 0; i < KF; ++i) f[i] = 0.f;
      if (t_idx >= my_tiles || m >= cand.M) return;
      if (fs.dense) {
        const float* srow = fs.dense + ((size_t)b * cand.M + m) * fs.dense_F;
#pragma unroll
        for (int i = 0; i < KF - 1; ++i)
          if (i < fs.F) f[i] = srow[fs.src[i]];
      } else {
        double x[MBO_MAX_D];
#pragma unroll
        for (int d = 0; d < MBO_MAX_D; ++d) x[d] = 0.0;
        load_candidate(cand, b, m, x);
        float mu = 0.f, sd = 0.f;
        if (!FUSE) { mu = mean[(size_t)b * cand.M + m]; sd = std_[(size_t)b * cand.M + m]; }
#pragma unroll
        for (int i = 0; i < KF - 1; ++i) {
          if (i < fs.F) {
            const int s = fs.src[i];
            float v;
            if (s == MBO_FEAT_MEAN) v = mu;
            else if (s == MBO_FEAT_STD) v = sd;
            else if (s >= MBO_FEAT_X0 && s < MBO_FEAT_X0 + MBO_MAX_D) {
              v = 0.f;
#pragma unroll
              for (int d = 0; d < MBO_MAX_D; ++d) if (s - MBO_FEAT_X0 == d) v = (float)x[d];
            } else v = epi[b * 4 + (s - MBO_FEAT_INCUMBENT)];
            f[i] = v;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < KF; ++i) if (i == fs.F) f[i] = 1.0f;    // bias column
    };
    float pf[KF];                                   // prefetched features of the next tile to be built

    auto build_features = [&](int t_idx) {
      if (FUSE) {   // posterior of this tile from the GP warps
        if (!mbar_wait(BAR_GPFULL(t_idx & 1), (uint32_t)(t_idx >> 1) & 1u, s_abort)) ok = false;
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
      gather_features(t_idx + 1, pf);                  // loads for the next tile are in flight from here on
    };

    if (my_tiles > 0 && half == 1) { gather_features(0, pf); build_features(0); }
    constexpr int MAXI = (C::NCHUNK + 1) / 2;
    for (int t = 0; t < my_tiles && ok; ++t) {
      // ---- E0..E2: relu(D + b) -> next layer's A operand, in place ----
      for (int l = 0; l < 3 && ok; ++l) {
        const uint32_t col = (l & 1) ? COL_R1 : COL_R0;
        if (!mbar_wait(BAR_DFULL(l & 1), dphase[l & 1], s_abort)) { ok = false; break; }
        dphase[l & 1] ^= 1u;
        tc_fence_after();
        const float* bias = s_tail + l * TC_N;
        uint32_t va[40], vb[40];
        auto ncols_of = [&](int c) { return (MODE == MBO_MLP_TF32) ? 40 : (c == C::NCHUNK - 1 ? 16 : 32); };
        ChunkOps<MODE>::load(tlane + col + half * C::CHUNK_COLS, va, ncols_of(half));
#pragma unroll
        for (int i = 0; i < MAXI; ++i) {
          const int c = half + 2 * i;
          if (c < C::NCHUNK) {
            uint32_t* cur = (i & 1) ? vb : va;
            uint32_t* nxt = (i & 1) ? va : vb;
            tmem_wait_ld();
            if (c + 2 < C::NCHUNK) ChunkOps<MODE>::load(tlane + col + (c + 2) * C::CHUNK_COLS, nxt, ncols_of(c + 2));
            ChunkOps<MODE>::convert_store(tlane + col + c * C::CHUNK_COLS, cur, bias + c * C::CHUNK_COLS, ncols_of(c));
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_AREADY(c));
          }
        }
        if (l == 2 && half == 1) {
          if (t + 1 < my_tiles) {
            // features of the next tile while L3 runs
            if (!mbar_wait(BAR_SFREE, sphase, s_abort)) { ok = false; break; }
            tc_fence_after();
            build_features(t + 1);
          }
          sphase ^= 1u;
        }
      }
      if (!ok) break;
      // ---- E3: logit = relu(D + b3) . w4 + b4, columns split between the two halves ----
      if (!mbar_wait(BAR_DFULL(1), dphase[1], s_abort)) { ok = false; break; }
      dphase[1] ^= 1u;
      tc_fence_after();
      float acc = 0.f;
      {
        const int cb = half * 104;
        uint32_t v0[64], v1[40];
        tmem_ld64(tlane + COL_R1 + cb, v0);
        tmem_wait_ld();
        tmem_ld32(tlane + COL_R1 + cb + 64, v1);
        tmem_ld8(tlane + COL_R1 + cb + 96, v1 + 32);
        const float4* b3 = reinterpret_cast<const float4*>(s_tail + 3 * TC_N + cb);
        const float4* w4 = reinterpret_cast<const float4*>(s_w4 + cb);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 b = b3[j], w = w4[j];
          acc = fmaf(fmaxf(__uint_as_float(v0[4 * j + 0]) + b.x, 0.f), w.x, acc);
          acc = fmaf(fmaxf(__uint_as_float(v0[4 * j + 1]) + b.y, 0.f), w.y, acc);
          acc = fmaf(fmaxf(__uint_as_float(v0[4 * j + 2]) + b.z, 0.f), w.z, acc);
          acc = fmaf(fmaxf(__uint_as_float(v0[4 * j + 3]) + b.w, 0.f), w.w, acc);
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const float4 b = b3[16 + j], w = w4[16 + j];
          acc = fmaf(fmaxf(__uint_as_float(v1[4 * j + 0]) + b.x, 0.f), w.x, acc);
          acc = fmaf(fmaxf(__uint_as_float(v1[4 * j + 1]) + b.y, 0.f), w.y, acc);
          acc = fmaf(fmaxf(__uint_as_float(v1[4 * j + 2]) + b.z, 0.f), w.z, acc);
          acc = fmaf(fmaxf(__uint_as_float(v1[4 * j + 3]) + b.w, 0.f), w.w, acc);
        }
      }
      tc_fence_before();   // order this tile's TMEM reads before the next hand-offs
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
      if (tk.k && half == 0) {
        // warp top-k over this warp's 32 rows (value descending, index ascending; NaN / padding never win), off the
        // pair barrier's critical path; two REDUX per round on an order-preserving integer key
        const bool valid = tile_m < cand.M && lgt == lgt;
        const uint32_t u = __float_as_uint(lgt + 0.0f);                                 // -0 -> +0 (compare equal as floats)
        uint32_t key = valid ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0u;     // 0 < every valid key
        int mi = valid ? tile_m : 0x7fffffff;
        uint32_t outk = 0u;
        int outm = 0x7fffffff;
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
    bool ok = true;
    for (int t = 0; t < my_tiles && ok; ++t) {
      const int buf = t & 1;
      if (!mbar_wait(BAR_GPEMPTY(buf), ((uint32_t)(t >> 1) & 1u) ^ 1u, s_abort)) { ok = false; break; }
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
  if (tid == 0 && *s_abort) atomicExch(err_flag, 1);
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

template <int MODE, bool FUSE, int K0S>
static int launch_tc_k(const mbo_policy* p, int B, const TcFeat& fs, const mbo_candidates& cand, const float* mean,
                     const float* std_, const float* epi, float* logits, int* err_flag, cudaStream_t s,
                     const GpFuse& gp, const TcTopk& tk) {
  using SM = TcSmem<MODE, FUSE>;
  auto kern = neural_af_tc_kernel<MODE, FUSE, K0S>;
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
      (const unsigned char*)p->d_tc, fs, cand, B, tiles_per_episode, mean, std_, epi, logits, err_flag,
      getenv("MBO_TC_DEBUG_NOLOAD") ? 1 : 0, gp, tk);
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
  if (mode == MBO_MLP_TF32)
    rc = launch_tc<MBO_MLP_TF32, false>(p, B, fs, *cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  else
    rc = launch_tc<MBO_MLP_BF16X3, false>(p, B, fs, *cand, mean, std_, epi, logits, err_flag, s, gp, tk);
  if (rc || !topk) return rc;
  topk_merge_kernel<<<B, 128, 0, s>>>(tiles * 4, topk, tk.part_val, tk.part_idx, topk_idx, topk_val);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
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
  MBO_CHECK_ARG(mode == MBO_MLP_TF32 || mode == MBO_MLP_BF16X3, "mbo_af_eval_fused: tensor-core modes only");
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
  GpFuse gp;
  gp.gp_state = gp_state;
  gp.nmax = nmax;
  gp.kid = kernel_id;
  gp.mean_out = mean_out;
  gp.std_out = std_out;
  cudaStream_t s = (cudaStream_t)stream;
  TcTopk tk;
  memset(&tk, 0, sizeof(tk));
  if (mode == MBO_MLP_TF32)
    return launch_tc<MBO_MLP_TF32, true>(p, B, fs, *cand, nullptr, nullptr, epi, logits, (int*)workspace, s, gp, tk);
  return launch_tc<MBO_MLP_BF16X3, true>(p, B, fs, *cand, nullptr, nullptr, epi, logits, (int*)workspace, s, gp, tk);
}
