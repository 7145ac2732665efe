// coresidency.cu -- why do the fp64 GP kernel and the persistent tcgen05 MLP kernel not run concurrently on two streams?
// A "big" persistent kernel (148 CTAs x 448 threads, 219 KB dynamic shared memory, ~118 registers, optional TMEM
// allocation) spins for a fixed number of clocks; a "small" kernel (many CTAs x 128 threads, 12 KB shared memory,
// ~80 registers) spins too.  Launched on two streams: if the small CTAs co-reside, total time ~ max, else ~ sum.
// Variants isolate the candidate blockers: TMEM ownership, shared-memory carve-out preference, big-kernel smem size.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o coresidency coresidency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// REGS_PAD values loaded through volatile pointers stay live in registers across the spin loop
#define HOLD_BEGIN(N) float r_[N]; _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_) r_[i_] = ((volatile float*)sm)[(threadIdx.x + i_) & 255];
#define HOLD_END(N) _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_) acc += r_[i_];

template <int HOLD>
__global__ void __launch_bounds__(448, 1) big(long long clocks, int use_tmem, float* sink, long long* t_first_last) {
  extern __shared__ unsigned char sm[];
  __shared__ uint32_t s_tmem;
  if ((use_tmem & 1) && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  sm[threadIdx.x] = (unsigned char)threadIdx.x;
  if (use_tmem & 2) {       // claim all 16 named barriers like the MLP kernel does
    asm volatile("bar.sync 1, 64;" ::: "memory"); asm volatile("bar.sync 2, 64;" ::: "memory"); asm volatile("bar.sync 3, 64;" ::: "memory");
    asm volatile("bar.sync 4, 64;" ::: "memory"); asm volatile("bar.sync 5, 64;" ::: "memory"); asm volatile("bar.sync 6, 64;" ::: "memory");
    asm volatile("bar.sync 7, 64;" ::: "memory"); asm volatile("bar.sync 8, 64;" ::: "memory"); asm volatile("bar.sync 9, 64;" ::: "memory");
    asm volatile("bar.sync 10, 64;" ::: "memory"); asm volatile("bar.sync 11, 64;" ::: "memory"); asm volatile("bar.sync 12, 64;" ::: "memory");
    asm volatile("bar.sync 13, 64;" ::: "memory"); asm volatile("bar.sync 14, 64;" ::: "memory"); asm volatile("bar.sync 15, 64;" ::: "memory");
  }
  float acc = 0.f;
  HOLD_BEGIN(HOLD)
  const long long t0 = clock64();
  while (clock64() - t0 < clocks) acc += sm[(threadIdx.x * 7) & 1023];
  HOLD_END(HOLD)
  __syncthreads();
  if ((use_tmem & 1) && threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
  if (acc == 1.2345f) sink[0] = acc;
}

__global__ void __launch_bounds__(128) small(long long clocks, float* sink, unsigned int* smid_hist) {
  extern __shared__ unsigned char sm[];
  sm[threadIdx.x] = (unsigned char)threadIdx.x;
  float acc = 0.f;
  HOLD_BEGIN(64)
  const long long t0 = clock64();
  while (clock64() - t0 < clocks) acc += sm[(threadIdx.x * 3) & 1023];
  HOLD_END(64)
  if (acc == 1.2345f) sink[0] = acc;
}

template <int HOLD>
static float run(int use_tmem, int big_smem, int small_carve, int order, int small_ctas) {
  auto big = ::big<HOLD>;
  cudaFuncAttributes fa, fb; cudaFuncGetAttributes(&fa, big); cudaFuncGetAttributes(&fb, small);
  printf("[big %d regs, small %d regs] ", fa.numRegs, fb.numRegs);
  float* sink; cudaMalloc(&sink, 64);
  cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288);
  if (small_carve >= 0) cudaFuncSetAttribute(small, cudaFuncAttributePreferredSharedMemoryCarveout, small_carve);
  const long long big_clk = 2000000, small_clk = 200000;     // ~1 ms and ~0.1 ms per CTA: 40 small CTAs per SM = 0.7 ms alone at 6 resident
  auto once = [&](int which) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    cudaStreamWaitEvent(s1, e0, 0); cudaStreamWaitEvent(s2, e0, 0);
    if (which & 1) { if (order == 0) big<<<148, 448, big_smem, s1>>>(big_clk, use_tmem, sink, nullptr); }
    if (which & 2) small<<<small_ctas, 128, 12288, s2>>>(small_clk, sink, nullptr);
    if (which & 1) { if (order == 1) big<<<148, 448, big_smem, s1>>>(big_clk, use_tmem, sink, nullptr); }
    cudaEvent_t j1, j2; cudaEventCreate(&j1); cudaEventCreate(&j2);
    cudaEventRecord(j1, s1); cudaEventRecord(j2, s2);
    cudaStreamWaitEvent(0, j1, 0); cudaStreamWaitEvent(0, j2, 0);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
  };
  once(3);
  const float tb = once(1), ts = once(2), both = once(3);
  printf("tmem=%d big_smem=%6d small_carveout=%4d order=%s small_ctas=%5d : big %.3f ms, small %.3f ms, both %.3f ms  -> %s (%s)\n", use_tmem,
         big_smem, small_carve, order ? "small-first" : "big-first", small_ctas, tb, ts, both,
         both < tb + 0.5f * ts ? "OVERLAP" : "serialised", cudaGetErrorString(cudaGetLastError()));
  cudaStreamDestroy(s1); cudaStreamDestroy(s2); cudaFree(sink);
  return both;
}

int main() {
  const int full = 217920;                 // the fp16 MLP kernel's dynamic shared memory
  run<100>(1, full, 100, 0, 148 * 40);     // ~110 registers: 4 warps x 112 x 32 = 14 336 of a sub-partition's 16 384
  run<92>(1, full, 100, 0, 148 * 40);
  run<84>(1, full, 100, 0, 148 * 40);
  run<84>(0, full, 100, 0, 148 * 40);
  run<84>(1, full, -1, 0, 148 * 40);
  run<84>(1, 100000, -1, 0, 148 * 40);
  run<76>(1, full, 100, 0, 148 * 40);
  run<60>(1, full, 100, 0, 148 * 40);
  run<84>(3, full, 100, 0, 148 * 40);      // + 16 named barriers
  run<84>(2, full, 100, 0, 148 * 40);
  return 0;
}
