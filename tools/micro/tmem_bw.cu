// tmem_bw.cu -- micro-benchmark: tcgen05.ld / tcgen05.st throughput and latency per SM on sm_100a.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../metabo_b200/csrc -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tmem_ldst.cuh"
using namespace mbo;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: back-to-back ld x32 (no wait between), one wait at the end of each group of `grp`
// mode 1: ld x32 + wait per iteration (latency chain)
// mode 2: st x32 back-to-back, wait at end of group
// mode 3: st x16 + wait per iteration
// mode 4: ld x32, wait, st x16, wait (epilogue iteration without ALU)
__global__ void k(int mode, int iters, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  uint32_t v[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  tmem_st32(tl, v);
  tmem_st32(tl + 32, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        tmem_ld32(tl, v);
        if ((it & 3) == 3) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += v[it & 31]; }
      } else if (mode == 1) {
        tmem_ld32(tl, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += v[it & 31];
      } else if (mode == 2) {
        tmem_st32(tl, v);
        if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else if (mode == 3) {
        tmem_st16(tl, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        tmem_ld32(tl, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        v[0] += acc;
        tmem_st16(tl + 32, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (acc == 0xdeadbeef) sink[threadIdx.x] = acc + v[3];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMallocManaged(&out, 8 * 32 * 148);
  cudaMalloc(&sink, 4096);
  const char* names[] = {"ld x32 pipelined", "ld x32 + wait", "st x32 pipelined", "st x16 + wait", "ld32 wait st16 wait"};
  const int bytes[] = {4096, 4096, 4096, 2048, 4096};
  for (int mode = 0; mode < 5; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      const int iters = 2000;
      k<<<1, 512>>>(mode, iters, nw, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = out[w] > mx ? out[w] : mx;
      printf("%-22s warps %2d: %7.1f cyc/iter/warp, %7.1f B/cyc/SM\n", names[mode], nw, (double)mx / iters,
             (double)bytes[mode] * nw * iters / mx);
    }
  return 0;
}
