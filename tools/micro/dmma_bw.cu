// dmma_bw.cu -- micro-benchmark: fp64 FMA pipe (DFMA) vs fp64 tensor instructions (mma.sync ... f64, SASS DMMA) on
// sm_100a, alone and together: do they share a pipe, and what is each worth?  (Decides whether the blocked triangular
// mat-vec of the GP posterior kernel should move to DMMA)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_bw dmma_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode 0 DFMA (8 chains / thread); 1 DMMA m8n8k4 (4 accumulator pairs); 2 DMMA m16n8k16 (2 accumulators);
// 3 DFMA + m8n8k4 interleaved in every warp; 4 even warps DFMA, odd warps m8n8k4; 5 m16n8k8; 6 DFMA + m16n8k16 interleaved
__global__ void __launch_bounds__(256) k(int mode, int iters, double* sink) {
  const int warp = threadIdx.x >> 5;
  double x[8], acc[8], a[8], b[4];
  for (int i = 0; i < 8; ++i) { x[i] = 1.0 + 1e-9 * (threadIdx.x + i); acc[i] = 0.5 * i; a[i] = 1.0 + 1e-7 * i; }
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-7 * i;
  const bool fma_w = mode == 0 || mode == 3 || mode == 6 || (mode == 4 && !(warp & 1));
  const bool mma_w = mode == 1 || mode == 3 || (mode == 4 && (warp & 1));
  for (int it = 0; it < iters; ++it) {
    if (fma_w) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], x[i], 1e-3);
      }
    }
    if (mma_w) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        dmma884(x[0], x[1], a[0], b[0]); dmma884(x[2], x[3], a[1], b[1]);
        dmma884(x[4], x[5], a[2], b[2]); dmma884(x[6], x[7], a[3], b[3]);
      }
    }
    if (mode == 2 || mode == 6) {
#pragma unroll
      for (int r = 0; r < 2; ++r) { dmma16816(x, a, b); dmma16816(x + 4, a, b); }
    }
    if (mode == 5) {
#pragma unroll
      for (int r = 0; r < 4; ++r) { dmma1688(x, a, b); dmma1688(x + 4, a + 4, b + 2); }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i] + x[i];
  if (s == 123.456) sink[0] = s;
}

int main() {
  double* sink; cudaMalloc(&sink, 8);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000, ctas = sms * 4;
  const char* names[] = {"DFMA only", "DMMA m8n8k4 only", "DMMA m16n8k16 only", "DFMA + m8n8k4 in every warp", "even warps DFMA / odd warps m8n8k4",
                         "DMMA m16n8k8 only", "DFMA + m16n8k16 in every warp"};
  for (int mode = 0; mode < 7; ++mode) {
    k<<<ctas, 256>>>(mode, 100, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<ctas, 256>>>(mode, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)ctas * 8, thr = warps * 32;
    double f_fma = 0, f_mma = 0;
    if (mode == 0 || mode == 3 || mode == 6) f_fma = thr * iters * 32.0 * 2;
    if (mode == 4) f_fma = thr / 2 * iters * 32.0 * 2;
    if (mode == 1 || mode == 3) f_mma = warps * iters * 16.0 * 512;
    if (mode == 4) f_mma = warps / 2 * iters * 16.0 * 512;
    if (mode == 2 || mode == 6) f_mma = warps * iters * 4.0 * 4096;
    if (mode == 5) f_mma = warps * iters * 8.0 * 2048;
    printf("mode %d %-36s %8.3f ms  DFMA %7.2f TFLOP/s  DMMA %7.2f TFLOP/s  sum %7.2f  (%s)\n", mode, names[mode], ms, f_fma / ms / 1e9,
           f_mma / ms / 1e9, (f_fma + f_mma) / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
