// fp64 tensor-core (mma.sync m8n8k4 f64) throughput on sm_100a vs warps per SM and independent
// accumulators per warp.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_kernel(double *out, int iters, double a0, double b0) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 0.5 * i; }
  double a = a0 + threadIdx.x * 1e-6, b = b0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static void run(int warps_per_sm, double *d_out) {
  const int sms = 148, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_kernel<ILP><<<sms, 32 * warps_per_sm>>>(d_out, 100, 1e-3, 1e-3);
  cudaEventRecord(e0);
  dmma_kernel<ILP><<<sms, 32 * warps_per_sm>>>(d_out, iters, 1e-3, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * warps_per_sm * iters * ILP;
  printf("ILP %2d warps/SM %2d: %.3f ms  %.1f TFLOP/s  (%.2f cycles per mma per SM at 1.965 GHz)\n", ILP,
         warps_per_sm, ms, mmas * 512 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (iters * ILP * warps_per_sm));
}

int main() {
  double *d_out; cudaMalloc(&d_out, 148 * 1024 * 8);
  for (int w : {4, 8, 16, 32}) { run<1>(w, d_out); run<2>(w, d_out); run<4>(w, d_out); run<8>(w, d_out); }
  return 0;
}
