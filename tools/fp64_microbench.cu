// fp64 pipe microbenchmark for sm_100a: DFMA throughput vs independent chains per thread and
// warps per SM sub-partition; dependent-issue latency.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static void run(int warps_per_sm, double *d_out) {
  int sms = 148;
  int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid(sms), block(32 * warps_per_sm);
  dfma_kernel<ILP><<<grid, block>>>(d_out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  dfma_kernel<ILP><<<grid, block>>>(d_out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)sms * warps_per_sm * iters * ILP;        // warp instructions
  double clk = 1.965e9;
  double per_smsp_per_clk = inst / (sms * 4) / (ms * 1e-3 * clk);
  double cyc_per_iter = ms * 1e-3 * clk / iters;                // cycles per loop iteration (per warp)
  printf("ILP %2d warps/SM %2d: %.3f ms  DFMA/clk/SMSP %.3f  cycles/iter %.1f  TFLOP/s %.1f\n", ILP, warps_per_sm, ms,
         per_smsp_per_clk, cyc_per_iter, inst * 32 * 2 / (ms * 1e-3) / 1e12);
}

int main() {
  double *d_out; cudaMalloc(&d_out, 148 * 1024 * 8);
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, d_out); run<2>(w, d_out); run<4>(w, d_out); run<8>(w, d_out);
  }
  return 0;
}
