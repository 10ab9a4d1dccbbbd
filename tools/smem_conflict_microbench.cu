// Shared-memory wavefronts per LDS.64 / STS.128 for the access patterns of transit_mma_kernel
// (A / B fragment reads, C fragment stores), measured as SM cycles per warp instruction with the
// shared-memory pipe saturated (8 warps, 8 independent accesses per iteration).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/smem_conflict_microbench tools/smem_conflict_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Pat { int off[32]; };

__global__ void lds64_kernel(Pat p, int iters, long long *cycles, double *sink) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + 8u * p.off[threadIdx.x & 31];
  double acc = 0.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + 2048u * u));
      acc += v;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  sink[threadIdx.x] = acc;
}

__global__ void sts128_kernel(Pat p, int iters, long long *cycles, double *sink) {
  extern __shared__ double sm[];
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + 8u * p.off[threadIdx.x & 31];
  const double a = threadIdx.x, b = a + 0.5;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + 8192u * u), "d"(a), "d"(b) : "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  __syncthreads();
  sink[threadIdx.x] = sm[threadIdx.x];
}

template <class K>
static void run(const char *name, K kern, const Pat &p, long long *d_c, double *d_s) {
  const int iters = 4000, warps = 8;
  kern<<<1, 32 * warps, 65536>>>(p, 10, d_c, d_s);
  kern<<<1, 32 * warps, 65536>>>(p, iters, d_c, d_s);
  long long c;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %.2f cycles per warp instruction\n", name, (double)c / ((double)iters * 8 * warps));
}

int main() {
  long long *d_c; double *d_s;
  cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 1024 * 8);
  cudaFuncSetAttribute(lds64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(sts128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  Pat p;
  auto fill = [&](auto f) { for (int l = 0; l < 32; l++) p.off[l] = f(l, l >> 2, l & 3); };
  fill([](int l, int g, int tg) { return l; });
  run("LDS.64 linear", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 20 + tg; });
  run("LDS.64 A fragment (row stride 20)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 36 + tg; });
  run("LDS.64 A fragment (row stride 36)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return tg * 64 + (g ^ (tg << 2)); });
  run("LDS.64 B fragment (xor 4 tg)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return tg * 64 + g; });
  run("LDS.64 B fragment (no swizzle)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return tg * 64 + (g ^ (tg << 3)); });
  run("LDS.64 B fragment (xor 8 tg)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return tg * 72 + g; });
  run("LDS.64 B fragment (row stride 72)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return tg * 68 + g; });
  run("LDS.64 B fragment (row stride 68)", lds64_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return 4 * g + tg; });
  run("LDS.64 B fragment (k-group packed: 4 g + tg)", lds64_kernel, p, d_c, d_s);
  // 16-byte stores: word offsets even
  fill([](int l, int g, int tg) { return 2 * l; });
  run("STS.128 linear", sts128_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 64 + ((2 * tg) ^ ((g & 1) << 3)); });
  run("STS.128 C fragment (xor 8 on odd rows)", sts128_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 64 + 2 * tg; });
  run("STS.128 C fragment (no swizzle)", sts128_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 64 + ((2 * tg) ^ ((g & 3) << 3)); });
  run("STS.128 C fragment (xor 8 (g & 3))", sts128_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 72 + 2 * tg; });
  run("STS.128 C fragment (row stride 72)", sts128_kernel, p, d_c, d_s);
  fill([](int l, int g, int tg) { return g * 8 + 2 * tg; });
  run("STS.128 C fragment (row stride 8 = packed)", sts128_kernel, p, d_c, d_s);
  return 0;
}
