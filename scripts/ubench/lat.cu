// Microbenchmarks: dependent-chain latencies of fp64 ops, shuffles, smem, L2 loads on B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dfma(double *o, double a, double b, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __fma_rn(x, a, b);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double *o, double a, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __dadd_rn(x, a);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmul(double *o, double a, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __dmul_rn(x, a);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ddiv(double *o, double a, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = __ddiv_rn(x, a) + 1.0;
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *o, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_up_sync(0xffffffffu, x, 1);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_add(double *o, double a, long long *cyc) {
  double x = o[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __dadd_rn(__shfl_up_sync(0xffffffffu, x, 1), a);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double *o, long long *cyc) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 33) & 1023;
  __syncthreads();
  int x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = s[x];
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_l2(const int *p, double *o, long long *cyc) {
  int x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 512; i++) x = *((volatile const int *)(p + x));
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_l2cg(const int *p, double *o, long long *cyc) {
  int x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 512; i++) x = __ldcg(p + x);
  long long t1 = clock64();
  o[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: many warps of independent DFMA
__global__ void k_dfma_tp(double *o, double a, double b) {
  double x0 = o[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < N; i++) {
    x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
    x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
  }
  o[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  double *o; long long *c, hc; int *p;
  cudaMalloc(&o, 1 << 24); cudaMalloc(&c, 64); cudaMalloc(&p, 1 << 20);
  cudaMemset(o, 0, 1 << 24);
  int hp[1 << 18]; for (int i = 0; i < (1 << 18); i++) hp[i] = (i + 4099) & ((1 << 18) - 1);
  cudaMemcpy(p, hp, sizeof(hp), cudaMemcpyHostToDevice);
#define RUN(name, launch, div) for (int r = 0; r < 2; r++) { launch; cudaDeviceSynchronize(); } cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("%-28s %8.2f cycles/op\n", name, (double)hc / (div));
  RUN("DFMA dependent (1 warp)", (k_dfma<<<1, 32>>>(o, 1.0000001, 1e-9, c)), N)
  RUN("DADD dependent (1 warp)", (k_dadd<<<1, 32>>>(o, 1e-9, c)), N)
  RUN("DMUL dependent (1 warp)", (k_dmul<<<1, 32>>>(o, 1.0000001, c)), N)
  RUN("DDIV+DADD dependent", (k_ddiv<<<1, 32>>>(o, 1.0000001, c)), N)
  RUN("SHFL.64 dependent", (k_shfl<<<1, 32>>>(o, c)), N)
  RUN("SHFL.64+DADD dependent", (k_shfl_add<<<1, 32>>>(o, 1e-9, c)), N)
  RUN("LDS dependent", (k_lds<<<1, 32>>>(o, c)), N)
  RUN("LDG volatile (L2) dependent", (k_l2<<<1, 32>>>(p, o, c)), 512)
  RUN("LDG .cg (L2) dependent", (k_l2cg<<<1, 32>>>(p, o, c)), 512)
  RUN("DFMA dependent (4 warps/SM)", (k_dfma<<<1, 128>>>(o, 1.0000001, 1e-9, c)), N)
  RUN("DFMA dependent (16 warps/SM)", (k_dfma<<<1, 512>>>(o, 1.0000001, 1e-9, c)), N)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_tp<<<148 * 8, 256>>>(o, 1.0000001, 1e-9);
  cudaEventRecord(e0); k_dfma_tp<<<148 * 8, 256>>>(o, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("DFMA throughput: %.2f TFLOP/s\n", 2.0 * 148 * 8 * 256 * 8.0 * N / (ms * 1e-3) / 1e12);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock rate attr %d kHz\n", clk);
  return 0;
}
