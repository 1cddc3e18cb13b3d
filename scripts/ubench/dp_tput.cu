// fp64 pipe: cycles per dependent DFMA per warp as a function of warps per SM and ILP (B200)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double b, double c) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __fma_rn(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP> void run(int warps, double *out, long long *cyc) {
  int iters = 512;
  k<ILP><<<148, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double dfma = (double)iters * 16 * ILP;            // per warp
  printf("warps/SM %2d ILP %d: %.2f cycles per DFMA per warp (chain step %.2f), SM throughput %.3f warp-DFMA/cycle\n", warps, ILP,
         h[0] / dfma, h[0] / ((double)iters * 16), dfma * warps / h[0]);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  for (int w : {1, 2, 4, 8, 16, 32}) { run<1>(w, out, cyc); }
  for (int w : {4, 8, 16}) { run<2>(w, out, cyc); run<4>(w, out, cyc); }
  return 0;
}
