// Store->remote-poll visibility latency between two SMs, with and without streaming loads in flight.
#include <cstdio>
#include <cuda_runtime.h>
#define ROUNDS 200
// block 0 = producer, block 1 = consumer (cooperative launch -> co-resident, different SMs)
template <int MODE>
__global__ void k(unsigned long long *box, const double *big, long long *tp, long long *tc, double *sink) {
  __shared__ double ring[4096];
  const int lane = threadIdx.x;
  double acc = 0;
  size_t off = (size_t)blockIdx.x * (1 << 24) + lane;
  for (int r = 0; r < ROUNDS; r++) {
    unsigned long long *e = box + (size_t)r * 16;   // 128 B apart
    if (MODE >= 1) {  // streaming traffic: uncoalesced DRAM-missing cp.async, like the ring loader
      for (int q = 0; q < 4; q++) {
        unsigned sa = (unsigned)__cvta_generic_to_shared(&ring[(q * 32 + lane) & 4095]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(big + off + (size_t)q * 4099 * 17));
      }
      asm volatile("cp.async.commit_group;");
      off += 4099 * 64;
    }
    if (blockIdx.x == 0) {
      // some work, then publish
      double x = 1.0 + r;
      for (int q = 0; q < 200; q++) x = __fma_rn(x, 1.0000001, 1e-9);
      acc += x;
      long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (lane == 0) {
        unsigned long long v = (unsigned long long)__double_as_longlong(x);
        if (MODE == 2) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"(v), "l"(v) : "memory");
        else asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"(v), "l"(v) : "memory");
        tp[r] = t;
      }
      // wait a bit so rounds do not overlap
      long long t2 = t; while (t2 - t < 20000) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
    } else {
      unsigned long long v0 = 0, v1 = 0;
      if (lane == 0) {
        do { asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(e)); } while (v0 == 0);
        long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tc[r] = t;
      }
      __syncwarp();
    }
    if (MODE >= 1) asm volatile("cp.async.wait_group 0;");
  }
  sink[blockIdx.x * 32 + lane] = acc + ring[lane];
}
int main() {
  unsigned long long *box; double *big, *sink; long long *tp, *tc;
  cudaMalloc(&box, ROUNDS * 128); cudaMalloc(&big, (size_t)1 << 31); cudaMalloc(&sink, 4096);
  cudaMalloc(&tp, ROUNDS * 8); cudaMalloc(&tc, ROUNDS * 8);
  long long hp[ROUNDS], hc[ROUNDS];
  for (int mode = 0; mode < 2; mode++) {
    cudaMemset(box, 0, ROUNDS * 128);
    void *args[] = {&box, &big, &tp, &tc, &sink};
    if (mode == 0) cudaLaunchCooperativeKernel((void *)k<0>, dim3(2), dim3(32), args, 0, 0);
    else cudaLaunchCooperativeKernel((void *)k<1>, dim3(2), dim3(32), args, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hp, tp, sizeof(hp), cudaMemcpyDeviceToHost); cudaMemcpy(hc, tc, sizeof(hc), cudaMemcpyDeviceToHost);
    double s = 0, mx = 0, mn = 1e9; int cnt = 0;
    for (int r = 20; r < ROUNDS; r++) { double d = (double)(hc[r] - hp[r]); s += d; cnt++; if (d > mx) mx = d; if (d < mn) mn = d; }
    printf("mode %d (%s): store->observed  mean %.0f ns  min %.0f  max %.0f   (%s)\n", mode, mode ? "with DRAM-missing cp.async in flight on both SMs" : "quiet", s / cnt, mn, mx, cudaGetErrorString(e));
  }
  return 0;
}
