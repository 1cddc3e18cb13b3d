// Does a relaxed/strong L2 load issued before a ~400-cycle dependent FP64 chain hide behind it?
#include <cstdio>
#include <cuda_runtime.h>
#define N 2000
template <int MODE>
__global__ void k(unsigned long long *buf, double *o, long long *cyc, double a, double b) {
  const int w = blockIdx.x;
  unsigned long long *mine = buf + (size_t)w * 4096 + threadIdx.x * 4;   // per-lane 32 B
  unsigned long long *outp = buf + (size_t)(w + 1024) * 4096 + threadIdx.x * 4;
  double x = o[threadIdx.x];
  unsigned long long p0 = 0, p1 = 0, p2 = 0, p3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < N; i++) {
    // consume previously loaded values at the top (like the mailbox check)
    if ((p0 | p1 | p2 | p3) == 0xFFF8DEADBEEF0001ull) x += 1.0;
    if (MODE >= 1) {
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(p0), "=l"(p1) : "l"(mine));
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(p2), "=l"(p3) : "l"(mine + 2));
    }
#pragma unroll
    for (int q = 0; q < 50; q++) x = __fma_rn(x, a, b);
    if (MODE >= 2) {
      unsigned long long v = (unsigned long long)__double_as_longlong(x);
      asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(outp), "l"(v), "l"(v) : "memory");
      asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(outp + 2), "l"(v), "l"(v) : "memory");
    }
    if (MODE >= 3) { mine += 4 * 32; outp += 4 * 32; if (i % 16 == 15) { mine -= 4 * 32 * 16; outp -= 4 * 32 * 16; } }
  }
  long long t1 = clock64();
  o[threadIdx.x] = x;
  if (threadIdx.x == 0 && w == 0) cyc[0] = t1 - t0;
}
int main() {
  unsigned long long *buf; double *o; long long *c, hc;
  cudaMalloc(&buf, (size_t)4096 * 4096 * 8); cudaMemset(buf, 0, (size_t)4096 * 4096 * 8);
  cudaMalloc(&o, 4096); cudaMemset(o, 0, 4096); cudaMalloc(&c, 8);
#define RUN(M, G) for (int r = 0; r < 2; r++) { k<M><<<G, 32>>>(buf, o, c, 1.0000001, 1e-9); cudaDeviceSynchronize(); } cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost); printf("mode %d grid %4d: %.1f cycles/iter\n", M, G, (double)hc / N);
  RUN(0, 1) RUN(1, 1) RUN(2, 1) RUN(3, 1)
  RUN(0, 512) RUN(1, 512) RUN(2, 512) RUN(3, 512)
  return 0;
}
