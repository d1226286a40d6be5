// What does compute-sanitizer's synccheck accept for a named barrier shared by two warps that run different code?
// (the warp-pair kernels: pair_exchange_sum, gp_dynamics.cuh). One variant per process: ./synccheck_probe <A|B|C|D>
//  A  bar.sync id, 64 inlined in two different functions (two program locations)      - what the pair kernels did
//  B  the same barrier inside one __noinline__ function called from both              - one program location
//  C  bar.sync id, 64 from one location, both warps running the same code             - control
//  D  barrier.sync (non-aligned form) inlined in two functions
#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ void bar_fn(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
template <int V> __device__ __forceinline__ void bar(int id) {
  if (V == 1) bar_fn(id);
  else if (V == 3) asm volatile("barrier.sync %0, 64;" ::"r"(id) : "memory");
  else asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
template <int V> __device__ double half0(double* x, int id, int lane) {
  double a = lane;
  for (int i = 0; i < 4; ++i) { x[lane] = a; bar<V>(id); a = a * 1.5 + x[32 + lane]; bar<V>(id); }
  return a;
}
template <int V> __device__ double half1(double* x, int id, int lane) {
  double a = 2 * lane;
  for (int i = 0; i < 4; ++i) { x[32 + lane] = a; bar<V>(id); a = a * 0.5 - x[lane] + 1.0; bar<V>(id); }
  return a;
}
template <int V> __global__ void k(double* out) {
  __shared__ double x[2][64];
  const int warp = threadIdx.x >> 5, pair = warp >> 1, lane = threadIdx.x & 31;
  double r;
  if (V == 2) r = half0<V>(x[pair], 1 + pair, lane) + (warp & 1);  // same code (and a wrong answer: control only)
  else if (warp & 1) r = half1<V>(x[pair], 1 + pair, lane);
  else r = half0<V>(x[pair], 1 + pair, lane);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main(int argc, char** argv) {
  const char v = argc > 1 ? argv[1][0] : 'A';
  double* d; cudaMalloc(&d, 4 * 128 * sizeof(double));
  if (v == 'A') k<0><<<4, 128>>>(d); else if (v == 'B') k<1><<<4, 128>>>(d); else if (v == 'C') k<2><<<4, 128>>>(d); else k<3><<<4, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  double h[128]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("variant %c: %s, out[1]=%g out[33]=%g\n", v, cudaGetErrorString(e), h[1], h[33]);
  return 0;
}
