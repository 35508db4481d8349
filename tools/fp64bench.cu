// FP64 pipe microbenchmark on B200: dependent-chain latency and throughput vs ILP and warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 2000;
  int threads = 32 * warps_per_sm;
  int blocks = 148, tpb = threads;
  if (threads > 1024) { blocks = 148 * 2; tpb = threads / 2; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<ILP><<<blocks, tpb>>>(out, 10, 1.0000001, 1e-9, cyc);
  cudaEventRecord(e0);
  k_dfma<ILP><<<blocks, tpb>>>(out, iters, 1.0000001, 1e-9, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double n_fma = (double)148 * threads * iters * 8 * ILP;
  printf("ILP=%d warps/SM=%2d: %.2f cycles per dependent DFMA step (per warp), %.1f TFLOP/s\n", ILP, warps_per_sm,
         (double)c / (iters * 8), 2 * n_fma / (ms * 1e-3) / 1e12);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 2048 * 8 * 2); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16, 32, 64}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  return 0;
}
