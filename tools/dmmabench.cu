// FP64 tensor-core (DMMA) microbenchmark on B200: throughput and dependent-chain latency of
// mma.sync.aligned.{m8n8k4,m16n8k4,m16n8k8,m16n8k16}.f64 vs independent chains per warp and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int KIND, int ILP>
__global__ void k_dmma(double* out, int iters, long long* cyc) {
  double c2[ILP][2], c4[ILP][4];
  double a[4], b[2];
  for (int i = 0; i < 4; i++) a[i] = 1e-3 * (threadIdx.x + i);
  for (int i = 0; i < 2; i++) b[i] = 1e-3 * (threadIdx.x * 3 + i);
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    c2[i][0] = c2[i][1] = i;
    c4[i][0] = c4[i][1] = c4[i][2] = c4[i][3] = i;
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (KIND == 0) mma884(c2[i], a[0], b[0]);
        if (KIND == 1) { double aa[2] = {a[0], a[1]}; mma1684(c4[i], aa, b[0]); }
        if (KIND == 2) mma1688(c4[i], a, b);
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c2[i][0] + c2[i][1] + c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int ILP>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 2000;
  int threads = 32 * warps_per_sm, blocks = 148, tpb = threads;
  if (threads > 1024) { blocks = 296; tpb = threads / 2; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dmma<KIND, ILP><<<blocks, tpb>>>(out, 10, cyc);
  cudaEventRecord(e0);
  k_dmma<KIND, ILP><<<blocks, tpb>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double flop_per = KIND == 0 ? 2. * 8 * 8 * 4 : KIND == 1 ? 2. * 16 * 8 * 4 : 2. * 16 * 8 * 8;
  const double n = (double)148 * warps_per_sm * iters * 4 * ILP;
  const char* nm = KIND == 0 ? "m8n8k4 " : KIND == 1 ? "m16n8k4" : "m16n8k8";
  printf("%s ILP=%d warps/SM=%2d: %.1f cycles per dependent step (per warp), %.2f TFLOP/s, %.2f Gmma/s per SM\n", nm, ILP,
         warps_per_sm, (double)c / (iters * 4), n * flop_per / (ms * 1e-3) / 1e12, n / (ms * 1e-3) / 148 / 1e9);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 2048 * 8 * 2); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16, 32}) {
    run<0, 1>(w, out, cyc); run<0, 4>(w, out, cyc); run<0, 8>(w, out, cyc);
    run<1, 1>(w, out, cyc); run<1, 4>(w, out, cyc); run<1, 8>(w, out, cyc);
    run<2, 1>(w, out, cyc); run<2, 4>(w, out, cyc);
  }
  return 0;
}
