// Streaming-read microbenchmark: which global->shared path sustains HBM bandwidth on B200?
//   mode 0: cp.async.bulk (TMA 1-D bulk), chunk bytes per copy, ring of `depth` chunks, 1 issuing thread
//   mode 1: cp.async (LDGSTS 16 B per thread), producer warps fill a ring
//   mode 2: plain LDG.128 into registers (all threads), summed
// Each CTA streams its own contiguous region of `per_cta` bytes.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok;
  do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory"); } while (!ok);
}
__device__ __forceinline__ void tma_bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}

__global__ void k_bulk(const unsigned char* src, size_t per_cta, uint32_t chunk, int depth, int pieces, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;
  unsigned char* ring = smem + 256;
  if (threadIdx.x == 0) { for (int i = 0; i < depth; i++) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
  const size_t n = per_cta / chunk;
  const uint32_t piece = chunk / pieces;
  if (threadIdx.x < 32) {
    // one warp: lane 0 waits/arms, lanes < pieces issue
    for (size_t i = 0; i < n + depth; i++) {
      if (i >= (size_t)depth) mbar_wait(&bar[i % depth], ((i / depth) - 1) & 1); // data of chunk i-depth landed
      if (i < n) {
        if (threadIdx.x == 0) mbar_expect_tx(&bar[i % depth], chunk);
        __syncwarp();
        if (threadIdx.x < pieces)
          tma_bulk(ring + (size_t)(i % depth) * chunk + threadIdx.x * piece, base + i * chunk + threadIdx.x * piece, piece, &bar[i % depth]);
      }
    }
  }
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = ring[0];
}

__global__ void k_ldgsts(const unsigned char* src, size_t per_cta, uint32_t chunk, int depth, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem + 256;
  const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
  const size_t n = per_cta / chunk;
  // all threads issue 16 B cp.async; commit per chunk; wait_group depth-1
  for (size_t i = 0; i < n; i++) {
    for (uint32_t o = threadIdx.x * 16; o < chunk; o += blockDim.x * 16)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + (size_t)(i % depth) * chunk + o)), "l"(base + i * chunk + o) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 3;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (threadIdx.x == 0 && sink) sink[blockIdx.x] = ring[0];
}

__global__ void k_ldg(const unsigned char* src, size_t per_cta, double* sink) {
  const double2* base = (const double2*)(src + (size_t)blockIdx.x * per_cta);
  const size_t n = per_cta / 16;
  double s = 0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x * 4) {
    double2 a = base[i], b = (i + blockDim.x < n) ? base[i + blockDim.x] : make_double2(0, 0);
    double2 c = (i + 2 * blockDim.x < n) ? base[i + 2 * blockDim.x] : make_double2(0, 0);
    double2 d = (i + 3 * blockDim.x < n) ? base[i + 3 * blockDim.x] : make_double2(0, 0);
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  if (s == 12345.678) sink[blockIdx.x] = s;
}

// write-only streams: each CTA writes its own region; piece = contiguous bytes per row, rows `stride` apart
__global__ void k_write(unsigned char* dst, size_t per_cta, int piece, int stride) {
  double2* base = (double2*)(dst + (size_t)blockIdx.x * per_cta);
  const size_t n = per_cta / 16;
  if (piece == 0) {
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) base[i] = make_double2(1.0, 2.0);
  } else {
    // rows of `piece` bytes at `stride`: thread t writes 8 B of row r (like the down pass: 32 lanes x 8 B x groups)
    const int per_row = piece / 8;
    double* b8 = (double*)base;
    const size_t rows = per_cta / stride;
    for (size_t r = threadIdx.x / per_row; r < rows; r += blockDim.x / per_row)
      b8[r * (stride / 8) + threadIdx.x % per_row] = 1.0;
  }
}

int main(int argc, char** argv) {
  const size_t total = (size_t)8 << 30;
  unsigned char* d; double* sink;
  cudaMalloc(&d, total); cudaMemset(d, 1, total); cudaMalloc(&sink, 1 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms) { printf("%-60s %8.3f ms  %8.1f GB/s\n", name, ms, total / ms / 1e6); fflush(stdout); };
  char name[256];
  for (int mode = 0; mode < 4; mode++) {
    const int grid = 148 * 16;
    const size_t per_cta = total / grid / 65536 * 65536;
    int piece = mode == 0 ? 0 : mode == 1 ? 2048 : mode == 2 ? 512 : 256, stride = 2048;
    k_write<<<grid, 256>>>(d, per_cta, piece, stride);
    cudaEventRecord(e0);
    k_write<<<grid, 256>>>(d, per_cta, piece, stride);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = piece == 0 ? (double)per_cta * grid : (double)(per_cta / stride) * piece * grid;
    printf("write-only piece=%d stride=%d: %8.3f ms %8.1f GB/s\n", piece, stride, ms, bytes / ms / 1e6);
  }
  {
    cudaEventRecord(e0);
    cudaMemsetAsync(d, 0, total);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset: %8.3f ms %8.1f GB/s\n", ms, total / ms / 1e6);
    cudaEventRecord(e0);
    cudaMemcpyAsync(d, d + total / 2, total / 2, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D (read+write bytes): %8.3f ms %8.1f GB/s\n", ms, total / ms / 1e6);
  }
  for (int ctas_per_sm : {1, 2, 4}) {
    const int grid = 148 * ctas_per_sm * 8; // 8 waves
    const size_t per_cta = total / grid / 65536 * 65536;
    for (uint32_t chunk : {2048u, 8192u, 32768u}) for (int depth : {2, 4, 8}) for (int pieces : {1, 4}) {
      size_t smem = 256 + (size_t)chunk * depth;
      if (smem * ctas_per_sm > 220 * 1024) continue;
      cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_bulk<<<grid, 64, smem>>>(d, per_cta, chunk, depth, pieces, sink);
      cudaEventRecord(e0);
      k_bulk<<<grid, 64, smem>>>(d, per_cta, chunk, depth, pieces, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      snprintf(name, sizeof name, "bulk ctas/sm=%d chunk=%u depth=%d pieces=%d (inflight/SM %zu KB)", ctas_per_sm, chunk, depth, pieces, (size_t)chunk * depth * ctas_per_sm / 1024);
      report(name, ms * (float)total / (float)(per_cta * grid));
      cudaError_t e = cudaGetLastError(); if (e) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    }
    for (uint32_t chunk : {8192u}) for (int depth : {4, 8}) {
      size_t smem = 256 + (size_t)chunk * depth;
      if (smem * ctas_per_sm > 220 * 1024) continue;
      cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_ldgsts<<<grid, 128, smem>>>(d, per_cta, chunk, depth, sink);
      cudaEventRecord(e0);
      k_ldgsts<<<grid, 128, smem>>>(d, per_cta, chunk, depth, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      snprintf(name, sizeof name, "ldgsts ctas/sm=%d chunk=%u depth=%d", ctas_per_sm, chunk, depth);
      report(name, ms * (float)total / (float)(per_cta * grid));
    }
    {
      k_ldg<<<grid, 256>>>(d, per_cta, sink);
      cudaEventRecord(e0);
      k_ldg<<<grid, 256>>>(d, per_cta, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      snprintf(name, sizeof name, "ldg.128 x4 grid=%d", grid);
      report(name, ms * (float)total / (float)(per_cta * grid));
    }
  }
  return 0;
}
