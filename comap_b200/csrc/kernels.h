// Launch interfaces of the CUDA kernels (implemented in k*.cu).
#pragma once
#include "common.h"

namespace cmb {

struct DevStream {           // an OpStream uploaded to the device
  DevBuf bytes, off, nbytes, nrec;
  uint32_t n_chunks = 0, cap = 0;
  void upload(const OpStream& s, cudaStream_t st);
  void release();
};

struct MapBuffers {          // per-batch device arrays, n_pad sites (multiple of 256)
  int64_t n = 0, n_pad = 0;
  const uint8_t* tips = nullptr;   // [T][n_pad]
  double* D = nullptr;             // [n_slots][C*A][n_pad]
  double* Lc = nullptr;            // [C][n_pad]
  double* invL = nullptr;          // [n_pad]
  double* loglik = nullptr;        // [n_pad]
  double* post_rate = nullptr;     // [n_pad]
  int32_t* rate_class = nullptr;   // [n_pad]
  double* out = nullptr;           // [B][n_pad]
  double* sum = nullptr;           // [n_pad]  sum_b n_b
  double* sumsq = nullptr;         // [n_pad]  sum_b n_b^2
};

struct MapModel {            // device-resident model constants
  int A = 0, C = 0, B = 0, n_slots = 0;
  const uint32_t* code_mask = nullptr; // [256]
  const double* pi = nullptr;          // [A]
  const double* rates = nullptr;       // [C]
  const double* probs = nullptr;       // [C]
};

// class block [c0, c0+cb) of the down (post-order) pass
void launch_map_down(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, int cb,
                     cudaStream_t st);
void launch_map_finish(const MapModel& m, const MapBuffers& b, cudaStream_t st);
// up (pre-order) pass + contraction; accumulate: add to out instead of overwrite;
// with_norms: also write sum / sumsq (only valid when one block covers all classes)
void launch_map_up(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, int cb,
                   bool accumulate, bool with_norms, cudaStream_t st);
void launch_map_norms(const MapModel& m, const MapBuffers& b, cudaStream_t st);
int map_class_block(int A, int C); // classes per pass for this (A, C)

void launch_simulate(const MapModel& m, const DevStream& s, uint64_t seed, int64_t first_site, int64_t n,
                     int64_t n_pad, int weighted, int root_node, uint8_t* tips, int32_t* classes,
                     cudaStream_t st);

// [B][n_pad] -> [n][B] for the host-facing site-major output
void launch_transpose_out(const double* out, int B, int64_t n, int64_t n_pad, double* dst, cudaStream_t st);

} // namespace cmb
