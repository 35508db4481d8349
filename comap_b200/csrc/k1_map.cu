// K1: per-site probabilistic substitution mapping on sm_100a.
//
// Replaces DRHomogeneousTreeLikelihood::initialize (Felsenstein post-order + pre-order
// conditional likelihoods) followed by LegacySubstitutionMappingTools::
// computeSubstitutionVectors -- reference call sites CoETools.cpp:209,358-359,397 and
// AnalysisTools.cpp:592-611.
//
// Design (DESIGN.md "K1"):
//   * one warp per (32-site group, rate class): lane = site, so every table read is a
//     warp-uniform 128-bit shared load and every partial access is a coalesced 256-byte
//     row; the C class-warps of a site group sit in the same CTA and combine their
//     class terms through shared memory in a fixed order (deterministic, no atomics);
//   * the tree walk is a precompiled op stream (schedule.cpp) whose records carry the
//     branch transition matrices P_c(b) and reward matrices W_c(b) = p_c P o n; the CTA
//     streams it through shared memory with double-buffered TMA bulk copies, so table
//     reads are warp-uniform 128-bit shared loads;
//   * down pass: each inner node's partial is written to HBM exactly once
//     ([slot][class*A+state][site], site contiguous -> coalesced); the message of the
//     larger child waits on a <= log2(T)-deep per-thread stack;
//   * up pass: each partial is read exactly once; both children of a node are expanded
//     together, the contraction sum_x Up[x] sum_y W[x][y] D[y] is fused, and the vector
//     entry is written as out[branch][site] (site contiguous);
//   * nothing is accumulated with atomics: results are deterministic.
#include "device_utils.cuh"
#include "kernels.h"

namespace cmb {

namespace {

struct ChunkMeta {
  const unsigned char* src;
  const uint32_t *off, *bytes, *nrec;
  uint32_t n_chunks, cap;
};

template <int A>
__device__ __forceinline__ void load_row(const double* __restrict__ row, double (&r)[A]) {
  if constexpr (A % 2 == 0) {
    const double2* p = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int i = 0; i < A / 2; i++) {
      double2 v = p[i];
      r[2 * i] = v.x;
      r[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < A; i++) r[i] = row[i];
  }
}

// o[c][x] = sum_y T[c][x][y] v[c][y]
template <int A, int CB>
__device__ __forceinline__ void matvec(const double* __restrict__ T, const double (&v)[CB * A],
                                       double (&o)[CB * A]) {
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int x = 0; x < A; x++) {
      double row[A];
      load_row<A>(T + (c * A + x) * A, row);
      double s = row[0] * v[c * A];
#pragma unroll
      for (int y = 1; y < A; y++) s = fma(row[y], v[c * A + y], s);
      o[c * A + x] = s;
    }
}

// o[c][x] = sum_y T[c][y][x] v[c][y]   (message travelling down the edge)
template <int A, int CB>
__device__ __forceinline__ void matvec_t(const double* __restrict__ T, const double (&v)[CB * A],
                                         double (&o)[CB * A]) {
#pragma unroll
  for (int c = 0; c < CB; c++) {
#pragma unroll
    for (int y = 0; y < A; y++) {
      double row[A];
      load_row<A>(T + (c * A + y) * A, row);
#pragma unroll
      for (int x = 0; x < A; x++) o[c * A + x] = (y == 0) ? row[x] * v[c * A] : fma(row[x], v[c * A + y], o[c * A + x]);
    }
  }
}

// column pick for a resolved tip: o[c][x] = T[c][x][state]
template <int A, int CB>
__device__ __forceinline__ void tip_column(const double* __restrict__ T, int state, double (&o)[CB * A]) {
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int x = 0; x < A; x++) o[c * A + x] = T[(c * A + x) * A + state];
}

template <int A, int CB>
__device__ __forceinline__ void tip_dense(uint32_t mask, double (&d)[CB * A]) {
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int y = 0; y < A; y++) d[c * A + y] = (mask >> y) & 1u ? 1. : 0.;
}

struct TipInfo {
  uint32_t mask;
  int state;
  bool fast; // warp-uniform: every lane's tip is a single resolved state
};
__device__ __forceinline__ TipInfo read_tip(const uint8_t* __restrict__ tips, const uint32_t* __restrict__ code_mask,
                                            int row, int64_t n_pad, int64_t site) {
  TipInfo t;
  t.mask = __ldg(code_mask + tips[(size_t)row * n_pad + site]);
  bool single = t.mask != 0 && (t.mask & (t.mask - 1)) == 0;
  t.state = __ffs(t.mask) - 1;
  t.fast = __all_sync(0xffffffffu, single);
  return t;
}

__device__ __forceinline__ TipInfo tip_from_code(const uint32_t* __restrict__ code_mask, uint32_t code) {
  TipInfo t;
  t.mask = __ldg(code_mask + code);
  bool single = t.mask != 0 && (t.mask & (t.mask - 1)) == 0;
  t.state = __ffs(t.mask) - 1;
  t.fast = __all_sync(0xffffffffu, single);
  return t;
}

// ------------------------------------------------------------------------------ down
struct WarpMap {
  int c;        // rate class of this warp
  int g;        // site group inside the CTA
  int64_t site; // site of this lane
};
__device__ __forceinline__ WarpMap warp_map(int C, int groups) {
  const int w = threadIdx.x >> 5;
  WarpMap m;
  m.c = w % C;
  m.g = w / C;
  m.site = ((int64_t)blockIdx.x * groups + m.g) * 32 + (threadIdx.x & 31);
  return m;
}

template <int A>
__global__ void __launch_bounds__(256) k1_down(MapModel m, MapBuffers b, ChunkMeta cm, int groups) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int AA = A * A;
  const WarpMap wm = warp_map(m.C, groups);
  const int64_t site = wm.site, n_pad = b.n_pad;
  const int N = m.C * AA; // doubles per table (all classes)
  ChunkStream cs{cm.src, cm.off, cm.bytes, cm.n_chunks, cm.cap, nullptr, nullptr};
  cs.start(smem + 128, reinterpret_cast<uint64_t*>(smem));

  double cur[A];
  double stk[kMaxStack][A];
  int sp = 0;
#pragma unroll
  for (int i = 0; i < A; i++) cur[i] = 0.;

  // software pipeline: the tip codes of the next record are loaded while this one computes
  auto fetch_tips = [&](const unsigned char* rec, uint32_t& ca, uint32_t& cb) {
    const int4 h = *reinterpret_cast<const int4*>(rec);
    ca = ((uint32_t)h.x & kDownTipA) ? b.tips[(size_t)h.y * n_pad + site] : 0u;
    cb = ((uint32_t)h.x & kDownTipB) ? b.tips[(size_t)h.z * n_pad + site] : 0u;
  };
  uint32_t code_a = 0, code_b = 0;
  const unsigned char* rp = cs.wait(0);
  fetch_tips(rp, code_a, code_b);
  for (uint32_t k = 0; k < cm.n_chunks; k++) {
    const uint32_t nrec = __ldg(cm.nrec + k);
    const unsigned char* next_chunk = nullptr;
    for (uint32_t r = 0; r < nrec; r++) {
      const int4 h = *reinterpret_cast<const int4*>(rp);
      const uint32_t flags = (uint32_t)h.x;
      const int ntab = 1 + ((flags & kDownTipA) ? 1 : 0) + ((flags & kDownPush) ? 1 : 0);
      const unsigned char* rp_next = rp + ((16 + (size_t)ntab * N * sizeof(double) + 15) & ~size_t(15));
      if (r + 1 == nrec) rp_next = next_chunk = (k + 1 < cm.n_chunks) ? cs.wait(k + 1) : nullptr;
      uint32_t nca = 0, ncb = 0;
      if (rp_next) fetch_tips(rp_next, nca, ncb);
      const double* tp = reinterpret_cast<const double*>(rp + 16) + wm.c * AA;
      double prod[A];
      if (flags & kDownTipB) {
        TipInfo t = tip_from_code(m.code_mask, code_b);
        if (t.fast) tip_column<A, 1>(tp, t.state, prod);
        else {
          double d[A];
          tip_dense<A, 1>(t.mask, d);
          matvec<A, 1>(tp, d, prod);
        }
      } else matvec<A, 1>(tp, cur, prod);
      tp += N;
      if (flags & kDownTipA) {
        double ma[A];
        TipInfo t = tip_from_code(m.code_mask, code_a);
        if (t.fast) tip_column<A, 1>(tp, t.state, ma);
        else {
          double d[A];
          tip_dense<A, 1>(t.mask, d);
          matvec<A, 1>(tp, d, ma);
        }
        tp += N;
#pragma unroll
        for (int i = 0; i < A; i++) prod[i] *= ma[i];
      } else {
        --sp;
#pragma unroll
        for (int i = 0; i < A; i++) prod[i] *= stk[sp][i];
      }
#pragma unroll
      for (int i = 0; i < A; i++) cur[i] = prod[i];
      if (h.w >= 0) {
        double* d = b.D + ((size_t)h.w * (m.C * A) + (size_t)wm.c * A) * n_pad + site;
#pragma unroll
        for (int i = 0; i < A; i++) d[(size_t)i * n_pad] = cur[i];
      }
      if (flags & kDownPush) {
        matvec<A, 1>(tp, cur, stk[sp]);
        ++sp;
      }
      rp = rp_next;
      code_a = nca;
      code_b = ncb;
    }
    cs.release(k);
  }
  // root: class likelihood L_c = sum_x pi_x root[c][x]
  double l = 0.;
#pragma unroll
  for (int x = 0; x < A; x++) l = fma(cur[x], __ldg(m.pi + x), l);
  b.Lc[(size_t)wm.c * n_pad + site] = l;
}

// ---------------------------------------------------------------------------- finish
// Site likelihood, log-likelihood, posterior rate, rate class with maximal likelihood
// (getLogLikelihoodPerSite / getPosteriorRatePerSite / getRateClassWithMaxPostProbPerSite,
// CoETools.cpp:507-510,669-670).
__global__ void k1_finish(MapModel m, MapBuffers b) {
  const int64_t site = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= b.n_pad) return;
  double L = 0.;
  for (int c = 0; c < m.C; c++) L += b.Lc[(size_t)c * b.n_pad + site] * __ldg(m.probs + c);
  double pr = 0., best = 0.;
  int rc = 0;
  for (int c = 0; c < m.C; c++) {
    double l = b.Lc[(size_t)c * b.n_pad + site];
    pr += (l / L) * __ldg(m.probs + c) * __ldg(m.rates + c);
    if (c == 0 || l > best) { best = l; rc = c; }
  }
  b.invL[site] = 1. / L;
  b.loglik[site] = log(L);
  b.post_rate[site] = pr;
  b.rate_class[site] = rc;
}

// -------------------------------------------------------------------------------- up
__device__ __forceinline__ void group_barrier(int g, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory");
}

template <int A>
__global__ void __launch_bounds__(256) k1_up(MapModel m, MapBuffers b, ChunkMeta cm, int groups) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int AA = A * A;
  const int C = m.C;
  const WarpMap wm = warp_map(C, groups);
  const int lane = threadIdx.x & 31;
  const int64_t site = wm.site, n_pad = b.n_pad;
  const int N = C * AA;
  // smem: [0,128) mbarriers | red[2][groups][C][2][32] doubles | chunk buffers
  double* red = reinterpret_cast<double*>(smem + 128);
  const size_t red_bytes = ((size_t)2 * groups * C * 2 * 32 * sizeof(double) + 127) & ~size_t(127);
  ChunkStream cs{cm.src, cm.off, cm.bytes, cm.n_chunks, cm.cap, nullptr, nullptr};
  cs.start(smem + 128 + red_bytes, reinterpret_cast<uint64_t*>(smem));

  double G[A];
  double stk[kMaxStack][A];
  int sp = 0;
#pragma unroll
  for (int x = 0; x < A; x++) G[x] = __ldg(m.pi + x);
  const double invL = b.invL[site];
  uint32_t node = 0;

  // software pipeline: the partials (or tip codes) of the next record are in flight while
  // this record computes
  const size_t rec_bytes = 32 + (size_t)4 * N * sizeof(double);
  auto fetch = [&](const unsigned char* rec, double (&da)[A], double (&db)[A], uint32_t& ca, uint32_t& cb) {
    const int4 h0 = *reinterpret_cast<const int4*>(rec);
    if ((uint32_t)h0.x & kUpTipA) ca = b.tips[(size_t)h0.y * n_pad + site];
    else {
      const double* d = b.D + ((size_t)h0.y * (C * A) + (size_t)wm.c * A) * n_pad + site;
#pragma unroll
      for (int i = 0; i < A; i++) da[i] = d[(size_t)i * n_pad];
    }
    if ((uint32_t)h0.x & kUpTipB) cb = b.tips[(size_t)h0.z * n_pad + site];
    else {
      const double* d = b.D + ((size_t)h0.z * (C * A) + (size_t)wm.c * A) * n_pad + site;
#pragma unroll
      for (int i = 0; i < A; i++) db[i] = d[(size_t)i * n_pad];
    }
  };
  double Da[A], Db[A];
  uint32_t code_a = 0, code_b = 0;
  const unsigned char* rp = cs.wait(0);
  fetch(rp, Da, Db, code_a, code_b);

  for (uint32_t k = 0; k < cm.n_chunks; k++) {
    const uint32_t nrec = __ldg(cm.nrec + k);
    for (uint32_t r = 0; r < nrec; r++, node++) {
      const int4 h0 = *reinterpret_cast<const int4*>(rp);
      const int4 h1 = *reinterpret_cast<const int4*>(rp + 16);
      const uint32_t flags = (uint32_t)h0.x;
      const int out_a = h0.w, out_b = h1.x;
      const double* Pa = reinterpret_cast<const double*>(rp + 32) + wm.c * AA;
      const double* Wa = Pa + N;
      const double* Pb = Wa + N;
      const double* Wb = Pb + N;
      const unsigned char* rp_next = rp + rec_bytes;
      if (r + 1 == nrec) rp_next = (k + 1 < cm.n_chunks) ? cs.wait(k + 1) : nullptr;
      double nDa[A], nDb[A];
      uint32_t nca = 0, ncb = 0;
      if (rp_next) fetch(rp_next, nDa, nDb, nca, ncb);

      double Ma[A], Mb[A];
      TipInfo ta{0, 0, false}, tb{0, 0, false};
      const bool tipa = flags & kUpTipA, tipb = flags & kUpTipB;
      if (tipa) ta = tip_from_code(m.code_mask, code_a);
      if (tipb) tb = tip_from_code(m.code_mask, code_b);
      if (tipa && !ta.fast) tip_dense<A, 1>(ta.mask, Da);
      if (tipb && !tb.fast) tip_dense<A, 1>(tb.mask, Db);
      const bool fasta = tipa && ta.fast, fastb = tipb && tb.fast;
      if (fasta) tip_column<A, 1>(Pa, ta.state, Ma); else matvec<A, 1>(Pa, Da, Ma);
      if (fastb) tip_column<A, 1>(Pb, tb.state, Mb); else matvec<A, 1>(Pb, Db, Mb);
#pragma unroll
      for (int i = 0; i < A; i++) {
        double g = G[i];
        double ua = g * Mb[i], ub = g * Ma[i];
        Mb[i] = ua;
        Ma[i] = ub;
      }
      double (&Ua)[A] = Mb;
      double (&Ub)[A] = Ma;
      // class term of the contraction with the reward tables (W includes p_c)
      double acc_a = 0., acc_b = 0.;
      if (out_a >= 0) {
        if (fasta) {
#pragma unroll
          for (int i = 0; i < A; i++) acc_a = fma(Ua[i], Wa[i * A + ta.state], acc_a);
        } else {
          double wd[A];
          matvec<A, 1>(Wa, Da, wd);
#pragma unroll
          for (int i = 0; i < A; i++) acc_a = fma(Ua[i], wd[i], acc_a);
        }
      }
      if (out_b >= 0) {
        if (fastb) {
#pragma unroll
          for (int i = 0; i < A; i++) acc_b = fma(Ub[i], Wb[i * A + tb.state], acc_b);
        } else {
          double wd[A];
          matvec<A, 1>(Wb, Db, wd);
#pragma unroll
          for (int i = 0; i < A; i++) acc_b = fma(Ub[i], wd[i], acc_b);
        }
      }
      // combine the C class terms of this site group in class order; warp 0 of the group
      // finishes branch a, warp 1 (if any) branch b
      double* rb = red + ((size_t)((node & 1) * groups + wm.g) * C) * 64;
      rb[(size_t)wm.c * 64 + lane] = acc_a;
      rb[(size_t)wm.c * 64 + 32 + lane] = acc_b;
      group_barrier(wm.g, 32 * C);
      const int wb = C > 1 ? 1 : 0;
      if (wm.c == 0 && out_a >= 0) {
        double t = 0.;
        for (int c = 0; c < C; c++) t += rb[(size_t)c * 64 + lane];
        b.out[(size_t)out_a * n_pad + site] = t * invL;
      }
      if (wm.c == wb && out_b >= 0) {
        double t = 0.;
        for (int c = 0; c < C; c++) t += rb[(size_t)c * 64 + 32 + lane];
        b.out[(size_t)out_b * n_pad + site] = t * invL;
      }
      // messages for the children that are expanded later
      if (flags & kUpTakeA) {
        if (flags & kUpPush) {
          matvec_t<A, 1>(Pb, Ub, stk[sp]);
          ++sp;
        }
        matvec_t<A, 1>(Pa, Ua, G);
      } else if (flags & kUpTakeB) {
        matvec_t<A, 1>(Pb, Ub, G);
      } else if (flags & kUpPop) {
        --sp;
#pragma unroll
        for (int i = 0; i < A; i++) G[i] = stk[sp][i];
      }
      rp = rp_next;
      code_a = nca;
      code_b = ncb;
#pragma unroll
      for (int i = 0; i < A; i++) { Da[i] = nDa[i]; Db[i] = nDb[i]; }
    }
    cs.release(k);
  }
}

__global__ void k_transpose_out(const double* __restrict__ out, int B, int64_t n, int64_t n_pad,
                                double* __restrict__ dst) {
  __shared__ double tile[32][33];
  int64_t s0 = (int64_t)blockIdx.x * 32;
  int b0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int br = b0 + i;
    int64_t s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (br < B && s < n) ? out[(size_t)br * n_pad + s] : 0.;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t s = s0 + i;
    int br = b0 + threadIdx.x;
    if (s < n && br < B) dst[(size_t)s * B + br] = tile[threadIdx.x][i];
  }
}

ChunkMeta meta_of(const DevStream& s) {
  return ChunkMeta{s.bytes.as<unsigned char>(), s.off.as<uint32_t>(), s.nbytes.as<uint32_t>(),
                   s.nrec.as<uint32_t>(), s.n_chunks, s.cap};
}

int groups_per_cta(int C) { return C >= 8 ? 1 : 8 / C; }

template <int A>
void run_down(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  const int groups = groups_per_cta(m.C);
  size_t smem = 128 + 2 * (size_t)s.cap;
  CMB_CUDA(cudaFuncSetAttribute(k1_down<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_down<A><<<(unsigned)(b.n_pad / (32 * groups)), 32 * groups * m.C, smem, st>>>(m, b, meta_of(s), groups);
  CMB_CUDA(cudaGetLastError());
}
template <int A>
void run_up(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  const int groups = groups_per_cta(m.C);
  size_t red = ((size_t)2 * groups * m.C * 2 * 32 * sizeof(double) + 127) & ~size_t(127);
  size_t smem = 128 + red + 2 * (size_t)s.cap;
  CMB_CUDA(cudaFuncSetAttribute(k1_up<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_up<A><<<(unsigned)(b.n_pad / (32 * groups)), 32 * groups * m.C, smem, st>>>(m, b, meta_of(s), groups);
  CMB_CUDA(cudaGetLastError());
}

} // namespace

void check_map_support(int A, int C) {
  if (A != 4 && A != 20)
    fail("mapping kernels are built for A = 4 (nucleotides) and A = 20 (proteins); got A = %d", A);
  if (C < 1 || C > 8) fail("mapping kernels support 1..8 rate classes; got C = %d", C);
}

void launch_map_down(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (b.n_pad % 256) fail("internal: n_pad must be a multiple of 256");
  if (m.A == 4) run_down<4>(m, b, s, st);
  else if (m.A == 20) run_down<20>(m, b, s, st);
  else fail("no mapping kernel for A = %d", m.A);
}
void launch_map_up(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.A == 4) run_up<4>(m, b, s, st);
  else if (m.A == 20) run_up<20>(m, b, s, st);
  else fail("no mapping kernel for A = %d", m.A);
}
void launch_map_finish(const MapModel& m, const MapBuffers& b, cudaStream_t st) {
  k1_finish<<<(unsigned)((b.n_pad + 255) / 256), 256, 0, st>>>(m, b);
  CMB_CUDA(cudaGetLastError());
}
void launch_transpose_out(const double* out, int B, int64_t n, int64_t n_pad, double* dst, cudaStream_t st) {
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((B + 31) / 32)), block(32, 8);
  k_transpose_out<<<grid, block, 0, st>>>(out, B, n, n_pad, dst);
  CMB_CUDA(cudaGetLastError());
}

} // namespace cmb
