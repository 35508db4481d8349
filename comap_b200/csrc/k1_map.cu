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
#include <algorithm>
#include <cstdlib>
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

// Up pass with a producer/consumer pipeline: two producer warps stream (i) the table
// chunks and (ii) the children's partials / tip codes of upcoming nodes into shared-memory
// rings with cp.async.bulk (TMA bulk copies, mbarrier full/empty pairs); G*C consumer
// warps never touch global memory on the read side, so the whole HBM latency is covered
// by the ring depth instead of by occupancy.
struct UpParams {
  ChunkMeta cm;
  const int4* refs;     // per node: flags, ref_a, ref_b, 0 (same values as the record header)
  uint32_t n_nodes;
  int groups, n_stages;
  uint32_t stage_bytes; // bytes of one ring stage
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int A>
__global__ void __launch_bounds__(320) k1_up(MapModel m, MapBuffers b, UpParams up) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int AA = A * A;
  const int C = m.C, groups = up.groups, NS = up.n_stages;
  const int n_cons_warps = groups * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_pad = b.n_pad;
  const int SG = 32 * groups;                    // sites per CTA
  const int64_t site0 = (int64_t)blockIdx.x * SG;
  const int N = C * AA;
  // smem: barriers | red | table chunk buffers | partial ring
  uint64_t* tab_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* tab_empty = tab_full + 2;
  uint64_t* d_full = tab_empty + 2;
  uint64_t* d_empty = d_full + 8;
  double* red = reinterpret_cast<double*>(smem + 256);
  const size_t red_bytes = ((size_t)2 * groups * C * 2 * 32 * sizeof(double) + 127) & ~size_t(127);
  unsigned char* tab_buf = smem + 256 + red_bytes;
  unsigned char* ring = tab_buf + 2 * (size_t)up.cm.cap;
  const uint32_t child_bytes = (uint32_t)(C * A) * SG * sizeof(double);
  const uint32_t tip_off = 2 * child_bytes; // two tip rows of SG bytes after the partials

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) { mbar_init(&tab_full[i], 1); mbar_init(&tab_empty[i], n_cons_warps); }
    for (int i = 0; i < NS; i++) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], n_cons_warps); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == n_cons_warps) {
    // ---- producer 1: table chunks
    if (lane == 0) {
      for (uint32_t k = 0; k < up.cm.n_chunks; k++) {
        if (k >= 2) mbar_wait(&tab_empty[k & 1], ((k >> 1) - 1) & 1);
        const uint32_t nb = __ldg(up.cm.bytes + k);
        mbar_expect_tx(&tab_full[k & 1], nb);
        tma_bulk_g2s(tab_buf + (size_t)(k & 1) * up.cm.cap, up.cm.src + __ldg(up.cm.off + k), nb, &tab_full[k & 1]);
      }
    }
    return;
  }
  if (warp == n_cons_warps + 1) {
    // ---- producer 2: partials and tip codes of node n into stage n % NS
    const int rows = C * A;
    int4 h_next = __ldg(up.refs);
    for (uint32_t n = 0; n < up.n_nodes; n++) {
      const int s = n % NS;
      const int4 h = h_next;
      if (n + 1 < up.n_nodes) h_next = __ldg(up.refs + n + 1);
      if (n >= (uint32_t)NS) mbar_wait(&d_empty[s], ((n / NS) - 1) & 1);
      const uint32_t flags = (uint32_t)h.x;
      unsigned char* st = ring + (size_t)s * up.stage_bytes;
      const uint32_t bytes = ((flags & kUpTipA) ? (uint32_t)SG : child_bytes) + ((flags & kUpTipB) ? (uint32_t)SG : child_bytes);
      if (lane == 0) mbar_expect_tx(&d_full[s], bytes);
      __syncwarp();
      if (flags & kUpTipA) {
        if (lane == 0) tma_bulk_g2s(st + tip_off, b.tips + (size_t)h.y * n_pad + site0, SG, &d_full[s]);
      } else {
        const double* src = b.D + ((size_t)h.y * rows) * n_pad + site0;
        for (int r = lane; r < rows; r += 32)
          tma_bulk_g2s(st + (size_t)r * SG * 8, src + (size_t)r * n_pad, SG * 8, &d_full[s]);
      }
      if (flags & kUpTipB) {
        if (lane == 0) tma_bulk_g2s(st + tip_off + SG, b.tips + (size_t)h.z * n_pad + site0, SG, &d_full[s]);
      } else {
        const double* src = b.D + ((size_t)h.z * rows) * n_pad + site0;
        for (int r = lane; r < rows; r += 32)
          tma_bulk_g2s(st + child_bytes + (size_t)r * SG * 8, src + (size_t)r * n_pad, SG * 8, &d_full[s]);
      }
    }
    return;
  }

  // ---- consumers: warp = (site group g, class c), lane = site
  const int c = warp % C, g = warp / C;
  const int64_t site = site0 + g * 32 + lane;
  double G[A];
  double stk[kMaxStack][A];
  int sp = 0;
#pragma unroll
  for (int x = 0; x < A; x++) G[x] = __ldg(m.pi + x);
  const double invL = b.invL[site];
  uint32_t node = 0;
  const size_t rec_bytes = 32 + (size_t)4 * N * sizeof(double);

  for (uint32_t k = 0; k < up.cm.n_chunks; k++) {
    mbar_wait(&tab_full[k & 1], (k >> 1) & 1);
    const unsigned char* rp = tab_buf + (size_t)(k & 1) * up.cm.cap;
    const uint32_t nrec = __ldg(up.cm.nrec + k);
    for (uint32_t r = 0; r < nrec; r++, node++, rp += rec_bytes) {
      const int4 h0 = *reinterpret_cast<const int4*>(rp);
      const int4 h1 = *reinterpret_cast<const int4*>(rp + 16);
      const uint32_t flags = (uint32_t)h0.x;
      const int out_a = h0.w, out_b = h1.x;
      const double* Pa = reinterpret_cast<const double*>(rp + 32) + c * AA;
      const double* Wa = Pa + N;
      const double* Pb = Wa + N;
      const double* Wb = Pb + N;

      // children data from the ring
      const int s = node % NS;
      mbar_wait(&d_full[s], (node / NS) & 1);
      const unsigned char* st = ring + (size_t)s * up.stage_bytes;
      double Da[A], Db[A], Ma[A], Mb[A];
      TipInfo ta{0, 0, false}, tb{0, 0, false};
      const bool tipa = flags & kUpTipA, tipb = flags & kUpTipB;
      uint32_t code_a = 0, code_b = 0;
      if (tipa) code_a = st[tip_off + g * 32 + lane];
      else {
        const double* d = reinterpret_cast<const double*>(st) + (size_t)(c * A) * SG + g * 32 + lane;
#pragma unroll
        for (int i = 0; i < A; i++) Da[i] = d[(size_t)i * SG];
      }
      if (tipb) code_b = st[tip_off + SG + g * 32 + lane];
      else {
        const double* d = reinterpret_cast<const double*>(st + child_bytes) + (size_t)(c * A) * SG + g * 32 + lane;
#pragma unroll
        for (int i = 0; i < A; i++) Db[i] = d[(size_t)i * SG];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[s]); // stage is in registers now
      if (tipa) ta = tip_from_code(m.code_mask, code_a);
      if (tipb) tb = tip_from_code(m.code_mask, code_b);
      if (tipa && !ta.fast) tip_dense<A, 1>(ta.mask, Da);
      if (tipb && !tb.fast) tip_dense<A, 1>(tb.mask, Db);
      const bool fasta = tipa && ta.fast, fastb = tipb && tb.fast;
      if (fasta) tip_column<A, 1>(Pa, ta.state, Ma); else matvec<A, 1>(Pa, Da, Ma);
      if (fastb) tip_column<A, 1>(Pb, tb.state, Mb); else matvec<A, 1>(Pb, Db, Mb);
#pragma unroll
      for (int i = 0; i < A; i++) {
        double gg = G[i];
        double ua = gg * Mb[i], ub = gg * Ma[i];
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
      double* rb = red + ((size_t)((node & 1) * groups + g) * C) * 64;
      rb[(size_t)c * 64 + lane] = acc_a;
      rb[(size_t)c * 64 + 32 + lane] = acc_b;
      group_barrier(g, 32 * C);
      const int wb = C > 1 ? 1 : 0;
      if (c == 0 && out_a >= 0) {
        double t = 0.;
        for (int cc = 0; cc < C; cc++) t += rb[(size_t)cc * 64 + lane];
        b.out[(size_t)out_a * n_pad + site] = t * invL;
      }
      if (c == wb && out_b >= 0) {
        double t = 0.;
        for (int cc = 0; cc < C; cc++) t += rb[(size_t)cc * 64 + 32 + lane];
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
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&tab_empty[k & 1]);
  }
}

// Up pass, nucleotide fast path: one thread per site with all CB = C rate classes in
// registers (no cross-warp class reduction, ~2.5x fewer instructions per site than the
// warp-per-class kernel), fed by the same producer warps / shared-memory rings.
template <int A, int CB>
__global__ void __launch_bounds__(320, 1) k1_up_site(MapModel m, MapBuffers b, UpParams up) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int N = CB * A * A;
  constexpr int SG = 256;      // sites per CTA = consumer threads
  constexpr int ROWS = CB * A;
  const int NS = up.n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_pad = b.n_pad;
  const int64_t site0 = (int64_t)blockIdx.x * SG;
  uint64_t* tab_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* tab_empty = tab_full + 2;
  uint64_t* d_full = tab_empty + 2;
  uint64_t* d_empty = d_full + 8;
  unsigned char* tab_buf = smem + 256;
  unsigned char* ring = tab_buf + 2 * (size_t)up.cm.cap;
  constexpr uint32_t child_bytes = (uint32_t)ROWS * SG * sizeof(double);
  constexpr uint32_t tip_off = 2 * child_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) { mbar_init(&tab_full[i], 1); mbar_init(&tab_empty[i], 8); }
    for (int i = 0; i < NS; i++) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 8); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 8) { // producer 1: table chunks
    if (lane == 0) {
      for (uint32_t k = 0; k < up.cm.n_chunks; k++) {
        if (k >= 2) mbar_wait(&tab_empty[k & 1], ((k >> 1) - 1) & 1);
        const uint32_t nb = __ldg(up.cm.bytes + k);
        mbar_expect_tx(&tab_full[k & 1], nb);
        tma_bulk_g2s(tab_buf + (size_t)(k & 1) * up.cm.cap, up.cm.src + __ldg(up.cm.off + k), nb, &tab_full[k & 1]);
      }
    }
    return;
  }
  if (warp == 9) { // producer 2: children partials / tip codes of node n into stage n % NS
    int4 h_next = __ldg(up.refs);
    for (uint32_t n = 0; n < up.n_nodes; n++) {
      const int s = n % NS;
      const int4 h = h_next;
      if (n + 1 < up.n_nodes) h_next = __ldg(up.refs + n + 1);
      if (n >= (uint32_t)NS) mbar_wait(&d_empty[s], ((n / NS) - 1) & 1);
      const uint32_t flags = (uint32_t)h.x;
      unsigned char* st = ring + (size_t)s * up.stage_bytes;
      const uint32_t bytes = ((flags & kUpTipA) ? (uint32_t)SG : child_bytes) + ((flags & kUpTipB) ? (uint32_t)SG : child_bytes);
      if (lane == 0) mbar_expect_tx(&d_full[s], bytes);
      __syncwarp();
      if (flags & kUpTipA) {
        if (lane == 0) tma_bulk_g2s(st + tip_off, b.tips + (size_t)h.y * n_pad + site0, SG, &d_full[s]);
      } else if (lane < ROWS) {
        const double* src = b.D + ((size_t)h.y * ROWS + lane) * n_pad + site0;
        tma_bulk_g2s(st + (size_t)lane * SG * 8, src, SG * 8, &d_full[s]);
      }
      if (flags & kUpTipB) {
        if (lane == 0) tma_bulk_g2s(st + tip_off + SG, b.tips + (size_t)h.z * n_pad + site0, SG, &d_full[s]);
      } else if (lane < ROWS) {
        const double* src = b.D + ((size_t)h.z * ROWS + lane) * n_pad + site0;
        tma_bulk_g2s(st + child_bytes + (size_t)lane * SG * 8, src, SG * 8, &d_full[s]);
      }
    }
    return;
  }

  // ---- consumers: thread = site, all classes in registers
  const int t = threadIdx.x;
  const int64_t site = site0 + t;
  double G[CB * A];
  double stk[kMaxStack][CB * A];
  int sp = 0;
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int x = 0; x < A; x++) G[c * A + x] = __ldg(m.pi + x);
  const double invL = b.invL[site];
  uint32_t node = 0;
  constexpr size_t rec_bytes = 32 + (size_t)4 * N * sizeof(double);

  for (uint32_t k = 0; k < up.cm.n_chunks; k++) {
    mbar_wait(&tab_full[k & 1], (k >> 1) & 1);
    const unsigned char* rp = tab_buf + (size_t)(k & 1) * up.cm.cap;
    const uint32_t nrec = __ldg(up.cm.nrec + k);
    for (uint32_t r = 0; r < nrec; r++, node++, rp += rec_bytes) {
      const int4 h0 = *reinterpret_cast<const int4*>(rp);
      const int4 h1 = *reinterpret_cast<const int4*>(rp + 16);
      const uint32_t flags = (uint32_t)h0.x;
      const int out_a = h0.w, out_b = h1.x;
      const double* Pa = reinterpret_cast<const double*>(rp + 32);
      const double* Wa = Pa + N;
      const double* Pb = Wa + N;
      const double* Wb = Pb + N;

      const int s = node % NS;
      mbar_wait(&d_full[s], (node / NS) & 1);
      const unsigned char* st = ring + (size_t)s * up.stage_bytes;
      double Da[CB * A], Db[CB * A], Ma[CB * A], Mb[CB * A];
      TipInfo ta{0, 0, false}, tb{0, 0, false};
      const bool tipa = flags & kUpTipA, tipb = flags & kUpTipB;
      uint32_t code_a = 0, code_b = 0;
      if (tipa) code_a = st[tip_off + t];
      else {
        const double* d = reinterpret_cast<const double*>(st) + t;
#pragma unroll
        for (int i = 0; i < CB * A; i++) Da[i] = d[(size_t)i * SG];
      }
      if (tipb) code_b = st[tip_off + SG + t];
      else {
        const double* d = reinterpret_cast<const double*>(st + child_bytes) + t;
#pragma unroll
        for (int i = 0; i < CB * A; i++) Db[i] = d[(size_t)i * SG];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty[s]);
      if (tipa) ta = tip_from_code(m.code_mask, code_a);
      if (tipb) tb = tip_from_code(m.code_mask, code_b);
      if (tipa && !ta.fast) tip_dense<A, CB>(ta.mask, Da);
      if (tipb && !tb.fast) tip_dense<A, CB>(tb.mask, Db);
      const bool fasta = tipa && ta.fast, fastb = tipb && tb.fast;
      if (fasta) tip_column<A, CB>(Pa, ta.state, Ma); else matvec<A, CB>(Pa, Da, Ma);
      if (fastb) tip_column<A, CB>(Pb, tb.state, Mb); else matvec<A, CB>(Pb, Db, Mb);
#pragma unroll
      for (int i = 0; i < CB * A; i++) {
        double gg = G[i];
        double ua = gg * Mb[i], ub = gg * Ma[i];
        Mb[i] = ua;
        Ma[i] = ub;
      }
      double (&Ua)[CB * A] = Mb;
      double (&Ub)[CB * A] = Ma;
      if (out_a >= 0) {
        double acc = 0.;
        if (fasta) {
#pragma unroll
          for (int i = 0; i < CB * A; i++) acc = fma(Ua[i], Wa[i * A + ta.state], acc);
        } else {
          double wd[CB * A];
          matvec<A, CB>(Wa, Da, wd);
#pragma unroll
          for (int i = 0; i < CB * A; i++) acc = fma(Ua[i], wd[i], acc);
        }
        b.out[(size_t)out_a * n_pad + site] = acc * invL;
      }
      if (out_b >= 0) {
        double acc = 0.;
        if (fastb) {
#pragma unroll
          for (int i = 0; i < CB * A; i++) acc = fma(Ub[i], Wb[i * A + tb.state], acc);
        } else {
          double wd[CB * A];
          matvec<A, CB>(Wb, Db, wd);
#pragma unroll
          for (int i = 0; i < CB * A; i++) acc = fma(Ub[i], wd[i], acc);
        }
        b.out[(size_t)out_b * n_pad + site] = acc * invL;
      }
      if (flags & kUpTakeA) {
        if (flags & kUpPush) {
          matvec_t<A, CB>(Pb, Ub, stk[sp]);
          ++sp;
        }
        matvec_t<A, CB>(Pa, Ua, G);
      } else if (flags & kUpTakeB) {
        matvec_t<A, CB>(Pb, Ub, G);
      } else if (flags & kUpPop) {
        --sp;
#pragma unroll
        for (int i = 0; i < CB * A; i++) G[i] = stk[sp][i];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&tab_empty[k & 1]);
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
template <int A, int CB>
bool run_up_site(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  int dev = 0, max_smem = 0;
  CMB_CUDA(cudaGetDevice(&dev));
  CMB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  UpParams up;
  up.cm = meta_of(s);
  up.refs = s.aux.as<int4>();
  up.n_nodes = s.n_records;
  up.groups = 8;
  const size_t stage = ((size_t)2 * CB * A * 256 * 8 + 2 * 256 + 127) & ~size_t(127);
  const size_t fixed = 256 + 2 * (size_t)s.cap;
  if ((size_t)max_smem < fixed + 2 * stage) return false;
  const int ns = (int)std::min<size_t>(8, ((size_t)max_smem - fixed) / stage);
  up.n_stages = ns;
  up.stage_bytes = (uint32_t)stage;
  const size_t smem = fixed + (size_t)ns * stage;
  CMB_CUDA(cudaFuncSetAttribute(k1_up_site<A, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_up_site<A, CB><<<(unsigned)(b.n_pad / 256), 320, smem, st>>>(m, b, up);
  CMB_CUDA(cudaGetLastError());
  return true;
}

template <int A>
void run_up(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  static const bool site_kernel = getenv("CMB_UP_SITE_KERNEL") != nullptr; // experiment switch
  if constexpr (A == 4) if (site_kernel) { // thread-per-site variant: all classes in one thread
    bool done = false;
    if (m.C == 4) done = run_up_site<4, 4>(m, b, s, st);
    else if (m.C == 3) done = run_up_site<4, 3>(m, b, s, st);
    else if (m.C == 2) done = run_up_site<4, 2>(m, b, s, st);
    else if (m.C == 1) done = run_up_site<4, 1>(m, b, s, st);
    if (done) return;
  }
  // site groups per CTA and ring depth from the shared-memory budget
  int dev = 0, max_smem = 0;
  CMB_CUDA(cudaGetDevice(&dev));
  CMB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int groups = groups_per_cta(m.C);
  UpParams up;
  up.cm = meta_of(s);
  up.refs = s.aux.as<int4>();
  up.n_nodes = s.n_records;
  size_t smem = 0;
  int ns = 0;
  for (;; groups = groups > 1 ? groups / 2 : 1) {
    const size_t red = ((size_t)2 * groups * m.C * 2 * 32 * sizeof(double) + 127) & ~size_t(127);
    const size_t stage = ((size_t)2 * m.C * A * 32 * groups * 8 + 2 * 32 * groups + 127) & ~size_t(127);
    const size_t fixed = 256 + red + 2 * (size_t)s.cap;
    // aim at two resident CTAs per SM when the tables are small, else one
    const size_t budget = fixed + 4 * stage <= (size_t)max_smem / 2 - 1024 ? (size_t)max_smem / 2 - 1024 : (size_t)max_smem;
    ns = (int)std::min<size_t>(8, budget > fixed ? (budget - fixed) / stage : 0);
    if (ns >= 2 || groups == 1) {
      up.stage_bytes = (uint32_t)stage;
      smem = fixed + (size_t)ns * stage;
      break;
    }
  }
  if (ns < 2) fail("mapping up pass: shared memory too small for A = %d, C = %d", A, m.C);
  up.groups = groups;
  up.n_stages = ns;
  CMB_CUDA(cudaFuncSetAttribute(k1_up<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_up<A><<<(unsigned)(b.n_pad / (32 * groups)), 32 * (groups * m.C + 2), smem, st>>>(m, b, up);
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
