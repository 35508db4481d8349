// K1: per-site probabilistic substitution mapping on sm_100a.
//
// Replaces DRHomogeneousTreeLikelihood::initialize (Felsenstein post-order + pre-order
// conditional likelihoods) followed by LegacySubstitutionMappingTools::
// computeSubstitutionVectors -- reference call sites CoETools.cpp:209,358-359,397 and
// AnalysisTools.cpp:592-611.
//
// Design (DESIGN.md "K1"):
//   * one thread per site, a block of CB rate classes in registers; 256 sites per CTA;
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

constexpr int NT = 256;

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

// ------------------------------------------------------------------------------ down
template <int A, int CB>
__global__ void __launch_bounds__(NT) k1_down(MapModel m, MapBuffers b, ChunkMeta cm, int c0) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int N = CB * A * A;
  const int64_t site = (int64_t)blockIdx.x * NT + threadIdx.x;
  const int64_t n_pad = b.n_pad;
  ChunkStream cs{cm.src, cm.off, cm.bytes, cm.n_chunks, cm.cap, nullptr, nullptr};
  cs.start(smem + 128, reinterpret_cast<uint64_t*>(smem));

  double cur[CB * A];
  double stk[kMaxStack][CB * A];
  int sp = 0;
#pragma unroll
  for (int i = 0; i < CB * A; i++) cur[i] = 0.;

  for (uint32_t k = 0; k < cm.n_chunks; k++) {
    const unsigned char* rp = cs.wait(k);
    const uint32_t nrec = __ldg(cm.nrec + k);
    for (uint32_t r = 0; r < nrec; r++) {
      const int4 h = *reinterpret_cast<const int4*>(rp);
      const uint32_t flags = (uint32_t)h.x;
      const double* tp = reinterpret_cast<const double*>(rp + 16);
      double prod[CB * A];
      if (flags & kDownTipB) {
        TipInfo t = read_tip(b.tips, m.code_mask, h.z, n_pad, site);
        if (t.fast) tip_column<A, CB>(tp, t.state, prod);
        else {
          double d[CB * A];
          tip_dense<A, CB>(t.mask, d);
          matvec<A, CB>(tp, d, prod);
        }
      } else matvec<A, CB>(tp, cur, prod);
      tp += N;
      if (flags & kDownTipA) {
        double ma[CB * A];
        TipInfo t = read_tip(b.tips, m.code_mask, h.y, n_pad, site);
        if (t.fast) tip_column<A, CB>(tp, t.state, ma);
        else {
          double d[CB * A];
          tip_dense<A, CB>(t.mask, d);
          matvec<A, CB>(tp, d, ma);
        }
        tp += N;
#pragma unroll
        for (int i = 0; i < CB * A; i++) prod[i] *= ma[i];
      } else {
        --sp;
#pragma unroll
        for (int i = 0; i < CB * A; i++) prod[i] *= stk[sp][i];
      }
#pragma unroll
      for (int i = 0; i < CB * A; i++) cur[i] = prod[i];
      if (h.w >= 0) {
        double* d = b.D + ((size_t)h.w * (m.C * A) + (size_t)c0 * A) * n_pad + site;
#pragma unroll
        for (int i = 0; i < CB * A; i++) d[(size_t)i * n_pad] = cur[i];
      }
      if (flags & kDownPush) {
        matvec<A, CB>(tp, cur, stk[sp]);
        ++sp;
        tp += N;
      }
      rp += (reinterpret_cast<const unsigned char*>(tp) - rp + 15) & ~size_t(15);
    }
    cs.release(k);
  }
  // root: class likelihoods L_c = sum_x pi_x root[c][x]
#pragma unroll
  for (int c = 0; c < CB; c++) {
    double l = 0.;
#pragma unroll
    for (int x = 0; x < A; x++) l = fma(cur[c * A + x], __ldg(m.pi + x), l);
    b.Lc[(size_t)(c0 + c) * n_pad + site] = l;
  }
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
template <int A, int CB>
__global__ void __launch_bounds__(NT) k1_up(MapModel m, MapBuffers b, ChunkMeta cm, int c0, int accumulate,
                                            int with_norms) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int N = CB * A * A;
  const int64_t site = (int64_t)blockIdx.x * NT + threadIdx.x;
  const int64_t n_pad = b.n_pad;
  ChunkStream cs{cm.src, cm.off, cm.bytes, cm.n_chunks, cm.cap, nullptr, nullptr};
  cs.start(smem + 128, reinterpret_cast<uint64_t*>(smem));

  double G[CB * A];
  double stk[kMaxStack][CB * A];
  int sp = 0;
#pragma unroll
  for (int c = 0; c < CB; c++)
#pragma unroll
    for (int x = 0; x < A; x++) G[c * A + x] = __ldg(m.pi + x);
  const double invL = b.invL[site];
  double s1 = 0., s2 = 0.;

  for (uint32_t k = 0; k < cm.n_chunks; k++) {
    const unsigned char* rp = cs.wait(k);
    const uint32_t nrec = __ldg(cm.nrec + k);
    for (uint32_t r = 0; r < nrec; r++) {
      const int4 h0 = *reinterpret_cast<const int4*>(rp);
      const int4 h1 = *reinterpret_cast<const int4*>(rp + 16);
      const uint32_t flags = (uint32_t)h0.x;
      const int ref_a = h0.y, ref_b = h0.z, out_a = h0.w, out_b = h1.x;
      const double* Pa = reinterpret_cast<const double*>(rp + 32);
      const double* Wa = Pa + N;
      const double* Pb = Wa + N;
      const double* Wb = Pb + N;
      rp += 32 + (size_t)4 * N * sizeof(double);

      double Da[CB * A], Db[CB * A], Ma[CB * A], Mb[CB * A];
      TipInfo ta{0, 0, false}, tb{0, 0, false};
      const bool tipa = flags & kUpTipA, tipb = flags & kUpTipB;
      // issue all global loads of this node first
      if (tipa) ta = read_tip(b.tips, m.code_mask, ref_a, n_pad, site);
      else {
        const double* d = b.D + ((size_t)ref_a * (m.C * A) + (size_t)c0 * A) * n_pad + site;
#pragma unroll
        for (int i = 0; i < CB * A; i++) Da[i] = d[(size_t)i * n_pad];
      }
      if (tipb) tb = read_tip(b.tips, m.code_mask, ref_b, n_pad, site);
      else {
        const double* d = b.D + ((size_t)ref_b * (m.C * A) + (size_t)c0 * A) * n_pad + site;
#pragma unroll
        for (int i = 0; i < CB * A; i++) Db[i] = d[(size_t)i * n_pad];
      }
      if (tipa && !ta.fast) tip_dense<A, CB>(ta.mask, Da);
      if (tipb && !tb.fast) tip_dense<A, CB>(tb.mask, Db);
      const bool fasta = tipa && ta.fast, fastb = tipb && tb.fast;
      if (fasta) tip_column<A, CB>(Pa, ta.state, Ma); else matvec<A, CB>(Pa, Da, Ma);
      if (fastb) tip_column<A, CB>(Pb, tb.state, Mb); else matvec<A, CB>(Pb, Db, Mb);
      // Ua = G o Mb (stored in Mb), Ub = G o Ma (stored in Ma)
#pragma unroll
      for (int i = 0; i < CB * A; i++) {
        double g = G[i];
        double ua = g * Mb[i], ub = g * Ma[i];
        Mb[i] = ua;
        Ma[i] = ub;
      }
      double (&Ua)[CB * A] = Mb;
      double (&Ub)[CB * A] = Ma;
      // contraction with the reward tables
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
        double* o = b.out + (size_t)out_a * n_pad + site;
        double val = acc * invL;
        if (accumulate) val += *o;
        *o = val;
        s1 += val;
        s2 = fma(val, val, s2);
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
        double* o = b.out + (size_t)out_b * n_pad + site;
        double val = acc * invL;
        if (accumulate) val += *o;
        *o = val;
        s1 += val;
        s2 = fma(val, val, s2);
      }
      // messages for the children that are expanded later
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
    cs.release(k);
  }
  if (with_norms) {
    b.sum[site] = s1;
    b.sumsq[site] = s2;
  }
}

// sum_b n_b and sum_b n_b^2 per site from the finished vectors (multi-block case)
__global__ void k1_norms(MapModel m, MapBuffers b) {
  const int64_t site = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= b.n_pad) return;
  double s1 = 0., s2 = 0.;
  for (int br = 0; br < m.B; br++) {
    double v = b.out[(size_t)br * b.n_pad + site];
    s1 += v;
    s2 = fma(v, v, s2);
  }
  b.sum[site] = s1;
  b.sumsq[site] = s2;
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

template <int A, int CB>
void run_down(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, cudaStream_t st) {
  size_t smem = 128 + 2 * (size_t)s.cap;
  CMB_CUDA(cudaFuncSetAttribute(k1_down<A, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_down<A, CB><<<(unsigned)(b.n_pad / NT), NT, smem, st>>>(m, b, meta_of(s), c0);
  CMB_CUDA(cudaGetLastError());
}
template <int A, int CB>
void run_up(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, bool acc, bool norms,
            cudaStream_t st) {
  size_t smem = 128 + 2 * (size_t)s.cap;
  CMB_CUDA(cudaFuncSetAttribute(k1_up<A, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_up<A, CB><<<(unsigned)(b.n_pad / NT), NT, smem, st>>>(m, b, meta_of(s), c0, acc ? 1 : 0, norms ? 1 : 0);
  CMB_CUDA(cudaGetLastError());
}

} // namespace

int map_class_block(int A, int C) {
  if (A == 4) return C < 4 ? C : 4;
  if (A == 20) return 1;
  fail("mapping kernels are built for A = 4 (nucleotides) and A = 20 (proteins); got A = %d", A);
}

#define CMB_DISPATCH(FN, ...)                                                     \
  do {                                                                            \
    if (m.A == 4 && cb == 4) FN<4, 4>(__VA_ARGS__);                               \
    else if (m.A == 4 && cb == 3) FN<4, 3>(__VA_ARGS__);                          \
    else if (m.A == 4 && cb == 2) FN<4, 2>(__VA_ARGS__);                          \
    else if (m.A == 4 && cb == 1) FN<4, 1>(__VA_ARGS__);                          \
    else if (m.A == 20 && cb == 1) FN<20, 1>(__VA_ARGS__);                        \
    else fail("no mapping kernel for A = %d, class block %d", m.A, cb);           \
  } while (0)

void launch_map_down(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, int cb,
                     cudaStream_t st) {
  if (b.n_pad % NT) fail("internal: n_pad must be a multiple of %d", NT);
  CMB_DISPATCH(run_down, m, b, s, c0, st);
}
void launch_map_up(const MapModel& m, const MapBuffers& b, const DevStream& s, int c0, int cb,
                   bool accumulate, bool with_norms, cudaStream_t st) {
  CMB_DISPATCH(run_up, m, b, s, c0, accumulate, with_norms, st);
}
void launch_map_finish(const MapModel& m, const MapBuffers& b, cudaStream_t st) {
  k1_finish<<<(unsigned)((b.n_pad + 255) / 256), 256, 0, st>>>(m, b);
  CMB_CUDA(cudaGetLastError());
}
void launch_map_norms(const MapModel& m, const MapBuffers& b, cudaStream_t st) {
  k1_norms<<<(unsigned)((b.n_pad + 255) / 256), 256, 0, st>>>(m, b);
  CMB_CUDA(cudaGetLastError());
}
void launch_transpose_out(const double* out, int B, int64_t n, int64_t n_pad, double* dst, cudaStream_t st) {
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((B + 31) / 32)), block(32, 8);
  k_transpose_out<<<grid, block, 0, st>>>(out, B, n, n_pad, dst);
  CMB_CUDA(cudaGetLastError());
}

} // namespace cmb
