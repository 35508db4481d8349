// K1 for nucleotides (A = 4) on the FP64 tensor cores (DMMA m8n8k4): down pass, up pass +
// contraction.
//
// Contract: Bio++ DRHomogeneousTreeLikelihood::initialize (computeSubtreeLikelihoodPostfix /
// Prefix) + LegacySubstitutionMappingTools::computeSubstitutionVectors; reference call sites
// CoETools.cpp:209,358-359,395-397, AnalysisTools.cpp:592-611; SURVEY.md s3.3, s8 a1/a4.
// The first, thread-per-site design (now k1_map.cu, proteins only) was bound by shared-memory
// bandwidth and issue slots: every lane re-read every 4x4 table as warp-uniform broadcast loads
// (ncu r1f: LSU 66 %, FP64 pipe 21 %).  Here a warp owns 8*NG sites and ALL rate classes; lane =
// (site s = lane / 4, state q = lane % 4) holds ONE double per 4-vector, and every 4x4
// matrix-vector product of 8 sites is one DMMA.8x8x4 whose B operand (the table) is one double
// per lane:
//   MMA1/2  D_child (8 sites x 4) x [P | W]      -> lane (s, q): S[q] = (P D)[q], T[q] = (W D)[q]
//   U_a = G o S_b, U_b = G o S_a;   n_a += U_a . T_a  (per lane product, summed over classes
//   in registers, then over q with a two-stage transposing shuffle reduction per node)
//   MMA3/4  U (8 sites x 4) x P^T                -> lane (s, q): message to an inner child
// B fragments come precomputed in the op stream (schedule.cpp build_up_mma_stream /
// build_down_mma_stream), so a table costs 256 B of shared-memory reads per warp instead of 32
// broadcast rows.  Resolved tips are column picks from the raw 4x4 tables (conflict-free);
// ambiguity codes fall back to 0/1 mask partials through the DMMA.  Node bodies are specialised
// by what the two children are, so the C x NG independent DMMA chains of a node interleave.
// DMMA.8x8x4 runs at the FP64 pipe's full rate (tools/dmmabench: 37 TFLOP/s, 4 cycles per SM):
// ~2-3 DMMA per (node, class, 8 sites).
//
// Partials: [128-site chunk][slot][class][site][state] -- a lane reads / writes its double at
// consecutive addresses (256 B per DMMA A operand) and one TMA bulk copy moves a child's whole
// chunk (C * 4 KB) into the stage ring.  Measured on B200 at config 4 (517 k sites): down 4.2 ms
// (5.1 TB/s of stores), up 10.8 ms.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "device_utils.cuh"
#include "kernels.h"

namespace cmb {

namespace {

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(0.), "d"(0.));
}

struct UpMmaParams {
  const unsigned char* stream;                   // records, one per node
  const uint32_t *rec_off, *rec_bytes;
  const int4* refs;     // per node 2 x int4: (flags, ref_a, ref_b, tips_off), (ref_a2, ref_b2, blk_off, 0)
  uint32_t n_nodes, stage_bytes; // stage_bytes: largest packed stage (record | tip rows | chunks), 128-aligned
  int n_stages;         // ring depth
};
constexpr int kMaxStages = 8;
constexpr int kSG = kChunkSites; // sites per CTA

// what a child is, warp-uniform
enum : int { kInner = 0, kTip = 1, kCherry = 2 };

struct NodePtrs {       // record and stage pointers of one node, lane offsets NOT applied
  const double* tab;                               // first table of the record (shared memory)
  const double *blk_a, *blk_b;                     // partial chunks of inner children
};
// Offsets (in doubles) of a record's tables given what the two children are (schedule.cpp
// build_up_mma_stream): F1a[C] F1b[C] | F3a[C] F3b[C] (non-tips) | leaf tables of cherries | raw
// P, W of tips.  Compile-time constants inside the specialised node bodies: eight pointers less
// to keep in (or spill from) 96 registers.
struct RecLayout { int F1a, F1b, F3a, F3b, Pxa, Pxb, Rta, Rtb; };
__host__ __device__ constexpr RecLayout rec_layout(int C, int ka, int kb) {
  RecLayout r{};
  r.F1a = 0;
  r.F1b = C * 32;
  r.F3a = 2 * C * 32;
  r.F3b = r.F3a + (ka != 1 /* kTip */ ? C * 32 : 0);
  r.Pxa = r.F3b + (kb != 1 ? C * 32 : 0);
  r.Pxb = r.Pxa + (ka == 2 /* kCherry */ ? 2 * C * 16 : 0);
  r.Rta = r.Pxb + (kb == 2 ? 2 * C * 16 : 0);
  r.Rtb = r.Rta + (ka == 1 ? 2 * C * 16 : 0);
  return r;
}

// One node, every leaf state below it resolved (no ambiguity codes in this warp's sites):
// straight-line code over the classes so the C x NG independent DMMA chains interleave.
//   inner child : D from the stage, (S, T) = DMMA(D, [P | W])
//   tip child   : S = P[q][state], T = W[q][state]   (column picks from the raw tables)
//   cherry child: D = P1[q][s1] * P2[q][s2], then as inner
template <int NG, int C, int KA, int KB>
__device__ __forceinline__ void node_fast(const NodePtrs& p, int lane, const int (&sa)[NG], const int (&sa2)[NG],
                                          const int (&sb)[NG], const int (&sb2)[NG], double (&G)[C][NG],
                                          double (*push)[NG], const double (*pop)[NG], double (&acc_a)[NG],
                                          double (&acc_b)[NG]) {
  const int q = lane & 3;
  constexpr RecLayout L = rec_layout(C, KA, KB);
#pragma unroll
  for (int c = 0; c < C; c++) {
    double Sa[NG], Ta[NG], Sb[NG], Tb[NG];
    auto child = [&](auto kind, const double* F1, const double* blk, const double* Px, const double* Rt,
                     const int (&s1)[NG], const int (&s2)[NG], double (&S)[NG], double (&T)[NG]) {
      constexpr int K = decltype(kind)::value;
      if constexpr (K == kTip) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
          S[g] = Rt[c * 16 + q * 4 + s1[g]];
          T[g] = Rt[(C + c) * 16 + q * 4 + s1[g]];
        }
      } else {
        double D[NG];
        if constexpr (K == kInner) {
#pragma unroll
          for (int g = 0; g < NG; g++) D[g] = blk[(size_t)c * (kSG * 4) + g * 32 + lane];
        } else {
#pragma unroll
          for (int g = 0; g < NG; g++) D[g] = Px[c * 16 + q * 4 + s1[g]] * Px[(C + c) * 16 + q * 4 + s2[g]];
        }
        const double f = F1[c * 32 + lane];
#pragma unroll
        for (int g = 0; g < NG; g++) dmma(S[g], T[g], D[g], f);
      }
    };
    child(std::integral_constant<int, KA>(), p.tab + L.F1a, p.blk_a, p.tab + L.Pxa, p.tab + L.Rta, sa, sa2, Sa, Ta);
    child(std::integral_constant<int, KB>(), p.tab + L.F1b, p.blk_b, p.tab + L.Pxb, p.tab + L.Rtb, sb, sb2, Sb, Tb);
    double Ua[NG], Ub[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) {
      Ua[g] = G[c][g] * Sb[g];
      Ub[g] = G[c][g] * Sa[g];
      acc_a[g] = c == 0 ? Ua[g] * Ta[g] : fma(Ua[g], Ta[g], acc_a[g]);
      acc_b[g] = c == 0 ? Ub[g] * Tb[g] : fma(Ub[g], Tb[g], acc_b[g]);
    }
    double unused;
    if constexpr (KA != kTip) {
      if constexpr (KB != kTip) { // both expanded later: b's message waits on the stack
        const double f3 = p.tab[L.F3b + c * 32 + lane];
#pragma unroll
        for (int g = 0; g < NG; g++) dmma(push[c][g], unused, Ub[g], f3);
      }
      const double f3 = p.tab[L.F3a + c * 32 + lane];
#pragma unroll
      for (int g = 0; g < NG; g++) dmma(G[c][g], unused, Ua[g], f3);
    } else if constexpr (KB != kTip) {
      const double f3 = p.tab[L.F3b + c * 32 + lane];
#pragma unroll
      for (int g = 0; g < NG; g++) dmma(G[c][g], unused, Ub[g], f3);
    } else if (pop) {
#pragma unroll
      for (int g = 0; g < NG; g++) G[c][g] = pop[c][g];
    }
  }
}

// Same node with ambiguity codes somewhere in the warp's sites (gaps, N, ...): a tip's partial
// is its 0/1 state mask and goes through the DMMA like an inner child's.
template <int NG, int C>
__device__ __forceinline__ void node_masks(const NodePtrs& p, int lane, int kind_a, int kind_b, const uint32_t (&ma)[NG],
                                        const uint32_t (&ma2)[NG], const uint32_t (&mb)[NG], const uint32_t (&mb2)[NG],
                                        double (&G)[C][NG], double (*push)[NG], const double (*pop)[NG],
                                        double (&acc_a)[NG], double (&acc_b)[NG]) {
  const int q = lane & 3;
  const RecLayout L = rec_layout(C, kind_a, kind_b);
#pragma unroll
  for (int c = 0; c < C; c++) {
    double Da[NG], Db[NG];
    auto child = [&](int kind, const double* blk, const uint32_t (&m1)[NG], const uint32_t (&m2)[NG], const double* Px,
                     double (&D)[NG]) {
      if (kind == kInner) {
#pragma unroll
        for (int g = 0; g < NG; g++) D[g] = blk[(size_t)c * (kSG * 4) + g * 32 + lane];
      } else if (kind == kTip) {
#pragma unroll
        for (int g = 0; g < NG; g++) D[g] = (m1[g] >> q) & 1u ? 1. : 0.;
      } else {
        const double* P1 = Px + c * 16 + q * 4;  // row q of the first leaf's table
        const double* P2 = P1 + C * 16;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          double u = 0., v = 0.;
#pragma unroll
          for (int y = 0; y < 4; y++) {
            u += (m1[g] >> y) & 1u ? P1[y] : 0.;
            v += (m2[g] >> y) & 1u ? P2[y] : 0.;
          }
          D[g] = u * v;
        }
      }
    };
    child(kind_a, p.blk_a, ma, ma2, p.tab + L.Pxa, Da);
    child(kind_b, p.blk_b, mb, mb2, p.tab + L.Pxb, Db);
    const double fa = p.tab[L.F1a + c * 32 + lane], fb = p.tab[L.F1b + c * 32 + lane];
    double Sa[NG], Ta[NG], Sb[NG], Tb[NG], Ua[NG], Ub[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) dmma(Sa[g], Ta[g], Da[g], fa);
#pragma unroll
    for (int g = 0; g < NG; g++) dmma(Sb[g], Tb[g], Db[g], fb);
#pragma unroll
    for (int g = 0; g < NG; g++) {
      Ua[g] = G[c][g] * Sb[g];
      Ub[g] = G[c][g] * Sa[g];
      acc_a[g] = c == 0 ? Ua[g] * Ta[g] : fma(Ua[g], Ta[g], acc_a[g]);
      acc_b[g] = c == 0 ? Ub[g] * Tb[g] : fma(Ub[g], Tb[g], acc_b[g]);
    }
    double unused;
    if (kind_a != kTip) {
      if (kind_b != kTip) {
        const double f3 = p.tab[L.F3b + c * 32 + lane];
#pragma unroll
        for (int g = 0; g < NG; g++) dmma(push[c][g], unused, Ub[g], f3);
      }
      const double f3 = p.tab[L.F3a + c * 32 + lane];
#pragma unroll
      for (int g = 0; g < NG; g++) dmma(G[c][g], unused, Ua[g], f3);
    } else if (kind_b != kTip) {
      const double f3 = p.tab[L.F3b + c * 32 + lane];
#pragma unroll
      for (int g = 0; g < NG; g++) dmma(G[c][g], unused, Ub[g], f3);
    } else if (pop) {
#pragma unroll
      for (int g = 0; g < NG; g++) G[c][g] = pop[c][g];
    }
  }
}

// One node's copies into ring stage s: record, tip rows, partial chunks of the stored children (one lane).
template <int C>
__device__ __forceinline__ void up_issue_node(const MapModel& m, const MapBuffers& b, const UpMmaParams& up, int4 r0, int4 r1,
                                              uint32_t roff, uint32_t rnb, unsigned char* st, uint64_t* full, int64_t site0) {
  constexpr uint32_t kBlock = C * kSG * 32;
  const int64_t n_pad = b.n_pad;
  const int64_t chunk = site0 / kChunkSites;
  const uint32_t flags = (uint32_t)r0.x;
  const int ref_a = r0.y, ref_b = r0.z, ref_a2 = r1.x, ref_b2 = r1.y;
  const uint32_t tips_off = (uint32_t)r0.w, blk_off = (uint32_t)r1.z;
  const bool tipa = flags & kUpTipA, tipb = flags & kUpTipB, cha = flags & kUpCherryA, chb = flags & kUpCherryB;
  const bool ina = !(tipa || cha), inb = !(tipb || chb);
  unsigned char* tp = st + tips_off;
  const uint32_t nrows = (tipa || cha) + (tipb || chb) + cha + chb;
  mbar_expect_tx(full, rnb + nrows * (uint32_t)kSG + ((uint32_t)ina + (uint32_t)inb) * kBlock);
  tma_bulk_g2s(st, up.stream + roff, rnb, full);
  if (tipa || cha) tma_bulk_g2s(tp, b.tips + (size_t)ref_a * n_pad + site0, kSG, full);
  if (tipb || chb) tma_bulk_g2s(tp + kSG, b.tips + (size_t)ref_b * n_pad + site0, kSG, full);
  if (cha) tma_bulk_g2s(tp + 2 * kSG, b.tips + (size_t)ref_a2 * n_pad + site0, kSG, full);
  if (chb) tma_bulk_g2s(tp + 3 * kSG, b.tips + (size_t)ref_b2 * n_pad + site0, kSG, full);
  // stored children, in order a then b, from blk_off
  if (ina) tma_bulk_g2s(st + blk_off, b.D + d_chunk(chunk, ref_a, m.n_slots, C), kBlock, full);
  if (inb) tma_bulk_g2s(st + blk_off + (ina ? kBlock : 0), b.D + d_chunk(chunk, ref_b, m.n_slots, C), kBlock, full);
}

// FOLD: no producer warp.  The CTA is kSG / (8 NG) consumer warps (8 at NG = 2 -> with two CTAs per SM four
// warps per SM sub-partition and 128 registers per thread, where the ninth warp capped them at 96 with spills);
// the refill of the stage node n - 1 used (with node n - 1 + NSTG) is issued by warp (n - 1) % W at the top of
// its iteration n, with the descriptors fetched one iteration earlier.
template <int NG, int C, int MINB, bool FOLD>
__global__ void __launch_bounds__(32 * (kSG / (8 * NG) + (FOLD ? 0 : 1)), MINB) k1_up_mma(MapModel m, MapBuffers b, UpMmaParams up) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int W = kSG / (8 * NG);                // consumer warps
  constexpr uint32_t kBlock = C * kSG * 32;        // bytes of one child's partial chunk
  const uint32_t stage_bytes = up.stage_bytes;     // packed per node: record | tip rows (a, b, a2, b2) | chunks
  const int NSTG = up.n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_pad = b.n_pad;
  const int64_t site0 = (int64_t)blockIdx.x * kSG;
  uint64_t* stg_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* stg_empty = stg_full + kMaxStages;
  unsigned char* stg_ring = smem + 128;

  __shared__ uint32_t cmask[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cmask[i] = __ldg(m.code_mask + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTG; i++) { mbar_init(&stg_full[i], 1); mbar_init(&stg_empty[i], W); }
    mbar_fence_init();
  }
  __syncthreads();

  if (!FOLD && warp == W) {
    // ---- producer: node n's record, tip rows and the partial chunks of its inner children go
    //      to stage n % NSTG; one mbarrier full / empty pair per stage.  The 32 lanes fetch the
    //      descriptors of 32 nodes at a time; lane 0 issues the copies.
    uint32_t s = 0, ph = 1;
    bool first = true;
    for (uint32_t n0 = 0; n0 < up.n_nodes; n0 += 32) {
      const uint32_t mine = min(n0 + lane, up.n_nodes - 1);
      const int4 r0 = __ldg(up.refs + 2 * mine), r1 = __ldg(up.refs + 2 * mine + 1);
      const uint32_t off = __ldg(up.rec_off + mine), nb = __ldg(up.rec_bytes + mine);
      const uint32_t cnt = min(32u, up.n_nodes - n0);
      for (uint32_t j = 0; j < cnt; j++) {
        int4 q0, q1;
        q0.x = __shfl_sync(0xffffffffu, r0.x, j); q0.y = __shfl_sync(0xffffffffu, r0.y, j);
        q0.z = __shfl_sync(0xffffffffu, r0.z, j); q0.w = __shfl_sync(0xffffffffu, r0.w, j);
        q1.x = __shfl_sync(0xffffffffu, r1.x, j); q1.y = __shfl_sync(0xffffffffu, r1.y, j);
        q1.z = __shfl_sync(0xffffffffu, r1.z, j); q1.w = 0;
        const uint32_t roff = __shfl_sync(0xffffffffu, off, j), rnb = __shfl_sync(0xffffffffu, nb, j);
        if (!first) mbar_wait_sleep(&stg_empty[s], ph, 200);
        if (lane == 0) up_issue_node<C>(m, b, up, q0, q1, roff, rnb, stg_ring + (size_t)s * stage_bytes, &stg_full[s], site0);
        __syncwarp();
        if (++s == (uint32_t)NSTG) { s = 0; ph ^= 1; first = false; }
      }
    }
    return;
  }
  if (FOLD) { // prologue: the first NSTG nodes, one per warp
    for (uint32_t n = warp; n < (uint32_t)NSTG && n < up.n_nodes; n += W)
      if (lane == 0)
        up_issue_node<C>(m, b, up, __ldg(up.refs + 2 * n), __ldg(up.refs + 2 * n + 1), __ldg(up.rec_off + n), __ldg(up.rec_bytes + n),
                         stg_ring + (size_t)n * stage_bytes, &stg_full[n], site0);
  }

  // ---- consumers
  const int q = lane & 3, s8 = lane >> 2;
  const int wsite = warp * (8 * NG);             // first site of this warp inside the CTA
  double G[C][NG];
  double stk[kMaxStack][C][NG];
  int sp = 0;
  {
    const double piq = __ldg(m.pi + q);
#pragma unroll
    for (int c = 0; c < C; c++)
#pragma unroll
      for (int g = 0; g < NG; g++) G[c][g] = piq;
  }
  // after the quad reduction lane q owns the items (branch = q >> 1, groups j * 2 + (q & 1))
  constexpr int NJ = NG / 2;
  double invL[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) invL[j] = b.invL[site0 + wsite + 8 * (2 * j + (q & 1)) + s8];

  uint32_t cs = 0, cph = 0;
  int4 pre0 = make_int4(0, 0, 0, 0), pre1 = pre0; // FOLD: descriptors of the node this warp issues next
  uint32_t pre_off = 0, pre_nb = 0;
  for (uint32_t node = 0; node < up.n_nodes; node++) {
    if (FOLD) {
      // refill duty of node - 1's stage (all W warps have arrived on its empty barrier before it is reused)
      if (node >= 1 && warp == (int)((node - 1) % W) && node - 1 + NSTG < up.n_nodes) {
        const uint32_t ps = cs == 0 ? (uint32_t)NSTG - 1 : cs - 1, pph = cs == 0 ? cph ^ 1 : cph;
        if (lane == 0) {
          mbar_wait(&stg_empty[ps], pph);
          up_issue_node<C>(m, b, up, pre0, pre1, pre_off, pre_nb, stg_ring + (size_t)ps * stage_bytes, &stg_full[ps], site0);
        }
        __syncwarp();
      }
      if (warp == (int)(node % W) && node + NSTG < up.n_nodes) { // used at the top of the next iteration
        const uint32_t nn = node + NSTG;
        pre0 = __ldg(up.refs + 2 * nn); pre1 = __ldg(up.refs + 2 * nn + 1);
        pre_off = __ldg(up.rec_off + nn); pre_nb = __ldg(up.rec_bytes + nn);
      }
    }
    mbar_wait(&stg_full[cs], cph);
    const unsigned char* stage = stg_ring + (size_t)cs * stage_bytes;
    const int4 h0 = *reinterpret_cast<const int4*>(stage);
    const int4 h1 = *reinterpret_cast<const int4*>(stage + 16);
    const uint32_t flags = (uint32_t)h0.x;
    const int kind_a = (flags & kUpTipA) ? kTip : (flags & kUpCherryA) ? kCherry : kInner;
    const int kind_b = (flags & kUpTipB) ? kTip : (flags & kUpCherryB) ? kCherry : kInner;
    NodePtrs p;
    p.tab = reinterpret_cast<const double*>(stage + 32);
    // h1.y / h1.z: offsets of the tip rows and of the first stored child's chunk in this stage
    const unsigned char* ts = stage + h1.y + wsite + s8; // tip code of (row, group g): ts[row * kSG + 8 g]
    p.blk_a = reinterpret_cast<const double*>(stage + h1.z) + (size_t)wsite * 4;
    p.blk_b = p.blk_a + (kind_a == kInner ? kBlock / 8 : 0);

    // ---- tips and cherries: states (or state masks) of this lane's NG sites
    double acc_a[NG], acc_b[NG];
    double (*push)[NG] = stk[sp];
    const double (*pop)[NG] = (flags & kUpPop) ? stk[sp > 0 ? sp - 1 : 0] : nullptr;
    int sa[NG], sa2[NG], sb[NG], sb2[NG];
    uint32_t ma[NG], ma2[NG], mb[NG], mb2[NG];
    bool fast = true;
    if (m.states_only) { // device-simulated alignment: the codes are the states
#pragma unroll
      for (int g = 0; g < NG; g++) {
        sa[g] = kind_a != kInner ? ts[8 * g] : 0;
        sa2[g] = kind_a == kCherry ? ts[2 * kSG + 8 * g] : 0;
        sb[g] = kind_b != kInner ? ts[kSG + 8 * g] : 0;
        sb2[g] = kind_b == kCherry ? ts[3 * kSG + 8 * g] : 0;
      }
    } else {
      bool single = true;
#pragma unroll
      for (int g = 0; g < NG; g++) {
        ma[g] = kind_a != kInner ? cmask[ts[8 * g]] : 1u;
        ma2[g] = kind_a == kCherry ? cmask[ts[2 * kSG + 8 * g]] : 1u;
        mb[g] = kind_b != kInner ? cmask[ts[kSG + 8 * g]] : 1u;
        mb2[g] = kind_b == kCherry ? cmask[ts[3 * kSG + 8 * g]] : 1u;
        single = single && __popc(ma[g]) == 1 && __popc(ma2[g]) == 1 && __popc(mb[g]) == 1 && __popc(mb2[g]) == 1;
        sa[g] = __ffs(ma[g]) - 1; sa2[g] = __ffs(ma2[g]) - 1;
        sb[g] = __ffs(mb[g]) - 1; sb2[g] = __ffs(mb2[g]) - 1;
      }
      fast = __all_sync(0xffffffffu, single);
    }
    if (fast) {
      switch (kind_a * 3 + kind_b) {
#define CMB_NODE(KA, KB) \
  case KA * 3 + KB: node_fast<NG, C, KA, KB>(p, lane, sa, sa2, sb, sb2, G, push, pop, acc_a, acc_b); break;
        CMB_NODE(kInner, kInner) CMB_NODE(kInner, kTip) CMB_NODE(kInner, kCherry)
        CMB_NODE(kTip, kInner) CMB_NODE(kTip, kTip) CMB_NODE(kTip, kCherry)
        CMB_NODE(kCherry, kInner) CMB_NODE(kCherry, kTip) CMB_NODE(kCherry, kCherry)
#undef CMB_NODE
      }
    } else {
      node_masks<NG, C>(p, lane, kind_a, kind_b, ma, ma2, mb, mb2, G, push, pop, acc_a, acc_b);
    }
    if (flags & kUpPush) ++sp;
    else if (flags & kUpPop) --sp;
    // the output rows are read again here rather than kept live (or spilled) across the node body
    const int ob = reinterpret_cast<const volatile int*>(stage)[3 + ((lane >> 1) & 1)]; // q & 2 ? out_b : out_a
    __syncwarp();
    if (lane == 0) mbar_arrive(&stg_empty[cs]);   // the stage has been read
    if (++cs == (uint32_t)NSTG) { cs = 0; cph ^= 1; }

    // ---- sum over the four state lanes of a site: 2 NG values per lane -> NG / 2 complete
    //      sums per lane (transposing reduction, 3 NG / 2 shuffles instead of 4 NG)
    double v[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) {
      const double send = (q & 2) ? acc_a[g] : acc_b[g];
      const double keep = (q & 2) ? acc_b[g] : acc_a[g];
      v[g] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      const double send = (q & 1) ? v[2 * j] : v[2 * j + 1];
      const double keep = (q & 1) ? v[2 * j + 1] : v[2 * j];
      const double t = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      if (ob >= 0) b.out[(size_t)ob * n_pad + site0 + wsite + 8 * (2 * j + (q & 1)) + s8] = t * invL[j];
    }
  }
}

template <int NG, int C, int MINB, bool FOLD>
bool try_up_mma(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.C != C) return false;
  int dev = 0, max_smem = 0;
  CMB_CUDA(cudaGetDevice(&dev));
  CMB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  max_smem = max_smem / MINB - 1024 - 1024;      // 1 KB system reserve per CTA, 1 KB static (cmask)
  const size_t stage = ((size_t)s.stage_bytes + 127) & ~size_t(127);
  const size_t fixed = 128;
  if ((size_t)max_smem < fixed + 2 * stage) return false;
  static const int stage_cap = getenv("CMB_UP_STAGES") ? atoi(getenv("CMB_UP_STAGES")) : kMaxStages;
  UpMmaParams up;
  up.stream = s.bytes.as<unsigned char>();
  up.rec_off = s.off.as<uint32_t>();
  up.rec_bytes = s.nbytes.as<uint32_t>();
  up.refs = s.aux.as<int4>();
  up.n_nodes = s.n_records;
  up.stage_bytes = (uint32_t)stage;
  up.n_stages = (int)std::min<size_t>(std::min(kMaxStages, stage_cap), ((size_t)max_smem - fixed) / stage);
  // Shared memory is carved out of the 256 KB it shares with L1, and the consumers' message stack and
  // register spills live in local memory behind that L1: a third stage at C = 4 (2 x 111 KB of shared
  // memory, ~28 KB of L1) ran 12.4 ms instead of 10.8 ms.  Keep a CTA's ring within 80 KB.
  static const size_t ring_kb = getenv("CMB_UP_RING_KB") ? (size_t)atoi(getenv("CMB_UP_RING_KB")) : 80;
  if (MINB > 1) up.n_stages = (int)std::max<size_t>(2, std::min<size_t>(up.n_stages, (ring_kb * 1024) / stage));
  const size_t smem = fixed + (size_t)up.n_stages * stage;
  constexpr int threads = 32 * (kSG / (8 * NG) + (FOLD ? 0 : 1));
  CMB_CUDA(cudaFuncSetAttribute(k1_up_mma<NG, C, MINB, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (getenv("CMB_UP_CARVEOUT"))
    CMB_CUDA(cudaFuncSetAttribute(k1_up_mma<NG, C, MINB, FOLD>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("CMB_UP_CARVEOUT"))));
  k1_up_mma<NG, C, MINB, FOLD><<<(unsigned)(b.n_pad / kSG), threads, smem, st>>>(m, b, up);
  CMB_CUDA(cudaGetLastError());
  return true;
}

template <int C>
bool up_mma_for(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.C != C) return false;
  // B200, config 4 (517 k sites): 2 groups/warp x 2 CTAs/SM (16 consumer warps) 10.8 ms;
  // 4 groups/warp x 2 CTAs/SM (8 warps, 168 regs) 11.2 ms; 2 groups/warp x 1 CTA/SM 12.7 ms
  if constexpr (C == 4) {
    static const int shape = getenv("CMB_UP_SHAPE") ? atoi(getenv("CMB_UP_SHAPE")) : 0; // experiment switch
    if (shape == 42) return try_up_mma<4, C, 2, false>(m, b, s, st);
    if (shape == 21) return try_up_mma<2, C, 1, false>(m, b, s, st);
  }
  static const int fold = getenv("CMB_UP_FOLD") ? atoi(getenv("CMB_UP_FOLD")) : 1; // experiment switch
  if (fold) return try_up_mma<2, C, 2, true>(m, b, s, st) || try_up_mma<2, C, 1, true>(m, b, s, st);
  return try_up_mma<2, C, 2, false>(m, b, s, st) || try_up_mma<2, C, 1, false>(m, b, s, st);
}

// ------------------------------------------------------------------------------ down
// Felsenstein post-order pass in the same lane mapping (lane = (site, state), warp = 8 NG
// sites x all classes).  A child's message through its edge, P D, is one DMMA per 8 sites
// (the second half of the B operand is unused) or, for a resolved tip, a column pick; the
// message of a larger sibling waits in a SHARED-MEMORY stack (depth <= log2 T, known from the
// schedule) instead of per-thread local memory; every stored partial leaves as one 256-byte
// row per (class, 8 sites).  The only global reads are the op stream and the tip rows, both
// through the producer warp's TMA ring.
struct DownMmaParams {
  const unsigned char* stream;
  const uint32_t *rec_off, *rec_bytes;
  const int4* refs;     // per node (flags, row_a, row_b, 0)
  uint32_t n_nodes, rec_cap;
  int n_stages, smem_levels;
};
constexpr int kDownStages = 8;

template <int NG, int C, int MINB>
__global__ void __launch_bounds__(32 * (kSG / (8 * NG) + 1), MINB) k1_down_mma(MapModel m, MapBuffers b, DownMmaParams dp) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int W = kSG / (8 * NG);
  constexpr uint32_t kEntry = C * kSG * 32;        // one stack level of the CTA
  const uint32_t stage_bytes = dp.rec_cap + 2 * kSG;
  const int NSTG = dp.n_stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_pad = b.n_pad;
  const int64_t site0 = (int64_t)blockIdx.x * kSG;
  uint64_t* stg_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* stg_empty = stg_full + kDownStages;
  unsigned char* stg_ring = smem + 128;
  double* stack = reinterpret_cast<double*>(stg_ring + (size_t)NSTG * stage_bytes);

  __shared__ uint32_t cmask[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cmask[i] = __ldg(m.code_mask + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTG; i++) { mbar_init(&stg_full[i], 1); mbar_init(&stg_empty[i], W); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == W) {
    // ---- producer: record + the tip rows of node n into stage n % NSTG
    uint32_t s = 0, ph = 1;
    bool first = true;
    for (uint32_t n0 = 0; n0 < dp.n_nodes; n0 += 32) {
      const uint32_t mine = min(n0 + lane, dp.n_nodes - 1);
      const int4 r0 = __ldg(dp.refs + mine);
      const uint32_t off = __ldg(dp.rec_off + mine), nb = __ldg(dp.rec_bytes + mine);
      const uint32_t cnt = min(32u, dp.n_nodes - n0);
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t flags = (uint32_t)__shfl_sync(0xffffffffu, r0.x, j);
        const int row_a = __shfl_sync(0xffffffffu, r0.y, j), row_b = __shfl_sync(0xffffffffu, r0.z, j);
        const uint32_t roff = __shfl_sync(0xffffffffu, off, j), rnb = __shfl_sync(0xffffffffu, nb, j);
        if (!first) mbar_wait_sleep(&stg_empty[s], ph, 200);
        if (lane == 0) {
          const bool tipa = flags & kDownTipA, tipb = flags & kDownTipB;
          unsigned char* st = stg_ring + (size_t)s * stage_bytes;
          mbar_expect_tx(&stg_full[s], rnb + ((uint32_t)tipa + (uint32_t)tipb) * (uint32_t)kSG);
          tma_bulk_g2s(st, dp.stream + roff, rnb, &stg_full[s]);
          if (tipa) tma_bulk_g2s(st + dp.rec_cap, b.tips + (size_t)row_a * n_pad + site0, kSG, &stg_full[s]);
          if (tipb) tma_bulk_g2s(st + dp.rec_cap + kSG, b.tips + (size_t)row_b * n_pad + site0, kSG, &stg_full[s]);
        }
        __syncwarp();
        if (++s == (uint32_t)NSTG) { s = 0; ph ^= 1; first = false; }
      }
    }
    return;
  }

  // ---- consumers
  const int q = lane & 3, s8 = lane >> 2;
  const int wsite = warp * (8 * NG);
  double cur[C][NG];
  double stk_l[kMaxStack][C][NG];                  // levels beyond the shared-memory stack
  int sp = 0;
#pragma unroll
  for (int c = 0; c < C; c++)
#pragma unroll
    for (int g = 0; g < NG; g++) cur[c][g] = 0.;
  double* my_stack = stack + (size_t)wsite * 4 + lane; // + level * kEntry / 8 + c * kSG * 4 + g * 32
  double* my_D = b.D + d_chunk(site0 / kChunkSites, 0, m.n_slots, C) + (size_t)wsite * 4 + lane;
  const size_t slot_stride = (size_t)C * kChunkSites * 4;

  uint32_t cs = 0, cph = 0;
  for (uint32_t node = 0; node < dp.n_nodes; node++) {
    mbar_wait(&stg_full[cs], cph);
    const unsigned char* stage = stg_ring + (size_t)cs * stage_bytes;
    const int4 h = *reinterpret_cast<const int4*>(stage);
    const uint32_t flags = (uint32_t)h.x;
    const bool tipa = flags & kDownTipA, tipb = flags & kDownTipB;
    const double* F = reinterpret_cast<const double*>(stage + 16);              // fragment of the running child
    const double* Ra = F + (tipa ? 0 : C * 32);                                  // raw P of tip a
    const double* Rb = Ra + (tipa ? C * 16 : 0);                                 // raw P of tip b
    const double* Fv = Rb + (tipb ? C * 16 : 0);                                 // fragment of v's own edge (push)
    const unsigned char* ts = stage + dp.rec_cap + wsite + s8;

    // message of a tip through its edge: lane q's component, all classes
    auto tip_message = [&](const unsigned char* codes, const double* R, double (&M)[C][NG]) {
      bool fast = true;
      uint32_t mk[NG];
      if (m.states_only) {
#pragma unroll
        for (int g = 0; g < NG; g++) mk[g] = codes[8 * g];
      } else {
        bool single = true;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          mk[g] = cmask[codes[8 * g]];
          single = single && __popc(mk[g]) == 1;
        }
        fast = __all_sync(0xffffffffu, single);
        if (fast) {
#pragma unroll
          for (int g = 0; g < NG; g++) mk[g] = __ffs(mk[g]) - 1;
        }
      }
      if (fast) {
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) M[c][g] = R[c * 16 + q * 4 + mk[g]];
      } else {
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) {
            double t = 0.;
#pragma unroll
            for (int y = 0; y < 4; y++) t += (mk[g] >> y) & 1u ? R[c * 16 + q * 4 + y] : 0.;
            M[c][g] = t;
          }
      }
    };
    auto edge_message = [&](const double* Fr, double (&M)[C][NG]) { // M = P cur, one DMMA per (class, 8 sites)
      double unused;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const double f = Fr[c * 32 + lane];
#pragma unroll
        for (int g = 0; g < NG; g++) dmma(M[c][g], unused, cur[c][g], f);
      }
    };

    double Ma[C][NG], Mb[C][NG];
    if (tipa) {                       // cherry
      tip_message(ts, Ra, Ma);
      tip_message(ts + kSG, Rb, Mb);
    } else if (tipb) {                // a's partial is the running one
      edge_message(F, Ma);
      tip_message(ts + kSG, Rb, Mb);
    } else {                          // a's message waits on the stack, b's partial is the running one
      edge_message(F, Mb);
      --sp;
      if (sp < dp.smem_levels) {
        const double* e = my_stack + (size_t)sp * (kEntry / 8);
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) Ma[c][g] = e[c * (kSG * 4) + g * 32];
      } else {
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) Ma[c][g] = stk_l[sp - dp.smem_levels][c][g];
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++)
#pragma unroll
      for (int g = 0; g < NG; g++) cur[c][g] = Ma[c][g] * Mb[c][g];
    if (h.w >= 0) {
      double* d = my_D + (size_t)h.w * slot_stride;
#pragma unroll
      for (int c = 0; c < C; c++)
#pragma unroll
        for (int g = 0; g < NG; g++) d[c * (kChunkSites * 4) + g * 32] = cur[c][g];
    }
    if (flags & kDownPush) {
      double M[C][NG];
      edge_message(Fv, M);
      if (sp < dp.smem_levels) {
        double* e = my_stack + (size_t)sp * (kEntry / 8);
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) e[c * (kSG * 4) + g * 32] = M[c][g];
      } else {
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
          for (int g = 0; g < NG; g++) stk_l[sp - dp.smem_levels][c][g] = M[c][g];
      }
      ++sp;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&stg_empty[cs]);
    if (++cs == (uint32_t)NSTG) { cs = 0; cph ^= 1; }
  }
  // ---- root: class likelihood L_c = sum_x pi_x root[c][x]
  const double piq = __ldg(m.pi + q);
#pragma unroll
  for (int c = 0; c < C; c++)
#pragma unroll
    for (int g = 0; g < NG; g++) {
      double v = cur[c][g] * piq;
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (q == 0) b.Lc[(size_t)c * n_pad + site0 + wsite + 8 * g + s8] = v;
    }
}

template <int NG, int C, int MINB>
bool try_down_mma(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.C != C) return false;
  int dev = 0, max_smem = 0;
  CMB_CUDA(cudaGetDevice(&dev));
  CMB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  max_smem = max_smem / MINB - 1024 - 1024;      // 1 KB system reserve per CTA, 1 KB static (cmask)
  const size_t stage = (size_t)s.cap + 2 * (size_t)kSG, entry = (size_t)C * kSG * 32;
  DownMmaParams dp;
  dp.stream = s.bytes.as<unsigned char>();
  dp.rec_off = s.off.as<uint32_t>();
  dp.rec_bytes = s.nbytes.as<uint32_t>();
  dp.refs = s.aux.as<int4>();
  dp.n_nodes = s.n_records;
  dp.rec_cap = s.cap;
  dp.n_stages = kDownStages;
  while (dp.n_stages > 2 && 128 + dp.n_stages * stage + entry > (size_t)max_smem) dp.n_stages /= 2;
  if (128 + dp.n_stages * stage > (size_t)max_smem) return false;
  // as many stack levels in shared memory as fit; deeper ones (rare) spill to local memory
  dp.smem_levels = (int)std::min<size_t>((size_t)s.stack_depth, ((size_t)max_smem - 128 - dp.n_stages * stage) / entry);
  if (MINB > 1 && dp.smem_levels < s.stack_depth) return false; // prefer one CTA per SM with the whole stack
  const size_t smem = 128 + dp.n_stages * stage + (size_t)dp.smem_levels * entry;
  constexpr int threads = 32 * (kSG / (8 * NG) + 1);
  CMB_CUDA(cudaFuncSetAttribute(k1_down_mma<NG, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k1_down_mma<NG, C, MINB><<<(unsigned)(b.n_pad / kSG), threads, smem, st>>>(m, b, dp);
  CMB_CUDA(cudaGetLastError());
  return true;
}

template <int C>
bool down_mma_for(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.C != C) return false;
  if constexpr (C == 4) {
    static const int shape = getenv("CMB_DOWN_SHAPE") ? atoi(getenv("CMB_DOWN_SHAPE")) : 0; // experiment switch
    if (shape == 21) return try_down_mma<2, C, 1>(m, b, s, st);
    if (shape == 11) return try_down_mma<1, C, 1>(m, b, s, st);
    if (shape == 12) return try_down_mma<1, C, 2>(m, b, s, st);
  }
  return try_down_mma<2, C, 2>(m, b, s, st) || try_down_mma<2, C, 1>(m, b, s, st);
}

} // namespace

void launch_map_up_mma(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.A != 4) fail("internal: the tensor-core up pass is built for A = 4");
  if (b.n_pad % kSG) fail("internal: n_pad must be a multiple of %d", kSG);
  const bool done = up_mma_for<1>(m, b, s, st) || up_mma_for<2>(m, b, s, st) || up_mma_for<3>(m, b, s, st) ||
                    up_mma_for<4>(m, b, s, st) || up_mma_for<5>(m, b, s, st) || up_mma_for<6>(m, b, s, st) ||
                    up_mma_for<7>(m, b, s, st) || up_mma_for<8>(m, b, s, st);
  if (!done) fail("mapping up pass: no launch shape fits shared memory for A = 4, C = %d", m.C);
}

} // namespace cmb

namespace cmb {
void launch_map_down_mma(const MapModel& m, const MapBuffers& b, const DevStream& s, cudaStream_t st) {
  if (m.A != 4) fail("internal: the tensor-core down pass is built for A = 4");
  if (b.n_pad % kSG) fail("internal: n_pad must be a multiple of %d", kSG);
  const bool done = down_mma_for<1>(m, b, s, st) || down_mma_for<2>(m, b, s, st) || down_mma_for<3>(m, b, s, st) ||
                    down_mma_for<4>(m, b, s, st) || down_mma_for<5>(m, b, s, st) || down_mma_for<6>(m, b, s, st) ||
                    down_mma_for<7>(m, b, s, st) || down_mma_for<8>(m, b, s, st);
  if (!done) fail("mapping down pass: no launch shape fits shared memory for A = 4, C = %d", m.C);
}
} // namespace cmb
