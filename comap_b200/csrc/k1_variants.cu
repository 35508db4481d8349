// K1 variants: nijt.average = no and / or nijt.joint = no (CoETools.cpp:393-407, AnalysisTools.cpp:597-633).
//
// The reference marks these two options "really for benchmarking only": instead of
// LegacySubstitutionMappingTools::computeSubstitutionVectors (the tensor-core path of k1_mma*.cu) the vectors
// come from ...Marginal (product of the two ends' marginal posteriors), ...NoAveraging (the most probable joint
// pair of ancestral states) or ...NoAveragingMarginal (the marginal reconstruction of every node).  The last two
// are what `statistic=MI` with `nijt=Label` needs (CoETools.cpp:577-589).  [Bio++ / from memory: none of the
// three bodies is in the reference tree and no golden exists -- parity unpinned; the formulas are stated in
// DESIGN.md s8 and in include/comap_b200.h (cmb_set_map_mode); sums run in a fixed order without fusing.]
//
// Layout: one thread per site walks the ORIGINAL (k-ary) tree: post-order for the partials below every node,
// pre-order for the partials above, both kept for all nodes in a site-minor scratch [node][class][state][site]
// (coalesced across the warp), chunked over sites so the scratch stays below a fixed budget.  Tables P and the
// count matrices N are read through the read-only path (every thread of a warp reads the same entry).  No fused
// multiply-adds: decisions (arg-max of a pair / of a state) then agree with a CPU evaluating the same order.
#include "kernels.h"
#include "device_utils.cuh"
#include <algorithm>

namespace cmb {
namespace {

struct VarArgs {
  int n_nodes, C, mode;       // mode: 1 marginal, 2 no averaging (joint pair), 3 no averaging, marginal states, 4 states only
  uint8_t* anc_out;           // mode 4: [n_nodes][n_pad] marginal state of every node (leaves: first compatible state)
  const int32_t* parent;      // [n_nodes]
  const int32_t* ch_off;      // [n_nodes + 1]
  const int32_t* ch;          // children lists in id order
  const int32_t* leaf_row;    // node -> alignment row or -1
  const double* P;            // [branch][C][A][A]
  const double* N;            // [branch][C][A][A]
  const double* pi;
  const double* probs;
  const uint32_t* code_mask;
  const uint8_t* tips;        // [T][n_pad]
  int64_t n_pad, site0, n;    // this launch covers sites [site0, site0 + n)
  int64_t chunk;              // scratch row length (sites)
  double* down;               // [n_nodes][C][A][chunk]
  double* up;                 // [n_nodes][C][A][chunk]
  uint8_t* anc;               // [n_nodes][chunk] (mode 3)
  double* out;                // [B][n_pad]
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// sum_y P[x][y] v[y]
template <int A>
__device__ __forceinline__ double row_dot(const double* __restrict__ P, int x, const double* v) {
  double l = 0.;
#pragma unroll 4
  for (int y = 0; y < A; y++) l = add(l, mul(__ldg(P + x * A + y), v[y]));
  return l;
}

// conditional likelihood of all the data given state x at inner / root node v and class c: sons in order, then
// the father's side (or the root frequencies) -- DRHomogeneousTreeLikelihood::computeLikelihoodAtNode
template <int A>
__device__ double node_lik(const VarArgs& a, int v, int c, int x, int64_t t, const double* uv /* [A] above v, class c */) {
  const size_t row = (size_t)a.chunk;
  double l = 1.;
  for (int k = a.ch_off[v]; k < a.ch_off[v + 1]; k++) {
    const int w = a.ch[k];
    const double* P = a.P + ((size_t)w * a.C + c) * (A * A);
    const double* dw = a.down + ((size_t)w * a.C + c) * A * row + t;
    double m = 0.;
#pragma unroll 4
    for (int y = 0; y < A; y++) m = add(m, mul(__ldg(P + x * A + y), dw[(size_t)y * row]));
    l = mul(l, m);
  }
  if (v == a.n_nodes - 1) return mul(l, __ldg(a.pi + x));
  const double* P = a.P + ((size_t)v * a.C + c) * (A * A);
  double m = 0.;
#pragma unroll 4
  for (int z = 0; z < A; z++) m = add(m, mul(uv[z], __ldg(P + z * A + x)));
  return mul(l, m);
}

template <int A>
__global__ void __launch_bounds__(128) k1_variant(VarArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n) return;
  const int64_t site = a.site0 + t;
  const int C = a.C, n = a.n_nodes, root = n - 1;
  const size_t row = (size_t)a.chunk;
  constexpr int AA = A * A;
  double vec[A], acc[A];

  // ---- postfix pass: down[v][c][x] = product over the sons of (P_son down[son])[x]; tips = 0/1 masks
  for (int v = 0; v < n; v++) {
    if (a.ch_off[v] == a.ch_off[v + 1]) {
      const uint32_t m = __ldg(a.code_mask + a.tips[(size_t)a.leaf_row[v] * a.n_pad + site]);
      for (int c = 0; c < C; c++)
        for (int x = 0; x < A; x++) a.down[((size_t)(v * C + c) * A + x) * row + t] = (m >> x) & 1u ? 1. : 0.;
      continue;
    }
    for (int c = 0; c < C; c++) {
#pragma unroll
      for (int x = 0; x < A; x++) acc[x] = 1.;
      for (int k = a.ch_off[v]; k < a.ch_off[v + 1]; k++) {
        const int w = a.ch[k];
        const double* dw = a.down + ((size_t)(w * C + c) * A) * row + t;
#pragma unroll
        for (int y = 0; y < A; y++) vec[y] = dw[(size_t)y * row];
        const double* P = a.P + ((size_t)w * C + c) * AA;
#pragma unroll 1
        for (int x = 0; x < A; x++) acc[x] = mul(acc[x], row_dot<A>(P, x, vec));
      }
#pragma unroll
      for (int x = 0; x < A; x++) a.down[((size_t)(v * C + c) * A + x) * row + t] = acc[x];
    }
  }
  // ---- site likelihood (the reference's own sum: per class over the root states, then over the classes)
  double L = 0.;
  for (int c = 0; c < C; c++) {
    double l = 0.;
    for (int x = 0; x < A; x++) l = add(l, mul(a.down[((size_t)(root * C + c) * A + x) * row + t], __ldg(a.pi + x)));
    L = add(L, mul(l, __ldg(a.probs + c)));
  }
  // ---- marginal state of the root (mode 3)
  if (a.mode >= 3) {
    double bv = -INFINITY; int best = 0;
    for (int x = 0; x < A; x++) {
      double l = 0.;
      for (int c = 0; c < C; c++) l = add(l, mul(node_lik<A>(a, root, c, x, t, nullptr), __ldg(a.probs + c)) / L);
      if (l > bv) { bv = l; best = x; }
    }
    a.anc[(size_t)root * row + t] = (uint8_t)best;
    if (a.mode == 4) a.anc_out[(size_t)root * a.n_pad + site] = (uint8_t)best;
  }
  // ---- prefix pass + the branch above every node (parents carry larger ids: walk the ids downwards)
  for (int v = n - 2; v >= 0; v--) {
    const int f = a.parent[v];
    const bool leaf = a.ch_off[v] == a.ch_off[v + 1];
    // up[v][c][x]: everything but v's subtree, state x at the father
    for (int c = 0; c < C; c++) {
#pragma unroll
      for (int x = 0; x < A; x++) acc[x] = 1.;
      for (int k = a.ch_off[f]; k < a.ch_off[f + 1]; k++) {
        const int w = a.ch[k];
        if (w == v) continue;
        const double* dw = a.down + ((size_t)(w * C + c) * A) * row + t;
#pragma unroll
        for (int y = 0; y < A; y++) vec[y] = dw[(size_t)y * row];
        const double* P = a.P + ((size_t)w * C + c) * AA;
#pragma unroll 1
        for (int x = 0; x < A; x++) acc[x] = mul(acc[x], row_dot<A>(P, x, vec));
      }
      if (f != root) {
        const double* uf = a.up + ((size_t)(f * C + c) * A) * row + t;
#pragma unroll
        for (int y = 0; y < A; y++) vec[y] = uf[(size_t)y * row];
        const double* P = a.P + ((size_t)f * C + c) * AA;
#pragma unroll 1
        for (int x = 0; x < A; x++) {
          double l = 0.;
#pragma unroll 4
          for (int y = 0; y < A; y++) l = add(l, mul(__ldg(P + y * A + x), vec[y]));
          acc[x] = mul(acc[x], l);
        }
      } else {
#pragma unroll
        for (int x = 0; x < A; x++) acc[x] = mul(acc[x], __ldg(a.pi + x));
      }
#pragma unroll
      for (int x = 0; x < A; x++) a.up[((size_t)(v * C + c) * A + x) * row + t] = acc[x];
    }
    const double* Pv = a.P + (size_t)v * C * AA;
    const double* Nv = a.N + (size_t)v * C * AA;
    const double* dn = a.down + (size_t)v * C * A * row + t; // [c][y] at (c * A + y) * row
    const double* uu = a.up + (size_t)v * C * A * row + t;
    double res = 0.;
    if (a.mode == 2) {
      // the joint pair (x at the father, y at the node) of largest probability summed over the classes
      double best = -INFINITY; int bx = 0, by = 0;
      for (int x = 0; x < A; x++)
        for (int y = 0; y < A; y++) {
          double pr = 0.;
          for (int c = 0; c < C; c++)
            pr = add(pr, mul(mul(mul(__ldg(a.probs + c), uu[(size_t)(c * A + x) * row]), __ldg(Pv + c * AA + x * A + y)), dn[(size_t)(c * A + y) * row]));
          if (pr > best) { best = pr; bx = x; by = y; }
        }
      double sc = 0.;
      for (int c = 0; c < C; c++)
        sc = add(sc, mul(mul(mul(mul(__ldg(a.probs + c), uu[(size_t)(c * A + bx) * row]), __ldg(Pv + c * AA + bx * A + by)), dn[(size_t)(c * A + by) * row]),
                         __ldg(Nv + c * AA + bx * A + by)));
      res = sc / best;
    } else if (a.mode >= 3) {
      int y = 0;
      if (leaf) { // whichMax of the leaf's 0/1 array: its first compatible state
        const uint32_t m = __ldg(a.code_mask + a.tips[(size_t)a.leaf_row[v] * a.n_pad + site]) & (A >= 32 ? 0xffffffffu : (1u << A) - 1u);
        y = m ? __ffs(m) - 1 : 0;
      } else {
        double bv = -INFINITY;
        for (int s = 0; s < A; s++) {
          double l = 0.;
          for (int c = 0; c < C; c++) {
#pragma unroll
            for (int z = 0; z < A; z++) vec[z] = uu[(size_t)(c * A + z) * row];
            l = add(l, mul(node_lik<A>(a, v, c, s, t, vec), __ldg(a.probs + c)) / L);
          }
          if (l > bv) { bv = l; y = s; }
        }
      }
      a.anc[(size_t)v * row + t] = (uint8_t)y;
      if (a.mode == 4) { a.anc_out[(size_t)v * a.n_pad + site] = (uint8_t)y; continue; }
      const int x = a.anc[(size_t)f * row + t];
      for (int c = 0; c < C; c++) res = add(res, mul(__ldg(Nv + c * AA + x * A + y), __ldg(a.probs + c)));
    } else {
      // marginal posteriors per (class, state) of the father and of the node, multiplied
      const double* uf = a.up + (size_t)f * C * A * row + t;
      double Lf = 0., Lv = 0.;
      for (int c = 0; c < C; c++) {
        if (f != root)
#pragma unroll
          for (int z = 0; z < A; z++) vec[z] = uf[(size_t)(c * A + z) * row];
        for (int x = 0; x < A; x++) Lf = add(Lf, mul(node_lik<A>(a, f, c, x, t, vec), __ldg(a.probs + c)));
        if (!leaf) {
#pragma unroll
          for (int z = 0; z < A; z++) vec[z] = uu[(size_t)(c * A + z) * row];
          for (int x = 0; x < A; x++) Lv = add(Lv, mul(node_lik<A>(a, v, c, x, t, vec), __ldg(a.probs + c)));
        }
      }
      for (int c = 0; c < C; c++) {
        const double pc = __ldg(a.probs + c);
        // the node's posteriors of this class first (acc), then the father's states one by one
#pragma unroll
        for (int z = 0; z < A; z++) vec[z] = uu[(size_t)(c * A + z) * row];
#pragma unroll 1
        for (int y = 0; y < A; y++)
          acc[y] = leaf ? mul(dn[(size_t)(c * A + y) * row], pc) : mul(node_lik<A>(a, v, c, y, t, vec), pc) / Lv;
        if (f != root)
#pragma unroll
          for (int z = 0; z < A; z++) vec[z] = uf[(size_t)(c * A + z) * row];
#pragma unroll 1
        for (int x = 0; x < A; x++) {
          const double pf = mul(node_lik<A>(a, f, c, x, t, vec), pc) / Lf;
          for (int y = 0; y < A; y++) res = add(res, mul(mul(pf, acc[y]), __ldg(Nv + c * AA + x * A + y)));
        }
      }
    }
    a.out[(size_t)v * a.n_pad + site] = res;
  }
}

} // namespace

static size_t map_variant_scratch_bytes(int n_nodes, int A, int C, int64_t chunk) {
  return (size_t)2 * n_nodes * C * A * chunk * sizeof(double) + (size_t)n_nodes * chunk;
}

int launch_map_variant(const MapModel& m, const MapBuffers& b, const VariantTables& vt, int mode, DevBuf& scratch, cudaStream_t st,
                       uint8_t* states_out) {
  if (mode < 1 || mode > 4 || (mode == 4) != (states_out != nullptr)) fail("internal: mapping variant %d", mode);
  // scratch budget 1 GiB: sites per launch, a multiple of 128
  const size_t per_site = map_variant_scratch_bytes(vt.n_nodes, m.A, m.C, 1);
  int64_t chunk = (int64_t)(((size_t)1 << 30) / per_site) / 128 * 128;
  if (chunk < 128) chunk = 128;
  if (chunk > (b.n + 127) / 128 * 128) chunk = (b.n + 127) / 128 * 128;
  scratch.reserve(map_variant_scratch_bytes(vt.n_nodes, m.A, m.C, chunk));
  VarArgs a;
  a.n_nodes = vt.n_nodes; a.C = m.C; a.mode = mode;
  a.parent = vt.parent; a.ch_off = vt.ch_off; a.ch = vt.ch; a.leaf_row = vt.leaf_row;
  a.P = vt.P; a.N = vt.N; a.pi = m.pi; a.probs = m.probs; a.code_mask = m.code_mask;
  a.tips = b.tips; a.n_pad = b.n_pad; a.chunk = chunk;
  a.down = scratch.as<double>();
  a.up = a.down + (size_t)vt.n_nodes * m.C * m.A * chunk;
  a.anc = reinterpret_cast<uint8_t*>(a.up + (size_t)vt.n_nodes * m.C * m.A * chunk);
  a.out = b.out;
  a.anc_out = states_out;
  int launches = 0;
  for (int64_t s0 = 0; s0 < b.n; s0 += chunk, launches++) {
    a.site0 = s0; a.n = std::min<int64_t>(chunk, b.n - s0);
    const unsigned grid = (unsigned)((a.n + 127) / 128);
    if (m.A == 4) k1_variant<4><<<grid, 128, 0, st>>>(a);
    else if (m.A == 20) k1_variant<20><<<grid, 128, 0, st>>>(a);
    else fail("no mapping kernel for A = %d", m.A);
    CMB_CUDA(cudaGetLastError());
  }
  return launches;
}

} // namespace cmb
