// K4: agglomerative clustering of sites on the device.
//
// Replaces Bio++ HierarchicalClustering(method, matrix, rootTree = false) + computeTree()
// (call sites CoMap.cpp:460-485, ClusterTools.cpp:260-262):
//   repeat while more than two clusters live: take the FIRST strictly smallest entry over
//   live pairs i<j in id order; the parent takes slot i, slot j dies; distances to every
//   other live k become w1 d(i,k) + w2 d(j,k) + w4 |d(i,k) - d(j,k)| with
//   (.5,.5,+.5) complete, (.5,.5,-.5) single, (n_i/(n_i+n_j), n_j/(n_i+n_j), 0) average;
//   height(parent) = d(i,j)/2; the last two clusters are joined at d/2.
// The reference rescans the whole matrix per merge (O(S^3)); here one persistent
// cooperative kernel keeps, per live row, the first minimum over live columns j>i
// (value + column), so a merge costs an argmin over S cached row minima, one row/column
// update and a rescan of the few rows whose cached minimum was invalidated -- O(S^2)
// in the common case -- with the reference's tie-breaking preserved exactly.
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <string>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace cmb {
namespace {

constexpr int CT = 512;

struct ClusterParams {
  int64_t S;
  int linkage;
  double* mat;          // [S][S] symmetric, consumed
  double* rmin_val;     // [S] first minimum over live j>i
  int32_t* rmin_idx;    // [S] its column, -1 if none
  uint8_t* alive;       // [S]
  double* len;          // [S] height of the cluster in the slot
  int32_t* nleaves;     // [S]
  int32_t* node;        // [S] dendrogram node in the slot
  int32_t* wl;          // [2][S] worklist: rows whose cached minimum has to be rescanned
  double* part_val;     // [grid] per-CTA first minimum of the new row a
  int32_t* part_idx;
  double* scan_val;     // [grid] per-CTA first minimum of its slice of the cached row minima
  int32_t* scan_idx;
  int32_t* scan_col;     // [grid] ... and the column of that minimum (saves a dependent load after the grid sync)
  double* seg_val;      // [grid] partial minima of row segments (a queued row is rescanned by several CTAs)
  int32_t* seg_idx;
  int32_t* seg_cnt;     // [grid] segments of a queued row finished so far
  int32_t* wl_count;    // [2]
  int32_t *left, *right; // [S-1]
  double* height;       // [S-1]
};

struct Best { double v; int i; };
__device__ __forceinline__ bool better(double v, int i, const Best& b) {
  return b.i < 0 || v < b.v || (v == b.v && i < b.i);
}
// Warp argmin with smallest-index tie-break in three REDUX steps on an order-preserving integer
// image of the double (high word, low word, index) instead of five shuffle + fp64-compare rounds;
// the result is valid in every lane.  -0.0 is folded into +0.0 first (they compare equal).
__device__ __forceinline__ Best warp_best(Best b) {
  const bool has = b.i >= 0;
  const long long bits = __double_as_longlong(__dadd_rn(b.v, 0.));
  const unsigned long long k = has ? (unsigned long long)(bits ^ ((bits >> 63) | (long long)0x8000000000000000ull)) : ~0ull;
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  const unsigned mi = __reduce_min_sync(0xffffffffu, has && hi == mhi && lo == mlo ? (unsigned)b.i : 0xffffffffu);
  if (mi == 0xffffffffu) return Best{0., -1};
  const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
  const long long mb = (mk >> 63) ? (long long)(mk ^ 0x8000000000000000ull) : (long long)~mk;
  return Best{__longlong_as_double(mb), (int)mi};
}
// block-wide argmin with smallest-index tie-break; result valid in every thread
__device__ Best block_best(Best b, Best* sh) {
  b = warp_best(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = b;
  __syncthreads();
  return warp_best(l < (int)(blockDim.x >> 5) ? sh[l] : Best{0., -1}); // every warp reduces the warp results itself
}

// first minimum of row r over live columns j>r (skipping `skip`).  The loads of a batch of
// U columns per thread are issued together (liveness byte and matrix entry are independent),
// so a 160 KB row costs a few L2 round trips instead of one per column.
__device__ Best scan_columns(const ClusterParams& p, int r, int skip, int64_t lo, int64_t hi, Best* sh) {
  constexpr int U = 8;
  Best b{0., -1};
  const double* row = p.mat + (size_t)r * p.S;
  for (int64_t base = lo; base < hi; base += (int64_t)blockDim.x * U) {
    double v[U];
    uint8_t al[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t j = base + (int64_t)u * blockDim.x + threadIdx.x;
      al[u] = j < hi ? p.alive[j] : 0;
      v[u] = j < hi ? row[j] : 0.;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t j = base + (int64_t)u * blockDim.x + threadIdx.x;
      if (al[u] && j != skip && !(v[u] != v[u]) && better(v[u], (int)j, b)) { b.v = v[u]; b.i = (int)j; }
    }
  }
  return block_best(b, sh);
}
__device__ void rescan_row(const ClusterParams& p, int r, int skip, Best* sh) {
  const Best b = scan_columns(p, r, skip, r + 1, p.S, sh);
  if (threadIdx.x == 0) { p.rmin_val[r] = b.v; p.rmin_idx[r] = b.i; }
}

__global__ void __launch_bounds__(CT) k4_init_rows(ClusterParams p) {
  __shared__ Best sh[32];
  for (int r = blockIdx.x; r < p.S; r += gridDim.x) {
    rescan_row(p, r, -1, sh);
    __syncthreads();
  }
}

// One merge = (1) every CTA finds the first global minimum among the cached row minima
// (dead rows carry index -1); (A) the grid rewrites row / column a, updates in place the cached
// minimum of every row k < a that only has to compare its new entry, queues the rows whose
// cached minimum pointed at a or b for a rescan, and reduces the new row a's own minimum on the
// fly (per-CTA partial); grid sync; (B) bookkeeping, row a's minimum from the partials, rescans;
// grid sync.  ncu / clock64 on config 5 (S = 20 000) before this layout: 73 us per merge, of
// which 21 us in (1) and 47 us in serial rescans of row a and ~40 queued rows.  Now (-DCMB_K4_TIMING,
// cycles per merge on a mid-grid CTA): scan 3300 | barrier 2800 | combine 3400 | update 1400 |
// barrier 6400 (waits for CTA 0's 4000) | rescans 2800 (slowest CTA 8200) | barrier 10000.  A barrier
// costs ~2500 cycles and a round trip to data another SM has just written ~1000, whatever the grid
// size (16..148 CTAs measured the same).  A two-barrier variant that rescans a queued row inside the
// CTA that found it (kept as tools/experiments/k4_two_barrier_variant.cu.txt, parity-green) was slower, 29 us per
// merge: one CTA needs five dependent batches per 160 KB row while the rest of the grid waits.  Per-CTA
// queue slots instead of the global atomic worklist, and keeping block 0 out of the row update, changed
// nothing either: the second barrier costs 6000 cycles with or without them (its fence waits for the
// scattered column-a stores).  What did help: argmin reductions through REDUX on an order-preserving
// integer image of the doubles instead of shuffle + fp64-compare rounds (every phase ends in one):
// 15.5 -> 12 us per merge (scan 1800 | barrier 2700 | combine 2000 | update 730 | barrier 4900 |
// rescans 2300, slowest CTA 6700 | barrier 8200).  A hand-written barrier (red.release.gpu on a monotone
// counter + ld.acquire.gpu poll) instead of cooperative_groups' grid.sync() measured the same 2700 cycles.
__global__ void __launch_bounds__(CT) k4_cluster(ClusterParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ Best sh[32];
  const int64_t S = p.S;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
#ifdef CMB_K4_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
#define K4_T(slot) { t1 = clock64(); tacc[slot] += t1 - t0; t0 = t1; }
#else
#define K4_T(slot)
#endif
  for (int64_t step = 0; step + 2 < S; step++) {
#ifdef CMB_K4_TIMING
    t0 = clock64();
#endif
    // (1) first global minimum of the cached row minima: every CTA scans its slice (one batch of
    //     loads), then all CTAs reduce the per-CTA partials; the extra grid sync (~1.4 us) costs
    //     less than the five dependent L2 round trips of a full scan per CTA
    Best b{0., -1};
    __shared__ int sh_col;
    int my_i = -1, my_col = -1; // this thread's candidate row and the column of its minimum
    {
      const int64_t per = (S + gridDim.x - 1) / gridDim.x, lo = (int64_t)blockIdx.x * per, hi = lo + per < S ? lo + per : S;
      for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int ix = p.rmin_idx[i];
        const double v = p.rmin_val[i];
        if (ix >= 0 && better(v, (int)i, b)) { b.v = v; b.i = (int)i; my_col = ix; }
      }
      my_i = b.i;
      b = block_best(b, sh);
      if (threadIdx.x == 0) { p.scan_val[blockIdx.x] = b.v; p.scan_idx[blockIdx.x] = b.i; }
      if (my_i >= 0 && my_i == b.i) p.scan_col[blockIdx.x] = my_col; // a row belongs to one thread: one writer
      K4_T(0)
      grid.sync();
      K4_T(1)
      // the column travels with the partial, so no load depends on the reduced row index
      b = Best{0., -1};
      my_i = -1;
      for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) {
        const int i = p.scan_idx[c];
        const double v = p.scan_val[c];
        const int col = p.scan_col[c];
        if (i >= 0 && better(v, i, b)) { b.v = v; b.i = i; my_col = col; }
      }
      my_i = b.i;
      __syncthreads(); // sh is reused
    }
    b = block_best(b, sh);
    const int a = b.i;
    if (a < 0) return; // nothing mergeable (NaN distances): host reports the error
    if (my_i == a) sh_col = my_col; // rows are unique across the partials: one writer
    __syncthreads();
    const int bb = sh_col;
    const double dab = b.v;
    double w1, w2, w4;
    if (p.linkage == 1) { w1 = .5; w2 = .5; w4 = -.5; }
    else if (p.linkage == 0) { w1 = .5; w2 = .5; w4 = .5; }
    else {
      double na = (double)p.nleaves[a], nb = (double)p.nleaves[bb];
      w1 = na / (na + nb); w2 = nb / (na + nb); w4 = 0.;
    }
    __syncthreads(); // sh is reused below
    K4_T(2)
    // (A) new distances to the merged cluster (slot a)
    int32_t* wl = p.wl + (step & 1) * S;
    int32_t* wlc = p.wl_count + (step & 1);
    Best ra{0., -1}; // first minimum of the new row a over live columns k > a
    // cached minima are still being read by CTAs that are in (1): in-place updates wait for (B)
    constexpr int kMaxPending = 4; // rows per thread per merge: S <= 4 * grid threads (300 k at 148 x 512)
    int pend_k[kMaxPending];
    double pend_v[kMaxPending];
    int n_pend = 0;
    for (int64_t k = gtid; k < S; k += gsz) {
      // one round trip: everything this row can need is loaded before the first branch
      const uint8_t live = p.alive[k];
      const double d1 = p.mat[(size_t)a * S + k], d2 = p.mat[(size_t)bb * S + k];
      const int ci = p.rmin_idx[k];
      const double cv = p.rmin_val[k];
      if (k == a || k == bb || !live) continue;
      // left-to-right, unfused, as the reference's C++ expression evaluates
      const double nd = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w1, d1), __dmul_rn(w2, d2)), __dmul_rn(0., dab)),
                                  __dmul_rn(w4, fabs(__dadd_rn(d1, -d2))));
      p.mat[(size_t)a * S + k] = nd;
      p.mat[(size_t)k * S + a] = nd;
      if (k > a) {
        if (!(nd != nd) && better(nd, (int)k, ra)) { ra.v = nd; ra.i = (int)k; }
        if (k < bb && ci == bb) wl[atomicAdd(wlc, 1)] = (int32_t)k; // lost its minimum's column
      } else {
        if (ci == a || ci == bb) wl[atomicAdd(wlc, 1)] = (int32_t)k;            // minimum pointed at a merged slot
        else if (!(nd != nd) && (ci < 0 || nd < cv || (nd == cv && a < ci)) && n_pend < kMaxPending) {
          pend_k[n_pend] = (int)k; pend_v[n_pend] = nd; n_pend++;              // only column a changed: compare
        }
      }
    }
    ra = block_best(ra, sh);
    if (threadIdx.x == 0) { p.part_val[blockIdx.x] = ra.v; p.part_idx[blockIdx.x] = ra.i; }
    K4_T(3)
    grid.sync();
    K4_T(4)
    // (B) deferred in-place updates, bookkeeping, row a's cached minimum from the per-CTA
    //     partials, queued rescans
#pragma unroll
    for (int q = 0; q < kMaxPending; q++)
      if (q < n_pend) { p.rmin_val[pend_k[q]] = pend_v[q]; p.rmin_idx[pend_k[q]] = a; }
    if (blockIdx.x == 0) {
      // slot state of a and bb, loaded before the reduction of the partials hides their latency
      double len_a = 0.; int node_a = 0, node_b = 0, nl_a = 0, nl_b = 0;
      if (threadIdx.x == 0) { len_a = p.len[a]; node_a = p.node[a]; node_b = p.node[bb]; nl_a = p.nleaves[a]; nl_b = p.nleaves[bb]; }
      Best t{0., -1};
      for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) {
        const int i = p.part_idx[c];
        const double v = p.part_val[c];
        if (i >= 0 && better(v, i, t)) { t.v = v; t.i = i; }
      }
      t = block_best(t, sh);
      if (threadIdx.x == 0) {
        const int32_t parent = (int32_t)(S + step);
        const double half = dab / 2.;
        const double d0 = half - len_a;
        p.left[step] = node_a;
        p.right[step] = node_b;
        p.height[step] = len_a + d0;
        p.node[a] = parent;
        p.len[a] = len_a + d0;
        p.nleaves[a] = nl_a + nl_b;
        p.alive[bb] = 0;
        p.rmin_idx[bb] = -1;  // dead rows drop out of (1)
        p.rmin_val[a] = t.v;
        p.rmin_idx[a] = t.i;
        p.wl_count[(step + 1) & 1] = 0;
      }
      __syncthreads();
    }
    const int n_wl = *wlc;
    // Every queued row is rescanned by nseg CTAs (one batch of loads each instead of up to five
    // dependent ones); the CTA that finishes a row's last segment combines the partial minima.
    const int G = (int)gridDim.x;
    const int nseg = n_wl > 0 && n_wl * 2 <= G ? (G / n_wl < 8 ? G / n_wl : 8) : 1;
    // block 0 is busy with the bookkeeping: the work starts at the other end of the grid
    for (int it = G - 1 - (int)blockIdx.x; it < n_wl * nseg; it += G) {
      const int w = it / nseg, seg = it % nseg;
      const int r = wl[w];
      if (nseg == 1) rescan_row(p, r, bb, sh);
      else {
        const int64_t L = S - r - 1, lo = r + 1 + L * seg / nseg, hi = r + 1 + L * (seg + 1) / nseg;
        const Best t = scan_columns(p, r, bb, lo, hi, sh);
        __shared__ int last;
        if (threadIdx.x == 0) {
          p.seg_val[w * nseg + seg] = t.v;
          p.seg_idx[w * nseg + seg] = t.i;
          __threadfence();
          last = atomicAdd(&p.seg_cnt[w], 1) == nseg - 1;
        }
        __syncthreads();
        if (last && threadIdx.x < 32) {
          __threadfence();
          Best c{0., -1};
          if ((int)threadIdx.x < nseg) {
            c.i = __ldcg(&p.seg_idx[w * nseg + threadIdx.x]);
            c.v = __ldcg(&p.seg_val[w * nseg + threadIdx.x]);
          }
          c = warp_best(c);
          if (threadIdx.x == 0) { p.rmin_val[r] = c.v; p.rmin_idx[r] = c.i; p.seg_cnt[w] = 0; }
        }
      }
      __syncthreads();
    }
    K4_T(5)
    grid.sync();
    K4_T(6)
  }
#ifdef CMB_K4_TIMING
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77 || blockIdx.x == gridDim.x - 1))
    printf("k4 timing cta %d: scan %lld sync1 %lld combine %lld update %lld sync2 %lld rescans %lld sync3 %lld (cycles per merge)\n",
           blockIdx.x, tacc[0] / (S - 2), tacc[1] / (S - 2), tacc[2] / (S - 2), tacc[3] / (S - 2), tacc[4] / (S - 2), tacc[5] / (S - 2), tacc[6] / (S - 2));
#endif
  // finalStep: join the last two clusters at d/2
  if (gtid == 0) {
    int i1 = -1, i2 = -1;
    for (int64_t i = 0; i < S; i++)
      if (p.alive[i]) { if (i1 < 0) i1 = (int)i; else i2 = (int)i; }
    const double d = p.mat[(size_t)i1 * S + i2] / 2;
    p.left[S - 2] = p.node[i1];
    p.right[S - 2] = p.node[i2];
    p.height[S - 2] = p.len[i1] + (d - p.len[i1]);
  }
}

// ---- the same merge loop inside ONE thread-block cluster ----------------------------------
// A grid barrier costs ~2500 cycles and every exchange through global memory ~1000 more; a merge
// has three of each.  Sixteen CTAs of one cluster hold the cached row minima in their shared memory
// (slice r of the rows in CTA r), exchange partial minima and the rescan queue with remote
// shared-memory stores and meet at cluster barriers; only the matrix itself stays in HBM / L2.
// Per merge: (P1) slice argmin -> partial to every CTA; barrier; (P2) every CTA reduces the 16
// partials, updates row / column a for the rows of its slice (cached minima local), sends its
// partial of the new row a to that row's owner and its queued rows to everybody; barrier; (P3) queued
// rows are dealt round-robin, each rescanned by one CTA, result stored into the owner's cache;
// barrier.  Measured (S = 20 000, cycles per merge): slice argmin 2100 | barrier 1100 | reduce 1250 | update 3600 |
// publish 2700 | barrier 1500+ | rescans up to 13 000 on the CTAs that hold a row | barrier: 16 us per merge.
// Splitting every queued row over the 16 CTAs (as the grid kernel does) is the missing step; until then the
// grid kernel is the default and this one is selected with CMB_K4_LAYOUT=cluster.
constexpr int NC = 16, NT = 1024, QMAX = 8;

struct DsmLayout {
  int per; size_t o_rv, o_pv, o_rav, o_ri, o_pi, o_pc, o_rai, o_rn, o_rl, o_ovf, o_alive, bytes;
  __host__ __device__ explicit DsmLayout(int64_t S) {
    per = (int)((S + NC - 1) / NC);
    size_t o = 0;
    o_rv = o; o += 8 * (size_t)per; o_pv = o; o += 8 * NC; o_rav = o; o += 8 * NC;
    o_ri = o; o += 4 * (size_t)per; o_pi = o; o += 4 * NC; o_pc = o; o += 4 * NC; o_rai = o; o += 4 * NC;
    o_rn = o; o += 4 * NC; o_rl = o; o += 4 * NC * QMAX; o_ovf = o; o += 4 * (size_t)per;
    o_alive = o; o += (size_t)S; bytes = (o + 15) & ~size_t(15);
  }
};

__device__ Best scan_row_dsm(const double* __restrict__ row, const uint8_t* alive_s, int r, int64_t S, Best* sh) {
  constexpr int U = 10; // 10 240 columns per batch of loads
  Best b{0., -1};
  for (int64_t base = r + 1; base < S; base += (int64_t)NT * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t j = base + (int64_t)u * NT + threadIdx.x;
      v[u] = j < S ? row[j] : 0.;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t j = base + (int64_t)u * NT + threadIdx.x;
      if (j < S && alive_s[j] && !(v[u] != v[u]) && better(v[u], (int)j, b)) { b.v = v[u]; b.i = (int)j; }
    }
  }
  return block_best(b, sh);
}

__global__ void __launch_bounds__(NT, 1) k4_cluster_dsm(ClusterParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), tid = (int)threadIdx.x;
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ Best sh[32];
  __shared__ int sh_col, sh_a, sh_bb, rs_n, sh_ntot, sh_list[NC * QMAX], rs_loc[QMAX];
  __shared__ double sh_d;
  const int64_t S = p.S;
  const DsmLayout L(S);
  const int per = L.per, lo = rank * per, hi = (int)((int64_t)lo + per < S ? lo + per : S);
  double* rv = (double*)(dsm + L.o_rv); double* pv = (double*)(dsm + L.o_pv); double* rav = (double*)(dsm + L.o_rav);
  int* ri = (int*)(dsm + L.o_ri); int* pi = (int*)(dsm + L.o_pi); int* pc = (int*)(dsm + L.o_pc);
  int* rai = (int*)(dsm + L.o_rai); int* rn = (int*)(dsm + L.o_rn); int* rl = (int*)(dsm + L.o_rl);
  int* ovf = (int*)(dsm + L.o_ovf); uint8_t* alive_s = dsm + L.o_alive;
  for (int i = lo + tid; i < hi; i += NT) { rv[i - lo] = p.rmin_val[i]; ri[i - lo] = p.rmin_idx[i]; }
  for (int64_t i = tid; i < S; i += NT) alive_s[i] = 1;
  __syncthreads();
  cluster.sync(); // every CTA of the cluster is running: remote shared memory may be written
#ifdef CMB_K4_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
#endif
  for (int64_t step = 0; step + 2 < S; step++) {
#ifdef CMB_K4_TIMING
    t0 = clock64();
#endif
    // (P1) first minimum of this CTA's slice of the cached row minima, published to every CTA
    Best b{0., -1};
    int my_col = -1;
    for (int i = lo + tid; i < hi; i += NT) {
      const int ix = ri[i - lo];
      const double v = rv[i - lo];
      if (ix >= 0 && better(v, i, b)) { b.v = v; b.i = i; my_col = ix; }
    }
    const int my_i = b.i;
    b = block_best(b, sh);
    if (my_i >= 0 && my_i == b.i) sh_col = my_col; // a row belongs to one thread: one writer
    __syncthreads();
    if (tid < NC) {
      cluster.map_shared_rank(pv, tid)[rank] = b.v;
      cluster.map_shared_rank(pi, tid)[rank] = b.i;
      cluster.map_shared_rank(pc, tid)[rank] = b.i >= 0 ? sh_col : -1;
    }
    K4_T(0)
    cluster.sync();
    K4_T(1)
    // (P2) every CTA reduces the 16 partials
    if (tid < 32) {
      Best c{0., -1};
      int col = -1;
      if (tid < NC) { c.v = pv[tid]; c.i = pi[tid]; col = pc[tid]; }
      const Best w = warp_best(c);
      if (w.i < 0) { if (tid == 0) sh_a = -1; }
      else if (c.i == w.i) { sh_a = w.i; sh_bb = col; sh_d = w.v; } // rows are unique across the partials
      if (tid == 0) rs_n = 0;
    }
    __syncthreads();
    const int a = sh_a;
    if (a < 0) return; // nothing mergeable (NaN distances), the same in every CTA: host reports the error
    const int bb = sh_bb;
    const double dab = sh_d;
    double w1, w2, w4;
    if (p.linkage == 1) { w1 = .5; w2 = .5; w4 = -.5; }
    else if (p.linkage == 0) { w1 = .5; w2 = .5; w4 = .5; }
    else {
      double na = (double)p.nleaves[a], nb = (double)p.nleaves[bb];
      w1 = na / (na + nb); w2 = nb / (na + nb); w4 = 0.;
    }
    // dendrogram bookkeeping (rank 0, thread 0): loads now, stores in (P3)
    double len_a = 0.; int node_a = 0, node_b = 0, nl_a = 0, nl_b = 0;
    if (rank == 0 && tid == 0) { len_a = p.len[a]; node_a = p.node[a]; node_b = p.node[bb]; nl_a = p.nleaves[a]; nl_b = p.nleaves[bb]; }
    if (tid == 0) alive_s[bb] = 0; // every CTA keeps its own copy of the liveness bytes
    K4_T(2)
    Best ra{0., -1}; // first minimum of the new row a over live columns k > a (this slice)
    for (int k = lo + tid; k < hi; k += NT) {
      const int kl = k - lo;
      const uint8_t live = alive_s[k];
      const double d1 = p.mat[(size_t)a * S + k], d2 = p.mat[(size_t)bb * S + k];
      const int ci = ri[kl];
      const double cv = rv[kl];
      if (k == bb) { ri[kl] = -1; continue; } // dead rows drop out of (P1)
      if (k == a || !live) continue;
      // left-to-right, unfused, as the reference's C++ expression evaluates
      const double nd = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w1, d1), __dmul_rn(w2, d2)), __dmul_rn(0., dab)),
                                  __dmul_rn(w4, fabs(__dadd_rn(d1, -d2))));
      p.mat[(size_t)a * S + k] = nd;
      p.mat[(size_t)k * S + a] = nd;
      bool rescan = false;
      if (k > a) {
        if (!(nd != nd) && better(nd, k, ra)) { ra.v = nd; ra.i = k; }
        rescan = k < bb && ci == bb;                                     // lost its minimum's column
      } else {
        if (ci == a || ci == bb) rescan = true;                         // minimum pointed at a merged slot
        else if (!(nd != nd) && (ci < 0 || nd < cv || (nd == cv && a < ci))) { rv[kl] = nd; ri[kl] = a; }
      }
      if (rescan) {
        const int q = atomicAdd(&rs_n, 1);
        if (q < QMAX) rs_loc[q] = k; else ovf[q - QMAX] = k;
      }
    }
    K4_T(3)
    ra = block_best(ra, sh); // its barriers also publish the queue and this CTA's column-a stores to the CTA
    const int n_found = rs_n, n_q = n_found < QMAX ? n_found : QMAX;
    if (tid == 0) {
      const int oa = a / per;
      cluster.map_shared_rank(rav, oa)[rank] = ra.v;
      cluster.map_shared_rank(rai, oa)[rank] = ra.i;
    }
    if (tid < NC) {
      cluster.map_shared_rank(rn, tid)[rank] = n_q;
      int* dst = cluster.map_shared_rank(rl, tid) + rank * QMAX;
      for (int q = 0; q < n_q; q++) dst[q] = rs_loc[q];
    }
    // more than QMAX queued rows in one slice (rare): rescanned here; row k changes only in column a,
    // which this CTA has just written
    for (int q = QMAX; q < n_found; q++) {
      __syncthreads();
      const int r = ovf[q - QMAX];
      const Best t = scan_row_dsm(p.mat + (size_t)r * S, alive_s, r, S, sh);
      if (tid == 0) { rv[r - lo] = t.v; ri[r - lo] = t.i; }
    }
    K4_T(4)
    cluster.sync();
    K4_T(5)
    // (P3) row a's minimum at its owner, queued rows dealt round-robin, bookkeeping
    if (a / per == rank && tid < 32) {
      Best c{0., -1};
      if (tid < NC) { c.v = rav[tid]; c.i = rai[tid]; }
      c = warp_best(c);
      if (tid == 0) { rv[a - lo] = c.v; ri[a - lo] = c.i; }
    }
    if (tid < NC * QMAX) {
      const int src = tid / QMAX, q = tid % QMAX;
      int pos = q, tot = 0;
      for (int r2 = 0; r2 < NC; r2++) { const int c = rn[r2]; if (r2 < src) pos += c; tot += c; }
      if (q < rn[src]) sh_list[pos] = rl[tid];
      if (tid == 0) sh_ntot = tot;
    }
    __syncthreads();
    const int n_tot = sh_ntot;
    for (int w = rank; w < n_tot; w += NC) {
      const int r = sh_list[w];
      const Best t = scan_row_dsm(p.mat + (size_t)r * S, alive_s, r, S, sh);
      if (tid == 0) {
        const int o = r / per;
        cluster.map_shared_rank(rv, o)[r - o * per] = t.v;
        cluster.map_shared_rank(ri, o)[r - o * per] = t.i;
      }
      __syncthreads();
    }
    if (rank == 0 && tid == 0) {
      const double half = dab / 2.;
      const double d0 = half - len_a;
      p.left[step] = node_a;
      p.right[step] = node_b;
      p.height[step] = len_a + d0;
      p.node[a] = (int32_t)(S + step);
      p.len[a] = len_a + d0;
      p.nleaves[a] = nl_a + nl_b;
    }
    K4_T(6)
    cluster.sync();
    K4_T(7)
  }
#ifdef CMB_K4_TIMING
  if (tid == 0 && (rank == 0 || rank == 7 || rank == 15))
    printf("k4 dsm timing rank %d: scan %lld b1 %lld reduce %lld update %lld publish %lld b2 %lld rescans %lld b3 %lld (cycles per merge)\n",
           rank, tacc[0] / (S - 2), tacc[1] / (S - 2), tacc[2] / (S - 2), tacc[3] / (S - 2), tacc[4] / (S - 2), tacc[5] / (S - 2), tacc[6] / (S - 2), tacc[7] / (S - 2));
#endif
  // finalStep: join the last two clusters at d/2
  if (rank == 0 && tid == 0) {
    int i1 = -1, i2 = -1;
    for (int64_t i = 0; i < S; i++)
      if (alive_s[i]) { if (i1 < 0) i1 = (int)i; else i2 = (int)i; }
    const double d = p.mat[(size_t)i1 * S + i2] / 2;
    p.left[S - 2] = p.node[i1];
    p.right[S - 2] = p.node[i2];
    p.height[S - 2] = p.len[i1] + (d - p.len[i1]);
  }
}

// launches the cluster kernel when its shared-memory caches fit and the device can place the cluster
bool try_cluster_dsm(ClusterParams& p, cudaStream_t st) {
  // opt-in (CMB_K4_LAYOUT=cluster): parity-green, but at S = 20 000 it needs 16 us per merge against 12 us for
  // the grid kernel -- a queued row is rescanned by ONE CTA (13 000 cycles) while the other 15 wait
  const char* e = std::getenv("CMB_K4_LAYOUT");
  if (!e || std::string(e) != "cluster") return false;
  const DsmLayout L(p.S);
  if (L.bytes > 200 * 1024) return false;
  if (cudaFuncSetAttribute(k4_cluster_dsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes) != cudaSuccess ||
      cudaFuncSetAttribute(k4_cluster_dsm, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = L.bytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&n_clusters, k4_cluster_dsm, &cfg) != cudaSuccess || n_clusters < 1) {
    cudaGetLastError();
    return false;
  }
  CMB_CUDA(cudaLaunchKernelEx(&cfg, k4_cluster_dsm, p));
  return true;
}

__global__ void k4_init_state(ClusterParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.S) return;
  p.alive[i] = 1;
  p.len[i] = 0.;
  p.nleaves[i] = 1;
  p.node[i] = (int32_t)i;
  if (i < 2) p.wl_count[i] = 0;
  if (i < p.S - 1) { p.left[i] = -1; p.right[i] = -1; p.height[i] = 0.; }
}

// Compensation group statistic (Statistics.h:267-294): 1 - ||sum_j v_j|| / sum_j ||v_j||,
// one thread per group, branches in id order, unfused arithmetic.
__global__ void k4_group_compensation(int64_t n_groups, const int32_t* __restrict__ members,
                                      const int64_t* __restrict__ offsets, int B, int64_t S_pad,
                                      const double* __restrict__ out, double* __restrict__ stat) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int64_t m0 = offsets[g], m1 = offsets[g + 1];
  double sumsq2 = 0., sumnorms = 0.;
  for (int b = 0; b < B; b++) {
    double s = 0.;
    for (int64_t m = m0; m < m1; m++) s = __dadd_rn(s, out[(size_t)b * S_pad + members[m]]);
    sumsq2 = __dadd_rn(sumsq2, __dmul_rn(s, s));
  }
  for (int64_t m = m0; m < m1; m++) {
    double q = 0.;
    for (int b = 0; b < B; b++) {
      double v = out[(size_t)b * S_pad + members[m]];
      q = __dadd_rn(q, __dmul_rn(v, v));
    }
    sumnorms = __dadd_rn(sumnorms, sqrt(q));
  }
  stat[g] = __dadd_rn(1., -(sqrt(sumsq2) / sumnorms));
}

} // namespace

int launch_cluster(int64_t S, int linkage, double* mat, DevBuf& work, int32_t* left_dev, int32_t* right_dev,
                   double* height_dev, cudaStream_t st) {
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t o_val = 0, o_idx = al(o_val + 8 * S), o_alive = al(o_idx + 4 * S), o_len = al(o_alive + S),
         o_nl = al(o_len + 8 * S), o_node = al(o_nl + 4 * S), o_wl = al(o_node + 4 * S), o_wlc = al(o_wl + 8 * S),
         o_pv = al(o_wlc + 64), o_pi = al(o_pv + 8 * 1024), o_sv = al(o_pi + 4 * 1024), o_si = al(o_sv + 8 * 1024),
         o_sc = al(o_si + 4 * 1024),
         o_gv = al(o_sc + 4 * 1024), o_gi = al(o_gv + 8 * 1024), o_gc = al(o_gi + 4 * 1024), o_end = al(o_gc + 4 * 1024);
  work.reserve(o_end);
  unsigned char* w = work.as<unsigned char>();
  ClusterParams p;
  p.S = S; p.linkage = linkage; p.mat = mat;
  p.rmin_val = (double*)(w + o_val); p.rmin_idx = (int32_t*)(w + o_idx); p.alive = w + o_alive;
  p.len = (double*)(w + o_len); p.nleaves = (int32_t*)(w + o_nl); p.node = (int32_t*)(w + o_node);
  p.wl = (int32_t*)(w + o_wl); p.wl_count = (int32_t*)(w + o_wlc);
  p.part_val = (double*)(w + o_pv); p.part_idx = (int32_t*)(w + o_pi);
  p.scan_val = (double*)(w + o_sv); p.scan_idx = (int32_t*)(w + o_si); p.scan_col = (int32_t*)(w + o_sc);
  p.seg_val = (double*)(w + o_gv); p.seg_idx = (int32_t*)(w + o_gi); p.seg_cnt = (int32_t*)(w + o_gc);
  CMB_CUDA(cudaMemsetAsync(p.seg_cnt, 0, 4 * 1024, st));
  p.left = left_dev; p.right = right_dev; p.height = height_dev;
  k4_init_state<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(p);
  CMB_CUDA(cudaGetLastError());
  int dev = 0, sms = 0, per_sm = 0;
  CMB_CUDA(cudaGetDevice(&dev));
  CMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_cluster, CT, 0));
  if (per_sm < 1) fail("k4_cluster cannot be made resident");
  int grid = std::min(sms, 1024); // one CTA per SM
  if (S > (int64_t)4 * grid * CT) fail("clustering: %lld sites exceed the %lld this device handles per merge pass", (long long)S,
                                       (long long)4 * grid * CT);
  k4_init_rows<<<grid, CT, 0, st>>>(p);
  CMB_CUDA(cudaGetLastError());
  if (try_cluster_dsm(p, st)) return 3;
  void* args[] = {&p};
  CMB_CUDA(cudaLaunchCooperativeKernel((void*)k4_cluster, dim3(grid), dim3(CT), args, 0, st));
  return 3;
}

void launch_group_compensation(int64_t n_groups, const int32_t* members, const int64_t* offsets, int B,
                               int64_t S_pad, const double* out, double* stat, cudaStream_t st) {
  if (n_groups == 0) return;
  k4_group_compensation<<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(n_groups, members, offsets, B, S_pad, out, stat);
  CMB_CUDA(cudaGetLastError());
}

} // namespace cmb
