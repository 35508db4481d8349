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
__device__ __forceinline__ Best warp_best(Best b) {
  for (int o = 16; o > 0; o >>= 1) {
    double v = __shfl_down_sync(0xffffffffu, b.v, o);
    int i = __shfl_down_sync(0xffffffffu, b.i, o);
    if (i >= 0 && better(v, i, b)) { b.v = v; b.i = i; }
  }
  return b;
}
// block-wide argmin with smallest-index tie-break; result valid in every thread
__device__ Best block_best(Best b, Best* sh) {
  b = warp_best(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = b;
  __syncthreads();
  if (w == 0) {
    Best t = l < (int)(blockDim.x >> 5) ? sh[l] : Best{0., -1};
    t = warp_best(t);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
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
// CTA that found it (kept as scratch/k4_two_barrier_variant.cu.txt, parity-green) was slower, 29 us per
// merge: one CTA needs five dependent batches per 160 KB row while the rest of the grid waits.  Per-CTA
// queue slots instead of the global atomic worklist, and keeping block 0 out of the row update, changed
// nothing either: the second barrier costs 6000 cycles with or without them (its fence waits for the
// scattered column-a stores).
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
