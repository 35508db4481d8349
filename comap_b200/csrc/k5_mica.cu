// K5: Mica's statistics (CoMap/Mica.cpp) -- mutual information, joint entropy and entropy of alignment COLUMNS
// (no substitution mapping: the mapping only supplies the norms that condition the null, Mica.cpp:336-339).
//
// Replaces SiteTools::mutualInformation / jointEntropy / entropy(site, resolveUnknowns = true) at Mica.cpp:92, 354,
// 357, 430-431, 524-525, 574-575, 660 [Bio++ bpp-seq, from memory; the reference ships no output of mica: parity
// unpinned].  A character compatible with k states adds 1/k to each of them (a pair 1/(k1 k2) to each combination),
// frequencies = counts / number of sequences, the joint table is renormalised by its total and the marginals are
// taken from it; sums run rows then columns, as restated in the oracle (orc_site_pair).
//
// Layout: the alignment is the tip matrix K1 reads, [T][S_pad] codes with the 256-entry code -> state-mask table.
// One thread per pair keeps the A x A joint table in local memory (A = 4: 16 doubles, A = 20: 400) -- T byte loads
// per column (coalesced over the 32 pairs of a warp, which share site i and walk consecutive sites j), A^2 updates
// and A^2 logarithms per pair.  Not a tiled kernel: the work per pair is a histogram, not a dot product.
#include "kernels.h"
#include "device_utils.cuh"

namespace cmb {
namespace {

template <int A>
__device__ __forceinline__ void pair_stats(const uint8_t* __restrict__ c1, size_t s1, const uint8_t* __restrict__ c2, size_t s2,
                                           int T, const uint32_t* __restrict__ cmask, double& mi, double& hj) {
  constexpr uint32_t full = A >= 32 ? 0xffffffffu : (1u << A) - 1u;
  double cnt[A * A];
#pragma unroll
  for (int k = 0; k < A * A; k++) cnt[k] = 0.;
  for (int t = 0; t < T; t++) {
    const uint32_t m1 = __ldg(cmask + c1[(size_t)t * s1]) & full, m2 = __ldg(cmask + c2[(size_t)t * s2]) & full;
    const int k1 = __popc(m1), k2 = __popc(m2);
    if (k1 == 1 && k2 == 1) cnt[(__ffs(m1) - 1) * A + __ffs(m2) - 1] += 1.;
    else if (k1 && k2) {
      const double w = 1. / ((double)k1 * (double)k2);
      for (int x = 0; x < A; x++)
        if ((m1 >> x) & 1u)
          for (int y = 0; y < A; y++)
            if ((m2 >> y) & 1u) cnt[x * A + y] += w;
    }
  }
  double p1[A], p2[A];
#pragma unroll
  for (int x = 0; x < A; x++) { p1[x] = 0.; p2[x] = 0.; }
  double tot = 0.;
  const double n = (double)T;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      const double pxy = cnt[x * A + y] / n;
      tot += pxy; p1[x] += pxy; p2[y] += pxy;
    }
  for (int x = 0; x < A; x++) { p1[x] /= tot; p2[x] /= tot; }
  double m = 0., h = 0.;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      const double pxy = cnt[x * A + y] / n / tot;
      if (pxy > 0.) { m += pxy * log(pxy / (p1[x] * p2[y])); h += pxy * log(pxy); }
    }
  mi = m; hj = -h;
}

template <int A>
__global__ void __launch_bounds__(128) k5_entropy(int T, int64_t n, int64_t n_pad, const uint8_t* __restrict__ tips,
                                                  const uint32_t* __restrict__ cmask, double* __restrict__ entropy) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  constexpr uint32_t full = A >= 32 ? 0xffffffffu : (1u << A) - 1u;
  double cnt[A];
#pragma unroll
  for (int x = 0; x < A; x++) cnt[x] = 0.;
  for (int t = 0; t < T; t++) {
    const uint32_t m = __ldg(cmask + tips[(size_t)t * n_pad + s]) & full;
    const int k = __popc(m);
    if (!k) continue;
    const double w = 1. / (double)k;
    for (int x = 0; x < A; x++)
      if ((m >> x) & 1u) cnt[x] += w;
  }
  double h = 0.;
  for (int x = 0; x < A; x++) {
    const double f = cnt[x] / (double)T;
    if (f != 0.) h += f * log(f);
  }
  entropy[s] = -h;
}

// all pairs i < j of one alignment, dense upper triangle in the reference's order (i ascending, j ascending)
template <int A>
__global__ void __launch_bounds__(128) k5_pairs(int T, int64_t S, int64_t n_pad, const uint8_t* __restrict__ tips,
                                                const uint32_t* __restrict__ cmask, double* __restrict__ mi,
                                                double* __restrict__ hj) {
  const int64_t i = blockIdx.y;
  const int64_t j = i + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  double m, h;
  pair_stats<A>(tips + i, (size_t)n_pad, tips + j, (size_t)n_pad, T, cmask, m, h);
  const int64_t idx = i * S - i * (i + 1) / 2 + (j - i - 1);
  mi[idx] = m;
  if (hj) hj[idx] = h;
}

// site a[r] against site b[r] (possibly of another tip matrix): the null's j <-> j pairs and the bootstrap's pair lists
template <int A>
__global__ void __launch_bounds__(128) k5_listed(int T, int64_t n, const uint8_t* __restrict__ t1, int64_t np1,
                                                 const uint8_t* __restrict__ t2, int64_t np2, const int32_t* __restrict__ a,
                                                 const int32_t* __restrict__ b, const uint32_t* __restrict__ cmask,
                                                 double* __restrict__ mi, double* __restrict__ hj) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t s1 = a ? a[r] : r, s2 = b ? b[r] : r;
  double m, h;
  pair_stats<A>(t1 + s1, (size_t)np1, t2 + s2, (size_t)np2, T, cmask, m, h);
  mi[r] = m;
  if (hj) hj[r] = h;
}

// Mica.cpp:341-361: average MI of every site with all the others, j ascending
__global__ void k5_average(int64_t S, const double* __restrict__ mi, double* __restrict__ avg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  double sum = 0.;
  for (int64_t j = 0; j < S; j++) {
    if (j == i) continue;
    const int64_t lo = j < i ? j : i, hi = j < i ? i : j;
    sum += mi[lo * S - lo * (lo + 1) / 2 + (hi - lo - 1)];
  }
  avg[i] = sum / (double)(S - 1);
}

// the other columns of a row of mica's table: i, j, Hmin = min entropy, Nmin = min norm (NaN without a mapping)
__global__ void k5_rows(int64_t S, const double* __restrict__ entropy, const double* __restrict__ norm, int32_t* __restrict__ oi,
                        int32_t* __restrict__ oj, double* __restrict__ hmin, double* __restrict__ nmin) {
  const int64_t i = blockIdx.y;
  const int64_t j = i + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  const int64_t idx = i * S - i * (i + 1) / 2 + (j - i - 1);
  oi[idx] = (int32_t)i; oj[idx] = (int32_t)j;
  const double a = entropy[i], b = entropy[j];
  hmin[idx] = a < b ? a : b;                       // std::min(entropy[i], entropy[j])
  if (nmin) {
    if (norm) { const double x = norm[i], y = norm[j]; nmin[idx] = x < y ? x : y; }
    else nmin[idx] = nan("");
  }
}

__global__ void k5_min2(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) out[r] = a[r] < b[r] ? a[r] : b[r];
}

} // namespace

void launch_mica_entropy(int A, int T, int64_t n, int64_t n_pad, const uint8_t* tips, const uint32_t* cmask, double* entropy,
                         cudaStream_t st) {
  const unsigned g = (unsigned)((n + 127) / 128);
  if (A == 4) k5_entropy<4><<<g, 128, 0, st>>>(T, n, n_pad, tips, cmask, entropy);
  else if (A == 20) k5_entropy<20><<<g, 128, 0, st>>>(T, n, n_pad, tips, cmask, entropy);
  else fail("mica kernels are built for A = 4 and A = 20 (got %d)", A);
  CMB_CUDA(cudaGetLastError());
}
void launch_mica_pairs(int A, int T, int64_t S, int64_t n_pad, const uint8_t* tips, const uint32_t* cmask, double* mi, double* hj,
                       cudaStream_t st) {
  if (S < 2) return;
  if (S > 65535) fail("mica: %lld sites exceed the 65535 rows one launch scores", (long long)S);
  dim3 grid((unsigned)((S - 1 + 127) / 128), (unsigned)(S - 1));
  if (A == 4) k5_pairs<4><<<grid, 128, 0, st>>>(T, S, n_pad, tips, cmask, mi, hj);
  else if (A == 20) k5_pairs<20><<<grid, 128, 0, st>>>(T, S, n_pad, tips, cmask, mi, hj);
  else fail("mica kernels are built for A = 4 and A = 20 (got %d)", A);
  CMB_CUDA(cudaGetLastError());
}
void launch_mica_listed(int A, int T, int64_t n, const uint8_t* t1, int64_t np1, const uint8_t* t2, int64_t np2, const int32_t* a,
                        const int32_t* b, const uint32_t* cmask, double* mi, double* hj, cudaStream_t st) {
  if (n == 0) return;
  const unsigned g = (unsigned)((n + 127) / 128);
  if (A == 4) k5_listed<4><<<g, 128, 0, st>>>(T, n, t1, np1, t2, np2, a, b, cmask, mi, hj);
  else if (A == 20) k5_listed<20><<<g, 128, 0, st>>>(T, n, t1, np1, t2, np2, a, b, cmask, mi, hj);
  else fail("mica kernels are built for A = 4 and A = 20 (got %d)", A);
  CMB_CUDA(cudaGetLastError());
}
void launch_mica_average(int64_t S, const double* mi, double* avg, cudaStream_t st) {
  k5_average<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(S, mi, avg);
  CMB_CUDA(cudaGetLastError());
}
void launch_mica_rows(int64_t S, const double* entropy, const double* norm, int32_t* oi, int32_t* oj, double* hmin, double* nmin,
                      cudaStream_t st) {
  if (S < 2) return;
  dim3 grid((unsigned)((S - 1 + 127) / 128), (unsigned)(S - 1));
  k5_rows<<<grid, 128, 0, st>>>(S, entropy, norm, oi, oj, hmin, nmin);
  CMB_CUDA(cudaGetLastError());
}
void launch_min2(int64_t n, const double* a, const double* b, double* out, cudaStream_t st) {
  if (n == 0) return;
  k5_min2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, a, b, out);
  CMB_CUDA(cudaGetLastError());
}

} // namespace cmb
