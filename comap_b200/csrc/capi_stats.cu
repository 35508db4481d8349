// C ABI, part 2: simulation, null distribution, pair statistics, clustering.
#include "../../include/comap_b200.h"
#include "context.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace cmb { extern thread_local std::string g_last_error; }
using namespace cmb;
struct cmb_ctx { Context c; };

#define CMB_TRY try {
#define CMB_CATCH                                   \
  }                                                 \
  catch (const std::exception& e) {                 \
    g_last_error = e.what();                        \
    return 1;                                       \
  }                                                 \
  catch (...) {                                     \
    g_last_error = "unknown error";                 \
    return 1;                                       \
  }                                                 \
  return 0;

namespace {

void check_stat(int stat_id) {
  if (stat_id < 0 || stat_id > 4) fail("unknown statistic id %d", stat_id);
}

// Bytes of device memory one simulated site needs through simulate -> map x2 -> paired.
size_t null_bytes_per_site(const Context& c) {
  size_t D = (size_t)c.tree.n_slots * c.C * c.A * 8;
  size_t out = (size_t)c.tree.B * 8 * 2;
  size_t tips = (size_t)c.tree.n_leaves * 2;
  return D + out + tips + (size_t)c.C * 8 + 16 * 8;
}

MapBuffers sim_buffers(Context& c, int k, int64_t n, int64_t n_pad) {
  const int A = c.A, C = c.C, B = c.tree.B, T = c.tree.n_leaves;
  c.s_tips[k].reserve((size_t)T * n_pad);
  c.s_D.reserve(sizeof(double) * (size_t)c.tree.n_slots * C * A * n_pad);
  c.s_Lc.reserve(sizeof(double) * (size_t)C * n_pad);
  c.s_invL.reserve(sizeof(double) * n_pad);
  c.s_loglik.reserve(sizeof(double) * n_pad);
  c.s_pr[k].reserve(sizeof(double) * n_pad);
  c.s_rc[k].reserve(sizeof(int32_t) * n_pad);
  c.s_out[k].reserve(sizeof(double) * (size_t)B * n_pad);
  c.s_sum[k].reserve(sizeof(double) * n_pad);
  c.s_sumsq[k].reserve(sizeof(double) * n_pad);
  MapBuffers b;
  b.n = n; b.n_pad = n_pad;
  b.tips = c.s_tips[k].as<uint8_t>();
  b.D = c.s_D.as<double>(); b.Lc = c.s_Lc.as<double>(); b.invL = c.s_invL.as<double>();
  b.loglik = c.s_loglik.as<double>(); b.post_rate = c.s_pr[k].as<double>(); b.rate_class = c.s_rc[k].as<int32_t>();
  b.out = c.s_out[k].as<double>(); b.sum = c.s_sum[k].as<double>(); b.sumsq = c.s_sumsq[k].as<double>();
  return b;
}

void null_load(Context& c, const double* stat_dev, const double* nmin_dev, int64_t n, int K, double nmax) {
  if (K < 1) fail("statistic.null.nb_rate_classes must be > 0 (Domain.cpp:49)");
  if (K > 4096) fail("too many null bins (%d)", K);
  if (nmax < 0.) {
    if (!c.mapped) fail("null binning with nmax < 0 needs a mapped alignment (cmb_map)");
    nmax = c.max_norm;
  }
  if (n > 0x7fffffff) fail("null distribution too large (%lld samples)", (long long)n);
  NullState& ns = c.null;
  ns.K = K; ns.nmax = nmax;
  ns.sorted.reserve(sizeof(double) * (size_t)std::max<int64_t>(n, 1));
  ns.bin_off_dev.reserve(sizeof(int64_t) * (K + 2));
  c.prof_begin("sort");
  int l = bin_and_sort(n, stat_dev, nmin_dev, K, nmax, c.scratch2, ns.sorted.as<double>(),
                       ns.bin_off_dev.as<int64_t>(), c.stream);
  c.prof_end(l);
  ns.bin_off.assign(K + 1, 0);
  CMB_CUDA(cudaMemcpyAsync(ns.bin_off.data(), ns.bin_off_dev.p, sizeof(int64_t) * (K + 1), cudaMemcpyDeviceToHost,
                           c.stream));
  CMB_CUDA(cudaStreamSynchronize(c.stream));
  ns.ready = true;
}

void null_core(Context& c, int stat_id, uint64_t seed, int rep_cpu, int rep_ram, int rep_begin, int rep_end,
               int weighted, int K, double nmax, double* raw, const uint8_t* sim1, const uint8_t* sim2) {
  check_stat(stat_id);
  if (rep_ram < 1 || rep_cpu < 0) fail("null: bad replicate counts");
  if (rep_begin < 0 || rep_end > rep_cpu || rep_begin > rep_end) fail("null: bad replicate range");
  CMB_CUDA(cudaSetDevice(c.device));
  c.ensure_streams();
  const int B = c.tree.B, T = c.tree.n_leaves;
  const int64_t R = rep_ram, nreps = rep_end - rep_begin, total = nreps * R;
  NullState& ns = c.null;
  ns.ready = false;
  ns.stat.reserve(sizeof(double) * (size_t)std::max<int64_t>(total, 1));
  ns.nmin.reserve(sizeof(double) * (size_t)std::max<int64_t>(total, 1));
  ns.n_samples = total;
  // batch as many outer replicates as fit comfortably in free HBM
  size_t freeb = 0, totalb = 0;
  CMB_CUDA(cudaMemGetInfo(&freeb, &totalb));
  size_t held = c.s_D.cap + c.s_out[0].cap + c.s_out[1].cap + c.s_tips[0].cap + c.s_tips[1].cap;
  size_t budget = (size_t)((freeb + held) * 0.6);
  int64_t max_sites = std::max<int64_t>(R, (int64_t)(budget / null_bytes_per_site(c)));
  max_sites = std::min<int64_t>(max_sites, (int64_t)1 << 20);
  int64_t rpb = std::max<int64_t>(1, max_sites / R);
  MapModel m = c.map_model();
  int64_t off = 0;
  for (int64_t r0 = rep_begin; r0 < rep_end; r0 += rpb) {
    const int64_t nb = std::min<int64_t>(rpb, rep_end - r0), n = nb * R, n_pad = pad_sites(n);
    MapBuffers b[2];
    for (int k = 0; k < 2; k++) {
      b[k] = sim_buffers(c, k, n, n_pad);
      if (sim1) {
        const uint8_t* src = k == 0 ? sim1 : sim2;
        for (int64_t r = 0; r < nb; r++)
          CMB_CUDA(cudaMemcpy2DAsync(c.s_tips[k].as<uint8_t>() + r * R, n_pad, src + (size_t)(r0 + r) * T * R, R, R, T,
                                     cudaMemcpyHostToDevice, c.stream));
      } else {
        c.prof_begin("simulate");
        launch_simulate(m, c.sim_stream, seed, (2 * r0 + k) * R, R, 2 * R, n, n_pad, weighted, c.tree.n_nodes - 1,
                        c.s_tips[k].as<uint8_t>(), nullptr, c.stream);
        c.prof_end(1);
      }
      c.run_map(b[k], true);
    }
    c.prof_begin("null_pairs");
    launch_paired(stat_id, B, n, n_pad, b[0].out, b[1].out, ns.stat.as<double>() + off, ns.nmin.as<double>() + off,
                  c.stream);
    c.prof_end(1);
    if (raw) {
      c.scratch.reserve(sizeof(double) * 4 * (size_t)n);
      launch_raw_rows(n, ns.stat.as<double>() + off, ns.nmin.as<double>() + off, b[0].rate_class, b[1].rate_class,
                      b[0].post_rate, b[1].post_rate, c.scratch.as<double>(), c.stream);
      c.prof.total_launches += 1;
      CMB_CUDA(cudaMemcpyAsync(raw + off * 4, c.scratch.p, sizeof(double) * 4 * (size_t)n, cudaMemcpyDeviceToHost,
                               c.stream));
      CMB_CUDA(cudaStreamSynchronize(c.stream));
    }
    off += n;
  }
  if (K > 0) null_load(c, ns.stat.as<double>(), ns.nmin.as<double>(), total, K, nmax);
  else CMB_CUDA(cudaStreamSynchronize(c.stream));
}

} // namespace

extern "C" {

int cmb_simulate(cmb_ctx* ctx, uint64_t seed, int64_t first_site, int64_t n, int32_t weighted_classes,
                 uint8_t* states, int32_t* classes) {
  CMB_TRY
  Context& c = ctx->c;
  CMB_CUDA(cudaSetDevice(c.device));
  c.ensure_streams();
  if (n < 1) fail("cmb_simulate: n must be positive");
  const int T = c.tree.n_leaves;
  const int64_t n_pad = pad_sites(n);
  c.s_tips[0].reserve((size_t)T * n_pad);
  c.s_cls.reserve(sizeof(int32_t) * n_pad);
  MapModel m = c.map_model();
  c.prof_begin("simulate");
  launch_simulate(m, c.sim_stream, seed, first_site, n, 0, n, n_pad, weighted_classes, c.tree.n_nodes - 1,
                  c.s_tips[0].as<uint8_t>(), c.s_cls.as<int32_t>(), c.stream);
  c.prof_end(1);
  CMB_CUDA(cudaMemcpy2DAsync(states, n, c.s_tips[0].p, n_pad, n, T, cudaMemcpyDeviceToHost, c.stream));
  if (classes) CMB_CUDA(cudaMemcpyAsync(classes, c.s_cls.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
  CMB_CUDA(cudaStreamSynchronize(c.stream));
  CMB_CATCH
}

int cmb_null_intra(cmb_ctx* ctx, int32_t stat_id, uint64_t seed, int32_t rep_cpu, int32_t rep_ram, int32_t rep_begin,
                   int32_t rep_end, int32_t weighted_classes, int32_t K, double nmax, double* raw) {
  CMB_TRY
  null_core(ctx->c, stat_id, seed, rep_cpu, rep_ram, rep_begin, rep_end, weighted_classes, K, nmax, raw, nullptr, nullptr);
  CMB_CATCH
}

int cmb_null_intra_from_alignments(cmb_ctx* ctx, int32_t stat_id, int32_t rep_cpu, int32_t rep_ram, const uint8_t* sim1,
                                   const uint8_t* sim2, int32_t K, double nmax, double* raw) {
  CMB_TRY
  if (!sim1 || !sim2) fail("cmb_null_intra_from_alignments: alignments missing");
  const int64_t tot = (int64_t)rep_cpu * ctx->c.tree.n_leaves * rep_ram;
  for (int64_t i = 0; i < tot; i++)
    if (sim1[i] >= ctx->c.A || sim2[i] >= ctx->c.A) fail("cmb_null_intra_from_alignments: state out of range");
  null_core(ctx->c, stat_id, 0, rep_cpu, rep_ram, 0, rep_cpu, 0, K, nmax, raw, sim1, sim2);
  CMB_CATCH
}

int cmb_null_samples_dev(cmb_ctx* ctx, const double** stat_dev, const double** nmin_dev, int64_t* n) {
  CMB_TRY
  *stat_dev = ctx->c.null.stat.as<double>();
  *nmin_dev = ctx->c.null.nmin.as<double>();
  *n = ctx->c.null.n_samples;
  CMB_CATCH
}

int cmb_null_load_dev(cmb_ctx* ctx, const double* stat_dev, const double* nmin_dev, int64_t n, int32_t K, double nmax) {
  CMB_TRY
  CMB_CUDA(cudaSetDevice(ctx->c.device));
  null_load(ctx->c, stat_dev, nmin_dev, n, K, nmax);
  CMB_CATCH
}

int cmb_null_get(cmb_ctx* ctx, int32_t* K, double* nmax, int64_t* bin_offsets, double* sorted, int64_t capacity) {
  CMB_TRY
  Context& c = ctx->c;
  if (!c.null.ready) fail("cmb_null_get: no null distribution loaded");
  if (K) *K = c.null.K;
  if (nmax) *nmax = c.null.nmax;
  if (bin_offsets) std::memcpy(bin_offsets, c.null.bin_off.data(), sizeof(int64_t) * (c.null.K + 1));
  if (sorted) {
    int64_t n = c.null.bin_off[c.null.K];
    if (capacity < n) fail("cmb_null_get: capacity %lld < %lld", (long long)capacity, (long long)n);
    CMB_CUDA(cudaMemcpyAsync(sorted, c.null.sorted.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
    CMB_CUDA(cudaStreamSynchronize(c.stream));
  }
  CMB_CATCH
}

// column ids: 0 i, 1 j, 2 stat, 3 rcmin, 4 prmin, 5 nmin, 6 pvalue, 7 nsim
static const size_t kColElt[8] = {4, 4, 8, 4, 8, 8, 8, 8};

int cmb_pairs_resident(cmb_ctx* ctx, int32_t stat_id, const cmb_filters* f, int32_t use_null, int32_t shard_index,
                       int32_t shard_count, uint32_t columns, int64_t* n_rows) {
  CMB_TRY
  Context& c = ctx->c;
  CMB_CUDA(cudaSetDevice(c.device));
  check_stat(stat_id);
  if (!c.mapped) fail("cmb_pairs: call cmb_map first");
  if (use_null && !c.null.ready) fail("cmb_pairs: no null distribution (cmb_null_intra / cmb_null_load_dev)");
  if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count) fail("cmb_pairs: bad shard");
  if (!use_null) columns &= ~(uint32_t)0xC0;
  const int64_t S = c.S;
  constexpr int TS = 64;
  // owned rows, dense offsets, tile list
  std::vector<int32_t> rows;
  std::vector<int64_t> row_off;
  int64_t total = 0;
  const int64_t period = 2 * (int64_t)shard_count;
  for (int64_t i = 0; i < S; i++) {
    int64_t r = i % period;
    if (shard_count == 1 || r == shard_index || r == period - 1 - shard_index) {
      rows.push_back((int32_t)i);
      row_off.push_back(total);
      total += S - 1 - i;
    }
  }
  std::vector<int2> tiles;
  const int64_t n_rows_owned = (int64_t)rows.size();
  for (int64_t ti = 0; ti * TS < n_rows_owned; ti++) {
    int64_t imin = rows[ti * TS];
    for (int64_t tj = (imin + 1) / TS; tj * TS < S; tj++) tiles.push_back(make_int2((int)ti, (int)tj));
  }
  const bool any_filter = f && (f->min_rate_class > 0 || f->min_rate > 0. || f->max_rate_class_diff >= 0 ||
                                f->max_rate_diff >= 0. || f->min_stat > 0.);
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t nT = std::max<size_t>(total, 1);
  size_t o_rows = 0, o_roff = al(o_rows + rows.size() * 4), o_tiles = al(o_roff + row_off.size() * 8),
         o_end = al(o_tiles + tiles.size() * 8);
  DevBuf& meta = c.scratch;
  meta.reserve(o_end + 256);
  unsigned char* mb = meta.as<unsigned char>();
  if (!rows.empty()) {
    CMB_CUDA(cudaMemcpyAsync(mb + o_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, c.stream));
    CMB_CUDA(cudaMemcpyAsync(mb + o_roff, row_off.data(), row_off.size() * 8, cudaMemcpyHostToDevice, c.stream));
  }
  if (!tiles.empty())
    CMB_CUDA(cudaMemcpyAsync(mb + o_tiles, tiles.data(), tiles.size() * 8, cudaMemcpyHostToDevice, c.stream));
  // dense columns and, when filtering, compacted copies
  size_t off[8] = {0}, off2[8] = {0}, cur = 0;
  for (int k = 0; k < 8; k++)
    if (columns >> k & 1) {
      off[k] = cur; cur = al(cur + kColElt[k] * nT);
      if (any_filter) { off2[k] = cur; cur = al(cur + kColElt[k] * nT); }
    }
  size_t o_keep = cur;
  cur = al(cur + nT);
  c.staging.reserve(cur + 256);
  unsigned char* sb = c.staging.as<unsigned char>();

  TilesLaunch L;
  L.stat_id = stat_id; L.B = c.tree.B; L.S = S; L.S_pad = c.S_pad; L.out = c.d_out.as<double>();
  L.mean = c.pairs_mean.as<double>(); L.sd = c.pairs_sd.as<double>(); L.norm = c.pairs_norm.as<double>();
  L.post_rate = c.d_pr.as<double>(); L.rate_class = c.d_rc.as<int32_t>();
  L.tiles = (const int2*)(mb + o_tiles); L.n_tiles = (int64_t)tiles.size();
  L.rows = shard_count == 1 ? nullptr : (const int32_t*)(mb + o_rows);
  L.n_rows = n_rows_owned; L.row_off = (const int64_t*)(mb + o_roff);
  if (f) {
    L.min_rate_class = f->min_rate_class; L.max_rate_class_diff = f->max_rate_class_diff;
    L.min_rate = f->min_rate; L.max_rate_diff = f->max_rate_diff; L.min_stat = f->min_stat;
  }
  L.any_filter = any_filter;
  if (use_null) {
    L.K = c.null.K; L.nmax = c.null.nmax;
    L.bin_off = c.null.bin_off_dev.as<int64_t>(); L.sorted = c.null.sorted.as<double>();
  }
  auto dp = [&](int k) -> void* { return (columns >> k & 1) ? sb + off[k] : nullptr; };
  L.o_i = (int32_t*)dp(0); L.o_j = (int32_t*)dp(1); L.o_stat = (double*)dp(2); L.o_rcmin = (int32_t*)dp(3);
  L.o_prmin = (double*)dp(4); L.o_nmin = (double*)dp(5); L.o_pvalue = (double*)dp(6); L.o_nsim = (int64_t*)dp(7);
  L.o_keep = sb + o_keep;
  c.prof_begin("pairs");
  int nl = launch_tiles(L, c.stream);
  c.prof_end(nl);
  int64_t kept = total;
  if (any_filter && total > 0) {
    int64_t* pos = nullptr;
    kept = compact_positions(total, L.o_keep, c.scratch2, &pos, c.stream);
    for (int k = 0; k < 8; k++) {
      if (!(columns >> k & 1)) continue;
      if (kColElt[k] == 4) compact_column<int32_t>(total, L.o_keep, pos, (int32_t*)(sb + off[k]), (int32_t*)(sb + off2[k]), c.stream);
      else if (k == 7) compact_column<int64_t>(total, L.o_keep, pos, (int64_t*)(sb + off[k]), (int64_t*)(sb + off2[k]), c.stream);
      else compact_column<double>(total, L.o_keep, pos, (double*)(sb + off[k]), (double*)(sb + off2[k]), c.stream);
      off[k] = off2[k];
      c.prof.total_launches += 1;
    }
  }
  for (int k = 0; k < 8; k++) c.pairs_col_off[k] = (columns >> k & 1) ? (int64_t)off[k] : -1;
  c.pairs_rows = kept;
  if (n_rows) *n_rows = kept;
  CMB_CATCH
}

int cmb_pairs_fetch(cmb_ctx* ctx, int32_t column, void* host, int64_t capacity) {
  CMB_TRY
  Context& c = ctx->c;
  if (column < 0 || column > 7) fail("cmb_pairs_fetch: bad column %d", column);
  if (c.pairs_rows < 0 || c.pairs_col_off[column] < 0) fail("cmb_pairs_fetch: column %d is not resident", column);
  if (capacity < c.pairs_rows) fail("cmb_pairs_fetch: capacity %lld < %lld rows", (long long)capacity, (long long)c.pairs_rows);
  if (c.pairs_rows > 0)
    CMB_CUDA(cudaMemcpyAsync(host, c.staging.as<unsigned char>() + c.pairs_col_off[column],
                             kColElt[column] * (size_t)c.pairs_rows, cudaMemcpyDeviceToHost, c.stream));
  CMB_CATCH
}

int cmb_pairs(cmb_ctx* ctx, int32_t stat_id, const cmb_filters* f, int32_t use_null, int32_t shard_index,
              int32_t shard_count, int64_t capacity, int32_t* out_i, int32_t* out_j, double* out_stat, int32_t* out_rcmin,
              double* out_prmin, double* out_nmin, double* out_pvalue, int64_t* out_nsim, int64_t* n_rows) {
  void* host[8] = {out_i, out_j, out_stat, out_rcmin, out_prmin, out_nmin, out_pvalue, out_nsim};
  uint32_t columns = 0;
  for (int k = 0; k < 8; k++) if (host[k]) columns |= 1u << k;
  int64_t kept = 0;
  int rc = cmb_pairs_resident(ctx, stat_id, f, use_null, shard_index, shard_count, columns, &kept);
  if (rc) return rc;
  CMB_TRY
  if (kept > capacity) fail("cmb_pairs: capacity %lld < %lld rows", (long long)capacity, (long long)kept);
  if (!use_null) columns &= ~(uint32_t)0xC0;
  for (int k = 0; k < 8; k++)
    if (columns >> k & 1) { if (cmb_pairs_fetch(ctx, k, host[k], capacity)) return 1; }
  CMB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  if (n_rows) *n_rows = kept;
  CMB_CATCH
}

#define CMB_TODO(name) CMB_TRY fail(name ": not implemented yet"); CMB_CATCH
int cmb_distance_matrix(cmb_ctx*, int32_t, double*) { CMB_TODO("cmb_distance_matrix") }
int cmb_cluster(cmb_ctx*, int32_t, int32_t*, int32_t*, double*) { CMB_TODO("cmb_cluster") }
int cmb_groups(cmb_ctx*, int32_t, int32_t, int32_t*, int64_t*, double*, double*, double*, int64_t*) { CMB_TODO("cmb_groups") }
int cmb_cluster_null(cmb_ctx*, int32_t, int32_t, uint64_t, int32_t, int32_t, int32_t, int32_t, int64_t, int64_t, int32_t*, int32_t*, double*, double*, double*, int32_t*, int64_t*, int64_t*) { CMB_TODO("cmb_cluster_null") }

} // extern "C"
