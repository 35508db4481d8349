/*
 * comap_oracle.h -- CPU restatement of CoMap's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle: a single-threaded, fp64, plain-C restatement of the
 * algorithms on the path BASELINE.json's north_star names.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product (comap_b200/) never links, imports or calls anything in oracle/.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - likelihood + mapping: PINNED against the reference's own golden outputs
 *     examples/Proteins/Benchmark/CoMap/Myo_unif.vec, Myo_decomp.vec, Myo.infos
 *     (tests/golden/myoglobin.npz, tests/test_oracle_golden.py).
 *   - pair statistics / null / clustering: the reference ships no expected outputs
 *     ("parity unpinned" by data); they are restated from the in-tree sources cited on
 *     each function and cross-checked with numpy/scipy in tests/.
 *
 * The Bio++ (bpp-phyl/bpp-core >= 3.0.0, CMakeLists.txt:113) routines CoMap calls are
 * not under /root/reference; their published algorithms are restated here and anchored
 * on CoMap's call sites (cited per function) and on the goldens above.
 *
 * Conventions shared with include/comap_b200.h:
 *   tree   : n_nodes nodes, ids in Newick post-order (children before parent),
 *            root = n_nodes-1, parent[root] = -1; branch b = edge above node b,
 *            B = n_nodes-1.  Leaves are the nodes without children; leaf k (k-th leaf
 *            in id order) is row k of every alignment.
 *   model  : A states, generator Q (row-major A*A, reversible w.r.t. pi), C rate
 *            classes (rates[c], probs[c]).
 *   codes  : alignments are uint8 codes, tip-major [T][S]; code_mask[code] is the
 *            bitmask of compatible states (ambiguity-aware tips).
 */
#ifndef COMAP_ORACLE_H
#define COMAP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_COUNT_UNIFORMIZATION = 0, ORC_COUNT_DECOMPOSITION = 1, ORC_COUNT_NAIVE = 2,
       ORC_COUNT_LAPLACE = 3 /* | trunc << 8; trunc 0 = 10 */, ORC_COUNT_LABEL = 4, ORC_COUNT_ONE_JUMP = 5 };
enum {
  ORC_STAT_CORRELATION = 0,
  ORC_STAT_COVARIANCE = 1,
  ORC_STAT_COSINUS = 2,
  ORC_STAT_COSUBSTITUTION = 3,
  ORC_STAT_COMPENSATION = 4,
  ORC_STAT_CORRECTED_CORRELATION = 5, /* needs orc_set_mean_vector */
  ORC_STAT_MI = 6, /* MI(threshold=..) without nijt=Label; threshold from orc_set_mi_threshold (0.99) */
  ORC_STAT_MI_LABEL = 7 /* MI with nijt=Label: one category per label; alphabet size from orc_set_mi_label */
};
enum { ORC_DIST_CORRELATION = 0, ORC_DIST_COMPENSATION = 1, ORC_DIST_EUCLIDIAN = 2 };
enum { ORC_LINK_COMPLETE = 0, ORC_LINK_SINGLE = 1, ORC_LINK_AVERAGE = 2 };

const char* orc_last_error(void);

/* P(t) = exp(Q t) by eigen-decomposition of the symmetrised generator. */
int orc_pmatrix(int A, const double* Q, const double* pi, double t, double* P);

/* E[#substitutions | x -> y, t], A*A, Total register (1 type), optional weights. */
int orc_counts(int method, int A, const double* Q, const double* pi, const double* weights,
               double t, double* N);

/* Likelihood + substitution mapping for S sites.  Outputs may be NULL. */
int orc_map(int n_nodes, const int32_t* parent, const double* brlen,
            int A, const double* Q, const double* pi,
            int C, const double* rates, const double* probs,
            int method, const double* weights,
            int64_t S, const uint8_t* codes, int n_codes, const uint32_t* code_mask,
            double* n_out /* [S][B] */, double* norm, double* post_rate,
            int32_t* rate_class, double* loglik);

/* One pair statistic on two branch vectors of length B. */
double orc_stat(int stat_id, int B, const double* v1, const double* v2);
/* CorrectedCorrelationStatistic (Statistics.h:176-205): the mean vector the statistic subtracts from
 * both sites, and how CoMap builds it from the mapping (CoMap.cpp:350-359: running sum over
 * sites in site order, divided by the number of sites). */
void orc_set_mean_vector(int B, const double* mv);
void orc_set_mi_threshold(double threshold);
void orc_set_mi_label(int n_states);
/* nijt.average / nijt.joint for every later mapping (orc_map, orc_null_intra, ...); default 1, 1 */
void orc_set_map_mode(int average, int joint);
/* asr.method = marginal: states [n_nodes][S] */
int orc_ancestral_states(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                         const double* pi, int C, const double* rates, const double* probs, int64_t S,
                         const uint8_t* codes, int n_codes, const uint32_t* code_mask, uint8_t* states);
void orc_set_mean_vectors(int B, const double* mv1, const double* mv2);
double orc_stat2(int stat_id, int B, const double* v1, const double* v2);
/* CoETools::computeInterStats, CoETools.cpp:732-840 */
int orc_pairs_inter(int stat_id, int64_t S1, int64_t S2, int B, const double* n1, const double* n2,
                    const double* norm1, const double* norm2, const double* pr1, const double* pr2,
                    const int32_t* rc1, const int32_t* rc2, int min_rate_class1, int min_rate_class2,
                    double min_rate1, double min_rate2, int max_rate_class_diff, double max_rate_diff,
                    double min_stat, int independent, int nmin_by_row, int64_t capacity, int32_t* out_i,
                    int32_t* out_j, double* out_stat, int32_t* out_rcmin, double* out_prmin,
                    double* out_nmin, int64_t* n_rows);
void orc_mean_vector(int64_t S, int B, const double* n, double* mv);
/* Group statistic (min over pairs; closed form for Compensation). idx = site rows. */
double orc_stat_group(int stat_id, int B, const double* n /* [S][B] */, int n_members,
                      const int32_t* members);

/* Equal-width Domain(0, nmax, K).getIndex(x); returns -1 when out of range. */
int orc_domain_index(double lo, double hi, int K, double x);

/* All pairs i<j with filters and p-values.  K = 0 disables the null columns. */
int orc_pairs(int stat_id, int64_t S, int B, const double* n, const double* norm,
              const double* post_rate, const int32_t* rate_class,
              int min_rate_class, double min_rate, int max_rate_class_diff,
              double max_rate_diff, double min_stat,
              int K, double nmax, const int64_t* bin_offsets, const double* sorted_null,
              int64_t capacity, int32_t* out_i, int32_t* out_j, double* out_stat,
              int32_t* out_rcmin, double* out_prmin, double* out_nmin,
              double* out_pvalue, int64_t* out_nsim, int64_t* n_rows);

/* Forward simulation of n sites (global site ids first_site .. first_site+n-1) with the
 * counter-based Philox4x32-10 generator shared with the device simulator. */
int orc_simulate(int n_nodes, const int32_t* parent, const double* brlen,
                 int A, const double* Q, const double* pi,
                 int C, const double* rates, const double* probs,
                 uint64_t seed, int64_t first_site, int64_t n, int weighted_classes,
                 uint8_t* states /* [T][n] */, int32_t* classes /* [n], nullable */);

/* Null distribution from given simulated alignments: sim1/sim2 are
 * [rep_cpu][T][rep_ram] state codes (0..A-1).  raw is [rep_cpu*rep_ram][4] =
 * (Stat, RCmin, PRmin, Nmin); sorted/bin_offsets receive the per-bin ascending
 * samples (K bins over [0, nmax)).  Any output may be NULL. */
int orc_null_intra(int n_nodes, const int32_t* parent, const double* brlen,
                   int A, const double* Q, const double* pi,
                   int C, const double* rates, const double* probs,
                   int method, const double* weights, int stat_id,
                   int rep_cpu, int rep_ram, const uint8_t* sim1, const uint8_t* sim2,
                   int K, double nmax, double* raw, int64_t* bin_offsets, double* sorted);

/* Distance matrix (full symmetric S*S, zero diagonal). */
int orc_distance_matrix(int dist_id, int64_t S, int B, const double* n, double* mat);

/* Agglomerative clustering.  mat is S*S and is overwritten.  Output dendrogram:
 * nodes 0..S-1 are leaves, S..2S-2 inner nodes in creation order (root last);
 * left/right/height are indexed by inner node - S; height = merge distance / 2. */
int orc_hclust(int linkage, int64_t S, double* mat, int32_t* left, int32_t* right,
               double* height);

/* Groups of the dendrogram in the reference's emission order (post-order of inner
 * nodes), keeping only groups with at most max_size members (CoMap.cpp:517).  members is
 * a flat list (capacity (S-1)*max_size), offsets has n_groups+1 entries (capacity S).
 * stat uses dist_id semantics: comp - 2*height for CORRELATION/EUCLIDIAN(2*height),
 * closed-form group compensation for COMPENSATION. */
int orc_groups(int dist_id, int64_t S, int B, const double* n, const double* norm,
               const int32_t* left, const int32_t* right, const double* height, int max_size,
               int32_t* members, int64_t* offsets, double* g_height, double* g_stat,
               double* g_nmin, int64_t* n_groups);

#ifdef __cplusplus
}
#endif
/* simulations.continuous = yes: one continuous rate per site (kind 1 constant, 2 gamma, 3 invariant + gamma) */
int orc_simulate_continuous(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                            const double* pi, int kind, double alpha, double p_inv, uint64_t seed,
                            int64_t first_site, int64_t n, uint8_t* states, double* rates_out);

/* Mica (CoMap/Mica.cpp): SiteTools::entropy / mutualInformation / jointEntropy with resolveUnknowns = true [Bio++ bpp-seq,
 * from memory]; columns are T codes, code_mask as everywhere else */
double orc_site_entropy(int T, const uint8_t* col, int A, int n_codes, const uint32_t* code_mask);
void orc_site_pair(int T, const uint8_t* c1, const uint8_t* c2, int A, int n_codes, const uint32_t* code_mask,
                   double* mi_out, double* hjoint_out);
void orc_mica_sites(int64_t S, int T, const uint8_t* codes, int A, int n_codes, const uint32_t* code_mask,
                    double* entropy, double* average_mi);

/* null.method = permutations: miTest (Mica.cpp:92-118) of one pair / of every pair i < j in mica's order */
void orc_mica_permutation_test(int T, const uint8_t* c1, const uint8_t* c2, int A, int n_codes, const uint32_t* code_mask,
                               uint64_t seed, uint64_t pair, int max_perm, double* mi_out, double* pvalue, int* nperm,
                               double* closest);
void orc_mica_permutations(int64_t S, int T, const uint8_t* codes, int A, int n_codes, const uint32_t* code_mask, uint64_t seed,
                           int max_perm, double* pvalue, int32_t* nperm, double* closest);

#endif
