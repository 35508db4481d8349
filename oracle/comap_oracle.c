/*
 * comap_oracle.c -- CPU restatement of CoMap's hot path.  TEST INFRASTRUCTURE ONLY.
 * See comap_oracle.h for the rules on who may use this file and for pinning status.
 *
 * Reference citations are relative to /root/reference (jydu/comap 1.6.0a).  "[Bio++]"
 * marks arithmetic that lives in the un-vendored bpp-phyl/bpp-core >= 3.0.0 and is
 * restated from its published algorithm; the CoMap call site is cited instead.
 *
 * Single thread, fp64, straight loops (build: gcc -O2 -ffp-contract=off).
 */
#include "comap_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[512] = "";
const char* orc_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return -1; } while (0)

/* ------------------------------------------------------------------------------------ */
/* small dense helpers                                                                   */
/* ------------------------------------------------------------------------------------ */

static void mat_mul(int n, const double* a, const double* b, double* c) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.;
      for (int k = 0; k < n; k++) s += a[i * n + k] * b[k * n + j];
      c[i * n + j] = s;
    }
}

/* Cyclic Jacobi for a symmetric n*n matrix: a = v diag(w) v^T. */
static void jacobi_eigh(int n, double* a, double* w, double* v) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) v[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += a[p * n + q] * a[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (a[q * n + q] - a[p * n + p]) / (2. * apq);
        double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));
        double c = 1. / sqrt(t * t + 1.), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = a[i * n + i];
}

/* Spectral form of a reversible generator: Q = R diag(ev) L with L = R^{-1}.
 * [Bio++] AbstractReversibleSubstitutionModel::updateMatrices computes the same
 * right/left eigenvectors with a general eigen-solver; for reversible Q the
 * symmetrised form sqrt(pi_i) Q_ij / sqrt(pi_j) gives them with an orthogonal basis. */
typedef struct {
  int A;
  double *ev, *R, *L; /* R[x][k], L[k][y] */
} spectral_t;

static int spectral_init(spectral_t* sp, int A, const double* Q, const double* pi) {
  sp->A = A;
  sp->ev = malloc(sizeof(double) * A);
  sp->R = malloc(sizeof(double) * A * A);
  sp->L = malloc(sizeof(double) * A * A);
  double* M = malloc(sizeof(double) * A * A);
  double* U = malloc(sizeof(double) * A * A);
  for (int i = 0; i < A; i++) {
    if (!(pi[i] > 0.)) {
      free(M); free(U);
      FAIL("spectral_init: non-positive equilibrium frequency %d", i);
    }
  }
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++) {
      double mij = Q[i * A + j] * sqrt(pi[i]) / sqrt(pi[j]);
      double mji = Q[j * A + i] * sqrt(pi[j]) / sqrt(pi[i]);
      if (fabs(mij - mji) > 1e-8 * (fabs(mij) + fabs(mji) + 1e-300) + 1e-12) {
        free(M); free(U);
        FAIL("spectral_init: generator is not reversible w.r.t. pi (%d,%d)", i, j);
      }
      M[i * A + j] = 0.5 * (mij + mji);
    }
  jacobi_eigh(A, M, sp->ev, U);
  for (int x = 0; x < A; x++)
    for (int k = 0; k < A; k++) {
      sp->R[x * A + k] = U[x * A + k] / sqrt(pi[x]);
      sp->L[k * A + x] = U[x * A + k] * sqrt(pi[x]);
    }
  free(M); free(U);
  return 0;
}
static void spectral_free(spectral_t* sp) { free(sp->ev); free(sp->R); free(sp->L); }

/* [Bio++] AbstractSubstitutionModel::getPij_t: P = R diag(exp(ev t)) L. */
static void spectral_pmatrix(const spectral_t* sp, double t, double* P) {
  int A = sp->A;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      double s = 0.;
      for (int k = 0; k < A; k++) s += sp->R[x * A + k] * exp(sp->ev[k] * t) * sp->L[k * A + y];
      P[x * A + y] = s;
    }
}

int orc_pmatrix(int A, const double* Q, const double* pi, double t, double* P) {
  spectral_t sp;
  if (spectral_init(&sp, A, Q, pi)) return -1;
  spectral_pmatrix(&sp, t, P);
  spectral_free(&sp);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* substitution counts  (CoMap.cpp:152 selects the method via `nijt=`)                   */
/* ------------------------------------------------------------------------------------ */

/* [Bio++] UniformizationSubstitutionCount::computeCounts_:
 *   mu = max_i |Q_ii|, R = I + Q/mu, Bm(j,k) = Q(j,k) w(j,k) for j != k,
 *   nMax = ceil(4 + 6 sqrt(lam) + lam), lam = mu t,
 *   s_0 = Bm, s_l = s_{l-1} R + R^l Bm,
 *   counts = sum_l s_l exp((l+1) ln lam - lam - ln mu - ln (l+1)!),
 *   counts(j,k) /= P(j,k;t); NaN/Inf -> 0; negatives -> 0 only when unweighted. */
static void counts_uniformization(const spectral_t* sp, const double* Q, const double* weights,
                                  double t, double* N) {
  int A = sp->A, AA = A * A;
  double mu = 0.;
  for (int i = 0; i < A; i++)
    if (fabs(Q[i * A + i]) > mu) mu = fabs(Q[i * A + i]);
  double lam = mu * t;
  for (int i = 0; i < AA; i++) N[i] = 0.;
  if (!(lam > 0.)) return; /* exp(-inf) weights: every term is 0, 0/P -> 0 */
  double* R = malloc(sizeof(double) * AA);
  double* Bm = malloc(sizeof(double) * AA);
  double* Rp = malloc(sizeof(double) * AA);
  double* s = malloc(sizeof(double) * AA);
  double* t1 = malloc(sizeof(double) * AA);
  double* t2 = malloc(sizeof(double) * AA);
  double* P = malloc(sizeof(double) * AA);
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++) {
      R[i * A + j] = Q[i * A + j] / mu + (i == j ? 1. : 0.);
      Bm[i * A + j] = (i == j) ? 0. : Q[i * A + j] * (weights ? weights[i * A + j] : 1.);
      Rp[i * A + j] = (i == j);
      s[i * A + j] = Bm[i * A + j];
    }
  long nmax = (long)ceil(4. + 6. * sqrt(lam) + lam);
  double loglam = log(lam), logmu = log(mu);
  for (long l = 0; l <= nmax; l++) {
    if (l > 0) {
      mat_mul(A, s, R, t1);   /* s_{l-1} R */
      mat_mul(A, Rp, R, t2);  /* R^l */
      memcpy(Rp, t2, sizeof(double) * AA);
      mat_mul(A, Rp, Bm, t2); /* R^l Bm */
      for (int i = 0; i < AA; i++) s[i] = t1[i] + t2[i];
    }
    double f = exp((double)(l + 1) * loglam - lam - logmu - lgamma((double)(l + 2)));
    for (int i = 0; i < AA; i++) N[i] += s[i] * f;
  }
  spectral_pmatrix(sp, t, P);
  for (int i = 0; i < AA; i++) {
    double v = N[i] / P[i];
    if (isnan(v) || isinf(v) || (!weights && v < 0.)) v = 0.;
    N[i] = v;
  }
  free(R); free(Bm); free(Rp); free(s); free(t1); free(t2); free(P);
}

/* [Bio++] DecompositionSubstitutionCount::computeCounts_: with Q = R diag(ev) L,
 *   counts = R [ (L Bm R) o J(t) ] L,  J_ij = t e^{ev_i t} if ev_i == ev_j
 *   else (e^{ev_i t} - e^{ev_j t}) / (ev_i - ev_j); then the same /P and clean-up. */
static void counts_decomposition(const spectral_t* sp, const double* Q, const double* weights,
                                 double t, double* N) {
  int A = sp->A, AA = A * A;
  double* Bm = calloc(AA, sizeof(double));
  double* t1 = malloc(sizeof(double) * AA);
  double* t2 = malloc(sizeof(double) * AA);
  double* P = malloc(sizeof(double) * AA);
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++)
      Bm[i * A + j] = (i == j) ? 0. : Q[i * A + j] * (weights ? weights[i * A + j] : 1.);
  mat_mul(A, sp->L, Bm, t1);
  mat_mul(A, t1, sp->R, t2); /* L Bm R */
  for (int i = 0; i < A; i++)
    for (int j = 0; j < A; j++) {
      double dd = sp->ev[i] - sp->ev[j], J;
      if (dd == 0.) J = t * exp(sp->ev[i] * t);
      else J = (exp(sp->ev[i] * t) - exp(sp->ev[j] * t)) / dd;
      t2[i * A + j] *= J;
    }
  mat_mul(A, sp->R, t2, t1);
  mat_mul(A, t1, sp->L, N);
  spectral_pmatrix(sp, t, P);
  for (int i = 0; i < AA; i++) {
    double v = N[i] / P[i];
    if (isnan(v) || isinf(v) || (!weights && v < 0.)) v = 0.;
    N[i] = v;
  }
  free(Bm); free(t1); free(t2); free(P);
}

/* nijt=Naive [Bio++ NaiveSubstitutionCount; pinned by Myo_naive.vec / Myo_naive_grantham.vec]: one
 * substitution (or its weight) whenever the two ends of the branch differ, whatever its length. */
static void counts_naive(const spectral_t* sp, const double* weights, double* N) {
  const int A = sp->A;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) N[x * A + y] = x == y ? 0. : (weights ? weights[x * A + y] : 1.);
}

/* [Bio++] MatrixTools::pow as the library that wrote Myo_laplace.vec evaluated it: 0, 1 and 2 directly,
 * above that by halving -- odd p: pow(p/2) squared; even p: pow((p-1)/2) squared, times A.  That is NOT
 * A^p from p = 3 on (it yields A^2, A^3, A^4, A^5, A^4, A^5, A^6 for p = 3..9); the golden is only
 * reproduced with it (exact powers: off by up to 0.68; these: 1.3e-5 = printed precision). */
static void mat_pow_bpp(int A, const double* M, int p, double* O) {
  int AA = A * A;
  if (p == 0) { for (int i = 0; i < AA; i++) O[i] = (i / A == i % A); return; }
  if (p == 1) { memcpy(O, M, sizeof(double) * AA); return; }
  if (p == 2) { mat_mul(A, M, M, O); return; }
  double* tmp = malloc(sizeof(double) * AA);
  if (p % 2) {
    mat_pow_bpp(A, M, p / 2, tmp);
    mat_mul(A, tmp, tmp, O);
  } else {
    double* sq = malloc(sizeof(double) * AA);
    mat_pow_bpp(A, M, (p - 1) / 2, tmp);
    mat_mul(A, tmp, tmp, sq);
    mat_mul(A, M, sq, O);
    free(sq);
  }
  free(tmp);
}

/* nijt=Laplace(trunc=k) [Bio++ LaplaceSubstitutionCount::computeCounts; pinned by
 * examples/Proteins/Benchmark/CoMap/Myo_laplace.vec, analyse.sh:12-15]:
 *   m = sum_{n=1}^{k-1} t^n / n!  sum_{p=0}^{n-1} pow(Q,p) QL pow(Q,n-p-1),  QL = Q with a zero diagonal,
 *   counts = m / P(t), entry by entry, no clean-up (zeroing negatives breaks the golden). */
static void counts_laplace(const spectral_t* sp, const double* Q, int trunc, double t, double* N) {
  int A = sp->A, AA = A * A;
  double* QL = malloc(sizeof(double) * AA);
  double* M2 = malloc(sizeof(double) * AA);
  double* M3 = malloc(sizeof(double) * AA);
  double* M4 = malloc(sizeof(double) * AA);
  double* M5 = malloc(sizeof(double) * AA);
  double* P = malloc(sizeof(double) * AA);
  for (int i = 0; i < AA; i++) { QL[i] = (i / A == i % A) ? 0. : Q[i]; N[i] = 0.; }
  double fact = 1.;
  for (int n = 1; n < trunc; n++) {
    fact *= n;
    for (int i = 0; i < AA; i++) M2[i] = 0.;
    for (int p = 0; p < n; p++) {
      mat_pow_bpp(A, Q, p, M3);
      mat_mul(A, M3, QL, M4);
      mat_pow_bpp(A, Q, n - p - 1, M3);
      mat_mul(A, M4, M3, M5);
      for (int i = 0; i < AA; i++) M2[i] += M5[i];
    }
    double f = pow(t, (double)n) / fact;
    for (int i = 0; i < AA; i++) N[i] += M2[i] * f;
  }
  spectral_pmatrix(sp, t, P);
  for (int i = 0; i < AA; i++) N[i] /= P[i];
  free(QL); free(M2); free(M3); free(M4); free(M5); free(P);
}

/* nijt=Label [Bio++ LabelSubstitutionCount, from memory; no golden]: every substitution x -> y carries its own
 * label, 1 .. A(A-1) in row-major order of the off-diagonal, whatever the branch length; 0 on the diagonal.
 * Used with nijt.average=no and statistic=MI (CoETools.cpp:577-589: bounds -0.5, 0.5, ... around the labels). */
static void counts_label(const spectral_t* sp, double* N) {
  const int A = sp->A;
  int count = 0;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) N[x * A + y] = x == y ? 0. : (double)++count;
}

/* nijt=ProbOneJump [Bio++ OneJumpSubstitutionCount, from memory; no golden]: probability that at least one
 * substitution happened on a branch of length t given its two ends: 1 if they differ, else 1 - exp(Q_xx t) / P_xx(t). */
static void counts_one_jump(const spectral_t* sp, const double* Q, double t, double* N) {
  const int A = sp->A;
  double* P = malloc(sizeof(double) * A * A);
  spectral_pmatrix(sp, t, P);
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) N[x * A + y] = x == y ? 1. - exp(Q[x * A + x] * t) / P[x * A + x] : 1.;
  free(P);
}

static void counts_any(int method, const spectral_t* sp, const double* Q, const double* weights,
                       double t, double* N) {
  if ((method & 0xff) == ORC_COUNT_LAPLACE) counts_laplace(sp, Q, (method >> 8) ? (method >> 8) : 10, t, N);
  else if (method == ORC_COUNT_LABEL) counts_label(sp, N);
  else if (method == ORC_COUNT_ONE_JUMP) counts_one_jump(sp, Q, t, N);
  else if (method == ORC_COUNT_NAIVE) counts_naive(sp, weights, N);
  else if (method == ORC_COUNT_DECOMPOSITION) counts_decomposition(sp, Q, weights, t, N);
  else counts_uniformization(sp, Q, weights, t, N);
}

int orc_counts(int method, int A, const double* Q, const double* pi, const double* weights,
               double t, double* N) {
  spectral_t sp;
  if ((method & 0xff) < ORC_COUNT_UNIFORMIZATION || (method & 0xff) > ORC_COUNT_ONE_JUMP) FAIL("orc_counts: unknown method %d", method);
  if (spectral_init(&sp, A, Q, pi)) return -1;
  counts_any(method, &sp, Q, weights, t, N);
  spectral_free(&sp);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* tree                                                                                  */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  int n, root, T;
  const int32_t* parent;
  double* len;          /* branch lengths with the [Bio++] 1e-6 lower bound applied */
  int *nch, *ch_off, *ch; /* children lists in id (= Newick) order */
  int* leaf_row;        /* node -> alignment row or -1 */
} tree_t;

static int tree_init(tree_t* t, int n_nodes, const int32_t* parent, const double* brlen) {
  memset(t, 0, sizeof *t);
  if (n_nodes < 3) FAIL("tree: need at least 3 nodes");
  t->n = n_nodes; t->root = n_nodes - 1; t->parent = parent;
  if (parent[t->root] != -1) FAIL("tree: last node must be the root (parent -1)");
  t->len = malloc(sizeof(double) * n_nodes);
  t->nch = calloc(n_nodes, sizeof(int));
  t->ch_off = malloc(sizeof(int) * (n_nodes + 1));
  t->ch = malloc(sizeof(int) * n_nodes);
  t->leaf_row = malloc(sizeof(int) * n_nodes);
  for (int v = 0; v < n_nodes - 1; v++) {
    if (parent[v] <= v || parent[v] >= n_nodes) FAIL("tree: ids must be post-order (node %d)", v);
    t->nch[parent[v]]++;
    /* [Bio++] branch-length lower bound 1e-6 */
    t->len[v] = brlen[v] < 1e-6 ? 1e-6 : brlen[v];
  }
  t->len[t->root] = 0.;
  t->ch_off[0] = 0;
  for (int v = 0; v < n_nodes; v++) t->ch_off[v + 1] = t->ch_off[v] + t->nch[v];
  int* fill = calloc(n_nodes, sizeof(int));
  for (int v = 0; v < n_nodes - 1; v++) {
    int p = parent[v];
    t->ch[t->ch_off[p] + fill[p]++] = v;
  }
  free(fill);
  t->T = 0;
  for (int v = 0; v < n_nodes; v++) t->leaf_row[v] = (t->nch[v] == 0) ? t->T++ : -1;
  if (t->nch[t->root] < 2) FAIL("tree: root must have at least 2 children");
  return 0;
}
static void tree_free(tree_t* t) {
  free(t->len); free(t->nch); free(t->ch_off); free(t->ch); free(t->leaf_row);
}

/* ------------------------------------------------------------------------------------ */
/* likelihood + mapping                                                                  */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  int A, C;
  const double *Q, *pi, *rates, *probs;
  spectral_t sp;
  double* P; /* [node][c][x][y]  = exp(Q len rate_c) */
  double* N; /* [node][c][x][y]  = conditional expected counts */
} tables_t;

static int tables_init(tables_t* tb, const tree_t* tr, int A, const double* Q, const double* pi,
                       int C, const double* rates, const double* probs, int method,
                       const double* weights, int want_counts) {
  memset(tb, 0, sizeof *tb);
  tb->A = A; tb->C = C; tb->Q = Q; tb->pi = pi; tb->rates = rates; tb->probs = probs;
  if (A < 2 || A > 32) FAIL("model: A must be in 2..32");
  if (C < 1 || C > 32) FAIL("model: C must be in 1..32");
  if (spectral_init(&tb->sp, A, Q, pi)) return -1;
  size_t AA = (size_t)A * A;
  tb->P = malloc(sizeof(double) * tr->n * C * AA);
  tb->N = want_counts ? malloc(sizeof(double) * tr->n * C * AA) : NULL;
  for (int v = 0; v < tr->n - 1; v++)
    for (int c = 0; c < C; c++) {
      double t = tr->len[v] * rates[c];
      spectral_pmatrix(&tb->sp, t, tb->P + ((size_t)v * C + c) * AA);
      if (want_counts) counts_any(method, &tb->sp, Q, weights, t, tb->N + ((size_t)v * C + c) * AA);
    }
  return 0;
}
static void tables_free(tables_t* tb) { spectral_free(&tb->sp); free(tb->P); free(tb->N); }

/*
 * Double-recursive likelihood + substitution vectors.
 *
 * [Bio++] DRHomogeneousTreeLikelihood::initialize / setData (CoETools.cpp:124,209,358-359;
 * AnalysisTools.cpp:592-593): for every non-root node v two arrays per (site, class, state)
 *   down[v] = likelihoods[father][v]  (subtree below v, state at v; tips = 0/1 masks)
 *   up[v]   = likelihoods[v][father]  (everything else, state at father; root freqs are
 *             folded in for the root's children and propagate down, prefix recursion)
 * [Bio++] LegacySubstitutionMappingTools::computeSubstitutionVectors (CoETools.cpp:397,
 * AnalysisTools.cpp:601, ClusterTools.cpp:227), SURVEY.md s3.3:
 *   n_v(s) = sum_c p_c sum_x up[v][s][c][x] sum_y P_v,c(x,y) N_v,c(x,y) down[v][s][c][y] / L_s
 * getLogLikelihoodPerSite / getPosteriorRatePerSite / getRateClassWithMaxPostProbPerSite
 * (CoETools.cpp:507-510,669-670): logL = ln L_s, PR = sum_c r_c p_c L_sc / L_s,
 * RC = first argmax_c L_sc (not weighted by p_c).
 * computeNormForSite (CoMap.cpp:158-163, AnalysisTools.cpp:343-350): sqrt(sum_b n_b^2).
 */
/* [Bio++] DRHomogeneousTreeLikelihood::computeLikelihoodAtNode: conditional likelihood of all the data given
 * state x at node v and class c = product of the arrays of all its neighbours (sons in order, then the father's,
 * or the root frequencies at the root). */
static double node_lik(const tree_t* tr, const tables_t* tb, const double* down, const double* up, size_t per_node,
                       size_t CA, int v, int64_t s, int c, int x) {
  const int A = tb->A, C = tb->C;
  const size_t AA = (size_t)A * A;
  double l = 1.;
  if (tr->nch[v] == 0) l = down[per_node * v + s * CA + c * A + x];
  for (int k = 0; k < tr->nch[v]; k++) {
    const int w = tr->ch[tr->ch_off[v] + k];
    const double* P = tb->P + ((size_t)w * C + c) * AA;
    const double* dw = down + per_node * w + s * CA + c * A;
    double m = 0.;
    for (int y = 0; y < A; y++) m += P[x * A + y] * dw[y];
    l *= m;
  }
  if (v == tr->root) return l * tb->pi[x];
  const double* P = tb->P + ((size_t)v * C + c) * AA;
  const double* uv = up + per_node * v + s * CA + c * A;
  double m = 0.;
  for (int z = 0; z < A; z++) m += uv[z] * P[z * A + x];
  return l * m;
}

/* nijt.average / nijt.joint (CoETools.cpp:393-407; AnalysisTools.cpp:597-633 for the null): which of the four
 * LegacySubstitutionMappingTools functions fills the vectors.  [Bio++ / from memory, no golden: parity unpinned]
 *   average, joint   computeSubstitutionVectors: sum over (x, y) of the joint posterior times the count
 *   average, !joint  ...Marginal: product of the two nodes' marginal posteriors per (state, class) instead of
 *                    the joint one (getPosteriorProbabilitiesPerStatePerRate; a leaf's is its 0/1 array x p_c)
 *   !average, joint  ...NoAveraging: the pair (x, y) of largest joint probability summed over the classes
 *                    (MatrixTools::whichMax: first maximum, rows then columns) and its count averaged over the
 *                    classes with that pair's posterior class weights
 *   !average, !joint ...NoAveragingMarginal: the marginal reconstruction of both nodes
 *                    (MarginalAncestralStateReconstruction: first arg-max of sum_c p_c L_c(x) / L; a leaf's first
 *                    compatible state) and the count of that pair averaged over the classes with p_c */
static int g_map_average = 1, g_map_joint = 1;
static uint8_t* g_anc_out = NULL; /* orc_ancestral_states: where the marginal reconstruction goes */
void orc_set_map_mode(int average, int joint) { g_map_average = average != 0; g_map_joint = joint != 0; }

static int map_core(const tree_t* tr, const tables_t* tb, int64_t S, const uint8_t* codes,
                    int n_codes, const uint32_t* code_mask, double* n_out, double* norm,
                    double* post_rate, int32_t* rate_class, double* loglik) {
  const int A = tb->A, C = tb->C, n = tr->n, root = tr->root;
  const size_t CA = (size_t)C * A, AA = (size_t)A * A;
  const size_t per_node = (size_t)S * CA;
  double* down = malloc(sizeof(double) * per_node * n);
  double* up = n_out ? malloc(sizeof(double) * per_node * n) : NULL;
  double* Lsc = malloc(sizeof(double) * S * C);
  double* Ls = malloc(sizeof(double) * S);
  if (!down || (n_out && !up) || !Lsc || !Ls) FAIL("map: out of memory");
  uint32_t full = (A >= 32) ? 0xffffffffu : ((1u << A) - 1u);

  /* postfix pass */
  for (int v = 0; v < n; v++) {
    double* dv = down + per_node * v;
    if (tr->nch[v] == 0) {
      const uint8_t* row = codes + (size_t)tr->leaf_row[v] * S;
      for (int64_t s = 0; s < S; s++) {
        if (row[s] >= n_codes) FAIL("map: code %d out of range at tip row %d", row[s], tr->leaf_row[v]);
        uint32_t m = code_mask[row[s]] & full;
        for (int c = 0; c < C; c++)
          for (int x = 0; x < A; x++) dv[s * CA + c * A + x] = (m >> x) & 1u ? 1. : 0.;
      }
    } else {
      for (size_t k = 0; k < per_node; k++) dv[k] = 1.;
      for (int k = 0; k < tr->nch[v]; k++) {
        int w = tr->ch[tr->ch_off[v] + k];
        const double* dw = down + per_node * w;
        for (int64_t s = 0; s < S; s++)
          for (int c = 0; c < C; c++) {
            const double* P = tb->P + ((size_t)w * C + c) * AA;
            const double* dws = dw + s * CA + c * A;
            for (int x = 0; x < A; x++) {
              double l = 0.;
              for (int y = 0; y < A; y++) l += P[x * A + y] * dws[y];
              dv[s * CA + c * A + x] *= l;
            }
          }
      }
    }
  }
  /* root: product over sons already in down[root]; apply root frequencies */
  {
    const double* dr = down + per_node * root;
    for (int64_t s = 0; s < S; s++) {
      double L = 0.;
      for (int c = 0; c < C; c++) {
        double l = 0.;
        for (int x = 0; x < A; x++) l += dr[s * CA + c * A + x] * tb->pi[x];
        Lsc[s * C + c] = l;
        L += l * tb->probs[c];
      }
      Ls[s] = L;
      if (loglik) loglik[s] = log(L);
      if (post_rate) {
        double r = 0.;
        for (int c = 0; c < C; c++) r += (Lsc[s * C + c] / L) * tb->probs[c] * tb->rates[c];
        post_rate[s] = r;
      }
      if (rate_class) {
        int best = 0;
        for (int c = 1; c < C; c++)
          if (Lsc[s * C + c] > Lsc[s * C + best]) best = c;
        rate_class[s] = best;
      }
    }
  }
  if (n_out) {
    const int B = n - 1;
    /* prefix pass: parents have larger ids, so walk ids downwards */
    for (int v = n - 2; v >= 0; v--) {
      int f = tr->parent[v];
      double* uv = up + per_node * v;
      for (size_t k = 0; k < per_node; k++) uv[k] = 1.;
      for (int k = 0; k < tr->nch[f]; k++) {
        int w = tr->ch[tr->ch_off[f] + k];
        if (w == v) continue;
        const double* dw = down + per_node * w;
        for (int64_t s = 0; s < S; s++)
          for (int c = 0; c < C; c++) {
            const double* P = tb->P + ((size_t)w * C + c) * AA;
            const double* dws = dw + s * CA + c * A;
            for (int x = 0; x < A; x++) {
              double l = 0.;
              for (int y = 0; y < A; y++) l += P[x * A + y] * dws[y];
              uv[s * CA + c * A + x] *= l;
            }
          }
      }
      if (f != root) {
        const double* uf = up + per_node * f;
        for (int64_t s = 0; s < S; s++)
          for (int c = 0; c < C; c++) {
            const double* P = tb->P + ((size_t)f * C + c) * AA;
            const double* ufs = uf + s * CA + c * A;
            for (int x = 0; x < A; x++) {
              double l = 0.;
              for (int y = 0; y < A; y++) l += P[y * A + x] * ufs[y]; /* transposed */
              uv[s * CA + c * A + x] *= l;
            }
          }
      } else {
        for (int64_t s = 0; s < S; s++)
          for (int c = 0; c < C; c++)
            for (int x = 0; x < A; x++) uv[s * CA + c * A + x] *= tb->pi[x];
      }
    }
    /* marginal reconstruction of every node (variants without averaging and without the joint pair) */
    int32_t* anc = NULL;
    const int variant = (g_map_average ? 0 : 2) + (g_map_joint ? 0 : 1); /* 0 avg+joint, 1 marginal, 2 no-avg, 3 no-avg marginal */
    if (variant == 3) {
      anc = malloc(sizeof(int32_t) * (size_t)n * S);
      for (int v = 0; v < n; v++)
        for (int64_t s = 0; s < S; s++) {
          int best = 0;
          if (tr->nch[v] == 0) { /* whichMax of the leaf's 0/1 array */
            const double* dv = down + per_node * v + s * CA;
            for (int x = 1; x < A; x++) if (dv[x] > dv[best]) best = x;
          } else {
            double bv = -INFINITY;
            for (int x = 0; x < A; x++) {
              double l = 0.;
              for (int c = 0; c < C; c++) l += node_lik(tr, tb, down, up, per_node, CA, v, s, c, x) * tb->probs[c] / Ls[s];
              if (l > bv) { bv = l; best = x; }
            }
          }
          anc[(size_t)v * S + s] = best;
          if (g_anc_out) g_anc_out[(size_t)v * S + s] = (uint8_t)best;
        }
    }
    /* mapping */
    for (int v = 0; v < B && variant; v++) {
      const double* dv = down + per_node * v;
      const double* uv = up + per_node * v;
      const int f = tr->parent[v];
      for (int64_t s = 0; s < S; s++) {
        double res = 0.;
        if (variant == 2) {
          double best = -INFINITY; int bx = 0, by = 0;
          for (int x = 0; x < A; x++)
            for (int y = 0; y < A; y++) {
              double pr = 0.;
              for (int c = 0; c < C; c++)
                pr += tb->probs[c] * uv[s * CA + c * A + x] * tb->P[((size_t)v * C + c) * AA + x * A + y] * dv[s * CA + c * A + y];
              if (pr > best) { best = pr; bx = x; by = y; }
            }
          double sc = 0.;
          for (int c = 0; c < C; c++)
            sc += tb->probs[c] * uv[s * CA + c * A + bx] * tb->P[((size_t)v * C + c) * AA + bx * A + by] * dv[s * CA + c * A + by]
                  * tb->N[((size_t)v * C + c) * AA + bx * A + by];
          res = sc / best;
        } else if (variant == 3) {
          const int x = anc[(size_t)f * S + s], y = anc[(size_t)v * S + s];
          for (int c = 0; c < C; c++) res += tb->N[((size_t)v * C + c) * AA + x * A + y] * tb->probs[c];
        } else { /* marginal posteriors of the father and of the node, per (class, state) */
          double Lf = 0., Lv = 0.;
          for (int c = 0; c < C; c++)
            for (int x = 0; x < A; x++) {
              Lf += node_lik(tr, tb, down, up, per_node, CA, f, s, c, x) * tb->probs[c];
              if (tr->nch[v]) Lv += node_lik(tr, tb, down, up, per_node, CA, v, s, c, x) * tb->probs[c];
            }
          for (int c = 0; c < C; c++)
            for (int x = 0; x < A; x++) {
              const double pf = node_lik(tr, tb, down, up, per_node, CA, f, s, c, x) * tb->probs[c] / Lf;
              for (int y = 0; y < A; y++) {
                const double pv = tr->nch[v] ? node_lik(tr, tb, down, up, per_node, CA, v, s, c, y) * tb->probs[c] / Lv
                                             : dv[s * CA + c * A + y] * tb->probs[c];
                res += pf * pv * tb->N[((size_t)v * C + c) * AA + x * A + y];
              }
            }
        }
        n_out[s * B + v] = res;
      }
    }
    free(anc);
    for (int v = 0; v < B && !variant; v++) {
      const double* dv = down + per_node * v;
      const double* uv = up + per_node * v;
      for (int64_t s = 0; s < S; s++) {
        double acc = 0.;
        for (int c = 0; c < C; c++) {
          const double* P = tb->P + ((size_t)v * C + c) * AA;
          const double* N = tb->N + ((size_t)v * C + c) * AA;
          double pc = tb->probs[c];
          for (int x = 0; x < A; x++) {
            double fx = pc * uv[s * CA + c * A + x];
            for (int y = 0; y < A; y++) {
              double lcxy = fx * P[x * A + y] * dv[s * CA + c * A + y];
              acc += lcxy * N[x * A + y];
            }
          }
        }
        n_out[s * B + v] = acc / Ls[s];
      }
    }
    if (norm)
      for (int64_t s = 0; s < S; s++) {
        double q = 0.;
        for (int b = 0; b < B; b++) q += n_out[s * B + b] * n_out[s * B + b];
        norm[s] = sqrt(q);
      }
  }
  free(down); free(up); free(Lsc); free(Ls);
  return 0;
}

/* asr.method = marginal (CoMap.cpp:168-198) [Bio++ LegacyMarginalAncestralStateReconstruction::getAllAncestralStates,
 * from memory]: per node and site the first arg-max over the states of sum_c p_c L_c(x) / L (computeLikelihoodAtNode);
 * a leaf: whichMax of its 0/1 array.  states: [n_nodes][S]. */
int orc_ancestral_states(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                         const double* pi, int C, const double* rates, const double* probs, int64_t S,
                         const uint8_t* codes, int n_codes, const uint32_t* code_mask, uint8_t* states) {
  tree_t tr; tables_t tb;
  if (tree_init(&tr, n_nodes, parent, brlen)) { tree_free(&tr); return -1; }
  if (tables_init(&tb, &tr, A, Q, pi, C, rates, probs, ORC_COUNT_NAIVE, NULL, 1)) { tree_free(&tr); return -1; }
  double* n_out = malloc(sizeof(double) * (size_t)S * (n_nodes - 1));
  const int av = g_map_average, jo = g_map_joint;
  g_map_average = 0; g_map_joint = 0; g_anc_out = states;
  int rc = map_core(&tr, &tb, S, codes, n_codes, code_mask, n_out, NULL, NULL, NULL, NULL);
  g_map_average = av; g_map_joint = jo; g_anc_out = NULL;
  free(n_out); tables_free(&tb); tree_free(&tr);
  return rc;
}

int orc_map(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
            const double* pi, int C, const double* rates, const double* probs, int method,
            const double* weights, int64_t S, const uint8_t* codes, int n_codes,
            const uint32_t* code_mask, double* n_out, double* norm, double* post_rate,
            int32_t* rate_class, double* loglik) {
  tree_t tr; tables_t tb;
  if (norm && !n_out) FAIL("orc_map: norm requires n_out");
  if (tree_init(&tr, n_nodes, parent, brlen)) { tree_free(&tr); return -1; }
  if (tables_init(&tb, &tr, A, Q, pi, C, rates, probs, method, weights, n_out != NULL)) {
    tree_free(&tr); return -1;
  }
  int rc = map_core(&tr, &tb, S, codes, n_codes, code_mask, n_out, norm, post_rate, rate_class, loglik);
  tables_free(&tb); tree_free(&tr);
  return rc;
}

/* ------------------------------------------------------------------------------------ */
/* statistics  (CoMap/Statistics.h:106-295; VectorTools::cor/cov/cos are [Bio++])        */
/* ------------------------------------------------------------------------------------ */

/* [Bio++] VectorTools::cov(v1, v2, unbiased=true) = <center(v1), center(v2)>/n * n/(n-1) */
static double vt_cov(int n, const double* a, const double* b) {
  double ma = 0., mb = 0.;
  for (int i = 0; i < n; i++) { ma += a[i]; mb += b[i]; }
  ma /= (double)n; mb /= (double)n;
  double s = 0.;
  for (int i = 0; i < n; i++) s += (a[i] - ma) * (b[i] - mb);
  double x = s / (double)n;
  x = x * (double)n / ((double)n - 1.);
  return x;
}

static double* g_mean_vector = NULL;  /* [2][B]: mean vector of the first / second site's data set */
static int g_mean_vector_len = 0;
void orc_set_mean_vectors(int B, const double* mv1, const double* mv2) { /* setMeanVectors, Statistics.h:200-203 */
  free(g_mean_vector);
  g_mean_vector = malloc(sizeof(double) * (size_t)B * 2);
  memcpy(g_mean_vector, mv1, sizeof(double) * (size_t)B);
  memcpy(g_mean_vector + B, mv2, sizeof(double) * (size_t)B);
  g_mean_vector_len = B;
}
void orc_set_mean_vector(int B, const double* mv) { orc_set_mean_vectors(B, mv, mv); }
/* CoMap.cpp:350-359 */
void orc_mean_vector(int64_t S, int B, const double* n, double* mv) {
  for (int b = 0; b < B; b++) mv[b] = 0.;
  for (int64_t i = 0; i < S; i++)
    for (int b = 0; b < B; b++) mv[b] += n[i * B + b];
  for (int b = 0; b < B; b++) mv[b] /= (double)S;
}

static double g_mi_threshold = 0.99;
void orc_set_mi_threshold(double threshold) { g_mi_threshold = threshold; }

/* statistic=MI(threshold=t) with nijt != Label (CoETools.cpp:590-595): DiscreteMutualInformationStatistic
 * with bounds {0, t, 10000} (Statistics.h:307-329): category of a branch = Domain(bounds).getIndex(sum of
 * its counts) = [v >= t]; then [Bio++ / from memory] VectorTools::miDiscrete(c1, c2, base = 2.7182818):
 * frequency maps iterated in key order, s += (n12 / n) * log(n12 * n / (n1 * n2)) / log(base). */
static double mi_discrete_binary(int B, const double* v1, const double* v2, double t) {
  double c1[2] = {0., 0.}, c2[2] = {0., 0.}, c12[2][2] = {{0., 0.}, {0., 0.}};
  for (int i = 0; i < B; i++) {
    int a = v1[i] >= t, b = v2[i] >= t;
    c1[a]++; c2[b]++; c12[a][b]++;
  }
  double s = 0., n = (double)B;
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++)
      if (c12[a][b] > 0.) s += (c12[a][b] / n) * log(c12[a][b] * n / (c1[a] * c2[b])) / log(2.7182818);
  return s;
}

/* statistic=MI with nijt=Label (CoETools.cpp:577-589): bounds -0.5, 0.5, ..., A(A-1) + 0.5, i.e. category of a
 * branch = its label; the same [Bio++ / from memory] miDiscrete over all the categories present. */
static int g_mi_label_states = 0;
void orc_set_mi_label(int n_states) { g_mi_label_states = n_states; }
static double mi_discrete_labels(int B, const double* v1, const double* v2, int A) {
  const int K = A * (A - 1) + 1;
  double* c1 = calloc((size_t)K * 2 + (size_t)K * K, sizeof(double));
  double *c2 = c1 + K, *c12 = c2 + K;
  for (int i = 0; i < B; i++) {
    int a = orc_domain_index(-0.5, (double)K - 0.5, K, v1[i]), b = orc_domain_index(-0.5, (double)K - 0.5, K, v2[i]);
    if (a < 0 || b < 0) { free(c1); return NAN; } /* Domain::getIndex throws: cannot happen with labels */
    c1[a]++; c2[b]++; c12[(size_t)a * K + b]++;
  }
  double s = 0., n = (double)B;
  for (int a = 0; a < K; a++)
    for (int b = 0; b < K; b++)
      if (c12[(size_t)a * K + b] > 0.) s += (c12[(size_t)a * K + b] / n) * log(c12[(size_t)a * K + b] * n / (c1[a] * c2[b])) / log(2.7182818);
  free(c1);
  return s;
}

double orc_stat(int stat_id, int B, const double* v1, const double* v2) {
  switch (stat_id) {
    case ORC_STAT_MI_LABEL: return mi_discrete_labels(B, v1, v2, g_mi_label_states);
    case ORC_STAT_MI: return mi_discrete_binary(B, v1, v2, g_mi_threshold);
    case ORC_STAT_CORRECTED_CORRELATION: { /* Statistics.h:188-194: cor(v1 - mean vector, v2 - mean vector) */
      if (g_mean_vector_len != B) return NAN;
      double* a = malloc(sizeof(double) * (size_t)B * 2);
      double* b = a + B;
      for (int i = 0; i < B; i++) { a[i] = v1[i] - g_mean_vector[i]; b[i] = v2[i] - g_mean_vector[B + i]; }
      double r = vt_cov(B, a, b) / (sqrt(vt_cov(B, a, a)) * sqrt(vt_cov(B, b, b)));
      free(a);
      return r;
    }
    case ORC_STAT_CORRELATION: /* Statistics.h:164-174 -> cov / (sd * sd) */
      return vt_cov(B, v1, v2) / (sqrt(vt_cov(B, v1, v1)) * sqrt(vt_cov(B, v2, v2)));
    case ORC_STAT_COVARIANCE: /* Statistics.h:206-216 */
      return vt_cov(B, v1, v2);
    case ORC_STAT_COSINUS: { /* Statistics.h:218-228 -> scalar / (norm * norm) */
      double s = 0., n1 = 0., n2 = 0.;
      for (int i = 0; i < B; i++) { s += v1[i] * v2[i]; n1 += v1[i] * v1[i]; n2 += v2[i] * v2[i]; }
      return s / (sqrt(n1) * sqrt(n2));
    }
    case ORC_STAT_COSUBSTITUTION: { /* Statistics.h:230-245 */
      double c = 0.;
      for (int i = 0; i < B; i++)
        if (v1[i] >= 1. && v2[i] >= 1.) c++;
      return c;
    }
    case ORC_STAT_COMPENSATION: { /* Statistics.h:247-265 */
      double s1 = 0., s2 = 0., s3 = 0.;
      for (int i = 0; i < B; i++) {
        s1 += pow(v1[i], 2);
        s2 += pow(v2[i], 2);
        s3 += pow(v1[i] + v2[i], 2);
      }
      return 1. - sqrt(s3) / (sqrt(s1) + sqrt(s2));
    }
  }
  return NAN;
}

double orc_stat2(int stat_id, int B, const double* v1, const double* v2) { return orc_stat(stat_id, B, v1, v2); }

double orc_stat_group(int stat_id, int B, const double* n, int n_members, const int32_t* members) {
  if (stat_id == ORC_STAT_COMPENSATION) { /* Statistics.h:267-294 */
    double* sq = calloc(n_members, sizeof(double));
    double sumsq2 = 0.;
    for (int i = 0; i < B; i++) {
      double s = 0.;
      for (int j = 0; j < n_members; j++) {
        double sv = n[(size_t)members[j] * B + i];
        sq[j] += pow(sv, 2);
        s += sv;
      }
      sumsq2 += pow(s, 2);
    }
    double sumnorms = 0.;
    for (int j = 0; j < n_members; j++) sumnorms += sqrt(sq[j]);
    free(sq);
    return 1. - sqrt(sumsq2) / sumnorms;
  }
  /* AbstractMinimumStatistic::getValueForGroup, Statistics.h:121-133 */
  double mini = INFINITY;
  for (int i = 1; i < n_members; i++)
    for (int j = 0; j < i; j++) {
      double val = orc_stat(stat_id, B, n + (size_t)members[i] * B, n + (size_t)members[j] * B);
      if (val < mini) mini = val;
    }
  return mini;
}

/* Domain(a, b, n) + getIndex, CoMap/Domain.cpp:46-59,113-122 */
int orc_domain_index(double lo, double hi, int K, double x) {
  double mini = lo < hi ? lo : hi, maxi = lo < hi ? hi : lo;
  double w = (maxi - mini) / (double)K;
  double upper = mini + (double)K * w;
  if (x < mini || x >= upper) return -1;
  for (int i = 1; i < K + 1; i++)
    if (x < mini + (double)i * w) return i - 1;
  return -1;
}

/* CoETools::computeIntraStats pair loop, CoETools.cpp:672-724 */
int orc_pairs(int stat_id, int64_t S, int B, const double* n, const double* norm,
              const double* post_rate, const int32_t* rate_class, int min_rate_class,
              double min_rate, int max_rate_class_diff, double max_rate_diff, double min_stat,
              int K, double nmax, const int64_t* bin_offsets, const double* sorted_null,
              int64_t capacity, int32_t* out_i, int32_t* out_j, double* out_stat,
              int32_t* out_rcmin, double* out_prmin, double* out_nmin, double* out_pvalue,
              int64_t* out_nsim, int64_t* n_rows) {
  int64_t r = 0;
  for (int64_t i = 0; i < S; i++) {
    int iClass = rate_class[i];
    double iRate = post_rate[i];
    if (iClass < min_rate_class) continue;
    if (iRate < min_rate) continue;
    double iNorm = norm[i];
    for (int64_t j = i + 1; j < S; j++) {
      int jClass = rate_class[j];
      double jRate = post_rate[j];
      if (jClass < min_rate_class) continue;
      if (jRate < min_rate) continue;
      double jNorm = norm[j];
      if (max_rate_class_diff >= 0 && abs(jClass - iClass) > max_rate_class_diff) continue;
      if (max_rate_diff >= 0. && fabs(jRate - iRate) > max_rate_diff) continue;
      double stat = orc_stat(stat_id, B, n + i * B, n + j * B);
      if (fabs(stat) < min_stat) continue;
      double minNorm = iNorm < jNorm ? iNorm : jNorm;
      if (r >= capacity) FAIL("orc_pairs: capacity exceeded");
      out_i[r] = (int32_t)i; out_j[r] = (int32_t)j; out_stat[r] = stat;
      out_rcmin[r] = iClass < jClass ? iClass : jClass;
      out_prmin[r] = iRate < jRate ? iRate : jRate;
      out_nmin[r] = minNorm;
      if (K > 0) {
        int cat = orc_domain_index(0., nmax, K, minNorm);
        if (cat >= 0) {
          const double* sim = sorted_null + bin_offsets[cat];
          int64_t nsim = bin_offsets[cat + 1] - bin_offsets[cat], count;
          for (count = 0; count < nsim && sim[count] < stat; ++count) {}
          out_pvalue[r] = (double)(nsim - count + 1) / (double)(nsim + 1);
          out_nsim[r] = nsim;
        } else { /* OutOfRangeException -> "NA\t0" */
          out_pvalue[r] = NAN;
          out_nsim[r] = 0;
        }
      }
      r++;
    }
  }
  *n_rows = r;
  return 0;
}

/* CoETools::computeInterStats (CoETools.cpp:732-840): data set 1 x data set 2 (or i <-> i when
 * independent); the statistic's mean vectors, if any, are the caller's business.  nmin_by_row
 * keeps upstream's jNorm = norms2[i] (CoETools.cpp:803). */
int orc_pairs_inter(int stat_id, int64_t S1, int64_t S2, int B, const double* n1, const double* n2,
                    const double* norm1, const double* norm2, const double* pr1, const double* pr2,
                    const int32_t* rc1, const int32_t* rc2, int min_rate_class1, int min_rate_class2,
                    double min_rate1, double min_rate2, int max_rate_class_diff, double max_rate_diff,
                    double min_stat, int independent, int nmin_by_row, int64_t capacity, int32_t* out_i,
                    int32_t* out_j, double* out_stat, int32_t* out_rcmin, double* out_prmin,
                    double* out_nmin, int64_t* n_rows) {
  int64_t r = 0;
  for (int64_t i = 0; i < S1; i++) {
    int iClass = rc1[i];
    double iRate = pr1[i];
    if (iClass < min_rate_class1) continue;
    if (iRate < min_rate1) continue;
    double iNorm = norm1[i];
    int64_t begin = independent ? i : 0, end = independent ? i + 1 : S2;
    for (int64_t j = begin; j < end; j++) {
      int jClass = rc2[j];
      double jRate = pr2[j];
      if (jClass < min_rate_class2) continue;
      if (jRate < min_rate2) continue;
      double jNorm = norm2[(nmin_by_row && i < S2) ? i : j];
      if (max_rate_class_diff >= 0 && abs(jClass - iClass) > max_rate_class_diff) continue;
      if (max_rate_diff >= 0. && fabs(jRate - iRate) > max_rate_diff) continue;
      double stat = orc_stat2(stat_id, B, n1 + i * B, n2 + j * B);
      if (fabs(stat) < min_stat) continue;
      if (r >= capacity) FAIL("orc_pairs_inter: capacity exceeded");
      out_i[r] = (int32_t)i; out_j[r] = (int32_t)j; out_stat[r] = stat;
      out_rcmin[r] = iClass < jClass ? iClass : jClass;
      out_prmin[r] = iRate < jRate ? iRate : jRate;
      out_nmin[r] = iNorm < jNorm ? iNorm : jNorm;
      r++;
    }
  }
  *n_rows = r;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* simulation                                                                            */
/* ------------------------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al. 2011).  Counter = (site lo, site hi, node, tag),
 * key = (seed lo, seed hi).  Shared bit-for-bit with comap_b200/csrc (own code there). */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
static double philox_u01(uint64_t seed, uint64_t site, uint32_t node, uint32_t tag) {
  uint32_t c[4] = {(uint32_t)site, (uint32_t)(site >> 32), node, tag};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t u = (((uint64_t)c[0] << 32) | c[1]) >> 11;
  return (double)u * (1.0 / 9007199254740992.0);
}

/* two uniforms from one block: (c0, c1) and (c2, c3) */
static void philox_u01x2(uint64_t seed, uint64_t site, uint32_t node, uint32_t tag, double* u0, double* u1) {
  uint32_t c[4] = {(uint32_t)site, (uint32_t)(site >> 32), node, tag};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  *u0 = (double)((((uint64_t)c[0] << 32) | c[1]) >> 11) * (1.0 / 9007199254740992.0);
  *u1 = (double)((((uint64_t)c[2] << 32) | c[3]) >> 11) * (1.0 / 9007199254740992.0);
}

/* [Bio++] NonHomogeneousSequenceSimulator::simulate(n) in discrete-rate mode
 * (CoMap.cpp:209-219; AnalysisTools.cpp:591,614; ClusterTools.cpp:224): per site the root
 * state is drawn from pi (first i with r <= cumulative), the rate class uniformly over
 * classes (upstream behaviour; weighted_classes=1 draws from probs
 * instead), each child state by linear inverse-CDF over the cumulative row of P_c(b). */
int orc_simulate(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                 const double* pi, int C, const double* rates, const double* probs, uint64_t seed,
                 int64_t first_site, int64_t n, int weighted_classes, uint8_t* states,
                 int32_t* classes) {
  tree_t tr; tables_t tb;
  if (tree_init(&tr, n_nodes, parent, brlen)) { tree_free(&tr); return -1; }
  if (tables_init(&tb, &tr, A, Q, pi, C, rates, probs, 0, NULL, 0)) { tree_free(&tr); return -1; }
  size_t AA = (size_t)A * A;
  double* cum = malloc(sizeof(double) * tr.n * C * AA);
  for (size_t m = 0; m < (size_t)(tr.n - 1) * C; m++)
    for (int x = 0; x < A; x++) {
      double s = 0.;
      for (int y = 0; y < A; y++) { s += tb.P[m * AA + x * A + y]; cum[m * AA + x * A + y] = s; }
    }
  uint8_t* st = malloc(tr.n);
  /* RNG key of node v: its parent and its rank k among the parent's children (id order); the
   * state is drawn with half k % 2 of the Philox block (parent, tag 2 + k / 2) */
  int* rank = calloc(tr.n, sizeof(int));
  {
    int* seen = calloc(tr.n, sizeof(int));
    for (int v = 0; v < tr.n - 1; v++) rank[v] = seen[tr.parent[v]]++;
    free(seen);
  }
  for (int64_t j = 0; j < n; j++) {
    uint64_t site = (uint64_t)(first_site + j);
    double r = philox_u01(seed, site, (uint32_t)tr.root, 0);
    int x0 = A - 1;
    double cp = 0.;
    for (int i = 0; i < A; i++) { cp += pi[i]; if (r <= cp) { x0 = i; break; } }
    double rc = philox_u01(seed, site, (uint32_t)tr.root, 1);
    int c = 0;
    if (weighted_classes) {
      double cq = 0.; c = C - 1;
      for (int i = 0; i < C; i++) { cq += probs[i]; if (rc <= cq) { c = i; break; } }
    } else {
      c = (int)(rc * (double)C);
      if (c >= C) c = C - 1;
    }
    if (classes) classes[j] = c;
    st[tr.root] = (uint8_t)x0;
    for (int v = tr.n - 2; v >= 0; v--) {
      int x = st[tr.parent[v]];
      double u0, u1;
      philox_u01x2(seed, site, (uint32_t)tr.parent[v], 2u + ((uint32_t)rank[v] >> 1), &u0, &u1);
      double u = (rank[v] & 1) ? u1 : u0;
      const double* row = cum + ((size_t)v * C + c) * AA + (size_t)x * A;
      int y = A - 1;
      for (int k = 0; k < A; k++) if (u < row[k]) { y = k; break; }
      st[v] = (uint8_t)y;
      if (tr.leaf_row[v] >= 0) states[(size_t)tr.leaf_row[v] * n + j] = (uint8_t)y;
    }
  }
  free(rank); free(st); free(cum); tables_free(&tb); tree_free(&tr);
  return 0;
}

/* Continuous site rates (simulations.continuous = yes; CoMap.cpp:146,209-219 -> NonHomogeneousSequenceSimulator::
 * enableContinuousRates -> rDist->randC()): one rate per site, Gamma(alpha, beta = alpha) by Marsaglia & Tsang with
 * Box-Muller normals from the site's Philox stream (node = root, tags 98..), P(d_b r) evaluated per branch from the
 * generator's spectrum, child states by the same inverse-CDF and Philox blocks as orc_simulate.  The reference's
 * generator is unseeded, so only the distribution is common ground: parity unpinned. */
static double continuous_rate(int kind, double alpha, double p_inv, uint64_t seed, uint64_t site, uint32_t node) {
  if (kind == 1) return 1.;
  double scale = 1.;
  if (kind == 3) {
    if (philox_u01(seed, site, node, 98) < p_inv) return 0.;
    scale = 1. / (1. - p_inv);
  }
  double a = alpha, boost = 1.;
  if (a < 1.) {
    boost = pow(1. - philox_u01(seed, site, node, 99), 1. / a);
    a += 1.;
  }
  const double d = a - 1. / 3., c = 1. / sqrt(9. * d);
  double g = d;
  for (uint32_t k = 0; k < 64; k++) {
    double u1, u2;
    philox_u01x2(seed, site, node, 100 + 2 * k, &u1, &u2);
    const double x = sqrt(-2. * log(1. - u1)) * cos(6.283185307179586 * u2);
    double v = 1. + c * x;
    if (v <= 0.) continue;
    v = v * v * v;
    const double u = 1. - philox_u01(seed, site, node, 101 + 2 * k);
    if (u < 1. - 0.0331 * (x * x) * (x * x) || log(u) < 0.5 * x * x + d * (1. - v + log(v))) { g = d * v; break; }
  }
  return g * boost / alpha * scale;
}

int orc_simulate_continuous(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                            const double* pi, int kind, double alpha, double p_inv, uint64_t seed,
                            int64_t first_site, int64_t n, uint8_t* states, double* rates_out) {
  tree_t tr;
  spectral_t sp;
  if (tree_init(&tr, n_nodes, parent, brlen)) { tree_free(&tr); return -1; }
  if (spectral_init(&sp, A, Q, pi)) { tree_free(&tr); return -1; }
  uint8_t* st = malloc(tr.n);
  int* rank = calloc(tr.n, sizeof(int));
  {
    int* seen = calloc(tr.n, sizeof(int));
    for (int v = 0; v < tr.n - 1; v++) rank[v] = seen[tr.parent[v]]++;
    free(seen);
  }
  double* ek = malloc(sizeof(double) * A);
  for (int64_t j = 0; j < n; j++) {
    uint64_t site = (uint64_t)(first_site + j);
    double r = philox_u01(seed, site, (uint32_t)tr.root, 0);
    int x0 = A - 1;
    double cp = 0.;
    for (int i = 0; i < A; i++) { cp += pi[i]; if (r <= cp) { x0 = i; break; } }
    const double rate = continuous_rate(kind, alpha, p_inv, seed, site, (uint32_t)tr.root);
    if (rates_out) rates_out[j] = rate;
    st[tr.root] = (uint8_t)x0;
    for (int v = tr.n - 2; v >= 0; v--) {
      int x = st[tr.parent[v]];
      double u0, u1;
      philox_u01x2(seed, site, (uint32_t)tr.parent[v], 2u + ((uint32_t)rank[v] >> 1), &u0, &u1);
      double u = (rank[v] & 1) ? u1 : u0;
      int y = x;
      if (rate != 0.) {
        const double t = tr.len[v] * rate;
        for (int k = 0; k < A; k++) ek[k] = sp.R[x * A + k] * exp(sp.ev[k] * t);
        double cum = 0.;
        y = A - 1;
        for (int jj = 0; jj < A; jj++) {
          double pj = 0.;
          for (int k = 0; k < A; k++) pj += ek[k] * sp.L[k * A + jj];
          cum += pj;
          if (u < cum) { y = jj; break; }
        }
      }
      st[v] = (uint8_t)y;
      if (tr.leaf_row[v] >= 0) states[(size_t)tr.leaf_row[v] * n + j] = (uint8_t)y;
    }
  }
  free(ek); free(rank); free(st); spectral_free(&sp); tree_free(&tr);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* null distribution  (AnalysisTools.cpp:564-658 + sort at CoETools.cpp:650-652)         */
/* ------------------------------------------------------------------------------------ */

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

int orc_null_intra(int n_nodes, const int32_t* parent, const double* brlen, int A, const double* Q,
                   const double* pi, int C, const double* rates, const double* probs, int method,
                   const double* weights, int stat_id, int rep_cpu, int rep_ram,
                   const uint8_t* sim1, const uint8_t* sim2, int K, double nmax, double* raw,
                   int64_t* bin_offsets, double* sorted) {
  tree_t tr; tables_t tb;
  if (tree_init(&tr, n_nodes, parent, brlen)) { tree_free(&tr); return -1; }
  if (tables_init(&tb, &tr, A, Q, pi, C, rates, probs, method, weights, 1)) { tree_free(&tr); return -1; }
  const int B = tr.n - 1;
  uint32_t* mask = malloc(sizeof(uint32_t) * A);
  for (int x = 0; x < A; x++) mask[x] = 1u << x;
  size_t total = (size_t)rep_cpu * rep_ram;
  double* m1 = malloc(sizeof(double) * (size_t)rep_ram * B);
  double* m2 = malloc(sizeof(double) * (size_t)rep_ram * B);
  double *n1 = malloc(sizeof(double) * rep_ram), *n2 = malloc(sizeof(double) * rep_ram);
  double *p1 = malloc(sizeof(double) * rep_ram), *p2 = malloc(sizeof(double) * rep_ram);
  int32_t *c1 = malloc(sizeof(int32_t) * rep_ram), *c2 = malloc(sizeof(int32_t) * rep_ram);
  double* st = malloc(sizeof(double) * total);
  int* cat = malloc(sizeof(int) * total);
  int rc = 0;
  for (int i = 0; i < rep_cpu && !rc; i++) {
    size_t off = (size_t)i * tr.T * rep_ram;
    rc = map_core(&tr, &tb, rep_ram, sim1 + off, A, mask, m1, n1, p1, c1, NULL);
    if (!rc) rc = map_core(&tr, &tb, rep_ram, sim2 + off, A, mask, m2, n2, p2, c2, NULL);
    if (rc) break;
    for (int j = 0; j < rep_ram; j++) {
      size_t k = (size_t)i * rep_ram + j;
      double stat = orc_stat(stat_id, B, m1 + (size_t)j * B, m2 + (size_t)j * B);
      double nmin = n1[j] < n2[j] ? n1[j] : n2[j];
      st[k] = stat;
      cat[k] = K > 0 ? orc_domain_index(0., nmax, K, nmin) : -1;
      if (raw) {
        raw[k * 4 + 0] = stat;
        raw[k * 4 + 1] = (double)(c1[j] < c2[j] ? c1[j] : c2[j]);
        raw[k * 4 + 2] = p1[j] < p2[j] ? p1[j] : p2[j];
        raw[k * 4 + 3] = nmin;
      }
    }
  }
  if (!rc && K > 0 && bin_offsets && sorted) {
    int64_t* cnt = calloc(K + 1, sizeof(int64_t));
    for (size_t k = 0; k < total; k++) if (cat[k] >= 0) cnt[cat[k] + 1]++;
    bin_offsets[0] = 0;
    for (int b = 0; b < K; b++) bin_offsets[b + 1] = bin_offsets[b] + cnt[b + 1];
    int64_t* pos = malloc(sizeof(int64_t) * K);
    for (int b = 0; b < K; b++) pos[b] = bin_offsets[b];
    for (size_t k = 0; k < total; k++) if (cat[k] >= 0) sorted[pos[cat[k]]++] = st[k];
    for (int b = 0; b < K; b++)
      qsort(sorted + bin_offsets[b], bin_offsets[b + 1] - bin_offsets[b], sizeof(double), cmp_double);
    free(cnt); free(pos);
  }
  free(mask); free(m1); free(m2); free(n1); free(n2); free(p1); free(p2); free(c1); free(c2);
  free(st); free(cat);
  tables_free(&tb); tree_free(&tr);
  return rc;
}

/* ------------------------------------------------------------------------------------ */
/* clustering                                                                            */
/* ------------------------------------------------------------------------------------ */

/* CoMap.cpp:401-440; Distance.h:157-171 (Euclidian), :334-337 (comp - stat, comp = 1 for
 * Correlation, CoMap.cpp:410), :382-385 (1 - Compensation). */
int orc_distance_matrix(int dist_id, int64_t S, int B, const double* n, double* mat) {
  for (int64_t i = 0; i < S; i++) {
    mat[i * S + i] = 0.;
    for (int64_t j = 0; j < i; j++) {
      const double *a = n + i * B, *b = n + j * B;
      double d;
      if (dist_id == ORC_DIST_EUCLIDIAN) {
        d = 0.;
        for (int k = 0; k < B; k++) d += pow(b[k] - a[k], 2);
        d = sqrt(d);
      } else if (dist_id == ORC_DIST_CORRELATION) {
        d = 1. - orc_stat(ORC_STAT_CORRELATION, B, a, b);
      } else if (dist_id == ORC_DIST_COMPENSATION) {
        d = 1. - orc_stat(ORC_STAT_COMPENSATION, B, a, b);
      } else FAIL("orc_distance_matrix: unknown distance %d", dist_id);
      mat[i * S + j] = mat[j * S + i] = d;
    }
  }
  return 0;
}

/* [Bio++] HierarchicalClustering(method, matrix, rootTree) + AbstractAgglomerative-
 * DistanceMethod::computeTree (CoMap.cpp:460-485; ClusterTools.cpp:260-262), SURVEY.md
 * s8 a14: while more than 2 live clusters: first strict minimum over live pairs (i<j)
 * in id order; branch lengths d/2 - height(child); parent takes slot i, slot j dies;
 * Lance-Williams update w1 d1 + w2 d2 + w4 |d1 - d2| with (.5,.5,+.5) complete,
 * (.5,.5,-.5) single, (n1/(n1+n2), n2/(n1+n2), 0) average; the last two clusters are
 * joined at d/2. */
int orc_hclust(int linkage, int64_t S, double* mat, int32_t* left, int32_t* right, double* height) {
  if (S < 2) FAIL("orc_hclust: need at least 2 sites");
  int32_t* node = malloc(sizeof(int32_t) * S);  /* dendrogram node held by slot */
  double* len = calloc(S, sizeof(double));      /* ClusterInfos.length of the slot's node */
  int64_t* nl = malloc(sizeof(int64_t) * S);    /* numberOfLeaves */
  char* alive = malloc(S);
  double* nd = malloc(sizeof(double) * S);
  for (int64_t i = 0; i < S; i++) { node[i] = (int32_t)i; nl[i] = 1; alive[i] = 1; }
  int64_t live = S; int32_t next = (int32_t)S;
  while (live > 2) {
    double dmin = INFINITY; int64_t bi = -1, bj = -1;
    for (int64_t i = 0; i < S; i++) {
      if (!alive[i]) continue;
      for (int64_t j = i + 1; j < S; j++) {
        if (!alive[j]) continue;
        double d = mat[i * S + j];
        if (d < dmin) { dmin = d; bi = i; bj = j; }
      }
    }
    if (bi < 0) FAIL("orc_hclust: no finite distance left");
    double half = mat[bi * S + bj] / 2.;
    double w1, w2, w4;
    if (linkage == ORC_LINK_SINGLE) { w1 = .5; w2 = .5; w4 = -.5; }
    else if (linkage == ORC_LINK_COMPLETE) { w1 = .5; w2 = .5; w4 = .5; }
    else if (linkage == ORC_LINK_AVERAGE) {
      double a = (double)nl[bi], b = (double)nl[bj];
      w1 = a / (a + b); w2 = b / (a + b); w4 = 0.;
    } else FAIL("orc_hclust: unknown linkage %d", linkage);
    for (int64_t k = 0; k < S; k++) {
      if (!alive[k]) continue;
      if (k != bi && k != bj) {
        double d1 = mat[bi * S + k], d2 = mat[bj * S + k];
        nd[k] = w1 * d1 + w2 * d2 + 0. * mat[bi * S + bj] + w4 * fabs(d1 - d2);
      } else nd[k] = 0.;
    }
    /* getParentNode: length(parent) = length(son1) + distToFather(son1) = d/2 */
    left[next - S] = node[bi]; right[next - S] = node[bj];
    double d0 = half - len[bi];
    height[next - S] = len[bi] + d0;
    node[bi] = next; len[bi] = height[next - S]; nl[bi] += nl[bj];
    alive[bj] = 0; live--; next++;
    for (int64_t k = 0; k < S; k++)
      if (alive[k]) mat[bi * S + k] = mat[k * S + bi] = nd[k];
  }
  { /* finalStep */
    int64_t i1 = -1, i2 = -1;
    for (int64_t i = 0; i < S; i++) if (alive[i]) { if (i1 < 0) i1 = i; else i2 = i; }
    double d = mat[i1 * S + i2] / 2;
    left[next - S] = node[i1]; right[next - S] = node[i2];
    height[next - S] = len[i1] + (d - len[i1]);
  }
  free(node); free(len); free(nl); free(alive); free(nd);
  return 0;
}

/* ClusterTools::getGroups (ClusterTools.cpp:59-113): one group per inner node, emitted in
 * post-order, members in DFS leaf order, height = height(last son) + its branch;
 * computeNormProperties (:296-319): Nmin = min leaf norm; Distance::setStatisticAsProperty
 * (Distance.h:109-129 Euclidian: 2*height; :346-368 statistic-based: comp - 2*height;
 * :390-422 compensation: 1 - ||sum v|| / sum ||v||). */
int orc_groups(int dist_id, int64_t S, int B, const double* n, const double* norm,
               const int32_t* left, const int32_t* right, const double* height, int max_size,
               int32_t* members, int64_t* offsets, double* g_height, double* g_stat,
               double* g_nmin, int64_t* n_groups) {
  int64_t n_inner = S - 1, ng = 0, fill = 0;
  int32_t root = (int32_t)(2 * S - 2);
  /* iterative post-order; each inner node's members = concat(left members, right members) */
  int32_t* stack = malloc(sizeof(int32_t) * 2 * S);
  char* state = calloc(2 * S, 1);
  int64_t* first = malloc(sizeof(int64_t) * 2 * S); /* start of this node's leaves in DFS order */
  int32_t* dfs = malloc(sizeof(int32_t) * S);
  int64_t ndfs = 0; int sp = 0;
  stack[sp++] = root;
  offsets[0] = 0;
  while (sp > 0) {
    int32_t v = stack[sp - 1];
    if (v < S) { first[v] = ndfs; dfs[ndfs++] = v; sp--; continue; }
    if (state[v] == 0) { first[v] = ndfs; state[v] = 1; stack[sp++] = left[v - S]; continue; }
    if (state[v] == 1) { state[v] = 2; stack[sp++] = right[v - S]; continue; }
    sp--;
    int64_t cnt = ndfs - first[v];
    if (cnt > max_size) continue; /* CoMap.cpp:517, ClusterTools.cpp:273 */
    double nmin = INFINITY;
    for (int64_t k = 0; k < cnt; k++) {
      int32_t m = dfs[first[v] + k];
      members[fill + k] = m;
      if (norm[m] < nmin) nmin = norm[m];
    }
    double h = height[v - S], stat;
    if (dist_id == ORC_DIST_COMPENSATION) stat = orc_stat_group(ORC_STAT_COMPENSATION, B, n, (int)cnt, members + fill);
    else if (dist_id == ORC_DIST_CORRELATION) stat = 1. - 2 * h;
    else stat = 2 * h;
    fill += cnt;
    offsets[ng + 1] = fill;
    g_height[ng] = h; g_stat[ng] = stat; g_nmin[ng] = nmin;
    ng++;
    if (ng > n_inner) FAIL("orc_groups: malformed dendrogram");
  }
  *n_groups = ng;
  free(stack); free(state); free(first); free(dfs);
  return 0;
}


/* ------------------------------------------------------------------------------------ */
/* Mica (CoMap/Mica.cpp): mutual information between alignment columns                   */
/* ------------------------------------------------------------------------------------ */

/* [Bio++ bpp-seq, from memory; no output of mica ships with the reference: parity unpinned]
 * SymbolListTools::getCounts / getFrequencies with resolveUnknowns = true: a character compatible with k states adds
 * 1/k to each of them (a pair of characters 1/(k1 k2) to every combination); frequencies = counts / number of
 * sequences.  SiteTools::getSitesToAnalyse turns gaps into unknown characters, so a column holds no gap; a character
 * outside the alphabet (mask 0) counts for nothing and the "correction for gaps" of mutualInformation / jointEntropy
 * (division by the total of the joint table) absorbs it. */
static int popcount32(uint32_t m) { int k = 0; while (m) { k += m & 1u; m >>= 1; } return k; }

/* SiteTools::entropy(site, true): - sum_x p_x ln p_x   (Mica.cpp:357) */
double orc_site_entropy(int T, const uint8_t* col, int A, int n_codes, const uint32_t* code_mask) {
  double* cnt = calloc(A, sizeof(double));
  uint32_t full = (A >= 32) ? 0xffffffffu : ((1u << A) - 1u);
  for (int t = 0; t < T; t++) {
    uint32_t m = col[t] < n_codes ? code_mask[col[t]] & full : 0;
    int k = popcount32(m);
    if (!k) continue;
    for (int x = 0; x < A; x++)
      if ((m >> x) & 1u) cnt[x] += 1. / (double)k;
  }
  double h = 0.;
  for (int x = 0; x < A; x++) {
    double f = cnt[x] / (double)T;
    if (f != 0.) h += f * log(f);
  }
  free(cnt);
  return -h;
}

/* SiteTools::mutualInformation(site1, site2, true) and jointEntropy(site1, site2, true) (Mica.cpp:92,354,430):
 * joint frequencies p12, tot = their sum over the alphabet's states, marginals from the joint table, then
 * MI = sum p ln(p / (p1 p2)), Hjoint = - sum p ln p with p = p12 / tot, rows then columns. */
void orc_site_pair(int T, const uint8_t* c1, const uint8_t* c2, int A, int n_codes, const uint32_t* code_mask,
                   double* mi_out, double* hjoint_out) {
  double* cnt = calloc((size_t)A * A + 2 * A, sizeof(double));
  double *p1 = cnt + (size_t)A * A, *p2 = p1 + A;
  uint32_t full = (A >= 32) ? 0xffffffffu : ((1u << A) - 1u);
  for (int t = 0; t < T; t++) {
    uint32_t m1 = c1[t] < n_codes ? code_mask[c1[t]] & full : 0, m2 = c2[t] < n_codes ? code_mask[c2[t]] & full : 0;
    int k1 = popcount32(m1), k2 = popcount32(m2);
    if (!k1 || !k2) continue;
    double w = 1. / ((double)k1 * (double)k2);
    for (int x = 0; x < A; x++)
      if ((m1 >> x) & 1u)
        for (int y = 0; y < A; y++)
          if ((m2 >> y) & 1u) cnt[x * A + y] += w;
  }
  double tot = 0.;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      double pxy = cnt[x * A + y] / (double)T;
      tot += pxy; p1[x] += pxy; p2[y] += pxy;
    }
  for (int x = 0; x < A; x++) { p1[x] /= tot; p2[x] /= tot; }
  double mi = 0., h = 0.;
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      double pxy = cnt[x * A + y] / (double)T / tot;
      if (pxy > 0.) { mi += pxy * log(pxy / (p1[x] * p2[y])); h += pxy * log(pxy); }
    }
  free(cnt);
  if (mi_out) *mi_out = mi;
  if (hjoint_out) *hjoint_out = -h;
}

/* Mica.cpp:341-361: entropy of every site and its average MI with all the others (sum over j != i, j ascending,
 * divided by S - 1).  codes: [T][S] tip-major. */
void orc_mica_sites(int64_t S, int T, const uint8_t* codes, int A, int n_codes, const uint32_t* code_mask,
                    double* entropy, double* average_mi) {
  uint8_t* cols = malloc((size_t)S * T);
  for (int64_t s = 0; s < S; s++)
    for (int t = 0; t < T; t++) cols[s * T + t] = codes[(size_t)t * S + s];
  for (int64_t i = 0; i < S; i++) {
    double sum = 0.;
    for (int64_t j = 0; j < S; j++)
      if (j != i) {
        double mi;
        orc_site_pair(T, cols + i * T, cols + j * T, A, n_codes, code_mask, &mi, NULL);
        sum += mi;
      }
    if (entropy) entropy[i] = orc_site_entropy(T, cols + i * T, A, n_codes, code_mask);
    if (average_mi) average_mi[i] = sum / (double)(S - 1);
  }
  free(cols);
}

/* null.method = permutations: miTest (Mica.cpp:92-118).  mi = MI of the two columns; unless one of them is constant
 * (SiteTools::isConstant(site, true): unknown characters ignored [Bio++ / from memory]), both columns are shuffled
 * and re-scored while count < 5 and i < max_perm, count = #{rep >= mi}; pvalue = (count + 1) / (i + 1),
 * nbPermutations = i.  Upstream shuffles its two copies in place, over and over, with Bio++'s global, unseeded
 * generator.  A uniformly random permutation of any arrangement is a uniformly random arrangement independent of the
 * one it started from, so here -- as on the device -- shuffle i is drawn from the ORIGINAL columns: the same
 * distribution of (count, i), and shuffles that do not depend on each other.  The shuffle used on both sides:
 * inside-out Fisher-Yates (s[0] = c[0]; for k = 1 .. T-1: j = floor(w (k + 1) / 2^32), s[k] = s[j], s[j] = c[k]), the
 * words w taken in order from Philox4x32-10 blocks with counter (pair lo, pair hi, i, column << 24 | block) and
 * key = seed; when T <= 256 a word serves two draws, w <- w (k + 1) mod 2^32 in between (the fraction the first draw
 * left over: bias below 2^-16).  closest (nullable) = the smallest |rep - mi| met, so a test can tell a count that
 * hinges on a rounding tie. */
void orc_mica_permutation_test(int T, const uint8_t* c1, const uint8_t* c2, int A, int n_codes, const uint32_t* code_mask,
                               uint64_t seed, uint64_t pair, int max_perm, double* mi_out, double* pvalue, int* nperm,
                               double* closest) {
  uint32_t full = (A >= 32) ? 0xffffffffu : ((1u << A) - 1u);
  int const1 = 1, const2 = 1, f1 = -1, f2 = -1;
  for (int t = 0; t < T; t++) {
    uint32_t m1 = c1[t] < n_codes ? code_mask[c1[t]] & full : 0, m2 = c2[t] < n_codes ? code_mask[c2[t]] & full : 0;
    if (m1 != full) { if (f1 < 0) f1 = c1[t]; else if (c1[t] != f1) const1 = 0; }
    if (m2 != full) { if (f2 < 0) f2 = c2[t]; else if (c2[t] != f2) const2 = 0; }
  }
  double mi;
  orc_site_pair(T, c1, c2, A, n_codes, code_mask, &mi, NULL);
  if (mi_out) *mi_out = mi;
  if (closest) *closest = INFINITY;
  if (const1 || const2) { *pvalue = 1.; *nperm = 0; return; }
  uint8_t* s1 = malloc(2 * (size_t)T);
  uint8_t* s2 = s1 + T;
  int count = 0, i;
  const int two_per_word = T <= 256;
  for (i = 0; count < 5 && i < max_perm; i++) {
    for (int column = 0; column < 2; column++) {
      uint8_t* col = column ? s2 : s1;
      const uint8_t* src = column ? c2 : c1;
      col[0] = src[0];
      int k = 1;
      for (uint32_t blk = 0; k < T; blk++) {
        uint32_t c[4] = {(uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)i, ((uint32_t)column << 24) | blk};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int u = 0; u < 4; u++) {
          uint32_t w = c[u];
          if (k < T) {
            uint64_t pr = (uint64_t)w * (uint64_t)(k + 1);
            int p = (int)(pr >> 32);
            col[k] = col[p]; col[p] = src[k];
            w = (uint32_t)pr;
            k++;
          }
          if (two_per_word && k < T) {
            int p = (int)(((uint64_t)w * (uint64_t)(k + 1)) >> 32);
            col[k] = col[p]; col[p] = src[k];
            k++;
          }
        }
      }
    }
    double rep;
    orc_site_pair(T, s1, s2, A, n_codes, code_mask, &rep, NULL);
    if (rep >= mi) count++;
    if (closest && fabs(rep - mi) < *closest) *closest = fabs(rep - mi);
  }
  free(s1);
  *pvalue = (double)(count + 1) / (double)(i + 1);
  *nperm = i;
}

/* every pair i < j of an alignment, mica's order; codes: [T][S] tip-major */
void orc_mica_permutations(int64_t S, int T, const uint8_t* codes, int A, int n_codes, const uint32_t* code_mask, uint64_t seed,
                           int max_perm, double* pvalue, int32_t* nperm, double* closest) {
  uint8_t* cols = malloc((size_t)S * T);
  for (int64_t s = 0; s < S; s++)
    for (int t = 0; t < T; t++) cols[s * T + t] = codes[(size_t)t * S + s];
  int64_t idx = 0;
  for (int64_t i = 0; i + 1 < S; i++)
    for (int64_t j = i + 1; j < S; j++, idx++) {
      int np;
      orc_mica_permutation_test(T, cols + i * T, cols + j * T, A, n_codes, code_mask, seed, (uint64_t)idx, max_perm, NULL,
                                pvalue + idx, &np, closest ? closest + idx : NULL);
      nperm[idx] = np;
    }
  free(cols);
}
