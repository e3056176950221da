/* oracle/gp_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's generalized-pruning
 * engine (phylovi/bito GPEngine), written from the behaviour of the reference sources —
 * each function cites the /root/reference/src file:line it follows. It exists so the CUDA
 * engine can be checked on machines where the reference itself (oracle/_ref) is not
 * built, and as a second, independent statement of the algorithm.
 *
 * PARITY PINNED: tests/test_oracle.py checks this file against (a) the reference's own
 * golden values (gp_engine.hpp:382-393, gp_doctest.cpp:119-131, 257-306, 310-346,
 * numerical_utils.hpp:78-113) and (b) outputs of the unmodified reference GPEngine
 * (oracle/_ref, or the fixtures it wrote under tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this. The product path (bito_b200/) must never call into oracle/.
 */
#include "gp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SPARE_NODES 16 /* PartialVectorHandler spare count, pv_handler.hpp:505 */
#define SPARE_EDGES 3  /* gp_engine.hpp:306 */

/* JC69 eigensystem, substitution_model.cpp:20-26 (row-major literals). */
static const double kJcV[4][4] = {
    {1.0, 2.0, 0.0, 0.5}, {1.0, -2.0, 0.5, 0.0}, {1.0, 2.0, 0.0, -0.5}, {1.0, -2.0, -0.5, 0.0}};
static const double kJcVinv[4][4] = {{0.25, 0.25, 0.25, 0.25},
                                     {0.125, -0.125, 0.125, -0.125},
                                     {0.0, 1.0, 0.0, -1.0},
                                     {1.0, 0.0, -1.0, 0.0}};
static const double kJcLambda[4] = {0.0, -1.3333333333333333, -1.3333333333333333,
                                    -1.3333333333333333};
static const double kJcPi[4] = {0.25, 0.25, 0.25, 0.25};
/* The model in use (process-wide, like the reference's one model per engine build): JC69 unless
 * gpo_set_model installed another eigensystem - GTR / HKY as the reference's GTRModel / HKYModel compute
 * them (substitution_model.cpp:79-186), taken from tests/golden/model_*.npz. */
static double kV[4][4] = {
    {1.0, 2.0, 0.0, 0.5}, {1.0, -2.0, 0.5, 0.0}, {1.0, 2.0, 0.0, -0.5}, {1.0, -2.0, -0.5, 0.0}};
static double kVinv[4][4] = {{0.25, 0.25, 0.25, 0.25},
                             {0.125, -0.125, 0.125, -0.125},
                             {0.0, 1.0, 0.0, -1.0},
                             {1.0, 0.0, -1.0, 0.0}};
static double kLambda[4] = {0.0, -1.3333333333333333, -1.3333333333333333, -1.3333333333333333};
static double kPi[4] = {0.25, 0.25, 0.25, 0.25};
void gpo_set_model(const double* v, const double* vinv, const double* lambda, const double* pi) {
  if (v == NULL) { /* back to JC69 */
    memcpy(kV, kJcV, sizeof kV);
    memcpy(kVinv, kJcVinv, sizeof kVinv);
    memcpy(kLambda, kJcLambda, sizeof kLambda);
    memcpy(kPi, kJcPi, sizeof kPi);
    return;
  }
  memcpy(kV, v, sizeof kV);
  memcpy(kVinv, vinv, sizeof kVinv);
  memcpy(kLambda, lambda, sizeof kLambda);
  memcpy(kPi, pi, sizeof kPi);
}

/* dag_branch_handler.hpp:266-295 */
static const double kDefaultBranchLength = 0.1;
static const double kMinLogBranchLength = -13.9;
static const double kMaxLogBranchLength = 1.1;
static const double kDenominatorToleranceForNewton = 1e-10;
static const double kStepSizeForOptimization = 5e-4;
static const double kStepSizeForLogSpaceOptimization = 1.0005;
static const int64_t kMaxIterForOptimization = 1000;
static const double kBranchLengthDifferenceThreshold = 1e-15;

struct gpo {
  int64_t taxa, P, sites, N, E;
  int64_t plv_count, padded_plv_count, padded_edge_count;
  double thr, log_thr;
  double* plv;   /* padded_plv_count x P x 4 (pattern-major: 4 contiguous states) */
  int32_t* counts;
  double *q, *inverted, *uncond;
  double *bl, *diff;
  double* hybrid;   /* hybrid marginal log-likelihoods per edge, -inf until computed (gp_engine.hpp:352) */
  double* ll;       /* padded_edge_count x P, row-major (gp_engine.hpp:345) */
  double* log_marg; /* P */
  double* weights;  /* P */
  double* scratch;  /* P */
  int method, sig_digits;
  int64_t opt_count, fevals;
  int strict; /* 1: the reference's Asserts fire (Debug build); 0: NDEBUG/Release semantics */
};

static char g_err[512];
const char* gpo_last_error(void) { return g_err; }
static int fail(const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return 1;
}

/* ---- numerical_utils.hpp:35-52 --------------------------------------------------- */
double gpo_log_add(double x, double y) {
  const double log_eps = log(2.220446049250313e-16);
  if (y > x) {
    double t = x;
    x = y;
    y = t;
  }
  if (x == -INFINITY) return x;
  double neg_diff = y - x;
  if (neg_diff < log_eps) return x;
  return x + log(1.0 + exp(neg_diff));
}

/* ---- gp_engine.cpp:341-358: M = (V * diag(f(lambda, t))) * V^-1 -------------------- */
static void eigen_product(const double d[4], double out[4][4]) {
  double vd[4][4];
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 4; ++k) vd[i][k] = kV[i][k] * d[k];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += vd[i][k] * kVinv[k][j];
      out[i][j] = s;
    }
}
static void transition(double t, double M[4][4]) {
  double d[4];
  for (int k = 0; k < 4; ++k) d[k] = exp(t * kLambda[k]);
  eigen_product(d, M);
}
static void transition_and_derivatives(double t, double M[4][4], double D1[4][4],
                                       double D2[4][4]) {
  double d[4], d1[4], d2[4];
  for (int k = 0; k < 4; ++k) {
    d[k] = exp(t * kLambda[k]);
    d1[k] = kLambda[k] * d[k];
    d2[k] = kLambda[k] * kLambda[k] * d[k];
  }
  eigen_product(d, M);
  eigen_product(d1, D1);
  eigen_product(d2, D2);
}
void gpo_transition_matrix(double t, double* out) {
  double M[4][4];
  transition(t, M);
  memcpy(out, M, sizeof M);
}

static inline double* PLV(const gpo* g, int64_t id) { return g->plv + id * g->P * 4; }

/* r^T M p for one pattern. */
static inline double quad(const double* r, const double M[4][4], const double* p) {
  double s = 0.0;
  for (int j = 0; j < 4; ++j) {
    double rm = 0.0;
    for (int i = 0; i < 4; ++i) rm += r[i] * M[i][j];
    s += rm * p[j];
  }
  return s;
}

/* ---- construction: gp_engine.cpp:9-43, 544-562 ------------------------------------- */
gpo* gpo_create(int64_t taxa, int64_t P, const uint8_t* symbols, const double* weights,
                int64_t site_count, int64_t node_count, int64_t edge_count, double thr,
                const double* sbn_prior, const double* uncond, const double* inverted,
                int use_gradients) {
  gpo* g = (gpo*)calloc(1, sizeof(gpo));
  g->taxa = taxa;
  g->P = P;
  g->sites = site_count;
  g->N = node_count;
  g->E = edge_count;
  g->plv_count = 6 * node_count;
  g->padded_plv_count = 6 * (node_count + SPARE_NODES);
  g->padded_edge_count = edge_count + SPARE_EDGES;
  g->thr = thr;
  g->log_thr = log(thr);
  g->plv = (double*)calloc((size_t)(g->padded_plv_count * P * 4), sizeof(double));
  g->counts = (int32_t*)calloc((size_t)g->padded_plv_count, sizeof(int32_t));
  g->q = (double*)malloc(sizeof(double) * g->padded_edge_count);
  g->inverted = (double*)malloc(sizeof(double) * g->padded_edge_count);
  g->uncond = (double*)malloc(sizeof(double) * (node_count + SPARE_NODES));
  g->bl = (double*)malloc(sizeof(double) * g->padded_edge_count);
  g->diff = (double*)calloc((size_t)g->padded_edge_count, sizeof(double));
  g->hybrid = (double*)malloc(sizeof(double) * g->padded_edge_count);
  g->ll = (double*)calloc((size_t)(g->padded_edge_count * P), sizeof(double));
  g->log_marg = (double*)malloc(sizeof(double) * P);
  g->weights = (double*)malloc(sizeof(double) * P);
  g->scratch = (double*)malloc(sizeof(double) * P);
  if (!g->plv || !g->ll) {
    fail("gpo_create: out of memory");
    gpo_destroy(g);
    return NULL;
  }
  for (int64_t i = 0; i < g->padded_edge_count; ++i) {
    g->q[i] = (sbn_prior && i < edge_count) ? sbn_prior[i] : 1.0;
    g->inverted[i] = (inverted && i < edge_count) ? inverted[i] : 1.0;
    g->bl[i] = kDefaultBranchLength;
    g->hybrid[i] = -INFINITY;
  }
  for (int64_t i = 0; i < node_count + SPARE_NODES; ++i)
    g->uncond[i] = (uncond && i < node_count) ? uncond[i] : 1.0;
  memcpy(g->weights, weights, sizeof(double) * P);
  for (int64_t p = 0; p < P; ++p) g->log_marg[p] = -INFINITY;
  /* InitializePLVsWithSitePatterns: one-hot, gap (symbol 4) = all ones. */
  for (int64_t t = 0; t < taxa; ++t) {
    double* plv = PLV(g, t);
    for (int64_t p = 0; p < P; ++p) {
      const int s = symbols[t * P + p];
      if (s == 4) {
        for (int i = 0; i < 4; ++i) plv[4 * p + i] = 1.0;
      } else if (s < 4) {
        plv[4 * p + s] = 1.0;
      }
    }
  }
  g->sig_digits = 10;
  g->method = use_gradients ? GPO_BRENT_WITH_GRADIENTS : GPO_BRENT; /* gp_engine.cpp:660-665 */
  return g;
}

void gpo_destroy(gpo* g) {
  if (!g) return;
  free(g->plv);
  free(g->counts);
  free(g->q);
  free(g->inverted);
  free(g->uncond);
  free(g->bl);
  free(g->diff);
  free(g->hybrid);
  free(g->ll);
  free(g->log_marg);
  free(g->weights);
  free(g->scratch);
  free(g);
}

/* ---- per-pattern log likelihood row: gp_engine.hpp:273-282, gp_engine.cpp:599-601 --- */
static double log_rescaling_for(const gpo* g, int64_t plv) {
  return (double)g->counts[plv] * g->log_thr;
}
static void per_pattern_loglik(const gpo* g, int64_t parent, int64_t child,
                               const double M[4][4], double* out) {
  const double* r = PLV(g, parent);
  const double* p = PLV(g, child);
  const double a = log_rescaling_for(g, parent), b = log_rescaling_for(g, child);
  for (int64_t k = 0; k < g->P; ++k) out[k] = log(quad(r + 4 * k, M, p + 4 * k)) + a + b;
}
static double dot_weights(const gpo* g, const double* v) {
  double s = 0.0;
  for (int64_t k = 0; k < g->P; ++k) s += v[k] * g->weights[k];
  return s;
}

/* gp_engine.cpp:470-542. out = {ll, d1, d2} at the edge's stored branch length. */
void gpo_loglik_and_derivatives(gpo* g, int64_t gpcsp, int64_t rootward, int64_t leafward,
                                double* out) {
  double M[4][4], D1[4][4], D2[4][4];
  transition_and_derivatives(g->bl[gpcsp], M, D1, D2);
  per_pattern_loglik(g, rootward, leafward, M, g->scratch);
  out[0] = dot_weights(g, g->scratch);
  const double* r = PLV(g, rootward);
  const double* p = PLV(g, leafward);
  double d1 = 0.0, d2 = 0.0;
  for (int64_t k = 0; k < g->P; ++k) {
    const double l = quad(r + 4 * k, M, p + 4 * k);
    const double l1 = quad(r + 4 * k, D1, p + 4 * k);
    const double l2 = quad(r + 4 * k, D2, p + 4 * k);
    d1 += (l1 / l) * g->weights[k];
    d2 += ((l2 * l - l1 * l1) / (l * l)) * g->weights[k];
  }
  out[1] = d1;
  out[2] = d2;
}

/* ---- objective functions installed by gp_engine.cpp:603-654 ------------------------- */
typedef struct {
  gpo* g;
  int64_t edge, parent, child;
} edge_ctx;

/* brent_nongrad_func: -ll at exp(log t); does NOT store the branch length. */
static double neg_loglik_at_log(double log_t, void* vctx) {
  edge_ctx* c = (edge_ctx*)vctx;
  double M[4][4];
  transition(exp(log_t), M);
  per_pattern_loglik(c->g, c->parent, c->child, M, c->g->scratch);
  c->g->fevals++;
  return -dot_weights(c->g, c->g->scratch);
}
/* brent_grad_func: stores t, returns (-ll, -t * dll/dt). */
static void neg_loglik_and_grad_at_log(double log_t, edge_ctx* c, double* f, double* df) {
  double out[3];
  const double t = exp(log_t);
  c->g->bl[c->edge] = t;
  gpo_loglik_and_derivatives(c->g, c->edge, c->parent, c->child, out);
  c->g->fevals++;
  *f = -out[0];
  *df = -t * out[1];
}

/* ---- optimization.hpp:71-188 (Boost-derived Brent with an initial guess) ------------- */
void gpo_brent_minimize(gpo_func f, void* ctx, double guess, double min, double max,
                        int significant_digits, int64_t max_iter, double* x_out,
                        double* fx_out) {
  const double tolerance = ldexp(1.0, 1 - significant_digits);
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  const double golden = 0.3819660f; /* float literal, as in the reference */
  w = v = x = guess;
  fw = fv = fx = f(x, ctx);
  delta2 = delta = 0;
  int64_t count = max_iter;
  do {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    int use_bisection = 1;
    if (fabs(delta2) > fract1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double p = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) p = -p;
      q = fabs(q);
      double td = delta2;
      delta2 = delta;
      if (!(fabs(p) >= fabs(q * td / 2)) && !(p <= q * (min - x)) && !(p >= q * (max - x))) {
        delta = p / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2))
          delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
        use_bisection = 0;
      }
    }
    if (use_bisection) {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta)
                                : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    fu = f(u, ctx);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (w == x)) {
        v = w; w = u;
        fv = fw; fw = fu;
      } else if ((fu <= fv) || (v == x) || (v == w)) {
        v = u;
        fv = fu;
      }
    }
  } while (--count);
  *x_out = x;
  *fx_out = fx;
}

/* ---- optimization.hpp:190-329 ------------------------------------------------------- */
static void brent_minimize_with_gradients(edge_ctx* c, double guess, double min, double max,
                                          int significant_digits, int64_t max_iter,
                                          double step_size, double* x_out, double* fx_out) {
  const double tolerance = ldexp(1.0, 1 - significant_digits);
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2, dummy;
  const double golden = 0.3819660f;
  w = v = x = guess;
  neg_loglik_and_grad_at_log(x, c, &fx, &dummy);
  fw = fv = fx;
  delta2 = delta = 0;
  int64_t count = max_iter;
  do {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    int use_bisection = 1;
    if (fabs(delta2) > fract1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double p = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) p = -p;
      q = fabs(q);
      double td = delta2;
      delta2 = delta;
      if (!(fabs(p) >= fabs(q * td / 2)) && !(p <= q * (min - x)) && !(p >= q * (max - x))) {
        delta = p / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2))
          delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
        use_bisection = 0;
      }
    }
    if (use_bisection) {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta)
                                : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    neg_loglik_and_grad_at_log(u, c, &fu, &dummy);
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      double f_prime_x, fu_;
      neg_loglik_and_grad_at_log(x, c, &dummy, &f_prime_x);
      const double u_ = x - step_size * f_prime_x;
      neg_loglik_and_grad_at_log(u_, c, &fu_, &dummy);
      if (fu_ <= fx) {
        if (u_ >= x) min = x; else max = x;
        v = w; w = x; x = u_;
        fv = fw; fw = fx; fx = fu_;
      } else {
        if (u < x) min = u; else max = u;
        if ((fu <= fw) || (w == x)) {
          v = w; w = u;
          fv = fw; fw = fu;
        } else if ((fu <= fv) || (v == x) || (v == w)) {
          v = u;
          fv = fu;
        }
      }
    }
  } while (--count);
  *x_out = x;
  *fx_out = fx;
}

/* ---- dag_branch_handler.cpp:123-280 ------------------------------------------------- */
static void optimize_branch_length(gpo* g, int64_t edge, int64_t parent, int64_t child) {
  const int check_convergence = g->opt_count != 0; /* !IsFirstOptimization() */
  if (check_convergence && g->diff[edge] < kBranchLengthDifferenceThreshold) return;
  edge_ctx c = {g, edge, parent, child};
  double out[3];
  switch (g->method) {
    case GPO_BRENT: {
      const double cur_log = log(g->bl[edge]);
      const double cur_nll = neg_loglik_at_log(cur_log, &c);
      double x, fx;
      gpo_brent_minimize(neg_loglik_at_log, &c, cur_log, kMinLogBranchLength,
                         kMaxLogBranchLength, g->sig_digits, kMaxIterForOptimization, &x, &fx);
      g->bl[edge] = (fx > cur_nll) ? exp(cur_log) : exp(x);
      g->diff[edge] = fabs(exp(cur_log) - g->bl[edge]);
      break;
    }
    case GPO_BRENT_WITH_GRADIENTS: {
      const double cur_log = log(g->bl[edge]);
      double cur_nll, dummy, x, fx;
      neg_loglik_and_grad_at_log(cur_log, &c, &cur_nll, &dummy);
      brent_minimize_with_gradients(&c, cur_log, kMinLogBranchLength, kMaxLogBranchLength,
                                    g->sig_digits, kMaxIterForOptimization,
                                    kStepSizeForLogSpaceOptimization, &x, &fx);
      g->bl[edge] = (fx > cur_nll) ? exp(cur_log) : exp(x);
      g->diff[edge] = fabs(exp(cur_log) - g->bl[edge]);
      break;
    }
    case GPO_GRADIENT_ASCENT: { /* optimization.hpp:331-345 */
      const double start = g->bl[edge];
      const double tolerance = pow(10, -g->sig_digits);
      double x = start;
      int64_t iter = 0;
      for (;;) {
        g->bl[edge] = x;
        gpo_loglik_and_derivatives(g, edge, parent, child, out);
        g->fevals++;
        const double new_x = x + out[1] * kStepSizeForOptimization;
        x = new_x > kMinLogBranchLength ? new_x : kMinLogBranchLength; /* sic: min_x */
        if (fabs(out[1]) < fabs(out[0]) * tolerance || iter >= kMaxIterForOptimization) break;
        ++iter;
      }
      g->bl[edge] = x;
      g->diff[edge] = fabs(start - g->bl[edge]);
      break;
    }
    case GPO_LOGSPACE_GRADIENT_ASCENT: { /* optimization.hpp:347-365 */
      const double start = g->bl[edge];
      const double tolerance = pow(10, -g->sig_digits);
      const double min_x = exp(kMinLogBranchLength);
      double x = start;
      int64_t iter = 0;
      for (;;) {
        const double y = log(x);
        g->bl[edge] = x;
        gpo_loglik_and_derivatives(g, edge, parent, child, out);
        g->fevals++;
        const double new_y = y + (x * out[1]) * kStepSizeForLogSpaceOptimization;
        const double new_x = exp(new_y);
        x = new_x > min_x ? new_x : min_x;
        if (fabs(out[1]) < fabs(out[0]) * tolerance || iter >= kMaxIterForOptimization) break;
        ++iter;
      }
      g->bl[edge] = x;
      g->diff[edge] = fabs(start - g->bl[edge]);
      break;
    }
    case GPO_NEWTON: { /* optimization.hpp:367-402 + gp_engine.cpp:641-653 */
      const double start = g->bl[edge];
      const double tolerance = pow(10, -g->sig_digits);
      const double min_x = kMinLogBranchLength, max_x = kMaxLogBranchLength;
      double x = log(start), new_x, delta;
      int64_t iter = 0;
      for (;;) {
        const double t = exp(x);
        g->bl[edge] = t;
        gpo_loglik_and_derivatives(g, edge, parent, child, out);
        g->fevals++;
        const double f_x = out[0];
        const double f_prime_y = t * out[1];
        const double f_double_prime_y = f_prime_y + pow(t, 2) * out[2];
        if (fabs(f_double_prime_y) < kDenominatorToleranceForNewton) break;
        new_x = x - f_prime_y / f_double_prime_y;
        if (new_x < min_x) new_x = x - 0.5 * (x - min_x);
        if (new_x > max_x) new_x = x - 0.5 * (x - max_x);
        delta = fabs(x - new_x);
        if (delta < tolerance || fabs(f_prime_y) < fabs(f_x) * tolerance ||
            iter == kMaxIterForOptimization)
          break;
        x = new_x;
        ++iter;
      }
      g->bl[edge] = exp(x);
      g->diff[edge] = fabs(start - g->bl[edge]);
      break;
    }
    default:
      break;
  }
}

/* ---- the ten operations: gp_engine.cpp:213-333 --------------------------------------- */
static int check_plv(const gpo* g, int64_t id) {
  return (id >= 0 && id < g->padded_plv_count) ? 0 : fail("PLV index out of range");
}
static int check_edge(const gpo* g, int64_t id) {
  return (id >= 0 && id < g->padded_edge_count) ? 0 : fail("GPCSP index out of range");
}

int gpo_run(gpo* g, const int64_t* ops, int64_t n, const int64_t* vec) {
  const int64_t P = g->P;
  for (int64_t o = 0; o < n; ++o) {
    const int64_t* r = ops + 6 * o;
    const int64_t a = r[1], b = r[2], c = r[3];
    switch (r[0]) {
      case 0: { /* ZeroPLV :213-216 */
        if (check_plv(g, a)) return 1;
        memset(PLV(g, a), 0, sizeof(double) * 4 * P);
        g->counts[a] = 0;
        break;
      }
      case 1: { /* SetToStationaryDistribution :218-227 */
        if (check_plv(g, a) || check_edge(g, b)) return 1;
        double* plv = PLV(g, a);
        for (int64_t k = 0; k < P; ++k)
          for (int i = 0; i < 4; ++i) plv[4 * k + i] = g->q[b] * kPi[i];
        g->counts[a] = 0;
        break;
      }
      case 2: { /* IncrementWithWeightedEvolvedPLV :229-249 (a=dest, b=gpcsp, c=src) */
        if (check_plv(g, a) || check_plv(g, c) || check_edge(g, b)) return 1;
        double M[4][4];
        transition(g->bl[b], M);
        const int diff = g->counts[c] - g->counts[a];
        if (diff < 0 && g->strict)
          return fail("dest_ rescaling too large in IncrementWithWeightedEvolvedPLV");
        const double factor = diff == 0 ? 1. : pow(g->thr, (double)diff);
        const double scale = factor * g->q[b];
        double* dest = PLV(g, a);
        const double* src = PLV(g, c);
        for (int64_t k = 0; k < P; ++k)
          for (int i = 0; i < 4; ++i) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += (scale * M[i][j]) * src[4 * k + j];
            dest[4 * k + i] += s;
          }
        break;
      }
      case 3: { /* Multiply :278-285 + RescalePLVIfNeeded :583-597 (a=dest, b=src1, c=src2) */
        if (check_plv(g, a) || check_plv(g, b) || check_plv(g, c)) return 1;
        double* dest = PLV(g, a);
        const double *s1 = PLV(g, b), *s2 = PLV(g, c);
        double mx = -INFINITY, mn = INFINITY;
        int finite = 1;
        for (int64_t k = 0; k < 4 * P; ++k) {
          const double v = s1[k] * s2[k];
          dest[k] = v;
          if (!isfinite(v)) finite = 0;
          if (v > mx) mx = v;
          if (v < mn) mn = v;
        }
        g->counts[a] = g->counts[b] + g->counts[c];
        if (!finite && g->strict) return fail("Multiply dest_ is not finite");
        if (mn < 0. && g->strict)
          return fail("PLV with negative entry passed to RescalePLVIfNeeded");
        if (mx == 0) break;
        int k = 0;
        while (mx < g->thr) {
          mx /= g->thr;
          k++;
        }
        if (k != 0) { /* RescalePLV :564-573 */
          const double d = pow(g->thr, (double)k);
          for (int64_t i = 0; i < 4 * P; ++i) dest[i] /= d;
          g->counts[a] += k;
        }
        break;
      }
      case 4: { /* Likelihood :287-291 (a=dest gpcsp, b=child, c=parent) */
        if (check_edge(g, a) || check_plv(g, b) || check_plv(g, c)) return 1;
        double M[4][4];
        transition(g->bl[a], M);
        per_pattern_loglik(g, c, b, M, g->ll + a * P);
        break;
      }
      case 5: { /* OptimizeBranchLength :293-295, 667-670 (a=leafward, b=rootward, c=gpcsp) */
        if (check_plv(g, a) || check_plv(g, b) || check_edge(g, c)) return 1;
        optimize_branch_length(g, c, b, a);
        break;
      }
      case 6: { /* UpdateSBNProbabilities :297-321 */
        const int64_t len = b - a;
        if (check_edge(g, a) || len < 1 || check_edge(g, b - 1)) return 1;
        if (len == 1) {
          g->q[a] = 1.;
        } else {
          double* x = (double*)malloc(sizeof(double) * len);
          double hybrid_min = INFINITY; /* :309-316: hybrid marginals win when all are set */
          for (int64_t i = 0; i < len; ++i)
            if (g->hybrid[a + i] < hybrid_min) hybrid_min = g->hybrid[a + i];
          for (int64_t i = 0; i < len; ++i)
            x[i] = (hybrid_min > -INFINITY ? g->hybrid[a + i] : dot_weights(g, g->ll + (a + i) * P)) +
                   log(g->q[a + i]);
          double norm = x[0]; /* NumericalUtils::LogSum = left fold, numerical_utils.cpp:8 */
          for (int64_t i = 1; i < len; ++i) norm = gpo_log_add(norm, x[i]);
          for (int64_t i = 0; i < len; ++i) g->q[a + i] = exp(x[i] - norm);
          free(x);
        }
        break;
      }
      case 7: /* ResetMarginalLikelihood :251-253 */
        for (int64_t k = 0; k < P; ++k) g->log_marg[k] = -INFINITY;
        break;
      case 8: { /* IncrementMarginalLikelihood :255-276 (a=stationary*prior, b=rootsplit, c=p) */
        if (check_plv(g, a) || check_edge(g, b) || check_plv(g, c)) return 1;
        if (g->counts[a] != 0 && g->strict)
          return fail("Surprise! Rescaled stationary distribution in IncrementMarginalLikelihood");
        const double *st = PLV(g, a), *pp = PLV(g, c);
        const double resc = log_rescaling_for(g, c), logq = log(g->q[b]);
        double* row = g->ll + b * P;
        for (int64_t k = 0; k < P; ++k) {
          double s = 0.0;
          for (int i = 0; i < 4; ++i) s += st[4 * k + i] * pp[4 * k + i];
          row[k] = log(s) + resc;
          g->log_marg[k] = gpo_log_add(g->log_marg[k], row[k]);
          row[k] -= logq;
        }
        break;
      }
      case 9: { /* PrepForMarginalization :323-333 */
        if (check_plv(g, a)) return 1;
        if (r[5] <= 0) return fail("Empty src_vector in PrepForMarginalization");
        int32_t m = INT32_MAX;
        for (int64_t i = 0; i < r[5]; ++i) {
          const int64_t s = vec[r[4] + i];
          if (check_plv(g, s)) return 1;
          if (g->counts[s] < m) m = g->counts[s];
        }
        g->counts[a] = m;
        break;
      }
      default:
        return fail("unknown GPOperation kind");
    }
  }
  return 0;
}

/* ---- getters: gp_engine.cpp:413-468 --------------------------------------------------- */
/* ---- quartet hybrid marginals: gp_engine.cpp:748-816 ---------------------------------- */
static void matvec4(const double M[4][4], const double* x, double* y) {
  for (int i = 0; i < 4; ++i) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += M[i][k] * x[k];
    y[i] = s;
  }
}

/* CalculateQuartetHybridLikelihoods (:748-808): tips = (tip_node_id, plv_idx, gpcsp_idx)
 * triples of the rootward, sister, rotated and sorted tips, counts[4] of each; out receives
 * one log-likelihood per (rootward, sister, rotated, sorted) choice, sorted innermost. */
int gpo_quartet_likelihoods(gpo* g, int64_t central, const int32_t* counts, const int64_t* tips,
                            double* out) {
  const int64_t* rw = tips;
  const int64_t* sis = rw + 3 * counts[0];
  const int64_t* rot = sis + 3 * counts[1];
  const int64_t* srt = rot + 3 * counts[2];
  const int64_t n_tips = (int64_t)counts[0] + counts[1] + counts[2] + counts[3];
  if (check_edge(g, central)) return 1;
  for (int64_t i = 0; i < n_tips; ++i) {
    if (check_plv(g, tips[3 * i + 1]) || check_edge(g, tips[3 * i + 2])) return 1;
    if (g->strict && g->counts[tips[3 * i + 1]] != 0)
      return fail("Rescaling not implemented in CalculateQuartetHybridLikelihoods.");
  }
  double Mrw[4][4], Ms[4][4], Mc[4][4], Mrot[4][4], Msrt[4][4];
  int64_t k = 0;
  for (int a = 0; a < counts[0]; ++a) {
    const double log_prior = log(g->uncond[rw[3 * a]]);
    transition(g->bl[rw[3 * a + 2]], Mrw);
    for (int b = 0; b < counts[1]; ++b) {
      transition(g->bl[sis[3 * b + 2]], Ms);
      transition(g->bl[central], Mc);
      for (int c = 0; c < counts[2]; ++c) {
        transition(g->bl[rot[3 * c + 2]], Mrot);
        for (int d = 0; d < counts[3]; ++d, ++k) {
          const double non_seq = log(g->inverted[rw[3 * a + 2]] * g->q[sis[3 * b + 2]] *
                                     g->q[rot[3 * c + 2]] * g->q[srt[3 * d + 2]]);
          transition(g->bl[srt[3 * d + 2]], Msrt);
          const double *x_rw = PLV(g, rw[3 * a + 1]), *x_s = PLV(g, sis[3 * b + 1]);
          const double *x_rot = PLV(g, rot[3 * c + 1]), *x_srt = PLV(g, srt[3 * d + 1]);
          for (int64_t p = 0; p < g->P; ++p) {
            double root[4], s[4], r_s[4], q_s[4], r[4], r_sorted[4];
            matvec4(Mrw, x_rw + 4 * p, root);
            matvec4(Ms, x_s + 4 * p, s);
            for (int i = 0; i < 4; ++i) r_s[i] = root[i] * s[i];
            matvec4(Mc, r_s, q_s);
            matvec4(Mrot, x_rot + 4 * p, r);
            for (int i = 0; i < 4; ++i) r_sorted[i] = q_s[i] * r[i];
            g->scratch[p] = log(quad(r_sorted, Msrt, x_srt + 4 * p)) - log_prior;
          }
          out[k] = non_seq + dot_weights(g, g->scratch);
        }
      }
    }
  }
  return 0;
}

/* ProcessQuartetHybridRequest (:810-816): LogSum (left fold of LogAdd, numerical_utils.cpp)
 * of a fully formed request's summands goes to hybrid_marginal_log_likelihoods_[central]. */
int gpo_process_quartet_requests(gpo* g, int64_t n, const int64_t* central, const int32_t* counts,
                                 const int64_t* tips) {
  const int64_t* t = tips;
  for (int64_t r = 0; r < n; ++r) {
    const int32_t* c = counts + 4 * r;
    const int64_t n_out = (int64_t)c[0] * c[1] * c[2] * c[3];
    if (n_out > 0) {
      double* out = (double*)malloc(sizeof(double) * (size_t)n_out);
      if (gpo_quartet_likelihoods(g, central[r], c, t, out) != 0) {
        free(out);
        return 1;
      }
      double acc = out[0];
      for (int64_t i = 1; i < n_out; ++i) acc = gpo_log_add(acc, out[i]);
      g->hybrid[central[r]] = acc;
      free(out);
    }
    t += 3 * ((int64_t)c[0] + c[1] + c[2] + c[3]);
  }
  return 0;
}

void gpo_get_hybrid_marginals(const gpo* g, double* out) {
  memcpy(out, g->hybrid, sizeof(double) * g->E);
}

int64_t gpo_plv_count(const gpo* g) { return g->plv_count; }
int64_t gpo_padded_plv_count(const gpo* g) { return g->padded_plv_count; }
int gpo_get_plv(const gpo* g, int64_t id, double* out) {
  if (check_plv(g, id)) return 1;
  memcpy(out, PLV(g, id), sizeof(double) * 4 * g->P);
  return 0;
}
int gpo_set_plv(gpo* g, int64_t id, const double* in, int32_t count) {
  if (check_plv(g, id)) return 1;
  memcpy(PLV(g, id), in, sizeof(double) * 4 * g->P);
  g->counts[id] = count;
  return 0;
}
void gpo_get_counts(const gpo* g, int32_t* out) {
  memcpy(out, g->counts, sizeof(int32_t) * g->padded_plv_count);
}
void gpo_get_loglik_matrix(const gpo* g, double* out) {
  memcpy(out, g->ll, sizeof(double) * g->E * g->P);
}
void gpo_get_per_pattern_marginal(const gpo* g, double* out) {
  memcpy(out, g->log_marg, sizeof(double) * g->P);
}
void gpo_get_per_gpcsp_loglik(const gpo* g, double* out) {
  for (int64_t e = 0; e < g->E; ++e) out[e] = dot_weights(g, g->ll + e * g->P);
}
void gpo_get_per_gpcsp_components(const gpo* g, double* out) {
  for (int64_t e = 0; e < g->E; ++e)
    out[e] = dot_weights(g, g->ll + e * g->P) + (double)g->sites * log(g->q[e]);
}
double gpo_get_log_marginal(const gpo* g) { return dot_weights(g, g->log_marg); }
void gpo_get_q(const gpo* g, double* out) { memcpy(out, g->q, sizeof(double) * g->E); }
void gpo_set_q(gpo* g, const double* in) { memcpy(g->q, in, sizeof(double) * g->E); }
void gpo_get_branch_lengths(const gpo* g, double* out) {
  memcpy(out, g->bl, sizeof(double) * g->E);
}
void gpo_set_branch_lengths(gpo* g, const double* in) {
  memcpy(g->bl, in, sizeof(double) * g->E);
}
void gpo_set_branch_lengths_constant(gpo* g, double t) {
  for (int64_t i = 0; i < g->padded_edge_count; ++i) g->bl[i] = t;
}
void gpo_get_branch_differences(const gpo* g, double* out) {
  memcpy(out, g->diff, sizeof(double) * g->E);
}
void gpo_set_optimization_method(gpo* g, int method) { g->method = method; }
void gpo_use_gradient_optimization(gpo* g, int use) {
  g->method = use ? GPO_BRENT_WITH_GRADIENTS : GPO_BRENT;
}
void gpo_set_significant_digits(gpo* g, int digits) { g->sig_digits = digits; }
void gpo_reset_optimization_count(gpo* g) { /* dag_branch_handler.hpp:49-52 */
  g->opt_count = 0;
  for (int64_t i = 0; i < g->padded_edge_count; ++i) g->diff[i] = 0.0;
}
void gpo_increment_optimization_count(gpo* g) { g->opt_count++; }
int64_t gpo_get_optimization_count(const gpo* g) { return g->opt_count; }
void gpo_set_null_prior(gpo* g) {
  for (int64_t i = 0; i < g->padded_edge_count; ++i) g->q[i] = 1.0;
}
int64_t gpo_feval_count(const gpo* g) { return g->fevals; }
void gpo_set_strict(gpo* g, int strict) { g->strict = strict; }
