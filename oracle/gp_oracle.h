/* oracle/gp_oracle.h — TEST INFRASTRUCTURE ONLY (see gp_oracle.c). */
#ifndef GP_ORACLE_H
#define GP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpo gpo;

/* Optimization::OptimizationMethod order, /root/reference/src/optimization.hpp:28-34 */
enum {
  GPO_BRENT = 0,
  GPO_BRENT_WITH_GRADIENTS = 1,
  GPO_GRADIENT_ASCENT = 2,
  GPO_LOGSPACE_GRADIENT_ASCENT = 3,
  GPO_NEWTON = 4
};

/* symbols: taxa x patterns row-major, 0..3 = ACGT, 4 = gap. Priors may be NULL (-> 1.0). */
gpo* gpo_create(int64_t taxon_count, int64_t pattern_count, const uint8_t* symbols,
                const double* weights, int64_t site_count, int64_t node_count,
                int64_t edge_count, double rescaling_threshold, const double* sbn_prior,
                const double* unconditional_node_probabilities,
                const double* inverted_sbn_prior, int use_gradients);
void gpo_destroy(gpo* g);
const char* gpo_last_error(void);

/* ops: int64[n][6] = {kind, a, b, c, vec_off, vec_len}; see include/bito_gp.h. */
int gpo_run(gpo* g, const int64_t* ops, int64_t n, const int64_t* vec);

int64_t gpo_plv_count(const gpo* g);        /* 6 * node_count              */
int64_t gpo_padded_plv_count(const gpo* g); /* 6 * (node_count + 16 spare) */
int gpo_get_plv(const gpo* g, int64_t plv_id, double* out /* P*4 */);
int gpo_set_plv(gpo* g, int64_t plv_id, const double* in /* P*4 */, int32_t count);
void gpo_get_counts(const gpo* g, int32_t* out /* padded plv count */);
void gpo_get_loglik_matrix(const gpo* g, double* out /* E*P */);
void gpo_get_per_pattern_marginal(const gpo* g, double* out /* P */);
void gpo_get_per_gpcsp_loglik(const gpo* g, double* out /* E */);
void gpo_get_per_gpcsp_components(const gpo* g, double* out /* E */);
double gpo_get_log_marginal(const gpo* g);
void gpo_get_q(const gpo* g, double* out);
void gpo_set_q(gpo* g, const double* in);
void gpo_get_branch_lengths(const gpo* g, double* out);
void gpo_set_branch_lengths(gpo* g, const double* in);
void gpo_set_branch_lengths_constant(gpo* g, double t);
void gpo_get_branch_differences(const gpo* g, double* out);
void gpo_set_optimization_method(gpo* g, int method);
void gpo_use_gradient_optimization(gpo* g, int use);
void gpo_set_significant_digits(gpo* g, int digits);
void gpo_reset_optimization_count(gpo* g);
void gpo_increment_optimization_count(gpo* g);
int64_t gpo_get_optimization_count(const gpo* g);
void gpo_set_null_prior(gpo* g);
/* out[0..2] = ll, d1, d2 at the edge's current branch length (gp_engine.cpp:470-542). */
void gpo_loglik_and_derivatives(gpo* g, int64_t gpcsp, int64_t rootward, int64_t leafward,
                                double* out);
void gpo_transition_matrix(double t, double* out /* 4x4 row-major */);
/* Process-wide substitution model: V, V^-1 (4x4 row-major), eigenvalues, frequencies; v == NULL restores JC69. */
void gpo_set_model(const double* v, const double* vinv, const double* lambda, const double* pi);
/* Quartet hybrid marginals (gp_engine.cpp:748-816); tips = (tip_node_id, plv_idx, gpcsp_idx)
 * triples, counts[4] = rootward, sister, rotated, sorted. */
int gpo_quartet_likelihoods(gpo* g, int64_t central, const int32_t* counts, const int64_t* tips,
                            double* out);
int gpo_process_quartet_requests(gpo* g, int64_t n, const int64_t* central, const int32_t* counts,
                                 const int64_t* tips);
void gpo_get_hybrid_marginals(const gpo* g, double* out /* E */);
/* number of objective evaluations performed by OptimizeBranchLength ops so far */
int64_t gpo_feval_count(const gpo* g);
/* strict != 0: the reference's Assert()s (sugar.hpp:103-111) raise errors, as in a Debug build.
 * Default 0 = the Release build (-DNDEBUG) the reference ships, where they compile away. */
void gpo_set_strict(gpo* g, int strict);

/* Scalar helpers, exported so tests can pin them against the reference's goldens. */
double gpo_log_add(double x, double y);
/* Brent on a caller-supplied function (for unit tests of the restated optimiser). */
typedef double (*gpo_func)(double x, void* ctx);
void gpo_brent_minimize(gpo_func f, void* ctx, double guess, double min, double max,
                        int significant_digits, int64_t max_iter, double* x_out,
                        double* fx_out);

#ifdef __cplusplus
}
#endif
#endif
