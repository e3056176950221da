/* include/bito_gp.h — C-ABI of the B200-native generalized-pruning (GP) engine.
 *
 * This is the drop-in boundary for ONE hot path of phylovi/bito: the GPEngine that
 * executes a GPOperationVector over site-pattern-length partial likelihood vectors
 * (PLVs). Every entry point below replaces a member of the reference's C++ class
 * `GPEngine` (/root/reference/src/gp_engine.hpp:24-236); the reference-side shim a bito
 * maintainer would add is shown in INTEGRATION.md. Plain pointers and sizes only: no
 * torch, Eigen or STL types cross this boundary.
 *
 * Conventions
 *  - All functions return 0 on success and non-zero on failure; bito_gp_last_error()
 *    then holds a message in the reference's Failwith style (sugar.hpp:120-130). The
 *    C++/Python wrappers rethrow it as std::runtime_error / RuntimeError.
 *  - Host pointers unless a name says `_device`. PLVs are pattern-contiguous: pattern p
 *    owns 4 consecutive doubles (A,C,G,T), i.e. the column-major 4xP layout of
 *    MmappedNucleotidePLV (mmapped_plv.hpp:14-50).
 *  - Index conventions are the reference's: PLV id =
 *    type * node_count + node_id with type order P, PHatRight, PHatLeft, RHat, RRight,
 *    RLeft (pv_handler.hpp:26-33, 487-490); spare PLV j = 6*node_count + j; edge
 *    ("GPCSP") ids index branch lengths, q and log-likelihood rows alike.
 *  - One host thread per engine; calls on one engine must not overlap (the reference is
 *    not re-entrant either, gp_engine.hpp:355-375).
 *  - Multi-GPU: one process per GPU. Each rank creates an engine over its contiguous
 *    shard of the site patterns and joins a communicator (bito_gp_comm_init). Per-edge
 *    and per-PLV scalars are all-reduced inside bito_gp_process_operations; every getter
 *    that returns a per-edge or whole-alignment scalar returns the GLOBAL value on every
 *    rank. Per-pattern getters return the local shard.
 */
#ifndef BITO_GP_H
#define BITO_GP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BITO_GP_ABI_VERSION 1

#if defined(__GNUC__)
#define BITO_GP_API __attribute__((visibility("default")))
#else
#define BITO_GP_API
#endif

/* GPOperation variant order, /root/reference/src/gp_operation.hpp:163-168. */
enum bito_gp_op_kind {
  BITO_GP_ZERO_PLV = 0,                            /* a=dest                            :29-33   */
  BITO_GP_SET_TO_STATIONARY_DISTRIBUTION = 1,      /* a=dest b=root_gpcsp_idx           :37-45   */
  BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV = 2, /* a=dest b=gpcsp c=src              :49-58   */
  BITO_GP_MULTIPLY = 3,                            /* a=dest b=src1 c=src2              :87-96   */
  BITO_GP_LIKELIHOOD = 4,                          /* a=dest(gpcsp) b=child c=parent    :102-111 */
  BITO_GP_OPTIMIZE_BRANCH_LENGTH = 5,              /* a=leafward b=rootward c=gpcsp     :118-127 */
  BITO_GP_UPDATE_SBN_PROBABILITIES = 6,            /* a=start b=stop                    :136-142 */
  BITO_GP_RESET_MARGINAL_LIKELIHOOD = 7,           /*                                   :61-64   */
  BITO_GP_INCREMENT_MARGINAL_LIKELIHOOD = 8,       /* a=stationary_times_prior b=rootsplit c=p :72-84 */
  BITO_GP_PREP_FOR_MARGINALIZATION = 9             /* a=dest, src_vector = vec[vec_off..+vec_len) :150-159 */
};

/* One flattened GPOperation: six int64, fields in the struct-member order of
 * gp_operation.hpp. An op list is bito_gp_op[n] plus one shared int64 pool for the
 * PrepForMarginalization source vectors. */
typedef struct bito_gp_op {
  int64_t kind;
  int64_t a, b, c;
  int64_t vec_off, vec_len;
} bito_gp_op;

/* Optimization::OptimizationMethod, /root/reference/src/optimization.hpp:28-34. */
enum bito_gp_optimization_method {
  BITO_GP_BRENT_OPTIMIZATION = 0,
  BITO_GP_BRENT_OPTIMIZATION_WITH_GRADIENTS = 1,
  BITO_GP_GRADIENT_ASCENT_OPTIMIZATION = 2,
  BITO_GP_LOGSPACE_GRADIENT_ASCENT_OPTIMIZATION = 3,
  BITO_GP_NEWTON_OPTIMIZATION = 4
};

/* Constructor arguments of GPEngine::GPEngine (gp_engine.hpp:26-29, gp_engine.cpp:9-43).
 * `mmap_file_path` has no counterpart: PLVs live in HBM. Zero-initialise, then fill. */
typedef struct bito_gp_config {
  int32_t abi_version;       /* BITO_GP_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal */
  int64_t taxon_count;
  int64_t pattern_count;     /* site patterns in THIS rank's shard */
  int64_t site_count;        /* alignment length (global), SitePattern::SiteCount() */
  int64_t node_count;        /* GPDAG::NodeCountWithoutDAGRoot() */
  int64_t gpcsp_count;       /* GPDAG::EdgeCountWithLeafSubsplits() */
  double rescaling_threshold;/* GPEngine::default_rescaling_threshold_ = 1e-40 */
  int32_t use_gradients;     /* gp_engine.cpp:660-665 */
  int32_t spare_node_count;  /* 0 -> 16 (pv_handler.hpp:505)  */
  int32_t spare_gpcsp_count; /* 0 -> 3  (gp_engine.hpp:306)   */
  int32_t flags;             /* BITO_GP_FLAG_* */
  int64_t max_device_bytes;  /* 0 -> 90% of free HBM at creation */
} bito_gp_config;

enum {
  BITO_GP_FLAG_NO_CUDA_GRAPHS = 1,   /* launch level by level instead of replaying graphs   */
  BITO_GP_FLAG_NO_FUSION = 2,        /* one kernel per reference op (debug / parity bisect) */
  BITO_GP_FLAG_NO_LOGLIK_MATRIX = 4, /* keep only per-edge sums; GetLogLikelihoodMatrix fails */
  /* The reference's Assert()s on this path (gp_engine.cpp:237-238, 256-257, 283, 585-586) compile
   * away in its Release build (sugar.hpp:103-111). Default: same - violations are only recorded in
   * bito_gp_stats.device_status_bits. With this flag they fail the call, as in a Debug build. */
  BITO_GP_FLAG_STRICT_ASSERTS = 8,
  /* OptimizeBranchLength on small alignments runs each edge's whole 1-D search inside one thread
   * block (no host round trips). This flag forces the round-per-launch scheme used for large ones. */
  BITO_GP_FLAG_NO_ONCHIP_OPTIMIZER = 16
};

typedef struct bito_gp_engine bito_gp_engine;

BITO_GP_API const char* bito_gp_last_error(void);
BITO_GP_API int bito_gp_abi_version(void);

/* ---- lifetime: GPEngine::GPEngine / ~GPEngine ------------------------------------ */
BITO_GP_API int bito_gp_create(const bito_gp_config* config, bito_gp_engine** out);
BITO_GP_API void bito_gp_destroy(bito_gp_engine* e);

/* SitePattern::GetPatterns()/GetWeights() (site_pattern.hpp:28-36) for this shard +
 * GPEngine::InitializePLVsWithSitePatterns (gp_engine.cpp:544-562). symbols is
 * taxon_count x pattern_count row-major; 0..3 = A,C,G,T; 4 = gap/ambiguous. */
BITO_GP_API int bito_gp_set_site_patterns(bito_gp_engine* e, const uint8_t* symbols, const double* weights);
/* Same, from buffers already resident on this engine's device. */
BITO_GP_API int bito_gp_set_site_patterns_device(bito_gp_engine* e, const uint8_t* symbols_device,
                                     const double* weights_device);

/* GPEngine::InitializePriors (gp_engine.cpp:45-58) / SetNullPrior (:60). */
BITO_GP_API int bito_gp_initialize_priors(bito_gp_engine* e, const double* sbn_prior,
                              const double* unconditional_node_probabilities,
                              const double* inverted_sbn_prior);
BITO_GP_API int bito_gp_set_null_prior(bito_gp_engine* e);

/* ---- the hot call: GPEngine::ProcessOperations (gp_engine.cpp:335-339) -------------- */
/* The list is compiled (or found in the program cache) and enqueued on the engine's stream; `ops` / `vec`
 * are not retained. On one GPU without BITO_GP_FLAG_STRICT_ASSERTS the call may return before the device
 * has finished (the caller's next list is hashed and launched meanwhile): every getter and setter is
 * ordered on the same stream and waits as needed, bito_gp_synchronize / bito_gp_get_stats collect the
 * device status word. With a communicator, or with STRICT_ASSERTS, the call waits and fails here. */
BITO_GP_API int bito_gp_process_operations(bito_gp_engine* e, const bito_gp_op* ops, int64_t n_ops,
                               const int64_t* vec, int64_t vec_len);

/* ---- branch lengths / optimiser: gp_engine.cpp:366-380, 656-674 --------------------- */
BITO_GP_API int bito_gp_set_branch_lengths(bito_gp_engine* e, const double* branch_lengths /* gpcsp_count */);
/* Writes [start, start + length) of the padded branch-length vector (spare edges included): what a
 * caller does through the reference's mutable DAGBranchHandler reference, e.g.
 * branch_handler(edge_id) = x in nni_evaluation_engine.cpp:108, 187. */
BITO_GP_API int bito_gp_set_branch_lengths_range(bito_gp_engine* e, int64_t start, int64_t length,
                                     const double* branch_lengths);
BITO_GP_API int bito_gp_set_branch_lengths_to_constant(bito_gp_engine* e, double branch_length);
BITO_GP_API int bito_gp_set_branch_lengths_to_default(bito_gp_engine* e);
BITO_GP_API int bito_gp_get_branch_lengths(bito_gp_engine* e, int64_t start, int64_t length, double* out);
BITO_GP_API int bito_gp_get_branch_length_differences(bito_gp_engine* e, double* out /* gpcsp_count */);
BITO_GP_API int bito_gp_set_optimization_method(bito_gp_engine* e, int method);
BITO_GP_API int bito_gp_use_gradient_optimization(bito_gp_engine* e, int use_gradients);
BITO_GP_API int bito_gp_set_significant_digits_for_optimization(bito_gp_engine* e, int significant_digits);
BITO_GP_API int64_t bito_gp_get_optimization_count(bito_gp_engine* e);
BITO_GP_API int bito_gp_reset_optimization_count(bito_gp_engine* e);
BITO_GP_API int bito_gp_increment_optimization_count(bito_gp_engine* e);
/* GPEngine::LogLikelihoodAndDerivative / ...AndFirstTwoDerivatives (gp_engine.cpp:470-542)
 * at the edge's current branch length. out[0..2] = ll, d ll/dt, d2 ll/dt2. */
BITO_GP_API int bito_gp_log_likelihood_and_derivatives(bito_gp_engine* e, int64_t gpcsp, int64_t rootward,
                                           int64_t leafward, double out[3]);
/* GPEngine::SetTransitionMatrixToHaveBranchLength + GetTransitionMatrix
 * (gp_engine.cpp:341-344), row-major 4x4, computed on the device. */
BITO_GP_API int bito_gp_get_transition_matrix(bito_gp_engine* e, double branch_length, double out[16]);
/* The substitution model as the engine uses it: the four getters of SubstitutionModel that GPEngine reads
 * (/root/reference/src/gp_engine.hpp:366-376, substitution_model.hpp:24-30): GetEigenvectors() and
 * GetInverseEigenvectors() as row-major 4x4, GetEigenvalues(), GetFrequencies(). The reference engine
 * hard-wires JC69Model, and so does bito_gp_create; this call installs any other reversible nucleotide
 * model - the reference's GTRModel / HKYModel eigendecompositions (substitution_model.cpp:79-186) are
 * what tests/test_models_gpu.py passes in. P(t) = V diag(exp(eigenvalues t)) V^-1. Eigenvalues that are
 * bit-equal share one exponential in the branch-length objective; with other than two distinct
 * eigenvalues the optimiser uses its general per-eigenvalue coefficient kernels. Engines of one process
 * that hold different models must not run concurrently (the eigensystem lives in constant memory). */
BITO_GP_API int bito_gp_set_substitution_model(bito_gp_engine* e, const double eigenvectors[16],
                                   const double inverse_eigenvectors[16], const double eigenvalues[4],
                                   const double frequencies[4]);

/* ---- read-back: gp_engine.cpp:413-468 ------------------------------------------------ */
BITO_GP_API int bito_gp_get_log_marginal_likelihood(bito_gp_engine* e, double* out);
BITO_GP_API int bito_gp_get_per_gpcsp_log_likelihoods(bito_gp_engine* e, int64_t start, int64_t length,
                                          double* out);
BITO_GP_API int bito_gp_get_per_gpcsp_components_of_full_log_marginal(bito_gp_engine* e, double* out);
/* Row-major gpcsp_count x pattern_count (local shard), gp_engine.hpp:340-345. */
BITO_GP_API int bito_gp_get_log_likelihood_matrix(bito_gp_engine* e, double* out);
BITO_GP_API int bito_gp_get_per_pattern_log_marginal(bito_gp_engine* e, double* out /* pattern_count */);
BITO_GP_API int bito_gp_get_sbn_parameters(bito_gp_engine* e, double* out /* gpcsp_count */);
BITO_GP_API int bito_gp_set_sbn_parameters(bito_gp_engine* e, const double* q /* gpcsp_count */);
/* GPEngine::GetPLV (gp_engine.hpp:147). PLVs live in HBM, so this is a staged copy, not
 * an Eigen::Ref; writes go through bito_gp_set_plv. out/in: pattern_count x 4. */
BITO_GP_API int bito_gp_get_plv(bito_gp_engine* e, int64_t plv_id, double* out);
BITO_GP_API int bito_gp_set_plv(bito_gp_engine* e, int64_t plv_id, const double* in, int32_t rescaling_count);
/* ---- quartet hybrid marginals: gp_engine.cpp:748-816, quartet_hybrid_request.hpp:11-43 -- */
/* QuartetTip (quartet_hybrid_request.hpp:11-18). */
typedef struct bito_gp_quartet_tip {
  int64_t tip_node_id;
  int64_t plv_idx;
  int64_t gpcsp_idx;
} bito_gp_quartet_tip;
/* GPEngine::CalculateQuartetHybridLikelihoods (gp_engine.cpp:748-808). tips = the request's
 * rootward, sister, rotated and sorted tips concatenated, tip_counts[4] of each. out receives
 * tip_counts[0]*[1]*[2]*[3] log-likelihoods in the reference's loop order (sorted innermost). */
BITO_GP_API int bito_gp_calculate_quartet_hybrid_likelihoods(bito_gp_engine* e, int64_t central_gpcsp_idx,
                                                 const bito_gp_quartet_tip* tips,
                                                 const int32_t tip_counts[4], double* out);
/* GPEngine::ProcessQuartetHybridRequest (gp_engine.cpp:810-816) for n_requests requests in one
 * launch (GPInstance::CalculateHybridMarginals, gp_instance.cpp:408-417, issues one per edge):
 * request r owns tip_counts[4r..4r+3] tips taken consecutively from `tips`; requests that are not
 * fully formed (an empty tip vector) leave their entry untouched. */
BITO_GP_API int bito_gp_process_quartet_hybrid_requests(bito_gp_engine* e, int64_t n_requests,
                                            const int64_t* central_gpcsp_idx,
                                            const int32_t* tip_counts,
                                            const bito_gp_quartet_tip* tips);
/* GPEngine::GetHybridMarginals (gp_engine.cpp:464-466); -inf where none was computed. */
BITO_GP_API int bito_gp_get_hybrid_marginals(bito_gp_engine* e, double* out /* gpcsp_count */);

/* rescaling_counts_ (gp_engine.hpp:317; private in the reference, exposed for parity). */
BITO_GP_API int bito_gp_get_rescaling_counts(bito_gp_engine* e, int32_t* out /* padded_plv_count */);

/* ---- counts: gp_engine.hpp:198-234 --------------------------------------------------- */
BITO_GP_API int64_t bito_gp_get_node_count(bito_gp_engine* e);
BITO_GP_API int64_t bito_gp_get_plv_count(bito_gp_engine* e);        /* 6 * node_count           */
BITO_GP_API int64_t bito_gp_get_padded_plv_count(bito_gp_engine* e); /* 6 * (node_count + spare) */
BITO_GP_API int64_t bito_gp_get_gpcsp_count(bito_gp_engine* e);
BITO_GP_API int64_t bito_gp_get_padded_gpcsp_count(bito_gp_engine* e);
BITO_GP_API int64_t bito_gp_get_site_pattern_count(bito_gp_engine* e);

/* ---- resize / copy (NNI-search caller surface): gp_engine.cpp:64-209, 386-409 -------- */
/* node_reindexer / gpcsp_reindexer: old index -> new index, length = new count, or NULL.
 * explicit_allocation < 0 means "not given". */
BITO_GP_API int bito_gp_grow_plvs(bito_gp_engine* e, int64_t new_node_count, const int64_t* node_reindexer,
                      int64_t explicit_allocation);
BITO_GP_API int bito_gp_grow_gpcsps(bito_gp_engine* e, int64_t new_gpcsp_count,
                        const int64_t* gpcsp_reindexer, int64_t explicit_allocation);
BITO_GP_API int bito_gp_grow_spare_plvs(bito_gp_engine* e, int64_t new_node_spare_count);
BITO_GP_API int bito_gp_grow_spare_gpcsps(bito_gp_engine* e, int64_t new_gpcsp_spare_count);
/* GPEngine::CopyNodeData (gp_engine.cpp:384-390): the unconditional node probability. */
BITO_GP_API int bito_gp_copy_node_data(bito_gp_engine* e, int64_t src_node_idx, int64_t dest_node_idx);
BITO_GP_API int bito_gp_copy_plv_data(bito_gp_engine* e, int64_t src_plv_idx, int64_t dest_plv_idx);
BITO_GP_API int bito_gp_copy_gpcsp_data(bito_gp_engine* e, int64_t src_gpcsp_idx, int64_t dest_gpcsp_idx);

/* ---- multi-GPU: no reference counterpart (the reference is single-process) ------------ */
/* Rank 0 makes an id, the launcher broadcasts the 128 bytes (e.g. torch.distributed),
 * every rank calls comm_init. Uses NCCL (dlopen'ed at call time). */
BITO_GP_API int bito_gp_comm_make_unique_id(uint8_t id[128]);
BITO_GP_API int bito_gp_comm_init(bito_gp_engine* e, int32_t n_ranks, int32_t rank, const uint8_t id[128]);

/* ---- streams, timing, statistics ------------------------------------------------------- */
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = engine's own. */
BITO_GP_API int bito_gp_set_stream(bito_gp_engine* e, void* cuda_stream);
BITO_GP_API int bito_gp_synchronize(bito_gp_engine* e);

typedef struct bito_gp_stats {
  int64_t kernel_launches;      /* kernels launched (or replayed inside graphs) so far   */
  int64_t graph_launches;
  int64_t process_calls;
  int64_t programs_compiled;    /* distinct op lists compiled into level schedules       */
  int64_t levels_last;          /* dependency levels of the last program                 */
  int64_t fused_ops_last;       /* macro-ops after fusion in the last program            */
  int64_t objective_evaluations;/* OptimizeBranchLength f-evals (all edges) so far       */
  int64_t collective_calls;     /* all-reduces issued so far (NCCL or peer memory)       */
  int64_t device_bytes_in_use;  /* PLV slabs + rows + scalars                            */
  int64_t plvs_resident;        /* PLVs that own HBM (the rest are symbolic or zero)     */
  double algorithmic_bytes_last;/* SURVEY 8(d) bytes per local pattern x patterns, last program */
  double last_process_ms;       /* device time of the last process_operations (events)   */
  int64_t device_status_bits;   /* OR of assert violations seen so far (see gp_types.h)  */
  /* How the most recent level of OptimizeBranchLength ops ran: 0 = one launch per objective round
   * over all edges of the level (rho streamed from HBM), 1 = one thread block per edge, 2 = one
   * thread-block cluster per edge (rho in distributed shared memory). The engine picks per level:
   * on-chip searches for levels of few edges, streaming for thousands of edges at 1e5 patterns. */
  int64_t optimizer_scheme;
  int64_t optimizer_cluster_size;    /* blocks per cluster for scheme 2, else 0              */
  int64_t optimizer_cluster_threads; /* threads per block for scheme 2, else 0               */
  int64_t optimizer_edges_in_flight; /* clusters resident on the device at once, scheme 2   */
  int64_t peer_collective_calls;     /* of collective_calls: one-kernel all-reduces over NVLink peer memory */
  int64_t programs_evicted;          /* compiled programs freed: stale after a resize, or least recently used */
  int64_t programs_cached;           /* compiled programs alive now                                          */
  /* Of objective_evaluations (streamed scheme, plain Brent, two-eigenvalue model): the evaluations that
   * cost a pass over the per-pattern coefficients in HBM. The others were answered from the power-sum
   * (Taylor) model of an earlier pass, within a quarter ulp of the objective (gp_types.h, OptPass). */
  int64_t objective_passes;
} bito_gp_stats;
BITO_GP_API int bito_gp_get_stats(bito_gp_engine* e, bito_gp_stats* out);

/* Per-kernel device time, measured with CUDA events on the launching stream around every
 * launch while profiling is on (graphs are bypassed meanwhile). algorithmic_bytes follows
 * compulsory bytes of the ops each launch executed. */
typedef struct bito_gp_kernel_profile {
  char name[32];
  int64_t launches;
  double total_ms;
  double algorithmic_bytes;
} bito_gp_kernel_profile;
BITO_GP_API int bito_gp_set_profiling(bito_gp_engine* e, int on);
BITO_GP_API int bito_gp_reset_kernel_profile(bito_gp_engine* e);
/* Writes up to `capacity` entries, returns the number written in *n_out. */
BITO_GP_API int bito_gp_get_kernel_profile(bito_gp_engine* e, bito_gp_kernel_profile* out, int capacity,
                               int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* BITO_GP_H */
