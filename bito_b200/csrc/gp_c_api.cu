// bito_b200/csrc/gp_c_api.cu — the extern "C" boundary declared in include/bito_gp.h.
// Every function catches C++ exceptions and turns them into a status + message.
#include <cstring>
#include <exception>

#include "gp_engine.h"

using bito_gp::Engine;

struct bito_gp_engine {
  Engine impl;
  explicit bito_gp_engine(const bito_gp_config& cfg) : impl(cfg) {}
};

namespace {
template <typename F>
int Guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    bito_gp::SetLastError(e.what());
    return 1;
  } catch (...) {
    bito_gp::SetLastError("unknown C++ exception");
    return 1;
  }
}
int NullEngine() {
  bito_gp::SetLastError("bito_gp: null engine handle");
  return 1;
}
}  // namespace

#define ENGINE_OR_FAIL(e) \
  if ((e) == nullptr) return NullEngine()

extern "C" {

const char* bito_gp_last_error(void) { return bito_gp::LastError(); }
int bito_gp_abi_version(void) { return BITO_GP_ABI_VERSION; }

int bito_gp_create(const bito_gp_config* config, bito_gp_engine** out) {
  if (config == nullptr || out == nullptr) {
    bito_gp::SetLastError("bito_gp_create: null argument");
    return 1;
  }
  *out = nullptr;
  return Guard([&] { *out = new bito_gp_engine(*config); });
}

void bito_gp_destroy(bito_gp_engine* e) {
  try {
    delete e;
  } catch (...) {
  }
}

int bito_gp_set_site_patterns(bito_gp_engine* e, const uint8_t* symbols, const double* weights) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetSitePatterns(symbols, weights, false); });
}
int bito_gp_set_site_patterns_device(bito_gp_engine* e, const uint8_t* symbols_device,
                                     const double* weights_device) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetSitePatterns(symbols_device, weights_device, true); });
}
int bito_gp_initialize_priors(bito_gp_engine* e, const double* sbn_prior,
                              const double* unconditional_node_probabilities,
                              const double* inverted_sbn_prior) {
  ENGINE_OR_FAIL(e);
  return Guard([&] {
    e->impl.InitializePriors(sbn_prior, unconditional_node_probabilities, inverted_sbn_prior);
  });
}
int bito_gp_set_null_prior(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetNullPrior(); });
}

int bito_gp_process_operations(bito_gp_engine* e, const bito_gp_op* ops, int64_t n_ops,
                               const int64_t* vec, int64_t vec_len) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.ProcessOperations(ops, n_ops, vec, vec_len); });
}

int bito_gp_set_branch_lengths(bito_gp_engine* e, const double* branch_lengths) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetBranchLengths(branch_lengths); });
}
int bito_gp_set_branch_lengths_range(bito_gp_engine* e, int64_t start, int64_t length,
                                     const double* branch_lengths) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetBranchLengthsRange(start, length, branch_lengths); });
}
int bito_gp_set_branch_lengths_to_constant(bito_gp_engine* e, double branch_length) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetBranchLengthsToConstant(branch_length); });
}
int bito_gp_set_branch_lengths_to_default(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetBranchLengthsToDefault(); });
}
int bito_gp_get_branch_lengths(bito_gp_engine* e, int64_t start, int64_t length, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetBranchLengths(start, length, out); });
}
int bito_gp_get_branch_length_differences(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetBranchLengthDifferences(out); });
}
int bito_gp_set_optimization_method(bito_gp_engine* e, int method) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetOptimizationMethod(method); });
}
int bito_gp_use_gradient_optimization(bito_gp_engine* e, int use_gradients) {
  ENGINE_OR_FAIL(e);
  return Guard([&] {
    e->impl.SetOptimizationMethod(use_gradients ? BITO_GP_BRENT_OPTIMIZATION_WITH_GRADIENTS
                                                : BITO_GP_BRENT_OPTIMIZATION);
  });
}
int bito_gp_set_significant_digits_for_optimization(bito_gp_engine* e, int significant_digits) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetSignificantDigits(significant_digits); });
}
int64_t bito_gp_get_optimization_count(bito_gp_engine* e) {
  return e == nullptr ? -1 : e->impl.optimization_count();
}
int bito_gp_reset_optimization_count(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.ResetOptimizationCount(); });
}
int bito_gp_increment_optimization_count(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.IncrementOptimizationCount(); });
}
int bito_gp_log_likelihood_and_derivatives(bito_gp_engine* e, int64_t gpcsp, int64_t rootward,
                                           int64_t leafward, double out[3]) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.LogLikelihoodAndDerivatives(gpcsp, rootward, leafward, out); });
}
int bito_gp_get_transition_matrix(bito_gp_engine* e, double branch_length, double out[16]) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetTransitionMatrix(branch_length, out); });
}

int bito_gp_set_substitution_model(bito_gp_engine* e, const double eigenvectors[16],
                                   const double inverse_eigenvectors[16], const double eigenvalues[4],
                                   const double frequencies[4]) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetSubstitutionModel(eigenvectors, inverse_eigenvectors, eigenvalues, frequencies); });
}

int bito_gp_get_log_marginal_likelihood(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { *out = e->impl.GetLogMarginalLikelihood(); });
}
int bito_gp_get_per_gpcsp_log_likelihoods(bito_gp_engine* e, int64_t start, int64_t length,
                                          double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetPerGpcspLogLikelihoods(start, length, out); });
}
int bito_gp_get_per_gpcsp_components_of_full_log_marginal(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetPerGpcspComponentsOfFullLogMarginal(out); });
}
int bito_gp_get_log_likelihood_matrix(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetLogLikelihoodMatrix(out); });
}
int bito_gp_get_per_pattern_log_marginal(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetPerPatternLogMarginal(out); });
}
int bito_gp_get_sbn_parameters(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetSbnParameters(out); });
}
int bito_gp_set_sbn_parameters(bito_gp_engine* e, const double* q) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetSbnParameters(q); });
}
int bito_gp_get_plv(bito_gp_engine* e, int64_t plv_id, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetPlv(plv_id, out); });
}
int bito_gp_set_plv(bito_gp_engine* e, int64_t plv_id, const double* in, int32_t rescaling_count) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetPlv(plv_id, in, rescaling_count); });
}
int bito_gp_calculate_quartet_hybrid_likelihoods(bito_gp_engine* e, int64_t central_gpcsp_idx,
                                                 const bito_gp_quartet_tip* tips,
                                                 const int32_t tip_counts[4], double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.QuartetHybrid(1, &central_gpcsp_idx, tip_counts, tips, out, false); });
}
int bito_gp_process_quartet_hybrid_requests(bito_gp_engine* e, int64_t n_requests,
                                            const int64_t* central_gpcsp_idx,
                                            const int32_t* tip_counts,
                                            const bito_gp_quartet_tip* tips) {
  ENGINE_OR_FAIL(e);
  return Guard([&] {
    e->impl.QuartetHybrid(n_requests, central_gpcsp_idx, tip_counts, tips, nullptr, true);
  });
}
int bito_gp_get_hybrid_marginals(bito_gp_engine* e, double* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetHybridMarginals(out); });
}
int bito_gp_get_rescaling_counts(bito_gp_engine* e, int32_t* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetRescalingCounts(out); });
}

int64_t bito_gp_get_node_count(bito_gp_engine* e) { return e ? e->impl.node_count() : -1; }
int64_t bito_gp_get_plv_count(bito_gp_engine* e) { return e ? e->impl.plv_count() : -1; }
int64_t bito_gp_get_padded_plv_count(bito_gp_engine* e) {
  return e ? e->impl.padded_plv_count() : -1;
}
int64_t bito_gp_get_gpcsp_count(bito_gp_engine* e) { return e ? e->impl.gpcsp_count() : -1; }
int64_t bito_gp_get_padded_gpcsp_count(bito_gp_engine* e) {
  return e ? e->impl.padded_gpcsp_count() : -1;
}
int64_t bito_gp_get_site_pattern_count(bito_gp_engine* e) {
  return e ? e->impl.pattern_count() : -1;
}

int bito_gp_grow_plvs(bito_gp_engine* e, int64_t new_node_count, const int64_t* node_reindexer,
                      int64_t explicit_allocation) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GrowPlvs(new_node_count, node_reindexer, explicit_allocation); });
}
int bito_gp_grow_gpcsps(bito_gp_engine* e, int64_t new_gpcsp_count,
                        const int64_t* gpcsp_reindexer, int64_t explicit_allocation) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GrowGpcsps(new_gpcsp_count, gpcsp_reindexer, explicit_allocation); });
}
int bito_gp_grow_spare_plvs(bito_gp_engine* e, int64_t new_node_spare_count) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GrowSparePlvs(new_node_spare_count); });
}
int bito_gp_grow_spare_gpcsps(bito_gp_engine* e, int64_t new_gpcsp_spare_count) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GrowSpareGpcsps(new_gpcsp_spare_count); });
}
int bito_gp_copy_node_data(bito_gp_engine* e, int64_t src_node_idx, int64_t dest_node_idx) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.CopyNodeData(src_node_idx, dest_node_idx); });
}
int bito_gp_copy_plv_data(bito_gp_engine* e, int64_t src_plv_idx, int64_t dest_plv_idx) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.CopyPlvData(src_plv_idx, dest_plv_idx); });
}
int bito_gp_copy_gpcsp_data(bito_gp_engine* e, int64_t src_gpcsp_idx, int64_t dest_gpcsp_idx) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.CopyGpcspData(src_gpcsp_idx, dest_gpcsp_idx); });
}

int bito_gp_comm_make_unique_id(uint8_t id[128]) {
  return Guard([&] { bito_gp::MakeNcclUniqueId(id); });
}
int bito_gp_comm_init(bito_gp_engine* e, int32_t n_ranks, int32_t rank, const uint8_t id[128]) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.CommInit(n_ranks, rank, id); });
}

int bito_gp_set_stream(bito_gp_engine* e, void* cuda_stream) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetStream(static_cast<cudaStream_t>(cuda_stream)); });
}
int bito_gp_synchronize(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.Synchronize(); });
}
int bito_gp_get_stats(bito_gp_engine* e, bito_gp_stats* out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.GetStats(out); });
}

int bito_gp_set_profiling(bito_gp_engine* e, int on) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.SetProfiling(on != 0); });
}
int bito_gp_reset_kernel_profile(bito_gp_engine* e) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { e->impl.ResetKernelProfile(); });
}
int bito_gp_get_kernel_profile(bito_gp_engine* e, bito_gp_kernel_profile* out, int capacity,
                               int* n_out) {
  ENGINE_OR_FAIL(e);
  return Guard([&] { *n_out = e->impl.GetKernelProfile(out, capacity); });
}

}  // extern "C"
