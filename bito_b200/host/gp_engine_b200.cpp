// bito_b200/host/gp_engine_b200.cpp — see gp_engine_b200.hpp. Every member is a forward to
// the C-ABI of include/bito_gp.h; the only real host code is the flattening of the
// GPOperation variant (gp_operation.hpp:163-170) into bito_gp_op rows and the three
// tree-collection helpers that never touched PLV arithmetic (gp_engine.cpp:676-746).
#include "gp_engine_b200.hpp"

#include <sstream>

#include "sugar.hpp"

using Self = BITO_B200_ENGINE_CLASS;

void Self::Check(int rc) const {
  if (rc != 0) Failwith(bito_gp_last_error());  // sugar.hpp:120-130 -> std::runtime_error
}

namespace {
int64_t I(size_t v) { return static_cast<int64_t>(v); }

// std::variant alternative order == bito_gp_op_kind (gp_operation.hpp:163-168).
struct Flatten {
  std::vector<bito_gp_op>& ops;
  std::vector<int64_t>& vec;
  bool changes_branch_lengths = false;  // the list holds an OptimizeBranchLength: the only op that writes them
  void Push(int64_t kind, int64_t a = 0, int64_t b = 0, int64_t c = 0, int64_t off = 0,
            int64_t len = 0) {
    ops.push_back(bito_gp_op{kind, a, b, c, off, len});
  }
  void operator()(const GPOperations::ZeroPLV& o) { Push(BITO_GP_ZERO_PLV, I(o.dest_)); }
  void operator()(const GPOperations::SetToStationaryDistribution& o) {
    Push(BITO_GP_SET_TO_STATIONARY_DISTRIBUTION, I(o.dest_), I(o.root_gpcsp_idx_));
  }
  void operator()(const GPOperations::IncrementWithWeightedEvolvedPLV& o) {
    Push(BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV, I(o.dest_), I(o.gpcsp_), I(o.src_));
  }
  void operator()(const GPOperations::ResetMarginalLikelihood&) {
    Push(BITO_GP_RESET_MARGINAL_LIKELIHOOD);
  }
  void operator()(const GPOperations::IncrementMarginalLikelihood& o) {
    Push(BITO_GP_INCREMENT_MARGINAL_LIKELIHOOD, I(o.stationary_times_prior_), I(o.rootsplit_),
         I(o.p_));
  }
  void operator()(const GPOperations::Multiply& o) {
    Push(BITO_GP_MULTIPLY, I(o.dest_), I(o.src1_), I(o.src2_));
  }
  void operator()(const GPOperations::Likelihood& o) {
    Push(BITO_GP_LIKELIHOOD, I(o.dest_), I(o.child_), I(o.parent_));
  }
  void operator()(const GPOperations::OptimizeBranchLength& o) {
    Push(BITO_GP_OPTIMIZE_BRANCH_LENGTH, I(o.leafward_), I(o.rootward_), I(o.gpcsp_));
    changes_branch_lengths = true;
  }
  void operator()(const GPOperations::UpdateSBNProbabilities& o) {
    Push(BITO_GP_UPDATE_SBN_PROBABILITIES, I(o.start_), I(o.stop_));
  }
  void operator()(const GPOperations::PrepForMarginalization& o) {
    Push(BITO_GP_PREP_FOR_MARGINALIZATION, I(o.dest_), 0, 0, I(vec.size()), I(o.src_vector_.size()));
    for (const size_t s : o.src_vector_) vec.push_back(I(s));
  }
};

void FlattenRequest(const QuartetHybridRequest& request, std::vector<bito_gp_quartet_tip>& tips,
                    std::vector<int32_t>& counts) {
  for (const QuartetTipVector* v : {&request.rootward_tips_, &request.sister_tips_,
                                    &request.rotated_tips_, &request.sorted_tips_}) {
    counts.push_back(static_cast<int32_t>(v->size()));
    for (const QuartetTip& t : *v)
      tips.push_back(bito_gp_quartet_tip{I(t.tip_node_id_), I(t.plv_idx_), I(t.gpcsp_idx_)});
  }
}
}  // namespace

// ---- construction: gp_engine.cpp:9-43 ---------------------------------------------------------
Self::BITO_B200_ENGINE_CLASS(SitePattern site_pattern, size_t node_count, size_t gpcsp_count,
                             const std::string& /*mmap_file_path: PLVs live in HBM*/,
                             double rescaling_threshold, EigenVectorXd sbn_prior,
                             EigenVectorXd unconditional_node_probabilities,
                             EigenVectorXd inverted_sbn_prior, bool use_gradients, int device,
                             int flags)
    : site_pattern_(std::move(site_pattern)) {
  bito_gp_config cfg{};
  cfg.abi_version = BITO_GP_ABI_VERSION;
  cfg.device = device;
  cfg.taxon_count = I(site_pattern_.SequenceCount());
  cfg.pattern_count = I(site_pattern_.PatternCount());
  cfg.site_count = I(site_pattern_.SiteCount());
  cfg.node_count = I(node_count);
  cfg.gpcsp_count = I(gpcsp_count);
  cfg.rescaling_threshold = rescaling_threshold;
  cfg.use_gradients = use_gradients ? 1 : 0;
  cfg.flags = flags;
  Check(bito_gp_create(&cfg, &handle_));
  // SitePattern::GetPatterns(): taxa x patterns of symbols 0..4 (site_pattern.cpp:16-46).
  const auto& patterns = site_pattern_.GetPatterns();
  const size_t P = site_pattern_.PatternCount();
  std::vector<uint8_t> symbols(patterns.size() * P);
  for (size_t t = 0; t < patterns.size(); ++t)
    for (size_t p = 0; p < P; ++p) {
      const auto s = patterns[t][p];
      symbols[t * P + p] = static_cast<uint8_t>(s > 4 ? 4 : s);
    }
  Check(bito_gp_set_site_patterns(H(), symbols.data(), site_pattern_.GetWeights().data()));
  // The reference moves the three vectors in whatever their size (gp_engine.hpp:382-393 builds an
  // engine with empty ones); only full-size priors are meaningful to upload.
  if (size_t(sbn_prior.size()) == gpcsp_count && size_t(inverted_sbn_prior.size()) == gpcsp_count &&
      size_t(unconditional_node_probabilities.size()) == node_count)
    Check(bito_gp_initialize_priors(H(), sbn_prior.data(),
                                    unconditional_node_probabilities.data(),
                                    inverted_sbn_prior.data()));
  transition_matrix_.setZero();
}

Self::~BITO_B200_ENGINE_CLASS() { bito_gp_destroy(handle_); }

void Self::InitializePriors(EigenVectorXd sbn_prior, EigenVectorXd unconditional_node_probabilities,
                            EigenVectorXd inverted_sbn_prior) {
  Assert(size_t(unconditional_node_probabilities.size()) == GetNodeCount(),
         "unconditional_node_probabilities is wrong size for GPEngine.");
  Assert(size_t(sbn_prior.size()) == GetGPCSPCount(), "sbn_prior is wrong size for GPEngine.");
  Assert(size_t(inverted_sbn_prior.size()) == GetGPCSPCount(),
         "inverted_sbn_prior is wrong size for GPEngine.");
  Check(bito_gp_initialize_priors(H(), sbn_prior.data(), unconditional_node_probabilities.data(),
                                  inverted_sbn_prior.data()));
}
void Self::SetNullPrior() { Check(bito_gp_set_null_prior(H())); }

// ---- resize / reindex: gp_engine.cpp:64-209 ---------------------------------------------------
void Self::GrowPLVs(const size_t node_count, std::optional<const Reindexer> node_reindexer,
                    std::optional<const size_t> explicit_allocation, const bool) {
  std::vector<int64_t> idx;
  if (node_reindexer.has_value())
    idx.assign(node_reindexer->GetData().begin(), node_reindexer->GetData().end());
  Check(bito_gp_grow_plvs(H(), I(node_count), node_reindexer.has_value() ? idx.data() : nullptr,
                          explicit_allocation.has_value() ? I(*explicit_allocation) : -1));
}
void Self::GrowGPCSPs(const size_t gpcsp_count, std::optional<const Reindexer> gpcsp_reindexer,
                      std::optional<const size_t> explicit_allocation, const bool) {
  std::vector<int64_t> idx;
  if (gpcsp_reindexer.has_value())
    idx.assign(gpcsp_reindexer->GetData().begin(), gpcsp_reindexer->GetData().end());
  Check(bito_gp_grow_gpcsps(H(), I(gpcsp_count),
                            gpcsp_reindexer.has_value() ? idx.data() : nullptr,
                            explicit_allocation.has_value() ? I(*explicit_allocation) : -1));
  AfterBranchLengthChange();
}
void Self::GrowSparePLVs(const size_t n) { Check(bito_gp_grow_spare_plvs(H(), I(n))); }
void Self::GrowSpareGPCSPs(const size_t n) {
  Check(bito_gp_grow_spare_gpcsps(H(), I(n)));
  AfterBranchLengthChange();
}

// ---- the hot call: gp_engine.cpp:335-339 ------------------------------------------------------
void Self::ProcessOperations(GPOperationVector operations) {
  std::vector<bito_gp_op> ops;
  std::vector<int64_t> vec;
  ops.reserve(operations.size());
  Flatten flatten{ops, vec};
  for (const auto& op : operations) std::visit(flatten, op);
  Check(bito_gp_process_operations(H(), ops.data(), I(ops.size()), vec.data(), I(vec.size())));
  // the DAGBranchHandler mirror is read back (a device -> host copy that waits for the list) only after a
  // list that can have changed branch lengths or differences; a likelihood pass returns as soon as it is queued
  if (flatten.changes_branch_lengths) AfterBranchLengthChange();
}

// ---- optimiser settings -----------------------------------------------------------------------
void Self::SetOptimizationMethod(const OptimizationMethod method) {
  Check(bito_gp_set_optimization_method(H(), static_cast<int>(method)));
}
void Self::UseGradientOptimization(const bool use_gradients) {
  Check(bito_gp_use_gradient_optimization(H(), use_gradients ? 1 : 0));
}
void Self::SetSignificantDigitsForOptimization(int significant_digits) {
  Check(bito_gp_set_significant_digits_for_optimization(H(), significant_digits));
}
size_t Self::GetOptimizationCount() {
  return static_cast<size_t>(bito_gp_get_optimization_count(H()));
}
void Self::ResetOptimizationCount() {
  Check(bito_gp_reset_optimization_count(H()));
  AfterBranchLengthChange();
}
void Self::IncrementOptimizationCount() { Check(bito_gp_increment_optimization_count(H())); }

void Self::SetSubstitutionModel(const SubstitutionModel& model) {
  Assert(model.GetStateCount() == 4, "GPEngine::SetSubstitutionModel needs a nucleotide model.");
  double v[16], vinv[16], lambda[4], pi[4];
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      v[4 * i + j] = model.GetEigenvectors()(i, j);
      vinv[4 * i + j] = model.GetInverseEigenvectors()(i, j);
    }
    lambda[i] = model.GetEigenvalues()[i];
    pi[i] = model.GetFrequencies()[i];
  }
  Check(bito_gp_set_substitution_model(H(), v, vinv, lambda, pi));
}
void Self::SetTransitionMatrixToHaveBranchLength(double branch_length) {
  double m[16];
  Check(bito_gp_get_transition_matrix(H(), branch_length, m));
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) transition_matrix_(i, j) = m[4 * i + j];
}
void Self::SetBranchLengths(EigenVectorXd branch_lengths) {
  Assert(size_t(branch_lengths.size()) == GetGPCSPCount(),
         "Size mismatch in GPEngine::SetBranchLengths.");
  Check(bito_gp_set_branch_lengths(H(), branch_lengths.data()));
  AfterBranchLengthChange();
}
void Self::SetBranchLengthsToConstant(double branch_length) {
  Check(bito_gp_set_branch_lengths_to_constant(H(), branch_length));
  AfterBranchLengthChange();
}
void Self::SetBranchLengthsToDefault() {
  Check(bito_gp_set_branch_lengths_to_default(H()));
  AfterBranchLengthChange();
}
void Self::ResetLogMarginalLikelihood() { (*this)(GPOperations::ResetMarginalLikelihood{}); }

void Self::CopyNodeData(const NodeId src, const NodeId dest) {
  Check(bito_gp_copy_node_data(H(), I(src.value_), I(dest.value_)));
}
void Self::CopyPLVData(const size_t src, const size_t dest) {
  Check(bito_gp_copy_plv_data(H(), I(src), I(dest)));
}
void Self::CopyGPCSPData(const EdgeId src, const EdgeId dest) {
  Check(bito_gp_copy_gpcsp_data(H(), I(src.value_), I(dest.value_)));
  AfterBranchLengthChange();
}

// ---- the DAGBranchHandler mirror ---------------------------------------------------------------
DAGBranchHandler& Self::GetBranchLengthHandler() {
  if (mirror_live_) PushMirror();
  PullMirror();
  mirror_live_ = true;
  return branch_mirror_;
}
const DAGBranchHandler& Self::GetBranchLengthHandler() const {
  if (mirror_live_) PushMirror();
  PullMirror();
  mirror_live_ = true;
  return branch_mirror_;
}
void Self::PullMirror() const {
  const size_t count = static_cast<size_t>(bito_gp_get_gpcsp_count(handle_));
  const size_t padded = static_cast<size_t>(bito_gp_get_padded_gpcsp_count(handle_));
  if (branch_mirror_.GetBranchLengths().GetCount() != count ||
      branch_mirror_.GetBranchLengths().GetSpareCount() != padded - count)
    branch_mirror_.Resize(count, padded - count, std::nullopt, std::nullopt);
  branch_mirror_.SetDefaultBranchLength(default_branch_length_);
  EigenVectorXd& bl = branch_mirror_.GetBranchLengthData();
  EigenVectorXd& diff = branch_mirror_.GetBranchDifferenceData();
  Assert(size_t(bl.size()) >= padded && size_t(diff.size()) >= count,
         "DAGBranchHandler mirror is smaller than the engine.");
  Check(bito_gp_get_branch_lengths(handle_, 0, I(padded), bl.data()));
  Check(bito_gp_get_branch_length_differences(handle_, diff.data()));
  mirror_synced_ = bl.head(padded);
}
void Self::PushMirror() const {
  const EigenVectorXd& bl = branch_mirror_.GetBranchLengthData();
  const size_t padded = static_cast<size_t>(mirror_synced_.size());
  if (size_t(bl.size()) < padded) return;  // the caller shrank it: nothing of ours left to compare
  size_t lo = padded, hi = 0;
  for (size_t i = 0; i < padded; ++i)
    if (bl[i] != mirror_synced_[i]) {
      lo = std::min(lo, i);
      hi = i + 1;
    }
  if (hi == 0) return;
  Check(bito_gp_set_branch_lengths_range(handle_, I(lo), I(hi - lo), bl.data() + lo));
  mirror_synced_.segment(lo, hi - lo) = bl.segment(lo, hi - lo);
}

// ---- read-back: gp_engine.cpp:413-468 ---------------------------------------------------------
EigenVectorXd Self::GetBranchLengths() const { return GetBranchLengths(0, GetGPCSPCount()); }
EigenVectorXd Self::GetBranchLengths(const size_t start, const size_t length) const {
  EigenVectorXd out(length);
  Check(bito_gp_get_branch_lengths(H(), I(start), I(length), out.data()));
  return out;
}
EigenVectorXd Self::GetSpareBranchLengths(const size_t start, const size_t length) const {
  return GetBranchLengths(GetSpareGPCSPIndex(start), length);
}
EigenVectorXd Self::GetBranchLengthDifferences() const {
  EigenVectorXd out(GetGPCSPCount());
  Check(bito_gp_get_branch_length_differences(H(), out.data()));
  return out;
}
EigenVectorXd Self::GetPerGPCSPLogLikelihoods() const {
  return GetPerGPCSPLogLikelihoods(0, GetGPCSPCount());
}
EigenVectorXd Self::GetPerGPCSPLogLikelihoods(const size_t start, const size_t length) const {
  EigenVectorXd out(length);
  Check(bito_gp_get_per_gpcsp_log_likelihoods(H(), I(start), I(length), out.data()));
  return out;
}
EigenVectorXd Self::GetSparePerGPCSPLogLikelihoods(const size_t start, const size_t length) const {
  return GetPerGPCSPLogLikelihoods(GetSpareGPCSPIndex(start), length);
}
EigenVectorXd Self::GetPerGPCSPComponentsOfFullLogMarginal() const {
  EigenVectorXd out(GetGPCSPCount());
  Check(bito_gp_get_per_gpcsp_components_of_full_log_marginal(H(), out.data()));
  return out;
}
const EigenMatrixXd& Self::GetLogLikelihoodMatrix() const {
  EigenMatrixXd& out = log_likelihood_matrix_copy_;
  out.resize(GetGPCSPCount(), GetSitePatternCount());  // row-major (eigen_sugar.hpp:20-22)
  Check(bito_gp_get_log_likelihood_matrix(H(), out.data()));
  return out;
}
const EigenVectorXd& Self::GetHybridMarginals() const {
  EigenVectorXd& out = hybrid_marginals_copy_;
  out.resize(GetGPCSPCount());
  Check(bito_gp_get_hybrid_marginals(H(), out.data()));
  return out;
}
const EigenVectorXd& Self::GetSBNParameters() const {
  EigenVectorXd& out = sbn_parameters_copy_;
  out.resize(GetGPCSPCount());
  Check(bito_gp_get_sbn_parameters(H(), out.data()));
  return out;
}
double Self::GetLogMarginalLikelihood() const {
  double v = 0.;
  Check(bito_gp_get_log_marginal_likelihood(H(), &v));
  return v;
}

// NucleotidePLV is 4 x P column-major (mmapped_plv.hpp:14): pattern p owns 4 consecutive doubles,
// which is the C-ABI's layout, so the Eigen buffer is handed over as is.
NucleotidePLV Self::GetPLV(const PVId plv_index) const {
  NucleotidePLV out(4, GetSitePatternCount());
  Check(bito_gp_get_plv(H(), I(plv_index.value_), out.data()));
  return out;
}
void Self::SetPLV(const PVId plv_index, const NucleotidePLV& plv, int rescaling_count) {
  Assert(size_t(plv.cols()) == GetSitePatternCount(), "SetPLV: wrong pattern count.");
  Check(bito_gp_set_plv(H(), I(plv_index.value_), plv.data(), rescaling_count));
}
PVId Self::GetSparePLVIndex(const PVId plv_index) const {  // pv_handler.hpp:227-232
  Assert(plv_index.value_ < GetSparePLVCount(),
         "Requested temporary pv_id outside of allocated scratch space.");
  return PVId(plv_index.value_ + GetPLVCount());
}
EigenVectorXi Self::GetRescalingCounts() const {
  static_assert(sizeof(int) == sizeof(int32_t), "EigenVectorXi must hold int32");
  EigenVectorXi out(GetPaddedPLVCount());
  Check(bito_gp_get_rescaling_counts(H(), out.data()));
  return out;
}

// ---- quartet hybrid marginals: gp_engine.cpp:748-816 -----------------------------------------
EigenVectorXd Self::CalculateQuartetHybridLikelihoods(const QuartetHybridRequest& request) {
  std::vector<bito_gp_quartet_tip> tips;
  std::vector<int32_t> counts;
  FlattenRequest(request, tips, counts);
  EigenVectorXd out(size_t(counts[0]) * counts[1] * counts[2] * counts[3]);
  Check(bito_gp_calculate_quartet_hybrid_likelihoods(H(), I(request.central_gpcsp_idx_),
                                                     tips.data(), counts.data(), out.data()));
  return out;
}
void Self::ProcessQuartetHybridRequest(const QuartetHybridRequest& request) {
  ProcessQuartetHybridRequests({request});
}
void Self::ProcessQuartetHybridRequests(const std::vector<QuartetHybridRequest>& requests) {
  std::vector<bito_gp_quartet_tip> tips;
  std::vector<int32_t> counts;
  std::vector<int64_t> central;
  for (const auto& request : requests) {
    central.push_back(I(request.central_gpcsp_idx_));
    FlattenRequest(request, tips, counts);
  }
  Check(bito_gp_process_quartet_hybrid_requests(H(), I(central.size()), central.data(),
                                                counts.data(), tips.data()));
}

// ---- branch lengths from a tree sample: gp_engine.cpp:676-746 (host-only, no PLV arithmetic) --
void Self::SetBranchLengthsFromTotals(std::vector<double>& totals, const std::vector<int>& seen,
                                      bool mean) {
  EigenVectorXd bl(totals.size());
  for (size_t e = 0; e < totals.size(); ++e)
    bl[e] = seen[e] == 0 ? default_branch_length_
                         : (mean ? totals[e] / static_cast<double>(seen[e]) : totals[e]);
  SetBranchLengths(std::move(bl));
}
void Self::HotStartBranchLengths(const RootedTreeCollection& tree_collection,
                                 const BitsetSizeMap& indexer) {
  const size_t n = GetGPCSPCount();
  std::vector<double> totals(n, 0.);
  std::vector<int> seen(n, 0);
  RootedSBNMaps::FunctionOverRootedTreeCollection(
      [&](EdgeId e, const Bitset&, const RootedTree& tree, const size_t, const Node* focal) {
        totals[e.value_] += tree.BranchLength(focal);
        seen[e.value_]++;
      },
      tree_collection, indexer, n);
  SetBranchLengthsFromTotals(totals, seen, true);
}
void Self::TakeFirstBranchLength(const RootedTreeCollection& tree_collection,
                                 const BitsetSizeMap& indexer) {
  const size_t n = GetGPCSPCount();
  std::vector<double> first(n, 0.);
  std::vector<int> seen(n, 0);
  RootedSBNMaps::FunctionOverRootedTreeCollection(
      [&](EdgeId e, const Bitset&, const RootedTree& tree, const size_t, const Node* focal) {
        if (seen[e.value_] == 0) {
          first[e.value_] = tree.BranchLength(focal);
          seen[e.value_] = 1;
        }
      },
      tree_collection, indexer, n);
  SetBranchLengthsFromTotals(first, seen, false);
}
SizeDoubleVectorMap Self::GatherBranchLengths(const RootedTreeCollection& tree_collection,
                                              const BitsetSizeMap& indexer) {
  SizeDoubleVectorMap by_gpcsp;
  RootedSBNMaps::FunctionOverRootedTreeCollection(
      [&](EdgeId e, const Bitset&, const RootedTree& tree, const size_t, const Node* focal) {
        by_gpcsp[e.value_].push_back(tree.BranchLength(focal));
      },
      tree_collection, indexer, GetGPCSPCount());
  return by_gpcsp;
}

// ---- derivatives: gp_engine.cpp:470-542 -------------------------------------------------------
DoublePair Self::LogLikelihoodAndDerivative(const GPOperations::OptimizeBranchLength& op) {
  return LogLikelihoodAndDerivative(op.gpcsp_, op.rootward_, op.leafward_);
}
DoublePair Self::LogLikelihoodAndDerivative(const size_t gpcsp, const size_t rootward,
                                            const size_t leafward) {
  double out[3];
  Check(bito_gp_log_likelihood_and_derivatives(H(), I(gpcsp), I(rootward), I(leafward), out));
  return {out[0], out[1]};
}
std::tuple<double, double, double> Self::LogLikelihoodAndFirstTwoDerivatives(
    const GPOperations::OptimizeBranchLength& op) {
  return LogLikelihoodAndFirstTwoDerivatives(op.gpcsp_, op.rootward_, op.leafward_);
}
std::tuple<double, double, double> Self::LogLikelihoodAndFirstTwoDerivatives(
    const size_t gpcsp, const size_t rootward, const size_t leafward) {
  double out[3];
  Check(bito_gp_log_likelihood_and_derivatives(H(), I(gpcsp), I(rootward), I(leafward), out));
  return {out[0], out[1], out[2]};
}

// ---- I/O --------------------------------------------------------------------------------------
std::string Self::PLVToString(const PVId plv_idx) const {
  // Debug printout, one PLV row (state) per line as pv_handler.hpp:381-386 prints spare PVs.
  const NucleotidePLV plv = GetPLV(plv_idx);
  std::stringstream out;
  out << "PV[" << plv_idx.value_ << "]: " << std::endl;
  for (Eigen::Index r = 0; r < plv.rows(); ++r) {
    for (Eigen::Index c = 0; c < plv.cols(); ++c) out << (c ? " " : "") << plv(r, c);
    out << std::endl;
  }
  return out.str();
}
std::string Self::LogLikelihoodMatrixToString() const {
  const EigenMatrixXd m = GetLogLikelihoodMatrix();
  std::stringstream out;
  for (Eigen::Index i = 0; i < m.rows(); i++) {
    for (Eigen::Index j = 0; j < m.cols(); j++) out << "[" << i << "," << j << "]: " << m(i, j) << "\t";
    out << std::endl;
  }
  return out.str();
}

// ---- counts -----------------------------------------------------------------------------------
size_t Self::GetNodeCount() const { return size_t(bito_gp_get_node_count(H())); }
size_t Self::GetSpareNodeCount() const { return GetSparePLVCount() / 6; }
size_t Self::GetPLVCount() const { return size_t(bito_gp_get_plv_count(H())); }
size_t Self::GetPaddedPLVCount() const { return size_t(bito_gp_get_padded_plv_count(H())); }
size_t Self::GetGPCSPCount() const { return size_t(bito_gp_get_gpcsp_count(H())); }
size_t Self::GetPaddedGPCSPCount() const { return size_t(bito_gp_get_padded_gpcsp_count(H())); }
size_t Self::GetSpareGPCSPIndex(const size_t gpcsp_offset) const {
  Assert(gpcsp_offset < GetSpareGPCSPCount(),
         "Requested gpcsp_offset outside of allocated scratch space.");
  return gpcsp_offset + GetGPCSPCount();
}
bito_gp_stats Self::Stats() const {
  bito_gp_stats s{};
  Check(bito_gp_get_stats(H(), &s));
  return s;
}
