// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
//
// A C-ABI driver around the UNMODIFIED reference GP path (GPDAG planner + CPU GPEngine),
// compiled against the reference headers where they lie under /root/reference/src (see
// oracle/Makefile). It clones what GPInstance::MakeGPEngine does
// (/root/reference/src/gp_instance.cpp:146-164) without touching BEAGLE, flattens
// GPOperationVectors (/root/reference/src/gp_operation.hpp:163-170) into the int64[n][6]
// table that include/bito_gp.h defines, and exposes the engine's private numeric state
// (rescaling_counts_, log_likelihoods_, q_; /root/reference/src/gp_engine.hpp:317-349)
// so parity tests can compare it bit for bit.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the resulting oracle/_ref/libbito_gp_ref.so. The product never does.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <variant>
#include <vector>

// Access to private numeric state of the reference classes (test-only). Access
// specifiers do not change object layout or symbol names under the Itanium ABI, so the
// library objects (compiled without this) and this TU agree.
#define private public
#define protected public
#include "gp_dag.hpp"
#include "gp_engine.hpp"
#include "driver.hpp"
#include "rooted_tree_collection.hpp"
#include "site_pattern.hpp"
#undef private
#undef protected

namespace {

thread_local std::string g_error;

struct RefInst {
  std::unique_ptr<GPDAG> dag;
  std::unique_ptr<GPEngine> engine;
  RootedTreeCollection trees;
  size_t taxon_count = 0;
  size_t site_count = 0;
};

enum OpKind : int64_t {
  kZeroPLV = 0,
  kSetToStationaryDistribution = 1,
  kIncrementWithWeightedEvolvedPLV = 2,
  kMultiply = 3,
  kLikelihood = 4,
  kOptimizeBranchLength = 5,
  kUpdateSBNProbabilities = 6,
  kResetMarginalLikelihood = 7,
  kIncrementMarginalLikelihood = 8,
  kPrepForMarginalization = 9
};

// Field order follows the struct member order in gp_operation.hpp.
struct Flattener {
  std::vector<int64_t>& ops;
  std::vector<int64_t>& vec;
  void Row(int64_t kind, int64_t a = 0, int64_t b = 0, int64_t c = 0, int64_t off = 0,
           int64_t len = 0) {
    ops.insert(ops.end(), {kind, a, b, c, off, len});
  }
  void operator()(const GPOperations::ZeroPLV& op) { Row(kZeroPLV, op.dest_); }
  void operator()(const GPOperations::SetToStationaryDistribution& op) {
    Row(kSetToStationaryDistribution, op.dest_, op.root_gpcsp_idx_);
  }
  void operator()(const GPOperations::IncrementWithWeightedEvolvedPLV& op) {
    Row(kIncrementWithWeightedEvolvedPLV, op.dest_, op.gpcsp_, op.src_);
  }
  void operator()(const GPOperations::Multiply& op) {
    Row(kMultiply, op.dest_, op.src1_, op.src2_);
  }
  void operator()(const GPOperations::Likelihood& op) {
    Row(kLikelihood, op.dest_, op.child_, op.parent_);
  }
  void operator()(const GPOperations::OptimizeBranchLength& op) {
    Row(kOptimizeBranchLength, op.leafward_, op.rootward_, op.gpcsp_);
  }
  void operator()(const GPOperations::UpdateSBNProbabilities& op) {
    Row(kUpdateSBNProbabilities, op.start_, op.stop_);
  }
  void operator()(const GPOperations::ResetMarginalLikelihood&) {
    Row(kResetMarginalLikelihood);
  }
  void operator()(const GPOperations::IncrementMarginalLikelihood& op) {
    Row(kIncrementMarginalLikelihood, op.stationary_times_prior_, op.rootsplit_, op.p_);
  }
  void operator()(const GPOperations::PrepForMarginalization& op) {
    const int64_t off = static_cast<int64_t>(vec.size());
    for (auto s : op.src_vector_) vec.push_back(static_cast<int64_t>(s));
    Row(kPrepForMarginalization, op.dest_, 0, 0, off,
        static_cast<int64_t>(op.src_vector_.size()));
  }
};

GPOperationVector Unflatten(const int64_t* ops, int64_t n, const int64_t* vec) {
  GPOperationVector out;
  out.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t* r = ops + 6 * i;
    const size_t a = r[1], b = r[2], c = r[3];
    switch (r[0]) {
      case kZeroPLV:
        out.push_back(GPOperations::ZeroPLV{a});
        break;
      case kSetToStationaryDistribution:
        out.push_back(GPOperations::SetToStationaryDistribution{a, b});
        break;
      case kIncrementWithWeightedEvolvedPLV:
        out.push_back(GPOperations::IncrementWithWeightedEvolvedPLV{a, b, c});
        break;
      case kMultiply:
        out.push_back(GPOperations::Multiply{a, b, c});
        break;
      case kLikelihood:
        out.push_back(GPOperations::Likelihood{a, b, c});
        break;
      case kOptimizeBranchLength:
        out.push_back(GPOperations::OptimizeBranchLength{a, b, c});
        break;
      case kUpdateSBNProbabilities:
        out.push_back(GPOperations::UpdateSBNProbabilities{a, b});
        break;
      case kResetMarginalLikelihood:
        out.push_back(GPOperations::ResetMarginalLikelihood{});
        break;
      case kIncrementMarginalLikelihood:
        out.push_back(GPOperations::IncrementMarginalLikelihood{a, b, c});
        break;
      case kPrepForMarginalization: {
        SizeVector src(vec + r[4], vec + r[4] + r[5]);
        out.push_back(GPOperations::PrepForMarginalization{a, std::move(src)});
        break;
      }
      default:
        throw std::runtime_error("ref_driver: unknown op kind");
    }
  }
  return out;
}

template <typename F>
int Guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }

// Clone of GPInstanceOfFiles (/root/reference/src/gp_doctest.cpp:45-56) +
// GPInstance::MakeGPEngine (/root/reference/src/gp_instance.cpp:146-164).
void* ref_open(const char* fasta_path, const char* newick_path, const char* mmap_path,
               double rescaling_threshold, int use_gradients) {
  RefInst* inst = nullptr;
  const int rc = Guard([&] {
    auto up = std::make_unique<RefInst>();
    Alignment alignment = Alignment::ReadFasta(fasta_path);
    Driver driver;
    driver.SetSortTaxa(false);
    up->trees = RootedTreeCollection::OfTreeCollection(driver.ParseNewickFile(newick_path));
    SitePattern site_pattern(alignment, up->trees.TagTaxonMap());
    up->taxon_count = site_pattern.TaxonCount();
    up->site_count = site_pattern.SiteCount();
    up->dag = std::make_unique<GPDAG>(up->trees);
    auto& dag = *up->dag;
    auto sbn_prior = dag.BuildUniformOnTopologicalSupportPrior();
    auto unconditional = dag.UnconditionalNodeProbabilities(sbn_prior);
    auto inverted = dag.InvertedGPCSPProbabilities(sbn_prior, unconditional);
    up->engine = std::make_unique<GPEngine>(
        std::move(site_pattern), dag.NodeCountWithoutDAGRoot(),
        dag.EdgeCountWithLeafSubsplits(), std::string(mmap_path), rescaling_threshold,
        std::move(sbn_prior), unconditional.segment(0, dag.NodeCountWithoutDAGRoot()),
        std::move(inverted), use_gradients != 0);
    inst = up.release();
  });
  return rc == 0 ? inst : nullptr;
}

// Engine without a DAG: for synthetic workloads whose op lists come from the caller.
// symbols is taxa x patterns (row-major, values 0..4), as SitePattern::GetPatterns().
void* ref_open_raw(int64_t taxon_count, int64_t pattern_count, const uint8_t* symbols,
                   const double* weights, int64_t site_count, int64_t node_count,
                   int64_t edge_count, const char* mmap_path, double rescaling_threshold,
                   const double* sbn_prior, const double* unconditional,
                   const double* inverted, int use_gradients) {
  RefInst* inst = nullptr;
  const int rc = Guard([&] {
    auto up = std::make_unique<RefInst>();
    SitePattern site_pattern;
    site_pattern.patterns_.resize(taxon_count);
    for (int64_t t = 0; t < taxon_count; ++t) {
      auto& row = site_pattern.patterns_[t];
      row.resize(pattern_count);
      for (int64_t p = 0; p < pattern_count; ++p) row[p] = symbols[t * pattern_count + p];
      site_pattern.tag_taxon_map_[PackInts(static_cast<uint32_t>(t), 1)] =
          "t" + std::to_string(t);
    }
    site_pattern.weights_.assign(weights, weights + pattern_count);
    site_pattern.alignment_ =
        Alignment({{"t0", std::string(static_cast<size_t>(site_count), 'A')}});
    up->taxon_count = taxon_count;
    up->site_count = site_count;
    EigenVectorXd q = Eigen::Map<const EigenVectorXd>(sbn_prior, edge_count);
    EigenVectorXd un = Eigen::Map<const EigenVectorXd>(unconditional, node_count);
    EigenVectorXd inv = Eigen::Map<const EigenVectorXd>(inverted, edge_count);
    up->engine = std::make_unique<GPEngine>(
        std::move(site_pattern), static_cast<size_t>(node_count),
        static_cast<size_t>(edge_count), std::string(mmap_path), rescaling_threshold,
        std::move(q), std::move(un), std::move(inv), use_gradients != 0);
    inst = up.release();
  });
  return rc == 0 ? inst : nullptr;
}

void ref_close(void* h) { delete static_cast<RefInst*>(h); }

// out: taxa, patterns, sites, nodes (without DAG root), edges, rootsplits, plv_count,
//      padded_plv_count, topology_count (as double bits rounded), has_dag
void ref_sizes(void* h, int64_t* out) {
  auto* inst = static_cast<RefInst*>(h);
  auto& e = *inst->engine;
  out[0] = inst->taxon_count;
  out[1] = e.GetSitePatternCount();
  out[2] = inst->site_count;
  out[3] = e.GetNodeCount();
  out[4] = e.GetGPCSPCount();
  out[5] = inst->dag ? inst->dag->RootsplitCount() : 0;
  out[6] = e.GetPLVCount();
  out[7] = e.GetPaddedPLVCount();
  out[8] = inst->dag ? static_cast<int64_t>(inst->dag->TopologyCount()) : 0;
  out[9] = inst->dag ? 1 : 0;
}

void ref_patterns(void* h, uint8_t* symbols, double* weights) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  const auto& pats = e.site_pattern_.GetPatterns();
  const size_t P = e.GetSitePatternCount();
  for (size_t t = 0; t < pats.size(); ++t)
    for (size_t p = 0; p < P; ++p) symbols[t * P + p] = static_cast<uint8_t>(pats[t][p]);
  const auto& w = e.site_pattern_.GetWeights();
  std::copy(w.begin(), w.end(), weights);
}

// q (E), unconditional node probabilities (N), inverted prior (E).
void ref_priors(void* h, double* q, double* unconditional, double* inverted) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t i = 0; i < e.GetGPCSPCount(); ++i) {
    q[i] = e.q_[i];
    inverted[i] = e.inverted_sbn_prior_[i];
  }
  for (size_t i = 0; i < e.GetNodeCount(); ++i)
    unconditional[i] = e.unconditional_node_probabilities_[i];
}

// DAG structure, per edge id: parent node, child node, 1 if the edge descends from the
// parent's left (rotated) clade.
int ref_edges(void* h, int64_t* parent, int64_t* child, int32_t* on_left) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    auto& dag = *inst->dag;
    for (size_t i = 0; i < dag.EdgeCountWithLeafSubsplits(); ++i) {
      auto line = dag.GetDAGEdge(EdgeId(i));
      parent[i] = line.GetParent().value_;
      child[i] = line.GetChild().value_;
      on_left[i] = line.GetSubsplitClade() == SubsplitClade::Left ? 1 : 0;
    }
  });
}

// Subsplit bitset of every node (incl. DAG root, last) as '0'/'1' chars, 2*taxa per node.
int ref_node_bitsets(void* h, char* out) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    auto& dag = *inst->dag;
    const size_t width = 2 * dag.TaxonCount();
    for (size_t i = 0; i < dag.NodeCount(); ++i) {
      const std::string s = dag.GetDAGNodeBitset(NodeId(i)).ToString();
      std::memcpy(out + i * width, s.data(), width);
    }
  });
}

int64_t ref_node_count_with_root(void* h) {
  return static_cast<RefInst*>(h)->dag->NodeCount();
}

// Taxon names in taxon-id order, '\n'-separated.
int64_t ref_taxon_names(void* h, char* out, int64_t cap) {
  auto* inst = static_cast<RefInst*>(h);
  std::string all;
  int64_t rc = Guard([&] {
    auto& dag = *inst->dag;
    std::vector<std::string> names(dag.TaxonCount());
    for (const auto& name : dag.BuildSetOfTaxonNames())
      names[dag.GetTaxonId(name).value_] = name;
    for (auto& n : names) all += n + "\n";
  });
  if (rc != 0) return -1;
  if (static_cast<int64_t>(all.size()) + 1 > cap) return -static_cast<int64_t>(all.size());
  std::memcpy(out, all.c_str(), all.size() + 1);
  return all.size();
}

// which: 0 PopulatePLVs, 1 ComputeLikelihoods, 2 MarginalLikelihood,
//        3 BranchLengthOptimization, 4 OptimizeSBNParameters, 5 RootwardPass,
//        6 LeafwardPass, 7 SetRootwardZero, 8 SetLeafwardZero, 9 SetRhatToStationary,
//        10 ApproximateBranchLengthOptimization.
// Returns the op count (call with ops == nullptr to size); -1 on error.
int64_t ref_oplist(void* h, int which, int64_t* ops, int64_t ops_cap, int64_t* vec,
                   int64_t vec_cap, int64_t* vec_len) {
  auto* inst = static_cast<RefInst*>(h);
  std::vector<int64_t> o, v;
  const int rc = Guard([&] {
    auto& dag = *inst->dag;
    GPOperationVector list;
    switch (which) {
      case 0: list = dag.PopulatePLVs(); break;
      case 1: list = dag.ComputeLikelihoods(); break;
      case 2: list = dag.MarginalLikelihood(); break;
      case 3: list = dag.BranchLengthOptimization(); break;
      case 4: list = dag.OptimizeSBNParameters(); break;
      case 5: list = dag.RootwardPass(); break;
      case 6: list = dag.LeafwardPass(); break;
      case 7: list = dag.SetRootwardZero(); break;
      case 8: list = dag.SetLeafwardZero(); break;
      case 9: list = dag.SetRhatToStationary(); break;
      case 10: list = dag.ApproximateBranchLengthOptimization(); break;
      default: throw std::runtime_error("ref_oplist: unknown list id");
    }
    Flattener fl{o, v};
    for (const auto& op : list) std::visit(fl, op);
  });
  if (rc != 0) return -1;
  const int64_t n = static_cast<int64_t>(o.size() / 6);
  *vec_len = static_cast<int64_t>(v.size());
  if (ops != nullptr) {
    if (n > ops_cap || *vec_len > vec_cap) {
      g_error = "ref_oplist: buffer too small";
      return -1;
    }
    std::copy(o.begin(), o.end(), ops);
    std::copy(v.begin(), v.end(), vec);
  }
  return n;
}

int ref_run(void* h, const int64_t* ops, int64_t n, const int64_t* vec) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] { inst->engine->ProcessOperations(Unflatten(ops, n, vec)); });
}

// Times `repeats` executions of the list with steady_clock; writes each duration (s).
int ref_time_run(void* h, const int64_t* ops, int64_t n, const int64_t* vec, int repeats,
                 double* seconds) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    const GPOperationVector list = Unflatten(ops, n, vec);
    for (int r = 0; r < repeats; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      inst->engine->ProcessOperations(list);
      const auto t1 = std::chrono::steady_clock::now();
      seconds[r] = std::chrono::duration<double>(t1 - t0).count();
    }
  });
}

// ---- state getters / setters -------------------------------------------------------

int ref_get_plv(void* h, int64_t plv_id, double* out /* 4*P, pattern-major */) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    const auto& plv = inst->engine->GetPLV(PVId(plv_id));
    for (Eigen::Index p = 0; p < plv.cols(); ++p)
      for (Eigen::Index r = 0; r < 4; ++r) out[4 * p + r] = plv(r, p);
  });
}

int ref_set_plv(void* h, int64_t plv_id, const double* in, int32_t count) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    auto& plv = inst->engine->GetPLV(PVId(plv_id));
    for (Eigen::Index p = 0; p < plv.cols(); ++p)
      for (Eigen::Index r = 0; r < 4; ++r) plv(r, p) = in[4 * p + r];
    inst->engine->rescaling_counts_(plv_id) = count;
  });
}

void ref_get_counts(void* h, int32_t* out /* padded plv count */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t i = 0; i < e.GetPaddedPLVCount(); ++i) out[i] = e.rescaling_counts_(i);
}

void ref_get_loglik_matrix(void* h, double* out /* E*P row-major */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  const size_t E = e.GetGPCSPCount(), P = e.GetSitePatternCount();
  for (size_t i = 0; i < E; ++i)
    for (size_t p = 0; p < P; ++p) out[i * P + p] = e.log_likelihoods_(i, p);
}

void ref_get_per_pattern_marginal(void* h, double* out /* P */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t p = 0; p < e.GetSitePatternCount(); ++p) out[p] = e.log_marginal_likelihood_[p];
}

void ref_get_per_gpcsp_loglik(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  EigenVectorXd v = e.GetPerGPCSPLogLikelihoods();
  std::copy(v.data(), v.data() + v.size(), out);
}

void ref_get_per_gpcsp_components(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  EigenVectorXd v = e.GetPerGPCSPComponentsOfFullLogMarginal();
  std::copy(v.data(), v.data() + e.GetGPCSPCount(), out);
}

double ref_get_log_marginal(void* h) {
  return static_cast<RefInst*>(h)->engine->GetLogMarginalLikelihood();
}

void ref_get_q(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t i = 0; i < e.GetGPCSPCount(); ++i) out[i] = e.q_[i];
}

void ref_set_q(void* h, const double* in /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t i = 0; i < e.GetGPCSPCount(); ++i) e.q_[i] = in[i];
}

void ref_get_branch_lengths(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  EigenVectorXd v = e.GetBranchLengths();
  std::copy(v.data(), v.data() + v.size(), out);
}

void ref_set_branch_lengths(void* h, const double* in /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  EigenVectorXd v = Eigen::Map<const EigenVectorXd>(in, e.GetGPCSPCount());
  e.SetBranchLengths(v);
}

void ref_set_branch_lengths_constant(void* h, double t) {
  static_cast<RefInst*>(h)->engine->SetBranchLengthsToConstant(t);
}

void ref_get_branch_differences(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  EigenVectorXd v = e.GetBranchLengthDifferences();
  std::copy(v.data(), v.data() + v.size(), out);
}

// method: Optimization::OptimizationMethod order (optimization.hpp:28-34).
void ref_set_optimization_method(void* h, int method) {
  static_cast<RefInst*>(h)->engine->SetOptimizationMethod(
      static_cast<OptimizationMethod>(method));
}
void ref_use_gradient_optimization(void* h, int use) {
  static_cast<RefInst*>(h)->engine->UseGradientOptimization(use != 0);
}
void ref_set_significant_digits(void* h, int digits) {
  static_cast<RefInst*>(h)->engine->SetSignificantDigitsForOptimization(digits);
}
void ref_reset_optimization_count(void* h) {
  static_cast<RefInst*>(h)->engine->ResetOptimizationCount();
}
void ref_increment_optimization_count(void* h) {
  static_cast<RefInst*>(h)->engine->IncrementOptimizationCount();
}
int64_t ref_get_optimization_count(void* h) {
  return static_cast<RefInst*>(h)->engine->GetOptimizationCount();
}

void ref_set_null_prior(void* h) { static_cast<RefInst*>(h)->engine->SetNullPrior(); }

// out: ll, d1 (and d2 when two_derivatives != 0); (gp_engine.cpp:470-542)
int ref_loglik_and_derivatives(void* h, int64_t gpcsp, int64_t rootward, int64_t leafward,
                               int two_derivatives, double* out) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    if (two_derivatives) {
      auto [a, b, c] =
          inst->engine->LogLikelihoodAndFirstTwoDerivatives(gpcsp, rootward, leafward);
      out[0] = a; out[1] = b; out[2] = c;
    } else {
      auto [a, b] = inst->engine->LogLikelihoodAndDerivative(gpcsp, rootward, leafward);
      out[0] = a; out[1] = b;
    }
  });
}

// JC69 transition matrix at branch length t (gp_engine.cpp:341-344), row-major 4x4.
// The eigensystem the engine was built with (JC69 in the stock build; whatever BITO_REF_MODEL selected in the
// `refmodel` build): V and V^-1 row-major, eigenvalues, stationary frequencies (gp_engine.hpp:366-376).
void ref_model_eigensystem(void* h, double* eigenvectors, double* inverse_eigenvectors, double* eigenvalues,
                           double* frequencies) {
  const GPEngine& e = *static_cast<RefInst*>(h)->engine;
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) {
      eigenvectors[4 * i + j] = e.eigenmatrix_(i, j);
      inverse_eigenvectors[4 * i + j] = e.inverse_eigenmatrix_(i, j);
    }
    eigenvalues[i] = e.eigenvalues_[i];
    frequencies[i] = e.stationary_distribution_[i];
  }
}

void ref_transition_matrix(void* h, double t, double* out) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  e.SetTransitionMatrixToHaveBranchLength(t);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = e.GetTransitionMatrix()(i, j);
}

// ---- quartet hybrid marginals (gp_engine.cpp:748-816, gp_dag.cpp:413-458) --------------------

namespace {
QuartetHybridRequest RequestOf(int64_t central, const int32_t* counts, const int64_t* tips) {
  QuartetTipVector v[4];
  const int64_t* t = tips;
  for (int k = 0; k < 4; ++k)
    for (int32_t i = 0; i < counts[k]; ++i, t += 3)
      v[k].emplace_back(static_cast<size_t>(t[0]), static_cast<size_t>(t[1]),
                        static_cast<size_t>(t[2]));
  return QuartetHybridRequest(static_cast<size_t>(central), std::move(v[0]), std::move(v[1]),
                              std::move(v[2]), std::move(v[3]));
}
}  // namespace

// Every request GPInstance::CalculateHybridMarginals issues (gp_instance.cpp:408-417), in its
// TopologicalEdgeTraversal order, flattened: central[r], tip_counts[4r..] (rootward, sister,
// rotated, sorted), tips = (tip_node_id, plv_idx, gpcsp_idx) triples. Call with central == nullptr
// to size. Returns the request count, -1 on error.
int64_t ref_quartet_requests(void* h, int64_t* central, int32_t* tip_counts, int64_t* tips,
                             int64_t cap_requests, int64_t cap_tips, int64_t* n_tips) {
  auto* inst = static_cast<RefInst*>(h);
  std::vector<int64_t> c, t;
  std::vector<int32_t> n;
  const int rc = Guard([&] {
    auto& dag = *inst->dag;
    dag.TopologicalEdgeTraversal([&](const NodeId parent_id, const bool is_edge_on_left,
                                     const NodeId child_id, const EdgeId edge_idx) {
      const QuartetHybridRequest req =
          dag.QuartetHybridRequestOf(parent_id, is_edge_on_left, child_id);
      c.push_back(static_cast<int64_t>(req.central_gpcsp_idx_));
      for (const QuartetTipVector* v : {&req.rootward_tips_, &req.sister_tips_,
                                        &req.rotated_tips_, &req.sorted_tips_}) {
        n.push_back(static_cast<int32_t>(v->size()));
        for (const auto& tip : *v) {
          t.push_back(static_cast<int64_t>(tip.tip_node_id_));
          t.push_back(static_cast<int64_t>(tip.plv_idx_));
          t.push_back(static_cast<int64_t>(tip.gpcsp_idx_));
        }
      }
    });
  });
  if (rc != 0) return -1;
  *n_tips = static_cast<int64_t>(t.size() / 3);
  if (central != nullptr) {
    if (static_cast<int64_t>(c.size()) > cap_requests || *n_tips > cap_tips) {
      g_error = "ref_quartet_requests: buffer too small";
      return -1;
    }
    std::copy(c.begin(), c.end(), central);
    std::copy(n.begin(), n.end(), tip_counts);
    std::copy(t.begin(), t.end(), tips);
  }
  return static_cast<int64_t>(c.size());
}

// GPEngine::CalculateQuartetHybridLikelihoods; out has prod(counts) entries.
int ref_quartet_likelihoods(void* h, int64_t central, const int32_t* counts, const int64_t* tips,
                            double* out) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    EigenVectorXd v = inst->engine->CalculateQuartetHybridLikelihoods(RequestOf(central, counts, tips));
    std::copy(v.data(), v.data() + v.size(), out);
  });
}

// GPEngine::ProcessQuartetHybridRequest for each flattened request.
int ref_process_quartet_requests(void* h, int64_t n, const int64_t* central, const int32_t* counts,
                                 const int64_t* tips) {
  auto* inst = static_cast<RefInst*>(h);
  return Guard([&] {
    const int64_t* t = tips;
    for (int64_t r = 0; r < n; ++r) {
      const int32_t* c = counts + 4 * r;
      inst->engine->ProcessQuartetHybridRequest(RequestOf(central[r], c, t));
      t += 3 * (static_cast<int64_t>(c[0]) + c[1] + c[2] + c[3]);
    }
  });
}

void ref_get_hybrid_marginals(void* h, double* out /* E */) {
  auto& e = *static_cast<RefInst*>(h)->engine;
  for (size_t i = 0; i < e.GetGPCSPCount(); ++i) out[i] = e.hybrid_marginal_log_likelihoods_[i];
}

}  // extern "C"
