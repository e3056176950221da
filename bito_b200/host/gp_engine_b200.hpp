// bito_b200/host/gp_engine_b200.hpp — the reference-side host class over the C-ABI.
//
// A C++ class with the public surface of the reference's `GPEngine`
// (/root/reference/src/gp_engine.hpp:24-236), written in the reference's own language and
// against the reference's own value types (SitePattern, GPOperationVector, EigenVectorXd,
// Reindexer, QuartetHybridRequest, ...), whose every member forwards to
// libbito_gp_b200.so through include/bito_gp.h. It is what a bito maintainer compiles in place
// of src/gp_engine.cpp: build bito with -DBITO_B200_ENGINE_CLASS=GPEngine and this header
// installed as gp_engine.hpp. The default class name GPEngineB200 lets the
// parity program tests/cpp/host_parity.cpp hold the reference GPEngine and this class side by side.
//
// Differences from the reference surface, all forced by PLVs living in HBM:
//  * GetPLV / GetSparePLV return copies, not Eigen::Ref into engine memory; writes go through SetPLV.
//    GetLogLikelihoodMatrix / GetHybridMarginals / GetSBNParameters return references to host copies
//    this object holds (valid until the next call of the same getter).
//  * GetPLVHandler() returns an index-only view (GetPVIndex / GetSparePVIndex / counts): PLV data
//    is not host-addressable. That is all NNIEvalEngineViaGP takes from it
//    (nni_evaluation_engine.cpp:233-421, 633, 813-923).
//  * GetBranchLengthHandler() returns a genuine reference DAGBranchHandler that MIRRORS the
//    device's branch lengths and differences: it is refreshed after every member that can change
//    them, and whatever the caller wrote through it (branch_handler(edge) = x,
//    nni_evaluation_engine.cpp:108, 187) is uploaded before the next member touches the device.
//    Optimiser settings are NOT read from the mirror: set them through the engine's own members.
//  * mmap_file_path is accepted and ignored.
#pragma once

#include <memory>
#include <optional>
#include <string>
#include <tuple>
#include <vector>

#include "eigen_sugar.hpp"
#include "gp_operation.hpp"
#include "mmapped_plv.hpp"
#include "optimization.hpp"
#include "quartet_hybrid_request.hpp"
#include "reindexer.hpp"
#include "rooted_tree_collection.hpp"
#include "sbn_maps.hpp"
#include "site_pattern.hpp"
#include "substitution_model.hpp"
#include "subsplit_dag_storage.hpp"
#include "pv_handler.hpp"
#include "dag_branch_handler.hpp"

#include "bito_gp.h"

#ifndef BITO_B200_ENGINE_CLASS
#define BITO_B200_ENGINE_CLASS GPEngineB200
#endif

class BITO_B200_ENGINE_CLASS {
 public:
  // gp_engine.hpp:26-29. `device` / `flags` (BITO_GP_FLAG_*) have defaults so the reference's call
  // sites (gp_instance.cpp:155-163) compile unchanged.
  BITO_B200_ENGINE_CLASS(SitePattern site_pattern, size_t node_count, size_t gpcsp_count,
                         const std::string& mmap_file_path, double rescaling_threshold,
                         EigenVectorXd sbn_prior, EigenVectorXd unconditional_node_probabilities,
                         EigenVectorXd inverted_sbn_prior, bool use_gradients, int device = 0,
                         int flags = 0);
  ~BITO_B200_ENGINE_CLASS();
  BITO_B200_ENGINE_CLASS(const BITO_B200_ENGINE_CLASS&) = delete;
  BITO_B200_ENGINE_CLASS& operator=(const BITO_B200_ENGINE_CLASS&) = delete;

  void InitializePriors(EigenVectorXd sbn_prior, EigenVectorXd unconditional_node_probabilities,
                        EigenVectorXd inverted_sbn_prior);
  void SetNullPrior();

  // ** Resizing and Reindexing (gp_engine.hpp:39-52)
  void GrowPLVs(const size_t node_count,
                std::optional<const Reindexer> node_reindexer = std::nullopt,
                std::optional<const size_t> explicit_allocation = std::nullopt,
                const bool on_initialization = false);
  void GrowGPCSPs(const size_t gpcsp_count,
                  std::optional<const Reindexer> gpcsp_reindexer = std::nullopt,
                  std::optional<const size_t> explicit_allocation = std::nullopt,
                  const bool on_intialization = false);
  void GrowSparePLVs(const size_t new_node_spare_count);
  void GrowSpareGPCSPs(const size_t new_gpcsp_spare_count);

  // ** GPOperations (gp_engine.hpp:55-67): a single op is a one-element list.
  template <typename Op>
  void operator()(const Op& op) {
    ProcessOperations(GPOperationVector{GPOperation(op)});
  }
  void ProcessOperations(GPOperationVector operations);

  // ** Branch Length Optimization (gp_engine.hpp:71-92)
  void OptimizeBranchLength(const GPOperations::OptimizeBranchLength& op) { (*this)(op); }
  void SetOptimizationMethod(const OptimizationMethod method);
  void UseGradientOptimization(const bool use_gradients);
  void SetSignificantDigitsForOptimization(int significant_digits);
  size_t GetOptimizationCount();
  void ResetOptimizationCount();
  void IncrementOptimizationCount();
  bool IsFirstOptimization() { return GetOptimizationCount() == 0; }

  // Extension: the reference engine hard-wires `JC69Model substitution_model_`
  // (gp_engine.hpp:366) although it only reads the four generic getters; any SubstitutionModel
  // (GTRModel, HKYModel: substitution_model.hpp:80-111) can be installed here.
  void SetSubstitutionModel(const SubstitutionModel& model);
  void SetTransitionMatrixToHaveBranchLength(double branch_length);
  const Eigen::Matrix4d& GetTransitionMatrix() const { return transition_matrix_; }
  void SetBranchLengths(EigenVectorXd branch_lengths);
  void SetBranchLengthsToConstant(double branch_length);
  void SetBranchLengthsToDefault();
  void ResetLogMarginalLikelihood();

  void CopyNodeData(const NodeId src_node_idx, const NodeId dest_node_idx);
  void CopyPLVData(const size_t src_plv_idx, const size_t dest_plv_idx);
  void CopyGPCSPData(const EdgeId src_gpcsp_idx, const EdgeId dest_gpcsp_idx);

  // ** Access (gp_engine.hpp:99-156)
  EigenVectorXd GetBranchLengths() const;
  EigenVectorXd GetBranchLengths(const size_t start, const size_t length) const;
  EigenVectorXd GetSpareBranchLengths(const size_t start, const size_t length) const;
  EigenVectorXd GetBranchLengthDifferences() const;
  EigenVectorXd GetPerGPCSPLogLikelihoods() const;
  EigenVectorXd GetPerGPCSPLogLikelihoods(const size_t start, const size_t length = 1) const;
  EigenVectorXd GetSparePerGPCSPLogLikelihoods(const size_t start, const size_t length = 1) const;
  EigenVectorXd GetPerGPCSPComponentsOfFullLogMarginal() const;
  // The reference returns Eigen::Ref views into engine memory here (gp_engine.hpp:136-138) and callers bind
  // them (GPInstance::GetSBNParameters returns such a Ref, gp_instance.cpp:419-421): these three return
  // references to host copies held by this object, refreshed by each call, so a bound Ref stays valid.
  const EigenMatrixXd& GetLogLikelihoodMatrix() const;
  const EigenVectorXd& GetHybridMarginals() const;
  const EigenVectorXd& GetSBNParameters() const;
  double GetLogMarginalLikelihood() const;

  NucleotidePLV GetPLV(const PVId plv_index) const;
  void SetPLV(const PVId plv_index, const NucleotidePLV& plv, int rescaling_count = 0);
  NucleotidePLV GetSparePLV(const PVId plv_index) const { return GetPLV(GetSparePLVIndex(plv_index)); }
  PVId GetSparePLVIndex(const PVId plv_index) const;
  EigenVectorXi GetRescalingCounts() const;  // rescaling_counts_ (private in the reference)

  // ** Handlers (gp_engine.hpp:145-156)
  // PLVNodeHandler's index arithmetic (pv_handler.hpp:26-33, 227-238, 487-490) without its storage.
  class PLVIndexView {
   public:
    using PLVType = PLVNodeHandler::PLVType;
    explicit PLVIndexView(const BITO_B200_ENGINE_CLASS& engine) : engine_(engine) {}
    PVId GetPVIndex(const PLVType plv_type, const NodeId node_id) const {
      return PLVNodeHandler::GetPVIndex(plv_type, node_id, engine_.GetNodeCount());
    }
    PVId GetSparePVIndex(const PVId pv_id) const { return engine_.GetSparePLVIndex(pv_id); }
    size_t GetNodeCount() const { return engine_.GetNodeCount(); }
    size_t GetSpareNodeCount() const { return engine_.GetSpareNodeCount(); }
    size_t GetPVCount() const { return engine_.GetPLVCount(); }
    size_t GetSparePVCount() const { return engine_.GetSparePLVCount(); }
    size_t GetPaddedPVCount() const { return engine_.GetPaddedPLVCount(); }
    // GPInstance::PrintStatus (gp_instance.cpp:33-36) reports the PLV memory: here HBM actually held, in bytes
    double GetByteCount() const { return static_cast<double>(engine_.Stats().device_bytes_in_use); }

   private:
    const BITO_B200_ENGINE_CLASS& engine_;
  };
  const PLVIndexView& GetPLVHandler() const { return plv_view_; }
  DAGBranchHandler& GetBranchLengthHandler();
  const DAGBranchHandler& GetBranchLengthHandler() const;

  // ** Other Operations (gp_engine.hpp:158-182)
  EigenVectorXd CalculateQuartetHybridLikelihoods(const QuartetHybridRequest& request);
  void ProcessQuartetHybridRequest(const QuartetHybridRequest& request);
  // One launch for many requests (GPInstance::CalculateHybridMarginals issues one per edge).
  void ProcessQuartetHybridRequests(const std::vector<QuartetHybridRequest>& requests);
  SizeDoubleVectorMap GatherBranchLengths(const RootedTreeCollection& tree_collection,
                                          const BitsetSizeMap& indexer);
  void HotStartBranchLengths(const RootedTreeCollection& tree_collection,
                             const BitsetSizeMap& indexer);
  void TakeFirstBranchLength(const RootedTreeCollection& tree_collection,
                             const BitsetSizeMap& indexer);
  DoublePair LogLikelihoodAndDerivative(const GPOperations::OptimizeBranchLength& op);
  DoublePair LogLikelihoodAndDerivative(const size_t gpcsp, const size_t rootward,
                                        const size_t leafward);
  std::tuple<double, double, double> LogLikelihoodAndFirstTwoDerivatives(
      const GPOperations::OptimizeBranchLength& op);
  std::tuple<double, double, double> LogLikelihoodAndFirstTwoDerivatives(
      const size_t gpcsp, const size_t rootward, const size_t leafward);

  // ** I/O
  std::string PLVToString(const PVId plv_idx) const;
  std::string LogLikelihoodMatrixToString() const;

  // ** Counts (gp_engine.hpp:193-235)
  size_t GetPLVCountPerNode() const { return 6; }
  size_t GetSitePatternCount() const { return site_pattern_.PatternCount(); }
  size_t GetNodeCount() const;
  size_t GetSpareNodeCount() const;
  size_t GetPaddedNodeCount() const { return GetNodeCount() + GetSpareNodeCount(); }
  size_t GetPLVCount() const;
  size_t GetSparePLVCount() const { return GetPaddedPLVCount() - GetPLVCount(); }
  size_t GetPaddedPLVCount() const;
  size_t GetGPCSPCount() const;
  size_t GetSpareGPCSPCount() const { return GetPaddedGPCSPCount() - GetGPCSPCount(); }
  size_t GetPaddedGPCSPCount() const;
  size_t GetSpareGPCSPIndex(const size_t gpcsp_offset) const;

  static constexpr double default_rescaling_threshold_ = 1e-40;

  // ** B200 extras (no reference counterpart)
  bito_gp_engine* Handle() const { return handle_; }
  bito_gp_stats Stats() const;

 private:
  void Check(int rc) const;
  void SetBranchLengthsFromTotals(std::vector<double>& totals, const std::vector<int>& seen,
                                  bool mean);

  // The device handle for a call. While a DAGBranchHandler mirror is out, first uploads what the
  // caller changed through it.
  bito_gp_engine* H() const {
    if (mirror_live_) PushMirror();
    return handle_;
  }
  void PullMirror() const;  // device -> mirror (sizes, branch lengths, differences)
  void PushMirror() const;  // mirror -> device, if the caller wrote to it
  void AfterBranchLengthChange() const {
    if (mirror_live_) PullMirror();
  }

  SitePattern site_pattern_;
  bito_gp_engine* handle_ = nullptr;
  PLVIndexView plv_view_{*this};
  mutable EigenMatrixXd log_likelihood_matrix_copy_;
  mutable EigenVectorXd hybrid_marginals_copy_, sbn_parameters_copy_;
  mutable DAGBranchHandler branch_mirror_{0};
  mutable EigenVectorXd mirror_synced_;  // branch lengths as last exchanged with the device
  mutable bool mirror_live_ = false;
  Eigen::Matrix4d transition_matrix_;
  static constexpr double default_branch_length_ = 0.1;  // dag_branch_handler.hpp:266
};
