// bito_b200/host/tp_eval_engine_b200.hpp — the TP likelihood evaluator as a class the reference's TPEngine can own
// (SURVEY.md 8f row 4): TPEvalEngineViaLikelihood's virtual interface (tp_evaluation_engine.hpp:156-262) served by a
// GP engine running the op lists of tp_likelihood_plan.hpp.
//
// TPEngine owns a std::unique_ptr<TPEvalEngineViaLikelihood> (tp_engine.hpp:535) and builds it with make_unique
// (tp_engine.cpp:1322-1327), so the swap is a SUBCLASS of the reference evaluator:
//   * every entry point the NNI search in TP mode goes through is virtual in the reference and overridden here:
//     Initialize, ComputeScores, GetTopTreeScoreWithProposedNNI, UpdateEngineAfterModifyingDAG, GrowEdgeData (and
//     through it GrowSpareEdgeData / GrowEngineForDAG / GrowEngineForAdjacentNNIs);
//   * the base subobject keeps doing what is host bookkeeping in the reference: the DAGBranchHandler that TPEngine
//     and the NNI engine read and write by reference (tp_engine.cpp:454, 481, 1016-1051), the top-tree score
//     vector, PV / temp-edge ID arithmetic (GetProposedNNIInfo, :643-721). Its mmapped PVs are never computed on;
//   * branch lengths live on the host handler between calls: they are uploaded before an op list runs and read back
//     after one that optimises;
//   * BranchLengthOptimization() is NOT virtual in the reference: the method of the same name here runs on the
//     engine, but a caller holding a TPEvalEngineViaLikelihood& (TPEngine::OptimizeBranchLengths, tp_engine.cpp:1424)
//     reaches the reference's CPU code. One `virtual` upstream closes that.
// TPEngineWithEvaluator<Evaluator> installs the evaluator in a TPEngine (the members are protected, tp_engine.hpp:530-537).
//
// Engine is any class with the GPEngine surface: the reference CPU GPEngine (tests/cpp/tp_search_parity.cpp uses it to
// check this class bit for bit without a GPU) or GPEngineB200 (the CUDA engine).
#pragma once

#include <memory>
#include <optional>
#include <string>

#include "tp_engine.hpp"
#include "tp_evaluation_engine.hpp"
#include "tp_likelihood_plan.hpp"

template <class Engine>
class TPEvalEngineOverGPEngine : public TPEvalEngineViaLikelihood {
 public:
  using Reference = TPEvalEngineViaLikelihood;

  TPEvalEngineOverGPEngine(TPEngine& tp_engine, const std::string& mmap_path)
      : Reference(tp_engine, mmap_path + ".host"), mmap_path_(mmap_path) {
    MakeEngine();
  }

  // TPEvalEngineViaLikelihood::Initialize (:120-129)
  void Initialize() override {
    const TPLikelihoodPlan plan = Plan();
    Upload();
    engine_->ProcessOperations(plan.InitializeOps());
  }

  // :921-935. The reference writes the per-pattern log-likelihood ROWS of the requested edges and then recomputes the
  // score of EVERY edge as row . weights - and GrowEdgeData (:166-196) resizes that matrix without reindexing it, so
  // after the DAG has grown the score of an edge that was not refreshed is the one of whichever edge owned its row
  // index before. row_scores_ keeps one dot product per row to reproduce exactly that.
  void ComputeScores(std::optional<EdgeIdVector> opt_edge_ids = std::nullopt) override {
    const TPLikelihoodPlan plan = Plan();
    Upload();
    const size_t E = GetTPEngine().GetEdgeCount();
    const EdgeIdVector edge_ids =
        opt_edge_ids.has_value() ? opt_edge_ids.value() : GetDAG().LeafwardEdgeTraversalTrace(true);
    engine_->ProcessOperations(plan.ComputeScoresOps(edge_ids));
    const EigenVectorXd scores = engine_->GetPerGPCSPLogLikelihoods();
    SizeRowScores();
    for (const auto edge_id : edge_ids) row_scores_[edge_id.value_] = scores[edge_id.value_];
    GetTopTreeScores() = row_scores_.head(E);
  }

  // :466-641
  double GetTopTreeScoreWithProposedNNI(const NNIOperation& post_nni, const NNIOperation& pre_nni,
                                        const size_t spare_offset = 0,
                                        std::optional<BitsetEdgeIdMap> best_edge_map = std::nullopt) override {
    const ProposedNNIInfo info = GetProposedNNIInfo(post_nni, pre_nni, spare_offset, best_edge_map);
    const bool init_with_dag = IsInitProposedBranchLengthsWithDAG() || best_edge_map.has_value();
    auto& handler = GetDAGBranchHandler();
    TPLikelihoodPlan::InitializeTempBranchLengths(handler, info, handler.GetDefaultBranchLength(), init_with_dag);
    const TPLikelihoodPlan plan = Plan();
    const auto ops = plan.ProposedNNIOps(info, init_with_dag, IsFixProposedBranchLengthsFromDAG(), spare_offset, 0);
    Upload();
    engine_->ResetOptimizationCount();
    engine_->ProcessOperations(ops.initialize);
    if (IsOptimizeNewEdges()) {
      for (size_t iter = 0; iter < GetOptimizationMaxIteration(); ++iter) {
        engine_->ProcessOperations(ops.iteration);
        engine_->IncrementOptimizationCount();
      }
      Download();
    }
    engine_->ProcessOperations(ops.score);
    SizeRowScores();
    row_scores_[ops.focal_gpcsp] = engine_->GetPerGPCSPLogLikelihoods(ops.focal_gpcsp, 1)[0];  // the temp edge's row
    return engine_->GetPerGPCSPLogLikelihoods(ops.focal_gpcsp, 1)[0];
  }

  // :267-460 (the engine was grown by GrowEdgeData below; TPEngine::UpdateChoiceMapAfterModifyingDAG ran before this)
  void UpdateEngineAfterModifyingDAG(const std::map<NNIOperation, NNIOperation>& nni_to_pre_nni,
                                     const size_t prev_node_count, const Reindexer& node_reindexer,
                                     const size_t prev_edge_count, const Reindexer& edge_reindexer) override {
    std::ignore = prev_node_count;
    std::ignore = node_reindexer;
    const TPLikelihoodPlan plan = Plan();
    const auto ops = plan.UpdateAfterModifyingDAGOps(nni_to_pre_nni, prev_edge_count, edge_reindexer);
    Upload();
    engine_->ResetOptimizationCount();
    engine_->ProcessOperations(ops.initialize);
    if (IsOptimizeNewEdges()) {
      for (size_t iter = 0; iter < GetOptimizationMaxIteration(); ++iter) engine_->ProcessOperations(ops.iteration);
      Download();
    }
    engine_->ProcessOperations(ops.score);
    const EigenVectorXd scores = engine_->GetPerGPCSPLogLikelihoods();
    SizeRowScores();
    for (const auto edge_id : ops.update_edges) row_scores_[edge_id.value_] = scores[edge_id.value_];
    GetTopTreeScores() = row_scores_.head(GetTPEngine().GetEdgeCount());  // as ComputeScores(update_edges) does
  }

  void UpdateEngineAfterDAGAddNodePair(const NNIOperation&, const NNIOperation&, std::optional<size_t>) override {
    Failwith("TPEvalEngineOverGPEngine: UpdateEngineAfterDAGAddNodePair is not served by the engine; the NNI search "
             "updates through UpdateEngineAfterModifyingDAG.");
  }

  // :166-196. The base resizes and reindexes the host handler and the score vector; the engine's PVs and per-edge
  // data follow with the same reindexer (its "node" slots are taxa, then edges).
  void GrowEdgeData(const size_t edge_count, std::optional<const Reindexer> edge_reindexer = std::nullopt,
                    std::optional<const size_t> explicit_alloc = std::nullopt, const bool on_init = false) override {
    Reference::GrowEdgeData(edge_count, edge_reindexer, explicit_alloc, on_init);
    if (engine_ == nullptr) return;  // the base constructor's call: MakeEngine sizes the engine afterwards
    const size_t taxa = GetDAG().TaxonCount();
    if (edge_count != engine_edge_count_ || edge_reindexer.has_value()) {
      if (edge_reindexer.has_value()) {
        Reindexer node_reindexer = Reindexer::IdentityReindexer(taxa + edge_reindexer.value().size());
        for (size_t i = 0; i < edge_reindexer.value().size(); ++i)
          node_reindexer.SetReindex(taxa + i, taxa + edge_reindexer.value().GetNewIndexByOldIndex(i));
        engine_->GrowPLVs(taxa + edge_count, node_reindexer);
        engine_->GrowGPCSPs(edge_count, edge_reindexer.value());
      } else {
        engine_->GrowPLVs(taxa + edge_count);
        engine_->GrowGPCSPs(edge_count);
      }
      engine_edge_count_ = edge_count;
    }
    GrowEngineSpares();
  }

  // The whole-DAG optimisation (:988-1001) on the engine. Hides, does not override, the reference's method.
  void BranchLengthOptimization(std::optional<bool> check_branch_convergence = std::nullopt) {
    const bool check = check_branch_convergence.has_value() ? check_branch_convergence.value() : !IsFirstOptimization();
    const TPLikelihoodPlan plan = Plan();
    const auto ops = plan.BranchLengthOptimizationOps();
    Upload();
    engine_->ResetOptimizationCount();
    if (check) engine_->IncrementOptimizationCount();  // the engine skips converged edges iff its count is not 0
    for (size_t round = 0; round < GetOptimizationMaxIteration(); ++round) {
      engine_->ProcessOperations(ops);
      IncrementOptimizationCount();
    }
    Download();
  }

  Engine& GetEngine() { return *engine_; }

 private:
  TPLikelihoodPlan Plan() const { return TPLikelihoodPlan(GetDAG(), GetTPEngine().GetChoiceMap()); }

  void MakeEngine() {
    const TPLikelihoodPlan plan = Plan();
    const size_t N = plan.EngineNodeCount(), G = plan.EngineGPCSPCount();
    const EigenVectorXd ones_g = EigenVectorXd::Ones(G), ones_n = EigenVectorXd::Ones(N);
    engine_ = std::make_unique<Engine>(SitePattern(GetSitePattern()), N, G, mmap_path_ + ".gp", 1e-40, ones_g, ones_n,
                                       ones_g, false);
    engine_->SetNullPrior();  // TP's evolve carries no prior weight
    engine_edge_count_ = G;
    GrowEngineSpares();
  }
  // temp PVs of one proposed NNI (18, tp_likelihood_plan.hpp) and as many temp edges as the TP engine keeps
  void GrowEngineSpares() {
    engine_->GrowSparePLVs(TPLikelihoodPlan::SpareNodesForBatch(1));
    engine_->GrowSpareGPCSPs(std::max<size_t>(GetTPEngine().GetSpareEdgeCount(), TPLikelihoodPlan::SpareGPCSPsForBatch(1)));
  }
  // host handler -> engine: DAG edges in one call, temp edges through the engine's mirror
  void Upload() {
    engine_->SetNullPrior();  // q = 1 on every edge, the temp edges grown since the last call included
    const auto& host = GetDAGBranchHandler().GetBranchLengthData();
    const size_t E = engine_->GetGPCSPCount();
    engine_->SetBranchLengths(host.head(E));
    auto& mirror = engine_->GetBranchLengthHandler();
    const size_t padded = std::min<size_t>(engine_->GetPaddedGPCSPCount(), host.size());
    for (size_t e = E; e < padded; ++e) mirror(EdgeId(e)) = host[e];
  }
  // engine -> host handler, after an op list that optimised
  void Download() {
    auto& host = GetDAGBranchHandler().GetBranchLengthData();
    const size_t padded = std::min<size_t>(engine_->GetPaddedGPCSPCount(), host.size());
    host.head(padded) = engine_->GetBranchLengths(0, padded);
  }

  // one entry per row of the reference's log-likelihood matrix (edges and temp edges), never reindexed
  void SizeRowScores() {
    const size_t rows = std::max<size_t>(GetMatrix().rows(), GetTPEngine().GetPaddedEdgeCount());
    if (size_t(row_scores_.size()) < rows) {
      const Eigen::Index old = row_scores_.size();
      row_scores_.conservativeResize(rows);
      row_scores_.tail(rows - old).setZero();
    }
  }

  std::string mmap_path_;
  EigenVectorXd row_scores_;
  std::unique_ptr<Engine> engine_;
  size_t engine_edge_count_ = 0;
};

// A TPEngine whose likelihood evaluator is `Evaluator` (a subclass of TPEvalEngineViaLikelihood). TPEvalEngine has
// no virtual destructor, so the object is destroyed through its own type.
template <class Evaluator>
class TPEngineWithEvaluator : public TPEngine {
 public:
  TPEngineWithEvaluator(GPDAG& dag, SitePattern& site_pattern, const std::string& mmap_likelihood_path,
                        const std::string& mmap_parsimony_path, const RootedTreeCollection& tree_collection,
                        const BitsetSizeMap& edge_indexer)
      : TPEngine(dag, site_pattern, mmap_likelihood_path + ".ref", mmap_parsimony_path, tree_collection, edge_indexer) {
    likelihood_engine_.reset();  // the reference evaluator the base constructor made
    evaluator_ = new Evaluator(*this, mmap_likelihood_path);
    likelihood_engine_.reset(evaluator_);
    eval_engine_ = evaluator_;
    eval_engine_in_use_[TPEvalEngineType::LikelihoodEvalEngine] = true;
    // what the base constructor does once its evaluators exist (tp_engine.cpp:42-45)
    GrowNodeData(GetDAG().NodeCount(), std::nullopt, std::nullopt, true);
    GrowEdgeData(GetDAG().EdgeCountWithLeafSubsplits(), std::nullopt, std::nullopt, true);
    InitializeScores();
  }
  ~TPEngineWithEvaluator() {
    likelihood_engine_.release();
    delete evaluator_;
  }
  Evaluator& GetEvaluator() { return *evaluator_; }

 private:
  Evaluator* evaluator_ = nullptr;
};
