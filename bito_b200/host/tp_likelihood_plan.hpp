// bito_b200/host/tp_likelihood_plan.hpp — the top-pruning (TP) likelihood evaluator as an op list
// for the GP engine.
//
// The reference's TPEvalEngineViaLikelihood (/root/reference/src/tp_evaluation_engine.cpp:120-158,
// 804-935, 1024-1155) keeps six partial vectors per DAG EDGE and fills them along the edge choices
// of a TPChoiceMap (one parent / sister / left child / right child edge per edge) with the same three
// primitives the GP engine runs per node: evolve a PV along an edge, multiply two PVs, and the
// per-pattern log-likelihood of an edge from its rootward and leafward PVs. So the whole
// Initialize() + ComputeScores() of that class is a GPOperationVector over an engine that is sized
// by edges instead of nodes, and the hand-written CUDA kernels run it unchanged:
//
//   engine "node" slots  = taxon_count + edge_count   (slot t < taxon_count: the taxon's site
//                                                       patterns, InitializePLVsWithSitePatterns;
//                                                       slot taxon_count + e: edge e)
//   engine GPCSP ids     = the DAG's edge ids (branch lengths, log-likelihood rows)
//   q (SBN prior)        = 1 everywhere (SetNullPrior): TP's evolve has no prior weight
//
// After ProcessOperations(InitializeOps()) and ProcessOperations(ComputeScoresOps()),
// GetPerGPCSPLogLikelihoods() holds what TPEngine::GetTopTreeLikelihoods() holds
// (tp_evaluation_engine.cpp:921-935): the log-likelihood of the best tree through every edge.
//
// ProposedNNIOps() does the same for GetTopTreeScoreWithProposedNNI (:466-641): the score of an NNI
// that is NOT in the DAG, computed in spare PVs / spare edges from the PVs of its pre-NNI, with the
// five edges around it optionally optimised (OptimizeBranchLength ops) for a few rounds.
#pragma once

#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "gp_dag.hpp"
#include "gp_operation.hpp"
#include "nni_operation.hpp"
#include "pv_handler.hpp"
#include "reindexer.hpp"
#include "tp_choice_map.hpp"
#include "tp_evaluation_engine.hpp"

class TPLikelihoodPlan {
 public:
  using PLVType = PLVNodeHandler::PLVType;

  TPLikelihoodPlan(const GPDAG& dag, const TPChoiceMap& choice_map)
      : dag_(dag), choice_map_(choice_map), taxon_count_(dag.TaxonCount()),
        edge_count_(dag.EdgeCountWithLeafSubsplits()) {}

  // How to size the engine that runs the lists.
  size_t EngineNodeCount() const { return taxon_count_ + edge_count_; }
  size_t EngineGPCSPCount() const { return edge_count_; }

  // PV of `type` of DAG edge `edge_id` as an engine PLV id. The P-PV of a leaf edge is its taxon's
  // site-pattern PLV (PopulateLeafPVsWithSitePatterns, :868-899, copies it into every leaf edge of
  // the taxon; here those edges read the one copy the engine already holds).
  size_t PV(const PLVType type, const EdgeId edge_id) const {
    if (type == PLVType::P && dag_.IsEdgeLeaf(edge_id)) {
      const NodeId leaf = dag_.GetDAGEdge(edge_id).GetChild();
      return PLVNodeHandler::GetPVIndex(PLVType::P, leaf, EngineNodeCount()).value_;  // leaf node id = taxon id
    }
    return PLVNodeHandler::GetPVIndex(type, NodeId(taxon_count_ + edge_id.value_), EngineNodeCount()).value_;
  }

  // TPEvalEngineViaLikelihood::Initialize (:120-129) minus the leaf copies.
  GPOperationVector InitializeOps() const {
    using namespace GPOperations;
    GPOperationVector ops;
    // PopulateRootPVsWithStationaryDistribution (:901-919)
    for (const auto edge_id : dag_.GetRootsplitEdgeIds())
      ops.push_back(SetToStationaryDistribution{PV(PLVType::RHat, edge_id), edge_id.value_});
    // PopulateRootwardPVs (:146-151, 804-838)
    for (const auto node_id : dag_.RootwardNodeTraversalTrace(false))
      for (const auto clade : {SubsplitClade::Left, SubsplitClade::Right})
        for (const auto adj : dag_.GetDAGNode(node_id).GetNeighbors(Direction::Rootward, clade))
          RootwardForEdge(ops, dag_.GetEdgeIdx(adj, node_id));
    // PopulateLeafwardPVs (:153-158, 840-866)
    for (const auto node_id : dag_.LeafwardNodeTraversalTrace(true))
      for (const auto clade : {SubsplitClade::Left, SubsplitClade::Right})
        for (const auto adj : dag_.GetDAGNode(node_id).GetNeighbors(Direction::Leafward, clade))
          LeafwardForEdge(ops, dag_.GetEdgeIdx(node_id, adj));
    return ops;
  }

  // TPEvalEngineViaLikelihood::ComputeScores (:921-935, 1042-1055, 1134-1140).
  GPOperationVector ComputeScoresOps() const {
    using namespace GPOperations;
    GPOperationVector ops;
    for (const auto edge_id : dag_.LeafwardEdgeTraversalTrace(true)) {
      const auto& choices = choice_map_.GetEdgeChoice(edge_id);
      const size_t parent_pv =
          choices.parent == NoId
              ? PV(PLVType::RHat, dag_.GetFirstRootsplitEdgeId())
              : PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(edge_id)), choices.parent);
      ops.push_back(Likelihood{edge_id.value_, PV(PLVType::P, edge_id), parent_pv});
    }
    return ops;
  }

  // ComputeScores(opt_edge_ids): the listed edges only, in the caller's order (:921-935).
  GPOperationVector ComputeScoresOps(const EdgeIdVector& edge_ids) const {
    using namespace GPOperations;
    GPOperationVector ops;
    for (const auto edge_id : edge_ids) {
      const auto& choices = choice_map_.GetEdgeChoice(edge_id);
      const size_t parent_pv =
          choices.parent == NoId
              ? PV(PLVType::RHat, dag_.GetFirstRootsplitEdgeId())
              : PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(edge_id)), choices.parent);
      ops.push_back(Likelihood{edge_id.value_, PV(PLVType::P, edge_id), parent_pv});
    }
    return ops;
  }

  // One round of TPEvalEngineViaLikelihood::BranchLengthOptimization (:988-1022) over every edge below a
  // rootsplit edge, rootward order: refresh the PVs around the edge, optimise it, push the result down.
  // The reference decides ONCE per call whether converged edges are skipped (check_branch_convergence =
  // !IsFirstOptimization() before its loop of GetOptimizationMaxIteration() rounds), so run this list
  // that many times and only then call IncrementOptimizationCount() as many times.
  GPOperationVector BranchLengthOptimizationOps() const {
    using namespace GPOperations;
    GPOperationVector ops;
    for (const auto edge_id : dag_.RootwardEdgeTraversalTrace(false)) {
      const auto& choices = choice_map_.GetEdgeChoice(edge_id);
      if (choices.parent == NoId) continue;
      RootwardForEdge(ops, edge_id);
      RootwardForEdge(ops, choices.parent);
      LeafwardForEdge(ops, choices.parent);
      ops.push_back(OptimizeBranchLength{PV(PLVType::P, edge_id),
                                         PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(edge_id)), choices.parent),
                                         edge_id.value_});
      LeafwardForEdge(ops, edge_id);
    }
    return ops;
  }

  // TPEvalEngineViaLikelihood::UpdateEngineAfterModifyingDAG (:267-460): the INCREMENTAL update after the NNI
  // search has added node pairs to the DAG. It is not a re-evaluation - only the edges around the new NNIs
  // are refreshed, in the reference's own order - so it is replayed step by step:
  //   initialize: rootsplit RHat PVs, then PopulateRootwardPVForEdge over `update_edges` sorted by parent
  //               node, PopulateLeafwardPVForEdge over them sorted by child node, descending (:306-330, 411-412);
  //   iteration : per NNI OptimizeEdge on left child, right child, sister, focal and (below the root) parent
  //               edge, then OptimizeEdge on every other new edge, then both local passes per NNI
  //               (:415-447); only NEW edges are optimised (:364-367), never skipped for convergence;
  //   score     : ComputeScores over `update_edges` (:455-457).
  // This plan is built for the GROWN DAG and choice map (after UpdateChoiceMapAfterModifyingDAG, which also
  // gives the new edges their starting branch lengths); the engine must have been grown with the edge
  // reindexer (EngineNodeReindexer below) and hold those branch lengths.
  // Run: ResetOptimizationCount; initialize; iteration x GetOptimizationMaxIteration() if IsOptimizeNewEdges(); score.
  struct ModifiedDAGUpdate {
    GPOperationVector initialize, iteration, score;
    std::vector<EdgeId> update_edges;
  };
  ModifiedDAGUpdate UpdateAfterModifyingDAGOps(const std::map<NNIOperation, NNIOperation>& nni_to_pre_nni,
                                               const size_t prev_edge_count, const Reindexer& edge_reindexer) const {
    using namespace GPOperations;
    ModifiedDAGUpdate out;
    std::set<EdgeId> new_edges, nni_edges, extra_edges, update_edges;
    for (size_t i = prev_edge_count; i < edge_reindexer.size(); i++) {
      const EdgeId edge_id = EdgeId(edge_reindexer.GetNewIndexByOldIndex(i));
      new_edges.insert(edge_id);
      extra_edges.insert(edge_id);
      update_edges.insert(edge_id);
    }
    for (const auto& [post_nni, pre_nni] : nni_to_pre_nni) {
      std::ignore = pre_nni;
      const auto edge_id = dag_.GetEdgeIdx(post_nni);
      const auto& choice = choice_map_.GetEdgeChoice(edge_id);
      nni_edges.insert(edge_id);
      for (const EdgeId e : {choice.right_child, choice.left_child, choice.sister, edge_id, choice.parent}) {
        extra_edges.erase(e);
        update_edges.insert(e);
      }
    }
    // the reference's (unstable) sorts on the same inputs with the same comparators
    std::vector<EdgeId> rootward_edges(update_edges.begin(), update_edges.end());
    std::sort(rootward_edges.begin(), rootward_edges.end(), [this](const EdgeId lhs, const EdgeId rhs) {
      return dag_.GetDAGEdge(lhs).GetParent() < dag_.GetDAGEdge(rhs).GetParent();
    });
    std::vector<EdgeId> leafward_edges(update_edges.begin(), update_edges.end());
    std::sort(leafward_edges.begin(), leafward_edges.end(), [this](const EdgeId lhs, const EdgeId rhs) {
      return dag_.GetDAGEdge(lhs).GetChild() > dag_.GetDAGEdge(rhs).GetChild();
    });
    out.update_edges.assign(update_edges.begin(), update_edges.end());
    // PopulateRootPVsWithStationaryDistribution (:901-919); the leaf P-PVs are the engine's site patterns
    for (const auto edge_id : dag_.GetRootsplitEdgeIds())
      out.initialize.push_back(SetToStationaryDistribution{PV(PLVType::RHat, edge_id), edge_id.value_});
    for (const auto edge_id : rootward_edges) RootwardForEdge(out.initialize, edge_id);
    for (const auto edge_id : leafward_edges) LeafwardForEdge(out.initialize, edge_id);

    auto& it = out.iteration;
    auto optimize_edge = [&](const EdgeId edge_id, const EdgeId parent_edge_id, const bool is_not_child_edge,
                             const bool is_not_parent_edge) {  // OptimizeEdge, :332-383
      const auto focal = dag_.GetFocalClade(edge_id);
      const auto sister = dag_.GetSisterClade(edge_id);
      if (is_not_child_edge)
        it.push_back(Multiply{PV(PLVType::P, edge_id), PV(PLVType::PHatLeft, edge_id), PV(PLVType::PHatRight, edge_id)});
      if (is_not_parent_edge) {
        Assert(!dag_.IsEdgeRoot(edge_id), "TPLikelihoodPlan: OptimizeEdge on a root edge with a parent edge.");
        it.push_back(Multiply{PV(PLVTypeEnum::RPLVType(focal), parent_edge_id), PV(PLVType::RHat, parent_edge_id),
                              PV(PLVTypeEnum::PPLVType(sister), parent_edge_id)});
      }
      if (new_edges.find(edge_id) != new_edges.end()) {
        const auto& choices = choice_map_.GetEdgeChoice(edge_id);  // GetPrimaryPVIdsOfEdge, :1042-1055
        const size_t parent_pv = choices.parent == NoId
                                     ? PV(PLVType::RHat, dag_.GetFirstRootsplitEdgeId())
                                     : PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(edge_id)), choices.parent);
        it.push_back(OptimizeBranchLength{PV(PLVType::P, edge_id), parent_pv, edge_id.value_});
      }
      if (is_not_parent_edge) {
        Evolve(it, PV(PLVTypeEnum::PPLVType(focal), parent_edge_id), edge_id, PV(PLVType::P, edge_id));
        it.push_back(Multiply{PV(PLVType::P, parent_edge_id), PV(PLVType::PHatLeft, parent_edge_id),
                              PV(PLVType::PHatRight, parent_edge_id)});
      }
    };
    for (const auto edge_id : nni_edges) {  // :417-432
      const auto& choice = choice_map_.GetEdgeChoice(edge_id);
      optimize_edge(choice.left_child, edge_id, false, true);
      optimize_edge(choice.right_child, edge_id, false, true);
      optimize_edge(choice.sister, choice.parent, false, true);
      optimize_edge(edge_id, choice.parent, true, true);
      if (!dag_.IsEdgeRoot(choice.parent)) {
        const auto& choice_2 = choice_map_.GetEdgeChoice(choice.parent);
        optimize_edge(choice.parent, choice_2.parent, true, false);
      }
    }
    for (const auto edge_id : extra_edges) {  // :433-441
      const auto& choice = choice_map_.GetEdgeChoice(edge_id);
      if (!dag_.IsEdgeRoot(choice.parent)) optimize_edge(edge_id, choice.parent, true, true);
    }
    for (const auto edge_id : nni_edges) {  // NNIUpdatePVs, :384-409, 443
      const auto& choice = choice_map_.GetEdgeChoice(edge_id);
      const auto focal = dag_.GetFocalClade(edge_id), sister = dag_.GetSisterClade(edge_id);
      // NNIRootwardPass: GetLocalPVIdsOfEdge (:1057-1097) names the PVs around the edge
      Evolve(it, PV(PLVType::PHatLeft, edge_id), choice.left_child, PV(PLVType::P, choice.left_child));
      Evolve(it, PV(PLVType::PHatRight, edge_id), choice.right_child, PV(PLVType::P, choice.right_child));
      it.push_back(Multiply{PV(PLVType::P, edge_id), PV(PLVType::PHatLeft, edge_id), PV(PLVType::PHatRight, edge_id)});
      Evolve(it, PV(PLVTypeEnum::PPLVType(sister), choice.parent), choice.sister, PV(PLVType::P, choice.sister));
      Evolve(it, PV(PLVTypeEnum::PPLVType(focal), choice.parent), edge_id, PV(PLVType::P, edge_id));
      it.push_back(Multiply{PV(PLVType::P, choice.parent), PV(PLVTypeEnum::PPLVType(focal), choice.parent),
                            PV(PLVTypeEnum::PPLVType(sister), choice.parent)});
      // NNILeafwardPass
      if (!dag_.IsEdgeRoot(choice.parent)) {
        const auto& choice_2 = choice_map_.GetEdgeChoice(choice.parent);
        Evolve(it, PV(PLVType::RHat, choice.parent), choice.parent,
               PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(choice.parent)), choice_2.parent));
      }
      it.push_back(Multiply{PV(PLVTypeEnum::RPLVType(focal), choice.parent), PV(PLVType::RHat, choice.parent),
                            PV(PLVTypeEnum::PPLVType(sister), choice.parent)});
      it.push_back(Multiply{PV(PLVTypeEnum::RPLVType(sister), choice.parent), PV(PLVType::RHat, choice.parent),
                            PV(PLVTypeEnum::PPLVType(focal), choice.parent)});
      Evolve(it, PV(PLVType::RHat, edge_id), edge_id, PV(PLVTypeEnum::RPLVType(focal), choice.parent));
      it.push_back(Multiply{PV(PLVType::RLeft, edge_id), PV(PLVType::RHat, edge_id), PV(PLVType::PHatRight, edge_id)});
      it.push_back(Multiply{PV(PLVType::RRight, edge_id), PV(PLVType::RHat, edge_id), PV(PLVType::PHatLeft, edge_id)});
    }
    for (const auto edge_id : out.update_edges) {  // ComputeScores(update_edges), :921-935
      const auto& choices = choice_map_.GetEdgeChoice(edge_id);
      const size_t parent_pv = choices.parent == NoId
                                   ? PV(PLVType::RHat, dag_.GetFirstRootsplitEdgeId())
                                   : PV(PLVTypeEnum::RPLVType(dag_.GetFocalClade(edge_id)), choices.parent);
      out.score.push_back(Likelihood{edge_id.value_, PV(PLVType::P, edge_id), parent_pv});
    }
    return out;
  }
  // The engine's "node" slots are taxa, then edges: growing it for a DAG whose edges were reindexed by
  // `edge_reindexer` (TPEvalEngine::GrowEdgeData, :34-49) is GrowPLVs(EngineNodeCount(), this) +
  // GrowGPCSPs(EngineGPCSPCount(), edge_reindexer) on the plan of the grown DAG.
  Reindexer EngineNodeReindexer(const Reindexer& edge_reindexer) const {
    Reindexer out = Reindexer::IdentityReindexer(taxon_count_ + edge_reindexer.size());
    for (size_t i = 0; i < edge_reindexer.size(); ++i)
      out.SetReindex(taxon_count_ + i, taxon_count_ + edge_reindexer.GetNewIndexByOldIndex(i));
    return out;
  }

  // TP PV id (PLVEdgeHandler numbering, pv_handler.hpp:487-490, 227-238: type * E + edge, spare j at
  // 6 E + j) -> engine PLV id (spare j -> the engine's spare PLV j).
  size_t EnginePV(const PVId tp_pv) const { return EnginePV(tp_pv, 0, 0); }
  // Same, for the `slot`-th NNI of a batch whose info was built with spare offset `tp_spare_offset`: the
  // reference hands NNI i the temp PVs [12 i, 12 i + 18) (GetTempLocalPVIdsForProposedNNIs, :742-776:
  // 18 ids at a stride of spare_nodes_per_nni_ = 12, so neighbours overlap - harmless there, one NNI is scored
  // at a time); a batch needs them disjoint, so temp j of NNI `slot` becomes engine spare PLV 18 slot + j.
  static constexpr size_t kTempPVsPerNNI = 18, kReferenceTempStride = 12;
  size_t EnginePV(const PVId tp_pv, const size_t tp_spare_offset, const size_t slot) const {
    const size_t v = tp_pv.value_;
    if (v >= 6 * edge_count_) {
      const size_t j = (v - 6 * edge_count_) - kReferenceTempStride * tp_spare_offset;
      return 6 * EngineNodeCount() + kTempPVsPerNNI * slot + j;
    }
    return PV(static_cast<PLVType>(v / edge_count_), EdgeId(v % edge_count_));
  }
  // Spare PLVs (in engine "nodes" of six PLVs) and spare GPCSPs an engine needs to score n NNIs in one batch.
  static size_t SpareNodesForBatch(const size_t n) { return (kTempPVsPerNNI * n + 5) / 6; }
  static size_t SpareGPCSPsForBatch(const size_t n) { return 5 * n; }

  // The three op lists of one proposed NNI, and which of its five edges get optimised.
  struct ProposedNNI {
    GPOperationVector initialize;  // RootwardPass + LeafwardPass (:501-537)
    GPOperationVector iteration;   // one round of OptimizeLeftChild .. OptimizeParent + both passes (:615-630)
    GPOperationVector score;       // ComputeLikelihood of the focal edge (:634-636)
    size_t focal_gpcsp;            // row that then holds the score: GetPerGPCSPLogLikelihoods(focal_gpcsp, 1)
    NNIAdjBools do_optimize_edge;
  };
  // `info` = TPEvalEngineViaLikelihood::GetProposedNNIInfo(post_nni, pre_nni, spare_offset) (:643-721);
  // the two flags are the evaluator's do_init_proposed_branch_lengths_with_dag_ /
  // do_fix_proposed_branch_lengths_from_dag_ (both true by default, tp_evaluation_engine.hpp:437-439).
  // Run: ResetOptimizationCount; initialize; [iteration; IncrementOptimizationCount] x max_iter; score.
  ProposedNNI ProposedNNIOps(const ProposedNNIInfo& info, const bool init_with_dag = true,
                             const bool fix_from_dag = true, const size_t tp_spare_offset = 0,
                             const size_t slot = 0) const {
    using namespace GPOperations;
    using Adj = NNIAdjacent;
    ProposedNNI out;
    auto EnginePV = [this, tp_spare_offset, slot](const PVId id) {  // temps of this NNI's slot
      return this->EnginePV(id, tp_spare_offset, slot);
    };
    const auto& t = info.temp_pv_ids;
    const auto& r = info.ref_pv_ids;
    const auto& te = info.temp_edge_ids;
    out.focal_gpcsp = te.focal.value_;
    out.do_optimize_edge = info.do_optimize_edge;
    for (auto adj : NNIAdjacentEnum::Iterator())  // :480-497
      if (init_with_dag && info.adj_edge_ids[adj] != NoId && fix_from_dag) out.do_optimize_edge[adj] = false;
    // parent_rhat: evolved down from the grandparent, or - below the DAG root - the reference PV itself
    // (TakePVValue, :524-529; the temp copy is only ever read, so the plan reads the original)
    const bool has_grandparent = r.grandparent_rfocal_ != NoId;
    const size_t parent_rhat = has_grandparent ? EnginePV(t.parent_rhat_) : EnginePV(r.parent_rhat_);
    auto rootward_pass = [&](GPOperationVector& ops) {  // :501-515
      Evolve(ops, EnginePV(t.child_phatleft_), te.left_child, EnginePV(r.leftchild_p_));
      Evolve(ops, EnginePV(t.child_phatright_), te.right_child, EnginePV(r.rightchild_p_));
      ops.push_back(Multiply{EnginePV(t.child_p_), EnginePV(t.child_phatleft_), EnginePV(t.child_phatright_)});
      Evolve(ops, EnginePV(t.parent_phatsister_), te.sister, EnginePV(r.sister_p_));
      Evolve(ops, EnginePV(t.parent_phatfocal_), te.focal, EnginePV(t.child_p_));
      ops.push_back(Multiply{EnginePV(t.parent_p_), EnginePV(t.parent_phatfocal_), EnginePV(t.parent_phatsister_)});
    };
    auto leafward_pass = [&](GPOperationVector& ops) {  // :516-537
      if (has_grandparent) Evolve(ops, parent_rhat, te.parent, EnginePV(r.grandparent_rfocal_));
      ops.push_back(Multiply{EnginePV(t.parent_rfocal_), parent_rhat, EnginePV(t.parent_phatsister_)});
      ops.push_back(Multiply{EnginePV(t.parent_rsister_), parent_rhat, EnginePV(t.parent_phatfocal_)});
      Evolve(ops, EnginePV(t.child_rhat_), te.focal, EnginePV(t.parent_rfocal_));
      ops.push_back(Multiply{EnginePV(t.child_rleft_), EnginePV(t.child_rhat_), EnginePV(t.child_phatright_)});
      ops.push_back(Multiply{EnginePV(t.child_rright_), EnginePV(t.child_rhat_), EnginePV(t.child_phatleft_)});
    };
    // OptimizeEdge (:540-569); PV arguments already as engine ids, kNone where the reference passes NoId
    constexpr size_t kNone = static_cast<size_t>(-1);
    auto optimize_edge = [&](GPOperationVector& ops, EdgeId edge, size_t parent_p, size_t parent_phatfocal,
                             size_t parent_phatsister, size_t prhat, size_t parent_rfocal, size_t child_p,
                             size_t child_phatleft, size_t child_phatright, bool update, bool not_child_edge,
                             bool not_parent_edge) {
      if (not_child_edge) ops.push_back(Multiply{child_p, child_phatleft, child_phatright});
      if (not_parent_edge) ops.push_back(Multiply{parent_rfocal, prhat, parent_phatsister});
      if (update) ops.push_back(OptimizeBranchLength{child_p, parent_rfocal, edge.value_});
      if (not_parent_edge) {
        Evolve(ops, parent_phatfocal, edge, child_p);
        ops.push_back(Multiply{parent_p, parent_phatfocal, parent_phatsister});
      }
    };
    rootward_pass(out.initialize);
    leafward_pass(out.initialize);
    auto& it = out.iteration;
    const auto& opt = out.do_optimize_edge;
    optimize_edge(it, te.left_child, EnginePV(t.child_p_), EnginePV(t.child_phatleft_), EnginePV(t.child_phatright_),
                  EnginePV(t.child_rhat_), EnginePV(t.child_rleft_), EnginePV(r.leftchild_p_), kNone, kNone,
                  opt[Adj::LeftChild], false, true);  // :571-578
    optimize_edge(it, te.right_child, EnginePV(t.child_p_), EnginePV(t.child_phatright_), EnginePV(t.child_phatleft_),
                  EnginePV(t.child_rhat_), EnginePV(t.child_rright_), EnginePV(r.rightchild_p_), kNone, kNone,
                  opt[Adj::RightChild], false, true);  // :579-586
    optimize_edge(it, te.sister, EnginePV(t.parent_p_), EnginePV(t.parent_phatsister_), EnginePV(t.parent_phatfocal_),
                  parent_rhat, EnginePV(t.parent_rsister_), EnginePV(r.sister_p_), kNone, kNone, opt[Adj::Sister],
                  false, true);  // :587-594
    optimize_edge(it, te.focal, EnginePV(t.parent_p_), EnginePV(t.parent_phatfocal_), EnginePV(t.parent_phatsister_),
                  parent_rhat, EnginePV(t.parent_rfocal_), EnginePV(t.child_p_), EnginePV(t.child_phatleft_),
                  EnginePV(t.child_phatright_), opt[Adj::Focal], true, true);  // :595-603
    if (!info.post_nni.GetParent().SubsplitIsRootsplit() && te.parent != NoId && has_grandparent)
      optimize_edge(it, te.parent, kNone, kNone, kNone, kNone, EnginePV(r.grandparent_rfocal_), EnginePV(t.parent_p_),
                    EnginePV(t.parent_phatfocal_), EnginePV(t.parent_phatsister_), opt[Adj::Parent], true,
                    false);  // :604-612, 621-625
    rootward_pass(it);
    leafward_pass(it);
    out.score.push_back(Likelihood{te.focal.value_, EnginePV(t.child_p_), EnginePV(t.parent_rfocal_)});
    return out;
  }
  // All the NNIs adjacent to the DAG in ONE set of lists (what NNIEvalEngineViaTP::ScoreAdjacentNNIs,
  // nni_evaluation_engine.cpp:1075-1086, does one NNI at a time): the NNIs are independent - each works in
  // its own temp PVs and temp edges - so the engine's level scheduler runs the k-th step of every NNI as one
  // batch of kernels, and a whole round of scoring costs the launches of a single NNI.
  // infos[i] = GetProposedNNIInfo(post_i, pre_i, i) (spare offset i: disjoint temp EDGES, 5 per NNI).
  // Run as for one NNI; the score of NNI i is then GetPerGPCSPLogLikelihoods(focal_gpcsp[i], 1).
  struct ProposedNNIBatch {
    GPOperationVector initialize, iteration, score;
    std::vector<size_t> focal_gpcsp;
  };
  ProposedNNIBatch BatchedProposedNNIOps(const std::vector<ProposedNNIInfo>& infos, const bool init_with_dag = true,
                                         const bool fix_from_dag = true) const {
    ProposedNNIBatch batch;
    for (size_t i = 0; i < infos.size(); ++i) {
      const ProposedNNI one = ProposedNNIOps(infos[i], init_with_dag, fix_from_dag, i, i);
      batch.initialize.insert(batch.initialize.end(), one.initialize.begin(), one.initialize.end());
      batch.iteration.insert(batch.iteration.end(), one.iteration.begin(), one.iteration.end());
      batch.score.insert(batch.score.end(), one.score.begin(), one.score.end());
      batch.focal_gpcsp.push_back(one.focal_gpcsp);
    }
    return batch;
  }

  // Branch lengths the reference gives the five temp edges before it scores (:480-497): the default,
  // or the pre-NNI's (reference) edge, or - when the edge already exists in the DAG - that edge's.
  template <typename BranchHandler>
  static NNIAdjDoubles TempBranchLengths(const BranchHandler& handler, const ProposedNNIInfo& info,
                                         const double default_branch_length, const bool init_with_dag = true) {
    NNIAdjDoubles out;
    for (auto adj : NNIAdjacentEnum::Iterator()) {
      double value = default_branch_length;
      if (init_with_dag) {
        if (info.ref_edge_ids[adj] != NoId) value = handler(info.ref_edge_ids[adj]);
        if (info.adj_edge_ids[adj] != NoId) value = handler(info.adj_edge_ids[adj]);
      }
      out[adj] = value;
    }
    return out;
  }
  template <typename BranchHandler>
  static void InitializeTempBranchLengths(BranchHandler& handler, const ProposedNNIInfo& info,
                                          const double default_branch_length, const bool init_with_dag = true) {
    const NNIAdjDoubles values = TempBranchLengths(handler, info, default_branch_length, init_with_dag);
    for (auto adj : NNIAdjacentEnum::Iterator()) handler(info.temp_edge_ids[adj]) = values[adj];
  }

 private:
  // dest = M(t_edge) src, SetToEvolvedPV (:1142-1146); q is 1, so the weighted increment is the evolve
  void Evolve(GPOperationVector& ops, size_t dest, EdgeId edge_id, size_t src) const {
    using namespace GPOperations;
    ops.push_back(ZeroPLV{dest});
    ops.push_back(IncrementWithWeightedEvolvedPLV{dest, edge_id.value_, src});
  }
  void RootwardForEdge(GPOperationVector& ops, const EdgeId edge_id) const {  // :814-838
    using namespace GPOperations;
    const auto& choices = choice_map_.GetEdgeChoice(edge_id);
    for (const EdgeId child_edge : {choices.left_child, choices.right_child}) {
      if (child_edge == NoId) continue;
      const auto focal = dag_.GetFocalClade(child_edge);  // EvolvePPVUpEdge, :1024-1031
      Evolve(ops, PV(PLVTypeEnum::PPLVType(focal), edge_id), child_edge, PV(PLVType::P, child_edge));
    }
    if (choices.left_child != NoId && choices.right_child != NoId)
      ops.push_back(Multiply{PV(PLVType::P, edge_id), PV(PLVType::PHatLeft, edge_id), PV(PLVType::PHatRight, edge_id)});
    // an edge with one child choice only (TakePVValue, :833-837) does not occur below a complete
    // subsplit; a leaf edge has none and keeps its site-pattern P-PV
    Assert((choices.left_child == NoId) == (choices.right_child == NoId),
           "TPLikelihoodPlan: edge with a single child choice.");
  }
  void LeafwardForEdge(GPOperationVector& ops, const EdgeId edge_id) const {  // :850-866
    using namespace GPOperations;
    const auto& choices = choice_map_.GetEdgeChoice(edge_id);
    if (choices.parent != NoId) {  // EvolveRPVDownEdge, :1033-1040
      const auto focal = dag_.GetFocalClade(edge_id);
      Evolve(ops, PV(PLVType::RHat, edge_id), edge_id, PV(PLVTypeEnum::RPLVType(focal), choices.parent));
    }
    ops.push_back(Multiply{PV(PLVType::RLeft, edge_id), PV(PLVType::RHat, edge_id), PV(PLVType::PHatRight, edge_id)});
    ops.push_back(Multiply{PV(PLVType::RRight, edge_id), PV(PLVType::RHat, edge_id), PV(PLVType::PHatLeft, edge_id)});
  }

  const GPDAG& dag_;
  const TPChoiceMap& choice_map_;
  size_t taxon_count_, edge_count_;
};
