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
#pragma once

#include <vector>

#include "gp_dag.hpp"
#include "gp_operation.hpp"
#include "pv_handler.hpp"
#include "tp_choice_map.hpp"

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
