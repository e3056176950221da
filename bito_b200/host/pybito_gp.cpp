// bito_b200/host/pybito_gp.cpp — the GP part of the `bito` Python module (pybito) over the B200 engine.
//
// The reference's pybito.cpp binds the BEAGLE-backed SBN instances next to the GP classes, so it cannot be
// built without libhmsbeagle. This TU is the GP-only surface SURVEY.md 8c lists, with the Python names and
// argument defaults of the reference so that a script written against `bito` runs unchanged:
//   gp_instance  /root/reference/src/pybito.cpp:614-776      dag       :780-837    graft_dag :839-864
//   gp_engine    :866-870        nni_engine :926-1063 (search loop, GP filters)    tp_engine :872-924 (counts)
//   RootedTree / RootedTreeCollection :114-192     node_topology :1086-1120     bitset, subsplit(), pcsp() :1122-1163
//   node_id / edge_id / taxon_id / tree_id :1165-1190        nni_op :1192-1215
// It is compiled against the reference's own headers with gp_engine.hpp swapped for the host class
// (bito_b200/host/gp_engine_b200.hpp, -DBITO_B200_ENGINE_CLASS=GPEngine), i.e. gp_instance.cpp, gp_dag.cpp,
// nni_engine.cpp ... are the reference's files, unchanged, and every GPEngine call lands in libbito_gp_b200.so
// through the C-ABI. Built without that swap the same TU binds the reference CPU GPEngine: that second build
// is what tests/test_pybito_gpu.py checks this one against (both recipes: `make pybito`).
// The two BEAGLE-backed members (get_likelihood_tree_engine, compute_tree_likelihood) raise RuntimeError.
#include <pybind11/eigen.h>
#include <pybind11/functional.h>
#include <pybind11/iostream.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <optional>
#include <string>

#include "gp_instance.hpp"

namespace py = pybind11;

namespace {

SubsplitClade CladeOfIndex(size_t i) {
  if (i >= 2) Failwith("child_count must be 0 (left) or 1 (right).");
  return i == 0 ? SubsplitClade::Left : SubsplitClade::Right;
}

[[noreturn]] void NoBeagle(const char *what) {
  Failwith(std::string(what) + ": BEAGLE is not available in the GP-only bito build.");
}

template <typename Id>
void BindId(py::module_ &m, const char *name, const char *doc) {
  py::class_<Id>(m, name, doc)
      .def(py::init<const size_t>())
      .def("__str__", [](const Id self) { return self.ToString(); })
      .def("__eq__", [](const Id lhs, const Id rhs) { return lhs == rhs; })
      .def("__hash__", [](const Id self) { return std::hash<size_t>()(self.value_); })
      .def("value", [](const Id &self) -> int { return static_cast<int>(self.value_); });
}

void BindTrees(py::module_ &m) {
  py::class_<RootedTree>(m, "RootedTree", "A rooted tree with branch lengths.")
      .def("__eq__", [](const RootedTree &a, const RootedTree &b) { return a == b; })
      .def("compare_by_topology",
           [](const RootedTree &a, const RootedTree &b) { return a.Topology() == b.Topology(); })
      .def("to_newick", [](const RootedTree &t) { return t.Newick(); }, "Output to Newick string with branch lengths.")
      .def("to_newick_topology", [](const RootedTree &t) { return t.NewickTopology(std::nullopt); },
           "Output to Newick string without branch lengths.")
      .def("parent_id_vector", &RootedTree::ParentIdVector)
      .def_static("example", &RootedTree::Example)
      .def_static("of_parent_id_vector", &RootedTree::OfParentIdVector)
      .def_readwrite("branch_lengths", &RootedTree::branch_lengths_)
      .def_readwrite("height_ratios", &RootedTree::height_ratios_)
      .def_readwrite("node_heights", &RootedTree::node_heights_)
      .def_readwrite("node_bounds", &RootedTree::node_bounds_)
      .def_readwrite("rates", &RootedTree::rates_)
      .def("topology", [](const RootedTree &t) { return t.Topology(); })
      .def("id", [](const RootedTree &t) { return t.Topology()->Id(); }, "Unique node id within topology.")
      .def("to_leaves", [](const RootedTree &t) { return t.Topology()->Leaves(); }, "Output node to leave bitset.")
      .def("build_subsplit", [](const RootedTree &t) { return t.Topology()->BuildSubsplit(); },
           "Build subsplit node bitset of node.")
      .def("build_pcsp",
           [](const RootedTree &t, const size_t child_id) { return t.Topology()->BuildPCSP(CladeOfIndex(child_id)); },
           "Build PCSP edge bitset of edge below node.")
      .def("build_set_of_subsplits", [](const RootedTree &t) { return t.Topology()->BuildSetOfSubsplits(); },
           "Build set of all subsplit bitsets for all nodes in topology.")
      .def("build_set_of_pcsps", [](const RootedTree &t) { return t.Topology()->BuildSetOfPCSPs(); },
           "Build set of all PCSP edge bitsets for all edges in topology.");

  py::class_<RootedTreeCollection>(m, "RootedTreeCollection", "A collection of rooted trees (member `trees`).")
      .def(py::init<RootedTree::RootedTreeVector>(), "The empty constructor.")
      .def(py::init<RootedTree::RootedTreeVector, TagStringMap>(),
           "Constructor from a vector of trees and a tags->taxon names map.")
      .def(py::init<RootedTree::RootedTreeVector, const std::vector<std::string> &>(),
           "Constructor from a vector of trees and a vector of taxon names.")
      .def("erase", &RootedTreeCollection::Erase, "Erase the specified range from the current tree collection.")
      .def("drop_first", &RootedTreeCollection::DropFirst, "Drop the first ``fraction`` trees from the tree collection.",
           py::arg("fraction"))
      .def("newick", &RootedTreeCollection::Newick, "Get the current set of trees as a big Newick string.")
      .def("tree_count", &RootedTreeCollection::TreeCount)
      .def_readwrite("trees", &RootedTreeCollection::trees_);

  py::class_<Node::Topology>(m, "node_topology", "A node in a node topology representing a tree.")
      .def("__str__", [](const Node::Topology &n) { return n->Leaves().ToString(); })
      .def("id", [](const Node::Topology &n) { return n->Id(); }, "Unique node id within topology.")
      .def("to_leaves", [](const Node::Topology &n) { return n->Leaves(); }, "Output node to leave bitset.")
      .def("build_subsplit", [](const Node::Topology &n) { return n->BuildSubsplit(); },
           "Build subsplit node bitset of node.")
      .def("build_pcsp", [](const Node::Topology &n, const size_t child_id) { return n->BuildPCSP(CladeOfIndex(child_id)); },
           "Build PCSP edge bitset of edge below node.")
      .def("build_set_of_subsplits", [](const Node::Topology &n) { return n->BuildSetOfSubsplits(); },
           "Build a vector of all subsplit bitsets for all nodes in topology.")
      .def("build_set_of_pcsps", [](const Node::Topology &n) { return n->BuildSetOfPCSPs(); },
           "Build vector of all PCSP edge bitsets for all edges in topology.")
      .def("to_newick", [](const Node::Topology &n) { return n->Newick(); }, "Output to Newick string.");
}

void BindBitsets(py::module_ &m) {
  py::class_<Bitset>(m, "bitset", "A bitset representing the taxon membership of a Subsplit or PCSP.")
      .def(py::init<const std::string &>())
      .def("__str__", &Bitset::ToString)
      .def("__repr__", [](const Bitset &b) { return b.ToHashString(); })
      .def("__eq__", [](const Bitset &a, const Bitset &b) { return a == b; })
      .def("__hash__", &Bitset::Hash)
      .def("to_string", &Bitset::ToString)
      .def("to_hash_string", &Bitset::ToHashString, py::arg("length") = 16)
      .def("subsplit_to_hash_string", &Bitset::SubsplitToHashString, py::arg("length") = 16)
      .def("pcsp_to_hash_string", &Bitset::PCSPToHashString, py::arg("length") = 16)
      .def("clade_get_count", &Bitset::Count)
      .def("subsplit_get_clade", [](const Bitset &b, const size_t i) { return b.SubsplitGetClade(CladeOfIndex(i)); })
      .def("subsplit_is_uca", &Bitset::SubsplitIsUCA)
      .def("subsplit_is_rootsplit", &Bitset::SubsplitIsRootsplit)
      .def("subsplit_is_leaf", &Bitset::SubsplitIsLeaf)
      .def("subsplit_to_string", &Bitset::SubsplitToString, "Output as Subsplit-style string.")
      .def("pcsp_to_string", &Bitset::PCSPToString, "Output as PCSP-style string.")
      .def("pcsp_get_parent_subsplit", &Bitset::PCSPGetParentSubsplit, "Get parent subsplit from PCSP.")
      .def("pcsp_get_child_subsplit", &Bitset::PCSPGetChildSubsplit, "Get child subsplit from PCSP.");
  m.def("subsplit", [](const std::string &left, const std::string &right) { return Bitset::Subsplit(left, right); },
        "A Subsplit Bitset constructed from two Bitset Clades.");
  m.def("pcsp", [](const Bitset &parent, const Bitset &child) { return Bitset::PCSP(parent, child); },
        "A PCSP Bitset constructed from two Bitset Subsplits.");

  BindId<NodeId>(m, "node_id", "An ID representing a unique node within a DAG.");
  BindId<EdgeId>(m, "edge_id", "An ID representing a unique edge within a DAG.");
  BindId<TaxonId>(m, "taxon_id", "An ID representing a unique taxon within a DAG.");
  BindId<TreeId>(m, "tree_id", "An ID representing a unique tree.");

  py::class_<NNIOperation>(m, "nni_op", "A proposed NNI Operation for the DAG. Repesents the PCSP to be added.")
      .def(py::init<const std::string &, const std::string &>())
      .def("__str__", &NNIOperation::ToString)
      .def("__repr__", [](const NNIOperation &n) { return n.ToHashString(); })
      .def("__eq__", [](const NNIOperation &a, const NNIOperation &b) { return a == b; })
      .def("__hash__", &NNIOperation::Hash)
      .def("to_hash_string", &NNIOperation::ToHashString, py::arg("length") = 16)
      .def("to_string", &NNIOperation::ToString)
      .def("get_parent", &NNIOperation::GetParent, "Get parent Subsplit of PCSP.")
      .def("get_child", &NNIOperation::GetChild, "Get child Subsplit of PCSP.")
      .def("get_central_edge_pcsp", &NNIOperation::GetCentralEdgePCSP, "Get central edge PCSP.")
      .def("is_valid", &NNIOperation::IsValid, "Checks that NNI Operation is a valid PCSP.");
}

void BindDags(py::module_ &m) {
  py::class_<GPDAG>(m, "dag", "Subsplit DAG for performing GPOperations.")
      .def("__eq__", [](const GPDAG &a, const GPDAG &b) { return a == b; })
      .def("node_count", &GPDAG::NodeCount, "Get number of nodes contained in DAG.")
      .def("edge_count", &GPDAG::EdgeCountWithLeafSubsplits, "Get number of edges contained in DAG.")
      .def("taxon_count", &GPDAG::TaxonCount, "Get number of taxa in DAG.")
      .def("topology_count", &GPDAG::TopologyCount, "Get number of unique topologies contained in DAG.")
      .def("get_nni", &GPDAG::GetNNI, "Get NNI for the given DAG edge.")
      .def("get_node_id", [](const GPDAG &d, const Bitset &b) { return d.GetDAGNodeId(b); })
      .def("get_edge_id", [](const GPDAG &d, const Bitset &b) { return d.GetEdgeIdx(b); })
      .def("get_edge_id", [](const GPDAG &d, const NNIOperation &n) { return d.GetEdgeIdx(n); })
      .def("get_taxon_map", &GPDAG::GetTaxonMap, "Get map of taxon names contained in DAG.")
      .def("build_set_of_node_bitsets", &GPDAG::BuildSetOfNodeBitsets,
           "Build a set of node Subsplit bitsets contained in DAG.")
      .def("build_set_of_edge_bitsets", &GPDAG::BuildSetOfEdgeBitsets,
           "Build a set of edge PCSP bitsets contained in DAG.")
      .def("contains_node", [](const GPDAG &d, const Bitset &b) { return d.ContainsNode(b); })
      .def("contains_edge", [](const GPDAG &d, const Bitset &b) { return d.ContainsEdge(b); })
      .def("contains_nni", &GPDAG::ContainsNNI)
      .def("contains_tree", &GPDAG::ContainsTree, "Check whether DAG contains tree.", py::arg("tree"),
           py::arg("is_quiet") = true)
      .def("contains_topology", &GPDAG::ContainsTopology, "Check whether DAG contains topology.")
      .def("is_valid_add_node_pair", &GPDAG::IsValidAddNodePair,
           "Checks whether a given parent/child subsplit pair is valid to be added to the DAG.")
      .def("add_node_pair", [](GPDAG &d, const Bitset &parent, const Bitset &child) { d.AddNodePair(parent, child); },
           "Add parent/child subsplit pair to DAG.")
      .def("add_nodes", &GPDAG::AddNodes)
      .def("add_edges", &GPDAG::AddEdges)
      .def("fully_connect", [](GPDAG &d) { d.FullyConnect(); }, "Adds all valid edges with present nodes to the DAG.")
      .def("tree_to_newick_topology", &GPDAG::TreeToNewickTopology)
      .def("tree_to_newick_tree", &GPDAG::TreeToNewickTree)
      .def("topology_to_newick_topology", &GPDAG::TopologyToNewickTopology)
      .def("generate_all_topologies", &GPDAG::GenerateAllTopologies)
      .def("to_newick_of_all_topologies", &GPDAG::ToNewickOfAllTopologies)
      .def("generate_covering_topologies", &GPDAG::GenerateCoveringTopologies)
      .def("to_newick_of_covering_topologies", &GPDAG::ToNewickOfCoveringTopologies);

  py::class_<GraftDAG>(m, "graft_dag", "Subsplit DAG for grafting nodes and edges.")
      .def("compare_to_dag", [](const GraftDAG &g, const GPDAG &d) { return g.CompareToDAG(d); })
      .def("graft_node_count", &GraftDAG::GraftNodeCount, "Get number of graft nodes appended to DAG.")
      .def("graft_edge_count", &GraftDAG::GraftEdgeCount, "Get number of graft edges appended to DAG.")
      .def("host_node_count", &GraftDAG::HostNodeCount, "Get number of host nodes contained in DAG.")
      .def("host_edge_count", &GraftDAG::HostEdgeCount, "Get number of host edges contained in DAG.")
      .def("is_valid_add_node_pair", &GraftDAG::IsValidAddNodePair)
      .def("add_node_pair",
           [](GraftDAG &g, const Bitset &parent, const Bitset &child) { g.AddNodePair(parent, child); });
}

void BindEngines(py::module_ &m) {
  py::class_<GPEngine>(m, "gp_engine", "An engine for computing Generalized Pruning.")
      .def("node_count", &GPEngine::GetNodeCount, "Get number of nodes.")
      .def("plv_count", &GPEngine::GetPLVCount, "Get number of PLVs.")
      .def("edge_count", &GPEngine::GetGPCSPCount, "Get number of edges.")
      // read-only views a user of the engine object asks for most (members of the reference class too,
      // gp_engine.hpp:113-143)
      .def("get_branch_lengths", [](const GPEngine &e) { return EigenVectorXd(e.GetBranchLengths()); })
      .def("get_per_gpcsp_log_likelihoods", [](GPEngine &e) { return EigenVectorXd(e.GetPerGPCSPLogLikelihoods()); })
      .def("get_log_marginal_likelihood", [](GPEngine &e) { return e.GetLogMarginalLikelihood(); })
      .def("get_sbn_parameters", [](GPEngine &e) { return EigenVectorXd(e.GetSBNParameters()); });

  py::class_<TPEngine>(m, "tp_engine", "An engine for computing Top Pruning.")
      .def("node_count", &TPEngine::GetNodeCount, "Get number of nodes.")
      .def("edge_count", &TPEngine::GetEdgeCount, "Get number of edges.")
      .def("get_top_tree_score", &TPEngine::GetTopTreeScore)
      .def("get_branch_lengths", [](TPEngine &e) { return e.GetBranchLengths(); })
      .def("optimize_branch_lengths", &TPEngine::OptimizeBranchLengths,
           py::arg("check_branch_convergence") = std::nullopt);

  py::class_<NNIEngine>(m, "nni_engine", "An engine for computing NNI Systematic Search.")
      .def("get_branch_lengths", &NNIEngine::GetBranchLengths, "Get DAG branch lengths.")
      .def("adjacent_nnis", &NNIEngine::GetAdjacentNNIs, "Get NNIs adjacent to DAG.")
      .def("new_adjacent_nnis", &NNIEngine::GetNewAdjacentNNIs, "Get new NNIs adjacent to DAG.")
      .def("accepted_nnis", &NNIEngine::GetAcceptedNNIs, "Get NNIs accepted into DAG.")
      .def("rejected_nnis", &NNIEngine::GetRejectedNNIs, "Get NNIs rejected from DAG.")
      .def("scored_nnis", &NNIEngine::GetScoredNNIs, "Get Scored NNIs of current iteration.")
      .def("past_scored_nnis", &NNIEngine::GetPastScoredNNIs, "Get scores from NNIs from previous iterations.")
      .def("adjacent_nni_count", &NNIEngine::GetAdjacentNNICount, "Get number of NNIs adjacent to DAG.")
      .def("new_adjacent_nni_count", &NNIEngine::GetNewAdjacentNNICount)
      .def("accepted_nni_count", &NNIEngine::GetAcceptedNNICount)
      .def("rejected_nni_count", &NNIEngine::GetRejectedNNICount)
      .def("past_accepted_nni_count", &NNIEngine::GetPastAcceptedNNICount)
      .def("past_rejected_nni_count", &NNIEngine::GetPastRejectedNNICount)
      .def("scored_nni_count", &NNIEngine::GetScoredNNICount, "Get number of current NNI scores.")
      .def("iter_count", &NNIEngine::GetIterationCount, "Get number of iterations of NNI search run.")
      .def("run", &NNIEngine::Run, "Primary runner for NNI systematic search.", py::arg("is_quiet") = true)
      .def("run_init", &NNIEngine::RunInit, "Run initialization step of NNI search.", py::arg("is_quiet") = true)
      .def("run_main_loop", &NNIEngine::RunMainLoop, "Run main loop of NNI search.", py::arg("is_quiet") = true)
      .def("run_post_loop", &NNIEngine::RunPostLoop, "Run post loop of NNI search.", py::arg("is_quiet") = true)
      .def("set_no_filter", &NNIEngine::SetNoFilter, py::arg("set_all_nni_to_accept"))
      .def("set_gp_likelihood_cutoff_filtering_scheme", &NNIEngine::SetGPLikelihoodCutoffFilteringScheme)
      .def("set_gp_likelihood_drop_filtering_scheme", &NNIEngine::SetGPLikelihoodDropFilteringScheme)
      .def("set_tp_likelihood_cutoff_filtering_scheme", &NNIEngine::SetTPLikelihoodCutoffFilteringScheme)
      .def("set_tp_likelihood_drop_filtering_scheme", &NNIEngine::SetTPLikelihoodDropFilteringScheme)
      .def("set_top_k_score_filtering_scheme", &NNIEngine::SetTopKScoreFilteringScheme, py::arg("top_k"),
           py::arg("max_is_best") = true)
      .def("set_include_rootsplits", &NNIEngine::SetIncludeRootsplitNNIs)
      .def("set_reevaluate_rejected_nnis", &NNIEngine::SetReevaluateRejectedNNIs)
      .def("set_rescore_rejected_nnis", &NNIEngine::SetRescoreRejectedNNIs)
      .def("get_score_by_nni", &NNIEngine::GetScoreByNNI, "Get score by NNI.")
      .def("get_score_by_edge", &NNIEngine::GetScoreByEdge, "Get score by EdgeId.");

  py::class_<SankoffHandler>(m, "parsimony_tree_engine", "An engine that computes parsimonies for tree topologies.")
      .def("compute_parsimony", [](SankoffHandler &s, const RootedTree &tree) {
        s.RunSankoff(tree.Topology());
        return s.ParsimonyScore();
      });
}

void BindInstance(py::module_ &m) {
  py::enum_<OptimizationMethod>(m, "optimization_method")
      .value("BrentOptimization", OptimizationMethod::BrentOptimization)
      .value("BrentOptimizationWithGradients", OptimizationMethod::BrentOptimizationWithGradients)
      .value("GradientAscentOptimization", OptimizationMethod::GradientAscentOptimization)
      .value("LogSpaceGradientAscentOptimization", OptimizationMethod::LogSpaceGradientAscentOptimization)
      .value("NewtonOptimization", OptimizationMethod::NewtonOptimization);

  py::class_<GPInstance>(m, "gp_instance", "A generalized pruning instance.")
      .def(py::init<const std::string &>())
      .def("print_status", &GPInstance::PrintStatus, "Print information about the instance.")
      .def("dag_summary_statistics", &GPInstance::DAGSummaryStatistics, "Return summary statistics about the DAG.")
      .def("make_dag", &GPInstance::MakeDAG, "Build subsplit DAG.")
      .def("print_dag", &GPInstance::PrintDAG, "Print the subsplit DAG.")
      // I/O
      .def("read_newick_file", &GPInstance::ReadNewickFile, py::arg("path"), py::arg("sort_taxa") = true,
           "Read trees from a Newick file.")
      .def("read_newick_file_gz", &GPInstance::ReadNewickFileGZ, py::arg("path"), py::arg("sort_taxa") = true)
      .def("read_nexus_file", &GPInstance::ReadNexusFile, py::arg("path"), py::arg("sort_taxa") = true)
      .def("read_nexus_file_gz", &GPInstance::ReadNexusFileGZ, py::arg("path"), py::arg("sort_taxa") = true)
      .def("read_fasta_file", &GPInstance::ReadFastaFile, "Read a sequence alignment from a FASTA file.")
      .def("sbn_parameters_to_csv", &GPInstance::SBNParametersToCSV)
      .def("sbn_prior_to_csv", &GPInstance::SBNPriorToCSV)
      .def("branch_lengths_to_csv", &GPInstance::BranchLengthsToCSV)
      .def("per_gpcsp_llhs_to_csv", &GPInstance::PerGPCSPLogLikelihoodsToCSV)
      .def("intermediate_bls_to_csv", &GPInstance::IntermediateBranchLengthsToCSV)
      .def("intermediate_per_gpcsp_llhs_to_csv", &GPInstance::IntermediatePerGPCSPLogLikelihoodsToCSV)
      .def("per_gpcsp_llh_surfaces_to_csv", &GPInstance::PerGPCSPLogLikelihoodSurfacesToCSV)
      .def("tracked_optim_values_to_csv", &GPInstance::TrackedOptimizationValuesToCSV)
      .def("export_trees", &GPInstance::ExportTrees, py::arg("out_path"))
      .def("currently_loaded_trees_with_gp_branch_lengths", &GPInstance::CurrentlyLoadedTreesWithGPBranchLengths,
           "Collection of all rooted trees loaded into DAG.")
      .def("generate_complete_rooted_tree_collection", &GPInstance::GenerateCompleteRootedTreeCollection,
           "Generate collection of all rooted trees expressed in DAG.")
      .def("export_all_generated_topologies", &GPInstance::ExportAllGeneratedTopologies, py::arg("out_path"))
      .def("export_all_generated_trees", &GPInstance::ExportAllGeneratedTrees, py::arg("out_path"))
      .def("export_trees_with_a_pcsp", &GPInstance::ExportTreesWithAPCSP, py::arg("pcsp_string"),
           py::arg("newick_path"))
      .def("subsplit_dag_to_dot", &GPInstance::SubsplitDAGToDot)
      .def("get_branch_lengths", &GPInstance::GetBranchLengths, "Return branch lengths from the GPInstance.")
      .def("build_edge_idx_to_pcsp_map", [](GPInstance &g) { return g.GetDAG().BuildInverseEdgeIndexer(); },
           "Build a map from DAG edge index to its corresponding PCSP bitset.")
      // Estimation
      .def("use_gradient_optimization", &GPInstance::UseGradientOptimization, py::arg("use_gradients") = false)
      .def("hot_start_branch_lengths", &GPInstance::HotStartBranchLengths,
           "Use given trees to initialize branch lengths.")
      .def("gather_branch_lengths", &GPInstance::GatherBranchLengths)
      .def("calculate_hybrid_marginals", &GPInstance::CalculateHybridMarginals, "Calculate hybrid marginals.")
      .def("estimate_sbn_parameters", &GPInstance::EstimateSBNParameters,
           "Estimate the SBN parameters based on current branch lengths.")
      .def("hot_start_branch_length", &GPInstance::HotStartBranchLengths)
      .def("take_first_branch_length", &GPInstance::TakeFirstBranchLength)
      .def("estimate_branch_lengths", &GPInstance::EstimateBranchLengths, "Estimate branch lengths for the GPInstance.",
           py::arg("tol"), py::arg("max_iter"), py::arg("quiet") = false,
           py::arg("track_intermediate_iterations") = false, py::arg("optimization_method") = std::nullopt)
      .def("get_perpcsp_llh_surface", &GPInstance::GetPerGPCSPLogLikelihoodSurfaces, py::arg("steps"),
           py::arg("scale_min"), py::arg("scale_max"))
      .def("perturb_and_track_optimization_values", &GPInstance::PerturbAndTrackValuesFromOptimization)
      // GP likelihoods
      .def("populate_plvs", &GPInstance::PopulatePLVs, "Populate PLVs.")
      .def("compute_likelihoods", &GPInstance::ComputeLikelihoods, "Compute Likelihoods.")
      .def("compute_marginal_likelihood", &GPInstance::ComputeMarginalLikelihood)
      .def("get_per_pcsp_log_likelihoods", &GPInstance::GetPerPCSPLogLikelihoods, "Get Per-PCSP Log Likelihoods.")
      .def("get_sbn_parameters", [](GPInstance &g) { return EigenVectorXd(g.GetSBNParameters()); })
      .def("get_log_marginal_likelihood", [](GPInstance &g) { return g.GetGPEngine().GetLogMarginalLikelihood(); })
      // DAG and engines
      .def("get_dag", [](GPInstance &g) -> GPDAG * { return &g.GetDAG(); }, py::return_value_policy::reference,
           "Get Subsplit DAG.")
      .def("make_gp_engine", &GPInstance::MakeGPEngine, "Initialize GP Engine.",
           py::arg("rescaling_threshold") = GPEngine::default_rescaling_threshold_, py::arg("use_gradients") = false)
      .def("get_gp_engine", [](GPInstance &g) -> GPEngine * { return &g.GetGPEngine(); },
           py::return_value_policy::reference, "Get GP Engine.")
      .def("make_nni_engine", &GPInstance::MakeNNIEngine, "Initialize NNI Engine.")
      .def("get_nni_engine", [](GPInstance &g) -> NNIEngine * { return &g.GetNNIEngine(); },
           py::return_value_policy::reference, "Get NNI Engine.")
      .def("make_tp_engine", &GPInstance::MakeTPEngine, "Initialize TP Engine.")
      .def("get_tp_engine", [](GPInstance &g) -> TPEngine * { return &g.GetTPEngine(); },
           py::return_value_policy::reference, "Get TP Engine.")
      .def("tp_engine_set_branch_lengths_by_taking_first", &GPInstance::TPEngineSetBranchLengthsByTakingFirst)
      .def("tp_engine_set_choice_map_by_taking_first", &GPInstance::TPEngineSetChoiceMapByTakingFirst,
           py::arg("use_subsplit_method") = true)
      // tree engines: parsimony works, the BEAGLE-backed likelihood engine is not part of this build
      .def("get_likelihood_tree_engine", [](GPInstance &) { NoBeagle("get_likelihood_tree_engine"); })
      .def("compute_tree_likelihood", [](const GPInstance &, const RootedTree &) -> double {
        NoBeagle("compute_tree_likelihood");
      })
      .def("get_parsimony_tree_engine", &GPInstance::GetParsimonyTreeEngine, py::return_value_policy::reference)
      .def("compute_tree_parsimony", [](const GPInstance &g, const RootedTree &tree) {
        auto site_pattern = g.MakeSitePattern();
        SankoffHandler engine(site_pattern, g.GetMMapFilePath() + ".sankoff");
        engine.RunSankoff(tree.Topology());
        return engine.ParsimonyScore();
      });
}

}  // namespace

#ifndef BITO_PYMODULE_NAME
#define BITO_PYMODULE_NAME bito
#endif

PYBIND11_MODULE(BITO_PYMODULE_NAME, m) {
  m.doc() = "bito (GP-only build): generalized pruning on subsplit DAGs.";
#ifdef BITO_B200_ENGINE_CLASS
  m.attr("gp_engine_backend") = "bito_b200 (CUDA, sm_100a) through the C-ABI of include/bito_gp.h";
#else
  m.attr("gp_engine_backend") = "reference CPU GPEngine";
#endif
  BindTrees(m);
  BindBitsets(m);
  BindDags(m);
  BindEngines(m);
  BindInstance(m);
  py::add_ostream_redirect(m, "ostream_redirect");
}
