// tests/cpp/nni_parity.cpp — the reference's NNI search on top of a GPEngine, printed as text.
//
// ONE source, compiled twice by oracle/Makefile (`make nniparity`):
//   oracle/_ref/nni_parity_ref   against the unmodified reference objects (CPU GPEngine);
//   oracle/_ref/nni_parity_b200  with bito_b200/host/gp_engine_b200.hpp installed as
//       gp_engine.hpp (class name GPEngine) and the reference's OWN nni_engine.cpp /
//       nni_evaluation_engine.cpp / tp_engine.cpp recompiled against it - i.e. exactly the swap
//       INTEGRATION.md describes - forwarding to libbito_gp_b200.so.
// Both run NNIEngine with the GP evaluation engine (nni_evaluation_engine.cpp:51-843: GrowPLVs /
// GrowGPCSPs with reindexers, spare PLVs and edges, CopyPLVData / CopyGPCSPData, graft-DAG scoring
// with hand-built op lists, branch lengths written through the DAGBranchHandler reference) the way
// test/nni_search.py --gp does (nni_search.py:601-622) and print every scored / accepted NNI per
// iteration; tests/test_nni_parity_gpu.py compares the two outputs.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

#include "alignment.hpp"
#include "driver.hpp"
#include "gp_dag.hpp"
#include "gp_engine.hpp"
#include "nni_engine.hpp"
#include "rooted_tree_collection.hpp"
#include "site_pattern.hpp"

namespace {
void PrintVector(const char* name, const EigenVectorXd& v) {
  std::printf("%s %zu", name, size_t(v.size()));
  for (Eigen::Index i = 0; i < v.size(); ++i) std::printf(" %.17g", v[i]);
  std::printf("\n");
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s fasta newick [iterations] [optimize_new_edges] [sweeps]\n", argv[0]);
    return 2;
  }
  const int iterations = argc > 3 ? std::atoi(argv[3]) : 3;
  const bool optimize_new_edges = argc > 4 ? std::atoi(argv[4]) != 0 : true;
  const int sweeps = argc > 5 ? std::atoi(argv[5]) : 2;
  try {
    Alignment alignment = Alignment::ReadFasta(argv[1]);
    Driver driver;
    driver.SetSortTaxa(false);
    RootedTreeCollection trees =
        RootedTreeCollection::OfTreeCollection(driver.ParseNewickFile(argv[2]));
    GPDAG dag(trees);
    // GPInstance::MakeGPEngine, gp_instance.cpp:146-164
    EigenVectorXd sbn_prior = dag.BuildUniformOnTopologicalSupportPrior();
    EigenVectorXd unconditional = dag.UnconditionalNodeProbabilities(sbn_prior);
    EigenVectorXd inverted = dag.InvertedGPCSPProbabilities(sbn_prior, unconditional);
    const size_t N = dag.NodeCountWithoutDAGRoot(), E = dag.EdgeCountWithLeafSubsplits();
    const std::string mmap_path = std::string("/tmp/gp_nni_parity_") + std::to_string(getpid()) + ".plv";
    GPEngine engine(SitePattern(alignment, trees.TagTaxonMap()), N, E, mmap_path,
                    GPEngine::default_rescaling_threshold_, sbn_prior, unconditional.segment(0, N),
                    inverted, false);
    std::printf("dag nodes %zu edges %zu patterns %zu\n", N, E, engine.GetSitePatternCount());

    // GPInstance::EstimateBranchLengths' loop, gp_instance.cpp:241-308, for a fixed number of sweeps
    engine.SetBranchLengthsToDefault();
    engine.ResetOptimizationCount();
    engine.ProcessOperations(dag.PopulatePLVs());
    for (int s = 0; s < sweeps; ++s) {
      engine.ProcessOperations(dag.BranchLengthOptimization());
      engine.IncrementOptimizationCount();
      engine.ProcessOperations(dag.PopulatePLVs());
    }
    engine.ProcessOperations(dag.ComputeLikelihoods());
    PrintVector("initial_branch_lengths", engine.GetBranchLengths());
    PrintVector("initial_per_gpcsp_llh", engine.GetPerGPCSPLogLikelihoods());

    // nni_search.py:601-622 (init_engine_for_gp_search)
    NNIEngine nni_engine(dag, nullptr, nullptr);
    nni_engine.MakeGPEvalEngine(&engine);
    nni_engine.GetGPEvalEngine().SetOptimizeNewEdges(optimize_new_edges);
    nni_engine.SetGPLikelihoodCutoffFilteringScheme(0.0);
    nni_engine.SetTopKScoreFilteringScheme(1);
    nni_engine.RunInit(true);
    for (int it = 0; it < iterations && nni_engine.GetAdjacentNNICount() > 0; ++it) {
      std::printf("iteration %d adjacent %zu\n", it, nni_engine.GetAdjacentNNICount());
      nni_engine.RunMainLoop(true);
      for (const auto& [nni, score] : nni_engine.GetScoredNNIs())  // std::map: NNIOperation order
        std::printf("scored %s %.17g\n", nni.ToHashString(16).c_str(), score);
      for (const auto& nni : nni_engine.GetAcceptedNNIs())
        std::printf("accepted %s\n", nni.ToHashString(16).c_str());
      nni_engine.RunPostLoop(true);
      std::printf("dag_after nodes %zu edges %zu engine_nodes %zu engine_gpcsps %zu\n",
                  dag.NodeCountWithoutDAGRoot(), dag.EdgeCountWithLeafSubsplits(),
                  engine.GetNodeCount(), engine.GetGPCSPCount());
      PrintVector("branch_lengths", engine.GetBranchLengths());
    }
    // the grown DAG through the grown engine: a fresh full pass
    engine.ProcessOperations(dag.PopulatePLVs());
    engine.ProcessOperations(dag.ComputeLikelihoods());
    PrintVector("final_per_gpcsp_llh", engine.GetPerGPCSPLogLikelihoods());
    std::printf("final_log_marginal %.17g\n", engine.GetLogMarginalLikelihood());
    unlink(mmap_path.c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "nni_parity: %s\n", e.what());
    return 1;
  }
  return 0;
}
