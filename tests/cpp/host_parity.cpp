// tests/cpp/host_parity.cpp — TEST PROGRAM (links the reference; never part of the product).
//
// Builds ONE GPDAG with the reference's own parser and planner, then constructs, from the same
// arguments (the clone of GPInstance::MakeGPEngine, /root/reference/src/gp_instance.cpp:146-164),
//   * the UNMODIFIED reference CPU `GPEngine`  (objects of oracle/_ref, compiled from
//     /root/reference/src where they lie), and
//   * `GPEngineB200` (bito_b200/host/gp_engine_b200.{hpp,cpp}), the host class a bito
//     maintainer drops in, which forwards to libbito_gp_b200.so through include/bito_gp.h,
// runs the same GPOperationVectors (std::variant lists straight from GPDAG) through both and
// compares every public read-back. Built here by `make -C oracle hostparity` into
// oracle/_ref/gp_host_parity (needs the reference headers); run on the GPU box by
// tests/test_host_shim_gpu.py with inputs the test writes.
//
//   gp_host_parity <fasta> <rooted-newick> [rescaling_threshold] [sweeps]
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <optional>
#include <queue>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <variant>
#include <vector>

#define private public  // rescaling_counts_ of the reference (test-only; layout unchanged)
#include "gp_engine.hpp"
#undef private
#include "driver.hpp"
#include "gp_dag.hpp"
#include "rooted_tree_collection.hpp"

#include "gp_engine_b200.hpp"

namespace {
int g_checks = 0, g_failures = 0;

void Report(const std::string& what, double err, double tol) {
  ++g_checks;
  const bool ok = err <= tol;  // NaN fails
  if (!ok) ++g_failures;
  std::printf("%-58s max err %.3e (tol %.1e) %s\n", what.c_str(), err, tol, ok ? "ok" : "FAIL");
}

template <typename A, typename B>
double RelErr(const A& got, const B& want) {
  if (got.size() != want.size()) return INFINITY;
  double worst = 0.;
  for (Eigen::Index i = 0; i < want.size(); ++i) {
    const double g = got.data()[i], w = want.data()[i];
    if (std::isinf(w) && g == w) continue;
    const double e = std::fabs(g - w) / std::max(1.0, std::fabs(w));
    if (!(e <= worst)) worst = e;
  }
  return worst;
}
template <typename A, typename B>
double AbsErr(const A& got, const B& want) {
  if (got.size() != want.size()) return INFINITY;
  double worst = 0.;
  for (Eigen::Index i = 0; i < want.size(); ++i) {
    const double e = std::fabs(got.data()[i] - want.data()[i]);
    if (!(e <= worst)) worst = e;
  }
  return worst;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s fasta newick [rescaling_threshold] [sweeps]\n", argv[0]);
    return 2;
  }
  const double threshold = argc > 3 ? std::atof(argv[3]) : GPEngine::default_rescaling_threshold_;
  const int sweeps = argc > 4 ? std::atoi(argv[4]) : 2;
  try {
    Alignment alignment = Alignment::ReadFasta(argv[1]);
    Driver driver;
    driver.SetSortTaxa(false);
    RootedTreeCollection trees =
        RootedTreeCollection::OfTreeCollection(driver.ParseNewickFile(argv[2]));
    GPDAG dag(trees);
    auto make_site_pattern = [&] { return SitePattern(alignment, trees.TagTaxonMap()); };
    EigenVectorXd sbn_prior = dag.BuildUniformOnTopologicalSupportPrior();
    EigenVectorXd unconditional = dag.UnconditionalNodeProbabilities(sbn_prior);
    EigenVectorXd inverted = dag.InvertedGPCSPProbabilities(sbn_prior, unconditional);
    const size_t N = dag.NodeCountWithoutDAGRoot(), E = dag.EdgeCountWithLeafSubsplits();
    const std::string mmap_path = std::string("/tmp/gp_host_parity_") + std::to_string(getpid()) + ".plv";

    GPEngine ref(make_site_pattern(), N, E, mmap_path, threshold, sbn_prior,
                 unconditional.segment(0, N), inverted, false);
    GPEngineB200 gpu(make_site_pattern(), N, E, mmap_path, threshold, sbn_prior,
                     unconditional.segment(0, N), inverted, false);
    std::printf("taxa %zu, patterns %zu, nodes %zu, edges %zu, threshold %g\n", dag.TaxonCount(),
                ref.GetSitePatternCount(), N, E, threshold);

    Report("counts (node, plv, padded plv, gpcsp, padded gpcsp)",
           double((ref.GetNodeCount() != gpu.GetNodeCount()) + (ref.GetPLVCount() != gpu.GetPLVCount()) +
                  (ref.GetPaddedPLVCount() != gpu.GetPaddedPLVCount()) +
                  (ref.GetGPCSPCount() != gpu.GetGPCSPCount()) +
                  (ref.GetPaddedGPCSPCount() != gpu.GetPaddedGPCSPCount()) +
                  (ref.GetSpareGPCSPIndex(1) != gpu.GetSpareGPCSPIndex(1)) +
                  (ref.GetSparePLVIndex(PVId(2)) != gpu.GetSparePLVIndex(PVId(2)))),
           0.);
    ref.SetTransitionMatrixToHaveBranchLength(0.75);
    gpu.SetTransitionMatrixToHaveBranchLength(0.75);
    Report("GetTransitionMatrix(0.75)", AbsErr(gpu.GetTransitionMatrix(), ref.GetTransitionMatrix()), 1e-14);

    // ---- full pass: GPInstance::PopulatePLVs + ComputeLikelihoods (gp_instance.cpp:231-235)
    const GPOperationVector populate = dag.PopulatePLVs(), likelihoods = dag.ComputeLikelihoods();
    for (int rep = 0; rep < 2; ++rep) {
      ref.ProcessOperations(populate);
      ref.ProcessOperations(likelihoods);
      gpu.ProcessOperations(populate);
      gpu.ProcessOperations(likelihoods);
    }
    Report("GetPerGPCSPLogLikelihoods", RelErr(gpu.GetPerGPCSPLogLikelihoods(), ref.GetPerGPCSPLogLikelihoods()), 1e-9);
    Report("GetPerGPCSPLogLikelihoods(1, 2)",
           RelErr(gpu.GetPerGPCSPLogLikelihoods(1, 2), ref.GetPerGPCSPLogLikelihoods(1, 2)), 1e-9);
    Report("GetLogMarginalLikelihood",
           std::fabs(gpu.GetLogMarginalLikelihood() - ref.GetLogMarginalLikelihood()) /
               std::fabs(ref.GetLogMarginalLikelihood()), 1e-9);
    {
      const EigenMatrixXd want = ref.GetLogLikelihoodMatrix();
      Report("GetLogLikelihoodMatrix (per pattern)", RelErr(gpu.GetLogLikelihoodMatrix(), want), 1e-9);
    }
    Report("GetPerGPCSPComponentsOfFullLogMarginal",
           RelErr(gpu.GetPerGPCSPComponentsOfFullLogMarginal(), ref.GetPerGPCSPComponentsOfFullLogMarginal()), 1e-9);
    {
      double worst = 0.;
      for (size_t id = 0; id < ref.GetPLVCount(); id += 1 + ref.GetPLVCount() / 97) {
        const NucleotidePLV want = ref.GetPLV(PVId(id));
        const double scale = std::max(want.cwiseAbs().maxCoeff(), 1e-300);
        worst = std::max(worst, AbsErr(gpu.GetPLV(PVId(id)), want) / scale);
      }
      Report("GetPLV (sampled ids, relative to the PLV's max)", worst, 1e-9);
    }
    {
      const EigenVectorXi got = gpu.GetRescalingCounts();
      double diff = 0.;
      for (size_t i = 0; i < ref.GetPaddedPLVCount(); ++i) diff += got[i] != ref.rescaling_counts_[i];
      Report("rescaling counts (bit-exact)", diff, 0.);
    }
    // ---- derivatives at every OptimizeBranchLength of the reference's own sweep list
    const GPOperationVector sweep = dag.BranchLengthOptimization();
    {
      double e0 = 0., e1 = 0., e2 = 0.;
      int n = 0;
      for (const auto& op : sweep)
        if (const auto* o = std::get_if<GPOperations::OptimizeBranchLength>(&op)) {
          if (n++ % 7) continue;
          const auto [a, b, c] = ref.LogLikelihoodAndFirstTwoDerivatives(*o);
          const auto [x, y, z] = gpu.LogLikelihoodAndFirstTwoDerivatives(*o);
          const auto [p, q] = gpu.LogLikelihoodAndDerivative(*o);
          e0 = std::max({e0, std::fabs(x - a) / std::fabs(a), std::fabs(p - a) / std::fabs(a)});
          e1 = std::max({e1, std::fabs(y - b) / std::max(1., std::fabs(b)), std::fabs(q - b) / std::max(1., std::fabs(b))});
          e2 = std::max(e2, std::fabs(z - c) / std::max(1., std::fabs(c)));
        }
      Report("LogLikelihoodAndFirstTwoDerivatives: ll", e0, 1e-9);
      Report("LogLikelihoodAndFirstTwoDerivatives: dll/dt", e1, 1e-7);
      Report("LogLikelihoodAndFirstTwoDerivatives: d2ll/dt2", e2, 1e-7);
    }
    // ---- quartet hybrid marginals: GPInstance::CalculateHybridMarginals (gp_instance.cpp:408-417)
    if (threshold <= 1e-30) {
      std::vector<QuartetHybridRequest> requests;
      dag.TopologicalEdgeTraversal([&](const NodeId parent, const bool on_left, const NodeId child,
                                       const EdgeId) {
        requests.push_back(dag.QuartetHybridRequestOf(parent, on_left, child));
      });
      double worst = 0.;
      size_t formed = 0;
      for (const auto& r : requests) {
        if (!r.IsFullyFormed()) continue;
        ++formed;
        worst = std::max(worst, RelErr(gpu.CalculateQuartetHybridLikelihoods(r), ref.CalculateQuartetHybridLikelihoods(r)));
        ref.ProcessQuartetHybridRequest(r);
      }
      gpu.ProcessQuartetHybridRequests(requests);
      std::printf("quartet requests %zu, fully formed %zu\n", requests.size(), formed);
      Report("CalculateQuartetHybridLikelihoods", worst, 1e-9);
      const EigenVectorXd want = ref.GetHybridMarginals().segment(0, E);
      Report("GetHybridMarginals after ProcessQuartetHybridRequest", RelErr(gpu.GetHybridMarginals(), want), 1e-9);
    }
    // ---- SBN parameters: GPInstance::EstimateSBNParameters (gp_instance.cpp:401-406)
    {
      const GPOperationVector sbn = dag.OptimizeSBNParameters();
      ref.ProcessOperations(sbn);
      gpu.ProcessOperations(sbn);
      const EigenVectorXd want = ref.GetSBNParameters().segment(0, E);
      Report("GetSBNParameters after OptimizeSBNParameters", AbsErr(gpu.GetSBNParameters(), want), 1e-6);
      ref.InitializePriors(sbn_prior, unconditional.segment(0, N), inverted);
      gpu.InitializePriors(sbn_prior, unconditional.segment(0, N), inverted);
    }
    // ---- branch lengths from the tree sample (gp_engine.cpp:676-746)
    {
      const BitsetSizeMap indexer = dag.BuildEdgeIndexer();
      ref.TakeFirstBranchLength(trees, indexer);
      gpu.TakeFirstBranchLength(trees, indexer);
      Report("TakeFirstBranchLength", AbsErr(gpu.GetBranchLengths(), ref.GetBranchLengths()), 0.);
      ref.HotStartBranchLengths(trees, indexer);
      gpu.HotStartBranchLengths(trees, indexer);
      Report("HotStartBranchLengths", AbsErr(gpu.GetBranchLengths(), ref.GetBranchLengths()), 1e-15);
      const auto a = ref.GatherBranchLengths(trees, indexer), b = gpu.GatherBranchLengths(trees, indexer);
      Report("GatherBranchLengths", a == b ? 0. : 1., 0.);
      ref.SetBranchLengthsToDefault();
      gpu.SetBranchLengthsToDefault();
    }
    // ---- GPInstance::EstimateBranchLengths' loop (gp_instance.cpp:241-308), the reference's own
    // Gauss-Seidel list, one OptimizeBranchLength per dependency level
    {
      const GPOperationVector marginal = dag.MarginalLikelihood();
      for (auto* unused : {&ref}) (void)unused;
      ref.ResetOptimizationCount();
      gpu.ResetOptimizationCount();
      ref.ProcessOperations(populate);
      ref.ProcessOperations(marginal);
      gpu.ProcessOperations(populate);
      gpu.ProcessOperations(marginal);
      for (int s = 0; s < sweeps; ++s) {
        ref.ProcessOperations(sweep);
        ref.ProcessOperations(populate);
        ref.ProcessOperations(marginal);
        gpu.ProcessOperations(sweep);
        gpu.ProcessOperations(populate);
        gpu.ProcessOperations(marginal);
        Report("sweep " + std::to_string(s) + ": GetBranchLengths", AbsErr(gpu.GetBranchLengths(), ref.GetBranchLengths()), 1e-6);
        Report("sweep " + std::to_string(s) + ": GetBranchLengthDifferences",
               AbsErr(gpu.GetBranchLengthDifferences(), ref.GetBranchLengthDifferences()), 1e-6);
        Report("sweep " + std::to_string(s) + ": GetLogMarginalLikelihood",
               std::fabs(gpu.GetLogMarginalLikelihood() - ref.GetLogMarginalLikelihood()) /
                   std::fabs(ref.GetLogMarginalLikelihood()), 1e-7);
        ref.IncrementOptimizationCount();
        gpu.IncrementOptimizationCount();
      }
      Report("GetOptimizationCount", double(ref.GetOptimizationCount() != gpu.GetOptimizationCount()), 0.);
    }
    // ---- single ops through operator() and the spare-PLV surface the NNI engine uses
    {
      ref.GrowSparePLVs(20);
      gpu.GrowSparePLVs(20);
      ref.GrowSpareGPCSPs(5);
      gpu.GrowSpareGPCSPs(5);
      const size_t spare = ref.GetSparePLVIndex(PVId(3)).value_;
      const size_t some_p = dag.TaxonCount();  // P-PLV of the first internal node
      ref.CopyPLVData(some_p, spare);
      gpu.CopyPLVData(some_p, spare);
      const GPOperations::Multiply mult{ref.GetSparePLVIndex(PVId(4)).value_, spare, some_p};
      ref(mult);
      gpu(mult);
      const NucleotidePLV want = ref.GetSparePLV(PVId(4));
      Report("CopyPLVData + operator()(Multiply) into spare PLVs",
             AbsErr(gpu.GetSparePLV(PVId(4)), want) / std::max(want.cwiseAbs().maxCoeff(), 1e-300), 1e-9);
      ref.CopyGPCSPData(EdgeId(1), EdgeId(ref.GetSpareGPCSPIndex(0)));
      gpu.CopyGPCSPData(EdgeId(1), EdgeId(gpu.GetSpareGPCSPIndex(0)));
      Report("CopyGPCSPData -> GetSpareBranchLengths", AbsErr(gpu.GetSpareBranchLengths(0, 1), ref.GetSpareBranchLengths(0, 1)), 0.);
      Report("counts after GrowSpare*", double((ref.GetPaddedPLVCount() != gpu.GetPaddedPLVCount()) +
                                               (ref.GetPaddedGPCSPCount() != gpu.GetPaddedGPCSPCount())), 0.);
    }
    // ---- GrowPLVs / GrowGPCSPs with Reindexers, as the NNI engine calls them after a DAG edit
    // (nni_evaluation_engine.cpp:55-137): two nodes and three edges are added and the new ids are
    // shifted into the middle of the old ranges (Reindexer::ReassignAndShift).
    {
      const size_t N2 = N + 2, E2 = E + 3;
      Reindexer node_reindexer = Reindexer::IdentityReindexer(N2);
      node_reindexer.ReassignAndShift(N2 - 1, dag.TaxonCount() + 1);
      node_reindexer.ReassignAndShift(N2 - 1, N / 2);
      Reindexer edge_reindexer = Reindexer::IdentityReindexer(E2);
      edge_reindexer.ReassignAndShift(E2 - 1, 2);
      edge_reindexer.ReassignAndShift(E2 - 1, E / 2);
      edge_reindexer.ReassignAndShift(E2 - 1, E / 3);
      ref.GrowPLVs(N2, node_reindexer);
      gpu.GrowPLVs(N2, node_reindexer);
      ref.GrowGPCSPs(E2, edge_reindexer);
      gpu.GrowGPCSPs(E2, edge_reindexer);
      Report("counts after GrowPLVs/GrowGPCSPs",
             double((ref.GetNodeCount() != gpu.GetNodeCount()) + (ref.GetPLVCount() != gpu.GetPLVCount()) +
                    (ref.GetPaddedPLVCount() != gpu.GetPaddedPLVCount()) + (ref.GetGPCSPCount() != gpu.GetGPCSPCount()) +
                    (ref.GetPaddedGPCSPCount() != gpu.GetPaddedGPCSPCount())), 0.);
      double worst = 0.;
      for (size_t id = 0; id < ref.GetPLVCount(); ++id) {
        const NucleotidePLV want = ref.GetPLV(PVId(id));
        const double scale = std::max(want.cwiseAbs().maxCoeff(), 1e-300);
        worst = std::max(worst, AbsErr(gpu.GetPLV(PVId(id)), want) / scale);
      }
      Report("GetPLV of every id after the node reindexing", worst, 1e-9);
      {
        const EigenVectorXi got = gpu.GetRescalingCounts();
        double diff = 0.;
        for (size_t i = 0; i < ref.GetPLVCount(); ++i) diff += got[i] != ref.rescaling_counts_[i];
        Report("rescaling counts after the node reindexing", diff, 0.);
      }
      Report("GetBranchLengths after the edge reindexing", AbsErr(gpu.GetBranchLengths(), ref.GetBranchLengths()), 1e-6);
      Report("GetSBNParameters after the edge reindexing",
             AbsErr(gpu.GetSBNParameters(), ref.GetSBNParameters().segment(0, E2)), 1e-12);
      if (threshold <= 1e-30)
        Report("GetHybridMarginals after the edge reindexing",
               RelErr(gpu.GetHybridMarginals(), ref.GetHybridMarginals().segment(0, E2)), 1e-9);
    }
    const bito_gp_stats st = gpu.Stats();
    std::printf("B200 engine: %lld kernel launches, %lld objective evaluations, %lld programs compiled\n",
                (long long)st.kernel_launches, (long long)st.objective_evaluations, (long long)st.programs_compiled);
    std::remove(mmap_path.c_str());
  } catch (const std::exception& e) {
    std::printf("FAIL: exception: %s\n", e.what());
    return 1;
  }
  std::printf("%s: %d checks, %d failures\n", g_failures ? "FAIL" : "PASS", g_checks, g_failures);
  return g_failures ? 1 : 0;
}
