// tests/cpp/tp_parity.cpp — the TP likelihood evaluator through GP op lists.
//
// Builds a GPDAG and the reference's TPEngine (tp_engine.cpp, tp_evaluation_engine.cpp, unmodified)
// from a fasta + newick pair, lets the reference compute its top-tree log-likelihoods
// (TPEvalEngineViaLikelihood::Initialize + ComputeScores), then turns the SAME TPChoiceMap into two
// GPOperationVectors with bito_b200/host/tp_likelihood_plan.hpp and runs them
//   (cpu) through the unmodified reference CPU GPEngine — this checks the plan itself, no GPU needed;
//   (gpu) with `--gpu`, through GPEngineB200 (the host class over libbito_gp_b200.so).
// Per-edge top-tree log-likelihoods must agree to 1e-9 relative.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>

#include "alignment.hpp"
#include "driver.hpp"
#include "gp_dag.hpp"
#include "gp_engine.hpp"
#include "gp_engine_b200.hpp"
#include "rooted_tree_collection.hpp"
#include "site_pattern.hpp"
#include "tp_engine.hpp"
#include "tp_likelihood_plan.hpp"

namespace {
int g_failures = 0;
double RelErr(const EigenVectorXd& a, const EigenVectorXd& b) {
  if (a.size() != b.size()) return 1e300;
  double worst = 0.;
  for (Eigen::Index i = 0; i < a.size(); ++i)
    worst = std::max(worst, std::abs(a[i] - b[i]) / std::max(1.0, std::abs(b[i])));
  return worst;
}
void Report(const char* what, double err, double tol) {
  const bool ok = err <= tol;
  std::printf("%s %-58s err %.3e (tol %.1e)\n", ok ? "ok  " : "FAIL", what, err, tol);
  if (!ok) ++g_failures;
}
template <typename Engine>
EigenVectorXd RunPlan(Engine& engine, const TPLikelihoodPlan& plan, const EigenVectorXd& branch_lengths) {
  engine.SetNullPrior();
  engine.SetBranchLengths(branch_lengths);
  engine.ProcessOperations(plan.InitializeOps());
  engine.ProcessOperations(plan.ComputeScoresOps());
  return engine.GetPerGPCSPLogLikelihoods();
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s fasta newick [--gpu]\n", argv[0]);
    return 2;
  }
  const bool with_gpu = argc > 3 && std::strcmp(argv[3], "--gpu") == 0;
  try {
    Alignment alignment = Alignment::ReadFasta(argv[1]);
    Driver driver;
    driver.SetSortTaxa(false);
    RootedTreeCollection trees =
        RootedTreeCollection::OfTreeCollection(driver.ParseNewickFile(argv[2]));
    GPDAG dag(trees);
    const size_t E = dag.EdgeCountWithLeafSubsplits();
    const std::string tag = std::string("/tmp/gp_tp_parity_") + std::to_string(getpid());
    SitePattern site_pattern(alignment, trees.TagTaxonMap());

    // the reference TP engine (gp_instance.cpp:791-800, gp_doctest.cpp:2688-2713)
    const auto edge_indexer = dag.BuildEdgeIndexer();
    TPEngine tp(dag, site_pattern, tag + ".tp_lik", tag + ".tp_pars", trees, edge_indexer);
    EigenVectorXd branch_lengths(E);
    for (size_t e = 0; e < E; ++e) branch_lengths[e] = 0.02 + 0.013 * double(e % 11);
    EigenVectorXd padded = tp.GetBranchLengths();
    padded.head(E) = branch_lengths;
    tp.SetBranchLengths(padded);
    tp.SetChoiceMapByTakingFirst(trees, edge_indexer);
    tp.GetLikelihoodEvalEngine().Initialize();
    tp.GetLikelihoodEvalEngine().ComputeScores();
    const EigenVectorXd want = tp.GetTopTreeLikelihoods().head(E);
    std::printf("taxa %zu, patterns %zu, nodes %zu, edges %zu; top-tree log-likelihood of edge 0: %.10f\n",
                dag.TaxonCount(), site_pattern.PatternCount(), dag.NodeCountWithoutDAGRoot(), E, want[0]);

    const TPLikelihoodPlan plan(dag, tp.GetChoiceMap());
    const size_t N = plan.EngineNodeCount(), G = plan.EngineGPCSPCount();
    const EigenVectorXd ones_g = EigenVectorXd::Ones(G), ones_n = EigenVectorXd::Ones(N);
    {
      GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n,
                   ones_g, false);
      Report("plan through the reference CPU GPEngine vs TPEngine", RelErr(RunPlan(cpu, plan, branch_lengths), want),
             1e-9);
    }
    if (with_gpu) {
      GPEngineB200 gpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n,
                       ones_g, false);
      const EigenVectorXd got = RunPlan(gpu, plan, branch_lengths);
      Report("plan through GPEngineB200 (CUDA) vs TPEngine", RelErr(got, want), 1e-9);
      // a second evaluation with other branch lengths on the same engine (graphs replay, matrices rebuilt)
      EigenVectorXd other = branch_lengths * 1.7;
      padded.head(E) = other;
      tp.SetBranchLengths(padded);
      tp.GetLikelihoodEvalEngine().Initialize();
      tp.GetLikelihoodEvalEngine().ComputeScores();
      Report("second evaluation, new branch lengths (CUDA) vs TPEngine",
             RelErr(RunPlan(gpu, plan, other), tp.GetTopTreeLikelihoods().head(E)), 1e-9);
    }
    for (const char* suffix : {".tp_lik", ".tp_pars", ".gp"}) unlink((tag + suffix).c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "tp_parity: %s\n", e.what());
    return 1;
  }
  std::printf("%s\n", g_failures == 0 ? "TP PARITY PASS" : "TP PARITY FAIL");
  return g_failures == 0 ? 0 : 1;
}
