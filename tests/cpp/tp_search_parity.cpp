// tests/cpp/tp_search_parity.cpp — the reference's NNI search in TP mode (test/nni_search.py --tp: NNIEngine +
// NNIEvalEngineViaTP + TPEngine, all unmodified) with the TPEngine's likelihood evaluator swapped for
// TPEvalEngineOverGPEngine (bito_b200/host/tp_eval_engine_b200.hpp):
//   (ref) TPEngine with its own TPEvalEngineViaLikelihood;
//   (cpu) the evaluator subclass over the reference CPU GPEngine: the class itself, no GPU needed - must reproduce
//         the reference search bit for bit (same scored NNIs and scores, same accepted NNIs, same grown DAG, same
//         top-tree likelihoods and branch lengths after every iteration);
//   (gpu) with --gpu, the same class over GPEngineB200, i.e. the CUDA kernels: identical scored / accepted NNIs,
//         scores to 1e-7, branch lengths to 1e-6, top-tree likelihoods to 1e-7.
// Each run builds its own DAG from the same trees (the search grows it).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unistd.h>
#include <vector>

#include "alignment.hpp"
#include "driver.hpp"
#include "gp_dag.hpp"
#include "gp_engine.hpp"
#include "gp_engine_b200.hpp"
#include "nni_engine.hpp"
#include "rooted_tree_collection.hpp"
#include "site_pattern.hpp"
#include "tp_engine.hpp"
#include "tp_eval_engine_b200.hpp"

namespace {
int g_failures = 0;
struct Iteration {
  std::map<std::string, double> scored;
  std::vector<std::string> accepted;
  size_t edges = 0;
  EigenVectorXd top_tree, branch_lengths;
};
using Trace = std::vector<Iteration>;

template <class TP>
Trace RunSearch(TP& tp, GPDAG& dag, const RootedTreeCollection& trees, const BitsetSizeMap& edge_indexer,
                const size_t iterations) {
  const size_t E = dag.EdgeCountWithLeafSubsplits();
  EigenVectorXd padded = tp.GetBranchLengths();
  for (size_t e = 0; e < E; ++e) padded[e] = 0.02 + 0.013 * double(e % 11);
  tp.SetBranchLengths(padded);
  tp.SetChoiceMapByTakingFirst(trees, edge_indexer);
  auto& eval = tp.GetLikelihoodEvalEngine();
  eval.SetOptimizeNewEdges(true);
  eval.Initialize();     // virtual: the swapped evaluator's
  eval.ComputeScores();
  if (getenv("TP_SEARCH_DEBUG")) {
    const EigenVectorXd t = tp.GetTopTreeLikelihoods().head(E);
    std::printf("  after ComputeScores: top-tree llh[0] %.12g min %.12g\n", t[0], t.minCoeff());
  }
  NNIEngine search(dag, nullptr, &tp);
  search.SetTPLikelihoodCutoffFilteringScheme(0.0);
  search.SetTopKScoreFilteringScheme(1);
  search.RunInit(true);
  Trace trace;
  for (size_t it = 0; it < iterations && search.GetAdjacentNNICount() > 0; ++it) {
    Iteration rec;
    search.RunMainLoop(true);
    for (const auto& [nni, score] : search.GetScoredNNIs()) rec.scored[nni.ToHashString(16)] = score;
    for (const auto& nni : search.GetAcceptedNNIs()) rec.accepted.push_back(nni.ToHashString(16));
    search.RunPostLoop(true);
    rec.edges = dag.EdgeCountWithLeafSubsplits();
    rec.top_tree = tp.GetTopTreeLikelihoods().head(rec.edges);
    rec.branch_lengths = eval.GetDAGBranchHandler().GetBranchLengthData().head(rec.edges);
    trace.push_back(rec);
  }
  return trace;
}

void Report(const char* what, bool same_sets, double score_err, double top_err, double bl_err, double score_tol,
            double top_tol, double bl_tol) {
  const bool ok = same_sets && score_err <= score_tol && top_err <= top_tol && bl_err <= bl_tol;
  std::printf("%s %-44s same NNIs %d  scores %.3e (tol %.0e)  top-tree llh %.3e (tol %.0e)  |dBL| %.3e (tol %.0e)\n",
              ok ? "ok  " : "FAIL", what, int(same_sets), score_err, score_tol, top_err, top_tol, bl_err, bl_tol);
  if (!ok) ++g_failures;
}
void Compare(const char* what, const Trace& got, const Trace& want, double score_tol, double top_tol, double bl_tol) {
  bool same = got.size() == want.size();
  double score_err = 0., top_err = 0., bl_err = 0.;
  for (size_t i = 0; same && i < want.size(); ++i) {
    same = got[i].accepted == want[i].accepted && got[i].edges == want[i].edges &&
           got[i].scored.size() == want[i].scored.size();
    if (!same) break;
    for (const auto& [nni, score] : want[i].scored) {
      const auto it = got[i].scored.find(nni);
      if (it == got[i].scored.end()) { same = false; break; }
      score_err = std::max(score_err, std::abs(it->second - score) / std::max(1.0, std::abs(score)));
    }
    for (Eigen::Index e = 0; e < want[i].top_tree.size(); ++e) {
      const double err = std::abs(got[i].top_tree[e] - want[i].top_tree[e]) / std::max(1.0, std::abs(want[i].top_tree[e]));
      if (getenv("TP_SEARCH_DEBUG") && err > 1e-9)
        std::printf("  it %zu edge %ld: got %.10g want %.10g\n", i, long(e), got[i].top_tree[e], want[i].top_tree[e]);
      top_err = std::max(top_err, err);
    }
    bl_err = std::max(bl_err, (got[i].branch_lengths - want[i].branch_lengths).cwiseAbs().maxCoeff());
  }
  if (!same || getenv("TP_SEARCH_DEBUG")) {
    std::printf("  iterations: got %zu want %zu\n", got.size(), want.size());
    for (size_t i = 0; i < std::max(got.size(), want.size()); ++i) {
      for (const Trace* t : {&got, &want}) {
        if (i >= t->size()) continue;
        std::printf("  %s it %zu: edges %zu scored %zu accepted", t == &got ? "got " : "want", i, (*t)[i].edges,
                    (*t)[i].scored.size());
        for (const auto& a : (*t)[i].accepted) std::printf(" %s", a.c_str());
        std::printf("\n");
        if (i == 0)
          for (const auto& [nni, score] : (*t)[i].scored) std::printf("      %s %.12g\n", nni.c_str(), score);
      }
    }
  }
  Report(what, same, score_err, top_err, bl_err, score_tol, top_tol, bl_tol);
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s fasta newick [--gpu]\n", argv[0]);
    return 2;
  }
  const bool with_gpu = argc > 3 && std::strcmp(argv[3], "--gpu") == 0;
  const size_t iterations = 3;
  try {
    Alignment alignment = Alignment::ReadFasta(argv[1]);
    Driver driver;
    driver.SetSortTaxa(false);
    RootedTreeCollection trees = RootedTreeCollection::OfTreeCollection(driver.ParseNewickFile(argv[2]));
    const std::string tag = std::string("/tmp/gp_tp_search_") + std::to_string(getpid());
    Trace ref, cpu, gpu;
    {
      GPDAG dag(trees);
      SitePattern site_pattern(alignment, trees.TagTaxonMap());
      const auto edge_indexer = dag.BuildEdgeIndexer();
      TPEngine tp(dag, site_pattern, tag + ".r_lik", tag + ".r_pars", trees, edge_indexer);
      ref = RunSearch(tp, dag, trees, edge_indexer, iterations);
      std::printf("reference search: %zu iterations, DAG %zu edges at the end\n", ref.size(),
                  ref.empty() ? size_t(0) : ref.back().edges);
    }
    {
      GPDAG dag(trees);
      SitePattern site_pattern(alignment, trees.TagTaxonMap());
      const auto edge_indexer = dag.BuildEdgeIndexer();
      TPEngineWithEvaluator<TPEvalEngineOverGPEngine<GPEngine>> tp(dag, site_pattern, tag + ".c_lik", tag + ".c_pars",
                                                                    trees, edge_indexer);
      cpu = RunSearch(tp, dag, trees, edge_indexer, iterations);
    }
    Compare("evaluator over the CPU GPEngine vs TPEngine", cpu, ref, 1e-12, 1e-12, 1e-12);
    if (with_gpu) {
      GPDAG dag(trees);
      SitePattern site_pattern(alignment, trees.TagTaxonMap());
      const auto edge_indexer = dag.BuildEdgeIndexer();
      TPEngineWithEvaluator<TPEvalEngineOverGPEngine<GPEngineB200>> tp(dag, site_pattern, tag + ".g_lik",
                                                                        tag + ".g_pars", trees, edge_indexer);
      gpu = RunSearch(tp, dag, trees, edge_indexer, iterations);
      Compare("evaluator over GPEngineB200 (CUDA) vs TPEngine", gpu, ref, 1e-7, 1e-7, 1e-6);
    }
    for (const char* suffix : {".r_lik", ".r_pars", ".c_lik", ".c_lik.ref", ".c_lik.host", ".c_lik.gp", ".c_pars",
                               ".g_lik", ".g_lik.ref", ".g_lik.host", ".g_lik.gp", ".g_pars"})
      unlink((tag + suffix).c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "tp_search_parity: %s\n", e.what());
    return 1;
  }
  std::printf("%s\n", g_failures == 0 ? "TP SEARCH PARITY PASS" : "TP SEARCH PARITY FAIL");
  return g_failures == 0 ? 0 : 1;
}
