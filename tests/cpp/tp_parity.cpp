// tests/cpp/tp_parity.cpp — the TP likelihood evaluator through GP op lists.
//
// Builds a GPDAG and the reference's TPEngine (tp_engine.cpp, tp_evaluation_engine.cpp, unmodified)
// from a fasta + newick pair, lets the reference compute its top-tree log-likelihoods
// (TPEvalEngineViaLikelihood::Initialize + ComputeScores), then turns the SAME TPChoiceMap into two
// GPOperationVectors with bito_b200/host/tp_likelihood_plan.hpp and runs them
//   (cpu) through the unmodified reference CPU GPEngine — this checks the plan itself, no GPU needed;
//   (gpu) with `--gpu`, through GPEngineB200 (the host class over libbito_gp_b200.so).
// Per-edge top-tree log-likelihoods must agree to 1e-9 relative.
// Then the NNIs adjacent to the DAG are scored as PROPOSED NNIs (GetTopTreeScoreWithProposedNNI, :466-641:
// spare PVs and edges, branch lengths taken over from the pre-NNI, optionally five rounds of
// OptimizeBranchLength on the new edges) by the reference and by TPLikelihoodPlan::ProposedNNIOps on the
// same engines: scores to 1e-9 with fixed branch lengths, 1e-7 with optimised ones (optimised lengths 1e-6).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unistd.h>

#include "alignment.hpp"
#include "driver.hpp"
#include "gp_dag.hpp"
#include "gp_engine.hpp"
#include "gp_engine_b200.hpp"
#include "nni_engine.hpp"
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
// Scores `info`'s NNI on `engine` (whose DAG-edge PVs and branch lengths are in place).
template <typename Engine>
double ScoreProposedNNI(Engine& engine, const TPLikelihoodPlan& plan, const ProposedNNIInfo& info,
                        const NNIAdjDoubles& temp_branch_lengths, const bool optimize, const size_t max_iter,
                        NNIAdjDoubles* optimised) {
  const auto ops = plan.ProposedNNIOps(info);
  auto& handler = engine.GetBranchLengthHandler();
  for (auto adj : NNIAdjacentEnum::Iterator()) handler(info.temp_edge_ids[adj]) = temp_branch_lengths[adj];
  engine.ResetOptimizationCount();
  engine.ProcessOperations(ops.initialize);
  if (optimize)
    for (size_t iter = 0; iter < max_iter; ++iter) {
      engine.ProcessOperations(ops.iteration);
      engine.IncrementOptimizationCount();
    }
  engine.ProcessOperations(ops.score);
  const EigenVectorXd all = engine.GetBranchLengths(0, engine.GetPaddedGPCSPCount());
  for (auto adj : NNIAdjacentEnum::Iterator()) (*optimised)[adj] = all[info.temp_edge_ids[adj].value_];
  return engine.GetPerGPCSPLogLikelihoods(ops.focal_gpcsp, 1)[0];
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
  // --gpu-batched: additionally run the batched proposed-NNI lists on the CUDA engine (not part of the committed
  // GPU tests yet: verified on the CPU engine only this round)
  const bool gpu_batched = argc > 3 && std::strcmp(argv[3], "--gpu-batched") == 0;
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
    // ---- proposed NNIs -------------------------------------------------------------------------
    {
      padded.head(E) = branch_lengths;
      tp.SetBranchLengths(padded);
      tp.SelectLikelihoodEvalEngine();
      auto& eval = tp.GetLikelihoodEvalEngine();
      eval.Initialize();
      eval.ComputeScores();
      tp.GrowSpareNodeData(tp.GetSpareNodesPerNNI());  // NNIEvalEngineViaTP::GrowEngineForAdjacentNNIs,
      tp.GrowSpareEdgeData(tp.GetSpareEdgesPerNNI());  // nni_evaluation_engine.cpp:1050-1065
      NNIEngine nni_engine(dag, nullptr, nullptr);
      nni_engine.SyncAdjacentNNIsWithDAG();
      GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n, ones_g,
                   false);
      cpu.GrowSpareGPCSPs(8);
      RunPlan(cpu, plan, branch_lengths);
      std::unique_ptr<GPEngineB200> gpu;
      if (with_gpu) {
        gpu = std::make_unique<GPEngineB200>(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40,
                                             ones_g, ones_n, ones_g, false);
        gpu->GrowSpareGPCSPs(8);
        RunPlan(*gpu, plan, branch_lengths);
      }
      double worst[2][3] = {{0., 0., 0.}, {0., 0., 0.}};  // [optimize][cpu score, gpu score, gpu lengths]
      double worst_cpu_bl = 0., moved = 0.;  // moved: how far the reference's optimiser took a new edge
      size_t n_scored = 0;
      for (const auto& post_nni : nni_engine.GetAdjacentNNIs()) {
        if (n_scored >= 8) break;
        const auto pre_nni = dag.FindNNINeighborInDAG(post_nni);
        for (const bool optimize : {false, true}) {
          eval.SetOptimizeNewEdges(optimize);
          const ProposedNNIInfo info = eval.GetProposedNNIInfo(post_nni, pre_nni, 0, std::nullopt);
          // the lengths the reference starts from: the same rule, read from ITS handler before it scores
          NNIAdjDoubles start;
          {
            auto& handler = eval.GetDAGBranchHandler();
            TPLikelihoodPlan::InitializeTempBranchLengths(handler, info, handler.GetDefaultBranchLength());
            for (auto adj : NNIAdjacentEnum::Iterator()) start[adj] = handler(info.temp_edge_ids[adj]);
          }
          const double want_score = tp.GetTopTreeScoreWithProposedNNI(post_nni, pre_nni, 0, std::nullopt);
          NNIAdjDoubles want_bl;
          for (auto adj : NNIAdjacentEnum::Iterator())
            want_bl[adj] = eval.GetDAGBranchHandler()(info.temp_edge_ids[adj]);
          if (optimize)
            for (auto adj : NNIAdjacentEnum::Iterator())
              moved = std::max(moved, std::abs(want_bl[adj] - start[adj]));
          NNIAdjDoubles got_bl;
          const double cpu_score =
              ScoreProposedNNI(cpu, plan, info, start, optimize, eval.GetOptimizationMaxIteration(), &got_bl);
          worst[optimize][0] = std::max(worst[optimize][0], std::abs(cpu_score - want_score) / std::abs(want_score));
          for (auto adj : NNIAdjacentEnum::Iterator())
            worst_cpu_bl = std::max(worst_cpu_bl, std::abs(got_bl[adj] - want_bl[adj]));
          if (with_gpu) {
            const double gpu_score =
                ScoreProposedNNI(*gpu, plan, info, start, optimize, eval.GetOptimizationMaxIteration(), &got_bl);
            worst[optimize][1] =
                std::max(worst[optimize][1], std::abs(gpu_score - want_score) / std::abs(want_score));
            for (auto adj : NNIAdjacentEnum::Iterator())
              worst[optimize][2] = std::max(worst[optimize][2], std::abs(got_bl[adj] - want_bl[adj]));
          }
        }
        ++n_scored;
      }
      std::printf("proposed NNIs scored: %zu (the reference's optimiser moved a new edge by up to %.3e)\n", n_scored,
                  moved);
      if (n_scored > 0 && !(moved > 1e-4)) {
        std::printf("FAIL the optimised case did not optimise anything\n");
        ++g_failures;
      }
      if (n_scored > 0) {
        Report("proposed NNIs, fixed lengths: plan on CPU GPEngine vs TPEngine", worst[0][0], 1e-9);
        Report("proposed NNIs, optimised: plan on CPU GPEngine vs TPEngine", worst[1][0], 1e-9);
        Report("proposed NNIs, optimised branch lengths on CPU GPEngine", worst_cpu_bl, 1e-9);
        if (with_gpu) {
          Report("proposed NNIs, fixed lengths: plan on CUDA vs TPEngine", worst[0][1], 1e-9);
          Report("proposed NNIs, optimised: plan on CUDA vs TPEngine", worst[1][1], 1e-7);
          Report("proposed NNIs, optimised branch lengths on CUDA", worst[1][2], 1e-6);
        }
      }
    }
    // ---- every adjacent NNI in ONE batch of lists (TPLikelihoodPlan::BatchedProposedNNIOps) against the
    // reference scoring them one at a time -----------------------------------------------------------
    {
      padded = tp.GetBranchLengths();
      padded.head(E) = branch_lengths;
      tp.SetBranchLengths(padded);
      auto& eval = tp.GetLikelihoodEvalEngine();
      eval.SetOptimizeNewEdges(true);
      eval.Initialize();
      eval.ComputeScores();
      NNIEngine nni_engine(dag, nullptr, nullptr);
      nni_engine.SyncAdjacentNNIsWithDAG();
      std::vector<NNIOperation> posts;
      for (const auto& nni : nni_engine.GetAdjacentNNIs())
        if (posts.size() < 12) posts.push_back(nni);
      const size_t n = posts.size();
      // the reference scores them one after the other in the same temp slot (spare offset 0,
      // nni_evaluation_engine.cpp:1081-1085); the batch gives NNI i the ids of spare offset i
      std::vector<ProposedNNIInfo> infos;
      std::vector<NNIAdjDoubles> starts(n);
      std::vector<double> want(n);
      for (size_t i = 0; i < n; ++i) {
        const auto pre_nni = dag.FindNNINeighborInDAG(posts[i]);
        infos.push_back(eval.GetProposedNNIInfo(posts[i], pre_nni, i, std::nullopt));
        const auto& handler = eval.GetDAGBranchHandler();
        starts[i] = TPLikelihoodPlan::TempBranchLengths(handler, infos[i], handler.GetDefaultBranchLength());
        want[i] = tp.GetTopTreeScoreWithProposedNNI(posts[i], pre_nni, 0, std::nullopt);
      }
      const auto batch = plan.BatchedProposedNNIOps(infos);
      auto run_batch = [&](auto& engine) {
        engine.GrowSparePLVs(TPLikelihoodPlan::SpareNodesForBatch(n));
        // The reference engine only rebuilds its PV index map when it reallocates (pv_handler.hpp:173-177 looks
        // every PV up through pv_reindexer_, sized at the last Resize): ask for an explicit allocation so that
        // the new spare PLVs are addressable there too. A no-op for the CUDA engine's slot table.
        engine.GrowPLVs(engine.GetNodeCount(), std::nullopt, engine.GetNodeCount());
        engine.GrowSpareGPCSPs(TPLikelihoodPlan::SpareGPCSPsForBatch(n) + n);
        RunPlan(engine, plan, branch_lengths);
        auto& handler = engine.GetBranchLengthHandler();
        for (size_t i = 0; i < n; ++i)
          for (auto adj : NNIAdjacentEnum::Iterator()) handler(infos[i].temp_edge_ids[adj]) = starts[i][adj];
        engine.ResetOptimizationCount();
        engine.ProcessOperations(batch.initialize);
        for (size_t iter = 0; iter < eval.GetOptimizationMaxIteration(); ++iter) {
          engine.ProcessOperations(batch.iteration);
          engine.IncrementOptimizationCount();
        }
        engine.ProcessOperations(batch.score);
        double worst = 0.;
        for (size_t i = 0; i < n; ++i) {
          const double got = engine.GetPerGPCSPLogLikelihoods(batch.focal_gpcsp[i], 1)[0];
          if (getenv("TP_PARITY_DEBUG")) std::printf("  nni %zu focal %zu got %.12g want %.12g\n", i, batch.focal_gpcsp[i], got, want[i]);
          worst = std::max(worst, std::abs(got - want[i]) / std::abs(want[i]));
        }
        return worst;
      };
      if (n > 0) {
        GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n, ones_g,
                     false);
        std::printf("batched proposed NNIs: %zu NNIs, %zu + %zu x %zu + %zu ops\n", n, batch.initialize.size(),
                    size_t(eval.GetOptimizationMaxIteration()), batch.iteration.size(), batch.score.size());
        Report("batched proposed NNIs (optimised): plan on CPU GPEngine vs TPEngine", run_batch(cpu), 1e-9);
        if (gpu_batched) {
          GPEngineB200 gpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n,
                           ones_g, false);
          Report("batched proposed NNIs (optimised): plan on CUDA vs TPEngine", run_batch(gpu), 1e-7);
          const bito_gp_stats st = gpu.Stats();
          std::printf("CUDA engine: %lld kernel launches, %lld levels in the last list\n",
                      static_cast<long long>(st.kernel_launches), static_cast<long long>(st.levels_last));
        }
      }
    }
    // ---- whole-DAG branch-length optimisation (:988-1022), two calls: first without, then with the
    // convergence check -------------------------------------------------------------------------
    {
      padded.head(E) = branch_lengths;
      tp.SetBranchLengths(padded);
      auto& eval = tp.GetLikelihoodEvalEngine();
      eval.ResetOptimizationCount();
      eval.Initialize();
      const size_t rounds = eval.GetOptimizationMaxIteration();
      const GPOperationVector blo = plan.BranchLengthOptimizationOps();
      auto run_ours = [&](auto& engine) {
        for (size_t r = 0; r < rounds; ++r) engine.ProcessOperations(blo);
        for (size_t r = 0; r < rounds; ++r) engine.IncrementOptimizationCount();
        return EigenVectorXd(engine.GetBranchLengths());
      };
      GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40, ones_g, ones_n, ones_g,
                   false);
      RunPlan(cpu, plan, branch_lengths);
      cpu.ResetOptimizationCount();
      std::unique_ptr<GPEngineB200> gpu;
      if (with_gpu) {
        gpu = std::make_unique<GPEngineB200>(SitePattern(alignment, trees.TagTaxonMap()), N, G, tag + ".gp", 1e-40,
                                             ones_g, ones_n, ones_g, false);
        RunPlan(*gpu, plan, branch_lengths);
        gpu->ResetOptimizationCount();
      }
      double worst_cpu = 0., worst_gpu = 0., moved = 0.;
      for (int call = 0; call < 2; ++call) {
        eval.BranchLengthOptimization();
        const EigenVectorXd want_bl = eval.GetDAGBranchHandler().GetBranchLengthData().head(E);
        moved = std::max(moved, (want_bl - branch_lengths).cwiseAbs().maxCoeff());
        worst_cpu = std::max(worst_cpu, (run_ours(cpu) - want_bl).cwiseAbs().maxCoeff());
        if (with_gpu) worst_gpu = std::max(worst_gpu, (run_ours(*gpu) - want_bl).cwiseAbs().maxCoeff());
      }
      std::printf("whole-DAG optimisation: %zu rounds x 2 calls, %zu ops per round, lengths moved by up to %.3e\n",
                  rounds, blo.size(), moved);
      Report("BranchLengthOptimization: plan on CPU GPEngine vs TPEngine, |dBL|", worst_cpu, 1e-9);
      if (with_gpu) Report("BranchLengthOptimization: plan on CUDA vs TPEngine, |dBL|", worst_gpu, 1e-6);
    }
    // ---- the reference's NNI search in TP mode grows the DAG (nni_search.py:624-642: top-1 filter, new edges
    // optimised); after every iteration the plan is rebuilt for the grown DAG and must reproduce what the
    // reference gets when it re-evaluates that DAG from scratch (Initialize + ComputeScores) with the branch
    // lengths the search has left behind, and the next round of proposed NNIs --------------------------------
    {
      auto& eval = tp.GetLikelihoodEvalEngine();
      eval.SetOptimizeNewEdges(true);
      NNIEngine search(dag, nullptr, &tp);
      search.SetTPLikelihoodCutoffFilteringScheme(0.0);
      search.SetTopKScoreFilteringScheme(1);
      search.RunInit(true);
      double worst_cpu = 0., worst_gpu = 0., worst_nni_cpu = 0., worst_nni_gpu = 0., stale = 0.;
      size_t iterations = 0, grown_edges = E;
      for (; iterations < 3 && search.GetAdjacentNNICount() > 0; ++iterations) {
        search.RunMainLoop(true);
        search.RunPostLoop(true);
        const size_t E2 = dag.EdgeCountWithLeafSubsplits();
        grown_edges = E2;
        const EigenVectorXd bl2 = eval.GetDAGBranchHandler().GetBranchLengthData().head(E2);
        const EigenVectorXd incremental = tp.GetTopTreeLikelihoods().head(E2);  // what the search itself holds
        eval.Initialize();
        eval.ComputeScores();
        const EigenVectorXd want2 = tp.GetTopTreeLikelihoods().head(E2);
        stale = std::max(stale, RelErr(incremental, want2));
        const TPLikelihoodPlan plan2(dag, tp.GetChoiceMap());
        const size_t N2 = plan2.EngineNodeCount(), G2 = plan2.EngineGPCSPCount();
        const EigenVectorXd ones_g2 = EigenVectorXd::Ones(G2), ones_n2 = EigenVectorXd::Ones(N2);
        GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N2, G2, tag + ".gp", 1e-40, ones_g2, ones_n2,
                     ones_g2, false);
        cpu.GrowSpareGPCSPs(8);
        worst_cpu = std::max(worst_cpu, RelErr(RunPlan(cpu, plan2, bl2), want2));
        std::unique_ptr<GPEngineB200> gpu;
        if (with_gpu) {
          gpu = std::make_unique<GPEngineB200>(SitePattern(alignment, trees.TagTaxonMap()), N2, G2, tag + ".gp",
                                               1e-40, ones_g2, ones_n2, ones_g2, false);
          gpu->GrowSpareGPCSPs(8);
          worst_gpu = std::max(worst_gpu, RelErr(RunPlan(*gpu, plan2, bl2), want2));
        }
        // the NNIs the search will score next, on the grown DAG
        size_t n = 0;
        for (const auto& post_nni : search.GetAdjacentNNIs()) {
          if (n++ >= 4) break;
          const auto pre_nni = dag.FindNNINeighborInDAG(post_nni);
          const ProposedNNIInfo info = eval.GetProposedNNIInfo(post_nni, pre_nni, 0, std::nullopt);
          NNIAdjDoubles start, unused;
          auto& handler = eval.GetDAGBranchHandler();
          TPLikelihoodPlan::InitializeTempBranchLengths(handler, info, handler.GetDefaultBranchLength());
          for (auto adj : NNIAdjacentEnum::Iterator()) start[adj] = handler(info.temp_edge_ids[adj]);
          const double want_score = tp.GetTopTreeScoreWithProposedNNI(post_nni, pre_nni, 0, std::nullopt);
          const double cpu_score =
              ScoreProposedNNI(cpu, plan2, info, start, true, eval.GetOptimizationMaxIteration(), &unused);
          worst_nni_cpu = std::max(worst_nni_cpu, std::abs(cpu_score - want_score) / std::abs(want_score));
          if (with_gpu) {
            const double gpu_score =
                ScoreProposedNNI(*gpu, plan2, info, start, true, eval.GetOptimizationMaxIteration(), &unused);
            worst_nni_gpu = std::max(worst_nni_gpu, std::abs(gpu_score - want_score) / std::abs(want_score));
          }
        }
      }
      std::printf("TP-mode NNI search: %zu iterations, DAG grew from %zu to %zu edges; the reference's incrementally "
                  "updated scores differ from its own re-evaluation by up to %.3e (relative)\n",
                  iterations, E, grown_edges, stale);
      if (iterations > 0) {
        Report("grown DAG: plan on CPU GPEngine vs TPEngine", worst_cpu, 1e-9);
        Report("grown DAG, next proposed NNIs (optimised): plan on CPU GPEngine", worst_nni_cpu, 1e-9);
        if (with_gpu) {
          Report("grown DAG: plan on CUDA vs TPEngine", worst_gpu, 1e-9);
          Report("grown DAG, next proposed NNIs (optimised): plan on CUDA", worst_nni_gpu, 1e-7);
        }
      }
    }
    // ---- the INCREMENTAL update the search itself performs (TPEvalEngineViaLikelihood::UpdateEngineAfterModifyingDAG,
    // tp_evaluation_engine.cpp:267-460), replayed by TPLikelihoodPlan::UpdateAfterModifyingDAGOps on engines that
    // PERSIST across the DAG's growth: they are grown with the search's reindexers, given the starting lengths the
    // TP engine gave the new edges, and must then hold the scores and branch lengths the reference holds after its
    // own partial update (which are NOT those of a re-evaluation: see `stale` above) --------------------------------
    {
      auto& eval = tp.GetLikelihoodEvalEngine();
      eval.SetOptimizeNewEdges(true);
      eval.Initialize();
      eval.ComputeScores();
      NNIEngine search(dag, nullptr, &tp);
      search.SetTPLikelihoodCutoffFilteringScheme(0.0);
      search.SetTopKScoreFilteringScheme(1);
      size_t E0 = dag.EdgeCountWithLeafSubsplits();
      auto plan0 = std::make_unique<TPLikelihoodPlan>(dag, tp.GetChoiceMap());
      const size_t N0 = plan0->EngineNodeCount(), G0 = plan0->EngineGPCSPCount();
      const EigenVectorXd ones_g0 = EigenVectorXd::Ones(G0), ones_n0 = EigenVectorXd::Ones(N0);
      GPEngine cpu(SitePattern(alignment, trees.TagTaxonMap()), N0, G0, tag + ".gp", 1e-40, ones_g0, ones_n0, ones_g0,
                   false);
      std::unique_ptr<GPEngineB200> gpu;
      if (with_gpu)
        gpu = std::make_unique<GPEngineB200>(SitePattern(alignment, trees.TagTaxonMap()), N0, G0, tag + ".gp", 1e-40,
                                             ones_g0, ones_n0, ones_g0, false);
      {
        const EigenVectorXd bl0 = eval.GetDAGBranchHandler().GetBranchLengthData().head(E0);
        RunPlan(cpu, *plan0, bl0);
        if (gpu) RunPlan(*gpu, *plan0, bl0);
      }
      double score_cpu = 0., score_gpu = 0., bl_cpu = 0., bl_gpu = 0.;
      size_t updates = 0, updated_edges = 0, optimised_edges = 0;
      auto replay = [&](auto& engine, const TPLikelihoodPlan& plan, const Reindexer& edge_reindexer,
                        const TPLikelihoodPlan::ModifiedDAGUpdate& ops, const EigenVectorXd& start_lengths,
                        const bool optimize, const size_t max_iter) {
        engine.GrowPLVs(plan.EngineNodeCount(), plan.EngineNodeReindexer(edge_reindexer));
        engine.GrowGPCSPs(plan.EngineGPCSPCount(), edge_reindexer);
        engine.SetNullPrior();
        engine.SetBranchLengths(start_lengths);
        engine.ResetOptimizationCount();
        engine.ProcessOperations(ops.initialize);
        if (optimize)
          for (size_t iter = 0; iter < max_iter; ++iter) engine.ProcessOperations(ops.iteration);
        engine.ProcessOperations(ops.score);
      };
      search.SetFilterPostModificationFunction(
          [&](NNIEngine& this_nni_engine, const SubsplitDAG::ModificationResult& mods,
              const std::map<NNIOperation, NNIOperation>& nni_to_pre_nni) {
            // what SetScoreViaEvalEngine's function does (nni_engine.cpp:458-468), with TPEngine::UpdateAfterModifyingDAG
            // (tp_engine.cpp:238-260) opened up so that the evaluator's part can be replayed beside it
            this_nni_engine.GrowEvalEngineForDAG(mods.node_reindexer, mods.edge_reindexer);
            tp.UpdateChoiceMapAfterModifyingDAG(nni_to_pre_nni, mods.prv_node_count, mods.node_reindexer,
                                                mods.prv_edge_count, mods.edge_reindexer);
            const size_t E2 = dag.EdgeCountWithLeafSubsplits();
            const EigenVectorXd start = eval.GetDAGBranchHandler().GetBranchLengthData().head(E2);
            const TPLikelihoodPlan plan2(dag, tp.GetChoiceMap());
            const auto ops = plan2.UpdateAfterModifyingDAGOps(nni_to_pre_nni, mods.prv_edge_count, mods.edge_reindexer);
            replay(cpu, plan2, mods.edge_reindexer, ops, start, eval.IsOptimizeNewEdges(),
                   eval.GetOptimizationMaxIteration());
            if (gpu)
              replay(*gpu, plan2, mods.edge_reindexer, ops, start, eval.IsOptimizeNewEdges(),
                     eval.GetOptimizationMaxIteration());
            eval.UpdateEngineAfterModifyingDAG(nni_to_pre_nni, mods.prv_node_count, mods.node_reindexer,
                                               mods.prv_edge_count, mods.edge_reindexer);
            const EigenVectorXd want_scores = tp.GetTopTreeLikelihoods().head(E2);
            const EigenVectorXd want_bl = eval.GetDAGBranchHandler().GetBranchLengthData().head(E2);
            auto compare = [&](auto& engine, double& worst_score, double& worst_bl) {
              const EigenVectorXd got = engine.GetPerGPCSPLogLikelihoods();
              for (const auto edge_id : ops.update_edges)
                worst_score = std::max(worst_score, std::abs(got[edge_id.value_] - want_scores[edge_id.value_]) /
                                                        std::max(1.0, std::abs(want_scores[edge_id.value_])));
              const EigenVectorXd bl = engine.GetBranchLengths(0, E2);
              worst_bl = std::max(worst_bl, (bl - want_bl).cwiseAbs().maxCoeff());
            };
            compare(cpu, score_cpu, bl_cpu);
            if (gpu) compare(*gpu, score_gpu, bl_gpu);
            ++updates;
            updated_edges += ops.update_edges.size();
            optimised_edges += E2 - mods.prv_edge_count;
          });
      search.RunInit(true);
      for (size_t iteration = 0; iteration < 3 && search.GetAdjacentNNICount() > 0; ++iteration) {
        search.RunMainLoop(true);
        search.RunPostLoop(true);
        // both sides back to the state of a fresh evaluation of the grown DAG (the next update starts from it)
        const size_t E2 = dag.EdgeCountWithLeafSubsplits();
        const EigenVectorXd bl2 = eval.GetDAGBranchHandler().GetBranchLengthData().head(E2);
        eval.Initialize();
        eval.ComputeScores();
        const TPLikelihoodPlan plan2(dag, tp.GetChoiceMap());
        RunPlan(cpu, plan2, bl2);
        if (gpu) RunPlan(*gpu, plan2, bl2);
        E0 = E2;
      }
      std::printf("incremental updates replayed: %zu (DAG now %zu edges; %zu edges refreshed, %zu new edges optimised)\n",
                  updates, E0, updated_edges, optimised_edges);
      if (updates > 0) {
        Report("incremental update: plan on CPU GPEngine vs TPEngine, scores", score_cpu, 1e-9);
        Report("incremental update: plan on CPU GPEngine vs TPEngine, |dBL|", bl_cpu, 1e-9);
        if (with_gpu) {
          Report("incremental update: plan on CUDA vs TPEngine, scores", score_gpu, 1e-7);
          Report("incremental update: plan on CUDA vs TPEngine, |dBL|", bl_gpu, 1e-6);
        }
      }
    }
    for (const char* suffix : {".tp_lik", ".tp_pars", ".gp"}) unlink((tag + suffix).c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "tp_parity: %s\n", e.what());
    return 1;
  }
  std::printf("%s\n", g_failures == 0 ? "TP PARITY PASS" : "TP PARITY FAIL");
  return g_failures == 0 ? 0 : 1;
}
