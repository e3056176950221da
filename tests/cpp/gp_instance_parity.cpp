// tests/cpp/gp_instance_parity.cpp — the reference's own GPInstance on top of a GPEngine, printed as text.
//
// ONE source, compiled twice by oracle/Makefile (`make gpinstanceparity`):
//   oracle/_ref/gp_instance_parity_ref   the reference's gp_instance.cpp + CPU GPEngine (only fat_beagle.hpp is a
//                                        stub: libhmsbeagle is not in the image and is not on the GP path);
//   oracle/_ref/gp_instance_parity_b200  the SAME gp_instance.cpp, unchanged, compiled against
//                                        bito_b200/host/gp_engine_b200.hpp installed as gp_engine.hpp.
// It drives GPInstance the way bito's users (and BASELINE.json configs[0..2]) do: MakeGPEngine, PopulatePLVs +
// ComputeLikelihoods (gp_instance.cpp:231-235), EstimateBranchLengths (:241-308) + ComputeMarginalLikelihood,
// EstimateSBNParameters (:401-406), CalculateHybridMarginals (:408-417), HotStart / TakeFirst branch lengths,
// the CSV exporters. tests/test_gp_instance_parity_gpu.py compares the two outputs.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <unistd.h>

#include "gp_instance.hpp"

namespace {
void PrintVector(const char* name, const EigenVectorXd& v) {
  std::printf("%s %zu", name, size_t(v.size()));
  for (Eigen::Index i = 0; i < v.size(); ++i) std::printf(" %.17g", v[i]);
  std::printf("\n");
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s fasta newick [rescaling_threshold] [max_iter] [other_build_branch_lengths]\n", argv[0]);
    return 2;
  }
  const double threshold = argc > 3 ? std::atof(argv[3]) : GPEngine::default_rescaling_threshold_;
  const size_t max_iter = argc > 4 ? size_t(std::atoi(argv[4])) : 3;
  const std::string mmap_path = std::string("/tmp/gp_instance_parity_") + std::to_string(getpid()) + ".plv";
  try {
    GPInstance inst(mmap_path);
    inst.ReadFastaFile(argv[1]);
    inst.ReadNewickFile(argv[2], false);
    inst.MakeGPEngine(threshold);
    GPEngine& engine = inst.GetGPEngine();
    std::printf("dag nodes %zu edges %zu patterns %zu\n", inst.GetDAG().NodeCountWithoutDAGRoot(),
                inst.GetDAG().EdgeCountWithLeafSubsplits(), engine.GetSitePatternCount());

    // configs[0]: full pass at the default branch lengths
    inst.PopulatePLVs();
    inst.ComputeLikelihoods();
    inst.ComputeMarginalLikelihood();
    PrintVector("pass_per_gpcsp_llh", engine.GetPerGPCSPLogLikelihoods());
    std::printf("pass_log_marginal %.17g\n", engine.GetLogMarginalLikelihood());

    // branch lengths from the loaded trees (gp_engine.cpp:676-746 through gp_instance.cpp:389-399)
    inst.TakeFirstBranchLength();
    PrintVector("take_first_branch_lengths", engine.GetBranchLengths());
    inst.HotStartBranchLengths();
    PrintVector("hot_start_branch_lengths", engine.GetBranchLengths());

    // configs[1]: EstimateBranchLengths + ComputeMarginalLikelihood
    inst.EstimateBranchLengths(1e-6, max_iter, true);
    inst.PopulatePLVs();
    inst.ComputeLikelihoods();
    inst.ComputeMarginalLikelihood();
    PrintVector("estimated_branch_lengths", engine.GetBranchLengths());
    PrintVector("estimated_per_gpcsp_llh", engine.GetPerGPCSPLogLikelihoods());
    std::printf("estimated_log_marginal %.17g\n", engine.GetLogMarginalLikelihood());
    if (argc > 5) {
      // Flatness check for edges on which the two builds' optimised lengths differ by more than 1e-6: the
      // objective of edge e (Likelihood(e): parent R PLV, M(t_e), child P PLV) depends on t_e only through M, so
      // with THIS build's PLVs in place, set those edges to the other build's lengths and score them again.
      // Prints, per such edge: index, log-likelihood at the own length, log-likelihood at the other build's.
      std::ifstream in(argv[5]);
      std::vector<double> other;
      for (double x; in >> x;) other.push_back(x);
      const EigenVectorXd own = engine.GetBranchLengths();
      const EigenVectorXd own_llh = engine.GetPerGPCSPLogLikelihoods();
      EigenVectorXd mixed = own;
      std::vector<Eigen::Index> off;
      for (Eigen::Index i = 0; i < own.size() && size_t(i) < other.size(); ++i)
        if (std::fabs(own[i] - other[size_t(i)]) > 1e-6) {
          mixed[i] = other[size_t(i)];
          off.push_back(i);
        }
      engine.SetBranchLengths(mixed);
      inst.ComputeLikelihoods();
      const EigenVectorXd other_llh = engine.GetPerGPCSPLogLikelihoods();
      std::printf("flatness %zu", off.size());
      for (Eigen::Index i : off) std::printf(" %ld %.17g %.17g", long(i), own_llh[i], other_llh[i]);
      std::printf("\n");
      engine.SetBranchLengths(own);
      inst.ComputeLikelihoods();
    }

    // configs[2]: SBN probability update, then with hybrid marginals
    inst.EstimateSBNParameters();
    PrintVector("sbn_parameters", inst.GetSBNParameters());
    inst.CalculateHybridMarginals();
    PrintVector("hybrid_marginals", engine.GetHybridMarginals());
    inst.EstimateSBNParameters();
    PrintVector("sbn_parameters_after_hybrid", inst.GetSBNParameters());

    // the exporters read everything back through the engine's getters
    const std::string csv = mmap_path + ".csv";
    inst.BranchLengthsToCSV(csv);
    inst.PerGPCSPLogLikelihoodsToCSV(csv);
    inst.SBNParametersToCSV(csv);
    unlink(csv.c_str());
    std::printf("exporters ok\n");
    unlink(mmap_path.c_str());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "gp_instance_parity: %s\n", e.what());
    return 1;
  }
  return 0;
}
