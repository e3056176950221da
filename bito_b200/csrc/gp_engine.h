// bito_b200/csrc/gp_engine.h — host side of the B200 GP engine.
//
// Owns every piece of numeric state the reference GPEngine owns
// (/root/reference/src/gp_engine.hpp:287-377) but in HBM, compiles each GPOperationVector
// into a level-scheduled program of fused macro-ops (the role of the serial visitor loop
// GPEngine::ProcessOperations, gp_engine.cpp:335-339) and launches the kernels of
// gp_kernels.cu. The C-ABI in gp_c_api.cu is a thin wrapper over this class.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bito_gp.h"
#include "gp_types.h"

namespace bito_gp {

struct GpError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Fixed-size slot allocator over a few large cudaMalloc chunks. PLVs (32*P_stride bytes)
// and log-likelihood rows (8*P_stride bytes) are handed out on first write, so a DAG only
// pays HBM for the vectors its op lists actually produce.
class SlabPool {
 public:
  SlabPool() = default;
  void Init(size_t slot_bytes, size_t target_chunk_bytes);
  void* Alloc();
  void Free(void* p) { free_.push_back(p); }
  void Release();
  size_t BytesReserved() const { return chunks_.size() * slots_per_chunk_ * slot_bytes_; }
  size_t SlotsInUse() const { return handed_out_ - free_.size(); }
  size_t slot_bytes() const { return slot_bytes_; }
  bool NeedsChunk() const { return free_.empty() && next_in_chunk_ == slots_per_chunk_; }
  size_t ChunkBytes() const { return slots_per_chunk_ * slot_bytes_; }

 private:
  size_t slot_bytes_ = 0, slots_per_chunk_ = 0, next_in_chunk_ = 0, handed_out_ = 0;
  std::vector<char*> chunks_;
  std::vector<void*> free_;
};

template <typename T>
struct DeviceArray {
  T* ptr = nullptr;
  size_t n = 0;
  void Resize(size_t count, bool keep, cudaStream_t stream);
  void Release();
};

struct PlvSlot {
  void* ptr = nullptr;
  int32_t kind = kPlvZero;
};

struct Level {
  int zero_off = 0, n_zero = 0;
  int scalar_off = 0, n_scalar = 0;
  int stat_off = 0, n_stat = 0;
  int node_off = 0, n_node = 0;  // fused accumulate/multiply macro-ops (k_node)
  int mult_off = 0, n_mult = 0;  // the Multiplies inside those nodes (rescale decisions)
  int lik_off = 0, n_lik = 0;
  int marg_off = 0, n_marg = 0, marg_reset = 0, has_marg = 0, marg_scatter_off = 0;
  int opt_off = 0, n_opt = 0;
  double node_bytes_per_pattern = 0.;
  bool rebuild_matrices = false;  // first level, or q may have changed (UpdateSBNProbabilities) since the tables were built
  bool rebuild_after_opt = false; // the previous level optimised branch lengths: a full rebuild only if it ran
                                  // the round scheme (the on-chip optimisers refresh their own edges' slots)
};

struct Program {
  uint64_t alloc_version = 0;
  uint64_t check_hash = 0;  // second hash of the op list, compared on every cache hit
  int64_t n_ops = 0, vec_len = 0;
  uint64_t last_used = 0;
  std::vector<Level> levels;
  // device tables (one arena)
  char* arena = nullptr;
  size_t arena_bytes = 0;
  ZeroOp* d_zero = nullptr;
  ScalarOp* d_scalar = nullptr;
  StatOp* d_stat = nullptr;
  NodeOp* d_node = nullptr;
  AccumItem* d_items = nullptr;
  MultOp* d_mult = nullptr;
  LikOp* d_lik = nullptr;
  MargItem* d_marg = nullptr;
  OptOp* d_opt = nullptr;
  int32_t* d_pool = nullptr;
  int32_t* d_lik_scatter = nullptr;   // per LikOp: edge
  int32_t* d_marg_scatter = nullptr;  // per marginal level: n_marg edges + marginal slot
  int n_mult_total = 0;
  int n_opt_total = 0;
  int n_opt_levels = 0;       // levels that hold OptimizeBranchLength ops
  int64_t n_items_total = 0;  // transition matrices of all accumulate items (mtab slots)
  int n_lik_total = 0;        // Likelihood ops (lik mtab slots)
  int64_t max_partials = 0;  // doubles of tile partials needed by any level
  int64_t max_packed = 0;    // reduced scalars needed by any level
  int64_t n_macro = 0;
  int64_t launches = 0;      // kernels per execution (without optimiser rounds)
  double alg_bytes_per_pattern = 0.;
  int64_t graph_opt_launches = 0;  // optimiser kernels inside the captured graph
  cudaGraphExec_t graph = nullptr;
  bool graph_tried = false;
};

enum ProfKind {
  kProfZero, kProfScalar, kProfStationary, kProfPrologue, kProfNode, kProfRescale, kProfLikelihood,
  kProfMarginal, kProfReduce, kProfOptPrepare, kProfOptEval, kProfOptStep, kProfOptBlock, kProfOptCluster,
  kProfKinds
};

class Engine;
// Brackets one kernel launch with CUDA events on the launching stream when profiling is on.
class ProfScope {
 public:
  ProfScope(Engine* e, int kind, double bytes);
  ~ProfScope();

 private:
  Engine* e_;
};

class Engine {
 public:
  friend class ProfScope;
  void SetProfiling(bool on);
  int GetKernelProfile(bito_gp_kernel_profile* out, int capacity);
  void ResetKernelProfile();

  explicit Engine(const bito_gp_config& cfg);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  void SetSitePatterns(const uint8_t* symbols, const double* weights, bool on_device);
  void InitializePriors(const double* sbn_prior, const double* unconditional,
                        const double* inverted);
  void SetNullPrior();
  void ProcessOperations(const bito_gp_op* ops, int64_t n, const int64_t* vec, int64_t vec_len);

  void SetBranchLengths(const double* bl);
  void SetBranchLengthsRange(int64_t start, int64_t length, const double* bl);
  void SetBranchLengthsToConstant(double v);
  void SetBranchLengthsToDefault();  // dag_branch_handler.hpp:266 default_branch_length_
  void GetBranchLengths(int64_t start, int64_t length, double* out);
  void GetBranchLengthDifferences(double* out);
  void SetOptimizationMethod(int m);
  void SetSignificantDigits(int d) { significant_digits_ = d; }
  int64_t optimization_count() const { return optimization_count_; }
  void ResetOptimizationCount();
  void IncrementOptimizationCount() { optimization_count_++; }
  void LogLikelihoodAndDerivatives(int64_t gpcsp, int64_t rootward, int64_t leafward,
                                   double out[3]);
  void GetTransitionMatrix(double t, double out[16]);
  void SetSubstitutionModel(const double* v, const double* vinv, const double* lambda, const double* pi);

  double GetLogMarginalLikelihood();
  void GetPerGpcspLogLikelihoods(int64_t start, int64_t length, double* out);
  void GetPerGpcspComponentsOfFullLogMarginal(double* out);
  void GetLogLikelihoodMatrix(double* out);
  void GetPerPatternLogMarginal(double* out);
  void GetSbnParameters(double* out);
  void SetSbnParameters(const double* q);
  void GetPlv(int64_t id, double* out);
  void SetPlv(int64_t id, const double* in, int32_t count);
  void GetRescalingCounts(int32_t* out);
  // Quartet hybrid marginals (gp_engine.cpp:748-816). Request r owns tip_counts[4r..4r+3] tips
  // (rootward, sister, rotated, sorted) taken consecutively from `tips`. likelihoods (nullable):
  // every summand in the reference's loop order; store: LogSum of each fully formed request goes to
  // hybrid_marginal_log_likelihoods_[central].
  void QuartetHybrid(int64_t n_requests, const int64_t* central, const int32_t* tip_counts,
                     const bito_gp_quartet_tip* tips, double* likelihoods, bool store);
  void GetHybridMarginals(double* out);

  int64_t node_count() const { return node_count_; }
  int64_t plv_count() const { return 6 * node_count_; }
  int64_t padded_plv_count() const { return 6 * (node_count_ + spare_nodes_); }
  int64_t gpcsp_count() const { return gpcsp_count_; }
  int64_t padded_gpcsp_count() const { return gpcsp_count_ + spare_gpcsps_; }
  int64_t pattern_count() const { return P_; }

  void GrowPlvs(int64_t new_node_count, const int64_t* reindexer, int64_t explicit_alloc);
  void GrowGpcsps(int64_t new_count, const int64_t* reindexer, int64_t explicit_alloc);
  void GrowSparePlvs(int64_t new_spare);
  void GrowSpareGpcsps(int64_t new_spare);
  void CopyNodeData(int64_t src, int64_t dest);
  void CopyPlvData(int64_t src, int64_t dest);
  void CopyGpcspData(int64_t src, int64_t dest);

  void CommInit(int n_ranks, int rank, const uint8_t id[128]);
  void SetStream(cudaStream_t s);
  void Synchronize();
  void GetStats(bito_gp_stats* out);

 private:
  // state helpers
  void Activate() const;
  void InstallModel(const double* v, const double* vinv, const double* lambda, const double* pi);
  void BindModel();  // makes the constant-memory eigensystem this engine's before it launches anything
  ModelConst model_{};
  uint64_t model_id_ = 0;
  DeviceState State() const;
  void AllocEdgeArrays(int64_t padded);
  void EnsureDense(int64_t plv_id);  // allocate (zero-filled / expanded) HBM for a PLV
  double* EnsureRow(int64_t edge);
  PlvRef Ref(int64_t plv_id) const;
  void CheckPlv(int64_t id, const char* what) const;
  void CheckEdge(int64_t id, const char* what) const;
  void CheckStatus();
  void InvalidatePrograms();

  // scheduling
  Program* Compile(const bito_gp_op* ops, int64_t n, const int64_t* vec, int64_t vec_len);
  void Execute(Program& prog);
  void ExecuteLevels(Program& prog, size_t first, size_t last);
  void RunOptimizeLevel(Program& prog, const Level& lv);
  int OptScheme(int n_ops, const OptClusterPlan** plan) const;
  bool ProgramOptimizesOnChip(const Program& prog) const;
  OptParams OptimizerParams(bool check_convergence) const;
  void RunOptimizer(const OptOp* d_ops, int n_ops, int method, bool check_convergence);
  void RunOptimizerPipelined(const OptOp* d_ops, int n_ops, const OptClusterPlan& plan);
  int EnsurePipelineBuffers(int n_ops, const OptClusterPlan& plan);  // returns the chunk size (edges)
  int64_t ring_rho_stride_ = 0;
  uint64_t layout_version_ = 0, ring_layout_version_ = 0;
  int64_t capture_opt_launches_ = 0;
  void FreeProgram(Program& p);
  void EvictPrograms(const Program* keep);
  static constexpr size_t kMaxCachedPrograms = 48;
  uint64_t program_clock_ = 0;
  void EnsureScratch(int64_t partial_doubles, int64_t packed_doubles);
  void BuildWeightClasses(const double* host_weights);
  void DropGraphs();
  void AllReduce(double* buf, int64_t n, bool max_op);

  bito_gp_config cfg_;
  int device_ = 0;
  cudaStream_t stream_ = nullptr;
  cudaStream_t own_stream_ = nullptr;
  cudaEvent_t ev_begin_ = nullptr, ev_end_ = nullptr;

  int64_t taxon_count_ = 0, P_ = 0, P_stride_ = 0, site_count_ = 0;
  int64_t node_count_ = 0, gpcsp_count_ = 0, spare_nodes_ = 16, spare_gpcsps_ = 3;
  double thr_ = 1e-40, total_weight_ = 0.;
  int method_ = 0, significant_digits_ = 10;
  int64_t optimization_count_ = 0;
  bool have_patterns_ = false;
  bool timing_pending_ = false;
  bool status_pending_ = false;  // a ProcessOperations call returned without reading the device status word
  int n_eigen_groups_ = 2;
  double group_lambda_[kMaxEigenGroups] = {0., 0., 0., 0.};
  int64_t opt_chunk_bytes_ = 0;  // coefficient scratch per optimiser batch (0: 1 GiB)

  SlabPool plv_pool_, row_pool_;
  std::vector<PlvSlot> plvs_;      // by logical PLV id (padded count)
  std::vector<double*> rows_;      // by edge id (padded count)
  uint64_t alloc_version_ = 1;
  int64_t max_device_bytes_ = 0;

  DeviceArray<uint8_t> d_symbols_;
  DeviceArray<OptControl> d_opt_ctl_;  // optimiser settings read by k_opt_block
  DeviceArray<double> d_weights_, d_log_marg_;
  DeviceArray<int32_t> d_counts_;
  DeviceArray<double> d_q_, d_bl_, d_diff_, d_hybrid_, d_ll_sum_, d_inverted_, d_uncond_;
  DeviceArray<uint32_t> d_status_;
  DeviceArray<unsigned long long> d_feval_total_;
  // scratch
  DeviceArray<double> d_partials_, d_packed_, d_level_max_, d_coef_, d_dense_tmp_, d_mtab_, d_mtab_lik_;
  DeviceArray<OptState> d_opt_states_;
  DeviceArray<OptPass> d_opt_pass_;  // Taylor-model Brent: per-edge pass request, cache and model
  DeviceArray<OptReq> d_opt_req_;    // the pass lists of the current and the next round
  OptClassStarts opt_class_starts_ = {};
  DeviceArray<double> d_opt_const_;
  DeviceArray<int32_t> d_active_, d_opt_active_, d_perm_, d_pos_w_;
  DeviceArray<double> d_wperm_;
  DeviceArray<uint8_t> d_row_class_;
  int64_t P_perm_ = 0;  // patterns in weight-class order, classes padded to 256-pattern rows
  // k_opt_cluster's layout: position -> pattern (-1 = padding), weights by position, class rows
  DeviceArray<int32_t> d_cluster_inv_perm_;
  DeviceArray<int32_t> d_cluster_pos_;  // pattern -> position (the inverse of d_cluster_inv_perm_)
  // Pipelined cluster scheme (RunOptimizerPipelined): rho of two chunks of edges, their K_e and the
  // producer's tile partials, double-buffered between the producer stream and the engine's stream
  DeviceArray<double> d_rho_ring_, d_ring_const_, d_ring_partials_;
  cudaStream_t prep_stream_ = nullptr, cons_stream_ = nullptr;
  cudaEvent_t ev_join_ = nullptr;
  int prep_blocks_per_sm_ = 0;    // BITO_GP_PREP_BLOCKS_PER_SM: fixed producer grid (0: one block per item)
  int opt_priority_env_ = 1;      // BITO_GP_OPT_PRIORITY=0: consumer on the engine's stream (no priority)
  cudaEvent_t ev_fork_ = nullptr, ev_ready_[2] = {nullptr, nullptr}, ev_free_[2] = {nullptr, nullptr};
  int opt_scheme_env_ = -1;       // BITO_GP_OPT_SCHEME (-1: automatic)
  int opt_model_env_ = 1;         // BITO_GP_OPT_MODEL=0: one streamed pass per objective evaluation (round-1 scheme)
  int opt_model_first_check_ = 4; // rounds before the host first looks at the active-edge counter
  double local_min_weight_ = 1., min_weight_ = 1.;  // smallest positive pattern weight: this rank / all ranks
  int opt_ring_edges_env_ = 0;    // BITO_GP_OPT_RING_EDGES (0: automatic)
  DeviceArray<double> d_cluster_wperm_;
  int32_t cluster_class_row_start_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<OptClusterPlan> cluster_plans_;  // every shape this device runs for this alignment
  int opt_cluster_env_ = -1;          // BITO_GP_OPT_CLUSTER (-1: automatic)
  int opt_cluster_threads_env_ = -1;  // BITO_GP_OPT_CLUSTER_THREADS
  int last_opt_scheme_ = 0;           // scheme and shape of the most recent optimiser level
  OptClusterPlan last_opt_plan_;
  bool capturing_ = false;            // inside cudaStreamBeginCapture .. EndCapture
  OptRefresh opt_refresh_{};          // where the on-chip optimisers refresh their edges' matrices
  bool matrices_stale_ = false;       // a round-scheme optimiser level changed branch lengths
  bool coef_padding_zeroed_ = false;
  std::vector<double> host_weights_cache_;
  DeviceArray<OptOp> d_single_opt_;
  DeviceArray<QuartetItem> d_quartet_items_;
  DeviceArray<double> d_quartet_mats_;
  void* pinned_ = nullptr;  // small pinned staging block

  std::unordered_map<uint64_t, std::unique_ptr<Program>> programs_;

  // NCCL (dlopen'ed)
  void* nccl_comm_ = nullptr;
  int n_ranks_ = 1, rank_ = 0;
  // NVLink peer memory for the scalar all-reduces (k_peer_allreduce); NCCL remains the fallback
  void SetUpPeerMemory();
  void ReleasePeerMemory();
  PeerComm peer_{};
  void* peer_local_ = nullptr;       // this rank's exchange buffer (cudaMalloc + IPC handle)
  std::vector<void*> peer_opened_;   // the other ranks' buffers as mapped here
  DeviceArray<unsigned long long> d_peer_epoch_, d_peer_seq_;
  DeviceArray<PeerComm> d_peer_comm_;  // peer_ for kernels whose lanes index base[] by rank
  PeerEdge PeerEdgeContext() const;
  void AgreeOnClusterScheme();
  int multi_rank_cluster_ops_ = 0;   // largest level (edges) that may take the cluster path, agreed by all ranks
  bool peer_ready_ = false;

  bito_gp_stats stats_{};

  struct ProfEvent {
    cudaEvent_t begin, end;
    int kind;
    double bytes;
  };
  void CollectProfile();
  bool profiling_ = false;
  std::vector<ProfEvent> prof_events_;
  bito_gp_kernel_profile prof_[kProfKinds] = {};
};

void MakeNcclUniqueId(uint8_t id[128]);

// Thread-local error text for the C-ABI.
void SetLastError(const std::string& msg);
const char* LastError();

}  // namespace bito_gp
