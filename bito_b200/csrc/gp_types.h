// bito_b200/csrc/gp_types.h — device-visible tables of the level-scheduled GP engine.
//
// The host scheduler (gp_engine.cu) compiles a GPOperationVector
// (/root/reference/src/gp_operation.hpp:163-170) into dependency levels of fused
// "macro-ops"; each level is executed by a handful of kernels (gp_kernels.cu) that read
// the tables below. PLV operands are resolved to device pointers at compile time.
#pragma once

#include <cstdint>

namespace bito_gp {

constexpr int kTile = 256;          // patterns per thread block (one pattern per thread)
constexpr int kItemChunk = 64;      // transition matrices staged in shared memory at a time
constexpr int kMaxEigenGroups = 4;  // distinct eigenvalues of the substitution model
constexpr int kClusterThreads = 256;  // k_opt_cluster block size = patterns per rho row
constexpr int kMaxOptCluster = 16;     // largest (non-portable) thread-block cluster on sm_100
constexpr size_t kOptClusterMaxSharedBytes = 200 * 1024;  // rho rows of one block of a cluster
constexpr int kOptPatternsPerThread = 8;  // k_opt_eval_ratio: patterns one thread folds into one log

// How a PLV operand is stored in HBM.
enum PlvKind : int32_t {
  kPlvDense = 0,    // pattern_stride x 4 doubles, 32 B per pattern
  kPlvSymbols = 1,  // leaf P-PLV kept as one byte per pattern (0..3 one-hot, 4 = all ones)
  kPlvZero = 2      // never written since the last ZeroPLV / construction: reads as 0
};

struct PlvRef {
  const void* ptr;
  int32_t kind;
  int32_t id;  // logical PLV id (indexes rescaling counts)
};

// Device error bits; each mirrors an Assert/Failwith on the reference path.
enum StatusBits : uint32_t {
  kErrRescalingDifference = 1u,  // gp_engine.cpp:237-238
  kErrMultiplyNotFinite = 2u,    // gp_engine.cpp:283, 575-577
  kErrNegativePLV = 4u,          // gp_engine.cpp:585-586
  kErrRescaledStationary = 8u,   // gp_engine.cpp:256-257
  kErrEmptyPrep = 16u,           // gp_engine.cpp:325
  kErrQuartetRescaled = 32u,     // gp_engine.cpp:750-753
  kErrPeerTimeout = 64u          // a peer GPU never delivered its share of an all-reduce
};

// Per-engine device pointers handed to every kernel by value.
struct DeviceState {
  int64_t P;         // local pattern count
  int64_t P_stride;  // allocation stride in patterns (multiple of 8)
  int32_t* counts;   // rescaling count per logical PLV id          (gp_engine.hpp:317)
  double* q;         // SBN parameters per edge, linear space       (gp_engine.hpp:337)
  double* bl;        // branch lengths per edge                     (dag_branch_handler.hpp:249)
  double* diff;      // last branch-length change per edge          (dag_branch_handler.hpp:252)
  double* hybrid;    // hybrid marginal log-likelihoods per edge    (gp_engine.hpp:352)
  double* inverted;  // inverted SBN prior per edge                 (gp_engine.hpp:338)
  double* uncond;    // unconditional node probabilities per node   (gp_engine.hpp:315)
  double* ll_sum;    // per-edge sum_p w_p * log_likelihoods_(e,p), GLOBAL over ranks
  double* weights;   // site pattern weights (local shard)
  double* log_marg;  // per-pattern log marginal (local shard)      (gp_engine.hpp:349)
  double* marg_sum;  // [0] = sum_p w_p * log_marg[p], GLOBAL over ranks
  uint32_t* status;  // StatusBits
  double thr;        // rescaling threshold
  double log_thr;
  double total_weight;  // sum of ALL ranks' weights
  unsigned long long* feval_total;  // objective evaluations of finished optimisations
};

// Substitution-model eigensystem (JC69 today: substitution_model.cpp:20-26) in constant
// memory. `group` maps each eigenvalue to its distinct-eigenvalue group for the
// branch-length objective: L_p(t) = sum_g coef[p][g] * exp(group_lambda[g] * t).
struct ModelConst {
  double V[16];     // eigenvectors, row-major
  double Vinv[16];  // inverse eigenvectors, row-major
  double lambda[4];
  double pi[4];
  int32_t group[4];
  int32_t n_groups;
  double group_lambda[kMaxEigenGroups];
  int32_t is_jc69;  // the eigensystem is bit for bit JC69Model's literals: ratio_coefficients takes its exact short form
  int32_t pad;
};

// ---- macro-ops ---------------------------------------------------------------------
enum CountMode : int32_t {
  kCountKeep = 0,  // dest count unchanged
  kCountZero = 1,  // a folded ZeroPLV: count[dest] = 0
  kCountPrep = 2   // a folded PrepForMarginalization: count[dest] = min count[src_vector]
};

// [ZeroPLV] [PrepForMarginalization] IncrementWithWeightedEvolvedPLV x n into one dest
// (gp_engine.cpp:213-216, 323-333, 229-249): dest is read at most once, written once.
struct AccumGroup {
  double* dest;
  int32_t dest_id;
  int32_t init_zero;
  int32_t count_mode;
  int32_t n_items;
  int32_t item_off;
  int32_t prep_off;
  int32_t prep_len;
  int32_t pad;
};
struct AccumItem {
  PlvRef src;
  int32_t edge;
  int32_t pad;
};

// Multiply (gp_engine.cpp:278-285); max_slot indexes the level's per-PLV maxima.
struct MultOp {
  double* dest;
  PlvRef s1, s2;
  int32_t dest_id;
  int32_t max_slot;
};

// A fused node macro-op: up to two accumulate groups followed by up to two Multiplies that
// consume their results from registers. Rootward GPDAG node: PHatRight, PHatLeft, P = PHatRight o
// PHatLeft (gp_dag.cpp:278-294); leafward node: RHat, RRight = RHat o PHatLeft, RLeft = RHat o
// PHatRight (gp_dag.cpp:260-276). Unfused groups and Multiplies are nodes with one member.
struct NodeMult {
  double* dest;
  PlvRef s1, s2;
  int32_t dest_id;
  int32_t max_slot;
  int32_t s1_group;  // >= 0: operand is the result of that group of this node (registers)
  int32_t s2_group;
};
struct NodeOp {
  int32_t n_groups;  // 0..2
  int32_t n_mults;   // 0..2
  AccumGroup g[2];   // items of g[1] follow those of g[0] in the item table
  NodeMult m[2];
};

// Likelihood (gp_engine.cpp:287-291).
struct LikOp {
  PlvRef parent, child;
  double* row;  // may be null with BITO_GP_FLAG_NO_LOGLIK_MATRIX
  int32_t edge;
  int32_t pad;
};

// [ResetMarginalLikelihood] IncrementMarginalLikelihood x n (gp_engine.cpp:251-276).
struct MargItem {
  PlvRef stationary, p;
  double* row;
  int32_t edge;
  int32_t pad;
};

struct StatOp {  // SetToStationaryDistribution (gp_engine.cpp:218-227)
  double* dest;
  int32_t dest_id;
  int32_t edge;
};

struct ZeroOp {  // a ZeroPLV that could not be folded away
  double* dest;
  int32_t dest_id;
  int32_t pad;
};

enum ScalarKind : int32_t {
  kScalarCountZero = 0,  // ZeroPLV of a PLV that owns no memory: count only
  kScalarPrep = 1,       // stand-alone PrepForMarginalization
  kScalarSbn = 2,        // UpdateSBNProbabilities (gp_engine.cpp:304-321)
  kScalarCountSum = 3    // Multiply whose result is identically zero: count[a] = count[b] + count[c]
};
struct ScalarOp {
  int32_t kind;
  int32_t a, b;  // count-zero/prep: a = dest id; sbn: [a, b); count-sum: dest, src1 (src2 in vec_off)
  int32_t vec_off, vec_len;
  int32_t pad;
};

// OptimizeBranchLength (gp_engine.cpp:293-295, 667-670; dag_branch_handler.cpp:123-280).
struct OptOp {
  PlvRef parent, child;  // rootward_ (r-PLV of the parent), leafward_ (p-PLV of the child)
  int32_t edge;
  // pool[fix_off .. fix_off + fix_n): the program's transition-matrix slots that hold this edge
  // (2 * accumulate item, or 2 * Likelihood op + 1). The on-chip optimisers refresh exactly those
  // when they finish, instead of a rebuild of the whole table after every optimiser level.
  int32_t fix_off;
  int32_t fix_n;
  int32_t pad;
};

// One summand of a quartet hybrid marginal (gp_engine.cpp:748-808, quartet_hybrid_request.hpp):
// a choice of (rootward, sister, rotated, sorted) tips around one central edge. mats points at
// the five transition matrices rootward, sister, central, rotated, sorted (80 doubles).
struct QuartetItem {
  PlvRef rootward, sister, rotated, sorted;
  int32_t edge[5];        // rootward, sister, central, rotated, sorted
  int32_t rootward_node;  // indexes the unconditional node probabilities
};

// Resumable optimiser state, one per OptimizeBranchLength in flight. The decision logic
// of optimization.hpp:71-402 is run one objective evaluation at a time (gp_kernels.cu,
// opt_step): `phase` says which evaluation is pending.
struct OptState {
  // pending evaluation
  double t_eval;    // branch length at which the objective is being evaluated
  double x_eval;    // the optimiser's own coordinate (log t for Brent/Newton, t for GA)
  double e[kMaxEigenGroups];  // exp(group_lambda[g] * t_eval), set with every request
  double x_ratio;             // e[1] / e[0]
  int32_t phase;
  int32_t done;
  int32_t method;
  int32_t edge;
  double ll_offset;  // (count[parent] + count[child]) * log thr * total_weight
  // Brent (names as in optimization.hpp:75-83)
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, min, max;
  double cur_x, cur_f;  // starting point and its objective value
  double u_alt;         // gradient-step candidate of BrentMinimizeWithGradients
  int64_t count;        // remaining iterations
  int64_t iter;
  int32_t evals;
  int32_t speculative;  // a copy advanced with made-up objective values to learn the next request: writes nothing
};

// ---- Taylor-model Brent (plain Brent on a two-eigenvalue model, ratio form) ---------------------
// The objective of an edge is S(x) = sum_p w_p log(1 + rho_p x), x = e^{(l1 - l0) t}. A streamed pass
// over rho (8 B per pattern) evaluates S at up to kOptPoints points AND the power sums
//   M_j = sum_p w_p z_p^j,  z_p = rho_p / (1 + rho_p c),  j = 1..kOptMoments,
// about one of them (c). Since log(1 + rho x) = log(1 + rho c) + log(1 + z (x - c)),
//   S(x) = S(c) + sum_{j < J} (-1)^{j+1} M_j (x - c)^j / j + R,   |R| <= 2 M_J |x - c|^J / J
// whenever max_p |z_p (x - c)| <= 1/2 (J = kOptMoments is even, so M_J >= 0 bounds every |z_p|).
// Later requests of the optimiser that fall inside the radius where that bound is below a quarter
// ulp of the objective are answered from the model without touching HBM (OptPass, k_opt_step_model).
constexpr int kOptPoints = 4;
constexpr int kOptMoments = 12;
constexpr int kOptPassValues = kOptPoints + kOptMoments;

// One entry of the pass list: what k_opt_eval_model evaluates for one still-active edge. Self-contained
// (no pointer chasing through the optimiser state), double-buffered by round parity.
struct OptReq {
  double x[kOptPoints];  // points ([0] only after the first pass of a search)
  double c;              // centre of the power sums
  int32_t o;             // edge index within the chunk; < 0: nothing to do
  int32_t pad;
};

// First tile group of each weight class of the streamed layout (Engine::BuildWeightClasses): class c
// (weight c + 1; 7 = general weights) owns tile groups [start[c], start[c + 1]).
struct OptClassStarts {
  int32_t start[9];
};

struct OptPass {
  // what the next streamed pass evaluates
  double px[kOptPoints];  // x of each point; [0] is the optimiser's pending request
  double ps[kOptPoints];  // the optimiser's own coordinate of each point (log t)
  double pt[kOptPoints];  // t of each point
  int32_t n_pts;
  int32_t centre;         // index of the point the moments are taken about
  // evaluated points the optimiser has not asked for (yet): speculative first requests
  double cache_s[kOptPoints], cache_ll[kOptPoints];
  int32_t n_cache;
  int32_t passes;         // streamed passes this search has cost so far
  // the model
  double c, S_c, radius;  // radius = largest |x - c|^J the model answers for (0: no model)
  double mj[kOptMoments]; // (-1)^{j+1} M_j / j, j = 1..J-1 at [j-1]; [J-1] = M_J
};

struct OptParams {
  int32_t significant_digits;
  int32_t check_convergence;  // !IsFirstOptimization()
  int64_t max_iter;
  double min_log_bl, max_log_bl;
  double denominator_tolerance;
  double step_size, log_step_size;
  double diff_threshold;
  // derived on the host so that the optimiser's single-thread step does not spend time on them
  double brent_tolerance;    // ldexp(1, 1 - significant_digits), optimization.hpp:84
  double decimal_tolerance;  // pow(10, -significant_digits), optimization.hpp:336, 352, 372
};

// What a captured OptimizeBranchLength launch must not bake in: the optimiser settings can change
// between replays of the same op list (method, significant digits, first-vs-later optimisation,
// dag_branch_handler.hpp:49-52), so k_opt_block reads them from device memory.
struct OptControl {
  OptParams prm;
  int32_t method;
  int32_t n_derivatives;
};

// Peer-memory all-reduce over NVLink (k_peer_allreduce): every rank owns one exchange buffer, mapped
// into every other rank's address space through CUDA IPC. Layout of a buffer, in doubles:
//   data[parity][source rank][capacity], then (as 64-bit words) flag[parity][source rank].
constexpr int kMaxPeerRanks = 8;
constexpr int64_t kPeerCapacity = 16384;  // doubles per all-reduce (the batched sweep of the bench DAG sums 11 139 per round); larger ones go through NCCL
struct PeerComm {
  double* base[kMaxPeerRanks];  // base[r] = rank r's exchange buffer as mapped in THIS process
  unsigned long long* epoch;    // all-reduces completed so far (device memory, so graphs replay)
  uint32_t* status;             // DeviceState::status
  int32_t n_ranks, rank;
};
// ... then, for the cluster-resident optimiser on several GPUs (k_opt_cluster*), kPeerEdgeSlots edge
// records [slot][bank][source rank] of {kPeerEdgeValues values, tag}: the sums of one pass of one edge
// (k_opt_cluster uses value 0 only; k_opt_cluster_model all of them: 4 objective sums, 12 power sums, K_e).
constexpr int kPeerEdgeSlots = 64;
constexpr int kPeerEdgeValues = 17;
constexpr int kPeerEdgeRecord = kPeerEdgeValues + 1;  // doubles per record (the tag last)
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t PeerEdgeOffsetDoubles(int n_ranks) {
  return size_t(2) * n_ranks * kPeerCapacity + size_t(2) * n_ranks;
}
inline size_t PeerBufferBytes(int n_ranks) {
  return (PeerEdgeOffsetDoubles(n_ranks) + size_t(kPeerEdgeSlots) * 4 * n_ranks * kPeerEdgeRecord) * sizeof(double);
}
// Cross-GPU part of k_opt_cluster's sums: enabled = 0 on a single rank.
struct PeerEdge {
  const PeerComm* pc;       // in device memory: lanes index base[] by rank
  unsigned long long* seq;  // [kPeerEdgeSlots] searches run so far per edge slot (device memory)
  int32_t enabled;
  int32_t pad;
};

// The transition-matrix tables of the running program and the slot pool OptOp::fix_off indexes.
struct OptRefresh {
  const int32_t* pool;
  double* mtab;      // 16 doubles per accumulate item: q[e] * M(t_e)
  double* mtab_lik;  // 16 doubles per Likelihood op: M(t_e)
};

// k_opt_cluster: rho rows (kClusterThreads patterns each, one weight class per row) of an edge are
// dealt to the blocks of its cluster, rows_per_block consecutive rows each.
struct OptClusterLayout {
  int32_t class_row_start[9];  // rows of class c: [class_row_start[c], class_row_start[c + 1])
  int32_t rows_total;
  int32_t rows_per_block;
};
struct OptClusterPlan {
  int threads = 0;          // threads per block: 256 or 1024
  int cluster_size = 0;     // blocks per cluster (0: the cluster path is not available)
  int rows_per_block = 0;
  int rows_total = 0;
  size_t shared_bytes = 0;
  int active_clusters = 0;  // edges resident on the chip at once (cudaOccupancyMaxActiveClusters)
};

}  // namespace bito_gp
