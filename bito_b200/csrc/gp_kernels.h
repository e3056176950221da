// bito_b200/csrc/gp_kernels.h — launch wrappers of the sm_100a GP kernels (gp_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "gp_types.h"

namespace bito_gp {

// Number of pattern tiles for P local patterns.
// Number of pattern tiles for P local patterns. An empty shard (P = 0, a rank of a multi-GPU run
// that owns no pattern) still gets one tile, so the per-macro-op bookkeeping (rescaling counts,
// optimiser states) runs there too; no thread of that tile is live.
inline int64_t TilesFor(int64_t P) { return P > 0 ? (P + kTile - 1) / kTile : 1; }

cudaError_t UploadModel(const ModelConst& model);

// q[e] * M(t_e) per accumulate item into mtab[16 * item], M(t_e) per Likelihood into
// mtab_lik[16 * lik]: once per program execution and after every level that changes t or q.
void LaunchBuildMatrices(cudaStream_t s, const DeviceState& st, const AccumItem* items, int n_items,
                         const LikOp* liks, int n_liks, double* mtab, double* mtab_lik);
void LaunchNodes(cudaStream_t s, const DeviceState& st, const NodeOp* nodes, const AccumItem* items,
                 const int32_t* pool, const double* mtab, int n_nodes, double* level_max);
// level_max[o] = maximum entry of ops[o].dest over all patterns (and ranks).
void LaunchRescale(cudaStream_t s, const DeviceState& st, const MultOp* ops, int n_ops,
                   const double* level_max);
// max over taxa x P symbols (rows of stride P_stride bytes) as a double into *out.
void LaunchMaxSymbol(cudaStream_t s, const uint8_t* symbols, int64_t rows, int64_t P, int64_t P_stride,
                     double* out);
// partials: n_ops rows of LikelihoodTileGroups(n_ops, P) tile-group sums.
int64_t LikelihoodTileGroups(int n_ops, int64_t P);
void LaunchLikelihood(cudaStream_t s, const DeviceState& st, const LikOp* ops, int n_ops,
                      const double* mtab, double* partials);
void LaunchMarginal(cudaStream_t s, const DeviceState& st, const MargItem* items, int n_items,
                    int reset, double* partials /* (n_items + 1) x tiles */);
void LaunchStationary(cudaStream_t s, const DeviceState& st, const StatOp* ops, int n_ops);
void LaunchZero(cudaStream_t s, const DeviceState& st, const ZeroOp* ops, int n_ops);
void LaunchScalar(cudaStream_t s, const DeviceState& st, const ScalarOp* ops,
                  const int32_t* pool, int n_ops);
// out[i] = sum_t partials[i * tiles + t] (fixed order); when scatter_idx != nullptr also
// scatter_dst[scatter_idx[i]] = out[i] for scatter_idx[i] >= 0.
void LaunchReducePartials(cudaStream_t s, const double* partials, int n_out, int64_t tiles,
                          double* out, const int32_t* scatter_idx, double* scatter_dst);
void LaunchScatter(cudaStream_t s, const double* packed, int n, const int32_t* scatter_idx,
                   double* scatter_dst);

// Branch-length optimisation.
void LaunchOptPrepare(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                      OptState* states, const OptParams& params, int method, double* coef,
                      int init_states);
void LaunchOptEval(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                   const double* coef, int n_derivatives, double* partials /* n_ops*3 x tiles */,
                   int n_groups);
// Value k of edge o is sums[o * value_stride + k] (multi-rank, after the all-reduce), or the sum of
// the n_parts entries of row (o * value_stride + k) of `partials`, reduced inside the step in a
// fixed order (single rank). edge_const (ratio form only): K_e, see k_opt_prepare_ratio.
void LaunchOptStep(cudaStream_t s, const DeviceState& st, int n_ops, OptState* states,
                   const OptParams& params, const double* sums, const double* partials, int n_parts,
                   int n_values, int value_stride, const double* edge_const,
                   int32_t* active_counter, int32_t* active, int active_capacity, int parity);
// Two-eigenvalue (JC69) Brent path: rho = c1 / c0 per pattern (8 B) stored at perm[p] (weight-class
// order, Engine::BuildWeightClasses), K_e partials per tile.
void LaunchOptPrepareRatio(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                           OptState* states, const OptParams& params, int method, double* rho,
                           const int32_t* perm,
                           const int32_t* pos_w /* per pattern: rho position << 3 | weight (1..7, 0 = look it up) */,
                           int64_t rho_stride,
                           double* partials /* n_ops x OptPrepareTileGroups(n_ops, P) */,
                           int32_t* active, int active_capacity);
// Small alignments: one block per edge runs the whole 1-D search on chip (coefficients in
// OptBlockSharedBytes(P, G) of dynamic shared memory), one launch per level, no host round trip.
size_t OptBlockSharedBytes(int64_t P, int n_groups);
// The optimiser settings live in device memory (*ctl, written in-stream by LaunchSetOptControl
// before the program runs) so that a captured launch picks up the settings of each replay.
void LaunchSetOptControl(cudaStream_t s, OptControl* ctl, const OptControl& value);
void LaunchOptBlock(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                    const OptControl* ctl, int n_groups, const OptRefresh& refresh);
// Large alignments (plain Brent, two-eigenvalue model, single rank): one thread-block cluster per
// edge keeps rho in distributed shared memory and runs the whole search there (k_opt_cluster).
// PlanOptCluster describes one cluster shape (threads per block 256 | 1024, cluster_size blocks) for
// rows_total rho rows of kClusterThreads patterns; false if the device cannot run it.
bool PlanOptCluster(int64_t rows_total, int threads, int cluster_size, OptClusterPlan* plan, bool model);
cudaError_t LaunchOptCluster(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                             const OptControl* ctl, const int32_t* inv_perm, const double* wperm,
                             const int32_t class_row_start[9], const OptClusterPlan& plan,
                             const OptRefresh& refresh, const PeerEdge& peer, const double* rho_in = nullptr,
                             int64_t rho_stride = 0, const double* edge_const_in = nullptr,
                             bool model = false /* k_opt_cluster_model: the Taylor-model search */,
                             double min_weight = 1.);
// Pipelined cluster scheme: rho (cluster layout, position cpos[p]) and K_e tile-group partials
// (n_ops x OptPrepareTileGroups) of a chunk of edges, written for k_opt_cluster<T, true>.
// max_blocks > 0: a fixed grid of that many blocks walks the (edge, tile group) items.
void LaunchOptPrepareCluster(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops, double* rho,
                             const int32_t* cpos, int64_t rho_stride, double* partials, int max_blocks);
int64_t OptPrepareTileGroups(int n_ops, int64_t P);
int64_t OptRatioTileGroups(int64_t P);
int64_t OptRatioPartials(int64_t P);  // partial sums per edge written by LaunchOptEvalRatio
void LaunchOptEvalRatio(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                        const double* rho, int64_t rho_stride, const double* wperm,
                        const uint8_t* row_class, double* partials /* n_ops x OptRatioPartials */,
                        int32_t* active /* [4 + 2 * capacity]: counts by parity, then the two lists */,
                        int active_capacity, int parity);

// Taylor-model Brent (gp_types.h, OptPass): the first requests of every search (LaunchOptPlan), one
// streamed pass = kOptPassValues sums per still-active edge (n_points = kOptPoints on the first pass of
// a search, 1 afterwards; partials [edge][value][segment], OptModelSegments), and the step that
// answers every later request it can from the cache / the model before asking for another pass.
void LaunchOptPlan(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                   const OptParams& params, OptPass* pass, OptReq* req /* [2][capacity] */,
                   int32_t* active /* [2]: entries of req by round parity */);
void OptModelSegments(int64_t rho_stride, int* seg_len, int* n_seg);
void LaunchOptEvalModel(cudaStream_t s, int n_ops, int n_points, const OptReq* req /* this round's half */,
                        const double* rho, int64_t rho_stride, const double* wperm,
                        const OptClassStarts& classes, double* partials, int32_t* active, int parity);
void LaunchOptStepModel(cudaStream_t s, const DeviceState& st, int n_ops, OptState* states, OptPass* pass,
                        OptReq* req, const OptParams& params, const double* sums /* multi-rank: all-reduced */,
                        const double* partials /* single rank */, int n_seg, const double* edge_const,
                        double min_weight, int32_t* active, int capacity, int parity);

// Quartet hybrid marginals (gp_engine.cpp:748-808): mats = 80 doubles per summand, partials =
// n_items x TilesFor(P) weighted tile sums, out[i] = non-sequence log-probability + sums[i].
void LaunchQuartetMatrices(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                           double* mats);
void LaunchQuartet(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                   const double* mats, double* partials);
void LaunchQuartetFinish(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                         const double* sums, double* out);

// In-place all-reduce (sum or max, fixed rank order) of n <= kPeerCapacity doubles across the ranks
// of `pc` by peer-memory stores over NVLink: one launch, no NCCL call.
void LaunchPeerAllReduce(cudaStream_t s, const PeerComm& pc, double* buf, int n, bool max_op);

// Utilities.
void LaunchExportPlv(cudaStream_t s, const DeviceState& st, PlvRef src, double* dense_out);
void LaunchFill(cudaStream_t s, double* dst, int64_t n, double value);
void LaunchTransitionMatrix(cudaStream_t s, double t, double* out16);
void LaunchWeightedSum(cudaStream_t s, const DeviceState& st, const double* values,
                       double* partials);

}  // namespace bito_gp
