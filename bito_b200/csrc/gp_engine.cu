// bito_b200/csrc/gp_engine.cu — host engine: HBM-resident state, op-list compiler
// (fusion + dependency levels), level-by-level / CUDA-graph execution, NCCL scalars.
#include "gp_engine.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>

#include "gp_kernels.h"

namespace bito_gp {

namespace {
thread_local std::string g_last_error;

#define GP_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t err__ = (expr);                                                        \
    if (err__ != cudaSuccess) {                                                        \
      std::ostringstream os__;                                                         \
      os__ << "CUDA error: " << cudaGetErrorString(err__) << " at " << __FILE__ << ":" \
           << __LINE__ << " (" #expr ")";                                              \
      throw GpError(os__.str());                                                       \
    }                                                                                  \
  } while (0)

[[noreturn]] void Fail(const std::string& msg) { throw GpError(msg); }

inline int64_t RoundUp(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// dag_branch_handler.hpp:266-295
constexpr double kDefaultBranchLength = 0.1;
constexpr double kMinLogBranchLength = -13.9;
constexpr double kMaxLogBranchLength = 1.1;
constexpr double kDenominatorToleranceForNewton = 1e-10;
constexpr double kStepSizeForOptimization = 5e-4;
constexpr double kStepSizeForLogSpaceOptimization = 1.0005;
constexpr int64_t kMaxIterForOptimization = 1000;
constexpr double kBranchLengthDifferenceThreshold = 1e-15;
// k_opt_block keeps G doubles per pattern in shared memory: up to 10240 patterns under JC69 (G = 2)
constexpr size_t kOptBlockMaxSharedBytes = 160 * 1024;

// Two independent 64-bit hashes of an op list in ONE pass over it: `key` keys the program cache, `check` is
// kept in the Program and compared on every hit, so a collision of the key alone cannot run another
// list. Multiply-xorshift over 8-byte words (the inputs are int64 tables) in four interleaved lanes per
// hash: a single dependent chain costs ~1 ms per hash on the bench's 84 000-op list (4 MB) - host time
// during which the GPU sits idle at the head of every call - eight independent chains stream it.
struct OpsHash {
  uint64_t key, check;
};
OpsHash HashOps(const bito_gp_op* ops, int64_t n, const int64_t* vec, int64_t vec_len) {
  constexpr uint64_t kMulA = 1099511628211ull, kMulB = 0xff51afd7ed558ccdull;
  uint64_t a[4] = {1469598103934665603ull, 0x9ae16a3b2f90404full, 0xc3a5c85c97cb3127ull, 0xb492b66fbe98f273ull};
  uint64_t b[4] = {0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull, 0x94d049bb133111ebull, 0x2545f4914f6cdd1dull};
  auto word = [&](int lane, uint64_t w) {
    a[lane] = (a[lane] ^ w) * kMulA;
    a[lane] ^= a[lane] >> 29;
    b[lane] = (b[lane] ^ w) * kMulB;
    b[lane] ^= b[lane] >> 33;
  };
  auto mix = [&](const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    const size_t words = bytes / 8;
    size_t i = 0;
    for (; i + 4 <= words; i += 4) {
      uint64_t w[4];
      std::memcpy(w, p + 8 * i, 32);
      word(0, w[0]);
      word(1, w[1]);
      word(2, w[2]);
      word(3, w[3]);
    }
    for (; i < words; ++i) {  // the tail goes through lane (i mod 4): position still matters
      uint64_t w;
      std::memcpy(&w, p + 8 * i, 8);
      word(static_cast<int>(i & 3), w);
    }
  };
  const int64_t head[2] = {n, vec_len};
  mix(head, sizeof head);
  mix(ops, static_cast<size_t>(n) * sizeof(bito_gp_op));
  if (vec_len > 0) mix(vec, static_cast<size_t>(vec_len) * sizeof(int64_t));
  OpsHash h = {a[0], b[0]};
  for (int lane = 1; lane < 4; ++lane) {
    h.key = (h.key ^ a[lane]) * kMulA;
    h.key ^= h.key >> 29;
    h.check = (h.check ^ b[lane]) * kMulB;
    h.check ^= h.check >> 33;
  }
  return h;
}

// ---- NCCL through dlopen: the library loads and runs single-GPU without NCCL ----------
struct NcclUniqueId {
  char internal[128];
};
using NcclComm = void*;
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool loaded = false;
};
NcclApi& Nccl() {
  static NcclApi api;
  if (api.loaded) return api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) Fail(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.GetErrorString =
      reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy)
    Fail("libnccl.so.2 lacks the expected symbols");
  api.loaded = true;
  return api;
}
void NcclCheck(int rc, const char* what) {
  if (rc != 0) {
    auto& api = Nccl();
    Fail(std::string("NCCL error in ") + what + ": " +
         (api.GetErrorString ? api.GetErrorString(rc) : "?"));
  }
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclInt8 = 0;

}  // namespace

void SetLastError(const std::string& msg) { g_last_error = msg; }
const char* LastError() { return g_last_error.c_str(); }

void MakeNcclUniqueId(uint8_t id[128]) {
  NcclUniqueId uid;
  NcclCheck(Nccl().GetUniqueId(&uid), "ncclGetUniqueId");
  std::memcpy(id, uid.internal, 128);
}

// ---- SlabPool / DeviceArray -----------------------------------------------------------
void SlabPool::Init(size_t slot_bytes, size_t target_chunk_bytes) {
  slot_bytes_ = RoundUp(static_cast<int64_t>(slot_bytes), 256);
  slots_per_chunk_ = std::max<size_t>(1, target_chunk_bytes / slot_bytes_);
  next_in_chunk_ = slots_per_chunk_;
}
void* SlabPool::Alloc() {
  handed_out_++;
  if (!free_.empty()) {
    void* p = free_.back();
    free_.pop_back();
    return p;
  }
  if (next_in_chunk_ == slots_per_chunk_) {
    char* c = nullptr;
    GP_CUDA(cudaMalloc(&c, slots_per_chunk_ * slot_bytes_));
    chunks_.push_back(c);
    next_in_chunk_ = 0;
  }
  return chunks_.back() + (next_in_chunk_++) * slot_bytes_;
}
void SlabPool::Release() {
  for (char* c : chunks_) cudaFree(c);
  chunks_.clear();
  free_.clear();
  next_in_chunk_ = slots_per_chunk_;
  handed_out_ = 0;
}

template <typename T>
void DeviceArray<T>::Resize(size_t count, bool keep, cudaStream_t stream) {
  if (count <= n) return;
  T* fresh = nullptr;
  GP_CUDA(cudaMalloc(&fresh, count * sizeof(T)));
  if (keep && ptr != nullptr && n > 0) {
    GP_CUDA(cudaMemcpyAsync(fresh, ptr, n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    GP_CUDA(cudaStreamSynchronize(stream));
  }
  if (ptr != nullptr) cudaFree(ptr);
  ptr = fresh;
  n = count;
}
template <typename T>
void DeviceArray<T>::Release() {
  if (ptr != nullptr) cudaFree(ptr);
  ptr = nullptr;
  n = 0;
}

// ---- construction: GPEngine::GPEngine, gp_engine.cpp:9-43 ------------------------------
Engine::Engine(const bito_gp_config& cfg) : cfg_(cfg) {
  if (cfg.abi_version != BITO_GP_ABI_VERSION) Fail("bito_gp_config.abi_version mismatch");
  if (cfg.taxon_count <= 0 || cfg.pattern_count < 0 || cfg.node_count <= 0 || cfg.gpcsp_count <= 0)
    Fail("bito_gp_create: taxon, node and gpcsp counts must be positive and pattern_count >= 0");
  if (!(cfg.rescaling_threshold > 0.) || !(cfg.rescaling_threshold < 1.))
    Fail("bito_gp_create: rescaling_threshold must lie in (0, 1)");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    Fail("bito_gp_create: no CUDA device is visible; this engine has no CPU fallback");
  if (cfg.device < 0 || cfg.device >= n_dev) Fail("bito_gp_create: bad device ordinal");
  device_ = cfg.device;
  GP_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  GP_CUDA(cudaGetDeviceProperties(&prop, device_));
  // the library carries sm_100a SASS only (no PTX): any other device would fail at its first launch
  if (prop.major != 10 || prop.minor != 0)
    Fail("bito_gp_create: kernels are built for sm_100a (B200) only; found sm_" +
         std::to_string(prop.major) + std::to_string(prop.minor));
  GP_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
  stream_ = own_stream_;
  GP_CUDA(cudaEventCreate(&ev_begin_));
  GP_CUDA(cudaEventCreate(&ev_end_));
  GP_CUDA(cudaMallocHost(&pinned_, 4096));

  taxon_count_ = cfg.taxon_count;
  P_ = cfg.pattern_count;
  P_stride_ = std::max<int64_t>(8, RoundUp(P_, 8));  // an empty shard still owns (unused) storage
  site_count_ = cfg.site_count;
  node_count_ = cfg.node_count;
  gpcsp_count_ = cfg.gpcsp_count;
  spare_nodes_ = cfg.spare_node_count > 0 ? cfg.spare_node_count : 16;
  spare_gpcsps_ = cfg.spare_gpcsp_count > 0 ? cfg.spare_gpcsp_count : 3;
  thr_ = cfg.rescaling_threshold;
  method_ = cfg.use_gradients ? BITO_GP_BRENT_OPTIMIZATION_WITH_GRADIENTS
                              : BITO_GP_BRENT_OPTIMIZATION;  // gp_engine.cpp:660-665
  if (taxon_count_ > node_count_) Fail("bito_gp_create: taxon_count exceeds node_count");

  size_t free_b = 0, total_b = 0;
  GP_CUDA(cudaMemGetInfo(&free_b, &total_b));
  max_device_bytes_ = cfg.max_device_bytes > 0 ? cfg.max_device_bytes
                                               : static_cast<int64_t>(free_b * 0.9);

  // JC69 eigensystem, substitution_model.cpp:20-26.
  {
    const double V[16] = {1.0, 2.0, 0.0, 0.5, 1.0, -2.0, 0.5, 0.0,
                          1.0, 2.0, 0.0, -0.5, 1.0, -2.0, -0.5, 0.0};
    const double Vinv[16] = {0.25, 0.25, 0.25, 0.25, 0.125, -0.125, 0.125, -0.125,
                             0.0, 1.0, 0.0, -1.0, 1.0, 0.0, -1.0, 0.0};
    const double lambda[4] = {0.0, -1.3333333333333333, -1.3333333333333333,
                              -1.3333333333333333};
    const double pi[4] = {0.25, 0.25, 0.25, 0.25};
    InstallModel(V, Vinv, lambda, pi);
  }
  if (const char* env = getenv("BITO_GP_OPT_CHUNK_MB")) opt_chunk_bytes_ = int64_t(atoll(env)) << 20;
  // BITO_GP_OPT_CLUSTER: 0 = never use the cluster-resident optimiser; N > 0 = use it with clusters
  // of exactly N blocks even where one block would do (tests); unset = automatic
  if (const char* env = getenv("BITO_GP_OPT_CLUSTER")) opt_cluster_env_ = atoi(env);
  // BITO_GP_OPT_CLUSTER_THREADS: 256 | 1024 = only that block size (with a forced cluster size: tests)
  if (const char* env = getenv("BITO_GP_OPT_CLUSTER_THREADS")) opt_cluster_threads_env_ = atoi(env);
  // BITO_GP_OPT_SCHEME: 0 = rounds, 2 = one cluster per edge reading the PLVs, 3 = pipelined clusters
  // (a streaming producer writes rho, clusters run the searches); unset = the engine's own choice
  if (const char* env = getenv("BITO_GP_OPT_SCHEME")) opt_scheme_env_ = atoi(env);
  if (const char* env = getenv("BITO_GP_OPT_MODEL")) opt_model_env_ = atoi(env);
  if (const char* env = getenv("BITO_GP_OPT_FIRST_CHECK")) opt_model_first_check_ = std::max(1, atoi(env));
  if (const char* env = getenv("BITO_GP_OPT_RING_EDGES")) opt_ring_edges_env_ = atoi(env);
  if (const char* env = getenv("BITO_GP_PREP_BLOCKS_PER_SM")) prep_blocks_per_sm_ = atoi(env);
  if (const char* env = getenv("BITO_GP_OPT_PRIORITY")) opt_priority_env_ = atoi(env);
  {
    // the producer runs at the lowest stream priority and the consumer clusters at the highest, so that a
    // cluster whose edge is ready is placed before more producer blocks are
    int least = 0, greatest = 0;
    GP_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    GP_CUDA(cudaStreamCreateWithPriority(&prep_stream_, cudaStreamNonBlocking, least));
    GP_CUDA(cudaStreamCreateWithPriority(&cons_stream_, cudaStreamNonBlocking, greatest));
  }
  GP_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
  GP_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  for (int b = 0; b < 2; ++b) {
    GP_CUDA(cudaEventCreateWithFlags(&ev_ready_[b], cudaEventDisableTiming));
    GP_CUDA(cudaEventCreateWithFlags(&ev_free_[b], cudaEventDisableTiming));
  }

  // PLV slabs: ~256 MiB chunks (or one PLV, whichever is larger).
  plv_pool_.Init(static_cast<size_t>(32 * P_stride_), size_t(256) << 20);
  row_pool_.Init(static_cast<size_t>(8 * P_stride_), size_t(64) << 20);
  plvs_.assign(static_cast<size_t>(padded_plv_count()), PlvSlot{});
  rows_.assign(static_cast<size_t>(padded_gpcsp_count()), nullptr);

  d_symbols_.Resize(static_cast<size_t>(taxon_count_ * P_stride_), false, stream_);
  d_weights_.Resize(static_cast<size_t>(P_stride_), false, stream_);
  d_log_marg_.Resize(static_cast<size_t>(P_stride_), false, stream_);
  d_counts_.Resize(static_cast<size_t>(padded_plv_count()), false, stream_);
  GP_CUDA(cudaMemsetAsync(d_counts_.ptr, 0, d_counts_.n * sizeof(int32_t), stream_));
  GP_CUDA(cudaMemsetAsync(d_weights_.ptr, 0, d_weights_.n * sizeof(double), stream_));
  LaunchFill(stream_, d_log_marg_.ptr, P_stride_, -std::numeric_limits<double>::infinity());
  d_status_.Resize(1, false, stream_);
  GP_CUDA(cudaMemsetAsync(d_status_.ptr, 0, sizeof(uint32_t), stream_));
  d_feval_total_.Resize(2, false, stream_);  // [0] objective evaluations, [1] streamed passes over rho
  GP_CUDA(cudaMemsetAsync(d_feval_total_.ptr, 0, 2 * sizeof(unsigned long long), stream_));
  d_active_.Resize(1, false, stream_);
  AllocEdgeArrays(padded_gpcsp_count());
  GP_CUDA(cudaStreamSynchronize(stream_));
}

Engine::~Engine() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  for (auto& kv : programs_) FreeProgram(*kv.second);
  ReleasePeerMemory();
  d_peer_epoch_.Release();
  d_peer_seq_.Release();
  d_peer_comm_.Release();
  if (nccl_comm_ != nullptr) Nccl().CommDestroy(nccl_comm_);
  plv_pool_.Release();
  row_pool_.Release();
  d_symbols_.Release(); d_weights_.Release(); d_log_marg_.Release(); d_counts_.Release();
  d_q_.Release(); d_bl_.Release(); d_diff_.Release(); d_hybrid_.Release(); d_ll_sum_.Release();
  d_inverted_.Release(); d_uncond_.Release(); d_status_.Release(); d_feval_total_.Release();
  d_partials_.Release(); d_packed_.Release(); d_level_max_.Release(); d_coef_.Release(); d_opt_ctl_.Release();
  d_dense_tmp_.Release(); d_mtab_.Release(); d_mtab_lik_.Release(); d_opt_states_.Release(); d_opt_pass_.Release(); d_opt_req_.Release(); d_opt_const_.Release(); d_opt_active_.Release(); d_perm_.Release(); d_pos_w_.Release(); d_wperm_.Release(); d_row_class_.Release(); d_active_.Release(); d_single_opt_.Release(); d_cluster_inv_perm_.Release(); d_cluster_wperm_.Release(); d_quartet_items_.Release(); d_quartet_mats_.Release();
  d_cluster_pos_.Release(); d_rho_ring_.Release(); d_ring_const_.Release(); d_ring_partials_.Release();
  if (prep_stream_) cudaStreamDestroy(prep_stream_);
  if (cons_stream_) cudaStreamDestroy(cons_stream_);
  if (ev_join_) cudaEventDestroy(ev_join_);
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  for (int b = 0; b < 2; ++b) {
    if (ev_ready_[b]) cudaEventDestroy(ev_ready_[b]);
    if (ev_free_[b]) cudaEventDestroy(ev_free_[b]);
  }
  if (pinned_ != nullptr) cudaFreeHost(pinned_);
  if (ev_begin_) cudaEventDestroy(ev_begin_);
  if (ev_end_) cudaEventDestroy(ev_end_);
  if (own_stream_) cudaStreamDestroy(own_stream_);
}

void Engine::Activate() const { GP_CUDA(cudaSetDevice(device_)); }

// Per-edge arrays; new entries get the reference defaults (gp_engine.cpp:112-162,
// dag_branch_handler.cpp:8-18): q = 1, inverted prior = 1, branch length 0.1, diff 0,
// hybrid marginal -inf.
void Engine::AllocEdgeArrays(int64_t padded) {
  const size_t old_n = d_bl_.n;
  const size_t want = static_cast<size_t>(padded);
  if (want <= old_n) return;
  d_q_.Resize(want, true, stream_);
  d_inverted_.Resize(want, true, stream_);
  d_bl_.Resize(want, true, stream_);
  d_diff_.Resize(want, true, stream_);
  d_hybrid_.Resize(want, true, stream_);
  // ll_sum has one extra trailing slot: the weighted total log marginal (marg_sum).
  d_ll_sum_.Resize(want + 1, true, stream_);
  const int64_t fresh = static_cast<int64_t>(want - old_n);
  LaunchFill(stream_, d_q_.ptr + old_n, fresh, 1.0);
  LaunchFill(stream_, d_inverted_.ptr + old_n, fresh, 1.0);
  LaunchFill(stream_, d_bl_.ptr + old_n, fresh, kDefaultBranchLength);
  LaunchFill(stream_, d_diff_.ptr + old_n, fresh, 0.0);
  LaunchFill(stream_, d_hybrid_.ptr + old_n, fresh, -std::numeric_limits<double>::infinity());
  LaunchFill(stream_, d_ll_sum_.ptr + old_n, fresh + 1, 0.0);
  {
    const size_t old_u = d_uncond_.n;
    d_uncond_.Resize(static_cast<size_t>(node_count_ + spare_nodes_), true, stream_);
    if (d_uncond_.n > old_u)
      LaunchFill(stream_, d_uncond_.ptr + old_u, static_cast<int64_t>(d_uncond_.n - old_u), 1.0);
  }
}

DeviceState Engine::State() const {
  DeviceState st{};
  st.P = P_;
  st.P_stride = P_stride_;
  st.counts = d_counts_.ptr;
  st.q = d_q_.ptr;
  st.bl = d_bl_.ptr;
  st.diff = d_diff_.ptr;
  st.hybrid = d_hybrid_.ptr;
  st.inverted = d_inverted_.ptr;
  st.uncond = d_uncond_.ptr;
  st.ll_sum = d_ll_sum_.ptr;
  st.weights = d_weights_.ptr;
  st.log_marg = d_log_marg_.ptr;
  st.marg_sum = d_ll_sum_.ptr + (d_ll_sum_.n - 1);
  st.status = d_status_.ptr;
  st.thr = thr_;
  st.log_thr = std::log(thr_);
  st.total_weight = total_weight_;
  st.feval_total = d_feval_total_.ptr;
  return st;
}

void Engine::CheckPlv(int64_t id, const char* what) const {
  if (id < 0 || id >= padded_plv_count())
    Fail(std::string(what) + ": PLV index " + std::to_string(id) + " out of range [0, " +
         std::to_string(padded_plv_count()) + ")");
}
void Engine::CheckEdge(int64_t id, const char* what) const {
  if (id < 0 || id >= padded_gpcsp_count())
    Fail(std::string(what) + ": GPCSP index " + std::to_string(id) + " out of range [0, " +
         std::to_string(padded_gpcsp_count()) + ")");
}

void Engine::InvalidatePrograms() { alloc_version_++; }

// ---- substitution model ---------------------------------------------------------------------------
namespace {
std::atomic<uint64_t> g_next_model_id{1};
// which engine's eigensystem the constant-memory copy of each device holds
uint64_t g_bound_model[64] = {0};
}  // namespace

void Engine::InstallModel(const double* v, const double* vinv, const double* lambda, const double* pi) {
  ModelConst m{};
  std::memcpy(m.V, v, sizeof m.V);
  std::memcpy(m.Vinv, vinv, sizeof m.Vinv);
  std::memcpy(m.lambda, lambda, sizeof m.lambda);
  std::memcpy(m.pi, pi, sizeof m.pi);
  m.n_groups = 0;
  for (int k = 0; k < 4; ++k) {  // bit-equal eigenvalues share one exponential of the objective
    int g = -1;
    for (int j = 0; j < m.n_groups; ++j)
      if (m.group_lambda[j] == lambda[k]) g = j;
    if (g < 0) {
      g = m.n_groups++;
      m.group_lambda[g] = lambda[k];
    }
    m.group[k] = g;
  }
  {
    const double jc_v[16] = {1.0, 2.0, 0.0, 0.5, 1.0, -2.0, 0.5, 0.0, 1.0, 2.0, 0.0, -0.5, 1.0, -2.0, -0.5, 0.0};
    const double jc_vinv[16] = {0.25, 0.25, 0.25, 0.25, 0.125, -0.125, 0.125, -0.125,
                                0.0, 1.0, 0.0, -1.0, 1.0, 0.0, -1.0, 0.0};
    m.is_jc69 = std::memcmp(m.V, jc_v, sizeof jc_v) == 0 && std::memcmp(m.Vinv, jc_vinv, sizeof jc_vinv) == 0 &&
                m.n_groups == 2 && m.group[0] == 0 && m.group[1] == 1 && m.group[2] == 1 && m.group[3] == 1;
  }
  model_ = m;
  model_id_ = g_next_model_id++;
  n_eigen_groups_ = m.n_groups;
  for (int g = 0; g < kMaxEigenGroups; ++g) group_lambda_[g] = g < m.n_groups ? m.group_lambda[g] : 0.;
}

void Engine::BindModel() {
  if (device_ < 64 && g_bound_model[device_] == model_id_) return;
  // The eigensystem is one __constant__ symbol per device, shared by every engine of the process: nothing
  // of ANY engine may be in flight while it changes (a ProcessOperations call can return before its
  // graph has finished, so the engine that bound the previous model may still be running).
  GP_CUDA(cudaDeviceSynchronize());
  GP_CUDA(UploadModel(model_));
  if (device_ < 64) g_bound_model[device_] = model_id_;
}

void Engine::SetSubstitutionModel(const double* v, const double* vinv, const double* lambda, const double* pi) {
  Activate();
  for (int k = 0; k < 16; ++k)
    if (!std::isfinite(v[k]) || !std::isfinite(vinv[k])) Fail("bito_gp_set_substitution_model: eigenvectors must be finite");
  double sum = 0.;
  for (int k = 0; k < 4; ++k) {
    if (!std::isfinite(lambda[k]) || !(pi[k] > 0.))
      Fail("bito_gp_set_substitution_model: eigenvalues must be finite and frequencies positive");
    sum += pi[k];
  }
  if (std::fabs(sum - 1.) >= 1e-3) Fail("bito_gp_set_substitution_model: frequencies do not sum to 1 +/- 0.001");
  GP_CUDA(cudaStreamSynchronize(stream_));
  InstallModel(v, vinv, lambda, pi);
  DropGraphs();  // which optimiser kernels a captured level launches depends on the number of eigenvalue groups
  AgreeOnClusterScheme();
}

// ---- site patterns: gp_engine.cpp:22-26, 544-562 -----------------------------------------
void Engine::SetSitePatterns(const uint8_t* symbols, const double* weights, bool on_device) {
  Activate();
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  // An empty shard (P_ == 0: a rank of a multi-GPU run that owns no pattern) uploads nothing but
  // runs every collective below with a zero contribution, so that all ranks issue the same sequence.
  if (P_ > 0) {
    if (P_ == P_stride_)  // rows are back to back on both sides: one linear copy instead of taxon_count_ rows
      GP_CUDA(cudaMemcpyAsync(d_symbols_.ptr, symbols, static_cast<size_t>(P_) * static_cast<size_t>(taxon_count_),
                              kind, stream_));
    else
      GP_CUDA(cudaMemcpy2DAsync(d_symbols_.ptr, static_cast<size_t>(P_stride_), symbols,
                                static_cast<size_t>(P_), static_cast<size_t>(P_),
                                static_cast<size_t>(taxon_count_), kind, stream_));
    GP_CUDA(cudaMemcpyAsync(d_weights_.ptr, weights, static_cast<size_t>(P_) * sizeof(double), kind,
                            stream_));
  }
  // Leaf P-PLVs (ids [0, taxa)) stay symbolic: 1 byte per pattern instead of 32.
  bool slots_changed = false;
  for (int64_t t = 0; t < taxon_count_; ++t) {
    PlvSlot& s = plvs_[static_cast<size_t>(t)];
    void* want = d_symbols_.ptr + t * P_stride_;
    if (s.kind == kPlvSymbols && s.ptr == want) continue;
    if (s.kind == kPlvDense) plv_pool_.Free(s.ptr);
    s.ptr = want;
    s.kind = kPlvSymbols;
    slots_changed = true;
  }
  // Total weight (all ranks) for the rescaling term of the branch-length objective.
  const int64_t tiles = TilesFor(P_);
  EnsureScratch(tiles, 2);
  LaunchFill(stream_, d_dense_tmp_.ptr, P_stride_, 1.0);
  LaunchWeightedSum(stream_, State(), d_dense_tmp_.ptr, d_partials_.ptr);  // P_ == 0: one tile, sum 0
  LaunchReducePartials(stream_, d_partials_.ptr, 1, tiles, d_packed_.ptr, nullptr, nullptr);
  AllReduce(d_packed_.ptr, 1, false);
  // The symbols are validated where they now live (a host-side scan of taxa x P bytes costs more
  // than the upload); the result rides on the read-back of the total weight.
  LaunchMaxSymbol(stream_, d_symbols_.ptr, taxon_count_, P_, P_stride_, d_packed_.ptr + 1);
  GP_CUDA(cudaMemcpyAsync(pinned_, d_packed_.ptr, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  const double new_total = static_cast<double*>(pinned_)[0];
  if (static_cast<double*>(pinned_)[1] > 4.) {
    have_patterns_ = false;
    Fail("bito_gp_set_site_patterns: symbol outside 0..4");
  }
  // captured kernels take DeviceState (total_weight) by value
  if (new_total != total_weight_) DropGraphs();
  total_weight_ = new_total;
  have_patterns_ = true;
  BuildWeightClasses(on_device && P_ > 0 ? nullptr : weights);
  AgreeOnClusterScheme();
  // Re-uploading an alignment of the same shape leaves every compiled program valid.
  if (slots_changed) InvalidatePrograms();
}

// The Brent objective (k_opt_eval_ratio) folds the per-pattern factors (1 + rho_p x)^{w_p} into a
// running product. To keep that branch-free the optimiser stores rho in a pattern order grouped by
// weight class: weights 1..7 (site-pattern multiplicities, almost always 1) get one class each,
// every other weight goes to a last "general" class evaluated with an explicit log. Classes are
// padded to whole tile groups; the padding holds rho = 0, i.e. a factor of exactly 1.
void Engine::BuildWeightClasses(const double* host_weights) {
  std::vector<double> w(static_cast<size_t>(P_));
  if (P_ == 0) {
    // nothing to read
  } else if (host_weights != nullptr) {
    std::memcpy(w.data(), host_weights, w.size() * sizeof(double));
  } else {
    GP_CUDA(cudaMemcpyAsync(w.data(), d_weights_.ptr, w.size() * sizeof(double), cudaMemcpyDeviceToHost,
                            stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
  }
  // smallest positive weight: bounds max |z| from the top power sum of the Taylor-model optimiser
  local_min_weight_ = std::numeric_limits<double>::infinity();
  for (double x : w)
    if (x > 0. && x < local_min_weight_) local_min_weight_ = x;
  if (n_ranks_ <= 1) min_weight_ = std::isfinite(local_min_weight_) ? local_min_weight_ : 1.;
  if (w == host_weights_cache_ && P_perm_ > 0 && P_ > 0) return;  // same weights: layout is current
  host_weights_cache_ = w;
  DropGraphs();  // captured optimiser launches bake the class boundaries (OptClusterLayout) in by value
  layout_version_++;
  auto cls = [](double x) {
    const int wi = static_cast<int>(x);
    return (x == static_cast<double>(wi) && wi >= 1 && wi <= 7) ? wi - 1 : 7;
  };
  // Classes are padded to whole tile groups (kTile * kOptPatternsPerThread patterns), so one
  // block-item of k_opt_eval_ratio sees a single weight.
  const int64_t group = static_cast<int64_t>(kTile) * kOptPatternsPerThread;
  int64_t n_in[8] = {0, 0, 0, 0, 0, 0, 0, 0}, start[8];
  for (double x : w) n_in[cls(x)]++;
  int64_t pos = 0;
  for (int c = 0; c < 8; ++c) {
    start[c] = pos;
    pos += RoundUp(n_in[c], group);
  }
  P_perm_ = std::max<int64_t>(pos, group);
  for (int c = 0; c < 8; ++c) opt_class_starts_.start[c] = static_cast<int32_t>(start[c] / group);
  opt_class_starts_.start[8] = static_cast<int32_t>(P_perm_ / group);
  std::vector<int32_t> perm(static_cast<size_t>(P_));
  std::vector<double> wperm(static_cast<size_t>(P_perm_), 0.);
  std::vector<uint8_t> row_class(static_cast<size_t>(P_perm_ / group), 0);
  int64_t next[8];
  for (int c = 0; c < 8; ++c) {
    next[c] = start[c];
    for (int64_t r = start[c] / group; r < (start[c] + RoundUp(n_in[c], group)) / group; ++r)
      row_class[static_cast<size_t>(r)] = static_cast<uint8_t>(c);
  }
  for (int64_t p = 0; p < P_; ++p) {
    const int c = cls(w[static_cast<size_t>(p)]);
    perm[static_cast<size_t>(p)] = static_cast<int32_t>(next[c]);
    wperm[static_cast<size_t>(next[c]++)] = w[static_cast<size_t>(p)];
  }
  // one word per pattern for k_opt_prepare_ratio: rho position and weight (1..7, else 0 = read weights[])
  std::vector<int32_t> pos_w(static_cast<size_t>(P_));
  const bool packable = P_perm_ < (int64_t(1) << 28);
  for (int64_t p = 0; p < P_ && packable; ++p) {
    const int c = cls(w[static_cast<size_t>(p)]);
    pos_w[static_cast<size_t>(p)] = (perm[static_cast<size_t>(p)] << 3) | (c < 7 ? c + 1 : 0);
  }
  GP_CUDA(cudaStreamSynchronize(stream_));
  d_pos_w_.Resize(packable ? pos_w.size() : 0, false, stream_);
  if (packable && P_ > 0)
    GP_CUDA(cudaMemcpyAsync(d_pos_w_.ptr, pos_w.data(), pos_w.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                            stream_));
  d_perm_.Resize(perm.size(), false, stream_);
  d_wperm_.Resize(wperm.size(), false, stream_);
  d_row_class_.Resize(row_class.size(), false, stream_);
  GP_CUDA(cudaMemcpyAsync(d_perm_.ptr, perm.data(), perm.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                          stream_));
  GP_CUDA(cudaMemcpyAsync(d_wperm_.ptr, wperm.data(), wperm.size() * sizeof(double),
                          cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaMemcpyAsync(d_row_class_.ptr, row_class.data(), row_class.size(), cudaMemcpyHostToDevice,
                          stream_));
  // The cluster-resident optimiser (k_opt_cluster) uses the same idea at a finer grain: classes
  // padded to rows of kClusterThreads patterns, addressed through the INVERSE permutation (a block
  // fills its own shared-memory rows, so it asks "which pattern sits at position q").
  std::vector<int32_t> inv(0);
  std::vector<double> cw(0);
  {
    const int64_t row = kClusterThreads;
    int64_t cstart[9];
    int64_t cpos = 0;
    for (int c = 0; c < 8; ++c) {
      cstart[c] = cpos;
      cpos += RoundUp(n_in[c], row);
    }
    cstart[8] = cpos;
    for (int c = 0; c < 9; ++c) cluster_class_row_start_[c] = static_cast<int32_t>(cstart[c] / row);
    inv.assign(static_cast<size_t>(std::max<int64_t>(cpos, row)), -1);
    cw.assign(inv.size(), 0.);
    int64_t cnext[8];
    for (int c = 0; c < 8; ++c) cnext[c] = cstart[c];
    for (int64_t p = 0; p < P_; ++p) {
      const int c = cls(w[static_cast<size_t>(p)]);
      inv[static_cast<size_t>(cnext[c])] = static_cast<int32_t>(p);
      cw[static_cast<size_t>(cnext[c]++)] = w[static_cast<size_t>(p)];
    }
    std::vector<int32_t> pos_of(static_cast<size_t>(std::max<int64_t>(P_, 1)), 0);
    for (size_t q = 0; q < inv.size(); ++q)
      if (inv[q] >= 0) pos_of[static_cast<size_t>(inv[q])] = static_cast<int32_t>(q);
    d_cluster_pos_.Resize(pos_of.size(), false, stream_);
    GP_CUDA(cudaMemcpyAsync(d_cluster_pos_.ptr, pos_of.data(), pos_of.size() * sizeof(int32_t),
                            cudaMemcpyHostToDevice, stream_));
    d_cluster_inv_perm_.Resize(inv.size(), false, stream_);
    d_cluster_wperm_.Resize(cw.size(), false, stream_);
    GP_CUDA(cudaMemcpyAsync(d_cluster_inv_perm_.ptr, inv.data(), inv.size() * sizeof(int32_t),
                            cudaMemcpyHostToDevice, stream_));
    GP_CUDA(cudaMemcpyAsync(d_cluster_wperm_.ptr, cw.data(), cw.size() * sizeof(double),
                            cudaMemcpyHostToDevice, stream_));
    // every cluster shape this device can run for this alignment; RunOptimizer picks one per level
    const size_t had = cluster_plans_.size();
    const int had_rows = had > 0 ? cluster_plans_[0].rows_total : -1;
    cluster_plans_.clear();
    if (opt_cluster_env_ != 0) {  // an empty shard (P = 0) plans one all-padding row: it still takes part in the exchange
      for (int threads : {256, 512, 1024}) {
        if (opt_cluster_threads_env_ > 0 && threads != opt_cluster_threads_env_) continue;
        if (opt_cluster_env_ > 0 && opt_cluster_threads_env_ <= 0 && threads != 256) continue;
        for (int c = 1; c <= kMaxOptCluster; ++c) {
          if (opt_cluster_env_ > 0 && c != opt_cluster_env_) continue;
          OptClusterPlan plan;
          if (PlanOptCluster(std::max<int64_t>(cpos / row, 1), threads, c, &plan, opt_model_env_ != 0 && opt_scheme_env_ != 3))
            cluster_plans_.push_back(plan);
        }
      }
    }
    if (getenv("BITO_GP_DEBUG_PLANS") != nullptr) {
      for (const OptClusterPlan& c : cluster_plans_)
        std::fprintf(stderr, "bito_gp plan: %2d blocks x %4d threads, %3d rows/block, %6zu B smem, %3d clusters resident\n",
                     c.cluster_size, c.threads, c.rows_per_block, c.shared_bytes, c.active_clusters);
    }
    if (had != cluster_plans_.size() ||
        (had > 0 && had_rows != cluster_plans_[0].rows_total))
      DropGraphs();  // captured optimiser launches bake the cluster shape in
  }
  GP_CUDA(cudaStreamSynchronize(stream_));  // host vectors die here
  coef_padding_zeroed_ = false;
}

void Engine::InitializePriors(const double* sbn_prior, const double* unconditional,
                              const double* inverted) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(d_q_.ptr, sbn_prior, gpcsp_count_ * sizeof(double),
                          cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaMemcpyAsync(d_inverted_.ptr, inverted, gpcsp_count_ * sizeof(double),
                          cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaMemcpyAsync(d_uncond_.ptr, unconditional, node_count_ * sizeof(double),
                          cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::SetNullPrior() {
  Activate();
  LaunchFill(stream_, d_q_.ptr, static_cast<int64_t>(d_q_.n), 1.0);
}

// ---- PLV / row residency -----------------------------------------------------------------
void Engine::EnsureDense(int64_t id) {
  PlvSlot& s = plvs_[static_cast<size_t>(id)];
  if (s.kind == kPlvDense) return;
  if (plv_pool_.NeedsChunk()) {
    const int64_t reserved =
        static_cast<int64_t>(plv_pool_.BytesReserved() + row_pool_.BytesReserved());
    if (reserved + static_cast<int64_t>(plv_pool_.ChunkBytes()) > max_device_bytes_)
      Fail("PLV storage would exceed the device memory budget (" +
           std::to_string(max_device_bytes_ >> 20) + " MiB): " +
           std::to_string(plv_pool_.SlotsInUse()) + " PLVs of " +
           std::to_string(plv_pool_.slot_bytes()) + " bytes are resident");
    // the optimiser's coefficient scratch may hold gigabytes it only needs during a sweep: PLVs come first
    if (d_coef_.n * sizeof(double) > (size_t(1) << 30)) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b < 2 * plv_pool_.ChunkBytes()) {
        GP_CUDA(cudaStreamSynchronize(stream_));
        d_coef_.Release();
        coef_padding_zeroed_ = false;
      }
    }
  }
  double* fresh = static_cast<double*>(plv_pool_.Alloc());
  if (s.kind == kPlvSymbols) {
    PlvRef src{s.ptr, kPlvSymbols, static_cast<int32_t>(id)};
    LaunchExportPlv(stream_, State(), src, fresh);
  } else {
    GP_CUDA(cudaMemsetAsync(fresh, 0, static_cast<size_t>(32 * P_stride_), stream_));
  }
  s.ptr = fresh;
  s.kind = kPlvDense;
  InvalidatePrograms();
}

double* Engine::EnsureRow(int64_t edge) {
  double*& r = rows_[static_cast<size_t>(edge)];
  if (r == nullptr) {
    if (cfg_.flags & BITO_GP_FLAG_NO_LOGLIK_MATRIX) return nullptr;
    r = static_cast<double*>(row_pool_.Alloc());
    GP_CUDA(cudaMemsetAsync(r, 0, static_cast<size_t>(8 * P_stride_), stream_));
    InvalidatePrograms();
  }
  return r;
}

PlvRef Engine::Ref(int64_t id) const {
  const PlvSlot& s = plvs_[static_cast<size_t>(id)];
  return PlvRef{s.ptr, s.kind, static_cast<int32_t>(id)};
}

void Engine::EnsureScratch(int64_t partial_doubles, int64_t packed_doubles) {
  bool grew = false;
  auto grow = [&](DeviceArray<double>& a, int64_t want) {
    if (static_cast<size_t>(want) > a.n) {
      // Round up generously so repeated growth (and graph invalidation) is rare.
      GP_CUDA(cudaStreamSynchronize(stream_));
      a.Resize(static_cast<size_t>(want + want / 2 + 64), false, stream_);
      grew = true;
    }
  };
  grow(d_partials_, partial_doubles);
  grow(d_packed_, packed_doubles);
  grow(d_dense_tmp_, 4 * P_stride_);
  if (grew) DropGraphs();  // captured graphs hold the old scratch addresses
}

void Engine::DropGraphs() {
  // a single-GPU ProcessOperations no longer waits for the device: do not destroy a graph that may be running
  if (stream_ != nullptr) cudaStreamSynchronize(stream_);
  for (auto& kv : programs_) {
    if (kv.second->graph != nullptr) {
      cudaGraphExecDestroy(kv.second->graph);
      kv.second->graph = nullptr;
    }
    kv.second->graph_tried = false;
  }
}

void Engine::AllReduce(double* buf, int64_t n, bool max_op) {
  if (n_ranks_ <= 1 || n <= 0) return;
  if (peer_ready_ && n <= kPeerCapacity) {  // a few scalars per edge: one kernel over NVLink peer memory
    LaunchPeerAllReduce(stream_, peer_, buf, static_cast<int>(n), max_op);
    stats_.collective_calls++;
    stats_.peer_collective_calls++;
    if (!capturing_) stats_.kernel_launches++;
    return;
  }
  NcclCheck(Nccl().AllReduce(buf, buf, static_cast<size_t>(n), kNcclFloat64,
                             max_op ? kNcclMax : kNcclSum, nccl_comm_, stream_),
            "ncclAllReduce");
  stats_.collective_calls++;
}

void Engine::CommInit(int n_ranks, int rank, const uint8_t id[128]) {
  Activate();
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) Fail("bito_gp_comm_init: bad rank / n_ranks");
  if (nccl_comm_ != nullptr) Fail("bito_gp_comm_init: communicator already initialised");
  n_ranks_ = n_ranks;
  rank_ = rank;
  if (n_ranks == 1) return;
  NcclUniqueId uid;
  std::memcpy(uid.internal, id, 128);
  NcclComm comm = nullptr;
  NcclCheck(Nccl().CommInitRank(&comm, n_ranks, uid, rank), "ncclCommInitRank");
  nccl_comm_ = comm;
  SetUpPeerMemory();
  if (have_patterns_) {  // total weight must now be global
    EnsureScratch(TilesFor(P_), 2);
    GP_CUDA(cudaMemcpyAsync(d_packed_.ptr, &total_weight_, sizeof(double), cudaMemcpyHostToDevice,
                            stream_));
    AllReduce(d_packed_.ptr, 1, false);
    GP_CUDA(cudaMemcpyAsync(pinned_, d_packed_.ptr, sizeof(double), cudaMemcpyDeviceToHost,
                            stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    total_weight_ = *static_cast<double*>(pinned_);
    AgreeOnClusterScheme();
  }
  InvalidatePrograms();
}

// Several ranks: whether a level's searches run in clusters must be the same decision everywhere
// (the clusters of one edge on different GPUs wait for each other). Each rank knows how many
// clusters its device keeps resident for its shard; the minimum over ranks bounds the edges of a
// level that may take the cluster path. Collective: called by every rank at the same points.
void Engine::AgreeOnClusterScheme() {
  const int old_ops = multi_rank_cluster_ops_;
  const double old_min_weight = min_weight_;
  multi_rank_cluster_ops_ = 0;
  if (n_ranks_ <= 1) return;
  int local = 0;
  for (const OptClusterPlan& c : cluster_plans_) local = std::max(local, c.active_clusters);
  if (!peer_ready_ || n_eigen_groups_ != 2) local = 0;
  EnsureScratch(TilesFor(P_), 2);
  // the smallest positive pattern weight of ANY rank rides along (Taylor-model optimiser)
  const double neg[2] = {-static_cast<double>(local), -std::min(local_min_weight_, 1e300)};
  GP_CUDA(cudaMemcpyAsync(d_packed_.ptr, neg, sizeof neg, cudaMemcpyHostToDevice, stream_));
  AllReduce(d_packed_.ptr, 2, true);  // max of the negatives = minus the minimum
  double out[2] = {0., 0.};
  GP_CUDA(cudaMemcpyAsync(out, d_packed_.ptr, sizeof out, cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  multi_rank_cluster_ops_ = std::min<int>(static_cast<int>(-out[0]), kPeerEdgeSlots);
  min_weight_ = (-out[1] > 0. && -out[1] < 1e300) ? -out[1] : 1.;
  // Captured levels bake both in (which optimiser kernel a level launches, the model's radius bound). A
  // re-upload of the same alignment - every step of a caller that streams its data in - changes neither:
  // keep the graphs (dropping them here cost a re-capture + re-instantiation of every program per upload).
  if (multi_rank_cluster_ops_ != old_ops || min_weight_ != old_min_weight) DropGraphs();
}

// Exchange buffers for k_peer_allreduce: one cudaMalloc per rank, its IPC handle all-gathered over
// the NCCL communicator that was just built, every peer's buffer mapped here (which also enables
// peer access). Any failure (no IPC in this container, no P2P path, more than 8 ranks, or
// BITO_GP_PEER_ALLREDUCE=0) leaves the engine on NCCL all-reduces: same results, more latency.
// All ranks must agree, so the outcome is itself all-reduced (min) before anyone uses the buffers.
void Engine::SetUpPeerMemory() {
  peer_ready_ = false;
  const char* env = getenv("BITO_GP_PEER_ALLREDUCE");
  bool ok = !(env != nullptr && atoi(env) == 0) && n_ranks_ <= kMaxPeerRanks && Nccl().AllGather != nullptr;
  const size_t bytes = PeerBufferBytes(n_ranks_);
  std::vector<cudaIpcMemHandle_t> handles(static_cast<size_t>(n_ranks_));
  if (ok) ok = cudaMalloc(&peer_local_, bytes) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(peer_local_, 0, bytes, stream_) == cudaSuccess;
  cudaIpcMemHandle_t mine{};
  if (ok) ok = cudaIpcGetMemHandle(&mine, peer_local_) == cudaSuccess;
  cudaGetLastError();
  // the gather runs on every rank whatever happened above (a collective must not be skipped by some)
  DeviceArray<uint8_t> d_handles;
  d_handles.Resize(sizeof(cudaIpcMemHandle_t) * static_cast<size_t>(n_ranks_ + 1), false, stream_);
  GP_CUDA(cudaMemcpyAsync(d_handles.ptr + sizeof(mine) * static_cast<size_t>(n_ranks_), &mine, sizeof(mine),
                          cudaMemcpyHostToDevice, stream_));
  if (Nccl().AllGather != nullptr) {
    NcclCheck(Nccl().AllGather(d_handles.ptr + sizeof(mine) * static_cast<size_t>(n_ranks_), d_handles.ptr,
                               sizeof(mine), kNcclInt8, nccl_comm_, stream_),
              "ncclAllGather");
    GP_CUDA(cudaMemcpyAsync(handles.data(), d_handles.ptr, sizeof(mine) * static_cast<size_t>(n_ranks_),
                            cudaMemcpyDeviceToHost, stream_));
  }
  GP_CUDA(cudaStreamSynchronize(stream_));
  d_handles.Release();
  for (int r = 0; r < n_ranks_ && ok; ++r) {
    if (r == rank_) {
      peer_.base[r] = static_cast<double*>(peer_local_);
      continue;
    }
    void* p = nullptr;
    ok = cudaIpcOpenMemHandle(&p, handles[static_cast<size_t>(r)], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    if (ok) {
      peer_.base[r] = static_cast<double*>(p);
      peer_opened_.push_back(p);
    }
  }
  cudaGetLastError();
  // agree: everyone or no one
  EnsureScratch(TilesFor(P_), 2);
  const double flag = ok ? 1. : 0.;
  GP_CUDA(cudaMemcpyAsync(d_packed_.ptr, &flag, sizeof(double), cudaMemcpyHostToDevice, stream_));
  NcclCheck(Nccl().AllReduce(d_packed_.ptr, d_packed_.ptr, 1, kNcclFloat64, /*ncclMin*/ 3, nccl_comm_, stream_),
            "ncclAllReduce");
  double all = 0.;
  GP_CUDA(cudaMemcpyAsync(&all, d_packed_.ptr, sizeof(double), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  if (all != 1.) {
    ReleasePeerMemory();
    return;
  }
  if (d_peer_epoch_.n == 0) d_peer_epoch_.Resize(1, false, stream_);
  GP_CUDA(cudaMemsetAsync(d_peer_epoch_.ptr, 0, sizeof(unsigned long long), stream_));
  if (d_peer_seq_.n == 0) d_peer_seq_.Resize(kPeerEdgeSlots, false, stream_);
  GP_CUDA(cudaMemsetAsync(d_peer_seq_.ptr, 0, kPeerEdgeSlots * sizeof(unsigned long long), stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  peer_.epoch = d_peer_epoch_.ptr;
  peer_.status = d_status_.ptr;
  peer_.n_ranks = n_ranks_;
  peer_.rank = rank_;
  if (d_peer_comm_.n == 0) d_peer_comm_.Resize(1, false, stream_);
  GP_CUDA(cudaMemcpyAsync(d_peer_comm_.ptr, &peer_, sizeof(PeerComm), cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  peer_ready_ = true;
}

PeerEdge Engine::PeerEdgeContext() const {
  PeerEdge px{};
  if (n_ranks_ > 1 && peer_ready_) {
    px.pc = d_peer_comm_.ptr;
    px.seq = d_peer_seq_.ptr;
    px.enabled = 1;
  }
  return px;
}

void Engine::ReleasePeerMemory() {
  peer_ready_ = false;
  for (void* p : peer_opened_) cudaIpcCloseMemHandle(p);
  peer_opened_.clear();
  if (peer_local_ != nullptr) cudaFree(peer_local_);
  peer_local_ = nullptr;
  cudaGetLastError();
}

void Engine::SetStream(cudaStream_t s) {
  Activate();
  GP_CUDA(cudaStreamSynchronize(stream_));
  stream_ = s != nullptr ? s : own_stream_;
  for (auto& kv : programs_) {  // graphs are stream-agnostic, but be conservative
    kv.second->graph_tried = false;
  }
}

void Engine::Synchronize() {
  Activate();
  GP_CUDA(cudaStreamSynchronize(stream_));
  CheckStatus();
}

// Surface device-side asserts with the reference's messages (gp_engine.cpp:237-238, 256-257,
// 283, 325, 585-586).
void Engine::CheckStatus() {
  status_pending_ = false;
  uint32_t* h = static_cast<uint32_t*>(pinned_);
  GP_CUDA(cudaMemcpyAsync(h, d_status_.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  const uint32_t bits = *h;
  if (bits == 0) return;
  GP_CUDA(cudaMemsetAsync(d_status_.ptr, 0, sizeof(uint32_t), stream_));
  stats_.device_status_bits |= bits;
  if (bits & kErrPeerTimeout) {
    // not an assert of the reference: a peer rank never delivered its share of an exchange. The bit
    // stays set on the device (kernels stop waiting for peers) and every call fails from here on.
    const uint32_t keep = kErrPeerTimeout;
    GP_CUDA(cudaMemcpyAsync(d_status_.ptr, &keep, sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    Fail("a peer GPU did not answer a peer-memory exchange within ~35 s (rank died or ranks issued different calls)");
  }
  if (!(cfg_.flags & BITO_GP_FLAG_STRICT_ASSERTS)) return;  // Release-build semantics
  std::string msg;
  if (bits & kErrRescalingDifference)
    msg += "dest_ rescaling too large in IncrementWithWeightedEvolvedPLV; ";
  if (bits & kErrMultiplyNotFinite) msg += "Multiply dest_ is not finite; ";
  if (bits & kErrNegativePLV) msg += "PLV with negative entry passed to RescalePLVIfNeeded; ";
  if (bits & kErrRescaledStationary)
    msg += "Surprise! Rescaled stationary distribution in IncrementMarginalLikelihood; ";
  if (bits & kErrEmptyPrep) msg += "Empty src_vector in PrepForMarginalization; ";
  if (bits & kErrQuartetRescaled)
    msg += "Rescaling not implemented in CalculateQuartetHybridLikelihoods.; ";
  msg += "(bito_gp device status " + std::to_string(bits) + ")";
  Fail(msg);
}

// ---- the op-list compiler -----------------------------------------------------------------
namespace {
enum MacroKind { kMkZero, kMkScalar, kMkStat, kMkAccum, kMkMult, kMkNode, kMkLik, kMkMarg, kMkOpt };
struct Macro {
  MacroKind kind;
  int idx;  // into the per-kind host vector
  int level = 0;
  std::vector<int64_t> reads, writes;
};
struct MargGroupHost {
  int item_off, n_items, reset;
};
}  // namespace

Program* Engine::Compile(const bito_gp_op* ops, int64_t n, const int64_t* vec, int64_t vec_len) {
  const bool fuse = !(cfg_.flags & BITO_GP_FLAG_NO_FUSION);
  const int64_t n_plv = padded_plv_count();
  const int64_t n_edge = padded_gpcsp_count();

  // -- pass 0a: which PLVs hold (or will hold) data once this list has run. A Multiply whose
  // destination owns no memory and one of whose operands is identically zero *for the whole list*
  // stays a count-only op; any other Multiply makes its destination dense. The decision must not
  // depend on the order in which later ops make an operand dense, so it is a fixpoint over the
  // whole list, taken before anything is allocated; pass 0 and pass 1 both read `elide`.
  std::vector<char> elide(static_cast<size_t>(n), 0);
  {
    std::vector<char> nonzero(static_cast<size_t>(n_plv), 0);
    for (int64_t k = 0; k < n_plv; ++k) nonzero[k] = plvs_[k].kind != kPlvZero;
    auto in_range = [&](int64_t id) { return id >= 0 && id < n_plv; };
    for (int64_t i = 0; i < n; ++i) {
      const bito_gp_op& op = ops[i];
      if ((op.kind == BITO_GP_SET_TO_STATIONARY_DISTRIBUTION ||
           op.kind == BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV) && in_range(op.a))
        nonzero[op.a] = 1;
    }
    bool changed = fuse;
    while (changed) {
      changed = false;
      for (int64_t i = 0; i < n; ++i) {
        const bito_gp_op& op = ops[i];
        if (op.kind != BITO_GP_MULTIPLY || !in_range(op.a) || !in_range(op.b) || !in_range(op.c)) continue;
        if (!nonzero[op.a] && nonzero[op.b] && nonzero[op.c]) {
          nonzero[op.a] = 1;
          changed = true;
        }
      }
    }
    if (fuse)
      for (int64_t i = 0; i < n; ++i) {
        const bito_gp_op& op = ops[i];
        if (op.kind == BITO_GP_MULTIPLY && in_range(op.a) && in_range(op.b) && in_range(op.c))
          elide[i] = !nonzero[op.a] && (!nonzero[op.b] || !nonzero[op.c]);
      }
  }

  // -- pass 0: validate and make every written PLV / row resident ------------------------
  for (int64_t i = 0; i < n; ++i) {
    const bito_gp_op& op = ops[i];
    switch (op.kind) {
      case BITO_GP_ZERO_PLV:
        CheckPlv(op.a, "ZeroPLV");
        if (plvs_[op.a].kind == kPlvSymbols) EnsureDense(op.a);
        break;
      case BITO_GP_SET_TO_STATIONARY_DISTRIBUTION:
        CheckPlv(op.a, "SetToStationaryDistribution");
        CheckEdge(op.b, "SetToStationaryDistribution");
        EnsureDense(op.a);
        break;
      case BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV:
        CheckPlv(op.a, "IncrementWithWeightedEvolvedPLV");
        CheckEdge(op.b, "IncrementWithWeightedEvolvedPLV");
        CheckPlv(op.c, "IncrementWithWeightedEvolvedPLV");
        EnsureDense(op.a);
        break;
      case BITO_GP_MULTIPLY:
        CheckPlv(op.a, "Multiply");
        CheckPlv(op.b, "Multiply");
        CheckPlv(op.c, "Multiply");
        // A product with a PLV that is (and stays) identically zero is zero: if the
        // destination owns no memory either it stays that way (e.g. the r-PLVs of leaves,
        // RHat o PHat with PHat never written).
        if (!elide[i]) EnsureDense(op.a);
        break;
      case BITO_GP_LIKELIHOOD:
        CheckEdge(op.a, "Likelihood");
        CheckPlv(op.b, "Likelihood");
        CheckPlv(op.c, "Likelihood");
        EnsureRow(op.a);
        break;
      case BITO_GP_OPTIMIZE_BRANCH_LENGTH:
        CheckPlv(op.a, "OptimizeBranchLength");
        CheckPlv(op.b, "OptimizeBranchLength");
        CheckEdge(op.c, "OptimizeBranchLength");
        break;
      case BITO_GP_UPDATE_SBN_PROBABILITIES:
        CheckEdge(op.a, "UpdateSBNProbabilities");
        if (op.b <= op.a) Fail("UpdateSBNProbabilities: empty range");
        CheckEdge(op.b - 1, "UpdateSBNProbabilities");
        break;
      case BITO_GP_RESET_MARGINAL_LIKELIHOOD:
        break;
      case BITO_GP_INCREMENT_MARGINAL_LIKELIHOOD:
        CheckPlv(op.a, "IncrementMarginalLikelihood");
        CheckEdge(op.b, "IncrementMarginalLikelihood");
        CheckPlv(op.c, "IncrementMarginalLikelihood");
        EnsureRow(op.b);
        break;
      case BITO_GP_PREP_FOR_MARGINALIZATION:
        CheckPlv(op.a, "PrepForMarginalization");
        if (op.vec_off < 0 || op.vec_len < 0 || op.vec_off + op.vec_len > vec_len)
          Fail("PrepForMarginalization: src_vector outside the vec pool");
        for (int64_t k = 0; k < op.vec_len; ++k)
          CheckPlv(vec[op.vec_off + k], "PrepForMarginalization");
        break;
      default:
        Fail("unknown GPOperation kind " + std::to_string(op.kind));
    }
  }

  // -- pass 1: fuse into macro-ops, in program order -----------------------------------------
  std::vector<ZeroOp> h_zero;
  std::vector<ScalarOp> h_scalar;
  std::vector<StatOp> h_stat;
  std::vector<AccumGroup> h_accum;
  std::vector<AccumItem> h_items;
  std::vector<int32_t> h_pool;
  std::vector<MultOp> h_mult;
  std::vector<LikOp> h_lik;
  std::vector<MargGroupHost> h_marg_groups;
  std::vector<MargItem> h_marg_items;
  std::vector<OptOp> h_opt;
  std::vector<Macro> macros;
  double alg_bytes = 0.;

  const int64_t res_bl = n_plv, res_q = n_plv + n_edge, res_row = n_plv + 2 * n_edge,
                res_marg = n_plv + 3 * n_edge;
  std::vector<char> pending_zero(static_cast<size_t>(n_plv), 0);

  auto emit_zero = [&](int64_t id) {
    Macro m;
    if (plvs_[id].kind == kPlvDense) {
      m.kind = kMkZero;
      m.idx = static_cast<int>(h_zero.size());
      h_zero.push_back(ZeroOp{static_cast<double*>(plvs_[id].ptr), static_cast<int32_t>(id), 0});
    } else {  // owns no memory: already reads as zero, only the count is reset
      m.kind = kMkScalar;
      m.idx = static_cast<int>(h_scalar.size());
      h_scalar.push_back(ScalarOp{kScalarCountZero, static_cast<int32_t>(id), 0, 0, 0, 0});
    }
    m.writes.push_back(id);
    macros.push_back(std::move(m));
  };
  auto before_read = [&](int64_t id) {
    if (pending_zero[id]) {
      pending_zero[id] = 0;
      emit_zero(id);
    }
  };
  auto add_pool = [&](const int64_t* src, int64_t len) {
    const int off = static_cast<int>(h_pool.size());
    for (int64_t k = 0; k < len; ++k) h_pool.push_back(static_cast<int32_t>(src[k]));
    return off;
  };

  int64_t i = 0;
  while (i < n) {
    const bito_gp_op& op = ops[i];
    switch (op.kind) {
      case BITO_GP_ZERO_PLV: {
        if (fuse) {
          pending_zero[op.a] = 1;  // a second ZeroPLV simply supersedes the first
        } else {
          emit_zero(op.a);
        }
        ++i;
        break;
      }
      case BITO_GP_SET_TO_STATIONARY_DISTRIBUTION: {
        pending_zero[op.a] = 0;  // fully overwritten
        Macro m;
        m.kind = kMkStat;
        m.idx = static_cast<int>(h_stat.size());
        h_stat.push_back(StatOp{static_cast<double*>(plvs_[op.a].ptr), static_cast<int32_t>(op.a),
                                static_cast<int32_t>(op.b)});
        m.writes.push_back(op.a);
        m.reads.push_back(res_q + op.b);
        macros.push_back(std::move(m));
        alg_bytes += 32;
        ++i;
        break;
      }
      case BITO_GP_PREP_FOR_MARGINALIZATION:
      case BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV: {
        // [Prep(dest)] Increment(dest, ..)+  -> one accumulate group.
        const int64_t dest = op.a;
        bool has_prep = false;
        int64_t j = i;
        if (op.kind == BITO_GP_PREP_FOR_MARGINALIZATION) {
          bool dest_in_src = false;
          for (int64_t k = 0; k < op.vec_len; ++k) dest_in_src |= (vec[op.vec_off + k] == dest);
          const bool followed = fuse && op.vec_len > 0 && !dest_in_src && i + 1 < n &&
                                ops[i + 1].kind == BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV &&
                                ops[i + 1].a == dest;
          if (!followed) {  // stand-alone Prep: scalar op
            before_read(dest);
            for (int64_t k = 0; k < op.vec_len; ++k) before_read(vec[op.vec_off + k]);
            Macro m;
            m.kind = kMkScalar;
            m.idx = static_cast<int>(h_scalar.size());
            h_scalar.push_back(ScalarOp{kScalarPrep, static_cast<int32_t>(dest), 0,
                                        add_pool(vec + op.vec_off, op.vec_len),
                                        static_cast<int32_t>(op.vec_len), 0});
            m.writes.push_back(dest);
            for (int64_t k = 0; k < op.vec_len; ++k) m.reads.push_back(vec[op.vec_off + k]);
            macros.push_back(std::move(m));
            ++i;
            break;
          }
          has_prep = true;
          j = i + 1;
        }
        AccumGroup g{};
        g.dest = static_cast<double*>(plvs_[dest].ptr);
        g.dest_id = static_cast<int32_t>(dest);
        g.item_off = static_cast<int>(h_items.size());
        Macro m;
        m.kind = kMkAccum;
        // Sources first: a pending ZeroPLV on a source must be materialised before us.
        int64_t end = j;
        while (end < n && ops[end].kind == BITO_GP_INCREMENT_WITH_WEIGHTED_EVOLVED_PLV &&
               ops[end].a == dest && (fuse || end == j)) {
          if (ops[end].c == dest && end != j) break;  // reads its own partial sum: new group
          ++end;
          if (ops[end - 1].c == dest) break;
        }
        for (int64_t k = j; k < end; ++k)
          if (ops[k].c != dest) before_read(ops[k].c);
        if (has_prep)
          for (int64_t k = 0; k < op.vec_len; ++k) before_read(vec[op.vec_off + k]);
        // Destination: fold a pending ZeroPLV.
        if (pending_zero[dest]) {
          pending_zero[dest] = 0;
          g.init_zero = 1;
          g.count_mode = kCountZero;
        }
        if (has_prep) {
          g.count_mode = kCountPrep;
          g.prep_off = add_pool(vec + op.vec_off, op.vec_len);
          g.prep_len = static_cast<int32_t>(op.vec_len);
          for (int64_t k = 0; k < op.vec_len; ++k) m.reads.push_back(vec[op.vec_off + k]);
        }
        for (int64_t k = j; k < end; ++k) {
          AccumItem it{};
          it.src = Ref(ops[k].c);
          it.edge = static_cast<int32_t>(ops[k].b);
          h_items.push_back(it);
          m.reads.push_back(ops[k].c);
          m.reads.push_back(res_bl + ops[k].b);
          m.reads.push_back(res_q + ops[k].b);
        }
        g.n_items = static_cast<int32_t>(end - j);
        m.idx = static_cast<int>(h_accum.size());
        h_accum.push_back(g);
        m.writes.push_back(dest);
        macros.push_back(std::move(m));
        alg_bytes += 32. * g.n_items + 32.;
        i = end;
        break;
      }
      case BITO_GP_MULTIPLY: {
        if (op.b != op.a) before_read(op.b);
        if (op.c != op.a) before_read(op.c);
        if (op.b == op.a || op.c == op.a) before_read(op.a);
        pending_zero[op.a] = 0;
        if (elide[i]) {
          // Decided in pass 0a for the list as a whole: the product is identically zero and the
          // destination keeps owning no memory. Only the count moves.
          Macro z;
          z.kind = kMkScalar;
          z.idx = static_cast<int>(h_scalar.size());
          h_scalar.push_back(ScalarOp{kScalarCountSum, static_cast<int32_t>(op.a),
                                      static_cast<int32_t>(op.b), static_cast<int32_t>(op.c), 0, 0});
          z.writes.push_back(op.a);
          z.reads.push_back(op.b);
          z.reads.push_back(op.c);
          macros.push_back(std::move(z));
          ++i;
          break;
        }
        Macro m;
        m.kind = kMkMult;
        m.idx = static_cast<int>(h_mult.size());
        MultOp mo{};
        mo.dest = static_cast<double*>(plvs_[op.a].ptr);
        mo.s1 = Ref(op.b);
        mo.s2 = Ref(op.c);
        mo.dest_id = static_cast<int32_t>(op.a);
        mo.max_slot = static_cast<int32_t>(h_mult.size());
        h_mult.push_back(mo);
        m.writes.push_back(op.a);
        m.reads.push_back(op.b);
        m.reads.push_back(op.c);
        macros.push_back(std::move(m));
        alg_bytes += 96;
        ++i;
        break;
      }
      case BITO_GP_LIKELIHOOD: {
        before_read(op.b);
        before_read(op.c);
        Macro m;
        m.kind = kMkLik;
        m.idx = static_cast<int>(h_lik.size());
        LikOp lo{};
        lo.parent = Ref(op.c);
        lo.child = Ref(op.b);
        lo.row = rows_[op.a];
        lo.edge = static_cast<int32_t>(op.a);
        h_lik.push_back(lo);
        m.writes.push_back(res_row + op.a);
        m.reads.push_back(op.b);
        m.reads.push_back(op.c);
        m.reads.push_back(res_bl + op.a);
        macros.push_back(std::move(m));
        alg_bytes += 72;
        ++i;
        break;
      }
      case BITO_GP_OPTIMIZE_BRANCH_LENGTH: {
        before_read(op.a);
        before_read(op.b);
        Macro m;
        m.kind = kMkOpt;
        m.idx = static_cast<int>(h_opt.size());
        OptOp oo{};
        oo.parent = Ref(op.b);
        oo.child = Ref(op.a);
        oo.edge = static_cast<int32_t>(op.c);
        h_opt.push_back(oo);
        m.reads.push_back(op.a);
        m.reads.push_back(op.b);
        m.writes.push_back(res_bl + op.c);
        macros.push_back(std::move(m));
        alg_bytes += 64;
        ++i;
        break;
      }
      case BITO_GP_UPDATE_SBN_PROBABILITIES: {
        Macro m;
        m.kind = kMkScalar;
        m.idx = static_cast<int>(h_scalar.size());
        h_scalar.push_back(ScalarOp{kScalarSbn, static_cast<int32_t>(op.a),
                                    static_cast<int32_t>(op.b), 0, 0, 0});
        for (int64_t e = op.a; e < op.b; ++e) {
          m.reads.push_back(res_row + e);
          m.writes.push_back(res_q + e);
        }
        macros.push_back(std::move(m));
        alg_bytes += 8. * static_cast<double>(op.b - op.a);
        ++i;
        break;
      }
      case BITO_GP_RESET_MARGINAL_LIKELIHOOD:
      case BITO_GP_INCREMENT_MARGINAL_LIKELIHOOD: {
        MargGroupHost g{static_cast<int>(h_marg_items.size()), 0, 0};
        Macro m;
        m.kind = kMkMarg;
        int64_t j = i;
        if (op.kind == BITO_GP_RESET_MARGINAL_LIKELIHOOD) {
          g.reset = 1;
          ++j;
        }
        while (j < n && ops[j].kind == BITO_GP_INCREMENT_MARGINAL_LIKELIHOOD) {
          const bito_gp_op& mo = ops[j];
          before_read(mo.a);
          before_read(mo.c);
          MargItem it{};
          it.stationary = Ref(mo.a);
          it.p = Ref(mo.c);
          it.row = rows_[mo.b];
          it.edge = static_cast<int32_t>(mo.b);
          h_marg_items.push_back(it);
          m.reads.push_back(mo.a);
          m.reads.push_back(mo.c);
          m.reads.push_back(res_q + mo.b);
          m.writes.push_back(res_row + mo.b);
          g.n_items++;
          alg_bytes += 88;
          ++j;
          if (!fuse) break;
        }
        m.writes.push_back(res_marg);
        m.idx = static_cast<int>(h_marg_groups.size());
        h_marg_groups.push_back(g);
        macros.push_back(std::move(m));
        i = j;
        break;
      }
      default:
        Fail("unknown GPOperation kind");
    }
  }
  // ZeroPLVs never followed by another access in this list.
  for (int64_t k = 0; k < n; ++k)
    if (ops[k].kind == BITO_GP_ZERO_PLV && pending_zero[ops[k].a]) {
      pending_zero[ops[k].a] = 0;
      emit_zero(ops[k].a);
    }

  // -- pass 1b: node fusion ------------------------------------------------------------------
  // Adjacent macro-ops in program order are merged into one NodeOp when a Multiply consumes the
  // result of the accumulate group(s) right before it: [A(X) [A(Y)] M(D = X o Y | X o Z) [M2]].
  // A contiguous run of a sequential program executed as one step is always legal as long as the
  // data flow inside the run is honoured; every site pattern is owned by one thread, which reads
  // the products' operands from its own registers. Conservative extra conditions (no other
  // aliasing inside the run) keep the kernel simple. Every remaining group / Multiply becomes a
  // single-member node, so one kernel serves them all.
  std::vector<NodeOp> h_node;
  {
    auto has = [](const std::vector<int64_t>& v, int64_t x) {
      return std::find(v.begin(), v.end(), x) != v.end();
    };
    std::vector<Macro> fused;
    fused.reserve(macros.size());
    size_t k = 0;
    while (k < macros.size()) {
      if (macros[k].kind != kMkAccum && macros[k].kind != kMkMult) {
        fused.push_back(std::move(macros[k]));
        ++k;
        continue;
      }
      NodeOp nd{};
      Macro nm;
      nm.kind = kMkNode;
      auto add_group = [&](const Macro& a) {
        nd.g[nd.n_groups++] = h_accum[a.idx];
        nm.reads.insert(nm.reads.end(), a.reads.begin(), a.reads.end());
        nm.writes.insert(nm.writes.end(), a.writes.begin(), a.writes.end());
      };
      auto group_of = [&](int64_t plv) {
        for (int g = 0; g < nd.n_groups; ++g)
          if (nd.g[g].dest_id == plv) return g;
        return -1;
      };
      auto add_mult = [&](const Macro& mm) {
        const MultOp& mo = h_mult[mm.idx];
        NodeMult x{};
        x.dest = mo.dest;
        x.s1 = mo.s1;
        x.s2 = mo.s2;
        x.dest_id = mo.dest_id;
        x.max_slot = 0;
        x.s1_group = group_of(mo.s1.id);
        x.s2_group = group_of(mo.s2.id);
        nd.m[nd.n_mults++] = x;
        for (int64_t r : mm.reads)
          if (group_of(r) < 0) nm.reads.push_back(r);
        nm.writes.insert(nm.writes.end(), mm.writes.begin(), mm.writes.end());
      };
      // A Multiply may join when it consumes a group result and aliases nothing else in the node.
      auto mult_joins = [&](const Macro& mm) {
        const MultOp& mo = h_mult[mm.idx];
        const int g1 = group_of(mo.s1.id), g2 = group_of(mo.s2.id);
        if (g1 < 0 && g2 < 0) return false;
        if (has(nm.writes, mo.dest_id) || has(nm.reads, mo.dest_id)) return false;
        if (g1 < 0 && has(nm.writes, mo.s1.id)) return false;
        if (g2 < 0 && has(nm.writes, mo.s2.id)) return false;
        return true;
      };
      size_t j = k + 1;
      if (macros[k].kind == kMkMult) {
        add_mult(macros[k]);
      } else {
        add_group(macros[k]);
        if (fuse) {
          // second group only if the Multiply right after it multiplies the two results
          if (j + 1 < macros.size() && macros[j].kind == kMkAccum && macros[j + 1].kind == kMkMult) {
            const AccumGroup& a1 = h_accum[macros[k].idx];
            const AccumGroup& a2 = h_accum[macros[j].idx];
            const MultOp& mo = h_mult[macros[j + 1].idx];
            const bool both = (mo.s1.id == a1.dest_id && mo.s2.id == a2.dest_id) ||
                              (mo.s1.id == a2.dest_id && mo.s2.id == a1.dest_id);
            const bool contiguous = a2.item_off == a1.item_off + a1.n_items;
            const bool clean = a1.dest_id != a2.dest_id && !has(macros[j].reads, a1.dest_id) &&
                               !has(macros[k].reads, a2.dest_id) && mo.dest_id != a1.dest_id &&
                               mo.dest_id != a2.dest_id && !has(macros[k].reads, mo.dest_id) &&
                               !has(macros[j].reads, mo.dest_id);
            if (both && contiguous && clean) {
              add_group(macros[j]);
              ++j;
            }
          }
          while (j < macros.size() && macros[j].kind == kMkMult && nd.n_mults < 2 &&
                 mult_joins(macros[j])) {
            add_mult(macros[j]);
            ++j;
          }
        }
      }
      nm.idx = static_cast<int>(h_node.size());
      h_node.push_back(nd);
      fused.push_back(std::move(nm));
      k = j;
    }
    macros.swap(fused);
  }

  // -- pass 2: dependency levels (RAW, WAW, WAR on PLVs, edge scalars, rows, marginal) ------
  const int64_t n_res = res_marg + 1;
  std::vector<int> last_write(static_cast<size_t>(n_res), -1), last_read(static_cast<size_t>(n_res), -1);
  int n_levels = 0;
  for (Macro& m : macros) {
    int lv = 0;
    for (int64_t r : m.reads) lv = std::max(lv, last_write[r] + 1);
    for (int64_t w : m.writes) lv = std::max(lv, std::max(last_write[w], last_read[w]) + 1);
    m.level = lv;
    for (int64_t r : m.reads) last_read[r] = std::max(last_read[r], lv);
    for (int64_t w : m.writes) last_write[w] = lv;
    n_levels = std::max(n_levels, lv + 1);
  }

  // -- pass 3: per-level contiguous tables ---------------------------------------------------
  std::vector<std::vector<int>> by_level(static_cast<size_t>(n_levels));
  for (size_t k = 0; k < macros.size(); ++k) by_level[macros[k].level].push_back(static_cast<int>(k));

  auto prog = std::make_unique<Program>();
  prog->levels.resize(static_cast<size_t>(n_levels));
  std::vector<ZeroOp> f_zero;
  std::vector<ScalarOp> f_scalar;
  std::vector<StatOp> f_stat;
  std::vector<NodeOp> f_node;
  std::vector<MultOp> f_mult;
  std::vector<LikOp> f_lik;
  std::vector<MargItem> f_marg;
  std::vector<OptOp> f_opt;
  std::vector<int32_t> f_lik_scatter, f_marg_scatter;
  const int64_t tiles = TilesFor(P_);
  const int32_t marg_slot = static_cast<int32_t>(d_ll_sum_.n - 1);
  bool dirty_matrices = false, after_opt = false;
  for (int lv = 0; lv < n_levels; ++lv) {
    Level& L = prog->levels[lv];
    L.zero_off = static_cast<int>(f_zero.size());
    L.scalar_off = static_cast<int>(f_scalar.size());
    L.stat_off = static_cast<int>(f_stat.size());
    L.node_off = static_cast<int>(f_node.size());
    L.mult_off = static_cast<int>(f_mult.size());
    L.lik_off = static_cast<int>(f_lik.size());
    L.marg_off = static_cast<int>(f_marg.size());
    L.marg_scatter_off = static_cast<int>(f_marg_scatter.size());
    L.opt_off = static_cast<int>(f_opt.size());
    for (int k : by_level[lv]) {
      const Macro& m = macros[k];
      switch (m.kind) {
        case kMkZero: f_zero.push_back(h_zero[m.idx]); L.n_zero++; break;
        case kMkScalar: f_scalar.push_back(h_scalar[m.idx]); L.n_scalar++; break;
        case kMkStat: f_stat.push_back(h_stat[m.idx]); L.n_stat++; break;
        case kMkNode: {
          NodeOp nd = h_node[m.idx];
          for (int g = 0; g < nd.n_groups; ++g)
            L.node_bytes_per_pattern += 32. * nd.g[g].n_items + 32.;
          for (int t = 0; t < nd.n_mults; ++t) {
            nd.m[t].max_slot = static_cast<int32_t>(f_mult.size());
            MultOp mo{};
            mo.dest = nd.m[t].dest;
            mo.s1 = nd.m[t].s1;
            mo.s2 = nd.m[t].s2;
            mo.dest_id = nd.m[t].dest_id;
            mo.max_slot = nd.m[t].max_slot;
            f_mult.push_back(mo);
            L.n_mult++;
            L.node_bytes_per_pattern += 96.;
          }
          f_node.push_back(nd);
          L.n_node++;
          break;
        }
        case kMkAccum:
        case kMkMult:
          Fail("internal: unfused accumulate/multiply macro after node fusion");
        case kMkLik:
          f_lik.push_back(h_lik[m.idx]);
          f_lik_scatter.push_back(h_lik[m.idx].edge);
          L.n_lik++;
          break;
        case kMkMarg: {
          const MargGroupHost& g = h_marg_groups[m.idx];
          if (L.has_marg) Fail("internal: two marginal groups in one level");
          L.has_marg = 1;
          L.marg_reset = g.reset;
          for (int t = 0; t < g.n_items; ++t) {
            f_marg.push_back(h_marg_items[g.item_off + t]);
            f_marg_scatter.push_back(h_marg_items[g.item_off + t].edge);
            L.n_marg++;
          }
          f_marg_scatter.push_back(marg_slot);
          break;
        }
        case kMkOpt: f_opt.push_back(h_opt[m.idx]); L.n_opt++; break;
      }
    }
    prog->max_partials = std::max<int64_t>(
        prog->max_partials, std::max<int64_t>(L.n_lik, L.has_marg ? L.n_marg + 1 : 0) * tiles);
    prog->max_packed = std::max<int64_t>(prog->max_packed,
                                         std::max<int64_t>(L.n_lik, L.n_marg + 1));
    // Transition matrices are (re)built before the first level and after any level that may have
    // changed branch lengths (OptimizeBranchLength) or q (UpdateSBNProbabilities).
    L.rebuild_matrices = (lv == 0) || dirty_matrices;
    L.rebuild_after_opt = after_opt && !L.rebuild_matrices;
    after_opt = L.n_opt > 0;
    dirty_matrices = false;
    for (int t = 0; t < L.n_scalar; ++t) dirty_matrices |= (f_scalar[L.scalar_off + t].kind == kScalarSbn);
    prog->launches += (L.n_zero > 0) + (L.n_scalar > 0) + (L.n_stat > 0) + (L.n_node > 0) +
                      (L.n_mult > 0) + 2 * (L.n_lik > 0) + 2 * L.has_marg + L.rebuild_matrices;
  }
  prog->n_mult_total = static_cast<int>(f_mult.size());
  prog->n_opt_total = static_cast<int>(f_opt.size());
  for (const Level& L : prog->levels) prog->n_opt_levels += (L.n_opt > 0);
  prog->n_items_total = static_cast<int64_t>(h_items.size());
  prog->n_lik_total = static_cast<int>(f_lik.size());
  prog->n_macro = static_cast<int64_t>(macros.size());
  prog->alg_bytes_per_pattern = alg_bytes;
  prog->alloc_version = alloc_version_;

  // Each OptimizeBranchLength carries the transition-matrix slots that hold its edge, so that the
  // on-chip optimisers refresh just those when they finish (see OptOp).
  if (!f_opt.empty()) {
    std::unordered_map<int32_t, std::vector<int32_t>> slots_of_edge;
    for (size_t i = 0; i < h_items.size(); ++i)
      slots_of_edge[h_items[i].edge].push_back(static_cast<int32_t>(2 * i));
    for (size_t j = 0; j < f_lik.size(); ++j)
      slots_of_edge[f_lik[j].edge].push_back(static_cast<int32_t>(2 * j + 1));
    for (OptOp& o : f_opt) {
      const std::vector<int32_t>& v = slots_of_edge[o.edge];
      o.fix_off = static_cast<int32_t>(h_pool.size());
      o.fix_n = static_cast<int32_t>(v.size());
      h_pool.insert(h_pool.end(), v.begin(), v.end());
    }
  }

  // -- pass 4: one device arena for all tables -------------------------------------------------
  size_t off = 0;
  auto reserve = [&off](size_t bytes) {
    const size_t at = off;
    off += RoundUp(static_cast<int64_t>(std::max<size_t>(bytes, 8)), 256);
    return at;
  };
  const size_t o_zero = reserve(f_zero.size() * sizeof(ZeroOp));
  const size_t o_scalar = reserve(f_scalar.size() * sizeof(ScalarOp));
  const size_t o_stat = reserve(f_stat.size() * sizeof(StatOp));
  const size_t o_node = reserve(f_node.size() * sizeof(NodeOp));
  const size_t o_items = reserve(h_items.size() * sizeof(AccumItem));
  const size_t o_mult = reserve(f_mult.size() * sizeof(MultOp));
  const size_t o_lik = reserve(f_lik.size() * sizeof(LikOp));
  const size_t o_marg = reserve(f_marg.size() * sizeof(MargItem));
  const size_t o_opt = reserve(f_opt.size() * sizeof(OptOp));
  const size_t o_pool = reserve(h_pool.size() * sizeof(int32_t));
  const size_t o_ls = reserve(f_lik_scatter.size() * sizeof(int32_t));
  const size_t o_ms = reserve(f_marg_scatter.size() * sizeof(int32_t));
  prog->arena_bytes = off;
  std::vector<char> host(off, 0);
  auto put = [&host](size_t at, const void* src, size_t bytes) {
    if (bytes > 0) std::memcpy(host.data() + at, src, bytes);
  };
  put(o_zero, f_zero.data(), f_zero.size() * sizeof(ZeroOp));
  put(o_scalar, f_scalar.data(), f_scalar.size() * sizeof(ScalarOp));
  put(o_stat, f_stat.data(), f_stat.size() * sizeof(StatOp));
  put(o_node, f_node.data(), f_node.size() * sizeof(NodeOp));
  put(o_items, h_items.data(), h_items.size() * sizeof(AccumItem));
  put(o_mult, f_mult.data(), f_mult.size() * sizeof(MultOp));
  put(o_lik, f_lik.data(), f_lik.size() * sizeof(LikOp));
  put(o_marg, f_marg.data(), f_marg.size() * sizeof(MargItem));
  put(o_opt, f_opt.data(), f_opt.size() * sizeof(OptOp));
  put(o_pool, h_pool.data(), h_pool.size() * sizeof(int32_t));
  put(o_ls, f_lik_scatter.data(), f_lik_scatter.size() * sizeof(int32_t));
  put(o_ms, f_marg_scatter.data(), f_marg_scatter.size() * sizeof(int32_t));
  GP_CUDA(cudaMalloc(&prog->arena, off));
  GP_CUDA(cudaMemcpyAsync(prog->arena, host.data(), off, cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));  // `host` dies at scope exit
  char* a = prog->arena;
  prog->d_zero = reinterpret_cast<ZeroOp*>(a + o_zero);
  prog->d_scalar = reinterpret_cast<ScalarOp*>(a + o_scalar);
  prog->d_stat = reinterpret_cast<StatOp*>(a + o_stat);
  prog->d_node = reinterpret_cast<NodeOp*>(a + o_node);
  prog->d_items = reinterpret_cast<AccumItem*>(a + o_items);
  prog->d_mult = reinterpret_cast<MultOp*>(a + o_mult);
  prog->d_lik = reinterpret_cast<LikOp*>(a + o_lik);
  prog->d_marg = reinterpret_cast<MargItem*>(a + o_marg);
  prog->d_opt = reinterpret_cast<OptOp*>(a + o_opt);
  prog->d_pool = reinterpret_cast<int32_t*>(a + o_pool);
  prog->d_lik_scatter = reinterpret_cast<int32_t*>(a + o_ls);
  prog->d_marg_scatter = reinterpret_cast<int32_t*>(a + o_ms);

  stats_.programs_compiled++;
  Program* raw = prog.get();
  const OpsHash hash = HashOps(ops, n, vec, vec_len);
  const uint64_t key = hash.key;
  prog->check_hash = hash.check;
  prog->n_ops = n;
  prog->vec_len = vec_len;
  prog->last_used = ++program_clock_;
  auto it = programs_.find(key);
  if (it != programs_.end()) FreeProgram(*it->second);
  programs_[key] = std::move(prog);
  EvictPrograms(raw);
  return raw;
}

// Programs compiled against an older PLV / edge allocation can never run again (ProcessOperations
// recompiles on a version mismatch), and an NNI search issues a new, larger list after every DAG
// growth plus one list per proposed NNI: free the dead ones now, and keep the live ones bounded
// (least recently used first). `keep` is the program that is about to run.
void Engine::EvictPrograms(const Program* keep) {
  for (auto it = programs_.begin(); it != programs_.end();) {
    Program* p = it->second.get();
    if (p != keep && p->alloc_version != alloc_version_) {
      FreeProgram(*p);
      it = programs_.erase(it);
      stats_.programs_evicted++;
    } else {
      ++it;
    }
  }
  while (programs_.size() > kMaxCachedPrograms) {
    auto victim = programs_.end();
    for (auto it = programs_.begin(); it != programs_.end(); ++it)
      if (it->second.get() != keep && (victim == programs_.end() || it->second->last_used < victim->second->last_used))
        victim = it;
    if (victim == programs_.end()) break;
    FreeProgram(*victim->second);
    programs_.erase(victim);
    stats_.programs_evicted++;
  }
}

void Engine::FreeProgram(Program& p) {
  if ((p.graph != nullptr || p.arena != nullptr) && stream_ != nullptr) cudaStreamSynchronize(stream_);  // may be running
  if (p.graph != nullptr) cudaGraphExecDestroy(p.graph);
  p.graph = nullptr;
  if (p.arena != nullptr) cudaFree(p.arena);
  p.arena = nullptr;
}

// ---- execution --------------------------------------------------------------------------------
void Engine::ExecuteLevels(Program& prog, size_t first, size_t last) {
  const DeviceState st = State();
  const int64_t tiles = TilesFor(P_);
  for (size_t li = first; li < last; ++li) {
    const Level& L = prog.levels[li];
    const double Pd = static_cast<double>(P_);
    if (L.n_zero > 0) {
      ProfScope ps(this, kProfZero, 32. * L.n_zero * Pd);
      LaunchZero(stream_, st, prog.d_zero + L.zero_off, L.n_zero);
    }
    if (L.n_scalar > 0) {
      ProfScope ps(this, kProfScalar, 0.);
      LaunchScalar(stream_, st, prog.d_scalar + L.scalar_off, prog.d_pool, L.n_scalar);
    }
    if (L.n_stat > 0) {
      ProfScope ps(this, kProfStationary, 32. * L.n_stat * Pd);
      LaunchStationary(stream_, st, prog.d_stat + L.stat_off, L.n_stat);
    }
    const bool stale = L.rebuild_after_opt && matrices_stale_;
    if (stale && !capturing_) stats_.kernel_launches++;  // not part of prog.launches
    if ((L.rebuild_matrices || stale) && (prog.n_items_total > 0 || prog.n_lik_total > 0)) {
      matrices_stale_ = false;
      ProfScope ps(this, kProfPrologue, 0.);
      LaunchBuildMatrices(stream_, st, prog.d_items, static_cast<int>(prog.n_items_total), prog.d_lik,
                          prog.n_lik_total, d_mtab_.ptr, d_mtab_lik_.ptr);
    }
    if (L.n_node > 0) {
      {
        ProfScope ps(this, kProfNode, L.node_bytes_per_pattern * Pd);
        LaunchNodes(stream_, st, prog.d_node + L.node_off, prog.d_items, prog.d_pool, d_mtab_.ptr,
                    L.n_node, d_level_max_.ptr);
      }
      if (L.n_mult > 0) {
        // The rescale decision needs the max over ALL patterns of the PLV (gp_engine.cpp:583-597).
        AllReduce(d_level_max_.ptr + L.mult_off, L.n_mult, true);
        ProfScope ps(this, kProfRescale, 0.);
        LaunchRescale(stream_, st, prog.d_mult + L.mult_off, L.n_mult, d_level_max_.ptr + L.mult_off);
      }
    }
    if (L.n_lik > 0) {
      {
        ProfScope ps(this, kProfLikelihood, 72. * L.n_lik * Pd);
        LaunchLikelihood(stream_, st, prog.d_lik + L.lik_off, L.n_lik,
                         d_mtab_lik_.ptr + 16 * static_cast<int64_t>(L.lik_off), d_partials_.ptr);
      }
      ProfScope ps(this, kProfReduce, 0.);
      const int64_t lik_groups = LikelihoodTileGroups(L.n_lik, P_);
      if (n_ranks_ == 1) {
        LaunchReducePartials(stream_, d_partials_.ptr, L.n_lik, lik_groups, d_packed_.ptr,
                             prog.d_lik_scatter + L.lik_off, st.ll_sum);
      } else {
        LaunchReducePartials(stream_, d_partials_.ptr, L.n_lik, lik_groups, d_packed_.ptr, nullptr,
                             nullptr);
        AllReduce(d_packed_.ptr, L.n_lik, false);
        LaunchScatter(stream_, d_packed_.ptr, L.n_lik, prog.d_lik_scatter + L.lik_off, st.ll_sum);
      }
    }
    if (L.has_marg) {
      {
        ProfScope ps(this, kProfMarginal, 88. * L.n_marg * Pd);
        LaunchMarginal(stream_, st, prog.d_marg + L.marg_off, L.n_marg, L.marg_reset,
                       d_partials_.ptr);
      }
      ProfScope ps(this, kProfReduce, 0.);
      const int n_out = L.n_marg + 1;
      if (n_ranks_ == 1) {
        LaunchReducePartials(stream_, d_partials_.ptr, n_out, tiles, d_packed_.ptr,
                             prog.d_marg_scatter + L.marg_scatter_off, st.ll_sum);
      } else {
        LaunchReducePartials(stream_, d_partials_.ptr, n_out, tiles, d_packed_.ptr, nullptr,
                             nullptr);
        AllReduce(d_packed_.ptr, n_out, false);
        LaunchScatter(stream_, d_packed_.ptr, n_out, prog.d_marg_scatter + L.marg_scatter_off,
                      st.ll_sum);
      }
    }
    if (L.n_opt > 0) RunOptimizeLevel(prog, L);
  }
}

// One block per edge can hold the per-pattern coefficients of its edge in shared memory: the whole
// 1-D search runs on chip. Needs every pattern on this rank (the objective is a sum over ALL patterns).
OptParams Engine::OptimizerParams(bool check_convergence) const {
  OptParams prm{};
  prm.significant_digits = significant_digits_;
  prm.check_convergence = check_convergence ? 1 : 0;
  prm.max_iter = kMaxIterForOptimization;
  prm.min_log_bl = kMinLogBranchLength;
  prm.max_log_bl = kMaxLogBranchLength;
  prm.denominator_tolerance = kDenominatorToleranceForNewton;
  prm.step_size = kStepSizeForOptimization;
  prm.log_step_size = kStepSizeForLogSpaceOptimization;
  prm.diff_threshold = kBranchLengthDifferenceThreshold;
  prm.brent_tolerance = std::ldexp(1.0, 1 - significant_digits_);
  prm.decimal_tolerance = std::pow(10., static_cast<double>(-significant_digits_));
  return prm;
}

// How a level of n_ops OptimizeBranchLength ops runs. 0: round-per-launch scheme (rho streamed from
// HBM every objective round; the host looks at a counter every few rounds, so no graph capture);
// 1: one block per edge (k_opt_block, every method); 2: one thread-block cluster per edge
// (k_opt_cluster: plain Brent on a two-eigenvalue model), *plan = its shape.
// The on-chip searches are bound by latency per edge (~16 dependent objective evaluations), the
// round scheme by HBM bandwidth: few edges -> spread each over as many SMs as a cluster has; a level
// with thousands of edges at 1e5 patterns -> stream.
int Engine::OptScheme(int n_ops, const OptClusterPlan** plan) const {
  if (plan != nullptr) *plan = nullptr;
  if (cfg_.flags & BITO_GP_FLAG_NO_ONCHIP_OPTIMIZER) return 0;
  if (n_ranks_ > 1) {
    // one cluster per edge on every GPU, their per-evaluation sums exchanged over NVLink inside the
    // kernel (peer_edge_sum): only when every cluster of the level is resident at once on every rank
    // (they wait for each other), which is the reference's Gauss-Seidel schedule; else rounds + all-reduce
    if (method_ != BITO_GP_BRENT_OPTIMIZATION || n_ops > multi_rank_cluster_ops_) return 0;
    const OptClusterPlan* best_plan = nullptr;
    double best = 0.;
    for (const OptClusterPlan& c : cluster_plans_) {
      if (c.active_clusters < n_ops) continue;
      const int rows_per_thread = (c.rows_per_block * kClusterThreads + c.threads - 1) / c.threads;
      const double us = opt_model_env_ != 0
                            ? 6.0 + 0.3 * ((rows_per_thread + 1) / 2) +
                                  5.0 * (3.0 + 0.12 * rows_per_thread + (c.threads > 512 ? 2.0 : 0.))
                            : 1.0 + 2.5 * ((rows_per_thread + 1) / 2) +
                                  14.5 * (2.4 + 0.05 * rows_per_thread + (c.threads > 256 ? 0.5 : 0.));
      if (best_plan == nullptr || us < best) {
        best = us;
        best_plan = &c;
      }
    }
    if (best_plan == nullptr) return 0;  // cannot happen: multi_rank_cluster_ops_ <= this rank's maximum
    if (plan != nullptr) *plan = best_plan;
    return 2;
  }
  if (P_ <= 0) return 0;
  const bool cluster_ok = !cluster_plans_.empty() && n_eigen_groups_ == 2 &&
                          method_ == BITO_GP_BRENT_OPTIMIZATION;
  const bool forced = opt_cluster_env_ > 0;
  if (opt_scheme_env_ == 0) return 0;
  if (!(cluster_ok && (forced || opt_scheme_env_ >= 2)) &&
      OptBlockSharedBytes(P_, n_eigen_groups_) <= kOptBlockMaxSharedBytes)
    return 1;
  if (!cluster_ok) return 0;
  // Scheme 3, pipelined clusters (BITO_GP_OPT_SCHEME=3 only): a streaming producer (HBM-bound) turns the
  // PLVs of the next chunk of edges into rho while the clusters run the searches of the current chunk
  // from shared memory, so shared memory only ever holds edges whose search is running. Measured on
  // B200 (profiles/r02_sweep_ab.md): the searches are bound by latency per edge - 21-30 edges fit the
  // chip's shared memory at 1e5 patterns and each takes ~70 us (~16 dependent objective evaluations,
  // each a cluster barrier plus a serial FP64 optimiser step; warps wait at barriers 47 % of the time)
  // - i.e. 45 ms for 11 139 edges whatever the producer does, against 46 ms for the WHOLE sweep with the
  // streaming scheme below. It stays available (and tested) but is never chosen automatically.
  if (opt_scheme_env_ == 3) {
    const OptClusterPlan* widest = nullptr;
    for (const OptClusterPlan& c : cluster_plans_) {
      if (widest == nullptr || c.active_clusters > widest->active_clusters ||
          (c.active_clusters == widest->active_clusters && c.threads < widest->threads))
        widest = &c;
    }
    if (plan != nullptr) *plan = widest;
    return 3;
  }
  // Cost model, microseconds, fitted to B200 timings (profiles/r01g_sweep_variants_*.log):
  //  on chip, per edge: two-row load trips of ~2.5 us, then ~14.5 dependent objective evaluations of
  //  2.4 us (reduction + cluster barrier + optimiser step) + 0.05 us per rho row a thread walks,
  //  +0.5 us with 1024-thread blocks (wider barriers); x1.4 once edges queue for SMs (co-resident
  //  clusters contend); streamed: 64 + 8 + 8 x evaluations bytes per pattern at 5.5 TB/s plus ~8 us
  //  of launches per round and the host's look at the active-edge counter.
  // With the Taylor model (k_opt_cluster_model, the default) a search is ~5 rounds of ~3 us + 0.12 us per rho
  // row a thread walks (twelve power sums per pattern; 1024-thread blocks are held to 64 registers and
  // spill: +2 us), and the streamed scheme re-reads rho ~3 times instead of ~15 (profiles/r02_sweep_model.md).
  const bool model = opt_model_env_ != 0;
  const double n_evals = model ? 5.0 : 14.5;
  double best = 0.;
  const OptClusterPlan* best_plan = nullptr;
  for (const OptClusterPlan& c : cluster_plans_) {
    const int rows_per_thread = (c.rows_per_block * kClusterThreads + c.threads - 1) / c.threads;
    // (the load is a latency-bound gather of two PLVs whatever the block size: measured 16 x 512 threads 767 ms,
    // 16 x 1024 threads 820 ms for the 11 139 single-edge levels of the 1000-taxon sweep, profiles/r02_sweep_model.md)
    const double load_us = model ? 6.0 + 0.3 * ((rows_per_thread + 1) / 2) : 1.0 + 2.5 * ((rows_per_thread + 1) / 2);
    const double eval_us = model ? 3.0 + 0.12 * rows_per_thread + (c.threads > 512 ? 2.0 : 0.)
                                 : 2.4 + 0.05 * rows_per_thread + (c.threads > 256 ? 0.5 : 0.);
    const double waves = std::ceil(static_cast<double>(n_ops) / c.active_clusters);
    const double us = waves * (load_us + n_evals * eval_us) * (waves > 1. ? 1.4 : 1.);
    if (best_plan == nullptr || us < best) {
      best = us;
      best_plan = &c;
    }
  }
  if (!forced) {
    const double streamed_bytes =
        static_cast<double>(n_ops) * static_cast<double>(P_) * (64. + 8. + 8. * (model ? 3.0 : n_evals));
    const double rounds_us = streamed_bytes / 5.5e6 + (model ? 10. : n_evals) * 8. + 30.;
    if (rounds_us < best) return 0;
  }
  if (plan != nullptr) *plan = best_plan;
  return 2;
}

// True when every OptimizeBranchLength level of the program runs on chip, i.e. the whole program
// is free of host round trips and can be captured as one CUDA graph.
bool Engine::ProgramOptimizesOnChip(const Program& prog) const {
  for (const Level& L : prog.levels)
    if (L.n_opt > 0 && OptScheme(L.n_opt, nullptr) == 0) return false;
  return true;
}

void Engine::RunOptimizeLevel(Program& prog, const Level& L) {
  opt_refresh_ = OptRefresh{prog.d_pool, d_mtab_.ptr, d_mtab_lik_.ptr};
  RunOptimizer(prog.d_opt + L.opt_off, L.n_opt, method_, optimization_count_ != 0);
  // the round scheme leaves the matrix tables to a full rebuild before the next level
  if (last_opt_scheme_ == 0) matrices_stale_ = true;
}

// Device-resident 1-D optimisers stepping every edge of the batch in lockstep: one objective
// evaluation per round (eval -> [reduce -> all-reduce] -> step). The PLVs of an edge are read once
// (k_opt_prepare*); every round streams only the per-pattern coefficients.
void Engine::RunOptimizer(const OptOp* d_ops, int n_ops, int method, bool check_convergence) {
  const DeviceState st = State();
  const int64_t tiles = TilesFor(P_);
  const int G = n_eigen_groups_;
  const int nd = method == BITO_GP_BRENT_OPTIMIZATION ? 0
                 : method == BITO_GP_NEWTON_OPTIMIZATION ? 2 : 1;
  // Plain Brent on a two-eigenvalue model (JC69, the only model GPEngine instantiates,
  // gp_engine.hpp:366): ratio form, 8 B per pattern and one log per 8 patterns.
  const OptClusterPlan* plan = nullptr;
  const int on_chip = OptScheme(n_ops, &plan);
  last_opt_scheme_ = on_chip;
  last_opt_plan_ = plan != nullptr ? *plan : OptClusterPlan();
  if (on_chip == 1 || on_chip == 2) {
    if (capturing_) capture_opt_launches_++; else stats_.kernel_launches++;
  }
  if (on_chip == 1) {
    // small alignment, single rank: every edge's whole search in one launch (k_opt_block); the
    // settings were written to d_opt_ctl_ by Execute, outside any captured graph
    ProfScope ps(this, kProfOptBlock, 64. * n_ops * static_cast<double>(P_));
    LaunchOptBlock(stream_, st, d_ops, n_ops, d_opt_ctl_.ptr, G, opt_refresh_);
    return;
  }
  if (on_chip == 3) {
    RunOptimizerPipelined(d_ops, n_ops, *plan);
    return;
  }
  if (on_chip == 2) {
    // large alignment, single rank, plain Brent: one thread-block cluster per edge (k_opt_cluster)
    ProfScope ps(this, kProfOptCluster, 64. * n_ops * static_cast<double>(P_));
    GP_CUDA(LaunchOptCluster(stream_, st, d_ops, n_ops, d_opt_ctl_.ptr, d_cluster_inv_perm_.ptr,
                             d_cluster_wperm_.ptr, cluster_class_row_start_, *plan, opt_refresh_,
                             PeerEdgeContext(), nullptr, 0, nullptr, opt_model_env_ != 0, min_weight_));
    return;
  }
  const OptParams prm = OptimizerParams(check_convergence);
  const bool ratio = (G == 2 && nd == 0);
  const int64_t coef_per_op = ratio ? P_perm_ : P_stride_ * G;
  // Coefficient scratch of one batch of edges: the whole level when HBM has room for it (fewer, fuller
  // launches: 25.6 -> 23.2 ms on the 1000-taxon shard), at most 40 % of what is free right now and
  // 16 GiB, at least 1 GiB; BITO_GP_OPT_CHUNK_MB pins it.
  int64_t budget_bytes = opt_chunk_bytes_;
  if (budget_bytes <= 0) {
    const int64_t have = static_cast<int64_t>(d_coef_.n * sizeof(double));  // already ours
    const int64_t want = static_cast<int64_t>(n_ops) * coef_per_op * static_cast<int64_t>(sizeof(double));
    budget_bytes = std::max<int64_t>(int64_t(1) << 30, have);
    if (want > budget_bytes) {  // ask the driver only when the scratch would have to grow (the query is not cheap)
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const int64_t room = static_cast<int64_t>(static_cast<double>(free_b) * 0.4) + have;
        budget_bytes = std::max<int64_t>(budget_bytes, std::min<int64_t>(room, int64_t(16) << 30));
      }
    }
  }
  const int64_t budget_doubles = budget_bytes / 8;
  const int chunk = static_cast<int>(
      std::max<int64_t>(1, std::min<int64_t>(n_ops, budget_doubles / coef_per_op)));
  if (static_cast<size_t>(chunk * coef_per_op) > d_coef_.n) {
    d_coef_.Resize(static_cast<size_t>(chunk * coef_per_op), false, stream_);
    coef_padding_zeroed_ = false;
  }
  if (ratio && !coef_padding_zeroed_) {  // padding rows of the weight-class layout: rho = 0
    GP_CUDA(cudaMemsetAsync(d_coef_.ptr, 0, d_coef_.n * sizeof(double), stream_));
    coef_padding_zeroed_ = true;
  }
  if (!ratio) coef_padding_zeroed_ = false;
  d_opt_states_.Resize(static_cast<size_t>(chunk), false, stream_);
  d_opt_const_.Resize(static_cast<size_t>(chunk), false, stream_);
  d_opt_active_.Resize(static_cast<size_t>(4 + 2 * chunk), false, stream_);
  // Taylor-model Brent (gp_types.h, OptPass): a pass yields kOptPassValues sums per edge and answers
  // several of the optimiser's requests; BITO_GP_OPT_MODEL=0 keeps one pass per objective evaluation.
  const bool model = ratio && opt_model_env_ != 0;
  int seg_len = 1, n_seg = 1;
  if (model) OptModelSegments(P_perm_, &seg_len, &n_seg);
  const int64_t groups = model ? n_seg : (ratio ? OptRatioPartials(P_perm_) : tiles);
  const int n_values = nd + 1, value_stride = model ? kOptPassValues : (ratio ? 1 : 3);
  EnsureScratch(static_cast<int64_t>(chunk) * std::max<int64_t>(value_stride * groups, tiles),
                static_cast<int64_t>(chunk) * std::max(3, value_stride));
  if (model) {
    d_opt_pass_.Resize(static_cast<size_t>(chunk), false, stream_);
    d_opt_req_.Resize(static_cast<size_t>(2 * chunk), false, stream_);
  }

  const int64_t max_rounds = 3 * kMaxIterForOptimization + 8;
  int32_t* h_active = static_cast<int32_t*>(pinned_);

  for (int c0 = 0; c0 < n_ops; c0 += chunk) {
    const int m = std::min(chunk, n_ops - c0);
    if (ratio) {
      {
        ProfScope ps(this, kProfOptPrepare, 64. * m * static_cast<double>(P_));
        LaunchOptPrepareRatio(stream_, st, d_ops + c0, m, d_opt_states_.ptr, prm, method, d_coef_.ptr,
                              d_perm_.ptr, d_pos_w_.n > 0 ? d_pos_w_.ptr : nullptr, P_perm_, d_partials_.ptr, d_opt_active_.ptr, chunk);
      }
      ProfScope ps(this, kProfReduce, 0.);
      LaunchReducePartials(stream_, d_partials_.ptr, m, OptPrepareTileGroups(m, P_), d_opt_const_.ptr,
                           nullptr, nullptr);
      AllReduce(d_opt_const_.ptr, m, false);
      stats_.kernel_launches += 2;
    } else {
      ProfScope ps(this, kProfOptPrepare, 64. * m * static_cast<double>(P_));
      LaunchOptPrepare(stream_, st, d_ops + c0, m, d_opt_states_.ptr, prm, method, d_coef_.ptr, 1);
      stats_.kernel_launches++;
    }
    int64_t rounds = 0;
    int32_t* act = ratio ? d_opt_active_.ptr : nullptr;
    if (model) {
      LaunchOptPlan(stream_, st, m, d_opt_states_.ptr, prm, d_opt_pass_.ptr, d_opt_req_.ptr, act);
      stats_.kernel_launches++;
      int batch = opt_model_first_check_;  // most searches need 1..8 passes
      for (;;) {
        for (int r = 0; r < batch; ++r) {
          const int parity = static_cast<int>((rounds + r) & 1);
          {
            ProfScope ps(this, kProfOptEval, 0.);
            LaunchOptEvalModel(stream_, m, rounds + r == 0 ? kOptPoints : 1,
                               d_opt_req_.ptr + static_cast<size_t>(parity) * chunk, d_coef_.ptr, P_perm_,
                               d_wperm_.ptr, opt_class_starts_, d_partials_.ptr, act, parity);
          }
          const double* sums = nullptr;
          const double* parts = d_partials_.ptr;
          if (n_ranks_ > 1) {
            ProfScope ps(this, kProfReduce, 0.);
            LaunchReducePartials(stream_, d_partials_.ptr, kOptPassValues * m, n_seg, d_packed_.ptr, nullptr,
                                 nullptr);
            AllReduce(d_packed_.ptr, static_cast<int64_t>(kOptPassValues) * m, false);
            sums = d_packed_.ptr;
            parts = nullptr;
            stats_.kernel_launches++;
          }
          ProfScope ps(this, kProfOptStep, 0.);
          LaunchOptStepModel(stream_, st, m, d_opt_states_.ptr, d_opt_pass_.ptr, d_opt_req_.ptr, prm, sums, parts,
                             n_seg, d_opt_const_.ptr, min_weight_, act, chunk, parity);
          stats_.kernel_launches += 2;
        }
        rounds += batch;
        GP_CUDA(cudaMemcpyAsync(h_active, act + (rounds & 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        GP_CUDA(cudaStreamSynchronize(stream_));
        if (*h_active == 0) break;
        if (rounds > max_rounds) Fail("OptimizeBranchLength: optimiser did not terminate");
        batch = 2;
      }
      continue;
    }
    int batch = method <= BITO_GP_BRENT_OPTIMIZATION_WITH_GRADIENTS ? 12 : 6;
    for (;;) {
      int32_t* counter = nullptr;
      for (int r = 0; r < batch; ++r) {
        const bool last = (r == batch - 1);
        const int parity = static_cast<int>((rounds + r) & 1);
        {
          ProfScope ps(this, kProfOptEval, 0.);
          if (ratio)
            LaunchOptEvalRatio(stream_, st, m, d_opt_states_.ptr, d_coef_.ptr, P_perm_, d_wperm_.ptr,
                               d_row_class_.ptr, d_partials_.ptr, act, chunk, parity);
          else
            LaunchOptEval(stream_, st, m, d_opt_states_.ptr, d_coef_.ptr, nd, d_partials_.ptr, G);
        }
        const double* sums = nullptr;
        const double* parts = d_partials_.ptr;
        if (n_ranks_ > 1) {
          ProfScope ps(this, kProfReduce, 0.);
          LaunchReducePartials(stream_, d_partials_.ptr, value_stride * m, groups, d_packed_.ptr,
                               nullptr, nullptr);
          AllReduce(d_packed_.ptr, static_cast<int64_t>(value_stride) * m, false);
          sums = d_packed_.ptr;
          parts = nullptr;
          stats_.kernel_launches++;
        }
        if (ratio) {
          counter = act + (parity ^ 1);  // edges still active after this round
        } else {
          counter = d_active_.ptr;
          if (last) GP_CUDA(cudaMemsetAsync(d_active_.ptr, 0, sizeof(int32_t), stream_));
        }
        ProfScope ps(this, kProfOptStep, 0.);
        LaunchOptStep(stream_, st, m, d_opt_states_.ptr, prm, sums, parts, static_cast<int>(groups),
                      n_values, value_stride, ratio ? d_opt_const_.ptr : nullptr,
                      (!ratio && last) ? d_active_.ptr : nullptr, act, chunk, parity);
        stats_.kernel_launches += 2;
      }
      rounds += batch;
      GP_CUDA(cudaMemcpyAsync(h_active, counter, sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      GP_CUDA(cudaStreamSynchronize(stream_));
      if (*h_active == 0) break;
      if (rounds > max_rounds) Fail("OptimizeBranchLength: optimiser did not terminate");
      batch = 4;
    }
  }
}

// Scheme 3. Two streams, two chunk buffers:
//   producer (prep_stream_): k_opt_prepare_cluster(chunk i) -> rho ring half i&1, K_e partials ->
//                            k_reduce_partials -> K_e                      [HBM-bound: 64 B / pattern / edge]
//   consumer (stream_):      k_opt_cluster<T, true>(chunk i): one cluster per edge copies its rho rows
//                            to shared memory and runs the whole Brent search there [latency-bound]
// chunk i + 1 is produced while chunk i is searched; events hand the halves back and forth. No
// host round trip, so the level is capturable (the producer stream joins the capture through the
// fork event and rejoins through the last `ready` event).
int Engine::EnsurePipelineBuffers(int n_ops, const OptClusterPlan& plan) {
  const int64_t rho_stride = static_cast<int64_t>(plan.rows_total) * kClusterThreads;
  // chunk: several waves of resident clusters (so that the tail of a launch, where its last searches
  // finish alone, stays small) within ~128 MiB of rho per half
  int chunk = opt_ring_edges_env_ > 0 ? opt_ring_edges_env_
                                      : std::max(6 * plan.active_clusters,
                                                 static_cast<int>((int64_t(128) << 20) / (8 * rho_stride)));
  chunk = std::max(1, std::min(chunk, n_ops));
  // the producer's tile-group count depends on the op count (TilesPerBlock): size for the worst case
  const int64_t max_groups = TilesFor(P_);
  const size_t ring_doubles = static_cast<size_t>(2) * chunk * rho_stride;
  if (d_rho_ring_.n < ring_doubles || ring_rho_stride_ != rho_stride) {
    if (capturing_) Fail("internal: pipelined optimiser buffers must exist before graph capture");
    GP_CUDA(cudaStreamSynchronize(stream_));
    GP_CUDA(cudaStreamSynchronize(prep_stream_));
    if (d_rho_ring_.n < ring_doubles) {
      d_rho_ring_.Release();
      d_rho_ring_.Resize(ring_doubles, false, stream_);
      DropGraphs();
    }
    // padding positions of the class layout hold rho = 0 (a factor of exactly 1) and are never
    // written by the producer; a new layout (other weights, other shape) moves them
    GP_CUDA(cudaMemsetAsync(d_rho_ring_.ptr, 0, d_rho_ring_.n * sizeof(double), stream_));
    ring_rho_stride_ = rho_stride;
    ring_layout_version_ = layout_version_;
  } else if (ring_layout_version_ != layout_version_) {
    if (capturing_) Fail("internal: pipelined optimiser buffers must exist before graph capture");
    GP_CUDA(cudaStreamSynchronize(prep_stream_));
    GP_CUDA(cudaMemsetAsync(d_rho_ring_.ptr, 0, d_rho_ring_.n * sizeof(double), stream_));
    ring_layout_version_ = layout_version_;
  }
  if (d_ring_const_.n < static_cast<size_t>(2 * chunk) ||
      d_ring_partials_.n < static_cast<size_t>(2 * chunk * max_groups)) {
    if (capturing_) Fail("internal: pipelined optimiser buffers must exist before graph capture");
    GP_CUDA(cudaStreamSynchronize(stream_));
    GP_CUDA(cudaStreamSynchronize(prep_stream_));
    d_ring_const_.Resize(static_cast<size_t>(2 * chunk), false, stream_);
    d_ring_partials_.Resize(static_cast<size_t>(2 * chunk * max_groups), false, stream_);
    DropGraphs();
  }
  return chunk;
}

void Engine::RunOptimizerPipelined(const OptOp* d_ops, int n_ops, const OptClusterPlan& plan) {
  const DeviceState st = State();
  const int64_t rho_stride = static_cast<int64_t>(plan.rows_total) * kClusterThreads;
  const int chunk = EnsurePipelineBuffers(n_ops, plan);
  const int64_t max_groups = TilesFor(P_);
  // with per-kernel profiling on, everything runs on the engine's stream (serialised, but timed)
  cudaStream_t ps = profiling_ ? stream_ : prep_stream_;
  cudaStream_t cs = (profiling_ || opt_priority_env_ == 0) ? stream_ : cons_stream_;
  if (ps != stream_) {
    GP_CUDA(cudaEventRecord(ev_fork_, stream_));
    GP_CUDA(cudaStreamWaitEvent(ps, ev_fork_, 0));
    if (cs != stream_) GP_CUDA(cudaStreamWaitEvent(cs, ev_fork_, 0));
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_);
  const int prep_max_blocks = (profiling_ || prep_blocks_per_sm_ <= 0) ? 0 : prep_blocks_per_sm_ * sms;
  int i = 0;
  for (int c0 = 0; c0 < n_ops; c0 += chunk, ++i) {
    const int m = std::min(chunk, n_ops - c0);
    const int b = i & 1;
    double* rho = d_rho_ring_.ptr + static_cast<size_t>(b) * chunk * rho_stride;
    double* consts = d_ring_const_.ptr + static_cast<size_t>(b) * chunk;
    double* parts = d_ring_partials_.ptr + static_cast<size_t>(b) * chunk * max_groups;
    if (ps != stream_ && i >= 2) GP_CUDA(cudaStreamWaitEvent(ps, ev_free_[b], 0));
    {
      ProfScope scope(this, kProfOptPrepare, 64. * m * static_cast<double>(P_));
      LaunchOptPrepareCluster(ps, st, d_ops + c0, m, rho, d_cluster_pos_.ptr, rho_stride, parts, prep_max_blocks);
    }
    {
      ProfScope scope(this, kProfReduce, 0.);
      LaunchReducePartials(ps, parts, m, OptPrepareTileGroups(m, P_), consts, nullptr, nullptr);
    }
    if (ps != stream_) {
      GP_CUDA(cudaEventRecord(ev_ready_[b], ps));
      GP_CUDA(cudaStreamWaitEvent(cs, ev_ready_[b], 0));
    }
    {
      ProfScope scope(this, kProfOptCluster, 0.);
      GP_CUDA(LaunchOptCluster(cs, st, d_ops + c0, m, d_opt_ctl_.ptr, d_cluster_inv_perm_.ptr,
                               d_cluster_wperm_.ptr, cluster_class_row_start_, plan, opt_refresh_,
                               PeerEdgeContext(), rho, rho_stride, consts));
    }
    if (ps != stream_) GP_CUDA(cudaEventRecord(ev_free_[b], cs));
    if (capturing_) capture_opt_launches_ += 3; else stats_.kernel_launches += 3;
  }
  if (cs != stream_) {  // the engine's stream continues after the last search
    GP_CUDA(cudaEventRecord(ev_join_, cs));
    GP_CUDA(cudaStreamWaitEvent(stream_, ev_join_, 0));
  }
}

void Engine::Execute(Program& prog) {
  EnsureScratch(prog.max_partials, prog.max_packed);
  {
    bool grew = false;
    auto grow = [&](DeviceArray<double>& a, int64_t want) {
      if (static_cast<size_t>(want) > a.n) {
        GP_CUDA(cudaStreamSynchronize(stream_));
        a.Resize(static_cast<size_t>(want + want / 4 + 64), false, stream_);
        grew = true;
      }
    };
    grow(d_level_max_, prog.n_mult_total);
    grow(d_mtab_, 16 * prog.n_items_total);
    grow(d_mtab_lik_, 16 * static_cast<int64_t>(prog.n_lik_total));
    if (grew) DropGraphs();  // captured graphs hold the old scratch addresses
  }
  // A program that holds a captured graph was captured under the current settings (everything the
  // per-level choice of optimiser kernel depends on drops the graphs when it changes), so a replay
  // does not walk its levels again: the reference's Gauss-Seidel list has 11 139 optimiser levels, and
  // two cost-model evaluations per level were milliseconds of host time in front of every launch.
  const bool replay = prog.graph != nullptr && !(cfg_.flags & BITO_GP_FLAG_NO_CUDA_GRAPHS) && !profiling_;
  const bool opt_on_chip = replay ? prog.n_opt_total > 0 : (prog.n_opt_total > 0 && ProgramOptimizesOnChip(prog));
  const bool want_graph = !(cfg_.flags & BITO_GP_FLAG_NO_CUDA_GRAPHS) &&
                          (prog.n_opt_total == 0 || opt_on_chip) && !profiling_;
  // optimiser launches are counted where they are issued (RunOptimizer), except inside a graph
  // replay, where every OptimizeBranchLength level is exactly one on-chip launch
  const int64_t launches = prog.launches;
  if (opt_on_chip && !replay) {  // the pipelined scheme's buffers cannot be (re)allocated inside a capture
    for (const Level& L : prog.levels) {
      const OptClusterPlan* plan = nullptr;
      if (L.n_opt > 0 && OptScheme(L.n_opt, &plan) == 3) EnsurePipelineBuffers(L.n_opt, *plan);
    }
  }
  if (prog.n_opt_total > 0) {  // read by k_opt_block / k_opt_cluster, whichever levels use them
    OptControl ctl{};
    ctl.prm = OptimizerParams(optimization_count_ != 0);
    ctl.method = method_;
    ctl.n_derivatives = method_ == BITO_GP_BRENT_OPTIMIZATION ? 0
                        : method_ == BITO_GP_NEWTON_OPTIMIZATION ? 2 : 1;
    if (d_opt_ctl_.n == 0) d_opt_ctl_.Resize(1, false, stream_);
    LaunchSetOptControl(stream_, d_opt_ctl_.ptr, ctl);
  }
  auto body = [&]() {
    if (prog.n_mult_total > 0)
      GP_CUDA(cudaMemsetAsync(d_level_max_.ptr, 0, prog.n_mult_total * sizeof(double), stream_));
    ExecuteLevels(prog, 0, prog.levels.size());
  };
  if (want_graph) {
    if (prog.graph == nullptr && !prog.graph_tried) {
      prog.graph_tried = true;
      cudaGraph_t graph = nullptr;
      GP_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
      capturing_ = true;
      capture_opt_launches_ = 0;
      try {
        body();
      } catch (...) {
        capturing_ = false;
        cudaStreamEndCapture(stream_, &graph);
        if (graph != nullptr) cudaGraphDestroy(graph);
        throw;
      }
      capturing_ = false;
      prog.graph_opt_launches = capture_opt_launches_;
      GP_CUDA(cudaStreamEndCapture(stream_, &graph));
      cudaError_t err = cudaGraphInstantiate(&prog.graph, graph, 0);
      cudaGraphDestroy(graph);
      if (err != cudaSuccess) {
        prog.graph = nullptr;
        cudaGetLastError();
      }
    }
    if (prog.graph != nullptr) {
      GP_CUDA(cudaGraphLaunch(prog.graph, stream_));
      stats_.graph_launches++;
      stats_.kernel_launches += prog.launches + prog.graph_opt_launches;
      return;
    }
  }
  body();
  stats_.kernel_launches += launches;
}

void Engine::ProcessOperations(const bito_gp_op* ops, int64_t n, const int64_t* vec,
                               int64_t vec_len) {
  Activate();
  if (!have_patterns_) Fail("ProcessOperations: call bito_gp_set_site_patterns first");
  BindModel();
  stats_.process_calls++;
  if (n == 0) return;
  const OpsHash hash = HashOps(ops, n, vec, vec_len);
  Program* prog = nullptr;
  auto it = programs_.find(hash.key);
  if (it != programs_.end() && it->second->alloc_version == alloc_version_ && it->second->n_ops == n &&
      it->second->vec_len == vec_len && it->second->check_hash == hash.check) {
    prog = it->second.get();
    prog->last_used = ++program_clock_;
  } else {
    prog = Compile(ops, n, vec, vec_len);
    // Compile may itself have made PLVs resident (bumping the version) before building the
    // tables, so the tables are current: stamp them with the final version.
    prog->alloc_version = alloc_version_;
  }
  stats_.levels_last = static_cast<int64_t>(prog->levels.size());
  stats_.fused_ops_last = prog->n_macro;
  stats_.algorithmic_bytes_last = prog->alg_bytes_per_pattern * static_cast<double>(P_);
  GP_CUDA(cudaEventRecord(ev_begin_, stream_));
  Execute(*prog);
  GP_CUDA(cudaEventRecord(ev_end_, stream_));
  timing_pending_ = true;
  // The device status word mirrors the reference's Asserts. On one GPU without STRICT_ASSERTS they are
  // Release-build no-ops that only accumulate in the statistics, so the call does not wait for the device:
  // the host hashes and launches the caller's next list while this one runs (every getter is ordered on
  // stream_ and synchronises it; Synchronize / GetStats collect the word). With peers a timed-out exchange
  // must fail THIS call, and STRICT_ASSERTS promises the reference's exception here: both wait.
  if (n_ranks_ > 1 || (cfg_.flags & BITO_GP_FLAG_STRICT_ASSERTS))
    CheckStatus();
  else
    status_pending_ = true;
}

// ---- branch lengths / optimiser settings ---------------------------------------------------------
void Engine::SetBranchLengths(const double* bl) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(d_bl_.ptr, bl, gpcsp_count_ * sizeof(double), cudaMemcpyHostToDevice,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::SetBranchLengthsRange(int64_t start, int64_t length, const double* bl) {
  Activate();
  if (start < 0 || length < 0 || start + length > padded_gpcsp_count())
    Fail("Requested range of BranchLengths is out-of-range.");
  GP_CUDA(cudaMemcpyAsync(d_bl_.ptr + start, bl, length * sizeof(double), cudaMemcpyHostToDevice,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::SetBranchLengthsToConstant(double v) {
  Activate();
  LaunchFill(stream_, d_bl_.ptr, static_cast<int64_t>(d_bl_.n), v);
}
void Engine::SetBranchLengthsToDefault() { SetBranchLengthsToConstant(kDefaultBranchLength); }
void Engine::GetBranchLengths(int64_t start, int64_t length, double* out) {
  Activate();
  if (start < 0 || length < 0 || start + length > padded_gpcsp_count())
    Fail("Requested range of BranchLengths is out-of-range.");
  GP_CUDA(cudaMemcpyAsync(out, d_bl_.ptr + start, length * sizeof(double), cudaMemcpyDeviceToHost,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetBranchLengthDifferences(double* out) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(out, d_diff_.ptr, gpcsp_count_ * sizeof(double), cudaMemcpyDeviceToHost,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::SetOptimizationMethod(int m) {
  if (m < 0 || m > 4) Fail("DAGBranchHandler::Optimization(): Invalid OptimizationMethod given.");
  if (m != method_) DropGraphs();  // which optimiser kernel a captured level launches depends on it
  method_ = m;
}
void Engine::ResetOptimizationCount() {  // dag_branch_handler.hpp:49-52
  Activate();
  optimization_count_ = 0;
  LaunchFill(stream_, d_diff_.ptr, static_cast<int64_t>(d_diff_.n), 0.0);
}

void Engine::LogLikelihoodAndDerivatives(int64_t gpcsp, int64_t rootward, int64_t leafward,
                                         double out[3]) {
  Activate();
  BindModel();
  CheckEdge(gpcsp, "LogLikelihoodAndDerivative");
  CheckPlv(rootward, "LogLikelihoodAndDerivative");
  CheckPlv(leafward, "LogLikelihoodAndDerivative");
  const DeviceState st = State();
  const int64_t tiles = TilesFor(P_);
  const int G = n_eigen_groups_;
  d_coef_.Resize(static_cast<size_t>(P_stride_ * G), false, stream_);
  coef_padding_zeroed_ = false;
  d_opt_states_.Resize(1, false, stream_);
  d_single_opt_.Resize(1, false, stream_);
  EnsureScratch(3 * tiles, 3);
  double bl = 0.;
  int32_t counts[2];
  GP_CUDA(cudaMemcpyAsync(&bl, d_bl_.ptr + gpcsp, sizeof(double), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaMemcpyAsync(&counts[0], d_counts_.ptr + rootward, sizeof(int32_t),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaMemcpyAsync(&counts[1], d_counts_.ptr + leafward, sizeof(int32_t),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  OptOp op{};
  op.parent = Ref(rootward);
  op.child = Ref(leafward);
  op.edge = static_cast<int32_t>(gpcsp);
  OptState s{};
  s.t_eval = bl;
  s.done = 0;
  for (int g = 0; g < n_eigen_groups_; ++g) s.e[g] = std::exp(group_lambda_[g] * bl);
  GP_CUDA(cudaMemcpyAsync(d_single_opt_.ptr, &op, sizeof op, cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaMemcpyAsync(d_opt_states_.ptr, &s, sizeof s, cudaMemcpyHostToDevice, stream_));
  OptParams prm{};
  LaunchOptPrepare(stream_, st, d_single_opt_.ptr, 1, d_opt_states_.ptr, prm, 0, d_coef_.ptr, 0);
  LaunchOptEval(stream_, st, 1, d_opt_states_.ptr, d_coef_.ptr, 2, d_partials_.ptr, G);
  LaunchReducePartials(stream_, d_partials_.ptr, 3, tiles, d_packed_.ptr, nullptr, nullptr);
  AllReduce(d_packed_.ptr, 3, false);
  GP_CUDA(cudaMemcpyAsync(out, d_packed_.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  stats_.kernel_launches += 3;
  out[0] += (static_cast<double>(counts[0]) * st.log_thr + static_cast<double>(counts[1]) * st.log_thr) *
            total_weight_;
}

// ---- quartet hybrid marginals: gp_engine.cpp:748-816 ---------------------------------------------
namespace {
// NumericalUtils::LogAdd (numerical_utils.hpp:35-52); LogSum is its left fold (numerical_utils.cpp).
double HostLogAdd(double x, double y) {
  if (y > x) std::swap(x, y);
  if (x == -std::numeric_limits<double>::infinity()) return x;
  const double neg_diff = y - x;
  if (neg_diff < std::log(std::numeric_limits<double>::epsilon())) return x;
  return x + std::log(1.0 + std::exp(neg_diff));
}
}  // namespace

void Engine::QuartetHybrid(int64_t n_requests, const int64_t* central, const int32_t* tip_counts,
                           const bito_gp_quartet_tip* tips, double* likelihoods, bool store) {
  Activate();
  if (n_requests < 0) Fail("QuartetHybrid: negative request count");
  BindModel();
  std::vector<QuartetItem> items;
  std::vector<int64_t> first_item(static_cast<size_t>(n_requests) + 1, 0);
  int64_t tip_off = 0;
  for (int64_t r = 0; r < n_requests; ++r) {
    const int32_t* n = tip_counts + 4 * r;
    for (int k = 0; k < 4; ++k)
      if (n[k] < 0) Fail("QuartetHybrid: negative tip count");
    const bito_gp_quartet_tip* rw = tips + tip_off;
    const bito_gp_quartet_tip* sis = rw + n[0];
    const bito_gp_quartet_tip* rot = sis + n[1];
    const bito_gp_quartet_tip* sorted = rot + n[2];
    tip_off += static_cast<int64_t>(n[0]) + n[1] + n[2] + n[3];
    CheckEdge(central[r], "QuartetHybridRequest");
    for (const bito_gp_quartet_tip* t = rw; t < sorted + n[3]; ++t) {
      CheckPlv(t->plv_idx, "QuartetHybridRequest");
      CheckEdge(t->gpcsp_idx, "QuartetHybridRequest");
    }
    first_item[r] = static_cast<int64_t>(items.size());
    // Loop nest of CalculateQuartetHybridLikelihoods: rootward, sister, rotated, sorted (innermost).
    for (int a = 0; a < n[0]; ++a) {
      if (rw[a].tip_node_id < 0 || rw[a].tip_node_id >= node_count_ + spare_nodes_)
        Fail("QuartetHybridRequest: rootward tip node id out of range");
      for (int b = 0; b < n[1]; ++b)
        for (int c = 0; c < n[2]; ++c)
          for (int d = 0; d < n[3]; ++d) {
            QuartetItem it{};
            it.rootward = Ref(rw[a].plv_idx);
            it.sister = Ref(sis[b].plv_idx);
            it.rotated = Ref(rot[c].plv_idx);
            it.sorted = Ref(sorted[d].plv_idx);
            it.edge[0] = static_cast<int32_t>(rw[a].gpcsp_idx);
            it.edge[1] = static_cast<int32_t>(sis[b].gpcsp_idx);
            it.edge[2] = static_cast<int32_t>(central[r]);
            it.edge[3] = static_cast<int32_t>(rot[c].gpcsp_idx);
            it.edge[4] = static_cast<int32_t>(sorted[d].gpcsp_idx);
            it.rootward_node = static_cast<int32_t>(rw[a].tip_node_id);
            items.push_back(it);
          }
    }
  }
  first_item[static_cast<size_t>(n_requests)] = static_cast<int64_t>(items.size());
  const int64_t n_items = static_cast<int64_t>(items.size());
  std::vector<double> result(static_cast<size_t>(n_items));
  if (n_items > 0) {
    if (n_items > (int64_t{1} << 30) / TilesFor(P_)) Fail("QuartetHybrid: too many summands in one call");
    const DeviceState st = State();
    const int64_t tiles = TilesFor(P_);
    d_quartet_items_.Resize(static_cast<size_t>(n_items), false, stream_);
    d_quartet_mats_.Resize(static_cast<size_t>(80 * n_items), false, stream_);
    EnsureScratch(n_items * tiles, 2 * n_items);
    GP_CUDA(cudaMemcpyAsync(d_quartet_items_.ptr, items.data(), items.size() * sizeof(QuartetItem),
                            cudaMemcpyHostToDevice, stream_));
    LaunchQuartetMatrices(stream_, st, d_quartet_items_.ptr, static_cast<int>(n_items), d_quartet_mats_.ptr);
    LaunchQuartet(stream_, st, d_quartet_items_.ptr, static_cast<int>(n_items), d_quartet_mats_.ptr,
                  d_partials_.ptr);
    LaunchReducePartials(stream_, d_partials_.ptr, static_cast<int>(n_items), tiles, d_packed_.ptr, nullptr,
                         nullptr);
    AllReduce(d_packed_.ptr, n_items, false);
    LaunchQuartetFinish(stream_, st, d_quartet_items_.ptr, static_cast<int>(n_items), d_packed_.ptr,
                        d_packed_.ptr + n_items);
    GP_CUDA(cudaMemcpyAsync(result.data(), d_packed_.ptr + n_items, n_items * sizeof(double),
                            cudaMemcpyDeviceToHost, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    stats_.kernel_launches += 4;
    CheckStatus();
  }
  if (likelihoods != nullptr && n_items > 0) std::memcpy(likelihoods, result.data(), n_items * sizeof(double));
  if (store) {
    for (int64_t r = 0; r < n_requests; ++r) {
      const int64_t b = first_item[r], e = first_item[r + 1];
      if (e == b) continue;  // not fully formed (gp_engine.cpp:811): entry stays as it is
      double acc = result[b];
      for (int64_t i = b + 1; i < e; ++i) acc = HostLogAdd(acc, result[i]);
      GP_CUDA(cudaMemcpyAsync(d_hybrid_.ptr + central[r], &acc, sizeof(double), cudaMemcpyHostToDevice,
                              stream_));
      GP_CUDA(cudaStreamSynchronize(stream_));
    }
  }
}

void Engine::GetHybridMarginals(double* out) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(out, d_hybrid_.ptr, gpcsp_count_ * sizeof(double), cudaMemcpyDeviceToHost,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::GetTransitionMatrix(double t, double out[16]) {
  Activate();
  BindModel();
  EnsureScratch(16, 16);
  LaunchTransitionMatrix(stream_, t, d_packed_.ptr);
  GP_CUDA(cudaMemcpyAsync(out, d_packed_.ptr, 16 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}

// ---- read-back: gp_engine.cpp:413-468 ---------------------------------------------------------------
double Engine::GetLogMarginalLikelihood() {
  Activate();
  double v = 0.;
  GP_CUDA(cudaMemcpyAsync(&v, d_ll_sum_.ptr + (d_ll_sum_.n - 1), sizeof(double),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
  return v;
}
void Engine::GetPerGpcspLogLikelihoods(int64_t start, int64_t length, double* out) {
  Activate();
  if (start < 0 || length < 0 || start + length > padded_gpcsp_count())
    Fail("Requested range of PerGPCSPLogLikelihoods is out-of-range.");
  GP_CUDA(cudaMemcpyAsync(out, d_ll_sum_.ptr + start, length * sizeof(double),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetPerGpcspComponentsOfFullLogMarginal(double* out) {
  std::vector<double> q(static_cast<size_t>(gpcsp_count_));
  GetPerGpcspLogLikelihoods(0, gpcsp_count_, out);
  GetSbnParameters(q.data());
  for (int64_t e = 0; e < gpcsp_count_; ++e)
    out[e] += static_cast<double>(site_count_) * std::log(q[e]);
}
void Engine::GetLogLikelihoodMatrix(double* out) {
  Activate();
  if (cfg_.flags & BITO_GP_FLAG_NO_LOGLIK_MATRIX)
    Fail("GetLogLikelihoodMatrix: engine was created with BITO_GP_FLAG_NO_LOGLIK_MATRIX");
  for (int64_t e = 0; e < gpcsp_count_; ++e) {
    if (rows_[e] == nullptr) {
      std::memset(out + e * P_, 0, static_cast<size_t>(P_) * sizeof(double));
    } else {
      GP_CUDA(cudaMemcpyAsync(out + e * P_, rows_[e], static_cast<size_t>(P_) * sizeof(double),
                              cudaMemcpyDeviceToHost, stream_));
    }
  }
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetPerPatternLogMarginal(double* out) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(out, d_log_marg_.ptr, static_cast<size_t>(P_) * sizeof(double),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetSbnParameters(double* out) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(out, d_q_.ptr, gpcsp_count_ * sizeof(double), cudaMemcpyDeviceToHost,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::SetSbnParameters(const double* q) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(d_q_.ptr, q, gpcsp_count_ * sizeof(double), cudaMemcpyHostToDevice,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetPlv(int64_t id, double* out) {
  Activate();
  CheckPlv(id, "GetPLV");
  const PlvSlot& s = plvs_[id];
  if (s.kind == kPlvZero) {
    std::memset(out, 0, static_cast<size_t>(4 * P_) * sizeof(double));
    return;
  }
  const double* src = static_cast<const double*>(s.ptr);
  if (s.kind == kPlvSymbols) {
    EnsureScratch(0, 0);
    LaunchExportPlv(stream_, State(), Ref(id), d_dense_tmp_.ptr);
    src = d_dense_tmp_.ptr;
  }
  GP_CUDA(cudaMemcpyAsync(out, src, static_cast<size_t>(4 * P_) * sizeof(double),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::SetPlv(int64_t id, const double* in, int32_t count) {
  Activate();
  CheckPlv(id, "SetPLV");
  if (plvs_[id].kind != kPlvDense) {
    // Content is about to be replaced: allocate without expanding.
    plvs_[id].kind = kPlvZero;
    EnsureDense(id);
  }
  GP_CUDA(cudaMemcpyAsync(plvs_[id].ptr, in, static_cast<size_t>(4 * P_) * sizeof(double),
                          cudaMemcpyHostToDevice, stream_));
  GP_CUDA(cudaMemcpyAsync(d_counts_.ptr + id, &count, sizeof(int32_t), cudaMemcpyHostToDevice,
                          stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}
void Engine::GetRescalingCounts(int32_t* out) {
  Activate();
  GP_CUDA(cudaMemcpyAsync(out, d_counts_.ptr, static_cast<size_t>(padded_plv_count()) * sizeof(int32_t),
                          cudaMemcpyDeviceToHost, stream_));
  GP_CUDA(cudaStreamSynchronize(stream_));
}

// ---- resize / copy: gp_engine.cpp:64-209, 386-409 ------------------------------------------------------
// PLV id = type * node_count + node, so a change of node_count moves every id. With PLVs
// behind a slot table this is a table permutation; no PLV data moves in HBM.
void Engine::GrowPlvs(int64_t new_node_count, const int64_t* reindexer, int64_t explicit_alloc) {
  Activate();
  (void)explicit_alloc;  // capacity is on-demand here; there is nothing to pre-allocate
  if (new_node_count < taxon_count_) Fail("GrowPLVs: node_count below taxon_count");
  const int64_t old_n = node_count_;
  const int64_t old_padded = padded_plv_count();
  std::vector<int32_t> old_counts(static_cast<size_t>(old_padded));
  GetRescalingCounts(old_counts.data());
  std::vector<PlvSlot> old_slots = plvs_;
  node_count_ = new_node_count;
  const int64_t new_padded = padded_plv_count();
  std::vector<PlvSlot> slots(static_cast<size_t>(new_padded));
  std::vector<int32_t> counts(static_cast<size_t>(new_padded), 0);
  std::vector<char> moved(static_cast<size_t>(old_padded), 0);
  const int64_t old_stride = old_n, new_stride = new_node_count;
  // Reindexer semantics (reindexer.hpp): new_index = reindexer[old_index]; nodes that did
  // not exist before keep fresh (zero) PLVs. Spare PLVs keep their offsets past 6N.
  for (int type = 0; type < 6; ++type) {
    for (int64_t node = 0; node < std::min(old_n, new_node_count); ++node) {
      const int64_t to = reindexer != nullptr ? reindexer[node] : node;
      if (to < 0 || to >= new_node_count) Fail("Node Reindexer is not valid.");
      const int64_t src = type * old_stride + node, dst = type * new_stride + to;
      slots[dst] = old_slots[src];
      counts[dst] = old_counts[src];
      moved[src] = 1;
    }
  }
  for (int64_t j = 0; j < 6 * spare_nodes_; ++j) {
    const int64_t src = 6 * old_stride + j, dst = 6 * new_stride + j;
    slots[dst] = old_slots[src];
    counts[dst] = old_counts[src];
    moved[src] = 1;
  }
  for (int64_t k = 0; k < old_padded; ++k)
    if (!moved[k] && old_slots[k].kind == kPlvDense) plv_pool_.Free(old_slots[k].ptr);
  plvs_ = std::move(slots);
  GP_CUDA(cudaStreamSynchronize(stream_));
  d_counts_.Resize(static_cast<size_t>(new_padded), false, stream_);
  GP_CUDA(cudaMemcpyAsync(d_counts_.ptr, counts.data(), counts.size() * sizeof(int32_t),
                          cudaMemcpyHostToDevice, stream_));
  // unconditional_node_probabilities_: reindexed with the nodes, 1 for nodes that are new
  // (gp_engine.cpp:100-104, 171-172).
  {
    std::vector<double> old_u(static_cast<size_t>(old_n + spare_nodes_), 1.);
    if (d_uncond_.ptr != nullptr)
      GP_CUDA(cudaMemcpyAsync(old_u.data(), d_uncond_.ptr,
                              std::min(old_u.size(), d_uncond_.n) * sizeof(double),
                              cudaMemcpyDeviceToHost, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    std::vector<double> u(static_cast<size_t>(node_count_ + spare_nodes_), 1.);
    for (int64_t node = 0; node < std::min(old_n, new_node_count); ++node)
      u[static_cast<size_t>(reindexer != nullptr ? reindexer[node] : node)] = old_u[static_cast<size_t>(node)];
    d_uncond_.Resize(u.size(), false, stream_);
    GP_CUDA(cudaMemcpyAsync(d_uncond_.ptr, u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice,
                            stream_));
  }
  GP_CUDA(cudaStreamSynchronize(stream_));
  InvalidatePrograms();
}

void Engine::GrowGpcsps(int64_t new_count, const int64_t* reindexer, int64_t explicit_alloc) {
  Activate();
  (void)explicit_alloc;
  const int64_t old_count = gpcsp_count_;
  const int64_t old_padded = padded_gpcsp_count();
  auto pull = [&](DeviceArray<double>& a, int64_t n) {
    std::vector<double> v(static_cast<size_t>(n));
    GP_CUDA(cudaMemcpyAsync(v.data(), a.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    return v;
  };
  std::vector<double> q = pull(d_q_, old_padded), inv = pull(d_inverted_, old_padded),
                      bl = pull(d_bl_, old_padded), diff = pull(d_diff_, old_padded),
                      hyb = pull(d_hybrid_, old_padded), lls = pull(d_ll_sum_, old_padded);
  const double marg = GetLogMarginalLikelihood();
  gpcsp_count_ = new_count;
  const int64_t new_padded = padded_gpcsp_count();
  const double ninf = -std::numeric_limits<double>::infinity();
  std::vector<double> nq(new_padded, 1.), ninv(new_padded, 1.), nbl(new_padded, kDefaultBranchLength),
      ndiff(new_padded, 0.), nhyb(new_padded, ninf), nlls(new_padded + 1, 0.);
  std::vector<double*> nrows(static_cast<size_t>(new_padded), nullptr);
  std::vector<char> moved(static_cast<size_t>(old_padded), 0);
  for (int64_t e = 0; e < std::min(old_count, new_count); ++e) {
    const int64_t to = reindexer != nullptr ? reindexer[e] : e;
    if (to < 0 || to >= new_count) Fail("GPCSP Reindexer is not valid for GPEngine size.");
    nq[to] = q[e]; ninv[to] = inv[e]; nbl[to] = bl[e]; ndiff[to] = diff[e]; nhyb[to] = hyb[e];
    nlls[to] = lls[e];
    nrows[to] = rows_[e];
    moved[e] = 1;
  }
  for (int64_t j = 0; j < spare_gpcsps_; ++j) {
    const int64_t src = old_count + j, dst = new_count + j;
    nq[dst] = q[src]; ninv[dst] = inv[src]; nbl[dst] = bl[src]; ndiff[dst] = diff[src];
    nhyb[dst] = hyb[src];
    nlls[dst] = lls[src];
    nrows[dst] = rows_[src];
    moved[src] = 1;
  }
  for (int64_t e = 0; e < old_padded; ++e)
    if (!moved[e] && rows_[e] != nullptr) row_pool_.Free(rows_[e]);
  nlls[new_padded] = marg;
  rows_ = std::move(nrows);
  auto push = [&](DeviceArray<double>& a, const std::vector<double>& v) {
    a.Resize(v.size(), false, stream_);
    GP_CUDA(cudaMemcpyAsync(a.ptr, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice,
                            stream_));
  };
  // The marginal slot is the LAST element of d_ll_sum_: keep n exact.
  if (d_ll_sum_.n != static_cast<size_t>(new_padded + 1)) {
    d_ll_sum_.Release(); d_q_.Release(); d_inverted_.Release(); d_bl_.Release(); d_diff_.Release();
    d_hybrid_.Release();
  }
  push(d_q_, nq); push(d_inverted_, ninv); push(d_bl_, nbl); push(d_diff_, ndiff);
  push(d_hybrid_, nhyb); push(d_ll_sum_, nlls);
  GP_CUDA(cudaStreamSynchronize(stream_));
  InvalidatePrograms();
}

void Engine::GrowSparePlvs(int64_t new_spare) {
  if (new_spare <= spare_nodes_) return;
  Activate();
  const int64_t old_padded = padded_plv_count();
  spare_nodes_ = new_spare;
  plvs_.resize(static_cast<size_t>(padded_plv_count()));
  GP_CUDA(cudaStreamSynchronize(stream_));
  d_counts_.Resize(static_cast<size_t>(padded_plv_count()), true, stream_);
  GP_CUDA(cudaMemsetAsync(d_counts_.ptr + old_padded, 0,
                          static_cast<size_t>(padded_plv_count() - old_padded) * sizeof(int32_t),
                          stream_));
  {
    const size_t old_n = d_uncond_.n;
    d_uncond_.Resize(static_cast<size_t>(node_count_ + spare_nodes_), true, stream_);
    if (d_uncond_.n > old_n) LaunchFill(stream_, d_uncond_.ptr + old_n, static_cast<int64_t>(d_uncond_.n - old_n), 1.0);
  }
  InvalidatePrograms();
}

void Engine::GrowSpareGpcsps(int64_t new_spare) {
  if (new_spare <= spare_gpcsps_) return;
  // Pull, extend with the reference defaults, push (values of existing edges kept).
  const int64_t keep = gpcsp_count_;
  const int64_t old_spare = spare_gpcsps_;
  Activate();
  auto extend = [&](DeviceArray<double>& a, int64_t old_n, int64_t new_n, double fill, bool tail_slot) {
    std::vector<double> v(static_cast<size_t>(old_n + (tail_slot ? 1 : 0)));
    GP_CUDA(cudaMemcpyAsync(v.data(), a.ptr, v.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
    std::vector<double> w(static_cast<size_t>(new_n + (tail_slot ? 1 : 0)), fill);
    std::copy(v.begin(), v.begin() + old_n, w.begin());
    if (tail_slot) w[new_n] = v[old_n];
    a.Release();
    a.Resize(w.size(), false, stream_);
    GP_CUDA(cudaMemcpyAsync(a.ptr, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, stream_));
    GP_CUDA(cudaStreamSynchronize(stream_));
  };
  const int64_t old_padded = keep + old_spare, new_padded = keep + new_spare;
  extend(d_q_, old_padded, new_padded, 1., false);
  extend(d_inverted_, old_padded, new_padded, 1., false);
  extend(d_bl_, old_padded, new_padded, kDefaultBranchLength, false);
  extend(d_diff_, old_padded, new_padded, 0., false);
  extend(d_hybrid_, old_padded, new_padded, -std::numeric_limits<double>::infinity(), false);
  extend(d_ll_sum_, old_padded, new_padded, 0., true);
  spare_gpcsps_ = new_spare;
  rows_.resize(static_cast<size_t>(new_padded), nullptr);
  InvalidatePrograms();
}

void Engine::CopyNodeData(int64_t src, int64_t dest) {  // gp_engine.cpp:384-390
  Activate();
  const int64_t padded = node_count_ + spare_nodes_;
  if (src < 0 || dest < 0 || src >= padded || dest >= padded)
    Fail("Cannot copy node data with src or dest index out-of-range.");
  if (src == dest) return;
  GP_CUDA(cudaMemcpyAsync(d_uncond_.ptr + dest, d_uncond_.ptr + src, sizeof(double),
                          cudaMemcpyDeviceToDevice, stream_));
}

void Engine::CopyPlvData(int64_t src, int64_t dest) {  // gp_engine.cpp:394-399
  Activate();
  if (src < 0 || dest < 0 || src >= padded_plv_count() || dest >= padded_plv_count())
    Fail("Cannot copy PLV data with src or dest index out-of-range.");
  if (src == dest) return;
  const PlvSlot s = plvs_[src];
  if (s.kind == kPlvZero) {
    if (plvs_[dest].kind == kPlvDense)
      GP_CUDA(cudaMemsetAsync(plvs_[dest].ptr, 0, static_cast<size_t>(32 * P_stride_), stream_));
    else if (plvs_[dest].kind == kPlvSymbols) {
      plvs_[dest].kind = kPlvZero;
      plvs_[dest].ptr = nullptr;
      InvalidatePrograms();
    }
  } else {
    if (plvs_[dest].kind != kPlvDense) {
      plvs_[dest].kind = kPlvZero;
      EnsureDense(dest);
    }
    if (s.kind == kPlvDense) {
      GP_CUDA(cudaMemcpyAsync(plvs_[dest].ptr, s.ptr, static_cast<size_t>(32 * P_stride_),
                              cudaMemcpyDeviceToDevice, stream_));
    } else {
      LaunchExportPlv(stream_, State(), Ref(src), static_cast<double*>(plvs_[dest].ptr));
    }
  }
  GP_CUDA(cudaMemcpyAsync(d_counts_.ptr + dest, d_counts_.ptr + src, sizeof(int32_t),
                          cudaMemcpyDeviceToDevice, stream_));
}

void Engine::CopyGpcspData(int64_t src, int64_t dest) {  // gp_engine.cpp:401-409
  Activate();
  if (src < 0 || dest < 0 || src >= padded_gpcsp_count() || dest >= padded_gpcsp_count())
    Fail("Cannot copy PLV data with src or dest index out-of-range.");
  auto cp = [&](DeviceArray<double>& a) {
    GP_CUDA(cudaMemcpyAsync(a.ptr + dest, a.ptr + src, sizeof(double), cudaMemcpyDeviceToDevice,
                            stream_));
  };
  cp(d_bl_);
  cp(d_q_);
  cp(d_inverted_);
}

// ---- per-kernel timing with CUDA events on the launching stream (bench.py's roofline) ---------------
const char* const kProfNames[kProfKinds] = {"k_zero", "k_scalar", "k_stationary", "k_prologue", "k_node",
                                            "k_rescale", "k_likelihood", "k_marginal", "k_reduce_partials",
                                            "k_opt_prepare", "k_opt_eval", "k_opt_step", "k_opt_block", "k_opt_cluster"};

ProfScope::ProfScope(Engine* e, int kind, double bytes) : e_(e->profiling_ ? e : nullptr) {
  if (e_ == nullptr) return;
  Engine::ProfEvent ev{};
  ev.kind = kind;
  ev.bytes = bytes;
  cudaEventCreate(&ev.begin);
  cudaEventCreate(&ev.end);
  cudaEventRecord(ev.begin, e_->stream_);
  e_->prof_events_.push_back(ev);
}
ProfScope::~ProfScope() {
  if (e_ != nullptr) cudaEventRecord(e_->prof_events_.back().end, e_->stream_);
}

void Engine::SetProfiling(bool on) {
  Activate();
  CollectProfile();
  profiling_ = on;
}

void Engine::CollectProfile() {
  if (prof_events_.empty()) return;
  GP_CUDA(cudaStreamSynchronize(stream_));
  for (ProfEvent& ev : prof_events_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.begin, ev.end) == cudaSuccess) {
      prof_[ev.kind].launches++;
      prof_[ev.kind].total_ms += ms;
      prof_[ev.kind].algorithmic_bytes += ev.bytes;
    }
    cudaEventDestroy(ev.begin);
    cudaEventDestroy(ev.end);
  }
  prof_events_.clear();
}

int Engine::GetKernelProfile(bito_gp_kernel_profile* out, int capacity) {
  Activate();
  CollectProfile();
  int n = 0;
  for (int k = 0; k < kProfKinds && n < capacity; ++k) {
    if (prof_[k].launches == 0) continue;
    out[n] = prof_[k];
    std::snprintf(out[n].name, sizeof out[n].name, "%s", kProfNames[k]);
    ++n;
  }
  return n;
}

void Engine::ResetKernelProfile() {
  CollectProfile();
  for (auto& p : prof_) p = bito_gp_kernel_profile{};
}

void Engine::GetStats(bito_gp_stats* out) {
  Activate();
  GP_CUDA(cudaStreamSynchronize(stream_));
  if (status_pending_) CheckStatus();
  if (timing_pending_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev_begin_, ev_end_) == cudaSuccess) stats_.last_process_ms = ms;
    timing_pending_ = false;
  }
  unsigned long long fevals[2] = {0, 0};
  GP_CUDA(cudaMemcpy(fevals, d_feval_total_.ptr, sizeof fevals, cudaMemcpyDeviceToHost));
  stats_.objective_evaluations = static_cast<int64_t>(fevals[0]);
  stats_.objective_passes = static_cast<int64_t>(fevals[1]);
  stats_.device_bytes_in_use =
      static_cast<int64_t>(plv_pool_.BytesReserved() + row_pool_.BytesReserved());
  int64_t resident = 0;
  for (const PlvSlot& s : plvs_) resident += (s.kind == kPlvDense);
  stats_.plvs_resident = resident;
  stats_.optimizer_scheme = last_opt_scheme_;
  stats_.optimizer_cluster_size = last_opt_scheme_ >= 2 ? last_opt_plan_.cluster_size : 0;
  stats_.optimizer_cluster_threads = last_opt_scheme_ >= 2 ? last_opt_plan_.threads : 0;
  stats_.optimizer_edges_in_flight = last_opt_scheme_ >= 2 ? last_opt_plan_.active_clusters : 0;
  stats_.programs_cached = static_cast<int64_t>(programs_.size());
  *out = stats_;
}

}  // namespace bito_gp
