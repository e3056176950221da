// tools/mix_bw.cu — what HBM delivers for the read:write mixes and stream counts of k_node.
//
// The roofline denominator of bench.py (MEASURED_PEAKS.json) is a 1:1 copy over two linear streams.
// k_node's levels are not that: a rootward node reads ~1 PLV tile and writes 3 (PHatRight, PHatLeft, P),
// a leafward one reads 2-4 and writes 3, and a level touches thousands of 4 MB PLVs at once (one 8 KB
// tile of each per block). This program times bare kernels with exactly those shapes - 256 threads per
// block, one 32-byte LDG.E.256 / STG.E.256 per thread and PLV, 8 consecutive tiles per block, blocks
// ordered tile-major over `nodes` independent "nodes" - so that k_node's GB/s can be read against the
// rate the machine reaches for ITS mix, not only against the copy peak.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mix_bw tools/mix_bw.cu
//   tools/_build/mix_bw [patterns=125000] [nodes=400]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct V4 { double a, b, c, d; };
__device__ __forceinline__ V4 ld256(const double* p) {
  V4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double* p, const V4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}

constexpr int kTile = 256;

// Node `o` owns R source PLVs and W destination PLVs of P patterns each, all distinct allocations inside
// one slab: PLV j of node o starts at slab + ((o * (R + W)) + j) * stride.
template <int R, int W>
__global__ void __launch_bounds__(kTile, 4)
    k_mix(double* __restrict__ slab, int64_t stride, int64_t P, int n_nodes, int tiles, int tiles_per_block) {
  const int tile_group = blockIdx.x / n_nodes;
  const int o = blockIdx.x - tile_group * n_nodes;
  double* base = slab + static_cast<int64_t>(o) * (R + W) * stride;
  const int t0 = tile_group * tiles_per_block;
  const int t1 = min(tiles, t0 + tiles_per_block);
  for (int tile = t0; tile < t1; ++tile) {
    const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
    if (p >= P) continue;
    V4 acc = {1., 2., 3., 4.};
    V4 x[R > 0 ? R : 1];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = ld256(base + r * stride + 4 * p);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc.a += x[r].a * 0.5; acc.b += x[r].b * 0.25; acc.c += x[r].c * 0.125; acc.d += x[r].d * 2.;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      V4 v = {acc.a + w, acc.b * (w + 1), acc.c - w, acc.d};
      st256(base + (R + w) * stride + 4 * p, v);
    }
  }
}


// ---- k_node-like dependencies ---------------------------------------------------------------------
// kind 0: level-1 shape: two 1-byte symbol loads per pattern (leaf children), results depend on them, 3 stores.
// kind 1: the same with the NEXT tile's symbols loaded before the current tile is finished (register prefetch).
// kind 2: 1 dense read : 3 writes with `work` dependent integer instructions between load and stores (the
//         ~370 warp instructions per tile k_node executes).
// kind 3: kind 2 with the next tile's load in flight (register prefetch).
// kind 4: kind 2 with a 3-stage cp.async ring in shared memory (two tiles ahead, private 32-byte slots).
template <int KIND>
__global__ void __launch_bounds__(kTile, 4)
    k_dep(double* __restrict__ slab, const uint8_t* __restrict__ sym, int64_t stride, int64_t P, int n_nodes, int tiles,
          int tiles_per_block, int work, const int* __restrict__ chain = nullptr, unsigned long long* mx = nullptr) {
  __shared__ __align__(32) double ring[KIND == 4 ? 3 * kTile * 4 : 4];
  __shared__ int s_chain;
  __shared__ double s_red[kTile / 32];
  if (chain != nullptr) {  // k_node's block prologue: op record -> item record -> counts (dependent loads), staged through shared memory
    if (threadIdx.x < 64) {
      int i = chain[(blockIdx.x % n_nodes) * 16];
      i = chain[i * 16 + (threadIdx.x & 3)];
      i = chain[i * 16 + 1];
      if (threadIdx.x == 0) s_chain = i;
    }
    __syncthreads();
    work += s_chain & 1;
  }
  const int tile_group = blockIdx.x / n_nodes;
  const int o = blockIdx.x - tile_group * n_nodes;
  double* base = slab + static_cast<int64_t>(o) * 4 * stride;
  const uint8_t* s0 = sym + static_cast<int64_t>(2 * o) * (stride / 4);
  const uint8_t* s1 = s0 + stride / 4;
  const int t0 = tile_group * tiles_per_block;
  const int t1 = min(tiles, t0 + tiles_per_block);
  auto finish = [&](V4 acc, int64_t p) {
    long long h = __double_as_longlong(acc.a);
    for (int k = 0; k < work; ++k) h = h * 6364136223846793005ll + 1442695040888963407ll;
    acc.d += (h == 42 ? 1. : 0.);
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      V4 v = {acc.a + w, acc.b * (w + 1), acc.c - w, acc.d};
      st256(base + (1 + w) * stride + 4 * p, v);
    }
  };
  if (KIND == 0 || KIND == 1) {
    int a = 0, b = 0;
    if (KIND == 1) {
      const int64_t p = static_cast<int64_t>(t0) * kTile + threadIdx.x;
      if (p < P) { a = s0[p]; b = s1[p]; }
    }
    for (int tile = t0; tile < t1; ++tile) {
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      int na = 0, nb = 0;
      if (KIND == 0) {
        if (p < P) { a = s0[p]; b = s1[p]; }
      } else {
        const int64_t pn = p + kTile;
        if (tile + 1 < t1 && pn < P) { na = s0[pn]; nb = s1[pn]; }
      }
      if (p < P) {
        V4 acc = {a == 0 ? 1. : 0.5, a == 1 ? 1. : 0.25, b == 2 ? 1. : 0.125, b == 3 ? 1. : 2.};
        finish(acc, p);
      }
      if (KIND == 1) { a = na; b = nb; }
    }
  } else if (KIND == 2) {
    for (int tile = t0; tile < t1; ++tile) {
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      if (p < P) finish(ld256(base + 4 * p), p);
    }
  } else if (KIND == 3) {
    V4 cur = {0, 0, 0, 0};
    {
      const int64_t p = static_cast<int64_t>(t0) * kTile + threadIdx.x;
      if (p < P) cur = ld256(base + 4 * p);
    }
    for (int tile = t0; tile < t1; ++tile) {
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      V4 nxt = {0, 0, 0, 0};
      if (tile + 1 < t1 && p + kTile < P) nxt = ld256(base + 4 * (p + kTile));
      if (p < P) finish(cur, p);
      cur = nxt;
    }
  } else {
    const unsigned slot0 = static_cast<unsigned>(__cvta_generic_to_shared(ring)) + 32u * threadIdx.x;
    auto issue = [&](int tile) {
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      if (tile < t1 && p < P) {
        const unsigned dst = slot0 + static_cast<unsigned>((tile - t0) % 3) * (kTile * 32u);
        const double* src = base + 4 * p;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 2) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(t0);
    issue(t0 + 1);
    for (int tile = t0; tile < t1; ++tile) {
      issue(tile + 2);
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      if (p < P) {
        const double* q = ring + static_cast<size_t>((tile - t0) % 3) * (kTile * 4) + 4 * threadIdx.x;
        V4 x = {q[0], q[1], q[2], q[3]};
        finish(x, p);
      }
    }
  }
  if (mx != nullptr) {  // k_node's epilogue: block maximum -> atomicMax
    double v = static_cast<double>(threadIdx.x);
    for (int o2 = 16; o2 > 0; o2 >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o2));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      v = threadIdx.x < kTile / 32 ? s_red[threadIdx.x] : 0.;
      for (int o2 = 16; o2 > 0; o2 >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o2));
      if (threadIdx.x == 0) atomicMax(mx + (blockIdx.x % n_nodes), static_cast<unsigned long long>(__double_as_longlong(v)));
    }
  }
}

template <int KIND>
double RunDep(double* slab, const uint8_t* sym, int64_t stride, int64_t P, int nodes, int tpb, int work, int reps, int occ = 8,
              const int* chain = nullptr, unsigned long long* mx = nullptr) {
  // cap the resident blocks per SM with dynamic shared memory (k_node runs 4 blocks of 256 threads per SM)
  const size_t dyn = occ >= 8 ? 0 : static_cast<size_t>(220 * 1024 / occ - (KIND == 4 ? 25 * 1024 : 1024));
  CK(cudaFuncSetAttribute(k_dep<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
  const int tiles = static_cast<int>((P + kTile - 1) / kTile);
  const unsigned grid = static_cast<unsigned>(nodes) * ((tiles + tpb - 1) / tpb);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) k_dep<KIND><<<grid, kTile, dyn>>>(slab, sym, stride, P, nodes, tiles, tpb, work, chain, mx);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_dep<KIND><<<grid, kTile, dyn>>>(slab, sym, stride, P, nodes, tiles, tpb, work, chain, mx);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double per_pattern = KIND <= 1 ? 3 * 32. + 2 : 4 * 32.;
  return per_pattern * P * nodes * reps / (ms * 1e-3) / 1e9;
}

template <int R, int W>
double Run(double* slab, int64_t stride, int64_t P, int nodes, int tpb, int reps) {
  const int tiles = static_cast<int>((P + kTile - 1) / kTile);
  const unsigned grid = static_cast<unsigned>(nodes) * ((tiles + tpb - 1) / tpb);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) k_mix<R, W><<<grid, kTile>>>(slab, stride, P, nodes, tiles, tpb);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_mix<R, W><<<grid, kTile>>>(slab, stride, P, nodes, tiles, tpb);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = 32. * P * (R + W) * nodes * reps;
  return bytes / (ms * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
  const int64_t P = argc > 1 ? atoll(argv[1]) : 125000;
  const int nodes_max = argc > 2 ? atoi(argv[2]) : 400;
  const int64_t stride = ((P + 7) / 8 * 8) * 4;  // doubles per PLV
  const int max_plvs = 7;
  const size_t bytes = static_cast<size_t>(nodes_max) * max_plvs * stride * sizeof(double);
  double* slab = nullptr;
  CK(cudaMalloc(&slab, bytes));
  CK(cudaMemset(slab, 0, bytes));
  printf("# P=%lld patterns per PLV (%.2f MB), up to %d nodes, slab %.1f GB; GB/s = (R+W) x 32 B x P x nodes / time\n",
         static_cast<long long>(P), stride * 8 / 1e6, nodes_max, bytes / 1e9);
  printf("| reads:writes per pattern | nodes | tiles/block | GB/s |\n|---|---|---|---|\n");
  const int reps = 5;
  for (int nodes : {nodes_max, 40, 1}) {
    for (int tpb : {8, 1}) {
      if (nodes == 1 && tpb == 8) continue;
      const int n = nodes;
#define ROW(R, W) printf("| %d:%d | %d | %d | %.0f |\n", R, W, n, tpb, Run<R, W>(slab, stride, P, n, tpb, reps)); fflush(stdout)
      ROW(1, 1);
      ROW(1, 3);
      ROW(2, 3);
      ROW(4, 3);
      ROW(2, 1);
      ROW(0, 3);
      ROW(3, 0);
#undef ROW
    }
  }
  // one node's PLVs as big as the whole slab: the linear-stream limit (what MEASURED_PEAKS times)
  {
    const int64_t bigP = static_cast<int64_t>(nodes_max) * (stride / 4) / 2 * 1;  // 2 PLVs of half the slab... per (R+W)
    const int64_t bstride = bigP * 4;
    printf("| 1:1 linear, one pair of %.1f GB streams | 1 | 8 | %.0f |\n", bstride * 8 / 1e9,
           Run<1, 1>(slab, bstride, bigP, 1, 8, reps));
    const int64_t P4 = bigP / 2;
    printf("| 1:3 linear, four %.1f GB streams | 1 | 8 | %.0f |\n", P4 * 32 / 1e9, Run<1, 3>(slab, P4 * 4, P4, 1, 8, reps));
  }
  {
    uint8_t* sym = nullptr;
    const size_t sym_bytes = static_cast<size_t>(2) * nodes_max * (stride / 4);
    CK(cudaMalloc(&sym, sym_bytes));
    CK(cudaMemset(sym, 1, sym_bytes));
    printf("\n| k_node-like dependency (400 nodes) | tiles/block | work (int instr pairs) | GB/s |\n|---|---|---|---|\n");
    for (int tpb : {8, 16}) {
#define DROW(K, W, LABEL) printf("| %s | %d | %d | %.0f |\n", LABEL, tpb, W, RunDep<K>(slab, sym, stride, P, nodes_max, tpb, W, reps)); fflush(stdout)
      DROW(0, 0, "0:3 + two symbol bytes, load then store");
      DROW(1, 0, "0:3 + two symbol bytes, next tile prefetched in registers");
      DROW(0, 100, "0:3 + two symbol bytes, load then store");
      DROW(1, 100, "0:3 + two symbol bytes, next tile prefetched in registers");
      DROW(2, 0, "1:3 load then store");
      DROW(2, 100, "1:3 load then store");
      DROW(2, 200, "1:3 load then store");
      DROW(3, 100, "1:3 next tile prefetched in registers");
      DROW(3, 200, "1:3 next tile prefetched in registers");
      DROW(4, 100, "1:3 cp.async ring, two tiles ahead");
      DROW(4, 200, "1:3 cp.async ring, two tiles ahead");
#undef DROW
    }
    printf("\n| resident blocks per SM (x 256 threads) | shape | GB/s |\n|---|---|---|\n");
    for (int occ : {3, 4, 5, 6, 8}) {
      printf("| %d | 0:3 + two symbol bytes, load then store | %.0f |\n", occ, RunDep<0>(slab, sym, stride, P, nodes_max, 8, 100, reps, occ));
      printf("| %d | 0:3 + two symbol bytes, next tile in registers | %.0f |\n", occ, RunDep<1>(slab, sym, stride, P, nodes_max, 8, 100, reps, occ));
      printf("| %d | 1:3 load then store | %.0f |\n", occ, RunDep<2>(slab, sym, stride, P, nodes_max, 8, 100, reps, occ));
      printf("| %d | 1:3 next tile in registers | %.0f |\n", occ, RunDep<3>(slab, sym, stride, P, nodes_max, 8, 100, reps, occ));
      printf("| %d | 1:3 cp.async ring, two tiles ahead | %.0f |\n", occ, RunDep<4>(slab, sym, stride, P, nodes_max, 8, 100, reps, occ));
      fflush(stdout);
    }
    // block prologue (three dependent small loads + barrier) and epilogue (block max + atomicMax) of k_node
    int* chain = nullptr;
    unsigned long long* mx = nullptr;
    CK(cudaMalloc(&chain, sizeof(int) * 16 * nodes_max));
    CK(cudaMalloc(&mx, sizeof(unsigned long long) * nodes_max));
    CK(cudaMemset(mx, 0, sizeof(unsigned long long) * nodes_max));
    {
      std::vector<int> h(16 * static_cast<size_t>(nodes_max));
      for (int i = 0; i < nodes_max; ++i)
        for (int k = 0; k < 16; ++k) h[16 * static_cast<size_t>(i) + k] = static_cast<int>((i * 7919ll + k * 104729ll + 13) % nodes_max);
      CK(cudaMemcpy(chain, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    printf("\n| 4 blocks per SM, with k_node's block prologue and epilogue | tiles/block | GB/s |\n|---|---|---|\n");
    for (int tpb : {4, 8, 16}) {
      printf("| 1:3 load then store, bare | %d | %.0f |\n", tpb, RunDep<2>(slab, sym, stride, P, nodes_max, tpb, 100, reps, 4));
      printf("| 1:3 load then store, prologue | %d | %.0f |\n", tpb, RunDep<2>(slab, sym, stride, P, nodes_max, tpb, 100, reps, 4, chain, nullptr));
      printf("| 1:3 load then store, prologue + epilogue | %d | %.0f |\n", tpb, RunDep<2>(slab, sym, stride, P, nodes_max, tpb, 100, reps, 4, chain, mx));
      printf("| 0:3 + symbols, prologue + epilogue | %d | %.0f |\n", tpb, RunDep<0>(slab, sym, stride, P, nodes_max, tpb, 100, reps, 4, chain, mx));
      fflush(stdout);
    }
  }
  return 0;
}
