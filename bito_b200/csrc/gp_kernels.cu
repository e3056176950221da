// bito_b200/csrc/gp_kernels.cu — hand-written sm_100a kernels of the GP likelihood engine.
//
// Every kernel is a streaming pass over pattern-contiguous FP64 PLVs: one thread owns one
// site pattern (4 states = one 32-byte LDG.E.256 / STG.E.256), a thread block owns a tile
// of kTile patterns of ONE macro-op, and a launch covers every macro-op of a dependency
// level. The 4x4 transition matrices are built on the device from the edge's branch
// length into shared memory. HBM-bound by construction (~0.6 flop/B):
// tensor cores are deliberately not used.
#include "gp_kernels.h"

#include <cfloat>
#include <cmath>
#include <cstdlib>

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace bito_gp {

__constant__ ModelConst c_model;

cudaError_t UploadModel(const ModelConst& model) {
  return cudaMemcpyToSymbol(c_model, &model, sizeof(ModelConst));
}

namespace {

struct V4 {
  double a, b, c, d;
};

// 256-bit global accesses (PTX ISA 8.8, sm_100+): one request per pattern.
__device__ __forceinline__ V4 ld256(const double* p) {
  V4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st256(double* p, const V4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c),
               "d"(v.d)
               : "memory");
}

// Same store with the cache-streaming hint: a PLV written by a level is next read a whole level
// (gigabytes of traffic) later, so it should not push the level's shared source tiles out of L2.
__device__ __forceinline__ void st256cs(double* p, const V4& v) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c),
               "d"(v.d)
               : "memory");
}
template <bool kStream>
__device__ __forceinline__ void st256p(double* p, const V4& v) {
  if (kStream) st256cs(p, v); else st256(p, v);
}

__device__ __forceinline__ V4 load_plv(const PlvRef& r, int64_t p) {
  if (r.kind == kPlvDense) return ld256(static_cast<const double*>(r.ptr) + 4 * p);
  V4 v = {0., 0., 0., 0.};
  if (r.kind == kPlvSymbols) {
    // InitializePLVsWithSitePatterns, gp_engine.cpp:544-562
    const int s = static_cast<const uint8_t*>(r.ptr)[p];
    const bool gap = (s == 4);
    v.a = (gap || s == 0) ? 1. : 0.;
    v.b = (gap || s == 1) ? 1. : 0.;
    v.c = (gap || s == 2) ? 1. : 0.;
    v.d = (gap || s == 3) ? 1. : 0.;
  }
  return v;
}

// M = scale * ((V * diag(f_k)) * V^-1), f_k = lambda_k^deriv * exp(lambda_k t)
// (gp_engine.cpp:341-358; evaluation order (V*D)*V^-1 as in Eigen).
__device__ void build_matrix(double t, int deriv, double scale, double* out16) {
  double d[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double l = c_model.lambda[k];
    const double e = exp(t * l);
    d[k] = deriv == 0 ? e : (deriv == 1 ? l * e : l * l * e);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = 0.;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += (c_model.V[4 * i + k] * d[k]) * c_model.Vinv[4 * k + j];
      out16[4 * i + j] = scale * s;
    }
}

__device__ __forceinline__ V4 matvec(const double* M, const V4& x) {
  V4 y;
  y.a = M[0] * x.a + M[1] * x.b + M[2] * x.c + M[3] * x.d;
  y.b = M[4] * x.a + M[5] * x.b + M[6] * x.c + M[7] * x.d;
  y.c = M[8] * x.a + M[9] * x.b + M[10] * x.c + M[11] * x.d;
  y.d = M[12] * x.a + M[13] * x.b + M[14] * x.c + M[15] * x.d;
  return y;
}

// r^T M p
__device__ __forceinline__ double quad(const V4& r, const double* M, const V4& p) {
  const double c0 = r.a * M[0] + r.b * M[4] + r.c * M[8] + r.d * M[12];
  const double c1 = r.a * M[1] + r.b * M[5] + r.c * M[9] + r.d * M[13];
  const double c2 = r.a * M[2] + r.b * M[6] + r.c * M[10] + r.d * M[14];
  const double c3 = r.a * M[3] + r.b * M[7] + r.c * M[11] + r.d * M[15];
  return c0 * p.a + c1 * p.b + c2 * p.c + c3 * p.d;
}

// NumericalUtils::LogAdd, numerical_utils.hpp:35-52
__device__ __forceinline__ double log_add(double x, double y) {
  if (y > x) {
    const double t = x;
    x = y;
    y = t;
  }
  if (x == -INFINITY) return x;
  const double neg_diff = y - x;
  if (neg_diff < -36.04365338911715 /* log(DBL_EPSILON) */) return x;
  return x + log(1.0 + exp(neg_diff));
}

// Deterministic block reductions (shuffle tree, then warp 0 over the warp leaders).
template <typename Op>
__device__ __forceinline__ double block_reduce(double v, Op op, double identity) {
  __shared__ double s_part[kTile / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect s_part across back-to-back calls
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? s_part[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;  // valid in thread 0
}
struct SumOp {
  __device__ double operator()(double a, double b) const { return a + b; }
};
struct MaxOp {
  __device__ double operator()(double a, double b) const { return a > b ? a : b; }
};

// ---- transition-matrix tables ------------------------------------------------------------------
// q[e] * M(t_e) for every IncrementWithWeightedEvolvedPLV of a program and M(t_e) for every
// Likelihood (gp_engine.cpp:229-249, 287-291, 341-344), built once per program execution (again
// after any level that changes branch lengths or q), so that no pattern tile ever waits on exp().
// The rescaling factor thr^(count[src]-count[dest]) is applied where the matrix is used.
__global__ void k_build_matrices(DeviceState st, const AccumItem* __restrict__ items, int n_items,
                                 const LikOp* __restrict__ liks, int n_liks,
                                 double* __restrict__ mtab, double* __restrict__ mtab_lik) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_items) {
    const int e = items[i].edge;
    build_matrix(st.bl[e], 0, st.q[e], mtab + 16 * static_cast<int64_t>(i));
  } else if (i < n_items + n_liks) {
    const int j = i - n_items;
    build_matrix(st.bl[liks[j].edge], 0, 1., mtab_lik + 16 * static_cast<int64_t>(j));
  }
}

__device__ __forceinline__ V4 load_item(const void* ptr, int kind, int64_t p) {
  PlvRef r;
  r.ptr = ptr;
  r.kind = kind;
  r.id = 0;
  return load_plv(r, p);
}

// acc += sum_{i in [b, e)} sM[i] * src_i[p], four independent 32-byte loads in flight.
__device__ __forceinline__ void accumulate_items(V4& acc, const double (*sM)[16],
                                                 const void* const* s_ptr, const int* s_kind,
                                                 int b, int e, int64_t p) {
  int i = b;
  for (; i + 4 <= e; i += 4) {
    const V4 x0 = load_item(s_ptr[i], s_kind[i], p);
    const V4 x1 = load_item(s_ptr[i + 1], s_kind[i + 1], p);
    const V4 x2 = load_item(s_ptr[i + 2], s_kind[i + 2], p);
    const V4 x3 = load_item(s_ptr[i + 3], s_kind[i + 3], p);
    const V4 y0 = matvec(sM[i], x0);
    acc.a += y0.a; acc.b += y0.b; acc.c += y0.c; acc.d += y0.d;
    const V4 y1 = matvec(sM[i + 1], x1);
    acc.a += y1.a; acc.b += y1.b; acc.c += y1.c; acc.d += y1.d;
    const V4 y2 = matvec(sM[i + 2], x2);
    acc.a += y2.a; acc.b += y2.b; acc.c += y2.c; acc.d += y2.d;
    const V4 y3 = matvec(sM[i + 3], x3);
    acc.a += y3.a; acc.b += y3.b; acc.c += y3.c; acc.d += y3.d;
  }
  for (; i < e; ++i) {
    const V4 x0 = load_item(s_ptr[i], s_kind[i], p);
    const V4 y0 = matvec(sM[i], x0);
    acc.a += y0.a; acc.b += y0.b; acc.c += y0.c; acc.d += y0.d;
  }
}

// ---- node macro-ops: accumulate groups fused with the Multiplies that consume them -------------
// Group g: dest (+)= sum_i Mtab_i * src_i in op-list order (gp_engine.cpp:229-249); Multiply:
// dest = s1 o s2 with operands taken from registers when a group of this node produced them
// (gp_engine.cpp:278-285), plus the per-PLV maximum for the rescale decision (:583-597).
// Blocks are ordered tile-major (all macro-ops of one pattern tile are neighbours in the grid), so
// a PLV tile read by several macro-ops of the level is served from L2 after its first use.
template <int kMinBlocks, bool kStream>
__global__ void __launch_bounds__(kTile, kMinBlocks)
    k_node(DeviceState st, const NodeOp* __restrict__ nodes, const AccumItem* __restrict__ items,
           const int32_t* __restrict__ pool, const double* __restrict__ mtab, int n_nodes, int tiles,
           int tiles_per_block, unsigned long long* __restrict__ level_max) {
  const int tile_group = blockIdx.x / n_nodes;
  const int o = blockIdx.x - tile_group * n_nodes;
  const NodeOp* nd = nodes + o;
  __shared__ __align__(32) double sM[kItemChunk][16];
  __shared__ const void* s_ptr[kItemChunk];
  __shared__ int s_kind[kItemChunk];
  __shared__ int s_src_count[kItemChunk];
  __shared__ int s_gcount[2];

  const int ng = nd->n_groups, nm = nd->n_mults;
  const int n0 = ng > 0 ? nd->g[0].n_items : 0;
  const int n_tot = n0 + (ng > 1 ? nd->g[1].n_items : 0);
  const int item_base = ng > 0 ? nd->g[0].item_off : 0;
  double* const dest0 = ng > 0 ? nd->g[0].dest : nullptr;
  double* const dest1 = ng > 1 ? nd->g[1].dest : nullptr;
  const bool keep0 = ng > 0 && !nd->g[0].init_zero;
  const bool keep1 = ng > 1 && !nd->g[1].init_zero;

  // Rescaling counts (the last warp; every block derives them, the block of tile group 0
  // publishes them): count[dest] of a group is 0 after a folded ZeroPLV, the minimum over the
  // src_vector after a folded PrepForMarginalization (gp_engine.cpp:213-216, 323-333), else
  // unchanged; count[dest] of a Multiply = count[s1] + count[s2] (gp_engine.cpp:281-282).
  if (threadIdx.x >= kTile - 32) {
    const int lane = threadIdx.x & 31;
    int gc0 = 0, gc1 = 0;
    for (int gi = 0; gi < ng; ++gi) {
      const AccumGroup* grp = &nd->g[gi];
      const int mode = grp->count_mode;
      int c;
      if (mode == kCountPrep) {
        c = INT_MAX;
        for (int i = lane; i < grp->prep_len; i += 32) c = min(c, st.counts[pool[grp->prep_off + i]]);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, sh));
      } else if (mode == kCountZero) {
        c = 0;
      } else {
        c = st.counts[grp->dest_id];
      }
      if (gi == 0) gc0 = c; else gc1 = c;
    }
    if (lane == 0) {
      s_gcount[0] = gc0;
      s_gcount[1] = gc1;
      if (tile_group == 0) {
        // Multiplies first: an in-place product reads its own old count.
        int mc[2] = {0, 0};
        for (int mi = 0; mi < nm; ++mi) {
          const NodeMult* m = &nd->m[mi];
          const int c1 = m->s1_group == 0 ? gc0 : (m->s1_group == 1 ? gc1 : st.counts[m->s1.id]);
          const int c2 = m->s2_group == 0 ? gc0 : (m->s2_group == 1 ? gc1 : st.counts[m->s2.id]);
          mc[mi] = c1 + c2;
        }
        if (ng > 0 && nd->g[0].count_mode != kCountKeep) st.counts[nd->g[0].dest_id] = gc0;
        if (ng > 1 && nd->g[1].count_mode != kCountKeep) st.counts[nd->g[1].dest_id] = gc1;
        for (int mi = 0; mi < nm; ++mi) st.counts[nd->m[mi].dest_id] = mc[mi];
      }
    }
  }

  // Stage (a chunk of) the node's transition matrices and sources: thread t moves quarter t&3 of
  // item t>>2.
  auto stage = [&](int base, int n) {
    const int it = threadIdx.x >> 2, quarter = threadIdx.x & 3;
    if (it < n) {
      const int64_t gi = item_base + base + it;
      const V4 m = ld256(mtab + 16 * gi + 4 * quarter);
      double* dst = &sM[it][4 * quarter];
      dst[0] = m.a; dst[1] = m.b; dst[2] = m.c; dst[3] = m.d;
      if (quarter == 0) {
        const PlvRef src = items[gi].src;
        s_ptr[it] = src.ptr;
        s_kind[it] = src.kind;
        const int dest_id = base + it < n0 ? nd->g[0].dest_id : nd->g[1].dest_id;
        // an increment that reads its own destination sees the count just set for it
        s_src_count[it] = src.id == dest_id ? INT_MIN : st.counts[src.id];
      }
    }
  };
  // Rescaling factor thr^(count[src] - count[dest]) (gp_engine.cpp:236-242): after the barrier
  // that follows stage(), fold it into the staged matrix. Almost always 1 (nothing to do).
  auto apply_factors = [&](int base, int n) {
    const int it = threadIdx.x >> 2, quarter = threadIdx.x & 3;
    if (it < n) {
      const int c = s_gcount[base + it < n0 ? 0 : 1];
      const int sc = s_src_count[it];
      const int diff = sc == INT_MIN ? 0 : sc - c;
      if (diff != 0) {
        if (diff < 0 && tile_group == 0 && quarter == 0) atomicOr(st.status, kErrRescalingDifference);
        const double factor = pow(st.thr, static_cast<double>(diff));
        double* dst = &sM[it][4 * quarter];
        dst[0] *= factor; dst[1] *= factor; dst[2] *= factor; dst[3] *= factor;
      }
    }
  };
  const bool single_chunk = n_tot <= kItemChunk;
  if (n_tot > 0 && single_chunk) stage(0, n_tot);
  __syncthreads();
  bool any_factor = false;
  if (n_tot > 0 && single_chunk) {
    // block-uniform decision, so the extra barrier is only paid when some factor is not 1
    for (int i = 0; i < n_tot; ++i) {
      const int sc = s_src_count[i];
      any_factor |= (sc != INT_MIN && sc != s_gcount[i < n0 ? 0 : 1]);
    }
    if (any_factor) {
      apply_factors(0, n_tot);
      __syncthreads();
    }
  }

  double mx0 = 0., mx1 = 0.;
  const int tile_begin = tile_group * tiles_per_block;
  const int tile_end = min(tiles, tile_begin + tiles_per_block);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
    const bool live = p < st.P;
    V4 acc0 = {0., 0., 0., 0.}, acc1 = {0., 0., 0., 0.};
    if (live && keep0) acc0 = ld256(dest0 + 4 * p);
    if (live && keep1) acc1 = ld256(dest1 + 4 * p);
    if (single_chunk) {
      if (live) {
        accumulate_items(acc0, sM, s_ptr, s_kind, 0, n0, p);
        accumulate_items(acc1, sM, s_ptr, s_kind, n0, n_tot, p);
      }
    } else {
      for (int base = 0; base < n_tot; base += kItemChunk) {
        const int n = min(kItemChunk, n_tot - base);
        __syncthreads();  // the previous chunk (or tile) is done with sM
        stage(base, n);
        __syncthreads();
        apply_factors(base, n);
        __syncthreads();
        if (live) {
          // chunk positions [0, n) hold items [base, base + n) of the node; the first n0 items
          // of the node belong to group 0.
          const int split = max(0, min(n, n0 - base));
          accumulate_items(acc0, sM, s_ptr, s_kind, 0, split, p);
          accumulate_items(acc1, sM, s_ptr, s_kind, split, n, p);
        }
      }
    }
    if (live) {
      if (ng > 0) st256p<kStream>(dest0 + 4 * p, acc0);
      if (ng > 1) st256p<kStream>(dest1 + 4 * p, acc1);
      for (int mi = 0; mi < nm; ++mi) {
        const NodeMult* m = &nd->m[mi];
        const int g1 = m->s1_group, g2 = m->s2_group;
        const V4 x = g1 == 0 ? acc0 : (g1 == 1 ? acc1 : load_plv(m->s1, p));
        const V4 y = g2 == 0 ? acc0 : (g2 == 1 ? acc1 : load_plv(m->s2, p));
        const V4 v = {x.a * y.a, x.b * y.b, x.c * y.c, x.d * y.d};
        st256p<kStream>(m->dest + 4 * p, v);
        const double hi = fmax(fmax(v.a, v.b), fmax(v.c, v.d));
        const double lo = fmin(fmin(v.a, v.b), fmin(v.c, v.d));
        // AssertPLVIsFinite (:575-577), non-negativity (:585-586). fmax/fmin drop NaNs, so
        // test the sum as well.
        if (!isfinite(v.a + v.b + v.c + v.d)) atomicOr(st.status, kErrMultiplyNotFinite);
        else if (lo < 0.) atomicOr(st.status, kErrNegativePLV);
        if (mi == 0) mx0 = fmax(mx0, hi); else mx1 = fmax(mx1, hi);
      }
    }
  }
  for (int mi = 0; mi < nm; ++mi) {
    const double mx = block_reduce(mi == 0 ? mx0 : mx1, MaxOp(), 0.);
    // Non-negative doubles order like their bit patterns.
    if (threadIdx.x == 0 && mx > 0.)
      atomicMax(level_max + nd->m[mi].max_slot,
                static_cast<unsigned long long>(__double_as_longlong(mx)));
  }
}

// RescalePLVIfNeeded + RescalePLV (gp_engine.cpp:564-573, 583-597). The maximum is over the
// whole PLV (all patterns, all ranks). Almost always there is nothing to do, so the grid is small
// and fixed. Every block scans the level's maxima (level_max[o] belongs to ops[o]) in chunks of
// kRescaleChunk Multiplies and compacts, in op order, the ones that need rescaling into shared
// memory; all blocks then share the (PLV, pattern tile) items of that list, so a PLV that does
// need it (deep DAGs: the levels next to the rootsplits) is rescaled by the whole machine, not by
// one block.
constexpr int kRescaleChunk = 4 * kTile;
__global__ void __launch_bounds__(kTile)
    k_rescale(DeviceState st, const MultOp* __restrict__ ops, int n_ops,
              const double* __restrict__ level_max, int tiles) {
  __shared__ int s_op[kRescaleChunk];
  __shared__ double s_divisor[kRescaleChunk];
  __shared__ int s_warp_count[kTile / 32];
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int chunk0 = 0; chunk0 < n_ops; chunk0 += kRescaleChunk) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int base = chunk0; base < min(n_ops, chunk0 + kRescaleChunk); base += kTile) {
      const int o = base + threadIdx.x;
      int rescaling_count = 0;
      if (o < n_ops) {
        double max_entry = level_max[o];
        if (max_entry != 0.) {  // max == 0: no rescale (gp_engine.cpp:587-589)
          while (max_entry < st.thr) {
            max_entry /= st.thr;
            rescaling_count++;
          }
        }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, rescaling_count > 0);
      if (lane == 0) s_warp_count[warp] = __popc(ballot);
      __syncthreads();
      int at = s_n;
      for (int w = 0; w < warp; ++w) at += s_warp_count[w];
      if (rescaling_count > 0) {
        at += __popc(ballot & ((1u << lane) - 1u));
        s_op[at] = o;
        s_divisor[at] = pow(st.thr, static_cast<double>(rescaling_count));
        if (blockIdx.x == 0) st.counts[ops[o].dest_id] += rescaling_count;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int total = s_n;
        for (int w = 0; w < kTile / 32; ++w) total += s_warp_count[w];
        s_n = total;
      }
      __syncthreads();
    }
    const int64_t n_items = static_cast<int64_t>(s_n) * tiles;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int k = static_cast<int>(item / tiles);
      const int tile = static_cast<int>(item - static_cast<int64_t>(k) * tiles);
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      if (p < st.P) {
        double* dest = ops[s_op[k]].dest;
        const double divisor = s_divisor[k];
        V4 v = ld256(dest + 4 * p);
        v.a /= divisor; v.b /= divisor; v.c /= divisor; v.d /= divisor;
        st256(dest + 4 * p, v);
      }
    }
    __syncthreads();  // the list is rebuilt for the next chunk
  }
}

// ---- Likelihood: row[p] = log(parent^T M(t_e) child) + (count_p + count_c) log thr -------
// (gp_engine.cpp:287-291, gp_engine.hpp:273-282); also the weighted tile partial of the row.
__global__ void __launch_bounds__(kTile)
    k_likelihood(DeviceState st, const LikOp* __restrict__ ops, int n_ops,
                 const double* __restrict__ mtab, int tiles, int tiles_per_block,
                 double* __restrict__ partials) {
  const int tile_group = blockIdx.x / n_ops;
  const int o = blockIdx.x - tile_group * n_ops;
  const int n_groups = gridDim.x / n_ops;
  const LikOp* op = ops + o;
  __shared__ double sM[16];
  if (threadIdx.x < 16) sM[threadIdx.x] = mtab[16 * static_cast<int64_t>(o) + threadIdx.x];
  const PlvRef parent = op->parent, child = op->child;
  double* const row = op->row;
  const double resc =
      static_cast<double>(st.counts[parent.id]) * st.log_thr +
      static_cast<double>(st.counts[child.id]) * st.log_thr;
  __syncthreads();
  double wsum = 0.;
  const int tile_begin = tile_group * tiles_per_block;
  const int tile_end = min(tiles, tile_begin + tiles_per_block);
  // Two pattern tiles per trip: the four 32-byte loads are issued before the first log(), so a
  // thread keeps 128 B in flight (the kernel is bound by load latency, not by the FP64 pipe).
  // The weighted sum keeps the one-tile-at-a-time order.
  int tile = tile_begin;
  for (; tile + 2 <= tile_end; tile += 2) {
    const int64_t p0 = static_cast<int64_t>(tile) * kTile + threadIdx.x;
    const int64_t p1 = p0 + kTile;
    const bool live0 = p0 < st.P, live1 = p1 < st.P;
    V4 r0 = {0., 0., 0., 0.}, c0 = r0, r1 = r0, c1 = r0;
    double w0 = 0., w1 = 0.;
    if (live0) { r0 = load_plv(parent, p0); c0 = load_plv(child, p0); w0 = st.weights[p0]; }
    if (live1) { r1 = load_plv(parent, p1); c1 = load_plv(child, p1); w1 = st.weights[p1]; }
    if (live0) {
      const double ll = log(quad(r0, sM, c0)) + resc;
      if (row != nullptr) row[p0] = ll;
      wsum += ll * w0;
    }
    if (live1) {
      const double ll = log(quad(r1, sM, c1)) + resc;
      if (row != nullptr) row[p1] = ll;
      wsum += ll * w1;
    }
  }
  for (; tile < tile_end; ++tile) {
    const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
    if (p < st.P) {
      const V4 r = load_plv(parent, p);
      const V4 c = load_plv(child, p);
      const double ll = log(quad(r, sM, c)) + resc;
      if (row != nullptr) row[p] = ll;
      wsum += ll * st.weights[p];
    }
  }
  wsum = block_reduce(wsum, SumOp(), 0.);
  if (threadIdx.x == 0) partials[static_cast<int64_t>(o) * n_groups + tile_group] = wsum;
}

// ---- [ResetMarginalLikelihood] IncrementMarginalLikelihood x n (gp_engine.cpp:251-276) ----
// One launch handles the whole group in op order, so the per-pattern LogAdd chain is the
// reference's. partials: rows 0..n-1 = per-rootsplit conditional rows, row n = marginal.
__global__ void __launch_bounds__(kTile)
    k_marginal(DeviceState st, const MargItem* __restrict__ items, int n_items, int reset,
               int tiles, double* __restrict__ partials) {
  const int tile = blockIdx.x;
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  const bool live = p < st.P;
  double lm = -INFINITY;
  double w = 0.;
  if (live) {
    w = st.weights[p];
    if (!reset) lm = st.log_marg[p];
  }
  for (int i = 0; i < n_items; ++i) {
    const MargItem it = items[i];
    if (threadIdx.x == 0 && tile == 0 && st.counts[it.stationary.id] != 0)
      atomicOr(st.status, kErrRescaledStationary);
    const double resc = static_cast<double>(st.counts[it.p.id]) * st.log_thr;
    const double log_q = log(st.q[it.edge]);
    double wsum = 0.;
    if (live) {
      const V4 s = load_plv(it.stationary, p);
      const V4 x = load_plv(it.p, p);
      double row = log(s.a * x.a + s.b * x.b + s.c * x.c + s.d * x.d) + resc;
      lm = log_add(lm, row);
      row -= log_q;
      if (it.row != nullptr) it.row[p] = row;
      wsum = row * w;
    }
    wsum = block_reduce(wsum, SumOp(), 0.);
    if (threadIdx.x == 0) partials[static_cast<int64_t>(i) * tiles + tile] = wsum;
  }
  double msum = 0.;
  if (live) {
    st.log_marg[p] = lm;
    msum = lm * w;
  }
  msum = block_reduce(msum, SumOp(), 0.);
  if (threadIdx.x == 0) partials[static_cast<int64_t>(n_items) * tiles + tile] = msum;
}

// ---- SetToStationaryDistribution (gp_engine.cpp:218-227) ---------------------------------
__global__ void __launch_bounds__(kTile)
    k_stationary(DeviceState st, const StatOp* __restrict__ ops, int tiles) {
  const int o = blockIdx.x / tiles;
  const int tile = blockIdx.x - o * tiles;
  const StatOp op = ops[o];
  const double qv = st.q[op.edge];
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  if (p < st.P) {
    V4 v = {qv * c_model.pi[0], qv * c_model.pi[1], qv * c_model.pi[2], qv * c_model.pi[3]};
    st256(op.dest + 4 * p, v);
  }
  if (tile == 0 && threadIdx.x == 0) st.counts[op.dest_id] = 0;
}

// ---- ZeroPLV that could not be folded (gp_engine.cpp:213-216) -----------------------------
__global__ void __launch_bounds__(kTile)
    k_zero(DeviceState st, const ZeroOp* __restrict__ ops, int tiles) {
  const int o = blockIdx.x / tiles;
  const int tile = blockIdx.x - o * tiles;
  const ZeroOp op = ops[o];
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  if (p < st.P) {
    V4 v = {0., 0., 0., 0.};
    st256(op.dest + 4 * p, v);
  }
  if (tile == 0 && threadIdx.x == 0) st.counts[op.dest_id] = 0;
}

// ---- scalar-only ops: count resets, stand-alone Prep, UpdateSBNProbabilities -------------
__global__ void k_scalar(DeviceState st, const ScalarOp* __restrict__ ops,
                         const int32_t* __restrict__ pool, int n_ops) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_ops) return;
  const ScalarOp op = ops[o];
  if (op.kind == kScalarCountZero) {
    st.counts[op.a] = 0;
  } else if (op.kind == kScalarCountSum) {  // gp_engine.cpp:281-282 (max == 0: no rescale, :587-589)
    st.counts[op.a] = st.counts[op.b] + st.counts[op.vec_off];
  } else if (op.kind == kScalarPrep) {  // gp_engine.cpp:323-333
    if (op.vec_len <= 0) {
      atomicOr(st.status, kErrEmptyPrep);
      return;
    }
    int c = INT_MAX;
    for (int i = 0; i < op.vec_len; ++i) c = min(c, st.counts[pool[op.vec_off + i]]);
    st.counts[op.a] = c;
  } else if (op.kind == kScalarSbn) {  // gp_engine.cpp:297-321
    const int start = op.a, stop = op.b;
    if (stop - start == 1) {
      st.q[start] = 1.;
      return;
    }
    double hybrid_min = INFINITY;
    for (int e = start; e < stop; ++e) hybrid_min = fmin(hybrid_min, st.hybrid[e]);
    const double* ll = hybrid_min > -INFINITY ? st.hybrid : st.ll_sum;
    // NumericalUtils::LogSum is a left fold of LogAdd (numerical_utils.cpp:8).
    double norm = ll[start] + log(st.q[start]);
    for (int e = start + 1; e < stop; ++e) norm = log_add(norm, ll[e] + log(st.q[e]));
    for (int e = start; e < stop; ++e) st.q[e] = exp((ll[e] + log(st.q[e])) - norm);
  }
}

// ---- deterministic second stage of every reduction ------------------------------------------
__global__ void k_reduce_partials(const double* __restrict__ partials, int n_out, int64_t tiles,
                                  double* __restrict__ out, const int32_t* __restrict__ scatter_idx,
                                  double* __restrict__ scatter_dst) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_out) return;
  const double* row = partials + static_cast<int64_t>(warp) * tiles;
  double v = 0.;
  for (int64_t t = lane; t < tiles; t += 32) v += row[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) {
    out[warp] = v;
    if (scatter_idx != nullptr && scatter_idx[warp] >= 0) scatter_dst[scatter_idx[warp]] = v;
  }
}

__global__ void k_scatter(const double* __restrict__ packed, int n,
                          const int32_t* __restrict__ scatter_idx, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && scatter_idx[i] >= 0) dst[scatter_idx[i]] = packed[i];
}

// ---- OptimizeBranchLength ----------------------------------------------------------------------
// The objective of one edge is l(t) = sum_p w_p log(r_p^T V diag(e^{lambda t}) V^-1 p_p) + const.
// Its PLVs do not change during the 1-D search, so the pass below reads them ONCE (64 B per
// pattern) and keeps, per pattern, one coefficient per distinct eigenvalue:
//   coef[p][g] = sum_{k in g} (V^T r_p)_k (V^-1 p_p)_k,   L_p(t) = sum_g coef[p][g] e^{lambda_g t}
// (2 doubles per pattern for JC69). Every later evaluation of l, l', l'' streams only those.

enum OptPhase : int32_t {
  kPhBrentInit = 0,
  kPhBrentU = 1,
  kPhBrentGradGx = 2,
  kPhBrentGradU2 = 3,
  kPhGradientAscent = 4,
  kPhLogSpaceGradientAscent = 5,
  kPhNewton = 6
};

__device__ void opt_request(OptState& s, double x, bool log_space) {
  s.x_eval = x;
  s.t_eval = log_space ? exp(x) : x;
  // diag(e^{lambda t}) once per edge and request, not once per pattern (gp_engine.cpp:341-344).
  // This runs on ONE thread with every other thread of the block (cluster) waiting for it, so the
  // exponentials are issued side by side (fully unrolled) and the zero eigenvalue every reversible
  // model has costs nothing: exp(0 * t) and x / 1 are exact, so skipping them changes no bit.
  double e[kMaxEigenGroups];
#pragma unroll
  for (int g = 0; g < kMaxEigenGroups; ++g) {
    const double l = c_model.group_lambda[g];
    e[g] = (g < c_model.n_groups && l != 0.) ? exp(l * s.t_eval) : 1.0;
  }
#pragma unroll
  for (int g = 0; g < kMaxEigenGroups; ++g)
    if (g < c_model.n_groups) s.e[g] = e[g];
  s.x_ratio = e[0] == 1.0 ? e[1] : e[1] / e[0];  // two-eigenvalue ratio form, k_opt_eval_ratio
  s.evals++;
}

// dag_branch_handler.cpp:123-146 (+ the first objective request of :150-280).
__device__ void opt_init(OptState& s, const DeviceState& st, const OptParams& prm, int method,
                         const OptOp& op) {
  s.method = method;
  s.edge = op.edge;
  s.done = 0;
  s.speculative = 0;
  s.evals = 0;
  s.iter = 0;
  s.ll_offset = (static_cast<double>(st.counts[op.parent.id]) * st.log_thr +
                 static_cast<double>(st.counts[op.child.id]) * st.log_thr) *
                st.total_weight;
  if (prm.check_convergence && st.diff[op.edge] < prm.diff_threshold) {
    s.done = 1;
    return;
  }
  const double bl = st.bl[op.edge];
  switch (method) {
    case 0:  // BrentOptimization
    case 1:  // BrentOptimizationWithGradients
      s.cur_x = log(bl);
      s.phase = kPhBrentInit;
      opt_request(s, s.cur_x, true);
      break;
    case 2:  // GradientAscentOptimization
      s.cur_x = bl;
      s.x = bl;
      s.phase = kPhGradientAscent;
      opt_request(s, s.x, false);
      break;
    case 3:  // LogSpaceGradientAscentOptimization
      s.cur_x = bl;
      s.x = bl;
      s.phase = kPhLogSpaceGradientAscent;
      opt_request(s, s.x, false);
      break;
    default:  // NewtonOptimization
      s.cur_x = bl;
      s.x = log(bl);
      s.phase = kPhNewton;
      opt_request(s, s.x, true);
      break;
  }
}

__device__ void opt_finish_brent(OptState& s, const DeviceState& st) {
  // dag_branch_handler.cpp:168-176: keep the old value if the search made things worse.
  s.done = 1;
  if (s.speculative) return;
  const double old_bl = exp(s.cur_x);
  const double new_bl = (s.fx > s.cur_f) ? old_bl : exp(s.x);
  st.bl[s.edge] = new_bl;
  st.diff[s.edge] = fabs(old_bl - new_bl);
}

// Top of the do-while body of BrentMinimize up to the objective call (optimization.hpp:95-148).
__device__ void brent_next(OptState& s, const DeviceState& st, const OptParams& prm) {
  const double tolerance = prm.brent_tolerance;  // 2^(1 - significant digits), optimization.hpp:84
  const double golden = 0.3819660f;
  const double mid = (s.min + s.max) / 2;
  const double fract1 = tolerance * fabs(s.x) + tolerance / 4;
  const double fract2 = 2 * fract1;
  if (fabs(s.x - mid) <= (fract2 - (s.max - s.min) / 2)) {
    opt_finish_brent(s, st);
    return;
  }
  bool use_bisection = true;
  if (fabs(s.delta2) > fract1) {
    double r = (s.x - s.w) * (s.fx - s.fv);
    double q = (s.x - s.v) * (s.fx - s.fw);
    double p = (s.x - s.v) * q - (s.x - s.w) * r;
    q = 2 * (q - r);
    if (q > 0) p = -p;
    q = fabs(q);
    const double td = s.delta2;
    s.delta2 = s.delta;
    if (((fabs(p) >= fabs(q * td / 2)) == false) && ((p <= q * (s.min - s.x)) == false) &&
        ((p >= q * (s.max - s.x)) == false)) {
      s.delta = p / q;
      s.u = s.x + s.delta;
      if (((s.u - s.min) < fract2) || ((s.max - s.u) < fract2)) {
        s.delta = (mid - s.x) < 0 ? -fabs(fract1) : fabs(fract1);
      }
      use_bisection = false;
    }
  }
  if (use_bisection) {
    s.delta2 = (s.x >= mid) ? s.min - s.x : s.max - s.x;
    s.delta = golden * s.delta2;
  }
  s.u = (fabs(s.delta) >= fract1) ? (s.x + s.delta)
                                  : (s.delta > 0 ? (s.x + fabs(fract1)) : (s.x - fabs(fract1)));
  s.phase = kPhBrentU;
  opt_request(s, s.u, true);
}

__device__ void brent_accept(OptState& s, double u, double fu) {  // optimization.hpp:150-163
  if (u >= s.x) s.min = s.x; else s.max = s.x;
  s.v = s.w; s.w = s.x; s.x = u;
  s.fv = s.fw; s.fw = s.fx; s.fx = fu;
}
__device__ void brent_reject(OptState& s, double u, double fu) {  // optimization.hpp:164-182
  if (u < s.x) s.min = u; else s.max = u;
  if ((fu <= s.fw) || (s.w == s.x)) {
    s.v = s.w; s.w = u;
    s.fv = s.fw; s.fw = fu;
  } else if ((fu <= s.fv) || (s.v == s.x) || (s.v == s.w)) {
    s.v = u;
    s.fv = fu;
  }
}
__device__ void brent_loop_end(OptState& s, const DeviceState& st, const OptParams& prm) {
  if (--s.count) brent_next(s, st, prm); else opt_finish_brent(s, st);
}

// Consumes one objective evaluation (ll, dll/dt, d2ll/dt2 at t_eval) and advances the
// optimiser to its next request or to completion. Decision-for-decision restatement of
// optimization.hpp:71-402 driven as dag_branch_handler.cpp:150-280 drives it.
__device__ void opt_advance(OptState& s, const DeviceState& st, const OptParams& prm, double ll,
                            double d1, double d2) {
  switch (s.phase) {
    case kPhBrentInit: {
      const double f = -ll;
      s.cur_f = f;
      s.w = s.v = s.x = s.cur_x;
      s.fw = s.fv = s.fx = f;
      s.delta2 = s.delta = 0;
      s.min = prm.min_log_bl;
      s.max = prm.max_log_bl;
      s.count = prm.max_iter;
      s.evals++;  // the reference evaluates the starting point twice (:161-162 and :89)
      brent_next(s, st, prm);
      break;
    }
    case kPhBrentU: {
      s.fu = -ll;
      if (s.fu <= s.fx) {
        brent_accept(s, s.u, s.fu);
        brent_loop_end(s, st, prm);
      } else if (s.method == 1) {
        // BrentMinimizeWithGradients: try a gradient step from x first (:286-289).
        s.phase = kPhBrentGradGx;
        opt_request(s, s.x, true);
      } else {
        brent_reject(s, s.u, s.fu);
        brent_loop_end(s, st, prm);
      }
      break;
    }
    case kPhBrentGradGx: {
      // brent_grad_func returns (-ll, -t * dll/dt), gp_engine.cpp:613-622.
      const double f_prime_x = -s.t_eval * d1;
      s.u_alt = s.x - prm.log_step_size * f_prime_x;
      s.phase = kPhBrentGradU2;
      opt_request(s, s.u_alt, true);
      break;
    }
    case kPhBrentGradU2: {
      const double fu_alt = -ll;
      if (fu_alt <= s.fx) brent_accept(s, s.u_alt, fu_alt); else brent_reject(s, s.u, s.fu);
      brent_loop_end(s, st, prm);
      break;
    }
    case kPhGradientAscent: {  // optimization.hpp:331-345 (min_x is the LOG bound, as there)
      const double tolerance = prm.decimal_tolerance;  // 10^-significant digits
      const double new_x = s.x + d1 * prm.step_size;
      s.x = fmax(new_x, prm.min_log_bl);
      if (fabs(d1) < fabs(ll) * tolerance || s.iter >= prm.max_iter) {
        st.bl[s.edge] = s.x;
        st.diff[s.edge] = fabs(s.cur_x - s.x);
        s.done = 1;
      } else {
        ++s.iter;
        opt_request(s, s.x, false);
      }
      break;
    }
    case kPhLogSpaceGradientAscent: {  // optimization.hpp:347-365
      const double tolerance = prm.decimal_tolerance;  // 10^-significant digits
      const double y = log(s.x);
      const double log_space_grad = s.x * d1;
      const double new_x = exp(y + log_space_grad * prm.log_step_size);
      s.x = fmax(new_x, exp(prm.min_log_bl));
      if (fabs(d1) < fabs(ll) * tolerance || s.iter >= prm.max_iter) {
        st.bl[s.edge] = s.x;
        st.diff[s.edge] = fabs(s.cur_x - s.x);
        s.done = 1;
      } else {
        ++s.iter;
        opt_request(s, s.x, false);
      }
      break;
    }
    default: {  // Newton, optimization.hpp:367-402 on gp_engine.cpp:641-653
      const double tolerance = prm.decimal_tolerance;  // 10^-significant digits
      const double t = s.t_eval;
      const double f_prime_y = t * d1;
      const double f_double_prime_y = f_prime_y + (t * t) * d2;
      bool stop = fabs(f_double_prime_y) < prm.denominator_tolerance;
      double new_x = s.x;
      if (!stop) {
        new_x = s.x - f_prime_y / f_double_prime_y;
        if (new_x < prm.min_log_bl) new_x = s.x - 0.5 * (s.x - prm.min_log_bl);
        if (new_x > prm.max_log_bl) new_x = s.x - 0.5 * (s.x - prm.max_log_bl);
        const double delta = fabs(s.x - new_x);
        stop = delta < tolerance || fabs(f_prime_y) < fabs(ll) * tolerance ||
               s.iter == prm.max_iter;
      }
      if (stop) {
        const double new_bl = exp(s.x);
        st.bl[s.edge] = new_bl;
        st.diff[s.edge] = fabs(s.cur_x - new_bl);
        s.done = 1;
      } else {
        s.x = new_x;
        ++s.iter;
        opt_request(s, s.x, true);
      }
      break;
    }
  }
}

__global__ void __launch_bounds__(kTile)
    k_opt_prepare(DeviceState st, const OptOp* __restrict__ ops, int tiles,
                  OptState* __restrict__ states, OptParams prm, int method,
                  double* __restrict__ coef, int init_states) {
  const int o = blockIdx.x / tiles;
  const int tile = blockIdx.x - o * tiles;
  const OptOp op = ops[o];
  if (init_states && tile == 0 && threadIdx.x == 0) {
    OptState s;
    opt_init(s, st, prm, method, op);
    states[o] = s;
  }
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  if (p >= st.P) return;
  const V4 r = load_plv(op.parent, p);
  const V4 c = load_plv(op.child, p);
  const int G = c_model.n_groups;
  double cg[kMaxEigenGroups] = {0., 0., 0., 0.};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double rv = r.a * c_model.V[k] + r.b * c_model.V[4 + k] + r.c * c_model.V[8 + k] +
                      r.d * c_model.V[12 + k];
    const double vp = c_model.Vinv[4 * k] * c.a + c_model.Vinv[4 * k + 1] * c.b +
                      c_model.Vinv[4 * k + 2] * c.c + c_model.Vinv[4 * k + 3] * c.d;
    const double term = rv * vp;
    const int g = c_model.group[k];
#pragma unroll
    for (int gg = 0; gg < kMaxEigenGroups; ++gg)
      if (gg == g) cg[gg] += term;
  }
  double* dst = coef + (static_cast<int64_t>(o) * st.P_stride + p) * G;
  for (int g = 0; g < G; ++g) dst[g] = cg[g];
}

template <int G>
__global__ void __launch_bounds__(kTile)
    k_opt_eval(DeviceState st, int tiles, const OptState* __restrict__ states,
               const double* __restrict__ coef, int n_derivatives,
               double* __restrict__ partials) {
  const int o = blockIdx.x / tiles;
  const int tile = blockIdx.x - o * tiles;
  if (states[o].done) return;  // sums of finished edges are never read
  double e[G], e1[G], e2[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const double l = c_model.group_lambda[g];
    e[g] = states[o].e[g];
    e1[g] = l * e[g];
    e2[g] = l * l * e[g];
  }
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  double f = 0., g1 = 0., g2 = 0.;
  if (p < st.P) {
    const double* c = coef + (static_cast<int64_t>(o) * st.P_stride + p) * G;
    double cg[G];
    if (G == 2) {
      const double2 v = *reinterpret_cast<const double2*>(c);
      cg[0] = v.x;
      cg[1] = v.y;
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) cg[g] = c[g];
    }
    double L = 0., L1 = 0., L2 = 0.;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      L += cg[g] * e[g];
      L1 += cg[g] * e1[g];
      L2 += cg[g] * e2[g];
    }
    const double w = st.weights[p];
    f = log(L) * w;
    if (n_derivatives >= 1) g1 = (L1 / L) * w;                       // gp_engine.cpp:493-496
    if (n_derivatives >= 2) g2 = ((L2 * L - L1 * L1) / (L * L)) * w;  // gp_engine.cpp:530-538
  }
  const int64_t base = static_cast<int64_t>(o) * 3 * tiles + tile;
  f = block_reduce(f, SumOp(), 0.);
  if (threadIdx.x == 0) partials[base] = f;
  if (n_derivatives >= 1) {
    g1 = block_reduce(g1, SumOp(), 0.);
    if (threadIdx.x == 0) partials[base + tiles] = g1;
  }
  if (n_derivatives >= 2) {
    g2 = block_reduce(g2, SumOp(), 0.);
    if (threadIdx.x == 0) partials[base + 2 * tiles] = g2;
  }
}

// c0, c1 = the two eigen-group coefficients of L_p(t) for one pattern (see k_opt_prepare_ratio);
// rho = c1 / c0.
// Splits t > 0 (finite, normal) into m * 2^e with m in [0.5, 1).
__device__ __forceinline__ bool split_positive(double t, double& m, int& e) {
  const int hi = __double2hiint(t);
  // biased exponent in [1, 2046] and sign bit clear
  if (static_cast<unsigned>(hi - 0x00100000) >= 0x7fe00000u) return false;
  e = (hi >> 20) - 1022;
  m = __hiloint2double((hi & 0x000fffff) | 0x3fe00000, __double2loint(t));
  return true;
}

__device__ __forceinline__ double pow_small(double t, int wi) {  // t^wi, 1 <= wi <= 7
  const double t2 = t * t;
  double r = (wi & 1) ? t : 1.;
  if (wi & 2) r *= t2;
  if (wi & 4) r *= t2 * t2;
  return r;
}

// Each coefficient is the sum of its group's eigen terms, in eigenvalue order: c_g = sum_{k in g} (V^T r)_k
// (V^-1 p)_k. JC69 - the model the reference engine instantiates (gp_engine.hpp:366) - takes a branch that
// produces the SAME BITS with 24 instead of 38 FP64 instructions: its eigenvector entries are 0, +-1/2, +-1,
// +-2, 1/4, 1/8 (substitution_model.cpp:20-26), products with those are exact and adding an exact zero
// changes nothing, so factoring the powers of two out of the general expression is not a re-association.
// (A cheaper form for any two-group model - only the smaller group summed, the other taken as r.p minus it -
// was measured and dropped: it moves rho by an ulp, which is enough to move a Brent termination decision of
// the `hello` fixture off the reference's side; profiles/r02_sweep_ab.md.)
__device__ __forceinline__ void eigen_group_coefficients(const V4& r, const V4& c, double& c0, double& c1) {
  if (c_model.is_jc69) {
    const double sr = ((r.a + r.b) + r.c) + r.d, sp = ((c.a + c.b) + c.c) + c.d;  // k = 0: V[:,0] = 1, Vinv[0,:] = 1/4
    const double ar = ((r.a - r.b) + r.c) - r.d, ap = ((c.a - c.b) + c.c) - c.d;  // k = 1: +-2, +-1/8
    const double p2 = (r.b - r.d) * (c.b - c.d);                                   // k = 2: +-1/2 on (C, T); +-1
    const double p3 = (r.a - r.c) * (c.a - c.c);                                   // k = 3: +-1/2 on (A, G); +-1
    c0 = 0.25 * (sr * sp);
    c1 = fma(0.5, p3, fma(0.5, p2, 0.25 * (ar * ap)));
  } else {
    c0 = 0.;
    c1 = 0.;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double rv = r.a * c_model.V[k] + r.b * c_model.V[4 + k] + r.c * c_model.V[8 + k] +
                        r.d * c_model.V[12 + k];
      const double vp = c_model.Vinv[4 * k] * c.a + c_model.Vinv[4 * k + 1] * c.b +
                        c_model.Vinv[4 * k + 2] * c.c + c_model.Vinv[4 * k + 3] * c.d;
      const double term = rv * vp;
      if (c_model.group[k] == 0) c0 += term; else c1 += term;
    }
  }
}
__device__ __forceinline__ void ratio_coefficients(const V4& r, const V4& c, double& rho, double& c0_out) {
  double c0, c1;
  eigen_group_coefficients(r, c, c0, c1);
  rho = c0 != 0. ? c1 / c0 : 0.;
  c0_out = c0;
}
// sum += w log c, through (mantissa product, exponent sum) for w = 1..7, so that a thread pays one log per
// item instead of one per pattern. c may be as small as thr^2 (1e-80): the exponent is split off BEFORE the
// power is taken. Everything else (c <= 0, other weights) goes through an explicit log.
struct LogSum {
  double prod = 1., slow = 0.;
  int esum = 0;
  __device__ __forceinline__ void add(double c, double w) {
    double m;
    int e;
    const int wi = static_cast<int>(w);
    if (w == static_cast<double>(wi) && wi >= 1 && wi <= 7 && split_positive(c, m, e)) {
      prod *= pow_small(m, wi);  // >= 2^-7
      esum += e * wi;
      if (split_positive(prod, m, e)) { prod = m; esum += e; }
    } else if (w != 0.) {
      slow += w * log(c);
    }
  }
  __device__ __forceinline__ double value() const {
    return log(prod) + static_cast<double>(esum) * 0.6931471805599453094 + slow;
  }
};

// The streaming loop of k_opt_prepare_ratio for the shapes every optimisation level of a GP op list has:
// a dense parent r-PLV and a child p-PLV that is dense or a leaf (1-byte symbols). One pattern tile per
// trip; the next tile's loads - PLVs (or the raw symbol byte) and the packed (rho position, weight code)
// word - are issued before the current tile's arithmetic and not touched until the next trip.
//  * rho = c1 / c0 is formed on operands scaled by 2^-exponent(c0) (exact): c0 can sit at thr^2 = 1e-80,
//    where the IEEE division leaves its fast path; the quotient is the same bits.
//  * K_e += w log c0 for w = 1..7 as mantissa^w (branch-free) and w * exponent: no log per pattern.
// Not inlined, so that the two instantiations stay two tight loops.
template <bool kSym>
__device__ __noinline__ void prepare_span(const double* __restrict__ parent, const void* __restrict__ child,
                                          int64_t P, int tile_begin, int tile_end,
                                          const int32_t* __restrict__ pos_w, const double* __restrict__ weights,
                                          double* __restrict__ rho_o, double* k_out) {
  double prod = 1., slow = 0.;
  int esum = 0;
  int64_t p = static_cast<int64_t>(tile_begin) * kTile + threadIdx.x;
  bool live = tile_begin < tile_end && p < P;
  V4 r = {1., 1., 1., 1.}, c = r;
  int sym = 4, code = 0;
  if (live) {
    r = ld256(parent + 4 * p);
    if (kSym) sym = static_cast<const uint8_t*>(child)[p]; else c = ld256(static_cast<const double*>(child) + 4 * p);
    code = pos_w[p];
  }
#pragma unroll 2
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t pn = p + kTile;
    const bool live_n = tile + 1 < tile_end && pn < P;
    V4 rn = r, cn = c;
    int sym_n = sym, code_n = code;
    if (live_n) {  // in flight during the arithmetic below
      rn = ld256(parent + 4 * pn);
      if (kSym) sym_n = static_cast<const uint8_t*>(child)[pn]; else cn = ld256(static_cast<const double*>(child) + 4 * pn);
      code_n = pos_w[pn];
    }
    if (live) {
      if (kSym) {  // InitializePLVsWithSitePatterns, gp_engine.cpp:544-562
        const bool gap = (sym == 4);
        c.a = (gap || sym == 0) ? 1. : 0.;
        c.b = (gap || sym == 1) ? 1. : 0.;
        c.c = (gap || sym == 2) ? 1. : 0.;
        c.d = (gap || sym == 3) ? 1. : 0.;
      }
      double c0, c1, m, rho;
      int e;
      eigen_group_coefficients(r, c, c0, c1);
      const int wi = code & 7;
      if (split_positive(c0, m, e)) {
        const double scale = __hiloint2double((1023 - e) << 20, 0);  // 2^-e: e in [-1021, 1024]
        rho = (c1 * scale) / m;
        if (wi != 0) {
          const double m2 = m * m, m4 = m2 * m2;
          prod *= ((wi & 1) ? m : 1.) * ((wi & 2) ? m2 : 1.) * ((wi & 4) ? m4 : 1.);  // >= 2^-7 per pattern
          esum += e * wi;
        } else {
          slow += weights[p] * (log(m) + static_cast<double>(e) * 0.6931471805599453094);
        }
      } else {  // c0 <= 0, subnormal or not finite
        rho = c0 != 0. ? c1 / c0 : 0.;
        const double w = wi != 0 ? static_cast<double>(wi) : weights[p];
        if (w != 0.) slow += w * log(c0);
      }
      rho_o[code >> 3] = rho;
    }
    r = rn;
    c = cn;
    sym = sym_n;
    code = code_n;
    p = pn;
    live = live_n;
  }
  // at most 32 tiles per block: prod >= 2^-224
  *k_out = log(prod) + static_cast<double>(esum) * 0.6931471805599453094 + slow;
}

// ---- Brent objective for two-eigenvalue models (JC69): ratio form -----------------------------
// L_p(t) = c0_p e^{l0 t} + c1_p e^{l1 t} = c0_p e^{l0 t} (1 + rho_p x),  rho_p = c1_p / c0_p,
// x = e^{(l1 - l0) t}. Hence  sum_p w_p log L_p = K + W l0 t + sum_p w_p log(1 + rho_p x)  with
// K = sum_p w_p log c0_p computed once per edge. Only rho (8 B per pattern) is kept, and the last
// sum is evaluated as the log of a running product with the binary exponents split off, so a
// thread pays one log() per kOptPatternsPerThread patterns instead of one per pattern.
__global__ void __launch_bounds__(kTile, 4)
    k_opt_prepare_ratio(DeviceState st, const OptOp* __restrict__ ops, int n_ops, int tiles,
                        int tiles_per_block, OptState* __restrict__ states, OptParams prm, int method,
                        double* __restrict__ rho, const int32_t* __restrict__ perm,
                        const int32_t* __restrict__ pos_w,
                        int64_t rho_stride, double* __restrict__ partials,
                        int32_t* __restrict__ active, int active_capacity) {
  // tile-major: the edges of one pattern tile group are neighbours in the grid, so a parent r-PLV
  // or child p-PLV tile shared by several edges is read from HBM once and from L2 afterwards
  const int n_groups = gridDim.x / n_ops;
  const int tile_group = blockIdx.x / n_ops;
  const int o = blockIdx.x - tile_group * n_ops;
  const OptOp op = ops[o];
  if (tile_group == 0 && threadIdx.x == 0) {
    OptState s;
    opt_init(s, st, prm, method, op);
    states[o] = s;
    // active[0..1]: number of edges still optimising, double-buffered by round parity;
    // active[4 + parity * capacity + i]: their indices (see k_opt_step).
    active[4 + o] = o;
    if (o == 0) {
      active[0] = n_ops;
      active[1] = 0;
    }
  }
  (void)active_capacity;
  const int tile_begin = tile_group * tiles_per_block;
  const int tile_end = min(tiles, tile_begin + tiles_per_block);
  double* const rho_o = rho + static_cast<int64_t>(o) * rho_stride;
  double k_mine = 0.;  // this thread's share of K_e = sum_p w_p log c0_p
  if (pos_w != nullptr && op.parent.kind == kPlvDense && op.child.kind == kPlvDense) {
    prepare_span<false>(static_cast<const double*>(op.parent.ptr), op.child.ptr, st.P, tile_begin, tile_end, pos_w,
                        st.weights, rho_o, &k_mine);
  } else if (pos_w != nullptr && op.parent.kind == kPlvDense && op.child.kind == kPlvSymbols) {
    prepare_span<true>(static_cast<const double*>(op.parent.ptr), op.child.ptr, st.P, tile_begin, tile_end, pos_w,
                       st.weights, rho_o, &k_mine);
  } else {  // any other pair of PLV kinds (hand-written op lists): the general loads
    LogSum k_sum;
#pragma unroll 1
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
      if (p < st.P) {
        const V4 r = load_plv(op.parent, p);
        const V4 c = load_plv(op.child, p);
        double rr, cc;
        ratio_coefficients(r, c, rr, cc);
        rho_o[perm[p]] = rr;
        k_sum.add(cc, st.weights[p]);
      }
    }
    k_mine = k_sum.value();
  }
  const double k_part = block_reduce(k_mine, SumOp(), 0.);
  if (threadIdx.x == 0) partials[static_cast<int64_t>(o) * n_groups + tile_group] = k_part;
}

// The same pass for the pipelined cluster scheme (Engine::RunOptimizerPipelined): rho goes to the
// CLUSTER weight-class layout (classes padded to rows of kClusterThreads patterns, position
// cpos[p]) of a chunk buffer that the consumer kernel k_opt_cluster<T, true> copies into shared
// memory; optimiser states are initialised by the consumer. partials: K_e tile-group sums.
__global__ void __launch_bounds__(kTile, 4)
    k_opt_prepare_cluster(DeviceState st, const OptOp* __restrict__ ops, int n_ops, int tiles,
                          int tiles_per_block, int n_groups, double* __restrict__ rho,
                          const int32_t* __restrict__ cpos, int64_t rho_stride,
                          double* __restrict__ partials) {
  // fixed grid walking the (edge, tile group) items edge-major within a tile group (a PLV tile shared by
  // several edges of the chunk is then served from L2): the producer never occupies more of an SM than
  // the launch asks for, so the consumer's clusters always find room next to it
  for (int item = blockIdx.x; item < n_ops * n_groups; item += gridDim.x) {
  const int tile_group = item / n_ops;
  const int o = item - tile_group * n_ops;
  const OptOp op = ops[o];
  LogSum k_sum;
  const int tile_begin = tile_group * tiles_per_block;
  const int tile_end = min(tiles, tile_begin + tiles_per_block);
  double* const rho_o = rho + static_cast<int64_t>(o) * rho_stride;
  // two pattern tiles per trip: four 256-bit loads in flight before the first divide
  int tile = tile_begin;
  for (; tile + 2 <= tile_end; tile += 2) {
    const int64_t p0 = static_cast<int64_t>(tile) * kTile + threadIdx.x, p1 = p0 + kTile;
    const bool live0 = p0 < st.P, live1 = p1 < st.P;
    V4 r0 = {1., 1., 1., 1.}, c0v = r0, r1 = r0, c1v = r0;
    int32_t q0 = 0, q1 = 0;
    double w0 = 0., w1 = 0.;
    if (live0) { r0 = load_plv(op.parent, p0); c0v = load_plv(op.child, p0); q0 = cpos[p0]; w0 = st.weights[p0]; }
    if (live1) { r1 = load_plv(op.parent, p1); c1v = load_plv(op.child, p1); q1 = cpos[p1]; w1 = st.weights[p1]; }
    double rr, cc;
    if (live0) {
      ratio_coefficients(r0, c0v, rr, cc);
      rho_o[q0] = rr;
      k_sum.add(cc, w0);
    }
    if (live1) {
      ratio_coefficients(r1, c1v, rr, cc);
      rho_o[q1] = rr;
      k_sum.add(cc, w1);
    }
  }
  for (; tile < tile_end; ++tile) {
    const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
    if (p < st.P) {
      const V4 r = load_plv(op.parent, p);
      const V4 c = load_plv(op.child, p);
      double rr, cc;
      ratio_coefficients(r, c, rr, cc);
      rho_o[cpos[p]] = rr;
      k_sum.add(cc, st.weights[p]);
    }
  }
  const double k_part = block_reduce(k_sum.value(), SumOp(), 0.);
  if (threadIdx.x == 0) partials[static_cast<int64_t>(o) * n_groups + tile_group] = k_part;
  }
}


// One item = one tile group (kTile * kOptPatternsPerThread = 2048 positions, all of one weight class) of
// one still-active edge. A thread owns 8 positions of the item as two 256-bit loads, and the loads of its
// NEXT item are issued before the current one is evaluated, so 128 B per thread stay in flight while the
// arithmetic runs (the kernel is bound by HBM: 8 B per pattern and evaluation, nothing is written but one
// partial per warp). Every factor t = 1 + rho x is 0 or lies in [2^-53, 4] (see k_opt_cluster), so the
// eight factors of a thread multiply up as they are: one log per 8 patterns and no exponent bookkeeping.
struct Rho8 {
  V4 lo, hi;
};
__device__ __forceinline__ Rho8 load_rho8(const double* item_base) {
  Rho8 r;
  r.lo = ld256(item_base + 4 * threadIdx.x);
  r.hi = ld256(item_base + 4 * kTile + 4 * threadIdx.x);
  return r;
}
__global__ void __launch_bounds__(kTile, 4)
    k_opt_eval_ratio(DeviceState st, int tile_groups, const OptState* __restrict__ states,
                     const double* __restrict__ rho, int64_t rho_stride,
                     const double* __restrict__ wperm, const uint8_t* __restrict__ group_class,
                     double* __restrict__ partials, int32_t* __restrict__ active,
                     int active_capacity, int parity) {
  // Fixed grid walking the (still active edge, tile group) items: late rounds, when most edges
  // of the batch have converged, cost only the work that is left.
  const int n_active = active[parity];
  const int32_t* list = active + 4 + parity * active_capacity;
  if (blockIdx.x == 0 && threadIdx.x == 0) active[parity ^ 1] = 0;  // filled by this round's step
  const int n_items = n_active * tile_groups;  // < 2^31: the batch's rho buffer is at most a few GiB
  constexpr int64_t kItem = static_cast<int64_t>(kTile) * kOptPatternsPerThread;
  auto item_base = [&](int item, int& o, int& tg) {
    const int a = item / tile_groups;
    tg = item - a * tile_groups;
    o = list[a];
    return rho + static_cast<int64_t>(o) * rho_stride + static_cast<int64_t>(tg) * kItem;
  };
  int item = blockIdx.x;
  if (item >= n_items) return;
  int o, tg;
  Rho8 cur = load_rho8(item_base(item, o, tg));
  for (;;) {
    const int next = item + gridDim.x;
    int o_next = 0, tg_next = 0;
    Rho8 nxt = cur;
    if (next < n_items) nxt = load_rho8(item_base(next, o_next, tg_next));  // in flight during the arithmetic below
    const double x = states[o].x_ratio;
    const int cls = group_class[tg];  // block-uniform: weight - 1, or 7 = general weights
    const double t0 = fma(cur.lo.a, x, 1.0), t1 = fma(cur.lo.b, x, 1.0), t2 = fma(cur.lo.c, x, 1.0),
                 t3 = fma(cur.lo.d, x, 1.0), t4 = fma(cur.hi.a, x, 1.0), t5 = fma(cur.hi.b, x, 1.0),
                 t6 = fma(cur.hi.c, x, 1.0), t7 = fma(cur.hi.d, x, 1.0);
    double f;
    if (cls == 0) {  // weight 1: the bulk of any alignment
      f = log(((t0 * t1) * (t2 * t3)) * ((t4 * t5) * (t6 * t7)));
    } else if (cls < 7) {  // weights 2..7: t^w >= 2^-371, two factors per log
      const int wi = cls + 1;
      f = log(pow_small(t0, wi) * pow_small(t1, wi)) + log(pow_small(t2, wi) * pow_small(t3, wi)) +
          log(pow_small(t4, wi) * pow_small(t5, wi)) + log(pow_small(t6, wi) * pow_small(t7, wi));
    } else {  // general weights: explicit log (padding has weight 0 and rho 0)
      const double* wp = wperm + static_cast<int64_t>(tg) * kItem;
      const V4 wl = ld256(wp + 4 * threadIdx.x), wh = ld256(wp + 4 * kTile + 4 * threadIdx.x);
      f = 0.;
      if (wl.a != 0.) f += wl.a * log(t0);
      if (wl.b != 0.) f += wl.b * log(t1);
      if (wl.c != 0.) f += wl.c * log(t2);
      if (wl.d != 0.) f += wl.d * log(t3);
      if (wh.a != 0.) f += wh.a * log(t4);
      if (wh.b != 0.) f += wh.b * log(t5);
      if (wh.c != 0.) f += wh.c * log(t6);
      if (wh.d != 0.) f += wh.d * log(t7);
    }
    // One partial per warp (fixed shuffle tree) and no block barrier: warps of a block run ahead
    // independently into the next item; k_opt_step sums the kTile/32 * tile_groups partials.
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) f += __shfl_down_sync(0xffffffffu, f, sh);
    if ((threadIdx.x & 31) == 0)
      partials[(static_cast<int64_t>(o) * tile_groups + tg) * (kTile / 32) + (threadIdx.x >> 5)] = f;
    if (next >= n_items) break;
    item = next;
    o = o_next;
    tg = tg_next;
    cur = nxt;
  }
}

// Consumes this round's objective sums and advances every optimiser of the batch. Single rank:
// `partials` holds n_parts partial sums per (edge, derivative) and is reduced here in a fixed
// order; multi-rank: `sums` holds the all-reduced totals. edge_const/time_coef add the terms of the
// ratio form that do not depend on the pattern: K_e + W * lambda_0 * t.
__global__ void k_opt_step(DeviceState st, int n_ops, OptState* __restrict__ states, OptParams prm,
                           const double* __restrict__ sums, const double* __restrict__ partials,
                           int n_parts, int n_values, int value_stride,
                           const double* __restrict__ edge_const,
                           int32_t* __restrict__ active_counter, int32_t* __restrict__ active,
                           int active_capacity, int parity) {
  // One warp per edge: the lanes sum the partials (lane-strided, then a fixed shuffle tree), lane
  // 0 advances the optimiser.
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  int o = a;
  if (active != nullptr) {  // compacted list of the edges still optimising
    if (a >= active[parity]) return;
    o = active[4 + parity * active_capacity + a];
  } else if (a >= n_ops) {
    return;
  }
  if (states[o].done) return;
  double v[3] = {0., 0., 0.};
  if (partials != nullptr) {
    for (int k = 0; k < n_values; ++k) {
      const double* row = partials + (static_cast<int64_t>(o) * value_stride + k) * n_parts;
      double acc = 0.;
      for (int t = lane; t < n_parts; t += 32) acc += row[t];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, sh);
      v[k] = acc;
    }
  } else {
    for (int k = 0; k < n_values; ++k) v[k] = sums[static_cast<int64_t>(o) * value_stride + k];
  }
  if (lane != 0) return;
  OptState s = states[o];
  double ll = v[0] + s.ll_offset;
  if (edge_const != nullptr)
    ll += edge_const[o] + st.total_weight * c_model.group_lambda[0] * s.t_eval;
  opt_advance(s, st, prm, ll, v[1], v[2]);
  states[o] = s;
  if (s.done) atomicAdd(st.feval_total, static_cast<unsigned long long>(s.evals));
  if (!s.done) {
    if (active != nullptr)
      active[4 + (parity ^ 1) * active_capacity + atomicAdd(active + (parity ^ 1), 1)] = o;
    if (active_counter != nullptr) atomicAdd(active_counter, 1);
  }
}


// ---- Taylor-model Brent: the streamed scheme for plain Brent on a two-eigenvalue model ------------
// (gp_types.h, OptPass.) Replaces "one pass over rho per objective evaluation" by "one pass per model":
// the reference's BrentMinimize (optimization.hpp:71-188) asks for ~16 objective values per edge, the
// first three at points that do not depend on the data (its start, a golden-section step, one of two
// golden-section steps) and the last ones within a fraction of a percent of each other.

// After k_opt_prepare_ratio: the first pass of every edge evaluates its pending request (the start)
// together with the next requests Brent can make, found by advancing COPIES of the optimiser with
// made-up objective values (only the order of the values matters for those steps: the parabolic fit
// through coincident points degenerates to p = q = 0 whatever they are). Nothing depends on the
// guesses being right: k_opt_step_model uses a cached value only if the real request equals its point.
__device__ void opt_first_pass(const OptState& s, const DeviceState& st, const OptParams& prm, OptPass& pp) {
  for (int k = 0; k < kOptPoints; ++k) {
    pp.px[k] = s.x_ratio;
    pp.ps[k] = s.x_eval;
    pp.pt[k] = s.t_eval;
    pp.cache_s[k] = 0.;
    pp.cache_ll[k] = 0.;
  }
  pp.n_pts = 1;
  pp.centre = 0;
  pp.n_cache = 0;
  pp.passes = 0;
  pp.c = 0.;
  pp.S_c = 0.;
  pp.radius = 0.;
  for (int j = 0; j < kOptMoments; ++j) pp.mj[j] = 0.;
  if (!s.done && s.method == 0 && s.phase == kPhBrentInit) {
    auto set = [&](int k, const OptState& a) {
      pp.px[k] = a.x_ratio;
      pp.ps[k] = a.x_eval;
      pp.pt[k] = a.t_eval;
    };
    OptState a = s;
    a.speculative = 1;
    opt_advance(a, st, prm, 0., 0., 0.);  // f(start) = 0
    if (!a.done) {
      set(1, a);
      OptState b = a;
      opt_advance(b, st, prm, 1., 0., 0.);  // f(u) = -1 <= f(x): accepted
      if (!b.done) set(2, b);
      OptState r = a;
      opt_advance(r, st, prm, -1., 0., 0.);  // f(u) = 1 > f(x): rejected
      if (!r.done) set(3, r);
      pp.n_pts = kOptPoints;
      // Where later requests will cluster is not known yet. A first optimisation starts from default
      // lengths and mostly walks towards the accepted golden-section side; a later one
      // (!IsFirstOptimization, dag_branch_handler.hpp:49-52) starts next to its optimum.
      pp.centre = (prm.check_convergence || b.done) ? 0 : 2;
    }
  }
}
__global__ void k_opt_plan(DeviceState st, int n_ops, const OptState* __restrict__ states, OptParams prm,
                           OptPass* __restrict__ pass, OptReq* __restrict__ req, int32_t* __restrict__ active) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_ops) return;
  if (o == 0) {
    active[0] = n_ops;  // round 0 lists every edge (finished ones as o < 0)
    active[1] = 0;
  }
  const OptState s = states[o];
  OptPass pp;
  opt_first_pass(s, st, prm, pp);
  pass[o] = pp;
  OptReq rq;
  for (int k = 0; k < kOptPoints; ++k) rq.x[k] = pp.px[k];
  rq.c = pp.px[pp.centre];
  rq.o = s.done ? -1 - o : o;
  rq.pad = 0;
  req[o] = rq;
}

// 1 / a for a normal a > 0: hardware seed (~2^-20) and two Newton steps; within an ulp or two, which is
// all the power sums need. a = 0 gives NaN (the model of that pass is dropped).
__device__ __forceinline__ double rcp_newton(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// A[j-1] += w (a^j + b^j + c^j + d^j), j = 1..12, through the elementary symmetric polynomials of the
// four values and Newton's identities p_j = e1 p_{j-1} - e2 p_{j-2} + e3 p_{j-3} - e4 p_{j-4}: 63 FP64
// instructions per four values instead of 92 for explicit powers. p_1 and p_2 (the terms that carry the
// model) are formed directly; the recurrence loses ~j^3 ulp against sum |z|^j by j = 12 (1e-12 relative),
// on terms that are below 1e-10 of the objective inside the model's radius.
__device__ __forceinline__ void add_moments4(double (&A)[kOptMoments], double a, double b, double c, double d,
                                             double w) {
  static_assert(kOptMoments == 12, "unrolled for 12 moments");
  const double s_ab = a + b, s_cd = c + d, p_ab = a * b, p_cd = c * d;
  const double e1 = s_ab + s_cd;
  const double e2 = fma(s_ab, s_cd, p_ab + p_cd);
  const double e3 = fma(p_ab, s_cd, p_cd * s_ab);
  const double e4 = p_ab * p_cd;
  const double p1 = e1;
  const double p2 = fma(a, a, fma(b, b, fma(c, c, d * d)));
  const double p3 = fma(e1, p2, fma(-e2, p1, 3. * e3));
  const double p4 = fma(e1, p3, fma(-e2, p2, fma(e3, p1, -4. * e4)));
  A[0] = fma(p1, w, A[0]);
  A[1] = fma(p2, w, A[1]);
  A[2] = fma(p3, w, A[2]);
  A[3] = fma(p4, w, A[3]);
  double q4 = p1, q3 = p2, q2 = p3, q1 = p4;  // p_{j-4} .. p_{j-1}
#pragma unroll
  for (int j = 4; j < kOptMoments; ++j) {
    const double pj = fma(e1, q1, fma(-e2, q2, fma(e3, q3, -(e4 * q4))));
    A[j] = fma(pj, w, A[j]);
    q4 = q3;
    q3 = q2;
    q2 = q1;
    q1 = pj;
  }
}

// z_i = r_i / t_i for eight t_i in one reciprocal (products pairwise up, one Newton reciprocal, back down):
// 21 multiplications + 5 instead of 8 x 6. Every t is 0 or in [2^-53, 4], so the product cannot leave the
// normal range; a zero factor turns all eight into NaN (the model of that pass is dropped). prod8 = the
// product of the eight t (the weight-1 log term of the centre).
__device__ __forceinline__ void ratios8(const double (&r)[8], const double (&t)[8], double (&z)[8], double& prod8) {
  const double t01 = t[0] * t[1], t23 = t[2] * t[3], t45 = t[4] * t[5], t67 = t[6] * t[7];
  const double t03 = t01 * t23, t47 = t45 * t67;
  prod8 = t03 * t47;
  const double inv = rcp_newton(prod8);
  const double i03 = inv * t47, i47 = inv * t03;
  const double i01 = i03 * t23, i23 = i03 * t01, i45 = i47 * t67, i67 = i47 * t45;
  z[0] = r[0] * (i01 * t[1]);
  z[1] = r[1] * (i01 * t[0]);
  z[2] = r[2] * (i23 * t[3]);
  z[3] = r[3] * (i23 * t[2]);
  z[4] = r[4] * (i45 * t[5]);
  z[5] = r[5] * (i45 * t[4]);
  z[6] = r[6] * (i67 * t[7]);
  z[7] = r[7] * (i67 * t[6]);
}

// Sum over the warp of 16 values per lane in 16 + 15 shuffles instead of 16 x 5: each step keeps the
// half of the values selected by one lane-id bit and hands the other half to the partner lane. Fixed
// order, so the result is deterministic. Afterwards lanes 2k and 2k+1 hold the warp total of value k in v[0].
__device__ __forceinline__ void warp_reduce16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? v[i] : v[i + half];
      const double keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Running sum of w log(product of factors) for the single-weight classes: the factors of a tile group
// multiply up raw (eight of them stay above 2^-424), the binary exponent of that product is split off and
// the mantissas keep multiplying (at most a few hundred per segment, each >= 2^-7), so a thread pays ONE
// log per point and segment.
struct ProductLog {
  double mant = 1., slow = 0.;
  int esum = 0;
  // *= prod8^wi, wi = 1..7 (the tile group's weight): mantissa^wi stays above 2^-7
  __device__ __forceinline__ void mul(double prod8, int wi) {
    double m;
    int e;
    if (split_positive(prod8, m, e)) {
      mant *= wi == 1 ? m : pow_small(m, wi);
      esum += e * wi;
    } else {
      slow += static_cast<double>(wi) * log(prod8);  // a zero factor: -inf, as the reference's log(0)
    }
  }
  __device__ __forceinline__ double value() const {
    return log(mant) + static_cast<double>(esum) * 0.6931471805599453094 + slow;
  }
};

// One item = one segment (seg_len tile groups of kTile * kOptPatternsPerThread positions) of one still
// active edge. NP = points evaluated (kOptPoints on the first pass of a search, 1 afterwards, where the
// point is also the centre). A thread owns 8 positions of every tile group, prefetches the next tile
// group (and the next item's request) while it works on the current one and keeps the NP log sums and the
// 12 power sums in registers for the whole segment; one block-wide reduction per segment.
// partials: [edge][value][segment]. The kernel is bound by FP64 issue, not by HBM (8 B per position).
template <int NP, int MINB>
__global__ void __launch_bounds__(kTile, MINB)
    k_opt_eval_model(int tile_groups, int seg_len, int n_seg, const OptReq* __restrict__ req,
                     const double* __restrict__ rho, int64_t rho_stride,
                     const double* __restrict__ wperm, OptClassStarts classes,
                     double* __restrict__ partials, int32_t* __restrict__ active, int parity) {
  __shared__ double s_red[kTile / 32][kOptPassValues];
  __shared__ double s_rq[2][sizeof(OptReq) / sizeof(double)];
  constexpr int kReqWords = static_cast<int>(sizeof(OptReq) / sizeof(double));
  const int n_active = active[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) active[parity ^ 1] = 0;  // filled by this round's step
  const int n_items = n_active * n_seg;
  constexpr int64_t kItem = static_cast<int64_t>(kTile) * kOptPatternsPerThread;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int item = blockIdx.x;
  if (item >= n_items) return;
  // The request of an item is block-uniform. Its words are fetched by the first threads one item ahead
  // (a per-thread load, so nothing waits for it) and handed over through shared memory at the barrier
  // that ends the item.
  if (threadIdx.x < kReqWords)
    s_rq[0][threadIdx.x] = reinterpret_cast<const double*>(req + item / n_seg)[threadIdx.x];
  __syncthreads();
  int buf = 0;
  for (; item < n_items; item += gridDim.x, buf ^= 1) {
    const int sg = item - (item / n_seg) * n_seg;
    const int next_item = item + static_cast<int>(gridDim.x);
    double next_word = 0.;
    if (threadIdx.x < kReqWords && next_item < n_items)
      next_word = reinterpret_cast<const double*>(req + next_item / n_seg)[threadIdx.x];  // in flight below
    OptReq cur_rq;
#pragma unroll
    for (int k = 0; k < kOptPoints; ++k) cur_rq.x[k] = s_rq[buf][k];
    cur_rq.c = s_rq[buf][kOptPoints];
    cur_rq.o = __double2loint(s_rq[buf][kOptPoints + 1]);
    const int o = cur_rq.o;
    const double c = NP == 1 ? cur_rq.x[0] : cur_rq.c;
    const int tg_begin = sg * seg_len;
    const int tg_end = o < 0 ? tg_begin : min(tile_groups, tg_begin + seg_len);  // o < 0: nothing to do
    const double* base = rho + static_cast<int64_t>(o < 0 ? 0 : o) * rho_stride;
    auto class_of = [&](int tg) {  // weight - 1, or 7 = general weights
      int cls = 0;
#pragma unroll
      for (int k = 1; k < 8; ++k) cls += tg >= classes.start[k];
      return cls;
    };
    ProductLog S[NP];
    double A[kOptMoments];
#pragma unroll
    for (int j = 0; j < kOptMoments; ++j) A[j] = 0.;
    Rho8 cur = {{0., 0., 0., 0.}, {0., 0., 0., 0.}};
    if (tg_begin < tg_end) cur = load_rho8(base + static_cast<int64_t>(tg_begin) * kItem);
#pragma unroll 1
    for (int tg = tg_begin; tg < tg_end; ++tg) {
      Rho8 nxt = cur;
      if (tg + 1 < tg_end) nxt = load_rho8(base + static_cast<int64_t>(tg + 1) * kItem);  // in flight below
      const int cls = class_of(tg);  // block-uniform
      const double r[8] = {cur.lo.a, cur.lo.b, cur.lo.c, cur.lo.d, cur.hi.a, cur.hi.b, cur.hi.c, cur.hi.d};
      double tc[8], z[8], prod_c;
#pragma unroll
      for (int i = 0; i < 8; ++i) tc[i] = fma(r[i], c, 1.0);
      ratios8(r, tc, z, prod_c);
      if (cls < 7) {  // one weight (1..7) for the whole tile group
        const int wi = cls + 1;
        const double wd = static_cast<double>(wi);
        add_moments4(A, z[0], z[1], z[2], z[3], wd);
        add_moments4(A, z[4], z[5], z[6], z[7], wd);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          double prod = prod_c;
          if (NP > 1) {
            const double xk = cur_rq.x[k];
            const double t0 = fma(r[0], xk, 1.0), t1 = fma(r[1], xk, 1.0), t2 = fma(r[2], xk, 1.0),
                         t3 = fma(r[3], xk, 1.0), t4 = fma(r[4], xk, 1.0), t5 = fma(r[5], xk, 1.0),
                         t6 = fma(r[6], xk, 1.0), t7 = fma(r[7], xk, 1.0);
            prod = ((t0 * t1) * (t2 * t3)) * ((t4 * t5) * (t6 * t7));
          }
          S[k].mul(prod, wi);
        }
      } else {  // general weights (padding: weight 0, rho 0): one position at a time, kept small
        const double* wp = wperm + static_cast<int64_t>(tg) * kItem;
        const V4 wl = ld256(wp + 4 * threadIdx.x), wh = ld256(wp + 4 * kTile + 4 * threadIdx.x);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const double wv = i == 0 ? wl.a : i == 1 ? wl.b : i == 2 ? wl.c : i == 3 ? wl.d
                          : i == 4 ? wh.a : i == 5 ? wh.b : i == 6 ? wh.c : wh.d;
          if (wv == 0.) continue;
          double zi = z[0], ri = r[0];
#pragma unroll
          for (int q = 1; q < 8; ++q) {
            zi = i == q ? z[q] : zi;
            ri = i == q ? r[q] : ri;
          }
          double zp = zi;
#pragma unroll
          for (int j = 0; j < kOptMoments; ++j) {
            A[j] = fma(zp, wv, A[j]);
            zp *= zi;
          }
#pragma unroll
          for (int k = 0; k < NP; ++k) S[k].slow += wv * log(fma(ri, cur_rq.x[k], 1.0));
        }
      }
      cur = nxt;
    }
    double v[kOptPassValues];
#pragma unroll
    for (int k = 0; k < kOptPoints; ++k) v[k] = k < NP ? S[k < NP ? k : 0].value() : 0.;
#pragma unroll
    for (int j = 0; j < kOptMoments; ++j) v[kOptPoints + j] = A[j];
    warp_reduce16(v);
    if ((lane & 1) == 0) s_red[warp][lane >> 1] = v[0];
    if (threadIdx.x < kReqWords) s_rq[buf ^ 1][threadIdx.x] = next_word;
    __syncthreads();
    if (threadIdx.x < kOptPassValues && o >= 0) {
      double acc = s_red[0][threadIdx.x];
#pragma unroll
      for (int wv = 1; wv < kTile / 32; ++wv) acc += s_red[wv][threadIdx.x];
      partials[(static_cast<int64_t>(o) * kOptPassValues + threadIdx.x) * n_seg + sg] = acc;
    }
    __syncthreads();  // s_red is free for the next item
  }
}

// S(x) from the model of the last pass; false outside its radius.
__device__ __forceinline__ bool opt_model_value(const OptPass& pp, double x, double& S) {
  const double d = x - pp.c;  // exact: x and c are within a factor of two of each other
  static_assert(kOptMoments == 12, "d^J below is written for J = 12");
  const double d2 = d * d, d4 = d2 * d2;
  if (!(d4 * d4 * d4 <= pp.radius)) return false;  // radius holds the J-th power
  double acc = pp.mj[kOptMoments - 2];
#pragma unroll
  for (int j = kOptMoments - 3; j >= 0; --j) acc = fma(acc, d, pp.mj[j]);
  S = fma(acc, d, pp.S_c);
  return true;
}

// One streamed (or on-chip) pass has produced v[0 .. kOptPassValues): the objective sums at the pass's points
// and the power sums about its centre. Builds the model, answers the pending request and then every
// further request that hits a cached point or falls inside the model's radius. On return either s.done
// or pp holds the next pass (one point: the optimiser's pending request, also the centre).
__device__ void opt_consume_pass(OptState& s, OptPass& pp, const double* v, double edge_const, double min_weight,
                                 const DeviceState& st, const OptParams& prm) {
  const double lin = st.total_weight * c_model.group_lambda[0];
  const double base = s.ll_offset + edge_const;
  pp.passes++;
  // the points of this pass other than the request: kept until the optimiser asks for them
  pp.n_cache = 0;
  for (int k = 1; k < pp.n_pts; ++k) {
    pp.cache_s[pp.n_cache] = pp.ps[k];
    pp.cache_ll[pp.n_cache] = v[k] + base + lin * pp.pt[k];
    pp.n_cache++;
  }
  const double ll0 = v[0] + base + lin * pp.pt[0];
  // the model about px[centre]
  {
    const int kc = pp.n_pts > 1 ? pp.centre : 0;
    pp.c = pp.px[kc];
    pp.S_c = v[kc];
    const double ll_c = v[kc] + base + lin * pp.pt[kc];
    const double MJ = v[kOptPoints + kOptMoments - 1];
    // (-1)^{j+1} / j
    constexpr double kInvJ[kOptMoments - 1] = {1., -1. / 2, 1. / 3, -1. / 4, 1. / 5, -1. / 6,
                                               1. / 7, -1. / 8, 1. / 9, -1. / 10, 1. / 11};
#pragma unroll
    for (int j = 1; j < kOptMoments; ++j) pp.mj[j - 1] = kInvJ[j - 1] * v[kOptPoints + j - 1];
    pp.mj[kOptMoments - 1] = MJ;
    // The model answers x when  2 M_J d^J / J <= tol  and  max |z| d <= 1/2, d = |x - c|; both are tests
    // on d^J (max |z|^J <= M_J / min_weight), so the radius is kept as its J-th power: no pow() here.
    pp.radius = 0.;
    bool finite = isfinite(ll_c);
#pragma unroll
    for (int j = 0; j < kOptMoments; ++j) finite = finite && isfinite(v[kOptPoints + j]);
    if (finite && MJ >= 0.) {
      if (MJ == 0.) {
        pp.radius = 1.;  // every z is 0: S is constant
      } else {
        const double J = static_cast<double>(kOptMoments);
        const double tol = 0x1p-55 * fabs(ll_c);  // a quarter ulp of the objective
        const double dj_err = tol * J / (2. * MJ);
        const double dj_half = min_weight / (MJ * static_cast<double>(1 << kOptMoments));
        const double r = 0.88 * fmin(dj_err, dj_half);  // 0.99^12
        if (r > 0. && isfinite(r)) pp.radius = r;
      }
    }
  }
  opt_advance(s, st, prm, ll0, 0., 0.);
  while (!s.done) {
    double ll = 0.;
    bool have = false;
    for (int k = 0; k < pp.n_cache; ++k) {
      if (pp.cache_s[k] == s.x_eval) {
        ll = pp.cache_ll[k];
        have = true;
      }
    }
    if (!have) {
      double S;
      if (opt_model_value(pp, s.x_ratio, S)) {
        ll = S + base + lin * s.t_eval;
        have = true;
      }
    }
    if (!have) break;
    opt_advance(s, st, prm, ll, 0., 0.);
  }
  if (s.done) return;
  pp.n_pts = 1;
  pp.centre = 0;
  pp.px[0] = s.x_ratio;
  pp.ps[0] = s.x_eval;
  pp.pt[0] = s.t_eval;
}

// Consumes one pass (the kOptPassValues sums of every still-active edge) and advances each optimiser as
// far as it can go without another pass: requests that hit a cached point or fall inside the model's
// radius are answered here. Single rank: partials [edge][value][segment], summed in a fixed order;
// multi-rank: `sums` [edge][value] all-reduced. One warp per edge, lane 0 decides.
__global__ void k_opt_step_model(DeviceState st, OptState* __restrict__ states, OptPass* __restrict__ pass,
                                 OptReq* __restrict__ req, OptParams prm, const double* __restrict__ sums,
                                 const double* __restrict__ partials, int n_seg,
                                 const double* __restrict__ edge_const, double min_weight,
                                 int32_t* __restrict__ active, int capacity, int parity) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= active[parity]) return;
  const int o = req[static_cast<int64_t>(parity) * capacity + a].o;
  if (o < 0) return;
  double mine = 0.;
  if (lane < kOptPassValues) {
    if (partials != nullptr) {
      const double* row = partials + (static_cast<int64_t>(o) * kOptPassValues + lane) * n_seg;
      for (int t = 0; t < n_seg; ++t) mine += row[t];
    } else {
      mine = sums[static_cast<int64_t>(o) * kOptPassValues + lane];
    }
  }
  double v[kOptPassValues];
#pragma unroll
  for (int k = 0; k < kOptPassValues; ++k) v[k] = __shfl_sync(0xffffffffu, mine, k);
  if (lane != 0) return;
  OptState s = states[o];
  OptPass pp = pass[o];
  opt_consume_pass(s, pp, v, edge_const[o], min_weight, st, prm);
  states[o] = s;
  if (s.done) {
    atomicAdd(st.feval_total, static_cast<unsigned long long>(s.evals));
    atomicAdd(st.feval_total + 1, static_cast<unsigned long long>(pp.passes));
    pass[o].passes = pp.passes;
    return;
  }
  pass[o] = pp;
  OptReq rq;
  for (int k = 0; k < kOptPoints; ++k) rq.x[k] = s.x_ratio;
  rq.c = s.x_ratio;
  rq.o = o;
  rq.pad = 0;
  req[static_cast<int64_t>(parity ^ 1) * capacity + atomicAdd(active + (parity ^ 1), 1)] = rq;
}

// After an on-chip search: rebuild the transition matrices of the program that hold this edge from
// its new branch length (k_build_matrices does the whole table; this does the edge's few slots).
__device__ __forceinline__ void refresh_edge_matrices(const DeviceState& st, const OptOp& op,
                                                      const OptRefresh& rf, int tid, int n_threads) {
  if (rf.pool == nullptr) return;
  const double t = st.bl[op.edge], q = st.q[op.edge];
  for (int k = tid; k < op.fix_n; k += n_threads) {
    const int v = rf.pool[op.fix_off + k];
    if (v & 1)
      build_matrix(t, 0, 1., rf.mtab_lik + 16 * static_cast<int64_t>(v >> 1));
    else
      build_matrix(t, 0, q, rf.mtab + 16 * static_cast<int64_t>(v >> 1));
  }
}

// ---- OptimizeBranchLength on chip (small alignments) -------------------------------------------
// Real alignments have 1e2..1e4 site patterns: there the round-per-launch scheme above is bound by
// launch and host-check latency, not by HBM. Here one block owns one edge for its whole 1-D search:
// it reads the two PLVs once (64 B per pattern), keeps the per-pattern eigen-coefficients in shared
// memory (G doubles per pattern), and loops objective evaluation -> block reduction -> optimiser
// decision (thread 0, the same opt_init/opt_advance state machine) until the optimiser is done. A
// level of k independent edges is k blocks of ONE launch with no host round trip, so whole
// Gauss-Seidel sweeps (GPDAG::BranchLengthOptimization) replay as a CUDA graph.
__global__ void __launch_bounds__(kTile)
    k_opt_block(DeviceState st, const OptOp* __restrict__ ops, const OptControl* __restrict__ ctl,
                OptRefresh refresh) {
  extern __shared__ __align__(16) double s_coef[];  // [G][P]
  __shared__ OptState s_state;
  const OptParams prm = ctl->prm;
  const int method = ctl->method, n_derivatives = ctl->n_derivatives;
  const OptOp op = ops[blockIdx.x];
  const int G = c_model.n_groups;
  const int P = static_cast<int>(st.P);
  if (threadIdx.x == 0) {
    OptState s;
    opt_init(s, st, prm, method, op);
    s_state = s;
  }
  for (int p = threadIdx.x; p < P; p += kTile) {
    const V4 r = load_plv(op.parent, p);
    const V4 c = load_plv(op.child, p);
    double cg[kMaxEigenGroups] = {0., 0., 0., 0.};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double rv = r.a * c_model.V[k] + r.b * c_model.V[4 + k] + r.c * c_model.V[8 + k] +
                        r.d * c_model.V[12 + k];
      const double vp = c_model.Vinv[4 * k] * c.a + c_model.Vinv[4 * k + 1] * c.b +
                        c_model.Vinv[4 * k + 2] * c.c + c_model.Vinv[4 * k + 3] * c.d;
      const double term = rv * vp;
      const int g = c_model.group[k];
#pragma unroll
      for (int gg = 0; gg < kMaxEigenGroups; ++gg)
        if (gg == g) cg[gg] += term;
    }
    for (int g = 0; g < G; ++g) s_coef[g * P + p] = cg[g];
  }
  __syncthreads();
  while (!s_state.done) {  // block-uniform: s_state only changes between the barriers below
    double e[kMaxEigenGroups], e1[kMaxEigenGroups], e2[kMaxEigenGroups];
#pragma unroll
    for (int g = 0; g < kMaxEigenGroups; ++g) {
      const double l = g < G ? c_model.group_lambda[g] : 0.;
      e[g] = g < G ? s_state.e[g] : 0.;
      e1[g] = l * e[g];
      e2[g] = l * l * e[g];
    }
    double f = 0., g1 = 0., g2 = 0.;
    for (int p = threadIdx.x; p < P; p += kTile) {
      double L = 0., L1 = 0., L2 = 0.;
#pragma unroll
      for (int g = 0; g < kMaxEigenGroups; ++g) {
        if (g < G) {
          const double c = s_coef[g * P + p];
          L += c * e[g];
          L1 += c * e1[g];
          L2 += c * e2[g];
        }
      }
      const double w = st.weights[p];
      f += log(L) * w;
      if (n_derivatives >= 1) g1 += (L1 / L) * w;                       // gp_engine.cpp:493-496
      if (n_derivatives >= 2) g2 += ((L2 * L - L1 * L1) / (L * L)) * w;  // gp_engine.cpp:530-538
    }
    f = block_reduce(f, SumOp(), 0.);
    if (n_derivatives >= 1) g1 = block_reduce(g1, SumOp(), 0.);
    if (n_derivatives >= 2) g2 = block_reduce(g2, SumOp(), 0.);
    if (threadIdx.x == 0) {
      OptState s = s_state;
      opt_advance(s, st, prm, f + s.ll_offset, g1, g2);
      s_state = s;
      if (s.done) atomicAdd(st.feval_total, static_cast<unsigned long long>(s.evals));
    }
    __syncthreads();
  }
  refresh_edge_matrices(st, op, refresh, threadIdx.x, kTile);  // st.bl[edge] was written before the barrier
}

// ---- OptimizeBranchLength on chip (large alignments): one thread-block CLUSTER per edge -----------
// At 1e5 patterns the ratio stream of an edge (8 B per pattern, see k_opt_prepare_ratio) is ~1 MB:
// too big for one SM, but it fits the shared memory of a cluster of 8-16 SMs. The cluster reads the
// edge's two PLVs from HBM exactly once (64 B per pattern, the algorithmic minimum of SURVEY 8d),
// keeps rho distributed over its blocks' shared memory, and runs the whole Brent search there: each
// evaluation is a pass over shared memory, a block reduction, one 8-byte store per peer block over
// distributed shared memory and one cluster barrier; every block then takes the (bitwise
// identical) optimiser decision itself, so there is no broadcast. The round-per-launch scheme
// re-streams rho from HBM for each of the ~16 evaluations instead.
// Pattern order: the cluster weight-class layout of Engine::BuildWeightClasses (classes padded to
// rows of kClusterThreads patterns; padding has inv_perm = -1 and contributes rho = 0).
// How long a kernel waits for a peer GPU before it gives up (clock64 ticks: ~35 s at 1.97 GHz). Ranks
// reach an exchange at slightly different times - one may still be compiling an op list on its
// host - so this is generous; it only exists so that a dead rank fails the others instead of
// hanging them.
constexpr long long kPeerTimeoutCycles = 1ll << 36;
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One value per rank summed across GPUs, for the edge that owns record slot `slot`: warp 0 of the
// cluster's block 0 stores (value, then tag with release) into every rank's record
// [slot][parity][my rank] over NVLink, waits for the tags of all ranks in its own buffer, and adds
// the values in rank order (the same on every GPU). Called by a full warp; returns the sum in
// every lane. `tag` grows with every exchange of this slot, so a stale record is never taken.
__device__ __forceinline__ double peer_edge_sum(const PeerEdge& px, int slot, int parity,
                                                unsigned long long tag, double v) {
  const PeerComm* pc = px.pc;
  const int R = pc->n_ranks, me = pc->rank, lane = threadIdx.x & 31;
  // bank = (search parity, round parity); tag = search number << 20 | round + 1
  const size_t bank = ((tag >> 20) & 1) * 2 + parity;
  const size_t rec0 = PeerEdgeOffsetDoubles(R) + (static_cast<size_t>(slot) * 4 + bank) * R * kPeerEdgeRecord;
  double got = 0.;
  if (lane < R) {
    double* theirs = pc->base[lane] + rec0 + kPeerEdgeRecord * me;
    *reinterpret_cast<volatile double*>(theirs) = v;
    st_release_sys(reinterpret_cast<unsigned long long*>(theirs + kPeerEdgeValues), tag);
    const double* mine = pc->base[me] + rec0 + kPeerEdgeRecord * lane;
    const long long t0 = clock64();
    // ~35 s without an answer: a peer died. Fail the call instead of hanging, and once that has
    // happened never wait again (every later exchange of this engine would time out too).
    const bool dead = (*reinterpret_cast<volatile uint32_t*>(pc->status) & kErrPeerTimeout) != 0;
    while (!dead && ld_acquire_sys(reinterpret_cast<const unsigned long long*>(mine + kPeerEdgeValues)) != tag) {
      if (clock64() - t0 > kPeerTimeoutCycles) {
        atomicOr(pc->status, kErrPeerTimeout);
        break;
      }
    }
    got = *reinterpret_cast<const volatile double*>(mine);
  }
  double acc = __shfl_sync(0xffffffffu, got, 0);
  for (int r = 1; r < R; ++r) acc += __shfl_sync(0xffffffffu, got, r);
  return acc;
}

// Sum of `v` over every thread of every block of the cluster, returned to all threads. s_slots is
// double-buffered by `parity`: a block can only be one barrier ahead of its peers.
template <int T>
__device__ __forceinline__ double cluster_sum(cg::cluster_group& cluster, double v, double* s_warp,
                                              double (*s_slots)[kMaxOptCluster], int parity,
                                              const PeerEdge& px, int slot, unsigned long long tag,
                                              double* s_global) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_down_sync(0xffffffffu, v, sh);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  const unsigned n_blocks = cluster.num_blocks();
  if (warp == 0) {
    double t = lane < T / 32 ? s_warp[lane] : 0.;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) t += __shfl_down_sync(0xffffffffu, t, sh);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (lane < static_cast<int>(n_blocks))  // lane k hands this block's sum to block k
      *cluster.map_shared_rank(&s_slots[parity][cluster.block_rank()], lane) = t;
  }
  cluster.sync();  // release/acquire: the remote stores above are visible to every block
  // fixed pairwise tree over the (at most 16) block sums: the same order in every block, and four
  // dependent additions instead of sixteen on the path every thread of the cluster waits on
  double v16[kMaxOptCluster];
#pragma unroll
  for (int k = 0; k < kMaxOptCluster; ++k)
    v16[k] = k < static_cast<int>(n_blocks) ? s_slots[parity][k] : 0.;
#pragma unroll
  for (int w = kMaxOptCluster / 2; w > 0; w >>= 1)
#pragma unroll
    for (int k = 0; k < w; ++k) v16[k] += v16[k + w];
  if (!px.enabled) return v16[0];
  // several GPUs: this GPU's total goes through NVLink, the global one comes back to every block
  if (cluster.block_rank() == 0 && warp == 0) {
    const double g = peer_edge_sum(px, slot, parity, tag, v16[0]);
    if (lane < static_cast<int>(n_blocks)) *cluster.map_shared_rank(&s_global[parity], lane) = g;
  }
  cluster.sync();
  return s_global[parity];
}

// T = threads per block: 256 (several blocks per SM: many edges resident, for levels with many
// edges) or 1024 (one edge spread over as many threads as a cluster has: the shortest time per
// edge, for the one-or-two-edge levels of a Gauss-Seidel sweep). Rows are kClusterThreads = 256
// patterns whatever T is; a block of T threads walks T / 256 rows at a time.
// kFromRho: the pipelined scheme. rho and K_e of this edge were produced by k_opt_prepare_cluster
// (rho_in: chunk buffer in the cluster layout, rho_stride doubles per edge; edge_const_in[o] = K_e),
// so the cluster copies its rows into shared memory (sequential 16-byte loads, served from L2 when
// the chunk is still resident) instead of reading the PLVs, and its shared memory is only ever
// occupied by an edge whose search is running.
template <int T, bool kFromRho>
// (pipelined variant: at most 64 registers, so that three resident blocks leave a quarter of the
// register file to the producer kernel's blocks that share the SM)
__global__ void __launch_bounds__(T, T == 256 ? (kFromRho ? 4 : 3) : (T == 512 ? 2 : 1))
    k_opt_cluster(DeviceState st, const OptOp* __restrict__ ops, const OptControl* __restrict__ ctl,
                  const int32_t* __restrict__ inv_perm, const double* __restrict__ wperm,
                  OptClusterLayout lay, OptRefresh refresh, PeerEdge px,
                  const double* __restrict__ rho_in, int64_t rho_stride,
                  const double* __restrict__ edge_const_in) {
  constexpr int S = T / kClusterThreads;           // rows walked per step
  constexpr int kStep = S * kClusterThreads;       // = T positions
  extern __shared__ __align__(16) double s_rho[];  // rows_per_block x kClusterThreads
  __shared__ OptState s_state;
  __shared__ double s_warp[T / 32];
  __shared__ double s_slots[2][kMaxOptCluster];
  __shared__ double s_global[2];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int o = blockIdx.x / cluster.num_blocks();
  // several GPUs: edge o of the level owns exchange slot o (the engine only takes this path when the
  // level has at most kPeerEdgeSlots edges and all its clusters are resident at once)
  const unsigned long long tag0 = px.enabled ? (px.seq[o] << 20) : 0ull;
  const OptOp op = ops[o];
  const OptParams prm = ctl->prm;
  if (threadIdx.x == 0) opt_init(s_state, st, prm, ctl->method, op);
  __syncthreads();
  // converged edges (dag_branch_handler.cpp:126-131): the same decision in every block
  if (s_state.done) return;
  const int row0 = rank * lay.rows_per_block;
  const int n_rows = max(0, min(lay.rows_per_block, lay.rows_total - row0));
  const int64_t q0 = static_cast<int64_t>(row0) * kClusterThreads;
  const int sub = threadIdx.x / kClusterThreads;  // this thread's rows: sub, sub + S, sub + 2 S, ...
  // the edge's PLVs, once: rho into shared memory, K_e = sum_p w_p log c0_p. Two rows per trip
  // with all four 256-bit loads issued before the first use (padding positions load nothing).
  double k_part = 0.;
  if (kFromRho) {
    const double2* src = reinterpret_cast<const double2*>(rho_in + static_cast<int64_t>(o) * rho_stride + q0);
    double2* dst = reinterpret_cast<double2*>(s_rho);
    const int n2 = n_rows * (kClusterThreads / 2);
    int i = threadIdx.x;
    for (; i + 3 * T < n2; i += 4 * T) {  // four independent 16-byte loads in flight per thread
      const double2 a = __ldcg(src + i), b = __ldcg(src + i + T), c = __ldcg(src + i + 2 * T),
                    d = __ldcg(src + i + 3 * T);
      dst[i] = a; dst[i + T] = b; dst[i + 2 * T] = c; dst[i + 3 * T] = d;
    }
    for (; i < n2; i += T) dst[i] = __ldcg(src + i);
  }
  for (int r = sub; !kFromRho && r < n_rows; r += 2 * S) {
    const int i0 = r * kClusterThreads + (threadIdx.x & (kClusterThreads - 1)), i1 = i0 + kStep;
    const bool two = r + S < n_rows;
    const int32_t pa = inv_perm[q0 + i0];
    const int32_t pb = two ? inv_perm[q0 + i1] : -1;
    V4 ra = {1., 1., 1., 1.}, ca = ra, rb = ra, cb = ra;  // padding: any finite value, masked below
    if (pa >= 0) ra = load_plv(op.parent, pa);
    if (pa >= 0) ca = load_plv(op.child, pa);
    if (pb >= 0) rb = load_plv(op.parent, pb);
    if (pb >= 0) cb = load_plv(op.child, pb);
    double rho, c0;
    ratio_coefficients(ra, ca, rho, c0);
    if (pa >= 0) k_part += wperm[q0 + i0] * log(c0);
    s_rho[i0] = pa >= 0 ? rho : 0.;
    if (two) {
      ratio_coefficients(rb, cb, rho, c0);
      if (pb >= 0) k_part += wperm[q0 + i1] * log(c0);
      s_rho[i1] = pb >= 0 ? rho : 0.;
    }
  }
  int round = 0;
  double edge_const;
  // every block of the cluster has started before the first store into a peer's shared memory (and s_rho is
  // complete before the first pass over it)
  cluster.sync();
  if (kFromRho && !px.enabled) {
    edge_const = edge_const_in[o];
  } else {
    // several GPUs: edge_const_in holds this rank's share of K_e; one exchange makes it global
    if (kFromRho) k_part = (rank == 0 && threadIdx.x == 0) ? edge_const_in[o] : 0.;
    edge_const =
        cluster_sum<T>(cluster, k_part, s_warp, s_slots, round & 1, px, o, tag0 + round + 1, s_global);
    ++round;
  }
  // this thread's rows by weight class: first row >= the class's first row that is = sub (mod S)
  int seg_begin[8], seg_end[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int b = min(max(lay.class_row_start[c] - row0, 0), n_rows);
    seg_begin[c] = b + ((sub - b) % S + S) % S;
    seg_end[c] = min(max(lay.class_row_start[c + 1] - row0, 0), n_rows);
  }
  const double* mine = s_rho + (threadIdx.x & (kClusterThreads - 1));
  for (;;) {
    const double x = s_state.x_ratio;
    double prod = 1., slow = 0.;
    int esum = 0;
    // Every factor t = 1 + rho x is either 0 or lies in [2^-53, 4] (L_p(t) >= 0 at t = 0 gives rho >= -1,
    // JC69 has rho <= 3, and 0 <= x <= 1), so a chain of up to 10 raw factors stays a normal double
    // (>= 2^-530, <= 2^20) and the product of two such chains too:
    // the chains multiply the factors as they are, and only every 32nd factor pays for the exponent
    // split (the mantissa bits are those of the product of mantissas: scaling by 2^e is exact).
    auto fold = [&](double q) {  // q = a partial product: its log goes into (prod, esum) or, if q <= 0, slow
      double m;
      int e;
      if (split_positive(q, m, e)) {
        prod *= m;  // >= 1/4
        esum += e;
        if (split_positive(prod, m, e)) { prod = m; esum += e; }
      } else {
        slow += log(q);  // log(0) = -inf as the reference's log of a zero likelihood; negative: NaN
      }
    };
    // weight 1: the bulk of any alignment; four independent chains
    {
      double p0 = 1., p1 = 1., p2 = 1., p3 = 1.;
      int r = seg_begin[0], trip = 0;
      for (; r + 3 * S < seg_end[0]; r += 4 * S, ++trip) {
        p0 *= fma(mine[(r + 0 * S) * kClusterThreads], x, 1.0);
        p1 *= fma(mine[(r + 1 * S) * kClusterThreads], x, 1.0);
        p2 *= fma(mine[(r + 2 * S) * kClusterThreads], x, 1.0);
        p3 *= fma(mine[(r + 3 * S) * kClusterThreads], x, 1.0);
        if ((trip & 7) == 7) {  // 8 factors per chain: each >= 2^-424
          fold(p0 * p1);
          fold(p2 * p3);
          p0 = p1 = p2 = p3 = 1.;
        }
      }
      for (; r < seg_end[0]; r += S) p0 *= fma(mine[r * kClusterThreads], x, 1.0);  // at most 3 more on chain 0
      fold(p0 * p1);
      fold(p2 * p3);
    }
#pragma unroll
    for (int cls = 1; cls < 7; ++cls) {  // weights 2..7: t^w by squaring (>= 2^-371), folded two at a time
      const int wi = cls + 1;
      double c = 1.;
      int k = 0;
      for (int r = seg_begin[cls]; r < seg_end[cls]; r += S, ++k) {
        const double tk = fma(mine[r * kClusterThreads], x, 1.0);
        const double t2 = tk * tk;
        double tw = (wi & 1) ? tk : 1.;
        if (wi & 2) tw *= t2;
        if (wi & 4) tw *= t2 * t2;
        c *= tw;
        if (k & 1) {
          fold(c);
          c = 1.;
        }
      }
      fold(c);
    }
    for (int r = seg_begin[7]; r < seg_end[7]; r += S) {  // general weights: explicit log
      const double w = wperm[q0 + r * kClusterThreads + (threadIdx.x & (kClusterThreads - 1))];
      if (w != 0.) slow += w * log(fma(mine[r * kClusterThreads], x, 1.0));
    }
    const double f = log(prod) + static_cast<double>(esum) * 0.6931471805599453094 + slow;
    const double total =
        cluster_sum<T>(cluster, f, s_warp, s_slots, round & 1, px, o, tag0 + round + 1, s_global);
    ++round;
    if (threadIdx.x == 0) {
      const double ll = total + s_state.ll_offset + edge_const +
                        st.total_weight * c_model.group_lambda[0] * s_state.t_eval;
      opt_advance(s_state, st, prm, ll, 0., 0.);
      if (s_state.done && rank == 0)
        atomicAdd(st.feval_total, static_cast<unsigned long long>(s_state.evals));
    }
    __syncthreads();
    if (s_state.done) break;  // every block of the cluster leaves in the same round
  }
  // st.bl[edge] was written by this block's thread 0 before the barrier above
  if (rank == 0) {
    refresh_edge_matrices(st, op, refresh, threadIdx.x, T);
    if (px.enabled && threadIdx.x == 0) px.seq[o] += 1;  // the next search of this slot uses fresh tags
  }
}

// ---- the cluster-resident search with the Taylor model (gp_types.h, OptPass) ------------------------
// Same layout and life cycle as k_opt_cluster; what changes is what one round does. A round walks the
// cluster's shared memory once and produces kClusterVec sums - the objective at the points of the pass
// (four on the first pass of a search: Brent's start and its data-independent next requests; one
// afterwards), twelve power sums about one of them, and on the first pass K_e - which cross the cluster
// (and, on several GPUs, NVLink) in ONE exchange. Every block then runs opt_consume_pass itself: all
// requests that hit a cached point or the model cost nothing, so a search is ~3 rounds instead of ~15.
constexpr int kClusterVec = kPeerEdgeValues;  // S[4], M[12], K_e

// Sums over ranks of kClusterVec values for the edge that owns record slot `slot` (see peer_edge_sum for the
// protocol: banked by (search parity, round parity), tagged search << 20 | round). Called by warp 0 of the
// cluster's block 0 with v = value `lane` of this GPU (lanes >= kClusterVec: anything); returns in lane l
// the global value l. Lane r < R ships the whole vector to rank r and waits for rank r's.
__device__ __forceinline__ double peer_edge_sum_vec(const PeerEdge& px, int slot, int parity,
                                                    unsigned long long tag, double v) {
  const PeerComm* pc = px.pc;
  const int R = pc->n_ranks, me = pc->rank, lane = threadIdx.x & 31;
  const size_t bank = ((tag >> 20) & 1) * 2 + parity;
  const size_t rec0 = PeerEdgeOffsetDoubles(R) + (static_cast<size_t>(slot) * 4 + bank) * R * kPeerEdgeRecord;
  double all[kClusterVec];
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) all[k] = __shfl_sync(0xffffffffu, v, k);
  if (lane < R) {
    double* theirs = pc->base[lane] + rec0 + kPeerEdgeRecord * me;
#pragma unroll
    for (int k = 0; k < kClusterVec; ++k) reinterpret_cast<volatile double*>(theirs)[k] = all[k];
    st_release_sys(reinterpret_cast<unsigned long long*>(theirs + kPeerEdgeValues), tag);
    const double* mine = pc->base[me] + rec0 + kPeerEdgeRecord * lane;
    const long long t0 = clock64();
    const bool dead = (*reinterpret_cast<volatile uint32_t*>(pc->status) & kErrPeerTimeout) != 0;
    while (!dead && ld_acquire_sys(reinterpret_cast<const unsigned long long*>(mine + kPeerEdgeValues)) != tag) {
      if (clock64() - t0 > kPeerTimeoutCycles) {
        atomicOr(pc->status, kErrPeerTimeout);
        break;
      }
    }
#pragma unroll
    for (int k = 0; k < kClusterVec; ++k) all[k] = reinterpret_cast<const volatile double*>(mine)[k];
  }
  // lane l collects value l of every rank, added in rank order (the same on every GPU)
  double acc = 0.;
#pragma unroll
  for (int k = 0; k < kClusterVec; ++k) {
    double sum = __shfl_sync(0xffffffffu, all[k], 0);
    for (int r = 1; r < R; ++r) sum += __shfl_sync(0xffffffffu, all[k], r);
    if (lane == k) acc = sum;
  }
  return acc;
}

template <int T>
__global__ void __launch_bounds__(T, T == 256 ? 2 : 1)
    k_opt_cluster_model(DeviceState st, const OptOp* __restrict__ ops, const OptControl* __restrict__ ctl,
                        const int32_t* __restrict__ inv_perm, const double* __restrict__ wperm,
                        OptClusterLayout lay, OptRefresh refresh, PeerEdge px, double min_weight) {
  constexpr int S = T / kClusterThreads;           // rows walked per step
  constexpr int kStep = S * kClusterThreads;       // = T positions
  constexpr int W = T / 32;
  extern __shared__ __align__(16) double s_rho[];  // rows_per_block x kClusterThreads
  __shared__ OptState s_state;
  __shared__ OptPass s_pass;
  __shared__ double s_warp[W][kClusterVec];
  __shared__ double s_block[kClusterVec];
  __shared__ double s_slots[2][kClusterVec][kMaxOptCluster];
  __shared__ double s_tot[2][kClusterVec];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int n_blocks = static_cast<int>(cluster.num_blocks());
  const int o = blockIdx.x / n_blocks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long tag0 = px.enabled ? (px.seq[o] << 20) : 0ull;
  const OptOp op = ops[o];
  const OptParams prm = ctl->prm;
  if (threadIdx.x == 0) {
    opt_init(s_state, st, prm, ctl->method, op);
    opt_first_pass(s_state, st, prm, s_pass);
  }
  __syncthreads();
  if (s_state.done) return;  // converged edges (dag_branch_handler.cpp:126-131): the same decision in every block
  const int row0 = rank * lay.rows_per_block;
  const int n_rows = max(0, min(lay.rows_per_block, lay.rows_total - row0));
  const int64_t q0 = static_cast<int64_t>(row0) * kClusterThreads;
  const int sub = threadIdx.x / kClusterThreads;  // this thread's rows: sub, sub + S, sub + 2 S, ...
  const int col = threadIdx.x & (kClusterThreads - 1);
  // the edge's PLVs, once: rho into shared memory, K_e = sum_p w_p log c0_p (as k_opt_cluster)
  double k_part = 0.;
  for (int r = sub; r < n_rows; r += 2 * S) {
    const int i0 = r * kClusterThreads + col, i1 = i0 + kStep;
    const bool two = r + S < n_rows;
    const int32_t pa = inv_perm[q0 + i0];
    const int32_t pb = two ? inv_perm[q0 + i1] : -1;
    V4 ra = {1., 1., 1., 1.}, ca = ra, rb = ra, cb = ra;  // padding: any finite value, masked below
    if (pa >= 0) ra = load_plv(op.parent, pa);
    if (pa >= 0) ca = load_plv(op.child, pa);
    if (pb >= 0) rb = load_plv(op.parent, pb);
    if (pb >= 0) cb = load_plv(op.child, pb);
    double rho, c0;
    ratio_coefficients(ra, ca, rho, c0);
    if (pa >= 0) k_part += wperm[q0 + i0] * log(c0);
    s_rho[i0] = pa >= 0 ? rho : 0.;
    if (two) {
      ratio_coefficients(rb, cb, rho, c0);
      if (pb >= 0) k_part += wperm[q0 + i1] * log(c0);
      s_rho[i1] = pb >= 0 ? rho : 0.;
    }
  }
  // s_rho is complete before the first pass over it - and every block of the cluster has started before the
  // first store into a peer's shared memory (distributed shared memory may only be touched once all blocks run)
  cluster.sync();
  // this thread's rows by weight class: first row >= the class's first row that is = sub (mod S)
  int seg_begin[8], seg_end[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int b = min(max(lay.class_row_start[c] - row0, 0), n_rows);
    seg_begin[c] = b + ((sub - b) % S + S) % S;
    seg_end[c] = min(max(lay.class_row_start[c + 1] - row0, 0), n_rows);
  }
  const double* mine = s_rho + col;
  double edge_const = 0.;
  for (int round = 0;; ++round) {
    const int parity = round & 1;
    const int np = s_pass.n_pts;
    const double c = s_pass.px[np > 1 ? s_pass.centre : 0];
    // ---- phase A: power sums about c, and the objective sum at c --------------------------------------
    double A[kOptMoments];
#pragma unroll
    for (int j = 0; j < kOptMoments; ++j) A[j] = 0.;
    ProductLog Sc;
    // four rows at a time (padding rows hold rho = 0: factor 1, z = 0); one reciprocal per four
    auto quad = [&](double r0, double r1, double r2, double r3, int wi) {
      const double t0 = fma(r0, c, 1.0), t1 = fma(r1, c, 1.0), t2 = fma(r2, c, 1.0), t3 = fma(r3, c, 1.0);
      const double t01 = t0 * t1, t23 = t2 * t3, prod4 = t01 * t23;
      const double inv = rcp_newton(prod4);
      const double i01 = inv * t23, i23 = inv * t01;
      add_moments4(A, r0 * (i01 * t1), r1 * (i01 * t0), r2 * (i23 * t3), r3 * (i23 * t2), static_cast<double>(wi));
      Sc.mul(prod4, wi);
    };
#pragma unroll 1
    for (int cls = 0; cls < 7; ++cls) {
      const int wi = cls + 1;
      int r = seg_begin[cls];
      for (; r + 3 * S < seg_end[cls]; r += 4 * S)
        quad(mine[r * kClusterThreads], mine[(r + S) * kClusterThreads], mine[(r + 2 * S) * kClusterThreads],
             mine[(r + 3 * S) * kClusterThreads], wi);
      if (r < seg_end[cls]) {
        const double r0 = mine[r * kClusterThreads];
        const double r1 = r + S < seg_end[cls] ? mine[(r + S) * kClusterThreads] : 0.;
        const double r2 = r + 2 * S < seg_end[cls] ? mine[(r + 2 * S) * kClusterThreads] : 0.;
        quad(r0, r1, r2, 0., wi);
      }
    }
    for (int r = seg_begin[7]; r < seg_end[7]; r += S) {  // general weights: one position at a time
      const double w = wperm[q0 + r * kClusterThreads + col];
      if (w == 0.) continue;
      const double rv = mine[r * kClusterThreads];
      const double tc = fma(rv, c, 1.0);
      const double z = rv * rcp_newton(tc);
      double zp = z;
#pragma unroll
      for (int j = 0; j < kOptMoments; ++j) {
        A[j] = fma(zp, w, A[j]);
        zp *= z;
      }
      Sc.slow += w * log(tc);
    }
    {
      double v16[16];
#pragma unroll
      for (int k = 0; k < kOptPoints; ++k) v16[k] = 0.;
      v16[0] = Sc.value();  // one point: the centre is the request
#pragma unroll
      for (int j = 0; j < kOptMoments; ++j) v16[kOptPoints + j] = A[j];
      warp_reduce16(v16);
      if ((lane & 1) == 0) s_warp[warp][lane >> 1] = v16[0];
      __syncwarp();
    }
    // ---- phase B (first pass of a search): the objective sums at all four points --------------------------
    if (np > 1) {
      ProductLog Sk[kOptPoints];
      double xk[kOptPoints];
#pragma unroll
      for (int k = 0; k < kOptPoints; ++k) xk[k] = s_pass.px[k];
#pragma unroll 1
      for (int cls = 0; cls < 7; ++cls) {
        const int wi = cls + 1;
        for (int r = seg_begin[cls]; r < seg_end[cls]; r += 2 * S) {
          const double r0 = mine[r * kClusterThreads];
          const double r1 = r + S < seg_end[cls] ? mine[(r + S) * kClusterThreads] : 0.;
#pragma unroll
          for (int k = 0; k < kOptPoints; ++k) Sk[k].mul(fma(r0, xk[k], 1.0) * fma(r1, xk[k], 1.0), wi);
        }
      }
      for (int r = seg_begin[7]; r < seg_end[7]; r += S) {
        const double w = wperm[q0 + r * kClusterThreads + col];
        if (w == 0.) continue;
        const double rv = mine[r * kClusterThreads];
#pragma unroll
        for (int k = 0; k < kOptPoints; ++k) Sk[k].slow += w * log(fma(rv, xk[k], 1.0));
      }
#pragma unroll
      for (int k = 0; k < kOptPoints; ++k) {
        double f = Sk[k].value();
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) f += __shfl_down_sync(0xffffffffu, f, sh);
        if (lane == 0) s_warp[warp][k] = f;
      }
    }
    {  // K_e rides on the first pass
      double f = round == 0 ? k_part : 0.;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) f += __shfl_down_sync(0xffffffffu, f, sh);
      if (lane == 0) s_warp[warp][kClusterVec - 1] = f;
    }
    __syncthreads();
    // ---- this block's totals -> every block of the cluster ---------------------------------------------------
    if (warp == 0) {
      if (lane < kClusterVec) {
        double t = s_warp[0][lane];
#pragma unroll
        for (int wv = 1; wv < W; ++wv) t += s_warp[wv][lane];
        s_block[lane] = t;
      }
    }
    __syncthreads();
    for (int b = warp; b < n_blocks; b += W)  // warp b hands this block's vector to block b
      if (lane < kClusterVec) *cluster.map_shared_rank(&s_slots[parity][lane][rank], b) = s_block[lane];
    cluster.sync();  // release/acquire: the remote stores above are visible to every block
    if (warp == 0) {
      double t = 0.;
      if (lane < kClusterVec) {
        // fixed pairwise tree over the (at most 16) block sums: the same order in every block
        double v16[kMaxOptCluster];
#pragma unroll
        for (int k = 0; k < kMaxOptCluster; ++k) v16[k] = k < n_blocks ? s_slots[parity][lane][k] : 0.;
#pragma unroll
        for (int w2 = kMaxOptCluster / 2; w2 > 0; w2 >>= 1)
#pragma unroll
          for (int k = 0; k < w2; ++k) v16[k] += v16[k + w2];
        t = v16[0];
      }
      if (!px.enabled) {
        if (lane < kClusterVec) s_tot[parity][lane] = t;
      } else if (rank == 0) {
        // several GPUs: this GPU's totals go through NVLink, the global ones come back to every block
        const double g = peer_edge_sum_vec(px, o, parity, tag0 + round + 1, t);
        if (lane < kClusterVec)
          for (int b = 0; b < n_blocks; ++b) *cluster.map_shared_rank(&s_tot[parity][lane], b) = g;
      }
    }
    if (px.enabled) cluster.sync(); else __syncthreads();
    if (threadIdx.x == 0) {
      if (round == 0) s_state.ll_offset += s_tot[parity][kClusterVec - 1];  // K_e: constant for the whole search
      double v[kOptPassValues];
#pragma unroll
      for (int k = 0; k < kOptPassValues; ++k) v[k] = s_tot[parity][k];
      opt_consume_pass(s_state, s_pass, v, 0., min_weight, st, prm);
      if (s_state.done && rank == 0) {
        atomicAdd(st.feval_total, static_cast<unsigned long long>(s_state.evals));
        atomicAdd(st.feval_total + 1, static_cast<unsigned long long>(s_pass.passes));
      }
    }
    __syncthreads();
    if (s_state.done) break;  // every block of the cluster leaves in the same round
  }
  (void)edge_const;
  // st.bl[edge] was written by this block's thread 0 before the barrier above
  if (rank == 0) {
    refresh_edge_matrices(st, op, refresh, threadIdx.x, T);
    if (px.enabled && threadIdx.x == 0) px.seq[o] += 1;  // the next search of this slot uses fresh tags
  }
}

// ---- all-reduce of per-edge scalars over NVLink peer memory --------------------------------------
// The collectives of this path carry a few scalars per edge: latency, not bandwidth.
// One block pushes its rank's n values into slot [parity][rank] of EVERY rank's exchange buffer
// (plain stores through the IPC mappings: NVLink writes), publishes them with a release store of
// the epoch number into the matching flag of every rank, waits until the flags of all ranks in its
// own buffer show this epoch, and reduces the n_ranks slots in rank order - the same order on every
// GPU, so all ranks get bitwise-identical sums and take identical optimiser / rescale decisions.
// Buffers are double-buffered by epoch parity: a rank can be at most one all-reduce ahead of a peer
// (it needs that peer's flag to finish its own), so the slot it writes is never one still being read.
constexpr int kPeerThreads = 256;
__global__ void __launch_bounds__(kPeerThreads)
    k_peer_allreduce(PeerComm pc, double* __restrict__ buf, int n, int max_op) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *pc.epoch + 1;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const int parity = static_cast<int>(epoch & 1), R = pc.n_ranks;
  const int64_t slot = (static_cast<int64_t>(parity) * R + pc.rank) * kPeerCapacity;
  for (int r = 0; r < R; ++r) {
    double* dst = pc.base[r] + slot;
    for (int i = threadIdx.x; i < n; i += kPeerThreads) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < R) {
    // flag[parity][source = me] of rank threadIdx.x, then wait for flag[parity][source = threadIdx.x] here
    unsigned long long* their = reinterpret_cast<unsigned long long*>(pc.base[threadIdx.x] + 2 * R * kPeerCapacity);
    st_release_sys(their + parity * R + pc.rank, epoch);
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(pc.base[pc.rank] + 2 * R * kPeerCapacity) + parity * R +
        threadIdx.x;
    const long long t0 = clock64();
    // ~35 s without an answer: a peer died; fail the call instead of hanging, and never wait again
    const bool dead = (*reinterpret_cast<volatile uint32_t*>(pc.status) & kErrPeerTimeout) != 0;
    while (!dead && ld_acquire_sys(mine) < epoch) {
      if (clock64() - t0 > kPeerTimeoutCycles) {
        atomicOr(pc.status, kErrPeerTimeout);
        break;
      }
    }
  }
  __syncthreads();
  const double* mine = pc.base[pc.rank] + static_cast<int64_t>(parity) * R * kPeerCapacity;
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    double acc = __ldcg(mine + i);
    for (int r = 1; r < R; ++r) {
      const double v = __ldcg(mine + r * kPeerCapacity + i);
      acc = max_op ? fmax(acc, v) : acc + v;
    }
    buf[i] = acc;
  }
  if (threadIdx.x == 0) *pc.epoch = epoch;
}
// ---- utilities ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile)
    k_export_plv(DeviceState st, PlvRef src, double* __restrict__ out) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kTile + threadIdx.x;
  if (p < st.P) st256(out + 4 * p, load_plv(src, p));
}

__global__ void k_fill(double* __restrict__ dst, int64_t n, double value) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

__global__ void k_transition_matrix(double t, double* out16) {
  if (threadIdx.x == 0 && blockIdx.x == 0) build_matrix(t, 0, 1., out16);
}

__global__ void __launch_bounds__(kTile)
    k_weighted_sum(DeviceState st, const double* __restrict__ values, double* __restrict__ partials) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kTile + threadIdx.x;
  double v = p < st.P ? values[p] * st.weights[p] : 0.;
  v = block_reduce(v, SumOp(), 0.);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

// Largest symbol of the uploaded alignment (valid symbols are 0..4, site_pattern.cpp symbol table):
// one 8-byte word per thread, the row padding beyond P masked out.
__global__ void k_max_symbol(const uint8_t* __restrict__ symbols, int64_t rows, int64_t P,
                             int64_t P_stride, unsigned* __restrict__ out) {
  const int64_t words_per_row = P_stride / 8;
  const int64_t n_words = rows * words_per_row;
  unsigned worst = 0;
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; w < n_words;
       w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = w / words_per_row;
    const int64_t col = (w - row * words_per_row) * 8;
    unsigned long long word = *reinterpret_cast<const unsigned long long*>(symbols + row * P_stride + col);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const unsigned v = static_cast<unsigned>(word & 0xffull);
      word >>= 8;
      if (col + b < P) worst = max(worst, v);
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, sh));
  if ((threadIdx.x & 31) == 0 && worst > 0) atomicMax(out, worst);
}
__global__ void k_u32_to_double(const unsigned* __restrict__ in, double* __restrict__ out) {
  *out = static_cast<double>(*in);
}

// ---- quartet hybrid marginal (gp_engine.cpp:748-808) -------------------------------------------
// The five transition matrices of every summand: rootward, sister, central, rotated, sorted.
__global__ void k_quartet_matrices(DeviceState st, const QuartetItem* __restrict__ items, int n_items,
                                   double* __restrict__ mats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 5 * n_items) return;
  const int item = i / 5, which = i - 5 * item;
  build_matrix(st.bl[items[item].edge[which]], 0, 1., mats + 16 * static_cast<int64_t>(i));
}

// One block = one 256-pattern tile of one summand. Per pattern, in the reference's order:
//   root = M_rw * plv_rw;  r_s = root o (M_sis * plv_sis);  q_s = M_central * r_s;
//   r_sorted = q_s o (M_rot * plv_rot);  L = r_sorted^T M_sorted plv_sorted;
//   partial += w_p * (log L - log(unconditional[rootward node])).
__global__ void __launch_bounds__(kTile)
    k_quartet(DeviceState st, const QuartetItem* __restrict__ items, const double* __restrict__ mats,
              int tiles, double* __restrict__ partials) {
  const int item = blockIdx.x / tiles;
  const int tile = blockIdx.x - item * tiles;
  const QuartetItem* it = items + item;
  __shared__ double sM[5][16];
  if (threadIdx.x < 80) (&sM[0][0])[threadIdx.x] = mats[80 * static_cast<int64_t>(item) + threadIdx.x];
  if (threadIdx.x == 0 && tile == 0 &&
      (st.counts[it->rootward.id] | st.counts[it->sister.id] | st.counts[it->rotated.id] |
       st.counts[it->sorted.id]) != 0)
    atomicOr(st.status, kErrQuartetRescaled);
  const double log_prior = log(st.uncond[it->rootward_node]);
  __syncthreads();
  const int64_t p = static_cast<int64_t>(tile) * kTile + threadIdx.x;
  double v = 0.;
  if (p < st.P) {
    const V4 x_rw = load_plv(it->rootward, p), x_sis = load_plv(it->sister, p);
    const V4 x_rot = load_plv(it->rotated, p), x_sorted = load_plv(it->sorted, p);
    const V4 root = matvec(sM[0], x_rw);
    const V4 sis = matvec(sM[1], x_sis);
    const V4 r_s = {root.a * sis.a, root.b * sis.b, root.c * sis.c, root.d * sis.d};
    const V4 q_s = matvec(sM[2], r_s);
    const V4 rot = matvec(sM[3], x_rot);
    const V4 r_sorted = {q_s.a * rot.a, q_s.b * rot.b, q_s.c * rot.c, q_s.d * rot.d};
    v = (log(quad(r_sorted, sM[4], x_sorted)) - log_prior) * st.weights[p];
  }
  v = block_reduce(v, SumOp(), 0.);
  if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

// out[i] = log(inverted[rootward edge] * q[sister] * q[rotated] * q[sorted]) + sums[i]
__global__ void k_quartet_finish(DeviceState st, const QuartetItem* __restrict__ items, int n_items,
                                 const double* __restrict__ sums, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const QuartetItem* it = items + i;
  out[i] = log(st.inverted[it->edge[0]] * st.q[it->edge[1]] * st.q[it->edge[3]] * st.q[it->edge[4]]) +
           sums[i];
}

inline unsigned Grid(int64_t n_ops, int64_t tiles) { return static_cast<unsigned>(n_ops * tiles); }

}  // namespace

void LaunchBuildMatrices(cudaStream_t s, const DeviceState& st, const AccumItem* items, int n_items,
                          const LikOp* liks, int n_liks, double* mtab, double* mtab_lik) {
  const int n = n_items + n_liks;
  if (n == 0) return;
  k_build_matrices<<<(n + 127) / 128, 128, 0, s>>>(st, items, n_items, liks, n_liks, mtab, mtab_lik);
}
// Pattern tiles one block walks. Small levels (the long tail of a DAG: a few nodes each) get
// exactly one wave of blocks on the 148 SMs x 4 resident blocks, each walking up to 16 tiles, so
// the per-block prologue (macro-op record, counts, matrix staging) is paid once per SM slot
// instead of once per tile; large levels use 8 tiles per block and many waves.
int TilesPerBlock(int64_t n_ops, int64_t tiles, const char* env_name, int max_tiles) {
  const char* e = getenv(env_name);
  const int forced = e != nullptr ? atoi(e) : 0;
  if (forced > 0) return forced;
  const int64_t blocks_one = n_ops * tiles;
  const int64_t slots = 148 * 4;
  int64_t t = blocks_one <= slots * 16 ? (blocks_one + slots - 1) / slots
                                       : (blocks_one >= slots * 256 ? max_tiles : 8);
  if (t < 1) t = 1;
  if (t > max_tiles) t = max_tiles;
  return static_cast<int>(t);
}
void LaunchNodes(cudaStream_t s, const DeviceState& st, const NodeOp* nodes, const AccumItem* items,
                 const int32_t* pool, const double* mtab, int n_nodes, double* level_max) {
  if (n_nodes == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const int tpb = TilesPerBlock(n_nodes, tiles, "BITO_GP_TILES_PER_BLOCK", 16);
  static const int occ = [] {
    const char* e = getenv("BITO_GP_NODE_OCC");
    return e != nullptr ? atoi(e) : 4;
  }();
  const unsigned grid = Grid(n_nodes, (tiles + tpb - 1) / tpb);
  unsigned long long* mx = reinterpret_cast<unsigned long long*>(level_max);
  static const bool stream = [] {  // BITO_GP_NODE_STORE=cs: cache-streaming stores for the PLVs a level writes
    const char* e = getenv("BITO_GP_NODE_STORE");
    return e != nullptr && e[0] == 'c';
  }();
  if (occ >= 4) {
    if (stream)
      k_node<4, true><<<grid, kTile, 0, s>>>(st, nodes, items, pool, mtab, n_nodes, tiles, tpb, mx);
    else
      k_node<4, false><<<grid, kTile, 0, s>>>(st, nodes, items, pool, mtab, n_nodes, tiles, tpb, mx);
  } else {
    k_node<3, false><<<grid, kTile, 0, s>>>(st, nodes, items, pool, mtab, n_nodes, tiles, tpb, mx);
  }
}
void LaunchRescale(cudaStream_t s, const DeviceState& st, const MultOp* ops, int n_ops,
                   const double* level_max) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const int64_t items = static_cast<int64_t>(n_ops) * tiles;
  const int grid = items < 592 ? static_cast<int>(items) : 592;  // 148 SMs x 4 resident blocks
  k_rescale<<<grid, kTile, 0, s>>>(st, ops, n_ops, level_max, tiles);
}
int64_t LikelihoodTileGroups(int n_ops, int64_t P) {
  const int64_t tiles = TilesFor(P);
  const int tpb = TilesPerBlock(n_ops, tiles, "BITO_GP_LIK_TILES_PER_BLOCK", 32);
  return (tiles + tpb - 1) / tpb;
}
void LaunchLikelihood(cudaStream_t s, const DeviceState& st, const LikOp* ops, int n_ops,
                      const double* mtab, double* partials) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const int tpb = TilesPerBlock(n_ops, tiles, "BITO_GP_LIK_TILES_PER_BLOCK", 32);
  k_likelihood<<<Grid(n_ops, (tiles + tpb - 1) / tpb), kTile, 0, s>>>(st, ops, n_ops, mtab, tiles, tpb,
                                                                     partials);
}
void LaunchMarginal(cudaStream_t s, const DeviceState& st, const MargItem* items, int n_items,
                    int reset, double* partials) {
  const int tiles = static_cast<int>(TilesFor(st.P));
  k_marginal<<<tiles, kTile, 0, s>>>(st, items, n_items, reset, tiles, partials);
}
void LaunchStationary(cudaStream_t s, const DeviceState& st, const StatOp* ops, int n_ops) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  k_stationary<<<Grid(n_ops, tiles), kTile, 0, s>>>(st, ops, tiles);
}
void LaunchZero(cudaStream_t s, const DeviceState& st, const ZeroOp* ops, int n_ops) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  k_zero<<<Grid(n_ops, tiles), kTile, 0, s>>>(st, ops, tiles);
}
void LaunchScalar(cudaStream_t s, const DeviceState& st, const ScalarOp* ops,
                  const int32_t* pool, int n_ops) {
  if (n_ops == 0) return;
  k_scalar<<<(n_ops + 63) / 64, 64, 0, s>>>(st, ops, pool, n_ops);
}
void LaunchReducePartials(cudaStream_t s, const double* partials, int n_out, int64_t tiles,
                          double* out, const int32_t* scatter_idx, double* scatter_dst) {
  if (n_out == 0) return;
  const int warps_per_block = 8;
  k_reduce_partials<<<(n_out + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                      s>>>(partials, n_out, tiles, out, scatter_idx, scatter_dst);
}
void LaunchScatter(cudaStream_t s, const double* packed, int n, const int32_t* scatter_idx,
                   double* scatter_dst) {
  if (n == 0) return;
  k_scatter<<<(n + 127) / 128, 128, 0, s>>>(packed, n, scatter_idx, scatter_dst);
}

void LaunchOptPrepare(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                      OptState* states, const OptParams& params, int method, double* coef,
                      int init_states) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  k_opt_prepare<<<Grid(n_ops, tiles), kTile, 0, s>>>(st, ops, tiles, states, params, method, coef,
                                                     init_states);
}
void LaunchOptEval(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                   const double* coef, int n_derivatives, double* partials, int n_groups) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const unsigned grid = Grid(n_ops, tiles);
  switch (n_groups) {
    case 1: k_opt_eval<1><<<grid, kTile, 0, s>>>(st, tiles, states, coef, n_derivatives, partials); break;
    case 2: k_opt_eval<2><<<grid, kTile, 0, s>>>(st, tiles, states, coef, n_derivatives, partials); break;
    case 3: k_opt_eval<3><<<grid, kTile, 0, s>>>(st, tiles, states, coef, n_derivatives, partials); break;
    default: k_opt_eval<4><<<grid, kTile, 0, s>>>(st, tiles, states, coef, n_derivatives, partials); break;
  }
}
void LaunchOptStep(cudaStream_t s, const DeviceState& st, int n_ops, OptState* states,
                   const OptParams& params, const double* sums, const double* partials, int n_parts,
                   int n_values, int value_stride, const double* edge_const,
                   int32_t* active_counter, int32_t* active, int active_capacity, int parity) {
  if (n_ops == 0) return;
  k_opt_step<<<(n_ops + 7) / 8, 256, 0, s>>>(st, n_ops, states, params, sums, partials, n_parts,
                                              n_values, value_stride, edge_const, active_counter,
                                              active, active_capacity, parity);
}
size_t OptBlockSharedBytes(int64_t P, int n_groups) {
  return static_cast<size_t>(P) * static_cast<size_t>(n_groups) * sizeof(double);
}
__global__ void k_set_opt_control(OptControl* ctl, OptControl value) { *ctl = value; }
void LaunchSetOptControl(cudaStream_t s, OptControl* ctl, const OptControl& value) {
  k_set_opt_control<<<1, 1, 0, s>>>(ctl, value);
}
void LaunchOptBlock(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                    const OptControl* ctl, int n_groups, const OptRefresh& refresh) {
  if (n_ops == 0) return;
  const size_t smem = OptBlockSharedBytes(st.P, n_groups);
  static size_t opted_in = 0;
  if (smem > 48 * 1024 && smem > opted_in) {
    cudaFuncSetAttribute(k_opt_block, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    opted_in = smem;
  }
  k_opt_block<<<n_ops, kTile, smem, s>>>(st, ops, ctl, refresh);
}
// One cluster shape of the on-chip optimiser: `threads` per block (256 or 1024), `cluster_size`
// blocks per edge. Fails when the edge's rho rows do not fit the cluster's shared memory or the
// device cannot place such a cluster; active_clusters = edges resident on the device at once.
template <int T>
static bool PlanOptClusterT(int64_t rows_total, int c, OptClusterPlan* plan, bool model) {
  static bool attrs_set = false;
  if (!attrs_set) {
    if (cudaFuncSetAttribute(k_opt_cluster<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kOptClusterMaxSharedBytes)) != cudaSuccess ||
        cudaFuncSetAttribute(k_opt_cluster<T, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
            cudaSuccess ||
        cudaFuncSetAttribute(k_opt_cluster<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kOptClusterMaxSharedBytes)) != cudaSuccess ||
        cudaFuncSetAttribute(k_opt_cluster<T, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
            cudaSuccess ||
        cudaFuncSetAttribute(k_opt_cluster_model<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(kOptClusterMaxSharedBytes)) != cudaSuccess ||
        cudaFuncSetAttribute(k_opt_cluster_model<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
            cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attrs_set = true;
  }
  const int64_t rpb = (rows_total + c - 1) / c;
  const size_t smem = static_cast<size_t>(rpb) * kClusterThreads * sizeof(double);
  if (smem > kOptClusterMaxSharedBytes) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(c) * 64u);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(c);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  // the pipelined variant has the same footprint as k_opt_cluster<T, false> (same shared memory, fewer
  // registers); the Taylor-model kernel has its own (more static shared memory)
  const cudaError_t rc = model ? cudaOccupancyMaxActiveClusters(&n, k_opt_cluster_model<T>, &cfg)
                               : cudaOccupancyMaxActiveClusters(&n, k_opt_cluster<T, false>, &cfg);
  if (rc != cudaSuccess || n < 1) {
    cudaGetLastError();
    return false;
  }
  plan->threads = T;
  plan->cluster_size = c;
  plan->rows_per_block = static_cast<int>(rpb);
  plan->rows_total = static_cast<int>(rows_total);
  plan->shared_bytes = smem;
  plan->active_clusters = n;
  return true;
}
bool PlanOptCluster(int64_t rows_total, int threads, int cluster_size, OptClusterPlan* plan, bool model) {
  *plan = OptClusterPlan();
  if (rows_total <= 0 || rows_total > (int64_t(1) << 30)) return false;
  if (cluster_size < 1 || cluster_size > kMaxOptCluster) return false;  // any size, not only powers of two
  if (threads == 256) return PlanOptClusterT<256>(rows_total, cluster_size, plan, model);
  if (threads == 512) return PlanOptClusterT<512>(rows_total, cluster_size, plan, model);
  if (threads == 1024) return PlanOptClusterT<1024>(rows_total, cluster_size, plan, model);
  return false;
}
cudaError_t LaunchOptCluster(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                             const OptControl* ctl, const int32_t* inv_perm, const double* wperm,
                             const int32_t class_row_start[9], const OptClusterPlan& plan,
                             const OptRefresh& refresh, const PeerEdge& peer, const double* rho_in,
                             int64_t rho_stride, const double* edge_const_in, bool model, double min_weight) {
  if (n_ops == 0) return cudaSuccess;
  OptClusterLayout lay;
  for (int c = 0; c < 9; ++c) lay.class_row_start[c] = class_row_start[c];
  lay.rows_total = plan.rows_total;
  lay.rows_per_block = plan.rows_per_block;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(n_ops) * static_cast<unsigned>(plan.cluster_size));
  cfg.blockDim = dim3(static_cast<unsigned>(plan.threads));
  cfg.dynamicSmemBytes = plan.shared_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(plan.cluster_size);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (model && rho_in == nullptr) {  // the Taylor-model search (the pipelined scheme keeps k_opt_cluster<T, true>)
#define GP_LAUNCH_CLUSTER_MODEL(T)                                                                          \
  return cudaLaunchKernelEx(&cfg, k_opt_cluster_model<T>, st, ops, ctl, inv_perm, wperm, lay, refresh, peer, \
                            min_weight)
    if (plan.threads == 1024) GP_LAUNCH_CLUSTER_MODEL(1024);
    if (plan.threads == 512) GP_LAUNCH_CLUSTER_MODEL(512);
    GP_LAUNCH_CLUSTER_MODEL(256);
#undef GP_LAUNCH_CLUSTER_MODEL
  }
#define GP_LAUNCH_CLUSTER(T, R)                                                                          \
  return cudaLaunchKernelEx(&cfg, k_opt_cluster<T, R>, st, ops, ctl, inv_perm, wperm, lay, refresh, peer, \
                            rho_in, rho_stride, edge_const_in)
  if (rho_in != nullptr) {
    if (plan.threads == 1024) GP_LAUNCH_CLUSTER(1024, true);
    if (plan.threads == 512) GP_LAUNCH_CLUSTER(512, true);
    GP_LAUNCH_CLUSTER(256, true);
  }
  if (plan.threads == 1024) GP_LAUNCH_CLUSTER(1024, false);
  if (plan.threads == 512) GP_LAUNCH_CLUSTER(512, false);
  GP_LAUNCH_CLUSTER(256, false);
#undef GP_LAUNCH_CLUSTER
}
void LaunchOptPrepareCluster(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops, double* rho,
                             const int32_t* cpos, int64_t rho_stride, double* partials, int max_blocks) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const int tpb = TilesPerBlock(n_ops, tiles, "BITO_GP_OPT_TILES_PER_BLOCK", 32);
  const int n_groups = (tiles + tpb - 1) / tpb;
  const int64_t items = static_cast<int64_t>(n_ops) * n_groups;
  const int64_t cap = max_blocks > 0 ? max_blocks : items;
  k_opt_prepare_cluster<<<static_cast<unsigned>(items < cap ? items : cap), kTile, 0, s>>>(
      st, ops, n_ops, tiles, tpb, n_groups, rho, cpos, rho_stride, partials);
}
int64_t OptPrepareTileGroups(int n_ops, int64_t P) {
  const int64_t tiles = TilesFor(P);
  const int tpb = TilesPerBlock(n_ops, tiles, "BITO_GP_OPT_TILES_PER_BLOCK", 32);
  return (tiles + tpb - 1) / tpb;
}
void LaunchOptPrepareRatio(cudaStream_t s, const DeviceState& st, const OptOp* ops, int n_ops,
                           OptState* states, const OptParams& params, int method, double* rho,
                           const int32_t* perm, const int32_t* pos_w, int64_t rho_stride, double* partials,
                           int32_t* active, int active_capacity) {
  if (n_ops == 0) return;
  const int tiles = static_cast<int>(TilesFor(st.P));
  const int tpb = TilesPerBlock(n_ops, tiles, "BITO_GP_OPT_TILES_PER_BLOCK", 32);
  static const bool lean = [] {  // BITO_GP_PREP_LEAN=0: the general loop for every edge (A/B)
    const char* e = getenv("BITO_GP_PREP_LEAN");
    return e == nullptr || atoi(e) != 0;
  }();
  k_opt_prepare_ratio<<<Grid(n_ops, (tiles + tpb - 1) / tpb), kTile, 0, s>>>(
      st, ops, n_ops, tiles, tpb, states, params, method, rho, perm, lean ? pos_w : nullptr, rho_stride,
      partials, active, active_capacity);
}
int64_t OptRatioTileGroups(int64_t P) {
  const int64_t per_block = static_cast<int64_t>(kTile) * kOptPatternsPerThread;
  return (P + per_block - 1) / per_block;
}
int64_t OptRatioPartials(int64_t P) { return OptRatioTileGroups(P) * (kTile / 32); }
void LaunchOptEvalRatio(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                        const double* rho, int64_t rho_stride, const double* wperm,
                        const uint8_t* row_class, double* partials, int32_t* active,
                        int active_capacity, int parity) {
  if (n_ops == 0) return;
  const int groups = static_cast<int>(OptRatioTileGroups(rho_stride));
  const int64_t items = static_cast<int64_t>(n_ops) * groups;
  static const int64_t cap = [] {  // exactly one resident wave: every SM slot walks the same share
    int per_sm = 0, sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_opt_eval_ratio, kTile, 0) != cudaSuccess ||
        per_sm < 1)
      per_sm = 4;
    return static_cast<int64_t>(per_sm) * sms;
  }();
  k_opt_eval_ratio<<<static_cast<unsigned>(items < cap ? items : cap), kTile, 0, s>>>(
      st, groups, states, rho, rho_stride, wperm, row_class, partials, active, active_capacity, parity);
}


void LaunchOptPlan(cudaStream_t s, const DeviceState& st, int n_ops, const OptState* states,
                   const OptParams& params, OptPass* pass, OptReq* req, int32_t* active) {
  if (n_ops == 0) return;
  k_opt_plan<<<(n_ops + 127) / 128, 128, 0, s>>>(st, n_ops, states, params, pass, req, active);
}
// Segments per edge of k_opt_eval_model: enough items to balance the grid, few enough that the
// block-wide reduction at the end of a segment stays small next to its 8+ tile groups.
void OptModelSegments(int64_t rho_stride, int* seg_len, int* n_seg) {
  const int groups = static_cast<int>(OptRatioTileGroups(rho_stride));
  static const int want = [] {
    const char* e = getenv("BITO_GP_OPT_SEGMENTS");
    return e != nullptr && atoi(e) > 0 ? atoi(e) : 8;
  }();
  int len = (groups + want - 1) / want;
  if (len < 1) len = 1;
  *seg_len = len;
  *n_seg = (groups + len - 1) / len;
}
void LaunchOptEvalModel(cudaStream_t s, int n_ops, int n_points, const OptReq* req, const double* rho,
                        int64_t rho_stride, const double* wperm, const OptClassStarts& classes, double* partials,
                        int32_t* active, int parity) {
  if (n_ops == 0) return;
  const int groups = static_cast<int>(OptRatioTileGroups(rho_stride));
  int seg_len = 1, n_seg = 1;
  OptModelSegments(rho_stride, &seg_len, &n_seg);
  const int64_t items = static_cast<int64_t>(n_ops) * n_seg;
  static const int occ = [] {
    const char* e = getenv("BITO_GP_OPT_EVAL_OCC");
    return e != nullptr && atoi(e) == 3 ? 3 : 2;
  }();
  static const int64_t cap = [] {  // one resident wave
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return static_cast<int64_t>(occ) * sms;
  }();
  const unsigned grid = static_cast<unsigned>(items < cap ? items : cap);
#define GP_EVAL_MODEL(NP, MINB)                                                                                 \
  k_opt_eval_model<NP, MINB><<<grid, kTile, 0, s>>>(groups, seg_len, n_seg, req, rho, rho_stride, wperm,       \
                                                    classes, partials, active, parity)
  if (n_points > 1) {
    if (occ == 3) GP_EVAL_MODEL(kOptPoints, 3); else GP_EVAL_MODEL(kOptPoints, 2);
  } else {
    if (occ == 3) GP_EVAL_MODEL(1, 3); else GP_EVAL_MODEL(1, 2);
  }
#undef GP_EVAL_MODEL
}
void LaunchOptStepModel(cudaStream_t s, const DeviceState& st, int n_ops, OptState* states, OptPass* pass,
                        OptReq* req, const OptParams& params, const double* sums, const double* partials,
                        int n_seg, const double* edge_const, double min_weight, int32_t* active, int capacity,
                        int parity) {
  if (n_ops == 0) return;
  k_opt_step_model<<<(n_ops + 7) / 8, 256, 0, s>>>(st, states, pass, req, params, sums, partials, n_seg,
                                                    edge_const, min_weight, active, capacity, parity);
}

void LaunchPeerAllReduce(cudaStream_t s, const PeerComm& pc, double* buf, int n, bool max_op) {
  if (n <= 0) return;
  k_peer_allreduce<<<1, kPeerThreads, 0, s>>>(pc, buf, n, max_op ? 1 : 0);
}

void LaunchExportPlv(cudaStream_t s, const DeviceState& st, PlvRef src, double* dense_out) {
  k_export_plv<<<static_cast<unsigned>(TilesFor(st.P)), kTile, 0, s>>>(st, src, dense_out);
}
void LaunchFill(cudaStream_t s, double* dst, int64_t n, double value) {
  if (n == 0) return;
  k_fill<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(dst, n, value);
}
void LaunchTransitionMatrix(cudaStream_t s, double t, double* out16) {
  k_transition_matrix<<<1, 32, 0, s>>>(t, out16);
}
void LaunchMaxSymbol(cudaStream_t s, const uint8_t* symbols, int64_t rows, int64_t P, int64_t P_stride,
                     double* out) {
  // *out doubles as the 4-byte accumulator (its low word), converted in place afterwards
  unsigned* acc = reinterpret_cast<unsigned*>(out);
  cudaMemsetAsync(out, 0, sizeof(double), s);
  const int64_t n_words = rows * (P_stride / 8);
  if (n_words > 0) {
    const int64_t want = (n_words + 255) / 256;
    k_max_symbol<<<static_cast<unsigned>(want < 148 * 8 ? want : 148 * 8), 256, 0, s>>>(symbols, rows, P,
                                                                                     P_stride, acc);
  }
  k_u32_to_double<<<1, 1, 0, s>>>(acc, out);
}
void LaunchWeightedSum(cudaStream_t s, const DeviceState& st, const double* values,
                       double* partials) {
  k_weighted_sum<<<static_cast<unsigned>(TilesFor(st.P)), kTile, 0, s>>>(st, values, partials);
}

void LaunchQuartetMatrices(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                           double* mats) {
  if (n_items == 0) return;
  k_quartet_matrices<<<(5 * n_items + 127) / 128, 128, 0, s>>>(st, items, n_items, mats);
}
void LaunchQuartet(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                   const double* mats, double* partials) {
  if (n_items == 0) return;
  const int64_t tiles = TilesFor(st.P);
  k_quartet<<<Grid(n_items, tiles), kTile, 0, s>>>(st, items, mats, static_cast<int>(tiles), partials);
}
void LaunchQuartetFinish(cudaStream_t s, const DeviceState& st, const QuartetItem* items, int n_items,
                         const double* sums, double* out) {
  if (n_items == 0) return;
  k_quartet_finish<<<(n_items + 127) / 128, 128, 0, s>>>(st, items, n_items, sums, out);
}

}  // namespace bito_gp
