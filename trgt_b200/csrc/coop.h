// coop.h -- cooperative-group shim shared by every kernel core in this directory.
//
// The DP cores (hmm_core.h, wfa_core.h) are written once, as templates over a "group" of
// lanes that cooperate on ONE work item: a warp (WarpGroup), a whole CTA (BlockGroup) or,
// in the CPU unit tests only, a single host lane (SerialGroup, tests/emul/).  All
// cross-lane communication goes through memory followed by g.sync(), or through the
// reductions below, so a loop of the form
//     for (int i = g.lane(); i < n; i += g.size()) { ... independent iterations ... }
//     g.sync();
// has identical semantics on the device and in the serial test build.  The serial build
// exists so the index arithmetic and tie-breaking of the cores can be checked against the
// oracle on a machine without a GPU; it is never part of libtrgt_b200.so.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TRGT_HD __host__ __device__ __forceinline__
#define TRGT_D __device__ __forceinline__
#else
#define TRGT_HD inline
#endif

namespace trgt {

// One lane doing everything (CPU tests; also usable for a thread-per-item kernel).
struct SerialGroup {
  TRGT_HD int lane() const { return 0; }
  TRGT_HD int size() const { return 1; }
  TRGT_HD void sync() const {}
  TRGT_HD int min_i(int v) const { return v; }
  TRGT_HD int max_i(int v) const { return v; }
  TRGT_HD int any(int p) const { return p; }
  TRGT_HD int bcast0(int v) const { return v; }
  TRGT_HD int bcast(int v, int /*src_lane*/) const { return v; }
  // exclusive prefix sum over the lanes; *total = sum over all lanes
  TRGT_HD int excl_scan_i(int v, int *total) const { *total = v; return 0; }
};

#if defined(__CUDACC__)
// 32 lanes of one warp; several warps of a CTA each own a different item.
struct WarpGroup {
  TRGT_D int lane() const { return (int)(threadIdx.x & 31u); }
  TRGT_D int size() const { return 32; }
  TRGT_D void sync() const { __syncwarp(); }
  TRGT_D int min_i(int v) const { return __reduce_min_sync(0xffffffffu, v); }
  TRGT_D int max_i(int v) const { return __reduce_max_sync(0xffffffffu, v); }
  TRGT_D int any(int p) const { return __any_sync(0xffffffffu, p); }
  TRGT_D int bcast0(int v) const { return __shfl_sync(0xffffffffu, v, 0); }
  TRGT_D int bcast(int v, int src_lane) const { return __shfl_sync(0xffffffffu, v, src_lane); }
  TRGT_D int excl_scan_i(int v, int *total) const {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if ((int)(threadIdx.x & 31u) >= d) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
  }
};

// N consecutive lanes of a warp (N = 2, 4, 8, 16) on one item: a warp works on 32 / N items at once,
// each tile following its own control flow (independent thread scheduling); every collective names
// the tile's own lanes in its mask.
template <int N>
struct TileGroup {
  unsigned mask;
  int l;
  TRGT_D TileGroup() {
    const unsigned lane = threadIdx.x & 31u;
    l = (int)(lane & (unsigned)(N - 1));
    mask = (unsigned)((1ull << N) - 1ull) << (lane & ~(unsigned)(N - 1));
  }
  TRGT_D int lane() const { return l; }
  TRGT_D int size() const { return N; }
  TRGT_D void sync() const { __syncwarp(mask); }
  TRGT_D int min_i(int v) const { return __reduce_min_sync(mask, v); }
  TRGT_D int max_i(int v) const { return __reduce_max_sync(mask, v); }
  TRGT_D int any(int p) const { return __any_sync(mask, p); }
  TRGT_D int bcast0(int v) const { return __shfl_sync(mask, v, 0, N); }
  TRGT_D int bcast(int v, int src_lane) const { return __shfl_sync(mask, v, src_lane, N); }
  TRGT_D int excl_scan_i(int v, int *total) const {
    int x = v;
#pragma unroll
    for (int d = 1; d < N; d <<= 1) {
      const int y = __shfl_up_sync(mask, x, d, N);
      if (l >= d) x += y;
    }
    *total = __shfl_sync(mask, x, N - 1, N);
    return x - v;
  }
  // bit i = predicate of the tile's lane i
  TRGT_D unsigned ballot(int p) const {
    return (__ballot_sync(mask, p) >> ((threadIdx.x & 31u) & ~(unsigned)(N - 1))) & (unsigned)((1ull << N) - 1ull);
  }
};

// The whole CTA (blockDim.x threads, a multiple of 32, <= 1024) on one item.
// `scratch` points at 33 ints of shared memory owned by the group.
struct BlockGroup {
  int *scratch;
  TRGT_D explicit BlockGroup(int *s) : scratch(s) {}
  TRGT_D int lane() const { return (int)threadIdx.x; }
  TRGT_D int size() const { return (int)blockDim.x; }
  TRGT_D void sync() const { __syncthreads(); }
  TRGT_D int min_i(int v) const {
    v = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = scratch[0];
    for (unsigned w = 1; w < (blockDim.x >> 5); w++) r = min(r, scratch[w]);
    __syncthreads();
    return r;
  }
  TRGT_D int max_i(int v) const {
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = scratch[0];
    for (unsigned w = 1; w < (blockDim.x >> 5); w++) r = max(r, scratch[w]);
    __syncthreads();
    return r;
  }
  TRGT_D int any(int p) const { return __syncthreads_or(p); }
  // exclusive prefix sum over the CTA's threads (scan within each warp, then across the warps' totals)
  TRGT_D int excl_scan_i(int v, int *total) const {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if ((int)(threadIdx.x & 31u) >= d) x += y;
    }
    const unsigned w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31u) == 31u) scratch[w] = x;
    __syncthreads();
    int base = 0, tot = 0;
    for (unsigned i = 0; i < nw; i++) {
      if (i < w) base += scratch[i];
      tot += scratch[i];
    }
    __syncthreads();
    *total = tot;
    return base + x - v;
  }
  TRGT_D int bcast0(int v) const { return bcast(v, 0); }
  TRGT_D int bcast(int v, int src_lane) const {
    if ((int)threadIdx.x == src_lane) scratch[32] = v;
    __syncthreads();
    int r = scratch[32];
    __syncthreads();
    return r;
  }
};
#endif

}  // namespace trgt
