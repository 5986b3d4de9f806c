// Exact block-level top-k (value descending, ties -> lower index first: Utils.topk, Utils.py:156-168) for
// kernels that see every element once and cannot afford per-element bookkeeping:
//   1. every thread reduces its elements to their maximum (a by-product of the softmax);
//   2. block_kth_max(): T = the k-th largest of the thread maxima.  The k largest maxima are k distinct
//      elements >= T, so every element of the true top-k is >= T;
//   3. threads append their elements >= T to a small shared-memory buffer (typically k .. 2k entries);
//   4. one warp selects the top-k of the buffer with a register-resident sorted list (WarpTopK).
// Pathological ties (more than `cap` elements >= T) are reported to the caller, which falls back to
// offering every element to per-warp lists.
#pragma once
#include "common.cuh"

namespace cb {

__device__ __forceinline__ bool tk_better(float v, int i, float v2, int i2) { return v > v2 || (v == v2 && i < i2); }

// ---- branch-free warp bitonic sorts (descending: lane 0 ends up with the largest)
__device__ __forceinline__ float warp_sort_desc(float v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0) == ((lane & j) == 0);   // this lane keeps the larger of the pair
      v = up ? fmaxf(v, o) : fminf(v, o);
    }
  }
  return v;
}
// merge step for two descending sequences a (this warp's registers) and b: the 32 largest of the union
__device__ __forceinline__ float warp_merge_top32(float a, float b) {
  const int lane = threadIdx.x & 31;
  float v = fmaxf(a, __shfl_sync(0xffffffffu, b, 31 - lane));    // bitonic sequence holding the top 32
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, v, j);
    v = ((lane & j) == 0) ? fmaxf(v, o) : fminf(v, o);
  }
  return v;
}
// 64-bit sort key of (value, index): larger key = better (value descending, then index ascending)
__device__ __forceinline__ unsigned long long tk_key(float v, int i) {
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);               // order-preserving map of floats to unsigned
  return ((unsigned long long)b << 32) | (uint32_t)(0x7fffffff - i);
}
__device__ __forceinline__ void tk_unkey(unsigned long long k, float& v, int& i) {
  uint32_t b = (uint32_t)(k >> 32);
  b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
  v = __uint_as_float(b);
  i = 0x7fffffff - (int)(uint32_t)(k & 0xffffffffu);
}
__device__ __forceinline__ unsigned long long warp_sort_desc_u64(unsigned long long v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0) == ((lane & j) == 0);
      v = up ? (v > o ? v : o) : (v < o ? v : o);
    }
  }
  return v;
}

// Lane k of a warp holds the entry of rank k (k < K <= 32) in registers.
struct WarpTopK {
  float ev;      // this lane's entry (lanes >= K: -inf)
  int ei;
  float thr_v;   // the K-th entry (warp-uniform copy)
  int thr_i;
  __device__ __forceinline__ void init() { ev = -INFINITY; ei = 0x7fffffff; thr_v = -INFINITY; thr_i = 0x7fffffff; }

  // every lane offers (x, ix); must be called by the whole warp.  Candidates are inserted one at a time,
  // re-voting after each insertion, so the loop runs once per entry that really enters the list.
  __device__ __forceinline__ void offer(int K, float x, int ix) {
    const int lane = threadIdx.x & 31;
    while (true) {
      const unsigned bal = __ballot_sync(0xffffffffu, tk_better(x, ix, thr_v, thr_i));
      if (bal == 0u) break;
      const int src = __ffs(bal) - 1;
      const float cx = __shfl_sync(0xffffffffu, x, src);
      const int ci = __shfl_sync(0xffffffffu, ix, src);
      if (lane == src) { x = -INFINITY; ix = 0x7fffffff; }      // consumed
      const unsigned ahead = __ballot_sync(0xffffffffu, lane < K && tk_better(ev, ei, cx, ci));
      const int p = __popc(ahead);                                // entries ahead of the candidate form a prefix
      const float pv = __shfl_up_sync(0xffffffffu, ev, 1);
      const int pi = __shfl_up_sync(0xffffffffu, ei, 1);
      if (lane < K) {
        if (lane == p) { ev = cx; ei = ci; }
        else if (lane > p) { ev = pv; ei = pi; }
      }
      thr_v = __shfl_sync(0xffffffffu, ev, K - 1);
      thr_i = __shfl_sync(0xffffffffu, ei, K - 1);
    }
  }
};

// Shared scratch of the block-level helpers (NW = warps per CTA, K <= 16, CAP = candidate buffer size)
template <int NW, int CAP>
struct TopKScratch {
  float lists_v[NW * 16];
  int lists_i[NW * 16];
  float buf_v[CAP];
  int buf_i[CAP];
  int count;
  float thr;
};

// k-th largest of the per-thread values `mine` (one per thread, all threads call).  Two barriers.
template <int NW, int CAP>
__device__ __forceinline__ float block_kth_max(TopKScratch<NW, CAP>& sc, int K, float mine) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float sorted = warp_sort_desc(mine);
  if (lane < K) sc.lists_v[warp * 16 + lane] = sorted;
  if (threadIdx.x == 0) sc.count = 0;
  __syncthreads();
  if (warp == 0) {
    float cur = -INFINITY;                                        // descending top-32 so far
    for (int base = 0; base < NW * K; base += 32) {              // NW * K values, 32 per round
      const int j = base + lane;
      const float x = warp_sort_desc(j < NW * K ? sc.lists_v[(j / K) * 16 + (j % K)] : -INFINITY);
      cur = base == 0 ? x : warp_merge_top32(cur, x);
    }
    const float t = __shfl_sync(0xffffffffu, cur, K - 1);
    if (lane == 0) sc.thr = t;
  }
  __syncthreads();
  return sc.thr;
}

// append one candidate (any thread, any time between block_kth_max and block_select)
template <int NW, int CAP>
__device__ __forceinline__ void topk_append(TopKScratch<NW, CAP>& sc, float v, int i) {
  const int pos = atomicAdd(&sc.count, 1);
  if (pos < CAP) { sc.buf_v[pos] = v; sc.buf_i[pos] = i; }
}

// After a barrier that follows the last append: warp 0 selects the top-K of the buffer; lane k < K of
// warp 0 returns rank k in (out_v, out_i).  Returns false (for every thread) when the buffer overflowed.
template <int NW, int CAP>
__device__ __forceinline__ bool block_select(TopKScratch<NW, CAP>& sc, int K, float& out_v, int& out_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = sc.count;
  if (n > CAP) return false;
  if (warp == 0) {
    if (n <= 32) {                                                // the common case: one branch-free sort
      const unsigned long long key = warp_sort_desc_u64(lane < n ? tk_key(sc.buf_v[lane], sc.buf_i[lane]) : 0ull);
      tk_unkey(key, out_v, out_i);
      if (key == 0ull) { out_v = -INFINITY; out_i = 0x7fffffff; }
    } else {
      WarpTopK wl;
      wl.init();
      for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        wl.offer(K, j < n ? sc.buf_v[j] : -INFINITY, j < n ? sc.buf_i[j] : 0x7fffffff);
      }
      out_v = wl.ev;
      out_i = wl.ei;
    }
  }
  return true;
}

// Fallback for massive ties: every warp has offered all of its elements to `wl`; merge the NW lists.
// Afterwards lane k < K of warp 0 holds rank k.  Two barriers.
template <int NW, int CAP>
__device__ __forceinline__ void block_merge_lists(TopKScratch<NW, CAP>& sc, WarpTopK& wl, int K) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane < K) { sc.lists_v[warp * 16 + lane] = wl.ev; sc.lists_i[warp * 16 + lane] = wl.ei; }
  __syncthreads();
  if (warp == 0) {
    wl.init();
    for (int w = 0; w < NW; ++w) {
      wl.offer(K, lane < K ? sc.lists_v[w * 16 + lane] : -INFINITY, lane < K ? sc.lists_i[w * 16 + lane] : 0x7fffffff);
    }
  }
}

}  // namespace cb
