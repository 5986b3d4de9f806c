// Fused vocabulary tail of a decode step, one 4-CTA cluster per decode row with the row's distribution
// held in (distributed) shared memory: [CaSE row finaliser: attention merge + mixture gates (Model.py:110-113, 39)] ->
// softmax of the logits x generation gate (Model.py:34, 41) -> copy scatter-add of both memories on the
// int map (Model.py:43, Utils.build_map) -> top-k (Utils.topk).  Replaces four launches and three
// passes over the 31 MB [R, V] tile (softmax_mix writes it, copy_scatter read-modify-writes it,
// topk_rows reads it) by one read of the logits; the [R, V] distribution is only written out when the
// caller asks for it (the `generate` face of the protocol).
#include "common.cuh"

namespace cb {

constexpr int TT = 256;          // threads per CTA
constexpr int TW = TT / 32;
constexpr int TCL = 4;           // CTAs per row (thread-block cluster): each owns a quarter of the vocabulary

__device__ __forceinline__ bool tail_better(float v, int i, float v2, int i2) { return v > v2 || (v == v2 && i < i2); }

__device__ __forceinline__ void tail_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tail_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tail_st_remote(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tail_st_remote(uint32_t addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// One thread-block cluster of 4 CTAs per decode row; CTA c holds vocabulary ids [c*Vq, (c+1)*Vq) of the
// row in shared memory (30 KB at V = 30522, so ~7 CTAs are resident per SM and the whole [R, V] tile is
// processed in one wave).  Cross-CTA traffic: one (max, sum) pair per CTA for the softmax, K (value,
// index) candidates per CTA for the top-k, both through distributed shared memory.
template <int K>
__global__ __launch_bounds__(TT) void row_tail_kernel(const case_tail_args_t a, int Vq, long long* dbg) {
  extern __shared__ __align__(16) float row[];           // [Vq]
  __shared__ float sh[TW * 3];
  __shared__ float st_s[2 * CASE_MAX_SPLIT * 4];         // attention split statistics of both memories
  __shared__ float e_s[2 * CASE_MAX_SPLIT];              // merge weights exp(m_j - M)
  __shared__ float xms[TCL * 2];                         // (max, sum) of every CTA of the cluster
  __shared__ float cand_v[TCL * CASE_MAX_W];             // top-k candidates of every CTA (leader only)
  __shared__ int cand_i[TCL * CASE_MAX_W];
  __shared__ float sv[TW];
  __shared__ int si[TW];
  __shared__ int swin;
  uint32_t rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int c = (int)rank_u;
  const int r = blockIdx.x / TCL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = r / a.W, V = a.V;
  const int v0 = c * Vq, nv = max(0, min(V, v0 + Vq) - v0);      // this CTA's id range [v0, v0 + nv)
  int dbg_n = 0;
  auto stamp = [&]() { if (dbg != nullptr && blockIdx.x == 0 && tid == 0) dbg[dbg_n++] = clock64(); };
  stamp();
  pdl_wait();
  stamp();
  // ---- this CTA's slice of the logits row starts landing in shared memory while the gates are computed
  {
    const float* x = a.logits + (size_t)r * a.ldl + v0;
    const uint32_t dst = smem_u32(row);
    for (int i = tid; i < (nv + 3) / 4; i += TT)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(x + i * 4) : "memory");
  }
  float g0, F[2] = {0.f, 0.f}, M[2] = {0.f, 0.f};
  if (a.do_finalize) {     // every CTA of the cluster computes the gates (rank 0 publishes them)
    const int n0 = a.ns[0] * 4, n1 = a.ns[1] * 4;
    if (tid < n0) st_s[tid] = a.stats[0][(size_t)r * n0 + tid];
    else if (tid < n0 + n1) st_s[CASE_MAX_SPLIT * 4 + tid - n0] = a.stats[1][(size_t)r * n1 + tid - n0];
    const float y = a.hN[(size_t)r * H + tid];
    __syncthreads();
    float Z[2], Q[2], cx[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ns = a.ns[i];
      const float* st = st_s + i * CASE_MAX_SPLIT * 4;
      float Mx = -INFINITY;
      for (int j = 0; j < ns; ++j) Mx = fmaxf(Mx, st[j * 4]);
      if (tid < ns) e_s[i * CASE_MAX_SPLIT + tid] = (st[tid * 4] == -INFINITY) ? 0.f : fexp(st[tid * 4] - Mx);
      M[i] = Mx;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ns = a.ns[i];
      const float* st = st_s + i * CASE_MAX_SPLIT * 4;
      const float* ee = e_s + i * CASE_MAX_SPLIT;
      const float* cp = a.ctxp[i] + (size_t)r * ns * H + tid;
      float z = 0.f, q = 0.f, acc0 = 0.f, acc1 = 0.f;
      int j = 0;
      for (; j + 1 < ns; j += 2) {
        const float x0 = cp[(size_t)j * H], x1 = cp[(size_t)(j + 1) * H];
        acc0 = fmaf(x0, ee[j], acc0);
        acc1 = fmaf(x1, ee[j + 1], acc1);
      }
      if (j < ns) acc0 = fmaf(cp[(size_t)j * H], ee[j], acc0);
      for (j = 0; j < ns; ++j) { z = fmaf(st[j * 4 + 1], ee[j], z); q = fmaf(st[j * 4 + 2], ee[j], q); }
      Z[i] = z; Q[i] = q;
      cx[i] = z > 0.f ? (acc0 + acc1) / z : 0.f;
      if (c == 0) a.ctx[i][(size_t)r * H + tid] = cx[i];
    }
    float part[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* wr = a.Wm + (size_t)k * 3 * H;
      part[k] = warp_sum(fmaf(__ldg(wr + tid), y, fmaf(__ldg(wr + H + tid), cx[0], __ldg(wr + 2 * H + tid) * cx[1])));
    }
    if (lane == 0) { sh[warp * 3] = part[0]; sh[warp * 3 + 1] = part[1]; sh[warp * 3 + 2] = part[2]; }
    __syncthreads();
    float lg[3] = {__ldg(a.bm), __ldg(a.bm + 1), __ldg(a.bm + 2)};
#pragma unroll
    for (int w = 0; w < TW; ++w) { lg[0] += sh[w * 3]; lg[1] += sh[w * 3 + 1]; lg[2] += sh[w * 3 + 2]; }
    const float mx = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
    const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx);
    const float inv = 1.f / (e0 + e1 + e2);
    g0 = e0 * inv;
    const float gi[2] = {e1 * inv, e2 * inv};
    // copy weight(r,i,s) = F_i * prior_i[s] * exp(e_i[s] - M_i) == gate_{i+1} * (w a) / (1e-8 + sum w a)
    // (Model.py:110-111, 42) with a = softmax(e): F_i = gate_{i+1} / (Z_i * (1e-8 + Q_i / Z_i))
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      F[i] = Z[i] > 0.f ? gi[i] / (Z[i] * (1e-8f + Q[i] / Z[i])) : 0.f;
      M[i] = Z[i] > 0.f ? M[i] : 0.f;
    }
    if (c == 0 && tid == 0) {
      float* gt = a.gates + (size_t)r * 4;
      gt[0] = g0; gt[1] = gi[0]; gt[2] = gi[1]; gt[3] = 0.f;
      float* f = a.fac + (size_t)r * a.fac_ld;
      f[a.fac_off[0]] = F[0]; f[a.fac_off[0] + 1] = M[0];
      f[a.fac_off[1]] = F[1]; f[a.fac_off[1] + 1] = M[1];
    }
  } else {
    g0 = a.gates[(size_t)r * 4];
    for (int i = 0; i < a.nmem; ++i) {
      F[i] = a.fac[(size_t)r * a.fac_ld + a.fac_off[i]];
      M[i] = a.fac[(size_t)r * a.fac_ld + a.fac_off[i] + 1];
    }
  }
  stamp();
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  stamp();
  if (a.mask_col0 && c == 0 && tid == 0) row[0] = -INFINITY;       // GTTP/Model.py:26
  __syncthreads();
  // ---- softmax x gate: local (max, sum) per CTA, merged over the cluster
  float m = -INFINITY;
  {
    int i = tid;
    for (; i + 3 * TT < nv; i += 4 * TT)
      m = fmaxf(fmaxf(m, row[i]), fmaxf(fmaxf(row[i + TT], row[i + 2 * TT]), row[i + 3 * TT]));
    for (; i < nv; i += TT) m = fmaxf(m, row[i]);
  }
  m = warp_max(m);
  if (lane == 0) sv[warp] = m;
  __syncthreads();
  m = sv[0];
#pragma unroll
  for (int w = 1; w < TW; ++w) m = fmaxf(m, sv[w]);
  float s = 0.f;
  {
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = tid;
    const float mm = (m > -INFINITY) ? m : 0.f;      // an all -inf slice gives exp(-inf - 0) = 0, never NaN
    for (; i + 3 * TT < nv; i += 4 * TT) {
      const float x0 = row[i], x1 = row[i + TT], x2 = row[i + 2 * TT], x3 = row[i + 3 * TT];
      const float y0 = fexp(x0 - mm), y1 = fexp(x1 - mm), y2 = fexp(x2 - mm), y3 = fexp(x3 - mm);
      row[i] = y0; row[i + TT] = y1; row[i + 2 * TT] = y2; row[i + 3 * TT] = y3;
      s += y0; s1 += y1; s2 += y2; s3 += y3;
    }
    for (; i < nv; i += TT) {
      const float y0 = fexp(row[i] - mm);
      row[i] = y0;
      s += y0;
    }
    s += s1 + s2 + s3;
  }
  s = warp_sum(s);
  __syncthreads();                 // sv is reused
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float t2 = 0.f;
#pragma unroll
    for (int w = 0; w < TW; ++w) t2 += sh[w];
    const uint32_t la = smem_u32(xms + 2 * c);
#pragma unroll
    for (int k = 0; k < TCL; ++k) {
      tail_st_remote(tail_mapa(la, k), m);
      tail_st_remote(tail_mapa(la + 4, k), t2);
    }
  }
  stamp();
  tail_cluster_sync();
  stamp();
  float scl;
  {
    float MM = -INFINITY;
#pragma unroll
    for (int k = 0; k < TCL; ++k) MM = fmaxf(MM, xms[2 * k]);
    float SS = 0.f;
#pragma unroll
    for (int k = 0; k < TCL; ++k) SS += (xms[2 * k] > -INFINITY) ? xms[2 * k + 1] * fexp(xms[2 * k] - MM) : 0.f;
    scl = (m > -INFINITY) ? g0 * fexp(m - MM) / SS : 0.f;
  }
  {
    int i = tid;
    for (; i + 3 * TT < nv; i += 4 * TT) {
      const float x0 = row[i], x1 = row[i + TT], x2 = row[i + 2 * TT], x3 = row[i + 3 * TT];
      row[i] = x0 * scl; row[i + TT] = x1 * scl; row[i + 2 * TT] = x2 * scl; row[i + 3 * TT] = x3 * scl;
    }
    for (; i < nv; i += TT) row[i] *= scl;
  }
  __syncthreads();
  stamp();
  // ---- copy scatter restricted to this CTA's id range: dist[map[b, s]] += F * prior[b, s] * exp(e[r, s] - M)
  for (int i = 0; i < a.nmem; ++i) {
    if (F[i] == 0.f) continue;
    const int S = a.S[i];
    const float* at = a.attn_un[i] + (size_t)r * S;
    const float* pr = a.prior[i] ? a.prior[i] + (size_t)b * S : nullptr;
    const int32_t* mp = a.map + (size_t)b * a.map_ld + a.map_off[i];
    for (int s0 = tid; s0 < S; s0 += 4 * TT) {
      int id[4];
      float ev[4], pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sidx = s0 + u * TT;
        id[u] = sidx < S ? __ldg(mp + sidx) - v0 : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sidx = s0 + u * TT;
        const bool mine = (unsigned)id[u] < (unsigned)nv;
        ev[u] = mine ? at[sidx] : -INFINITY;
        pv[u] = (mine && pr) ? __ldg(pr + sidx) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ev[u] == -INFINITY) continue;                 // masked source position, or another CTA's id
        const float cw = F[i] * pv[u] * fexp(ev[u] - M[i]);
        if (cw != 0.f) atomicAdd(row + id[u], cw);
      }
    }
  }
  __syncthreads();
  stamp();
  if (a.dist != nullptr) {
    float* d = a.dist + (size_t)r * a.ldd + v0;
    for (int i = tid; i < nv; i += TT) d[i] = row[i];
  }
  if (a.top_idx == nullptr) return;      // uniform over the cluster; no remote access is outstanding
  // ---- top-k: per-thread sorted list over ascending indices (strict > keeps the lower index on ties),
  // K rounds of block arg-max -> this CTA's K candidates -> leader merges the 4 K candidates
  float tv[K];
  int ti[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { tv[k] = -INFINITY; ti[k] = 0x7fffffff; }
  auto push = [&](float v, int i) {
    if (v > tv[K - 1]) {
      tv[K - 1] = v; ti[K - 1] = i;
#pragma unroll
      for (int k = K - 1; k > 0; --k) {
        if (tv[k] > tv[k - 1]) {
          const float fv = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = fv;
          const int fi = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = fi;
        }
      }
    }
  };
  {
    int i = tid;
    for (; i + 3 * TT < nv; i += 4 * TT) {
      const float x0 = row[i], x1 = row[i + TT], x2 = row[i + 2 * TT], x3 = row[i + 3 * TT];
      push(x0, v0 + i); push(x1, v0 + i + TT); push(x2, v0 + i + 2 * TT); push(x3, v0 + i + 3 * TT);
    }
    for (; i < nv; i += TT) push(row[i], v0 + i);
  }
  stamp();
  const uint32_t cv0 = tail_mapa(smem_u32(cand_v + c * CASE_MAX_W), 0), ci0 = tail_mapa(smem_u32(cand_i + c * CASE_MAX_W), 0);
  for (int round = 0; round < a.K; ++round) {
    float bv = tv[0];
    int bi = ti[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (tail_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float wv = sv[0];
      int wi = si[0];
#pragma unroll
      for (int w = 1; w < TW; ++w)
        if (tail_better(sv[w], si[w], wv, wi)) { wv = sv[w]; wi = si[w]; }
      tail_st_remote(cv0 + 4 * round, wv);
      tail_st_remote(ci0 + 4 * round, wi);
      swin = wi;
    }
    __syncthreads();
    if (ti[0] == swin) {   // owner pops its head
#pragma unroll
      for (int k = 0; k < K - 1; ++k) { tv[k] = tv[k + 1]; ti[k] = ti[k + 1]; }
      tv[K - 1] = -INFINITY; ti[K - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
  stamp();
  tail_cluster_sync();
  stamp();
  if (c == 0 && tid == 0) {              // every CTA's list is sorted: a K-round 4-way merge
    int head[TCL] = {0, 0, 0, 0};
    for (int round = 0; round < a.K; ++round) {
      int best = 0;
      float wv = -INFINITY;
      int wi = 0x7fffffff;
#pragma unroll
      for (int k = 0; k < TCL; ++k) {
        if (head[k] < a.K) {
          const float v = cand_v[k * CASE_MAX_W + head[k]];
          const int ix = cand_i[k * CASE_MAX_W + head[k]];
          if (tail_better(v, ix, wv, wi)) { wv = v; wi = ix; best = k; }
        }
      }
#pragma unroll
      for (int k = 0; k < TCL; ++k) head[k] += (k == best);
      a.top_vals[(size_t)r * a.K + round] = wv;
      a.top_idx[(size_t)r * a.K + round] = wi;
    }
  }
}

}  // namespace cb

using namespace cb;

static thread_local long long* g_tail_dbg = nullptr;   // debugging aid of the calling thread
/* debugging aid (not part of the stable ABI): clock64() stamps of CTA 0 at the phase boundaries */
extern "C" int case_debug_tail_timing(void* buf) { g_tail_dbg = (long long*)buf; return 0; }

extern "C" int case_row_tail_max_vocab(void) { return TCL * ((200 * 1024) / 4); }

extern "C" int case_row_tail(const case_tail_args_t* a, case_stream_t stream) {
  CB_REQUIRE(a && a->logits && a->gates && a->fac && a->map, "case_row_tail: null pointer");
  CB_REQUIRE(a->R > 0 && a->W >= 1 && a->V > 0 && a->V <= case_row_tail_max_vocab(), "case_row_tail: bad sizes (V too large for the cluster's shared memory)");
  CB_REQUIRE(a->ldl % 4 == 0 && a->ldl >= ((a->V + 3) / 4) * 4 && (uintptr_t)a->logits % 16 == 0, "case_row_tail: logits rows must be 16-byte aligned and padded to a multiple of 4");
  CB_REQUIRE(a->nmem >= 1 && a->nmem <= 2, "case_row_tail: nmem must be 1 or 2");
  CB_REQUIRE(!a->do_finalize || (a->nmem == 2 && a->hN && a->stats[0] && a->stats[1] && a->ctxp[0] && a->ctxp[1] && a->Wm && a->bm && a->ctx[0] && a->ctx[1] && a->ns[0] >= 1 && a->ns[1] >= 1 && a->ns[0] <= CASE_MAX_SPLIT && a->ns[1] <= CASE_MAX_SPLIT),
             "case_row_tail: the CaSE finaliser needs hN, stats, ctxp, Wm, bm, ctx for both memories");
  CB_REQUIRE(a->top_idx == nullptr || (a->top_vals && a->K >= 1 && a->K <= CASE_MAX_W && a->K <= a->V), "case_row_tail: k out of range");
  CB_REQUIRE(a->top_idx != nullptr || a->dist != nullptr, "case_row_tail: nothing to produce");
  for (int i = 0; i < a->nmem; ++i) CB_REQUIRE(a->attn_un[i] && a->S[i] > 0, "case_row_tail: attn_un / S missing");
  const int Vq = (((a->V + TCL - 1) / TCL) + 3) / 4 * 4;          // ids per CTA, a multiple of 4 (16-byte copies)
  const size_t smem = (size_t)Vq * 4;
  const int K = a->top_idx ? a->K : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a->R * TCL); cfg.blockDim = dim3(TT); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = TCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = launch_opts().pdl ? 2 : 1;
  long long* dbg = g_tail_dbg;
  void* pa[] = {(void*)a, (void*)&Vq, (void*)&dbg};
#define TAIL_LAUNCH(KK)                                                                                        \
  do {                                                                                                          \
    ensure_smem<row_tail_kernel<KK>>(200 * 1024); \
    launch_err() = cudaLaunchKernelExC(&cfg, (const void*)row_tail_kernel<KK>, pa);                            \
  } while (0)
  if (K == 1) TAIL_LAUNCH(1);
  else if (K == 2) TAIL_LAUNCH(2);
  else if (K <= 4) TAIL_LAUNCH(4);
  else TAIL_LAUNCH(8);
#undef TAIL_LAUNCH
  return check_launch("case_row_tail");
}
