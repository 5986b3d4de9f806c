// Cross-attention over a COMPACTED memory with a balanced static partition (bf16 K/V).
//
// Padding keys contribute exactly zero to softmax(QK^T)V (TransformerDecoder.py:81 masks them with -inf),
// and attention does not depend on key order, so the prefill packs only the valid keys of every query,
// contiguously, into the K|V tile stream (case_pack_kv_tiles_gather): at CAsT-shaped inputs ~30 % fewer
// bytes per step.  Queries then have different tile counts, so the (query, head, tile) stream - T tiles in
// all, T only known on the device - is cut into NW equal contiguous ranges, one per warp of a persistent
// grid (148 CTAs x 12 warps: every SM streams, nothing waits for a straggler).  A warp walks its range
// through a 2-stage ring of 8 KB bulk copies (12 x 2 rather than 8 x 3: the same 192 KB in flight per SM, but a tile's
// dependent wait -> ldmatrix -> MMA -> softmax -> MMA chain hides behind 11 other warps instead of 7: 25.2 -> 24.4 us
// per launch in the decode graph) that runs ahead ACROSS (query, head) boundaries; for every
// (query, head) it touches it writes one flash-decoding partial.  The partials of a (query, head) are
// the consecutive warps that share its tiles, slot = warp - first warp; unused slots are filled with
// (m = -inf, l = 0) by the first of them, so case_layer_chain merges a fixed number of slots.
#include "common.cuh"

namespace cb {

constexpr int XP_WARPS = 12;
constexpr int XP_TILE = 64;
constexpr int XP_TILE_BYTES = XP_TILE * HD * 2;     // 4 KB of K, then 4 KB of V
constexpr int XP_STAGE = 2 * XP_TILE_BYTES;
constexpr int XP_NS = 2;
constexpr int XP_WARP_BYTES = XP_NS * XP_STAGE + 64;
constexpr int XP_MIN_TILES = 5;                     // a warp's range is never shorter (bounds the slot count)

__device__ __forceinline__ uint32_t xp_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void xp_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void xp_ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xp_ldsm4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xp_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xp_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ bool xp_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t xp_swz(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

// position in the (query, head, tile) stream
struct XpPos {
  int b, head, tile, nt;          // nt = tiles of query b
};

// tp: per-head tile prefix [B + 1] (shared memory).  g in [0, NH * tp[B])
__device__ __forceinline__ XpPos xp_locate(const int* tp, int B, long long g) {
  int lo = 0, hi = B;              // largest b with NH * tp[b] <= g
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)NH * tp[mid] <= g) lo = mid; else hi = mid;
  }
  XpPos p;
  p.b = lo;
  p.nt = tp[lo + 1] - tp[lo];
  const int rem = (int)(g - (long long)NH * tp[lo]);
  p.head = rem / p.nt;
  p.tile = rem - p.head * p.nt;
  return p;
}
__device__ __forceinline__ void xp_advance(XpPos& p, const int* tp, int B) {   // next tile of the stream
  if (++p.tile < p.nt) return;
  p.tile = 0;
  if (++p.head < NH) return;
  p.head = 0;
  do { ++p.b; } while (p.b < B && tp[p.b + 1] == tp[p.b]);     // queries without valid keys own no tiles
  p.nt = p.b < B ? tp[p.b + 1] - tp[p.b] : 1;
}

__global__ __launch_bounds__(XP_WARPS * 32) void cross_attn_part_kernel(
    const float* __restrict__ q2, const bf16* __restrict__ KV, const int32_t* __restrict__ ncount,
    const int32_t* __restrict__ tile_prefix, int B, int W, int ntile_all, int nslot, float* __restrict__ part_ml,
    float* __restrict__ part_acc, int evict) {
  const uint64_t pol = l2_stream_policy(evict);
  extern __shared__ __align__(128) unsigned char xp_smem[];   // [warp][XP_NS stages of K|V][barriers], then tp[B+1]
  int* tp = reinterpret_cast<int*>(xp_smem + (size_t)XP_WARPS * XP_WARP_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g4 = lane >> 2, t4 = lane & 3;
  unsigned char* wsm = xp_smem + (size_t)warp * XP_WARP_BYTES;
  const uint32_t sbase = smem_u32(wsm), bars = sbase + XP_NS * XP_STAGE;
  if (lane == 0) {
    for (int i = 0; i < XP_NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i <= B; i += blockDim.x) tp[i] = tile_prefix[i];   // written by the prefill, not by a neighbour launch
  __syncthreads();
  const long long T = (long long)NH * tp[B];
  // the effective number of ranges: never fewer than XP_MIN_TILES tiles per warp
  const int NWall = gridDim.x * XP_WARPS;
  long long NWl = T / XP_MIN_TILES;
  const int NW = (int)(NWl < 1 ? 1 : (NWl > NWall ? NWall : NWl));
  const int w = blockIdx.x * XP_WARPS + warp;
  // spread the active ranges over the CTAs: range index r = w when NW == NWall; otherwise warps of the
  // same CTA take ranges NCTA apart so that every SM keeps streaming
  const int r = warp * gridDim.x + blockIdx.x;
  auto bnd = [&](int k) -> long long { return ((long long)k * T) / NW; };
  auto range_of = [&](long long g) -> int { return (int)((((g + 1) * NW + T - 1) / T) - 1); };
  (void)w;
  pdl_trigger();
  const bool active = r < NW && T > 0;
  const long long g_begin = active ? bnd(r) : 0, g_end = active ? bnd(r + 1) : 0;
  const int my_tiles = (int)(g_end - g_begin);

  // the first tiles are requested BEFORE waiting for the producer of q2: K|V were written by the prefill
  XpPos lp, cp;
  auto tile_src = [&](const XpPos& p) -> const char* {
    return reinterpret_cast<const char*>(KV) + ((size_t)(p.b * NH + p.head) * ntile_all + p.tile) * XP_STAGE;
  };
  int loaded = 0;
  if (my_tiles > 0) {
    lp = xp_locate(tp, B, g_begin);            // load iterator
    cp = lp;                                   // compute iterator
    if (lane == 0) {
      for (; loaded < XP_NS && loaded < my_tiles; ++loaded) {
        xp_expect(bars + 8 * loaded, XP_STAGE);
        xp_bulk(sbase + loaded * XP_STAGE, tile_src(lp), XP_STAGE, bars + 8 * loaded, pol);
        xp_advance(lp, tp, B);
      }
    }
    loaded = __shfl_sync(0xffffffffu, loaded, 0);
  }
  pdl_wait();            // q2 comes from the preceding layer_chain launch (which also read the old partials)
  // queries without any valid key own no tile: one CTA marks all their partial slots empty
  if (blockIdx.x == gridDim.x - 1) {
    for (int b = warp; b < B; b += XP_WARPS) {
      if (tp[b + 1] != tp[b]) continue;
      for (int idx = lane; idx < W * NH * nslot; idx += 32) {
        part_ml[((size_t)b * W * NH * nslot + idx) * 2] = -INFINITY;
        part_ml[((size_t)b * W * NH * nslot + idx) * 2 + 1] = 0.f;
      }
    }
  }
  if (my_tiles <= 0) return;

  int done = 0;
  while (done < my_tiles) {
    // ---- one (query, head) segment: tiles [cp.tile, seg_end) of it
    const int b = cp.b, hh = cp.head, nt = cp.nt;
    const int seg_n = min(nt - cp.tile, my_tiles - done);
    const int nvalid = ncount[b];
    uint32_t qa[2][2];
    {
      const float* qp = q2 + ((size_t)(b * W + (g4 < W ? g4 : 0))) * H + hh * HD;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float2 lo = *reinterpret_cast<const float2*>(qp + ks * 16 + 2 * t4);
        const float2 hi = *reinterpret_cast<const float2*>(qp + ks * 16 + 8 + 2 * t4);
        qa[ks][0] = g4 < W ? xp_pack(lo.x, lo.y) : 0u;
        qa[ks][1] = g4 < W ? xp_pack(hi.x, hi.y) : 0u;
      }
    }
    float m = -INFINITY, l = 0.f;
    float o[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) { o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f; }
    for (int i = 0; i < seg_n; ++i) {
      const int it = done + i, stage = it % XP_NS;
      const int key0 = (cp.tile + i) * XP_TILE;
      while (!xp_try_wait(bars + 8 * stage, (uint32_t)(it / XP_NS) & 1u)) {}
      const uint32_t kt = sbase + stage * XP_STAGE, vt = kt + XP_TILE_BYTES;
      const int nkey = nvalid - key0;           // keys of this tile that exist (>= 64: all)
      float sc[XP_TILE / 8][2];
      float tmax = -INFINITY;
#pragma unroll
      for (int kb = 0; kb < XP_TILE / 8; ++kb) {
        uint32_t kf[4];
        xp_ldsm4(kf, kt + xp_swz(kb * 8 + (lane & 7), lane >> 3));
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        xp_mma(c, qa[0][0], 0u, qa[0][1], 0u, kf[0], kf[1]);
        xp_mma(c, qa[1][0], 0u, qa[1][1], 0u, kf[2], kf[3]);
        sc[kb][0] = c[0];
        sc[kb][1] = c[1];
      }
      if (nkey < XP_TILE) {                      // only the last tile of a query has keys past its end (warp-uniform)
#pragma unroll
        for (int kb = 0; kb < XP_TILE / 8; ++kb) {
          const int kcol = kb * 8 + 2 * t4;
          if (kcol >= nkey) sc[kb][0] = -INFINITY;
          if (kcol + 1 >= nkey) sc[kb][1] = -INFINITY;
        }
      }
#pragma unroll
      for (int kb = 0; kb < XP_TILE / 8; ++kb) tmax = fmaxf(tmax, fmaxf(sc[kb][0], sc[kb][1]));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
      const float mn = fmaxf(m, tmax);
      const float scale = (m == -INFINITY) ? 0.f : fexp(m - mn);
      m = mn;
      l *= scale;
      const float mnl = -mn * 1.4426950408889634f;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) { o[nb][0] *= scale; o[nb][1] *= scale; }
#pragma unroll
      for (int kk = 0; kk < XP_TILE / 16; ++kk) {
        float p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          // exp(sv - mn) as one FFMA + ex2; a key past the end has sv = -inf -> 0 (mn is finite: every tile of the
          // compacted stream holds at least one valid key)
          p[u] = exp2f(fmaf(sc[2 * kk + (u >> 1)][u & 1], 1.4426950408889634f, mnl));
          l += p[u];
        }
        const uint32_t pa0 = xp_pack(p[0], p[1]), pa2 = xp_pack(p[2], p[3]);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t vf[4];
          const int mi = lane >> 3;
          xp_ldsm4_t(vf, vt + xp_swz(kk * 16 + (mi & 1) * 8 + (lane & 7), c2 * 2 + (mi >> 1)));
          xp_mma(o[c2 * 2], pa0, 0u, pa2, 0u, vf[0], vf[1]);
          xp_mma(o[c2 * 2 + 1], pa0, 0u, pa2, 0u, vf[2], vf[3]);
        }
      }
      __syncwarp();                              // every lane is done with this stage
      if (lane == 0 && loaded < my_tiles) {      // refill it with the tile XP_NS ahead (possibly of the next segment)
        xp_expect(bars + 8 * stage, XP_STAGE);
        xp_bulk(sbase + stage * XP_STAGE, tile_src(lp), XP_STAGE, bars + 8 * stage, pol);
        xp_advance(lp, tp, B);
      }
      if (loaded < my_tiles) ++loaded;
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    // ---- partial slot of this segment: ranges that share (b, hh) are consecutive
    const long long gs = (long long)NH * tp[b] + (long long)hh * nt;          // first tile of (b, hh) in the stream
    const int r0 = range_of(gs), slot = r - r0;
    if (g4 < W) {
      const size_t oidx = (((size_t)(b * W + g4)) * NH + hh) * nslot + slot;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        *reinterpret_cast<float2*>(part_acc + oidx * HD + nb * 8 + 2 * t4) = make_float2(o[nb][0], o[nb][1]);
      if (t4 == 0) { part_ml[oidx * 2] = m; part_ml[oidx * 2 + 1] = l; }
    }
    if (cp.tile == 0) {                          // the first range of (b, hh) marks the unused slots empty
      const int nseg = range_of(gs + nt - 1) - r0 + 1;
      for (int idx = lane; idx < W * (nslot - nseg); idx += 32) {
        const int row = idx / (nslot - nseg), sl = nseg + idx % (nslot - nseg);
        const size_t oidx = (((size_t)(b * W + row)) * NH + hh) * nslot + sl;
        part_ml[oidx * 2] = -INFINITY;
        part_ml[oidx * 2 + 1] = 0.f;
      }
    }
    done += seg_n;
    for (int i = 0; i < seg_n; ++i) xp_advance(cp, tp, B);
  }
}

// ------------------------------------------------------------------------------------------ prefill packing
// like pack_kv_tiles_kernel, but tile (b, j) holds the valid keys cidx[b][64 j .. 64 j + 63] (original
// positions, ascending) and keys >= ncount[b] are zero
struct XpOut { bf16* p[4]; };
__global__ __launch_bounds__(256) void pack_kv_gather_kernel(const bf16* __restrict__ kv, int ldkv, int B, int S,
                                                             const int32_t* __restrict__ cidx,
                                                             const int32_t* __restrict__ ncount, int nl, XpOut out) {
  pdl_wait();
  const int ntile = (S + 63) / 64;
  const size_t per_row = (size_t)nl * 2 * NH * 4;                 // 16-byte chunks per source row
  const size_t total = (size_t)B * ntile * 64 * per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 4), hh = (int)((i / 4) % NH), j = (int)((i / (4 * NH)) % 2), l = (int)((i / (8 * NH)) % nl);
    const size_t row = i / per_row;
    const int key = (int)(row % 64), tile = (int)((row / 64) % ntile), b = (int)(row / ((size_t)64 * ntile));
    const int sc = tile * 64 + key;
    const int nv = ncount[b];
    if (tile * 64 >= nv && tile > 0) continue;                    // tiles past the query's last one are never read
    uint4 v = make_uint4(0, 0, 0, 0);
    if (sc < nv) {
      const int s = cidx[(size_t)b * S + sc];
      v = __ldg(reinterpret_cast<const uint4*>(kv + ((size_t)b * S + s) * ldkv + ((l * 2 + j) * NH + hh) * HD + c * 8));
    }
    char* dst = reinterpret_cast<char*>(out.p[l]) + ((((size_t)(b * NH + hh) * ntile + tile) * 2 + j) * 64 + key) * 64 +
                ((c ^ ((key >> 1) & 3)) << 4);
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

}  // namespace cb

using namespace cb;

extern "C" int case_cross_attn_part_slots(int S) { return ((S + 63) / 64 + XP_MIN_TILES - 1) / XP_MIN_TILES + 2; }

extern "C" int case_cross_attn_part(const float* q2, const void* KV, const int32_t* ncount, const int32_t* tile_prefix,
                                    int B, int W, int S, int nslot, float* part_ml, float* part_acc,
                                    case_stream_t stream) {
  CB_REQUIRE(q2 && KV && ncount && tile_prefix && part_ml && part_acc, "case_cross_attn_part: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0, "case_cross_attn_part: bad sizes");
  CB_REQUIRE(nslot >= case_cross_attn_part_slots(S) && nslot <= CASE_MAX_XSPLIT, "case_cross_attn_part: nslot must be >= case_cross_attn_part_slots(S)");
  CB_REQUIRE((uintptr_t)KV % 16 == 0, "case_cross_attn_part: KV must be 16-byte aligned");
  const size_t smem = (size_t)XP_WARPS * XP_WARP_BYTES + (size_t)(B + 1) * 4;
  CB_REQUIRE(smem <= 227 * 1024, "case_cross_attn_part: too many queries for the shared-memory prefix table");
  ensure_smem<cross_attn_part_kernel>(227 * 1024);
  int dev = 0, nsm = 0;                                  // persistent grid: one CTA per SM of the CURRENT device
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if (nsm <= 0) nsm = 148;
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(cross_attn_part_kernel, nsm, XP_WARPS * 32, smem, st, q2, (const bf16*)KV, ncount, tile_prefix, B, W,
           (S + 63) / 64, nslot, part_ml, part_acc, launch_opts().evict_first);
  return check_launch("case_cross_attn_part");
}

extern "C" int case_pack_kv_tiles_gather(const void* kv, int ldkv, int B, int S, const int32_t* cidx,
                                         const int32_t* ncount, int nl, void* const* out, case_stream_t stream) {
  CB_REQUIRE(kv && out && cidx && ncount && B > 0 && S > 0 && nl >= 1 && nl <= 4, "case_pack_kv_tiles_gather: bad arguments");
  CB_REQUIRE(ldkv % 8 == 0 && (uintptr_t)kv % 16 == 0, "case_pack_kv_tiles_gather: source rows must be 16-byte aligned");
  XpOut o;
  for (int l = 0; l < 4; ++l) o.p[l] = l < nl ? (bf16*)out[l] : nullptr;
  launch_k(pack_kv_gather_kernel, 148 * 8, 256, 0, (cudaStream_t)stream, (const bf16*)kv, ldkv, B, S, cidx, ncount, nl, o);
  return check_launch("case_pack_kv_tiles_gather");
}
