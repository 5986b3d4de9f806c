// Memory-streaming attention kernels of the decode step.
//
//  * cross_attn_partial : multi-head cross-attention of the newest position over a projected memory
//    (TransformerDecoder.py:81).  One CTA per (query, head, key split) streams that head's K and V
//    exactly once and serves all W beam rows of the query from the same shared-memory tile.
//  * additive_attn      : the "BilinearAttention" of the reference (actually additive/Bahdanau,
//    BilinearAttention.py:24-60) fused with its softmax statistics, the prior re-weighting sums of
//    CaSE/Model.py:110-111 and the context reduction.  One CTA per (query, key split).
//
// Both are HBM-bound on their K/V (resp. Uk.mem / mem) streams; algorithmic bytes per launch are
// B * 2 * S * H * sizeof(T).
#include "common.cuh"

namespace cb {

// ------------------------------------------------------------------------------------------ cross attention
constexpr int XT = 128;   // keys per tile == threads per CTA

template <typename T, int WMAX>
__global__ __launch_bounds__(XT) void cross_attn_partial_kernel(
    const float* __restrict__ q2, const T* __restrict__ Kmem, const T* __restrict__ Vmem,
    const uint8_t* __restrict__ mask, int W, int S, int nsplit, float* __restrict__ part_ml,
    float* __restrict__ part_acc) {
  __shared__ float qs[WMAX][HD];
  __shared__ float Ks[XT][HD + 1];
  __shared__ __align__(16) float Vs[XT][HD];
  __shared__ float ps[WMAX][XT];
  __shared__ float wred[2][XT / 32][WMAX];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, hh = blockIdx.y, sp = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = split_chunk(S, nsplit, XT);
  const int s_begin = sp * chunk, s_end = min(S, s_begin + chunk);

  for (int i = tid; i < W * HD; i += XT) {
    const int w = i / HD, d = i % HD;
    qs[w][d] = q2[((size_t)(b * W + w)) * H + hh * HD + d];
  }
  float m[WMAX], l[WMAX], acc[(WMAX + 3) / 4];
#pragma unroll
  for (int w = 0; w < WMAX; ++w) { m[w] = -INFINITY; l[w] = 0.f; }
#pragma unroll
  for (int i = 0; i < (WMAX + 3) / 4; ++i) acc[i] = 0.f;
  const T* Kb = Kmem + ((size_t)(b * NH + hh)) * S * HD;
  const T* Vb = Vmem + ((size_t)(b * NH + hh)) * S * HD;
  const uint8_t* mb = mask + (size_t)b * S;
  __syncthreads();

  for (int s0 = s_begin; s0 < s_end; s0 += XT) {
    const int cnt = min(XT, s_end - s0);
    // tile = cnt*HD contiguous elements of K and of V
    for (int e = tid * 8; e < cnt * HD; e += XT * 8) {
      float kv[8], vv[8];
      ld8(Kb + (size_t)s0 * HD + e, kv);
      ld8(Vb + (size_t)s0 * HD + e, vv);
      const int s = e / HD, d = e % HD;
#pragma unroll
      for (int u = 0; u < 8; ++u) { Ks[s][d + u] = kv[u]; Vs[s][d + u] = vv[u]; }
    }
    __syncthreads();
    // phase 1: thread = key
    float sc[WMAX];
    const bool valid = tid < cnt && mb[s0 + tid] != 0;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) sc[w] = 0.f;
    if (valid) {
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        const float kd = Ks[tid][d];
#pragma unroll
        for (int w = 0; w < WMAX; ++w) sc[w] = fmaf(qs[w][d], kd, sc[w]);   // rows >= W read zeros/garbage, unused
      }
    }
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      sc[w] = valid ? sc[w] : -INFINITY;
      const float mx = warp_max(sc[w]);
      if (lane == 0) wred[0][warp][w] = mx;
    }
    __syncthreads();
    float scale[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      float tm = fmaxf(fmaxf(wred[0][0][w], wred[0][1][w]), fmaxf(wred[0][2][w], wred[0][3][w]));
      const float mn = fmaxf(m[w], tm);
      scale[w] = (m[w] == -INFINITY) ? 0.f : fexp(m[w] - mn);
      m[w] = mn;
      const float p = (sc[w] == -INFINITY) ? 0.f : fexp(sc[w] - mn);
      ps[w][tid] = p;
      const float su = warp_sum(p);
      if (lane == 0) wred[1][warp][w] = su;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < WMAX; ++w)
      l[w] = fmaf(l[w], scale[w], (wred[1][0][w] + wred[1][1][w]) + (wred[1][2][w] + wred[1][3][w]));
    // phase 2: thread = (w group, d)
#pragma unroll
    for (int i = 0; i < (WMAX + 3) / 4; ++i) {
      const int w = warp + 4 * i;
      if (w < WMAX && w < W) {
        float a = acc[i] * scale[w];
        for (int s = 0; s < cnt; ++s) a = fmaf(ps[w][s], Vs[s][lane], a);
        acc[i] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < (WMAX + 3) / 4; ++i) {
    const int w = warp + 4 * i;
    if (w < WMAX && w < W) {
      const size_t o = (((size_t)(b * W + w)) * NH + hh) * nsplit + sp;
      part_acc[o * HD + lane] = acc[i];
    }
  }
  if (tid < W) {
    float mm = -INFINITY, ll = 0.f;
#pragma unroll
    for (int w = 0; w < WMAX; ++w)
      if (w == tid) { mm = m[w]; ll = l[w]; }
    const size_t o = (((size_t)(b * W + tid)) * NH + hh) * nsplit + sp;
    part_ml[o * 2] = mm;
    part_ml[o * 2 + 1] = ll;
  }
}

// ------------------------------------------------------------------------------------------ additive attention
constexpr int AT = 256;    // threads
constexpr int ATB = 128;   // keys per tile in pass B

template <typename T> struct ATile { static constexpr int TS = 32; static constexpr int PAD = 4; };
template <> struct ATile<bf16> { static constexpr int TS = 64; static constexpr int PAD = 8; };

template <typename T, int WMAX, bool FAST>
__global__ __launch_bounds__(AT) void additive_attn_kernel(
    const float* __restrict__ qa, const T* __restrict__ U, const T* __restrict__ Mv,
    const float* __restrict__ vvec, const uint8_t* __restrict__ mask, const float* __restrict__ prior,
    const int32_t* __restrict__ tok, int tok_ld, int t, int W, int S, int DV, int nsplit,
    float* attn_un, float* __restrict__ stats, float* __restrict__ ctx_part) {
  constexpr int TS = ATile<T>::TS, LD = H + ATile<T>::PAD, NG = AT / TS, HG = H / NG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qas = reinterpret_cast<float*>(smem_raw);             // [WMAX][H]
  float* vs = qas + WMAX * H;                                  // [H]
  float* er = vs + H;                                          // [NG][WMAX][TS]  (pass A) / scratch (pass B)
  float* mrow = er + NG * WMAX * TS;                           // [8] block max per row
  float* wsum = mrow + 8;                                      // [8 warps][8]
  T* Us = reinterpret_cast<T*>(wsum + 128);                    // [TS][LD]
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, sp = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = split_chunk(S, nsplit, ATB);
  const int s_begin = sp * chunk, s_end = min(S, s_begin + chunk);
  const int r0 = b * W;

  for (int i = tid; i < W * H; i += AT) qas[i] = qa[(size_t)r0 * H + i];
  for (int i = tid; i < H; i += AT) vs[i] = vvec[i];
  __syncthreads();

  // ---------------- pass A: raw masked scores -> attn_un
  const T* Ub = U + (size_t)b * S * H;
  const uint8_t* mb = mask + (size_t)b * S;
  for (int s0 = s_begin; s0 < s_end; s0 += TS) {
    const int cnt = min(TS, s_end - s0);
    for (int e = tid * 8; e < cnt * H; e += AT * 8) {
      const int s = e / H, c = e % H;
      // raw copy of 8 elements (16 B bf16 / 32 B fp32)
      if (sizeof(T) == 2) {
        *reinterpret_cast<uint4*>(Us + s * LD + c) = __ldg(reinterpret_cast<const uint4*>(Ub + (size_t)s0 * H + e));
      } else {
        const float4* src = reinterpret_cast<const float4*>(Ub + (size_t)s0 * H + e);
        float4* dst = reinterpret_cast<float4*>(Us + s * LD + c);
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1);
      }
    }
    __syncthreads();
    {
      const int s = tid % TS, g = tid / TS;
      float e[WMAX];
#pragma unroll
      for (int w = 0; w < WMAX; ++w) e[w] = 0.f;
      if (s < cnt) {
        const T* up = Us + s * LD + g * HG;
        for (int k = 0; k < HG; k += 8) {
          float u[8];
          ld8c(up + k, u);
          const int hh = g * HG + k;
#pragma unroll
          for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = qas[w * H + hh + i] + u[i];
                e[w] = fmaf(vs[hh + i], FAST ? tanh_fast(x) : tanh_acc(x), e[w]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int w = 0; w < WMAX; ++w) er[(g * WMAX + w) * TS + s] = e[w];
    }
    __syncthreads();
    for (int i = tid; i < W * cnt; i += AT) {
      const int w = i / cnt, s = i % cnt;
      float e = 0.f;
#pragma unroll
      for (int g = 0; g < NG; ++g) e += er[(g * WMAX + w) * TS + s];
      const bool rv = tok ? tok[(size_t)(r0 + w) * tok_ld + t] != 0 : true;
      const bool ok = rv && mb[s0 + s] != 0;
      attn_un[(size_t)(r0 + w) * S + s0 + s] = ok ? e : -INFINITY;
    }
    __syncthreads();
  }

  // ---------------- block max per row over this split
  for (int w = 0; w < W; ++w) {
    float mx = -INFINITY;
    for (int s = s_begin + tid; s < s_end; s += AT) mx = fmaxf(mx, attn_un[(size_t)(r0 + w) * S + s]);
    mx = warp_max(mx);
    if (lane == 0) wsum[warp * 8 + w] = mx;
  }
  __syncthreads();
  if (tid < W) {
    float mx = -INFINITY;
    for (int i = 0; i < AT / 32; ++i) mx = fmaxf(mx, wsum[i * 8 + tid]);
    mrow[tid] = mx;
  }
  __syncthreads();

  // ---------------- pass B: p = exp(e - m), sums, context partial
  float* pt = er;                                   // [WMAX][ATB]
  const int ncol2 = DV / 2;                         // column pairs
  const int npc = ncol2 / 128;                      // pairs per thread (1 for DV=256, 2 for DV=512)
  const int hp = tid & 127, kp = tid >> 7;
  float acc[WMAX][2][2];
#pragma unroll
  for (int w = 0; w < WMAX; ++w)
#pragma unroll
    for (int c = 0; c < 2; ++c) { acc[w][c][0] = 0.f; acc[w][c][1] = 0.f; }
  float lsum = 0.f, lwsum = 0.f;                    // this thread's (w = tid / ATB... ) partial sums
  const T* Mb = Mv + (size_t)b * S * DV;
  const float* pb = prior ? prior + (size_t)b * S : nullptr;
  for (int s0 = s_begin; s0 < s_end; s0 += ATB) {
    const int cnt = min(ATB, s_end - s0);
    for (int i = tid; i < W * ATB; i += AT) {
      const int w = i / ATB, s = i % ATB;
      float p = 0.f;
      if (s < cnt) {
        const float e = attn_un[(size_t)(r0 + w) * S + s0 + s];
        p = (e == -INFINITY) ? 0.f : fexp(e - mrow[w]);     // attn_un keeps the raw score for the scatter
      }
      pt[w * ATB + s] = p;
    }
    __syncthreads();
    // row sums of this tile: warp w' handles row w' (W <= 8 warps)
    if (warp < W) {
      float a = 0.f, aw = 0.f;
      for (int s = lane; s < cnt; s += 32) {
        const float p = pt[warp * ATB + s];
        a += p;
        aw = fmaf(pb ? pb[s0 + s] : 1.f, p, aw);
      }
      lsum += a; lwsum += aw;
    }
    for (int s = kp; s < cnt; s += 2) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < npc) {
          float mv[2];
          ld2(Mb + (size_t)(s0 + s) * DV + (c * 128 + hp) * 2, mv);
#pragma unroll
          for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
              const float p = pt[w * ATB + s];
              acc[w][c][0] = fmaf(p, mv[0], acc[w][c][0]);
              acc[w][c][1] = fmaf(p, mv[1], acc[w][c][1]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (warp < W) {
    lsum = warp_sum(lsum);
    lwsum = warp_sum(lwsum);
    if (lane == 0) {
      float* st = stats + ((size_t)(r0 + warp) * nsplit + sp) * 4;
      st[0] = mrow[warp]; st[1] = lsum; st[2] = lwsum; st[3] = 0.f;
    }
  }
  // combine the two key-parity halves through shared memory (reuses the U tile region)
  float* cred = reinterpret_cast<float*>(Us);      // [WMAX][DV] needs WMAX*DV*4 <= TS*LD*sizeof(T)
  for (int w = 0; w < W; ++w) {
    if (kp == 1) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (c < npc) {
          cred[(c * 128 + hp) * 2] = acc[w][c][0];
          cred[(c * 128 + hp) * 2 + 1] = acc[w][c][1];
        }
    }
    __syncthreads();
    if (kp == 0) {
      float* dst = ctx_part + ((size_t)(r0 + w) * nsplit + sp) * DV;
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (c < npc) {
          const int col = (c * 128 + hp) * 2;
          dst[col] = acc[w][c][0] + cred[col];
          dst[col + 1] = acc[w][c][1] + cred[col + 1];
        }
    }
    __syncthreads();
  }
}

template <typename T, int WMAX, bool FAST>
static int launch_additive(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                           const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int DV,
                           int nsplit, float* attn_un, float* stats, float* ctx_part, cudaStream_t st) {
  constexpr int TS = ATile<T>::TS, LD = H + ATile<T>::PAD, NG = AT / TS;
  static_assert(NG * TS >= ATB, "pass-B tile must fit in the pass-A reduction buffer");
  const size_t fl = (size_t)WMAX * H + H + (size_t)NG * WMAX * TS + 8 + 128;
  const size_t smem = fl * sizeof(float) + (size_t)TS * LD * sizeof(T);
  auto kern = additive_attn_kernel<T, WMAX, FAST>;
  ensure_smem<additive_attn_kernel<T, WMAX, FAST>>(96 * 1024);
  launch_k(kern, dim3(B, nsplit), AT, smem, st, qa, (const T*)U, (const T*)Mv, v, mask, prior, tok, tok_ld, t, W, S, DV,
                                          nsplit, attn_un, stats, ctx_part);
  return check_launch("case_additive_attn");
}

template <typename T, bool FAST>
static int dispatch_additive_w(int W, const float* qa, const void* U, const void* Mv, const float* v,
                               const uint8_t* mask, const float* prior, const int32_t* tok, int tok_ld, int t, int B,
                               int S, int DV, int nsplit, float* attn_un, float* stats, float* ctx_part,
                               cudaStream_t st) {
  if (W <= 1) return launch_additive<T, 1, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats, ctx_part, st);
  if (W <= 2) return launch_additive<T, 2, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats, ctx_part, st);
  if (W <= 4) return launch_additive<T, 4, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats, ctx_part, st);
  return launch_additive<T, 8, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats, ctx_part, st);
}

template <typename T, int WMAX>
static int launch_cross(const float* q2, const void* K, const void* V, const uint8_t* mask, int B, int W, int S,
                        int nsplit, float* part_ml, float* part_acc, cudaStream_t st) {
  launch_k(cross_attn_partial_kernel<T, WMAX>, dim3(B, NH, nsplit), XT, 0, st, q2, (const T*)K, (const T*)V, mask, W, S,
                                                                          nsplit, part_ml, part_acc);
  return check_launch("case_cross_attn_partial");
}

template <typename T>
static int dispatch_cross_w(const float* q2, const void* K, const void* V, const uint8_t* mask, int B, int W, int S,
                            int nsplit, float* part_ml, float* part_acc, cudaStream_t st) {
  if (W <= 1) return launch_cross<T, 1>(q2, K, V, mask, B, W, S, nsplit, part_ml, part_acc, st);
  if (W <= 2) return launch_cross<T, 2>(q2, K, V, mask, B, W, S, nsplit, part_ml, part_acc, st);
  if (W <= 4) return launch_cross<T, 4>(q2, K, V, mask, B, W, S, nsplit, part_ml, part_acc, st);
  return launch_cross<T, 8>(q2, K, V, mask, B, W, S, nsplit, part_ml, part_acc, st);
}

}  // namespace cb

using namespace cb;

extern "C" int case_cross_attn_partial(const float* q2, const void* Kmem, const void* Vmem, const uint8_t* mask,
                                       int B, int W, int S, int nsplit, float* part_ml, float* part_acc, int dtype,
                                       case_stream_t stream) {
  CB_REQUIRE(q2 && Kmem && Vmem && mask && part_ml && part_acc, "case_cross_attn_partial: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0, "case_cross_attn_partial: bad sizes");
  CB_REQUIRE(nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "case_cross_attn_partial: nsplit out of range");
  if (dtype == CASE_BF16)
    return dispatch_cross_w<bf16>(q2, Kmem, Vmem, mask, B, W, S, nsplit, part_ml, part_acc, (cudaStream_t)stream);
  return dispatch_cross_w<float>(q2, Kmem, Vmem, mask, B, W, S, nsplit, part_ml, part_acc, (cudaStream_t)stream);
}

int case_additive_attn_bf16(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                            const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int DV,
                            int nsplit, float* scores, float* stats, float* ctx_part, int fast_tanh, const int32_t* cidx,
                            const int32_t* ncount, const int32_t* qorder, cudaStream_t st);

extern "C" int case_additive_attn(const float* qa, const void* U, const void* Mv, const float* v,
                                  const uint8_t* mask, const float* prior, const int32_t* tok, int tok_ld, int t,
                                  int B, int W, int S, int DV, int nsplit, float* attn_un, float* stats,
                                  float* ctx_part, int fast_tanh, int dtype, case_stream_t stream) {
  CB_REQUIRE(qa && U && Mv && v && mask && attn_un && stats && ctx_part, "case_additive_attn: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0, "case_additive_attn: bad sizes");
  CB_REQUIRE(DV == 256 || DV == 512, "case_additive_attn: DV must be 256 or 512");
  CB_REQUIRE(nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "case_additive_attn: nsplit out of range");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CASE_BF16)     // fused single-pass kernel of additive_v2.cu (the two-pass kernel here is the fp32 form)
    return case_additive_attn_bf16(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats,
                                   ctx_part, fast_tanh, nullptr, nullptr, nullptr, st);
  if (fast_tanh) return dispatch_additive_w<float, true>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, DV, nsplit, attn_un, stats, ctx_part, st);
  return dispatch_additive_w<float, false>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, DV, nsplit, attn_un, stats, ctx_part, st);
}

// =============================================================================== tensor-core cross attention
// bf16 K/V only.  FlashAttention-2 style decode step on mma.sync.m16n8k16 tiles: the W (<= 8) beam rows
// of one query are the M rows 0..7 of the tile (rows 8..15 are zero padding), so one pass over a
// head's K/V serves every beam.  K and V of a head live interleaved in ONE array, tile by tile
// ([64 keys x 32] K then [64 x 32] V, 16-byte chunks XOR-swizzled by ((key >> 1) & 3) so both the
// row-wise and the transposed ldmatrix reads are bank-conflict free), which lets each warp stream its
// tiles with one 8 KB bulk async copy per stage into a private 3-stage ring (mbarrier per stage; a
// warp keeps one bulk copy in flight, eight warps per SM saturate HBM - profiles/micro/
// hbm_stream_bench.cu).  A CTA is one (query, key split); its eight warps are the eight heads, each
// walking its head's tiles on its own - no block-level synchronisation at all - and writing its own
// partial per (row, head, split) for case_layer_back.  HBM traffic = K and V once.
namespace cb {

constexpr int XM_WARPS = 8;                        // one warp per head
constexpr int XM_TILE = 64;                        // keys per tile
constexpr int XM_TILE_BYTES = XM_TILE * HD * 2;    // 4 KB of K, then 4 KB of V
constexpr int XM_STAGE = 2 * XM_TILE_BYTES;        // 8 KB per stage
constexpr int XM_NS = 3;                           // ring depth per warp
constexpr int XM_WARP_BYTES = XM_NS * XM_STAGE + 64;
constexpr int XM_SMEM = XM_WARPS * XM_WARP_BYTES;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void xm_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xm_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool xm_try_wait(uint32_t bar, uint32_t parity) {
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
// byte offset of (key row, 16-byte chunk) inside a [64][32] bf16 tile (the swizzle is already in memory)
__device__ __forceinline__ uint32_t xm_swz(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

__global__ __launch_bounds__(XM_WARPS * 32) void cross_attn_mma_kernel(
    const float* __restrict__ q2, const bf16* __restrict__ KV, const uint8_t* __restrict__ mask, int W, int S,
    int nsplit, float* __restrict__ part_ml, float* __restrict__ part_acc) {
  extern __shared__ __align__(128) unsigned char xm_smem[];   // [warp][XM_NS stages of K|V][barriers]
  const int b = blockIdx.x, sp = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hh = warp;
  const int ntile_all = (S + XM_TILE - 1) / XM_TILE;
  const int chunk = split_chunk(S, nsplit, XM_TILE);
  const int s_begin = sp * chunk, s_end = min(S, s_begin + chunk);
  const int tile0 = s_begin / XM_TILE;                                  // this split's first tile
  const int tile_end = (s_end + XM_TILE - 1) / XM_TILE;
  const int my_tiles = tile0 < tile_end ? tile_end - tile0 : 0;
  const char* kvb = reinterpret_cast<const char*>(KV) + ((size_t)(b * NH + hh)) * ntile_all * XM_STAGE;
  const uint8_t* mb = mask + (size_t)b * S;
  unsigned char* wsm = xm_smem + (size_t)warp * XM_WARP_BYTES;
  const uint32_t sbase = smem_u32(wsm), bars = sbase + XM_NS * XM_STAGE;

  if (lane == 0) {
    for (int i = 0; i < XM_NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_trigger();
  pdl_wait();            // K/V come from the prefill, q2 from the preceding layer_front
  if (lane == 0)
    for (int p = 0; p < XM_NS && p < my_tiles; ++p) {
      xm_expect(bars + 8 * p, XM_STAGE);
      xm_bulk(sbase + p * XM_STAGE, kvb + (size_t)(tile0 + p) * XM_STAGE, XM_STAGE, bars + 8 * p);
    }

  // Q fragments (A operand), rows >= W are zero
  uint32_t qa[2][2];
  {
    const float* qp = q2 + ((size_t)(b * W + (g < W ? g : 0))) * H + hh * HD;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const float2 lo = *reinterpret_cast<const float2*>(qp + ks * 16 + 2 * t);
      const float2 hi = *reinterpret_cast<const float2*>(qp + ks * 16 + 8 + 2 * t);
      qa[ks][0] = g < W ? pack_bf16(lo.x, lo.y) : 0u;
      qa[ks][1] = g < W ? pack_bf16(hi.x, hi.y) : 0u;
    }
  }
  float m = -INFINITY, l = 0.f;
  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) { o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f; }

  // key-validity bits of a tile: lane owns keys 2*lane, 2*lane+1 -> two ballots (even keys, odd keys).
  // The bytes of tile i+1 are requested before waiting for tile i, so their latency is never exposed.
  auto key_ok = [&](int s) { return s >= s_begin && s < s_end && mb[s] != 0; };
  uint32_t be = 0, bo = 0;
  if (my_tiles > 0) {
    const int s = tile0 * XM_TILE + 2 * lane;
    be = __ballot_sync(0xffffffffu, key_ok(s));
    bo = __ballot_sync(0xffffffffu, key_ok(s + 1));
  }
  for (int i = 0; i < my_tiles; ++i) {
    const int stage = i % XM_NS;
    const int tile_s0 = (tile0 + i) * XM_TILE;
    const int ns = tile_s0 + XM_TILE + 2 * lane;            // this lane's keys in the next tile
    const bool ne = (i + 1 < my_tiles) && key_ok(ns), no = (i + 1 < my_tiles) && key_ok(ns + 1);
    while (!xm_try_wait(bars + 8 * stage, (uint32_t)(i / XM_NS) & 1u)) {}
    const uint32_t kt = sbase + stage * XM_STAGE, vt = kt + XM_TILE_BYTES;
    // ---- S = Q K^T for 8 key blocks of 8
    float sc[XM_TILE / 8][2];
    float tmax = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < XM_TILE / 8; ++kb) {
      uint32_t kf[4];
      ldsm_x4(kf, kt + xm_swz(kb * 8 + (lane & 7), lane >> 3));
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(c, qa[0][0], 0u, qa[0][1], 0u, kf[0], kf[1]);
      mma_bf16_16816(c, qa[1][0], 0u, qa[1][1], 0u, kf[2], kf[3]);
      sc[kb][0] = ((be >> (kb * 4 + t)) & 1u) ? c[0] : -INFINITY;
      sc[kb][1] = ((bo >> (kb * 4 + t)) & 1u) ? c[1] : -INFINITY;
      tmax = fmaxf(tmax, fmaxf(sc[kb][0], sc[kb][1]));
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float mn = fmaxf(m, tmax);
    const float scale = (m == -INFINITY) ? 0.f : fexp(m - mn);
    m = mn;
    l *= scale;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) { o[nb][0] *= scale; o[nb][1] *= scale; }
    // ---- P (bf16) and O += P V, 16 keys at a time
#pragma unroll
    for (int kk = 0; kk < XM_TILE / 16; ++kk) {
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sv = sc[2 * kk + (u >> 1)][u & 1];
        p[u] = (sv == -INFINITY) ? 0.f : fexp(sv - mn);
        l += p[u];
      }
      const uint32_t pa0 = pack_bf16(p[0], p[1]), pa2 = pack_bf16(p[2], p[3]);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t vf[4];
        const int mi = lane >> 3;
        ldsm_x4_t(vf, vt + xm_swz(kk * 16 + (mi & 1) * 8 + (lane & 7), c2 * 2 + (mi >> 1)));
        mma_bf16_16816(o[c2 * 2], pa0, 0u, pa2, 0u, vf[0], vf[1]);
        mma_bf16_16816(o[c2 * 2 + 1], pa0, 0u, pa2, 0u, vf[2], vf[3]);
      }
    }
    be = __ballot_sync(0xffffffffu, ne);                  // also: every lane is done with this stage
    bo = __ballot_sync(0xffffffffu, no);
    if (lane == 0 && i + XM_NS < my_tiles) {
      xm_expect(bars + 8 * stage, XM_STAGE);
      xm_bulk(sbase + stage * XM_STAGE, kvb + (size_t)(tile0 + i + XM_NS) * XM_STAGE, XM_STAGE, bars + 8 * stage);
    }
  }
  l += __shfl_xor_sync(0xffffffffu, l, 1);
  l += __shfl_xor_sync(0xffffffffu, l, 2);
  if (g < W) {
    const size_t oidx = (((size_t)(b * W + g)) * NH + hh) * nsplit + sp;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
      *reinterpret_cast<float2*>(part_acc + oidx * HD + nb * 8 + 2 * t) = make_float2(o[nb][0], o[nb][1]);
    if (t == 0) { part_ml[oidx * 2] = m; part_ml[oidx * 2 + 1] = l; }
  }
}

}  // namespace cb

extern "C" int case_cross_attn_partial_tc(const float* q2, const void* KV, const uint8_t* mask, int B, int W, int S,
                                          int nsplit, float* part_ml, float* part_acc, case_stream_t stream) {
  CB_REQUIRE(q2 && KV && mask && part_ml && part_acc, "case_cross_attn_partial_tc: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0, "case_cross_attn_partial_tc: bad sizes");
  CB_REQUIRE(nsplit >= 1 && nsplit <= CASE_MAX_XSPLIT, "case_cross_attn_partial_tc: nsplit out of range");
  CB_REQUIRE((uintptr_t)KV % 16 == 0, "case_cross_attn_partial_tc: KV must be 16-byte aligned");
  ensure_smem<cb::cross_attn_mma_kernel>(cb::XM_SMEM);
  launch_k(cb::cross_attn_mma_kernel, dim3(B, nsplit), cb::XM_WARPS * 32, cb::XM_SMEM, (cudaStream_t)stream,
           q2, (const cb::bf16*)KV, mask, W, S, nsplit, part_ml, part_acc);
  return cb::check_launch("case_cross_attn_partial_tc");
}
