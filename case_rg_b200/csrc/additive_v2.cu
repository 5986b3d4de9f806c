// Additive ("bilinear") attention, fused single pass for bf16 storage (BilinearAttention.py:24-60 +
// CaSE/Model.py:110-111 sums).  attention.cu keeps the two-pass fp32 form.
//
// One CTA per (query, key split) walks its keys in tiles of 32 with a cp.async double buffer holding
// the tile of Uk.mem rows AND the tile of value rows, so the HBM stream of the next tile hides under
// the MUFU-bound tanh work of the current one.  Per tile:
//   scores   warp g <-> 32 hidden units, lane <-> key:  partial e[w][key] over the warp's units
//   softmax  warp w <-> beam row: sum the 8 partials, mask, write the raw score, online (max, sum,
//            prior-weighted sum) update, p = exp(e - running max)
//   context  warp kp <-> 4 keys of the tile, lane <-> 8 value columns: acc[w] = acc[w]*scale + p*Mv
// Tiles whose 32 keys are all padding are skipped outright (no loads, no tanh).
// Outputs: raw masked scores e [R][S], per split (max, sum, prior-weighted sum) and the context
// partial relative to that max.  The tanh count (R*S*H per launch, one MUFU op each at 16/clk/SM)
// is the floor of this kernel; its HBM stream is B*2*S*H*2 bytes.
#include <cuda_fp16.h>

#include "common.cuh"

namespace cb {

int g_additive_impl = 3;      // 3 = warp-autonomous kernel, 2 = block-synchronous tiles (A/B)
static const int32_t* g_add_cidx = nullptr;     // compaction tables of the NEXT launch (case_additive_attn_compact)
static const int32_t* g_add_ncount = nullptr;
static const int32_t* g_add_qorder = nullptr;

constexpr int A2T = 256;      // threads
constexpr int A2K = 32;       // keys per tile
constexpr int A2_SPLIT = 32;  // key splits are whole tiles of 32 (the fp32 kernels of attention.cu use AATTN_TILE)
constexpr int A2ULD = H + 8;  // padded bf16 row of the U tile

__device__ __forceinline__ void a2_cp16(uint32_t dst, const void* src, int nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void a2_cp16_hint(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void a2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void a2_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int WMAX, int DV, bool FAST>
__global__ __launch_bounds__(A2T) void additive_attn_v2_kernel(
    const float* __restrict__ qa, const bf16* __restrict__ U, const bf16* __restrict__ Mv,
    const float* __restrict__ vvec, const uint8_t* __restrict__ mask, const float* __restrict__ prior,
    const int32_t* __restrict__ tok, int tok_ld, int t, int W, int S, int nsplit, float* __restrict__ scores,
    float* __restrict__ stats, float* __restrict__ ctx_part) {
  constexpr int CG = DV / 256;                 // 8-column groups per lane in the context phase
  constexpr int U_STAGE = A2K * A2ULD * 2;     // bytes
  constexpr int M_STAGE = A2K * DV * 2;
  extern __shared__ __align__(128) unsigned char sm[];
  bf16* Us = reinterpret_cast<bf16*>(sm);                                   // [2][A2K][A2ULD]
  bf16* Ms = reinterpret_cast<bf16*>(sm + 2 * U_STAGE);                     // [2][A2K][DV]
  float* qas = reinterpret_cast<float*>(sm + 2 * U_STAGE + 2 * M_STAGE);    // [WMAX][H]
  float* vs = qas + WMAX * H;                                               // [H]
  float* er = vs + H;                                                       // [8 groups][WMAX][A2K]
  float* ps = er + 8 * WMAX * A2K;                                          // [WMAX][A2K]
  float* sscale = ps + WMAX * A2K;                                          // [8]
  uint32_t* tvalid = reinterpret_cast<uint32_t*>(sscale + 8);               // [ntiles] any-valid flags
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, sp = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = split_chunk(S, nsplit, A2_SPLIT);
  const int s_begin = sp * chunk, s_end = min(S, s_begin + chunk);
  const int ntiles = s_end > s_begin ? (s_end - s_begin + A2K - 1) / A2K : 0;
  const int r0 = b * W;
  const uint8_t* mb = mask + (size_t)b * S;
  const bf16* Ub = U + (size_t)b * S * H;
  const bf16* Mb = Mv + (size_t)b * S * DV;
  const float* pb = prior ? prior + (size_t)b * S : nullptr;

  for (int i = tid; i < W * H; i += A2T) qas[i] = qa[(size_t)r0 * H + i];
  for (int i = tid; i < H; i += A2T) vs[i] = vvec[i];
  for (int i = warp; i < ntiles; i += A2T / 32) {
    const int s = s_begin + i * A2K + lane;
    const unsigned any = __ballot_sync(0xffffffffu, s < s_end && mb[s] != 0);
    if (lane == 0) tvalid[i] = any;
  }
  __syncthreads();

  auto load_tile = [&](int ti, int stage) {
    const int s0 = s_begin + ti * A2K;
    const uint32_t ud = smem_u32(Us) + stage * U_STAGE, md = smem_u32(Ms) + stage * M_STAGE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {                         // U: 32 rows x 32 chunks of 16 B
      const int ci = tid + A2T * i, row = ci >> 5, ch = ci & 31, s = s0 + row;
      a2_cp16(ud + row * (A2ULD * 2) + ch * 16, Ub + (size_t)min(s, S - 1) * H + ch * 8, s < s_end ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < 4 * CG; ++i) {                    // Mv: 32 rows x (DV/8) chunks
      const int ci = tid + A2T * i, row = ci / (DV / 8), ch = ci % (DV / 8), s = s0 + row;
      a2_cp16(md + row * (DV * 2) + ch * 16, Mb + (size_t)min(s, S - 1) * DV + ch * 8, s < s_end ? 16 : 0);
    }
  };
  int next = 0;                                           // next tile to load (skipping all-padding tiles)
  while (next < ntiles && tvalid[next] == 0) ++next;
  if (next < ntiles) load_tile(next, 0);
  a2_commit();

  // per-row running statistics live in the lanes of warp w (uniform across the warp)
  float m_run = -INFINITY, l_run = 0.f, lw_run = 0.f;
  bool rowvalid = true;
  if (warp < W && tok) rowvalid = tok[(size_t)(r0 + warp) * tok_ld + t] != 0;
  float acc[WMAX][CG][8];
#pragma unroll
  for (int w = 0; w < WMAX; ++w)
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[w][c][i] = 0.f;

  int stage = 0;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int s0 = s_begin + ti * A2K;
    if (tvalid[ti] == 0) {                               // all padding: scores are -inf, nothing else to do
      if (warp < W && s0 + lane < s_end) scores[(size_t)(r0 + warp) * S + s0 + lane] = -INFINITY;
      continue;
    }
    int nn = ti + 1;
    while (nn < ntiles && tvalid[nn] == 0) ++nn;
    if (nn < ntiles) load_tile(nn, stage ^ 1);
    a2_commit();
    a2_wait<1>();
    __syncthreads();                                      // tile ti (U and Mv) visible to every warp

    // ---- scores: warp = 32 hidden units, lane = key
    {
      const bf16* up = Us + (size_t)stage * (A2K * A2ULD) + lane * A2ULD + warp * 32;
      float e[WMAX];
#pragma unroll
      for (int w = 0; w < WMAX; ++w) e[w] = 0.f;
      if ((tvalid[ti] >> lane) & 1u) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          float u[8];
          ld8c(up + k, u);
          const float4 v0 = *reinterpret_cast<const float4*>(vs + warp * 32 + k);
          const float4 v1 = *reinterpret_cast<const float4*>(vs + warp * 32 + k + 4);
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
              const float4 q0 = *reinterpret_cast<const float4*>(qas + w * H + warp * 32 + k);
              const float4 q1 = *reinterpret_cast<const float4*>(qas + w * H + warp * 32 + k + 4);
              const float qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = qq[i] + u[i];
                e[w] = fmaf(vv[i], FAST ? tanh_fast(x) : tanh_acc(x), e[w]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int w = 0; w < WMAX; ++w) er[(warp * WMAX + w) * A2K + lane] = e[w];
    }
    __syncthreads();
    // ---- softmax bookkeeping: warp w = beam row w, lane = key
    if (warp < W) {
      const int s = s0 + lane;
      float e = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) e += er[(g * WMAX + warp) * A2K + lane];
      const bool ok = rowvalid && ((tvalid[ti] >> lane) & 1u);
      e = ok ? e : -INFINITY;
      if (s < s_end) scores[(size_t)(r0 + warp) * S + s] = e;
      const float tmax = warp_max(e);
      const float mn = fmaxf(m_run, tmax);
      const float sc = (m_run == -INFINITY) ? 0.f : fexp(m_run - mn);
      const float p = (e == -INFINITY) ? 0.f : fexp(e - mn);
      const float pw = (pb && s < s_end) ? pb[s] * p : p;
      l_run = fmaf(l_run, sc, warp_sum(p));
      lw_run = fmaf(lw_run, sc, warp_sum(pw));
      m_run = mn;
      ps[warp * A2K + lane] = p;
      if (lane == 0) sscale[warp] = sc;
    }
    __syncthreads();
    // ---- context: warp = 4 keys of the tile, lane = 8 (x CG) value columns
    {
      const bf16* mp = Ms + (size_t)stage * (A2K * DV);
#pragma unroll
      for (int w = 0; w < WMAX; ++w) {
        if (w < W) {
          const float sc = sscale[w];
#pragma unroll
          for (int c = 0; c < CG; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[w][c][i] *= sc;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int key = warp * 4 + kk;
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          float mv[8];
          ld8c(mp + key * DV + c * 256 + lane * 8, mv);
#pragma unroll
          for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
              const float p = ps[w * A2K + key];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[w][c][i] = fmaf(p, mv[i], acc[w][c][i]);
            }
          }
        }
      }
    }
    __syncthreads();                                      // stage and er / ps may be overwritten next iteration
    stage ^= 1;
  }
  a2_wait<0>();
  __syncthreads();
  // ---- reduce the 8 key-phase warps: cred[warp][w][DV] through the (now idle) tile buffers
  float* cred = reinterpret_cast<float*>(sm);             // needs 8 * W * DV * 4 bytes <= 2*U_STAGE + 2*M_STAGE (W <= 4)
  constexpr int CRED_ROWS = (2 * U_STAGE + 2 * M_STAGE) / (8 * DV * 4);   // rows of W that fit at once
  for (int w0 = 0; w0 < W; w0 += CRED_ROWS) {
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      if (w >= w0 && w < w0 + CRED_ROWS && w < W) {
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          float* d = cred + ((size_t)(warp * CRED_ROWS + (w - w0)) * DV) + c * 256 + lane * 8;
          *reinterpret_cast<float4*>(d) = make_float4(acc[w][c][0], acc[w][c][1], acc[w][c][2], acc[w][c][3]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(acc[w][c][4], acc[w][c][5], acc[w][c][6], acc[w][c][7]);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < min(CRED_ROWS, W - w0) * DV; i += A2T) {
      const int wl = i / DV, col = i % DV;
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s += cred[(size_t)(g * CRED_ROWS + wl) * DV + col];
      ctx_part[((size_t)(r0 + w0 + wl) * nsplit + sp) * DV + col] = s;
    }
    __syncthreads();
  }
  if (warp < W && lane == 0) {
    float* st = stats + ((size_t)(r0 + warp) * nsplit + sp) * 4;
    st[0] = m_run; st[1] = l_run; st[2] = lw_run; st[3] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------ v3
// Warp-autonomous form: no block barrier inside the key loop.  A tile is 8 warps x KPT keys; warp w owns
// keys 4w..4w+3 of every tile and streams exactly those rows (Uk.mem row + value row) into a private
// double buffer with cp.async - padding keys are neither loaded nor evaluated, at key granularity.
// Lane l owns hidden units 8l..8l+7 (scores) and value columns 8l..8l+7 (context): per key a lane
// evaluates 8 x W tanh terms, the 4 x W partial sums of a tile are reduced across the warp with a
// 16-shuffle multi-value butterfly (lane l ends up with sum number l >> SH), and every warp keeps its
// own online softmax (max, sum, prior-weighted sum) and context accumulators - merged across the 8
// warps once, at the end.  The MUFU pipe (one tanh per (row, key, hidden unit)) is the only shared
// resource the warps contend for.
template <int WMAX, int DV, bool FAST>
__global__ __launch_bounds__(A2T) void additive_attn_v3_kernel(
    const float* __restrict__ qa, const bf16* __restrict__ U, const bf16* __restrict__ Mv,
    const float* __restrict__ vvec, const uint8_t* __restrict__ mask, const float* __restrict__ prior,
    const int32_t* __restrict__ tok, int tok_ld, int t, int W, int S, int nsplit, float* __restrict__ scores,
    float* __restrict__ stats, float* __restrict__ ctx_part, const int32_t* __restrict__ cidx,
    const int32_t* __restrict__ ncount, const int32_t* __restrict__ qorder) {
  constexpr int CG = DV / 256;
  constexpr int KPT = WMAX == 8 ? 2 : 4;               // keys per warp per tile
  constexpr int NV = KPT * WMAX;                       // partial sums per warp per tile (4, 8, 16)
  constexpr int LW = WMAX == 1 ? 0 : (WMAX == 2 ? 1 : (WMAX == 4 ? 2 : 3));
  constexpr int LNV = (KPT == 4 ? 2 : 1) + LW;         // log2(NV)
  constexpr int SH = 5 - LNV;                          // lane l holds sum number l >> SH
  constexpr int TILEK = 8 * KPT;
  constexpr int ROWB = (H + DV) * 2;                   // staged bytes per key: U row, then value row
  constexpr int WSTAGE = KPT * ROWB;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ float wst[8][WMAX][3];
  pdl_trigger();
  pdl_wait();
  // compacted form (cidx != NULL): the split walks the VALID keys of the query, cidx[b][j] = position of the
  // j-th one, every split of a query gets the same number of them, and heavy queries are launched first
  const int b = qorder ? qorder[blockIdx.x] : blockIdx.x, sp = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Sv = cidx ? ncount[b] : S;                 // keys to walk
  const int chunk = split_chunk(Sv, nsplit, A2_SPLIT);
  const int s_begin = sp * chunk, s_end = min(Sv, s_begin + chunk);
  const int ntiles = s_end > s_begin ? (s_end - s_begin + TILEK - 1) / TILEK : 0;
  const int r0 = b * W;
  const uint8_t* mb = mask + (size_t)b * S;
  const int32_t* cb = cidx ? cidx + (size_t)b * S : nullptr;
  const bf16* Ub = U + (size_t)b * S * H;
  const bf16* Mb = Mv + (size_t)b * S * DV;
  const float* pb = prior ? prior + (size_t)b * S : nullptr;
  unsigned char* wbuf = sm + (size_t)warp * 2 * WSTAGE;
  const uint32_t wbuf_s = smem_u32(wbuf);

  // this lane's slice of the queries and of v
  float qv[WMAX][8], vv[8];
#pragma unroll
  for (int w = 0; w < WMAX; ++w) {
    const float* q = qa + (size_t)(r0 + min(w, W - 1)) * H + lane * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(q), a1 = *reinterpret_cast<const float4*>(q + 4);
    qv[w][0] = a0.x; qv[w][1] = a0.y; qv[w][2] = a0.z; qv[w][3] = a0.w;
    qv[w][4] = a1.x; qv[w][5] = a1.y; qv[w][6] = a1.z; qv[w][7] = a1.w;
  }
  {
    const float4 a0 = *reinterpret_cast<const float4*>(vvec + lane * 8), a1 = *reinterpret_cast<const float4*>(vvec + lane * 8 + 4);
    vv[0] = a0.x; vv[1] = a0.y; vv[2] = a0.z; vv[3] = a0.w; vv[4] = a1.x; vv[5] = a1.y; vv[6] = a1.z; vv[7] = a1.w;
  }
  // the (key, row) this lane represents after the butterfly
  const int myj = lane >> SH, myk = myj >> LW, myw = myj & (WMAX - 1);
  bool rowvalid = myw < W;
  if (rowvalid && tok) rowvalid = tok[(size_t)(r0 + myw) * tok_ld + t] != 0;

  auto key_of = [&](int ti, int k) { return s_begin + ti * TILEK + warp * KPT + k; };
  // valid bits (bit k) and, in the compacted form, the memory positions of the group's keys (lane k holds key k's)
  int pos_next = 0;
  auto valid_bits = [&](int ti) -> unsigned {
    bool ok = false;
    pos_next = 0;
    if (lane < KPT && ti < ntiles) {
      const int s = key_of(ti, lane);
      if (cb) { ok = s < s_end; pos_next = ok ? cb[s] : 0; }
      else { ok = s < s_end && mb[s] != 0; pos_next = s; }
    }
    return __ballot_sync(0xffffffffu, ok);
  };
  auto issue = [&](int stage, unsigned vbits, int pos_lane) {
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int s = __shfl_sync(0xffffffffu, pos_lane, k);
      if ((vbits >> k) & 1u) {
        const uint32_t dst = wbuf_s + stage * WSTAGE + k * ROWB;
        a2_cp16(dst + lane * 16, Ub + (size_t)s * H + lane * 8, 16);
#pragma unroll
        for (int c = 0; c < CG; ++c)
          a2_cp16(dst + H * 2 + (c * 32 + lane) * 16, Mb + (size_t)s * DV + (c * 32 + lane) * 8, 16);
      }
    }
  };

  float m_run = -INFINITY, l_run = 0.f, lw_run = 0.f;  // statistics of row myw (replicated over its lanes)
  float acc[WMAX][CG][8];
#pragma unroll
  for (int w = 0; w < WMAX; ++w)
#pragma unroll
    for (int c = 0; c < CG; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[w][c][i] = 0.f;

  unsigned vb = valid_bits(0);
  int pos = pos_next;                                  // lane k: memory position of key k of the current tile
  issue(0, vb, pos);
  a2_commit();
  for (int ti = 0; ti < ntiles; ++ti) {
    const int stage = ti & 1;
    const unsigned vb_next = valid_bits(ti + 1);
    const int pos_n = pos_next;
    if (ti + 1 < ntiles) issue(stage ^ 1, vb_next, pos_n);
    a2_commit();
    a2_wait<1>();
    __syncwarp();
    const unsigned char* st = wbuf + stage * WSTAGE;
    // ---- partial scores of this lane's 8 hidden units
    float part[NV];
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
#pragma unroll
      for (int w = 0; w < WMAX; ++w) part[k * WMAX + w] = 0.f;
      if ((vb >> k) & 1u) {
        float u[8];
        ld8c(reinterpret_cast<const bf16*>(st + k * ROWB) + lane * 8, u);
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
          if (w < W) {
            float e = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = qv[w][i] + u[i];
              e = fmaf(vv[i], FAST ? tanh_fast(x) : tanh_acc(x), e);
            }
            part[k * WMAX + w] = e;
          }
        }
      }
    }
    // ---- multi-value butterfly: NV sums over 32 lanes in NV - 1 + SH shuffles
    {
      int n = NV;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
          n >>= 1;
          const bool hi = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < NV / 2; ++i) {
            if (i < n) {
              const float keep = hi ? part[i + n] : part[i];
              const float send = hi ? part[i] : part[i + n];
              part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        } else {
          part[0] += __shfl_xor_sync(0xffffffffu, part[0], off);
        }
      }
    }
    const int skey = __shfl_sync(0xffffffffu, pos, myk);        // memory position of this lane's key
    const bool inrange = key_of(ti, myk) < s_end;
    const bool ok = rowvalid && ((vb >> myk) & 1u);
    const float e = ok ? part[0] : -INFINITY;
    // (compacted form: padding positions are never visited; their scores were set to -inf at prefill)
    if ((lane & ((1 << SH) - 1)) == 0 && myw < W && inrange) scores[(size_t)(r0 + myw) * S + skey] = e;
    // ---- online softmax of row myw over the group's keys (the key index sits in the upper bits of myj)
    float tmax = e;
#pragma unroll
    for (int o = (1 << (SH + LW)); o < 32; o <<= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    const float mn = fmaxf(m_run, tmax);
    const float sc = (m_run == -INFINITY) ? 0.f : fexp(m_run - mn);
    const float p = (e == -INFINITY) ? 0.f : fexp(e - mn);
    float psum = p, pwsum = (pb && ok) ? pb[skey] * p : p;
#pragma unroll
    for (int o = (1 << (SH + LW)); o < 32; o <<= 1) {
      psum += __shfl_xor_sync(0xffffffffu, psum, o);
      pwsum += __shfl_xor_sync(0xffffffffu, pwsum, o);
    }
    l_run = fmaf(l_run, sc, psum);
    lw_run = fmaf(lw_run, sc, pwsum);
    m_run = mn;
    // ---- context: every lane needs all p[k][w] and the rescale factors
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      if (w < W) {
        const float scw = __shfl_sync(0xffffffffu, sc, w << SH);
#pragma unroll
        for (int c = 0; c < CG; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[w][c][i] *= scw;
      }
    }
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      float pk[WMAX];
#pragma unroll
      for (int w = 0; w < WMAX; ++w) pk[w] = __shfl_sync(0xffffffffu, p, (k * WMAX + w) << SH);
      if ((vb >> k) & 1u) {
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          float mv[8];
          ld8c(reinterpret_cast<const bf16*>(st + k * ROWB + H * 2) + c * 256 + lane * 8, mv);
#pragma unroll
          for (int w = 0; w < WMAX; ++w) {
            if (w < W) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[w][c][i] = fmaf(pk[w], mv[i], acc[w][c][i]);
            }
          }
        }
      }
    }
    __syncwarp();                                         // this stage is refilled two iterations from now
    vb = vb_next;
    pos = pos_n;
  }
  a2_wait<0>();
  // ---- merge the 8 warps: statistics, then the contexts through the (now idle) staging memory
  if ((lane & ((1 << SH) - 1)) == 0 && myk == 0 && myw < W) {
    wst[warp][myw][0] = m_run; wst[warp][myw][1] = l_run; wst[warp][myw][2] = lw_run;
  }
  __syncthreads();
  float fw[WMAX];                                         // this warp's rescale factor per row
#pragma unroll
  for (int w = 0; w < WMAX; ++w) {
    fw[w] = 0.f;
    if (w < W) {
      float Mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < 8; ++g) Mx = fmaxf(Mx, wst[g][w][0]);
      const float mw = wst[warp][w][0];
      fw[w] = (mw == -INFINITY) ? 0.f : fexp(mw - Mx);
    }
  }
  float* cred = reinterpret_cast<float*>(sm);
  constexpr int CRED_ROWS = (8 * 2 * WSTAGE) / (8 * DV * 4) < WMAX ? (8 * 2 * WSTAGE) / (8 * DV * 4) : WMAX;
  for (int w0 = 0; w0 < W; w0 += CRED_ROWS) {
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      if (w >= w0 && w < w0 + CRED_ROWS && w < W) {
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          float* d = cred + ((size_t)(warp * CRED_ROWS + (w - w0)) * DV) + c * 256 + lane * 8;
          *reinterpret_cast<float4*>(d) = make_float4(acc[w][c][0] * fw[w], acc[w][c][1] * fw[w], acc[w][c][2] * fw[w], acc[w][c][3] * fw[w]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(acc[w][c][4] * fw[w], acc[w][c][5] * fw[w], acc[w][c][6] * fw[w], acc[w][c][7] * fw[w]);
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < min(CRED_ROWS, W - w0) * DV; i += A2T) {
      const int wl = i / DV, col = i % DV;
      float s2 = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s2 += cred[(size_t)(g * CRED_ROWS + wl) * DV + col];
      ctx_part[((size_t)(r0 + w0 + wl) * nsplit + sp) * DV + col] = s2;
    }
    __syncthreads();
  }
  if (tid < W) {
    float Mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < 8; ++g) Mx = fmaxf(Mx, wst[g][tid][0]);
    float l = 0.f, lw = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float f = (wst[g][tid][0] == -INFINITY) ? 0.f : fexp(wst[g][tid][0] - Mx);
      l = fmaf(wst[g][tid][1], f, l);
      lw = fmaf(wst[g][tid][2], f, lw);
    }
    float* so = stats + ((size_t)(r0 + tid) * nsplit + sp) * 4;
    so[0] = Mx; so[1] = l; so[2] = lw; so[3] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------ gate form
// CaSE reads the contexts m_i = sum_s a_i[s] mem_i[s] ONLY inside the 3-way mixture gate
// softmax(W_m [h; m_0; m_1] + b_m) (Model.py:39,117: c_m goes nowhere else), and the gate logits are linear
// in m_i:  W_m,i . m_i = sum_s a_i[s] (W_m,i . mem_i[s]).  The prefill therefore projects every key once to
// G[b][s][0..2] = W_m[:, H(1+i):H(2+i)] . mem_i[b][s] (fp32), and the step accumulates 3 numbers per
// (row, key) instead of H: the value rows - half of the HBM stream and 32 FFMA per key per lane of the
// kernel above - disappear, what is left is the tanh count.  Same warp-autonomous walk as v3 (per-warp
// cp.async ring over the Uk.mem rows of the warp's keys, padding skipped per key, multi-value butterfly);
// lane (key k, row w) keeps the three gate sums of its own (key, row) pairs and they are reduced once.
// Splits are work-proportional (nsq[b] of the nsplit slots per query) so that the launch is ONE resident wave
// of equally long CTAs (3 per SM).
// Outputs: scores, stats as above, gate_part [R][nsplit][4] = sum exp(e - m) * G (relative to stats' m).
// (A lane = (key, hidden quarter) mapping with q in shared memory - two shuffles per key instead of the
//  butterfly - executes as many instructions (its per-warp prologue and the LDS traffic eat the gain) and was
//  slower at every split policy: 52-62 us against 48 us at the BASELINE shape.)
constexpr int AG_NST = 3;     // ring stages per warp
template <int WMAX, bool FAST, bool FULLW>    // FULLW: W == WMAX, the row checks of the tanh block fold away
__global__ __launch_bounds__(A2T, WMAX <= 4 ? 3 : 2) void additive_attn_gate_kernel(
    const float* __restrict__ qa, const bf16* __restrict__ U, const float4* __restrict__ G,
    const float* __restrict__ vvec, const uint8_t* __restrict__ mask, const float* __restrict__ prior,
    const int32_t* __restrict__ tok, int tok_ld, int t, int W, int S, int nsplit, float* __restrict__ scores,
    float* __restrict__ stats, float* __restrict__ gate_part, const int32_t* __restrict__ cidx,
    const int32_t* __restrict__ ncount, const int32_t* __restrict__ qorder, const int32_t* __restrict__ nsq, int evict) {
  const uint64_t pol = l2_stream_policy(evict);
  constexpr int KPT = WMAX == 8 ? 2 : 4;               // keys per warp per tile
  constexpr int NV = KPT * WMAX;                       // partial sums per warp per tile (4, 8, 16)
  constexpr int LW = WMAX == 1 ? 0 : (WMAX == 2 ? 1 : (WMAX == 4 ? 2 : 3));
  constexpr int LNV = (KPT == 4 ? 2 : 1) + LW;         // log2(NV)
  constexpr int SH = 5 - LNV;                          // lane l holds sum number l >> SH
  constexpr int TILEK = 8 * KPT;
  constexpr int ROWB = H * 2;                          // staged bytes per key: the U row
  constexpr int WSTAGE = KPT * ROWB;
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ float wst[8][WMAX][6];
  pdl_trigger();
  // everything up to the first key rows in flight reads prefill data only (split plan, key positions, Uk.mem);
  // the wait for the producer of qa / tok comes after it
  const int b = qorder ? qorder[blockIdx.x] : blockIdx.x, sp = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Sv = cidx ? ncount[b] : S;                 // keys to walk
  const int nsb = nsq ? max(1, min(nsq[b], nsplit)) : nsplit;
  if (sp >= nsb) {
    pdl_wait();
    if (tid < W) {
      const size_t o = (size_t)(b * W + tid) * nsplit + sp;
      reinterpret_cast<float4*>(stats)[o] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(gate_part)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int chunk = split_chunk(Sv, nsb, A2_SPLIT);
  const int s_begin = sp * chunk, s_end = min(Sv, s_begin + chunk);
  const int ntiles = s_end > s_begin ? (s_end - s_begin + TILEK - 1) / TILEK : 0;
  const int r0 = b * W;
  const uint8_t* mb = mask + (size_t)b * S;
  const int32_t* cb = cidx ? cidx + (size_t)b * S : nullptr;
  const bf16* Ub = U + (size_t)b * S * H;
  const float4* Gb = G + (size_t)b * S;
  const float* pb = prior ? prior + (size_t)b * S : nullptr;
  unsigned char* wbuf = sm + (size_t)warp * AG_NST * WSTAGE;
  const uint32_t wbuf_s = smem_u32(wbuf);

  const int myj = lane >> SH, myk = myj >> LW, myw = myj & (WMAX - 1);

  auto key_of = [&](int ti, int k) { return s_begin + ti * TILEK + warp * KPT + k; };
  int pos_next = 0;
  auto valid_bits = [&](int ti) -> unsigned {
    bool ok = false;
    pos_next = 0;
    if (lane < KPT && ti < ntiles) {
      const int s = key_of(ti, lane);
      if (cb) { ok = s < s_end; pos_next = ok ? cb[s] : 0; }
      else { ok = s < s_end && mb[s] != 0; pos_next = s; }
    }
    return __ballot_sync(0xffffffffu, ok);
  };
  auto issue = [&](int stage, unsigned vbits, int pos_lane) {
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int s = __shfl_sync(0xffffffffu, pos_lane, k);
      if ((vbits >> k) & 1u) a2_cp16_hint(wbuf_s + stage * WSTAGE + k * ROWB + lane * 16, Ub + (size_t)s * H + lane * 8, pol);
    }
  };

  float m_run = -INFINITY, l_run = 0.f, lw_run = 0.f;  // statistics of row myw (replicated over its lanes)
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;                  // gate sums of this lane's (key slot, row) pairs

  // tiles ti .. ti + AG_NST - 2 are in flight at the top of iteration ti
  unsigned vbq[AG_NST - 1];
  int posq[AG_NST - 1];
#pragma unroll
  for (int j = 0; j < AG_NST - 1; ++j) {
    vbq[j] = valid_bits(j);
    posq[j] = pos_next;
    if (j < ntiles) issue(j, vbq[j], posq[j]);
    a2_commit();
  }
  pdl_wait();
  float qv[WMAX][8], vv[8];
#pragma unroll
  for (int w = 0; w < WMAX; ++w) {
    const float* q = qa + (size_t)(r0 + min(w, W - 1)) * H + lane * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(q), a1 = *reinterpret_cast<const float4*>(q + 4);
    qv[w][0] = a0.x; qv[w][1] = a0.y; qv[w][2] = a0.z; qv[w][3] = a0.w;
    qv[w][4] = a1.x; qv[w][5] = a1.y; qv[w][6] = a1.z; qv[w][7] = a1.w;
  }
  {
    const float4 a0 = *reinterpret_cast<const float4*>(vvec + lane * 8), a1 = *reinterpret_cast<const float4*>(vvec + lane * 8 + 4);
    vv[0] = a0.x; vv[1] = a0.y; vv[2] = a0.z; vv[3] = a0.w; vv[4] = a1.x; vv[5] = a1.y; vv[6] = a1.z; vv[7] = a1.w;
  }
  bool rowvalid = myw < W;
  if (rowvalid && tok) rowvalid = tok[(size_t)(r0 + myw) * tok_ld + t] != 0;
  int stage = 0, fstage = AG_NST - 1;
  for (int ti = 0; ti < ntiles; ++ti) {
    const unsigned vb_new = valid_bits(ti + AG_NST - 1);
    const int pos_new = pos_next;
    if (ti + AG_NST - 1 < ntiles) issue(fstage, vb_new, pos_new);
    a2_commit();
    const unsigned vb = vbq[0];
    const int pos = posq[0];
    // gate projections and prior of this lane's key: in flight during the tanh block
    const int skey = __shfl_sync(0xffffffffu, pos, myk);
    const bool ok = rowvalid && ((vb >> myk) & 1u);
    float4 gk = make_float4(0.f, 0.f, 0.f, 0.f);
    float pr = 1.f;
    if (ok) {
      gk = __ldg(Gb + skey);
      if (pb) pr = __ldg(pb + skey);
    }
    a2_wait<AG_NST - 1>();
    __syncwarp();
    const unsigned char* st = wbuf + stage * WSTAGE;
    float part[NV];
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
#pragma unroll
      for (int w = 0; w < WMAX; ++w) part[k * WMAX + w] = 0.f;
      if ((vb >> k) & 1u) {
        float u[8];
        ld8c(reinterpret_cast<const bf16*>(st + k * ROWB) + lane * 8, u);
#pragma unroll
        for (int w = 0; w < WMAX; ++w) {
          if (FULLW || w < W) {
            float e = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = qv[w][i] + u[i];
              e = fmaf(vv[i], FAST ? tanh_fast(x) : tanh_acc(x), e);
            }
            part[k * WMAX + w] = e;
          }
        }
      }
    }
    {
      int n = NV;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
          n >>= 1;
          const bool hi = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < NV / 2; ++i) {
            if (i < n) {
              const float keep = hi ? part[i + n] : part[i];
              const float send = hi ? part[i] : part[i + n];
              part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        } else {
          part[0] += __shfl_xor_sync(0xffffffffu, part[0], off);
        }
      }
    }
    const bool inrange = key_of(ti, myk) < s_end;
    const float e = ok ? part[0] : -INFINITY;
    if ((lane & ((1 << SH) - 1)) == 0 && myw < W && inrange) scores[(size_t)(r0 + myw) * S + skey] = e;
    float tmax = e;
#pragma unroll
    for (int o = (1 << (SH + LW)); o < 32; o <<= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    const float mn = fmaxf(m_run, tmax);
    const float sc = (m_run == -INFINITY) ? 0.f : fexp(m_run - mn);
    const float p = (e == -INFINITY) ? 0.f : fexp(e - mn);
    float psum = p, pwsum = pr * p;
#pragma unroll
    for (int o = (1 << (SH + LW)); o < 32; o <<= 1) {
      psum += __shfl_xor_sync(0xffffffffu, psum, o);
      pwsum += __shfl_xor_sync(0xffffffffu, pwsum, o);
    }
    l_run = fmaf(l_run, sc, psum);
    lw_run = fmaf(lw_run, sc, pwsum);
    m_run = mn;
    g0 = fmaf(g0, sc, p * gk.x);
    g1 = fmaf(g1, sc, p * gk.y);
    g2 = fmaf(g2, sc, p * gk.z);
    __syncwarp();                                         // this stage is refilled next iteration
#pragma unroll
    for (int j = 0; j + 1 < AG_NST - 1; ++j) { vbq[j] = vbq[j + 1]; posq[j] = posq[j + 1]; }
    vbq[AG_NST - 2] = vb_new;
    posq[AG_NST - 2] = pos_new;
    fstage = stage;
    stage = stage + 1 == AG_NST ? 0 : stage + 1;
  }
  a2_wait<0>();
  // the key slots of a row sit in the upper lane bits: one reduction for the whole walk
#pragma unroll
  for (int o = (1 << (SH + LW)); o < 32; o <<= 1) {
    g0 += __shfl_xor_sync(0xffffffffu, g0, o);
    g1 += __shfl_xor_sync(0xffffffffu, g1, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
  }
  if ((lane & ((1 << SH) - 1)) == 0 && myk == 0 && myw < W) {
    float* d = wst[warp][myw];
    d[0] = m_run; d[1] = l_run; d[2] = lw_run; d[3] = g0; d[4] = g1; d[5] = g2;
  }
  __syncthreads();
  if (tid < W) {
    float Mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < 8; ++g) Mx = fmaxf(Mx, wst[g][tid][0]);
    float l = 0.f, lw = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float f = (wst[g][tid][0] == -INFINITY) ? 0.f : fexp(wst[g][tid][0] - Mx);
      l = fmaf(wst[g][tid][1], f, l);
      lw = fmaf(wst[g][tid][2], f, lw);
      a0 = fmaf(wst[g][tid][3], f, a0);
      a1 = fmaf(wst[g][tid][4], f, a1);
      a2 = fmaf(wst[g][tid][5], f, a2);
    }
    const size_t o = (size_t)(r0 + tid) * nsplit + sp;
    reinterpret_cast<float4*>(stats)[o] = make_float4(Mx, l, lw, 0.f);
    reinterpret_cast<float4*>(gate_part)[o] = make_float4(a0, a1, a2, 0.f);
  }
}

// ------------------------------------------------------------------------------------------ gate form, f16 / tensor-core
// The kernel above is issue-bound, not MUFU-bound: 6.2 warp instructions per tanh (FADD + MUFU + FFMA, bf16
// unpacks, the butterfly's SHFL + FSEL).  Here the additions and the tanh are packed (HADD2, tanh.approx.f16x2:
// one instruction per two elements at the same element rate) and the weighted sum over the hidden units - the
// v . tanh(...) reduction - is done by the tensor core: the tanh values of 16 (key, row) pairs x 16 hidden units
// are exactly the A fragment of mma.m16n8k16 as the lanes produce them, B = v in every column, so D[row][*] is
// the score and the cross-lane reduction costs nothing.  Uk.mem is stored in f16 for this kernel (prefill: 10
// mantissa bits instead of bf16's 7, the dominant error of the bf16 form), q and v are rounded to f16,
// accumulation is fp32.
// MEASURED (B200, BASELINE shape): ptxas turns tanh.approx.f16x2 into TWO MUFU.TANH.F16 plus a PRMT, so the MUFU
// count is unchanged (3.5 M); instructions drop from 21.7 M to 17.1 M but the launch takes 56 us against 47 us
// for the fp32-math kernel above (each HMMA waits for eight MUFU results and chains on one accumulator).  Off
// by default (case_set_gate_f16), kept with its tests as the record of that experiment.
//   tile      = KPT = 16 / WMAX keys x WMAX rows = the 16 rows of the MMA, row = key * WMAX + w
//   lane      = (g = lane / 4, tq = lane % 4): rows g and g + 8 (same w, two keys), hidden units 64 tq .. 64 tq + 63
//   k order   = hidden unit 64 tq + 4 c + j is column (j < 2 ? 2 tq + j : 2 tq + 6 + j) of k-step c (any fixed
//               permutation of the hidden units is fine: q, U and v use the same one)
//   smem row  = the 512-byte U row as four 128-byte quarters at a 144-byte stride (conflict-free LDS.128)
constexpr int AH_QS = 144;                // quarter stride (bytes)
constexpr int AH_ROWB = 4 * AH_QS;        // staged bytes per key
__device__ __forceinline__ uint32_t ah_tanh2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t ah_add2(uint32_t a, uint32_t b) {
  uint32_t y;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b));
  return y;
}
__device__ __forceinline__ uint32_t ah_pack(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void ah_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int WMAX>
__global__ __launch_bounds__(A2T, WMAX == 4 ? 3 : 2) void additive_attn_gate_h_kernel(
    const float* __restrict__ qa, const __half* __restrict__ U, const float4* __restrict__ G,
    const float* __restrict__ vvec, const uint8_t* __restrict__ mask, const float* __restrict__ prior,
    const int32_t* __restrict__ tok, int tok_ld, int t, int W, int S, int nsplit, float* __restrict__ scores,
    float* __restrict__ stats, float* __restrict__ gate_part, const int32_t* __restrict__ cidx,
    const int32_t* __restrict__ ncount, const int32_t* __restrict__ qorder, const int32_t* __restrict__ nsq) {
  static_assert(WMAX == 2 || WMAX == 4 || WMAX == 8, "one MMA tile is 16 / WMAX keys");
  constexpr int KPT = 16 / WMAX;                       // keys per warp per tile
  constexpr int TILEK = 8 * KPT;
  constexpr int WSTAGE = KPT * AH_ROWB;
  extern __shared__ __align__(128) unsigned char sm[];
  __half* v_s = reinterpret_cast<__half*>(sm + 8 * AG_NST * WSTAGE);      // v in f16, quarters at the 144-byte stride
  __shared__ float wst[8][WMAX][6];
  pdl_trigger();
  const int b = qorder ? qorder[blockIdx.x] : blockIdx.x, sp = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Sv = cidx ? ncount[b] : S;                 // keys to walk
  const int nsb = nsq ? max(1, min(nsq[b], nsplit)) : nsplit;
  if (sp >= nsb) {
    pdl_wait();
    if (tid < W) {
      const size_t o = (size_t)(b * W + tid) * nsplit + sp;
      reinterpret_cast<float4*>(stats)[o] = make_float4(-INFINITY, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(gate_part)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int chunk = split_chunk(Sv, nsb, A2_SPLIT);
  const int s_begin = sp * chunk, s_end = min(Sv, s_begin + chunk);
  const int ntiles = s_end > s_begin ? (s_end - s_begin + TILEK - 1) / TILEK : 0;
  const int r0 = b * W;
  const uint8_t* mb = mask + (size_t)b * S;
  const int32_t* cb = cidx ? cidx + (size_t)b * S : nullptr;
  const __half* Ub = U + (size_t)b * S * H;
  const float4* Gb = G + (size_t)b * S;
  const float* pb = prior ? prior + (size_t)b * S : nullptr;
  unsigned char* wbuf = sm + (size_t)warp * AG_NST * WSTAGE;
  const uint32_t wbuf_s = smem_u32(wbuf);
  const int g = lane >> 2, tq = lane & 3;
  const int keyA = g / WMAX, keyB = (g + 8) / WMAX, myw = g % WMAX;

  auto key_of = [&](int ti, int k) { return s_begin + ti * TILEK + warp * KPT + k; };
  // lane k < KPT owns the bookkeeping of key k of a tile: memory position (or, masked form, ~position) and validity
  int pos_next = 0;
  auto valid_bits = [&](int ti) -> unsigned {
    bool ok = false;
    pos_next = 0;
    if (lane < KPT && ti < ntiles) {
      const int s = key_of(ti, lane);
      if (cb) { ok = s < s_end; pos_next = ok ? cb[s] : 0; }
      else { ok = s < s_end && mb[s] != 0; pos_next = s; }
    }
    return __ballot_sync(0xffffffffu, ok);
  };
  auto issue = [&](int stage, unsigned vbits, int pos_lane) {
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
      const int s = __shfl_sync(0xffffffffu, pos_lane, k);
      if ((vbits >> k) & 1u)
        a2_cp16(wbuf_s + stage * WSTAGE + k * AH_ROWB + (lane >> 3) * AH_QS + (lane & 7) * 16, Ub + (size_t)s * H + lane * 8, 16);
    }
  };
  unsigned vbq[AG_NST - 1];
  int posq[AG_NST - 1];
#pragma unroll
  for (int j = 0; j < AG_NST - 1; ++j) {
    vbq[j] = valid_bits(j);
    posq[j] = pos_next;
    if (j < ntiles) issue(j, vbq[j], posq[j]);
    a2_commit();
  }
  // v (a weight) in f16 -> shared memory
  for (int i = tid; i < H; i += A2T) v_s[(i >> 6) * (AH_QS / 2) + (i & 63)] = __float2half_rn(vvec[i]);
  pdl_wait();
  // this lane's 64 hidden units of q[row myw] as 32 f16 pairs: qh[2 c + j2] = units 64 tq + 4 c + 2 j2, + 1
  uint32_t qh[32];
  {
    const float* q = qa + (size_t)(r0 + min(myw, W - 1)) * H + tq * 64;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 x = *reinterpret_cast<const float4*>(q + c * 4);
      qh[2 * c] = ah_pack(x.x, x.y);
      qh[2 * c + 1] = ah_pack(x.z, x.w);
    }
  }
  bool rowvalid = myw < W;
  if (rowvalid && tok) rowvalid = tok[(size_t)(r0 + myw) * tok_ld + t] != 0;
  __syncthreads();                                      // v_s
  const uint32_t v_q = smem_u32(v_s) + tq * AH_QS;

  float m_run = -INFINITY, l_run = 0.f, lw_run = 0.f;  // row myw over the keys this lane group sees (merged per tile)
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;                  // gate sums of this lane's two (key slot, row) pairs
  int stage = 0, fstage = AG_NST - 1;
  for (int ti = 0; ti < ntiles; ++ti) {
    const unsigned vb_new = valid_bits(ti + AG_NST - 1);
    const int pos_new = pos_next;
    if (ti + AG_NST - 1 < ntiles) issue(fstage, vb_new, pos_new);
    a2_commit();
    const unsigned vb = vbq[0];
    const int pos = posq[0];
    const int posA = __shfl_sync(0xffffffffu, pos, keyA), posB = __shfl_sync(0xffffffffu, pos, keyB);
    const bool okA = rowvalid && ((vb >> keyA) & 1u), okB = rowvalid && ((vb >> keyB) & 1u);
    float4 gA = make_float4(0.f, 0.f, 0.f, 0.f), gB = gA;
    float prA = 1.f, prB = 1.f;
    if (okA) { gA = __ldg(Gb + posA); if (pb) prA = __ldg(pb + posA); }
    if (okB) { gB = __ldg(Gb + posB); if (pb) prB = __ldg(pb + posB); }
    a2_wait<AG_NST - 1>();
    __syncwarp();
    const uint32_t ua_s = wbuf_s + stage * WSTAGE + keyA * AH_ROWB + tq * AH_QS;
    const uint32_t ub_s = wbuf_s + stage * WSTAGE + keyB * AH_ROWB + tq * AH_QS;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (vb != 0u) {
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2) {                    // two k-steps of the MMA per 16-byte load
        uint32_t ua[4], ub[4], vv[4];
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ua[0]), "=r"(ua[1]), "=r"(ua[2]), "=r"(ua[3]) : "r"(ua_s + c2 * 16));
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ub[0]), "=r"(ub[1]), "=r"(ub[2]), "=r"(ub[3]) : "r"(ub_s + c2 * 16));
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]) : "r"(v_q + c2 * 16));
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const uint32_t q0 = qh[4 * c2 + 2 * h2], q1 = qh[4 * c2 + 2 * h2 + 1];
          const uint32_t a0 = ah_tanh2(ah_add2(q0, ua[2 * h2])), a1 = ah_tanh2(ah_add2(q0, ub[2 * h2]));
          const uint32_t a2 = ah_tanh2(ah_add2(q1, ua[2 * h2 + 1])), a3 = ah_tanh2(ah_add2(q1, ub[2 * h2 + 1]));
          ah_mma(acc, a0, a1, a2, a3, vv[2 * h2], vv[2 * h2 + 1]);
        }
      }
    }
    // every column of D holds the score of its row: rows g (key A) and g + 8 (key B)
    const float eA = okA ? acc[0] : -INFINITY, eB = okB ? acc[2] : -INFINITY;
    if (myw < W) {
      if (tq == 0 && ((vb >> keyA) & 1u || (!cb && key_of(ti, keyA) < s_end))) scores[(size_t)(r0 + myw) * S + posA] = eA;
      if (tq == 1 && ((vb >> keyB) & 1u || (!cb && key_of(ti, keyB) < s_end))) scores[(size_t)(r0 + myw) * S + posB] = eB;
    }
    // online softmax of row myw over the tile's keys: two here, the others WMAX * 4 .. 16 lanes away
    float tmax = fmaxf(eA, eB);
#pragma unroll
    for (int o = WMAX * 4; o < 32; o <<= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    const float mn = fmaxf(m_run, tmax);
    const float sc = (m_run == -INFINITY) ? 0.f : fexp(m_run - mn);
    const float pA = (eA == -INFINITY) ? 0.f : fexp(eA - mn), pB = (eB == -INFINITY) ? 0.f : fexp(eB - mn);
    float psum = pA + pB, pwsum = fmaf(prA, pA, prB * pB);
#pragma unroll
    for (int o = WMAX * 4; o < 32; o <<= 1) {
      psum += __shfl_xor_sync(0xffffffffu, psum, o);
      pwsum += __shfl_xor_sync(0xffffffffu, pwsum, o);
    }
    l_run = fmaf(l_run, sc, psum);
    lw_run = fmaf(lw_run, sc, pwsum);
    m_run = mn;
    g0 = fmaf(g0, sc, fmaf(pA, gA.x, pB * gB.x));
    g1 = fmaf(g1, sc, fmaf(pA, gA.y, pB * gB.y));
    g2 = fmaf(g2, sc, fmaf(pA, gA.z, pB * gB.z));
    __syncwarp();                                         // this stage is refilled next iteration
#pragma unroll
    for (int j = 0; j + 1 < AG_NST - 1; ++j) { vbq[j] = vbq[j + 1]; posq[j] = posq[j + 1]; }
    vbq[AG_NST - 2] = vb_new;
    posq[AG_NST - 2] = pos_new;
    fstage = stage;
    stage = stage + 1 == AG_NST ? 0 : stage + 1;
  }
  a2_wait<0>();
#pragma unroll
  for (int o = WMAX * 4; o < 32; o <<= 1) {
    g0 += __shfl_xor_sync(0xffffffffu, g0, o);
    g1 += __shfl_xor_sync(0xffffffffu, g1, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
  }
  if (tq == 0 && g < WMAX && myw < W) {
    float* d = wst[warp][myw];
    d[0] = m_run; d[1] = l_run; d[2] = lw_run; d[3] = g0; d[4] = g1; d[5] = g2;
  }
  __syncthreads();
  if (tid < W) {
    float Mx = -INFINITY;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) Mx = fmaxf(Mx, wst[gg][tid][0]);
    float l = 0.f, lw = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) {
      const float f = (wst[gg][tid][0] == -INFINITY) ? 0.f : fexp(wst[gg][tid][0] - Mx);
      l = fmaf(wst[gg][tid][1], f, l);
      lw = fmaf(wst[gg][tid][2], f, lw);
      a0 = fmaf(wst[gg][tid][3], f, a0);
      a1 = fmaf(wst[gg][tid][4], f, a1);
      a2 = fmaf(wst[gg][tid][5], f, a2);
    }
    const size_t o = (size_t)(r0 + tid) * nsplit + sp;
    reinterpret_cast<float4*>(stats)[o] = make_float4(Mx, l, lw, 0.f);
    reinterpret_cast<float4*>(gate_part)[o] = make_float4(a0, a1, a2, 0.f);
  }
}

template <int WMAX>
static int launch_gate_h(const float* qa, const void* U, const float* G, const float* v, const uint8_t* mask,
                         const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int nsplit,
                         float* scores, float* stats, float* gate_part, const int32_t* cidx, const int32_t* ncount,
                         const int32_t* qorder, const int32_t* nsq, cudaStream_t st) {
  constexpr int KPT = 16 / WMAX;
  const size_t smem = (size_t)8 * AG_NST * KPT * AH_ROWB + 4 * AH_QS;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(additive_attn_gate_h_kernel<WMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    attr = true;
  }
  launch_k(additive_attn_gate_h_kernel<WMAX>, dim3(B, nsplit), A2T, smem, st, qa, (const __half*)U, (const float4*)G, v, mask,
           prior, tok, tok_ld, t, W, S, nsplit, scores, stats, gate_part, cidx, ncount, qorder, nsq);
  return check_launch("case_additive_attn_gate_h");
}

// ---- prefill side of the gate form
// G[n][0..2] = Wg[0..2][:] . mem[n][:] for N key rows (bf16 storage, fp32 accumulate): one warp per row, HBM-bound
__global__ __launch_bounds__(256) void gate_project_kernel(const bf16* __restrict__ mem, const float* __restrict__ Wg,
                                                           float4* __restrict__ G, long long N) {
  const int lane = threadIdx.x & 31;
  float wg[3][8];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(Wg + j * H + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(Wg + j * H + lane * 8 + 4));
    wg[j][0] = a.x; wg[j][1] = a.y; wg[j][2] = a.z; wg[j][3] = a.w; wg[j][4] = b.x; wg[j][5] = b.y; wg[j][6] = b.z; wg[j][7] = b.w;
  }
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N; n += nw) {
    float m[8];
    ld8(mem + n * H + lane * 8, m);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { d0 = fmaf(wg[0][i], m[i], d0); d1 = fmaf(wg[1][i], m[i], d1); d2 = fmaf(wg[2][i], m[i], d2); }
    d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
    if (lane == 0) G[n] = make_float4(d0, d1, d2, 0.f);
  }
}

// work-proportional split plan: keys per CTA = max(ceil(sum / slots), ceil(max / max_split)) rounded up to 32,
// nsq[b] = clamp(ceil(count[b] / that), 1, max_split).  One CTA (B is at most a few hundred).
__global__ __launch_bounds__(256) void split_plan_kernel(const int32_t* __restrict__ count, int B, int slots, int max_split,
                                                         int32_t* __restrict__ nsq) {
  __shared__ long long ssum[8];
  __shared__ int smax[8];
  long long sum = 0;
  int mx = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) { const int c = count[b]; sum += c; mx = max(mx, c); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  sum = 0; mx = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sum += ssum[w]; mx = max(mx, smax[w]); }
  long long chunk = (sum + slots - 1) / slots;
  const long long cmin = (mx + max_split - 1) / max_split;
  if (cmin > chunk) chunk = cmin;
  chunk = (chunk + 31) / 32 * 32;
  if (chunk < 32) chunk = 32;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int n = (int)((count[b] + chunk - 1) / chunk);
    nsq[b] = n < 1 ? 1 : (n > max_split ? max_split : n);
  }
}

template <int WMAX>
static int launch_gate(int fast, const float* qa, const void* U, const float* G, const float* v, const uint8_t* mask,
                       const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int nsplit,
                       float* scores, float* stats, float* gate_part, const int32_t* cidx, const int32_t* ncount,
                       const int32_t* qorder, const int32_t* nsq, cudaStream_t st) {
  constexpr int KPT = WMAX == 8 ? 2 : 4;
  const size_t smem = (size_t)8 * AG_NST * KPT * H * 2;
  const bool full = W == WMAX;
#define AG_LAUNCH(FAST_, FULL_)                                                                                           \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      cudaFuncSetAttribute(additive_attn_gate_kernel<WMAX, FAST_, FULL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    launch_k(additive_attn_gate_kernel<WMAX, FAST_, FULL_>, dim3(B, nsplit), A2T, smem, st, qa, (const bf16*)U,            \
             (const float4*)G, v, mask, prior, tok, tok_ld, t, W, S, nsplit, scores, stats, gate_part, cidx, ncount, qorder, nsq, g_evict_first); \
  } while (0)
  if (fast) { if (full) AG_LAUNCH(true, true); else AG_LAUNCH(true, false); }
  else { if (full) AG_LAUNCH(false, true); else AG_LAUNCH(false, false); }
#undef AG_LAUNCH
  return check_launch("case_additive_attn_gate");
}

template <int WMAX, int DV, bool FAST>
static int launch_v2(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                     const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int nsplit,
                     float* scores, float* stats, float* ctx_part, cudaStream_t st) {
  const int32_t* cidx = g_add_cidx; const int32_t* ncount = g_add_ncount; const int32_t* qorder = g_add_qorder;
  g_add_cidx = g_add_ncount = g_add_qorder = nullptr;          // one-shot: set by case_additive_attn_compact
  if (g_additive_impl == 2) {
    const int chunk = split_chunk(S, nsplit, A2_SPLIT);
    const size_t smem = (size_t)2 * A2K * A2ULD * 2 + (size_t)2 * A2K * DV * 2 +
                        sizeof(float) * ((size_t)WMAX * H + H + 8 * WMAX * A2K + WMAX * A2K + 8) +
                        sizeof(uint32_t) * (size_t)(chunk / A2K + 1);
    auto kern = additive_attn_v2_kernel<WMAX, DV, FAST>;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr = true;
    }
    launch_k(kern, dim3(B, nsplit), A2T, smem, st, qa, (const bf16*)U, (const bf16*)Mv, v, mask, prior, tok, tok_ld, t, W, S,
                                             nsplit, scores, stats, ctx_part);
    return check_launch("case_additive_attn(v2)");
  }
  constexpr int KPT = WMAX == 8 ? 2 : 4;
  const size_t smem = (size_t)8 * 2 * KPT * (H + DV) * 2;
  auto kern = additive_attn_v3_kernel<WMAX, DV, FAST>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  launch_k(kern, dim3(B, nsplit), A2T, smem, st, qa, (const bf16*)U, (const bf16*)Mv, v, mask, prior, tok, tok_ld, t, W, S,
                                           nsplit, scores, stats, ctx_part, cidx, ncount, qorder);
  return check_launch("case_additive_attn(v3)");
}

template <int DV, bool FAST>
static int dispatch_v2(int W, const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                       const float* prior, const int32_t* tok, int tok_ld, int t, int B, int S, int nsplit,
                       float* scores, float* stats, float* ctx_part, cudaStream_t st) {
  if (W <= 1) return launch_v2<1, DV, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, scores, stats, ctx_part, st);
  if (W <= 2) return launch_v2<2, DV, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, scores, stats, ctx_part, st);
  if (W <= 4) return launch_v2<4, DV, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, scores, stats, ctx_part, st);
  return launch_v2<8, DV, FAST>(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, scores, stats, ctx_part, st);
}

}  // namespace cb

int case_additive_attn_v2(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                          const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int DV,
                          int nsplit, float* scores, float* stats, float* ctx_part, int fast_tanh, cudaStream_t st) {
  using namespace cb;
  if (DV == 256) {
    if (fast_tanh) return dispatch_v2<256, true>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, nsplit, scores, stats, ctx_part, st);
    return dispatch_v2<256, false>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, nsplit, scores, stats, ctx_part, st);
  }
  if (fast_tanh) return dispatch_v2<512, true>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, nsplit, scores, stats, ctx_part, st);
  return dispatch_v2<512, false>(W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, nsplit, scores, stats, ctx_part, st);
}

extern "C" int case_set_additive_impl(int impl) {
  const int old = cb::g_additive_impl;
  cb::g_additive_impl = impl == 2 ? 2 : 3;
  return old;
}

/* case_additive_attn over the VALID keys only (bf16, warp-autonomous kernel): cidx int32 [B][S] positions
 * of the valid keys (ascending), ncount int32 [B], qorder int32 [B] launch order of the queries (heaviest
 * first; may be NULL).  scores[r][s] of padding positions are NOT written: fill them with -inf once. */
extern "C" int case_additive_attn_compact(const float* qa, const void* U, const void* Mv, const float* v,
                                          const uint8_t* mask, const float* prior, const int32_t* tok, int tok_ld, int t,
                                          int B, int W, int S, int DV, int nsplit, float* attn_un, float* stats,
                                          float* ctx_part, int fast_tanh, const int32_t* cidx, const int32_t* ncount,
                                          const int32_t* qorder, case_stream_t stream) {
  if (!cidx || !ncount) { cb::set_error("case_additive_attn_compact: cidx / ncount missing"); return CASE_EINVAL; }
  cb::g_add_cidx = cidx; cb::g_add_ncount = ncount; cb::g_add_qorder = qorder;
  const int old = cb::g_additive_impl;
  cb::g_additive_impl = 3;
  const int rc = case_additive_attn_v2(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats,
                                       ctx_part, fast_tanh, (cudaStream_t)stream);
  cb::g_additive_impl = old;
  cb::g_add_cidx = cb::g_add_ncount = cb::g_add_qorder = nullptr;
  return rc;
}

/* Gate form of the additive attention (bf16 keys): G fp32 [B][S][4] = (W_m slice of memory i) . mem[b][s]
 * (3 gate logits' worth per key, 4th unused) replaces the value rows; gate_part [R][nsplit][4] =
 * sum exp(e - m) * G relative to the split's m in stats.  cidx / ncount / qorder as in
 * case_additive_attn_compact, or all NULL to walk every position under the mask.  nsq int32 [B] (may be NULL):
 * query b uses only its first nsq[b] <= nsplit splits (work-proportional splitting), the other slots are
 * written as empty partials (m = -inf). */
extern "C" int case_additive_attn_gate(const float* qa, const void* U, const float* G, const float* v, const uint8_t* mask,
                                       const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S,
                                       int nsplit, float* attn_un, float* stats, float* gate_part, int fast_tanh,
                                       const int32_t* cidx, const int32_t* ncount, const int32_t* qorder,
                                       const int32_t* nsq, case_stream_t stream) {
  using namespace cb;
  CB_REQUIRE(qa && U && G && v && mask && attn_un && stats && gate_part, "case_additive_attn_gate: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0 && nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "case_additive_attn_gate: bad sizes");
  CB_REQUIRE((cidx == nullptr) == (ncount == nullptr), "case_additive_attn_gate: cidx and ncount go together");
  CB_REQUIRE((uintptr_t)G % 16 == 0 && (uintptr_t)U % 16 == 0 && (uintptr_t)stats % 16 == 0 && (uintptr_t)gate_part % 16 == 0,
             "case_additive_attn_gate: U, G, stats, gate_part must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (W <= 1) return launch_gate<1>(fast_tanh, qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
  if (W <= 2) return launch_gate<2>(fast_tanh, qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
  if (W <= 4) return launch_gate<4>(fast_tanh, qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
  return launch_gate<8>(fast_tanh, qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
}

/* Prefill of the gate form: G fp32 [N][4] = (Wg[0] . mem[n], Wg[1] . mem[n], Wg[2] . mem[n], 0) for N key rows of
 * bf16 [N][H]; Wg fp32 [3][H] = W_m[:, H(1+i):H(2+i)] (CaSE/Model.py:36,39). */
extern "C" int case_gate_project(const void* mem, const float* Wg, float* G, long long N, case_stream_t stream) {
  using namespace cb;
  CB_REQUIRE(mem && Wg && G && N > 0, "case_gate_project: bad arguments");
  CB_REQUIRE((uintptr_t)mem % 16 == 0 && (uintptr_t)Wg % 16 == 0 && (uintptr_t)G % 16 == 0, "case_gate_project: 16-byte alignment required");
  const long long want = (N + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  launch_k(gate_project_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16*)mem, Wg, (float4*)G, N);
  return check_launch("case_gate_project");
}

/* Work-proportional split plan for case_additive_attn_gate: nsq[b] = clamp(ceil(count[b] / c), 1, max_split) with
 * c = max(ceil(sum(count) / slots), ceil(max(count) / max_split)) rounded up to a multiple of 32 keys, so the
 * launch has about `slots` equally long CTAs (+ at most one short CTA per query). */
extern "C" int case_split_plan(const int32_t* count, int B, int slots, int max_split, int32_t* nsq, case_stream_t stream) {
  using namespace cb;
  CB_REQUIRE(count && nsq && B > 0 && slots > 0 && max_split >= 1 && max_split <= CASE_MAX_SPLIT, "case_split_plan: bad arguments");
  launch_k(split_plan_kernel, 1, 256, 0, (cudaStream_t)stream, count, B, slots, max_split, nsq);
  return check_launch("case_split_plan");
}

/* case_additive_attn_gate with Uk.mem stored in f16 (U f16 [B][S][H]) for 2 <= W <= 8: packed f16 additions and
 * tanh.approx.f16x2, the v-weighted sum over the hidden units on the tensor core (mma.m16n8k16, fp32
 * accumulation).  Always the approximate tanh.  Everything else as case_additive_attn_gate. */
extern "C" int case_additive_attn_gate_h(const float* qa, const void* U, const float* G, const float* v, const uint8_t* mask,
                                         const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S,
                                         int nsplit, float* attn_un, float* stats, float* gate_part, const int32_t* cidx,
                                         const int32_t* ncount, const int32_t* qorder, const int32_t* nsq,
                                         case_stream_t stream) {
  using namespace cb;
  CB_REQUIRE(qa && U && G && v && mask && attn_un && stats && gate_part, "case_additive_attn_gate_h: null pointer");
  CB_REQUIRE(B > 0 && W >= 2 && W <= CASE_MAX_W && S > 0 && nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "case_additive_attn_gate_h: bad sizes (2 <= W <= 8)");
  CB_REQUIRE((cidx == nullptr) == (ncount == nullptr), "case_additive_attn_gate_h: cidx and ncount go together");
  CB_REQUIRE((uintptr_t)G % 16 == 0 && (uintptr_t)U % 16 == 0 && (uintptr_t)stats % 16 == 0 && (uintptr_t)gate_part % 16 == 0 && (uintptr_t)qa % 16 == 0,
             "case_additive_attn_gate_h: qa, U, G, stats, gate_part must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (W <= 2) return launch_gate_h<2>(qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
  if (W <= 4) return launch_gate_h<4>(qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
  return launch_gate_h<8>(qa, U, G, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, attn_un, stats, gate_part, cidx, ncount, qorder, nsq, st);
}
