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
#include "common.cuh"

namespace cb {

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
    ensure_smem<additive_attn_gate_kernel<WMAX, FAST_, FULL_>>(64 * 1024);                                                \
    launch_k(additive_attn_gate_kernel<WMAX, FAST_, FULL_>, dim3(B, nsplit), A2T, smem, st, qa, (const bf16*)U,            \
             (const float4*)G, v, mask, prior, tok, tok_ld, t, W, S, nsplit, scores, stats, gate_part, cidx, ncount, qorder, nsq, launch_opts().evict_first); \
  } while (0)
  if (fast) { if (full) AG_LAUNCH(true, true); else AG_LAUNCH(true, false); }
  else { if (full) AG_LAUNCH(false, true); else AG_LAUNCH(false, false); }
#undef AG_LAUNCH
  return check_launch("case_additive_attn_gate");
}

template <int WMAX, int DV, bool FAST>
static int launch_v3(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                     const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int nsplit,
                     float* scores, float* stats, float* ctx_part, const int32_t* cidx, const int32_t* ncount,
                     const int32_t* qorder, cudaStream_t st) {
  constexpr int KPT = WMAX == 8 ? 2 : 4;
  const size_t smem = (size_t)8 * 2 * KPT * (H + DV) * 2;
  auto kern = additive_attn_v3_kernel<WMAX, DV, FAST>;
  ensure_smem<additive_attn_v3_kernel<WMAX, DV, FAST>>(160 * 1024);
  launch_k(kern, dim3(B, nsplit), A2T, smem, st, qa, (const bf16*)U, (const bf16*)Mv, v, mask, prior, tok, tok_ld, t, W, S,
                                           nsplit, scores, stats, ctx_part, cidx, ncount, qorder);
  return check_launch("case_additive_attn(bf16)");
}

template <int DV, bool FAST>
static int dispatch_v3(int W, const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                       const float* prior, const int32_t* tok, int tok_ld, int t, int B, int S, int nsplit,
                       float* scores, float* stats, float* ctx_part, const int32_t* cidx, const int32_t* ncount,
                       const int32_t* qorder, cudaStream_t st) {
#define A3_ARGS qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, nsplit, scores, stats, ctx_part, cidx, ncount, qorder, st
  if (W <= 1) return launch_v3<1, DV, FAST>(A3_ARGS);
  if (W <= 2) return launch_v3<2, DV, FAST>(A3_ARGS);
  if (W <= 4) return launch_v3<4, DV, FAST>(A3_ARGS);
  return launch_v3<8, DV, FAST>(A3_ARGS);
#undef A3_ARGS
}

}  // namespace cb

/* bf16 form of case_additive_attn (attention.cu dispatches here); cidx / ncount / qorder: the compaction tables of
 * case_additive_attn_compact, or all NULL to walk every position under the mask. */
int case_additive_attn_bf16(const float* qa, const void* U, const void* Mv, const float* v, const uint8_t* mask,
                            const float* prior, const int32_t* tok, int tok_ld, int t, int B, int W, int S, int DV,
                            int nsplit, float* scores, float* stats, float* ctx_part, int fast_tanh, const int32_t* cidx,
                            const int32_t* ncount, const int32_t* qorder, cudaStream_t st) {
  using namespace cb;
#define A3_ARGS W, qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, S, nsplit, scores, stats, ctx_part, cidx, ncount, qorder, st
  if (DV == 256) {
    if (fast_tanh) return dispatch_v3<256, true>(A3_ARGS);
    return dispatch_v3<256, false>(A3_ARGS);
  }
  if (fast_tanh) return dispatch_v3<512, true>(A3_ARGS);
  return dispatch_v3<512, false>(A3_ARGS);
#undef A3_ARGS
}

/* case_additive_attn over the VALID keys only (bf16, warp-autonomous kernel): cidx int32 [B][S] positions
 * of the valid keys (ascending), ncount int32 [B], qorder int32 [B] launch order of the queries (heaviest
 * first; may be NULL).  scores[r][s] of padding positions are NOT written: fill them with -inf once. */
extern "C" int case_additive_attn_compact(const float* qa, const void* U, const void* Mv, const float* v,
                                          const uint8_t* mask, const float* prior, const int32_t* tok, int tok_ld, int t,
                                          int B, int W, int S, int DV, int nsplit, float* attn_un, float* stats,
                                          float* ctx_part, int fast_tanh, const int32_t* cidx, const int32_t* ncount,
                                          const int32_t* qorder, case_stream_t stream) {
  using namespace cb;
  CB_REQUIRE(cidx && ncount, "case_additive_attn_compact: cidx / ncount missing");
  CB_REQUIRE(qa && U && Mv && v && mask && attn_un && stats && ctx_part, "case_additive_attn_compact: null pointer");
  CB_REQUIRE(B > 0 && W >= 1 && W <= CASE_MAX_W && S > 0 && nsplit >= 1 && nsplit <= CASE_MAX_SPLIT && (DV == 256 || DV == 512),
             "case_additive_attn_compact: bad sizes");
  return case_additive_attn_bf16(qa, U, Mv, v, mask, prior, tok, tok_ld, t, B, W, S, DV, nsplit, attn_un, stats, ctx_part,
                                 fast_tanh, cidx, ncount, qorder, (cudaStream_t)stream);
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
