// Pre-decode producers of CaSE: what CaSE.do_test runs before the decoder (CaSE/Model.py:313-331) -
// the shared TransformerSeqEncoder over query and passages (common/TransformerSeqEncoderDecoder.py:14-45,
// TransformerEncoder.py:54-77), Interaction (common/Interaction.py:15-76), the TransformerBlock stacks of passage selection
// and supporting-token identification (common/TransformerBlock.py:22-33, CaSE/Model.py:127-215) and the prior / answer
// representation of ResponseGeneration.action (CaSE/Model.py:230-245).
//
// Kernels here are everything that is not a plain GEMM:
//   enc_embed          E[tok] * sqrt(H) + pe[pos]                                     (fp32 rows)
//   ln_rows_wide       LayerNorm of [M][C] rows, C = 256 or 1280, optional second addend, bf16 + fp32 outputs
//   enc_attention      nn.MultiheadAttention self-attention with a key padding mask, FlashAttention-2 style on
//                      mma.sync.m16n8k16 (head dim 32 for the C = 256 layers, 160 for the 5H = 1280 blocks)
//   interaction        the dual attention of Interaction.forward WITHOUT its [B*NP, Lp, Lq, 3H] tensor: the score matrix
//                      U [Lp][Lq] of one (query, passage) pair lives in shared memory, both softmaxes are taken from it,
//                      the four attended tensors are accumulated in registers, the 5H-wide outputs are written once
//   rows_dot           scorer Linear(H, 1) over rows
//   prior_answer       prior = sigmoid(passage score) * sigmoid(token score), normalised per query, and the answer
//                      representation sum_s prior[s] * mem_p[s]
// The GEMMs go through case_gemm_rows_tc (gemm_rows.cu): tcgen05 / TMEM with bias / activation / residual / row-mask
// epilogues.  bf16 storage for GEMM operands, fp32 accumulation, statistics and residual streams.
#include "common.cuh"

namespace cb {

__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------------------------------ embedding
__global__ __launch_bounds__(256) void enc_embed_kernel(const float* __restrict__ E, const float* __restrict__ pe,
                                                        const int32_t* __restrict__ tok, long long M, int L, float scale,
                                                        float* __restrict__ x) {
  pdl_wait();
  const long long m = (long long)blockIdx.x * 4 + (threadIdx.x >> 6);
  if (m >= M) return;
  const int c = (threadIdx.x & 63) * 4, pos = (int)(m % L);
  const int id = tok[m];
  const float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)id * H + c));
  const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (size_t)pos * H + c));
  *reinterpret_cast<float4*>(x + (size_t)m * H + c) =
      make_float4(fmaf(e.x, scale, p.x), fmaf(e.y, scale, p.y), fmaf(e.z, scale, p.z), fmaf(e.w, scale, p.w));
}

// ------------------------------------------------------------------------------------------ LayerNorm
// one warp per row; C / 32 elements per lane held in registers, in 16-byte chunks interleaved over the lanes (chunk i of
// lane l = chunk 32 i + l of the row), so that every load / store instruction of the warp covers one contiguous run.
// x (and the optional addend) fp32 or bf16.
template <int C, bool XBF>
__global__ __launch_bounds__(256) void ln_rows_wide_kernel(const void* __restrict__ x_, const void* __restrict__ add_,
                                                           const float* __restrict__ g, const float* __restrict__ b,
                                                           bf16* __restrict__ y16, float* __restrict__ y32, long long M) {
  pdl_wait();
  constexpr int PER = C / 32;                      // 8 or 40 per lane
  constexpr int VEC = XBF ? 8 : 4;                 // elements per 16-byte chunk of the input
  constexpr int NCH = PER / VEC;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  float v[PER];
  auto load = [&](const void* p, float (&o)[PER]) {
    if (XBF) {
      const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p) + (size_t)m * C) + lane;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const uint4 w = r[32 * i];
        o[8 * i] = bf_lo(w.x); o[8 * i + 1] = bf_hi(w.x); o[8 * i + 2] = bf_lo(w.y); o[8 * i + 3] = bf_hi(w.y);
        o[8 * i + 4] = bf_lo(w.z); o[8 * i + 5] = bf_hi(w.z); o[8 * i + 6] = bf_lo(w.w); o[8 * i + 7] = bf_hi(w.w);
      }
    } else {
      const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + (size_t)m * C) + lane;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const float4 w = r[32 * i];
        o[4 * i] = w.x; o[4 * i + 1] = w.y; o[4 * i + 2] = w.z; o[4 * i + 3] = w.w;
      }
    }
  };
  load(x_, v);
  if (add_ != nullptr) {
    float a[PER];
    load(add_, a);
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] += a[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  s = warp_sum(s);
  const float mean = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  q = warp_sum(q);
  const float rstd = rsqrtf(q * (1.f / C) + LN_EPS);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c0 = (32 * i + lane) * VEC;          // first column of the chunk
#pragma unroll
    for (int e = 0; e < VEC; e += 4) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c0 + e)), bb = __ldg(reinterpret_cast<const float4*>(b + c0 + e));
      float* t = v + VEC * i + e;
      t[0] = (t[0] - mean) * rstd * gg.x + bb.x; t[1] = (t[1] - mean) * rstd * gg.y + bb.y;
      t[2] = (t[2] - mean) * rstd * gg.z + bb.z; t[3] = (t[3] - mean) * rstd * gg.w + bb.w;
    }
    if (y16 != nullptr) {
      bf16* o = y16 + (size_t)m * C + c0;
      if (VEC == 8) {
        *reinterpret_cast<uint4*>(o) = make_uint4(pk2(v[8 * i], v[8 * i + 1]), pk2(v[8 * i + 2], v[8 * i + 3]),
                                                  pk2(v[8 * i + 4], v[8 * i + 5]), pk2(v[8 * i + 6], v[8 * i + 7]));
      } else {
        *reinterpret_cast<uint2*>(o) = make_uint2(pk2(v[4 * i], v[4 * i + 1]), pk2(v[4 * i + 2], v[4 * i + 3]));
      }
    }
    if (y32 != nullptr) {
      float* o = y32 + (size_t)m * C + c0;
#pragma unroll
      for (int e = 0; e < VEC; e += 4)
        *reinterpret_cast<float4*>(o + e) = make_float4(v[VEC * i + e], v[VEC * i + e + 1], v[VEC * i + e + 2], v[VEC * i + e + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------ self-attention
__device__ __forceinline__ void pa_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pa_ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pa_ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// softmax(Q K^T / sqrt(HD) + key padding mask) V for one (64-query tile, head, sequence).  qkv: bf16 [M][3C] rows =
// tokens (Q | K | V column blocks, head h = columns h * HD ..), kmask uint8 [M] (1 = valid key), out bf16 [M][C].
// Warp w owns query rows 16 w .. 16 w + 15 of the tile; keys stream through a two-stage cp.async ring of 64-key tiles.
template <int HD_>
__global__ __launch_bounds__(128) void enc_attention_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ kmask,
                                                            int L, int C, float scale, bf16* __restrict__ out) {
  constexpr int LD = HD_ + 8;                       // padded row (conflict-free ldmatrix)
  constexpr int KS = HD_ / 16;                      // k-steps of Q K^T
  constexpr int ND = HD_ / 8;                       // n-tiles of the output
  constexpr int CH = HD_ / 8;                       // 16-byte chunks per row
  extern __shared__ __align__(128) unsigned char sm[];
  bf16* Qs = reinterpret_cast<bf16*>(sm);
  bf16* KV = Qs + 64 * LD;                          // two stages of (K tile, V tile), 64 x LD each
  uint8_t* msk = reinterpret_cast<uint8_t*>(KV + 4 * 64 * LD);     // [2][64]
  pdl_wait();
  const int qt = blockIdx.x, h = blockIdx.y, seq = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const size_t row0 = (size_t)seq * L;
  const int ld = 3 * C;
  // 16-byte cp.async with zero fill for the positions past the sequence
  auto cp16 = [](const bf16* dst, const bf16* src, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  };
  auto fetch_kv = [&](int kt) {                     // key tile kt -> stage kt & 1
    bf16* Kd = KV + (kt & 1) * 2 * 64 * LD;
    bf16* Vd = Kd + 64 * LD;
    for (int i = tid; i < 64 * CH; i += 128) {
      const int r = i / CH, c = i - r * CH, pos = kt * 64 + r;
      const bool ok = pos < L;
      const bf16* src = qkv + (row0 + (ok ? pos : 0)) * ld + C + h * HD_ + c * 8;
      cp16(Kd + r * LD + c * 8, src, ok);
      cp16(Vd + r * LD + c * 8, src + C, ok);
    }
    if (tid < 64) { const int pos = kt * 64 + tid; msk[(kt & 1) * 64 + tid] = pos < L ? kmask[row0 + pos] : 0; }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Q tile, then the first key tile; the loop keeps one tile in flight behind the one it computes on
  for (int i = tid; i < 64 * CH; i += 128) {
    const int r = i / CH, c = i - r * CH, pos = qt * 64 + r;
    const bool ok = pos < L;
    cp16(Qs + r * LD + c * 8, qkv + (row0 + (ok ? pos : 0)) * ld + h * HD_ + c * 8, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  fetch_kv(0);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    pa_ldsm4(qf[ks], smem_u32(Qs + (16 * warp + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8));
  float o[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int nkt = (L + 63) / 64;
  for (int kt = 0; kt < nkt; ++kt) {
    __syncthreads();                                // the stage the next fetch overwrites is consumed
    if (kt + 1 < nkt) {
      fetch_kv(kt + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const bf16* Ks = KV + (kt & 1) * 2 * 64 * LD;
    const bf16* Vs = Ks + 64 * LD;
    const uint8_t* ms = msk + (kt & 1) * 64;
    // ---- S = Q K^T for the 64 keys of the tile (8 n-tiles)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {              // n-tile pair (2 np, 2 np + 1)
        uint32_t b[4];
        // matrices: (keys 16 np .. +7, k lo), (same keys, k hi), (keys 16 np + 8 .., k lo), (.., k hi)
        pa_ldsm4(b, smem_u32(Ks + (16 * np + (lane & 7) + ((lane >> 4) << 3)) * LD + ks * 16 + ((lane >> 3) & 1) * 8));
        pa_mma(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b[0], b[1]);
        pa_mma(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b[2], b[3]);
      }
    }
    // ---- mask, online softmax (rows g and g + 8; a row's 64 scores sit in the 4 lanes of a quad)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const bool ok0 = ms[8 * nt + 2 * tq] != 0, ok1 = ms[8 * nt + 2 * tq + 1] != 0;
      s[nt][0] = ok0 ? s[nt][0] * scale : -INFINITY; s[nt][1] = ok1 ? s[nt][1] * scale : -INFINITY;
      s[nt][2] = ok0 ? s[nt][2] * scale : -INFINITY; s[nt][3] = ok1 ? s[nt][3] * scale : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = (m0 == -INFINITY) ? 0.f : fexp(m0 - mn0), c1 = (m1 == -INFINITY) ? 0.f : fexp(m1 - mn1);
    const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;   // all-masked tile: p = 0
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];                              // P as bf16 A fragments, k-step kk = keys 16 kk .. 16 kk + 15
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fexp(s[nt][0] - e0), p1 = fexp(s[nt][1] - e0), p2 = fexp(s[nt][2] - e1), p3 = fexp(s[nt][3] - e1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = pk2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pk2(p2, p3);
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = fmaf(l0, c0, rs0); l1 = fmaf(l1, c1, rs1);
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) { o[nd][0] *= c0; o[nd][1] *= c0; o[nd][2] *= c1; o[nd][3] *= c1; }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < ND / 2; ++dp) {         // output n-tile pair (dims 16 dp .. 16 dp + 15)
        uint32_t b[4];
        // trans: matrices (keys 16 kk .. +7, dims 16 dp ..), (keys +8.., same dims), (keys .., dims 16 dp + 8 ..), (keys + 8, ..)
        pa_ldsm4t(b, smem_u32(Vs + (16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + 16 * dp + (lane >> 4) * 8));
        pa_mma(o[2 * dp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b[0], b[1]);
        pa_mma(o[2 * dp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b[2], b[3]);
      }
    }
  }
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int p0 = qt * 64 + 16 * warp + g, p1 = p0 + 8;
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    if (p0 < L) *reinterpret_cast<uint32_t*>(out + (row0 + p0) * C + h * HD_ + 8 * nd + 2 * tq) = pk2(o[nd][0] * i0, o[nd][1] * i0);
    if (p1 < L) *reinterpret_cast<uint32_t*>(out + (row0 + p1) * C + h * HD_ + 8 * nd + 2 * tq) = pk2(o[nd][2] * i1, o[nd][3] * i1);
  }
}

// ------------------------------------------------------------------------------------------ Interaction
// One CTA per (query, passage) pair.  E_q fp32 [B][Lq][H] (one query sequence per query), E_p fp32 [B*NP][Lp][H].
//   U[i][j] = w1.E_q[j] + w2.E_p[i] + (w3 * E_p[i]).E_q[j], -inf outside the mask       (Interaction.py:36-44)
//   A = softmax_j U, B = softmax_i U (zero outside the mask)                            (:45-49)
//   A1 = A E_q, B1 = B^T E_p, A2 = A B1, B2 = B^T A1                                     (:51-55)
//   G_q_p[i] = [E_p, A1, A2, E_p*A1, E_p*A2] (bf16, zero for PAD rows)                   (:68, 74)
//   G_p_q[j] = [E_q, B1, B2, E_q*B1, E_q*B2] (fp32 per passage, zero for PAD columns; max over passages follows)
// Row-local products (A1, A2) are computed by the warp that owns row i; the reductions over i (B1, B2) by the warp that
// owns the column: warp w owns columns w, w + 8, ... and walks all rows with its 8-wide slice of the hidden dimension
// per lane.  A1 goes through a global scratch (fp32 [B*NP*Lp][H]) between the two.
constexpr int IT_LQ = 64;                           // largest Lq
constexpr int IT_EQLD = H + 8;                      // bf16 row stride of E_q / B1 in shared memory: 16-byte aligned rows, and
                                                    // lane j reading 16 bytes of row j is conflict-free (132 words: 8 lanes, 8 bank groups)
__global__ __launch_bounds__(256) void interaction_kernel(const float* __restrict__ Eq, const float* __restrict__ Ep,
                                                          const uint8_t* __restrict__ qmask, const uint8_t* __restrict__ pmask,
                                                          const float* __restrict__ w, int NP, int Lq, int Lp,
                                                          float* __restrict__ A1s, float* __restrict__ Gq,
                                                          bf16* __restrict__ Gp) {
  extern __shared__ __align__(128) unsigned char sm[];
  bf16* eq = reinterpret_cast<bf16*>(sm);                               // [Lq][IT_EQLD]
  bf16* b1 = eq + IT_LQ * IT_EQLD;                                      // [Lq][IT_EQLD]
  float* U = reinterpret_cast<float*>(b1 + IT_LQ * IT_EQLD);            // [Lp][Lq + 1]
  const int UL = Lq + 1;
  float* rowb = U + (((size_t)Lp * UL + 3) & ~(size_t)3);               // [8][H]  w3 * E_p[i] of the warp's current row (16-byte aligned)
  float* aj = rowb + 8 * H;                                             // [IT_LQ]
  float* cmax = aj + IT_LQ;                                             // [IT_LQ]
  float* csum = cmax + IT_LQ;                                           // [IT_LQ]
  float* rmax = csum + IT_LQ;                                           // [Lp]
  float* rsum = rmax + Lp;                                              // [Lp]
  int* vidx = reinterpret_cast<int*>(rsum + Lp);                        // [Lp] indices of the valid passage rows
  int* nvalid = vidx + Lp;
  pdl_wait();
  const int s = blockIdx.x, b = s / NP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Eqb = Eq + (size_t)b * Lq * H;
  const float* Epb = Ep + (size_t)s * Lp * H;
  const uint8_t* qm = qmask + (size_t)b * Lq;
  const uint8_t* pm = pmask + (size_t)s * Lp;
  // ---- P0: E_q -> shared (bf16), a_j = w1 . E_q[j]
  for (int j = warp; j < Lq; j += 8) {
    const float4 x0 = *reinterpret_cast<const float4*>(Eqb + (size_t)j * H + lane * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(Eqb + (size_t)j * H + lane * 8 + 4);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + lane * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + lane * 8 + 4));
    uint32_t* d = reinterpret_cast<uint32_t*>(eq + j * IT_EQLD + lane * 8);
    d[0] = pk2(x0.x, x0.y); d[1] = pk2(x0.z, x0.w); d[2] = pk2(x1.x, x1.y); d[3] = pk2(x1.z, x1.w);
    float a = x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
    a = warp_sum(a);
    if (lane == 0) aj[j] = a;
  }
  __syncthreads();
  // ---- P1: U and the row statistics
  for (int i = warp; i < Lp; i += 8) {
    if (pm[i] == 0) {
      for (int j = lane; j < Lq; j += 32) U[i * UL + j] = -INFINITY;
      if (lane == 0) { rmax[i] = -INFINITY; rsum[i] = 0.f; }
      continue;
    }
    const float4 x0 = *reinterpret_cast<const float4*>(Epb + (size_t)i * H + lane * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(Epb + (size_t)i * H + lane * 8 + 4);
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(w + H + lane * 8)), u1 = __ldg(reinterpret_cast<const float4*>(w + H + lane * 8 + 4));
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(w + 2 * H + lane * 8)), t1 = __ldg(reinterpret_cast<const float4*>(w + 2 * H + lane * 8 + 4));
    float bi = x0.x * u0.x + x0.y * u0.y + x0.z * u0.z + x0.w * u0.w + x1.x * u1.x + x1.y * u1.y + x1.z * u1.z + x1.w * u1.w;
    bi = warp_sum(bi);
    float* rb = rowb + warp * H + lane * 8;
    *reinterpret_cast<float4*>(rb) = make_float4(x0.x * t0.x, x0.y * t0.y, x0.z * t0.z, x0.w * t0.w);
    *reinterpret_cast<float4*>(rb + 4) = make_float4(x1.x * t1.x, x1.y * t1.y, x1.z * t1.z, x1.w * t1.w);
    __syncwarp();
    float uv[2];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int j = lane + 32 * c;
      uv[c] = -INFINITY;
      if (j < Lq && qm[j] != 0) {
        const uint4* er = reinterpret_cast<const uint4*>(eq + j * IT_EQLD);
        const float4* rr = reinterpret_cast<const float4*>(rowb + warp * H);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll 4
        for (int k = 0; k < H / 8; ++k) {
          const uint4 e = er[k];
          const float4 r0 = rr[2 * k], r1 = rr[2 * k + 1];
          d0 = fmaf(r0.x, bf_lo(e.x), d0); d1 = fmaf(r0.y, bf_hi(e.x), d1);
          d0 = fmaf(r0.z, bf_lo(e.y), d0); d1 = fmaf(r0.w, bf_hi(e.y), d1);
          d0 = fmaf(r1.x, bf_lo(e.z), d0); d1 = fmaf(r1.y, bf_hi(e.z), d1);
          d0 = fmaf(r1.z, bf_lo(e.w), d0); d1 = fmaf(r1.w, bf_hi(e.w), d1);
        }
        uv[c] = aj[j] + bi + d0 + d1;
      }
      if (j < Lq) U[i * UL + j] = uv[c];
      mx = fmaxf(mx, uv[c]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) se += (uv[c] == -INFINITY) ? 0.f : __expf(uv[c] - mx);
    se = warp_sum(se);
    if (lane == 0) { rmax[i] = mx; rsum[i] = se; }
    __syncwarp();
  }
  __syncthreads();
  // ---- P2: column statistics; list of the valid rows (warp 7: ballot compaction, order kept)
  if (warp == 7) {
    int n = 0;
    for (int i0 = 0; i0 < Lp; i0 += 32) {
      const int i = i0 + lane;
      const bool ok = i < Lp && pm[i] != 0;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) vidx[n + __popc(bal & ((1u << lane) - 1u))] = i;
      n += __popc(bal);
    }
    if (lane == 0) *nvalid = n;
  }
  if (tid < Lq) {
    float mx = -INFINITY;
    for (int i = 0; i < Lp; ++i) mx = fmaxf(mx, U[i * UL + tid]);
    float se = 0.f;
    if (mx > -INFINITY)
      for (int i = 0; i < Lp; ++i) { const float u = U[i * UL + tid]; se += (u == -INFINITY) ? 0.f : __expf(u - mx); }
    cmax[tid] = mx; csum[tid] = se;
  }
  __syncthreads();
  // row-local product with a [Lq][H] bf16 operand in shared memory: acc[8 dims of the lane] = sum_j A[i][j] * X[j]
  auto row_prod = [&](int i, const bf16* X, float (&acc)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const float mxr = rmax[i], inv = rsum[i] > 0.f ? 1.f / rsum[i] : 0.f;
    for (int j = 0; j < Lq; ++j) {
      const float u = U[i * UL + j];
      if (u == -INFINITY) continue;                                     // uniform over the warp
      const float a = __expf(u - mxr) * inv;
      const uint4 x = *reinterpret_cast<const uint4*>(X + j * IT_EQLD + lane * 8);
      acc[0] = fmaf(a, bf_lo(x.x), acc[0]); acc[1] = fmaf(a, bf_hi(x.x), acc[1]);
      acc[2] = fmaf(a, bf_lo(x.y), acc[2]); acc[3] = fmaf(a, bf_hi(x.y), acc[3]);
      acc[4] = fmaf(a, bf_lo(x.z), acc[4]); acc[5] = fmaf(a, bf_hi(x.z), acc[5]);
      acc[6] = fmaf(a, bf_lo(x.w), acc[6]); acc[7] = fmaf(a, bf_hi(x.w), acc[7]);
    }
  };
  // reduction over the rows for the warp's columns: acc[c][8 dims] = sum_i B[i][j_c] * X[i]  (X fp32 rows in global memory).
  // Only the valid rows are walked (vidx: their indices, built once), four at a time with all eight 16-byte loads of the
  // batch in flight together - the loop is bound by the latency of these loads otherwise.
  auto col_prod = [&](const float* X, float (&acc)[8][8]) {
    float cm[8], ci[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = warp + 8 * c;
      cm[c] = j < Lq ? cmax[j] : 0.f;
      ci[c] = (j < Lq && csum[j] > 0.f) ? 1.f / csum[j] : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
    }
    const int nv = *nvalid;
    for (int i0 = 0; i0 < nv; i0 += 4) {
      int ii[4];
      float4 x0[4], x1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ii[u] = vidx[min(i0 + u, nv - 1)];
        x0[u] = *reinterpret_cast<const float4*>(X + (size_t)ii[u] * H + lane * 8);
        x1[u] = *reinterpret_cast<const float4*>(X + (size_t)ii[u] * H + lane * 8 + 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u >= nv) break;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j = warp + 8 * c;
          if (j >= Lq) break;
          const float uu = U[ii[u] * UL + j];
          const float bw = (uu == -INFINITY) ? 0.f : __expf(uu - cm[c]) * ci[c];
          acc[c][0] = fmaf(bw, x0[u].x, acc[c][0]); acc[c][1] = fmaf(bw, x0[u].y, acc[c][1]);
          acc[c][2] = fmaf(bw, x0[u].z, acc[c][2]); acc[c][3] = fmaf(bw, x0[u].w, acc[c][3]);
          acc[c][4] = fmaf(bw, x1[u].x, acc[c][4]); acc[c][5] = fmaf(bw, x1[u].y, acc[c][5]);
          acc[c][6] = fmaf(bw, x1[u].z, acc[c][6]); acc[c][7] = fmaf(bw, x1[u].w, acc[c][7]);
        }
      }
    }
  };
  float* A1b = A1s + (size_t)s * Lp * H;
  float* Gqb = Gq + (size_t)s * Lq * 5 * H;
  // ---- P3a: A1 = A E_q (row-local) -> scratch
  for (int i = warp; i < Lp; i += 8) {
    if (pm[i] == 0) continue;
    float acc[8];
    row_prod(i, eq, acc);
    *reinterpret_cast<float4*>(A1b + (size_t)i * H + lane * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(A1b + (size_t)i * H + lane * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  // ---- P3b: B1 = B^T E_p for the warp's columns -> shared (bf16) + G_p_q slot 1
  {
    float acc[8][8];
    col_prod(Epb, acc);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = warp + 8 * c;
      if (j >= Lq) break;
      uint32_t* d = reinterpret_cast<uint32_t*>(b1 + j * IT_EQLD + lane * 8);
      d[0] = pk2(acc[c][0], acc[c][1]); d[1] = pk2(acc[c][2], acc[c][3]); d[2] = pk2(acc[c][4], acc[c][5]); d[3] = pk2(acc[c][6], acc[c][7]);
      float* o = Gqb + (size_t)j * 5 * H + H + lane * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
    }
  }
  __syncthreads();                 // A1 (global, written by this CTA) and B1 (shared) are complete
  // ---- P4a: A2 = A B1 (row-local) and the passage-side output rows
  for (int i = warp; i < Lp; i += 8) {
    bf16* gp = Gp + ((size_t)s * Lp + i) * 5 * H;
    if (pm[i] == 0) {
      for (int k = lane; k < 5 * H / 8; k += 32) reinterpret_cast<uint4*>(gp)[k] = make_uint4(0, 0, 0, 0);
      continue;
    }
    float a2[8];
    row_prod(i, b1, a2);
    const float4 e0 = *reinterpret_cast<const float4*>(Epb + (size_t)i * H + lane * 8);
    const float4 e1 = *reinterpret_cast<const float4*>(Epb + (size_t)i * H + lane * 8 + 4);
    const float4 p0 = *reinterpret_cast<const float4*>(A1b + (size_t)i * H + lane * 8);
    const float4 p1 = *reinterpret_cast<const float4*>(A1b + (size_t)i * H + lane * 8 + 4);
    const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    const float a1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    auto st = [&](int slot, const float (&v)[8]) {
      *reinterpret_cast<uint4*>(gp + slot * H + lane * 8) = make_uint4(pk2(v[0], v[1]), pk2(v[2], v[3]), pk2(v[4], v[5]), pk2(v[6], v[7]));
    };
    float m1[8], m2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { m1[k] = e[k] * a1[k]; m2[k] = e[k] * a2[k]; }
    st(0, e); st(1, a1); st(2, a2); st(3, m1); st(4, m2);
  }
  // ---- P4b: B2 = B^T A1 for the warp's columns and the query-side output rows (fp32, per passage)
  {
    float acc[8][8];
    col_prod(A1b, acc);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = warp + 8 * c;
      if (j >= Lq) break;
      float* o = Gqb + (size_t)j * 5 * H;
      const float4 q0 = *reinterpret_cast<const float4*>(Eqb + (size_t)j * H + lane * 8);
      const float4 q1 = *reinterpret_cast<const float4*>(Eqb + (size_t)j * H + lane * 8 + 4);
      const float4 c0 = *reinterpret_cast<const float4*>(o + H + lane * 8), c1 = *reinterpret_cast<const float4*>(o + H + lane * 8 + 4);
      const bool ok = qm[j] != 0;
      const float e[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      const float bb1[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      auto st = [&](int slot, const float (&v)[8]) {
        float* d = o + slot * H + lane * 8;
        *reinterpret_cast<float4*>(d) = ok ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(d + 4) = ok ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      float m1[8], m2[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { m1[k] = e[k] * bb1[k]; m2[k] = e[k] * acc[c][k]; }
      st(0, e); st(1, bb1); st(2, acc[c]); st(3, m1); st(4, m2);
    }
  }
}

// G_p_q[b][j][c] = max over the NP passages of Gq[b][p][j][c]   (Interaction.py:73-74) -> bf16 rows
__global__ __launch_bounds__(256) void interaction_qmax_kernel(const float* __restrict__ Gq, int NP, long long per, long long total,
                                                               bf16* __restrict__ out) {
  pdl_wait();
  const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) * 2;       // (b, j, c) flattened, two columns per thread
  if (idx >= total) return;
  const long long b = idx / per, r = idx - b * per;
  float2 m = *reinterpret_cast<const float2*>(Gq + (b * NP) * per + r);
  for (int p = 1; p < NP; ++p) {
    const float2 v = *reinterpret_cast<const float2*>(Gq + (b * NP + p) * per + r);
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y);
  }
  *reinterpret_cast<uint32_t*>(out + idx) = pk2(m.x, m.y);
}

// ------------------------------------------------------------------------------------------ scorers, prior, answer
// y[m] = w . x[m] + b over fp32 rows [M][H], rows sampled with a stride (stride = L, offset 0: the [CLS] rows)
__global__ __launch_bounds__(256) void rows_dot_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                       long long nrows, long long row_stride, float* __restrict__ y) {
  pdl_wait();
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= nrows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)r * row_stride * H + lane * 8;
  const float4 x0 = *reinterpret_cast<const float4*>(xr), x1 = *reinterpret_cast<const float4*>(xr + 4);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + lane * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + lane * 8 + 4));
  float d = x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
  d = warp_sum(d);
  if (lane == 0) y[r] = d + __ldg(b);
}

// per query: prior[s] = sigmoid(ps[p(s)]) * sigmoid(ts[s]) (0 at PAD), normalised by 1e-8 + sum; answer = sum prior * mem_p
// (CaSE/Model.py:239-243).  One CTA of 256 threads per query; thread = hidden column for the answer.
__global__ __launch_bounds__(256) void prior_answer_kernel(const float* __restrict__ pscore, const float* __restrict__ tscore,
                                                           const uint8_t* __restrict__ pmask, const float* __restrict__ memp,
                                                           int NP, int Lp, float* __restrict__ prior, float* __restrict__ answer) {
  extern __shared__ float pr[];                     // [NP * Lp]
  __shared__ float red[8];
  pdl_wait();
  const int b = blockIdx.x, S = NP * Lp, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float part = 0.f;
  for (int s = tid; s < S; s += 256) {
    const float ps = pscore[b * NP + s / Lp];
    // token_score was masked_fill(~mask, -1e6) and clamped (Model.py:205-206): sigmoid(-1e6) == 0 in fp32
    const float ts = pmask[(size_t)b * S + s] ? fminf(fmaxf(tscore[(size_t)b * S + s], -1e6f), 1e6f) : -1e6f;
    const float v = (1.f / (1.f + __expf(-ps))) * (ts < -80.f ? 0.f : 1.f / (1.f + __expf(-ts)));
    pr[s] = v;
    part += v;
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w2 = 0; w2 < 8; ++w2) tot += red[w2];
  const float inv = 1.f / (1e-8f + tot);
  for (int s = tid; s < S; s += 256) {
    const float v = pr[s] * inv;
    pr[s] = v;
    prior[(size_t)b * S + s] = v;
  }
  __syncthreads();
  float acc = 0.f;
  const float* mp = memp + (size_t)b * S * H + tid;
  for (int s = 0; s < S; ++s) {
    const float v = pr[s];
    if (v != 0.f) acc = fmaf(v, mp[(size_t)s * H], acc);
  }
  answer[(size_t)b * H + tid] = acc;
}

}  // namespace cb

using namespace cb;

extern "C" int case_enc_embed(const float* E, const float* pe, const int32_t* tok, long long M, int L, float scale, float* x,
                              case_stream_t stream) {
  CB_REQUIRE(E && pe && tok && x && M > 0 && L > 0, "case_enc_embed: bad arguments");
  launch_k(enc_embed_kernel, (unsigned)((M + 3) / 4), 256, 0, (cudaStream_t)stream, E, pe, tok, M, L, scale, x);
  return check_launch("case_enc_embed");
}

extern "C" int case_ln_rows_wide(const void* x, const void* add, int in_dtype, const float* g, const float* b, void* y16,
                                 float* y32, long long M, int C, case_stream_t stream) {
  CB_REQUIRE(x && g && b && (y16 || y32) && M > 0 && (C == 256 || C == 1280), "case_ln_rows_wide: C must be 256 or 1280");
  CB_REQUIRE(in_dtype == CASE_F32 || in_dtype == CASE_BF16, "case_ln_rows_wide: in_dtype");
  const unsigned grid = (unsigned)((M + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 256) {
    if (in_dtype == CASE_BF16) launch_k(ln_rows_wide_kernel<256, true>, grid, 256, 0, st, x, add, g, b, (bf16*)y16, y32, M);
    else launch_k(ln_rows_wide_kernel<256, false>, grid, 256, 0, st, x, add, g, b, (bf16*)y16, y32, M);
  } else {
    if (in_dtype == CASE_BF16) launch_k(ln_rows_wide_kernel<1280, true>, grid, 256, 0, st, x, add, g, b, (bf16*)y16, y32, M);
    else launch_k(ln_rows_wide_kernel<1280, false>, grid, 256, 0, st, x, add, g, b, (bf16*)y16, y32, M);
  }
  return check_launch("case_ln_rows_wide");
}

extern "C" int case_enc_attention(const void* qkv, const uint8_t* kmask, int nseq, int L, int C, int nhead, void* out,
                                  case_stream_t stream) {
  CB_REQUIRE(qkv && kmask && out && nseq > 0 && L > 0, "case_enc_attention: bad arguments");
  CB_REQUIRE((C == 256 || C == 1280) && nhead == 8, "case_enc_attention: 8 heads of 32 (C = 256) or 160 (C = 1280)");
  CB_REQUIRE(nseq <= 65535, "case_enc_attention: too many sequences for one launch");
  const int hd = C / nhead;
  const float scale = 1.f / sqrtf((float)hd);
  dim3 grid((L + 63) / 64, nhead, nseq);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)5 * 64 * (hd + 8) * 2 + 128;
  if (hd == 32) {
    launch_k(enc_attention_kernel<32>, grid, 128, smem, st, (const bf16*)qkv, kmask, L, C, scale, (bf16*)out);
  } else {
    ensure_smem<enc_attention_kernel<160>>((int)smem);
    launch_k(enc_attention_kernel<160>, grid, 128, smem, st, (const bf16*)qkv, kmask, L, C, scale, (bf16*)out);
  }
  return check_launch("case_enc_attention");
}

extern "C" size_t case_interaction_smem_bytes(int Lq, int Lp) {
  return (size_t)2 * IT_LQ * IT_EQLD * 2 + ((((size_t)Lp * (Lq + 1) + 3) & ~(size_t)3) + 8 * H + 3 * IT_LQ + 3 * (size_t)Lp + 4) * 4;
}

extern "C" int case_interaction(const float* Eq, const float* Ep, const uint8_t* qmask, const uint8_t* pmask, const float* w,
                                int B, int NP, int Lq, int Lp, float* A1_scratch, float* Gq_scratch, void* Gq_out, void* Gp_out,
                                case_stream_t stream) {
  CB_REQUIRE(Eq && Ep && qmask && pmask && w && A1_scratch && Gq_scratch && Gq_out && Gp_out, "case_interaction: null pointer");
  CB_REQUIRE(B > 0 && NP > 0 && Lq >= 1 && Lq <= IT_LQ && Lp >= 1, "case_interaction: Lq must be 1..64");
  const size_t smem = case_interaction_smem_bytes(Lq, Lp);
  CB_REQUIRE(smem <= 227 * 1024, "case_interaction: passage too long for the shared-memory score matrix");
  cudaStream_t st = (cudaStream_t)stream;
  ensure_smem<interaction_kernel>(227 * 1024);
  launch_k(interaction_kernel, B * NP, 256, smem, st, Eq, Ep, qmask, pmask, w, NP, Lq, Lp, A1_scratch, Gq_scratch, (bf16*)Gp_out);
  int rc = check_launch("case_interaction");
  if (rc) return rc;
  const long long per = (long long)Lq * 5 * H, total = (long long)B * per;
  launch_k(interaction_qmax_kernel, (unsigned)((total / 2 + 255) / 256), 256, 0, st, (const float*)Gq_scratch, NP, per, total, (bf16*)Gq_out);
  return check_launch("case_interaction(qmax)");
}

extern "C" int case_rows_dot(const float* x, const float* w, const float* b, long long nrows, long long row_stride, float* y,
                             case_stream_t stream) {
  CB_REQUIRE(x && w && b && y && nrows > 0 && row_stride >= 1, "case_rows_dot: bad arguments");
  launch_k(rows_dot_kernel, (unsigned)((nrows + 7) / 8), 256, 0, (cudaStream_t)stream, x, w, b, nrows, row_stride, y);
  return check_launch("case_rows_dot");
}

extern "C" int case_prior_answer(const float* pscore, const float* tscore, const uint8_t* pmask, const float* memp, int B, int NP,
                                 int Lp, float* prior, float* answer, case_stream_t stream) {
  CB_REQUIRE(pscore && tscore && pmask && memp && prior && answer && B > 0 && NP > 0 && Lp > 0, "case_prior_answer: bad arguments");
  const size_t smem = (size_t)NP * Lp * 4;
  CB_REQUIRE(smem <= 200 * 1024, "case_prior_answer: too many source positions");
  ensure_smem<prior_answer_kernel>(200 * 1024);
  launch_k(prior_answer_kernel, B, 256, smem, (cudaStream_t)stream, pscore, tscore, pmask, memp, NP, Lp, prior, answer);
  return check_launch("case_prior_answer");
}
