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
  float* msk = reinterpret_cast<float*>(KV + 4 * 64 * LD);         // [2][64] additive key bias: 0 (valid) or -inf (padding)
  pdl_wait();
  const int qt = blockIdx.x, h = blockIdx.y, seq = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const size_t row0 = (size_t)seq * L;
  const int ld = 3 * C;
  const float sc2 = scale * 1.4426950408889634f;     // softmax in the base-2 domain
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
    if (tid < 64) { const int pos = kt * 64 + tid; msk[(kt & 1) * 64 + tid] = (pos < L && kmask[row0 + pos] != 0) ? 0.f : -INFINITY; }
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
    const float* ms = msk + (kt & 1) * 64;
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
    // ---- mask, online softmax (rows g and g + 8; a row's 64 scores sit in the 4 lanes of a quad).  Scores are taken to
    // the base-2 domain with the padding mask in one FFMA each: t = s * (scale * log2 e) + bias_key (0 or -inf)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 kb = *reinterpret_cast<const float2*>(ms + 8 * nt + 2 * tq);
      s[nt][0] = fmaf(s[nt][0], sc2, kb.x); s[nt][1] = fmaf(s[nt][1], sc2, kb.y);
      s[nt][2] = fmaf(s[nt][2], sc2, kb.x); s[nt][3] = fmaf(s[nt][3], sc2, kb.y);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = (m0 == -INFINITY) ? 0.f : exp2f(m0 - mn0), c1 = (m1 == -INFINITY) ? 0.f : exp2f(m1 - mn1);
    const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;   // all-masked tile: p = 0
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];                              // P as bf16 A fragments, k-step kk = keys 16 kk .. 16 kk + 15
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - e0), p1 = exp2f(s[nt][1] - e0), p2 = exp2f(s[nt][2] - e1), p3 = exp2f(s[nt][3] - e1);
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
// All five products run on mma.sync.m16n8k16 (bf16 operands, fp32 accumulation) over passage tiles of 64 rows; the score
// matrix is never stored - its 64 x 64 tile is recomputed from the operands where it is needed, directly in the fragment
// layout of the product that consumes it (U for the row softmax A, U^T for the column softmax B^T):
//   pass 1  B1: FlashAttention over the passage rows with the query positions as "queries" (online column max / sum,
//           rescaled accumulator); leaves the final column statistics for pass 2
//   pass 2  per tile: A (row softmax, complete within the tile's 64 columns) -> A1 = A E_q and A2 = A B1 -> the tile's
//           G_q_p rows; B^T from the final column statistics -> B2 += B^T A1
// Warp w: rows 16 (w & 3) .. + 15 of the 64-row operand, hidden dimensions 128 (w >> 2) .. + 127 of the products.
constexpr int IT_LQ = 64;                           // largest Lq
constexpr int IT_LD = H + 8;                        // bf16 row stride in shared memory (528 bytes: conflict-free ldmatrix)
constexpr int IT_BUF = IT_LQ * IT_LD;               // elements of one [64][IT_LD] operand buffer
constexpr int IT_SMEM = 6 * IT_BUF * 2 + 4 * 64 * 4 + 128;

__global__ __launch_bounds__(256, 1) void interaction_kernel(const float* __restrict__ Eq, const float* __restrict__ Ep,
                                                             const uint8_t* __restrict__ qmask, const uint8_t* __restrict__ pmask,
                                                             const float* __restrict__ w, int NP, int Lq, int Lp,
                                                             float* __restrict__ Gq, bf16* __restrict__ Gp) {
  extern __shared__ __align__(128) unsigned char sm[];
  bf16* EqS = reinterpret_cast<bf16*>(sm);          // E_q
  bf16* EqW = EqS + IT_BUF;                         // E_q * w3
  bf16* EpT = EqW + IT_BUF;                         // the current tile of E_p
  bf16* B1S = EpT + IT_BUF;
  bf16* A1T = B1S + IT_BUF;                         // A1 / A2 rows of the current tile
  bf16* A2T = A1T + IT_BUF;
  float* c1 = reinterpret_cast<float*>(A2T + IT_BUF);   // [64] w1 . E_q[j]
  float* r2 = c1 + 64;                              // [64] w2 . E_p[i] of the tile
  float* cmax = r2 + 64;                            // [64] column max / 1 / column sum (final, after pass 1)
  float* cinv = cmax + 64;
  uint8_t* qm = reinterpret_cast<uint8_t*>(cinv + 64);  // [64]
  uint8_t* pmt = qm + 64;                           // [64] mask of the tile's rows
  pdl_wait();
  const int pair = blockIdx.x, b = pair / NP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int mt = warp & 3, dh = warp >> 2;
  const float* eqg = Eq + (size_t)b * Lq * H;
  const float* epg = Ep + (size_t)pair * Lp * H;
  const uint8_t* pmg = pmask + (size_t)pair * Lp;
  auto pack8 = [](const float4& x0, const float4& x1) {
    return make_uint4(pk2(x0.x, x0.y), pk2(x0.z, x0.w), pk2(x1.x, x1.y), pk2(x1.z, x1.w));
  };
  auto dot8 = [](const float4& x0, const float4& x1, const float4& w0, const float4& w1) {
    return x0.x * w0.x + x0.y * w0.y + x0.z * w0.z + x0.w * w0.w + x1.x * w1.x + x1.y * w1.y + x1.z * w1.z + x1.w * w1.w;
  };
  // ---- E_q -> bf16 (plain and scaled by w3), c1[j] = w1 . E_q[j]; rows >= Lq are zero
  {
    const float4 wa0 = __ldg(reinterpret_cast<const float4*>(w + lane * 8)), wa1 = __ldg(reinterpret_cast<const float4*>(w + lane * 8 + 4));
    const float4 wc0 = __ldg(reinterpret_cast<const float4*>(w + 2 * H + lane * 8)),
                 wc1 = __ldg(reinterpret_cast<const float4*>(w + 2 * H + lane * 8 + 4));
    for (int j = warp; j < IT_LQ; j += 8) {
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (j < Lq) {
        x0 = *reinterpret_cast<const float4*>(eqg + (size_t)j * H + lane * 8);
        x1 = *reinterpret_cast<const float4*>(eqg + (size_t)j * H + lane * 8 + 4);
      }
      const float d = warp_sum(dot8(x0, x1, wa0, wa1));
      *reinterpret_cast<uint4*>(EqS + j * IT_LD + lane * 8) = pack8(x0, x1);
      const float4 y0 = make_float4(x0.x * wc0.x, x0.y * wc0.y, x0.z * wc0.z, x0.w * wc0.w),
                   y1 = make_float4(x1.x * wc1.x, x1.y * wc1.y, x1.z * wc1.z, x1.w * wc1.w);
      *reinterpret_cast<uint4*>(EqW + j * IT_LD + lane * 8) = pack8(y0, y1);
      if (lane == 0) { c1[j] = d; qm[j] = j < Lq ? qmask[(size_t)b * Lq + j] : 0; }
    }
  }
  const float4 wb0 = __ldg(reinterpret_cast<const float4*>(w + H + lane * 8)), wb1 = __ldg(reinterpret_cast<const float4*>(w + H + lane * 8 + 4));
  // tile t of E_p -> bf16 rows, r2[i] = w2 . E_p[i], row masks (8 rows per warp, all 16 loads of a lane in flight)
  auto load_tile = [&](int t) {
    float4 x0[8], x1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = t * 64 + warp * 8 + k;
      x0[k] = make_float4(0.f, 0.f, 0.f, 0.f); x1[k] = x0[k];
      if (i < Lp) {
        x0[k] = *reinterpret_cast<const float4*>(epg + (size_t)i * H + lane * 8);
        x1[k] = *reinterpret_cast<const float4*>(epg + (size_t)i * H + lane * 8 + 4);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = warp * 8 + k, i = t * 64 + r;
      const float d = warp_sum(dot8(x0[k], x1[k], wb0, wb1));
      *reinterpret_cast<uint4*>(EpT + r * IT_LD + lane * 8) = pack8(x0[k], x1[k]);
      if (lane == 0) { r2[r] = d; pmt[r] = i < Lp ? pmg[i] : 0; }
    }
  };
  // 64 x 64 score tile in accumulator layout: rows (16 mt + g, + 8) of the operand `rows`, columns 8 nt + 2 tq (+ 1) = rows
  // of the operand `cols`; + the two rank-1 terms; -inf outside the mask
  auto score_tile = [&](const bf16* rows, const bf16* cols, const float* rterm, const float* cterm, const uint8_t* rmask,
                        const uint8_t* cmask, float (&s)[8][4]) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll 4
    for (int ks = 0; ks < H / 16; ++ks) {
      uint32_t af[4];
      pa_ldsm4(af, smem_u32(rows + (16 * mt + (lane & 15)) * IT_LD + ks * 16 + (lane >> 4) * 8));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bb[4];
        pa_ldsm4(bb, smem_u32(cols + (16 * np + (lane & 7) + ((lane >> 4) << 3)) * IT_LD + ks * 16 + ((lane >> 3) & 1) * 8));
        pa_mma(s[2 * np], af[0], af[1], af[2], af[3], bb[0], bb[1]);
        pa_mma(s[2 * np + 1], af[0], af[1], af[2], af[3], bb[2], bb[3]);
      }
    }
    const int ra = 16 * mt + g, rb = ra + 8;
    const float ta = rterm[ra], tb = rterm[rb];
    const bool oka = rmask[ra] != 0, okb = rmask[rb] != 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int ca = 8 * nt + 2 * tq;
      const float u0 = cterm[ca], u1 = cterm[ca + 1];
      const bool k0 = cmask[ca] != 0, k1 = cmask[ca + 1] != 0;
      s[nt][0] = (oka && k0) ? s[nt][0] + ta + u0 : -INFINITY; s[nt][1] = (oka && k1) ? s[nt][1] + ta + u1 : -INFINITY;
      s[nt][2] = (okb && k0) ? s[nt][2] + tb + u0 : -INFINITY; s[nt][3] = (okb && k1) ? s[nt][3] + tb + u1 : -INFINITY;
    }
  };
  // acc[16][4] (rows 16 mt + g / + 8, dims 128 dh + 8 nd + 2 tq / + 1) += P (A fragments over the 64 k rows) . X[k][dims]
  auto prod_tile = [&](const uint32_t (&pf)[4][4], const bf16* X, float (&acc)[16][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {
        uint32_t bb[4];
        pa_ldsm4t(bb, smem_u32(X + (16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * IT_LD + 128 * dh + 16 * dp + (lane >> 4) * 8));
        pa_mma(acc[2 * dp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], bb[0], bb[1]);
        pa_mma(acc[2 * dp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], bb[2], bb[3]);
      }
    }
  };
  auto quad_max = [](float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  };
  auto quad_sum = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
  };
  const int ntile = (Lp + 63) / 64;
  const int ja = 16 * mt + g, jb = ja + 8;           // the warp's query positions in the column-side products
  // G_p_q segments seg0 / seg1 of rows ja, jb: value and E_q * value (zeros for PAD columns)
  auto write_gq = [&](const float (&acc)[16][4], float sa, float sb, int seg0, int seg1) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = half ? jb : ja;
      if (j >= Lq) continue;
      const bool ok = qm[j] != 0;
      const float sc = half ? sb : sa;
      float* o = Gq + ((size_t)pair * Lq + j) * (5 * H);
#pragma unroll
      for (int nd = 0; nd < 16; ++nd) {
        const int col = 128 * dh + 8 * nd + 2 * tq;
        const float2 e = *reinterpret_cast<const float2*>(eqg + (size_t)j * H + col);
        const float v0 = ok ? acc[nd][2 * half] * sc : 0.f, v1 = ok ? acc[nd][2 * half + 1] * sc : 0.f;
        *reinterpret_cast<float2*>(o + seg0 * H + col) = make_float2(v0, v1);
        *reinterpret_cast<float2*>(o + seg1 * H + col) = make_float2(ok ? e.x * v0 : 0.f, ok ? e.y * v1 : 0.f);
        if (seg0 == 1) *reinterpret_cast<float2*>(o + col) = ok ? e : make_float2(0.f, 0.f);
      }
    }
  };

  // =============================== pass 1: B1 = B^T E_p with an online column softmax
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  {
    float b1[16][4];
#pragma unroll
    for (int nd = 0; nd < 16; ++nd) { b1[nd][0] = b1[nd][1] = b1[nd][2] = b1[nd][3] = 0.f; }
    for (int t = 0; t < ntile; ++t) {
      __syncthreads();                               // the previous tile is consumed (t = 0: nothing yet)
      load_tile(t);
      __syncthreads();
      float s[8][4];
      score_tile(EqW, EpT, c1, r2, qm, pmt, s);      // U^T: rows j, columns i
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float c0 = (m0 == -INFINITY) ? 0.f : fexp(m0 - mn0), cc1 = (m1 == -INFINITY) ? 0.f : fexp(m1 - mn1);
      const float e0 = (mn0 == -INFINITY) ? 0.f : mn0, e1 = (mn1 == -INFINITY) ? 0.f : mn1;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = fexp(s[nt][0] - e0), p1 = fexp(s[nt][1] - e0), p2 = fexp(s[nt][2] - e1), p3 = fexp(s[nt][3] - e1);
        rs0 += p0 + p1; rs1 += p2 + p3;
        pf[nt >> 1][(nt & 1) * 2] = pk2(p0, p1);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pk2(p2, p3);
      }
      rs0 = quad_sum(rs0); rs1 = quad_sum(rs1);
      l0 = fmaf(l0, c0, rs0); l1 = fmaf(l1, cc1, rs1);
      m0 = mn0; m1 = mn1;
#pragma unroll
      for (int nd = 0; nd < 16; ++nd) { b1[nd][0] *= c0; b1[nd][1] *= c0; b1[nd][2] *= cc1; b1[nd][3] *= cc1; }
      prod_tile(pf, EpT, b1);
    }
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int nd = 0; nd < 16; ++nd) {
      const int col = 128 * dh + 8 * nd + 2 * tq;
      *reinterpret_cast<uint32_t*>(B1S + ja * IT_LD + col) = pk2(b1[nd][0] * i0, b1[nd][1] * i0);
      *reinterpret_cast<uint32_t*>(B1S + jb * IT_LD + col) = pk2(b1[nd][2] * i1, b1[nd][3] * i1);
    }
    if (dh == 0 && tq == 0) { cmax[ja] = m0; cinv[ja] = i0; cmax[jb] = m1; cinv[jb] = i1; }
    write_gq(b1, i0, i1, 1, 3);
  }

  // =============================== pass 2: A1, A2 and the G_q_p rows per tile; B2 = B^T A1
  float b2[16][4];
#pragma unroll
  for (int nd = 0; nd < 16; ++nd) { b2[nd][0] = b2[nd][1] = b2[nd][2] = b2[nd][3] = 0.f; }
  const float e0 = (m0 == -INFINITY) ? 0.f : m0, e1 = (m1 == -INFINITY) ? 0.f : m1;
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  for (int t = 0; t < ntile; ++t) {
    __syncthreads();                                 // the previous tile's buffers are consumed; B1S is complete
    load_tile(t);
    __syncthreads();
    {
      float s[8][4];
      score_tile(EpT, EqW, r2, c1, pmt, qm, s);      // U: rows i, columns j (the whole row: Lq <= 64)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      const float f0 = (mx0 == -INFINITY) ? 0.f : mx0, f1 = (mx1 == -INFINITY) ? 0.f : mx1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = fexp(s[nt][0] - f0); s[nt][1] = fexp(s[nt][1] - f0); s[nt][2] = fexp(s[nt][2] - f1); s[nt][3] = fexp(s[nt][3] - f1);
        rs0 += s[nt][0] + s[nt][1]; rs1 += s[nt][2] + s[nt][3];
      }
      rs0 = quad_sum(rs0); rs1 = quad_sum(rs1);
      const float n0 = rs0 > 0.f ? 1.f / rs0 : 0.f, n1 = rs1 > 0.f ? 1.f / rs1 : 0.f;
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        pf[nt >> 1][(nt & 1) * 2] = pk2(s[nt][0] * n0, s[nt][1] * n0);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pk2(s[nt][2] * n1, s[nt][3] * n1);
      }
      const int ra = 16 * mt + g, rb = ra + 8;
#pragma unroll
      for (int which = 0; which < 2; ++which) {      // A1 = A E_q, A2 = A B1
        float a[16][4];
#pragma unroll
        for (int nd = 0; nd < 16; ++nd) { a[nd][0] = a[nd][1] = a[nd][2] = a[nd][3] = 0.f; }
        prod_tile(pf, which ? B1S : EqS, a);
        bf16* dst = which ? A2T : A1T;
#pragma unroll
        for (int nd = 0; nd < 16; ++nd) {
          const int col = 128 * dh + 8 * nd + 2 * tq;
          *reinterpret_cast<uint32_t*>(dst + ra * IT_LD + col) = pk2(a[nd][0], a[nd][1]);
          *reinterpret_cast<uint32_t*>(dst + rb * IT_LD + col) = pk2(a[nd][2], a[nd][3]);
        }
      }
    }
    __syncthreads();                                 // A1T / A2T complete
    {
      float s[8][4];
      score_tile(EqW, EpT, c1, r2, qm, pmt, s);      // U^T again, now against the final column statistics
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        pf[nt >> 1][(nt & 1) * 2] = pk2(fexp(s[nt][0] - e0) * i0, fexp(s[nt][1] - e0) * i0);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pk2(fexp(s[nt][2] - e1) * i1, fexp(s[nt][3] - e1) * i1);
      }
      prod_tile(pf, A1T, b2);
    }
    // the tile's G_q_p rows: [E_p, A1, A2, E_p * A1, E_p * A2], 16 bytes per lane and segment
#pragma unroll 2
    for (int k = 0; k < 8; ++k) {
      const int r = warp * 8 + k, i = t * 64 + r;
      if (i >= Lp) break;
      uint4* o = reinterpret_cast<uint4*>(Gp + ((size_t)pair * Lp + i) * (5 * H) + lane * 8);
      if (pmt[r] == 0) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int sgi = 0; sgi < 5; ++sgi) o[sgi * (H / 8)] = z;
        continue;
      }
      const uint4 e = *reinterpret_cast<const uint4*>(EpT + r * IT_LD + lane * 8);
      const uint4 x1 = *reinterpret_cast<const uint4*>(A1T + r * IT_LD + lane * 8);
      const uint4 x2 = *reinterpret_cast<const uint4*>(A2T + r * IT_LD + lane * 8);
      auto mul2 = [](uint32_t p, uint32_t q) { return pk2(bf_lo(p) * bf_lo(q), bf_hi(p) * bf_hi(q)); };
      o[0] = e; o[H / 8] = x1; o[2 * (H / 8)] = x2;
      o[3 * (H / 8)] = make_uint4(mul2(e.x, x1.x), mul2(e.y, x1.y), mul2(e.z, x1.z), mul2(e.w, x1.w));
      o[4 * (H / 8)] = make_uint4(mul2(e.x, x2.x), mul2(e.y, x2.y), mul2(e.z, x2.z), mul2(e.w, x2.w));
    }
  }
  write_gq(b2, 1.f, 1.f, 2, 4);
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
  const size_t smem = (size_t)5 * 64 * (hd + 8) * 2 + 512;
  if (hd == 32) {
    launch_k(enc_attention_kernel<32>, grid, 128, smem, st, (const bf16*)qkv, kmask, L, C, scale, (bf16*)out);
  } else {
    ensure_smem<enc_attention_kernel<160>>((int)smem);
    launch_k(enc_attention_kernel<160>, grid, 128, smem, st, (const bf16*)qkv, kmask, L, C, scale, (bf16*)out);
  }
  return check_launch("case_enc_attention");
}

extern "C" size_t case_interaction_smem_bytes(int Lq, int Lp) {
  (void)Lq; (void)Lp;                              // six 64-row operand buffers, whatever the passage length
  return IT_SMEM;
}

extern "C" int case_interaction(const float* Eq, const float* Ep, const uint8_t* qmask, const uint8_t* pmask, const float* w,
                                int B, int NP, int Lq, int Lp, float* Gq_scratch, void* Gq_out, void* Gp_out,
                                case_stream_t stream) {
  CB_REQUIRE(Eq && Ep && qmask && pmask && w && Gq_scratch && Gq_out && Gp_out, "case_interaction: null pointer");
  CB_REQUIRE(B > 0 && NP > 0 && Lq >= 1 && Lq <= IT_LQ && Lp >= 1, "case_interaction: Lq must be 1..64");
  cudaStream_t st = (cudaStream_t)stream;
  ensure_smem<interaction_kernel>(IT_SMEM);
  launch_k(interaction_kernel, B * NP, 256, IT_SMEM, st, Eq, Ep, qmask, pmask, w, NP, Lq, Lp, Gq_scratch, (bf16*)Gp_out);
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
