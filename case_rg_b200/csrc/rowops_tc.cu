// Tensor-core forms of the row-local kernels for bf16 storage (rowops.cu keeps the fp32 CUDA-core
// forms).  A CTA owns RB = 8 decode rows, held as the first 8 rows of a 16-row bf16 A tile; every
// linear is D[16 x 256] += A[16 x K] . W^T on mma.sync.m16n8k16 with the eight consumer warps each
// owning 32 output columns.  Weights arrive as 16 KB slabs [256 n][32 k] (pre-swizzled in global
// memory so ldmatrix is bank-conflict free) through a ring of bulk async copies driven by four
// producer-only warps: full/empty mbarriers per stage, no block-wide barrier inside a linear.
//
// Packed weight layout (bf16), nn.Linear weight W[N][K]:
//   [N/256][K/32] slabs of 16 KB; inside a slab element (n, k) sits at byte
//   (n%256)*64 + ((((k%32)/8) ^ (((n%256)>>1)&3)) * 16) + (k%8)*2
#include "common.cuh"

namespace cb {

constexpr int TRB = 8;              // real rows per CTA (rows 8..15 of the MMA tile are zero)
constexpr int TCT = 256;            // consumer threads (8 warps)
constexpr int TPW = 4;              // producer warps: a warp has ONE bulk copy in flight at a time
                                    // (~0.36 us each, profiles/micro/bulk_bench.cu), so the ring is fed by four
constexpr int TNT = TCT + 32 * TPW;
constexpr int TSLAB = 16384;
constexpr int TNS = 6;              // ring stages
constexpr int ALD = H + 8;          // bf16 A-tile row stride for K = 256 (conflict-free ldmatrix)
constexpr int FLD = H + 4;          // fp32 row stride of the row buffers

__device__ __forceinline__ void tmb_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tmb_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmb_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tmb_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmb_wait(uint64_t* bar, uint32_t parity) {
  while (!tmb_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tbulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tmma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t tpack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct Ring {
  uint64_t* full;     // [TNS]
  uint64_t* empty;    // [TNS]
  char* stage;        // [TNS][TSLAB]
  int g;              // consumer: next slab (warp-uniform)
};

struct SlabSrc {      // up to 3 weight matrices consumed back to back
  const char* base[3];
  int end[3];
  int total;
  __device__ __forceinline__ const char* at(int p) const {
    int s = 0, first = 0;
    if (p >= end[0]) { s = 1; first = end[0]; }
    if (p >= end[1]) { s = 2; first = end[1]; }
    return base[s] + (size_t)(p - first) * TSLAB;
  }
};

__device__ __forceinline__ void ring_setup(Ring& rg, unsigned char* smem_raw) {
  rg.stage = reinterpret_cast<char*>(smem_raw);
  rg.full = reinterpret_cast<uint64_t*>(smem_raw + TNS * TSLAB);
  rg.empty = rg.full + TNS;
  rg.g = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TNS; ++s) { tmb_init(&rg.full[s], 1); tmb_init(&rg.empty[s], TCT / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();     // all TNT threads, once
}

// producer warps: warp i feeds slabs i, i + TPW, ... of the kernel, then exits
__device__ __forceinline__ void producer_run(const Ring& rg, const SlabSrc& src) {
  if ((threadIdx.x & 31) == 0) {
    for (int p = (threadIdx.x >> 5) - TCT / 32; p < src.total; p += TPW) {
      const int s = p % TNS;
      if (p >= TNS) tmb_wait(&rg.empty[s], (uint32_t)((p / TNS) - 1) & 1u);
      tmb_expect_tx(&rg.full[s], TSLAB);
      tbulk_g2s(rg.stage + (size_t)s * TSLAB, src.at(p), TSLAB, &rg.full[s]);
    }
  }
}

// D[16 x 256] = A[16 x K] . Wtile^T : warp w accumulates columns 32w .. 32w+31 (4 n-subtiles of 8)
// abuf: bf16 [16][lda] in shared memory; consumes K/32 slabs of the ring.
__device__ __forceinline__ void linear16(Ring& rg, const bf16* abuf, int lda, int K, float (&acc)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < 4; ++s) { acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f; }
  const uint32_t a_base = smem_u32(abuf) + (uint32_t)((lane & 15) * lda + (lane >> 4) * 8) * 2;
  const int brow = warp * 32 + (lane & 7), bch = lane >> 3;
  for (int kt = 0; kt < K / 32; ++kt) {
    const int g = rg.g, st = g % TNS;
    tmb_wait(&rg.full[st], (uint32_t)(g / TNS) & 1u);
    const uint32_t sb = smem_u32(rg.stage + (size_t)st * TSLAB);
    uint32_t a0[4], a1[4];
    tldsm4(a0, a_base + (uint32_t)(kt * 32) * 2);
    tldsm4(a1, a_base + (uint32_t)(kt * 32 + 16) * 2);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = brow + 8 * s;
      uint32_t b[4];
      tldsm4(b, sb + (uint32_t)(n * 64 + ((bch ^ ((n >> 1) & 3)) << 4)));
      tmma(acc[s], a0, b[0], b[1]);
      tmma(acc[s], a1, b[2], b[3]);
    }
    __syncwarp();
    if (lane == 0) tmb_arrive(&rg.empty[st]);
    rg.g = g + 1;
  }
}

// LayerNorm of row `warp` (8 consumer warps <-> 8 rows), fp32 in place, and its bf16 copy into the A tile
__device__ __forceinline__ void ln_row_warp(float* x /*[FLD]*/, const float* __restrict__ g,
                                            const float* __restrict__ b, bf16* arow /*[ALD]*/) {
  const int lane = threadIdx.x & 31;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = x[lane * 8 + i]; s += v[i]; }
  const float mean = warp_sum(s) * (1.f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / H) + LN_EPS);
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = lane * 8 + i;
    y[i] = (v[i] - mean) * rstd * __ldg(g + n) + __ldg(b + n);
    x[n] = y[i];
  }
  uint4 pk = make_uint4(tpack(y[0], y[1]), tpack(y[2], y[3]), tpack(y[4], y[5]), tpack(y[6], y[7]));
  *reinterpret_cast<uint4*>(arow + lane * 8) = pk;
}

// ------------------------------------------------------------------------------------------ generic
__global__ __launch_bounds__(TNT) void row_linear_tc_kernel(case_rowlin_args_t a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring rg;
  ring_setup(rg, smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int lda = a.K + 8;
  bf16* abuf = reinterpret_cast<bf16*>(smem_raw + TNS * TSLAB + 128);
  if (warp >= TCT / 32) {
    SlabSrc src;
    src.base[0] = src.base[1] = src.base[2] = reinterpret_cast<const char*>(a.Wt) + (size_t)blockIdx.y * (a.K / 32) * TSLAB;
    src.total = a.K / 32;
    src.end[0] = src.end[1] = src.end[2] = src.total;
    producer_run(rg, src);
    return;
  }
  pdl_trigger();
  pdl_wait();
  const int r0 = blockIdx.x * TRB;
  // gather + convert the input rows (one simple strided loop per segment so the loads pipeline);
  // rows >= TRB (and rows past R) are zero
  for (int i = tid; i < 8 * lda / 2; i += TCT) reinterpret_cast<uint32_t*>(abuf + 8 * lda)[i] = 0u;
  {
    int off = 0;
    for (int s = 0; s < a.nseg; ++s) {
      const case_seg_t sg = a.seg[s];
      const int hw = sg.width >> 1;
#pragma unroll 4
      for (int i = tid; i < TRB * hw; i += TCT) {
        const int rb = i / hw, c = (i - rb * hw) * 2, r = r0 + rb;
        float2 v = make_float2(0.f, 0.f);
        if (r < a.R) {
          int rr = sg.gather ? a.gather_idx[r] : r;
          rr /= sg.div;
          v = *reinterpret_cast<const float2*>(sg.p + (size_t)rr * sg.ld + c);
        }
        *reinterpret_cast<uint32_t*>(abuf + rb * lda + off + c) = tpack(v.x, v.y);
      }
      off += sg.width;
    }
  }
  consumer_sync();
  float acc[4][4];
  linear16(rg, abuf, lda, a.K, acc);
  const int g = lane >> 2, t = lane & 3, r = r0 + g;
  if (r < a.R) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = blockIdx.y * 256 + warp * 32 + s * 8 + 2 * t;
      float y0 = acc[s][0] + (a.bias ? __ldg(a.bias + n) : 0.f);
      float y1 = acc[s][1] + (a.bias ? __ldg(a.bias + n + 1) : 0.f);
      if (a.act == 1) { y0 = gelu_erf(y0); y1 = gelu_erf(y1); }
      if (a.res) { y0 += a.res[(size_t)r * a.ldres + n]; y1 += a.res[(size_t)r * a.ldres + n + 1]; }
      *reinterpret_cast<float2*>(a.out + (size_t)r * a.ldo + n) = make_float2(y0, y1);
    }
  }
}

// ------------------------------------------------------------------------------------------ layer front
// smem after the ring: [abuf0][abuf1] bf16 16 x ALD | xs, qs, hs fp32 8 x FLD | sc fp32 [8][8][CASE_MAX_T]
constexpr int T_ABUF_BYTES = 16 * ALD * 2;
constexpr int T_FRONT_SMEM =
    TNS * TSLAB + 128 + 2 * T_ABUF_BYTES + 3 * TRB * FLD * 4 + TRB * NH * CASE_MAX_T * 4 + TRB * CASE_MAX_T * 4;
constexpr int T_BACK_SMEM = TNS * TSLAB + 128 + 2 * T_ABUF_BYTES + 2 * TRB * FLD * 4;

__global__ __launch_bounds__(TNT) void layer_front_tc_kernel(const float* __restrict__ h, case_layer_weights_t w,
                                                             bf16* kc, bf16* vc, const int32_t* __restrict__ anc,
                                                             int anc_ld, const int32_t* __restrict__ tok, int tok_ld,
                                                             int t, int Tmax, float* __restrict__ b_out,
                                                             float* __restrict__ q2_out, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring rg;
  ring_setup(rg, smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (warp >= TCT / 32) {
    SlabSrc src;
    src.base[0] = reinterpret_cast<const char*>(w.Wqkv_t);
    src.base[1] = reinterpret_cast<const char*>(w.Wo_t);
    src.base[2] = reinterpret_cast<const char*>(w.Wq2_t);
    src.end[0] = 24; src.end[1] = 32; src.end[2] = 40;
    src.total = 40;
    producer_run(rg, src);
    return;
  }
  pdl_trigger();
  pdl_wait();
  unsigned char* p = smem_raw + TNS * TSLAB + 128;
  bf16* abuf0 = reinterpret_cast<bf16*>(p);
  bf16* abuf1 = reinterpret_cast<bf16*>(p + T_ABUF_BYTES);
  float* xs = reinterpret_cast<float*>(p + 2 * T_ABUF_BYTES);     // a = LN1(h)
  float* qs = xs + TRB * FLD;                                     // q (pre-scaled)
  float* hs = qs + TRB * FLD;                                     // h1, then b = LN2(h1)
  float* sc = hs + TRB * FLD;                                     // [row][head][key]
  const int r0 = blockIdx.x * TRB, g = lane >> 2, tq = lane & 3;

  // zero the padding rows of both A tiles, load the 8 input rows
  for (int i = tid; i < 8 * ALD / 2; i += TCT) {
    reinterpret_cast<uint32_t*>(abuf0 + 8 * ALD)[i] = 0u;
    reinterpret_cast<uint32_t*>(abuf1 + 8 * ALD)[i] = 0u;
  }
  {
    const int r = r0 + warp;
    float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
    if (r < R) {
      v0 = *reinterpret_cast<const float4*>(h + (size_t)r * H + lane * 8);
      v1 = *reinterpret_cast<const float4*>(h + (size_t)r * H + lane * 8 + 4);
    }
    *reinterpret_cast<float4*>(xs + warp * FLD + lane * 8) = v0;
    *reinterpret_cast<float4*>(xs + warp * FLD + lane * 8 + 4) = v1;
    __syncwarp();
    ln_row_warp(xs + warp * FLD, w.ln1_g, w.ln1_b, abuf0 + warp * ALD);
  }
  consumer_sync();

  float acc[4][4];
  const int rrow = r0 + g;
  for (int nt = 0; nt < 3; ++nt) {          // q | k | v
    linear16(rg, abuf0, ALD, H, acc);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = warp * 32 + s * 8 + 2 * tq;
      const float y0 = acc[s][0] + __ldg(w.bqkv + nt * 256 + n), y1 = acc[s][1] + __ldg(w.bqkv + nt * 256 + n + 1);
      if (nt == 0) {
        *reinterpret_cast<float2*>(qs + g * FLD + n) = make_float2(y0, y1);
      } else if (rrow < R) {
        bf16* dst = (nt == 1 ? kc : vc) + ((size_t)rrow * Tmax + t) * H + n;
        *reinterpret_cast<uint32_t*>(dst) = tpack(y0, y1);
      }
    }
  }
  consumer_sync();      // q in shared memory, newest K/V rows visible to the whole CTA

  // self-attention: warp <-> row; lane = (head, quarter of the head's 32 dims)
  {
    const int r = r0 + warp, hh = lane >> 2, qd = lane & 3;
    float* scr = sc + (size_t)warp * NH * CASE_MAX_T;
    float ctx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r < R) {
      float q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = qs[warp * FLD + hh * HD + qd * 8 + i];
      // physical row (bit 30 set = masked: PAD key, Model.py:106) of every history position
      int* prow = reinterpret_cast<int*>(sc + (size_t)TRB * NH * CASE_MAX_T) + warp * CASE_MAX_T;
      for (int j = lane; j <= t; j += 32) {
        const int pr = (j == t) ? r : anc[(size_t)r * anc_ld + j];
        prow[j] = pr | (tok[(size_t)pr * tok_ld + j] != 0 ? 0 : 0x40000000);
      }
      __syncwarp();
      // scores: 8 keys per batch, all loads of a batch issued before any reduction
      float mx = -INFINITY;
      const size_t coff = (size_t)hh * HD + qd * 8;
      for (int jb = 0; jb <= t; jb += 8) {
        uint4 raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(jb + u, t);
          raw[u] = *reinterpret_cast<const uint4*>(kc + ((size_t)(prow[j] & 0x3fffffff) * Tmax + j) * H + coff);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = jb + u;
          if (j <= t) {
            const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              d = fmaf(q[2 * i], __uint_as_float(w4[i] << 16), d);
              d = fmaf(q[2 * i + 1], __uint_as_float(w4[i] & 0xffff0000u), d);
            }
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            const float sv = (prow[j] & 0x40000000) ? -INFINITY : d;
            if (qd == 0) scr[hh * CASE_MAX_T + j] = sv;
            mx = fmaxf(mx, sv);
          }
        }
      }
      __syncwarp();
      float sum = 0.f;
      for (int jb = 0; jb <= t; jb += 8) {
        uint4 raw[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = min(jb + u, t);
          raw[u] = *reinterpret_cast<const uint4*>(vc + ((size_t)(prow[j] & 0x3fffffff) * Tmax + j) * H + coff);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = jb + u;
          if (j <= t) {
            const float sv = scr[hh * CASE_MAX_T + j];
            const float pexp = (sv == -INFINITY) ? 0.f : fexp(sv - mx);
            sum += pexp;
            const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ctx[2 * i] = fmaf(pexp, __uint_as_float(w4[i] << 16), ctx[2 * i]);
              ctx[2 * i + 1] = fmaf(pexp, __uint_as_float(w4[i] & 0xffff0000u), ctx[2 * i + 1]);
            }
          }
        }
      }
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ctx[i] *= inv;
    }
    uint4 pk = make_uint4(tpack(ctx[0], ctx[1]), tpack(ctx[2], ctx[3]), tpack(ctx[4], ctx[5]), tpack(ctx[6], ctx[7]));
    *reinterpret_cast<uint4*>(abuf1 + warp * ALD + hh * HD + qd * 8) = pk;
  }
  consumer_sync();

  // h1 = a + c.Wo + bo
  linear16(rg, abuf1, ALD, H, acc);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int n = warp * 32 + s * 8 + 2 * tq;
    const float2 a2 = *reinterpret_cast<const float2*>(xs + g * FLD + n);
    *reinterpret_cast<float2*>(hs + g * FLD + n) =
        make_float2(a2.x + acc[s][0] + __ldg(w.bo + n), a2.y + acc[s][1] + __ldg(w.bo + n + 1));
  }
  consumer_sync();
  ln_row_warp(hs + warp * FLD, w.ln2_g, w.ln2_b, abuf0 + warp * ALD);
  {
    const int r = r0 + warp;
    if (r < R) {
      __syncwarp();
      *reinterpret_cast<float4*>(b_out + (size_t)r * H + lane * 8) = *reinterpret_cast<const float4*>(hs + warp * FLD + lane * 8);
      *reinterpret_cast<float4*>(b_out + (size_t)r * H + lane * 8 + 4) = *reinterpret_cast<const float4*>(hs + warp * FLD + lane * 8 + 4);
    }
  }
  consumer_sync();
  linear16(rg, abuf0, ALD, H, acc);
  if (rrow < R) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = warp * 32 + s * 8 + 2 * tq;
      *reinterpret_cast<float2*>(q2_out + (size_t)rrow * H + n) =
          make_float2(acc[s][0] + __ldg(w.bq2 + n), acc[s][1] + __ldg(w.bq2 + n + 1));
    }
  }
}

// ------------------------------------------------------------------------------------------ layer back
__global__ __launch_bounds__(TNT) void layer_back_tc_kernel(const float* __restrict__ b_in,
                                                            const float* __restrict__ part_ml,
                                                            const float* __restrict__ part_acc, int nsplit,
                                                            case_layer_weights_t w, float* __restrict__ h_out, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring rg;
  ring_setup(rg, smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (warp >= TCT / 32) {
    SlabSrc src;
    src.base[0] = reinterpret_cast<const char*>(w.Wo2_t);
    src.base[1] = reinterpret_cast<const char*>(w.W1_t);
    src.base[2] = reinterpret_cast<const char*>(w.W2_t);
    src.end[0] = 8; src.end[1] = 16; src.end[2] = 24;
    src.total = 24;
    producer_run(rg, src);
    return;
  }
  pdl_trigger();
  pdl_wait();
  unsigned char* p = smem_raw + TNS * TSLAB + 128;
  bf16* abuf0 = reinterpret_cast<bf16*>(p);
  bf16* abuf1 = reinterpret_cast<bf16*>(p + T_ABUF_BYTES);
  float* bs = reinterpret_cast<float*>(p + 2 * T_ABUF_BYTES);    // b (residual of the cross-attention block)
  float* ys = bs + TRB * FLD;                                    // h2 -> c = LN3(h2)
  const int r0 = blockIdx.x * TRB, g = lane >> 2, tq = lane & 3, rrow = r0 + g;
  for (int i = tid; i < 8 * ALD / 2; i += TCT) {
    reinterpret_cast<uint32_t*>(abuf0 + 8 * ALD)[i] = 0u;
    reinterpret_cast<uint32_t*>(abuf1 + 8 * ALD)[i] = 0u;
  }
  // merge the cross-attention partials: thread = column, 8 rows.  All loads of a phase are
  // independent and issued together (this prologue is pure L2 latency otherwise).
  {
    const int hh = tid / HD, d = tid % HD;
    float* wgt = bs;                                   // [TRB][NH][nsplit] merge weights e_j / Z (bs | ys are free here)
    if (nsplit == 1) {
      if (tid < TRB * NH) {
        const int rb = tid / NH, r = r0 + rb;
        const float l = r < R ? part_ml[((size_t)r * NH + tid % NH) * 2 + 1] : 0.f;
        wgt[tid] = l > 0.f ? 1.f / l : 0.f;
      }
    } else if (nsplit > 16) {
      // many partials (compacted long memories: up to CASE_MAX_XSPLIT slots): the weights alone fill bs | ys
      // (64 x 64 floats), so the (m, l) pairs are read straight from global memory instead of being staged
      if (tid < TRB * NH) {
        const int r = r0 + tid / NH;
        const float2* mlp = reinterpret_cast<const float2*>(part_ml) + ((size_t)r * NH + tid % NH) * nsplit;
        float M = -INFINITY, Z = 0.f;
        if (r < R) {
          for (int j = 0; j < nsplit; ++j) M = fmaxf(M, mlp[j].x);
          for (int j = 0; j < nsplit; ++j) {
            const float2 ml = mlp[j];
            Z = fmaf(ml.y, (ml.x == -INFINITY) ? 0.f : fexp(ml.x - M), Z);
          }
        }
        for (int j = 0; j < nsplit; ++j) {
          const float mj = r < R ? mlp[j].x : -INFINITY;
          wgt[tid * nsplit + j] = (Z > 0.f && mj != -INFINITY) ? fexp(mj - M) / Z : 0.f;
        }
      }
    } else {
      float* mls = wgt + TRB * NH * nsplit;            // staged (m, l) pairs
      for (int i = tid; i < TRB * NH * nsplit; i += TCT) {
        const int rb = i / (NH * nsplit), r = r0 + rb;
        float2 ml = make_float2(-INFINITY, 0.f);
        if (r < R) ml = *reinterpret_cast<const float2*>(part_ml + ((size_t)r * NH * nsplit + (i - rb * NH * nsplit)) * 2);
        mls[2 * i] = ml.x; mls[2 * i + 1] = ml.y;
      }
      consumer_sync();
      if (tid < TRB * NH) {
        float M = -INFINITY, Z = 0.f;
        for (int j = 0; j < nsplit; ++j) M = fmaxf(M, mls[2 * (tid * nsplit + j)]);
        for (int j = 0; j < nsplit; ++j) {
          const float mj = mls[2 * (tid * nsplit + j)];
          Z = fmaf(mls[2 * (tid * nsplit + j) + 1], (mj == -INFINITY) ? 0.f : fexp(mj - M), Z);
        }
        for (int j = 0; j < nsplit; ++j) {
          const float mj = mls[2 * (tid * nsplit + j)];
          wgt[tid * nsplit + j] = (Z > 0.f && mj != -INFINITY) ? fexp(mj - M) / Z : 0.f;
        }
      }
    }
    float bv[TRB], cacc[TRB];
#pragma unroll
    for (int rb = 0; rb < TRB; ++rb) {
      const int r = r0 + rb;
      bv[rb] = r < R ? b_in[(size_t)r * H + tid] : 0.f;
      cacc[rb] = 0.f;
    }
    consumer_sync();
    for (int j = 0; j < nsplit; ++j) {
      float a[TRB];
#pragma unroll
      for (int rb = 0; rb < TRB; ++rb) {
        const int r = r0 + rb;
        a[rb] = r < R ? part_acc[(((size_t)r * NH + hh) * nsplit + j) * HD + d] : 0.f;
      }
#pragma unroll
      for (int rb = 0; rb < TRB; ++rb) cacc[rb] = fmaf(a[rb], wgt[(rb * NH + hh) * nsplit + j], cacc[rb]);
    }
    consumer_sync();                                   // wgt (aliasing bs | ys) is dead from here on
#pragma unroll
    for (int rb = 0; rb < TRB; ++rb) {
      bs[rb * FLD + tid] = bv[rb];
      abuf0[rb * ALD + tid] = __float2bfloat16_rn(cacc[rb]);
    }
  }
  consumer_sync();
  float acc[4][4];
  linear16(rg, abuf0, ALD, H, acc);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int n = warp * 32 + s * 8 + 2 * tq;
    const float2 b2 = *reinterpret_cast<const float2*>(bs + g * FLD + n);
    *reinterpret_cast<float2*>(ys + g * FLD + n) =
        make_float2(b2.x + acc[s][0] + __ldg(w.bo2 + n), b2.y + acc[s][1] + __ldg(w.bo2 + n + 1));
  }
  consumer_sync();
  ln_row_warp(ys + warp * FLD, w.ln3_g, w.ln3_b, abuf1 + warp * ALD);
  consumer_sync();
  linear16(rg, abuf1, ALD, H, acc);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int n = warp * 32 + s * 8 + 2 * tq;
    *reinterpret_cast<uint32_t*>(abuf0 + g * ALD + n) =
        tpack(gelu_erf(acc[s][0] + __ldg(w.b1 + n)), gelu_erf(acc[s][1] + __ldg(w.b1 + n + 1)));
  }
  consumer_sync();
  linear16(rg, abuf0, ALD, H, acc);
  if (rrow < R) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = warp * 32 + s * 8 + 2 * tq;
      const float2 c2 = *reinterpret_cast<const float2*>(ys + g * FLD + n);
      *reinterpret_cast<float2*>(h_out + (size_t)rrow * H + n) =
          make_float2(c2.x + acc[s][0] + __ldg(w.b2 + n), c2.y + acc[s][1] + __ldg(w.b2 + n + 1));
    }
  }
}

}  // namespace cb

using namespace cb;

int case_row_linear_tc(const case_rowlin_args_t* a, cudaStream_t st) {
  const size_t smem = (size_t)TNS * TSLAB + 128 + (size_t)16 * (a->K + 8) * 2;
  ensure_smem<row_linear_tc_kernel>(200 * 1024);
  dim3 grid((a->R + TRB - 1) / TRB, a->N / 256);
  launch_k(row_linear_tc_kernel, grid, TNT, smem, st, *a);
  return check_launch("case_row_linear(tc)");
}

int case_layer_front_tc(const float* h, const case_layer_weights_t* w, void* kcache, void* vcache, const int32_t* anc,
                        int anc_ld, const int32_t* tok, int tok_ld, int t, int Tmax, float* b_out, float* q2_out, int R,
                        cudaStream_t st) {
  ensure_smem<layer_front_tc_kernel>(T_FRONT_SMEM);
  launch_k(layer_front_tc_kernel, (R + TRB - 1) / TRB, TNT, T_FRONT_SMEM, st, h, *w, (bf16*)kcache, (bf16*)vcache, anc, anc_ld,
                                                                        tok, tok_ld, t, Tmax, b_out, q2_out, R);
  return check_launch("case_layer_front(tc)");
}

int case_layer_back_tc(const float* b_in, const float* part_ml, const float* part_acc, int nsplit,
                       const case_layer_weights_t* w, float* h_out, int R, cudaStream_t st) {
  CB_REQUIRE(nsplit <= CASE_MAX_XSPLIT, "case_layer_back: at most CASE_MAX_XSPLIT partials per (row, head)");
  static_assert(TRB * NH * CASE_MAX_XSPLIT * 4 <= 2 * TRB * FLD * 4, "merge weights must fit the row buffers");
  ensure_smem<layer_back_tc_kernel>(T_BACK_SMEM);
  launch_k(layer_back_tc_kernel, (R + TRB - 1) / TRB, TNT, T_BACK_SMEM, st, b_in, part_ml, part_acc, nsplit, *w, h_out, R);
  return check_launch("case_layer_back(tc)");
}
