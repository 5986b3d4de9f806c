// Vocabulary projection on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   logits[r][v] = sum_k f[r][k] * Wv[v][k] (+ bias[v])        r < R, v < V, K = 256
//
// The weight side is the big, static operand, so it is the UMMA "A" (M) operand: one CTA owns a
// tile of 128 vocabulary rows, the decode rows are the "B" (N) operand in blocks of up to 256, and
// the fp32 accumulator D[128 x N] lives in tensor memory.  Both operands are K-major and are stored
// in global memory already in the no-swizzle UMMA canonical layout (8x8 core matrices of 128 B,
// k-chunk major), so plain 1-D bulk async copies (cp.async.bulk -> UBLKCP) land them ready for the
// tensor core; no tensor maps are needed.  K is split into 4 chunks with one mbarrier each so the
// first MMAs start while later chunks are still in flight.  The epilogue reads TMEM with
// tcgen05.ld (32 lanes x 32 columns per warp) - lane = vocabulary row, so every store instruction
// writes 32 consecutive floats of one logits row.
//
// Packed layouts (bf16, element (row, k) of a tile with ROWS rows):
//   byte offset = (k / 8) * (ROWS / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2
//   weights:     [ceil(V/128)] tiles of ROWS = 128   (64 KB each; rows >= V are zero)
//   activations: [ceil(R/256)] blocks at a 128 KB stride, ROWS = the block's row count rounded up
//                to 16 (rows >= R are zero), so each 64-wide K chunk of a block is one contiguous run
#include "common.cuh"

namespace cb {

constexpr int TC_K = 256;
constexpr int TC_M = 128;                       // vocabulary rows per CTA
constexpr int TC_NB = 256;                      // decode rows per block
constexpr int TC_KCH = 4;                       // K chunks (pipeline stages)
constexpr int TC_W_BYTES = TC_M * TC_K * 2;     // 64 KB
constexpr int TC_A_BYTES = TC_NB * TC_K * 2;    // 128 KB
constexpr int TC_SMEM = TC_W_BYTES + TC_A_BYTES + 128;
constexpr int TC_THREADS = 256;                 // warps 0..3 feed the pipeline; all 8 drain TMEM (two per lane quarter)

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tc_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_expect_tx_only(uint32_t bar, uint32_t bytes) {   // no arrival
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool tc_try_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
  while (!tc_try_wait(bar, parity)) {}
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) leading-dim byte offset>>4 (between the two 8-element k-chunks of one
// K=16 MMA), [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 [R][256] -> bf16 canonical blocks of 256 rows (zero padded)
__global__ void pack_activations_kernel(const float* __restrict__ f, bf16* __restrict__ out, int R) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // one thread per (row, k-chunk of 8)
  const int nblk = (R + TC_NB - 1) / TC_NB;
  if (idx >= nblk * TC_NB * (TC_K / 8)) return;
  const int kc = idx % (TC_K / 8), row = idx / (TC_K / 8);
  const int blk = row / TC_NB, rr = row % TC_NB;
  const int npad = (min(TC_NB, R - blk * TC_NB) + 15) & ~15;     // rows of this block, padded to the MMA N granule
  if (rr >= npad) return;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < R) {
    const float4 a = *reinterpret_cast<const float4*>(f + (size_t)row * TC_K + kc * 8);
    const float4 b = *reinterpret_cast<const float4*>(f + (size_t)row * TC_K + kc * 8 + 4);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
    v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
  }
  const size_t off = (size_t)blk * TC_A_BYTES + (size_t)kc * (npad / 8) * 128 + (rr / 8) * 128 + (rr % 8) * 16;
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + off) = v;
}

__global__ __launch_bounds__(TC_THREADS, 1) void vocab_gemm_tc_kernel(const bf16* __restrict__ Wp, const bf16* __restrict__ Ap,
                                                               const float* __restrict__ bias,
                                                               float* __restrict__ logits, int R, int V, int ldl) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s_w = smem_u32(smem), s_a = s_w + TC_W_BYTES;
  const uint32_t s_bar = s_a + TC_A_BYTES;             // [0..3] chunk-full, [4] mma-done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TC_W_BYTES + TC_A_BYTES + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int nblk = (R + TC_NB - 1) / TC_NB;

  if (tid == 0) {
    for (int i = 0; i < TC_KCH + 1; ++i) tc_mbar_init(s_bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  constexpr uint32_t W_CH = TC_W_BYTES / TC_KCH;       // 16 KB of weights per K chunk
  // one issuing lane per K chunk, in four different warps: a warp keeps only one bulk copy in
  // flight at a time (profiles/micro/bulk_bench.cu), so spreading the issue overlaps the chunks.
  // The weight tile is a constant: it is requested before the dependency wait.
  if (lane == 0 && warp < TC_KCH) {
    const char* wsrc = reinterpret_cast<const char*>(Wp) + (size_t)tile * TC_W_BYTES;
    tc_expect_tx_only(s_bar + 8 * warp, W_CH);
    tc_bulk_g2s(s_w + warp * W_CH, wsrc + (size_t)warp * W_CH, W_CH, s_bar + 8 * warp);
  }
  pdl_trigger();
  pdl_wait();
  for (int blk = 0; blk < nblk; ++blk) {
    const int rows = min(TC_NB, R - blk * TC_NB);
    const int N = (rows + 15) & ~15;
    const uint32_t par = blk & 1;
    if (lane == 0 && warp < TC_KCH) {
      const int c = warp;                                    // one issuing warp per K chunk
      const char* asrc = reinterpret_cast<const char*>(Ap) + (size_t)blk * TC_A_BYTES;
      const uint32_t a_ch = (uint32_t)N * (TC_K / TC_KCH) * 2;     // N rows x 64 k of bf16, contiguous
      const uint32_t bar = s_bar + 8 * c;
      tc_expect_tx(bar, a_ch);                               // the (single) arrival of this phase
      tc_bulk_g2s(s_a + c * a_ch, asrc + (size_t)c * a_ch, a_ch, bar);
    }
    if (tid == 0) {
      const uint32_t idesc = umma_idesc(TC_M, N);
      for (int c = 0; c < TC_KCH; ++c) {
        tc_wait(s_bar + 8 * c, par);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                     // 4 MMAs of K = 16 per 64-wide chunk
          const int kc = c * 8 + ks * 2;                     // first 8-element k-chunk of this MMA
          const uint64_t ad = umma_desc(s_w + kc * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
          const uint64_t bd = umma_desc(s_a + kc * (N / 8) * 128, (N / 8) * 128, 128);
          umma_bf16(tmem, ad, bd, idesc, (c | ks) != 0);
        }
      }
      umma_commit(s_bar + 8 * TC_KCH);
    }
    // ---- epilogue: a warp reads the TMEM lanes of its quarter (warp % 4 = 32 vocabulary rows); the two warps
    // of a quarter take alternate 32-column groups, so twice as many stores are in flight
    tc_wait(s_bar + 8 * TC_KCH, par);
    __syncwarp();         // lane 0 of warp 0 arrives here from the issue path; tcgen05.ld is .aligned
    tc_fence_after();
    const int lq = warp & 3;
    const int v = tile * TC_M + lq * 32 + lane;
    const float bv = (bias && v < V) ? __ldg(bias + v) : 0.f;
    for (int c0 = (warp >> 2) * 32; c0 < N; c0 += 64) {
      uint32_t acc[32];
      tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + c0, acc);
      tmem_ld_wait();
      if (v < V) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int r = blk * TC_NB + c0 + j;
          if (r < R) logits[(size_t)r * ldl + v] = __uint_as_float(acc[j]) + bv;
        }
      }
    }
    tc_fence_before();
    __syncthreads();      // TMEM and the activation buffer are free for the next block
    tc_fence_after();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}


// ------------------------------------------------------------------------------------------ prefill projections
// Once per batch and memory: the cross-attention K|V of all layers of a stack (TransformerDecoder.py:81 recomputes
// them every step) and Uk.mem of the additive attention (BilinearAttention.py:34) are ONE GEMM over the memory rows,
//   Y[key][n] = sum_k mem[key][k] * W[n][k] (+ bias[n]),   n = (layer, K|V, head, dim) ++ (Uk column)
// whose epilogue writes the consumers' layouts directly: the swizzled bf16 K|V tiles of case_cross_attn_part /
// case_layer_stack (valid keys first when cidx / ncount are given - the compaction happens on the LOAD side, the
// A tiles are gathered) and U row-major at the keys' original positions.  No [B*S][2048] intermediate (687 MB at the
// BASELINE shape) is written or re-read.
//   CTA = PF_KEYS keys of one query (UMMA M = 128 tiles, TMEM lanes = keys) x a sequence of PF_N-column weight
//   blocks.  The A tile(s) stay in shared memory; weight blocks (canonical layout, L2-resident) stream through two
//   buffers; the accumulators are double buffered in TMEM: the MMAs of block j+1 run under the epilogue of block j.
//   Measured at the BASELINE shape (B = 64, S = 2560 + 60, ~1750 valid keys per query): 0.31-0.33 ms for both
//   memories against 0.20 (cuBLAS) + 0.28 (packing pass) before.  What bounds it (A/B with the MMAs and / or the
//   epilogue switched off: 121 us with neither, 210 us MMAs only, 257 us epilogue only): every CTA re-streams all
//   1.15 MB of weight blocks from L2 for its 128 keys, 148 of them together ask ~9 TB/s of L2 - the weight stream
//   alone is the 121 us - and the SS-mode MMAs (128 B/clk of shared-memory operand reads at M = N = 128) share the
//   shared-memory port with the landing copies.  256 keys x 64-column blocks (half the L2 stream) measured 380 us:
//   N = 64 MMAs need 192 B/clk.  Next step: a 4-CTA cluster whose CTAs each fetch a quarter of a weight block and
//   multicast it (quarter of the L2 stream), or the A tile in TMEM.
constexpr int PF_M = 128, PF_TILES = 1, PF_N = 128;
constexpr int PF_KEYS = PF_M * PF_TILES;             // keys per CTA
constexpr int PF_A_BYTES = PF_M * TC_K * 2;          // 64 KB per A tile
constexpr int PF_W_BYTES = PF_N * TC_K * 2;          // bytes per weight block
constexpr int PF_CG = PF_N / 32;                     // 32-column groups (= heads) per block
constexpr int PF_MAX_BIAS = 4 * 2 * TC_K;            // 4 layers x (K, V) x 256
constexpr int PF_EPI_WARPS = 16;                     // epilogue warps: one (tile, head) pair of a lane quarter each
constexpr int PF_THREADS = (PF_EPI_WARPS + 1) * 32;  // + warp 16: its lane 0 issues the MMAs and nothing else
constexpr int PF_SMEM = PF_TILES * PF_A_BYTES + 2 * PF_W_BYTES + PF_MAX_BIAS * 4 + PF_KEYS * 4 + 128;
static_assert(PF_TILES * PF_CG == 4, "the epilogue gives every warp one (tile, head) pair of its lane quarter");
struct PfOut { bf16* p[4]; };

__global__ __launch_bounds__(PF_THREADS, 1) void prefill_project_tc_kernel(
    const bf16* __restrict__ mem, const bf16* __restrict__ Wp, const float* __restrict__ bias, int S,
    const int32_t* __restrict__ cidx, const int32_t* __restrict__ ncount, int nl, PfOut out, bf16* __restrict__ U) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s_a = smem_u32(smem), s_w = s_a + PF_TILES * PF_A_BYTES;
  float* s_bias = reinterpret_cast<float*>(smem + PF_TILES * PF_A_BYTES + 2 * PF_W_BYTES);
  int* s_src = reinterpret_cast<int*>(s_bias + PF_MAX_BIAS);
  const uint32_t s_bar = smem_u32(s_src + PF_KEYS);        // [0,1] weights landed, [2,3] MMAs done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_src + PF_KEYS) + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, b = blockIdx.y;
  const int ntile = (S + 63) / 64;
  pdl_wait();
  const int nv = ncount ? ncount[b] : S;
  const int nkv = nl * 2 * TC_K / PF_N, nu = U ? TC_K / PF_N : 0;
  const bool kv_live = mt * PF_KEYS < nv || mt == 0;      // tiles past the query's last one are never read
  const int jb0 = kv_live ? 0 : nkv, nblk = nkv + nu - jb0;
  if (nblk <= 0) return;
  const int ntl = min(PF_TILES, (S - mt * PF_KEYS + PF_M - 1) / PF_M);   // A tiles wholly past the memory are skipped

  if (tid == 0) {
    tc_mbar_init(s_bar, 4); tc_mbar_init(s_bar + 8, 4);              // weights landed
    tc_mbar_init(s_bar + 16, 1); tc_mbar_init(s_bar + 24, 1);        // MMAs done (tcgen05.commit)
    tc_mbar_init(s_bar + 32, PF_EPI_WARPS); tc_mbar_init(s_bar + 40, PF_EPI_WARPS);   // accumulator drained
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < PF_KEYS) {
    const int sc = mt * PF_KEYS + tid;
    s_src[tid] = sc < S ? (cidx ? cidx[(size_t)b * S + sc] : sc) : -1;
  }
  for (int i = tid; i < nl * 2 * TC_K; i += PF_THREADS) s_bias[i] = __ldg(bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // weight block jj (local index) -> buffer jj & 1: four bulk copies issued from four warps
  auto issue_w = [&](int jj) {
    const uint32_t bar = s_bar + 8 * (jj & 1);
    const char* src = reinterpret_cast<const char*>(Wp) + (size_t)(jb0 + jj) * PF_W_BYTES + (size_t)warp * (PF_W_BYTES / 4);
    tc_expect_tx(bar, PF_W_BYTES / 4);
    tc_bulk_g2s(s_w + (jj & 1) * PF_W_BYTES + warp * (PF_W_BYTES / 4), src, PF_W_BYTES / 4, bar);
  };
  if (lane == 0 && warp < 4) {
    issue_w(0);
    if (nblk > 1) issue_w(1);
  }
  // A tiles: 128 gathered memory rows each -> canonical layout.  lane = (k-chunk & 1, row & 15): 32-byte runs of a
  // source row per 2 lanes, 2-way bank conflicts on the store side
  for (int tl = 0; tl < (warp < PF_EPI_WARPS ? ntl : 0); ++tl) {
    const int r16 = lane & 15, kc = warp * 2 + (lane >> 4);
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                        // warp: rows 16 i .. 16 i + 15, k-chunks 2*warp, 2*warp + 1
      const int s = s_src[tl * PF_M + i * 16 + r16];
      v[i] = make_uint4(0, 0, 0, 0);
      if (s >= 0) v[i] = __ldg(reinterpret_cast<const uint4*>(mem + ((size_t)b * S + s) * TC_K + kc * 8));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<uint4*>(smem + tl * PF_A_BYTES + kc * (PF_M / 8) * 128 + (i * 16 + r16) * 16) = v[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
  __syncthreads();

  constexpr uint32_t idesc = umma_idesc(PF_M, PF_N);
  auto issue_mma = [&](int jj) {
    const int buf = jj & 1;
    tc_wait(s_bar + 8 * buf, (jj >> 1) & 1);
    tc_fence_after();
    for (int tl = 0; tl < ntl; ++tl) {
#pragma unroll
      for (int ks = 0; ks < TC_K / 16; ++ks) {
        const int kc = ks * 2;
        const uint64_t ad = umma_desc(s_a + tl * PF_A_BYTES + kc * (PF_M / 8) * 128, (PF_M / 8) * 128, 128);
        const uint64_t bd = umma_desc(s_w + buf * PF_W_BYTES + kc * (PF_N / 8) * 128, (PF_N / 8) * 128, 128);
        umma_bf16(tmem + buf * (PF_TILES * PF_N) + tl * PF_N, ad, bd, idesc, ks != 0);
      }
    }
    umma_commit(s_bar + 16 + 8 * buf);
  };
  // warp 16: the MMA issuer.  tcgen05.mma issue blocks its thread for most of the MMAs' duration, so an issuer that
  // is also an epilogue warp serialises weights -> MMAs -> epilogue (measured: 2.1 us per block = the sum of the
  // three); a dedicated warp runs ahead, bounded only by landed weights and drained accumulators.
  if (warp == PF_EPI_WARPS) {
    if (lane == 0) {
      for (int jj = 0; jj < nblk; ++jj) {
        if (jj >= 2) tc_wait(s_bar + 32 + 8 * (jj & 1), ((jj >> 1) + 1) & 1);    // epilogue of block jj - 2 done
        issue_mma(jj);
      }
    }
  } else {
  // epilogue roles: warp & 3 = TMEM lane quarter (32 keys), warp >> 2 = one of the quarter's four (tile, head) pairs.
  // A thread drains the 32 columns of ITS key (tcgen05.ld: lane = key) = one head's K or V row = 64 contiguous bytes
  // of the tile stream (a transposition through shared memory to 512-byte store runs measured no faster).
  const int q = warp & 3, combo = warp >> 2, tl = combo / PF_CG, c0 = (combo % PF_CG) * 32;
  const int km = tl * PF_M + q * 32 + lane, sc = mt * PF_KEYS + km;
  const int tile = sc >> 6, key = sc & 63;
  const int s_orig = s_src[km];
  const bool kv_write = tile < ntile && (tile * 64 < nv || tile == 0);
  const bool kv_valid = sc < nv;

  for (int jj = 0; jj < nblk; ++jj) {
    const int buf = jj & 1, j = jb0 + jj;
    tc_wait(s_bar + 16 + 8 * buf, (jj >> 1) & 1);
    tc_fence_after();
    if (lane == 0 && warp < 4 && jj + 2 < nblk) issue_w(jj + 2);   // MMAs of block jj are done: its buffer is free
    __syncwarp();
    if (tl < ntl) {
      const int n0 = j * PF_N + c0;                         // first of the 32 output columns
      uint32_t acc[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * (PF_TILES * PF_N) + tl * PF_N + c0, acc);
      tmem_ld_wait();
      uint4 o[4];
      if (j < nkv) {
        const float4* bp = reinterpret_cast<const float4*>(s_bias + n0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = bp[2 * c], b1 = bp[2 * c + 1];
          __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c]) + b0.x, __uint_as_float(acc[8 * c + 1]) + b0.y);
          __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 2]) + b0.z, __uint_as_float(acc[8 * c + 3]) + b0.w);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 4]) + b1.x, __uint_as_float(acc[8 * c + 5]) + b1.y);
          __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 6]) + b1.z, __uint_as_float(acc[8 * c + 7]) + b1.w);
          o[c].x = *reinterpret_cast<uint32_t*>(&p0); o[c].y = *reinterpret_cast<uint32_t*>(&p1);
          o[c].z = *reinterpret_cast<uint32_t*>(&p2); o[c].w = *reinterpret_cast<uint32_t*>(&p3);
          if (!kv_valid) o[c] = make_uint4(0, 0, 0, 0);    // keys past the count are zero rows (pack_kv_gather_kernel)
        }
        if (kv_write) {
          const int l = n0 / (2 * TC_K), kvj = (n0 / TC_K) & 1, hh = (n0 >> 5) & 7;
          char* dst = reinterpret_cast<char*>(out.p[l]) + ((((size_t)(b * NH + hh) * ntile + tile) * 2 + kvj) * 64 + key) * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + ((c ^ ((key >> 1) & 3)) << 4)) = o[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c]), __uint_as_float(acc[8 * c + 1]));
          __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 2]), __uint_as_float(acc[8 * c + 3]));
          __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 4]), __uint_as_float(acc[8 * c + 5]));
          __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(acc[8 * c + 6]), __uint_as_float(acc[8 * c + 7]));
          o[c].x = *reinterpret_cast<uint32_t*>(&p0); o[c].y = *reinterpret_cast<uint32_t*>(&p1);
          o[c].z = *reinterpret_cast<uint32_t*>(&p2); o[c].w = *reinterpret_cast<uint32_t*>(&p3);
        }
        if (s_orig >= 0) {
          uint4* dst = reinterpret_cast<uint4*>(U + ((size_t)b * S + s_orig) * TC_K + (n0 - nkv * PF_N));
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = o[c];
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_bar + 32 + 8 * buf) : "memory");
  }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
  }
}

}  // namespace cb

using namespace cb;

extern "C" size_t case_vocab_tc_workspace_bytes(int R) { return (size_t)((R + TC_NB - 1) / TC_NB) * TC_A_BYTES; }
extern "C" size_t case_vocab_tc_packed_weight_bytes(int V) { return (size_t)((V + TC_M - 1) / TC_M) * TC_W_BYTES; }

// Wp: packed weights (see header comment); workspace: case_vocab_tc_workspace_bytes(R) bytes, 16-byte aligned.
extern "C" int case_vocab_gemm_tc(const float* f, const void* Wp, const float* bias, float* logits, int R, int V,
                                  int ldl, void* workspace, case_stream_t stream) {
  CB_REQUIRE(f && Wp && logits && workspace && R > 0 && V > 0 && ldl >= V, "case_vocab_gemm_tc: bad arguments");
  CB_REQUIRE(((uintptr_t)Wp % 16 == 0) && ((uintptr_t)workspace % 16 == 0), "case_vocab_gemm_tc: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  ensure_smem<vocab_gemm_tc_kernel>(TC_SMEM);
  const int nblk = (R + TC_NB - 1) / TC_NB;
  const int n = nblk * TC_NB * (TC_K / 8);
  launch_k(pack_activations_kernel, (n + 255) / 256, 256, 0, st, f, (bf16*)workspace, R);
  int rc = check_launch("case_vocab_gemm_tc(pack)");
  if (rc) return rc;
  launch_k(vocab_gemm_tc_kernel, (V + TC_M - 1) / TC_M, TC_THREADS, TC_SMEM, st, (const bf16*)Wp, (const bf16*)workspace, bias,
                                                                     logits, R, V, ldl);
  return check_launch("case_vocab_gemm_tc");
}

extern "C" int case_prefill_project_tc(const void* mem, const void* Wp, const float* bias, int B, int S, const int32_t* cidx,
                                       const int32_t* ncount, int nl, void* const* out, void* U, case_stream_t stream) {
  CB_REQUIRE(mem && Wp && bias && out && B > 0 && S > 0 && nl >= 1 && nl <= 4, "case_prefill_project_tc: bad arguments");
  CB_REQUIRE(((uintptr_t)mem % 16 == 0) && ((uintptr_t)Wp % 16 == 0) && ((uintptr_t)U % 16 == 0),
             "case_prefill_project_tc: 16-byte alignment required");
  CB_REQUIRE((cidx == nullptr) == (ncount == nullptr), "case_prefill_project_tc: cidx and ncount go together");
  CB_REQUIRE(B <= 65535, "case_prefill_project_tc: too many queries for one launch");
  PfOut o;
  for (int l = 0; l < 4; ++l) {
    o.p[l] = l < nl ? (bf16*)out[l] : nullptr;
    CB_REQUIRE(l >= nl || (out[l] && (uintptr_t)out[l] % 16 == 0), "case_prefill_project_tc: K|V outputs must be 16-byte aligned");
  }
  ensure_smem<prefill_project_tc_kernel>(PF_SMEM);
  launch_k(prefill_project_tc_kernel, dim3((S + PF_KEYS - 1) / PF_KEYS, B), PF_THREADS, PF_SMEM, (cudaStream_t)stream,
           (const bf16*)mem, (const bf16*)Wp, bias, S, cidx, ncount, nl, o, (bf16*)U);
  return check_launch("case_prefill_project_tc");
}
