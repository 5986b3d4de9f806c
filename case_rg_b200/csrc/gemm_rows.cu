// Row GEMM on tcgen05 / TMEM for the Linear layers of the pre-decode producers:
//
//   Y[M][N] = mask_rows( act( X[M][K] . W[N][K]^T + bias[N] ) + residual[M][N] )
//
// X: bf16 row-major (the LayerNorm / attention / previous GEMM output), K % 64 == 0; W: the nn.Linear weight packed per
// 64-wide K block into 256-row tiles of the UMMA K-major no-swizzle canonical layout ([K/64][N/256][32 KB]); Y: bf16 or
// fp32 row-major.  Persistent grid (one CTA per SM): a work unit is MT row tiles of 128 (UMMA M = 128 = the TMEM lanes) x one
// chunk of 256 columns (UMMA N = 256: 96 B/clk of shared-memory operand reads, against 128 B/clk at N = 128); CTAs stride
// over the chunk-minor unit list.  MT = 1: the accumulator D[128 x 256] (fp32) of a unit lives in one half of tensor memory
// while the epilogue drains the other half; MT = 2 (wide 1280-deep layers): both halves hold the unit's two accumulators,
// one W stage feeds 256 rows.  The K loop runs over a ring of (MT A tiles 128 x 64, W tile 256 x 64) stages - 4 x 48 KB or
// 3 x 64 KB - that runs on across units:
//   warp 0           producer (one thread): per stage one TMA tensor load of the X box (64 k x 128 MT rows; the tensor map
//                    carries the 128-byte swizzle the UMMA descriptor of the A operand expects, rows past M arrive as
//                    zeros) and one bulk copy of the pre-packed W tile, both credited to the stage's "full" mbarrier
//   warp 1           MMA issuer (one thread): 4 MT tcgen05.mma (M = 128, N = 256, K = 16) per stage; tcgen05.commit signals
//                    "stage free" and "accumulator complete"
//   warps 2-9        epilogue: tcgen05.ld (lane = row, 32 columns at a time), bias / gelu / relu / residual / row mask,
//                    residual loads and output stores staged through shared memory so that global memory sees whole
//                    sectors (4 lanes per row, 64 contiguous bytes)
// (First versions gathered A with 16-byte cp.async from four producer warps: 552 TFLOP/s with a no-swizzle layout - half-used
// sectors, L1TEX-bound - 1.0-1.19 PFLOP/s with the swizzled layout; the TMA load reaches 1.25 PFLOP/s on 3840 x 1280.)
#include <cuda.h>

#include "common.cuh"

namespace cb {

constexpr int GR_M = 128, GR_NC = 256, GR_KB = 64;
constexpr int GR_A_BYTES = GR_M * GR_KB * 2;         // 16 KB: one 128-byte-swizzle atom column (64 k) of 128 rows
constexpr int GR_W_BYTES = GR_NC * GR_KB * 2;        // 32 KB
// MT = row tiles of 128 per work unit.  MT = 1: 48 KB stages x 4, the two halves of tensor memory alternate between units.
// MT = 2: 64 KB stages x 3, both halves hold the unit's two accumulators: a W stage is read from L2 once for 256 rows, which
// is what the 1280-wide layers need - at MT = 1 they sit on the L2 -> SM throughput cap (960 KB per 128 x 256 x 1280 unit =
// 12 TB/s over 148 SMs at 1.05 PFLOP/s).
template <int MT> struct GrCfg {
  static constexpr int NS = MT == 1 ? 4 : 3;
  static constexpr int STAGE = MT * GR_A_BYTES + GR_W_BYTES;
};
constexpr int GR_THREADS = 10 * 32;                  // warp 0: producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int GR_STG = 32 * 80;                       // epilogue staging block of a warp: 32 rows x (64 + 16) bytes
constexpr int GR_RING = 192 * 1024;                   // NS x STAGE for both configurations
constexpr int GR_SMEM = GR_RING + 256 + 8 * GR_STG;

__device__ __forceinline__ void gr_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gr_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gr_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gr_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 2-D TMA tile load (box of the tensor map at element coordinates (c0 = k, c1 = row)) into shared memory; the bytes of the
// whole box - rows past the end of the tensor arrive as zeros - are credited to the mbarrier
__device__ __forceinline__ void gr_tma_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void gr_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint64_t gr_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// K-major operand in the 128-byte-swizzle layout: rows of 64 bf16 (128 B), 8-row groups 1024 B apart, the 16-byte chunk c of
// row r stored at chunk position c ^ (r & 7).  The leading-dimension offset is not used by this layout.
__device__ __forceinline__ uint64_t gr_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t gr_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void gr_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void gr_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gr_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gr_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gr_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void gr_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t gr_pk2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Tail of the epilogue for one 32-column group of a warp's 32 rows (v: this lane's row): + residual, row mask, store.
// Residual and output go through the warp's staging block (32 rows x 64 bytes, 80-byte row stride): global memory is
// touched with 4 lanes per row and 8 rows per instruction - whole 32-byte sectors, 64 contiguous bytes per row - while a
// lane reads / writes its own row in shared memory.
__device__ __forceinline__ void gr_finish_group(float (&v)[32], unsigned char* stg, int lane, long long mw, long long M, int N,
                                                int n0, const void* res, int res_bf16, bool keep, void* Y, int y_bf16) {
  if (res != nullptr) {
    const int nsb = res_bf16 ? 1 : 2;            // 64-byte sub-blocks per 32 columns
    const size_t rstride = (size_t)N * (res_bf16 ? 2 : 4);
    const char* rbase = reinterpret_cast<const char*>(res) + (size_t)n0 * (res_bf16 ? 2 : 4);
    for (int sb = 0; sb < nsb; ++sb) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        if (mw + r < M)
          *reinterpret_cast<uint4*>(stg + r * 80 + (lane & 3) * 16) =
              __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)(mw + r) * rstride + sb * 64 + (lane & 3) * 16));
      }
      __syncwarp();
      if (res_bf16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 w = *reinterpret_cast<const uint4*>(stg + lane * 80 + j * 16);
          v[8 * j] += __uint_as_float(w.x << 16); v[8 * j + 1] += __uint_as_float(w.x & 0xffff0000u);
          v[8 * j + 2] += __uint_as_float(w.y << 16); v[8 * j + 3] += __uint_as_float(w.y & 0xffff0000u);
          v[8 * j + 4] += __uint_as_float(w.z << 16); v[8 * j + 5] += __uint_as_float(w.z & 0xffff0000u);
          v[8 * j + 6] += __uint_as_float(w.w << 16); v[8 * j + 7] += __uint_as_float(w.w & 0xffff0000u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(stg + lane * 80 + j * 16);
          if (sb == 0) { v[4 * j] += w.x; v[4 * j + 1] += w.y; v[4 * j + 2] += w.z; v[4 * j + 3] += w.w; }
          else { v[16 + 4 * j] += w.x; v[16 + 4 * j + 1] += w.y; v[16 + 4 * j + 2] += w.z; v[16 + 4 * j + 3] += w.w; }
        }
      }
      __syncwarp();
    }
  }
  if (!keep) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  const int nsb = y_bf16 ? 1 : 2;
  const size_t ystride = (size_t)N * (y_bf16 ? 2 : 4);
  char* ybase = reinterpret_cast<char*>(Y) + (size_t)n0 * (y_bf16 ? 2 : 4);
  for (int sb = 0; sb < nsb; ++sb) {
    if (y_bf16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stg + lane * 80 + j * 16) =
            make_uint4(gr_pk2(v[8 * j], v[8 * j + 1]), gr_pk2(v[8 * j + 2], v[8 * j + 3]), gr_pk2(v[8 * j + 4], v[8 * j + 5]),
                       gr_pk2(v[8 * j + 6], v[8 * j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (sb == 0) *reinterpret_cast<float4*>(stg + lane * 80 + j * 16) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        else *reinterpret_cast<float4*>(stg + lane * 80 + j * 16) = make_float4(v[16 + 4 * j], v[16 + 4 * j + 1], v[16 + 4 * j + 2], v[16 + 4 * j + 3]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {             // streaming stores: Y passes through L2 once
      const int r = it * 8 + (lane >> 2);
      if (mw + r < M)
        __stcs(reinterpret_cast<uint4*>(ybase + (size_t)(mw + r) * ystride + sb * 64 + (lane & 3) * 16),
               *reinterpret_cast<const uint4*>(stg + r * 80 + (lane & 3) * 16));
    }
    __syncwarp();
  }
}

struct GemmRowsArgs {
  const bf16* X; const bf16* Wp; const float* bias;
  long long M; int N, K;
  int act;                       // 0 none, 1 gelu (erf), 2 relu
  const void* res; int res_bf16; // residual [M][N] fp32 or bf16, may be NULL
  const uint8_t* row_mask;       // [M] (1 = keep), may be NULL: masked rows are written as zeros
  void* Y; int y_bf16;
};

template <int MT>
__global__ __launch_bounds__(GR_THREADS, 1) void gemm_rows_tc_kernel(const GemmRowsArgs a, const __grid_constant__ CUtensorMap tmA) {
  constexpr int GR_NS = GrCfg<MT>::NS, GR_STAGE = GrCfg<MT>::STAGE;
  static_assert(GR_NS * GR_STAGE == GR_RING, "ring size");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_bar = s_base + GR_RING;
  const uint32_t BAR_A = s_bar, BAR_W = s_bar + 8 * GR_NS, BAR_FREE = s_bar + 16 * GR_NS, BAR_ACC = s_bar + 24 * GR_NS,
                 BAR_ACCF = BAR_ACC + 16;
  // barriers: per stage a_full (128 gather threads), w_full (4 issuing lanes + bytes), free (MMAs done); per accumulator
  // slot acc_full and acc_free (the epilogue warps that drain the slot)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + GR_RING + 192);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent: a work unit is (MT row tiles, 256-column chunk); this CTA owns the units blockIdx.x, blockIdx.x + gridDim.x,
  // ... of the chunk-minor unit list (neighbouring CTAs share the row tiles and all of them share the few weight chunks in
  // flight, so both operands are L2 hits); the stage ring runs on across units.  Accumulator slots (256 TMEM columns each):
  // MT = 1: slot = unit parity, drained by all 8 epilogue warps, so the epilogue of one unit overlaps the next one's MMAs;
  // MT = 2: slot = row tile of the unit, drained by 4 warps each.
  const int KB = a.K / GR_KB, nchunk = a.N / GR_NC;
  const long long total = ((a.M + MT * GR_M - 1) / (MT * GR_M)) * nchunk;
  const int nunit = (long long)blockIdx.x < total ? (int)((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int nit = nunit * KB;                           // ring iterations

  if (tid == 0) {
    for (int i = 0; i < GR_NS; ++i) {
      gr_mbar_init(BAR_A + 8 * i, 1);                   // stage full: one issuing thread + the bytes of the A box and the W tile
      gr_mbar_init(BAR_W + 8 * i, 1);                   // (unused)
      gr_mbar_init(BAR_FREE + 8 * i, 1);                // the MMAs that read the stage are done
    }
    gr_mbar_init(BAR_ACC, 1); gr_mbar_init(BAR_ACC + 8, 1);                       // accumulator slot complete
    gr_mbar_init(BAR_ACCF, 8 / MT); gr_mbar_init(BAR_ACCF + 8, 8 / MT);           // ... drained by its epilogue warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  gr_fence_before();
  __syncthreads();
  gr_fence_after();
  const uint32_t tmem = *s_tmem;
  pdl_wait();

  if (warp == 0) {
    // ================= producer: one thread feeds the ring - a TMA box (MT x 128 rows x 64 k of X, landed in the 128-byte-
    // swizzle K-major layout by the copy engine; rows past M arrive as zeros) and the pre-packed W tile as a bulk copy, both
    // credited to the stage's "full" barrier
    if (lane == 0) {
      for (int it = 0; it < nit; ++it) {
        const int g = it % GR_NS, use = it / GR_NS;
        const uint32_t s_a = s_base + g * GR_STAGE, s_w = s_a + MT * GR_A_BYTES;
        const int u = it / KB, kb = it - u * KB;
        const long long ug = (long long)blockIdx.x + (long long)u * gridDim.x;
        const long long mt = ug / nchunk;
        const int c = (int)(ug - mt * nchunk);
        if (use >= 1) gr_wait(BAR_FREE + 8 * g, (uint32_t)(use - 1) & 1u);      // the MMAs of this stage's previous use are done
        gr_expect_tx(BAR_A + 8 * g, MT * GR_A_BYTES + GR_W_BYTES);
        gr_tma_2d(s_a, &tmA, kb * GR_KB, (int)(mt * (MT * GR_M)), BAR_A + 8 * g);
        gr_bulk(s_w, reinterpret_cast<const char*>(a.Wp) + ((size_t)kb * nchunk + c) * GR_W_BYTES, GR_W_BYTES, BAR_A + 8 * g);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = gr_idesc(GR_M, GR_NC);
      int it = 0;
      for (int c = 0; c < nunit; ++c) {
        if (MT == 1 && c >= 2) gr_wait(BAR_ACCF + 8 * (c & 1), (uint32_t)((c >> 1) - 1) & 1u);    // epilogue of unit c - 2 done
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int g = it % GR_NS;
          const uint32_t par = (uint32_t)(it / GR_NS) & 1u;
          gr_wait(BAR_A + 8 * g, par);
          gr_fence_after();
          const uint32_t s_a = s_base + g * GR_STAGE, s_w = s_a + MT * GR_A_BYTES;
#pragma unroll
          for (int tile = 0; tile < MT; ++tile) {
            const int slot = MT == 1 ? (c & 1) : tile;
            // MT = 2: the slot still holds unit c - 1 until its four epilogue warps have drained it
            if (MT == 2 && kb == 0 && c >= 1) { gr_wait(BAR_ACCF + 8 * slot, (uint32_t)(c - 1) & 1u); gr_fence_after(); }
#pragma unroll
            for (int ks = 0; ks < GR_KB / 16; ++ks) {
              const int kc = ks * 2;
              const uint64_t ad = gr_desc_sw128(s_a + tile * GR_A_BYTES + ks * 32);
              const uint64_t bd = gr_desc(s_w + kc * (GR_NC / 8) * 128, (GR_NC / 8) * 128, 128);
              gr_umma(tmem + slot * GR_NC, ad, bd, idesc, (kb | ks) != 0);
            }
          }
          gr_commit(BAR_FREE + 8 * g);                   // stage free once these MMAs have read it
        }
        if (MT == 1) {
          gr_commit(BAR_ACC + 8 * (c & 1));              // accumulator of the unit complete
        } else {
          gr_commit(BAR_ACC); gr_commit(BAR_ACC + 8);    // both row tiles of the unit complete
        }
      }
    }
  } else {
    // ================= epilogue (warps 2-9): warp & 3 = TMEM lane quarter; hh = (warp - 2) >> 2 is, for MT = 1, the parity
    // of the 32-column groups the warp takes and, for MT = 2, the row tile (= accumulator slot) it drains
    const int q = warp & 3, hh = (warp - 2) >> 2;
    unsigned char* stg = smem + GR_RING + 256 + (warp - 2) * GR_STG;      // this warp's staging block
    for (int u = 0; u < nunit; ++u) {
      const long long ug = (long long)blockIdx.x + (long long)u * gridDim.x;
      const long long mt = ug / nchunk;
      const int c = (int)(ug - mt * nchunk);
      const long long m = mt * (MT * GR_M) + (MT == 2 ? hh * GR_M : 0) + q * 32 + lane;
      const bool rowok = m < a.M;
      const bool keep = rowok && (a.row_mask == nullptr || a.row_mask[m] != 0);
      const int tb = MT == 1 ? (u & 1) : hh;           // accumulator slot
      gr_wait(BAR_ACC + 8 * tb, (uint32_t)(MT == 1 ? (u >> 1) : u) & 1u);
      gr_fence_after();
      const long long mw = m - lane;                   // first row of the warp's 32
      for (int cg = (MT == 1 ? hh : 0); cg < GR_NC / 32; cg += (MT == 1 ? 2 : 1)) {
        const int n0 = c * GR_NC + cg * 32;
        uint32_t acc[32];
        gr_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tb * GR_NC + cg * 32, acc);
        gr_tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + j));
          v[j] = __uint_as_float(acc[j]) + b4.x; v[j + 1] = __uint_as_float(acc[j + 1]) + b4.y;
          v[j + 2] = __uint_as_float(acc[j + 2]) + b4.z; v[j + 3] = __uint_as_float(acc[j + 3]) + b4.w;
        }
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (a.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        gr_finish_group(v, stg, lane, mw, a.M, a.N, n0, a.res, a.res_bf16, keep, a.Y, a.y_bf16);
      }
      gr_fence_before();
      __syncwarp();
      if (lane == 0) gr_arrive(BAR_ACCF + 8 * tb);
    }
  }
  gr_fence_before();
  __syncthreads();
  gr_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}


// ------------------------------------------------------------------------------------------ fused feed-forward
//   Y[M][256] = mask_rows( act( X[M][K1] . W1[256][K1]^T + b1 ) . W2[256][256]^T + b2 + residual )
// TransformerEncoderLayer / TransformerBlock feed-forward (common/TransformerEncoder.py:73-76, TransformerBlock.py:30-32) in
// one launch: the hidden activations [M][256] never leave the SM (84 MB written + 84 MB read per call at the BASELINE
// shape otherwise).  Same roles and ring as gemm_rows_tc_kernel<1>; a unit = 128 rows runs KB1 = K1 / 64 ring stages of
// (A gathered, W1 block) into accumulator 1 (TMEM columns 0-255), the epilogue warps turn it into the bf16 hidden tile -
// written, 64-wide K block by K block, straight into the A halves of the four ring stages the second product uses (their
// W halves receive the four W2 blocks meanwhile) - and four more stages accumulate hidden . W2^T into accumulator 2
// (columns 256-511), whose epilogue (+ b2, residual, row mask, store) runs under the first product of the next unit.
struct FfnRowsArgs {
  const bf16* X; const bf16* W1p; const float* b1; const bf16* W2p; const float* b2;
  long long M; int K1; int act;
  const void* res; int res_bf16; const uint8_t* row_mask; void* Y; int y_bf16;
};

__global__ __launch_bounds__(GR_THREADS, 1) void ffn_rows_tc_kernel(const FfnRowsArgs a, const __grid_constant__ CUtensorMap tmA) {
  constexpr int NS = GrCfg<1>::NS, STAGE = GrCfg<1>::STAGE, KB2 = GR_NC / GR_KB;      // 4 stages of 48 KB; 4 K blocks of the hidden
  static_assert(KB2 == NS, "the hidden tile lives in the A halves of exactly one ring round");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_bar = s_base + GR_RING;
  const uint32_t BAR_A = s_bar, BAR_W = s_bar + 32, BAR_FREE = s_bar + 64, BAR_ACC1 = s_bar + 96, BAR_ACC2 = s_bar + 104,
                 BAR_H = s_bar + 112, BAR_ACC2F = s_bar + 120;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + GR_RING + 192);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB1 = a.K1 / GR_KB, KU = KB1 + KB2;          // ring iterations per unit
  const long long total = (a.M + GR_M - 1) / GR_M;
  const int nunit = (long long)blockIdx.x < total ? (int)((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int nit = nunit * KU;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      gr_mbar_init(BAR_A + 8 * i, 1);                     // stage full: one issuing thread + bytes
      gr_mbar_init(BAR_W + 8 * i, 1);                     // (unused)
      gr_mbar_init(BAR_FREE + 8 * i, 1);
    }
    gr_mbar_init(BAR_ACC1, 1); gr_mbar_init(BAR_ACC2, 1);
    gr_mbar_init(BAR_H, 8);                               // hidden tile written (8 epilogue warps) = accumulator 1 drained
    gr_mbar_init(BAR_ACC2F, 8);                           // accumulator 2 drained
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  gr_fence_before();
  __syncthreads();
  gr_fence_after();
  const uint32_t tmem = *s_tmem;
  pdl_wait();

  if (warp == 0) {
    // ================= producer (one thread): first product - TMA box of X + W1 block; second product - the W2 block only
    // (its A half is written by the epilogue warps)
    if (lane == 0) {
      for (int it = 0; it < nit; ++it) {
        const int g = it % NS, use = it / NS;
        const uint32_t s_a = s_base + g * STAGE, s_w = s_a + GR_A_BYTES;
        const int u = it / KU, j = it - u * KU;
        const long long m0 = ((long long)blockIdx.x + (long long)u * gridDim.x) * GR_M;
        if (use >= 1) gr_wait(BAR_FREE + 8 * g, (uint32_t)(use - 1) & 1u);
        if (j < KB1) {
          gr_expect_tx(BAR_A + 8 * g, GR_A_BYTES + GR_W_BYTES);
          gr_tma_2d(s_a, &tmA, j * GR_KB, (int)m0, BAR_A + 8 * g);
          gr_bulk(s_w, reinterpret_cast<const char*>(a.W1p) + (size_t)j * GR_W_BYTES, GR_W_BYTES, BAR_A + 8 * g);
        } else {
          gr_expect_tx(BAR_A + 8 * g, GR_W_BYTES);
          gr_bulk(s_w, reinterpret_cast<const char*>(a.W2p) + (size_t)(j - KB1) * GR_W_BYTES, GR_W_BYTES, BAR_A + 8 * g);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = gr_idesc(GR_M, GR_NC);
      int it = 0;
      for (int u = 0; u < nunit; ++u) {
        for (int j = 0; j < KU; ++j, ++it) {
          const int g = it % NS;
          const uint32_t par = (uint32_t)(it / NS) & 1u;
          gr_wait(BAR_A + 8 * g, par);
          if (j == KB1) {
            gr_wait(BAR_H, (uint32_t)u & 1u);            // hidden tile in place (and accumulator 1 drained)
            if (u >= 1) gr_wait(BAR_ACC2F, (uint32_t)(u - 1) & 1u);      // accumulator 2 of the previous unit drained
          }
          gr_fence_after();
          const uint32_t s_a = s_base + g * STAGE, s_w = s_a + GR_A_BYTES;
          const bool second = j >= KB1;
          const int jj = second ? j - KB1 : j;
#pragma unroll
          for (int ks = 0; ks < GR_KB / 16; ++ks) {
            const int kc = ks * 2;
            const uint64_t ad = gr_desc_sw128(s_a + ks * 32);
            const uint64_t bd = gr_desc(s_w + kc * (GR_NC / 8) * 128, (GR_NC / 8) * 128, 128);
            gr_umma(tmem + (second ? GR_NC : 0), ad, bd, idesc, (jj | ks) != 0);
          }
          gr_commit(BAR_FREE + 8 * g);
          if (j == KB1 - 1) gr_commit(BAR_ACC1);
          if (j == KU - 1) gr_commit(BAR_ACC2);
        }
      }
    }
  } else {
    // ================= epilogue warps: hidden tile, then the output
    const int q = warp & 3, hh = (warp - 2) >> 2;
    unsigned char* stg = smem + GR_RING + 256 + (warp - 2) * GR_STG;
    const int row = q * 32 + lane;                       // row of the tile = TMEM lane
    for (int u = 0; u < nunit; ++u) {
      const long long m = ((long long)blockIdx.x + (long long)u * gridDim.x) * GR_M + row;
      const bool rowok = m < a.M;
      const bool keep = rowok && (a.row_mask == nullptr || a.row_mask[m] != 0);
      const long long mw = m - lane;
      // ---- accumulator 1 -> act(. + b1) -> bf16 hidden tile in the A halves of the second product's stages
      gr_wait(BAR_ACC1, (uint32_t)u & 1u);
      gr_fence_after();
      for (int cg = hh; cg < GR_NC / 32; cg += 2) {
        uint32_t acc[32];
        gr_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cg * 32, acc);
        gr_tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b1 + cg * 32 + j));
          v[j] = __uint_as_float(acc[j]) + b4.x; v[j + 1] = __uint_as_float(acc[j + 1]) + b4.y;
          v[j + 2] = __uint_as_float(acc[j + 2]) + b4.z; v[j + 3] = __uint_as_float(acc[j + 3]) + b4.w;
        }
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (a.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const int g2 = (u * KU + KB1 + (cg >> 1)) % NS;  // the stage whose A half holds hidden columns 64 (cg >> 1) ..
        unsigned char* hb = smem + g2 * STAGE + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int kk8 = (cg & 1) * 4 + c4;             // 16-byte chunk of the 64-wide K block
          *reinterpret_cast<uint4*>(hb + ((kk8 ^ (row & 7)) << 4)) =
              make_uint4(gr_pk2(v[8 * c4], v[8 * c4 + 1]), gr_pk2(v[8 * c4 + 2], v[8 * c4 + 3]), gr_pk2(v[8 * c4 + 4], v[8 * c4 + 5]),
                         gr_pk2(v[8 * c4 + 6], v[8 * c4 + 7]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      gr_fence_before();
      __syncwarp();
      if (lane == 0) gr_arrive(BAR_H);
      // ---- accumulator 2 -> + b2, residual, row mask -> Y
      gr_wait(BAR_ACC2, (uint32_t)u & 1u);
      gr_fence_after();
      for (int cg = hh; cg < GR_NC / 32; cg += 2) {
        uint32_t acc[32];
        gr_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + GR_NC + cg * 32, acc);
        gr_tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b2 + cg * 32 + j));
          v[j] = __uint_as_float(acc[j]) + b4.x; v[j + 1] = __uint_as_float(acc[j + 1]) + b4.y;
          v[j + 2] = __uint_as_float(acc[j + 2]) + b4.z; v[j + 3] = __uint_as_float(acc[j + 3]) + b4.w;
        }
        gr_finish_group(v, stg, lane, mw, a.M, GR_NC, cg * 32, a.res, a.res_bf16, keep, a.Y, a.y_bf16);
      }
      gr_fence_before();
      __syncwarp();
      if (lane == 0) gr_arrive(BAR_ACC2F);
    }
  }
  gr_fence_before();
  __syncthreads();
  gr_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace cb

using namespace cb;

// TMA descriptor of X [M][K] bf16 row-major: boxes of 64 k x box_rows rows, 128-byte swizzle (what the UMMA descriptors of the
// A operand expect), zero fill past the end.  cuTensorMapEncodeTiled is fetched through the runtime (no libcuda link).
static int gr_make_tmap(CUtensorMap* tm, const void* X, long long M, int K, int box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn enc = nullptr;
  if (enc == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess) {
      cb::set_error("cuTensorMapEncodeTiled is not available from this driver");
      return (int)cudaErrorNotSupported;
    }
    enc = (encode_fn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)GR_KB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cb::set_error("cuTensorMapEncodeTiled failed for the row GEMM's X operand");
    return (int)cudaErrorNotSupported;
  }
  return 0;
}

/* packed weight bytes for an [N][K] Linear: K / 64 blocks x N / 256 tiles of 32 KB */
extern "C" size_t case_gemm_rows_packed_weight_bytes(int N, int K) {
  return (size_t)(K / GR_KB) * ((N + GR_NC - 1) / GR_NC) * GR_W_BYTES;
}

extern "C" int case_gemm_rows_tc(const void* X, const void* Wp, const float* bias, long long M, int N, int K, int act,
                                 const void* residual, int residual_dtype, const uint8_t* row_mask, void* Y, int y_dtype,
                                 case_stream_t stream) {
  CB_REQUIRE(X && Wp && bias && Y && M > 0, "case_gemm_rows_tc: null pointer");
  CB_REQUIRE(N > 0 && N % GR_NC == 0 && K > 0 && K % GR_KB == 0, "case_gemm_rows_tc: N must be a multiple of 256 and K of 64");
  CB_REQUIRE(act >= 0 && act <= 2, "case_gemm_rows_tc: act is 0 (none), 1 (gelu) or 2 (relu)");
  CB_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)Wp % 16 == 0) && ((uintptr_t)Y % 16 == 0) && ((uintptr_t)bias % 16 == 0) &&
                 ((uintptr_t)residual % 16 == 0),
             "case_gemm_rows_tc: 16-byte alignment required");
  CB_REQUIRE((y_dtype == CASE_F32 || y_dtype == CASE_BF16) && (!residual || residual_dtype == CASE_F32 || residual_dtype == CASE_BF16),
             "case_gemm_rows_tc: dtypes are CASE_F32 / CASE_BF16");
  CB_REQUIRE((M + GR_M - 1) / GR_M <= 0x7fffffffLL, "case_gemm_rows_tc: too many rows");
  GemmRowsArgs a;
  a.X = (const bf16*)X; a.Wp = (const bf16*)Wp; a.bias = bias; a.M = M; a.N = N; a.K = K; a.act = act;
  a.res = residual; a.res_bf16 = residual_dtype == CASE_BF16; a.row_mask = row_mask; a.Y = Y; a.y_bf16 = y_dtype == CASE_BF16;
  // two row tiles per unit where the weight stream is what bounds the launch: the 1280-deep, wide layers (3840 x 1280:
  // 1.58 -> 1.35 ms at M = 163,840; with a single 256-column chunk the un-overlapped epilogue of this form costs more
  // than the shared weight stage saves)
  const bool mt2 = K >= 1024 && N >= 2 * GR_NC && M >= 2 * GR_M;
  const long long units = ((M + (mt2 ? 2 : 1) * GR_M - 1) / ((mt2 ? 2 : 1) * GR_M)) * (N / GR_NC);
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if (nsm <= 0) nsm = 148;
  const unsigned grid = (unsigned)(units < nsm ? units : nsm);
  CUtensorMap tm;
  if (int rc = gr_make_tmap(&tm, X, M, K, (mt2 ? 2 : 1) * GR_M)) return rc;
  if (mt2) {
    ensure_smem<gemm_rows_tc_kernel<2>>(GR_SMEM);
    launch_k(gemm_rows_tc_kernel<2>, grid, GR_THREADS, GR_SMEM, (cudaStream_t)stream, a, tm);
  } else {
    ensure_smem<gemm_rows_tc_kernel<1>>(GR_SMEM);
    launch_k(gemm_rows_tc_kernel<1>, grid, GR_THREADS, GR_SMEM, (cudaStream_t)stream, a, tm);
  }
  return check_launch("case_gemm_rows_tc");
}

extern "C" int case_ffn_rows_tc(const void* X, const void* W1p, const float* b1, int K1, int act, const void* W2p, const float* b2,
                                long long M, const void* residual, int residual_dtype, const uint8_t* row_mask, void* Y,
                                int y_dtype, case_stream_t stream) {
  CB_REQUIRE(X && W1p && b1 && W2p && b2 && Y && M > 0, "case_ffn_rows_tc: null pointer");
  CB_REQUIRE(K1 > 0 && K1 % GR_KB == 0, "case_ffn_rows_tc: K1 must be a multiple of 64");
  CB_REQUIRE(act == 1 || act == 2, "case_ffn_rows_tc: act is 1 (gelu) or 2 (relu)");
  CB_REQUIRE(((uintptr_t)X % 16 == 0) && ((uintptr_t)W1p % 16 == 0) && ((uintptr_t)W2p % 16 == 0) && ((uintptr_t)Y % 16 == 0) &&
                 ((uintptr_t)b1 % 16 == 0) && ((uintptr_t)b2 % 16 == 0) && ((uintptr_t)residual % 16 == 0),
             "case_ffn_rows_tc: 16-byte alignment required");
  CB_REQUIRE((y_dtype == CASE_F32 || y_dtype == CASE_BF16) && (!residual || residual_dtype == CASE_F32 || residual_dtype == CASE_BF16),
             "case_ffn_rows_tc: dtypes are CASE_F32 / CASE_BF16");
  FfnRowsArgs a;
  a.X = (const bf16*)X; a.W1p = (const bf16*)W1p; a.b1 = b1; a.W2p = (const bf16*)W2p; a.b2 = b2; a.M = M; a.K1 = K1; a.act = act;
  a.res = residual; a.res_bf16 = residual_dtype == CASE_BF16; a.row_mask = row_mask; a.Y = Y; a.y_bf16 = y_dtype == CASE_BF16;
  const long long units = (M + GR_M - 1) / GR_M;
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if (nsm <= 0) nsm = 148;
  CUtensorMap tm;
  if (int rc = gr_make_tmap(&tm, X, M, K1, GR_M)) return rc;
  ensure_smem<ffn_rows_tc_kernel>(GR_SMEM);
  launch_k(ffn_rows_tc_kernel, (unsigned)(units < nsm ? units : nsm), GR_THREADS, GR_SMEM, (cudaStream_t)stream, a, tm);
  return check_launch("case_ffn_rows_tc");
}
