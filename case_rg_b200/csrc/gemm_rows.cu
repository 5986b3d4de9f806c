// Row GEMM on tcgen05 / TMEM for the Linear layers of the pre-decode producers:
//
//   Y[M][N] = mask_rows( act( X[M][K] . W[N][K]^T + bias[N] ) + residual[M][N] )
//
// X: bf16 row-major (the LayerNorm / attention / previous GEMM output), K % 128 == 0; W: the nn.Linear weight packed per
// 128-wide K block into 256-row tiles of the UMMA K-major no-swizzle canonical layout ([K/128][N/256][64 KB]); Y: bf16 or
// fp32 row-major.  One CTA owns 128 rows (UMMA M = 128 = the TMEM lanes) and walks N in chunks of 256 columns (UMMA
// N = 256: 96 B/clk of shared-memory operand reads, against 128 B/clk at N = 128); a chunk's accumulator D[128 x 256]
// (fp32) lives in one half of tensor memory while the epilogue drains the other half.  The K loop runs over a two-stage
// ring of (A tile 128 x 128, W tile 256 x 128) = 96 KB per stage:
//   warps 0-3 / 4-7  producers of the even / odd stages: the A tile is gathered with 16-byte cp.async (16 per thread, all
//                    in flight, a full 128-byte line per quarter warp; rows past M are zero-filled) into the 128-byte-swizzle
//                    K-major layout (conflict-free on the shared-memory side), the pre-packed W tile arrives as four bulk
//                    copies; two stages are in flight at any time, so the gather latency of one hides behind the other
//   warp 16          MMA issuer (one thread): 8 tcgen05.mma (M = 128, N = 256, K = 16) per stage; tcgen05.commit signals
//                    "stage free" and "accumulator complete"
//   warps 8-15       epilogue: tcgen05.ld (lane = row, 32 columns at a time), bias / gelu / relu / residual / row mask,
//                    64-byte (bf16) or 128-byte (fp32) stores per thread
#include "common.cuh"

namespace cb {

constexpr int GR_M = 128, GR_NC = 256, GR_KB = 128;
constexpr int GR_A_BYTES = GR_M * GR_KB * 2;         // 32 KB
constexpr int GR_W_BYTES = GR_NC * GR_KB * 2;        // 64 KB
constexpr int GR_STAGE = GR_A_BYTES + GR_W_BYTES;    // 96 KB
constexpr int GR_THREADS = 17 * 32;
constexpr int GR_SMEM = 2 * GR_STAGE + 256;

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

struct GemmRowsArgs {
  const bf16* X; const bf16* Wp; const float* bias;
  long long M; int N, K;
  int act;                       // 0 none, 1 gelu (erf), 2 relu
  const void* res; int res_bf16; // residual [M][N] fp32 or bf16, may be NULL
  const uint8_t* row_mask;       // [M] (1 = keep), may be NULL: masked rows are written as zeros
  void* Y; int y_bf16;
};

__global__ __launch_bounds__(GR_THREADS, 1) void gemm_rows_tc_kernel(const GemmRowsArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_bar = s_base + 2 * GR_STAGE;
  // barriers: [0,1] a_full (128 gather threads), [2,3] w_full (4 issuing lanes + bytes), [4,5] stage free (MMAs done),
  //           [6,7] acc_full, [8,9] acc_free (8 epilogue warps)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 2 * GR_STAGE + 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // persistent: this CTA owns the row tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a work unit is (row tile, 256-column
  // chunk), the stage ring and the two accumulators run on across units so the epilogue of one overlaps the next one's MMAs
  const int KB = a.K / GR_KB, nchunk = a.N / GR_NC;
  const int ntile = (int)((a.M + GR_M - 1) / GR_M);
  const int mine = ((int)blockIdx.x < ntile) ? (ntile - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nunit = mine * nchunk;
  const int nit = nunit * KB;                           // ring iterations

  if (tid == 0) {
    gr_mbar_init(s_bar, 128); gr_mbar_init(s_bar + 8, 128);
    gr_mbar_init(s_bar + 16, 4); gr_mbar_init(s_bar + 24, 4);
    gr_mbar_init(s_bar + 32, 1); gr_mbar_init(s_bar + 40, 1);
    gr_mbar_init(s_bar + 48, 1); gr_mbar_init(s_bar + 56, 1);
    gr_mbar_init(s_bar + 64, 8); gr_mbar_init(s_bar + 72, 8);
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

  if (warp < 8) {
    // ================= producers: group g = warp / 4 feeds the stages it with (it & 1) == g
    const int g = warp >> 2, w4 = warp & 3;
    const uint32_t s_a = s_base + g * GR_STAGE, s_w = s_a + GR_A_BYTES;
    // gather role: a quarter warp (8 lanes) copies one 128-byte run of a source row - one full cache line per request - into
    // the 128-byte-swizzle layout, where its eight 16-byte chunks land in eight different bank groups
    const int rq = w4 * 4 + (lane >> 3), ch = lane & 7;
    for (int it = g; it < nit; it += 2) {
      const int u = it / KB, kb = it - u * KB, use = it >> 1;
      const int mt = u / nchunk, c = u - mt * nchunk;
      const long long m0 = ((long long)blockIdx.x + (long long)mt * gridDim.x) * GR_M;
      if (use >= 1) gr_wait(s_bar + 32 + 8 * g, (uint32_t)(use - 1) & 1u);      // the MMAs of this stage's previous use are done
      if (lane == 0) {
        const char* src = reinterpret_cast<const char*>(a.Wp) + ((size_t)kb * nchunk + c) * GR_W_BYTES + (size_t)w4 * (GR_W_BYTES / 4);
        gr_expect_tx(s_bar + 16 + 8 * g, GR_W_BYTES / 4);
        gr_bulk(s_w + w4 * (GR_W_BYTES / 4), src, GR_W_BYTES / 4, s_bar + 16 + 8 * g);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {                      // the two 64-wide K atoms of the stage
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = i * 16 + rq;
          const long long m = m0 + row;
          const bf16* src = a.X + (size_t)(m < a.M ? m : 0) * a.K + kb * GR_KB + h * 64 + ch * 8;
          const uint32_t dst = s_a + h * (GR_A_BYTES / 2) + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
          const int nbytes = m < a.M ? 16 : 0;           // rows past M are zero-filled
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      gr_arrive(s_bar + 8 * g);
    }
  } else if (warp == 16) {
    // ================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = gr_idesc(GR_M, GR_NC);
      int it = 0;
      for (int c = 0; c < nunit; ++c) {
        const int tb = c & 1;
        if (c >= 2) gr_wait(s_bar + 64 + 8 * tb, (uint32_t)((c >> 1) - 1) & 1u);      // epilogue of unit c - 2 done
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int g = it & 1;
          const uint32_t par = (uint32_t)(it >> 1) & 1u;
          gr_wait(s_bar + 8 * g, par);
          gr_wait(s_bar + 16 + 8 * g, par);
          gr_fence_after();
          const uint32_t s_a = s_base + g * GR_STAGE, s_w = s_a + GR_A_BYTES;
#pragma unroll
          for (int ks = 0; ks < GR_KB / 16; ++ks) {
            const int kc = ks * 2;
            const uint64_t ad = gr_desc_sw128(s_a + (ks >> 2) * (GR_A_BYTES / 2) + (ks & 3) * 32);
            const uint64_t bd = gr_desc(s_w + kc * (GR_NC / 8) * 128, (GR_NC / 8) * 128, 128);
            gr_umma(tmem + tb * GR_NC, ad, bd, idesc, (kb | ks) != 0);
          }
          gr_commit(s_bar + 32 + 8 * g);                 // stage free once these MMAs have read it
        }
        gr_commit(s_bar + 48 + 8 * tb);                  // accumulator of the chunk complete
      }
    }
  } else {
    // ================= epilogue (warps 8-15): warp & 3 = TMEM lane quarter, (warp - 8) >> 2 = which 32-column groups
    const int q = warp & 3, hh = (warp - 8) >> 2;
    for (int u = 0; u < nunit; ++u) {
      const int mt = u / nchunk, c = u - mt * nchunk;
      const long long m = ((long long)blockIdx.x + (long long)mt * gridDim.x) * GR_M + q * 32 + lane;
      const bool rowok = m < a.M;
      const bool keep = rowok && (a.row_mask == nullptr || a.row_mask[m] != 0);
      const int tb = u & 1;
      gr_wait(s_bar + 48 + 8 * tb, (uint32_t)(u >> 1) & 1u);
      gr_fence_after();
      for (int cg = hh; cg < GR_NC / 32; cg += 2) {
        const int n0 = c * GR_NC + cg * 32;
        uint32_t acc[32];
        gr_tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tb * GR_NC + cg * 32, acc);
        gr_tmem_ld_wait();
        if (rowok) {
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
          if (a.res != nullptr) {
            if (a.res_bf16) {
              const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.res) + (size_t)m * a.N + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 w = rp[j];
                v[8 * j] += __uint_as_float(w.x << 16); v[8 * j + 1] += __uint_as_float(w.x & 0xffff0000u);
                v[8 * j + 2] += __uint_as_float(w.y << 16); v[8 * j + 3] += __uint_as_float(w.y & 0xffff0000u);
                v[8 * j + 4] += __uint_as_float(w.z << 16); v[8 * j + 5] += __uint_as_float(w.z & 0xffff0000u);
                v[8 * j + 6] += __uint_as_float(w.w << 16); v[8 * j + 7] += __uint_as_float(w.w & 0xffff0000u);
              }
            } else {
              const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.res) + (size_t)m * a.N + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w = rp[j];
                v[4 * j] += w.x; v[4 * j + 1] += w.y; v[4 * j + 2] += w.z; v[4 * j + 3] += w.w;
              }
            }
          }
          if (!keep) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (a.y_bf16) {
            uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.Y) + (size_t)m * a.N + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j)     // streaming stores: Y passes through L2 once, X and W tiles are re-read from it
              __stcs(yp + j, make_uint4(gr_pk2(v[8 * j], v[8 * j + 1]), gr_pk2(v[8 * j + 2], v[8 * j + 3]),
                                        gr_pk2(v[8 * j + 4], v[8 * j + 5]), gr_pk2(v[8 * j + 6], v[8 * j + 7])));
          } else {
            float4* yp = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.Y) + (size_t)m * a.N + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) __stcs(yp + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
        }
      }
      gr_fence_before();
      __syncwarp();
      if (lane == 0) gr_arrive(s_bar + 64 + 8 * tb);
    }
  }
  gr_fence_before();
  __syncthreads();
  gr_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace cb

using namespace cb;

/* packed weight bytes for an [N][K] Linear: K / 128 blocks x N / 256 tiles of 64 KB */
extern "C" size_t case_gemm_rows_packed_weight_bytes(int N, int K) {
  return (size_t)(K / GR_KB) * ((N + GR_NC - 1) / GR_NC) * GR_W_BYTES;
}

extern "C" int case_gemm_rows_tc(const void* X, const void* Wp, const float* bias, long long M, int N, int K, int act,
                                 const void* residual, int residual_dtype, const uint8_t* row_mask, void* Y, int y_dtype,
                                 case_stream_t stream) {
  CB_REQUIRE(X && Wp && bias && Y && M > 0, "case_gemm_rows_tc: null pointer");
  CB_REQUIRE(N > 0 && N % GR_NC == 0 && K > 0 && K % GR_KB == 0, "case_gemm_rows_tc: N must be a multiple of 256 and K of 128");
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
  ensure_smem<gemm_rows_tc_kernel>(GR_SMEM);
  const long long ntile = (M + GR_M - 1) / GR_M;
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if (nsm <= 0) nsm = 148;
  launch_k(gemm_rows_tc_kernel, (unsigned)(ntile < nsm ? ntile : nsm), GR_THREADS, GR_SMEM, (cudaStream_t)stream, a);
  return check_launch("case_gemm_rows_tc");
}
