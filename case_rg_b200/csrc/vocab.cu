// Vocabulary-side kernels: output projection (SIMT fp32 form; the tcgen05 form lives in
// gemm_tcgen05.cu), softmax x mixture gate, copy-scatter into the distribution, per-row top-k.
#include "common.cuh"

namespace cb {

// ------------------------------------------------------------------------------------------ SIMT GEMM
// C[R][V] = A[R][H] . B[V][H]^T (+ bias), 64x64 tile, BK = 16, 256 threads x (4x4).
template <typename T>
__global__ __launch_bounds__(256) void vocab_gemm_simt_kernel(const float* __restrict__ A, const T* __restrict__ B,
                                                              const float* __restrict__ bias, float* __restrict__ C,
                                                              int R, int V, int ldc) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row lr (0..63), k offset lk (0,4,8,12)
  float acc[4][4] = {};
  for (int k0 = 0; k0 < H; k0 += BK) {
    float a4[4] = {0, 0, 0, 0}, b4[4] = {0, 0, 0, 0};
    if (m0 + lr < R) ld4(A + (size_t)(m0 + lr) * H + k0 + lk, a4);
    if (n0 + lr < V) ld4(B + (size_t)(n0 + lr) * H + k0 + lk, b4);
#pragma unroll
    for (int u = 0; u < 4; ++u) { As[lk + u][lr] = a4[u]; Bs[lk + u][lr] = b4[u]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < V) C[(size_t)m * ldc + n] = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------ softmax x gate
// dist[r, v] = gates[r][0] * exp(l[v] - max) / sum; one CTA per row, single online (max, sum) pass
// over the (L2-resident) logits row, then one write pass.
constexpr int SMT = 512;

__global__ __launch_bounds__(SMT) void softmax_mix_kernel(const float* __restrict__ logits, int ldl,
                                                          const float* __restrict__ gates, float* __restrict__ dist,
                                                          int ldd, int V, int mask_col0) {
  __shared__ float sm[SMT / 32], ss[SMT / 32];
  __shared__ float bm, bs;
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (size_t)r * ldl;
  const int V4 = V & ~3;
  float m = -INFINITY, s = 0.f;
  for (int i = tid * 4; i < V4; i += SMT * 4) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    if (mask_col0 && i == 0) v.x = -INFINITY;
    const float cm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (cm > m) { s *= fexp(m - cm); m = cm; }     // m = -inf -> fexp(-inf) = 0, s was 0
    if (m > -INFINITY) s += (fexp(v.x - m) + fexp(v.y - m)) + (fexp(v.z - m) + fexp(v.w - m));
  }
  for (int i = V4 + tid; i < V; i += SMT) {
    float v = x[i];
    if (mask_col0 && i == 0) v = -INFINITY;
    if (v > m) { s *= fexp(m - v); m = v; }
    if (m > -INFINITY) s += fexp(v - m);
  }
  // combine (m, s) pairs
  float M = warp_max(m);
  s = (m > -INFINITY) ? s * fexp(m - M) : 0.f;
  s = warp_sum(s);
  if (lane == 0) { sm[warp] = M; ss[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    float mm = lane < SMT / 32 ? sm[lane] : -INFINITY;
    float s2 = lane < SMT / 32 ? ss[lane] : 0.f;
    const float MM = warp_max(mm);
    s2 = (mm > -INFINITY) ? s2 * fexp(mm - MM) : 0.f;
    s2 = warp_sum(s2);
    if (lane == 0) { bm = MM; bs = s2; }
  }
  __syncthreads();
  const float MM = bm;
  const float sc = gates[(size_t)r * 4] / bs;
  float* d = dist + (size_t)r * ldd;
  for (int i = tid * 4; i < V4; i += SMT * 4) {
    float4 v = *reinterpret_cast<const float4*>(x + i);
    float4 o = make_float4(sc * fexp(v.x - MM), sc * fexp(v.y - MM), sc * fexp(v.z - MM), sc * fexp(v.w - MM));
    if (mask_col0 && i == 0) o.x = 0.f;
    *reinterpret_cast<float4*>(d + i) = o;
  }
  for (int i = V4 + tid; i < V; i += SMT) d[i] = (mask_col0 && i == 0) ? 0.f : sc * fexp(x[i] - MM);
}

// ------------------------------------------------------------------------------------------ copy scatter
// One thread per 4 consecutive source positions of one row: coalesced 16-byte reads of the
// attention weights, prior and map; one RED.ADD.F32 per non-zero weight.
__global__ __launch_bounds__(256) void copy_scatter_kernel(const int32_t* __restrict__ map, int map_ld, int map_off,
                                                           const float* __restrict__ prior,
                                                           const float* __restrict__ attn_un,
                                                           const float* __restrict__ fac, int fac_ld,
                                                           float* __restrict__ dist, int ldd, int W, int S, int V,
                                                           int vec_ok) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.y, b = r / W;
  const int s4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (s4 >= S) return;
  const float* a = attn_un + (size_t)r * S;
  const float* p = prior ? prior + (size_t)b * S : nullptr;
  const int32_t* mp = map + (size_t)b * map_ld + map_off;
  const float F = fac[(size_t)r * fac_ld], M = fac[(size_t)r * fac_ld + 1];
  if (F == 0.f) return;
  float* d = dist + (size_t)r * ldd;
  float av[4], pv[4] = {1.f, 1.f, 1.f, 1.f};
  int iv[4];
  const int n = min(4, S - s4);
  if (vec_ok && n == 4) {
    const float4 t = *reinterpret_cast<const float4*>(a + s4);
    av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
    if (p) { const float4 q = __ldg(reinterpret_cast<const float4*>(p + s4)); pv[0] = q.x; pv[1] = q.y; pv[2] = q.z; pv[3] = q.w; }
    const int4 ii = __ldg(reinterpret_cast<const int4*>(mp + s4));
    iv[0] = ii.x; iv[1] = ii.y; iv[2] = ii.z; iv[3] = ii.w;
  } else {
    for (int u = 0; u < 4; ++u) {
      av[u] = u < n ? a[s4 + u] : -INFINITY;
      if (p) pv[u] = u < n ? p[s4 + u] : 0.f;
      iv[u] = u < n ? mp[s4 + u] : 0;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (u >= n) break;
    if (av[u] == -INFINITY) continue;                       // masked source position
    const float c = F * pv[u] * fexp(av[u] - M);
    if (c != 0.f && (unsigned)iv[u] < (unsigned)V) atomicAdd(d + iv[u], c);
  }
}

// ------------------------------------------------------------------------------------------ top-k
struct VI { float v; int i; };
__device__ __forceinline__ bool better(float v, int i, float v2, int i2) { return v > v2 || (v == v2 && i < i2); }

template <int K>
__global__ __launch_bounds__(256) void topk_rows_kernel(const float* __restrict__ dist, int ldd, int V,
                                                        float* __restrict__ vals, int32_t* __restrict__ idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int swin;
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = dist + (size_t)r * ldd;
  float tv[K];
  int ti[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { tv[k] = -INFINITY; ti[k] = 0x7fffffff; }
  // indices visited in increasing order per thread -> strict '>' keeps the lower index on ties;
  // 16-byte loads, two in flight per thread (rows are 16-byte aligned: ldd % 4 == 0)
  auto push = [&](float v, int i) {
    if (v > tv[K - 1]) {
      tv[K - 1] = v; ti[K - 1] = i;
#pragma unroll
      for (int k = K - 1; k > 0; --k) {
        if (tv[k] > tv[k - 1]) {
          const float fv = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = fv;
          const int fi = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = fi;
        }
      }
    }
  };
  const int V4 = V & ~3;
  int i = tid * 4;
  for (; i + 1024 < V4; i += 2048) {
    const float4 a = *reinterpret_cast<const float4*>(x + i);
    const float4 b = *reinterpret_cast<const float4*>(x + i + 1024);
    push(a.x, i); push(a.y, i + 1); push(a.z, i + 2); push(a.w, i + 3);
    push(b.x, i + 1024); push(b.y, i + 1025); push(b.z, i + 1026); push(b.w, i + 1027);
  }
  for (; i < V4; i += 1024) {
    const float4 a = *reinterpret_cast<const float4*>(x + i);
    push(a.x, i); push(a.y, i + 1); push(a.z, i + 2); push(a.w, i + 3);
  }
  for (int j = V4 + tid; j < V; j += 256) push(x[j], j);
  for (int round = 0; round < K; ++round) {
    float bv = tv[0];
    int bi = ti[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float wv = sv[0];
      int wi = si[0];
      for (int w = 1; w < 8; ++w)
        if (better(sv[w], si[w], wv, wi)) { wv = sv[w]; wi = si[w]; }
      vals[(size_t)r * K + round] = wv;
      idx[(size_t)r * K + round] = wi;
      swin = wi;
    }
    __syncthreads();
    if (ti[0] == swin) {   // owner pops its head
#pragma unroll
      for (int k = 0; k < K - 1; ++k) { tv[k] = tv[k + 1]; ti[k] = ti[k + 1]; }
      tv[K - 1] = -INFINITY; ti[K - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}

}  // namespace cb

using namespace cb;

extern "C" int case_vocab_gemm(const float* f, const void* Wv, const float* bias, float* logits, int R, int V,
                               int ldl, int dtype, int impl, void* workspace, case_stream_t stream) {
  CB_REQUIRE(f && Wv && logits && R > 0 && V > 0 && ldl >= V, "case_vocab_gemm: bad arguments");
  if (impl == 1) {
    CB_REQUIRE(dtype == CASE_BF16, "case_vocab_gemm: the tensor-core path needs bf16 weights");
    return case_vocab_gemm_tc(f, Wv, bias, logits, R, V, ldl, workspace, stream);
  }
  dim3 grid((V + 63) / 64, (R + 63) / 64);
  if (dtype == CASE_BF16)
    launch_k(vocab_gemm_simt_kernel<bf16>, grid, 256, 0, (cudaStream_t)stream, f, (const bf16*)Wv, bias, logits, R, V, ldl);
  else
    launch_k(vocab_gemm_simt_kernel<float>, grid, 256, 0, (cudaStream_t)stream, f, (const float*)Wv, bias, logits, R, V, ldl);
  return check_launch("case_vocab_gemm");
}

extern "C" int case_softmax_mix(const float* logits, int ldl, const float* gates, float* dist, int ldd, int R, int V,
                                int mask_col0, case_stream_t stream) {
  CB_REQUIRE(logits && gates && dist && R > 0 && V > 0, "case_softmax_mix: bad arguments");
  CB_REQUIRE(ldl % 4 == 0 && ldd % 4 == 0 && ldl >= V && ldd >= V, "case_softmax_mix: row strides must be multiples of 4");
  CB_REQUIRE(((uintptr_t)logits % 16 == 0) && ((uintptr_t)dist % 16 == 0), "case_softmax_mix: 16-byte alignment required");
  launch_k(softmax_mix_kernel, R, SMT, 0, (cudaStream_t)stream, logits, ldl, gates, dist, ldd, V, mask_col0);
  return check_launch("case_softmax_mix");
}

extern "C" int case_copy_scatter(const int32_t* map, int map_ld, int map_off, const float* prior,
                                 const float* attn_un, const float* fac, int fac_ld, float* dist,
                                 int ldd, int B, int W, int S, int V, case_stream_t stream) {
  CB_REQUIRE(map && attn_un && fac && dist && B > 0 && W > 0 && S > 0 && fac_ld >= 2, "case_copy_scatter: bad arguments");
  const int vec_ok = (S % 4 == 0) && (map_ld % 4 == 0) && (map_off % 4 == 0) && ((uintptr_t)map % 16 == 0) &&
                     ((uintptr_t)attn_un % 16 == 0) && (!prior || (uintptr_t)prior % 16 == 0);
  dim3 grid((S + 1023) / 1024, B * W);
  launch_k(copy_scatter_kernel, grid, 256, 0, (cudaStream_t)stream, map, map_ld, map_off, prior, attn_un, fac, fac_ld,
                                                              dist, ldd, W, S, V, vec_ok);
  return check_launch("case_copy_scatter");
}

namespace cb {
// one CTA per row: thread d folds dynamic entry d into its vocabulary id and applies the overlap mask
__global__ __launch_bounds__(256) void oov_fold_kernel(float* __restrict__ gen, int ldg, int V, int D,
                                                       const int32_t* __restrict__ vmap, const float* __restrict__ overlap,
                                                       int rows_per_map) {
  pdl_wait();
  const int r = blockIdx.x, q = r / rows_per_map;
  float* row = gen + (size_t)r * ldg;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float w = row[V + d];
    const int v = vmap[(size_t)q * D + d];
    if (w != 0.f && (unsigned)v < (unsigned)V) atomicAdd(row + v, w);   // several OOV words fold onto UNK
    row[V + d] = w * overlap[(size_t)q * D + d];
  }
}
}  // namespace cb

extern "C" int case_oov_fold(float* gen, int ldg, int R, int V, int D, const int32_t* vocab_map, const float* overlap,
                             int rows_per_map, case_stream_t stream) {
  CB_REQUIRE(gen && vocab_map && overlap && R > 0 && V > 0 && D > 0 && rows_per_map >= 1, "case_oov_fold: bad arguments");
  CB_REQUIRE(ldg >= V + D, "case_oov_fold: rows must hold V + D entries");
  launch_k(cb::oov_fold_kernel, R, 256, 0, (cudaStream_t)stream, gen, ldg, V, D, vocab_map, overlap, rows_per_map);
  return check_launch("case_oov_fold");
}

extern "C" int case_topk_rows(const float* dist, int ldd, int R, int V, int k, float* vals, int32_t* idx,
                              case_stream_t stream) {
  CB_REQUIRE(dist && vals && idx && R > 0 && V > 0, "case_topk_rows: bad arguments");
  CB_REQUIRE(k >= 1 && k <= CASE_MAX_W && k <= V, "case_topk_rows: k out of range");
  CB_REQUIRE(ldd % 4 == 0 && (uintptr_t)dist % 16 == 0, "case_topk_rows: rows must be 16-byte aligned (ldd % 4 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  // vals / idx are [R][k]; kernels are instantiated for k = 1, 2, 4, 8 and others fall to the next size up
  switch (k) {
    case 1: launch_k(topk_rows_kernel<1>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 2: launch_k(topk_rows_kernel<2>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 3: launch_k(topk_rows_kernel<3>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 4: launch_k(topk_rows_kernel<4>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 5: launch_k(topk_rows_kernel<5>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 6: launch_k(topk_rows_kernel<6>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    case 7: launch_k(topk_rows_kernel<7>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
    default: launch_k(topk_rows_kernel<8>, R, 256, 0, st, dist, ldd, V, vals, idx); break;
  }
  return check_launch("case_topk_rows");
}

// ------------------------------------------------------------------------------------------ prefill packing
// fp32 projected memory rows [B*S][ldkv] (columns: layer, K|V, head, dim - the output of the prefill
// GEMM) -> per layer the tensor-core cross-attention layout: bf16 [B][NH][ceil(S/64)][2][64][32] with
// the 16-byte chunks of every key row XOR-swizzled.  One thread per 16-byte output chunk; consecutive
// threads walk one source row, so reads are fully coalesced and writes land as 64-byte runs.
namespace cb {
struct KvOut { bf16* p[4]; };
template <typename TS>
__global__ __launch_bounds__(256) void pack_kv_tiles_kernel(const TS* __restrict__ kv, int ldkv, int B, int S,
                                                            int nl, KvOut out) {
  pdl_trigger();
  pdl_wait();
  const int ntile = (S + 63) / 64;
  const size_t per_row = (size_t)nl * 2 * NH * 4;                 // 16-byte chunks per source row
  const size_t total = (size_t)B * ntile * 64 * per_row;          // padded keys included (zero-filled)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 4), hh = (int)((i / 4) % NH), j = (int)((i / (4 * NH)) % 2), l = (int)((i / (8 * NH)) % nl);
    const size_t row = i / per_row;
    const int key = (int)(row % 64), tile = (int)((row / 64) % ntile), b = (int)(row / ((size_t)64 * ntile));
    const int s = tile * 64 + key;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s < S) {
      const TS* src = kv + ((size_t)b * S + s) * ldkv + ((l * 2 + j) * NH + hh) * HD + c * 8;
      if (sizeof(TS) == 2) {
        v = __ldg(reinterpret_cast<const uint4*>(src));
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), d = __ldg(reinterpret_cast<const float4*>(src) + 1);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(d.x, d.y), p3 = __floats2bfloat162_rn(d.z, d.w);
        v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
      }
    }
    char* dst = reinterpret_cast<char*>(out.p[l]) + ((((size_t)(b * NH + hh) * ntile + tile) * 2 + j) * 64 + key) * 64 +
                ((c ^ ((key >> 1) & 3)) << 4);
    *reinterpret_cast<uint4*>(dst) = v;
  }
}
}  // namespace cb

extern "C" int case_pack_kv_tiles(const void* kv, int src_dtype, int ldkv, int B, int S, int nl, void* const* out,
                                  case_stream_t stream) {
  CB_REQUIRE(kv && out && B > 0 && S > 0 && nl >= 1 && nl <= 4, "case_pack_kv_tiles: bad arguments");
  CB_REQUIRE(ldkv % 8 == 0 && (uintptr_t)kv % 16 == 0, "case_pack_kv_tiles: source rows must be 16-byte aligned");
  cb::KvOut o;
  for (int l = 0; l < 4; ++l) o.p[l] = l < nl ? (cb::bf16*)out[l] : nullptr;
  if (src_dtype == CASE_BF16)
    launch_k(cb::pack_kv_tiles_kernel<cb::bf16>, 148 * 8, 256, 0, (cudaStream_t)stream, (const cb::bf16*)kv, ldkv, B, S,
             nl, o);
  else
    launch_k(cb::pack_kv_tiles_kernel<float>, 148 * 8, 256, 0, (cudaStream_t)stream, (const float*)kv, ldkv, B, S, nl, o);
  return cb::check_launch("case_pack_kv_tiles");
}
