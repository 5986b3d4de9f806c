// Row-local pieces of the decode step: embedding, LayerNorm, the k-major row linear, the two
// fused halves of a decoder layer, the row finalisers and the GRU cell.
//
// Everything here works on RB rows per CTA with the activations in shared memory and streams the
// (L2-resident) weight matrices once per CTA through a bulk-async-copy ring (see WStream).
#include "common.cuh"

namespace cb {

constexpr int RB = 4;       // rows per CTA
constexpr int NT = 256;     // threads per CTA (== one output tile of 256 columns)

// ------------------------------------------------------------------------------------------
// Weight streaming.  Every weight matrix is stored "n-tile major": [N/256][K][256] in T, so the
// weights one CTA needs for a 256-column output tile are one contiguous run, cut into 16 KB tiles
// of KT k-rows.  A single thread feeds a ring of NSTAGE shared-memory stages with 1-D bulk async
// copies (cp.async.bulk -> UBLKCP, completion on an mbarrier); all 256 threads consume a tile, hit
// one __syncthreads and the freed stage is re-armed for the tile NSTAGE ahead.  The stream runs
// across the consecutive linears of a fused kernel, so the next matrix is already in flight while
// LayerNorm / self-attention run.
constexpr int TILE_BYTES = 16384;
constexpr int NSTAGE = 6;
template <typename T> struct KTile { static constexpr int KT = TILE_BYTES / (256 * (int)sizeof(T)); };

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
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

struct WStream {
  const char* base[3];
  int end[3];          // cumulative tile counts of the (up to 3) weight matrices, in consumption order
  int total;
  uint64_t* full;      // [NSTAGE] mbarriers (shared)
  char* stage;         // [NSTAGE][TILE_BYTES] (shared)
  int g;               // next tile to consume (block-uniform)

  __device__ __forceinline__ const char* src(int p) const {
    int s = 0, first = 0;
    if (p >= end[0]) { s = 1; first = end[0]; }
    if (p >= end[1]) { s = 2; first = end[1]; }
    return base[s] + (size_t)(p - first) * TILE_BYTES;
  }
  __device__ __forceinline__ void issue(int p) const {
    const int s = p % NSTAGE;
    mbar_expect_tx(&full[s], TILE_BYTES);
    bulk_g2s(stage + (size_t)s * TILE_BYTES, src(p), TILE_BYTES, &full[s]);
  }
  // call once by all threads, before any consumption
  __device__ __forceinline__ void start() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    // a warp keeps one bulk copy in flight at a time (profiles/micro/bulk_bench.cu): spread the issue
    if ((threadIdx.x & 31) == 0) {
      const int p = threadIdx.x >> 5;
      if (p < NSTAGE && p < total) issue(p);
    }
    g = 0;
  }
};

__device__ __forceinline__ void lds4(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void lds4(const bf16* p, float (&o)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
}

// One 256-column output tile over K inputs for RB rows: consumes K/KT tiles of the stream.  Thread
// (kg = tid/64, n4 = tid%64) owns columns 4*n4..+3 and a quarter of each tile's k-rows; the four
// k-quarters are parked in red[kg][rb][col] for the caller to sum after a __syncthreads().
template <typename T>
__device__ __forceinline__ void rows_linear_stream(WStream& ws, const float* __restrict__ xs, int ldx, int K,
                                                   float* __restrict__ red) {
  constexpr int KT = KTile<T>::KT, KPT = KT / 4;
  const int tid = threadIdx.x, n4 = tid & 63, kg = tid >> 6;
  float acc[RB][4];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[rb][j] = 0.f;
  for (int kt = 0; kt < K / KT; ++kt) {
    const int g = ws.g, s = g % NSTAGE;
    const uint32_t par = (uint32_t)(g / NSTAGE) & 1u;
    while (!mbar_try_wait(&ws.full[s], par)) {}
    const T* wt = reinterpret_cast<const T*>(ws.stage + (size_t)s * TILE_BYTES) + (kg * KPT) * 256 + n4 * 4;
    const float* xp = xs + kt * KT + kg * KPT;
#pragma unroll
    for (int kk = 0; kk < KPT; kk += 4) {
      float w[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) lds4(wt + (kk + u) * 256, w[u]);
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + rb * ldx + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[rb][j] = fmaf(xv.x, w[0][j], acc[rb][j]);
          acc[rb][j] = fmaf(xv.y, w[1][j], acc[rb][j]);
          acc[rb][j] = fmaf(xv.z, w[2][j], acc[rb][j]);
          acc[rb][j] = fmaf(xv.w, w[3][j], acc[rb][j]);
        }
      }
    }
    __syncthreads();                                   // every thread is done reading stage s
    if ((tid & 31) == 0 && (tid >> 5) == (g & 7) && g + NSTAGE < ws.total) ws.issue(g + NSTAGE);
    ws.g = g + 1;
  }
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
    *reinterpret_cast<float4*>(red + (kg * RB + rb) * 256 + n4 * 4) =
        make_float4(acc[rb][0], acc[rb][1], acc[rb][2], acc[rb][3]);
}

__device__ __forceinline__ float red_sum(const float* red, int rb, int col) {
  return (red[(0 * RB + rb) * 256 + col] + red[(1 * RB + rb) * 256 + col]) +
         (red[(2 * RB + rb) * 256 + col] + red[(3 * RB + rb) * 256 + col]);
}

// LayerNorm of RB rows of width H held in shared memory (row stride ld), in place; warp rb does row rb.
__device__ __forceinline__ void ln_rows_smem(float* xs, int ld, const float* __restrict__ g,
                                             const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < RB) {
    float* x = xs + warp * ld;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = x[lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) * (1.f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / H) + LN_EPS);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = lane + 32 * i;
      x[n] = (v[i] - mean) * rstd * __ldg(g + n) + __ldg(b + n);
    }
  }
}

// ------------------------------------------------------------------------------------------ embed
__global__ void embed_kernel(const float* __restrict__ E, const float* __restrict__ pe,
                             const int32_t* __restrict__ tok, int tok_ld, int t, float scale,
                             float* __restrict__ x, int R) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 6);
  if (r >= R) return;
  const int c = (threadIdx.x & 63) * 4;
  const int id = tok[(size_t)r * tok_ld + t];
  float4 e = __ldg(reinterpret_cast<const float4*>(E + (size_t)id * H + c));
  float4 p = pe ? __ldg(reinterpret_cast<const float4*>(pe + (size_t)t * H + c)) : make_float4(0, 0, 0, 0);
  *reinterpret_cast<float4*>(x + (size_t)r * H + c) =
      make_float4(fmaf(e.x, scale, p.x), fmaf(e.y, scale, p.y), fmaf(e.z, scale, p.z), fmaf(e.w, scale, p.w));
}

__global__ void layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                      const float* __restrict__ b, float* __restrict__ y, int R) {
  __shared__ __align__(16) float xs[RB][H];
  pdl_trigger();
  pdl_wait();
  const int r0 = blockIdx.x * RB;
  for (int i = threadIdx.x; i < RB * H; i += NT) {
    const int rb = i / H, r = r0 + rb;
    xs[rb][i % H] = r < R ? x[(size_t)r * H + i % H] : 0.f;
  }
  __syncthreads();
  ln_rows_smem(&xs[0][0], H, g, b);
  __syncthreads();
  for (int i = threadIdx.x; i < RB * H; i += NT) {
    const int rb = i / H, r = r0 + rb;
    if (r < R) y[(size_t)r * H + i % H] = xs[rb][i % H];
  }
}

// ------------------------------------------------------------------------------------------ shared memory plan
// [ring: NSTAGE x 16 KB][mbarriers: 64 B][kernel-specific floats]
constexpr int RING_BYTES = NSTAGE * TILE_BYTES + 64;

__device__ __forceinline__ void stream_setup(WStream& ws, unsigned char* smem_raw) {
  ws.stage = reinterpret_cast<char*>(smem_raw);
  ws.full = reinterpret_cast<uint64_t*>(smem_raw + NSTAGE * TILE_BYTES);
}

// ------------------------------------------------------------------------------------------ generic
template <typename T>
__global__ __launch_bounds__(NT) void row_linear_kernel(case_rowlin_args_t a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw + RING_BYTES);   // [RB][K]
  float* red = xs + RB * a.K;                                    // [4][RB][256]
  const int r0 = blockIdx.x * RB, n0 = blockIdx.y * 256, tid = threadIdx.x;
  WStream ws;
  stream_setup(ws, smem_raw);
  ws.base[0] = ws.base[1] = ws.base[2] =
      reinterpret_cast<const char*>(a.Wt) + (size_t)blockIdx.y * a.K * 256 * sizeof(T);
  ws.total = a.K / KTile<T>::KT;
  ws.end[0] = ws.end[1] = ws.end[2] = ws.total;
  ws.start();          // weights are constants: their stream starts before the dependency wait
  pdl_trigger();
  pdl_wait();
  int off = 0;
  for (int s = 0; s < a.nseg; ++s) {
    const case_seg_t sg = a.seg[s];
    for (int i = tid; i < RB * sg.width; i += NT) {
      const int rb = i / sg.width, c = i - rb * sg.width, r = r0 + rb;
      float v = 0.f;
      if (r < a.R) {
        int rr = sg.gather ? a.gather_idx[r] : r;
        rr /= sg.div;
        v = sg.p[(size_t)rr * sg.ld + c];
      }
      xs[rb * a.K + off + c] = v;
    }
    off += sg.width;
  }
  __syncthreads();
  rows_linear_stream<T>(ws, xs, a.K, a.K, red);
  __syncthreads();
  const int n = n0 + tid;
  const float bias = a.bias ? __ldg(a.bias + n) : 0.f;
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int r = r0 + rb;
    if (r >= a.R) break;
    float y = red_sum(red, rb, tid) + bias;
    if (a.act == 1) y = gelu_erf(y);
    if (a.res) y += a.res[(size_t)r * a.ldres + n];
    a.out[(size_t)r * a.ldo + n] = y;
  }
}

// ------------------------------------------------------------------------------------------ layer front
constexpr int FRONT_FLOATS = RB * H + RB * 3 * H + RB * H + 4 * RB * 256 + (NT / 32) * CASE_MAX_T;
constexpr int BACK_FLOATS = 3 * RB * H + 4 * RB * 256;

template <typename T>
__global__ __launch_bounds__(NT) void layer_front_kernel(const float* __restrict__ h, case_layer_weights_t w,
                                                         T* kc, T* vc,
                                                         const int32_t* __restrict__ anc, int anc_ld,
                                                         const int32_t* __restrict__ tok, int tok_ld, int t,
                                                         int Tmax, float* __restrict__ b_out,
                                                         float* __restrict__ q2_out, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* fl = reinterpret_cast<float*>(smem_raw + RING_BYTES);
  float (*xs)[H] = reinterpret_cast<float (*)[H]>(fl);                          // h -> a = LN1(h)
  float (*qkv)[3 * H] = reinterpret_cast<float (*)[3 * H]>(fl + RB * H);        // q | k | v, later h1 -> b
  float (*cs)[H] = reinterpret_cast<float (*)[H]>(fl + RB * H + RB * 3 * H);    // self-attention context
  float* red = fl + RB * H + RB * 3 * H + RB * H;
  float (*sc)[CASE_MAX_T] = reinterpret_cast<float (*)[CASE_MAX_T]>(red + 4 * RB * 256);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = blockIdx.x * RB;
  constexpr int TPM = H / KTile<T>::KT;     // tiles per 256x256 matrix

  WStream ws;
  stream_setup(ws, smem_raw);
  ws.base[0] = reinterpret_cast<const char*>(w.Wqkv_t);
  ws.base[1] = reinterpret_cast<const char*>(w.Wo_t);
  ws.base[2] = reinterpret_cast<const char*>(w.Wq2_t);
  ws.end[0] = 3 * TPM; ws.end[1] = 4 * TPM; ws.end[2] = 5 * TPM;
  ws.total = 5 * TPM;
  ws.start();          // weights are constants: their stream starts before the dependency wait
  pdl_trigger();
  pdl_wait();

  for (int i = tid; i < RB * H; i += NT) {
    const int rb = i / H, r = r0 + rb;
    xs[rb][i % H] = r < R ? h[(size_t)r * H + i % H] : 0.f;
  }
  __syncthreads();
  ln_rows_smem(&xs[0][0], H, w.ln1_g, w.ln1_b);
  __syncthreads();

  for (int tile = 0; tile < 3; ++tile) {
    rows_linear_stream<T>(ws, &xs[0][0], H, H, red);
    __syncthreads();
    const float bias = __ldg(w.bqkv + tile * 256 + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) qkv[rb][tile * 256 + tid] = red_sum(red, rb, tid) + bias;
    __syncthreads();
  }
  // newest K / V rows -> cache at (row, position t); re-read below through the cache so the
  // value used now is the (possibly bf16-rounded) one later steps will see.
  for (int i = tid; i < RB * H; i += NT) {
    const int rb = i / H, n = i % H, r = r0 + rb;
    if (r < R) {
      st1(kc + ((size_t)r * Tmax + t) * H + n, qkv[rb][H + n]);
      st1(vc + ((size_t)r * Tmax + t) * H + n, qkv[rb][2 * H + n]);
    }
  }
  __syncthreads();

  // self-attention: one warp per (row, head); keys = positions 0..t of the row's ancestry
  for (int p = warp; p < RB * NH; p += NT / 32) {
    const int rb = p / NH, hh = p % NH, r = r0 + rb;
    if (r >= R) continue;
    const float* q = &qkv[rb][hh * HD];
    float m = -INFINITY;
    for (int j = lane; j <= t; j += 32) {
      const int pr = (j == t) ? r : anc[(size_t)r * anc_ld + j];
      const bool valid = tok[(size_t)pr * tok_ld + j] != 0;
      const T* kp = kc + ((size_t)pr * Tmax + j) * H + hh * HD;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 8) {
        float kv[8];
        ld8c(kp + d, kv);
#pragma unroll
        for (int u = 0; u < 8; ++u) s = fmaf(q[d + u], kv[u], s);
      }
      s = valid ? s : -INFINITY;
      sc[warp][j] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j <= t; j += 32) {
      const float s = sc[warp][j];
      const float e = (s == -INFINITY) ? 0.f : fexp(s - m);
      sc[warp][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int j = 0; j <= t; ++j) {
      const int pr = (j == t) ? r : anc[(size_t)r * anc_ld + j];
      acc = fmaf(sc[warp][j], ld1c(vc + ((size_t)pr * Tmax + j) * H + hh * HD + lane), acc);
    }
    cs[rb][hh * HD + lane] = sum > 0.f ? acc / sum : 0.f;
    __syncwarp();
  }
  __syncthreads();

  // h1 = a + c.Wo + bo   (residual on the normalised tensor, TransformerDecoder.py:76-79)
  rows_linear_stream<T>(ws, &cs[0][0], H, H, red);
  __syncthreads();
  {
    const float bias = __ldg(w.bo + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) qkv[rb][tid] = xs[rb][tid] + red_sum(red, rb, tid) + bias;
  }
  __syncthreads();
  ln_rows_smem(&qkv[0][0], 3 * H, w.ln2_g, w.ln2_b);
  __syncthreads();
  rows_linear_stream<T>(ws, &qkv[0][0], 3 * H, H, red);
  __syncthreads();
  {
    const float bias = __ldg(w.bq2 + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int r = r0 + rb;
      if (r < R) {
        q2_out[(size_t)r * H + tid] = red_sum(red, rb, tid) + bias;
        b_out[(size_t)r * H + tid] = qkv[rb][tid];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ layer back
template <typename T>
__global__ __launch_bounds__(NT) void layer_back_kernel(const float* __restrict__ b_in,
                                                        const float* __restrict__ part_ml,
                                                        const float* __restrict__ part_acc, int nsplit,
                                                        case_layer_weights_t w, float* __restrict__ h_out, int R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* fl = reinterpret_cast<float*>(smem_raw + RING_BYTES);
  float (*bs)[H] = reinterpret_cast<float (*)[H]>(fl);
  float (*cs)[H] = reinterpret_cast<float (*)[H]>(fl + RB * H);
  float (*ys)[H] = reinterpret_cast<float (*)[H]>(fl + 2 * RB * H);
  float* red = fl + 3 * RB * H;
  const int tid = threadIdx.x, r0 = blockIdx.x * RB;
  const int hh = tid / HD, d = tid % HD;
  constexpr int TPM = H / KTile<T>::KT;

  WStream ws;
  stream_setup(ws, smem_raw);
  ws.base[0] = reinterpret_cast<const char*>(w.Wo2_t);
  ws.base[1] = reinterpret_cast<const char*>(w.W1_t);
  ws.base[2] = reinterpret_cast<const char*>(w.W2_t);
  ws.end[0] = TPM; ws.end[1] = 2 * TPM; ws.end[2] = 3 * TPM;
  ws.total = 3 * TPM;
  ws.start();          // weights are constants: their stream starts before the dependency wait
  pdl_trigger();
  pdl_wait();

#pragma unroll
  for (int rb = 0; rb < RB; ++rb) {
    const int r = r0 + rb;
    float c = 0.f, bv = 0.f;
    if (r < R) {
      bv = b_in[(size_t)r * H + tid];
      const size_t base = ((size_t)r * NH + hh) * nsplit;
      float M = -INFINITY;
      for (int j = 0; j < nsplit; ++j) M = fmaxf(M, part_ml[(base + j) * 2]);
      float Z = 0.f, a = 0.f;
      for (int j = 0; j < nsplit; ++j) {
        const float mj = part_ml[(base + j) * 2];
        const float e = (mj == -INFINITY) ? 0.f : fexp(mj - M);
        Z = fmaf(part_ml[(base + j) * 2 + 1], e, Z);
        a = fmaf(part_acc[(base + j) * HD + d], e, a);
      }
      c = Z > 0.f ? a / Z : 0.f;
    }
    bs[rb][tid] = bv;
    cs[rb][tid] = c;
  }
  __syncthreads();
  rows_linear_stream<T>(ws, &cs[0][0], H, H, red);
  __syncthreads();
  {
    const float bias = __ldg(w.bo2 + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) ys[rb][tid] = bs[rb][tid] + red_sum(red, rb, tid) + bias;
  }
  __syncthreads();
  ln_rows_smem(&ys[0][0], H, w.ln3_g, w.ln3_b);
  __syncthreads();
  rows_linear_stream<T>(ws, &ys[0][0], H, H, red);
  __syncthreads();
  {
    const float bias = __ldg(w.b1 + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) cs[rb][tid] = gelu_erf(red_sum(red, rb, tid) + bias);
  }
  __syncthreads();
  rows_linear_stream<T>(ws, &cs[0][0], H, H, red);
  __syncthreads();
  {
    const float bias = __ldg(w.b2 + tid);
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int r = r0 + rb;
      if (r < R) h_out[(size_t)r * H + tid] = ys[rb][tid] + red_sum(red, rb, tid) + bias;
    }
  }
}

// ------------------------------------------------------------------------------------------ finalisers
__device__ __forceinline__ float block_sum_256(float v, float* sh /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += sh[i];
  return s;
}

// merge the per-split partials of one additive attention for row r, column n (DV columns looped by caller)
struct MergeStat { float e[CASE_MAX_SPLIT]; float Z, Q, M; };
__device__ __forceinline__ void merge_stats(const float* __restrict__ stats, int r, int nsplit, MergeStat& ms) {
  float M = -INFINITY;
  for (int j = 0; j < nsplit; ++j) M = fmaxf(M, stats[((size_t)r * nsplit + j) * 4]);
  ms.Z = 0.f; ms.Q = 0.f; ms.M = M;
  for (int j = 0; j < nsplit; ++j) {
    const float* s = stats + ((size_t)r * nsplit + j) * 4;
    const float e = (s[0] == -INFINITY) ? 0.f : fexp(s[0] - M);
    ms.e[j] = e;
    ms.Z = fmaf(s[1], e, ms.Z);
    ms.Q = fmaf(s[2], e, ms.Q);
  }
}

__global__ __launch_bounds__(NT) void finalize_rows_kernel(
    const float* __restrict__ h, const float* __restrict__ g, const float* __restrict__ b,
    const float* __restrict__ stats0, const float* __restrict__ ctxp0, int ns0,
    const float* __restrict__ stats1, const float* __restrict__ ctxp1, int ns1,
    const float* __restrict__ Wm, const float* __restrict__ bm, float* __restrict__ hN,
    float* __restrict__ ctx0, float* __restrict__ ctx1, float* __restrict__ gates, float* __restrict__ fac) {
  __shared__ float sh[8];
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, n = threadIdx.x;
  const float x = h[(size_t)r * H + n];
  const float mean = block_sum_256(x, sh) * (1.f / H);
  const float dx = x - mean;
  const float var = block_sum_256(dx * dx, sh) * (1.f / H);
  const float y = dx * rsqrtf(var + LN_EPS) * __ldg(g + n) + __ldg(b + n);
  hN[(size_t)r * H + n] = y;

  MergeStat m0, m1;
  merge_stats(stats0, r, ns0, m0);
  merge_stats(stats1, r, ns1, m1);
  float c0 = 0.f, c1 = 0.f;
  for (int j = 0; j < ns0; ++j) c0 = fmaf(ctxp0[((size_t)r * ns0 + j) * H + n], m0.e[j], c0);
  for (int j = 0; j < ns1; ++j) c1 = fmaf(ctxp1[((size_t)r * ns1 + j) * H + n], m1.e[j], c1);
  c0 = m0.Z > 0.f ? c0 / m0.Z : 0.f;
  c1 = m1.Z > 0.f ? c1 / m1.Z : 0.f;
  ctx0[(size_t)r * H + n] = c0;
  ctx1[(size_t)r * H + n] = c1;

  float lg[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* wr = Wm + (size_t)c * 3 * H;
    float part = fmaf(__ldg(wr + n), y, fmaf(__ldg(wr + H + n), c0, __ldg(wr + 2 * H + n) * c1));
    lg[c] = block_sum_256(part, sh) + __ldg(bm + c);
  }
  const float mx = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
  const float e0 = expf(lg[0] - mx), e1 = expf(lg[1] - mx), e2 = expf(lg[2] - mx);
  const float inv = 1.f / (e0 + e1 + e2);
  const float g0 = e0 * inv, g1 = e1 * inv, g2 = e2 * inv;
  if (n == 0) {
    gates[(size_t)r * 4 + 0] = g0; gates[(size_t)r * 4 + 1] = g1;
    gates[(size_t)r * 4 + 2] = g2; gates[(size_t)r * 4 + 3] = 0.f;
  }
  // copy weight(r,i,s) = F_i * prior_i[s] * exp(e_i[s] - M_i)  ==  gate_{i+1} * (w a) / (1e-8 + sum w a)
  // (Model.py:110-111,42) with a = softmax(e): F_i = gate_{i+1} / (Z_i * (1e-8 + Q_i / Z_i))
  if (n == 0) {
    float* f = fac + (size_t)r * 2 * CASE_MAX_SPLIT;
    f[0] = m0.Z > 0.f ? g1 / (m0.Z * (1e-8f + m0.Q / m0.Z)) : 0.f;
    f[1] = m0.Z > 0.f ? m0.M : 0.f;
    f[CASE_MAX_SPLIT] = m1.Z > 0.f ? g2 / (m1.Z * (1e-8f + m1.Q / m1.Z)) : 0.f;
    f[CASE_MAX_SPLIT + 1] = m1.Z > 0.f ? m1.M : 0.f;
  }
}

__global__ void attn_merge_kernel(const float* __restrict__ stats, const float* __restrict__ ctxp, int nsplit,
                                  int DV, float* __restrict__ ctx, float* __restrict__ fac, int fac_ld) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x;
  MergeStat ms;
  merge_stats(stats, r, nsplit, ms);
  for (int n = threadIdx.x; n < DV; n += blockDim.x) {
    float c = 0.f;
    for (int j = 0; j < nsplit; ++j) c = fmaf(ctxp[((size_t)r * nsplit + j) * DV + n], ms.e[j], c);
    ctx[(size_t)r * DV + n] = ms.Z > 0.f ? c / ms.Z : 0.f;
  }
  if (fac && threadIdx.x == 0) {     // normalised attention(r, s) = fac[0] * exp(e[s] - fac[1])
    fac[(size_t)r * fac_ld] = ms.Z > 0.f ? 1.f / ms.Z : 0.f;
    fac[(size_t)r * fac_ld + 1] = ms.Z > 0.f ? ms.M : 0.f;
  }
}

__global__ void gttp_gates_kernel(const float* __restrict__ f, const float* __restrict__ wc,
                                  const float* __restrict__ bc, float* __restrict__ gates,
                                  float* __restrict__ fac, int fac_ld, int nsplit) {
  __shared__ float sh[8];
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, n = threadIdx.x;
  const float z = block_sum_256(f[(size_t)r * H + n] * __ldg(wc + n), sh) + __ldg(bc);
  const float pc = 1.f / (1.f + expf(-z));
  if (n == 0) {
    gates[(size_t)r * 4 + 0] = 1.f - pc; gates[(size_t)r * 4 + 1] = pc;
    gates[(size_t)r * 4 + 2] = 0.f; gates[(size_t)r * 4 + 3] = 0.f;
  }
  if (n == 0) fac[(size_t)r * fac_ld] *= pc;
}

__global__ void gru_cell_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                const float* __restrict__ hp, const int32_t* __restrict__ gidx,
                                float* __restrict__ ho) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, n = threadIdx.x;
  const float* a = gi + (size_t)r * 3 * H;
  const float* c = gh + (size_t)r * 3 * H;
  const int src = gidx ? gidx[r] : r;
  const float rg = 1.f / (1.f + expf(-(a[n] + c[n])));
  const float zg = 1.f / (1.f + expf(-(a[H + n] + c[H + n])));
  const float ng = tanhf(a[2 * H + n] + rg * c[2 * H + n]);
  ho[(size_t)r * H + n] = (1.f - zg) * ng + zg * hp[(size_t)src * H + n];
}

}  // namespace cb

// =============================================================================== C ABI
using namespace cb;

extern "C" int case_embed_rows(const float* E, const float* pe, const int32_t* tok, int tok_ld, int t,
                               float scale, float* x, int R, case_stream_t stream) {
  CB_REQUIRE(E && tok && x && R > 0 && t >= 0, "case_embed_rows: bad arguments");
  launch_k(embed_kernel, (R + 3) / 4, 256, 0, (cudaStream_t)stream, E, pe, tok, tok_ld, t, scale, x, R);
  return check_launch("case_embed_rows");
}

extern "C" int case_layernorm_rows(const float* x, const float* g, const float* b, float* y, int R,
                                   case_stream_t stream) {
  CB_REQUIRE(x && g && b && y && R > 0, "case_layernorm_rows: bad arguments");
  launch_k(layernorm_rows_kernel, (R + RB - 1) / RB, NT, 0, (cudaStream_t)stream, x, g, b, y, R);
  return check_launch("case_layernorm_rows");
}

int case_row_linear_tc(const case_rowlin_args_t* a, cudaStream_t st);
int case_layer_front_tc(const float* h, const case_layer_weights_t* w, void* kcache, void* vcache, const int32_t* anc,
                        int anc_ld, const int32_t* tok, int tok_ld, int t, int Tmax, float* b_out, float* q2_out, int R,
                        cudaStream_t st);
int case_layer_back_tc(const float* b_in, const float* part_ml, const float* part_acc, int nsplit,
                       const case_layer_weights_t* w, float* h_out, int R, cudaStream_t st);

extern "C" int case_row_linear(const case_rowlin_args_t* a, case_stream_t stream) {
  CB_REQUIRE(a && a->Wt && a->out && a->R > 0, "case_row_linear: null pointer");
  CB_REQUIRE(a->nseg >= 1 && a->nseg <= 4, "case_row_linear: nseg must be 1..4");
  int K = 0;
  for (int s = 0; s < a->nseg; ++s) {
    CB_REQUIRE(a->seg[s].p && a->seg[s].div >= 1 && a->seg[s].width > 0, "case_row_linear: bad segment");
    CB_REQUIRE(!a->seg[s].gather || a->gather_idx, "case_row_linear: gather without gather_idx");
    K += a->seg[s].width;
  }
  CB_REQUIRE(K == a->K && K % 32 == 0 && K <= 2048, "case_row_linear: K must equal the segment widths, %32, <=2048");
  CB_REQUIRE(a->N % 256 == 0 && a->N > 0, "case_row_linear: N must be a multiple of 256");
  if (a->dtype == CASE_BF16) return case_row_linear_tc(a, (cudaStream_t)stream);
  const size_t smem = RING_BYTES + (size_t)(RB * a->K + 4 * RB * 256) * sizeof(float);
  dim3 grid((a->R + RB - 1) / RB, a->N / 256);
  ensure_smem<row_linear_kernel<bf16>>(200 * 1024);
  ensure_smem<row_linear_kernel<float>>(200 * 1024);
  if (a->dtype == CASE_BF16) {
    launch_k(row_linear_kernel<bf16>, grid, NT, smem, (cudaStream_t)stream, *a);
  } else {
    launch_k(row_linear_kernel<float>, grid, NT, smem, (cudaStream_t)stream, *a);
  }
  return check_launch("case_row_linear");
}

extern "C" int case_layer_front(const float* h, const case_layer_weights_t* w, void* kcache, void* vcache,
                                const int32_t* anc, int anc_ld, const int32_t* tok, int tok_ld, int t, int Tmax,
                                float* b_out, float* q2_out, int R, int dtype, case_stream_t stream) {
  CB_REQUIRE(h && w && kcache && vcache && anc && tok && b_out && q2_out && R > 0, "case_layer_front: null pointer");
  CB_REQUIRE(t >= 0 && t < Tmax && Tmax <= CASE_MAX_T, "case_layer_front: t / Tmax out of range");
  if (dtype == CASE_BF16)
    return case_layer_front_tc(h, w, kcache, vcache, anc, anc_ld, tok, tok_ld, t, Tmax, b_out, q2_out, R,
                               (cudaStream_t)stream);
  const int grid = (R + RB - 1) / RB;
  const size_t smem = RING_BYTES + (size_t)FRONT_FLOATS * sizeof(float);
  ensure_smem<layer_front_kernel<bf16>>(200 * 1024);
  ensure_smem<layer_front_kernel<float>>(200 * 1024);
  if (dtype == CASE_BF16)
    launch_k(layer_front_kernel<bf16>, grid, NT, smem, (cudaStream_t)stream, h, *w, (bf16*)kcache, (bf16*)vcache, anc, anc_ld,
                                                                     tok, tok_ld, t, Tmax, b_out, q2_out, R);
  else
    launch_k(layer_front_kernel<float>, grid, NT, smem, (cudaStream_t)stream, h, *w, (float*)kcache, (float*)vcache, anc,
                                                                      anc_ld, tok, tok_ld, t, Tmax, b_out, q2_out, R);
  return check_launch("case_layer_front");
}

extern "C" int case_layer_back(const float* b_in, const float* part_ml, const float* part_acc, int nsplit,
                               const case_layer_weights_t* w, float* h_out, int R, int dtype,
                               case_stream_t stream) {
  CB_REQUIRE(b_in && part_ml && part_acc && w && h_out && R > 0, "case_layer_back: null pointer");
  CB_REQUIRE(nsplit >= 1 && nsplit <= CASE_MAX_XSPLIT, "case_layer_back: nsplit out of range");
  if (dtype == CASE_BF16) return case_layer_back_tc(b_in, part_ml, part_acc, nsplit, w, h_out, R, (cudaStream_t)stream);
  const int grid = (R + RB - 1) / RB;
  const size_t smem = RING_BYTES + (size_t)BACK_FLOATS * sizeof(float);
  ensure_smem<layer_back_kernel<bf16>>(200 * 1024);
  ensure_smem<layer_back_kernel<float>>(200 * 1024);
  if (dtype == CASE_BF16)
    launch_k(layer_back_kernel<bf16>, grid, NT, smem, (cudaStream_t)stream, b_in, part_ml, part_acc, nsplit, *w, h_out, R);
  else
    launch_k(layer_back_kernel<float>, grid, NT, smem, (cudaStream_t)stream, b_in, part_ml, part_acc, nsplit, *w, h_out, R);
  return check_launch("case_layer_back");
}

extern "C" int case_finalize_rows(const float* h, const float* lnN_g, const float* lnN_b, const float* stats0,
                                  const float* ctxp0, int nsplit0, const float* stats1, const float* ctxp1,
                                  int nsplit1, const float* Wm, const float* bm, float* hN, float* ctx0,
                                  float* ctx1, float* gates, float* fac, int R, case_stream_t stream) {
  CB_REQUIRE(h && lnN_g && lnN_b && stats0 && ctxp0 && stats1 && ctxp1 && Wm && bm && hN && ctx0 && ctx1 &&
                 gates && fac && R > 0, "case_finalize_rows: null pointer");
  CB_REQUIRE(nsplit0 >= 1 && nsplit0 <= CASE_MAX_SPLIT && nsplit1 >= 1 && nsplit1 <= CASE_MAX_SPLIT,
             "case_finalize_rows: nsplit out of range");
  launch_k(finalize_rows_kernel, R, NT, 0, (cudaStream_t)stream, h, lnN_g, lnN_b, stats0, ctxp0, nsplit0, stats1, ctxp1,
                                                            nsplit1, Wm, bm, hN, ctx0, ctx1, gates, fac);
  return check_launch("case_finalize_rows");
}

extern "C" int case_attn_merge(const float* stats, const float* ctx_part, int nsplit, int DV, float* ctx,
                               float* fac, int fac_ld, int R, case_stream_t stream) {
  CB_REQUIRE(stats && ctx_part && ctx && R > 0 && DV > 0, "case_attn_merge: bad arguments");
  CB_REQUIRE(nsplit >= 1 && nsplit <= CASE_MAX_SPLIT, "case_attn_merge: nsplit out of range");
  launch_k(attn_merge_kernel, R, 256, 0, (cudaStream_t)stream, stats, ctx_part, nsplit, DV, ctx, fac, fac_ld);
  return check_launch("case_attn_merge");
}

extern "C" int case_gttp_gates(const float* f, const float* wc, const float* bc, float* gates, float* fac,
                               int fac_ld, int nsplit, int R, case_stream_t stream) {
  CB_REQUIRE(f && wc && bc && gates && fac && R > 0, "case_gttp_gates: bad arguments");
  launch_k(gttp_gates_kernel, R, NT, 0, (cudaStream_t)stream, f, wc, bc, gates, fac, fac_ld, nsplit);
  return check_launch("case_gttp_gates");
}

extern "C" int case_gru_cell(const float* gi, const float* gh, const float* h_prev, const int32_t* gather_idx,
                             float* h_out, int R, case_stream_t stream) {
  CB_REQUIRE(gi && gh && h_prev && h_out && R > 0, "case_gru_cell: bad arguments");
  launch_k(gru_cell_kernel, R, NT, 0, (cudaStream_t)stream, gi, gh, h_prev, gather_idx, h_out);
  return check_launch("case_gru_cell");
}
