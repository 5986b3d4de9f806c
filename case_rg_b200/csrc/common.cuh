// Shared device helpers for the case_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/case_b200.h"

namespace cb {

constexpr int H = CASE_H;
constexpr int NH = CASE_NH;
constexpr int HD = CASE_HD;
constexpr float LN_EPS = 1e-5f;

void set_error(const char* msg);
int check_launch(const char* what);

// Per-THREAD launch options (step.cu): nothing in this library is process-global.  The step orchestrators install the
// options of their argument block (case_step_args_t.opt) for the duration of the call; direct launcher calls use the
// calling thread's options (case_thread_options, default: everything on).
struct LaunchOpts {
  int pdl;           // programmatic dependent launch attribute on every kernel launch
  int evict_first;   // the once-per-step K|V / Uk.mem streams are loaded with an L2 evict-first policy
};
LaunchOpts& launch_opts();
cudaError_t& launch_err();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: set it once per (kernel
// instantiation, device) - not once per process, which would leave every device but the first at the 48 KB default.
template <auto Kern>
inline void ensure_smem(int bytes) {
  static unsigned long long done = 0;          // one static per kernel; bit d = device d is set up
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(done & bit)) {
    cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    done |= bit;                               // (a racing second thread repeats the same idempotent call)
  }
}

// Launch helper: same as kernel<<<grid, block, smem, stream>>>(args...) plus the programmatic
// stream-serialization attribute, so the next kernel's CTAs may be scheduled (and run their
// prologue: barrier init, weight prefetch) while this kernel drains.  Every kernel calls
// pdl_wait() before it touches anything an earlier kernel wrote.
template <typename T> struct ident { typedef T type; };
template <typename... KA>
inline void launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     typename ident<KA>::type... args) {
  void* pa[] = {(void*)&args...};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = launch_opts().pdl ? 1 : 0;
  launch_err() = cudaLaunchKernelExC(&cfg, (const void*)kern, pa);
}

#define CB_REQUIRE(cond, msg)        \
  do {                               \
    if (!(cond)) {                   \
      cb::set_error(msg);            \
      return CASE_EINVAL;            \
    }                                \
  } while (0)

typedef __nv_bfloat16 bf16;

// key range of one split: ceil(S / nsplit) rounded up to a whole number of tiles
__host__ __device__ inline int split_chunk(int S, int nsplit, int tile) {
  const int c = (S + nsplit - 1) / nsplit;
  return (c + tile - 1) / tile * tile;
}
constexpr int XATTN_TILE = 128;   // cross-attention keys per tile
constexpr int AATTN_TILE = 128;   // additive-attention split granularity

// ---- storage-type loads (fp32 or bf16 -> fp32 registers), vector forms
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void ld4(const float* p, float (&o)[4]) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&o)[4]) {
  uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ void ld2(const float* p, float (&o)[2]) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  o[0] = v.x; o[1] = v.y;
}
__device__ __forceinline__ void ld2(const bf16* p, float (&o)[2]) {
  uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
  o[0] = __uint_as_float(v << 16); o[1] = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ void ld8(const float* p, float (&o)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&o)[8]) {
  uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  o[4] = __uint_as_float(v.z << 16); o[5] = __uint_as_float(v.z & 0xffff0000u);
  o[6] = __uint_as_float(v.w << 16); o[7] = __uint_as_float(v.w & 0xffff0000u);
}

// coherent (non-__ldg) forms for buffers written earlier in the same kernel
__device__ __forceinline__ float ld1c(const float* p) { return *p; }
__device__ __forceinline__ float ld1c(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void ld8c(const float* p, float (&o)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void ld8c(const bf16* p, float (&o)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
  o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  o[4] = __uint_as_float(v.z << 16); o[5] = __uint_as_float(v.z & 0xffff0000u);
  o[6] = __uint_as_float(v.w << 16); o[7] = __uint_as_float(v.w & 0xffff0000u);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// programmatic dependent launch: let the dependent grid start early / wait for the producer grid
// (an explicit early trigger made the step slower: dependents that are resident but blocked in
//  pdl_wait() take shared memory / CTA slots from multi-wave producers such as the cross-attention,
//  so the trigger is left implicit - it fires as each producer CTA exits)
#ifdef CASE_PDL_EARLY_TRIGGER
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() {}
#endif
// L2 cache policy of a stream that is read once per step and is larger than L2 (cross-attention K|V, Uk.mem): with
// evict-first its lines do not push out what the step re-reads (layer weights, the vocabulary weight, the logits tile,
// the self-attention history)
__device__ __forceinline__ uint64_t l2_stream_policy(int evict_first) {
  uint64_t pol;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exp with ex2.approx (rel. error ~2^-22 * |x|); arguments here are always <= 0 after max-shift.
__device__ __forceinline__ float fexp(float x) { return exp2f(x * 1.4426950408889634f); }

// tanh, two flavours: 1 MUFU (tanh.approx, abs err ~5e-4) or ex2+rcp based (abs err ~2e-7).
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_acc(float x) {
  float ax = fminf(fabsf(x), 15.f);
  float e = exp2f(ax * 2.8853900817779268f);  // exp(2|x|)
  float r = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(r, x);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

}  // namespace cb
