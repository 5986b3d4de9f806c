// Micro-benchmark: can a FEW CTAs (one per query) pull HBM-resident streams at the chip's bandwidth?
// Each CTA streams its own distinct region (larger than L2 in aggregate) with W warps, each warp
// running a ring of `stages` bulk copies of `slab` bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_stream profiles/micro/hbm_stream_bench.cu && /tmp/hbm_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ bool mb_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok;
}

__global__ void stream(const char* src, size_t per_cta, int slab, int stages, float* sink) {
  extern __shared__ __align__(128) unsigned char sm0[];
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  unsigned char* sm = sm0 + (size_t)warp * ((size_t)slab * stages + 128);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)slab * stages);
  const size_t per_warp = per_cta / nw;
  const char* base = src + (size_t)blockIdx.x * per_cta + (size_t)warp * per_warp;
  const int total = (int)(per_warp / slab);
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < stages; ++s) mb_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto issue = [&](int p) { const int s = p % stages; mb_expect(&full[s], slab); bulk(sm + (size_t)s * slab, base + (size_t)p * slab, slab, &full[s]); };
    for (int p = 0; p < stages && p < total; ++p) issue(p);
    float acc = 0.f;
    for (int g = 0; g < total; ++g) {
      const int s = g % stages;
      while (!mb_try(&full[s], (g / stages) & 1)) {}
      acc += reinterpret_cast<float*>(sm + (size_t)s * slab)[g & 63];
      if (g + stages < total) issue(g + stages);
    }
    sink[blockIdx.x * 32 + warp] = acc;
  }
}

int main() {
  const size_t per_cta = 12ull << 20;            // 12 MB per CTA (about one query's K/V of a stack at C2)
  const int max_ctas = 148;
  char* src; float* sink;
  cudaMalloc(&src, per_cta * max_ctas); cudaMalloc(&sink, 148 * 32 * 4);
  cudaMemset(src, 1, per_cta * max_ctas);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("%6s %6s %7s %7s | %10s %14s %12s\n", "ctas", "warps", "slabKB", "stages", "us", "GB/s per CTA", "TB/s total");
  struct Cfg { int warps, slab, stages; } cfgs[] = {{8, 8192, 2}, {8, 8192, 3}, {8, 4096, 4}, {8, 16384, 1}, {12, 8192, 2}, {16, 4096, 3}, {4, 16384, 3}, {8, 12288, 2}};
  for (int ctas : {32, 64, 128, 148})
    for (auto c : cfgs) {
      const size_t smem = ((size_t)c.slab * c.stages + 128) * c.warps;
      if (smem > 220 * 1024) continue;
      stream<<<ctas, 32 * c.warps, smem>>>(src, per_cta, c.slab, c.stages, sink);   // warm-up (also evicts)
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      stream<<<ctas, 32 * c.warps, smem>>>(src, per_cta, c.slab, c.stages, sink);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double per = (double)per_cta / (ms * 1e-3) / 1e9;
      printf("%6d %6d %7d %7d | %10.1f %14.1f %12.2f\n", ctas, c.warps, c.slab / 1024, c.stages, ms * 1e3, per, per * ctas / 1e3);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
