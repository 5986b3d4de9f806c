// Micro-benchmark: per-SM throughput of cp.async.bulk (global/L2 -> shared) rings, the mechanism the
// row-block kernels use to stream weights.  Varies slab size, ring depth, CTA count, whether all
// CTAs read the same bytes, and 1 vs 2 issuing threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_bench profiles/micro/bulk_bench.cu && /tmp/bulk_bench
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

// lane 0 of every warp drives its own ring (its own smem partition); `split` > 1 cuts each slab
// into that many copies
__global__ void ring(const char* src, size_t per_cta_stride, int slab, int stages, int total, int split, float* sink) {
  extern __shared__ __align__(128) unsigned char sm0[];
  const int warp = threadIdx.x >> 5;
  unsigned char* sm = sm0 + (size_t)warp * ((size_t)slab * stages + 256);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)slab * stages);
  const char* base = src + (size_t)blockIdx.x * per_cta_stride + (size_t)warp * 40 * slab;
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < stages; ++s) mb_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    auto issue = [&](int p) {
      const int s = p % stages;
      mb_expect(&full[s], slab);
      for (int q = 0; q < split; ++q)
        bulk(sm + (size_t)s * slab + (size_t)q * (slab / split), base + ((size_t)(p % 40) * slab) + (size_t)q * (slab / split), slab / split, &full[s]);
    };
    for (int p = 0; p < stages && p < total; ++p) issue(p);
    float acc = 0.f;
    for (int g = 0; g < total; ++g) {
      const int s = g % stages;
      while (!mb_try(&full[s], (g / stages) & 1)) {}
      acc += reinterpret_cast<float*>(sm + (size_t)s * slab)[g & 63];
      if (g + stages < total) issue(g + stages);
    }
    sink[blockIdx.x * 8 + warp] = acc;
  }
}

int main() {
  const size_t bytes = 160ull << 20;
  char* src; float* sink;
  cudaMalloc(&src, bytes); cudaMalloc(&sink, 4096);
  cudaMemset(src, 1, bytes);
  cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("%6s %6s %6s %6s %6s | %10s %12s %12s\n", "ctas", "slabKB", "stages", "shared", "split", "us", "GB/s per SM", "TB/s total");
  const int ctas_l[] = {1, 32, 148};
  struct Cfg { int slab, stages, split, warps; } cfgs[] = {{16384, 6, 1, 1}, {16384, 3, 1, 2}, {16384, 3, 1, 4}, {8192, 3, 1, 8}, {16384, 2, 1, 6},
                                                           {32768, 3, 1, 2}, {32768, 2, 1, 3}, {65536, 3, 1, 1}, {98304, 2, 1, 1}, {65536, 1, 1, 3}};
  printf("(columns: ctas slabKB stages shared split/warps)\n");
  for (int shared = 1; shared >= 1; --shared)
    for (int ci = 0; ci < 3; ++ci)
      for (auto c : cfgs) {
        const int ctas = ctas_l[ci];
        const int total = (4 << 20) / c.slab;                       // 4 MB per warp
        const size_t stride = 0;
        const size_t smem = ((size_t)c.slab * c.stages + 256) * c.warps;
        ring<<<ctas, 32 * c.warps, smem>>>(src, stride, c.slab, c.stages, 8, c.split, sink);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        ring<<<ctas, 32 * c.warps, smem>>>(src, stride, c.slab, c.stages, total, c.split, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double per = (double)total * c.slab * c.warps / (ms * 1e-3) / 1e9;
        printf("%6d %6d %6d %6d %6d | %10.1f %12.1f %12.2f\n", ctas, c.slab / 1024, c.stages, shared, c.warps, ms * 1e3, per, per * ctas / 1e3);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
