// Tuning probe (not part of the product): how fast can a persistent kernel stream a
// byte table from HBM on B200, (a) through a ring of TMA bulk copies into shared
// memory, (b) through plain 128-bit loads into registers?  Varies tile size, ring
// depth, CTA size and CTAs per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <functional>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void k_tma(const uint8_t *src, uint64_t bytes, int tile, int stages, unsigned long long *sink, int touch, int shift, int extra)
{
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  unsigned char *buf = smem + 128;
  const int tid = threadIdx.x;
  const uint64_t ntiles = bytes / tile;
  if (tid == 0)
  {
    for (int s = 0; s < stages; s++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](uint64_t t, int s)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[s])), "r"(tile + extra) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(buf + (size_t) s * (tile + extra))), "l"(src + t * tile + shift), "r"(tile + extra), "r"(smem_u32(&bars[s])) : "memory");
  };
  if (tid == 0)
    for (int j = 0; j < stages - 1; j++)
      if (blockIdx.x + (uint64_t) j * gridDim.x < ntiles) issue(blockIdx.x + (uint64_t) j * gridDim.x, j);
  uint32_t acc = 0;
  uint32_t it = 0;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, it++)
  {
    const int s = it % stages;
    if (tid == 0 && t + (uint64_t) (stages - 1) * gridDim.x < ntiles)
      issue(t + (uint64_t) (stages - 1) * gridDim.x, (it + stages - 1) % stages);
    const uint32_t parity = (it / stages) & 1;
    uint32_t done;
    do
    {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(parity) : "memory");
    } while (!done);
    if (touch)
      for (int o = tid * 16; o < tile; o += blockDim.x * 16)
      {
        const uint4 v = *reinterpret_cast<const uint4 *>(buf + (size_t) s * (tile + extra) + o);
        acc += v.x ^ v.y ^ v.z ^ v.w;
      }
    __syncthreads();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int UNROLL>
__global__ void k_ldg(const uint4 *src, uint64_t n16, unsigned long long *sink)
{
  uint32_t acc = 0;
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride)
  {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

static float time_it(std::function<void()> f, void *flush, size_t flush_bytes)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  std::vector<float> ts;
  for (int r = 0; r < 8; r++)
  {
    CK(cudaMemsetAsync(flush, r, flush_bytes));
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

#include <functional>
int main(int argc, char **argv)
{
  const uint64_t bytes = argc > 1 ? strtoull(argv[1], 0, 10) : 400000000ull;
  uint8_t *d; unsigned long long *sink; void *flush;
  const uint64_t alloc = (bytes + (1 << 20)) & ~((1ull << 20) - 1);
  CK(cudaMalloc(&d, alloc)); CK(cudaMemset(d, 1, alloc));
  CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&flush, 512u << 20));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("bytes %llu, %d SMs\n", (unsigned long long) bytes, sms);
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int tiles[] = {16384};
  for (int shift : {0, 16, 112})
   for (int extra : {0, 32})
    for (int tile : tiles)
      for (int stages : {3})
        for (int threads : {512})
          for (int cps : {2})
          {
            const int touch = 1;
            const size_t smem = 128 + (size_t) (tile + extra) * stages;
            if (smem * cps > 220 * 1024 || threads * cps > 2048) continue;
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tma, threads, smem));
            if (occ < cps) continue;
            const int grid = sms * cps;
            float ms = time_it([&] { k_tma<<<grid, threads, smem>>>(d, bytes - 65536, tile, stages, sink, touch, shift, extra); }, flush, 512u << 20);
            CK(cudaGetLastError());
            printf("tma shift=%d extra=%d touch=%d tile=%5d stages=%d threads=%d ctas/sm=%d : %7.1f us  %6.0f GB/s\n", shift, extra, touch, tile, stages,
                   threads, cps, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
          }
  for (int threads : {256, 512})
    for (int cps : {2, 4, 8})
    {
      if (threads * cps > 2048) continue;
      const int grid = sms * cps;
      float m1 = time_it([&] { k_ldg<1><<<grid, threads>>>((const uint4 *) d, bytes / 16, sink); }, flush, 512u << 20);
      float m4 = time_it([&] { k_ldg<4><<<grid, threads>>>((const uint4 *) d, bytes / 16, sink); }, flush, 512u << 20);
      float m8 = time_it([&] { k_ldg<8><<<grid, threads>>>((const uint4 *) d, bytes / 16, sink); }, flush, 512u << 20);
      printf("ldg threads=%d ctas/sm=%d : unroll1 %6.0f GB/s  unroll4 %6.0f GB/s  unroll8 %6.0f GB/s\n", threads, cps,
             bytes / (m1 * 1e-3) / 1e9, bytes / (m4 * 1e-3) / 1e9, bytes / (m8 * 1e-3) / 1e9);
    }
  return 0;
}
