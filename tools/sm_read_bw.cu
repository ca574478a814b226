// Experiment: read-only HBM bandwidth as a function of how many SMs pull, with 1-D TMA bulk loads (global -> shared)
// of `chunk` bytes through a ring of `stages` buffers; nothing consumes the data.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/sm_read_bw.cu -o tools/sm_read_bw.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { auto e = (x); if (e != cudaSuccess) { printf("fail %s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(32, 1) pull(const unsigned char* __restrict__ src, long long bytes_per_cta, int chunk, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned char* p = src + static_cast<long long>(blockIdx.x) * bytes_per_cta;
    const long long n = bytes_per_cta / chunk;
    for (long long i = 0; i < n + stages; ++i) {
      const int s = static_cast<int>(i % stages);
      if (i >= stages) {  // wait for the load issued `stages` iterations ago
        const uint32_t parity = static_cast<uint32_t>(((i / stages) - 1) & 1);
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                       : "=r"(ok) : "r"(s32(&bars[s])), "r"(parity) : "memory");
      }
      if (i < n) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(ring + static_cast<size_t>(s) * chunk)), "l"(p + i * chunk), "r"(chunk), "r"(s32(&bars[s])) : "memory");
      }
    }
  }
}

int main() {
  const long long total = 2LL << 30;  // 2 GiB source, far beyond L2
  unsigned char* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  CK(cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grids[] = {19, 37, 74, 111, 148};
  const int chunks[] = {8192, 16384, 32768};
  for (int chunk : chunks)
    for (int in_flight_kb : {64, 128, 192})
      for (int g : grids) {
        const int stages = in_flight_kb * 1024 / chunk;
        const long long per_cta = (total / 148 / chunk) * chunk / 2;  // every CTA streams the same amount (~7 MiB)
        pull<<<g, 32, stages * chunk + 1024>>>(src, per_cta, chunk, stages);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) pull<<<g, 32, stages * chunk + 1024>>>(src + (r % 2) * (total / 2), per_cta, chunk, stages);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gbs = 5.0 * g * per_cta / (ms * 1e-3) / 1e9;
        printf("chunk %5d B, %3d KiB in flight/SM, %3d CTAs: %7.0f GB/s total, %6.1f GB/s per SM\n", chunk, in_flight_kb, g, gbs, gbs / g);
      }
  return 0;
}
