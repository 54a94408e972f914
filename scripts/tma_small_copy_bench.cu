// Microbenchmark: throughput of small (1 KB) cp.async.bulk copies global -> shared from scattered, 16-byte aligned addresses,
// as a producer warp would issue them for the neighbour cells of a cell batch (k = 4: 1000 B per cell).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_small_copy_bench scripts/tma_small_copy_bench.cu
//   build/tma_small_copy_bench            (prints cycles per copy for several issuer counts / ring depths)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ISSUERS warps per CTA, lane 0 of each issues `n_copies` bulk copies of `bytes` into its own ring of DEPTH slots
template<int ISSUERS, int DEPTH>
__global__ void __launch_bounds__(256, 2) bench(const double * __restrict__ src, long long n_cells, int n_copies, uint32_t bytes, long long * cycles, double * sink)
{
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem);                 // [ISSUERS][DEPTH]
  unsigned char * ring = smem + 1024;                                   // [ISSUERS][DEPTH][1024]
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ISSUERS * DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  double acc = 0.0;
  if (w < ISSUERS && lane == 0) {
    uint64_t rng = 0x9E3779B97F4A7C15ull * (blockIdx.x * ISSUERS + w + 1);
    for (int i = 0; i < n_copies; ++i) {
      const int s = i % DEPTH;
      uint64_t * bar = bars + w * DEPTH + s;
      unsigned char * dst = ring + (size_t)(w * DEPTH + s) * 1024;
      if (i >= DEPTH) {
        const uint32_t parity = (uint32_t)((i / DEPTH - 1) & 1);
        uint32_t done;
        do {
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
        } while (!done);
        acc += *reinterpret_cast<double *>(dst + 8 * (i & 63)); // consume something
      }
      rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
      const long long cell = (long long)(rng % (unsigned long long)n_cells) & ~1ll; // even cells: 16-byte aligned
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src + cell * 125), "r"(bytes), "r"(s32(bar)) : "memory");
    }
    for (int s = 0; s < DEPTH && s < n_copies; ++s) { // drain
      const int last = ((n_copies - 1 - s) / DEPTH) * DEPTH + s; // last copy that used slot s... parity of its completion
      const uint32_t parity = (uint32_t)((last / DEPTH) & 1);
      uint64_t * bar = bars + w * DEPTH + s;
      uint32_t done;
      do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)), "r"(parity) : "memory");
      } while (!done);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456) sink[0] = acc;
}

template<int ISSUERS, int DEPTH>
void run(const double * src, long long n_cells, uint32_t bytes)
{
  const int grid = 148 * 2, n_copies = 4000;
  const size_t smem = 1024 + (size_t)ISSUERS * DEPTH * 1024;
  cudaFuncSetAttribute(bench<ISSUERS, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long * cycles; double * sink;
  cudaMalloc(&cycles, grid * sizeof(long long)); cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<ISSUERS, DEPTH><<<grid, 256, smem>>>(src, n_cells, 200, bytes, cycles, sink);
  cudaEventRecord(e0);
  bench<ISSUERS, DEPTH><<<grid, 256, smem>>>(src, n_cells, n_copies, bytes, cycles, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const cudaError_t err = cudaGetLastError();
  long long h[296]; cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
  const double total_copies = (double)grid * ISSUERS * n_copies;
  std::printf("issuers/CTA %d depth %2d bytes %4u: %.3f ms, %.1f cycles per copy per issuer, %.1f cycles per copy per SM (2 CTAs/SM), %.0f GB/s aggregate  [%s]\n", ISSUERS, DEPTH, bytes, ms,
              mean / n_copies, mean / (n_copies * ISSUERS * 2.0), total_copies * bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
  cudaFree(cycles); cudaFree(sink);
}

int main()
{
  const long long n_cells = 884736; // 96^3 cells of 125 doubles
  double * src; cudaMalloc(&src, (size_t)(n_cells + 2) * 1000);
  cudaMemset(src, 0, (size_t)(n_cells + 2) * 1000);
  for (uint32_t bytes : {1008u, 512u}) {
    run<1, 8>(src, n_cells, bytes); run<1, 16>(src, n_cells, bytes); run<1, 32>(src, n_cells, bytes);
    run<2, 16>(src, n_cells, bytes); run<4, 8>(src, n_cells, bytes); run<4, 16>(src, n_cells, bytes);
  }
  return 0;
}
