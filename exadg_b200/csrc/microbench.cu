// FP64 pipe microbenchmarks: the denominators for the "FP64 pipe utilisation" evidence next to the
// HBM roofline (MEASURED_PEAKS.json has HBM and bf16 only; the affine vmult is FP64-bound).
//   dfma: register-resident chains of fma.rn.f64 (8 independent accumulators per thread)
//   dmma: mma.sync.aligned.m8n8k4.row.col.f64 (legacy FP64 tensor path; tcgen05 has no FP64 kind)
#include "operator.cuh"

namespace exadg_b200
{
namespace
{
__global__ void __launch_bounds__(256) dfma_kernel(double * out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) dmma_kernel(double * out, int iters, double a, double b)
{
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  const double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
} // namespace

void fp64_peak(double * dfma_tflops, double * dmma_tflops)
{
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  double * out = nullptr;
  CUDA_CHECK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) { // last repetition counts (warm clocks)
    CUDA_CHECK(cudaEventRecord(e0));
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
    CUDA_CHECK(cudaEventRecord(e1)); CUDA_CHECK(cudaEventSynchronize(e1));
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  }
  if (dfma_tflops) *dfma_tflops = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_CHECK(cudaEventRecord(e0));
    dmma_kernel<<<blocks, threads>>>(out, iters, 0.5, 0.25);
    CUDA_CHECK(cudaEventRecord(e1)); CUDA_CHECK(cudaEventSynchronize(e1));
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  }
  // per warp and mma: 2*8*8*4 = 512 flop; 16 mma per iteration
  if (dmma_tflops) *dmma_tflops = 512.0 * 16.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
}

} // namespace exadg_b200
