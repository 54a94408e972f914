// Kernel tournament for the affine fast path at k = 4 (C ABI only, no Python): times the pipelined kernel (variant 0) and
// the warp-specialised kernel (variants 1, 2, 3) on the same operator and vectors, and checks that both give the same vmult /
// vmult_add.  Usage: ws_tournament only_variant|-1 (n_sub refine steps)...   (default: -1 3 5 100 = 96^3 cells, 100 steps)
//   nvcc -O2 -std=c++17 scripts/ws_tournament.cpp -Iinclude -Lexadg_b200 -lexadg_b200 -Xlinker -rpath='$ORIGIN/../exadg_b200' -o build/ws_tournament
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "exadg_b200.h"

#define CK(x) do { auto e_ = (x); if (e_ != 0) { std::printf("FAILED %s -> %d (%s)\n", #x, (int)e_, exadg_b200_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA FAILED %s -> %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

static double rel_diff(const std::vector<double> & a, const std::vector<double> & b)
{
  double d = 0, n = 0;
  for (size_t i = 0; i < a.size(); ++i) { d += (a[i] - b[i]) * (a[i] - b[i]); n += b[i] * b[i]; }
  return std::sqrt(d / n);
}

static int run_mesh(int n_sub, int refine, int steps, int only)
{
  exadg_b200_hypercube_desc d{};
  d.degree = 4; d.n_subdivisions = n_sub; d.n_refinements = refine; d.mapping_degree = 1; d.deformation = 0.0; d.frequency = 2;
  d.ip_factor = 1.0; d.rank = 0; d.world = 1; d.force_general = 0;
  exadg_b200_operator * op = nullptr;
  CK(exadg_b200_create_hypercube(&d, &op));
  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));
  CK(exadg_b200_set_stream(op, stream));
  const int64_t n = exadg_b200_local_size(op);
  std::printf("cells %d^3 dofs %lld cartesian_path %d\n", n_sub << refine, (long long)n, exadg_b200_is_cartesian_path(op));
  // 0 pipelined, 1 WS depth 8, 2 WS depth 12, 3 WS with 4 producer warps (setmaxnreg), 4 / 5 warp-private kernel, 6 WS with 4 producer warps
  // and neighbour cells staged in shared memory by cp.async.  TOURNAMENT_VARIANTS=036 restricts the list (0 is the parity reference).
  constexpr int NV = 7;
  bool use[NV];
  for (int v = 0; v < NV; ++v) use[v] = true;
  if (const char * e = std::getenv("TOURNAMENT_VARIANTS")) {
    for (int v = 0; v < NV; ++v) use[v] = false;
    for (const char * c = e; *c; ++c) if (*c >= '0' && *c < '0' + NV) use[*c - '0'] = true;
  }
  double * src = nullptr, * dst[NV] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  CK(exadg_b200_initialize_dof_vector(op, &src));
  for (int v = 0; v < NV; ++v) if (use[v]) CK(exadg_b200_initialize_dof_vector(op, &dst[v]));
  {
    std::vector<double> h(n);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int64_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0; }
    CU(cudaMemcpy(src, h.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  for (int v = 0; v < NV; ++v) {
    if (only >= 0 && v != only) continue;
    if (!use[v]) continue;
    if (only < 0 && v == 2 && std::getenv("TOURNAMENT_SKIP2")) continue;
    exadg_b200_cartesian_kernel(v);
    for (int i = 0; i < 5; ++i) CK(exadg_b200_vmult(op, dst[v], src));
    CU(cudaStreamSynchronize(stream));
    float best = 1e30f, total = 0;
    const int reps = 3;
    for (int r = 0; r < reps; ++r) {
      CU(cudaEventRecord(e0, stream));
      for (int i = 0; i < steps; ++i) CK(exadg_b200_vmult(op, dst[v], src));
      CU(cudaEventRecord(e1, stream));
      CU(cudaEventSynchronize(e1));
      float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
      best = std::min(best, ms / steps); total += ms / steps;
    }
    std::printf("variant %d: %.4f ms/vmult best of %d (mean %.4f)  %.2f GDoF/s\n", v, best, reps, total / reps, n / best * 1e-6);
    std::fflush(stdout);
  }
  if (only < 0) { // parity: vmult of every variant against variant 0, vmult_add against 2 * vmult
    std::vector<double> y0(n), y(n);
    CU(cudaMemcpy(y0.data(), dst[0], n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int v = 1; v < NV; ++v) {
      if (!use[v] || !use[0]) continue;
      CU(cudaMemcpy(y.data(), dst[v], n * sizeof(double), cudaMemcpyDeviceToHost));
      std::printf("vmult     rel l2 (variant %d vs 0): %.3e\n", v, rel_diff(y, y0));
      exadg_b200_cartesian_kernel(v);
      CK(exadg_b200_vmult_add(op, dst[v], src)); // dst = 2 A src through the bulk add-reduction
      CU(cudaStreamSynchronize(stream));
      CU(cudaMemcpy(y.data(), dst[v], n * sizeof(double), cudaMemcpyDeviceToHost));
      for (auto & x : y) x *= 0.5;
      std::printf("vmult_add rel l2 (variant %d, half of it vs 0): %.3e\n", v, rel_diff(y, y0));
      std::fflush(stdout);
    }
  }
  exadg_b200_cartesian_kernel(0);
  exadg_b200_free_dof_vector(src);
  for (int v = 0; v < NV; ++v) if (dst[v]) exadg_b200_free_dof_vector(dst[v]);
  exadg_b200_destroy(op);
  return 0;
}

// ws_tournament only_variant (n_sub refine steps)...      e.g.  ws_tournament -1 3 5 100 1 2 5 5 0 5
int main(int argc, char ** argv)
{
  const int only = argc > 1 ? std::atoi(argv[1]) : -1;
  if (argc < 5) { const int rc = run_mesh(3, 5, 100, only); std::printf("TOURNAMENT DONE\n"); return rc; }
  for (int a = 2; a + 2 < argc; a += 3) {
    const int rc = run_mesh(std::atoi(argv[a]), std::atoi(argv[a + 1]), std::atoi(argv[a + 2]), only);
    if (rc != 0) return rc;
  }
  std::printf("TOURNAMENT DONE\n");
  return 0;
}
