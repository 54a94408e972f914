#include "vector_ops.cuh"

#include <algorithm>

#include "operator.cuh"

namespace exadg_b200
{
namespace
{

// CTA-wide sum in fixed order; valid in thread 0
__device__ __forceinline__ double block_sum(double v)
{
  __shared__ double sh[RED_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < RED_THREADS / 32; ++i) s += sh[i];
  return s;
}

// every CTA stores its partial sum; the CTA that arrives last adds all partials in index order
__device__ __forceinline__ void finish_reduction(double v, double * partial, double * result, int slot)
// (result[8 + slot] reinterpreted as the ticket counter of the slot)
{
  const double s = block_sum(v);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[slot * RED_BLOCKS + blockIdx.x] = s;
    __threadfence();
    unsigned int * ticket = reinterpret_cast<unsigned int *>(result + 8) + slot;
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) acc += ((volatile double *)partial)[slot * RED_BLOCKS + i];
    const double tot = block_sum(acc);
    if (threadIdx.x == 0) { result[slot] = tot; *(reinterpret_cast<unsigned int *>(result + 8) + slot) = 0u; }
  }
}

// All streaming kernels below move 16 bytes per access (double2; the vectors of the C ABI are 16-byte aligned, check_ptr in
// c_api.cu) with two independent accesses per thread in flight; an odd tail element is handled by one thread.  The summation
// order of the reductions is fixed by (grid, thread, element order) and therefore reproducible run to run.
__device__ __forceinline__ double2 ld2(const double * p, int64_t i2) { return reinterpret_cast<const double2 *>(p)[i2]; }
__device__ __forceinline__ void st2(double * p, int64_t i2, double2 v) { reinterpret_cast<double2 *>(p)[i2] = v; }

__global__ void __launch_bounds__(RED_THREADS) dot_kernel(const double * __restrict__ a, const double * __restrict__ b, int64_t n, double * partial, double * result, int slot)
{
  double acc0 = 0.0, acc1 = 0.0;
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n2; i += stride) {
    const double2 x = ld2(a, i), y = ld2(b, i);
    acc0 = fma(x.x, y.x, acc0); acc1 = fma(x.y, y.y, acc1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc0 = fma(a[n - 1], b[n - 1], acc0);
  finish_reduction(acc0 + acc1, partial, result, slot);
}

__global__ void __launch_bounds__(RED_THREADS) sum_kernel(const double * __restrict__ a, int64_t n, double * partial, double * result, int slot)
{
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * RED_THREADS) acc += a[i];
  finish_reduction(acc, partial, result, slot);
}

__global__ void __launch_bounds__(RED_THREADS) cg_update_x_g_kernel(double * __restrict__ x, const double * __restrict__ d, double * __restrict__ g, const double * __restrict__ h,
                                                                  int64_t n, double * partial, double * result, int slot, int num, int den)
{
  const double alpha = result[num] / result[den];
  double acc0 = 0.0, acc1 = 0.0;
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n2; i += stride) {
    const double2 xi = ld2(x, i), di = ld2(d, i), gi = ld2(g, i), hi = ld2(h, i);
    double2 xo, go;
    xo.x = fma(alpha, di.x, xi.x); xo.y = fma(alpha, di.y, xi.y);
    go.x = fma(alpha, hi.x, gi.x); go.y = fma(alpha, hi.y, gi.y);
    st2(x, i, xo); st2(g, i, go);
    acc0 = fma(go.x, go.x, acc0); acc1 = fma(go.y, go.y, acc1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    x[i] = fma(alpha, d[i], x[i]);
    const double gi = fma(alpha, h[i], g[i]);
    g[i] = gi; acc0 = fma(gi, gi, acc0);
  }
  finish_reduction(acc0 + acc1, partial, result, slot);
}

__global__ void __launch_bounds__(RED_THREADS) jacobi_dot_kernel(double * __restrict__ z, const double * __restrict__ inv_diag, const double * __restrict__ g, int64_t n,
                                                               double * partial, double * result, int slot)
{
  double acc0 = 0.0, acc1 = 0.0;
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t i = blockIdx.x * (int64_t)RED_THREADS + threadIdx.x; i < n2; i += stride) {
    const double2 gi = ld2(g, i), pi = ld2(inv_diag, i);
    double2 zi; zi.x = pi.x * gi.x; zi.y = pi.y * gi.y;
    st2(z, i, zi);
    acc0 = fma(gi.x, zi.x, acc0); acc1 = fma(gi.y, zi.y, acc1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) { const int64_t i = n - 1; const double gi = g[i], zi = inv_diag[i] * gi; z[i] = zi; acc0 = fma(gi, zi, acc0); }
  finish_reduction(acc0 + acc1, partial, result, slot);
}

__global__ void cg_update_d_kernel(double * __restrict__ d, const double * __restrict__ z, int64_t n, const double * result, int num, int den)
{
  const double beta = result[num] / result[den];
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 di = ld2(d, i), zi = ld2(z, i);
    double2 o; o.x = fma(beta, di.x, -zi.x); o.y = fma(beta, di.y, -zi.y);
    st2(d, i, o);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = fma(beta, d[n - 1], -z[n - 1]);
}

__global__ void axpby_kernel(double a, const double * __restrict__ x, double b, double * __restrict__ y, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}
__global__ void scale_copy_kernel(double a, const double * __restrict__ x, double * __restrict__ y, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a * x[i];
}
__global__ void fill_kernel(double * __restrict__ x, double v, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void invert_diagonal_kernel(double * __restrict__ d, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = d[i];
    d[i] = (fabs(v) > 1.0e-10) ? 1.0 / v : 1.0; // invert_diagonal.h:41-45
  }
}
__global__ void cheb_first_kernel(double * __restrict__ x, double * __restrict__ xold, const double * __restrict__ inv_diag, const double * __restrict__ b, const double * __restrict__ r,
                                  double inv_theta, int zero_start, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (zero_start) { xold[i] = 0.0; x[i] = inv_diag[i] * b[i] * inv_theta; }
    else { const double xi = x[i]; xold[i] = xi; x[i] = xi + inv_diag[i] * (b[i] - r[i]) * inv_theta; }
  }
}
__global__ void cheb_step_kernel(double * __restrict__ x, double * __restrict__ xold, const double * __restrict__ inv_diag, const double * __restrict__ b, const double * __restrict__ r,
                                 double f1, double f2, int64_t n)
{
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
    const double2 xi = ld2(x, i), xo = ld2(xold, i), pi = ld2(inv_diag, i), bi = ld2(b, i), ri = ld2(r, i);
    double2 o;
    o.x = xi.x + f1 * (xi.x - xo.x) + f2 * pi.x * (bi.x - ri.x);
    o.y = xi.y + f1 * (xi.y - xo.y) + f2 * pi.y * (bi.y - ri.y);
    st2(x, i, o); st2(xold, i, xi);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1; const double xi = x[i];
    x[i] = xi + f1 * (xi - xold[i]) + f2 * inv_diag[i] * (b[i] - r[i]);
    xold[i] = xi;
  }
}
__global__ void fill_mod11_kernel(double * __restrict__ x, int64_t off, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = (double)((i + off) % 11);
}
__global__ void add_scalar_kernel(double * __restrict__ x, double a, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += a;
}

// one CTA per COARSE cell; three tensor sweeps through shared memory per fine cell (p: the same cell, h: its eight children
// 8 c + child, child = x + 2 y + 4 z).  PROLONG: in = coarse, out = fine, matrices I_d; otherwise in = fine, out = coarse, matrices
// I_d^T, the children summed in child order.  Fixed summation order, no atomics.
template<bool PROLONG>
__global__ void __launch_bounds__(128) transfer_kernel(const TransferTable t, double * __restrict__ out, const double * __restrict__ in, int64_t n_coarse_cells)
{
  __shared__ double A[512], B[512], C[512];
  const int nc = t.nc, nf = t.nf;
  const int ni = PROLONG ? nc : nf, no = PROLONG ? nf : nc;
  const int n_child = t.h ? 8 : 1;
  for (int64_t cell = blockIdx.x; cell < n_coarse_cells; cell += gridDim.x) {
    if (!PROLONG) for (int i = threadIdx.x; i < no * no * no; i += blockDim.x) C[i] = 0.0;
    for (int child = 0; child < n_child; ++child) {
      const int64_t fine = t.h ? cell * 8 + child : cell;
      const double * src = in + (PROLONG ? cell * (int64_t)(nc * nc * nc) : fine * (int64_t)(nf * nf * nf));
      const double * Mx = t.I[t.h ? (child & 1) : 0], * My = t.I[t.h ? ((child >> 1) & 1) : 0], * Mz = t.I[t.h ? ((child >> 2) & 1) : 0];
      for (int i = threadIdx.x; i < ni * ni * ni; i += blockDim.x) A[i] = src[i];
      __syncthreads();
      // x: B[o, j, k] = sum_i M[o][i] A[i, j, k]      (extents: no x ni x ni)
      for (int e = threadIdx.x; e < no * ni * ni; e += blockDim.x) {
        const int o = e % no, jk = e / no;
        double v = 0.0;
        for (int i = 0; i < ni; ++i) v = fma(PROLONG ? Mx[o * nc + i] : Mx[i * nc + o], A[i + ni * jk], v);
        B[e] = v;
      }
      __syncthreads();
      // y: A[o1, o, k] = sum_j M[o][j] B[o1, j, k]    (no x no x ni)
      for (int e = threadIdx.x; e < no * no * ni; e += blockDim.x) {
        const int o1 = e % no, o = (e / no) % no, k = e / (no * no);
        double v = 0.0;
        for (int j = 0; j < ni; ++j) v = fma(PROLONG ? My[o * nc + j] : My[j * nc + o], B[o1 + no * (j + ni * k)], v);
        A[e] = v;
      }
      __syncthreads();
      // z: out[o1, o2, o] += sum_k M[o][k] A[o1, o2, k]
      double * dst = out + fine * (int64_t)(nf * nf * nf);
      for (int e = threadIdx.x; e < no * no * no; e += blockDim.x) {
        const int o12 = e % (no * no), o = e / (no * no);
        double v = 0.0;
        for (int k = 0; k < ni; ++k) v = fma(PROLONG ? Mz[o * nc + k] : Mz[k * nc + o], A[o12 + no * no * k], v);
        if (PROLONG) dst[e] += v; else C[e] += v; // every thread owns its entries of C
      }
      __syncthreads();
    }
    if (!PROLONG) {
      double * dst = out + cell * (int64_t)(nc * nc * nc);
      for (int e = threadIdx.x; e < no * no * no; e += blockDim.x) dst[e] += C[e];
    }
  }
}

inline unsigned ew_grid(int64_t n) { const int64_t g = (n + 255) / 256; return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); }
inline unsigned red_grid(int64_t n) { const int64_t g = (n + RED_THREADS - 1) / RED_THREADS; return (unsigned)(g < 1 ? 1 : (g > RED_BLOCKS ? RED_BLOCKS : g)); }
} // namespace

void reducer_init(Reducer & r)
{
  CUDA_CHECK(cudaMalloc(&r.partial, 8 * RED_BLOCKS * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&r.result, 16 * sizeof(double))); // 8 scalars + 8 ticket counters
  CUDA_CHECK(cudaMemset(r.result, 0, 16 * sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&r.host, 8 * sizeof(double)));
}
void reducer_free(Reducer & r)
{
  cudaFree(r.partial); cudaFree(r.result); cudaFreeHost(r.host);
  r.partial = r.result = r.host = nullptr;
}

void dot(const Reducer & r, int slot, const double * a, const double * b, int64_t n, cudaStream_t s)
{ dot_kernel<<<red_grid(n), RED_THREADS, 0, s>>>(a, b, n, r.partial, r.result, slot); }
void prolongate_add(const TransferTable & t, double * fine, const double * coarse, int64_t n_coarse_cells, cudaStream_t s)
{ if (n_coarse_cells > 0) transfer_kernel<true><<<(unsigned)std::min<int64_t>(n_coarse_cells, 148 * 16), 128, 0, s>>>(t, fine, coarse, n_coarse_cells); }
void restrict_add(const TransferTable & t, double * coarse, const double * fine, int64_t n_coarse_cells, cudaStream_t s)
{ if (n_coarse_cells > 0) transfer_kernel<false><<<(unsigned)std::min<int64_t>(n_coarse_cells, 148 * 16), 128, 0, s>>>(t, coarse, fine, n_coarse_cells); }
void sum(const Reducer & r, int slot, const double * a, int64_t n, cudaStream_t s)
{ sum_kernel<<<red_grid(n), RED_THREADS, 0, s>>>(a, n, r.partial, r.result, slot); }
void cg_update_x_g(const Reducer & r, int slot, int num, int den, double * x, const double * d, double * g, const double * h, int64_t n, cudaStream_t s)
{ cg_update_x_g_kernel<<<red_grid(n), RED_THREADS, 0, s>>>(x, d, g, h, n, r.partial, r.result, slot, num, den); }
void jacobi_dot(const Reducer & r, int slot, double * z, const double * inv_diag, const double * g, int64_t n, cudaStream_t s)
{ jacobi_dot_kernel<<<red_grid(n), RED_THREADS, 0, s>>>(z, inv_diag, g, n, r.partial, r.result, slot); }
void cg_update_d(const Reducer & r, int num, int den, double * d, const double * z, int64_t n, cudaStream_t s)
{ cg_update_d_kernel<<<ew_grid(n), 256, 0, s>>>(d, z, n, r.result, num, den); }
void axpby(double a, const double * x, double b, double * y, int64_t n, cudaStream_t s) { axpby_kernel<<<ew_grid(n), 256, 0, s>>>(a, x, b, y, n); }
void scale_copy(double a, const double * x, double * y, int64_t n, cudaStream_t s) { scale_copy_kernel<<<ew_grid(n), 256, 0, s>>>(a, x, y, n); }
void fill(double * x, double v, int64_t n, cudaStream_t s) { fill_kernel<<<ew_grid(n), 256, 0, s>>>(x, v, n); }
void invert_diagonal(double * d, int64_t n, cudaStream_t s) { invert_diagonal_kernel<<<ew_grid(n), 256, 0, s>>>(d, n); }
void cheb_first(double * x, double * xold, const double * inv_diag, const double * b, const double * r, double inv_theta, bool zero_start, int64_t n, cudaStream_t s)
{ cheb_first_kernel<<<ew_grid(n), 256, 0, s>>>(x, xold, inv_diag, b, r, inv_theta, zero_start ? 1 : 0, n); }
void cheb_step(double * x, double * xold, const double * inv_diag, const double * b, const double * r, double f1, double f2, int64_t n, cudaStream_t s)
{ cheb_step_kernel<<<ew_grid(n), 256, 0, s>>>(x, xold, inv_diag, b, r, f1, f2, n); }
void fill_mod11(double * x, int64_t global_offset, int64_t n, cudaStream_t s) { fill_mod11_kernel<<<ew_grid(n), 256, 0, s>>>(x, global_offset, n); }
void add_scalar(double * x, double a, int64_t n, cudaStream_t s) { add_scalar_kernel<<<ew_grid(n), 256, 0, s>>>(x, a, n); }

} // namespace exadg_b200
