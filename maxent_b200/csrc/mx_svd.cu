// One-sided (Hestenes) Jacobi SVD on the device, replacing np.linalg.svd in KernelSVD.svd
// (python/kernels.py:53-64).  K[m, n] (m >= n) = U diag(S) V^T, S descending.
// Columns of K are kept as contiguous rows of `work` ([n][m]); one CTA rotates one column pair,
// n/2 disjoint pairs per step (round-robin tournament), n-1 steps per sweep.  Pairs whose columns
// are both below eps * |K|_F are treated as converged (the kernel matrix has numerical rank ~55,
// the remaining columns are rounding noise whose mutual angles are meaningless).
#include "mx_common.cuh"

namespace mx {

__global__ void svd_init_kernel(const double* __restrict__ K, int m, int n, double* __restrict__ At, double* __restrict__ Vw) {
    const int64_t tot = (int64_t)m * n;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < tot; o += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(o / m), i = (int)(o - (int64_t)j * m);
        At[o] = K[(int64_t)i * n + j];
    }
    const int64_t tv = (int64_t)n * n;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < tv; o += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(o / n), q = (int)(o - (int64_t)p * n);
        Vw[o] = p == q ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ double block_sum256(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(256) svd_step_kernel(double* __restrict__ At, double* __restrict__ Vw, int m, int n,
                                                       int np, int step, double tol, double tiny2, int* __restrict__ nrot) {
    __shared__ double red[8];
    // round-robin tournament on np (even) players; player np-1 is fixed
    const int t = blockIdx.x;
    int p, q;
    if (t == 0) { p = np - 1; q = step % (np - 1); }
    else { p = (step + t) % (np - 1); q = (step - t + (np - 1)) % (np - 1); }
    if (p >= n || q >= n) return;               // dummy player for odd n
    if (p > q) { const int x = p; p = q; q = x; }
    double* ap = At + (int64_t)p * m;
    double* aq = At + (int64_t)q * m;
    double a = 0, b = 0, g = 0;
    for (int i = threadIdx.x; i < m; i += 256) { const double x = ap[i], y = aq[i]; a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g); }
    a = block_sum256(a, red); b = block_sum256(b, red); g = block_sum256(g, red);
    const double lim = sqrt(a) * sqrt(b);
    bool rot = fabs(g) > tol * lim && a * b > 0.0 && !(a < tiny2 && b < tiny2);
    double c = 1.0, s = 0.0;
    if (rot) {
        const double zeta = (b - a) / (2.0 * g);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        c = 1.0 / sqrt(1.0 + tt * tt); s = c * tt;
    }
    // keep the larger column at the lower index (de Rijk): swap if needed
    const double an = rot ? (c * c * a - 2 * c * s * g + s * s * b) : a;
    const double bn = rot ? (s * s * a + 2 * c * s * g + c * c * b) : b;
    const bool swp = bn > an;
    if (!rot && !swp) return;
    if (threadIdx.x == 0 && rot) atomicAdd(nrot, 1);
    for (int i = threadIdx.x; i < m; i += 256) {
        const double x = ap[i], y = aq[i];
        const double xn = c * x - s * y, yn = s * x + c * y;
        ap[i] = swp ? yn : xn; aq[i] = swp ? xn : yn;
    }
    double* vp = Vw + (int64_t)p * n;
    double* vq = Vw + (int64_t)q * n;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double x = vp[i], y = vq[i];
        const double xn = c * x - s * y, yn = s * x + c * y;
        vp[i] = swp ? yn : xn; vq[i] = swp ? xn : yn;
    }
}

__global__ void __launch_bounds__(256) svd_norm_kernel(const double* __restrict__ At, int m, int n, double* __restrict__ S) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    double a = 0;
    for (int i = threadIdx.x; i < m; i += 256) { const double x = At[(int64_t)j * m + i]; a = fma(x, x, a); }
    a = block_sum256(a, red);
    if (threadIdx.x == 0) S[j] = sqrt(a);
}

// rank columns by norm (descending, ties by index) and scatter into U, S, V
__global__ void __launch_bounds__(256) svd_finish_kernel(const double* __restrict__ At, const double* __restrict__ Vw,
                                                         const double* __restrict__ Sun, int m, int n,
                                                         double* __restrict__ U, double* __restrict__ S, double* __restrict__ V) {
    const int j = blockIdx.x;
    const double sj = Sun[j];
    __shared__ int rank_s;
    __shared__ int cnt[8];
    int c = 0;
    for (int i = threadIdx.x; i < n; i += 256) { const double si = Sun[i]; if (si > sj || (si == sj && i < j)) ++c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += cnt[w]; rank_s = t; S[t] = sj; }
    __syncthreads();
    const int r = rank_s;
    const double inv = sj > 0.0 ? 1.0 / sj : 0.0;
    for (int i = threadIdx.x; i < m; i += 256) U[(int64_t)i * n + r] = At[(int64_t)j * m + i] * inv;
    for (int i = threadIdx.x; i < n; i += 256) V[(int64_t)i * n + r] = Vw[(int64_t)j * n + i];
}

__global__ void svd_fro_kernel(const double* __restrict__ S, int n, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { double t = 0; for (int i = 0; i < n; ++i) t += S[i] * S[i]; out[0] = t; }
}

int svd_jacobi(const double* K, int m, int n, double* U, double* S, double* V, double* work,
               int max_sweeps, int* sweeps_done, cudaStream_t stream) {
    // work: At [n*m] ; we borrow V as Vw during the iteration?  No: V is the output layout, keep
    // a separate Vw inside work: work must hold n*m + n*n + n + 2 doubles.
    double* At = work;
    double* Vw = work + (int64_t)m * n;
    double* Sun = Vw + (int64_t)n * n;
    double* fro = Sun + n;
    int* nrot = reinterpret_cast<int*>(fro + 1);
    svd_init_kernel<<<1184, 256, 0, stream>>>(K, m, n, At, Vw);
    svd_norm_kernel<<<n, 256, 0, stream>>>(At, m, n, Sun);
    svd_fro_kernel<<<1, 32, 0, stream>>>(Sun, n, fro);
    double fro2 = 0.0;
    if (cudaMemcpyAsync(&fro2, fro, sizeof(double), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return MX_ERR_CUDA;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return MX_ERR_CUDA;
    const double eps = 2.220446049250313e-16;
    const double tiny2 = fro2 * eps * eps;      // squared norm below which a column is rounding noise
    const double tol = 1e-15;
    const int np = n + (n & 1);
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (cudaMemsetAsync(nrot, 0, sizeof(int), stream) != cudaSuccess) return MX_ERR_CUDA;
        for (int step = 0; step < np - 1; ++step)
            svd_step_kernel<<<np / 2, 256, 0, stream>>>(At, Vw, m, n, np, step, tol, tiny2, nrot);
        int h = 0;
        if (cudaMemcpyAsync(&h, nrot, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return MX_ERR_CUDA;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return MX_ERR_CUDA;
        if (h == 0) { ++sweep; break; }
    }
    if (sweeps_done) *sweeps_done = sweep;
    svd_norm_kernel<<<n, 256, 0, stream>>>(At, m, n, Sun);
    svd_finish_kernel<<<n, 256, 0, stream>>>(At, Vw, Sun, m, n, U, S, V);
    return cudaGetLastError() == cudaSuccess ? MX_OK : MX_ERR_CUDA;
}

}  // namespace mx
