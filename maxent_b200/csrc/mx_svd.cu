// Device SVD of the continuation kernel, replacing np.linalg.svd in KernelSVD.svd (python/kernels.py:53-64).
//
// The kernels of analytic continuation have a numerical rank of a few dozen (57 at 2000 x 1000, 54 at 10000 x 2000:
// the singular values fall below eps * S[0] after that), and only the triplets above the caller's cut -- at most the
// numerical rank -- enter the MaxEnt loop (python/kernels.py:101-122).  svd_truncated therefore computes the LEADING p
// triplets only:
//     Omega [m, p] pseudo-random          Y^T = Omega^T K   spans the row space of K up to rounding
//     Q = orth(Y)                         Gram-Schmidt with re-orthogonalisation (three passes: orthonormal to eps however
//                                         ill-conditioned Y is), columns that vanish at the rounding floor dropped
//     B = K Q  = U diag(S) W^T            one-sided (Hestenes) Jacobi on the p columns of B  ->  V = Q W
// K = U S (Q W)^T holds to eps * S[0] whenever range(Q) contains the row space of K numerically, i.e. whenever the
// p-th singular value returned is at the rounding floor -- the host checks exactly that and asks for more columns
// otherwise.  The Jacobi iteration is ONE persistent cooperative kernel per matrix (one CTA per column pair,
// round-robin tournament, grid-wide barrier between steps, convergence decided on the device: no host round trip);
// svd_jacobi (full SVD of an m x n matrix, every column) runs the same kernel on K itself.
#include <cooperative_groups.h>
#include "mx_common.cuh"

namespace cg = cooperative_groups;

namespace mx {

// ------------------------------------------------------------------------------------------------------------
// small dense products (set-up only): C[M, N] = A[M, K] * B[K, N]  or  A[M, K] * B[N, K]^T, row-major, FP64
// ------------------------------------------------------------------------------------------------------------
template <bool BT>
__global__ void __launch_bounds__(256) gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                                   int M, int N, int K) {
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ double As[TK][TM + 1];
    __shared__ double Bs[TK][TN + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            const int mm = i / TK, kk = i % TK;
            As[kk][mm] = (m0 + mm < M && k0 + kk < K) ? A[(int64_t)(m0 + mm) * K + k0 + kk] : 0.0;
        }
        if (BT) {
            for (int i = threadIdx.x; i < TN * TK; i += 256) {
                const int nn = i / TK, kk = i % TK;
                Bs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? B[(int64_t)(n0 + nn) * K + k0 + kk] : 0.0;
            }
        } else {
            for (int i = threadIdx.x; i < TN * TK; i += 256) {
                const int kk = i / TN, nn = i % TN;
                Bs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? B[(int64_t)(k0 + kk) * N + n0 + nn] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int mm = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
            if (mm < M && nn < N) C[(int64_t)mm * N + nn] = acc[i][j];
        }
}

static void gemm(bool bt, const double* A, const double* B, double* C, int M, int N, int K, cudaStream_t st) {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    if (bt) gemm_kernel<true><<<grid, 256, 0, st>>>(A, B, C, M, N, K);
    else gemm_kernel<false><<<grid, 256, 0, st>>>(A, B, C, M, N, K);
}

// Omega^T [p, m]: reproducible standard normals (counter-based hash + Box-Muller); no state, no library
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}
__global__ void randn_kernel(double* __restrict__ out, int64_t n, uint64_t seed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t a = mix64(seed + 2 * (uint64_t)i + 1), b = mix64(seed ^ (0x9e3779b97f4a7c15ull * (2 * (uint64_t)i + 2)));
        const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
        out[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
}

__global__ void transpose_kernel(const double* __restrict__ K, int m, int n, double* __restrict__ At) {
    const int64_t tot = (int64_t)m * n;
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < tot; o += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(o / m), i = (int)(o - (int64_t)j * m);
        At[o] = K[(int64_t)i * n + j];
    }
}

// ------------------------------------------------------------------------------------------------------------
// one-sided Jacobi, persistent cooperative kernel
// ------------------------------------------------------------------------------------------------------------
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

// At: the p columns of the matrix as contiguous rows [p][len]; Wt: accumulated right rotations [p][p] (identity on
// entry is written here) or nullptr; ctl: [0] = sum of squared column norms (double), ints behind it: rotations per
// sweep [max_sweeps], sweeps done.  Pairs whose columns are BOTH below eps * |A|_F are treated as converged (their
// mutual angles are rounding noise); a column at the noise floor is still rotated against the meaningful ones.
__global__ void __launch_bounds__(256) jacobi_kernel(double* __restrict__ At, double* __restrict__ Wt, int len, int p,
                                                     int max_sweeps, double tol, double* __restrict__ ctl) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[8];
    int* ictl = reinterpret_cast<int*>(ctl + 2);
    const int np = p + (p & 1);
    if (Wt) {
        for (int64_t o = blockIdx.x * 256ll + threadIdx.x; o < (int64_t)p * p; o += gridDim.x * 256ll) Wt[o] = (o / p == o % p) ? 1.0 : 0.0;
    }
    if (blockIdx.x == 0) {                  // |A|_F^2 in a fixed summation order (one CTA): the noise floor is reproducible
        double a = 0.0;
        for (int64_t i = threadIdx.x; i < (int64_t)p * len; i += 256) { const double x = At[i]; a = fma(x, x, a); }
        a = block_sum256(a, red);
        if (threadIdx.x == 0) ctl[0] = a;
    }
    grid.sync();
    const double eps = 2.220446049250313e-16;
    const double tiny2 = __ldcg(ctl) * eps * eps;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int step = 0; step < np - 1; ++step) {
            for (int t = blockIdx.x; t < np / 2; t += gridDim.x) {
                // round-robin tournament on np (even) players; player np-1 is fixed
                int a_ = (t == 0) ? np - 1 : (step + t) % (np - 1);
                int b_ = (t == 0) ? step % (np - 1) : (step - t + (np - 1)) % (np - 1);
                if (a_ >= p || b_ >= p) continue;       // dummy player for odd p
                if (a_ > b_) { const int x = a_; a_ = b_; b_ = x; }
                double* ap = At + (int64_t)a_ * len;
                double* aq = At + (int64_t)b_ * len;
                double a = 0, b = 0, g = 0;
                // columns move between CTAs (SMs) from step to step: read and write them at L2 (.cg), never through L1
                for (int i = threadIdx.x; i < len; i += 256) { const double x = __ldcg(ap + i), y = __ldcg(aq + i); a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g); }
                a = block_sum256(a, red); b = block_sum256(b, red); g = block_sum256(g, red);
                const double lim = sqrt(a) * sqrt(b);
                const bool rot = fabs(g) > tol * lim && a * b > 0.0 && !(a < tiny2 && b < tiny2);
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
                if (!rot && !swp) continue;
                if (threadIdx.x == 0 && rot) atomicAdd(ictl + sweep, 1);
                for (int i = threadIdx.x; i < len; i += 256) {
                    const double x = __ldcg(ap + i), y = __ldcg(aq + i);
                    const double xn = c * x - s * y, yn = s * x + c * y;
                    __stcg(ap + i, swp ? yn : xn); __stcg(aq + i, swp ? xn : yn);
                }
                if (Wt) {
                    double* vp = Wt + (int64_t)a_ * p;
                    double* vq = Wt + (int64_t)b_ * p;
                    for (int i = threadIdx.x; i < p; i += 256) {
                        const double x = __ldcg(vp + i), y = __ldcg(vq + i);
                        const double xn = c * x - s * y, yn = s * x + c * y;
                        __stcg(vp + i, swp ? yn : xn); __stcg(vq + i, swp ? xn : yn);
                    }
                }
            }
            grid.sync();
        }
        if (__ldcg(ictl + sweep) == 0) { ++sweep; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ictl[max_sweeps] = sweep;
}

// Classical Gram-Schmidt with re-orthogonalisation of the p columns of Y (rows of Yt, length len), one CTA: column j is
// projected three times against the finished columns ("twice is enough" holds for residual ratios down to sqrt(eps);
// Y has nearly dependent columns by construction, the third pass covers ratios down to the rounding floor), then
// normalised.  A column whose residual is below drop_rel times its original norm is numerically in the span of the
// previous ones: it is zeroed.  (One-sided Jacobi is NOT used here: on a matrix whose columns are nearly parallel and
// of equal norm it loses the small directions; on B = K Q, whose columns are graded, it is accurate.)
__global__ void __launch_bounds__(1024) cgs_kernel(double* __restrict__ Yt, int len, int p, double drop_rel) {
    __shared__ double r[512];
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += red[w];
        return t;
    };
    for (int j = 0; j < p; ++j) {
        double* yj = Yt + (int64_t)j * len;
        double n0 = 0.0;
        for (int k = tid; k < len; k += 1024) { const double x = yj[k]; n0 = fma(x, x, n0); }
        n0 = block_sum(n0);
        for (int pass = 0; pass < 3 && j > 0; ++pass) {
            for (int i = warp; i < j; i += 32) {
                const double* qi = Yt + (int64_t)i * len;
                double d = 0.0;
                for (int k = lane; k < len; k += 32) d = fma(qi[k], yj[k], d);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) r[i] = d;
            }
            __syncthreads();
            for (int k = tid; k < len; k += 1024) {
                double acc = yj[k];
#pragma unroll 8
                for (int i = 0; i < j; ++i) acc = fma(-r[i], Yt[(int64_t)i * len + k], acc);
                yj[k] = acc;
            }
            __syncthreads();
        }
        double n1 = 0.0;
        for (int k = tid; k < len; k += 1024) { const double x = yj[k]; n1 = fma(x, x, n1); }
        n1 = block_sum(n1);
        const double inv = (n1 > drop_rel * drop_rel * n0 && n1 > 0.0) ? rsqrt(n1) : 0.0;
        for (int k = tid; k < len; k += 1024) yj[k] *= inv;
        __syncthreads();
    }
}

int gram_schmidt_rows(double* Yt, int len, int p, double drop_rel, cudaStream_t stream) {
    cgs_kernel<<<1, 1024, 0, stream>>>(Yt, len, p, drop_rel);
    return cudaGetLastError() == cudaSuccess ? MX_OK : MX_ERR_CUDA;
}

__global__ void __launch_bounds__(256) col_norm_kernel(const double* __restrict__ At, int len, int p, double* __restrict__ S) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    double a = 0;
    for (int i = threadIdx.x; i < len; i += 256) { const double x = At[(int64_t)j * len + i]; a = fma(x, x, a); }
    a = block_sum256(a, red);
    if (threadIdx.x == 0) S[j] = sqrt(a);
}

// rank columns by norm (descending, ties by index) and scatter into U [len, ldu], S, V [nv, ldu] (V = Vsrc columns)
__global__ void __launch_bounds__(256) finish_kernel(const double* __restrict__ At, const double* __restrict__ Vt, const double* __restrict__ Sun,
                                                     int len, int nv, int p, double* __restrict__ U, double* __restrict__ S, double* __restrict__ V) {
    const int j = blockIdx.x;
    const double sj = Sun[j];
    __shared__ int rank_s;
    __shared__ int cnt[8];
    int c = 0;
    for (int i = threadIdx.x; i < p; i += 256) { const double si = Sun[i]; if (si > sj || (si == sj && i < j)) ++c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += cnt[w]; rank_s = t; S[t] = sj; }
    __syncthreads();
    const int r = rank_s;
    const double inv = sj > 0.0 ? 1.0 / sj : 0.0;
    for (int i = threadIdx.x; i < len; i += 256) U[(int64_t)i * p + r] = At[(int64_t)j * len + i] * inv;
    for (int i = threadIdx.x; i < nv; i += 256) V[(int64_t)i * p + r] = Vt[(int64_t)j * nv + i];
}

static int launch_jacobi(double* At, double* Wt, int len, int p, int max_sweeps, double* ctl, cudaStream_t stream) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MX_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel, 256, 0) != cudaSuccess || per_sm < 1) return MX_ERR_CUDA;
    int grid = (p + 1) / 2;
    if (grid > per_sm * sms) grid = per_sm * sms;
    if (grid < 1) grid = 1;
    if (max_sweeps > 60) max_sweeps = 60;
    if (cudaMemsetAsync(ctl, 0, 64 * sizeof(double), stream) != cudaSuccess) return MX_ERR_CUDA;
    double tol = 1e-15;
    void* args[] = {&At, &Wt, &len, &p, &max_sweeps, &tol, &ctl};
    if (cudaLaunchCooperativeKernel((void*)jacobi_kernel, dim3(grid), dim3(256), args, 0, stream) != cudaSuccess) return MX_ERR_CUDA;
    return MX_OK;
}

// full thin SVD of K[m, n] (m >= n): every column.  work: m*n + n*n + n + 64 doubles
int svd_jacobi(const double* K, int m, int n, double* U, double* S, double* V, double* work,
               int max_sweeps, int* sweeps_done, cudaStream_t stream) {
    double* At = work;
    double* Wt = work + (int64_t)m * n;
    double* Sun = Wt + (int64_t)n * n;
    double* ctl = Sun + n;
    transpose_kernel<<<1184, 256, 0, stream>>>(K, m, n, At);
    int rc = launch_jacobi(At, Wt, m, n, max_sweeps, ctl, stream);
    if (rc != MX_OK) return rc;
    col_norm_kernel<<<n, 256, 0, stream>>>(At, m, n, Sun);
    finish_kernel<<<n, 256, 0, stream>>>(At, Wt, Sun, m, n, n, U, S, V);
    if (sweeps_done) {       // stream-ordered copy: valid after the caller synchronises the stream
        const int ms = max_sweeps > 60 ? 60 : max_sweeps;
        if (cudaMemcpyAsync(sweeps_done, reinterpret_cast<int*>(ctl + 2) + ms, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            return MX_ERR_CUDA;
    }
    return cudaGetLastError() == cudaSuccess ? MX_OK : MX_ERR_CUDA;
}

int64_t svd_truncated_work_doubles(int m, int n, int p) {
    const int64_t big = m > n ? m : n;
    return (int64_t)p * m + 2 * (int64_t)p * big + (int64_t)p * p + (int64_t)p * n + p + 2 * 64;
}

// leading p triplets of K[m, n] (any shape, p <= min(m, n)): U[m, p], S[p] (descending), V[n, p]
int svd_truncated(const double* K, int m, int n, int p, double* U, double* S, double* V, double* work, uint64_t seed,
                  cudaStream_t stream) {
    const int64_t big = m > n ? m : n;
    double* Ot = work;                       // Omega^T [p][m]
    double* Yt = Ot + (int64_t)p * m;        // (Omega^T K) [p][n] -> Q^T (orthonormal rows)
    double* Bt = Yt + (int64_t)p * big;      // (K Q)^T [p][m]
    double* Wt = Bt + (int64_t)p * big;      // right rotations of the second Jacobi [p][p]
    double* Vt = Wt + (int64_t)p * p;        // (Q W)^T [p][n]
    double* Sun = Vt + (int64_t)p * n;
    double* ctl = Sun + p;
    randn_kernel<<<592, 256, 0, stream>>>(Ot, (int64_t)p * m, seed);
    gemm(false, Ot, K, Yt, p, n, m, stream);                          // Y^T = Omega^T K
    if (p > 512) return MX_ERR_UNSUPPORTED;
    cgs_kernel<<<1, 1024, 0, stream>>>(Yt, n, p, 4 * 2.220446049250313e-16);   // Q = orth(Y)
    gemm(true, Yt, K, Bt, p, m, n, stream);                           // B^T = Q^T K^T
    int rc = launch_jacobi(Bt, Wt, m, p, 40, ctl + 64, stream);
    if (rc != MX_OK) return rc;
    gemm(false, Wt, Yt, Vt, p, n, p, stream);                         // V^T = W^T Q^T
    col_norm_kernel<<<p, 256, 0, stream>>>(Bt, m, p, Sun);
    finish_kernel<<<p, 256, 0, stream>>>(Bt, Vt, Sun, m, n, p, U, S, V);
    return cudaGetLastError() == cudaSuccess ? MX_OK : MX_ERR_CUDA;
}

}  // namespace mx
