// The inner step of the cost pass (T-pass) of the sweep kernel in isolation: the k-tile already sits in shared memory,
// no TMA, no mbarriers -- what does ONE step cost a warp (x = V' t for 8 trials, two exponentials per lane, entropy
// terms, y += H^T V'), alone on its scheduler and with 2 / 4 warps per scheduler?  Variants switch parts off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Imaxent_b200/csrc -o tools/tpass_loop_bench tools/tpass_loop_bench.cu
#include <cstdio>
#include <cstdlib>
#include "mx_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NT = 7;
constexpr int ITERS = 512;
using mx::dmma;
using mx::tile_off;

// MODE bit 0: x-MMA, bit 1: pointwise (exp + entropy), bit 2: y-MMA, bit 3: global store of w
template <int MODE>
__global__ void step_kernel(const double* __restrict__ D, double* __restrict__ wout, double* out, long long* cyc) {
    extern __shared__ __align__(128) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, q = lane & 3;
    const int nw = blockDim.x >> 5;
    for (int i = tid; i < nw * NT * 64; i += blockDim.x) sm[i] = 1e-3 * ((i * 7) % 13 - 6);
    __syncthreads();
    const int offX = tile_off(r, 2 * q), offY0 = tile_off(2 * q, r), offY1 = tile_off(2 * q + 1, r);
    double tA[NT][2], yacc[NT][2];
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) { tA[jt][0] = 1e-2 * (r + jt); tA[jt][1] = 1e-2 * (q - jt); yacc[jt][0] = 0; yacc[jt][1] = 0; }
    double sacc = 0.0;
    const double* tile = sm + warp * NT * 64;
    double2 Dv = *reinterpret_cast<const double2*>(D + 2 * q);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        const double2 Dn = *reinterpret_cast<const double2*>(D + ((it + 1) & 63) * 8 + 2 * q);
        double C0[2] = {0.0, 0.0}, C1[2] = {0.0, 0.0};
        if (MODE & 1) {
#pragma unroll
            for (int jt = 0; jt < NT; ++jt) {
                const double2 vv = *reinterpret_cast<const double2*>(tile + jt * 64 + offX);
                dmma(C0, tA[jt][0], vv.x);
                dmma(C1, tA[jt][1], vv.y);
            }
        } else { C0[0] = sacc * 1e-9; C1[1] = tA[0][0]; }
        const double x0 = C0[0] + C1[0], x1 = C0[1] + C1[1];
        double Hv[2] = {x0, x1};
        if (MODE & 2) {
            double ex2[2];
            if (__any_sync(0xffffffffu, mx::exp_is_special(x0) || mx::exp_is_special(x1))) { ex2[0] = exp(x0); ex2[1] = exp(x1); }
            else { ex2[0] = mx::exp_main(x0); ex2[1] = mx::exp_main(x1); }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double Dk = i ? Dv.y : Dv.x, x = i ? x1 : x0, ex = ex2[i];
                double H = Dk * ex;
                const double lg = (ex <= 1e-100) ? -230.25850929940458 : x;
                double st_ = H - Dk - H * lg;
                if (Dk == 0.0) { H = 0.0; st_ = 0.0; }
                sacc += st_;
                Hv[i] = H;
            }
        }
        if (MODE & 8) *reinterpret_cast<double2*>(wout + ((size_t)blockIdx.x * nw + warp) * 8192 + (it & 127) * 64 + 2 * lane) = make_double2(Hv[0], Hv[1]);
        if (MODE & 4) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int off = e ? offY1 : offY0;
#pragma unroll
                for (int jt = 0; jt < NT; ++jt) dmma(yacc[jt], Hv[e], tile[jt * 64 + off]);
            }
        } else { sacc += Hv[0] + Hv[1]; }
        Dv = Dn;
        __syncwarp();
    }
    const long long t1 = clock64();
    double s = sacc;
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) s += yacc[jt][0] + yacc[jt][1];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int nwarps, const double* D, double* w, double* out, long long* cyc, bool last = false) {
    long long h;
    const size_t smem = (size_t)nwarps * NT * 64 * 8;
    CK(cudaFuncSetAttribute(step_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int rep = 0; rep < 2; ++rep) { step_kernel<MODE><<<1, nwarps * 32, smem>>>(D, w, out, cyc); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("\"%s_%dwarps_clk_per_step\": %.1f%s", name, nwarps, (double)h / ITERS, last ? "" : ", ");
}

int main() {
    double *D, *w, *out;
    long long* cyc;
    CK(cudaMalloc(&D, 4096 * 8));
    CK(cudaMalloc(&w, (size_t)16 * 8192 * 8 * 4));
    CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMalloc(&cyc, 1024));
    double hD[4096];
    for (int i = 0; i < 4096; ++i) hD[i] = 1e-3 + 1e-6 * i;
    CK(cudaMemcpy(D, hD, sizeof(hD), cudaMemcpyHostToDevice));
    printf("{");
    for (int nw : {4, 8, 16}) {
        run<15>("full", nw, D, w, out, cyc);
        run<7>("no_store", nw, D, w, out, cyc);
        run<5>("mma_only", nw, D, w, out, cyc);
        run<1>("xmma_only", nw, D, w, out, cyc);
        run<4>("ymma_only", nw, D, w, out, cyc);
        run<2>("pointwise_only", nw, D, w, out, cyc);
        run<3>("xmma_pointwise", nw, D, w, out, cyc, nw == 16);
    }
    printf("}\n");
    return 0;
}
