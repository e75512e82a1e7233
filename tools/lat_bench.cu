// Instruction-latency microbenchmarks for the FP64 path of sm_100a (B200): what one warp sees.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_bench tools/lat_bench.cu && tools/lat_bench
// One JSON object on stdout (cycles per operation, SM clock).  These numbers size the dependent chains of the sweep
// kernel (exp in the cost pass, Cholesky pivots, triangular solves) against the issue cost of DMMA.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int N = 2048;

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- dependent DFMA chains, ILP independent chains per thread ----
template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- DMMA: ILP independent accumulators per warp ----
template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- one dependent DFMA chain interleaved with NM independent DMMAs per chain step ----
template <int NM>
__global__ void k_mix(double* out, long long* cyc, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    double x = threadIdx.x * 1e-3, y = x + 1.0;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NM) dmma(c[i], a, b);
            x = fma(x, a, b);
            y = fma(y, a, b);
        }
    }
    long long t1 = clock64();
    double s = x + y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- shuffle -> fma dependent chain (Cholesky pivot broadcast) ----
__global__ void k_shfl(double* out, long long* cyc, double a, double b) {
    double x = threadIdx.x * 1e-3 + 1.0;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
        x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 5) & 31);
        x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- reciprocal: MUFU.RCP64H seed + two Newton steps, dependent ----
__global__ void k_rcp(double* out, long long* cyc, double a0) {
    double a = a0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
        double e = fma(-a, y, 1.0);
        y = fma(y, e, y);
        e = fma(-a, y, 1.0);
        a = fma(y, e, y) + 1.5;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_mufu(double* out, long long* cyc, double a0) {
    double a = a0 + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
        double y;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
        a = y;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- LDS dependent (pointer chase) ----
__global__ void k_lds(double* out, long long* cyc) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) p = idx[p];
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- mbarrier try_wait on a completed phase / test_wait / arrive ----
__global__ void k_mbar(double* out, long long* cyc) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) {
        uint32_t ok;
        asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        acc += ok;
    }
    long long t1 = clock64();
    for (int it = 0; it < N; ++it) {
        uint32_t ok;
        asm volatile("{ .reg .pred P1; mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        acc += ok;
    }
    long long t2 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}

// ---- CTA barrier round trip with W warps ----
__global__ void k_bar(double* out, long long* cyc) {
    long long t0 = clock64();
    for (int it = 0; it < N; ++it) __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = 0;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// exp_main (the straight-line exp of the cost pass), CH chains per thread
__device__ __forceinline__ double exp_main(double x) {
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const double n = t - 6755399441055744.0;
    double r = fma(n, -0.6931471805599453, x);
    r = fma(n, -2.3190468138462996e-17, r);
    double p = fma(r, 2.502232253650299e-08, 2.763090348817311e-07);
    p = fma(r, p, 2.755751454588244e-06);
    p = fma(r, p, 2.4801491039099165e-05);
    p = fma(r, p, 0.00019841269589115497);
    p = fma(r, p, 0.001388888894591638);
    p = fma(r, p, 0.008333333333455043);
    p = fma(r, p, 0.041666666666519754);
    p = fma(r, p, 0.16666666666666477);
    p = fma(r, p, 0.5000000000000012);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    return __hiloint2double(__double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
}
template <int CHN>
__global__ void k_exp(double* out, long long* cyc, double a) {
    double x[CHN];
#pragma unroll
    for (int i = 0; i < CHN; ++i) x[i] = threadIdx.x * 1e-3 + i * 0.1;
    long long t0 = clock64();
    for (int it = 0; it < N / 8; ++it) {
#pragma unroll
        for (int i = 0; i < CHN; ++i) x[i] = exp_main(x[i]) * a;
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHN; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMalloc(&cyc, 1024));
    long long h[4];
    printf("{");
    auto report = [&](const char* name, double per, bool last = false) { printf("\"%s\": %.2f%s", name, per, last ? "" : ", "); };
#define RUN1(name, launch, div) do { launch; CK(cudaDeviceSynchronize()); launch; CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost)); report(name, (double)h[0] / (div)); } while (0)
    // one warp alone on an SM
    RUN1("dfma_dep_1chain_clk", (k_dfma<1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N);
    RUN1("dfma_2chains_clk_per_instr", (k_dfma<2><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 2.0 * N);
    RUN1("dfma_4chains_clk_per_instr", (k_dfma<4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 4.0 * N);
    RUN1("dfma_8chains_clk_per_instr", (k_dfma<8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 8.0 * N);
    RUN1("dfma_16chains_clk_per_instr", (k_dfma<16><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 16.0 * N);
    RUN1("dmma_dep_1acc_clk", (k_dmma<1><<<1, 32>>>(out, cyc, 1e-3, 1e-3)), N);
    RUN1("dmma_2acc_clk_per_instr", (k_dmma<2><<<1, 32>>>(out, cyc, 1e-3, 1e-3)), 2.0 * N);
    RUN1("dmma_4acc_clk_per_instr", (k_dmma<4><<<1, 32>>>(out, cyc, 1e-3, 1e-3)), 4.0 * N);
    RUN1("dmma_8acc_clk_per_instr", (k_dmma<8><<<1, 32>>>(out, cyc, 1e-3, 1e-3)), 8.0 * N);
    RUN1("dmma_14acc_clk_per_instr", (k_dmma<14><<<1, 32>>>(out, cyc, 1e-3, 1e-3)), 14.0 * N);
    // two and four warps on the same SM sub-partition (block of 8 / 16 warps: warps w, w+4, ... share a scheduler)
    RUN1("dmma_8acc_2warps_per_smsp_clk_per_instr_per_warp", (k_dmma<8><<<1, 256>>>(out, cyc, 1e-3, 1e-3)), 8.0 * N);
    RUN1("dmma_8acc_4warps_per_smsp_clk_per_instr_per_warp", (k_dmma<8><<<1, 512>>>(out, cyc, 1e-3, 1e-3)), 8.0 * N);
    RUN1("dfma_dep_2warps_per_smsp_clk", (k_dfma<1><<<1, 256>>>(out, cyc, 1.0000001, 1e-9)), N);
    RUN1("dfma_dep_4warps_per_smsp_clk", (k_dfma<1><<<1, 512>>>(out, cyc, 1.0000001, 1e-9)), N);
    RUN1("mix_2dfma_chains_0dmma_clk_per_step", (k_mix<0><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 8.0 * N);
    RUN1("mix_2dfma_chains_4dmma_of_8_clk_per_step", (k_mix<4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 8.0 * N);
    RUN1("mix_2dfma_chains_8dmma_of_8_clk_per_step", (k_mix<8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), 8.0 * N);
    RUN1("shfl_then_dfma_dep_clk", (k_shfl<<<1, 32>>>(out, cyc, 1.0000001, 1e-9)), N);
    RUN1("rcp_nr_dep_clk", (k_rcp<<<1, 32>>>(out, cyc, 3.0)), N);
    RUN1("mufu_rcp64h_dep_clk", (k_mufu<<<1, 32>>>(out, cyc, 3.0)), N);
    RUN1("lds_dep_clk", (k_lds<<<1, 32>>>(out, cyc)), N);
    RUN1("exp_main_1chain_clk", (k_exp<1><<<1, 32>>>(out, cyc, 1e-3)), N / 8);
    RUN1("exp_main_2chains_clk_per_exp", (k_exp<2><<<1, 32>>>(out, cyc, 1e-3)), 2.0 * (N / 8));
    RUN1("exp_main_4chains_clk_per_exp", (k_exp<4><<<1, 32>>>(out, cyc, 1e-3)), 4.0 * (N / 8));
    RUN1("exp_main_2chains_2warps_per_smsp_clk_per_exp", (k_exp<2><<<1, 256>>>(out, cyc, 1e-3)), 2.0 * (N / 8));
    RUN1("syncthreads_8warps_clk", (k_bar<<<1, 256>>>(out, cyc)), N);
    RUN1("syncthreads_4warps_clk", (k_bar<<<1, 128>>>(out, cyc)), N);
    k_mbar<<<1, 32>>>(out, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    report("mbar_try_wait_complete_clk", (double)h[0] / N);
    report("mbar_test_wait_complete_clk", (double)h[1] / N, true);
    printf("}\n");
    return 0;
}
