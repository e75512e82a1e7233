// FP64 microbenchmarks for B200 (sm_100a): DFMA peak, DMMA (mma.sync f64) peak for the
// shapes ptxas accepts, mixed DFMA+DMMA, exp() throughput and L2->SM read bandwidth.
// Output: one JSON object on stdout. These numbers are the roofline denominators for the
// alpha-sweep kernel (MEASURED_PEAKS.json carries no FP64 figure).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
    double c[NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE>  // 4, 8, 16 = k of m16n8kX
__global__ void __launch_bounds__(256) k_dmma16(double* out, double a, double b) {
    constexpr int NT = 6;
    double c[NT][4];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            if (SHAPE == 4) dmma1684(c[i], av[0], av[1], bv[0]);
            if (SHAPE == 8) { double a4[4] = {av[0], av[1], av[2], av[3]}; double b2[2] = {bv[0], bv[1]}; dmma1688(c[i], a4, b2); }
            if (SHAPE == 16) dmma16816(c[i], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per loop 8 DMMA884 + 16 DFMA (does DFMA ride for free next to DMMA?)
__global__ void __launch_bounds__(256) k_mixed(double* out, double a, double b) {
    double c[8][2], d[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dmma884(c[i][0], c[i][1], a, b); d[2 * i] = fma(d[2 * i], a, b); d[2 * i + 1] = fma(d[2 * i + 1], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_exp(double* out, double a) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = a * (threadIdx.x + i);
    double s = 0;
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s += exp(x[i]); x[i] += 1e-9; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every CTA streams the same `n` doubles `reps` times: L2 (or L1) -> SM bandwidth
template <bool CG>
__global__ void __launch_bounds__(256) k_l2read(const double2* __restrict__ buf, size_t n2, int reps, double* out) {
    double s = 0;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = threadIdx.x; i < n2; i += blockDim.x) {
            double2 v;
            if (CG) asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(buf + i));
            else    asm volatile("ld.global.ca.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(buf + i));
            s += v.x + v.y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[0];
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    for (int cps : {1, 2, 4}) {   // CTAs per SM of 256 threads
        int grid = sms * cps;
        double ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dfma_tflops_cps%d\": %.3f", cps, 2.0 * 16 * ITERS * 256.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma884_tflops_cps%d\": %.3f", cps, 2.0 * 256 * 8 * ITERS * 8.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma16<4><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma1684_tflops_cps%d\": %.3f", cps, 2.0 * 512 * 6 * ITERS * 8.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma16<8><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma1688_tflops_cps%d\": %.3f", cps, 2.0 * 1024 * 6 * ITERS * 8.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma16<16><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma16816_tflops_cps%d\": %.3f", cps, 2.0 * 2048 * 6 * ITERS * 8.0 * grid / ms / 1e9);
    }
    {
        int grid = sms * 2;
        double ms = time_ms([&] { k_dmma884<2><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma884_ilp2_tflops\": %.3f", 2.0 * 256 * 2 * ITERS * 8.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma884<4><<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma884_ilp4_tflops\": %.3f", 2.0 * 256 * 4 * ITERS * 8.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_dmma884<1><<<sms, 32>>>(out, 1.0000001, 1e-9); });
        printf(", \"dmma884_dep_latency_clk_est_ms\": %.4f", ms);
        ms = time_ms([&] { k_mixed<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        printf(", \"mixed_ms\": %.4f, \"mixed_dmma_tflops\": %.3f, \"mixed_dfma_tflops\": %.3f", ms,
               2.0 * 256 * 8 * ITERS * 8.0 * grid / ms / 1e9, 2.0 * 16 * ITERS * 256.0 * grid / ms / 1e9);
        ms = time_ms([&] { k_exp<<<grid * 2, 256>>>(out, 1e-3); });
        printf(", \"exp_gexp_per_s\": %.3f", 4.0 * 512 * 256.0 * grid * 2 / ms / 1e6);
    }
    for (size_t kb : {64, 416, 4096}) {
        size_t n2 = kb * 1024 / 16;
        double2* buf; CK(cudaMalloc(&buf, n2 * 16)); CK(cudaMemset(buf, 0, n2 * 16));
        int reps = (int)std::max<size_t>(4, 65536 / kb);
        for (int cps : {1, 2, 4}) {
            int grid = sms * cps;
            double ms = time_ms([&] { k_l2read<true><<<grid, 256>>>(buf, n2, reps, out); });
            printf(", \"l2read_cg_%zukb_cps%d_gbs\": %.1f", kb, cps, (double)n2 * 16 * reps * grid / ms / 1e6);
            ms = time_ms([&] { k_l2read<false><<<grid, 256>>>(buf, n2, reps, out); });
            printf(", \"l1read_ca_%zukb_cps%d_gbs\": %.1f", kb, cps, (double)n2 * 16 * reps * grid / ms / 1e6);
        }
        CK(cudaFree(buf));
    }
    printf("}\n");
    return 0;
}
