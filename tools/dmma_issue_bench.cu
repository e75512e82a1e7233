// How many warps per SM sub-partition does DMMA.8x8x4 need to saturate the FP64 pipe?  (B200, sm_100a)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_issue_bench tools/dmma_issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 8192;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void k(double* out, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA interleaved with dependent DFMA chains (like exp) to see co-issue
template <int ILP>
__global__ void kmix(double* out, double a, double b) {
    double c[ILP][2];
    double x = a, y = b;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { dmma884(c[i][0], c[i][1], a, b); x = fma(x, a, b); y = fma(y, b, a); }
    }
    double s = x + y;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 4);
    printf("{\"sms\": %d", sms);
#define RUN(ILP, THREADS) { float ms = time_ms([&] { k<ILP><<<sms, THREADS>>>(out, 1.0000001, 1e-9); }); \
        printf(", \"dmma_ilp%d_warps%d_tflops\": %.2f", ILP, THREADS / 32, 2.0 * 256 * ILP * (double)ITERS * (THREADS / 32) * sms / ms / 1e9); }
    RUN(1, 128) RUN(2, 128) RUN(4, 128) RUN(8, 128) RUN(14, 128)
    RUN(1, 256) RUN(2, 256) RUN(4, 256) RUN(8, 256)
    RUN(1, 512) RUN(2, 512) RUN(4, 512)
    RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32)
#define RUNM(ILP, THREADS) { float ms = time_ms([&] { kmix<ILP><<<sms, THREADS>>>(out, 1.0000001, 1e-9); }); \
        printf(", \"mix_ilp%d_warps%d_dmma_tflops\": %.2f", ILP, THREADS / 32, 2.0 * 256 * ILP * (double)ITERS * (THREADS / 32) * sms / ms / 1e9); }
    RUNM(4, 128) RUNM(4, 256) RUNM(8, 128)
    printf("}\n");
    return 0;
}
