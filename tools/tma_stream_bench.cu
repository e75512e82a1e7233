// How fast can every SM of a B200 stream the same L2-resident buffer (the re-tiled V', 448 KB) into shared memory
// with 1-D TMA bulk copies?  One thread per CTA issues the copies into a ring of `depth` stages of `chunk` bytes and
// waits for each on its mbarrier; nothing is computed.  Sweeps ring depth, chunk size, CTAs per SM and the number of
// replicas of the buffer (CTA b reads replica b % replicas: does it matter that all SMs hit the same L2 lines?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_bench tools/tma_stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void stream_kernel(const char* buf, size_t bytes, int replicas, int chunk, int depth, int passes, long long* cyc) {
    extern __shared__ __align__(128) char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    char* stage0 = sm + 128;
    const char* src = buf + (size_t)(blockIdx.x % replicas) * bytes;
    const int nch = (int)(bytes / chunk);
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bars + i)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    unsigned g = 0;          // chunks issued so far (global), gw = chunks waited for
    unsigned gw = 0;
    const unsigned total = (unsigned)nch * passes;
    auto issue = [&](unsigned gi) {
        const int st = gi % depth;
        const int c = gi % nch;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bars + st)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(smem_u32(stage0 + (size_t)st * chunk)), "l"(src + (size_t)c * chunk), "r"(chunk), "r"(smem_u32(bars + st)) : "memory");
    };
    for (; g < (unsigned)depth && g < total; ++g) issue(g);
    for (; gw < total; ++gw) {
        const int st = gw % depth;
        const uint32_t parity = (gw / depth) & 1;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra WAIT_LOOP;\n"
            "DONE:\n"
            "}\n" ::"r"(smem_u32(bars + st)), "r"(parity) : "memory");
        if (g < total) { issue(g); ++g; }
    }
    cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t bytes = 125 * 7 * 64 * 8;      // 448,000 B: V' of the benchmark shape
    const int max_rep = 16;
    char* buf;
    long long* cyc;
    CK(cudaMalloc(&buf, bytes * max_rep));
    CK(cudaMemset(buf, 1, bytes * max_rep));
    CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int passes = 40;
    printf("[");
    bool first = true;
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
        for (int replicas : {1, 4, 16})
            for (int chunk : {3584, 7168, 14336, 28672})
                for (int depth : {2, 3, 4, 6, 8}) {
                    size_t smem = 128 + (size_t)chunk * depth;
                    if (smem > 100 * 1024) continue;
                    // occupancy: force `per_sm` CTAs per SM by padding the dynamic shared memory request
                    size_t req = per_sm == 1 ? 120 * 1024 : 100 * 1024;
                    if (req < smem) req = smem;
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    stream_kernel<<<sms * per_sm, 32, req>>>(buf, bytes, replicas, chunk, depth, 2, cyc);
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    stream_kernel<<<sms * per_sm, 32, req>>>(buf, bytes, replicas, chunk, depth, passes, cyc);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double us_per_pass = ms * 1e3 / passes;
                    const double gbs_per_cta = bytes / (us_per_pass * 1e-6) / 1e9;
                    printf("%s{\"ctas_per_sm\": %d, \"replicas\": %d, \"chunk\": %d, \"depth\": %d, \"us_per_448KB_pass\": %.2f, "
                           "\"GBs_per_cta\": %.1f, \"TBs_chip\": %.2f}", first ? "" : ",\n ", per_sm, replicas, chunk, depth,
                           us_per_pass, gbs_per_cta, gbs_per_cta * sms * per_sm / 1e3);
                    first = false;
                }
    printf("]\n");
    return 0;
}
