// Checks that mx::exp_main (csrc/mx_common.cuh) is bit-identical to the CUDA library exp() on its domain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -I maxent_b200/csrc tools/exp_check.cu -o tools/exp_check && tools/exp_check
#include <cstdio>
#include <cstring>
#include "mx_common.cuh"
__global__ void check(unsigned long long seed, unsigned long long* nbad, unsigned long long* ntested, double* worst) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0, n = 0;
    for (int i = 0; i < 4096; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const double u = (double)(s >> 11) / 9007199254740992.0;
        double x;
        switch (i & 3) {
            case 0: x = -60.0 + 70.0 * u; break;
            case 1: x = -708.39 + 1416.78 * u; break;
            case 2: x = (u - 0.5) * 1e-3; break;
            default: x = __longlong_as_double((long long)(s >> 1)); break;      // arbitrary bit patterns
        }
        if (mx::exp_is_special(x) || x != x) continue;
        ++n;
        const double a = mx::exp_main(x), b = exp(x);
        if (__double_as_longlong(a) != __double_as_longlong(b)) { ++bad; *worst = x; }
    }
    atomicAdd(nbad, bad); atomicAdd(ntested, n);
}
int main() {
    unsigned long long *nbad, *nt; double* worst;
    cudaMallocManaged(&nbad, 8); cudaMallocManaged(&nt, 8); cudaMallocManaged(&worst, 8);
    *nbad = 0; *nt = 0; *worst = 0;
    check<<<592, 256>>>(12345, nbad, nt, worst);
    cudaDeviceSynchronize();
    printf("exp_main vs exp: %llu mismatches in %llu arguments (last mismatch at x=%.17g)\n", *nbad, *nt, *worst);
    return *nbad ? 1 : 0;
}
