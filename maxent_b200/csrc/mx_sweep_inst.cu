// One translation unit per singular-space tile count NT (compiled in parallel with -DMX_NT=<n>).
#include "mx_sweep.cuh"
namespace mx {
#define MX_CAT2(a, b) a##b
#define MX_CAT(a, b) MX_CAT2(a, b)
int MX_CAT(sweep_nt, MX_NT)(const SweepArgs& a, cudaStream_t stream, bool query, int* o_t, int* o_smem) {
    return pick_T<MX_NT>(a, stream, query, o_t, o_smem);
}
}  // namespace mx
