// One translation unit per singular-space tile count NT of the spectrum-per-CTA sweep (-DMX_NT=<n>).
#include "mx_sweep2.cuh"
namespace mx2 {
#define MX_CAT2(a, b) a##b
#define MX_CAT(a, b) MX_CAT2(a, b)
int MX_CAT(sweep2_nt, MX_NT)(const mx::SweepArgs& a, cudaStream_t stream, bool query, int* o_smem, int* o_grid) {
    return launch_sweep2<MX_NT>(a, stream, query, o_smem, o_grid);
}
#if MX_NT == 7
int sweep2_threads() { return threads_per_cta(); }
long long sweep2_wide_doubles(int nt) { return wide_doubles(nt); }
#endif
}  // namespace mx2
