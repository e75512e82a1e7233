// Dispatch of the fused sweep over the compiled (NT, T) instantiations.
#include "mx_common.cuh"
namespace mx {
int sweep_nt4(const SweepArgs&, cudaStream_t, bool, int*, int*);
int sweep_nt5(const SweepArgs&, cudaStream_t, bool, int*, int*);
int sweep_nt6(const SweepArgs&, cudaStream_t, bool, int*, int*);
int sweep_nt7(const SweepArgs&, cudaStream_t, bool, int*, int*);
int sweep_nt8(const SweepArgs&, cudaStream_t, bool, int*, int*);

int dispatch_sweep(SweepArgs& a, cudaStream_t stream, bool query, int* o_t, int* o_smem) {
    const int s = a.n_sv;
    if (s < 1) return MX_ERR_BAD_ARG;
    int pk = (s * (s + 1)) / 2;
    pk = (pk + 1) & ~1;
    a.pk = pk;
    const int nt = (s + 7) / 8;
    switch (nt) {
        case 1: case 2: case 3: case 4: return sweep_nt4(a, stream, query, o_t, o_smem);
        case 5: return sweep_nt5(a, stream, query, o_t, o_smem);
        case 6: return sweep_nt6(a, stream, query, o_t, o_smem);
        case 7: return sweep_nt7(a, stream, query, o_t, o_smem);
        case 8: return sweep_nt8(a, stream, query, o_t, o_smem);
        default: return MX_ERR_UNSUPPORTED;
    }
}
}  // namespace mx
