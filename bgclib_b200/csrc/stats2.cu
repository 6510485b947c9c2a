// K7 second form: translation unit and launcher (gotoh_stats2.cuh).
#include "gotoh_stats2.cuh"
#include "launchers.h"

namespace bgca {

template <int MODE, int KS>
static cudaError_t launch_one(const Stats2Params &p, int n_sm, cudaStream_t st)
{
    using Cfg = Stats2Cfg<KS>;
    static int ctas_per_sm = 0;
    auto kern = gotoh_stats2_kernel<MODE, KS>;
    if (ctas_per_sm == 0) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (err != cudaSuccess) return err;
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ST2_THREADS, Cfg::SMEM_BYTES);
        if (err != cudaSuccess) return err;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > p.n_items) grid = p.n_items;
    kern<<<grid, ST2_THREADS, Cfg::SMEM_BYTES, st>>>(p);
    return cudaGetLastError();
}

// rows per lane of the instantiation that covers a query of La residues in one pass
int stats2_rows_per_lane(int La)
{
    static const int ks[] = {BGCA_STATS2_KS_LIST};
    for (int k : ks)
        if (32 * k >= La) return k;
    return 0;
}

cudaError_t launch_stats2(int mode, int KS, const Stats2Params &p, int n_sm, cudaStream_t st)
{
#define BGCA_ST2_CASE(K)                                                                    \
    case K:                                                                                 \
        return mode == BGCA_MODE_LOCAL ? launch_one<BGCA_MODE_LOCAL, K>(p, n_sm, st)       \
                                       : launch_one<BGCA_MODE_GLOBAL, K>(p, n_sm, st);
    switch (KS) {
        BGCA_ST2_CASE(4) BGCA_ST2_CASE(8) BGCA_ST2_CASE(12) BGCA_ST2_CASE(16) BGCA_ST2_CASE(24) BGCA_ST2_CASE(32)
        default:
            return cudaErrorInvalidValue;
    }
#undef BGCA_ST2_CASE
}

}  // namespace bgca
