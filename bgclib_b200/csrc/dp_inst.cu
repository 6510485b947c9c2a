// Instantiations of the packed DP kernel for one (G, MODE, PRESET); compiled 18
// times (G in {8,16,32} x MODE in {local, global} x gap preset) so the build
// parallelises.
#include "gotoh_pair16.cuh"
#include "launchers.h"
#include "variants.h"

#if !defined(BGCA_INST_G) || !defined(BGCA_INST_MODE) || !defined(BGCA_INST_PRESET)
#error "define BGCA_INST_G, BGCA_INST_MODE and BGCA_INST_PRESET"
#endif

namespace bgca {

template <int G, int K, int MODE, int PRESET>
static cudaError_t launch_one(const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out)
{
    using Cfg = Pair16Cfg<G, K>;
    static int ctas_per_sm = 0;   // per instantiation, resolved on first use
    auto kern = gotoh_pair16_kernel<G, K, MODE, PRESET>;
    if (ctas_per_sm == 0) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               Cfg::SMEM_BYTES);
        if (err != cudaSuccess) return err;
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA_THREADS, Cfg::SMEM_BYTES);
        if (err != cudaSuccess) return err;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid_out) *grid_out = grid;
    kern<<<grid, CTA_THREADS, Cfg::SMEM_BYTES, st>>>(p);
    return cudaGetLastError();
}

#define BGCA_CAT2(a, b, c, d, e, f) a##b##c##d##e##f
#define BGCA_CAT(a, b, c, d, e, f) BGCA_CAT2(a, b, c, d, e, f)

cudaError_t BGCA_CAT(launch_pair16_g, BGCA_INST_G, _m, BGCA_INST_MODE, _p, BGCA_INST_PRESET)(
    int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out)
{
    switch (K) {
#define X(G, KK) \
    case KK:     \
        return launch_one<BGCA_INST_G, KK, BGCA_INST_MODE, BGCA_INST_PRESET>(p, n_sm, st, grid_out);
        BGCA_K_LIST(X, BGCA_INST_G)
#if BGCA_INST_G == 32
        BGCA_K_LIST_WIDE(X, BGCA_INST_G)
#endif
#undef X
        default:
            return cudaErrorInvalidValue;
    }
}

}  // namespace bgca
