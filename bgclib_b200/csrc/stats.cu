// K7 translation unit: alignment statistics without traceback (gotoh_stats.cuh).
#include "gotoh_stats.cuh"
#include "launchers.h"

namespace bgca {

// compact != 0: every sequence in the pair list is at most 1023 residues long (Tup64 fits)
cudaError_t launch_stats(int mode, int compact, const ScoreParams &p, const int32_t *pq, const int32_t *ps,
                         long long npairs, uint4 *scratch, int scratch_stride, bgca_alnstats *out,
                         int grid, cudaStream_t st)
{
#define BGCA_LAUNCH_STATS(M, T) \
    gotoh_stats_kernel<M, T><<<grid, CTA_THREADS, 0, st>>>(p, pq, ps, npairs, scratch, scratch_stride, out)
    if (mode == BGCA_MODE_LOCAL) {
        if (compact) BGCA_LAUNCH_STATS(BGCA_MODE_LOCAL, Tup64);
        else BGCA_LAUNCH_STATS(BGCA_MODE_LOCAL, Tup);
    } else {
        if (compact) BGCA_LAUNCH_STATS(BGCA_MODE_GLOBAL, Tup64);
        else BGCA_LAUNCH_STATS(BGCA_MODE_GLOBAL, Tup);
    }
#undef BGCA_LAUNCH_STATS
    return cudaGetLastError();
}

}  // namespace bgca
