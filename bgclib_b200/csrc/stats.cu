// K7 translation unit: alignment statistics without traceback (gotoh_stats.cuh).
#include "gotoh_stats.cuh"
#include "launchers.h"

namespace bgca {

cudaError_t launch_stats(int mode, const ScoreParams &p, const int32_t *pq, const int32_t *ps,
                         long long npairs, uint4 *scratch, int scratch_stride, bgca_alnstats *out,
                         int grid, cudaStream_t st)
{
    if (mode == BGCA_MODE_LOCAL)
        gotoh_stats_kernel<BGCA_MODE_LOCAL><<<grid, CTA_THREADS, 0, st>>>(p, pq, ps, npairs, scratch,
                                                                         scratch_stride, out);
    else
        gotoh_stats_kernel<BGCA_MODE_GLOBAL><<<grid, CTA_THREADS, 0, st>>>(p, pq, ps, npairs, scratch,
                                                                          scratch_stride, out);
    return cudaGetLastError();
}

}  // namespace bgca
