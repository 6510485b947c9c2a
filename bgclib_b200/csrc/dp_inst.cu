// Instantiations of the packed DP kernel for one (G, MODE, PRESET); compiled once per
// (G, MODE in {local, global}, gap preset) — and once more per (MODE, preset) with
// -DBGCA_INST_MP=1 for the multi-pass whole-warp form — so the build parallelises.
#include "gotoh_pair16.cuh"
#include "launchers.h"
#include "variants.h"

#if !defined(BGCA_INST_G) || !defined(BGCA_INST_MODE) || !defined(BGCA_INST_PRESET)
#error "define BGCA_INST_G, BGCA_INST_MODE and BGCA_INST_PRESET"
#endif

#ifndef BGCA_INST_MP
#define BGCA_INST_MP 0
#endif
#ifndef BGCA_INST_W32
#define BGCA_INST_W32 0      // 1 (with BGCA_INST_MP): one alignment per 32-bit word, any score range
#endif

namespace bgca {

constexpr int kMpMaxCtasPerSm = 3;   // the multi-pass kernels' scratch is sized for this many (K <= 18 runs three)

template <int G, int K, int MODE, int PRESET, int MP, int W32 = 0>
static cudaError_t launch_one(const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out)
{
    using Cfg = Pair16Cfg<G, K>;
    constexpr int SMEM = MP ? Cfg::SMEM_BYTES_MP : Cfg::SMEM_BYTES;
    static int ctas_per_sm = 0;   // per instantiation, resolved on first use
    auto kern = gotoh_pair16_kernel<G, K, MODE, PRESET, MP, W32>;
    if (ctas_per_sm == 0) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (err != cudaSuccess) return err;
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA_THREADS, SMEM);
        if (err != cudaSuccess) return err;
        ctas_per_sm = occ > 0 ? occ : 1;
        if (MP && ctas_per_sm > kMpMaxCtasPerSm) ctas_per_sm = kMpMaxCtasPerSm;
    }
    int grid = n_sm * ctas_per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid_out) *grid_out = grid;
    kern<<<grid, CTA_THREADS, SMEM, st>>>(p);
    return cudaGetLastError();
}

#define BGCA_CAT2(a, b, c, d, e, f) a##b##c##d##e##f
#define BGCA_CAT(a, b, c, d, e, f) BGCA_CAT2(a, b, c, d, e, f)

#if BGCA_INST_MP
#if BGCA_INST_W32
#define BGCA_MP_NAME launch_pair16_mpw
#else
#define BGCA_MP_NAME launch_pair16_mp
#endif
cudaError_t BGCA_CAT(BGCA_MP_NAME, , _m, BGCA_INST_MODE, _p, BGCA_INST_PRESET)(
    int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out)
{
    switch (K) {
#define X(G, KK) \
    case KK:     \
        return launch_one<32, KK, BGCA_INST_MODE, BGCA_INST_PRESET, 1, BGCA_INST_W32>(p, n_sm, st, grid_out);
        BGCA_K_LIST_MP(X, 32)
#undef X
        default:
            return cudaErrorInvalidValue;
    }
}
#if BGCA_INST_MODE == 0 && BGCA_INST_PRESET == 0 && !BGCA_INST_W32
int pair16_mp_max_grid(int n_sm) { return n_sm * kMpMaxCtasPerSm; }
#endif
#else
cudaError_t BGCA_CAT(launch_pair16_g, BGCA_INST_G, _m, BGCA_INST_MODE, _p, BGCA_INST_PRESET)(
    int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out)
{
    switch (K) {
#define X(G, KK) \
    case KK:     \
        return launch_one<BGCA_INST_G, KK, BGCA_INST_MODE, BGCA_INST_PRESET, 0>(p, n_sm, st, grid_out);
        BGCA_K_LIST(X, BGCA_INST_G)
#undef X
        default:
            return cudaErrorInvalidValue;
    }
}
#endif

}  // namespace bgca
