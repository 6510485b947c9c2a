// Host-side launch entry points of the per-(G, MODE) kernel translation units.
#pragma once
#include "common.cuh"

namespace bgca {

using Pair16Launcher = cudaError_t (*)(int K, const ScoreParams &p, int n_sm, cudaStream_t st,
                                       int *grid_out);
#define BGCA_DECL_LAUNCH(G, M, P) \
    cudaError_t launch_pair16_g##G##_m##M##_p##P(int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out);
#define BGCA_DECL_LAUNCH_GM(G, M) BGCA_DECL_LAUNCH(G, M, 0) BGCA_DECL_LAUNCH(G, M, 1) BGCA_DECL_LAUNCH(G, M, 2)
BGCA_DECL_LAUNCH_GM(1, 0)
BGCA_DECL_LAUNCH_GM(1, 1)
BGCA_DECL_LAUNCH_GM(2, 0)
BGCA_DECL_LAUNCH_GM(2, 1)
BGCA_DECL_LAUNCH_GM(4, 0)
BGCA_DECL_LAUNCH_GM(4, 1)
BGCA_DECL_LAUNCH_GM(8, 0)
BGCA_DECL_LAUNCH_GM(8, 1)
BGCA_DECL_LAUNCH_GM(16, 0)
BGCA_DECL_LAUNCH_GM(16, 1)
BGCA_DECL_LAUNCH_GM(32, 0)
BGCA_DECL_LAUNCH_GM(32, 1)
// multi-pass (G = 32) instantiations
#define BGCA_DECL_LAUNCH_MP(M, P) \
    cudaError_t launch_pair16_mp_m##M##_p##P(int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out); \
    cudaError_t launch_pair16_mpw_m##M##_p##P(int K, const ScoreParams &p, int n_sm, cudaStream_t st, int *grid_out);
BGCA_DECL_LAUNCH_MP(0, 0) BGCA_DECL_LAUNCH_MP(0, 1) BGCA_DECL_LAUNCH_MP(0, 2)
BGCA_DECL_LAUNCH_MP(1, 0) BGCA_DECL_LAUNCH_MP(1, 1) BGCA_DECL_LAUNCH_MP(1, 2)
#undef BGCA_DECL_LAUNCH_MP
// CTAs the multi-pass kernels are launched with at most (sizes their scratch lines)
int pair16_mp_max_grid(int n_sm);
#undef BGCA_DECL_LAUNCH_GM
#undef BGCA_DECL_LAUNCH

cudaError_t launch_s32(int mode, const ScoreParams &p, const int32_t *pq, const int32_t *ps,
                       long long npairs, int2 *scratch, int scratch_stride, int32_t *out_list,
                       int grid, cudaStream_t st);

// K7 (stats.cu)
cudaError_t launch_stats(int mode, int compact, const ScoreParams &p, const int32_t *pq, const int32_t *ps,
                         long long npairs, uint4 *scratch, int scratch_stride, bgca_alnstats *out,
                         int grid, cudaStream_t st);

// K7 second form (stats2.cu): sequences of at most 1023 residues, pairs grouped by query
#define BGCA_STATS2_KS_LIST 4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32
struct Stats2Params;
struct Stats2Item;
int stats2_rows_per_lane(int La);
cudaError_t launch_stats2(int mode, int KS, const Stats2Params &p, int n_sm, cudaStream_t st);
// pair selection by score on the device (stats2.cu)
cudaError_t launch_select_count(const int32_t *scores, int n, int min_score, const int32_t *type_sorted,
                                const int32_t *inv, const int32_t *len_sorted, int item_pairs, int32_t *cnt,
                                int64_t *row_start, int32_t *cls_items, cudaStream_t st);
cudaError_t launch_select_fill(const int32_t *scores, int n, int min_score, const int32_t *type_sorted,
                               const int32_t *inv, const int32_t *len_sorted, const int64_t *row_start,
                               int item_pairs, const int32_t *cls_base, int32_t *cls_fill, int32_t *pairs_out,
                               int32_t *q_sorted, int32_t *s_sorted, int32_t *hint, Stats2Item *items,
                               cudaStream_t st);

// K1 / K5 kernels (pack_reduce.cu)
cudaError_t launch_pack(const uint8_t *ascii, const int64_t *src_off_sorted, const int32_t *order,
                        const uint32_t *seq_off, const int32_t *seq_len, int32_t n, uint8_t *codes,
                        unsigned long long *err_word, cudaStream_t st);
cudaError_t launch_bgc_reduce(const int32_t *rowbest, const int32_t *selfsc, const int32_t *bgc_start,
                              const int32_t *bgc_doms, int32_t n_bgc, int64_t *best, int64_t *selfsum,
                              double *sim, cudaStream_t st);

// sorted-space int16 score matrix -> caller's int32 matrix (pack_reduce.cu)
#define BGCA_S16_UNWRITTEN 0x8080          /* cudaMemset(0x80): below every score the s16x2 kernels can report */
constexpr int kPermuteMaxN = 100000;       // a sorted row of n int16 has to fit in shared memory
cudaError_t launch_permute16(const int16_t *S, int stride, const int32_t *inv, int n, int32_t *out,
                             int skip_unwritten, int n_sm, cudaStream_t st);

int dpx_probe_run(int device, double *out, int n);

}  // namespace bgca
