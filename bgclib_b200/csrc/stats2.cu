// K7 second form: translation unit and launcher (gotoh_stats2.cuh).
#include "gotoh_stats2.cuh"
#include "launchers.h"

namespace bgca {

template <int MODE, int KS, int HINT>
static cudaError_t launch_one(const Stats2Params &p, int n_sm, cudaStream_t st)
{
    using Cfg = Stats2Cfg<KS>;
    static int ctas_per_sm = 0;
    auto kern = gotoh_stats2_kernel<MODE, KS, HINT>;
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

// p.hint != NULL (local mode only): the kernels that take the best score of every pair as known
cudaError_t launch_stats2(int mode, int KS, const Stats2Params &p, int n_sm, cudaStream_t st)
{
#define BGCA_ST2_CASE(K)                                                                               \
    case K:                                                                                            \
        return mode != BGCA_MODE_LOCAL ? launch_one<BGCA_MODE_GLOBAL, K, 0>(p, n_sm, st)              \
               : p.hint              ? launch_one<BGCA_MODE_LOCAL, K, 1>(p, n_sm, st)                 \
                                     : launch_one<BGCA_MODE_LOCAL, K, 0>(p, n_sm, st);
    switch (KS) {
        BGCA_ST2_CASE(4) BGCA_ST2_CASE(6) BGCA_ST2_CASE(8) BGCA_ST2_CASE(10) BGCA_ST2_CASE(12) BGCA_ST2_CASE(14)
        BGCA_ST2_CASE(16) BGCA_ST2_CASE(20) BGCA_ST2_CASE(24) BGCA_ST2_CASE(28) BGCA_ST2_CASE(32)
        default:
            return cudaErrorInvalidValue;
    }
#undef BGCA_ST2_CASE
}

}  // namespace bgca

// ---------------------------------------------------------------------------
// Pair selection on the device (statistics "for the pairs that are similar at
// all"): from the n x n score matrix of the scoring kernels (ORIGINAL index space) every pair
// i < j with score >= min_score [and the same domain type], in row-major order — which is
// already the grouping by query that K7 wants — with its score as the hint, without the list
// ever visiting the host.
// ---------------------------------------------------------------------------
namespace bgca {

__device__ __forceinline__ bool sel_keep(const int32_t *__restrict__ scores, int n, int i, int j, int min_score,
                                         const int32_t *__restrict__ type_sorted, const int32_t *__restrict__ inv,
                                         int type_i)
{
    if (j <= i || j >= n) return false;
    if (scores[(size_t)i * n + j] < min_score) return false;
    return !type_sorted || type_sorted[inv[j]] == type_i;
}

__global__ void __launch_bounds__(256)
select_count_kernel(const int32_t *__restrict__ scores, int n, int min_score, const int32_t *__restrict__ type_sorted,
                    const int32_t *__restrict__ inv, int32_t *__restrict__ cnt)
{
    const int i = blockIdx.x;
    const int type_i = type_sorted ? type_sorted[inv[i]] : 0;
    int c = 0;
    for (int j = i + 1 + threadIdx.x; j < n; j += 256) c += sel_keep(scores, n, i, j, min_score, type_sorted, inv, type_i);
    __shared__ int part[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        cnt[i] = t;
    }
}

// one CTA: row_start[0..n] = exclusive prefix sums of cnt; cls_items[c] = work items of class c
__global__ void __launch_bounds__(1024)
select_scan_kernel(const int32_t *__restrict__ cnt, int n, const int32_t *__restrict__ len_sorted,
                   const int32_t *__restrict__ inv, int item_pairs, int64_t *__restrict__ row_start,
                   int32_t *__restrict__ cls_items)
{
    __shared__ long long carry;
    __shared__ long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int ks_list[] = {BGCA_STATS2_KS_LIST};
    constexpr int n_cls = sizeof(ks_list) / sizeof(ks_list[0]);
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = (i < n) ? cnt[i] : 0;
        long long x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long w = wsum[threadIdx.x];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long y = __shfl_up_sync(0xFFFFFFFFu, w, off);
                if (threadIdx.x >= off) w += y;
            }
            wsum[threadIdx.x] = w;
        }
        __syncthreads();
        const long long before = carry + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < n) {
            row_start[i] = before;
            if (v > 0 && cls_items) {
                const int L = len_sorted[inv[i]];
                int c = 0;
                while (c < n_cls - 1 && 32 * ks_list[c] < L) ++c;
                atomicAdd(&cls_items[c], (int)((v + item_pairs - 1) / item_pairs));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) row_start[n] = carry;
}

// CTA per row: the kept columns in ascending order behind row_start[i]; work items of the row
__global__ void __launch_bounds__(256)
select_fill_kernel(const int32_t *__restrict__ scores, int n, int min_score, const int32_t *__restrict__ type_sorted,
                   const int32_t *__restrict__ inv, const int32_t *__restrict__ len_sorted,
                   const int64_t *__restrict__ row_start, int item_pairs, const int32_t *__restrict__ cls_base,
                   int32_t *__restrict__ cls_fill, int32_t *__restrict__ pairs_out, int32_t *__restrict__ q_sorted,
                   int32_t *__restrict__ s_sorted, int32_t *__restrict__ hint, Stats2Item *__restrict__ items)
{
    const int i = blockIdx.x;
    const long long first = row_start[i];
    const int total = (int)(row_start[i + 1] - first);
    if (total == 0) return;
    const int qi = inv[i];
    const int type_i = type_sorted ? type_sorted[qi] : 0;
    __shared__ int wcount[8];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = i + 1; base < n; base += 256) {
        const int j = base + threadIdx.x;
        const bool keep = sel_keep(scores, n, i, j, min_score, type_sorted, inv, type_i);
        const unsigned b = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) wcount[warp] = __popc(b);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += wcount[w];
        if (keep) {
            const long long pos = first + before + __popc(b & ((1u << lane) - 1u));
            if (pairs_out) { pairs_out[2 * pos] = i; pairs_out[2 * pos + 1] = j; }
            q_sorted[pos] = qi;
            s_sorted[pos] = inv[j];
            hint[pos] = scores[(size_t)i * n + j];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += wcount[w];
            running += t;
        }
        __syncthreads();
    }
    if (items && threadIdx.x == 0) {
        const int ks_list[] = {BGCA_STATS2_KS_LIST};
        constexpr int n_cls = sizeof(ks_list) / sizeof(ks_list[0]);
        const int L = len_sorted[qi];
        int c = 0;
        while (c < n_cls - 1 && 32 * ks_list[c] < L) ++c;
        const int n_items = (total + item_pairs - 1) / item_pairs;
        const int at = cls_base[c] + atomicAdd(&cls_fill[c], n_items);
        for (int t = 0; t < n_items; ++t)
            items[at + t] = Stats2Item{qi, (int32_t)(first + (long long)t * item_pairs),
                                       min(item_pairs, total - t * item_pairs), 0};
    }
}

cudaError_t launch_select_count(const int32_t *scores, int n, int min_score, const int32_t *type_sorted,
                                const int32_t *inv, const int32_t *len_sorted, int item_pairs, int32_t *cnt,
                                int64_t *row_start, int32_t *cls_items, cudaStream_t st)
{
    select_count_kernel<<<n, 256, 0, st>>>(scores, n, min_score, type_sorted, inv, cnt);
    select_scan_kernel<<<1, 1024, 0, st>>>(cnt, n, len_sorted, inv, item_pairs, row_start, cls_items);
    return cudaGetLastError();
}

cudaError_t launch_select_fill(const int32_t *scores, int n, int min_score, const int32_t *type_sorted,
                               const int32_t *inv, const int32_t *len_sorted, const int64_t *row_start,
                               int item_pairs, const int32_t *cls_base, int32_t *cls_fill, int32_t *pairs_out,
                               int32_t *q_sorted, int32_t *s_sorted, int32_t *hint, Stats2Item *items,
                               cudaStream_t st)
{
    select_fill_kernel<<<n, 256, 0, st>>>(scores, n, min_score, type_sorted, inv, len_sorted, row_start, item_pairs,
                                          cls_base, cls_fill, pairs_out, q_sorted, s_sorted, hint, items);
    return cudaGetLastError();
}

}  // namespace bgca
