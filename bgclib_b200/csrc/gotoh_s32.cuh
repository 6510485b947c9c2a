// K4 — 32-bit fallback DP kernel (any length, any score range), sm_100a.
//
// One warp per (query, subject) pair taken from an explicit list in sorted
// index space.  Same systolic sweep as the packed kernel with 32-bit DPX ops
// (VIADDMNMX[.RELU]); lane t owns K32 query rows, a pass covers 32*K32 rows and
// longer queries take several passes, the bottom row of a pass (H, F^ per
// subject column) parked in a per-warp scratch line in HBM/L2.  No profile:
// the 24x24 matrix sits in shared memory and each lane keeps its rows' offsets.
// This path exists for correctness at the edges (sequences longer than the
// packed kernel's 1024-row pass, or whose score bound leaves int16); it is not
// the throughput path.
#pragma once
#include "common.cuh"

namespace bgca {

constexpr int K32 = 16;
constexpr int PASS32 = 32 * K32;
constexpr int NEG32 = -(1 << 28);

template <int MODE>
__global__ void __launch_bounds__(CTA_THREADS)
gotoh_s32_kernel(const ScoreParams p, const int32_t *__restrict__ pq, const int32_t *__restrict__ ps,
                 long long npairs, int2 *scratch, int scratch_stride,
                 int32_t *out_list /* nullable: out[pair] instead of the epilogue */)
{
    constexpr bool LOCAL = (MODE == BGCA_MODE_LOCAL);
    constexpr unsigned FULL = 0xFFFFFFFFu;
    __shared__ int8_t ssub[NSYM * NSYM];
    for (int i = threadIdx.x; i < NSYM * NSYM; i += CTA_THREADS) ssub[i] = p.sub[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (CTA_THREADS / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (CTA_THREADS / 32);
    int2 *bnd = scratch + w * scratch_stride;
    const int open = p.open, ext = p.extend;

    for (long long pair = w; pair < npairs; pair += nw) {
        const int q = pq[pair], s = ps[pair];
        const int La = p.seq_len[q], Ls = p.seq_len[s];
        const uint8_t *qp = p.codes + p.seq_off[q];
        const uint8_t *sp = p.codes + p.seq_off[s];
        int best = 0, cap = 0;
        const int npass = (La + PASS32 - 1) / PASS32;

        for (int pass = 0; pass < npass; ++pass) {
            const int i0 = pass * PASS32 + lane * K32;
            int qrow[K32], Hp[K32], Eh[K32];
#pragma unroll
            for (int k = 0; k < K32; ++k) {
                qrow[k] = (i0 + k < La) ? (int)qp[i0 + k] * NSYM : -1;
                Hp[k] = LOCAL ? 0 : open + (i0 + k) * ext;
                Eh[k] = LOCAL ? 0 : NEG32;
            }
            int hin_prev = (LOCAL || i0 == 0) ? 0 : open + (i0 - 1) * ext;
            int h_out = 0, f_out = 0;
            int j = -lane;
            int c_next = ((unsigned)j < (unsigned)Ls) ? (int)__ldg(sp + j) : 0;
            int2 b_next = make_int2(0, 0);
            if (lane == 0 && pass > 0) b_next = __ldcg(bnd);
            const bool park = (pass + 1 < npass);

            for (int tau = 0; tau < Ls + 31; ++tau, ++j) {
                int h_in = __shfl_up_sync(FULL, h_out, 1);
                int f_in = __shfl_up_sync(FULL, f_out, 1);
                const int c = c_next;
                c_next = ((unsigned)(j + 1) < (unsigned)Ls) ? (int)__ldg(sp + j + 1) : 0;
                if (lane == 0) {
                    if (pass == 0) {
                        h_in = LOCAL ? 0 : open + j * ext;
                        f_in = LOCAL ? 0 : NEG32;
                    } else {
                        h_in = b_next.x;
                        f_in = b_next.y;
                        if (j + 1 < Ls) b_next = __ldcg(bnd + j + 1);
                    }
                }
                if ((unsigned)j < (unsigned)Ls) {
                    int diag = hin_prev;
                    hin_prev = h_in;
                    int up = h_in, F = f_in;
#pragma unroll
                    for (int k = 0; k < K32; ++k) {
                        const int sc = (qrow[k] >= 0) ? (int)ssub[qrow[k] + c] : 0;
                        Eh[k] = __viaddmax_s32(Eh[k], ext, Hp[k]);
                        F = __viaddmax_s32(F, ext, up);
                        const int u = max(Eh[k], F);
                        const int tt = diag + sc;
                        const int h = LOCAL ? __viaddmax_s32_relu(u, open, tt) : __viaddmax_s32(u, open, tt);
                        diag = Hp[k];
                        Hp[k] = h;
                        up = h;
                        if (LOCAL) best = max(best, h);
                    }
                    h_out = up;
                    f_out = F;
                    if (park && lane == 31) __stcg(bnd + j, make_int2(up, F));
                    if (!LOCAL && j == Ls - 1) {
#pragma unroll
                        for (int k = 0; k < K32; ++k)
                            if (i0 + k == La - 1) cap = Hp[k];
                    }
                }
            }
            __syncwarp();   // parked row of this pass is visible to lane 0 of the next
        }

        int r;
        if (LOCAL) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(FULL, best, off));
            r = best;
        } else {
            r = __shfl_sync(FULL, cap, ((La - 1) % PASS32) / K32);
        }
        if (lane == 0) {
            if (out_list) out_list[pair] = r;
            else emit_pair(p, min(q, s), max(q, s), r);
        }
    }
}

}  // namespace bgca
