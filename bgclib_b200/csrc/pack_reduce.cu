// K1 (packer), K4 launcher, K5 (BGC reduction) — the HBM-bound stages.
#include "gotoh_s32.cuh"
#include "launchers.h"

namespace bgca {

// ---------------------------------------------------------------------------
// K1: ASCII -> residue codes, written in SORTED order, each sequence followed by
// at least one BGCA_PAD_CODE byte (the DP kernels read it as the end-of-subject
// token) and padded to a 16-byte multiple.  One warp per sequence; each lane
// encodes 16 residues and issues one 128-bit store, so the write side is
// fully coalesced; the read side is byte loads that L1 merges into full lines.
// Algorithmic bytes: sum(L) read + sum(roundup16(L+1)) written.
// A letter outside the alphabet is reported through err_word =
// (original sequence index << 32) | position, smallest wins; ~0 = no error.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t encode_residue(uint32_t ch)
{
    // ARNDCQEGHILKMFPSTWYVBZX*  -> 0..23 ; anything else -> 255
    switch (ch) {
        case 'A': return 0;  case 'R': return 1;  case 'N': return 2;  case 'D': return 3;
        case 'C': return 4;  case 'Q': return 5;  case 'E': return 6;  case 'G': return 7;
        case 'H': return 8;  case 'I': return 9;  case 'L': return 10; case 'K': return 11;
        case 'M': return 12; case 'F': return 13; case 'P': return 14; case 'S': return 15;
        case 'T': return 16; case 'W': return 17; case 'Y': return 18; case 'V': return 19;
        case 'B': return 20; case 'Z': return 21; case 'X': return 22; case '*': return 23;
        default: return 255;
    }
}

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ offsets_orig,
            const int32_t *__restrict__ order, const uint32_t *__restrict__ seq_off,
            const int32_t *__restrict__ seq_len, int32_t n, uint8_t *__restrict__ codes,
            unsigned long long *err_word)
{
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)encode_residue(threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps_per_grid) {
        const int orig = order[r];
        const uint8_t *src = ascii + offsets_orig[orig];
        const int len = seq_len[r];
        const int padded = (len + 1 + 15) & ~15;
        uint4 *dst = reinterpret_cast<uint4 *>(codes + seq_off[r]);
        for (int base = lane * 16; base < padded; base += 32 * 16) {
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int i = base + b;
                uint32_t code = BGCA_PAD_CODE;
                if (i < len) {
                    code = lut[__ldg(src + i)];
                    if (code == 255u) {
                        atomicMin(err_word, ((unsigned long long)(uint32_t)orig << 32) |
                                                (unsigned long long)(uint32_t)i);
                        code = 22;  // 'X' keeps the batch well-formed; the call still fails
                    }
                }
                w[b >> 2] |= code << (8 * (b & 3));
            }
            dst[base >> 4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

cudaError_t launch_pack(const uint8_t *ascii, const int64_t *offsets_orig, const int32_t *order,
                        const uint32_t *seq_off, const int32_t *seq_len, int32_t n, uint8_t *codes,
                        unsigned long long *err_word, cudaStream_t st)
{
    int grid = (n + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    pack_kernel<<<grid, 256, 0, st>>>(ascii, offsets_orig, order, seq_off, seq_len, n, codes, err_word);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K4 launcher
// ---------------------------------------------------------------------------
cudaError_t launch_s32(int mode, const ScoreParams &p, const int32_t *pq, const int32_t *ps,
                       long long npairs, int2 *scratch, int scratch_stride, int32_t *out_list,
                       int grid, cudaStream_t st)
{
    if (mode == BGCA_MODE_LOCAL)
        gotoh_s32_kernel<BGCA_MODE_LOCAL><<<grid, CTA_THREADS, 0, st>>>(p, pq, ps, npairs, scratch,
                                                                       scratch_stride, out_list);
    else
        gotoh_s32_kernel<BGCA_MODE_GLOBAL><<<grid, CTA_THREADS, 0, st>>>(p, pq, ps, npairs, scratch,
                                                                        scratch_stride, out_list);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K5 second stage.  best[a][b] = sum over the domains d of BGC a of
// rowbest[d][b] (int64).  Grid (ceil(B/1024), B): a thread owns 4 adjacent
// columns (one 128-bit load per row), rows are walked 4 at a time so 4 loads are
// in flight per thread; every warp reads 512 contiguous bytes per row.
// Algorithmic bytes: 4*N*B read + 8*B*B written.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bgc_sum_kernel(const int32_t *__restrict__ rowbest, const int32_t *__restrict__ selfsc,
               const int32_t *__restrict__ bgc_start, const int32_t *__restrict__ bgc_doms,
               int32_t n_bgc, int64_t *__restrict__ best, int64_t *__restrict__ selfsum)
{
    const int a = blockIdx.y;
    const int b0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int d0 = bgc_start[a], d1 = bgc_start[a + 1];
    if (b0 < n_bgc) {
        long long acc[4] = {0, 0, 0, 0};
        if ((n_bgc & 3) == 0) {             // rows are 16-byte aligned: vector loads
            int i = d0;
            for (; i + 4 <= d1; i += 4) {
                int4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = __ldg(reinterpret_cast<const int4 *>(rowbest + (size_t)bgc_doms[i + u] * n_bgc + b0));
#pragma unroll
                for (int u = 0; u < 4; ++u) { acc[0] += v[u].x; acc[1] += v[u].y; acc[2] += v[u].z; acc[3] += v[u].w; }
            }
            for (; i < d1; ++i) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(rowbest + (size_t)bgc_doms[i] * n_bgc + b0));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            }
        } else {
            for (int i = d0; i < d1; ++i) {
                const int32_t *row = rowbest + (size_t)bgc_doms[i] * n_bgc + b0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (b0 + u < n_bgc) acc[u] += row[u];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (b0 + u < n_bgc) best[(size_t)a * n_bgc + b0 + u] = acc[u];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && selfsum) {
        long long acc = 0;
        for (int i = d0; i < d1; ++i) acc += selfsc[bgc_doms[i]];
        selfsum[a] = acc;
    }
}

// sim[a][b] = (best[a][b] + best[b][a]) / (selfsum[a] + selfsum[b]).  32x32 tiles; the
// transposed operand goes through shared memory so both reads are coalesced.
__global__ void __launch_bounds__(256)
bgc_sim_kernel(const int64_t *__restrict__ best, const int64_t *__restrict__ selfsum, int32_t n_bgc,
               double *__restrict__ sim)
{
    __shared__ long long tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    // tile[i][j] = best[b0 + i][a0 + j]
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int bi = b0 + r, aj = a0 + tx;
        tile[r][tx] = (bi < n_bgc && aj < n_bgc) ? best[(size_t)bi * n_bgc + aj] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int a = a0 + r, b = b0 + tx;
        if (a < n_bgc && b < n_bgc) {
            const long long sa = selfsum[a], sb = selfsum[b];
            const long long den = sa + sb;
            double v = 0.0;
            if (den > 0) {
                const long long num = best[(size_t)a * n_bgc + b] + tile[tx][r];
                v = __ddiv_rn((double)num, (double)den);   // one correctly-rounded IEEE divide
            }
            if (a == b && sa > 0) v = 1.0;
            sim[(size_t)a * n_bgc + b] = v;
        }
    }
}

cudaError_t launch_bgc_reduce(const int32_t *rowbest, const int32_t *selfsc, const int32_t *bgc_start,
                              const int32_t *bgc_doms, int32_t n_bgc, int64_t *best, int64_t *selfsum,
                              double *sim, cudaStream_t st)
{
    dim3 grid((n_bgc + 1023) / 1024, n_bgc);
    bgc_sum_kernel<<<grid, 256, 0, st>>>(rowbest, selfsc, bgc_start, bgc_doms, n_bgc, best, selfsum);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess || !sim) return err;
    dim3 grid2((n_bgc + 31) / 32, (n_bgc + 31) / 32);
    bgc_sim_kernel<<<grid2, 256, 0, st>>>(best, selfsum, n_bgc, sim);
    return cudaGetLastError();
}

}  // namespace bgca
