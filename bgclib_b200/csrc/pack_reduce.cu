// K1 (packer), K4 launcher, K5 (BGC reduction) — the HBM-bound stages.
#include <algorithm>

#include "gotoh_s32.cuh"
#include "launchers.h"

namespace bgca {

// ---------------------------------------------------------------------------
// K1: ASCII -> residue codes, written in SORTED order, each sequence followed by
// at least one BGCA_PAD_CODE byte (the DP kernels read it as the end-of-subject
// token) and padded to a 16-byte multiple.
//
// HBM-bound by design: a lane owns one 16-byte output chunk; it fetches the (at
// most two) 16-byte ALIGNED source words that cover its 16 residues with 128-bit
// loads — the source strings sit at arbitrary byte offsets — realigns them with
// funnel shifts, encodes through a 256-byte shared-memory table (letters A..Z
// fall in 7 distinct banks: no conflicts) and issues one 128-bit store.  A warp
// takes PACK_SEQS consecutive sorted sequences per iteration and issues all
// their loads before using any (the unrolled first 512 bytes of each), with the
// next iteration's metadata prefetched, so the kernel is limited by bytes in
// flight rather than by dependent-load latency.  Only aligned words that hold
// at least one valid residue are touched, so nothing outside the caller's
// buffer is read.
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

constexpr int PACK_SEQS = 4;     // sequences per warp iteration
constexpr int PACK_ROUNDS = 4;   // unrolled rounds of 32 chunks whose loads are all in flight together

// One 16-byte output chunk.  Source side: five 4-byte-aligned words cover the 16 residues
// whatever the byte offset of the string (strings sit at arbitrary offsets); only words that
// hold at least one residue of this sequence are read, so nothing outside the caller's buffer
// is touched.
struct PackChunk {
    uint32_t w[5];
};

// abase = the ASCII buffer aligned down to 4 bytes, mis = its misalignment; so = byte offset of
// the sequence in the buffer.  All index math is 32-bit (a batch is < 4 GiB); the word index is
// clamped to the word that holds the sequence's last residue, so every load is in bounds and
// there is no branch around it.
__device__ __forceinline__ PackChunk pack_fetch(const uint32_t *__restrict__ abase, uint32_t mis, uint32_t so,
                                                int len, int base)
{
    PackChunk c;
    const uint32_t w = (so + mis + (uint32_t)base) >> 2;
    const uint32_t last = (so + mis + (uint32_t)len - 1u) >> 2;
#pragma unroll
    for (int i = 0; i < 5; ++i) c.w[i] = __ldg(abase + min(w + (uint32_t)i, last));
    return c;
}

// realign (one funnel shift per word), encode through the shared table, blank what lies past
// the sequence with the pad code, one 128-bit store.  Returns the OR of the encoded words: a
// set bit 7 in any byte = an illegal letter somewhere in this chunk.
__device__ __forceinline__ uint32_t pack_emit(const PackChunk &c, uint32_t mis, uint32_t so, int len, int base,
                                              uint8_t *dst, const uint8_t *lut)
{
    const uint32_t bits = ((so + mis + (uint32_t)base) & 3u) * 8u;
    uint32_t out[4], bad = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t raw = __funnelshift_r(c.w[i], c.w[i + 1], bits);
        const uint32_t c0 = lut[raw & 0xFFu], c1 = lut[(raw >> 8) & 0xFFu];
        const uint32_t c2 = lut[(raw >> 16) & 0xFFu], c3 = lut[raw >> 24];
        uint32_t o = __byte_perm(__byte_perm(c0, c1, 0x0040), __byte_perm(c2, c3, 0x0040), 0x5410);
        // residues left in this word (0..4): keep that many low bytes, pad code in the others
        const int nb = min(max(len - (base + 4 * i), 0), 4);
        const uint32_t keep = __funnelshift_lc(0xFFFFFFFFu, 0u, 8 * nb);
        o = (o & keep) | ((BGCA_PAD_CODE * 0x01010101u) & ~keep);
        bad |= o;
        out[i] = o;
    }
    *reinterpret_cast<uint4 *>(dst + base) = make_uint4(out[0], out[1], out[2], out[3]);
    return bad & 0x80808080u;
}

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ ascii, const int64_t *__restrict__ src_off_sorted,
            const int32_t *__restrict__ order, const uint32_t *__restrict__ seq_off,
            const int32_t *__restrict__ seq_len, int32_t n, uint8_t *__restrict__ codes,
            unsigned long long *err_word)
{
    constexpr unsigned FULL = 0xFFFFFFFFu;
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)encode_residue(threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(ascii) & 3);
    const uint32_t *abase = reinterpret_cast<const uint32_t *>(ascii - mis);

    // lanes 0..PACK_SEQS-1 hold the metadata of the block's sequences
    auto fetch_meta = [&](int r0, uint32_t &so, int &len, uint32_t &doff) {
        const int r = r0 + lane;
        const bool v = lane < PACK_SEQS && r < n;
        so = v ? (uint32_t)src_off_sorted[r] : 0u;
        len = v ? seq_len[r] : 0;
        doff = v ? seq_off[r] : 0u;
    };
    uint32_t so_n = 0;
    int len_n = 0;
    uint32_t doff_n = 0;
    int r0 = gw * PACK_SEQS;
    if (r0 < n) fetch_meta(r0, so_n, len_n, doff_n);
    for (; r0 < n; r0 += warps_per_grid * PACK_SEQS) {
        const uint32_t so = so_n;
        const int len = len_n;
        const uint32_t doff = doff_n;
        if (r0 + warps_per_grid * PACK_SEQS < n) fetch_meta(r0 + warps_per_grid * PACK_SEQS, so_n, len_n, doff_n);

        // the block's chunks as one flat index space [0, total): sequence u owns [first[u], first[u+1]).
        // The metadata stays in lanes 0..PACK_SEQS-1 and is fetched with lane-indexed shuffles.
        int first[PACK_SEQS + 1];
        first[0] = 0;
#pragma unroll
        for (int u = 0; u < PACK_SEQS; ++u) {
            const int Lu = __shfl_sync(FULL, len, u);     // 0 past the end of the batch
            first[u + 1] = first[u] + (Lu > 0 ? (Lu + 16) >> 4 : 0);
        }
        const int total = first[PACK_SEQS];
        auto locate = [&](int f, uint32_t &s_, uint8_t *&d_, int &l_, int &base_, int &u_) {
            u_ = 0;
            int f0 = 0;
#pragma unroll
            for (int u = 1; u < PACK_SEQS; ++u) {
                const int in = (f >= first[u]);
                u_ += in;
                f0 += in * (first[u] - first[u - 1]);
            }
            s_ = __shfl_sync(FULL, so, u_);
            d_ = codes + __shfl_sync(FULL, doff, u_);
            l_ = __shfl_sync(FULL, len, u_);
            base_ = (f - f0) * 16;
        };
        for (int f_base = 0; f_base < total; f_base += 32 * PACK_ROUNDS) {
            PackChunk ch[PACK_ROUNDS];
            uint32_t s_[PACK_ROUNDS];
            uint8_t *d_[PACK_ROUNDS];
            int l_[PACK_ROUNDS], b_[PACK_ROUNDS], u_[PACK_ROUNDS];
            bool have[PACK_ROUNDS];
#pragma unroll
            for (int k = 0; k < PACK_ROUNDS; ++k) {
                const int f = f_base + 32 * k + lane;
                have[k] = f < total;                              // a lane without a chunk in round k
                locate(min(f, total - 1), s_[k], d_[k], l_[k], b_[k], u_[k]);   // re-reads the last one
                ch[k] = pack_fetch(abase, mis, s_[k], l_[k], b_[k]);
            }
#pragma unroll
            for (int k = 0; k < PACK_ROUNDS; ++k) {
                if (!have[k]) continue;
                if (pack_emit(ch[k], mis, s_[k], l_[k], b_[k], d_[k], lut)) {
                    // cold path: name the first illegal letter of this chunk, put 'X' in its place
#pragma unroll 1
                    for (int i = 0; i < 16 && b_[k] + i < l_[k]; ++i)
                        if (lut[ascii[s_[k] + b_[k] + i]] == 255u) {
                            atomicMin(err_word, ((unsigned long long)(uint32_t)order[r0 + u_[k]] << 32) |
                                                    (unsigned long long)(uint32_t)(b_[k] + i));
                            d_[k][b_[k] + i] = 22;
                        }
                }
            }
        }
    }
}

cudaError_t launch_pack(const uint8_t *ascii, const int64_t *src_off_sorted, const int32_t *order,
                        const uint32_t *seq_off, const int32_t *seq_len, int32_t n, uint8_t *codes,
                        unsigned long long *err_word, cudaStream_t st)
{
    // persistent-ish grid: a multiple of the SM count, never more warps than sequence blocks
    const int blocks_needed = (n + PACK_SEQS * 8 - 1) / (PACK_SEQS * 8);
    int grid = 148 * 4;   // 64 registers -> 4 resident CTAs per SM; every warp loops over its blocks
    if (grid > blocks_needed) grid = blocks_needed;
    if (grid < 1) grid = 1;
    pack_kernel<<<grid, 256, 0, st>>>(ascii, src_off_sorted, order, seq_off, seq_len, n, codes, err_word);
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
// Score matrix, sorted space (int16) -> caller's index space (int32):
//     out[i][j] = S[inv[i]][inv[j]]       for every entry the s16x2 kernels wrote.
// A CTA takes output rows one at a time: the sorted row inv[i] (n int16, 16-byte aligned rows)
// goes to shared memory with 128-bit loads, every thread gathers T[inv[j]] for its columns and
// the row leaves as coalesced 128-byte lines.  HBM-bound: 2*n*n read + 4*n*n written (+ the
// inv table, L2-resident).  Entries that still hold the fill value were not computed by these
// kernels (another rank's tiles, 32-bit tiles, masked type pairs) and are left as the caller
// initialised them.
// ---------------------------------------------------------------------------
constexpr int PERM_THREADS = 512;
__global__ void __launch_bounds__(PERM_THREADS)
permute16_kernel(const int16_t *__restrict__ S, int stride, const int32_t *__restrict__ inv, int n,
                 int32_t *__restrict__ out, int skip_unwritten)
{
    extern __shared__ uint4 perm_smem[];
    int16_t *T = reinterpret_cast<int16_t *>(perm_smem);
    const int nvec = (n + 7) >> 3;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(S + (size_t)inv[i] * stride);
        __syncthreads();                    // the previous row has been read out
        for (int v = threadIdx.x; v < nvec; v += PERM_THREADS) perm_smem[v] = __ldg(src + v);
        __syncthreads();
        int32_t *dst = out + (size_t)i * n;
        for (int j = threadIdx.x; j < n; j += PERM_THREADS) {
            const int16_t v = T[__ldg(inv + j)];
            if (!skip_unwritten || v != (int16_t)BGCA_S16_UNWRITTEN) dst[j] = (int32_t)v;
        }
    }
}

cudaError_t launch_permute16(const int16_t *S, int stride, const int32_t *inv, int n, int32_t *out,
                             int skip_unwritten, int n_sm, cudaStream_t st)
{
    const size_t smem = (size_t)((n + 7) >> 3) * 16;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t err = cudaFuncSetAttribute(permute16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        smem_set = smem;
    }
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
    int grid = std::min(n, n_sm * per_sm);
    permute16_kernel<<<grid, PERM_THREADS, smem, st>>>(S, stride, inv, n, out, skip_unwritten);
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
