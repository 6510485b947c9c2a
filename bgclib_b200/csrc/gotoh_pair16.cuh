// K2/K3 — packed (s16x2 DPX) affine-gap DP kernel, sm_100a.
//
// Layout of the work.  A tile is a *query pair* (qa, qb) against a run of
// subjects.  The two queries live in the two 16-bit halves of every register,
// so each DPX instruction advances two alignments that read the same subject
// residue.  A group of G lanes sweeps one subject systolically: lane t owns
// query rows [t*K, (t+1)*K) (H and E^ kept in registers for both queries), at
// step tau it processes subject column j = tau - t; the bottom H / F^ of a
// lane travel to lane t+1 with __shfl_up_sync.  G = 32 is the warp-cooperative
// ("intra-sequence") form for long domains, G = 8/16 runs 4/2 subjects per
// warp for short ones ("inter-sequence" in the limit).
//
// The substitution scores come from a query-PAIR profile in shared memory:
// prof[b][chunk][lane][4] holds, for subject letter b and query row i, the
// word (sub[qa_i][b] | sub[qb_i][b] << 16) — already in s16x2 form, so the inner
// loop needs no unpack/permute.  The [chunk][lane] interleave makes every
// LDS.128 of a quarter-warp hit 8 distinct 16-byte bank groups whatever letter
// each lane looks up (row stride is a multiple of 128 B).
//
// Per cell pair ("Ho" form: the registers keep Ho = H + open, the profile holds
// s - open, so the diagonal needs no second copy of H):
//   E  = viaddmax(E, ext, Ho_left)       VIADDMNMX.S16x2        alu pipe
//   F  = viaddmax(F, ext, Ho_up)         VIADDMNMX.S16x2        alu pipe
//   t  = Ho_diag + (s - open)            VIADD.16x2             fma pipe (ncu: pipe_fma)
//   H  = max3(t, E, F) [relu]            VIMNMX3.S16x2[.RELU]   alu pipe
//   Ho = H + open                        VIADD.16x2             fma pipe
//   best = max3(best, H_even, H_odd)     VIMNMX3.S16x2          alu pipe, every 2nd cell (local)
// = 5.5 issue slots per 2 cells of which 3.5 on the alu pipe (local); 5 / 3 (global).
#pragma once
#include "common.cuh"

namespace bgca {

constexpr uint32_t NEG16x2 = 0x8AD08AD0u;  // (-30000, -30000): "-inf" that survives +open/+extend

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

template <int G, int K>
struct Pair16Cfg {
    static constexpr int CH = (K + 3) / 4;          // 16-byte chunks per lane per letter
    static constexpr int LPAD = G * K;              // query rows covered in one pass
    static constexpr int ROW_U4 = CH * G;           // uint4 per profile row
    static constexpr int PROF_BYTES = NSYM * ROW_U4 * 16;
    static constexpr int SMEM_BYTES = PROF_BYTES + 2 * LPAD + NSYM * NSYM + 16;
    // registers: 2K state + 4CH profile + ~24 -> pick CTAs/SM so ptxas has room
    static constexpr int MIN_CTAS = (K <= 8) ? 5 : (K <= 16) ? 4 : (K <= 24) ? 3 : 2;
};

// Gap presets: with the gap scores known at compile time the DPX ops take them as
// immediates, which saves one register-file read per op (the register file, not a
// pipe, is what the mixed alu/fma stream saturates — see DESIGN.md "What bounds it").
// PRESET 0 = run-time gap scores (any open <= extend <= 0).
template <int PRESET> struct GapPreset { static constexpr int OPEN = 0, EXT = 0; };
template <> struct GapPreset<1> { static constexpr int OPEN = -12, EXT = -1; };   // Biopython blastp preset
template <> struct GapPreset<2> { static constexpr int OPEN = -11, EXT = -1; };   // BLAST "11/1" read as open+extend
constexpr int NUM_GAP_PRESETS = 3;

__host__ __device__ constexpr uint32_t cpack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

template <int G, int K, int MODE, int PRESET>
__global__ void __launch_bounds__(CTA_THREADS, (Pair16Cfg<G, K>::MIN_CTAS))
gotoh_pair16_kernel(const ScoreParams p)
{
    using Cfg = Pair16Cfg<G, K>;
    constexpr int CH = Cfg::CH, LPAD = Cfg::LPAD;
    constexpr bool LOCAL = (MODE == BGCA_MODE_LOCAL);
    constexpr unsigned FULL = 0xFFFFFFFFu;

    extern __shared__ uint4 smem_u4[];
    uint4 *prof = smem_u4;                                            // [NSYM][CH][G]
    uint8_t *qcodes = reinterpret_cast<uint8_t *>(smem_u4) + Cfg::PROF_BYTES;  // [2][LPAD]
    int8_t *ssub = reinterpret_cast<int8_t *>(qcodes + 2 * LPAD);
    __shared__ int s_tile;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int t = lane % G;                 // lane within the alignment group
    constexpr int NGROUPS = CTA_THREADS / G;

    for (int i = tid; i < NSYM * NSYM; i += CTA_THREADS) ssub[i] = p.sub[i];

    // (e,e) / (o,o): immediates under a preset, pre-packed kernel parameters otherwise
    const uint32_t ext2 = PRESET ? cpack2(GapPreset<PRESET>::EXT) : p.ext2;
    const uint32_t open2 = PRESET ? cpack2(GapPreset<PRESET>::OPEN) : p.open2;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
        __syncthreads();                    // also: every warp is done with the old profile
        const int tile = s_tile;
        if (tile >= p.n_tiles) break;
        const bgca_tile T = p.tiles[tile];
        const int qa = T.qa, qb = T.qb;
        const int La = p.seq_len[qa];
        const int Lb = (qb >= 0) ? p.seq_len[qb] : 0;

        // ---- stage the two queries, then build the pair profile ----------------
        {
            const uint8_t *ca = p.codes + p.seq_off[qa];
            const uint8_t *cb = p.codes + p.seq_off[qb >= 0 ? qb : qa];
            for (int i = tid; i < LPAD; i += CTA_THREADS) {
                qcodes[i] = (i < La) ? ca[i] : (uint8_t)255;
                qcodes[LPAD + i] = (i < Lb) ? cb[i] : (uint8_t)255;
            }
        }
        __syncthreads();
        {
            uint32_t *pw = reinterpret_cast<uint32_t *>(prof);
            for (int idx = tid; idx < NSYM * LPAD; idx += CTA_THREADS) {
                const int b = idx / LPAD, pos = idx - b * LPAD;
                const int ca = qcodes[pos], cb = qcodes[LPAD + pos];
                // profile word = score - open; rows past the query score 0
                const int sa = ((ca < NSYM) ? ssub[ca * NSYM + b] : 0) - p.open;
                const int sb = ((cb < NSYM) ? ssub[cb * NSYM + b] : 0) - p.open;
                const int tt = pos / K, k = pos - tt * K;
                pw[(((b * CH) + (k >> 2)) * G + tt) * 4 + (k & 3)] =
                    ((uint32_t)sa & 0xFFFFu) | ((uint32_t)sb << 16);
            }
        }
        __syncthreads();

        // column-boundary values of this lane's rows (H[i][-1]); local: all 0
        const int i0 = t * K;

        // ---- subjects of this tile, NGROUPS at a time (warp-uniform trip count) --
        for (int sbase = T.s_begin + (tid >> 5) * (32 / G); sbase < T.s_end; sbase += NGROUPS) {
            const int s = sbase + (lane / G);
            const bool valid = s < T.s_end;
            const int Ls = valid ? p.seq_len[s] : 0;
            const uint8_t *sp = p.codes + (valid ? p.seq_off[s] : 0u);
            int Lmax = Ls;
            if (G < 32) {
#pragma unroll
                for (int off = G; off < 32; off <<= 1) Lmax = max(Lmax, __shfl_xor_sync(FULL, Lmax, off));
            }

            uint32_t Ho[K], E[K];     // Ho[k] = H[i0+k][j-1] + open,  E[k] = E[i0+k][j-1]
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Ho[k] = LOCAL ? open2 : pack2(2 * p.open + (i0 + k) * p.extend);
                E[k] = LOCAL ? 0u : NEG16x2;
            }
            uint32_t best = 0;
            uint32_t h_out = 0, f_out = 0;
            // Ho[i0-1][j-1]: column boundary of the row above this lane
            uint32_t hin_prev = (LOCAL || t == 0) ? open2 : pack2(2 * p.open + (i0 - 1) * p.extend);
            uint32_t capA = 0, capB = 0;
            int j = -t;
            uint32_t c_next = ((unsigned)j < (unsigned)Ls) ? (uint32_t)__ldg(sp + j) : 0u;

            const int nsteps = Lmax + G - 1;
            for (int tau = 0; tau < nsteps; ++tau, ++j) {
                uint32_t h_in = __shfl_up_sync(FULL, h_out, 1, G);
                uint32_t f_in = __shfl_up_sync(FULL, f_out, 1, G);
                const uint32_t c = c_next;
                c_next = ((unsigned)(j + 1) < (unsigned)Ls) ? (uint32_t)__ldg(sp + j + 1) : 0u;
                if (t == 0) {                       // top boundary row: Ho[-1][j], F[-1][j]
                    h_in = LOCAL ? open2 : pack2(2 * p.open + j * p.extend);
                    f_in = LOCAL ? 0u : NEG16x2;
                }
                if ((unsigned)j < (unsigned)Ls) {
                    const uint4 *prow = prof + c * (CH * G) + t;
                    uint32_t P[CH * 4];
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        const uint4 v = prow[ch * G];
                        P[ch * 4 + 0] = v.x; P[ch * 4 + 1] = v.y;
                        P[ch * 4 + 2] = v.z; P[ch * 4 + 3] = v.w;
                    }
                    // phase 1 — independent adds (fma pipe): T[k] = H_diag + s, in place in P
                    P[0] = __vadd2(hin_prev, P[0]);
#pragma unroll
                    for (int k = 1; k < K; ++k) P[k] = __vadd2(Ho[k - 1], P[k]);
                    hin_prev = h_in;
                    // phase 2 — the dependent chain down this lane's rows
                    uint32_t up = h_in, F = f_in;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        E[k] = __viaddmax_s16x2(E[k], ext2, Ho[k]);
                        F = __viaddmax_s16x2(F, ext2, up);
                        const uint32_t h = LOCAL ? __vimax3_s16x2_relu(P[k], E[k], F)
                                                 : __vimax3_s16x2(P[k], E[k], F);
                        up = __vadd2(h, open2);
                        Ho[k] = up;
                        if (LOCAL) {
                            // The best cell of a local alignment is always entered by a
                            // substitution (E/F only ever lower an earlier H), so tracking
                            // max(H_diag + s) is exact — and gives the sum a second reader,
                            // which keeps ptxas from folding the add into a VIADDMNMX on the
                            // alu pipe.
                            if (k & 1) best = __vimax3_s16x2(best, P[k - 1], P[k]);
                            else if (k == K - 1) best = __vmaxs2(best, P[k]);
                        }
                    }
                    h_out = up;
                    f_out = F;
                    if (!LOCAL) {
                        if (j == Ls - 1) {          // last column: pick H[L-1][Ls-1] of each half
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                if (i0 + k == La - 1) capA = Ho[k];
                                if (i0 + k == Lb - 1) capB = Ho[k];
                            }
                        }
                    }
                }
            }

            int32_t ra, rb;
            if (LOCAL) {
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1)
                    best = __vmaxs2(best, __shfl_xor_sync(FULL, best, off, G));
                ra = (int32_t)(int16_t)(best & 0xFFFFu);
                rb = (int32_t)(int16_t)(best >> 16);
            } else {
                const uint32_t a = __shfl_sync(FULL, capA, (La - 1) / K, G);
                const uint32_t b = __shfl_sync(FULL, capB, (Lb > 0 ? Lb - 1 : 0) / K, G);
                ra = (int32_t)(int16_t)(a & 0xFFFFu) - p.open;      // captured Ho = H + open
                rb = (int32_t)(int16_t)(b >> 16) - p.open;
            }
            if (t == 0 && valid) {
                emit_pair(p, qa, s, ra);
                emit_pair(p, qb, s, rb);
            }
        }
    }
}

}  // namespace bgca
