// K7 — alignment statistics without traceback, sm_100a.
//
// For an explicit list of (query, subject) pairs: score, identities, alignment length and
// the start/end coordinates of the alignment that is the lexicographic maximum of
//   (score, identities, -columns, q_start, s_start, -s_end, -q_end)
// over ALL alignments of the affine-gap model (definition: DESIGN.md "K7"; CPU statement:
// oracle/gotoh_oracle.c bgco_stats_pair).  The value (score, identities, -columns) is
// additive along a path and the start is fixed at the path's origin, so the Gotoh
// recurrence over tuples computes the maximum exactly — no traceback matrix, and the result
// does not depend on evaluation order, which is what lets the systolic sweep below be
// bit-exact against the row-major CPU loop.
//
// Tuple = 64-bit key (score << 32 | identities << 16 | 65535 - columns) + 32-bit payload
// (q_start << 16 | s_start); "one more column" is a single 64-bit add.  Same warp-per-pair
// systolic sweep as K4 (lane t owns KS query rows, bottom row of a 32*KS-row pass parked in
// an L2-resident scratch line).  96-bit state per DP value, no DPX: this is the detail path
// for selected pairs, not the all-vs-all throughput path.
#pragma once
#include "common.cuh"

namespace bgca {

constexpr int KS = 8;
constexpr int PASS_ST = 32 * KS;

// Wide tuple: any length up to 32767.
struct Tup {
    long long k;
    uint32_t p;
    static __device__ __forceinline__ Tup zero(int i, int j) { return Tup{65535LL, ((uint32_t)i << 16) | (uint32_t)j}; }
    static __device__ __forceinline__ Tup neg() { return Tup{(long long)(-(1 << 28)) * 4294967296LL + 65535, 0u}; }
    static __device__ __forceinline__ long long delta(int score, int ident)
    {
        return (long long)score * 4294967296LL + (long long)(ident << 16) - 1;
    }
    // a gap of `len` >= 1 residues from the corner (global boundary)
    static __device__ __forceinline__ Tup edge(int open, int ext, int len)
    {
        return Tup{(long long)(open + (len - 1) * ext) * 4294967296LL + (65535 - len), 0u};
    }
    __device__ __forceinline__ bool gt(const Tup &o) const { return k > o.k || (k == o.k && p > o.p); }
    __device__ __forceinline__ bool same(const Tup &o) const { return k == o.k && p == o.p; }
    __device__ __forceinline__ Tup shfl_up() const
    {
        return Tup{__shfl_up_sync(0xFFFFFFFFu, k, 1), __shfl_up_sync(0xFFFFFFFFu, p, 1)};
    }
    __device__ __forceinline__ Tup shfl_xor(int off) const
    {
        return Tup{__shfl_xor_sync(0xFFFFFFFFu, k, off), __shfl_xor_sync(0xFFFFFFFFu, p, off)};
    }
    __device__ __forceinline__ Tup shfl(int src) const
    {
        return Tup{__shfl_sync(0xFFFFFFFFu, k, src), __shfl_sync(0xFFFFFFFFu, p, src)};
    }
    __device__ __forceinline__ int score() const { return (int)(k >> 32); }
    __device__ __forceinline__ int identities() const { return (int)((k >> 16) & 0xFFFF); }
    __device__ __forceinline__ int columns() const { return 65535 - (int)(k & 0xFFFF); }
    __device__ __forceinline__ int q_start() const { return (int)(p >> 16); }
    __device__ __forceinline__ int s_start() const { return (int)(p & 0xFFFF); }
    __device__ __forceinline__ uint4 park_lo(const Tup &f) const
    {
        return make_uint4((uint32_t)k, (uint32_t)((unsigned long long)k >> 32), p, (uint32_t)f.k);
    }
    __device__ __forceinline__ uint4 park_hi(const Tup &f) const
    {
        return make_uint4((uint32_t)((unsigned long long)f.k >> 32), f.p, 0u, 0u);
    }
    static __device__ __forceinline__ void unpark(const uint4 &b0, const uint4 &b1, Tup &h, Tup &f)
    {
        h = Tup{(long long)(((unsigned long long)b0.y << 32) | b0.x), b0.z};
        f = Tup{(long long)(((unsigned long long)b1.x << 32) | b0.w), b1.y};
    }
};

// Compact tuple for sequences of at most 1023 residues (every protein domain): the same
// lexicographic order (score, identities, -columns, q_start, s_start) in ONE 64-bit integer —
// score:23 | identities:10 | 2047-columns:11 | q_start:10 | s_start:10 — so a max is one 64-bit
// compare and an add touches one value.
struct Tup64 {
    long long k;
    static __device__ __forceinline__ Tup64 zero(int i, int j) { return Tup64{(2047LL << 20) | ((long long)i << 10) | j}; }
    static __device__ __forceinline__ Tup64 neg() { return Tup64{((long long)(-(1 << 20)) << 41) + (2047LL << 20)}; }
    static __device__ __forceinline__ long long delta(int score, int ident)
    {
        return ((long long)score << 41) + ((long long)ident << 31) - (1LL << 20);
    }
    static __device__ __forceinline__ Tup64 edge(int open, int ext, int len)
    {
        return Tup64{((long long)(open + (len - 1) * ext) << 41) + ((long long)(2047 - len) << 20)};
    }
    __device__ __forceinline__ bool gt(const Tup64 &o) const { return k > o.k; }
    __device__ __forceinline__ bool same(const Tup64 &o) const { return k == o.k; }
    __device__ __forceinline__ Tup64 shfl_up() const { return Tup64{__shfl_up_sync(0xFFFFFFFFu, k, 1)}; }
    __device__ __forceinline__ Tup64 shfl_xor(int off) const { return Tup64{__shfl_xor_sync(0xFFFFFFFFu, k, off)}; }
    __device__ __forceinline__ Tup64 shfl(int src) const { return Tup64{__shfl_sync(0xFFFFFFFFu, k, src)}; }
    __device__ __forceinline__ int score() const { return (int)(k >> 41); }
    __device__ __forceinline__ int identities() const { return (int)((k >> 31) & 1023); }
    __device__ __forceinline__ int columns() const { return 2047 - (int)((k >> 20) & 2047); }
    __device__ __forceinline__ int q_start() const { return (int)((k >> 10) & 1023); }
    __device__ __forceinline__ int s_start() const { return (int)(k & 1023); }
    __device__ __forceinline__ uint4 park_lo(const Tup64 &f) const
    {
        return make_uint4((uint32_t)k, (uint32_t)((unsigned long long)k >> 32), (uint32_t)f.k,
                          (uint32_t)((unsigned long long)f.k >> 32));
    }
    __device__ __forceinline__ uint4 park_hi(const Tup64 &) const { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ void unpark(const uint4 &b0, const uint4 &, Tup64 &h, Tup64 &f)
    {
        h = Tup64{(long long)(((unsigned long long)b0.y << 32) | b0.x)};
        f = Tup64{(long long)(((unsigned long long)b0.w << 32) | b0.z)};
    }
};

template <typename T> __device__ __forceinline__ T tup_max(const T &a, const T &b) { return b.gt(a) ? b : a; }
template <typename T> __device__ __forceinline__ T tup_add(T a, long long d) { a.k += d; return a; }

template <int MODE, typename T>
__global__ void __launch_bounds__(CTA_THREADS)
gotoh_stats_kernel(const ScoreParams p, const int32_t *__restrict__ pq, const int32_t *__restrict__ ps,
                   long long npairs, uint4 *scratch, int scratch_stride, bgca_alnstats *__restrict__ out)
{
    constexpr bool LOCAL = (MODE == BGCA_MODE_LOCAL);
    constexpr unsigned FULL = 0xFFFFFFFFu;
    __shared__ int8_t ssub[NSYM * NSYM];
    for (int i = threadIdx.x; i < NSYM * NSYM; i += CTA_THREADS) ssub[i] = p.sub[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (CTA_THREADS / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (CTA_THREADS / 32);
    uint4 *bnd = scratch + w * 2 * (long long)scratch_stride;   // two uint4 per subject column
    const long long d_open = T::delta(p.open, 0), d_ext = T::delta(p.extend, 0);
    const T NEG = T::neg();
    // H[r+1][0] and H[0][j+1] of the global boundary: a gap of r+1 (j+1) columns from the corner
    auto edge = [&](int len) { return T::edge(p.open, p.extend, len); };

    for (long long pair = w; pair < npairs; pair += nw) {
        const int q = pq[pair], s = ps[pair];
        const int La = p.seq_len[q], Ls = p.seq_len[s];
        const uint8_t *qp = p.codes + p.seq_off[q];
        const uint8_t *sp = p.codes + p.seq_off[s];
        T best = T::zero(0, 0);             // best of this lane over all passes
        uint32_t best_end = 0;              // s_end << 16 | q_end (DP corner), smaller wins ties
        T cap = T::zero(0, 0);
        const int npass = (La + PASS_ST - 1) / PASS_ST;

        for (int pass = 0; pass < npass; ++pass) {
            const int i0 = pass * PASS_ST + lane * KS;
            int qrow[KS];
            T Hp[KS], Eh[KS];
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const int r = i0 + k;
                qrow[k] = (r < La) ? (int)qp[r] * NSYM : -NSYM;
                Hp[k] = LOCAL ? T::zero(r + 1, 0) : edge(r + 1);
                Eh[k] = NEG;
            }
            T hin_prev = LOCAL ? T::zero(i0, 0) : (i0 == 0 ? T::zero(0, 0) : edge(i0));
            T h_out = NEG, f_out = NEG;
            T pbest = T::zero(0, 0);        // best of this pass (visited column-major: strict > keeps
            uint32_t pbest_end = 0;         // the smallest s_end, then the smallest q_end)
            int j = -lane;
            int c_next = ((unsigned)j < (unsigned)Ls) ? (int)__ldg(sp + j) : 0;
            uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
            if (lane == 0 && pass > 0) { b0 = __ldcg(bnd); b1 = __ldcg(bnd + 1); }
            const bool park = (pass + 1 < npass);

            for (int tau = 0; tau < Ls + 31; ++tau, ++j) {
                T h_in = h_out.shfl_up();
                T f_in = f_out.shfl_up();
                const int c = c_next;
                c_next = ((unsigned)(j + 1) < (unsigned)Ls) ? (int)__ldg(sp + j + 1) : 0;
                if (lane == 0) {
                    if (pass == 0) {
                        h_in = LOCAL ? T::zero(0, j + 1) : edge(j + 1);
                        f_in = NEG;
                    } else {
                        T::unpark(b0, b1, h_in, f_in);
                        if (j + 1 < Ls) { b0 = __ldcg(bnd + 2 * (j + 1)); b1 = __ldcg(bnd + 2 * (j + 1) + 1); }
                    }
                }
                if ((unsigned)j < (unsigned)Ls) {
                    T diag = hin_prev;
                    hin_prev = h_in;
                    T up = h_in, F = f_in;
                    const int crow = c * NSYM;
#pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        const bool real = qrow[k] >= 0;
                        const int sc = real ? (int)ssub[qrow[k] + c] : 0;
                        Eh[k] = tup_max(tup_add(Eh[k], d_ext), tup_add(Hp[k], d_open));
                        F = tup_max(tup_add(F, d_ext), tup_add(up, d_open));
                        T h = tup_add(diag, T::delta(sc, qrow[k] == crow));
                        h = tup_max(tup_max(h, Eh[k]), F);
                        if (LOCAL) {
                            h = tup_max(h, T::zero(i0 + k + 1, j + 1));
                            if (real && h.gt(pbest)) {
                                pbest = h;
                                pbest_end = ((uint32_t)(j + 1) << 16) | (uint32_t)(i0 + k + 1);
                            }
                        }
                        diag = Hp[k];
                        Hp[k] = h;
                        up = h;
                    }
                    h_out = up;
                    f_out = F;
                    if (park && lane == 31) {
                        __stcg(bnd + 2 * j, up.park_lo(F));
                        __stcg(bnd + 2 * j + 1, up.park_hi(F));
                    }
                    if (!LOCAL && j == Ls - 1) {
#pragma unroll
                        for (int k = 0; k < KS; ++k)
                            if (i0 + k == La - 1) cap = Hp[k];
                    }
                }
            }
            if (LOCAL) {
                // passes are visited in row order: an equal (key, start) of a later pass wins only
                // with a smaller s_end
                if (pbest.gt(best) || (pbest.same(best) && pbest_end < best_end)) {
                    best = pbest;
                    best_end = pbest_end;
                }
            }
            __syncwarp();   // parked row of this pass is visible to lane 0 of the next
        }

        if (LOCAL) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T o = best.shfl_xor(off);
                const uint32_t oe = __shfl_xor_sync(FULL, best_end, off);
                if (o.gt(best) || (o.same(best) && oe < best_end)) {
                    best = o;
                    best_end = oe;
                }
            }
        } else {
            const int src = ((La - 1) % PASS_ST) / KS;
            best = cap.shfl(src);
            best_end = ((uint32_t)Ls << 16) | (uint32_t)La;
        }
        if (lane == 0) {
            bgca_alnstats r;
            r.score = best.score();
            r.identities = best.identities();
            r.columns = best.columns();
            if (r.columns == 0) {
                r.q_start = r.q_end = r.s_start = r.s_end = -1;
                r.gaps = 0;
            } else {
                r.q_start = best.q_start();
                r.s_start = best.s_start();
                r.q_end = (int32_t)(best_end & 0xFFFF);
                r.s_end = (int32_t)(best_end >> 16);
                r.gaps = 2 * r.columns - (r.q_end - r.q_start) - (r.s_end - r.s_start);
            }
            out[pair] = r;
        }
    }
}

}  // namespace bgca
