// Shared device/host declarations of the BGC aligner engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bgca.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "bgclib_b200 is written for sm_100a (B200) only"
#endif

namespace bgca {

constexpr int NSYM = BGCA_NSYM;
constexpr int CTA_THREADS = 256;  // 8 warps per CTA for every DP kernel

// Everything a DP kernel needs; passed by value (fits the 4 KB param space).
struct ScoreParams {
    const uint8_t *codes;      // packed residues, sorted order, 16 B aligned per sequence
    const uint32_t *seq_off;   // [n]   byte offset of each sorted sequence
    const int32_t *seq_len;    // [n]
    const int32_t *order;      // [n]   sorted -> original index
    const int32_t *type_of;    // [n]   sorted -> type id   (nullable)
    const int32_t *bgc_of;     // [n]   sorted -> BGC index (nullable)
    const bgca_tile *tiles;    // tiles of ONE kernel variant
    int32_t n_tiles;
    int32_t *tile_counter;     // persistent-CTA work queue head
    int32_t *scores;           // [n*n]      original space, nullable
    // [n * s16_stride] SORTED space, nullable: where the s16x2 kernels of a square plan leave their
    // scores (they fit int16 by construction) for permute16_kernel to carry into `scores`.  In
    // sorted space a tile's row segment is contiguous and the mirrored entries of neighbouring
    // query pairs share sectors, so L2 merges the stores and each sector reaches DRAM once; written
    // straight into the caller's index space the same stores cost 6.7x their bytes in DRAM traffic.
    int16_t *scores16;
    int32_t s16_stride;
    int32_t *rowbest;          // [n*n_bgc]  original space, nullable
    int32_t *selfsc;           // [n]        original space, nullable
    int32_t n;
    int32_t n_bgc;
    int32_t open, extend;      // negative scores
    uint32_t open2, ext2;      // the same, replicated in both 16-bit halves
    int32_t type_check;
    // query-vs-database plan (cross_nq > 0): ORIGINAL indices [cross_q0, cross_q0 + cross_nq) are the
    // queries, the rest the database.  cross_q0 = 0 for a batch packed [queries | database] in one
    // call, = n_database when queries were appended behind a persistent database.
    int32_t cross_nq, cross_q0;
    // multi-pass sweep of queries longer than one pass (G*K rows): per CTA and alignment group a
    // line of mp_stride (Ho, F) pairs — the bottom row of a pass, one entry per stream position —
    // parked in L2/HBM for the next pass to take as its top boundary
    int2 *mp_scratch;
    int32_t mp_stride;
    int8_t sub[NSYM * NSYM];
};

// K5 epilogue, fused into every DP kernel: one finished alignment (q, s) in
// sorted space.  Emits only s >= q so that each unordered pair is written once
// even though the packed halves also sweep the mirrored first pair.
__device__ __forceinline__ void emit_pair(const ScoreParams &p, int q, int s, int32_t score)
{
    if (q < 0) return;
    const int oq = p.order[q], os = p.order[s];
    if (p.cross_nq > 0) {
        // query-vs-database: only cross pairs and self pairs count (the planner lists every cross
        // pair once, in either order); scores is [n_query x n_db]
        const bool q_is_query = (unsigned)(oq - p.cross_q0) < (unsigned)p.cross_nq;
        const bool s_is_query = (unsigned)(os - p.cross_q0) < (unsigned)p.cross_nq;
        if (q != s && q_is_query == s_is_query) return;
        if (p.scores && q != s) {
            const int oqq = q_is_query ? oq : os, odb = q_is_query ? os : oq;
            const int row = oqq - p.cross_q0;
            const int col = p.cross_q0 == 0 ? odb - p.cross_nq : odb;      // database index
            p.scores[(size_t)row * (p.n - p.cross_nq) + col] = score;
        }
    } else if (s < q) {
        return;
    } else if (p.scores16) {
        p.scores16[(size_t)q * p.s16_stride + s] = (int16_t)score;
        p.scores16[(size_t)s * p.s16_stride + q] = (int16_t)score;
    } else if (p.scores) {
        p.scores[(size_t)oq * p.n + os] = score;
        p.scores[(size_t)os * p.n + oq] = score;
    }
    if (q == s) {
        if (p.selfsc) p.selfsc[oq] = score;
    }
    if (p.rowbest) {
        if (!p.type_check || p.type_of[q] == p.type_of[s]) {
            atomicMax(&p.rowbest[(size_t)oq * p.n_bgc + p.bgc_of[s]], score);
            if (q != s) atomicMax(&p.rowbest[(size_t)os * p.n_bgc + p.bgc_of[q]], score);
        }
    }
}

}  // namespace bgca
