// K2/K3 — packed (s16x2 DPX) affine-gap DP kernel, sm_100a.
//
// Layout of the work.  A tile is a *query pair* (qa, qb) against a run of
// subjects.  The two queries live in the two 16-bit halves of every register,
// so each DPX instruction advances two alignments that read the same subject
// residue.  A group of G lanes sweeps subjects systolically: lane t owns query
// rows [t*K, (t+1)*K) (Ho = H + open and E kept in registers for both queries),
// at every step it consumes one token of the group's subject stream, lagging
// lane t-1 by one step; the bottom Ho / F of a lane travel to lane t+1 with
// __shfl_up_sync.  G = 32 is the warp-cooperative ("intra-sequence") form for
// long domains, G = 8/16 runs 4/2 streams per warp for short ones
// ("inter-sequence" in the limit).
//
// Streaming.  A group does not drain its pipeline between subjects: its
// subjects (s_begin + gid, + NGROUPS, ...) are consumed back to back as ONE token
// stream — residues (< 24), then the pad byte the packer always leaves behind a
// sequence (24 = "boundary"), then the next subject.  At a boundary token a lane
// posts its best, the last lane collects and emits the two scores, state is
// reset; lanes that have not started / have finished see IDLE tokens (25).
// Boundary and idle handling sit on a cold path, so the hot loop carries no
// per-step bounds checks and the G-1 step wind-up is paid once per tile, not
// once per subject.
//
// The substitution scores come from a query-PAIR profile in shared memory:
// prof[b][chunk][lane][4] holds, for subject letter b and query row i, the
// word ((sub[qa_i][b]-open) | (sub[qb_i][b]-open) << 16) — already in s16x2 form,
// so the inner loop needs no unpack/permute.  The [chunk][lane] interleave makes
// every LDS.128 of a quarter-warp hit 8 distinct 16-byte bank groups whatever
// letter each lane looks up (row stride is a multiple of 128 B).
//
// Multi-pass (MP = 1, G = 32).  A query pair longer than one pass (32*K rows) is swept in
// ceil(L / (32*K)) passes over the same subject stream: the last lane parks the (Ho, F) it hands
// on — the bottom row of the pass, one pair per stream position — in an L2-resident scratch
// line, and lane 0 of the next pass takes its top boundary from there (prefetched two steps
// ahead like the tokens) instead of the constant boundary row.  Local best scores are carried
// from pass to pass per subject; global scores are emitted by the pass that owns the last row.
// Any query length the packer accepts runs this way at the occupancy of the K <= 32 variants
// (two CTAs per SM), so the 32-bit kernel only serves pairs whose scores leave int16.
//
// Per cell pair ("next-T" form).  The per-row state is not H but what the NEXT
// column needs from it: Tn[k] = Ho[k-1] + (s_next - open) (the diagonal term,
// built with the profile row of the NEXT token, which is prefetched two steps
// ahead) and E[k] already advanced to the next column.  Every state update is
// then in place — no register copies, no saved diagonal:
//   F   = viaddmax(F, ext, Ho_up)        VIADDMNMX.S16x2        alu pipe
//   H   = max3(Tn[k], E[k], F) [relu]    VIMNMX3.S16x2[.RELU]   alu pipe
//   Tn[k] = Ho_up + Pnext[k]             VIADD.16x2             fma pipe (ncu: pipe_fma)
//   Ho  = H + open                       VIADD.16x2             fma pipe
//   E[k] = viaddmax(E[k], ext, Ho)       VIADDMNMX.S16x2        alu pipe
//   best = max3(best, H_even, H_odd)     VIMNMX3.S16x2          alu pipe, every 2nd cell (local)
// = 5.5 issue slots per 2 cells of which 3.5 on the alu pipe (local); 5 / 3 (global).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace bgca {

constexpr uint32_t NEG16x2 = 0x8AD08AD0u;  // (-30000, -30000): "-inf" that survives +open/+extend
constexpr uint32_t TOK_BOUNDARY = BGCA_PAD_CODE;   // 24: the pad byte behind every packed sequence
constexpr uint32_t TOK_IDLE = 25;                  // fills the 64-byte header of the packed batch

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
__host__ __device__ constexpr uint32_t cpack2(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }

// The arithmetic of a DP word.  W32 = 0: two alignments per word in s16x2 (the throughput form).
// W32 = 1: ONE alignment per word in s32 — the same kernel for pairs whose scores can leave int16
// (whole megasynthase proteins against each other, W-rich sequences, custom matrices); half the
// cells per instruction, any score range.
template <int W32> struct DpOps;
template <> struct DpOps<0> {
    static constexpr uint32_t NEG = 0x8AD08AD0u;   // NEG16x2
    static __device__ __forceinline__ uint32_t pack(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
    static __host__ __device__ constexpr uint32_t cpack(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
    static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
    static __device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t max3_relu(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2_relu(a, b, c); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
    static __device__ __forceinline__ int hi(uint32_t v) { return (int)(int16_t)(v >> 16); }
};
template <> struct DpOps<1> {
    static constexpr uint32_t NEG = (uint32_t)(-(1 << 28));
    static __device__ __forceinline__ uint32_t pack(int v) { return (uint32_t)v; }
    static __host__ __device__ constexpr uint32_t cpack(int v) { return (uint32_t)v; }
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return a + b; }
    static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return a - b; }
    static __device__ __forceinline__ uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t max3_relu(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)__vimax3_s32_relu((int)a, (int)b, (int)c); }
    static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return (uint32_t)max((int)a, (int)b); }
    static __device__ __forceinline__ int lo(uint32_t v) { return (int)v; }
    static __device__ __forceinline__ int hi(uint32_t) { return 0; }
};

constexpr int PROF_ROWS = NSYM + 2;   // 24 residues + the two token rows (never scored, must be loadable)

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// one token (byte) of the packed batch, zero-extended by the load itself
__device__ __forceinline__ uint32_t ldg_token(const uint8_t *p)
{
    uint32_t v;
    asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// The packed batch never crosses a 4 GiB boundary (DevBuf::alloc / window4g_start in bgca_api.cu
// see to it), so the token pointer of the hot loop advances with ONE 32-bit add on its low half instead
// of a 64-bit add (IADD3 + IADD3.X) on top of a base pointer reloaded from the constant bank.
__device__ __forceinline__ void advance_lo32(uint64_t &ptr)
{
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tadd.u32 lo, lo, 1;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+l"(ptr));
}

template <int G, int K>
struct Pair16Cfg {
    static constexpr int CH = (K + 3) / 4;          // 16-byte chunks per lane per letter
    static constexpr int LPAD = G * K;              // query rows covered in one pass
    // A quarter-warp (8 lanes) is served by one LDS.128 wavefront when its lanes hit 8 distinct
    // 16-byte bank groups.  With G >= 8 the 8 lanes are 8 different t of one group: 8 distinct
    // columns of the same row.  With G < 8 they belong to 8/G groups that look up DIFFERENT
    // letters (rows) at the same t, so the profile is replicated 8/G times and every lane of a
    // quarter-warp owns its own column (lane & 7) whatever row it reads.
    static constexpr int REP = (G < 8) ? 8 / G : 1;
    static constexpr int W = G * REP;               // 16-byte columns per chunk row
    static constexpr int ROW_BYTES = CH * W * 16;   // one profile row (multiple of 128 B)
    static constexpr int NGROUPS = CTA_THREADS / G; // alignment groups (subject streams) per CTA
    static constexpr int PROF_BYTES = PROF_ROWS * ROW_BYTES;
    // per group a ring of (lo, hi) best slots: lane 0 can run at most G-1 subjects ahead of the
    // last lane (one-residue subjects), so G slots never collide
    static constexpr int RING = G;
    static constexpr int BEST_BYTES = NGROUPS * RING * 2 * 4;
    // Groups that share a warp keep their subject streams in step (G <= 4: 8+ streams per warp,
    // short subjects): after a subject a group idles until the longest subject of the warp's
    // batch has ended, so the cold boundary path runs once per batch and warp instead of once per
    // group — unsynchronised, a 20-residue batch spends more issue slots in boundary code than in
    // DP cells.
    static constexpr bool ALIGN_STREAMS = (G <= 4);
    static constexpr int SMEM_BYTES = PROF_BYTES + BEST_BYTES + 2 * LPAD + NSYM * NSYM + 16;
    // multi-pass: best-so-far of every subject of the tile, carried between passes
    static constexpr int MP_MAX_SUBJECTS = 64;                 // per group and tile (planner: spg <= 64)
    static constexpr int MP_BEST_BYTES = NGROUPS * MP_MAX_SUBJECTS * 2 * 4;
    static constexpr int SMEM_BYTES_MP = SMEM_BYTES + MP_BEST_BYTES;
    // registers: 2K state + 4CH profile + ~24 -> pick CTAs/SM so ptxas has room.
    static constexpr int MIN_CTAS_LOCAL = (K <= 6) ? 5 : (K <= 14) ? 4 : (K <= 18) ? 3 : (K <= 32) ? 2 : 1;
    static constexpr int MIN_CTAS_GLOBAL = (K <= 4) ? 5 : (K <= 9) ? 4 : (K <= 14) ? 3 : (K <= 32) ? 2 : 1;   // a few more live values (boundary walk)
};

// Gap presets: with the gap scores known at compile time the DPX ops take them as
// immediates, which saves one register-file read per op (the register file, not a
// pipe, is what the mixed alu/fma stream saturates — see DESIGN.md "What bounds it").
// PRESET 0 = run-time gap scores (any open <= extend <= 0).
template <int PRESET> struct GapPreset { static constexpr int OPEN = 0, EXT = 0; };
template <> struct GapPreset<1> { static constexpr int OPEN = -12, EXT = -1; };   // Biopython blastp preset
template <> struct GapPreset<2> { static constexpr int OPEN = -11, EXT = -1; };   // BLAST "11/1" read as open+extend
constexpr int NUM_GAP_PRESETS = 3;

template <int G, int K, int MODE, int PRESET, int MP = 0, int W32 = 0>
__global__ void __launch_bounds__(CTA_THREADS, (MP ? (K <= 18 && !W32 ? 3 : 2) : MODE == BGCA_MODE_LOCAL ? Pair16Cfg<G, K>::MIN_CTAS_LOCAL
                                                                              : Pair16Cfg<G, K>::MIN_CTAS_GLOBAL))
gotoh_pair16_kernel(const ScoreParams p)
{
    static_assert(!MP || G == 32, "the multi-pass sweep is built for whole-warp groups");
    static_assert(!W32 || MP, "the 32-bit form is instantiated on the multi-pass kernel only");
    using Ops = DpOps<W32>;
    using Cfg = Pair16Cfg<G, K>;
    constexpr int CH = Cfg::CH, LPAD = Cfg::LPAD, NGROUPS = Cfg::NGROUPS;
    constexpr bool LOCAL = (MODE == BGCA_MODE_LOCAL);
    constexpr unsigned FULL = 0xFFFFFFFFu;

    extern __shared__ uint4 smem_u4[];
    uint4 *prof = smem_u4;                                                    // [PROF_ROWS][CH][W]
    int *sbest = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(smem_u4) + Cfg::PROF_BYTES);
    uint8_t *qcodes = reinterpret_cast<uint8_t *>(sbest) + Cfg::BEST_BYTES;   // [2][LPAD]
    int8_t *ssub = reinterpret_cast<int8_t *>(qcodes + 2 * LPAD);
    // [NGROUPS][MP_MAX_SUBJECTS][2], touched only by the last lane of each group
    int *pbest = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(smem_u4) + ((Cfg::SMEM_BYTES + 15) & ~15));
    __shared__ int s_tile;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int t = lane % G;                 // lane within the alignment group
    const int gid = tid / G;                // group (= subject stream) within the CTA
    const int i0 = t * K;                   // first query row of this lane

    for (int i = tid; i < NSYM * NSYM; i += CTA_THREADS) ssub[i] = p.sub[i];
    for (int i = tid; i < Cfg::BEST_BYTES / 4; i += CTA_THREADS) sbest[i] = 0;

    // (e,e) / (o,o): immediates under a preset, pre-packed kernel parameters otherwise
    const uint32_t ext2 = PRESET ? Ops::cpack(GapPreset<PRESET>::EXT) : (W32 ? (uint32_t)p.extend : p.ext2);
    const uint32_t open2 = PRESET ? Ops::cpack(GapPreset<PRESET>::OPEN) : (W32 ? (uint32_t)p.open : p.open2);
    int *const my_best = sbest + gid * (Cfg::RING * 2);
    // this lane's 16-byte column of the profile, as a 32-bit shared-window address
    uint32_t prof_lane_v = (uint32_t)__cvta_generic_to_shared(prof) + (uint32_t)(G < 8 ? (lane & 7) : t) * 16u;
    // opaque: ptxas otherwise rebuilds this address (S2UR, ULEA, shift, mask, add: 7 instructions) in
    // every step of the hot loop instead of keeping it in a register
    asm volatile("" : "+r"(prof_lane_v));
    const uint32_t prof_lane = prof_lane_v;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
        __syncthreads();                    // also: every warp is done with the old profile
        const int tile = s_tile;
        if (tile >= p.n_tiles) break;
        const bgca_tile T = p.tiles[tile];
        const int qa = T.qa, qb = T.qb;
        const int La = p.seq_len[qa];
        const int Lb = (qb >= 0) ? p.seq_len[qb] : 0;
        const int npass = MP ? (max(La, Lb) + LPAD - 1) / LPAD : 1;
        // this group's line of parked (Ho, F) pairs, indexed by stream position
        int2 *const mp_line = MP ? p.mp_scratch + ((size_t)blockIdx.x * NGROUPS + gid) * (size_t)p.mp_stride : nullptr;

      for (int pass = 0; pass < npass; ++pass) {
        const int row0 = pass * LPAD;           // first query row of this pass
        const int gi0 = row0 + i0;              // first query row of this lane
        const bool from_park = MP && pass > 0;  // lane 0 takes its top boundary from the parked row
        const bool park = MP && pass + 1 < npass;
        if (MP && pass > 0) __syncthreads();    // every warp is done with the profile of the last pass
                                                // (and its parked row is visible to the whole CTA)

        // ---- stage the two queries, then build the pair profile ----------------
        {
            const uint8_t *ca = p.codes + p.seq_off[qa] + row0;
            const uint8_t *cb = p.codes + p.seq_off[qb >= 0 ? qb : qa] + row0;
            for (int i = tid; i < LPAD; i += CTA_THREADS) {
                qcodes[i] = (row0 + i < La) ? ca[i] : (uint8_t)255;
                qcodes[LPAD + i] = (row0 + i < Lb) ? cb[i] : (uint8_t)255;
            }
        }
        __syncthreads();
        {
            uint32_t *pw = reinterpret_cast<uint32_t *>(prof);
            for (int idx = tid; idx < PROF_ROWS * LPAD; idx += CTA_THREADS) {
                const int b = idx / LPAD, pos = idx - b * LPAD;
                const int ca = qcodes[pos], cb = qcodes[LPAD + pos];
                // profile word = score - open; rows past the query (and the token rows) score 0
                const int sa = ((ca < NSYM && b < NSYM) ? ssub[ca * NSYM + b] : 0) - p.open;
                const int sb = ((cb < NSYM && b < NSYM) ? ssub[cb * NSYM + b] : 0) - p.open;
                const int tt = pos / K, k = pos - tt * K;
                const uint32_t word = W32 ? (uint32_t)sa : (((uint32_t)sa & 0xFFFFu) | ((uint32_t)sb << 16));
#pragma unroll
                for (int r = 0; r < Cfg::REP; ++r)
                    pw[(((b * CH) + (k >> 2)) * Cfg::W + r * G + tt) * 4 + (k & 3)] = word;
            }
        }
        __syncthreads();

        // ---- this group's subject stream ------------------------------------------
        const int s_first = T.s_begin + gid;
        const int n_g = (s_first < T.s_end) ? (T.s_end - s_first + NGROUPS - 1) / NGROUPS : 0;
        // the subject of the last group of this warp in batch o: the longest of the warp's batch
        // when lengths ascend (they do inside a type block); any other case only aligns less well
        const int s_wlast = T.s_begin + (tid | 31) / G;
        auto stream_len = [&](int o) {          // steps batch o takes in this group's stream
            const int own = p.seq_len[s_first + o * NGROUPS];
            if (!Cfg::ALIGN_STREAMS) return own + 1;
            return max(own, p.seq_len[min(s_wlast + o * NGROUPS, T.s_end - 1)]) + 1;
        };
        int nsteps = 0;
        for (int o = t; o < n_g; o += G) nsteps += stream_len(o);
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) nsteps += __shfl_xor_sync(FULL, nsteps, off);
        nsteps = (n_g > 0) ? nsteps + G - 1 : 0;
        if (G < 32) {
#pragma unroll
            for (int off = G; off < 32; off <<= 1) nsteps = max(nsteps, __shfl_xor_sync(FULL, nsteps, off));
        }
        // the same value in every lane; REDUX leaves it in a uniform register, so the step counter
        // and its compare run on the uniform datapath instead of the alu pipe
        nsteps = (int)__reduce_max_sync(FULL, (unsigned)nsteps);

        // ---- lane state ------------------------------------------------------------
        // Tn[k] = Ho[i0+k-1][j-1] + (sub(q_{i0+k}, residue j) - open)   diagonal term of the NEXT cell
        // E[k]  = E[i0+k][j]                                            already advanced to column j
        uint32_t Tn[K], E[K];
        uint32_t best = 0;
        uint32_t h_out = 0, f_out = 0;
        uint32_t top_h = LOCAL ? open2 : Ops::pack(2 * p.open);     // lane 0: Ho[-1][j]
        // column boundary (j = -1) of the row above this lane: Ho[i0-1][-1]
        const uint32_t hin0 = (LOCAL || gi0 == 0) ? open2 : Ops::pack(2 * p.open + (gi0 - 1) * p.extend);
        // local mode: the top boundary row is constant (Ho = open, F = 0), so lane 0 takes it with
        // one LOP3 per value instead of a compare and two selects per step
        uint32_t keep_in = (t == 0 && !from_park) ? 0u : 0xFFFFFFFFu;
        uint32_t top_or = (t == 0 && !from_park) ? open2 : 0u;
        // through a shuffle with the lane's own index: opaque to ptxas (an empty asm is opaque to the
        // front end only), which otherwise turns the masks back into a predicate that it rebuilds in
        // every step, and two selects
        if (G > 1) {
            keep_in = __shfl_sync(FULL, keep_in, lane);
            top_or = __shfl_sync(FULL, top_or, lane);
        }

        // start of a subject whose first residue is c0: state for column 0
        auto begin_subject = [&](uint32_t c0) {
            const uint32_t a = prof_lane + c0 * (uint32_t)Cfg::ROW_BYTES;
            uint32_t P[CH * 4];
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const uint4 v = lds128(a + ch * (Cfg::W * 16));
                P[ch * 4 + 0] = v.x; P[ch * 4 + 1] = v.y; P[ch * 4 + 2] = v.z; P[ch * 4 + 3] = v.w;
            }
            // Ho[i0+k][-1], walked down the rows.  The empty asm makes the start value opaque so
            // the K loop-invariant boundary words are rebuilt here (cold path) instead of being
            // hoisted into 2K registers that would stay live across the hot loop.
            uint32_t col = LOCAL ? open2 : Ops::pack(2 * p.open + gi0 * p.extend);
            asm volatile("" : "+r"(col));
            uint32_t above = hin0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Tn[k] = Ops::add(above, P[k]);
                // E advanced to column 0 = max(-inf + ext, Ho[i0+k][-1]); local: any value <= 0 is inert
                E[k] = LOCAL ? 0u : col;
                above = col;
                if (!LOCAL) col = Ops::add(col, ext2);
            }
            best = 0;
            top_h = LOCAL ? open2 : Ops::pack(2 * p.open);
        };

        int ord = 0;                                            // subjects this lane has finished
        int wait = (n_g > 0) ? t : 0;                           // idle steps before this lane starts
        // address of the last token fetched.  Every step advances it by one (low half only, see
        // advance_lo32) and loads through it; a lane on idle tokens is put back on the header of idle
        // bytes in the cold path, so the hot loop carries no test of where the pointer stands
        uint64_t tokp = (uint64_t)(uintptr_t)p.codes;           // address of the last token fetched (p.codes = idle header)
        uint32_t c = TOK_IDLE, c1 = TOK_IDLE;                   // token of this step / of the next step
#pragma unroll
        for (int k = 0; k < K; ++k) { Tn[k] = 0; E[k] = 0; }
        if (n_g > 0 && t == 0) {
            const uint8_t *s0 = p.codes + p.seq_off[s_first];
            c = __ldg(s0);
            c1 = __ldg(s0 + 1);
            tokp = (uint64_t)(uintptr_t)(s0 + 1);
            begin_subject(c);
        }

        // multi-pass: the parked row of the previous pass, two steps ahead of its use (lane 0 only;
        // one step ahead leaves the L2 latency exposed: 2 048 residues 7.59 -> 7.26 TCUPS)
        // The sweep is instantiated per role of the pass — RD: it reads the row the previous pass
        // parked, WR: it parks its own bottom row — so that a first pass issues none of the reader's
        // instructions and a last pass none of the writer's (single-pass kernels: neither).
        auto sweep = [&](auto rd_tag, auto wr_tag) {
        constexpr bool RD = MP && decltype(rd_tag)::value, WR = MP && decltype(wr_tag)::value;
        const bool park_reader = RD && t == 0;
        const int park_writer = (WR && t == G - 1) ? 1 : 0;
        int2 bnd = make_int2(0, 0), bnd1 = make_int2(0, 0);
        if (park_reader) { bnd = __ldcg(mp_line); bnd1 = __ldcg(mp_line + 1); }

        for (int tau = 0; tau < nsteps; ++tau) {
            // G = 1: the lane holds the whole query, its top boundary is the constant row below
            uint32_t h_in = (G > 1) ? __shfl_up_sync(FULL, h_out, 1, G) : 0u;
            uint32_t f_in = (G > 1) ? __shfl_up_sync(FULL, f_out, 1, G) : 0u;
            advance_lo32(tokp);
            uint32_t c2 = ldg_token(reinterpret_cast<const uint8_t *>(tokp));
            int2 bnd2 = make_int2(0, 0);
            if (RD) {
                if (park_reader) {              // Ho / F of row row0-1 at this stream position
                    bnd2 = __ldcg(mp_line + tau + 2);
                    h_in = (uint32_t)bnd.x;
                    f_in = (uint32_t)bnd.y;
                }
            }
            if (LOCAL) {                        // top boundary row: Ho[-1][j] = open, F[-1][j] = 0
                if (G > 1) {
                    asm("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(h_in) : "r"(keep_in), "r"(top_or));   // (h & keep) | top
                    f_in &= keep_in;
                } else {
                    h_in = open2;
                }
            } else if (t == 0 && !from_park) {  // global: Ho[-1][j] walks down the boundary
                h_in = top_h;
                f_in = Ops::NEG;
            }
            if (c < TOK_BOUNDARY) {
                // ================= hot path: one residue of the current subject ========
                {
                    const uint32_t a = prof_lane + c1 * (uint32_t)Cfg::ROW_BYTES;   // row of the NEXT token
                    uint32_t P[CH * 4];
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        const uint4 v = lds128(a + ch * (Cfg::W * 16));
                        P[ch * 4 + 0] = v.x; P[ch * 4 + 1] = v.y; P[ch * 4 + 2] = v.z; P[ch * 4 + 3] = v.w;
                    }
                    uint32_t up = h_in, F = f_in, h_prev = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        F = Ops::addmax(F, ext2, up);
                        const uint32_t h = LOCAL ? Ops::max3_relu(Tn[k], E[k], F)
                                                 : Ops::max3(Tn[k], E[k], F);
                        Tn[k] = Ops::add(up, P[k]);              // diagonal term of (k, j+1)
                        up = Ops::add(h, open2);                 // Ho[k][j]
                        E[k] = Ops::addmax(E[k], ext2, up);   // E[k][j+1]
                        if (LOCAL) {
                            if (k & 1) best = Ops::max3(best, h_prev, h);
                            else if (k == K - 1) best = Ops::max2(best, h);
                        }
                        h_prev = h;
                    }
                    h_out = up;
                    f_out = F;
                    if (!LOCAL) top_h = Ops::add(top_h, ext2);
                    if (WR) {                   // bottom row of this pass, by stream position: one
                                                // predicated store, no branch around it
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.global.cg.v2.s32 [%1], {%2, %3};\n\t}"
                                     :: "r"(park_writer), "l"(mp_line + (tau - (G - 1))), "r"((int)up), "r"((int)F) : "memory");
                    }
                }
            } else if (c == TOK_BOUNDARY) {
                // ================= cold path: the subject just ended ====================
                const int s = s_first + ord * NGROUPS;
                if (LOCAL) {
                    int *slot = my_best + (ord & (Cfg::RING - 1)) * 2;
                    const int lo = Ops::lo(best), hi = Ops::hi(best);
                    if (G == 1) {               // the lane holds the whole query: nothing to collect
                        emit_pair(p, qa, s, max(lo, 0));
                        emit_pair(p, qb, s, max(hi, 0));
                    } else if (t != G - 1) {
                        if (lo > 0) atomicMax(slot, lo);
                        if (hi > 0) atomicMax(slot + 1, hi);
                    } else {                    // every other lane passed this boundary earlier
                        int a = max(lo, atomicExch(slot, 0));
                        int b = max(hi, atomicExch(slot + 1, 0));
                        if (MP) {               // best of the earlier passes of this subject
                            int *pb = pbest + (gid * Cfg::MP_MAX_SUBJECTS + ord) * 2;
                            if (pass > 0) { a = max(a, pb[0]); b = max(b, pb[1]); }
                            if (park) { pb[0] = a; pb[1] = b; }
                        }
                        if (!park) {
                            emit_pair(p, qa, s, a);
                            emit_pair(p, qb, s, b);
                        }
                    }
                } else {
                    // H[row][last column] is still in the lane's state: the last column pass built
                    // Tn[r+1] = Ho[r] + (profile row of the boundary token = 0 - open) = H[r], and the
                    // bottom row of the lane left h_out = Ho[K-1].  Picked out here, on the cold path,
                    // so the hot loop carries no capture code.
                    auto final_h = [&](int r) {
                        uint32_t v = Ops::sub(h_out, open2);
                        // one predicated select per row, as opaque asm: written as a plain
                        // "if (r + 1 == k) v = Tn[k]" chain, ptxas turns Tn[] into an indexed local
                        // array and stores it to the stack in every hot step
#pragma unroll
                        for (int k = 1; k < K; ++k)
                            asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %1, %2;\n\tselp.b32 %0, %3, %0, q;\n\t}"
                                : "+r"(v) : "r"(r + 1), "r"(k), "r"(Tn[k]));
                        return v;
                    };
                    const int ra = La - 1 - gi0, rb = Lb - 1 - gi0;   // owned by this lane in this pass?
                    if (ra >= 0 && ra < K) emit_pair(p, qa, s, Ops::lo(final_h(ra)));
                    if (rb >= 0 && rb < K) emit_pair(p, qb, s, Ops::hi(final_h(rb)));
                }
                // idle steps that keep the warp's streams in step (0 unless ALIGN_STREAMS)
                const int pad = Cfg::ALIGN_STREAMS ? stream_len(ord) - 1 - p.seq_len[s] : 0;
                ++ord;
                if (ord < n_g && pad == 0) {
                    const uint8_t *sn = p.codes + p.seq_off[s_first + ord * NGROUPS];
                    c1 = __ldg(sn);
                    c2 = __ldg(sn + 1);
                    tokp = (uint64_t)(uintptr_t)(sn + 1);
                    begin_subject(c1);
                } else {
                    c1 = TOK_IDLE;
                    c2 = TOK_IDLE;
                    tokp = (uint64_t)(uintptr_t)p.codes;
                    wait = (ord < n_g) ? pad : 0;
                }
            } else {
                // ================= cold path: idle token (wind-up of lane t, between two subjects of
                // an aligned stream, or stream finished) ====
                tokp = (uint64_t)(uintptr_t)p.codes;        // stay on the header of idle bytes
                if (wait > 0 && --wait == 0) {
                    const uint8_t *s0 = p.codes + p.seq_off[s_first + ord * NGROUPS];
                    c1 = __ldg(s0);
                    c2 = __ldg(s0 + 1);
                    tokp = (uint64_t)(uintptr_t)(s0 + 1);
                    begin_subject(c1);
                }
            }
            c = c1;
            c1 = c2;
            if (RD) { bnd = bnd1; bnd1 = bnd2; }
        }
        };  // sweep
        if (!MP) sweep(std::false_type{}, std::false_type{});
        else if (from_park && park) sweep(std::true_type{}, std::true_type{});
        else if (from_park) sweep(std::true_type{}, std::false_type{});
        else if (park) sweep(std::false_type{}, std::true_type{});
        else sweep(std::false_type{}, std::false_type{});
      }   // pass
    }
}

}  // namespace bgca
