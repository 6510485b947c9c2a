// K7 (second form) — alignment statistics without traceback for sequences of at most 1023
// residues, sm_100a.  Same definition as gotoh_stats.cuh / oracle bgco_stats_pair: the alignment
// reported is the lexicographic maximum of
//     (score, identities, -columns, q_start, s_start, -s_end, -q_end)
// over all alignments of the affine-gap model; bit-exact against the oracle.
//
// What makes it fast.  A DP value is the 64-bit word
//     hi = (score + 2^19) << 10 | identities          lo = w << 20 | q_start << 10 | s_start
// with w = (aligned residue pairs) + q_start + s_start.  At ONE cell (i, j) every path has
// columns = i + j - w, so ordering by w is ordering by -columns, and — unlike a column count —
// w does not change on a gap step: E and F carry lo untouched and only the diagonal step adds
// to it.  The bias keeps hi positive with an exponent field that is never all ones, so the word
// read as an IEEE double is a positive finite number and
//     lexicographic max of (hi, lo)  ==  max of the doubles:
// one DSETP on the FP64 pipe (idle in every other kernel here) plus two selects, instead of two
// integer compares and two selects on the alu pipe (tools/fp64_probe.cu: 43 vs 28 max/clk/SM).
// Where only hi changes (E, F) hi itself comes from one DPX VIADDMNMX.  Per cell: 4 DSETP,
// 7 selects, 2 VIADDMNMX, 1 VIMNMX, 1 ISETP, 5 adds (+ 6 for the running best in local mode).
//
// HINT (local mode): when the caller knows the best score of every pair — it does whenever the
// pairs were chosen by score from the scoring kernels' matrix — only cells that reach that score
// can end the reported alignment.  The running best leaves the hot loop: per four rows one max
// and one compare against the threshold, and a cold block that does the full lexicographic
// comparison for the cells that pass (a handful per pair).  A hint that is too low only makes
// the cold block run more often; one that is too high finds no candidate, which the record
// shows (score < hint) and the host answers by running the pair again without hints.
//
// Layout of the work.  Pairs arrive grouped by query.  A CTA (4 warps) takes a work item = one
// query and a run of its subjects, builds the query's profile in shared memory —
// prof[letter][chunk][lane][4] = ((sub[q_i][letter] - open) << 10) | (q_i == letter), the same
// conflict-free layout as the scoring kernel — and every warp sweeps its subjects systolically
// (lane t owns rows [t*KS, (t+1)*KS), lags lane t-1 by one column, hands (Ho, Hoe, lo) and F down
// with shuffles) as ONE token stream, so the 31-step wind-up is paid once per work item.  At the
// end of a subject each lane posts its best to a shared-memory slot (plain read-modify-write:
// the lanes of a warp reach that point in different steps), the last lane writes the record.
#pragma once
#include "common.cuh"

namespace bgca {

constexpr int ST2_THREADS = 128;
constexpr int ST2_WARPS = ST2_THREADS / 32;
constexpr int ST2_BIAS = 1 << 19;
constexpr uint32_t ST2_ZHI = (uint32_t)ST2_BIAS << 10;     // hi of the empty alignment (score 0, 0 identities)
constexpr uint32_t ST2_DCOL = 1u << 20;                    // one aligned residue pair in lo
constexpr int ST2_MAXLEN = 1023;
constexpr uint32_t ST2_TOK_BOUNDARY = BGCA_PAD_CODE, ST2_TOK_IDLE = 25;
constexpr int ST2_PROF_ROWS = NSYM + 2;

struct Stats2Item {
    int32_t q;        // query, SORTED index
    int32_t first;    // its subjects are subj[first, first + count)
    int32_t count;
    int32_t pad;
};

struct Stats2Params {
    const uint8_t *codes;
    const uint32_t *seq_off;
    const int32_t *seq_len;
    const Stats2Item *items;
    int32_t n_items;
    int32_t *item_counter;
    const int32_t *subj;       // [npairs] SORTED subject index, grouped by query
    const int32_t *out_idx;    // [npairs] position of the pair in the caller's list (NULL: its own index)
    const int32_t *hint;       // [npairs] HINT kernels: best local score of the pair, grouped order
    int32_t *redo_count;       // HINT kernels: pairs whose hint turned out too high
    bgca_alnstats *out;
    int32_t open, extend;
    int32_t openh, exth;       // open * 1024, extend * 1024: the gap scores in units of hi
    int8_t sub[NSYM * NSYM];
};

template <int KS>
struct Stats2Cfg {
    static constexpr int CH = (KS + 3) / 4;
    static constexpr int LPAD = 32 * KS;
    static constexpr int ROW_BYTES = CH * 32 * 16;
    static constexpr int PROF_BYTES = ST2_PROF_ROWS * ROW_BYTES;
    static constexpr int SLOT_BYTES = ST2_WARPS * 32 * 16;           // per warp a ring of 32 best slots
    static constexpr int SMEM_BYTES = PROF_BYTES + SLOT_BYTES + LPAD + NSYM * NSYM + 16;
    static constexpr int MIN_CTAS = (KS <= 8) ? 6 : (KS <= 12) ? 5 : (KS <= 16) ? 4 : (KS <= 20) ? 3 : 2;
    // The kernels with score hints and the global ones carry a few more live values: one CTA per SM
    // less where the tighter register budget made ptxas spill inside the step (measured with hints /
    // global, GCUPS: 240-residue domains 706 / 746 -> 795 / 818, 370: 762 / 810 -> 787 / 831,
    // 500: 713 / 782 -> 730 / 797; the 20-row class is better off where it was).
    // The 14-row class with hints keeps four (12 bytes of spill; at three: 823 -> 767 GCUPS).
    static constexpr int MIN_CTAS_HINT = (KS <= 8) ? 5 : (KS <= 14) ? 4 : (KS <= 20) ? 3 : 2;
    static constexpr int MIN_CTAS_GLOBAL = (KS <= 8) ? 5 : (KS <= 12) ? 4 : (KS <= 20) ? 3 : 2;
};

// A DP value lives in a double (the 64-bit word described above, a positive finite number): the
// register allocator then keeps hi and lo in an aligned pair, which is what DSETP wants — held as
// two 32-bit variables the pairs are assembled with moves in front of every compare.
using st2_t = double;
__device__ __forceinline__ st2_t st2_make(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint32_t st2_hi(st2_t v) { return (uint32_t)__double2hiint(v); }
__device__ __forceinline__ uint32_t st2_lo(st2_t v) { return (uint32_t)__double2loint(v); }

template <int MODE, int KS, int HINT = 0>
__global__ void __launch_bounds__(ST2_THREADS, (HINT ? Stats2Cfg<KS>::MIN_CTAS_HINT : MODE == BGCA_MODE_GLOBAL ? Stats2Cfg<KS>::MIN_CTAS_GLOBAL : Stats2Cfg<KS>::MIN_CTAS))
gotoh_stats2_kernel(const Stats2Params p)
{
    using Cfg = Stats2Cfg<KS>;
    constexpr int CH = Cfg::CH, LPAD = Cfg::LPAD;
    constexpr bool LOCAL = (MODE == BGCA_MODE_LOCAL);
    static_assert(!HINT || LOCAL, "score hints only make sense for local alignments");
    constexpr unsigned FULL = 0xFFFFFFFFu;

    extern __shared__ uint4 smem_u4[];
    uint4 *prof = smem_u4;                                                        // [PROF_ROWS][CH][32]
    uint4 *slots = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(smem_u4) + Cfg::PROF_BYTES);
    uint8_t *qcodes = reinterpret_cast<uint8_t *>(slots) + Cfg::SLOT_BYTES;       // [LPAD]
    int8_t *ssub = reinterpret_cast<int8_t *>(qcodes + LPAD);
    __shared__ int s_item;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = lane * KS;
    for (int i = tid; i < NSYM * NSYM; i += ST2_THREADS) ssub[i] = p.sub[i];
    for (int i = tid; i < ST2_WARPS * 32; i += ST2_THREADS) slots[i] = make_uint4(0, 0, 0, 0);

    // gap scores in hi units, multiplied on the host: computed here, ptxas folds the shift into
    // every use as an IMAD and the fused add-max (VIADDMNMX) falls apart into IMAD + VIMNMX
    const int EXTH = p.exth, OPENH = p.openh;
    const uint32_t prof_lane = (uint32_t)__cvta_generic_to_shared(prof) + (uint32_t)lane * 16u;
    volatile uint4 *const my_slots = slots + warp * 32;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.item_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= p.n_items) break;
        const Stats2Item it = p.items[item];
        const int La = p.seq_len[it.q];
        {
            const uint8_t *cq = p.codes + p.seq_off[it.q];
            for (int i = tid; i < LPAD; i += ST2_THREADS) qcodes[i] = (i < La) ? cq[i] : (uint8_t)255;
        }
        __syncthreads();
        {
            uint32_t *pw = reinterpret_cast<uint32_t *>(prof);
            for (int idx = tid; idx < ST2_PROF_ROWS * LPAD; idx += ST2_THREADS) {
                const int b = idx / LPAD, pos = idx - b * LPAD;
                const int cq = qcodes[pos];
                // rows past the query: local mode takes a score that always loses to the empty
                // alignment; the token rows (never scored) and global padding rows score 0
                int s = (cq < NSYM && b < NSYM) ? (int)ssub[cq * NSYM + b] : (LOCAL && b < NSYM ? -4096 : 0);
                const int word = (s - p.open) * 1024 + ((cq < NSYM && cq == b) ? 1 : 0);
                const int tt = pos / KS, k = pos - tt * KS;
                pw[(((b * CH) + (k >> 2)) * 32 + tt) * 4 + (k & 3)] = (uint32_t)word;
            }
        }
        __syncthreads();

        // ---- this warp's subject stream: pairs first + warp, + ST2_WARPS, ... ----------------
        const int n_w = (it.count > warp) ? (it.count - warp + ST2_WARPS - 1) / ST2_WARPS : 0;
        auto pair_of = [&](int o) { return it.first + warp + o * ST2_WARPS; };
        int nsteps = 0;
        for (int o = lane; o < n_w; o += 32) nsteps += p.seq_len[p.subj[pair_of(o)]] + 1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) nsteps += __shfl_xor_sync(FULL, nsteps, off);
        nsteps = (n_w > 0) ? nsteps + 31 : 0;
        // the same value in every lane; out of a REDUX it sits in a uniform register and the step
        // counter leaves the alu pipe (724 -> 766 GCUPS local, 792 -> 823 with hints, 823 -> 849 global)
        nsteps = (int)__reduce_max_sync(FULL, (unsigned)nsteps);

        // ---- lane state ---------------------------------------------------------------------
        // Tn = H[i0+k-1][j-1] + delta(q_{i0+k}, residue j): the diagonal term of the NEXT column
        // E  = E[i0+k][j], already advanced to column j            (hi in H units, lo as above)
        st2_t Tn[KS], E[KS];
        const st2_t EMPTY = st2_make(ST2_ZHI, 2047u << 20);              // position-independent form
        st2_t best = EMPTY;                                               // local: best so far
        uint32_t bend = 0xFFFFFFFFu;
        st2_t cap = 0.0;                                                  // global: H[La-1][Ls-1]
        uint32_t ho_out = 0, hoe_out = 0, hl_out = 0, fh_out = 0, fl_out = 0;
        int ord = 0, wait = (n_w > 0) ? lane : 0;
        int j = 0;                                   // column of the token this lane handles now
        int Ls_cur = 0;
        uint32_t off = 0, c = ST2_TOK_IDLE, c1 = ST2_TOK_IDLE;
        // global: Ho / Hoe of the top boundary row H[-1][j] = open + j*ext (lane 0)
        uint32_t top_ho = 0, top_hoe = 0;
        int thr = 0x7FFFFFFF;                        // HINT: hi of (hinted score, 0 identities)

        auto begin_subject = [&](uint32_t c0, int s, int pair) {
            const uint32_t a = prof_lane + c0 * (uint32_t)Cfg::ROW_BYTES;
            uint32_t P[CH * 4];
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const uint4 v = *reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(a + ch * 512));
                P[ch * 4 + 0] = v.x; P[ch * 4 + 1] = v.y; P[ch * 4 + 2] = v.z; P[ch * 4 + 3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const int i = i0 + k;                // query row
                if (LOCAL) {
                    // H[i-1][-1] = empty alignment starting at (i, 0): Ho = ZHI + open
                    Tn[k] = st2_make(ST2_ZHI + (uint32_t)OPENH + P[k],
                                     (((uint32_t)i << 20) | ((uint32_t)i << 10)) + ST2_DCOL);
                    E[k] = 0.0;                      // -inf: loses to everything
                } else {
                    // H[i-1][-1] = a gap of i residues from the corner (0 for i = 0); lo = 0
                    const int hprev = (i == 0) ? 0 : p.open + (i - 1) * p.extend;
                    Tn[k] = st2_make((uint32_t)((hprev + ST2_BIAS) * 1024 + OPENH) + P[k], ST2_DCOL);
                    // E advanced to column 0 = H[i][-1] + open
                    E[k] = st2_make((uint32_t)((2 * p.open + i * p.extend + ST2_BIAS) * 1024), 0u);
                }
            }
            best = EMPTY; bend = 0xFFFFFFFFu;
            if (HINT) thr = (min(max(p.hint[pair], 0), 1 << 20) + ST2_BIAS) * 1024;   // any caller value stays in range
            j = 0;
            Ls_cur = p.seq_len[s];
            top_ho = (uint32_t)((2 * p.open + ST2_BIAS) * 1024);
            top_hoe = top_ho - (uint32_t)EXTH;
        };

        if (n_w > 0 && lane == 0) {
            const int s = p.subj[pair_of(0)];
            const uint8_t *s0 = p.codes + p.seq_off[s];
            c = __ldg(s0);
            c1 = __ldg(s0 + 1);
            off = p.seq_off[s] + 1;
            begin_subject(c, s, pair_of(0));
        } else {
#pragma unroll
            for (int k = 0; k < KS; ++k) { Tn[k] = 0.0; E[k] = 0.0; }
        }

        for (int tau = 0; tau < nsteps; ++tau) {
            uint32_t ho_in = __shfl_up_sync(FULL, ho_out, 1);
            uint32_t hoe_in = __shfl_up_sync(FULL, hoe_out, 1);
            uint32_t hl_in = __shfl_up_sync(FULL, hl_out, 1);
            uint32_t fh_in = __shfl_up_sync(FULL, fh_out, 1);
            uint32_t fl_in = __shfl_up_sync(FULL, fl_out, 1);
            off += 1;
            uint32_t c2 = __ldg(p.codes + off);
            if (lane == 0) {
                if (LOCAL) {                    // H[-1][j] = empty alignment starting at (0, j+1)
                    ho_in = ST2_ZHI + (uint32_t)OPENH;
                    hoe_in = ho_in - (uint32_t)EXTH;
                    hl_in = ((uint32_t)(j + 1) << 20) | (uint32_t)(j + 1);
                } else {
                    ho_in = top_ho;
                    hoe_in = top_hoe;
                    hl_in = 0;
                }
                fh_in = 0;                      // -inf
                fl_in = 0;
            }
            if (c < ST2_TOK_BOUNDARY) {
                // ================= hot path: one residue of the current subject ===============
                const uint32_t a = prof_lane + c1 * (uint32_t)Cfg::ROW_BYTES;    // row of the NEXT token
                uint32_t P[CH * 4];
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) {
                    uint4 v;
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                        : "r"(a + ch * 512));
                    P[ch * 4 + 0] = v.x; P[ch * 4 + 1] = v.y; P[ch * 4 + 2] = v.z; P[ch * 4 + 3] = v.w;
                }
                uint32_t uho = ho_in;                              // Ho of the row above, this column
                st2_t uhoe = st2_make(hoe_in, hl_in);              // the same value less ext, with its lo
                st2_t f = st2_make(fh_in, fl_in);
                // local: lo of the empty alignment that starts BEHIND cell (i0+k, j): (i+1, j+1), and
                // the term that turns w into 2047 - columns for a path that ends there
                const uint32_t z0 = (((uint32_t)(i0 + 1 + j + 1)) << 20) | ((uint32_t)(i0 + 1) << 10) | (uint32_t)(j + 1);
                const uint32_t cz0 = (uint32_t)(2047 - (i0 + 1 + j + 1)) << 20;
                const bool last_col = !LOCAL && (j == Ls_cur - 1);
                uint32_t h4h[4], h4l[4];                           // HINT: the last four rows' H
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    // F[k][j] = max(F[k-1][j] + ext, H[k-1][j] + open): hi by DPX, lo follows the compare
                    const bool pf = f > uhoe;
                    f = st2_make((uint32_t)__viaddmax_s32((int)st2_hi(f), EXTH, (int)uho),
                                 pf ? st2_lo(f) : st2_lo(uhoe));
                    // off the critical chain: max(diagonal term, E[, empty])
                    st2_t m = (Tn[k] > E[k]) ? Tn[k] : E[k];
                    if (LOCAL) {                // the empty alignment wins every tie on hi (fewer columns)
                        const bool pz = (int)st2_hi(m) > (int)ST2_ZHI;
                        m = st2_make((uint32_t)max((int)st2_hi(m), (int)ST2_ZHI),
                                     pz ? st2_lo(m) : z0 + (uint32_t)k * ((1u << 20) | (1u << 10)));
                    }
                    const st2_t h = (m > f) ? m : f;
                    // diagonal term of (k, j+1): H[k-1][j] + delta
                    Tn[k] = st2_make(uho + P[k], st2_lo(uhoe) + ST2_DCOL);
                    uho = st2_hi(h) + (uint32_t)OPENH;
                    uhoe = st2_make(uho - (uint32_t)EXTH, st2_lo(h));
                    // E[k][j+1] = max(E[k][j] + ext, H[k][j] + open)
                    const bool pe = E[k] > uhoe;
                    E[k] = st2_make((uint32_t)__viaddmax_s32((int)st2_hi(E[k]), EXTH, (int)uho),
                                    pe ? st2_lo(E[k]) : st2_lo(h));
                    if (LOCAL && !HINT) {
                        // running best: strict > keeps the first cell in column-major order, i.e. the
                        // smallest s_end, then the smallest q_end
                        const st2_t cand = st2_make(st2_hi(h), st2_lo(h) + cz0 - (uint32_t)k * (1u << 20));
                        if (cand > best) {
                            best = cand;
                            bend = ((uint32_t)(j + 1) << 16) | (uint32_t)(i0 + k + 1);
                        }
                    } else if (LOCAL) {
                        h4h[k & 3] = st2_hi(h);
                        h4l[k & 3] = st2_lo(h);
                        if ((k & 3) == 3 || k == KS - 1) {
                            const int k0 = k & ~3;
                            int mx = (int)h4h[0];
#pragma unroll
                            for (int r = 1; r <= k - k0; ++r) mx = max(mx, (int)h4h[r]);
                            if (mx >= thr) {        // cold: some cell of these rows reaches the hinted score
#pragma unroll
                                for (int r = 0; r <= k - k0; ++r) {
                                    const st2_t cand = st2_make(h4h[r], h4l[r] + cz0 - (uint32_t)(k0 + r) * (1u << 20));
                                    if ((int)h4h[r] >= thr && cand > best) {
                                        best = cand;
                                        bend = ((uint32_t)(j + 1) << 16) | (uint32_t)(i0 + k0 + r + 1);
                                    }
                                }
                            }
                        }
                    } else if (last_col) {
                        if (i0 + k == La - 1) cap = h;
                    }
                }
                const uint32_t fh = st2_hi(f), fl = st2_lo(f), uhoe_h = st2_hi(uhoe), ul = st2_lo(uhoe);
                ho_out = uho; hoe_out = uhoe_h; hl_out = ul; fh_out = fh; fl_out = fl;
                if (!LOCAL) { top_ho += (uint32_t)EXTH; top_hoe += (uint32_t)EXTH; }
                ++j;
            } else if (c == ST2_TOK_BOUNDARY) {
                // ================= cold path: the subject just ended ============================
                volatile uint4 *slot = my_slots + (ord & 31);
                if (LOCAL) {
                    // lanes reach this point in consecutive steps, so the slot is updated one lane
                    // at a time; smaller s_end, then q_end, wins a full tie
                    const uint4 cur = make_uint4(slot->x, slot->y, slot->z, slot->w);
                    const uint32_t bh = st2_hi(best), bl = st2_lo(best);
                    const bool better = cur.w == 0 || best > st2_make(cur.x, cur.y) ||
                                        (bh == cur.x && bl == cur.y && bend < cur.z);
                    if (better) { slot->x = bh; slot->y = bl; slot->z = bend; slot->w = 1; }
                } else if (La - 1 >= i0 && La - 1 < i0 + KS) {
                    slot->x = st2_hi(cap); slot->y = st2_lo(cap); slot->z = 0; slot->w = 1;
                }
                if (lane == 31) {
                    const uint32_t rh = slot->x, rl = slot->y, re = slot->z;
                    slot->w = 0;
                    const int pair = pair_of(ord);
                    bgca_alnstats r;
                    r.score = (int)(rh >> 10) - ST2_BIAS;
                    if (HINT) {                 // no cell reached the hinted score: the hint was too high
                        if (r.score < min(max(p.hint[pair], 0), 1 << 20)) atomicAdd(p.redo_count, 1);
                    }
                    r.identities = (int)(rh & 1023u);
                    if (LOCAL) {
                        r.columns = 2047 - (int)(rl >> 20);
                        if (r.columns == 0) {
                            r.q_start = r.q_end = r.s_start = r.s_end = -1;
                            r.gaps = 0;
                        } else {
                            r.q_start = (int)((rl >> 10) & 1023u);
                            r.s_start = (int)(rl & 1023u);
                            r.q_end = (int)(re & 0xFFFFu);
                            r.s_end = (int)(re >> 16);
                            r.gaps = 2 * r.columns - (r.q_end - r.q_start) - (r.s_end - r.s_start);
                        }
                    } else {
                        const int Ls = p.seq_len[p.subj[pair]];
                        r.columns = La + Ls - (int)(rl >> 20);
                        r.q_start = 0; r.s_start = 0; r.q_end = La; r.s_end = Ls;
                        r.gaps = 2 * r.columns - La - Ls;
                    }
                    p.out[p.out_idx ? p.out_idx[pair] : pair] = r;
                }
                ++ord;
                if (ord < n_w) {
                    const int s = p.subj[pair_of(ord)];
                    const uint8_t *sn = p.codes + p.seq_off[s];
                    c1 = __ldg(sn);
                    c2 = __ldg(sn + 1);
                    off = p.seq_off[s] + 1;
                    begin_subject(c1, s, pair_of(ord));
                } else {
                    c1 = ST2_TOK_IDLE;
                    c2 = ST2_TOK_IDLE;
                    off = 0;
                }
            } else {
                // ================= cold path: idle token (wind-up, or stream finished) ===========
                off = 0;
                if (wait > 0 && --wait == 0) {
                    const int s = p.subj[pair_of(0)];
                    const uint8_t *s0 = p.codes + p.seq_off[s];
                    c1 = __ldg(s0);
                    c2 = __ldg(s0 + 1);
                    off = p.seq_off[s] + 1;
                    begin_subject(c1, s, pair_of(0));
                }
            }
            __syncwarp();       // slot updates of this step are visible to the lanes that follow
            c = c1;
            c1 = c2;
        }
    }
}

}  // namespace bgca
