// C ABI of the BGC aligner engine (include/bgca.h): engine lifetime, packer
// driver, tile planner (K6) and kernel dispatch.  Host code; the kernels live in
// gotoh_pair16.cuh / gotoh_s32.cuh / pack_reduce.cu.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "gotoh_stats2.cuh"
#include "launchers.h"
#include "variants.h"

using namespace bgca;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(BGCA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct VariantDesc { int G, K; };
const VariantDesc kVariants[] = {
#define X(G, K) {G, K},
    BGCA_VARIANTS(X)
#undef X
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kSubjectsPerGroup = 16;
constexpr uint64_t kPackedHeader = 64;    // bytes of TOK_IDLE (25) in front of the first sequence
constexpr int kMpTokensPerGroup = 32768;  // multi-pass tiles: target length of one group's subject stream

// Relative issue cost per subject column with variant (G,K): (instructions per step:
// 5.5 per row + ~21 of step overhead + one LDS.128 per 4 rows) x (share of the warp one
// alignment uses).  Subjects are streamed, so there is no per-subject wind-up term.
double variant_cost(int G, int K, int /*Lq*/)
{
    return (5.5 * K + 21.0 + (K + 3) / 4) * (double)G;
}

// BGCA_MIN_G=n keeps the planner off group widths below n (tuning aid, e.g. 4 = the round-1 set).
int min_group_width()
{
    const char *s = std::getenv("BGCA_MIN_G");
    return s ? std::atoi(s) : 1;
}

// BGCA_FORCE_G=4|8|16|32 restricts the planner to one group width (tuning aid).
int forced_group_width()
{
    const char *s = std::getenv("BGCA_FORCE_G");
    return s ? std::atoi(s) : 0;
}

// n_subjects: subjects the query pair meets in its block.  G = 4 / 2 / 1 run 64 / 128 / 256 subject
// streams per CTA and are only offered when there are enough subjects to keep them busy.
int pick_variant_uncached(int Lq, int force_g, int min_g, int n_subjects)
{
    int best = 0;
    double best_cost = 0.0;
    if (Lq > BGCA_ONE_PASS_ROWS) {
        // multi-pass whole-warp sweep: the fewest passes of <= 1024 rows, rows spread evenly
        const int npass = (Lq + BGCA_ONE_PASS_ROWS - 1) / BGCA_ONE_PASS_ROWS;
        const int K = (Lq + 32 * npass - 1) / (32 * npass);      // 17..32
        return BGCA_VARIANT_MP | BGCA_VARIANT_ID(32, K);
    }
    for (int v = 0; v < kNumVariants; ++v) {
        const int G = kVariants[v].G, K = kVariants[v].K;
        if (G * K < Lq) continue;
        if (G < min_g) continue;
        if (G <= 4 && force_g != G && n_subjects < 4 * (CTA_THREADS / G)) continue;
        if (force_g && G != force_g && force_g * 32 >= Lq) continue;
        const double c = variant_cost(G, K, Lq);
        if (!best || c < best_cost) {
            best = BGCA_VARIANT_ID(G, K);
            best_cost = c;
        }
    }
    return best;   // 0: no packed variant covers Lq
}

// One plan asks for the variant of every query pair: the answer depends on the query length and
// on which of the narrow group widths (G = 4, 2, 1) the number of subjects allows, so a plan
// keeps a small table instead of scanning the variant list (and the environment) per pair.
struct VariantPicker {
    int force_g = forced_group_width(), min_g = min_group_width();
    std::vector<int> table[4];    // by narrow-width class, then by Lq (<= BGCA_ONE_PASS_ROWS)
    int operator()(int Lq, int n_subjects)
    {
        if (Lq > BGCA_ONE_PASS_ROWS) return pick_variant_uncached(Lq, force_g, min_g, n_subjects);
        const int cls = (n_subjects >= 4 * CTA_THREADS) ? 3 : (n_subjects >= 2 * CTA_THREADS) ? 2 :
                        (n_subjects >= CTA_THREADS) ? 1 : 0;
        std::vector<int> &t = table[cls];
        if (t.empty()) t.assign(BGCA_ONE_PASS_ROWS + 1, -1);
        if (t[Lq] < 0) {
            // the smallest subject count of the class stands for all of it (thresholds are 4 * 256 / G)
            static const int rep[4] = {1, CTA_THREADS, 2 * CTA_THREADS, 4 * CTA_THREADS};
            t[Lq] = pick_variant_uncached(Lq, force_g, min_g, rep[cls]);
        }
        return t[Lq];
    }
};

// the 32-bit form of the multi-pass kernel for a query of L residues: fewest passes, rows spread evenly
int w32_variant(int L)
{
    const int npass = (L + BGCA_ONE_PASS_ROWS - 1) / BGCA_ONE_PASS_ROWS;
    const int K = std::max(17, (L + 32 * npass - 1) / (32 * npass));
    return BGCA_VARIANT_MP | BGCA_VARIANT_W32 | BGCA_VARIANT_ID(32, K);
}

// int16 safety of one tile: local scores are bounded by max_sub * min(len),
// global intermediates by the all-gap path.
bool fits_int16(int mode, int max_abs_sub, int open, int extend, int Lq, int Ls_max)
{
    const long long hi = (long long)max_abs_sub * std::min(Lq, Ls_max);
    if (hi > 30000) return false;
    if (mode == BGCA_MODE_GLOBAL) {
        const long long lo = 3LL * -open + (long long)(-extend) * (Lq + Ls_max + 2) +
                             (long long)max_abs_sub;   // magnitude of the most negative cell
        if (lo > 29000) return false;
    }
    return true;
}

// cross plans take the queries of a block as its leading run of role-0 sequences: true when
// every block the planner will cut ([0,n) or one per type) is sorted queries-first.
bool roles_partitioned(const int32_t *type, const int32_t *role, int32_t n, int same_type_only)
{
    for (int i = 1; i < n; ++i) {
        const bool same_block = !(same_type_only && type) || type[i] == type[i - 1];
        if (same_block && role[i] < role[i - 1]) return false;
    }
    return true;
}

struct PlanResult {
    std::vector<bgca_tile> tiles;   // this rank's tiles, grouped by variant, heavy first
    bgca_plan_info info{};
};

// K6 — tile scheduler.  See bgca.h.  All indices are SORTED-space.
// role (nullable): 0 = query, 1 = database; sorted as (type, role, length).  cross_only: only
// query x database pairs are planned, plus one self-only tile per sequence pair (reserved = 1)
// so that the self scores the similarity is normalised with exist.
// n_db_first > 0 (cross plans only): sorted [0, n_db_first) is a persistent database sorted by
// (type, length) and [n_db_first, n) the appended queries, sorted the same way; cross_only == 2
// then leaves out the database's self tiles (its self scores are cached by the caller).
PlanResult make_plan(const int32_t *len, const int32_t *type, const int32_t *role, int32_t n,
                     int same_type_only, int cross_only, int32_t n_db_first, int mode, int max_abs_sub,
                     int open, int extend, int rank, int world)
{
    PlanResult R;
    VariantPicker pick_variant;
    (void)mode;   // local and global kernels have the same register budget
    std::vector<int64_t> pre(n + 1, 0);
    for (int i = 0; i < n; ++i) pre[i + 1] = pre[i] + len[i];
    // The packer sorts by (type, role, length): lengths are non-decreasing only inside one run.
    // run_end[i] = end of the non-decreasing run that holds i, so the longest subject of any
    // range is the maximum over the last elements of the few runs it crosses.
    std::vector<int32_t> run_end(n);
    for (int i = n - 1; i >= 0; --i) run_end[i] = (i + 1 < n && len[i + 1] >= len[i]) ? run_end[i + 1] : i + 1;
    auto max_len_in = [&](int s0, int s1) {
        int mx = 0;
        for (int i = s0; i < s1;) {
            const int e = std::min<int>(run_end[i], s1);
            mx = std::max(mx, len[e - 1]);
            i = e;
        }
        return mx;
    };

    // subjects per alignment group and tile: 16 keeps >= 100 tiles per CTA slot on small
    // batches; larger batches take longer tiles (still > 150 per CTA slot at 10k sequences) so
    // that the tile list, and with it the plan time inside every end-to-end call, stays small
    int spg = std::min(64, std::max(kSubjectsPerGroup, n / 300));

    // exact accounting of unique pairs (i<=j) per tile
    auto tile_pairs = [&](const bgca_tile &T, int64_t &pairs, int64_t &cells) {
        if (T.reserved) {                    // self-only tile of a cross plan
            pairs = (T.qb >= 0) ? 2 : 1;
            cells = (int64_t)len[T.qa] * len[T.qa] + (T.qb >= 0 ? (int64_t)len[T.qb] * len[T.qb] : 0);
            return;
        }
        const int64_t sumL = pre[T.s_end] - pre[T.s_begin];
        pairs = T.s_end - T.s_begin;
        cells = (int64_t)len[T.qa] * sumL;
        if (T.qb >= 0) {
            int64_t p2 = T.s_end - T.s_begin, sl2 = sumL;
            if (T.s_begin <= T.qa && T.qa < T.s_end) { p2 -= 1; sl2 -= len[T.qa]; }   // (qb, qa) mirrored
            pairs += p2;
            cells += (int64_t)len[T.qb] * sl2;
        }
    };

    // Assignment over the exact cell counts: greedy "least-loaded rank takes the next tile".
    // Small batches collect every tile and visit them heaviest-first (LPT: the order matters when
    // a rank gets only a handful of tiles).  Large batches assign each tile as it is generated —
    // list scheduling is within one tile of the mean whatever the order, and one tile is < 1e-4
    // of a rank's load there — so no global tile list is built and nothing is sorted.
    const bool small_plan = n < 4096;
    std::vector<int64_t> load(world, 0);
    auto consume = [&](const bgca_tile &T) {
        int owner = 0;
        for (int r = 1; r < world; ++r)
            if (load[r] < load[owner]) owner = r;
        load[owner] += T.cells;
        int64_t pr, ce;
        tile_pairs(T, pr, ce);
        R.info.total_pairs += pr;
        R.info.total_cells += ce;
        if (owner != rank) return;
        R.info.n_pairs += pr;
        R.info.cells += ce;
        R.info.cells_computed += T.cells;
        if (T.variant == 0 || (T.variant & BGCA_VARIANT_W32)) R.info.n_fallback32++;
        R.tiles.push_back(T);
    };
    std::vector<bgca_tile> all;
    auto produce = [&](const bgca_tile &T) {
        if (small_plan) all.push_back(T);
        else consume(T);
    };
    if (!small_plan) R.tiles.reserve(((size_t)n * ((size_t)n / (2 * 128 * (size_t)spg) + 2) / 2) / (size_t)world + 64);

    bool counting = false;          // pre-pass: only add up the subjects visited (in all, and per variant)
    int64_t visits = 0;
    std::map<int, int64_t> visits_of_variant;
    std::map<int, int> spg_of_variant;     // multi-rank plans: shorter tiles for runs with few of them
    auto self_tile = [&](int qa, int qb) {
        if (counting) return;
        const int La = len[qa], Lb = (qb >= 0) ? len[qb] : 0;
        const int Lq = std::max(La, Lb);
        bgca_tile T{};
        T.qa = qa; T.qb = qb; T.s_begin = qa; T.s_end = (qb >= 0 ? qb : qa) + 1;
        T.variant = pick_variant(Lq, 1);
        T.reserved = 1;
        if (T.variant && !fits_int16(mode, max_abs_sub, open, extend, Lq, Lq)) T.variant = 0;   // two pairs: K4
        T.cells = (int64_t)(La + Lb) * (pre[T.s_end] - pre[T.s_begin]);
        produce(T);
    };
    // query pair (qa, qb) against the subjects [sub0, te), cut into chunks
    auto row_tiles = [&](int qa, int qb, int sub0, int te) {
        const int La = len[qa], Lb = (qb >= 0) ? len[qb] : 0;
        const int Lq = std::max(La, Lb);
        const int S = te - sub0;
        if (S <= 0) return;
        const int variant = pick_variant(Lq, S);
        if (counting) { visits += S; visits_of_variant[variant] += S; return; }
        const int G = (variant >> 8) & 0xFF;
        // does any subject of the row take the scores out of int16?  Then its tiles may end up on
        // the 32-bit form of the multi-pass kernel and are cut for that kernel's 8 groups
        const bool fits_all = variant && fits_int16(mode, max_abs_sub, open, extend, Lq, max_len_in(sub0, te));
        // subjects per tile: ~spg per alignment group (= subject stream) of the CTA,
        // split evenly over the row and rounded up to a whole number per group
        const int ngroups = !variant ? 16 : fits_all ? CTA_THREADS / G : std::min(CTA_THREADS / G, CTA_THREADS / 32);
        int spg_row = spg;
        {
            auto it = spg_of_variant.find(variant);
            if (it != spg_of_variant.end()) spg_row = it->second;
        }
        if ((variant & BGCA_VARIANT_MP) || !fits_all)   // bounds the line of parked rows a group keeps between passes
            spg_row = std::max(1, std::min(spg_row, kMpTokensPerGroup / (max_len_in(sub0, te) + 1)));
        const int s_max = ngroups * spg_row;
        const int nchunks = (S + s_max - 1) / s_max;
        int chunk = (S + nchunks - 1) / nchunks;
        chunk = (chunk + ngroups - 1) / ngroups * ngroups;
        for (int s0 = sub0; s0 < te; s0 += chunk) {
            const int s1 = std::min(te, s0 + chunk);
            if (variant && (fits_all || fits_int16(mode, max_abs_sub, open, extend, Lq, max_len_in(s0, s1)))) {
                bgca_tile T{};
                T.qa = qa; T.qb = qb; T.s_begin = s0; T.s_end = s1;
                T.variant = variant;
                T.cells = (int64_t)(La + Lb) * (pre[s1] - pre[s0]);
                produce(T);
                continue;
            }
            // scores beyond int16 (or a query no packed variant covers): one tile per QUERY on the
            // 32-bit form of the multi-pass kernel; very short queries on the pair-list kernel (K4)
            for (int h = 0; h < 2; ++h) {
                const int q = h ? qb : qa;
                if (q < 0) continue;
                // square plans: the pair (qb, qa) belongs to qa's tile
                const int s0q = (h && !cross_only) ? std::max(s0, qb) : s0;
                if (s0q >= s1) continue;
                bgca_tile T{};
                T.qa = q; T.qb = -1; T.s_begin = s0q; T.s_end = s1;
                T.variant = (len[q] >= BGCA_W32_MIN_ROWS) ? w32_variant(len[q]) : 0;
                T.cells = (int64_t)len[q] * (pre[s1] - pre[s0q]);
                produce(T);
            }
        }
    };

    auto generate = [&]() {
    int tb = 0;
    if (cross_only && n_db_first > 0) {
        // persistent database in front, queries behind: every query type block meets the
        // database's run of the same type (or the whole database)
        const int nd = n_db_first;
        for (int q0 = nd; q0 < n;) {
            int q1 = n, d0 = 0, d1 = nd;
            if (same_type_only && type) {
                q1 = q0 + 1;
                while (q1 < n && type[q1] == type[q0]) ++q1;
                d0 = (int)(std::lower_bound(type, type + nd, type[q0]) - type);
                d1 = (int)(std::upper_bound(type, type + nd, type[q0]) - type);
            }
            for (int qa = q0; qa < q1; qa += 2) {
                const int qb = (qa + 1 < q1) ? qa + 1 : -1;
                row_tiles(qa, qb, d0, d1);
                self_tile(qa, qb);
            }
            q0 = q1;
        }
        if (cross_only == 1)                 // the database's own self scores too
            for (int d = 0; d < nd;) {
                int de = nd;
                if (same_type_only && type) {
                    de = d + 1;
                    while (de < nd && type[de] == type[d]) ++de;
                }
                for (int qa = d; qa < de; qa += 2) self_tile(qa, (qa + 1 < de) ? qa + 1 : -1);
                d = de;
            }
        tb = n;                              // the generic loop below has nothing left to do
    }
    while (tb < n) {
        int te = n;
        if (same_type_only && type) {
            te = tb + 1;
            while (te < n && type[te] == type[tb]) ++te;
        }
        // queries [tb, tm) and subjects' range; square mode: tm = te and subjects start at qa
        int tm = te;
        if (cross_only) {
            tm = tb;
            while (tm < te && role[tm] == 0) ++tm;
        }
        for (int qa = tb; qa < tm; qa += 2) {
            const int qb = (qa + 1 < tm) ? qa + 1 : -1;
            row_tiles(qa, qb, cross_only ? tm : qa, te);
        }
        if (cross_only) {
            // self-only tiles for every sequence of the block (queries and database alike)
            for (int r0 = tb; r0 < te;) {
                const int r_end = (r0 < tm) ? tm : te;          // pairs never straddle the role boundary
                for (int qa = r0; qa < r_end; qa += 2) self_tile(qa, (qa + 1 < r_end) ? qa + 1 : -1);
                r0 = r_end;
            }
        }
        tb = te;
    }
    };   // generate
    if (cross_only || small_plan) {
        // a few queries against a large database, or a small batch, would otherwise make a few
        // very long tiles: size them so that every CTA slot of every rank still gets >= 16
        counting = true;
        generate();
        counting = false;
        const int64_t target_tiles = 16LL * 2 * 148 * world;
        spg = (int)std::max<int64_t>(2, std::min<int64_t>(spg, visits / (target_tiles * 16)));
    } else if (world > 1) {
        // Large plan shared by several ranks: a kernel run (one variant) that leaves a rank fewer
        // than ~8 tiles per CTA slot ends in a tail of half-idle SMs — with many variants (mixed
        // domain lengths) that was 1.3 % of the step at 8 GPUs.  Such runs get shorter tiles.
        counting = true;
        generate();
        counting = false;
        const int64_t want_tiles = 8LL * 2 * 148 * world;
        for (const auto &kv : visits_of_variant) {
            const int G = (kv.first >> 8) & 0xFF;
            const int ngroups = kv.first ? CTA_THREADS / std::max(G, 1) : 16;
            const int64_t s = kv.second / (want_tiles * ngroups);
            if (s < spg) spg_of_variant[kv.first] = (int)std::max<int64_t>(4, s);
        }
    }
    generate();

    if (small_plan) {
        std::vector<int> idx(all.size());
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return all[a].cells > all[b].cells; });
        for (int i : idx) consume(all[i]);
        // one contiguous run per kernel variant, larger variants first, heaviest tile first
        std::stable_sort(R.tiles.begin(), R.tiles.end(), [](const bgca_tile &a, const bgca_tile &b) {
            if (a.variant != b.variant) return a.variant > b.variant;
            return a.cells > b.cells;
        });
    } else {
        // one contiguous run per kernel variant (bgca_score launches run by run), larger variants
        // first, generation order inside a run: a stable bucket pass over the few distinct variants
        std::vector<int> keys;
        for (const auto &T : R.tiles)
            if (std::find(keys.begin(), keys.end(), T.variant) == keys.end()) keys.push_back(T.variant);
        std::sort(keys.begin(), keys.end(), std::greater<int>());
        std::vector<size_t> start(keys.size() + 1, 0);
        auto bucket = [&](int v) { return (size_t)(std::find(keys.begin(), keys.end(), v) - keys.begin()); };
        for (const auto &T : R.tiles) start[bucket(T.variant) + 1]++;
        for (size_t k = 0; k < keys.size(); ++k) start[k + 1] += start[k];
        std::vector<bgca_tile> grouped(R.tiles.size());
        for (const auto &T : R.tiles) grouped[start[bucket(T.variant)]++] = T;
        R.tiles.swap(grouped);
    }
    R.info.n_tiles = (int64_t)R.tiles.size();
    int nv = 0, last = -1;
    for (const auto &T : R.tiles)
        if (T.variant != last) { ++nv; last = T.variant; }
    R.info.n_variants = nv;
    return R;
}

// Where a buffer of `bytes` bytes may start inside an allocation that begins at `raw` so that it
// does not cross a 4 GiB address boundary: `raw` itself when [raw, raw + bytes) lies inside one
// window, else the next boundary (the caller then needs raw + 2 * bytes of room: the boundary is
// less than `bytes` away).  0: no such start (bytes > 4 GiB).
uintptr_t window4g_start(uintptr_t raw, size_t bytes)
{
    if (bytes == 0) return raw;
    if (bytes > ((size_t)1 << 32)) return 0;
    if ((raw >> 32) == ((raw + bytes - 1) >> 32)) return raw;
    return ((raw >> 32) + 1) << 32;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    void *base = nullptr;       // what cudaMalloc returned (p may sit further in: see alloc)
    // window4g: the buffer must not cross a 4 GiB address boundary — the DP kernels advance their
    // token pointer with a 32-bit add on its low half (gotoh_pair16.cuh, advance_lo32).  An
    // allocation that crosses one is made again at twice the size and used from the boundary on.
    static cudaError_t alloc(size_t bytes, bool window4g, void **base_out, T **p_out)
    {
        void *raw = nullptr;
        bytes = std::max<size_t>(bytes, 1);
        cudaError_t e = cudaMalloc(&raw, bytes);
        if (e != cudaSuccess) return e;
        T *use = static_cast<T *>(raw);
        if (window4g && window4g_start((uintptr_t)raw, bytes) != (uintptr_t)raw) {
            cudaFree(raw);
            if (bytes > ((size_t)1 << 32)) return cudaErrorInvalidValue;
            e = cudaMalloc(&raw, 2 * bytes);
            if (e != cudaSuccess) return e;
            use = reinterpret_cast<T *>(window4g_start((uintptr_t)raw, bytes));
        }
        *base_out = raw;
        *p_out = use;
        return cudaSuccess;
    }
    cudaError_t reserve(size_t count, bool window4g = false)
    {
        if (count <= cap && p) return cudaSuccess;
        release();
        cudaError_t e = alloc(count * sizeof(T), window4g, &base, &p);
        if (e == cudaSuccess) cap = count;
        return e;
    }
    // grow to `count` elements, keeping the first `keep` (device-to-device copy on `st`)
    cudaError_t grow_keep(size_t count, size_t keep, cudaStream_t st, bool window4g = false)
    {
        if (count <= cap && p) return cudaSuccess;
        T *np_ = nullptr;
        void *nbase = nullptr;
        const size_t want = count + count / 4;          // headroom: query sets vary from call to call
        cudaError_t e = alloc(want * sizeof(T), window4g, &nbase, &np_);
        if (e != cudaSuccess) return e;
        if (p && keep) {
            e = cudaMemcpyAsync(np_, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(nbase); return e; }
        }
        if (base) cudaFree(base);
        p = np_;
        base = nbase;
        cap = want;
        return cudaSuccess;
    }
    void release()
    {
        if (base) cudaFree(base);
        p = nullptr;
        base = nullptr;
        cap = 0;
    }
};

}  // namespace

struct bgca_engine {
    int device = 0;
    int n_sm = 0;
    // scoring
    int8_t sub[NSYM * NSYM];
    int open = -12, extend = -1, mode = BGCA_MODE_LOCAL, max_abs_sub = 11;
    bool scoring_set = false;
    // packed batch (host mirrors, SORTED space unless named *_orig)
    int32_t n = 0, n_bgc = 0;
    std::vector<int32_t> len, type, bgc, role, order, inv_order;
    int32_t n_query = 0;          // sequences with role 0
    // persistent database: sequences [0, n_db) of BOTH index spaces are a database that stays
    // packed while query sets are appended behind it (bgca_pack_append); 0 = none
    int32_t n_db = 0;
    uint64_t db_bytes = 0;        // packed bytes of the database part
    bool has_role = false, cross_planned = false;
    std::vector<uint32_t> seq_off;
    std::vector<int32_t> bgc_start, bgc_doms;   // CSR over ORIGINAL indices
    int64_t packed_bytes = 0;
    bool packed = false, has_type = false, has_bgc = false;
    // device buffers
    DevBuf<uint8_t> d_ascii, d_codes;
    DevBuf<int64_t> d_offsets_orig;
    DevBuf<int32_t> d_order, d_len, d_type, d_bgc, d_bgc_start, d_bgc_doms, d_counters, d_pq, d_ps;
    DevBuf<uint32_t> d_seq_off;
    DevBuf<unsigned long long> d_err;
    DevBuf<bgca_tile> d_tiles;
    DevBuf<int2> d_scratch, d_mp_scratch;
    DevBuf<int16_t> d_s16;                // sorted-space staging of the score matrix (bgca_score)
    DevBuf<uint4> d_scratch_st;
    DevBuf<int32_t> d_scores_host_call;   // bgca_score_all_host's n x n staging matrix
    DevBuf<int32_t> d_st2_items;          // K7 work items (Stats2Item = 4 x int32)
    DevBuf<int32_t> d_st2_hint;           // K7: known best scores of the grouped pairs
    // pair selection by score (bgca_select_pairs -> bgca_pair_stats_selected)
    DevBuf<int32_t> d_inv_order, d_sel_cnt, d_sel_cls;
    DevBuf<int64_t> d_sel_rowstart;
    bool sel_valid = false;
    int32_t sel_min_score = 0, sel_same_type = 0;
    int64_t sel_total = 0;
    int32_t sel_cls_items[16] = {0};
    // plan
    std::vector<bgca_tile> tiles;
    bgca_plan_info plan_info{};
    bool planned = false;
    // stats
    int last_launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    cudaEvent_t evp0 = nullptr, evp1 = nullptr;   // around the K1 kernel of the last bgca_pack
    bool pack_timed = false;
    // side streams: the per-variant DP kernels are persistent, so launching them on
    // different streams lets the tail of one overlap the head of the next
    static constexpr int kSideStreams = 4;
    cudaStream_t side[kSideStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kSideStreams] = {nullptr, nullptr, nullptr, nullptr};
    // pipelined host call (bgca_score_all_host): finished row blocks of the sorted-space int16
    // matrix travel to the host on the copy engine while the remaining kernel runs compute
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_pool;     // per-run / per-block events, grown on demand
    int16_t *h_stage = nullptr;           // pinned copy of the staging matrix (SORTED space, int16)
    size_t h_stage_elems = 0;
};

// where bgca_score_all_host wants the matrix: straight into host memory, block by block
struct HostSink {
    int32_t *scores_host;
    bool used = false;                    // set by score_impl when it took the pipelined path
};

extern "C" {

int bgca_version(void) { return BGCA_VERSION; }

const char *bgca_last_error(void) { return g_last_error.c_str(); }

int bgca_window4g_start(uint64_t raw, uint64_t bytes, uint64_t *start)
{
    if (!start) return fail(BGCA_ERR_INVALID, "bgca_window4g_start: start is NULL");
    *start = (uint64_t)window4g_start((uintptr_t)raw, (size_t)bytes);
    return *start || !bytes ? BGCA_OK : fail(BGCA_ERR_INVALID, "bgca_window4g_start: more than 4 GiB");
}

int bgca_create(int device, bgca_engine **out)
try {
    if (!out) return fail(BGCA_ERR_INVALID, "bgca_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BGCA_ERR_CUDA, std::string("bgca_create: no usable CUDA device (") +
                                       cudaGetErrorString(e) + "); this engine has no CPU fallback");
    if (device < 0 || device >= count) return fail(BGCA_ERR_INVALID, "bgca_create: bad device index");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(BGCA_ERR_CUDA, "bgca_create: device is not sm_100 (B200) class");
    std::unique_ptr<bgca_engine, int (*)(bgca_engine *)> guard(new bgca_engine(), bgca_destroy);
    bgca_engine *E = guard.get();
    E->device = device;
    E->n_sm = prop.multiProcessorCount;
    CUDA_TRY(cudaEventCreate(&E->ev0));
    CUDA_TRY(cudaEventCreate(&E->ev1));
    CUDA_TRY(cudaEventCreate(&E->evp0));
    CUDA_TRY(cudaEventCreate(&E->evp1));
    CUDA_TRY(cudaEventCreateWithFlags(&E->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < bgca_engine::kSideStreams; ++i) {
        CUDA_TRY(cudaStreamCreateWithFlags(&E->side[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&E->ev_join[i], cudaEventDisableTiming));
    }
    {
        // highest priority: the DP kernels are persistent, so a row-block permutation can only start
        // in the slots a finishing run frees — it has to win them against the next run's CTAs
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&E->copy_stream, cudaStreamNonBlocking, prio_hi));
    }
    *out = guard.release();
    return BGCA_OK;
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_create: ") + ex.what());
}

int bgca_destroy(bgca_engine *e)
{
    if (!e) return BGCA_OK;
    cudaSetDevice(e->device);
    e->d_ascii.release(); e->d_codes.release(); e->d_offsets_orig.release(); e->d_order.release();
    e->d_len.release(); e->d_type.release(); e->d_bgc.release(); e->d_bgc_start.release();
    e->d_bgc_doms.release(); e->d_counters.release(); e->d_pq.release(); e->d_ps.release();
    e->d_seq_off.release(); e->d_err.release(); e->d_tiles.release(); e->d_scratch.release(); e->d_mp_scratch.release(); e->d_s16.release(); e->d_scratch_st.release(); e->d_scores_host_call.release(); e->d_st2_items.release(); e->d_st2_hint.release();
    e->d_inv_order.release(); e->d_sel_cnt.release(); e->d_sel_cls.release(); e->d_sel_rowstart.release();
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->evp0) cudaEventDestroy(e->evp0);
    if (e->evp1) cudaEventDestroy(e->evp1);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (int i = 0; i < bgca_engine::kSideStreams; ++i) {
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
        if (e->side[i]) cudaStreamDestroy(e->side[i]);
    }
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    delete e;
    return BGCA_OK;
}

int bgca_set_scoring(bgca_engine *e, const int8_t *sub, int32_t open, int32_t extend, int32_t mode)
{
    if (!e || !sub) return fail(BGCA_ERR_INVALID, "bgca_set_scoring: NULL argument");
    if (mode != BGCA_MODE_LOCAL && mode != BGCA_MODE_GLOBAL)
        return fail(BGCA_ERR_INVALID, "bgca_set_scoring: mode must be 0 (local) or 1 (global)");
    if (open > 0 || extend > 0 || open > extend)
        return fail(BGCA_ERR_INVALID, "bgca_set_scoring: need open <= extend <= 0 (Gotoh H/E/F form)");
    if (open < -1000 || extend < -1000) return fail(BGCA_ERR_RANGE, "bgca_set_scoring: gap score too large");
    int mx = 0;
    for (int i = 0; i < NSYM; ++i)
        for (int j = 0; j < NSYM; ++j) {
            if (sub[i * NSYM + j] != sub[j * NSYM + i])
                return fail(BGCA_ERR_INVALID, "bgca_set_scoring: substitution matrix must be symmetric");
            mx = std::max(mx, std::abs((int)sub[i * NSYM + j]));
        }
    std::memcpy(e->sub, sub, NSYM * NSYM);
    e->open = open; e->extend = extend; e->mode = mode; e->max_abs_sub = std::max(mx, 1);
    e->scoring_set = true;
    e->planned = false;
    return BGCA_OK;
}

// Shared body of bgca_pack / bgca_pack_spans: sequence i is ascii[src_off[i], src_off[i] + src_len[i])
// of a buffer of ascii_bytes bytes (host or device).  Sequences may overlap or leave gaps — domains
// are slices of their proteins' sequences.
static int pack_impl(bgca_engine *e, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *src_off,
                     const int64_t *src_len, int32_t n, const int32_t *type_of_seq, const int32_t *bgc_of_seq,
                     int32_t n_bgc, int32_t n_query, int32_t ascii_on_device, void *stream)
{
    if (n <= 0) return fail(BGCA_ERR_INVALID, "bgca_pack: need at least one sequence");
    if (bgc_of_seq && (n_bgc <= 0 || n_bgc > 65535))
        return fail(BGCA_ERR_RANGE, "bgca_pack: n_bgc must be in 1..65535");
    if (n_query < 0 || n_query > n) return fail(BGCA_ERR_INVALID, "bgca_pack: n_query must be in 0..n");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    e->packed = false;
    e->planned = false;
    e->n_db = 0;

    std::vector<int32_t> len_orig(n);
    for (int i = 0; i < n; ++i) {
        const int64_t l = src_len[i];
        if (l < 0) return fail(BGCA_ERR_INVALID, "bgca_pack: offsets must be non-decreasing");
        if (src_off[i] < 0 || src_off[i] + l > ascii_bytes)
            return fail(BGCA_ERR_INVALID, "sequence " + std::to_string(i) + " lies outside the residue buffer");
        if (l == 0) return fail(BGCA_ERR_EMPTY, "sequence " + std::to_string(i) + " has zero length");
        if (l > 65535) return fail(BGCA_ERR_RANGE, "sequence " + std::to_string(i) + " longer than 65535");
        len_orig[i] = (int32_t)l;
    }
    if (bgc_of_seq)
        for (int i = 0; i < n; ++i)
            if (bgc_of_seq[i] < 0 || bgc_of_seq[i] >= n_bgc)
                return fail(BGCA_ERR_INVALID, "bgca_pack: bgc_of_seq out of range");

    // sort by (type, length, original index)
    e->order.resize(n);
    std::iota(e->order.begin(), e->order.end(), 0);
    std::stable_sort(e->order.begin(), e->order.end(), [&](int a, int b) {
        const int ta = type_of_seq ? type_of_seq[a] : 0, tb = type_of_seq ? type_of_seq[b] : 0;
        if (ta != tb) return ta < tb;
        if (n_query > 0) {                       // queries (original index < n_query) before database
            const int ra = a >= n_query, rb = b >= n_query;
            if (ra != rb) return ra < rb;
        }
        return len_orig[a] < len_orig[b];
    });
    e->n = n;
    e->n_bgc = bgc_of_seq ? n_bgc : 0;
    e->has_type = type_of_seq != nullptr;
    e->has_bgc = bgc_of_seq != nullptr;
    e->n_query = n_query;
    e->has_role = n_query > 0;
    e->cross_planned = false;
    e->len.resize(n); e->type.assign(n, 0); e->bgc.assign(n, 0); e->role.assign(n, 0); e->seq_off.resize(n);
    e->inv_order.resize(n);
    uint64_t off = kPackedHeader;   // header of idle tokens (see gotoh_pair16.cuh)
    for (int r = 0; r < n; ++r) {
        const int o = e->order[r];
        e->inv_order[o] = r;
        e->len[r] = len_orig[o];
        if (type_of_seq) e->type[r] = type_of_seq[o];
        if (bgc_of_seq) e->bgc[r] = bgc_of_seq[o];
        e->role[r] = (n_query > 0 && o >= n_query) ? 1 : 0;
        e->seq_off[r] = (uint32_t)off;
        off += (uint64_t)((len_orig[o] + 1 + 15) & ~15);   // >= 1 pad byte = end-of-subject token
        if (off > 0xFFFFFFF0ull) return fail(BGCA_ERR_RANGE, "bgca_pack: packed batch exceeds 4 GiB");
    }
    e->packed_bytes = (int64_t)off;

    // CSR of domains per BGC, ORIGINAL indices
    e->bgc_start.assign((size_t)e->n_bgc + 1, 0);
    e->bgc_doms.assign(bgc_of_seq ? n : 0, 0);
    if (bgc_of_seq) {
        for (int i = 0; i < n; ++i) e->bgc_start[bgc_of_seq[i] + 1]++;
        for (int b = 0; b < n_bgc; ++b) e->bgc_start[b + 1] += e->bgc_start[b];
        std::vector<int32_t> cur(e->bgc_start.begin(), e->bgc_start.end() - 1);
        for (int i = 0; i < n; ++i) e->bgc_doms[cur[bgc_of_seq[i]]++] = i;
    }

    const uint8_t *d_src = ascii;
    if (!ascii_on_device) {
        CUDA_TRY(e->d_ascii.reserve((size_t)ascii_bytes + 16));   // K1 reads whole aligned 4-byte words
        if (ascii_bytes > 0)
            CUDA_TRY(cudaMemcpyAsync(e->d_ascii.p, ascii, (size_t)ascii_bytes, cudaMemcpyHostToDevice, st));
        d_src = e->d_ascii.p;
    }
    // source byte offset of every SORTED sequence (relative to d_src): one gather less on the device
    std::vector<int64_t> off_rel(n);
    for (int r = 0; r < n; ++r) off_rel[r] = src_off[e->order[r]];
    CUDA_TRY(e->d_offsets_orig.reserve(n));
    CUDA_TRY(e->d_order.reserve(n));
    CUDA_TRY(e->d_len.reserve(n));
    CUDA_TRY(e->d_type.reserve(n));
    CUDA_TRY(e->d_bgc.reserve(n));
    CUDA_TRY(e->d_seq_off.reserve(n));
    CUDA_TRY(e->d_codes.reserve((size_t)off + 16, /*window4g=*/true));
    CUDA_TRY(e->d_err.reserve(1));
    CUDA_TRY(e->d_bgc_start.reserve(e->bgc_start.size()));
    CUDA_TRY(e->d_bgc_doms.reserve(std::max<size_t>(e->bgc_doms.size(), 1)));
    CUDA_TRY(cudaMemcpyAsync(e->d_offsets_orig.p, off_rel.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_order.p, e->order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_len.p, e->len.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_type.p, e->type.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_bgc.p, e->bgc.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_seq_off.p, e->seq_off.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_bgc_start.p, e->bgc_start.data(), sizeof(int32_t) * e->bgc_start.size(), cudaMemcpyHostToDevice, st));
    if (!e->bgc_doms.empty())
        CUDA_TRY(cudaMemcpyAsync(e->d_bgc_doms.p, e->bgc_doms.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(e->d_inv_order.reserve(n));
    CUDA_TRY(cudaMemcpyAsync(e->d_inv_order.p, e->inv_order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(e->d_err.p, 0xFF, sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(e->d_codes.p, 25, kPackedHeader, st));
    CUDA_TRY(cudaMemsetAsync(e->d_codes.p + off, BGCA_PAD_CODE, 16, st));
    CUDA_TRY(cudaEventRecord(e->evp0, st));
    CUDA_TRY(launch_pack(d_src, e->d_offsets_orig.p, e->d_order.p, e->d_seq_off.p, e->d_len.p, n,
                         e->d_codes.p, e->d_err.p, st));
    CUDA_TRY(cudaEventRecord(e->evp1, st));
    e->pack_timed = true;
    unsigned long long err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, e->d_err.p, sizeof(err), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (err != ~0ULL) {
        const unsigned seq = (unsigned)(err >> 32), pos = (unsigned)(err & 0xFFFFFFFFu);
        return fail(BGCA_ERR_ALPHABET, "sequence " + std::to_string(seq) + " contains a letter not in the "
                                       "alphabet ARNDCQEGHILKMFPSTWYVBZX* at position " + std::to_string(pos));
    }
    e->packed = true;
    return BGCA_OK;
}

int bgca_pack(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n,
              const int32_t *type_of_seq, const int32_t *bgc_of_seq, int32_t n_bgc,
              int32_t n_query, int32_t ascii_on_device, void *stream)
try {
    if (!e || !ascii || !offsets) return fail(BGCA_ERR_INVALID, "bgca_pack: NULL argument");
    if (n <= 0) return fail(BGCA_ERR_INVALID, "bgca_pack: need at least one sequence");
    std::vector<int64_t> off(n), len(n);
    for (int i = 0; i < n; ++i) {
        off[i] = offsets[i] - offsets[0];
        len[i] = offsets[i + 1] - offsets[i];
    }
    return pack_impl(e, ascii + offsets[0], offsets[n] - offsets[0], off.data(), len.data(), n, type_of_seq,
                     bgc_of_seq, n_bgc, n_query, ascii_on_device, stream);
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_pack: ") + ex.what());
}

int bgca_pack_spans(bgca_engine *e, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *span_start,
                    const int32_t *span_len, int32_t n, const int32_t *type_of_seq, const int32_t *bgc_of_seq,
                    int32_t n_bgc, int32_t n_query, int32_t ascii_on_device, void *stream)
try {
    if (!e || !ascii || !span_start || !span_len || ascii_bytes < 0)
        return fail(BGCA_ERR_INVALID, "bgca_pack_spans: NULL argument");
    if (n <= 0) return fail(BGCA_ERR_INVALID, "bgca_pack_spans: need at least one sequence");
    std::vector<int64_t> len(n);
    for (int i = 0; i < n; ++i) len[i] = span_len[i];
    return pack_impl(e, ascii, ascii_bytes, span_start, len.data(), n, type_of_seq, bgc_of_seq, n_bgc, n_query,
                     ascii_on_device, stream);
} catch (const std::exception &ex) {
    return fail(BGCA_ERR_RANGE, std::string("bgca_pack_spans: ") + ex.what());
}

int bgca_pack_append(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n_new,
                     const int32_t *type_of_seq, const int32_t *bgc_of_seq, int32_t n_bgc_total,
                     int32_t ascii_on_device, void *stream)
try {
    if (!e || !ascii || !offsets) return fail(BGCA_ERR_INVALID, "bgca_pack_append: NULL argument");
    if (!e->packed) return fail(BGCA_ERR_INVALID, "bgca_pack_append: pack the database first (bgca_pack, n_query = 0)");
    if (e->n_db == 0 && e->n_query != 0)
        return fail(BGCA_ERR_INVALID, "bgca_pack_append: the packed batch already holds queries (n_query != 0)");
    if (n_new <= 0) return fail(BGCA_ERR_INVALID, "bgca_pack_append: need at least one query sequence");
    if ((type_of_seq != nullptr) != e->has_type || (bgc_of_seq != nullptr) != e->has_bgc)
        return fail(BGCA_ERR_INVALID, "bgca_pack_append: give type_of_seq / bgc_of_seq exactly as for the database");
    if (bgc_of_seq && (n_bgc_total <= 0 || n_bgc_total > 65535))
        return fail(BGCA_ERR_RANGE, "bgca_pack_append: n_bgc_total must be in 1..65535");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    // every host-side check comes before the engine is touched, so a rejected query set leaves
    // the database exactly as it was
    const int n_db = e->n_db ? e->n_db : e->n;
    const uint64_t db_bytes = e->n_db ? e->db_bytes : (uint64_t)e->packed_bytes;
    if ((int64_t)n_db + n_new > 0x7FFFFFF0) return fail(BGCA_ERR_RANGE, "bgca_pack_append: too many sequences");
    std::vector<int32_t> len_new(n_new);
    {
        uint64_t bytes = db_bytes;
        for (int i = 0; i < n_new; ++i) {
            const int64_t l = offsets[i + 1] - offsets[i];
            if (l < 0) return fail(BGCA_ERR_INVALID, "bgca_pack_append: offsets must be non-decreasing");
            if (l == 0) return fail(BGCA_ERR_EMPTY, "sequence " + std::to_string(i) + " has zero length");
            if (l > 65535) return fail(BGCA_ERR_RANGE, "sequence " + std::to_string(i) + " longer than 65535");
            len_new[i] = (int32_t)l;
            bytes += (uint64_t)((l + 1 + 15) & ~15);
        }
        if (bytes > 0xFFFFFFF0ull) return fail(BGCA_ERR_RANGE, "bgca_pack_append: packed batch exceeds 4 GiB");
    }
    if (bgc_of_seq)
        for (int i = 0; i < n_new; ++i)
            if (bgc_of_seq[i] < 0 || bgc_of_seq[i] >= n_bgc_total)
                return fail(BGCA_ERR_INVALID, "bgca_pack_append: bgc_of_seq out of range");
    e->n_db = n_db;                          // first query set: everything packed so far is the database
    e->db_bytes = db_bytes;
    e->packed = false;
    e->planned = false;
    // any failure from here on (allocation, upload, an illegal letter) puts the engine back on
    // the bare database: the next bgca_pack_append starts over
    struct Restore {
        bgca_engine *e;
        bool armed = true;
        ~Restore()
        {
            if (!armed) return;
            e->n = e->n_db; e->n_query = 0; e->packed_bytes = (int64_t)e->db_bytes; e->packed = true;
            e->len.resize(e->n); e->type.resize(e->n); e->bgc.resize(e->n); e->role.resize(e->n);
            e->seq_off.resize(e->n); e->order.resize(e->n); e->inv_order.resize(e->n);
        }
    } restore{e};

    // queries sorted by (type, length) among themselves, placed behind the database in BOTH
    // index spaces: sorted r = n_db + k  <->  original n_db + (position in the query list)
    std::vector<int> qord(n_new);
    std::iota(qord.begin(), qord.end(), 0);
    std::stable_sort(qord.begin(), qord.end(), [&](int a, int b) {
        const int ta = type_of_seq ? type_of_seq[a] : 0, tb = type_of_seq ? type_of_seq[b] : 0;
        if (ta != tb) return ta < tb;
        return len_new[a] < len_new[b];
    });
    const int n = n_db + n_new;
    e->len.resize(n); e->type.resize(n); e->bgc.resize(n); e->role.resize(n); e->seq_off.resize(n);
    e->order.resize(n); e->inv_order.resize(n);
    for (int r = 0; r < n_db; ++r) e->role[r] = 1;
    uint64_t off = e->db_bytes;
    for (int k = 0; k < n_new; ++k) {
        const int r = n_db + k, o = qord[k];
        e->order[r] = n_db + o;
        e->inv_order[n_db + o] = r;
        e->len[r] = len_new[o];
        e->type[r] = type_of_seq ? type_of_seq[o] : 0;
        e->bgc[r] = bgc_of_seq ? bgc_of_seq[o] : 0;
        e->role[r] = 0;
        e->seq_off[r] = (uint32_t)off;
        off += (uint64_t)((len_new[o] + 1 + 15) & ~15);
    }
    e->n = n;
    e->n_query = n_new;
    e->has_role = true;
    e->cross_planned = false;
    e->packed_bytes = (int64_t)off;
    if (bgc_of_seq) {
        e->n_bgc = n_bgc_total;
        e->bgc_start.assign((size_t)e->n_bgc + 1, 0);
        e->bgc_doms.assign(n, 0);
        for (int r = 0; r < n; ++r) e->bgc_start[e->bgc[r] + 1]++;
        for (int b = 0; b < e->n_bgc; ++b) e->bgc_start[b + 1] += e->bgc_start[b];
        std::vector<int32_t> cur(e->bgc_start.begin(), e->bgc_start.end() - 1);
        std::vector<int32_t> bgc_orig(n);
        for (int r = 0; r < n; ++r) bgc_orig[e->order[r]] = e->bgc[r];
        for (int i = 0; i < n; ++i) e->bgc_doms[cur[bgc_orig[i]]++] = i;      // ORIGINAL indices, ascending per BGC
    }

    const int64_t total_ascii = offsets[n_new] - offsets[0];
    const uint8_t *d_src = ascii + offsets[0];
    if (!ascii_on_device) {
        CUDA_TRY(e->d_ascii.reserve((size_t)total_ascii + 16));
        CUDA_TRY(cudaMemcpyAsync(e->d_ascii.p, ascii + offsets[0], (size_t)total_ascii, cudaMemcpyHostToDevice, st));
        d_src = e->d_ascii.p;
    }
    std::vector<int64_t> off_rel(n_new);
    for (int k = 0; k < n_new; ++k) off_rel[k] = offsets[qord[k]] - offsets[0];
    CUDA_TRY(e->d_codes.grow_keep((size_t)off + 16, (size_t)e->db_bytes, st, /*window4g=*/true));
    CUDA_TRY(e->d_offsets_orig.reserve(n_new));
    CUDA_TRY(e->d_order.reserve(n));
    CUDA_TRY(e->d_len.reserve(n));
    CUDA_TRY(e->d_type.reserve(n));
    CUDA_TRY(e->d_bgc.reserve(n));
    CUDA_TRY(e->d_seq_off.reserve(n));
    CUDA_TRY(e->d_err.reserve(1));
    CUDA_TRY(e->d_bgc_start.reserve(e->bgc_start.size()));
    CUDA_TRY(e->d_bgc_doms.reserve(std::max<size_t>(e->bgc_doms.size(), 1)));
    CUDA_TRY(cudaMemcpyAsync(e->d_offsets_orig.p, off_rel.data(), sizeof(int64_t) * n_new, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_order.p, e->order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_len.p, e->len.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_type.p, e->type.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_bgc.p, e->bgc.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_seq_off.p, e->seq_off.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    if (bgc_of_seq) {
        CUDA_TRY(cudaMemcpyAsync(e->d_bgc_start.p, e->bgc_start.data(), sizeof(int32_t) * e->bgc_start.size(), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->d_bgc_doms.p, e->bgc_doms.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemsetAsync(e->d_err.p, 0xFF, sizeof(unsigned long long), st));
    CUDA_TRY(cudaMemsetAsync(e->d_codes.p + off, BGCA_PAD_CODE, 16, st));
    CUDA_TRY(cudaEventRecord(e->evp0, st));
    CUDA_TRY(launch_pack(d_src, e->d_offsets_orig.p, e->d_order.p + n_db, e->d_seq_off.p + n_db, e->d_len.p + n_db,
                         n_new, e->d_codes.p, e->d_err.p, st));
    CUDA_TRY(cudaEventRecord(e->evp1, st));
    e->pack_timed = true;
    unsigned long long err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, e->d_err.p, sizeof(err), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (err != ~0ULL) {
        const unsigned seq = (unsigned)(err >> 32) - (unsigned)n_db, pos = (unsigned)(err & 0xFFFFFFFFu);
        return fail(BGCA_ERR_ALPHABET, "sequence " + std::to_string(seq) + " contains a letter not in the "
                                       "alphabet ARNDCQEGHILKMFPSTWYVBZX* at position " + std::to_string(pos));
    }
    restore.armed = false;
    e->packed = true;
    return BGCA_OK;
} catch (const std::exception &ex) {
    return fail(BGCA_ERR_RANGE, std::string("bgca_pack_append: ") + ex.what());
}

int64_t bgca_packed_bytes(const bgca_engine *e) { return (e && e->packed) ? e->packed_bytes : 0; }

int bgca_get_packed(const bgca_engine *e, uint8_t *codes_out, int64_t *offsets_sorted_out,
                    int32_t *order_out, void *stream)
try {
    if (!e || !e->packed) return fail(BGCA_ERR_INVALID, "bgca_get_packed: nothing packed");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    if (codes_out) {
        CUDA_TRY(cudaMemcpyAsync(codes_out, e->d_codes.p, (size_t)e->packed_bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (offsets_sorted_out) {
        for (int r = 0; r < e->n; ++r) offsets_sorted_out[r] = e->seq_off[r];
        offsets_sorted_out[e->n] = e->packed_bytes;
    }
    if (order_out) std::memcpy(order_out, e->order.data(), sizeof(int32_t) * e->n);
    return BGCA_OK;
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_get_packed: ") + ex.what());
}

int bgca_plan_host(const int32_t *len_sorted, const int32_t *type_sorted, const int32_t *role_sorted,
                   int32_t n, int32_t same_type_only, int32_t cross_only, int32_t n_db_first, int32_t mode,
                   int32_t max_abs_sub, int32_t open, int32_t extend, int32_t rank, int32_t world,
                   bgca_tile *tiles_out, int64_t capacity, int64_t *n_tiles_out, bgca_plan_info *info)
try {
    if (!len_sorted || n <= 0 || world <= 0 || rank < 0 || rank >= world)
        return fail(BGCA_ERR_INVALID, "bgca_plan_host: bad arguments");
    if (n_db_first < 0 || n_db_first >= n || (n_db_first > 0 && !cross_only))
        return fail(BGCA_ERR_INVALID, "bgca_plan_host: n_db_first needs a cross plan and 0 <= n_db_first < n");
    if (cross_only && n_db_first == 0 && !role_sorted)
        return fail(BGCA_ERR_INVALID, "bgca_plan_host: cross_only needs roles");
    if (cross_only && n_db_first == 0 && !roles_partitioned(type_sorted, role_sorted, n, same_type_only))
        return fail(BGCA_ERR_INVALID, "bgca_plan_host: cross_only needs every planned block sorted queries-first");
    PlanResult R = make_plan(len_sorted, type_sorted, role_sorted, n, same_type_only, cross_only, n_db_first,
                             mode, max_abs_sub, open, extend, rank, world);
    if (n_tiles_out) *n_tiles_out = (int64_t)R.tiles.size();
    if (info) *info = R.info;
    if (tiles_out) {
        const int64_t m = std::min<int64_t>(capacity, (int64_t)R.tiles.size());
        std::memcpy(tiles_out, R.tiles.data(), sizeof(bgca_tile) * (size_t)m);
    }
    return BGCA_OK;
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_plan_host: ") + ex.what());
}

int bgca_plan(bgca_engine *e, int32_t same_type_only, int32_t cross_only, int32_t rank, int32_t world,
              bgca_plan_info *info)
try {
    if (!e || !e->packed) return fail(BGCA_ERR_INVALID, "bgca_plan: call bgca_pack first");
    if (!e->scoring_set) return fail(BGCA_ERR_INVALID, "bgca_plan: call bgca_set_scoring first");
    if (world <= 0 || rank < 0 || rank >= world) return fail(BGCA_ERR_INVALID, "bgca_plan: bad rank/world");
    if (same_type_only && !e->has_type)
        return fail(BGCA_ERR_INVALID, "bgca_plan: same_type_only needs type_of_seq at pack time");
    if (cross_only && (!e->has_role || e->n_query >= e->n))
        return fail(BGCA_ERR_INVALID, "bgca_plan: cross_only needs 0 < n_query < n at pack time");
    if (cross_only && e->n_db == 0 &&
        !roles_partitioned(e->has_type ? e->type.data() : nullptr, e->role.data(), e->n, same_type_only))
        return fail(BGCA_ERR_INVALID, "bgca_plan: cross_only without same_type_only needs a batch packed "
                                      "without type_of_seq (the packer sorts by type before role)");
    if (!cross_only && e->n_db > 0)
        return fail(BGCA_ERR_INVALID, "bgca_plan: a database with appended queries takes a cross plan "
                                      "(cross_only = 1, or 2 to leave out the database's self tiles)");
    CUDA_TRY(cudaSetDevice(e->device));
    PlanResult R = make_plan(e->len.data(), e->has_type ? e->type.data() : nullptr, e->role.data(), e->n,
                             same_type_only, cross_only, e->n_db, e->mode, e->max_abs_sub, e->open, e->extend,
                             rank, world);
    e->cross_planned = cross_only != 0;
    e->tiles.swap(R.tiles);
    e->plan_info = R.info;
    // the persistent kernels of an earlier bgca_score() may still be reading d_tiles on the caller's
    // stream or on the side streams: its closing event (recorded after the side streams joined)
    // must have passed before the list is replaced or its buffer reallocated
    if (e->timed) CUDA_TRY(cudaEventSynchronize(e->ev1));
    CUDA_TRY(e->d_tiles.reserve(e->tiles.size()));
    if (!e->tiles.empty())
        CUDA_TRY(cudaMemcpy(e->d_tiles.p, e->tiles.data(), sizeof(bgca_tile) * e->tiles.size(),
                            cudaMemcpyHostToDevice));
    e->planned = true;
    if (info) *info = e->plan_info;
    return BGCA_OK;
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_plan: ") + ex.what());
}

static void fill_params(const bgca_engine *e, ScoreParams &p, int32_t *scores, int32_t *rowbest,
                        int32_t *selfsc, int type_check)
{
    std::memset(&p, 0, sizeof(p));
    p.codes = e->d_codes.p;
    p.seq_off = e->d_seq_off.p;
    p.seq_len = e->d_len.p;
    p.order = e->d_order.p;
    p.type_of = e->d_type.p;
    p.bgc_of = e->d_bgc.p;
    p.scores = scores;
    p.rowbest = rowbest;
    p.selfsc = selfsc;
    p.n = e->n;
    p.n_bgc = e->n_bgc;
    p.open = e->open;
    p.extend = e->extend;
    p.open2 = ((uint32_t)e->open & 0xFFFFu) * 0x10001u;
    p.ext2 = ((uint32_t)e->extend & 0xFFFFu) * 0x10001u;
    p.type_check = type_check;
    p.cross_nq = e->cross_planned ? e->n_query : 0;
    p.cross_q0 = e->cross_planned ? e->n_db : 0;       // queries sit behind a persistent database
    std::memcpy(p.sub, e->sub, NSYM * NSYM);
}

static int run_s32_pairs(bgca_engine *e, ScoreParams &p, const int32_t *d_pq, const int32_t *d_ps,
                         long long npairs, int32_t *out_list, cudaStream_t st)
{
    int max_len = 0;
    for (int r = 0; r < e->n; ++r) max_len = std::max(max_len, e->len[r]);
    int grid = e->n_sm * 4;
    const long long warps_needed = (npairs + 0);
    if ((long long)grid * (CTA_THREADS / 32) > warps_needed)
        grid = (int)std::max<long long>(1, (warps_needed + CTA_THREADS / 32 - 1) / (CTA_THREADS / 32));
    const int stride = (max_len + 3) & ~3;
    CUDA_TRY(e->d_scratch.reserve((size_t)grid * (CTA_THREADS / 32) * stride));
    CUDA_TRY(launch_s32(e->mode, p, d_pq, d_ps, npairs, e->d_scratch.p, stride, out_list, grid, st));
    e->last_launches++;
    return BGCA_OK;
}

static int score_impl(bgca_engine *e, int32_t *scores, int32_t *rowbest, int32_t *selfsc, int32_t type_check,
                      void *stream, HostSink *sink)
{
    if (!e || !e->planned) return fail(BGCA_ERR_INVALID, "bgca_score: call bgca_plan first");
    if (rowbest && !e->has_bgc) return fail(BGCA_ERR_INVALID, "bgca_score: rowbest needs bgc_of_seq at pack time");
    if (rowbest && type_check && !e->has_type)
        return fail(BGCA_ERR_INVALID, "bgca_score: type_check needs type_of_seq at pack time");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    e->last_launches = 0;
    ScoreParams p;
    fill_params(e, p, scores, rowbest, selfsc, type_check);

    // one launch per kernel variant present in the plan (tiles are grouped by variant)
    std::vector<std::pair<size_t, size_t>> runs;   // [begin, end) per variant
    for (size_t i = 0; i < e->tiles.size();) {
        size_t j = i;
        while (j < e->tiles.size() && e->tiles[j].variant == e->tiles[i].variant) ++j;
        runs.push_back({i, j});
        i = j;
    }
    // multi-pass runs: one line of parked (Ho, F) pairs per CTA and group, as long as the longest
    // subject stream a group of any such tile reads (+ wind-up and the two-step prefetch); runs
    // execute concurrently on the side streams, so each gets its own lines
    int mp_stride = 0;
    size_t mp_runs = 0, mp_lines_used = 0;
    for (const auto &run : runs) {
        if (!(e->tiles[run.first].variant & BGCA_VARIANT_MP)) continue;
        ++mp_runs;
        const int ng = CTA_THREADS / 32;
        for (size_t i = run.first; i < run.second; ++i) {
            const bgca_tile &T = e->tiles[i];
            for (int g = 0; g < ng; ++g) {
                int tokens = 0;
                for (int s = T.s_begin + g; s < T.s_end; s += ng) tokens += e->len[s] + 1;
                mp_stride = std::max(mp_stride, tokens);
            }
        }
    }
    if (mp_runs) {
        mp_stride = (mp_stride + 32 + 2 + 15) & ~15;
        CUDA_TRY(e->d_mp_scratch.reserve(mp_runs * (size_t)pair16_mp_max_grid(e->n_sm) * (CTA_THREADS / 32) *
                                         (size_t)mp_stride));
    }
    // square plans: the s16x2 kernels stage their scores in sorted space as int16 (coalescing
    // stores, see ScoreParams::scores16) and one permutation kernel carries them into `scores`
    const bool stage16 = scores && !e->cross_planned && e->n_db == 0 && e->n <= kPermuteMaxN &&
                         !std::getenv("BGCA_NO_STAGE16");
    const int s16_stride = (e->n + 7) & ~7;
    // does this call write EVERY entry through the s16x2 kernels?  Otherwise the staging matrix
    // starts out "unwritten" and the permutation leaves such entries of `scores` alone
    const bool stage_full = e->plan_info.n_pairs == (int64_t)e->n * (e->n + 1) / 2 && e->plan_info.n_fallback32 == 0;
    if (stage16) {
        CUDA_TRY(e->d_s16.reserve((size_t)e->n * s16_stride));
        if (!stage_full)
            CUDA_TRY(cudaMemsetAsync(e->d_s16.p, BGCA_S16_UNWRITTEN & 0xFF, sizeof(int16_t) * (size_t)e->n * s16_stride, st));
    }
    // Pipelined host sink: with several runs, every entry staged in sorted space and nothing on the
    // 32-bit kernels, sorted row x is final once every run that holds a tile with qa <= x has
    // finished.  Runs are then launched in ascending order of their highest query; behind each one
    // its row block of the int16 staging matrix goes to pinned host memory on the copy engine (no
    // SM needed: the persistent DP kernels hold them all) and host threads permute its columns,
    // widen to int32 and put the rows in their place in the caller's matrix — while later runs compute.
    bool pipelined = sink && stage16 && stage_full && runs.size() > 1;
    if (pipelined) {
        // pinned copy of the staging matrix; a host that cannot pin that much takes the plain path
        const size_t elems = (size_t)e->n * s16_stride;
        if (e->h_stage_elems < elems) {
            if (e->h_stage) cudaFreeHost(e->h_stage);
            e->h_stage = nullptr; e->h_stage_elems = 0;
            if (cudaHostAlloc(&e->h_stage, elems * sizeof(int16_t), cudaHostAllocDefault) == cudaSuccess)
                e->h_stage_elems = elems;
            else {
                (void)cudaGetLastError();
                e->h_stage = nullptr;
                pipelined = false;
            }
        }
    }
    std::vector<size_t> run_order(runs.size());
    std::iota(run_order.begin(), run_order.end(), 0);
    std::vector<int> run_min_q(runs.size(), 0), run_max_q(runs.size(), 0);
    if (pipelined) {
        for (size_t r = 0; r < runs.size(); ++r) {
            int lo = e->n, hi = -1;
            for (size_t i = runs[r].first; i < runs[r].second; ++i) {
                const bgca_tile &T = e->tiles[i];
                lo = std::min(lo, T.qa);
                hi = std::max(hi, std::max(T.qa, T.qb));
            }
            run_min_q[r] = lo; run_max_q[r] = hi;
        }
        std::stable_sort(run_order.begin(), run_order.end(), [&](size_t a, size_t b) { return run_max_q[a] < run_max_q[b]; });
        while (e->ev_pool.size() < 2 * runs.size()) {
            cudaEvent_t ev;
            CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->ev_pool.push_back(ev);
        }
        sink->used = true;
    }
    CUDA_TRY(e->d_counters.reserve(runs.size() + 1));
    CUDA_TRY(cudaMemsetAsync(e->d_counters.p, 0, sizeof(int32_t) * (runs.size() + 1), st));
    CUDA_TRY(cudaEventRecord(e->ev0, st));
    CUDA_TRY(cudaEventRecord(e->ev_fork, st));
    int n_side_used = 0;
    for (size_t ri = 0; ri < runs.size(); ++ri) {
        const size_t r = run_order[ri];
        const size_t b = runs[r].first, en = runs[r].second;
        const int variant = e->tiles[b].variant;
        // 32-bit kernels (scores beyond int16) write the caller's matrix directly
        const bool to_s16 = stage16 && variant != 0 && !(variant & BGCA_VARIANT_W32);
        p.scores = to_s16 ? nullptr : scores;
        p.scores16 = to_s16 ? e->d_s16.p : nullptr;
        p.s16_stride = s16_stride;
        if (variant == 0) {
            // 32-bit fallback: expand the tiles into an explicit (q, s) list
            std::vector<int32_t> pq, ps;
            for (size_t i = b; i < en; ++i) {
                const bgca_tile &T = e->tiles[i];
                if (T.reserved) {            // self-only tile of a cross plan
                    pq.push_back(T.qa); ps.push_back(T.qa);
                    if (T.qb >= 0) { pq.push_back(T.qb); ps.push_back(T.qb); }
                    continue;
                }
                // square plans list the mirrored pair (qb, s < qb) on the qa side already; cross
                // plans pair queries with database subjects that may sort before OR behind them
                // (a persistent database sits in front of its queries), so nothing is mirrored there
                for (int s = T.s_begin; s < T.s_end; ++s) {
                    pq.push_back(T.qa); ps.push_back(s);
                    if (T.qb >= 0 && (e->cross_planned || s >= T.qb)) { pq.push_back(T.qb); ps.push_back(s); }
                }
            }
            CUDA_TRY(e->d_pq.reserve(pq.size()));
            CUDA_TRY(e->d_ps.reserve(ps.size()));
            CUDA_TRY(cudaMemcpyAsync(e->d_pq.p, pq.data(), sizeof(int32_t) * pq.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(e->d_ps.p, ps.data(), sizeof(int32_t) * ps.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));   // pq/ps are stack-owned host vectors
            int rc = run_s32_pairs(e, p, e->d_pq.p, e->d_ps.p, (long long)pq.size(), nullptr, st);
            if (rc != BGCA_OK) return rc;
            continue;
        }
        p.tiles = e->d_tiles.p + b;
        p.n_tiles = (int32_t)(en - b);
        p.tile_counter = e->d_counters.p + r;
        const int G = (variant >> 8) & 0xFF, K = variant & 0xFF;
        const bool multipass = (variant & BGCA_VARIANT_MP) != 0;
        cudaError_t err = cudaErrorInvalidValue;
        int grid = 0;
        const bool local = e->mode == BGCA_MODE_LOCAL;
        // [G index][mode][gap preset]; preset 0 takes the gap scores at run time
        static const Pair16Launcher table[6][2][3] = {
            {{launch_pair16_g1_m0_p0, launch_pair16_g1_m0_p1, launch_pair16_g1_m0_p2},
             {launch_pair16_g1_m1_p0, launch_pair16_g1_m1_p1, launch_pair16_g1_m1_p2}},
            {{launch_pair16_g2_m0_p0, launch_pair16_g2_m0_p1, launch_pair16_g2_m0_p2},
             {launch_pair16_g2_m1_p0, launch_pair16_g2_m1_p1, launch_pair16_g2_m1_p2}},
            {{launch_pair16_g4_m0_p0, launch_pair16_g4_m0_p1, launch_pair16_g4_m0_p2},
             {launch_pair16_g4_m1_p0, launch_pair16_g4_m1_p1, launch_pair16_g4_m1_p2}},
            {{launch_pair16_g8_m0_p0, launch_pair16_g8_m0_p1, launch_pair16_g8_m0_p2},
             {launch_pair16_g8_m1_p0, launch_pair16_g8_m1_p1, launch_pair16_g8_m1_p2}},
            {{launch_pair16_g16_m0_p0, launch_pair16_g16_m0_p1, launch_pair16_g16_m0_p2},
             {launch_pair16_g16_m1_p0, launch_pair16_g16_m1_p1, launch_pair16_g16_m1_p2}},
            {{launch_pair16_g32_m0_p0, launch_pair16_g32_m0_p1, launch_pair16_g32_m0_p2},
             {launch_pair16_g32_m1_p0, launch_pair16_g32_m1_p1, launch_pair16_g32_m1_p2}}};
        const int gi = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : (G == 8) ? 3 : (G == 16) ? 4 : (G == 32) ? 5 : -1;
        int preset = 0;
        if (e->open == -12 && e->extend == -1) preset = 1;
        else if (e->open == -11 && e->extend == -1) preset = 2;
        if (std::getenv("BGCA_NO_PRESET")) preset = 0;   // tuning aid: force the run-time-gap kernels
        // variants round-robin over the side streams; the caller's stream joins them below
        cudaStream_t ks = st;
        if (runs.size() > 1) {
            const int si = (int)(ri % bgca_engine::kSideStreams);
            ks = e->side[si];
            if (si >= n_side_used) {
                CUDA_TRY(cudaStreamWaitEvent(ks, e->ev_fork, 0));
                n_side_used = si + 1;
            }
        }
        if (multipass) {
            static const Pair16Launcher mp_table[2][2][3] = {
                {{launch_pair16_mp_m0_p0, launch_pair16_mp_m0_p1, launch_pair16_mp_m0_p2},
                 {launch_pair16_mp_m1_p0, launch_pair16_mp_m1_p1, launch_pair16_mp_m1_p2}},
                {{launch_pair16_mpw_m0_p0, launch_pair16_mpw_m0_p1, launch_pair16_mpw_m0_p2},
                 {launch_pair16_mpw_m1_p0, launch_pair16_mpw_m1_p1, launch_pair16_mpw_m1_p2}}};
            p.mp_scratch = e->d_mp_scratch.p + mp_lines_used * (size_t)mp_stride;
            p.mp_stride = mp_stride;
            mp_lines_used += (size_t)pair16_mp_max_grid(e->n_sm) * (CTA_THREADS / 32);
            err = mp_table[(variant & BGCA_VARIANT_W32) ? 1 : 0][local ? 0 : 1][preset](K, p, e->n_sm, ks, &grid);
        } else if (gi >= 0) err = table[gi][local ? 0 : 1][preset](K, p, e->n_sm, ks, &grid);
        if (err != cudaSuccess)
            return fail(BGCA_ERR_CUDA, std::string("DP kernel launch (G=") + std::to_string(G) + ",K=" +
                                           std::to_string(K) + "): " + cudaGetErrorString(err));
        e->last_launches++;
        if (pipelined) CUDA_TRY(cudaEventRecord(e->ev_pool[r], ks));
    }
    if (pipelined) {
        // row blocks in launch order: (previous highest query, this run's highest query]
        int row_lo = 0;
        std::vector<std::pair<int, int>> blocks;
        for (size_t ri = 0; ri < runs.size(); ++ri) {
            const size_t r = run_order[ri];
            const int row_hi = (ri + 1 == runs.size()) ? e->n : run_max_q[r] + 1;
            blocks.push_back({row_lo, std::max(row_lo, row_hi)});
            if (row_hi > row_lo) {
                for (size_t r2 = 0; r2 < runs.size(); ++r2)       // every run that can still write these rows
                    if (run_min_q[r2] < row_hi) CUDA_TRY(cudaStreamWaitEvent(e->copy_stream, e->ev_pool[r2], 0));
                CUDA_TRY(cudaMemcpyAsync(e->h_stage + (size_t)row_lo * s16_stride, e->d_s16.p + (size_t)row_lo * s16_stride,
                                         sizeof(int16_t) * (size_t)(row_hi - row_lo) * s16_stride, cudaMemcpyDeviceToHost,
                                         e->copy_stream));
                row_lo = row_hi;
            }
            CUDA_TRY(cudaEventRecord(e->ev_pool[runs.size() + ri], e->copy_stream));
        }
        for (int i = 0; i < n_side_used; ++i) {
            CUDA_TRY(cudaEventRecord(e->ev_join[i], e->side[i]));
            CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join[i], 0));
        }
        CUDA_TRY(cudaEventRecord(e->ev1, st));
        e->timed = true;
        // host side: as each block lands in pinned memory, out[order[x]][j] = S[x][inv[j]]
        const int n = e->n;
        const int32_t *inv = e->inv_order.data(), *order = e->order.data();
        const int16_t *hs = e->h_stage;
        int32_t *out = sink->scores_host;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n_threads = (int)std::min(8u, hw);
        for (size_t ri = 0; ri < blocks.size(); ++ri) {
            CUDA_TRY(cudaEventSynchronize(e->ev_pool[runs.size() + ri]));
            const int x0 = blocks[ri].first, x1 = blocks[ri].second;
            if (x1 <= x0) continue;
            auto work = [=](int t) {
                for (int x = x0 + t; x < x1; x += n_threads) {
                    const int16_t *src = hs + (size_t)x * s16_stride;
                    int32_t *dst = out + (size_t)order[x] * n;
                    for (int j = 0; j < n; ++j) dst[j] = src[inv[j]];
                }
            };
            const int use = std::min(n_threads, x1 - x0);
            std::vector<std::thread> pool;
            for (int t = 1; t < use; ++t) pool.emplace_back(work, t);
            work(0);
            for (int t = use; t < n_threads; ++t) work(t);      // fewer rows than threads: leftovers stay here
            for (auto &th : pool) th.join();
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        return BGCA_OK;
    }
    for (int i = 0; i < n_side_used; ++i) {
        CUDA_TRY(cudaEventRecord(e->ev_join[i], e->side[i]));
        CUDA_TRY(cudaStreamWaitEvent(st, e->ev_join[i], 0));
    }
    CUDA_TRY(cudaEventRecord(e->ev1, st));
    e->timed = true;
    if (stage16) {
        CUDA_TRY(launch_permute16(e->d_s16.p, s16_stride, e->d_inv_order.p, e->n, scores, stage_full ? 0 : 1,
                                  e->n_sm, st));
        e->last_launches++;
    }
    return BGCA_OK;
}

int bgca_score(bgca_engine *e, int32_t *scores, int32_t *rowbest, int32_t *selfsc, int32_t type_check,
               void *stream)
try {
    return score_impl(e, scores, rowbest, selfsc, type_check, stream, nullptr);
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_score: ") + ex.what());
}

// ORIGINAL-space device pair lists -> SORTED-space lists in e->d_pq / e->d_ps (lists are small
// on these explicit-pair paths, so the translation runs on the host)
static int stage_pair_list(bgca_engine *e, const int32_t *pi, const int32_t *pj, int64_t npairs,
                           int max_len_allowed, const char *who, cudaStream_t st, int *max_len_seen = nullptr,
                           std::vector<int32_t> *sorted_q = nullptr, std::vector<int32_t> *sorted_s = nullptr)
{
    int longest = 0;
    std::vector<int32_t> hi(npairs), hj(npairs);
    CUDA_TRY(cudaMemcpyAsync(hi.data(), pi, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(hj.data(), pj, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int64_t k = 0; k < npairs; ++k) {
        if (hi[k] < 0 || hi[k] >= e->n || hj[k] < 0 || hj[k] >= e->n)
            return fail(BGCA_ERR_INVALID, std::string(who) + ": index out of range");
        hi[k] = e->inv_order[hi[k]];
        hj[k] = e->inv_order[hj[k]];
        longest = std::max(longest, std::max(e->len[hi[k]], e->len[hj[k]]));
        if (e->len[hi[k]] > max_len_allowed || e->len[hj[k]] > max_len_allowed)
            return fail(BGCA_ERR_RANGE, std::string(who) + ": sequence longer than " +
                                            std::to_string(max_len_allowed) + " residues");
    }
    CUDA_TRY(e->d_pq.reserve(npairs));
    CUDA_TRY(e->d_ps.reserve(npairs));
    CUDA_TRY(cudaMemcpyAsync(e->d_pq.p, hi.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_ps.p, hj.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (max_len_seen) *max_len_seen = longest;
    if (sorted_q) sorted_q->swap(hi);
    if (sorted_s) sorted_s->swap(hj);
    return BGCA_OK;
}

// K7, second form.  Work items of at most kStats2ItemPairs pairs of one query, one item list and
// launch per rows-per-lane class of the query.
constexpr int kStats2ItemPairs = 32;
static const int kStats2Ks[] = {BGCA_STATS2_KS_LIST};
constexpr int kStats2Classes = sizeof(kStats2Ks) / sizeof(kStats2Ks[0]);

// Launches over device-resident grouped pairs: items [cls_first[c], cls_first[c+1]) belong to
// class c.  hint != NULL (local mode): the kernels that take the best score as known; should a
// hint prove too high the whole list runs again without hints (exact either way).
static int launch_stats2_classes(bgca_engine *e, const Stats2Item *d_items, const size_t *cls_first,
                                 const int32_t *d_subj, const int32_t *d_out_idx, const int32_t *d_hint,
                                 bgca_alnstats *out, cudaStream_t st)
{
    Stats2Params sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.codes = e->d_codes.p;
    sp.seq_off = e->d_seq_off.p;
    sp.seq_len = e->d_len.p;
    sp.subj = d_subj;
    sp.out_idx = d_out_idx;
    sp.out = out;
    sp.open = e->open;
    sp.extend = e->extend;
    sp.openh = e->open * 1024;
    sp.exth = e->extend * 1024;
    std::memcpy(sp.sub, e->sub, NSYM * NSYM);
    const bool hinted = d_hint != nullptr && e->mode == BGCA_MODE_LOCAL;
    CUDA_TRY(e->d_counters.reserve(2 * kStats2Classes + 2));
    e->last_launches = 0;
    CUDA_TRY(cudaEventRecord(e->ev0, st));
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool use_hint = hinted && attempt == 0;
        CUDA_TRY(cudaMemsetAsync(e->d_counters.p, 0, sizeof(int32_t) * (kStats2Classes + 2), st));
        sp.hint = use_hint ? d_hint : nullptr;
        sp.redo_count = e->d_counters.p + kStats2Classes;
        for (int c = 0; c < kStats2Classes; ++c) {
            if (cls_first[c + 1] == cls_first[c]) continue;
            sp.items = d_items + cls_first[c];
            sp.n_items = (int32_t)(cls_first[c + 1] - cls_first[c]);
            sp.item_counter = e->d_counters.p + c;
            CUDA_TRY(launch_stats2(e->mode, kStats2Ks[c], sp, e->n_sm, st));
            e->last_launches++;
        }
        if (!use_hint) break;
        CUDA_TRY(cudaEventRecord(e->ev1, st));
        int32_t redo = 0;
        CUDA_TRY(cudaMemcpyAsync(&redo, sp.redo_count, sizeof(redo), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (redo == 0) { e->timed = true; return BGCA_OK; }
    }
    CUDA_TRY(cudaEventRecord(e->ev1, st));
    e->timed = true;
    return BGCA_OK;
}

// explicit pair list (SORTED indices hq, hs; optional known scores): grouped by query on the host
static int run_stats2(bgca_engine *e, const std::vector<int32_t> &hq, const std::vector<int32_t> &hs,
                      const std::vector<int32_t> *known, bgca_alnstats *out, cudaStream_t st)
{
    const int64_t npairs = (int64_t)hq.size();
    std::vector<int32_t> idx(npairs);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return hq[a] < hq[b]; });
    std::vector<int32_t> subj(npairs), hint(known ? npairs : 0);
    for (int64_t k = 0; k < npairs; ++k) subj[k] = hs[idx[k]];
    if (known)
        for (int64_t k = 0; k < npairs; ++k) hint[k] = (*known)[idx[k]];
    std::vector<Stats2Item> items[kStats2Classes];
    for (int64_t a = 0; a < npairs;) {
        int64_t b = a;
        const int q = hq[idx[a]];
        while (b < npairs && hq[idx[b]] == q) ++b;
        const int ks = stats2_rows_per_lane(e->len[q]);
        int cls = 0;
        while (kStats2Ks[cls] != ks) ++cls;
        for (int64_t f = a; f < b; f += kStats2ItemPairs)
            items[cls].push_back(Stats2Item{q, (int32_t)f, (int32_t)std::min<int64_t>(kStats2ItemPairs, b - f), 0});
        a = b;
    }
    std::vector<Stats2Item> flat;
    size_t cls_first[kStats2Classes + 1] = {0};
    for (int c = 0; c < kStats2Classes; ++c) {
        flat.insert(flat.end(), items[c].begin(), items[c].end());
        cls_first[c + 1] = flat.size();
    }
    CUDA_TRY(e->d_pq.reserve(npairs));                      // subjects, grouped
    CUDA_TRY(e->d_ps.reserve(npairs));                      // caller's position of each grouped pair
    CUDA_TRY(e->d_st2_items.reserve(flat.size() * 4));
    CUDA_TRY(cudaMemcpyAsync(e->d_pq.p, subj.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_ps.p, idx.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_st2_items.p, flat.data(), sizeof(Stats2Item) * flat.size(), cudaMemcpyHostToDevice, st));
    if (known) {
        CUDA_TRY(e->d_st2_hint.reserve(npairs));
        CUDA_TRY(cudaMemcpyAsync(e->d_st2_hint.p, hint.data(), sizeof(int32_t) * npairs, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));                    // the host vectors above die with this frame
    return launch_stats2_classes(e, reinterpret_cast<const Stats2Item *>(e->d_st2_items.p), cls_first, e->d_pq.p,
                                 e->d_ps.p, known ? e->d_st2_hint.p : nullptr, out, st);
}

// the FP64-compare form of K7 (gotoh_stats2.cuh) serves a batch when every sequence has at most
// 1023 residues and every DP value stays well inside the +-2^19 its biased score field holds (it
// does for any protein matrix and gap scores in use)
static bool stats2_serves(const bgca_engine *e, int longest)
{
    const long long worst = 3LL * -e->open + (long long)-e->extend * (2 * longest + 2) +
                            (long long)e->max_abs_sub * longest + 4096;
    return longest <= ST2_MAXLEN && worst < (1 << 19) - 4096 && !std::getenv("BGCA_STATS_V1");
}

int bgca_score_pairs32(bgca_engine *e, const int32_t *pi, const int32_t *pj, int64_t npairs,
                       int32_t *out, void *stream)
try {
    if (!e || !e->packed || !e->scoring_set) return fail(BGCA_ERR_INVALID, "bgca_score_pairs32: pack and set scoring first");
    if (!pi || !pj || !out || npairs < 0) return fail(BGCA_ERR_INVALID, "bgca_score_pairs32: bad arguments");
    if (npairs == 0) return BGCA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    int rc = stage_pair_list(e, pi, pj, npairs, 65535, "bgca_score_pairs32", st);
    if (rc != BGCA_OK) return rc;
    ScoreParams p;
    fill_params(e, p, nullptr, nullptr, nullptr, 0);
    e->last_launches = 0;
    return run_s32_pairs(e, p, e->d_pq.p, e->d_ps.p, npairs, out, st);
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_score_pairs32: ") + ex.what());
}

static int pair_stats_impl(bgca_engine *e, const int32_t *pi, const int32_t *pj, const int32_t *known_scores,
                           int64_t npairs, bgca_alnstats *out, void *stream, const char *who)
{
    if (!e || !e->packed || !e->scoring_set) return fail(BGCA_ERR_INVALID, std::string(who) + ": pack and set scoring first");
    if (!pi || !pj || !out || npairs < 0) return fail(BGCA_ERR_INVALID, std::string(who) + ": bad arguments");
    if (npairs == 0) return BGCA_OK;
    if (npairs > 0x7FFFFFF0) return fail(BGCA_ERR_RANGE, std::string(who) + ": too many pairs for one call");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    int longest = 0;
    std::vector<int32_t> hq, hs;
    int rc = stage_pair_list(e, pi, pj, npairs, 32767, who, st, &longest, &hq, &hs);
    if (rc != BGCA_OK) return rc;
    // every protein domain fits the one-word tuple; longer sequences take the 96-bit one
    const int compact = (longest <= 1023 && e->max_abs_sub * 1023 < (1 << 21) && -e->open < 1000 && -e->extend < 1000);
    if (compact && stats2_serves(e, longest)) {
        std::vector<int32_t> known;
        if (known_scores && e->mode == BGCA_MODE_LOCAL) {
            known.resize(npairs);
            CUDA_TRY(cudaMemcpyAsync(known.data(), known_scores, sizeof(int32_t) * npairs, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        return run_stats2(e, hq, hs, known.empty() ? nullptr : &known, out, st);
    }
    ScoreParams p;
    fill_params(e, p, nullptr, nullptr, nullptr, 0);
    int max_len = 0;
    for (int r = 0; r < e->n; ++r) max_len = std::max(max_len, e->len[r]);
    const int warps_per_cta = CTA_THREADS / 32;
    int grid = e->n_sm * 4;
    if ((long long)grid * warps_per_cta > npairs)
        grid = (int)std::max<long long>(1, (npairs + warps_per_cta - 1) / warps_per_cta);
    const int stride = (max_len + 3) & ~3;
    CUDA_TRY(e->d_scratch_st.reserve((size_t)grid * warps_per_cta * 2 * stride));
    CUDA_TRY(cudaEventRecord(e->ev0, st));
    CUDA_TRY(launch_stats(e->mode, compact, p, e->d_pq.p, e->d_ps.p, npairs, e->d_scratch_st.p, stride, out, grid, st));
    CUDA_TRY(cudaEventRecord(e->ev1, st));
    e->timed = true;
    e->last_launches = 1;
    return BGCA_OK;
}

int bgca_pair_stats(bgca_engine *e, const int32_t *pi, const int32_t *pj, int64_t npairs,
                    bgca_alnstats *out, void *stream)
try {
    return pair_stats_impl(e, pi, pj, nullptr, npairs, out, stream, "bgca_pair_stats");
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_pair_stats: ") + ex.what());
}

int bgca_pair_stats_hinted(bgca_engine *e, const int32_t *pi, const int32_t *pj, const int32_t *known_scores,
                           int64_t npairs, bgca_alnstats *out, void *stream)
try {
    return pair_stats_impl(e, pi, pj, known_scores, npairs, out, stream, "bgca_pair_stats_hinted");
} catch (const std::exception &ex) {
    return fail(BGCA_ERR_RANGE, std::string("bgca_pair_stats_hinted: ") + ex.what());
}

int bgca_select_pairs(bgca_engine *e, const int32_t *scores, int32_t min_score, int32_t same_type_only,
                      int64_t *count, void *stream)
try {
    if (!e || !e->packed || !scores || !count) return fail(BGCA_ERR_INVALID, "bgca_select_pairs: pack first; NULL argument");
    if (e->n_query != 0 || e->n_db != 0)
        return fail(BGCA_ERR_INVALID, "bgca_select_pairs: needs one undivided collection (a square score matrix)");
    if (same_type_only && !e->has_type)
        return fail(BGCA_ERR_INVALID, "bgca_select_pairs: same_type_only needs type_of_seq at pack time");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    const int n = e->n;
    CUDA_TRY(e->d_inv_order.reserve(n));
    CUDA_TRY(e->d_sel_cnt.reserve(n));
    CUDA_TRY(e->d_sel_rowstart.reserve((size_t)n + 1));
    CUDA_TRY(e->d_sel_cls.reserve(3 * 16));                 // [items per class | base per class | fill cursors]
    CUDA_TRY(cudaMemcpyAsync(e->d_inv_order.p, e->inv_order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(e->d_sel_cls.p, 0, sizeof(int32_t) * 48, st));
    CUDA_TRY(launch_select_count(scores, n, min_score, same_type_only ? e->d_type.p : nullptr, e->d_inv_order.p,
                                 e->d_len.p, kStats2ItemPairs, e->d_sel_cnt.p, e->d_sel_rowstart.p, e->d_sel_cls.p, st));
    int64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, e->d_sel_rowstart.p + n, sizeof(total), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(e->sel_cls_items, e->d_sel_cls.p, sizeof(int32_t) * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (total > 0x7FFFFFF0) return fail(BGCA_ERR_RANGE, "bgca_select_pairs: more than 2^31 pairs selected; raise min_score");
    e->sel_valid = true;
    e->sel_min_score = min_score;
    e->sel_same_type = same_type_only;
    e->sel_total = total;
    *count = total;
    return BGCA_OK;
} catch (const std::exception &ex) {
    return fail(BGCA_ERR_RANGE, std::string("bgca_select_pairs: ") + ex.what());
}

int bgca_pair_stats_selected(bgca_engine *e, const int32_t *scores, int32_t *pairs_out, bgca_alnstats *out,
                             void *stream)
try {
    if (!e || !e->sel_valid || !e->packed) return fail(BGCA_ERR_INVALID, "bgca_pair_stats_selected: call bgca_select_pairs first");
    if (!scores || !out) return fail(BGCA_ERR_INVALID, "bgca_pair_stats_selected: NULL argument");
    if (!e->scoring_set) return fail(BGCA_ERR_INVALID, "bgca_pair_stats_selected: call bgca_set_scoring first");
    e->sel_valid = false;                                   // one selection, one statistics pass
    const int64_t total = e->sel_total;
    if (total == 0) return BGCA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    const int n = e->n;
    int longest = 0;
    for (int r = 0; r < n; ++r) longest = std::max(longest, e->len[r]);
    if (longest > 32767) return fail(BGCA_ERR_RANGE, "bgca_pair_stats_selected: sequence longer than 32767 residues");
    const int compact = (longest <= 1023 && e->max_abs_sub * 1023 < (1 << 21) && -e->open < 1000 && -e->extend < 1000);
    const bool v2 = compact && stats2_serves(e, longest);
    size_t cls_first[kStats2Classes + 1] = {0};
    int32_t base[16] = {0};
    for (int c = 0; c < kStats2Classes; ++c) {
        base[c] = (int32_t)cls_first[c];
        cls_first[c + 1] = cls_first[c] + (size_t)e->sel_cls_items[c];
    }
    CUDA_TRY(e->d_pq.reserve(total));                       // query of every pair, SORTED index
    CUDA_TRY(e->d_ps.reserve(total));                       // subject
    CUDA_TRY(e->d_st2_hint.reserve(total));
    CUDA_TRY(e->d_st2_items.reserve(std::max<size_t>(cls_first[kStats2Classes], 1) * 4));
    CUDA_TRY(cudaMemcpyAsync(e->d_sel_cls.p + 16, base, sizeof(int32_t) * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(e->d_sel_cls.p + 32, 0, sizeof(int32_t) * 16, st));
    CUDA_TRY(launch_select_fill(scores, n, e->sel_min_score, e->sel_same_type ? e->d_type.p : nullptr,
                                e->d_inv_order.p, e->d_len.p, e->d_sel_rowstart.p, kStats2ItemPairs,
                                e->d_sel_cls.p + 16, e->d_sel_cls.p + 32, pairs_out, e->d_pq.p, e->d_ps.p,
                                e->d_st2_hint.p, reinterpret_cast<Stats2Item *>(e->d_st2_items.p), st));
    if (v2)
        return launch_stats2_classes(e, reinterpret_cast<const Stats2Item *>(e->d_st2_items.p), cls_first, e->d_ps.p,
                                     nullptr, e->d_st2_hint.p, out, st);
    // sequences beyond the compact form: the first K7 kernel over the same device lists
    ScoreParams p;
    fill_params(e, p, nullptr, nullptr, nullptr, 0);
    const int warps_per_cta = CTA_THREADS / 32;
    int grid = e->n_sm * 4;
    if ((long long)grid * warps_per_cta > total)
        grid = (int)std::max<long long>(1, (total + warps_per_cta - 1) / warps_per_cta);
    const int stride = (longest + 3) & ~3;
    CUDA_TRY(e->d_scratch_st.reserve((size_t)grid * warps_per_cta * 2 * stride));
    CUDA_TRY(cudaEventRecord(e->ev0, st));
    CUDA_TRY(launch_stats(e->mode, compact, p, e->d_pq.p, e->d_ps.p, total, e->d_scratch_st.p, stride, out, grid, st));
    CUDA_TRY(cudaEventRecord(e->ev1, st));
    e->timed = true;
    e->last_launches = 1;
    return BGCA_OK;
} catch (const std::exception &ex) {
    return fail(BGCA_ERR_RANGE, std::string("bgca_pair_stats_selected: ") + ex.what());
}

int bgca_reduce_bgc(bgca_engine *e, const int32_t *rowbest, const int32_t *selfsc, int64_t *best,
                    int64_t *selfsum, double *sim, void *stream)
{
    if (!e || !e->packed || !e->has_bgc) return fail(BGCA_ERR_INVALID, "bgca_reduce_bgc: pack with bgc_of_seq first");
    if (!rowbest || !selfsc || !best || !selfsum) return fail(BGCA_ERR_INVALID, "bgca_reduce_bgc: NULL argument");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(launch_bgc_reduce(rowbest, selfsc, e->d_bgc_start.p, e->d_bgc_doms.p, e->n_bgc, best, selfsum,
                               sim, (cudaStream_t)stream));
    return BGCA_OK;
}

int bgca_score_all_host(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n,
                        int32_t *scores_host)
try {
    if (!e || !scores_host) return fail(BGCA_ERR_INVALID, "bgca_score_all_host: NULL argument");
    int rc = bgca_pack(e, ascii, offsets, n, nullptr, nullptr, 0, 0, 0, nullptr);
    if (rc != BGCA_OK) return rc;
    rc = bgca_plan(e, 0, 0, 0, 1, nullptr);
    if (rc != BGCA_OK) return rc;
    // the device-side matrix is engine-owned scratch, kept across calls (repeated calls on
    // same-sized batches pay no cudaMalloc/cudaFree)
    const size_t bytes = sizeof(int32_t) * (size_t)n * (size_t)n;
    CUDA_TRY(e->d_scores_host_call.reserve((size_t)n * (size_t)n));
    int32_t *d_scores = e->d_scores_host_call.p;
    // the pipelined sink needs no cleared matrix (every entry is written); decided inside
    HostSink sink{scores_host};
    cudaError_t err = cudaSuccess;
    rc = score_impl(e, d_scores, nullptr, nullptr, 0, nullptr, std::getenv("BGCA_NO_PIPELINE") ? nullptr : &sink);
    if (rc == BGCA_OK && !sink.used) err = cudaMemcpy(scores_host, d_scores, bytes, cudaMemcpyDeviceToHost);
    if (rc != BGCA_OK) return rc;
    if (err != cudaSuccess) return fail(BGCA_ERR_CUDA, std::string("bgca_score_all_host: ") + cudaGetErrorString(err));
    return BGCA_OK;
} catch (const std::exception &ex) {   // std::bad_alloc and friends must not cross the C ABI
    return fail(BGCA_ERR_RANGE, std::string("bgca_score_all_host: ") + ex.what());
}

int bgca_last_launches(const bgca_engine *e) { return e ? e->last_launches : 0; }

int bgca_last_dp_ms(bgca_engine *e, float *ms)
{
    if (!e || !ms || !e->timed) return fail(BGCA_ERR_INVALID, "bgca_last_dp_ms: no timed bgca_score yet");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaEventSynchronize(e->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return BGCA_OK;
}

int bgca_last_pack_ms(bgca_engine *e, float *ms)
{
    if (!e || !ms || !e->pack_timed) return fail(BGCA_ERR_INVALID, "bgca_last_pack_ms: no bgca_pack yet");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaEventSynchronize(e->evp1));
    CUDA_TRY(cudaEventElapsedTime(ms, e->evp0, e->evp1));
    return BGCA_OK;
}

int bgca_dpx_probe(int device, double *out, int32_t n)
{
    if (!out || n <= 0) return fail(BGCA_ERR_INVALID, "bgca_dpx_probe: bad arguments");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(BGCA_ERR_CUDA, "bgca_dpx_probe: no usable CUDA device");
    if (dpx_probe_run(device, out, n) != 0) return fail(BGCA_ERR_CUDA, "bgca_dpx_probe: CUDA failure");
    return BGCA_OK;
}

}  // extern "C"
