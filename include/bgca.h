/*
 * bgca.h — C ABI of the B200-native all-vs-all domain alignment engine
 * ("BGC aligner").  Plain pointers and sizes only; no torch / C++ types.
 *
 * What this boundary replaces in the reference (jorgecnavarrom/BGClib):
 * NOTHING EXISTING — the reference has no FFI, no aligner and no distance code
 * (SURVEY.md §0 F2/F3, §8b).  The feature is only announced: "Distance
 * calculation (between e.g. two organism's BGCs ...)" at BGClib/Readme.md:16-18
 * and the commented-out `--comparative` switch at BGCtoolkit.py:157-161.  The
 * entry points below are therefore what a binding for that announced feature
 * would call; each one names the reference symbols whose data it consumes.
 *
 * All functions return 0 on success or a negative bgca_status; the message of
 * the last failure on the calling thread is returned by bgca_last_error().
 * Index spaces: "original" = order in which sequences were handed to
 * bgca_pack(); "sorted" = engine-internal order (by type, then length).
 * Every caller-visible array is in ORIGINAL order unless stated otherwise.
 */
#ifndef BGCA_H
#define BGCA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGCA_VERSION 105

#define BGCA_MODE_LOCAL 0  /* Smith-Waterman/Gotoh, result = best cell, >= 0        */
#define BGCA_MODE_GLOBAL 1 /* Needleman-Wunsch/Gotoh, end gaps cost like internal   */

#define BGCA_NSYM 24       /* ARNDCQEGHILKMFPSTWYVBZX*  (Biopython BLOSUM62 order)  */
#define BGCA_PAD_CODE 24   /* filler written after each packed sequence             */

typedef enum bgca_status {
    BGCA_OK = 0,
    BGCA_ERR_INVALID = -1,   /* bad argument / call order                           */
    BGCA_ERR_ALPHABET = -2,  /* a residue outside the 24-letter alphabet (ValueError
                                in Biopython's PairwiseAligner)   */
    BGCA_ERR_EMPTY = -3,     /* a zero-length sequence (ValueError in Biopython)     */
    BGCA_ERR_CUDA = -4,      /* CUDA runtime failure, including "no device"         */
    BGCA_ERR_RANGE = -5      /* sizes beyond what the engine supports               */
} bgca_status;

typedef struct bgca_engine bgca_engine;

/* One unit of work for the DP kernels, in SORTED index space: the query pair
 * (qa, qb) — packed in the two 16-bit halves of every DPX op — against the
 * subjects [s_begin, s_end).  qb == -1 : single query.  variant encodes the
 * kernel instantiation ((G << 8) | K : G lanes per alignment, K query rows per
 * lane; | 0x10000 = multi-pass sweep for queries longer than one pass; | 0x20000 =
 * its 32-bit form, single query), 0 = 32-bit pair-list kernel. */
typedef struct bgca_tile {
    int32_t qa, qb, s_begin, s_end;
    int32_t variant;
    int32_t reserved;
    int64_t cells; /* (len[qa]+len[qb]) * sum(len[s]) — scheduling weight */
} bgca_tile;

typedef struct bgca_plan_info {
    int64_t n_tiles;       /* tiles owned by this rank                              */
    int64_t n_pairs;       /* unique pairs i<=j (original space) owned by this rank */
    int64_t cells;         /* sum len_i*len_j over those unique pairs (true lengths,
                              the GCUPS numerator)                   */
    int64_t cells_computed;/* cells the kernels actually sweep (incl. the redundant
                              half of self/first pairs, excl. lane padding)         */
    int64_t total_pairs;   /* unique pairs over all ranks                           */
    int64_t total_cells;   /* GCUPS numerator over all ranks                        */
    int32_t n_variants;    /* distinct kernel instantiations used                   */
    int32_t n_fallback32;  /* tiles routed to a 32-bit kernel (scores beyond int16)  */
} bgca_plan_info;

int bgca_version(void);
const char *bgca_last_error(void);

/* Engine lifetime.  One engine per device per process (one process per GPU
 * under torchrun).  Fails with BGCA_ERR_CUDA when no CUDA device is usable —
 * there is no CPU fallback. */
int bgca_create(int device, bgca_engine **out);
int bgca_destroy(bgca_engine *e);

/* Scoring parameters (Biopython PairwiseAligner semantics —
 * a gap of length k scores open + (k-1)*extend; defaults BLOSUM62, -12, -1).
 * sub: 24x24 row-major int8 in BGCA alphabet order; must be symmetric. */
int bgca_set_scoring(bgca_engine *e, const int8_t *sub, int32_t open, int32_t extend, int32_t mode);

/* K1 packer.  Consumes what BGCDomain.get_sequence() yields
 * (BGClib/BGClib.py:3240-3241: protein.sequence[ali_from:ali_to], upper-cased by
 * the setter at BGClib/BGClib.py:1903-1907), concatenated as ASCII, with
 * offsets[n+1].  type_of_seq / bgc_of_seq (nullable, length n) carry the domain
 * type key (alias-or-ID rule, BGCtoolkit.py:1511-1513) and the index of the
 * owning BGC (BGCProtein.parent_cluster_id, BGClib/BGClib.py:1835).
 * n_query > 0 declares the first n_query sequences "queries" and the rest a "database"
 * (the query-vs-database use BGClib/Readme.md:18 describes: "a new set of BGCs against a
 * precomputed database"); 0 = one undivided collection.
 * ascii_on_device != 0: `ascii` is a device pointer (offsets/types stay host).
 * The engine sorts by (type, length), encodes to 0..23, pads every sequence to
 * a 16-byte multiple with BGCA_PAD_CODE and keeps the packed batch in HBM. */
int bgca_pack(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n,
              const int32_t *type_of_seq, const int32_t *bgc_of_seq, int32_t n_bgc,
              int32_t n_query, int32_t ascii_on_device, void *stream);

/* The same for sequences given as SLICES of a residue buffer: sequence i is
 * ascii[span_start[i], span_start[i] + span_len[i]).  This is the array form of
 * BGCDomain.get_sequence() (BGClib/BGClib.py:3240-3241): the buffer holds every protein's
 * sequence once, a domain is (offset of its protein + ali_from, ali_to - ali_from), and the
 * slicing happens in K1 on the device instead of 10^5 Python string slices.  Slices may overlap
 * and leave gaps; each must lie inside the buffer. */
int bgca_pack_spans(bgca_engine *e, const uint8_t *ascii, int64_t ascii_bytes, const int64_t *span_start,
                    const int32_t *span_len, int32_t n, const int32_t *type_of_seq,
                    const int32_t *bgc_of_seq, int32_t n_bgc, int32_t n_query, int32_t ascii_on_device,
                    void *stream);

/* Persistent database ("a new set of BGCs against a precomputed database",
 * BGClib/Readme.md:18): keeps the batch packed by bgca_pack(..., n_query = 0) resident in HBM
 * as a database and appends n_new query sequences behind it (a later call replaces the previous
 * query set; only the queries are uploaded and encoded).  Afterwards the ORIGINAL index space is
 * [database | queries]: query i is sequence n_database + i, outputs of bgca_score() after a
 * cross plan are scores[n_new x n_database], rowbest[(n_database + n_new) x n_bgc_total],
 * selfsc[n_database + n_new].  type_of_seq / bgc_of_seq must be given exactly when they were
 * given for the database (same type-id space; query BGC indices continue after the
 * database's, n_bgc_total = all of them). */
int bgca_pack_append(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n_new,
                     const int32_t *type_of_seq, const int32_t *bgc_of_seq, int32_t n_bgc_total,
                     int32_t ascii_on_device, void *stream);

/* Introspection of the packed batch (host outputs; any pointer may be NULL).
 * codes_out must hold bgca_packed_bytes() bytes. */
int64_t bgca_packed_bytes(const bgca_engine *e);
int bgca_get_packed(const bgca_engine *e, uint8_t *codes_out, int64_t *offsets_sorted_out,
                    int32_t *order_out /* sorted -> original */, void *stream);

/* K6 tile scheduler.  Enumerates the upper-triangular pair matrix (i<=j, self
 * pairs included) as tiles, optionally only within a domain type, and assigns
 * tiles to `world` ranks by longest-processing-time on the exact cell count.
 * cross_only != 0 (needs n_query at pack time, or queries appended with bgca_pack_append):
 * only query x database pairs are planned, plus the self pair of every sequence (the
 * similarity is normalised with self scores); cross_only == 2 leaves out the self pairs of a
 * persistent database (the caller keeps its self scores from an earlier run).
 * bgca_plan keeps this rank's tiles on the device for bgca_score(). */
int bgca_plan(bgca_engine *e, int32_t same_type_only, int32_t cross_only, int32_t rank, int32_t world,
              bgca_plan_info *info);

/* Host-only planner (no CUDA calls): same algorithm on caller-supplied SORTED
 * lengths/types; writes up to `capacity` tiles of `rank` and the tile count.
 * Used by the CPU tests of the scheduler. */
int bgca_plan_host(const int32_t *len_sorted, const int32_t *type_sorted, const int32_t *role_sorted,
                   int32_t n, int32_t same_type_only, int32_t cross_only,
                   int32_t n_db_first /* > 0: sorted [0, n_db_first) is a persistent database */,
                   int32_t mode,
                   int32_t max_abs_sub, int32_t open, int32_t extend,
                   int32_t rank, int32_t world,
                   bgca_tile *tiles_out, int64_t capacity, int64_t *n_tiles_out,
                   bgca_plan_info *info);

/* K2/K3/K4 DP kernels + K5 epilogue over this rank's tiles.  Device outputs,
 * all optional (NULL to skip), all in ORIGINAL index space, caller-owned and
 * caller-initialised (zero; global mode: any value < every score):
 *   scores   int32[n*n]      s(i,j) and s(j,i) for every owned pair i<=j
 *            (after a cross_only plan: int32[n_query * (n - n_query)], row = query,
 *            column = database sequence)
 *   rowbest  int32[n*n_bgc]  atomicMax_{e in bgc b, type(e)==type(d)} s(d,e)
 *   selfsc   int32[n]        s(d,d)
 * type_check != 0 applies the same-type condition inside the epilogue (needed
 * when the plan covers all pairs but the aggregation is type-restricted). */
int bgca_score(bgca_engine *e, int32_t *scores, int32_t *rowbest, int32_t *selfsc,
               int32_t type_check, void *stream);

/* Explicit pair list through the 32-bit kernel (any lengths, any score range):
 * out[p] = s(pi[p], pj[p]); pi/pj/out are device pointers, ORIGINAL indices. */
int bgca_score_pairs32(bgca_engine *e, const int32_t *pi, const int32_t *pj, int64_t npairs,
                       int32_t *out, void *stream);

/* K7 — alignment statistics without traceback for an explicit pair list (SURVEY.md §8f-1;
 * the quantity a BiG-SCAPE-style distance on the same domains needs; the reference has no
 * such code).  Query = sequence pi[p], subject = sequence pj[p] (device pointers, ORIGINAL
 * indices); out is a device array of npairs records.  The alignment described is the
 * lexicographic maximum of (score, identities, -columns, q_start, s_start, -s_end, -q_end)
 * over all alignments of the affine-gap model in the engine's mode; score equals what
 * bgca_score() reports for the pair.  Coordinates are 0-based and end-exclusive like
 * BGCDomain.ali_from/ali_to (BGClib/BGClib.py:3222-3241); the empty local alignment
 * (score 0) has columns 0 and all four coordinates -1.  gaps = gap columns
 * (2*columns - query span - subject span); aligned residue pairs = columns - gaps.
 * Sequences longer than 32767 residues -> BGCA_ERR_RANGE. */
typedef struct bgca_alnstats {
    int32_t score, identities, columns, q_start, q_end, s_start, s_end, gaps;
} bgca_alnstats;
int bgca_pair_stats(bgca_engine *e, const int32_t *pi, const int32_t *pj, int64_t npairs,
                    bgca_alnstats *out, void *stream);

/* The same with the best score of every pair known (known_scores[p], device, what bgca_score()
 * wrote for the pair): in local mode the running best leaves the DP inner loop and only cells
 * that reach the known score are examined.  The result is exact whatever the hints are — a hint
 * that is too low costs time, one that is too high is detected and the list is run again without
 * hints.  Global mode ignores the hints. */
int bgca_pair_stats_hinted(bgca_engine *e, const int32_t *pi, const int32_t *pj, const int32_t *known_scores,
                           int64_t npairs, bgca_alnstats *out, void *stream);

/* Statistics for the pairs that are similar at all, chosen ON THE DEVICE from the score matrix
 * bgca_score() filled (device, int32[n*n], ORIGINAL index space; one undivided collection):
 * every pair i < j with scores[i][j] >= min_score [and the same domain type], in row-major
 * order.  bgca_select_pairs() counts them (*count; the caller sizes its outputs);
 * bgca_pair_stats_selected() then writes pairs_out[count][2] = (i, j) (device, nullable) and
 * out[count] (device) in that order, with the selected scores as hints.  The pair list never
 * visits the host.  One selection serves one statistics pass. */
int bgca_select_pairs(bgca_engine *e, const int32_t *scores, int32_t min_score, int32_t same_type_only,
                      int64_t *count, void *stream);
int bgca_pair_stats_selected(bgca_engine *e, const int32_t *scores, int32_t *pairs_out, bgca_alnstats *out,
                             void *stream);

/* K5 second stage (after ranks combined rowbest/selfsc with max):
 *   best[a,b] = sum_{d in a} rowbest[d,b]   int64[n_bgc*n_bgc]
 *   selfsum[a] = sum_{d in a} selfsc[d]     int64[n_bgc]
 *   sim[a,b] = (best[a,b]+best[b,a]) / (selfsum[a]+selfsum[b])  float64, 0/0 -> 0,
 *              sim[a,a] = 1 when selfsum[a] > 0           */
int bgca_reduce_bgc(bgca_engine *e, const int32_t *rowbest, const int32_t *selfsc,
                    int64_t *best, int64_t *selfsum, double *sim, void *stream);

/* End-to-end convenience with HOST buffers (the call bench.py's e2e leg times):
 * pack + plan(rank 0 of 1) + score + D2H.  scores_host: int32[n*n] (pageable or pinned).
 * Pipelined: the kernel runs are launched in ascending order of their queries and each
 * finished block of rows is copied to the host (copy engine, engine-owned pinned staging) and
 * put in place by up to 8 host threads while the remaining runs compute, so the call ends
 * within ~0.4 % of the device time of its DP kernels.  Synchronous; uses the legacy default
 * stream and the engine's own side streams. */
int bgca_score_all_host(bgca_engine *e, const uint8_t *ascii, const int64_t *offsets, int32_t n,
                        int32_t *scores_host);

/* Counters of the last bgca_score(): kernel launches issued, and the device
 * time of the DP kernels in ms (CUDA events on `stream`; valid after the
 * stream is synchronised). */
int bgca_last_launches(const bgca_engine *e);
int bgca_last_dp_ms(bgca_engine *e, float *ms);
/* Device time of the K1 kernel of the last bgca_pack() in ms (CUDA events on its stream). */
int bgca_last_pack_ms(bgca_engine *e, float *ms);

/* The placement rule of the packed batch, exposed for the host tests: the DP kernels advance
 * their token pointer with a 32-bit add on its low half, so the batch must lie inside one 4 GiB
 * address window.  Given an allocation that starts at `raw`, *start = where a buffer of `bytes`
 * bytes may begin: `raw` when [raw, raw + bytes) crosses no 4 GiB boundary, else the next
 * boundary (less than `bytes` away, so an allocation of 2 * bytes always has room).
 * BGCA_ERR_INVALID for more than 4 GiB (cannot happen: packed offsets are 32-bit). */
int bgca_window4g_start(uint64_t raw, uint64_t bytes, uint64_t *start);

/* Calibration microbenchmark: issue rate of the DPX/ALU ops
 * the DP kernels are made of.  Fills out[0..n) with lane-ops per clock per SM
 * for: 0 VIADDMNMX.S16x2, 1 VIMNMX3.S16x2, 2 VIMNMX.S16x2, 3 VIADD.16x2
 * (__vadd2), 4 IADD3, 5 IMAD, 6 PRMT, 7 VIADDMNMX (s32), 8 the bare instruction
 * stream of the packed local cell (5.5 ops per 2 cells, K = 26, the kernel's
 * occupancy; reported as lane-ops: cell pairs per clock per SM = value / 5.5),
 * 9 measured SM clock in MHz,
 * 10-13 two-op mixes (VIADDMNMX+VIADD.16x2, VIADDMNMX+IMAD, VIMNMX3+VIADD.16x2,
 * VIADD.16x2+IMAD: > 64 means the two ops issue to different pipes),
 * 14-17 dependent-chain latency in cycles (VIADDMNMX, VIMNMX3, VIADD.16x2, and
 * the 3-op F->H->Ho cell chain). n <= 20. */
int bgca_dpx_probe(int device, double *out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* BGCA_H */
