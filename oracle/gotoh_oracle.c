/*
 * oracle/gotoh_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the scoring step that the B200 engine accelerates:
 * affine-gap pairwise alignment *score* (Gotoh), local and global, over a
 * 24-letter protein alphabet with an integer substitution matrix (BLOSUM62).
 *
 * PARITY UNPINNED BY THE REFERENCE: jorgecnavarrom/BGClib contains no aligner
 * (only the intent at BGClib/Readme.md:16-18 and the
 * commented-out --comparative flag at BGCtoolkit.py:157-161).  The semantics
 * restated here are those of Biopython 1.83 (pyproject.toml:13)
 * Bio.Align.PairwiseAligner.score() with substitution_matrix=BLOSUM62 and
 * open_gap_score/extend_gap_score set: a gap of length k
 * scores open + (k-1)*extend, end gaps cost the same as internal gaps in
 * global mode, local mode clamps at 0 and returns the best cell.  Biopython
 * is not installed in this image, so the pin is (i) the four SHA-256 digests
 * of SURVEY.md App. B on examples/all_KS_domains.fasta and (ii) agreement of
 * two independent formulations: the classic H/E/F recurrence (this file,
 * bgco_score) and the 3-state M/Ix/Iy recurrence Biopython uses
 * (bgco_score_3state below and oracle/gotoh_py.py).
 *
 * Input semantics follow the reference: the unit aligned is
 * BGCDomain.get_sequence() == protein.sequence[ali_from:ali_to]
 * (BGClib/BGClib.py:3240-3241), upper-cased by the sequence setter
 * (BGClib/BGClib.py:1903-1907).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BGCO_MODE_LOCAL 0
#define BGCO_MODE_GLOBAL 1
#define BGCO_NSYM 24
#define BGCO_NEG (-(1 << 28))

static inline int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t max3(int32_t a, int32_t b, int32_t c) { return max2(max2(a, b), c); }

/* Classic Gotoh H/E/F, one row of H and E kept; O(lb) memory.
 * a: la codes (rows), b: lb codes (columns), sub: 24x24 row-major int8.
 * open/extend are the (negative) scores of the first / each further gapped
 * residue.  work: caller scratch of 2*(lb+1) int32. */
int32_t bgco_score_w(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb,
                     const int8_t *sub, int32_t open, int32_t extend, int32_t mode,
                     int32_t *work)
{
    int32_t *H = work;          /* H[j] : row i-1 on entry, row i on exit   */
    int32_t *E = work + lb + 1; /* E[j] : gap-in-a state ending at (i, j), carried down rows */
    int32_t best = 0;
    const int local = (mode == BGCO_MODE_LOCAL);

    H[0] = 0;
    for (int32_t j = 1; j <= lb; ++j) {
        H[j] = local ? 0 : open + (j - 1) * extend;
        E[j] = BGCO_NEG;
    }
    for (int32_t i = 1; i <= la; ++i) {
        const int8_t *srow = sub + (size_t)a[i - 1] * BGCO_NSYM;
        int32_t diag = H[0];
        int32_t left = local ? 0 : open + (i - 1) * extend; /* H[i][0] */
        int32_t F = BGCO_NEG;                               /* gap-in-b state, carried along the row */
        H[0] = left;
        for (int32_t j = 1; j <= lb; ++j) {
            int32_t e = max2(E[j] + extend, H[j] + open); /* H[j] still holds row i-1 */
            F = max2(F + extend, left + open);
            int32_t h = max3(diag + srow[b[j - 1]], e, F);
            if (local) {
                if (h < 0) h = 0;
                if (h > best) best = h;
            }
            diag = H[j];
            H[j] = h;
            E[j] = e;
            left = h;
        }
    }
    return local ? best : H[lb];
}

int32_t bgco_score(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb,
                   const int8_t *sub, int32_t open, int32_t extend, int32_t mode)
{
    int32_t *work = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(lb + 1));
    int32_t s = bgco_score_w(a, la, b, lb, sub, open, extend, mode, work);
    free(work);
    return s;
}

/* Independent second formulation: three explicit states as in Biopython's
 * Gotoh (M = ends in a substitution, Ix = ends in a gap in b, Iy = ends in a
 * gap in a), with Ix<->Iy transitions allowed at `open` cost.  Local mode:
 * every state clamped at 0, result = max over M.  Full O(la*lb) tables —
 * deliberately naive; used only to cross-check bgco_score on small inputs. */
int32_t bgco_score_3state(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb,
                          const int8_t *sub, int32_t open, int32_t extend, int32_t mode)
{
    const int local = (mode == BGCO_MODE_LOCAL);
    size_t W = (size_t)lb + 1, n = ((size_t)la + 1) * W;
    int32_t *M = (int32_t *)malloc(3 * n * sizeof(int32_t));
    int32_t *Ix = M + n, *Iy = Ix + n;
    int32_t best = 0;
    for (size_t t = 0; t < 3 * n; ++t) M[t] = BGCO_NEG;
    M[0] = 0;
    if (local) {
        for (int32_t i = 0; i <= la; ++i) M[i * W] = Ix[i * W] = Iy[i * W] = 0;
        for (int32_t j = 0; j <= lb; ++j) M[j] = Ix[j] = Iy[j] = 0;
    } else {
        for (int32_t i = 1; i <= la; ++i) Iy[i * W] = open + (i - 1) * extend;
        for (int32_t j = 1; j <= lb; ++j) Ix[j] = open + (j - 1) * extend;
    }
    for (int32_t i = 1; i <= la; ++i) {
        for (int32_t j = 1; j <= lb; ++j) {
            size_t c = i * W + j, d = (i - 1) * W + (j - 1), l = i * W + (j - 1), u = (i - 1) * W + j;
            int32_t m = max3(M[d], Ix[d], Iy[d]) + sub[a[i - 1] * BGCO_NSYM + b[j - 1]];
            int32_t x = max3(M[l] + open, Ix[l] + extend, Iy[l] + open);
            int32_t y = max3(M[u] + open, Iy[u] + extend, Ix[u] + open);
            if (local) {
                if (m < 0) m = 0;
                if (x < 0) x = 0;
                if (y < 0) y = 0;
                if (m > best) best = m;
            }
            M[c] = m; Ix[c] = x; Iy[c] = y;
        }
    }
    int32_t r = local ? best : max3(M[n - 1], Ix[n - 1], Iy[n - 1]);
    free(M);
    return r;
}

/* All pairs i<=j of n sequences (codes concatenated, offsets[n+1]); writes the
 * full symmetric n x n int32 matrix.  nthreads<=0 -> OpenMP default. */
void bgco_all_pairs(const uint8_t *codes, const int64_t *offsets, int32_t n,
                    const int8_t *sub, int32_t open, int32_t extend, int32_t mode,
                    int32_t *out, int32_t nthreads)
{
    int32_t maxlen = 0;
    for (int32_t i = 0; i < n; ++i) {
        int32_t l = (int32_t)(offsets[i + 1] - offsets[i]);
        if (l > maxlen) maxlen = l;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int32_t *work = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(maxlen + 1));
#pragma omp for schedule(dynamic, 1)
        for (int32_t i = 0; i < n; ++i) {
            const uint8_t *a = codes + offsets[i];
            int32_t la = (int32_t)(offsets[i + 1] - offsets[i]);
            for (int32_t j = i; j < n; ++j) {
                const uint8_t *b = codes + offsets[j];
                int32_t lb = (int32_t)(offsets[j + 1] - offsets[j]);
                int32_t s = bgco_score_w(a, la, b, lb, sub, open, extend, mode, work);
                out[(size_t)i * n + j] = s;
                out[(size_t)j * n + i] = s;
            }
        }
        free(work);
    }
}

/* Explicit pair list: out[p] = score(seq[pi[p]], seq[pj[p]]). */
void bgco_pair_list(const uint8_t *codes, const int64_t *offsets, int32_t n,
                    const int32_t *pi, const int32_t *pj, int64_t npairs,
                    const int8_t *sub, int32_t open, int32_t extend, int32_t mode,
                    int32_t *out, int32_t nthreads)
{
    int32_t maxlen = 0;
    for (int32_t i = 0; i < n; ++i) {
        int32_t l = (int32_t)(offsets[i + 1] - offsets[i]);
        if (l > maxlen) maxlen = l;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        int32_t *work = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(maxlen + 1));
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < npairs; ++p) {
            int32_t i = pi[p], j = pj[p];
            out[p] = bgco_score_w(codes + offsets[i], (int32_t)(offsets[i + 1] - offsets[i]),
                                  codes + offsets[j], (int32_t)(offsets[j + 1] - offsets[j]),
                                  sub, open, extend, mode, work);
        }
        free(work);
    }
}

int32_t bgco_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
