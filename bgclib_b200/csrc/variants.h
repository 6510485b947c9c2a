// The (G, K) instantiations of the packed DP kernel — one list shared by the
// kernel translation units and the tile planner.  G = lanes per alignment,
// K = query rows per lane; a variant covers queries up to G*K residues.
#pragma once

#define BGCA_K_LIST(X, G) \
    X(G, 2) X(G, 3) X(G, 4) X(G, 5) X(G, 6) X(G, 7) X(G, 8) X(G, 9) X(G, 10) X(G, 11) X(G, 12) \
    X(G, 13) X(G, 14) X(G, 15) X(G, 16) X(G, 17) X(G, 18) X(G, 19) X(G, 20) X(G, 21) X(G, 22)  \
    X(G, 23) X(G, 24) X(G, 25) X(G, 26) X(G, 27) X(G, 28) X(G, 29) X(G, 30) X(G, 31) X(G, 32)

// Multi-pass whole-warp variants: a query of more than 1024 residues is swept in
// ceil(L / (32*K)) passes of 32*K rows (K = 17..32 covers every length), see gotoh_pair16.cuh.
#define BGCA_K_LIST_MP(X, G) \
    X(G, 17) X(G, 18) X(G, 19) X(G, 20) X(G, 21) X(G, 22) X(G, 23) X(G, 24) X(G, 25) X(G, 26) \
    X(G, 27) X(G, 28) X(G, 29) X(G, 30) X(G, 31) X(G, 32)

#define BGCA_VARIANTS(X) \
    BGCA_K_LIST(X, 1) BGCA_K_LIST(X, 2) BGCA_K_LIST(X, 4) BGCA_K_LIST(X, 8) BGCA_K_LIST(X, 16) BGCA_K_LIST(X, 32)

#define BGCA_VARIANT_MP 0x10000   /* flag in bgca_tile.variant: multi-pass instantiation */
#define BGCA_VARIANT_W32 0x20000  /* with MP: one alignment per 32-bit word (single-query tiles) */
#define BGCA_W32_MIN_ROWS 272     /* shorter queries leave more than half of a 17-row pass empty */
#define BGCA_VARIANT_ID(G, K) (((G) << 8) | (K))
#define BGCA_ONE_PASS_ROWS 1024   /* longest query of a single-pass variant (G = 32, K = 32) */
