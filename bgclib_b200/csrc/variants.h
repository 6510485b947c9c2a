// The (G, K) instantiations of the packed DP kernel — one list shared by the
// kernel translation units and the tile planner.  G = lanes per alignment,
// K = query rows per lane; a variant covers queries up to G*K residues.
#pragma once

#define BGCA_K_LIST(X, G) \
    X(G, 2) X(G, 3) X(G, 4) X(G, 5) X(G, 6) X(G, 7) X(G, 8) X(G, 9) X(G, 10) X(G, 11) X(G, 12) \
    X(G, 13) X(G, 14) X(G, 15) X(G, 16) X(G, 17) X(G, 18) X(G, 19) X(G, 20) X(G, 21) X(G, 22)  \
    X(G, 23) X(G, 24) X(G, 25) X(G, 26) X(G, 27) X(G, 28) X(G, 29) X(G, 30) X(G, 31) X(G, 32)

// Whole-warp variants with more than 32 rows per lane: queries of 1025..2048 residues (whole
// proteins) in ONE pass, at one CTA per SM (the profile alone takes up to 213 KB of shared memory).
#define BGCA_K_LIST_WIDE(X, G) X(G, 40) X(G, 48) X(G, 56) X(G, 64)

#define BGCA_VARIANTS(X) \
    BGCA_K_LIST(X, 4) BGCA_K_LIST(X, 8) BGCA_K_LIST(X, 16) BGCA_K_LIST(X, 32) BGCA_K_LIST_WIDE(X, 32)

#define BGCA_VARIANT_ID(G, K) (((G) << 8) | (K))
