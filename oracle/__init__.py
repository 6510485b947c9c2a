"""CPU oracle for the all-vs-all domain alignment path — TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product (``bgclib_b200``) never does; it fails loudly when its CUDA library is
missing.

PARITY UNPINNED BY THE REFERENCE: jorgecnavarrom/BGClib holds no aligner, no
call into ``Bio.Align`` and no test with scores.
What pins this oracle instead is stated in ``oracle/gotoh_oracle.c``.
"""
from .cbind import (  # noqa: F401
    ALPHABET, BLOSUM62, MODE_GLOBAL, MODE_LOCAL, all_pairs, bgc_similarity_np,
    STATS_DTYPE, encode, load, num_threads, pair_list, score, score_3state, stats_list,
)
