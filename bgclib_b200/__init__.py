"""bgclib_b200 — B200-native all-vs-all domain alignment scoring for BGClib
collections (Gotoh local/global, BLOSUM62), reduced on-device to BGC-pair
similarity.  Hand-written sm_100a CUDA behind a C ABI (include/bgca.h);
PyTorch only carries device memory, streams and torch.distributed.
"""
from .alphabet import ALPHABET, BLOSUM62  # noqa: F401
from .api import (BGCSimilarity, DomainPairScores, DomainPairStats, PackedDatabase, attach, bgc_similarity,  # noqa: F401
                  domain_pair_scores, domain_pair_stats, get_engine)
from .engine import Engine, dpx_probe, load_library, plan_host  # noqa: F401
from .extract import DomainBatch, DomainRecord, extract  # noqa: F401

__version__ = "0.1.0"
