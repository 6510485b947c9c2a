import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no libbgca.so (built artefacts are git-ignored): build it once, as
    # __graft_entry__.build() does; nvcc cross-compiles without a GPU
    from bgclib_b200 import build as _build
    if not _build.LIB.exists():
        import shutil
        if shutil.which(_build.NVCC) or Path(_build.NVCC).exists():
            _build.build(verbose=True)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ks_records():
    """The 15 KS domains of examples/all_KS_domains.fasta as the reference parsed them."""
    return json.loads((GOLDEN / "ks_domains.json").read_text())


@pytest.fixture(scope="session")
def ks_golden():
    return json.loads((GOLDEN / "ks_scores.json").read_text())


@pytest.fixture(scope="session")
def af293():
    return json.loads((GOLDEN / "af293_arch.json").read_text())
