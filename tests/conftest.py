import gzip
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def agx():
    import agxpkg
    return agxpkg.load()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def read_golden(name: str) -> bytes:
    p = GOLDEN / name
    if p.exists():
        return p.read_bytes()
    gz = GOLDEN / (name + ".gz")
    return gzip.decompress(gz.read_bytes())


def ref_scores(name: str):
    """Scores from a recorded reference stdout (tests/golden/*.ref.out)."""
    text = (GOLDEN / name).read_text()
    return [int(l.split()[1]) for l in text.splitlines() if l.startswith("Score:")]


@pytest.fixture(scope="session")
def gpu_lib(agx):
    """libagx bound to GPU 0.  GPU tests must run the CUDA path: no fallback, fail loudly."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    agx.capi.init(1)
    yield agx.capi
    agx.capi.shutdown()
