import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native pieces are built in-tree (git-ignored); rebuild when missing or older than their sources
    import __graft_entry__ as entry
    entry.build()


@pytest.fixture(autouse=True)
def reset_globals():
    """Same hygiene as the reference's tests/conftest.py:12-31: the partition module keeps run state in globals."""
    yield
    try:
        from ppanggolin_b200.nem import partition as part
        part.samples = []
    except Exception:
        pass
