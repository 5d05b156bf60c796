"""pytest configuration: the `gpu` marker and shared fixtures."""
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


class Golden:
    """Read-only view of tests/golden/*.npz (vectors from the unmodified reference)."""

    def __init__(self, path):
        self._z = np.load(path)
        self.cases = sorted({k.split("/")[0] for k in self._z.files})

    def case(self, name):
        pre = name + "/"
        d = {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}
        if "bc_idx" in d:
            d["bc"] = {int(i): float(v) for i, v in zip(d["bc_idx"], d["bc_val"])}
        return d


@pytest.fixture(scope="session")
def golden():
    return Golden(ROOT / "tests" / "golden" / "ref_small.npz")


@pytest.fixture(scope="session")
def golden_big():
    p = ROOT / "tests" / "golden" / "ref_big.npz"
    if not p.exists():
        pytest.skip("ref_big.npz not generated")
    return Golden(p)
