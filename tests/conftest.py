"""pytest wiring: the `gpu` marker, repo-root imports, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # a fresh checkout has no binaries (they are git-ignored): build the library and the checkers once (nvcc cross-compiles)
    lib = os.path.join(ROOT, "depecher_b200", "lib", "libdepecher_b200.so")
    if not os.path.exists(lib) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import subprocess
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True)


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_vectors():
    return dict(np.load(os.path.join(GOLDEN, "ref_vectors.npz")))


@pytest.fixture(scope="session")
def testdata():
    z = np.load(os.path.join(GOLDEN, "testdata.npz"))
    out = {k: z[k] for k in z.files}
    out["X"] = out.pop("X_1e4").astype(np.float64) / 1e4
    return out


@pytest.fixture(scope="session")
def engine():
    """The CUDA library through its C ABI (fails loudly without a device)."""
    import depecher_b200 as dp
    return dp.Engine()
