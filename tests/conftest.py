import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Fresh checkout: the built artefacts are git-ignored, so build them once (nvcc cross-compiles without a GPU)."""
    import subprocess
    need = [os.path.join(ROOT, "seqhasher_b200", "libseqhash_b200.so"), os.path.join(ROOT, "seqhasher_b200", "bin", "seqhasher-b200"),
            os.path.join(ROOT, "seqhasher_b200", "bin", "shb-zcat"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so"), os.path.join(ROOT, "oracle", "_build", "cpu_bench")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", os.path.join(ROOT, "seqhasher_b200", "csrc")], check=True, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
