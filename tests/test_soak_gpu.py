"""A short randomised soak of the C ABI against the oracle (tools/soak_gpu.py): random record counts, length
distributions, alphabets with whitespace / lower case / arbitrary bytes, flag combinations that force each kernel
family, random algorithm subsets.  It found the one parity bug of round 1 that the structured tests had missed."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [7, 2024])
def test_random_batches_match_the_oracle(seed):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_gpu.py"), "--iters", "150", "--seed", str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "0 mismatches" in p.stdout


@pytest.mark.gpu
def test_random_bgzf_files_match_the_plain_path():
    """tools/soak_bgzf.py: random FASTA / FASTQ / protein / tandem-repeat texts, every zlib level and strategy, member sizes
    64 B - 64 KiB, random chunking and look-ahead: the device-inflate path must print what the plain-text path prints."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak_bgzf.py"), "--iters", "60", "--seed", "11"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "0 mismatches" in p.stdout
