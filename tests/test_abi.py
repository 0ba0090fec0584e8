"""The C-ABI library loads without a GPU and exports every symbol include/seqhash_b200.h declares."""
import ctypes
import os
import re

from seqhasher_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "seqhash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(binding.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/seqhash_b200.h but not exported"
    assert sorted(binding.EXPORTS) == declared


def test_name_tables_need_no_gpu():
    L = binding.lib()
    assert [L.shb_algo_name(i).decode() for i in range(12)] == binding.ALGO_NAMES
    assert [L.shb_digest_size(i) for i in range(12)] == binding.DIGEST_SIZES
    assert [binding.algo_id(n) for n in binding.ALGO_NAMES] == list(range(12))
    # isValidHashType is exact-match (seqhasher.go:142-149)
    for bad in ("", "invalid", " md5", "SHA1", "sha1 "):
        assert binding.algo_id(bad) == -1
    assert L.shb_digest_size(12) == 0 and L.shb_algo_name(-1) is None
    assert sum(binding.DIGEST_SIZES) == 328


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    try:
        binding.init(0)
    except binding.SeqhashError as e:
        assert "no CUDA device" in str(e) or "failed" in str(e)
    else:
        raise AssertionError("shb_init succeeded without a GPU")
