"""Error behaviour of the C ABI: return codes + shb_last_error text, no exceptions, nothing hashed on bad input."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shb():
    from seqhasher_b200 import binding
    binding.init(1)
    return binding


def test_capacity_and_offsets(shb):
    b = shb.Batch(1000, 10)
    with pytest.raises(shb.SeqhashError, match="capacity"):
        b.load(np.zeros(2000, np.uint8), np.array([0, 2000], dtype=np.int64))
    b.offsets[:3] = [0, 10, 5]                       # not monotone
    with pytest.raises(shb.SeqhashError, match="-5"):
        b.submit(["sha1"], 0, n=2)
    b.offsets[:2] = [0, 2000]                        # more bytes than the batch holds
    with pytest.raises(shb.SeqhashError, match="-3"):
        b.submit(["sha1"], 0, n=1)
    with pytest.raises(shb.SeqhashError, match="-3"):
        b.submit(["sha1"], 0, n=11)                  # more records than the batch holds
    b.offsets[0] = 1
    with pytest.raises(shb.SeqhashError, match="-5"):
        b.submit(["sha1"], 0, n=1)
    b.close()


def test_bad_algo_ids(shb):
    L = shb.lib()
    b = shb.Batch(100, 4)
    b.offsets[:2] = [0, 4]
    arr = (ctypes.c_int * 1)(12)
    assert L.shb_submit(b._h, 1, arr, 1, 0) == -2
    assert b"algo id" in L.shb_last_error()
    with pytest.raises(shb.SeqhashError):
        b.submit(["sha2"], 0, n=1)
    b.close()


def test_device_api_argument_checks(shb):
    import torch
    L = shb.lib()
    res = torch.zeros(64 + 16, dtype=torch.uint8, device="cuda")
    offs = torch.tensor([0, 64], dtype=torch.int64, device="cuda")
    out = torch.zeros(20, dtype=torch.uint8, device="cuda")
    ids = (ctypes.c_int * 1)(0)
    ptrs = (ctypes.c_void_p * 1)(out.data_ptr())
    # strip without scratch
    assert L.shb_hash_device(0, res.data_ptr(), offs.data_ptr(), 1, 64, ids, 1, 2, ptrs, None, None, None) == -2
    # misaligned residues
    assert L.shb_hash_device(0, res.data_ptr() + 1, offs.data_ptr(), 1, 63, ids, 1, 0, ptrs, None, None, None) == -2
    # device not selected
    assert L.shb_hash_device(99, res.data_ptr(), offs.data_ptr(), 1, 64, ids, 1, 0, ptrs, None, None, None) == -2
    # zero algorithms and zero records are fine
    assert L.shb_hash_device(0, res.data_ptr(), offs.data_ptr(), 1, 64, ids, 0, 0, ptrs, None, None, None) == 0
    assert L.shb_hash_device(0, res.data_ptr(), offs.data_ptr(), 0, 0, ids, 1, 0, ptrs, None, None, None) == 0
    torch.cuda.synchronize()


def test_textproc_capacity(shb):
    tp = shb.TextProc(1 << 12)
    with pytest.raises(shb.SeqhashError, match="capacity|-3"):
        tp.run((1 << 12) + 1, shb.FORMAT_FASTA, True, [0], 1, True, b"f")
    with pytest.raises(shb.SeqhashError, match="-2"):
        tp.run(10, 7, True, [0], 1, True, b"f")      # unknown format
    out, consumed, nrec, nempty = tp.run(0, shb.FORMAT_FASTA, True, [0], 1, True, b"f")
    assert out.size == 0 and nrec == 0
    tp.close()
