"""ctypes binding of libseqhash_b200.so (the C ABI in include/seqhash_b200.h).

This is the only way Python reaches the CUDA kernels.  There is no fallback: if the shared
library is missing the import of any compute entry point raises, and every call fails with the
library's own error when no sm_100 GPU is usable.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHB_LIB") or os.path.join(_PKG, "libseqhash_b200.so")   # SHB_LIB: A/B-test another build

ALGO_NAMES = ["sha1", "sha3", "md5", "xxhash", "xxh3", "cityhash", "murmur3", "nthash", "blake3", "k12",
              "rapidhash", "rapidhash32"]
DIGEST_SIZES = [20, 64, 16, 8, 8, 16, 16, 8, 32, 128, 8, 4]
FLAG_UPPER = 1
FLAG_STRIP = 2
FLAG_EMIT_SEQ = 4
FLAG_CHECK_ASCII = 8
FLAG_FORCE_DIRECT = 0x100
FLAG_NO_LONG_PATH = 0x200
FLAG_NO_LENGTH_SORT = 0x400
FLAG_NO_FUSED_B3NT = 0x800

# every symbol include/seqhash_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "shb_init", "shb_shutdown", "shb_last_error", "shb_version", "shb_algo_id", "shb_algo_name",
    "shb_digest_size", "shb_batch_create", "shb_batch_create_on", "shb_batch_destroy", "shb_batch_residues",
    "shb_batch_offsets", "shb_submit", "shb_wait", "shb_digests", "shb_out_lengths", "shb_out_offsets",
    "shb_out_residues", "shb_batch_nonascii", "shb_textproc_nonascii", "shb_scratch_bytes", "shb_hash_device", "shb_launch_count", "shb_synth_fill",
    "shb_probe_int_peak", "shb_probe_last_checksum", "shb_probe_h2d", "shb_debug_oob", "shb_textproc_create", "shb_textproc_destroy",
    "shb_textproc_input", "shb_textproc_capacity", "shb_textproc_run", "shb_textproc_output", "shb_textproc_consumed",
    "shb_textproc_records", "shb_textproc_empty_records", "shb_textproc_last_ms",
    "shb_textproc_bgzf_input", "shb_textproc_run_bgzf", "shb_textproc_resume", "shb_textproc_pending", "shb_textproc_last_inflate_ms",
]
PARTIAL_FORMAT = 1   # shb_textproc_run: SHB_PARTIAL_FORMAT
NEED_MORE = 2        # shb_textproc_run_bgzf: SHB_NEED_MORE
ERR_INFLATE = -7
FORMAT_FASTA = 1
FORMAT_FASTQ = 2


class SeqhashError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeqhashError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). seqhasher_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    L.shb_init.argtypes = [c.c_int, c.c_int]; L.shb_init.restype = c.c_int
    L.shb_shutdown.argtypes = []; L.shb_shutdown.restype = None
    L.shb_last_error.argtypes = []; L.shb_last_error.restype = c.c_char_p
    L.shb_version.argtypes = []; L.shb_version.restype = c.c_char_p
    L.shb_algo_id.argtypes = [c.c_char_p]; L.shb_algo_id.restype = c.c_int
    L.shb_algo_name.argtypes = [c.c_int]; L.shb_algo_name.restype = c.c_char_p
    L.shb_digest_size.argtypes = [c.c_int]; L.shb_digest_size.restype = c.c_size_t
    L.shb_batch_create.argtypes = [c.c_size_t, c.c_size_t]; L.shb_batch_create.restype = c.c_void_p
    L.shb_batch_create_on.argtypes = [c.c_int, c.c_size_t, c.c_size_t]; L.shb_batch_create_on.restype = c.c_void_p
    L.shb_batch_destroy.argtypes = [c.c_void_p]; L.shb_batch_destroy.restype = None
    L.shb_batch_residues.argtypes = [c.c_void_p]; L.shb_batch_residues.restype = c.c_void_p
    L.shb_batch_offsets.argtypes = [c.c_void_p]; L.shb_batch_offsets.restype = c.c_void_p
    L.shb_submit.argtypes = [c.c_void_p, c.c_size_t, c.POINTER(c.c_int), c.c_int, c.c_uint]; L.shb_submit.restype = c.c_int
    L.shb_wait.argtypes = [c.c_void_p]; L.shb_wait.restype = c.c_int
    L.shb_digests.argtypes = [c.c_void_p, c.c_int]; L.shb_digests.restype = c.c_void_p
    L.shb_out_lengths.argtypes = [c.c_void_p]; L.shb_out_lengths.restype = c.c_void_p
    L.shb_out_offsets.argtypes = [c.c_void_p]; L.shb_out_offsets.restype = c.c_void_p
    L.shb_out_residues.argtypes = [c.c_void_p]; L.shb_out_residues.restype = c.c_void_p
    L.shb_batch_nonascii.argtypes = [c.c_void_p]; L.shb_batch_nonascii.restype = c.c_size_t
    L.shb_textproc_nonascii.argtypes = [c.c_void_p]; L.shb_textproc_nonascii.restype = c.c_size_t
    L.shb_scratch_bytes.argtypes = [c.c_size_t, c.c_size_t]; L.shb_scratch_bytes.restype = c.c_size_t
    L.shb_hash_device.argtypes = [c.c_int, c.c_void_p, c.c_void_p, c.c_size_t, c.c_size_t, c.POINTER(c.c_int), c.c_int,
                                  c.c_uint, c.POINTER(c.c_void_p), c.c_void_p, c.c_void_p, c.c_void_p]
    L.shb_hash_device.restype = c.c_int
    L.shb_launch_count.argtypes = []; L.shb_launch_count.restype = c.c_ulonglong
    L.shb_synth_fill.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_uint64, c.c_uint, c.c_uint, c.c_void_p]
    L.shb_synth_fill.restype = c.c_int
    L.shb_probe_int_peak.argtypes = [c.c_int, c.c_int, c.c_int]; L.shb_probe_int_peak.restype = c.c_double
    L.shb_probe_last_checksum.argtypes = []; L.shb_probe_last_checksum.restype = c.c_uint
    L.shb_probe_h2d.argtypes = [c.c_int, c.c_size_t, c.c_size_t, c.c_double, c.c_int]; L.shb_probe_h2d.restype = c.c_double
    L.shb_debug_oob.argtypes = []; L.shb_debug_oob.restype = c.c_longlong
    L.shb_textproc_create.argtypes = [c.c_int, c.c_size_t]; L.shb_textproc_create.restype = c.c_void_p
    L.shb_textproc_destroy.argtypes = [c.c_void_p]; L.shb_textproc_destroy.restype = None
    L.shb_textproc_input.argtypes = [c.c_void_p]; L.shb_textproc_input.restype = c.c_void_p
    L.shb_textproc_capacity.argtypes = [c.c_void_p]; L.shb_textproc_capacity.restype = c.c_size_t
    L.shb_textproc_run.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.c_int, c.POINTER(c.c_int), c.c_int, c.c_uint, c.c_int,
                                   c.c_char_p]
    L.shb_textproc_run.restype = c.c_int
    L.shb_textproc_bgzf_input.argtypes = [c.c_void_p, c.c_size_t]; L.shb_textproc_bgzf_input.restype = c.c_void_p
    L.shb_textproc_run_bgzf.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_size_t, c.c_longlong, c.c_int, c.c_int,
                                        c.POINTER(c.c_int), c.c_int, c.c_uint, c.c_int, c.c_char_p]
    L.shb_textproc_run_bgzf.restype = c.c_int
    L.shb_textproc_resume.argtypes = [c.c_void_p, c.c_int, c.POINTER(c.c_int), c.c_int, c.c_uint, c.c_int, c.c_char_p]
    L.shb_textproc_resume.restype = c.c_int
    L.shb_textproc_pending.argtypes = [c.c_void_p]; L.shb_textproc_pending.restype = c.c_int
    L.shb_textproc_last_inflate_ms.argtypes = [c.c_void_p]; L.shb_textproc_last_inflate_ms.restype = c.c_float
    L.shb_textproc_output.argtypes = [c.c_void_p, c.POINTER(c.c_size_t)]; L.shb_textproc_output.restype = c.c_void_p
    L.shb_textproc_last_ms.argtypes = [c.c_void_p, c.POINTER(c.c_float)]; L.shb_textproc_last_ms.restype = None
    for nm in ("shb_textproc_consumed", "shb_textproc_records", "shb_textproc_empty_records"):
        getattr(L, nm).argtypes = [c.c_void_p]; getattr(L, nm).restype = c.c_size_t
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc < 0:
        raise SeqhashError(f"{what} failed ({rc}): {lib().shb_last_error().decode()}")


def init(n_gpus: int = 0) -> int:
    rc = lib().shb_init(n_gpus, 0)
    _check(rc, "shb_init")
    return rc


def algo_id(name: str) -> int:
    return lib().shb_algo_id(name.encode())


def _ids(algos) -> list[int]:
    out = []
    for a in algos:
        i = a if isinstance(a, int) else algo_id(a)
        if i < 0 or i >= len(ALGO_NAMES):
            raise SeqhashError(f"unknown hash type {a!r}")
        out.append(i)
    return out


class Batch:
    """One shb_batch: pinned staging + device buffers + a stream.  Fill, submit, wait, read."""

    def __init__(self, max_bytes: int, max_records: int, device: int = 0):
        L = lib()
        self._h = L.shb_batch_create_on(device, max_bytes, max_records)
        if not self._h:
            raise SeqhashError(f"shb_batch_create failed: {L.shb_last_error().decode()}")
        self.max_bytes, self.max_records = max_bytes, max_records
        rp = L.shb_batch_residues(self._h)
        op = L.shb_batch_offsets(self._h)
        self.residues = np.ctypeslib.as_array(ctypes.cast(rp, ctypes.POINTER(ctypes.c_uint8)), shape=(max_bytes + 16,))
        self.offsets = np.ctypeslib.as_array(ctypes.cast(op, ctypes.POINTER(ctypes.c_int64)), shape=(max_records + 1,))
        self.n = 0
        self.algos: list[int] = []
        self.flags = 0

    def close(self) -> None:
        if self._h:
            lib().shb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, residues: np.ndarray, offsets: np.ndarray) -> None:
        """Copy a CSR batch into the pinned staging buffers."""
        n = len(offsets) - 1
        total = int(offsets[-1]) if n >= 0 and len(offsets) else 0
        if n > self.max_records or total > self.max_bytes:
            raise SeqhashError("batch exceeds capacity")
        self.residues[:total] = residues[:total]
        self.offsets[: n + 1] = offsets
        self.n = n

    def submit(self, algos, flags: int, n: int | None = None) -> None:
        if n is not None:
            self.n = n
        self.algos = _ids(algos)
        self.flags = flags
        arr = (ctypes.c_int * len(self.algos))(*self.algos)
        _check(lib().shb_submit(self._h, self.n, arr, len(self.algos), flags), "shb_submit")

    def wait(self) -> None:
        _check(lib().shb_wait(self._h), "shb_wait")

    def digests(self, slot: int) -> np.ndarray:
        p = lib().shb_digests(self._h, slot)
        ds = DIGEST_SIZES[self.algos[slot]]
        if self.n == 0:
            return np.zeros((0, ds), dtype=np.uint8)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(self.n, ds))

    def nonascii(self) -> int:
        """residue bytes >= 0x80 of the last batch (submit with FLAG_CHECK_ASCII)"""
        return lib().shb_batch_nonascii(self._h)

    def out_lengths(self) -> np.ndarray:
        p = lib().shb_out_lengths(self._h)
        if self.n == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int64)), shape=(self.n,))

    def out_sequences(self):
        """(residues, offsets) of the normalised sequences; needs FLAG_EMIT_SEQ."""
        po, pr = lib().shb_out_offsets(self._h), lib().shb_out_residues(self._h)
        if not po or not pr:
            raise SeqhashError("submit with FLAG_EMIT_SEQ to get the normalised sequences")
        offs = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_int64)), shape=(self.n + 1,))
        total = int(offs[self.n])
        res = np.ctypeslib.as_array(ctypes.cast(pr, ctypes.POINTER(ctypes.c_uint8)), shape=(max(total, 1),))[:total]
        return res, offs


def hash_batch(residues: np.ndarray, offsets: np.ndarray, algos, flags: int, device: int = 0, emit: bool = False):
    """Convenience one-shot through the host-buffer path: returns ([n x size uint8], lengths[, (res, offs)])."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    b = Batch(max(int(offsets[-1]), 1), max(n, 1), device)
    try:
        b.load(residues, offsets)
        b.submit(algos, flags | (FLAG_EMIT_SEQ if emit else 0))
        b.wait()
        outs = [b.digests(k).copy() for k in range(len(b.algos))]
        lens = b.out_lengths().copy()
        if emit:
            r, o = b.out_sequences()
            return outs, lens, (r.copy(), o.copy())
        return outs, lens
    finally:
        b.close()


def hash_device(d_residues, d_offsets, n: int, total_bytes: int, algos, flags: int, d_digests, d_out_lengths=None,
                d_scratch=None, stream: int = 0, device: int = 0) -> None:
    """Device-resident path: arguments are raw device pointers (ints), e.g. torch tensors' data_ptr()."""
    ids = _ids(algos)
    arr = (ctypes.c_int * len(ids))(*ids)
    ptrs = (ctypes.c_void_p * len(ids))(*d_digests)
    _check(lib().shb_hash_device(device, d_residues, d_offsets, n, total_bytes, arr, len(ids), flags, ptrs,
                                 d_out_lengths, d_scratch, stream), "shb_hash_device")


def scratch_bytes(total_bytes: int, n: int) -> int:
    return lib().shb_scratch_bytes(total_bytes, n)


def synth_fill(d_residues: int, total_bytes: int, seed: int, lower_permille: int = 0, n_permille: int = 0,
               stream: int = 0, device: int = 0) -> None:
    _check(lib().shb_synth_fill(device, d_residues, total_bytes, seed & 0xFFFFFFFFFFFFFFFF, lower_permille, n_permille,
                                stream), "shb_synth_fill")


def launch_count() -> int:
    return lib().shb_launch_count()


def probe_int_peak(kind: int, iters: int = 2000, device: int = 0) -> float:
    return lib().shb_probe_int_peak(device, kind, iters)


def probe_h2d(device: int = 0, nbytes: int = 1 << 30, piece: int = 64 << 20, seconds: float = 2.0, write_combined: bool = False) -> float:
    """pinned host -> device copy rate of this process in GB/s"""
    return lib().shb_probe_h2d(device, nbytes, piece, seconds, 1 if write_combined else 0)


def probe_last_checksum() -> int:
    return lib().shb_probe_last_checksum()


class TextProc:
    """shb_textproc: the whole record loop of processSequences on the GPU for chunks of decompressed text."""

    def __init__(self, max_text_bytes: int = 64 << 20, device: int = 0):
        L = lib()
        self._h = L.shb_textproc_create(device, max_text_bytes)
        if not self._h:
            raise SeqhashError(f"shb_textproc_create failed: {L.shb_last_error().decode()}")
        self.capacity = L.shb_textproc_capacity(self._h)
        p = L.shb_textproc_input(self._h)
        self.input = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(self.capacity + 32,))

    def close(self) -> None:
        if self._h:
            lib().shb_textproc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_bytes: int, fmt: int, is_final: bool, algos, flags: int, headers_only: bool, label: bytes | None):
        """Returns (output bytes view, consumed, n_records, n_empty).  The view is valid until the next run."""
        L = lib()
        ids = list(algos)
        arr = (ctypes.c_int * max(len(ids), 1))(*ids)
        rc = L.shb_textproc_run(self._h, n_bytes, fmt, 1 if is_final else 0, arr, len(ids), flags,
                                1 if headers_only else 0, label)
        _check(rc, "shb_textproc_run")
        self.nonascii = L.shb_textproc_nonascii(self._h)   # sequence bytes >= 0x80 in this chunk
        self.partial_format = rc == PARTIAL_FORMAT   # the text at `consumed` is malformed; the output holds what precedes it
        nb = ctypes.c_size_t(0)
        p = L.shb_textproc_output(self._h, ctypes.byref(nb))
        # zero-copy view of the pinned output buffer (valid until the next run)
        out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nb.value,)) if nb.value else np.zeros(0, np.uint8)
        return out, L.shb_textproc_consumed(self._h), L.shb_textproc_records(self._h), L.shb_textproc_empty_records(self._h)

    def _results(self, rc: int):
        L = lib()
        self.nonascii = L.shb_textproc_nonascii(self._h)
        self.partial_format = rc == PARTIAL_FORMAT
        nb = ctypes.c_size_t(0)
        p = L.shb_textproc_output(self._h, ctypes.byref(nb))
        out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nb.value,)) if nb.value else np.zeros(0, np.uint8)
        return out, L.shb_textproc_consumed(self._h), L.shb_textproc_records(self._h), L.shb_textproc_empty_records(self._h)

    def bgzf_input(self, comp_capacity: int) -> np.ndarray:
        """Pinned staging for the bytes of BGZF members (shb_textproc_bgzf_input), at least comp_capacity bytes."""
        p = lib().shb_textproc_bgzf_input(self._h, comp_capacity)
        if not p:
            raise SeqhashError(f"shb_textproc_bgzf_input failed: {lib().shb_last_error().decode()}")
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(comp_capacity,))

    def run_bgzf(self, n_comp_bytes: int, members: np.ndarray, n_own: int, n_extra: int, first_start: int, at_eof: bool,
                 fmt: int, algos, flags: int, headers_only: bool, label: bytes | None):
        """shb_textproc_run_bgzf: members = uint64 array (n_own + n_extra, 3): offset of the raw DEFLATE data in the staging
        buffer, its length, ISIZE | CRC32 << 32.  Returns None for SHB_NEED_MORE (stage more extra members and call again),
        else (output view, consumed, n_records, n_empty); the rest of a partly consumed chunk (self.pending()) goes through
        resume()."""
        ids = list(algos)
        arr = (ctypes.c_int * max(len(ids), 1))(*ids)
        m = np.ascontiguousarray(members, dtype=np.uint64).reshape(-1, 3)
        assert len(m) == n_own + n_extra
        rc = lib().shb_textproc_run_bgzf(self._h, n_comp_bytes, m.ctypes.data if len(m) else None, n_own, n_extra, first_start,
                                         1 if at_eof else 0, fmt, arr, len(ids), flags, 1 if headers_only else 0, label)
        _check(rc, "shb_textproc_run_bgzf")
        if rc == NEED_MORE:
            return None
        return self._results(rc)

    def last_inflate_ms(self) -> float:
        return float(lib().shb_textproc_last_inflate_ms(self._h))

    def pending(self) -> bool:
        return bool(lib().shb_textproc_pending(self._h))

    def resume(self, fmt: int, algos, flags: int, headers_only: bool, label: bytes | None):
        ids = list(algos)
        arr = (ctypes.c_int * max(len(ids), 1))(*ids)
        rc = lib().shb_textproc_resume(self._h, fmt, arr, len(ids), flags, 1 if headers_only else 0, label)
        _check(rc, "shb_textproc_resume")
        return self._results(rc)

    def last_ms(self) -> dict:
        arr = (ctypes.c_float * 5)()
        lib().shb_textproc_last_ms(self._h, arr)
        return dict(zip(["h2d", "split", "hash", "format", "d2h"], [float(x) for x in arr]))
