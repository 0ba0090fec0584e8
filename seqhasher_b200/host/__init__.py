"""Host-side mirror of the reference interface for the hash path (seqhasher.go): parse_flags, get_hash_func,
process_sequences, run — same names, argument meaning and error behaviour, with the per-record
strip/upper/hash statements (seqhasher.go:246-259) replaced by batched calls into libseqhash_b200.so."""
from .cli import Config, main, parse_flags, run  # noqa: F401  (python -m seqhasher_b200.host.cli runs cli.main)
from .hasher import GpuHasher, get_hash_func, is_valid_hash_type, SUPPORTED_HASH_TYPES  # noqa: F401
from .process import process_sequences, process_sequences_gpu  # noqa: F401
