"""processSequences (seqhasher.go:210-286) with the per-record hash statements replaced by batched GPU calls.
Header layout, --headersonly and FASTA/FASTQ formatting follow seqhasher.go:261-282."""
from __future__ import annotations

import io

from .fastx import FastxError, read_records
from .hasher import GpuHasher, shared_hasher

BATCH_RECORDS = 1 << 18
BATCH_BYTES = 48 << 20


def process_sequences(input_stream, output_stream, cfg, hasher: GpuHasher | None = None) -> None:
    """Raises RuntimeError("Failed to create reader: ...") / ("Error reading record: ...") like the reference."""
    hasher = hasher or shared_hasher()
    input_file_name = cfg.input_file_name
    no_file_name = cfg.no_file_name
    if cfg.name_override != "":
        input_file_name = cfg.name_override
    elif input_file_name == "-":
        no_file_name = True   # seqhasher.go:217-219
    out = io.BufferedWriter(output_stream) if isinstance(output_stream, io.RawIOBase) else output_stream
    file_label = input_file_name.encode()

    def flush(batch):
        if not batch:
            return
        hexes, norm = hasher.hash_records([r.seq for r in batch], cfg.hash_types, cfg.case_sensitive,
                                          want_seq=not cfg.headers_only)
        for idx, rec in enumerate(batch):
            hashes = ";".join(hexes[idx]).encode()
            if no_file_name:
                name = hashes + b";" + rec.name if cfg.hash_types else rec.name
            else:
                name = (file_label + b";" + hashes + b";" + rec.name) if cfg.hash_types else (file_label + b";" + rec.name)
            if cfg.headers_only:
                out.write(name + b"\n")
            elif rec.qual is not None and len(rec.qual) > 0:   # record.Format(0): FASTQ iff quality present
                out.write(b"@" + name + b"\n" + norm[idx] + b"\n+\n" + rec.qual + b"\n")
            else:
                out.write(b">" + name + b"\n" + norm[idx] + b"\n")

    try:
        reader = read_records(input_stream)
        batch, nbytes = [], 0
        first = True
        while True:
            try:
                rec = next(reader)
            except StopIteration:
                break
            except FastxError as e:
                if first:
                    raise RuntimeError(f"Failed to create reader: {e}") from None
                raise RuntimeError(f"Error reading record: {e}") from None
            first = False
            batch.append(rec)
            nbytes += len(rec.seq)
            if len(batch) >= BATCH_RECORDS or nbytes >= BATCH_BYTES:
                flush(batch)
                batch, nbytes = [], 0
        flush(batch)
    finally:
        out.flush()
