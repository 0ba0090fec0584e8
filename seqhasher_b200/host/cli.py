"""CLI contract of seqhasher (seqhasher.go:56-149): flags, positionals, version/usage short-circuits,
error texts.  Flag parsing follows Go's `flag` package: -x and --x alike, `-x=v` or `-x v`, parsing stops
at the first positional or after `--`."""
from __future__ import annotations

import sys
from dataclasses import dataclass, field

from .hasher import SUPPORTED_HASH_TYPES, is_valid_hash_type

VERSION = "1.2.0"
DEFAULT_HASH_TYPE = "sha1"


@dataclass
class Config:
    headers_only: bool = False
    hash_types: list = field(default_factory=lambda: [DEFAULT_HASH_TYPE])
    no_file_name: bool = False
    case_sensitive: bool = False
    input_file_name: str = ""
    output_file_name: str = ""
    name_override: str = ""
    show_version: bool = False


class FlagError(Exception):
    pass


class HelpRequested(Exception):
    pass


_BOOL = {"headersonly": "headers_only", "o": "headers_only", "nofilename": "no_file_name", "n": "no_file_name",
         "casesensitive": "case_sensitive", "c": "case_sensitive", "version": "show_version", "v": "show_version"}
_STR = {"hash": "hash", "H": "hash", "name": "name_override", "f": "name_override"}


def parse_flags(argv: list[str]) -> Config:
    """parseFlags (seqhasher.go:101-140).  Raises ValueError with the reference's text for a bad hash type."""
    cfg = Config()
    hash_string = DEFAULT_HASH_TYPE
    i = 0
    while i < len(argv):
        a = argv[i]
        if len(a) < 2 or a[0] != "-":
            break                      # first positional ("-" included) ends flag parsing
        if a == "--":
            i += 1
            break
        body = a[2:] if a.startswith("--") else a[1:]
        if not body or body[0] in "-=":
            raise FlagError(f"bad flag syntax: {a}")
        name, eq, val = body.partition("=")
        if name in ("h", "help"):
            raise HelpRequested()
        if name in _BOOL:
            v = True
            if eq:
                if val.lower() in ("1", "t", "true"):
                    v = True
                elif val.lower() in ("0", "f", "false"):
                    v = False
                else:
                    raise FlagError(f'invalid boolean value "{val}" for -{name}: parse error')
            setattr(cfg, _BOOL[name], v)
        elif name in _STR:
            if not eq:
                i += 1
                if i >= len(argv):
                    raise FlagError(f"flag needs an argument: -{name}")
                val = argv[i]
            if _STR[name] == "hash":
                hash_string = val
            else:
                cfg.name_override = val
        else:
            raise FlagError(f"flag provided but not defined: -{name}")
        i += 1
    rest = argv[i:]
    cfg.input_file_name = rest[0] if len(rest) > 0 else ""
    cfg.output_file_name = rest[1] if len(rest) > 1 else ""
    # validated after TrimSpace but stored untrimmed (seqhasher.go:132-137)
    cfg.hash_types = hash_string.split(",")
    for ht in cfg.hash_types:
        if not is_valid_hash_type(ht.strip(" \t\n\v\f\r")):
            raise ValueError(f"Invalid hash type: {ht}. Supported types are: {', '.join(SUPPORTED_HASH_TYPES)}")
    return cfg


def print_usage(w, prog: str = "seqhasher") -> None:
    """Short usage (the reference's default layout, seqhasher.go:197-207); the coloured -h layout is cosmetics."""
    w.write(f"SeqHasher v{VERSION}\n")
    w.write(f"Usage: {prog} [options] <input_file> [output_file]\n")
    w.write("Options:\n")
    w.write("  -o, --headersonly    Output only headers\n")
    w.write("  -H, --hash string    Hash type(s) (comma-separated: " + ", ".join(SUPPORTED_HASH_TYPES) + ') (default "sha1")\n')
    w.write("  -n, --nofilename     Do not include file name in output\n")
    w.write("  -c, --casesensitive  Case-sensitive hashing\n")
    w.write("  -f, --name string    Override input file name in output\n")
    w.write("  -v, --version        Show version information\n")
    w.write(f"\nSupported hash types: {', '.join(SUPPORTED_HASH_TYPES)}\n")
    w.write("If input_file is '-' or omitted, reads from stdin.\n")
    w.write("If output_file is '-' or omitted, writes to stdout.\n")
    w.write("\nFor more detailed help, use -h or --help.\n")


def run(argv: list[str], w=None, stdin=None) -> None:
    """run (seqhasher.go:62-99).  w is a text writer for version/usage and a binary sink for records
    (sys.stdout by default).  Raises RuntimeError / ValueError with the reference's error texts."""
    from .process import process_sequences_gpu as process_sequences   # whole record loop on the GPU
    w = w or sys.stdout
    cfg = parse_flags(argv)
    if cfg.show_version:
        w.write(f"SeqHasher {VERSION}\n")
        return
    if cfg.input_file_name == "":
        print_usage(w)
        return
    try:
        inp = (stdin or sys.stdin.buffer) if cfg.input_file_name == "-" else open(cfg.input_file_name, "rb")
    except OSError as e:
        raise RuntimeError(f"Error opening input: open {cfg.input_file_name}: {e.strerror.lower()}") from None
    try:
        if cfg.output_file_name not in ("", "-"):
            try:
                outf = open(cfg.output_file_name, "wb")
            except OSError as e:
                raise RuntimeError(f"Error opening output: open {cfg.output_file_name}: {e.strerror.lower()}") from None
            with outf:
                process_sequences(inp, outf, cfg)
        else:
            sink = getattr(w, "buffer", w)
            process_sequences(inp, sink, cfg)
            if hasattr(w, "flush"):
                w.flush()
    finally:
        if inp is not (stdin or sys.stdin.buffer):
            inp.close()


def main(argv: list[str] | None = None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    try:
        run(argv)
    except HelpRequested:
        print_usage(sys.stderr)
        return 0
    except FlagError as e:
        print(e, file=sys.stderr)
        print_usage(sys.stderr)
        return 2
    except (RuntimeError, ValueError) as e:   # log.Fatalf("%v", err): seqhasher.go:57-59
        print(e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    import os
    code = main()
    sys.stdout.flush()
    sys.stderr.flush()
    # skip interpreter / CUDA teardown: unpinning and freeing the staging buffers costs ~1 s and the OS does it anyway
    os._exit(code)
