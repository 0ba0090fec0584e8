"""Aggregate an `ncu --page source --csv` dump into executed warp-instructions per SASS opcode.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_opmix.py src.csv <units> [label]
<units> = number of warp-level work units (e.g. SHA-1 blocks / 32) to normalise by."""
import collections
import csv
import sys

ALU = {"LOP3", "SHF", "IADD3", "PRMT", "LEA", "ISETP", "SEL", "VIMNMX", "IABS", "FLO", "POPC", "BREV", "SGXT", "BMSK",
       "LOP", "IMNMX", "VIADDMNMX", "PLOP3", "P2R", "R2P", "FMNMX", "FSEL", "FSETP"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "VIADD", "MOV", "HFMA2", "IDP", "IDP4A"}


def main():
    path, units = sys.argv[1], float(sys.argv[2])
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(lines[start:]))
    tot = collections.Counter()
    for r in rows:
        toks = r["Source"].strip().split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        parts = op.rstrip(";").split(".")
        k = parts[0]
        if len(parts) > 1 and parts[1] in ("WIDE", "HI", "IADD", "MOV", "SHL"):
            k += "." + parts[1]
        tot[k] += int(r["Instructions Executed"])
    s = sum(tot.values())
    alu = sum(v for k, v in tot.items() if k.split(".")[0] in ALU)
    fma = sum(v for k, v in tot.items() if k.split(".")[0] in FMA)
    print(f"total {s}  per unit {s / units:.1f}   ALU-pipe ~{alu / units:.1f}  FMA-pipe ~{fma / units:.1f}")
    for k, v in tot.most_common(40):
        print(f"  {k:14s} {v / units:9.2f}")


if __name__ == "__main__":
    main()
