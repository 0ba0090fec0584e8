"""Phase breakdown of the parallel gzip reader on the box (SHB_TIMING=1), by thread count."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_c5_cli import make_fasta, gzip_single_member
text, _ = make_fasta(400_000)
open("/tmp/p.fa.gz", "wb").write(gzip_single_member(text, os.cpu_count()))
Z = os.path.join(ROOT, "seqhasher_b200", "bin", "shb-zcat")
for t in (1, 2, 4, 8, 12, 16):
    for rep in range(2):
        p = subprocess.run([Z, "-q", "-t", str(t), "/tmp/p.fa.gz"], capture_output=True, env=dict(os.environ, SHB_TIMING="1"))
    print(t, p.stdout.decode().strip(), p.stderr.decode().strip(), flush=True)
