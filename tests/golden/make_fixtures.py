"""Copies the reference's DATA fixtures (test/*.fasta, *.fastq and their compressed twins — inputs, not code)
into tests/golden/fixtures/ so the GPU box, where /root/reference does not exist, can run the same inputs, and
records their sha256 in MANIFEST.json.  Run here (in the build container): python tests/golden/make_fixtures.py"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/test"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")
os.makedirs(DST, exist_ok=True)
manifest = {}
for name in sorted(os.listdir(SRC)):
    if name.endswith(".sh"):
        continue   # the reference's bash test script is code, not a fixture; its cases are restated in tests/
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    manifest[name] = hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest()
json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
print(f"{len(manifest)} fixtures -> {DST}")
