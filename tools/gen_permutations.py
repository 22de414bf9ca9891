"""Extract the reference's 100x255 MinHash permutation table as DATA.

The table in the reference (permutations.c:7-1808) was produced by a rand()-based
generator whose sequence is libc-specific (generatepermutations.c:31,45,75) and cannot be
regenerated, so it is a golden data constant.  This script reads it out of the compiled
reference library (oracle/_ref, via get_permutation()) and writes it as a flat hex
include used by the CUDA product and (a separate copy) by the C oracle.

    python tools/gen_permutations.py
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref import Ref  # noqa: E402

perm = Ref().permutations()
assert perm.shape == (100, 255) and perm.max() < 8192
digest = hashlib.sha256(perm.astype("<u2").tobytes()).hexdigest()
lines = [f"/* 100 x 255 u16 MinHash bit positions; data extracted from the reference's",
         f" * get_permutation() (permutations.c:7-1813) by tools/gen_permutations.py.",
         f" * sha256(le u16 bytes) = {digest} */"]
flat = perm.reshape(-1)
for i in range(0, len(flat), 15):
    lines.append(",".join(f"0x{v:04x}" for v in flat[i:i + 15]) + ",")
text = "\n".join(lines) + "\n"
for rel in ("mnemophonix_b200/data/permutations_table.inc", "oracle/permutations_table.inc"):
    with open(os.path.join(ROOT, rel), "w") as f:
        f.write(text)
print("wrote", len(flat), "values, sha256", digest)
