#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped libb200pt.so: TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), shared-memory
atomics / loads / stores, warp votes, code size — and sha256 of every prebuilt binary that travels to the GPU box.
    python scripts/sass_evidence.py [out.txt]"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpu-pathtracer_b200", "csrc", "libb200pt.so")
KEYS = ["UBLKCP", "SYNCS", "ATOMS", "LDS", "STS", "VOTE", "SHFL", "BAR", "LDG", "STG", "FFMA", "FMNMX"]
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_grep.txt")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
rows, name, cnt, n = [], None, None, 0
ins = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if name:
            rows.append((name, n, cnt))
        name, n, cnt = m.group(1), 0, {k: 0 for k in KEYS}
        continue
    m = ins.match(ln)
    if m and name:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k):
                cnt[k] += 1
if name:
    rows.append((name, n, cnt))
demangle = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
with open(out, "w") as f:
    f.write(f"# cuobjdump -sass gpu-pathtracer_b200/csrc/libb200pt.so — instruction counts per kernel (tree at git {head} + working changes)\n")
    f.write("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier try_wait / arrive, ATOMS = shared-memory atomic, VOTE = warp ballot\n")
    f.write(f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{k:>6s}" for k in KEYS) + "\n")
    for (nm, n, c), dn in zip(rows, demangle):
        dn = re.sub(r"\(.*", "", dn).replace("void ", "")
        f.write(f"{dn[:78]:78s} {n:7d} " + " ".join(f"{c[k]:6d}" for k in KEYS) + "\n")
    f.write("\n# sha256 of the prebuilt binaries that travel with gpurun (recipes: gpu-pathtracer_b200/csrc/Makefile, oracle/build_ref.sh, tests/emu/Makefile)\n")
    for p in [LIB, os.path.join(ROOT, "oracle", "libpt_oracle.so"), os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")] + sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "*.so"))):
        if os.path.exists(p):
            f.write(f"{hashlib.sha256(open(p, 'rb').read()).hexdigest()}  {os.path.relpath(p, ROOT)}\n")
print(open(out).read())
