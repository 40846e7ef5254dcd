#!/usr/bin/env python
"""TEST INFRASTRUCTURE: align the probe output of the reference's probe build and the product's probe build
(scripts/parity_diag.py) and report, per sample, every field of every per-bounce record that differs (in ulps)."""
import re
import struct
import sys


def ulps(a, b):
    def key(x):
        i = struct.unpack("<i", struct.pack("<f", x))[0]
        return i if i >= 0 else -(i & 0x7fffffff)
    try:
        return abs(key(a) - key(b))
    except (OverflowError, struct.error):
        return -1


def parse(lines):
    recs = []
    for ln in lines:
        p = ln.split()
        if not p or p[0] not in "HLMCDVE" or len(p[0]) != 1:
            continue
        vals, names = [], []
        name = ""
        for tok in p[1:]:
            try:
                v = float.fromhex(tok) if ("x" in tok or "nan" in tok or "inf" in tok) else float(int(tok))
                vals.append(v); names.append(name)
            except ValueError:
                name = tok
        recs.append((p[0], names, vals, ln.strip()))
    return recs


def main(path):
    txt = open(path).read().splitlines()
    i = 0
    while i < len(txt):
        if txt[i].startswith("PROBE-REF-BEGIN"):
            title = txt[i]
            j = txt.index("PROBE-REF-END", i)
            ref = parse(txt[i + 1:j])
            k = next(n for n in range(j, len(txt)) if txt[n].startswith("PROBE-OURS-BEGIN"))
            m = txt.index("PROBE-OURS-END", k)
            ours = parse(txt[k + 1:m])
            print("=" * 100); print(title, f"ref records {len(ref)} ours {len(ours)}")
            for tag in "HVLMCDE":
                a = [r for r in ref if r[0] == tag]; b = [r for r in ours if r[0] == tag]
                if len(a) != len(b):
                    print(f"  tag {tag}: ref has {len(a)} records, ours {len(b)}")
                for n, (x, y) in enumerate(zip(a, b)):
                    diffs = []
                    for nm, u, v in zip(x[1], x[2], y[2]):
                        if u != v and not (u != u and v != v):
                            diffs.append(f"{nm}:{ulps(u, v)}ulp({u:.9g} vs {v:.9g})")
                    if diffs:
                        print(f"  {tag}[{n}] " + " ".join(diffs))
            i = m
        i += 1


if __name__ == "__main__":
    main(sys.argv[1])
