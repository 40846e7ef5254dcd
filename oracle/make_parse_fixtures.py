#!/usr/bin/env python
"""TEST INFRASTRUCTURE — tests/golden/parse/<scene>.json.bin.gz: what the reference's own scene parser (oracle/_ref/parse_tool,
built by oracle/build_parse_tool.sh; this container only) produces for the scenes tests/parse_cases.py stages.

    python oracle/make_parse_fixtures.py"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import parse_cases as pc  # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "parse_tool")
OUT = os.path.join(ROOT, "tests", "golden", "parse")


def main():
    os.makedirs(OUT, exist_ok=True)
    dst = os.path.join(tempfile.mkdtemp(prefix="b200pt_parse_"), "cornell_box")
    for name in pc.stage(dst):
        raw = os.path.join(dst, name + ".bin")
        subprocess.run([TOOL, os.path.join(dst, name), raw], check=True, stdout=subprocess.DEVNULL)
        with open(raw, "rb") as f, gzip.GzipFile(os.path.join(OUT, name + ".bin.gz"), "wb", mtime=0) as g:
            g.write(f.read())
        print(name, os.path.getsize(raw), "->", os.path.getsize(os.path.join(OUT, name + ".bin.gz")), "bytes")


if __name__ == "__main__":
    main()
