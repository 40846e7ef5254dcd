#!/usr/bin/env python
"""TEST INFRASTRUCTURE — fixtures that pin gpu-pathtracer_b200/xform.py against the GLM the reference vendors
(oracle/_ref/glm_tool, built by oracle/build_exr_tool.sh from oracle/refbuild/glm_tool.cpp; this container only):

  tests/golden/glm/trs.npz      inputs (scale, translate, rotate in degrees, a vertex, a normal) and what the reference's
                                call sequence gives: trs, transpose(inverse(trs)), the moved vertex, the moved normal
  tests/golden/glm/frames.npz   "rotate" triples / "matrix" 16-tuples of an infinite light and the frame u, v, w

    python oracle/make_glm_fixtures.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "glm_tool")
OUT = os.path.join(ROOT, "tests", "golden", "glm")


def run(tag, rows):
    rows = np.ascontiguousarray(rows, np.float32)
    text = "".join(tag + " " + " ".join(f"{w:08x}" for w in r.view(np.uint32)) + "\n" for r in rows)
    res = subprocess.run([TOOL], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(res) == len(rows)
    return np.array([[int(w, 16) for w in line.split()] for line in res], np.uint32).view(np.float32)


def trs_inputs(n=400):
    rng = np.random.default_rng(20261018)
    rows = np.empty((n, 15), np.float32)
    rows[:, 0:3] = np.exp(rng.uniform(-3, 3, (n, 3)))                       # scale
    rows[:, 3:6] = rng.uniform(-50, 50, (n, 3))                             # translate
    rows[:, 6:9] = rng.uniform(-360, 360, (n, 3))                           # rotate, degrees
    rows[:, 9:12] = rng.uniform(-10, 10, (n, 3))                            # a vertex
    nn = rng.normal(size=(n, 3)); rows[:, 12:15] = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    # the shapes the shipped scenes use: uniform scale only, translate only, one right angle, identity
    rows[0, :9] = (1, 1, 1, 0, 0, 0, 0, 0, 0)
    rows[1, :9] = (25, 25, 25, 0, 0, 0, 0, 0, 0)
    rows[2, :9] = (0.198, 0.198, 0.198, 0, 0.15, 0, 0, 0, 0)
    rows[3, :9] = (1, 1, 1, 0, 15.0, 0, 0, 0, 0)
    rows[4, :9] = (1, 1, 1, 0, 0, 0, 0, 90, 0)
    rows[5, :9] = (1, 1, 1, 0, 0, 0, 180, 0, -90)
    rows[6, :9] = (2, 1, 0.5, 1, 2, 3, 30, 45, 60)
    rows[7, 9:12] = (-0.0, 0.0, -0.0); rows[7, :9] = (1, 1, 1, 0, 0, 0, 0, 0, 0)   # signed zeros through the identity
    return rows


def main():
    os.makedirs(OUT, exist_ok=True)
    rows = trs_inputs()
    out = run("T", rows)
    np.savez_compressed(os.path.join(OUT, "trs.npz"), inputs=rows, trs=out[:, 0:16], inv_t=out[:, 16:32], v=out[:, 32:35], n=out[:, 35:38])
    rng = np.random.default_rng(7)
    rot = rng.uniform(-360, 360, (200, 3)).astype(np.float32)
    rot[0] = (0, 0, 0); rot[1] = (0, 90, 0); rot[2] = (-90, 0, 180)
    fr = run("R", rot)
    mats = rng.normal(size=(200, 16)).astype(np.float32)
    mats[0] = np.eye(4, dtype=np.float32).ravel()
    a = np.radians(30.0)
    mats[1] = np.array([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], np.float32).ravel()
    fm = run("M", mats)
    np.savez_compressed(os.path.join(OUT, "frames.npz"), rotate=rot, rotate_frames=fr, matrix=mats, matrix_frames=fm)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
