#!/usr/bin/env python
"""tests/golden/spoon_mesh.npz from the reference's test/data/spoon.obj (2504 vertices, 2502 quadrilateral faces): the surface mesh
of configuration C2's spoon scene (test/spoon.jl:36-41).  /root/reference is not available on the GPU boxes, so the vertex / face
arrays are committed as a compressed fixture (data, not source); run this script where the reference is mounted to regenerate it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/data/spoon.obj"

v, f = [], []
for line in open(SRC):
    t = line.split()
    if not t:
        continue
    if t[0] == "v":
        v.append([float(t[1]), float(t[2]), float(t[3])])
    elif t[0] == "f":
        f.append([int(tok.split("/")[0]) - 1 for tok in t[1:]])
v = np.array(v, np.float64)
assert all(len(q) == 4 for q in f), "spoon.obj has quadrilateral faces only"
quad = np.array(f, np.int32)
out = os.path.join(ROOT, "tests", "golden", "spoon_mesh.npz")
np.savez_compressed(out, vertices=v, quads=quad, source="ryanelandt/PressureFieldContact.jl test/data/spoon.obj")
print("wrote", out, os.path.getsize(out), "bytes:", v.shape, quad.shape)
