#!/usr/bin/env python
"""Small driver for compute-sanitizer over the round-2x additions: scgpu_wl_order (all methods, mesh fill + union-find hole search),
scgpu_wl_mesh (labels + export) and scgpu_set_particle_type followed by energies, on Tests/test_mempore (1 500 particles).
    compute-sanitizer --tool memcheck python scripts/sanitize_wl.py"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine                           # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

inp = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "test_mempore.inputs.json.gz")).read().decode())
hs = HostSystem(inp["top.init"], inp["config.init"])
eng = Engine(0, "fast").load(hs)
for wlm in (1, 3, 4, 7, 8, 9):
    w = eng.wl_order(wlm, wlmtype=2, minorder=-3.0, dorder=0.25)
    print("wlm", wlm, w.order[0], w.raw[0])
for ms in (1.0 / 3.0, 0.125):
    w = eng.wl_order((2, 1), wlmtype=2, meshsize=ms)
    lab = eng.wl_mesh(w.mesh_dim[:])
    print("wlm 2 mesh", tuple(w.mesh_dim[:]), "largest hole", w.raw[0], "holes", int(lab.max()), "occupied", w.mesh_occupied)
e0 = eng.one_to_all(5)
eng.set_particle_type(5, 1 if hs.type[5] == 2 else 2)
print("type switch:", e0, "->", eng.one_to_all(5), "total", eng.all_to_all())
eng.close()
