#!/usr/bin/env python
"""Debug: per-phase cycle counts of k_sweep_cells (needs sc_b200/libscgpu_prof.so built with -DSW_PROFILE; extra -D flags, e.g.
-DSW_FORCE_K=1, select the grid):
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=true -DSCG_FAST_DIV -DSW_PROFILE -shared -Xcompiler -fPIC \
       -o sc_b200/libscgpu_prof.so sc_b200/csrc/scgpu.cu -ldl"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sc_b200                                       # noqa: E402
B = sys.modules["sc_b200.build"]
B.VARIANTS["prof"] = ("libscgpu_prof.so", ["-fmad=true", "-DSW_PROFILE"])
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.engine import MoveParams, load_library  # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
eng = Engine(0, "prof").load(hs)
L = load_library("prof")
mp = MoveParams()
mp.temper = 0.1
for k in range(40):
    mp.trans_mx[k] = 2.0 * 0.0212
    mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
mp.n_sub = 1
for sw in range(3):
    eng.sweep(mp, 12345, sw)
buf = (C.c_ulonglong * 16)()
L.scgpu_sweep_profile(buf, 1)
t0 = time.perf_counter()
NS = 10
for sw in range(3, 3 + NS):
    eng.sweep(mp, 12345, sw)
eng.sync()
t1 = time.perf_counter()
L.scgpu_sweep_profile(buf, 0)
v = list(buf)
ntr = v[15]
names = ["0 rng+pick", "1 load rec+apply", "2 -", "3 gate+cheap", "4 -", "5 patch", "6 reduce", "7 metropolis+commit"]
tot = sum(v[:8])
print("trials", ntr, "ms/sweep", (t1 - t0) / NS * 1e3, "non-empty active cells (blocks)", v[14], "prologue cycles/block %.0f" % (v[8] / max(1, v[14])),
      "trials/block %.2f" % (ntr / max(1, v[14])))
for k in range(8):
    print("%-22s %9.0f cycles/trial  %5.1f %%" % (names[k], v[k] / ntr, 100.0 * v[k] / tot))
print("total cycles/trial %.0f" % (tot / ntr))
