#!/usr/bin/env python
"""Small driver for ncu: loads the 65 536-PSC workload and runs a few passes of one entry point.
usage: python scripts/profile_step.py [everyone|all_to_all|cells|sweep] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "everyone"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = "fast"
if len(sys.argv) > 3:                                # experiment builds: python scripts/profile_step.py everyone 3 libscgpu_x46.so
    import sc_b200 as _pkg                           # noqa: F401
    sys.modules["sc_b200.build"].VARIANTS["x"] = (sys.argv[3], [])
    variant = "x"
if what.startswith("membrane"):                       # BASELINE configs[2]: 265 041-particle lipid membrane + CPSC, allToAll
    import gzip
    import json
    inp = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(21, 21, inp["top.init"], inp["config.init"])
else:
    top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
eng = Engine(0, variant).load(hs)
eng.build_cells()
for _ in range(reps):
    if what == "everyone":
        eng.one_to_all_everyone(fetch=False)
    elif what in ("all_to_all", "membrane"):
        eng.all_to_all(fetch=False)
    elif what == "cells":
        eng.set_particles(hs.state, hs.type, hs.moltype)
        eng.build_cells()
    elif what == "sweep":
        from sc_b200.engine import MoveParams
        mp = MoveParams()
        mp.temper = 0.1
        for k in range(40):
            mp.trans_mx[k] = 0.0424
            mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
        mp.n_sub = 1
        eng.sweep(mp, 12345, _)
eng.sync()
print("done", what, reps)
if what in ("everyone", "all_to_all"):                # device time per pass (not under a profiler)
    ms = []
    for k in range(20):
        eng.flush_l2()
        eng.timer_start()
        if what == "everyone":
            eng.one_to_all_everyone(fetch=False)
        else:
            eng.all_to_all(fetch=False)
        ms.append(eng.timer_stop())
    print("ms_per_pass %.4f (min %.4f)" % (sum(ms) / len(ms), min(ms)))
if what == "sweep":                                  # wall-clock per sweep (not under a profiler)
    import time
    t0 = time.perf_counter()
    for k in range(reps, reps + 10):
        eng.sweep(mp, 12345, k)
    eng.sync()
    print("ms_per_sweep %.3f" % ((time.perf_counter() - t0) / 10 * 1e3))
