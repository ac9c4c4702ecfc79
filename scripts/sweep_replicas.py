#!/usr/bin/env python
"""Sweep throughput of 1 and R replicas of the 65 536-PSC workload on ONE GPU (asynchronous sweeps on R contexts):
python scripts/sweep_replicas.py [R] [library file]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.engine import MoveParams                # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variant = "fast"
if len(sys.argv) > 2:
    sys.modules["sc_b200.build"].VARIANTS["x"] = (sys.argv[2], [])
    variant = "x"
top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
mp = MoveParams()
mp.temper = 0.1
for k in range(40):
    mp.trans_mx[k] = 2.0 * 0.0212
    mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
mp.n_sub = 1
for reps in (1, R):
    engines = [Engine(0, variant).load(hs) for _ in range(reps)]
    for k in range(2):
        for r, e in enumerate(engines):
            e.sweep(mp, 777 + r, k, stats=False)
    for e in engines:
        e.sync()
    nsw = 10
    t0 = time.perf_counter()
    for k in range(nsw):
        for r, e in enumerate(engines):
            e.sweep(mp, 777 + r, 2 + k, stats=False)
    for e in engines:
        e.sync()
    dt = time.perf_counter() - t0
    print("replicas %d: %.3f ms per sweep round, %.1f sweeps/s aggregate" % (reps, dt / nsw * 1e3, reps * nsw / dt))
    for e in engines:
        e.close()
