#!/usr/bin/env python
"""Where the end-to-end step spends its time: host wall clock around each C-ABI call (65 536 PSC, pinned host buffers)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
eng = Engine(0, "fast").load(hs)
pin = torch.empty((n, 9), dtype=torch.float64).pin_memory().numpy()
pin[:] = hs.state[:, :9]
out = torch.empty((n,), dtype=torch.float64).pin_memory().numpy()
eng.set_particles_compact(pin, hs.type, hs.moltype)
T = {"upload": 0.0, "cells": 0.0, "energy+fetch": 0.0, "upload(types)": 0.0, "energy+fetch(pageable)": 0.0}
R = 20
for r in range(R + 2):
    t0 = time.perf_counter(); eng.set_particles_compact(pin, None, None)
    t1 = time.perf_counter(); eng.build_cells(); eng.sync()
    t2 = time.perf_counter(); eng.one_to_all_everyone(fetch=True, out=out)
    t3 = time.perf_counter(); eng.set_particles_compact(pin, hs.type, hs.moltype)
    t4 = time.perf_counter(); eng.one_to_all_everyone(fetch=True)
    t5 = time.perf_counter()
    if r >= 2:
        T["upload"] += t1 - t0; T["cells"] += t2 - t1; T["energy+fetch"] += t3 - t2; T["upload(types)"] += t4 - t3; T["energy+fetch(pageable)"] += t5 - t4
for k, v in T.items():
    print("%-26s %8.1f us" % (k, v / R * 1e6))
x = torch.empty(5242880 // 8, dtype=torch.float64).pin_memory()
d = torch.empty_like(x, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print("pinned H2D 5.2 MB: %.1f us (%.1f GB/s)" % ((time.perf_counter() - t0) / 20 * 1e6, 5242880 * 20 / (time.perf_counter() - t0) / 1e9))
