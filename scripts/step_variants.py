#!/usr/bin/env python
"""Per-kernel microseconds of one every-particle pass (65 536 PSC, L2 flushed) for one or more builds of the library, plus a
parity check of the pass against the default build's result. usage: python scripts/step_variants.py [lib.so ...]
Each library runs in a process of its own (SCGPU_LIB_FAST)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    from sc_b200 import Engine, synth
    from sc_b200.host import HostSystem
    what = sys.argv[2]
    if what == "psc":
        top, cfg, n = synth.psc_bulk()
    else:
        import json
        name = "test_14_normal_SPA_PSC_CPSC" if what == "mix14" else "test_20_chain_bond12"
        inp = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".inputs.json")))
        top, cfg, n = synth.tile(inp["top.init"], inp["config.init"], 12, 12, 12)
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    ev = eng.one_to_all_everyone()
    tot = eng.all_to_all()
    us = np.zeros(4)
    reps = 20
    for _ in range(reps):
        eng.flush_l2()
        us += np.array(eng.profile_everyone())
    us /= reps
    ms = []
    for _ in range(20):
        eng.flush_l2()
        eng.timer_start()
        eng.one_to_all_everyone(fetch=False)
        ms.append(eng.timer_stop())
    print("%-8s gate %.1f cheap %.1f patch %.1f combine %.1f us | pass %.4f ms | sum(ev) %.12e total %.12e" %
          (what, us[0], us[1], us[2], us[3], sum(ms) / len(ms), float(ev.sum()), tot), flush=True)
    sys.exit(0)

libs = sys.argv[1:] or [""]
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["SCGPU_LIB_FAST"] = os.path.abspath(lib)
    print("==", lib or "default build", flush=True)
    for what in ("psc", "mix14"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", what], env=env)
