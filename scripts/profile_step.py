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
elif what.startswith("mix14") or what.startswith("chains20"):          # BASELINE configs[3] tiled 12^3
    import json
    name = "test_14_normal_SPA_PSC_CPSC" if what.startswith("mix14") else "test_20_chain_bond12"
    inp = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".inputs.json")))
    top, cfg, n = synth.tile(inp["top.init"], inp["config.init"], 12, 12, 12)
    what = "everyone"
else:
    top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
eng = Engine(0, variant).load(hs)
eng.build_cells()
if what in ("everyone", "all_to_all", "membrane"):     # one synchronous call first: list sizes and the gate's kernel choice settle here
    eng.all_to_all() if what != "everyone" else eng.one_to_all_everyone()
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
        mp.grid_k = int(os.environ.get("SWEEP_GRID_K", "0"))
        mp.trial_rule = int(os.environ.get("SWEEP_TRIAL_RULE", "0"))
        eng.sweep(mp, 12345, _)
    elif what == "membrane_chainsweep":               # configs[2] with the reference's move mix for lipids: chain moves of whole 3-bead molecules
        from sc_b200.engine import MoveParams, ChainMoves
        mp = MoveParams()
        mp.temper = 1.0
        cm = ChainMoves()
        cm.chainprob = 0.5
        for k in range(40):
            mp.trans_mx[k] = 0.1
            mp.rot_angle[k] = 10.0 / 180.0 * 1.5707963267948966 * 0.5
        for k in range(32):
            cm.chainm_mx[k] = 0.2
            cm.chainr_angle[k] = 10.0 / 180.0 * 1.5707963267948966
        mp.n_sub = 1
        st, cst = eng.sweep_chains(mp, cm, 4242, _)
        print("sweep", _, "particle acc/rej", st.trans_acc + st.rot_acc, st.trans_rej + st.rot_rej, "chain acc/rej", cst.chainm_acc + cst.chainr_acc,
              cst.chainm_rej + cst.chainr_rej, "chain cell_rej", cst.cell_rej)
eng.sync()
print("done", what, reps)
if what == "membrane_chainsweep":
    import time
    t0 = time.perf_counter()
    for k in range(reps, reps + 3):
        eng.sweep_chains(mp, cm, 4242, k)
    eng.sync()
    print("ms_per_sweep %.3f (N = %d, chainprob 0.5)" % ((time.perf_counter() - t0) / 3 * 1e3, n))
if what in ("everyone", "all_to_all", "membrane"):    # device time per pass (not under a profiler)
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
