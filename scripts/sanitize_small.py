#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck): the thread-per-target gate on the rods_wide system (one-to-all of
everyone + allToAll), the forced fallback, sweeps of the round kernel and the cell walk, and a few chain-move sweeps on the chain fluid.
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.engine import MoveParams, ChainMoves    # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "rows"):
    top, cfg = synth.small_case("rods_wide")
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    ev = eng.one_to_all_everyone()
    tot = eng.all_to_all()
    print("rows: sum", float(ev.sum()), "2*total", 2 * tot)
    eng.close()
if what in ("all", "rounds"):          # the round kernel (systems without bonds), all three trial rules, and the cell walk on the same system
    for kind in ("psc_lattice", "mix"):
        top, cfg = synth.small_case(kind)
        hs = HostSystem(top, cfg)
        eng = Engine(0, "fast").load(hs)
        mp = MoveParams()
        mp.temper = 0.5
        mp.n_sub = 1
        for k in range(40):
            mp.trans_mx[k] = 0.1
            mp.rot_angle[k] = 0.1
        e0 = eng.all_to_all()
        tot = 0.0
        for sw, rule in enumerate((0, 1, 2)):
            mp.trial_rule = rule
            tot += eng.sweep(mp, 9, sw).energy_delta
        os.environ["SCGPU_SWEEP_KERNEL"] = "cells"
        tot += eng.sweep(mp, 9, 3).energy_delta
        del os.environ["SCGPU_SWEEP_KERNEL"]
        print("rounds", kind, ": drift", eng.all_to_all() - e0 - tot)
        eng.close()
if what in ("all", "chains"):
    top, cfg = synth.small_case("chain_fluid")
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mp = MoveParams()
    mp.temper = 1.0
    mp.n_sub = 1
    cm = ChainMoves()
    cm.chainprob = 0.5
    for k in range(40):
        mp.trans_mx[k] = 0.2
        mp.rot_angle[k] = 0.1
    for k in range(32):
        cm.chainm_mx[k] = 0.3
        cm.chainr_angle[k] = 0.2
    e0 = eng.all_to_all()
    tot = 0.0
    for sw in range(2):
        st, cst = eng.sweep_chains(mp, cm, 5, sw)
        tot += st.energy_delta + cst.energy_delta
    print("chains: drift", eng.all_to_all() - e0 - tot)
    st = eng.pressure_move(1, 0.1, 1.0, 0.1, 3, 3)
    print("pressure move accepted", st.accepted)
    eng.close()
