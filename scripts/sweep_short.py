#!/usr/bin/env python
"""A handful of checkerboard sweeps of the 65 536-rod benchmark system (the short command ncu wraps)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth                    # noqa: E402
from sc_b200.engine import MoveParams                # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402
top, cfg, n = synth.psc_bulk()
hs = HostSystem(top, cfg)
eng = Engine(0, "fast").load(hs)
mp = MoveParams()
mp.temper = 0.1
for k in range(40):
    mp.trans_mx[k] = 2.0 * 0.0212
    mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
mp.n_sub = 1
for sw in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    st = eng.sweep(mp, 12345, sw)
print("acc", st.trans_acc + st.rot_acc, "rej", st.trans_rej + st.rot_rej)
