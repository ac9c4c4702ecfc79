#!/usr/bin/env python
"""Where a sweep of the 265 041-particle lipid membrane (configs[2]) spends its time: single-bead passes on the grids K = 1..3, chain passes."""
import gzip, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth
from sc_b200.engine import MoveParams, ChainMoves
from sc_b200.host import HostSystem

inp = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "membrane601.inputs.json.gz")).read().decode())
top, cfg, n = synth.membrane(21, 21, inp["top.init"], inp["config.init"])
hs = HostSystem(top, cfg)
eng = Engine(0, "fast").load(hs)
mp = MoveParams()
mp.temper = 1.0
mp.n_sub = 1
for k in range(40):
    mp.trans_mx[k] = 0.1
    mp.rot_angle[k] = 10.0 / 180.0 * 1.5707963267948966 * 0.5
for gk in (1, 2, 3):
    mp.grid_k = gk
    eng.sweep(mp, 1, 0)
    t0 = time.perf_counter()
    for k in range(2):
        st = eng.sweep(mp, 1, 1 + k)
    print("single-bead sweep, grid K=%d: %.1f ms  (acc %d rej %d cell_rej %d)" % (gk, (time.perf_counter() - t0) / 2 * 1e3, st.trans_acc + st.rot_acc, st.trans_rej + st.rot_rej, st.cell_rej), flush=True)
cm = ChainMoves()
cm.chainprob = 1.0
for k in range(32):
    cm.chainm_mx[k] = 0.2
    cm.chainr_angle[k] = 10.0 / 180.0 * 1.5707963267948966
mp.grid_k = 0
eng.sweep_chains(mp, cm, 2, 0)
t0 = time.perf_counter()
for k in range(2):
    st, cst = eng.sweep_chains(mp, cm, 2, 1 + k)
print("chain-only sweep (chainprob 1): %.1f ms (acc %d rej %d cell_rej %d)" % ((time.perf_counter() - t0) / 2 * 1e3, cst.chainm_acc + cst.chainr_acc, cst.chainm_rej + cst.chainr_rej, cst.cell_rej))
