"""Probe 2: relaxation curve, trials per sweep and energy books of the checkerboard sweeps on the tiled CPSC system at one temperature,
next to the reference's curve (golden). SCGPU_SWEEP_ONE_CELL=1 in the environment removes the decomposition.
Usage: python scripts/cpsc_bias_probe2.py [temper] [W] [n_sub]"""
import gzip, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine
from sc_b200.engine import MoveParams
from sc_b200.host import HostSystem

PIH = 1.57079632679489661923132169163975
temper = float(sys.argv[1]) if len(sys.argv) > 1 else 0.16
W = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
n_sub = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gold = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "sweep_cpsc_temps.json.gz"), "rt").read())
top = gold["top"]["cpsc800"]
cfg = next(r["config"] for r in gold["runs"] if r["system"] == "cpsc800" and r["config"] and abs(r["temper"] - temper) < 1e-9)
marks = [50, 250, 500, 1000, 2000, 3000, 4000, 6000, 8000, 10000, 15000, 20000]


def run(seed):
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mp = MoveParams()
    mp.temper = temper
    for k in range(40):
        mp.trans_mx[k] = 0.06
        mp.rot_angle[k] = 15.0 / 180.0 * PIH * 0.5
    mp.n_sub, mp.grid_k, mp.trial_rule = n_sub, 0, int(os.environ.get("TRIAL_RULE", "0"))
    e = e0 = eng.all_to_all()
    curve = {}
    trials = cellrej = 0
    worst = 0.0
    for k in range(W // n_sub):
        st = eng.sweep(mp, 1000 + seed, k)
        e += st.energy_delta
        trials += st.trans_acc + st.rot_acc + st.trans_rej + st.rot_rej
        cellrej += st.cell_rej
        sw = (k + 1) * n_sub
        if sw in marks or sw % 500 == 0:
            ex = eng.all_to_all()
            worst = max(worst, abs(ex - e))
            e = ex
            if sw in marks:
                curve[sw] = ex
    eng.close(); hs.close()
    return e0, curve, trials / W, cellrej / max(1, trials), worst


ref = [r for r in gold["runs"] if r["system"] == "cpsc800" and abs(r["temper"] - temper) < 1e-9]
print("one cell" if os.environ.get("SCGPU_SWEEP_ONE_CELL") else "checkerboard", "T", temper, "n_sub", n_sub)
print("reference ", {m: round(float(np.mean([r["energy"][r["sweep"].index(m)] for r in ref])), 1) for m in marks if m <= W})
t0 = time.time()
with ThreadPoolExecutor(8) as ex:
    res = list(ex.map(run, range(5, 13)))
print("gpu       ", {m: round(float(np.mean([r[1][m] for r in res])), 1) for m in marks if m <= W})
print("E0 %.4f trials/sweep %.1f cell_rej share %.4f worst book error %.2e  wall %.0f s" % (res[0][0], np.mean([r[2] for r in res]), np.mean([r[3] for r in res]), max(r[4] for r in res), time.time() - t0))
