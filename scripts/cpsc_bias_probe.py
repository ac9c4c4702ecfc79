"""Probe: window means of <E> for the tiled CPSC system (Tests/System_averages_tests/CPSC/various_temp, 2x2x2 tiles = 800 particles)
under the variants of the checkerboard sweep (trial_rule, n_sub, grid_k), next to the reference's sequential runs of the golden file.
Usage: python scripts/cpsc_bias_probe.py [temper] [W]"""
import gzip, json, math, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine
from sc_b200.engine import MoveParams
from sc_b200.host import HostSystem

PIH = 1.57079632679489661923132169163975
temper = float(sys.argv[1]) if len(sys.argv) > 1 else 0.16
W = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
system = sys.argv[3] if len(sys.argv) > 3 else "cpsc800"
gold = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "sweep_cpsc_temps.json.gz"), "rt").read())
top = gold["top"][system]
cfg = next(r["config"] for r in gold["runs"] if r["system"] == system and r["config"] and abs(r["temper"] - temper) < 1e-9)
seeds = list(range(5, 13))
variants = [("per-cell n_sub50", 0, 50, 0), ("per-particle n_sub50", 1, 50, 0), ("per-particle n_sub1", 1, 1, 0), ("per-cell n_sub1", 0, 1, 0)]


def run(job):
    (name, rule, n_sub, gk), seed = job
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mp = MoveParams()
    mp.temper = temper
    for k in range(40):
        mp.trans_mx[k] = 2.0 * 0.03
        mp.rot_angle[k] = 15.0 / 180.0 * PIH * 0.5
    mp.n_sub, mp.grid_k, mp.trial_rule = n_sub, gk, rule
    en, sw = [], []
    ta = tt = 0
    for k in range(W // n_sub):
        st = eng.sweep(mp, 1000 + seed, k)
        ta += st.trans_acc + st.rot_acc
        tt += st.trans_acc + st.rot_acc + st.trans_rej + st.rot_rej
        if ((k + 1) * n_sub) % 50 == 0:
            sw.append((k + 1) * n_sub)
            en.append(eng.all_to_all())
    eng.close(); hs.close()
    e = np.array([v for s_, v in zip(sw, en) if s_ > W // 3])
    return name, float(e.mean()), ta / max(1, tt), float(en[-1])


ref = [r for r in gold["runs"] if r["system"] == system and abs(r["temper"] - temper) < 1e-9]
rm = [np.mean([v for s_, v in zip(r["sweep"], r["energy"]) if W // 3 < s_ <= W]) for r in ref]
print("reference: mean %.2f  sd %.2f  per seed %s" % (np.mean(rm), np.std(rm, ddof=1), np.round(rm, 1)))
t0 = time.time()
with ThreadPoolExecutor(16) as ex:
    res = list(ex.map(run, [(v, s) for v in variants for s in seeds]))
for v in variants:
    mine = [r for r in res if r[0] == v[0]]
    gm = [r[1] for r in mine]
    print("%-18s mean %.2f  sd %.2f  acc %.3f  per seed %s" % (v[0], np.mean(gm), np.std(gm, ddof=1), np.mean([r[2] for r in mine]), np.round(gm, 1)))
print("wall %.0f s" % (time.time() - t0))
