#!/usr/bin/env python
"""Statistical golden for batched NPT (checkerboard sweeps + scgpu_pressure_move), produced by the REFERENCE program itself
(oracle/_ref/SC_testing = unmodified sources, sequential sweeps with shave = 1 volume move per sweep, ptype 1): volume and energy
time series of the 1280-rod PSC fluid ('psc_lattice' start) -> sweep_npt_psc1280.json. The volume comes from the box line of
the movie frames (Updater::dumpMovie, mc/updater.h:82-93). Runs only in the build container (needs /root/reference)."""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sc_b200 import synth  # noqa: E402

SC = os.path.join(ROOT, "oracle", "_ref", "SC_testing")
PARAMS = dict(temper=1.0, press=0.1, ptype=1, shave=1, edge_mx=0.08, transmx=0.15, rotmx=20.0, nsweeps=3200, report=8)
OPTIONS = """ptype = %(ptype)d
press = %(press)g
paralpress = %(press)g
shave = %(shave)d
nequil = 0
adjust = 0
nsweeps  = %(nsweeps)d
paramfrq = 0
report   = %(report)d
nrepchange = 0
nGrandCanon = 0
nClustMove = 0
movie    = %(report)d
chainprob = 0.0
transmx = %(transmx)g
rotmx = %(rotmx)g
edge_mx = %(edge_mx)g
chainmmx = 0.0
chainrmx = 0.0
temper = %(temper)g
paraltemper = %(temper)g
wlm = 0
wlmtype = 0
switchprob = 0.00
pairlist_update = 8
seed = %(seed)d
write_cluster = 0
"""


def run(seed, nsweeps=None):
    top, cfg = synth.small_case("psc_lattice")
    tmp = tempfile.mkdtemp(prefix="swnptgold_")
    p = dict(PARAMS)
    p["seed"] = seed
    if nsweeps:
        p["nsweeps"] = nsweeps
    for fn, txt in (("top.init", top), ("config.init", cfg), ("options", OPTIONS % p)):
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    out = subprocess.run([SC], cwd=tmp, capture_output=True, text=True).stdout
    sweeps, vols = [], []
    for fn in os.listdir(tmp):
        if fn.startswith("movie"):
            for line in open(os.path.join(tmp, fn)):
                m = re.match(r"sweep (\d+); box (\S+) (\S+) (\S+)", line)
                if m:
                    sweeps.append(int(m.group(1)))
                    vols.append(float(m.group(2)) * float(m.group(3)) * float(m.group(4)))
    energy = []
    with open(os.path.join(tmp, "energy.dat")) as f:
        for line in f:
            t = line.replace(";", " ").split()
            if len(t) >= 2 and not line.lstrip().startswith("#"):
                energy.append((int(float(t[0])), float(t[1])))
    acc = [l.strip() for l in out.splitlines() if "acc" in l.lower() or "pressure" in l.lower() or "volume" in l.lower()]
    shutil.rmtree(tmp)
    return sweeps, vols, energy, acc


def main():
    if "probe" in sys.argv[1:]:
        sw, v, e, acc = run(11, 600)
        print(len(v), v[:3], v[-5:], "\n".join(acc[:10]))
        return
    res = {"params": PARAMS, "system": "synth.small_case('psc_lattice')", "runs": []}
    from concurrent.futures import ThreadPoolExecutor
    seeds = (11, 22, 33, 44, 55, 66, 77, 88, 99, 110, 121, 132)
    with ThreadPoolExecutor(4) as ex:
        results = list(ex.map(run, seeds))
    for seed, (sw, v, e, acc) in zip(seeds, results):
        res["runs"].append({"seed": seed, "sweep": sw, "volume": v, "energy_sweep": [a for a, _ in e], "energy": [b for _, b in e], "stdout_stats": acc[:12]})
        print("seed", seed, "frames", len(v), "V_first", v[0], "V_last", v[-1])
    with open(os.path.join(HERE, "sweep_npt_psc1280.json"), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
