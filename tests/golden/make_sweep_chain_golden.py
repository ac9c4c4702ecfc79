#!/usr/bin/env python
"""Statistical golden for the device-side chain moves, produced by the REFERENCE program itself
(oracle/_ref/SC_testing = unmodified sources, sequential sweeps with chainprob > 0): energy time series of the 'chain_fluid'
system (480 bonded SPN-SPA-SPA trimers + 160 PSC rods) at fixed step sizes, several seeds -> sweep_chain_fluid.json.
Runs only in the build container (needs /root/reference)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sc_b200 import synth  # noqa: E402

SC = os.path.join(ROOT, "oracle", "_ref", "SC_testing")
PARAMS = dict(temper=1.0, transmx=0.12, rotmx=15.0, chainprob=0.3, chainmmx=0.15, chainrmx=12.0, nsweeps=2400, report=8)
OPTIONS = """ptype = 1
press = 0
paralpress = 0
shave = 0
nequil = 0
adjust = 0
nsweeps  = %(nsweeps)d
paramfrq = 0
report   = %(report)d
nrepchange = 0
nGrandCanon = 0
nClustMove = 0
movie    = 0
chainprob = %(chainprob)g
transmx = %(transmx)g
rotmx = %(rotmx)g
edge_mx = 0.0
chainmmx = %(chainmmx)g
chainrmx = %(chainrmx)g
temper = %(temper)g
paraltemper = %(temper)g
wlm = 0
wlmtype = 0
switchprob = 0.00
pairlist_update = 8
seed = %(seed)d
write_cluster = 0
"""


def run(seed):
    top, cfg = synth.small_case("chain_fluid")
    tmp = tempfile.mkdtemp(prefix="swchgold_")
    p = dict(PARAMS)
    p["seed"] = seed
    for fn, txt in (("top.init", top), ("config.init", cfg), ("options", OPTIONS % p)):
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    out = subprocess.run([SC], cwd=tmp, capture_output=True, text=True).stdout
    series = []
    with open(os.path.join(tmp, "energy.dat")) as f:
        for line in f:
            t = line.replace(";", " ").split()
            if len(t) >= 2 and not line.lstrip().startswith("#"):
                series.append((int(float(t[0])), float(t[1])))
    acc = [l.strip() for l in out.splitlines() if "acc" in l.lower() or "ratio" in l.lower() or "chain" in l.lower()]
    shutil.rmtree(tmp)
    return series, acc


def main():
    res = {"params": PARAMS, "system": "synth.small_case('chain_fluid')", "runs": []}
    from concurrent.futures import ThreadPoolExecutor
    seeds = (11, 22, 33, 44)
    with ThreadPoolExecutor(4) as ex:
        results = list(ex.map(run, seeds))
    for seed, (series, acc) in zip(seeds, results):
        res["runs"].append({"seed": seed, "sweep": [s for s, _ in series], "energy": [e for _, e in series], "stdout_stats": acc[:16]})
        print("seed", seed, "points", len(series), "E_first", series[0][1], "E_last", series[-1][1])
        print("\n".join(acc[:16]))
    with open(os.path.join(HERE, "sweep_chain_fluid.json"), "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
