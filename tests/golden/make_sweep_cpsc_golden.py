#!/usr/bin/env python
"""Statistical goldens from the reference's own 'comprehensive' inputs: Tests/System_averages_tests/CPSC/various_temp/T_0.1, T_0.16,
T_0.22 (100 CPSC in a 10^3 box), run by the UNMODIFIED reference program
(oracle/_ref/SC_testing, sequential sweeps, Ran2) for 2*10^5 sweeps and 8 seeds each, plus the same configuration tiled 2 x 2 x 2
(800 particles, 20^3 box: several cells per axis, so the checkerboard decomposition is exercised) for 2*10^4 sweeps. Stored per run:
the energy time series and the acceptance percentages the reference prints (Statistics::print).
The shipped top.init is in the pre-PARALLEL_EPS format the current parser rejects; the one type in use (F, CPSC) is rewritten in
the current format with PARALLEL_EPS = 0 -- same physics. Runs only in the build container (needs /root/reference).
Output: sweep_cpsc_temps.json.gz"""
import gzip
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/Tests/System_averages_tests/CPSC/various_temp"
SC = os.path.join(ROOT, "oracle", "_ref", "SC_testing")
TEMPS = ["0.1", "0.16", "0.22"]
SEEDS = [101, 202, 303, 404, 505, 606, 707, 808]
TOP = """[Types]
#NAME NUMBER GEOTYPE EPSILON SIGMA ATTRACT_DIST ATTRACT_SWITCH PATCH_ANGLE PATCH_SWITCH SC_LENGTH PARALLEL_EPS
F     6      CPSC     1.0     1.0   1.122        0.878          30          5.0          2.0     0.0
[Molecules]
F: {
particles:   6
}
[System]
F %d
"""


def tiled_config(text, k):
    lines = [l for l in text.splitlines() if l.strip()]
    box = [float(x) for x in lines[0].split()]
    out = [" ".join("%.8e" % (b * k) for b in box)]
    for iz in range(k):
        for iy in range(k):
            for ix in range(k):
                for l in lines[1:]:
                    t = l.split()
                    p = [float(t[0]) + ix * box[0], float(t[1]) + iy * box[1], float(t[2]) + iz * box[2]]
                    out.append(" ".join("%.8e" % x for x in p) + "   " + " ".join(t[3:]))
    return "\n".join(out) + "\n"


def run(job):
    system, temp, seed = job
    src = os.path.join(REF, "T_" + temp)
    opt = open(os.path.join(src, "options")).read()
    cfg = open(os.path.join(src, "config.init")).read()
    k = 1 if system == "cpsc100" else 2
    nsweeps, report = (200000, 500) if k == 1 else (20000, 50)
    opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = %d" % nsweeps, opt)
    opt = re.sub(r"(?m)^report\s*=\s*\d+", "report = %d" % report, opt)
    opt = re.sub(r"(?m)^write_cluster\s*=\s*\d+", "write_cluster = 0", opt)
    opt = re.sub(r"(?m)^seed\s*=\s*\d+", "seed = %d" % seed, opt)
    tmp = tempfile.mkdtemp(prefix="cpscgold_")
    cfg_used = tiled_config(cfg, k)
    for fn, txt in (("top.init", TOP % (100 * k ** 3)), ("config.init", cfg_used), ("options", opt)):
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    out = subprocess.run([SC], cwd=tmp, capture_output=True, text=True).stdout
    sweeps, energy = [], []
    with open(os.path.join(tmp, "energy.dat")) as f:
        for line in f:
            t = line.replace(";", " ").split()
            if len(t) >= 2 and not line.lstrip().startswith("#"):
                sweeps.append(int(float(t[0])))
                energy.append(float(t[1]))
    acc = {}
    m = re.search(r"Single particle translation:\s+(\d+)\s+(\d+)\s+(\d+)", out)
    if m:
        acc["trans_acc_pct"], acc["trans_steps"] = int(m.group(1)), int(m.group(3))
    m = re.search(r"Single particle rotation:\s+(\d+)\s+(\d+)\s+(\d+)", out)
    if m:
        acc["rot_acc_pct"], acc["rot_steps"] = int(m.group(1)), int(m.group(3))
    shutil.rmtree(tmp)
    print(system, temp, seed, "points", len(energy), "E_last", energy[-1], acc, flush=True)
    return {"system": system, "temper": float(temp), "seed": seed, "sweep": sweeps, "energy": energy, "acceptance": acc,
            "options": opt if seed == SEEDS[0] else None, "config": cfg_used if seed == SEEDS[0] else None}      # every T_x directory ships its own equilibrated config.init


def main():
    jobs = [(s, t, seed) for s in ("cpsc100", "cpsc800") for t in TEMPS for seed in SEEDS]
    with ThreadPoolExecutor(8) as ex:
        runs = list(ex.map(run, jobs))
    res = {"top": {"cpsc100": TOP % 100, "cpsc800": TOP % 800}, "runs": runs,
           "source": "Tests/System_averages_tests/CPSC/various_temp/T_{0.1,0.16,0.22} run by oracle/_ref/SC_testing (unmodified reference, -DTESTING)"}
    with gzip.open(os.path.join(HERE, "sweep_cpsc_temps.json.gz"), "wt") as f:
        json.dump(res, f)


def add_configs():
    """patch an existing golden file: the starting configuration of every (system, temperature), without re-running the reference"""
    path = os.path.join(HERE, "sweep_cpsc_temps.json.gz")
    res = json.loads(gzip.open(path, "rt").read())
    for r in res["runs"]:
        if r["seed"] == SEEDS[0]:
            temp = next(t for t in TEMPS if abs(float(t) - r["temper"]) < 1e-9)
            r["config"] = tiled_config(open(os.path.join(REF, "T_" + temp, "config.init")).read(), 1 if r["system"] == "cpsc100" else 2)
    with gzip.open(path, "wt") as f:
        json.dump(res, f)


if __name__ == "__main__":
    add_configs() if "--add-configs" in sys.argv else main()
