#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`); the fixtures it writes are committed so that tests never read
/root/reference at run time.

For every Tests/test_*/new and Tests/volumeChange/*/new directory of the reference:
  <name>_init.ref.gz   reference-driver dump of the shipped initial configuration
  <name>.config.last   trajectory end state of the unmodified reference program (-DTESTING, Ran2)
  <name>_end.ref.gz    reference-driver dump of that end state fed back in as config.init
  <name>.inputs.json   options / top.init / config.init text (so tests can re-parse them)
Pose grids modelled on Tests/Interactions_tests (test.sh, main.cpp:137-197):
  grid_<A>_<B>.npz     top text + pose-set kind + E(0,j), E(j,0) for every pose j
  grid_config_<kind>.txt.gz   the config.init text of each pose set (shared)
"""
import gzip
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile

import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "sc_ref_driver")
SC = os.path.join(ROOT, "oracle", "_ref", "SC_testing")
sys.path.insert(0, ROOT)

BIG = {"test_mempore", "test_pscthrough", "test_wallfibril"}  # init dump only (size)


def run(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def do_case(name, src):
    tmp = tempfile.mkdtemp(prefix="gold_")
    for fn in os.listdir(src):
        shutil.copy(os.path.join(src, fn), tmp)
    inputs = {}
    for fn in ("options", "top.init", "config.init"):
        with open(os.path.join(tmp, fn)) as f:
            inputs[fn] = f.read()
    if os.path.exists(os.path.join(tmp, "wl.dat")):
        with open(os.path.join(tmp, "wl.dat")) as f:
            inputs["wl.dat"] = f.read()
    big = name in BIG
    if not big:
        with open(os.path.join(HERE, name + ".inputs.json"), "w") as f:
            json.dump(inputs, f)
    else:
        with gzip.GzipFile(os.path.join(HERE, name + ".inputs.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(inputs).encode())
    run([DRIVER, "dump", "ref_dump.txt"], tmp)
    gz_copy(os.path.join(tmp, "ref_dump.txt"), os.path.join(HERE, name + "_init.ref.gz"))
    if not big:
        run([SC], tmp)
        shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, name + ".config.last"))
        shutil.copy(os.path.join(tmp, "config.last"), os.path.join(tmp, "config.init"))
        run([DRIVER, "dump", "ref_dump_end.txt"], tmp)
        gz_copy(os.path.join(tmp, "ref_dump_end.txt"), os.path.join(HERE, name + "_end.ref.gz"))
    shutil.rmtree(tmp)
    print("golden:", name)


# ---- pose grids (our own generator, same construction as Tests/Interactions_tests/main.cpp:137-197)
def fibonacci_sphere(samples):
    offset = 2.0 / samples
    inc = math.pi * (3.0 - math.sqrt(5.0))
    out = []
    for i in range(samples):
        z = ((i * offset) - 1) + (offset / 2)
        r = math.sqrt(1 - z * z)
        phi = i * inc
        out.append((math.cos(phi) * r, math.sin(phi) * r, z))
    return out


def rotate(v, axis, angle):
    c, s = math.cos(angle), math.sin(angle)
    qw, qx, qy, qz = c, axis[0] * s, axis[1] * s, axis[2] * s
    t2, t3, t4 = qw * qx, qw * qy, qw * qz
    t5, t6, t7 = -qx * qx, qx * qy, qx * qz
    t8, t9, t10 = -qy * qy, qy * qz, -qz * qz
    x, y, z = v
    return (2.0 * ((t8 + t10) * x + (t6 - t4) * y + (t3 + t7) * z) + x,
            2.0 * ((t4 + t6) * x + (t5 + t10) * y + (t9 - t2) * z) + y,
            2.0 * ((t7 - t3) * x + (t2 + t9) * y + (t5 + t8) * z) + z)


def grid_config(kind, samples=10, angle_inc=1.0):
    """config.init text: particle 0 = A at the box centre (dir x, patch y), then B poses."""
    lines = ["10 10 10", "5 5 5   1 0 0  0 1 0     0 0"]
    if kind == "sc":
        coords = [1, 4, 7]
        beads = fibonacci_sphere(samples)
        for x in coords:
            for y in coords:
                for z in coords:
                    for d in beads:
                        j = 0.0
                        while j < 6.28318530718:
                            par = (d[0], d[1], -(d[0] + d[1]) / d[2])
                            par = rotate(par, d, j)
                            lines.append("%d %d %d %g %g %g %g %g %g 0 0" % (x, y, z, d[0], d[1], d[2], par[0], par[1], par[2]))
                            j += angle_inc
    elif kind == "sphere":
        v = 1.0
        while v <= 9.0 + 1e-9:
            lines.append("%.4f 5 5  1 0 0  0 1 0  0 0" % v)
            lines.append("5 %.4f 5  1 0 0  0 1 0  0 0" % v)
            v += 0.0333
    else:  # sphere partner around a rod
        v = 1.0
        while v <= 9.0 + 1e-9:
            w = 1.0
            while w <= 9.0 + 1e-9:
                z = 1.0
                while z <= 5.0 + 1e-9:
                    lines.append("%.2f %.2f %.2f  1 0 0  0 1 0  0 0" % (v, w, z))
                    z += 0.77
                w += 0.77
            v += 0.77
    return "\n".join(lines) + "\n"


SPHERE = {"SPN": "SPN 1.33 0.95", "SPA": "SPA 1.33 1.2 1.346954458 1.6"}
SC_T = {
    "PSC": "PSC 1.33 1.2 1.346954458 0.3 30.0 0.0 3 0.0",
    "CPSC": "CPSC 1.33 1.2 1.346954458 0.3 30.0 0.0 3 0.0",
    "CHPSC": "CHPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 5.0",
    "CHCPSC": "CHCPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 5.0",
    "TPSC": "TPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 180.0 90 5.0",
    "TCPSC": "TCPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 180.0 90 5.0",
    "TCHPSC": "TCHPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 180.0 90 5.0 5.0",
    "TCHCPSC": "TCHCPSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 180.0 90 5.0 5.0",
    # extra coverage beyond the reference's grid: parallel-eps term and the rarely used SCN / SCA rods
    "PSCpar": "PSC 1.333333 1.2 1.346954458 0.3 90 5.0 3 0.7",
    "SCN": "SCN 1.33 1.2 3",
    "SCA": "SCA 1.33 1.2 1.346954458 0.3 3",
}


def top_text(defA, defB, nB):
    return ("[Types]\nA 1 %s\nB 2 %s\n[Molecules]\nA: {\nparticles: 1\n}\nB: {\nparticles: 2\n}\n[System]\nA 1\nB %d\n"
            % (defA, defB, nB))


def do_grid(nameA, defA, nameB, defB, kind, options_text):
    cfg = grid_config(kind)
    nB = len(cfg.strip().split("\n")) - 2
    top = top_text(defA, defB, nB)
    tmp = tempfile.mkdtemp(prefix="grid_")
    for fn, txt in (("top.init", top), ("config.init", cfg), ("options", options_text)):
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    run([DRIVER, "dump0", "ref_dump.txt"], tmp)
    e0j = np.zeros(nB + 1)
    ej0 = np.zeros(nB + 1)
    ov0j = np.zeros(nB + 1, dtype=np.int8)
    ovj0 = np.zeros(nB + 1, dtype=np.int8)
    with open(os.path.join(tmp, "ref_dump.txt")) as f:
        for line in f:
            t = line.split()
            if t and t[0] == "OV0":
                ov0j[int(t[1])] = int(t[2])
                ovj0[int(t[1])] = int(t[3])
            if t and t[0] == "E":
                i, j, e = int(t[1]), int(t[2]), float.fromhex(t[3])
                if i == 0:
                    e0j[j] = e
                else:
                    ej0[i] = e
    cfg_file = os.path.join(HERE, "grid_config_%s.txt.gz" % kind)   # shared by every grid of this kind
    if not os.path.exists(cfg_file):
        with gzip.GzipFile(cfg_file, "wb", mtime=0) as g:
            g.write(cfg.encode())
    np.savez_compressed(os.path.join(HERE, "grid_%s_%s.npz" % (nameA, nameB)), top=top, kind=kind, e0j=e0j, ej0=ej0, ov0j=ov0j, ovj0=ovj0)
    shutil.rmtree(tmp)
    print("grid:", nameA, nameB, nB, "poses; nonzero", int(np.count_nonzero(e0j)), "overlaps", int(ov0j.sum()))


def do_extras():
    """multi-cell synthetic systems (sc_b200/synth.py) through the reference: pins angle1/angle2/bondh/bondd,
    exclusions, negative parallel-eps and two-patch chiral mixes, which the reference's own test inputs never use"""
    from sc_b200 import synth
    with open(os.path.join(REF, "Tests", "test_01_normal_PSC", "new", "options")) as f:
        opts = f.read()
    for kind in ("chains", "mix"):
        top, cfg = synth.small_case(kind)
        tmp = tempfile.mkdtemp(prefix="extra_")
        for fn, txt in (("top.init", top), ("config.init", cfg), ("options", opts)):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        run([DRIVER, "dump", "ref_dump.txt"], tmp)
        gz_copy(os.path.join(tmp, "ref_dump.txt"), os.path.join(HERE, "extra_%s_init.ref.gz" % kind))
        with gzip.GzipFile(os.path.join(HERE, "extra_%s.inputs.json.gz" % kind), "wb", mtime=0) as g:
            g.write(json.dumps({"options": opts, "top.init": top, "config.init": cfg}).encode())
        shutil.rmtree(tmp)
        print("extra:", kind)


def do_short():
    """short trajectories of the unmodified reference program for the sequential GPU-path test (that path pays ~0.1 ms of
    launch + FP64 latency per trial energy, so full 2 000-20 000 sweep runs are kept for a few cases only):
    every Tests/test_0x/1x/2x case at 300 sweeps, every volumeChange case at 500 sweeps"""
    import re
    cases = []
    for d in sorted(os.listdir(os.path.join(REF, "Tests"))):
        if re.match(r"test_(0|1|2)\d_", d):
            cases.append((d, os.path.join(REF, "Tests", d, "new"), 300))
    for d in sorted(os.listdir(os.path.join(REF, "Tests", "volumeChange"))):
        src = os.path.join(REF, "Tests", "volumeChange", d, "new")
        if os.path.isdir(src):
            cases.append(("volumeChange_" + d, src, 500))
    for name, src, ns in cases:
        tmp = tempfile.mkdtemp(prefix="short_")
        for fn in os.listdir(src):
            shutil.copy(os.path.join(src, fn), tmp)
        opt = open(os.path.join(tmp, "options")).read()
        opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = %d" % ns, opt)
        with open(os.path.join(tmp, "options"), "w") as f:
            f.write(opt)
        run([SC], tmp)
        shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, "%s.short%d.config.last" % (name, ns)))
        shutil.rmtree(tmp)
        print("short:", name, ns)


def do_equil():
    """runs WITH an equilibration phase (nequil > 0, adjust > 0): the reference adapts its maximum step sizes every `adjust` sweeps
    of the first half of the equilibration (Updater::optimizeStep / optimizeRot, mc/updater.cpp:238-253, 395-465), runs the second
    half and the production at the adapted sizes. NVT rods, bonded chains with chain moves, and an NPT case (box-edge step)."""
    import re
    cases = [("test_01_normal_PSC", os.path.join(REF, "Tests", "test_01_normal_PSC", "new")),
             ("test_20_chain_bond12", os.path.join(REF, "Tests", "test_20_chain_bond12", "new")),
             ("volumeChange_0h", os.path.join(REF, "Tests", "volumeChange", "0h", "new"))]
    for name, src in cases:
        tmp = tempfile.mkdtemp(prefix="equil_")
        for fn in os.listdir(src):
            shutil.copy(os.path.join(src, fn), tmp)
        opt = open(os.path.join(tmp, "options")).read()
        opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 100", opt)
        opt = re.sub(r"(?m)^nequil\s*=\s*\d+", "nequil = 120", opt)
        opt = re.sub(r"(?m)^adjust\s*=\s*\d+", "adjust = 10", opt)
        with open(os.path.join(tmp, "options"), "w") as f:
            f.write(opt)
        out = run([SC], tmp)
        shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, "%s.equil.config.last" % name))
        with open(os.path.join(HERE, "%s.equil.options" % name), "w") as f:
            f.write(opt)
        shutil.rmtree(tmp)
        print("equil:", name)


def do_membrane():
    """BASELINE.json configs[2]: the CPSC + SPN-SPA-SPA lipid membrane of Tests/SC_PSC_MEMBRANE_WANG (601 particles); the
    inputs are committed so that sc_b200.synth.membrane() can tile them to 265 k particles on the GPU box"""
    src = os.path.join(REF, "Tests", "SC_PSC_MEMBRANE_WANG", "scOOP_test")
    with open(os.path.join(REF, "Tests", "test_01_normal_PSC", "new", "options")) as f:
        opts = f.read()       # plain NVT options: the driver only evaluates energies (the shipped options switch Wang-Landau on)
    inputs = {"options": opts}
    for fn in ("top.init", "config.init"):
        with open(os.path.join(src, fn)) as f:
            inputs[fn] = f.read()
    tmp = tempfile.mkdtemp(prefix="mem_")
    for fn, txt in inputs.items():
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    run([DRIVER, "dump", "ref_dump.txt"], tmp)
    gz_copy(os.path.join(tmp, "ref_dump.txt"), os.path.join(HERE, "membrane601_init.ref.gz"))
    with gzip.GzipFile(os.path.join(HERE, "membrane601.inputs.json.gz"), "wb", mtime=0) as g:
        g.write(json.dumps(inputs).encode())
    shutil.rmtree(tmp)
    print("membrane601 done")


def do_ptype45():
    """pressureMove ptype 4 and 5 (mc/movecreator.cpp:480-550): the reference ships volumeChange inputs for ptype 0-3 only; these
    are Tests/volumeChange/0h and 0l with the coupling type switched, 500 sweeps by the unmodified reference program"""
    import re
    for pt in (4, 5):
        for hl in "hl":
            src = os.path.join(REF, "Tests", "volumeChange", "0" + hl, "new")
            name = "volumeChange_%d%s" % (pt, hl)
            tmp = tempfile.mkdtemp(prefix="pt45_")
            inputs = {}
            for fn in ("options", "top.init", "config.init"):
                with open(os.path.join(src, fn)) as f:
                    inputs[fn] = f.read()
            inputs["options"] = re.sub(r"(?m)^ptype\s*=\s*\d+", "ptype = %d" % pt, inputs["options"])
            with open(os.path.join(HERE, name + ".inputs.json"), "w") as f:
                json.dump(inputs, f)
            opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 500", inputs["options"])
            for fn, txt in (("options", opt), ("top.init", inputs["top.init"]), ("config.init", inputs["config.init"])):
                with open(os.path.join(tmp, fn), "w") as f:
                    f.write(txt)
            run([SC], tmp)
            shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, "%s.short500.config.last" % name))
            shutil.rmtree(tmp)
            print("ptype45:", name)


WALL_TYPES = """[Types]
PSC     1   PSC      1.0  1.2  1.346954458  0.5  80.0  5.0  3.0  0.0
CPSC    2   CPSC     1.4  1.0  1.12246205   1.0  80.0  5.0  3.0  0.0
CHPSC   3   CHPSC    1.0  1.2  1.346954458  0.5  90.0  5.0  3.0  0.0  10.0
CHCPSC  4   CHCPSC   3.5  1.0  1.12246205   1.0 170.0  5.0  3.0  0.0  10.0
TPSC    5   TPSC     1.0  1.2  1.346954458  0.5  80.0  5.0  3.0  0.0  180.0  60.0  5.0
TCPSC   6   TCPSC    1.0  1.2  1.346954458  0.5  80.0  5.0  3.0  0.0  90.0   60.0  5.0
TCHPSC  7   TCHPSC   1.0  1.2  1.346954458  0.5  80.0  5.0  3.0  0.0  180.0  60.0  5.0  10.0
TCHCPSC 8   TCHCPSC  1.0  1.2  1.346954458  0.5  80.0  5.0  3.0  0.0  90.0   60.0  5.0  10.0
SPN     9   SPN      1.0  1.2
SPA     10  SPA      1.0  1.2  1.346954458  0.5
[Molecules]
"""


def do_wall():
    """[EXTER] wall potential: ExternalEnergyCalculator::extere2 of every particle (oracle/ref_driver.cpp `exter`) on (a) the initial
    configuration of Tests/test_wallfibril and (b) a synthetic slab with every geotype at random heights and orientations around the
    wall (the wall is the plane z = 0 of the periodic box), plus a 300-sweep trajectory of test_wallfibril by the reference program"""
    import random
    import re
    rnd = random.Random(4711)
    names = ["PSC", "CPSC", "CHPSC", "CHCPSC", "TPSC", "TCPSC", "TCHPSC", "TCHCPSC", "SPN", "SPA"]
    per = 48
    top = WALL_TYPES
    for k, nm in enumerate(names):
        top += "%s: {\nparticles: %d\n}\n" % (nm, k + 1)
    top += "[System]\n" + "".join("%s %d\n" % (nm, per) for nm in names) + "[EXTER]\n5.0 4.00 1.0\n"
    box = (30.0, 30.0, 24.0)
    lines = ["%.8e %.8e %.8e" % box]
    for nm in names:
        for _ in range(per):
            z = rnd.uniform(-7.0, 7.0)
            pos = (rnd.uniform(-15, 15), rnd.uniform(-15, 15), z)
            while True:
                d = [rnd.gauss(0, 1) for _ in range(3)]
                n = math.sqrt(sum(x * x for x in d))
                if n > 1e-3:
                    break
            d = [x / n for x in d]
            if rnd.random() < 0.08:
                d = [1.0, 0.0, 0.0] if rnd.random() < 0.5 else [0.0, 0.0, 1.0]        # rods exactly parallel / perpendicular to the wall
            while True:
                p = [rnd.gauss(0, 1) for _ in range(3)]
                dp = sum(a * b for a, b in zip(p, d))
                p = [a - dp * b for a, b in zip(p, d)]
                n = math.sqrt(sum(x * x for x in p))
                if n > 1e-3:
                    break
            p = [x / n for x in p]
            lines.append(" ".join("%.8e" % x for x in list(pos) + d + p) + " 0")
    cfg = "\n".join(lines) + "\n"
    with open(os.path.join(REF, "Tests", "Interactions_tests", "options")) as f:
        opts = f.read()
    cases = {"wall_mix": {"options": opts, "top.init": top, "config.init": cfg}}
    src = os.path.join(REF, "Tests", "test_wallfibril", "new")
    cases["wall_fibril"] = {fn: open(os.path.join(src, fn)).read() for fn in ("options", "top.init", "config.init")}
    for name, inp in cases.items():
        tmp = tempfile.mkdtemp(prefix="wall_")
        for fn, txt in inp.items():
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        run([DRIVER, "exter", "ref_exter.txt"], tmp)
        gz_copy(os.path.join(tmp, "ref_exter.txt"), os.path.join(HERE, name + ".exter.gz"))
        run([DRIVER, "dump", "ref_dump.txt"], tmp)
        gz_copy(os.path.join(tmp, "ref_dump.txt"), os.path.join(HERE, name + "_init.ref.gz"))
        with gzip.GzipFile(os.path.join(HERE, name + ".inputs.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(inp).encode())
        if name == "wall_fibril":
            opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 300", inp["options"])
            with open(os.path.join(tmp, "options"), "w") as f:
                f.write(opt)
            run([SC], tmp)
            shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, "test_wallfibril.short300.config.last"))
        shutil.rmtree(tmp)
        print("wall:", name)


def do_wanglandau():
    """Tests/test_mempore (wlm = 2: hole in the xy plane of a lipid membrane, NPT ptype 2, chain moves) and Tests/test_pscthrough
    (wlm = 1: z position of a PSC crossing the membrane), 300 sweeps of the unmodified reference program: config.last and the
    Wang-Landau weights it writes (wl-new.dat). The order parameters stay the reference's host code above the calculator seam;
    through oracle/_ref/SC_scgpu every energy of these runs comes from the device (tests/test_gpu_dropin.py)."""
    for name in ("test_mempore", "test_pscthrough"):
        src = os.path.join(REF, "Tests", name, "new")
        inp = {fn: open(os.path.join(src, fn)).read() for fn in ("options", "top.init", "config.init", "wl.dat")}
        inp["options"] = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 300", inp["options"])
        inp["options"] = re.sub(r"(?m)^movie\s*=\s*\d+", "movie = 0", inp["options"])
        tmp = tempfile.mkdtemp(prefix="wl_")
        for fn, txt in inp.items():
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        run([SC], tmp)
        with gzip.GzipFile(os.path.join(HERE, name + ".inputs.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(inp).encode())
        gz_copy(os.path.join(tmp, "config.last"), os.path.join(HERE, name + ".short300.config.last.gz"))
        gz_copy(os.path.join(tmp, "wl-new.dat"), os.path.join(HERE, name + ".short300.wl-new.dat.gz"))
        shutil.rmtree(tmp)
        print("wanglandau:", name)


def do_othermoves():
    """Moves that stay the reference's host code above the calculator seam, on Tests/test_01_normal_PSC: geometric cluster moves
    (nClustMove = 10; MoveCreator::clusterMoveGeom, movecreator.cpp:54-172: allToAll + p2p) and grand-canonical insertion / deletion
    (nGrandCanon = 5 with an activity in top.init; movecreator.cpp:797-924: the particle count changes, update(EMResize)).
    300 sweeps of the unmodified reference program; through oracle/_ref/SC_scgpu every energy comes from the device.
    Cluster moves: the golden comes from the reference built with ITS OWN alternative calculator TotalEFull<PairE> (oracle/_ref/SC_full,
    totalenergycalculator.h:1228). clusterMoveGeom writes conf->pvec without calling update() (movecreator.cpp:164), so the default
    TotalEMatrix goes on with a STALE energy matrix after every cluster move (its trajectory differs from TotalEFull's from the first
    cluster move on; without cluster moves the two calculators give byte-identical runs). TotalEFull recomputes from conf->pvec."""
    src = os.path.join(REF, "Tests", "test_01_normal_PSC", "new")
    base = {fn: open(os.path.join(src, fn)).read() for fn in ("options", "top.init", "config.init")}
    base["options"] = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 300", base["options"])
    cases = {}
    c = dict(base)
    c["options"] = re.sub(r"(?m)^nClustMove\s*=\s*\d+", "nClustMove = 10", c["options"])
    cases["test_01_clustermoves"] = c
    c = dict(base)
    c["options"] = re.sub(r"(?m)^nGrandCanon\s*=\s*\d+", "nGrandCanon = 5", c["options"])
    c["top.init"] = c["top.init"].replace("particles:   1\n}", "particles:   1\nactivity: 0.05\n}")
    assert "activity" in c["top.init"]
    cases["test_01_grandcanonical"] = c
    for name, inp in cases.items():
        tmp = tempfile.mkdtemp(prefix="moves_")
        for fn, txt in inp.items():
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        run([os.path.join(ROOT, "oracle", "_ref", "SC_full") if name == "test_01_clustermoves" else SC], tmp)
        with open(os.path.join(HERE, name + ".inputs.json"), "w") as f:
            json.dump(inp, f)
        shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, name + ".short300.config.last"))
        shutil.rmtree(tmp)
        print("othermoves:", name)


def do_switchmoves():
    """Type-switch moves (MoveCreator::switchTypeMove, movecreator.cpp:233-303) on Tests/test_wallfibril: every CPSC (type 1) may switch to
    the chiral CHCPSC (type 2) with delta_mu = 5 -- the line the shipped top.init carries as a comment (`#particles: 1 2 5`) and
    Tests/test_wallfibril/old/options enables with switchprob -- at switchprob = 0.02 per step (~2400 attempts in 300 sweeps), external
    wall present. 300 sweeps of the unmodified reference program; config.last records every particle's `switched` flag."""
    src = os.path.join(REF, "Tests", "test_wallfibril", "new")
    inp = {fn: open(os.path.join(src, fn)).read() for fn in ("options", "top.init", "config.init")}
    inp["options"] = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 300", inp["options"])
    inp["options"] = re.sub(r"(?m)^switchprob\s*=\s*[0-9.]+", "switchprob = 0.02", inp["options"])
    assert "switchprob = 0.02" in inp["options"]
    lines = inp["top.init"].split("\n")
    k = [i for i, l in enumerate(lines) if l.strip().startswith("particles:")]
    assert len(k) == 1
    lines[k[0]] = "particles:   1        2           5"
    inp["top.init"] = "\n".join(lines)
    name = "test_wallfibril_switchmoves"
    tmp = tempfile.mkdtemp(prefix="switch_")
    for fn, txt in inp.items():
        with open(os.path.join(tmp, fn), "w") as f:
            f.write(txt)
    run([SC], tmp)
    with open(os.path.join(HERE, name + ".inputs.json"), "w") as f:
        json.dump(inp, f)
    shutil.copy(os.path.join(tmp, "config.last"), os.path.join(HERE, name + ".short300.config.last"))
    last = open(os.path.join(tmp, "config.last")).read().split("\n")[1:]
    print("switchmoves:", name, "particles switched at the end:", sum(1 for l in last if l.split() and l.split()[-1] == "1"))
    shutil.rmtree(tmp)


def do_verify_grids():
    """The pose grids of tests/golden/grid_config_*.txt.gz are written by the Python generator above; this check builds the REFERENCE's own
    generator (Tests/Interactions_tests/main.cpp -> Spc_config_generator) and replays Tests/Interactions_tests/test.sh's loops
    (Grid_energy_test_Spherocylinder.sh "..." 10 1.0 3.0; Grid_energy_test_Spherocylinder_Sphere.sh "..." 0.77) on it: the concatenated
    config.init texts must be byte-identical to ours. Writes tests/golden/grid_config.sha256 (checked by tests/test_oracle_golden.py)."""
    import hashlib
    tmp = tempfile.mkdtemp(prefix="grids_")
    gen = os.path.join(tmp, "Spc_config_generator")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-w", os.path.join(REF, "Tests", "Interactions_tests", "main.cpp"), "-o", gen])
    pre = open(os.path.join(REF, "Tests", "Interactions_tests", "config.init_prescript")).read()
    sc = pre
    for x in (1, 4, 7):                                   # seq 1.0 3.0 9.0
        for y in (1, 4, 7):
            for z in (1, 4, 7):
                out = subprocess.run([gen, str(x), str(y), str(z), "10", "1.0"], capture_output=True, text=True, check=True).stdout
                sc += "\n".join(out.split("\n")[:-2]) + "\n"       # head -n -1: the last line is the particle count
    seq = lambda hi: subprocess.run(["seq", "1.0", "0.77", hi], capture_output=True, text=True, check=True).stdout.split()
    scsp = pre
    for x in seq("9.0"):
        for y in seq("9.0"):
            for z in seq("5.0"):
                scsp += "%s %s %s %s\n" % (x, y, z, " 1 0 0  0 1 0  0 0")
    res = {}
    for kind, ref_text in (("sc", sc), ("scsp", scsp)):
        ours = gzip.open(os.path.join(HERE, "grid_config_%s.txt.gz" % kind), "rt").read()
        assert ours == ref_text, "pose grid %s differs from the reference generator's output" % kind
        res[kind] = hashlib.sha256(ref_text.encode()).hexdigest()
    with open(os.path.join(HERE, "grid_config.sha256"), "w") as f:
        json.dump({"what": "sha256 of the config.init text the reference's own Tests/Interactions_tests generator + test.sh loops produce "
                           "(tests/golden/make_golden.py verify_grids); grid_config_<kind>.txt.gz must hash to the same value", **res}, f, indent=1)
    shutil.rmtree(tmp)
    print("verify_grids: sc and scsp pose grids byte-identical to the reference generator's output", res)


def do_wlorder():
    """Wang-Landau order parameters of whole configurations, from the reference's own members (oracle/ref_driver.cpp `wlorder`):
    Tests/test_mempore (the membrane whose hole wlm 2 measures), Tests/test_pscthrough (wlm 1: a PSC through the membrane), the
    multi-type mixture extra_mix, and test_mempore's end state after 300 sweeps of the reference (a configuration off the lattice)."""
    for name in ("test_mempore", "test_pscthrough", "extra_mix", "test_mempore.short300"):
        base = name.split(".")[0]
        inp = json.loads(gzip.open(os.path.join(HERE, base + ".inputs.json.gz")).read().decode())
        tmp = tempfile.mkdtemp(prefix="wlo_")
        for fn in ("options", "top.init", "config.init"):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(inp[fn])
        if name.endswith(".short300"):
            with open(os.path.join(tmp, "config.init"), "wb") as f:
                f.write(gzip.open(os.path.join(HERE, name + ".config.last.gz")).read())
        run([DRIVER, "wlorder", "ref_wlorder.txt"], tmp)
        gz_copy(os.path.join(tmp, "ref_wlorder.txt"), os.path.join(HERE, name + ".wlorder.gz"))
        shutil.rmtree(tmp)
        print("wlorder:", name)


def main():
    if "verify_grids" in sys.argv[1:]:
        return do_verify_grids()
    if "wlorder" in sys.argv[1:]:
        return do_wlorder()
    if "switchmoves" in sys.argv[1:]:
        return do_switchmoves()
    if "othermoves" in sys.argv[1:]:
        return do_othermoves()
    if "wanglandau" in sys.argv[1:]:
        return do_wanglandau()
    if "wall" in sys.argv[1:]:
        return do_wall()
    if "ptype45" in sys.argv[1:]:
        return do_ptype45()
    if "membrane" in sys.argv[1:]:
        return do_membrane()
    if "extras" in sys.argv[1:]:
        return do_extras()
    if "short" in sys.argv[1:]:
        return do_short()
    if "equil" in sys.argv[1:]:
        return do_equil()
    if not (os.path.exists(DRIVER) and os.path.exists(SC)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    tests = sorted(d for d in os.listdir(os.path.join(REF, "Tests")) if d.startswith("test_"))
    if "grids" in sys.argv[1:]:
        tests = []
    for d in tests:
        do_case(d, os.path.join(REF, "Tests", d, "new"))
    for d in sorted(os.listdir(os.path.join(REF, "Tests", "volumeChange"))):
        if "grids" in sys.argv[1:]:
            break
        p = os.path.join(REF, "Tests", "volumeChange", d, "new")
        if os.path.isdir(p):
            do_case("volumeChange_" + d, p)
    with open(os.path.join(REF, "Tests", "Interactions_tests", "options")) as f:
        opts = f.read()
    names = list(SC_T)
    for a in range(len(names)):
        for b in range(a, len(names)):
            do_grid(names[a], SC_T[names[a]], names[b], SC_T[names[b]], "sc", opts)
    sn = list(SPHERE)
    for a in range(len(sn)):
        for b in range(a, len(sn)):
            do_grid(sn[a], SPHERE[sn[a]], sn[b], SPHERE[sn[b]], "sphere", opts)
    for a in names:
        for b in sn:
            do_grid(a, SC_T[a], b, SPHERE[b], "scsp", opts)
            do_grid(b, SPHERE[b], a, SC_T[a], "sc", opts)   # sphere first, rod second (role swap branch)


if __name__ == "__main__":
    main()
