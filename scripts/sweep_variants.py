#!/usr/bin/env python
"""Measurement driver for experimental builds of the sweep kernel (test infrastructure of the optimisation work, not product).

    python scripts/sweep_variants.py build            compile the variants here (nvcc cross-compiles; the .so files travel to the GPU box)
    python scripts/sweep_variants.py run              on the GPU box: time every variant, single replica and 8 replicas on one GPU

Each variant is libscgpu.so compiled with extra -D flags; it is loaded through the ordinary bindings via SCGPU_LIB_FAST."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build_variants")
VARIANTS = {
    "t320_mb12": ["-DSW_MINBLOCKS=12", "-DSW_TILE_N=320"],
    "t320_mb10": ["-DSW_MINBLOCKS=10", "-DSW_TILE_N=320"],
    "t320_mb16": ["-DSW_MINBLOCKS=16", "-DSW_TILE_N=224"],
}


def build():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "sc_b200", "csrc", "scgpu.cu")
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(OUT, "libscgpu_%s.so" % name)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-diag-suppress", "177", "-shared",
               "-Xcompiler", "-fPIC", "-fmad=true", "-DSCG_FAST_DIV"] + flags + ["-o", out, src, "-ldl"]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        assert p.wait() == 0


def run_one(name):
    import numpy as np
    from sc_b200 import Engine, synth
    from sc_b200.engine import MoveParams
    from sc_b200.host import HostSystem
    top, cfg, n = synth.psc_bulk()
    hs = HostSystem(top, cfg)
    mp = MoveParams()
    mp.temper = 0.1
    for k in range(40):
        mp.trans_mx[k] = 2.0 * 0.0212
        mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
    mp.n_sub = 1
    eng = Engine(0, "fast").load(hs)
    e0 = eng.all_to_all()
    de = 0.0
    acc = tot = 0
    for k in range(3):
        st = eng.sweep(mp, 12345, k)
        de += st.energy_delta
    e1 = eng.all_to_all()
    drift = abs((e1 - e0) - de) / max(1.0, abs(e1))
    nsw = 10
    eng.sync()
    t0 = time.perf_counter()
    for k in range(nsw):
        st = eng.sweep(mp, 12345, 3 + k)
        acc += st.trans_acc + st.rot_acc
        tot += st.trans_acc + st.rot_acc + st.trans_rej + st.rot_rej
    dt1 = time.perf_counter() - t0
    # asynchronous submission (no statistics read-back)
    t0 = time.perf_counter()
    for k in range(nsw):
        eng.sweep(mp, 12345, 13 + k, stats=False)
    eng.sync()
    dt1a = time.perf_counter() - t0
    R = 8
    engines = [eng] + [Engine(0, "fast").load(hs) for _ in range(R - 1)]
    for k in range(2):
        for r, e in enumerate(engines):
            e.sweep(mp, 777 + r, k, stats=False)
    for e in engines:
        e.sync()
    t0 = time.perf_counter()
    for k in range(nsw):
        for r, e in enumerate(engines):
            e.sweep(mp, 777 + r, 2 + k, stats=False)
    for e in engines:
        e.sync()
    dt8 = time.perf_counter() - t0
    mp.grid_k = 1
    for r, e in enumerate(engines):
        e.sweep(mp, 777 + r, 100, stats=False)
    for e in engines:
        e.sync()
    t0 = time.perf_counter()
    for k in range(nsw):
        for r, e in enumerate(engines):
            e.sweep(mp, 777 + r, 101 + k, stats=False)
    for e in engines:
        e.sync()
    dt8k1 = time.perf_counter() - t0
    print(json.dumps({"variant": name, "sweeps_per_s_8_grid_k1": R * nsw / dt8k1, "sweeps_per_s_1": nsw / dt1, "sweeps_per_s_1_async": nsw / dt1a, "sweeps_per_s_8": R * nsw / dt8, "acceptance": acc / max(1, tot),
                      "trials_per_sweep": tot / nsw, "bookkeeping_rel_err": drift, "e0": e0, "e1": e1}), flush=True)


def run():
    for name in list(VARIANTS) + ["default"]:
        env = dict(os.environ)
        if name != "default":
            env["SCGPU_LIB_FAST"] = os.path.join(OUT, "libscgpu_%s.so" % name)
        subprocess.call([sys.executable, os.path.abspath(__file__), "one", name], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        run()
    else:
        run_one(sys.argv[2])
