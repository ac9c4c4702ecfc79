#!/usr/bin/env python
"""bench.py -- throughput of the scOOP energy hot path on B200, with its roofline and the CPU reference beside it.

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code on this box's host cores

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "L"): 65 536 PSC rods, bulk NVT lattice, box 89.6 x 89.6 x 70.4,
synthetic (sc_b200/synth.py, seed 12345). A STEP = one pass of the pair-energy path over the whole system: the
one-to-all trial energy of every particle against its 27-cell neighbourhood (N = 65 536 oneToAllTrial evaluations,
one warp each, one kernel launch). Metric = gated pair-energy evaluations per second (pairs that pass the
reference's sqmaxcut gate and reach a functor, PairE::operator(), scOOP/mc/paire.h:1209-1220).
At N > 1 every rank holds its own replica of the energy workload (weak scaling, no data-path collective); the parallel-tempering
leg (secondary.parallel_tempering, BASELINE configs[4]) runs 8 replicas on the N GPUs with a replica exchange through the
C ABI (scgpu_replica_exchange: device-side records, one ncclAllGather, device decision) every 10 sweeps, timed inside.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pair_energy_evals_per_s"
UNIT = "gated pair-energy evals/s"
WORKLOAD = "65536 PSC bulk NVT (configs[1]), one-to-all trial energy of every particle over cell lists"
FLOP_WEIGHTS = {"add": 1, "mul": 1, "div": 1, "sqrt": 1, "cos": 20, "acos": 20, "pow": 3}


def workload_texts():
    from sc_b200 import synth
    return synth.psc_bulk()     # (top_text, config_text, n)


def options_text():
    # the reference needs an `options` file; these are Tests/test_01_normal_PSC's values
    return """ptype = 1
press = 0
paralpress = 0
shave = 0
nequil = 0
adjust = 0
nsweeps  = 2
paramfrq = 0
report   = 0
nrepchange = 0
nGrandCanon = 0
nClustMove = 0
movie    = 0
chainprob = 0.0
transmx = 0.0212
rotmx = 7.5
edge_mx = 0.0
chainmmx = 0.0
chainrmx = 0.0
temper = 0.1
paraltemper = 0.1
wlm = 0
wlmtype = 0
switchprob = 0.00
pairlist_update = 10
seed = 145658
write_cluster = 0
"""


# ------------------------------------------------------------------------------------------------
# CPU reference timing (oracle/_ref/sc_ref_fast = our driver + the unmodified reference sources, -Ofast)
# ------------------------------------------------------------------------------------------------
def cpu_reference(ntargets, reps, nproc):
    """Runs `nproc` independent copies of the reference's oneToAll loop (its only parallel model is independent
    replicas/ranks) over a bounded sample: `ntargets` evenly spaced trial particles x `reps` repetitions each."""
    top, cfg, n = workload_texts()
    drv = os.path.join(ROOT, "oracle", "_ref", "sc_ref_fast")
    if not os.path.exists(drv):
        return None
    tmp = tempfile.mkdtemp(prefix="scref_")
    try:
        with open(os.path.join(tmp, "top.init"), "w") as f:
            f.write(top.replace("A %d" % n, "A 1"))      # one molecule; the driver replicates it n times (MAXN bypass)
        with open(os.path.join(tmp, "config.init"), "w") as f:
            f.write(cfg)
        with open(os.path.join(tmp, "options"), "w") as f:
            f.write(options_text())
        procs = [subprocess.Popen([drv, "time", str(ntargets), str(reps), str(n)], cwd=tmp, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for _ in range(nproc)]
        outs = [p.communicate()[0] for p in procs]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    recs = []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("REFJSON "):
                recs.append(json.loads(line[8:]))
    if len(recs) != nproc:
        return None
    gated = sum(r["gated_pairs"] for r in recs)
    tmax = max(r["one_to_all_s"] for r in recs)
    return {"value": gated / tmax, "unit": UNIT, "cores": nproc, "kind": "reference",
            "sample": "%d trial particles x %d reps per process, %d independent processes (reference is single-threaded); "
                      "TotalEFull<PairE>::oneToAll over the reference's neighbour lists, -Ofast -march=x86-64-v3; "
                      "list construction (%.2f s per process for the sample) not charged" % (ntargets, reps, nproc, recs[0]["list_build_s"]),
            "seconds": tmax, "gated_pairs": gated}


def cpu_reference_membrane(mtop, mcfg, ntargets=512):
    """The reference's per-particle energy loop on the 265 041-particle membrane (configs[2]), one process, bounded sample: `ntargets`
    evenly spaced particles over the reference's neighbour lists. The reference's allToAll() visits all N^2/2 pairs without lists
    (totalenergycalculator.h:502-521) -- hours at this size; N/2 x the measured per-particle time is a LOWER bound of it."""
    import re as _re
    drv = os.path.join(ROOT, "oracle", "_ref", "sc_ref_fast")
    if not os.path.exists(drv):
        return None
    m = _re.search(r"\[System\]\s*\n\s*(\w+)\s+(\d+)\s*\n\s*(\w+)\s+(\d+)", mtop)
    if not m:
        return None
    counts = [int(m.group(2)), int(m.group(4))]
    top1 = mtop[:m.start()] + "[System]\n%s 1\n%s 1\n" % (m.group(1), m.group(3))      # one molecule of each type; the driver replicates (MAXN bypass)
    tmp = tempfile.mkdtemp(prefix="scref_mem_")
    try:
        for fn, txt in (("top.init", top1), ("config.init", mcfg), ("options", options_text())):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        out = subprocess.run([drv, "time", str(ntargets), "1"] + [str(c) for c in counts], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600).stdout
    except Exception:
        return None
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    rec = None
    for line in out.splitlines():
        if line.startswith("REFJSON "):
            rec = json.loads(line[8:])
    if not rec:
        return None
    per = rec["one_to_all_s"] / max(1, rec["targets"])
    return {"kind": "reference", "cores": 1, "us_per_particle_one_to_all": per * 1e6, "all_to_all_ms_lower_bound": per * rec["n"] / 2 * 1e3,
            "sample": "%d evenly spaced particles, TotalEFull<PairE>::oneToAll over the reference's neighbour lists, -Ofast; list construction "
                      "(%.1f s for the sample) not charged; the reference's own allToAll() loops over all N^2/2 = 3.5e10 pairs without lists, "
                      "N/2 x the per-particle time is a lower bound" % (rec["targets"], rec["list_build_s"])}


def cpu_reference_sweeps(ntargets, ntrials, nproc):
    """Measured sequential sweeps of the reference's own move code (MoveCreator::partDisplace / partRotate over TotalEFull<PairE> and the
    reference's neighbour lists; oracle/_ref/sc_ref_full, see oracle/ref_driver.cpp do_sweep): `nproc` independent processes (the
    reference is single-threaded; independent replicas are its parallel model), each `ntrials` trial moves among `ntargets` particles."""
    top, cfg, n = workload_texts()
    drv = os.path.join(ROOT, "oracle", "_ref", "sc_ref_full")
    if not os.path.exists(drv):
        return None
    tmp = tempfile.mkdtemp(prefix="scsweep_")
    try:
        with open(os.path.join(tmp, "top.init"), "w") as f:
            f.write(top.replace("A %d" % n, "A 1"))
        with open(os.path.join(tmp, "config.init"), "w") as f:
            f.write(cfg)
        with open(os.path.join(tmp, "options"), "w") as f:
            f.write(options_text())
        procs = [subprocess.Popen([drv, "sweep", str(ntargets), str(ntrials), str(n)], cwd=tmp, stdout=subprocess.PIPE,
                                  stderr=subprocess.DEVNULL, text=True) for _ in range(nproc)]
        outs = [p.communicate()[0] for p in procs]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    recs = [json.loads(line[8:]) for o in outs for line in o.splitlines() if line.startswith("REFJSON ")]
    if len(recs) != nproc:
        return None
    tmax = max(r["trial_s"] for r in recs)
    return {"sweeps_per_s": nproc * ntrials / tmax / n, "sweeps_per_s_one_core": float(np.mean([r["sweeps_per_s_one_core"] for r in recs])), "cores": nproc,
            "us_per_trial_one_core": tmax / ntrials * 1e6, "acceptance": recs[0]["accepted"] / max(1, recs[0]["accepted"] + recs[0]["rejected"]),
            "kind": "reference", "sample": "%d trial moves among %d particles per process, %d independent processes; MoveCreator::partDisplace/partRotate over "
            "TotalEFull<PairE> (old AND trial energy evaluated per move: the reference's default TotalEMatrix caches the old one but needs 34 GB at this size) "
            "and the reference's neighbour lists, -Ofast; list construction (%.2f s per process for the sample) not charged" % (ntrials, ntargets, nproc, recs[0]["list_build_s"])}


def cpu_port_fallback(ntargets):
    """oracle port, 1 core -- only if oracle/_ref did not travel"""
    from oracle import oracle as O
    top, cfg, n = workload_texts()
    s = O.system_from_text(top, cfg)
    cells = s.cells()
    t0 = time.perf_counter()
    gated = 0
    for t in range(0, n, max(1, n // ntargets)):
        gated += s.one_to_all_cells(t, cells)[2]
    dt = time.perf_counter() - t0
    return {"value": gated / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": "%d trial particles, C oracle over cell lists" % ntargets,
            "seconds": dt, "gated_pairs": gated}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock and throttle reasons sampled DURING the timed region. The region is a few milliseconds long, far shorter than the
    start-up of an `nvidia-smi -lms` child, so the samples come from NVML in this process (what nvidia-smi itself reads): one
    sample when the region opens, a polling thread while it runs, one sample when it closes."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.sm, self.mx, self.reasons = [], None, set()
        self.h = None
        self.run = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:
            self.err = repr(ex)
            self.h = None
            return
        self._sample()
        self.run = True
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _sample(self):
        try:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while self.run:
            self._sample()
            time.sleep(0.0005)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")], "samples": 0}
        self.run = False
        self.t.join(timeout=1.0)
        self._sample()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "NVML (nvmlDeviceGetClockInfo / CurrentClocksEventReasons) polled in-process over the timed region"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # one process per GPU on a multi-socket host: run on the CPUs next to the GPU, so that the page-locked buffers of the end-to-end leg
    # are first touched on its NUMA node (NVML knows the topology). The CPU legs below lift the restriction again.
    full_affinity = None
    if world > 1 or os.environ.get("BENCH_AFFINITY") == "1":
        try:
            import pynvml
            pynvml.nvmlInit()
            full_affinity = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            full_affinity = None
    from sc_b200 import Engine
    from sc_b200.host import HostSystem
    top, cfg, n = workload_texts()
    hs = HostSystem(top, cfg)                 # product-side parser (C++ host mirror), not the oracle
    eng = Engine(local, "fast")
    eng.load(hs)
    if world > 1:
        # parallel tempering: every replica starts from the same lattice but lives at its own temperature; for the
        # energy pass only the configuration matters, so perturb it per rank to make the replicas distinct
        rng = np.random.default_rng(1000 + rank)
        st = hs.state.copy()
        st[:, 0:3] += rng.normal(scale=2e-4, size=(n, 3))
        eng.set_particles(st, hs.type, hs.moltype)
    _, ncand, ngate = eng.one_to_all_everyone(fetch=True, count=True)   # also builds the cell list
    peak = eng.fp64_peak()
    import torch

    def step():
        eng.one_to_all_everyone(fetch=False)

    for _ in range(max(3, args.warmup)):
        step()
    eng.sync()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    clocks = Clocks(local) if rank == 0 else None
    l0 = eng.launches()
    kernel_ms = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()                        # inputs (18 MB) are smaller than L2: flush between timed iterations
        eng.timer_start()
        step()
        kernel_ms.append(eng.timer_stop())
    stage_us = eng.profile_everyone()         # per-launch microseconds of one more pass, taken while the GPU is still at speed
    eng.sync()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    wall = time.perf_counter() - t_wall0
    launches = eng.launches() - l0 - args.steps            # minus the flush launches (not part of the step)
    clk = clocks.stop() if clocks else None
    total_ms = float(np.sum(kernel_ms))
    pairs = float(ngate) * args.steps
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        p = torch.tensor([pairs, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
        pairs, launches = float(p[0].item()), int(p[1].item())
    value = pairs / (total_ms * 1e-3)

    # ---- end to end through the C ABI with HOST buffers: H2D of the configuration, cell build, energy pass, D2H of energies
    # page-locked host buffers (the contract's "pinned host memory"): the library DMAs straight from them. The upload is what a
    # configuration file holds -- 9 doubles per particle (position, direction, patch direction); patch sides are derived on the device
    # exactly as the reference's partVecInit does after reading config.init.
    import torch as _t
    pinned = _t.empty((n, 9), dtype=_t.float64).pin_memory()
    state_host = pinned.numpy()
    state_host[:] = hs.state[:, :9]
    e_pinned = _t.empty((n,), dtype=_t.float64).pin_memory()
    e_host = e_pinned.numpy()
    e2e_steps = max(4, min(args.steps, 20))
    eng.set_particles_compact(state_host, hs.type, hs.moltype)        # types travel once: they never change along a Monte Carlo run
    for _ in range(2):
        eng.set_particles_compact(state_host, None, None)
        eng.one_to_all_everyone(fetch=True, out=e_host)
    if world > 1:
        dist.barrier()
    # (a) one configuration at a time: upload -> cell build -> energies -> read-back, each step waits for the previous one
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.set_particles_compact(state_host, None, None)
        eng.one_to_all_everyone(fetch=True, out=e_host)
    eng.sync()
    e2e_serial_s = time.perf_counter() - t0
    e_serial = e_host.copy()
    # (b) the headline: the same per-step work through scgpu_submit_everyone on FOUR contexts used in turn (further replicas
    # of the system on the same GPU), so the host<->device copies of one configuration overlap the kernels of the other.
    # Every step still uploads its configuration from page-locked memory and reads its energies back inside the timed region.
    nctx = int(os.environ.get("BENCH_E2E_CONTEXTS", "4"))
    engs, ins, outs = [eng], [pinned], [e_pinned]
    for _ in range(nctx - 1):
        e2 = Engine(local, "fast")
        e2.load(hs)
        e2.set_particles_compact(state_host, hs.type, hs.moltype)
        engs.append(e2)
        ins.append(_t.empty((n, 9), dtype=_t.float64).pin_memory())
        ins[-1].numpy()[:] = state_host
        outs.append(_t.empty((n,), dtype=_t.float64).pin_memory())

    def submit(k):
        engs[k].submit_everyone(ins[k].numpy(), outs[k].numpy())

    for k in list(range(nctx)) * 2:          # warm-up; a list that had to grow is reported by sync(): submit again
        for attempt in range(4):
            submit(k)
            try:
                engs[k].sync()
                break
            except RuntimeError:
                continue
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        k = i % nctx
        if i >= nctx:
            engs[k].sync()                   # the previous step of this context is complete: its energies are in outs[k]
        submit(k)
    for e_ in engs:
        e_.sync()
    e2e_s = time.perf_counter() - t0
    assert all(np.array_equal(o.numpy(), e_serial) for o in outs), "pipelined e2e differs from the serial call"
    for e_ in engs[1:]:
        e_.close()
    if world > 1:
        t = torch.tensor([e2e_s, e2e_serial_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_serial_s = float(t[0].item()), float(t[1].item())
    e2e_val = float(ngate) * e2e_steps * world / e2e_s
    h2d = state_host.nbytes
    d2h = e_host.nbytes

    # ---- latency of ONE calculator call: what the literal drop-in pays per virtual (integration/totalegpu.h issues one scgpu_one_to_all
    # per oneToAll / oneToAllTrial: upload of the trial record, three launches, read-back, one synchronisation)
    eng.set_particles(hs.state, hs.type, hs.moltype)
    eng.build_cells()
    trial = hs.state[1234].copy()
    trial[0:3] += 1e-4
    for k in range(50):
        eng.one_to_all(1234, trial)
    t0 = time.perf_counter()
    ncall = 400
    for k in range(ncall):
        eng.one_to_all(1234 + 97 * k, trial if (k & 1) else None)
    single_call = {"us_per_call": (time.perf_counter() - t0) / ncall * 1e6, "calls": ncall,
                   "what": "scgpu_one_to_all on the 65 536-particle system through ctypes (trial state every other call): the cost of one oneToAll()/oneToAllTrial() "
                           "virtual of the drop-in class; a sequential trial move costs two of them plus update() on acceptance"}

    # ---- cell-list build (counting sort by cell + SoA permute): the HBM-bound kernel group of the path
    cb_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        eng.flush_l2()
        eng.timer_start()
        eng.build_cells()
        cb_ms.append(eng.timer_stop())
    cb_bytes = 544.0 * n          # 24 B position + 8 B cell id + 2 x 256 B record per particle
    peaks, peak_kind = measured_peaks()
    cell_build = {"bound": "hbm", "ms": float(np.mean(cb_ms)), "achieved": cb_bytes / (np.mean(cb_ms) * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                  "unit": "GB/s", "frac": cb_bytes / (np.mean(cb_ms) * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                  "bytes_per_build": cb_bytes, "kernels": "k_cell_count (histogram + scan through a last-block ticket) + k_cell_fill + k_cell_place (3 launches; launch latency is a large share at 65k particles)"}

    # ---- full-system energy (allToAll: what every NPT volume move and every replica exchange needs): every pair once
    fe_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        eng.flush_l2()
        eng.timer_start()
        eng.all_to_all(fetch=False)
        fe_ms.append(eng.timer_stop())
    full_energy = {"workload": "allToAll over 65536 PSC (every unordered pair once, conlist of the higher index)", "ms": float(np.mean(fe_ms)),
                   "gated_pairs": int(ngate) // 2, "pair_evals_per_s": (int(ngate) // 2) / (np.mean(fe_ms) * 1e-3)}
    if args.with_membrane and world == 1:
        import gzip as _gz
        inp = json.loads(_gz.open(os.path.join(ROOT, "tests", "golden", "membrane601.inputs.json.gz")).read().decode())
        from sc_b200 import synth as _synth
        mtop, mcfg, mn = _synth.membrane(21, 21, inp["top.init"], inp["config.init"])
        mhs = HostSystem(mtop, mcfg)
        meng = Engine(local, "fast").load(mhs)
        meng.all_to_all()
        mm = []
        for _ in range(3):
            meng.flush_l2()
            meng.timer_start()
            meng.all_to_all(fetch=False)
            mm.append(meng.timer_stop())
        full_energy["membrane_265041"] = {"workload": "BASELINE configs[2]: CPSC + SPN-SPA-SPA lipid membrane tiled 21x21 (265 041 particles), one allToAll = one NPT trial energy",
                                          "ms": float(np.mean(mm))}
        # sweeps of the membrane with the reference's move mix for lipids: single-bead moves + rigid moves of whole SPN-SPA-SPA molecules
        from sc_b200.engine import MoveParams as _MP, ChainMoves as _CM
        mmp, mcm = _MP(), _CM()
        mmp.temper, mmp.n_sub, mcm.chainprob = 1.0, 1, 0.5
        for k in range(40):
            mmp.trans_mx[k] = 0.1
            mmp.rot_angle[k] = 10.0 / 180.0 * 1.5707963267948966 * 0.5
        for k in range(32):
            mcm.chainm_mx[k] = 0.2
            mcm.chainr_angle[k] = 10.0 / 180.0 * 1.5707963267948966
        for k in range(2):
            meng.sweep_chains(mmp, mcm, 4242, k)
        t0m = time.perf_counter()
        nacc = ntot = 0
        for k in range(3):
            st_, cst_ = meng.sweep_chains(mmp, mcm, 4242, 2 + k)
            nacc += st_.trans_acc + st_.rot_acc + cst_.chainm_acc + cst_.chainr_acc
            ntot += st_.trans_acc + st_.rot_acc + st_.trans_rej + st_.rot_rej + cst_.chainm_acc + cst_.chainr_acc + cst_.chainm_rej + cst_.chainr_rej
        msw = (time.perf_counter() - t0m) / 3
        full_energy["membrane_265041"]["sweep_with_chain_moves"] = {
            "ms_per_sweep": msw * 1e3, "sweeps_per_s": 1.0 / msw, "trial_moves_per_s": ntot / 3 / msw, "acceptance": nacc / max(1, ntot),
            "chainprob": 0.5, "temper": 1.0,
            "note": "scgpu_sweep_checkerboard_chains: single-bead passes (k_sweep_cells, fine grid) + rigid moves of whole lipids (k_sweep_chain_colour, coarse grid)"}
        # Wang-Landau order parameters of that configuration on the device (row (f)3): the membrane hole (wlm 2, mesh of sigma / 3) + the
        # z distance of particle 0 from the centre of mass (wlm 1) in one call; after the sweeps the newest configuration is device-resident
        try:
            tail_t = int(mhs.type[-1])
            msz = float(mhs.ia[tail_t, tail_t, 3]) / 3.0
            wl0 = meng.wl_order((2, 1), wlmtype=tail_t, minorder=(0.0, -30.0), dorder=(1.0, 0.01), meshsize=msz)
            t0w = time.perf_counter()
            for _ in range(20):
                wl0 = meng.wl_order((2, 1), wlmtype=tail_t, minorder=(0.0, -30.0), dorder=(1.0, 0.01), meshsize=msz)
            wl_us = (time.perf_counter() - t0w) / 20 * 1e6
            wlo = {"us_per_call": wl_us, "mesh": [int(wl0.mesh_dim[0]), int(wl0.mesh_dim[1])], "largest_hole_mesh_points": int(wl0.raw[0]),
                   "what": "scgpu_wl_order, wlm = (2, 1): mass centre + mesh fill + union-find hole search (7 launches) + one 88-byte read-back, synchronous call through ctypes"}
            if not args.no_cpu:
                from oracle import oracle as _O
                st_now = meng.download_particles()
                osys = _O.System(st_now, mhs.type, mhs.moltype, mhs.ia, mhs.mol, mhs.box, mhs.sqmaxcut, mhs.maxcut)
                t0w = time.perf_counter()
                for _ in range(5):
                    href = _O.wl_raw(osys, 2, wlmtype=tail_t, meshsize=msz)
                    zref = _O.wl_raw(osys, 1)
                wlo["cpu_port_us_per_call"] = (time.perf_counter() - t0w) / 5 * 1e6
                wlo["cpu_note"] = "oracle/wl_order.c (Mesh::meshInit + Conf::massCenter restated, one core, -O2); the configuration is already on the host for it"
                wlo["matches_cpu"] = bool(int(href[0]) == int(wl0.raw[0]) and abs(zref - wl0.raw[1]) < 1e-9)
            full_energy["membrane_265041"]["wl_order"] = wlo
        except Exception as e:                       # a secondary leg must not take the headline down with it
            full_energy["membrane_265041"]["wl_order"] = {"error": repr(e)[:200]}
        if not args.no_cpu:
            mref = cpu_reference_membrane(mtop, mcfg)
            if mref:
                full_energy["membrane_265041"]["cpu_reference"] = mref
        meng.close()
        mhs.close()

    # ---- second metric of BASELINE.json: MC sweeps/s (batched checkerboard displacement/rotation sweeps, N trials each)
    from sc_b200.engine import MoveParams
    mp = MoveParams()
    mp.temper = 0.1
    for k in range(40):
        mp.trans_mx[k] = 2.0 * 0.0212
        mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
    mp.n_sub = 1
    nsw = max(3, min(args.steps, 10))

    def time_sweeps(rule):
        mp.trial_rule = rule
        eng.set_particles(hs.state, hs.type, hs.moltype)
        for k in range(2):
            eng.sweep(mp, 12345 + rank, k)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        a = t = 0
        for k in range(nsw):
            st = eng.sweep(mp, 12345 + rank, 2 + k)
            a += st.trans_acc + st.rot_acc
            t += st.trans_acc + st.rot_acc + st.trans_rej + st.rot_rej
        return time.perf_counter() - t0, a, t

    rule0_s, _, _ = time_sweeps(0)           # equal number of trials per non-empty cell, drawn with replacement
    sweep_s, acc, tot = time_sweeps(2)       # the headline: every particle exactly once per sweep, random order inside its cell
    if world > 1:
        t = torch.tensor([sweep_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sweep_s = float(t.item())
    sweeps = {"metric": "mc_sweeps_per_s", "value": nsw * world / sweep_s, "unit": "sweeps/s (1 sweep = N = 65536 trial moves per replica)",
              "ms_per_sweep": sweep_s / nsw * 1e3, "trial_moves_per_s": tot * world / sweep_s, "acceptance": acc / max(1, tot),
              "temper": 0.1, "transmx": 0.0212, "rotmx_deg": 7.5, "sweeps_timed": nsw,
              "trial_rule": 2, "ms_per_sweep_trial_rule_0": rule0_s / nsw * 1e3,
              "note": "whole scgpu_sweep_checkerboard call: shifted cell build + 8 colour passes + statistics read-back. trial_rule 2 (every particle exactly "
                      "once per sweep in a fresh random order inside its cell) runs every pass as four dense launches (sweep_phased.cuh: k_sweep_propose -> "
                      "k_cheap_flat -> k_patch_flat -> k_sweep_resolve); rule 0 (the same number of trials in every non-empty cell, with replacement) runs "
                      "the round kernel (k_sweep_rounds: a block per active cell, up to 16 trials evaluated at once and resolved in sequence)"}

    # ---- BASELINE configs[4]: parallel tempering, 8 replicas x 65 536 PSC on `world` GPUs (8 / world replicas per GPU, each on its
    # own stream), one scgpu_replica_exchange every nrepchange = 10 sweeps: allToAll() of every replica, records packed on the device,
    # ONE ncclAllGather, device decision kernel (sc_b200/csrc/comm.cuh). Timed on the device: events on every replica's stream
    # bracket the loop (all streams idle at the start), max over replicas and ranks.
    from sc_b200 import replica as _rep
    from sc_b200.engine import Comm
    R_total = 8
    nlocal = max(1, R_total // world)
    engines = [eng] + [Engine(local, "fast").load(hs) for _ in range(nlocal - 1)]
    for e in engines:
        e.set_particles(hs.state, hs.type, hs.moltype)
    if world > 1:
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, device=torch.device("cuda", local))
        comm = Comm(local, world, rank, box[0])
    else:
        comm = Comm(local, 1, 0)
    ptr = _rep.ParallelTempering(comm, engines, 0.1, 0.13, 0.0212, 7.5, nrepchange=10, seed=145658)
    nsw_pt = 20 if args.steps >= 10 else 10
    for k in range(1, 31):                   # warm-up: 30 sweeps and three exchanges (NCCL sets its channels up lazily over the first collectives)
        ptr.sweep(k)
    for e in engines:
        e.sync()
    ptr.exchange_us.clear()
    acc0, rej0 = sum(ptr.acc), sum(ptr.rej)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    for e in engines:
        e.timer_start()
    t0 = time.perf_counter()
    t_exch = 0.0
    for k in range(31, 31 + nsw_pt):
        if k % 10 == 0:
            te = time.perf_counter()
            ptr.sweep(k)
            t_exch += time.perf_counter() - te      # host view of a sweep submission + the whole exchange call
        else:
            ptr.sweep(k)
    pt_ms = max(e.timer_stop() for e in engines)
    pt_wall = time.perf_counter() - t0
    acc1, rej1 = sum(ptr.acc) - acc0, sum(ptr.rej) - rej0
    vals = [pt_ms, float(np.mean(ptr.exchange_us)) if ptr.exchange_us else 0.0, float(acc1), float(rej1)]
    if world > 1:
        t = torch.tensor(vals[:2], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c2 = torch.tensor(vals[2:], dtype=torch.float64, device="cuda")
        dist.all_reduce(c2, op=dist.ReduceOp.SUM)
        vals = [float(t[0].item()), float(t[1].item()), float(c2[0].item()), float(c2[1].item())]
    pt_ms, exch_us, acc1, rej1 = vals
    n_exch = nsw_pt // 10
    # the full energy of one replica alone (what every exchange needs first), for the share of the exchange that is energy
    fe = []
    for _ in range(3):
        eng.timer_start()
        eng.all_to_all(fetch=False)
        fe.append(eng.timer_stop())
    pt = {"workload": "BASELINE configs[4]: parallel tempering, %d replicas x %d PSC, T = 0.1 .. 0.13 (ladder of sim.h:389-393), exchange every 10 sweeps" % (ptr.R, n),
          "replicas": ptr.R, "gpus": world, "replicas_per_gpu": nlocal, "sweeps_timed": nsw_pt, "exchanges_timed": n_exch,
          "aggregate_sweeps_per_s": ptr.R * nsw_pt / (pt_ms * 1e-3), "ms_per_sweep_all_replicas": pt_ms / nsw_pt,
          "trial_moves_per_s": ptr.R * nsw_pt * n / (pt_ms * 1e-3),
          "exchange_us_after_energy": exch_us, "all_to_all_us_one_replica": float(np.mean(fe)) * 1e3,
          "exchange_share_of_time": (exch_us + float(np.mean(fe)) * 1e3 * nlocal) * n_exch * 1e-3 / pt_ms,
          "exchange_call_host_ms": t_exch / max(1, n_exch) * 1e3,
          "pair_acceptance": acc1 / max(1.0, acc1 + rej1), "pairs_attempted": int((acc1 + rej1) // 2),
          "analytic_estimate": _rep.switch_probability_estimate(n, ptr.dtemp),
          "analytic_note": "exp(-0.5 N dT^2 / (1 + dT)) as the reference prints it (mc/inicializer.cpp:52-55); with N = 65 536 and 8 rungs between T = 0.1 and 0.13 it is ~0: the ladder of BASELINE configs[4] is far too coarse for this system size, exchanges are attempted and (correctly) rejected",
          "timing": "CUDA events on every replica's stream around the whole loop (sweeps + exchanges), max over replicas and ranks; wall %.4f s" % pt_wall,
          "path": "scgpu_replica_exchange (C ABI): device-side records, %s, device decision kernel; no host round trip of the energies" % ("one ncclAllGather of %d x 512 B" % ptr.R if world > 1 else "one process: device copy instead of NCCL")}
    # the same loop on a ladder the system size allows (the reference's estimate exp(-0.5 N dT^2 / (1 + dT)) ~ 0.3): exchanges do happen
    ptt = _rep.ParallelTempering(comm, engines, 0.1, 0.100426, 0.0212, 7.5, nrepchange=10, seed=145658)
    for k in range(1001, 1041):
        ptt.sweep(k)
    tv = [float(sum(ptt.acc)), float(sum(ptt.rej))]
    if world > 1:
        c3 = torch.tensor(tv, dtype=torch.float64, device="cuda")
        dist.all_reduce(c3, op=dist.ReduceOp.SUM)
        tv = [float(c3[0].item()), float(c3[1].item())]
    pt["tight_ladder"] = {"T": [0.1, 0.100426], "sweeps": 40, "pairs_attempted": int((tv[0] + tv[1]) // 2), "pair_acceptance": tv[0] / max(1.0, tv[0] + tv[1]),
                          "analytic_estimate": _rep.switch_probability_estimate(n, ptt.dtemp),
                          "note": "all replicas start from the same lattice and have not equilibrated at their rungs: a plausibility check of the exchange rule, not a converged ratio"}
    comm.close()
    for e in engines[1:]:
        e.close()
    sweeps["parallel_tempering"] = pt

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if full_affinity is not None:
        try:
            os.sched_setaffinity(0, full_affinity)       # the CPU reference legs use every host core
        except Exception:
            pass
    # ---- roofline of the dominant kernel (k_one_to_all): FP64 pipe. Algorithmic flops counted by the op-counting oracle
    roof = {"bound": "fp64", "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None, "traffic": None}
    try:
        from oracle import oracle as O
        s = O.system_from_text(top, cfg)
        sample = sorted(np.random.default_rng(2024).choice(n, 1024, replace=False).tolist())      # unbiased: random targets, not a lattice stride
        d = O.count_flops(s, sample)
        flops_sample = sum(FLOP_WEIGHTS[k] * d[k] for k in FLOP_WEIGHTS)
        gate_sample = sum(FLOP_WEIGHTS[k] * d["gate"][k] for k in FLOP_WEIGHTS)
        scale = n / len(sample)
        flops_step, flops_gate = flops_sample * scale, gate_sample * scale
        flops_functor = flops_step - flops_gate
        kms = float(np.mean(kernel_ms))
        ach = flops_step / (kms * 1e-3) / 1e12
        us = stage_us                         # one pass with events between the launches (warm L2; shares, not absolutes)
        us_tot = sum(us)
        fp64_us = us[1] + us[2]
        roof.update({"achieved": ach, "frac": ach / peak, "flops_per_launch": flops_step,
                     "flops_gate": flops_gate, "flops_functor": flops_functor,
                     "flops_per_candidate_gate": gate_sample / max(1, d["candidates"]),
                     "flops_per_gated_pair_functor": (flops_sample - gate_sample) / max(1, d["gated"]),
                     "candidates_per_target_sampled": d["candidates"] / len(sample), "gated_per_target_sampled": d["gated"] / len(sample),
                     "counting": "op-counting oracle over 1024 random targets: the cutoff gate of PairE::operator() (image + |r|^2, 17 flop) ONCE per candidate + the functor work of every gated pair; add/sub/mul/div/sqrt = 1, cos/acos = 20, pow = 3",
                     "peak_source": "DFMA-chain microbenchmark in this process (scgpu_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                     "kernel_ms": kms,
                     "per_kernel_us": {"gate": us[0], "cheap": us[1], "patch": us[2], "combine": us[3]},
                     "per_kernel_share": {"gate": us[0] / us_tot, "cheap": us[1] / us_tot, "patch": us[2] / us_tot, "combine": us[3] / us_tot},
                     "functor_frac_over_fp64_kernels": flops_functor / (fp64_us * 1e-6) / 1e12 / peak,
                     "functor_frac_over_step": flops_functor / (kms * 1e-3) / 1e12 / peak,
                     "note": "the gate runs in FP32 (its 17 flop per candidate are algorithmic FP64 flops of the reference, executed as 3 FFMA + compare); the FP64 pipe sees the functor work only -- compare functor_frac_over_fp64_kernels with ncu sm__pipe_fp64_cycles_active of the cheap/patch kernels in profiles/",
                     "kernels": "one step = gate (k_gate_rows) + cheap terms (k_cheap_flat) + patch terms (k_patch_flat) + k_combine_flat; per_kernel_us measured live by scgpu_profile_everyone"})
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            roof["traffic"] = json.load(open(prof)).get("pipeline_dram_bytes_per_step_total")
    except Exception as ex:      # the roofline numerator needs the oracle; never let it kill the bench line
        roof["error"] = repr(ex)
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_reference(2048, 96, os.cpu_count() or 1) or cpu_port_fallback(256)      # ~0.7 s per process: ~10 core-seconds on the 16-core box
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "particles_per_replica": n, "replicas": world, "l2": "flushed between timed steps",
                      "gated_pairs_per_step": int(ngate), "candidates_per_step": int(ncand), "library": "libscgpu.so (-fmad=true)",
                      "parallelism": "replica-per-GPU"},
           "roofline": roof, "cpu_baseline": cpu,
           "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                   "mode": "scgpu_submit_everyone, %d contexts used in turn (copies of one configuration overlap the kernels of the others)" % nctx,
                   "one_at_a_time": {"value": float(ngate) * e2e_steps * world / e2e_serial_s, "ms_per_step": e2e_serial_s / e2e_steps * 1e3}},
           "gpu_launches": int(launches), "clocks": clk, "wall_s_timed_region": wall, "secondary": sweeps, "cell_build": cell_build, "full_energy": full_energy, "single_call": single_call}
    if world == 1 and not args.no_cpu:
        cs = cpu_reference_sweeps(2048, 32768, os.cpu_count() or 1)
        if cs:
            out["secondary"]["cpu_reference_sweeps"] = cs
            out["single_call"]["cpu_us_per_trial_one_core"] = cs["us_per_trial_one_core"]
    print(json.dumps(out), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ntargets, reps = 2048, 32      # ~0.2 s per process and step
    vals, secs = [], []
    r = None
    for k in range(args.warmup + args.steps):
        r = cpu_reference(ntargets, reps, cores)
        if r is None:
            r = cpu_port_fallback(256)
        if k >= args.warmup:
            vals.append(r["gated_pairs"])
            secs.append(r["seconds"])
    value = float(np.sum(vals) / np.sum(secs))
    cb = dict(r)
    cb["value"] = value
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": float(np.mean(secs) * 1e3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": WORKLOAD, "particles_per_replica": 65536, "step_sample": cb["sample"]},
           "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), file=_REAL_STDOUT, flush=True)


_REAL_STDOUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--with-membrane", action="store_true", default=True, help="also time the 265k-particle membrane full-energy pass (configs[2]); on by default")
    ap.add_argument("--no-membrane", dest="with_membrane", action="store_false", help="skip the membrane leg")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 on their own (NCCL prints its version banner
    # there when NCCL_DEBUG is set) are sent to stderr for the duration of the run, the JSON line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    main()
