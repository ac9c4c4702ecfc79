#!/usr/bin/env python
"""Parallel tempering, one replica per GPU (BASELINE.json configs[4]): batched checkerboard sweeps on every
rank, a replica-exchange attempt every `nrepchange` sweeps through ONE NCCL all-gather of the packed record
{E, V, N, T, P, pseudoRank} (the full energy comes from the device-side reduction, scgpu_replica_record).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/parallel_tempering.py [--small] [--sweeps 60]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth, replica          # noqa: E402
from sc_b200.engine import MoveParams                # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="1280 rods instead of 65 536")
    ap.add_argument("--sweeps", type=int, default=60)
    ap.add_argument("--nrepchange", type=int, default=10)
    ap.add_argument("--temper", type=float, default=0.1)
    ap.add_argument("--paraltemper", type=float, default=0.13)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.small:
        top, cfg = synth.small_case("psc_lattice")
    else:
        top, cfg, _ = synth.psc_bulk()
    hs = HostSystem(top, cfg)
    eng = Engine(local, "fast").load(hs)
    ladder, dtemp = replica.temperature_ladder(args.temper, args.paraltemper, world)
    st = replica.ReplicaState(rank, ladder[rank])
    mp = MoveParams()
    for k in range(40):
        mp.trans_mx[k] = 2.0 * 0.0212
        mp.rot_angle[k] = 7.5 / 180.0 * 1.5707963267948966 * 0.5
    mp.n_sub = 1
    seed = 145658 + rank                                  # sim.h:400  seed += mpirank
    log = []
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    t_exch = 0.0
    for sweep in range(1, args.sweeps + 1):
        mp.temper = st.temper
        eng.sweep(mp, seed, sweep, stats=False)
        if sweep % args.nrepchange == 0:
            eng.sync()
            te = time.perf_counter()
            dec, rec = replica.exchange(eng, st, sweep, args.nrepchange, 4242, dtemp, 0.0, dist)
            t_exch += time.perf_counter() - te
            log.append([sweep, [(a, b, int(c)) for (a, b, c, _) in dec], [float(x) for x in rec[:, 0]]])
    eng.sync()
    dist.barrier()
    dt = time.perf_counter() - t0
    out = {"rank": rank, "T_final": st.temper, "pseudo_rank": st.pseudo_rank, "acc": st.acc, "rej": st.rej}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps({"replicas": world, "particles": hs.n, "sweeps": args.sweeps, "seconds": dt, "aggregate_sweeps_per_s": world * args.sweeps / dt,
                          "exchange_attempts": len(log), "exchange_ms_each": t_exch / max(1, len(log)) * 1e3, "ranks": gathered,
                          "last_energies": log[-1][2] if log else None, "ladder": ladder}))
        temps = sorted(g["T_final"] for g in gathered)
        assert np.allclose(temps, sorted(ladder)), "temperatures must be a permutation of the ladder"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
