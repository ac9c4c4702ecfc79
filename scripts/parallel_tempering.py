#!/usr/bin/env python
"""Parallel tempering (BASELINE.json configs[4]): R replicas of one system on N GPUs (R/N per GPU, each on
its own stream), batched checkerboard sweeps, a replica-exchange attempt every `nrepchange` sweeps through the C ABI's
scgpu_replica_exchange: allToAll() on the device, records packed on the device, ONE ncclAllGather, a device decision kernel.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/parallel_tempering.py [--small] [--replicas 8] [--sweeps 60]
    python scripts/parallel_tempering.py --small         (one GPU, all replicas on it, no NCCL)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sc_b200 import Engine, synth, replica          # noqa: E402
from sc_b200.engine import Comm                      # noqa: E402
from sc_b200.host import HostSystem                  # noqa: E402


def make_comm(local, world, rank):
    """the unique id of rank 0 travels through torch.distributed (the reference's main() would MPI_Bcast it)"""
    if world == 1:
        return Comm(local, 1, 0)
    import torch
    import torch.distributed as dist
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, device=torch.device("cuda", local))
    return Comm(local, world, rank, box[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="1280 rods instead of 65 536")
    ap.add_argument("--replicas", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=60)
    ap.add_argument("--nrepchange", type=int, default=10)
    ap.add_argument("--temper", type=float, default=0.1)
    ap.add_argument("--paraltemper", type=float, default=0.13)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.small:
        top, cfg = synth.small_case("psc_lattice")
    else:
        top, cfg, _ = synth.psc_bulk()
    hs = HostSystem(top, cfg)
    nlocal = max(1, args.replicas // world)
    engines = [Engine(local, "fast").load(hs) for _ in range(nlocal)]
    comm = make_comm(local, world, rank)
    pt = replica.ParallelTempering(comm, engines, args.temper, args.paraltemper, 0.0212, 7.5, args.nrepchange)
    for e in engines:
        e.sync()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for sweep in range(1, args.sweeps + 1):
        pt.sweep(sweep)
    for e in engines:
        e.sync()
    if dist:
        dist.barrier()
    dt = time.perf_counter() - t0
    out = {"rank": rank, "T": [pt.states[k].temper for k in range(nlocal)], "pseudo": [pt.states[k].pseudo_rank for k in range(nlocal)],
           "acc": pt.acc, "rej": pt.rej, "exchange_us": float(np.mean(pt.exchange_us)) if pt.exchange_us else None}
    gathered = [out]
    if dist:
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
    if rank == 0:
        temps = sorted(t for g in gathered for t in g["T"])
        pseudo = sorted(p for g in gathered for p in g["pseudo"])
        acc, rej = sum(sum(g["acc"]) for g in gathered), sum(sum(g["rej"]) for g in gathered)
        print(json.dumps({"replicas": pt.R, "gpus": world, "particles": hs.n, "sweeps": args.sweeps, "seconds": dt,
                          "aggregate_sweeps_per_s": pt.R * args.sweeps / dt, "exchange_attempts": pt.exchanges,
                          "exchange_us_after_energy": gathered[0]["exchange_us"], "pair_acceptance": acc / max(1, acc + rej),
                          "analytic_estimate": replica.switch_probability_estimate(hs.n, pt.dtemp), "ladder": pt.ladder, "ranks": gathered}))
        if args.check:
            assert np.allclose(temps, sorted(pt.ladder)), "temperatures must be a permutation of the ladder"
            assert pseudo == list(range(pt.R)), "pseudo-ranks must be a permutation"
            assert acc % 2 == 0 and (acc + rej) > 0
    comm.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
