"""CPU: replica-exchange plumbing (sc_b200/replica.py) with torch.distributed gloo, world_size 2 and 4:
every rank reaches identical decisions from one all-gather, swaps are consistent, the ladder matches sim.h."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sc_b200 import replica


def test_ladder_matches_reference_formula():
    lad, dtemp = replica.temperature_ladder(0.1, 0.13, 8)
    assert lad[0] == 0.1 and abs(lad[-1] - 0.13) < 1e-12
    inv = [1.0 / t for t in lad]
    assert np.allclose(np.diff(inv), -dtemp)


def test_decisions_follow_the_reference_rule():
    # two replicas: lower T holds the HIGHER energy -> change > 0 -> always accepted
    lad, dtemp = replica.temperature_ladder(0.1, 0.12, 2)
    rec = np.zeros((2, 8))
    rec[0] = [-100.0, 1000.0, 64, lad[0], 0.0, 0, 0, 0]
    rec[1] = [-120.0, 1000.0, 64, lad[1], 0.0, 1, 0, 0]
    d = replica.decide_exchanges(rec, sweep=10, nrepchange=10, seed=1, dtemp=dtemp)
    assert len(d) == 1 and d[0][2] is True
    assert abs(d[0][3] - (1 / lad[0] - 1 / (lad[0] + dtemp)) * 20.0) < 1e-12
    # strongly unfavourable: never accepted
    rec[0, 0], rec[1, 0] = -200.0, -100.0
    assert replica.decide_exchanges(rec, 10, 10, 1, dtemp)[0][2] is False
    # odd/even alternation with 4 replicas (movecreator.cpp:616-623)
    rec4 = np.zeros((4, 8))
    for r in range(4):
        rec4[r] = [-100.0, 1.0, 1, 0.1 + 0.01 * r, 0.0, r, 0, 0]
    pairs_a = [(a, b) for (a, b, _, _) in replica.decide_exchanges(rec4, 20, 10, 1, 0.01)]
    pairs_b = [(a, b) for (a, b, _, _) in replica.decide_exchanges(rec4, 30, 10, 1, 0.01)]
    assert pairs_a == [(0, 1), (2, 3)] and pairs_b == [(1, 2)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lad, dtemp = replica.temperature_ladder(0.1, 0.13, world)
    st = replica.ReplicaState(rank, lad[rank])
    rng = np.random.default_rng(100 + rank)
    log = []
    for sweep in range(10, 210, 10):
        e = -1000.0 - 0.5 * rng.random() + 0.1 * st.pseudo_rank       # fake full energies
        rec = torch.tensor([e, 5.0e5, 65536.0, st.temper, st.press, float(st.pseudo_rank), 0, 0], dtype=torch.float64)
        gathered = [torch.zeros(8, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, rec)
        records = torch.stack(gathered).numpy()
        dec = replica.decide_exchanges(records, sweep, 10, 4242, dtemp)
        replica.apply_exchanges(st, records, dec)
        log.append((sweep, tuple((a, b, c) for (a, b, c, _) in dec), st.temper, st.pseudo_rank))
    q.put((rank, log, st.acc, st.rej))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_ranks_agree(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, log, acc, rej = q.get(timeout=120)
        res[rank] = (log, acc, rej)
    for p in procs:
        p.join(timeout=60)
    lad, _ = replica.temperature_ladder(0.1, 0.13, world)
    nsteps = len(res[0][0])
    for k in range(nsteps):
        decs = {res[r][0][k][1] for r in range(world)}
        assert len(decs) == 1                                  # identical decisions on every rank
        temps = sorted(res[r][0][k][2] for r in range(world))
        assert np.allclose(temps, sorted(lad))                 # temperatures are permuted, never lost
        assert sorted(res[r][0][k][3] for r in range(world)) == list(range(world))
    assert sum(res[r][1] for r in range(world)) > 0            # some exchanges were accepted
    assert sum(res[r][1] for r in range(world)) % 2 == 0       # always in pairs
