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


def _recs(rows):
    return np.stack([replica.make_record(*r) for r in rows])


def test_decisions_follow_the_reference_rule():
    # two replicas: lower T holds the HIGHER energy -> change > 0 -> always accepted
    lad, dtemp = replica.temperature_ladder(0.1, 0.12, 2)
    rec = _recs([(-100.0, 1000.0, 64, lad[0], 0.0, 0, 0), (-120.0, 1000.0, 64, lad[1], 0.0, 1, 1)])
    out, d = replica.decide_exchanges(rec, sweep=10, nrepchange=10, seed=1, dtemp=dtemp)
    assert len(d) == 1 and d[0][2] is True
    assert abs(d[0][3] - (1 / lad[0] - 1 / (lad[0] + dtemp)) * 20.0) < 1e-12
    # temperatures, pseudo-ranks (and payloads) changed hands; energies and volumes stayed
    assert out[0, replica.RX_T] == lad[1] and out[1, replica.RX_T] == lad[0]
    assert out[0, replica.RX_PSEUDO] == 1 and out[1, replica.RX_PSEUDO] == 0
    assert out[0, replica.RX_E] == -100.0 and out[0, replica.RX_PARTNER] == 1 and out[1, replica.RX_PARTNER] == 0
    # strongly unfavourable: never accepted
    rec[0, 0], rec[1, 0] = -200.0, -100.0
    out, d = replica.decide_exchanges(rec, 10, 10, 1, dtemp)
    assert d[0][2] is False and out[0, replica.RX_T] == lad[0] and out[0, replica.RX_ATTEMPTED] == 1
    # odd/even alternation with 4 replicas (movecreator.cpp:616-623)
    rec4 = _recs([(-100.0, 1.0, 1, 0.1 + 0.01 * r, 0.0, r, r) for r in range(4)])
    pairs_a = [(a, b) for (a, b, _, _) in replica.decide_exchanges(rec4, 20, 10, 1, 0.01)[1]]
    pairs_b = [(a, b) for (a, b, _, _) in replica.decide_exchanges(rec4, 30, 10, 1, 0.01)[1]]
    assert pairs_a == [(0, 1), (2, 3)] and pairs_b == [(1, 2)]


def test_pressure_mu_and_wl_terms():
    # isobaric term (movecreator.cpp:729-730), grand-canonical term (:733-736), Wang-Landau term (:738-742)
    T, dT, dP = 0.5, 0.05, 0.2
    a = replica.make_record(-10.0, 900.0, 50, T, 1.0, 0, 0, wl_order=(2, 0), part_num=[50, 3])
    b = replica.make_record(-12.0, 950.0, 50, T + 0.07, 1.2, 1, 1, wl_order=(4, 0), part_num=[50, 7])
    wl = np.arange(16, dtype=np.float64).reshape(2, 8) * 0.25
    out, d = replica.decide_exchanges(np.stack([a, b]), 20, 10, 7, dT, dP, chempot=[0.0, 1.5], wl_all=wl, wl_len0=8)
    temp = 1 / T - 1 / (T + dT)
    want = temp * 2.0 + (1.0 / T - 1.2 / (T + dT)) * (900.0 - 950.0) + temp * 1.5 * T * (3 - 7)
    want += (-wl[0, 2] + wl[0, 4]) / T + (-wl[1, 4] + wl[1, 2]) / (T + dT)
    assert abs(d[0][3] - want) < 1e-12
    assert out[0, replica.RX_PWL0] == 4 and out[1, replica.RX_PWL0] == 2


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10 (counter = key = 0 / all ones / pi digits)
    assert replica.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert replica.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert replica.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_wl_update_as_written():
    # WangLandau::update (wanglandau.h:66-123): flat when T * log(max / min [integer division]) < WL_GERR and min > WL_MINHIST
    w = np.array([3.0, 5.0, 4.0]); h = np.array([1500, 2900, 2000], dtype=np.int64)
    alpha, mn, wmin, mx, halved, conv = replica.wl_update(w, h, 1.0, 0.01)
    assert halved and alpha == 0.005 and mn == 1500 and mx == 2900 and wmin == 3.0       # 2900 // 1500 == 1 -> log 1 = 0
    assert np.array_equal(h, [0, 0, 0]) and np.array_equal(w, [0.0, 2.0, 1.0])
    w = np.array([3.0, 5.0]); h = np.array([1500, 3000], dtype=np.int64)
    assert replica.wl_update(w, h, 1.0, 0.01)[4] is False                                 # 3000 // 1500 == 2
    h = np.array([900, 901], dtype=np.int64)
    assert replica.wl_update(w, h, 1.0, 0.01)[4] is False                                 # min <= WL_MINHIST
    h = np.array([1500, 1501], dtype=np.int64)
    assert replica.wl_update(w, h, 1.0, 1e-9)[5] is True                                  # converged


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
    temper, pseudo, payload = lad[rank], rank, np.full(40, float(rank))
    acc = rej = 0
    rng = np.random.default_rng(100 + rank)
    log = []
    for sweep in range(10, 210, 10):
        e = -1000.0 - 0.5 * rng.random() + 0.1 * pseudo       # fake full energies
        rec = torch.from_numpy(replica.make_record(e, 5.0e5, 65536.0, temper, 0.0, pseudo, rank, payload=payload))
        gathered = [torch.zeros(replica.RX, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, rec)
        records = torch.stack(gathered).numpy()
        out, dec = replica.decide_exchanges(records, sweep, 10, 4242, dtemp)
        mine = out[rank]
        if mine[replica.RX_ATTEMPTED] > 0:
            if mine[replica.RX_ACCEPTED] > 0:
                acc += 1
            else:
                rej += 1
        temper, pseudo, payload = float(mine[replica.RX_T]), int(mine[replica.RX_PSEUDO]), mine[replica.RX_PAYLOAD:].copy()
        assert payload[0] == pseudo                          # the payload travels with the temperature
        log.append((sweep, tuple((a, b, c) for (a, b, c, _) in dec), temper, pseudo))
    q.put((rank, log, acc, rej))
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
