"""GPU: the multi-GPU layer of the C ABI (scgpu_replica_exchange, scgpu_wl_merge) on ONE device, several replicas per
process -- the 1-GPU corner of BASELINE configs[4] (8 replicas on 1/2/4/8 GPUs). The device decision kernel is compared
with its Python restatement (sc_b200/replica.py, which follows MoveCreator::replicaExchangeMove,
scOOP/mc/movecreator.cpp:552-795, line by line); the energies that enter the rule are the device's allToAll() sums.
The N > 1 path (NCCL all-gather between processes) needs >= 2 GPUs: tests/test_gpu_replica.py::test_two_ranks runs
only where two devices exist; scripts/parallel_tempering.py and bench.py --gpus N exercise it on the multi-GPU box.
"""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from sc_b200 import Engine, replica, synth
from sc_b200.engine import Comm, ExchangeParams, ReplicaState
from sc_b200.host import HostSystem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _engines(nrep, jitter=2e-3):
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    engs = []
    for r in range(nrep):
        st = hs.state.copy()
        st[:, 0:3] += np.random.default_rng(50 + r).normal(scale=jitter * (1 + r), size=(hs.n, 3))
        e = Engine(0, "fast").load(hs)
        e.set_particles(st, hs.type, hs.moltype)
        engs.append(e)
    return hs, engs


def _states(nrep, ladder, press=0.0):
    arr = (ReplicaState * nrep)()
    for r in range(nrep):
        arr[r].temper, arr[r].press, arr[r].pseudo_rank = ladder[r], press + 0.1 * r, r
        for k in range(40):
            arr[r].payload[k] = 100.0 * r + k
        arr[r].part_num[0] = 1280.0
    return arr


@pytest.mark.parametrize("nrep", [2, 4, 8])
def test_exchange_matches_python_restatement(nrep):
    hs, engs = _engines(nrep)
    ladder, dtemp = replica.temperature_ladder(0.25, 0.40, nrep)
    comm = Comm(0, 1, 0)
    states = _states(nrep, ladder, press=1.0)
    p = ExchangeParams()
    p.nrepchange, p.dtemp, p.dpress, p.seed = 10, dtemp, 0.1, 99
    energies = [e.all_to_all() for e in engs]
    vol = float(np.prod(hs.box))
    n_acc = 0
    for sweep in range(10, 130, 10):
        # what the device should do, from the same inputs
        recs = np.stack([replica.make_record(energies[r], vol, hs.n, states[r].temper, states[r].press, states[r].pseudo_rank, r,
                                             part_num=[1280.0], payload=list(states[r].payload)) for r in range(nrep)])
        want, dec = replica.decide_exchanges(recs, sweep, 10, 99, dtemp, 0.1)
        comm.exchange(engs, states, p, sweep)
        for r in range(nrep):
            s = states[r]
            assert s.replica == r
            assert s.energy == energies[r] and abs(s.volume - vol) <= 1e-12 * vol          # E is the device's allToAll(), bit for bit
            assert s.temper == want[r, replica.RX_T] and s.press == want[r, replica.RX_P] and s.pseudo_rank == int(want[r, replica.RX_PSEUDO])
            assert s.attempted == int(want[r, replica.RX_ATTEMPTED]) and s.accepted == int(want[r, replica.RX_ACCEPTED])
            assert s.partner == int(want[r, replica.RX_PARTNER])
            assert abs(s.change - want[r, replica.RX_CHANGE]) <= 1e-12 * max(1.0, abs(s.change))
            assert abs(s.edrift - want[r, replica.RX_EDRIFT]) <= 1e-9 * max(1.0, abs(s.edrift))
            assert list(s.payload) == list(want[r, replica.RX_PAYLOAD:])
            n_acc += s.accepted
        # temperatures are permuted, never lost
        assert sorted(states[r].temper for r in range(nrep)) == sorted(ladder)
        assert sorted(states[r].pseudo_rank for r in range(nrep)) == list(range(nrep))
    assert n_acc > 0 and n_acc % 2 == 0
    assert 0.0 < comm.last_exchange_us() < 5000.0
    comm.close()
    for e in engs:
        e.close()


def test_exchange_wang_landau_term():
    hs, engs = _engines(2)
    comm = Comm(0, 1, 0)
    ladder, dtemp = replica.temperature_ladder(0.25, 0.30, 2)
    states = _states(2, ladder)
    states[0].wl_order[0], states[1].wl_order[0] = 3, 5
    wl = [np.linspace(0.0, 4.0, 12), np.linspace(1.0, -2.0, 12)]
    p = ExchangeParams()
    p.nrepchange, p.dtemp, p.seed, p.wl_len, p.wl_len0 = 10, dtemp, 5, 12, 12
    energies = [e.all_to_all() for e in engs]
    recs = np.stack([replica.make_record(energies[r], float(np.prod(hs.box)), hs.n, ladder[r], 0.1 * r, r, r, wl_order=(states[r].wl_order[0], 0),
                                         part_num=[1280.0], payload=list(states[r].payload)) for r in range(2)])
    want, dec = replica.decide_exchanges(recs, 10, 10, 5, dtemp, 0.0, wl_all=np.stack(wl), wl_len0=12)
    comm.exchange(engs, states, p, 10, wl_weights=wl)
    assert abs(states[0].change - dec[0][3]) <= 1e-12 * max(1.0, abs(dec[0][3]))
    assert states[0].partner_wl_order[0] == 5 and states[1].partner_wl_order[0] == 3
    assert states[0].accepted == int(dec[0][2])
    comm.close()
    for e in engs:
        e.close()


def test_wl_merge_single_walker():
    # one walker: the merge is the identity, then WangLandau::update as written (wanglandau.h:66-123)
    comm = Comm(0, 1, 0)
    rng = np.random.default_rng(3)
    L = 937
    w = rng.normal(size=L); h = rng.integers(1500, 2900, size=L).astype(np.int64)
    wb, hb = w - rng.random(L) * 0.01, h - 3
    w0, h0 = w.copy(), h.copy()
    wr, hr = w.copy(), h.copy()
    alpha, mn, wmin, mx, halved, conv = replica.wl_update(wr, hr, 1.0, 0.02)
    st = comm.wl_merge(w, h, wb, hb, 1.0, 0.02)
    assert halved and st.halved == 1 and st.alpha == alpha and st.min == mn and st.max == mx and st.wmin == wmin and st.converged == 0
    assert np.array_equal(w, wr) and np.array_equal(h, hr) and np.array_equal(wb, w) and np.array_equal(hb, h)
    # not flat: arrays unchanged, alpha kept
    w, h = w0.copy(), h0.copy(); h[7] = 9000
    wb, hb = w.copy(), h.copy()
    st = comm.wl_merge(w, h, wb, hb, 1.0, 0.02)
    assert st.halved == 0 and st.alpha == 0.02 and np.array_equal(w, w0) and h[7] == 9000 and st.max == 9000
    comm.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device); covered by bench.py --gpus N on the multi-GPU box")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "parallel_tempering.py"), "--small", "--replicas", "4",
                          "--sweeps", "40", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
