"""Parallel tempering across GPUs: host-side mirror of the exchange logic that the C ABI runs on the device.

The product path is `scgpu_replica_exchange` (include/scgpu.h, sc_b200/csrc/comm.cuh; Python: `sc_b200.engine.Comm.exchange`):
allToAll() of every local replica, records packed on the device, ONE NCCL all-gather, a device kernel that takes every
pair's decision. This module holds what a host program needs around it -- the temperature ladder of Sim::readOptions
(scOOP/structures/sim.h:384-403), the analytic acceptance estimate the reference prints (scOOP/mc/inicializer.cpp:52-55) --
and a pure-Python restatement of the decision kernel (`decide_exchanges`, same Philox4x32-10 stream, same record layout) used
by the CPU tests (gloo, world size 2 and 4) and as the checker of the device kernel in the GPU tests.

Reference: MoveCreator::replicaExchangeMove (scOOP/mc/movecreator.cpp:552-795). The reference does an MPI_Alltoall of
MpiExchangeData so that every rank can find its partner, then 4 point-to-point messages per pair; configurations never
move -- temperature, pressure, pseudo-rank and statistics do. Acceptance rule as written (evaluated for the lower partner):
    change = (1/T - 1/(T+dT)) (E_here - E_recv) + (P/T - (P+dP)/(T+dT)) (V_here - V_recv) + mu terms + WL terms
    accept if change > 0 or u < exp(change)
"""
import math

import numpy as np

RX = 64          # doubles per packed record (sc_b200/csrc/comm.cuh)
RX_E, RX_V, RX_N, RX_T, RX_P, RX_PSEUDO, RX_REPLICA, RX_WL0, RX_WL1, RX_ATTEMPTED, RX_ACCEPTED, RX_PARTNER, RX_CHANGE, RX_PWL0, RX_PWL1, RX_EDRIFT = range(16)
RX_PARTNUM, RX_PAYLOAD = 16, 24
NPAYLOAD, NMOLTYPES = 40, 8


def temperature_ladder(temper, paraltemper, nprocs):
    """Sim::readOptions (sim.h:384-396): geometric in 1/T between temper and paraltemper; returns (pTemp[], dtemp)."""
    if nprocs <= 1 or paraltemper == temper:
        return [temper] * max(1, nprocs), 0.0
    dtemp = ((1.0 / temper) - (1.0 / paraltemper)) / (nprocs - 1)
    ladder = []
    for i in range(nprocs):
        ladder.append(temper / (1.0 - i * temper * dtemp))
    return ladder, dtemp


def switch_probability_estimate(n_particles, dtemp):
    """'Probability to switch replicas is roughly' (mc/inicializer.cpp:52-55)"""
    return math.exp(-0.5 * n_particles * dtemp * dtemp / (1.0 + dtemp))


# ---- Philox4x32-10, the generator of the device kernels (sc_b200/csrc/sweep.cuh) ------------------------------------------------
def philox4x32(c0, c1, c2, c3, k0, k1):
    M0, M1, MASK = 0xD2511F53, 0xCD9E8D57, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + 0x9E3779B9) & MASK, (k1 + 0xBB67AE85) & MASK
    return c0, c1, c2, c3


def u01(a, b):
    """53-bit uniform of the device code: ((a << 21) ^ (b >> 11) + 0.5) * 2^-53 * (1 - 2^-53)"""
    return (float((a << 21) ^ (b >> 11)) + 0.5) * (1.0 / 9007199254740992.0) * 0.99999999999999989


def exchange_uniform(seed, sweep, lo):
    seed, sweep = int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32(sweep & 0xFFFFFFFF, sweep >> 32, int(lo), 0x52455058, seed & 0xFFFFFFFF, seed >> 32)
    return u01(r[0], r[1])


def make_record(energy, volume, n, temper, press, pseudo_rank, replica, wl_order=(0, 0), part_num=None, payload=None):
    r = np.zeros(RX)
    r[RX_E], r[RX_V], r[RX_N], r[RX_T], r[RX_P], r[RX_PSEUDO], r[RX_REPLICA] = energy, volume, n, temper, press, pseudo_rank, replica
    r[RX_WL0], r[RX_WL1] = wl_order
    r[RX_PARTNER] = -1
    if part_num is not None:
        r[RX_PARTNUM:RX_PARTNUM + len(part_num)] = part_num
    if payload is not None:
        r[RX_PAYLOAD:RX_PAYLOAD + len(payload)] = payload
    return r


def decide_exchanges(records, sweep, nrepchange, seed, dtemp, dpress=0.0, chempot=None, wl_all=None, wl_len0=1):
    """records: [R, RX] as gathered. Returns (out_records [R, RX], decisions) where decisions is a list of
    (replica_lo, replica_hi, accepted, change); identical on every rank. Restates k_replica_decide line by line."""
    rec = np.asarray(records, dtype=np.float64).reshape(-1, RX)
    R = rec.shape[0]
    out = rec.copy()
    oddoreven = 1 if (int(sweep) % (2 * int(nrepchange))) == 0 else 0      # movecreator.cpp:616-623
    if R == 2:
        oddoreven = 1
    by_pseudo = {int(rec[g, RX_PSEUDO]): g for g in range(R)}
    decisions = []
    for hi in range(1, R):
        if hi % 2 != oddoreven:
            continue
        lo = hi - 1
        gl, gh = by_pseudo[lo], by_pseudo[hi]
        L, H = rec[gl], rec[gh]
        T, P = L[RX_T], L[RX_P]
        temp = (1 / T - 1 / (T + dtemp))
        change = temp * (L[RX_E] - H[RX_E])
        change += (P / T - (P + dpress) / (T + dtemp)) * (L[RX_V] - H[RX_V])
        if chempot is not None:
            for i, mu in enumerate(chempot):
                if mu != 0.0:
                    change += temp * mu * T * (L[RX_PARTNUM + i] - H[RX_PARTNUM + i])
        if wl_all is not None:
            wl_all = np.asarray(wl_all, dtype=np.float64).reshape(R, -1)
            localwl = int(L[RX_WL0]) + int(L[RX_WL1]) * int(wl_len0)
            receivedwl = int(H[RX_WL0]) + int(H[RX_WL1]) * int(wl_len0)
            if 0 <= localwl < wl_all.shape[1] and 0 <= receivedwl < wl_all.shape[1]:
                change += (-wl_all[gl, localwl] + wl_all[gl, receivedwl]) / T + (-wl_all[gh, receivedwl] + wl_all[gh, localwl]) / (T + dtemp)
        accepted = (change > 0) or (exchange_uniform(seed, sweep, lo) < math.exp(change))
        for g, o, src in ((gl, gh, H), (gh, gl, L)):
            out[g, RX_ATTEMPTED], out[g, RX_ACCEPTED], out[g, RX_PARTNER], out[g, RX_CHANGE] = 1.0, float(accepted), float(o), change
            out[g, RX_PWL0], out[g, RX_PWL1] = src[RX_WL0], src[RX_WL1]
        if accepted:
            for g, me, ot in ((gl, L, H), (gh, H, L)):
                out[g, RX_T], out[g, RX_P], out[g, RX_PSEUDO] = ot[RX_T], ot[RX_P], ot[RX_PSEUDO]
                out[g, RX_PAYLOAD:] = ot[RX_PAYLOAD:]
                entrophy = me[RX_P] * me[RX_V] - me[RX_N] * math.log(me[RX_V]) * me[RX_T]
                ed = me[RX_P] * (ot[RX_V] - me[RX_V]) - me[RX_N] * math.log(ot[RX_V] / me[RX_V]) * me[RX_T]
                ed += (ot[RX_P] * me[RX_V] - me[RX_N] * math.log(me[RX_V]) * ot[RX_T]) - entrophy
                out[g, RX_EDRIFT] = ed
        decisions.append((gl, gh, bool(accepted), float(change)))
    return out, decisions


# ---- WangLandau::update on merged arrays (scOOP/mc/wanglandau.h:66-123), restated for the tests of scgpu_wl_merge ----------------
WL_GERR, WL_ALPHATOL, WL_MINHIST = 0.0001, 1.0e-8, 1000


def wl_update(weights, hist, temper, alpha):
    """in place; returns (alpha, min, wmin, max, halved, converged). NB max/min is an INTEGER division in the reference."""
    mn, mx = int(hist.min()), int(hist.max())
    halved = converged = False
    wmin = 0.0
    if mn > WL_MINHIST:
        if temper * math.log(mx // mn) < WL_GERR:
            if alpha < WL_ALPHATOL:
                converged = True
            else:
                alpha /= 2
                halved = True
                wmin = float(weights[0])
                hist[:] = 0
                weights -= wmin
    return alpha, mn, wmin, mx, halved, converged


def wl_merge_host(deltas_w, deltas_h, base_w, base_h):
    """what scgpu_wl_merge computes before the update: base += sum over walkers of (mine - base)"""
    return base_w + np.sum(deltas_w, axis=0), base_h + np.sum(deltas_h, axis=0)


# ---- the parallel-tempering loop a host program runs around the C ABI (BASELINE configs[4]) ----------------------------------------
class ParallelTempering:
    """R = nranks * len(engines) replicas of one system on a temperature ladder; replica g = rank * nlocal + k starts at
    pTemp[g] (sim.h:389-396) with the random stream seed + g (sim.h:400). Every sweep is one batched checkerboard sweep per
    replica (each on its own stream); every `nrepchange` sweeps one scgpu_replica_exchange. The step sizes travel with the
    temperature in the payload (the reference swaps its Statistics block): payload[0:20] = trans_mx, [20:40] = rot_angle per type."""

    def __init__(self, comm, engines, temper, paraltemper, transmx, rotmx_deg, nrepchange=10, seed=145658, press=0.0, paralpress=0.0, grid_k=0, trial_rule=2):
        from .engine import ExchangeParams, MoveParams, ReplicaState
        self.comm, self.engines = comm, engines
        self.nlocal = len(engines)
        self.R = comm.nranks * self.nlocal
        self.ladder, self.dtemp = temperature_ladder(temper, paraltemper, self.R)
        self.dpress = (paralpress - press) / (self.R - 1) if self.R > 1 else 0.0
        self.nrepchange, self.seed = int(nrepchange), int(seed)
        self.states = (ReplicaState * self.nlocal)()
        for k in range(self.nlocal):
            g = comm.rank * self.nlocal + k
            s = self.states[k]
            s.temper, s.press, s.pseudo_rank, s.replica = self.ladder[g], press + self.dpress * g, g, g
            for t in range(20):
                s.payload[t] = 2.0 * transmx                                   # sim.h:365
                s.payload[20 + t] = rotmx_deg / 180.0 * 1.5707963267948966 * 0.5   # sim.h:360
        self.params = ExchangeParams()
        self.params.nrepchange, self.params.dtemp, self.params.dpress, self.params.seed = self.nrepchange, self.dtemp, self.dpress, self.seed
        self.mp = [MoveParams() for _ in range(self.nlocal)]
        self.grid_k = int(grid_k)
        self.trial_rule = int(trial_rule)        # scgpu_moveparams::trial_rule (2: every particle once per sweep)
        self.acc = [0] * self.nlocal
        self.rej = [0] * self.nlocal
        self.exchanges = 0
        self.exchange_us = []          # device time after the energy kernels: pack + all-gather + decision + copy back
        self._refresh()

    def _refresh(self):
        for k in range(self.nlocal):
            s, mp = self.states[k], self.mp[k]
            mp.temper, mp.n_sub, mp.grid_k, mp.trial_rule = s.temper, 1, self.grid_k, self.trial_rule
            for t in range(40):
                mp.trans_mx[t] = s.payload[min(t, 19)]
                mp.rot_angle[t] = s.payload[20 + min(t, 19)]

    def sweep(self, sweep):
        for k, e in enumerate(self.engines):
            e.sweep(self.mp[k], self.seed + self.comm.rank * self.nlocal + k, sweep, stats=False)
        if sweep % self.nrepchange == 0:
            self.comm.exchange(self.engines, self.states, self.params, sweep)
            self.exchanges += 1
            self.exchange_us.append(self.comm.last_exchange_us())
            for k in range(self.nlocal):
                if self.states[k].attempted:
                    if self.states[k].accepted:
                        self.acc[k] += 1
                    else:
                        self.rej[k] += 1
            self._refresh()
