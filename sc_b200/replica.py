"""Parallel tempering across GPUs: one replica per rank, energies exchanged with ONE collective.

Restates MoveCreator::replicaExchangeMove (scOOP/mc/movecreator.cpp:552-795) and the temperature ladder of
Sim::readOptions (scOOP/structures/sim.h:384-403). The reference does an MPI_Alltoall of MpiExchangeData so that every
rank can find its partner, then 4 point-to-point messages per pair; configurations never move -- temperature,
pressure, pseudo-rank and statistics do. Here every rank all-gathers a packed 8-double record
    {E, V, N, T, P, pseudoRank, 0, 0}
(NCCL on GPUs, gloo in the CPU tests), the full energy E coming straight from the device-side reduction
(scgpu_replica_record) without visiting the host, and then every rank evaluates the SAME deterministic odd/even
pairing and the same Metropolis test with a counter-based random number keyed on (seed, sweep, lower pseudo-rank):
no second round trip is needed. Acceptance rule, as written in the reference (evaluated for the lower-T partner):
    change = (1/T - 1/(T+dT)) (E_here - E_recv) + (P/T - (P+dP)/(T+dT)) (V_here - V_recv);  accept if change > 0 or u < exp(change)
"""
import math

import numpy as np

REC = 8


def temperature_ladder(temper, paraltemper, nprocs):
    """Sim::readOptions (sim.h:384-396): geometric in 1/T between temper and paraltemper; returns (pTemp[], dtemp)."""
    if nprocs <= 1 or paraltemper == temper:
        return [temper] * max(1, nprocs), 0.0
    dtemp = ((1.0 / temper) - (1.0 / paraltemper)) / (nprocs - 1)
    ladder = []
    for i in range(nprocs):
        ladder.append(temper / (1.0 - i * temper * dtemp))
    return ladder, dtemp


class ReplicaState:
    """what a rank swaps on an accepted exchange (sim->temper, sim->press, sim->pseudoRank + acceptance counters)"""

    def __init__(self, rank, temper, press=0.0):
        self.rank = rank
        self.temper = temper
        self.press = press
        self.pseudo_rank = rank
        self.acc = 0
        self.rej = 0


def _uniform(seed, sweep, lo):
    g = np.random.Generator(np.random.Philox(key=int(seed) & 0xFFFFFFFFFFFFFFFF, counter=[int(sweep), int(lo), 0, 0]))
    return float(g.random())


def decide_exchanges(records, sweep, nrepchange, seed, dtemp, dpress=0.0):
    """records: array [world, 8] = {E, V, N, T, P, pseudoRank, ...} as gathered. Returns a list of
    (rank_lo, rank_hi, accepted, change) for every attempted pair; identical on every rank."""
    records = np.asarray(records, dtype=np.float64).reshape(-1, REC)
    world = records.shape[0]
    oddoreven = 1 if (sweep % (2 * nrepchange)) == 0 else 0      # movecreator.cpp:616-623
    if world == 2:
        oddoreven = 1
    by_pseudo = {int(round(records[r, 5])): r for r in range(world)}
    out = []
    for hi in range(1, world):
        if hi % 2 != oddoreven:
            continue
        lo = hi - 1
        r_lo, r_hi = by_pseudo[lo], by_pseudo[hi]
        E_l, V_l, _, T_l, P_l = records[r_lo, 0:5]
        E_h, V_h = records[r_hi, 0], records[r_hi, 1]
        temp = (1.0 / T_l - 1.0 / (T_l + dtemp))
        change = temp * (E_l - E_h)
        change += (P_l / T_l - (P_l + dpress) / (T_l + dtemp)) * (V_l - V_h)
        accepted = (change > 0) or (_uniform(seed, sweep, lo) < math.exp(change))
        out.append((r_lo, r_hi, bool(accepted), float(change)))
    return out


def apply_exchanges(state, records, decisions):
    """swap (T, P, pseudoRank) of this rank with its partner's if its pair was accepted"""
    records = np.asarray(records, dtype=np.float64).reshape(-1, REC)
    for (r_lo, r_hi, acc, _) in decisions:
        if state.rank not in (r_lo, r_hi):
            continue
        other = r_hi if state.rank == r_lo else r_lo
        if acc:
            state.temper = float(records[other, 3])
            state.press = float(records[other, 4])
            state.pseudo_rank = int(round(records[other, 5]))
            state.acc += 1
        else:
            state.rej += 1
    return state


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def record_tensor(engine):
    """device tensor view of the engine's replica record {E, V, N, 0...}; E = full energy computed on this call"""
    import torch
    ptr = engine.replica_record_ptr()
    return torch.as_tensor(_DevArray(ptr, REC), device="cuda")


def exchange(engine, state, sweep, nrepchange, seed, dtemp, dpress=0.0, dist=None):
    """one replica-exchange attempt on GPUs: full-energy kernel -> all-gather -> identical decisions everywhere"""
    import torch
    rec = record_tensor(engine).clone()
    rec[3:6] = torch.tensor([state.temper, state.press, float(state.pseudo_rank)], dtype=torch.float64, device="cuda")
    world = dist.get_world_size()
    gathered = torch.zeros(world * REC, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(gathered, rec)
    records = gathered.cpu().numpy().reshape(world, REC)
    decisions = decide_exchanges(records, sweep, nrepchange, seed, dtemp, dpress)
    apply_exchanges(state, records, decisions)
    return decisions, records
