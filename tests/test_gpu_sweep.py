"""GPU: batched checkerboard displacement/rotation sweeps (row A14). They cannot reproduce the reference's sequential
trajectory (different proposal order and RNG), so they are validated by invariants and STATISTICALLY:
  * energy bookkeeping: E_total(after) - E_total(before) == sum of accepted dE (the reference's own drift check,
    scOOP/mc/updater.cpp:345-352)
  * reproducibility: same (seed, configuration) -> bit-identical trajectory
  * rigid-body invariants of every particle survive thousands of rotations
  * <E> over a production window agrees with sequential sweeps of the REFERENCE program (tests/golden/sweep_psc1280.json,
    made by tests/golden/make_sweep_golden.py from the unmodified reference sources) within the statistical error
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import Engine, synth
from sc_b200.engine import MoveParams
from sc_b200.host import HostSystem

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIH = 1.57079632679489661923132169163975


def move_params(temper, transmx, rotmx, n_sub=1):
    mp = MoveParams()
    mp.temper = temper
    for k in range(40):
        mp.trans_mx[k] = 2.0 * transmx                    # sim.h:365  transmx *= 2
        mp.rot_angle[k] = rotmx / 180.0 * PIH * 0.5       # sim.h:360
    mp.n_sub = n_sub
    return mp


@pytest.mark.parametrize("kind", ["psc_lattice", "mix", "chains"])
def test_energy_bookkeeping_and_invariants(kind):
    top, cfg = synth.small_case(kind)
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    e0 = eng.all_to_all()
    mp = move_params(0.5 if kind != "psc_lattice" else 0.25, 0.05, 8.0)
    tot_de, acc, rej, cell_rej = 0.0, 0, 0, 0
    nsw = 12
    for sw in range(nsw):
        st = eng.sweep(mp, 777, sw)
        tot_de += st.energy_delta
        acc += st.trans_acc + st.rot_acc
        rej += st.trans_rej + st.rot_rej
        cell_rej += st.cell_rej
    # one sweep = N trials in expectation: every non-empty cell does N / (non-empty cells) of them, stochastically rounded
    assert abs((acc + rej) - nsw * hs.n) <= 6.0 * np.sqrt(nsw * 0.25 * hs.n) + 1
    e1 = eng.all_to_all()
    assert acc > 0 and rej > 0
    assert cell_rej < 0.2 * (acc + rej)
    scale = max(abs(e0), abs(e1), 1.0)
    assert abs((e1 - e0) - tot_de) <= 1e-9 * scale, (e0, e1, tot_de)
    state = eng.download_particles()
    # particles stay rigid: |dir| = 1, patchdir perpendicular to dir (rods only)
    for i in range(hs.n):
        g = int(hs.ia[hs.type[i], hs.type[i], 0])
        if g < 30:
            d, p = state[i, 3:6], state[i, 6:9]
            assert abs(np.dot(d, d) - 1.0) < 1e-9
            if g >= 12:
                assert abs(np.dot(p, p) - 1.0) < 1e-9 and abs(np.dot(d, p)) < 1e-9
    # the oracle agrees with the engine on the evolved configuration
    s = O.system_from_text(top, cfg)
    s.state[:] = state
    for t in range(0, hs.n, max(1, hs.n // 16)):
        a, b = eng.one_to_all(t), s.one_to_all(t)
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b)) + 1e-10
    eng.close()


def test_trial_rule_per_particle_gives_exactly_n_trials_and_keeps_the_books():
    """scgpu_moveparams::trial_rule = 1, 2: a cell performs (its population) trials per sweep -- N in total, exactly (no rounding), on
    every grid fineness; with rule 2 every particle exactly once; energies stay exact"""
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    for gk, rule in ((0, 1), (1, 1), (2, 1), (0, 2), (1, 2), (2, 2)):
        eng = Engine(0, "fast").load(hs)
        mp = move_params(0.25, 0.05, 8.0)
        mp.trial_rule, mp.grid_k = rule, gk
        e0 = eng.all_to_all()
        de = 0.0
        for sw in range(6):
            st = eng.sweep(mp, 31, sw)
            assert st.trans_acc + st.trans_rej + st.rot_acc + st.rot_rej == hs.n
            de += st.energy_delta
        e1 = eng.all_to_all()
        assert abs((e1 - e0) - de) <= 1e-9 * max(abs(e0), abs(e1), 1.0)
        eng.close()
    mp.trial_rule = 3
    eng = Engine(0, "fast").load(hs)
    with pytest.raises(Exception):
        eng.sweep(mp, 31, 0)
    eng.close()
    hs.close()


@pytest.mark.parametrize("kind", ["psc_lattice", "mix"])
def test_cell_walk_kernel_on_unbonded_systems(kind, monkeypatch):
    """systems without bonds take the round kernel by default; the one-warp-per-cell walk (k_sweep_cells, what bonded systems use)
    must keep the same exact books on them (SCGPU_SWEEP_KERNEL=cells)"""
    monkeypatch.setenv("SCGPU_SWEEP_KERNEL", "cells")
    top, cfg = synth.small_case(kind)
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mp = move_params(0.5 if kind != "psc_lattice" else 0.25, 0.05, 8.0)
    e0 = eng.all_to_all()
    de, acc = 0.0, 0
    for sw in range(8):
        st = eng.sweep(mp, 778, sw)
        de += st.energy_delta
        acc += st.trans_acc + st.rot_acc
    e1 = eng.all_to_all()
    assert acc > 0
    assert abs((e1 - e0) - de) <= 1e-9 * max(abs(e0), abs(e1), 1.0)
    eng.close()
    hs.close()


@pytest.mark.parametrize("kind", ["psc_lattice", "mix", "psc_8192"])
def test_phased_sweep_walks_the_same_chain_as_the_round_kernel(kind, monkeypatch):
    """trial_rule 2 on the coarse grid runs as four dense launches per colour pass (sweep_phased.cuh: k_sweep_propose, the energy
    pipeline's k_cheap_flat / k_patch_flat, k_sweep_resolve); SCGPU_SWEEP_KERNEL=rounds forces the one-kernel form. Both draw the same
    permutation and the same proposals and take the decisions of the same sequential walk, so the configurations after a few sweeps
    are THE SAME (the sums are taken in different orders: a decision could only flip on a 1e-16 tie)."""
    if kind == "psc_8192":         # 8 x 8 x 6 cells: no axis can wrap, the gate of k_sweep_propose works in length units without the fold
        top, cfg, _ = synth.psc_bulk(32, 32, 8, seed=7, tilt=0.1)
    else:                          # 4 x 4 x 4 cells: the fold path
        top, cfg = synth.small_case(kind)
    hs = HostSystem(top, cfg)
    finals, books, launches = [], [], []
    for kern in ("phased", "rounds"):
        monkeypatch.setenv("SCGPU_SWEEP_KERNEL", kern)
        eng = Engine(0, "fast").load(hs)
        mp = move_params(0.5 if kind != "psc_lattice" else 0.25, 0.05, 8.0)
        mp.trial_rule, mp.grid_k = 2, 1
        e0 = eng.all_to_all()
        de = 0.0
        nacc = 0
        for sw in range(4):
            st = eng.sweep(mp, 4711, sw)
            assert st.trans_acc + st.trans_rej + st.rot_acc + st.rot_rej == hs.n
            de += st.energy_delta
            nacc += st.trans_acc + st.rot_acc
        l0 = eng.launches()
        de += eng.sweep(mp, 4711, 4).energy_delta
        launches.append(eng.launches() - l0)
        e1 = eng.all_to_all()
        assert abs((e1 - e0) - de) <= 1e-9 * max(abs(e0), abs(e1), 1.0), kern
        finals.append(eng.download_particles())
        books.append((nacc, e1))
        eng.close()
    assert launches[0] - launches[1] == 3 * 8, launches           # four launches per colour pass against one: the two forms really ran
    assert books[0][0] == books[1][0] and books[0][0] > 0
    assert np.array_equal(finals[0], finals[1])
    hs.close()


def test_reproducible_trajectory():
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    finals = []
    for _ in range(2):
        eng = Engine(0, "fast").load(hs)
        mp = move_params(0.25, 0.08, 12.0)
        for sw in range(6):
            eng.sweep(mp, 4242, sw)
        finals.append(eng.download_particles())
        eng.close()
    assert np.array_equal(finals[0], finals[1])
    eng = Engine(0, "fast").load(hs)
    mp = move_params(0.25, 0.08, 12.0)
    for sw in range(6):
        eng.sweep(mp, 4243, sw)
    assert not np.array_equal(finals[0], eng.download_particles())       # another seed, another trajectory
    eng.close()


def _block_stderr(x, nblocks=8):
    x = np.asarray(x)
    m = len(x) // nblocks
    means = [x[k * m:(k + 1) * m].mean() for k in range(nblocks)]
    return float(np.std(means, ddof=1) / math.sqrt(nblocks))


@pytest.mark.parametrize("rule", [0, 2])
def test_average_energy_matches_reference_sequential_sweeps(rule):
    """rule 0: the same number of trials in every cell, with replacement; rule 2: every particle once per sweep in random order
    (scgpu_moveparams::trial_rule) -- both through the round kernel (k_sweep_rounds), this system has no bonds"""
    gold = json.load(open(os.path.join(G, "sweep_psc1280.json")))
    P = gold["params"]
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    skip = P["nsweeps"] // 3
    ref_means, ref_errs = [], []
    for run in gold["runs"]:
        sw, e = np.array(run["sweep"]), np.array(run["energy"])
        w = e[sw > skip]
        ref_means.append(w.mean())
        ref_errs.append(_block_stderr(w))
    ref_mean = float(np.mean(ref_means))
    ref_err = max(float(np.std(ref_means, ddof=1) / math.sqrt(len(ref_means))), float(np.mean(ref_errs)) / math.sqrt(len(ref_means)))
    def one_seed(seed):          # the seeds run side by side (one context and stream each): a 1 280-particle sweep leaves the GPU almost idle
        eng = Engine(0, "fast").load(hs)
        mp = move_params(P["temper"], P["transmx"], P["rotmx"])
        mp.trial_rule = rule
        e = eng.all_to_all()
        series = []
        a_t = a_r = n_t = n_r = 0
        for sw in range(1, P["nsweeps"] + 1):
            st = eng.sweep(mp, seed, sw)
            e += st.energy_delta
            a_t += st.trans_acc; n_t += st.trans_acc + st.trans_rej
            a_r += st.rot_acc; n_r += st.rot_acc + st.rot_rej
            if sw > skip and sw % P["report"] == 0:
                series.append(e)
        e_check = eng.all_to_all()
        assert abs(e_check - e) <= 1e-7 * abs(e_check)                   # running sum of dE == recomputed total
        eng.close()
        return np.mean(series), _block_stderr(series), a_t, n_t, a_r, n_r

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(one_seed, (101, 202, 303, 404)))
    gpu_means, gpu_errs = [r[0] for r in res], [r[1] for r in res]
    acc_t, tot_t, acc_r, tot_r = (sum(r[k] for r in res) for k in (2, 3, 4, 5))
    gpu_mean = float(np.mean(gpu_means))
    gpu_err = max(float(np.std(gpu_means, ddof=1) / math.sqrt(len(gpu_means))), float(np.mean(gpu_errs)) / math.sqrt(len(gpu_means)))
    sigma = math.sqrt(ref_err ** 2 + gpu_err ** 2)
    print("reference <E> = %.3f +- %.3f ; checkerboard <E> = %.3f +- %.3f ; acceptance trans %.3f rot %.3f"
          % (ref_mean, ref_err, gpu_mean, gpu_err, acc_t / tot_t, acc_r / tot_r))
    assert abs(gpu_mean - ref_mean) <= 4.0 * sigma + 2e-3 * abs(ref_mean), (gpu_mean, ref_mean, sigma)


def test_dense_membrane_sweeps_fall_back_to_global_scan():
    """CPSC + lipid membrane (bonded 3-bead chains, neighbourhoods far larger than the staged tile): sweeps still keep exact books"""
    import gzip
    inp = json.loads(gzip.open(os.path.join(G, "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(4, 4, inp["top.init"], inp["config.init"])
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    e0 = eng.all_to_all()
    mp = move_params(1.0, 0.05, 10.0)
    tot = 0.0
    acc = 0
    for sw in range(4):
        st = eng.sweep(mp, 99, sw)
        tot += st.energy_delta
        acc += st.trans_acc + st.rot_acc
    e1 = eng.all_to_all()
    assert acc > 0
    assert abs((e1 - e0) - tot) <= 1e-9 * max(abs(e0), abs(e1)), (e0, e1, tot)
    eng.close()


# ------------------------------------------------------------------------------------------------
# chain moves on the device (scOOP/mc/movecreator.cpp:304-328, 1075-1256, 1308-1392)
# ------------------------------------------------------------------------------------------------
def chain_moves(chainprob, chainmmx, chainrmx):
    from sc_b200.engine import ChainMoves
    cm = ChainMoves()
    cm.chainprob = chainprob
    for k in range(32):
        cm.chainm_mx[k] = 2.0 * chainmmx                   # sim.h:366  chainmmx *= 2
        cm.chainr_angle[k] = chainrmx / 180.0 * PIH        # sim.h:362
    return cm


def _molecules(hs):
    """(first index, size) of every molecule, from the molecule-type table the engine itself uses"""
    out, i = [], 0
    while i < hs.n:
        m = int(hs.mol[hs.moltype[i], 12])
        out.append((i, m))
        i += m
    return out


def _intramolecular(state, box, mols):
    d = []
    for first, m in mols:
        for a in range(first, first + m):
            for b in range(a + 1, first + m):
                r = (state[a, 0:3] - state[b, 0:3]) * box      # chains live in unwrapped coordinates: no image
                d.append(math.sqrt(float(np.dot(r, r))))
            if m > 1:
                d.append(float(np.dot(state[first, 3:6], state[first + 1, 3:6])))
    return np.array(d)


@pytest.mark.parametrize("kind", ["chain_fluid", "chains"])
def test_chain_moves_only_are_rigid_and_keep_exact_books(kind):
    """chainprob = 1: every trial is a chain move. Molecules move rigidly (all intramolecular distances and relative
    orientations unchanged), the summed accepted dE equals the change of the recomputed total energy, and the oracle agrees on
    the evolved configuration."""
    top, cfg = synth.small_case(kind)
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mols = _molecules(hs)
    assert any(m > 1 for _, m in mols)
    s0 = eng.download_particles()
    g0 = _intramolecular(s0, hs.box, mols)
    e0 = eng.all_to_all()
    mp = move_params(1.0, 0.1, 10.0)
    cm = chain_moves(1.0, 0.2, 15.0)
    tot, acc_m, acc_r, tried, cell_rej = 0.0, 0, 0, 0, 0
    for sw in range(10):
        st, cst = eng.sweep_chains(mp, cm, 4242, sw)
        assert st.trans_acc + st.trans_rej + st.rot_acc + st.rot_rej == 0          # no single-particle trials at chainprob = 1
        tot += cst.energy_delta
        acc_m += cst.chainm_acc; acc_r += cst.chainr_acc
        tried += cst.chainm_acc + cst.chainm_rej + cst.chainr_acc + cst.chainr_rej
        cell_rej += cst.cell_rej
    assert acc_m > 0 and acc_r > 0 and tried > 0
    assert cell_rej < tried
    e1 = eng.all_to_all()
    assert abs((e1 - e0) - tot) <= 1e-9 * max(abs(e0), abs(e1), 1.0), (e0, e1, tot)
    s1 = eng.download_particles()
    assert np.max(np.abs(s1[:, 0:3] - s0[:, 0:3])) > 1e-3                             # something moved
    g1 = _intramolecular(s1, hs.box, mols)
    assert np.max(np.abs(g1 - g0)) < 1e-9, float(np.max(np.abs(g1 - g0)))
    for i in range(hs.n):
        g = int(hs.ia[hs.type[i], hs.type[i], 0])
        if g < 30:
            d = s1[i, 3:6]
            assert abs(np.dot(d, d) - 1.0) < 1e-9
    s = O.system_from_text(top, cfg)
    s.state[:] = s1
    for t in range(0, hs.n, max(1, hs.n // 16)):
        a, b = eng.one_to_all(t), s.one_to_all(t)
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b)) + 1e-10
    eng.close()


def test_mixed_sweeps_with_chain_moves_keep_exact_books_and_are_reproducible():
    top, cfg = synth.small_case("chain_fluid")
    hs = HostSystem(top, cfg)
    finals = []
    for rep in range(2):
        eng = Engine(0, "fast").load(hs)
        e0 = eng.all_to_all()
        mp = move_params(1.0, 0.12, 15.0)
        cm = chain_moves(0.3, 0.15, 12.0)
        tot, n_part, n_chain, n_noop = 0.0, 0, 0, 0
        for sw in range(8):
            st, cst = eng.sweep_chains(mp, cm, 777, sw)
            tot += st.energy_delta + cst.energy_delta
            n_part += st.trans_acc + st.trans_rej + st.rot_acc + st.rot_rej
            n_chain += cst.chainm_acc + cst.chainm_rej + cst.chainr_acc + cst.chainr_rej
            n_noop += cst.noop
        e1 = eng.all_to_all()
        assert abs((e1 - e0) - tot) <= 1e-9 * max(abs(e0), abs(e1), 1.0), (e0, e1, tot)
        # the trial mix of Updater::simulate (updater.cpp:206-230): a share chainprob of the N trials of a sweep are chain moves;
        # picks that land on a one-particle molecule (the rods: 160 of 1600 particles) are no-ops and not counted
        assert abs(n_part - 0.7 * 8 * hs.n) <= 6.0 * math.sqrt(8 * hs.n) + 8
        assert abs(n_chain + n_noop - 0.3 * 8 * hs.n) <= 6.0 * math.sqrt(8 * hs.n) + 8
        assert abs(n_noop - 0.3 * 8 * hs.n * 0.1) <= 6.0 * math.sqrt(8 * hs.n) + 8
        finals.append(eng.download_particles())
        eng.close()
    assert np.array_equal(finals[0], finals[1])


def test_average_energy_with_chain_moves_matches_reference_sequential_sweeps():
    gold = json.load(open(os.path.join(G, "sweep_chain_fluid.json")))
    P = gold["params"]
    top, cfg = synth.small_case("chain_fluid")
    hs = HostSystem(top, cfg)
    skip = P["nsweeps"] // 3
    ref_means, ref_errs = [], []
    for run in gold["runs"]:
        sw, e = np.array(run["sweep"]), np.array(run["energy"])
        w = e[sw > skip]
        ref_means.append(w.mean())
        ref_errs.append(_block_stderr(w))
    ref_mean = float(np.mean(ref_means))
    ref_err = max(float(np.std(ref_means, ddof=1) / math.sqrt(len(ref_means))), float(np.mean(ref_errs)) / math.sqrt(len(ref_means)))
    def one_seed(seed):
        eng = Engine(0, "fast").load(hs)
        mp = move_params(P["temper"], P["transmx"], P["rotmx"])
        cm = chain_moves(P["chainprob"], P["chainmmx"], P["chainrmx"])
        e = eng.all_to_all()
        series = []
        for sw in range(1, P["nsweeps"] + 1):
            st, cst = eng.sweep_chains(mp, cm, seed, sw)
            e += st.energy_delta + cst.energy_delta
            if sw > skip and sw % P["report"] == 0:
                series.append(e)
        e_check = eng.all_to_all()
        assert abs(e_check - e) <= 1e-7 * max(abs(e_check), 1.0)
        eng.close()
        return np.mean(series), _block_stderr(series)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(one_seed, (101, 202, 303, 404)))
    gpu_means, gpu_errs = [r[0] for r in res], [r[1] for r in res]
    gpu_mean = float(np.mean(gpu_means))
    gpu_err = max(float(np.std(gpu_means, ddof=1) / math.sqrt(len(gpu_means))), float(np.mean(gpu_errs)) / math.sqrt(len(gpu_means)))
    sigma = math.sqrt(ref_err ** 2 + gpu_err ** 2)
    print("reference <E> = %.3f +- %.3f ; checkerboard with chain moves <E> = %.3f +- %.3f" % (ref_mean, ref_err, gpu_mean, gpu_err))
    assert abs(gpu_mean - ref_mean) <= 4.0 * sigma + 2e-3 * abs(ref_mean), (gpu_mean, ref_mean, sigma)


def test_membrane_lipids_chain_sweeps():
    """the move mix the reference uses on the lipid membrane (configs[2]): single-bead moves plus rigid moves of whole SPN-SPA-SPA
    lipids, in neighbourhoods far denser than the staged tile of the single-particle kernel"""
    import gzip
    inp = json.loads(gzip.open(os.path.join(G, "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(4, 4, inp["top.init"], inp["config.init"])
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    mols = _molecules(hs)
    s0 = eng.download_particles()
    g0 = _intramolecular(s0, hs.box, mols)
    e0 = eng.all_to_all()
    mp = move_params(1.0, 0.05, 10.0)
    mp_none = move_params(1.0, 0.05, 10.0)
    cm = chain_moves(1.0, 0.1, 10.0)
    tot = 0.0
    for sw in range(5):
        st, cst = eng.sweep_chains(mp_none, cm, 99, sw)
        tried = cst.chainm_acc + cst.chainm_rej + cst.chainr_acc + cst.chainr_rej
        # every pick lands on a lipid bead except those on the 16 CPSC rods (no-ops)
        assert abs(tried + cst.noop - n) <= 6.0 * math.sqrt(n) + 20, (sw, tried, cst.noop, n)
        assert cst.noop < 0.02 * n
        assert cst.chainm_acc > 0 and cst.chainr_acc > 0
        tot += cst.energy_delta
    s1 = eng.download_particles()
    assert np.all(np.isfinite(s1))
    assert np.max(np.abs(_intramolecular(s1, hs.box, mols) - g0)) < 1e-9
    e1 = eng.all_to_all()
    assert abs((e1 - e0) - tot) <= 1e-9 * max(abs(e0), abs(e1)), (e0, e1, tot)
    cm = chain_moves(0.5, 0.1, 10.0)
    for sw in range(5, 8):
        st, cst = eng.sweep_chains(mp, cm, 99, sw)
        tot += st.energy_delta + cst.energy_delta
        tried = cst.chainm_acc + cst.chainm_rej + cst.chainr_acc + cst.chainr_rej
        assert abs(tried + cst.noop - 0.5 * n) <= 6.0 * math.sqrt(n) + 20, (sw, tried, cst.noop, n)
    e2 = eng.all_to_all()
    assert abs((e2 - e0) - tot) <= 1e-9 * max(abs(e0), abs(e2)), (e0, e2, tot)
    eng.close()


# ------------------------------------------------------------------------------------------------
# volume moves between batched sweeps (MoveCreator::pressureMove, scOOP/mc/movecreator.cpp:330-550)
# ------------------------------------------------------------------------------------------------
def test_pressure_move_mechanics():
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    eng = Engine(0, "fast").load(hs)
    box = np.array(hs.box, dtype=np.float64)
    n_acc = 0
    for step in range(40):
        ptype = step % 4
        e_before = eng.all_to_all()
        st = eng.pressure_move(ptype, 0.08, 0.5, 0.3, 2024, step)
        assert st.energy_old == e_before
        nb = np.array(list(st.box))
        if st.accepted:
            n_acc += 1
            assert not np.array_equal(nb, box)
            assert eng.all_to_all() == st.energy_new                       # the accepted box is the one the device now holds
            if ptype == 0:
                assert np.sum(nb != box) == 1                               # one edge
            elif ptype == 1:
                assert np.allclose(nb - box, (nb - box)[0], rtol=0, atol=1e-12)
            elif ptype == 2:
                assert nb[2] == box[2] and abs((nb[0] - box[0]) - (nb[1] - box[1])) < 1e-12
            else:
                assert abs(nb.prod() - box.prod()) < 1e-9 * box.prod()
            dv = nb.prod() - box.prod()
            de = st.energy_new - st.energy_old
            if ptype == 1:
                assert abs(st.enthalpy_delta - (de + 0.08 * dv - hs.n * 0.5 * math.log(nb.prod() / box.prod()))) < 1e-9 * max(1.0, abs(de))
            box = nb
        else:
            assert np.array_equal(nb, box) and st.enthalpy_delta == 0.0
            assert eng.all_to_all() == e_before                            # the old box is back, to the last bit
    assert 0 < n_acc < 40
    # a pure function of (seed, step, configuration)
    a = eng.pressure_move(1, 0.08, 0.5, 0.3, 7, 7)
    if a.accepted:
        eng.set_box(box)
    b = eng.pressure_move(1, 0.08, 0.5, 0.3, 7, 7)
    assert (a.accepted, a.energy_new) == (b.accepted, b.energy_new)
    eng.close()


@pytest.mark.parametrize("rule", [0, 2])
def test_npt_average_volume_matches_reference_sequential_sweeps(rule):
    """checkerboard sweeps + one volume move per sweep (the reference draws a volume move with probability shave/N per step, i.e.
    `shave` per sweep on average) against the reference's own NPT run: <V> and <E> over the second half"""
    gold = json.load(open(os.path.join(G, "sweep_npt_psc1280.json")))
    P = gold["params"]
    top, cfg = synth.small_case("psc_lattice")
    hs = HostSystem(top, cfg)
    skip = P["nsweeps"] // 2
    ref_v, ref_e = [], []
    for run in gold["runs"]:
        sw, v = np.array(run["sweep"]), np.array(run["volume"])
        ref_v.append(v[sw > skip].mean())
        se, e = np.array(run["energy_sweep"]), np.array(run["energy"])
        ref_e.append(e[se > skip].mean())
    def one_seed(seed):
        eng = Engine(0, "fast").load(hs)
        mp = move_params(P["temper"], P["transmx"], P["rotmx"])
        mp.trial_rule = rule                # 0: round kernel, 2: phased sweep (the box changes between the sweeps)
        vs, es = [], []
        n_acc = 0
        e = eng.all_to_all()
        for sw in range(1, P["nsweeps"] + 1):
            st = eng.sweep(mp, seed, sw)
            e += st.energy_delta
            pm = eng.pressure_move(P["ptype"], P["press"], P["temper"], 2.0 * P["edge_mx"], seed, sw)
            if pm.accepted:
                n_acc += 1
                e = pm.energy_new
            if sw > skip and sw % P["report"] == 0:
                vs.append(float(np.prod(list(pm.box))))
                es.append(e)
        assert abs(eng.all_to_all() - e) <= 1e-7 * max(1.0, abs(e))
        assert 0.2 < n_acc / P["nsweeps"] < 0.9
        eng.close()
        return np.mean(vs), np.mean(es)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one_seed, (101, 202, 303, 404, 505, 606, 707, 808)))
    gpu_v, gpu_e = [r[0] for r in res], [r[1] for r in res]
    rv, gv = float(np.mean(ref_v)), float(np.mean(gpu_v))
    sv = math.sqrt(np.var(ref_v, ddof=1) / len(ref_v) + np.var(gpu_v, ddof=1) / len(gpu_v))
    re_, ge = float(np.mean(ref_e)), float(np.mean(gpu_e))
    se_ = math.sqrt(np.var(ref_e, ddof=1) / len(ref_e) + np.var(gpu_e, ddof=1) / len(gpu_e))
    print("reference <V> = %.1f, checkerboard NPT <V> = %.1f (sigma %.1f); <E> %.2f vs %.2f (sigma %.2f)" % (rv, gv, sv, re_, ge, se_))
    assert abs(gv - rv) <= 4.0 * sv + 5e-3 * rv, (gv, rv, sv)
    assert abs(ge - re_) <= 4.0 * se_ + 2e-2 * abs(re_), (ge, re_, se_)


# ---- the reference's own "comprehensive" inputs: Tests/System_averages_tests/CPSC/various_temp --------------
def _cpsc_golden():
    import gzip
    return json.loads(gzip.open(os.path.join(G, "sweep_cpsc_temps.json.gz"), "rt").read())


def _window_stats(sweeps, energy, lo, hi):
    e = np.array([v for s_, v in zip(sweeps, energy) if lo < s_ <= hi])
    return float(e.mean()), _block_stderr(e, 4) if len(e) >= 8 else float(e.std() / math.sqrt(max(1, len(e))))


@pytest.mark.parametrize("system", ["cpsc100", "cpsc800"])
def test_cpsc_system_averages_three_temperatures(system):
    """<E> AND the per-move acceptance ratios at T = 0.1, 0.16, 0.22 against sequential runs of the unmodified reference program
    (8 seeds each; tests/golden/make_sweep_cpsc_golden.py). cpsc100 is the input as shipped (one cell: the proposal / acceptance logic
    alone); cpsc800 is the same configuration tiled 2 x 2 x 2 (4 cells per axis: the checkerboard decomposition, where a bias would
    show first at the lowest temperature). Both sides start from the same configuration and are compared over the same window of
    sweeps, (W/3, W], W = 5 000, and from the configuration the reference itself starts from at that temperature (every T_x directory
    ships its own equilibrated config.init)."""
    from concurrent.futures import ThreadPoolExecutor
    gold = _cpsc_golden()
    W = 5000
    top = gold["top"][system]
    cfgs = {round(r["temper"], 3): r["config"] for r in gold["runs"] if r["system"] == system and r["config"]}
    temps = [0.1, 0.16, 0.22]
    seeds = [5, 6, 7, 8, 9, 10, 11, 12]       # as many as the reference: a run is a single warp (cpsc100) or a few (cpsc800), they overlap on the GPU
    chunk = 50

    def run(job):
        temper, seed = job
        hs = HostSystem(top, cfgs[round(temper, 3)])
        eng = Engine(0, "fast").load(hs)
        mp = move_params(temper, 0.03, 15.0, n_sub=chunk)
        mp.trial_rule = 1 if system == "cpsc100" else 2          # per-particle rates with replacement / every particle once per sweep
        sw, en = [], []
        ta = tr = ra = rr = 0
        for k in range(W // chunk):
            st = eng.sweep(mp, 1000 + seed, k)
            ta += st.trans_acc; tr += st.trans_rej; ra += st.rot_acc; rr += st.rot_rej
            sw.append((k + 1) * chunk)
            en.append(eng.all_to_all())
        eng.close()
        hs.close()
        return temper, sw, en, ta / max(1, ta + tr), ra / max(1, ra + rr)

    with ThreadPoolExecutor(len(temps) * len(seeds)) as ex:
        results = list(ex.map(run, [(t, s_) for t in temps for s_ in seeds]))
    for temper in temps:
        ref = [r for r in gold["runs"] if r["system"] == system and abs(r["temper"] - temper) < 1e-9]
        assert len(ref) == 8
        rm = [_window_stats(r["sweep"], r["energy"], W // 3, W)[0] for r in ref]
        ref_mean, ref_se = float(np.mean(rm)), float(np.std(rm, ddof=1) / math.sqrt(len(rm)))
        mine = [r for r in results if r[0] == temper]
        gm = [_window_stats(r[1], r[2], W // 3, W)[0] for r in mine]
        g_mean = float(np.mean(gm))
        # seed-to-seed scatter is the honest error bar here (near T = 0.16 the window means of the REFERENCE's own seeds differ by 4 %:
        # the correlation time is a good part of the window); both sides are means over 8 seeds
        sd = max(float(np.std(rm, ddof=1)), float(np.std(gm, ddof=1)))
        tol = 4.0 * sd * math.sqrt(1.0 / len(rm) + 1.0 / len(gm)) + 0.005 * abs(ref_mean)
        assert abs(g_mean - ref_mean) <= tol, (system, temper, g_mean, ref_mean, tol)
        # acceptance ratios: the reference prints whole per cent of the full run (Statistics::print)
        ref_t = float(np.mean([r["acceptance"]["trans_acc_pct"] for r in ref])) / 100.0
        ref_r = float(np.mean([r["acceptance"]["rot_acc_pct"] for r in ref])) / 100.0
        g_t, g_r = float(np.mean([r[3] for r in mine])), float(np.mean([r[4] for r in mine]))
        slack = 0.03 if system == "cpsc100" else 0.04          # + moves rejected for leaving their cell on the tiled system
        assert abs(g_t - ref_t) <= slack and abs(g_r - ref_r) <= slack, (system, temper, g_t, ref_t, g_r, ref_r)
