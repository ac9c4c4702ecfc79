"""GPU: parity of the CUDA path (through the C ABI, include/scgpu.h) against the CPU oracle and the
committed reference goldens. Bars: energies <= 1e-10 relative (absolute floor
1e-12 for sums that cancel to ~0), cell assignment / sort order / overlap flags bit-exact.
Both library variants are checked: `strict` (-fmad=false) and `fast` (-fmad=true, the benchmarked one).
"""
import glob
import gzip
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import topo as otopo
from sc_b200 import Engine
from sc_b200 import synth

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(G, "*.ref.gz")))
GRIDS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(G, "grid_*.npz")))
VARIANTS = ["strict", "fast"]
RTOL = 1e-10
ATOL = 1e-12


def close(a, b, scale=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= RTOL * np.maximum(np.abs(a), np.abs(b)) + ATOL * scale)


def worst(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300))) if a.size else 0.0


@pytest.fixture(scope="module", params=VARIANTS)
def engine(request):
    e = Engine(0, request.param)
    yield e
    e.close()


def eps_scale(s):
    return max(1.0, float(np.max(np.abs(s.ia[:, :, 4]))))


@pytest.mark.parametrize("name", DUMPS)
def test_golden_configs(engine, name):
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    s = r.system
    engine.load(s)
    sc = eps_scale(s)
    # cell assignment and stable sort: bit-exact with the oracle's definition C1
    ncell, cell_of, order, start = s.cells()
    g_cell, g_nc = engine.cell_assignment()
    assert np.array_equal(g_nc, ncell) and np.array_equal(g_cell, cell_of)
    g_order, g_start = engine.cell_order(int(ncell.prod()))
    assert np.array_equal(g_order, order) and np.array_equal(g_start, start)
    targets = range(s.n) if s.n <= 100 else range(0, s.n, max(1, s.n // 25))
    for t in targets:
        e, ep = engine.one_to_all(t, pairs=True)
        ref = np.array([r.pairs.get((t, j), 0.0) for j in range(s.n)])
        assert close(ep, ref, sc), (name, t, worst(ep, ref))
        assert close(e, r.one[t], sc), (name, t, e, r.one[t])
    # batch == singles
    tl = list(targets)
    eb = engine.one_to_all_batch(tl)
    assert close(eb, [r.one[t] for t in tl], sc)
    ev = engine.one_to_all_everyone()
    assert close(ev[tl], [r.one[t] for t in tl], sc)
    if s.n <= 500:
        tot, rows = engine.all_to_all(rows=True)
        otot, orows = s.all_to_all(rows=True)
        assert close(rows, orows, sc), worst(rows, orows)
        assert close(tot, r.total, sc * 10), (tot, r.total)
    for (first, m, e) in r.mol2o[:16]:
        assert close(engine.mol_to_others(first, m), e, sc)
    # overlap: as written in the reference (variant 0), bit-exact flags
    if s.n <= 500:
        assert engine.overlap_all(0) == (1 if r.overlaps else 0)
        for t in list(targets)[:10]:
            assert engine.overlap_one(t, None, 0) == s.overlap_one(t, None, 0)
            assert engine.overlap_one(t, None, 1) == s.overlap_one(t, None, 1)


@pytest.mark.parametrize("name", GRIDS)
def test_pose_grids(engine, name):
    """Interactions_tests-style degenerate pose grids: E(0,j) and E(j,0) against the reference's values"""
    z = np.load(os.path.join(G, name + ".npz"))
    cfg = gzip.open(os.path.join(G, "grid_config_%s.txt.gz" % str(z["kind"]))).read().decode()
    s = O.system_from_text(str(z["top"]), cfg)
    engine.load(s)
    sc = eps_scale(s)
    e0, ep = engine.one_to_all(0, pairs=True)
    ref = z["e0j"]
    # overlapping poses produce energies up to 1e20: relative bar only
    assert close(ep, ref, sc), (name, worst(ep, ref), int(np.argmax(np.abs(ep - ref))))
    # E(j,0): every pose as the first particle -- all_to_all rows are sum_{k<j} E(j,k); pick E(j,0) through pairs of j
    js = np.nonzero(z["ej0"])[0][:40]
    for j in js:
        _, pj = engine.one_to_all(int(j), pairs=True)
        assert close(pj[0], z["ej0"][j], sc), (name, j, pj[0], z["ej0"][j])
    for j in range(1, s.n, max(1, s.n // 60)):
        assert engine.overlap_one(0, None, 0) in (0, 1)
    ov = [engine.overlap_one(int(j), None, 0) for j in np.nonzero(z["ovj0"])[0][:5]]
    assert all(v == 1 for v in ov)


def _trial_moves(s, rng, k):
    """random displacement / rotation proposals in the reference's style (movecreator.cpp:947-1028)"""
    out = []
    for _ in range(k):
        t = int(rng.integers(s.n))
        st = s.state[t].copy()
        g = int(s.ia[s.type[t], s.type[t], 0])
        if rng.random() < 0.5 or g >= otopo.SP:
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            st[0:3] += 0.3 * v / s.box
        else:
            ax = rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            st = O.psc_rotate(st, g, 0.2 * rng.random(), ax, int(rng.integers(2)))
        out.append((t, st))
    return out


@pytest.mark.parametrize("name", ["test_01_normal_PSC_end", "test_08_normal_TCHCPSC_end", "test_14_normal_SPA_PSC_CPSC_end",
                                  "test_20_chain_bond12_end", "test_21_chain_bondd2_end", "test_mempore_init"])
def test_trial_states(engine, name):
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    s = r.system
    engine.load(s)
    sc = eps_scale(s)
    rng = np.random.default_rng(7)
    moves = _trial_moves(s, rng, 24)
    for (t, st) in moves:
        assert close(engine.one_to_all(t, st), s.one_to_all(t, st), sc)
        assert engine.overlap_one(t, st, 0) == s.overlap_one(t, st, 0)
    tl = [m[0] for m in moves]
    sts = np.array([m[1] for m in moves])
    eb = engine.one_to_all_batch(tl, sts)
    assert close(eb, [s.one_to_all(t, st) for (t, st) in moves], sc)
    # update(): commit a move, energies must follow
    t, st = moves[0]
    engine.update_particle(t, st)
    s.state[t] = st
    for q in range(0, s.n, max(1, s.n // 12)):
        assert close(engine.one_to_all(q), s.one_to_all(q), sc)
    if s.n <= 500:
        assert close(engine.all_to_all(), s.all_to_all(), sc * 10)


@pytest.mark.parametrize("kind", ["psc_lattice", "psc_gas", "mix", "chains"])
def test_multicell_systems(engine, kind):
    """systems with a real 3-D cell grid: pair-set equivalence of the cell path with the all-pairs oracle"""
    top, cfg = synth.small_case(kind)
    s = O.system_from_text(top, cfg)
    engine.load(s)
    sc = eps_scale(s)
    ncell, cell_of, order, start = s.cells()
    assert ncell.prod() > 27
    g_cell, g_nc = engine.cell_assignment()
    assert np.array_equal(g_nc, ncell) and np.array_equal(g_cell, cell_of)
    g_order, g_start = engine.cell_order(int(ncell.prod()))
    assert np.array_equal(g_order, order) and np.array_equal(g_start, start)
    ev, ncand, ngate = engine.one_to_all_everyone(count=True)
    oc, og = 0, 0
    for t in range(0, s.n, max(1, s.n // 200)):
        assert close(ev[t], s.one_to_all(t), sc), (kind, t, ev[t], s.one_to_all(t))
    for t in range(s.n):
        _, c, g = s.one_to_all_cells(t, (ncell, cell_of, order, start))
        oc += c
        og += g
    assert (ncand, ngate) == (oc, og)      # integer work counters: bit-exact
    tot, rows = engine.all_to_all(rows=True)
    assert close(np.sum(rows), tot, sc * 10)
    assert close(2 * tot, np.sum(ev), sc * 100)    # every pair is in exactly two one-to-all sums (energies are symmetric)
    if s.n <= 3000:
        assert close(tot, s.all_to_all(), sc * 10)
    assert engine.overlap_all(0) == s.overlap_all(0)
    assert engine.overlap_all(1) == s.overlap_all(1)


def _pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64).pin_memory()


def test_row_unit_gate_matches_cell_gate_and_oracle(engine):
    """rods-only system on a grid without wrap: the every-particle passes run k_gate_rows (thread per target); with work counters
    requested they run k_gate_cells. Same pair set -> energies agree to rounding (the summation order differs), both agree
    with the oracle; the asynchronous whole-configuration call returns the same bits as the synchronous one."""
    top, cfg = synth.small_case("rods_wide")
    s = O.system_from_text(top, cfg)
    engine.load(s)
    sc = eps_scale(s)
    ncell = s.cells()[0]
    assert ncell[0] >= 7 and ncell[1] >= 5 and ncell[2] >= 5
    ev_cells, ncand, ngate = engine.one_to_all_everyone(count=True)
    ev_rows = engine.one_to_all_everyone()
    assert close(ev_rows, ev_cells, sc * 10), worst(ev_rows, ev_cells)
    assert np.array_equal(ev_rows, engine.one_to_all_everyone())          # run-to-run identical
    for t in range(0, s.n, max(1, s.n // 150)):
        assert close(ev_rows[t], s.one_to_all(t), sc), (t, ev_rows[t], s.one_to_all(t))
    tot, rows = engine.all_to_all(rows=True)
    assert close(np.sum(rows), tot, sc * 10)
    assert close(2 * tot, np.sum(ev_rows), sc * 100)
    assert close(tot, s.all_to_all(), sc * 10)
    # a shrunk box (NPT): fewer cells per axis, the wrap variant of the cell gate takes over
    state9, out = _pinned((s.n, 9)), _pinned((s.n,))
    state9.numpy()[:] = s.state[:, :9]
    engine.set_particles_compact(s.state[:, :9], s.type, s.moltype)
    ref = engine.one_to_all_everyone()
    for _ in range(3):
        engine.submit_everyone(state9.numpy(), out.numpy())
        engine.sync()
        assert np.array_equal(out.numpy(), ref)
    from sc_b200 import ScgpuError
    with pytest.raises(ScgpuError):
        engine.submit_everyone(np.zeros((s.n, 9)), out.numpy())          # pageable memory is refused


def test_row_unit_gate_falls_back_when_the_layout_does_not_fit(engine):
    """four times denser: a unit's neighbourhood exceeds the staged tile of k_gate_rows, the launch raises the fallback flag and
    is repeated with k_gate_cells -- silently for the caller, with the same results as the counting pass and the oracle"""
    top, cfg = synth.small_case("rods_dense")
    s = O.system_from_text(top, cfg)
    engine.load(s)
    sc = eps_scale(s)
    l0 = engine.launches()
    ev = engine.one_to_all_everyone()
    l1 = engine.launches()
    ev2 = engine.one_to_all_everyone()
    l2 = engine.launches()
    assert l1 - l0 > l2 - l1 > 0                       # the first call launched twice (rows, then cells); later calls go straight to cells
    assert np.array_equal(ev, ev2)
    ev_cells, _, _ = engine.one_to_all_everyone(count=True)
    assert close(ev, ev_cells, sc * 10)
    for t in range(0, s.n, max(1, s.n // 100)):
        assert close(ev[t], s.one_to_all(t), sc), (t, ev[t], s.one_to_all(t))


def test_row_unit_gate_splits_off_targets_with_very_many_partners(engine):
    """spheres with ~20 partners and a few long rods with ~300: the rods' targets go to their own gate. With sub-cells (the host sees
    that the spheres' reach fits a third of a cell) the split is known before the first launch; without them (SCGPU_NO_SUBCELLS) the
    general row-unit gate notices that the rods overflow its per-target buffers and the launch is repeated with the rod type split
    off. Either way: five launches once settled, same results as the cell gate alone (counting pass) and the oracle, bit-reproducible."""
    top, cfg = synth.small_case("rods_in_spheres")
    s = O.system_from_text(top, cfg)
    engine.load(s)
    sc = eps_scale(s)
    ncell = s.cells()[0]
    assert min(ncell) >= 5
    ev_cells, _, _ = engine.one_to_all_everyone(count=True)
    l0 = engine.launches()
    ev = engine.one_to_all_everyone()
    l1 = engine.launches()
    ev2 = engine.one_to_all_everyone()
    l2 = engine.launches()
    assert l2 - l1 == 5 and l1 - l0 >= 5              # settled: light-type gate + gate of the rods + cheap + patch + combine
    assert np.array_equal(ev, ev2)
    assert close(ev, ev_cells, sc * 10), worst(ev, ev_cells)
    rods = np.where(s.type == s.type[-1])[0]
    for t in list(rods[:12]) + list(range(0, s.n, max(1, s.n // 60))):
        assert close(ev[t], s.one_to_all(t), sc), (t, ev[t], s.one_to_all(t))
    tot, rows = engine.all_to_all(rows=True)
    assert close(np.sum(rows), tot, sc * 10)
    assert close(2 * tot, np.sum(ev), sc * 100)


def test_determinism_and_box_change(engine):
    top, cfg = synth.small_case("psc_gas")
    s = O.system_from_text(top, cfg)
    engine.load(s)
    a = engine.one_to_all_everyone()
    b = engine.one_to_all_everyone()
    assert np.array_equal(a, b)                      # fixed-order reductions: run-to-run identical
    t1 = engine.all_to_all()
    assert t1 == engine.all_to_all()
    # NPT: the caller changes the box (positions are box-fractional), calculator must follow (paire.h:1205,1211)
    for f in (0.97, 1.05, 0.6):
        s.box = s.box * f
        engine.set_box(s.box)
        for t in range(0, s.n, max(1, s.n // 20)):
            assert close(engine.one_to_all(t), s.one_to_all(t), eps_scale(s))
        nc, _, _, _ = s.cells()
        assert np.array_equal(engine.cell_assignment()[1], nc)


def test_errors_are_loud(engine):
    from sc_b200 import ScgpuError
    top, cfg = synth.small_case("psc_gas")
    s = O.system_from_text(top, cfg)
    engine.load(s)
    with pytest.raises(ScgpuError):
        engine.one_to_all(s.n + 5)
    with pytest.raises(ScgpuError):
        engine.set_box([0.0, 1.0, 1.0])
    with pytest.raises(ScgpuError):
        engine.overlap_all(7)


@pytest.mark.parametrize("name", ["test_01_normal_PSC_init", "test_08_normal_TCHCPSC_init", "test_06_normal_TCPSC_init", "test_14_normal_SPA_PSC_CPSC_init",
                                  "extra_mix_init", "extra_chains_init"])
def test_compact_upload_derives_the_same_particles(engine, name):
    """scgpu_set_particles_compact: 9 doubles per particle, Particle::init on the device == the host/reference derivation"""
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    s = r.system
    engine.set_topology(s.ia, s.mol, s.sqmaxcut, s.maxcut)
    engine.set_box(s.box)
    engine.set_particles_compact(s.state[:, :9], s.type, s.moltype)
    got = engine.download_particles()
    from test_oracle_golden import used_fields
    for i in range(s.n):
        f = used_fields(int(s.ia[s.type[i], s.type[i], 0]))
        assert np.max(np.abs(got[i, f] - s.state[i, f])) < 1e-14, (i, got[i, f] - s.state[i, f])   # re-normalising a unit vector moves last bits
    sc = eps_scale(s)
    tl = list(range(0, s.n, max(1, s.n // 20)))
    assert close(engine.one_to_all_batch(tl), [r.one[t] for t in tl], sc)
    # types NULL = "unchanged since the previous upload": only coordinates travel
    engine.set_particles_compact(s.state[:, :9], None, None)
    assert close(engine.one_to_all_batch(tl), [r.one[t] for t in tl], sc)
    from sc_b200 import ScgpuError
    with pytest.raises(ScgpuError):
        engine.set_particles_compact(s.state[:-1, :9], None, None)        # a different particle count needs the types again
    engine.set_particles_compact(s.state[:, :9], s.type, s.moltype)
