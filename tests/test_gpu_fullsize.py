"""GPU: the BASELINE.json configurations at FULL size, checked through size-independent properties and sampled oracle
comparisons (the oracle finishes a one-to-all over cell lists in microseconds even at 265 k particles):
  L  65 536 PSC bulk (configs[1])             M  265 041-particle CPSC + lipid membrane, NPT box changes (configs[2])
  X  Tests/test_14 (SPA + PSC + CPSC) and Tests/test_20 (bonded TCPSC chains) tiled 12^3 -> 65 664 / 69 120 particles (configs[3])
properties: cell ids / sort order bit-exact; sampled one-to-all energies <= 1e-10; sum of row sums == total;
sum of all one-to-all energies == 2 x total (every pair is in exactly two of them); run-to-run bit reproducibility.
"""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import Engine, synth
from sc_b200.host import HostSystem

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(a, b, rtol=1e-10, atol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol)


def _check_system(top, cfg, variant, nsample, box_scales=()):
    hs = HostSystem(top, cfg)
    s = O.system_from_text(top, cfg)
    assert np.array_equal(hs.state, s.state) and np.array_equal(hs.ia, s.ia)      # product parser == oracle parser
    eng = Engine(0, variant).load(hs)
    cells = s.cells()
    ncell, cell_of, order, start = cells
    g_cell, g_nc = eng.cell_assignment()
    assert np.array_equal(g_nc, ncell) and np.array_equal(g_cell, cell_of)
    g_order, g_start = eng.cell_order(int(ncell.prod()))
    assert np.array_equal(g_order, order) and np.array_equal(g_start, start)
    ev, ncand, ngate = eng.one_to_all_everyone(count=True)
    sample = list(range(0, s.n, max(1, s.n // nsample)))
    oc = og = 0
    for t in sample:
        e, c, g = s.one_to_all_cells(t, cells)
        assert close(ev[t], e), (t, ev[t], e)
        oc += c
        og += g
    eb = eng.one_to_all_batch(sample)
    assert close(eb, ev[sample])
    ev2 = eng.one_to_all_everyone()
    assert np.array_equal(ev2, eng.one_to_all_everyone())               # deterministic reductions: the same call gives the same bits
    assert close(ev, ev2)                                               # (the counting pass lists extra zero-energy pairs: other lane order)
    tot, rows = eng.all_to_all(rows=True)
    assert close(rows.sum(), tot, atol=1e-7)
    assert close(ev.sum(), 2.0 * tot, rtol=1e-9, atol=1e-6)
    # NPT: the caller rescales the box (positions are fractional); energies must follow, grid must re-form when needed
    for f in box_scales:
        s.box = s.box * np.array(f)
        eng.set_box(s.box)
        cells = s.cells()
        assert np.array_equal(eng.cell_assignment()[1], cells[0])
        for t in sample[:: max(1, len(sample) // 24)]:
            assert close(eng.one_to_all(t), s.one_to_all_cells(t, cells)[0])
        tot2, rows2 = eng.all_to_all(rows=True)
        assert close(rows2.sum(), tot2, atol=1e-7)
    eng.close()
    return ncand, ngate, tot


@pytest.mark.parametrize("variant", ["fast", "strict"])
def test_L_65536_psc_bulk(variant):
    top, cfg, n = synth.psc_bulk()
    assert n == 65536
    ncand, ngate, tot = _check_system(top, cfg, variant, 256, box_scales=[(1.02, 1.02, 1.02)])
    assert ngate == 4063232 and ncand == 31385024                      # the bench's work counters for this configuration
    assert abs(tot - (-36919.0)) < 5.0                                 # the reference printed E_start = -36919 for this family of lattices


def test_M_265k_membrane_npt():
    inp = json.loads(gzip.open(os.path.join(G, "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(21, 21, inp["top.init"], inp["config.init"])
    assert n == 265041
    _check_system(top, cfg, "fast", 128, box_scales=[(1.01, 1.01, 1.0), (0.97, 0.97, 1.0)])      # ptype 2: xy isotropic, z constant


def test_M_molecule_energies():
    """mol2others over lipids (3-bead chains with bond1 + bond2) of a 4 x 4 membrane tile against the oracle"""
    inp = json.loads(gzip.open(os.path.join(G, "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(4, 4, inp["top.init"], inp["config.init"])
    hs = HostSystem(top, cfg)
    s = O.system_from_text(top, cfg)
    eng = Engine(0, "fast").load(hs)
    first_lipid = 16
    for k in range(0, 3200, 211):
        f = first_lipid + 3 * k
        assert close(eng.mol_to_others(f, 3), s.mol_to_others(f, 3))
    eng.close()


@pytest.mark.parametrize("name", ["test_14_normal_SPA_PSC_CPSC", "test_20_chain_bond12"])
def test_X_mixed_types_and_bonded_chains_tiled(name):
    """BASELINE configs[3] at throughput size: every pair-potential branch (sphere-sphere, rod-sphere, PSC-CPSC) and bonded chains
    (bond1 + bond2 along five-rod molecules) through the flat pipeline of the general (non rods-only) kernels"""
    inp = json.load(open(os.path.join(G, name + ".inputs.json")))
    top, cfg, n = synth.tile(inp["top.init"], inp["config.init"], 12, 12, 12)
    assert n in (65664, 69120)
    ncand, ngate, tot = _check_system(top, cfg, "fast", 96, box_scales=[(0.98, 0.98, 0.98)])
    # periodic tiling: the total is 12^3 times the energy of the shipped configuration (coordinates are printed with 9 digits)
    s0 = O.system_from_text(inp["top.init"], inp["config.init"])
    assert abs(tot - 1728 * s0.all_to_all()) <= 1e-6 * abs(tot)
    assert 0 < ngate <= ncand
