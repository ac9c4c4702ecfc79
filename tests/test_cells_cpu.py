"""CPU: the cell-list definition C1 (ours -- the reference has no cell list) is pair-set equivalent to the
reference's all-pairs sums: the oracle's 27-cell + bonded-partner one-to-all equals its all-pairs one-to-all
bit for bit (skipped pairs contribute exactly 0)."""
import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import synth


@pytest.mark.parametrize("kind", ["psc_lattice", "psc_gas", "mix", "chains"])
def test_cell_path_equals_all_pairs(kind):
    top, cfg = synth.small_case(kind)
    s = O.system_from_text(top, cfg)
    cells = s.cells()
    ncell, cell_of, order, start = cells
    assert ncell.prod() > 27 and start[-1] == s.n
    assert sorted(order.tolist()) == list(range(s.n))
    # stable: ascending original index inside every cell, cells ascending
    assert all(np.all(np.diff(order[start[c]:start[c + 1]]) > 0) for c in range(len(start) - 1))
    assert np.all(np.diff(cell_of[order]) >= 0)
    for t in range(0, s.n, 7):
        e_cells, ncand, ngate = s.one_to_all_cells(t, cells)
        assert e_cells == s.one_to_all(t)
        assert ngate <= ncand < s.n


def test_cell_binning_edges():
    """INBOX convention (macros.h:119): u <= 0 wraps by +1, exact 1.0 products fold back to cell 0"""
    top, cfg = synth.small_case("psc_gas")
    s = O.system_from_text(top, cfg)
    s.state[0, 0:3] = [0.0, -0.0, 1.0]
    s.state[1, 0:3] = [-1e-20, 1.0 - 1e-17, -3.25]
    s.state[2, 0:3] = [5.75, -7.0, 0.999999999999]
    ncell, cell_of, order, start = s.cells()
    nx, ny, nz = ncell

    def coords(c):
        return (c % nx, (c // nx) % ny, c // (nx * ny))

    assert coords(cell_of[0]) == (0, 0, 0)
    assert coords(cell_of[1]) == (0, 0, int(0.75 * nz))
    assert coords(cell_of[2]) == (int(0.75 * nx), 0, nz - 1)


def test_small_box_degrades_to_one_cell():
    """test_01-sized boxes (10^3, maxcut 5.11): fewer than 3 cells per axis -> a single cell, no double images"""
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    r = O.load_ref_dump(os.path.join(G, "test_01_normal_PSC_init.ref.gz"))
    s = r.system
    cells = s.cells()
    assert cells[0].tolist() == [1, 1, 1]
    for t in range(s.n):
        assert s.one_to_all_cells(t, cells)[0] == r.one[t]


def test_tiler_keeps_energy_per_copy():
    """synth.tile (configs[3] at throughput size): a periodic 2 x 2 x 2 replication has 8 times the energy of the original,
    also for bonded chains whose members sit on opposite sides of the small box"""
    import json
    import os
    from oracle import oracle as O
    from sc_b200 import synth
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("test_14_normal_SPA_PSC_CPSC", "test_20_chain_bond12"):
        inp = json.load(open(os.path.join(g, name + ".inputs.json")))
        top, cfg, n = synth.tile(inp["top.init"], inp["config.init"], 2, 2, 2)
        s1, s2 = O.system_from_text(inp["top.init"], inp["config.init"]), O.system_from_text(top, cfg)
        assert s2.n == 8 * s1.n == n
        assert abs(s2.all_to_all() - 8 * s1.all_to_all()) <= 1e-6 * abs(s2.all_to_all())
