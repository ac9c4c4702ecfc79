"""GPU: Wang-Landau order parameters of a whole configuration on the device (sc_b200/csrc/wl_order.cuh, scgpu_wl_order) against the
REFERENCE's own members -- WangLandau::zOrder / zOrient / twoPartDist / contParticlesAll / boxSize_x / boxSize_y, Conf::massCenter,
Mesh::meshInit (mesh fill + hole search) as dumped by oracle/ref_driver.cpp `wlorder` from the unmodified reference sources
(tests/golden/*.wlorder.gz) -- and against the oracle (oracle/wl_order.c, itself bit-exact against those dumps) where the
reference has no fixture: the 265 041-particle membrane, configurations changed by device sweeps, random sparse meshes.
Bins, hole sizes, occupied mesh points, contact counts: bit-exact. Centre of mass: 1e-12 (block-wise summation order)."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import Engine, synth
from sc_b200.engine import MoveParams, ScgpuError

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["test_mempore", "test_pscthrough", "extra_mix", "test_mempore.short300"]


def load_case(name):
    base = name.split(".")[0]
    inp = json.loads(gzip.open(os.path.join(G, base + ".inputs.json.gz")).read().decode())
    cfg = inp["config.init"]
    if name.endswith(".short300"):
        cfg = gzip.open(os.path.join(G, name + ".config.last.gz")).read().decode()
    return O.system_from_text(inp["top.init"], cfg), O.load_wlorder_dump(os.path.join(G, name + ".wlorder.gz"))


@pytest.mark.parametrize("variant", ["fast", "strict"])
@pytest.mark.parametrize("name", CASES)
def test_order_parameters_against_reference_dumps(name, variant):
    s, d = load_case(name)
    eng = Engine(0, variant).load(s)
    w = eng.wl_order(1, minorder=0.0, dorder=1.0)
    assert np.allclose(np.array(w.syscm[:]), d["syscm"], rtol=0, atol=1e-12) and abs(w.sysvolume - d["sysvolume"]) <= 1e-12 * d["sysvolume"]
    for (mn, dd), ref in d["bins"].items():
        for wlm, key in ((1, "W1"), (3, "W3"), (4, "W4"), (8, "W8"), (9, "W9")):
            w = eng.wl_order(wlm, minorder=mn, dorder=dd)
            assert w.order[0] == ref[key], (wlm, mn, dd, w.raw[0])
            assert abs(w.raw[0] - O.wl_raw(s, wlm)) <= 1e-12 * max(1.0, abs(w.raw[0]))
        for t, (cnt, order) in ref["W7"].items():
            w = eng.wl_order(7, wlmtype=t, minorder=mn, dorder=dd)
            assert w.raw[0] == cnt and w.order[0] == order, (t, mn, dd)
    for t, meshsize, d0, d1, maxsize, occupied, order in d["mesh"]:
        w = eng.wl_order(2, wlmtype=t, minorder=1.0, dorder=4.0, meshsize=meshsize)
        assert tuple(w.mesh_dim[:]) == (d0, d1) and w.mesh_skipped == 0
        assert w.mesh_occupied == occupied, (t, meshsize)                   # Mesh::meshFill
        assert w.raw[0] == maxsize and w.order[0] == order, (t, meshsize)   # Mesh::findHoles
    for (t, meshsize, d0, d1, maxsize, occupied, order), (nholes, fnv) in zip(d["mesh"], d["mesh_labels"]):
        eng.wl_order(2, wlmtype=t, meshsize=meshsize)
        lab = eng.wl_mesh((d0, d1))                                          # Mesh::data as findHoles leaves it, bit for bit
        assert max(int(lab.max()), 0) == nholes and O.mesh_hash(lab) == fnv, (t, meshsize)
        assert np.array_equal(lab, eng.wl_mesh((d0, d1)))                    # a second read returns the same array
    # two dimensions at once, as `wlm = 2 1` of an options file
    t, meshsize, _, _, maxsize, _, _ = d["mesh"][1]
    (mn, dd), ref = next(iter(d["bins"].items()))
    w = eng.wl_order((2, 1), wlmtype=t, minorder=(0.0, mn), dorder=(1.0, dd), meshsize=meshsize)
    assert w.order[0] == maxsize and w.order[1] == ref["W1"]
    # repeated calls give the same bits (fixed summation order, integer atomics)
    a = eng.wl_order(1, minorder=0.0, dorder=1.0)
    b = eng.wl_order(1, minorder=0.0, dorder=1.0)
    assert a.raw[0] == b.raw[0] and a.syscm[:] == b.syscm[:]
    eng.close()


def test_refused_methods_and_arguments():
    s, d = load_case("test_pscthrough")
    eng = Engine(0, "fast").load(s)
    for wlm in (5, 6, 10, -1):
        with pytest.raises(ScgpuError):
            eng.wl_order(wlm)
    with pytest.raises(ScgpuError):
        eng.wl_order(2, wlmtype=1, meshsize=0.0)
    with pytest.raises(ScgpuError):
        eng.wl_order(2, wlmtype=99, meshsize=0.3)
    with pytest.raises(ScgpuError):
        eng.wl_order(1, dorder=0.0)
    with pytest.raises(ScgpuError):
        eng.wl_mesh((10, 10))                                                # no mesh yet
    w = eng.wl_order(2, wlmtype=1, meshsize=0.5)
    with pytest.raises(ScgpuError):
        eng.wl_mesh((w.mesh_dim[0] + 1, w.mesh_dim[1]))                      # wrong size
    eng.close()


def test_after_device_sweeps_and_a_box_change():
    """the sweeps leave the newest configuration in the cell-sorted arrays: the order parameters must see it"""
    s, d = load_case("test_mempore")
    eng = Engine(0, "fast").load(s)
    mp = MoveParams()
    mp.temper, mp.n_sub, mp.trial_rule = 1.0, 1, 1
    for k in range(40):
        mp.trans_mx[k] = 0.2
        mp.rot_angle[k] = 0.1
    for sweep in range(3):
        eng.sweep(mp, 77, sweep)
    w1 = eng.wl_order((2, 7), wlmtype=2, minorder=(0.0, 0.0), dorder=(1.0, 1.0), meshsize=1.0 / 8.0)
    wz = eng.wl_order(1, minorder=-30.0, dorder=0.125)
    st = eng.download_particles()
    assert not np.array_equal(st[:, :3], s.state[:, :3])
    s2 = O.System(st, s.type, s.moltype, s.ia, s.mol, s.box, s.sqmaxcut, s.maxcut)
    m, dim, occ, skip = O.wl_raw(s2, 2, wlmtype=2, meshsize=1.0 / 8.0)
    assert (w1.raw[0], tuple(w1.mesh_dim[:]), w1.mesh_occupied, w1.mesh_skipped) == (m, dim, occ, skip)
    assert w1.raw[1] == O.wl_raw(s2, 7, wlmtype=2)
    z = O.wl_raw(s2, 1)
    assert abs(wz.raw[0] - z) <= 1e-11 and wz.order[0] == O.wl_bin(1, wz.raw[0], -30.0, 0.125)
    # NPT: the mesh follows the box (Mesh::meshInit takes conf->geo.box), fractional positions stay
    box2 = s.box * np.array([1.07, 0.93, 1.0])
    eng.set_box(box2)
    s2.box = np.ascontiguousarray(box2)
    w2 = eng.wl_order((2, 8), wlmtype=2, minorder=(0.0, 10.0), dorder=(1.0, 0.5), meshsize=1.0 / 8.0)
    m, dim, occ, skip = O.wl_raw(s2, 2, wlmtype=2, meshsize=1.0 / 8.0)
    assert (w2.raw[0], tuple(w2.mesh_dim[:]), w2.mesh_occupied) == (m, dim, occ)
    assert w2.order[1] == O.wl_bin(8, box2[0], 10.0, 0.5)
    eng.close()


def test_hole_search_on_random_sparse_meshes():
    """union-find against the breadth-first walk on meshes near the percolation threshold (many holes of every size, wrapped
    around the periodic box), incl. meshes one or two points wide"""
    rng = np.random.default_rng(20)
    s, _ = load_case("test_pscthrough")
    eng = Engine(0, "fast")
    for n, box, meshsize in ((4000, (60.0, 45.0, 10.0), 0.5), (20000, (200.0, 200.0, 10.0), 0.5), (12, (40.0, 1.9, 10.0), 1.0),
                             (50, (30.0, 1.0, 5.0), 1.0), (2500, (64.0, 64.0, 5.0), 1.0), (9000, (333.0, 77.0, 5.0), 0.37)):
        st = np.zeros((n, 30))
        st[:, :3] = rng.uniform(-1.5, 1.5, size=(n, 3))           # box fractions, not wrapped: INBOX() must fold them
        st[:, 5] = 1.0
        typ = np.ones(n, dtype=np.int32)
        typ[rng.uniform(size=n) < 0.3] = 2
        sys2 = O.System(st, typ, np.zeros(n, dtype=np.int32), s.ia, s.mol, np.array(box), s.sqmaxcut, s.maxcut)
        eng.load(sys2)
        for t in (1, 2):
            w = eng.wl_order(2, wlmtype=t, meshsize=meshsize)
            m, dim, occ, skip = O.wl_raw(sys2, 2, wlmtype=t, meshsize=meshsize)
            assert (w.raw[0], tuple(w.mesh_dim[:]), w.mesh_occupied, w.mesh_skipped) == (m, dim, occ, skip), (n, box, t)
            assert m < dim[0] * dim[1]
            assert np.array_equal(eng.wl_mesh(dim), O.wl_mesh_labels(sys2, t, meshsize)), (n, box, t)
    eng.close()


def test_membrane_265k_hole_and_centre_of_mass():
    """BASELINE configs[2]: the tiled membrane; mesh of 685 x 685 points (sigma / 3 of the lipid tails) and 1371 x 1371"""
    inp = json.loads(gzip.open(os.path.join(G, "membrane601.inputs.json.gz")).read().decode())
    top, cfg, n = synth.membrane(21, 21, inp["top.init"], inp["config.init"])
    s = O.system_from_text(top, cfg)
    assert s.n == 265041
    eng = Engine(0, "fast").load(s)
    tail = int(s.type[-1])
    sig = float(s.ia[tail, tail, 3])
    for meshsize in (sig / 3.0, sig / 6.0):
        w = eng.wl_order((2, 1), wlmtype=tail, minorder=(0.0, -30.0), dorder=(1.0, 0.01), meshsize=meshsize)
        m, dim, occ, skip = O.wl_raw(s, 2, wlmtype=tail, meshsize=meshsize)
        assert (w.raw[0], tuple(w.mesh_dim[:]), w.mesh_occupied, w.mesh_skipped) == (m, dim, occ, skip)
        cm, vol = O.wl_mass_center(s)
        assert np.allclose(np.array(w.syscm[:]), cm, rtol=0, atol=1e-12) and abs(w.sysvolume - vol) <= 1e-10 * vol      # 265 041 terms: the sequential sum carries ~3e-12
        assert abs(w.raw[1] - O.wl_raw(s, 1)) <= 1e-10
        assert np.array_equal(eng.wl_mesh(dim), O.wl_mesh_labels(s, tail, meshsize))
    w = eng.wl_order(7, wlmtype=tail)
    assert w.raw[0] == O.wl_raw(s, 7, wlmtype=tail) and w.raw[0] > 0
    eng.close()
