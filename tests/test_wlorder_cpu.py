"""CPU: the oracle's restatement of the Wang-Landau order parameters of a whole configuration (oracle/wl_order.c) against the
REFERENCE itself: WangLandau::zOrder / zOrient / twoPartDist / contParticlesAll / boxSize_x / boxSize_y, Conf::massCenter and
Mesh::meshInit called by oracle/ref_driver.cpp `wlorder` (unmodified reference sources) on Tests/test_mempore, Tests/test_pscthrough,
a multi-type mixture and test_mempore after 300 reference sweeps (tests/golden/*.wlorder.gz, made by tests/golden/make_golden.py
wlorder). Integers bit for bit; the centre of mass bit for bit (same summation order)."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["test_mempore", "test_pscthrough", "extra_mix", "test_mempore.short300"]


def load_case(name):
    base = name.split(".")[0]
    inp = json.loads(gzip.open(os.path.join(G, base + ".inputs.json.gz")).read().decode())
    cfg = inp["config.init"]
    if name.endswith(".short300"):
        cfg = gzip.open(os.path.join(G, name + ".config.last.gz")).read().decode()
    return inp["top.init"], cfg, O.load_wlorder_dump(os.path.join(G, name + ".wlorder.gz"))


@pytest.mark.parametrize("name", CASES)
def test_wl_order_oracle_is_bit_exact_against_the_reference(name):
    top, cfg, d = load_case(name)
    s = O.system_from_text(top, cfg)
    assert s.n == d["n"] and np.array_equal(s.box, d["box"])
    vol = O.type_volumes(s)
    for t, v in d["vol"].items():
        assert vol[t] == v, t                                              # Ia_param::volume, topo.cpp:388-392
    cm, sysvol = O.wl_mass_center(s)
    assert sysvol == d["sysvolume"] and np.array_equal(cm, d["syscm"])    # Conf::massCenter, inicializer.cpp:32-34
    assert len(d["bins"]) == 3
    for (mn, dd), w in d["bins"].items():
        assert O.wl_bin(1, O.wl_raw(s, 1), mn, dd) == w["W1"]
        assert O.wl_bin(3, O.wl_raw(s, 3), mn, dd) == w["W3"]
        assert O.wl_bin(4, O.wl_raw(s, 4), mn, dd) == w["W4"]
        assert O.wl_bin(8, O.wl_raw(s, 8), mn, dd) == w["W8"]
        assert O.wl_bin(9, O.wl_raw(s, 9), mn, dd) == w["W9"]
        for t, (cnt, order) in w["W7"].items():
            c = O.wl_raw(s, 7, wlmtype=t)
            assert c == cnt and O.wl_bin(7, c, mn, dd) == order, t
    assert len(d["mesh"]) >= 8
    holes = set()
    for (t, meshsize, d0, d1, maxsize, occupied, order), (nholes, fnv) in zip(d["mesh"], d["mesh_labels"]):
        lab = O.wl_mesh_labels(s, t, meshsize)                            # Mesh::data after findHoles: hole numbers in scan order
        assert lab.shape == (d1, d0) and max(int(lab.max()), 0) == nholes and O.mesh_hash(lab) == fnv, (t, meshsize)
        assert np.count_nonzero(lab < 0) == occupied and (nholes == 0 or np.bincount(lab[lab > 0]).max() == maxsize)
        m, dim, occ, skip = O.wl_raw(s, 2, wlmtype=t, meshsize=meshsize)
        assert dim == (d0, d1) and skip == 0
        assert occ == occupied, (t, meshsize)                             # Mesh::meshFill
        assert m == maxsize, (t, meshsize)                                # Mesh::findHoles
        assert O.wl_bin(2, m, 1.0, 4.0) == order
        holes.add(m)
    assert len(holes) >= 5                                               # the fixtures really hold holes of several sizes


def test_mesh_hole_of_known_shapes():
    """hand-made meshes: one particle occupies its 3 x 3 block; holes wrap around the periodic box"""
    box = np.array([10.0, 10.0, 10.0])
    st = np.zeros((2, 30))
    typ = np.array([1, 1], dtype=np.int32)
    st[0, :3] = [0.05, 0.05, 0.5]          # mesh point (0, 0) of a 10 x 10 mesh: block wraps to rows / columns 9, 0, 1
    st[1, :3] = [0.55, 0.55, 0.5]          # (5, 5): block 4..6
    L = O.lib()
    dim = np.zeros(2, dtype=np.int32)
    data = np.zeros(100, dtype=np.int32)
    import ctypes as C
    occ, skip = C.c_long(), C.c_long()
    m = L.sco_wl_mesh_hole(2, O._d(st), O._i(typ), 1, O._d(box), 1.0, O._i(dim), O._i(data), C.byref(occ), C.byref(skip))
    assert tuple(dim) == (10, 10) and occ.value == 18 and skip.value == 0
    assert m == 100 - 18                    # everything else is one connected hole
    grid = data.reshape(10, 10)             # [y][x]
    assert grid[0, 0] == -1 and grid[9, 9] == -1 and grid[1, 9] == -1 and grid[5, 5] == -1 and grid[2, 2] == 0
    # a wall of particles along x cuts nothing (periodic in y), two walls make two holes
    n = 20
    st = np.zeros((n, 30))
    typ = np.ones(n, dtype=np.int32)
    for k in range(10):
        st[k, :3] = [(k + 0.5) / 10, 0.15, 0.5]       # rows 0..2 occupied
        st[10 + k, :3] = [(k + 0.5) / 10, 0.65, 0.5]  # rows 5..7 occupied
    m = L.sco_wl_mesh_hole(n, O._d(st), O._i(typ), 1, O._d(box), 1.0, O._i(dim), None, C.byref(occ), C.byref(skip))
    assert occ.value == 60 and m == 20      # free rows 3, 4 (20 points) and 8, 9 (20 points)
    # particles of another type leave the mesh empty: one hole of every point
    m = L.sco_wl_mesh_hole(n, O._d(st), O._i(typ), 2, O._d(box), 1.0, O._i(dim), None, C.byref(occ), C.byref(skip))
    assert occ.value == 0 and m == 100
