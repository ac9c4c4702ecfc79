"""CPU: pin the oracle (oracle/sc_oracle.c + oracle/topo.py) against outputs of the reference itself.

Fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the unmodified reference
sources (oracle/_ref/sc_ref_driver, SC_testing). The bar here is BIT-EXACT: the oracle is compiled
IEEE-strict in the reference's operation order.
"""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import topo as otopo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(G, "*.ref.gz")))
GRIDS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(G, "grid_*.npz")))

# fields of the 30-double state record that the reference initialises, per geotype
def used_fields(g):
    f = list(range(0, 9))
    if g >= otopo.SP or g in (otopo.SCN, otopo.SCA):
        return f
    f += list(range(12, 18))
    if g in (otopo.TPSC, otopo.TCPSC, otopo.TCHPSC, otopo.TCHCPSC):
        f += list(range(9, 12)) + list(range(18, 24))
    if g in (otopo.CHPSC, otopo.CHCPSC, otopo.TCHPSC, otopo.TCHCPSC):
        f += list(range(24, 27))
    if g in (otopo.TCHPSC, otopo.TCHCPSC):
        f += list(range(27, 30))
    return f


def test_fixture_inventory():
    assert len(DUMPS) >= 50 and len(GRIDS) >= 100


@pytest.mark.parametrize("name", DUMPS)
def test_pair_one_total_bitexact(name):
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    s = r.system
    targets = range(s.n) if s.n <= 500 else range(0, s.n, s.n // 40)
    for t in targets:
        e, ep = s.one_to_all(t, pairs=True)
        ref = np.array([r.pairs.get((t, j), 0.0) for j in range(s.n)])
        assert np.array_equal(ep, ref), (name, t)
        assert e == r.one[t]
        cl = s.conlist(t)
        assert [cl.is_empty] + list(cl.con) == list(r.conlists[t])
    if s.n <= 500:
        assert s.all_to_all() == r.total
        ov = {(i, j) for i in range(s.n) for j in range(i + 1, s.n) if s.overlap_pair(i, j)}
        assert ov == r.overlaps
    for (first, m, e) in r.mol2o[:64]:
        assert s.mol_to_others(first, m) == e


@pytest.mark.parametrize("name", DUMPS)
def test_parsers_and_particle_init(name):
    """oracle-side top.init/config.init parsing + mixing rules + Particle::init == reference dump"""
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    base = name.rsplit("_", 1)[0]
    p = os.path.join(G, base + ".inputs.json")
    if os.path.exists(p):
        inputs = json.load(open(p))
    else:
        inputs = json.loads(gzip.open(p + ".gz").read().decode())
    cfg = inputs["config.init"]
    if name.endswith("_end"):
        cfg = open(os.path.join(G, base + ".config.last")).read()
    s = O.system_from_text(inputs["top.init"], cfg)
    ref = r.system
    assert s.n == ref.n and np.array_equal(s.type, ref.type) and np.array_equal(s.moltype, ref.moltype)
    assert np.array_equal(s.box, ref.box) and s.sqmaxcut == ref.sqmaxcut and s.maxcut == ref.maxcut
    used = sorted(set(ref.type.tolist()))
    for a in used:
        for b in used:
            assert np.array_equal(s.ia[a, b], ref.ia[a, b]), (a, b, np.nonzero(s.ia[a, b] != ref.ia[a, b]))
    assert np.array_equal(s.mol, ref.mol)
    for i in range(s.n):
        f = used_fields(int(ref.ia[ref.type[i], ref.type[i], 0]))
        assert np.array_equal(s.state[i, f], ref.state[i, f]), (i, s.state[i, f] - ref.state[i, f])


@pytest.mark.parametrize("name", GRIDS)
def test_pose_grid_bitexact(name):
    """Interactions_tests-style pose grids (deliberately degenerate axis-aligned poses): E(0,j), E(j,0), overlap"""
    z = np.load(os.path.join(G, name + ".npz"))
    cfg = gzip.open(os.path.join(G, "grid_config_%s.txt.gz" % str(z["kind"]))).read().decode()
    s = O.system_from_text(str(z["top"]), cfg)
    e0, ep = s.one_to_all(0, pairs=True)
    assert np.array_equal(ep, z["e0j"])
    ej0 = np.array([0.0] + [s.pair(j, 0) for j in range(1, s.n)])
    assert np.array_equal(ej0, z["ej0"])
    ov0j = np.array([0] + [s.overlap_pair(0, j) for j in range(1, s.n)])
    ovj0 = np.array([0] + [s.overlap_pair(j, 0) for j in range(1, s.n)])
    assert np.array_equal(ov0j, z["ov0j"]) and np.array_equal(ovj0, z["ovj0"])


def test_pose_grids_are_the_reference_generators_own():
    """the spherocylinder and spherocylinder-sphere pose sets are byte-identical to what the reference's own generator
    (Tests/Interactions_tests/main.cpp) and test.sh loops write (tests/golden/make_golden.py verify_grids recorded the hashes)"""
    import hashlib
    import json
    want = json.load(open(os.path.join(G, "grid_config.sha256")))
    for kind in ("sc", "scsp"):
        txt = gzip.open(os.path.join(G, "grid_config_%s.txt.gz" % kind)).read()
        assert hashlib.sha256(txt).hexdigest() == want[kind], kind


def test_known_answers_test01():
    """two full-precision known answers from test_01's initial configuration"""
    r = O.load_ref_dump(os.path.join(G, "test_01_normal_PSC_init.ref.gz"))
    s = r.system
    assert s.pair(0, 1) == -3.9999990000000003
    assert s.pair(0, 32) == -3.9848704916940747
    assert abs(s.all_to_all() - (-79.0682)) < 1e-4
