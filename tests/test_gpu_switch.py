"""GPU: scgpu_set_particle_type (MoveCreator::switchTypeMove, scOOP/mc/movecreator.cpp:233-303: the type of one particle changes, its
state record stays the caller's business). Energies after type changes against the oracle evaluated with the same type array: on a
two-type rod mixture (Tests/test_11: PSC + CPSC) driven through every specialisation the census of types selects -- mixed, a single type
left (the one-type kernels), mixed again -- and on the sphere / rod mixture of Tests/test_14. The byte-identical run of the reference's
own switch moves is in tests/test_gpu_dropin.py."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import Engine

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(a, b, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= 1e-10 * np.maximum(np.abs(a), np.abs(b)) + 1e-10 * scale)


def check(eng, s):
    sc = max(1.0, float(np.max(np.abs(s.ia[:, :, 4]))))
    ev = eng.one_to_all_everyone()
    want = np.array([s.one_to_all(t) for t in range(s.n)])
    assert close(ev, want, sc), float(np.max(np.abs(ev - want)))
    tot = eng.all_to_all()
    assert close(tot, s.all_to_all(), sc * 10)
    for t in range(0, s.n, 9):
        assert close(eng.one_to_all(t), want[t], sc)


@pytest.mark.parametrize("variant", ["fast", "strict"])
def test_type_switches_follow_through_every_specialisation(variant):
    r = O.load_ref_dump(os.path.join(G, "test_11_normal_PSC_CPSC_init.ref.gz"))
    s = r.system
    a, b = [int(t) for t in np.unique(s.type)]
    eng = Engine(0, variant).load(s)
    check(eng, s)
    rng = np.random.default_rng(3)
    for i in rng.choice(s.n, size=12, replace=False):                 # a few switches in a mixture
        new = b if s.type[i] == a else a
        s.type[i] = new
        eng.set_particle_type(int(i), new)
    check(eng, s)
    for i in np.nonzero(s.type == b)[0]:                              # every particle of type b becomes a: a single type is left
        s.type[i] = a
        eng.set_particle_type(int(i), a)
    assert len(np.unique(s.type)) == 1
    check(eng, s)
    for i in range(0, s.n, 5):                                        # and back to a mixture
        s.type[i] = b
        eng.set_particle_type(int(i), b)
    check(eng, s)
    # a trial of a particle whose type has just changed (the order of calls of switchTypeMove): old energy, new type, trial energy
    i = 7
    e_old = eng.one_to_all(i)
    new = a if s.type[i] == b else b
    s.type[i] = new
    eng.set_particle_type(i, new)
    st = s.state[i].copy()
    e_new = eng.one_to_all(i, st)
    assert close(e_new, s.one_to_all(i, st), 10.0) and e_new != e_old
    eng.set_particle_type(i, new)                                     # same type again: no-op
    with pytest.raises(Exception):
        eng.set_particle_type(i, 99)
    eng.close()


def test_type_switch_between_sphere_and_rod_types():
    r = O.load_ref_dump(os.path.join(G, "test_14_normal_SPA_PSC_CPSC_init.ref.gz"))
    s = r.system
    types = [int(t) for t in np.unique(s.type)]
    eng = Engine(0, "fast").load(s)
    check(eng, s)
    rods = [t for t in types if int(s.ia[t, t, 0]) < 30]
    assert len(rods) >= 2
    idx = np.nonzero(s.type == rods[0])[0][:10]
    for i in idx:                                                     # PSC <-> CPSC inside a system that also holds spheres
        s.type[i] = rods[1]
        eng.set_particle_type(int(i), rods[1])
    check(eng, s)
    eng.close()
