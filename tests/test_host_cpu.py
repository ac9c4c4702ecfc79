"""CPU: the C++ host mirror (sc_b200/csrc/host) parses the reference's own input files to exactly the state,
interaction table and cutoffs that the reference itself produced (golden dumps), bit for bit."""
import glob
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200.host import HostSystem, HostError
from test_oracle_golden import used_fields

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = sorted(os.path.basename(p)[:-7] for p in glob.glob(os.path.join(G, "*.ref.gz")))


@pytest.mark.parametrize("name", DUMPS)
def test_host_parser_matches_reference(name):
    r = O.load_ref_dump(os.path.join(G, name + ".ref.gz"))
    base = name.rsplit("_", 1)[0]
    p = os.path.join(G, base + ".inputs.json")
    inputs = json.load(open(p)) if os.path.exists(p) else json.loads(gzip.open(p + ".gz").read().decode())
    cfg = inputs["config.init"]
    if name.endswith("_end"):
        cfg = open(os.path.join(G, base + ".config.last")).read()
    h = HostSystem(inputs["top.init"], cfg)
    ref = r.system
    assert h.n == ref.n and np.array_equal(h.type, ref.type) and np.array_equal(h.moltype, ref.moltype)
    assert np.array_equal(h.box, ref.box) and h.sqmaxcut == ref.sqmaxcut and h.maxcut == ref.maxcut
    used = sorted(set(ref.type.tolist()))
    for a in used:
        for b in used:
            assert np.array_equal(h.ia[a, b], ref.ia[a, b]), (a, b)
    assert np.array_equal(h.mol, ref.mol)
    for i in range(h.n):
        f = used_fields(int(ref.ia[ref.type[i], ref.type[i], 0]))
        assert np.array_equal(h.state[i, f], ref.state[i, f]), i
    h.close()


def test_host_parser_errors():
    with pytest.raises(HostError):
        HostSystem("[Types]\nA 1 PSC 1.0 1.2\n[Molecules]\nA: {\nparticles: 1\n}\n[System]\nA 1\n", "10 10 10\n1 1 1 1 0 0 0 1 0 0\n")
    with pytest.raises(HostError):
        HostSystem("[Types]\nA 1 BOGUS 1.0 1.2\n", "10 10 10\n")
    with pytest.raises(HostError):   # null direction vector for a rod
        HostSystem("[Types]\nA 1 SCN 1.0 1.2 3\n[Molecules]\nA: {\nparticles: 1\n}\n[System]\nA 1\n", "10 10 10\n1 1 1 0 0 0 0 1 0 0\n")
