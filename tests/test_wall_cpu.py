"""CPU: the oracle's restatement of the [EXTER] wall potential (oracle/sc_oracle.c: wall_energy, sco_exter_params) against the
REFERENCE itself: ExternalEnergyCalculator::extere2 of every particle and the derived parameters topo.exter.interactions[] as dumped
by oracle/ref_driver.cpp `exter` (unmodified reference sources) for (a) Tests/test_wallfibril's initial configuration and (b) a slab
with every geotype at random heights and orientations around the wall (tests/golden/make_golden.py wall). Bit for bit."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["wall_mix", "wall_fibril"])
def test_wall_oracle_is_bit_exact_against_the_reference(name):
    inp = json.loads(gzip.open(os.path.join(G, name + ".inputs.json.gz")).read().decode())
    s = O.system_from_text(inp["top.init"], inp["config.init"])
    d = O.load_exter_dump(os.path.join(G, name + ".exter.gz"))
    assert s.exter is not None and d["exter"][0] == 1
    assert tuple(s.exter) == d["exter"][1:4]
    par, sq = s.exter_setup()
    assert sq == d["exter"][4]                                   # topo.exter.sqmaxcut (topo.cpp:151-152)
    for t, v in d["params"].items():
        assert int(s.ia[t, t, 0]) == v[0] and np.array_equal(par[t], np.array(v[1:])), t      # topo.cpp:120-130
    e = np.array([s.extere2(i) for i in range(s.n)])
    assert np.array_equal(e, d["ext"])
    assert np.count_nonzero(e) > 90                              # the slab really touches the wall
    if name == "wall_mix":                                       # every geotype contributes
        for t in range(1, 11):
            assert np.count_nonzero(e[s.type == t]) > 5, t
