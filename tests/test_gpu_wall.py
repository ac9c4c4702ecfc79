"""GPU: the [EXTER] wall potential on the device (sc_b200/csrc/wall.cuh, scgpu_set_exter). Every energy entry point adds
ExternalEnergyCalculator::extere2 of the particles it sums over, as the reference's calculators do
(scOOP/mc/totalenergycalculator.h:348-350, 377-378, 410-411, 435-449); the sweeps include it in both energies of a trial.
Checked against the reference's own dumps (per-particle extere2, pair sums), the oracle, and -- the reference's regression
method -- the byte-identical config.last of Tests/test_wallfibril through the sequential driver and through the reference's
own main() on the GPU calculator (oracle/_ref/SC_scgpu)."""
import gzip
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle as O
from sc_b200 import Engine
from sc_b200.engine import MoveParams
from sc_b200.host import HostSystem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def close(a, b, rtol=1e-10, atol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol)


def _inputs(name):
    return json.loads(gzip.open(os.path.join(G, name + ".inputs.json.gz")).read().decode())


@pytest.mark.parametrize("variant", ["fast", "strict"])
@pytest.mark.parametrize("name", ["wall_mix", "wall_fibril"])
def test_wall_terms_against_reference_dumps(name, variant):
    inp = _inputs(name)
    hs = HostSystem(inp["top.init"], inp["config.init"])
    assert hs.exter is not None
    s = O.system_from_text(inp["top.init"], inp["config.init"])
    d = O.load_exter_dump(os.path.join(G, name + ".exter.gz"))
    r = O.load_ref_dump(os.path.join(G, name + "_init.ref.gz"))
    ext = d["ext"]
    eng = Engine(0, variant).load(hs)
    ev = eng.one_to_all_everyone()
    want = np.array([r.one[t] for t in range(s.n)]) + ext            # oneToAll = pair sum + extere2(target)
    scale = max(1.0, float(np.max(np.abs(want))))
    assert close(ev, want, atol=1e-10 * scale), float(np.max(np.abs(ev - want)))
    tot, rows = eng.all_to_all(rows=True)
    assert close(tot, r.total + ext.sum(), rtol=1e-10, atol=1e-9 * scale)         # allToAll = pairs + sum of extere2
    for t in range(0, s.n, 7):
        assert close(eng.one_to_all(t), want[t], atol=1e-10 * scale)
    # trial states: the target moved towards / away from the wall and turned
    rng = np.random.default_rng(5)
    for t in range(0, s.n, 11):
        st = s.state[t].copy()
        st[2] += rng.normal(scale=0.02)
        e = eng.one_to_all(t, st)
        o = s.one_to_all(t, st) + s.extere2(t, st)
        assert close(e, o, atol=1e-10 * scale), (t, e, o)
    # without the wall the same engine gives the pair sums alone
    eng.set_exter(None)
    assert close(eng.one_to_all_everyone(), np.array([r.one[t] for t in range(s.n)]), atol=1e-10 * scale)
    eng.close()
    hs.close()


@pytest.mark.parametrize("rule", [0, 2])
def test_sweeps_with_wall_keep_the_energy_books(rule):
    """rule 0: the round kernel; rule 2: the phased sweep (the wall term enters in k_sweep_resolve)"""
    inp = _inputs("wall_fibril")
    hs = HostSystem(inp["top.init"], inp["config.init"])
    eng = Engine(0, "fast").load(hs)
    mp = MoveParams()
    mp.trial_rule = rule
    mp.temper = 1.0
    for k in range(40):
        mp.trans_mx[k] = 0.1
        mp.rot_angle[k] = 0.1
    mp.n_sub = 1
    e0 = eng.all_to_all()
    de = 0.0
    acc = 0
    for sw in range(20):
        st = eng.sweep(mp, 99, sw)
        de += st.energy_delta
        acc += st.trans_acc + st.rot_acc
    e1 = eng.all_to_all()
    assert acc > 100
    assert abs((e1 - e0) - de) <= 1e-9 * max(abs(e0), abs(e1), 1.0), (e0, e1, de)
    eng.close()
    hs.close()


def test_wallfibril_trajectory_is_byte_identical():
    """300 sweeps of Tests/test_wallfibril (400 CPSC at a wall) through the sequential driver: every energy, wall term included, from the device"""
    inp = _inputs("wall_fibril")
    hs = HostSystem(inp["top.init"], inp["config.init"])
    st = hs.run_mc(inp["options"], 0, 300)
    got = hs.config_last(True)
    hs.close()
    assert abs(st["drift"]) < 1e-8 * max(1.0, abs(st["e_end"]))
    assert got == open(os.path.join(G, "test_wallfibril.short300.config.last")).read(), st


def test_wallfibril_through_the_reference_main():
    sc = os.path.join(ROOT, "oracle", "_ref", "SC_scgpu")
    if not os.path.exists(sc):
        pytest.skip("oracle/_ref/SC_scgpu was not built")
    inp = _inputs("wall_fibril")
    with tempfile.TemporaryDirectory(prefix="wall_") as tmp:
        opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = 300", inp["options"])
        for fn, text in (("options", opt), ("top.init", inp["top.init"]), ("config.init", inp["config.init"])):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(text)
        r = subprocess.run([sc], cwd=tmp, capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        got = open(os.path.join(tmp, "config.last")).read()
    assert got == open(os.path.join(G, "test_wallfibril.short300.config.last")).read()
