"""GPU: whole reference runs reproduced through the CUDA energy path. The C++ host mirror replays the reference's
sequential Monte-Carlo loop (Ran2 stream, proposals, Metropolis test: sc_b200/csrc/host/mc_driver.cpp) with every energy
coming from the C ABI (TotalEGpu -> scgpu_*), and the resulting config.last must be BYTE-IDENTICAL to the one the unmodified
reference program wrote (tests/golden/*.config.last, made by tests/golden/make_golden.py): 76 000 accept/reject decisions
per case all have to come out the same. This is the reference's own regression method (Tests/test: diff config.last)."""
import json
import os
import re

import pytest

from sc_b200.host import HostSystem

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NVT = ["test_01_normal_PSC", "test_02_normal_CPSC", "test_03_normal_CHPSC", "test_04_normal_CHCPSC", "test_05_normal_TPSC",
       "test_06_normal_TCPSC", "test_07_normal_TCHPSC", "test_08_normal_TCHCPSC", "test_09_normal_SPN", "test_10_normal_SPA",
       "test_11_normal_PSC_CPSC", "test_12_normal_SPA_CPSC", "test_13_normal_SPA_PSC", "test_14_normal_SPA_PSC_CPSC",
       "test_20_chain_bond12", "test_21_chain_bondd2"]
NPT = ["volumeChange_%d%s" % (k, hl) for k in range(6) for hl in "hl"]      # 4, 5: tests/golden/make_golden.py ptype45


def _run(name, golden_file, nsweeps=0):
    inputs = json.load(open(os.path.join(G, name + ".inputs.json")))
    hs = HostSystem(inputs["top.init"], inputs["config.init"])
    st = hs.run_mc(inputs["options"], 0, nsweeps)
    got = hs.config_last(True)
    want = open(os.path.join(G, golden_file)).read()
    hs.close()
    return got, want, st


@pytest.mark.parametrize("name", NVT)
def test_nvt_trajectory_is_byte_identical(name):
    """first 300 sweeps (11 400 - 12 000 trial moves) of every NVT regression case of the reference"""
    got, want, st = _run(name, name + ".short300.config.last", 300)
    assert abs(st["drift"]) < 1e-8 * max(1.0, abs(st["e_end"]))        # the reference's own energy-drift self check
    assert got == want, (name, st)


@pytest.mark.parametrize("name", ["test_01_normal_PSC", "test_14_normal_SPA_PSC_CPSC", "test_20_chain_bond12"])
def test_full_length_run_is_byte_identical(name):
    """the complete runs exactly as Tests/test performs them (1 000 - 2 000 sweeps)"""
    got, want, st = _run(name, name + ".config.last")
    assert got == want, (name, st)


@pytest.mark.parametrize("name", NPT)
def test_npt_trajectory_is_byte_identical(name):
    """pressure moves (ptype 0-5, high and low pressure): allToAll / allToAllTrial / update() through the GPU path"""
    got, want, st = _run(name, name + ".short500.config.last", 500)
    assert st["edge_acc"] + st["edge_rej"] > 100
    assert got == want, (name, st)


@pytest.mark.parametrize("name", ["test_01_normal_PSC", "test_20_chain_bond12", "volumeChange_0h"])
def test_equilibration_with_step_size_adaptation_is_byte_identical(name):
    """nequil = 120, adjust = 10, nsweeps = 100: Updater::optimizeStep / optimizeRot (mc/updater.cpp:238-253, 395-465) adapt the
    per-type displacement, chain-displacement and box-edge steps during the first half of the equilibration; the run then
    continues at the adapted sizes (main.cpp:270-297). Every later proposal depends on the adapted sizes, so a byte-identical
    config.last pins the whole adaptation history."""
    inputs = json.load(open(os.path.join(G, name + ".inputs.json")))
    options = open(os.path.join(G, name + ".equil.options")).read()
    assert re.search(r"(?m)^nequil\s*=\s*120", options) and re.search(r"(?m)^adjust\s*=\s*10", options)
    hs = HostSystem(inputs["top.init"], inputs["config.init"])
    st = hs.run_mc(options, 0, 0)
    got = hs.config_last(True)
    hs.close()
    want = open(os.path.join(G, name + ".equil.config.last")).read()
    assert got == want, (name, st)
