"""GPU: the drop-in, literally. oracle/_ref/SC_scgpu is the reference's OWN program -- its main.cpp, parsers, move code, random
stream -- with two lines changed by sed on a scratch copy (oracle/Makefile, target scgpu_ref):
    #include "totalegpu.h"
    typedef TotalEGpu<PairE> TotalEnergyCalculator;            (scOOP/mc/totalenergycalculator.h:1226-1228)
i.e. the calculator class of integration/totalegpu.h forwarding every virtual of TotalE<> to the C ABI.
It is run on the reference's own regression inputs (Tests/test_*, Tests/volumeChange/*; committed as tests/golden/*.inputs.json)
and its config.last must be BYTE-IDENTICAL to the one the unmodified reference program wrote (tests/golden/*.config.last) -- the
reference's own regression method (Tests/test: diff config.last config.last2). The binary is built where /root/reference exists
and travels to the GPU box; nothing here reads /root/reference."""
import json
import os
import re
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
SC = os.path.join(ROOT, "oracle", "_ref", "SC_scgpu")

NVT = ["test_01_normal_PSC", "test_02_normal_CPSC", "test_03_normal_CHPSC", "test_04_normal_CHCPSC", "test_05_normal_TPSC",
       "test_06_normal_TCPSC", "test_07_normal_TCHPSC", "test_08_normal_TCHCPSC", "test_09_normal_SPN", "test_10_normal_SPA",
       "test_11_normal_PSC_CPSC", "test_12_normal_SPA_CPSC", "test_13_normal_SPA_PSC", "test_14_normal_SPA_PSC_CPSC",
       "test_20_chain_bond12", "test_21_chain_bondd2"]
NPT = ["volumeChange_%d%s" % (k, hl) for k in range(6) for hl in "hl"]      # 4, 5: tests/golden/make_golden.py ptype45


def run_reference_program(name, nsweeps):
    if not os.path.exists(SC):
        pytest.skip("oracle/_ref/SC_scgpu was not built (make -C oracle scgpu_ref needs /root/reference)")
    inputs = json.load(open(os.path.join(G, name + ".inputs.json")))
    with tempfile.TemporaryDirectory(prefix="dropin_") as tmp:
        opt = inputs["options"]
        if nsweeps:
            opt = re.sub(r"(?m)^nsweeps\s*=\s*\d+", "nsweeps = %d" % nsweeps, opt)
        for fn, text in (("options", opt), ("top.init", inputs["top.init"]), ("config.init", inputs["config.init"])):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(text)
        r = subprocess.run([SC], cwd=tmp, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        return open(os.path.join(tmp, "config.last")).read(), r.stdout


@pytest.mark.parametrize("name", NVT)
def test_reference_main_on_gpu_calculator_nvt(name):
    got, out = run_reference_program(name, 300)
    assert got == open(os.path.join(G, name + ".short300.config.last")).read(), out[-800:]


@pytest.mark.parametrize("name", NPT)
def test_reference_main_on_gpu_calculator_npt(name):
    got, out = run_reference_program(name, 500)
    assert got == open(os.path.join(G, name + ".short500.config.last")).read(), out[-800:]


def test_reference_main_full_length_test_01():
    """Tests/test_01_normal_PSC exactly as Tests/test runs it (BASELINE configs[0])"""
    got, out = run_reference_program("test_01_normal_PSC", 0)
    assert got == open(os.path.join(G, "test_01_normal_PSC.config.last")).read(), out[-800:]


@pytest.mark.parametrize("name", ["test_mempore", "test_pscthrough"])
def test_reference_main_wang_landau_runs(name):
    """Tests/test_mempore (Wang-Landau on the hole in the xy plane of a 500-lipid membrane, NPT ptype 2, chain moves) and
    Tests/test_pscthrough (Wang-Landau on the z position of a PSC crossing the membrane): the order parameters, the mesh hole search
    and the histogram are the reference's own host code above the calculator seam (mc/wanglandau.h, mc/mesh.cpp); every energy --
    single-particle trials, chain moves (mol2others), volume moves (allToAll) -- comes from the device. config.last AND the
    Wang-Landau weights written at the end (wl-new.dat) must equal the unmodified reference's byte for byte."""
    import gzip
    if not os.path.exists(SC):
        pytest.skip("oracle/_ref/SC_scgpu was not built (make -C oracle scgpu_ref needs /root/reference)")
    inputs = json.loads(gzip.open(os.path.join(G, name + ".inputs.json.gz")).read().decode())
    with tempfile.TemporaryDirectory(prefix="dropin_wl_") as tmp:
        for fn in ("options", "top.init", "config.init", "wl.dat"):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(inputs[fn])
        r = subprocess.run([SC], cwd=tmp, capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        got_cfg = open(os.path.join(tmp, "config.last")).read()
        got_wl = open(os.path.join(tmp, "wl-new.dat")).read()
    assert got_cfg == gzip.open(os.path.join(G, name + ".short300.config.last.gz"), "rt").read(), r.stdout[-800:]
    assert got_wl == gzip.open(os.path.join(G, name + ".short300.wl-new.dat.gz"), "rt").read()


def test_reference_main_type_switch_moves():
    """MoveCreator::switchTypeMove (movecreator.cpp:233-303): the reference changes conf->pvec[target].type in place before oneToAllTrial and
    restores it behind the calculator's back on rejection; TotalEGpu follows through scgpu_set_particle_type. Tests/test_wallfibril with
    the switch line of its top.init enabled (CPSC <-> CHCPSC, delta_mu 5, ~2400 attempts, external wall): config.last -- which records
    every particle's `switched` flag, 143 particles end up switched -- byte-identical to the unmodified reference
    (tests/golden/make_golden.py switchmoves)."""
    got, out = run_reference_program("test_wallfibril_switchmoves", 0)
    want = open(os.path.join(G, "test_wallfibril_switchmoves.short300.config.last")).read()
    assert sum(1 for l in want.split("\n")[1:] if l.split() and l.split()[-1] == "1") > 100
    assert got == want, out[-800:]


def test_reference_main_with_the_hole_search_on_the_device():
    """oracle/_ref/SC_scgpu_wl = SC_scgpu with ONE more call of the reference redirected (integration/wl_gpu_hook.h; oracle/Makefile
    scgpu_wl_ref): WangLandau::holeXYPlane(wli) -- the from-scratch membrane-hole order parameter WangLandau::runPress evaluates after every
    volume move (mc/wanglandau.h:168-196, 313-317) -- calls scgpu_mesh_init instead of Mesh::meshInit: mesh fill + hole search on the device
    (scgpu_wl_order), Mesh::data handed back (scgpu_wl_mesh) for the incremental updates of the following single-particle and chain moves,
    which stay the reference's host code. Tests/test_mempore (wlm 2, NPT ptype 2, ~1 volume move per sweep): config.last AND wl-new.dat
    byte-identical to the unmodified reference."""
    import gzip
    sc_wl = os.path.join(ROOT, "oracle", "_ref", "SC_scgpu_wl")
    if not os.path.exists(sc_wl):
        pytest.skip("oracle/_ref/SC_scgpu_wl was not built (make -C oracle scgpu_wl_ref needs /root/reference)")
    name = "test_mempore"
    inputs = json.loads(gzip.open(os.path.join(G, name + ".inputs.json.gz")).read().decode())
    with tempfile.TemporaryDirectory(prefix="dropin_wlgpu_") as tmp:
        for fn in ("options", "top.init", "config.init", "wl.dat"):
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(inputs[fn])
        r = subprocess.run([sc_wl], cwd=tmp, capture_output=True, text=True, timeout=1500)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        got_cfg = open(os.path.join(tmp, "config.last")).read()
        got_wl = open(os.path.join(tmp, "wl-new.dat")).read()
    assert "Mesh::meshInit runs on the device" in r.stderr            # the hook was really taken
    assert got_cfg == gzip.open(os.path.join(G, name + ".short300.config.last.gz"), "rt").read(), r.stdout[-800:]
    assert got_wl == gzip.open(os.path.join(G, name + ".short300.wl-new.dat.gz"), "rt").read()


@pytest.mark.parametrize("name", ["test_01_clustermoves", "test_01_grandcanonical"])
def test_reference_main_cluster_and_grand_canonical_moves(name):
    """the moves SURVEY.md section 8(f) rank 4 lists stay the reference's own host code (MoveCreator::clusterMoveGeom,
    movecreator.cpp:54-172; muVTMove, :797-924) and run unchanged on the GPU calculator: allToAll / p2p for the cluster moves,
    update(EMResize) -> initEM() when an insertion or deletion changes the particle count. Byte-identical config.last after 300 sweeps
    of Tests/test_01_normal_PSC with nClustMove = 10, resp. nGrandCanon = 5 and an activity (tests/golden/make_golden.py othermoves).
    The cluster-move golden is written by the reference built with its own TotalEFull<PairE> calculator (oracle/_ref/SC_full): the cluster
    move writes conf->pvec behind the calculator's back (movecreator.cpp:164) and the default TotalEMatrix carries on with a stale
    matrix; TotalEGpu::allToAll() re-reads the host configuration, as TotalEFull does."""
    got, out = run_reference_program(name, 0)
    assert got == open(os.path.join(G, name + ".short300.config.last")).read(), out[-800:]
