"""CPU: the C-ABI libraries load and export every symbol include/scgpu.h declares; record sizes match the header.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re

import pytest

from sc_b200 import engine
from sc_b200.build import lib_path, HOST_LIB

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "scgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(scgpu_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("variant", ["fast", "strict"])
def test_every_declared_symbol_is_exported(variant):
    syms = declared_symbols()
    assert len(syms) >= 25
    assert set(syms) == set(engine.SYMBOLS)
    L = ctypes.CDLL(lib_path(variant))
    for s in syms:
        assert hasattr(L, s), s


def test_record_sizes_match_header():
    txt = open(os.path.join(ROOT, "include", "scgpu.h")).read()
    assert "#define SCGPU_STATE_DOUBLES 30" in txt and "#define SCGPU_IAPARAM_DOUBLES 48" in txt
    assert ctypes.sizeof(engine.MoveParams) == 8 + 40 * 8 * 2 + 16      # temper, trans_mx, rot_angle, {n_sub, grid_k, trial_rule, reserved}
    assert ctypes.sizeof(engine.SweepStats) == 6 * 8
    assert ctypes.sizeof(engine.ChainMoves) == 8 + 32 * 8 * 2 and ctypes.sizeof(engine.ChainStats) == 7 * 8
    assert ctypes.sizeof(engine.PressureParams) == 3 * 8 + 8 and ctypes.sizeof(engine.PressureStats) == 8 + 6 * 8
    assert ctypes.sizeof(engine.WlOrder) == 16 + 4 * 8 + 8 + 2 * 8 + 2 * 8 + 4 * 8 + 8 + 2 * 8      # scgpu_wlorder


def test_host_library_loads():
    L = ctypes.CDLL(HOST_LIB)
    for s in ("schost_load_text", "schost_export", "schost_calc_create", "schost_calc_one_to_all_trial", "schost_calc_all_to_all"):
        assert hasattr(L, s)


def test_no_gpu_means_loud_failure():
    """there is no CPU fallback: without a device, creating a context must raise"""
    L = engine.load_library("fast")
    if L.scgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.ScgpuError):
        engine.Engine(0, "fast")
