"""ctypes binding of the C++ host mirror (sc_b200/csrc/host): scOOP's top.init / config.init semantics and the
TotalEGpu calculator class with the reference's method names. Product code: no oracle involved."""
import ctypes as C
import os

import numpy as np

from .build import HOST_LIB

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


class HostError(RuntimeError):
    pass


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB):
            raise HostError("host library %s is missing: run `python -m sc_b200.build`" % HOST_LIB)
        L = C.CDLL(HOST_LIB)
        L.schost_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.schost_load_text.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_long), C.c_int, C.POINTER(vp)]
        L.schost_free.argtypes = [vp]
        L.schost_dims.argtypes = [vp, _ip, _ip, _ip]
        L.schost_export.argtypes = [vp, _dp, _ip, _ip, _dp, _dp, _dp, _dp]
        L.schost_set_state.argtypes = [vp, C.c_int, _dp]
        L.schost_exter.argtypes = [vp, _dp]
        L.schost_set_box.argtypes = [vp, _dp]
        L.schost_calc_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.schost_calc_free.argtypes = [vp]
        for name in ("schost_calc_all_to_all", "schost_calc_all_to_all_trial"):
            getattr(L, name).argtypes = [vp, _dp]
        for name in ("schost_calc_one_to_all", "schost_calc_one_to_all_trial"):
            getattr(L, name).argtypes = [vp, C.c_int, _dp]
        L.schost_calc_p2p.argtypes = [vp, C.c_int, C.c_int, _dp]
        L.schost_calc_mol2others.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp]
        L.schost_calc_update_particle.argtypes = [vp, C.c_int]
        L.schost_calc_update_box.argtypes = [vp]
        L.schost_calc_ctx.argtypes = [vp]
        L.schost_calc_ctx.restype = vp
        L.schost_mc_last_error.restype = C.c_char_p
        L.schost_run_mc.argtypes = [vp, C.c_char_p, C.c_int, C.c_long, _dp]
        L.schost_config_last.argtypes = [vp, C.c_int, C.c_char_p, C.c_long]
        L.schost_config_last.restype = C.c_long
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise HostError(_load().schost_last_error().decode())


class HostSystem:
    """A configuration loaded by the C++ host mirror. Attribute names match what Engine.load() expects."""

    def __init__(self, top_text, config_text, counts=None):
        L = _load()
        self.h = C.c_void_p()
        if counts is not None:
            arr = (C.c_long * len(counts))(*counts)
            _ck(L.schost_load_text(top_text.encode(), config_text.encode(), arr, len(counts), C.byref(self.h)))
        else:
            _ck(L.schost_load_text(top_text.encode(), config_text.encode(), None, 0, C.byref(self.h)))
        n, T, M = C.c_int(), C.c_int(), C.c_int()
        L.schost_dims(self.h, C.byref(n), C.byref(T), C.byref(M))
        self.n, self.ntypes, self.nmol = n.value, T.value, M.value
        self.state = np.zeros((self.n, 30))
        self.type = np.zeros(self.n, dtype=np.int32)
        self.moltype = np.zeros(self.n, dtype=np.int32)
        self.ia = np.zeros((self.ntypes, self.ntypes, 48))
        self.mol = np.zeros((self.nmol, 16))
        self.box = np.zeros(3)
        cut = np.zeros(2)
        L.schost_export(self.h, self.state.ctypes.data_as(_dp), self.type.ctypes.data_as(_ip), self.moltype.ctypes.data_as(_ip),
                        self.ia.ctypes.data_as(_dp), self.mol.ctypes.data_as(_dp), self.box.ctypes.data_as(_dp), cut.ctypes.data_as(_dp))
        self.sqmaxcut, self.maxcut = float(cut[0]), float(cut[1])
        ex = np.zeros(4)
        L.schost_exter(self.h, ex.ctypes.data_as(_dp))
        self.exter = [float(ex[1]), float(ex[2]), float(ex[3])] if ex[0] else None      # [EXTER]: thickness, epsilon, attraction switch

    def set_state(self, idx, state):
        st = np.ascontiguousarray(state, dtype=np.float64)
        _ck(_load().schost_set_state(self.h, int(idx), st.ctypes.data_as(_dp)))
        self.state[idx] = st

    def set_box(self, box):
        b = np.ascontiguousarray(box, dtype=np.float64)
        _ck(_load().schost_set_box(self.h, b.ctypes.data_as(_dp)))
        self.box[:] = b

    def refresh(self):
        """re-read the arrays after the C++ side changed the configuration (sequential MC driver)"""
        cut = np.zeros(2)
        _load().schost_export(self.h, self.state.ctypes.data_as(_dp), self.type.ctypes.data_as(_ip), self.moltype.ctypes.data_as(_ip),
                              self.ia.ctypes.data_as(_dp), self.mol.ctypes.data_as(_dp), self.box.ctypes.data_as(_dp), cut.ctypes.data_as(_dp))

    def run_mc(self, options_text, device=0, nsweeps=0):
        """production run of the reference program (Updater::simulate) through TotalEGpu; returns a stats dict"""
        L = _load()
        out = np.zeros(13)
        if L.schost_run_mc(self.h, options_text.encode(), int(device), int(nsweeps), out.ctypes.data_as(_dp)) != 0:
            raise HostError(L.schost_mc_last_error().decode())
        self.refresh()
        keys = ["trans_acc", "trans_rej", "rot_acc", "rot_rej", "chainm_acc", "chainm_rej", "chainr_acc", "chainr_rej", "edge_acc", "edge_rej",
                "e_start", "e_end", "drift"]
        return dict(zip(keys, out.tolist()))

    def config_last(self, testing_format=True):
        L = _load()
        n = L.schost_config_last(self.h, 1 if testing_format else 0, None, 0)
        buf = C.create_string_buffer(n + 1)
        L.schost_config_last(self.h, 1 if testing_format else 0, buf, n)
        return buf.raw[:n].decode()

    def close(self):
        if self.h:
            _load().schost_free(self.h)
            self.h = C.c_void_p()


class Calculator:
    """TotalEGpu through its C entry points; method names are the reference's (totalenergycalculator.h:135-297)."""

    def __init__(self, hsys, device=0):
        self.sys = hsys
        self.h = C.c_void_p()
        _ck(_load().schost_calc_create(hsys.h, int(device), C.byref(self.h)))

    def _d(self, fn, *args):
        out = C.c_double(0.0)
        _ck(fn(self.h, *args, C.byref(out)))
        return out.value

    def allToAll(self):
        return self._d(_load().schost_calc_all_to_all)

    def allToAllTrial(self):
        return self._d(_load().schost_calc_all_to_all_trial)

    def oneToAll(self, target):
        return self._d(_load().schost_calc_one_to_all, int(target))

    def oneToAllTrial(self, target):
        return self._d(_load().schost_calc_one_to_all_trial, int(target))

    def p2p(self, a, b):
        return self._d(_load().schost_calc_p2p, int(a), int(b))

    def mol2others(self, first, m):
        return self._d(_load().schost_calc_mol2others, int(first), int(m), 0)

    def mol2othersTrial(self, first, m):
        return self._d(_load().schost_calc_mol2others, int(first), int(m), 1)

    def update(self, target=None):
        if target is None:
            _ck(_load().schost_calc_update_box(self.h))
        else:
            _ck(_load().schost_calc_update_particle(self.h, int(target)))

    def close(self):
        if self.h:
            _load().schost_calc_free(self.h)
            self.h = C.c_void_p()
