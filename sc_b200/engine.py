"""ctypes binding of the C ABI (include/scgpu.h). Mirrors the calculator interface of the reference
(TotalE<>, scOOP/mc/totalenergycalculator.h:135-297): one_to_all / one_to_all_trial / all_to_all /
mol_to_others / update. Fails loudly when the CUDA library or a GPU is missing -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np

from .build import lib_path

STATE = 30
IA = 48
MOL = 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)


class ScgpuError(RuntimeError):
    pass


class MoveParams(C.Structure):
    _fields_ = [("temper", C.c_double), ("trans_mx", C.c_double * 40), ("rot_angle", C.c_double * 40),
                ("n_sub", C.c_int), ("grid_k", C.c_int), ("trial_rule", C.c_int), ("reserved", C.c_int)]


class SweepStats(C.Structure):
    _fields_ = [("trans_acc", C.c_int64), ("trans_rej", C.c_int64), ("rot_acc", C.c_int64), ("rot_rej", C.c_int64),
                ("cell_rej", C.c_int64), ("energy_delta", C.c_double)]


class ChainMoves(C.Structure):
    _fields_ = [("chainprob", C.c_double), ("chainm_mx", C.c_double * 32), ("chainr_angle", C.c_double * 32)]


class ChainStats(C.Structure):
    _fields_ = [("chainm_acc", C.c_int64), ("chainm_rej", C.c_int64), ("chainr_acc", C.c_int64), ("chainr_rej", C.c_int64),
                ("cell_rej", C.c_int64), ("energy_delta", C.c_double), ("noop", C.c_int64)]


class PressureParams(C.Structure):
    _fields_ = [("temper", C.c_double), ("press", C.c_double), ("edge_mx", C.c_double), ("ptype", C.c_int), ("reserved", C.c_int)]


class PressureStats(C.Structure):
    _fields_ = [("accepted", C.c_int), ("reserved", C.c_int), ("energy_old", C.c_double), ("energy_new", C.c_double),
                ("enthalpy_delta", C.c_double), ("box", C.c_double * 3)]


class WlOrder(C.Structure):
    """scgpu_wlorder (include/scgpu.h): Wang-Landau order parameters of the whole configuration, in / out"""
    _fields_ = [("wlm", C.c_int * 2), ("wlmtype", C.c_int), ("reserved", C.c_int), ("minorder", C.c_double * 2), ("dorder", C.c_double * 2),
                ("meshsize", C.c_double), ("order", C.c_int64 * 2), ("raw", C.c_double * 2), ("syscm", C.c_double * 3), ("sysvolume", C.c_double),
                ("mesh_dim", C.c_int * 2), ("mesh_occupied", C.c_int64), ("mesh_skipped", C.c_int64)]


class ReplicaState(C.Structure):
    """scgpu_replica_state (include/scgpu.h): what a replica owns and what travels on an accepted exchange"""
    _fields_ = [("temper", C.c_double), ("press", C.c_double), ("pseudo_rank", C.c_int), ("replica", C.c_int),
                ("wl_order", C.c_int64 * 2), ("part_num", C.c_double * 8), ("payload", C.c_double * 40),
                ("attempted", C.c_int), ("accepted", C.c_int), ("partner", C.c_int), ("reserved", C.c_int),
                ("partner_wl_order", C.c_int64 * 2), ("change", C.c_double), ("energy", C.c_double), ("volume", C.c_double),
                ("edrift", C.c_double)]


class ExchangeParams(C.Structure):
    _fields_ = [("nrepchange", C.c_int), ("wl_len", C.c_int), ("wl_len0", C.c_int64), ("dtemp", C.c_double), ("dpress", C.c_double),
                ("chempot", C.c_double * 8), ("seed", C.c_uint64)]


class WLState(C.Structure):
    _fields_ = [("alpha", C.c_double), ("wmin", C.c_double), ("min", C.c_int64), ("max", C.c_int64), ("halved", C.c_int), ("converged", C.c_int)]


SYMBOLS = ["scgpu_last_error", "scgpu_device_count", "scgpu_create", "scgpu_destroy", "scgpu_set_topology",
           "scgpu_set_particles", "scgpu_set_particles_compact", "scgpu_set_exter", "scgpu_set_box", "scgpu_update_particle", "scgpu_set_particle_type", "scgpu_download_particles",
           "scgpu_build_cells", "scgpu_cell_assignment", "scgpu_cell_order", "scgpu_one_to_all",
           "scgpu_one_to_all_batch", "scgpu_one_to_all_everyone", "scgpu_submit_everyone", "scgpu_mol_to_others", "scgpu_all_to_all",
           "scgpu_overlap_one", "scgpu_overlap_all", "scgpu_sweep_checkerboard", "scgpu_sweep_checkerboard_chains", "scgpu_pressure_move",
           "scgpu_comm_unique_id", "scgpu_comm_create", "scgpu_comm_attach", "scgpu_comm_destroy", "scgpu_replica_exchange",
           "scgpu_comm_last_exchange_us", "scgpu_wl_merge", "scgpu_wl_order", "scgpu_wl_mesh",
           "scgpu_timer_start", "scgpu_timer_stop", "scgpu_sync", "scgpu_fp64_peak", "scgpu_profile_everyone", "scgpu_flush_l2",
           "scgpu_kernel_launches"]

_libs = {}


def load_library(variant="fast"):
    """dlopen the in-tree CUDA library; raises if it has not been built (never falls back to anything)."""
    if variant in _libs:
        return _libs[variant]
    path = lib_path(variant)
    if not os.path.exists(path):
        raise ScgpuError("CUDA library %s is missing: run `python -m sc_b200.build` (no CPU fallback exists)" % path)
    L = C.CDLL(path)
    L.scgpu_last_error.restype = C.c_char_p
    vp = C.c_void_p
    L.scgpu_create.argtypes = [C.POINTER(vp), C.c_int]
    L.scgpu_destroy.argtypes = [vp]
    L.scgpu_set_topology.argtypes = [vp, C.c_int, _dp, C.c_double, C.c_double, C.c_int, _dp]
    L.scgpu_set_particles.argtypes = [vp, C.c_int, _dp, _ip, _ip]
    L.scgpu_set_particles_compact.argtypes = [vp, C.c_int, _dp, _ip, _ip]
    L.scgpu_set_box.argtypes = [vp, _dp]
    L.scgpu_set_exter.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.scgpu_update_particle.argtypes = [vp, C.c_int, _dp]
    L.scgpu_set_particle_type.argtypes = [vp, C.c_int, C.c_int]
    L.scgpu_download_particles.argtypes = [vp, _dp]
    L.scgpu_build_cells.argtypes = [vp]
    L.scgpu_cell_assignment.argtypes = [vp, _ip, _ip]
    L.scgpu_cell_order.argtypes = [vp, _ip, _ip]
    L.scgpu_one_to_all.argtypes = [vp, C.c_int, _dp, _dp, _dp]
    L.scgpu_one_to_all_batch.argtypes = [vp, C.c_int, _ip, _dp, _dp]
    L.scgpu_one_to_all_everyone.argtypes = [vp, _dp, _i64p, _i64p]
    L.scgpu_submit_everyone.argtypes = [vp, _dp, _dp]
    L.scgpu_mol_to_others.argtypes = [vp, C.c_int, C.c_int, _dp, _dp]
    L.scgpu_all_to_all.argtypes = [vp, _dp, _dp]
    L.scgpu_overlap_one.argtypes = [vp, C.c_int, _dp, C.c_int, _ip]
    L.scgpu_overlap_all.argtypes = [vp, C.c_int, _ip]
    L.scgpu_sweep_checkerboard.argtypes = [vp, C.POINTER(MoveParams), C.c_uint64, C.c_uint64, C.POINTER(SweepStats)]
    L.scgpu_pressure_move.argtypes = [vp, C.POINTER(PressureParams), C.c_uint64, C.c_uint64, C.POINTER(PressureStats)]
    L.scgpu_wl_order.argtypes = [vp, C.POINTER(WlOrder)]
    L.scgpu_wl_mesh.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
    L.scgpu_sweep_checkerboard_chains.argtypes = [vp, C.POINTER(MoveParams), C.POINTER(ChainMoves), C.c_uint64, C.c_uint64,
                                                  C.POINTER(SweepStats), C.POINTER(ChainStats)]
    L.scgpu_comm_unique_id.argtypes = [C.c_char_p]
    L.scgpu_comm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_char_p]
    L.scgpu_comm_attach.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int]
    L.scgpu_comm_destroy.argtypes = [vp]
    L.scgpu_replica_exchange.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(ReplicaState), C.POINTER(ExchangeParams), C.c_uint64, C.POINTER(_dp)]
    L.scgpu_comm_last_exchange_us.argtypes = [vp, C.POINTER(C.c_float)]
    L.scgpu_wl_merge.argtypes = [vp, C.c_int, _dp, _i64p, _dp, _i64p, C.c_double, C.POINTER(WLState)]
    L.scgpu_timer_start.argtypes = [vp]
    L.scgpu_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.scgpu_sync.argtypes = [vp]
    L.scgpu_fp64_peak.argtypes = [vp, _dp]
    L.scgpu_flush_l2.argtypes = [vp]
    L.scgpu_profile_everyone.argtypes = [vp, C.POINTER(C.c_float)]
    L.scgpu_kernel_launches.argtypes = [vp, _i64p]
    _libs[variant] = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Engine:
    """One context = one GPU = one stream. Arrays follow the C ABI: state[n,30], ia[T,T,48], mol[M,16]."""

    def __init__(self, device=0, variant="fast"):
        self.L = load_library(variant)
        self.variant = variant
        self.h = C.c_void_p()
        self._ck(self.L.scgpu_create(C.byref(self.h), int(device)))
        self.n = 0

    def _ck(self, rc):
        if rc != 0:
            raise ScgpuError("scgpu error %d: %s" % (rc, self.L.scgpu_last_error().decode()))

    def close(self):
        if self.h:
            self.L.scgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_topology(self, ia, mol, sqmaxcut, maxcut):
        ia = np.ascontiguousarray(ia, dtype=np.float64)
        mol = np.ascontiguousarray(mol, dtype=np.float64).reshape(-1, MOL)
        T = ia.shape[0]
        assert ia.shape == (T, T, IA)
        self._ck(self.L.scgpu_set_topology(self.h, T, _d(ia), float(sqmaxcut), float(maxcut), mol.shape[0], _d(mol)))

    def set_particles(self, state, types, moltypes):
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, STATE)
        types = np.ascontiguousarray(types, dtype=np.int32)
        moltypes = np.ascontiguousarray(moltypes, dtype=np.int32)
        self.n = state.shape[0]
        self._ck(self.L.scgpu_set_particles(self.h, self.n, _d(state), _i(types), _i(moltypes)))

    def set_particles_compact(self, state9, types, moltypes):
        """state9[n,9] = pos (box-fractional), dir, patchdir as config.init holds them; the rest is derived on the device"""
        state9 = np.ascontiguousarray(state9, dtype=np.float64).reshape(-1, 9)
        self.n = state9.shape[0]
        if types is None and moltypes is None:       # same particle count and types as the previous upload: only coordinates travel
            self._ck(self.L.scgpu_set_particles_compact(self.h, self.n, _d(state9), None, None))
            return
        types = np.ascontiguousarray(types, dtype=np.int32)
        moltypes = np.ascontiguousarray(moltypes, dtype=np.int32)
        self._ck(self.L.scgpu_set_particles_compact(self.h, self.n, _d(state9), _i(types), _i(moltypes)))

    def set_box(self, box):
        box = np.ascontiguousarray(box, dtype=np.float64)
        self._ck(self.L.scgpu_set_box(self.h, _d(box)))

    def set_exter(self, exter):
        """exter: None, or (thickness, epsilon, attraction switch) of the [EXTER] wall; after set_topology"""
        if exter is None:
            self._ck(self.L.scgpu_set_exter(self.h, 0, 0.0, 0.0, 0.0))
        else:
            self._ck(self.L.scgpu_set_exter(self.h, 1, float(exter[0]), float(exter[1]), float(exter[2])))

    def load(self, sysobj):
        """sysobj: anything with .ia .mol .sqmaxcut .maxcut .state .type .moltype .box (and optionally .exter)"""
        self.set_topology(sysobj.ia, sysobj.mol, sysobj.sqmaxcut, sysobj.maxcut)
        self.set_exter(getattr(sysobj, "exter", None))
        self.set_box(sysobj.box)
        self.set_particles(sysobj.state, sysobj.type, sysobj.moltype)
        return self

    def update_particle(self, idx, state):
        st = np.ascontiguousarray(state, dtype=np.float64)
        self._ck(self.L.scgpu_update_particle(self.h, int(idx), _d(st)))

    def set_particle_type(self, idx, ptype):
        """switchTypeMove: the type of one particle changes on the device"""
        self._ck(self.L.scgpu_set_particle_type(self.h, int(idx), int(ptype)))

    def download_particles(self):
        out = np.zeros((self.n, STATE))
        self._ck(self.L.scgpu_download_particles(self.h, _d(out)))
        return out

    # ---- cells
    def build_cells(self):
        self._ck(self.L.scgpu_build_cells(self.h))

    def cell_assignment(self):
        cell = np.zeros(self.n, dtype=np.int32)
        nc = np.zeros(3, dtype=np.int32)
        self._ck(self.L.scgpu_cell_assignment(self.h, _i(cell), _i(nc)))
        return cell, nc

    def cell_order(self, ncells):
        order = np.zeros(self.n, dtype=np.int32)
        start = np.zeros(ncells + 1, dtype=np.int32)
        self._ck(self.L.scgpu_cell_order(self.h, _i(order), _i(start)))
        return order, start

    # ---- energies (names of the reference's calculator)
    def one_to_all(self, target, trial_state=None, pairs=False):
        e = C.c_double(0.0)
        ep = np.zeros(self.n) if pairs else None
        ts = None if trial_state is None else np.ascontiguousarray(trial_state, dtype=np.float64)
        self._ck(self.L.scgpu_one_to_all(self.h, int(target), None if ts is None else _d(ts), C.byref(e),
                                         None if ep is None else _d(ep)))
        return (e.value, ep) if pairs else e.value

    one_to_all_trial = one_to_all

    def one_to_all_batch(self, targets, trial_states=None):
        targets = np.ascontiguousarray(targets, dtype=np.int32)
        out = np.zeros(len(targets))
        ts = None if trial_states is None else np.ascontiguousarray(trial_states, dtype=np.float64)
        self._ck(self.L.scgpu_one_to_all_batch(self.h, len(targets), _i(targets), None if ts is None else _d(ts), _d(out)))
        return out

    def one_to_all_everyone(self, fetch=True, count=False, out=None):
        """out: optional caller-owned float64[n] result buffer (page-locked memory is read back by DMA without staging)"""
        if out is None:
            out = np.zeros(self.n) if fetch else None
        nc, ng = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.scgpu_one_to_all_everyone(self.h, None if out is None else _d(out),
                                                  C.byref(nc) if count else None, C.byref(ng) if count else None))
        if count:
            return out, nc.value, ng.value
        return out

    def submit_everyone(self, state9_pinned, out_pinned):
        """asynchronous: upload a configuration (page-locked float64[n,9]), rebuild cells, oneToAll of every particle, energies into
        the page-locked float64[n] buffer; valid after sync(). The caller keeps both arrays alive and untouched until then."""
        assert state9_pinned.dtype == np.float64 and state9_pinned.flags.c_contiguous and state9_pinned.size == 9 * self.n
        assert out_pinned.dtype == np.float64 and out_pinned.flags.c_contiguous and out_pinned.size == self.n
        self._ck(self.L.scgpu_submit_everyone(self.h, _d(state9_pinned), _d(out_pinned)))

    def mol_to_others(self, first, m, trial_states=None):
        e = C.c_double(0.0)
        ts = None if trial_states is None else np.ascontiguousarray(trial_states, dtype=np.float64)
        self._ck(self.L.scgpu_mol_to_others(self.h, int(first), int(m), None if ts is None else _d(ts), C.byref(e)))
        return e.value

    def all_to_all(self, rows=False, fetch=True):
        e = C.c_double(0.0)
        er = np.zeros(self.n) if rows else None
        self._ck(self.L.scgpu_all_to_all(self.h, C.byref(e) if fetch else None, None if er is None else _d(er)))
        return (e.value, er) if rows else e.value

    def overlap_one(self, target, trial_state=None, variant=0):
        f = C.c_int(0)
        ts = None if trial_state is None else np.ascontiguousarray(trial_state, dtype=np.float64)
        self._ck(self.L.scgpu_overlap_one(self.h, int(target), None if ts is None else _d(ts), int(variant), C.byref(f)))
        return f.value

    def overlap_all(self, variant=0):
        f = C.c_int(0)
        self._ck(self.L.scgpu_overlap_all(self.h, int(variant), C.byref(f)))
        return f.value

    def sweep(self, mp, seed, sweep, stats=True):
        """one checkerboard sweep; stats=False enqueues it without any read-back (several engines can then overlap on one GPU)"""
        if not stats:
            self._ck(self.L.scgpu_sweep_checkerboard(self.h, C.byref(mp), int(seed), int(sweep), None))
            return None
        st = SweepStats()
        self._ck(self.L.scgpu_sweep_checkerboard(self.h, C.byref(mp), int(seed), int(sweep), C.byref(st)))
        return st

    def sweep_chains(self, mp, cm, seed, sweep, stats=True):
        """one checkerboard sweep in which a share cm.chainprob of the trials are chain moves -> (SweepStats, ChainStats)"""
        if not stats:
            self._ck(self.L.scgpu_sweep_checkerboard_chains(self.h, C.byref(mp), C.byref(cm), int(seed), int(sweep), None, None))
            return None
        st, cst = SweepStats(), ChainStats()
        self._ck(self.L.scgpu_sweep_checkerboard_chains(self.h, C.byref(mp), C.byref(cm), int(seed), int(sweep), C.byref(st), C.byref(cst)))
        return st, cst

    def pressure_move(self, ptype, press, temper, edge_mx, seed, step):
        """one volume move (pressureMove, ptype 0-3); edge_mx as stat.edge.mx (= 2 * the options file's edge_mx) -> PressureStats"""
        pp = PressureParams()
        pp.temper, pp.press, pp.edge_mx, pp.ptype = float(temper), float(press), float(edge_mx), int(ptype)
        st = PressureStats()
        self._ck(self.L.scgpu_pressure_move(self.h, C.byref(pp), int(seed), int(step), C.byref(st)))
        return st

    def wl_order(self, wlm, wlmtype=0, minorder=(0.0, 0.0), dorder=(1.0, 1.0), meshsize=0.0):
        """Wang-Landau order parameters of the configuration on the device (WangLandau::init / runPress forms, scOOP/mc/wanglandau.h:168-196,
        wanglandau.cpp:56-125); wlm = one method or a pair (second 0 = unused) -> WlOrder (order, raw, syscm, mesh_dim, ...)"""
        w = WlOrder()
        wl = (int(wlm), 0) if np.isscalar(wlm) else tuple(int(x) for x in wlm)
        mn = (float(minorder), 0.0) if np.isscalar(minorder) else tuple(float(x) for x in minorder)
        dd = (float(dorder), 1.0) if np.isscalar(dorder) else tuple(float(x) for x in dorder)
        w.wlm[0], w.wlm[1] = wl
        w.minorder[0], w.minorder[1] = mn
        w.dorder[0], w.dorder[1] = dd
        w.wlmtype, w.meshsize = int(wlmtype), float(meshsize)
        self._ck(self.L.scgpu_wl_order(self.h, C.byref(w)))
        return w

    def wl_mesh(self, dim):
        """Mesh::data of the last wl_order call with wlm 2 (occupied: -count, free: hole number) as an int32 array [dim1][dim0]"""
        out = np.zeros(int(dim[0]) * int(dim[1]), dtype=np.int32)
        self._ck(self.L.scgpu_wl_mesh(self.h, out.ctypes.data_as(C.POINTER(C.c_int)), out.size))
        return out.reshape(int(dim[1]), int(dim[0]))

    # ---- measurement
    def timer_start(self):
        self._ck(self.L.scgpu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0.0)
        self._ck(self.L.scgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def sync(self):
        self._ck(self.L.scgpu_sync(self.h))

    def fp64_peak(self):
        t = C.c_double(0.0)
        self._ck(self.L.scgpu_fp64_peak(self.h, C.byref(t)))
        return t.value

    def profile_everyone(self):
        """microseconds of the four launches of one every-particle pass: gate, cheap terms, patch terms, combine"""
        us = (C.c_float * 4)()
        self._ck(self.L.scgpu_profile_everyone(self.h, us))
        return [float(x) for x in us]

    def flush_l2(self):
        self._ck(self.L.scgpu_flush_l2(self.h))

    def launches(self):
        v = C.c_int64(0)
        self._ck(self.L.scgpu_kernel_launches(self.h, C.byref(v)))
        return v.value


class Comm:
    """scgpu_comm: the replicas / walkers of `nranks` processes (one per GPU), NCCL underneath when nranks > 1.
    unique_id: the 128 bytes of Comm.unique_id() drawn on rank 0 and handed to every rank by the host program."""

    def __init__(self, device=0, nranks=1, rank=0, unique_id=None, variant="fast"):
        self.L = load_library(variant)
        self.h = C.c_void_p()
        self.nranks, self.rank = int(nranks), int(rank)
        rc = self.L.scgpu_comm_create(C.byref(self.h), int(device), self.nranks, self.rank, unique_id)
        if rc != 0:
            raise ScgpuError("scgpu error %d: %s" % (rc, self.L.scgpu_last_error().decode()))

    @staticmethod
    def unique_id(variant="fast"):
        L = load_library(variant)
        buf = C.create_string_buffer(128)
        rc = L.scgpu_comm_unique_id(buf)
        if rc != 0:
            raise ScgpuError("scgpu error %d: %s" % (rc, L.scgpu_last_error().decode()))
        return buf.raw

    def _ck(self, rc):
        if rc != 0:
            raise ScgpuError("scgpu error %d: %s" % (rc, self.L.scgpu_last_error().decode()))

    def close(self):
        if self.h:
            self.L.scgpu_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exchange(self, engines, states, params, sweep, wl_weights=None):
        """MoveCreator::replicaExchangeMove for this process's replicas. engines: list of Engine, states: ctypes array of
        ReplicaState (updated in place), params: ExchangeParams, wl_weights: None or a list of float64 arrays."""
        n = len(engines)
        hs = (C.c_void_p * n)(*[e.h for e in engines])
        wl = None
        if wl_weights is not None:
            wl = (_dp * n)(*[_d(w) for w in wl_weights])
        self._ck(self.L.scgpu_replica_exchange(self.h, n, hs, states, C.byref(params), int(sweep), wl))

    def last_exchange_us(self):
        v = C.c_float(0.0)
        self._ck(self.L.scgpu_comm_last_exchange_us(self.h, C.byref(v)))
        return v.value

    def wl_merge(self, weights, hist, weights_base, hist_base, temper, alpha):
        """merge this walker's Wang-Landau arrays with everybody's (all float64 / int64 numpy arrays, updated in place) -> WLState"""
        st = WLState()
        st.alpha = float(alpha)
        for a in (weights, weights_base):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        for a in (hist, hist_base):
            assert a.dtype == np.int64 and a.flags.c_contiguous
        self._ck(self.L.scgpu_wl_merge(self.h, len(weights), _d(weights), hist.ctypes.data_as(_i64p), _d(weights_base),
                                       hist_base.ctypes.data_as(_i64p), float(temper), C.byref(st)))
        return st
