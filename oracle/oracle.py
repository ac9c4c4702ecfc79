"""TEST INFRASTRUCTURE -- ctypes front end of the C oracle (oracle/libscoracle.so) and the parser of
reference-driver dumps (tests/golden/*.ref.gz). NOT part of the product: only tests/, bench.py's CPU
legs (cpu_baseline / --impl reference) and __graft_entry__.smoke() may import this module.
"""
import ctypes as C
import gzip
import math
import os
import subprocess

import numpy as np

from . import topo as otopo

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscoracle.so")
STATE = 30
IA_FIELDS = 48
MOL_FIELDS = 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _SysC(C.Structure):
    _fields_ = [("n", C.c_int), ("ntypes", C.c_int), ("nmoltypes", C.c_int),
                ("state", _dp), ("type", _ip), ("moltype", _ip), ("ia", _dp), ("mol", _dp),
                ("box", C.c_double * 3), ("sqmaxcut", C.c_double), ("maxcut", C.c_double)]


class _ConC(C.Structure):
    _fields_ = [("is_empty", C.c_int), ("con", C.c_int * 4), ("sp", C.c_double), ("mod", C.c_double * 2),
                ("c", C.c_double * 2), ("eq", C.c_double * 2)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.sco_extere2.restype = C.c_double
        L.sco_extere2.argtypes = [_dp, C.c_int, _dp, C.c_double, C.c_double]
        L.sco_exter_params.argtypes = [_dp, C.c_double, C.c_double, C.c_double, _dp]
        L.sco_pair_energy.restype = C.c_double
        L.sco_pair_energy.argtypes = [C.POINTER(_SysC), _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.POINTER(_ConC)]
        L.sco_one_to_all.restype = C.c_double
        L.sco_one_to_all.argtypes = [C.POINTER(_SysC), C.c_int, _dp, _dp]
        L.sco_all_to_all.restype = C.c_double
        L.sco_all_to_all.argtypes = [C.POINTER(_SysC), _dp]
        L.sco_mol_to_others.restype = C.c_double
        L.sco_mol_to_others.argtypes = [C.POINTER(_SysC), C.c_int, C.c_int]
        L.sco_get_conlist.argtypes = [C.POINTER(_SysC), C.c_int, C.POINTER(_ConC)]
        L.sco_particle_init.argtypes = [_dp, _dp]
        L.sco_overlap_pair.restype = C.c_int
        L.sco_overlap_pair.argtypes = [C.POINTER(_SysC), _dp, C.c_int, _dp, C.c_int, C.c_int]
        L.sco_overlap_one.restype = C.c_int
        L.sco_overlap_one.argtypes = [C.POINTER(_SysC), C.c_int, _dp, C.c_int]
        L.sco_overlap_all.restype = C.c_int
        L.sco_overlap_all.argtypes = [C.POINTER(_SysC), C.c_int]
        L.sco_cell_dims.argtypes = [C.POINTER(_SysC), _ip]
        L.sco_cell_assign.argtypes = [C.POINTER(_SysC), _ip, _ip]
        L.sco_cell_sort.argtypes = [C.POINTER(_SysC), _ip, C.c_int, _ip, _ip]
        L.sco_one_to_all_cells.restype = C.c_double
        L.sco_one_to_all_cells.argtypes = [C.POINTER(_SysC), C.c_int, _dp, _ip, _ip, _ip, _ip,
                                           C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.sco_psc_rotate.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_int]
        L.sco_min_dist_segments.argtypes = [_dp, _dp, C.c_double, C.c_double, _dp, _dp]
        L.sco_image.argtypes = [_dp, _dp, _dp, _dp]
        # Wang-Landau order parameters (oracle/wl_order.c)
        L.sco_wl_mass_center.argtypes = [C.c_int, _dp, _ip, _dp, _dp]
        L.sco_wl_z.restype = C.c_double
        L.sco_wl_z.argtypes = [_dp, _dp, _dp]
        L.sco_wl_two_part_dist.restype = C.c_double
        L.sco_wl_two_part_dist.argtypes = [_dp, _dp]
        L.sco_wl_contacts.restype = C.c_long
        L.sco_wl_contacts.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp]
        L.sco_wl_bin.restype = C.c_long
        L.sco_wl_bin.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.sco_wl_mesh_hole.restype = C.c_int
        L.sco_wl_mesh_hole.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, C.c_double, _ip, _ip, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.sco_wl_mesh_labels.restype = C.c_int
        L.sco_wl_mesh_labels.argtypes = [C.c_int, _dp, _ip, C.c_int, _dp, C.c_double, _ip, _ip]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class System:
    """Packed system: the same buffers go to the oracle and (through the C ABI) to the product."""

    def __init__(self, state, types, moltypes, ia, mol, box, sqmaxcut, maxcut):
        self.state = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, STATE)
        self.type = np.ascontiguousarray(types, dtype=np.int32)
        self.moltype = np.ascontiguousarray(moltypes, dtype=np.int32)
        self.ia = np.ascontiguousarray(ia, dtype=np.float64)
        self.mol = np.ascontiguousarray(mol, dtype=np.float64).reshape(-1, MOL_FIELDS)
        self.box = np.ascontiguousarray(box, dtype=np.float64)
        self.sqmaxcut = float(sqmaxcut)
        self.maxcut = float(maxcut)
        self.n = self.state.shape[0]
        self.ntypes = self.ia.shape[0]
        assert self.ia.shape == (self.ntypes, self.ntypes, IA_FIELDS)

    def c(self):
        s = _SysC()
        s.n, s.ntypes, s.nmoltypes = self.n, self.ntypes, self.mol.shape[0]
        s.state, s.type, s.moltype = _d(self.state), _i(self.type), _i(self.moltype)
        s.ia, s.mol = _d(self.ia), _d(self.mol)
        s.box[0], s.box[1], s.box[2] = self.box
        s.sqmaxcut, s.maxcut = self.sqmaxcut, self.maxcut
        return s

    # ---- [EXTER] wall potential (ExternalEnergyCalculator::extere2, scOOP/mc/externalenergycalculator.cpp:5-500)
    exter = None          # (thickness, epsilon, attraction switch) of the topology's [EXTER] section, or None

    def exter_setup(self):
        """topo.exter as Topo::genParamPairs / genTopoParams build it (structures/topo.cpp:120-130, 151-152):
        -> (params[ntypes, 8], exter_sqmaxcut)"""
        T = self.ntypes
        par = np.zeros((T, 8))
        sq = 0.0
        maxlength = 0.0
        L = lib()
        for i in range(T):
            q = self.ia[i, i]
            maxlength = max(maxlength, q[41])                 # len[0]
            if int(q[0]) == 0:
                continue
            L.sco_exter_params(_d(np.ascontiguousarray(q)), float(self.exter[0]), float(self.exter[1]), float(self.exter[2]), _d(par[i]))
            if par[i, 3] > sq:
                sq = par[i, 3]
        sq += maxlength
        sq *= sq * 1.1
        return par, sq

    def extere2(self, i, state=None):
        if self.exter is None:
            return 0.0
        if not hasattr(self, "_exter_cache"):
            self._exter_cache = self.exter_setup()
        par, sq = self._exter_cache
        st = np.ascontiguousarray(self.state[i] if state is None else state, dtype=np.float64)
        t = int(self.type[i])
        return lib().sco_extere2(_d(st), int(self.ia[t, t, 0]), _d(par[t]), float(sq), float(self.box[2]))

    # ---- oracle calls
    def conlist(self, i):
        cl = _ConC()
        s = self.c()
        lib().sco_get_conlist(C.byref(s), i, C.byref(cl))
        return cl

    def pair(self, i, j, cl=None, state_i=None):
        s = self.c()
        if cl is None:
            cl = self.conlist(i)
        si = np.ascontiguousarray(self.state[i] if state_i is None else state_i)
        return lib().sco_pair_energy(C.byref(s), _d(si), int(self.type[i]), int(self.moltype[i]), i,
                                     _d(self.state[j]), int(self.type[j]), j, C.byref(cl))

    def one_to_all(self, target, trial_state=None, pairs=False):
        s = self.c()
        ep = np.zeros(self.n) if pairs else None
        ts = None if trial_state is None else np.ascontiguousarray(trial_state, dtype=np.float64)
        e = lib().sco_one_to_all(C.byref(s), target, None if ts is None else _d(ts), None if ep is None else _d(ep))
        return (e, ep) if pairs else e

    def all_to_all(self, rows=False):
        s = self.c()
        er = np.zeros(self.n) if rows else None
        e = lib().sco_all_to_all(C.byref(s), None if er is None else _d(er))
        return (e, er) if rows else e

    def mol_to_others(self, first, m):
        s = self.c()
        return lib().sco_mol_to_others(C.byref(s), first, m)

    def overlap_pair(self, i, j, variant=0):
        s = self.c()
        return lib().sco_overlap_pair(C.byref(s), _d(self.state[i]), int(self.type[i]), _d(self.state[j]), int(self.type[j]), variant)

    def overlap_one(self, target, trial_state=None, variant=0):
        s = self.c()
        ts = None if trial_state is None else np.ascontiguousarray(trial_state, dtype=np.float64)
        return lib().sco_overlap_one(C.byref(s), target, None if ts is None else _d(ts), variant)

    def overlap_all(self, variant=0):
        s = self.c()
        return lib().sco_overlap_all(C.byref(s), variant)

    def cells(self):
        s = self.c()
        ncell = np.zeros(3, dtype=np.int32)
        cell_of = np.zeros(self.n, dtype=np.int32)
        lib().sco_cell_assign(C.byref(s), _i(cell_of), _i(ncell))
        ncells = int(ncell.prod())
        order = np.zeros(self.n, dtype=np.int32)
        start = np.zeros(ncells + 1, dtype=np.int32)
        lib().sco_cell_sort(C.byref(s), _i(cell_of), ncells, _i(order), _i(start))
        return ncell, cell_of, order, start

    def one_to_all_cells(self, target, cells, trial_state=None):
        s = self.c()
        ncell, cell_of, order, start = cells
        nc, ng = C.c_long(0), C.c_long(0)
        ts = None if trial_state is None else np.ascontiguousarray(trial_state, dtype=np.float64)
        e = lib().sco_one_to_all_cells(C.byref(s), target, None if ts is None else _d(ts), _i(cell_of), _i(ncell),
                                       _i(order), _i(start), C.byref(nc), C.byref(ng))
        return e, nc.value, ng.value

    def init_particles(self):
        """Conf::partVecInit, structures/Conf.cpp:98-103: derive patch sides / 2nd patch / chiral axes."""
        for i in range(self.n):
            t = int(self.type[i])
            if self.ia[t, t, 0] < otopo.SP:
                lib().sco_particle_init(_d(self.ia[t, t]), _d(self.state[i]))


_clib = None


def count_flops(system, targets):
    """Algorithmic operation counts of the cell-path one-to-all over `targets` (oracle/flopcount.cpp):
    returns dict(add, mul, div, sqrt, cos, acos, pow, candidates, gated, gate=dict(...)): the totals count the cutoff gate of
    PairE::operator() (image + |r|^2) ONCE per candidate; `gate` is that share, the rest is functor work of the gated pairs."""
    global _clib
    if _clib is None:
        path = os.path.join(HERE, "libscoracle_count.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", HERE, "count"])
        _clib = C.CDLL(path)
        _clib.cnt_sco_one_to_all_cells.restype = C.c_double
        _clib.cnt_sco_one_to_all_cells.argtypes = [C.POINTER(_SysC), C.c_int, _dp, _ip, _ip, _ip, _ip,
                                                   C.POINTER(C.c_long), C.POINTER(C.c_long)]
    s = system.c()
    ncell, cell_of, order, start = system.cells()
    _clib.cnt_reset()
    cand = gated = 0
    for t in targets:
        nc, ng = C.c_long(0), C.c_long(0)
        _clib.cnt_sco_one_to_all_cells(C.byref(s), int(t), None, _i(cell_of), _i(ncell), _i(order), _i(start), C.byref(nc), C.byref(ng))
        cand += nc.value
        gated += ng.value
    out = (C.c_longlong * 7)()
    _clib.cnt_get(out)
    names = ["add", "mul", "div", "sqrt", "cos", "acos", "pow"]
    d = {k: int(out[i]) for i, k in enumerate(names)}
    d["candidates"], d["gated"] = cand, gated
    _clib.cnt_get_gate(out)
    d["gate"] = {k: int(out[i]) for i, k in enumerate(names)}
    return d


def psc_rotate(state, geotype, angle, axis, positive):
    st = np.ascontiguousarray(state, dtype=np.float64).copy()
    ax = np.ascontiguousarray(axis, dtype=np.float64)
    lib().sco_psc_rotate(_d(st), int(geotype), float(angle), _d(ax), int(positive))
    return st


# ------------------------------------------------------------------------------------------------
# Wang-Landau order parameters of a whole configuration (oracle/wl_order.c)
# ------------------------------------------------------------------------------------------------
def type_volumes(system):
    """Ia_param::volume per particle type (scOOP/structures/topo.cpp:388-392): sphere + cylinder for the spherocylinder geotypes"""
    v = np.zeros(system.ntypes)
    for t in range(system.ntypes):
        sig, g, ln = system.ia[t, t, 3], int(system.ia[t, t, 0]), system.ia[t, t, 41]
        v[t] = 4.0 / 3.0 * math.pi * math.pow(sig / 2.0, 3.0)
        if 0 < g < 30:
            v[t] = 4.0 / 3.0 * math.pi * math.pow(sig / 2.0, 3.0) + math.pi / 2.0 * ln * math.pow(sig / 2.0, 2.0)
    return v


def wl_mass_center(system, volumes=None):
    """-> (syscm[3], sysvolume), Conf::massCenter"""
    v = np.ascontiguousarray(type_volumes(system) if volumes is None else volumes, dtype=np.float64)
    out = np.zeros(4)
    lib().sco_wl_mass_center(system.n, _d(system.state), _i(system.type), _d(v), _d(out))
    return out[:3].copy(), float(out[3])


def wl_raw(system, wlm, wlmtype=0, meshsize=0.0, volumes=None):
    """the quantity before binning of order parameter wlm (1, 2, 3, 4, 7, 8, 9); wlm 2 -> (maxsize, dim, occupied, skipped)"""
    L = lib()
    if wlm == 1:
        cm, _ = wl_mass_center(system, volumes)
        return L.sco_wl_z(_d(system.state), _d(np.ascontiguousarray(cm)), _d(system.box))
    if wlm == 2:
        dim = np.zeros(2, dtype=np.int32)
        occ, skip = C.c_long(0), C.c_long(0)
        m = L.sco_wl_mesh_hole(system.n, _d(system.state), _i(system.type), int(wlmtype), _d(system.box), float(meshsize), _i(dim), None,
                               C.byref(occ), C.byref(skip))
        return m, (int(dim[0]), int(dim[1])), occ.value, skip.value
    if wlm == 3:
        return float(system.state[0, 5])
    if wlm == 4:
        return L.sco_wl_two_part_dist(_d(system.state), _d(system.box))
    if wlm == 7:
        return L.sco_wl_contacts(system.n, _d(system.state), _i(system.type), int(wlmtype), _d(system.box))
    if wlm in (8, 9):
        return float(system.box[wlm - 8])
    raise ValueError("wlm %d" % wlm)


def wl_mesh_labels(system, wlmtype, meshsize):
    """Mesh::data after Mesh::meshInit: int32 [dim1][dim0], occupied < 0, free = hole number"""
    dim = np.zeros(2, dtype=np.int32)
    cap = int(system.box[0] / meshsize) * int(system.box[1] / meshsize)
    out = np.zeros(max(cap, 1), dtype=np.int32)
    lib().sco_wl_mesh_labels(system.n, _d(system.state), _i(system.type), int(wlmtype), _d(system.box), float(meshsize), _i(dim), _i(out))
    return out[:dim[0] * dim[1]].reshape(int(dim[1]), int(dim[0]))


def mesh_hash(a):
    """FNV-1a over the int32 values of a mesh (the hash oracle/ref_driver.cpp `wlorder` prints for Mesh::data)"""
    h = 0xcbf29ce484222325
    for v in np.asarray(a, dtype=np.int32).ravel().view(np.uint32).tolist():
        h = ((h ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def wl_bin(wlm, raw, minorder, dorder):
    return lib().sco_wl_bin(int(wlm), float(raw), float(minorder), float(dorder))


def load_wlorder_dump(path):
    """oracle/ref_driver.cpp `wlorder` dump -> dict: n, box, syscm, sysvolume, vol{type}, bins[(min, dorder)] -> {W1, W3, W4, W8, W9: order,
    W7: {type: (contacts, order)}}, mesh: list of (type, meshsize, dim0, dim1, maxsize, occupied, order at min 1 / dorder 4)"""
    op = gzip.open if path.endswith(".gz") else open
    out = {"vol": {}, "bins": {}, "mesh": []}
    cur = None
    with op(path, "rt") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "N":
                out["n"] = int(t[1])
            elif t[0] == "BOX":
                out["box"] = np.array([_hx(x) for x in t[1:4]])
            elif t[0] == "SYSCM":
                out["syscm"] = np.array([_hx(x) for x in t[1:4]])
                out["sysvolume"] = _hx(t[4])
            elif t[0] == "VOL":
                out["vol"][int(t[1])] = _hx(t[2])
            elif t[0] == "BIN":
                cur = out["bins"].setdefault((_hx(t[1]), _hx(t[2])), {"W7": {}})
            elif t[0] in ("W1", "W3", "W4", "W8", "W9"):
                cur[t[0]] = int(t[1])
            elif t[0] == "W7":
                cur["W7"][int(t[1])] = (int(t[2]), int(t[3]))
            elif t[0] == "W2":
                out["mesh"].append((int(t[1]), _hx(t[2]), int(t[3]), int(t[4]), int(t[5]), int(t[6]), int(t[7])))
                out.setdefault("mesh_labels", []).append((int(t[8]), int(t[9], 16)))       # number of holes, FNV-1a of Mesh::data
    return out


def system_from_text(top_text, config_text, counts=None):
    """options-free load: top.init + config.init text -> System (oracle-side parsers)."""
    t = otopo.read_top(top_text)
    types, moltypes, first = otopo.build_particle_lists(t, counts)
    box, state = otopo.read_config(config_text, len(types))
    otopo.make_chains_whole(t, moltypes, first, box, state)
    ia = otopo.pack_tables(t, types)
    mol = otopo.pack_mols(t, first)
    s = System(state, types, moltypes, ia, mol, box, t.sqmaxcut, t.maxcut)
    s.exter = t.exter
    s.init_particles()
    return s


def system_from_dir(path, counts=None):
    with open(os.path.join(path, "top.init")) as f:
        top = f.read()
    with open(os.path.join(path, "config.init")) as f:
        cfg = f.read()
    return system_from_text(top, cfg, counts)


def load_exter_dump(path):
    """oracle/ref_driver.cpp `exter` dump -> dict(exter=(exist, thickness, epsilon, attraction, sqmaxcut), params={type: [geotype, 8 values]},
    ext=array of extere2 per particle, state[n,30], type[n])"""
    op = gzip.open if path.endswith(".gz") else open
    out = {"params": {}, "ext": {}, "parts": []}
    with op(path, "rt") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "EXTP":
                out["exter"] = (int(t[1]),) + tuple(_hx(x) for x in t[2:6])
            elif t[0] == "EXTI":
                out["params"][int(t[1])] = [int(t[2])] + [_hx(x) for x in t[3:11]]
            elif t[0] == "EXT":
                out["ext"][int(t[1])] = _hx(t[2])
            elif t[0] == "P":
                out["parts"].append((int(t[2]), [_hx(x) for x in t[4:34]]))
            elif t[0] == "BOX":
                out["box"] = np.array([_hx(x) for x in t[1:4]])
    n = len(out["parts"])
    out["type"] = np.array([p[0] for p in out["parts"]], dtype=np.int32)
    out["state"] = np.array([p[1] for p in out["parts"]], dtype=np.float64).reshape(n, STATE)
    out["ext"] = np.array([out["ext"][i] for i in range(n)])
    return out


# ------------------------------------------------------------------------------------------------
# reference-driver dump (oracle/ref_driver.cpp `dump`) parser
# ------------------------------------------------------------------------------------------------
class RefDump:
    pass


def _hx(tok):
    return float.fromhex(tok)


def load_ref_dump(path):
    op = gzip.open if path.endswith(".gz") else open
    r = RefDump()
    ia_rows, mol_rows, parts = [], [], []
    r.pairs = {}
    r.one = {}
    r.mol2o = []
    r.overlaps = set()
    with op(path, "rt") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            k = t[0]
            if k == "N":
                r.n = int(t[1])
            elif k == "BOX":
                r.box = np.array([_hx(x) for x in t[1:4]])
            elif k == "CUT":
                r.sqmaxcut, r.maxcut = _hx(t[1]), _hx(t[2])
            elif k == "IA":
                ia_rows.append((int(t[1]), int(t[2]), int(t[3]), int(t[4]), int(t[5]), [_hx(x) for x in t[6:]]))
            elif k == "MOL":
                mol_rows.append((int(t[1]), int(t[2]), int(t[3]), [_hx(x) for x in t[4:]]))
            elif k == "P":
                parts.append((int(t[2]), int(t[3]), [_hx(x) for x in t[4:34]], [int(x) for x in t[34:39]]))
            elif k == "E":
                r.pairs[(int(t[1]), int(t[2]))] = _hx(t[3])
            elif k == "ONE":
                r.one[int(t[1])] = _hx(t[2])
            elif k == "TOTAL":
                r.total = _hx(t[1])
            elif k == "MOL2O":
                r.mol2o.append((int(t[2]), int(t[3]), _hx(t[4])))
            elif k == "OV":
                r.overlaps.add((int(t[1]), int(t[2])))
            elif k == "NPAIR":
                r.npair = int(t[1])
            elif k == "NOV":
                r.nov = int(t[1])
    types = np.array([p[0] for p in parts], dtype=np.int32)
    moltypes = np.array([p[1] for p in parts], dtype=np.int32)
    state = np.array([p[2] for p in parts], dtype=np.float64).reshape(-1, STATE)
    r.conlists = np.array([p[3] for p in parts], dtype=np.int32).reshape(-1, 5)
    T = int(types.max()) + 1 if len(types) else 1
    ia = np.zeros((T, T, IA_FIELDS))
    for (a, b, g0, g1, excl, v) in ia_rows:
        # v: sigma eps A B pdis pswitch pswitchINV rcut rcutSq rcutwca rcutwcaSq parallel | len0 len1 hl0 hl1 |
        #    pangl4 panglsw4 pcangl4 pcanglsw4 pcoshalfi4 psinhalfi4 | csec2 ssec2 chcos2 chsin2
        rec = np.zeros(IA_FIELDS)
        rec[0], rec[1], rec[2] = g0, g1, excl
        rec[3:15] = v[0:12]
        rec[41:43] = v[12:14]
        rec[15:17] = v[14:16]
        rec[17:21] = v[24:28]
        rec[21:25] = v[28:32]
        rec[25:29] = v[32:36]
        rec[29:33] = v[36:40]
        rec[33:35] = v[40:42]
        rec[35:37] = v[42:44]
        rec[37:39] = v[44:46]
        rec[39:41] = v[46:48]
        ia[a, b] = rec
    mol = np.zeros((len(mol_rows), MOL_FIELDS))
    for (m, msz, first, v) in mol_rows:
        mol[m, :12] = v[:12]
        mol[m, 12] = msz
        mol[m, 13] = first
    r.system = System(state, types, moltypes, ia, mol, r.box, r.sqmaxcut, r.maxcut)
    return r
