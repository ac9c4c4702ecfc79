"""TEST INFRASTRUCTURE -- oracle-side restatement (numpy / pure Python) of the reference's input
semantics for the hot path. NOT part of the product.

Restates, citing /root/reference/scOOP:
  * top.init parsing and per-type fill      mc/inicializer.cpp:453-585, 660-852, 929-1108
  * pair mixing rules                        structures/topo.cpp:5-138   (Topo::genParamPairs)
  * global cutoff                            structures/topo.cpp:140-153 (Topo::genTopoParams)
  * config.init parsing                      mc/inicializer.cpp:89-240   (Inicializer::initConfig)
  * particle list / group list               mc/inicializer.cpp:409-450, 330-372

Pinned against the reference driver's IA / MOL / P dumps (tests/test_oracle_golden.py).
Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may import this.
"""
import math
import numpy as np

SCN, SCA, PSC, CPSC, CHPSC, CHCPSC, TPSC, TCPSC, TCHPSC, TCHCPSC = range(10, 20)
SP = SPN = 30
SPA = 31
GEOTYPES = {"SCN": SCN, "SCA": SCA, "PSC": PSC, "CPSC": CPSC, "CHPSC": CHPSC, "CHCPSC": CHCPSC,
            "TPSC": TPSC, "TCPSC": TCPSC, "TCHPSC": TCHPSC, "TCHCPSC": TCHCPSC, "SP": SP, "SPN": SPN, "SPA": SPA}
PI = 3.141592653589793238462643383279
DEGTORAD = 0.017453292519943295769236907683
MAXT = 40
IA_FIELDS = 48
MOL_FIELDS = 16
STATE = 30

# field offsets inside the 48-double record (must match sco_iaparam / scgpu_iaparam)
F = dict(geotype=0, exclude=2, sigma=3, epsilon=4, A=5, B=6, pdis=7, pswitch=8, pswitchINV=9, rcut=10, rcutSq=11,
         rcutwca=12, rcutwcaSq=13, parallel=14, half_len=15, pcangl=17, pcanglsw=21, pcoshalfi=25, psinhalfi=29,
         csecpatchrot=33, ssecpatchrot=35, chiral_cos=37, chiral_sin=39, len=41)


class IaParam:
    """Ia_param defaults, structures/structures.h:222-253 (plus zero for members the ctor leaves unset
    but the parser always overwrites before use)."""

    def __init__(self):
        self.geotype = [0, 0]
        self.sigma = self.epsilon = self.A = self.B = 0.0
        self.pdis = self.pswitch = self.pswitchINV = 0.0
        self.rcut = self.rcutSq = self.rcutwca = self.rcutwcaSq = 0.0
        self.parallel = 0.0
        self.half_len = [0.0, 0.0]
        self.len = [0.0, 0.0]
        self.pangl = [0.0] * 4
        self.panglsw = [0.0] * 4
        self.pcangl = [0.0] * 4
        self.pcanglsw = [0.0] * 4
        self.pcoshalfi = [0.0] * 4
        self.psinhalfi = [0.0] * 4
        self.csecpatchrot = [0.0, 0.0]
        self.ssecpatchrot = [0.0, 0.0]
        self.chiral_cos = [0.0, 0.0]
        self.chiral_sin = [0.0, 0.0]
        self.exclude = False

    def pack(self):
        r = np.zeros(IA_FIELDS)
        r[0:2] = self.geotype
        r[2] = 1.0 if self.exclude else 0.0
        r[3:15] = [self.sigma, self.epsilon, self.A, self.B, self.pdis, self.pswitch, self.pswitchINV,
                   self.rcut, self.rcutSq, self.rcutwca, self.rcutwcaSq, self.parallel]
        r[15:17] = self.half_len
        r[17:21] = self.pcangl
        r[21:25] = self.pcanglsw
        r[25:29] = self.pcoshalfi
        r[29:33] = self.psinhalfi
        r[33:35] = self.csecpatchrot
        r[35:37] = self.ssecpatchrot
        r[37:39] = self.chiral_cos
        r[39:41] = self.chiral_sin
        r[41:43] = self.len
        return r


class MolParam:
    """MoleculeParams defaults, structures/moleculeparams.h:35-38"""

    def __init__(self, name):
        self.name = name
        self.bond1eq = self.bond1c = self.bond2eq = self.bond2c = -1.0
        self.bonddeq = self.bonddc = self.bondheq = self.bondhc = -1.0
        self.angle1eq = self.angle1c = self.angle2eq = self.angle2c = -1.0
        self.particle_types = []


def _strip(line):
    i = line.find("#")
    if i >= 0:
        line = line[:i]
    return line.strip(" \t\r\n")


class Topology:
    def __init__(self):
        self.ia = [[IaParam() for _ in range(MAXT)] for _ in range(MAXT)]
        self.mols = []
        self.system = []          # (molname, count)
        self.exclusions = set()
        self.sqmaxcut = 0.0
        self.maxcut = 0.0
        self.exter = None

    # ---- Inicializer::fillTypes, mc/inicializer.cpp:660-852
    def fill_type(self, line):
        tok = line.split()
        name, typ, geo = tok[0], int(tok[1]), tok[2]
        param = [float(x) for x in tok[3:15]] + [0.0] * 12
        fields = min(len(tok), 15) - 5
        g = GEOTYPES.get(geo, 0)
        if not g:
            raise ValueError("TOPOLOGY ERROR: Unknown GEOTYPE: %s" % geo)
        need = {SPN: 0, SCN: 1, SPA: 2, SCA: 3, PSC: 6, CPSC: 6, CHPSC: 7, CHCPSC: 7, TPSC: 9, TCPSC: 9, TCHPSC: 10, TCHCPSC: 10}[g]
        if fields != need:
            raise ValueError("TOPOLOGY ERROR: wrong number of parameters for %s" % geo)
        p = self.ia[typ][typ]
        p.geotype = [g, g]
        p.epsilon = param[0]
        p.sigma = param[1]
        p.A = 4 * p.epsilon * math.pow(p.sigma, 12)
        p.B = 4 * p.epsilon * math.pow(p.sigma, 6)
        p.rcutwca = p.sigma * math.pow(2.0, 1.0 / 6.0)
        p.rcutwcaSq = p.rcutwca * p.rcutwca
        if fields > 0 and fields != 1 and fields != 3:
            p.pdis = param[2]
            p.pswitch = param[3]
            p.pswitchINV = 1.0 / param[3]
            p.rcut = (p.pswitch + p.pdis) if g != SPN else 0.0
            p.rcutSq = p.rcut * p.rcut
        if fields == 1:
            p.len = [param[2]] * 2
            p.half_len = [param[2] / 2] * 2
        if fields == 3:
            p.len = [param[4]] * 2
            p.half_len = [param[4] / 2] * 2
        if fields > 2 and fields != 3:
            for i in range(2):
                p.len[i] = param[6]
                p.half_len[i] = param[6] / 2
                p.pangl[i] = param[4]
                p.panglsw[i] = param[5]
                p.pcangl[i] = math.cos(param[4] / 2.0 / 180 * PI)
                p.pcanglsw[i] = math.cos((param[4] / 2.0 + param[5]) / 180 * PI)
                p.pcoshalfi[i] = math.cos((param[4] / 2.0 + param[5]) / 2.0 / 180 * PI)
                p.psinhalfi[i] = math.sqrt(1.0 - p.pcoshalfi[i] * p.pcoshalfi[i])
                p.parallel = param[7]
        if fields == 7:
            for i in range(2):
                p.chiral_cos[i] = math.cos(param[8] / 360 * PI)
                p.chiral_sin[i] = math.sqrt(1 - p.chiral_cos[i] * p.chiral_cos[i])
        if fields in (9, 10):
            for i in range(2):
                p.csecpatchrot[i] = math.cos(param[8] / 360 * PI)
                p.ssecpatchrot[i] = math.sqrt(1 - p.csecpatchrot[i] * p.csecpatchrot[i])
                p.pangl[i + 2] = param[9]
                p.panglsw[i + 2] = param[10]
                p.pcangl[i + 2] = math.cos(param[9] / 2.0 / 180 * PI)
                p.pcanglsw[i + 2] = math.cos((param[9] / 2.0 + param[10]) / 180 * PI)
                p.pcoshalfi[i + 2] = math.cos((param[9] / 2.0 + param[10]) / 2.0 / 180 * PI)
                p.psinhalfi[i + 2] = math.sqrt(1.0 - p.pcoshalfi[i + 2] * p.pcoshalfi[i + 2])
        if fields == 10:
            for i in range(2):
                p.chiral_cos[i] = math.cos(param[11] / 360 * PI)
                p.chiral_sin[i] = math.sqrt(1 - p.chiral_cos[i] * p.chiral_cos[i])
        if p.rcutwca > self.sqmaxcut:
            self.sqmaxcut = p.rcutwca
        if p.rcut > self.sqmaxcut:
            self.sqmaxcut = p.rcut

    # ---- Inicializer::fillMol, mc/inicializer.cpp:929-1108
    def fill_mol(self, mol, body):
        body = body.split("}")[0]
        if "{" in body:
            body = body.split("{", 1)[1]
        body = body.strip()
        if not body:
            return
        cmd, _, params = body.partition(":")
        cmd = cmd.strip().upper()
        vals = params.split()
        if cmd == "PARTICLES":
            mol.particle_types.append(int(vals[0]))
            return
        k, eq = float(vals[0]), float(vals[1])
        if cmd == "BOND1":
            mol.bond1c, mol.bond1eq = k, eq
        elif cmd == "BOND2":
            mol.bond2c, mol.bond2eq = k, eq
        elif cmd == "BONDD":
            mol.bonddc, mol.bonddeq = k, eq
        elif cmd == "BONDH":
            mol.bondhc, mol.bondheq = k, eq
        elif cmd == "ANGLE1":
            mol.angle1c, mol.angle1eq = k, eq * DEGTORAD
        elif cmd == "ANGLE2":
            mol.angle2c, mol.angle2eq = k, eq * DEGTORAD
        else:
            raise ValueError("TOPOLOGY ERROR: unknown parameter: %s" % cmd)

    # ---- Topo::genParamPairs, structures/topo.cpp:5-138
    def gen_param_pairs(self):
        ia = self.ia
        for i in range(MAXT):
            for j in range(MAXT):
                if i == j:
                    continue
                if ia[j][j].geotype[0] == 0 or ia[i][i].geotype[0] == 0:
                    continue
                a = (i, j)
                q = ia[i][j]
                for k in range(2):
                    s = ia[a[k]][a[k]]
                    q.geotype[k] = s.geotype[0]
                    q.len[k] = s.len[0]
                    q.half_len[k] = s.half_len[0]
                    if PSC <= q.geotype[k] < SP:
                        q.pangl[k] = s.pangl[0]
                        q.panglsw[k] = s.panglsw[0]
                        q.pcangl[k] = math.cos(q.pangl[k] / 2.0 / 180 * PI)
                        q.pcanglsw[k] = math.cos((q.pangl[k] / 2.0 + q.panglsw[k]) / 180 * PI)
                        q.pcoshalfi[k] = math.cos((q.pangl[k] / 2.0 + q.panglsw[k]) / 2.0 / 180 * PI)
                        q.psinhalfi[k] = math.sqrt(1.0 - q.pcoshalfi[k] * q.pcoshalfi[k])
                    if q.geotype[k] in (CHCPSC, CHPSC, TCHCPSC, TCHPSC):
                        q.chiral_cos[k] = s.chiral_cos[0]
                        q.chiral_sin[k] = s.chiral_sin[0]
                    if q.geotype[k] in (TCPSC, TPSC, TCHCPSC, TCHPSC):
                        q.csecpatchrot[k] = s.csecpatchrot[0]
                        q.ssecpatchrot[k] = s.ssecpatchrot[0]
                        q.pangl[k + 2] = s.pangl[2]
                        q.panglsw[k + 2] = s.panglsw[2]
                        q.pcangl[k + 2] = math.cos(q.pangl[k + 2] / 2.0 / 180 * PI)
                        q.pcanglsw[k + 2] = math.cos((q.pangl[k + 2] / 2.0 + q.panglsw[k + 2]) / 180 * PI)
                        q.pcoshalfi[k + 2] = math.cos((q.pangl[k + 2] / 2.0 + q.panglsw[k + 2]) / 2.0 / 180 * PI)
                        q.psinhalfi[k + 2] = math.sqrt(1.0 - q.pcoshalfi[k + 2] * q.pcoshalfi[k + 2])
                pi_, pj_ = ia[i][i], ia[j][j]
                q.sigma = (pi_.sigma + pj_.sigma) * 0.5
                q.epsilon = math.sqrt(pi_.epsilon * pj_.epsilon)
                q.A = 4 * q.epsilon * math.pow(q.sigma, 12)
                q.B = 4 * q.epsilon * math.pow(q.sigma, 6)
                q.pswitch = (pi_.pswitch + pj_.pswitch) * 0.5
                q.pswitchINV = (1.0 / q.pswitch) if q.pswitch != 0.0 else math.inf
                q.rcutwca = q.sigma * math.pow(2.0, 1.0 / 6.0)
                q.rcutwcaSq = q.rcutwca * q.rcutwca
                if pi_.parallel > 0 and pj_.parallel > 0:
                    q.parallel = math.sqrt(pi_.parallel * pj_.parallel)
                if pi_.parallel < 0 and pj_.parallel < 0:
                    q.parallel = -math.sqrt(pi_.parallel * pj_.parallel)
                q.pdis = ((pi_.pdis - pi_.rcutwca) + (pj_.pdis - pj_.rcutwca)) * 0.5 + q.rcutwca
                if q.geotype[0] == SPN or q.geotype[1] == SPN:
                    q.rcut = 0.0
                else:
                    q.rcut = q.pswitch + q.pdis
                q.rcutSq = q.rcut * q.rcut
                if q.rcutwca > self.sqmaxcut:
                    self.sqmaxcut = q.rcutwca
                if q.rcut > self.sqmaxcut:
                    self.sqmaxcut = q.rcut
        for i in range(MAXT):
            for j in range(MAXT):
                ia[i][j].exclude = (i, j) in self.exclusions

    # ---- Topo::genTopoParams, structures/topo.cpp:140-153
    def gen_topo_params(self):
        maxlength = 0.0
        for i in range(MAXT):
            if maxlength < self.ia[i][i].len[0]:
                maxlength = self.ia[i][i].len[0]
        self.sqmaxcut += maxlength
        self.sqmaxcut *= 1.1
        self.maxcut = self.sqmaxcut
        self.sqmaxcut = self.sqmaxcut * self.sqmaxcut


def read_top(text):
    """Inicializer::readTopoFile, mc/inicializer.cpp:453-585 -> Topology (pairs generated)."""
    t = Topology()
    key = ""
    cur = None
    lines = text.split("\n")
    li = 0
    while li < len(lines):
        raw = lines[li]
        li += 1
        while raw.rstrip().endswith("\\") and li < len(lines):
            raw = raw.rstrip()[:-1] + lines[li]
            li += 1
        line = _strip(raw)
        if not line:
            continue
        if line[0] == "[":
            key = line[1:].split("]")[0].strip().upper()
            continue
        if key == "TYPES":
            t.fill_type(line)
        elif key == "MOLECULES":
            if cur is None:
                name = line.split(":")[0].strip()
                cur = MolParam(name)
                t.mols.append(cur)
            t.fill_mol(cur, line)
            if "}" in line:
                cur = None
        elif key == "SYSTEM":
            tok = line.split()
            t.system.append((tok[0], int(tok[1])))
        elif key == "POOL":
            pass
        elif key == "EXTER":
            t.exter = [float(x) for x in line.split()[:3]]
        elif key == "EXCLUDE":
            nums = [int(x) for x in line.split()]
            for a, b in zip(nums[0::2], nums[1::2]):
                t.exclusions.add((a, b))
                t.exclusions.add((b, a))
        else:
            raise ValueError("TOPOLOGY ERROR: invalid keyword: %s" % key)
    t.gen_param_pairs()
    t.gen_topo_params()
    return t


def build_particle_lists(t, counts=None):
    """Inicializer::setParticlesParamss + initGroupLists (mc/inicializer.cpp:409-450, 330-372).
    Returns type[n], moltype[n], first[nmol+1]. `counts` overrides [System] counts per entry."""
    names = [m.name for m in t.mols]
    types, moltypes = [], []
    for si, (name, cnt) in enumerate(t.system):
        mol = names.index(name)
        c = cnt if counts is None else counts[si]
        for _ in range(c):
            for ty in t.mols[mol].particle_types:
                types.append(ty)
                moltypes.append(mol)
    types = np.asarray(types, dtype=np.int32)
    moltypes = np.asarray(moltypes, dtype=np.int32)
    nmol = len(t.mols)
    first = np.zeros(nmol + 1, dtype=np.int64)
    # particles are grouped by molecule type in [System] order == molecule-type order in the reference's tests
    for m in range(nmol):
        idx = np.nonzero(moltypes == m)[0]
        first[m] = idx[0] if len(idx) else (first[m - 1] + np.count_nonzero(moltypes == m - 1) if m > 0 else 0)
    first[nmol] = len(types)
    return types, moltypes, first


def _use_pbc(v):
    """Cuboid::usePBC(Vector&), structures/geometry.h:79-98"""
    while v < 0.0:
        v += 1.0
    while v > 1.0:
        v -= 1.0
    return v


def read_config(text, n):
    """Inicializer::initConfig, mc/inicializer.cpp:89-240: returns box[3], state[n,30] with pos scaled to
    the unit box and wrapped, dir normalised, patchdir orthonormalised. Derived vectors are NOT yet set."""
    lines = [_strip(l) for l in text.split("\n")]
    lines = [l for l in lines]
    li = 0
    while li < len(lines) and not lines[li]:
        li += 1
    box = np.array([float(x) for x in lines[li].split()[:3]])
    li += 1
    state = np.zeros((n, STATE))
    for i in range(n):
        tok = lines[li].split()
        li += 1
        vals = [float(x) for x in tok[:9]]
        pos = [_use_pbc(vals[d] / box[d]) for d in range(3)]
        d = np.array(vals[3:6])
        # Vector::normalise, structures/Vector.h:56-64
        tot = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if tot != 0.0:
            tot = 1.0 / tot
            d = np.array([d[0] * tot, d[1] * tot, d[2] * tot])
        p = np.array(vals[6:9])
        dp = p[0] * d[0] + p[1] * d[1] + p[2] * d[2]
        p = np.array([p[0] - dp * d[0], p[1] - dp * d[1], p[2] - dp * d[2]])
        tot = math.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2])
        if tot != 0.0:
            tot = 1.0 / tot
            p = np.array([p[0] * tot, p[1] * tot, p[2] * tot])
        state[i, 0:3] = pos
        state[i, 3:6] = d
        state[i, 6:9] = p
    return box, state


def pack_tables(t, types):
    """Compact T x T x 48 table indexed by the reference's type numbers, T = max type in use + 1;
    and the nmol x 16 molecule table (first[] filled by the caller)."""
    T = int(max(types)) + 1
    ia = np.zeros((T, T, IA_FIELDS))
    for i in range(T):
        for j in range(T):
            ia[i, j] = t.ia[i][j].pack()
    return ia


def pack_mols(t, first):
    mol = np.zeros((len(t.mols), MOL_FIELDS))
    for m, q in enumerate(t.mols):
        mol[m, :12] = [q.bond1eq, q.bond1c, q.bond2eq, q.bond2c, q.bonddeq, q.bonddc, q.bondheq, q.bondhc,
                       q.angle1eq, q.angle1c, q.angle2eq, q.angle2c]
        mol[m, 12] = len(q.particle_types)
        mol[m, 13] = first[m]
    return mol


def make_chains_whole(t, moltypes, first, box, state):
    """Inicializer::initConfig tail (mc/inicializer.cpp:268-278) + Conf::makeMoleculeWhole
    (structures/Conf.h:371-380): each chain particle is placed next to its predecessor's nearest image."""
    n = len(moltypes)
    i = 0
    while i < n:
        m = int(moltypes[i])
        msz = len(t.mols[m].particle_types)
        if msz > 1:
            for k in range(i + 1, i + msz):
                for d in range(3):
                    r = state[k, d] - state[k - 1, d]
                    r = box[d] * (r - float(np.rint(r)))
                    r /= box[d]
                    state[k, d] = state[k - 1, d] + r
        i += msz
