"""Deterministic synthetic inputs in the reference's own text formats (top.init / config.init).

`psc_bulk` is configuration "L" of SURVEY.md section 8(d) / BASELINE.json configs[1]: N = nx*ny*nz PSC rods of
Tests/test_01_normal_PSC's type on a 1.4 x 1.4 x 4.4 lattice (rho = 0.116), directions z + N(0, 0.02) tilt,
patch azimuth uniform, Python random.seed(12345). The other generators are small multi-cell systems for the
parity tests (random gas, type mix, bonded chains). Numbers are written with the precision the reference
writes config.last with (%15.8e), so the reference, the oracle and the product all read identical doubles.
"""
import math
import random

PSC_TYPE = "PSC    1.333333  1.2  1.346954458        0.3          90           5.0          3       0.0"

TOP_PSC = """[Types]
A       1         %s
[Molecules]
A: {
particles:   1
}
[System]
A %%d
""" % PSC_TYPE


def _fmt(v):
    return " ".join("%15.8e" % x for x in v)


def _rand_unit(rnd):
    while True:
        x, y, z = rnd.gauss(0, 1), rnd.gauss(0, 1), rnd.gauss(0, 1)
        n = math.sqrt(x * x + y * y + z * z)
        if n > 1e-6:
            return (x / n, y / n, z / n)


def _perp(rnd, d):
    while True:
        p = _rand_unit(rnd)
        dp = p[0] * d[0] + p[1] * d[1] + p[2] * d[2]
        q = (p[0] - dp * d[0], p[1] - dp * d[1], p[2] - dp * d[2])
        n = math.sqrt(q[0] ** 2 + q[1] ** 2 + q[2] ** 2)
        if n > 1e-3:
            return (q[0] / n, q[1] / n, q[2] / n)


def psc_bulk(nx=64, ny=64, nz=16, seed=12345, tilt=0.02):
    """-> (top_text, config_text, n). 64 x 64 x 16 = 65 536 rods in a 89.6 x 89.6 x 70.4 box."""
    rnd = random.Random(seed)
    sx, sy, sz = 1.4, 1.4, 4.4
    box = (nx * sx, ny * sy, nz * sz)
    lines = [_fmt(box)]
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                pos = ((ix + 0.5) * sx, (iy + 0.5) * sy, (iz + 0.5) * sz)
                d = (rnd.gauss(0, tilt), rnd.gauss(0, tilt), 1.0)
                n = math.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
                d = (d[0] / n, d[1] / n, d[2] / n)
                az = rnd.uniform(0.0, 2.0 * math.pi)
                # a vector perpendicular to d at azimuth az
                ex = (1.0 - d[0] * d[0], -d[0] * d[1], -d[0] * d[2])
                en = math.sqrt(ex[0] ** 2 + ex[1] ** 2 + ex[2] ** 2)
                ex = (ex[0] / en, ex[1] / en, ex[2] / en)
                ey = (d[1] * ex[2] - d[2] * ex[1], d[2] * ex[0] - d[0] * ex[2], d[0] * ex[1] - d[1] * ex[0])
                p = tuple(math.cos(az) * ex[k] + math.sin(az) * ey[k] for k in range(3))
                lines.append(_fmt(pos) + "   " + _fmt(d) + "   " + _fmt(p) + " 0")
    n = nx * ny * nz
    return TOP_PSC % n, "\n".join(lines) + "\n", n


def _gas(rnd, n, box, lines):
    for _ in range(n):
        pos = (rnd.uniform(0, box[0]), rnd.uniform(0, box[1]), rnd.uniform(0, box[2]))
        d = _rand_unit(rnd)
        p = _perp(rnd, d)
        lines.append(_fmt(pos) + "   " + _fmt(d) + "   " + _fmt(p) + " 0")


def small_case(kind, seed=2024):
    """small systems whose box holds a real 3-D cell grid (> 27 cells) -> (top_text, config_text)"""
    rnd = random.Random(seed)
    if kind == "psc_lattice":
        top, cfg, _ = psc_bulk(16, 16, 5, seed=seed, tilt=0.15)
        return top, cfg
    box = (26.0, 25.5, 33.0) if kind == "chains" else (22.0, 21.0, 23.5)
    lines = [_fmt(box)]
    if kind == "psc_gas":
        _gas(rnd, 700, box, lines)
        return TOP_PSC % 700, "\n".join(lines) + "\n"
    if kind in ("rods_wide", "rods_dense"):
        # un-bonded rods of four types with different cutoffs on a grid of >= 7 x 5 x 5 cells: the configuration the thread-per-target
        # gate (k_gate_rows) takes; clustered along x so that cell populations are very uneven (empty cells, cells beyond 32 rods)
        top = """[Types]
P1 1 PSC    1.333333  1.2  1.346954458  0.3  90  5.0  3  0.5
C2 2 CPSC   1.1       1.1  1.30         0.45 120 5.0  3  0.5
T3 3 TCPSC  1.0       1.0  1.2          0.4  80  8.0  3  -0.4 150.0 90 5.0
H4 4 CHPSC  1.2       1.2  1.346954458  0.3  90  5.0  3  0.0  10.0
[Molecules]
A: {
particles: 1
}
B: {
particles: 2
}
C: {
particles: 3
}
D: {
particles: 4
}
[System]
A %d
B %d
C %d
D %d
"""
        # rods_dense: four times the particles -> neighbourhoods above the staged tile of k_gate_rows: the launch must fall back
        mult = 4 if kind == "rods_dense" else 1
        top = top % (900 * mult, 700 * mult, 600 * mult, 500 * mult)
        box = (44.0, 29.0, 31.0)
        lines = [_fmt(box)]
        for i in range(2700 * mult):
            if i % 3 == 0:
                pos = (rnd.gauss(0.3, 0.07) % 1.0 * box[0], rnd.uniform(0, box[1]), rnd.gauss(0.5, 0.2) % 1.0 * box[2])
            else:
                pos = (rnd.uniform(0, box[0]), rnd.uniform(0, box[1]), rnd.uniform(0, box[2]))
            d = _rand_unit(rnd)
            p = _perp(rnd, d)
            lines.append(_fmt(pos) + "   " + _fmt(d) + "   " + _fmt(p) + " 0")
        return top, "\n".join(lines) + "\n"
    if kind == "chain_fluid":
        # 480 bonded SPN-SPA-SPA trimers + 160 PSC rods on a lattice without overlaps: the system of the chain-move sweep tests
        top = """[Types]
H2 2 SPN    1.0 0.95
T3 3 SPA    1.0 1.0 1.12246205 0.6
Q4 4 PSC    1.0 1.2 1.346954458 0.3 120 5.0 3 0.3
[Molecules]
B: {
bond1: 10.0 1.0
particles: 2
particles: 3
particles: 3
}
E: {
particles: 4
}
[System]
B 480
E 160
"""
        box = (26.0, 25.6, 33.0)
        sx, sy, sz = box[0] / 8, box[1] / 8, box[2] / 10
        rods, trimers = [], []
        for iz in range(10):
            for iy in range(8):
                for ix in range(8):
                    c = ((ix + 0.5) * sx, (iy + 0.5) * sy, (iz + 0.5) * sz)
                    if (ix + iy) % 2 == 0 and iz % 2 == 0:
                        rods.append(c)
                    else:
                        trimers.append(c)
        lines = [_fmt(box)]
        for c in trimers:
            for k in (-1, 0, 1):
                pos = (c[0] + k * 1.0, c[1], c[2])
                lines.append(_fmt(pos) + "   " + _fmt((0.0, 0.0, 1.0)) + "   " + _fmt((1.0, 0.0, 0.0)) + " 0")
        for c in rods:
            az = rnd.uniform(0.0, 2.0 * math.pi)
            lines.append(_fmt(c) + "   " + _fmt((0.0, 1.0, 0.0)) + "   " + _fmt((math.cos(az), 0.0, math.sin(az))) + " 0")
        return top, "\n".join(lines) + "\n"
    if kind == "rods_in_spheres":
        # 48 000 small attractive spheres and 40 long patchy rods on a 5 x 5 x 5 grid: a sphere has ~20 partners, a rod ~300 -- more
        # than the per-target buffers of k_gate_rows_gen hold, so the rod type is split off to k_gate_cells (adaptive, per launch)
        top = """[Types]
S1 1 SPA    1.0  1.0  1.12246205  1.0
R2 2 CPSC   1.0  1.0  1.12246205  1.0  120 5.0  6  0.0
[Molecules]
A: {
particles: 1
}
B: {
particles: 2
}
[System]
A 48000
B 40
"""
        box = (46.0, 46.0, 46.0)
        lines = [_fmt(box)]
        _gas(rnd, 48040, box, lines)
        return top, "\n".join(lines) + "\n"
    if kind == "mix":
        top = """[Types]
S1 1 SPA    1.333333  1.2  1.346954458  0.3
P2 2 PSC    1.333333  1.2  1.346954458  0.3  90  5.0  3  0.5
C3 3 CPSC   1.1       1.1  1.30         0.35 120 5.0  3  0.5
T4 4 TCHPSC 1.2       1.2  1.346954458  0.3  90  5.0  3  0.0  180.0 60 5.0 10.0
N5 5 SPN    1.0  0.95
D6 6 TCPSC  1.0       1.0  1.2          0.4  80  8.0  3  -0.4 150.0 90 5.0
[Molecules]
A: {
particles: 1
}
B: {
particles: 2
}
C: {
particles: 3
}
D: {
particles: 4
}
E: {
particles: 5
}
F: {
particles: 6
}
[System]
A 120
B 150
C 150
D 120
E 80
F 100
[EXCLUDE]
1 3
"""
        _gas(rnd, 720, box, lines)
        return top, "\n".join(lines) + "\n"
    if kind == "chains":
        top = """[Types]
R1 1 TCPSC  1.333333 1.2 1.346954458 0.3 90 5.0 3 0.0 180.0 90 5.0
H2 2 SPN    1.0 0.95
T3 3 SPA    1.0 1.0 1.12246205 1.6
Q4 4 PSC    1.0 1.2 1.346954458 0.3 120 5.0 3 0.3
[Molecules]
A: {
particles: 1
particles: 1
particles: 1
particles: 1
particles: 1
bond1: 1.0 1.0
bond2: 1.0 2.0
angle1: 2.0 20.0
angle2: 1.5 30.0
}
B: {
bond1: 90.0 0.4
bond2: 10.0 4.0
particles: 2
particles: 3
particles: 3
}
C: {
particles: 4
particles: 3
particles: 4
bondh: 5.0 0.5
angle1: 3.0 10.0
}
D: {
particles: 4
particles: 4
bondd: 4.0 0.7
}
E: {
particles: 4
}
[System]
A 40
B 100
C 30
D 30
E 100
"""
        def chain(nmem, step):
            pos = [rnd.uniform(0, box[0]), rnd.uniform(0, box[1]), rnd.uniform(0, box[2])]
            d = _rand_unit(rnd)
            for _ in range(nmem):
                dd = _rand_unit(rnd)
                d2 = tuple(d[k] + 0.25 * dd[k] for k in range(3))
                nn = math.sqrt(sum(x * x for x in d2))
                d2 = tuple(x / nn for x in d2)
                p = _perp(rnd, d2)
                lines.append(_fmt(pos) + "   " + _fmt(d2) + "   " + _fmt(p) + " 0")
                jit = _perp(rnd, d2)   # off-axis jitter: a partner exactly on a rod's axis is a 0/0 in the reference
                pos = [pos[k] + step * d2[k] + 0.3 * jit[k] for k in range(3)]
        for _ in range(40):
            chain(5, 4.0)
        for _ in range(100):
            chain(3, 1.0)
        for _ in range(30):
            chain(3, 2.6)
        for _ in range(30):
            chain(2, 4.2)
        _gas(rnd, 100, box, lines)
        return top, "\n".join(lines) + "\n"
    raise ValueError(kind)


def membrane(tiles_x, tiles_y, top_text, config_text):
    """BASELINE.json configs[2] / SURVEY.md 8(d) "M": tile the 601-particle membrane of Tests/SC_PSC_MEMBRANE_WANG
    (1 CPSC + 200 SPN-SPA-SPA lipids, box 10.888 x 10.888 x 50) tiles_x x tiles_y times in the bilayer plane.
    21 x 21 -> 441 CPSC + 88 200 lipids = 265 041 particles. Returns (top_text, config_text, n)."""
    lines = [l for l in config_text.split("\n") if l.strip()]
    box = [float(x) for x in lines[0].split()[:3]]
    parts = [l.split() for l in lines[1:]]
    # [System] of the shipped topology: A 1 (the CPSC), B 200 (lipids of 3 beads)
    head, tail = parts[:1], parts[1:]
    out = [_fmt((box[0] * tiles_x, box[1] * tiles_y, box[2]))]

    def emit(block):
        for ty in range(tiles_y):
            for tx in range(tiles_x):
                for t in block:
                    v = [float(x) for x in t[:9]]
                    v[0] += tx * box[0]
                    v[1] += ty * box[1]
                    out.append(_fmt(v[0:3]) + "   " + _fmt(v[3:6]) + "   " + _fmt(v[6:9]) + " 0")
    emit(head)
    emit(tail)
    nt = tiles_x * tiles_y
    top = []
    in_system = False
    for l in top_text.split("\n"):
        s_ = l.strip()
        if s_.upper().startswith("[SYSTEM]"):
            in_system = True
            top.append(l)
            continue
        if in_system and s_ and not s_.startswith("#") and not s_.startswith("["):
            name, cnt = s_.split()[:2]
            top.append("%s %d" % (name, int(cnt) * nt))
            continue
        if s_.startswith("["):
            in_system = False
        top.append(l)
    n = len(parts) * nt
    return "\n".join(top) + "\n", "\n".join(out) + "\n", n


def tile(top_text, config_text, kx, ky, kz):
    """Replicate ANY (top.init, config.init) pair kx x ky x kz times (SURVEY.md 8(d) "X": Tests/test_14 and Tests/test_20 tiled
    12^3 -> 65 664 / 69 120 particles). Particles stay grouped by the molecule blocks of [System], so molecule indices,
    chain positions and connectivity lists of every copy are those of the original. Returns (top_text, config_text, n)."""
    lines = [l for l in config_text.split("\n") if l.strip()]
    box = [float(x) for x in lines[0].split()[:3]]
    parts = [l.split() for l in lines[1:]]
    # molecule sizes from [Molecules], block counts from [System]
    sizes, counts, order = {}, [], []
    section, cur = "", None
    for l in top_text.split("\n"):
        s_ = l.split("#")[0].strip()
        if not s_:
            continue
        if s_.startswith("["):
            section = s_.upper()
            continue
        if section.startswith("[MOLECULES]"):
            if s_.endswith("{"):
                cur = s_.split(":")[0].strip()
                sizes[cur] = 0
            elif s_.startswith("}"):
                cur = None
            elif cur is not None and s_.lower().startswith("particles"):
                sizes[cur] += 1
        elif section.startswith("[SYSTEM]"):
            name, cnt = s_.split()[:2]
            counts.append((name, int(cnt)))
    out = [_fmt((box[0] * kx, box[1] * ky, box[2] * kz))]
    at = 0
    for name, cnt in counts:
        blk = [list(t) for t in parts[at:at + cnt * sizes[name]]]
        at += cnt * sizes[name]
        # config files hold positions wrapped into the box: bring every chain member next to its predecessor first, otherwise a
        # bond that crosses the boundary of the small box would be stretched over it in the large one
        for m0 in range(0, len(blk), sizes[name]):
            for k in range(m0 + 1, m0 + sizes[name]):
                for d in range(3):
                    delta = float(blk[k][d]) - float(blk[k - 1][d])
                    blk[k][d] = repr(float(blk[k][d]) - box[d] * round(delta / box[d]))
        for tz in range(kz):
            for ty in range(ky):
                for tx in range(kx):
                    for t in blk:
                        v = [float(x) for x in t[:9]]
                        v[0] += tx * box[0]; v[1] += ty * box[1]; v[2] += tz * box[2]
                        out.append(_fmt(v[0:3]) + "   " + _fmt(v[3:6]) + "   " + _fmt(v[6:9]) + " 0")
    assert at == len(parts), "config.init does not match the [System] block of top.init"
    nt = kx * ky * kz
    top, in_system = [], False
    for l in top_text.split("\n"):
        s_ = l.split("#")[0].strip()
        if s_.upper().startswith("[SYSTEM]"):
            in_system = True
            top.append(l)
            continue
        if in_system and s_ and not s_.startswith("["):
            name, cnt = s_.split()[:2]
            top.append("%s %d" % (name, int(cnt) * nt))
            continue
        if s_.startswith("["):
            in_system = False
        top.append(l)
    return "\n".join(top) + "\n", "\n".join(out) + "\n", len(parts) * nt
