// See topology.hpp. Semantics follow the reference's parsers line by line (citations inline); the code is ours.
#include "topology.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace schost {

static const double PI = 3.141592653589793238462643383279;          // macros.h:65
static const double DEGTORAD = 0.017453292519943295769236907683;    // macros.h:67

enum { SCN = 10, SCA, PSC, CPSC, CHPSC, CHCPSC, TPSC, TCPSC, TCHPSC, TCHCPSC, SP = 30, SPN = 30, SPA = 31 };

static std::string strip(const std::string& raw) {   // strip_comment + trim (simlib.cpp:7-49)
    std::string s = raw.substr(0, raw.find('#'));
    size_t b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) return "";
    size_t e = s.find_last_not_of(" \t\r\n");
    return s.substr(b, e - b + 1);
}

static std::vector<std::string> split(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}

static std::string upper(std::string s) {
    for (auto& c : s) c = (char)toupper((unsigned char)c);
    return s;
}

static int convertGeotype(const std::string& g) {   // Inicializer::convertGeotype (inicializer.cpp:855-885)
    static const std::pair<const char*, int> tab[] = {{"SCN", SCN}, {"SCA", SCA}, {"PSC", PSC}, {"CPSC", CPSC}, {"CHPSC", CHPSC},
        {"CHCPSC", CHCPSC}, {"TPSC", TPSC}, {"TCPSC", TCPSC}, {"TCHPSC", TCHPSC}, {"TCHCPSC", TCHCPSC}, {"SP", SP}, {"SPN", SPN}, {"SPA", SPA}};
    for (auto& p : tab) if (g == p.first) return p.second;
    return 0;
}

scgpu_iaparam TypeParams::pack() const {
    scgpu_iaparam r;
    memset(&r, 0, sizeof r);
    r.geotype[0] = geotype[0]; r.geotype[1] = geotype[1];
    r.exclude = exclude ? 1.0 : 0.0;
    r.sigma = sigma; r.epsilon = epsilon; r.A = A; r.B = B; r.pdis = pdis; r.pswitch = pswitch; r.pswitchINV = pswitchINV;
    r.rcut = rcut; r.rcutSq = rcutSq; r.rcutwca = rcutwca; r.rcutwcaSq = rcutwcaSq; r.parallel = parallel;
    for (int k = 0; k < 2; k++) {
        r.half_len[k] = half_len[k]; r.len[k] = len[k];
        r.csecpatchrot[k] = csecpatchrot[k]; r.ssecpatchrot[k] = ssecpatchrot[k];
        r.chiral_cos[k] = chiral_cos[k]; r.chiral_sin[k] = chiral_sin[k];
    }
    for (int k = 0; k < 4; k++) {
        r.pcangl[k] = pcangl[k]; r.pcanglsw[k] = pcanglsw[k]; r.pcoshalfi[k] = pcoshalfi[k]; r.psinhalfi[k] = psinhalfi[k];
    }
    return r;
}

Topology::Topology() : ia(MAXT, std::vector<TypeParams>(MAXT)) {}

// Inicializer::fillTypes (inicializer.cpp:660-852)
void Topology::fillType(const std::string& line) {
    auto tok = split(line);
    if (tok.size() < 5) throw Error("TOPOLOGY ERROR: in reading types: " + line);
    int type = atoi(tok[1].c_str());
    if (type < 0 || type >= MAXT) throw Error("TOPOLOGY ERROR: type number out of range 0-39: " + line);
    int g = convertGeotype(tok[2]);
    if (!g) throw Error("TOPOLOGY ERROR: Unknown GEOTYPE: " + tok[2]);
    double param[12] = {0};
    int fields = 0;
    for (size_t k = 3; k < tok.size() && k < 15; k++) { param[k - 3] = strtod(tok[k].c_str(), nullptr); fields++; }
    fields -= 2;   // EPSILON and SIGMA are not counted (inicializer.cpp:691)
    int need = -1;
    switch (g) {
        case SPN: need = 0; break; case SCN: need = 1; break; case SPA: need = 2; break; case SCA: need = 3; break;
        case PSC: case CPSC: need = 6; break; case CHPSC: case CHCPSC: need = 7; break;
        case TPSC: case TCPSC: need = 9; break; case TCHPSC: case TCHCPSC: need = 10; break;
    }
    if (fields != need) throw Error("TOPOLOGY ERROR: wrong number of parameters for " + tok[2]);
    TypeParams& p = ia[type][type];
    p.name = tok[0];
    p.geotype[0] = p.geotype[1] = g;
    p.epsilon = param[0];
    p.sigma = param[1];
    p.A = 4 * p.epsilon * pow(p.sigma, 12);
    p.B = 4 * p.epsilon * pow(p.sigma, 6);
    p.rcutwca = p.sigma * pow(2.0, 1.0 / 6.0);
    p.rcutwcaSq = p.rcutwca * p.rcutwca;
    if (fields > 0 && fields != 1 && fields != 3) {
        p.pdis = param[2];
        p.pswitch = param[3];
        p.pswitchINV = 1.0 / param[3];
        p.rcut = (g != SPN) ? p.pswitch + p.pdis : 0.0;
        p.rcutSq = p.rcut * p.rcut;
    }
    if (fields == 1) for (int i = 0; i < 2; i++) { p.len[i] = param[2]; p.half_len[i] = param[2] / 2; }
    if (fields == 3) for (int i = 0; i < 2; i++) { p.len[i] = param[4]; p.half_len[i] = param[4] / 2; }
    if (fields > 2 && fields != 3) {
        for (int i = 0; i < 2; i++) {
            p.len[i] = param[6];
            p.half_len[i] = param[6] / 2;
            p.pangl[i] = param[4];
            p.panglsw[i] = param[5];
            p.pcangl[i] = cos(param[4] / 2.0 / 180 * PI);
            p.pcanglsw[i] = cos((param[4] / 2.0 + param[5]) / 180 * PI);
            p.pcoshalfi[i] = cos((param[4] / 2.0 + param[5]) / 2.0 / 180 * PI);
            p.psinhalfi[i] = sqrt(1.0 - p.pcoshalfi[i] * p.pcoshalfi[i]);
            p.parallel = param[7];
        }
    }
    if (fields == 7) for (int i = 0; i < 2; i++) {
        p.chiral_cos[i] = cos(param[8] / 360 * PI);
        p.chiral_sin[i] = sqrt(1 - p.chiral_cos[i] * p.chiral_cos[i]);
    }
    if (fields == 9 || fields == 10) for (int i = 0; i < 2; i++) {
        p.csecpatchrot[i] = cos(param[8] / 360 * PI);
        p.ssecpatchrot[i] = sqrt(1 - p.csecpatchrot[i] * p.csecpatchrot[i]);
        p.pangl[i + 2] = param[9];
        p.panglsw[i + 2] = param[10];
        p.pcangl[i + 2] = cos(param[9] / 2.0 / 180 * PI);
        p.pcanglsw[i + 2] = cos((param[9] / 2.0 + param[10]) / 180 * PI);
        p.pcoshalfi[i + 2] = cos((param[9] / 2.0 + param[10]) / 2.0 / 180 * PI);
        p.psinhalfi[i + 2] = sqrt(1.0 - p.pcoshalfi[i + 2] * p.pcoshalfi[i + 2]);
    }
    if (fields == 10) for (int i = 0; i < 2; i++) {
        p.chiral_cos[i] = cos(param[11] / 360 * PI);
        p.chiral_sin[i] = sqrt(1 - p.chiral_cos[i] * p.chiral_cos[i]);
    }
    if (g < SP) p.volume = 4.0 / 3.0 * PI * pow(p.sigma / 2.0, 3.0) + PI / 2.0 * p.len[0] * pow(p.sigma / 2.0, 2.0);
    else p.volume = 4.0 / 3.0 * PI * pow(p.sigma / 2.0, 3.0);
    if (p.rcutwca > sqmaxcut) sqmaxcut = p.rcutwca;   // un-squared at this stage (inicializer.cpp:845-848)
    if (p.rcut > sqmaxcut) sqmaxcut = p.rcut;
}

// Inicializer::fillMol (inicializer.cpp:929-1108)
void Topology::fillMol(MoleculeType& mol, const std::string& line) {
    std::string body = line.substr(0, line.find('}'));
    size_t ob = body.find('{');
    if (ob != std::string::npos) body = body.substr(ob + 1);
    body = strip(body);
    if (body.empty()) return;
    size_t sep = body.find(':');
    std::string cmd = upper(strip(body.substr(0, sep)));
    auto vals = split(sep == std::string::npos ? "" : body.substr(sep + 1));
    if (cmd == "PARTICLES") {
        if (vals.empty()) throw Error("TOPOLOGY ERROR: could not read a pacticle.");
        int t = atoi(vals[0].c_str());
        if (t < 0 || t >= MAXT) throw Error("TOPOLOGY ERROR: pacticles include type out of range");
        mol.particleTypes.push_back(t);
        return;
    }
    if (cmd == "ACTIVITY") return;   // muVT is outside the hot path
    if (vals.size() < 2) throw Error("TOPOLOGY ERROR: wrong number of parameters for " + cmd);
    double k = strtod(vals[0].c_str(), nullptr), eq = strtod(vals[1].c_str(), nullptr);
    if (eq < 0) throw Error("TOPOLOGY ERROR: equilibrium value cannot be negative");
    if (cmd == "BOND1") { mol.bond1c = k; mol.bond1eq = eq; }
    else if (cmd == "BOND2") { mol.bond2c = k; mol.bond2eq = eq; }
    else if (cmd == "BONDD") { mol.bonddc = k; mol.bonddeq = eq; }
    else if (cmd == "BONDH") { mol.bondhc = k; mol.bondheq = eq; }
    else if (cmd == "ANGLE1") { mol.angle1c = k; mol.angle1eq = eq * DEGTORAD; }
    else if (cmd == "ANGLE2") { mol.angle2c = k; mol.angle2eq = eq * DEGTORAD; }
    else throw Error("TOPOLOGY ERROR: unknown parameter: " + cmd);
}

// Topo::genParamPairs (topo.cpp:5-138)
void Topology::genParamPairs() {
    double length = 0;
    for (int i = 0; i < MAXT; i++) {
        for (int j = 0; j < MAXT; j++) {
            if (i == j) continue;
            if (ia[j][j].geotype[0] == 0 || ia[i][i].geotype[0] == 0) continue;
            int a[2] = {i, j};
            TypeParams& q = ia[i][j];
            for (int k = 0; k < 2; k++) {
                const TypeParams& s = ia[a[k]][a[k]];
                q.geotype[k] = s.geotype[0];
                q.len[k] = s.len[0];
                if (s.len[0] > 0) {
                    if (length == 0) length = s.len[0];
                    else if (length != s.len[0]) throw Error("Error: Different lengths for spherocylinders have not been implemented yet!");
                }
                q.half_len[k] = s.half_len[0];
                if (q.geotype[k] >= PSC && q.geotype[k] < SP) {
                    q.pangl[k] = s.pangl[0];
                    q.panglsw[k] = s.panglsw[0];
                    q.pcangl[k] = cos(q.pangl[k] / 2.0 / 180 * PI);
                    q.pcanglsw[k] = cos((q.pangl[k] / 2.0 + q.panglsw[k]) / 180 * PI);
                    q.pcoshalfi[k] = cos((q.pangl[k] / 2.0 + q.panglsw[k]) / 2.0 / 180 * PI);
                    q.psinhalfi[k] = sqrt(1.0 - q.pcoshalfi[k] * q.pcoshalfi[k]);
                }
                int g = q.geotype[k];
                if (g == CHCPSC || g == CHPSC || g == TCHCPSC || g == TCHPSC) {
                    q.chiral_cos[k] = s.chiral_cos[0];
                    q.chiral_sin[k] = s.chiral_sin[0];
                }
                if (g == TCPSC || g == TPSC || g == TCHCPSC || g == TCHPSC) {
                    q.csecpatchrot[k] = s.csecpatchrot[0];
                    q.ssecpatchrot[k] = s.ssecpatchrot[0];
                    q.pangl[k + 2] = s.pangl[2];
                    q.panglsw[k + 2] = s.panglsw[2];
                    q.pcangl[k + 2] = cos(q.pangl[k + 2] / 2.0 / 180 * PI);
                    q.pcanglsw[k + 2] = cos((q.pangl[k + 2] / 2.0 + q.panglsw[k + 2]) / 180 * PI);
                    q.pcoshalfi[k + 2] = cos((q.pangl[k + 2] / 2.0 + q.panglsw[k + 2]) / 2.0 / 180 * PI);
                    q.psinhalfi[k + 2] = sqrt(1.0 - q.pcoshalfi[k + 2] * q.pcoshalfi[k + 2]);
                }
            }
            const TypeParams& pi = ia[i][i];
            const TypeParams& pj = ia[j][j];
            q.name = pi.name;
            q.sigma = (pi.sigma + pj.sigma) * 0.5;
            q.epsilon = sqrt(pi.epsilon * pj.epsilon);
            q.A = 4 * q.epsilon * pow(q.sigma, 12);
            q.B = 4 * q.epsilon * pow(q.sigma, 6);
            q.pswitch = (pi.pswitch + pj.pswitch) * 0.5;
            q.pswitchINV = 1.0 / q.pswitch;
            q.rcutwca = q.sigma * pow(2.0, 1.0 / 6.0);
            q.rcutwcaSq = q.rcutwca * q.rcutwca;
            if (pi.parallel > 0 && pj.parallel > 0) q.parallel = sqrt(pi.parallel * pj.parallel);
            if (pi.parallel < 0 && pj.parallel < 0) q.parallel = -sqrt(pi.parallel * pj.parallel);
            q.pdis = ((pi.pdis - pi.rcutwca) + (pj.pdis - pj.rcutwca)) * 0.5 + q.rcutwca;
            if (q.geotype[0] == SPN || q.geotype[1] == SPN) q.rcut = 0.0;
            else q.rcut = q.pswitch + q.pdis;
            q.rcutSq = q.rcut * q.rcut;
            if (q.rcutwca > sqmaxcut) sqmaxcut = q.rcutwca;
            if (q.rcut > sqmaxcut) sqmaxcut = q.rcut;
        }
    }
    for (int i = 0; i < MAXT; i++)
        for (int j = 0; j < MAXT; j++) ia[i][j].exclude = exclusions.count({i, j}) > 0;
}

// Topo::genTopoParams (topo.cpp:140-153)
void Topology::genTopoParams() {
    double maxlength = 0;
    for (int i = 0; i < MAXT; i++) if (maxlength < ia[i][i].len[0]) maxlength = ia[i][i].len[0];
    sqmaxcut += maxlength;
    sqmaxcut *= 1.1;
    maxcut = sqmaxcut;
    sqmaxcut = sqmaxcut * sqmaxcut;
}

// Inicializer::readTopoFile (inicializer.cpp:453-585)
Topology Topology::fromText(const std::string& text) {
    Topology t;
    std::istringstream in(text);
    std::string raw, key;
    MoleculeType* cur = nullptr;
    while (std::getline(in, raw)) {
        // continuation lines end with a backslash (simlib.cpp:54-64)
        std::string r = raw;
        while (true) {
            size_t e = r.find_last_not_of(" \t\r\n");
            if (e != std::string::npos && r[e] == '\\') {
                std::string more;
                r = r.substr(0, e);
                if (!std::getline(in, more)) break;
                r += more;
            } else break;
        }
        std::string line = strip(r);
        if (line.empty()) continue;
        if (line[0] == '[') {
            key = upper(strip(line.substr(1, line.find(']') - 1)));
            continue;
        }
        if (key == "TYPES") t.fillType(line);
        else if (key == "MOLECULES") {
            if (!cur) {
                if ((int)t.mols.size() >= MAXMT) throw Error("TOPOLOGY ERROR: too many molecule types");
                t.mols.emplace_back();
                cur = &t.mols.back();
                cur->name = strip(line.substr(0, line.find(':')));
            }
            t.fillMol(*cur, line);
            if (line.find('}') != std::string::npos) cur = nullptr;
        } else if (key == "SYSTEM") {
            auto tok = split(line);
            if (tok.size() < 2) throw Error("TOPOLOGY ERROR: failed reading system from (" + line + ")");
            t.system.emplace_back(tok[0], atol(tok[1].c_str()));
        } else if (key == "POOL") {
            // muVT pool: outside the hot path
        } else if (key == "EXTER") {
            auto tok = split(line);
            for (size_t k = 0; k < tok.size() && k < 3; k++) t.exter[k] = strtod(tok[k].c_str(), nullptr);
            t.exterExist = !tok.empty();
        } else if (key == "EXCLUDE") {
            auto tok = split(line);
            if (tok.size() % 2) throw Error("Error in readin Topology exclusions, probably there is not even number of types");
            for (size_t k = 0; k + 1 < tok.size(); k += 2) {
                int a = atoi(tok[k].c_str()), b = atoi(tok[k + 1].c_str());
                t.exclusions.insert({a, b});
                t.exclusions.insert({b, a});
            }
        } else throw Error("TOPOLOGY ERROR: invalid keyword:" + key);
    }
    t.genParamPairs();
    t.genTopoParams();
    return t;
}

int Topology::maxTypeInUse(const std::vector<int>& types) const {
    int m = 0;
    for (int t : types) if (t > m) m = t;
    return m;
}

// Vector::rotate(axis, cos, sin) (Vector.h:138-160)
static void rotateQ(const double* p, const double* axis, double cosAngle, double sinAngle, double* out) {
    double qw = cosAngle, qx = (axis[0] * sinAngle), qy = (axis[1] * sinAngle), qz = (axis[2] * sinAngle);
    double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy, t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t10 = -qz * qz;
    double x = p[0], y = p[1], z = p[2];
    out[0] = 2.0 * ((t8 + t10) * x + (t6 - t4) * y + (t3 + t7) * z) + x;
    out[1] = 2.0 * ((t4 + t6) * x + (t5 + t10) * y + (t9 - t2) * z) + y;
    out[2] = 2.0 * ((t7 - t3) * x + (t2 + t9) * y + (t5 + t8) * z) + z;
}
static void normalise(double* v) {   // Vector::normalise (Vector.h:56-64)
    double tot = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (tot != 0.0) { tot = 1.0 / tot; v[0] *= tot; v[1] *= tot; v[2] *= tot; }
}
static void ortogonalise(double* a, const double* b) {   // Vector::ortogonalise (Vector.h:133-135)
    double dp = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    a[0] -= dp * b[0]; a[1] -= dp * b[1]; a[2] -= dp * b[2];
}

// Particle::init (particle.cpp:3-79). state30: pos 0 | dir 3 | patchdir0 6 | patchdir1 9 | sides 12,15,18,21 | chdir 24,27
void particleInit(const TypeParams& ia, double* st) {
    int g = ia.geotype[0];
    if (g == SCA || g == SCN) return;
    double* dir = st + 3; double* pd0 = st + 6; double* pd1 = st + 9;
    normalise(dir);
    ortogonalise(pd0, dir);
    normalise(pd0);
    bool two = (g == TPSC || g == TCPSC || g == TCHPSC || g == TCHCPSC);
    bool chiral = (g == CHPSC || g == CHCPSC || g == TCHPSC || g == TCHCPSC);
    if (g == PSC || g == CPSC || g == TPSC || g == TCPSC) {
        rotateQ(pd0, dir, ia.pcoshalfi[0], ia.psinhalfi[0], st + 12);
        rotateQ(pd0, dir, ia.pcoshalfi[0], -1.0 * ia.psinhalfi[0], st + 15);
    }
    if (two) {
        double tmp[3];
        rotateQ(pd0, dir, ia.csecpatchrot[0], ia.ssecpatchrot[0], tmp);
        ortogonalise(tmp, dir);
        normalise(tmp);
        pd1[0] = tmp[0]; pd1[1] = tmp[1]; pd1[2] = tmp[2];
    }
    if (g == TPSC || g == TCPSC) {
        rotateQ(pd1, dir, ia.pcoshalfi[2], ia.psinhalfi[2], st + 18);
        rotateQ(pd1, dir, ia.pcoshalfi[2], -1.0 * ia.psinhalfi[2], st + 21);
    }
    if (chiral) {
        rotateQ(dir, pd0, ia.chiral_cos[0], ia.chiral_sin[0], st + 24);
        rotateQ(pd0, st + 24, ia.pcoshalfi[0], ia.psinhalfi[0], st + 12);
        rotateQ(pd0, st + 24, ia.pcoshalfi[0], -1.0 * ia.psinhalfi[0], st + 15);
    }
    if (g == TCHPSC || g == TCHCPSC) {
        rotateQ(dir, pd1, ia.chiral_cos[0], ia.chiral_sin[0], st + 27);
        rotateQ(pd1, st + 27, ia.pcoshalfi[2], ia.psinhalfi[2], st + 18);
        rotateQ(pd1, st + 27, ia.pcoshalfi[2], -1.0 * ia.psinhalfi[2], st + 21);
    }
}

void System::initParticle(int i) {
    const TypeParams& self = topo.ia[type[i]][type[i]];
    if (self.geotype[0] < SP) particleInit(self, &state[(size_t)i * 30]);
}

static double usePBC(double v) {   // Cuboid::usePBC(Vector&) (geometry.h:79-98)
    while (v < 0.0) v += 1.0;
    while (v > 1.0) v -= 1.0;
    return v;
}

System System::fromText(const std::string& topText, const std::string& configText, const std::vector<long>& countsOverride) {
    System s;
    s.topo = Topology::fromText(topText);
    Topology& t = s.topo;
    if (!countsOverride.empty() && countsOverride.size() != t.system.size())
        throw Error("counts override must have one entry per [System] line");
    // Inicializer::setParticlesParamss (inicializer.cpp:409-450)
    for (size_t si = 0; si < t.system.size(); si++) {
        int mol = -1;
        for (size_t m = 0; m < t.mols.size(); m++) if (t.mols[m].name == t.system[si].first) { mol = (int)m; break; }
        if (mol < 0) throw Error("TOPOLOGY ERROR: molecules " + t.system[si].first + " is not defined.");
        long cnt = countsOverride.empty() ? t.system[si].second : countsOverride[si];
        for (long c = 0; c < cnt; c++)
            for (int ty : t.mols[mol].particleTypes) { s.type.push_back(ty); s.moltype.push_back(mol); }
    }
    s.n = (int)s.type.size();
    if (s.n == 0) throw Error("TOPOLOGY ERROR: no particles in [System]");
    // Inicializer::initGroupLists (inicializer.cpp:330-372): first index of every molecule type
    int nmol = (int)t.mols.size();
    s.first.assign(nmol + 1, 0);
    {
        int i = 0;
        for (int m = 0; m < nmol; m++) {
            while (i < s.n && s.moltype[i] < m) i++;
            s.first[m] = i;
            while (i < s.n && s.moltype[i] == m) i++;
        }
        s.first[nmol] = s.n;
        for (int i2 = 1; i2 < s.n; i2++)
            if (s.moltype[i2] < s.moltype[i2 - 1]) throw Error("TOPOLOGY ERROR: [System] must list molecule types in the order they are defined");
    }
    // Inicializer::initConfig (inicializer.cpp:89-240)
    std::istringstream in(configText);
    std::string raw, line;
    while (std::getline(in, raw)) { line = strip(raw); if (!line.empty()) break; }
    {
        auto tok = split(line);
        if (tok.size() < 3) throw Error("ERROR: Could not read box size (Inicializer::initConfig)");
        for (int d = 0; d < 3; d++) s.box[d] = strtod(tok[d].c_str(), nullptr);
    }
    s.state.assign((size_t)s.n * 30, 0.0);
    s.switched.assign(s.n, 0);
    for (int i = 0; i < s.n; i++) {
        if (!std::getline(in, raw)) throw Error("ERROR: Could not read coordinates for particle " + std::to_string(i + 1));
        auto tok = split(strip(raw));
        if (tok.size() < 9) throw Error("ERROR: Could not read coordinates for particle " + std::to_string(i + 1));
        double v[9];
        for (int k = 0; k < 9; k++) v[k] = strtod(tok[k].c_str(), nullptr);
        if (tok.size() >= 10) s.switched[i] = atoi(tok[9].c_str());
        double* st = &s.state[(size_t)i * 30];
        for (int d = 0; d < 3; d++) st[d] = usePBC(v[d] / s.box[d]);
        int g = t.ia[s.type[i]][s.type[i]].geotype[0];
        st[3] = v[3]; st[4] = v[4]; st[5] = v[5];
        st[6] = v[6]; st[7] = v[7]; st[8] = v[8];
        if (g < SP && (st[3] * st[3] + st[4] * st[4] + st[5] * st[5]) < 1.0e-12)
            throw Error("ERROR: Null direction vector supplied for particle " + std::to_string(i + 1));
        normalise(st + 3);
        if (g < SP && g != SCN && (st[6] * st[6] + st[7] * st[7] + st[8] * st[8]) < 1.0e-12)
            throw Error("ERROR: Null patch vector supplied for particle " + std::to_string(i + 1));
        ortogonalise(st + 6, st + 3);
        normalise(st + 6);
    }
    // make chains whole (inicializer.cpp:268-278, Conf.h:371-380)
    for (int i = 0; i < s.n;) {
        int msz = t.mols[s.moltype[i]].molSize();
        if (msz > 1) {
            for (int k = i + 1; k < i + msz && k < s.n; k++) {
                for (int d = 0; d < 3; d++) {
                    double r = s.state[(size_t)k * 30 + d] - s.state[(size_t)(k - 1) * 30 + d];
                    r = s.box[d] * (r - rint(r));
                    r /= s.box[d];
                    s.state[(size_t)k * 30 + d] = s.state[(size_t)(k - 1) * 30 + d] + r;
                }
            }
        }
        i += msz;
    }
    // Conf::partVecInit (Conf.cpp:98-103)
    for (int i = 0; i < s.n; i++) s.initParticle(i);
    // packed tables, indexed by the reference's type numbers
    s.ntypes = t.maxTypeInUse(s.type) + 1;
    s.iaTable.resize((size_t)s.ntypes * s.ntypes);
    for (int a = 0; a < s.ntypes; a++)
        for (int b = 0; b < s.ntypes; b++) s.iaTable[(size_t)a * s.ntypes + b] = t.ia[a][b].pack();
    s.molTable.resize(nmol);
    for (int m = 0; m < nmol; m++) {
        const MoleculeType& q = t.mols[m];
        scgpu_molparam r;
        memset(&r, 0, sizeof r);
        r.bond1eq = q.bond1eq; r.bond1c = q.bond1c; r.bond2eq = q.bond2eq; r.bond2c = q.bond2c;
        r.bonddeq = q.bonddeq; r.bonddc = q.bonddc; r.bondheq = q.bondheq; r.bondhc = q.bondhc;
        r.angle1eq = q.angle1eq; r.angle1c = q.angle1c; r.angle2eq = q.angle2eq; r.angle2c = q.angle2c;
        r.mol_size = q.molSize(); r.first = s.first[m];
        s.molTable[m] = r;
    }
    return s;
}

static std::string slurp(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw Error("ERROR: Could not open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

System System::fromFiles(const std::string& topPath, const std::string& configPath) {
    return fromText(slurp(topPath), slurp(configPath));
}

// main.cpp:304-311 + Conf::draw (Conf.h:388-425)
std::string System::configLast(bool testingFormat) const {
    std::string out;
    char buf[512];
    const char* f3 = testingFormat ? "%15.6e %15.6e %15.6e\n" : "%15.8e %15.8e %15.8e\n";
    snprintf(buf, sizeof buf, f3, box[0], box[1], box[2]);
    out += buf;
    for (int i = 0; i < n; i++) {
        const double* st = &state[(size_t)i * 30];
        double p[3];
        for (int d = 0; d < 3; d++) p[d] = box[d] * (st[d] - rint(st[d]));
        if (testingFormat)
            snprintf(buf, sizeof buf, "%15.6e %15.6e %15.6e   %15.6e %15.6e %15.6e   %15.6e %15.6e %15.6e %d\n",
                     p[0], p[1], p[2], st[3], st[4], st[5], st[6], st[7], st[8], switched[i]);
        else
            snprintf(buf, sizeof buf, "%15.8e %15.8e %15.8e   %15.8e %15.8e %15.8e   %15.8e %15.8e %15.8e %d %d\n",
                     p[0], p[1], p[2], st[3], st[4], st[5], st[6], st[7], st[8], switched[i], moltype[i]);
        out += buf;
    }
    return out;
}

}  // namespace schost
