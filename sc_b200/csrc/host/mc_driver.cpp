// Sequential Monte-Carlo driver on top of TotalEGpu: the CALLER side of the hot path, mirrored from the reference so that a
// whole reference run (options + top.init + config.init -> config.last) can be reproduced through the CUDA energy path:
//   Sim::readOptions            scOOP/structures/sim.h:123-406          (the keys the tests use; step sizes, probabilities)
//   Ran2                        scOOP/mc/randomGenerator.cpp:20-54      (selected by -DTESTING; as written, incl. the un-warmed start)
//   Updater::simulate           scOOP/mc/updater.cpp:45-389             (step loop; production run only)
//   MoveCreator::particleMove / partDisplace / partRotate       scOOP/mc/movecreator.cpp:11-33, 947-1028
//   MoveCreator::chainMove / chainDisplace / chainRotate / clusterRotate   :304-328, 1075-1392
//   MoveCreator::pressureMove (ptype 0-5) / moveTry             :330-550, movecreator.h:175-187
//   Particle::pscRotate, Vector::randomUnitSphere               scOOP/structures/particle.h:182-272, Vector.h:190-205
// Not mirrored (outside the hot path's callers this round): Wang-Landau, muVT, cluster and switch moves, the wall
// potential, step-size adaptation (adjust/nequil must be 0, as in every Tests/test_* input).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>

#include "calculator.hpp"

namespace schost {

static const double PIH = 1.57079632679489661923132169163975;

class Ran2 {   // Numerical Recipes ran2 as the reference calls it: a POSITIVE seed skips the warm-up branch entirely
    long seed;
    long idum2 = 123456789, iy = 0, iv[32];
public:
    explicit Ran2(long s) : seed(s) { for (auto& v : iv) v = 0; }
    double operator()() {
        const long IM1 = 2147483563, IM2 = 2147483399, IMM1 = IM1 - 1, IA1 = 40014, IA2 = 40692, IQ1 = 53668, IQ2 = 52774, IR1 = 12211, IR2 = 3791;
        const int NTAB = 32;
        const long NDIV = 1 + IMM1 / NTAB;
        const double AM = 1.0 / IM1, RNMX = 1.0 - 1.2e-7;
        long k;
        int j;
        if (seed <= 0) {
            if (-(seed) < 1) seed = 1; else seed = -(seed);
            idum2 = seed;
            for (j = NTAB + 7; j >= 0; j--) {
                k = seed / IQ1;
                seed = IA1 * (seed - k * IQ1) - k * IR1;
                if (seed < 0) seed += IM1;
                if (j < NTAB) iv[j] = seed;
            }
            iy = iv[0];
        }
        k = seed / IQ1;
        seed = IA1 * (seed - k * IQ1) - k * IR1;
        if (seed < 0) seed += IM1;
        k = idum2 / IQ2;
        idum2 = IA2 * (idum2 - k * IQ2) - k * IR2;
        if (idum2 < 0) idum2 += IM2;
        j = (int)(iy / NDIV);
        iy = iv[j] - idum2;
        iv[j] = seed;
        if (iy < 1) iy += IMM1;
        double temp = AM * iy;
        return temp > RNMX ? RNMX : temp;
    }
};

struct Options {
    std::map<std::string, double> v;
    double get(const char* k, double dflt = 0.0) const { auto it = v.find(k); return it == v.end() ? dflt : it->second; }
    static Options parse(const std::string& text) {   // key = value # comment (sim.h:176-230)
        Options o;
        std::istringstream in(text);
        std::string line;
        while (std::getline(in, line)) {
            line = line.substr(0, line.find('#'));
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string key = line.substr(0, eq), val = line.substr(eq + 1);
            auto trim = [](std::string s) {
                size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
                return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
            };
            key = trim(key);
            val = trim(val);
            if (key.empty() || val.empty()) continue;
            o.v[key] = strtod(val.c_str(), nullptr);
        }
        return o;
    }
};

struct McStats {
    long trans_acc = 0, trans_rej = 0, rot_acc = 0, rot_rej = 0, chainm_acc = 0, chainm_rej = 0, chainr_acc = 0, chainr_rej = 0, edge_acc = 0, edge_rej = 0;
    double e_start = 0, e_end = 0, drift = 0;
};

class SequentialMC {
    System* conf;
    TotalEGpu* calc;
    Ran2 ran2;
    double temper, press, shprob, chainprob;
    int ptype;
    // Disp (structures/statistics.h:55-71): step size + acceptance counters per particle type / molecule type, adapted during
    // equilibration by optimizeStep / optimizeRot (mc/updater.cpp:395-465)
    struct Disp {
        double mx = 0.0, angle = 0.0, oldrmsd = 0.0, oldmx = 0.0;
        long acc = 0, rej = 0;
        double ratio() const { return (acc + rej) > 0 ? 1.0 * acc / (acc + rej) : 0.0; }
    };
    Disp trans[40], rot[40], chainm[100], chainr[100], edge;      // MAXT, MAXMT (structures/macros.h:87-88)
    long nequil = 0, adjust = 0;
    std::vector<std::pair<int, int>> chains;   // (first index, size) of every molecule with more than one particle
public:
    McStats st;
    SequentialMC(System* c, TotalEGpu* calc_, const Options& o)
        : conf(c), calc(calc_), ran2((long)o.get("seed")) {
        nequil = (long)o.get("nequil"); adjust = (long)o.get("adjust");
        // (nrepchange is not an obstacle: without MPI the reference's replicaExchangeMove is an empty function, movecreator.cpp:552-795)
        if (o.get("wlm") != 0 || o.get("nGrandCanon") != 0 || o.get("nClustMove") != 0 || o.get("switchprob") != 0)
            throw Error("sequential driver: Wang-Landau / muVT / cluster / switch moves are outside the mirrored callers");
        temper = o.get("temper"); press = o.get("press"); ptype = (int)o.get("ptype");
        for (int i = 0; i < 40; i++) {                          // sim.h:358-374
            trans[i].mx = 2.0 * o.get("transmx");
            rot[i].mx = cos(o.get("rotmx") / 180.0 * PIH);      // the cosine is what optimizeRot adapts; the moves read .angle, which it never touches
            rot[i].angle = o.get("rotmx") / 180.0 * PIH * 0.5;
        }
        for (int i = 0; i < 100; i++) {                         // sim.h:376-382
            chainm[i].mx = 2.0 * o.get("chainmmx");
            chainr[i].mx = cos(o.get("chainrmx") / 180.0 * PIH);
            chainr[i].angle = o.get("chainrmx") / 180.0 * PIH;
        }
        edge.mx = 2.0 * o.get("edge_mx");                       // sim.h:364
        for (int i = 0; i < conf->n;) {
            int msz = conf->topo.mols[conf->moltype[i]].molSize();
            if (msz > 1) chains.emplace_back(i, msz);
            i += msz;
        }
        chainprob = chains.empty() ? 0.0 : o.get("chainprob");   // Inicializer::testChains (inicializer.cpp:301-306)
        shprob = o.get("shave") / (double)conf->n;              // updater.cpp:87-91
    }

    double* P(int i) { return &conf->state[(size_t)i * 30]; }
    int geotype(int i) const { return conf->topo.ia[conf->type[i]][conf->type[i]].geotype[0]; }

    void randomUnitSphere(double v[3]) {      // Vector.h:190-205
        double a, xi1, xi2;
        do {
            xi1 = 1.0 - 2.0 * ran2();
            xi2 = 1.0 - 2.0 * ran2();
            a = xi1 * xi1 + xi2 * xi2;
        } while (a > 1.0);
        double b = 2.0 * sqrt(1.0 - a);
        v[0] = xi1 * b; v[1] = xi2 * b; v[2] = 1.0 - 2.0 * a;
    }
    bool moveTry(double eold, double enew) {   // true = REJECT (movecreator.h:175-187)
        if (enew <= eold) return false;
        return !(exp(-1.0 * (enew - eold) / temper) > ran2());
    }
    static void rotVec(double* p, const double d[9]) {
        double x = p[0], y = p[1], z = p[2];
        p[0] = 2.0 * (d[0] * x + d[1] * y + d[2] * z) + x;
        p[1] = 2.0 * (d[3] * x + d[4] * y + d[5] * z) + y;
        p[2] = 2.0 * (d[6] * x + d[7] * y + d[8] * z) + z;
    }
    static void quatMatrix(double qw, double qx, double qy, double qz, double d[9]) {   // particle.h:203-221 / Vector.h:162-183
        double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy, t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t10 = -qz * qz;
        d[0] = t8 + t10; d[1] = t6 - t4; d[2] = t3 + t7; d[3] = t4 + t6; d[4] = t5 + t10; d[5] = t9 - t2; d[6] = t7 - t3; d[7] = t2 + t9; d[8] = t5 + t8;
    }

    double partDisplace(int target) {          // movecreator.cpp:947-994
        double orig[3] = {P(target)[0], P(target)[1], P(target)[2]};
        double energy = calc->oneToAll(target);
        double dr[3];
        randomUnitSphere(dr);
        Disp& ds = trans[conf->type[target]];
        dr[0] *= ds.mx / conf->box[0]; dr[1] *= ds.mx / conf->box[1]; dr[2] *= ds.mx / conf->box[2];
        P(target)[0] += dr[0]; P(target)[1] += dr[1]; P(target)[2] += dr[2];
        double enermove = calc->oneToAllTrial(target);
        if (moveTry(energy, enermove)) {
            P(target)[0] = orig[0]; P(target)[1] = orig[1]; P(target)[2] = orig[2];
            st.trans_rej++; ds.rej++;
            return 0.0;
        }
        st.trans_acc++; ds.acc++;
        calc->update(target);
        return enermove - energy;
    }

    void pscRotate(double* p, double angle, int g, const double axis[3]) {   // particle.h:182-272, clockwise == 2
        double vc = cos(angle), vs;
        if (ran2() < 0.5) vs = sqrt(1.0 - vc * vc); else vs = -sqrt(1.0 - vc * vc);
        double d[9];
        quatMatrix(vc, axis[0] * vs, axis[1] * vs, axis[2] * vs, d);
        rotVec(p + 3, d);
        if (g != 10 && g != 11) {
            int m = (g == 16 || g == 17 || g == 18 || g == 19) ? 2 : 1;
            for (int k = 0; k < m; k++) { rotVec(p + 6 + 3 * k, d); rotVec(p + 12 + 6 * k, d); rotVec(p + 15 + 6 * k, d); }
        }
        if (g == 14 || g == 15 || g == 18 || g == 19) {
            int m = (g == 18 || g == 19) ? 2 : 1;
            for (int k = 0; k < m; k++) rotVec(p + 24 + 3 * k, d);
        }
    }

    double partRotate(int target) {            // movecreator.cpp:996-1028
        double orig[30];
        memcpy(orig, P(target), sizeof orig);
        double energy = calc->oneToAll(target);
        // rotateRandom: pscRotate(max_angle * ran2(), geotype, Vector::getRandomUnitSphere()) -- the reference binary (GCC, x86-64)
        // evaluates the arguments right to left: the axis is drawn BEFORE the angle (particle.h:171-173)
        double axis[3];
        randomUnitSphere(axis);
        Disp& ds = rot[conf->type[target]];
        double angle = ds.angle * ran2();
        pscRotate(P(target), angle, geotype(target), axis);
        {   // patchdir[0].ortogonalise(dir)
            double* p = P(target);
            double dp = p[6] * p[3] + p[7] * p[4] + p[8] * p[5];
            p[6] -= dp * p[3]; p[7] -= dp * p[4]; p[8] -= dp * p[5];
        }
        double enermove = calc->oneToAllTrial(target);
        if (moveTry(energy, enermove)) {
            memcpy(P(target), orig, sizeof orig);
            st.rot_rej++; ds.rej++;
            return 0.0;
        }
        st.rot_acc++; ds.acc++;
        calc->update(target);
        return enermove - energy;
    }

    double particleMove() {                    // movecreator.cpp:11-33
        int target = (int)(ran2() * (long)conf->n);
        if ((ran2() < 0.5) || geotype(target) >= 30) return partDisplace(target);
        return partRotate(target);
    }

    double chainDisplace(int ch) {             // movecreator.cpp:1075-1136
        Molecule mol;
        for (int k = 0; k < chains[ch].second; k++) mol.push_back(chains[ch].first + k);
        std::vector<double> orig(mol.size() * 3);
        for (size_t k = 0; k < mol.size(); k++) memcpy(&orig[3 * k], P(mol[k]), 3 * sizeof(double));
        double energy = calc->mol2others(mol);
        double dr[3];
        randomUnitSphere(dr);
        Disp& ds = chainm[conf->moltype[mol[0]]];
        dr[0] *= ds.mx / conf->box[0]; dr[1] *= ds.mx / conf->box[1]; dr[2] *= ds.mx / conf->box[2];
        for (int i : mol) { P(i)[0] += dr[0]; P(i)[1] += dr[1]; P(i)[2] += dr[2]; }
        double enermove = calc->mol2othersTrial(mol);
        if (moveTry(energy, enermove)) {
            for (size_t k = 0; k < mol.size(); k++) memcpy(P(mol[k]), &orig[3 * k], 3 * sizeof(double));
            st.chainm_rej++; ds.rej++;
            return 0.0;
        }
        st.chainm_acc++; ds.acc++;
        calc->update(mol);
        return enermove - energy;
    }

    double chainRotate(int ch) {               // movecreator.cpp:1138-1256, clusterRotate :1339-1392
        Molecule mol;
        for (int k = 0; k < chains[ch].second; k++) mol.push_back(chains[ch].first + k);
        std::vector<double> orig(mol.size() * 30);
        for (size_t k = 0; k < mol.size(); k++) memcpy(&orig[30 * k], P(mol[k]), 30 * sizeof(double));
        double energy = calc->mol2others(mol);
        double cm[3] = {0, 0, 0}, vol = 0.0;
        for (int i : mol) {
            double v = conf->topo.ia[conf->type[i]][conf->type[i]].volume;
            cm[0] += P(i)[0] * v; cm[1] += P(i)[1] * v; cm[2] += P(i)[2] * v;
            vol += v;
        }
        cm[0] /= vol; cm[1] /= vol; cm[2] /= vol;
        double axis[3];
        randomUnitSphere(axis);
        Disp& ds = chainr[conf->moltype[mol[0]]];
        double vc = cos(ds.angle * ran2()), vs;
        if (ran2() < 0.5) vs = sqrt(1.0 - vc * vc); else vs = -sqrt(1.0 - vc * vc);
        double d[9];
        quatMatrix(vc, axis[0] * vs, axis[1] * vs, axis[2] * vs, d);
        for (int i : mol) {
            double* p = P(i);
            for (int k = 0; k < 3; k++) { p[k] -= cm[k]; p[k] *= conf->box[k]; }
            rotVec(p, d);        // pos
            rotVec(p + 3, d);    // dir
            rotVec(p + 6, d); rotVec(p + 9, d);      // patchdir[0], [1]
            rotVec(p + 24, d); rotVec(p + 27, d);    // chdir[0], [1]
            rotVec(p + 12, d); rotVec(p + 15, d); rotVec(p + 18, d); rotVec(p + 21, d);   // patchsides
            for (int k = 0; k < 3; k++) { p[k] /= conf->box[k]; p[k] += cm[k]; }
        }
        double enermove = calc->mol2othersTrial(mol);
        if (moveTry(energy, enermove)) {
            for (size_t k = 0; k < mol.size(); k++) memcpy(P(mol[k]), &orig[30 * k], 30 * sizeof(double));
            st.chainr_rej++; ds.rej++;
            return 0.0;
        }
        st.chainr_acc++; ds.acc++;
        calc->update(mol);
        return enermove - energy;
    }

    double chainMove() {                       // movecreator.cpp:304-328
        if (chains.empty()) return 0.0;
        int target = (int)(ran2() * (double)chains.size());
        if (ran2() < 0.5) return chainDisplace(target);
        return chainRotate(target);
    }

    double pressureMove() {                    // movecreator.cpp:330-550, ptype 0-5
        double energy = calc->allToAll();
        double enermove = 0.0;
        auto& box = conf->box;
        const double N = (double)conf->n;
        bool reject;
        if (ptype == 0) {
            double rsave = ran2();
            int side;
            double area;
            if (rsave < 1.0 / 3.0) { side = 0; area = box[1] * box[2]; }
            else if (rsave < 2.0 / 3.0) { side = 1; area = box[0] * box[2]; }
            else { side = 2; area = box[0] * box[1]; }
            double old_side = box[side];
            box[side] += edge.mx * (ran2() - 0.5);
            enermove = press * area * (box[side] - old_side) - N * temper * log(box[side] / old_side);
            enermove += calc->allToAllTrial();
            reject = box[side] <= 0.0 || moveTry(energy, enermove);
            if (reject) box[side] = old_side;
        } else if (ptype == 1) {
            double psch = edge.mx * (ran2() - 0.5);
            double pvol = box[0] * box[1] * box[2];
            box[0] += psch; box[1] += psch; box[2] += psch;
            double pvoln = box[0] * box[1] * box[2];
            enermove = press * (pvoln - pvol) - N * temper * log(pvoln / pvol);
            enermove += calc->allToAllTrial();
            reject = moveTry(energy, enermove);
            if (reject) { box[0] -= psch; box[1] -= psch; box[2] -= psch; }
        } else if (ptype == 2) {
            double psch = edge.mx * (ran2() - 0.5);
            double pvol = box[0] * box[1];
            box[0] += psch; box[1] += psch;
            double pvoln = box[0] * box[1];
            enermove = press * box[2] * (pvoln - pvol) - N * temper * log(pvoln / pvol);
            enermove += calc->allToAllTrial();
            reject = moveTry(energy, enermove);
            if (reject) { box[0] -= psch; box[1] -= psch; }
        } else if (ptype == 3) {
            double psch = edge.mx * (ran2() - 0.5);
            double pvol = box[0] * box[1] * box[2];
            box[0] += psch; box[1] += psch;
            box[2] = pvol / box[0] / box[1];
            enermove += calc->allToAllTrial();
            reject = moveTry(energy, enermove);
            if (reject) { box[0] -= psch; box[1] -= psch; box[2] = pvol / box[0] / box[1]; }
        } else if (ptype == 4) {           // :480-514 "anisotropic in xy, z const". As written `if (ran2() - 0.5)` tests a non-zero double: always
            double pvol = box[0] * box[1];       // true, so only the x edge ever changes -- two random numbers are drawn all the same
            double psx = 0.0, psy = 0.0;
            if (ran2() - 0.5) psx = edge.mx * (ran2() - 0.5);
            else psy = edge.mx * (ran2() - 0.5);
            box[0] += psx; box[1] += psy;
            double pvoln = box[0] * box[1];
            enermove = press * box[2] * (pvoln - pvol) - N * temper * log(pvoln / pvol);
            enermove += calc->allToAllTrial();
            reject = moveTry(energy, enermove);
            if (reject) { box[0] -= psx; box[1] -= psy; }
        } else if (ptype == 5) {           // :515-541 box change along y only
            double pvol = box[1];
            double psy = edge.mx * (ran2() - 0.5);
            box[1] += psy;
            double pvoln = box[1];
            enermove = press * box[2] * box[0] * (pvoln - pvol) - N * temper * log(pvoln / pvol);
            enermove += calc->allToAllTrial();
            reject = moveTry(energy, enermove);
            if (reject) box[1] -= psy;
        } else throw Error("sequential driver: unknown type of pressure coupling");
        if (reject) { st.edge_rej++; edge.rej++; return 0.0; }     // the calculator re-reads the restored box on its next call
        st.edge_acc++; edge.acc++;
        calc->update();
        return enermove - energy;
    }

    static void optimizeStep(Disp& x, double hi, double lo) {      // updater.cpp:395-429
        const double newrmsd = x.mx * x.ratio();
        if (x.oldrmsd > 0) {
            if (newrmsd < x.oldrmsd) {
                if (x.oldmx > 1) { x.mx /= 1.05; x.oldmx = 0.95; } else { x.mx *= 1.05; x.oldmx = 1.05; }
            } else {
                if (x.oldmx > 1) { x.mx *= 1.05; x.oldmx = 1.05; } else { x.mx /= 1.05; x.oldmx = 0.95; }
            }
        }
        if (newrmsd > 0) x.oldrmsd = newrmsd;
        else { x.oldrmsd = 0.0; x.mx /= 1.05; x.oldmx = 0.95; }
        if (x.mx > hi) x.mx = hi;
        if (x.mx < lo) x.mx = lo;
        x.acc = x.rej = 0;
    }
    static void optimizeRot(Disp& x, double hi, double lo) {       // updater.cpp:431-465 (adapts the cosine .mx; the moves read .angle)
        const double newrmsd = x.mx * x.ratio();
        if (x.oldrmsd > 0) {
            if (newrmsd > x.oldrmsd) {
                if (x.oldmx > 1) { x.mx *= 0.99; x.oldmx *= 0.99; } else { x.mx *= 1.01; x.oldmx *= 1.01; }
            } else {
                if (x.oldmx > 1) { x.mx *= 1.01; x.oldmx *= 1.01; } else { x.mx *= 0.99; x.oldmx *= 0.99; }
            }
        }
        if (newrmsd > 0) x.oldrmsd = newrmsd;
        else { x.oldrmsd = 0.0; x.mx *= 1.01; x.oldmx = 1.01; }
        if (x.mx > hi) x.mx = hi;
        if (x.mx < lo) x.mx = lo;
        x.acc = x.rej = 0;
    }

    // one Updater::simulate call (updater.cpp:45-389): `adjust` > 0 adapts the step sizes every `adjust` sweeps (:238-253)
    void simulate(long nsweeps, long adjust_every = 0, bool reinit = false) {
        double edriftchanges = 0.0;
        if (reinit) for (int i = 0; i < conf->n; i++) conf->initParticle(i);      // Updater::initValues -> partVecInit (:31)
        calc->initEM();
        st.e_start = calc->allToAll();
        long next_adjust = adjust_every;
        for (long sweep = 1; sweep <= nsweeps; sweep++) {
            for (long step = 1; step <= (long)conf->n; step++) {
                double moveprobab = ran2();
                if (moveprobab < shprob) { edriftchanges += pressureMove(); continue; }
                if (moveprobab < shprob + chainprob) { edriftchanges += chainMove(); continue; }
                edriftchanges += particleMove();
            }
            if (sweep == next_adjust) {
                for (int i = 0; i < 40; i++) {
                    if (trans[i].acc > 0 || trans[i].rej > 0) optimizeStep(trans[i], 1.5, 0.0);
                    if (rot[i].acc > 0 || rot[i].rej > 0) optimizeRot(rot[i], 5.0, 0.01);
                }
                for (int i = 0; i < 100; i++) {
                    if (chainm[i].acc > 0 || chainm[i].rej > 0) optimizeStep(chainm[i], 1.5, 0.0);
                    if (chainr[i].acc > 0 || chainr[i].rej > 0) optimizeRot(chainr[i], 5.0, 0.01);
                }
                optimizeStep(edge, 1.0, 0.0);
                next_adjust += adjust_every;
            }
            if (!(sweep % 100000)) for (int i = 0; i < conf->n; i++) conf->initParticle(i);
        }
        st.e_end = calc->allToAll();
        st.drift = st.e_end - st.e_start - edriftchanges;
    }

    // main.cpp:270-297: equilibration (first half adapting the step sizes, second half at the adapted sizes), then production
    void run(long nsweeps) {
        bool again = false;
        if (nequil) {
            simulate(nequil / 2, adjust, false);
            simulate(nequil / 2, 0, true);
            again = true;
        }
        simulate(nsweeps, 0, again);
    }
    double transMx(int type) const { return trans[type].mx; }
};

}  // namespace schost

// ---- C entry points ----
using namespace schost;
static thread_local std::string g_mc_err;

extern "C" {

const char* schost_mc_last_error(void) { return g_mc_err.c_str(); }

// runs the production part of a reference run on `sys` (state is updated in place); stats10: acc/rej pairs + energies
int schost_run_mc(void* sys, const char* options_text, int device, long nsweeps_override, double* out13) {
    try {
        System* s = (System*)sys;
        Options o = Options::parse(options_text);
        TotalEGpu calc(s, device);
        SequentialMC mc(s, &calc, o);
        long ns = nsweeps_override > 0 ? nsweeps_override : (long)o.get("nsweeps");
        mc.run(ns);
        const McStats& t = mc.st;
        double v[13] = {(double)t.trans_acc, (double)t.trans_rej, (double)t.rot_acc, (double)t.rot_rej, (double)t.chainm_acc, (double)t.chainm_rej,
                        (double)t.chainr_acc, (double)t.chainr_rej, (double)t.edge_acc, (double)t.edge_rej, t.e_start, t.e_end, t.drift};
        memcpy(out13, v, sizeof v);
        return 0;
    } catch (const std::exception& e) { g_mc_err = e.what(); return -1; }
}

// config.last text of the current state (main.cpp:304-311); returns the length, copies at most cap bytes
long schost_config_last(void* sys, int testing_format, char* buf, long cap) {
    System* s = (System*)sys;
    std::string t = s->configLast(testing_format != 0);
    long n = (long)t.size();
    if (buf && cap > 0) memcpy(buf, t.data(), (size_t)(n < cap ? n : cap));
    return n;
}

}  // extern "C"
