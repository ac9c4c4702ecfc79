// TEST INFRASTRUCTURE -- not part of the product.
//
// Reference driver: a small main() of OUR OWN that links the UNMODIFIED reference sources where
// they lie under /root/reference/scOOP (everything except the reference's main.cpp) and calls the
// reference's own classes -- Sim, Inicializer, Conf, PairE, ParticleVector::getConlist,
// Conf::overlap -- to dump, at full precision (C99 hex floats), what the hot path computes on the
// configuration found in the current directory (options / top.init / config.init):
//
//   * the derived per-particle state after Conf::partVecInit()   (particle.cpp:3-79)
//   * the interaction table entries in use                         (topo.cpp:5-153)
//   * every non-zero PairE(i,j,conlist(i)) in BOTH orders          (paire.h:1209-1220)
//   * per-particle sums in TotalEFull::oneToAll order              (totalenergycalculator.h:563-583)
//   * the total in TotalEMatrix::allToAll(matrix) order            (totalenergycalculator.h:502-521)
//   * Conf::overlap(i,j) flags                                     (Conf.cpp:104-239)
//
// It is built ONLY by oracle/Makefile into oracle/_ref/ (git-ignored) and used ONLY to
// (1) generate tests/golden/*.ref fixtures (tests/golden/make_golden.py) and
// (2) serve as the `--impl reference` / cpu_baseline "reference" timing arm of bench.py.
// No reference source text is copied here; this file only *calls* the reference API.
//
// usage:  cd <dir with options,top.init,config.init> && sc_ref_driver dump|dump0 [out]   (dump0: pairs with particle 0 only)
//         sc_ref_driver wlorder [out]   Wang-Landau order parameters of the configuration, from-scratch forms (wanglandau.h:125-290, 516-660, mesh.cpp:10-187)
//         sc_ref_driver time <ntargets> <reps> [count_per_moltype ...]  -> "REFJSON {...}" line: oneToAll timing
#include <cstdio>
#include <cstring>
#include <ctime>
#include <chrono>
#include <string>
#include <vector>

#include "mc/inicializer.h"
#include "mc/totalenergycalculator.h"
#include "mc/randomGenerator.h"
// the `wlorder` mode calls the reference's own order-parameter members (WangLandau::zOrder, zOrient, twoPartDist, contParticlesAll,
// boxSize_x/y: private inline members of mc/wanglandau.h): the class is opened for this one header (access only; everything the
// header includes has been included above or has no private part)
#define private public
#include "mc/wanglandau.h"
#undef private
#ifdef REF_FULL      // sc_ref_full: built from a scratch copy whose calculator typedef is TotalEFull (oracle/Makefile); the sweep mode calls the
#include <iomanip>   // reference's own MoveCreator::partDisplace / partRotate on chosen targets (private members: opened for this one header,
#include "mc/wanglandau.h"   // everything it includes has been included above)
#define private public
#include "mc/movecreator.h"
#undef private
#endif

using namespace std;

// globals the reference expects its main.cpp to define (main.cpp:18-30)
MpiCout mcout(0);
Topo topo;
#ifdef RAN2
Ran2 ran2;
#else
#ifdef DSFMT
Dsfmt ran2;
#else
MersenneTwister ran2;
#endif
#endif

static void pv(FILE* f, const Vector& v) { fprintf(f, " %a %a %a", v.x, v.y, v.z); }

static void load(Conf& conf, Sim*& sim, FileNames& files, long big_counts_n, const long* big_counts) {
    sim = new Sim(&conf, &files, 0, 1);
    Inicializer init(sim, &conf, &files);
    init.initTop();
    init.testChains();
    if (big_counts_n > 0) {
        // Systems above the reference's MAXN: top.init lists ONE molecule of each molecule type;
        // replicate each molecule big_counts[k] times here, then fix the group list the same way
        // Inicializer::initGroupLists does (inicializer.cpp:330-372).
        ParticleVector proto = conf.pvec;
        conf.pvec.clear();
        int ntypes = proto.molTypeCount;
        int at = 0;
        for (int t = 0; t < ntypes; t++) {
            int msz = topo.moleculeParam[t].molSize();
            conf.pvec.first[t] = (int)conf.pvec.size();
            for (long c = 0; c < big_counts[t]; c++)
                for (int k = 0; k < msz; k++) conf.pvec.push_back(proto[at + k]);
            at += msz;
        }
        conf.pvec.first[ntypes] = (int)conf.pvec.size();
        conf.pvec.molTypeCount = ntypes;
        conf.pvec.calcChainCount();
    }
    FILE* infile = fopen(files.configurationInFile, "r");
    if (!infile) { fprintf(stderr, "cannot open config.init\n"); exit(1); }
    if (!init.initConfig(&infile, conf.pvec)) exit(1);
    fclose(infile);
    conf.partVecInit();   // Updater::simulate does this before any energy (updater.cpp:45-60)
}

static int conidx(Conf& conf, Particle* p) { return p ? (int)(p - &conf.pvec[0]) : -1; }

static int do_dump(const char* outname, bool only0) {
    FileNames files(0);
    Conf conf;
    Sim* sim = nullptr;
    load(conf, sim, files, 0, nullptr);
    FILE* f = fopen(outname, "w");
    int n = (int)conf.pvec.size();
    fprintf(f, "N %d\n", n);
    fprintf(f, "BOX %a %a %a\n", conf.geo.box.x, conf.geo.box.y, conf.geo.box.z);
    fprintf(f, "CUT %a %a\n", topo.sqmaxcut, topo.maxcut);
    // types in use
    bool used[MAXT] = {false};
    for (int i = 0; i < n; i++) used[conf.pvec[i].type] = true;
    for (int a = 0; a < MAXT; a++) for (int b = 0; b < MAXT; b++) {
        if (!used[a] || !used[b]) continue;
        const Ia_param& p = topo.ia_params[a][b];
        fprintf(f, "IA %d %d %d %d %d", a, b, p.geotype[0], p.geotype[1], (int)p.exclude);
        fprintf(f, " %a %a %a %a %a %a %a %a %a %a %a %a", p.sigma, p.epsilon, p.A, p.B, p.pdis, p.pswitch,
                p.pswitchINV, p.rcut, p.rcutSq, p.rcutwca, p.rcutwcaSq, p.parallel);
        fprintf(f, " %a %a %a %a", p.len[0], p.len[1], p.half_len[0], p.half_len[1]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.pangl[k]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.panglsw[k]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.pcangl[k]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.pcanglsw[k]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.pcoshalfi[k]);
        for (int k = 0; k < 4; k++) fprintf(f, " %a", p.psinhalfi[k]);
        fprintf(f, " %a %a %a %a %a %a %a %a", p.csecpatchrot[0], p.csecpatchrot[1], p.ssecpatchrot[0], p.ssecpatchrot[1],
                p.chiral_cos[0], p.chiral_cos[1], p.chiral_sin[0], p.chiral_sin[1]);
        fprintf(f, "\n");
    }
    for (int m = 0; m < conf.pvec.molTypeCount; m++) {
        MoleculeParams& q = topo.moleculeParam[m];
        fprintf(f, "MOL %d %d %d %a %a %a %a %a %a %a %a %a %a %a %a\n", m, q.molSize(), conf.pvec.first[m],
                q.bond1eq, q.bond1c, q.bond2eq, q.bond2c, q.bonddeq, q.bonddc, q.bondheq, q.bondhc,
                q.angle1eq, q.angle1c, q.angle2eq, q.angle2c);
    }
    for (int i = 0; i < n; i++) {
        Particle& p = conf.pvec[i];
        fprintf(f, "P %d %d %d", i, p.type, p.molType);
        pv(f, p.pos); pv(f, p.dir); pv(f, p.patchdir[0]); pv(f, p.patchdir[1]);
        for (int k = 0; k < 4; k++) pv(f, p.patchsides[k]);
        pv(f, p.chdir[0]); pv(f, p.chdir[1]);
        ConList c = conf.pvec.getConlist(i);
        fprintf(f, " %d %d %d %d %d\n", (int)c.isEmpty, conidx(conf, c.conlist[0]), conidx(conf, c.conlist[1]),
                conidx(conf, c.conlist[2]), conidx(conf, c.conlist[3]));
    }
    PairE pairE(&conf.geo);
    // pairs, both argument orders, non-zero only
    long npair = 0;
    for (int i = 0; i < n; i++) {
        ConList ci = conf.pvec.getConlist(i);
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            if (only0 && i != 0 && j != 0) continue;   // pose grids: pairs with particle 0 only
            double e = pairE(&conf.pvec[i], &conf.pvec[j], &ci);
            if (e != 0.0) { fprintf(f, "E %d %d %a\n", i, j, e); npair++; }
        }
    }
    fprintf(f, "NPAIR %ld\n", npair);
    if (only0) {   // pose grids: overlap flags of particle 0 with every pose, both argument orders
        for (int j = 1; j < n; j++) {
            int a = conf.overlap(&conf.pvec[0], &conf.pvec[j], topo.ia_params);
            int b = conf.overlap(&conf.pvec[j], &conf.pvec[0], topo.ia_params);
            if (a || b) fprintf(f, "OV0 %d %d %d\n", j, a, b);
        }
        fclose(f);
        return 0;
    }
    // oneToAll in TotalEFull order (ascending i != target)
    for (int t = 0; t < n; t++) {
        ConList ct = conf.pvec.getConlist(t);
        double energy = 0.0;
        for (int i = 0; i < n; i++) if (i != t) energy += pairE(&conf.pvec[t], &conf.pvec[i], &ct);
        fprintf(f, "ONE %d %a\n", t, energy);
    }
    // total in TotalEMatrix::allToAll(matrix) order: i from 1, j<i, conlist(i)
    {
        double energy = 0.0;
        for (int i = 1; i < n; i++) {
            ConList ci = conf.pvec.getConlist(i);
            for (int j = 0; j < i; j++) energy += pairE(&conf.pvec[i], &conf.pvec[j], &ci);
        }
        fprintf(f, "TOTAL %a\n", energy);
    }
    // mol2othersTrial-style sums (empty conlist, partners outside the molecule) for chain molecules
    for (int c = 0; c < conf.pvec.getChainCount(); c++) {
        Molecule mol = conf.pvec.getChain(c);
        ConList empty;
        double energy = 0.0;
        for (unsigned j = 0; j < mol.size(); j++) {
            for (int i = 0; i < mol[0]; i++) energy += pairE(&conf.pvec[mol[j]], &conf.pvec[i], &empty);
            for (int i = mol.back() + 1; i < n; i++) energy += pairE(&conf.pvec[mol[j]], &conf.pvec[i], &empty);
        }
        fprintf(f, "MOL2O %d %d %d %a\n", c, mol[0], (int)mol.size(), energy);
    }
    // overlap flags as written (dead code in the reference, Conf.cpp:104-239); flagged pairs only
    long nov = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) {
        if (conf.overlap(&conf.pvec[i], &conf.pvec[j], topo.ia_params)) { fprintf(f, "OV %d %d\n", i, j); nov++; }
    }
    fprintf(f, "NOV %ld\n", nov);
    fclose(f);
    return 0;
}

// [EXTER] wall potential: the parameters the reference derives (topo.exter, structures/topo.cpp:120-130, 151-152) and
// ExternalEnergyCalculator::extere2 of every particle of the configuration in the current directory
static int do_exter(const char* outname) {
    FileNames files(0);
    Conf conf;
    Sim* sim = nullptr;
    load(conf, sim, files, 0, nullptr);
    FILE* f = fopen(outname, "w");
    int n = (int)conf.pvec.size();
    fprintf(f, "N %d\n", n);
    fprintf(f, "BOX %a %a %a\n", conf.geo.box.x, conf.geo.box.y, conf.geo.box.z);
    fprintf(f, "EXTP %d %a %a %a %a\n", (int)topo.exter.exist, topo.exter.thickness, topo.exter.epsilon, topo.exter.attraction, topo.exter.sqmaxcut);
    bool used[MAXT] = {false};
    for (int i = 0; i < n; i++) used[conf.pvec[i].type] = true;
    for (int t = 0; t < MAXT; t++) if (used[t]) {
        const Ia_param& p = topo.exter.interactions[t];
        fprintf(f, "EXTI %d %d %a %a %a %a %a %a %a %a\n", t, p.geotype[0], p.sigma, p.epsilon, p.rcutwca, p.rcut, p.pdis, p.pswitch, p.len[0], p.half_len[0]);
    }
    ExternalEnergyCalculator ex(&conf.geo.box);
    for (int i = 0; i < n; i++) {
        Particle& p = conf.pvec[i];
        fprintf(f, "P %d %d %d", i, p.type, p.molType);
        pv(f, p.pos); pv(f, p.dir); pv(f, p.patchdir[0]); pv(f, p.patchdir[1]);
        for (int k = 0; k < 4; k++) pv(f, p.patchsides[k]);
        pv(f, p.chdir[0]); pv(f, p.chdir[1]);
        fprintf(f, "\n");
        fprintf(f, "EXT %d %a\n", i, ex.extere2(&conf.pvec[i]));
    }
    fclose(f);
    return 0;
}

// Timing arm: the reference's own TotalEFull<PairE> (the compile-time alternative calculator,
// totalenergycalculator.h:525-619; the default TotalEMatrix needs 2 x N^2/2 doubles and cannot hold 65k particles)
// looping over the reference's neighbour lists (conf.neighborList, filled here for the sampled targets with the
// cut-off rule of Updater::genSimplePairList, updater.cpp:484-552, and main.cpp:117-126). Only the energy loop is
// timed -- the reference's O(N^2) list construction is reported separately, not charged. `reps` repeats the sample.
static int do_time(int argc, char** argv) {
    long ntargets = argc > 2 ? atol(argv[2]) : 64;
    int reps = argc > 3 ? atoi(argv[3]) : 1;
    vector<long> counts;
    for (int a = 4; a < argc; a++) counts.push_back(atol(argv[a]));
    FileNames files(0);
    Conf conf;
    Sim* sim = nullptr;
    load(conf, sim, files, (long)counts.size(), counts.data());
    long n = (long)conf.pvec.size();
    for (int i = 0; i < MAXT; i++) for (int j = 0; j < MAXT; j++) {   // main.cpp:117-126
        double m = AVER(sim->stat.trans[i].mx, sim->stat.trans[j].mx);
        m *= (1 + sim->pairlist_update) * 2;
        m += topo.maxcut;
        sim->max_dist_squared[i][j] = m * m;
    }
    long stride = n / ntargets; if (stride < 1) stride = 1;
    vector<int> targets;
    for (long t = 0; t < n && (long)targets.size() < ntargets; t += stride) targets.push_back((int)t);
    conf.neighborList.resize(n);
    conf.pairlist_update = true;
    double sq = topo.sqmaxcut;
    long gated = 0, cand = 0;
    auto l0 = chrono::steady_clock::now();
    for (int t : targets) {
        vector<long> nb;
        ConList ct = conf.pvec.getConlist(t);
        for (long i = 0; i < n; i++) if (i != t) {
            Vector r = conf.geo.image(&conf.pvec[t].pos, &conf.pvec[i].pos);
            double r2 = r.dot(r);
            bool bonded = false;
            for (int q = 0; q < 4; q++) if (ct.conlist[q] == &conf.pvec[i]) bonded = true;
            if (r2 <= sim->max_dist_squared[conf.pvec[t].type][conf.pvec[i].type] || bonded) {
                nb.push_back(i);
                cand++;
                if (!(r2 > sq && ct.isEmpty)) gated++;
            }
        }
        conf.neighborList[t].neighborID = (long*)malloc(sizeof(long) * (nb.size() + 1));
        memcpy(conf.neighborList[t].neighborID, nb.data(), sizeof(long) * nb.size());
        conf.neighborList[t].neighborCount = (long)nb.size();
    }
    auto l1 = chrono::steady_clock::now();
    double list_s = chrono::duration<double>(l1 - l0).count();
    TotalEFull<PairE> calc(sim, &conf);
    calc.pairListUpdate = true;
    auto t0 = chrono::steady_clock::now();
    double acc = 0.0;
    for (int r = 0; r < reps; r++)
        for (int t : targets) acc += calc.oneToAll(t);
    auto t1 = chrono::steady_clock::now();
    double one_s = chrono::duration<double>(t1 - t0).count();
    for (int t : targets) { free(conf.neighborList[t].neighborID); conf.neighborList[t].neighborID = NULL; }
    conf.neighborList.clear();
    printf("REFJSON {\"n\": %ld, \"targets\": %ld, \"reps\": %d, \"one_to_all_s\": %.6f, \"list_candidates\": %ld, "
           "\"gated_pairs\": %ld, \"list_build_s\": %.6f, \"sum\": %.17g}\n",
           n, (long)targets.size(), reps, one_s, cand * reps, gated * reps, list_s, acc);
    return 0;
}

#ifdef REF_FULL
// Measured sequential sweeps of the reference's own move code at sizes its default calculator cannot hold: MoveCreator::partDisplace /
// partRotate (mc/movecreator.cpp:947-1028) over TotalEFull<PairE> (the reference's compile-time alternative calculator, :525-619)
// with the reference's neighbour lists (cut-off rule of Updater::genSimplePairList, updater.cpp:484-552, main.cpp:117-126), on a
// bounded sample: `ntargets` evenly spaced particles get their lists (the O(N) scan per target is reported, not charged -- the
// reference amortises it over pairlist_update sweeps), then `ntrials` trial moves are drawn among them exactly as particleMove()
// does (movecreator.cpp:11-33). One sweep = N such trials (updater.cpp:206).
static int do_sweep(int argc, char** argv) {
    long ntargets = argc > 2 ? atol(argv[2]) : 1024;
    long ntrials = argc > 3 ? atol(argv[3]) : 8192;
    vector<long> counts;
    for (int a = 4; a < argc; a++) counts.push_back(atol(argv[a]));
    FileNames files(0);
    Conf conf;
    Sim* sim = nullptr;
    load(conf, sim, files, (long)counts.size(), counts.data());
    long n = (long)conf.pvec.size();
    for (int i = 0; i < MAXT; i++) for (int j = 0; j < MAXT; j++) {   // main.cpp:117-126
        double m = AVER(sim->stat.trans[i].mx, sim->stat.trans[j].mx);
        m *= (1 + sim->pairlist_update) * 2;
        m += topo.maxcut;
        sim->max_dist_squared[i][j] = m * m;
    }
    long stride = n / ntargets; if (stride < 1) stride = 1;
    vector<int> targets;
    for (long t = 0; t < n && (long)targets.size() < ntargets; t += stride) targets.push_back((int)t);
    conf.neighborList.resize(n);
    conf.pairlist_update = true;
    auto l0 = chrono::steady_clock::now();
    for (int t : targets) {
        vector<long> nb;
        ConList ct = conf.pvec.getConlist(t);
        for (long i = 0; i < n; i++) if (i != t) {
            Vector r = conf.geo.image(&conf.pvec[t].pos, &conf.pvec[i].pos);
            bool bonded = false;
            for (int q = 0; q < 4; q++) if (ct.conlist[q] == &conf.pvec[i]) bonded = true;
            if (r.dot(r) <= sim->max_dist_squared[conf.pvec[t].type][conf.pvec[i].type] || bonded) nb.push_back(i);
        }
        conf.neighborList[t].neighborID = (long*)malloc(sizeof(long) * (nb.size() + 1));
        memcpy(conf.neighborList[t].neighborID, nb.data(), sizeof(long) * nb.size());
        conf.neighborList[t].neighborCount = (long)nb.size();
    }
    double list_s = chrono::duration<double>(chrono::steady_clock::now() - l0).count();
    TotalEnergyCalculator calc(sim, &conf);      // = TotalEFull<PairE> in this build
    calc.pairListUpdate = true;
    MoveCreator move(sim, &conf, &calc);
    auto t0 = chrono::steady_clock::now();
    double edrift = 0.0;
    for (long k = 0; k < ntrials; k++) {
        long target = targets[(size_t)(ran2() * (double)targets.size())];
        if ((ran2() < 0.5) || (topo.ia_params[conf.pvec[target].type][conf.pvec[target].type].geotype[0] >= SP)) edrift += move.partDisplace(target);
        else edrift += move.partRotate(target);
    }
    double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    long acc = 0, rej = 0;
    for (int t = 0; t < MAXT; t++) { acc += sim->stat.trans[t].acc + sim->stat.rot[t].acc; rej += sim->stat.trans[t].rej + sim->stat.rot[t].rej; }
    printf("REFJSON {\"n\": %ld, \"targets\": %ld, \"trials\": %ld, \"trial_s\": %.6f, \"list_build_s\": %.6f, \"accepted\": %ld, \"rejected\": %ld, "
           "\"sweeps_per_s_one_core\": %.6f, \"edrift\": %.17g}\n", n, (long)targets.size(), ntrials, s, list_s, acc, rej, (double)ntrials / s / (double)n, edrift);
    return 0;
}
#endif

// Wang-Landau order parameters of the whole configuration in the current directory, as the reference computes them from scratch
// (WangLandau::init, wanglandau.cpp:56-125, and the runPress forms, wanglandau.h:170-196): Conf::massCenter + zOrder (wlm 1),
// Mesh::meshInit = meshFill + findHoles (wlm 2, mesh.cpp:10-187), zOrient (3), twoPartDist (4), contParticlesAll (7), boxSize_x/y (8, 9).
// wlm 5 / 6 (radiusholeAll, wanglandau.cpp:306-338) are not dumped: a local variable shadows the member radiusholemax, the array is never
// allocated and the first call writes through a null pointer.
static int do_wlorder(const char* outname) {
    FileNames files(0);
    Conf conf;
    Sim* sim = nullptr;
    load(conf, sim, files, 0, nullptr);
    FILE* f = fopen(outname, "w");
    const long n = (long)conf.pvec.size();
    fprintf(f, "N %ld\n", n);
    fprintf(f, "BOX %a %a %a\n", conf.geo.box.x, conf.geo.box.y, conf.geo.box.z);
    conf.massCenter();                                         // Conf.cpp:78-95
    fprintf(f, "SYSCM %a %a %a %a\n", conf.syscm.x, conf.syscm.y, conf.syscm.z, conf.sysvolume);
    bool used[MAXT] = {false};
    for (long i = 0; i < n; i++) used[conf.pvec[i].type] = true;
    for (int t = 0; t < MAXT; t++) if (used[t]) fprintf(f, "VOL %d %a\n", t, topo.ia_params[t][t].volume);
    WangLandau wl(&conf, sim);
    wl.currorder[0] = wl.currorder[1] = 0;
    wl.neworder[0] = wl.neworder[1] = 0;
    wl.mesh.data = NULL; wl.mesh.tmp = NULL;
    const double bins[3][2] = {{-3.0, 0.25}, {0.0, 1.0}, {-1.0, 0.0078125}};      // (minorder, dorder) pairs
    for (int b = 0; b < 3; b++) {
        wl.minorder[0] = bins[b][0]; wl.dorder[0] = bins[b][1];
        fprintf(f, "BIN %a %a\n", wl.minorder[0], wl.dorder[0]);
        fprintf(f, "W1 %ld\n", wl.zOrder(0));
        wl.zOrient(0, 0);
        fprintf(f, "W3 %ld\n", wl.neworder[0]);
        if (n > 1) fprintf(f, "W4 %ld\n", wl.twoPartDist(0));
        wl.boxSize_x(0);
        fprintf(f, "W8 %ld\n", wl.neworder[0]);
        wl.boxSize_y(0);
        fprintf(f, "W9 %ld\n", wl.neworder[0]);
        for (int t = 0; t < MAXT; t++) if (used[t]) {
            wl.wlmtype = t;
            long o = wl.contParticlesAll(0);
            fprintf(f, "W7 %d %ld %ld\n", t, wl.partincontact, o);
        }
    }
    // hole in the xy plane: the mesh of every particle type present, at the reference's own mesh size sigma / 3 (wanglandau.cpp:79)
    // and at finer ones (more and larger holes); order = (long)((maxsize - minorder) / dorder) as wanglandau.cpp:84-88
    for (int t = 0; t < MAXT; t++) if (used[t]) {
        const double sig = topo.ia_params[t][t].sigma;
        const double sizes[4] = {sig / 3.0, sig / 5.0, sig / 8.0, sig / 12.0};
        for (int k = 0; k < 4; k++) {
            if ((int)(conf.geo.box.x / sizes[k]) < 1 || (int)(conf.geo.box.y / sizes[k]) < 1) continue;
            int maxsize = wl.mesh.meshInit(sizes[k], n, t, conf.geo.box, &conf.pvec);
            long occupied = 0;
            for (int i = 0; i < wl.mesh.dim[0] * wl.mesh.dim[1]; i++) if (wl.mesh.data[i] < 0) occupied++;
            // Mesh::data as findHoles left it (occupied < 0, free = hole number): number of holes and an FNV-1a hash of the array
            int nholes = 0;
            unsigned long long h = 0xcbf29ce484222325ull;
            for (int i = 0; i < wl.mesh.dim[0] * wl.mesh.dim[1]; i++) {
                if (wl.mesh.data[i] > nholes) nholes = wl.mesh.data[i];
                h = (h ^ (unsigned long long)(unsigned int)wl.mesh.data[i]) * 0x100000001b3ull;
            }
            fprintf(f, "W2 %d %a %d %d %d %ld %ld %d %llx\n", t, sizes[k], wl.mesh.dim[0], wl.mesh.dim[1], maxsize, occupied,
                    (long)((maxsize - 1.0) / 4.0), nholes, h);
        }
    }
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
#ifdef REF_FULL
    if (argc >= 2 && !strcmp(argv[1], "sweep")) return do_sweep(argc, argv);
#endif
    if (argc >= 2 && !strcmp(argv[1], "dump")) return do_dump(argc > 2 ? argv[2] : "ref_dump.txt", false);
    if (argc >= 2 && !strcmp(argv[1], "dump0")) return do_dump(argc > 2 ? argv[2] : "ref_dump.txt", true);
    if (argc >= 2 && !strcmp(argv[1], "exter")) return do_exter(argc > 2 ? argv[2] : "ref_exter.txt");
    if (argc >= 2 && !strcmp(argv[1], "wlorder")) return do_wlorder(argc > 2 ? argv[2] : "ref_wlorder.txt");
    if (argc >= 2 && !strcmp(argv[1], "time")) return do_time(argc, argv);
    fprintf(stderr, "usage: sc_ref_driver dump|dump0 [out] | time <ntargets> <reps> [count_per_moltype ...]\n");
    return 2;
}
