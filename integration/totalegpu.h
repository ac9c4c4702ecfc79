// TotalEGpu -- the energy calculator a maintainer of robertvacha/SC (scOOP) adds next to TotalEMatrix / TotalEFull to run the
// pair-energy hot path on a B200 through the scgpu C ABI (include/scgpu.h). It is the binding INTEGRATION.md describes, as a file:
//
//   scOOP/mc/totalenergycalculator.h, before the typedef block (:1226-1228):
//       #include "totalegpu.h"
//       typedef TotalEGpu<PairE> TotalEnergyCalculator;
//   link:  -I<repo>/integration -I<repo>/include -L<repo>/sc_b200 -lscgpu
//
// Nothing else of the reference changes: topology / options / configuration parsing, moves, Wang-Landau, output stay as they are.
// oracle/Makefile (target `scgpu_ref`) applies exactly these two lines with sed to a scratch copy of the reference under
// oracle/_ref/ and builds oracle/_ref/SC_scgpu = the reference's own main.cpp on the GPU calculator; tests/test_gpu_dropin.py runs
// the reference's regression inputs (Tests/test_*, Tests/volumeChange/*) through it and compares config.last byte for byte with
// what the unmodified reference wrote.
//
// Protocol (TotalE<>, scOOP/mc/totalenergycalculator.h:135-297): the caller mutates conf->pvec[target] (or conf->geo.box) in place,
// calls the ...Trial() method, then either restores the particle / box itself (reject) or calls update(...) (accept). The device
// keeps the last committed configuration; trial states travel as the 30-double record of the mutated particle(s).
#ifndef TOTALEGPU_H
#define TOTALEGPU_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "scgpu.h"
#include "wl_gpu_hook.h"      // scgpu_dropin_ctx(): lets the optional Wang-Landau hook (Mesh::meshInit on the device) find this calculator's context

template<typename pairEFce>
class TotalEGpu : public TotalE<pairEFce> {
    scgpu_ctx* ctx;
    std::vector<double> st;                 // packed particle records (30 doubles each)
    std::vector<int> type, moltype;
    using TotalE<pairEFce>::conf;

    static void fail(const char* what) { fprintf(stderr, "scgpu %s: %s\n", what, scgpu_last_error()); exit(1); }
    static void pack(const Particle& p, double* o) {     // member order of Particle (structures/particle.h:26-30)
        const Vector* v[10] = {&p.pos, &p.dir, &p.patchdir[0], &p.patchdir[1], &p.patchsides[0], &p.patchsides[1],
                               &p.patchsides[2], &p.patchsides[3], &p.chdir[0], &p.chdir[1]};
        for (int k = 0; k < 10; k++) { o[3 * k] = v[k]->x; o[3 * k + 1] = v[k]->y; o[3 * k + 2] = v[k]->z; }
    }
    // Type-switch moves (MoveCreator::switchTypeMove, mc/movecreator.cpp:233-303) change conf->pvec[t].type in place before oneToAllTrial(t)
    // and, when the move is rejected, change it back WITHOUT telling the calculator. The device holds each particle's type: it follows the host
    // at the trial (syncType) and the particle is remembered, so that whichever entry point is called next settles a rejected switch first.
    int switched;
    void syncType(int t) {
        if (conf->pvec[t].type != type[t]) { type[t] = conf->pvec[t].type; if (scgpu_set_particle_type(ctx, t, type[t])) fail("set_particle_type"); }
    }
    void settle() { if (switched >= 0) { int t = switched; switched = -1; syncType(t); } }
    void pushBox() { settle(); double b[3] = {conf->geo.box.x, conf->geo.box.y, conf->geo.box.z}; if (scgpu_set_box(ctx, b)) fail("set_box"); }
    void pushTopology() {
        int T = 0, M = conf->pvec.molTypeCount;
        for (unsigned i = 0; i < conf->pvec.size(); i++) if (conf->pvec[i].type + 1 > T) T = conf->pvec[i].type + 1;
        for (int m = 0; m < M; m++) {                         // types a switch move may turn a particle into (moleculeparams.h:36-38)
            MoleculeParams& q = topo.moleculeParam[m];
            for (unsigned k = 0; k < q.switchTypes.size(); k++) if (q.switchTypes[k] + 1 > T) T = q.switchTypes[k] + 1;
            for (unsigned k = 0; k < q.particleTypes.size(); k++) if (q.particleTypes[k] + 1 > T) T = q.particleTypes[k] + 1;
        }
        std::vector<scgpu_iaparam> tab((size_t)T * T);
        for (int a = 0; a < T; a++) for (int b = 0; b < T; b++) {
            const Ia_param& q = topo.ia_params[a][b];
            scgpu_iaparam& r = tab[(size_t)a * T + b];
            memset(&r, 0, sizeof r);
            r.geotype[0] = q.geotype[0]; r.geotype[1] = q.geotype[1]; r.exclude = q.exclude;
            r.sigma = q.sigma; r.epsilon = q.epsilon; r.A = q.A; r.B = q.B; r.pdis = q.pdis; r.pswitch = q.pswitch;
            r.pswitchINV = q.pswitchINV; r.rcut = q.rcut; r.rcutSq = q.rcutSq; r.rcutwca = q.rcutwca; r.rcutwcaSq = q.rcutwcaSq;
            r.parallel = q.parallel;
            for (int k = 0; k < 2; k++) {
                r.half_len[k] = q.half_len[k]; r.len[k] = q.len[k]; r.csecpatchrot[k] = q.csecpatchrot[k];
                r.ssecpatchrot[k] = q.ssecpatchrot[k]; r.chiral_cos[k] = q.chiral_cos[k]; r.chiral_sin[k] = q.chiral_sin[k];
            }
            for (int k = 0; k < 4; k++) { r.pcangl[k] = q.pcangl[k]; r.pcanglsw[k] = q.pcanglsw[k]; r.pcoshalfi[k] = q.pcoshalfi[k]; r.psinhalfi[k] = q.psinhalfi[k]; }
        }
        std::vector<scgpu_molparam> mol(M);
        for (int m = 0; m < M; m++) {
            MoleculeParams& q = topo.moleculeParam[m];
            scgpu_molparam& r = mol[m];
            memset(&r, 0, sizeof r);
            r.bond1eq = q.bond1eq; r.bond1c = q.bond1c; r.bond2eq = q.bond2eq; r.bond2c = q.bond2c; r.bonddeq = q.bonddeq; r.bonddc = q.bonddc;
            r.bondheq = q.bondheq; r.bondhc = q.bondhc; r.angle1eq = q.angle1eq; r.angle1c = q.angle1c; r.angle2eq = q.angle2eq; r.angle2c = q.angle2c;
            r.mol_size = q.molSize(); r.first = conf->pvec.first[m];
        }
        if (scgpu_set_topology(ctx, T, tab.data(), topo.sqmaxcut, topo.maxcut, M, mol.data())) fail("set_topology");
        // the [EXTER] wall: the library derives topo.exter.interactions[] from the table and adds extere2 wherever the reference's calculators do
        if (scgpu_set_exter(ctx, topo.exter.exist ? 1 : 0, topo.exter.thickness, topo.exter.epsilon, topo.exter.attraction)) fail("set_exter");
    }
public:
    // the non-virtual helpers of the base class that muVT / cluster / analysis code calls stay visible (and stay host code)
    using TotalE<pairEFce>::mol2others;
    using TotalE<pairEFce>::oneToAll;

    TotalEGpu(Sim* sim, Conf* conf) : TotalE<pairEFce>(sim, conf), ctx(NULL), switched(-1) { if (scgpu_create(&ctx, sim->mpirank)) fail("create"); scgpu_dropin_ctx() = ctx; }   // one replica per GPU
    ~TotalEGpu() { if (scgpu_dropin_ctx() == ctx) scgpu_dropin_ctx() = NULL; scgpu_destroy(ctx); }

    void initEM() override {                                   // replaces allToAll(eMat.energyMatrix)
        size_t n = conf->pvec.size();
        st.resize(n * 30); type.resize(n); moltype.resize(n);
        for (size_t i = 0; i < n; i++) { pack(conf->pvec[i], &st[i * 30]); type[i] = conf->pvec[i].type; moltype[i] = conf->pvec[i].molType; }
        switched = -1;
        pushTopology();
        pushBox();
        if (scgpu_set_particles(ctx, (int)n, st.data(), type.data(), moltype.data())) fail("set_particles");
    }
    void update() override { pushBox(); }                       // accepted volume move
    void update(EMResize) override { initEM(); }                // muVT changed the particle count
    void update(int t) override {
        settle();
        syncType(t);                                            // an accepted type switch keeps the new type
        double* s = &st[(size_t)t * 30]; pack(conf->pvec[t], s); if (scgpu_update_particle(ctx, t, s)) fail("update");
    }
    void update(Molecule m) override { for (unsigned k = 0; k < m.size(); k++) update(m[k]); }

    // Some callers write conf->pvec WITHOUT telling the calculator: the geometric cluster move reflects a whole cluster in place and then
    // asks for allToAll() (movecreator.cpp:164, 172). TotalEFull recomputes from conf->pvec and is right; the default TotalEMatrix sums its
    // (now stale) matrix. allToAll() here re-reads the host configuration and sends what changed -- the TotalEFull semantics.
    void resync() {
        size_t n = conf->pvec.size();
        if (n * 30 != st.size()) { initEM(); return; }
        double s[30];
        for (size_t i = 0; i < n; i++) {
            pack(conf->pvec[i], s);
            syncType((int)i);
            if (memcmp(s, &st[i * 30], sizeof s)) { memcpy(&st[i * 30], s, sizeof s); if (scgpu_update_particle(ctx, (int)i, s)) fail("update"); }
        }
    }
    double allToAll() override { double e; resync(); pushBox(); if (scgpu_all_to_all(ctx, &e, NULL)) fail("all_to_all"); return e; }
    double allToAllTrial() override { double e; pushBox(); if (scgpu_all_to_all(ctx, &e, NULL)) fail("all_to_all"); return e; }      // the caller has already changed conf->geo.box
    // NB the reference restores conf->geo.box behind the calculator's back on a rejected volume move: the box is re-sent on every call
    double oneToAll(int t) override { double e; pushBox(); if (scgpu_one_to_all(ctx, t, NULL, &e, NULL)) fail("one_to_all"); return e; }
    double oneToAllTrial(int t) override {                      // the caller has already mutated conf->pvec[t]
        double s[30], e;
        pack(conf->pvec[t], s);
        pushBox();
        if (conf->pvec[t].type != type[t]) { syncType(t); switched = t; }      // switchTypeMove: the trial carries a new type
        if (scgpu_one_to_all(ctx, t, s, &e, NULL)) fail("one_to_all");
        return e;
    }
    double mol2others(Molecule& m) override { double e; pushBox(); if (scgpu_mol_to_others(ctx, m[0], (int)m.size(), NULL, &e)) fail("mol2others"); return e; }
    double mol2othersTrial(Molecule& m) override {
        std::vector<double> s(m.size() * 30);
        for (size_t k = 0; k < m.size(); k++) pack(conf->pvec[m[k]], &s[k * 30]);
        double e;
        pushBox();
        if (scgpu_mol_to_others(ctx, m[0], (int)m.size(), s.data(), &e)) fail("mol2others");
        return e;
    }
};

#endif
