#include "calculator.hpp"

namespace schost {

void TotalEGpu::check(int rc, const char* what) const {
    if (rc != SCGPU_OK) throw Error(std::string(what) + ": " + scgpu_last_error());
}

TotalEGpu::TotalEGpu(System* c, int device) : conf(c) {
    check(scgpu_create(&ctx, device), "scgpu_create");
    initEM();
}

TotalEGpu::~TotalEGpu() { scgpu_destroy(ctx); }

void TotalEGpu::pushBox() { check(scgpu_set_box(ctx, conf->box.data()), "scgpu_set_box"); }

// The reference fills its N x N energy matrix here; ours uploads the configuration and sorts it into cells.
void TotalEGpu::initEM() {
    check(scgpu_set_topology(ctx, conf->ntypes, conf->iaTable.data(), conf->topo.sqmaxcut, conf->topo.maxcut,
                             (int)conf->molTable.size(), conf->molTable.data()), "scgpu_set_topology");
    check(scgpu_set_exter(ctx, conf->topo.exterExist ? 1 : 0, conf->topo.exter[0], conf->topo.exter[1], conf->topo.exter[2]), "scgpu_set_exter");
    pushBox();
    check(scgpu_set_particles(ctx, conf->n, conf->state.data(), conf->type.data(), conf->moltype.data()), "scgpu_set_particles");
    check(scgpu_build_cells(ctx), "scgpu_build_cells");
}

void TotalEGpu::update() { pushBox(); }
void TotalEGpu::update(int target) { check(scgpu_update_particle(ctx, target, &conf->state[(size_t)target * 30]), "scgpu_update_particle"); }
void TotalEGpu::update(const Molecule& mol) { for (int i : mol) update(i); }
void TotalEGpu::update(EMResize) { initEM(); }

double TotalEGpu::allToAll() {
    double e = 0.0;
    pushBox();
    check(scgpu_all_to_all(ctx, &e, nullptr), "scgpu_all_to_all");
    return e;
}
double TotalEGpu::allToAllTrial() { return allToAll(); }   // box already mutated by the caller; pushBox() picks it up

// The reference reads the box through a pointer (PairE::pbc), so a REJECTED volume move restores it behind the calculator's
// back: every entry point therefore re-sends the current host box first (three doubles; re-gridding only if ncell changes).
double TotalEGpu::oneToAll(int target) {
    double e = 0.0;
    pushBox();
    check(scgpu_one_to_all(ctx, target, nullptr, &e, nullptr), "scgpu_one_to_all");
    return e;
}
double TotalEGpu::oneToAllTrial(int target) {
    double e = 0.0;
    pushBox();
    check(scgpu_one_to_all(ctx, target, &conf->state[(size_t)target * 30], &e, nullptr), "scgpu_one_to_all");
    return e;
}
double TotalEGpu::mol2others(const Molecule& mol) {
    double e = 0.0;
    pushBox();
    check(scgpu_mol_to_others(ctx, mol.front(), (int)mol.size(), nullptr, &e), "scgpu_mol_to_others");
    return e;
}
double TotalEGpu::mol2othersTrial(const Molecule& mol) {
    double e = 0.0;
    pushBox();
    check(scgpu_mol_to_others(ctx, mol.front(), (int)mol.size(), &conf->state[(size_t)mol.front() * 30], &e), "scgpu_mol_to_others");
    return e;
}
double TotalEGpu::p2p(int part1, int part2) {
    scratch.assign(conf->n, 0.0);
    double e = 0.0;
    pushBox();
    check(scgpu_one_to_all(ctx, part1, nullptr, &e, scratch.data()), "scgpu_one_to_all");
    return scratch[part2];
}
int TotalEGpu::overlapAll(int target, int variant) {
    int f = 0;
    pushBox();
    check(scgpu_overlap_one(ctx, target, &conf->state[(size_t)target * 30], variant, &f), "scgpu_overlap_one");
    return f;
}
int TotalEGpu::checkall(int variant) {
    int f = 0;
    pushBox();
    check(scgpu_overlap_all(ctx, variant, &f), "scgpu_overlap_all");
    return f;
}

}  // namespace schost
