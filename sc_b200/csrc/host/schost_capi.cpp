// C entry points over the host mirror, for the Python tests / bench (ctypes). Pointers and sizes only.
#include <cstring>
#include <string>

#include "calculator.hpp"
#include "topology.hpp"

using namespace schost;

static thread_local std::string g_err;

extern "C" {

const char* schost_last_error(void) { return g_err.c_str(); }

// parse top.init + config.init text with the reference's semantics; counts (optional) override the [System] counts
int schost_load_text(const char* top_text, const char* config_text, const long* counts, int ncounts, void** out) {
    try {
        std::vector<long> c;
        if (counts && ncounts > 0) c.assign(counts, counts + ncounts);
        System* s = new System(System::fromText(top_text, config_text, c));
        *out = s;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

void schost_free(void* sys) { delete (System*)sys; }

int schost_dims(void* sys, int* n, int* ntypes, int* nmol) {
    System* s = (System*)sys;
    *n = s->n; *ntypes = s->ntypes; *nmol = (int)s->molTable.size();
    return 0;
}

int schost_export(void* sys, double* state30, int* type, int* moltype, double* ia, double* mol, double* box3, double* cut2) {
    System* s = (System*)sys;
    memcpy(state30, s->state.data(), s->state.size() * sizeof(double));
    memcpy(type, s->type.data(), s->type.size() * sizeof(int));
    memcpy(moltype, s->moltype.data(), s->moltype.size() * sizeof(int));
    memcpy(ia, s->iaTable.data(), s->iaTable.size() * sizeof(scgpu_iaparam));
    memcpy(mol, s->molTable.data(), s->molTable.size() * sizeof(scgpu_molparam));
    for (int d = 0; d < 3; d++) box3[d] = s->box[d];
    cut2[0] = s->topo.sqmaxcut; cut2[1] = s->topo.maxcut;
    return 0;
}

// [EXTER] section: {exists, thickness, epsilon, attraction switch}
int schost_exter(void* sys, double* out4) {
    System* s = (System*)sys;
    out4[0] = s->topo.exterExist ? 1.0 : 0.0; out4[1] = s->topo.exter[0]; out4[2] = s->topo.exter[1]; out4[3] = s->topo.exter[2];
    return 0;
}
int schost_set_state(void* sys, int idx, const double* state30) {
    System* s = (System*)sys;
    if (idx < 0 || idx >= s->n) { g_err = "schost_set_state: index out of range"; return -1; }
    memcpy(&s->state[(size_t)idx * 30], state30, 30 * sizeof(double));
    return 0;
}
int schost_set_box(void* sys, const double* box3) {
    System* s = (System*)sys;
    for (int d = 0; d < 3; d++) s->box[d] = box3[d];
    return 0;
}

// ---- calculator (TotalEGpu) ----
int schost_calc_create(void* sys, int device, void** out) {
    try { *out = new TotalEGpu((System*)sys, device); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
void schost_calc_free(void* calc) { delete (TotalEGpu*)calc; }

#define CALC_D(name, expr)                                                            \
    int name {                                                                        \
        try { TotalEGpu* c = (TotalEGpu*)calc; *out = (expr); return 0; }             \
        catch (const std::exception& e) { g_err = e.what(); return -1; }              \
    }
CALC_D(schost_calc_all_to_all(void* calc, double* out), c->allToAll())
CALC_D(schost_calc_all_to_all_trial(void* calc, double* out), c->allToAllTrial())
CALC_D(schost_calc_one_to_all(void* calc, int target, double* out), c->oneToAll(target))
CALC_D(schost_calc_one_to_all_trial(void* calc, int target, double* out), c->oneToAllTrial(target))
CALC_D(schost_calc_p2p(void* calc, int a, int b, double* out), c->p2p(a, b))
int schost_calc_mol2others(void* calc, int first, int m, int trial, double* out) {
    try {
        TotalEGpu* c = (TotalEGpu*)calc;
        Molecule mol;
        for (int i = 0; i < m; i++) mol.push_back(first + i);
        *out = trial ? c->mol2othersTrial(mol) : c->mol2others(mol);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int schost_calc_update_particle(void* calc, int target) {
    try { ((TotalEGpu*)calc)->update(target); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int schost_calc_update_box(void* calc) {
    try { ((TotalEGpu*)calc)->update(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}
void* schost_calc_ctx(void* calc) { return ((TotalEGpu*)calc)->ctx; }

}  // extern "C"
