// Multi-GPU layer of the C ABI: parallel-tempering replica exchange and multiple-walker Wang-Landau merges over NCCL.
// (textually included by scgpu.cu after scgpu_ctx, launch_all_to_all, overflow_then_grow and philox4x32 are defined)
//
// Reference: MoveCreator::replicaExchangeMove (scOOP/mc/movecreator.cpp:552-795), the ladder of Sim::readOptions
// (scOOP/structures/sim.h:384-403), MpiExchangeData (scOOP/structures/structures.h:280-357), WangLandau::initCalc shared windows
// (scOOP/mc/wanglandau.cpp:190-212), WangLandau::accept / update (scOOP/mc/wanglandau.h:66-123, 282-289).
// The reference needs an MPI_Alltoall so that every rank can FIND its partner, then four point-to-point messages per pair, and takes
// the decision on one side. Here the packed records of all replicas are all-gathered once and every rank evaluates all pairs
// with the same counter-based random numbers: no second message, and the full energy goes from the reduction kernel into the
// collective without visiting the host.
#pragma once
#include <dlfcn.h>

// ---- NCCL, bound at run time (libnccl.so.2: inside a torch process this is the copy torch already loaded) ----------------------
struct NcclId { char internal[SCGPU_UNIQUE_ID_BYTES]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId /* ncclUniqueId travels by value: 128 bytes */, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
enum { NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };     // ncclDataType_t / ncclRedOp_t values (nccl.h, stable since 2.0)

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) return nullptr;
#define NCCL_SYM(field, name) *(void**)(&api.field) = dlsym(api.handle, name); if (!api.field) { api.handle = nullptr; return nullptr; }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(Broadcast, "ncclBroadcast")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return &api;
}

#define NCK(call)                                                                                          \
    do {                                                                                                   \
        int r_ = (call);                                                                                   \
        if (r_ != 0) {                                                                                     \
            char b_[512];                                                                                  \
            snprintf(b_, sizeof b_, "%s:%d %s: NCCL error %d (%s)", __FILE__, __LINE__, #call, r_, nccl_api()->GetErrorString(r_)); \
            g_err = b_;                                                                                    \
            return SCGPU_ERR_CUDA;                                                                         \
        }                                                                                                  \
    } while (0)

// ---- the packed per-replica record that travels through the collective (doubles) ---------------------------------------------
constexpr int RX = 64;
enum { RX_E = 0, RX_V, RX_N, RX_T, RX_P, RX_PSEUDO, RX_REPLICA, RX_WL0, RX_WL1, RX_ATTEMPTED, RX_ACCEPTED, RX_PARTNER, RX_CHANGE,
       RX_PWL0, RX_PWL1, RX_EDRIFT, RX_PARTNUM = 16, RX_PAYLOAD = 24 };
static_assert(RX_PARTNUM + SCGPU_REPLICA_MOLTYPES == RX_PAYLOAD && RX_PAYLOAD + SCGPU_REPLICA_PAYLOAD == RX, "record layout");

struct PackArgs {
    const double* energy[SCGPU_MAX_LOCAL_REPLICAS];      // device: the total written by k_reduce_fixed of each local replica
    const int* overflow[SCGPU_MAX_LOCAL_REPLICAS];       // device: that context's sticky work-list flags
    double volume[SCGPU_MAX_LOCAL_REPLICAS];
    double npart[SCGPU_MAX_LOCAL_REPLICAS];
};

// one block of RX threads per local replica: thermodynamic state (uploaded) + energy (device) -> record. A launch whose energy
// kernels overflowed a work list marks the record: nobody swaps and every rank repeats the call.
__global__ void k_replica_pack(int base_replica, const double* __restrict__ d_in, PackArgs a, double* __restrict__ d_send) {
    const int r = blockIdx.x, f = threadIdx.x;
    double v = d_in[r * RX + f];
    if (f == RX_E) v = *a.energy[r];
    else if (f == RX_V) v = a.volume[r];
    else if (f == RX_N) v = a.npart[r];
    else if (f == RX_REPLICA) v = (double)(base_replica + r);
    else if (f == RX_ATTEMPTED) v = (*a.overflow[r] != 0) ? -1.0 : 0.0;      // -1: energy invalid
    else if (f == RX_ACCEPTED || f == RX_CHANGE || f == RX_EDRIFT) v = 0.0;
    else if (f == RX_PARTNER) v = -1.0;
    d_send[r * RX + f] = v;
}

struct DecideArgs {
    int R, nrepchange, wl_len;
    long long wl_len0;
    double dtemp, dpress;
    double chempot[SCGPU_REPLICA_MOLTYPES];
    unsigned long long seed, sweep;
};

// replicaExchangeMove's decision for EVERY pair, one thread per pair; every rank runs this on the same gathered records and
// gets the same bits. Pairing (movecreator.cpp:616-623, 652-654, 702-703): the replica whose pseudo-rank `hi` has
// hi % 2 == oddoreven and hi > 0 asks for the temperature of pseudo-rank hi - 1; the decision is evaluated with the lower
// replica's temperature and pressure ("here" = lo, "received" = hi, :722-745).
__global__ void k_replica_decide(DecideArgs a, const double* __restrict__ d_all, double* __restrict__ d_out, const double* __restrict__ wl_all) {
    __shared__ int map[256];
    __shared__ int bad;
    const int R = a.R;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < R * RX; k += blockDim.x) d_out[k] = d_all[k];
    for (int g = threadIdx.x; g < R; g += blockDim.x) {
        const int ps = (int)d_all[g * RX + RX_PSEUDO];
        if (ps >= 0 && ps < R) map[ps] = g;
        if (d_all[g * RX + RX_ATTEMPTED] < 0.0) bad = 1;
    }
    __syncthreads();
    if (bad) {       // some replica's energy is invalid: no decisions, everybody is told
        for (int g = threadIdx.x; g < R; g += blockDim.x) d_out[g * RX + RX_ATTEMPTED] = -1.0;
        return;
    }
    int oddoreven = (a.sweep % (2ull * (unsigned long long)a.nrepchange)) == 0 ? 1 : 0;
    if (R == 2) oddoreven = 1;
    for (int hi = 1 + (int)threadIdx.x; hi < R; hi += blockDim.x) {
        if (hi % 2 != oddoreven) continue;
        const int lo = hi - 1;
        const int gl = map[lo], gh = map[hi];
        const double* L = d_all + gl * RX;
        const double* H = d_all + gh * RX;
        const double T = L[RX_T], P = L[RX_P];
        const double temp = (1 / T - 1 / (T + a.dtemp));
        double change = temp * (L[RX_E] - H[RX_E]);                                              // canonical
        change += (P / T - (P + a.dpress) / (T + a.dtemp)) * (L[RX_V] - H[RX_V]);                // isobaric-isothermal
        for (int i = 0; i < SCGPU_REPLICA_MOLTYPES; i++)                                          // grand canonical, chempot stored as mu/kT
            if (a.chempot[i] != 0.0) change += temp * a.chempot[i] * T * (L[RX_PARTNUM + i] - H[RX_PARTNUM + i]);
        if (wl_all && a.wl_len > 0) {
            const long long localwl = (long long)L[RX_WL0] + (long long)L[RX_WL1] * a.wl_len0;
            const long long receivedwl = (long long)H[RX_WL0] + (long long)H[RX_WL1] * a.wl_len0;
            if (localwl >= 0 && localwl < a.wl_len && receivedwl >= 0 && receivedwl < a.wl_len) {
                const double* wl_l = wl_all + (size_t)gl * a.wl_len;
                const double* wl_h = wl_all + (size_t)gh * a.wl_len;
                change += (-wl_l[localwl] + wl_l[receivedwl]) / T + (-wl_h[receivedwl] + wl_h[localwl]) / (T + a.dtemp);
            }
        }
        const uint4 rnd = philox4x32((uint32_t)a.sweep, (uint32_t)(a.sweep >> 32), (uint32_t)lo, 0x52455058u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const bool accept = (change > 0) || (u01(rnd.x, rnd.y) < exp(change));
        double* OL = d_out + gl * RX;
        double* OH = d_out + gh * RX;
        OL[RX_ATTEMPTED] = 1.0; OH[RX_ATTEMPTED] = 1.0;
        OL[RX_ACCEPTED] = accept ? 1.0 : 0.0; OH[RX_ACCEPTED] = accept ? 1.0 : 0.0;
        OL[RX_PARTNER] = (double)gh; OH[RX_PARTNER] = (double)gl;
        OL[RX_CHANGE] = change; OH[RX_CHANGE] = change;
        OL[RX_PWL0] = H[RX_WL0]; OL[RX_PWL1] = H[RX_WL1]; OH[RX_PWL0] = L[RX_WL0]; OH[RX_PWL1] = L[RX_WL1];
        if (accept) {
            // temperature, pressure, pseudo-rank and the payload change hands; configurations stay where they are
            OL[RX_T] = H[RX_T]; OL[RX_P] = H[RX_P]; OL[RX_PSEUDO] = H[RX_PSEUDO];
            OH[RX_T] = L[RX_T]; OH[RX_P] = L[RX_P]; OH[RX_PSEUDO] = L[RX_PSEUDO];
            for (int k = 0; k < SCGPU_REPLICA_PAYLOAD; k++) { OL[RX_PAYLOAD + k] = H[RX_PAYLOAD + k]; OH[RX_PAYLOAD + k] = L[RX_PAYLOAD + k]; }
            // drift bookkeeping, the same two lines on both sides (:676-684, 751-755)
            for (int side = 0; side < 2; side++) {
                const double* me = side ? H : L;
                const double* ot = side ? L : H;
                const double entrophy = me[RX_P] * me[RX_V] - me[RX_N] * log(me[RX_V]) * me[RX_T];
                double ed = me[RX_P] * (ot[RX_V] - me[RX_V]) - me[RX_N] * log(ot[RX_V] / me[RX_V]) * me[RX_T];
                ed += (ot[RX_P] * me[RX_V] - me[RX_N] * log(me[RX_V]) * ot[RX_T]) - entrophy;
                (side ? OH : OL)[RX_EDRIFT] = ed;
            }
        }
    }
}

struct scgpu_comm {
    int device = 0, nranks = 1, rank = 0;
    void* nccl = nullptr;
    bool owns = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t join[SCGPU_MAX_LOCAL_REPLICAS] = {};
    double *d_in = nullptr, *d_send = nullptr, *d_all = nullptr, *d_out = nullptr;
    double* h_pin = nullptr;           // [2][SCGPU_MAX_LOCAL_REPLICAS][RX]: upload, download
    int rec_cap = 0;                   // records d_all / d_out hold
    double *d_wl = nullptr, *d_wl_all = nullptr;
    size_t wl_cap = 0, wl_all_cap = 0;
    // Wang-Landau merge scratch
    double *d_w = nullptr, *d_wb = nullptr, *d_dw = nullptr;
    long long *d_h = nullptr, *d_hb = nullptr, *d_dh = nullptr;
    double* d_wlstate = nullptr;       // 8 doubles
    int wlm_cap = 0;
    float last_us = 0.f;
};

extern "C" int scgpu_comm_unique_id(char id[SCGPU_UNIQUE_ID_BYTES]) {
    ARG(id != nullptr, "scgpu_comm_unique_id: NULL argument");
    NcclApi* n = nccl_api();
    if (!n) { g_err = "scgpu_comm_unique_id: libnccl.so.2 could not be loaded"; return SCGPU_ERR_CUDA; }
    NcclId u;
    NCK(n->GetUniqueId(&u));
    memcpy(id, u.internal, SCGPU_UNIQUE_ID_BYTES);
    return SCGPU_OK;
}

static int comm_alloc(scgpu_comm* c) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    for (int k = 0; k < SCGPU_MAX_LOCAL_REPLICAS; k++) CK(cudaEventCreateWithFlags(&c->join[k], cudaEventDisableTiming));
    CK(cudaMalloc(&c->d_in, SCGPU_MAX_LOCAL_REPLICAS * RX * sizeof(double)));
    CK(cudaMalloc(&c->d_send, SCGPU_MAX_LOCAL_REPLICAS * RX * sizeof(double)));
    CK(cudaMallocHost((void**)&c->h_pin, 2 * SCGPU_MAX_LOCAL_REPLICAS * RX * sizeof(double)));
    CK(cudaMalloc(&c->d_wlstate, 8 * sizeof(double)));
    return SCGPU_OK;
}

extern "C" int scgpu_comm_create(scgpu_comm** out, int device, int nranks, int rank, const char id[SCGPU_UNIQUE_ID_BYTES]) {
    ARG(out != nullptr, "scgpu_comm_create: out is NULL");
    ARG(nranks >= 1 && rank >= 0 && rank < nranks, "scgpu_comm_create: need 0 <= rank < nranks");
    ARG(nranks == 1 || id != nullptr, "scgpu_comm_create: more than one rank needs the unique id of rank 0");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    ARG(device >= 0 && device < ndev, "scgpu_comm_create: bad device ordinal");
    scgpu_comm* c = new scgpu_comm();
    c->device = device; c->nranks = nranks; c->rank = rank;
    if (int r = comm_alloc(c)) { delete c; return r; }
    if (nranks > 1) {
        NcclApi* n = nccl_api();
        if (!n) { g_err = "scgpu_comm_create: libnccl.so.2 could not be loaded (needed for more than one rank)"; delete c; return SCGPU_ERR_CUDA; }
        NcclId u;
        memcpy(u.internal, id, SCGPU_UNIQUE_ID_BYTES);
        NCK(n->CommInitRank(&c->nccl, nranks, u, rank));
        c->owns = true;
    }
    *out = c;
    return SCGPU_OK;
}

extern "C" int scgpu_comm_attach(scgpu_comm** out, int device, void* nccl_comm, int nranks, int rank) {
    ARG(out != nullptr && nccl_comm != nullptr, "scgpu_comm_attach: NULL argument");
    ARG(nranks >= 1 && rank >= 0 && rank < nranks, "scgpu_comm_attach: need 0 <= rank < nranks");
    if (!nccl_api()) { g_err = "scgpu_comm_attach: libnccl.so.2 could not be loaded"; return SCGPU_ERR_CUDA; }
    scgpu_comm* c = new scgpu_comm();
    c->device = device; c->nranks = nranks; c->rank = rank; c->nccl = nccl_comm; c->owns = false;
    if (int r = comm_alloc(c)) { delete c; return r; }
    *out = c;
    return SCGPU_OK;
}

extern "C" int scgpu_comm_destroy(scgpu_comm* c) {
    if (!c) return SCGPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->owns && c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
    cudaFree(c->d_in); cudaFree(c->d_send); cudaFree(c->d_all); cudaFree(c->d_out); cudaFree(c->d_wl); cudaFree(c->d_wl_all);
    cudaFree(c->d_w); cudaFree(c->d_wb); cudaFree(c->d_dw); cudaFree(c->d_h); cudaFree(c->d_hb); cudaFree(c->d_dh); cudaFree(c->d_wlstate);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    for (int k = 0; k < SCGPU_MAX_LOCAL_REPLICAS; k++) cudaEventDestroy(c->join[k]);
    cudaStreamDestroy(c->stream);
    delete c;
    return SCGPU_OK;
}

extern "C" int scgpu_comm_last_exchange_us(scgpu_comm* c, float* us) {
    ARG(c && us, "scgpu_comm_last_exchange_us: NULL argument");
    *us = c->last_us;
    return SCGPU_OK;
}

static void state_to_record(const scgpu_replica_state& s, double* r) {
    for (int k = 0; k < RX; k++) r[k] = 0.0;
    r[RX_T] = s.temper; r[RX_P] = s.press; r[RX_PSEUDO] = (double)s.pseudo_rank;
    r[RX_WL0] = (double)s.wl_order[0]; r[RX_WL1] = (double)s.wl_order[1];
    for (int k = 0; k < SCGPU_REPLICA_MOLTYPES; k++) r[RX_PARTNUM + k] = s.part_num[k];
    for (int k = 0; k < SCGPU_REPLICA_PAYLOAD; k++) r[RX_PAYLOAD + k] = s.payload[k];
}
static void record_to_state(const double* r, scgpu_replica_state& s) {
    s.temper = r[RX_T]; s.press = r[RX_P]; s.pseudo_rank = (int)r[RX_PSEUDO]; s.replica = (int)r[RX_REPLICA];
    for (int k = 0; k < SCGPU_REPLICA_PAYLOAD; k++) s.payload[k] = r[RX_PAYLOAD + k];
    s.attempted = r[RX_ATTEMPTED] > 0.0 ? 1 : 0; s.accepted = (int)r[RX_ACCEPTED]; s.partner = (int)r[RX_PARTNER]; s.reserved = 0;
    s.partner_wl_order[0] = (int64_t)r[RX_PWL0]; s.partner_wl_order[1] = (int64_t)r[RX_PWL1];
    s.change = r[RX_CHANGE]; s.energy = r[RX_E]; s.volume = r[RX_V]; s.edrift = r[RX_EDRIFT];
}

extern "C" int scgpu_replica_exchange(scgpu_comm* c, int nlocal, scgpu_ctx* const* ctxs, scgpu_replica_state* states,
                                      const scgpu_exchangeparams* p, uint64_t sweep, const double* const* wl_weights) {
    ARG(c && ctxs && states && p, "scgpu_replica_exchange: NULL argument");
    ARG(nlocal >= 1 && nlocal <= SCGPU_MAX_LOCAL_REPLICAS, "scgpu_replica_exchange: nlocal must be in 1 .. 16");
    ARG(p->nrepchange >= 1, "scgpu_replica_exchange: nrepchange must be positive");
    ARG(wl_weights == nullptr || p->wl_len > 0, "scgpu_replica_exchange: Wang-Landau weights need wl_len");
    const int R = c->nranks * nlocal;
    ARG(R >= 2 && R <= 256, "scgpu_replica_exchange: 2 .. 256 replicas in total");
    for (int k = 0; k < nlocal; k++) ARG(ctxs[k] && ctxs[k]->device == c->device && ctxs[k]->n > 0, "scgpu_replica_exchange: every context must live on the communicator's device and hold particles");
    CK(cudaSetDevice(c->device));
    if (R > c->rec_cap) {
        cudaFree(c->d_all); cudaFree(c->d_out);
        c->d_all = c->d_out = nullptr; c->rec_cap = 0;
        CK(cudaMalloc(&c->d_all, (size_t)R * RX * sizeof(double)));
        CK(cudaMalloc(&c->d_out, (size_t)R * RX * sizeof(double)));
        c->rec_cap = R;
    }
    const bool wl = wl_weights != nullptr;
    if (wl) {
        const size_t need = (size_t)nlocal * p->wl_len, need_all = (size_t)R * p->wl_len;
        if (need > c->wl_cap) { cudaFree(c->d_wl); c->d_wl = nullptr; CK(cudaMalloc(&c->d_wl, need * sizeof(double))); c->wl_cap = need; }
        if (need_all > c->wl_all_cap) { cudaFree(c->d_wl_all); c->d_wl_all = nullptr; CK(cudaMalloc(&c->d_wl_all, need_all * sizeof(double))); c->wl_all_cap = need_all; }
    }
    NcclApi* n = c->nranks > 1 ? nccl_api() : nullptr;
    double* h_up = c->h_pin;
    double* h_down = c->h_pin + SCGPU_MAX_LOCAL_REPLICAS * RX;
    for (int attempt = 0; attempt < 8; attempt++) {
        // ---- full energy of every local replica, each on its own stream (they overlap on the GPU)
        PackArgs pa;
        memset(&pa, 0, sizeof pa);
        for (int k = 0; k < nlocal; k++) {
            scgpu_ctx* x = ctxs[k];
            if (int r = ensure_cells(x)) return r;
            if (launch_all_to_all(x)) return SCGPU_ERR_CUDA;
            CK(cudaEventRecord(c->join[k], x->stream));
            CK(cudaStreamWaitEvent(c->stream, c->join[k], 0));
            pa.energy[k] = x->d_scalar; pa.overflow[k] = x->d_pl_overflow;
            pa.volume[k] = x->box[0] * x->box[1] * x->box[2]; pa.npart[k] = (double)x->n;
            state_to_record(states[k], h_up + k * RX);
        }
        CK(cudaEventRecord(c->ev0, c->stream));
        CK(cudaMemcpyAsync(c->d_in, h_up, (size_t)nlocal * RX * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_replica_pack<<<nlocal, RX, 0, c->stream>>>(c->rank * nlocal, c->d_in, pa, c->d_send);
        if (wl) for (int k = 0; k < nlocal; k++)
            CK(cudaMemcpyAsync(c->d_wl + (size_t)k * p->wl_len, wl_weights[k], (size_t)p->wl_len * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (n) {
            NCK(n->AllGather(c->d_send, c->d_all, (size_t)nlocal * RX, NCCL_FLOAT64, c->nccl, c->stream));
            if (wl) NCK(n->AllGather(c->d_wl, c->d_wl_all, (size_t)nlocal * p->wl_len, NCCL_FLOAT64, c->nccl, c->stream));
        } else {
            CK(cudaMemcpyAsync(c->d_all, c->d_send, (size_t)nlocal * RX * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if (wl) CK(cudaMemcpyAsync(c->d_wl_all, c->d_wl, (size_t)nlocal * p->wl_len * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
        DecideArgs da;
        da.R = R; da.nrepchange = p->nrepchange; da.wl_len = wl ? p->wl_len : 0; da.wl_len0 = p->wl_len0; da.dtemp = p->dtemp; da.dpress = p->dpress;
        for (int k = 0; k < SCGPU_REPLICA_MOLTYPES; k++) da.chempot[k] = p->chempot[k];
        da.seed = p->seed; da.sweep = sweep;
        k_replica_decide<<<1, 128, 0, c->stream>>>(da, c->d_all, c->d_out, wl ? c->d_wl_all : nullptr);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_down, c->d_out + (size_t)c->rank * nlocal * RX, (size_t)nlocal * RX * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaEventSynchronize(c->ev1));
        CK(cudaEventElapsedTime(&c->last_us, c->ev0, c->ev1));
        c->last_us *= 1000.f;
        ctxs[0]->launches += 2;
        if (h_down[RX_ATTEMPTED] >= 0.0) {
            for (int k = 0; k < nlocal; k++) record_to_state(h_down + k * RX, states[k]);
            return SCGPU_OK;
        }
        // a work list of some replica (here or on another rank) overflowed: grow what is ours, then everybody repeats
        for (int k = 0; k < nlocal; k++) { bool rep = false; if (int r = overflow_then_grow(ctxs[k], &rep)) return r; }
    }
    g_err = "scgpu_replica_exchange: the energy work lists kept overflowing";
    return SCGPU_ERR_STATE;
}

// ---- multiple-walker Wang-Landau merge -------------------------------------------------------------------------------------------
__global__ void k_wl_delta(int len, const double* __restrict__ w, const double* __restrict__ wb, const long long* __restrict__ h,
                           const long long* __restrict__ hb, double* __restrict__ dw, long long* __restrict__ dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) { dw[i] = w[i] - wb[i]; dh[i] = h[i] - hb[i]; }
}
__global__ void k_wl_apply(int len, double* __restrict__ wb, long long* __restrict__ hb, const double* __restrict__ dw, const long long* __restrict__ dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) { wb[i] += dw[i]; hb[i] += dh[i]; }
}
// WangLandau::update (wanglandau.h:66-123) on the merged arrays; one block. st: {alpha, min, wmin, max, halved, converged}
__global__ void k_wl_update(int len, double* __restrict__ w, long long* __restrict__ h, double temper, double* __restrict__ st) {
    __shared__ long long smin[256], smax[256];
    long long mn = h[0], mx = h[0];
    for (int i = threadIdx.x; i < len; i += blockDim.x) { const long long v = h[i]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { smin[threadIdx.x] = min(smin[threadIdx.x], smin[threadIdx.x + o]); smax[threadIdx.x] = max(smax[threadIdx.x], smax[threadIdx.x + o]); }
        __syncthreads();
    }
    mn = smin[0]; mx = smax[0];
    double alpha = st[0], wmin = st[2];
    int halved = 0, converged = 0;
    if (mn > 1000 /* WL_MINHIST */) {
        if (temper * log((double)(mx / mn)) < 0.0001 /* WL_GERR; max/min is an INTEGER division in the reference */) {
            if (alpha < 1.0e-8 /* WL_ALPHATOL */) converged = 1;
            else {
                alpha /= 2;
                halved = 1;
                wmin = w[0];
                __syncthreads();
                for (int i = threadIdx.x; i < len; i += blockDim.x) { h[i] = 0; w[i] -= wmin; }
            }
        }
    }
    if (threadIdx.x == 0) { st[0] = alpha; st[1] = (double)mn; st[2] = wmin; st[3] = (double)mx; st[4] = (double)halved; st[5] = (double)converged; }
}

extern "C" int scgpu_wl_merge(scgpu_comm* c, int len, double* weights, int64_t* hist, double* weights_base, int64_t* hist_base,
                              double temper, scgpu_wlstate* st) {
    ARG(c && weights && hist && weights_base && hist_base && st, "scgpu_wl_merge: NULL argument");
    ARG(len > 0, "scgpu_wl_merge: len must be positive");
    CK(cudaSetDevice(c->device));
    if (len > c->wlm_cap) {
        cudaFree(c->d_w); cudaFree(c->d_wb); cudaFree(c->d_dw); cudaFree(c->d_h); cudaFree(c->d_hb); cudaFree(c->d_dh);
        c->d_w = c->d_wb = c->d_dw = nullptr; c->d_h = c->d_hb = c->d_dh = nullptr; c->wlm_cap = 0;
        CK(cudaMalloc(&c->d_w, (size_t)len * sizeof(double))); CK(cudaMalloc(&c->d_wb, (size_t)len * sizeof(double))); CK(cudaMalloc(&c->d_dw, (size_t)len * sizeof(double)));
        CK(cudaMalloc(&c->d_h, (size_t)len * sizeof(long long))); CK(cudaMalloc(&c->d_hb, (size_t)len * sizeof(long long))); CK(cudaMalloc(&c->d_dh, (size_t)len * sizeof(long long)));
        c->wlm_cap = len;
    }
    const size_t wb = (size_t)len * sizeof(double), hb = (size_t)len * sizeof(long long);
    CK(cudaMemcpyAsync(c->d_w, weights, wb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_wb, weights_base, wb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_h, hist, hb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_hb, hist_base, hb, cudaMemcpyHostToDevice, c->stream));
    double hs[8] = {st->alpha, 0, st->wmin, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(c->d_wlstate, hs, sizeof hs, cudaMemcpyHostToDevice, c->stream));
    const int nb = (len + 255) / 256;
    k_wl_delta<<<nb, 256, 0, c->stream>>>(len, c->d_w, c->d_wb, c->d_h, c->d_hb, c->d_dw, c->d_dh);
    if (c->nranks > 1) {
        NcclApi* n = nccl_api();
        NCK(n->AllReduce(c->d_dw, c->d_dw, (size_t)len, NCCL_FLOAT64, NCCL_SUM, c->nccl, c->stream));
        NCK(n->AllReduce(c->d_dh, c->d_dh, (size_t)len, NCCL_INT64, NCCL_SUM, c->nccl, c->stream));
    }
    k_wl_apply<<<nb, 256, 0, c->stream>>>(len, c->d_wb, c->d_hb, c->d_dw, c->d_dh);
    k_wl_update<<<1, 256, 0, c->stream>>>(len, c->d_wb, c->d_hb, temper, c->d_wlstate);
    CK(cudaGetLastError());
    if (c->nranks > 1)      // shared_A_min_wmin: rank 0's values are the ones everybody uses (identical by construction)
        NCK(nccl_api()->Broadcast(c->d_wlstate, c->d_wlstate, 8, NCCL_FLOAT64, 0, c->nccl, c->stream));
    CK(cudaMemcpyAsync(weights, c->d_wb, wb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hist, c->d_hb, hb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(hs, c->d_wlstate, sizeof hs, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(weights_base, weights, wb);
    memcpy(hist_base, hist, hb);
    st->alpha = hs[0]; st->min = (int64_t)hs[1]; st->wmin = hs[2]; st->max = (int64_t)hs[3]; st->halved = (int)hs[4]; st->converged = (int)hs[5];
    return SCGPU_OK;
}
