// scgpu.cu -- kernels + C ABI (include/scgpu.h) of the B200-native energy engine. sm_100a only.
//
// Data layout in HBM (per context; N particles, C cells):
//   d_api   [N][30] double   particles in ORIGINAL order, the C-ABI record (source of truth between rebuilds)
//   d_posw  [N]     double4  cell-sorted: fractional x,y,z and w = bit-packed {type | moltype<<8, original index}
//   d_rec   [N][32] double   cell-sorted internal records (pair_energy.cuh), 256 B = two full 128-B lines each
//   d_cell_start[C+1], d_order[N] (slot -> original), d_slot_of[N] (original -> slot), d_cell_of[N]
// N = 65 536 -> 2 MB + 16 MB: everything is L2-resident on B200 (126 MB), so the energy kernels are bound by
// the FP64 pipe, the cell build by HBM/L2 bandwidth and launch latency.
//
// Kernels:
//   k_cell_count (+ scan by its last block) / k_cell_fill / k_cell_place   counting sort by cell, stable (row C1)
//   k_gate_cheap<MODE> / k_patch / k_combine                    one-to-all energies in three dense phases (rows A1-A11, A13)
//   k_reduce_fixed                                             fixed-order total for allToAll
//   k_overlap                                                  warp-vote early exit (row A12)
//   k_sweep_colour                                             checkerboard displacement/rotation trials (row A14)
//   k_fp64_peak, k_flush                                       measurement helpers
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <type_traits>
#include <vector>
#include <algorithm>

#include "pair_energy.cuh"
#include "wall.cuh"
#include "wl_order.cuh"

using namespace scg;

static thread_local std::string g_err;
extern "C" const char* scgpu_last_error(void) { return g_err.c_str(); }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            g_err = b_;                                                                            \
            return SCGPU_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

#define ARG(cond, msg)                    \
    do {                                  \
        if (!(cond)) { g_err = msg; return SCGPU_ERR_ARG; } \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device view of a system
// ------------------------------------------------------------------------------------------------
struct DevSys {
    int n, ntypes, nmol;
    int nc[3];
    int ncells;
    const double* api;
    const double4* posw;
    const double* rec;
    const float4* p32;    // cell-sorted FP32 copy for the gates: x, y, z = position wrapped into [0, 1) (box fractions), w = original index | type << 24
    const float4* d32;    // cell-sorted FP32 axis (dir); w unused
    const int* cell_start;
    const int* fine_start;   // sub > 1: every cell is cut into sub^3 sub-cells and sorted by them; [cell * sub^3 + (sz * sub + sy) * sub + sx] -> first slot
    int sub;                 // (== cell_start when sub == 1)
    const int* order;
    const int* slot_of;
    const int* type;
    const int* moltype;
    const scgpu_iaparam* ia;
    const scgpu_molparam* mol;
    const float* reach2;  // per ordered type pair: squared centre distance beyond which the pair energy is exactly 0 (x 1.001);
                          // [T*T .. T*T+T): the largest of each row; [T*T+T .. 2*T*T+T): cut2 = max(rcut, rcutwca)^2 x 1.001 of rod pairs
                          // (0 for other kinds: no segment bound applies); [2*T*T+T .. 2*T*T+2*T): half length of each type (0 = sphere)
    double box[3];
    double shift[3];     // fractional grid shift used for the current cell assignment (0 for the energy API)
    double sqmaxcut;
    const WallParam* wall;   // [EXTER] wall: per particle type, or nullptr when the topology has none
    double exter_sqmaxcut;
};

__device__ __forceinline__ double pack_w(int type, int moltype, int orig) { return __hiloint2double(type | (moltype << 8), orig); }
__device__ __forceinline__ int w_type(double w) { return __double2hiint(w) & 0xff; }
__device__ __forceinline__ int w_moltype(double w) { return __double2hiint(w) >> 8; }
__device__ __forceinline__ int w_orig(double w) { return __double2loint(w); }

// cell coordinate, definition C1: f = INBOX(u) (scOOP/structures/macros.h:119, as used by
// Mesh::addPart, scOOP/mc/mesh.cpp:49-54), c = (int)(f*ncell), c == ncell -> 0
__host__ __device__ __forceinline__ int cell_coord(double u, int nc) {
    double ip;
    double f = (u > 0) ? modf(u, &ip) : modf(u, &ip) + 1;
    int c = (int)(f * nc);
    if (c == nc) c = 0;
    return c;
}
// the same with the sub-cell (0 .. S-1) the particle sits in along this axis
__host__ __device__ __forceinline__ int cell_coord_sub(double u, int nc, int S, int& sx) {
    double ip;
    double f = (u > 0) ? modf(u, &ip) : modf(u, &ip) + 1;
    const double t = f * nc;
    int c = (int)t;
    sx = (int)((t - (double)c) * S);
    if (sx >= S) sx = S - 1;
    if (sx < 0) sx = 0;
    if (c == nc) { c = 0; sx = 0; }
    return c;
}
__host__ __device__ __forceinline__ int fine_index(const double* pos, const double* shift, const int* nc, int S, int* coarse) {
    int sx, sy, sz;
    const int cx = cell_coord_sub(pos[0] + shift[0], nc[0], S, sx);
    const int cy = cell_coord_sub(pos[1] + shift[1], nc[1], S, sy);
    const int cz = cell_coord_sub(pos[2] + shift[2], nc[2], S, sz);
    const int c = (cz * nc[1] + cy) * nc[0] + cx;
    if (coarse) *coarse = c;
    return c * (S * S * S) + (sz * S + sy) * S + sx;
}
__host__ __device__ __forceinline__ int cell_index(const double* pos, const double* shift, const int* nc) {
    int cx = cell_coord(pos[0] + shift[0], nc[0]);
    int cy = cell_coord(pos[1] + shift[1], nc[1]);
    int cz = cell_coord(pos[2] + shift[2], nc[2]);
    return (cz * nc[1] + cy) * nc[0] + cx;
}

// C-ABI record (30) -> internal record (32): internal[k] = api[API_OF[k]]
__constant__ int c_api_of[30] = {3, 4, 5, 6, 7, 8, 12, 13, 14, 15, 16, 17, 9, 10, 11, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 0, 1, 2};
static const int h_api_of[30] = {3, 4, 5, 6, 7, 8, 12, 13, 14, 15, 16, 17, 9, 10, 11, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 0, 1, 2};

// ------------------------------------------------------------------------------------------------
// cell list: counting sort by cell, stable in the original index
// ------------------------------------------------------------------------------------------------
__device__ void cell_scan_block(int ncells, int* __restrict__ counts, int* __restrict__ cell_start, int* __restrict__ cursor);
// one thread per particle: cell (and sub-cell) index, histogram. The block that finishes last turns the histogram into the exclusive
// scan (one launch instead of memset + count + scan; ticket: one unsigned the last block leaves at zero again).
__global__ void __launch_bounds__(1024)
k_cell_count(DevSys s, int nsort, int* __restrict__ cell_of, int* __restrict__ fine_of, int* __restrict__ counts,
             int* __restrict__ sort_start, int* __restrict__ cursor, unsigned* __restrict__ ticket) {
    __shared__ bool last;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < s.n) {
        int c;
        const int f = s.sub > 1 ? fine_index(s.api + (size_t)i * 30, s.shift, s.nc, s.sub, &c) : (c = cell_index(s.api + (size_t)i * 30, s.shift, s.nc));
        cell_of[i] = c;
        if (s.sub > 1) fine_of[i] = f;
        atomicAdd(&counts[f], 1);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    cell_scan_block(nsort, counts, sort_start, cursor);
    if (threadIdx.x == 0) *ticket = 0u;
}
// sub > 1: the start of every (coarse) cell out of the sub-cell starts
__global__ void k_coarse_start(int ncells, int s3, const int* __restrict__ fine_start, int* __restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= ncells) cell_start[c] = fine_start[c * s3];
    if (c == ncells + 1) cell_start[c] = 0;
}

// single-block exclusive scan of counts[0..ncells) -> cell_start[0..ncells]; cursor := cell_start
__device__ void cell_scan_block(int ncells, int* __restrict__ counts, int* __restrict__ cell_start, int* __restrict__ cursor) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ncells; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < ncells) ? __ldcg(counts + i) : 0;        // written by other blocks' atomics: read at L2
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        int excl = carry + (wid ? warp_tot[wid - 1] : 0) + x - v;
        if (i < ncells) { cell_start[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_start[ncells] = carry;
    // [ncells + 1]: number of non-empty cells (the sweep kernel spreads its trials over them); the counts are consumed here and left at
    // zero for the next build (no memset launch)
    __syncthreads();
    int ne = 0;
    for (int i = threadIdx.x; i < ncells; i += blockDim.x) { ne += __ldcg(counts + i) > 0; counts[i] = 0; }
    ne = __reduce_add_sync(0xffffffffu, ne);
    if (lane == 0) warp_tot[wid] = ne;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += warp_tot[w]; cell_start[ncells + 1] = t; }
}

constexpr int HV_MAX = 4096;      // heavy-type particles the sorted list holds (beyond that the sub-cell gate is not used)
__global__ void k_heavy_collect(int n, const double4* __restrict__ posw, unsigned heavy, int* __restrict__ list, int* __restrict__ count) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    if ((heavy >> (__double2hiint(posw[slot].w) & 0xff)) & 1u) { const int p = atomicAdd(count, 1); if (p < HV_MAX) list[p] = slot; }
}
// one block: ascending slots (bitonic, padded with INT_MAX) -> the list does not depend on the order the atomics came in
__global__ void __launch_bounds__(1024) k_heavy_sort(int* __restrict__ list, int* __restrict__ count) {
    __shared__ int v[HV_MAX];
    const int n = min(count[0], HV_MAX);
    for (int k = threadIdx.x; k < HV_MAX; k += blockDim.x) v[k] = k < n ? list[k] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= HV_MAX; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int k = threadIdx.x; k < HV_MAX / 2; k += blockDim.x) {
                const int i = 2 * k - (k & (stride - 1)), j = i + stride;
                const bool up = (i & size) == 0;
                const int a = v[i], b = v[j];
                if ((a > b) == up) { v[i] = b; v[j] = a; }
            }
            __syncthreads();
        }
    for (int k = threadIdx.x; k < n; k += blockDim.x) list[k] = v[k];
    if (threadIdx.x == 0) count[1] = count[0] > HV_MAX ? 1 : 0;      // overflow: the host falls back to the cell gate
}

// unordered fill of each cell's segment with original indices (order fixed afterwards by k_cell_place)
__global__ void k_cell_fill(int n, const int* __restrict__ cell_of, int* __restrict__ cursor, int* __restrict__ tmp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = atomicAdd(&cursor[cell_of[i]], 1);
    tmp[p] = i;
}

// stable placement: slot = cell_start[c] + #{members of c with a smaller original index}; permute the records
__device__ __forceinline__ float4 make_p32(double x, double y, double z, int orig, int type) {
    return make_float4((float)(x - floor(x)), (float)(y - floor(y)), (float)(z - floor(z)), __int_as_float(orig | (type << 24)));
}
// FP32 copies of the sorted positions / axes (the gates stage these: 16 contiguous bytes per candidate, no FP64 arithmetic)
__global__ void k_make_f32(int n, const double4* __restrict__ posw, const double* __restrict__ rec, float4* __restrict__ p32, float4* __restrict__ d32) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const double4 pw = posw[slot];
    const double4 dw = ldg256(rec + (size_t)slot * REC + R_DIR);
    p32[slot] = make_p32(pw.x, pw.y, pw.z, w_orig(pw.w), w_type(pw.w));
    d32[slot] = make_float4((float)dw.x, (float)dw.y, (float)dw.z, 0.f);
}

__global__ void k_cell_place(DevSys s, const int* __restrict__ cell_of, const int* __restrict__ tmp,
                             int* __restrict__ order, int* __restrict__ slot_of, double4* __restrict__ posw, double* __restrict__ rec,
                             float4* __restrict__ p32, float4* __restrict__ d32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < s.n;
    int slot = 0;
    if (valid) {
        int c = cell_of[i];          // (the caller passes the sub-cell index of every particle when sub > 1)
        int b = s.fine_start[c], e = s.fine_start[c + 1];
        int r = 0;
        for (int k = b; k < e; k++) r += (tmp[k] < i);
        slot = b + r;
        order[slot] = i;
        slot_of[i] = slot;
    }
    // record permute, warp-cooperative: for each of the warp's 32 particles the 32 lanes copy one 240-byte record together, so every
    // load and store instruction touches one or two contiguous cache lines instead of 32 scattered ones
    const int lane = threadIdx.x & 31;
    if (!valid) slot = -1;
    const int perm = lane < 30 ? c_api_of[lane] : 0;
    for (int j0 = 0; j0 < 32; j0 += 8) {          // 8 records in flight per lane: loads first, then stores
        double v[8];
        int dst[8], srcs[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            srcs[u] = __shfl_sync(0xffffffffu, i, j0 + u);
            dst[u] = __shfl_sync(0xffffffffu, slot, j0 + u);
            v[u] = (dst[u] >= 0 && lane < 30) ? s.api[(size_t)srcs[u] * 30 + perm] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (dst[u] < 0) continue;
            rec[(size_t)dst[u] * REC + lane] = v[u];
            if (lane == 31) {
                const double* a = s.api + (size_t)srcs[u] * 30;
                const int ty = s.type[srcs[u]];
                posw[dst[u]] = make_double4(a[0], a[1], a[2], pack_w(ty, s.moltype[srcs[u]], srcs[u]));
                p32[dst[u]] = make_p32(a[0], a[1], a[2], srcs[u], ty);
            }
            if (lane == 30) {
                const double* a = s.api + (size_t)srcs[u] * 30;
                d32[dst[u]] = make_float4((float)a[3], (float)a[4], (float)a[5], 0.f);
            }
        }
    }
}

// Particle::init on the device (scOOP/structures/particle.cpp:3-79): derive patch sides, second patch and chiral axes from
// (dir, patchdir) and the type's own parameters -- what Conf::partVecInit does after config.init has been read. Lets a caller
// upload the 9 doubles per particle that config.init holds instead of the 30-double record.
__device__ __forceinline__ void rot_q(const double* p, const double* axis, double c, double sn, double* out) {   // Vector::rotate (Vector.h:138-160)
    double qw = c, qx = axis[0] * sn, qy = axis[1] * sn, qz = axis[2] * sn;
    double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy, t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t10 = -qz * qz;
    double x = p[0], y = p[1], z = p[2];
    out[0] = 2.0 * ((t8 + t10) * x + (t6 - t4) * y + (t3 + t7) * z) + x;
    out[1] = 2.0 * ((t4 + t6) * x + (t5 + t10) * y + (t9 - t2) * z) + y;
    out[2] = 2.0 * ((t7 - t3) * x + (t2 + t9) * y + (t5 + t8) * z) + z;
}
__device__ __forceinline__ void normalise3(double* v) {
    double tot = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (tot != 0.0) { tot = 1.0 / tot; v[0] *= tot; v[1] *= tot; v[2] *= tot; }
}
__device__ __forceinline__ void ortho3(double* a, const double* b) {
    double dp = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    a[0] -= dp * b[0]; a[1] -= dp * b[1]; a[2] -= dp * b[2];
}
__global__ void k_particle_init(int n, int ntypes, const double* __restrict__ compact9, const int* __restrict__ type,
                                const scgpu_iaparam* __restrict__ ia_tab, double* __restrict__ api) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double st[30];
#pragma unroll
    for (int k = 0; k < 9; k++) st[k] = compact9[(size_t)i * 9 + k];
#pragma unroll
    for (int k = 9; k < 30; k++) st[k] = 0.0;
    const scgpu_iaparam& ia = ia_tab[type[i] * ntypes + type[i]];
    const int g = (int)ia.geotype[0];
    if (g < SCGPU_SPN && g != SCGPU_SCA && g != SCGPU_SCN) {
        double* dir = st + 3; double* pd0 = st + 6; double* pd1 = st + 9;
        normalise3(dir);
        ortho3(pd0, dir);
        normalise3(pd0);
        const bool two = is_two_patch(g), chiral = is_chiral(g);
        if (g == SCGPU_PSC || g == SCGPU_CPSC || g == SCGPU_TPSC || g == SCGPU_TCPSC) {
            rot_q(pd0, dir, ia.pcoshalfi[0], ia.psinhalfi[0], st + 12);
            rot_q(pd0, dir, ia.pcoshalfi[0], -1.0 * ia.psinhalfi[0], st + 15);
        }
        if (two) {
            double tmp[3];
            rot_q(pd0, dir, ia.csecpatchrot[0], ia.ssecpatchrot[0], tmp);
            ortho3(tmp, dir);
            normalise3(tmp);
            pd1[0] = tmp[0]; pd1[1] = tmp[1]; pd1[2] = tmp[2];
        }
        if (g == SCGPU_TPSC || g == SCGPU_TCPSC) {
            rot_q(pd1, dir, ia.pcoshalfi[2], ia.psinhalfi[2], st + 18);
            rot_q(pd1, dir, ia.pcoshalfi[2], -1.0 * ia.psinhalfi[2], st + 21);
        }
        if (chiral) {
            rot_q(dir, pd0, ia.chiral_cos[0], ia.chiral_sin[0], st + 24);
            rot_q(pd0, st + 24, ia.pcoshalfi[0], ia.psinhalfi[0], st + 12);
            rot_q(pd0, st + 24, ia.pcoshalfi[0], -1.0 * ia.psinhalfi[0], st + 15);
        }
        if (g == SCGPU_TCHPSC || g == SCGPU_TCHCPSC) {
            rot_q(dir, pd1, ia.chiral_cos[0], ia.chiral_sin[0], st + 27);
            rot_q(pd1, st + 27, ia.pcoshalfi[2], ia.psinhalfi[2], st + 18);
            rot_q(pd1, st + 27, ia.pcoshalfi[2], -1.0 * ia.psinhalfi[2], st + 21);
        }
    }
#pragma unroll
    for (int k = 0; k < 30; k++) api[(size_t)i * 30 + k] = st[k];
}

// rewrite one particle's sorted record in place (update(int target) when it stayed in its cell)
__global__ void k_update_one(DevSys s, int idx, double4* __restrict__ posw, double* __restrict__ rec, float4* __restrict__ p32, float4* __restrict__ d32) {
    int k = threadIdx.x;
    int slot = s.slot_of[idx];
    const double* a = s.api + (size_t)idx * 30;
    if (k < 30) rec[(size_t)slot * REC + k] = a[c_api_of[k]];
    if (k == 31) { posw[slot] = make_double4(a[0], a[1], a[2], pack_w(s.type[idx], s.moltype[idx], idx)); p32[slot] = make_p32(a[0], a[1], a[2], idx, s.type[idx]); }
    if (k == 30) d32[slot] = make_float4((float)a[3], (float)a[4], (float)a[5], 0.f);
}

// sorted -> original order (after device-side sweeps changed the sorted arrays)
__global__ void k_unsort(DevSys s, double* __restrict__ api) {
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= s.n) return;
    int i = s.order[slot];
    const double* r = s.rec + (size_t)slot * REC;
    double* a = api + (size_t)i * 30;
#pragma unroll
    for (int k = 0; k < 30; k++) a[c_api_of[k]] = r[k];
}

// ------------------------------------------------------------------------------------------------
// one-to-all energy: the heart of the hot path, in three launches
//   k_gate_cheap  one warp-group per trial particle: scan the 27 neighbour cells, warp-ballot-compact the candidates that
//                 pass the sqmaxcut gate, evaluate everything EXCEPT the rod-rod patch attraction on dense 32-lane batches,
//                 and compact the (few) pairs that owe a patch evaluation into a global work list
//   k_patch       one thread per listed pair: the patch geometry + attraction (the expensive ~10 %), densely packed
//   k_combine     per trial particle: partial + its patch terms, summed in list order (deterministic)
// ------------------------------------------------------------------------------------------------
constexpr int QCAP = 96;      // per-warp gated-candidate queue (ints, shared memory): < 32 left over + up to 64 new
constexpr int PCAP = 64;      // per-warp patch-pair buffer (ints, shared memory), flushed to the global list in chunks
// a warp that collects more than PCAP patch partners for one particle flushes several chunks; they are chained

struct Filter {
    int self;        // original index never paired with itself
    int excl_lo, excl_hi;   // [lo,hi) original indices skipped (molecule members for mol2others)
    int max_idx;     // only partners with original index < max_idx (allToAll rows); INT_MAX otherwise
};

__device__ __forceinline__ bool filt(const Filter& f, int orig) {
    return orig != f.self && !(orig >= f.excl_lo && orig < f.excl_hi) && orig < f.max_idx;
}

__device__ __forceinline__ double warp_sum(double v) {
    // fixed butterfly order -> bitwise reproducible
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct PatchList {            // global work list of pairs that owe pair_energy_patch()
    int2* pair;               // x: slot of the first particle (>= 0) or -1-t for a trial record; y: slot of the second
    double* e;                // result of k_patch, same index
    int* total;               // running fill (atomicAdd); reset by k_combine
    int cap;
    int* warp_head;           // per warp: id of its first chunk, -1 if it has none
    int4* chunks;             // {start in pair[]/e[], length, id of the warp's next chunk or -1, unused}
    int* chunk_count;         // running fill of chunks[] (atomicAdd); reset by k_combine
    int chunk_cap;
    int* overflow;            // sticky flag: a list was too small -> results invalid, the host grows the lists and repeats
    double* trial_rec;        // [m][REC] internal records of trial states (only when the caller passed trial states)
};

// Energy of one (trial) particle against the 27-cell neighbourhood + its bonded partners, WITHOUT rod-rod patch terms.
// s1: this particle's internal record in shared memory. Executed by `gw` cooperating warps (group warp id `gwid`).
// Returns this WARP's partial sum (identical in all lanes). queue / pbuf: shared memory private to the warp.
// first_ref: what k_patch needs to find this particle's record (slot >= 0, or -1-t for a trial record).
template <bool RODS>
__device__ double warp_gate_cheap(const DevSys& s, const double* s1, int type1, int moltype1, const ConList& cl, const Filter& f,
                                  int gw, int gwid, int* queue, int* pbuf, int first_ref, int warp_global, const PatchList& pl,
                                  double* e_pairs, unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const v3 p1 = ld3(s1 + R_POS);
    const int cx = cell_coord(p1.x + s.shift[0], s.nc[0]);
    const int cy = cell_coord(p1.y + s.shift[1], s.nc[1]);
    const int cz = cell_coord(p1.z + s.shift[2], s.nc[2]);
    const int nx = s.nc[0] == 1 ? 1 : 3, ny = s.nc[1] == 1 ? 1 : 3, nz = s.nc[2] == 1 ? 1 : 3;
    const int ncell_nb = nx * ny * nz;
    double acc = 0.0;
    int qn = 0, pn = 0, last_chunk = -1, head = -1;
    unsigned n_cand = 0, n_gate = 0;

    // hand the buffered patch pairs to k_patch as one contiguous run of the global list
    auto flush_chunk = [&]() {
        int base = 0, cid = 0;
        if (lane == 0) { base = atomicAdd(pl.total, pn); cid = atomicAdd(pl.chunk_count, 1); }
        base = __shfl_sync(0xffffffffu, base, 0);
        cid = __shfl_sync(0xffffffffu, cid, 0);
        bool ok = (cid < pl.chunk_cap) && (base + pn <= pl.cap);
        if (ok) {
            for (int k = lane; k < pn; k += 32) pl.pair[base + k] = make_int2(first_ref, pbuf[k]);
            if (lane == 0) {
                pl.chunks[cid] = make_int4(base, pn, -1, 0);
                if (last_chunk >= 0) pl.chunks[last_chunk].z = cid;
            }
            if (head < 0) head = cid;
            last_chunk = cid;
        } else if (lane == 0) atomicOr(pl.overflow, 1);
        pn = 0;
        __syncwarp();
    };
    auto eval_slot = [&](int slot, bool on) {
        bool np = false;
        if (on) {
            double4 pw = s.posw[slot];
            v3 r_cm = image(s.box, p1, mk(pw.x, pw.y, pw.z));
            double dotrcm = dot(r_cm, r_cm);
            int orig = w_orig(pw.w);
            double e = pair_energy_cheap<RODS>(s.box, s.ia, s.ntypes, s.mol, r_cm, dotrcm, s1, type1, moltype1,
                                         s.rec + (size_t)slot * REC, w_type(pw.w), orig, cl, np);
            if (e_pairs) e_pairs[orig] = e;
            acc += e;
            n_gate++;
        }
        unsigned m = __ballot_sync(0xffffffffu, np);
        if (m) {
            int c = __popc(m);
            if (pn + c > PCAP) flush_chunk();
            if (np) pbuf[pn + __popc(m & lt_mask)] = slot;
            pn += c;
            __syncwarp();
        }
    };

    for (int k = gwid; k < ncell_nb; k += gw) {
        int dx = k % nx, dy = (k / nx) % ny, dz = k / (nx * ny);
        int ccx = nx == 1 ? 0 : (cx + dx - 1 + s.nc[0]) % s.nc[0];
        int ccy = ny == 1 ? 0 : (cy + dy - 1 + s.nc[1]) % s.nc[1];
        int ccz = nz == 1 ? 0 : (cz + dz - 1 + s.nc[2]) % s.nc[2];
        int c = (ccz * s.nc[1] + ccy) * s.nc[0] + ccx;
        int b = s.cell_start[c], e = s.cell_start[c + 1];
        for (int base = b; base < e; base += 32) {
            int j = base + lane;
            bool pass = false;
            if (j < e) {
                double4 pw = s.posw[j];
                int orig = w_orig(pw.w);
                bool bonded = !RODS && (orig == cl.con[0] || orig == cl.con[1] || orig == cl.con[2] || orig == cl.con[3]);
                if (filt(f, orig) && !bonded) {      // bonded partners are handled (and counted) once, below
                    n_cand++;
                    v3 r_cm = image(s.box, p1, mk(pw.x, pw.y, pw.z));
                    pass = (dot(r_cm, r_cm) <= s.sqmaxcut);   // PairE gate (mc/paire.h:1214)
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) queue[qn + __popc(m & lt_mask)] = j;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                eval_slot(queue[lane], true);
                int rest = qn - 32;
                int mv = (lane < rest) ? queue[32 + lane] : 0;
                __syncwarp();
                if (lane < rest) queue[lane] = mv;
                qn = rest;
                __syncwarp();
            }
        }
    }
    if (qn > 0) eval_slot(lane < qn ? queue[lane] : 0, lane < qn);
    // bonded partners: never gated (conlist not empty, mc/paire.h:1214), wherever they are
    if (!RODS && gwid == 0 && !cl.is_empty) {
        int orig = lane < 4 ? cl.con[lane] : -1;
        bool on = orig >= 0 && filt(f, orig);
        if (on) n_cand++;
        eval_slot(on ? s.slot_of[orig] : 0, on);
    }
    if (pn > 0) flush_chunk();
    if (lane == 0) pl.warp_head[warp_global] = head;
    if (counters) {
        n_cand = __reduce_add_sync(0xffffffffu, n_cand);
        n_gate = __reduce_add_sync(0xffffffffu, n_gate);
        if (lane == 0) { atomicAdd(&counters[0], (unsigned long long)n_cand); atomicAdd(&counters[1], (unsigned long long)n_gate); }
    }
    return warp_sum(acc);
}

constexpr int OTA_THREADS = 128;
constexpr int OTA_WARPS = OTA_THREADS / 32;

// MODE 0: targets[] (+ optional trial states in C-ABI layout); MODE 1: every particle, own state, all partners;
// MODE 2: allToAll rows (partners with a smaller original index); MODE 3: molecule members vs non-members, empty conlist
template <int MODE, bool RODS>
__global__ void __launch_bounds__(OTA_THREADS, RODS ? 5 : 3)
k_gate_cheap(DevSys s, int m, int gw, const int* __restrict__ targets, const double* __restrict__ trial_states,
             int excl_lo, int excl_hi, PatchList pl, double* __restrict__ warp_partial, double* __restrict__ e_pairs,
             unsigned long long* counters) {
    __shared__ double sh_rec[OTA_WARPS][REC];
    __shared__ int sh_queue[OTA_WARPS][QCAP];
    __shared__ int sh_pbuf[OTA_WARPS][PCAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int groups_per_block = OTA_WARPS / gw;
    const int g = wid / gw, gwid = wid % gw;
    const int t = blockIdx.x * groups_per_block + g;
    const bool active = t < m;
    int target = 0;
    if (active) target = (MODE == 0) ? targets[t] : (MODE == 3 ? excl_lo + t : t);
    double* rec1 = sh_rec[g * gw];
    const bool trial = (MODE == 0 || MODE == 3) && trial_states != nullptr;
    if (active && gwid == 0) {
        if (trial) {
            double v = (lane < 30) ? trial_states[(size_t)t * 30 + c_api_of[lane]] : 0.0;
            rec1[lane] = v;
            pl.trial_rec[(size_t)t * REC + lane] = v;      // k_patch reads the trial record from here
        } else {
            rec1[lane] = s.rec[(size_t)s.slot_of[target] * REC + lane];
        }
    }
    __syncthreads();
    if (!active) return;
    int type1 = s.type[target], moltype1 = s.moltype[target];
    ConList cl;
    if (MODE == 3 || RODS) { cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0; }
    else get_conlist(s.mol, moltype1, target, cl);
    Filter f;
    f.self = target;
    f.excl_lo = (MODE == 3) ? excl_lo : 0;
    f.excl_hi = (MODE == 3) ? excl_hi : 0;
    f.max_idx = (MODE == 2) ? target : 0x7fffffff;
    const int first_ref = trial ? (-1 - t) : s.slot_of[target];
    const int wg = t * gw + gwid;
    double part = warp_gate_cheap<RODS>(s, rec1, type1, moltype1, cl, f, gw, gwid, sh_queue[wid], sh_pbuf[wid], first_ref, wg, pl, e_pairs, counters);
    if (lane == 0) warp_partial[wg] = part;
}

__device__ __forceinline__ double rel_frac(double u, double cen) { double d = u - cen; return d - rint(d); }

// ------------------------------------------------------------------------------------------------
// FLAT pipeline for "every particle" passes (MODE 1 one-to-all of everyone, MODE 2 allToAll rows): four launches, each dense
//   k_gate_cells    block per cell; FP32 conservative pre-gate over the staged neighbourhood; survivors -> global pair list
//                   (per-particle chunk chains, so that the final sum has a defined order)
//   k_cheap_flat    thread per listed pair: exact FP64 gate + everything but the rod-rod patch term; pairs that owe a
//                   patch term are appended (warp-aggregated) to a second list
//   k_patch_flat    thread per patch pair
//   k_combine_flat  per particle: sum of (cheap + patch) over its chunk chain, in list order
// Compared with evaluating inside the gate kernel this keeps the gate kernel at ~40 registers (3x the resident warps) and
// gives the FP64 work perfectly packed lanes.
// ------------------------------------------------------------------------------------------------
struct FlatList {
    int2* pair;          // (slot of the first particle, slot of the second)
    double2* e;          // {cheap part, patch part}
    int* total; int cap;
    int* head;           // per particle (original index): first chunk id or -1
    int4* chunks; int* chunk_count; int chunk_cap;
    int* plist; int* ptotal;     // indices into pair[] that owe a patch term
    int* overflow;
    int* heavy;                  // bit t set by k_gate_rows_gen: a target of type t had more partners than its shared-memory buffer holds
    unsigned heavy_types;        // types handled by k_gate_cells (targets of the other types by k_gate_rows_gen); 0 = no split
};

constexpr int GT_TILE = 1024;     // staged candidates per tile of a cell neighbourhood: 1024 x 20 B = 20 KB
constexpr int GT_WARPS = 4;
constexpr int GT_NMASK = 512;     // 64-bit gate masks kept from the counting pass (4 KB): the writing pass replays them instead of re-testing
constexpr int GT_MAXP = 256;      // particles of one cell handled as one batch (counts and write cursors live in shared memory)

// WRAP: some axis has fewer than 5 cells, so a neighbour can be more than half a box away from the target and the FP32
// separation needs the minimum-image fold. With >= 5 cells per axis every candidate of the 27-cell neighbourhood is within
// 2/5 of the box of the target once both are taken relative to the cell centre, and the fold is skipped.
//
// One block per cell, two passes over the same staged neighbourhood: pass 0 COUNTS the listed partners of every particle
// of the cell, the block then reserves one contiguous span of the global list with a single atomicAdd (thousands of
// same-address atomics per launch instead of hundreds of thousands -- they serialise in L2), and pass 1 repeats the scan and
// writes each particle's partners, in candidate order, into its own sub-span. The FP32 scan is a few instructions per
// candidate, so running it twice is cheaper than any queue + flush scheme; spans make the per-particle sum trivial.
template <int MODE, bool RODS, bool WRAP>
__global__ void __launch_bounds__(GT_WARPS * 32, 8)
k_gate_cells(DevSys s, FlatList fl, unsigned long long* counters) {
    __shared__ float4 t_pf[GT_TILE];      // x,y,z: FP32 coordinates relative to the cell centre, in length units; w: original index | type << 24
    __shared__ int t_slot[GT_TILE];
    __shared__ int sh_cnt[GT_MAXP], sh_pos[GT_MAXP];     // per particle of the current batch: listed partners, write cursor
    __shared__ unsigned long long sh_mask[GT_NMASK];
    __shared__ unsigned sh_grp[GT_TILE / 64];          // per 64-candidate group of the staged tile: which neighbour cells it holds
    __shared__ unsigned sh_ctypes[28];                 // per neighbour cell: bit t set = a particle of type t lives there
    __shared__ unsigned sh_tcm[GT_MAXP];               // per particle of the batch: neighbour cells within its reach (computed once)
    __shared__ unsigned sh_tilecells;                  // neighbour cells with entries in the staged tile
    __shared__ int sh_b[28], sh_off[28];
    __shared__ int sh_ok;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int c0 = blockIdx.x;
    const int tb = s.cell_start[c0], te = s.cell_start[c0 + 1];
    if (tb == te) return;
    if (fl.heavy_types) {        // split launch: only the targets of the heavy types are this kernel's; most cells hold none
        int any = 0;
        for (int k = tb + (int)threadIdx.x; k < te; k += blockDim.x) any |= (fl.heavy_types >> w_type(s.posw[k].w)) & 1u;
        if (!__syncthreads_or(any)) return;
    }
    const int cx = c0 % s.nc[0], cy = (c0 / s.nc[0]) % s.nc[1], cz = c0 / (s.nc[0] * s.nc[1]);
    const int nx = s.nc[0] == 1 ? 1 : 3, ny = s.nc[1] == 1 ? 1 : 3, nz = s.nc[2] == 1 ? 1 : 3;
    const int ncell_nb = nx * ny * nz;
    if (wid == 0) {
        int len = 0, b = 0;
        if (lane < ncell_nb) {
            int dx = lane % nx, dy = (lane / nx) % ny, dz = lane / (nx * ny);
            int ccx = nx == 1 ? 0 : (cx + dx - 1 + s.nc[0]) % s.nc[0];
            int ccy = ny == 1 ? 0 : (cy + dy - 1 + s.nc[1]) % s.nc[1];
            int ccz = nz == 1 ? 0 : (cz + dz - 1 + s.nc[2]) % s.nc[2];
            int c = (ccz * s.nc[1] + ccy) * s.nc[0] + ccx;
            b = s.cell_start[c];
            len = s.cell_start[c + 1] - b;
        }
        int x = len;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < 28) { sh_b[lane] = b; sh_off[lane] = x - len; }
    }
    __syncthreads();
    const int C = sh_off[ncell_nb];
    const int ntiles = (C + GT_TILE - 1) / GT_TILE;      // neighbourhoods larger than one tile are processed tile by tile
    if (!RODS || ntiles > 1) {       // culling data: the particle types present in every neighbour cell (types >= 32: assume all)
        for (int k = wid; k < ncell_nb; k += GT_WARPS) {
            const int b = sh_b[k], len = sh_off[k + 1] - sh_off[k];
            unsigned m = 0;
            for (int idx = lane; idx < len; idx += 32) { const int ty = w_type(s.posw[b + idx].w); m |= ty < 32 ? 1u << ty : 0xffffffffu; }
            m = __reduce_or_sync(0xffffffffu, m);
            if (lane == 0) sh_ctypes[k] = s.ntypes > 32 ? 0xffffffffu : m;
        }
        __syncthreads();
    }
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    const float ibox[3] = {(float)(1.0 / s.box[0]), (float)(1.0 / s.box[1]), (float)(1.0 / s.box[2])};
    const float cut_hi = (float)(s.sqmaxcut * 1.001), cut_lo = (float)(s.sqmaxcut * 0.999);
    const float hcell[3] = {(float)(0.5 * s.box[0] / s.nc[0]), (float)(0.5 * s.box[1] / s.nc[1]), (float)(0.5 * s.box[2] / s.nc[2])};
    auto staged = [&](int slot) {
        double4 pw = s.posw[slot];
        return make_float4((float)(rel_frac(pw.x + s.shift[0], ccen[0]) * s.box[0]), (float)(rel_frac(pw.y + s.shift[1], ccen[1]) * s.box[1]),
                           (float)(rel_frac(pw.z + s.shift[2], ccen[2]) * s.box[2]), __int_as_float(w_orig(pw.w) | (w_type(pw.w) << 24)));
    };
    int staged_tile = -1;
    auto stage = [&](int tile) {        // block-wide; one warp per neighbour cell: contiguous 32-byte loads, no index search
        if (tile == staged_tile) return;
        const int t0 = tile * GT_TILE, TC = min(GT_TILE, C - t0);
        __syncthreads();                // previous tile fully consumed
        for (int k = wid; k < ncell_nb; k += GT_WARPS) {
            const int b = sh_b[k], off = sh_off[k], len = sh_off[k + 1] - off;
            const int lo = max(off, t0), hi = min(off + len, t0 + TC);
            for (int p = lo + lane; p < hi; p += 32) {
                t_pf[p - t0] = staged(b + (p - off));
                t_slot[p - t0] = b + (p - off);
            }
        }
        for (int p = TC + threadIdx.x; p < ((TC + 63) & ~63); p += blockDim.x) {     // padding: NaN coordinates fail every comparison of the gate
            const float qnan = __int_as_float(0x7fc00000);
            t_pf[p] = make_float4(qnan, qnan, qnan, __int_as_float(0xffffff));
            t_slot[p] = 0;
        }
        if (threadIdx.x < GT_TILE / 64) {
            const int g0 = t0 + 64 * (int)threadIdx.x, g1 = g0 + 64;
            unsigned m = 0;
            for (int k = 0; k < ncell_nb; k++) if (sh_off[k + 1] > sh_off[k] && sh_off[k] < g1 && sh_off[k + 1] > g0) m |= 1u << k;
            sh_grp[threadIdx.x] = m;
        }
        if (threadIdx.x == 0) {
            unsigned m = 0;
            for (int k = 0; k < ncell_nb; k++) if (sh_off[k + 1] > sh_off[k] && sh_off[k] < t0 + TC && sh_off[k + 1] > t0) m |= 1u << k;
            sh_tilecells = m;
        }
        staged_tile = tile;
        __syncthreads();
    };
    // the block's own cell is the centre of its neighbourhood: where its particles sit in the (first) staged tile
    const int centre_off = sh_off[(nz == 1 ? 0 : nx * ny) + (ny == 1 ? 0 : nx) + (nx == 1 ? 0 : 1)];
    const bool count = counters != nullptr;
    for (int bt = tb; bt < te; bt += GT_MAXP) {          // batches of particles of this cell (one batch unless the cell is huge)
        const int nb = min(GT_MAXP, te - bt);
        const int W = (C + 63) >> 6;                               // 64-candidate groups of a single-tile neighbourhood
        const bool use_masks = ntiles == 1 && nb * W <= GT_NMASK;  // the usual case; otherwise pass 1 re-tests
        __syncthreads();
        for (int k = threadIdx.x; k < nb; k += blockDim.x) sh_cnt[k] = 0;
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {
            for (int tile = 0; tile < ntiles; tile++) {
                stage(tile);
                const int TC = min(GT_TILE, C - tile * GT_TILE);
                for (int ti = bt + wid; ti < bt + nb; ti += GT_WARPS) {
                    if (RODS && pass && use_masks) {       // replay the masks of the counting pass: nothing about the target is needed
                        int cur = sh_pos[ti - bt];
                        const unsigned long long* mrow = sh_mask + (ti - bt) * W;
                        for (int it = 0; it < W; it++) {
                            const unsigned long long m = mrow[it];
                            if (m == 0) continue;
                            const unsigned ma = (unsigned)m, mb = (unsigned)(m >> 32);
                            const int na = __popc(ma);
                            if ((ma >> lane) & 1u) fl.pair[cur + __popc(ma & lt_mask)] = make_int2(ti, t_slot[it * 64 + lane]);
                            if ((mb >> lane) & 1u) fl.pair[cur + na + __popc(mb & lt_mask)] = make_int2(ti, t_slot[it * 64 + 32 + lane]);
                            cur += na + __popc(mb);
                        }
                        continue;
                    }
                    const bool cull = !count && (!RODS || ntiles > 1);      // rods in reach-sized cells need nearly all 27 cells: not worth the test
                    const bool first_visit = pass == 0 && tile == 0;
                    // a particle whose reach touches none of the cells of this tile has nothing to do here (except on the last
                    // tile, where bonded partners are appended by index)
                    if (cull && !first_visit && !(sh_tcm[ti - bt] & sh_tilecells) && (RODS || tile != ntiles - 1)) continue;
                    int target, ttype;
                    float t1x, t1y, t1z;
                    int con0 = -1, con1 = -1, con2 = -1, con3 = -1;
                    if (RODS && ntiles == 1) {             // the target is itself an entry of the staged tile (its cell is the centre one)
                        const float4 q = t_pf[centre_off + (ti - tb)];
                        t1x = q.x; t1y = q.y; t1z = q.z;
                        target = __float_as_int(q.w) & 0xffffff;
                        ttype = __float_as_int(q.w) >> 24;
                    } else {
                        const double4 tpw = s.posw[ti];
                        target = w_orig(tpw.w);
                        ttype = w_type(tpw.w);
                        if (fl.heavy_types && !((fl.heavy_types >> ttype) & 1u)) continue;      // warp-uniform: k_gate_rows_gen has listed this target
                        if (!RODS) {
                            ConList cl;
                            get_conlist(s.mol, w_moltype(tpw.w), target, cl);
                            con0 = cl.con[0]; con1 = cl.con[1]; con2 = cl.con[2]; con3 = cl.con[3];
                        }
                        t1x = (float)(rel_frac(tpw.x + s.shift[0], ccen[0]) * s.box[0]); t1y = (float)(rel_frac(tpw.y + s.shift[1], ccen[1]) * s.box[1]);
                        t1z = (float)(rel_frac(tpw.z + s.shift[2], ccen[2]) * s.box[2]);
                    }
                    const float* reach_row = s.reach2 + ttype * s.ntypes;
                    const float reach_same = s.reach2[s.ntypes * s.ntypes + ttype];      // RODS: the largest reach of this type (conservative for mixed rod types)
                    // neighbour cells whose nearest face is beyond the largest reach of this particle hold no partner: their
                    // 64-candidate groups are skipped. Not while counting work: the reference's sqmaxcut gate is wider than reach.
                    unsigned cellmask = 0xffffffffu;
                    if (cull && !first_visit) cellmask = sh_tcm[ti - bt];
                    else if (cull) {
                        float rmax2 = 0.f;          // the largest reach of this particle towards the types present in neighbour cell `lane`
                        float g2 = 0.f;
                        if (lane < ncell_nb) {
                            unsigned tm = sh_ctypes[lane];
                            if (tm == 0xffffffffu) rmax2 = s.reach2[s.ntypes * s.ntypes + ttype];
                            else while (tm) { const int b = __ffs(tm) - 1; tm &= tm - 1; rmax2 = fmaxf(rmax2, reach_row[b]); }
                            const int ox = nx == 1 ? 0 : lane % nx - 1, oy = ny == 1 ? 0 : (lane / nx) % ny - 1, oz = nz == 1 ? 0 : lane / (nx * ny) - 1;
                            const float gx = ox < 0 ? t1x + hcell[0] : ox > 0 ? hcell[0] - t1x : 0.f;
                            const float gy = oy < 0 ? t1y + hcell[1] : oy > 0 ? hcell[1] - t1y : 0.f;
                            const float gz = oz < 0 ? t1z + hcell[2] : oz > 0 ? hcell[2] - t1z : 0.f;
                            g2 = fmaxf(gx, 0.f) * fmaxf(gx, 0.f) + fmaxf(gy, 0.f) * fmaxf(gy, 0.f) + fmaxf(gz, 0.f) * fmaxf(gz, 0.f);
                        }
                        cellmask = __ballot_sync(0xffffffffu, lane < ncell_nb && g2 <= rmax2 * 1.001f);
                        if (lane == 0) sh_tcm[ti - bt] = cellmask;
                    }
                    unsigned n_cand = 0, n_sure = 0;     // n_sure: pairs surely inside sqmaxcut AND surely beyond reach: gated, energy exactly 0, not listed
                    int cur = pass ? sh_pos[ti - bt] : 0;      // pass 0: partners counted so far in this tile; pass 1: write cursor
                    auto scan = [&](auto write_c, auto count_c, auto cull_c) {        // the edge band is listed exactly when work is being counted, in BOTH passes
                        constexpr bool WRITE = decltype(write_c)::value, BAND = decltype(count_c)::value, COUNT = BAND && !WRITE, CULL = decltype(cull_c)::value;
                        for (int base = 0; base < TC; base += 64) {
                            if (CULL && !(sh_grp[base >> 6] & cellmask)) {       // warp-uniform
                                if (!WRITE && use_masks && lane == 0) sh_mask[(ti - bt) * W + (base >> 6)] = 0ull;
                                continue;
                            }
                            bool pa = false, pb = false;
                            int sa = 0, sb = 0;
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int p = base + 32 * h + lane;        // the tile is padded to a multiple of 64: no bounds test
                                const float4 q = t_pf[p];
                                const int wbits = __float_as_int(q.w);
                                const int orig = wbits & 0xffffff;
                                float dx = t1x - q.x, dy = t1y - q.y, dz = t1z - q.z;
                                if (WRAP) {
                                    dx -= boxf[0] * rintf(dx * ibox[0]); dy -= boxf[1] * rintf(dy * ibox[1]); dz -= boxf[2] * rintf(dz * ibox[2]);
                                }
                                const float d2 = dx * dx + dy * dy + dz * dz;
                                const float reach = RODS ? reach_same : reach_row[wbits >> 24];
                                bool ok = MODE == 2 ? orig < target : orig != target;
                                if (!RODS) ok = ok & !((orig == con0) | (orig == con1) | (orig == con2) | (orig == con3));
                                // listed: may interact (inside reach) or sits on the edge of the sqmaxcut gate (needs the exact FP64 test to be counted)
                                // (the edge band only matters for the work counters: without them, beyond reach means exactly zero energy)
                                const bool pass_gate = BAND ? (ok & ((d2 <= reach) | ((d2 > cut_lo) & (d2 <= cut_hi))))
                                                            : (ok & (d2 <= reach));      // bitwise on purpose: no branches in this loop
                                if (COUNT && ok && p < TC) { n_cand++; if (!pass_gate && d2 <= cut_lo) n_sure++; }
                                if (h == 0) { pa = pass_gate; if (WRITE) sa = t_slot[p]; } else { pb = pass_gate; if (WRITE) sb = t_slot[p]; }
                            }
                            const unsigned ma = __ballot_sync(0xffffffffu, pa), mb = __ballot_sync(0xffffffffu, pb);
                            const int na = __popc(ma);
                            if (WRITE) {
                                if (pa) fl.pair[cur + __popc(ma & lt_mask)] = make_int2(ti, sa);
                                if (pb) fl.pair[cur + na + __popc(mb & lt_mask)] = make_int2(ti, sb);
                            } else if (use_masks && lane == 0) {
                                sh_mask[(ti - bt) * W + (base >> 6)] = (unsigned long long)ma | ((unsigned long long)mb << 32);
                            }
                            cur += na + __popc(mb);
                        }
                    };
                    if (pass && use_masks) {       // replay the masks of the counting pass
                        const unsigned long long* mrow = sh_mask + (ti - bt) * W;
                        for (int it = 0; it < W; it++) {
                            const unsigned long long m = mrow[it];
                            if (m == 0) continue;
                            const unsigned ma = (unsigned)m, mb = (unsigned)(m >> 32);
                            const int na = __popc(ma);
                            if ((ma >> lane) & 1u) fl.pair[cur + __popc(ma & lt_mask)] = make_int2(ti, t_slot[it * 64 + lane]);
                            if ((mb >> lane) & 1u) fl.pair[cur + na + __popc(mb & lt_mask)] = make_int2(ti, t_slot[it * 64 + 32 + lane]);
                            cur += na + __popc(mb);
                        }
                    }
                    else if (pass) {
                        if (count) scan(std::true_type{}, std::true_type{}, std::false_type{});
                        else if (cull) scan(std::true_type{}, std::false_type{}, std::true_type{});
                        else scan(std::true_type{}, std::false_type{}, std::false_type{});
                    }
                    else if (count) scan(std::false_type{}, std::true_type{}, std::false_type{});
                    else if (cull) scan(std::false_type{}, std::false_type{}, std::true_type{});
                    else scan(std::false_type{}, std::false_type{}, std::false_type{});
                    if (!RODS && tile == ntiles - 1) {       // bonded partners by index (never gated, mc/paire.h:1214)
                        int orig = lane == 0 ? con0 : lane == 1 ? con1 : lane == 2 ? con2 : lane == 3 ? con3 : -1;
                        bool on = orig >= 0 && orig != target && (MODE != 2 || orig < target);
                        if (on && !pass) n_cand++;
                        unsigned m = __ballot_sync(0xffffffffu, on);
                        if (pass && on) fl.pair[cur + __popc(m & lt_mask)] = make_int2(ti, s.slot_of[orig]);
                        cur += __popc(m);
                    }
                    __syncwarp();        // every lane has read its cursor (racecheck: make the order of the lane-0 update explicit)
                    if (lane == 0) { if (pass) sh_pos[ti - bt] = cur; else sh_cnt[ti - bt] += cur; }
                    __syncwarp();        // ... and the same warp reads this entry again on the next tile
                    if (!pass && count) {
                        n_cand = __reduce_add_sync(0xffffffffu, n_cand);
                        n_sure = __reduce_add_sync(0xffffffffu, n_sure);
                        if (lane == 0) { atomicAdd(&counters[0], (unsigned long long)n_cand); atomicAdd(&counters[1], (unsigned long long)n_sure); }
                    }
                }
            }
            if (pass == 0) {      // reserve the batch's span, hand every particle its sub-span
                __syncthreads();
                if (wid == 0) {
                    int carry = 0;
                    for (int k0 = 0; k0 < nb; k0 += 32) {
                        const int k = k0 + lane;
                        const int v = k < nb ? sh_cnt[k] : 0;
                        int x = v;
                        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                        if (k < nb) sh_pos[k] = carry + x - v;
                        carry += __shfl_sync(0xffffffffu, x, 31);
                    }
                    int base = 0;
                    if (lane == 0) {
                        base = atomicAdd(fl.total, carry);
                        const bool ok = base + carry <= fl.cap;
                        if (!ok) atomicOr(fl.overflow, 2);
                        sh_ok = ok ? 1 : 0;
                    }
                    base = __shfl_sync(0xffffffffu, base, 0);
                    for (int k = lane; k < nb; k += 32) {
                        const double tw = s.posw[bt + k].w;
                        const int target = w_orig(tw);
                        sh_pos[k] += base;
                        if (fl.heavy_types && !((fl.heavy_types >> w_type(tw)) & 1u)) continue;
                        fl.chunks[target] = make_int4(sh_pos[k], sh_cnt[k], -1, 0);      // the particle's span of the list
                        fl.head[target] = target;
                    }
                }
                __syncthreads();
                if (!sh_ok) return;          // list too short: the host grows it and repeats the launch
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_gate_rows: the gate of the flat pipeline for rods-only systems on grids without wrap (>= 5 cells per axis), THREAD per
// target. k_gate_cells gives a warp to every target and spends ~60 warp instructions per 32 candidate tests on ballots,
// compaction, mask replay and per-cell set-up (cells hold ~17 particles, a block amortises its prologue over very few
// targets). Here a block takes 32 CONSECUTIVE slots of one cell row (cells of a row are contiguous in the sorted arrays, so a
// unit is always full whatever the cell populations are) and stages the union of their neighbourhoods: the cells
// [cxa-1, cxb+1] of the 9 neighbouring rows, <= 18 contiguous slot ranges. Lane t owns target t; the four warps split the
// staged candidates round-robin (warp w takes c = w, w+4, ...), so every lane of a warp reads the SAME candidate (one
// broadcast LDS.128) and runs the FP32 test privately: ~9 instructions per test, no votes. Hits go to a per-(target, warp)
// buffer in shared memory; afterwards the block reserves one span of the global list and every target's hits are written
// contiguously (warp-0 hits first, then warp 1, ...: a fixed order, so the combine step stays bit-reproducible).
// Anything the layout cannot hold (a unit spanning too many cells for the box, a neighbourhood above the tile, a buffer
// overflow in a very dense spot) raises bit 2 of the overflow word: the host then repeats the launch with k_gate_cells.
// ------------------------------------------------------------------------------------------------
constexpr int GR_T = 32;            // targets per unit (= lanes)
constexpr int GR_SL = 4;            // candidate slices (= warps per block)
constexpr int GR_TILE = 1024;       // staged candidates per unit
#ifndef GR_CAP_N
#define GR_CAP_N 39
#endif
constexpr int GR_CAP = GR_CAP_N;    // hits per (target, slice)
constexpr int GR_STRIDE = 34;       // halfwords per buffer row: row k of target t at k * 34 + t -> the write-out (fixed t, k = lane) is conflict-free

template <int MODE>
__global__ void __launch_bounds__(GR_SL * 32, 6)
k_gate_rows(DevSys s, FlatList fl) {
    __shared__ float4 t_pf[GR_TILE];                // x, y, z: FP32 coordinates relative to the unit centre, length units; w = x^2 + y^2 + z^2
    __shared__ __align__(16) int t_orig[MODE == 2 ? GR_TILE : 4];      // allToAll rows compare original indices; every-particle passes only need "not myself"
    __shared__ uint2 t_dq[GR_TILE];                 // axis of every staged rod, 3 x snorm16 (segment lower bound; its margin covers the quantisation)
    __shared__ unsigned short sh_hit[GR_SL][GR_CAP * GR_STRIDE];
    __shared__ int sh_cnt[GR_SL][GR_T];
    __shared__ int sh_off[GR_SL][GR_T];
    __shared__ int sh_sb[20], sh_soff[40];          // staged segments: first slot, offset in the tile ([19..39]: INT_MAX, the slot search reads past 18)
    __shared__ int sh_cx[3];
    __shared__ int sh_ok;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.y;
    const int nx = s.nc[0], ny = s.nc[1];
    const int cy = row % ny, cz = row / ny;
    const int rs = s.cell_start[row * nx], re = s.cell_start[(row + 1) * nx];
    if (threadIdx.x >= 19 && threadIdx.x < 40) sh_soff[threadIdx.x] = 0x7fffffff;
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    // A unit is 32 consecutive slots of the row. Where the row is sparse such a unit would span many cells (and its neighbourhood
    // the whole row): it is then processed in sub-units of at most kmax cells each (dense rows: one sub-unit, all lanes busy).
    const int kmax = min(6, nx - 3);
    for (int ufirst = rs + GR_T * blockIdx.x; ufirst < re; ufirst += GR_T * gridDim.x) {
      const int ulast = min(ufirst + GR_T, re) - 1;
      for (int first = ufirst; first <= ulast;) {
        __syncthreads();                 // previous unit fully written out
        // ---- cells of the first and the last target (warp 0: lanes over the cells of the row)
        if (wid == 0) {
            int cxa = 0, cxb = 0;
            for (int c0 = 0; c0 < nx; c0 += 32) {
                const int cxl = c0 + lane;
                int b = 0, e = 0;
                if (cxl < nx) { b = s.cell_start[row * nx + cxl]; e = s.cell_start[row * nx + cxl + 1]; }
                const unsigned ma = __ballot_sync(0xffffffffu, cxl < nx && b <= first && first < e);
                const unsigned mb = __ballot_sync(0xffffffffu, cxl < nx && b <= ulast && ulast < e);
                if (ma) cxa = c0 + __ffs(ma) - 1;
                if (mb) cxb = c0 + __ffs(mb) - 1;
            }
            int lastv = ulast;
            if (cxb - cxa + 1 > kmax) { cxb = cxa + kmax - 1; lastv = s.cell_start[row * nx + cxb + 1] - 1; }
            // ---- the <= 18 contiguous slot ranges of the neighbourhood: 9 rows x (one range, or two when the x range wraps)
            const int k = cxb - cxa + 1;
            const bool fits = nx >= k + 3;
            int b = 0, len = 0;
            if (fits && lane < 18) {
                const int rr = lane >> 1, part = lane & 1;
                const int yy = (cy + rr % 3 - 1 + ny) % ny, zz = (cz + rr / 3 - 1 + s.nc[2]) % s.nc[2];
                const int rbase = (zz * ny + yy) * nx;
                const int lo = cxa - 1, hi = cxb + 1;                // inclusive cell range, may stick out of [0, nx)
                int a0, a1;                                          // this part's range, empty when a0 > a1
                if (lo < 0) { if (part == 0) { a0 = 0; a1 = hi; } else { a0 = lo + nx; a1 = nx - 1; } }
                else if (hi >= nx) { if (part == 0) { a0 = lo; a1 = nx - 1; } else { a0 = 0; a1 = hi - nx; } }
                else { a0 = part == 0 ? lo : 1; a1 = part == 0 ? hi : 0; }
                if (a0 <= a1) { b = s.cell_start[rbase + a0]; len = s.cell_start[rbase + a1 + 1] - b; }
            }
            int x = len;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane < 19) { sh_sb[lane] = b; sh_soff[lane] = x - len; }
            if (lane == 0) { sh_cx[0] = cxa; sh_cx[1] = cxb; sh_cx[2] = lastv; }
            const int C = __shfl_sync(0xffffffffu, x, 18);
            if (lane == 0) {
                sh_ok = (fits && C <= GR_TILE) ? 1 : 0;
                if (!sh_ok) atomicOr(fl.overflow, 4);
            }
        }
        __syncthreads();
        if (!sh_ok) return;
        const int last = sh_cx[2];
        const int count = last - first + 1;
        const int C = sh_soff[18];
#ifndef GATE_NO_MASKS
        const int Cpad = (C + 63) / 64 * 64;          // the scan below takes the tile in half-words of 64 candidates (16 per warp)
#else
        const int Cpad = (C + 4 * GR_SL - 1) / (4 * GR_SL) * (4 * GR_SL);
#endif
        // unit centre in box fractions (the staged FP32 positions are wrapped into [0, 1): the difference is folded once)
        const float cen[3] = {(float)(0.5 * (sh_cx[0] + sh_cx[1] + 1) / nx - s.shift[0]), (float)((cy + 0.5) / ny - s.shift[1]), (float)((cz + 0.5) / s.nc[2] - s.shift[2])};
        auto staged = [&](const float4& f) {
            float x = f.x - cen[0], y = f.y - cen[1], z = f.z - cen[2];
            x = (x - rintf(x)) * boxf[0]; y = (y - rintf(y)) * boxf[1]; z = (z - rintf(z)) * boxf[2];
            return make_float4(x, y, z, x * x + y * y + z * z);
        };
        auto quant = [](const float4& d) {          // 3 x snorm16
            const int qx = __float2int_rn(d.x * 32767.f), qy = __float2int_rn(d.y * 32767.f), qz = __float2int_rn(d.z * 32767.f);
            return make_uint2((unsigned)(qx & 0xffff) | ((unsigned)qy << 16), (unsigned)(qz & 0xffff));
        };
        // ---- stage: a warp per segment, contiguous 16-byte loads of the FP32 copies
        for (int k = wid; k < 18; k += GR_SL) {
            const int b = sh_sb[k], off = sh_soff[k], len = sh_soff[k + 1] - off;
            for (int idx = lane; idx < len; idx += 32) {
                const float4 f = s.p32[b + idx], d = s.d32[b + idx];
                t_pf[off + idx] = staged(f);
                t_dq[off + idx] = quant(d);
                if (MODE == 2) t_orig[off + idx] = __float_as_int(f.w) & 0xffffff;
            }
        }
        for (int p = C + threadIdx.x; p < Cpad; p += blockDim.x) {          // padding: an infinite |q|^2 fails the comparison
            t_pf[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
            t_dq[p] = make_uint2(0u, 0u);
            if (MODE == 2) t_orig[p] = -1;
        }
        // ---- this lane's target: |t - q|^2 <= reach  <=>  |q|^2 - 2 t.q <= reach - |t|^2 (three FMAs and a compare per test; the
        // rounding error, ~1e-6 relative at these magnitudes, is far inside the 0.1 % margin carried by reach)
        float m2x = 0.f, m2y = 0.f, m2z = 0.f, thr = __int_as_float(0xff800000);
        float tdx = 0.f, tdy = 0.f, tdz = 1.f, hl = 0.f, cut2 = 0.f;
        int target = -2, c_self = -1;
        if (lane < count) {
            const float4 f = s.p32[first + lane], d = s.d32[first + lane];
            const float4 q = staged(f);
            m2x = -2.f * q.x; m2y = -2.f * q.y; m2z = -2.f * q.z;
            tdx = d.x; tdy = d.y; tdz = d.z;
            const int wb = __float_as_int(f.w);
            target = wb & 0xffffff;
            c_self = sh_soff[8] + (first + lane - sh_sb[8]);      // the targets sit in the centre row of their own neighbourhood (segment 8)
            const int T = s.ntypes, tt = wb >> 24;
            thr = s.reach2[T * T + tt] - q.w;     // the largest reach of this type: conservative for mixed rod types
            hl = s.reach2[2 * T * T + T + tt];    // all rods of a system have one length (Topo::genParamPairs enforces it, topo.cpp:22-33)
            for (int b = 0; b < T; b++) cut2 = fmaxf(cut2, s.reach2[T * T + T + tt * T + b]);      // the largest surface cutoff of this type's rod pairs
        }
        __syncthreads();
        // ---- scan: every lane tests its own target against the candidates of this warp's slice. Hits are appended through a
        // per-lane cursor that saturates four entries below the capacity (then the launch is repeated by k_gate_cells).
        unsigned short* buf = sh_hit[wid];
        int cur = lane;
        const int cur_max = lane + (GR_CAP - 4) * GR_STRIDE;
#ifndef GATE_NO_MASKS
        // Pass A: the centre-distance test of 32 candidates of this warp's slice leaves one BIT each in a register word (a broadcast
        // LDS.128, three FMAs, a compare and a predicated OR per test -- no store, no cursor arithmetic, no branch); word j, bit b <->
        // candidate 128 j + 16 (b >> 2) + 4 wid + (b & 3). The bits of a word are then expanded into the per-lane hit buffer
        // (only hits cost instructions there; self-exclusion / the j < i rule of allToAll rows is applied at this point).
        for (int j = 0; 128 * j < Cpad; j++) {
            unsigned m = 0u;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (128 * j + 64 * h < Cpad) {          // warp-uniform
                    const float4* tq = t_pf + 128 * j + 64 * h + 4 * wid;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float4 q = tq[16 * t + u];
                            const float sq = fmaf(m2z, q.z, fmaf(m2y, q.y, fmaf(m2x, q.x, q.w)));
                            m |= sq <= thr ? 1u << (16 * h + 4 * t + u) : 0u;
                        }
                    }
                }
            }
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                const int c = 128 * j + 16 * (b >> 2) + 4 * wid + (b & 3);
                buf[cur] = (unsigned short)c;
                const bool take = MODE == 2 ? t_orig[c] < target : c != c_self;
                cur = min(cur + (take ? GR_STRIDE : 0), cur_max);
            }
        }
#else
        // warp w takes the candidates 16 i + 4 w .. 16 i + 4 w + 3: four float4 reads (and, for allToAll rows, ONE int4 read of the
        // original indices), fetched one trip ahead so that the shared-memory latency hides behind the arithmetic of the current trip
        float4 qn[4];
        int4 on = make_int4(-1, -1, -1, -1);
        {
            const int c0 = 4 * wid;       // Cpad >= 16: always inside the padded tile
#pragma unroll
            for (int u = 0; u < 4; u++) qn[u] = t_pf[c0 + u];
            if (MODE == 2) on = *reinterpret_cast<const int4*>(t_orig + c0);
        }
        for (int c0 = 4 * wid; c0 < Cpad; c0 += 4 * GR_SL) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = qn[u];
            const int ob[4] = {on.x, on.y, on.z, on.w};
            {
                const int cn = min(c0 + 4 * GR_SL, Cpad - 4 * GR_SL + 4 * wid);      // the last trip re-reads its own entries
#pragma unroll
                for (int u = 0; u < 4; u++) qn[u] = t_pf[cn + u];
                if (MODE == 2) on = *reinterpret_cast<const int4*>(t_orig + cn);
            }
            bool hit[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {        // four independent tests first (instruction-level parallelism), appends afterwards
                const float sq = fmaf(m2z, q[u].z, fmaf(m2y, q[u].y, fmaf(m2x, q[u].x, q[u].w)));
                hit[u] = (sq <= thr) & (MODE == 2 ? ob[u] < target : (c0 + u) != c_self);
            }
            if (hit[0] | hit[1] | hit[2] | hit[3]) {
#pragma unroll
                for (int u = 0; u < 4; u++) { buf[cur] = (unsigned short)(c0 + u); cur += hit[u] ? GR_STRIDE : 0; }
                cur = min(cur, cur_max);
            }
        }
#endif
        if (__any_sync(0xffffffffu, cur == cur_max) && lane == 0) atomicOr(fl.overflow, 4);
        // ---- segment lower bound (lb_beyond_one, pair_energy.cuh): of the candidates within the centre-distance reach only those whose
        // rods can come within the surface cutoff interact at all -- a sixth of them in a dense rod fluid. Every lane filters its own
        // hits in place.
        {
            const int nk = (cur - lane) / GR_STRIDE;
            const float tx = -0.5f * m2x, ty = -0.5f * m2y, tz = -0.5f * m2z;
            int wr = lane;
            for (int k = 0, rd = lane; k < nk; k++, rd += GR_STRIDE) {
                const int c = buf[rd];
                const float4 q = t_pf[c];
                const uint2 dq = t_dq[c];
                const float bx = (float)(short)(dq.x & 0xffff) * (1.f / 32767.f), by = (float)(short)(dq.x >> 16) * (1.f / 32767.f), bz = (float)(short)(dq.y & 0xffff) * (1.f / 32767.f);
                const float rx = tx - q.x, ry = ty - q.y, rz = tz - q.z;
                // 7.5e-3: |sin| of the angle between the axes can be that much larger than computed from a 16-bit quantised axis
                if (!lb_beyond_one(rx, ry, rz, rx * rx + ry * ry + rz * rz, tdx, tdy, tdz, bx, by, bz, hl, hl, cut2, 7.5e-3f)) { buf[wr] = (unsigned short)c; wr += GR_STRIDE; }
            }
            cur = wr;
        }
        const int cnt = (cur - lane) / GR_STRIDE;
        sh_cnt[wid][lane] = cnt;
        __syncthreads();
        // ---- reserve the unit's span of the list, hand every (target, slice) its sub-span
        if (wid == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < GR_SL; w++) tot += sh_cnt[w][lane];
            int x = tot;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            const int total = __shfl_sync(0xffffffffu, x, 31);
            int base = 0;
            if (lane == 0) {
                base = atomicAdd(fl.total, total);
                const bool ok = base + total <= fl.cap;
                if (!ok) atomicOr(fl.overflow, 2);
                sh_ok = ok ? 1 : 0;
            }
            base = __shfl_sync(0xffffffffu, base, 0) + x - tot;
            {
                int o = base;
#pragma unroll
                for (int w = 0; w < GR_SL; w++) { sh_off[w][lane] = o; o += sh_cnt[w][lane]; }
            }
            if (lane < count) {
                fl.chunks[target] = make_int4(base, tot, -1, 0);
                fl.head[target] = target;
            }
        }
        __syncthreads();
        if (!sh_ok) return;
        // ---- write-out: after the segment bound a (target, slice) keeps one or two partners, so every lane writes its own target's
        // hits (behind those of the slices before it: a fixed order, the combine step stays bit-reproducible)
        {
            const int my_off = sh_off[wid][lane];
            for (int k = 0, rd = lane; k < cnt; k++, rd += GR_STRIDE) {
                const int c = buf[rd];
                int sg = 0;                    // tile index -> sorted slot: which of the 18 staged segments holds it (sh_soff is padded with INT_MAX)
                sg += c >= sh_soff[sg + 16] ? 16 : 0;
                sg += c >= sh_soff[sg + 8] ? 8 : 0;
                sg += c >= sh_soff[sg + 4] ? 4 : 0;
                sg += c >= sh_soff[sg + 2] ? 2 : 0;
                sg += c >= sh_soff[sg + 1] ? 1 : 0;
                fl.pair[my_off + k] = make_int2(first + lane, sh_sb[sg] + (c - sh_soff[sg]));
            }
        }
        first = last + 1;
      }
    }
}

// ------------------------------------------------------------------------------------------------
// k_gate_rows_gen: the same thread-per-target gate for systems with spheres, mixed types and bonded molecules (the lipid
// membrane: ~80 particles per cell, 2 000 candidates per neighbourhood of which 2 % interact -- k_gate_cells spends 93 % of the
// membrane's full-energy pass there). Differences from k_gate_rows:
//   * the neighbourhood of a unit is processed in tiles of GR_TILE candidates; hits (sorted slots, 32 bit) stay in shared
//     memory across the tiles
//   * the reach depends on the (target type, candidate type) pair: per-lane row of a T x T table in shared memory (T <= 8)
//   * chain neighbours: candidates whose original index is within +-2 of the target's are left out of the distance scan
//     (one subtract and one unsigned compare per test, which also drops the target itself) and decided per target
//     afterwards: a bonded partner (ParticleVector::getConlist, structures/Conf.h:90-147) is always listed -- the cutoff does
//     not apply to it (mc/paire.h:1214) -- any other one by its exact FP64 minimum-image distance.
// ------------------------------------------------------------------------------------------------
constexpr int GG_CAP = 48;          // hits per (target, slice)
constexpr int GG_TILE = 768;        // staged candidates per tile
constexpr int GG_STRIDE = 33;       // words per buffer row
constexpr int GG_MAXT = 8;          // particle types the shared reach table holds

template <int MODE>
__global__ void __launch_bounds__(GR_SL * 32, 4)
k_gate_rows_gen(DevSys s, FlatList fl) {
    __shared__ float4 t_pf[GG_TILE];                // x, y, z relative to the unit centre (length units), w = x^2 + y^2 + z^2
    __shared__ __align__(16) int t_ot[GG_TILE];     // original index | type << 24
    __shared__ __align__(16) int t_slot[GG_TILE];
    __shared__ int sh_hit[GR_SL][GG_CAP * GG_STRIDE];
    __shared__ int sh_cnt[GR_SL][GR_T];
    __shared__ int sh_off[GR_SL][GR_T];
    __shared__ float sh_reach[GG_MAXT * GG_MAXT];
    __shared__ int sh_sb[20], sh_soff[20];
    __shared__ int sh_cx[3];
    __shared__ int sh_ok;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row = blockIdx.y;
    const int nx = s.nc[0], ny = s.nc[1];
    const int cy = row % ny, cz = row / ny;
    const int rs = s.cell_start[row * nx], re = s.cell_start[(row + 1) * nx];
    const int T = s.ntypes;
    for (int k = threadIdx.x; k < T * T; k += blockDim.x) sh_reach[k] = s.reach2[k];
    // A unit is 32 consecutive slots of the row. Where the row is sparse such a unit would span many cells (and its neighbourhood
    // the whole row): it is then processed in sub-units of at most kmax cells each (dense rows: one sub-unit, all lanes busy).
    const int kmax = min(6, nx - 3);
    for (int ufirst = rs + GR_T * blockIdx.x; ufirst < re; ufirst += GR_T * gridDim.x) {
      const int ulast = min(ufirst + GR_T, re) - 1;
      for (int first = ufirst; first <= ulast;) {
        __syncthreads();
        if (wid == 0) {
            int cxa = 0, cxb = 0;
            for (int c0 = 0; c0 < nx; c0 += 32) {
                const int cxl = c0 + lane;
                int b = 0, e = 0;
                if (cxl < nx) { b = s.cell_start[row * nx + cxl]; e = s.cell_start[row * nx + cxl + 1]; }
                const unsigned ma = __ballot_sync(0xffffffffu, cxl < nx && b <= first && first < e);
                const unsigned mb = __ballot_sync(0xffffffffu, cxl < nx && b <= ulast && ulast < e);
                if (ma) cxa = c0 + __ffs(ma) - 1;
                if (mb) cxb = c0 + __ffs(mb) - 1;
            }
            int lastv = ulast;
            if (cxb - cxa + 1 > kmax) { cxb = cxa + kmax - 1; lastv = s.cell_start[row * nx + cxb + 1] - 1; }
            const int k = cxb - cxa + 1;
            const bool fits = nx >= k + 3;
            int b = 0, len = 0;
            if (fits && lane < 18) {
                const int rr = lane >> 1, part = lane & 1;
                const int yy = (cy + rr % 3 - 1 + ny) % ny, zz = (cz + rr / 3 - 1 + s.nc[2]) % s.nc[2];
                const int rbase = (zz * ny + yy) * nx;
                const int lo = cxa - 1, hi = cxb + 1;
                int a0, a1;
                if (lo < 0) { if (part == 0) { a0 = 0; a1 = hi; } else { a0 = lo + nx; a1 = nx - 1; } }
                else if (hi >= nx) { if (part == 0) { a0 = lo; a1 = nx - 1; } else { a0 = 0; a1 = hi - nx; } }
                else { a0 = part == 0 ? lo : 1; a1 = part == 0 ? hi : 0; }
                if (a0 <= a1) { b = s.cell_start[rbase + a0]; len = s.cell_start[rbase + a1 + 1] - b; }
            }
            int x = len;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane < 19) { sh_sb[lane] = b; sh_soff[lane] = x - len; }
            if (lane == 0) { sh_cx[0] = cxa; sh_cx[1] = cxb; sh_cx[2] = lastv; sh_ok = fits ? 1 : 0; if (!fits) atomicOr(fl.overflow, 4); }
        }
        __syncthreads();
        if (!sh_ok) return;
        const int last = sh_cx[2];
        const int count = last - first + 1;
        const int C = sh_soff[18];
        const double ccen[3] = {0.5 * (sh_cx[0] + sh_cx[1] + 1) / nx, (cy + 0.5) / ny, (cz + 0.5) / s.nc[2]};
        auto staged = [&](const double4& pw) {
            const float x = (float)(rel_frac(pw.x + s.shift[0], ccen[0]) * s.box[0]), y = (float)(rel_frac(pw.y + s.shift[1], ccen[1]) * s.box[1]),
                        z = (float)(rel_frac(pw.z + s.shift[2], ccen[2]) * s.box[2]);
            return make_float4(x, y, z, x * x + y * y + z * z);
        };
        // ---- this lane's target
        float m2x = 0.f, m2y = 0.f, m2z = 0.f, tt2 = __int_as_float(0x7f800000);      // idle lanes: threshold -inf, never a hit
        int target = -0x40000000, ttype = 0, tmol = 0;
        double4 tpw = make_double4(0.0, 0.0, 0.0, 0.0);
        if (lane < count) {
            tpw = s.posw[first + lane];
            const float4 q = staged(tpw);
            m2x = -2.f * q.x; m2y = -2.f * q.y; m2z = -2.f * q.z; tt2 = q.w;
            target = w_orig(tpw.w); ttype = w_type(tpw.w); tmol = w_moltype(tpw.w);
        }
        // targets of the heavy types (a rod among lipid beads has hundreds of partners) are left to k_gate_cells, launched next
        const bool mine = lane < count && !((fl.heavy_types >> ttype) & 1u);
        if (!mine) { tt2 = __int_as_float(0x7f800000); target = -0x40000000; }
        const float* reach_row = sh_reach + ttype * T;
        int* buf = sh_hit[wid];
        int cur = lane;
        const int cur_max = lane + (GG_CAP - 4) * GG_STRIDE;
        for (int t0 = 0; t0 < C; t0 += GG_TILE) {
            const int TC = min(GG_TILE, C - t0);
            const int TCpad = (TC + 4 * GR_SL - 1) / (4 * GR_SL) * (4 * GR_SL);
            __syncthreads();            // previous tile consumed
            for (int k = wid; k < 18; k += GR_SL) {
                const int b = sh_sb[k], off = sh_soff[k], len = sh_soff[k + 1] - off;
                const int lo = max(off, t0), hi = min(off + len, t0 + TC);
                for (int p = lo + lane; p < hi; p += 32) {
                    const double4 pw = s.posw[b + (p - off)];
                    t_pf[p - t0] = staged(pw); t_ot[p - t0] = w_orig(pw.w) | (w_type(pw.w) << 24); t_slot[p - t0] = b + (p - off);
                }
            }
            for (int p = TC + threadIdx.x; p < TCpad; p += blockDim.x) {
                t_pf[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
                t_ot[p] = 0x00ffffff; t_slot[p] = 0;
            }
            __syncthreads();
            {
                for (int c0 = 4 * wid; c0 < TCpad; c0 += 4 * GR_SL) {
                    const int4 ot4 = *reinterpret_cast<const int4*>(t_ot + c0);
                    const int ot[4] = {ot4.x, ot4.y, ot4.z, ot4.w};
                    bool hit[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float4 q = t_pf[c0 + u];
                        const float sq = fmaf(m2z, q.z, fmaf(m2y, q.y, fmaf(m2x, q.x, q.w)));
                        const int ob = ot[u] & 0xffffff;
                        const float thr = reach_row[ot[u] >> 24] - tt2;
                        // the target itself and its chain neighbours (original index within +-2) are decided after the scan
                        hit[u] = (sq <= thr) & ((unsigned)(ob - target + 2) > 4u) & (MODE == 2 ? ob < target : true);
                    }
                    if (hit[0] | hit[1] | hit[2] | hit[3]) {
                        const int4 sl4 = *reinterpret_cast<const int4*>(t_slot + c0);
                        const int sl[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) { buf[cur] = sl[u]; cur += hit[u] ? GG_STRIDE : 0; }
                        cur = min(cur, cur_max);
                    }
                }
            }
        }
        // ---- chain neighbours and other particles whose original index is within +-2 (slice 0 only)
        if (wid == 0 && mine) {
            ConList cl;
            get_conlist(s.mol, tmol, target, cl);
            const v3 tp = mk(tpw.x, tpw.y, tpw.z);
#pragma unroll
            for (int dd = 0; dd < 4; dd++) {
                const int partner = target + (dd == 0 ? -2 : dd == 1 ? -1 : dd == 2 ? 1 : 2);
                if (partner < 0 || partner >= s.n) continue;
                if (MODE == 2 && partner > target) continue;
                const int ps = s.slot_of[partner];
                bool list = (partner == cl.con[0]) | (partner == cl.con[1]) | (partner == cl.con[2]) | (partner == cl.con[3]);
                if (!list) {
                    const double4 pw = s.posw[ps];
                    const v3 r = image(s.box, tp, mk(pw.x, pw.y, pw.z));
                    list = dot(r, r) <= (double)reach_row[w_type(pw.w)];
                }
                if (list) { buf[cur] = ps; cur = min(cur + GG_STRIDE, cur_max); }
            }
        }
        if (cur == cur_max) { atomicOr(fl.overflow, 4); atomicOr(fl.heavy, 1 << ttype); }       // rare: the host repeats with this type split off
        const int cnt = (cur - lane) / GG_STRIDE;
        sh_cnt[wid][lane] = cnt;
        __syncthreads();
        if (wid == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < GR_SL; w++) tot += sh_cnt[w][lane];
            int x = tot;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            const int total = __shfl_sync(0xffffffffu, x, 31);
            int base = 0;
            if (lane == 0) {
                base = atomicAdd(fl.total, total);
                const bool ok = base + total <= fl.cap;
                if (!ok) atomicOr(fl.overflow, 2);
                sh_ok = ok ? 1 : 0;
            }
            base = __shfl_sync(0xffffffffu, base, 0) + x - tot;
            {
                int o = base;
#pragma unroll
                for (int w = 0; w < GR_SL; w++) { sh_off[w][lane] = o; o += sh_cnt[w][lane]; }
            }
            if (mine) {
                fl.chunks[target] = make_int4(base, tot, -1, 0);
                fl.head[target] = target;
            }
        }
        __syncthreads();
        if (!sh_ok) return;
        const int my_off = sh_off[wid][lane];
        const bool big = __any_sync(0xffffffffu, cnt > 32);
        const int* my_row = buf + lane * GG_STRIDE;
#pragma unroll
        for (int t = 0; t < GR_T; t++) {
            if (t >= count) break;
            const int nh = __shfl_sync(0xffffffffu, cnt, t), off = __shfl_sync(0xffffffffu, my_off, t);
            if (lane < nh) fl.pair[off + lane] = make_int2(first + t, my_row[t]);
            if (big && lane + 32 < nh) fl.pair[off + lane + 32] = make_int2(first + t, my_row[32 * GG_STRIDE + t]);
        }
        first = last + 1;
      }
    }
}

// ------------------------------------------------------------------------------------------------
// k_gate_fine: the thread-per-target gate on the SUB-CELL level (build_cells_impl: cells cut into sub^3 sub-cells, particles sorted by
// (cell, sub-cell, index)) for systems whose cell edge is set by a few long particles while most have short interactions -- the
// lipid membrane: 265 041 beads with reaches of 1.1 - 2.7 in cells 9.9 wide, k_gate_rows_gen tested ~7 000 candidates per bead for
// ~40 partners. A unit is 32 consecutive slots of ONE cell; its neighbourhood is the bounding box of its sub-cells grown by one
// sub-cell: rows of sub-cells, each row up to three contiguous slot ranges (one per cell it crosses). Only pairs of LIGHT types
// (reach <= sub-cell edge, chosen by the host) are found here; a light target finds its few HEAVY partners in the sorted list of
// heavy particles (binary search for the 27 surrounding cells), heavy targets are listed by k_gate_cells as before.
// ------------------------------------------------------------------------------------------------
constexpr int GF_SEG = 112;         // staged slot ranges per unit: <= 36 rows x 3 runs (sub <= 4)

template <int MODE>
__global__ void __launch_bounds__(GR_SL * 32, 4)
k_gate_fine(DevSys s, FlatList fl, const int* __restrict__ heavy_list, const int* __restrict__ heavy_count, unsigned heavy_mask, unsigned skip_mask) {
    // heavy_mask: types whose pairs are not looked for among the sub-cells (their particles are in heavy_list); skip_mask: types whose
    // TARGETS other gates list (the heavy ones and any type whose targets overflowed a buffer of this kernel)
    __shared__ float4 t_pf[GG_TILE];                // x, y, z relative to the cell centre (length units), w = x^2 + y^2 + z^2
    __shared__ __align__(16) int t_ot[GG_TILE];     // original index | type << 24
    __shared__ __align__(16) int t_slot[GG_TILE];
    __shared__ int sh_hit[GR_SL][GG_CAP * GG_STRIDE];
    __shared__ int sh_cnt[GR_SL][GR_T];
    __shared__ int sh_off[GR_SL][GR_T];
    __shared__ float sh_reach[GG_MAXT * GG_MAXT];
    __shared__ int sh_sb[GF_SEG + 16], sh_soff[GF_SEG + 16];
    __shared__ int sh_hr[27][2];                    // per surrounding cell: its run in the heavy list
    __shared__ int sh_ok;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c0 = blockIdx.y;
    const int nx = s.nc[0], ny = s.nc[1], nz = s.nc[2];
    const int S = s.sub, S3 = S * S * S;
    const int cx = c0 % nx, cy = (c0 / nx) % ny, cz = c0 / (nx * ny);
    const int cs = s.cell_start[c0], ce = s.cell_start[c0 + 1];
    if (cs == ce) return;
    const int T = s.ntypes;
    // reach table of the scan: pairs with a heavy partner are not looked for among the sub-cells
    for (int k = threadIdx.x; k < T * T; k += blockDim.x) sh_reach[k] = ((heavy_mask >> (k % T)) & 1u) ? -1.f : s.reach2[k];
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    const float cen[3] = {(float)((cx + 0.5) / nx - s.shift[0]), (float)((cy + 0.5) / ny - s.shift[1]), (float)((cz + 0.5) / nz - s.shift[2])};
    auto staged = [&](const float4& f) {
        float x = f.x - cen[0], y = f.y - cen[1], z = f.z - cen[2];
        x = (x - rintf(x)) * boxf[0]; y = (y - rintf(y)) * boxf[1]; z = (z - rintf(z)) * boxf[2];
        return make_float4(x, y, z, x * x + y * y + z * z);
    };
    if (heavy_mask && wid == 1) {       // the heavy particles of the 27 surrounding cells: runs of the sorted heavy list
        const int nh = min(heavy_count[0], HV_MAX);
        int lo = 0, hi = 0;
        if (lane < 27) {
            const int ccx = (cx + lane % 3 - 1 + nx) % nx, ccy = (cy + (lane / 3) % 3 - 1 + ny) % ny, ccz = (cz + lane / 9 - 1 + nz) % nz;
            const int cc = (ccz * ny + ccy) * nx + ccx;
            const int b = s.cell_start[cc], e = s.cell_start[cc + 1];
            int a0 = 0, a1 = nh;      // first entry >= b
            while (a0 < a1) { const int m = (a0 + a1) >> 1; if (heavy_list[m] < b) a0 = m + 1; else a1 = m; }
            lo = a0; a1 = nh;         // first entry >= e
            while (a0 < a1) { const int m = (a0 + a1) >> 1; if (heavy_list[m] < e) a0 = m + 1; else a1 = m; }
            hi = a0;
            sh_hr[lane][0] = lo; sh_hr[lane][1] = hi;
        }
    }
    for (int ufirst = cs + GR_T * blockIdx.x; ufirst < ce; ufirst += GR_T * gridDim.x) {
        const int first = ufirst, last = min(ufirst + GR_T, ce) - 1;
        const int count = last - first + 1;
        __syncthreads();
        if (wid == 0) {
            // ---- sub-cells of the first and the last target, bounding box of everything between them
            const int fb = c0 * S3;
            int sa = 0, sb = 0;
            for (int f0 = 0; f0 < S3; f0 += 32) {
                const int f = f0 + lane;
                int b = 0, e = 0;
                if (f < S3) { b = s.fine_start[fb + f]; e = s.fine_start[fb + f + 1]; }
                const unsigned ma = __ballot_sync(0xffffffffu, f < S3 && b <= first && first < e);
                const unsigned mb = __ballot_sync(0xffffffffu, f < S3 && b <= last && last < e);
                if (ma) sa = f0 + __ffs(ma) - 1;
                if (mb) sb = f0 + __ffs(mb) - 1;
            }
            int x0 = S, x1 = -1, y0 = S, y1 = -1, z0 = S, z1 = -1;
            for (int f = sa + lane; f <= sb; f += 32) {
                const int sx = f % S, sy = (f / S) % S, sz = f / (S * S);
                x0 = min(x0, sx); x1 = max(x1, sx); y0 = min(y0, sy); y1 = max(y1, sy); z0 = min(z0, sz); z1 = max(z1, sz);
            }
            x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
            y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
            z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
            const int gx0 = cx * S + x0 - 1, gx1 = cx * S + x1 + 1, gy0 = cy * S + y0 - 1, gz0 = cz * S + z0 - 1;
            const int wy = y1 - y0 + 3, wz = z1 - z0 + 3;
            const int nq = wy * wz * 3;                              // (row, run) pairs: a row crosses at most three cells
            const int cxa = gx0 >= 0 ? gx0 / S : -1;                  // the (unwrapped) cell the row starts in
            int carry = 0;
            for (int q0 = 0; q0 < nq; q0 += 32) {
                const int q = q0 + lane;
                int b = 0, len = 0;
                if (q < nq) {
                    const int r = q / 3, u = q % 3;
                    int gy = gy0 + r % wy, gz = gz0 + r / wy;
                    gy = (gy + ny * S) % (ny * S); gz = (gz + nz * S) % (nz * S);
                    const int ccy = gy / S, sy = gy % S, ccz = gz / S, sz = gz % S;
                    const int cr = cxa + u;                           // unwrapped cell of this run
                    const int lo = max(gx0, cr * S), hi = min(gx1, cr * S + S - 1);
                    if (lo <= hi) {
                        const int ccx = (cr + nx) % nx;
                        const int base = ((ccz * ny + ccy) * nx + ccx) * S3 + (sz * S + sy) * S;
                        b = s.fine_start[base + (lo - cr * S)];
                        len = s.fine_start[base + (hi - cr * S) + 1] - b;
                    }
                }
                int x = len;
                for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (q < nq) { sh_sb[q] = b; sh_soff[q] = carry + x - len; }
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) { sh_soff[nq] = carry; sh_sb[GF_SEG + 8] = nq; sh_ok = nq <= GF_SEG ? 1 : 0; if (!sh_ok) atomicOr(fl.overflow, 4); }
        }
        __syncthreads();
        if (!sh_ok) return;
        const int nq = sh_sb[GF_SEG + 8];
        const int C = sh_soff[nq];
        // ---- this lane's target
        float m2x = 0.f, m2y = 0.f, m2z = 0.f, tt2 = __int_as_float(0x7f800000);      // idle lanes: threshold -inf, never a hit
        int target = -0x40000000, ttype = 0, tmol = 0;
        double4 tpw = make_double4(0.0, 0.0, 0.0, 0.0);
        if (lane < count) {
            tpw = s.posw[first + lane];
            const float4 q = staged(s.p32[first + lane]);
            m2x = -2.f * q.x; m2y = -2.f * q.y; m2z = -2.f * q.z; tt2 = q.w;
            target = w_orig(tpw.w); ttype = w_type(tpw.w); tmol = w_moltype(tpw.w);
        }
        // targets of the heavy types are left to k_gate_heavy / k_gate_cells, launched next
        const bool mine = lane < count && !((skip_mask >> ttype) & 1u);
        if (!mine) { tt2 = __int_as_float(0x7f800000); target = -0x40000000; }
        const float* reach_row = sh_reach + ttype * T;
        int* buf = sh_hit[wid];
        int cur = lane;
        const int cur_max = lane + (GG_CAP - 4) * GG_STRIDE;
        for (int t0 = 0; t0 < C; t0 += GG_TILE) {
            const int TC = min(GG_TILE, C - t0);
            const int TCpad = (TC + 4 * GR_SL - 1) / (4 * GR_SL) * (4 * GR_SL);
            __syncthreads();            // previous tile consumed
            for (int k = wid; k < nq; k += GR_SL) {
                const int b = sh_sb[k], off = sh_soff[k], len = sh_soff[k + 1] - off;
                const int lo = max(off, t0), hi = min(off + len, t0 + TC);
                for (int p = lo + lane; p < hi; p += 32) {
                    const float4 f = s.p32[b + (p - off)];
                    t_pf[p - t0] = staged(f); t_ot[p - t0] = __float_as_int(f.w); t_slot[p - t0] = b + (p - off);
                }
            }
            for (int p = TC + threadIdx.x; p < TCpad; p += blockDim.x) {
                t_pf[p] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
                t_ot[p] = 0x00ffffff; t_slot[p] = 0;
            }
            __syncthreads();
            for (int c4 = 4 * wid; c4 < TCpad; c4 += 4 * GR_SL) {
                const int4 ot4 = *reinterpret_cast<const int4*>(t_ot + c4);
                const int ot[4] = {ot4.x, ot4.y, ot4.z, ot4.w};
                bool hit[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float4 q = t_pf[c4 + u];
                    const float sq = fmaf(m2z, q.z, fmaf(m2y, q.y, fmaf(m2x, q.x, q.w)));
                    const int ob = ot[u] & 0xffffff;
                    const float thr = reach_row[ot[u] >> 24] - tt2;
                    // the target itself and its chain neighbours (original index within +-2) are decided after the scan
                    hit[u] = (sq <= thr) & ((unsigned)(ob - target + 2) > 4u) & (MODE == 2 ? ob < target : true);
                }
                if (hit[0] | hit[1] | hit[2] | hit[3]) {
                    const int4 sl4 = *reinterpret_cast<const int4*>(t_slot + c4);
                    const int sl[4] = {sl4.x, sl4.y, sl4.z, sl4.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) { buf[cur] = sl[u]; cur += hit[u] ? GG_STRIDE : 0; }
                    cur = min(cur, cur_max);
                }
            }
        }
        // ---- chain neighbours and other particles whose original index is within +-2 (slice 0 only): exact, whatever their type
        if (wid == 0 && mine) {
            ConList cl;
            get_conlist(s.mol, tmol, target, cl);
            const v3 tp = mk(tpw.x, tpw.y, tpw.z);
#pragma unroll
            for (int dd = 0; dd < 4; dd++) {
                const int partner = target + (dd == 0 ? -2 : dd == 1 ? -1 : dd == 2 ? 1 : 2);
                if (partner < 0 || partner >= s.n) continue;
                if (MODE == 2 && partner > target) continue;
                const int ps = s.slot_of[partner];
                bool list = (partner == cl.con[0]) | (partner == cl.con[1]) | (partner == cl.con[2]) | (partner == cl.con[3]);
                if (!list) {
                    const double4 pw = s.posw[ps];
                    const v3 r = image(s.box, tp, mk(pw.x, pw.y, pw.z));
                    list = dot(r, r) <= (double)s.reach2[ttype * T + w_type(pw.w)];
                }
                if (list) { buf[cur] = ps; cur = min(cur + GG_STRIDE, cur_max); }
            }
        }
        // ---- heavy partners of a light target (slice 1): the heavy particles of the 27 surrounding cells, exact distance
        if (wid == 1 && mine && heavy_mask) {
            const v3 tp = mk(tpw.x, tpw.y, tpw.z);
            for (int k = 0; k < 27; k++) {
                for (int h = sh_hr[k][0]; h < sh_hr[k][1]; h++) {
                    const int hs = heavy_list[h];
                    const double4 pw = s.posw[hs];
                    const int ob = w_orig(pw.w);
                    if ((unsigned)(ob - target + 2) <= 4u) continue;                 // decided above
                    if (MODE == 2 && ob > target) continue;
                    const v3 r = image(s.box, tp, mk(pw.x, pw.y, pw.z));
                    if (dot(r, r) <= (double)s.reach2[ttype * T + w_type(pw.w)]) { buf[cur] = hs; cur = min(cur + GG_STRIDE, cur_max); }
                }
            }
        }
        if (cur == cur_max) { atomicOr(fl.overflow, 4); atomicOr(fl.heavy, 1 << ttype); }       // rare: the host repeats with this type split off
        const int cnt = (cur - lane) / GG_STRIDE;
        sh_cnt[wid][lane] = cnt;
        __syncthreads();
        if (wid == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < GR_SL; w++) tot += sh_cnt[w][lane];
            int x = tot;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            const int total = __shfl_sync(0xffffffffu, x, 31);
            int base = 0;
            if (lane == 0) {
                base = atomicAdd(fl.total, total);
                const bool ok = base + total <= fl.cap;
                if (!ok) atomicOr(fl.overflow, 2);
                sh_ok = ok ? 1 : 0;
            }
            base = __shfl_sync(0xffffffffu, base, 0) + x - tot;
            {
                int o = base;
#pragma unroll
                for (int w = 0; w < GR_SL; w++) { sh_off[w][lane] = o; o += sh_cnt[w][lane]; }
            }
            if (mine) {
                fl.chunks[target] = make_int4(base, tot, -1, 0);
                fl.head[target] = target;
            }
        }
        __syncthreads();
        if (!sh_ok) return;
        {
            const int my_off = sh_off[wid][lane];
            for (int k = 0, rd = lane; k < cnt; k++, rd += GG_STRIDE) fl.pair[my_off + k] = make_int2(first + lane, buf[rd]);
        }
    }
}

// k_gate_heavy: the partner list of ONE heavy-type particle per block (sub-cell mode; the heavy particles are the few long rods of a
// system of short-range beads). All 256 threads walk the 27 cells around it -- thousands of candidates, read as contiguous 16-byte
// FP32 records -- each warp over its own contiguous share, hits compacted by ballot in candidate order; the shares are counted
// first, the block reserves the span, a second walk writes it: the list is in candidate order, whatever the timing.
template <int MODE>
__global__ void __launch_bounds__(256)
k_gate_heavy(DevSys s, FlatList fl, const int* __restrict__ heavy_list, int nheavy) {
    __shared__ int sh_b[32], sh_off[64];
    __shared__ int sh_wcnt[8], sh_woff[8];
    __shared__ int sh_extra[4];
    __shared__ int sh_nextra, sh_base, sh_ok;
    if ((int)blockIdx.x >= nheavy) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int hs = heavy_list[blockIdx.x];
    const double4 tpw = s.posw[hs];
    const int target = w_orig(tpw.w), ttype = w_type(tpw.w), tmol = w_moltype(tpw.w);
    const int nx = s.nc[0], ny = s.nc[1], nz = s.nc[2], T = s.ntypes;
    const int cx = cell_coord(tpw.x + s.shift[0], nx), cy = cell_coord(tpw.y + s.shift[1], ny), cz = cell_coord(tpw.z + s.shift[2], nz);
    if (threadIdx.x >= 28 && threadIdx.x < 64) sh_off[threadIdx.x] = 0x7fffffff;
    if (wid == 0) {
        int b = 0, len = 0;
        if (lane < 27) {
            const int ccx = (cx + lane % 3 - 1 + nx) % nx, ccy = (cy + (lane / 3) % 3 - 1 + ny) % ny, ccz = (cz + lane / 9 - 1 + nz) % nz;
            const int cc = (ccz * ny + ccy) * nx + ccx;
            b = s.cell_start[cc];
            len = s.cell_start[cc + 1] - b;
        }
        int x = len;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < 28) { sh_b[lane] = b; sh_off[lane] = x - len; }       // [27] = total
    }
    if (threadIdx.x == 32) {        // chain neighbours / original index within +-2: decided exactly, wherever they are
        ConList cl;
        get_conlist(s.mol, tmol, target, cl);
        const v3 tp = mk(tpw.x, tpw.y, tpw.z);
        int ne = 0;
        for (int dd = 0; dd < 4; dd++) {
            const int partner = target + (dd == 0 ? -2 : dd == 1 ? -1 : dd == 2 ? 1 : 2);
            if (partner < 0 || partner >= s.n) continue;
            if (MODE == 2 && partner > target) continue;
            const int ps = s.slot_of[partner];
            bool list = (partner == cl.con[0]) | (partner == cl.con[1]) | (partner == cl.con[2]) | (partner == cl.con[3]);
            if (!list) {
                const double4 pw = s.posw[ps];
                const v3 r = image(s.box, tp, mk(pw.x, pw.y, pw.z));
                list = dot(r, r) <= (double)s.reach2[ttype * T + w_type(pw.w)];
            }
            if (list) sh_extra[ne++] = ps;
        }
        sh_nextra = ne;
    }
    __syncthreads();
    const int C = sh_off[27];
    const int per_warp = ((C + 7) / 8 + 31) & ~31;
    const int wlo = wid * per_warp, whi = min(C, wlo + per_warp);
    const float4 tf = s.p32[hs];
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    const float* reach_row = s.reach2 + ttype * T;
    for (int pass = 0; pass < 2; pass++) {
        int cnt = 0;
        const int out0 = pass ? sh_base + sh_nextra + sh_woff[wid] : 0;
        for (int p0 = wlo; p0 < whi; p0 += 32) {
            const int p = p0 + lane;
            bool hit = false;
            int slot = 0;
            if (p < whi) {
                int sg = 0;
                sg += p >= sh_off[sg + 16] ? 16 : 0;
                sg += p >= sh_off[sg + 8] ? 8 : 0;
                sg += p >= sh_off[sg + 4] ? 4 : 0;
                sg += p >= sh_off[sg + 2] ? 2 : 0;
                sg += p >= sh_off[sg + 1] ? 1 : 0;
                slot = sh_b[sg] + (p - sh_off[sg]);
                const float4 f = s.p32[slot];
                const int wb = __float_as_int(f.w), ob = wb & 0xffffff;
                float dx = tf.x - f.x, dy = tf.y - f.y, dz = tf.z - f.z;
                dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                hit = (dx * dx + dy * dy + dz * dz <= reach_row[wb >> 24]) & ((unsigned)(ob - target + 2) > 4u) & (MODE == 2 ? ob < target : true);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (pass && hit) fl.pair[out0 + cnt + __popc(m & lt_mask)] = make_int2(hs, slot);
            cnt += __popc(m);
        }
        if (pass == 0) {
            if (lane == 0) sh_wcnt[wid] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
                int tot = sh_nextra;
                for (int w = 0; w < 8; w++) { sh_woff[w] = tot - sh_nextra; tot += sh_wcnt[w]; }
                const int base = atomicAdd(fl.total, tot);
                const bool ok = base + tot <= fl.cap;
                if (!ok) atomicOr(fl.overflow, 2);
                sh_ok = ok ? 1 : 0;
                sh_base = base;
                if (ok) {
                    for (int k = 0; k < sh_nextra; k++) fl.pair[base + k] = make_int2(hs, sh_extra[k]);
                    fl.chunks[target] = make_int4(base, tot, -1, 0);
                    fl.head[target] = target;
                }
            }
            __syncthreads();
            if (!sh_ok) return;
        }
    }
}

#ifndef CHEAP_MINB
#define CHEAP_MINB 3
#endif
constexpr int CHEAP_BUF = 256;    // patch-list entries buffered per warp between global appends
// ONE: every particle present has the same type, so there is a single interaction-table entry: it travels as a kernel parameter
// (constant bank, read by LDC on the uniform path) instead of being fetched field by field through the load/store unit
template <bool RODS, bool ONE, bool MIRROR>
__global__ void __launch_bounds__(256, RODS ? CHEAP_MINB : 2)
k_cheap_flat(DevSys s, FlatList fl, unsigned long long* counters, const __grid_constant__ scgpu_iaparam ia1) {
    if (*fl.overflow) return;         // the gate could not finish its list (the host repeats the launch): spans may be missing or out of range
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int total = *fl.total;
    if (total > fl.cap) total = fl.cap;
    unsigned n_gate = 0;
    // Each block owns one contiguous slice of the list. Pairs that owe a patch term are collected in a per-warp shared-memory
    // buffer and appended to the global patch list with ONE atomicAdd per buffer-full (normally once per warp and launch): an
    // append per warp and trip meant ~130 000 atomics on a single address per launch, which serialise in L2 and cost a third
    // of the kernel. No block barrier anywhere: warps whose pairs take the long path do not hold the others up.
    __shared__ int sh_pl[256 / 32][CHEAP_BUF];
    int* wbuf = sh_pl[threadIdx.x >> 5];
    int wn = 0;                       // entries in this warp's buffer (warp-uniform)
    const int per_block = (((total + (int)gridDim.x - 1) / (int)gridDim.x) + 255) & ~255;
    const int lo = blockIdx.x * per_block, hi = min(total, lo + per_block);
    auto flush = [&]() {              // warp-wide
        int gbase = 0;
        if (lane == 0) gbase = atomicAdd(fl.ptotal, wn);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        for (int k = lane; k < wn; k += 32) fl.plist[gbase + k] = wbuf[k];      // plist has the capacity of pair[]: cannot overflow
        __syncwarp();
        wn = 0;
    };
    int p = lo + threadIdx.x;
    int2 pr_next = p < hi ? fl.pair[p] : make_int2(0, 0);
    for (int p0 = lo; p0 < hi; p0 += 256, p += 256) {      // block-uniform trip count
        const int2 pr = pr_next;
        // everything the rod path reads from memory, issued together: one L2 round trip, not three
        // (staging these operands through shared memory with cp.async one trip ahead was measured: 25 % slower -- eight 16-byte
        // copies per thread throttle the memory pipe harder than the exposed latency of four 32-byte loads costs)
        const double4 pi = ldg256(s.posw + pr.x), pj = ldg256(s.posw + pr.y);
        v3 di, dj;
        if (RODS) {      // dir + one more double of the record: a single 32-byte request per particle
            const double4 ri = ldg256(s.rec + (size_t)pr.x * REC + R_DIR), rj = ldg256(s.rec + (size_t)pr.y * REC + R_DIR);
            di = mk(ri.x, ri.y, ri.z); dj = mk(rj.x, rj.y, rj.z);
        }
        pr_next = (p + 256 < hi) ? fl.pair[p + 256] : make_int2(0, 0);
        if (wn > CHEAP_BUF - 32) flush();
        bool np = false;
        if (p < hi) {
            v3 r_cm = image(s.box, mk(pi.x, pi.y, pi.z), mk(pj.x, pj.y, pj.z));
            double dotrcm = dot(r_cm, r_cm);
            int oi = w_orig(pi.w), oj = w_orig(pj.w);
            ConList cl;
            cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0;
            bool bonded = false;
            if (!RODS) {
                get_conlist(s.mol, w_moltype(pi.w), oi, cl);
                bonded = !cl.is_empty && (oj == cl.con[0] || oj == cl.con[1] || oj == cl.con[2] || oj == cl.con[3]);
            }
            double e = 0.0;
            if (dotrcm <= s.sqmaxcut || bonded) {           // the exact PairE gate (mc/paire.h:1214)
                if (RODS) e = pair_energy_cheap_rods(ONE ? ia1 : s.ia[w_type(pi.w) * s.ntypes + w_type(pj.w)], r_cm, dotrcm, di, dj, np);
                else e = pair_energy_cheap<false>(s.box, s.ia, s.ntypes, s.mol, r_cm, dotrcm, s.rec + (size_t)pr.x * REC, w_type(pi.w), w_moltype(pi.w),
                                                  s.rec + (size_t)pr.y * REC, w_type(pj.w), oj, cl, np);
                n_gate++;
            }
            // every-particle passes list a patch pair from both sides: the patch term is evaluated once, by the side whose first particle
            // has the larger original index; k_patch_flat writes a non-zero result into the mirror entry as well
            if (MIRROR && np && oi < oj) np = false;
            fl.e[p] = make_double2(e, 0.0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, np);
        if (np) wbuf[wn + __popc(m & lt_mask)] = p;
        wn += __popc(m);
        __syncwarp();
    }
    if (wn) flush();
    if (counters) {
        n_gate = __reduce_add_sync(0xffffffffu, n_gate);
        if (lane == 0 && n_gate) atomicAdd(&counters[1], (unsigned long long)n_gate);
    }
}

// Patch terms in three dense phases: most pairs drop out after the FIRST patch_intersect() (the partner is simply not inside the
// patch wedge), so intersect #1, intersect #2 and atr_e() run as separate phases with a compaction of the survivors in between,
// which keeps the lanes of the expensive later phases full instead of leaving a few lanes per warp on the long path.
// The compaction is per WARP: every warp owns two ring buffers in shared memory (survivors of phase 1, survivors of phase 2) and
// runs a later phase as soon as 32 entries wait for it -- no block barrier anywhere (the block-wide version of this kernel spent
// 45 % of its stall cycles at its four __syncthreads per batch), warps drift apart and hide each other's gather latency.
#ifndef PF_THREADS_N
#define PF_THREADS_N 128
#endif
constexpr int PF_THREADS = PF_THREADS_N;
constexpr int PF_Q = 64;              // ring capacity per warp: at most 31 waiting + 32 new entries

// the four vectors of one patch out of a cell-sorted record (layout: dir | pd0 | s0 | s1 | pd1 | s2 | s3 | ch0 | ch1 | pos) in
// three or four 32-byte requests instead of a dozen narrower ones: the gathers of this kernel are bound by L1 requests
__device__ __forceinline__ void load_patch_args(const double* rec, int pn, bool chiral, PatchArgs& P) {
    if (!pn) {
        const double4 a = ldg256(rec), b = ldg256(rec + 4), c = ldg256(rec + 8);
        P.dir = mk(a.x, a.y, a.z); P.pdir = mk(a.w, b.x, b.y); P.s0 = mk(b.z, b.w, c.x); P.s1 = mk(c.y, c.z, c.w);
        if (chiral) { const double4 d = ldg256(rec + 20); P.dir = mk(d.y, d.z, d.w); }                       // R_CH0 = 21
    } else {
        const double4 a = ldg256(rec + 12), b = ldg256(rec + 16), c = ldg256(rec + 20);
        P.pdir = mk(a.x, a.y, a.z); P.s0 = mk(a.w, b.x, b.y); P.s1 = mk(b.z, b.w, c.x);                       // R_PD1 = 12, R_S2 = 15, R_S3 = 18
        if (chiral) { const double4 d = ldg256(rec + 24); P.dir = mk(d.x, d.y, d.z); }                       // R_CH1 = 24
        else { const double4 d = ldg256(rec); P.dir = mk(d.x, d.y, d.z); }
    }
}

#ifndef PATCH_MINB
#define PATCH_MINB 4
#endif
// an item carries the slots and the separation vector, so that the later phases do not re-load the pair, both positions and
// re-image them
struct PfItem1 { int p, si, sj, tb; double rx, ry, rz, T1, T2; };
struct PfItem2 { int p, si, sj, tb; double rx, ry, rz, T1, T2, S1, S2; };

template <bool ONE>
__global__ void __launch_bounds__(PF_THREADS, PATCH_MINB)
k_patch_flat(DevSys s, FlatList fl, int any_two_patch, int mirror, const __grid_constant__ scgpu_iaparam ia1) {
    const scgpu_iaparam* ia_one = ONE ? &ia1 : nullptr;
    if (*fl.overflow) return;
    __shared__ PfItem1 sh_q1[PF_THREADS / 32][PF_Q];
    __shared__ PfItem2 sh_q2[PF_THREADS / 32][PF_Q];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    PfItem1* q1 = sh_q1[wid];
    PfItem2* q2 = sh_q2[wid];
    const int total = *fl.ptotal;
    const int ncombo = any_two_patch ? 4 : 1;
    const int gw = blockIdx.x * (PF_THREADS / 32) + wid, nwarps = gridDim.x * (PF_THREADS / 32);
    // The patch combinations of two-patch types run one after another, each drained completely before the next starts: a pair is
    // handled by the same warp in every combination (the batch -> warp map is fixed), so its slot of fl.e has one writer and the
    // additions happen in program order -- bit-reproducible.
    for (int combo = 0; combo < ncombo; combo++) {
        const int pn1 = combo & 1, pn2 = combo >> 1;
        int h1 = 0, n1 = 0, h2 = 0, n2 = 0;              // ring heads and fill counts (warp-uniform)
        // ---- phase 3: attraction of the pairs both of whose rods reach into the other's patch
        auto phase3 = [&](int cnt) {
            if (lane < cnt) {
                const PfItem2 it = q2[(h2 + lane) & (PF_Q - 1)];
                const scgpu_iaparam* ia = ia_one ? ia_one : &s.ia[it.tb];
                const int g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
                PatchArgs P1, P2;
                load_patch_args(s.rec + (size_t)it.si * REC, pn1, is_chiral(g0), P1);
                load_patch_args(s.rec + (size_t)it.sj * REC, pn2, is_chiral(g1), P2);
                const double e = atr_e(*ia, P1.dir, P2.dir, P1.pdir, P2.pdir, mk(it.rx, it.ry, it.rz), pn1, pn2, it.S1, it.S2, it.T1, it.T2);
                fl.e[it.p].y += e;
                if (mirror && e != 0.0) {      // every-particle pass: the partner lists this pair too -- find it in the partner's span (a few entries)
                    const int4 cj = fl.chunks[w_orig(s.posw[it.sj].w)];
                    bool found = false;
                    for (int k = 0; k < cj.y; k++) if (fl.pair[cj.x + k].y == it.si) { fl.e[cj.x + k].y += e; found = true; break; }
                    if (!found) atomicOr(fl.overflow, 8);      // cannot happen: a pair inside the patch range is listed by both sides
                }
            }
            __syncwarp();
            h2 = (h2 + cnt) & (PF_Q - 1);
            n2 -= cnt;
        };
        // ---- phase 2: rod 1 against the patch of rod 2 (S1, S2), survivors of phase 1 only
        auto phase2 = [&](int cnt) {
            bool keep = false;
            PfItem2 out;
            if (lane < cnt) {
                const PfItem1 it = q1[(h1 + lane) & (PF_Q - 1)];
                const scgpu_iaparam* ia = ia_one ? ia_one : &s.ia[it.tb];
                const int kind = (int)ia->reserved[0], g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
                const bool second_psc = (kind == K_SC_PSC) || (kind == K_SC_PSCCPSC && !is_psc_family(g0));
                PatchArgs P1, P2;
                load_patch_args(s.rec + (size_t)it.si * REC, pn1, is_chiral(g0), P1);
                load_patch_args(s.rec + (size_t)it.sj * REC, pn2, is_chiral(g1), P2);
                const v3 vec1 = neg(mk(it.rx, it.ry, it.rz));
                double S1 = 0.0, S2 = 0.0;
                const int nn = second_psc ? patch_intersect<false>(P2.dir, P1.dir, P2, vec1, S1, S2, ia->pcanglsw[2 * pn2 + 1], ia->rcutSq, ia->half_len[1], ia->half_len[0])
                                          : patch_intersect<true>(P2.dir, P1.dir, P2, vec1, S1, S2, ia->pcanglsw[2 * pn2 + 1], ia->rcutSq, ia->half_len[1], ia->half_len[0]);
                keep = nn >= 2;
                out.p = it.p; out.si = it.si; out.sj = it.sj; out.tb = it.tb; out.rx = it.rx; out.ry = it.ry; out.rz = it.rz;
                out.T1 = it.T1; out.T2 = it.T2; out.S1 = S1; out.S2 = S2;
            }
            __syncwarp();
            h1 = (h1 + cnt) & (PF_Q - 1);
            n1 -= cnt;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) q2[(h2 + n2 + __popc(m & lt_mask)) & (PF_Q - 1)] = out;
            n2 += __popc(m);
            __syncwarp();
        };
        // ---- phase 1: rod 2 against the patch of rod 1 (T1, T2), every listed pair, 32 at a time
        auto phase1 = [&](int base) {
            const int q = base + lane;
            bool keep = false;
            PfItem1 it;
            if (q < total) {
                const int p = fl.plist[q];
                const int2 pr = fl.pair[p];
                const double4 pi = ldg256(s.posw + pr.x), pj = ldg256(s.posw + pr.y);
                const v3 r_cm = image(s.box, mk(pi.x, pi.y, pi.z), mk(pj.x, pj.y, pj.z));
                const int tb = w_type(pi.w) * s.ntypes + w_type(pj.w);
                const scgpu_iaparam* ia = ia_one ? ia_one : &s.ia[tb];
                const int kind = (int)ia->reserved[0], g0 = (int)ia->geotype[0], g1 = (int)ia->geotype[1];
                const bool firstT = is_two_patch(g0), secondT = is_two_patch(g1);
                const bool ok = (combo == 0) || (combo == 1 && firstT) || (combo == 2 && secondT) || (combo == 3 && firstT && secondT);
                if (ok) {
                    const bool first_psc = (kind == K_SC_PSC) || (kind == K_SC_PSCCPSC && is_psc_family(g0));
                    PatchArgs P1, P2;
                    load_patch_args(s.rec + (size_t)pr.x * REC, pn1, is_chiral(g0), P1);
                    load_patch_args(s.rec + (size_t)pr.y * REC, pn2, is_chiral(g1), P2);
                    double T1 = 0.0, T2 = 0.0;
                    const int nn = first_psc ? patch_intersect<false>(P1.dir, P2.dir, P1, r_cm, T1, T2, ia->pcanglsw[2 * pn1], ia->rcutSq, ia->half_len[0], ia->half_len[1])
                                             : patch_intersect<true>(P1.dir, P2.dir, P1, r_cm, T1, T2, ia->pcanglsw[2 * pn1], ia->rcutSq, ia->half_len[0], ia->half_len[1]);
                    keep = nn >= 2;
                    it.p = p; it.si = pr.x; it.sj = pr.y; it.tb = tb; it.rx = r_cm.x; it.ry = r_cm.y; it.rz = r_cm.z; it.T1 = T1; it.T2 = T2;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) q1[(h1 + n1 + __popc(m & lt_mask)) & (PF_Q - 1)] = it;
            n1 += __popc(m);
            __syncwarp();
        };
        // every phase appears ONCE in the code (the three bodies are large; duplicates would thrash the instruction cache): a small
        // warp-uniform scheduler runs the latest phase that has a full warp of work, and drains the rings at the end
        int base = gw * 32;
        bool draining = false;
        for (;;) {
            if (n2 >= 32 || (draining && n1 == 0 && n2 > 0)) { phase3(min(n2, 32)); continue; }
            if (n1 >= 32 || (draining && n1 > 0)) { phase2(min(n1, 32)); continue; }
            if (base < total) { phase1(base); base += nwarps * 32; continue; }
            if (!draining) { draining = true; continue; }
            break;
        }
    }
}

// eight lanes per particle: each group reads its particle's span of the list 128 bytes at a time, lane k of the group sums
// entries k, k+8, ... in order, then a fixed three-step shuffle tree -> the same bits on every run
__global__ void __launch_bounds__(256) k_combine_flat(int n, FlatList fl, const double4* __restrict__ posw, double* __restrict__ out) {
    if (*fl.overflow) return;         // nothing valid to sum; the host resets the counters and repeats the launch
    const int sub = threadIdx.x & 7;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *fl.total = 0; *fl.chunk_count = 0; *fl.ptotal = 0; }   // lists consumed (stream order makes the reset safe)
    double e = 0.0;
    if (t < n) {
        const int4 ch = fl.chunks[t];        // every gate of the flat pipeline hands a particle ONE span of the list
        for (int r = sub; r < ch.y; r += 8) {
            const double2 v = fl.e[ch.x + r];
            e += v.x + v.y;      // per pair (cheap + patch), then accumulate
        }
    }
    for (int o = 4; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (t < n && sub == 0) out[t] = e;
}

__device__ __noinline__ double wall_energy_rec(const DevSys& s, const double* r, int type) {        // r: internal record (pair_energy.cuh layout)
    return wall_energy(s.wall[type], s.exter_sqmaxcut, s.box[2], r[R_POS + 2], ld3(r + R_DIR), ld3(r + R_PD0), ld3(r + R_PD1), ld3(r + R_CH0), ld3(r + R_CH1));
}

#include "sweep.cuh"
#include "sweep_rounds.cuh"
#include "sweep_phased.cuh"

// one thread per listed pair; grid-stride because the list length lives on the device
__global__ void __launch_bounds__(128, 4)
k_patch(DevSys s, PatchList pl, const int* __restrict__ targets, int mode, int excl_lo, double* __restrict__ e_pairs) {
    int total = *pl.total;
    if (total > pl.cap) total = pl.cap;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        int2 pr = pl.pair[p];
        const double* s1;
        int type1;
        if (pr.x >= 0) { s1 = s.rec + (size_t)pr.x * REC; type1 = w_type(s.posw[pr.x].w); }
        else {
            int t = -1 - pr.x;
            s1 = pl.trial_rec + (size_t)t * REC;
            type1 = s.type[mode == 0 ? targets[t] : excl_lo + t];
        }
        double4 pw = s.posw[pr.y];
        v3 r_cm = image(s.box, ld3(s1 + R_POS), mk(pw.x, pw.y, pw.z));
        double e = pair_energy_patch(s.ia[type1 * s.ntypes + w_type(pw.w)], r_cm, s1, s.rec + (size_t)pr.y * REC);
        pl.e[p] = e;
        if (e_pairs) e_pairs[w_orig(pw.w)] += e;     // single-target calls only: one writer per partner
    }
}

// per trial particle: partial sums of its warps, then its patch terms in list order
__global__ void k_combine(int m, int gw, PatchList pl, const double* __restrict__ warp_partial, double* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { *pl.total = 0; *pl.chunk_count = 0; }   // the lists are consumed: reset for the next launch (stream order makes this safe)
    if (t >= m) return;
    double e = 0.0;
    for (int k = 0; k < gw; k++) e += warp_partial[t * gw + k];
    for (int k = 0; k < gw; k++) {
        for (int cid = pl.warp_head[t * gw + k]; cid >= 0;) {
            int4 ch = pl.chunks[cid];
            for (int r = 0; r < ch.y; r++) e += pl.e[ch.x + r];
            cid = ch.z;
        }
    }
    out[t] = e;
}

// deterministic total: one block, each thread strides in a fixed pattern, fixed tree
// Sum of n doubles in an order that depends on n alone: block b of RF_BLOCKS sums its contiguous chunk (thread t takes every
// RF_THREADS-th element, then a shared-memory tree), and whichever block finishes last adds the RF_BLOCKS partial sums in index
// order. scratch: RF_BLOCKS doubles followed by one unsigned counter that the last block leaves at zero again.
constexpr int RF_BLOCKS = 128, RF_THREADS = 256;
__global__ void __launch_bounds__(RF_THREADS) k_reduce_fixed(int n, const double* __restrict__ v, double* __restrict__ out, double* __restrict__ scratch) {
    __shared__ double sh[RF_THREADS];
    __shared__ bool last;
    const int chunk = (n + RF_BLOCKS - 1) / RF_BLOCKS;
    const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = lo + threadIdx.x;
    for (; i + 3 * RF_THREADS < hi; i += 4 * RF_THREADS) { a0 += v[i]; a1 += v[i + RF_THREADS]; a2 += v[i + 2 * RF_THREADS]; a3 += v[i + 3 * RF_THREADS]; }
    for (; i < hi; i += RF_THREADS) a0 += v[i];
    sh[threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    for (int o = RF_THREADS >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    unsigned* counter = (unsigned*)(scratch + RF_BLOCKS);
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(counter, 1u) == RF_BLOCKS - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    sh[threadIdx.x] = threadIdx.x < RF_BLOCKS ? ((volatile double*)scratch)[threadIdx.x] : 0.0;
    __syncthreads();
    for (int o = RF_THREADS >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0]; *counter = 0u; }
}

// ------------------------------------------------------------------------------------------------
// external wall ([EXTER]): O(N) per-particle term the reference's calculators add to every sum when topo.exter.exist
// (mc/totalenergycalculator.h:348-350, 377-378, 410-411, 435-449, 476-495, 516-518)
// ------------------------------------------------------------------------------------------------
// out[t] += extere2(particle of result t). mode 0: targets[] (+ trial states in C-ABI layout); 1 / 2: every particle, out indexed by
// original index; 3: molecule members first .. first + m (+ trial states)
__global__ void k_add_exter(DevSys s, int mode, int m, const int* __restrict__ targets, const double* __restrict__ trial_states, int first,
                            double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int target = mode == 0 ? targets[t] : (mode == 3 ? first + t : t);
    double r[REC];
    if ((mode == 0 || mode == 3) && trial_states) {
#pragma unroll
        for (int k = 0; k < 30; k++) r[k] = trial_states[(size_t)t * 30 + c_api_of[k]];
    } else {
        const double* g = s.rec + (size_t)s.slot_of[target] * REC;
#pragma unroll
        for (int k = 0; k < 30; k++) r[k] = g[k];
    }
    out[t] += wall_energy_rec(s, r, s.type[target]);
}

// ------------------------------------------------------------------------------------------------
// overlap: one warp per target, warp-vote early exit (Conf::overlapAll / checkall)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OTA_THREADS)
k_overlap(DevSys s, int m, const int* __restrict__ targets, const double* __restrict__ trial_states, int all_pairs, int variant,
          int* __restrict__ flag_out) {
    __shared__ double sh_rec[OTA_WARPS][REC];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = blockIdx.x * OTA_WARPS + wid;
    if (t >= m) return;
    int target = all_pairs ? t : targets[t];
    double* s1 = sh_rec[wid];
    if (!all_pairs && trial_states) { if (lane < 30) s1[lane] = trial_states[(size_t)t * 30 + c_api_of[lane]]; }
    else s1[lane] = s.rec[(size_t)s.slot_of[target] * REC + lane];
    __syncwarp();
    const int type1 = s.type[target];
    const v3 p1 = ld3(s1 + R_POS);
    const int cx = cell_coord(p1.x + s.shift[0], s.nc[0]), cy = cell_coord(p1.y + s.shift[1], s.nc[1]), cz = cell_coord(p1.z + s.shift[2], s.nc[2]);
    const int nx = s.nc[0] == 1 ? 1 : 3, ny = s.nc[1] == 1 ? 1 : 3, nz = s.nc[2] == 1 ? 1 : 3;
    int found = 0;
    for (int k = 0; k < nx * ny * nz && !found; k++) {
        int dx = k % nx, dy = (k / nx) % ny, dz = k / (nx * ny);
        int ccx = nx == 1 ? 0 : (cx + dx - 1 + s.nc[0]) % s.nc[0];
        int ccy = ny == 1 ? 0 : (cy + dy - 1 + s.nc[1]) % s.nc[1];
        int ccz = nz == 1 ? 0 : (cz + dz - 1 + s.nc[2]) % s.nc[2];
        int c = (ccz * s.nc[1] + ccy) * s.nc[0] + ccx;
        int b = s.cell_start[c], e = s.cell_start[c + 1];
        for (int base = b; base < e && !found; base += 32) {
            int j = base + lane;
            int hit = 0;
            if (j < e) {
                double4 pw = s.posw[j];
                int orig = w_orig(pw.w);
                // checkall tests each unordered pair once as (i, j>i) (Conf.cpp:259-265)
                if (orig != target && (!all_pairs || orig > target)) {
                    v3 r_cm = image(s.box, p1, mk(pw.x, pw.y, pw.z));
                    if (dot(r_cm, r_cm) <= s.sqmaxcut)
                        hit = overlap_pair(s.box, s.ia, s.ntypes, r_cm, s1, type1, s.rec + (size_t)j * REC, w_type(pw.w), variant);
                }
            }
            found = __any_sync(0xffffffffu, hit);   // early exit by warp vote
        }
    }
    if (lane == 0 && found) {
        if (all_pairs) atomicOr(flag_out, 1); else flag_out[t] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_flush(double* buf, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = buf[i] * 0.5 + 1.0;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct scgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // timer
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;      // bounce-buffer hand-off in scgpu_set_particles
    int sm_count = 148;
    // topology
    int ntypes = 0, nmol = 0;
    std::vector<scgpu_iaparam> h_ia;
    std::vector<scgpu_molparam> h_mol;
    float* d_reach2 = nullptr;
    scgpu_iaparam* d_ia = nullptr;
    scgpu_molparam* d_mol = nullptr;
    double sqmaxcut = 0, maxcut = 0;
    bool exter = false;              // [EXTER] wall potential present (scgpu_set_exter)
    double exter_sqmaxcut = 0;
    WallParam* d_wall = nullptr;
    double min_reach2 = 0;           // the smallest non-zero squared reach of the table (bounds the FP32 rounding the row-unit gates may carry)
    // particles
    int n = 0, cap = 0;
    double box[3] = {0, 0, 0};
    double shift[3] = {0, 0, 0};
    std::vector<int> h_cell_of;
    double* d_api = nullptr;
    double* d_compact = nullptr;     // staging of 9-double records (scgpu_set_particles_compact)
    double4* d_posw = nullptr;
    double* d_rec = nullptr;
    float4 *d_p32 = nullptr, *d_d32 = nullptr;      // FP32 copies of position / axis for the gates (k_cell_place, k_make_f32)
    bool f32_valid = false;
    int *d_type = nullptr, *d_moltype = nullptr, *d_cell_of = nullptr, *d_order = nullptr, *d_slot_of = nullptr, *d_tmp = nullptr;
    int *d_counts = nullptr, *d_cell_start = nullptr, *d_cursor = nullptr;
    int cells_cap = 0;
    unsigned* d_ticket = nullptr;    // last-block ticket of k_cell_count
    int *d_fine_of = nullptr, *d_fine_start = nullptr;      // sub-cell sort (sub > 1): per particle, per sub-cell
    int fine_cap = 0;
    int sub = 1;                     // sub-cells per cell and axis of the current cell list
    unsigned present_mask = 0;       // particle types present (bit per type, types < 32)
    long long type_count[40] = {0};
    std::vector<char> mol_used;      // molecule types present
    unsigned fine_heavy = 0;         // types whose reach exceeds a sub-cell: their pairs stay on the coarse grid
    int *d_heavy = nullptr, *d_nheavy = nullptr;             // sorted slots of the heavy-type particles
    int heavy_cap = 0;
    int nc[3] = {1, 1, 1};
    int ncells = 1;
    bool cells_valid = false;
    bool types_valid = false;        // d_type / d_moltype hold the types of the current n particles
    bool api_stale = false;          // sorted arrays are newer than d_api (after device sweeps)
    // scratch
    double* d_out = nullptr;         // n doubles
    double* d_pairs = nullptr;       // n doubles
    // patch work list (k_gate_cheap -> k_patch -> k_combine)
    int2* d_pl_pair = nullptr;
    double* d_pl_e = nullptr;
    int* d_pl_total = nullptr;
    int* d_pl_overflow = nullptr;
    int pl_cap = 0;
    bool rods_only = false;          // every particle an un-bonded rod: the specialised kernels apply
    int last_overflow = 0;
    unsigned heavy_types = 0;        // particle types whose targets k_gate_rows_gen could not hold: they go to k_gate_cells
    int one_type = -1;               // >= 0: every particle present has this type (single table entry -> kernel parameter)
    bool use_rows = true;            // k_gate_rows until a launch reports a layout it cannot hold (then k_gate_cells for good)
    bool any_two_patch = false;      // some particle type present carries a second patch (TPSC/TCPSC/TCHPSC/TCHCPSC)
    int* d_warp_head = nullptr;
    int4* d_chunks = nullptr;
    int chunk_cap = 0;
    double* d_warp_partial = nullptr;
    double* d_trial_rec = nullptr;   // [trial_cap][REC]
    // flat pipeline lists
    int2* d_fl_pair = nullptr;
    double2* d_fl_e = nullptr;
    int* d_fl_plist = nullptr;
    int* d_fl_head = nullptr;
    int4* d_fl_chunks = nullptr;
    int fl_cap = 0, fl_chunk_cap = 0;
    double* d_trial = nullptr;       // staging for trial states
    int* d_targets = nullptr;
    int trial_cap = 0;
    double* d_scalar = nullptr;      // 16 doubles: [0] total, [8..] replica record
    double* d_reduce = nullptr;      // k_reduce_fixed scratch: RF_BLOCKS partial sums + counter
    int* d_flags = nullptr;          // n ints
    unsigned long long* d_counters = nullptr;   // 8
    void* d_sweep_acc = nullptr;
    int sweep_acc_cap = 0;
    int* d_sw_maxc = nullptr; int* h_sw_maxc = nullptr;      // largest staged neighbourhood the cell walk met in the last sweep (device, pinned host copy)
    SwTrial* d_sw_trials = nullptr; int sw_trial_cap = 0;      // phased sweeps: per-trial records, per-cell records, per-pair flags
    SwCell* d_sw_cells = nullptr; int sw_cell_cap = 0;
    unsigned short* d_sw_meta = nullptr; int sw_meta_cap = 0;
    bool sweep_one_cell = getenv("SCGPU_SWEEP_ONE_CELL") != nullptr;      // diagnostic: sweeps without the cell decomposition
    double* d_flush = nullptr;
    size_t flush_n = 0;
    char* h_small = nullptr;         // 1 KB pinned: single-call inputs [0..255], results [256..511], update staging [512..751]
    char* d_single = nullptr;        // 256 B device staging of a single call: trial record [0..239], target index [240..243]
    char* d_wl = nullptr;            // scgpu_wl_order: block partials, results, counters (WL_BYTES)
    int *d_wl_mesh = nullptr, *d_wl_parent = nullptr, *d_wl_size = nullptr;      // the hole mesh of wlm 2 and its union-find arrays
    int wl_mesh_cap = 0;
    int wl_mesh_len = 0;             // mesh points of the last wlm-2 call (0: none); wl_mesh_labelled: d_wl_mesh already holds Mesh::data
    bool wl_mesh_labelled = false;
    void* h_pinned = nullptr;        // pinned staging, grown on demand
    size_t pinned_bytes = 0;
    int64_t launches = 0;
};

// what the kernels' specialisations depend on, from the census of particle types (type_count) and molecule types (mol_used) present
static void derive_type_flags(scgpu_ctx* c) {
    c->use_rows = true;
    c->heavy_types = 0;
    std::vector<char> tu(c->ntypes, 0);
    for (int t = 0; t < c->ntypes; t++) tu[t] = c->type_count[t] > 0;
    c->present_mask = 0;
    for (int t = 0; t < c->ntypes && t < 32; t++) if (tu[t]) c->present_mask |= 1u << t;
    bool rods = true;
    for (int a = 0; a < c->ntypes && rods; a++) for (int b = 0; b < c->ntypes && rods; b++) {
        if (!tu[a] || !tu[b]) continue;
        int k = (int)c->h_ia[(size_t)a * c->ntypes + b].reserved[0];
        if (!(k == K_SC_PSCCPSC || k == K_SC_CPSC || k == K_SC_PSC || k == K_SC_SCN)) rods = false;
    }
    for (int mm = 0; mm < c->nmol && rods; mm++) if (c->mol_used[mm] && c->h_mol[mm].mol_size != 1.0) rods = false;
    c->rods_only = rods;
    {
        int nt = 0, last = -1;
        for (int a = 0; a < c->ntypes; a++) if (tu[a]) { nt++; last = a; }
        c->one_type = nt == 1 ? last : -1;
    }
    c->any_two_patch = false;
    for (int a = 0; a < c->ntypes; a++) if (tu[a]) {
        int g = (int)c->h_ia[(size_t)a * c->ntypes + a].geotype[0];
        if (g == SCGPU_TPSC || g == SCGPU_TCPSC || g == SCGPU_TCHPSC || g == SCGPU_TCHCPSC) c->any_two_patch = true;
    }
}

static int ensure_pinned(scgpu_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_bytes) return 0;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr; c->pinned_bytes = 0;
    CK(cudaMallocHost(&c->h_pinned, bytes));
    c->pinned_bytes = bytes;
    return 0;
}

static DevSys view(const scgpu_ctx* c) {
    DevSys s;
    s.n = c->n; s.ntypes = c->ntypes; s.nmol = c->nmol;
    for (int d = 0; d < 3; d++) { s.nc[d] = c->nc[d]; s.box[d] = c->box[d]; s.shift[d] = c->shift[d]; }
    s.ncells = c->ncells;
    s.api = c->d_api; s.posw = c->d_posw; s.rec = c->d_rec; s.p32 = c->d_p32; s.d32 = c->d_d32; s.cell_start = c->d_cell_start; s.order = c->d_order;
    s.fine_start = c->sub > 1 ? c->d_fine_start : c->d_cell_start; s.sub = c->sub;
    s.slot_of = c->d_slot_of; s.type = c->d_type; s.moltype = c->d_moltype; s.ia = c->d_ia; s.mol = c->d_mol; s.reach2 = c->d_reach2;
    s.sqmaxcut = c->sqmaxcut;
    s.wall = c->exter ? c->d_wall : nullptr;
    s.exter_sqmaxcut = c->exter_sqmaxcut;
    return s;
}

extern "C" int scgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int scgpu_create(scgpu_ctx** out, int device) {
    ARG(out != nullptr, "scgpu_create: out is NULL");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) { g_err = "scgpu_create: no CUDA device (this library has no CPU fallback)"; return SCGPU_ERR_CUDA; }
    ARG(device >= 0 && device < ndev, "scgpu_create: bad device ordinal");
    CK(cudaSetDevice(device));
    scgpu_ctx* c = new scgpu_ctx();
    c->device = device;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->ev2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev3, cudaEventDisableTiming));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaMallocHost((void**)&c->h_small, 1024));
    CK(cudaMalloc((void**)&c->d_single, 256));
    CK(cudaMalloc(&c->d_scalar, 16 * sizeof(double)));
    CK(cudaMalloc(&c->d_reduce, (RF_BLOCKS + 2) * sizeof(double)));
    CK(cudaMemset(c->d_reduce, 0, (RF_BLOCKS + 2) * sizeof(double)));
    CK(cudaMalloc(&c->d_counters, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&c->d_ticket, 2 * sizeof(unsigned)));
    CK(cudaMemset(c->d_ticket, 0, 2 * sizeof(unsigned)));
    CK(cudaMalloc(&c->d_pl_total, 8 * sizeof(int)));     // [0] patch pairs, [1] overflow flag, [2] patch chunks, [3] flat pairs, [4] flat chunks, [5] flat patch pairs
    CK(cudaMemset(c->d_pl_total, 0, 8 * sizeof(int)));
    c->d_pl_overflow = c->d_pl_total + 1;
    CK(cudaMemset(c->d_scalar, 0, 16 * sizeof(double)));
    *out = c;
    return SCGPU_OK;
}

static void free_particles(scgpu_ctx* c) {
    cudaFree(c->d_api); cudaFree(c->d_compact); cudaFree(c->d_posw); cudaFree(c->d_rec); cudaFree(c->d_type); cudaFree(c->d_moltype);
    cudaFree(c->d_p32); cudaFree(c->d_d32); c->d_p32 = c->d_d32 = nullptr;
    cudaFree(c->d_cell_of); cudaFree(c->d_order); cudaFree(c->d_slot_of); cudaFree(c->d_tmp); cudaFree(c->d_out);
    cudaFree(c->d_pairs); cudaFree(c->d_flags);
    cudaFree(c->d_fl_pair); cudaFree(c->d_fl_e); cudaFree(c->d_fl_plist); cudaFree(c->d_fl_head); cudaFree(c->d_fl_chunks);
    c->d_fl_pair = nullptr; c->d_fl_e = nullptr; c->d_fl_plist = nullptr; c->d_fl_head = nullptr; c->d_fl_chunks = nullptr;
    cudaFree(c->d_pl_pair); cudaFree(c->d_pl_e); cudaFree(c->d_warp_head); cudaFree(c->d_chunks); cudaFree(c->d_warp_partial);
    c->d_pl_pair = nullptr; c->d_pl_e = nullptr; c->d_warp_head = nullptr; c->d_chunks = nullptr; c->d_warp_partial = nullptr;
    c->d_api = nullptr; c->d_compact = nullptr; c->d_posw = nullptr; c->d_rec = nullptr; c->d_type = c->d_moltype = c->d_cell_of = c->d_order = c->d_slot_of = c->d_tmp = nullptr;
    c->d_out = c->d_pairs = nullptr; c->d_flags = nullptr;
    c->cap = 0;
}

extern "C" int scgpu_destroy(scgpu_ctx* c) {
    if (!c) return SCGPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_particles(c);
    cudaFree(c->d_ia); cudaFree(c->d_mol); cudaFree(c->d_reach2); cudaFree(c->d_counts); cudaFree(c->d_cell_start); cudaFree(c->d_cursor);
    cudaFree(c->d_ticket); cudaFree(c->d_fine_of); cudaFree(c->d_fine_start); cudaFree(c->d_heavy); cudaFree(c->d_nheavy); cudaFree(c->d_wall);
    cudaFree(c->d_sw_maxc); if (c->h_sw_maxc) cudaFreeHost(c->h_sw_maxc);
    cudaFree(c->d_sw_trials); cudaFree(c->d_sw_cells); cudaFree(c->d_sw_meta);
    cudaFree(c->d_sweep_acc);
    cudaFree(c->d_trial); cudaFree(c->d_trial_rec); cudaFree(c->d_pl_total); cudaFree(c->d_targets); cudaFree(c->d_scalar); cudaFree(c->d_reduce); cudaFree(c->d_counters); cudaFree(c->d_flush);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_small) cudaFreeHost(c->h_small);
    cudaFree(c->d_single);
    cudaFree(c->d_wl); cudaFree(c->d_wl_mesh); cudaFree(c->d_wl_parent); cudaFree(c->d_wl_size);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->ev3);
    cudaStreamDestroy(c->stream);
    delete c;
    return SCGPU_OK;
}

// functor kind of PairE::initIntFCE (scOOP/mc/paire.cpp:6-80): the ifs are evaluated in the reference's order,
// later matches overwrite earlier ones
static int functor_kind(int g, int o) {
    auto psc = [](int x) { return x == SCGPU_PSC || x == SCGPU_CHPSC || x == SCGPU_TPSC || x == SCGPU_TCHPSC; };
    auto cpsc = [](int x) { return x == SCGPU_CPSC || x == SCGPU_CHCPSC || x == SCGPU_TCPSC || x == SCGPU_TCHCPSC; };
    auto sph = [](int x) { return x == SCGPU_SPA || x == SCGPU_SPN; };
    int k = K_EBASIC;
    if ((cpsc(g) && psc(o)) || (psc(g) && cpsc(o))) k = K_SC_PSCCPSC;
    if (cpsc(g) && cpsc(o)) k = K_SC_CPSC;
    if (psc(g) && psc(o)) k = K_SC_PSC;
    if (g == SCGPU_SCN && o == SCGPU_SCN) k = K_SC_SCN;
    if (g == SCGPU_SCA && o == SCGPU_SCA) k = K_SC_SCA;
    if (g == SCGPU_SPN || o == SCGPU_SPN) k = K_SP_WCA;
    if (g == SCGPU_SPA && o == SCGPU_SPA) k = K_SP_COS2;
    if ((g == SCGPU_SCA && o == SCGPU_SPA) || (g == SCGPU_SPA && o == SCGPU_SCA)) k = K_MIX_SCASPA;
    if ((psc(g) && sph(o)) || (sph(g) && psc(o))) k = K_MIX_PSCSPA;
    if ((cpsc(g) && sph(o)) || (sph(g) && cpsc(o))) k = K_MIX_CPSCSPA;
    return k;
}

extern "C" int scgpu_set_topology(scgpu_ctx* c, int ntypes, const scgpu_iaparam* table, double sqmaxcut, double maxcut,
                                  int nmoltypes, const scgpu_molparam* mol) {
    ARG(c && table && mol, "scgpu_set_topology: NULL argument");
    ARG(ntypes > 0 && ntypes <= 40 && nmoltypes > 0, "scgpu_set_topology: 1 .. 40 particle types (MAXT, scOOP/structures/macros.h:87) and at least one molecule type");
    ARG(maxcut > 0 && sqmaxcut > 0, "scgpu_set_topology: cutoff must be positive");
    CK(cudaSetDevice(c->device));
    c->types_valid = false;          // rods_only / any_two_patch are derived from (topology, types): the next upload must bring types
    c->h_ia.assign(table, table + (size_t)ntypes * ntypes);
    for (auto& p : c->h_ia) {
        p.reserved[0] = (double)functor_kind((int)p.geotype[0], (int)p.geotype[1]);
        // rod pairs: squared centre distance beyond which every term is exactly 0 (pair_energy_cheap's shortcut)
        const double reach = sqrt(fmax(p.rcutSq, p.rcutwcaSq)) + p.half_len[0] + p.half_len[1];
        p.reserved[1] = reach * reach * 1.000001;
        // Ia_param::volume (mc/inicializer.cpp:840-844), the weight of clusterCM (mc/movecreator.cpp:1323-1337); meaningful on the diagonal
        const double hs = p.sigma / 2.0;
        p.reserved[2] = 4.0 / 3.0 * 3.14159265358979323846 * hs * hs * hs + ((int)p.geotype[0] < SCGPU_SPN ? 3.14159265358979323846 / 2.0 * p.len[0] * hs * hs : 0.0);
    }
    c->h_mol.assign(mol, mol + nmoltypes);
    cudaFree(c->d_ia); cudaFree(c->d_mol); cudaFree(c->d_reach2);
    c->d_ia = nullptr; c->d_mol = nullptr; c->d_reach2 = nullptr;
    {   // beyond `reach` (centre distance) every term of the pair energy is EXACTLY zero: rods are at least |r| - l1/2 - l2/2 apart,
        // a rod and a sphere |r| - l/2, spheres |r|; the cutoffs are max(rcut, rcutwca). Stored squared, with a 0.1 % FP32 safety margin.
        std::vector<float> reach((size_t)ntypes * ntypes);
        reach.reserve((size_t)ntypes * ntypes + ntypes);
        for (int a = 0; a < ntypes; a++) for (int b = 0; b < ntypes; b++) {
            const scgpu_iaparam& q = c->h_ia[(size_t)a * ntypes + b];
            int k = (int)q.reserved[0];
            double cut = sqrt(q.rcutSq > q.rcutwcaSq ? q.rcutSq : q.rcutwcaSq);
            if (q.rcut > cut) cut = q.rcut;
            if (q.rcutwca > cut) cut = q.rcutwca;
            double r = 0.0;
            if (k >= K_SC_PSCCPSC && k <= K_SC_SCA) r = cut + q.half_len[0] + q.half_len[1];
            else if (k == K_SP_WCA || k == K_SP_COS2) r = cut;
            else if (k >= K_MIX_SCASPA) r = cut + q.half_len[0] + q.half_len[1];    // one of the two half lengths is 0 (the sphere)
            reach[(size_t)a * ntypes + b] = (float)(r * r * 1.001);
        }
        // [ntypes*ntypes + a]: the largest of row a -- no partner of a type-a particle interacts beyond it (neighbour-cell culling)
        for (int a = 0; a < ntypes; a++) {
            float m = 0.f;
            for (int b = 0; b < ntypes; b++) m = fmaxf(m, reach[(size_t)a * ntypes + b]);
            reach.push_back(m);
        }
        // segment lower-bound filter (sweep.cuh, lb_reject): squared surface cutoff of rod-rod pairs and the rods' half lengths
        for (int a = 0; a < ntypes; a++) for (int b = 0; b < ntypes; b++) {
            const scgpu_iaparam& q = c->h_ia[(size_t)a * ntypes + b];
            const int k = (int)q.reserved[0];
            double cut = sqrt(q.rcutSq > q.rcutwcaSq ? q.rcutSq : q.rcutwcaSq);
            if (q.rcut > cut) cut = q.rcut;
            if (q.rcutwca > cut) cut = q.rcutwca;
            reach.push_back((k >= K_SC_PSCCPSC && k <= K_SC_SCA) ? (float)(cut * cut * 1.001) : 0.f);
        }
        for (int a = 0; a < ntypes; a++) {
            const scgpu_iaparam& q = c->h_ia[(size_t)a * ntypes + a];
            reach.push_back((int)q.geotype[0] < SCGPU_SPN ? (float)q.half_len[0] : 0.f);
        }
        c->min_reach2 = 0.0;
        for (size_t k = 0; k < (size_t)ntypes * ntypes; k++) if (reach[k] > 0.f && (c->min_reach2 == 0.0 || reach[k] < c->min_reach2)) c->min_reach2 = reach[k];
        CK(cudaMalloc(&c->d_reach2, reach.size() * sizeof(float)));
        CK(cudaMemcpy(c->d_reach2, reach.data(), reach.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&c->d_ia, c->h_ia.size() * sizeof(scgpu_iaparam)));
    CK(cudaMalloc(&c->d_mol, c->h_mol.size() * sizeof(scgpu_molparam)));
    CK(cudaMemcpyAsync(c->d_ia, c->h_ia.data(), c->h_ia.size() * sizeof(scgpu_iaparam), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_mol, c->h_mol.data(), c->h_mol.size() * sizeof(scgpu_molparam), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->ntypes = ntypes; c->nmol = nmoltypes; c->sqmaxcut = sqmaxcut; c->maxcut = maxcut;
    c->cells_valid = false;
    c->exter = false;                // the wall parameters derive from this table: scgpu_set_exter comes after scgpu_set_topology
    return SCGPU_OK;
}

// topo.exter (scOOP/structures/structures.h:259-275), filled by Topo::genParamPairs / genTopoParams (structures/topo.cpp:120-130, 151-152):
// every particle type interacts with the wall through its own Ia_param with the wall's mixing rules
extern "C" int scgpu_set_exter(scgpu_ctx* c, int exist, double thickness, double epsilon, double attraction) {
    ARG(c, "scgpu_set_exter: NULL context");
    ARG(c->ntypes > 0, "scgpu_set_exter: call scgpu_set_topology first");
    CK(cudaSetDevice(c->device));
    c->exter = false;
    if (!exist) return SCGPU_OK;
    std::vector<WallParam> w(c->ntypes);
    double sq = 0.0, maxlength = 0.0;
    for (int i = 0; i < c->ntypes; i++) {
        const scgpu_iaparam& q = c->h_ia[(size_t)i * c->ntypes + i];
        memset(&w[i], 0, sizeof(WallParam));
        if (maxlength < q.len[0]) maxlength = q.len[0];
        if ((int)q.geotype[0] == 0) continue;
        w[i].geotype = (int)q.geotype[0];
        w[i].len0 = q.len[0]; w[i].half_len0 = q.half_len[0];
        w[i].sigma = (q.sigma + thickness) * 0.5;
        w[i].rcutwca = (w[i].sigma) * pow(2.0, 1.0 / 6.0);
        w[i].epsilon = sqrt(q.epsilon * epsilon);
        w[i].pswitch = (q.pswitch + attraction) * 0.5;
        w[i].pdis = (q.pdis - q.rcutwca + 0.0) * 0.5 + w[i].rcutwca;
        w[i].rcut = w[i].pswitch + w[i].pdis;
        if (w[i].rcut > sq) sq = w[i].rcut;
    }
    sq += maxlength;
    sq *= sq * 1.1;
    c->exter_sqmaxcut = sq;
    if (!c->d_wall) CK(cudaMalloc(&c->d_wall, 40 * sizeof(WallParam)));
    CK(cudaMemcpyAsync(c->d_wall, w.data(), (size_t)c->ntypes * sizeof(WallParam), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->exter = true;
    return SCGPU_OK;
}

static int set_particles_impl(scgpu_ctx* c, int n, const double* state30, const int* type, const int* moltype, bool compact);
static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int scgpu_set_particles(scgpu_ctx* c, int n, const double* state30, const int* type, const int* moltype) {
    return set_particles_impl(c, n, state30, type, moltype, false);
}
extern "C" int scgpu_set_particles_compact(scgpu_ctx* c, int n, const double* state9, const int* type, const int* moltype) {
    return set_particles_impl(c, n, state9, type, moltype, true);
}

static int set_particles_impl(scgpu_ctx* c, int n, const double* state30, const int* type, const int* moltype, bool compact) {
    ARG(c && state30, "scgpu_set_particles: NULL argument");
    ARG((type == nullptr) == (moltype == nullptr), "scgpu_set_particles: type and moltype must both be given or both be NULL");
    ARG(n > 0 && n < (1 << 24), "scgpu_set_particles: n must be in 1 .. 16 777 215");
    ARG(c->ntypes > 0, "scgpu_set_particles: call scgpu_set_topology first");
    // type == moltype == NULL: same particle count and the same types as the previous upload (a Monte Carlo caller's types
    // never change between configurations); only the coordinates travel
    const bool same_types = type == nullptr;
    ARG(!same_types || (n == c->n && c->types_valid), "scgpu_set_particles: NULL types need a previous upload of the same particle count");
    for (int i = 0; i < n && !same_types; i++) {
        ARG(type[i] >= 0 && type[i] < c->ntypes, "scgpu_set_particles: particle type outside the topology table");
        ARG(moltype[i] >= 0 && moltype[i] < c->nmol, "scgpu_set_particles: molecule type outside the topology table");
    }
    CK(cudaSetDevice(c->device));
    if (n > c->cap) {
        free_particles(c);
        size_t N = (size_t)n;
        CK(cudaMalloc(&c->d_api, N * 30 * sizeof(double)));
        CK(cudaMalloc(&c->d_compact, N * 9 * sizeof(double)));
        CK(cudaMalloc(&c->d_posw, (2 * N + 64) * sizeof(double4)));      // [N, 2N): spare slots for the trial states of a sweep pass (sweep_phased.cuh)
        CK(cudaMalloc(&c->d_p32, N * sizeof(float4)));
        CK(cudaMalloc(&c->d_d32, N * sizeof(float4)));
        CK(cudaMalloc(&c->d_rec, (2 * N + 64) * REC * sizeof(double)));
        CK(cudaMalloc(&c->d_type, N * sizeof(int)));
        CK(cudaMalloc(&c->d_moltype, N * sizeof(int)));
        CK(cudaMalloc(&c->d_cell_of, N * sizeof(int)));
        CK(cudaMalloc(&c->d_order, N * sizeof(int)));
        CK(cudaMalloc(&c->d_slot_of, N * sizeof(int)));
        CK(cudaMalloc(&c->d_tmp, N * sizeof(int)));
        CK(cudaMalloc(&c->d_out, N * sizeof(double)));
        CK(cudaMalloc(&c->d_pairs, N * sizeof(double)));
        CK(cudaMalloc(&c->d_flags, N * sizeof(int)));
        c->pl_cap = (int)(N * 24 < 16384 ? 16384 : N * 24);
        CK(cudaMalloc(&c->d_pl_pair, (size_t)c->pl_cap * sizeof(int2)));
        CK(cudaMalloc(&c->d_pl_e, (size_t)c->pl_cap * sizeof(double)));
        c->fl_cap = (int)(N * 96 < 65536 ? 65536 : N * 96);
        c->fl_chunk_cap = (int)(N + 64) + c->fl_cap / 64 + 64;
        CK(cudaMalloc(&c->d_fl_pair, (size_t)c->fl_cap * sizeof(int2)));
        CK(cudaMalloc(&c->d_fl_e, (size_t)c->fl_cap * sizeof(double2)));
        CK(cudaMalloc(&c->d_fl_plist, (size_t)c->fl_cap * sizeof(int)));
        CK(cudaMalloc(&c->d_fl_head, (N + 64) * sizeof(int)));
        CK(cudaMalloc(&c->d_fl_chunks, (size_t)c->fl_chunk_cap * sizeof(int4)));
        c->chunk_cap = (int)(N + 64) + c->pl_cap / PCAP + 64;
        CK(cudaMalloc(&c->d_warp_head, (N + 64) * sizeof(int)));
        CK(cudaMalloc(&c->d_chunks, (size_t)c->chunk_cap * sizeof(int4)));
        CK(cudaMalloc(&c->d_warp_partial, (N + 64) * sizeof(double)));
        c->cap = n;
    }
    c->n = n;
    // Host -> device. If the caller's buffers are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)
    // the DMA reads them directly; otherwise they are staged through a pinned bounce buffer in 2 MB chunks so that the
    // host memcpy of chunk k+1 overlaps the DMA of chunk k.
    size_t bytes = (size_t)n * (compact ? 9 : 30) * sizeof(double);
    double* d_dst = compact ? c->d_compact : c->d_api;
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    if (is_pinned(state30)) {
        CK(cudaMemcpyAsync(d_dst, state30, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        const size_t CH = 2u << 20;
        if (ensure_pinned(c, 2 * CH)) return SCGPU_ERR_CUDA;
        char* pin = (char*)c->h_pinned;
        cudaEvent_t evs[2] = {c->ev2, c->ev3};
        int k = 0;
        for (size_t off = 0; off < bytes; off += CH, k ^= 1) {
            size_t len = bytes - off < CH ? bytes - off : CH;
            if (off >= 2 * CH) CK(cudaEventSynchronize(evs[k]));          // this half of the bounce buffer is free again
            memcpy(pin + k * CH, (const char*)state30 + off, len);
            CK(cudaMemcpyAsync((char*)d_dst + off, pin + k * CH, len, cudaMemcpyHostToDevice, c->stream));
            CK(cudaEventRecord(evs[k], c->stream));
        }
    }
    if (!same_types) {
        CK(cudaMemcpyAsync(c->d_type, type, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_moltype, moltype, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    if (compact) {
        k_particle_init<<<(n + 127) / 128, 128, 0, c->stream>>>(n, c->ntypes, c->d_compact, c->d_type, c->d_ia, c->d_api);
        c->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(c->stream));
    // specialisation switch: only rod-rod functors and no bonded molecule among the particles present
    if (!same_types) {
        c->types_valid = true;
        for (int t = 0; t < 40; t++) c->type_count[t] = 0;
        for (int i = 0; i < n; i++) c->type_count[type[i]]++;
        c->mol_used.assign(c->nmol, 0);
        for (int i = 0; i < n; i++) c->mol_used[moltype[i]] = 1;
        derive_type_flags(c);
    }
    c->cells_valid = false;
    c->api_stale = false;
    return SCGPU_OK;
}

extern "C" int scgpu_set_box(scgpu_ctx* c, const double box[3]) {
    ARG(c && box, "scgpu_set_box: NULL argument");
    ARG(box[0] > 0 && box[1] > 0 && box[2] > 0, "scgpu_set_box: box edges must be positive");
    // cell ids depend only on fractional coordinates and on ncell = floor(box/maxcut): re-grid only when that changes
    bool regrid = !c->cells_valid;
    for (int d = 0; d < 3; d++) {
        c->box[d] = box[d];
        if (c->maxcut > 0) {
            int nc = (int)floor(box[d] / c->maxcut);
            if (nc < 3) nc = 1;
            if (nc != c->nc[d]) regrid = true;
        }
    }
    if (regrid) c->cells_valid = false;
    return SCGPU_OK;
}

static int sync_api_from_sorted(scgpu_ctx* c) {
    if (!c->api_stale) return 0;
    DevSys s = view(c);
    k_unsort<<<(c->n + 255) / 256, 256, 0, c->stream>>>(s, c->d_api);
    c->launches++;
    CK(cudaGetLastError());
    c->api_stale = false;
    return 0;
}

// shift: fractional grid offset; sweep_k = 0: the energy grid (cells >= maxcut); sweep_k = K >= 1: the checkerboard grid of the
// sweeps: cells of edge >= maxcut / K, a multiple of K + 1 cells per axis (>= 2 (K + 1), so that the 2K + 1 neighbour cells of an
// axis are distinct), else that axis is one cell
static int build_cells_impl(scgpu_ctx* c, const double shift[3], int sweep_k) {
    ARG(c, "scgpu_build_cells: NULL context");
    ARG(c->n > 0 && c->ntypes > 0 && c->box[0] > 0, "scgpu_build_cells: topology, particles and box must be set first");
    CK(cudaSetDevice(c->device));
    if (sync_api_from_sorted(c)) return SCGPU_ERR_CUDA;
    for (int d = 0; d < 3; d++) {
        int nc = (int)floor(c->box[d] / c->maxcut);
        if (nc < 3) nc = 1;      // fewer than 3 cells: the +-1 neighbours would alias through the periodic image
        if (sweep_k > 0) {
            nc = (int)floor(c->box[d] * sweep_k / c->maxcut);
            nc -= nc % (sweep_k + 1);
            if (nc < 2 * (sweep_k + 1)) nc = 1;
            if (c->sweep_one_cell) nc = 1;      // diagnostic (SCGPU_SWEEP_ONE_CELL): no decomposition, one warp walks the whole system
        }
        c->nc[d] = nc;
        c->shift[d] = shift[d];
    }
    long long ncells = (long long)c->nc[0] * c->nc[1] * c->nc[2];
    ARG(ncells < (1ll << 30), "scgpu_build_cells: too many cells");
    c->ncells = (int)ncells;
    // ---- sub-cells. The cell edge is set by the LONGEST interaction of the system (maxcut); where most particles only have short
    // ones -- lipid beads among a few long rods: cells 9.6 wide, bead reach 2.7 -- a particle would test thousands of candidates for
    // a few dozen partners. The cells are then cut into sub^3 sub-cells, particles sorted by (cell, sub-cell, index), and the
    // every-particle gate of the "light" types scans the 27 sub-cells around a target (k_gate_fine); pairs that involve a "heavy" type
    // (reach beyond a sub-cell) stay on the cell level. The cell assignment itself (definition C1) is unchanged.
    c->sub = 1;
    c->fine_heavy = 0;
    if (sweep_k == 0 && c->types_valid && !c->rods_only && c->ntypes <= GG_MAXT && c->n >= 8192 && c->nc[0] >= 3 && c->nc[1] >= 3 && c->nc[2] >= 3
        && c->use_rows && getenv("SCGPU_NO_SUBCELLS") == nullptr) {
        double edge = c->box[0] / c->nc[0];
        for (int d = 1; d < 3; d++) edge = fmin(edge, c->box[d] / c->nc[d]);
        unsigned light = c->present_mask;
        const int T = c->ntypes;
        auto pair_reach = [&](int a, int b) {
            const scgpu_iaparam& q = c->h_ia[(size_t)a * T + b];
            const int k = (int)q.reserved[0];
            double cut = sqrt(q.rcutSq > q.rcutwcaSq ? q.rcutSq : q.rcutwcaSq);
            if (q.rcut > cut) cut = q.rcut;
            if (q.rcutwca > cut) cut = q.rcutwca;
            if (k == K_EBASIC) return 0.0;
            return cut + q.half_len[0] + q.half_len[1];
        };
        int best = 1;
        while (light) {
            double rmax = 0.0;
            int worst = -1;
            for (int a = 0; a < T; a++) for (int b = 0; b < T; b++)
                if (((light >> a) & 1u) && ((light >> b) & 1u)) { const double r = pair_reach(a, b); if (r > rmax) { rmax = r; worst = a; } }
            const int S = rmax > 0 ? (int)floor(edge / (1.02 * rmax)) : 1;
            if (S >= 2) { best = S > 4 ? 4 : S; break; }
            if (worst < 0) break;
            light &= ~(1u << worst);
        }
        // worth it only if the light types are the bulk of the system
        if (best >= 2 && light) {
            long long nlight = 0;
            for (int t = 0; t < T; t++) if ((light >> t) & 1u) nlight += c->type_count[t];
            if (nlight * 4 >= (long long)c->n * 3 && (long long)c->n - nlight <= HV_MAX) { c->sub = best; c->fine_heavy = c->present_mask & ~light; }
        }
    }
    const int s3 = c->sub * c->sub * c->sub;
    const long long nfine = ncells * s3;
    ARG(nfine < (1ll << 30), "scgpu_build_cells: too many sub-cells");
    if (c->ncells + 2 > c->cells_cap) {
        cudaFree(c->d_cell_start);
        c->cells_cap = c->ncells + 2;
        CK(cudaMalloc(&c->d_cell_start, (size_t)c->cells_cap * sizeof(int)));
    }
    if ((int)nfine + 2 > c->fine_cap) {
        cudaFree(c->d_counts); cudaFree(c->d_cursor); cudaFree(c->d_fine_start);
        c->fine_cap = (int)nfine + 2;
        CK(cudaMalloc(&c->d_counts, (size_t)c->fine_cap * sizeof(int)));
        CK(cudaMemsetAsync(c->d_counts, 0, (size_t)c->fine_cap * sizeof(int), c->stream));
        CK(cudaMalloc(&c->d_cursor, (size_t)c->fine_cap * sizeof(int)));
        CK(cudaMalloc(&c->d_fine_start, (size_t)c->fine_cap * sizeof(int)));
    }
    if (c->sub > 1 && !c->d_fine_of) CK(cudaMalloc(&c->d_fine_of, (size_t)c->cap * sizeof(int)));
    DevSys s = view(c);
    int nb = (c->n + 255) / 256;
    // (the histogram array is zero here: cleared when allocated and again by every scan)
    const int nb1 = (c->n + 1023) / 1024;
    if (c->sub > 1) {
        k_cell_count<<<nb1, 1024, 0, c->stream>>>(s, (int)nfine, c->d_cell_of, c->d_fine_of, c->d_counts, c->d_fine_start, c->d_cursor, c->d_ticket);
        k_coarse_start<<<(c->ncells + 2 + 255) / 256, 256, 0, c->stream>>>(c->ncells, s3, c->d_fine_start, c->d_cell_start);
        k_cell_fill<<<nb, 256, 0, c->stream>>>(c->n, c->d_fine_of, c->d_cursor, c->d_tmp);
        k_cell_place<<<nb, 256, 0, c->stream>>>(s, c->d_fine_of, c->d_tmp, c->d_order, c->d_slot_of, c->d_posw, c->d_rec, c->d_p32, c->d_d32);
        // the heavy-type particles, as a sorted list of slots (a few hundred among hundreds of thousands)
        if (!c->d_nheavy) CK(cudaMalloc(&c->d_nheavy, 4 * sizeof(int)));
        if (c->heavy_cap < HV_MAX) { cudaFree(c->d_heavy); CK(cudaMalloc(&c->d_heavy, HV_MAX * sizeof(int))); c->heavy_cap = HV_MAX; }
        CK(cudaMemsetAsync(c->d_nheavy, 0, 4 * sizeof(int), c->stream));
        k_heavy_collect<<<nb, 256, 0, c->stream>>>(c->n, c->d_posw, c->fine_heavy, c->d_heavy, c->d_nheavy);
        k_heavy_sort<<<1, 1024, 0, c->stream>>>(c->d_heavy, c->d_nheavy);
        c->launches += 3;
    } else {
        k_cell_count<<<nb1, 1024, 0, c->stream>>>(s, c->ncells, c->d_cell_of, c->d_fine_of, c->d_counts, c->d_cell_start, c->d_cursor, c->d_ticket);
        k_cell_fill<<<nb, 256, 0, c->stream>>>(c->n, c->d_cell_of, c->d_cursor, c->d_tmp);
        k_cell_place<<<nb, 256, 0, c->stream>>>(s, c->d_cell_of, c->d_tmp, c->d_order, c->d_slot_of, c->d_posw, c->d_rec, c->d_p32, c->d_d32);
    }
    c->f32_valid = true;
    c->launches += 3;
    CK(cudaGetLastError());
    c->cells_valid = true;
    c->h_cell_of.clear();   // host mirror fetched lazily by scgpu_update_particle
    return SCGPU_OK;
}

extern "C" int scgpu_build_cells(scgpu_ctx* c) {
    const double zero[3] = {0.0, 0.0, 0.0};
    return build_cells_impl(c, zero, 0);
}

static int ensure_cells(scgpu_ctx* c) {
    if (c->cells_valid) return 0;
    return scgpu_build_cells(c);
}

extern "C" int scgpu_cell_assignment(scgpu_ctx* c, int* cell_of_particle, int ncell3[3]) {
    ARG(c && cell_of_particle && ncell3, "scgpu_cell_assignment: NULL argument");
    if (int r = ensure_cells(c)) return r;
    CK(cudaMemcpyAsync(cell_of_particle, c->d_cell_of, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; d++) ncell3[d] = c->nc[d];
    return SCGPU_OK;
}

extern "C" int scgpu_cell_order(scgpu_ctx* c, int* order, int* cell_start) {
    ARG(c && order && cell_start, "scgpu_cell_order: NULL argument");
    if (int r = ensure_cells(c)) return r;
    CK(cudaMemcpyAsync(order, c->d_order, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cell_start, c->d_cell_start, (size_t)(c->ncells + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // definition C1: ascending original index inside a cell. With sub-cells the internal order is (sub-cell, index): re-sort per cell
    if (c->sub > 1) for (int k = 0; k < c->ncells; k++) std::sort(order + cell_start[k], order + cell_start[k + 1]);
    return SCGPU_OK;
}

extern "C" int scgpu_update_particle(scgpu_ctx* c, int idx, const double* state30) {
    ARG(c && state30, "scgpu_update_particle: NULL argument");
    ARG(idx >= 0 && idx < c->n, "scgpu_update_particle: index out of range");
    CK(cudaSetDevice(c->device));
    if (sync_api_from_sorted(c)) return SCGPU_ERR_CUDA;
    CK(cudaStreamSynchronize(c->stream));        // the staging slot is reused: the previous update must have landed
    memcpy(c->h_small + 512, state30, 30 * sizeof(double));
    CK(cudaMemcpyAsync(c->d_api + (size_t)idx * 30, c->h_small + 512, 30 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->cells_valid) {
        if (c->h_cell_of.empty()) {          // (sub-cells: the mirror holds the sub-cell index -- a particle must stay in its sub-cell to be updated in place)
            c->h_cell_of.resize(c->n);
            CK(cudaMemcpyAsync(c->h_cell_of.data(), c->sub > 1 ? c->d_fine_of : c->d_cell_of, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        int newc = c->sub > 1 ? fine_index(state30, c->shift, c->nc, c->sub, nullptr) : cell_index(state30, c->shift, c->nc);
        if (newc != c->h_cell_of[idx]) {
            c->cells_valid = false;      // left its cell: the next energy call re-sorts
        } else {
            DevSys s = view(c);
            k_update_one<<<1, 32, 0, c->stream>>>(s, idx, c->d_posw, c->d_rec, c->d_p32, c->d_d32);
            c->launches++;
            CK(cudaGetLastError());
        }
    }
    return SCGPU_OK;
}

// MoveCreator::switchTypeMove (scOOP/mc/movecreator.cpp:233-303) changes conf->pvec[target].type in place (and re-derives the patch
// vectors with Particle::init for the new type) before it asks for oneToAllTrial(target), and keeps the new type when the move is
// accepted. The type of a particle lives on the device (uploaded by scgpu_set_particles): this call changes it for ONE particle.
// The cell-sorted arrays carry the type of every partner, and the kernels' specialisations depend on which types are present: both
// are refreshed (the next energy call re-sorts). Type switches are rare (switchprob per sweep), the re-sort is not on a hot path.
extern "C" int scgpu_set_particle_type(scgpu_ctx* c, int idx, int type) {
    ARG(c, "scgpu_set_particle_type: NULL context");
    ARG(c->types_valid && idx >= 0 && idx < c->n, "scgpu_set_particle_type: index out of range (or no particles with types uploaded)");
    ARG(type >= 0 && type < c->ntypes, "scgpu_set_particle_type: not a particle type of the topology");
    CK(cudaSetDevice(c->device));
    if (sync_api_from_sorted(c)) return SCGPU_ERR_CUDA;
    CK(cudaStreamSynchronize(c->stream));        // the staging slot is reused
    int* h = (int*)(c->h_small + 768);
    CK(cudaMemcpyAsync(h, c->d_type + idx, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int old = h[0];
    if (old == type) return SCGPU_OK;
    h[1] = type;
    CK(cudaMemcpyAsync(c->d_type + idx, h + 1, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->type_count[old]--;
    c->type_count[type]++;
    derive_type_flags(c);
    c->cells_valid = false;
    c->f32_valid = false;
    return SCGPU_OK;
}

extern "C" int scgpu_download_particles(scgpu_ctx* c, double* state30) {
    ARG(c && state30, "scgpu_download_particles: NULL argument");
    CK(cudaSetDevice(c->device));
    if (sync_api_from_sorted(c)) return SCGPU_ERR_CUDA;
    CK(cudaMemcpyAsync(state30, c->d_api, (size_t)c->n * 30 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SCGPU_OK;
}

static int ensure_trial(scgpu_ctx* c, int m) {
    if (m <= c->trial_cap) return 0;
    cudaFree(c->d_trial); cudaFree(c->d_targets); cudaFree(c->d_trial_rec);
    c->d_trial = nullptr; c->d_targets = nullptr; c->d_trial_rec = nullptr; c->trial_cap = 0;
    CK(cudaMalloc(&c->d_trial_rec, (size_t)m * REC * sizeof(double)));
    CK(cudaMalloc(&c->d_trial, (size_t)m * 30 * sizeof(double)));
    CK(cudaMalloc(&c->d_targets, (size_t)m * sizeof(int)));
    c->trial_cap = m;
    return 0;
}

// the three launches behind every energy entry point
static int launch_energy(scgpu_ctx* c, int mode, int m, int gw, const int* d_targets, const double* d_trial, int excl_lo, int excl_hi,
                         double* d_out, double* d_pairs, unsigned long long* d_counters, cudaEvent_t* stage_ev = nullptr) {
    // stage_ev (measurement only): 5 events recorded before / between / after the launches of the flat pipeline
    DevSys s = view(c);
    PatchList pl;
    pl.pair = c->d_pl_pair; pl.e = c->d_pl_e; pl.total = c->d_pl_total; pl.cap = c->pl_cap;
    pl.warp_head = c->d_warp_head; pl.chunks = c->d_chunks; pl.chunk_count = c->d_pl_total + 2; pl.chunk_cap = c->chunk_cap;
    pl.trial_rec = c->d_trial_rec; pl.overflow = c->d_pl_overflow;
    int groups_per_block = OTA_WARPS / gw;
    int blocks = (m + groups_per_block - 1) / groups_per_block;
    if (mode == 1 || mode == 2) {        // every-particle passes: the flat four-launch pipeline
        FlatList fl;
        fl.pair = c->d_fl_pair; fl.e = c->d_fl_e; fl.total = c->d_pl_total + 3; fl.cap = c->fl_cap; fl.head = c->d_fl_head;
        fl.chunks = c->d_fl_chunks; fl.chunk_count = c->d_pl_total + 4; fl.chunk_cap = c->fl_chunk_cap;
        fl.plist = c->d_fl_plist; fl.ptotal = c->d_pl_total + 5; fl.overflow = c->d_pl_overflow;
        fl.heavy = c->d_pl_total + 6; fl.heavy_types = 0;
        const bool wrap = c->nc[0] < 5 || c->nc[1] < 5 || c->nc[2] < 5;
        const bool one = c->one_type >= 0 && c->rods_only;
        const scgpu_iaparam& ia1 = c->h_ia[one ? (size_t)c->one_type * c->ntypes + c->one_type : 0];
        const int nrows = c->nc[1] * c->nc[2];
        // The row-unit gates test |q|^2 - 2 t.q <= reach^2 - |t|^2 in FP32 coordinates relative to the unit centre: the rounding
        // error of that form grows with R^2 (R = half-span of a unit's neighbourhood: up to (kmax + 2) / 2 cells along x, 1.5 along
        // y and z), not with the reach. The reach carries a 0.1 % margin; where ~4 roundings of relative size 2^-24 on R^2 would
        // eat more than half of it (long rods set the cells, short-reach beads are the pair) the cell gate is used instead -- its
        // test is a plain |t - q|^2, whose error is relative to the distance itself.
        bool rows_exact = true;
        if (!wrap) {
            const double bmax = fmax(c->box[0], fmax(c->box[1], c->box[2]));
            double hx, hy, hz;
            if (c->sub > 1 && !c->rods_only) {      // k_gate_fine: coordinates relative to the cell centre, candidates at most one sub-cell outside the cell
                const double f = 0.5 + 1.0 / c->sub;
                hx = f * c->box[0] / c->nc[0]; hy = f * c->box[1] / c->nc[1]; hz = f * c->box[2] / c->nc[2];
            } else {
                const double kx = 0.5 * (double)((c->nc[0] - 3 < 6 ? c->nc[0] - 3 : 6) + 2);
                hx = kx * c->box[0] / c->nc[0]; hy = 1.5 * c->box[1] / c->nc[1]; hz = 1.5 * c->box[2] / c->nc[2];
            }
            const double R2 = hx * hx + hy * hy + hz * hz;
            // + the staged positions are FP32 box fractions: 2^-24 of the box per coordinate, entering d^2 as 2 sqrt(reach2) sqrt(3) times that
            rows_exact = R2 * 4.0 * 5.96e-8 + 2.0 * sqrt(c->min_reach2) * 1.74 * bmax * 1.2e-7 <= 0.5 * 0.001 * c->min_reach2;
        }
        const bool rows_ok = rows_exact && c->use_rows && !wrap && !d_counters && nrows <= 65535;
        if (!c->f32_valid) {
            k_make_f32<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->n, c->d_posw, c->d_rec, c->d_p32, c->d_d32);
            c->launches++;
            c->f32_valid = true;
        }
        auto mark = [&](int k) { if (stage_ev) cudaEventRecord(stage_ev[k], c->stream); };
        auto launch_cheap_rods = [&]() {        // MIRROR (every-particle passes): a patch pair is evaluated by one of its two sides only
            const int nb = c->sm_count * CHEAP_MINB * 2;
            if (one) { if (mode == 1) k_cheap_flat<true, true, true><<<nb, 256, 0, c->stream>>>(s, fl, d_counters, ia1); else k_cheap_flat<true, true, false><<<nb, 256, 0, c->stream>>>(s, fl, d_counters, ia1); }
            else { if (mode == 1) k_cheap_flat<true, false, true><<<nb, 256, 0, c->stream>>>(s, fl, d_counters, ia1); else k_cheap_flat<true, false, false><<<nb, 256, 0, c->stream>>>(s, fl, d_counters, ia1); }
        };
        mark(0);
        if (c->rods_only && rows_ok) {
            const dim3 grid((unsigned)(c->n / nrows / GR_T + 2), (unsigned)nrows);
            if (mode == 1) k_gate_rows<1><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl);
            else k_gate_rows<2><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl);
            mark(1);
            launch_cheap_rods();
        } else if (!c->rods_only && rows_ok && c->sub > 1 && c->ncells <= 65535) {
            // sub-cell gate for the light types, cell gate for the targets of the heavy ones (and of any type whose targets overflowed a buffer)
            const dim3 grid(4, (unsigned)c->ncells);
            const unsigned skip = c->fine_heavy | c->heavy_types;
            if (mode == 1) k_gate_fine<1><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl, c->d_heavy, c->d_nheavy, c->fine_heavy, skip);
            else k_gate_fine<2><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl, c->d_heavy, c->d_nheavy, c->fine_heavy, skip);
            if (c->fine_heavy) {        // the heavy particles themselves: a block each
                int nheavy = 0;
                for (int t = 0; t < c->ntypes && t < 32; t++) if ((c->fine_heavy >> t) & 1u) nheavy += (int)c->type_count[t];
                if (mode == 1) k_gate_heavy<1><<<nheavy, 256, 0, c->stream>>>(s, fl, c->d_heavy, nheavy);
                else k_gate_heavy<2><<<nheavy, 256, 0, c->stream>>>(s, fl, c->d_heavy, nheavy);
                c->launches++;
            }
            if (c->heavy_types & ~c->fine_heavy) {      // types whose targets overflowed k_gate_fine's buffers: the cell gate
                fl.heavy_types = c->heavy_types & ~c->fine_heavy;
                if (mode == 1) k_gate_cells<1, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, nullptr);
                else k_gate_cells<2, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, nullptr);
                c->launches++;
            }
            fl.heavy_types = 0;
            mark(1);
            if (mode == 1) k_cheap_flat<false, false, true><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1); else k_cheap_flat<false, false, false><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1);
        } else if (!c->rods_only && rows_ok && c->ntypes <= GG_MAXT) {
            const dim3 grid((unsigned)(c->n / nrows / GR_T + 2), (unsigned)nrows);
            fl.heavy_types = c->heavy_types;
            if (mode == 1) k_gate_rows_gen<1><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl);
            else k_gate_rows_gen<2><<<grid, GR_SL * 32, 0, c->stream>>>(s, fl);
            if (c->heavy_types) {      // the few targets with very many partners: the cell gate, restricted to their types
                if (mode == 1) k_gate_cells<1, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, nullptr);
                else k_gate_cells<2, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, nullptr);
                c->launches++;
            }
            fl.heavy_types = 0;
            mark(1);
            if (mode == 1) k_cheap_flat<false, false, true><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1); else k_cheap_flat<false, false, false><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1);
        } else if (c->rods_only) {
            if (mode == 1) { if (wrap) k_gate_cells<1, true, true><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); else k_gate_cells<1, true, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); }
            else { if (wrap) k_gate_cells<2, true, true><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); else k_gate_cells<2, true, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); }
            mark(1);
            launch_cheap_rods();
        } else {
            if (mode == 1) { if (wrap) k_gate_cells<1, false, true><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); else k_gate_cells<1, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); }
            else { if (wrap) k_gate_cells<2, false, true><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); else k_gate_cells<2, false, false><<<c->ncells, GT_WARPS * 32, 0, c->stream>>>(s, fl, d_counters); }
            mark(1);
            if (mode == 1) k_cheap_flat<false, false, true><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1); else k_cheap_flat<false, false, false><<<c->sm_count * 4, 256, 0, c->stream>>>(s, fl, d_counters, ia1);
        }
        mark(2);
        if (one) k_patch_flat<true><<<c->sm_count * PATCH_MINB, PF_THREADS, 0, c->stream>>>(s, fl, c->any_two_patch ? 1 : 0, mode == 1 ? 1 : 0, ia1); else k_patch_flat<false><<<c->sm_count * PATCH_MINB, PF_THREADS, 0, c->stream>>>(s, fl, c->any_two_patch ? 1 : 0, mode == 1 ? 1 : 0, ia1);
        mark(3);
        k_combine_flat<<<(c->n + 31) / 32, 256, 0, c->stream>>>(c->n, fl, c->d_posw, d_out);
        mark(4);
        c->launches += 4;
        if (c->exter) { k_add_exter<<<(c->n + 127) / 128, 128, 0, c->stream>>>(s, mode, c->n, nullptr, nullptr, 0, d_out); c->launches++; }
        CK(cudaGetLastError());
        return 0;
    }
#define LAUNCH_GC(M, R) k_gate_cheap<M, R><<<blocks, OTA_THREADS, 0, c->stream>>>(s, m, gw, d_targets, d_trial, excl_lo, excl_hi, pl, c->d_warp_partial, d_pairs, d_counters)
    if (c->rods_only) {
        switch (mode) { case 0: LAUNCH_GC(0, true); break; case 1: LAUNCH_GC(1, true); break; case 2: LAUNCH_GC(2, true); break; default: LAUNCH_GC(3, true); break; }
    } else {
        switch (mode) { case 0: LAUNCH_GC(0, false); break; case 1: LAUNCH_GC(1, false); break; case 2: LAUNCH_GC(2, false); break; default: LAUNCH_GC(3, false); break; }
    }
#undef LAUNCH_GC
    int pblocks = m <= 64 ? 8 : c->sm_count * 8;
    k_patch<<<pblocks, 128, 0, c->stream>>>(s, pl, d_targets, mode, excl_lo, d_pairs);
    k_combine<<<(m + 255) / 256, 256, 0, c->stream>>>(m, gw, pl, c->d_warp_partial, d_out);
    c->launches += 3;
    if (c->exter) { k_add_exter<<<(m + 127) / 128, 128, 0, c->stream>>>(s, mode, m, d_targets, d_trial, excl_lo, d_out); c->launches++; }
    CK(cudaGetLastError());
    return 0;
}

// the patch work list is sized for ~24 patch partners per particle on average; if a launch overflowed it (sticky device
// flag) the results of that launch are invalid: grow the list 4x and let the caller repeat the launch
static int overflow_then_grow(scgpu_ctx* c, bool* repeat) {
    int* hflag = (int*)(c->h_small + 256 + 64);
    CK(cudaMemcpyAsync(hflag, c->d_pl_overflow, 6 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));      // [0] flags ... [5] heavy-type mask
    CK(cudaStreamSynchronize(c->stream));
    *repeat = false;
    if (!*hflag) return 0;
    const int which = *hflag;      // bit 0: patch work list, bit 1: flat pair list, bit 2: k_gate_rows cannot hold this configuration
    c->last_overflow = which;
    if (which & 8) {
        CK(cudaMemsetAsync(c->d_pl_total, 0, 8 * sizeof(int), c->stream));
        g_err = "internal error: a patch pair was listed by one side of an every-particle pass only";
        return SCGPU_ERR_STATE;
    }
    if (which & (16 | 32 | 64)) {      // raised by the phased sweep (sweep_phased.cuh): its passes stop at the first cell or trial they cannot hold
        CK(cudaMemsetAsync(c->d_pl_total, 0, 8 * sizeof(int), c->stream));
        g_err = "a batched sweep could not be completed by the phased sweep kernels (a cell above " + std::to_string(SP_TR) + " particles, a neighbourhood above " +
                std::to_string(SP_TILE) + " candidates or a trial with more than " + std::to_string(SP_EB) + " pair terms; flags " + std::to_string(which) +
                "): the remaining passes were not performed; set SCGPU_SWEEP_KERNEL=rounds for this system";
        return SCGPU_ERR_STATE;
    }
    if (which & 4) {
        const unsigned heavy = (unsigned)hflag[5];       // d_pl_total[6]
        unsigned present = 0;
        for (int t = 0; t < c->ntypes && t < 32; t++) present |= 1u << t;
        if (!c->rods_only && heavy && ((c->heavy_types | heavy) & present) != present) c->heavy_types |= heavy;     // split those types off and try again
        else c->use_rows = false;
    }
    if (which & 2) {
        if ((long long)c->fl_cap * 2 > (1ll << 30)) { g_err = "flat pair list overflow: more gated pairs than the list can ever hold"; return SCGPU_ERR_STATE; }
        cudaFree(c->d_fl_pair); cudaFree(c->d_fl_e); cudaFree(c->d_fl_plist); cudaFree(c->d_fl_chunks);
        c->d_fl_pair = nullptr; c->d_fl_e = nullptr; c->d_fl_plist = nullptr; c->d_fl_chunks = nullptr;
        c->fl_cap *= 2;
        c->fl_chunk_cap = c->cap + 64 + c->fl_cap / 64 + 64;
        CK(cudaMalloc(&c->d_fl_pair, (size_t)c->fl_cap * sizeof(int2)));
        CK(cudaMalloc(&c->d_fl_e, (size_t)c->fl_cap * sizeof(double2)));
        CK(cudaMalloc(&c->d_fl_plist, (size_t)c->fl_cap * sizeof(int)));
        CK(cudaMalloc(&c->d_fl_chunks, (size_t)c->fl_chunk_cap * sizeof(int4)));
    }
    if (which & 1) {
        if ((long long)c->pl_cap * 4 > (1ll << 30)) { g_err = "patch work list overflow: more patch pairs than the list can ever hold"; return SCGPU_ERR_STATE; }
        cudaFree(c->d_pl_pair); cudaFree(c->d_pl_e); cudaFree(c->d_chunks);
        c->d_pl_pair = nullptr; c->d_pl_e = nullptr; c->d_chunks = nullptr;
        c->pl_cap *= 4;
        c->chunk_cap = c->cap + 64 + c->pl_cap / PCAP + 64;
        CK(cudaMalloc(&c->d_pl_pair, (size_t)c->pl_cap * sizeof(int2)));
        CK(cudaMalloc(&c->d_pl_e, (size_t)c->pl_cap * sizeof(double)));
        CK(cudaMalloc(&c->d_chunks, (size_t)c->chunk_cap * sizeof(int4)));
    }
    CK(cudaMemsetAsync(c->d_pl_total, 0, 8 * sizeof(int), c->stream));
    *repeat = true;
    return 0;
}

extern "C" int scgpu_one_to_all(scgpu_ctx* c, int target, const double* trial_state30, double* e_sum, double* e_pairs) {
    ARG(c && e_sum, "scgpu_one_to_all: NULL argument");
    ARG(target >= 0 && target < c->n, "scgpu_one_to_all: target out of range");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    if (ensure_trial(c, 1)) return SCGPU_ERR_CUDA;
    // the sequential-MC path calls this twice per trial move: one small pinned upload (trial record + target index), three
    // launches, one small pinned read-back, one synchronisation
    if (trial_state30) memcpy(c->h_small, trial_state30, 30 * sizeof(double));
    memcpy(c->h_small + 240, &target, sizeof(int));
    double* hres = (double*)(c->h_small + 256);
    for (bool repeat = true; repeat;) {
        CK(cudaMemcpyAsync(c->d_single, c->h_small, 256, cudaMemcpyHostToDevice, c->stream));
        if (e_pairs) CK(cudaMemsetAsync(c->d_pairs, 0, (size_t)c->n * sizeof(double), c->stream));
        // a single trial: all 4 warps of one block share the 27 neighbour cells
        if (launch_energy(c, 0, 1, OTA_WARPS, (const int*)(c->d_single + 240), trial_state30 ? (const double*)c->d_single : nullptr, 0, 0, c->d_out,
                          e_pairs ? c->d_pairs : nullptr, nullptr)) return SCGPU_ERR_CUDA;
        CK(cudaMemcpyAsync(hres, c->d_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (e_pairs) CK(cudaMemcpyAsync(e_pairs, c->d_pairs, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (int r = overflow_then_grow(c, &repeat)) return r;
    }
    *e_sum = *hres;
    return SCGPU_OK;
}

extern "C" int scgpu_one_to_all_batch(scgpu_ctx* c, int m, const int* targets, const double* trial_states30, double* e_sums) {
    ARG(c && targets && e_sums, "scgpu_one_to_all_batch: NULL argument");
    ARG(m > 0, "scgpu_one_to_all_batch: m must be positive");
    for (int i = 0; i < m; i++) ARG(targets[i] >= 0 && targets[i] < c->n, "scgpu_one_to_all_batch: target out of range");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    if (ensure_trial(c, m)) return SCGPU_ERR_CUDA;
    size_t tb = (size_t)m * sizeof(int), sb = trial_states30 ? (size_t)m * 30 * sizeof(double) : 0, ob = (size_t)m * sizeof(double);
    if (ensure_pinned(c, tb + sb + ob)) return SCGPU_ERR_CUDA;
    char* pin = (char*)c->h_pinned;
    memcpy(pin, targets, tb);
    if (sb) memcpy(pin + tb, trial_states30, sb);
    CK(cudaMemcpyAsync(c->d_targets, pin, tb, cudaMemcpyHostToDevice, c->stream));
    if (sb) CK(cudaMemcpyAsync(c->d_trial, pin + tb, sb, cudaMemcpyHostToDevice, c->stream));
    double* d_res = c->d_out;
    if (m > c->n) { g_err = "scgpu_one_to_all_batch: m larger than the particle count is not supported"; return SCGPU_ERR_ARG; }
    for (bool repeat = true; repeat;) {
        if (launch_energy(c, 0, m, 1, c->d_targets, sb ? c->d_trial : nullptr, 0, 0, d_res, nullptr, nullptr)) return SCGPU_ERR_CUDA;
        CK(cudaMemcpyAsync(pin + tb + sb, d_res, ob, cudaMemcpyDeviceToHost, c->stream));
        if (int r = overflow_then_grow(c, &repeat)) return r;
    }
    memcpy(e_sums, pin + tb + sb, ob);
    return SCGPU_OK;
}

extern "C" int scgpu_one_to_all_everyone(scgpu_ctx* c, double* e_host, int64_t* n_candidates, int64_t* n_gated) {
    ARG(c, "scgpu_one_to_all_everyone: NULL context");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    bool count = n_candidates || n_gated;
    unsigned long long h[2] = {0, 0};
    for (bool repeat = true; repeat;) {
        if (count) CK(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), c->stream));
        if (launch_energy(c, 1, c->n, 1, nullptr, nullptr, 0, 0, c->d_out, nullptr, count ? c->d_counters : nullptr)) return SCGPU_ERR_CUDA;
        if (e_host) CK(cudaMemcpyAsync(e_host, c->d_out, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (count) CK(cudaMemcpyAsync(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        repeat = false;
        if (e_host || count) { if (int r = overflow_then_grow(c, &repeat)) return r; }   // asynchronous calls are checked by scgpu_sync
    }
    if (n_candidates) *n_candidates = (int64_t)h[0];
    if (n_gated) *n_gated = (int64_t)h[1];
    return SCGPU_OK;
}

// measurement: one every-particle pass with an event between its launches -> microseconds of gate, cheap, patch, combine
extern "C" int scgpu_profile_everyone(scgpu_ctx* c, float* us4) {
    ARG(c && us4, "scgpu_profile_everyone: NULL argument");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    cudaEvent_t ev[5];
    for (int k = 0; k < 5; k++) CK(cudaEventCreate(&ev[k]));
    for (int pass = 0; pass < 4; pass++)       // the first passes bring the clocks up after an idle period; the last one is reported
        for (bool repeat = true; repeat;) {
            if (launch_energy(c, 1, c->n, 1, nullptr, nullptr, 0, 0, c->d_out, nullptr, nullptr, ev)) return SCGPU_ERR_CUDA;
            if (int r = overflow_then_grow(c, &repeat)) return r;
        }
    for (int k = 0; k < 4; k++) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); us4[k] = ms * 1000.f; }
    for (int k = 0; k < 5; k++) cudaEventDestroy(ev[k]);
    return SCGPU_OK;
}

extern "C" int scgpu_submit_everyone(scgpu_ctx* c, const double* state9, double* e_out) {
    ARG(c && state9 && e_out, "scgpu_submit_everyone: NULL argument");
    ARG(c->n > 0 && c->types_valid, "scgpu_submit_everyone: needs a previous upload that brought the particle types");
    ARG(is_pinned_host(state9) && is_pinned_host(e_out), "scgpu_submit_everyone: both buffers must be page-locked host memory");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->d_compact, state9, (size_t)c->n * 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_particle_init<<<(c->n + 127) / 128, 128, 0, c->stream>>>(c->n, c->ntypes, c->d_compact, c->d_type, c->d_ia, c->d_api);
    c->launches++;
    CK(cudaGetLastError());
    c->cells_valid = false;
    c->api_stale = false;
    if (int r = scgpu_build_cells(c)) return r;
    if (launch_energy(c, 1, c->n, 1, nullptr, nullptr, 0, 0, c->d_out, nullptr, nullptr)) return SCGPU_ERR_CUDA;
    CK(cudaMemcpyAsync(e_out, c->d_out, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return SCGPU_OK;          // list overflows are reported by scgpu_sync
}

extern "C" int scgpu_mol_to_others(scgpu_ctx* c, int first, int m, const double* trial_states30, double* e_sum) {
    ARG(c && e_sum, "scgpu_mol_to_others: NULL argument");
    ARG(first >= 0 && m > 0 && first + m <= c->n, "scgpu_mol_to_others: molecule range out of bounds");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    if (trial_states30) {
        if (ensure_trial(c, m)) return SCGPU_ERR_CUDA;
        CK(cudaMemcpyAsync(c->d_trial, trial_states30, (size_t)m * 30 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    for (bool repeat = true; repeat;) {
        if (launch_energy(c, 3, m, 1, nullptr, trial_states30 ? c->d_trial : nullptr, first, first + m, c->d_out, nullptr, nullptr)) return SCGPU_ERR_CUDA;
        k_reduce_fixed<<<RF_BLOCKS, RF_THREADS, 0, c->stream>>>(m, c->d_out, c->d_scalar, c->d_reduce);
        c->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(e_sum, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (int r = overflow_then_grow(c, &repeat)) return r;
    }
    return SCGPU_OK;
}

static int launch_all_to_all(scgpu_ctx* c) {
    if (launch_energy(c, 2, c->n, 1, nullptr, nullptr, 0, 0, c->d_out, nullptr, nullptr)) return SCGPU_ERR_CUDA;
    k_reduce_fixed<<<RF_BLOCKS, RF_THREADS, 0, c->stream>>>(c->n, c->d_out, c->d_scalar, c->d_reduce);
    c->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int scgpu_all_to_all(scgpu_ctx* c, double* e_total, double* e_per_particle) {
    ARG(c, "scgpu_all_to_all: NULL argument");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    for (bool repeat = true; repeat;) {
        if (launch_all_to_all(c)) return SCGPU_ERR_CUDA;
        if (e_total) CK(cudaMemcpyAsync(e_total, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (e_per_particle) CK(cudaMemcpyAsync(e_per_particle, c->d_out, (size_t)c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        repeat = false;
        if (e_total || e_per_particle) { if (int r = overflow_then_grow(c, &repeat)) return r; }
    }
    return SCGPU_OK;
}

extern "C" int scgpu_overlap_one(scgpu_ctx* c, int target, const double* trial_state30, int variant, int* flag) {
    ARG(c && flag, "scgpu_overlap_one: NULL argument");
    ARG(target >= 0 && target < c->n, "scgpu_overlap_one: target out of range");
    ARG(variant == 0 || variant == 1, "scgpu_overlap_one: variant must be 0 or 1");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    if (ensure_trial(c, 1)) return SCGPU_ERR_CUDA;
    CK(cudaMemcpyAsync(c->d_targets, &target, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (trial_state30) CK(cudaMemcpyAsync(c->d_trial, trial_state30, 30 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    DevSys s = view(c);
    k_overlap<<<1, OTA_THREADS, 0, c->stream>>>(s, 1, c->d_targets, trial_state30 ? c->d_trial : nullptr, 0, variant, c->d_flags);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flag, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SCGPU_OK;
}

extern "C" int scgpu_overlap_all(scgpu_ctx* c, int variant, int* flag) {
    ARG(c && flag, "scgpu_overlap_all: NULL argument");
    ARG(variant == 0 || variant == 1, "scgpu_overlap_all: variant must be 0 or 1");
    CK(cudaSetDevice(c->device));
    if (int r = ensure_cells(c)) return r;
    CK(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
    DevSys s = view(c);
    k_overlap<<<(c->n + OTA_WARPS - 1) / OTA_WARPS, OTA_THREADS, 0, c->stream>>>(s, c->n, nullptr, nullptr, 1, variant, c->d_flags);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flag, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SCGPU_OK;
}

static inline unsigned long long splitmix64(unsigned long long& x) {
    unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int sweep_impl(scgpu_ctx* c, const scgpu_moveparams* mp, const scgpu_chainmoves* cm, uint64_t seed, uint64_t sweep, scgpu_sweepstats* stats,
                      scgpu_chainstats* cstats, double single_scale = -1.0, bool chains_only = false);
extern "C" int scgpu_sweep_checkerboard(scgpu_ctx* c, const scgpu_moveparams* mp, uint64_t seed, uint64_t sweep, scgpu_sweepstats* stats) {
    return sweep_impl(c, mp, nullptr, seed, sweep, stats, nullptr);
}
extern "C" int scgpu_sweep_checkerboard_chains(scgpu_ctx* c, const scgpu_moveparams* mp, const scgpu_chainmoves* cm, uint64_t seed, uint64_t sweep,
                                               scgpu_sweepstats* stats, scgpu_chainstats* cstats) {
    // Single-particle trials and chain trials of a sweep run as two series of colour passes, each on the grid that suits it: the chain
    // kernel needs whole molecules inside a cell of the COARSE grid, the single-particle walk is several times faster on the fine one
    // (lipid membrane, 265 041 particles: 160 ms against 380 ms per sweep). A composition of moves that each leave the Boltzmann
    // distribution invariant; SCGPU_SWEEP_INTERLEAVED=1 restores the colour-by-colour interleaving on the coarse grid.
    const bool chains = cm && cm->chainprob > 0.0;
    if (!chains || getenv("SCGPU_SWEEP_INTERLEAVED") != nullptr) return sweep_impl(c, mp, cm, seed, sweep, stats, cstats);
    ARG(c && mp, "scgpu_sweep_checkerboard: NULL argument");
    ARG(cm->chainprob <= 1.0, "scgpu_sweep_checkerboard_chains: chainprob must be in [0, 1]");
    if (stats) memset(stats, 0, sizeof *stats);
    if (cm->chainprob < 1.0)
        if (int r = sweep_impl(c, mp, nullptr, seed, sweep, stats, nullptr, 1.0 - cm->chainprob, false)) return r;
    return sweep_impl(c, mp, cm, seed, sweep, nullptr, cstats, 0.0, true);
}

static int ensure_sw_maxc(scgpu_ctx* c) {       // [0]: cell walk, [1]: chain kernel -- largest neighbourhood met in the last sweep
    if (c->d_sw_maxc) return 0;
    CK(cudaMalloc(&c->d_sw_maxc, 2 * sizeof(int)));
    CK(cudaMemset(c->d_sw_maxc, 0, 2 * sizeof(int)));
    CK(cudaMallocHost(&c->h_sw_maxc, 2 * sizeof(int)));
    c->h_sw_maxc[0] = c->h_sw_maxc[1] = 0;
    return 0;
}

static int sweep_impl(scgpu_ctx* c, const scgpu_moveparams* mp, const scgpu_chainmoves* cm, uint64_t seed, uint64_t sweep, scgpu_sweepstats* stats,
                      scgpu_chainstats* cstats, double single_scale, bool chains_only) {
    ARG(c && mp, "scgpu_sweep_checkerboard: NULL argument");
    const bool chains = cm && cm->chainprob > 0.0;
    ARG(!chains || (cm->chainprob <= 1.0 && c->nmol <= CH_MAXMT), "scgpu_sweep_checkerboard_chains: chainprob must be in [0, 1] and at most 32 molecule types");
    if (chains) for (int t = 0; t < c->nmol; t++) ARG(c->h_mol[t].mol_size <= CH_MAX, "scgpu_sweep_checkerboard_chains: a molecule is longer than MAXCHL = 20");
    ARG(mp->temper > 0 && mp->n_sub >= 1, "scgpu_sweep_checkerboard: temperature and n_sub must be positive");
    ARG(mp->grid_k >= 0 && mp->grid_k <= 3, "scgpu_sweep_checkerboard: grid_k must be 0 (automatic), 1, 2 or 3");
    ARG(mp->trial_rule >= 0 && mp->trial_rule <= 2, "scgpu_sweep_checkerboard: trial_rule must be 0 (per cell), 1 (per particle) or 2 (every particle once)");
    ARG(c->n > 0 && c->ntypes > 0 && c->ntypes <= 40, "scgpu_sweep_checkerboard: set topology (<= 40 types) and particles first");
    CK(cudaSetDevice(c->device));
    // random grid shift and colour order for this sweep: a pure function of (seed, sweep)
    unsigned long long st = seed * 0xD1342543DE82EF95ull + sweep * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull;
    double shift[3];
    for (int d = 0; d < 3; d++) shift[d] = (double)(splitmix64(st) >> 11) * (1.0 / 9007199254740992.0);
    // Fineness of the checkerboard: the serial chain of trials inside a cell is what a sweep costs, (K+1)^3 N / cells(K) trials long;
    // the largest K <= 3 that still leaves about two particles per cell. Chain moves need every member of a molecule in one cell and
    // the 27-cell neighbourhood of their kernel: sweeps with chain moves stay on the coarse grid.
    // Systems without bonded molecules take the round kernel (sweep_rounds.cuh): the trials of a cell are evaluated several at a time,
    // so the coarse grid (more trials per cell and pass, fewer passes) is the fast one. SCGPU_SWEEP_KERNEL=cells|rounds forces one.
    bool bonded_any = false;
    for (int t = 0; t < c->nmol; t++) bonded_any = bonded_any || c->h_mol[t].mol_size > 1.0;
    bool rounds = !chains && !bonded_any;
    bool phased_wanted = rounds && mp->trial_rule == 2;      // sweep_phased.cuh: a pass as four dense launches
    if (const char* e = getenv("SCGPU_SWEEP_KERNEL")) {
        if (!strcmp(e, "cells")) { rounds = false; phased_wanted = false; }
        else if (!strcmp(e, "rounds")) phased_wanted = false;
    }
    int K = 1;
    if (!chains && !rounds) {
        for (int k = 3; k >= 2; k--) {
            double cells = 1.0;
            bool fits = true;
            for (int d = 0; d < 3; d++) {
                int nc = (int)floor(c->box[d] * k / c->maxcut);
                nc -= nc % (k + 1);
                if (nc < 2 * (k + 1)) { fits = false; nc = 1; }
                cells *= nc;
            }
            if (fits && (double)c->n / cells >= 2.0) { K = k; break; }
        }
    }
    if (!chains && mp->grid_k >= 1 && mp->grid_k <= 3) K = mp->grid_k;       // the caller's choice (a grid that does not fit falls back to one cell per axis)
#ifdef SW_FORCE_K
    if (!chains) K = SW_FORCE_K;
#endif
    if (int r = build_cells_impl(c, shift, K)) return r;
    SweepGrid grid;
    for (int d = 0; d < 3; d++) { grid.k[d] = c->nc[d] > 1 ? K : 0; grid.ncol[d] = c->nc[d] > 1 ? K + 1 : 1; }
    int3 ncol = make_int3(grid.ncol[0], grid.ncol[1], grid.ncol[2]);
    int ncolours = ncol.x * ncol.y * ncol.z;
    int order[64];
    for (int k = 0; k < ncolours; k++) order[k] = k;
    for (int k = ncolours - 1; k > 0; k--) { int j = (int)(splitmix64(st) % (unsigned long long)(k + 1)); int t = order[k]; order[k] = order[j]; order[j] = t; }
    SweepParams sp;
    sp.temper = mp->temper;
    sp.n_sub = mp->n_sub;
    sp.trial_rule = mp->trial_rule;
    sp.trial_scale = single_scale >= 0.0 ? single_scale : (chains ? 1.0 - cm->chainprob : 1.0);
    ChainParams cp;
    memset(&cp, 0, sizeof cp);
    if (chains) {
        cp.temper = mp->temper;
        cp.trials_per_particle = cm->chainprob * (double)mp->n_sub;
        for (int t = 0; t < CH_MAXMT; t++) { cp.chainm_mx[t] = cm->chainm_mx[t]; cp.chainr_angle[t] = cm->chainr_angle[t]; }
    }
    for (int t = 0; t < 40; t++) {
        sp.trans_mx[t] = mp->trans_mx[t];
        sp.rot_angle[t] = mp->rot_angle[t];
        sp.geotype_of_type[t] = (t < c->ntypes) ? (int)c->h_ia[(size_t)t * c->ntypes + t].geotype[0] : 0;
    }
    if (c->sweep_acc_cap < c->ncells) {        // [0, ncells): single-particle passes, [ncells, 2 ncells): chain passes
        cudaFree(c->d_sweep_acc);
        c->d_sweep_acc = nullptr;
        CK(cudaMalloc(&c->d_sweep_acc, (size_t)2 * c->ncells * sizeof(SweepAcc)));
        c->sweep_acc_cap = c->ncells;
    }
    CK(cudaMemsetAsync(c->d_sweep_acc, 0, (size_t)2 * c->ncells * sizeof(SweepAcc), c->stream));
    SweepAcc* d_chain_acc = (SweepAcc*)c->d_sweep_acc + c->sweep_acc_cap;
    DevSys s = view(c);
    int nactive = (c->nc[0] / ncol.x) * (c->nc[1] / ncol.y) * (c->nc[2] / ncol.z);
    const bool one = c->one_type >= 0 && c->rods_only;
    const scgpu_iaparam& ia1 = c->h_ia[one ? (size_t)c->one_type * c->ntypes + c->one_type : 0];
    const bool phased = phased_wanted && K == 1 && (double)c->n / (double)c->ncells <= 24.0;      // cells of at most SP_TR particles, neighbourhoods inside the tile
    if (phased) {
        static bool pattr_done = false;
        if (!pattr_done) {
            CK(cudaFuncSetAttribute(k_sweep_propose<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpShared)));
            CK(cudaFuncSetAttribute(k_sweep_propose<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpShared)));
            CK(cudaFuncSetAttribute(k_sweep_propose<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpShared)));
            pattr_done = true;
        }
        if (c->sw_trial_cap < c->n + 64) { cudaFree(c->d_sw_trials); c->d_sw_trials = nullptr; CK(cudaMalloc(&c->d_sw_trials, (size_t)(c->n + 64) * sizeof(SwTrial))); c->sw_trial_cap = c->n + 64; }
        if (c->sw_cell_cap < c->ncells) { cudaFree(c->d_sw_cells); c->d_sw_cells = nullptr; CK(cudaMalloc(&c->d_sw_cells, (size_t)c->ncells * sizeof(SwCell))); c->sw_cell_cap = c->ncells; }
        if (c->sw_meta_cap < c->fl_cap) { cudaFree(c->d_sw_meta); c->d_sw_meta = nullptr; CK(cudaMalloc(&c->d_sw_meta, (size_t)c->fl_cap * sizeof(unsigned short))); c->sw_meta_cap = c->fl_cap; }
    }
    if (rounds && !phased) {
        static bool attr_done = false;
        if (!attr_done) {
            CK(cudaFuncSetAttribute(k_sweep_rounds<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SrShared)));
            CK(cudaFuncSetAttribute(k_sweep_rounds<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SrShared)));
            CK(cudaFuncSetAttribute(k_sweep_rounds<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SrShared)));
            attr_done = true;
        }
    }
    // staged tile of the cell walk: the default, or as much as the mean neighbourhood needs (with head room), up to 96 KB
    int sw_tile = SW_TILE;
    if (!rounds && !chains_only) {
        double meanC = (double)c->n / (double)c->ncells;
        for (int d = 0; d < 3; d++) meanC *= (double)(2 * grid.k[d] + 1);
        int want = (int)(meanC * 1.5) + 64;
        if (int r = ensure_sw_maxc(c)) return r;
        // (the mean counts empty cells too -- a membrane in water: what the previous sweep met on the device is the better guide; the copy
        // is asynchronous, a value one sweep old is as good)
        if (*c->h_sw_maxc > 0) want = *c->h_sw_maxc + *c->h_sw_maxc / 16 + 32;
        CK(cudaMemsetAsync(c->d_sw_maxc, 0, sizeof(int), c->stream));
        if (const char* e = getenv("SCGPU_SWEEP_TILE")) want = atoi(e);
        if (want > SW_TILE) {
            sw_tile = want > SW_TILE_MAX ? SW_TILE_MAX : (want + 63) / 64 * 64;
            static bool tattr_done = false;
            if (!tattr_done) {
                CK(cudaFuncSetAttribute(k_sweep_cells<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_TILE_MAX * 2 * (int)sizeof(float4)));
                CK(cudaFuncSetAttribute(k_sweep_cells<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_TILE_MAX * 2 * (int)sizeof(float4)));
                CK(cudaFuncSetAttribute(k_sweep_cells<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_TILE_MAX * 2 * (int)sizeof(float4)));
                tattr_done = true;
            }
        }
    }
    // staged tile of the chain kernel (16 B per candidate of the 27-cell neighbourhood), sized the same way
    int ch_tile = 1024;
    if (chains) {
        if (int r = ensure_sw_maxc(c)) return r;
        int want = (int)(27.0 * (double)c->n / (double)c->ncells * 1.5) + 64;
        if (c->h_sw_maxc[1] > 0) want = c->h_sw_maxc[1] + c->h_sw_maxc[1] / 16 + 32;
        if (const char* e = getenv("SCGPU_CHAIN_TILE")) want = atoi(e);
        ch_tile = want < 1024 ? 1024 : (want > CH_TILE_MAX ? CH_TILE_MAX : (want + 63) / 64 * 64);
        static bool cattr_done = false;
        if (!cattr_done) { CK(cudaFuncSetAttribute(k_sweep_chain_colour, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_TILE_MAX * (int)sizeof(float4))); cattr_done = true; }
        CK(cudaMemsetAsync(c->d_sw_maxc + 1, 0, sizeof(int), c->stream));
    }
    FlatList sfl;
    SweepAux sax;
    memset(&sfl, 0, sizeof sfl);
    memset(&sax, 0, sizeof sax);
    if (phased) {
        sfl.pair = c->d_fl_pair; sfl.e = c->d_fl_e; sfl.total = c->d_pl_total + 3; sfl.cap = c->fl_cap; sfl.head = c->d_fl_head;
        sfl.chunks = c->d_fl_chunks; sfl.chunk_count = c->d_pl_total + 4; sfl.chunk_cap = c->fl_chunk_cap;
        sfl.plist = c->d_fl_plist; sfl.ptotal = c->d_pl_total + 5; sfl.overflow = c->d_pl_overflow;
        sfl.heavy = c->d_pl_total + 6; sfl.heavy_types = 0;
        sax.trials = c->d_sw_trials; sax.trial_total = c->d_pl_total + 7; sax.trial_cap = c->n; sax.cells = c->d_sw_cells; sax.meta = c->d_sw_meta;
    }
    const int nwalks = phased ? mp->n_sub : 1;          // the phased form walks the system n_sub times on the same grid, colour by colour
    for (int sub = 0; sub < nwalks; sub++)
    for (int k = 0; k < ncolours; k++) {
        if (phased) {
            SweepAcc* ao = (SweepAcc*)c->d_sweep_acc;
            if (c->rods_only && one) k_sweep_propose<true, true><<<nactive, SP_THREADS, sizeof(SpShared), c->stream>>>(s, sp, seed, sweep, order[k], sub, grid, c->d_posw, c->d_rec, sfl, sax, ao, ia1);
            else if (c->rods_only) k_sweep_propose<true, false><<<nactive, SP_THREADS, sizeof(SpShared), c->stream>>>(s, sp, seed, sweep, order[k], sub, grid, c->d_posw, c->d_rec, sfl, sax, ao, ia1);
            else k_sweep_propose<false, false><<<nactive, SP_THREADS, sizeof(SpShared), c->stream>>>(s, sp, seed, sweep, order[k], sub, grid, c->d_posw, c->d_rec, sfl, sax, ao, ia1);
            if (c->rods_only) {
                const int nb = c->sm_count * CHEAP_MINB * 2;
                if (one) k_cheap_flat<true, true, false><<<nb, 256, 0, c->stream>>>(s, sfl, nullptr, ia1);
                else k_cheap_flat<true, false, false><<<nb, 256, 0, c->stream>>>(s, sfl, nullptr, ia1);
            } else k_cheap_flat<false, false, false><<<c->sm_count * 4, 256, 0, c->stream>>>(s, sfl, nullptr, ia1);
            if (one) k_patch_flat<true><<<c->sm_count * PATCH_MINB, PF_THREADS, 0, c->stream>>>(s, sfl, c->any_two_patch ? 1 : 0, 0, ia1);
            else k_patch_flat<false><<<c->sm_count * PATCH_MINB, PF_THREADS, 0, c->stream>>>(s, sfl, c->any_two_patch ? 1 : 0, 0, ia1);
            k_sweep_resolve<<<nactive, SP_RTHREADS, 0, c->stream>>>(s, sp, order[k], sub, grid, c->d_posw, c->d_rec, sfl, sax, ao);
            c->launches += 4;
            continue;
        }
        if (rounds) {
            if (c->rods_only && one) k_sweep_rounds<true, true><<<nactive, SR_THREADS, sizeof(SrShared), c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, ia1);
            else if (c->rods_only) k_sweep_rounds<true, false><<<nactive, SR_THREADS, sizeof(SrShared), c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, ia1);
            else k_sweep_rounds<false, false><<<nactive, SR_THREADS, sizeof(SrShared), c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, ia1);
            c->launches++;
            continue;
        }
        if (!chains_only) {
            const size_t dyn = (size_t)sw_tile * 2 * sizeof(float4);
            if (c->rods_only && one) k_sweep_cells<true, true><<<nactive, 32, dyn, c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, sw_tile, c->d_sw_maxc, ia1);
            else if (c->rods_only) k_sweep_cells<true, false><<<nactive, 32, dyn, c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, sw_tile, c->d_sw_maxc, ia1);
            else k_sweep_cells<false, false><<<nactive, 32, dyn, c->stream>>>(s, sp, seed, sweep, order[k], grid, c->d_posw, c->d_rec, (SweepAcc*)c->d_sweep_acc, sw_tile, c->d_sw_maxc, ia1);
            c->launches++;
        }
        if (chains) {
            k_sweep_chain_colour<<<nactive, CH_THREADS, (size_t)ch_tile * sizeof(float4), c->stream>>>(s, cp, seed, sweep, order[k], ncol, c->d_posw, c->d_rec, d_chain_acc, ch_tile, c->d_sw_maxc + 1);
            c->launches++;
        }
    }
    if (!rounds && !chains_only && c->d_sw_maxc) CK(cudaMemcpyAsync(c->h_sw_maxc, c->d_sw_maxc, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (chains && c->d_sw_maxc) CK(cudaMemcpyAsync(c->h_sw_maxc + 1, c->d_sw_maxc + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaGetLastError());
    c->api_stale = true;           // the cell-sorted arrays are now the newest copy of the configuration
    c->f32_valid = false;          // ... and their FP32 copies are out of date
    c->h_cell_of.clear();
    if (K > 1) c->cells_valid = false;      // the energy kernels need cells of edge >= maxcut: the next energy call re-sorts
    if (!stats && !cstats) return SCGPU_OK;       // asynchronous form: nothing is read back
    std::vector<SweepAcc> acc(c->ncells), cacc(chains ? c->ncells : 0);
    CK(cudaMemcpyAsync(acc.data(), c->d_sweep_acc, (size_t)c->ncells * sizeof(SweepAcc), cudaMemcpyDeviceToHost, c->stream));
    if (chains) CK(cudaMemcpyAsync(cacc.data(), d_chain_acc, (size_t)c->ncells * sizeof(SweepAcc), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (cstats) {
        memset(cstats, 0, sizeof *cstats);
        for (size_t i = 0; i < cacc.size(); i++) {       // fixed order
            cstats->chainm_acc += cacc[i].trans_acc; cstats->chainm_rej += cacc[i].trans_rej;
            cstats->chainr_acc += cacc[i].rot_acc; cstats->chainr_rej += cacc[i].rot_rej;
            cstats->cell_rej += cacc[i].cell_rej; cstats->energy_delta += cacc[i].de; cstats->noop += cacc[i].pad;
        }
    }
    if (phased) {
        int* hflag = (int*)(c->h_small + 256 + 64);
        CK(cudaMemcpyAsync(hflag, c->d_pl_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (*hflag) {
            const int which = *hflag;
            CK(cudaMemsetAsync(c->d_pl_total, 0, 8 * sizeof(int), c->stream));
            g_err = "scgpu_sweep_checkerboard: a pass of the phased sweep could not list its pair terms (flags " + std::to_string(which) + "): its trials were not performed; use SCGPU_SWEEP_KERNEL=rounds for this system";
            return SCGPU_ERR_STATE;
        }
    }
    if (rounds) {
        long long lost = 0;
        for (int i = 0; i < c->ncells; i++) lost += acc[i].pad;
        if (lost) { g_err = "scgpu_sweep_checkerboard: " + std::to_string(lost) + " trial(s) could not be evaluated by this sweep kernel (more partners than its work list holds, or a cell too dense); use SCGPU_SWEEP_KERNEL=cells (or rounds) for this system [grid " + std::to_string(c->nc[0]) + "x" + std::to_string(c->nc[1]) + "x" + std::to_string(c->nc[2]) + ", K " + std::to_string(K) + "]"; return SCGPU_ERR_STATE; }
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int i = 0; i < c->ncells; i++) {       // fixed order
            stats->trans_acc += acc[i].trans_acc; stats->trans_rej += acc[i].trans_rej;
            stats->rot_acc += acc[i].rot_acc; stats->rot_rej += acc[i].rot_rej;
            stats->cell_rej += acc[i].cell_rej; stats->energy_delta += acc[i].de;
        }
    }
    return SCGPU_OK;
}

extern "C" int scgpu_pressure_move(scgpu_ctx* c, const scgpu_pressureparams* pp, uint64_t seed, uint64_t step, scgpu_pressurestats* out) {
    ARG(c && pp && out, "scgpu_pressure_move: NULL argument");
    ARG(pp->temper > 0 && pp->ptype >= 0 && pp->ptype <= 5, "scgpu_pressure_move: temperature must be positive and ptype in 0..5");
    ARG(c->n > 0 && c->box[0] > 0, "scgpu_pressure_move: particles and box must be set first");
    unsigned long long st = seed * 0xA0761D6478BD642Full + step * 0xE7037ED1A0B428DBull + 0x8EBC6AF09C88C6E3ull;
    auto u01h = [&]() { return (double)(splitmix64(st) >> 11) * (1.0 / 9007199254740992.0); };
    double energy = 0.0, enermove = 0.0, etrial = 0.0;
    if (int r = scgpu_all_to_all(c, &energy, nullptr)) return r;
    const double old[3] = {c->box[0], c->box[1], c->box[2]};
    double nb[3] = {old[0], old[1], old[2]};
    const double N = (double)c->n;
    bool positive = true;
    if (pp->ptype == 0) {               // movecreator.cpp:345-396: one edge, chosen at random
        const double rsave = u01h();
        const int side = rsave < 1.0 / 3.0 ? 0 : (rsave < 2.0 / 3.0 ? 1 : 2);
        const double area = old[(side + 1) % 3] * old[(side + 2) % 3];
        nb[side] += pp->edge_mx * (u01h() - 0.5);
        positive = nb[side] > 0.0;
        if (positive) enermove = pp->press * area * (nb[side] - old[side]) - N * pp->temper * log(nb[side] / old[side]);
    } else if (pp->ptype == 1) {        // :398-431 isotropic
        const double psch = pp->edge_mx * (u01h() - 0.5);
        nb[0] += psch; nb[1] += psch; nb[2] += psch;
        positive = nb[0] > 0 && nb[1] > 0 && nb[2] > 0;
        const double pvol = old[0] * old[1] * old[2], pvoln = nb[0] * nb[1] * nb[2];
        if (positive) enermove = pp->press * (pvoln - pvol) - N * pp->temper * log(pvoln / pvol);
    } else if (pp->ptype == 2) {        // :433-466 isotropic in xy, z constant
        const double psch = pp->edge_mx * (u01h() - 0.5);
        nb[0] += psch; nb[1] += psch;
        positive = nb[0] > 0 && nb[1] > 0;
        const double pvol = old[0] * old[1], pvoln = nb[0] * nb[1];
        if (positive) enermove = pp->press * old[2] * (pvoln - pvol) - N * pp->temper * log(pvoln / pvol);
    } else if (pp->ptype == 3) {        // :468-503 xy at constant volume
        const double psch = pp->edge_mx * (u01h() - 0.5);
        nb[0] += psch; nb[1] += psch;
        positive = nb[0] > 0 && nb[1] > 0;
        if (positive) nb[2] = old[0] * old[1] * old[2] / nb[0] / nb[1];
    } else if (pp->ptype == 4) {        // :480-514 "anisotropic in xy, z constant": as written only the x edge ever changes (the branch
        u01h();                         // `if (ran2() - 0.5)` tests a non-zero double), two random numbers are drawn all the same
        nb[0] += pp->edge_mx * (u01h() - 0.5);
        positive = nb[0] > 0;
        const double pvol = old[0] * old[1], pvoln = nb[0] * nb[1];
        if (positive) enermove = pp->press * old[2] * (pvoln - pvol) - N * pp->temper * log(pvoln / pvol);
    } else {                            // :515-541 the y edge only
        nb[1] += pp->edge_mx * (u01h() - 0.5);
        positive = nb[1] > 0;
        if (positive) enermove = pp->press * old[2] * old[0] * (nb[1] - old[1]) - N * pp->temper * log(nb[1] / old[1]);
    }
    bool accept = false;
    if (positive) {
        if (int r = scgpu_set_box(c, nb)) return r;
        if (int r = scgpu_all_to_all(c, &etrial, nullptr)) return r;
        enermove += etrial;
        accept = (enermove <= energy) || (exp(-(enermove - energy) / pp->temper) > u01h());      // moveTry (movecreator.h:175-187)
        if (!accept) { if (int r = scgpu_set_box(c, old)) return r; }
    }
    out->accepted = accept ? 1 : 0;
    out->reserved = 0;
    out->energy_old = energy;
    out->energy_new = positive ? etrial : energy;
    out->enthalpy_delta = accept ? enermove - energy : 0.0;
    for (int d = 0; d < 3; d++) out->box[d] = c->box[d];
    return SCGPU_OK;
}

extern "C" int scgpu_timer_start(scgpu_ctx* c) {
    ARG(c, "scgpu_timer_start: NULL context");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return SCGPU_OK;
}
extern "C" int scgpu_timer_stop(scgpu_ctx* c, float* ms) {
    ARG(c && ms, "scgpu_timer_stop: NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SCGPU_OK;
}
extern "C" int scgpu_sync(scgpu_ctx* c) {
    ARG(c, "scgpu_sync: NULL context");
    CK(cudaSetDevice(c->device));
    bool grew = false;
    if (int r = overflow_then_grow(c, &grew)) return r;
    if (grew) {
        g_err = "scgpu_sync: an asynchronous energy launch could not complete (a work list overflowed and has been grown, or the row-unit gate met a layout it cannot hold and has been switched off); its results are invalid: repeat the call [flags " +
                std::to_string(c->last_overflow) + "]";
        return SCGPU_ERR_STATE;
    }
    return SCGPU_OK;
}
extern "C" int scgpu_kernel_launches(scgpu_ctx* c, int64_t* launches) {
    ARG(c && launches, "scgpu_kernel_launches: NULL argument");
    *launches = c->launches;
    return SCGPU_OK;
}

extern "C" int scgpu_fp64_peak(scgpu_ctx* c, double* tflops) {
    ARG(c && tflops, "scgpu_fp64_peak: NULL argument");
    CK(cudaSetDevice(c->device));
    const int threads = 256, blocks = c->sm_count * 8, iters = 20000;
    double* d = nullptr;
    CK(cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)));
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(d, 2000);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(c->ev0, c->stream));
        k_fp64_peak<<<blocks, threads, 0, c->stream>>>(d, iters);
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaEventSynchronize(c->ev1));
        float ms;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (ms < best) best = ms;
    }
    c->launches += 6;
    cudaFree(d);
    double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return SCGPU_OK;
}

extern "C" int scgpu_flush_l2(scgpu_ctx* c) {
    ARG(c, "scgpu_flush_l2: NULL context");
    CK(cudaSetDevice(c->device));
    if (!c->d_flush) {
        c->flush_n = (size_t)256 * 1024 * 1024 / sizeof(double);   // 256 MB > 126 MB L2
        CK(cudaMalloc(&c->d_flush, c->flush_n * sizeof(double)));
        CK(cudaMemsetAsync(c->d_flush, 0, c->flush_n * sizeof(double), c->stream));
    }
    k_flush<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_flush, c->flush_n);
    c->launches++;
    CK(cudaGetLastError());
    return SCGPU_OK;
}

// ---- Wang-Landau order parameters of the whole configuration (wl_order.cuh) ------------------------------------------------------
#define WL_OFF_PARTIAL 0
#define WL_OFF_PARTIAL_C (WL_OFF_PARTIAL + WL_MAX_BLOCKS * 4 * sizeof(double))
#define WL_OFF_OUT (WL_OFF_PARTIAL_C + WL_MAX_BLOCKS * sizeof(long long))
#define WL_OFF_OUT_C (WL_OFF_OUT + 8 * sizeof(double))
#define WL_OFF_COUNTERS (WL_OFF_OUT_C + sizeof(long long))
#define WL_BYTES (WL_OFF_COUNTERS + 4 * sizeof(int))
#define WL_RESULT_BYTES (WL_BYTES - WL_OFF_OUT)

extern "C" int scgpu_wl_order(scgpu_ctx* c, scgpu_wlorder* io) {
    ARG(c && io, "scgpu_wl_order: NULL argument");
    ARG(c->n > 0 && c->ntypes > 0 && c->types_valid && c->box[0] > 0, "scgpu_wl_order: topology, particles (with types) and box must be set first");
    bool need_mesh = false;
    for (int w = 0; w < 2; w++) {
        const int m = io->wlm[w];
        ARG(m >= 0 && m <= 9, "scgpu_wl_order: wlm must be 0 .. 9");
        ARG(m != 5 && m != 6, "scgpu_wl_order: wlm 5 / 6 (pore radius) have no defined behaviour in the reference: WangLandau::radiusholeAll "
                              "(scOOP/mc/wanglandau.cpp:306-338) never allocates its array (a local shadows radiusholemax) and writes through a null pointer");
        if (m == 0) continue;
        ARG(io->dorder[w] > 0, "scgpu_wl_order: dorder must be positive");
        if (m == 2) need_mesh = true;
        if (m == 4) ARG(c->n >= 2, "scgpu_wl_order: wlm 4 needs two particles");
    }
    if (need_mesh || io->wlm[0] == 7 || io->wlm[1] == 7) ARG(io->wlmtype >= 0 && io->wlmtype < c->ntypes, "scgpu_wl_order: wlmtype is not a particle type of the topology");
    CK(cudaSetDevice(c->device));
    if (int r = sync_api_from_sorted(c)) return r;      // after device sweeps the cell-sorted arrays hold the newest configuration
    if (!c->d_wl) CK(cudaMalloc(&c->d_wl, WL_BYTES));
    if (int r = ensure_pinned(c, WL_RESULT_BYTES)) return r;
    double* d_partial = (double*)(c->d_wl + WL_OFF_PARTIAL);
    long long* d_partial_c = (long long*)(c->d_wl + WL_OFF_PARTIAL_C);
    double* d_out = (double*)(c->d_wl + WL_OFF_OUT);
    long long* d_out_c = (long long*)(c->d_wl + WL_OFF_OUT_C);
    int* d_counters = (int*)(c->d_wl + WL_OFF_COUNTERS);
    CK(cudaMemsetAsync(c->d_wl + WL_OFF_OUT, 0, WL_RESULT_BYTES, c->stream));
    int d0 = 0, d1 = 0;
    {   // the centre of mass and the system volume are always returned: O(N), one pass
        int nb = (c->n + WL_BLOCK * 4 - 1) / (WL_BLOCK * 4);
        if (nb > WL_MAX_BLOCKS) nb = WL_MAX_BLOCKS;
        if (nb < 1) nb = 1;
        k_wl_partial<<<nb, WL_BLOCK, 0, c->stream>>>(c->d_api, c->d_type, c->n, c->d_ia, c->ntypes, c->box[0], c->box[1], c->box[2],
                                                     io->wlmtype, d_partial, d_partial_c);
        k_wl_final<<<1, 32, 0, c->stream>>>(c->d_api, c->n, nb, c->box[0], c->box[1], c->box[2], d_partial, d_partial_c, d_out, d_out_c);
        c->launches += 2;
        CK(cudaGetLastError());
    }
    if (need_mesh) {
        ARG(io->meshsize > 0, "scgpu_wl_order: meshsize must be positive (wlm 2)");
        const double f0 = c->box[0] / io->meshsize, f1 = c->box[1] / io->meshsize;      // Mesh::meshInit, mesh.cpp:15-16
        ARG(f0 >= 1 && f1 >= 1 && f0 * f1 < 1.0e9, "scgpu_wl_order: the mesh must have 1 .. 1e9 points");
        d0 = (int)f0; d1 = (int)f1;
        const int len = d0 * d1;
        if (len > c->wl_mesh_cap) {
            cudaFree(c->d_wl_mesh); cudaFree(c->d_wl_parent); cudaFree(c->d_wl_size);
            c->d_wl_mesh = c->d_wl_parent = c->d_wl_size = nullptr; c->wl_mesh_cap = 0;
            const int cap = len + len / 4;
            CK(cudaMalloc(&c->d_wl_mesh, sizeof(int) * (size_t)cap));
            CK(cudaMalloc(&c->d_wl_parent, sizeof(int) * (size_t)cap));
            CK(cudaMalloc(&c->d_wl_size, sizeof(int) * (size_t)cap));
            c->wl_mesh_cap = cap;
        }
        CK(cudaMemsetAsync(c->d_wl_mesh, 0, sizeof(int) * (size_t)len, c->stream));
        const int gm = (len + WL_BLOCK - 1) / WL_BLOCK;
        k_mesh_fill<<<(c->n + WL_BLOCK - 1) / WL_BLOCK, WL_BLOCK, 0, c->stream>>>(c->d_api, c->d_type, c->n, io->wlmtype, d0, d1, c->d_wl_mesh, d_counters);
        k_mesh_init<<<gm, WL_BLOCK, 0, c->stream>>>(c->d_wl_mesh, len, c->d_wl_parent, c->d_wl_size, d_counters);
        k_mesh_union<<<gm, WL_BLOCK, 0, c->stream>>>(d0, d1, c->d_wl_parent);
        k_mesh_count<<<gm, WL_BLOCK, 0, c->stream>>>(len, c->d_wl_parent, c->d_wl_size);
        k_mesh_max<<<gm, WL_BLOCK, 0, c->stream>>>(len, c->d_wl_size, d_counters);
        c->launches += 5;
        CK(cudaGetLastError());
        c->wl_mesh_len = len; c->wl_mesh_labelled = false;
    }
    CK(cudaMemcpyAsync(c->h_pinned, c->d_wl + WL_OFF_OUT, WL_RESULT_BYTES, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const char* h = (const char*)c->h_pinned;
    const double* out = (const double*)h;
    const long long contacts = *(const long long*)(h + (WL_OFF_OUT_C - WL_OFF_OUT));
    const int* counters = (const int*)(h + (WL_OFF_COUNTERS - WL_OFF_OUT));
    for (int k = 0; k < 3; k++) io->syscm[k] = out[k];
    io->sysvolume = out[3];
    io->mesh_dim[0] = d0; io->mesh_dim[1] = d1;
    io->mesh_occupied = need_mesh ? counters[1] : 0;
    io->mesh_skipped = need_mesh ? counters[2] : 0;
    for (int w = 0; w < 2; w++) {
        double raw = 0;
        int64_t o = 0;
        const double mn = io->minorder[w], dd = io->dorder[w];
        switch (io->wlm[w]) {
            case 1: raw = out[4]; o = (int64_t)ceil((raw - mn) / dd); break;                 // zOrder, wanglandau.h:551-554
            case 2: raw = (double)counters[0]; o = (int64_t)((raw - mn) / dd); break;        // holeXYPlane(wli), wanglandau.h:313-317
            case 3: raw = out[5]; o = (int64_t)floor((raw - mn) / dd); break;                // zOrient, wanglandau.h:321-324
            case 4: raw = out[6]; o = (int64_t)ceil((raw - mn) / dd); break;                 // twoPartDist, wanglandau.h:395-398
            case 7: raw = (double)contacts; o = (int64_t)ceil((raw - mn) / dd); break;       // contParticlesOrder, wanglandau.h:652-654
            case 8: raw = c->box[0]; o = (int64_t)ceil((raw - mn) / dd); break;              // boxSize_x, wanglandau.h:375-377
            case 9: raw = c->box[1]; o = (int64_t)ceil((raw - mn) / dd); break;              // boxSize_y, wanglandau.h:380-382
            default: break;
        }
        io->raw[w] = raw;
        io->order[w] = o;
    }
    return SCGPU_OK;
}

extern "C" int scgpu_wl_mesh(scgpu_ctx* c, int* data, int len) {
    ARG(c && data, "scgpu_wl_mesh: NULL argument");
    ARG(c->wl_mesh_len > 0, "scgpu_wl_mesh: no mesh yet (call scgpu_wl_order with wlm 2 first)");
    ARG(len == c->wl_mesh_len, "scgpu_wl_mesh: len must be mesh_dim[0] * mesh_dim[1] of the last scgpu_wl_order call");
    CK(cudaSetDevice(c->device));
    if (!c->wl_mesh_labelled) {
        k_mesh_label<<<1, 1024, 0, c->stream>>>(len, c->d_wl_parent, c->d_wl_size);        // the hole sizes are not needed any more: d_wl_size takes the labels
        k_mesh_export<<<(len + WL_BLOCK - 1) / WL_BLOCK, WL_BLOCK, 0, c->stream>>>(len, c->d_wl_mesh, c->d_wl_parent, c->d_wl_size);
        c->launches += 2;
        CK(cudaGetLastError());
        c->wl_mesh_labelled = true;
    }
    CK(cudaMemcpyAsync(data, c->d_wl_mesh, sizeof(int) * (size_t)len, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SCGPU_OK;
}

#include "comm.cuh"

#ifdef SW_PROFILE
extern "C" int scgpu_sweep_profile(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, sw_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(sw_prof, z, sizeof z); }
    return 0;
}
#endif
