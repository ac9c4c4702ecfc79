// Wang-Landau order parameters of a WHOLE configuration on the device: the from-scratch forms the
// reference evaluates in WangLandau::init (scOOP/mc/wanglandau.cpp:56-125) and after every volume / switch move in WangLandau::runPress
// and runSwitch (scOOP/mc/wanglandau.h:168-196, 220-238). On the host they are O(N) loops over conf->pvec plus, for the membrane
// hole (wlm 2), a mesh fill and a breadth-first cluster search (Mesh::meshInit, scOOP/mc/mesh.cpp:11-185); a caller that keeps its
// configuration on the device (the batched sweeps, scgpu_pressure_move) would otherwise have to download it for each of them.
//
//   k_wl_partial / k_wl_final   Conf::massCenter (structures/Conf.cpp:78-95), the system volume (mc/inicializer.cpp:32-34), the contact
//                               count of wlm 7 (wanglandau.h:603-615, 661-678) -- block partial sums written to fixed slots and added in
//                               slot order (the same bits on every run) -- and the one- / two-particle quantities of wlm 1, 3, 4
//   k_mesh_fill                 Mesh::meshFill / addPart (mesh.cpp:37-65): -1 on the nine mesh points around every particle of wlmtype
//   k_mesh_init, k_mesh_union, k_mesh_count, k_mesh_max
//                               Mesh::findHoles (mesh.cpp:135-185) as a lock-free union-find over the free mesh points (4-neighbour,
//                               periodic): the hole sizes are those of the reference's breadth-first walk, whatever the order of the unions
//
// Included by scgpu.cu (needs scgpu_iaparam and the CK / ARG macros of that file).
#pragma once

#define WL_BLOCK 256
#define WL_MAX_BLOCKS 592          // 4 x 148: block partials of k_wl_partial
#define WL_CONTACTS_SQ 36.0        // WL_CONTACTS, scOOP/structures/macros.h:107

// INBOX, scOOP/structures/macros.h:119
__device__ __forceinline__ double wl_inbox(double a) {
    double ip;
    return a > 0 ? modf(a, &ip) : modf(a, &ip) + 1;
}

// partial[b][0..3] = sum of volume, sum of (pos - anInt(pos)) * volume (x, y, z); partial_c[b] = contacts with particle 0
__global__ void __launch_bounds__(WL_BLOCK) k_wl_partial(const double* __restrict__ api, const int* __restrict__ type, int n,
                                                         const scgpu_iaparam* __restrict__ ia, int ntypes, double bx, double by, double bz,
                                                         int wlmtype, double* __restrict__ partial, long long* __restrict__ partial_c) {
    __shared__ double sh[4][WL_BLOCK];
    __shared__ int shc[WL_BLOCK];
    const int tid = threadIdx.x;
    double v = 0, cx = 0, cy = 0, cz = 0;
    int cont = 0;
    const double x0 = api[0], y0 = api[1], z0 = api[2];
    // every block owns one contiguous run of particles, every thread walks it with stride WL_BLOCK: a fixed summation tree
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(n, lo + per);
    for (int i = lo + tid; i < hi; i += WL_BLOCK) {
        const double* p = api + (size_t)i * 30;
        const int t = type[i];
        const double vol = ia[(size_t)t * ntypes + t].reserved[2];      // Ia_param::volume of the type (set by scgpu_set_topology)
        const double px = p[0], py = p[1], pz = p[2];
        v += vol;
        cx += __dmul_rn(px - rint(px), vol);
        cy += __dmul_rn(py - rint(py), vol);
        cz += __dmul_rn(pz - rint(pz), vol);
        if (i > 0 && t == wlmtype) {                                    // particlesInContact, wanglandau.h:661-678 (no FMA: strict '<')
            double x = px - x0, y = py - y0, z = pz - z0;
            x = __dmul_rn(bx, x - rint(x));
            y = __dmul_rn(by, y - rint(y));
            z = __dmul_rn(bz, z - rint(z));
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
            if (d2 < WL_CONTACTS_SQ) cont++;
        }
    }
    sh[0][tid] = v; sh[1][tid] = cx; sh[2][tid] = cy; sh[3][tid] = cz; shc[tid] = cont;
    __syncthreads();
    for (int s = WL_BLOCK / 2; s > 0; s >>= 1) {
        if (tid < s) {
            sh[0][tid] += sh[0][tid + s]; sh[1][tid] += sh[1][tid + s]; sh[2][tid] += sh[2][tid + s]; sh[3][tid] += sh[3][tid + s];
            shc[tid] += shc[tid + s];
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int k = 0; k < 4; k++) partial[(size_t)blockIdx.x * 4 + k] = sh[k][0];
        partial_c[blockIdx.x] = shc[0];
    }
}

// out[0..2] syscm, [3] sysvolume, [4] (pos0.z - syscm.z) * box.z (zOrder, wanglandau.h:551-554), [5] dir0.z (zOrient, :321-324),
// [6] xy distance of particles 0 and 1 (twoPartDist, :395-398; 0 when n < 2), [7] contacts as a double; out_c[0] contacts
__global__ void k_wl_final(const double* __restrict__ api, int n, int nblocks, double bx, double by, double bz,
                           const double* __restrict__ partial, const long long* __restrict__ partial_c, double* __restrict__ out,
                           long long* __restrict__ out_c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v = 0, cx = 0, cy = 0, cz = 0;
    long long cont = 0;
    for (int b = 0; b < nblocks; b++) {
        v += partial[(size_t)b * 4]; cx += partial[(size_t)b * 4 + 1]; cy += partial[(size_t)b * 4 + 2]; cz += partial[(size_t)b * 4 + 3];
        cont += partial_c[b];
    }
    out[0] = cx / v; out[1] = cy / v; out[2] = cz / v; out[3] = v;
    out[4] = __dmul_rn(api[2] - out[2], bz);
    out[5] = api[5];
    double d = 0;
    if (n > 1) {
        double rx = api[0] - api[30], ry = api[1] - api[31];
        rx = __dmul_rn(bx, rx - rint(rx));
        ry = __dmul_rn(by, ry - rint(ry));
        d = sqrt(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)));
    }
    out[6] = d;
    out[7] = (double)cont;
    out_c[0] = cont;
}

// Mesh::meshFill + addPart + meshSquare (mesh.cpp:37-65, 81-112). A coordinate whose INBOX() is exactly 1 (a non-positive integer)
// indexes one past the row in the reference (undefined behaviour there): skipped and counted in counters[2].
__global__ void __launch_bounds__(WL_BLOCK) k_mesh_fill(const double* __restrict__ api, const int* __restrict__ type, int n, int wlmtype,
                                                        int d0, int d1, int* __restrict__ mesh, int* __restrict__ counters) {
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    if (i >= n || type[i] != wlmtype) return;
    const double* p = api + (size_t)i * 30;
    const int x = (int)__dmul_rn(wl_inbox(p[0]), (double)d0);
    const int y = (int)__dmul_rn(wl_inbox(p[1]), (double)d1);
    if (x < 0 || y < 0 || x >= d0 || y >= d1) { atomicAdd(&counters[2], 1); return; }
    const int xs[3] = {x, x - 1 < 0 ? d0 - 1 : x - 1, x + 1 == d0 ? 0 : x + 1};
    const int ys[3] = {y, y - 1 < 0 ? d1 - 1 : y - 1, y + 1 == d1 ? 0 : y + 1};
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
        for (int a = 0; a < 3; a++) atomicSub(&mesh[xs[a] + d0 * ys[b]], 1);
}

// free mesh points become their own parent, occupied ones -1; counters[1] = occupied points
__global__ void __launch_bounds__(WL_BLOCK) k_mesh_init(const int* __restrict__ mesh, int len, int* __restrict__ parent, int* __restrict__ size,
                                                        int* __restrict__ counters) {
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    int occ = 0;
    if (i < len) {
        occ = mesh[i] != 0;           // after a fill from zero the mesh is <= 0 everywhere (findHoles' `> 0 -> 0` pass is a no-op)
        parent[i] = occ ? -1 : i;
        size[i] = 0;
    }
    const int total = __syncthreads_count(occ);
    if (threadIdx.x == 0 && total) atomicAdd(&counters[1], total);
}

__device__ __forceinline__ int uf_find(int* parent, int i) {
    volatile int* p = parent;
    int q = p[i];
    while (q != i) { i = q; q = p[i]; }
    return i;
}

// the larger root is hooked under the smaller one with a compare-and-swap; a failed swap means another thread hooked it first: retry
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        if (atomicCAS(&parent[b], b, a) == b) return;
    }
}

// Mesh::meshNeighbors (mesh.cpp:114-133): every free point joins its free right and upper neighbour (the other two directions are
// some other point's right / upper neighbour)
__global__ void __launch_bounds__(WL_BLOCK) k_mesh_union(int d0, int d1, int* parent) {
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    if (i >= d0 * d1 || parent[i] < 0) return;
    const int x = i % d0, y = i / d0;
    const int r = (x + 1 == d0 ? 0 : x + 1) + d0 * y;
    const int u = x + d0 * (y + 1 == d1 ? 0 : y + 1);
    if (r != i && ((volatile int*)parent)[r] >= 0) uf_unite(parent, i, r);
    if (u != i && ((volatile int*)parent)[u] >= 0) uf_unite(parent, i, u);
}

__global__ void __launch_bounds__(WL_BLOCK) k_mesh_count(int len, int* parent, int* size) {
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    if (i >= len || parent[i] < 0) return;
    atomicAdd(&size[uf_find(parent, i)], 1);
}

// counters[0] = the largest hole (mesh points)
__global__ void __launch_bounds__(WL_BLOCK) k_mesh_max(int len, const int* __restrict__ size, int* __restrict__ counters) {
    __shared__ int sh[WL_BLOCK / 32];
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    int m = i < len ? size[i] : 0;
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < WL_BLOCK / 32; w++) m = max(m, sh[w]);
        if (m > 0) atomicMax(&counters[0], m);
    }
}

// The mesh as Mesh::findHoles leaves it in Mesh::data (mesh.cpp:135-185): occupied points keep their (negative) count, every free point carries
// the number of its hole, holes numbered 1, 2, ... in the order the reference's scan meets them = by their smallest mesh index -- which is
// the root of the union-find (the larger root is always hooked under the smaller). The reference's INCREMENTAL updates
// (Mesh::addPart / removePart, meshOrderMoveMolecule) go on from exactly this array, so a caller that lets the device do the from-scratch
// search and the host the incremental steps needs it bit for bit (scgpu_wl_mesh).
__global__ void __launch_bounds__(1024) k_mesh_label(int len, const int* __restrict__ parent, int* __restrict__ label) {
    __shared__ int sh[1024];
    const int tid = threadIdx.x;
    const int per = (len + 1023) / 1024;
    const int lo = min(len, tid * per), hi = min(len, lo + per);
    int cnt = 0;
    for (int i = lo; i < hi; i++) cnt += parent[i] == i;
    sh[tid] = cnt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // inclusive scan of the per-thread root counts
        const int v = tid >= off ? sh[tid - off] : 0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    int run = sh[tid] - cnt;
    for (int i = lo; i < hi; i++) if (parent[i] == i) label[i] = ++run;
}

__global__ void __launch_bounds__(WL_BLOCK) k_mesh_export(int len, int* mesh, int* parent, const int* __restrict__ label) {
    const int i = blockIdx.x * WL_BLOCK + threadIdx.x;
    if (i >= len || parent[i] < 0) return;
    mesh[i] = label[uf_find(parent, i)];
}
