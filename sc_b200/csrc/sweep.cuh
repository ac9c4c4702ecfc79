// Checkerboard cell-sublattice sweeps of displacement / rotation trial moves.
//
// The reference's sweep (Updater::simulate, scOOP/mc/updater.cpp:206-230) is N sequential single-particle trials, each
//   partDisplace / partRotate (scOOP/mc/movecreator.cpp:947-1028): old energy, proposal, trial energy, moveTry (movecreator.h:175-187).
// Here the box is cut into cells of edge >= maxcut on an EVEN grid, the cells are 2-coloured per axis (<= 8 colours) and
// all cells of one colour are updated concurrently, one thread block per cell: two active cells are separated by a full
// cell (>= maxcut), so their trial energies never involve each other's moving particles, provided a move that would take a
// particle out of its cell is rejected. Inside a cell the trials are sequential, exactly as in the reference:
// uniformly chosen particle (with replacement), displacement with p = 1/2 (always for spheres) else rotation, fixed-length
// displacement trans_mx in a uniform direction (Vector::randomUnitSphere, scOOP/structures/Vector.h:190-205), rotation by
// angle*u about a uniform axis with random sense (Particle::pscRotate, scOOP/structures/particle.h:182-272), Metropolis test
// dE <= 0 or exp(-dE/T) > u. The grid is shifted by a random vector every sweep so that cell walls do not pin anything.
// Random numbers: Philox4x32-10 keyed by (seed), counter (sweep, colour, cell, trial) -> the trajectory is a pure function of
// (seed, configuration) and reproducible run to run. Validated STATISTICALLY against sequential sweeps (tests/).
#pragma once
// (textually included by scgpu.cu after DevSys, cell_index, rel_frac and warp_sum are defined)

struct SweepParams {
    double temper;
    double trans_mx[40];
    double rot_angle[40];
    int n_sub;
    int geotype_of_type[40];
};

struct SweepAcc {      // per cell, written once per colour pass (summed on the host side in a fixed order)
    int trans_acc, trans_rej, rot_acc, rot_rej, cell_rej, pad;
    double de;
};

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, 4 x 32 random bits per call
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ inline uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {    // 53-bit uniform in [0,1)
    return ((double)(((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11)) + 0.5) * (1.0 / 9007199254740992.0) * 0.99999999999999989;
}

// Particle::pscRotate on an internal record (quaternion half-angle convention of the reference: vc = cos(angle))
__device__ inline void rotate_record(double* r, int geotype, double angle, const v3& axis, bool positive) {
    double vc = cos(angle);
    double vs = positive ? sqrt(1.0 - vc * vc) : -sqrt(1.0 - vc * vc);
    double qw = vc, qx = axis.x * vs, qy = axis.y * vs, qz = axis.z * vs;
    double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy, t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t10 = -qz * qz;
    double d1 = t8 + t10, d2 = t6 - t4, d3 = t3 + t7, d4 = t4 + t6, d5 = t5 + t10, d6 = t9 - t2, d7 = t7 - t3, d8 = t2 + t9, d9 = t5 + t8;
    auto rot = [&](int off) {
        double x = r[off], y = r[off + 1], z = r[off + 2];
        r[off] = 2.0 * (d1 * x + d2 * y + d3 * z) + x;
        r[off + 1] = 2.0 * (d4 * x + d5 * y + d6 * z) + y;
        r[off + 2] = 2.0 * (d7 * x + d8 * y + d9 * z) + z;
    };
    rot(R_DIR);
    if (geotype != SCGPU_SCN && geotype != SCGPU_SCA) {
        rot(R_PD0); rot(R_S0); rot(R_S1);
        if (is_two_patch(geotype)) { rot(R_PD1); rot(R_S2); rot(R_S3); }
    }
    if (is_chiral(geotype)) {
        rot(R_CH0);
        if (geotype == SCGPU_TCHPSC || geotype == SCGPU_TCHCPSC) rot(R_CH1);
    }
}

#ifndef SW_WARPS_N
#define SW_WARPS_N 2
#endif
constexpr int SW_WARPS = SW_WARPS_N;
#ifndef SW_MINBLOCKS
#define SW_MINBLOCKS 1
#endif
constexpr int SW_TILE = 1280;     // staged neighbourhood (FP32 relative coordinates + slot): 1280 x 20 B = 25 KB

// one block per ACTIVE cell of the current colour
__global__ void __launch_bounds__(SW_WARPS * 32, SW_MINBLOCKS)
k_sweep_colour(DevSys s, SweepParams sp, unsigned long long seed, unsigned long long sweep, int colour, int3 ncol,
               double4* posw, double* rec, SweepAcc* acc_out, int* fail_flag) {
    __shared__ float4 t_pf[SW_TILE];
    __shared__ int t_slot[SW_TILE];
    __shared__ double sh_old[REC], sh_new[REC];
    __shared__ int sh_queue[SW_WARPS][96];
    __shared__ int sh_pl[SW_WARPS][32];       // per-warp lists of (slot, state) entries that owe a patch evaluation
    __shared__ int sh_pc[SW_WARPS];
    __shared__ double sh_eo[SW_WARPS], sh_en[SW_WARPS];
    __shared__ int sh_b[28], sh_off[28];
    __shared__ int sh_ctl[4];     // [0] accept, [1] picked slot, [2] displacement?, [3] stayed in its cell?
    __shared__ double sh_u[12];   // the trial's uniforms
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    // active cell of this block
    const int ax = s.nc[0] / ncol.x, ay = s.nc[1] / ncol.y;
    const int bx = blockIdx.x % ax, by = (blockIdx.x / ax) % ay, bz = blockIdx.x / (ax * ay);
    const int cx = bx * ncol.x + (colour % ncol.x), cy = by * ncol.y + ((colour / ncol.x) % ncol.y), cz = bz * ncol.z + (colour / (ncol.x * ncol.y));
    const int c0 = (cz * s.nc[1] + cy) * s.nc[0] + cx;
    const int tb = s.cell_start[c0], te = s.cell_start[c0 + 1];
    const int npart = te - tb;
    SweepAcc acc = {0, 0, 0, 0, 0, 0, 0.0};
    if (npart == 0) { if (threadIdx.x == 0) acc_out[c0] = acc; return; }
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
    const bool tiled = C <= SW_TILE;        // denser neighbourhoods are scanned from global memory (slower, same results)
    (void)fail_flag;
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    const float pre_cut = (float)(s.sqmaxcut * 1.001);
    auto slot_of_p = [&](int p) {
        int k = 0;
        while (k + 1 < ncell_nb && sh_off[k + 1] <= p) k++;
        return sh_b[k] + (p - sh_off[k]);
    };
    auto staged = [&](int slot) {
        double4 pw = posw[slot];
        return make_float4((float)rel_frac(pw.x + s.shift[0], ccen[0]), (float)rel_frac(pw.y + s.shift[1], ccen[1]),
                           (float)rel_frac(pw.z + s.shift[2], ccen[2]), 0.f);
    };
    if (tiled) {
        for (int k = wid; k < ncell_nb; k += SW_WARPS) {
            const int b = sh_b[k], off = sh_off[k], len = sh_off[k + 1] - off;
            for (int idx = lane; idx < len; idx += 32) { t_pf[off + idx] = staged(b + idx); t_slot[off + idx] = b + idx; }
        }
    }
    __syncthreads();
    const int ntrial = npart * sp.n_sub;
    for (int trial = 0; trial < ntrial; trial++) {
        // ---- random numbers of this trial (thread 0), Philox counter = (sweep, colour, cell, 3*trial + k)
        if (threadIdx.x == 0) {
            const uint32_t c1 = (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 28);
            for (int k = 0; k < 3; k++) {
                uint4 r = philox4x32((uint32_t)sweep, c1, (uint32_t)c0, (uint32_t)(3 * trial + k), (uint32_t)seed, (uint32_t)(seed >> 32));
                sh_u[2 * k] = u01(r.x, r.y);
                sh_u[2 * k + 1] = u01(r.z, r.w);
            }
            int pick = tb + (int)(sh_u[0] * npart);          // uniformly chosen particle of this cell, with replacement
            if (pick >= te) pick = te - 1;
            sh_ctl[1] = pick;
        }
        __syncthreads();
        const int tslot = sh_ctl[1];
        if (threadIdx.x < REC) { double v = rec[(size_t)tslot * REC + threadIdx.x]; sh_old[threadIdx.x] = v; sh_new[threadIdx.x] = v; }
        const double4 tpw = posw[tslot];
        const int target = w_orig(tpw.w), type1 = w_type(tpw.w), moltype1 = w_moltype(tpw.w);
        __syncthreads();
        if (threadIdx.x == 0) {
            int g = sp.geotype_of_type[type1];
            bool displace = (g >= SCGPU_SPN) || (sh_u[1] < 0.5);                 // particleMove (movecreator.cpp:11-33)
            double z = 1.0 - 2.0 * sh_u[2], phi = 6.283185307179586476925 * sh_u[3];
            double rr = sqrt(fmax(0.0, 1.0 - z * z));
            v3 u = mk(rr * cos(phi), rr * sin(phi), z);                           // uniform on the unit sphere
            if (displace) {            // partDisplace (movecreator.cpp:947-994): fixed length trans_mx, uniform direction
                double mx = sp.trans_mx[type1];
                sh_new[R_POS] += u.x * mx / s.box[0];
                sh_new[R_POS + 1] += u.y * mx / s.box[1];
                sh_new[R_POS + 2] += u.z * mx / s.box[2];
            } else {                   // partRotate (movecreator.cpp:996-1028)
                rotate_record(sh_new, g, sp.rot_angle[type1] * sh_u[4], u, sh_u[5] < 0.5);
            }
            sh_ctl[2] = displace ? 1 : 0;
            // a move that leaves the cell would break the independence of the active cells: reject it
            sh_ctl[3] = (cell_index(sh_new + R_POS, s.shift, s.nc) == c0) ? 1 : 0;
            sh_ctl[0] = 0;
        }
        __syncthreads();
        const bool in_cell = sh_ctl[3] != 0;
        const bool displace = sh_ctl[2] != 0;
        const double u_acc = sh_u[1] < 0.5 ? 2.0 * sh_u[1] : 2.0 * sh_u[1] - 1.0;   // the move-type bit is used up; the rest is still uniform
        double e_old = 0.0, e_new = 0.0;
        if (in_cell) {
            ConList cl;
            get_conlist(s.mol, moltype1, target, cl);
            const v3 po = ld3(sh_old + R_POS), pn = ld3(sh_new + R_POS);
            const float ox = (float)rel_frac(po.x + s.shift[0], ccen[0]), oy = (float)rel_frac(po.y + s.shift[1], ccen[1]), oz = (float)rel_frac(po.z + s.shift[2], ccen[2]);
            const float nxf = (float)rel_frac(pn.x + s.shift[0], ccen[0]), nyf = (float)rel_frac(pn.y + s.shift[1], ccen[1]), nzf = (float)rel_frac(pn.z + s.shift[2], ccen[2]);
            int* queue = sh_queue[wid];
            int qn = 0;
            double lo = 0.0, ln = 0.0;
            // one queue entry = (partner slot, which state): the old and the new state of a trial are evaluated on DIFFERENT
            // lanes. Phase A (all warps): exact gate + everything but the rod-rod patch term; entries that owe a patch term are
            // collected per warp. Phase B (warp 0 only): the patch terms of the whole trial, packed on as few lanes as there
            // are entries -- the other warps wait at the barrier instead of issuing 3-lanes-active patch code four times over.
            int pc = 0;
            auto patch_entry = [&](int entry) {
                const int slot = entry >> 1;
                const bool is_new = entry & 1;
                double4 pw = posw[slot];
                v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                double e = pair_energy_patch(s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC);
                if (is_new) ln += e; else lo += e;
            };
            auto eval = [&](int entry, bool on) {
                bool np = false;
                if (on) {
                    const int slot = entry >> 1;
                    const bool is_new = entry & 1;
                    double4 pw = posw[slot];
                    int orig = w_orig(pw.w);
                    v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                    double d = dot(r, r);
                    bool bonded = !cl.is_empty && (orig == cl.con[0] || orig == cl.con[1] || orig == cl.con[2] || orig == cl.con[3]);
                    if (d <= s.sqmaxcut || bonded) {
                        double e = pair_energy_cheap<false>(s.box, s.ia, s.ntypes, s.mol, r, d, is_new ? sh_new : sh_old, type1, moltype1,
                                                            rec + (size_t)slot * REC, w_type(pw.w), orig, cl, np);
                        if (is_new) ln += e; else lo += e;
                    }
                }
                unsigned m = __ballot_sync(0xffffffffu, np);
                if (m) {
                    int c = __popc(m);
                    if (pc + c > 32) {                 // list full (never in physical configurations): evaluate in place
                        if (np) patch_entry(entry);
                    } else {
                        if (np) sh_pl[wid][pc + __popc(m & lt_mask)] = entry;
                        pc += c;
                    }
                    __syncwarp();
                }
            };
            for (int base = wid * 32; base < C; base += SW_WARPS * 32) {
                int p = base + lane;
                bool pass_o = false, pass_n = false;
                int slot = 0;
                if (p < C) {
                    slot = tiled ? t_slot[p] : slot_of_p(p);
                    if (slot != tslot) {
                        float4 q = tiled ? t_pf[p] : staged(slot);
                        float dx = ox - q.x, dy = oy - q.y, dz = oz - q.z;
                        dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                        float ex = nxf - q.x, ey = nyf - q.y, ez = nzf - q.z;
                        ex = (ex - rintf(ex)) * boxf[0]; ey = (ey - rintf(ey)) * boxf[1]; ez = (ez - rintf(ez)) * boxf[2];
                        pass_o = (dx * dx + dy * dy + dz * dz <= pre_cut);
                        pass_n = (ex * ex + ey * ey + ez * ez <= pre_cut);
                        if (!cl.is_empty) {       // bonded partners are evaluated by index below
                            int orig = w_orig(posw[slot].w);
                            if (orig == cl.con[0] || orig == cl.con[1] || orig == cl.con[2] || orig == cl.con[3]) { pass_o = false; pass_n = false; }
                        }
                    }
                }
                unsigned mo = __ballot_sync(0xffffffffu, pass_o), mn = __ballot_sync(0xffffffffu, pass_n);
                int no = __popc(mo);
                if (pass_o) queue[qn + __popc(mo & lt_mask)] = slot * 2;
                if (pass_n) queue[qn + no + __popc(mn & lt_mask)] = slot * 2 + 1;
                qn += no + __popc(mn);
                __syncwarp();
                while (qn >= 32) {
                    eval(queue[lane], true);
                    int rest = qn - 32;
                    int mv0 = (lane < rest) ? queue[32 + lane] : 0;
                    int mv1 = (lane + 32 < rest) ? queue[64 + lane] : 0;
                    __syncwarp();
                    if (lane < rest) queue[lane] = mv0;
                    if (lane + 32 < rest) queue[32 + lane] = mv1;
                    qn = rest;
                    __syncwarp();
                }
            }
            if (qn > 0) eval(lane < qn ? queue[lane] : 0, lane < qn);
            if (wid == 0 && !cl.is_empty) {
                bool on = lane < 8 && cl.con[lane >> 1] >= 0;
                eval(on ? s.slot_of[cl.con[lane >> 1]] * 2 + (lane & 1) : 0, on);
            }
            if (lane == 0) sh_pc[wid] = pc;
            __syncthreads();
            if (wid == 0) {        // phase B: all patch terms of this trial on one warp, two lanes per term, in warp/list order
                int tot = 0;
                for (int w = 0; w < SW_WARPS; w++) tot += sh_pc[w];
                for (int base = 0; base < 2 * tot; base += 32) {
                    const int idx = (base + lane) >> 1;
                    const bool act = idx < tot;
                    int entry = 0;
                    if (act) {
                        int w = 0, k = idx;
                        while (k >= sh_pc[w]) { k -= sh_pc[w]; w++; }
                        entry = sh_pl[w][k];
                    }
                    const int slot = entry >> 1;
                    const bool is_new = entry & 1;
                    double4 pw = posw[slot];
                    v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                    double e = pair_energy_patch_two_lanes(s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC, act);
                    if (is_new) ln += e; else lo += e;
                }
            }
            e_old = warp_sum(lo);
            e_new = warp_sum(ln);
        } else {
            __syncthreads();                           // keep the barrier count identical on both paths
        }
        __syncthreads();
        if (lane == 0) { sh_eo[wid] = e_old; sh_en[wid] = e_new; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double eo = 0.0, en = 0.0;
            for (int k = 0; k < SW_WARPS; k++) { eo += sh_eo[k]; en += sh_en[k]; }     // fixed order
            bool accept = false;
            if (!in_cell) acc.cell_rej++;
            else {
                double de = en - eo;
                accept = (de <= 0.0) || (exp(-de / sp.temper) > u_acc);                // moveTry (movecreator.h:175-187)
                if (accept) acc.de += de;
            }
            if (displace) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
            else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
            sh_ctl[0] = accept ? 1 : 0;
        }
        __syncthreads();
        if (sh_ctl[0]) {          // commit in place: sorted record, position word, staged FP32 copy
            if (threadIdx.x < REC) rec[(size_t)tslot * REC + threadIdx.x] = sh_new[threadIdx.x];
            if (threadIdx.x == 32) posw[tslot] = make_double4(sh_new[R_POS], sh_new[R_POS + 1], sh_new[R_POS + 2], tpw.w);
            for (int p = threadIdx.x; tiled && p < C; p += blockDim.x)
                if (t_slot[p] == tslot)
                    t_pf[p] = make_float4((float)rel_frac(sh_new[R_POS] + s.shift[0], ccen[0]), (float)rel_frac(sh_new[R_POS + 1] + s.shift[1], ccen[1]),
                                          (float)rel_frac(sh_new[R_POS + 2] + s.shift[2], ccen[2]), 0.f);
            __threadfence_block();
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) acc_out[c0] = acc;
}

