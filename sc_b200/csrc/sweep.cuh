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
    double trial_scale;             // share of the sweep's trials that are single-particle moves (1 - chainprob)
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

// Particle::pscRotate on an internal record (quaternion half-angle convention of the reference: vc = cos(angle)),
// split into the coefficients (which depend on angle and axis only) and their application to one vector of the record
__device__ inline void rotation_coefficients(double* d, double angle, const v3& axis, bool positive) {
    double vc = cos(angle);
    double vs = positive ? sqrt(1.0 - vc * vc) : -sqrt(1.0 - vc * vc);
    double qw = vc, qx = axis.x * vs, qy = axis.y * vs, qz = axis.z * vs;
    double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy, t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t10 = -qz * qz;
    d[0] = t8 + t10; d[1] = t6 - t4; d[2] = t3 + t7; d[3] = t4 + t6; d[4] = t5 + t10; d[5] = t9 - t2; d[6] = t7 - t3; d[7] = t2 + t9; d[8] = t5 + t8;
}
__device__ __forceinline__ void rotate_vector(double* r, const double* d) {
    double x = r[0], y = r[1], z = r[2];
    r[0] = 2.0 * (d[0] * x + d[1] * y + d[2] * z) + x;
    r[1] = 2.0 * (d[3] * x + d[4] * y + d[5] * z) + y;
    r[2] = 2.0 * (d[6] * x + d[7] * y + d[8] * z) + z;
}
// which vectors of the record (index = offset / 3: dir, pd0, s0, s1, pd1, s2, s3, ch0, ch1) a rotation touches for a geotype
__device__ __forceinline__ bool record_vector_rotates(int geotype, int v) {
    if (v == 0) return true;
    if (v <= 3) return geotype != SCGPU_SCN && geotype != SCGPU_SCA;
    if (v <= 6) return geotype != SCGPU_SCN && geotype != SCGPU_SCA && is_two_patch(geotype);
    if (v == 7) return is_chiral(geotype);
    return geotype == SCGPU_TCHPSC || geotype == SCGPU_TCHCPSC;
}
__device__ inline void rotate_record(double* r, int geotype, double angle, const v3& axis, bool positive) {
    double d[9];
    rotation_coefficients(d, angle, axis, positive);
    for (int v = 0; v < 9; v++) if (record_vector_rotates(geotype, v)) rotate_vector(r + 3 * v, d);
}

#ifndef SW_WARPS_N
#define SW_WARPS_N 2
#endif
constexpr int SW_WARPS = SW_WARPS_N;
#ifndef SW_MINBLOCKS
#define SW_MINBLOCKS 1
#endif
constexpr int SW_TILE = 1280;     // staged neighbourhood (FP32 relative coordinates + slot): 1280 x 20 B = 25 KB
constexpr int SW_BATCH = 32;      // trials whose random numbers and proposal geometry are prepared together, one thread each

struct SweepProposal {            // everything of a trial that does not depend on the outcome of earlier trials
    int slot, displace;
    double u_acc;
    double m[9];                  // displacement (m[0..2], box-fractional) or the rotation coefficients d1..d9 of pscRotate
};

#ifdef SW_PROFILE      // debug build only: cycles per phase of a trial, summed over blocks (thread 0's view)
__device__ unsigned long long sw_prof[16];
#define SWP_MARK(k) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&sw_prof[k], (unsigned long long)(t_ - swp_t)); swp_t = t_; } } while (0)
#else
#define SWP_MARK(k) do { } while (0)
#endif

// one block per ACTIVE cell of the current colour
// ONE (with RODS): a single particle type is present -> its interaction-table entry is a kernel parameter (constant bank); in this
// latency-bound kernel every field fetched through the load/store unit sits on the serial path of a trial
template <bool RODS, bool ONE>
__global__ void __launch_bounds__(SW_WARPS * 32, SW_MINBLOCKS)
k_sweep_colour(DevSys s, SweepParams sp, unsigned long long seed, unsigned long long sweep, int colour, int3 ncol,
               double4* posw, double* rec, SweepAcc* acc_out, int* fail_flag, const __grid_constant__ scgpu_iaparam ia1) {
    __shared__ float4 t_pf[SW_TILE];
    __shared__ int t_slot[SW_TILE];
    __shared__ double sh_old[REC], sh_new[REC];
    __shared__ int sh_queue[SW_WARPS][96];
    __shared__ int sh_pl[SW_WARPS][32];       // per-warp lists of (slot, state) entries that owe a patch evaluation
    __shared__ double sh_eo[SW_WARPS], sh_en[SW_WARPS];
    __shared__ int sh_b[28], sh_off[28];
    __shared__ SweepProposal sh_prop[SW_BATCH];
    static_assert(SW_WARPS * 32 >= SW_BATCH && SW_WARPS * 32 >= REC + 1, "block too small");
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
    // the active cell is the centre of its own neighbourhood: where its particles sit in the staged tile
    const int k_centre = (nz == 1 ? 0 : nx * ny) + (ny == 1 ? 0 : nx) + (nx == 1 ? 0 : 1);
    const int centre_off = sh_off[k_centre];
    // every non-empty cell performs the same number of trials (n_sub * N / non-empty cells, stochastically rounded): the count
    // does not depend on anything a trial can change (particles never leave their cell within a pass), so detailed balance
    // holds, and all blocks of a pass finish together instead of waiting for the fullest cell
    int ntrial;
    {
        const double avg = sp.trial_scale * (double)sp.n_sub * (double)s.n / (double)s.cell_start[s.ncells + 1];
        const uint4 r = philox4x32((uint32_t)sweep, (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 28), (uint32_t)c0, 0xffffffffu, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double fl = floor(avg);
        ntrial = (int)fl + (u01(r.x, r.y) < avg - fl ? 1 : 0);
    }
#ifdef SW_PROFILE
    long long swp_t = clock64();
    if (threadIdx.x == 0) atomicAdd(&sw_prof[15], (unsigned long long)ntrial);
#endif
    for (int trial = 0; trial < ntrial; trial++) {
        const int bi = trial % SW_BATCH;
        if (bi == 0) {
            // ---- random numbers and proposal geometry of the next SW_BATCH trials, one thread per trial.
            // Philox counter = (sweep, colour, cell, 3*trial + k); nothing here depends on earlier acceptances.
            __syncthreads();
            const int tr = trial + (int)threadIdx.x;
            if (threadIdx.x < SW_BATCH && tr < ntrial) {
                double u[6];
                const uint32_t c1 = (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 28);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    uint4 r = philox4x32((uint32_t)sweep, c1, (uint32_t)c0, (uint32_t)(3 * tr + k), (uint32_t)seed, (uint32_t)(seed >> 32));
                    u[2 * k] = u01(r.x, r.y);
                    u[2 * k + 1] = u01(r.z, r.w);
                }
                int pick = tb + (int)(u[0] * npart);          // uniformly chosen particle of this cell, with replacement
                if (pick >= te) pick = te - 1;
                const int ty = w_type(posw[pick].w);
                const int g = sp.geotype_of_type[ty];
                const bool displace = (g >= SCGPU_SPN) || (u[1] < 0.5);                 // particleMove (movecreator.cpp:11-33)
                const double z = 1.0 - 2.0 * u[2], phi = 6.283185307179586476925 * u[3];
                const double rr = sqrt(fmax(0.0, 1.0 - z * z));
                const v3 ax3 = mk(rr * cos(phi), rr * sin(phi), z);                     // uniform on the unit sphere
                SweepProposal& P = sh_prop[threadIdx.x];
                P.slot = pick;
                P.displace = displace ? 1 : 0;
                P.u_acc = u[1] < 0.5 ? 2.0 * u[1] : 2.0 * u[1] - 1.0;   // the move-type bit is used up; the rest is still uniform
                if (displace) {            // partDisplace (movecreator.cpp:947-994): fixed length trans_mx, uniform direction
                    const double mx = sp.trans_mx[ty];
                    P.m[0] = ax3.x * mx / s.box[0]; P.m[1] = ax3.y * mx / s.box[1]; P.m[2] = ax3.z * mx / s.box[2];
                } else {                   // partRotate (movecreator.cpp:996-1028)
                    rotation_coefficients(P.m, sp.rot_angle[ty] * u[4], ax3, u[5] < 0.5);
                }
            }
            __syncthreads();
        }
        SWP_MARK(0);
        const int tslot = sh_prop[bi].slot;
        const bool displace = sh_prop[bi].displace != 0;
        const double u_acc = sh_prop[bi].u_acc;
        if (threadIdx.x < REC) { double v = rec[(size_t)tslot * REC + threadIdx.x]; sh_old[threadIdx.x] = v; sh_new[threadIdx.x] = v; }
        const double4 tpw = posw[tslot];
        const int target = w_orig(tpw.w), type1 = w_type(tpw.w), moltype1 = w_moltype(tpw.w);
        __syncthreads();
        SWP_MARK(1);
        if (displace) {
            if (threadIdx.x < 3) sh_new[R_POS + threadIdx.x] += sh_prop[bi].m[threadIdx.x];
        } else if (threadIdx.x < 9) {      // one thread per vector of the record
            const int g = sp.geotype_of_type[type1];
            const int off = 3 * threadIdx.x;                   // R_DIR, R_PD0, R_S0, R_S1, R_PD1, R_S2, R_S3, R_CH0, R_CH1
            if (record_vector_rotates(g, threadIdx.x)) rotate_vector(sh_new + off, sh_prop[bi].m);
        }
        __syncthreads();
        SWP_MARK(2);
        // a move that leaves the cell would break the independence of the active cells: reject it
        const bool in_cell = cell_index(sh_new + R_POS, s.shift, s.nc) == c0;
        double e_old = 0.0, e_new = 0.0;
        if (in_cell) {
            ConList cl;
            if (RODS) { cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0; }
            else get_conlist(s.mol, moltype1, target, cl);
            const v3 po = ld3(sh_old + R_POS), pn = ld3(sh_new + R_POS);
            const float ox = (float)rel_frac(po.x + s.shift[0], ccen[0]), oy = (float)rel_frac(po.y + s.shift[1], ccen[1]), oz = (float)rel_frac(po.z + s.shift[2], ccen[2]);
            const float nxf = (float)rel_frac(pn.x + s.shift[0], ccen[0]), nyf = (float)rel_frac(pn.y + s.shift[1], ccen[1]), nzf = (float)rel_frac(pn.z + s.shift[2], ccen[2]);
            int* queue = sh_queue[wid];
            int qn = 0;
            double lo = 0.0, ln = 0.0;
            // one queue entry = (partner slot, which state): the old and the new state of a trial are evaluated on DIFFERENT
            // lanes. Phase A: exact gate + everything but the rod-rod patch term; entries that owe a patch term are collected
            // per warp. Phase B: each warp evaluates its own patch terms, two lanes per term.
            int pc = 0;
            auto patch_entry = [&](int entry) {
                const int slot = entry >> 1;
                const bool is_new = entry & 1;
                double4 pw = posw[slot];
                v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                double e = pair_energy_patch(ONE ? ia1 : s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC);
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
                    bool bonded = !RODS && !cl.is_empty && (orig == cl.con[0] || orig == cl.con[1] || orig == cl.con[2] || orig == cl.con[3]);
                    if (d <= s.sqmaxcut || bonded) {
                        double e;
                        if (RODS && ONE) e = pair_energy_cheap_rods(ia1, r, d, ld3((is_new ? sh_new : sh_old) + R_DIR), ld3(rec + (size_t)slot * REC + R_DIR), np);
                        else e = pair_energy_cheap<RODS>(s.box, s.ia, s.ntypes, s.mol, r, d, is_new ? sh_new : sh_old, type1, moltype1,
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
                        if (!RODS && !cl.is_empty) {       // bonded partners are evaluated by index below
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
            if (!RODS && wid == 0 && !cl.is_empty) {
                bool on = lane < 8 && cl.con[lane >> 1] >= 0;
                eval(on ? s.slot_of[cl.con[lane >> 1]] * 2 + (lane & 1) : 0, on);
            }
            SWP_MARK(3);
            for (int base = 0; base < 2 * pc; base += 32) {      // phase B: this warp's patch terms, two lanes per term, in list order
                const int idx = (base + lane) >> 1;
                const bool act = idx < pc;
                const int entry = act ? sh_pl[wid][idx] : 0;
                const int slot = entry >> 1;
                const bool is_new = entry & 1;
                double4 pw = posw[slot];
                v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                double e = pair_energy_patch_two_lanes(ONE ? ia1 : s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC, act);
                if (is_new) ln += e; else lo += e;
            }
            SWP_MARK(5);
            e_old = warp_sum(lo);
            e_new = warp_sum(ln);
        }
        if (lane == 0) { sh_eo[wid] = e_old; sh_en[wid] = e_new; }
        __syncthreads();
        SWP_MARK(6);
        bool accept = false;
        if (in_cell) {
            double de = 0.0;
            if (lane == 0) {          // every warp takes the same decision from the same numbers (no broadcast barrier)
                double eo = 0.0, en = 0.0;
                for (int k = 0; k < SW_WARPS; k++) { eo += sh_eo[k]; en += sh_en[k]; }     // fixed order
                de = en - eo;
                accept = (de <= 0.0) || (exp(-de / sp.temper) > u_acc);                    // moveTry (movecreator.h:175-187)
            }
            accept = __shfl_sync(0xffffffffu, accept ? 1 : 0, 0) != 0;
            if (threadIdx.x == 0 && accept) acc.de += de;
        }
        if (threadIdx.x == 0) {
            if (!in_cell) acc.cell_rej++;
            if (displace) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
            else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
        }
        if (accept) {          // commit in place: sorted record, position word, staged FP32 copy
            if (threadIdx.x < REC) rec[(size_t)tslot * REC + threadIdx.x] = sh_new[threadIdx.x];
            if (threadIdx.x == REC) {
                posw[tslot] = make_double4(sh_new[R_POS], sh_new[R_POS + 1], sh_new[R_POS + 2], tpw.w);
                if (tiled) t_pf[centre_off + (tslot - tb)] = make_float4((float)rel_frac(sh_new[R_POS] + s.shift[0], ccen[0]), (float)rel_frac(sh_new[R_POS + 1] + s.shift[1], ccen[1]),
                                                                         (float)rel_frac(sh_new[R_POS + 2] + s.shift[2], ccen[2]), 0.f);
            }
        }
        __syncthreads();
        SWP_MARK(7);
    }
    if (threadIdx.x == 0) acc_out[c0] = acc;
}


// ------------------------------------------------------------------------------------------------
// Chain moves on the device: MoveCreator::chainMove / chainDisplace / chainRotate /
// clusterRotate / clusterCM (scOOP/mc/movecreator.cpp:304-328, 1075-1256, 1308-1392) as checkerboard passes.
// A trial picks a uniformly random particle of the active cell and moves the WHOLE molecule it belongs to: with p = 1/2 a
// displacement of fixed length chainm[molType].mx in a uniform direction, else a rotation by chainr[molType].angle * u about a
// uniform axis through the volume-weighted centre (all ten vectors of every member turn, positions in real units, exactly as
// clusterRotate does). Energy before and after is mol2others (totalenergycalculator.h:436-494): every member against every
// NON-member, with an empty connectivity list; the intramolecular terms do not change under a rigid move. Acceptance is
// moveTry. The independence argument of the checkerboard needs every member of the molecule inside the active cell before
// AND after the move, otherwise the trial is rejected (the grid is re-drawn with a random shift every sweep, so every
// molecule that fits into a cell is mobile). A molecule is picked with probability (members in the cell) / (cell population)
// in both directions of a move -- detailed balance holds; particles of one-particle molecules make the trial a no-op
// (counted in scgpu_chainstats::noop).
// Chain passes are separate launches between the single-particle passes of the same sweep: a composition of moves that
// each satisfy detailed balance. Validated by the energy-drift identity, rigid-body invariants and against <E> of the
// reference's sequential sweeps with chainprob > 0.
// ------------------------------------------------------------------------------------------------
constexpr int CH_MAX = 20;          // MAXCHL (scOOP/structures/macros.h:62)
constexpr int CH_THREADS = 128;
constexpr int CH_MAXMT = 32;        // molecule types with their own chain step sizes

struct ChainParams {
    double temper;
    double trials_per_particle;     // chainprob * n_sub: a cell performs (this x its population) trials, stochastically rounded
    double chainm_mx[CH_MAXMT];     // stat.chainm[molType].mx (= 2 * chainmmx, sim.h:366)
    double chainr_angle[CH_MAXMT];  // stat.chainr[molType].angle (radians, sim.h:362)
};

__global__ void __launch_bounds__(CH_THREADS)
k_sweep_chain_colour(DevSys s, ChainParams cp, unsigned long long seed, unsigned long long sweep, int colour, int3 ncol,
                     double4* posw, double* rec, SweepAcc* acc_out) {
    __shared__ double sh_old[CH_MAX][REC], sh_new[CH_MAX][REC];
    __shared__ int sh_mslot[CH_MAX], sh_mtype[CH_MAX];
    __shared__ double sh_red[2][CH_THREADS / 32];
    __shared__ int sh_b[28], sh_off[28];
    __shared__ double sh_u[8];
    __shared__ double sh_rot[9], sh_cm[3];
    __shared__ int sh_ok, sh_accept;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
        if (lane < 28) { sh_b[lane] = b; sh_off[lane] = len; }
    }
    __syncthreads();
    int ntrial;
    {
        const uint4 r = philox4x32((uint32_t)sweep, (uint32_t)(sweep >> 32) ^ ((uint32_t)(colour | 8) << 28), (uint32_t)c0, 0xffffffffu, (uint32_t)seed, (uint32_t)(seed >> 32));
        // proportional to the cell population (constant during a pass: no particle leaves its cell), so that cells holding only
        // one-particle molecules do not burn trials on no-ops
        const double avg = cp.trials_per_particle * (double)npart;
        const double fl = floor(avg);
        ntrial = (int)fl + (u01(r.x, r.y) < avg - fl ? 1 : 0);
    }
    ConList cl;        // mol2others evaluates with an EMPTY connectivity list (totalenergycalculator.h:442-446)
    cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0;
    for (int trial = 0; trial < ntrial; trial++) {
        __syncthreads();
        if (threadIdx.x < 3) {      // Philox counter = (sweep, colour | chain flag, cell, 3 * trial + k)
            const uint4 r = philox4x32((uint32_t)sweep, (uint32_t)(sweep >> 32) ^ ((uint32_t)(colour | 8) << 28), (uint32_t)c0, (uint32_t)(3 * trial + threadIdx.x),
                                       (uint32_t)seed, (uint32_t)(seed >> 32));
            sh_u[2 * threadIdx.x] = u01(r.x, r.y);
            sh_u[2 * threadIdx.x + 1] = u01(r.z, r.w);
        }
        __syncthreads();
        int pick = tb + (int)(sh_u[0] * npart);
        if (pick >= te) pick = te - 1;
        const double4 ppw = posw[pick];
        const int moltype = w_moltype(ppw.w);
        const scgpu_molparam& mpar = s.mol[moltype];
        const int m = (int)mpar.mol_size;
        if (m <= 1 || m > CH_MAX) { if (threadIdx.x == 0) acc.pad++; continue; }        // block-uniform: not a chain (or longer than the reference allows)
        const int mfirst = (int)mpar.first + ((w_orig(ppw.w) - (int)mpar.first) / m) * m;
        const bool displace = sh_u[1] < 0.5;
        const double u_acc = sh_u[1] < 0.5 ? 2.0 * sh_u[1] : 2.0 * sh_u[1] - 1.0;
        // ---- members: records into shared memory (old and new copy)
        if (threadIdx.x < m) {
            const int sl = s.slot_of[mfirst + threadIdx.x];
            sh_mslot[threadIdx.x] = sl;
            sh_mtype[threadIdx.x] = w_type(posw[sl].w);
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < m * REC; idx += blockDim.x) {
            const int k = idx / REC, f = idx % REC;
            const double v = rec[(size_t)sh_mslot[k] * REC + f];
            sh_old[k][f] = v; sh_new[k][f] = v;
        }
        __syncthreads();
        // ---- proposal
        if (threadIdx.x == 0) {
            const double z = 1.0 - 2.0 * sh_u[2], phi = 6.283185307179586476925 * sh_u[3];
            const double rr = sqrt(fmax(0.0, 1.0 - z * z));
            const v3 ax3 = mk(rr * cos(phi), rr * sin(phi), z);
            if (displace) {           // chainDisplace (movecreator.cpp:1091-1094)
                const double mx = cp.chainm_mx[moltype];
                sh_cm[0] = ax3.x * mx / s.box[0]; sh_cm[1] = ax3.y * mx / s.box[1]; sh_cm[2] = ax3.z * mx / s.box[2];
            } else {                  // clusterCM (:1323-1337) + the quaternion of clusterRotate (:1347-1353)
                double cmx = 0.0, cmy = 0.0, cmz = 0.0, vol = 0.0;
                for (int k = 0; k < m; k++) {
                    const double v = s.ia[sh_mtype[k] * s.ntypes + sh_mtype[k]].reserved[2];
                    cmx += sh_old[k][R_POS] * v; cmy += sh_old[k][R_POS + 1] * v; cmz += sh_old[k][R_POS + 2] * v;
                    vol += v;
                }
                sh_cm[0] = cmx / vol; sh_cm[1] = cmy / vol; sh_cm[2] = cmz / vol;
                double d[9];
                rotation_coefficients(d, cp.chainr_angle[moltype] * sh_u[4], ax3, sh_u[5] < 0.5);
                for (int k = 0; k < 9; k++) sh_rot[k] = d[k];
            }
        }
        __syncthreads();
        if (displace) {
            if (threadIdx.x < 3 * m) sh_new[threadIdx.x / 3][R_POS + threadIdx.x % 3] += sh_cm[threadIdx.x % 3];
        } else {
            for (int idx = threadIdx.x; idx < 10 * m; idx += blockDim.x) {
                const int k = idx / 10, v = idx % 10;
                if (v < 9) rotate_vector(&sh_new[k][3 * v], sh_rot);       // dir, patchdir[2], patchsides[4], chdir[2]: all of them, as clusterRotate does
                else {
                    double p[3] = {(sh_old[k][R_POS] - sh_cm[0]) * s.box[0], (sh_old[k][R_POS + 1] - sh_cm[1]) * s.box[1], (sh_old[k][R_POS + 2] - sh_cm[2]) * s.box[2]};
                    rotate_vector(p, sh_rot);
                    sh_new[k][R_POS] = p[0] / s.box[0] + sh_cm[0]; sh_new[k][R_POS + 1] = p[1] / s.box[1] + sh_cm[1]; sh_new[k][R_POS + 2] = p[2] / s.box[2] + sh_cm[2];
                }
            }
        }
        if (threadIdx.x == 0) sh_ok = 1;
        __syncthreads();
        if (threadIdx.x < 2 * m) {      // every member inside the active cell, before and after
            const double* r = (threadIdx.x < m) ? sh_old[threadIdx.x] : sh_new[threadIdx.x - m];
            if (cell_index(r + R_POS, s.shift, s.nc) != c0) sh_ok = 0;
        }
        __syncthreads();
        const bool in_cell = sh_ok != 0;
        double eo = 0.0, en = 0.0;
        if (in_cell) {
            for (int seg = 0; seg < ncell_nb; seg++) {
                const int b = sh_b[seg], len = sh_off[seg];
                for (int i = threadIdx.x; i < len; i += blockDim.x) {
                    const int slot = b + i;
                    const double4 pw = posw[slot];
                    const int orig = w_orig(pw.w);
                    if (orig >= mfirst && orig < mfirst + m) continue;
                    const v3 pc = mk(pw.x, pw.y, pw.z);
                    const int type2 = w_type(pw.w);
                    for (int k = 0; k < m; k++) {
                        v3 r = image(s.box, ld3(&sh_old[k][R_POS]), pc);
                        double d = dot(r, r);
                        if (d <= s.sqmaxcut) eo += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_old[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
                        r = image(s.box, ld3(&sh_new[k][R_POS]), pc);
                        d = dot(r, r);
                        if (d <= s.sqmaxcut) en += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_new[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
                    }
                }
            }
        }
        eo = warp_sum(eo); en = warp_sum(en);
        if (lane == 0) { sh_red[0][wid] = eo; sh_red[1][wid] = en; }
        __syncthreads();
        if (threadIdx.x == 0) {
            bool accept = false;
            double de = 0.0;
            if (in_cell) {
                double a = 0.0, b2 = 0.0;
                for (int k = 0; k < CH_THREADS / 32; k++) { a += sh_red[0][k]; b2 += sh_red[1][k]; }      // fixed order
                de = b2 - a;
                accept = (de <= 0.0) || (exp(-de / cp.temper) > u_acc);       // moveTry (movecreator.h:175-187)
            } else acc.cell_rej++;
            if (accept) acc.de += de;
            if (displace) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
            else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
            sh_accept = accept ? 1 : 0;
        }
        __syncthreads();
        if (sh_accept) {
            for (int idx = threadIdx.x; idx < m * REC; idx += blockDim.x) {
                const int k = idx / REC, f = idx % REC;
                rec[(size_t)sh_mslot[k] * REC + f] = sh_new[k][f];
            }
            if (threadIdx.x < m) {
                const int sl = sh_mslot[threadIdx.x];
                const double w = posw[sl].w;
                posw[sl] = make_double4(sh_new[threadIdx.x][R_POS], sh_new[threadIdx.x][R_POS + 1], sh_new[threadIdx.x][R_POS + 2], w);
            }
        }
    }
    if (threadIdx.x == 0) acc_out[c0] = acc;
}
