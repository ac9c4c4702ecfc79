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
    int trial_rule;                 // 0: the same number of trials in every non-empty cell; 1: trials of a cell = n_sub x its population
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

#ifndef SW_MINBLOCKS
#define SW_MINBLOCKS 8
#endif
constexpr int SW_WARPS = 1;       // one warp per active cell: no block barrier anywhere in a trial
#ifndef SW_TILE_N
#define SW_TILE_N 576
#endif
constexpr int SW_TILE_MAX = 6144;   // largest staged tile the host may ask for (192 KB of dynamic shared memory: one block per SM)
constexpr int SW_TILE = SW_TILE_N;      // staged neighbourhood: FP32 position + direction + slot, 32 B per candidate = 18 KB (8 warps per SM resident)
constexpr int SW_BATCH = 32;      // trials whose random numbers and proposal geometry are prepared together, one lane each
constexpr int SW_MAXROWS = 49;    // (2K+1)^2 rows of neighbour cells, K <= 3
constexpr int SW_MAXT = 8;        // particle types whose reach / cutoff tables are kept in shared memory
constexpr int SW_VLP = 8;         // particles of the active cell that get a partner list of their own ...
constexpr int SW_VLC = 96;        // ... of at most this many candidates (anything beyond falls back to the full scan)

struct SweepGrid {                // fine checkerboard: cells of edge >= maxcut / K, K + 1 colours per axis, active cells K cells apart
    int k[3];                     // per axis: half-width of the neighbourhood in cells (K, or 0 where the axis is a single cell)
    int ncol[3];                  // per axis: colours (K + 1, or 1)
};

struct SweepProposal {            // everything of a trial that does not depend on the outcome of earlier trials
    int slot, displace;
    double u_acc;
    double m[9];                  // displacement (m[0..2], box-fractional) or the rotation coefficients d1..d9 of pscRotate
};

#ifdef SW_PROFILE      // debug build only: cycles per phase of a trial, summed over blocks (lane 0's view)
__device__ unsigned long long sw_prof[16];
#define SWP_MARK(k) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&sw_prof[k], (unsigned long long)(t_ - swp_t)); swp_t = t_; } } while (0)
#else
#define SWP_MARK(k) do { } while (0)
#endif

// the never-taken overflow path of the patch list: out of line, so that it does not bloat the kernel's instruction footprint
__device__ __noinline__ double pair_energy_patch_outofline(const scgpu_iaparam& ia, const v3& r_cm, const double* s1, const double* s2) {
    return pair_energy_patch(ia, r_cm, s1, s2);
}

// One WARP per ACTIVE cell of the current colour, sequential trials inside it, no block barrier anywhere.
// The grid is FINE: cells of edge >= maxcut / K (K = 1, 2, 3, chosen by the host from the mean cell population), K + 1 colours
// per axis, so that two active cells are K cells (>= maxcut) apart and the (2K+1)^3 cells around an active cell hold every
// partner of its particles. Compared with the coarse grid (K = 1) a pass has (2K/(K+1))^3 times more independent cells in
// flight and a sweep (K+1)^3 N / cells serial trials per cell instead of 8 N / cells -- the length of that serial chain is what
// a sweep costs, because a single trial is bound by the latency of its own FP64 dependency chain.
// Per trial: the staged neighbourhood (FP32 position and direction of every candidate) is scanned by the 32 lanes for the old AND
// the new state: centre distance against the exact reach of the type pair, then the segment lower bound above; the few
// survivors are evaluated in FP64 on different lanes (old and new state side by side), patch terms with two lanes per term.
// ONE (with RODS): a single particle type is present -> its interaction-table entry is a kernel parameter (constant bank).
template <bool RODS, bool ONE>
__global__ void __launch_bounds__(32, SW_MINBLOCKS)
k_sweep_cells(DevSys s, SweepParams sp, unsigned long long seed, unsigned long long sweep, int colour, SweepGrid g,
              double4* posw, double* rec, SweepAcc* acc_out, int tile_cap, int* max_c, const __grid_constant__ scgpu_iaparam ia1) {
    // the staged neighbourhood lives in dynamic shared memory: tile_cap candidates (SW_TILE by default; the host asks for more where
    // the neighbourhoods are large -- a lipid membrane on a grid set by a few long rods -- so that those cells keep the staged path
    // and its per-particle partner lists instead of scanning global memory in every trial)
    extern __shared__ __align__(16) unsigned char sw_dyn[];
    float4* t_pf = reinterpret_cast<float4*>(sw_dyn);          // x, y, z: box-fractional position relative to the cell centre; w: original index | type << 24
    float4* t_df = t_pf + tile_cap;                            // direction; w: slot
    __shared__ double sh_old[REC], sh_new[REC];
    __shared__ int sh_queue[128];
    __shared__ int sh_pl[32];                 // (tile entry, state) pairs that owe a patch evaluation
    __shared__ int sh_b[2 * SW_MAXROWS + 2], sh_off[2 * SW_MAXROWS + 2];
    __shared__ SweepProposal sh_prop[SW_BATCH];
    __shared__ float sh_tab[2 * SW_MAXT * SW_MAXT + SW_MAXT];
    __shared__ unsigned short sh_vl[SW_VLP][SW_VLC];
    __shared__ int sh_vn[SW_VLP];
    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    // active cell of this warp
    const int ax = s.nc[0] / g.ncol[0], ay = s.nc[1] / g.ncol[1];
    const int bx = blockIdx.x % ax, by = (blockIdx.x / ax) % ay, bz = blockIdx.x / (ax * ay);
    const int cx = bx * g.ncol[0] + (colour % g.ncol[0]), cy = by * g.ncol[1] + ((colour / g.ncol[0]) % g.ncol[1]), cz = bz * g.ncol[2] + (colour / (g.ncol[0] * g.ncol[1]));
    const int c0 = (cz * s.nc[1] + cy) * s.nc[0] + cx;
    const int tb = s.cell_start[c0], te = s.cell_start[c0 + 1];
    const int npart = te - tb;
    SweepAcc acc = {0, 0, 0, 0, 0, 0, 0.0};
    if (npart == 0) { if (lane == 0) acc_out[c0] = acc; return; }
#ifdef SW_PROFILE
    const long long swp_t0 = clock64();
#endif
    // ---- the neighbourhood: (2ky+1)(2kz+1) rows of cells, each row one contiguous slot range [cx-kx, cx+kx] or two where it wraps
    const int wy = 2 * g.k[1] + 1, wz = 2 * g.k[2] + 1, nrows = wy * wz;
    const int T = s.ntypes;
    for (int part = 0; part < 2; part++) {
        int b = 0, len = 0;
        for (int r0 = 0; r0 < nrows; r0 += 32) {
            const int r = r0 + lane;
            b = 0; len = 0;
            if (r < nrows) {
                const int yy = (cy + r % wy - g.k[1] + s.nc[1]) % s.nc[1], zz = (cz + r / wy - g.k[2] + s.nc[2]) % s.nc[2];
                const int rbase = (zz * s.nc[1] + yy) * s.nc[0];
                const int lo = cx - g.k[0], hi = cx + g.k[0];
                int a0, a1;
                if (lo < 0) { if (part == 0) { a0 = 0; a1 = hi; } else { a0 = lo + s.nc[0]; a1 = s.nc[0] - 1; } }
                else if (hi >= s.nc[0]) { if (part == 0) { a0 = lo; a1 = s.nc[0] - 1; } else { a0 = 0; a1 = hi - s.nc[0]; } }
                else { a0 = part == 0 ? lo : 1; a1 = part == 0 ? hi : 0; }
                if (a0 <= a1) { b = s.cell_start[rbase + a0]; len = s.cell_start[rbase + a1 + 1] - b; }
                sh_b[2 * r + part] = b; sh_off[2 * r + part] = len;      // lengths first, offsets below
            }
        }
    }
    __syncwarp();
    int C = 0;
    {   // exclusive scan of the 2 * nrows segment lengths (at most 98: four per lane)
        int v[4], tot = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) { const int k = 4 * lane + u; v[u] = k < 2 * nrows ? sh_off[k] : 0; tot += v[u]; }
        int x = tot;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        C = __shfl_sync(0xffffffffu, x, 31);
        int run = x - tot;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; u++) { const int k = 4 * lane + u; if (k < 2 * nrows) sh_off[k] = run; run += v[u]; }
        if (lane == 0) sh_off[2 * nrows] = C;
    }
    const bool tabs = !(RODS && ONE) && T <= SW_MAXT;       // reach / cutoff / half-length tables in shared memory
    if (tabs) for (int k = lane; k < 2 * T * T + T; k += 32) sh_tab[k] = k < T * T ? s.reach2[k] : s.reach2[k + T];
    __syncwarp();
    const int nseg = 2 * nrows;
    const bool tiled = C <= tile_cap;       // denser neighbourhoods are scanned from global memory (slower, same results)
    if (lane == 0) atomicMax(max_c, C);     // the host sizes the tile of the NEXT sweep by what this one met
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    auto slot_of_p = [&](int p) {
        int lo = 0, hi = nseg - 1;             // the last segment whose offset is <= p (offsets are non-decreasing; empty segments repeat them)
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sh_off[mid] <= p) lo = mid; else hi = mid - 1; }
        return sh_b[lo] + (p - sh_off[lo]);
    };
    auto staged_pos = [&](const double4& pw) {
        return make_float4((float)rel_frac(pw.x + s.shift[0], ccen[0]), (float)rel_frac(pw.y + s.shift[1], ccen[1]),
                           (float)rel_frac(pw.z + s.shift[2], ccen[2]), __int_as_float(w_orig(pw.w) | (w_type(pw.w) << 24)));
    };
    auto staged_dir = [&](int slot) {
        const double4 d = ldg256(rec + (size_t)slot * REC + R_DIR);
        return make_float4((float)d.x, (float)d.y, (float)d.z, __int_as_float(slot));
    };
    if (tiled) {
        // slots first (shared memory only), then one flat pass with every lane busy and two candidates in flight per lane: the
        // segments are short (a few cells each), a pass per segment would expose one L2 round trip per segment
        for (int k = 0; k < nseg; k++) {
            const int b = sh_b[k], off = sh_off[k], len = sh_off[k + 1] - off;
            for (int idx = lane; idx < len; idx += 32) t_df[off + idx].w = __int_as_float(b + idx);
        }
        __syncwarp();
        for (int p0 = 0; p0 < C; p0 += 64) {
            const int pa = p0 + lane, pb = p0 + 32 + lane;
            const int sa = pa < C ? __float_as_int(t_df[pa].w) : tb, sb = pb < C ? __float_as_int(t_df[pb].w) : tb;
            const double4 wa = posw[sa], wb = posw[sb];
            const float4 da = staged_dir(sa), db = staged_dir(sb);
            if (pa < C) { t_pf[pa] = staged_pos(wa); t_df[pa] = da; }
            if (pb < C) { t_pf[pb] = staged_pos(wb); t_df[pb] = db; }
        }
    }
    __syncwarp();          // the staged tile is read by other lanes from here on (racecheck: the partner lists below)
    // the active cell is the centre of its own neighbourhood: where its particles sit in the staged tile
    const int centre_seg = 2 * (g.k[2] * wy + g.k[1]);
    const int centre_off = sh_off[centre_seg] + (tb - sh_b[centre_seg]);
    // trials of this cell: the same number in every non-empty cell (rule 0: n_sub * N / non-empty cells; all cells of a pass finish
    // together instead of waiting for the fullest one) or n_sub x its population (rule 1: every particle is picked once per sweep on
    // average, the reference's rates). Either count does not depend on anything a trial can change (particles never leave their
    // cell within a pass), so detailed balance holds. Fractional counts are stochastically rounded.
    int ntrial;
    {
        const double avg = sp.trial_rule >= 1 ? sp.trial_scale * (double)sp.n_sub * (double)npart
                                              : sp.trial_scale * (double)sp.n_sub * (double)s.n / (double)s.cell_start[s.ncells + 1];
        const uint4 r = philox4x32((uint32_t)sweep, (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 24), (uint32_t)c0, 0xffffffffu, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double fl = floor(avg);
        ntrial = (int)fl + (u01(r.x, r.y) < avg - fl ? 1 : 0);
    }
    // ---- partner lists of the cell's own particles, built once per pass: a candidate can only become a partner during the pass if
    // it is within reach + skin now, skin = the farthest the particle itself can travel (ntrial displacements of fixed length;
    // candidates outside the cell do not move during the pass, those inside are few and the same skin covers them twice over:
    // 2 * skin). A trial then scans its particle's list (a few dozen entries) instead of the whole neighbourhood.
    if (tiled) {
        for (int a = 0; a < SW_VLP && a < npart; a++) {
            const int ia = centre_off + a;
            const float4 me = t_pf[ia];
            const int mytype = __float_as_int(me.w) >> 24;
            float tmax = 0.f;
            for (int t = 0; t < T; t++) tmax = fmaxf(tmax, (float)sp.trans_mx[t]);
            const float skin = 2.f * tmax * (float)ntrial * 1.0001f + 1e-3f;
            const float* rrow = (RODS && ONE) ? nullptr : (tabs ? sh_tab + mytype * T : s.reach2 + mytype * T);
            const float r_one = (RODS && ONE) ? (float)(ia1.reserved[1] * 1.001) : 0.f;
            int cnt = 0;
            for (int base = 0; base < C; base += 32) {
                const int p = base + lane;
                bool keep = false;
                if (p < C && p != ia) {
                    const float4 q = t_pf[p];
                    float dx = me.x - q.x, dy = me.y - q.y, dz = me.z - q.z;
                    dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                    const float rr = (RODS && ONE) ? r_one : rrow[__float_as_int(q.w) >> 24];
                    const float lim = sqrtf(rr) + skin;
                    keep = dx * dx + dy * dy + dz * dz <= lim * lim;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int at = cnt + __popc(m & lt_mask);
                if (keep && at < SW_VLC) sh_vl[a][at] = (unsigned short)p;
                cnt += __popc(m);
            }
            if (lane == 0) sh_vn[a] = cnt <= SW_VLC ? cnt : -1;
        }
    }
    __syncwarp();
#ifdef SW_PROFILE
    if (lane == 0) { atomicAdd(&sw_prof[15], (unsigned long long)ntrial); atomicAdd(&sw_prof[14], 1ull); atomicAdd(&sw_prof[8], (unsigned long long)(clock64() - swp_t0)); }
    long long swp_t = clock64();
#endif
    __syncwarp();
    for (int trial = 0; trial < ntrial; trial++) {
        const int bi = trial % SW_BATCH;
        if (bi == 0) {
            // ---- random numbers and proposal geometry of the next SW_BATCH trials, one lane per trial.
            // Philox counter = (sweep, colour, cell, 3*trial + k); nothing here depends on earlier acceptances.
            __syncwarp();
            const int tr = trial + lane;
            if (tr < ntrial) {
                double u[6];
                const uint32_t c1 = (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 24);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    uint4 r = philox4x32((uint32_t)sweep, c1, (uint32_t)c0, (uint32_t)(3 * tr + k), (uint32_t)seed, (uint32_t)(seed >> 32));
                    u[2 * k] = u01(r.x, r.y);
                    u[2 * k + 1] = u01(r.z, r.w);
                }
                int pick = tb + (int)(u[0] * npart);          // uniformly chosen particle of this cell, with replacement
                if (pick >= te) pick = te - 1;
                const int ty = w_type(posw[pick].w);
                const int gt = sp.geotype_of_type[ty];
                const bool displace = (gt >= SCGPU_SPN) || (u[1] < 0.5);                 // particleMove (movecreator.cpp:11-33)
                const double z = 1.0 - 2.0 * u[2], phi = 6.283185307179586476925 * u[3];
                const double rr = sqrt(fmax(0.0, 1.0 - z * z));
                const v3 ax3 = mk(rr * cos(phi), rr * sin(phi), z);                     // uniform on the unit sphere
                SweepProposal& P = sh_prop[lane];
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
            __syncwarp();
        }
        SWP_MARK(0);
        const int tslot = sh_prop[bi].slot;
        const bool displace = sh_prop[bi].displace != 0;
        const double u_acc = sh_prop[bi].u_acc;
        { const double v = rec[(size_t)tslot * REC + lane]; sh_old[lane] = v; sh_new[lane] = v; }       // REC == 32 == the warp
        const double4 tpw = posw[tslot];
        const int target = w_orig(tpw.w), type1 = w_type(tpw.w), moltype1 = w_moltype(tpw.w);
        __syncwarp();
        if (displace) {
            if (lane < 3) sh_new[R_POS + lane] += sh_prop[bi].m[lane];
        } else if (lane < 9) {      // one lane per vector of the record
            const int gt = sp.geotype_of_type[type1];
            if (record_vector_rotates(gt, lane)) rotate_vector(sh_new + 3 * lane, sh_prop[bi].m);      // R_DIR, R_PD0, R_S0, R_S1, R_PD1, R_S2, R_S3, R_CH0, R_CH1
        }
        __syncwarp();
        SWP_MARK(1);
        // a move that leaves the cell would break the independence of the active cells: reject it
        bool in_cell = cell_index(sh_new + R_POS, s.shift, s.nc) == c0;
        ConList cl;
        if (RODS) { cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0; }
        else {
            get_conlist(s.mol, moltype1, target, cl);
            // a bonded partner is evaluated wherever it is; if it sits in ANOTHER active cell it may be moving right now: reject
            // (the test depends only on the partner's position, which this trial does not change: detailed balance holds)
            if (!cl.is_empty && in_cell) {
                bool clash = false;
                if (lane < 4 && cl.con[lane] >= 0) {
                    const double4 pw = posw[s.slot_of[cl.con[lane]]];
                    const int px = cell_coord(pw.x + s.shift[0], s.nc[0]), py = cell_coord(pw.y + s.shift[1], s.nc[1]), pz = cell_coord(pw.z + s.shift[2], s.nc[2]);
                    const bool active = px % g.ncol[0] == colour % g.ncol[0] && py % g.ncol[1] == (colour / g.ncol[0]) % g.ncol[1] && pz % g.ncol[2] == colour / (g.ncol[0] * g.ncol[1]);
                    clash = active && !(px == cx && py == cy && pz == cz);
                }
                if (__any_sync(0xffffffffu, clash)) in_cell = false;
            }
        }
        double e_old = 0.0, e_new = 0.0;
        if (in_cell) {
            const v3 po = ld3(sh_old + R_POS), pn = ld3(sh_new + R_POS);
            const float ox = (float)rel_frac(po.x + s.shift[0], ccen[0]), oy = (float)rel_frac(po.y + s.shift[1], ccen[1]), oz = (float)rel_frac(po.z + s.shift[2], ccen[2]);
            const float nxf = (float)rel_frac(pn.x + s.shift[0], ccen[0]), nyf = (float)rel_frac(pn.y + s.shift[1], ccen[1]), nzf = (float)rel_frac(pn.z + s.shift[2], ccen[2]);
            const float dox = (float)sh_old[R_DIR], doy = (float)sh_old[R_DIR + 1], doz = (float)sh_old[R_DIR + 2];
            const float dnx = (float)sh_new[R_DIR], dny = (float)sh_new[R_DIR + 1], dnz = (float)sh_new[R_DIR + 2];
            const float* reach_row = tabs ? sh_tab + type1 * T : s.reach2 + type1 * T;
            const float* cut_row = tabs ? sh_tab + T * T + type1 * T : s.reach2 + T * T + T + type1 * T;
            const float* hl_tab = tabs ? sh_tab + 2 * T * T : s.reach2 + 2 * T * T + T;
            const float h1 = (RODS && ONE) ? (float)ia1.half_len[0] : hl_tab[type1];
            const float reach_one = (float)(ia1.reserved[1] * 1.001), cut_one = (float)(fmax(ia1.rcutSq, ia1.rcutwcaSq) * 1.001);
            int qn = 0;
            double lo = 0.0, ln = 0.0;
            // one queue entry = (candidate, which state): the old and the new state of a trial are evaluated on DIFFERENT lanes.
            // Phase A: exact gate + everything but the rod-rod patch term; entries that owe a patch term are collected.
            // Phase B: the patch terms, two lanes per term.
            int pc = 0;
            auto entry_slot = [&](int entry) { const int p = entry >> 1; return (tiled && p >= 0) ? __float_as_int(t_df[p].w) : (p >= 0 ? slot_of_p(p) : -1 - p); };
            auto patch_entry = [&](int entry) {
                const int slot = entry_slot(entry);
                const bool is_new = entry & 1;
                double4 pw = posw[slot];
                v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                double e = pair_energy_patch_outofline(ONE ? ia1 : s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC);
                if (is_new) ln += e; else lo += e;
            };
            auto eval = [&](int entry, bool on) {
                bool np = false;
                if (on) {
                    const int slot = entry_slot(entry);
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
                        if (np) sh_pl[pc + __popc(m & lt_mask)] = entry;
                        pc += c;
                    }
                    __syncwarp();
                }
            };
            const int va = tslot - tb;
            const int vn = (tiled && va < SW_VLP) ? sh_vn[va] : -1;
            const int nscan = vn >= 0 ? vn : C;
            for (int base = 0; base < nscan; base += 32) {
                int p = base + lane;
                bool pass_o = false, pass_n = false;
                if (p < nscan) {
                    if (vn >= 0) p = sh_vl[va][p];
                    float4 q, qd;
                    if (tiled) { q = t_pf[p]; qd = t_df[p]; }
                    else { const int sl = slot_of_p(p); q = staged_pos(posw[sl]); qd = staged_dir(sl); }
                    const int wbits = __float_as_int(q.w);
                    const int orig = wbits & 0xffffff, ctype = wbits >> 24;
                    if (orig != target) {
                        float dx = ox - q.x, dy = oy - q.y, dz = oz - q.z;
                        dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                        float ex = nxf - q.x, ey = nyf - q.y, ez = nzf - q.z;
                        ex = (ex - rintf(ex)) * boxf[0]; ey = (ey - rintf(ey)) * boxf[1]; ez = (ez - rintf(ez)) * boxf[2];
                        const float d2o = dx * dx + dy * dy + dz * dz, d2n = ex * ex + ey * ey + ez * ez;
                        const float reach = (RODS && ONE) ? reach_one : reach_row[ctype];
                        pass_o = d2o <= reach;
                        pass_n = d2n <= reach;
                        const float cut2 = (RODS && ONE) ? cut_one : cut_row[ctype];
                        if (cut2 > 0.f && (pass_o || pass_n)) {     // rod pair: the segment bound
                            const float h2 = (RODS && ONE) ? h1 : hl_tab[ctype];
                            if (pass_o && lb_beyond(dx, dy, dz, d2o, dox, doy, doz, qd.x, qd.y, qd.z, h1, h2, cut2)) pass_o = false;
                            if (pass_n && lb_beyond(ex, ey, ez, d2n, dnx, dny, dnz, qd.x, qd.y, qd.z, h1, h2, cut2)) pass_n = false;
                        }
                        if (!RODS && !cl.is_empty) {       // bonded partners are evaluated by index below
                            if (orig == cl.con[0] || orig == cl.con[1] || orig == cl.con[2] || orig == cl.con[3]) { pass_o = false; pass_n = false; }
                        }
                    }
                }
                unsigned mo = __ballot_sync(0xffffffffu, pass_o), mn = __ballot_sync(0xffffffffu, pass_n);
                int no = __popc(mo);
                if (pass_o) sh_queue[qn + __popc(mo & lt_mask)] = p * 2;
                if (pass_n) sh_queue[qn + no + __popc(mn & lt_mask)] = p * 2 + 1;
                qn += no + __popc(mn);
                __syncwarp();
                while (qn >= 32) {
                    eval(sh_queue[lane], true);
                    int rest = qn - 32;
                    int mv0 = (lane < rest) ? sh_queue[32 + lane] : 0;
                    int mv1 = (lane + 32 < rest) ? sh_queue[64 + lane] : 0;
                    __syncwarp();
                    if (lane < rest) sh_queue[lane] = mv0;
                    if (lane + 32 < rest) sh_queue[32 + lane] = mv1;
                    qn = rest;
                    __syncwarp();
                }
            }
            if (qn > 0) eval(lane < qn ? sh_queue[lane] : 0, lane < qn);
            if (!RODS && !cl.is_empty) {      // bonded partners by index, both states: entry = (-1 - slot) * 2 + state
                bool on = lane < 8 && cl.con[lane >> 1] >= 0;
                eval(on ? (-1 - s.slot_of[cl.con[lane >> 1]]) * 2 + (lane & 1) : 0, on);
            }
            SWP_MARK(3);
            for (int base = 0; base < 2 * pc; base += 32) {      // phase B: the patch terms, two lanes per term, in list order
                const int idx = (base + lane) >> 1;
                const bool act = idx < pc;
                const int entry = act ? sh_pl[idx] : 0;
                const int slot = act ? entry_slot(entry) : tslot;
                const bool is_new = entry & 1;
                double4 pw = posw[slot];
                v3 r = image(s.box, is_new ? pn : po, mk(pw.x, pw.y, pw.z));
                double e = pair_energy_patch_two_lanes(ONE ? ia1 : s.ia[type1 * s.ntypes + w_type(pw.w)], r, is_new ? sh_new : sh_old, rec + (size_t)slot * REC, act);
                if (is_new) ln += e; else lo += e;
            }
            SWP_MARK(5);
            e_old = warp_sum(lo);
            e_new = warp_sum(ln);
            if (s.wall != nullptr) {       // [EXTER] wall: added to both energies of the trial, as oneToAll / oneToAllTrial do (totalenergycalculator.h:377-378, 410-411)
                double wv = 0.0;
                if (lane < 2) wv = wall_energy_rec(s, lane ? sh_new : sh_old, type1);
                e_old += __shfl_sync(0xffffffffu, wv, 0);
                e_new += __shfl_sync(0xffffffffu, wv, 1);
            }
        }
        SWP_MARK(6);
        bool accept = false;
        double de = 0.0;
        if (in_cell) {        // every lane takes the same decision from the same numbers
            de = e_new - e_old;
            accept = (de <= 0.0) || (exp(-de / sp.temper) > u_acc);                    // moveTry (movecreator.h:175-187)
        }
        if (lane == 0) {
            if (accept) acc.de += de;
            if (!in_cell) acc.cell_rej++;
            if (displace) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
            else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
        }
        if (accept) {          // commit in place: sorted record, position word, staged FP32 copies
            rec[(size_t)tslot * REC + lane] = sh_new[lane];
            if (lane == 0) {
                const double4 npw = make_double4(sh_new[R_POS], sh_new[R_POS + 1], sh_new[R_POS + 2], tpw.w);
                posw[tslot] = npw;
                if (tiled) {
                    t_pf[centre_off + (tslot - tb)] = staged_pos(npw);
                    t_df[centre_off + (tslot - tb)] = make_float4((float)sh_new[R_DIR], (float)sh_new[R_DIR + 1], (float)sh_new[R_DIR + 2], __int_as_float(tslot));
                }
            }
        }
        __syncwarp();
        SWP_MARK(7);
    }
    if (lane == 0) acc_out[c0] = acc;
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
#ifndef CH_THREADS_N
#define CH_THREADS_N 512
#endif
constexpr int CH_THREADS = CH_THREADS_N;
constexpr int CH_TILE_MAX = 12288;  // largest staged neighbourhood of the chain kernel (192 KB of dynamic shared memory)
constexpr int CH_MAXMT = 32;        // molecule types with their own chain step sizes

struct ChainParams {
    double temper;
    double trials_per_particle;     // chainprob * n_sub: a cell performs (this x its population) trials, stochastically rounded
    double chainm_mx[CH_MAXMT];     // stat.chainm[molType].mx (= 2 * chainmmx, sim.h:366)
    double chainr_angle[CH_MAXMT];  // stat.chainr[molType].angle (radians, sim.h:362)
};

__global__ void __launch_bounds__(CH_THREADS)
k_sweep_chain_colour(DevSys s, ChainParams cp, unsigned long long seed, unsigned long long sweep, int colour, int3 ncol,
                     double4* posw, double* rec, SweepAcc* acc_out, int tile_cap, int* max_c) {
    // the 27-cell neighbourhood staged once per pass in dynamic shared memory: FP32 position relative to the cell centre (box fractions)
    // and slot | type << 24. A trial tests the staged candidates against the exact reach of the type pairs (+ 1 % and the FP32 slack)
    // for every member in both states and goes to global memory and FP64 only for those that pass -- a lipid bead among 6 000
    // candidates of a grid set by a few long rods has a few dozen partners. Neighbourhoods above tile_cap: the scan of global memory.
    extern __shared__ __align__(16) unsigned char ch_dyn[];
    float4* tile = reinterpret_cast<float4*>(ch_dyn);
    __shared__ float4 sh_fo[CH_MAX], sh_fn[CH_MAX];
    __shared__ int sh_toff[29];
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
        int x = len;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane < 28) sh_toff[lane] = x - len;
        if (lane == 27) sh_toff[28] = x;
    }
    __syncthreads();
    const int Ctot = sh_toff[28];
    const bool staged = Ctot <= tile_cap;
    if (threadIdx.x == 0) atomicMax(max_c, Ctot);      // the host sizes the tile of the next sweep by what this one met
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    auto stage_entry = [&](const double4& pw, int slot) {
        return make_float4((float)rel_frac(pw.x + s.shift[0], ccen[0]), (float)rel_frac(pw.y + s.shift[1], ccen[1]),
                           (float)rel_frac(pw.z + s.shift[2], ccen[2]), __int_as_float(slot | (w_type(pw.w) << 24)));
    };
    int centre_toff = 0;
    for (int seg = 0; seg < ncell_nb; seg++) if (sh_b[seg] == tb && sh_off[seg] == npart) centre_toff = sh_toff[seg];
    if (staged) {
        for (int seg = 0; seg < ncell_nb; seg++) {
            const int b = sh_b[seg], len = sh_off[seg], off = sh_toff[seg];
            for (int i = threadIdx.x; i < len; i += blockDim.x) tile[off + i] = stage_entry(posw[b + i], b + i);
        }
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
            const float4 f = make_float4((float)rel_frac(r[R_POS] + s.shift[0], ccen[0]), (float)rel_frac(r[R_POS + 1] + s.shift[1], ccen[1]),
                                         (float)rel_frac(r[R_POS + 2] + s.shift[2], ccen[2]), 0.f);
            if (threadIdx.x < m) sh_fo[threadIdx.x] = f; else sh_fn[threadIdx.x - m] = f;
        }
        __syncthreads();
        const bool in_cell = sh_ok != 0;
        double eo = 0.0, en = 0.0;
        auto evaluate = [&](int slot) {       // one candidate against every member, old and new state, exact
            const double4 pw = posw[slot];
            const int orig = w_orig(pw.w);
            if (orig >= mfirst && orig < mfirst + m) return;
            const v3 pc = mk(pw.x, pw.y, pw.z);
            const int type2 = w_type(pw.w);
            for (int k = 0; k < m; k++) {
                const double reach = (double)s.reach2[sh_mtype[k] * s.ntypes + type2];
                v3 r = image(s.box, ld3(&sh_old[k][R_POS]), pc);
                double d = dot(r, r);
                if (d <= s.sqmaxcut && d <= reach) eo += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_old[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
                r = image(s.box, ld3(&sh_new[k][R_POS]), pc);
                d = dot(r, r);
                if (d <= s.sqmaxcut && d <= reach) en += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_new[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
            }
        };
        if (in_cell && staged) {
            for (int p = threadIdx.x; p < Ctot; p += blockDim.x) {
                const float4 q = tile[p];
                const int sbits = __float_as_int(q.w);
                const int type2 = sbits >> 24;
                bool any = false;
                for (int k = 0; k < m; k++) {
                    const float lim = s.reach2[sh_mtype[k] * s.ntypes + type2] * 1.01f + 1e-3f;
                    float dx = sh_fo[k].x - q.x, dy = sh_fo[k].y - q.y, dz = sh_fo[k].z - q.z;
                    dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                    any = any || dx * dx + dy * dy + dz * dz <= lim;
                    dx = sh_fn[k].x - q.x; dy = sh_fn[k].y - q.y; dz = sh_fn[k].z - q.z;
                    dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                    any = any || dx * dx + dy * dy + dz * dz <= lim;
                }
                if (any) evaluate(sbits & 0xffffff);
            }
        } else if (in_cell) {
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
                        // beyond the reach of the type pair every term is exactly 0 (the connectivity list is empty here): the cutoff the
                        // reference applies (sqmaxcut, set by the LONGEST interaction of the system) lets a lipid bead call the functor for
                        // every bead within 10 sigma; the exact reach of a bead pair is 2.7
                        const double reach = (double)s.reach2[sh_mtype[k] * s.ntypes + type2];
                        v3 r = image(s.box, ld3(&sh_old[k][R_POS]), pc);
                        double d = dot(r, r);
                        if (d <= s.sqmaxcut && d <= reach) eo += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_old[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
                        r = image(s.box, ld3(&sh_new[k][R_POS]), pc);
                        d = dot(r, r);
                        if (d <= s.sqmaxcut && d <= reach) en += pair_energy_gated(s.box, s.ia, s.ntypes, s.mol, r, d, sh_new[k], sh_mtype[k], moltype, rec + (size_t)slot * REC, type2, orig, cl);
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
                if (s.wall != nullptr)       // [EXTER] wall term of every member (mol2others / mol2othersTrial, totalenergycalculator.h:435-449)
                    for (int k = 0; k < m; k++) { a += wall_energy_rec(s, sh_old[k], sh_mtype[k]); b2 += wall_energy_rec(s, sh_new[k], sh_mtype[k]); }
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
                const double4 npw = make_double4(sh_new[threadIdx.x][R_POS], sh_new[threadIdx.x][R_POS + 1], sh_new[threadIdx.x][R_POS + 2], w);
                posw[sl] = npw;
                if (staged) tile[centre_toff + (sl - tb)] = stage_entry(npw, sl);      // the members sit in the active cell
            }
        }
    }
    if (threadIdx.x == 0) acc_out[c0] = acc;
}
