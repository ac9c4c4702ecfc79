// Checkerboard sweeps, third form: a colour pass as FOUR DENSE LAUNCHES instead of one kernel that does everything per cell.
//
//   k_sweep_propose   block per active cell: every particle of the cell gets one trial (scgpu_moveparams::trial_rule 2: a fresh random
//                     order), the new states are written to SPARE SLOTS behind the N sorted particles, the gate lists every pair term
//                     the pass can need as (slot, slot) pairs in the flat list of the energy pipeline: trial particle (old / new slot)
//                     x partner, and for two trial particles of the same cell the four old / new combinations, flagged conditional;
//   k_cheap_flat      } the energy pipeline's own kernels over that list: a thread per listed pair, patch terms compacted and
//   k_patch_flat      } evaluated in three dense phases -- the trial states are ordinary slots to them;
//   k_sweep_resolve   block per active cell: sums of every trial's unconditional terms in parallel, then ONE warp walks the trials in
//                     order (conditional terms that match the outcome of earlier trials, moveTry) and the accepted states are copied
//                     from the spare slots over the old ones.
//
// The decisions are those of the sequential walk through the cell, as in k_sweep_rounds (sweep_rounds.cuh) -- but there the terms of a
// cell were evaluated by the 128 threads of its block between barriers, here all terms of all cells of the pass (a few hundred
// thousand) feed kernels that run at the throughput of the energy pipeline.
// Eligible: no bonded molecules, trial_rule 2 (n_sub walks = n_sub times the colour passes on the same grid), coarse grid, cells of at
// most SP_TR particles whose neighbourhood fits the staged tile
// (anything else takes k_sweep_rounds).
#pragma once

#ifndef SP_THREADS_N
#define SP_THREADS_N 256
#endif
#ifndef SP_MINB
#define SP_MINB 3
#endif
constexpr int SP_THREADS = SP_THREADS_N;      // k_sweep_propose: the grid is only as large as the number of active cells (a few hundred), so the warps must come from the block
#ifndef SP_RTHREADS_N
#define SP_RTHREADS_N 128
#endif
constexpr int SP_RTHREADS = SP_RTHREADS_N;               // k_sweep_resolve
constexpr int SP_TILE = 768;           // staged neighbourhood (FP32 position + direction + slot, 32 B per candidate)
constexpr int SP_TR = 64;              // trials (= particles) per cell and pass
constexpr int SP_EB = 192;             // pair terms of one trial
constexpr int SP_HB = 512;             // centre-distance hits a warp collects before it runs the segment bound over them
constexpr int SP_CPOOL = 1536;         // conditional terms of a cell kept in shared memory by k_sweep_resolve

struct SwTrial {                       // one trial of the pass (global memory, indexed by SwCell::trial_base + order in the cell)
    int span, cnt, ncond;              // its pair terms in the flat list: [span, span + cnt), the first ncond of them conditional
    int slot_old, slot_new;            // sorted slot of the particle, spare slot of its trial state
    int flags;                         // bit 0: the move stays inside the cell (else rejected), bit 1: displacement (else rotation)
    int type, pad;
    double u_acc;
};
struct SwCell { int trial_base, nt; };
struct SweepAux {
    SwTrial* trials; int* trial_total; int trial_cap;
    SwCell* cells;
    unsigned short* meta;              // per listed pair: bit 0 own state (0 old / 1 new), bit 1 partner in its new state, bit 2 conditional,
};                                     // bits 3..: order of the partner's trial in the cell

struct SpShared {
    float4 t_pf[SP_TILE];
    float4 t_df[SP_TILE];
    float4 fo[SP_TR], fdo[SP_TR], fn[SP_TR], fdn[SP_TR];      // the trial particles, old and new state: FP32 position (w: type) and axis
    int slot[SP_TR], valid[SP_TR];
    double recbuf[SP_THREADS / 32][REC];
    int hit[SP_THREADS / 32][SP_HB];
    int2 epair[SP_THREADS / 32][SP_EB];
    unsigned short emeta[SP_THREADS / 32][SP_EB];
    int seg_b[2 * SW_MAXROWS + 2], seg_off[2 * SW_MAXROWS + 2];
    float tab[2 * SW_MAXT * SW_MAXT + SW_MAXT];
    unsigned int key[SP_TR];
    unsigned short perm[SP_TR];
    int tbase, ok;
};

template <bool RODS, bool ONE>
__global__ void __launch_bounds__(SP_THREADS, SP_MINB)
k_sweep_propose(DevSys s, SweepParams sp, unsigned long long seed, unsigned long long sweep, int colour, int sub, SweepGrid g,
                double4* posw, double* rec, FlatList fl, SweepAux ax, SweepAcc* acc_out, const __grid_constant__ scgpu_iaparam ia1) {
    extern __shared__ __align__(16) unsigned char sp_raw[];
    SpShared& S = *reinterpret_cast<SpShared*>(sp_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SP_THREADS / 32;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int ax_ = s.nc[0] / g.ncol[0], ay = s.nc[1] / g.ncol[1];
    const int bx = blockIdx.x % ax_, by = (blockIdx.x / ax_) % ay, bz = blockIdx.x / (ax_ * ay);
    const int cx = bx * g.ncol[0] + (colour % g.ncol[0]), cy = by * g.ncol[1] + ((colour / g.ncol[0]) % g.ncol[1]), cz = bz * g.ncol[2] + (colour / (g.ncol[0] * g.ncol[1]));
    const int c0 = (cz * s.nc[1] + cy) * s.nc[0] + cx;
    const int tb = s.cell_start[c0], te = s.cell_start[c0 + 1];
    const int npart = te - tb;
    if (npart == 0 || *fl.overflow) { if (tid == 0) { ax.cells[c0].trial_base = 0; ax.cells[c0].nt = 0; } return; }
    // ---- the neighbourhood: (2ky+1)(2kz+1) rows of cells, each row one contiguous slot range [cx-kx, cx+kx] or two where it wraps
    const int wy = 2 * g.k[1] + 1, wz = 2 * g.k[2] + 1, nrows = wy * wz, nseg = 2 * nrows;
    const int T = s.ntypes;
    if (tid < nseg) {
        const int r = tid >> 1, part = tid & 1;
        const int yy = (cy + r % wy - g.k[1] + s.nc[1]) % s.nc[1], zz = (cz + r / wy - g.k[2] + s.nc[2]) % s.nc[2];
        const int rbase = (zz * s.nc[1] + yy) * s.nc[0];
        const int lo = cx - g.k[0], hi = cx + g.k[0];
        int a0, a1, b = 0, len = 0;
        if (lo < 0) { if (part == 0) { a0 = 0; a1 = hi; } else { a0 = lo + s.nc[0]; a1 = s.nc[0] - 1; } }
        else if (hi >= s.nc[0]) { if (part == 0) { a0 = lo; a1 = s.nc[0] - 1; } else { a0 = 0; a1 = hi - s.nc[0]; } }
        else { a0 = part == 0 ? lo : 1; a1 = part == 0 ? hi : 0; }
        if (a0 <= a1) { b = s.cell_start[rbase + a0]; len = s.cell_start[rbase + a1 + 1] - b; }
        S.seg_b[tid] = b; S.seg_off[tid] = len;
    }
    const bool tabs = !(RODS && ONE) && T <= SW_MAXT;
    if (tabs) for (int k = tid; k < 2 * T * T + T; k += SP_THREADS) S.tab[k] = k < T * T ? s.reach2[k] : s.reach2[k + T];
    __syncthreads();
    if (wid == 0) {
        int v[4], tot = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) { const int k = 4 * lane + u; v[u] = k < nseg ? S.seg_off[k] : 0; tot += v[u]; }
        int x = tot;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        const int Ctot = __shfl_sync(0xffffffffu, x, 31);
        int run = x - tot;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; u++) { const int k = 4 * lane + u; if (k < nseg) S.seg_off[k] = run; run += v[u]; }
        if (lane == 0) S.seg_off[nseg] = Ctot;
    }
    __syncthreads();
    const int C = S.seg_off[nseg];
    if (C > SP_TILE || npart > SP_TR) {        // not a cell for this path: no trials, REPORTED through SweepAcc::pad (the host raises an error)
        if (tid == 0) { SweepAcc a = {0, 0, 0, 0, 0, npart, 0.0}; acc_out[c0] = a; ax.cells[c0].trial_base = 0; ax.cells[c0].nt = -1; atomicOr(fl.overflow, 64); }
        return;                                // (the sticky flag stops this and all later passes; the next synchronous call reports it)
    }
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    // With >= 5 cells on every axis two particles of one neighbourhood are less than half a box apart once both are taken relative
    // to the cell centre: the FP32 separation needs no minimum-image fold, and the staged coordinates are kept in LENGTH units
    // (a subtraction and a multiply-add per component and test instead of a fold). Otherwise box fractions and the fold.
    const bool nowrap = s.nc[0] >= 5 && s.nc[1] >= 5 && s.nc[2] >= 5;
    const float scx = nowrap ? boxf[0] : 1.f, scy = nowrap ? boxf[1] : 1.f, scz = nowrap ? boxf[2] : 1.f;
    auto staged_xyz = [&](double x, double y, double z, int wbits) {
        return make_float4((float)rel_frac(x + s.shift[0], ccen[0]) * scx, (float)rel_frac(y + s.shift[1], ccen[1]) * scy,
                           (float)rel_frac(z + s.shift[2], ccen[2]) * scz, __int_as_float(wbits));
    };
    for (int k = wid; k < nseg; k += NW) {
        const int b = S.seg_b[k], off = S.seg_off[k], len = S.seg_off[k + 1] - off;
        for (int idx = lane; idx < len; idx += 32) {
            const double4 pw = posw[b + idx];
            const double4 d = ldg256(rec + (size_t)(b + idx) * REC + R_DIR);
            S.t_df[off + idx] = make_float4((float)d.x, (float)d.y, (float)d.z, __int_as_float(b + idx));
            S.t_pf[off + idx] = staged_xyz(pw.x, pw.y, pw.z, w_orig(pw.w) | (w_type(pw.w) << 24));
        }
    }
    const int centre_seg = 2 * (g.k[2] * wy + g.k[1]);
    const int centre_off = S.seg_off[centre_seg] + (tb - S.seg_b[centre_seg]);
    // (sub: which of the n_sub walks through the system this pass belongs to -- its own permutation and proposals)
    const uint32_t ctr1 = (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 24) ^ ((uint32_t)sub << 8);
    const int nt = npart;
    // ---- a fresh random order of the cell's particles
    if (tid < nt) {
        const uint4 r = philox4x32((uint32_t)sweep, ctr1, (uint32_t)c0, (uint32_t)tid, (uint32_t)seed ^ 0xA511E9B3u, (uint32_t)(seed >> 32));
        S.key[tid] = r.x;
    }
    if (tid == 0) {
        const int base = atomicAdd(ax.trial_total, nt);
        S.ok = base + nt <= ax.trial_cap ? 1 : 0;
        if (!S.ok) atomicOr(fl.overflow, 16);
        S.tbase = base;
    }
    __syncthreads();
    if (!S.ok) { if (tid == 0) { ax.cells[c0].trial_base = 0; ax.cells[c0].nt = 0; } return; }
    if (tid < nt) {
        const unsigned int ka = S.key[tid];
        int rank = 0;
        for (int b = 0; b < nt; b++) { const unsigned int kb = S.key[b]; rank += (kb < ka || (kb == ka && b < tid)) ? 1 : 0; }
        S.perm[rank] = (unsigned short)tid;
    }
    __syncthreads();
    const int tbase = S.tbase;
    // ---- proposals, a warp per trial: the record goes through shared memory (one lane per vector for a rotation) into its spare slot
    for (int i = wid; i < nt; i += NW) {
        const int slot = tb + S.perm[i], nslot = s.n + tbase + i;
        double u[6];
#pragma unroll
        for (int k = 0; k < 3; k++) {      // Philox counter = (sweep, colour, cell, 3 i + k); every lane computes the same numbers
            const uint4 r = philox4x32((uint32_t)sweep, ctr1, (uint32_t)c0, (uint32_t)(3 * i + k), (uint32_t)seed, (uint32_t)(seed >> 32));
            u[2 * k] = u01(r.x, r.y);
            u[2 * k + 1] = u01(r.z, r.w);
        }
        const double4 pw = posw[slot];
        const int ty = w_type(pw.w), gt = sp.geotype_of_type[ty];
        const bool displace = (gt >= SCGPU_SPN) || (u[1] < 0.5);                 // particleMove (movecreator.cpp:11-33)
        const double z = 1.0 - 2.0 * u[2], phi = 6.283185307179586476925 * u[3];
        const double rr = sqrt(fmax(0.0, 1.0 - z * z));
        const v3 ax3 = mk(rr * cos(phi), rr * sin(phi), z);                     // uniform on the unit sphere
        double* rb = S.recbuf[wid];
        const double vold = rec[(size_t)slot * REC + lane];
        rb[lane] = vold;
        __syncwarp();
        if (lane == 0) {
            S.fo[i] = staged_xyz(rb[R_POS], rb[R_POS + 1], rb[R_POS + 2], ty);
            S.fdo[i] = make_float4((float)rb[R_DIR], (float)rb[R_DIR + 1], (float)rb[R_DIR + 2], 0.f);
        }
        __syncwarp();
        if (displace) {            // partDisplace (movecreator.cpp:947-994): fixed length trans_mx, uniform direction
            const double mx = sp.trans_mx[ty];
            if (lane == 0) { rb[R_POS] += ax3.x * mx / s.box[0]; rb[R_POS + 1] += ax3.y * mx / s.box[1]; rb[R_POS + 2] += ax3.z * mx / s.box[2]; }
        } else {                   // partRotate (movecreator.cpp:996-1028)
            double m[9];
            rotation_coefficients(m, sp.rot_angle[ty] * u[4], ax3, u[5] < 0.5);
            if (lane < 9 && record_vector_rotates(gt, lane)) rotate_vector(rb + 3 * lane, m);
        }
        __syncwarp();
        rec[(size_t)nslot * REC + lane] = rb[lane];
        if (lane == 0) {
            posw[nslot] = make_double4(rb[R_POS], rb[R_POS + 1], rb[R_POS + 2], pw.w);
            S.fn[i] = staged_xyz(rb[R_POS], rb[R_POS + 1], rb[R_POS + 2], ty);
            S.fdn[i] = make_float4((float)rb[R_DIR], (float)rb[R_DIR + 1], (float)rb[R_DIR + 2], 0.f);
            S.slot[i] = slot;
            // a move that leaves the cell would break the independence of the active cells: rejected
            const int inside = cell_index(rb + R_POS, s.shift, s.nc) == c0 ? 1 : 0;
            S.valid[i] = inside;
            SwTrial t;
            t.span = 0; t.cnt = 0; t.ncond = 0; t.slot_old = slot; t.slot_new = nslot; t.flags = inside | (displace ? 2 : 0); t.type = ty; t.pad = 0;
            t.u_acc = u[1] < 0.5 ? 2.0 * u[1] : 2.0 * u[1] - 1.0;   // the move-type bit is used up; the rest is still uniform
            ax.trials[tbase + i] = t;
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- gate, a warp per trial
    const float* hl_tab = tabs ? S.tab + 2 * T * T : s.reach2 + 2 * T * T + T;
    const float reach_one = (float)(ia1.reserved[1] * 1.001), cut_one = (float)(fmax(ia1.rcutSq, ia1.rcutwcaSq) * 1.001);
    for (int i = wid; i < nt; i += NW) {
        if (!S.valid[i]) continue;         // (span, cnt, ncond stay 0)
        int cnt = 0;
        bool over = false;
        const float4 fo = S.fo[i], fn = S.fn[i], fdo = S.fdo[i], fdn = S.fdn[i];
        const int type1 = __float_as_int(fo.w);
        const bool moved = fo.x != fn.x || fo.y != fn.y || fo.z != fn.z;
        const int slot_old = S.slot[i], slot_new = s.n + tbase + i;
        const float* reach_row = tabs ? S.tab + type1 * T : s.reach2 + type1 * T;
        const float* cut_row = tabs ? S.tab + T * T + type1 * T : s.reach2 + T * T + T + type1 * T;
        const float h1 = (RODS && ONE) ? (float)ia1.half_len[0] : hl_tab[type1];
        int2* ep = S.epair[wid];
        unsigned short* em = S.emeta[wid];
        auto emit = [&](bool on, int bslot, int meta) {
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (m == 0u) return;
            const int at = cnt + __popc(m & lt_mask);
            if (cnt + __popc(m) > SP_EB) { over = true; return; }
            if (on) { ep[at] = make_int2((meta & 1) ? slot_new : slot_old, bslot); em[at] = (unsigned short)meta; }
            cnt += __popc(m);
        };
        auto centre_d2 = [&](const float4& a, const float4& b, float& dx, float& dy, float& dz) {
            dx = a.x - b.x; dy = a.y - b.y; dz = a.z - b.z;
            if (!nowrap) { dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2]; }
            return dx * dx + dy * dy + dz * dz;
        };
        auto reach_of = [&](int ctype) { return (RODS && ONE) ? reach_one : reach_row[ctype]; };
        auto rods_apart = [&](int ctype, float dx, float dy, float dz, float d2, const float4& da, float bdx, float bdy, float bdz) {
            const float cut2 = (RODS && ONE) ? cut_one : cut_row[ctype];
            const float h2 = (RODS && ONE) ? h1 : hl_tab[ctype];
            return cut2 > 0.f && lb_beyond_fast(dx, dy, dz, d2, da.x, da.y, da.z, bdx, bdy, bdz, h1, h2, cut2);
        };
        // (1) partners that move EARLIER in this pass (same cell): both of their states against both of ours, conditional on their outcome
        for (int j0 = 0; j0 < i && !over; j0 += 8) {
            const int jj = j0 + (lane >> 2), x = lane & 1, y = (lane >> 1) & 1;
            bool pass = false;
            int bslot = 0;
            if (jj < i && S.valid[jj]) {
                const float4 q2 = y ? S.fn[jj] : S.fo[jj], d2v = y ? S.fdn[jj] : S.fdo[jj];
                const int ctype = __float_as_int(q2.w);
                float dx, dy, dz;
                const float d2 = centre_d2(x ? fn : fo, q2, dx, dy, dz);
                pass = d2 <= reach_of(ctype) && !rods_apart(ctype, dx, dy, dz, d2, x ? fdn : fdo, d2v.x, d2v.y, d2v.z);
                bslot = y ? s.n + tbase + jj : S.slot[jj];
            }
            emit(pass, bslot, x | (y << 1) | 4 | (jj << 3));
        }
        const int ncond = cnt;
        // (2) everybody else: centre-distance test of the whole neighbourhood with all lanes, then the segment bound over the hits
        int* hb = S.hit[wid];
        for (int base = 0; base < C && !over;) {
            int nh = 0;
            for (; base < C && nh <= SP_HB - 64; base += 32) {
                const int p = base + lane;
                bool ho = false, hn = false;
                if (p < C) {
                    const float4 q = S.t_pf[p];
                    const int wbits = __float_as_int(q.w);
                    const float reach = reach_of(wbits >> 24);
                    float dx, dy, dz;
                    ho = centre_d2(fo, q, dx, dy, dz) <= reach;
                    hn = moved ? centre_d2(fn, q, dx, dy, dz) <= reach : ho;      // a rotation leaves the centre where it was
                }
                const unsigned mo = __ballot_sync(0xffffffffu, ho), mn = __ballot_sync(0xffffffffu, hn);
                if (ho) hb[nh + __popc(mo & lt_mask)] = p * 2;
                nh += __popc(mo);
                if (hn) hb[nh + __popc(mn & lt_mask)] = p * 2 + 1;
                nh += __popc(mn);
            }
            __syncwarp();
            for (int k0 = 0; k0 < nh && !over; k0 += 32) {
                const int k = k0 + lane;
                bool pass = false;
                int bslot = 0, x = 0;
                if (k < nh) {
                    const int h = hb[k];
                    const int p = h >> 1;
                    x = h & 1;
                    const float4 q = S.t_pf[p], qd = S.t_df[p];
                    bslot = __float_as_int(qd.w);
                    const int ctype = __float_as_int(q.w) >> 24;
                    float dx, dy, dz;
                    const float d2 = centre_d2(x ? fn : fo, q, dx, dy, dz);
                    pass = bslot != slot_old && !rods_apart(ctype, dx, dy, dz, d2, x ? fdn : fdo, qd.x, qd.y, qd.z);
                    if (pass && p >= centre_off && p < centre_off + npart) {      // a particle of this cell that moves earlier: listed in (1)
                        for (int j = 0; j < i; j++) if (S.slot[j] == bslot && S.valid[j]) pass = false;
                    }
                }
                emit(pass, bslot, x);
            }
            __syncwarp();
        }
        // ---- the trial's span of the flat list
        int span = 0;
        if (lane == 0) {
            if (over) { cnt = 0; atomicOr(fl.overflow, 32); }
            span = atomicAdd(fl.total, cnt);
            if (span + cnt > fl.cap) { atomicOr(fl.overflow, 2); cnt = 0; }
            SwTrial* t = ax.trials + tbase + i;
            t->span = span; t->cnt = cnt; t->ncond = ncond < cnt ? ncond : cnt;
        }
        span = __shfl_sync(0xffffffffu, span, 0);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        for (int k = lane; k < cnt; k += 32) { fl.pair[span + k] = ep[k]; ax.meta[span + k] = em[k]; }
        __syncwarp();
    }
    if (tid == 0) { ax.cells[c0].trial_base = tbase; ax.cells[c0].nt = nt; }
}

struct SpResolveShared {
    double lo[SP_TR], ln[SP_TR], u_acc[SP_TR];
    double cval[SP_CPOOL];
    unsigned short cmeta[SP_CPOOL];
    int coff[SP_TR + 1], span[SP_TR], ncond[SP_TR], flags[SP_TR], slot_old[SP_TR], slot_new[SP_TR], type[SP_TR];
    int acc[SP_TR];
};

__global__ void __launch_bounds__(SP_RTHREADS)
k_sweep_resolve(DevSys s, SweepParams sp, int colour, int sub, SweepGrid g, double4* posw, double* rec, FlatList fl, SweepAux ax, SweepAcc* acc_out) {
    __shared__ SpResolveShared S;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SP_RTHREADS / 32;
    const int ax_ = s.nc[0] / g.ncol[0], ay = s.nc[1] / g.ncol[1];
    const int bx = blockIdx.x % ax_, by = (blockIdx.x / ax_) % ay, bz = blockIdx.x / (ax_ * ay);
    const int cx = bx * g.ncol[0] + (colour % g.ncol[0]), cy = by * g.ncol[1] + ((colour / g.ncol[0]) % g.ncol[1]), cz = bz * g.ncol[2] + (colour / (g.ncol[0] * g.ncol[1]));
    const int c0 = (cz * s.nc[1] + cy) * s.nc[0] + cx;
    if (blockIdx.x == 0 && tid == 0) { *fl.total = 0; *fl.ptotal = 0; *ax.trial_total = 0; }      // lists consumed by this launch's own reads below (by index, not by count)
    const SwCell cell = ax.cells[c0];
    const int nt = cell.nt;
    if (nt < 0) return;                // k_sweep_propose reported the cell (SweepAcc::pad)
    SweepAcc acc = {0, 0, 0, 0, 0, 0, 0.0};
    if (sub > 0 && tid == 0) acc = acc_out[c0];        // the statistics of a cell add up over the walks of one call
    if (nt == 0 || *fl.overflow) { if (tid == 0) acc_out[c0] = acc; return; }
    const SwTrial* tr = ax.trials + cell.trial_base;
    if (tid < nt) {
        const SwTrial t = tr[tid];
        S.span[tid] = t.span; S.ncond[tid] = t.ncond; S.flags[tid] = t.flags; S.slot_old[tid] = t.slot_old; S.slot_new[tid] = t.slot_new;
        S.type[tid] = t.type; S.u_acc[tid] = t.u_acc; S.acc[tid] = 0;
        S.lo[tid] = (double)t.cnt;     // (count parked here until the sums below replace it)
    }
    __syncthreads();
    if (wid == 0) {                    // where every trial's conditional terms sit in the pool
        int run = 0;
        for (int i0 = 0; i0 < nt; i0 += 32) {
            const int i = i0 + lane;
            const int v = i < nt ? S.ncond[i] : 0;
            int x = v;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (i < nt) S.coff[i] = run + x - v;
            run += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) S.coff[nt] = run;
    }
    __syncthreads();
    const bool pooled = S.coff[nt] <= SP_CPOOL;
    // ---- per trial, in parallel: the sums of its unconditional terms (old and new state), its conditional terms into the pool
    for (int i = wid; i < nt; i += NW) {
        const int span = S.span[i], cnt = (int)S.lo[i], nc = S.ncond[i];
        __syncwarp();
        double lo = 0.0, ln = 0.0;
        for (int k = lane; k < cnt; k += 32) {
            const double2 e = fl.e[span + k];
            const unsigned short m = ax.meta[span + k];
            const double v = e.x + e.y;
            if (k < nc) { if (pooled) { S.cval[S.coff[i] + k] = v; S.cmeta[S.coff[i] + k] = m; } }
            else if (m & 1) ln += v; else lo += v;
        }
        lo = warp_sum(lo); ln = warp_sum(ln);
        if (s.wall != nullptr) {       // [EXTER] wall: in both energies of the trial (totalenergycalculator.h:377-378, 410-411)
            double wv = 0.0;
            if (lane < 2) wv = wall_energy_rec(s, rec + (size_t)(lane ? S.slot_new[i] : S.slot_old[i]) * REC, S.type[i]);
            lo += __shfl_sync(0xffffffffu, wv, 0);
            ln += __shfl_sync(0xffffffffu, wv, 1);
        }
        if (lane == 0) { S.lo[i] = lo; S.ln[i] = ln; }
    }
    __syncthreads();
    // ---- the walk, one warp, trial by trial: plus the conditional terms of the state every earlier partner ended up in, then moveTry
    if (wid == 0) {
        for (int i = 0; i < nt; i++) {
            bool accept = false;
            double de = 0.0;
            const int fl_ = S.flags[i];
            if (fl_ & 1) {
                double lo = 0.0, ln = 0.0;
                const int nc = S.ncond[i];
                if (nc > 0) {
                    for (int k = lane; k < nc; k += 32) {
                        double v; unsigned short m;
                        if (pooled) { v = S.cval[S.coff[i] + k]; m = S.cmeta[S.coff[i] + k]; }
                        else { const double2 e = fl.e[S.span[i] + k]; v = e.x + e.y; m = ax.meta[S.span[i] + k]; }
                        if (((m >> 1) & 1) == S.acc[m >> 3]) { if (m & 1) ln += v; else lo += v; }
                    }
                    lo = warp_sum(lo); ln = warp_sum(ln);
                }
                const double e_old = S.lo[i] + lo, e_new = S.ln[i] + ln;
                de = e_new - e_old;
                accept = (de <= 0.0) || (exp(-de / sp.temper) > S.u_acc[i]);                    // moveTry (movecreator.h:175-187)
            }
            if (lane == 0) {
                S.acc[i] = accept ? 1 : 0;
                if (accept) acc.de += de;
                if (!(fl_ & 1)) acc.cell_rej++;
                if (fl_ & 2) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
                else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
            }
            __syncwarp();
        }
        if (lane == 0) acc_out[c0] = acc;
    }
    __syncthreads();
    // ---- commit: the accepted trial states move from their spare slots over the old ones
    for (int i0 = wid; i0 < nt; i0 += 4 * NW) {          // four records in flight per warp
        double v[4];
        double4 pw[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * NW;
            on[u] = i < nt && S.acc[i];
            if (on[u]) { v[u] = rec[(size_t)S.slot_new[i] * REC + lane]; if (lane == 0) pw[u] = posw[S.slot_new[i]]; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * NW;
            if (on[u]) { rec[(size_t)S.slot_old[i] * REC + lane] = v[u]; if (lane == 0) posw[S.slot_old[i]] = pw[u]; }
        }
    }
}
