// Checkerboard sweeps, second kernel: ROUNDS of trials evaluated in parallel, resolved in sequence.
//
// k_sweep_cells (sweep.cuh) walks the trials of an active cell one after another; a trial is a chain of dependent FP64
// operations (segment distance, two patch intersections, acos / cos) of ~25 000 cycles that no amount of lanes shortens, so a
// sweep costs (trials per cell and pass) x (passes) of those latencies however idle the GPU is. Here a BLOCK owns an active cell
// and takes its trials in rounds: the next trials of the cell's (unchanged, Philox-keyed) trial sequence up to the first one that
// picks a particle already in the round. All pair terms the round can possibly need are evaluated at once by all threads:
//   * every trial particle in its old and in its new state against every partner outside the round,
//   * for two trial particles t' < t of the same round: t (old, new) against t' (old, new) -- four terms, of which the
//     resolution uses the two that match the outcome of t'.
// Then ONE warp resolves the round in trial order (sum of the matching terms, moveTry), exactly the decisions the sequential
// walk would take: the Markov chain is the same sequential chain of single-particle Metropolis moves inside the cell, only
// the arithmetic is hoisted out of it. (Sums are taken in a different order than k_sweep_cells takes them, so the two kernels
// agree to rounding, not bit for bit; each is bit-reproducible on its own.)
// Used for systems without bonded molecules; chains and lipids stay on k_sweep_cells / k_sweep_chain_colour.
#pragma once

constexpr int SR_THREADS = 128;
constexpr int SR_TILE = 768;           // staged neighbourhood (FP32 position + direction + slot, 32 B per candidate)
constexpr int SR_TR = 16;              // trials per round at most
constexpr int SR_Q = 64;               // work-list entries per trial of a round; a trial that needs more runs alone with the whole list
constexpr int SR_ENT = SR_TR * SR_Q;
constexpr int SR_HB = SR_ENT * 2 / (SR_THREADS / 32);      // (the buffers live in SrShared::e, idle during the gate)
//           // centre-distance hits a warp collects before it runs the segment bound over them
constexpr int SR_PERM = 1024;          // largest cell population that is walked as a permutation (trial_rule 2); larger cells draw with replacement

struct SrEntry {
    int p;                    // partner: index into the staged neighbourhood
    unsigned short t;         // trial of the round
    unsigned char x;          // state of the trial's own particle: 0 old, 1 new
    unsigned char flags;      // bit 0: partner in the NEW state of its own (earlier) trial; bit 1: conditional on that trial's outcome;
};                            // bits 4..7: that trial's index in the round

struct SrShared {
    float4 t_pf[SR_TILE];
    float4 t_df[SR_TILE];
    double rec_old[SR_TR][REC], rec_new[SR_TR][REC];
    double e[SR_ENT];
    SrEntry ent[SR_ENT];
    unsigned short plist[SR_ENT];
    int seg_b[2 * SW_MAXROWS + 2], seg_off[2 * SW_MAXROWS + 2];
    float tab[2 * SW_MAXT * SW_MAXT + SW_MAXT];
    double u_acc[SR_TR];
    int pick[SR_TR], valid[SR_TR], disp[SR_TR], type[SR_TR], molt[SR_TR], orig[SR_TR], cnt[SR_TR], base[SR_TR + 1], acc[SR_TR];
    int nt, npl, solo;
    unsigned short perm[SR_PERM];
};
static_assert(SR_PERM * sizeof(unsigned int) <= SR_ENT * sizeof(double), "the permutation keys borrow SrShared::e");

template <bool RODS, bool ONE>
__global__ void __launch_bounds__(SR_THREADS, 4)
k_sweep_rounds(DevSys s, SweepParams sp, unsigned long long seed, unsigned long long sweep, int colour, SweepGrid g,
               double4* posw, double* rec, SweepAcc* acc_out, const __grid_constant__ scgpu_iaparam ia1) {
    extern __shared__ __align__(16) unsigned char sr_raw[];
    SrShared& S = *reinterpret_cast<SrShared*>(sr_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = SR_THREADS / 32;
    const unsigned lt_mask = (1u << lane) - 1u;
    // active cell of this block
    const int ax = s.nc[0] / g.ncol[0], ay = s.nc[1] / g.ncol[1];
    const int bx = blockIdx.x % ax, by = (blockIdx.x / ax) % ay, bz = blockIdx.x / (ax * ay);
    const int cx = bx * g.ncol[0] + (colour % g.ncol[0]), cy = by * g.ncol[1] + ((colour / g.ncol[0]) % g.ncol[1]), cz = bz * g.ncol[2] + (colour / (g.ncol[0] * g.ncol[1]));
    const int c0 = (cz * s.nc[1] + cy) * s.nc[0] + cx;
    const int tb = s.cell_start[c0], te = s.cell_start[c0 + 1];
    const int npart = te - tb;
    SweepAcc acc = {0, 0, 0, 0, 0, 0, 0.0};
    if (npart == 0) { if (tid == 0) acc_out[c0] = acc; return; }
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
        S.seg_b[tid] = b; S.seg_off[tid] = len;           // lengths first, offsets below
    }
    const bool tabs = !(RODS && ONE) && T <= SW_MAXT;       // reach / cutoff / half-length tables in shared memory
    if (tabs) for (int k = tid; k < 2 * T * T + T; k += SR_THREADS) S.tab[k] = k < T * T ? s.reach2[k] : s.reach2[k + T];
    __syncthreads();
    if (wid == 0) {       // exclusive scan of the segment lengths (at most 98: four per lane)
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
    const bool tiled = C <= SR_TILE;        // denser neighbourhoods are scanned from global memory (slower, same results)
    const double ccen[3] = {(cx + 0.5) / s.nc[0], (cy + 0.5) / s.nc[1], (cz + 0.5) / s.nc[2]};
    const float boxf[3] = {(float)s.box[0], (float)s.box[1], (float)s.box[2]};
    auto slot_of_p = [&](int p) {
        int lo = 0, hi = nseg - 1;             // the last segment whose offset is <= p (offsets are non-decreasing; empty segments repeat them)
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.seg_off[mid] <= p) lo = mid; else hi = mid - 1; }
        return S.seg_b[lo] + (p - S.seg_off[lo]);
    };
    auto staged_xyz = [&](double x, double y, double z, int wbits) {
        return make_float4((float)rel_frac(x + s.shift[0], ccen[0]), (float)rel_frac(y + s.shift[1], ccen[1]),
                           (float)rel_frac(z + s.shift[2], ccen[2]), __int_as_float(wbits));
    };
    auto staged_pos = [&](const double4& pw) { return staged_xyz(pw.x, pw.y, pw.z, w_orig(pw.w) | (w_type(pw.w) << 24)); };
    auto staged_dir = [&](int slot) {
        const double4 d = ldg256(rec + (size_t)slot * REC + R_DIR);
        return make_float4((float)d.x, (float)d.y, (float)d.z, __int_as_float(slot));
    };
    if (tiled) {
        for (int k = wid; k < nseg; k += NW) {
            const int b = S.seg_b[k], off = S.seg_off[k], len = S.seg_off[k + 1] - off;
            for (int idx = lane; idx < len; idx += 32) {
                const double4 pw = posw[b + idx];
                S.t_df[off + idx] = staged_dir(b + idx);
                S.t_pf[off + idx] = staged_pos(pw);
            }
        }
    }
    // the active cell is the centre of its own neighbourhood: where its particles sit in the staged order
    const int centre_seg = 2 * (g.k[2] * wy + g.k[1]);
    const int centre_off = S.seg_off[centre_seg] + (tb - S.seg_b[centre_seg]);
    // trials of this cell in this pass: as in k_sweep_cells (SweepParams::trial_rule)
    int ntrial;
    {
        const double avg = sp.trial_rule >= 1 ? sp.trial_scale * (double)sp.n_sub * (double)npart
                                              : sp.trial_scale * (double)sp.n_sub * (double)s.n / (double)s.cell_start[s.ncells + 1];
        const uint4 r = philox4x32((uint32_t)sweep, (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 24), (uint32_t)c0, 0xffffffffu, (uint32_t)seed, (uint32_t)(seed >> 32));
        const double fl = floor(avg);
        ntrial = (int)fl + (u01(r.x, r.y) < avg - fl ? 1 : 0);
    }
    // trial_rule 2: the particles of the cell are walked in a fresh random order n_sub times (every particle exactly once per sweep);
    // a round then never meets a particle twice and is as long as the list allows
    const bool permute = sp.trial_rule == 2 && npart <= SR_PERM;
    const uint32_t ctr1 = (uint32_t)(sweep >> 32) ^ ((uint32_t)colour << 24);
    const float* hl_tab = tabs ? S.tab + 2 * T * T : s.reach2 + 2 * T * T + T;
    const float reach_one = (float)(ia1.reserved[1] * 1.001), cut_one = (float)(fmax(ia1.rcutSq, ia1.rcutwcaSq) * 1.001);
    if (tid == 0) S.solo = 0;
    __syncthreads();

    for (int tr0 = 0; tr0 < ntrial;) {
        // ---- the round: the next trials up to the first that picks a particle already taken (Philox counter = (sweep, colour, cell, 3 trial + k))
        if (permute && tr0 % npart == 0) {       // a new walk through the cell: rank the particles by fresh random keys
            const int sub = tr0 / npart;
            unsigned int* key = reinterpret_cast<unsigned int*>(S.e);          // idle until the round's terms are evaluated
            for (int a = tid; a < npart; a += SR_THREADS) {
                const uint4 r = philox4x32((uint32_t)sweep, ctr1, (uint32_t)c0, (uint32_t)(sub * SR_PERM + a), (uint32_t)seed ^ 0xA511E9B3u, (uint32_t)(seed >> 32));
                key[a] = r.x;
            }
            __syncthreads();
            for (int a = tid; a < npart; a += SR_THREADS) {
                const unsigned int ka = key[a];
                int rank = 0;
                for (int b = 0; b < npart; b++) { const unsigned int kb = key[b]; rank += (kb < ka || (kb == ka && b < a)) ? 1 : 0; }
                S.perm[rank] = (unsigned short)a;
            }
            __syncthreads();
        }
        if (tid < SR_TR) {
            const int tr = tr0 + tid;
            int pick = -1;
            if (tr < ntrial) {
                if (permute) { if (tr0 % npart + tid < npart) pick = tb + S.perm[tr0 % npart + tid]; }      // a round does not run into the next walk
                else {
                    const uint4 r = philox4x32((uint32_t)sweep, ctr1, (uint32_t)c0, (uint32_t)(3 * tr), (uint32_t)seed, (uint32_t)(seed >> 32));
                    pick = tb + (int)(u01(r.x, r.y) * npart);          // uniformly chosen particle of this cell, with replacement
                    if (pick >= te) pick = te - 1;
                }
            }
            S.pick[tid] = pick;
        }
        __syncthreads();
        if (wid == 0) {
            bool stop = lane >= SR_TR || S.pick[lane < SR_TR ? lane : 0] < 0;
            if (!stop && !permute) for (int j = 0; j < lane; j++) if (S.pick[j] == S.pick[lane]) stop = true;
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            int n = __ffs(m) - 1;
            if (S.solo) n = 1;
            if (lane == 0) { S.nt = n; S.npl = 0; }
        }
        __syncthreads();
        int nt = S.nt;
        const bool solo = S.solo != 0;
        // ---- records of the round's particles (a warp per record: REC == 32 == the warp)
        for (int i = wid; i < nt; i += NW) { const double v = rec[(size_t)S.pick[i] * REC + lane]; S.rec_old[i][lane] = v; S.rec_new[i][lane] = v; }
        __syncthreads();
        // ---- proposals, one thread per trial: nothing here depends on the outcome of other trials
        if (tid < nt) {
            const int tr = tr0 + tid, slot = S.pick[tid];
            double u[6];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint4 r = philox4x32((uint32_t)sweep, ctr1, (uint32_t)c0, (uint32_t)(3 * tr + k), (uint32_t)seed, (uint32_t)(seed >> 32));
                u[2 * k] = u01(r.x, r.y);
                u[2 * k + 1] = u01(r.z, r.w);
            }
            const double w = posw[slot].w;
            const int ty = w_type(w), gt = sp.geotype_of_type[ty];
            const bool displace = (gt >= SCGPU_SPN) || (u[1] < 0.5);                 // particleMove (movecreator.cpp:11-33)
            const double z = 1.0 - 2.0 * u[2], phi = 6.283185307179586476925 * u[3];
            const double rr = sqrt(fmax(0.0, 1.0 - z * z));
            const v3 ax3 = mk(rr * cos(phi), rr * sin(phi), z);                     // uniform on the unit sphere
            double* rn = S.rec_new[tid];
            if (displace) {            // partDisplace (movecreator.cpp:947-994): fixed length trans_mx, uniform direction
                const double mx = sp.trans_mx[ty];
                rn[R_POS] += ax3.x * mx / s.box[0]; rn[R_POS + 1] += ax3.y * mx / s.box[1]; rn[R_POS + 2] += ax3.z * mx / s.box[2];
            } else {                   // partRotate (movecreator.cpp:996-1028)
                double m[9];
                rotation_coefficients(m, sp.rot_angle[ty] * u[4], ax3, u[5] < 0.5);
                for (int v = 0; v < 9; v++) if (record_vector_rotates(gt, v)) rotate_vector(rn + 3 * v, m);
            }
            S.disp[tid] = displace ? 1 : 0;
            S.u_acc[tid] = u[1] < 0.5 ? 2.0 * u[1] : 2.0 * u[1] - 1.0;   // the move-type bit is used up; the rest is still uniform
            S.type[tid] = ty;
            S.orig[tid] = w_orig(w);
            S.molt[tid] = w_moltype(w);
            // a move that leaves the cell would break the independence of the active cells: rejected
            S.valid[tid] = cell_index(rn + R_POS, s.shift, s.nc) == c0 ? 1 : 0;
            S.acc[tid] = 0;
        }
        __syncthreads();
        // ---- gate, a warp per trial: the staged neighbourhood against the old AND the new state (centre distance against the exact
        // reach of the type pair, then the segment lower bound); a partner that has an EARLIER trial in this round is listed in both
        // of its states, conditionally
        for (int i = wid; i < nt; i += NW) {
            int cnt = 0;
            bool over = false;
            if (S.valid[i]) {
                const int ebase = solo ? 0 : i * SR_Q, ecap = solo ? SR_ENT : SR_Q;
                const int type1 = S.type[i], target = S.orig[i];
                const double* ro = S.rec_old[i];
                const double* rn = S.rec_new[i];
                const float4 fo = staged_xyz(ro[R_POS], ro[R_POS + 1], ro[R_POS + 2], 0), fn = staged_xyz(rn[R_POS], rn[R_POS + 1], rn[R_POS + 2], 0);
                const float dox = (float)ro[R_DIR], doy = (float)ro[R_DIR + 1], doz = (float)ro[R_DIR + 2];
                const float dnx = (float)rn[R_DIR], dny = (float)rn[R_DIR + 1], dnz = (float)rn[R_DIR + 2];
                const float* reach_row = tabs ? S.tab + type1 * T : s.reach2 + type1 * T;
                const float* cut_row = tabs ? S.tab + T * T + type1 * T : s.reach2 + T * T + T + type1 * T;
                const float h1 = (RODS && ONE) ? (float)ia1.half_len[0] : hl_tab[type1];
                auto emit = [&](bool on, int p, int x, int flags) {
                    const unsigned m = __ballot_sync(0xffffffffu, on);
                    if (m == 0u) return;
                    const int at = cnt + __popc(m & lt_mask);
                    if (cnt + __popc(m) > ecap) { over = true; return; }
                    if (on) { SrEntry en; en.p = p; en.t = (unsigned short)i; en.x = (unsigned char)x; en.flags = (unsigned char)flags; S.ent[ebase + at] = en; }
                    cnt += __popc(m);
                };
                // the pair test in FP32: centre distance against the exact reach of the type pair, then the segment lower bound
                auto centre_d2 = [&](const float4& a, const float4& b, float& dx, float& dy, float& dz) {
                    dx = a.x - b.x; dy = a.y - b.y; dz = a.z - b.z;
                    dx = (dx - rintf(dx)) * boxf[0]; dy = (dy - rintf(dy)) * boxf[1]; dz = (dz - rintf(dz)) * boxf[2];
                    return dx * dx + dy * dy + dz * dz;
                };
                auto reach_of = [&](int ctype) { return (RODS && ONE) ? reach_one : reach_row[ctype]; };
                auto rods_apart = [&](int ctype, float dx, float dy, float dz, float d2, float adx, float ady, float adz, float bdx, float bdy, float bdz) {
                    const float cut2 = (RODS && ONE) ? cut_one : cut_row[ctype];
                    const float h2 = (RODS && ONE) ? h1 : hl_tab[ctype];
                    return cut2 > 0.f && lb_beyond_fast(dx, dy, dz, d2, adx, ady, adz, bdx, bdy, bdz, h1, h2, cut2);
                };
                // (1) partners that move EARLIER in this round: both of their states against both of ours, conditional on their outcome;
                //     eight partners per trip, four lanes each
                for (int j0 = 0; j0 < i && !over; j0 += 8) {
                    const int jj = j0 + (lane >> 2), x = lane & 1, y = (lane >> 1) & 1;
                    bool pass = false;
                    int p = 0;
                    if (jj < i && S.valid[jj]) {
                        const double* r2 = y ? S.rec_new[jj] : S.rec_old[jj];
                        const float4 q2 = staged_xyz(r2[R_POS], r2[R_POS + 1], r2[R_POS + 2], 0);
                        const int ctype = S.type[jj];
                        float dx, dy, dz;
                        const float d2 = centre_d2(x ? fn : fo, q2, dx, dy, dz);
                        pass = d2 <= reach_of(ctype) &&
                               !rods_apart(ctype, dx, dy, dz, d2, x ? dnx : dox, x ? dny : doy, x ? dnz : doz, (float)r2[R_DIR], (float)r2[R_DIR + 1], (float)r2[R_DIR + 2]);
                        p = centre_off + (S.pick[jj] - tb);
                    }
                    emit(pass, p, x, 2 | y | (jj << 4));
                }
                // (2) everybody else, two steps: the centre-distance test of the whole neighbourhood with all lanes (hits are compacted into
                //     the warp's buffer), then the segment bound over the hits, again with all lanes
                int* hb = reinterpret_cast<int*>(S.e) + wid * SR_HB;
                for (int base = 0; base < C && !over;) {
                    int nh = 0;
                    for (; base < C && nh <= SR_HB - 64; base += 32) {
                        const int p = base + lane;
                        bool ho = false, hn = false;
                        if (p < C) {
                            float4 q;
                            if (tiled) q = S.t_pf[p]; else q = staged_pos(posw[slot_of_p(p)]);
                            const int wbits = __float_as_int(q.w);
                            if ((wbits & 0xffffff) != target) {
                                const float reach = reach_of(wbits >> 24);
                                float dx, dy, dz;
                                ho = centre_d2(fo, q, dx, dy, dz) <= reach;
                                hn = centre_d2(fn, q, dx, dy, dz) <= reach;
                            }
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
                        int p = 0, x = 0;
                        if (k < nh) {
                            const int h = hb[k];
                            p = h >> 1; x = h & 1;
                            float4 q, qd;
                            if (tiled) { q = S.t_pf[p]; qd = S.t_df[p]; }
                            else { const int sl = slot_of_p(p); q = staged_pos(posw[sl]); qd = staged_dir(sl); }
                            const int ctype = __float_as_int(q.w) >> 24;
                            float dx, dy, dz;
                            const float d2 = centre_d2(x ? fn : fo, q, dx, dy, dz);
                            pass = !rods_apart(ctype, dx, dy, dz, d2, x ? dnx : dox, x ? dny : doy, x ? dnz : doz, qd.x, qd.y, qd.z);
                            if (pass && p >= centre_off && p < centre_off + npart) {      // a particle of this cell that moves earlier in the round: listed in (1)
                                const int sl = tb + (p - centre_off);
                                for (int j = 0; j < i; j++) if (S.pick[j] == sl && S.valid[j]) pass = false;
                            }
                        }
                        emit(pass, p, x, 0);
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) S.cnt[i] = over ? -1 : cnt;
        }
        __syncthreads();
        // ---- a trial whose partners do not fit its share of the list ends the round before it; it then runs alone with the whole list
        if (wid == 0) {
            const int cv = lane < nt ? S.cnt[lane] : 0;
            const unsigned mo = __ballot_sync(0xffffffffu, cv < 0);
            int n = mo ? __ffs(mo) - 1 : nt;
            int x = (lane < n) ? cv : 0;
            const int own = x;
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane <= n && lane <= SR_TR) S.base[lane] = x - own;          // exclusive prefix; base[n] = total
            if (lane == 0) {
                if (n == 0) {
                    if (solo) { acc.pad++; S.valid[0] = 0; S.cnt[0] = 0; S.base[1] = 0; n = 1; }      // cannot be evaluated: rejected and REPORTED (SweepAcc::pad)
                    else S.solo = 1;
                } else S.solo = 0;
                S.nt = n;
            }
        }
        __syncthreads();
        nt = S.nt;
        if (nt == 0) continue;             // repeat this trial alone
        const int total = S.base[nt];
        // ---- everything but the rod-rod patch term, a thread per listed term; terms that owe a patch evaluation are collected
        for (int idx0 = 0; idx0 < total; idx0 += SR_THREADS) {
            const int idx = idx0 + tid;
            bool np = false;
            int eidx = 0;
            if (idx < total) {
                int i = 0;
                while (S.base[i + 1] <= idx) i++;
                eidx = (solo ? 0 : i * SR_Q) + (idx - S.base[i]);
                const SrEntry en = S.ent[eidx];
                const double* s1 = en.x ? S.rec_new[i] : S.rec_old[i];
                const int type1 = S.type[i];
                const double* s2;
                int type2, orig2;
                v3 p2;
                if (en.flags & 1) { const int j = en.flags >> 4; s2 = S.rec_new[j]; type2 = S.type[j]; orig2 = S.orig[j]; p2 = ld3(s2 + R_POS); }
                else {
                    const int slot = tiled ? __float_as_int(S.t_df[en.p].w) : slot_of_p(en.p);
                    const double4 pw = posw[slot];
                    s2 = rec + (size_t)slot * REC; type2 = w_type(pw.w); orig2 = w_orig(pw.w); p2 = mk(pw.x, pw.y, pw.z);
                }
                const v3 r = image(s.box, ld3(s1 + R_POS), p2);
                const double d = dot(r, r);
                double e = 0.0;
                if (d <= s.sqmaxcut) {
                    if (RODS && ONE) e = pair_energy_cheap_rods(ia1, r, d, ld3(s1 + R_DIR), ld3(s2 + R_DIR), np);
                    else {
                        ConList cl;
                        cl.is_empty = 1; cl.con[0] = cl.con[1] = cl.con[2] = cl.con[3] = -1; cl.sp = cl.mod0 = cl.mod1 = cl.c0 = cl.c1 = cl.eq0 = cl.eq1 = 0.0;
                        e = pair_energy_cheap<RODS>(s.box, s.ia, s.ntypes, s.mol, r, d, s1, type1, S.molt[i], s2, type2, orig2, cl, np);
                    }
                }
                S.e[eidx] = e;
            }
            const unsigned m = __ballot_sync(0xffffffffu, np);
            if (m) {
                int at = 0;
                if (lane == 0) at = atomicAdd(&S.npl, __popc(m));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (np) S.plist[at + __popc(m & lt_mask)] = (unsigned short)eidx;
            }
        }
        __syncthreads();
        // ---- the patch terms, a thread per term (each adds to its own slot: the order of the list does not matter)
        const int npl = S.npl;
        for (int k = tid; k < npl; k += SR_THREADS) {
            const int eidx = S.plist[k];
            const SrEntry en = S.ent[eidx];
            const int i = en.t;
            const double* s1 = en.x ? S.rec_new[i] : S.rec_old[i];
            const double* s2;
            int type2;
            v3 p2;
            if (en.flags & 1) { const int j = en.flags >> 4; s2 = S.rec_new[j]; type2 = S.type[j]; p2 = ld3(s2 + R_POS); }
            else {
                const int slot = tiled ? __float_as_int(S.t_df[en.p].w) : slot_of_p(en.p);
                const double4 pw = posw[slot];
                s2 = rec + (size_t)slot * REC; type2 = w_type(pw.w); p2 = mk(pw.x, pw.y, pw.z);
            }
            const v3 r = image(s.box, ld3(s1 + R_POS), p2);
            S.e[eidx] += pair_energy_patch(ONE ? ia1 : s.ia[S.type[i] * s.ntypes + type2], r, s1, s2);
        }
        __syncthreads();
        // ---- resolution, one warp, trial by trial: the terms against partners outside the round plus, for every partner that moved
        // earlier in the round, the terms of the state it ended up in; then moveTry (movecreator.h:175-187)
        if (wid == 0) {
            for (int i = 0; i < nt; i++) {
                bool accept = false;
                double de = 0.0;
                if (S.valid[i]) {
                    const int ebase = solo ? 0 : i * SR_Q, n = S.cnt[i];
                    double lo = 0.0, ln = 0.0;
                    for (int k = lane; k < n; k += 32) {
                        const SrEntry en = S.ent[ebase + k];
                        const bool use = !(en.flags & 2) || ((en.flags & 1) == S.acc[en.flags >> 4]);
                        if (use) { if (en.x) ln += S.e[ebase + k]; else lo += S.e[ebase + k]; }
                    }
                    double e_old = warp_sum(lo), e_new = warp_sum(ln);
                    if (s.wall != nullptr) {       // [EXTER] wall: in both energies of the trial (totalenergycalculator.h:377-378, 410-411)
                        double wv = 0.0;
                        if (lane < 2) wv = wall_energy_rec(s, lane ? S.rec_new[i] : S.rec_old[i], S.type[i]);
                        e_old += __shfl_sync(0xffffffffu, wv, 0);
                        e_new += __shfl_sync(0xffffffffu, wv, 1);
                    }
                    de = e_new - e_old;
                    accept = (de <= 0.0) || (exp(-de / sp.temper) > S.u_acc[i]);
                }
                if (lane == 0) {
                    S.acc[i] = accept ? 1 : 0;
                    if (accept) acc.de += de;
                    if (!S.valid[i]) acc.cell_rej++;
                    if (S.disp[i]) { if (accept) acc.trans_acc++; else acc.trans_rej++; }
                    else { if (accept) acc.rot_acc++; else acc.rot_rej++; }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- commit in place: sorted record, position word, staged FP32 copies
        for (int i = wid; i < nt; i += NW) {
            if (!S.acc[i]) continue;
            const int slot = S.pick[i];
            rec[(size_t)slot * REC + lane] = S.rec_new[i][lane];
            if (lane == 0) {
                const double* rn = S.rec_new[i];
                const double4 npw = make_double4(rn[R_POS], rn[R_POS + 1], rn[R_POS + 2], posw[slot].w);
                posw[slot] = npw;
                if (tiled) {
                    S.t_pf[centre_off + (slot - tb)] = staged_pos(npw);
                    S.t_df[centre_off + (slot - tb)] = make_float4((float)rn[R_DIR], (float)rn[R_DIR + 1], (float)rn[R_DIR + 2], __int_as_float(slot));
                }
            }
        }
        __syncthreads();
        tr0 += nt;
    }
    if (tid == 0) acc_out[c0] = acc;
}
