/* scgpu -- C ABI of the B200-native energy engine for patchy-spherocylinder Monte Carlo.
 *
 * This is the drop-in boundary for the ONE data-parallel hot path of robertvacha/SC (scOOP): pair energy
 * and overlap evaluation over cell lists, the full-system energy sum, batched checkerboard trial moves.
 * The reference has no plugin/FFI layer; its seam is the compile-time calculator class
 *     typedef TotalEMatrix<PairE> TotalEnergyCalculator;      (scOOP/mc/totalenergycalculator.h:1227)
 * whose virtual interface is TotalE<> (scOOP/mc/totalenergycalculator.h:135-297). A reference-side
 * TotalEGpu : TotalE<PairE> (shown in INTEGRATION.md; our own host mirror is sc_b200/csrc/host/) forwards each virtual
 * to one entry point below. Each entry point cites the reference member it replaces.
 *
 * Conventions: every function returns 0 on success, <0 on error (message: scgpu_last_error()). One context
 * = one CUDA device = one stream; a context is not thread-safe. The caller owns all host buffers, the
 * library owns all device memory. All reals are IEEE double; positions are in BOX-FRACTION units exactly
 * as in the reference's Particle::pos; everything else is in real units. There is no CPU fallback: every
 * call fails with SCGPU_ERR_CUDA if no CUDA device is usable.
 */
#ifndef SCGPU_H
#define SCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCGPU_OK 0
#define SCGPU_ERR_ARG (-1)
#define SCGPU_ERR_CUDA (-2)
#define SCGPU_ERR_STATE (-3)

#define SCGPU_STATE_DOUBLES 30
#define SCGPU_IAPARAM_DOUBLES 48
#define SCGPU_MOLPARAM_DOUBLES 16

/* geotype codes (scOOP/structures/macros.h:69-85) */
enum {
    SCGPU_SCN = 10, SCGPU_SCA = 11, SCGPU_PSC = 12, SCGPU_CPSC = 13, SCGPU_CHPSC = 14, SCGPU_CHCPSC = 15,
    SCGPU_TPSC = 16, SCGPU_TCPSC = 17, SCGPU_TCHPSC = 18, SCGPU_TCHCPSC = 19, SCGPU_SPN = 30, SCGPU_SPA = 31
};

/* Packed copy of the pair-path fields of Ia_param (scOOP/structures/structures.h:187-254); one entry per
 * ordered type pair, 48 doubles. geotype[] and exclude hold small integers. */
typedef struct scgpu_iaparam {
    double geotype[2];
    double exclude;
    double sigma, epsilon, A, B;
    double pdis, pswitch, pswitchINV;
    double rcut, rcutSq, rcutwca, rcutwcaSq;
    double parallel;
    double half_len[2];
    double pcangl[4], pcanglsw[4];
    double pcoshalfi[4], psinhalfi[4];
    double csecpatchrot[2], ssecpatchrot[2];
    double chiral_cos[2], chiral_sin[2];
    double len[2];
    double reserved[5];
} scgpu_iaparam;

/* Per molecule type: bond/angle constants of MoleculeParams (scOOP/structures/moleculeparams.h:10-45) plus
 * the molecule size and ParticleVector::first[] (scOOP/structures/Conf.h:38-42) that getConlist needs. */
typedef struct scgpu_molparam {
    double bond1eq, bond1c, bond2eq, bond2c, bonddeq, bonddc, bondheq, bondhc;
    double angle1eq, angle1c, angle2eq, angle2c;
    double mol_size, first;
    double reserved[2];
} scgpu_molparam;

/* A particle is 30 doubles, the vector members of Particle in declaration order
 * (scOOP/structures/particle.h:26-30): pos[3] dir[3] patchdir[2][3] patchsides[4][3] chdir[2][3]. */

typedef struct scgpu_ctx scgpu_ctx;

/* Batched trial moves (replaces the per-step body of Updater::simulate, scOOP/mc/updater.cpp:206-230, for
 * displacement/rotation: MoveCreator::partDisplace/partRotate, scOOP/mc/movecreator.cpp:947-1028). */
typedef struct scgpu_moveparams {
    double temper;                 /* Sim::temper */
    double trans_mx[40];           /* per particle type: stat.trans[type].mx  (= 2*transmx, sim.h:365) */
    double rot_angle[40];          /* per particle type: stat.rot[type].angle (radians, sim.h:360) */
    int n_sub;                     /* sweeps per call: n_sub * N trials in total, shared among the cells by trial_rule */
    int grid_k;                    /* fineness of the checkerboard: 0 = chosen by the library (the finest grid that leaves ~2 particles per
                                      cell: shortest serial chain, best for ONE system on the GPU); 1, 2, 3 = cells of edge >= maxcut / k.
                                      Many replicas sharing a GPU are throughput-bound and run best on the coarse grid (1) */
    int trial_rule;                /* how the n_sub * N trials are shared among the cells of the checkerboard.
                                      0: every non-empty cell performs the same number, n_sub * N / (non-empty cells) -- all cells of a pass
                                      finish together (shortest pass; the rule of HOOMD-blue's HPMC, a fixed number of trials per cell);
                                      particles of sparse cells are moved more often than those of dense ones;
                                      1: a cell performs n_sub * (its population) trials -- every particle is picked once per sweep on
                                      average, the reference's per-particle rates (Updater::simulate draws uniformly among N,
                                      updater.cpp:206-230), at the price of every pass lasting as long as its fullest cell.
                                      2: every particle is tried exactly once per sweep: the particles of a cell are walked in a fresh
                                      random order (a random-order sequential sweep; balance holds, the order is independent of the
                                      configuration). Kernels that draw with replacement (bonded systems, chain sweeps) treat 2 as 1.
                                      On a system without bonds this rule runs as four dense launches per colour pass
                                      (sweep_phased.cuh) -- the fastest form.
                                      All rules leave the Boltzmann distribution invariant (the count is fixed before the pass and no
                                      particle leaves its cell within one); fractional counts are stochastically rounded */
    int reserved;
} scgpu_moveparams;

typedef struct scgpu_sweepstats {
    int64_t trans_acc, trans_rej, rot_acc, rot_rej, cell_rej; /* cell_rej: moves rejected for leaving the cell */
    double energy_delta;                                       /* sum of accepted dE */
} scgpu_sweepstats;

/* Chain moves of the same sweep (MoveCreator::chainMove / chainDisplace / chainRotate, scOOP/mc/movecreator.cpp:304-328,
 * 1075-1256): whole molecules are displaced or rotated rigidly, energy = mol2others before and after. */
typedef struct scgpu_chainmoves {
    double chainprob;              /* Sim::chainprob: share of the sweep's trials that are chain moves (updater.cpp:215) */
    double chainm_mx[32];          /* per molecule type: stat.chainm[molType].mx    (= 2*chainmmx, sim.h:366) */
    double chainr_angle[32];       /* per molecule type: stat.chainr[molType].angle (radians, sim.h:362) */
} scgpu_chainmoves;

typedef struct scgpu_chainstats {
    int64_t chainm_acc, chainm_rej, chainr_acc, chainr_rej, cell_rej; /* cell_rej: a member outside the active cell before or after */
    double energy_delta;                                               /* sum of accepted dE */
    int64_t noop;                                                      /* picks that landed on a one-particle molecule: no move */
} scgpu_chainstats;

/* One volume move (MoveCreator::pressureMove, scOOP/mc/movecreator.cpp:330-550, ptype 0-5) for callers that run the batched
 * sweeps: positions are box-fractional, so only the box changes; both energies are full-system sums on the device. */
typedef struct scgpu_pressureparams {
    double temper, press;          /* Sim::temper, Sim::press */
    double edge_mx;                /* stat.edge.mx (= 2*edge_mx of the options file, sim.h:364) */
    int ptype;                     /* 0 anisotropic (one random edge), 1 isotropic, 2 isotropic in xy (z constant), 3 xy at constant volume,
                                      4 "anisotropic in xy" (as written in the reference: the x edge only), 5 the y edge only */
    int reserved;
} scgpu_pressureparams;

typedef struct scgpu_pressurestats {
    int accepted, reserved;
    double energy_old, energy_new; /* allToAll() before / allToAllTrial() at the proposed box */
    double enthalpy_delta;         /* accepted moves: what the reference adds to its drift sum (dE + P dV - N T ln(V'/V)), else 0 */
    double box[3];                 /* the box after the move */
} scgpu_pressurestats;

const char* scgpu_last_error(void);
int scgpu_device_count(void);

/* ctor of TotalE<> (totalenergycalculator.h:145-146) */
int scgpu_create(scgpu_ctx** out, int device);
int scgpu_destroy(scgpu_ctx* ctx);

/* global `topo` (ia_params, moleculeParam, sqmaxcut, maxcut: scOOP/structures/topo.h:21-27); table is ntypes*ntypes,
 * row-major by (type of first particle, type of second particle) */
int scgpu_set_topology(scgpu_ctx* ctx, int ntypes, const scgpu_iaparam* table, double sqmaxcut, double maxcut,
                       int nmoltypes, const scgpu_molparam* mol);
/* the [EXTER] wall potential of the topology (topo.exter, scOOP/structures/structures.h:259-275; Topo::genParamPairs / genTopoParams,
 * structures/topo.cpp:120-130, 151-152): a structureless wall in the plane z = 0 of the periodic box. Once set, every energy entry point
 * adds ExternalEnergyCalculator::extere2 (scOOP/mc/externalenergycalculator.cpp:5-105) of the particles it sums over, exactly where
 * the reference's calculators add it (mc/totalenergycalculator.h:348-350, 377-378, 410-411, 435-449), and the batched sweeps include
 * it in both energies of a trial. Call after scgpu_set_topology (the wall parameters derive from the table); exist = 0 removes it. */
int scgpu_set_exter(scgpu_ctx* ctx, int exist, double thickness, double epsilon, double attraction);
/* conf->pvec (scOOP/structures/Conf.h:305); also what initEM() / update(EMResize) need after a particle-count change.
 * type == moltype == NULL: n and every particle's type are those of the previous upload (only coordinates travel).
 * Page-locked host buffers are read by DMA directly; pageable ones are staged through the library's own pinned buffer. */
int scgpu_set_particles(scgpu_ctx* ctx, int n, const double* state30, const int* type, const int* moltype);
/* the same from the 9 doubles per particle that config.init holds (pos[3] box-fractional, dir[3], patchdir[3]); patch sides,
 * second patch and chiral axes are derived on the device as Conf::partVecInit / Particle::init do
 * (scOOP/structures/Conf.cpp:98-103, scOOP/structures/particle.cpp:3-79) */
int scgpu_set_particles_compact(scgpu_ctx* ctx, int n, const double* state9, const int* type, const int* moltype);
/* conf->geo.box, read through PairE::pbc in the reference (scOOP/mc/paire.h:1205,1211) */
int scgpu_set_box(scgpu_ctx* ctx, const double box[3]);
/* update(int target) after an accepted single-particle move (totalenergycalculator.h:326-328) */
int scgpu_update_particle(scgpu_ctx* ctx, int idx, const double* state30);
/* MoveCreator::switchTypeMove (scOOP/mc/movecreator.cpp:233-303) changes conf->pvec[target].type in place before oneToAllTrial(target) and keeps
 * the new type on acceptance: the type of ONE particle changes on the device (the state record travels as usual, through the trial state /
 * scgpu_update_particle). The kernels' specialisations follow the census of types present; the next energy call re-sorts the cells. */
int scgpu_set_particle_type(scgpu_ctx* ctx, int idx, int type);
int scgpu_download_particles(scgpu_ctx* ctx, double* state30);

/* Replaces Updater::genSimplePairList (scOOP/mc/updater.cpp:484-552): counting sort by cell. Called implicitly by
 * the energy entry points when the list is stale. */
int scgpu_build_cells(scgpu_ctx* ctx);
int scgpu_cell_assignment(scgpu_ctx* ctx, int* cell_of_particle, int ncell3[3]);
/* sorted slot -> original index, and cell_start[ncells+1] (bit-exact checks of the stable sort) */
int scgpu_cell_order(scgpu_ctx* ctx, int* order, int* cell_start);

/* oneToAllTrial(target) / oneToAll(target) (totalenergycalculator.h:355-415, 563-583). trial_state30 == NULL:
 * current state. e_pairs (optional, n doubles, indexed by original j) = the reference's `changes[]`. */
int scgpu_one_to_all(scgpu_ctx* ctx, int target, const double* trial_state30, double* e_sum, double* e_pairs);
/* m independent oneToAllTrial evaluations in one launch (one warp per trial); trial_states may be NULL */
int scgpu_one_to_all_batch(scgpu_ctx* ctx, int m, const int* targets, const double* trial_states30, double* e_sums);
/* same over every particle's current state, results stay on the device (bench: inputs resident in HBM);
 * e_host may be NULL. n_gated / n_candidates (optional) return the pair counters of that launch. */
int scgpu_one_to_all_everyone(scgpu_ctx* ctx, double* e_host, int64_t* n_candidates, int64_t* n_gated);
/* ASYNCHRONOUS whole-configuration pass for callers that stream configurations (replica workers, analysis of stored
 * trajectories): upload a configuration of the SAME particle count and types as the previous upload (state9 = what
 * config.init holds per particle: pos, dir, patchdir), derive the patch vectors on the device (Particle::init,
 * structures/particle.cpp:3-79), rebuild the cell list, evaluate oneToAll of every particle (the loop a TotalE<>::initEM()
 * + per-particle oneToAll performs, totalenergycalculator.h:135-297) and copy the n energies to e_out. Both buffers must be
 * page-locked; the call returns at once, the data is valid after scgpu_sync(ctx) returns SCGPU_OK, and state9 must not be
 * modified before that. Two contexts used alternately overlap the copies of one configuration with the kernels of another. */
int scgpu_submit_everyone(scgpu_ctx* ctx, const double* state9_pinned, double* e_out_pinned);
/* mol2others(mol) / mol2othersTrial(mol) for a molecule of m consecutive particles starting at `first`
 * (totalenergycalculator.h:417-499): members x non-members with an EMPTY conlist. trial_states30 (m*30, optional)
 * = the members' states as mutated by the caller. */
int scgpu_mol_to_others(scgpu_ctx* ctx, int first, int m, const double* trial_states30, double* e_sum);
/* allToAllTrial() / allToAll() / initEM() (totalenergycalculator.h:314-353, 502-521): every pair once, conlist of the
 * higher index. e_per_particle (optional, n) = row sums  sum_{j<i} E(i,j). */
int scgpu_all_to_all(scgpu_ctx* ctx, double* e_total, double* e_per_particle);

/* Conf::overlapAll / Conf::checkall (scOOP/structures/Conf.cpp:244-267). variant 0 = as written in the reference,
 * 1 = documented intent. */
int scgpu_overlap_one(scgpu_ctx* ctx, int target, const double* trial_state30, int variant, int* flag);
int scgpu_overlap_all(scgpu_ctx* ctx, int variant, int* flag);

/* one checkerboard sweep of displacement/rotation trials, validated statistically against sequential sweeps */
int scgpu_sweep_checkerboard(scgpu_ctx* ctx, const scgpu_moveparams* mp, uint64_t seed, uint64_t sweep,
                             scgpu_sweepstats* stats);

/* the same sweep with a share `chainprob` of its trials made chain moves (cm == NULL or chainprob == 0: identical to the call above) */
int scgpu_sweep_checkerboard_chains(scgpu_ctx* ctx, const scgpu_moveparams* mp, const scgpu_chainmoves* cm, uint64_t seed, uint64_t sweep,
                                    scgpu_sweepstats* stats, scgpu_chainstats* chain_stats);

/* volume move between sweeps; the random numbers are a pure function of (seed, step) */
int scgpu_pressure_move(scgpu_ctx* ctx, const scgpu_pressureparams* pp, uint64_t seed, uint64_t step, scgpu_pressurestats* out);

/* ---- multi-GPU: parallel-tempering replicas / Wang-Landau walkers, one process per GPU, NCCL over NVLink -------------------
 * The reference's only parallel mode is one MPI rank per replica/walker (ENABLE_MPI): MoveCreator::replicaExchangeMove
 * (scOOP/mc/movecreator.cpp:552-795) and the shared Wang-Landau arrays (scOOP/mc/wanglandau.cpp:190-212, wanglandau.h:66-123,
 * 282-289). A communicator spans `nranks` processes, each holding `nlocal` replicas on its GPU (8 replicas on 1/2/4/8 GPUs =
 * 8/4/2/1 per GPU). NCCL is loaded at run time (libnccl.so.2) and only when nranks > 1. */
typedef struct scgpu_comm scgpu_comm;
#define SCGPU_UNIQUE_ID_BYTES 128
#define SCGPU_REPLICA_PAYLOAD 40
#define SCGPU_REPLICA_MOLTYPES 8
#define SCGPU_MAX_LOCAL_REPLICAS 16

/* rank 0 draws the NCCL unique id (ncclGetUniqueId), the host program hands it to the other ranks (MPI_Bcast in the reference's
 * own main(), a torch.distributed store in ours) and every rank creates the communicator (ncclCommInitRank; replaces MPI_Init +
 * MPI_Comm_rank/size, scOOP/main.cpp:37-45). nranks == 1 needs no id and no NCCL. */
int scgpu_comm_unique_id(char id[SCGPU_UNIQUE_ID_BYTES]);
int scgpu_comm_create(scgpu_comm** out, int device, int nranks, int rank, const char id[SCGPU_UNIQUE_ID_BYTES]);
/* the same around an ncclComm_t the caller already owns (it is not destroyed by scgpu_comm_destroy) */
int scgpu_comm_attach(scgpu_comm** out, int device, void* nccl_comm, int nranks, int rank);
int scgpu_comm_destroy(scgpu_comm* comm);

/* What a replica owns and what travels on an accepted exchange: temperature, pressure, pseudo-rank and a payload (the reference
 * swaps its whole Statistics block, i.e. the adapted step sizes and counters that belong to the temperature:
 * movecreator.cpp:670-671, 766-767; the caller packs what it wants to travel -- per-type trans/rot maxima, edge maximum). */
typedef struct scgpu_replica_state {
    double temper, press;                       /* Sim::temper, Sim::press */
    int pseudo_rank;                            /* Sim::pseudoRank: position on the temperature ladder */
    int replica;                                /* out: global replica index = rank * nlocal + local index (MpiExchangeData::mpiRank) */
    int64_t wl_order[2];                        /* wl.currorder (only read when Wang-Landau weights are passed) */
    double part_num[SCGPU_REPLICA_MOLTYPES];    /* molecules per molecule type (grand-canonical term, movecreator.cpp:736-739) */
    double payload[SCGPU_REPLICA_PAYLOAD];
    /* results of the last call */
    int attempted, accepted;                    /* this replica was part of a pair / the pair was swapped */
    int partner;                                /* global replica index of the partner, -1 if none */
    int reserved;
    int64_t partner_wl_order[2];                /* the partner's wl.currorder (the lower replica adopts it on acceptance, :757-762) */
    double change;                              /* the exponent of the acceptance rule (:722-745) */
    double energy, volume;                      /* this replica's allToAll() and box volume as used in the rule */
    double edrift;                              /* what the reference adds to its drift sum on acceptance (:676-684, 751-755) */
} scgpu_replica_state;

typedef struct scgpu_exchangeparams {
    int nrepchange;                             /* Sim::nrepchange: decides which neighbours pair up on this sweep (:616-623) */
    int wl_len;                                 /* 0, or length of every replica's Wang-Landau weight array (wl.length[0] * max(1, wl.length[1])) */
    int64_t wl_len0;                            /* wl.length[0] (index = order[0] + order[1] * wl_len0) */
    double dtemp, dpress;                       /* Sim::dtemp, Sim::dpress (sim.h:389-399) */
    double chempot[SCGPU_REPLICA_MOLTYPES];     /* MoleculeParams::chemPot where activity != -1, else 0 */
    uint64_t seed;
} scgpu_exchangeparams;

/* MoveCreator::replicaExchangeMove for the `nlocal` replicas of this process (ctxs[k] holds replica k's configuration, states[k]
 * its thermodynamic state; in/out). Per call: allToAll() of every local replica on its own stream, a kernel packs the records
 * {E, V, N, T, P, pseudoRank, wl order, particle numbers, payload} on the device (the energy never visits the host), ONE
 * ncclAllGather (replaces MPI_Alltoall + 4 point-to-point messages per pair), a device kernel takes every pair's decision -- the
 * reference's odd/even pairing and acceptance rule with a counter-based uniform keyed on (seed, sweep, lower pseudo-rank), hence
 * identical on every rank -- and swaps {T, P, pseudoRank, payload}; the states of the local replicas come back in one small
 * page-locked copy. wl_weights: NULL, or nlocal arrays of wl_len doubles (host) for the Wang-Landau term of the rule. */
int scgpu_replica_exchange(scgpu_comm* comm, int nlocal, scgpu_ctx* const* ctxs, scgpu_replica_state* states,
                           const scgpu_exchangeparams* params, uint64_t sweep, const double* const* wl_weights);
/* device time of the last scgpu_replica_exchange after the energy kernels: pack + all-gather + decision + copy back */
int scgpu_comm_last_exchange_us(scgpu_comm* comm, float* us);

/* Multiple-walker Wang-Landau: the reference's walkers share weights/hist through an MPI-3 shared window and update them in
 * place (WangLandau::accept, wanglandau.h:282-289); here every walker applies its accepts to its own copy and the walkers merge
 * every K sweeps: delta = mine - base, ncclAllReduce(sum) of the deltas, base += sum, then WangLandau::update
 * (wanglandau.h:66-123: flatness test, alpha /= 2, weights -= wmin, hist = 0) evaluated on the device -- identically on every
 * rank -- and {alpha, min, wmin} broadcast from rank 0 as the reference's shared_A_min_wmin. */
typedef struct scgpu_wlstate {
    double alpha;                               /* in/out: wl.alpha */
    double wmin;                                /* out */
    int64_t min, max;                           /* out: histogram extremes */
    int halved, converged;                      /* out: alpha was halved on this call / alpha < WL_ALPHATOL (update returns true) */
} scgpu_wlstate;
int scgpu_wl_merge(scgpu_comm* comm, int len, double* weights, int64_t* hist, double* weights_base, int64_t* hist_base,
                   double temper, scgpu_wlstate* st);

/* ---- Wang-Landau order parameters of the whole configuration, evaluated on the device -----------------------------------------
 * The from-scratch forms the reference computes in WangLandau::init (scOOP/mc/wanglandau.cpp:56-125) and after every volume or
 * type-switch move (WangLandau::runPress / runSwitch, scOOP/mc/wanglandau.h:168-196, 220-238), for callers whose configuration
 * lives on the device (scgpu_sweep_checkerboard*, scgpu_pressure_move): no download of the particles, one small read-back.
 *   wlm 1  z of particle 0 from the system centre of mass: Conf::massCenter (scOOP/structures/Conf.cpp:78-95) + zOrder (wanglandau.h:551-554)
 *   wlm 2  largest hole of the membrane in the xy plane: Mesh::meshInit = meshFill + findHoles (scOOP/mc/mesh.cpp:11-185), holeXYPlane(wli) (wanglandau.h:313-317)
 *   wlm 3  z component of particle 0's axis: zOrient (wanglandau.h:321-324)
 *   wlm 4  xy distance of particles 0 and 1: twoPartDist (wanglandau.h:395-398)
 *   wlm 7  particles of type wlmtype within sqrt(WL_CONTACTS) of particle 0: contParticlesAll (wanglandau.h:603-615, 652-678)
 *   wlm 8, 9  box edge x / y: boxSize_x / boxSize_y (wanglandau.h:375-382)
 * wlm 5 and 6 (pore radius, radiusholeAll, wanglandau.cpp:306-338) are refused: the reference never allocates the array they fill.
 * order[] is binned exactly as the reference bins (ceil for 1, 4, 7, 8, 9; floor for 3; truncation for 2); the caller compares it
 * with wl.length[] and looks up its weights. The incremental forms used by single-particle moves (meshOrderMoveMolecule,
 * contParticlesMoveMolecule) remain host logic of the reference: they update this state trial by trial. */
typedef struct scgpu_wlorder {
    int wlm[2];                                 /* in: Sim::wlm[] (0 = dimension not used) */
    int wlmtype;                                /* in: Sim::wlmtype, the particle type wlm 2 and 7 look at */
    int reserved;
    double minorder[2], dorder[2];              /* in: wl.minorder[], wl.dorder[] (wl.dat) */
    double meshsize;                            /* in: wl.wl_meshsize (wlm 2; the reference uses sigma(wlmtype) / 3, wanglandau.cpp:79) */
    int64_t order[2];                           /* out: wl.neworder[] */
    double raw[2];                              /* out: the quantity before binning (z distance, hole size in mesh points, dir.z, distance, contacts, edge) */
    double syscm[3], sysvolume;                 /* out: conf->syscm and conf->sysvolume of the current configuration */
    int mesh_dim[2];                            /* out: Mesh::dim (wlm 2) */
    int64_t mesh_occupied, mesh_skipped;        /* out: occupied mesh points; particles whose INBOX() coordinate was exactly 1 (the reference
                                                   writes one past its mesh row for them: undefined there, skipped here) */
} scgpu_wlorder;
int scgpu_wl_order(scgpu_ctx* ctx, scgpu_wlorder* io);
/* Mesh::data as Mesh::meshInit leaves it (scOOP/mc/mesh.cpp:11-35, 135-185) for the mesh of the last scgpu_wl_order call with wlm 2:
 * occupied points hold minus the number of particles covering them, every free point the number of its hole (1, 2, ... in the order
 * the reference's scan meets the holes). len = mesh_dim[0] * mesh_dim[1]. The reference's incremental updates of single-particle and
 * chain moves (Mesh::addPart / removePart, WangLandau::meshOrderMoveMolecule, scOOP/mc/wanglandau.cpp:7-52) carry on from this array:
 * a caller that lets the device do the from-scratch search after a volume move hands it to its host-side Mesh (integration/wl_gpu_hook.h). */
int scgpu_wl_mesh(scgpu_ctx* ctx, int* data, int len);

/* measurement helpers (CUDA events on the context's stream; FP64 FMA-chain peak microbenchmark) */
int scgpu_timer_start(scgpu_ctx* ctx);
int scgpu_timer_stop(scgpu_ctx* ctx, float* ms);
/* waits for the context's stream. Returns SCGPU_ERR_STATE when an asynchronous energy launch (e_host == NULL forms,
 * scgpu_submit_everyone) could not complete -- a work list overflowed and has been grown, or the thread-per-target gate met
 * a layout it cannot hold and the library has switched those targets (or the whole pass) to the cell gate: the results of
 * that launch are invalid, repeat the call (synchronous calls repeat internally and never report this). */
int scgpu_sync(scgpu_ctx* ctx);
int scgpu_fp64_peak(scgpu_ctx* ctx, double* tflops);
/* one scgpu_one_to_all_everyone pass with an event between its launches: microseconds of {gate, cheap terms, patch terms, combine} */
int scgpu_profile_everyone(scgpu_ctx* ctx, float us[4]);
int scgpu_flush_l2(scgpu_ctx* ctx);
int scgpu_kernel_launches(scgpu_ctx* ctx, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
