/* pic_b200.h -- C ABI of libpic_b200.so: the B200 (sm_100a) implementation of
 * pyPIC's per-timestep particle-in-cell loop.
 *
 * The reference (drobnyjt/pyPIC) has no FFI: its hot path is the set of
 * module-level Python functions cited next to each entry point below.  The
 * drop-in Python modules in this repository (pypic.py, PIC_L.py, PIC_L_DD.py,
 * pygcpic.py) bind these symbols through ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PIC_ERR_* code and never
 *     throws; pic_last_error() returns a static description of the last failure
 *     on the calling thread;
 *   - "pic_dev_*"  : all array arguments are DEVICE pointers (fp64 unless noted),
 *     the call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - "pic_host_*" : all array arguments are HOST pointers; the call uploads,
 *     launches the same pic_dev_* kernels, downloads and synchronises;
 *   - no allocation is done by pic_dev_* calls; the caller owns every buffer;
 *   - all arithmetic is IEEE fp64 (the reference's precision); indices are int32;
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef PIC_B200_H
#define PIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIC_OK 0
#define PIC_ERR_CUDA (-1)     /* a CUDA runtime call / kernel launch failed        */
#define PIC_ERR_ARG (-2)      /* invalid argument (null pointer, bad size, ...)     */
#define PIC_ERR_NODEVICE (-3) /* no CUDA device visible                            */
#define PIC_ERR_RANGE (-4)    /* a particle index fell outside the grid (reference
                                 undefined behaviour)              */

const char* pic_last_error(void);
int pic_version(void);
/* number of SMs / device name of the current device (diagnostics for bench.py) */
int pic_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* Raw device-memory helpers for hosts that hold plain pointers (asynchronous on
 * `stream` except pic_dev_read, which synchronises the stream before returning). */
int pic_dev_read(const void* dev, void* host, int64_t bytes, void* stream);
int pic_dev_write(void* dev, const void* host, int64_t bytes, void* stream);
int pic_dev_zero(void* dev, int64_t bytes, void* stream);
int pic_dev_copy(void* dst, const void* src, int64_t bytes, void* stream);
/* Start of a sheath timestep (PIC_L_DD.py:452-456: Es = E0, k = 0, r = 1): Es <- E0 and the cumulative
 * wall counts (4 doubles), the step statistics (nstats doubles) and the loop flag cleared, in one call */
int pic_dev_dd_step_begin(double* Es, const double* E0, int Ng, double* wall_cum, double* stats, int64_t nstats,
                          int32_t* ctl, void* stream);
int pic_stream_sync(void* stream);
/* frees the device workspaces cached by the pic_host_* entry points */
int pic_host_release(void);

/* ------------------------------------------------------------------------- *
 * Grid kernels (one CTA; Ng is small).  F/out are DEVICE fp64[n].
 * ------------------------------------------------------------------------- */
/* variant 0: periodic (pypic.smooth_field_p pypic.py:63-76)
 * variant 1: ends pinned (PIC_L_DD.smoothField :216-221, Grid.smooth_rho pygcpic.py:1055-1060) */
int pic_dev_smooth(const double* F, double* out, int n, int variant, void* stream);
/* variant 0: pypic.differentiate_p pypic.py:185-214        (+dF/dx, periodic, *(0.5/dx))
 * variant 1: PIC_L_DD.differentiateField :192-203           (-dF/dx, one-sided ends, /dx*0.5)
 * variant 2: PIC_L.differentiateFieldPeriodic :235-246      (-dF/dx, periodic over n nodes, /dx*0.5)
 * variant 3: Grid.differentiate_phi_to_E_dirichlet pygcpic.py:932-936 (-dF/dx, /dx/2.) */
int pic_dev_differentiate(const double* F, double* out, int n, double dx, int variant, void* stream);
/* PIC_L_DD.integrateField :205-214 : out[i] = -trapz(F[:i+1],dx) as a block scan;
 * if subtract_max != 0 also applies the caller's ``- max`` (PIC_L_DD.py:519,523). */
int pic_dev_integrate_field(const double* F, double* out, int n, double dx, int subtract_max, void* stream);
/* out = F - max(F) (mode 0) or F - min(F) (mode 1): pypic.py:553, pygcpic.py:1002,1052 */
int pic_dev_shift_extreme(const double* F, double* out, int n, int mode, void* stream);

/* Tridiagonal solve by parallel cyclic reduction (PCR), no pivoting.
 * a=sub (a[0] ignored), b=diag, c=super (c[n-1] ignored), d=rhs -> x.  n<=PIC_PCR_SMEM_MAX
 * runs in one CTA in shared memory; larger n uses global-memory PCR passes and
 * `work` (8*n doubles) must be provided. */
#define PIC_PCR_SMEM_MAX 6144
int pic_dev_tridiag_pcr(const double* a, const double* b, const double* c, const double* d,
                        double* x, int n, double* work, void* stream);
/* pypic.solve_poisson_p pypic.py:359-382 / PIC_L.solvePoissonPeriodicElectronsNeutralized
 * PIC_L.py:208-220: periodic [1,-2,1] phi = -dx^2(c0+c2), c0=-mean(rho)/eps0, c2=rho/eps0,
 * gauge phi[n-1]=0 then ``- max(phi)`` if subtract_max.  work: 13*n doubles. */
int pic_dev_poisson_periodic(const double* rho, double* phi, int n, double dx, int subtract_max,
                             double* work, void* stream);
/* Grid.solve_for_phi_dirichlet pygcpic.py:987-1003 (rows 0,n-1 identity, no eps0, -min). work 13*n */
int pic_dev_poisson_dirichlet(const double* rho, double* phi, int n, double dx, double* work, void* stream);
/* Newton-Boltzmann solves, whole Newton loop in ONE kernel launch (n <= PIC_PCR_SMEM_MAX):
 * bc 0: Grid.solve_for_phi_dirichlet_boltzmann pygcpic.py:1005-1053 (cold start, c2=rho/eps0,
 *       stop dphi.dphi<=tol, exact tridiagonal step instead of bicgstab)
 * bc 1: Grid.solve_for_phi_dirichlet_neumann_boltzmann :1062-1109 (warm start from phi,
 *       src = n (number density, c2=e*n/eps0), last row [1,-4,3], stop |dphi|<=tol)
 * iters_out: device int[1] receiving the Newton iteration count. */
int pic_dev_newton_boltzmann(const double* src, double* phi, int n, double dx, double n0, double Te,
                             int bc, double tol, int iter_max, int* iters_out, void* stream);
/* The Boltzmann-Newton solves of the other two codes, as written there (reference node n/2,
 * `while resid > tol and k <= maxiter`, resid = |dphi|_2, start from the phi passed in):
 * periodic 0: PIC_L.solvePoisson :146-177 == PIC_L_DD.solvePoisson :116-147 (laplacian1D with its
 *             three-entry last row [1,1,-2], F[0]=phi[0], F[-1]=phi[-1]);
 * periodic 1: PIC_L.solvePoissonPeriodic :179-206 (n = Ng+1) == PIC_L_DD.solvePoissonPeriodic
 *             :149-176 (n = Ng): cyclic system by Sherman-Morrison, work = n doubles of scratch.
 * The reference inverts J with scipy.sparse.linalg.inv; here PCR.  n <= PIC_PCR_SMEM_MAX. */
int pic_dev_newton_boltzmann_l(const double* rho, double* phi, int n, double dx, double kBT, double tol, int maxiter,
                               int periodic, double* work, int* iters_out, void* stream);

/* ------------------------------------------------------------------------- *
 * Host draw service of the RNG-parity mode (no GPU involved): NumPy's legacy MT19937 stream
 * (np.random.get_state(): key uint32[624], pos in [0,624], has_gauss, cached_gaussian), which
 * PIC_L_DD.py:419-450 consumes in particle index order.
 *   pic_mt_jump_poly   g(t) = t^nwords mod phi(t) (phi: characteristic polynomial of MT19937,
 *                      found once by Berlekamp-Massey), 19937 coefficients packed in uint32[624];
 *   pic_mt_jump        advances (key,pos) by the number of 32-bit outputs g encodes (~1 ms however
 *                      large the jump): the thermostat's one uniform per ACTIVE particle
 *                      (PIC_L_DD.py:421; two words each) is skipped without generating it;
 *   pic_mt_skip        the same by plain generation (small remainders);
 *   pic_mt_sheath_draws  per dead slot x = uniform(0,L), u,v,w = normal(0,sigma[k]) exactly as
 *                      np.random.uniform / np.random.normal would return them (:433-436, :443-446);
 *                      xd == NULL only advances the stream (slots owned by other ranks);
 *   pic_mt_sheath_thermostat  gamma != 0 (:419-427): n_active uniforms in index order, every
 *                      u < gamma followed by three normals of sigma0 (ordinal < k_split) / sigma1;
 *                      the hits' ordinals and draws are returned (capacity cap; on overflow the state
 *                      is left untouched, *nhits holds the count and PIC_ERR_ARG is returned).
 * ------------------------------------------------------------------------- */
int pic_mt_jump_poly(uint64_t nwords, uint32_t* g624);
int pic_mt_jump(uint32_t* key624, int32_t* pos, const uint32_t* g624);
int pic_mt_skip(uint32_t* key624, int32_t* pos, uint64_t nwords);
int pic_mt_sheath_draws(uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t n,
                        const double* sigma, double L, double* xd, double* ud, double* vd, double* wd);
int pic_mt_sheath_thermostat(uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t n_active,
                             int64_t k_split, double gamma, double sigma0, double sigma1, int64_t cap,
                             int64_t* hit_k, double* hu, double* hv, double* hw, int64_t* nhits);

/* ------------------------------------------------------------------------- *
 * PIC_L_DD.py -- bounded two-species implicit sheath (the benchmark path)
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t N;        /* particles in this shard                                   */
    int64_t n_split;  /* local index where species 2 starts: [0,n_split) use q[0],m[0] */
    int32_t Ng;       /* grid NODES (PIC_L_DD.py:324)                               */
    int32_t flags;    /* 0 (default): register-window deposit over contiguous chunks, tiles in
                         shared memory; bit0: plain shared-memory atomics per particle (the
                         "fast atomicAdd" variant benchmarked alongside); bit2: grid-stride
                         kernel with warp-uniform pre-reduction; bit3: the register-prefetch
                         (non-TMA) build of the window kernel; bit1: keep the grid tiles in
                         global memory (grids too large for shared memory); bit4: force the
                         large-grid build of the window kernel (per-warp field windows of 32
                         nodes instead of the whole-grid tile, deposit windows flushed every 1-16
                         rows) -- taken automatically when the grid does not fit shared memory;
                         bit5: pic_dev_dd_sort_by_cell takes its global-memory cursor path (faster
                         on a nearly sorted store); bit6: static round-robin of whole chunks over
                         the CTAs instead of dynamically scheduled 1024-particle slices;
                         bit7: REPRODUCIBLE build of the default window kernel -- every addition to
                         the global accumulators is made on a pair of 64-bit fixed-point words
                         (integer atomics: the sum does not depend on the order of the additions,
                         i.e. not on scheduling, grid size or the number of ranks).  `acc` then has
                         2*Ng+4 doubles FOLLOWED by 4*Ng int64 words [hi(2Ng) | lo(2Ng)], all zero
                         on entry; pic_dev_dd_field_update / pic_dev_dd_j1_finish convert the words
                         (one rounding per node) and clear them.  A sharded run all-reduces the
                         words as int64 and the four counts as fp64.  One hi unit is 2^-s with
                         2^(31-s) > max|q|*p2c/dx*c (c = speed of light), lo adds 32 bits; a
                         contribution 2048 times beyond that bound is counted in range_err. */
    double dx, dt, L, p2c;
    double q[2], m[2];
} pic_dd_params;

/* Function-level drop-ins (per-particle q, active as fp64 1/0/-1 like the reference):
 * PIC_L_DD.interpolateField :32-39 (vectorised over x) */
int pic_dev_dd_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng, double dx,
                           int* range_err, void* stream);
/* PIC_L_DD.weightCurrents :41-68 (v != NULL) / weightDensities :70-88 (v == NULL).
 * out fp64[Ng] is overwritten; wall-charge terms and edge fold included for currents. */
int pic_dev_dd_weight(const double* x, const double* q, const double* v, const double* active,
                      double* out, int64_t N, int Ng, double dx, double dt, double p2c,
                      int* range_err, void* stream);

/* One Picard iteration of PIC_L_DD.main_i's particle phase, fused
 * (gather :470-474, CN push :477-491, absorb :494-505, deposit of jh and j1 :509,513):
 *   x0,u0   : state at time n (read only)
 *   x1,u1   : scratch = state at n+1 of the previous iteration (read unless `first`,
 *             always written); the half-step position is recomputed as (x0+x1)*0.5
 *   active  : int8 {1,0,-1}; read unless `first`, written when a particle is absorbed
 *   Es      : field used for the gather (fp64[Ng])
 *   acc     : fp64[2*Ng+4] accumulators, must be zero on entry:
 *             [0,Ng) raw CIC jh, [Ng,2Ng) raw CIC j1, then the numbers of particles
 *             absorbed IN THIS ITERATION: left-wall sp1, left-wall sp2, right-wall sp1,
 *             right-wall sp2 (exact integers stored as fp64 so one allreduce covers all)
 *   range_err: device int, incremented for every out-of-grid index (clamped). */
int pic_dev_dd_picard_iter(const pic_dd_params* p, const double* x0, const double* u0,
                           double* x1, double* u1, int8_t* active, const double* Es,
                           double* acc, int first, int* range_err, void* stream);
/* The same iteration with the n+1 positions PING-PONGED between two buffers (x1_in is read,
 * x1_out written; they may be equal) and the velocity store OPTIONAL: with u1 == NULL the
 * iteration streams 32 instead of 40 bytes per particle.  u1 is only needed by the commit, so
 * the host asks for it in the iteration it expects to be the last one (the contraction of the
 * residual is very regular) and otherwise repairs it with pic_dev_dd_commit_u. */
int pic_dev_dd_picard_iter2(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, void* stream);
/* Enqueue-ahead variant: `done` (device int32, may be NULL) is read at kernel entry and a non-zero
 * value makes the launch a no-op.  Together with pic_dev_dd_field_update2, which raises the flag
 * when the loop condition `r > tol and k < maxiter` (PIC_L_DD.py:452) fails, the host can queue
 * the iterations it expects back to back and read the outcome once per step instead of once per
 * iteration. */
int pic_dev_dd_picard_iter3(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, void* stream);
/* u1 of the last iteration after the fact: x1_prev/x1_last are that iteration's input and output
 * positions, Es the field it gathered with, `first` whether it was the first iteration of the
 * step.  Particles absorbed before it get the reference's 0.0 (PIC_L_DD.py:459-462). */
/* pic_dev_dd_picard_iter3 with the ABSORPTION LOG: every particle this launch absorbs appends the
 * int32 quadruple {slot, original index (orig[slot], or the slot when orig == NULL), iteration, 0}
 * to dead_buf = int32[4 + 4*dead_cap] (16-byte aligned), whose word 0 is the device counter (it may
 * exceed dead_cap: entries beyond the capacity are dropped and the caller falls back to scanning
 * the flags).  The re-injection visits the logged slots instead of scanning N flags, in the
 * reference's index order, and the iteration number orders the vionout tally (PIC_L_DD.py:497-503). */
int pic_dev_dd_picard_iter4(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, int32_t* dead_buf, int32_t dead_cap,
                            const int32_t* orig, int32_t iteration, void* stream);
/* ... and with the velocity moments of the step's INITIAL state for free: the first iteration (first != 0)
 * streams u0 anyway, so with moments != NULL it adds sum(u0) and sum(u0*u0) over its particles to
 * moments[0..1] (zero on entry) -- np.std(u0) of PIC_L_DD.py:417 and the kinetic energy of :549 without a
 * pass of their own.  The sums describe the state AFTER re-injection; pic_dev_dd_apply_draws3 accumulates
 * the correction back to the state before it. */
int pic_dev_dd_picard_iter5(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_in,
                            double* x1_out, double* u1, int8_t* active, const double* Es, double* acc, int first,
                            int* range_err, const int32_t* done, int32_t* dead_buf, int32_t dead_cap,
                            const int32_t* orig, int32_t iteration, double* moments, void* stream);
int pic_dev_dd_commit_u(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_prev,
                        const double* x1_last, const int8_t* active, const double* Es, double* u1, int first,
                        int* range_err, void* stream);
/* Light iterations: pic_dev_dd_picard_iter2 called with u1 == NULL neither stores the velocities nor
 * deposits j1 (the current at n+1, PIC_L_DD.py:513, is only used after the Picard loop).  If the loop
 * ends on such an iteration, pic_dev_dd_commit_u2 recomputes u1 AND deposits the survivors' j1 into
 * acc[Ng..2Ng) (acc != NULL), and -- after the caller's all-reduce of acc when sharded --
 * pic_dev_dd_j1_finish applies the wall terms from the cumulative counts and the edge fold
 * (PIC_L_DD.py:55-66), writes j1[Ng] and stats[1] = mean(j1), and zeroes the accumulator. */
int pic_dev_dd_commit_u2(const pic_dd_params* p, const double* x0, const double* u0, const double* x1_prev,
                         const double* x1_last, const int8_t* active, const double* Es, double* u1,
                         int first, double* acc, int* range_err, void* stream);
int pic_dev_dd_j1_finish(const pic_dd_params* p, double* acc, const double* wall_cum, double* j1,
                         double* stats, void* stream);
/* Device self test: compares the constant-divisor division and the fast cell lookup used
 * on the hot path against the IEEE operations for n pseudo-random / adversarial operands;
 * *mismatches_dev (device uint64, zeroed by the caller) must stay 0. */
int pic_dev_selftest_div(double b, uint64_t n, uint64_t seed, uint64_t* mismatches_dev, void* stream);
/* Debug aid: when buf != NULL (device uint64[2*gridDim]) the fused Picard kernels record per-CTA
 * (start,end) %globaltimer values; pass NULL to switch it off. */
int pic_dev_debug_cta_timer(uint64_t* buf);
/* Field phase of the same iteration (PIC_L_DD.py:55-66,516-527), one CTA:
 *   wall_cum fp64[4] += acc[2Ng..2Ng+3]; jh,j1 get wall terms + edge fold;
 *   E1 = E0 + (dt/eps0)(mean(jh) - jh); Eh=(E1+E0)/2; r=|Es-Eh|_2; Es=Eh;
 *   acc is zeroed for the next iteration.  stats (device fp64[8]):
 *   [0]=r, [1]=mean(j1), [2]=sum(eps0*E1^2*dx/2), [3]=iteration counter (incremented);
 *   [4..7] must be zero on entry: reduction scratch of the cooperative multi-CTA build used
 *   for Ng > 32768 (its sums are re-associated relative to the one-CTA kernel). */
int pic_dev_dd_field_update(const pic_dd_params* p, double* acc, double* wall_cum, const double* E0,
                            double* Es, double* E1, double* j1, double* stats, void* stream);
/* The same field phase for the enqueue-ahead loop.  All extra arguments may be NULL / 0:
 *   Es_prev fp64[Ng] receives the field the iteration gathered with (Es on entry) -- what
 *           pic_dev_dd_commit_u2 needs if the loop ends on a light iteration;
 *   rhist   fp64[maxiter]: rhist[k-1] = residual of the k-th iteration of the step (k = stats[3]);
 *   ctl     device int32: a non-zero value on entry makes the launch a no-op; set to 1 on exit
 *           when not (r > tol) or k >= maxiter, i.e. when PIC_L_DD.py:452 leaves the loop. */
int pic_dev_dd_field_update2(const pic_dd_params* p, double* acc, double* wall_cum, const double* E0,
                             double* Es, double* E1, double* j1, double* stats, double* Es_prev,
                             double* rhist, int32_t* ctl, double tol, int maxiter, void* stream);
/* Particle decomposition without a library collective: the field kernel itself exchanges and sums the
 * ranks' accumulators over NVLink peer memory (one process per GPU; buffers shared through CUDA IPC).
 *   pic_p2p_alloc   allocates this rank's buffer -- nacc doubles (the accumulators the particle
 *                   kernels add to, zero; nacc even), an inbox of 2 x world x nacc doubles the PEERS
 *                   write to, 2*world + 2 uint32 of flags / counters -- and writes its 64-byte IPC handle;
 *   pic_p2p_open    maps another rank's buffer from its handle (exchange the handles with any host
 *                   collective); pic_p2p_close / pic_p2p_free undo the two.
 *   pic_dev_dd_field_update_p2p = pic_dev_dd_field_update2 with the all-reduce inside: peers_dev is
 *                   a DEVICE array of the world buffer pointers in rank order (own entry = the local
 *                   pointer).  Every rank pushes its accumulators into its slot of every rank's
 *                   inbox (posted NVLink stores, double-buffered by the parity of a reduction counter
 *                   kept on the device), zeroes them, raises a flag in every rank's buffer, waits for
 *                   the world flags in its OWN buffer and sums the slots in rank order from local
 *                   memory into acc_sum (fp64[2*Ng+4], local) -- the same bits on every rank -- which
 *                   the field phase consumes.  seq is ignored (kept for ABI stability).  Waits are
 *                   bounded (seconds); *err (device int) is set to 1 on a time-out.  Ng <= 32768.
 *   pic_dev_p2p_reduce  the reduction alone (sum[nacc] local), e.g. for the j1 repair pass. */
int pic_p2p_alloc(int64_t nacc, int world, void** dev_ptr, void* handle64);
int pic_p2p_open(const void* handle64, void** peer_ptr);
/* bound of the flag waits: 2^log2_spins polls of >= 200 ns (default 25, i.e. 7 s or more) */
int pic_p2p_set_timeout(int log2_spins);
int pic_p2p_close(void* peer_ptr);
int pic_p2p_free(void* dev_ptr);
int pic_dev_p2p_reduce(const double* const* peers_dev, int rank, int world, uint32_t seq, int64_t nacc,
                       double* sum, int* err, void* stream);
int pic_dev_dd_field_update_p2p(const pic_dd_params* p, const double* const* peers_dev, int rank, int world,
                                uint32_t seq, double* acc_sum, double* wall_cum, const double* E0, double* Es,
                                double* E1, double* j1, double* stats, double* Es_prev, double* rhist,
                                int32_t* ctl, double tol, int maxiter, int* err, void* stream);
/* Field phase of the same iteration on SPATIAL SLABS (pypic_b200/spatial.py, BASELINE config 5; formulas of
 * pic_dev_dd_field_update2 = PIC_L_DD.py:55-66,516-527, evaluated by every rank on its own nodes and `guard`
 * nodes either side).  Rank r owns the cells [c0, c1); its particles deposit on the nodes [c0-guard, c1+guard].
 * Per iteration: pic_dev_slab_pack -> all-gather of the messages -> pic_dev_slab_field_update -> all-gather of the
 * two partial sums -> pic_dev_slab_finish.  The collectives are the caller's (NCCL); every kernel is a no-op when
 * *ctl != 0 (enqueue-ahead loop).
 *   pic_slab_message_len(guard)  doubles per rank message: the raw jh, j1 of the 2*guard+1 nodes around each
 *                       slab boundary, the sums of all raw deposits of the rank (+ its fold nodes), the four
 *                       absorbed counts of the iteration and the fold nodes j[1], j[Ng-2];
 *   pic_slab_work_len()  doubles of device scratch (zero-initialised once by the caller);
 *   pic_dev_slab_pack    acc fp64[2Ng+4] (raw accumulators of the particle kernels) -> msg; adds the four counts to
 *                       absorbed_local (fp64[4], may be NULL: this rank's own absorptions of the step) and clears them;
 *   pic_dev_slab_field_update  gathered = the messages of all ranks in rank order; completes jh, j1 on the band,
 *                       applies wall terms and fold, E1 = E0 + (dt/eps0)(mean(jh) - jh), Eh, Es := Eh and j1 on the
 *                       band; clears the band of acc; partial[2] = (sum over OWNED nodes of (Es-Eh)^2, field energy);
 *                       neighbours compute identical bits on the nodes they share;
 *   pic_dev_slab_finish partials = partial[2] of all ranks in rank order: wall_cum += counts, stats[0..3] and rhist
 *                       as pic_dev_dd_field_update2, raises *ctl when the loop ends (same value on every rank). */
int pic_slab_message_len(int guard);
int pic_slab_work_len(void);
int pic_dev_slab_pack(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, double* acc, double* msg,
                      double* work, double* absorbed_local, const int32_t* ctl, void* stream);
int pic_dev_slab_field_update(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, double* acc,
                              const double* gathered, const double* wall_cum, const double* E0, double* Es, double* E1,
                              double* j1, double* partial, double* work, const int32_t* ctl, void* stream);
int pic_dev_slab_finish(const pic_dd_params* p, int c0, int c1, int guard, int rank, int world, const double* gathered,
                        const double* partials, double* wall_cum, double* stats, double* rhist, int32_t* ctl, double tol,
                        int maxiter, void* stream);
/* Re-injection (PIC_L_DD.py:429-450).  Host-RNG parity mode: compact the indices of
 * inactive slots in index order (pic_dev_compact_flags), draw on the host with the
 * legacy MT19937 stream, then scatter: */
int pic_dev_dd_apply_draws(const int32_t* idx, const double* xd, const double* ud, const double* vd,
                           const double* wd, int64_t n, double* x0, double* u0, double* v0, double* w0,
                           int8_t* active, void* stream);
/* The same for a cell-sorted store that keeps the reference's particle numbering: slot[t] is where
 * the particle sits now, orig[t] its original index (NULL: = slot).  x0,u0,active are written at
 * the slot, the passive v0,w0 -- kept in ORIGINAL order, they are never streamed by the Picard
 * kernels -- at the original index.  xd == NULL / active == NULL leave x0 / the flags alone (the
 * thermostat, PIC_L_DD.py:419-427, redraws only the velocities). */
int pic_dev_dd_apply_draws2(const int32_t* slot, const int32_t* orig, const double* xd, const double* ud,
                            const double* vd, const double* wd, int64_t n, double* x0, double* u0, double* v0,
                            double* w0, int8_t* active, void* stream);
/* pic_dev_dd_apply_draws2 that also accumulates corr[0] += sum(u_new - u_old), corr[1] += sum(u_new^2 -
 * u_old^2) over the slots it rewrites (corr may be NULL) */
int pic_dev_dd_apply_draws3(const int32_t* slot, const int32_t* orig, const double* xd, const double* ud,
                            const double* vd, const double* wd, int64_t n, double* x0, double* u0, double* v0,
                            double* w0, int8_t* active, double* corr, void* stream);
/* Start of a sheath timestep in one launch: re-injection of the slots the previous step absorbed
 * (PIC_L_DD.py:429-450) and the per-step clears of pic_dev_dd_step_begin.
 *   philox != 0: device draws for the slots named in `log` (the absorption log the previous step's
 *     iterations wrote; flag scan when it overflowed), as pic_dev_dd_reinject_philox_log / _philox2;
 *   n_draws > 0: host draws (legacy MT19937 order) applied as pic_dev_dd_apply_draws3 does;
 *   next_log: header of the log the coming step writes (a different buffer than `log`), cleared. */
typedef struct {
    const int32_t* log; int32_t* next_log; int32_t log_cap; int32_t philox;
    double sigma[2]; uint64_t seed, step; int64_t global_offset;
    const int32_t* slot; const int32_t* orig_of_draw; const double* xd; const double* ud; const double* vd; const double* wd;
    int64_t n_draws; double* corr;
    double* x0; double* u0; double* v0; double* w0; int8_t* active; const int32_t* orig;
    double* Es; const double* E0; double* wall_cum; double* stats; int64_t nstats; int32_t* ctl;
} pic_dd_prologue;
int pic_dev_dd_step_prologue(const pic_dd_params* p, const pic_dd_prologue* a, void* stream);
/* Device-mode re-injection of the slots named in the absorption log (no flag scan), and the
 * device-mode thermostat (every active particle redraws u,v,w from sigma[species] with probability
 * gamma).  Philox is keyed by the ORIGINAL global index orig[i] + global_offset (orig == NULL: the
 * slot), and v0/w0 are addressed by it. */
int pic_dev_dd_reinject_philox_log(const pic_dd_params* p, const int32_t* dead_buf, int32_t dead_cap, double* x0,
                                   double* u0, double* v0, double* w0, int8_t* active, const int32_t* orig,
                                   const double sigma[2], uint64_t seed, uint64_t step, int64_t global_offset,
                                   void* stream);
/* pic_dev_dd_reinject_philox with the original-index payload (draws keyed by orig[i] + global_offset,
 * v0/w0 written at orig[i]) and a guard: when dead_count != NULL and *dead_count <= dead_cap the
 * kernel returns at entry, because pic_dev_dd_reinject_philox_log has visited every dead slot (and
 * vice versa: on overflow the log kernel does nothing and this scan re-injects). */
int pic_dev_dd_reinject_philox2(const pic_dd_params* p, double* x0, double* u0, double* v0, double* w0,
                                int8_t* active, const int32_t* orig, const double sigma[2], uint64_t seed,
                                uint64_t step, int64_t global_offset, const int32_t* dead_count, int32_t dead_cap,
                                void* stream);
int pic_dev_dd_thermostat_philox(const pic_dd_params* p, double* u0, double* v0, double* w0, const int8_t* active,
                                 const int32_t* orig, double gamma, const double sigma[2], uint64_t seed, uint64_t step,
                                 int64_t global_offset, void* stream);
/* Device mode (benchmark sizes, statistical parity only): Philox4x32-10 keyed by
 * (seed, step, global particle id); x~U(0,L), u,v,w~N(0,sqrt(kT/m)). v0/w0 may be NULL. */
int pic_dev_dd_reinject_philox(const pic_dd_params* p, double* x0, double* u0, double* v0, double* w0,
                               int8_t* active, const double sigma[2], uint64_t seed, uint64_t step,
                               int64_t global_offset, void* stream);
/* Commit helper: after the last iteration particles that were already inactive when it
 * started hold x1=u1=0 like the reference (PIC_L_DD.py:459-462); the commit itself is a
 * pointer swap on the host.  KE diagnostic sum(me*u^2/2) (PIC_L_DD.py:549): */
int pic_dev_sum_sq(const double* u, int64_t N, double scale, double* out1, void* stream);
/* out2 = {sum u, sum u*u} in one pass: np.std(u0) (PIC_L_DD.py:417) and KE (:549) together; per-CTA partial sums added
 * in CTA order (for a given particle order the result does not depend on scheduling).  The partial sums live in
 * one static device buffer: calls must be stream-ordered with respect to each other (one stream per process, as
 * all drivers of this library do). */
int pic_dev_moments(const double* u, int64_t N, double* out2, void* stream);
/* Counting sort by (species, cell) of the n-level state; out-of-place.  Keeps species
 * ranges contiguous; order inside a cell is unspecified (benchmark mode only).
 * counts: int32 scratch of 2*Ng+2 + ceil(2*Ng/1024)+2 entries (the tail is used by the
 * global-memory histogram path taken when 2*Ng counters do not fit shared memory).
 * v0/w0 (and outputs) may be NULL. */
int pic_dev_dd_sort_by_cell(const pic_dd_params* p, const double* x0, const double* u0,
                            const double* v0, const double* w0, double* x0s, double* u0s,
                            double* v0s, double* w0s, int32_t* counts, void* stream);
/* The same sort carrying the ORIGINAL INDEX of every particle as an int32 payload (orig == NULL on
 * the first sort: the identity), so that a sorted store still knows the reference's particle
 * numbering: re-injection draws in index order (PIC_L_DD.py:429-450), vionout, downloads. */
int pic_dev_dd_sort_by_cell2(const pic_dd_params* p, const double* x0, const double* u0, const int32_t* orig,
                             double* x0s, double* u0s, int32_t* origs, int32_t* counts, void* stream);
/* STABLE sort by (species, cell): LSD radix sort of each species block with 8-bit digits of the
 * cell index (per-tile digit histograms, exclusive scan, scatter with index-order ranks).  Equal
 * cells keep their previous order, so the result is independent of scheduling -- the sort of
 * the reproducible build (flags bit7).  The passes ping-pong between (x0,u0) and (xs,us):
 * *result_in_scratch (host int) = 1 when the sorted state ends in (xs,us), 0 when it is back in
 * (x0,u0).  scratch: int32[scratch_entries], at least 256*T + ceil(256*T/1024) + 2 entries with
 * T = ceil(max species block / 4096). */
int pic_dev_dd_sort_by_cell_stable(const pic_dd_params* p, double* x0, double* u0, double* xs, double* us,
                                   int32_t* scratch, int64_t scratch_entries, int32_t* result_in_scratch,
                                   void* stream);
/* The stable sort carrying the ORIGINAL INDEX of every particle (int32) through its passes, for a reproducible
 * run that keeps the reference's particle numbering (pic_dev_dd_sort_by_cell2's role in the reproducible build).
 * orig / origs ping-pong like (x0,u0) / (xs,us); identity != 0: orig's content is ignored and the numbering
 * starts as the slot index (first sort). */
int pic_dev_dd_sort_by_cell_stable2(const pic_dd_params* p, double* x0, double* u0, double* xs, double* us, int32_t* orig,
                                    int32_t* origs, int identity, int32_t* scratch, int64_t scratch_entries,
                                    int32_t* result_in_scratch, void* stream);
/* The same counting sort for any structure-of-arrays store: xs receives the sorted positions and
 * perm (int32[N]) the source slot of every output slot; apply it to the other arrays with
 * pic_dev_soa_permute.  Only p->N, n_split, Ng, dx are read. */
int pic_dev_sort_perm_by_cell(const pic_dd_params* p, const double* x, double* xs, int32_t* perm,
                              int32_t* counts, void* stream);
/* The counting sort carrying up to three fp64 payload arrays THROUGH the scatter (a; b, c may be NULL)
 * and, when perm != NULL, the source slot of every output slot for the remaining arrays
 * (pic_dev_soa_permute).  The pygcpic store sorts x with vx, vy, vz this way: the scatter writes runs,
 * where a gather by permutation reads 8 bytes per 32-byte sector. */
int pic_dev_sort_by_cell_payload(const pic_dd_params* p, const double* x, const double* a, const double* b, const double* c,
                                 double* xs, double* as, double* bs, double* cs, int32_t* perm, int32_t* counts,
                                 void* stream);
/* dst[t] = src[perm[t]] for up to 12 fp64, 2 int32 and 6 int8 arrays in ONE pass (arrays of device
 * pointers passed from the host). */
int pic_dev_soa_permute(const int32_t* perm, int64_t n, const double* const* src_f64,
                        double* const* dst_f64, int n_f64, const int32_t* const* src_i32,
                        int32_t* const* dst_i32, int n_i32, const int8_t* const* src_i8,
                        int8_t* const* dst_i8, int n_i8, void* stream);

/* ------------------------------------------------------------------------- *
 * pypic.py -- periodic implicit PIC
 * ------------------------------------------------------------------------- */
/* pypic.interpolate_p pypic.py:28-61 */
int pic_dev_pypic_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng,
                              double dx, int* range_err, void* stream);
/* pypic.weight_current_p :91-136 (v != NULL, wR=(x%dx)*idx) / weight_density_p :138-183
 * (v == NULL, wR=(x%dx)/dx).  p2c is the value AFTER int32 truncation. out zeroed by caller. */
int pic_dev_pypic_weight(const double* x, const double* q, const double* v, double* out, int64_t N,
                         int Ng, double dx, double p2c, int* range_err, void* stream);
/* The same deposit with ORDER-INDEPENDENT accumulation (fixed-point words, integer atomics; the initial rho of a
 * reproducible implicit_pic run): out += the deposit, one rounding per node.  qmax = max |q[i]| (sets the scale). */
int pic_dev_pypic_weight_fixed(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng,
                               double dx, double p2c, double qmax, int* range_err, void* stream);
typedef struct {
    int64_t N;
    int32_t Ng;
    int32_t flags;         /* 0 (default): TMA-staged private-window kernel over whole 16384-particle
                              chunks (fast on a store sorted by cell), grid-stride kernel on the tail;
                              bit0: plain shared-memory atomics per contribution; bit2: grid-stride kernel
                              with warp-uniform pre-reduction for every particle (any particle order);
                              bit1: x0 holds UNWRAPPED positions, ``x % L`` (pypic.py:277) is applied on load;
                              bit3: light iteration (no v1 store, no j1 deposit, see pic_dev_pypic_j1_repair);
                              bit4: large-grid build of the window kernel (per-warp field windows);
                              bit7: REPRODUCIBLE build (as pic_dd_params'): acc is fp64[2Ng] followed by the
                              fixed-point words int64[4Ng] = [hi(2Ng) | lo(2Ng)]; the window kernel and the j1
                              repair add to the words with integer atomics (order-independent),
                              pic_dev_pypic_field_update2 / pic_dev_pypic_j1_finish fold them back into acc */
    double dx, dt, L, p2c; /* p2c already truncated */
    double q, m;           /* single species (electrons)            */
} pic_pypic_params;
/* One Picard iteration of pypic.particle_push_p :259-279, fused: gather at xs with the
 * SMOOTHED field Fs, CN push, wrap, deposit jh(xh,vh) and j1(x1,v1).
 *   x1 holds the UNWRAPPED n+1 position of the previous iteration (xs=((x0+x1)*.5)%L);
 *   acc fp64[2*Ng] zero on entry. */
int pic_dev_pypic_picard_iter(const pic_pypic_params* p, const double* x0, const double* v0,
                              double* x1, double* v1, const double* Fs, double* acc, int first,
                              int* range_err, void* stream);
/* pypic.py:283-292: E1=E0+(dt/eps0)(mean(jh)-smooth(jh)); Eh; r=sum((Es-Eh)^2); Es=Eh;
 * Fs=smooth(Es) for the next gather; j1 copied out; acc zeroed.
 * stats fp64[4]: r, mean(j1), sum(eps0 E1^2 dx/2), iteration counter. */
int pic_dev_pypic_field_update(const pic_pypic_params* p, double* acc, const double* E0, double* Es,
                               double* Fs, double* E1, double* j1, double* stats, void* stream);
/* The same iteration with separate input / output buffers for the n+1 positions (ping-pong), so that
 * the inputs of the last iteration survive it. */
int pic_dev_pypic_picard_iter2(const pic_pypic_params* p, const double* x0, const double* v0,
                               const double* x1_in, double* x1_out, double* v1, const double* Fs,
                               double* acc, int first, int* range_err, void* stream);
/* Enqueue-ahead variants (see pic_dev_dd_picard_iter3 / pic_dev_dd_field_update2): `done` / `ctl`
 * is a device int32 read at kernel entry, a non-zero value makes the launch a no-op; the field kernel
 * sets it when `r > tol and k < maxiter` (pypic.py:259) fails, writes the smoothed field the
 * iteration gathered with to Fs_prev (what pic_dev_pypic_j1_repair needs) and the residual of the
 * k-th iteration to rhist[k-1].  Fs_prev, rhist, ctl may be NULL. */
int pic_dev_pypic_picard_iter3(const pic_pypic_params* p, const double* x0, const double* v0,
                               const double* x1_in, double* x1_out, double* v1, const double* Fs,
                               double* acc, int first, int* range_err, const int32_t* done, void* stream);
/* One Picard iteration with PER-PARTICLE charge and mass (particle_push_p takes q, m as arrays,
 * pypic.py:248: q_m = q / m; the deposit uses q[i]): grid-stride kernel, any particle order, every
 * iteration stores v1 and deposits j1.  p->q, p->m are ignored. */
int pic_dev_pypic_picard_iter_qm(const pic_pypic_params* p, const double* x0, const double* v0, const double* x1i,
                                 double* x1, double* v1, const double* q, const double* m, const double* Fs, double* acc,
                                 int first, int* range_err, const int32_t* done_flag, void* stream);
int pic_dev_pypic_field_update2(const pic_pypic_params* p, double* acc, const double* E0, double* Es,
                                double* Fs, double* E1, double* j1, double* stats, double* Fs_prev,
                                double* rhist, int32_t* ctl, double tol, int maxiter, void* stream);
/* Light iterations (flags bit3 of pic_pypic_params): the iteration neither stores v1 nor deposits j1 --
 * both are only used after the Picard loop.  If the loop ends on one, pic_dev_pypic_j1_repair recomputes
 * v1 = v0 + dt*(q/m)*E(xs) from that iteration's inputs (x1_prev, the smoothed field Fs_prev it
 * gathered from) and deposits weight_current_p(x1 % L, q, v1) (pypic.py:277-279) into acc[Ng..2Ng);
 * after the caller's all-reduce when sharded, pic_dev_pypic_j1_finish writes j1[Ng], stats[1] =
 * mean(j1) and zeroes the accumulator. */
int pic_dev_pypic_j1_repair(const pic_pypic_params* p, const double* x0, const double* v0,
                            const double* x1_prev, const double* x1_last, const double* Fs_prev, double* v1,
                            int first, double* acc, int* range_err, void* stream);
int pic_dev_pypic_j1_finish(const pic_pypic_params* p, double* acc, double* j1, double* stats, void* stream);
/* x[i] = x[i] % L in place (pypic.py:277 applied to the committed positions) */
int pic_dev_wrap_periodic(double* x, int64_t N, double L, void* stream);

/* ------------------------------------------------------------------------- *
 * PIC_L.py -- periodic explicit leapfrog, Poisson every step
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t N;
    int64_t n_split;
    int32_t Ng;    /* cells; the grid has Ng+1 nodes (PIC_L.py:646,667) */
    int32_t flags;
    double dx, dt, L, p2c; /* wrap length is L+dx (PIC_L.py:285) */
    double q[2], m[2];
} pic_l_params;
/* PIC_L.interpolateFieldPeriodic :39-46 */
int pic_dev_l_interpolate(const double* F, const double* x, double* out, int64_t N, int Ng, double dx,
                          int* range_err, void* stream);
/* PIC_L.weightDensitiesPeriodic :100-118 (v==NULL) / weightCurrentsPeriodic :62-80; folds included */
int pic_dev_l_weight(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng,
                     double dx, double p2c, int* range_err, void* stream);
/* PIC_L.weightDensities :83-98 (v==NULL) / weightCurrents :48-60: bounded CIC on Ng nodes, no folds */
int pic_dev_l_weight_bounded(const double* x, const double* q, const double* v, double* out, int64_t N, int Ng,
                             double dx, double p2c, int* range_err, void* stream);
/* PIC_L.pushParticlesImplicit :261-270 (function form of the Crank-Nicolson push; per-particle q, m) */
int pic_dev_l_push_implicit(const double* x0, const double* xh, const double* v, const double* q, const double* m,
                            const double* Eh, double* xout, double* vout, int64_t N, int Ng, double dx, double dt,
                            int* range_err, void* stream);
/* PIC_L.applyBoundaryConditions :272-282, device half: flags[i] = 0 where x > L or x <= 0, else 1 (the host
 * then redraws those particles from the legacy stream in index order and scatters them back) */
int pic_dev_l_outside_flags(const double* x, int8_t* flags, int64_t N, double L, void* stream);
/* Fused explicit step, particle phase (PIC_L.py:767-768 + next step's :763):
 * gather E at x, kick-drift-kick, wrap x%(L+dx), and deposit rho of the NEW positions
 * into rho_acc fp64[Ng+1] (zero on entry, raw CIC; fold applied by pic_dev_l_field_solve).
 * flags 0 (default): TMA-staged private-window kernel over whole 16384-particle chunks (fast on
 * a store sorted by (species, cell)), grid-stride kernel on the tail; bit0: plain shared-memory
 * atomics; bit2: grid-stride kernel with warp-uniform pre-reduction for every particle (any
 * order); bit1: function-level PIC_L.pushParticlesExplicit :248-259 -- x,v receive the UNWRAPPED
 * xout,vout and nothing is deposited.
 * bit7: REPRODUCIBLE build (as pic_dd_params'): rho_acc is fp64[Ng+1] followed by int64[2*(Ng+1)] fixed-point words
 * [hi | lo]; every global addition of the window kernel is an integer addition on the words (order-independent),
 * which pic_dev_l_field_solve folds back into rho_acc.  Together with the stable sort
 * (pic_dev_dd_sort_by_cell_stable[2]) two runs are bit-identical.  pic_dev_l_deposit_fixed: the initial rho of
 * the current positions (PIC_L.py:763 before the first step) on the words. */
int pic_dev_l_deposit_fixed(const pic_l_params* p, const double* x, double* rho_acc, int* range_err, void* stream);
int pic_dev_l_push_deposit(const pic_l_params* p, double* x, double* v, const double* E,
                           double* rho_acc, int* range_err, void* stream);
/* PIC_L.py:763-766 field phase: fold rho_acc -> rho; periodic Poisson; -max; E=-dphi/dx.
 * work: 13*(Ng+1) doubles.  rho_acc is zeroed.  stats[0] = sum(eps0*E*E/2). */
int pic_dev_l_field_solve(const pic_l_params* p, double* rho_acc, double* rho, double* phi, double* E,
                          double* work, double* stats, void* stream);

/* ------------------------------------------------------------------------- *
 * pygcpic.py -- Particle / Grid (Boris 1D3V, guiding centre, Boltzmann electrons)
 * Particle store: SoA r0..r6 (x,y,z,vx,vy,vz,t), charge_state/m/p2c fp64, flags int8.
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t N;
    int32_t ng;     /* grid nodes */
    int32_t flags;  /* bit0: the Egrid argument of push_boris / push_rk4 holds one already-gathered
                       E_x per particle (function-level Particle.push_6D / push_GC); bit1: push_boris
                       applies no boundary test */
    double dx, dt, length;
    double B[3];
    double Eyz[2];  /* the shared E[1],E[2] of Particle.E (only E[0] is ever gathered) */
} pic_gc_params;
/* Particle.interpolate_electric_field_dirichlet pygcpic.py:325-348 (MIRRORED weights) */
int pic_dev_gc_interpolate(const double* E, const double* x, double* out, int64_t N, int ng, double dx,
                           int* range_err, void* stream);
/* Grid.weight_particles_to_grid_boltzmann :868-883: rho and n of active particles (zeroed by caller) */
int pic_dev_gc_weight(const double* x, const double* charge_state, const double* p2c,
                      const int8_t* active, double* rho, double* n, int64_t N, int ng, double dx,
                      int* range_err, void* stream);
/* Fused particle phase of a pygcpic step for active particles (pygcpic.py:1500-1502):
 * mirrored gather -> Particle.push_6D :460-507 -> apply_BCs_dirichlet :668-689.
 * r: 7 device arrays; hit_count: device int64 incremented per wall hit this call;
 * hit_flag (may be NULL): int8[N], 1 for particles absorbed by this call (tallies T1). */
int pic_dev_gc_push_boris(const pic_gc_params* p, double* const r[7], const double* charge_state,
                          const double* m, int8_t* active, int8_t* at_wall, int8_t* hit_flag,
                          const double* Egrid, long long* hit_count, int* range_err, void* stream);
/* The same step for a SPECIES-UNIFORM store (every particle has this charge_state, m, p2c), on the
 * TMA-ring / private-window design, optionally fused with the CIC deposit of the number density
 * at the NEW positions of the particles still active (pygcpic.py:871-883, the next step's D5):
 * n_acc fp64[ng] is accumulated (not zeroed) when non-NULL.  hit_flag is only written for
 * particles that are inactive or absorbed by this call (it must start zeroed and be cleared
 * when a slot is re-activated).  The arrays must be 16-byte aligned. */
int pic_dev_gc_push_boris_uniform(const pic_gc_params* p, double* const r[7], double charge_state, double m,
                                  double p2c, int8_t* active, int8_t* at_wall, int8_t* hit_flag,
                                  const double* Egrid, double* n_acc, long long* hit_count, int* range_err,
                                  void* stream);
/* The same with the LEAN store mode: lean != 0 streams x, vx, vy, vz only (64 B per particle-step, the
 * Boris row of SURVEY.md 8(d)); y, z and the per-particle clock r[6] are not advanced (nothing on the
 * path reads them; r[1], r[2] may be NULL) and a particle absorbed by this push gets r[6] = t_now, so
 * the clock of every particle is still known: t_now for the active ones, the time of death otherwise. */
int pic_dev_gc_push_boris_uniform2(const pic_gc_params* p, double* const r[7], double charge_state, double m, double p2c,
                                   int lean, double t_now, int8_t* active, int8_t* at_wall, int8_t* hit_flag,
                                   const double* Egrid, double* n_acc, long long* hit_count, int* range_err,
                                   void* stream);
/* Post-push pass of pic_bca_aps' particle loop (pygcpic.py:1509-1541), N3: per particle the
 * ionisation eligibility (Z==1 & charge 0: attempt_first_ionization :350-395; Z==5 & charge<3:
 * attempt_nth_ionization :397-458) and probability density^2*rate*dx*dt/p2c (density = CIC gather
 * of n_grid; rate[4] = np.interp'd coefficients for (Z,charge) = (1,0),(5,0),(5,1),(5,2)); the
 * mid-domain exit of wall-born particles :1530-1541 (active <- 0, flag returned); and the
 * particle's deterministic contribution to the running source-ion count of :1544.  The uniform
 * draws and the decisions are made on the host in index order (legacy stream parity). */
/* The fused step for a MIXED store (ions, neutrals, several charge states in one list, as in the
 * reference's loop pygcpic.py:1498-1549): per-particle charge_state, m, p2c; results bit-identical to
 * pic_dev_gc_push_boris; with n_acc != NULL both deposits of Grid.weight_particles_to_grid_boltzmann
 * :871-883 are fused (n_acc += p2c/dx*w, rho_acc += charge_state*e*p2c/dx*w at the new positions of the
 * particles still active).  lean as in pic_dev_gc_push_boris_uniform2. */
int pic_dev_gc_push_boris_mixed(const pic_gc_params* p, double* const r[7], const double* charge_state, const double* m,
                                const double* p2c, int lean, double t_now, int8_t* active, int8_t* at_wall,
                                int8_t* hit_flag, const double* Egrid, double* n_acc, double* rho_acc,
                                long long* hit_count, int* range_err, void* stream);
int pic_dev_gc_post_push(const double* x, const double* p2c, const double* charge_state, const int32_t* Z,
                         const int8_t* from_wall, int8_t* active, const double* n_grid, int ng, double dx,
                         double dt, double length, const double rate[4], int source_Z, double* prob,
                         int8_t* eligible, int8_t* midexit, int8_t* contrib, int64_t N, int* range_err,
                         void* stream);
/* n = n_acc ; rho = charge_state*e*n_acc for a species-uniform store */
int pic_dev_gc_uniform_finish(const double* n_acc, double* n, double* rho, int ng, double charge_state,
                              void* stream);
/* adds the CIC number density of the slots idx[0..M) (int64 indices; re-activated particles) */
int pic_dev_gc_deposit_idx(const double* x, const int64_t* idx, int64_t M, double p2c, double dx, int ng,
                           double* n_acc, int* range_err, void* stream);
/* Particle.apply_BCs_dirichlet :668-689 alone */
int pic_dev_gc_apply_bcs(const double* x, int8_t* active, int8_t* at_wall, int64_t N, double length,
                         void* stream);
/* Particle.transform_6D_to_GC :509-551 / transform_GC_to_6D :553-596 (a = uniform draws, N x 3
 * as three arrays) / push_GC :598-645 (RK4, E gathered once with mirrored weights if Egrid!=NULL,
 * else Ex=0) -- applied to particles with active==1 */
int pic_dev_gc_to_gc(const pic_gc_params* p, double* const r[7], const double* charge_state,
                     const double* m, const int8_t* active, void* stream);
int pic_dev_gc_to_6d(const pic_gc_params* p, double* const r[7], const double* charge_state,
                     const double* m, const int8_t* active, const double* a0, const double* a1,
                     const double* a2, void* stream);
int pic_dev_gc_push_rk4(const pic_gc_params* p, double* const r[7], const double* charge_state,
                        const double* m, const int8_t* active, const double* Egrid, int* range_err,
                        void* stream);
/* The same step for a species-uniform store (scalars instead of the charge_state / m arrays):
 * the ExB drift terms are computed once per particle and the divisions by uniform quantities
 * use a correctly rounded constant-divisor division (bit-identical results). */
int pic_dev_gc_push_rk4_uniform(const pic_gc_params* p, double* const r[7], double charge_state, double m,
                                const int8_t* active, const double* Egrid, int* range_err, void* stream);
/* Boltzmann reference-density update, Grid.weight_particles_to_grid_boltzmann :889-904.
 * domain = Grid.domain (np.linspace(0,length,ng)) so np.trapz's spacings are reproduced;
 * state fp64[3] = {n0, p_old, initialised(0/1)} on the device. */
int pic_dev_gc_n0_update(const double* phi, const double* n, const double* domain, int ng, double Te,
                         double ve, double added_particles, double dt, double* state, void* stream);
/* Reactivate-or-delete rule pygcpic.py:1543-1549 as an exclusive prefix scan:
 * contrib_entry/contrib_after int8 (is an active source ion before / after its own push),
 * inactive_entry int8.  Outputs decision int8: 0 none, 1 reactivate, 2 delete.
 * idx_scratch/base_scratch: int32[N]; scratch: int64[4 + 2*(ceil(N/2048)+1)], on return
 * scratch[0]=#inactive at entry, [1]=#source ions at entry, [2]=#reactivated, [3]=#deleted. */
int pic_dev_gc_decide(const int8_t* inactive_entry, const int8_t* contrib_entry, const int8_t* contrib_after,
                      int8_t* decision, int64_t N, int64_t source_N, int32_t* idx_scratch,
                      int32_t* base_scratch, int64_t* scratch, void* stream);

/* ------------------------------------------------------------------------- *
 * Stream compaction (warp-ballot prefix sum, stable = index order preserved)
 * ------------------------------------------------------------------------- */
/* idx_out receives the indices i (ascending) with flags[i] != keep_value... see mode:
 * mode 0: select flags[i] != 1 (inactive slots of the sheath), mode 1: select flags[i]==0,
 * mode 2: select flags[i] != 2 (survivors of the pygcpic deletion).
 * count_out: device int64[1]. block_counts: int64 scratch of 2*(ceil(N/2048)+1) entries. */
int pic_dev_compact_flags(const int8_t* flags, int64_t N, int mode, int32_t* idx_out, int64_t* count_out,
                          int64_t* block_counts, void* stream);
/* dst[k] = src[idx[k]] for k<n (fp64 / int8 payloads) */
int pic_dev_gather_f64(const double* src, const int32_t* idx, double* dst, int64_t n, void* stream);
int pic_dev_gather_i8(const int8_t* src, const int32_t* idx, int8_t* dst, int64_t n, void* stream);
int pic_dev_gather_i32(const int32_t* src, const int32_t* idx, int32_t* dst, int64_t n, void* stream);
/* dst[idx[t]] = src[t] (a sorted store back to the original order; idx == NULL: copy); inv[perm[t]] = t */
int pic_dev_scatter_f64(const double* src, const int32_t* idx, double* dst, int64_t n, void* stream);
int pic_dev_scatter_i8(const int8_t* src, const int32_t* idx, int8_t* dst, int64_t n, void* stream);
int pic_dev_invert_perm(const int32_t* perm, int32_t* inv, int64_t n, void* stream);

/* ------------------------------------------------------------------------- *
 * Device-side initialisers and IEAD histogram (N1)
 * ------------------------------------------------------------------------- */
/* x ~ U(xlo,xhi); v0 ~ N(mean[s],sigma[s]); v1,v2 ~ N(0,sigma[s]); s = (i >= n_split).  Philox4x32-10
 * keyed by (seed, stream_id, global_offset+i): independent of the sharding.  NULL arrays are
 * skipped.  Distribution parity with PIC_L_DD.initialize :223-314 / Particle._initialize_6D
 * pygcpic.py:277-304 (the host initialisers of the drop-in modules keep MT19937 stream parity). */
int pic_dev_init_uniform_maxwellian(double* x, double* v0, double* v1, double* v2, int64_t N, int64_t n_split,
                                    double xlo, double xhi, const double sigma[2], const double mean[2],
                                    uint64_t seed, uint64_t stream_id, int64_t global_offset, void* stream);
/* pypic.initialize_p :457-467 perturbation loader: global particle g < prefix[Ng] is placed uniformly
 * in the cell c with prefix[c] <= g < prefix[c+1] (prefix = exclusive cumsum of int(F[i]), Ng+1
 * int64 on the device; X = Ng+1 cell edges). */
int pic_dev_pypic_perturb_positions(double* x, int64_t N, const int64_t* prefix, const double* X, int Ng,
                                    uint64_t seed, int64_t global_offset, void* stream);
/* hist[n_e_bins*n_a_bins] (fp64 counts, accumulated) += numpy.histogram2d(kinetic_energy/e, angle) of the
 * particles with select[i]==1 (and Z[i]==Z_select when Z != NULL): pygcpic.py:1516-1527, 1574-1584. */
int pic_dev_gc_iead_hist(const double* vx, const double* vy, const double* vz, const double* m,
                         const int8_t* select, const int32_t* Z, int Z_select, int64_t N,
                         const double* e_edges, int n_e_bins, const double* a_edges, int n_a_bins,
                         double* hist, void* stream);

/* ------------------------------------------------------------------------- *
 * Host-buffer entry points: the calls a binding of the reference would make with
 * NumPy arrays (same argument meaning as the Python functions they replace).
 * ------------------------------------------------------------------------- */
int pic_host_pypic_interpolate_p(const double* F, const double* x, int Ng, int64_t N, double dx, double* out);
int pic_host_pypic_weight_current_p(const double* x, const double* q, const double* v, int p2c, int Ng,
                                    int64_t N, double dx, double* j);
int pic_host_pypic_weight_density_p(const double* x, const double* q, int p2c, int Ng, int64_t N,
                                    double dx, double* rho);
int pic_host_dd_interpolateField(const double* F, const double* x, int Ng, int64_t N, double dx, double* out);
int pic_host_dd_weightCurrents(const double* x, const double* q, const double* v, double p2c, int Ng,
                               int64_t N, double dx, double dt, const double* active, double* j);
int pic_host_dd_weightDensities(const double* x, const double* q, double p2c, int Ng, int64_t N,
                                double dx, const double* active, double* rho);
/* Whole sheath timestep with HOST buffers (bench.py's e2e leg): uploads x0,u0,E0 (all particles
 * active, i.e. after re-injection), runs the Picard loop of PIC_L_DD.py:452-545 and downloads
 * x1,u1,active,E1,j1.  Returns the iteration count in *iters and the residual in *resid. */
int pic_host_dd_step(const pic_dd_params* p, const double* x0, const double* u0, const double* E0,
                     double tol, int maxiter, double* x1, double* u1, int8_t* active, double* E1,
                     double* j1, int* iters, double* resid);
/* The same for `nbatch` INDEPENDENT states (arrays of nbatch host pointers; pinned memory makes
 * the copies asynchronous): batches are pipelined through two device slots so that the upload
 * of batch b+1 and the download of batch b-1 overlap the Picard loop of batch b.  iters/resid
 * receive nbatch entries. */
int pic_host_dd_step_batches(const pic_dd_params* p, int nbatch, const double* const* x0,
                             const double* const* u0, const double* const* E0, double tol, int maxiter,
                             double* const* x1, double* const* u1, int8_t* const* active,
                             double* const* E1, double* const* j1, int* iters, double* resid);

#ifdef __cplusplus
}
#endif
#endif /* PIC_B200_H */
