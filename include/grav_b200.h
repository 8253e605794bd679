/*
 * grav_b200.h -- thin C ABI of the B200 (sm_100a) acceleration path for grav_sim.
 *
 * This is the whole device boundary: plain pointers, sizes and scalars; no CUDA, torch or
 * reference types in any signature.  Every entry returns 0 on success or one of the
 * GRAV_B200_E* codes below; grav_b200_last_error() then returns a thread-local message
 * (CUDA / NCCL error string included).  There is NO CPU fallback: if no sm_100-class device
 * or no CUDA driver is present every compute entry fails with GRAV_B200_ENODEV.
 *
 * Two families of entry points:
 *
 *  (1) host-pointer one-shots.  These are what the reference's C dispatch binds to
 *      (grav_sim_shim.c forwards `acceleration()` & co. here).  Inputs are the reference's
 *      host AoS arrays (x[3N], m[N]); the result is written to the caller's host a[3N].
 *      Each cites the reference function it replaces.
 *
 *  (2) a context (grav_b200_ctx) that keeps particle state resident in HBM as packed
 *      (x,y,z,m) records + AoS velocity/acceleration, so integrator sub-steps do not
 *      round-trip to the host, and that shards work over ranks (one process per GPU,
 *      NCCL all-gather of positions per force call).
 *
 * Reference citations are relative to the upstream tree (alvinng4/Gravity-Simulator).
 */
#ifndef GRAV_B200_H
#define GRAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes ------------------------------------------------------------------ */
#define GRAV_B200_OK        0
#define GRAV_B200_EINVAL    1   /* bad argument (maps to GRAV_VALUE_ERROR / GRAV_POINTER_ERROR) */
#define GRAV_B200_ENOMEM    2   /* cudaMalloc / malloc failed (maps to GRAV_MEMORY_ERROR)        */
#define GRAV_B200_ECUDA     3   /* any other CUDA runtime failure (maps to GRAV_FAILURE)         */
#define GRAV_B200_ENODEV    4   /* no usable device / driver (maps to GRAV_FAILURE)              */
#define GRAV_B200_ENCCL     5   /* NCCL failure (maps to GRAV_FAILURE)                           */
#define GRAV_B200_ETREE     6   /* the octree outgrew its device buffers (pathologically deep chains).  Builds are queued
                                 * without host synchronisation, so this surfaces at the next synchronising call; the
                                 * host-pointer entries retry with larger buffers by themselves (maps to GRAV_FAILURE)  */

/* acceleration methods: same encoding as src/acceleration.h:16-18 */
#define GRAV_B200_METHOD_PAIRWISE   1
#define GRAV_B200_METHOD_MASSLESS   2
#define GRAV_B200_METHOD_BARNES_HUT 3

/* Largest supported particle count (2^24, the largest size in BASELINE.json): work-unit indices are 32-bit and
 * the walk records pack the particle count in 26 bits.  Larger systems are rejected with GRAV_B200_EINVAL. */
#define GRAV_B200_MAX_PARTICLES (1 << 24)

/* Barnes-Hut walk semantics (see DESIGN.md "BH modes") */
#define GRAV_B200_BH_REFERENCE 0  /* bug-for-bug with src/acceleration_barnes_hut.c:78-248 (default) */
#define GRAV_B200_BH_FIXED     1  /* opt-in: correct inclusion test, leaves never dropped            */

const char *grav_b200_last_error(void);

/* Number of usable CUDA devices (0 if none / no driver).  Never fails. */
int grav_b200_device_count(void);

/* ---- (1) host-pointer one-shots ---------------------------------------------------- */

/* replaces acceleration_pairwise, src/acceleration.c:177-234
 * a[3n] (out, fully overwritten), x[3n], m[n] host arrays; Plummer softening eps>=0. */
int grav_b200_acceleration_pairwise(double *a, int n, const double *x, const double *m,
                                    double G, double softening_length);

/* replaces acceleration_massless, src/acceleration.c:236-367 (incl. its m[rank] quirk, :357-359) */
int grav_b200_acceleration_massless(double *a, int n, const double *x, const double *m,
                                    double G, double softening_length);

/* replaces acceleration_barnes_hut, src/acceleration_barnes_hut.c:33-76
 * (auto bounding box, tree built and dropped inside the call, like the reference) */
int grav_b200_acceleration_barnes_hut(double *a, int n, const double *x, const double *m,
                                      double G, double softening_length,
                                      double opening_angle, int max_num_particles_per_leaf);

/* replaces the O(N^2) pair loop of compute_energy, src/utils.c:27-59 (and of compute_energy_python,
 * src/python_interface.c:195-243, per snapshot): sum_i m_i |v_i|^2 / 2 - G sum_{i<j} m_i m_j / |x_i - x_j|,
 * unsoftened.  Summation order differs from the serial loop: agrees to ~1e-13 relative. */
int grav_b200_compute_energy(double *energy, int n, const double *x, const double *v, const double *m, double G);

/* replaces whfast_acceleration_pairwise / _massless, src/integrator_whfast.c:839-957, 959-1264
 * jacobi_x[3n], eta[n] as computed by the WHFast caller.  Softening is r^3 + eps^3 here. */
int grav_b200_whfast_acceleration_pairwise(double *a, int n, const double *x, const double *m,
                                           double G, const double *jacobi_x, const double *eta,
                                           double softening_length);
int grav_b200_whfast_acceleration_massless(double *a, int n, const double *x, const double *m,
                                           double G, const double *jacobi_x, const double *eta,
                                           double softening_length);

/* replaces construct_octree, src/linear_octree.c:825-962.
 * Builds the linear octree on the device and copies it out.  box_center may be NULL and
 * box_width <= 0 for the automatic bounding box (src/linear_octree.c:856-867).
 * The caller passes the addresses of ten NULL pointers; on success each is a malloc()ed host
 * array the caller frees with free(): keys[n] (sorted), sorted_indices[n] and eight
 * per-node arrays of *num_nodes entries (layout of LinearOctree, src/linear_octree.h:20-57).
 * first_internal_children_idx is -1 for leaves (the reference leaves it uninitialised). */
int grav_b200_construct_octree(int n, const double *x, const double *m,
                               int max_num_particles_per_leaf,
                               const double *box_center, double box_width,
                               double *out_box_width, int *out_num_nodes,
                               int64_t **keys, int **sorted_indices,
                               int **tree_num_particles, int **tree_num_internal_children,
                               int **tree_first_particle_sorted_idx,
                               int **tree_first_internal_children_idx,
                               double **tree_mass, double **tree_com_x,
                               double **tree_com_y, double **tree_com_z);

/* Morton keys only (stage BH-2, src/linear_octree.c:159-202), for stage-level parity tests.
 * keys_unsorted[n] in particle order.  out_center[3], out_width may be NULL. */
int grav_b200_morton_keys(int n, const double *x, int64_t *keys_unsorted,
                          double *out_center, double *out_width);

/* BH walk mode for the one-shot entries and new contexts; also settable with the
 * environment variable GRAV_B200_BH_MODE=reference|fixed (read once). */
int grav_b200_set_bh_mode(int mode);
int grav_b200_get_bh_mode(void);

/* Barnes-Hut walk arithmetic.  Every accept / open / leaf decision is always taken with the reference's IEEE operations
 * in the reference's order, so the set of sources per particle is the reference's in both settings.
 *   0 (default)  warp-cooperative walk; accepted sources are evaluated with the direct sum's fused arithmetic and summed
 *                in a different order: accelerations agree with src/acceleration_barnes_hut.c:78-248 to <= 1e-12 relative
 *                (the north-star tolerance; observed ~1e-15).
 *   1            per-lane walk in the reference's depth-first order with sqrt / div / separate multiplies: accelerations
 *                BIT-IDENTICAL to the x86-64 reference build, about 2-3x slower.
 * Also settable with the environment variable GRAV_B200_BH_EXACT=1 (read once).  Applies to the one-shot entries and to
 * contexts created afterwards. */
int grav_b200_set_bh_exact(int on);
int grav_b200_get_bh_exact(void);

/* How the pairwise direct sum (src/acceleration.c:177-234) is evaluated.
 *   -1 (default)  automatic: systems large enough to balance evaluate every unordered pair ONCE and apply it to both
 *                 particles, like the reference's i < j loop (10 FP64 instructions per ordered interaction, 9 with equal
 *                 masses); smaller ones evaluate every ordered interaction (16 instructions, finer work units)
 *    0            ordered interactions always
 *    1            pair-once whenever the system has at least 256 particles
 * Both agree with the reference to <= 1e-12 relative.  Also settable with GRAV_B200_DS_SYM (read once). */
int grav_b200_set_direct_sum_mode(int mode);
int grav_b200_get_direct_sum_mode(void);
/* Test hooks (host code, no GPU needed): the work decomposition of the pair-once kernel (direct_sum_sym.cu).
 * _segments: the (row, [group_begin, group_end)) pieces CTA `cta` of `ctas_total` works through for a system of n particles
 * (rows of _row() particles, groups of 32; returns how many, -1 if the system has fewer than two rows); _touches: whether
 * the finishing kernel reads that CTA's private sums for block `block` of _row() particles. */
int grav_b200_debug_pair_once_segments(int n, int ctas_total, int cta, int max_segments, int *rows, int *group_begin, int *group_end);
int grav_b200_debug_pair_once_touches(int n, int ctas_total, int cta, int block);
int grav_b200_debug_pair_once_row(void);

/* ---- (2) device-resident context ---------------------------------------------------- */

typedef struct grav_b200_ctx grav_b200_ctx;

/* Create a context on `device`.  world_size==1: nccl_unique_id may be NULL.
 * world_size>1: one process per GPU; every rank passes the same 128-byte id obtained from
 * grav_b200_nccl_unique_id() on rank 0 (exchange it with any host mechanism). */
int grav_b200_ctx_create(grav_b200_ctx **out, int device, int rank, int world_size,
                         const void *nccl_unique_id);
void grav_b200_ctx_destroy(grav_b200_ctx *ctx);
int grav_b200_nccl_unique_id(void *out128);

/* Device team: num_devices GPUs of this box driven from the ONE calling thread (the reference's callers,
 * src/integrator.c:963,1021 and grav_sim/simulator.py:66-102, are single-process and single-threaded, so the
 * one-process-per-GPU form above is out of their reach).  Returns the LEADER context (rank 0 on devices[0]; devices NULL =
 * 0..num_devices-1); k-1 worker threads own the contexts of the other devices (ranks 1..k-1 of an in-process NCCL
 * communicator).  Every grav_b200_ctx_* entry called on the leader runs on all members -- same sharding as with one
 * process per GPU (direct sum by target range + all-gather of positions; Barnes-Hut replicated build + sharded walk) --
 * and returns when all have returned; uploads read the caller's host arrays from every member, downloads write each
 * member's own slice.  Not available for the resident WHFast (one GPU by design).  Destroy the leader to end the team.
 * ctx_create_auto(): device GRAV_B200_DEVICE (default 0), or a team of GRAV_B200_DEVICES devices starting there -- what
 * the host-pointer one-shots (the drop-in acceleration()) and the resident time loops of the drop-in build use. */
int grav_b200_ctx_create_team(grav_b200_ctx **out, int num_devices, const int *devices);
int grav_b200_ctx_create_auto(grav_b200_ctx **out);
int grav_b200_ctx_team_size(const grav_b200_ctx *ctx);

/* Upload the full system (host AoS, as in struct System, src/system.h:12-20).  v may be NULL.
 * With world_size>1 every rank passes the same arrays; rank r owns targets
 * [r*n/world, (r+1)*n/world) and keeps only their v / a up to date. */
int grav_b200_ctx_set_system(grav_b200_ctx *ctx, int n, const double *x, const double *v,
                             const double *m, double G);
int grav_b200_ctx_set_positions(grav_b200_ctx *ctx, const double *x);
int grav_b200_ctx_num_particles(const grav_b200_ctx *ctx);
void grav_b200_ctx_owned_range(const grav_b200_ctx *ctx, int *lo, int *hi);

/* One force evaluation on the resident state; result stays in HBM.
 * With world_size>1 it starts with the all-gather of the owned position shards. */
int grav_b200_ctx_acceleration(grav_b200_ctx *ctx, int method, double softening_length,
                               double opening_angle, int max_num_particles_per_leaf);

/* Download (host AoS).  world_size>1: all ranks' shards are gathered first, so the call is
 * collective and every rank receives the full array. */
int grav_b200_ctx_get_positions(grav_b200_ctx *ctx, double *x);
int grav_b200_ctx_get_velocities(grav_b200_ctx *ctx, double *v);
int grav_b200_ctx_get_accelerations(grav_b200_ctx *ctx, double *a);

/* Leapfrog (kick-drift-kick with compensated summation) on the resident state; mirrors the loop of
 * src/integrator.c:894-1121.  leapfrog_begin() zeroes the error terms, evaluates a(x0) and does the initial half
 * kick with step dt (:963-982), leaving v at the half step.  leapfrog_steps() advances num_steps steps of size dt:
 * drift (:1009-1018), force (:1021), full kick (:1030-1039).  While a leapfrog is running, get_velocities() and
 * energy() return velocities brought back to the position time level (v - a dt/2, the reference's snapshot
 * convention, :1048-1056) without disturbing the state; leapfrog_end() applies that shift in place (:1088-1094). */
int grav_b200_ctx_leapfrog_begin(grav_b200_ctx *ctx, int method, double softening_length,
                                 double opening_angle, int max_num_particles_per_leaf, double dt);
int grav_b200_ctx_leapfrog_steps(grav_b200_ctx *ctx, double dt, int64_t num_steps);
int grav_b200_ctx_leapfrog_end(grav_b200_ctx *ctx);

/* Euler, Euler-Cromer and RK4 on the resident state; mirror euler(), euler_cromer() and rk4(), src/integrator.c:281-454,
 * :456-628, :630-892 (same update formulas, compensated summation and operation order).  fixed_begin() zeroes the
 * error terms and fixes the force parameters; fixed_steps() advances num_steps steps of size dt (one force evaluation
 * per step, four for RK4).  Positions/velocities are read back with get_positions()/get_velocities(). */
#define GRAV_B200_INTEGRATOR_EULER        1   /* src/integrator.h:17-19 */
#define GRAV_B200_INTEGRATOR_EULER_CROMER 2
#define GRAV_B200_INTEGRATOR_RK4          3
int grav_b200_ctx_fixed_begin(grav_b200_ctx *ctx, int integrator, int method, double softening_length,
                              double opening_angle, int max_num_particles_per_leaf);
int grav_b200_ctx_fixed_steps(grav_b200_ctx *ctx, double dt, int64_t num_steps);

/* Device-resident WHFast on the resident state (one GPU); mirrors whfast(), src/integrator_whfast.c:200-407, with
 * the particle order, eta, the Kepler drift, both coordinate transforms and the kick on the device -- bit-identical
 * to the reference after any number of steps.
 *   whfast_begin(): particle_ids[n] host array or NULL (= 0..n-1); sorts by distance from id 0 (:242), eta and
 *                   cartesian_to_jacobi (:266-267), the first interaction acceleration and the half kick (:268-273).
 *                   method: pairwise or massless (:817-837).  remove_invalid_particles as IntegratorParam's flag.
 *   whfast_steps(): num_steps times { sort by Jacobi distance (:301-311), eta (:312), Kepler drift with the optional
 *                   removal of unsolvable particles (:315-327), jacobi_to_cartesian (:330), acceleration (:333),
 *                   kick (:340) }.  The particle count can shrink (ctx_num_particles()).
 *   whfast_get_state(): downloads n, ids, x, v, m (any may be NULL) in the current particle order.  snapshot != 0
 *                   first brings the velocities back by half a step, as the reference does before an output
 *                   (:346-351) -- and like the reference leaves x/v in that state; snapshot == 0 returns what the
 *                   last step left (v from the half-step Jacobi velocities), the reference's state at loop exit.
 *   whfast_end():   releases the integrator state.
 * Ordering assumptions behind "bit-identical": equal distances keep index order (stable sort -- glibc's qsort is a stable
 * merge sort for these sizes; the C standard leaves tie order open) and invalid particles are removed in ascending index
 * order (the reference's serial order; its OpenMP build fills the removal list in thread-arrival order).
 *   whfast_set_verbose(): the caller's Settings.verbose.  At GRAV_VERBOSITY_VERBOSE (3) a removal prints the reference's
 *                   "whfast_drift: Removing N invalid particles. Particle IDs: [...]" line (:609-623) to stderr.  The
 *                   per-particle "Kepler's equation did not converge" lines and the Stumpff NaN/Inf warning (:486-549)
 *                   are not printed by this loop (the iteration's error value stays on the device). */
int grav_b200_ctx_whfast_begin(grav_b200_ctx *ctx, const int *particle_ids, int method, double softening_length,
                               double dt, int remove_invalid_particles);
int grav_b200_ctx_whfast_set_verbose(grav_b200_ctx *ctx, int level);
int grav_b200_ctx_whfast_steps(grav_b200_ctx *ctx, double dt, int64_t num_steps);
int grav_b200_ctx_whfast_get_state(grav_b200_ctx *ctx, int snapshot, int *n_out, int *particle_ids, double *x,
                                   double *v, double *m);
int grav_b200_ctx_whfast_end(grav_b200_ctx *ctx);

/* Total energy of the resident state, same definition as compute_energy, src/utils.c:27-59
 * (unsoftened potential).  Collective when world_size>1. */
int grav_b200_ctx_energy(grav_b200_ctx *ctx, double *energy);

/* Block until all queued device work of this context finished. */
int grav_b200_ctx_synchronize(grav_b200_ctx *ctx);

/* Device-side timing of the most recent ctx_acceleration(): milliseconds measured with CUDA
 * events on the context's stream.  stage: 0 total, 1 all-gather, 2 direct-sum / walk kernel,
 * 3 bbox+Morton, 4 radix sort, 5 tree build + moments.  Blocks until that work finished. */
int grav_b200_ctx_last_timing_ms(grav_b200_ctx *ctx, int stage, float *ms);
/* Benchmark plumbing.  event_record() records one of 8 user events (slot 0..7) on the context's
 * stream; event_elapsed_ms() blocks until slot_b completed and returns the device time between
 * the two.  flush_l2() overwrites a 256 MiB scratch buffer on the stream (evicts the 126 MB L2).
 * mark_positions_sharded() declares that only the owned shard of the positions is current, as
 * after a drift step, so the next ctx_acceleration() starts with the all-gather (no-op for
 * world_size 1).  host_register()/host_unregister() pin caller memory for faster H2D/D2H. */
int grav_b200_ctx_event_record(grav_b200_ctx *ctx, int slot);
int grav_b200_ctx_event_elapsed_ms(grav_b200_ctx *ctx, int slot_a, int slot_b, float *ms);
int grav_b200_ctx_flush_l2(grav_b200_ctx *ctx);
int grav_b200_ctx_mark_positions_sharded(grav_b200_ctx *ctx);
/* Which formulation the context's last pairwise force evaluation used (measurement / reporting only):
 * *pair_once = 1 for the pair-once kernel, 0 for ordered interactions; *equal_mass = 1 when that evaluation found all
 * masses equal and factored the mass out.  Synchronises the context's stream. */
int grav_b200_ctx_direct_sum_path(grav_b200_ctx *ctx, int *pair_once, int *equal_mass);
int grav_b200_host_register(void *ptr, uint64_t bytes);
int grav_b200_host_unregister(void *ptr);
/* Number of kernels this library launched since process start (all contexts). */
int64_t grav_b200_kernel_launch_count(void);

/* Measure the device's FP64 FMA peak with a register-resident DFMA loop (roofline
 * denominator for the direct sum).  Returns TFLOP/s (2 flop per DFMA). */
int grav_b200_measure_fp64_peak(int device, double *tflops, double *sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* GRAV_B200_H */
