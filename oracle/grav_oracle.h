/*
 * grav_oracle.h -- CPU restatement of grav_sim's acceleration hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libgrav_b200.so, libgrav_sim_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here against
 * the unmodified reference compiled to oracle/_ref/libgrav_sim_ref.so (bit-exact for keys,
 * permutation, node arrays, moments and BH accelerations; bit-exact for the pairwise / massless /
 * WHFast sums, which use the reference's summation order), and tests/golden/ holds vectors minted
 * from that reference build by tests/golden/make_golden.py.
 *
 * Every function cites the reference lines it restates (paths relative to the upstream tree).
 * Compile with -ffp-contract=off: the reference is built without FMA contraction
 * (CMakeLists.txt:56-61: -O3, no -march, no -ffast-math).
 */
#ifndef GRAV_ORACLE_H
#define GRAV_ORACLE_H
#include <stdint.h>

/* src/acceleration.c:177-234 */
void oracle_pairwise(double *a, int n, const double *x, const double *m, double G, double eps);
/* the entries targets[0..nt) of oracle_pairwise()'s result, each with the reference's operation order for that target
 * (bit-identical to the full loop); _ld: the same sums in long double, as a yardstick for rounding noise */
void oracle_pairwise_targets(double *a, int nt, const int *targets, int n, const double *x, const double *m, double G,
                             double eps);
void oracle_pairwise_targets_ld(double *a, int nt, const int *targets, int n, const double *x, const double *m, double G,
                                double eps);
/* src/acceleration.c:236-367 (with the m[rank] indexing of :357-359) */
int oracle_massless(double *a, int n, const double *x, const double *m, double G, double eps);
/* src/integrator_whfast.c:839-957 and :959-1264; aux[] defined as zero-initialised */
void oracle_whfast_pairwise(double *a, int n, const double *x, const double *m, double G,
                            const double *jacobi_x, const double *eta, double eps);
int oracle_whfast_massless(double *a, int n, const double *x, const double *m, double G,
                           const double *jacobi_x, const double *eta, double eps);

/* src/linear_octree.c:113-146 */
void oracle_bounding_box(double center[3], double *width, int n, const double *x);
/* src/linear_octree.c:159-202 */
void oracle_morton_keys(int64_t *keys, int n, const double *x, const double center[3], double width);
/* src/linear_octree.c:216-326: result == stable sort by key; keys permuted in place, perm out */
int oracle_sort_keys(int64_t *keys, int *perm, int n);

typedef struct OracleTree {
    double box_width;
    int n, num_nodes;
    int64_t *keys;   /* sorted */
    int *perm;       /* sorted_indices */
    int *num_particles, *num_children, *first_particle, *first_child; /* first_child = -1 for leaves */
    double *mass, *com_x, *com_y, *com_z;
} OracleTree;

/* src/linear_octree.c:406-576, 590-736, 825-962, restated level-by-level.
 * box_center NULL or box_width <= 0: automatic bounding box. Returns 0, or -1 on allocation failure. */
int oracle_build_tree(OracleTree *t, int n, const double *x, const double *m, int max_leaf,
                      const double *box_center, double box_width);
void oracle_free_tree(OracleTree *t);

/* src/acceleration_barnes_hut.c:78-248.  fixed_mode 0 = bug-for-bug, 1 = corrected walk. */
void oracle_bh_walk(double *a, const OracleTree *t, const double *x, const double *m, double G,
                    double eps, double theta, int fixed_mode);
/* the walk for the targets at the given sorted positions only (a[3k..] belongs to particle perm[positions[k]]);
 * stats: NULL or 4 counters per target (visits, accepts, opened, leaf particles) */
void oracle_bh_walk_targets(double *a, long long *stats, const OracleTree *t, const double *x, const double *m, double G,
                            double eps, double theta, int fixed_mode, int nt, const int *positions);
/* src/acceleration_barnes_hut.c:33-76 */
int oracle_barnes_hut(double *a, int n, const double *x, const double *m, double G, double eps,
                      double theta, int max_leaf, int fixed_mode);

/* src/utils.c:27-59 */
double oracle_energy(int n, const double *x, const double *v, const double *m, double G);

/* ---- WHFast step pieces (whfast_oracle.c), SURVEY.md section 8f row N2 ---- */
/* src/system.c:1210-1335 (stable by distance from the particle whose id is primary_id); 0 ok, -1 id not found */
int oracle_sort_by_distance(int n, int *ids, double *x, double *v, double *m, int primary_id);
/* src/integrator_whfast.c:1266-1279 */
void oracle_whfast_eta(double *eta, int n, const double *m);
/* :682-724 and :726-772 */
void oracle_cartesian_to_jacobi(double *jx, double *jv, int n, const double *x, const double *v, const double *m,
                                const double *eta);
void oracle_jacobi_to_cartesian(double *x, double *v, int n, const double *jx, const double *jv, const double *m,
                                const double *eta);
/* :774-815, c = {c0, c1, c2, c3} */
void oracle_stumpff(double c[4], double z);
/* :424-680; returns the particle count after the removal of invalid particles */
int oracle_whfast_drift(int n, double *jx, double *jv, int *ids, double *x, double *v, double *m, double *eta, double G,
                        double dt, int remove_invalid);
/* :200-407 with output disabled; returns the final particle count (<0: error) */
int oracle_whfast_integrate(int n, int *ids, double *x, double *v, double *m, double G, double dt, double tf, int method,
                            double eps, int remove_invalid, int64_t max_steps, int snapshot, double *a_out);

#endif
