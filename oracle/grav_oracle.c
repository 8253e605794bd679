/*
 * grav_oracle.c -- CPU restatement of grav_sim's acceleration hot path.  TEST INFRASTRUCTURE ONLY
 * (see grav_oracle.h for who may use it and for the parity status: PINNED against
 * oracle/_ref/libgrav_sim_ref.so and tests/golden/).
 *
 * The direct-sum functions keep the reference's summation order so they can be compared bit for
 * bit.  The octree is NOT built the way the reference builds it (a serial depth-first walk with
 * realloc); it is restated in the level-parallel form the GPU kernels use -- breadth-first node
 * discovery, ids from a prefix sum over start positions, moments by reverse sweep -- and the
 * tests prove the two give identical arrays.
 */
#include "grav_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LEVEL 21

/* ------------------------------------------------------------------------------------------ */
/* direct sums                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* one unordered pair, both partners updated: src/acceleration.c:204-230 */
static inline void pair_kick(double *a, const double *x, int i, int j, double mi, double mj, double G, double eps2)
{
    const double rx = x[3 * i] - x[3 * j], ry = x[3 * i + 1] - x[3 * j + 1], rz = x[3 * i + 2] - x[3 * j + 2];
    const double r = sqrt(rx * rx + ry * ry + rz * rz + eps2);
    const double f = G / (r * r * r);
    const double fx = f * rx, fy = f * ry, fz = f * rz;
    a[3 * i] -= fx * mj; a[3 * i + 1] -= fy * mj; a[3 * i + 2] -= fz * mj;
    a[3 * j] += fx * mi; a[3 * j + 1] += fy * mi; a[3 * j + 2] += fz * mi;
}

void oracle_pairwise(double *a, int n, const double *x, const double *m, double G, double eps)
{
    const double eps2 = eps * eps;
    memset(a, 0, sizeof(double) * 3 * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) pair_kick(a, x, i, j, m[i], m[j], G, eps2);
}

/* One target of oracle_pairwise(): the operations the reference's i<j loop applies to a[i], in the order it applies
 * them -- pairs (j,i) with j < i arrive first, as the "+= fx*mi" half of pair_kick(j, i) in outer iteration j, then the
 * pairs (i,j) with j > i in outer iteration i (src/acceleration.c:198-231).  Bit-identical to oracle_pairwise()[i]
 * (tests/test_oracle.py), which makes reference comparisons affordable at N = 2^20 (a few hundred targets). */
void oracle_pairwise_targets(double *a, int nt, const int *targets, int n, const double *x, const double *m, double G,
                             double eps)
{
    const double eps2 = eps * eps;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < nt; t++) {
        const int i = targets[t];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < i; j++) {
            const double rx = x[3 * j] - x[3 * i], ry = x[3 * j + 1] - x[3 * i + 1], rz = x[3 * j + 2] - x[3 * i + 2];
            const double r = sqrt(rx * rx + ry * ry + rz * rz + eps2);
            const double f = G / (r * r * r);
            const double fx = f * rx, fy = f * ry, fz = f * rz;
            ax += fx * m[j]; ay += fy * m[j]; az += fz * m[j];
        }
        for (int j = i + 1; j < n; j++) {
            const double rx = x[3 * i] - x[3 * j], ry = x[3 * i + 1] - x[3 * j + 1], rz = x[3 * i + 2] - x[3 * j + 2];
            const double r = sqrt(rx * rx + ry * ry + rz * rz + eps2);
            const double f = G / (r * r * r);
            const double fx = f * rx, fy = f * ry, fz = f * rz;
            ax -= fx * m[j]; ay -= fy * m[j]; az -= fz * m[j];
        }
        a[3 * t] = ax; a[3 * t + 1] = ay; a[3 * t + 2] = az;
    }
}

/* The same sum in extended precision (x87 long double, 64-bit significand): the yardstick that tells the reference's own
 * rounding noise (serial double sum of N terms) from a real disagreement when N = 2^20. */
void oracle_pairwise_targets_ld(double *a, int nt, const int *targets, int n, const double *x, const double *m, double G,
                                double eps)
{
    const long double eps2 = (long double)eps * eps;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < nt; t++) {
        const int i = targets[t];
        long double ax = 0.0L, ay = 0.0L, az = 0.0L;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            const long double rx = (long double)x[3 * i] - x[3 * j], ry = (long double)x[3 * i + 1] - x[3 * j + 1],
                              rz = (long double)x[3 * i + 2] - x[3 * j + 2];
            const long double r = sqrtl(rx * rx + ry * ry + rz * rz + eps2);
            const long double f = (long double)m[j] / (r * r * r);
            ax -= f * rx; ay -= f * ry; az -= f * rz;
        }
        a[3 * t] = (double)(G * ax); a[3 * t + 1] = (double)(G * ay); a[3 * t + 2] = (double)(G * az);
    }
}

static int split_by_mass(int n, const double *m, int **massive, int *n_massive, int **massless, int *n_massless)
{
    int *hv = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *lt = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (!hv || !lt) { free(hv); free(lt); return -1; }
    int nh = 0, nl = 0;
    for (int i = 0; i < n; i++) {
        if (m[i] != 0.0) hv[nh++] = i; else lt[nl++] = i;
    }
    *massive = hv; *n_massive = nh; *massless = lt; *n_massless = nl;
    return 0;
}

int oracle_massless(double *a, int n, const double *x, const double *m, double G, double eps)
{
    const double eps2 = eps * eps;
    int *hv, *lt, nh, nl;
    if (split_by_mass(n, m, &hv, &nh, &lt, &nl)) return -1;
    memset(a, 0, sizeof(double) * 3 * (size_t)n);
    /* massive x massive, Newton-3 (src/acceleration.c:299-333) */
    for (int p = 0; p < nh; p++)
        for (int q = p + 1; q < nh; q++) pair_kick(a, x, hv[p], hv[q], m[hv[p]], m[hv[q]], G, eps2);
    /* massive -> massless (:336-361): outer loop over massive rank p, mass taken as m[p] (sic) */
    for (int p = 0; p < nh; p++) {
        const int i = hv[p];
        for (int q = 0; q < nl; q++) {
            const int j = lt[q];
            const double rx = x[3 * i] - x[3 * j], ry = x[3 * i + 1] - x[3 * j + 1], rz = x[3 * i + 2] - x[3 * j + 2];
            const double r = sqrt(rx * rx + ry * ry + rz * rz + eps2);
            const double f = G / (r * r * r);
            a[3 * j] += f * rx * m[p];
            a[3 * j + 1] += f * ry * m[p];
            a[3 * j + 2] += f * rz * m[p];
        }
    }
    free(hv);
    free(lt);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* WHFast interaction-term accelerations (Jacobi coordinates)                                 */
/* ------------------------------------------------------------------------------------------ */

static inline double norm3(const double *v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* d = x[to] - x[from]; returns |d|^3 + eps^3 (src/integrator_whfast.c:871-874) */
static inline double sep_cubed(double d[3], const double *x, int to, int from, double eps3)
{
    d[0] = x[3 * to] - x[3 * from];
    d[1] = x[3 * to + 1] - x[3 * from + 1];
    d[2] = x[3 * to + 2] - x[3 * from + 2];
    const double r = norm3(d);
    return r * r * r + eps3;
}

/* the central-body + Jacobi term written straight into a[i] (:866-883); ratio_on selects eta_i/eta_{i-1} */
static void whfast_central_term(double *a, int i, const double *x, const double *m, double G, const double *jx,
                                const double *eta, double eps3, int ratio_on)
{
    double d[3];
    const double c0 = sep_cubed(d, x, i, 0, eps3);
    const double jn = norm3(&jx[3 * i]);
    const double cj = jn * jn * jn + eps3;
    for (int k = 0; k < 3; k++) {
        if (ratio_on)
            a[3 * i + k] = G * m[0] * eta[i] / eta[i - 1] * (jx[3 * i + k] / cj - d[k] / c0);
        else
            a[3 * i + k] = G * m[0] * (jx[3 * i + k] / cj - d[k] / c0);
    }
}

void oracle_whfast_pairwise(double *a, int n, const double *x, const double *m, double G, const double *jx,
                            const double *eta, double eps)
{
    const double eps3 = eps * eps * eps;
    for (int i = 1; i < n; i++) {
        double s[3] = {0.0, 0.0, 0.0}, d[3];
        whfast_central_term(a, i, x, m, G, jx, eta, eps3, 1);
        for (int j = 1; j < i; j++) {               /* inner planets, :885-899 */
            const double c = sep_cubed(d, x, i, j, eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[j] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] -= s[k] * eta[i] / eta[i - 1]; s[k] = 0.0; }
        for (int j = i + 1; j < n; j++) {           /* outer planets, :905-919 */
            const double c = sep_cubed(d, x, j, i, eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[j] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] += s[k]; s[k] = 0.0; }
        for (int j = 0; j < i; j++)                 /* pairs straddling i, :925-944 */
            for (int q = i + 1; q < n; q++) {
                const double c = sep_cubed(d, x, q, j, eps3);
                for (int k = 0; k < 3; k++) s[k] += G * m[j] * m[q] * d[k] / c;
            }
        for (int k = 0; k < 3; k++) a[3 * i + k] -= s[k] / eta[i - 1];
    }
}

int oracle_whfast_massless(double *a, int n, const double *x, const double *m, double G, const double *jx,
                           const double *eta, double eps)
{
    const double eps3 = eps * eps * eps;
    int *hv, *lt, nh, nl;
    if (split_by_mass(n, m, &hv, &nh, &lt, &nl)) return -1;
    /* massive targets, :1012-1121 (list position p >= 1; the list's first entry is skipped as "the star") */
    for (int p = 1; p < nh; p++) {
        const int i = hv[p];
        double s[3] = {0.0, 0.0, 0.0}, d[3];
        whfast_central_term(a, i, x, m, G, jx, eta, eps3, 1);
        for (int q = 1; q < p; q++) {
            const double c = sep_cubed(d, x, i, hv[q], eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] -= s[k] * eta[i] / eta[i - 1]; s[k] = 0.0; }
        for (int q = p + 1; q < nh; q++) {
            const double c = sep_cubed(d, x, hv[q], i, eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] += s[k]; s[k] = 0.0; }
        for (int q = 0; q < p; q++)
            for (int r = p + 1; r < nh; r++) {
                const double c = sep_cubed(d, x, hv[r], hv[q], eps3);
                for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * m[hv[r]] * d[k] / c;
            }
        for (int k = 0; k < 3; k++) a[3 * i + k] -= s[k] / eta[i - 1];
    }
    /* massless targets, :1127-1258 */
    for (int t = 0; t < nl; t++) {
        const int i = lt[t];
        if (i == 0) continue;
        double s[3] = {0.0, 0.0, 0.0}, d[3];
        whfast_central_term(a, i, x, m, G, jx, eta, eps3, 0);
        for (int q = 1; q < nh; q++) {              /* massive particles before i */
            if (hv[q] >= i) break;
            const double c = sep_cubed(d, x, i, hv[q], eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] -= s[k]; s[k] = 0.0; }
        for (int q = 1; q < nh; q++) {              /* massive particles after i */
            if (hv[q] <= i) continue;
            const double c = sep_cubed(d, x, hv[q], i, eps3);
            for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * d[k] / c;
        }
        for (int k = 0; k < 3; k++) { a[3 * i + k] += s[k]; s[k] = 0.0; }
        for (int q = 0; q < nh; q++) {              /* massive pairs straddling i */
            if (hv[q] >= i) break;
            for (int r = q + 1; r < nh; r++) {
                if (hv[r] <= i) continue;
                const double c = sep_cubed(d, x, hv[r], hv[q], eps3);
                for (int k = 0; k < 3; k++) s[k] += G * m[hv[q]] * m[hv[r]] * d[k] / c;
            }
        }
        for (int k = 0; k < 3; k++) a[3 * i + k] -= s[k] / eta[i - 1];
    }
    free(hv);
    free(lt);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Barnes-Hut: bounding box, keys, sort                                                       */
/* ------------------------------------------------------------------------------------------ */

void oracle_bounding_box(double center[3], double *width, int n, const double *x)
{
    double lo[3] = {x[0], x[1], x[2]}, hi[3] = {x[0], x[1], x[2]};
    for (int i = 1; i < n; i++)
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], x[3 * i + k]);
            hi[k] = fmax(hi[k], x[3 * i + k]);
        }
    for (int k = 0; k < 3; k++) center[k] = (hi[k] + lo[k]) / 2.0;
    *width = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
}

/* 21 bits -> every third bit (masks of src/linear_octree.c:180-184) */
static inline int64_t spread3(int64_t v)
{
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffff;
    v = (v | v << 16) & 0x1f0000ff0000ff;
    v = (v | v << 8) & 0x100f00f00f00f00f;
    v = (v | v << 4) & 0x10c30c30c30c30c3;
    v = (v | v << 2) & 0x1249249249249249;
    return v;
}

/* (int64) of a double as x86-64 does it (cvttsd2si): NaN and out-of-range give INT64_MIN, which is
 * what the reference binary produces for the width==0 case; spelled out to avoid C UB here. */
static inline int64_t trunc_to_i64(double u)
{
    if (!(u > -9.3e18 && u < 9.3e18)) return INT64_MIN;
    return (int64_t)u;
}

void oracle_morton_keys(int64_t *keys, int n, const double *x, const double center[3], double width)
{
    for (int i = 0; i < n; i++) {
        int64_t c[3];
        for (int k = 0; k < 3; k++) {
            const double u = (x[3 * i + k] - center[k]) / width + 0.5;
            c[k] = trunc_to_i64(u * (double)(1 << 21));   /* u == 1.0 wraps to cell 0 after the mask (reference quirk) */
        }
        keys[i] = spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
    }
}

/* Stable LSD sort, 4 passes of 16 bits (the reference does 7 of 9: same result, any stable sort is). */
int oracle_sort_keys(int64_t *keys, int *perm, int n)
{
    int64_t *k2 = malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int *p2 = malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    size_t *cnt = malloc(sizeof(size_t) * 65537);
    if (!k2 || !p2 || !cnt) { free(k2); free(p2); free(cnt); return -1; }
    for (int i = 0; i < n; i++) perm[i] = i;
    int64_t *ka = keys, *kb = k2;
    int *pa = perm, *pb = p2;
    for (int pass = 0; pass < 4; pass++) {
        const int sh = 16 * pass;
        memset(cnt, 0, sizeof(size_t) * 65537);
        for (int i = 0; i < n; i++) cnt[((uint64_t)ka[i] >> sh & 0xffff) + 1]++;
        for (int d = 0; d < 65536; d++) cnt[d + 1] += cnt[d];
        for (int i = 0; i < n; i++) {
            const size_t dst = cnt[(uint64_t)ka[i] >> sh & 0xffff]++;
            kb[dst] = ka[i];
            pb[dst] = pa[i];
        }
        int64_t *tk = ka; ka = kb; kb = tk;
        int *tp = pa; pa = pb; pb = tp;
    }
    /* 4 passes: data is back in the caller's arrays */
    free(k2);
    free(p2);
    free(cnt);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Barnes-Hut: tree                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct Expanded {   /* a node that has children, in breadth-first discovery order */
    int level, s, e;        /* covers sorted positions [s, e) */
    int parent, rank;       /* index of the parent in this list (-1 root), position among its siblings */
    int same_start_before;  /* children owned by expanded ancestors that start at the same s */
    int nch, ch_s[8], ch_n[8];
    int first_child, id;
} Expanded;

void oracle_free_tree(OracleTree *t)
{
    free(t->keys); free(t->perm); free(t->num_particles); free(t->num_children); free(t->first_particle);
    free(t->first_child); free(t->mass); free(t->com_x); free(t->com_y); free(t->com_z);
    memset(t, 0, sizeof(*t));
}

int oracle_build_tree(OracleTree *t, int n, const double *x, const double *m, int max_leaf, const double *box_center,
                      double box_width)
{
    memset(t, 0, sizeof(*t));
    t->n = n;
    double center[3];
    if (!box_center || box_width <= 0.0) {
        oracle_bounding_box(center, &t->box_width, n, x);
    } else {
        t->box_width = box_width;
        memcpy(center, box_center, sizeof(center));
    }
    t->keys = malloc(sizeof(int64_t) * (size_t)n);
    t->perm = malloc(sizeof(int) * (size_t)n);
    if (!t->keys || !t->perm) { oracle_free_tree(t); return -1; }
    oracle_morton_keys(t->keys, n, x, center, t->box_width);
    if (oracle_sort_keys(t->keys, t->perm, n)) { oracle_free_tree(t); return -1; }
    const int64_t *K = t->keys;

    /* --- node discovery, one level at a time ------------------------------
     * The root is always expanded (src/linear_octree.c:639-650).  The children of an expanded
     * node at level l-1 are the maximal runs of equal (key >> 3(21-l)) inside its range, in key
     * order (:342-393, :446-572).  A child is expanded iff it holds more than max_leaf particles
     * and l < 21 (:662). */
    size_t cap = 1024, ne = 0;
    Expanded *ex = malloc(sizeof(Expanded) * cap);
    int *W = calloc((size_t)n + 1, sizeof(int));       /* children created by nodes starting at position p */
    if (!ex || !W) { free(ex); free(W); oracle_free_tree(t); return -1; }
    memset(&ex[0], 0, sizeof(Expanded));
    ex[0].level = 0; ex[0].s = 0; ex[0].e = n; ex[0].parent = -1;
    ne = 1;
    for (size_t u = 0; u < ne; u++) {
        const int l = ex[u].level + 1, sh = 3 * (MAX_LEVEL - l);
        int nch = 0;
        for (int p = ex[u].s; p < ex[u].e;) {
            int q = p + 1;
            while (q < ex[u].e && (K[q] >> sh) == (K[p] >> sh)) q++;
            ex[u].ch_s[nch] = p;
            ex[u].ch_n[nch] = q - p;
            nch++;
            p = q;
        }
        ex[u].nch = nch;
        W[ex[u].s] += nch;
        for (int k = 0; k < nch; k++) {
            if (ex[u].ch_n[k] > max_leaf && l < MAX_LEVEL) {
                if (ne == cap) {
                    cap *= 2;
                    Expanded *tmp = realloc(ex, sizeof(Expanded) * cap);
                    if (!tmp) { free(ex); free(W); oracle_free_tree(t); return -1; }
                    ex = tmp;
                }
                Expanded *c = &ex[ne++];
                memset(c, 0, sizeof(*c));
                c->level = l; c->s = ex[u].ch_s[k]; c->e = c->s + ex[u].ch_n[k];
                c->parent = (int)u; c->rank = k;
                c->same_start_before = (c->s == ex[u].s) ? ex[u].same_start_before + nch : 0;
            }
        }
    }

    /* --- numbering --------------------------------------------------------------------------
     * The reference hands out ids when a node is expanded, all its children at once, and expands
     * depth-first (:457, :560-572, :652-705): expanded nodes are therefore served in the order
     * (start position, level), and first_child = 1 + children of everything served earlier
     *             = 1 + [children of nodes starting left of s] + [children of ancestors starting at s]. */
    int run = 0;
    for (int p = 0; p <= n; p++) { const int w = W[p]; W[p] = run; run += w; }   /* exclusive scan */
    const int M = 1 + run;
    t->num_nodes = M;
    for (size_t u = 0; u < ne; u++) ex[u].first_child = 1 + W[ex[u].s] + ex[u].same_start_before;
    ex[0].id = 0;
    for (size_t u = 1; u < ne; u++) ex[u].id = ex[ex[u].parent].first_child + ex[u].rank;   /* parents precede children */

    t->num_particles = malloc(sizeof(int) * (size_t)M);
    t->num_children = malloc(sizeof(int) * (size_t)M);
    t->first_particle = malloc(sizeof(int) * (size_t)M);
    t->first_child = malloc(sizeof(int) * (size_t)M);
    t->mass = calloc((size_t)M, sizeof(double));
    t->com_x = calloc((size_t)M, sizeof(double));
    t->com_y = calloc((size_t)M, sizeof(double));
    t->com_z = calloc((size_t)M, sizeof(double));
    double *sum_mx = calloc(3 * (size_t)M, sizeof(double));
    if (!t->num_particles || !t->num_children || !t->first_particle || !t->first_child || !t->mass || !t->com_x ||
        !t->com_y || !t->com_z || !sum_mx) {
        free(ex); free(W); free(sum_mx); oracle_free_tree(t); return -1;
    }
    t->num_particles[0] = n; t->first_particle[0] = 0;
    for (size_t u = 0; u < ne; u++)
        for (int k = 0; k < ex[u].nch; k++) {
            const int id = ex[u].first_child + k;
            t->num_particles[id] = ex[u].ch_n[k];
            t->first_particle[id] = ex[u].ch_s[k];
            t->num_children[id] = 0;      /* leaf until proven otherwise */
            t->first_child[id] = -1;      /* the reference leaves this uninitialised for leaves */
        }
    for (size_t u = 0; u < ne; u++) {
        t->num_children[ex[u].id] = ex[u].nch;
        t->first_child[ex[u].id] = ex[u].first_child;
    }

    /* --- moments (:667-674, :712-730): children in id order; a leaf child adds its particles one by
     * one in sorted order, an expanded child adds its finished sums; leaves keep mass = com = 0 (:567-570).
     * Deeper nodes come later in breadth-first order, so a reverse sweep sees children first. */
    for (size_t u = ne; u-- > 0;) {
        double tot = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
        for (int k = 0; k < ex[u].nch; k++) {
            const int id = ex[u].first_child + k;
            if (t->num_children[id] == 0) {
                for (int p = ex[u].ch_s[k]; p < ex[u].ch_s[k] + ex[u].ch_n[k]; p++) {
                    const int i = t->perm[p];
                    tot += m[i];
                    sx += m[i] * x[3 * i];
                    sy += m[i] * x[3 * i + 1];
                    sz += m[i] * x[3 * i + 2];
                }
            } else {
                tot += t->mass[id];
                sx += sum_mx[3 * id]; sy += sum_mx[3 * id + 1]; sz += sum_mx[3 * id + 2];
            }
        }
        const int id = ex[u].id;
        t->mass[id] = tot;
        sum_mx[3 * id] = sx; sum_mx[3 * id + 1] = sy; sum_mx[3 * id + 2] = sz;
        t->com_x[id] = sx / tot; t->com_y[id] = sy / tot; t->com_z[id] = sz / tot;
    }
    free(sum_mx);
    free(ex);
    free(W);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Barnes-Hut: walk                                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct WalkCtx {
    const OracleTree *t;
    const double *x, *m;
    double G, eps2, theta2, box_length;
    int fixed;
    /* per target */
    int self;
    int64_t key;
    double xi[3], acc[3];
    /* statistics of the current target (oracle_bh_walk_stats) */
    long long visits, accepts, opened, leaf_particles;
    int *opened_ids;        /* optional: ids of the opened nodes, in walk order */
    long long opened_cap;
} WalkCtx;

static void walk_children(WalkCtx *w, int node, int level)
{
    const OracleTree *t = w->t;
    const int sh = 3 * (MAX_LEVEL - level);
    for (int k = 0; k < t->num_children[node]; k++) {
        const int c = t->first_child[node] + k;
        const int s = t->first_particle[c];
        const int is_leaf = t->num_children[c] <= 0;
        /* reference mode: both keys are fetched from the SORTED array with ORIGINAL particle ids
         * (src/acceleration_barnes_hut.c:120,145) */
        const int64_t kc = w->fixed ? t->keys[s] : t->keys[t->perm[s]];
        const int inside = (w->key >> sh) == (kc >> sh);
        w->visits++;
        if (!inside && !(w->fixed && is_leaf)) {
            const double rx = w->xi[0] - t->com_x[c], ry = w->xi[1] - t->com_y[c], rz = w->xi[2] - t->com_z[c];
            const double cell = w->box_length / (2 << level);
            const double d2 = rx * rx + ry * ry + rz * rz;
            if (cell * cell < w->theta2 * d2) {          /* accepted; a leaf has mass 0 here, i.e. is dropped */
                const double r = sqrt(d2 + w->eps2);
                const double f = w->G * t->mass[c] / (r * r * r);
                w->acc[0] -= f * rx; w->acc[1] -= f * ry; w->acc[2] -= f * rz;
                w->accepts++;
                continue;
            }
        }
        if (is_leaf) {
            for (int p = s; p < s + t->num_particles[c]; p++) {
                const int j = t->perm[p];
                if (j == w->self) continue;
                w->leaf_particles++;
                const double rx = w->xi[0] - w->x[3 * j], ry = w->xi[1] - w->x[3 * j + 1], rz = w->xi[2] - w->x[3 * j + 2];
                const double r = sqrt(rx * rx + ry * ry + rz * rz + w->eps2);
                const double f = w->G * w->m[j] / (r * r * r);
                w->acc[0] -= f * rx; w->acc[1] -= f * ry; w->acc[2] -= f * rz;
            }
        } else {
            if (w->opened_ids && w->opened < w->opened_cap) w->opened_ids[w->opened] = c;
            w->opened++;
            walk_children(w, c, level + 1);
        }
    }
}

static void walk_one(WalkCtx *w, const OracleTree *t, const double *x, int p, int fixed_mode);

/* ids of the nodes the walk of the target at sorted position p opens (design studies of shared traversals); returns the count */
long long oracle_bh_walk_opened(const OracleTree *t, const double *x, const double *m, double theta, int fixed_mode, int p,
                                int *ids, long long cap)
{
    WalkCtx w;
    w.t = t; w.x = x; w.m = m; w.G = 1.0; w.eps2 = 0.0; w.theta2 = theta * theta;
    w.box_length = t->box_width * 2.0;
    w.fixed = fixed_mode;
    w.opened_ids = ids; w.opened_cap = cap;
    walk_one(&w, t, x, p, fixed_mode);
    return w.opened;
}

void oracle_bh_walk(double *a, const OracleTree *t, const double *x, const double *m, double G, double eps,
                    double theta, int fixed_mode)
{
    WalkCtx w;
    w.opened_ids = NULL; w.opened_cap = 0;
    w.t = t; w.x = x; w.m = m; w.G = G; w.eps2 = eps * eps; w.theta2 = theta * theta;
    w.box_length = t->box_width * 2.0;
    w.fixed = fixed_mode;
    for (int p = 0; p < t->n; p++) {
        const int i = t->perm[p];
        walk_one(&w, t, x, p, fixed_mode);
        a[3 * i] = w.acc[0]; a[3 * i + 1] = w.acc[1]; a[3 * i + 2] = w.acc[2];
    }
}

static void walk_one(WalkCtx *w, const OracleTree *t, const double *x, int p, int fixed_mode)
{
    const int i = t->perm[p];
    w->self = i;
    w->key = fixed_mode ? t->keys[p] : t->keys[i];
    w->xi[0] = x[3 * i]; w->xi[1] = x[3 * i + 1]; w->xi[2] = x[3 * i + 2];
    w->acc[0] = w->acc[1] = w->acc[2] = 0.0;
    w->visits = w->accepts = w->opened = w->leaf_particles = 0;
    walk_children(w, 0, 1);
}

/* oracle_bh_walk() for a subset of targets given by SORTED POSITION; a[3*k..] is the acceleration of particle
 * perm[positions[k]].  Same code path per target, so bit-identical to the full walk (tests/test_oracle.py); this is what
 * makes a reference comparison at N = 2^24 affordable.  stats (may be NULL): 4 counters per target -- nodes visited,
 * nodes accepted, nodes opened, leaf particles summed directly. */
void oracle_bh_walk_targets(double *a, long long *stats, const OracleTree *t, const double *x, const double *m, double G,
                            double eps, double theta, int fixed_mode, int nt, const int *positions)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int k = 0; k < nt; k++) {
        WalkCtx w;
        w.t = t; w.x = x; w.m = m; w.G = G; w.eps2 = eps * eps; w.theta2 = theta * theta;
        w.box_length = t->box_width * 2.0;
        w.fixed = fixed_mode;
        w.opened_ids = NULL; w.opened_cap = 0;
        walk_one(&w, t, x, positions[k], fixed_mode);
        a[3 * k] = w.acc[0]; a[3 * k + 1] = w.acc[1]; a[3 * k + 2] = w.acc[2];
        if (stats) {
            stats[4 * k] = w.visits; stats[4 * k + 1] = w.accepts; stats[4 * k + 2] = w.opened;
            stats[4 * k + 3] = w.leaf_particles;
        }
    }
}

int oracle_barnes_hut(double *a, int n, const double *x, const double *m, double G, double eps, double theta,
                      int max_leaf, int fixed_mode)
{
    OracleTree t;
    if (oracle_build_tree(&t, n, x, m, max_leaf, NULL, -1.0)) return -1;
    oracle_bh_walk(a, &t, x, m, G, eps, theta, fixed_mode);
    oracle_free_tree(&t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */

double oracle_energy(int n, const double *x, const double *v, const double *m, double G)
{
    double e = 0.0;
    for (int i = 0; i < n; i++) {
        const double vn = norm3(&v[3 * i]);
        e += 0.5 * m[i] * vn * vn;
        for (int j = i + 1; j < n; j++) {
            const double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
            e -= G * m[i] * m[j] / norm3(d);
        }
    }
    return e;
}
