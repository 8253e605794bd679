/*
 * whfast_oracle.c -- CPU restatement of the WHFast step pieces around the acceleration kernel (SURVEY.md
 * section 8f row N2).  TEST INFRASTRUCTURE ONLY; the product never links or calls it.
 *
 * Restates, with the reference's operation order and no FMA contraction (-ffp-contract=off):
 *   system_sort_by_distance   src/system.c:1210-1335      (glibc qsort == stable merge sort: ties keep index order)
 *   whfast_compute_eta        src/integrator_whfast.c:1266-1279
 *   cartesian_to_jacobi       :682-724        jacobi_to_cartesian   :726-772
 *   stumpff_functions         :774-815        whfast_drift          :424-680 (Kepler solver :459-578, removal :607-671)
 *   whfast_kick               :409-422        whfast() time loop    :200-407 (output disabled)
 * Parity status: PINNED -- tests/test_oracle.py runs the unmodified reference's launch_simulation_python with
 * integrator=whfast on the same inputs and requires bit-identical final x, v, m, ids (including a run in which
 * whfast_drift removes particles), and probes the static stage functions through oracle/ref_whfast_probe.c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "grav_oracle.h"

#define KEPLER_TOL 1e-12      /* :31 */
#define KEPLER_MAX_ITER 500   /* :32 */
#define INVALID_TOL 1e-5      /* :33 */

static double norm3(const double *p) { return sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]); }

/* src/system.c:1232-1335 */
typedef struct { double d; int idx; } DistRec;

static void merge_sort_dist(DistRec *a, DistRec *tmp, int n)
{
    if (n < 2) return;
    const int h = n / 2;
    merge_sort_dist(a, tmp, h);
    merge_sort_dist(a + h, tmp, n - h);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].d < a[i].d) ? a[j++] : a[i++];   /* stable: left wins ties (and NaN) */
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(DistRec) * (size_t)n);
}

int oracle_sort_by_distance(int n, int *ids, double *x, double *v, double *m, int primary_id)
{
    int p = -1;
    if (primary_id < n && ids[primary_id] == primary_id) p = primary_id;
    else for (int i = 0; i < n; i++) if (ids[i] == primary_id) { p = i; break; }
    if (p < 0) return -1;
    DistRec *r = malloc(sizeof(DistRec) * (size_t)n * 2);
    int *nid = malloc(sizeof(int) * (size_t)n);
    double *nx = malloc(sizeof(double) * 7 * (size_t)n);
    if (!r || !nid || !nx) { free(r); free(nid); free(nx); return -2; }
    double *nv = nx + 3 * (size_t)n, *nm = nx + 6 * (size_t)n;
    for (int i = 0; i < n; i++) {
        const double d[3] = {x[3 * i] - x[3 * p], x[3 * i + 1] - x[3 * p + 1], x[3 * i + 2] - x[3 * p + 2]};
        r[i].d = norm3(d);
        r[i].idx = i;
    }
    r[p].d = 0.0;
    merge_sort_dist(r, r + n, n);
    for (int i = 0; i < n; i++) {
        const int s = r[i].idx;
        nid[i] = ids[s];
        memcpy(nx + 3 * i, x + 3 * s, 24);
        memcpy(nv + 3 * i, v + 3 * s, 24);
        nm[i] = m[s];
    }
    memcpy(ids, nid, sizeof(int) * (size_t)n);
    memcpy(x, nx, 24 * (size_t)n);
    memcpy(v, nv, 24 * (size_t)n);
    memcpy(m, nm, 8 * (size_t)n);
    free(r); free(nid); free(nx);
    return 0;
}

/* :1266-1279 */
void oracle_whfast_eta(double *eta, int n, const double *m)
{
    eta[0] = m[0];
    for (int i = 1; i < n; i++) eta[i] = eta[i - 1] + m[i];
}

/* :682-724 */
void oracle_cartesian_to_jacobi(double *jx, double *jv, int n, const double *x, const double *v, const double *m,
                                const double *eta)
{
    double xc[3], vc[3];
    for (int k = 0; k < 3; k++) { xc[k] = m[0] * x[k]; vc[k] = m[0] * v[k]; }
    for (int i = 1; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            jx[3 * i + k] = x[3 * i + k] - xc[k] / eta[i - 1];
            jv[3 * i + k] = v[3 * i + k] - vc[k] / eta[i - 1];
            xc[k] = xc[k] * (1.0 + m[i] / eta[i - 1]) + m[i] * jx[3 * i + k];
            vc[k] = vc[k] * (1.0 + m[i] / eta[i - 1]) + m[i] * jv[3 * i + k];
        }
    }
    for (int k = 0; k < 3; k++) { jx[k] = xc[k] / eta[n - 1]; jv[k] = vc[k] / eta[n - 1]; }
}

/* :726-772 */
void oracle_jacobi_to_cartesian(double *x, double *v, int n, const double *jx, const double *jv, const double *m,
                                const double *eta)
{
    double xc[3], vc[3];
    for (int k = 0; k < 3; k++) { xc[k] = eta[n - 1] * jx[k]; vc[k] = eta[n - 1] * jv[k]; }
    for (int i = n - 1; i > 0; i--) {
        for (int k = 0; k < 3; k++) {
            xc[k] = (xc[k] - m[i] * jx[3 * i + k]) / eta[i];
            vc[k] = (vc[k] - m[i] * jv[3 * i + k]) / eta[i];
            x[3 * i + k] = jx[3 * i + k] + xc[k];
            v[3 * i + k] = jv[3 * i + k] + vc[k];
            xc[k] = eta[i - 1] * xc[k];
            vc[k] = eta[i - 1] * vc[k];
        }
    }
    for (int k = 0; k < 3; k++) { x[k] = xc[k] / m[0]; v[k] = vc[k] / m[0]; }
}

/* :774-815 */
void oracle_stumpff(double c[4], double z)
{
    int n = 0;
    while (fabs(z) > 0.1) { z /= 4.0; n++; }
    double c3 = (1.0 - z / 20.0 * (1.0 - z / 42.0 * (1.0 - z / 72.0 * (1.0 - z / 110.0 * (1.0 - z / 156.0 * (1.0 - z / 210.0)))))) / 6.0;
    double c2 = (1.0 - z / 12.0 * (1.0 - z / 30.0 * (1.0 - z / 56.0 * (1.0 - z / 90.0 * (1.0 - z / 132.0 * (1.0 - z / 182.0)))))) / 2.0;
    double c1 = 1.0 - z * c3;
    double c0 = 1.0 - z * c2;
    for (; n > 0; n--) {
        c3 = (c2 + c0 * c3) / 4.0;
        c2 = (c1 * c1) / 2.0;
        c1 = c0 * c1;
        c0 = (2.0 * c0 * c0) - 1.0;
    }
    c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3;
}

/* One particle of whfast_drift (:459-578).  Returns 1 when the particle qualifies for removal (:551). */
static int kepler_drift_one(double *px, double *pv, double gm, double dt)
{
    const double x[3] = {px[0], px[1], px[2]}, v[3] = {pv[0], pv[1], pv[2]};
    const double xn = norm3(x), vn = norm3(v);
    const double rv = (x[0] * v[0] + x[1] * v[1] + x[2] * v[2]) / xn;
    const double alpha = 2.0 * gm / xn - (vn * vn);
    double s = dt / xn;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    int converged = 0, z_bad = 0;
    for (int it = 0; it < KEPLER_MAX_ITER; it++) {
        const double z = alpha * (s * s);
        if (!isfinite(z)) { z_bad = 1; break; }
        oracle_stumpff(c, z);
        const double F = xn * s * c[1] + xn * rv * (s * s) * c[2] + gm * (s * s * s) * c[3] - dt;
        const double dF = xn * c[0] + xn * rv * s * c[1] + gm * (s * s) * c[2];
        const double ds = -F / dF;
        s += ds;
        if (fabs(ds) < KEPLER_TOL) { converged = 1; break; }
    }
    const double r = xn * c[0] + xn * rv * s * c[1] + gm * (s * s) * c[2];
    int remove = 0;
    if (!converged) {
        const double err = (xn * s * c[1] + xn * rv * (s * s) * c[2] + gm * (s * s * s) * c[3] - dt) / r;
        if (err > INVALID_TOL || z_bad) remove = 1;
    }
    const double f = 1.0 - gm * (s * s) * c[2] / xn;
    const double g = dt - gm * (s * s * s) * c[3];
    const double df = -gm * s * c[1] / (r * xn);
    const double dg = 1.0 - gm * (s * s) * c[2] / r;
    for (int k = 0; k < 3; k++) {
        px[k] = f * x[k] + g * v[k];
        pv[k] = df * x[k] + dg * v[k];
    }
    return remove;
}

/* :424-680 with the serial build's removal order (ascending index == stable compaction, src/system.c:444-528).
 * Returns the new particle count. */
int oracle_whfast_drift(int n, double *jx, double *jv, int *ids, double *x, double *v, double *m, double *eta, double G,
                        double dt, int remove_invalid)
{
    char *bad = calloc((size_t)n, 1);
    int nbad = 0;
    for (int i = 1; i < n; i++) {
        const double gm = G * m[0] * eta[i] / eta[i - 1];
        if (kepler_drift_one(jx + 3 * i, jv + 3 * i, gm, dt) && remove_invalid) { bad[i] = 1; nbad++; }
    }
    if (nbad > 0) {
        int w = 0;
        for (int i = 0; i < n; i++) {
            if (bad[i]) continue;
            if (w != i) {
                ids[w] = ids[i]; m[w] = m[i]; eta[w] = eta[i];
                memcpy(x + 3 * w, x + 3 * i, 24); memcpy(v + 3 * w, v + 3 * i, 24);
                memcpy(jx + 3 * w, jx + 3 * i, 24); memcpy(jv + 3 * w, jv + 3 * i, 24);
            }
            w++;
        }
        n = w;
        oracle_whfast_eta(eta, n, m);
    }
    free(bad);
    return n;
}

/* whfast(), :200-407, output disabled.  x, v, m, ids are updated in place the way the reference leaves them
 * (ids/m in the last sorted order; x, v from the last jacobi_to_cartesian, i.e. v at the half step).
 * max_steps < 0: run to tf.  snapshot != 0: finish like the output branch (:346-351), velocities kicked back by -dt/2
 * and converted once more.  Returns the final particle count or a negative error code. */
int oracle_whfast_integrate(int n, int *ids, double *x, double *v, double *m, double G, double dt, double tf, int method,
                            double eps, int remove_invalid, int64_t max_steps, int snapshot, double *a_out)
{
    if (method != 1 && method != 2) return -3;
    /* the reference mallocs `a` and never writes a[0..2] (:231, :857), yet kicks jacobi_v[0] with it (:416-421): a read of
     * uninitialised memory that is zero in practice (fresh heap / mmap pages).  Defined as zero here and on the GPU. */
    double *jx = calloc(3 * (size_t)n, 8), *jv = malloc(24 * (size_t)n), *a = calloc(3 * (size_t)n, 8);
    double *eta = malloc(8 * (size_t)n);
    if (!jx || !jv || !a || !eta) return -2;
    int rc = oracle_sort_by_distance(n, ids, x, v, m, 0);
    if (rc) goto done;
    oracle_whfast_eta(eta, n, m);
    oracle_cartesian_to_jacobi(jx, jv, n, x, v, m, eta);
#define ACCEL() do { if (method == 1) oracle_whfast_pairwise(a, n, x, m, G, jx, eta, eps); \
                     else oracle_whfast_massless(a, n, x, m, G, jx, eta, eps); } while (0)
#define KICK(h) do { for (int q = 0; q < 3 * n; q++) jv[q] += a[q] * (h); } while (0)
    ACCEL();
    KICK(0.5 * dt);
    const int64_t total = (int64_t)ceil(tf / dt);
    double t = 0.0;
    for (int64_t s = 0; s < total && (max_steps < 0 || s < max_steps); ) {
        if (t + dt > tf) dt = tf - t;
        rc = oracle_sort_by_distance(n, ids, jx, jv, m, 0);     /* system->x/v swapped to the Jacobi arrays, :301-311 */
        if (rc) goto done;
        oracle_whfast_eta(eta, n, m);
        n = oracle_whfast_drift(n, jx, jv, ids, x, v, m, eta, G, dt, remove_invalid);
        oracle_jacobi_to_cartesian(x, v, n, jx, jv, m, eta);
        ACCEL();
        KICK(dt);
        s++;
        t = (double)s * dt;
    }
    if (snapshot) {
        double *tv = malloc(24 * (size_t)n);
        if (!tv) { rc = -2; goto done; }
        for (int q = 0; q < 3 * n; q++) tv[q] = jv[q] + a[q] * (-0.5 * dt);
        oracle_jacobi_to_cartesian(x, v, n, jx, tv, m, eta);
        free(tv);
    }
    if (a_out) memcpy(a_out, a, 24 * (size_t)n);
    rc = n;
done:
    free(jx); free(jv); free(a); free(eta);
    return rc;
}
