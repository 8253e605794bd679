/*
 * ref_whfast_probe.c -- TEST INFRASTRUCTURE.  The reference's WHFast acceleration kernels are
 * `static` (src/integrator_whfast.c:136-193), so its shared library does not export them.  This
 * translation unit includes that reference source file *at compile time, from where it lies under
 * /root/reference* (nothing is copied into the repo) and exports two thin wrappers so the oracle's
 * restatement of those kernels can be pinned against the real thing.  Output: _ref/libwhfast_probe.so.
 */
#include "integrator_whfast.c"

int probe_whfast_acceleration(double *a, int n, double *x, double *m, double G, const double *jacobi_x,
                              const double *eta, int method, double eps)
{
    System s;
    s.num_particles = n; s.particle_ids = NULL; s.x = x; s.v = NULL; s.m = m; s.G = G;
    AccelerationParam p;
    p.method = method; p.opening_angle = 1.0; p.softening_length = eps; p.max_num_particles_per_leaf = 1;
    ErrorStatus st = whfast_acceleration(a, &s, jacobi_x, eta, &p);
    if (st.traceback) free(st.traceback);
    return st.return_code;
}
