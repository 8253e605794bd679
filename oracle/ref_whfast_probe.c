/*
 * ref_whfast_probe.c -- TEST INFRASTRUCTURE.  The reference's WHFast acceleration kernels are
 * `static` (src/integrator_whfast.c:136-193), so its shared library does not export them.  This
 * translation unit includes that reference source file *at compile time, from where it lies under
 * /root/reference* (nothing is copied into the repo) and exports two thin wrappers so the oracle's
 * restatement of those kernels can be pinned against the real thing.  Output: _ref/libwhfast_probe.so.
 */
#include "integrator_whfast.c"
#include "integrator.h"

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

/* whfast() itself (src/integrator_whfast.c:200-407) with output and progress bar disabled, called with the
 * reference's own parameter structs.  launch_simulation_python does not report the particle count after
 * whfast_drift removed particles, so this wrapper returns it. */
int probe_whfast_run(int n, int *ids, double *x, double *v, double *m, double G, double dt, double tf, int method,
                     double eps, int remove_invalid, int verbose)
{
    System s = get_new_system();
    s.num_particles = n; s.particle_ids = ids; s.x = x; s.v = v; s.m = m; s.G = G;
    AccelerationParam ap = get_new_acceleration_param();
    ap.method = method; ap.softening_length = eps;
    IntegratorParam ip = get_new_integrator_param();
    ip.integrator = INTEGRATOR_WHFAST; ip.dt = dt; ip.whfast_remove_invalid_particles = remove_invalid;
    OutputParam op = get_new_output_param();
    op.method = OUTPUT_METHOD_DISABLED;
    SimulationStatus st;
    Settings set = get_new_settings();
    bool is_exit = false;
    set.verbose = verbose; set.enable_progress_bar = false; set.is_exit_ptr = &is_exit;
    ErrorStatus es = whfast(&s, &ip, &ap, &op, &st, &set, tf);
    if (es.return_code != GRAV_SUCCESS) { if (es.traceback) free(es.traceback); return -es.return_code - 1000; }
    return s.num_particles;
}

/* the static stage functions, for stage-level pins */
void probe_whfast_c2j(double *jx, double *jv, int n, double *x, double *v, double *m, const double *eta)
{
    System s; s.num_particles = n; s.particle_ids = NULL; s.x = x; s.v = v; s.m = m; s.G = 1.0;
    cartesian_to_jacobi(jx, jv, &s, eta);
}
void probe_whfast_j2c(double *x, double *v, int n, double *m, const double *jx, const double *jv, const double *eta)
{
    System s; s.num_particles = n; s.particle_ids = NULL; s.x = x; s.v = v; s.m = m; s.G = 1.0;
    jacobi_to_cartesian(&s, jx, jv, eta);
}
void probe_stumpff(double c[4], double z) { stumpff_functions(&c[0], &c[1], &c[2], &c[3], z); }
int probe_whfast_drift(int n, int *ids, double *x, double *v, double *m, double G, double *jx, double *jv, double *eta,
                       double dt, int remove_invalid)
{
    System s; s.num_particles = n; s.particle_ids = ids; s.x = x; s.v = v; s.m = m; s.G = G;
    ErrorStatus es = whfast_drift(jx, jv, &s, eta, dt, remove_invalid, 0);
    if (es.return_code != GRAV_SUCCESS) { if (es.traceback) free(es.traceback); return -1; }
    return s.num_particles;
}
