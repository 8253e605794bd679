// Temporary: entries not implemented yet fail loudly.
#include "internal.cuh"
namespace gb {
static int nyi(const char *w) { set_error("%s: not implemented yet", w); return GRAV_B200_ECUDA; }
int whfast_accel(grav_b200_ctx *, const double *, const double *, double, bool) { return nyi("whfast_accel"); }
}
using namespace gb;
extern "C" {
int grav_b200_ctx_leapfrog_begin(grav_b200_ctx *, int, double, double, int) { return nyi("leapfrog_begin"); }
int grav_b200_ctx_leapfrog_steps(grav_b200_ctx *, double, int64_t) { return nyi("leapfrog_steps"); }
int grav_b200_ctx_energy(grav_b200_ctx *, double *) { return nyi("energy"); }
}
