// WHFast interaction-term accelerations in Jacobi coordinates (config 3, Kirkwood-gap runs).
//
// Reference: whfast_acceleration_pairwise / _massless, static in src/integrator_whfast.c:839-957 and
// :959-1264.  Every target particle is independent there (OpenMP over i), so one GPU thread per target
// reproduces the reference's serial per-target arithmetic operation for operation -- IEEE mul/add/div/sqrt
// without FMA contraction, same association -- and the result is bit-identical.  The reference reads aux[3]
// uninitialised in its first inner loop (:861,896; :1026,1063; :1141,1182); we define it as zero, which is what
// the -O3 x86-64 build does in practice.
//
// Softening here is  |r|^3 + eps^3  (not (r^2+eps^2)^{3/2}), :854,874.
#include "internal.cuh"
#include "whfast_device.cuh"

namespace gb {

__global__ void __launch_bounds__(128) whfast_kernel(int n, const double4 *__restrict__ posm, double G,
                                                    const double *__restrict__ jx, const double *__restrict__ eta,
                                                    double eps3, const int *__restrict__ list, int nl,
                                                    const int *__restrict__ rank, double *__restrict__ acc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) whfast_accel_one(i, n, posm, G, jx, eta, eps3, list, nl, rank, acc);
}

int massive_list(grav_b200_ctx *c, int *n_massive);   // direct_sum.cu: fills c->msrc_id (ids) and c->stage_d2 ranks

// resident WHFast (whfast_resident.cu): the caller already holds the massive list of the current particle order
int whfast_accel_with_list(grav_b200_ctx *c, const double *d_jx, const double *d_eta, double eps, const int *list, int nl,
                           const int *rank)
{
    const int n = c->n;
    whfast_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n, c->posm.as<double4>(), c->G, d_jx, d_eta, eps * eps * eps, list, nl, rank,
                                                         c->acc.as<double>());
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

int whfast_accel(grav_b200_ctx *c, const double *d_jx, const double *d_eta, double eps, bool massless)
{
    const int n = c->n;
    const double eps3 = eps * eps * eps;
    const int *list = nullptr, *rank = nullptr;
    int nl = 0;
    if (massless) {
        GB_TRY(massive_list(c, &nl));
        list = c->msrc_id.as<int>();
        rank = c->mrank.as<int>();
    }
    whfast_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n, c->posm.as<double4>(), c->G, d_jx, d_eta, eps3, list, nl, rank,
                                                         c->acc.as<double>());
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

}  // namespace gb
