// Device-resident leapfrog (kick-drift-kick with compensated summation) and the energy diagnostic.
//
// Reference: leapfrog(), src/integrator.c:894-1121 -- initial a(x0) and half kick (:963-982), per step a drift
// (:1009-1018), a force evaluation (:1021) and a full kick (:1030-1039), with Kahan-style error terms for x and v;
// velocities are brought back to the position time level with v -= a dt / 2 only for snapshots and at the end
// (:1048-1056, :1088-1094).  Energy: compute_energy(), src/utils.c:27-59 (unsoftened potential).
//
// The update kernels use the reference's exact operation order without FMA contraction, so with bit-identical
// accelerations (Barnes-Hut in reference mode) the trajectory is bit-identical to the reference's; with the
// direct sum it differs only through the 1e-15-level differences of the accelerations.
// Every rank updates its owned particle range only; the position shards are gathered by the next force call.
#include "internal.cuh"

namespace gb {

// c += a*scale*dt ; v = v0 + c ; c += v0 - v      (reference: v_err += [0.5 *] a * dt; v = tmp + v_err; v_err += tmp - v)
__global__ void __launch_bounds__(256) kick_kernel(double *__restrict__ v, double *__restrict__ vc, const double *__restrict__ a,
                                                  size_t lo, size_t hi, double dt, int half)
{
    const size_t k = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    const double inc = half ? __dmul_rn(__dmul_rn(0.5, a[k]), dt) : __dmul_rn(a[k], dt);
    double c = __dadd_rn(vc[k], inc);
    const double v0 = v[k];
    const double v1 = __dadd_rn(v0, c);
    c = __dadd_rn(c, __dsub_rn(v0, v1));
    v[k] = v1;
    vc[k] = c;
}

__global__ void __launch_bounds__(256) drift_kernel(double4 *__restrict__ posm, double *__restrict__ xc, const double *__restrict__ v,
                                                   int lo, int hi, double dt)
{
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double4 q = posm[i];
    double *pc = xc + 3 * (size_t)i;
    const double *pv = v + 3 * (size_t)i;
    double x0[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double c = __dadd_rn(pc[k], __dmul_rn(pv[k], dt));
        const double x1 = __dadd_rn(x0[k], c);
        c = __dadd_rn(c, __dsub_rn(x0[k], x1));
        pc[k] = c;
        x0[k] = x1;
    }
    q.x = x0[0]; q.y = x0[1]; q.z = x0[2];
    posm[i] = q;
}

// out = v - 0.5*a*dt  (velocities at the position time level)
__global__ void __launch_bounds__(256) vsync_kernel(const double *__restrict__ v, const double *__restrict__ a, double *__restrict__ out,
                                                   size_t lo, size_t hi, double dt)
{
    const size_t k = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    out[k] = __dsub_rn(v[k], __dmul_rn(__dmul_rn(0.5, a[k]), dt));
}


// ---- Euler, Euler-Cromer, RK4 (src/integrator.c:281-454, :456-628, :630-892) ---------------------------------------
// Same conventions as the leapfrog kernels: the reference's formulas and association, IEEE without contraction, one
// thread per owned particle.  mode 0: Euler (x and v advance from the old v and a, :382-395); mode 1: Euler-Cromer
// (v first, x with the NEW v, :557-569).
__global__ void __launch_bounds__(256) euler_kernel(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acc,
                                                   double *__restrict__ xc, double *__restrict__ vc, int lo, int hi, double dt, int cromer)
{
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double4 q = posm[i];
    double x0[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const size_t e = 3 * (size_t)i + k;
        const double v0 = vel[e], a = acc[e];
        double cv = __dadd_rn(vc[e], __dmul_rn(a, dt));
        const double v1 = __dadd_rn(v0, cv);
        cv = __dadd_rn(cv, __dsub_rn(v0, v1));
        double cx = __dadd_rn(xc[e], __dmul_rn(cromer ? v1 : v0, dt));
        const double x1 = __dadd_rn(x0[k], cx);
        cx = __dadd_rn(cx, __dsub_rn(x0[k], x1));
        vel[e] = v1; vc[e] = cv; xc[e] = cx;
        x0[k] = x1;
    }
    q.x = x0[0]; q.y = x0[1]; q.z = x0[2];
    posm[i] = q;
}

// RK4 stage s = 1..3 (:746-819): record xk_s = v and vk_s = a, then x = x_0 + [0.5] xk_s dt, v = v_0 + [0.5] vk_s dt.
// Stage 1 also stores x_0, v_0 (:742-743).
__global__ void __launch_bounds__(256) rk4_stage_kernel(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acc,
                                                       double *__restrict__ x_0, double *__restrict__ v_0, double *__restrict__ xk,
                                                       double *__restrict__ vk, int lo, int hi, double dt, int stage)
{
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double4 q = posm[i];
    double xs[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const size_t e = 3 * (size_t)i + k;
        const double v = vel[e], a = acc[e];
        if (stage == 1) { x_0[e] = xs[k]; v_0[e] = v; }
        xk[e] = v; vk[e] = a;
        const double bx = x_0[e], bv = v_0[e];
        if (stage < 3) {
            xs[k] = __dadd_rn(bx, __dmul_rn(__dmul_rn(0.5, v), dt));
            vel[e] = __dadd_rn(bv, __dmul_rn(__dmul_rn(0.5, a), dt));
        } else {
            xs[k] = __dadd_rn(bx, __dmul_rn(v, dt));
            vel[e] = __dadd_rn(bv, __dmul_rn(a, dt));
        }
    }
    q.x = xs[0]; q.y = xs[1]; q.z = xs[2];
    posm[i] = q;
}

// RK4 update (:837-850): xk4 = v, vk4 = a, weighted sum through the compensated-summation terms.
__global__ void __launch_bounds__(256) rk4_final_kernel(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acc,
                                                       const double *__restrict__ x_0, const double *__restrict__ v_0,
                                                       const double *__restrict__ xk1, const double *__restrict__ xk2,
                                                       const double *__restrict__ xk3, const double *__restrict__ vk1,
                                                       const double *__restrict__ vk2, const double *__restrict__ vk3,
                                                       double *__restrict__ xc, double *__restrict__ vc, int lo, int hi, double dt)
{
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    double4 q = posm[i];
    double xs[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const size_t e = 3 * (size_t)i + k;
        const double xk4 = vel[e], vk4 = acc[e];
        const double sv = __dadd_rn(__dadd_rn(__dadd_rn(vk1[e], __dmul_rn(2.0, vk2[e])), __dmul_rn(2.0, vk3[e])), vk4);
        const double sx = __dadd_rn(__dadd_rn(__dadd_rn(xk1[e], __dmul_rn(2.0, xk2[e])), __dmul_rn(2.0, xk3[e])), xk4);
        double cv = __dadd_rn(vc[e], __ddiv_rn(__dmul_rn(sv, dt), 6.0));
        double cx = __dadd_rn(xc[e], __ddiv_rn(__dmul_rn(sx, dt), 6.0));
        const double bv = v_0[e], bx = x_0[e];
        const double v1 = __dadd_rn(bv, cv), x1 = __dadd_rn(bx, cx);
        cv = __dadd_rn(cv, __dsub_rn(bv, v1));
        cx = __dadd_rn(cx, __dsub_rn(bx, x1));
        vel[e] = v1; vc[e] = cv; xc[e] = cx;
        xs[k] = x1;
    }
    q.x = xs[0]; q.y = xs[1]; q.z = xs[2];
    posm[i] = q;
}

static int run_force(grav_b200_ctx *c)
{
    return grav_b200_ctx_acceleration(c, c->lf_method, c->lf_eps, c->lf_theta, c->lf_leaf);
}

// First force evaluation of a resident run.  Later Barnes-Hut builds are queued without waiting for their sizes, so this
// one waits, rebuilds with more room if the tree did not fit, and leaves 2x headroom over what this tree needed.
static int first_force_checked(grav_b200_ctx *c)
{
    for (;;) {
        GB_TRY(run_force(c));
        if (c->lf_method != GRAV_B200_METHOD_BARNES_HUT) return GRAV_B200_OK;
        GB_CUDA(cudaStreamSynchronize(c->stream));
        const int rc = bh_check(c);
        if (rc == GRAV_B200_ETREE && c->tree.slack < 16) { c->tree.slack *= 2; continue; }
        if (rc != GRAV_B200_OK) return rc;
        if (2LL * c->tree.h_meta->num_expanded > c->tree.ne_cap && c->tree.slack < 16) c->tree.slack *= 2;
        return GRAV_B200_OK;
    }
}

static int kick(grav_b200_ctx *c, double dt, int half)
{
    const size_t lo = 3 * (size_t)c->lo, hi = 3 * (size_t)c->hi;
    if (hi <= lo) return GRAV_B200_OK;
    kick_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, c->stream>>>(c->vel.as<double>(), c->vcomp.as<double>(), c->acc.as<double>(),
                                                                           lo, hi, dt, half);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

// velocities synchronised with the positions: the stored ones while no leapfrog is running, else v - a dt/2
int synced_velocities(grav_b200_ctx *c, double **d_out)
{
    if (!c->lf_ready) { *d_out = c->vel.as<double>(); return GRAV_B200_OK; }
    GB_TRY(c->stage_b.reserve(sizeof(double) * 3 * (size_t)c->n));
    const size_t lo = 3 * (size_t)c->lo, hi = 3 * (size_t)c->hi;
    if (hi > lo) {
        vsync_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, c->stream>>>(c->vel.as<double>(), c->acc.as<double>(),
                                                                                c->stage_b.as<double>(), lo, hi, c->lf_dt);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    *d_out = c->stage_b.as<double>();
    return GRAV_B200_OK;
}

// ---- energy -------------------------------------------------------------------------------------------
// E = sum_i m_i |v_i|^2 / 2  -  G sum_{i<j} m_i m_j / |x_i - x_j|  (compute_energy, src/utils.c:27-59; unsoftened).
// Same tile machinery as the direct sum (direct_sum.cu): TI targets per thread in registers, sources streamed through a
// shared-memory tile, and m_j / r_ij from the MUFU.RSQ64H seed y0 (|e| <~ 2^-20, e = 1 - r2 y0^2) with one correction,
// r2^-1/2 = y0 (1 + e (1/2 + 3/8 e)) -- the dropped term 5/16 e^3 is < 2^-61 -- i.e. 13 FP64-pipe instructions per
// pair instead of the ~45 of sqrt + division.  Every unordered pair is visited ONCE, like the reference's i < j loop: a
// target sums m_j / r_ij over the sources j > i only, so a block of targets streams the tiles from its own onwards.  To
// keep the CTAs equally loaded the unit of work is target block q TOGETHER with its mirror image NB - 1 - q (NT + TI tiles
// whatever q is), cut into a few equal shares of that tile list (energy_kernel below); the ranks of a multi-GPU run take
// the work items in turn.  The kinetic term is summed over the rank's own targets
// (velocities are sharded).  Block partial sums are written out and added in block order by one thread: deterministic.
// Coincident particles give a non-finite energy like the reference's 1/0.

// tiles [tile_lo, tile_hi) (256 sources each, absolute indices) against target block blk
template <int TI>
__device__ __forceinline__ double energy_block_potential(const double4 *__restrict__ posm, int n, int blk, int tile_lo, int tile_hi, double4 *tile)
{
    double xi[TI], yi[TI], zi[TI], mi[TI], pot[TI];
    int ii[TI];
#pragma unroll
    for (int t = 0; t < TI; t++) {
        ii[t] = (blk * TI + t) * 256 + threadIdx.x;
        const double4 me = (ii[t] < n) ? posm[ii[t]] : make_double4(0.0, 0.0, 0.0, 0.0);
        xi[t] = me.x; yi[t] = me.y; zi[t] = me.z; mi[t] = me.w;
        pot[t] = 0.0;
    }
    for (int jt = tile_lo; jt < tile_hi; jt++) {
        const int j0 = jt * 256;
        __syncthreads();
        tile[threadIdx.x] = posm[j0 + threadIdx.x];
        __syncthreads();
        const int jcount = min(256, n - j0);          // padding sources are never visited
#pragma unroll 2
        for (int j = 0; j < jcount; j++) {
            const double4 q = tile[j];
#pragma unroll
            for (int t = 0; t < TI; t++) {
                const double dx = q.x - xi[t], dy = q.y - yi[t], dz = q.z - zi[t];
                const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const double y = rsqrt_seed(r2);
                const double e = fma(-r2, y * y, 1.0);
                const double my = q.w * y;
                double w = fma(my, e * fma(0.375, e, 0.5), my);     // m_j / r
                if (j0 + j <= ii[t]) w = 0.0;                         // only the partners behind the target (and not itself)
                pot[t] += w;
            }
        }
    }
    double e = 0.0;
#pragma unroll
    for (int t = 0; t < TI; t++) {
        if (ii[t] < n && pot[t] != pot[t]) {
            // r = 0 (another particle at the same point) turns the seed into inf and the correction into NaN; the reference
            // divides by zero there (m / 0 = inf, or NaN for a massless partner): repeat this target's share with the division
            double p = 0.0;
            for (int j = max(ii[t] + 1, tile_lo * 256); j < min(n, tile_hi * 256); j++) {
                const double4 q = posm[j];
                const double dx = q.x - xi[t], dy = q.y - yi[t], dz = q.z - zi[t];
                p += q.w / sqrt(fma(dz, dz, fma(dy, dy, dx * dx)));
            }
            pot[t] = p;
        }
        if (ii[t] < n) e -= mi[t] * pot[t];
    }
    return e;      // - sum over the block's targets of m_i sum_{j>i, j in the tiles} m_j / r_ij   (G applied by the caller)
}

// Work item = (mirrored pair q of target blocks, split s of `splits`): the NT - b TI tiles of block q followed by those of block
// NB - 1 - q form one list of ~NT + TI tiles whatever q is; the item takes the s-th share of that list.  The splits only serve
// the balance: 256 equal pairs on 148 SMs would run as long as 296.
template <int TI>
__global__ void __launch_bounds__(256) energy_kernel(const double4 *__restrict__ posm, const double *__restrict__ v, int n, int n_pad,
                                                    int lo, int hi, int rank, int world, int splits, double G,
                                                    double *__restrict__ block_out)
{
    __shared__ double4 tile[256];
    __shared__ double red[256];
    const int NB = (n + 256 * TI - 1) / (256 * TI), NT = n_pad / 256;
    const int item = blockIdx.x * world + rank;
    const int q = item / splits, s = item - q * splits;
    double e = 0.0;
    if (q < (NB + 1) / 2) {
        const int b0 = q, b1 = NB - 1 - q;
        const int T0 = NT - b0 * TI, T1 = (b1 != b0) ? NT - b1 * TI : 0;
        const int a = (int)(((long long)s * (T0 + T1)) / splits), z = (int)(((long long)(s + 1) * (T0 + T1)) / splits);
        if (a < T0) e = energy_block_potential<TI>(posm, n, b0, b0 * TI + a, b0 * TI + min(z, T0), tile);
        if (z > T0) e += energy_block_potential<TI>(posm, n, b1, b1 * TI + max(a - T0, 0), b1 * TI + (z - T0), tile);
        e *= G;
    }
    // kinetic energy of the rank's own targets, spread over the CTAs
    for (int i = lo + blockIdx.x * 256 + threadIdx.x; i < hi; i += gridDim.x * 256) {
        const double vx = v[3 * (size_t)i], vy = v[3 * (size_t)i + 1], vz = v[3 * (size_t)i + 2];
        e += 0.5 * posm[i].w * (vx * vx + vy * vy + vz * vz);
    }
    red[threadIdx.x] = e;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_out[blockIdx.x] = red[0];
}

__global__ void sum_blocks_kernel(const double *__restrict__ in, int count, double *__restrict__ out)
{
    double s = 0.0;
    for (int k = 0; k < count; k++) s += in[k];
    out[0] = s;
}

}  // namespace gb

using namespace gb;

extern "C" {

int grav_b200_ctx_leapfrog_begin(grav_b200_ctx *c, int method, double eps, double theta, int max_leaf, double dt)
{
    GB_TEAM(c, grav_b200_ctx_leapfrog_begin(r_, method, eps, theta, max_leaf, dt));
    if (!c || c->n < 1) { set_error("context has no system"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    const size_t b3 = sizeof(double) * 3 * (size_t)c->n;
    GB_TRY(c->xcomp.reserve(b3));
    GB_TRY(c->vcomp.reserve(b3));
    GB_CUDA(cudaMemsetAsync(c->xcomp.p, 0, b3, c->stream));
    GB_CUDA(cudaMemsetAsync(c->vcomp.p, 0, b3, c->stream));
    c->lf_method = method; c->lf_eps = eps; c->lf_theta = theta; c->lf_leaf = max_leaf;
    c->lf_ready = false;
    GB_TRY(first_force_checked(c));     // a(x0), src/integrator.c:963
    GB_TRY(kick(c, dt, 1));             // v_{1/2}, :973-982
    c->lf_dt = dt;
    c->lf_ready = true;
    return GRAV_B200_OK;
}

int grav_b200_ctx_leapfrog_steps(grav_b200_ctx *c, double dt, int64_t num_steps)
{
    GB_TEAM(c, grav_b200_ctx_leapfrog_steps(r_, dt, num_steps));
    if (!c || !c->lf_ready) { set_error("leapfrog_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    for (int64_t s = 0; s < num_steps; s++) {
        const int cnt = c->hi - c->lo;
        if (cnt > 0) {
            drift_kernel<<<(cnt + 255) / 256, 256, 0, c->stream>>>(c->posm.as<double4>(), c->xcomp.as<double>(), c->vel.as<double>(),
                                                                  c->lo, c->hi, dt);   // :1009-1018
            GB_LAUNCH_CHECK();
            count_launch();
        }
        if (c->world > 1) c->posm_gathered = false;
        GB_TRY(run_force(c));           // :1021
        GB_TRY(kick(c, dt, 0));         // :1030-1039
        c->lf_dt = dt;
    }
    return GRAV_B200_OK;
}

int grav_b200_ctx_leapfrog_end(grav_b200_ctx *c)
{
    GB_TEAM(c, grav_b200_ctx_leapfrog_end(r_));
    if (!c || !c->lf_ready) { set_error("leapfrog_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    const size_t lo = 3 * (size_t)c->lo, hi = 3 * (size_t)c->hi;
    if (hi > lo) {   // v -= a dt / 2 in place, :1088-1094
        vsync_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, c->stream>>>(c->vel.as<double>(), c->acc.as<double>(), c->vel.as<double>(),
                                                                                lo, hi, c->lf_dt);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    c->lf_ready = false;
    return GRAV_B200_OK;
}


int grav_b200_ctx_fixed_begin(grav_b200_ctx *c, int integrator, int method, double eps, double theta, int max_leaf)
{
    GB_TEAM(c, grav_b200_ctx_fixed_begin(r_, integrator, method, eps, theta, max_leaf));
    if (!c || c->n < 1) { set_error("context has no system"); return GRAV_B200_EINVAL; }
    if (integrator != GRAV_B200_INTEGRATOR_EULER && integrator != GRAV_B200_INTEGRATOR_EULER_CROMER &&
        integrator != GRAV_B200_INTEGRATOR_RK4) {
        set_error("Unknown fixed-step integrator. Got: %d", integrator);
        return GRAV_B200_EINVAL;
    }
    GB_CUDA(cudaSetDevice(c->device));
    const size_t b3 = sizeof(double) * 3 * (size_t)c->n;
    GB_TRY(c->xcomp.reserve(b3));
    GB_TRY(c->vcomp.reserve(b3));
    GB_CUDA(cudaMemsetAsync(c->xcomp.p, 0, b3, c->stream));     // calloc'ed error terms (:313-314, :488-489, :672-673)
    GB_CUDA(cudaMemsetAsync(c->vcomp.p, 0, b3, c->stream));
    if (integrator == GRAV_B200_INTEGRATOR_RK4) GB_TRY(c->rk_buf.reserve(8 * b3));   // x_0, v_0, xk1-3, vk1-3
    c->lf_method = method; c->lf_eps = eps; c->lf_theta = theta; c->lf_leaf = max_leaf;
    c->lf_ready = false;
    c->fixed_integrator = integrator;
    if (method == GRAV_B200_METHOD_BARNES_HUT) {   // size the tree buffers on a throw-away build (see first_force_checked)
        GB_TRY(bh_build_checked(c, max_leaf == -1 ? 1 : max_leaf, nullptr, -1.0));
        if (2LL * c->tree.h_meta->num_expanded > c->tree.ne_cap && c->tree.slack < 16) c->tree.slack *= 2;
    }
    return GRAV_B200_OK;
}

int grav_b200_ctx_fixed_steps(grav_b200_ctx *c, double dt, int64_t num_steps)
{
    GB_TEAM(c, grav_b200_ctx_fixed_steps(r_, dt, num_steps));
    if (!c || !c->fixed_integrator) { set_error("fixed_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    const int cnt = c->hi - c->lo;
    const unsigned blocks = (unsigned)((cnt + 255) / 256);
    double4 *posm = c->posm.as<double4>();
    double *vel = c->vel.as<double>(), *acc = c->acc.as<double>(), *xc = c->xcomp.as<double>(), *vc = c->vcomp.as<double>();
    const size_t n3 = 3 * (size_t)c->n;
    for (int64_t s = 0; s < num_steps; s++) {
        if (c->fixed_integrator != GRAV_B200_INTEGRATOR_RK4) {
            GB_TRY(run_force(c));
            if (cnt > 0) {
                euler_kernel<<<blocks, 256, 0, c->stream>>>(posm, vel, acc, xc, vc, c->lo, c->hi, dt,
                                                           c->fixed_integrator == GRAV_B200_INTEGRATOR_EULER_CROMER);
                GB_LAUNCH_CHECK();
                count_launch();
            }
            if (c->world > 1) c->posm_gathered = false;
            continue;
        }
        double *r = c->rk_buf.as<double>();
        double *x_0 = r, *v_0 = r + n3, *xk = r + 2 * n3, *vk = r + 5 * n3;     // xk[3][n3], vk[3][n3]
        for (int st = 1; st <= 3; st++) {
            GB_TRY(run_force(c));
            if (cnt > 0) {
                rk4_stage_kernel<<<blocks, 256, 0, c->stream>>>(posm, vel, acc, x_0, v_0, xk + (st - 1) * n3, vk + (st - 1) * n3,
                                                               c->lo, c->hi, dt, st);
                GB_LAUNCH_CHECK();
                count_launch();
            }
            if (c->world > 1) c->posm_gathered = false;
        }
        GB_TRY(run_force(c));
        if (cnt > 0) {
            rk4_final_kernel<<<blocks, 256, 0, c->stream>>>(posm, vel, acc, x_0, v_0, xk, xk + n3, xk + 2 * n3, vk, vk + n3, vk + 2 * n3,
                                                           xc, vc, c->lo, c->hi, dt);
            GB_LAUNCH_CHECK();
            count_launch();
        }
        if (c->world > 1) c->posm_gathered = false;
    }
    return GRAV_B200_OK;
}

int grav_b200_ctx_energy(grav_b200_ctx *c, double *energy)
{
    if (team_active(c)) {   // every member reduces to the same total; only the leader writes the caller's variable
        return team_run(c, [=](grav_b200_ctx *r_) -> int { double other; return grav_b200_ctx_energy(r_, r_->rank == 0 ? energy : &other); });
    }
    if (!c || !energy || c->n < 1) { set_error("context has no system / NULL pointer"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    if (c->world > 1 && !c->posm_gathered) { GB_TRY(comm_allgather_posm(c)); c->posm_gathered = true; }
    double *d_v;
    GB_TRY(synced_velocities(c, &d_v));
    // every unordered pair once: a CTA takes a block of targets and its mirror image, the ranks take the CTAs in turn.
    // Four targets per thread once that still leaves two CTAs per SM; one target per thread below (small systems need the CTAs)
    auto pairs_for = [&](int ti) { const int nb = (c->n + 256 * ti - 1) / (256 * ti); return (nb + 1) / 2; };
    const bool ti4 = (pairs_for(4) + c->world - 1) / c->world >= 2 * c->sm_count;
    const int pairs = ti4 ? pairs_for(4) : pairs_for(1);
    // split the pairs' tile lists until a rank has at least ~6 items per SM (rounding loss of the last wave < 15 %)
    int splits = 1;
    const int NT = c->n_pad / 256;
    while (splits < 16 && splits * 2 <= NT && (long long)pairs * splits < 6LL * c->sm_count * c->world) splits *= 2;
    const int blocks = (int)(((long long)pairs * splits + c->world - 1) / c->world);
    GB_TRY(c->misc.reserve(sizeof(double) * ((size_t)blocks + 2)));
    double *part = c->misc.as<double>();
    if (blocks > 0) {
        if (ti4) energy_kernel<4><<<blocks, 256, 0, c->stream>>>(c->posm.as<double4>(), d_v, c->n, c->n_pad, c->lo, c->hi, c->rank, c->world, splits, c->G, part + 1);
        else energy_kernel<1><<<blocks, 256, 0, c->stream>>>(c->posm.as<double4>(), d_v, c->n, c->n_pad, c->lo, c->hi, c->rank, c->world, splits, c->G, part + 1);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    sum_blocks_kernel<<<1, 1, 0, c->stream>>>(part + 1, blocks, part);
    GB_LAUNCH_CHECK();
    count_launch();
    GB_TRY(comm_allreduce_sum(c, part, 1));
    GB_CUDA(cudaMemcpyAsync(energy, part, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return bh_check(c);
}

}  // extern "C"
