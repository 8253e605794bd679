// Device-resident WHFast: the whole Wisdom-Holman step stays in HBM (config 3,
// Sun + planets + 1e5 massless asteroids).
//
// Reference: whfast(), src/integrator_whfast.c:200-407.  Per step it (1) sorts the particles by Jacobi distance from
// particle id 0 (system_sort_by_distance, src/system.c:1210-1335, a stable sort), (2) recomputes eta = prefix sums of
// the masses (:1266-1279), (3) Kepler-drifts every particle but the first (whfast_drift :424-680, Newton iteration on
// the universal-variable Kepler equation with Stumpff series :774-815, optional removal of particles whose solve
// failed :551,607-671), (4) converts Jacobi -> Cartesian (:726-772), (5) evaluates the interaction acceleration
// (whfast.cu) and (6) kicks the Jacobi velocities (:409-422).
//
// Everything is IEEE mul/add/div/sqrt in the reference's association without FMA contraction, so the state after any
// number of steps is bit-identical to the reference's (tests/test_whfast_resident_gpu.py).  Two of the pieces are
// serial recurrences on the CPU; here they are split into a tiny serial "skeleton" over the MASSIVE particles and a
// parallel fill for the massless ones:
//   * eta[i] = eta[i-1] + m[i]: adding a zero mass is exact, so eta only changes at massive particles.
//   * cartesian_to_jacobi: the running centre-of-mass term c <- c*(1 + m_i/eta) + m_i*J_i is unchanged by m_i = 0.
//   * jacobi_to_cartesian: walking down a run of massless particles the reference applies t <- fl(c/eta),
//     c <- fl(eta*t) once per particle, with the same eta.  That map is monotone, so the sequence t_0, t_1, ... is
//     monotone and in practice reaches a fixed point after one or two particles; the skeleton iterates it until two
//     consecutive values are bit-identical, stores the short transient in a table and the fill kernel looks the value
//     of every particle up by its position in the run (an explicit per-particle fallback covers a transient longer
//     than the table).
//   (Signed zeros: 0*J is taken as +0 in these shortcuts; a centre-of-mass component that is exactly -0.0 can come out
//   as +0.0.  Values never differ.)
// Removal of invalid particles needs the host (the particle count changes), but checking for it every step would
// serialise launch and execution.  Steps therefore run optimistically in batches from a checkpoint of the Jacobi
// state; the drift kernel records the first step of the batch that flagged a particle, and only then the batch is
// replayed up to that step, which is finished "carefully" (flags -> stable compaction -> new eta), exactly like
// the serial reference build does (ascending index order, src/system.c:444-528).
//
// A step of the common case (massless method, <= 64 massive bodies, <= 131072 particles) is seven kernels, two steps
// per CUDA graph launch:
//   wh_dist_kernel -> sort_small_all_kernel (bh_sort.cu) -> wh_gather_kernel -> wh_skel_kernel -> wh_drift_kernel ->
//   wh_j2c_skel_kernel (recurrence, then massive targets and gap pair sums, one warp per item) -> wh_tail_kernel
// Larger systems, more massive bodies and the pairwise method fall back to separate flag/scan/list, fill,
// acceleration and kick kernels (tests force those paths through GRAV_B200_WHFAST_SKEL_MAX_K / _PAIR_MAX_K).
#include "internal.cuh"
#include "whfast_device.cuh"

namespace gb {

int radix_pass(grav_b200_ctx *c, const long long *kin, const int *vin, long long *kout, int *vout, int n, int shift);   // bh_sort.cu
int radix_sort_buffers(grav_b200_ctx *c, long long *ka, int *pa, long long *kb, int *pb, int n);                         // bh_sort.cu
int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp);                                // scan.cu
int whfast_accel_with_list(grav_b200_ctx *c, const double *d_jx, const double *d_eta, double eps, const int *list, int nl,
                           const int *rank);                                                                              // whfast.cu

constexpr int WH_KMAX = 16;          // table entries per (massless run, component)
constexpr int WH_BATCH = 32;         // optimistic steps between removal checks
constexpr int WH_NO_BAD = 0x7fffffff;

struct WhfastState {
    DevBuf jx[2], jv[2], m[2], ids[2];       // ping-pong: [cur] is live, [cur^1] is the gather target
    int cur = 0;
    DevBuf eta, etaM, keys[2], perm[2];
    DevBuf flag, rank, list, ulist;          // massive flags, exclusive scan (n+1), massive indices (sorted / as appended)
    DevBuf tab, tabinfo, cstate, texp;       // jacobi<->cartesian skeleton outputs
    DevBuf bad, status;                      // per-particle removal flags; status[0] = primary slot, [1] = first bad step,
                                             // [2] = removal count, [3] = error bits
    DevBuf ck_jx, ck_jv, ck_m, ck_ids;       // checkpoint of the Jacobi state at the start of a batch
    int *h_status = nullptr;                 // pinned
    int K = 0;                               // massive particles
    int skel_max_k = 1024, pair_max_k = 64;  // path thresholds (WH_SKEL_MAX_K / WH_PAIR_MAX_K; lowered by tests through the environment)
    // two consecutive steps captured as one CUDA graph (the ping-pong buffers are back in place after two), per
    // starting parity; valid for one (n, K, dt)
    cudaGraphExec_t pair_exec[2] = {nullptr, nullptr};
    int pair_n[2] = {0, 0}, pair_K[2] = {0, 0}, pair_launches[2] = {0, 0};
    uint64_t pair_gen[2] = {0, 0};           // g_alloc_generation at capture: a buffer that moved since invalidates the graph
    double pair_dt[2] = {0.0, 0.0};
    int method = 0;
    double eps = 0.0;
    bool remove_invalid = false;
    int verbose = 0;                         // settings->verbose of the caller (GRAV_VERBOSITY_*): >= 3 prints the removal message
    bool ready = false;
    double last_dt = 0.0;

    double *JX() { return jx[cur].as<double>(); }
    double *JV() { return jv[cur].as<double>(); }
    double *M() { return m[cur].as<double>(); }
    int *IDS() { return ids[cur].as<int>(); }
};

void whfast_state_free(grav_b200_ctx *c)
{
    WhfastState *w = (WhfastState *)c->wh;
    if (!w) return;
    DevBuf *bufs[] = {&w->jx[0], &w->jx[1], &w->jv[0], &w->jv[1], &w->m[0], &w->m[1], &w->ids[0], &w->ids[1], &w->eta, &w->etaM,
                      &w->keys[0], &w->keys[1], &w->perm[0], &w->perm[1], &w->flag, &w->rank, &w->list, &w->ulist, &w->tab, &w->tabinfo,
                      &w->cstate, &w->texp, &w->bad, &w->status, &w->ck_jx, &w->ck_jv, &w->ck_m, &w->ck_ids};
    for (DevBuf *b : bufs) b->release();
    if (w->h_status) cudaFreeHost(w->h_status);
    for (int k = 0; k < 2; k++) if (w->pair_exec[k]) cudaGraphExecDestroy(w->pair_exec[k]);
    delete w;
    c->wh = nullptr;
}

// ---- small device helpers (IEEE, no contraction) ------------------------------------------------------------------
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double nrm3(double x, double y, double z)
{
    return __dsqrt_rn(add(add(mul(x, x), mul(y, y)), mul(z, z)));
}

// ---- sort by distance (src/system.c:1232-1280) ---------------------------------------------------------------------
__global__ void wh_primary_kernel(int n, const int *__restrict__ ids, int primary_id, int *__restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ids[i] == primary_id) atomicMin(&status[0], i);    // first index holding the id (:1234-1251)
}

// key = bit pattern of the distance (non-negative doubles order like their bit patterns); the primary gets 0 (:1277)
// slot: status word holding the primary's index; reset_slot (>= 0): word to re-arm for the gather kernel of this step
__global__ void __launch_bounds__(256) wh_dist_kernel(int n, const double *__restrict__ pos, int *__restrict__ status, int slot,
                                                      int reset_slot, long long *__restrict__ keys, int *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = status[slot];
    if (i == 0 && reset_slot >= 0) status[reset_slot] = 0x7f7f7f7f;
    if (p < 0 || p >= n) {          // id not present: the host turns this into GRAV_VALUE_ERROR
        if (i == 0) atomicOr(&status[3], 1);
        p = 0;
    }
    const double d = nrm3(sub(pos[3 * (size_t)i], pos[3 * (size_t)p]), sub(pos[3 * (size_t)i + 1], pos[3 * (size_t)p + 1]),
                          sub(pos[3 * (size_t)i + 2], pos[3 * (size_t)p + 2]));
    keys[i] = (i == p) ? 0LL : __double_as_longlong(d);
    perm[i] = i;
}

__global__ void __launch_bounds__(256) wh_gather_kernel(int n, const int *__restrict__ perm, const double *__restrict__ a_in,
                                                        const double *__restrict__ b_in, const double *__restrict__ m_in,
                                                        const int *__restrict__ ids_in, double *__restrict__ a_out,
                                                        double *__restrict__ b_out, double *__restrict__ m_out,
                                                        int *__restrict__ ids_out, int *__restrict__ status, int next_slot,
                                                        int *__restrict__ ulist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t s = (size_t)perm[i];
    if (ids_in[s] == 0) atomicMin(&status[next_slot], i);        // where the next sort finds particle id 0 (:1234-1251)
    if (m_in[s] != 0.0) ulist[atomicAdd(&status[6], 1)] = i;     // massive particles, order fixed by wh_skel_kernel
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a_out[3 * (size_t)i + k] = a_in[3 * s + k];
        b_out[3 * (size_t)i + k] = b_in[3 * s + k];
    }
    m_out[i] = m_in[s];
    ids_out[i] = ids_in[s];
}

// ---- massive list ----------------------------------------------------------------------------------------------------
__global__ void wh_flag_kernel(int n, const double *__restrict__ m, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = m[i] != 0.0;      // the reference's test (src/integrator_whfast.c:985, src/acceleration.c:262)
    if (i == n) flag[i] = 0;
}
__global__ void wh_list_kernel(int n, const int *__restrict__ flag, const int *__restrict__ rank, int *__restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[rank[i]] = i;
}

// eta over the massive particles only (:1266-1279): etaM[k] = eta[list[k]]
__global__ void wh_eta_skel_kernel(int K, const int *__restrict__ list, const double *__restrict__ m, double *__restrict__ etaM,
                                   int *__restrict__ status)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    status[5] += 1;          // step tag for the drift kernel that follows (kernel parameters are frozen inside a graph)
    status[6] = 0;           // massive indices appended by the gather kernel are not used on this path
    double e = 0.0;
    for (int k = 0; k < K; k++) {
        const double mk = m[list[k]];
        e = (k == 0) ? mk : add(e, mk);
        etaM[k] = e;
    }
}
// Small-K path: the gather kernel appended the massive indices in arbitrary order; one CTA ranks them (K <= 1024),
// re-arms the append counter and runs the eta recurrence.
constexpr int WH_SKEL_MAX_K = 1024;
__global__ void __launch_bounds__(256) wh_skel_kernel(int K, const int *__restrict__ ulist, int *__restrict__ list,
                                                      const double *__restrict__ m, double *__restrict__ etaM, int *__restrict__ status)
{
    __shared__ int s_u[WH_SKEL_MAX_K];
    const int cnt = status[6];
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_u[k] = (k < cnt) ? ulist[k] : 0x7fffffff;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int mine = s_u[k];
        int r = 0;
        for (int j = 0; j < K; j++) r += (s_u[j] < mine);        // indices are distinct
        list[r] = mine;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (cnt != K) atomicOr(&status[3], 2);                   // the massive count changed behind the host's back
        status[6] = 0;
        status[5] += 1;
        double e = 0.0;
        for (int k = 0; k < K; k++) {
            const double mk = m[list[k]];
            e = (k == 0) ? mk : add(e, mk);
            etaM[k] = e;
        }
    }
}
// number of massive particles with index < i (list sorted ascending)
__device__ __forceinline__ int massive_before(const int *__restrict__ list, int K, int i)
{
    int lo = 0, hi = K;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < i) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// eta at particle i, given cnt = number of massive particles with index <= i
__device__ __forceinline__ double eta_at(const double *__restrict__ etaM, int cnt) { return cnt > 0 ? etaM[cnt - 1] : 0.0; }

__global__ void wh_eta_fill_kernel(int n, const int *__restrict__ rank, const double *__restrict__ etaM, double *__restrict__ eta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) eta[i] = eta_at(etaM, rank[i + 1]);
}

// ---- Kepler drift (:455-600) ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stumpff(double z, double &c0, double &c1, double &c2, double &c3)   // :774-815
{
    int n = 0;
    while (fabs(z) > 0.1) { z = dvd(z, 4.0); n++; }
    double t = sub(1.0, dvd(z, 210.0));
    t = sub(1.0, mul(dvd(z, 156.0), t));
    t = sub(1.0, mul(dvd(z, 110.0), t));
    t = sub(1.0, mul(dvd(z, 72.0), t));
    t = sub(1.0, mul(dvd(z, 42.0), t));
    t = sub(1.0, mul(dvd(z, 20.0), t));
    double k3 = dvd(t, 6.0);
    t = sub(1.0, dvd(z, 182.0));
    t = sub(1.0, mul(dvd(z, 132.0), t));
    t = sub(1.0, mul(dvd(z, 90.0), t));
    t = sub(1.0, mul(dvd(z, 56.0), t));
    t = sub(1.0, mul(dvd(z, 30.0), t));
    t = sub(1.0, mul(dvd(z, 12.0), t));
    double k2 = dvd(t, 2.0);
    double k1 = sub(1.0, mul(z, k3));
    double k0 = sub(1.0, mul(z, k2));
    for (; n > 0; n--) {
        k3 = dvd(add(k2, mul(k0, k3)), 4.0);
        k2 = dvd(mul(k1, k1), 2.0);
        k1 = mul(k0, k1);
        k0 = sub(mul(mul(2.0, k0), k0), 1.0);
    }
    c0 = k0; c1 = k1; c2 = k2; c3 = k3;
}

__global__ void __launch_bounds__(128) wh_drift_kernel(int n, double *__restrict__ jx, double *__restrict__ jv,
                                                       const double *__restrict__ m, const int *__restrict__ rank,
                                                       const double *__restrict__ etaM, double *__restrict__ eta, double G,
                                                       double dt, int remove_invalid, char *__restrict__ bad,
                                                       int *__restrict__ status, const int *__restrict__ list, int K,
                                                       int *__restrict__ rank_out, int *__restrict__ flag_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r0, r1;     // massive particles with index < i, <= i
    if (list) {     // small-K path: rank / flag arrays are produced here for the kernels that follow
        const int f = m[i] != 0.0;
        r0 = massive_before(list, K, i);
        r1 = r0 + f;
        rank_out[i] = r0;
        flag_out[i] = f;
        if (i == n - 1) { rank_out[n] = r1; flag_out[n] = 0; }
    } else {
        r0 = rank[i]; r1 = rank[i + 1];
    }
    const double eta_i = eta_at(etaM, r1);
    eta[i] = eta_i;
    bad[i] = 0;
    if (i == 0) return;
    const double eta_im1 = eta_at(etaM, r0);
    const double gm = dvd(mul(mul(G, m[0]), eta_i), eta_im1);                       // :459
    const double x0 = jx[3 * (size_t)i], x1 = jx[3 * (size_t)i + 1], x2 = jx[3 * (size_t)i + 2];
    const double v0 = jv[3 * (size_t)i], v1 = jv[3 * (size_t)i + 1], v2 = jv[3 * (size_t)i + 2];
    const double xn = nrm3(x0, x1, x2), vn = nrm3(v0, v1, v2);
    const double rv = dvd(add(add(mul(x0, v0), mul(x1, v1)), mul(x2, v2)), xn);     // :467
    const double alpha = sub(dvd(mul(2.0, gm), xn), mul(vn, vn));                   // :469
    const double xr = mul(xn, rv);
    double s = dvd(dt, xn);                                                         // :474
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
    bool converged = false, z_bad = false;
    for (int it = 0; it < 500; it++) {                                              // :484-522
        const double ss = mul(s, s);
        const double z = mul(alpha, ss);
        if (!isfinite(z)) { z_bad = true; break; }
        stumpff(z, c0, c1, c2, c3);
        const double F = sub(add(add(mul(mul(xn, s), c1), mul(mul(xr, ss), c2)), mul(mul(gm, mul(ss, s)), c3)), dt);
        const double dF = add(add(mul(xn, c0), mul(mul(xr, s), c1)), mul(mul(gm, ss), c2));
        const double ds = dvd(-F, dF);
        s = add(s, ds);
        if (fabs(ds) < 1e-12) { converged = true; break; }
    }
    const double ss = mul(s, s);
    const double r = add(add(mul(xn, c0), mul(mul(xr, s), c1)), mul(mul(gm, ss), c2));                  // :528
    if (!converged) {
        const double err = dvd(sub(add(add(mul(mul(xn, s), c1), mul(mul(xr, ss), c2)), mul(mul(gm, mul(ss, s)), c3)), dt), r);
        if ((err > 1e-5 || z_bad) && remove_invalid) {                                                  // :551
            bad[i] = 1;
            atomicMin(&status[1], status[5] - 1);
            atomicAdd(&status[2], 1);
        }
    }
    const double gss2 = mul(mul(gm, ss), c2);
    const double f = sub(1.0, dvd(gss2, xn));                                       // :567
    const double g = sub(dt, mul(mul(gm, mul(ss, s)), c3));                         // :568
    const double df = dvd(mul(mul(-gm, s), c1), mul(r, xn));                        // :570
    const double dg = sub(1.0, dvd(gss2, r));                                       // :571
    jx[3 * (size_t)i + 0] = add(mul(f, x0), mul(g, v0));
    jx[3 * (size_t)i + 1] = add(mul(f, x1), mul(g, v1));
    jx[3 * (size_t)i + 2] = add(mul(f, x2), mul(g, v2));
    jv[3 * (size_t)i + 0] = add(mul(df, x0), mul(dg, v0));
    jv[3 * (size_t)i + 1] = add(mul(df, x1), mul(dg, v1));
    jv[3 * (size_t)i + 2] = add(mul(df, x2), mul(dg, v2));
}

// ---- stable compaction after a removal (src/system.c:444-528, :530-...) ---------------------------------------------------
__global__ void wh_keep_flag_kernel(int n, const char *__restrict__ bad, int *__restrict__ keep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = bad[i] ? 0 : 1;
    if (i == n) keep[i] = 0;
}
__global__ void wh_compact_kernel(int n, const int *__restrict__ keep, const int *__restrict__ pos, const double *__restrict__ a_in,
                                  const double *__restrict__ b_in, const double *__restrict__ m_in, const int *__restrict__ ids_in,
                                  double *__restrict__ a_out, double *__restrict__ b_out, double *__restrict__ m_out,
                                  int *__restrict__ ids_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const size_t d = (size_t)pos[i];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a_out[3 * d + k] = a_in[3 * (size_t)i + k];
        b_out[3 * d + k] = b_in[3 * (size_t)i + k];
    }
    m_out[d] = m_in[i];
    ids_out[d] = ids_in[i];
}

// ---- Jacobi <-> Cartesian ----------------------------------------------------------------------------------------------
// component c of lane: 0..2 position, 3..5 velocity
__device__ __forceinline__ double ld_cart(const double4 *posm, const double *vel, int i, int c)
{
    return c < 3 ? reinterpret_cast<const double *>(posm)[4 * (size_t)i + c] : vel[3 * (size_t)i + (c - 3)];
}
__device__ __forceinline__ void st_cart(double4 *posm, double *vel, int i, int c, double val)
{
    if (c < 3) reinterpret_cast<double *>(posm)[4 * (size_t)i + c] = val;
    else vel[3 * (size_t)i + (c - 3)] = val;
}
__device__ __forceinline__ double ld_jac(const double *jx, const double *jv, int i, int c)
{
    return c < 3 ? jx[3 * (size_t)i + c] : jv[3 * (size_t)i + (c - 3)];
}
__device__ __forceinline__ void st_jac(double *jx, double *jv, int i, int c, double val)
{
    if (c < 3) jx[3 * (size_t)i + c] = val;
    else jv[3 * (size_t)i + (c - 3)] = val;
}

// cartesian_to_jacobi (:682-724), skeleton: cstate[q][c] = running centre-of-mass term after the first q massive
// particles; massive particles get their Jacobi coordinates here, particle 0 at the end.
__global__ void wh_c2j_skel_kernel(int n, int K, const int *__restrict__ list, const double *__restrict__ m,
                                   const double *__restrict__ etaM, const double4 *__restrict__ posm, const double *__restrict__ vel,
                                   double *__restrict__ jx, double *__restrict__ jv, double *__restrict__ cstate)
{
    const int c = threadIdx.x;
    if (blockIdx.x != 0 || c >= 6) return;
    double cm = mul(m[0], ld_cart(posm, vel, 0, c));                       // :697-703
    cstate[c] = cm;
    for (int q = 1; q <= K; q++) {
        const int mi = list[q - 1];
        if (mi != 0) {
            const double e_prev = eta_at(etaM, q - 1);
            const double mk = m[mi];
            const double j = sub(ld_cart(posm, vel, mi, c), dvd(cm, e_prev));      // :709
            st_jac(jx, jv, mi, c, j);
            cm = add(mul(cm, add(1.0, dvd(mk, e_prev))), mul(mk, j));              // :712
        }
        cstate[6 * q + c] = cm;
    }
    st_jac(jx, jv, 0, c, dvd(cm, eta_at(etaM, K)));                          // :717-723
}
__global__ void __launch_bounds__(256) wh_c2j_fill_kernel(int n, const int *__restrict__ flag, const int *__restrict__ rank,
                                                          const double *__restrict__ etaM, const double *__restrict__ cstate,
                                                          const double4 *__restrict__ posm, const double *__restrict__ vel,
                                                          double *__restrict__ jx, double *__restrict__ jv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= n || flag[i]) return;
    const int q = rank[i];
    const double e_prev = eta_at(etaM, q);
#pragma unroll
    for (int c = 0; c < 6; c++) st_jac(jx, jv, i, c, sub(ld_cart(posm, vel, i, c), dvd(cstate[6 * q + c], e_prev)));
}

// jacobi_to_cartesian (:726-772), skeleton.  Run r (r = 0..K) is the stretch of massless particles that follows massive
// list entry r-1 (r = 0: before the first massive particle); for it, per component:
//   tabinfo[(r*6+c)*2 + 0] = nconv: steps j >= nconv all use the fixed value; tabinfo[..+1] = 1 when steps
//   WH_KMAX <= j < nconv are stored per particle in texp (transient longer than the table);
//   tab[((r*(WH_KMAX+1)) + j)*6 + c] = t_j for j < min(nconv, WH_KMAX), slot WH_KMAX = the fixed value.
__device__ __forceinline__ double wh_run(double cm, double eta, int len, int hi, int r, int c, double *__restrict__ tab,
                                         int *__restrict__ tabinfo, double *__restrict__ texp)
{
    int nconv = 0, explicit_used = 0;
    if (len > 0) {
        double tprev = 0.0;
        int j = 0;
        for (;;) {
            const double t = dvd(cm, eta);          // :754 with m_i = 0
            cm = mul(eta, t);                       // :760 (eta[i-1] == eta[i] inside the run)
            if (j > 0 && __double_as_longlong(t) == __double_as_longlong(tprev)) { nconv = j; break; }
            if (j < WH_KMAX) tab[((size_t)r * (WH_KMAX + 1) + j) * 6 + c] = t;
            else { texp[6 * (size_t)(hi - j) + c] = t; explicit_used = 1; }
            tprev = t;
            j++;
            if (j == len) { nconv = len; break; }
        }
        tab[((size_t)r * (WH_KMAX + 1) + WH_KMAX) * 6 + c] = tprev;
    }
    tabinfo[(r * 6 + c) * 2 + 0] = nconv;
    tabinfo[(r * 6 + c) * 2 + 1] = explicit_used;
    return cm;
}

// pair_out != nullptr (massless method, K <= WH_PAIR_MAX_K): afterwards the warp also tabulates, for every gap g between
// massive particles, the sum over massive pairs straddling the gap that every massless target of that gap subtracts
// (src/integrator_whfast.c:1225-1252: same loops, same order, evaluated once instead of once per target).
constexpr int WH_PAIR_MAX_K = 64;
__device__ __forceinline__ void wh_j2c_skel_lane(int c, int n, int K, const int *__restrict__ list, const double *__restrict__ m,
                                                 const double *__restrict__ etaM, const double *__restrict__ jx,
                                                 const double *__restrict__ jv, double4 *posm, double *__restrict__ vel,
                                                 double *__restrict__ tab, int *__restrict__ tabinfo, double *__restrict__ texp)
{
    double cm = mul(eta_at(etaM, K), ld_jac(jx, jv, 0, c));                 // :742-748
    int hi = n - 1;
    bool reached_zero = false;
    for (int k = K - 1; k >= 0; k--) {
        const int mi = list[k];
        const double e_k = etaM[k];
        cm = wh_run(cm, e_k, hi - mi, hi, k + 1, c, tab, tabinfo, texp);
        if (mi == 0) { reached_zero = true; break; }
        const double j = ld_jac(jx, jv, mi, c);
        if (c == 0) reinterpret_cast<double *>(posm)[4 * (size_t)mi + 3] = m[mi];
        const double t = dvd(sub(cm, mul(m[mi], j)), e_k);                  // :754
        st_cart(posm, vel, mi, c, add(j, t));                               // :757
        cm = mul(eta_at(etaM, k), t);                                       // :760
        hi = mi - 1;
    }
    if (!reached_zero) cm = wh_run(cm, 0.0, hi, hi, 0, c, tab, tabinfo, texp);   // particle 0 itself is massless
    else if (c == 0) { for (int q = 0; q < 12; q++) tabinfo[q] = 0; }
    if (c == 0) reinterpret_cast<double *>(posm)[4 * 0 + 3] = m[0];
    st_cart(posm, vel, 0, c, dvd(cm, m[0]));                                // :765-771
}

__global__ void __launch_bounds__(256) wh_j2c_skel_kernel(int n, int K, const int *__restrict__ list, const double *__restrict__ m,
                                                          const double *__restrict__ etaM, const double *__restrict__ jx, double *jv,
                                                          double4 *posm, double *__restrict__ vel, double *__restrict__ tab,
                                                          int *__restrict__ tabinfo, double *__restrict__ texp, double G, double eps3,
                                                          const double *__restrict__ eta, double *acc, double h,
                                                          double *__restrict__ pair_out)
{
    if (blockIdx.x != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x < 6) wh_j2c_skel_lane(threadIdx.x, n, K, list, m, etaM, jx, jv, posm, vel, tab, tabinfo, texp);
    if (!pair_out) return;
    // massless method, few massive bodies: the rest of the step for the MASSIVE particles happens here too, one warp per
    // item -- K accelerations + kicks (:1006-1128, :409-422) and the K + 1 gap sums the massless targets will subtract
    __threadfence_block();
    __syncthreads();
    for (int it = warp; it < 2 * K + 1; it += nwarps) {
        if (it < K) {
            whfast_accel_massive_warp(it, posm, G, jx, eta, eps3, list, K, acc);
            __syncwarp();
            if (lane < 3) {
                const size_t e = 3 * (size_t)list[it] + lane;
                jv[e] = add(jv[e], mul(acc[e], h));
            }
        } else {
            const int g = it - K;
            const V3 s = whfast_gap_pair_sum_warp(g, posm, G, eps3, list, K);
            if (lane == 0) { pair_out[3 * g + 0] = s.x; pair_out[3 * g + 1] = s.y; pair_out[3 * g + 2] = s.z; }
        }
    }
}

// Cartesian state of massless particle i from the skeleton's tables (massive particles and particle 0 are written by
// the skeleton itself, masses included)
__device__ __forceinline__ void wh_j2c_fill_one(int i, int n, int K, const int *__restrict__ flag, const int *__restrict__ rank,
                                                const int *__restrict__ list, const double *__restrict__ m,
                                                const double *__restrict__ jx, const double *jv,
                                                const double *__restrict__ tab, const int *__restrict__ tabinfo,
                                                const double *__restrict__ texp, double4 *posm, double *__restrict__ vel)
{
    if (i == 0 || flag[i]) return;
    reinterpret_cast<double *>(posm)[4 * (size_t)i + 3] = m[i];
    const int r = rank[i];
    const int hi = (r < K) ? list[r] - 1 : n - 1;
    const int j = hi - i;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const int nconv = tabinfo[(r * 6 + c) * 2];
        double t;
        if (j >= nconv) t = tab[((size_t)r * (WH_KMAX + 1) + WH_KMAX) * 6 + c];
        else if (j < WH_KMAX) t = tab[((size_t)r * (WH_KMAX + 1) + j) * 6 + c];
        else t = texp[6 * (size_t)i + c];
        st_cart(posm, vel, i, c, add(ld_jac(jx, jv, i, c), t));
    }
}

__global__ void __launch_bounds__(256) wh_j2c_fill_kernel(int n, int K, const int *__restrict__ flag, const int *__restrict__ rank,
                                                          const int *__restrict__ list, const double *__restrict__ m,
                                                          const double *__restrict__ jx, const double *__restrict__ jv,
                                                          const double *__restrict__ tab, const int *__restrict__ tabinfo,
                                                          const double *__restrict__ texp, double4 *__restrict__ posm,
                                                          double *__restrict__ vel)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) wh_j2c_fill_one(i, n, K, flag, rank, list, m, jx, jv, tab, tabinfo, texp, posm, vel);
}

// Massless method only: the rest of the step for particle i in one kernel -- its Cartesian state, its interaction
// acceleration (which reads the particle's own record and the massive particles', all final after the skeleton kernel)
// and its kick.  With the pairwise method a target reads every other particle's record, so the three stay separate.
__global__ void __launch_bounds__(128) wh_tail_kernel(int n, int K, const int *__restrict__ flag, const int *__restrict__ rank,
                                                      const int *__restrict__ list, const double *__restrict__ m,
                                                      const double *__restrict__ jx, double *jv,
                                                      const double *__restrict__ tab, const int *__restrict__ tabinfo,
                                                      const double *__restrict__ texp, double4 *posm,
                                                      double *__restrict__ vel, double G, const double *__restrict__ eta, double eps3,
                                                      double *acc, double h, const double *__restrict__ pair_sum)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pair_sum && flag[i]) return;     // massive particles were finished by the skeleton kernel
    wh_j2c_fill_one(i, n, K, flag, rank, list, m, jx, jv, tab, tabinfo, texp, posm, vel);
    whfast_accel_one(i, n, posm, G, jx, eta, eps3, list, K, rank, acc, pair_sum);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const size_t e = 3 * (size_t)i + k;
        jv[e] = add(jv[e], mul(acc[e], h));
    }
}

// jacobi_v += a*h  (:409-422); out may alias in
__global__ void __launch_bounds__(256) wh_kick_kernel(size_t n3, const double *__restrict__ in, const double *__restrict__ a, double h,
                                                      double *__restrict__ out)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n3) out[k] = add(in[k], mul(a[k], h));
}

__global__ void wh_unpack_kernel(int n, const double4 *__restrict__ posm, double *__restrict__ x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 q = posm[i];
    x[3 * (size_t)i] = q.x; x[3 * (size_t)i + 1] = q.y; x[3 * (size_t)i + 2] = q.z;
}
__global__ void wh_pack_kernel(int n, const double *__restrict__ x, const double *__restrict__ m, double4 *__restrict__ posm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    posm[i] = make_double4(x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], m[i]);
}

// ---- host side -------------------------------------------------------------------------------------------------------------
#define WH_GRID(n, b) (unsigned)(((n) + (b) - 1) / (b)), (b), 0, c->stream
#define WH_LAUNCHED() do { GB_LAUNCH_CHECK(); count_launch(); } while (0)

static int wh_reserve(grav_b200_ctx *c, WhfastState *w, int n)
{
    const size_t b3 = sizeof(double) * 3 * (size_t)n, b1 = sizeof(double) * (size_t)n;
    for (int k = 0; k < 2; k++) {
        GB_TRY(w->jx[k].reserve(b3)); GB_TRY(w->jv[k].reserve(b3)); GB_TRY(w->m[k].reserve(b1));
        GB_TRY(w->ids[k].reserve(sizeof(int) * (size_t)n));
        GB_TRY(w->keys[k].reserve(sizeof(long long) * (size_t)n)); GB_TRY(w->perm[k].reserve(sizeof(int) * (size_t)n));
    }
    GB_TRY(w->eta.reserve(b1));
    GB_TRY(w->flag.reserve(sizeof(int) * (size_t)(n + 1)));
    GB_TRY(w->rank.reserve(sizeof(int) * (size_t)(n + 1)));
    GB_TRY(w->list.reserve(sizeof(int) * (size_t)n));
    GB_TRY(w->etaM.reserve(b1));
    GB_TRY(w->texp.reserve(2 * b3));
    GB_TRY(w->bad.reserve((size_t)n + 16));
    GB_TRY(w->status.reserve(sizeof(int) * 16));
    GB_TRY(w->ulist.reserve(sizeof(int) * (size_t)n));
    GB_TRY(w->ck_jx.reserve(b3)); GB_TRY(w->ck_jv.reserve(b3)); GB_TRY(w->ck_m.reserve(b1));
    GB_TRY(w->ck_ids.reserve(sizeof(int) * (size_t)n));
    if (!w->h_status) GB_CUDA(cudaMallocHost(&w->h_status, sizeof(int) * 16));
    return GRAV_B200_OK;
}

// Index of the particle with id 0 in the live arrays -> status[8 + cur] (start-up, after a removal, after a restore;
// inside the loop the gather kernel of the previous step provides it).
static int wh_find_primary(grav_b200_ctx *c, WhfastState *w)
{
    int *slot = w->status.as<int>() + 8 + w->cur;
    GB_CUDA(cudaMemsetAsync(slot, 0x7f, sizeof(int), c->stream));
    wh_primary_kernel<<<WH_GRID(c->n, 256)>>>(c->n, w->IDS(), 0, slot);
    WH_LAUNCHED();
    return GRAV_B200_OK;
}

// Stable sort of the live arrays by distance of the live position array from the particle with id 0 (Cartesian
// positions at start-up, Jacobi positions inside the loop).  The gather also records where id 0 went and appends the
// massive particles for the skeleton kernel.
static int wh_sort(grav_b200_ctx *c, WhfastState *w)
{
    const int n = c->n, par = w->cur;
    int *st = w->status.as<int>();
    long long *ka = w->keys[0].as<long long>(), *kb = w->keys[1].as<long long>();
    int *pa = w->perm[0].as<int>(), *pb = w->perm[1].as<int>();
    wh_dist_kernel<<<WH_GRID(n, 256)>>>(n, w->JX(), st, 8 + par, 8 + (par ^ 1), ka, pa);
    WH_LAUNCHED();
    static const bool two_launch = getenv("GRAV_B200_SORT_SMALL_TWO_LAUNCH") && atoi(getenv("GRAV_B200_SORT_SMALL_TWO_LAUNCH")) != 0;
    if (!two_launch) {
        GB_TRY(radix_sort_buffers(c, ka, pa, kb, pb, n));     // 8 passes: the result is back in ka / pa
    } else {
        for (int pass = 0; pass < 8; pass++) {
            GB_TRY(radix_pass(c, ka, pa, kb, pb, n, pass * 8));
            long long *tk = ka; ka = kb; kb = tk;
            int *tp = pa; pa = pb; pb = tp;
        }
    }
    const int o = par ^ 1;
    wh_gather_kernel<<<WH_GRID(n, 256)>>>(n, pa, w->JX(), w->JV(), w->M(), w->IDS(), w->jx[o].as<double>(), w->jv[o].as<double>(),
                                          w->m[o].as<double>(), w->ids[o].as<int>(), st, 8 + o, w->ulist.as<int>());
    WH_LAUNCHED();
    w->cur = o;
    return GRAV_B200_OK;
}

// flags -> scan -> list -> eta over the massive particles.  count_massive: also read K back (synchronises).
static int wh_massive(grav_b200_ctx *c, WhfastState *w, bool count_massive)
{
    const int n = c->n;
    int *flag = w->flag.as<int>(), *rank = w->rank.as<int>();
    wh_flag_kernel<<<WH_GRID(n + 1, 256)>>>(n, w->M(), flag);
    WH_LAUNCHED();
    GB_TRY(exclusive_scan_int(c, flag, rank, n + 1, c->misc));
    if (count_massive) {
        GB_CUDA(cudaMemcpyAsync(w->h_status + 4, rank + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        w->K = w->h_status[4];
        GB_TRY(w->tab.reserve(sizeof(double) * 6 * (WH_KMAX + 1) * (size_t)(w->K + 1)));
        GB_TRY(w->tabinfo.reserve(sizeof(int) * 12 * (size_t)(w->K + 1)));
        GB_TRY(w->cstate.reserve(sizeof(double) * 6 * (size_t)(w->K + 1)));
    }
    wh_list_kernel<<<WH_GRID(n, 256)>>>(n, flag, rank, w->list.as<int>());
    WH_LAUNCHED();
    wh_eta_skel_kernel<<<1, 32, 0, c->stream>>>(w->K, w->list.as<int>(), w->M(), w->etaM.as<double>(), w->status.as<int>());
    WH_LAUNCHED();
    return GRAV_B200_OK;
}

static int wh_j2c(grav_b200_ctx *c, WhfastState *w, const double *d_jv)
{
    const int n = c->n;
    wh_j2c_skel_kernel<<<1, 32, 0, c->stream>>>(n, w->K, w->list.as<int>(), w->M(), w->etaM.as<double>(), w->JX(),
                                                const_cast<double *>(d_jv), c->posm.as<double4>(), c->vel.as<double>(),
                                                w->tab.as<double>(), w->tabinfo.as<int>(), w->texp.as<double>(), c->G, 0.0,
                                                nullptr, nullptr, 0.0, nullptr);
    WH_LAUNCHED();
    wh_j2c_fill_kernel<<<WH_GRID(n, 256)>>>(n, w->K, w->flag.as<int>(), w->rank.as<int>(), w->list.as<int>(), w->M(), w->JX(), d_jv,
                                            w->tab.as<double>(), w->tabinfo.as<int>(), w->texp.as<double>(),
                                            c->posm.as<double4>(), c->vel.as<double>());
    WH_LAUNCHED();
    return GRAV_B200_OK;
}

static int wh_accel_kick(grav_b200_ctx *c, WhfastState *w, double h)
{
    const bool massless = w->method == GRAV_B200_METHOD_MASSLESS;
    GB_TRY(whfast_accel_with_list(c, w->JX(), w->eta.as<double>(), w->eps, massless ? w->list.as<int>() : nullptr, w->K,
                                  massless ? w->rank.as<int>() : nullptr));
    const size_t n3 = 3 * (size_t)c->n;
    wh_kick_kernel<<<WH_GRID(n3, 256)>>>(n3, w->JV(), c->acc.as<double>(), h, w->JV());
    WH_LAUNCHED();
    return GRAV_B200_OK;
}

static int wh_check_status(grav_b200_ctx *c, WhfastState *w)
{
    if (w->h_status[3] & 1) { set_error("Primary particle ID not found in system"); return GRAV_B200_EINVAL; }
    if (w->h_status[3] & 2) { set_error("whfast: the number of massive particles changed inside a batch"); return GRAV_B200_ECUDA; }
    (void)c;
    return GRAV_B200_OK;
}

// front half of a step: sort, eta, drift (:299-327 without the removal)
static int wh_step_front(grav_b200_ctx *c, WhfastState *w, double dt)
{
    GB_TRY(wh_sort(c, w));
    const bool small_k = w->K <= w->skel_max_k;
    if (small_k) {
        wh_skel_kernel<<<1, 256, 0, c->stream>>>(w->K, w->ulist.as<int>(), w->list.as<int>(), w->M(), w->etaM.as<double>(),
                                                 w->status.as<int>());
        WH_LAUNCHED();
    } else {
        GB_TRY(wh_massive(c, w, false));
    }
    wh_drift_kernel<<<WH_GRID(c->n, 128)>>>(c->n, w->JX(), w->JV(), w->M(), w->rank.as<int>(), w->etaM.as<double>(),
                                            w->eta.as<double>(), c->G, dt, w->remove_invalid ? 1 : 0, w->bad.as<char>(),
                                            w->status.as<int>(), small_k ? w->list.as<int>() : nullptr, w->K, w->rank.as<int>(),
                                            w->flag.as<int>());
    WH_LAUNCHED();
    return GRAV_B200_OK;
}
// back half: Jacobi -> Cartesian, acceleration, kick (:329-340)
static int wh_step_back(grav_b200_ctx *c, WhfastState *w, double dt)
{
    if (w->method == GRAV_B200_METHOD_MASSLESS) {
        const int n = c->n;
        const double eps3 = w->eps * w->eps * w->eps;
        double *pairs = w->K <= w->pair_max_k ? w->cstate.as<double>() : nullptr;     // cstate (6 (K+1) doubles) is free after begin()
        wh_j2c_skel_kernel<<<1, pairs ? 256 : 32, 0, c->stream>>>(n, w->K, w->list.as<int>(), w->M(), w->etaM.as<double>(), w->JX(),
                                                                  w->JV(), c->posm.as<double4>(), c->vel.as<double>(),
                                                                  w->tab.as<double>(), w->tabinfo.as<int>(), w->texp.as<double>(),
                                                                  c->G, eps3, w->eta.as<double>(), c->acc.as<double>(), dt, pairs);
        WH_LAUNCHED();
        wh_tail_kernel<<<WH_GRID(n, 128)>>>(n, w->K, w->flag.as<int>(), w->rank.as<int>(), w->list.as<int>(), w->M(), w->JX(), w->JV(),
                                            w->tab.as<double>(), w->tabinfo.as<int>(), w->texp.as<double>(), c->posm.as<double4>(),
                                            c->vel.as<double>(), c->G, w->eta.as<double>(), eps3, c->acc.as<double>(), dt, pairs);
        WH_LAUNCHED();
    } else {
        GB_TRY(wh_j2c(c, w, w->JV()));
        GB_TRY(wh_accel_kick(c, w, dt));
    }
    w->last_dt = dt;
    return GRAV_B200_OK;
}

// Two steps as one graph launch.  The first use with a given (parity, n, K, dt) captures the launches of two ordinary
// steps from the stream; every buffer already has its final size by then (at least one plain step has run).
static int wh_step_pair(grav_b200_ctx *c, WhfastState *w, double dt)
{
    const int par = w->cur;
    if (!w->pair_exec[par] || w->pair_n[par] != c->n || w->pair_K[par] != w->K || w->pair_dt[par] != dt ||
        w->pair_gen[par] != g_alloc_generation) {
        if (w->pair_exec[par]) { cudaGraphExecDestroy(w->pair_exec[par]); w->pair_exec[par] = nullptr; }
        const int64_t l0 = g_launch_count;
        GB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = wh_step_front(c, w, dt);
        if (rc == GRAV_B200_OK) rc = wh_step_back(c, w, dt);
        if (rc == GRAV_B200_OK) rc = wh_step_front(c, w, dt);
        if (rc == GRAV_B200_OK) rc = wh_step_back(c, w, dt);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        const int captured = (int)(g_launch_count - l0);
        __atomic_fetch_sub(&g_launch_count, (int64_t)captured, __ATOMIC_RELAXED);      // nothing ran yet
        if (rc != GRAV_B200_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        GB_CUDA(e);
        const cudaError_t ei = cudaGraphInstantiate(&w->pair_exec[par], graph, 0);
        cudaGraphDestroy(graph);
        GB_CUDA(ei);
        w->pair_n[par] = c->n; w->pair_K[par] = w->K; w->pair_dt[par] = dt; w->pair_launches[par] = captured;
        w->pair_gen[par] = g_alloc_generation;
    }
    GB_CUDA(cudaGraphLaunch(w->pair_exec[par], c->stream));
    count_launch(w->pair_launches[par]);
    w->last_dt = dt;
    return GRAV_B200_OK;
}

static int wh_steps_plain_or_graph(grav_b200_ctx *c, WhfastState *w, double dt, int count, bool use_graph)
{
    int s = 0;
    if (use_graph && count >= 3) {
        GB_TRY(wh_step_front(c, w, dt));       // sizes every scratch buffer for this n before a capture
        GB_TRY(wh_step_back(c, w, dt));
        s = 1;
        for (; s + 2 <= count; s += 2) GB_TRY(wh_step_pair(c, w, dt));
    }
    for (; s < count; s++) {
        GB_TRY(wh_step_front(c, w, dt));
        GB_TRY(wh_step_back(c, w, dt));
    }
    return GRAV_B200_OK;
}

// The reference's message before a removal (src/integrator_whfast.c:609-623, verbose >= GRAV_VERBOSITY_VERBOSE): ids of the
// flagged particles in index order.  The host is synchronised here anyway (it just read the status words).
static int wh_print_removed(grav_b200_ctx *c, WhfastState *w, int n_removed)
{
    const int n = c->n;
    char *bad = (char *)malloc((size_t)n);
    int *ids = (int *)malloc(sizeof(int) * (size_t)n);
    if (!bad || !ids) { free(bad); free(ids); set_error("out of host memory"); return GRAV_B200_ECUDA; }
    cudaError_t e = cudaMemcpyAsync(bad, w->bad.as<char>(), (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ids, w->IDS(), sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { free(bad); free(ids); return cuda_fail(e, "download of the removal flags", __FILE__, __LINE__); }
    fprintf(stderr, "whfast_drift: Removing %d invalid particles. Particle IDs: [", n_removed);
    bool first = true;
    for (int i = 0; i < n; i++) {
        if (!bad[i]) continue;
        fprintf(stderr, first ? "%d" : ", %d", ids[i]);
        first = false;
    }
    fputs("]\n", stderr);
    free(bad); free(ids);
    return GRAV_B200_OK;
}

// the removal itself (:607-671): stable compaction of the Jacobi state, masses and ids; eta recomputed
static int wh_remove_flagged(grav_b200_ctx *c, WhfastState *w, int n_removed)
{
    const int n = c->n;
    int *keep = w->flag.as<int>(), *pos = w->rank.as<int>();
    wh_keep_flag_kernel<<<WH_GRID(n + 1, 256)>>>(n, w->bad.as<char>(), keep);
    WH_LAUNCHED();
    GB_TRY(exclusive_scan_int(c, keep, pos, n + 1, c->misc));
    const int o = w->cur ^ 1;
    wh_compact_kernel<<<WH_GRID(n, 256)>>>(n, keep, pos, w->JX(), w->JV(), w->M(), w->IDS(), w->jx[o].as<double>(),
                                           w->jv[o].as<double>(), w->m[o].as<double>(), w->ids[o].as<int>());
    WH_LAUNCHED();
    w->cur = o;
    const int n_new = n - n_removed;
    const int pad_old = c->n_pad;
    c->n = n_new;
    c->n_pad = ((n_new + SRC_PAD - 1) / SRC_PAD) * SRC_PAD;
    c->lo = 0; c->hi = n_new;
    // keep the "padding is zero" invariant of posm for whoever uses the context next
    GB_CUDA(cudaMemsetAsync(c->posm.as<double4>() + n_new, 0, sizeof(double4) * (size_t)(pad_old - n_new), c->stream));
    GB_TRY(wh_massive(c, w, true));
    wh_eta_fill_kernel<<<WH_GRID(n_new, 256)>>>(n_new, w->rank.as<int>(), w->etaM.as<double>(), w->eta.as<double>());
    WH_LAUNCHED();
    GB_TRY(wh_find_primary(c, w));      // the compaction moved the particles and flipped the live buffers
    return GRAV_B200_OK;
}

static int wh_reset_batch_status(grav_b200_ctx *c, WhfastState *w)
{
    // constants only: the pinned words may still be in flight from an earlier reset
    w->h_status[8] = 0; w->h_status[9] = WH_NO_BAD; w->h_status[10] = 0; w->h_status[11] = 0; w->h_status[12] = 0; w->h_status[13] = 0;
    GB_CUDA(cudaMemcpyAsync(w->status.as<int>(), w->h_status + 8, sizeof(int) * 6, cudaMemcpyHostToDevice, c->stream));
    return GRAV_B200_OK;
}
static int wh_read_status(grav_b200_ctx *c, WhfastState *w)
{
    GB_CUDA(cudaMemcpyAsync(w->h_status, w->status.as<int>(), sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return wh_check_status(c, w);
}

}  // namespace gb

using namespace gb;

extern "C" {

int grav_b200_ctx_whfast_begin(grav_b200_ctx *c, const int *particle_ids, int method, double eps, double dt,
                               int remove_invalid_particles)
{
    if (c && c->team) { set_error("the device-resident WHFast runs on one GPU: create its context with grav_b200_ctx_create(), not as a device team"); return GRAV_B200_EINVAL; }
    if (!c || c->n < 1) { set_error("context has no system"); return GRAV_B200_EINVAL; }
    if (c->world != 1) { set_error("device-resident WHFast runs on one GPU"); return GRAV_B200_EINVAL; }
    if (method != GRAV_B200_METHOD_PAIRWISE && method != GRAV_B200_METHOD_MASSLESS) {
        set_error("Invalid acceleration method for WHFast integrator. Only pairwise and massless are supported");
        return GRAV_B200_EINVAL;
    }
    if (eps < 0.0) { set_error("Softening length is negative. Got: %.3g", eps); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    if (!c->wh) c->wh = new WhfastState();
    WhfastState *w = (WhfastState *)c->wh;
    const int n = c->n;
    GB_TRY(wh_reserve(c, w, n));
    w->method = method; w->eps = eps; w->remove_invalid = remove_invalid_particles != 0; w->ready = false; w->cur = 0;
    for (int k = 0; k < 2; k++)      // method, eps and the flags are baked into captured steps
        if (w->pair_exec[k]) { cudaGraphExecDestroy(w->pair_exec[k]); w->pair_exec[k] = nullptr; }
    {   // test hooks: force the large-K code paths with a handful of massive bodies
        const char *e1 = getenv("GRAV_B200_WHFAST_SKEL_MAX_K"), *e2 = getenv("GRAV_B200_WHFAST_PAIR_MAX_K");
        w->skel_max_k = e1 ? (atoi(e1) < WH_SKEL_MAX_K ? atoi(e1) : WH_SKEL_MAX_K) : WH_SKEL_MAX_K;
        w->pair_max_k = e2 ? (atoi(e2) < WH_PAIR_MAX_K ? atoi(e2) : WH_PAIR_MAX_K) : WH_PAIR_MAX_K;
    }
    c->lf_ready = false;
    c->mlist_valid = false;
    c->sym_eqm_valid = false;      // this integrator reorders (and may remove) particles: posm's masses move
    // live arrays <- Cartesian state (x unpacked from posm, v, m) + ids
    wh_unpack_kernel<<<WH_GRID(n, 256)>>>(n, c->posm.as<double4>(), w->JX());
    WH_LAUNCHED();
    GB_CUDA(cudaMemcpyAsync(w->JV(), c->vel.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    GB_CUDA(cudaMemcpy2DAsync(w->M(), sizeof(double), reinterpret_cast<const double *>(c->posm.p) + 3, sizeof(double4), sizeof(double),
                              (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    if (particle_ids) {
        GB_CUDA(cudaMemcpyAsync(w->IDS(), particle_ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    } else {
        int *h = (int *)malloc(sizeof(int) * (size_t)n);
        if (!h) { set_error("out of host memory"); return GRAV_B200_ENOMEM; }
        for (int i = 0; i < n; i++) h[i] = i;
        cudaError_t e = cudaMemcpyAsync(w->IDS(), h, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        free(h);
        GB_CUDA(e);
    }
    GB_TRY(wh_reset_batch_status(c, w));
    GB_CUDA(cudaMemsetAsync(w->status.as<int>() + 6, 0, sizeof(int), c->stream));
    GB_TRY(wh_find_primary(c, w));
    GB_TRY(wh_sort(c, w));                                   // :242, on the Cartesian positions
    GB_TRY(wh_read_status(c, w));
    // sorted Cartesian state back into posm / vel
    wh_pack_kernel<<<WH_GRID(n, 256)>>>(n, w->JX(), w->M(), c->posm.as<double4>());
    WH_LAUNCHED();
    GB_CUDA(cudaMemcpyAsync(c->vel.p, w->JV(), sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    GB_TRY(wh_massive(c, w, true));                          // :266
    wh_eta_fill_kernel<<<WH_GRID(n, 256)>>>(n, w->rank.as<int>(), w->etaM.as<double>(), w->eta.as<double>());
    WH_LAUNCHED();
    // cartesian_to_jacobi (:267); jacobi_x starts zeroed (calloc, :228), `a` is defined as zero where the reference
    // leaves it unwritten (entry 0)
    GB_CUDA(cudaMemsetAsync(w->JX(), 0, sizeof(double) * 3 * (size_t)n, c->stream));
    GB_CUDA(cudaMemsetAsync(c->acc.p, 0, sizeof(double) * 3 * (size_t)n, c->stream));
    wh_c2j_skel_kernel<<<1, 32, 0, c->stream>>>(n, w->K, w->list.as<int>(), w->M(), w->etaM.as<double>(), c->posm.as<double4>(),
                                                c->vel.as<double>(), w->JX(), w->JV(), w->cstate.as<double>());
    WH_LAUNCHED();
    wh_c2j_fill_kernel<<<WH_GRID(n, 256)>>>(n, w->flag.as<int>(), w->rank.as<int>(), w->etaM.as<double>(), w->cstate.as<double>(),
                                            c->posm.as<double4>(), c->vel.as<double>(), w->JX(), w->JV());
    WH_LAUNCHED();
    GB_TRY(wh_accel_kick(c, w, 0.5 * dt));                   // :268-273
    w->last_dt = dt;
    w->ready = true;
    return GRAV_B200_OK;
}

int grav_b200_ctx_whfast_set_verbose(grav_b200_ctx *c, int level)
{
    WhfastState *w = c ? (WhfastState *)c->wh : nullptr;
    if (!w) { set_error("whfast_begin() has not been called"); return GRAV_B200_EINVAL; }
    w->verbose = level;
    return GRAV_B200_OK;
}

int grav_b200_ctx_whfast_steps(grav_b200_ctx *c, double dt, int64_t num_steps)
{
    WhfastState *w = c ? (WhfastState *)c->wh : nullptr;
    if (!w || !w->ready) { set_error("whfast_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    int64_t remaining = num_steps;
    const char *genv = getenv("GRAV_B200_WHFAST_GRAPH");
    const bool use_graph = !(genv && genv[0] == '0');
    while (remaining > 0) {
        const int B = (int)(remaining < WH_BATCH ? remaining : WH_BATCH);
        const int n = c->n;
        if (w->remove_invalid) {     // checkpoint of everything a step depends on
            GB_CUDA(cudaMemcpyAsync(w->ck_jx.p, w->JX(), sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
            GB_CUDA(cudaMemcpyAsync(w->ck_jv.p, w->JV(), sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
            GB_CUDA(cudaMemcpyAsync(w->ck_m.p, w->M(), sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
            GB_CUDA(cudaMemcpyAsync(w->ck_ids.p, w->IDS(), sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        }
        GB_TRY(wh_reset_batch_status(c, w));
        GB_TRY(wh_steps_plain_or_graph(c, w, dt, B, use_graph));
        GB_TRY(wh_read_status(c, w));
        const int first_bad = w->h_status[1];
        if (!w->remove_invalid || first_bad == WH_NO_BAD) { remaining -= B; continue; }
        // replay: restore, redo the clean steps, then finish the flagged step with the removal
        GB_CUDA(cudaMemcpyAsync(w->JX(), w->ck_jx.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        GB_CUDA(cudaMemcpyAsync(w->JV(), w->ck_jv.p, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        GB_CUDA(cudaMemcpyAsync(w->M(), w->ck_m.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        GB_CUDA(cudaMemcpyAsync(w->IDS(), w->ck_ids.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        GB_TRY(wh_reset_batch_status(c, w));
        GB_TRY(wh_find_primary(c, w));
        GB_TRY(wh_steps_plain_or_graph(c, w, dt, first_bad, use_graph));
        GB_TRY(wh_reset_batch_status(c, w));
        GB_TRY(wh_step_front(c, w, dt));
        GB_TRY(wh_read_status(c, w));
        const int n_removed = w->h_status[2];
        if (n_removed <= 0 || n_removed >= n) { set_error("whfast replay lost the flagged particles (%d of %d)", n_removed, n); return GRAV_B200_ECUDA; }
        if (w->verbose >= 3) GB_TRY(wh_print_removed(c, w, n_removed));
        GB_TRY(wh_remove_flagged(c, w, n_removed));
        GB_TRY(wh_step_back(c, w, dt));
        remaining -= (first_bad + 1);
    }
    return GRAV_B200_OK;
}

// snapshot != 0: velocities kicked back by -dt/2 before the conversion, like the reference's output branch (:346-351),
// which also leaves system->x / system->v in that state.  snapshot == 0: the state the last step left.
int grav_b200_ctx_whfast_get_state(grav_b200_ctx *c, int snapshot, int *n_out, int *particle_ids, double *x, double *v, double *m)
{
    WhfastState *w = c ? (WhfastState *)c->wh : nullptr;
    if (!w || !w->ready) { set_error("whfast_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    const int n = c->n;
    const size_t b3 = sizeof(double) * 3 * (size_t)n;
    if (snapshot) {
        double *tmp = w->jv[w->cur ^ 1].as<double>();
        const size_t n3 = 3 * (size_t)n;
        wh_kick_kernel<<<WH_GRID(n3, 256)>>>(n3, w->JV(), c->acc.as<double>(), -0.5 * w->last_dt, tmp);
        WH_LAUNCHED();
        GB_TRY(wh_j2c(c, w, tmp));
    }
    if (n_out) *n_out = n;
    if (x) {
        GB_TRY(c->stage_a.reserve(b3));
        wh_unpack_kernel<<<WH_GRID(n, 256)>>>(n, c->posm.as<double4>(), c->stage_a.as<double>());
        WH_LAUNCHED();
        GB_CUDA(cudaMemcpyAsync(x, c->stage_a.p, b3, cudaMemcpyDeviceToHost, c->stream));
    }
    if (v) GB_CUDA(cudaMemcpyAsync(v, c->vel.p, b3, cudaMemcpyDeviceToHost, c->stream));
    if (m) GB_CUDA(cudaMemcpyAsync(m, w->M(), sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (particle_ids) GB_CUDA(cudaMemcpyAsync(particle_ids, w->IDS(), sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return GRAV_B200_OK;
}

int grav_b200_ctx_whfast_end(grav_b200_ctx *c)
{
    WhfastState *w = c ? (WhfastState *)c->wh : nullptr;
    if (!w) { set_error("whfast_begin() has not been called"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    whfast_state_free(c);
    return GRAV_B200_OK;
}

}  // extern "C"
