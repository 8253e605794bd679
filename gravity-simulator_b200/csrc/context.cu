// Context life-cycle, host<->device staging and the C ABI entry points of libgrav_b200.so.
#include <stdarg.h>
#include <mutex>
#include "internal.cuh"

namespace gb {

// ---- errors -----------------------------------------------------------------------------
static thread_local char t_err[512] = "";
int64_t g_launch_count = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    const char *base = strrchr(file, '/');
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, base ? base + 1 : file, line);
    if (e == cudaErrorMemoryAllocation) return GRAV_B200_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return GRAV_B200_ENODEV;
    return GRAV_B200_ECUDA;
}

uint64_t g_alloc_generation = 0;   // bumped whenever a device buffer moves: captured CUDA graphs hold raw pointers

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap && p) return GRAV_B200_OK;
    __atomic_fetch_add(&g_alloc_generation, (uint64_t)1, __ATOMIC_RELAXED);
    if (bytes == 0) bytes = 256;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    // grow geometrically so per-call scratch settles quickly
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    cap = want;
    return GRAV_B200_OK;
}

void DevBuf::release()
{
    if (p) __atomic_fetch_add(&g_alloc_generation, (uint64_t)1, __ATOMIC_RELAXED);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ---- AoS <-> packed records ---------------------------------------------------------------
// posm[i] = (x[3i], x[3i+1], x[3i+2], m[i]); entries n..n_pad-1 are zero.
__global__ void pack_posm_kernel(const double *__restrict__ x, const double *__restrict__ m, double4 *__restrict__ posm,
                                 int n, int n_pad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double4 q = make_double4(0.0, 0.0, 0.0, 0.0);
    if (i < n) {
        q.x = x[3 * (size_t)i + 0];
        q.y = x[3 * (size_t)i + 1];
        q.z = x[3 * (size_t)i + 2];
        q.w = m[i];
    }
    posm[i] = q;
}

__global__ void pack_positions_kernel(const double *__restrict__ x, double4 *__restrict__ posm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 q = posm[i];
    q.x = x[3 * (size_t)i + 0];
    q.y = x[3 * (size_t)i + 1];
    q.z = x[3 * (size_t)i + 2];
    posm[i] = q;
}

__global__ void unpack_positions_kernel(const double4 *__restrict__ posm, double *__restrict__ x, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 q = posm[i];
    x[3 * (size_t)i + 0] = q.x;
    x[3 * (size_t)i + 1] = q.y;
    x[3 * (size_t)i + 2] = q.z;
}

int pack_posm(grav_b200_ctx *c, const double *d_x, const double *d_m)
{
    pack_posm_kernel<<<(c->n_pad + 255) / 256, 256, 0, c->stream>>>(d_x, d_m, c->posm.as<double4>(), c->n, c->n_pad);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}
int pack_positions(grav_b200_ctx *c, const double *d_x)
{
    pack_positions_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(d_x, c->posm.as<double4>(), c->n);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}
int unpack_positions(grav_b200_ctx *c, double *d_x)
{
    unpack_positions_kernel<<<(c->n + 255) / 256, 256, 0, c->stream>>>(c->posm.as<double4>(), d_x, c->n);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

// ---- default context for the host-pointer one-shots -------------------------------------
static std::mutex g_mu;
static grav_b200_ctx *g_default = nullptr;
static int g_bh_mode = -1;
static int g_bh_exact = -1;

static int default_ctx(grav_b200_ctx **out)
{
    if (!g_default) GB_TRY(grav_b200_ctx_create_auto(&g_default));   // GRAV_B200_DEVICE / GRAV_B200_DEVICES (team.cu)
    // the reference calls acceleration() from whatever thread owns the simulation
    // (a non-main Python thread, grav_sim/simulator.py:61-103): bind the device there
    GB_CUDA(cudaSetDevice(g_default->device));
    const int mode = grav_b200_get_bh_mode(), exact = grav_b200_get_bh_exact();
    if (g_default->bh_mode != mode || g_default->bh_exact != exact) {
        if (g_default->team) GB_TRY(team_run(g_default, [=](grav_b200_ctx *r) { r->bh_mode = mode; r->bh_exact = exact; return GRAV_B200_OK; }));
        g_default->bh_mode = mode;
        g_default->bh_exact = exact;
    }
    *out = g_default;
    return GRAV_B200_OK;
}

int default_ctx_locked_begin(grav_b200_ctx **out)
{
    g_mu.lock();
    const int rc = default_ctx(out);
    if (rc != GRAV_B200_OK) g_mu.unlock();
    return rc;
}
void default_ctx_locked_end() { g_mu.unlock(); }

static int check_sys(const void *a, int n, const void *x, const void *m)
{
    if (!a || !x || !m) { set_error("NULL array pointer"); return GRAV_B200_EINVAL; }
    if (n < 1) { set_error("num_particles must be >= 1, got %d", n); return GRAV_B200_EINVAL; }
    return GRAV_B200_OK;
}

}  // namespace gb

using namespace gb;

extern "C" {

const char *grav_b200_last_error(void) { return t_err; }

int grav_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t grav_b200_kernel_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

int grav_b200_set_bh_mode(int mode)
{
    if (mode != GRAV_B200_BH_REFERENCE && mode != GRAV_B200_BH_FIXED) { set_error("unknown BH mode %d", mode); return GRAV_B200_EINVAL; }
    g_bh_mode = mode;
    return GRAV_B200_OK;
}

int grav_b200_get_bh_mode(void)
{
    if (g_bh_mode < 0) {
        const char *e = getenv("GRAV_B200_BH_MODE");
        g_bh_mode = (e && strcmp(e, "fixed") == 0) ? GRAV_B200_BH_FIXED : GRAV_B200_BH_REFERENCE;
    }
    return g_bh_mode;
}

int grav_b200_set_bh_exact(int on)
{
    g_bh_exact = on ? 1 : 0;
    return GRAV_B200_OK;
}

static int g_ds_mode = -2;   // -2: not read from the environment yet
int grav_b200_get_direct_sum_mode(void)
{
    if (g_ds_mode == -2) {
        const char *e = getenv("GRAV_B200_DS_SYM");
        g_ds_mode = e ? (atoi(e) > 0 ? 1 : (atoi(e) == 0 ? 0 : -1)) : -1;
    }
    return g_ds_mode;
}
int grav_b200_set_direct_sum_mode(int mode)
{
    g_ds_mode = mode > 0 ? 1 : (mode == 0 ? 0 : -1);
    return GRAV_B200_OK;
}

int grav_b200_get_bh_exact(void)
{
    if (g_bh_exact < 0) {
        const char *e = getenv("GRAV_B200_BH_EXACT");
        g_bh_exact = (e && atoi(e) != 0) ? 1 : 0;
    }
    return g_bh_exact;
}

int grav_b200_ctx_create(grav_b200_ctx **out, int device, int rank, int world_size, const void *uid)
{
    if (!out) { set_error("NULL out pointer"); return GRAV_B200_EINVAL; }
    *out = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size) { set_error("bad rank/world_size %d/%d", rank, world_size); return GRAV_B200_EINVAL; }
    if (world_size > 1 && !uid) { set_error("world_size > 1 needs an NCCL unique id"); return GRAV_B200_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libgrav_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return GRAV_B200_ENODEV;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range [0,%d)", device, ndev); return GRAV_B200_ENODEV; }
    GB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return GRAV_B200_ENODEV;
    }
    grav_b200_ctx *c = new grav_b200_ctx();
    c->device = device;
    c->rank = rank;
    c->world = world_size;
    c->sm_count = prop.multiProcessorCount;
    c->bh_mode = grav_b200_get_bh_mode();
    c->bh_exact = grav_b200_get_bh_exact();
    if (const char *sl = getenv("GRAV_B200_TREE_SLACK")) { const int v = atoi(sl); if (v >= 1 && v <= 16) c->tree.slack = v; }
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete c; return cuda_fail(se, "cudaStreamCreate", __FILE__, __LINE__); }
    for (int i = 0; i < 2 * ST_COUNT && se == cudaSuccess; i++) se = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 8 && se == cudaSuccess; i++) se = cudaEventCreate(&c->user_ev[i]);
    if (se != cudaSuccess) { grav_b200_ctx_destroy(c); return cuda_fail(se, "cudaEventCreate", __FILE__, __LINE__); }
    if (world_size > 1) {
        int rc = comm_init(c, uid);
        if (rc != GRAV_B200_OK) { grav_b200_ctx_destroy(c); return rc; }
    }
    *out = c;
    return GRAV_B200_OK;
}

void grav_b200_ctx_destroy(grav_b200_ctx *c)
{
    if (!c) return;
    if (team_active(c)) team_destroy(c);   // the workers destroy their own contexts on their own threads
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    whfast_state_free(c);
    comm_destroy(c);
    DevBuf *bufs[] = {&c->posm, &c->vel, &c->acc, &c->xcomp, &c->vcomp, &c->stage_a, &c->stage_b, &c->stage_c, &c->stage_d,
                      &c->partials, &c->sym_priv, &c->sym_flag, &c->misc, &c->rk_buf, &c->msrc, &c->msrc_id, &c->msrc_altm, &c->l2_flush, &c->mflag, &c->mrank};
    for (DevBuf *b : bufs) b->release();
    DevTree &t = c->tree;
    DevBuf *tb[] = {&t.keys_unsorted, &t.keys, &t.perm, &t.keys_tmp, &t.perm_tmp, &t.hist, &t.bbox, &t.exp_rec,
                    &t.wsum, &t.wscan, &t.scan_tmp, &t.meta, &t.node_mtd, &t.node_walk, &t.posm_sorted, &t.ki, &t.tord,
                    &t.walk_out, &t.xport};
    for (DevBuf *b : tb) b->release();
    if (t.h_meta) cudaFreeHost(t.h_meta);
    t.h_meta = nullptr;
    for (int i = 0; i < 2 * ST_COUNT; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 8; i++) if (c->user_ev[i]) cudaEventDestroy(c->user_ev[i]);
    mailbox_free(c);
    if (c->small_pinned) cudaFreeHost(c->small_pinned);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int grav_b200_ctx_num_particles(const grav_b200_ctx *c) { return c ? c->n : 0; }
void grav_b200_ctx_owned_range(const grav_b200_ctx *c, int *lo, int *hi)
{
    if (lo) *lo = c ? c->lo : 0;
    if (hi) *hi = c ? c->hi : 0;
}

// need_vel = false: the caller never reads velocities (host-pointer one-shots), so a NULL v costs nothing
static int set_system_impl(grav_b200_ctx *c, int n, const double *x, const double *v, const double *m, double G, bool need_vel)
{
    if (!c) { set_error("NULL context"); return GRAV_B200_EINVAL; }
    if (!x || !m) { set_error("NULL array pointer"); return GRAV_B200_EINVAL; }
    if (n < 1) { set_error("num_particles must be >= 1, got %d", n); return GRAV_B200_EINVAL; }
    if (n > GRAV_B200_MAX_PARTICLES) {   // 32-bit work-unit and packed level/count fields; 2^24 is BASELINE.json's largest size
        set_error("num_particles %d exceeds the supported maximum %d", n, GRAV_B200_MAX_PARTICLES);
        return GRAV_B200_EINVAL;
    }
    GB_CUDA(cudaSetDevice(c->device));
    c->n = n;
    c->n_pad = ((n + SRC_PAD - 1) / SRC_PAD) * SRC_PAD;
    c->G = G;
    c->lo = (int)(((long long)c->rank * n) / c->world);
    c->hi = (int)(((long long)(c->rank + 1) * n) / c->world);
    c->lf_ready = false;
    c->fixed_integrator = 0;
    c->mlist_valid = false;
    c->sym_eqm_valid = false;
    const size_t b3 = sizeof(double) * 3 * (size_t)n;
    GB_TRY(c->posm.reserve(sizeof(double4) * (size_t)c->n_pad));
    GB_TRY(c->acc.reserve(b3));
    GB_TRY(c->vel.reserve(b3));
    GB_TRY(c->stage_a.reserve(b3));
    GB_TRY(c->stage_b.reserve(sizeof(double) * (size_t)n));
    GB_CUDA(cudaMemcpyAsync(c->stage_a.p, x, b3, cudaMemcpyHostToDevice, c->stream));
    GB_CUDA(cudaMemcpyAsync(c->stage_b.p, m, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    GB_TRY(pack_posm(c, c->stage_a.as<double>(), c->stage_b.as<double>()));
    if (v) GB_CUDA(cudaMemcpyAsync(c->vel.p, v, b3, cudaMemcpyHostToDevice, c->stream));
    else if (need_vel) GB_CUDA(cudaMemsetAsync(c->vel.p, 0, b3, c->stream));
    c->posm_gathered = true;
    return GRAV_B200_OK;
}

int grav_b200_ctx_set_system(grav_b200_ctx *c, int n, const double *x, const double *v, const double *m, double G)
{
    GB_TEAM(c, grav_b200_ctx_set_system(r_, n, x, v, m, G));
    return set_system_impl(c, n, x, v, m, G, true);
}

// the one-shots' upload: velocities are never read
static int set_system_no_velocities(grav_b200_ctx *c, int n, const double *x, const double *m, double G)
{
    GB_TEAM(c, set_system_no_velocities(r_, n, x, m, G));
    return set_system_impl(c, n, x, nullptr, m, G, false);
}

int grav_b200_ctx_set_positions(grav_b200_ctx *c, const double *x)
{
    GB_TEAM(c, grav_b200_ctx_set_positions(r_, x));
    if (!c || !x || c->n < 1) { set_error("context has no system / NULL pointer"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    const size_t b3 = sizeof(double) * 3 * (size_t)c->n;
    GB_CUDA(cudaMemcpyAsync(c->stage_a.p, x, b3, cudaMemcpyHostToDevice, c->stream));
    GB_TRY(pack_positions(c, c->stage_a.as<double>()));
    c->posm_gathered = true;
    return GRAV_B200_OK;
}

int grav_b200_ctx_acceleration(grav_b200_ctx *c, int method, double eps, double theta, int max_leaf)
{
    GB_TEAM(c, grav_b200_ctx_acceleration(r_, method, eps, theta, max_leaf));
    if (!c || c->n < 1) { set_error("context has no system"); return GRAV_B200_EINVAL; }
    if (eps < 0.0) { set_error("Softening length is negative. Got: %.3g", eps); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    for (int s = 0; s < ST_COUNT; s++) c->ev_valid[s] = false;
    stage_begin(c, ST_TOTAL);
    if (c->world > 1 && !c->posm_gathered) {
        stage_begin(c, ST_GATHER);
        GB_TRY(comm_allgather_posm(c));
        stage_end(c, ST_GATHER);
        c->posm_gathered = true;
    }
    int rc;
    switch (method) {
        case GRAV_B200_METHOD_PAIRWISE:
            stage_begin(c, ST_FORCE);
            rc = direct_sum_pairwise(c, eps);
            stage_end(c, ST_FORCE);
            break;
        case GRAV_B200_METHOD_MASSLESS:
            stage_begin(c, ST_FORCE);
            rc = direct_sum_massless(c, eps);
            stage_end(c, ST_FORCE);
            break;
        case GRAV_B200_METHOD_BARNES_HUT:
            if (theta < 0.0) { set_error("Opening angle is negative. Got: %.3g", theta); return GRAV_B200_EINVAL; }
            if (max_leaf == -1) max_leaf = 1;
            if (max_leaf < 1) { set_error("Maximum number of particles per leaf must be positive. Got: %d", max_leaf); return GRAV_B200_EINVAL; }
            rc = bh_build(c, max_leaf, nullptr, -1.0);
            if (rc == GRAV_B200_OK) {
                stage_begin(c, ST_FORCE);
                rc = bh_walk(c, eps, theta);
                stage_end(c, ST_FORCE);
            }
            break;
        default:
            set_error("Unknown acceleration method. Got: %d", method);
            return GRAV_B200_EINVAL;
    }
    stage_end(c, ST_TOTAL);
    return rc;
}

static int download_aos3(grav_b200_ctx *c, double *d_src, double *h_dst, bool sharded)
{
    GB_CUDA(cudaSetDevice(c->device));
    if (c->team) {
        // in-process team: the host array is shared by all members, so each writes the slice of its own targets (valid on
        // every rank whatever produced it: owned shards of v and of the direct-sum a, the full Barnes-Hut a, own positions)
        const size_t lo = 3 * (size_t)c->lo, cnt = 3 * (size_t)(c->hi - c->lo);
        if (cnt) GB_CUDA(cudaMemcpyAsync(h_dst + lo, d_src + lo, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        return bh_check(c);
    }
    if (sharded && c->world > 1) GB_TRY(comm_allgather_aos3(c, d_src));
    GB_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double) * 3 * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return bh_check(c);
}

int grav_b200_ctx_get_positions(grav_b200_ctx *c, double *x)
{
    if (!c || !x || c->n < 1) { set_error("context has no system / NULL pointer"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_get_positions(r_, x));
    GB_CUDA(cudaSetDevice(c->device));
    if (c->world > 1 && !c->posm_gathered && !c->team) { GB_TRY(comm_allgather_posm(c)); c->posm_gathered = true; }
    GB_TRY(unpack_positions(c, c->stage_a.as<double>()));
    return download_aos3(c, c->stage_a.as<double>(), x, false);
}
int grav_b200_ctx_get_velocities(grav_b200_ctx *c, double *v)
{
    if (!c || !v || c->n < 1) { set_error("context has no system / NULL pointer"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_get_velocities(r_, v));
    GB_CUDA(cudaSetDevice(c->device));
    double *d_v;
    GB_TRY(synced_velocities(c, &d_v));
    return download_aos3(c, d_v, v, true);
}
int grav_b200_ctx_get_accelerations(grav_b200_ctx *c, double *a)
{
    if (!c || !a || c->n < 1) { set_error("context has no system / NULL pointer"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_get_accelerations(r_, a));
    return download_aos3(c, c->acc.as<double>(), a, true);
}

int grav_b200_ctx_synchronize(grav_b200_ctx *c)
{
    if (!c) { set_error("NULL context"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_synchronize(r_));
    GB_CUDA(cudaSetDevice(c->device));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return bh_check(c);
}

int grav_b200_ctx_last_timing_ms(grav_b200_ctx *c, int stage, float *ms)
{
    if (!c || !ms || stage < 0 || stage >= ST_COUNT) { set_error("bad timing query"); return GRAV_B200_EINVAL; }
    *ms = 0.0f;
    if (!c->ev_valid[stage]) return GRAV_B200_OK;
    GB_CUDA(cudaSetDevice(c->device));
    GB_CUDA(cudaEventSynchronize(c->ev[2 * stage + 1]));
    GB_CUDA(cudaEventElapsedTime(ms, c->ev[2 * stage], c->ev[2 * stage + 1]));
    return GRAV_B200_OK;
}

int grav_b200_ctx_event_record(grav_b200_ctx *c, int slot)
{
    if (!c || slot < 0 || slot >= 8) { set_error("bad event slot"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    GB_CUDA(cudaEventRecord(c->user_ev[slot], c->stream));
    return GRAV_B200_OK;
}

int grav_b200_ctx_event_elapsed_ms(grav_b200_ctx *c, int a, int b, float *ms)
{
    if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) { set_error("bad event slot"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    GB_CUDA(cudaEventSynchronize(c->user_ev[b]));
    GB_CUDA(cudaEventElapsedTime(ms, c->user_ev[a], c->user_ev[b]));
    return GRAV_B200_OK;
}

int grav_b200_ctx_direct_sum_path(grav_b200_ctx *c, int *pair_once, int *equal_mass)
{
    if (!c || !pair_once || !equal_mass) { set_error("NULL pointer"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(c->device));
    *pair_once = c->last_ds_sym;
    *equal_mass = 0;
    if (c->last_ds_sym && c->sym_flag.p) {
        int flag = 0;
        GB_CUDA(cudaMemcpyAsync(&flag, c->sym_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        *equal_mass = flag != 0;
    }
    return GRAV_B200_OK;
}

int grav_b200_ctx_flush_l2(grav_b200_ctx *c)
{
    if (!c) { set_error("NULL context"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_flush_l2(r_));
    GB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)256 << 20;
    GB_TRY(c->l2_flush.reserve(bytes));
    GB_CUDA(cudaMemsetAsync(c->l2_flush.p, 0, bytes, c->stream));
    return GRAV_B200_OK;
}

int grav_b200_ctx_mark_positions_sharded(grav_b200_ctx *c)
{
    if (!c) { set_error("NULL context"); return GRAV_B200_EINVAL; }
    GB_TEAM(c, grav_b200_ctx_mark_positions_sharded(r_));
    if (c->world > 1) c->posm_gathered = false;
    return GRAV_B200_OK;
}

int grav_b200_host_register(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes) { set_error("bad host range"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return GRAV_B200_OK;
}

int grav_b200_host_unregister(void *ptr)
{
    GB_CUDA(cudaHostUnregister(ptr));
    return GRAV_B200_OK;
}

// ---- host-pointer one-shots ---------------------------------------------------------------
static int one_shot(double *a, int n, const double *x, const double *m, double G, int method, double eps, double theta, int leaf)
{
    GB_TRY(check_sys(a, n, x, m));
    std::lock_guard<std::mutex> lk(g_mu);
    grav_b200_ctx *c;
    GB_TRY(default_ctx(&c));
    GB_TRY(set_system_no_velocities(c, n, x, m, G));
    for (;;) {
        GB_TRY(grav_b200_ctx_acceleration(c, method, eps, theta, leaf));
        const int rc = grav_b200_ctx_get_accelerations(c, a);
        // Barnes-Hut builds are queued without waiting for their sizes; a tree that outgrew its buffers is reported by
        // the download's synchronisation: rebuild with twice the room (the system is still resident)
        if (rc != GRAV_B200_ETREE || c->tree.slack >= 16) return rc;
        if (c->team) GB_TRY(team_run(c, [](grav_b200_ctx *r) { r->tree.slack *= 2; return GRAV_B200_OK; }));
        else c->tree.slack *= 2;
    }
}

int grav_b200_acceleration_pairwise(double *a, int n, const double *x, const double *m, double G, double eps)
{
    if (n >= 1 && n <= 256 && a && x && m && eps >= 0.0) {   // small systems: one launch, zero-copy (direct_sum.cu)
        std::lock_guard<std::mutex> lk(g_mu);
        grav_b200_ctx *c;
        GB_TRY(default_ctx(&c));
        // a resident kernel answers from a mailbox in pinned host memory (small_mailbox.cu); GRAV_B200_SMALL_MAILBOX=0 goes
        // back to one launch + one synchronisation per call
        static const bool use_mailbox = !(getenv("GRAV_B200_SMALL_MAILBOX") && atoi(getenv("GRAV_B200_SMALL_MAILBOX")) == 0);
        return use_mailbox ? mailbox_pairwise(c, a, n, x, m, G, eps) : direct_sum_small_host(c, a, n, x, m, G, eps);
    }
    return one_shot(a, n, x, m, G, GRAV_B200_METHOD_PAIRWISE, eps, 0.0, 1);
}
int grav_b200_acceleration_massless(double *a, int n, const double *x, const double *m, double G, double eps)
{
    return one_shot(a, n, x, m, G, GRAV_B200_METHOD_MASSLESS, eps, 0.0, 1);
}
int grav_b200_acceleration_barnes_hut(double *a, int n, const double *x, const double *m, double G, double eps,
                                      double theta, int leaf)
{
    return one_shot(a, n, x, m, G, GRAV_B200_METHOD_BARNES_HUT, eps, theta, leaf);
}

int grav_b200_compute_energy(double *energy, int n, const double *x, const double *v, const double *m, double G)
{
    if (!energy || !v) { set_error("NULL array pointer"); return GRAV_B200_EINVAL; }
    GB_TRY(check_sys(energy, n, x, m));
    std::lock_guard<std::mutex> lk(g_mu);
    grav_b200_ctx *c;
    GB_TRY(default_ctx(&c));
    GB_TRY(grav_b200_ctx_set_system(c, n, x, v, m, G));
    return grav_b200_ctx_energy(c, energy);
}

static int whfast_one_shot(double *a, int n, const double *x, const double *m, double G, const double *jx,
                           const double *eta, double eps, bool massless)
{
    GB_TRY(check_sys(a, n, x, m));
    if (!jx || !eta) { set_error("NULL jacobi_x / eta"); return GRAV_B200_EINVAL; }
    std::lock_guard<std::mutex> lk(g_mu);
    grav_b200_ctx *c;
    GB_TRY(default_ctx(&c));
    GB_TRY(set_system_impl(c, n, x, nullptr, m, G, false));
    const size_t b3 = sizeof(double) * 3 * (size_t)n;
    GB_TRY(c->stage_c.reserve(b3));
    GB_TRY(c->stage_d.reserve(sizeof(double) * (size_t)n));
    GB_CUDA(cudaMemcpyAsync(c->stage_c.p, jx, b3, cudaMemcpyHostToDevice, c->stream));
    GB_CUDA(cudaMemcpyAsync(c->stage_d.p, eta, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    // the reference leaves some entries of a[] untouched (particle 0 and, in the massless variant, the first
    // massive particle: loops start at 1, src/integrator_whfast.c:858,1012,1131): start from the caller's values
    GB_CUDA(cudaMemcpyAsync(c->acc.p, a, b3, cudaMemcpyHostToDevice, c->stream));
    GB_TRY(whfast_accel(c, c->stage_c.as<double>(), c->stage_d.as<double>(), eps, massless));
    return grav_b200_ctx_get_accelerations(c, a);
}

int grav_b200_whfast_acceleration_pairwise(double *a, int n, const double *x, const double *m, double G,
                                           const double *jx, const double *eta, double eps)
{
    return whfast_one_shot(a, n, x, m, G, jx, eta, eps, false);
}
int grav_b200_whfast_acceleration_massless(double *a, int n, const double *x, const double *m, double G,
                                           const double *jx, const double *eta, double eps)
{
    return whfast_one_shot(a, n, x, m, G, jx, eta, eps, true);
}

}  // extern "C"
