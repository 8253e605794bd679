// Internal declarations shared by the .cu files of libgrav_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/grav_b200.h"

struct ncclComm;
namespace gb { struct Team; }

namespace gb {

// ---- error plumbing -----------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define GB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return gb::cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)
#define GB_TRY(call)                                  \
    do {                                              \
        int rc_ = (call);                             \
        if (rc_ != GRAV_B200_OK) return rc_;          \
    } while (0)
#define GB_LAUNCH_CHECK() GB_CUDA(cudaGetLastError())

extern int64_t g_launch_count;   // kernels launched by this library
extern uint64_t g_alloc_generation;   // device-buffer (re)allocations so far (graph caches compare it)
inline void count_launch(int k = 1) { __atomic_fetch_add(&g_launch_count, (int64_t)k, __ATOMIC_RELAXED); }

// ---- geometry of the direct-sum kernel ----------------------------------------------
constexpr int DS_BLOCK = 256;   // threads per CTA
constexpr int DS_TJ    = 256;   // sources per shared-memory tile (one per thread on load)
constexpr int SRC_PAD  = DS_TJ; // particle buffers are zero-padded to a multiple of this

// ---- m * r^-3 for the direct sum and the fast Barnes-Hut evaluation -------------------------
// The seed y0 = rsqrt.approx.ftz.f64(r2) only looks at the high word of r2 (|e| <~ 2^-20 with e = 1 - r2*y0^2);
// r2^-3/2 = y0^3 (1-e)^-3/2 = y0^3 (1 + e(3/2 + 15/8 e) + 35/16 e^3 + ...), and the dropped term is < 2^-58.
#ifdef __CUDACC__
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// m_j * r2^(-3/2), 7 FP64-pipe instructions + 1 MUFU.
__device__ __forceinline__ double inv_r3_times_m(double r2, double mj)
{
    const double y = rsqrt_seed(r2);
    const double t = y * y;                 // exact: y has <= 24 significant bits... (<=53 anyway)
    const double e = fma(-r2, t, 1.0);      // 1 - r2*y^2, |e| <~ 2^-20
    const double my = mj * y;
    const double y3m = my * t;              // m * y^3
    const double p = fma(1.875, e, 1.5);    // 3/2 + 15/8 e
    const double q = e * p;
    // m*y^3*(1 + q).  Written so the DFMA has only two distinct register sources: on sm_100 a
    // DFMA with three distinct register operands occupies the FP64 pipe for 3 cycles instead of 2
    // (measured, scratch/fp64_micro.cu; DESIGN.md "FP64 pipe").
    return fma(q, y3m, y3m);
}
#endif

// A growable device buffer (never shrinks; contents undefined after a grow).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T *as() const { return (T *)p; }
};

// Timing stages (grav_b200_ctx_last_timing_ms)
enum { ST_TOTAL = 0, ST_GATHER = 1, ST_FORCE = 2, ST_MORTON = 3, ST_SORT = 4, ST_BUILD = 5, ST_COUNT = 6 };

// What the tree walk reads per node: two 32-byte records (geometry, topology), stored as two planes
// (geo[m_cap] then topo[m_cap] in one buffer) rather than one 64-byte record: the walk is bound by L1 wavefronts =
// distinct 128-byte lines per load instruction, and lanes that are a few siblings apart then share lines
// (4 records per line instead of 2).
// `next` is the "rope": the node that follows in depth-first order when this node's subtree is skipped
// (next sibling, else the parent's rope; -1 ends the walk), which makes the per-lane walk stackless.
struct WalkGeo {
    double cx, cy, cz;
    long long kq;      // key the inclusion test compares against (mode dependent)
};
struct WalkTopo {
    unsigned fcn;      // first child id << 4 | number of children; the low 4 bits are 0 for a leaf.  Also the walk's stack entry.
    int next;
    int first;         // sorted position of the node's first particle
    unsigned shift_count;   // 3 (21 - level) << 26 | particles in the node (N <= 2^24 < 2^26): the key shift of the inclusion test
    double mass;
    double cell2;      // (box_length / (2 << level))^2, the left-hand side of the opening test for this node's level
};
constexpr int WALK_COUNT_BITS = 26;
static_assert(sizeof(WalkGeo) == 32 && sizeof(WalkTopo) == 32, "walk record layout");

// Device-side bookkeeping of one tree build.  Everything the later stages need to know about the sizes the earlier
// stages produced lives HERE, on the device, so the host never waits for a count (no cudaStreamSynchronize between the
// first and the last kernel of a force evaluation).  A copy lands in pinned host memory at the end of the build and is
// looked at by the next call that synchronises anyway (download, energy, ctx_synchronize).
struct TreeMeta {
    int lvl_cnt[24];    // expanded records per level
    int lvl_off[24];    // first record of each level
    int num_expanded;   // records in total
    int num_nodes;      // M = 1 + children of all expanded nodes
    int levels;         // levels that hold expanded nodes
    int overflow;       // bit 0: more expanded nodes than exp_rec holds, bit 1: more nodes than the node planes hold
    unsigned coop_barrier;   // arrival counter of the single-launch build's grid barrier
    unsigned pad_;
    double box_width;
    double cell2[24];   // (box_length / (2 << level))^2 per child level, src/acceleration_barnes_hut.c:157,162
};
constexpr int TREE_OVERFLOW_EXPANDED = 1, TREE_OVERFLOW_NODES = 2;

// Device-side linear octree (layout in DESIGN.md "BH data layout")
struct DevTree {
    int n = 0;
    int ne_cap = 0, m_cap = 0;   // capacities of exp_rec and of the node planes in the current build
    int slack = 1;               // capacity multiplier (doubled after an overflow)
    bool built = false;          // a build has been queued since the last successful check
    DevBuf keys_unsorted, keys, perm;       // int64[n], int64[n], int[n]
    DevBuf keys_tmp, perm_tmp, hist;        // radix sort ping-pong + histograms
    DevBuf bbox;                            // double[8]: min xyz, max xyz (ordered-int encoded), then center xyz + width as double[4]
    DevBuf exp_rec;                         // expanded-node records, level by level
    DevBuf wsum, wscan;                     // int[n+1] children-per-start-position and its exclusive scan
    DevBuf scan_tmp;
    DevBuf meta;                            // TreeMeta
    TreeMeta *h_meta = nullptr;             // pinned host mirror (valid after the stream reached the end of the build)
    DevBuf node_walk;                       // geo[m_cap] then topo[m_cap]
    DevBuf node_mtd;                        // double[3*m_cap] mass-weighted position sums
    DevBuf posm_sorted;                     // double4[n]: posm in sorted (Morton) order, for the walk's leaf sums
    DevBuf ki, tord;                        // walk keys and target order (optional walk-key grouping of the per-lane walk)
    DevBuf walk_out;                        // multi-GPU: per-rank walk results in slot order + the gathered copy
    DevBuf xport;                           // construct_octree(): the LinearOctree arrays, derived from the planes on demand
    WalkGeo *geo() const { return node_walk.as<WalkGeo>(); }
    WalkTopo *topo() const { return reinterpret_cast<WalkTopo *>(node_walk.as<WalkGeo>() + m_cap); }
};

}  // namespace gb

struct grav_b200_ctx {
    int device = 0, rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    ncclComm *comm = nullptr;
    gb::Team *team = nullptr;   // member of an in-process device team (team.cu): public entries called on the leader fan out

    int n = 0;            // particles
    int n_pad = 0;        // n rounded up to SRC_PAD
    int lo = 0, hi = 0;   // owned target range
    double G = 0.0;
    bool posm_gathered = true;   // false when only the owned shard of posm is current

    gb::DevBuf posm;      // double4[n_pad] packed (x,y,z,m); padding is all-zero
    gb::DevBuf vel, acc;  // double[3n] AoS
    gb::DevBuf xcomp, vcomp;      // compensated-summation error terms, double[3n]
    gb::DevBuf stage_a, stage_b, stage_c, stage_d;  // H2D/D2H staging
    gb::DevBuf partials;  // direct-sum split-segment partial sums
    gb::DevBuf sym_priv;  // pair-once direct sum: one private accumulation array per CTA (direct_sum_sym.cu), all zero between calls
    gb::DevBuf sym_flag;  // pair-once direct sum: equal-mass flag + the common mass
    bool sym_priv_clean = false, sym_attr_set = false;
    long long sym_layout = -1;    // (n, CTAs, first CTA) the private arrays were last used with
    bool sym_eqm_valid = false;   // sym_flag describes the resident masses (reset with mlist_valid whenever the masses change)
    int last_ds_sym = 0;  // the last pairwise force evaluation took the pair-once path
    gb::DevBuf misc;      // small scratch (reductions)
    // massless method scratch
    gb::DevBuf msrc, msrc_id, msrc_altm, mflag, mrank;
    bool mlist_valid = false;     // mflag / mrank / the massive count describe the resident masses (reset by set_system)
    int mlist_count = 0;
    gb::DevTree tree;

    // leapfrog bookkeeping
    int lf_method = 0, lf_leaf = 1;
    double lf_eps = 0.0, lf_theta = 1.0, lf_dt = 0.0;
    bool lf_ready = false;
    int fixed_integrator = 0;     // Euler / Euler-Cromer / RK4 on the resident state (integrate.cu)
    gb::DevBuf rk_buf;            // RK4: x_0, v_0, xk1..3, vk1..3

    void *wh = nullptr;   // gb::WhfastState (whfast_resident.cu)
    // per-device launch state of the radix sort (bh_sort.cu)
    int sort_resident_ctas = -1;
    int tree_coop_ctas = -1;          // co-resident CTAs of the single-launch tree build on this device (bh_build.cu)
    bool sort_attr_scatter = false, sort_attr_onesweep = false;

    int bh_mode = 0;
    int bh_exact = 0;             // 1: bit-identical per-lane walk (GRAV_B200_BH_EXACT), 0: warp-cooperative walk, <= 1e-12
    cudaEvent_t user_ev[8] = {};
    gb::DevBuf l2_flush;
    void *mailbox = nullptr;          // gb::MailboxState: resident small-system kernel (small_mailbox.cu)
    double *small_pinned = nullptr;   // mapped pinned host staging of the small-N one-shot path (in: 4n doubles, out: 3n)
    cudaEvent_t ev[2 * gb::ST_COUNT] = {};
    bool ev_valid[gb::ST_COUNT] = {};
};

namespace gb {
// direct_sum.cu
int direct_sum_pairwise(grav_b200_ctx *c, double eps);
int direct_sum_massless(grav_b200_ctx *c, double eps);
// direct_sum_sym.cu
constexpr int GRAV_B200_ENOMEM_SYM = -1000;   // internal: the private arrays of the pair-once kernel could not be allocated
bool direct_sum_sym_wanted(const grav_b200_ctx *c);
int direct_sum_pairwise_sym(grav_b200_ctx *c, double eps);
int direct_sum_small_host(grav_b200_ctx *c, double *a, int n, const double *x, const double *m, double G, double eps);
// small_mailbox.cu
int mailbox_pairwise(grav_b200_ctx *c, double *a, int n, const double *x, const double *m, double G, double eps);
int mailbox_stop(grav_b200_ctx *c);
void mailbox_free(grav_b200_ctx *c);
// pack.cu
int pack_posm(grav_b200_ctx *c, const double *d_x_aos, const double *d_m);   // device AoS -> posm
int pack_positions(grav_b200_ctx *c, const double *d_x_aos);
int unpack_positions(grav_b200_ctx *c, double *d_x_aos);
// whfast.cu
int whfast_accel(grav_b200_ctx *c, const double *d_jacobi_x, const double *d_eta, double eps, bool massless);
// bh_*.cu
int bh_build(grav_b200_ctx *c, int max_leaf, const double *box_center, double box_width);
int bh_walk(grav_b200_ctx *c, double eps, double theta);
int bh_build_checked(grav_b200_ctx *c, int max_leaf, const double *box_center, double box_width);   // build + sync + retry
int bh_check(grav_b200_ctx *c);   // after a host sync: did the last queued build fit its buffers?  (GRAV_B200_ETREE if not)
// whfast_resident.cu
void whfast_state_free(grav_b200_ctx *c);
// comm.cu
int comm_init(grav_b200_ctx *c, const void *uid);
void comm_destroy(grav_b200_ctx *c);
int comm_allgather_posm(grav_b200_ctx *c);
int comm_allgather_aos3(grav_b200_ctx *c, double *d_buf);   // gathers owned [3*lo,3*hi) slices in place
int comm_allreduce_sum(grav_b200_ctx *c, double *d_val, int count);
int comm_allgather_equal(grav_b200_ctx *c, double *d_base, size_t count_per_rank);   // in place: rank r's part at d_base + r * count
// integrate.cu
int leapfrog_kick(grav_b200_ctx *c, double dt_half_or_full);
int leapfrog_drift(grav_b200_ctx *c, double dt);
int synced_velocities(grav_b200_ctx *c, double **d_out);
// team.cu
bool team_active(const grav_b200_ctx *c);        // c leads (or belongs to) a team and this thread is not already inside a team call
int team_run(grav_b200_ctx *c, const std::function<int(grav_b200_ctx *)> &f);
void team_destroy(grav_b200_ctx *leader);
// first statement of a public context entry: forward the call to every member of the team (r_ is the member's context)
#define GB_TEAM(c, expr)                                                                              \
    do {                                                                                              \
        if (gb::team_active(c)) return gb::team_run((c), [=](grav_b200_ctx *r_) -> int { return (expr); }); \
    } while (0)
// timing helpers
inline void stage_begin(grav_b200_ctx *c, int st) { cudaEventRecord(c->ev[2 * st], c->stream); }
inline void stage_end(grav_b200_ctx *c, int st) { cudaEventRecord(c->ev[2 * st + 1], c->stream); c->ev_valid[st] = true; }
}  // namespace gb
