// Internal declarations shared by the .cu files of libgrav_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/grav_b200.h"

struct ncclComm;

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

// What the tree walk reads per node: one 64-byte record.  The first 48 bytes are needed at every visit,
// the mass only when the node is accepted.
// `next` is the "rope": the node that follows in depth-first order when this node's subtree is skipped
// (next sibling, else the parent's rope; -1 ends the walk), which makes the walk stackless.
// Stored as two planes of 32-byte records (geo[M] then topo[M] in one buffer) rather than one 64-byte record: the
// walk is bound by L1 wavefronts = distinct 128-byte lines per load instruction, and lanes that are a few siblings
// apart then share lines (4 records per line instead of 2).
struct WalkGeo {
    double cx, cy, cz;
    long long kq;      // key the inclusion test compares against (mode dependent)
};
struct WalkTopo {
    int fc;            // id of the first child, -1 for a leaf
    int next;
    int first;         // sorted position of the node's first particle
    int level_count;   // level << 26 | particles in the node (N <= 2^24 < 2^26)
    double mass;
    long long pad;
};
constexpr int WALK_COUNT_BITS = 26;
static_assert(sizeof(WalkGeo) == 32 && sizeof(WalkTopo) == 32, "walk record layout");

// Device-side linear octree (layout in DESIGN.md "BH data layout")
struct DevTree {
    int n = 0;
    int num_nodes = 0;           // valid after build (host copy)
    int num_expanded = 0;
    int max_level = 0;           // number of levels that hold expanded nodes
    int level_off[24] = {};      // expanded-node records of level l are [level_off[l], level_off[l+1])
    double box_width = 0.0;
    DevBuf keys_unsorted, keys, perm;       // int64[n], int64[n], int[n]
    DevBuf keys_tmp, perm_tmp, hist;        // radix sort ping-pong + histograms
    DevBuf bbox;                            // double[8]: min xyz, max xyz (ordered-int encoded), then center xyz + width as double[4]
    DevBuf exp_rec;                         // expanded-node records in BFS order
    DevBuf wsum, wscan;                     // int[n+1] children-per-start-position and its exclusive scan
    DevBuf scan_tmp;
    DevBuf fc;                              // int[num_expanded] first-child id per expanded record
    DevBuf node_np, node_nch, node_first, node_fc;      // int[num_nodes]
    DevBuf node_mass, node_cx, node_cy, node_cz;        // double[num_nodes]
    DevBuf node_mtd;                        // double[3*num_nodes] mass-weighted position sums
    DevBuf node_walk;                       // packed 64-byte walk records
    DevBuf posm_sorted;                     // double4[n]: posm in sorted (Morton) order, for the walk's leaf sums
    DevBuf ki, tord;                        // walk keys and target order (optional walk-key grouping)
    DevBuf counters;                        // misc device ints
};

}  // namespace gb

struct grav_b200_ctx {
    int device = 0, rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    ncclComm *comm = nullptr;

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

    int bh_mode = 0;
    cudaEvent_t user_ev[8] = {};
    gb::DevBuf l2_flush;
    double *small_pinned = nullptr;   // mapped pinned host staging of the small-N one-shot path (in: 4n doubles, out: 3n)
    cudaEvent_t ev[2 * gb::ST_COUNT] = {};
    bool ev_valid[gb::ST_COUNT] = {};
};

namespace gb {
// direct_sum.cu
int direct_sum_pairwise(grav_b200_ctx *c, double eps);
int direct_sum_massless(grav_b200_ctx *c, double eps);
int direct_sum_small_host(grav_b200_ctx *c, double *a, int n, const double *x, const double *m, double G, double eps);
// pack.cu
int pack_posm(grav_b200_ctx *c, const double *d_x_aos, const double *d_m);   // device AoS -> posm
int pack_positions(grav_b200_ctx *c, const double *d_x_aos);
int unpack_positions(grav_b200_ctx *c, double *d_x_aos);
// whfast.cu
int whfast_accel(grav_b200_ctx *c, const double *d_jacobi_x, const double *d_eta, double eps, bool massless);
// bh_*.cu
int bh_build(grav_b200_ctx *c, int max_leaf, const double *box_center, double box_width);
int bh_walk(grav_b200_ctx *c, double eps, double theta);
// whfast_resident.cu
void whfast_state_free(grav_b200_ctx *c);
// comm.cu
int comm_init(grav_b200_ctx *c, const void *uid);
void comm_destroy(grav_b200_ctx *c);
int comm_allgather_posm(grav_b200_ctx *c);
int comm_allgather_aos3(grav_b200_ctx *c, double *d_buf);   // gathers owned [3*lo,3*hi) slices in place
int comm_allreduce_sum(grav_b200_ctx *c, double *d_val, int count);
// integrate.cu
int leapfrog_kick(grav_b200_ctx *c, double dt_half_or_full);
int leapfrog_drift(grav_b200_ctx *c, double dt);
int synced_velocities(grav_b200_ctx *c, double **d_out);
// timing helpers
inline void stage_begin(grav_b200_ctx *c, int st) { cudaEventRecord(c->ev[2 * st], c->stream); }
inline void stage_end(grav_b200_ctx *c, int st) { cudaEventRecord(c->ev[2 * st + 1], c->stream); c->ev_valid[st] = true; }
}  // namespace gb
