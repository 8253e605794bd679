// Barnes-Hut stage BH-8 on the device: the tree walk.
//
// Reference: helper_compute_acceleration (src/acceleration_barnes_hut.c:78-248): one thread per particle
// (in Morton order) runs a depth-first walk with a private stack.  Here a WARP walks the tree for 32
// Morton-adjacent targets at once: control flow, the stack (shared memory, <= 22 frames) and the node
// records (one 64-byte load, same address for all lanes) are warp-uniform; each frame carries the mask of
// lanes that still need the subtree.  A lane that accepts a node leaves the mask for that subtree, so every
// lane sees exactly the nodes, in exactly the depth-first order, that the reference's per-particle walk sees.
// All arithmetic that feeds a decision or the result uses IEEE operations without FMA contraction
// (__dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn) in the reference's order, so in reference mode the output is
// bit-identical to the x86-64 reference build.
//
// Modes (grav_b200_set_bh_mode):
//   reference  bug-for-bug: the inclusion test compares keys fetched from the SORTED key array with ORIGINAL
//              particle ids (:120,:145), and the opening test is applied to leaves too, whose mass is 0, so
//              a far leaf is silently dropped (:150-176 precedes :181).
//   fixed      inclusion from the particle's / node's own key; leaves are never approximated.
#include "internal.cuh"

namespace gb {

constexpr int MAX_LEVEL = 21;
constexpr int WALK_WARPS = 8;

struct WalkNode {     // must match bh_build.cu
    double mass, cx, cy, cz;
    long long kq;
    int first, count, nch, fc;
    long long pad;
};

struct WalkArgs {
    const WalkNode *nodes;
    const long long *K;      // sorted keys
    const int *perm;         // sorted position -> particle id
    const double4 *posm;
    int p_lo, p_hi;          // sorted positions handled by this launch
    double G, eps2, theta2;
    double cell2[MAX_LEVEL + 2];   // (box_length / (2 << level))^2 per child level
    double *acc;             // AoS [3n] by particle id
};

__device__ __forceinline__ WalkNode load_node(const WalkNode *p)
{
    // four 16-byte read-only loads; every lane of the warp reads the same address (one transaction each)
    const int4 *q = reinterpret_cast<const int4 *>(p);
    int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    WalkNode w;
    w.mass = __hiloint2double(a.y, a.x);
    w.cx = __hiloint2double(a.w, a.z);
    w.cy = __hiloint2double(b.y, b.x);
    w.cz = __hiloint2double(b.w, b.z);
    w.kq = ((long long)c.y << 32) | (unsigned)c.x;
    w.first = c.z;
    w.count = c.w;
    const int2 d = __ldg(reinterpret_cast<const int2 *>(p) + 6);
    w.nch = d.x;
    w.fc = d.y;
    w.pad = 0;
    return w;
}

template <bool FIXED>
__global__ void __launch_bounds__(WALK_WARPS * 32) walk_kernel(const WalkArgs a)
{
    // saved frames of the enclosing levels; the current frame lives in (warp-uniform) registers
    __shared__ int s_fc[WALK_WARPS][MAX_LEVEL + 1], s_n[WALK_WARPS][MAX_LEVEL + 1], s_j[WALK_WARPS][MAX_LEVEL + 1];
    __shared__ unsigned s_mask[WALK_WARPS][MAX_LEVEL + 1];
    __shared__ double s_cell2[MAX_LEVEL + 2];
    if (threadIdx.x < MAX_LEVEL + 2) s_cell2[threadIdx.x] = a.cell2[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = a.p_lo + (blockIdx.x * WALK_WARPS + warp) * 32 + lane;
    const bool valid = p < a.p_hi;
    int idx = -1;
    double xi = 0.0, yi = 0.0, zi = 0.0;
    long long ki = 0;
    if (valid) {
        idx = a.perm[p];
        const double4 q = a.posm[idx];
        xi = q.x; yi = q.y; zi = q.z;
        ki = FIXED ? a.K[p] : a.K[idx];
    }
    double ax = 0.0, ay = 0.0, az = 0.0;
    const unsigned all = __ballot_sync(0xffffffffu, valid);
    if (all == 0u) return;

    int depth = 0;
    int cur_fc, cur_n, cur_j = 0;
    unsigned cur_mask = all;
    {
        const WalkNode root = load_node(a.nodes);
        cur_fc = root.fc;
        cur_n = root.nch;
    }
    while (true) {
        if (cur_j >= cur_n) {           // frame exhausted: pop
            if (depth == 0) break;
            depth--;
            cur_fc = s_fc[warp][depth]; cur_n = s_n[warp][depth]; cur_j = s_j[warp][depth]; cur_mask = s_mask[warp][depth];
            continue;
        }
        const int c = cur_fc + cur_j;
        cur_j++;
        const int level = depth + 1;
        const int shift = 3 * (MAX_LEVEL - level);
        const WalkNode nd = load_node(a.nodes + c);
        const bool leaf = nd.nch <= 0;
        bool need = false;
        if ((cur_mask >> lane) & 1u) {
            const bool inside = (ki >> shift) == (nd.kq >> shift);
            bool accepted = false;
            if (FIXED ? (!inside && !leaf) : !inside) {
                const double rx = __dsub_rn(xi, nd.cx), ry = __dsub_rn(yi, nd.cy), rz = __dsub_rn(zi, nd.cz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                if (s_cell2[level] < __dmul_rn(a.theta2, d2)) {
                    const double r = __dsqrt_rn(__dadd_rn(d2, a.eps2));
                    const double f = __ddiv_rn(__dmul_rn(a.G, nd.mass), __dmul_rn(__dmul_rn(r, r), r));
                    ax = __dsub_rn(ax, __dmul_rn(f, rx));
                    ay = __dsub_rn(ay, __dmul_rn(f, ry));
                    az = __dsub_rn(az, __dmul_rn(f, rz));
                    accepted = true;
                }
            }
            need = !accepted;
        }
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask == 0u) continue;
        if (leaf) {
            // direct sum over the leaf's particles (sorted order), skipping the target itself
            for (int k = 0; k < nd.count; k++) {
                const int jdx = __ldg(a.perm + nd.first + k);
                const double4 q = a.posm[jdx];
                if (need && jdx != idx) {
                    const double rx = __dsub_rn(xi, q.x), ry = __dsub_rn(yi, q.y), rz = __dsub_rn(zi, q.z);
                    const double d2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)), a.eps2);
                    const double r = __dsqrt_rn(d2);
                    const double f = __ddiv_rn(__dmul_rn(a.G, q.w), __dmul_rn(__dmul_rn(r, r), r));
                    ax = __dsub_rn(ax, __dmul_rn(f, rx));
                    ay = __dsub_rn(ay, __dmul_rn(f, ry));
                    az = __dsub_rn(az, __dmul_rn(f, rz));
                }
            }
        } else {
            // push: every lane stores the same (uniform) values; only this warp ever touches its rows, and the
            // reads on pop are separated from these writes by the ballot above / program order within a lane
            s_fc[warp][depth] = cur_fc; s_n[warp][depth] = cur_n; s_j[warp][depth] = cur_j; s_mask[warp][depth] = cur_mask;
            depth++;
            cur_fc = nd.fc; cur_n = nd.nch; cur_j = 0; cur_mask = need_mask;
        }
    }
    if (valid) {
        a.acc[3 * (size_t)idx + 0] = ax;
        a.acc[3 * (size_t)idx + 1] = ay;
        a.acc[3 * (size_t)idx + 2] = az;
    }
}

int bh_pack_walk_nodes(grav_b200_ctx *c);

int bh_walk(grav_b200_ctx *c, double eps, double theta)
{
    DevTree &t = c->tree;
    GB_TRY(bh_pack_walk_nodes(c));
    WalkArgs a{};
    a.nodes = t.node_walk.as<WalkNode>();
    a.K = t.keys.as<long long>();
    a.perm = t.perm.as<int>();
    a.posm = c->posm.as<double4>();
    // ranks share the walk by sorted position (Morton-contiguous), not by particle id
    a.p_lo = (int)(((long long)c->rank * c->n) / c->world);
    a.p_hi = (int)(((long long)(c->rank + 1) * c->n) / c->world);
    a.G = c->G;
    a.eps2 = eps * eps;
    a.theta2 = theta * theta;
    const double box_length = t.box_width * 2.0;          // src/acceleration_barnes_hut.c:100
    for (int level = 0; level <= MAX_LEVEL; level++) {
        const double bl = box_length / (double)(2 << level);   // :157
        a.cell2[level] = bl * bl;                              // :162 (left-hand side)
    }
    a.acc = c->acc.as<double>();
    if (c->world > 1) GB_CUDA(cudaMemsetAsync(a.acc, 0, sizeof(double) * 3 * (size_t)c->n, c->stream));
    const int npos = a.p_hi - a.p_lo;
    if (npos > 0) {
        const int blocks = (npos + WALK_WARPS * 32 - 1) / (WALK_WARPS * 32);
        if (c->bh_mode == GRAV_B200_BH_FIXED) walk_kernel<true><<<blocks, WALK_WARPS * 32, 0, c->stream>>>(a);
        else walk_kernel<false><<<blocks, WALK_WARPS * 32, 0, c->stream>>>(a);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    if (c->world > 1) GB_TRY(comm_allreduce_sum(c, a.acc, 3 * c->n));   // disjoint targets + zeros: exact
    return GRAV_B200_OK;
}

}  // namespace gb
