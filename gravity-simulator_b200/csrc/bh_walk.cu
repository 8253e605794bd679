// Barnes-Hut stage BH-8 on the device: the tree walk.
//
// Reference: helper_compute_acceleration (src/acceleration_barnes_hut.c:78-248): one CPU thread per particle
// (in Morton order) runs a depth-first walk with a private 22-frame stack.
//
// Here: one GPU thread per target, targets taken in Morton order so the 32 lanes of a warp follow nearly the
// same path (their node loads then hit the same sectors), and the walk is STACKLESS: every node record carries
// a rope (`next`, the node that follows in depth-first order when the subtree is skipped), so a visit is
//     accept or leaf  ->  node = next          open  ->  node = first child
// and each lane advances through exactly the nodes, in exactly the order, the reference visits for that
// particle.  A first version shared one traversal per warp (union of the lanes' trees, per-lane masks); the ncu
// profile (profiles/r1_walk_warp_union.txt) showed it issue-bound with 7 of 32 lanes active in the accept path
// and 4.4x more node visits than a single particle needs, because the opening test separates Morton
// neighbours.  Independent lanes remove both costs.
//
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
constexpr int WALK_BLOCK = 128;
constexpr int WALK_POOL = WALK_BLOCK;       // targets per CTA; > WALK_BLOCK enables per-lane work fetching (measured slower: 46 vs 41 ms, N=2^20 Plummer)

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

struct NodeRec {   // the 48 bytes of a WalkNode every visit needs, as loaded (three 16-byte read-only loads)
    int4 A, B, C;
};
__device__ __forceinline__ NodeRec load_rec(const WalkNode *nodes, int node)
{
    const int4 *q = reinterpret_cast<const int4 *>(nodes + node);
    NodeRec r;
    r.A = __ldg(q); r.B = __ldg(q + 1); r.C = __ldg(q + 2);
    return r;
}

// Per-lane state machine.  Each trip of the loop handles ONE item for the lane -- a node visit, or the next
// particle of a leaf that is being summed directly -- in three steps that are the same code for every lane:
//   decide    opening test on the current node record -> the next node id, and possibly a source (R, d2, m)
//   prefetch  the record of the next node is requested BEFORE the arithmetic below, so its L1/L2 latency
//             overlaps the sqrt/div sequence
//   evaluate  f = G m / r^3 and the three accumulator updates, shared by accepted nodes and leaf particles
//             (both are sqrt(((rx^2+ry^2)+rz^2)+eps^2) in the reference, :164-171 and :198-213)
template <bool FIXED>
__global__ void __launch_bounds__(WALK_BLOCK) walk_kernel(const WalkArgs a)
{
    __shared__ double s_cell2[MAX_LEVEL + 2];
    __shared__ int s_next;   // next unclaimed target of this CTA's pool
    if (threadIdx.x < MAX_LEVEL + 2) s_cell2[threadIdx.x] = a.cell2[threadIdx.x];
    // Work distribution: a CTA owns a pool of WALK_POOL Morton-consecutive targets.  A lane that finishes its
    // particle claims the next one of the pool (shared-memory counter), so lanes with short walks do not idle
    // behind lanes with long ones.  In reference mode walk lengths differ wildly even between Morton
    // neighbours, because the inclusion test uses an unrelated particle's key (see file header).
    const int pool_lo = a.p_lo + blockIdx.x * WALK_POOL;
    const int pool_hi = min(pool_lo + WALK_POOL, a.p_hi);
    if (threadIdx.x == 0) s_next = pool_lo + WALK_BLOCK;
    __syncthreads();
    int p = pool_lo + threadIdx.x;
    const int root_fc = __ldg(&a.nodes[0].fc);   // the root is always expanded

    while (p < pool_hi) {
        const int idx = a.perm[p];
        const double4 me = a.posm[idx];
        const double xi = me.x, yi = me.y, zi = me.z;
        const long long ki = FIXED ? a.K[p] : a.K[idx];
        double ax = 0.0, ay = 0.0, az = 0.0;

        int node = root_fc;
        NodeRec rec = load_rec(a.nodes, node);
        int leaf_pos = 0, leaf_rem = 0, leaf_next = -1;
        while (node >= 0) {
            bool have = false;
            double rx = 0.0, ry = 0.0, rz = 0.0, d2 = 0.0, msrc = 0.0;
            int new_node = node;
            if (leaf_rem == 0) {
                const double cx = __hiloint2double(rec.A.y, rec.A.x), cy = __hiloint2double(rec.A.w, rec.A.z);
                const double cz = __hiloint2double(rec.B.y, rec.B.x);
                const long long kq = ((long long)rec.B.w << 32) | (unsigned)rec.B.z;
                const int fc = rec.C.x, next = rec.C.y, level = rec.C.z, count = rec.C.w;
                const int shift = 3 * (MAX_LEVEL - level);
                const bool leaf = fc < 0;
                const bool inside = ((ki ^ kq) >> shift) == 0;
                bool accepted = false;
                if (FIXED ? (!inside && !leaf) : !inside) {
                    rx = __dsub_rn(xi, cx); ry = __dsub_rn(yi, cy); rz = __dsub_rn(zi, cz);
                    d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                    accepted = s_cell2[level] < __dmul_rn(a.theta2, d2);
                }
                if (accepted) {
                    msrc = __ldg(&a.nodes[node].mass);
                    have = true;
                    new_node = next;
                } else if (leaf) {
                    leaf_pos = __ldg(&a.nodes[node].first);
                    leaf_rem = count;
                    leaf_next = next;
                } else {
                    new_node = fc;
                }
            }
            if (leaf_rem > 0) {   // one particle of the leaf per trip (sorted order), skipping the target itself
                const int jdx = __ldg(a.perm + leaf_pos);
                leaf_pos++;
                leaf_rem--;
                if (jdx != idx) {
                    const double4 pj = a.posm[jdx];
                    rx = __dsub_rn(xi, pj.x); ry = __dsub_rn(yi, pj.y); rz = __dsub_rn(zi, pj.z);
                    d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                    msrc = pj.w;
                    have = true;
                }
                if (leaf_rem == 0) new_node = leaf_next;
            }
            if (new_node != node && new_node >= 0) rec = load_rec(a.nodes, new_node);
            node = new_node;
            if (have) {
                const double s = __dadd_rn(d2, a.eps2);
                const double gm = __dmul_rn(a.G, msrc);
                double f;
                if (gm == 0.0 && s > 1e-200 && s < 1e200) {
                    f = gm;     // 0 / r^3 with r^3 finite and positive: exactly the signed zero gm (dropped zero-mass leaves)
                } else {
                    const double r = __dsqrt_rn(s);
                    f = __ddiv_rn(gm, __dmul_rn(__dmul_rn(r, r), r));
                }
                ax = __dsub_rn(ax, __dmul_rn(f, rx));
                ay = __dsub_rn(ay, __dmul_rn(f, ry));
                az = __dsub_rn(az, __dmul_rn(f, rz));
            }
        }
        a.acc[3 * (size_t)idx + 0] = ax;
        a.acc[3 * (size_t)idx + 1] = ay;
        a.acc[3 * (size_t)idx + 2] = az;
        p = atomicAdd(&s_next, 1);
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
        const int blocks = (npos + WALK_POOL - 1) / WALK_POOL;
        if (c->bh_mode == GRAV_B200_BH_FIXED) walk_kernel<true><<<blocks, WALK_BLOCK, 0, c->stream>>>(a);
        else walk_kernel<false><<<blocks, WALK_BLOCK, 0, c->stream>>>(a);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    if (c->world > 1) GB_TRY(comm_allreduce_sum(c, a.acc, 3 * c->n));   // disjoint targets + zeros: exact
    return GRAV_B200_OK;
}

}  // namespace gb
