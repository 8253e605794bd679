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
// particle.  History (ncu evidence under profiles/):
//   r1_walk_warp_union        one traversal per warp with lane masks: issue-bound, 7/32 lanes in the accept
//                             path, 4.4x the node visits one particle needs                     67.5 ms
//   r1_walk_stackless_exact   independent lanes + ropes: half the executed instructions          41.7 ms
//   r1_walk_state_machine     one item per trip; still ran a 3-load leaf path for 2/32 lanes in 77 % of trips
//   (this version)            accepted nodes and leaf particles share ONE source path: a source is a packed
//                             (x,y,z,m) record -- the node's (com, mass) or the particle's record from a
//                             Morton-sorted copy -- so the self test is a position compare and no lane waits
//                             on a perm -> posm chain                                   (N=2^20 Plummer, 1 GPU)
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
constexpr int WALK_CHUNK = 2048;    // multi-GPU: granularity of the interleaved target partition (16 CTAs)
// Tuning knobs kept for A/B builds (measured at N=2^20 Plummer, reference mode):
//   WALK_VARIANT 0  next node's record prefetched into registers before the arithmetic        39.3 ms  <- default
//   WALK_VARIANT 1  prefetch.global.L1 + reload at loop top (12 fewer live registers)          41.9 ms
//   WALK_MINB 10/12 (48/40 registers, more resident warps)                                41.6 / 50.7 ms
// More occupancy does not help: the kernel is bound by instructions per item and lane divergence, not latency.
#ifndef WALK_VARIANT
#define WALK_VARIANT 0
#endif
#ifndef WALK_MINB
#define WALK_MINB 8
#endif

struct WalkArgs {
    const WalkGeo *geo;        // plane 0: com + key
    const WalkTopo *topo;      // plane 1: topology + mass
    const long long *K;        // sorted keys
    const int *perm;           // sorted position -> particle id
    const double4 *psorted;    // particle records in sorted order
    const int *tord;           // optional: thread q handles sorted position tord[q] (walk-key grouping), else q
    int p_lo, p_hi;            // sorted positions (or slots of tord) handled by this launch
    int chunk, rank, world;    // world > 1: the launch covers chunks rank, rank + world, ... of `chunk` consecutive slots
    double G, eps2, theta2;
    double cell2[MAX_LEVEL + 2];   // (box_length / (2 << level))^2 per child level
    double *acc;               // AoS [3n] by particle id
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// 256-bit read-only global load (sm_100: LDG.E.ENL2.256).  The walk is bound by L1 data-pipe wavefronts
// (l1tex__data_pipe_lsu_wavefronts 78 % of peak, profiles/r1_walk_state_machine_n256k.txt): every load instruction
// of a warp costs one wavefront per distinct 128-byte line its lanes touch, so a node visit is served by TWO load
// instructions (32 + 16 bytes) instead of three, and a particle record by ONE instead of two.
__device__ __forceinline__ void ldg256(const void *p, long long &a, long long &b, long long &c, long long &d)
{
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

struct NodeRec {   // a whole WalkNode as two 256-bit loads (the mass rides along: no third load on accept)
    long long cx, cy, cz, kq;   // raw bits
    long long fc_next, first_lc, mass, pad;
};
__device__ __forceinline__ NodeRec load_rec(const WalkGeo *geo, const WalkTopo *topo, int node)
{
    NodeRec r;
    ldg256(geo + node, r.cx, r.cy, r.cz, r.kq);
    ldg256(topo + node, r.fc_next, r.first_lc, r.mass, r.pad);
    return r;
}

// Per-lane state machine.  Each trip of the loop handles ONE item for the lane -- a node visit, or the next
// particle of a leaf that is being summed directly -- in steps that are the same code for every lane:
//   decide    opening test on the current node record -> the next node id, and whether there is a source
//   fetch     the next node's record is requested BEFORE the arithmetic below (its latency overlaps the sqrt/div);
//             the source's mass (accepted node) or record (leaf particle) is loaded
//   evaluate  R = x_i - x_src, f = G m / r^3 and the three accumulator updates: sqrt(((rx^2+ry^2)+rz^2)+eps^2) for
//             both kinds of source, as in the reference (:164-171 and :198-213)
template <bool FIXED>
__global__ void __launch_bounds__(WALK_BLOCK, WALK_MINB) walk_kernel(const WalkArgs a)
{
    __shared__ double s_cell2[MAX_LEVEL + 2];
    if (threadIdx.x < MAX_LEVEL + 2) s_cell2[threadIdx.x] = a.cell2[threadIdx.x];
    __syncthreads();
    int q = a.p_lo + blockIdx.x * WALK_BLOCK + threadIdx.x;
    if (a.world > 1) {
        // interleaved chunks: walk cost varies smoothly along the Morton curve (dense centre vs halo), so every rank
        // gets a sample of all regions instead of one contiguous stretch
        const int t = blockIdx.x * WALK_BLOCK + threadIdx.x;
        q = ((t / a.chunk) * a.world + a.rank) * a.chunk + t % a.chunk;
    }
    if (q >= a.p_hi) return;
    const int p = a.tord ? a.tord[q] : q;
    const int idx = a.perm[p];
    const double4 me = a.psorted[p];
    const double xi = me.x, yi = me.y, zi = me.z;
    const long long ki = FIXED ? a.K[p] : a.K[idx];
    double ax = 0.0, ay = 0.0, az = 0.0;

    int node = __ldg(&a.topo[0].fc);   // the root is always expanded
#if WALK_VARIANT == 0
    NodeRec rec = load_rec(a.geo, a.topo, node);
#endif
    int leaf_pos = 0, leaf_rem = 0, leaf_next = -1;
    while (node >= 0) {
#if WALK_VARIANT == 1
        const NodeRec rec = load_rec(a.geo, a.topo, node);
#endif
        bool have = false;              // this trip produced a source
        bool from_node = false;         // ... which is the current node (else: particle at sorted position src_pos)
        int src_pos = 0;
        double sx = 0.0, sy = 0.0, sz = 0.0, msrc_node = 0.0;
        int new_node = node;
        if (leaf_rem == 0) {
            const double cx = __longlong_as_double(rec.cx), cy = __longlong_as_double(rec.cy), cz = __longlong_as_double(rec.cz);
            const long long kq = rec.kq;
            const int fc = (int)rec.fc_next, next = (int)(rec.fc_next >> 32), first = (int)rec.first_lc;
            const int lc = (int)(rec.first_lc >> 32);
            const int level = lc >> WALK_COUNT_BITS, count = lc & ((1 << WALK_COUNT_BITS) - 1);
            const int shift = 3 * (MAX_LEVEL - level);
            const bool leaf = fc < 0;
            const bool inside = ((ki ^ kq) >> shift) == 0;
            bool accepted = false;
            if (FIXED ? (!inside && !leaf) : !inside) {
                const double rx = __dsub_rn(xi, cx), ry = __dsub_rn(yi, cy), rz = __dsub_rn(zi, cz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                accepted = s_cell2[level] < __dmul_rn(a.theta2, d2);
            }
            if (accepted) {
                have = true; from_node = true;
                sx = cx; sy = cy; sz = cz; msrc_node = __longlong_as_double(rec.mass);
                new_node = next;
            } else if (leaf) {
                leaf_pos = first;
                leaf_rem = count;
                leaf_next = next;
            } else {
                new_node = fc;
            }
        }
        if (leaf_rem > 0) {   // one particle of the leaf per trip (sorted order), skipping the target itself
            src_pos = leaf_pos;
            have = (leaf_pos != p);
            leaf_pos++;
            leaf_rem--;
            if (leaf_rem == 0) new_node = leaf_next;
        }
        // fetch: source data of this trip, then the next node's record
        double msrc = 0.0;
        if (have) {
            if (from_node) {
                msrc = msrc_node;
            } else {
                long long qx, qy, qz, qw;
                ldg256(a.psorted + src_pos, qx, qy, qz, qw);
                sx = __longlong_as_double(qx); sy = __longlong_as_double(qy); sz = __longlong_as_double(qz);
                msrc = __longlong_as_double(qw);
            }
        }
#if WALK_VARIANT == 0
        if (new_node != node && new_node >= 0) rec = load_rec(a.geo, a.topo, new_node);
#else
        if (new_node != node && new_node >= 0) prefetch_l1(a.geo + new_node);
#endif
        node = new_node;
        if (have) {
            const double rx = __dsub_rn(xi, sx), ry = __dsub_rn(yi, sy), rz = __dsub_rn(zi, sz);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
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
}

int bh_pack_walk_nodes(grav_b200_ctx *c);
int radix_pass(grav_b200_ctx *c, const long long *kin, const int *vin, long long *kout, int *vout, int n, int shift);

// walk key of every sorted position (reference mode: the sorted key array indexed by the ORIGINAL particle id)
__global__ void __launch_bounds__(256) walk_keys_kernel(const long long *__restrict__ K, const int *__restrict__ perm, int n,
                                                       long long *__restrict__ ki, int *__restrict__ pos)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) { ki[p] = K[perm[p]]; pos[p] = p; }
}

int bh_walk(grav_b200_ctx *c, double eps, double theta)
{
    DevTree &t = c->tree;
    GB_TRY(bh_pack_walk_nodes(c));
    WalkArgs a{};
    a.geo = t.node_walk.as<WalkGeo>();
    a.topo = reinterpret_cast<const WalkTopo *>(a.geo + t.num_nodes);
    a.K = t.keys.as<long long>();
    a.perm = t.perm.as<int>();
    a.psorted = t.posm_sorted.as<double4>();
    // ranks share the walk by sorted position, not by particle id: interleaved chunks of WALK_CHUNK positions
    // (contiguous equal-count slices left the rank holding a Plummer sphere's centre 12 % behind at N = 2^24)
    a.p_lo = 0;
    a.p_hi = c->n;
    a.chunk = WALK_CHUNK; a.rank = c->rank; a.world = c->world;
    a.G = c->G;
    a.eps2 = eps * eps;
    a.theta2 = theta * theta;
    const double box_length = t.box_width * 2.0;          // src/acceleration_barnes_hut.c:100
    for (int level = 0; level <= MAX_LEVEL; level++) {
        const double bl = box_length / (double)(2 << level);   // :157
        a.cell2[level] = bl * bl;                              // :162 (left-hand side)
    }
    a.acc = c->acc.as<double>();
    // Reference mode only: the walk length of a target depends on the top bits of its (unrelated) walk key, so
    // Morton neighbours diverge.  Grouping targets by the key's level-1 octant (stable, Morton order inside a group)
    // raises the item-count lane efficiency from 0.73 to 0.87 on a Plummer sphere (oracle statistics).
    static const int group_bits = getenv("GRAV_B200_WALK_GROUP_BITS") ? atoi(getenv("GRAV_B200_WALK_GROUP_BITS")) : 0;
    if (group_bits > 0 && c->bh_mode != GRAV_B200_BH_FIXED) {
        const int n = c->n;
        GB_TRY(t.ki.reserve(sizeof(long long) * 2 * (size_t)n));
        GB_TRY(t.tord.reserve(sizeof(int) * 2 * (size_t)n));
        long long *ki = t.ki.as<long long>();
        int *pos = t.tord.as<int>();
        walk_keys_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(a.K, a.perm, n, ki, pos + n);
        GB_LAUNCH_CHECK();
        count_launch();
        GB_TRY(radix_pass(c, ki, pos + n, ki + n, pos, n, 63 - group_bits));   // digit = top group_bits bits (<= 8)
        a.tord = pos;
    }
    if (c->world > 1) GB_CUDA(cudaMemsetAsync(a.acc, 0, sizeof(double) * 3 * (size_t)c->n, c->stream));
    int npos = a.p_hi - a.p_lo;
    if (c->world > 1) {   // slots of this rank: its share of the WALK_CHUNK-sized chunks (the last one may be partial)
        const int chunks = (c->n + WALK_CHUNK - 1) / WALK_CHUNK;
        const int mine = chunks > c->rank ? (chunks - c->rank + c->world - 1) / c->world : 0;
        npos = mine * WALK_CHUNK;
    }
    if (npos > 0) {
        const int blocks = (npos + WALK_BLOCK - 1) / WALK_BLOCK;
        if (c->bh_mode == GRAV_B200_BH_FIXED) walk_kernel<true><<<blocks, WALK_BLOCK, 0, c->stream>>>(a);
        else walk_kernel<false><<<blocks, WALK_BLOCK, 0, c->stream>>>(a);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    if (c->world > 1) GB_TRY(comm_allreduce_sum(c, a.acc, 3 * c->n));   // disjoint targets + zeros: exact
    return GRAV_B200_OK;
}

}  // namespace gb
