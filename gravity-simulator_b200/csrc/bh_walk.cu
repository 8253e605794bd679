// Barnes-Hut stage BH-8 on the device: the tree walk.
//
// Reference: helper_compute_acceleration (src/acceleration_barnes_hut.c:78-248): one CPU thread per particle
// (in Morton order) runs a depth-first walk with a private 22-frame stack.
//
// Two kernels, same visit set (every accept / open / leaf decision is taken per target with IEEE operations without FMA
// contraction, in the reference's operation order -- a flipped decision would change a force at the 1e-3 level):
//
//   walk_coop_kernel   DEFAULT.  One WARP per target, shared-memory stack of nodes to open.  Per trip the warp pops up to
//                      four opened nodes; the 8 lanes of an octet take the (<= 8, contiguous) children of one of them:
//                      coalesced record loads, the opening test in parallel, accepted nodes and single-particle leaves
//                      evaluated at once into lane-private partial sums (m r^-3 from the direct sum's rsqrt seed + one
//                      correction, FMAs), opened children pushed with a ballot.  The partial sums are added across the
//                      warp at the end, so the SUMMATION ORDER differs from the reference: results agree to ~1e-15
//                      relative (tests gate at 1e-12, the north-star tolerance).  Idle lanes are the price of the
//                      cooperative scheme (6.7-7.1 of 8 children per opened node, measured with the oracle), but there is
//                      no per-lane state machine, no divergent traversal, and the loads of a trip are 2-3 lines per octet.
//   walk_kernel        GRAV_B200_BH_EXACT=1 (or grav_b200_set_bh_exact).  One THREAD per target, stackless via ropes,
//                      one item (node visit or leaf particle) per loop trip, sources in the reference's depth-first order
//                      with sqrt / div / separate multiplies: output bit-identical to the x86-64 reference build.
//                      Round-1 history of this kernel (ncu evidence under profiles/r1_walk_*): one traversal per warp
//                      with lane masks 67.5 ms -> independent lanes + ropes 41.7 -> one item per trip 39.2 -> 256-bit
//                      loads and two 32-byte record planes 30.8 ms (N = 2^20 Plummer, theta = 0.5, one GPU).
//                      Its <FIXED, false> instantiation (GRAV_B200_WALK_KERNEL=lane) keeps the per-lane traversal but
//                      uses the fast evaluation: the A/B point between the two designs.
//
// Modes (grav_b200_set_bh_mode):
//   reference  bug-for-bug: the inclusion test compares keys fetched from the SORTED key array with ORIGINAL
//              particle ids (:120,:145), and the opening test is applied to leaves too, whose mass is 0, so
//              a far leaf is silently dropped (:150-176 precedes :181).
//   fixed      inclusion from the particle's / node's own key; leaves are never approximated.
#include "internal.cuh"

namespace gb {

constexpr int MAX_LEVEL = 21;
constexpr int COUNT_MASK = (1 << WALK_COUNT_BITS) - 1;
constexpr int WALK_BLOCK = 128;
constexpr int WALK_CHUNK = 2048;    // multi-GPU: granularity of the interleaved target partition
// Tuning knobs of the per-lane kernel kept for A/B builds (measured at N=2^20 Plummer, reference mode, round 1):
//   WALK_MINB 10/12 (48/40 registers, more resident warps)                                41.6 / 50.7 ms
#ifndef WALK_MINB
#define WALK_MINB 8
#endif

struct WalkArgs {
    const WalkGeo *geo;        // plane 0: com + key
    const WalkTopo *topo;      // plane 1: topology + mass
    const TreeMeta *meta;      // device-side build bookkeeping: overflow flag, cell sizes
    const long long *K;        // sorted keys
    const int *perm;           // sorted position -> particle id
    const double4 *psorted;    // particle records in sorted order
    const int *tord;           // optional (per-lane kernel): slot q handles sorted position tord[q] (walk-key grouping)
    int n;                     // sorted positions are [0, n)
    int chunk, rank, world;    // world > 1: this rank's slot t is position ((t / chunk) * world + rank) * chunk + t % chunk
    double G, eps2, theta2;
    double *acc;               // AoS [3n] by particle id (world == 1)
    double *out_slots;         // world > 1: results in slot order, [3 * slots]
};

// 256-bit read-only global load (sm_100: LDG.E.ENL2.256).  Every load instruction of a warp costs one L1 wavefront per
// distinct 128-byte line its lanes touch, so a node visit is served by TWO load instructions (geo + topo plane) and a
// particle record by ONE.
__device__ __forceinline__ void ldg256(const void *p, long long &a, long long &b, long long &c, long long &d)
{
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

struct NodeRec {   // a node's two records as two 256-bit loads (mass and cell size ride along: no third load)
    long long cx, cy, cz, kq;   // raw bits
    long long fcn_next, first_sc, mass, cell2;
};
__device__ __forceinline__ NodeRec load_rec(const WalkGeo *geo, const WalkTopo *topo, unsigned node)
{
    NodeRec r;
    ldg256(geo + node, r.cx, r.cy, r.cz, r.kq);
    ldg256(topo + node, r.fcn_next, r.first_sc, r.mass, r.cell2);
    return r;
}

__device__ __forceinline__ long long slot_to_position(const WalkArgs &a, long long t)
{
    // interleaved chunks: walk cost varies smoothly along the Morton curve (dense centre vs halo), so every rank
    // gets a sample of all regions instead of one contiguous stretch
    return a.world > 1 ? ((t / a.chunk) * a.world + a.rank) * a.chunk + t % a.chunk : t;
}

// fast evaluation of one source: the direct sum's 16 FP64 instructions; the sums carry G-less
// "source minus target" terms, G is applied once at the end
__device__ __forceinline__ void eval_fast(double sx, double sy, double sz, double sm, double xi, double yi, double zi,
                                          double eps2, double &ax, double &ay, double &az)
{
    const double dx = sx - xi, dy = sy - yi, dz = sz - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    const double s = inv_r3_times_m(r2, sm);
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
}

// ---------------------------------------------------------------------------------------------------------------------
// warp-cooperative walk
// ---------------------------------------------------------------------------------------------------------------------
#ifndef CW_WARPS_DEF
#define CW_WARPS_DEF 4
#endif
#ifndef CW_SYNC
#define CW_SYNC 0
#endif
#ifndef CW_OPAQUE
#define CW_OPAQUE 0     // 1: ~18 fewer instructions per trip (105 vs 123) and 0.7 ms SLOWER: the walk is latency bound, the rematerialised address arithmetic fills idle issue slots (profiles/r2_walk_opaque_ab.txt)
#endif
constexpr int CW_WARPS = CW_WARPS_DEF;       // warps per CTA
constexpr int CW_STACK = 512;     // stack entries per warp (4 bytes each: first child id << 4 | number of children)
// Above CW_HIGH entries the warp pops ONE opened node per trip, i.e. walks depth-first, where the stack grows by at most
// 7 entries per level, 140 in all; four at a time it grows by at most 28 per trip.  320 + 28 + 140 < 512.
constexpr int CW_HIGH = 320;

#ifndef CW_MINB
#define CW_MINB 8     // 64 registers, 32 warps per SM: the walk is latency bound (8 CTAs 18.9 ms, 6 CTAs / 80 registers 22.9 ms at N = 2^20 Plummer)
#endif

template <bool FIXED>
__global__ void __launch_bounds__(CW_WARPS * 32, CW_MINB) walk_coop_kernel(const WalkArgs a, int tpw, long long slots)
{
    __shared__ unsigned s_stack[CW_WARPS][CW_STACK];
    if (a.meta->overflow) return;   // the build did not fit its buffers: the planes are not a tree (reported by the host)
    const int lane = threadIdx.x & 31;
#if CW_OPAQUE
    // Per-lane constants handed to the compiler as opaque register values: under the 64-register cap ptxas otherwise
    // rematerialises them from SR_TID / SR_CgaCtaId on every trip (~20 of 123 instructions, profiles/r2_walk_coop_v2_n1m.txt).
    unsigned stk_s, e_o, k_o, lt_o;
    asm volatile("mov.u32 %0, %1;" : "=r"(stk_s) : "r"((unsigned)__cvta_generic_to_shared(s_stack[threadIdx.x >> 5])));
    asm volatile("mov.u32 %0, %1;" : "=r"(e_o) : "r"((unsigned)(lane >> 3)));
    asm volatile("mov.u32 %0, %1;" : "=r"(k_o) : "r"((unsigned)(lane & 7)));
    asm volatile("mov.u32 %0, %1;" : "=r"(lt_o) : "r"((1u << lane) - 1u));
    const int e = (int)e_o;
    const unsigned k = k_o, lt = lt_o;
#define STK_ST(idx, val) asm volatile("st.shared.u32 [%0], %1;" ::"r"(stk_s + 4u * (unsigned)(idx)), "r"(val) : "memory")
#define STK_LD(dst, idx) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dst) : "r"(stk_s + 4u * (unsigned)(idx)) : "memory")
#else
    unsigned *const stk = s_stack[threadIdx.x >> 5];
    const int e = lane >> 3;
    const unsigned k = lane & 7;
    const unsigned lt = (1u << lane) - 1u;
#define STK_ST(idx, val) stk[idx] = (val)
#define STK_LD(dst, idx) dst = stk[idx]
#endif
    const unsigned root_ent = __ldg(&a.topo[0].fcn);   // the root is always expanded
    const long long t0 = ((long long)blockIdx.x * CW_WARPS + (threadIdx.x >> 5)) * tpw;
    for (int tt = 0; tt < tpw; tt++) {
        const long long t = t0 + tt;                     // slot of this rank
        const long long q = slot_to_position(a, t);
        if (t >= slots || q >= a.n) {                    // positions grow with the slot: nothing further for this warp
#if CW_SYNC
            __syncthreads();
            continue;
#else
            break;
#endif
        }
        const int p = (int)q;
        const int idx = __ldg(&a.perm[p]);
        long long mx, my, mz, mw;
        ldg256(a.psorted + p, mx, my, mz, mw);
        const double xi = __longlong_as_double(mx), yi = __longlong_as_double(my), zi = __longlong_as_double(mz);
        const long long ki = FIXED ? __ldg(&a.K[p]) : __ldg(&a.K[idx]);
        double ax = 0.0, ay = 0.0, az = 0.0;

        int sp = 0;
        unsigned ent = (e == 0) ? root_ent : 0u;
        bool active = k < (ent & 15u);
        NodeRec rec = {};
        if (active) rec = load_rec(a.geo, a.topo, (ent >> 4) + k);
        for (;;) {
            bool open = false, have = false;
            int lf_first = 0, lf_count = 0;
            double dx = 0.0, dy = 0.0, dz = 0.0, d2 = 0.0, sm = 0.0;   // source minus target, squared distance, source mass
            unsigned child_ent = 0;
            if (active) {
                sm = __longlong_as_double(rec.mass);
                child_ent = (unsigned)rec.fcn_next;
                const unsigned sc = (unsigned)(rec.first_sc >> 32);
                const bool leaf = (child_ent & 15u) == 0u;
                const bool inside = ((unsigned long long)(ki ^ rec.kq) >> (sc >> WALK_COUNT_BITS)) == 0ull;
                // the decision arithmetic of src/acceleration_barnes_hut.c:150-162, operation for operation (the reference
                // forms x_i - com; com - x_i is its exact negative, so the squares and their sum are the same doubles)
                dx = __dsub_rn(__longlong_as_double(rec.cx), xi);
                dy = __dsub_rn(__longlong_as_double(rec.cy), yi);
                dz = __dsub_rn(__longlong_as_double(rec.cz), zi);
                d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const bool far = __longlong_as_double(rec.cell2) < __dmul_rn(a.theta2, d2);
                const bool accepted = FIXED ? (!inside && !leaf && far) : (!inside && far);
                open = !accepted && !leaf;
                have = accepted && sm != 0.0;      // d2 > 0 here, so a zero-mass node (a dropped leaf) contributes exactly 0
                if (!accepted && leaf) {
                    lf_first = (int)rec.first_sc;
                    lf_count = (int)(sc & (unsigned)COUNT_MASK);
                    if (lf_first != p) {          // first (usually only) particle of the leaf: same path as an accepted node
                        long long qx, qy, qz, qw;
                        ldg256(a.psorted + lf_first, qx, qy, qz, qw);
                        dx = __longlong_as_double(qx) - xi; dy = __longlong_as_double(qy) - yi; dz = __longlong_as_double(qz) - zi;
                        d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                        sm = __longlong_as_double(qw);
                        have = true;
                    }
                }
            }
            const unsigned om = __ballot_sync(0xffffffffu, open);
            if (open) STK_ST(sp + __popc(om & lt), child_ent);
            sp += __popc(om);
            __syncwarp();
            // next batch: its record loads are in flight while this trip's sources are evaluated
            const bool more = sp > 0;
            const int take = sp > CW_HIGH ? 1 : min(sp, 4);       // 0 when the stack is empty: every lane goes idle
            STK_LD(ent, max(sp - 1 - e, 0));
            if (e >= take) ent = 0u;
            __syncwarp();
            sp -= take;
            active = k < (ent & 15u);
            if (active) rec = load_rec(a.geo, a.topo, (ent >> 4) + k);
            if (have) {                                // the separation is already there: m r^-3 and three FMAs
                const double s = inv_r3_times_m(d2 + a.eps2, sm);
                ax = fma(s, dx, ax);
                ay = fma(s, dy, ay);
                az = fma(s, dz, az);
            }
            if (lf_count > 1) {                       // max_leaf > 1, or duplicates at level 21: the rest of the leaf
                for (int j = 1; j < lf_count; j++) {
                    const int pos = lf_first + j;
                    if (pos == p) continue;
                    long long qx, qy, qz, qw;
                    ldg256(a.psorted + pos, qx, qy, qz, qw);
                    eval_fast(__longlong_as_double(qx), __longlong_as_double(qy), __longlong_as_double(qz), __longlong_as_double(qw),
                              xi, yi, zi, a.eps2, ax, ay, az);
                }
            }
            if (!more) break;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, d);
            ay += __shfl_xor_sync(0xffffffffu, ay, d);
            az += __shfl_xor_sync(0xffffffffu, az, d);
        }
        if (lane < 3) {
            const double v = a.G * (lane == 0 ? ax : (lane == 1 ? ay : az));
            if (a.out_slots) a.out_slots[3 * (size_t)t + lane] = v;
            else a.acc[3 * (size_t)idx + lane] = v;
        }
#if CW_SYNC
        __syncthreads();     // keeps the CTA's warps on neighbouring targets in phase (they stream the same nodes through L1)
#else
        __syncwarp();
#endif
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-lane walk (bit-exact mode, and the fast-evaluation A/B variant)
// ---------------------------------------------------------------------------------------------------------------------
// Per-lane state machine.  Each trip of the loop handles ONE item for the lane -- a node visit, or the next
// particle of a leaf that is being summed directly -- in steps that are the same code for every lane:
//   decide    opening test on the current node record -> the next node id, and whether there is a source
//   fetch     the next node's record is requested BEFORE the arithmetic below (its latency overlaps the sqrt/div);
//             the source's mass (accepted node) or record (leaf particle) is loaded
//   evaluate  EXACT: R = x_i - x_src, f = G m / r^3 with sqrt(((rx^2+ry^2)+rz^2)+eps^2), r*r*r and a division, three
//             multiply-subtract pairs, as in the reference (:164-171 and :198-213); else eval_fast()
template <bool FIXED, bool EXACT>
__global__ void __launch_bounds__(WALK_BLOCK, WALK_MINB) walk_kernel(const WalkArgs a, long long slots)
{
    if (a.meta->overflow) return;
    const long long t = (long long)blockIdx.x * WALK_BLOCK + threadIdx.x;
    if (t >= slots) return;
    const long long q = slot_to_position(a, t);
    if (q >= a.n) return;
    const int p = a.tord ? a.tord[q] : (int)q;
    const int idx = a.perm[p];
    const double4 me = a.psorted[p];
    const double xi = me.x, yi = me.y, zi = me.z;
    const long long ki = FIXED ? a.K[p] : a.K[idx];
    double ax = 0.0, ay = 0.0, az = 0.0;

    int node = (int)(__ldg(&a.topo[0].fcn) >> 4);   // the root is always expanded
    NodeRec rec = load_rec(a.geo, a.topo, node);
    int leaf_pos = 0, leaf_rem = 0, leaf_next = -1;
    while (node >= 0) {
        bool have = false;              // this trip produced a source
        bool from_node = false;         // ... which is the current node (else: particle at sorted position src_pos)
        int src_pos = 0;
        double sx = 0.0, sy = 0.0, sz = 0.0, msrc_node = 0.0;
        int new_node = node;
        if (leaf_rem == 0) {
            const double cx = __longlong_as_double(rec.cx), cy = __longlong_as_double(rec.cy), cz = __longlong_as_double(rec.cz);
            const long long kq = rec.kq;
            const unsigned fcn = (unsigned)rec.fcn_next;
            const int fc = (int)(fcn >> 4), next = (int)(rec.fcn_next >> 32), first = (int)rec.first_sc;
            const unsigned sc = (unsigned)(rec.first_sc >> 32);
            const int shift = (int)(sc >> WALK_COUNT_BITS), count = (int)(sc & (unsigned)COUNT_MASK);
            const bool leaf = (fcn & 15u) == 0u;
            const bool inside = ((ki ^ kq) >> shift) == 0;
            bool accepted = false;
            if (FIXED ? (!inside && !leaf) : !inside) {
                const double rx = __dsub_rn(xi, cx), ry = __dsub_rn(yi, cy), rz = __dsub_rn(zi, cz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                accepted = __longlong_as_double(rec.cell2) < __dmul_rn(a.theta2, d2);
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
        if (new_node != node && new_node >= 0) rec = load_rec(a.geo, a.topo, new_node);
        node = new_node;
        if (have) {
            if (EXACT) {
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
            } else {
                eval_fast(sx, sy, sz, msrc, xi, yi, zi, a.eps2, ax, ay, az);
            }
        }
    }
    if (!EXACT) { ax *= a.G; ay *= a.G; az *= a.G; }
    double *dst = a.out_slots ? a.out_slots + 3 * (size_t)t : a.acc + 3 * (size_t)idx;
    dst[0] = ax;
    dst[1] = ay;
    dst[2] = az;
}

int radix_pass(grav_b200_ctx *c, const long long *kin, const int *vin, long long *kout, int *vout, int n, int shift);

// walk key of every sorted position (reference mode: the sorted key array indexed by the ORIGINAL particle id)
__global__ void __launch_bounds__(256) walk_keys_kernel(const long long *__restrict__ K, const int *__restrict__ perm, int n,
                                                       long long *__restrict__ ki, int *__restrict__ pos)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) { ki[p] = K[perm[p]]; pos[p] = p; }
}

// multi-GPU: the gathered per-rank slot-ordered results -> acc[particle id]
__global__ void __launch_bounds__(256) walk_scatter_kernel(const double *__restrict__ gathered, const int *__restrict__ perm, int n,
                                                          int chunk, int world, long long slots_per_rank, double *__restrict__ acc)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;     // sorted position
    if (q >= n) return;
    const int ch = q / chunk;
    const int r = ch % world;
    const long long t = (long long)(ch / world) * chunk + q % chunk;
    const double *src = gathered + 3 * ((size_t)r * slots_per_rank + (size_t)t);
    const int idx = perm[q];
    acc[3 * (size_t)idx + 0] = src[0];
    acc[3 * (size_t)idx + 1] = src[1];
    acc[3 * (size_t)idx + 2] = src[2];
}

int bh_walk(grav_b200_ctx *c, double eps, double theta)
{
    DevTree &t = c->tree;
    const int n = c->n;
    WalkArgs a{};
    a.geo = t.geo();
    a.topo = t.topo();
    a.meta = t.meta.as<TreeMeta>();
    a.K = t.keys.as<long long>();
    a.perm = t.perm.as<int>();
    a.psorted = t.posm_sorted.as<double4>();
    a.n = n;
    a.chunk = WALK_CHUNK; a.rank = c->rank; a.world = c->world;
    a.G = c->G;
    a.eps2 = eps * eps;
    a.theta2 = theta * theta;
    a.acc = c->acc.as<double>();
    const bool fixed = c->bh_mode == GRAV_B200_BH_FIXED;
    const bool exact = c->bh_exact != 0;
    static const char *kernel_env = getenv("GRAV_B200_WALK_KERNEL");
    const bool per_lane = exact || (kernel_env && strcmp(kernel_env, "lane") == 0);

    // Per-lane kernel, reference mode only: the walk length of a target depends on the top bits of its (unrelated) walk
    // key, so Morton neighbours diverge.  Grouping targets by the key's level-1 octant (stable, Morton order inside a
    // group) raises the item-count lane efficiency from 0.73 to 0.87 on a Plummer sphere (Plummer -7 %, uniform +15 %).
    static const int group_bits = getenv("GRAV_B200_WALK_GROUP_BITS") ? atoi(getenv("GRAV_B200_WALK_GROUP_BITS")) : 0;
    if (per_lane && group_bits > 0 && !fixed && c->world == 1) {
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

    // ranks share the walk by sorted position, not by particle id: interleaved chunks of WALK_CHUNK positions
    // (contiguous equal-count slices left the rank holding a Plummer sphere's centre 12 % behind at N = 2^24)
    long long slots = n, slots_per_rank = 0;
    double *gathered = nullptr;
    if (c->world > 1) {
        const long long chunks = ((long long)n + WALK_CHUNK - 1) / WALK_CHUNK;
        slots = chunks > c->rank ? ((chunks - c->rank + c->world - 1) / c->world) * WALK_CHUNK : 0;
        slots_per_rank = ((chunks + c->world - 1) / c->world) * WALK_CHUNK;
        GB_TRY(t.walk_out.reserve(sizeof(double) * 3 * (size_t)slots_per_rank * c->world));
        gathered = t.walk_out.as<double>();
        a.out_slots = gathered + 3 * (size_t)slots_per_rank * c->rank;
    }
    if (slots > 0) {
        if (per_lane) {
            const unsigned blocks = (unsigned)((slots + WALK_BLOCK - 1) / WALK_BLOCK);
            if (exact) {
                if (fixed) walk_kernel<true, true><<<blocks, WALK_BLOCK, 0, c->stream>>>(a, slots);
                else walk_kernel<false, true><<<blocks, WALK_BLOCK, 0, c->stream>>>(a, slots);
            } else {
                if (fixed) walk_kernel<true, false><<<blocks, WALK_BLOCK, 0, c->stream>>>(a, slots);
                else walk_kernel<false, false><<<blocks, WALK_BLOCK, 0, c->stream>>>(a, slots);
            }
        } else {
            // consecutive targets of a warp revisit almost the same nodes (L1 hits); fewer targets per warp when the
            // system is too small to give every SM four waves of warps
            int tpw = 8;
            while (tpw > 1 && (slots + tpw - 1) / tpw < (long long)c->sm_count * 32 * 4) tpw >>= 1;
            const long long warps = (slots + tpw - 1) / tpw;
            const unsigned blocks = (unsigned)((warps + CW_WARPS - 1) / CW_WARPS);
            if (fixed) walk_coop_kernel<true><<<blocks, CW_WARPS * 32, 0, c->stream>>>(a, tpw, slots);
            else walk_coop_kernel<false><<<blocks, CW_WARPS * 32, 0, c->stream>>>(a, tpw, slots);
        }
        GB_LAUNCH_CHECK();
        count_launch();
    }
    if (c->world > 1) {
        // every rank receives every rank's slot-ordered results (24 N bytes in all, half of what the round-1 all-reduce of
        // zero-padded full arrays moved, and no memset), then scatters them to particle order
        GB_TRY(comm_allgather_equal(c, gathered, 3 * (size_t)slots_per_rank));
        walk_scatter_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(gathered, a.perm, n, WALK_CHUNK, c->world, slots_per_rank, a.acc);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    return GRAV_B200_OK;
}

}  // namespace gb
