// Barnes-Hut stages BH-4..BH-5 on the device: linear-octree construction with the reference's node
// numbering, and the mass / centre-of-mass moments.
//
// Reference: setup_node + binary_search_num_particles_per_octant + helper_construct_octree
// (src/linear_octree.c:342-393, 406-576, 590-823), a serial depth-first walk that hands out node ids as
// nodes are expanded.  Here the same tree is produced level by level (validated bit-exact against the
// reference through oracle/grav_oracle.c, which restates this algorithm on the CPU):
//
//   discovery   breadth-first.  The root is always expanded.  The children of an expanded node of level
//               l-1 are the maximal runs of equal (key >> 3(21-l)) in its sorted range; a child is
//               expanded iff it holds more than max_leaf particles and l < 21.  Eight lanes per expanded
//               node find the octant boundaries by binary search and write the records of the children that
//               are expanded in turn; a CTA takes the slots for its 32 nodes' children with one atomic.
//   numbering   the reference serves expanded nodes in (start position, level) order and gives each all of
//               its children at once, so   first_child(u) = 1 + P[s_u] + same_start(u)   where P is the
//               exclusive prefix sum over sorted positions of "children created by expanded nodes starting
//               here" and same_start(u) counts the children of u's expanded ancestors that start at s_u.
//               id(child k of u) = first_child(u) + k.  Ids depend on positions only, never on the order of
//               the records inside a level, so the atomic slot allocation above does not show in the result.
//   moments     levels deepest-first, eight lanes per expanded node, children in id order; a leaf child adds
//               its particles one at a time in sorted order, an expanded child adds its finished sums;
//               separate multiply and add (no FMA) and IEEE division, as the reference's x86-64 build
//               does.  Leaves keep mass = com = 0 (src/linear_octree.c:567-570).
//
// Round 2: NO host synchronisation anywhere in the build.  Level populations, the record and node totals, the box
// width and the per-level cell sizes stay on the device (TreeMeta); every kernel is launched with a grid derived from
// an upper bound and loops over the device-side count.  Buffers are sized from n (records: 2n, nodes: 3n, times
// `slack`); a tree that does not fit -- only pathologically deep chains do that -- raises a flag in TreeMeta that the
// next synchronising call reports (GRAV_B200_ETREE), and the host-pointer entries retry with twice the room.
// The node arrays of LinearOctree (src/linear_octree.h:20-57) are no longer produced on the hot path: the walk records
// (two 32-byte planes) are written directly, and construct_octree() derives the arrays from them on demand.
#include "internal.cuh"

namespace gb {

constexpr int MAX_LEVEL = 21;
constexpr int EXPAND_THREADS = 256;                  // 32 nodes (octets) per CTA round
constexpr int COUNT_MASK = (1 << WALK_COUNT_BITS) - 1;

struct ExpRec {       // one expanded (= has children) node, 64 bytes
    int b[9];         // child k covers sorted positions [b[k], b[k+1]); before expansion b[0] = s, b[1] = e
    int level;        // level of this node; its children are level + 1
    int parent;       // index of the parent record, -1 for the root
    int rank;         // position among the parent's children
    int same_start;   // children of expanded ancestors that start at the same position
    int nch;
    int first_child;  // node id of child 0
    int id;           // node id of this node
};
static_assert(sizeof(ExpRec) == 64, "ExpRec layout");

int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp);
int bh_keys(grav_b200_ctx *c, const double *box_center, double box_width, bool want_unsorted);
int radix_sort_pairs(grav_b200_ctx *c);

__global__ void tree_init_kernel(ExpRec *rec, TreeMeta *meta, int n)
{
    // every word of the bookkeeping starts defined (the whole struct is copied to the host at the end of the build)
    for (int w = threadIdx.x; w < (int)(sizeof(TreeMeta) / sizeof(int)); w += blockDim.x) reinterpret_cast<int *>(meta)[w] = 0;
    __syncthreads();
    if (threadIdx.x == 0) meta->lvl_cnt[0] = 1;
    if (threadIdx.x == 0) {
        meta->num_expanded = 0; meta->num_nodes = 0; meta->levels = 0; meta->overflow = 0; meta->coop_barrier = 0;
        ExpRec r;
        for (int k = 0; k < 9; k++) r.b[k] = 0;
        r.b[0] = 0; r.b[1] = n;
        r.level = 0; r.parent = -1; r.rank = 0; r.same_start = 0; r.nch = 0; r.first_child = 1; r.id = 0;
        rec[0] = r;
    }
}

// One level of the discovery.  Eight lanes per expanded node, one per octant: lane o finds the first position of the
// node's range whose level-l digit is >= o (all keys of the range share the higher bits), so the eight binary searches --
// chains of dependent L2 loads -- run side by side.  The non-empty octants, in order, are the children; those that hold
// more than max_leaf particles get a record in the next level.  Record slots: the CTA adds up what its 32 nodes need and
// takes them with ONE atomicAdd on the next level's counter (an atomic per node would serialise ~10^6 operations on one
// address at N = 2^24).
// (bid, nblocks): the calling grid; rec / meta / W are also written by other phases of the single-launch build below, so they
// are plain pointers here (no read-only loads).
__device__ __forceinline__ void expand_level(ExpRec *rec, TreeMeta *meta, int l, const long long *__restrict__ K, int max_leaf,
                                             int *W, int ne_cap, int bid, int nblocks)
{
    __shared__ int s_warp_tot[EXPAND_THREADS / 32];
    __shared__ int s_base;
    const int begin = meta->lvl_off[l];
    int count = meta->lvl_cnt[l];
    if (begin + count > ne_cap) count = max(0, ne_cap - begin);   // overflow (flagged below): those records were never written
    const int next_begin = begin + count;
    if (bid == 0 && threadIdx.x == 0) meta->lvl_off[l + 1] = next_begin;
    const int o = threadIdx.x & 7, oct = threadIdx.x >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g0 = lane & ~7;
    const unsigned gmask = 0xffu << g0;
    const int cl = l + 1, shift = 3 * (MAX_LEVEL - cl);            // level of the children
    for (int t0 = bid * (EXPAND_THREADS / 8); t0 < count; t0 += nblocks * (EXPAND_THREADS / 8)) {   // CTA-uniform
        const int t = t0 + oct;
        const bool valid = t < count;
        ExpRec *r = rec + begin + (valid ? t : 0);
        int s = 0, e = 0, lo = 0, end = 0, nch = 0, my_rank = 0, same_start = 0;
        unsigned gr_mask = 0;
        bool grows = false;
        if (valid) {                                               // whole octets take this branch together
            s = r->b[0]; e = r->b[1];
            same_start = r->same_start;
            lo = s;
            int hi = e;
            if (o > 0) {
                while (lo < hi) {
                    const int mid = lo + ((hi - lo) >> 1);
                    if ((int)((K[mid] >> shift) & 7) < o) lo = mid + 1; else hi = mid;
                }
            }
            const int up = __shfl_down_sync(gmask, lo, 1, 8);
            end = (o == 7) ? e : up;
            const bool nonempty = end > lo;
            grows = nonempty && (end - lo > max_leaf) && cl < MAX_LEVEL;
            const unsigned ne_mask = (__ballot_sync(gmask, nonempty) >> g0) & 0xffu;
            gr_mask = (__ballot_sync(gmask, grows) >> g0) & 0xffu;
            nch = __popc(ne_mask);
            my_rank = __popc(ne_mask & ((1u << o) - 1u));
            if (nonempty) r->b[my_rank] = lo;
            // slots nch .. 8 hold the end of the range (lane o takes slot nch + o; nch >= 1, so 8 lanes cover them)
            if (nch + o <= 8) r->b[nch + o] = e;
            if (o == 0) {
                r->nch = nch;
                atomicAdd(&W[s], nch);
            }
        }
        // record slots of the expanded children: prefix over the CTA's 32 octets
        const int g = __popc(gr_mask);                             // same on the 8 lanes of an octet, 0 for idle octets
        const int q0 = __shfl_sync(0xffffffffu, g, 0), q1 = __shfl_sync(0xffffffffu, g, 8),
                  q2 = __shfl_sync(0xffffffffu, g, 16), q3 = __shfl_sync(0xffffffffu, g, 24);
        const int in_warp = (g0 >= 8 ? q0 : 0) + (g0 >= 16 ? q1 : 0) + (g0 >= 24 ? q2 : 0);
        if (lane == 0) s_warp_tot[warp] = q0 + q1 + q2 + q3;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < EXPAND_THREADS / 32; w++) tot += s_warp_tot[w];
            s_base = tot > 0 ? atomicAdd(&meta->lvl_cnt[cl], tot) : 0;
        }
        __syncthreads();
        if (grows) {
            int slot = next_begin + s_base + in_warp + __popc(gr_mask & ((1u << o) - 1u));
            for (int w = 0; w < warp; w++) slot += s_warp_tot[w];
            if (slot < ne_cap) {
                ExpRec c;
#pragma unroll
                for (int j = 0; j < 9; j++) c.b[j] = 0;
                c.b[0] = lo; c.b[1] = end;
                c.level = cl; c.parent = begin + t; c.rank = my_rank;
                c.same_start = (lo == s) ? same_start + nch : 0;
                c.nch = 0; c.first_child = 0; c.id = 0;
                rec[slot] = c;
            } else {
                atomicOr(&meta->overflow, TREE_OVERFLOW_EXPANDED);
            }
        }
        __syncthreads();   // s_warp_tot / s_base are rewritten by the next round
    }
}

__global__ void __launch_bounds__(EXPAND_THREADS) expand_kernel(ExpRec *rec, TreeMeta *meta, int l, const long long *__restrict__ K,
                                                               int max_leaf, int *W, int ne_cap)
{
    expand_level(rec, meta, l, K, max_leaf, W, ne_cap, blockIdx.x, gridDim.x);
}

// After the discovery and the scan of W: totals, overflow flags, box width and the walk's cell sizes.
__device__ __forceinline__ void tree_meta_body(TreeMeta *meta, const int *P, int n, const double *__restrict__ box, int ne_cap, int m_cap)
{
    int ne = 0, levels = 0;
    for (int l = 0; l <= MAX_LEVEL; l++) {
        const int cnt = meta->lvl_cnt[l];
        if (cnt > 0) levels = l + 1;
        ne += cnt;
    }
    if (ne > ne_cap) { atomicOr(&meta->overflow, TREE_OVERFLOW_EXPANDED); ne = ne_cap; }
    const int M = 1 + P[n];
    if (M > m_cap) atomicOr(&meta->overflow, TREE_OVERFLOW_NODES);
    meta->num_expanded = ne;
    meta->num_nodes = M;
    meta->levels = levels;
    const double w = box[3];
    meta->box_width = w;
    const double box_length = __dmul_rn(w, 2.0);                          // src/acceleration_barnes_hut.c:100
    for (int level = 0; level <= MAX_LEVEL; level++) {
        const double bl = __ddiv_rn(box_length, (double)(2 << level));    // :157
        meta->cell2[level] = __dmul_rn(bl, bl);                            // :162 (left-hand side)
    }
}

__global__ void tree_meta_kernel(TreeMeta *meta, const int *P, int n, const double *__restrict__ box, int ne_cap, int m_cap)
{
    if (threadIdx.x == 0) tree_meta_body(meta, P, n, box, ne_cap, m_cap);
}

// Numbering, topology and ropes in one pass over the expanded records, eight lanes per record (one per child).
//   first_child(u) = 1 + P[s_u] + same_start(u);  id(u) = first_child(parent) + rank  -- both pure functions of the
//   records' positions, so nothing here waits for another record's result.
//   A parent writes first particle, level/count, rope and inclusion key of all its children and the leaf defaults of the
//   children that stay leaves; an expanded node writes its own first-child id and child count; the moments kernel
//   writes mass and centre of mass of expanded nodes later.  No field has two writers.
//   Ropes: child k < nch-1 is followed by its next sibling (id + 1); the last child inherits its parent's rope, found by
//   climbing while the node is itself a last child (1.3 steps on average); the root's rope ends the walk.
//   Inclusion key: reference mode reproduces the walk's fetch keys[sorted_indices[first_particle]] -- the SORTED key
//   array indexed by an ORIGINAL particle id (src/acceleration_barnes_hut.c:143-147); fixed mode uses the node's own key.
__device__ __forceinline__ void fill_nodes_body(ExpRec *rec, const TreeMeta *meta, int n, int max_leaf, const int *P,
                                                const long long *__restrict__ K, const int *__restrict__ perm, int fixed_mode,
                                                WalkGeo *geo, WalkTopo *topo, int bid, int nblocks)
{
    if (meta->overflow) return;                      // ids would run past the planes; the host reports the flag
    const int ne = meta->num_expanded;
    for (int gt = bid * blockDim.x + threadIdx.x; (gt >> 3) < ne; gt += nblocks * blockDim.x) {
        const int u = gt >> 3, k = gt & 7;
        const ExpRec r = rec[u];
        const int fcu = 1 + P[r.b[0]] + r.same_start;
        if (k == 0) {
            int id = 0;
            if (r.parent >= 0) {
                const int ps = rec[r.parent].b[0], pss = rec[r.parent].same_start;
                id = 1 + P[ps] + pss + r.rank;
            } else {
                topo[0].next = -1; topo[0].first = 0; topo[0].shift_count = (unsigned)n;   // root (never tested: shift and cell2 unused)
                topo[0].cell2 = 0.0;
                geo[0].kq = 0;
            }
            rec[u].first_child = fcu;
            rec[u].id = id;
            topo[id].fcn = ((unsigned)fcu << 4) | (unsigned)r.nch;
        }
        if (k < r.nch) {
            const int cid = fcu + k, first = r.b[k], cnt = r.b[k + 1] - first;
            const int cl = r.level + 1;
            int nxt = cid + 1;
            if (k == r.nch - 1) {                     // last child: the rope of u
                nxt = -1;
                int v = u;
                while (true) {
                    const int parent = rec[v].parent;
                    if (parent < 0) break;
                    const int vr = rec[v].rank;
                    if (vr < rec[parent].nch - 1) { nxt = 1 + P[rec[parent].b[0]] + rec[parent].same_start + vr + 1; break; }
                    v = parent;
                }
            }
            WalkTopo *w = topo + cid;
            w->next = nxt;
            w->first = first;
            w->shift_count = ((unsigned)(3 * (MAX_LEVEL - cl)) << WALK_COUNT_BITS) | (unsigned)cnt;
            w->cell2 = meta->cell2[cl];
            geo[cid].kq = fixed_mode ? K[first] : K[perm[first]];
            if (!(cnt > max_leaf && cl < MAX_LEVEL)) {   // stays a leaf
                w->fcn = 0;
                w->mass = 0.0;
                geo[cid].cx = 0.0; geo[cid].cy = 0.0; geo[cid].cz = 0.0;
            }
        }
    }
}

__global__ void __launch_bounds__(256) fill_nodes_kernel(ExpRec *rec, const TreeMeta *meta, int n, int max_leaf, const int *P,
                                                        const long long *__restrict__ K, const int *__restrict__ perm, int fixed_mode,
                                                        WalkGeo *geo, WalkTopo *topo)
{
    fill_nodes_body(rec, meta, n, max_leaf, P, K, perm, fixed_mode, geo, topo, blockIdx.x, gridDim.x);
}

// One level of the moments.  Eight lanes per expanded node, one per child: the loads of the children's contributions (a
// leaf's particle from the Morton-sorted copy, an expanded child's finished sums) go out together; then every lane of
// the octet adds the items up in the reference's order (children in id order, the particles of a leaf one at a time), so
// the sums are bit-identical to the serial loop.  Leaves with several particles (max_leaf > 1, or duplicates at level
// 21) are read in the ordered phase by all eight lanes (same address: one transaction).
__device__ __forceinline__ void moments_level(const ExpRec *rec, const TreeMeta *meta, int l, int max_leaf,
                                              const double4 *__restrict__ psorted, WalkGeo *geo, WalkTopo *topo, double *mtd,
                                              int bid, int nblocks)
{
    if (meta->overflow) return;
    const int begin = meta->lvl_off[l], count = meta->lvl_cnt[l];
    const int o = threadIdx.x & 7;
    const int g0 = (threadIdx.x & 31) & ~7;
    const unsigned gmask = 0xffu << g0;
    for (int gt = bid * blockDim.x + threadIdx.x; (gt >> 3) < count; gt += nblocks * blockDim.x) {   // whole octets
        const ExpRec *r = rec + begin + (gt >> 3);
        const int nch = r->nch, fcid = r->first_child;
        const bool child_level_expandable = (r->level + 1) < MAX_LEVEL;
        double im = 0.0, ix = 0.0, iy = 0.0, iz = 0.0;
        int start = 0, cnt = 0, multi = 0;
        if (o < nch) {
            const int cid = fcid + o;
            start = r->b[o];
            cnt = r->b[o + 1] - start;
            if (cnt > max_leaf && child_level_expandable) {
                im = topo[cid].mass;
                ix = mtd[3 * (size_t)cid + 0]; iy = mtd[3 * (size_t)cid + 1]; iz = mtd[3 * (size_t)cid + 2];
            } else if (cnt == 1) {
                const double4 q = psorted[start];
                im = q.w;
                ix = __dmul_rn(q.w, q.x); iy = __dmul_rn(q.w, q.y); iz = __dmul_rn(q.w, q.z);
            } else {
                multi = 1;
            }
        }
        double tot = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
        for (int k = 0; k < nch; k++) {
            const int mk = __shfl_sync(gmask, multi, k, 8);
            if (!mk) {
                tot = __dadd_rn(tot, __shfl_sync(gmask, im, k, 8));
                sx = __dadd_rn(sx, __shfl_sync(gmask, ix, k, 8));
                sy = __dadd_rn(sy, __shfl_sync(gmask, iy, k, 8));
                sz = __dadd_rn(sz, __shfl_sync(gmask, iz, k, 8));
            } else {
                const int ps = __shfl_sync(gmask, start, k, 8), pc = __shfl_sync(gmask, cnt, k, 8);
                for (int p = ps; p < ps + pc; p++) {
                    const double4 q = psorted[p];
                    tot = __dadd_rn(tot, q.w);
                    sx = __dadd_rn(sx, __dmul_rn(q.w, q.x));
                    sy = __dadd_rn(sy, __dmul_rn(q.w, q.y));
                    sz = __dadd_rn(sz, __dmul_rn(q.w, q.z));
                }
            }
        }
        if (o == 0) {
            const int id = r->id;
            topo[id].mass = tot;
            mtd[3 * (size_t)id + 0] = sx; mtd[3 * (size_t)id + 1] = sy; mtd[3 * (size_t)id + 2] = sz;
            geo[id].cx = __ddiv_rn(sx, tot); geo[id].cy = __ddiv_rn(sy, tot); geo[id].cz = __ddiv_rn(sz, tot);
        }
    }
}

__global__ void __launch_bounds__(128) moments_kernel(const ExpRec *rec, const TreeMeta *meta, int l, int max_leaf,
                                                     const double4 *__restrict__ psorted, WalkGeo *geo, WalkTopo *topo, double *mtd)
{
    moments_level(rec, meta, l, max_leaf, psorted, geo, topo, mtd, blockIdx.x, gridDim.x);
}

// ---- the whole build in ONE cooperative launch ------------------------------------------------------------------------
// Small and medium systems (config 4: N = 60000) spend the build waiting on ~50 launches, half of them for levels the
// tree does not have.  Here the levels are loops inside one kernel with a grid barrier in between -- the launch is
// cooperative, so all CTAs are resident and the barrier cannot starve -- and both loops stop where the tree stops:
// discovery (a barrier per level), the exclusive scan of the children-per-start-position array (three phases), the
// bookkeeping, the node planes, the moments (a barrier per level).  Same device functions as the one-launch-per-level path.
struct TreeCoopArgs {
    ExpRec *rec;
    TreeMeta *meta;
    const long long *K;
    const int *perm;
    const double4 *psorted;
    const double *box;
    int *W, *P, *partial;       // W: children created per start position, P: its exclusive scan, partial: per-CTA sums
    WalkGeo *geo;
    WalkTopo *topo;
    double *mtd;
    unsigned *barrier;          // monotone arrival counter (zeroed before the launch)
    int n, max_leaf, ne_cap, m_cap, fixed_mode;
};

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();                       // this CTA's writes are visible before it arrives
        atomicAdd(bar, 1u);
        unsigned spins = 0;
        while (*(volatile unsigned *)bar < target) {
            if (++spins > (1u << 28)) break;   // never in a cooperative launch; bounds a mistake instead of hanging the GPU
        }
        __threadfence();                       // also drops this SM's L1 lines: the next phase reads what other SMs wrote
    }
    __syncthreads();
}

constexpr int TREE_COOP_THREADS = EXPAND_THREADS;

__global__ void __launch_bounds__(TREE_COOP_THREADS) tree_coop_kernel(const TreeCoopArgs a)
{
    __shared__ int s_scan[TREE_COOP_THREADS / 32];
    __shared__ int s_carry;
    unsigned target = 0;
    const int bid = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
    volatile TreeMeta *vm = a.meta;

    // discovery
    int levels = 1;
    for (int l = 0; l < MAX_LEVEL; l++) {
        expand_level(a.rec, a.meta, l, a.K, a.max_leaf, a.W, a.ne_cap, bid, nb);
        grid_barrier(a.barrier, target);
        levels = l + 1;
        if (vm->lvl_cnt[l + 1] == 0) break;     // the same value on every CTA (read after the barrier)
    }

    // numbering: P = exclusive scan of W[0..n]; every CTA owns one contiguous chunk
    const int total = a.n + 1;
    const int chunk = ((total + nb - 1) / nb + TREE_COOP_THREADS - 1) / TREE_COOP_THREADS * TREE_COOP_THREADS;
    const int c0 = min(bid * chunk, total), c1 = min(c0 + chunk, total);
    auto block_sum = [&](int v) -> int {        // total over the CTA, valid on every thread
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        __syncthreads();
        if ((tid & 31) == 0) s_scan[tid >> 5] = v;
        __syncthreads();
        int t = 0;
        for (int w = 0; w < TREE_COOP_THREADS / 32; w++) t += s_scan[w];
        return t;
    };
    {
        int v = 0;
        for (int i = c0 + tid; i < c1; i += TREE_COOP_THREADS) v += a.W[i];
        const int t = block_sum(v);
        if (tid == 0) a.partial[bid] = t;
    }
    grid_barrier(a.barrier, target);
    {
        // offset of this chunk: sum of the partials of the CTAs before it (nb <= a few hundred: one strided pass)
        int v = 0;
        for (int b = tid; b < bid; b += TREE_COOP_THREADS) v += ((volatile int *)a.partial)[b];
        int carry = block_sum(v);
        for (int i0 = c0; i0 < c1; i0 += TREE_COOP_THREADS) {
            const int i = i0 + tid;
            const int w = (i < c1) ? a.W[i] : 0;
            int inc = w;                          // inclusive scan over the CTA
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if ((tid & 31) >= d) inc += t;
            }
            __syncthreads();
            if ((tid & 31) == 31) s_scan[tid >> 5] = inc;
            __syncthreads();
            int before = 0, all = 0;
            for (int q = 0; q < TREE_COOP_THREADS / 32; q++) {
                const int t = s_scan[q];
                if (q < (tid >> 5)) before += t;
                all += t;
            }
            if (i < c1) a.P[i] = carry + before + inc - w;
            carry += all;
        }
    }
    grid_barrier(a.barrier, target);
    if (bid == 0 && tid == 0) tree_meta_body(a.meta, a.P, a.n, a.box, a.ne_cap, a.m_cap);
    grid_barrier(a.barrier, target);
    fill_nodes_body(a.rec, a.meta, a.n, a.max_leaf, a.P, a.K, a.perm, a.fixed_mode, a.geo, a.topo, bid, nb);
    grid_barrier(a.barrier, target);
    for (int l = levels - 1; l >= 0; l--) {
        moments_level(a.rec, a.meta, l, a.max_leaf, a.psorted, a.geo, a.topo, a.mtd, bid, nb);
        if (l > 0) grid_barrier(a.barrier, target);
    }
    (void)s_carry;
}

__global__ void __launch_bounds__(256) gather_sorted_kernel(const double4 *__restrict__ posm, const int *__restrict__ perm, int n,
                                                           double4 *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = posm[perm[p]];
}

// construct_octree() only: the eight per-node arrays of LinearOctree from the walk planes.
__global__ void __launch_bounds__(256) export_nodes_kernel(int M, const WalkGeo *__restrict__ geo, const WalkTopo *__restrict__ topo,
                                                          int *__restrict__ np, int *__restrict__ nchild, int *__restrict__ first,
                                                          int *__restrict__ fc, double *__restrict__ mass, double *__restrict__ cx,
                                                          double *__restrict__ cy, double *__restrict__ cz)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= M) return;
    const WalkTopo w = topo[id];
    const WalkGeo g = geo[id];
    np[id] = (int)(w.shift_count & (unsigned)COUNT_MASK);
    nchild[id] = (int)(w.fcn & 15u);
    first[id] = w.first;
    fc[id] = (w.fcn & 15u) ? (int)(w.fcn >> 4) : -1;   // -1 for leaves (the reference leaves it uninitialised)
    mass[id] = w.mass;
    cx[id] = g.cx; cy[id] = g.cy; cz[id] = g.cz;
}

static int grid_for(long long threads_needed, int block, int max_blocks)
{
    long long b = (threads_needed + block - 1) / block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

// Builds keys, permutation, walk planes and moments on the device.  Leaves c->tree ready for bh_walk().
// Queues work only: no host synchronisation (see the header comment).
int bh_build(grav_b200_ctx *c, int max_leaf, const double *box_center, double box_width)
{
    DevTree &t = c->tree;
    const int n = c->n;
    stage_begin(c, ST_MORTON);
    GB_TRY(bh_keys(c, box_center, box_width, false));
    stage_end(c, ST_MORTON);
    stage_begin(c, ST_SORT);
    GB_TRY(radix_sort_pairs(c));
    stage_end(c, ST_SORT);

    stage_begin(c, ST_BUILD);
    const long long *K = t.keys.as<long long>();
    // capacities: typical trees have 0.48 n expanded nodes and 1.48 n nodes; leaves are disjoint and non-empty, so
    // nodes <= 1 + n + expanded holds for any tree
    const long long ne_cap = (2LL * n + 1024) * t.slack, m_cap = ne_cap + n + 1;
    if (m_cap >= (1LL << 28)) { set_error("tree capacity %lld exceeds the 2^28 nodes a walk-stack entry can address", m_cap); return GRAV_B200_EINVAL; }
    t.ne_cap = (int)ne_cap;
    t.m_cap = (int)m_cap;
    GB_TRY(t.wsum.reserve(sizeof(int) * ((size_t)n + 1)));
    GB_TRY(t.wscan.reserve(sizeof(int) * ((size_t)n + 1)));
    GB_TRY(t.exp_rec.reserve(sizeof(ExpRec) * (size_t)ne_cap));
    GB_TRY(t.meta.reserve(sizeof(TreeMeta)));
    GB_TRY(t.node_walk.reserve((sizeof(WalkGeo) + sizeof(WalkTopo)) * (size_t)m_cap));
    GB_TRY(t.node_mtd.reserve(sizeof(double) * 3 * (size_t)m_cap));
    GB_TRY(t.posm_sorted.reserve(sizeof(double4) * (size_t)n));
    if (!t.h_meta) GB_CUDA(cudaHostAlloc((void **)&t.h_meta, sizeof(TreeMeta), cudaHostAllocDefault));
    ExpRec *rec = t.exp_rec.as<ExpRec>();
    TreeMeta *meta = t.meta.as<TreeMeta>();
    WalkGeo *geo = t.geo();
    WalkTopo *topo = t.topo();
    const int max_blocks = c->sm_count * 8;

    GB_CUDA(cudaMemsetAsync(t.wsum.p, 0, sizeof(int) * ((size_t)n + 1), c->stream));
    tree_init_kernel<<<1, 32, 0, c->stream>>>(rec, meta, n);
    GB_LAUNCH_CHECK();
    gather_sorted_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->posm.as<double4>(), t.perm.as<int>(), n, t.posm_sorted.as<double4>());
    GB_LAUNCH_CHECK();
    count_launch(2);

    // Single cooperative launch (levels as loops with grid barriers) unless the system is large enough for launch overheads
    // not to matter, the device refuses the cooperative grid, or GRAV_B200_TREE_KERNEL=levels asks for one launch per level.
    const long long most = (long long)n / (max_leaf + 1) + 1;
    static const char *tree_env = getenv("GRAV_B200_TREE_KERNEL");
    bool coop = tree_env ? strcmp(tree_env, "coop") == 0 : n <= (1 << 22);
    if (coop) {
        if (c->tree_coop_ctas < 0) {
            int per_sm = 0;
            GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tree_coop_kernel, TREE_COOP_THREADS, 0));
            c->tree_coop_ctas = per_sm * c->sm_count;
        }
        static const int ctas_env = getenv("GRAV_B200_TREE_CTAS") ? atoi(getenv("GRAV_B200_TREE_CTAS")) : 0;
        // measured (profiles/r2_tree_coop_grid.txt): N = 60000 is fastest with 2 CTAs per SM (0.21 ms), N = 2^20 with 4 or more
        int grid = grid_for(most * 8, TREE_COOP_THREADS, ctas_env > 0 ? ctas_env : c->sm_count * (n <= (1 << 17) ? 2 : 4));
        if (grid > c->tree_coop_ctas) grid = c->tree_coop_ctas;
        if (grid < 1) coop = false;
        if (coop) {
            GB_TRY(t.scan_tmp.reserve(sizeof(int) * ((size_t)grid + 16)));
            TreeCoopArgs ta{};
            ta.rec = rec; ta.meta = meta; ta.K = K; ta.perm = t.perm.as<int>(); ta.psorted = t.posm_sorted.as<double4>();
            ta.box = t.bbox.as<double>() + 8;
            ta.W = t.wsum.as<int>(); ta.P = t.wscan.as<int>(); ta.partial = t.scan_tmp.as<int>();
            ta.geo = geo; ta.topo = topo; ta.mtd = t.node_mtd.as<double>();
            ta.barrier = &meta->coop_barrier;
            ta.n = n; ta.max_leaf = max_leaf; ta.ne_cap = t.ne_cap; ta.m_cap = t.m_cap;
            ta.fixed_mode = c->bh_mode == GRAV_B200_BH_FIXED ? 1 : 0;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(TREE_COOP_THREADS);
            cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative;
            at[0].val.cooperative = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const cudaError_t e = cudaLaunchKernelEx(&cfg, tree_coop_kernel, ta);
            if (e == cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); coop = false; }
            else { GB_CUDA(e); count_launch(); }
        }
    }
    if (!coop) {
    // discovery: level l holds at most min(8^l, n / (max_leaf + 1)) expanded nodes
    long long pow8 = 1;
    for (int l = 0; l < MAX_LEVEL; l++) {
        const long long ub = pow8 < most ? pow8 : most;
        expand_kernel<<<grid_for(ub * 8, EXPAND_THREADS, max_blocks), EXPAND_THREADS, 0, c->stream>>>(rec, meta, l, K, max_leaf,
                                                                                                     t.wsum.as<int>(), t.ne_cap);
        GB_LAUNCH_CHECK();
        count_launch();
        if (pow8 < most) pow8 *= 8;
    }
    // numbering
    GB_TRY(exclusive_scan_int(c, t.wsum.as<int>(), t.wscan.as<int>(), n + 1, t.scan_tmp));
    tree_meta_kernel<<<1, 32, 0, c->stream>>>(meta, t.wscan.as<int>(), n, t.bbox.as<double>() + 8, t.ne_cap, t.m_cap);
    GB_LAUNCH_CHECK();
    fill_nodes_kernel<<<grid_for(most * 8, 256, max_blocks * 2), 256, 0, c->stream>>>(rec, meta, n, max_leaf, t.wscan.as<int>(), K,
                                                                                     t.perm.as<int>(), c->bh_mode == GRAV_B200_BH_FIXED ? 1 : 0,
                                                                                     geo, topo);
    GB_LAUNCH_CHECK();
    count_launch(2);
    // moments, deepest level first
    pow8 = 1;
    long long ubs[MAX_LEVEL];
    for (int l = 0; l < MAX_LEVEL; l++) { ubs[l] = pow8 < most ? pow8 : most; if (pow8 < most) pow8 *= 8; }
    for (int l = MAX_LEVEL - 1; l >= 0; l--) {
        moments_kernel<<<grid_for(ubs[l] * 8, 128, max_blocks * 2), 128, 0, c->stream>>>(rec, meta, l, max_leaf, t.posm_sorted.as<double4>(),
                                                                                        geo, topo, t.node_mtd.as<double>());
        GB_LAUNCH_CHECK();
        count_launch();
    }
    }
    GB_CUDA(cudaMemcpyAsync(t.h_meta, meta, sizeof(TreeMeta), cudaMemcpyDeviceToHost, c->stream));
    t.built = true;
    stage_end(c, ST_BUILD);
    return GRAV_B200_OK;
}

// To be called after the stream has been synchronised: reports a build that did not fit its buffers.
int bh_check(grav_b200_ctx *c)
{
    DevTree &t = c->tree;
    if (!t.built || !t.h_meta) return GRAV_B200_OK;
    t.built = false;
    if (t.h_meta->overflow) {
        set_error("the octree needs more room than was reserved (%d expanded nodes / %d nodes for n = %d; flags %d): "
                  "deep chains of close particles.  The host-pointer entries retry by themselves; for a resident run raise "
                  "GRAV_B200_TREE_SLACK (currently %d)", t.ne_cap, t.m_cap, t.n, t.h_meta->overflow, t.slack);
        return GRAV_B200_ETREE;
    }
    return GRAV_B200_OK;
}

// Build, wait, and if the tree did not fit rebuild with more room (the positions are still resident).  For callers
// that synchronise anyway: construct_octree(), the first force evaluation of a resident integration.
int bh_build_checked(grav_b200_ctx *c, int max_leaf, const double *box_center, double box_width)
{
    for (;;) {
        GB_TRY(bh_build(c, max_leaf, box_center, box_width));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        const int rc = bh_check(c);
        if (rc != GRAV_B200_ETREE) return rc;
        if (c->tree.slack >= 16) return rc;   // 32 n records cover the worst case of 21 levels x n / 2 nodes
        c->tree.slack *= 2;
    }
}

}  // namespace gb

// ---- host-pointer entries for stage-level parity tests and the construct_octree() drop-in ---------------
using namespace gb;

namespace gb {
int default_ctx_locked_begin(grav_b200_ctx **out);   // context.cu: takes the one-shot mutex
void default_ctx_locked_end();
}

template <class T>
static int copy_out(grav_b200_ctx *c, T **dst, const void *d_src, size_t count)
{
    *dst = (T *)malloc(sizeof(T) * (count ? count : 1));
    if (!*dst) { set_error("malloc failed for %zu bytes", sizeof(T) * count); return GRAV_B200_ENOMEM; }
    GB_CUDA(cudaMemcpyAsync(*dst, d_src, sizeof(T) * count, cudaMemcpyDeviceToHost, c->stream));
    return GRAV_B200_OK;
}

extern "C" int grav_b200_construct_octree(int n, const double *x, const double *m, int max_leaf, const double *box_center,
                                          double box_width, double *out_box_width, int *out_num_nodes, int64_t **keys,
                                          int **sorted_indices, int **tree_num_particles, int **tree_num_internal_children,
                                          int **tree_first_particle_sorted_idx, int **tree_first_internal_children_idx,
                                          double **tree_mass, double **tree_com_x, double **tree_com_y, double **tree_com_z)
{
    if (!x || !m || !out_box_width || !out_num_nodes || !keys || !sorted_indices || !tree_num_particles ||
        !tree_num_internal_children || !tree_first_particle_sorted_idx || !tree_first_internal_children_idx || !tree_mass ||
        !tree_com_x || !tree_com_y || !tree_com_z) {
        set_error("NULL pointer argument");
        return GRAV_B200_EINVAL;
    }
    if (n < 1) { set_error("num_particles must be >= 1, got %d", n); return GRAV_B200_EINVAL; }
    if (max_leaf == -1) max_leaf = 1;
    if (max_leaf < 1) { set_error("Maximum number of particles per leaf must be positive. Got: %d", max_leaf); return GRAV_B200_EINVAL; }
    grav_b200_ctx *c;
    GB_TRY(default_ctx_locked_begin(&c));
    int rc = grav_b200_ctx_set_system(c, n, x, nullptr, m, 1.0);
    if (rc == GRAV_B200_OK) rc = bh_build_checked(c, max_leaf, box_center, box_width);
    if (rc == GRAV_B200_OK) {
        DevTree &t = c->tree;
        const size_t M = (size_t)t.h_meta->num_nodes;
        *out_box_width = t.h_meta->box_width;
        *out_num_nodes = (int)M;
        // the LinearOctree arrays are not part of the hot path any more: derive them from the walk planes
        const size_t Mp = (M + 1) & ~(size_t)1;                      // keeps the double arrays 8-byte aligned
        rc = t.xport.reserve((4 * sizeof(int) + 4 * sizeof(double)) * Mp);
        if (rc == GRAV_B200_OK) {
            int *xi = t.xport.as<int>();
            double *xd = reinterpret_cast<double *>(xi + 4 * Mp);
            export_nodes_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>((int)M, t.geo(), t.topo(), xi, xi + Mp, xi + 2 * Mp,
                                                                                  xi + 3 * Mp, xd, xd + Mp, xd + 2 * Mp, xd + 3 * Mp);
            cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess) rc = cuda_fail(le, "export_nodes_kernel", __FILE__, __LINE__);
            count_launch();
            if (rc == GRAV_B200_OK) rc = copy_out(c, keys, t.keys.p, (size_t)n);
            if (rc == GRAV_B200_OK) rc = copy_out(c, sorted_indices, t.perm.p, (size_t)n);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_num_particles, xi, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_num_internal_children, xi + Mp, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_first_particle_sorted_idx, xi + 2 * Mp, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_first_internal_children_idx, xi + 3 * Mp, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_mass, xd, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_x, xd + Mp, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_y, xd + 2 * Mp, M);
            if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_z, xd + 3 * Mp, M);
        }
        if (rc == GRAV_B200_OK) {
            cudaError_t e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
        }
    }
    default_ctx_locked_end();
    return rc;
}

extern "C" int grav_b200_morton_keys(int n, const double *x, int64_t *keys_unsorted, double *out_center, double *out_width)
{
    if (!x || !keys_unsorted) { set_error("NULL pointer argument"); return GRAV_B200_EINVAL; }
    if (n < 1) { set_error("num_particles must be >= 1, got %d", n); return GRAV_B200_EINVAL; }
    grav_b200_ctx *c;
    GB_TRY(default_ctx_locked_begin(&c));
    // masses are irrelevant for keys: reuse x's first n values as a stand-in to avoid an extra host buffer
    double *ones = (double *)calloc((size_t)n, sizeof(double));
    int rc = ones ? GRAV_B200_OK : GRAV_B200_ENOMEM;
    if (rc == GRAV_B200_OK) rc = grav_b200_ctx_set_system(c, n, x, nullptr, ones, 1.0);
    if (rc == GRAV_B200_OK) rc = bh_keys(c, nullptr, -1.0, true);
    if (rc == GRAV_B200_OK) {
        cudaError_t e = cudaMemcpyAsync(keys_unsorted, c->tree.keys_unsorted.p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        double box[4];
        if (e == cudaSuccess) e = cudaMemcpyAsync(box, c->tree.bbox.as<double>() + 8, sizeof(box), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "morton_keys copy-out", __FILE__, __LINE__);
        else {
            if (out_center) { out_center[0] = box[0]; out_center[1] = box[1]; out_center[2] = box[2]; }
            if (out_width) *out_width = box[3];
        }
    }
    free(ones);
    default_ctx_locked_end();
    return rc;
}
