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
//               node find the octant boundaries by binary search; the next level's records are
//               placed with a prefix sum, so record order is deterministic (by start position).
//   numbering   the reference serves expanded nodes in (start position, level) order and gives each all of
//               its children at once, so   first_child(u) = 1 + P[s_u] + same_start(u)   where P is the
//               exclusive prefix sum over sorted positions of "children created by expanded nodes starting
//               here" and same_start(u) counts the children of u's expanded ancestors that start at s_u.
//               id(child k of u) = first_child(u) + k.
//   moments     levels deepest-first, eight lanes per expanded node, children in id order; a leaf child adds
//               its particles one at a time in sorted order, an expanded child adds its finished sums;
//               separate multiply and add (no FMA) and IEEE division, as the reference's x86-64 build
//               does.  Leaves keep mass = com = 0 (src/linear_octree.c:567-570).
#include "internal.cuh"

namespace gb {

constexpr int MAX_LEVEL = 21;

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

__global__ void root_init_kernel(ExpRec *rec, int n)
{
    ExpRec r;
    for (int k = 0; k < 9; k++) r.b[k] = 0;
    r.b[0] = 0; r.b[1] = n;
    r.level = 0; r.parent = -1; r.rank = 0; r.same_start = 0; r.nch = 0; r.first_child = 1; r.id = 0;
    rec[0] = r;
}

// Eight lanes per expanded node, one per octant: lane o finds the first position of the node's range whose level-l
// digit is >= o (all keys of the range share the higher bits), so the eight binary searches -- chains of dependent L2
// loads -- run side by side instead of one after the other (16 x 25 us -> 16 x 6 us of build time at N = 60000).  The
// non-empty octants, in order, are the children.
__global__ void __launch_bounds__(128) expand_kernel(ExpRec *__restrict__ rec, int begin, int count,
                                                    const long long *__restrict__ K, int max_leaf,
                                                    int *__restrict__ W, int *__restrict__ nexp)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = gt >> 3, o = gt & 7;
    if (t >= count) {
        if (t == count && o == 0) nexp[t] = 0;   // sentinel so the scan also yields the total
        return;                                   // count * 8 is a multiple of 8: whole octets leave together
    }
    ExpRec *r = rec + begin + t;
    const int s = r->b[0], e = r->b[1];
    const int l = r->level + 1, shift = 3 * (MAX_LEVEL - l);
    // lower bound of digit o in [s, e)
    int lo = s, hi = e;
    if (o > 0) {
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((int)((K[mid] >> shift) & 7) < o) lo = mid + 1; else hi = mid;
        }
    }
    const int lane = threadIdx.x & 31, g0 = lane & ~7;
    const unsigned gmask = 0xffu << g0;
    const int up = __shfl_down_sync(gmask, lo, 1, 8);
    const int end = (o == 7) ? e : up;
    const bool nonempty = end > lo;
    const bool grows = nonempty && (end - lo > max_leaf) && l < MAX_LEVEL;
    const unsigned ne_mask = (__ballot_sync(gmask, nonempty) >> g0) & 0xffu;
    const unsigned gr_mask = (__ballot_sync(gmask, grows) >> g0) & 0xffu;
    const int nch = __popc(ne_mask);
    if (nonempty) r->b[__popc(ne_mask & ((1u << o) - 1u))] = lo;
    // slots nch .. 8 hold the end of the range (lane o takes slot nch + o; nch >= 1, so 8 lanes cover them)
    if (nch + o <= 8) r->b[nch + o] = e;
    if (o == 0) {
        r->nch = nch;
        nexp[t] = __popc(gr_mask);
        atomicAdd(&W[s], nch);
    }
}

__global__ void __launch_bounds__(128) emit_kernel(ExpRec *__restrict__ rec, int begin, int count, int next_begin,
                                                  const int *__restrict__ off, int max_leaf)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const ExpRec r = rec[begin + t];
    const int l = r.level + 1;
    if (l >= MAX_LEVEL) return;
    int o = next_begin + off[t];
    for (int k = 0; k < r.nch; k++) {
        const int cs = r.b[k], ce = r.b[k + 1];
        if (ce - cs > max_leaf) {
            ExpRec c;
            for (int j = 0; j < 9; j++) c.b[j] = 0;
            c.b[0] = cs; c.b[1] = ce;
            c.level = l; c.parent = begin + t; c.rank = k;
            c.same_start = (cs == r.b[0]) ? r.same_start + r.nch : 0;
            c.nch = 0; c.first_child = 0; c.id = 0;
            rec[o++] = c;
        }
    }
}

// Numbering and node arrays in one pass over the expanded records.  first_child(u) = 1 + P[s_u] + same_start(u); the id of
// u itself is first_child(parent) + rank (the parent's first_child is recomputed from the parent's record, so no kernel
// boundary is needed).  A parent writes count / first position of all its children and the leaf defaults of the children
// that stay leaves; an expanded child writes its own child count and first-child id (same rule as emit_kernel decides
// who is expanded), so no entry has two writers.
__global__ void __launch_bounds__(256) fill_nodes_kernel(ExpRec *__restrict__ rec, int ne, int n, int max_leaf,
                                                        const int *__restrict__ P, int *__restrict__ np,
                                                        int *__restrict__ nchild, int *__restrict__ first, int *__restrict__ fc)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ne) return;
    const ExpRec r = rec[u];
    const int fcu = 1 + P[r.b[0]] + r.same_start;
    int id = 0;
    if (r.parent >= 0) {
        const int ps = rec[r.parent].b[0], pss = rec[r.parent].same_start;
        id = 1 + P[ps] + pss + r.rank;
    } else {
        np[0] = n; first[0] = 0;
    }
    rec[u].first_child = fcu;
    rec[u].id = id;
    nchild[id] = r.nch;
    fc[id] = fcu;
    const int l = r.level + 1;
    for (int k = 0; k < r.nch; k++) {
        const int cid = fcu + k, cnt = r.b[k + 1] - r.b[k];
        np[cid] = cnt;
        first[cid] = r.b[k];
        if (!(cnt > max_leaf && l < MAX_LEVEL)) {
            nchild[cid] = 0;
            fc[cid] = -1;       // the reference leaves this uninitialised for leaves
        }
    }
}

// Eight lanes per expanded node, one per child: the loads of the children's contributions (a leaf's particle through
// perm -> posm, an expanded child's finished sums) go out together; then every lane of the octet adds the items up in
// the reference's order (children in id order, the particles of a leaf one at a time), so the sums are bit-identical to
// the serial loop.  Leaves with several particles (max_leaf > 1, or duplicates at level 21) are read in the ordered
// phase by all eight lanes (same address: one transaction).
__global__ void __launch_bounds__(128) moments_kernel(const ExpRec *__restrict__ rec, int begin, int count,
                                                     const int *__restrict__ perm, const double4 *__restrict__ posm,
                                                     const int *__restrict__ nchild, double *__restrict__ mass,
                                                     double *__restrict__ mtd, double *__restrict__ cx,
                                                     double *__restrict__ cy, double *__restrict__ cz)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = gt >> 3, o = gt & 7;
    if (t >= count) return;                      // whole octets leave together
    const ExpRec *r = rec + begin + t;
    const int nch = r->nch, fcid = r->first_child;
    const int g0 = (threadIdx.x & 31) & ~7;
    const unsigned gmask = 0xffu << g0;
    double im = 0.0, ix = 0.0, iy = 0.0, iz = 0.0;
    int start = 0, cnt = 0, multi = 0;
    if (o < nch) {
        const int cid = fcid + o;
        start = r->b[o];
        cnt = r->b[o + 1] - start;
        if (nchild[cid] != 0) {
            im = mass[cid];
            ix = mtd[3 * (size_t)cid + 0]; iy = mtd[3 * (size_t)cid + 1]; iz = mtd[3 * (size_t)cid + 2];
        } else if (cnt == 1) {
            const double4 q = posm[perm[start]];
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
                const double4 q = posm[perm[p]];
                tot = __dadd_rn(tot, q.w);
                sx = __dadd_rn(sx, __dmul_rn(q.w, q.x));
                sy = __dadd_rn(sy, __dmul_rn(q.w, q.y));
                sz = __dadd_rn(sz, __dmul_rn(q.w, q.z));
            }
        }
    }
    if (o == 0) {
        const int id = r->id;
        mass[id] = tot;
        mtd[3 * (size_t)id + 0] = sx; mtd[3 * (size_t)id + 1] = sy; mtd[3 * (size_t)id + 2] = sz;
        cx[id] = __ddiv_rn(sx, tot); cy[id] = __ddiv_rn(sy, tot); cz[id] = __ddiv_rn(sz, tot);
    }
}

// Packed per-node record for the walk.  Reference mode reproduces the walk's key fetch
// keys[sorted_indices[first_particle]] -- the SORTED key array indexed by an ORIGINAL particle id
// (src/acceleration_barnes_hut.c:143-147); fixed mode uses the node's own first key.
__global__ void __launch_bounds__(256) walk_nodes_kernel(int M, const int *__restrict__ np, const int *__restrict__ nchild,
                                                        const int *__restrict__ first, const int *__restrict__ fc,
                                                        const double *__restrict__ mass, const double *__restrict__ cx,
                                                        const double *__restrict__ cy, const double *__restrict__ cz,
                                                        const long long *__restrict__ K, const int *__restrict__ perm,
                                                        int fixed_mode, WalkGeo *__restrict__ geo, WalkTopo *__restrict__ topo)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= M) return;
    WalkGeo *g = geo + id;
    WalkTopo *w = topo + id;     // next / level are written by rope_kernel
    const int f = first[id];
    g->cx = cx[id]; g->cy = cy[id]; g->cz = cz[id];
    g->kq = fixed_mode ? K[f] : K[perm[f]];
    w->fc = nchild[id] > 0 ? fc[id] : -1;
    w->first = f;
    w->mass = mass[id];
    w->pad = 0;
    // next / level_count of every other node are written by rope_kernel (launched after this kernel)
    if (id == 0) { w->next = -1; w->level_count = np[id]; }
}

__global__ void __launch_bounds__(256) gather_sorted_kernel(const double4 *__restrict__ posm, const int *__restrict__ perm, int n,
                                                           double4 *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = posm[perm[p]];
}

// Ropes in one launch: child k < nch-1 is followed by its next sibling (id + 1); the last child inherits its parent's
// rope, found by climbing while the node is itself a last child (1.3 steps on average); the root's rope ends the walk.
__global__ void __launch_bounds__(256) rope_kernel(const ExpRec *__restrict__ rec, int ne, WalkTopo *__restrict__ nodes)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= ne) return;
    const ExpRec r = rec[u];
    int my_next = -1;
    int v = u;
    while (true) {
        const int parent = rec[v].parent;
        if (parent < 0) break;                                   // reached the root: -1
        if (rec[v].rank < rec[parent].nch - 1) { my_next = rec[v].id + 1; break; }
        v = parent;
    }
    for (int k = 0; k < r.nch; k++) {
        const int cid = r.first_child + k;
        nodes[cid].next = (k < r.nch - 1) ? cid + 1 : my_next;
        nodes[cid].level_count = ((r.level + 1) << WALK_COUNT_BITS) | (r.b[k + 1] - r.b[k]);
    }
}

static int grow_preserve(grav_b200_ctx *c, DevBuf &buf, size_t used_bytes, size_t need_bytes)
{
    if (need_bytes <= buf.cap && buf.p) return GRAV_B200_OK;
    size_t want = need_bytes + need_bytes / 2;
    void *np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = need_bytes;
        e = cudaMalloc(&np, want);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
    }
    if (buf.p && used_bytes) {
        e = cudaMemcpyAsync(np, buf.p, used_bytes, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { cudaFree(np); return cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__); }
    }
    if (buf.p) cudaFree(buf.p);
    buf.p = np;
    buf.cap = want;
    return GRAV_B200_OK;
}

// Builds keys, permutation, node arrays and moments on the device.  Leaves c->tree ready for bh_walk().
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
    GB_TRY(t.wsum.reserve(sizeof(int) * ((size_t)n + 1)));
    GB_TRY(t.wscan.reserve(sizeof(int) * ((size_t)n + 1)));
    GB_CUDA(cudaMemsetAsync(t.wsum.p, 0, sizeof(int) * ((size_t)n + 1), c->stream));
    // a generous first guess for the record count; grown on demand (deep chains can exceed it)
    GB_TRY(grow_preserve(c, t.exp_rec, 0, sizeof(ExpRec) * ((size_t)n / 2 + 1024)));
    root_init_kernel<<<1, 1, 0, c->stream>>>(t.exp_rec.as<ExpRec>(), n);
    GB_LAUNCH_CHECK();
    count_launch();

    int level_off[MAX_LEVEL + 2];
    level_off[0] = 0;
    level_off[1] = 1;
    int levels = 1;   // number of levels holding expanded nodes
    for (int l = 0; l < MAX_LEVEL; l++) {
        const int begin = level_off[l], count = level_off[l + 1] - begin;
        GB_TRY(t.counters.reserve(sizeof(int) * ((size_t)count + 1)));
        int *nexp = t.counters.as<int>();
        expand_kernel<<<(8 * (count + 1) + 127) / 128, 128, 0, c->stream>>>(t.exp_rec.as<ExpRec>(), begin, count, K, max_leaf,
                                                                     t.wsum.as<int>(), nexp);
        GB_LAUNCH_CHECK();
        count_launch();
        if (l + 1 >= MAX_LEVEL) { levels = l + 1; break; }   // level-21 nodes are never expanded
        GB_TRY(exclusive_scan_int(c, nexp, nexp, count + 1, t.scan_tmp));
        int total = 0;
        GB_CUDA(cudaMemcpyAsync(&total, nexp + count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        levels = l + 1;
        if (total == 0) break;
        const size_t used = sizeof(ExpRec) * (size_t)level_off[l + 1];
        GB_TRY(grow_preserve(c, t.exp_rec, used, used + sizeof(ExpRec) * (size_t)total));
        emit_kernel<<<(count + 127) / 128, 128, 0, c->stream>>>(t.exp_rec.as<ExpRec>(), begin, count, level_off[l + 1], nexp,
                                                               max_leaf);
        GB_LAUNCH_CHECK();
        count_launch();
        level_off[l + 2] = level_off[l + 1] + total;
    }
    const int ne = level_off[levels];
    t.num_expanded = ne;
    t.max_level = levels;
    for (int l = 0; l <= levels; l++) t.level_off[l] = level_off[l];

    // numbering
    GB_TRY(exclusive_scan_int(c, t.wsum.as<int>(), t.wscan.as<int>(), n + 1, t.scan_tmp));
    int total_children = 0;
    GB_CUDA(cudaMemcpyAsync(&total_children, t.wscan.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaMemcpyAsync(&t.box_width, t.bbox.as<double>() + 8 + 3, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    const int M = 1 + total_children;
    t.num_nodes = M;
    const size_t mi = sizeof(int) * (size_t)M, md = sizeof(double) * (size_t)M;
    GB_TRY(t.node_np.reserve(mi));
    GB_TRY(t.node_nch.reserve(mi));
    GB_TRY(t.node_first.reserve(mi));
    GB_TRY(t.node_fc.reserve(mi));
    GB_TRY(t.node_mass.reserve(md));
    GB_TRY(t.node_cx.reserve(md));
    GB_TRY(t.node_cy.reserve(md));
    GB_TRY(t.node_cz.reserve(md));
    GB_TRY(t.node_mtd.reserve(3 * md));
    GB_CUDA(cudaMemsetAsync(t.node_mass.p, 0, md, c->stream));
    GB_CUDA(cudaMemsetAsync(t.node_cx.p, 0, md, c->stream));
    GB_CUDA(cudaMemsetAsync(t.node_cy.p, 0, md, c->stream));
    GB_CUDA(cudaMemsetAsync(t.node_cz.p, 0, md, c->stream));
    ExpRec *rec = t.exp_rec.as<ExpRec>();
    const int eb = (ne + 255) / 256;
    fill_nodes_kernel<<<eb, 256, 0, c->stream>>>(rec, ne, n, max_leaf, t.wscan.as<int>(), t.node_np.as<int>(), t.node_nch.as<int>(),
                                                t.node_first.as<int>(), t.node_fc.as<int>());
    GB_LAUNCH_CHECK();
    count_launch();

    // moments, deepest level first
    for (int l = levels - 1; l >= 0; l--) {
        const int begin = level_off[l], count = level_off[l + 1] - begin;
        if (count <= 0) continue;
        moments_kernel<<<(8 * count + 127) / 128, 128, 0, c->stream>>>(rec, begin, count, t.perm.as<int>(), c->posm.as<double4>(),
                                                                  t.node_nch.as<int>(), t.node_mass.as<double>(),
                                                                  t.node_mtd.as<double>(), t.node_cx.as<double>(),
                                                                  t.node_cy.as<double>(), t.node_cz.as<double>());
        GB_LAUNCH_CHECK();
        count_launch();
    }
    stage_end(c, ST_BUILD);
    return GRAV_B200_OK;
}

int bh_pack_walk_nodes(grav_b200_ctx *c)
{
    DevTree &t = c->tree;
    const int M = t.num_nodes;
    GB_TRY(t.node_walk.reserve((sizeof(WalkGeo) + sizeof(WalkTopo)) * (size_t)M));
    WalkGeo *geo = t.node_walk.as<WalkGeo>();
    WalkTopo *nodes = reinterpret_cast<WalkTopo *>(geo + M);   // second plane
    walk_nodes_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(M, t.node_np.as<int>(), t.node_nch.as<int>(), t.node_first.as<int>(),
                                                            t.node_fc.as<int>(), t.node_mass.as<double>(), t.node_cx.as<double>(),
                                                            t.node_cy.as<double>(), t.node_cz.as<double>(), t.keys.as<long long>(),
                                                            t.perm.as<int>(), c->bh_mode == GRAV_B200_BH_FIXED ? 1 : 0, geo, nodes);
    GB_LAUNCH_CHECK();
    count_launch();
    GB_TRY(t.posm_sorted.reserve(sizeof(double4) * (size_t)t.n));
    gather_sorted_kernel<<<(t.n + 255) / 256, 256, 0, c->stream>>>(c->posm.as<double4>(), t.perm.as<int>(), t.n, t.posm_sorted.as<double4>());
    GB_LAUNCH_CHECK();
    count_launch();
    rope_kernel<<<(t.num_expanded + 255) / 256, 256, 0, c->stream>>>(t.exp_rec.as<ExpRec>(), t.num_expanded, nodes);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
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
    if (rc == GRAV_B200_OK) rc = bh_build(c, max_leaf, box_center, box_width);
    if (rc == GRAV_B200_OK) {
        DevTree &t = c->tree;
        const size_t M = (size_t)t.num_nodes;
        *out_box_width = t.box_width;
        *out_num_nodes = t.num_nodes;
        rc = copy_out(c, keys, t.keys.p, (size_t)n);
        if (rc == GRAV_B200_OK) rc = copy_out(c, sorted_indices, t.perm.p, (size_t)n);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_num_particles, t.node_np.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_num_internal_children, t.node_nch.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_first_particle_sorted_idx, t.node_first.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_first_internal_children_idx, t.node_fc.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_mass, t.node_mass.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_x, t.node_cx.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_y, t.node_cy.p, M);
        if (rc == GRAV_B200_OK) rc = copy_out(c, tree_com_z, t.node_cz.p, M);
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
