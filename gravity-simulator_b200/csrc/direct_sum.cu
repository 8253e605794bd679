// FP64 O(N^2) direct sum on sm_100a.
//
// Replaces the j-loops of acceleration_pairwise (reference src/acceleration.c:198-231) and
// acceleration_massless (:298-361).  Same mathematics -- a_i = -G sum_j m_j R_ij /
// (|R_ij|^2 + eps^2)^{3/2}, R_ij = x_i - x_j, no self term -- but organised for the GPU:
// every ordered interaction i<-j is evaluated (no Newton-3 scatter), i-accumulators live in
// registers, sources stream through shared memory as packed (x,y,z,m) records, and G is
// applied once at the end.
//
// Cost model (DESIGN.md "direct sum"): 16 FP64-pipe instructions per ordered interaction
//   3 DADD  (dx,dy,dz)            3 DFMA (r2 = eps2 + dx^2 + dy^2 + dz^2)
//   7       (r2^-3/2 * m_j from the MUFU.RSQ64H seed with one cubically convergent step)
//   3 DFMA  (accumulate)
// The seed y0 = rsqrt.approx.ftz.f64(r2) only looks at the high word of r2 (|e| <~ 2^-20 with
// e = 1 - r2*y0^2); r2^-3/2 = y0^3 (1-e)^-3/2 = y0^3 (1 + e(3/2 + 15/8 e) + 35/16 e^3 + ...),
// and the dropped term is < 2^-58, i.e. below double rounding.
//
// Work decomposition: "units" = (block of IB=256*TI targets) x (tile of 256 sources), flattened
// target-block-major and split EVENLY over a persistent grid of sm_count*occupancy CTAs
// (stream-K style), so there is no tail wave whatever N is.  A CTA whose range ends inside a
// target block writes its partial sums to a scratch slot; a tiny fix-up kernel adds the slots
// of each split block in CTA order, so results are deterministic (no floating-point atomics).
#include "internal.cuh"

namespace gb {

struct DSArgs {
    const double4 *src;       // sources, zero padded to a multiple of DS_TJ
    const int *src_id;        // MASSLESS: particle id of each source; PAIRWISE: nullptr (id == j)
    const double *src_altm;   // MASSLESS: mass used when the target is massless (m[rank] quirk)
    int n_src;
    const double4 *tgt;       // targets, indexed by particle id
    int i_lo, i_hi;           // target range handled by this launch
    double eps2, G;
    double *acc;              // AoS [3n], indexed by particle id
    double *partials;         // [grid][2][3][IB]
    int NB, NT;               // target blocks, source tiles
    int gpt;                  // work granules per source tile (8 = 32 sources each, 1 = whole tiles): the unit of the even split
    int skip_special;         // fast kernel: leave the diagonal / padded tiles to direct_sum_special_kernel
};

template <int TI>
struct Acc {
    double x[TI], y[TI], z[TI];
};

// One ordered interaction target <- source (16 FP64-pipe instructions in the unchecked form).
//   CHECK: mask the self term (source id == target id) and padding sources.
template <bool CHECK>
__device__ __forceinline__ void interaction(const double4 pj, double mj, bool masked, double xi, double yi, double zi,
                                            double eps2, double &ax, double &ay, double &az)
{
    const double dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double s = inv_r3_times_m(r2, mj);
    if (CHECK) {
        if (masked) s = 0.0;
    }
    ax = fma(s, dx, ax);
    ay = fma(s, dy, ay);
    az = fma(s, dz, az);
}

// One shared-memory tile against the TI register-resident targets of this thread.
//   MASSLESS: per-target choice between the true source mass and the quirk mass
template <int TI, bool CHECK, bool MASSLESS>
__device__ __forceinline__ void tile_interactions(const double4 *__restrict__ tile, const int *__restrict__ tile_id,
                                                  const double *__restrict__ tile_altm, int j_base, int n_src,
                                                  const double (&xi)[TI], const double (&yi)[TI],
                                                  const double (&zi)[TI], const int (&ii)[TI],
                                                  const bool (&alt)[TI], double eps2, Acc<TI> &a, int ja = 0, int jb = DS_TJ)
{
    // sources [ja, jb) of the tile (whole granules: multiples of 32); checked loops stop at the last real source (the
    // massless method has a handful of sources in a 256-wide tile)
    const int jcount = CHECK ? min(jb, n_src - j_base) : jb;
#pragma unroll 2
    for (int j = ja; j < jcount; j++) {
        const double4 pj = tile[j];
        int jid = j_base + j;
        double altm = 0.0;
        if (MASSLESS) {
            jid = tile_id[j];
            altm = tile_altm[j];
        }
        const bool jpad = CHECK && (j_base + j) >= n_src;
#pragma unroll
        for (int t = 0; t < TI; t++) {
            double mj = pj.w;
            if (MASSLESS) mj = alt[t] ? altm : mj;
            interaction<CHECK>(pj, mj, CHECK && (jpad || jid == ii[t]), xi[t], yi[t], zi[t], eps2, a.x[t], a.y[t], a.z[t]);
        }
    }
}

__host__ __device__ inline long long unit_begin(long long c, long long U, long long C) { return (c * U) / C; }

// A source tile needs the checked loop when it can contain the self term (it overlaps the target block)
// or padding (it is the last, partial tile).
__device__ __forceinline__ bool tile_is_special(int jt, int n_src, int blk_lo, int blk_hi)
{
    const int j_base = jt * DS_TJ;
    return (j_base + DS_TJ > n_src) || (j_base < blk_hi && j_base + DS_TJ > blk_lo);
}

// Main kernel.  CHECK=false: the fast loop only; special tiles are SKIPPED (direct_sum_special_kernel adds
// them afterwards), which keeps this kernel's code identical to the tuned loop (one inner-loop variant, no
// extra live registers).  CHECK=true: every tile goes through the checked loop (small problems, massless).
//   GRAN: the even split cuts at 32-source granules (small problems) instead of whole tiles.  The variable inner-loop bounds
//   cost the tuned loop 6 % (128 registers, no slack for ptxas), which pays only while a CTA gets fewer than ~16 tiles.
template <int TI, bool CHECK, bool MASSLESS, bool GRAN>
__global__ void __launch_bounds__(DS_BLOCK, 2)   // tuned on B200: TI=4, 2 CTAs/SM, unroll 2 (scratch/ds_tune.cu)
direct_sum_kernel(const DSArgs p)
{
    constexpr int IB = DS_BLOCK * TI;
    __shared__ double4 tile[DS_TJ];
    __shared__ int tile_id[MASSLESS ? DS_TJ : 1];
    __shared__ double tile_altm[MASSLESS ? DS_TJ : 1];

    const int tid = threadIdx.x;
    // Granule range of this CTA.  The work of a target block is cut into granules of DS_TJ / gpt sources (32 when the
    // problem is small enough to care: with whole tiles as the unit N = 16384 gave the CTAs 3 or 4 tiles each and the SMs
    // 6 to 8, a 16 % tail).  64-bit only for the product c*U; the host keeps U below 2^31.
    const int gpt = GRAN ? p.gpt : 1;
    const int NG = p.NT * gpt;                          // granules per target block
    const int gsz = DS_TJ / gpt;                        // sources per granule
    const long long U = (long long)p.NB * NG;
    const int u0 = (int)unit_begin(blockIdx.x, U, gridDim.x);
    const int u1 = (int)unit_begin(blockIdx.x + 1, U, gridDim.x);
    if (u0 >= u1) return;
    const int ib_first = u0 / NG;

    int u = u0;
    while (u < u1) {
        const int ib = u / NG;
        const int ga = u - ib * NG;
        const int seg_end = min(u1, (ib + 1) * NG);
        const int gb = seg_end - ib * NG;
        const bool full = (ga == 0 && gb == NG);

        // targets of this thread
        double xi[TI], yi[TI], zi[TI];
        int ii[TI];
        bool alt[TI];
        Acc<TI> a;
        const int blk_lo = p.i_lo + ib * IB;
#pragma unroll
        for (int t = 0; t < TI; t++) {
            const int i = blk_lo + t * DS_BLOCK + tid;
            ii[t] = (CHECK || MASSLESS) ? i : 0;     // only the checked loop needs the ids kept live
            const double4 q = (i < p.i_hi) ? p.tgt[i] : make_double4(0.0, 0.0, 0.0, 1.0);
            xi[t] = q.x; yi[t] = q.y; zi[t] = q.z;
            alt[t] = MASSLESS ? (q.w == 0.0) : false;
            a.x[t] = 0.0; a.y[t] = 0.0; a.z[t] = 0.0;
        }
        const int blk_hi = min(blk_lo + IB, p.i_hi);   // exclusive

        // stream the source tiles through shared memory (the second CTA of the SM computes meanwhile); the first and the
        // last tile of the segment may be entered / left part-way
        const int jt0 = ga / gpt, jt1 = (gb - 1) / gpt + 1;
        for (int jt = jt0; jt < jt1; jt++) {
            if (!CHECK && p.skip_special && tile_is_special(jt, p.n_src, blk_lo, blk_hi)) continue;   // block-uniform
            __syncthreads();
            tile[tid] = p.src[(size_t)jt * DS_TJ + tid];
            if (MASSLESS) {
                tile_id[tid] = p.src_id[(size_t)jt * DS_TJ + tid];
                tile_altm[tid] = p.src_altm[(size_t)jt * DS_TJ + tid];
            }
            __syncthreads();
            if (GRAN) {
                const int ja = max(ga - jt * gpt, 0) * gsz, jb = min(gb - jt * gpt, gpt) * gsz;
                tile_interactions<TI, CHECK, MASSLESS>(tile, tile_id, tile_altm, jt * DS_TJ, p.n_src, xi, yi, zi, ii, alt, p.eps2, a, ja, jb);
            } else {
                tile_interactions<TI, CHECK, MASSLESS>(tile, tile_id, tile_altm, jt * DS_TJ, p.n_src, xi, yi, zi, ii, alt, p.eps2, a);
            }
        }

        if (full) {
#pragma unroll
            for (int t = 0; t < TI; t++) {
                const int i = blk_lo + t * DS_BLOCK + tid;
                if (i < p.i_hi) {
                    // a_i = -G * sum (x_i - x_j) s = G * sum (x_j - x_i) s
                    p.acc[3 * (size_t)i + 0] = p.G * a.x[t];
                    p.acc[3 * (size_t)i + 1] = p.G * a.y[t];
                    p.acc[3 * (size_t)i + 2] = p.G * a.z[t];
                }
            }
        } else {
            const int slot = (ib == ib_first) ? 0 : 1;
            double *dst = p.partials + ((size_t)blockIdx.x * 2 + slot) * 3 * IB;
#pragma unroll
            for (int t = 0; t < TI; t++) {
                const int li = t * DS_BLOCK + tid;
                dst[0 * IB + li] = a.x[t];
                dst[1 * IB + li] = a.y[t];
                dst[2 * IB + li] = a.z[t];
            }
        }
        u = seg_end;
    }
}

// The tiles the fast kernel skipped: for each of its target blocks (IB_MAIN targets), the source tiles that overlap
// the block plus the partial last tile, through the checked loop; acc += G * sum.  One target per thread and one
// CTA per 256 targets (4x the CTAs of the main decomposition, the same 4-5 tiles each); runs after the main kernel
// and its fix-up on the same stream, so the result is deterministic.
template <int IB_MAIN>
__global__ void __launch_bounds__(DS_BLOCK) direct_sum_special_kernel(const DSArgs p)
{
    __shared__ double4 tile[DS_TJ];
    const int tid = threadIdx.x;
    const int my_lo = p.i_lo + blockIdx.x * DS_BLOCK;
    // the main-kernel block these targets belong to decides which tiles were skipped for them
    const int blk_lo = p.i_lo + ((my_lo - p.i_lo) / IB_MAIN) * IB_MAIN;
    const int blk_hi = min(blk_lo + IB_MAIN, p.i_hi);
    double xi[1], yi[1], zi[1];
    int ii[1];
    bool alt[1] = {false};
    Acc<1> a;
    ii[0] = my_lo + tid;
    const double4 q = (ii[0] < p.i_hi) ? p.tgt[ii[0]] : make_double4(0.0, 0.0, 0.0, 1.0);
    xi[0] = q.x; yi[0] = q.y; zi[0] = q.z;
    a.x[0] = 0.0; a.y[0] = 0.0; a.z[0] = 0.0;
    const int jt_a = blk_lo / DS_TJ, jt_b = min((blk_hi - 1) / DS_TJ, p.NT - 1);
    const int last = p.NT - 1;
    const bool last_partial = (p.n_src % DS_TJ) != 0 && last > jt_b;
    for (int k = jt_a; k <= jt_b + (last_partial ? 1 : 0); k++) {
        const int jt = (k <= jt_b) ? k : last;
        __syncthreads();
        tile[tid] = p.src[(size_t)jt * DS_TJ + tid];
        __syncthreads();
        tile_interactions<1, true, false>(tile, nullptr, nullptr, jt * DS_TJ, p.n_src, xi, yi, zi, ii, alt, p.eps2, a);
    }
    if (ii[0] < p.i_hi) {
        p.acc[3 * (size_t)ii[0] + 0] += p.G * a.x[0];
        p.acc[3 * (size_t)ii[0] + 1] += p.G * a.y[0];
        p.acc[3 * (size_t)ii[0] + 2] += p.G * a.z[0];
    }
}

// Adds the partial sums of every target block that was split over several CTAs, in CTA order.  One CTA per 256
// targets (grid = NB x TI); the CTA range of the block and the slot each of those CTAs used are worked out once per
// CTA (the 64-bit divisions of unit_begin() per target and contributor made the first version of this kernel cost
// 16 % of a force evaluation at N = 16384).
constexpr int DS_FIXUP_MAX_CONTRIB = 1024;
template <int TI>
__global__ void __launch_bounds__(DS_BLOCK) direct_sum_fixup_kernel(const DSArgs p, int C)
{
    constexpr int IB = DS_BLOCK * TI;
    __shared__ int s_range[2];
    __shared__ signed char s_slot[DS_FIXUP_MAX_CONTRIB];     // 0 / 1: partial slot of contributor c_lo + k, -1: none
    const int ib = blockIdx.x;
    const int NG = p.NT * p.gpt;
    const long long U = (long long)p.NB * NG;
    if (threadIdx.x == 0) {
        const long long ua = (long long)ib * NG, ub = ua + NG - 1;
        // owner(u) = largest c with unit_begin(c) <= u
        auto owner = [&](long long u) {
            int lo = 0, hi = C - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (unit_begin(mid, U, C) <= u) lo = mid; else hi = mid - 1;
            }
            return lo;
        };
        s_range[0] = owner(ua);
        s_range[1] = owner(ub);
    }
    __syncthreads();
    const int c_lo = s_range[0], c_hi = s_range[1];
    if (c_lo == c_hi) return;   // one CTA had the whole block and wrote acc itself
    for (int k = threadIdx.x; k <= c_hi - c_lo; k += DS_BLOCK) {
        const int c = c_lo + k;
        const long long b = unit_begin(c, U, C), e = unit_begin(c + 1, U, C);
        s_slot[k] = (b >= e) ? -1 : (((int)(b / NG) == ib) ? 0 : 1);
    }
    __syncthreads();
    const int li = blockIdx.y * DS_BLOCK + threadIdx.x;
    const int i = p.i_lo + ib * IB + li;
    if (i >= p.i_hi) return;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int k = 0; k <= c_hi - c_lo; k++) {
        const int slot = s_slot[k];
        if (slot < 0) continue;
        const double *src = p.partials + ((size_t)(c_lo + k) * 2 + slot) * 3 * IB;
        sx += src[0 * IB + li];
        sy += src[1 * IB + li];
        sz += src[2 * IB + li];
    }
    p.acc[3 * (size_t)i + 0] = p.G * sx;
    p.acc[3 * (size_t)i + 1] = p.G * sy;
    p.acc[3 * (size_t)i + 2] = p.G * sz;
}

constexpr int DS_TI = 4;

template <bool MASSLESS, int TI>
static int launch_direct_sum_t(grav_b200_ctx *c, DSArgs &a)
{
    constexpr int IB = DS_BLOCK * TI;
    const int n_tgt = a.i_hi - a.i_lo;
    if (n_tgt <= 0) return GRAV_B200_OK;
    a.NB = (n_tgt + IB - 1) / IB;
    a.NT = (a.n_src + DS_TJ - 1) / DS_TJ;
    if (a.NT == 0) {   // no sources at all: a = 0
        GB_CUDA(cudaMemsetAsync(a.acc + 3 * (size_t)a.i_lo, 0, sizeof(double) * 3 * (size_t)n_tgt, c->stream));
        return GRAV_B200_OK;
    }
    // Small source counts (and the massless method) run everything through the checked loop in one kernel;
    // otherwise fast kernel + special-tile kernel.
    static const int checked_max_tiles = getenv("GRAV_B200_DS_CHECKED_TILES") ? atoi(getenv("GRAV_B200_DS_CHECKED_TILES")) : 64;
    // With a softening length the self term needs no check at all: dx = dy = dz = 0 gives r2 = eps^2 > 0, a finite
    // s = m / eps^3 and the contribution fma(s, 0, a) = a exactly; padding sources (m = 0 at the origin) contribute
    // fma(0, dx, a) = a.  The fast loop then runs over every tile, at every N (eps >= 1e-30 keeps m / eps^3 far from
    // overflow).  Only eps = 0 needs the checked loop / the special-tile kernel.
    const bool softened = !MASSLESS && a.eps2 >= 1e-60;
    const bool all_checked = !softened && (MASSLESS || a.NT <= checked_max_tiles);
    a.skip_special = softened ? 0 : 1;
    // granules of 32 sources as the unit of the split while a CTA gets fewer than 16 tiles (N <~ 40000 on one GPU)
    long long grid = (long long)c->sm_count * 2;
    a.gpt = ((long long)a.NB * a.NT < 16 * grid) ? 8 : 1;
    const long long U = (long long)a.NB * a.NT * a.gpt;
    if (grid > DS_FIXUP_MAX_CONTRIB) grid = DS_FIXUP_MAX_CONTRIB;   // the fix-up kernel's per-block contributor table
    if (grid > U) grid = U;
    GB_TRY(c->partials.reserve((size_t)grid * 2 * 3 * IB * sizeof(double)));
    a.partials = c->partials.as<double>();
    if (all_checked) {
        if (a.gpt > 1) direct_sum_kernel<TI, true, MASSLESS, true><<<(unsigned)grid, DS_BLOCK, 0, c->stream>>>(a);
        else direct_sum_kernel<TI, true, MASSLESS, false><<<(unsigned)grid, DS_BLOCK, 0, c->stream>>>(a);
    } else {
        if (a.gpt > 1) direct_sum_kernel<TI, false, false, true><<<(unsigned)grid, DS_BLOCK, 0, c->stream>>>(a);
        else direct_sum_kernel<TI, false, false, false><<<(unsigned)grid, DS_BLOCK, 0, c->stream>>>(a);
    }
    GB_LAUNCH_CHECK();
    count_launch();
    // a fix-up is needed iff some target block is split, i.e. unless every CTA boundary is a block boundary
    bool split = false;
    for (long long k = 1; k < grid && !split; k++) split = (unit_begin(k, U, grid) % ((long long)a.NT * a.gpt)) != 0;
    if (split) {
        direct_sum_fixup_kernel<TI><<<dim3(a.NB, TI), DS_BLOCK, 0, c->stream>>>(a, (int)grid);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    if (!all_checked && !softened) {
        // one target per thread here: 4x the CTAs of the main decomposition, one or two tiles each
        direct_sum_special_kernel<IB><<<(n_tgt + DS_BLOCK - 1) / DS_BLOCK, DS_BLOCK, 0, c->stream>>>(a);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    return GRAV_B200_OK;
}

// Few targets: with 1024-target blocks a small system has fewer work units than the GPU has CTA slots (N = 4096: 64 units
// of 33 us each on 296 slots).  One target per thread gives four times as many, shorter units.
template <bool MASSLESS>
static int launch_direct_sum(grav_b200_ctx *c, DSArgs &a)
{
    const long long nb4 = (a.i_hi - a.i_lo + DS_BLOCK * DS_TI - 1) / (DS_BLOCK * DS_TI);
    const long long nt = (a.n_src + DS_TJ - 1) / DS_TJ;
    static const int ti1_units = getenv("GRAV_B200_DS_TI1_UNITS") ? atoi(getenv("GRAV_B200_DS_TI1_UNITS")) : 0;
    if (nb4 * nt < (ti1_units > 0 ? ti1_units : c->sm_count)) return launch_direct_sum_t<MASSLESS, 1>(c, a);
    return launch_direct_sum_t<MASSLESS, DS_TI>(c, a);
}

int direct_sum_pairwise(grav_b200_ctx *c, double eps)
{
    // large systems: every unordered pair once (direct_sum_sym.cu), 10 instead of 16 FP64 instructions per ordered interaction
    c->last_ds_sym = direct_sum_sym_wanted(c) ? 1 : 0;
    if (c->last_ds_sym) {
        const int rc = direct_sum_pairwise_sym(c, eps);
        if (rc != GRAV_B200_ENOMEM_SYM) return rc;
        c->last_ds_sym = 0;      // no room for the private accumulation arrays on this device: ordered interactions instead
    }
    DSArgs a{};
    a.src = c->posm.as<double4>();
    a.n_src = c->n;
    a.tgt = c->posm.as<double4>();
    a.i_lo = c->lo;
    a.i_hi = c->hi;
    a.eps2 = eps * eps;
    a.G = c->G;
    a.acc = c->acc.as<double>();
    return launch_direct_sum<false>(c, a);
}

// ---- small systems through the host-pointer API (config 1: N = 9 under IAS15 makes ~8 million force calls) --------
// One CTA, no staging copies: the kernel reads the caller's x and m from mapped pinned host memory, keeps all sources in
// shared memory and writes a[] straight back to mapped host memory, so a call is one launch and one synchronise
// (~10 us instead of ~80 us through the general path).  Same arithmetic as the main kernel, every pair checked for self.
constexpr int DS_SMALL_MAX = 256;     // one target per thread; beyond this the general path is faster (measured)

__global__ void __launch_bounds__(256) direct_sum_small_kernel(const double *__restrict__ x, const double *__restrict__ m, int n,
                                                              double G, double eps2, double *__restrict__ acc)
{
    __shared__ double4 src[DS_SMALL_MAX];
    for (int j = threadIdx.x; j < n; j += blockDim.x) src[j] = make_double4(x[3 * j], x[3 * j + 1], x[3 * j + 2], m[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double4 me = src[i];
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < n; j++) interaction<true>(src[j], src[j].w, j == i, me.x, me.y, me.z, eps2, ax, ay, az);
        acc[3 * i + 0] = G * ax;
        acc[3 * i + 1] = G * ay;
        acc[3 * i + 2] = G * az;
    }
}

// a, x, m: caller's host arrays.  Returns GRAV_B200_OK, or a negative value if the system is too large for this path.
int direct_sum_small_host(grav_b200_ctx *c, double *a, int n, const double *x, const double *m, double G, double eps)
{
    if (n > DS_SMALL_MAX) return -1;
    if (!c->small_pinned) {
        GB_CUDA(cudaHostAlloc((void **)&c->small_pinned, sizeof(double) * 7 * DS_SMALL_MAX, cudaHostAllocMapped));
    }
    double *hx = c->small_pinned, *hm = hx + 3 * DS_SMALL_MAX, *ha = hm + DS_SMALL_MAX;
    memcpy(hx, x, sizeof(double) * 3 * (size_t)n);
    memcpy(hm, m, sizeof(double) * (size_t)n);
    direct_sum_small_kernel<<<1, 256, 0, c->stream>>>(hx, hm, n, G, eps * eps, ha);   // UVA: host pointer == device pointer
    GB_LAUNCH_CHECK();
    count_launch();
    GB_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(a, ha, sizeof(double) * 3 * (size_t)n);
    return GRAV_B200_OK;
}

// ---- massless method ------------------------------------------------------------------
// Reference semantics (src/acceleration.c:236-367): sources are the particles with m != 0, in
// index order ("massive list").  A massive target feels every other massive particle with its
// true mass.  A massless target feels massive particle number r of that list with mass m[r]
// -- the reference indexes m by list rank, not by particle id (:357-359) -- reproduced here.
__global__ void massless_flags_kernel(const double4 *__restrict__ posm, int n, int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (posm[i].w != 0.0) ? 1 : 0;
}

__global__ void massless_compact_kernel(const double4 *__restrict__ posm, int n, const int *__restrict__ flag,
                                        const int *__restrict__ rank, double4 *__restrict__ src,
                                        int *__restrict__ src_id, double *__restrict__ src_altm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        const int r = rank[i];
        src[r] = posm[i];
        src_id[r] = i;
        src_altm[r] = posm[r].w;   // m[rank]
    }
}

int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp);   // scan.cu

// Massive-particle list of the resident system: c->msrc (packed records), c->msrc_id (particle ids, index order),
// c->msrc_altm (m[rank] quirk masses), c->mrank (list position of every massive particle).  Shared by the massless
// direct sum and the WHFast massless kernel.
int massive_list(grav_b200_ctx *c, int *n_massive_out)
{
    const int n = c->n;
    GB_TRY(c->mflag.reserve(sizeof(int) * (size_t)(n + 1)));
    GB_TRY(c->mrank.reserve(sizeof(int) * (size_t)(n + 1)));
    int *flag = c->mflag.as<int>();
    int *rank = c->mrank.as<int>();
    const int nb = (n + 255) / 256;
    int n_massive = c->mlist_count;
    if (!c->mlist_valid) {
        // which particles are massive only changes with the masses (set_system): inside a resident integration the flags,
        // ranks and the count are reused and every later call just refreshes the packed source records
        GB_CUDA(cudaMemsetAsync(flag + n, 0, sizeof(int), c->stream));
        massless_flags_kernel<<<nb, 256, 0, c->stream>>>(c->posm.as<double4>(), n, flag);
        GB_LAUNCH_CHECK();
        count_launch();
        GB_TRY(exclusive_scan_int(c, flag, rank, n + 1, c->misc));
        GB_CUDA(cudaMemcpyAsync(&n_massive, rank + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        c->mlist_count = n_massive;
    }
    const int pad = ((n_massive + SRC_PAD - 1) / SRC_PAD) * SRC_PAD;
    GB_TRY(c->msrc.reserve(sizeof(double4) * (size_t)(pad ? pad : SRC_PAD)));
    GB_TRY(c->msrc_id.reserve(sizeof(int) * (size_t)(pad ? pad : SRC_PAD)));
    GB_TRY(c->msrc_altm.reserve(sizeof(double) * (size_t)(pad ? pad : SRC_PAD)));
    if (pad) {
        if (!c->mlist_valid) {      // padding entries; the real ones are rewritten by every compaction
            GB_CUDA(cudaMemsetAsync(c->msrc.p, 0, sizeof(double4) * (size_t)pad, c->stream));
            GB_CUDA(cudaMemsetAsync(c->msrc_id.p, 0xff, sizeof(int) * (size_t)pad, c->stream));
            GB_CUDA(cudaMemsetAsync(c->msrc_altm.p, 0, sizeof(double) * (size_t)pad, c->stream));
        }
        massless_compact_kernel<<<nb, 256, 0, c->stream>>>(c->posm.as<double4>(), n, flag, rank, c->msrc.as<double4>(),
                                                         c->msrc_id.as<int>(), c->msrc_altm.as<double>());
        GB_LAUNCH_CHECK();
        count_launch();
    }
    c->mlist_valid = true;
    *n_massive_out = n_massive;
    return GRAV_B200_OK;
}

int direct_sum_massless(grav_b200_ctx *c, double eps)
{
    int n_massive = 0;
    GB_TRY(massive_list(c, &n_massive));
    DSArgs a{};
    a.src = c->msrc.as<double4>();
    a.src_id = c->msrc_id.as<int>();
    a.src_altm = c->msrc_altm.as<double>();
    a.n_src = n_massive;
    a.tgt = c->posm.as<double4>();
    a.i_lo = c->lo;
    a.i_hi = c->hi;
    a.eps2 = eps * eps;
    a.G = c->G;
    a.acc = c->acc.as<double>();
    return launch_direct_sum<true>(c, a);
}

}  // namespace gb
