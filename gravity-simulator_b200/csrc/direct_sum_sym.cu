// FP64 direct sum, every unordered pair evaluated ONCE (Newton's third law), on sm_100a.
//
// The reference's acceleration_pairwise (src/acceleration.c:198-231) loops over i < j and applies each pair to both
// particles.  direct_sum.cu evaluates every ordered interaction instead (16 FP64-pipe instructions each), which is the
// simple GPU formulation but does the r^-3 work twice.  Here a pair costs 20 instructions for BOTH directions
//   3 DADD (dx,dy,dz)  3 DFMA (r2)  6 (r2^-3/2 from the MUFU.RSQ64H seed, one cubic correction)
//   2 DMUL (w*m_j, w*m_i)  3 DFMA (a_i += s_j d)  3 DFMA (a_j -= s_i d)
// i.e. 10 per ordered interaction (9 when all masses are equal: the two DMULs go and the common mass is applied at
// the end).
//
// What makes the j-side sum cheap is a register rotation.  A warp holds a ROW of 256 particles (8 per lane: positions,
// masses and 24 accumulators in registers) and meets one GROUP of 32 j-particles at a time: lane l starts with
// j-particle l, and after every step (8 pairs per lane) the three j-accumulators move one lane down with SHFL while the
// lane reads the j-position of the step after from the warp's shared-memory slice.  After 32 steps every lane has met
// every j-particle, the j-accumulators are back in their home lane and complete for this row, and go out with three
// fire-and-forget RED.ADD.F64.  No cross-lane reduction, no block barrier in the loop.
//
// Decomposition: units = (row A of 256 particles) x (group g of 32 particles beyond the row, g >= 8(A+1)), flattened
// row-major and split EVENLY over a persistent grid of one 256-thread CTA per SM (and over the ranks of a multi-GPU
// run).  The 8 warps of a CTA all hold the same row and take the groups of the CTA's segment round-robin; the row's
// accumulators are combined through shared memory in warp order when the segment ends.  Everything a CTA produces goes
// into a PRIVATE accumulation array (3 n doubles per CTA), so the order of floating-point additions is fixed by the
// decomposition: same input, same grid -> same bits.  A finishing kernel adds the private arrays in CTA order, adds
// the pairs inside each 256-particle block (the diagonal of the pair matrix, ordered interactions), applies G and
// leaves the arrays zeroed for the next call.  With several ranks each one does its share of the units and of the
// diagonal blocks for ALL particles and the results are added with one all-reduce of 24 n bytes.
//
// Tuning (profiles/r2_sym_variants.txt; the SY_*_ macros are the knobs of scripts/build_variant.sh): 8 particles per lane at 240 registers and 8 warps per SM beat 4 per lane at
// 128 registers and 16 warps by 8 % -- the FP64 chains of 8 independent pairs hide the pipe latency better than more
// warps do; reading the next j-position one step ahead is worth 7 % of that.
#include "internal.cuh"

namespace gb {

#ifndef SY_TI_
#define SY_TI_ 8
#endif
#ifndef SY_THREADS_
#define SY_THREADS_ 256
#endif
#ifndef SY_UNROLL_
#define SY_UNROLL_ 1
#endif
constexpr int SY_TI = SY_TI_;            // particles per lane in a row
constexpr int SY_ROW = 32 * SY_TI;       // 256
constexpr int SY_GPR = SY_ROW / 32;      // groups per row width (8)
constexpr int SY_THREADS = SY_THREADS_;
constexpr int SY_UNROLL = SY_UNROLL_;    // steps of the rotation loop per trip
constexpr int SY_WARPS = SY_THREADS / 32;

struct SymArgs {
    const double4 *posm;     // packed (x,y,z,m), zero padded to a multiple of 256
    int n;                   // real particles
    int NR, NG;              // rows of SY_ROW, groups of 32
    long long U;             // units in total: sum over rows A <= NR-2 of NG - SY_GPR (A+1)
    int cta0, ctas_total;    // this launch runs CTAs cta0 .. cta0 + gridDim.x - 1 of ctas_total (ranks share the unit range)
    double eps2, G;
    double *priv;            // [gridDim.x][stride] private sums, AoS by particle, all zero on entry
    long long stride;        // doubles per CTA (3 * n_pad)
    const int *eqm_flag;     // device: 1 when all real masses are equal (then *eqm_mass is that mass)
    const double *eqm_mass;
    double *acc;             // finish kernel: AoS [3n]
    int pad_in_last_group;   // n is not a multiple of 32
    int diag_rank, diag_world;   // finish kernel: this rank adds the in-block pairs of blocks b with b % world == rank
};

__host__ __device__ inline long long sym_row_start(long long A, long long NG) { return A * NG - SY_GPR * (A * (A + 1) / 2); }
__host__ __device__ inline long long sym_unit_begin(long long c, long long U, long long C) { return (c * U) / C; }
// largest A in [0, NR-2] with row_start(A) <= u   (callers guarantee 0 <= u < U, NR >= 2)
__host__ __device__ inline int sym_row_of(long long u, int NG, int NR)
{
    int lo = 0, hi = NR - 2;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sym_row_start(mid, NG) <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// r2^(-3/2) without the mass: 6 FP64-pipe instructions + 1 MUFU (see internal.cuh, inv_r3_times_m)
__device__ __forceinline__ double inv_r3(double r2)
{
    const double y = rsqrt_seed(r2);
    const double t = y * y;
    const double e = fma(-r2, t, 1.0);
    const double y3 = y * t;
    const double p = fma(1.875, e, 1.5);
    const double q = e * p;
    return fma(q, y3, y3);
}

template <bool EQM, bool CHECK>
__device__ __forceinline__ void sym_pair(double xi, double yi, double zi, double mi, const double4 pj, bool valid, double eps2,
                                         double &aix, double &aiy, double &aiz, double &ajx, double &ajy, double &ajz)
{
    const double dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    double w = inv_r3(r2);
    if (CHECK) {
        if (!valid) w = 0.0;
    }
    const double sj = EQM ? w : w * pj.w;
    const double si = EQM ? w : w * mi;
    aix = fma(sj, dx, aix);
    aiy = fma(sj, dy, aiy);
    aiz = fma(sj, dz, aiz);
    ajx = fma(-si, dx, ajx);
    ajy = fma(-si, dy, ajy);
    ajz = fma(-si, dz, ajz);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One group of 32 j-particles against the row: 32 steps of SY_TI pairs per lane, the j-accumulators rotating through the lanes.
template <bool EQM, bool CHECK>
__device__ __forceinline__ void sym_group(const double4 *__restrict__ sl, int jbase, int n, const double (&xi)[SY_TI], const double (&yi)[SY_TI],
                                          const double (&zi)[SY_TI], const double (&mi)[SY_TI], double eps2, double (&ax)[SY_TI],
                                          double (&ay)[SY_TI], double (&az)[SY_TI], double &ajx, double &ajy, double &ajz)
{
    const int lane = threadIdx.x & 31;
    double4 pn = sl[lane];                   // the j-particle of a step is read one step ahead (+7 %)
#pragma unroll SY_UNROLL
    for (int s = 0; s < 32; s++) {
        const int jj = (lane + s) & 31;
        const double4 pj = pn;
        pn = sl[(lane + s + 1) & 31];
        const bool jv = CHECK ? (jbase + jj < n) : true;
#pragma unroll
        for (int t = 0; t < SY_TI; t++)
            sym_pair<EQM, CHECK>(xi[t], yi[t], zi[t], mi[t], pj, jv, eps2, ax[t], ay[t], az[t], ajx, ajy, ajz);
        // the accumulators of j-particle (lane + s) & 31 move on to the lane that meets it next
        ajx = __shfl_sync(0xffffffffu, ajx, (lane + 1) & 31);
        ajy = __shfl_sync(0xffffffffu, ajy, (lane + 1) & 31);
        ajz = __shfl_sync(0xffffffffu, ajz, (lane + 1) & 31);
    }
}

// One segment = groups [ka, kb) of row A (absolute group SY_GPR (A+1) + k), shared round-robin by the warps of the CTA.
template <bool EQM>
__device__ __forceinline__ void sym_segment(const SymArgs &p, int A, int ka, int kb, double4 *slice, double *part, double *P)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double xi[SY_TI], yi[SY_TI], zi[SY_TI], mi[SY_TI], ax[SY_TI], ay[SY_TI], az[SY_TI];
#pragma unroll
    for (int t = 0; t < SY_TI; t++) {
        const int i = A * SY_ROW + t * 32 + lane;
        const double4 q = p.posm[i];
        xi[t] = q.x; yi[t] = q.y; zi[t] = q.z; mi[t] = q.w;
        ax[t] = 0.0; ay[t] = 0.0; az[t] = 0.0;
    }
    const int g0 = SY_GPR * (A + 1);
    int k = ka + warp, buf = 0;
    if (k < kb) {
        const double4 *src = p.posm + (size_t)(g0 + k) * 32 + lane;
        cp_async16(&slice[lane].x, &src->x);
        cp_async16(&slice[lane].z, &src->z);
    }
    cp_async_commit();
    for (; k < kb; k += SY_WARPS) {
        const int kn = k + SY_WARPS;
        if (kn < kb) {
            const double4 *src = p.posm + (size_t)(g0 + kn) * 32 + lane;
            cp_async16(&slice[(buf ^ 1) * 32 + lane].x, &src->x);
            cp_async16(&slice[(buf ^ 1) * 32 + lane].z, &src->z);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const double4 *sl = slice + buf * 32;
        const int jbase = (g0 + k) * 32;
        double ajx = 0.0, ajy = 0.0, ajz = 0.0;
        // only the last group of the system can hold padding (zero-mass records at the origin: harmless with a
        // softening length and true masses, but not when the masses are factored out or a real particle can sit at
        // distance 0 from them)
        if (p.pad_in_last_group && g0 + k == p.NG - 1) sym_group<EQM, true>(sl, jbase, p.n, xi, yi, zi, mi, p.eps2, ax, ay, az, ajx, ajy, ajz);
        else sym_group<EQM, false>(sl, jbase, p.n, xi, yi, zi, mi, p.eps2, ax, ay, az, ajx, ajy, ajz);
        // home again: lane holds the sums of j-particle `lane` of this group over the whole row
        double *dst = P + 3 * (size_t)(jbase + lane);
        atomicAdd(dst + 0, ajx);
        atomicAdd(dst + 1, ajy);
        atomicAdd(dst + 2, ajz);
        __syncwarp();      // everybody is done with this buffer before the copy of the group after next lands in it
        buf ^= 1;
    }
    cp_async_wait<0>();
    // row sums: the warps' partial sums are added in warp order (deterministic), then go to the private array
#pragma unroll
    for (int t = 0; t < SY_TI; t++) {
        double *d = part + ((size_t)warp * SY_ROW + t * 32 + lane) * 3;
        d[0] = ax[t]; d[1] = ay[t]; d[2] = az[t];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < SY_ROW * 3; e += SY_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < SY_WARPS; w++) s += part[(size_t)w * SY_ROW * 3 + e];
        atomicAdd(P + 3 * (size_t)A * SY_ROW + e, s);
    }
    // the next segment's (or nobody's) additions to the same addresses must come after these
    __threadfence();
    __syncthreads();
}

__global__ void __launch_bounds__(SY_THREADS, 1) direct_sum_sym_kernel(const SymArgs p)
{
    extern __shared__ double4 sy_smem[];
    const int warp = threadIdx.x >> 5;
    double4 *slice = sy_smem + warp * 64;                        // two buffers of 32 j-particles per warp
    double *part = reinterpret_cast<double *>(sy_smem + SY_WARPS * 64);   // [warps][SY_ROW][3]
    const long long cg = p.cta0 + blockIdx.x;
    long long u = sym_unit_begin(cg, p.U, p.ctas_total);
    const long long u1 = sym_unit_begin(cg + 1, p.U, p.ctas_total);
    double *P = p.priv + (size_t)blockIdx.x * p.stride;
    const bool eqm = (*p.eqm_flag != 0);
    while (u < u1) {
        const int A = sym_row_of(u, p.NG, p.NR);
        const long long rs = sym_row_start(A, p.NG);
        const int L = p.NG - SY_GPR * (A + 1);
        const int ka = (int)(u - rs);
        const long long rem = u1 - rs;
        const int kb = rem < (long long)L ? (int)rem : L;
        if (eqm) sym_segment<true>(p, A, ka, kb, slice, part, P);
        else sym_segment<false>(p, A, ka, kb, slice, part, P);
        u = rs + kb;
    }
}

// Whether CTA cg of the decomposition added anything to the entries of block b (particles [b SY_ROW, (b+1) SY_ROW)): as one
// of its rows (i-side) or inside the group ranges it met (j-side).  May say yes for a CTA that did not (reading zeros is
// harmless); never says no for one that did -- the finishing kernel leaves exactly the touched entries zeroed.
__host__ __device__ inline bool sym_cta_touches(const SymArgs &p, long long cg, int b)
{
    const long long u0 = sym_unit_begin(cg, p.U, p.ctas_total), u1 = sym_unit_begin(cg + 1, p.U, p.ctas_total);
    if (u0 >= u1) return false;
    const int A0 = sym_row_of(u0, p.NG, p.NR), A1 = sym_row_of(u1 - 1, p.NG, p.NR);
    if (b < A0) return false;
    if (b <= A1) return true;                     // one of its rows
    if (A1 - A0 >= 2) return true;                // row A0 + 1 is covered completely and lies before b
    // group range of block b in the k-coordinate of row A:  [SY_GPR (b - A - 1), SY_GPR (b - A - 1) + SY_GPR)
    const long long ka0 = u0 - sym_row_start(A0, p.NG);
    const long long kend0 = (A1 > A0) ? (long long)(p.NG - SY_GPR * (A0 + 1)) : (u1 - sym_row_start(A0, p.NG));
    const long long kb_lo0 = (long long)SY_GPR * (b - A0 - 1);
    if (kb_lo0 < kend0 && kb_lo0 + SY_GPR > ka0) return true;
    if (A1 > A0) {
        const long long kend1 = u1 - sym_row_start(A1, p.NG);
        if ((long long)SY_GPR * (b - A1 - 1) < kend1) return true;
    }
    return false;
}

// acc[i] = G * (sum of the private arrays that can hold something for i, in CTA order, + the pairs inside i's own
// block of SY_ROW particles); the private entries are zeroed on the way.  One thread per particle, one CTA per block.
__global__ void __launch_bounds__(SY_ROW) direct_sum_sym_finish_kernel(const SymArgs p, int ctas_local)
{
    __shared__ double4 blk[SY_ROW];
    __shared__ unsigned char s_touch[1024];          // per local CTA: did it add anything to this block's entries?
    const int b = blockIdx.x, tid = threadIdx.x;
    const int i = b * SY_ROW + tid;
    const bool inb = 3 * (long long)i < p.stride;      // inside the padded arrays
    blk[tid] = inb ? p.posm[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    for (int c = tid; c < ctas_local; c += SY_ROW) s_touch[c] = sym_cta_touches(p, (long long)p.cta0 + c, b) ? 1 : 0;
    __syncthreads();
    const int cmax = ctas_local - 1;
    // the block's 3 * SY_ROW sums as flat, fully coalesced columns (thread t takes elements t, t + SY_ROW, t + 2 SY_ROW of
    // every contributing array, four arrays in flight), then regrouped per particle through shared memory
    __shared__ double s_sum[3 * SY_ROW];
    {
        double f0 = 0.0, f1 = 0.0, f2 = 0.0;
        const size_t e0 = 3 * (size_t)b * SY_ROW + tid;
        const bool in0 = (long long)(e0) < p.stride, in1 = (long long)(e0 + SY_ROW) < p.stride, in2 = (long long)(e0 + 2 * SY_ROW) < p.stride;
        int c = 0;
        for (; c + 4 <= cmax + 1; c += 4) {
            double v[4][3];
            bool tc[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double *P = p.priv + (size_t)(c + q) * p.stride + e0;
                tc[q] = s_touch[c + q] != 0;
                v[q][0] = (tc[q] && in0) ? P[0] : 0.0; v[q][1] = (tc[q] && in1) ? P[SY_ROW] : 0.0; v[q][2] = (tc[q] && in2) ? P[2 * SY_ROW] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                double *P = p.priv + (size_t)(c + q) * p.stride + e0;
                if (!tc[q]) continue;              // adding its zeros would not change a bit (x + 0.0 = x; the sums start at +0.0)
                f0 += v[q][0]; f1 += v[q][1]; f2 += v[q][2];
                if (in0) P[0] = 0.0;
                if (in1) P[SY_ROW] = 0.0;
                if (in2) P[2 * SY_ROW] = 0.0;
            }
        }
        for (; c <= cmax; c++) {
            if (!s_touch[c]) continue;
            double *P = p.priv + (size_t)c * p.stride + e0;
            if (in0) { f0 += P[0]; P[0] = 0.0; }
            if (in1) { f1 += P[SY_ROW]; P[SY_ROW] = 0.0; }
            if (in2) { f2 += P[2 * SY_ROW]; P[2 * SY_ROW] = 0.0; }
        }
        s_sum[tid] = f0; s_sum[tid + SY_ROW] = f1; s_sum[tid + 2 * SY_ROW] = f2;
    }
    __syncthreads();
    double sx = s_sum[3 * tid + 0], sy = s_sum[3 * tid + 1], sz = s_sum[3 * tid + 2];
    const bool eqm = (*p.eqm_flag != 0);
    if (eqm) {
        const double m0 = *p.eqm_mass;
        sx *= m0; sy *= m0; sz *= m0;
    }
    if (b % p.diag_world == p.diag_rank) {
        // ordered interactions inside the block, self term and padding masked (the checked form of direct_sum.cu)
        const double4 me = blk[tid];
        double dx_ = 0.0, dy_ = 0.0, dz_ = 0.0;
        const int jn = min(SY_ROW, p.n - b * SY_ROW);
        for (int j = 0; j < jn; j++) {
            const double4 pj = blk[j];
            const double dx = pj.x - me.x, dy = pj.y - me.y, dz = pj.z - me.z;
            double r2 = fma(dx, dx, p.eps2);
            r2 = fma(dy, dy, r2);
            r2 = fma(dz, dz, r2);
            double s = inv_r3_times_m(r2, pj.w);
            if (j == tid) s = 0.0;
            dx_ = fma(s, dx, dx_);
            dy_ = fma(s, dy, dy_);
            dz_ = fma(s, dz, dz_);
        }
        sx += dx_; sy += dy_; sz += dz_;
    }
    if (i < p.n) {
        p.acc[3 * (size_t)i + 0] = p.G * sx;
        p.acc[3 * (size_t)i + 1] = p.G * sy;
        p.acc[3 * (size_t)i + 2] = p.G * sz;
    }
}

// flag = 1 and mass = m[0] when every real particle has the same (non-zero, finite) mass
__global__ void sym_equal_mass_kernel(const double4 *__restrict__ posm, int n, int *__restrict__ flag, double *__restrict__ mass)
{
    const double m0 = posm[0].w;
    int differs = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) differs |= (posm[i].w != m0);
    if (differs) atomicAnd(flag, 0);      // any mismatch clears the flag (initialised to 1 by the caller)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *mass = m0;
        if (!(m0 != 0.0) || !(fabs(m0) < 1.79e308)) atomicAnd(flag, 0);
    }
}

// Whether this system takes the pair-once path: enough units per CTA for the even split to balance, and the private
// arrays within a memory budget.  grav_b200_set_direct_sum_mode() / GRAV_B200_DS_SYM force the choice.
bool direct_sum_sym_wanted(const grav_b200_ctx *c)
{
    const int mode = grav_b200_get_direct_sum_mode();
    if (mode == 0) return false;
    const long long n = c->n;
    if (n < 2 * SY_ROW) return false;
    const long long NR = (n + SY_ROW - 1) / SY_ROW, NG = (n + 31) / 32;
    const long long U = sym_row_start(NR - 1, NG);
    const long long ctas = (long long)c->sm_count * c->world;
    const size_t bytes = (size_t)c->sm_count * 3 * (size_t)c->n_pad * sizeof(double);
    static const long long budget_gb = getenv("GRAV_B200_DS_SYM_MAX_GB") ? atoll(getenv("GRAV_B200_DS_SYM_MAX_GB")) : 40;
    if (bytes > (size_t)budget_gb << 30) return false;
    if (mode == 1) return U >= 1;
    // measured crossover on one GPU (profiles/r2_sym_variants.txt): a tie at N = 8192 (27 units per CTA), +25 % at 12288 (61);
    // with several ranks the all-reduce of the results has to be paid for too
    static const long long min_units = getenv("GRAV_B200_DS_SYM_MIN_UNITS") ? atoll(getenv("GRAV_B200_DS_SYM_MIN_UNITS")) : 40;
    if (c->world > 1) return U >= 2 * min_units * ctas;
    return U >= min_units * ctas;
}

int direct_sum_pairwise_sym(grav_b200_ctx *c, double eps)
{
    SymArgs a{};
    a.posm = c->posm.as<double4>();
    a.n = c->n;
    a.NR = (c->n + SY_ROW - 1) / SY_ROW;
    a.NG = (c->n + 31) / 32;
    a.U = sym_row_start(a.NR - 1, a.NG);
    const int ctas = c->sm_count < 1024 ? c->sm_count : 1024;     // one per SM (the finishing kernel's contributor table holds 1024)
    a.cta0 = c->rank * ctas;
    a.ctas_total = c->world * ctas;
    a.diag_rank = c->rank;
    a.diag_world = c->world;
    // measurement hook: the share of rank 0 of a k-rank run on one GPU (results are partial sums then)
    static const int fake_world = getenv("GRAV_B200_DS_SYM_FAKE_WORLD") ? atoi(getenv("GRAV_B200_DS_SYM_FAKE_WORLD")) : 0;
    if (fake_world > 1 && c->world == 1) { a.ctas_total = fake_world * ctas; a.diag_world = fake_world; }
    a.eps2 = eps * eps;
    a.G = c->G;
    a.stride = 3 * (long long)c->n_pad;
    a.acc = c->acc.as<double>();
    const size_t bytes = (size_t)ctas * (size_t)a.stride * sizeof(double);
    // the arrays are all zero between calls as long as every call finishes (and zeroes) what it touched; a different
    // system size or rank layout starts from a fresh memset anyway
    const long long layout = ((long long)c->n << 20) ^ ((long long)a.ctas_total << 8) ^ a.cta0;
    if (c->sym_priv.cap < bytes || !c->sym_priv_clean || c->sym_layout != layout) {
        const int rc_alloc = c->sym_priv.reserve(bytes);
        if (rc_alloc != GRAV_B200_OK) {
            if (c->world > 1) return rc_alloc;     // the ranks must take the same path (collectives): report, do not switch
            cudaGetLastError();
            return GRAV_B200_ENOMEM_SYM;
        }
        GB_CUDA(cudaMemsetAsync(c->sym_priv.p, 0, c->sym_priv.cap, c->stream));
    }
    a.priv = c->sym_priv.as<double>();
    c->sym_priv_clean = false;
    c->sym_layout = layout;
    // equal masses?  Decided on the device whenever the masses may have changed (set_system), like the massive list
    GB_TRY(c->sym_flag.reserve(64));
    int *flag = c->sym_flag.as<int>();
    double *mass = reinterpret_cast<double *>(c->sym_flag.as<char>() + 8);
    static const int eqm_allowed = getenv("GRAV_B200_DS_EQUAL_MASS") ? atoi(getenv("GRAV_B200_DS_EQUAL_MASS")) : 1;
    if (!c->sym_eqm_valid) {
        GB_CUDA(cudaMemsetAsync(flag, 0, 16, c->stream));
        if (eqm_allowed) {
            GB_CUDA(cudaMemsetAsync(flag, 1, sizeof(int), c->stream));     // any non-zero value means "equal so far"
            sym_equal_mass_kernel<<<c->sm_count, 256, 0, c->stream>>>(a.posm, a.n, flag, mass);
            GB_LAUNCH_CHECK();
            count_launch();
        }
        c->sym_eqm_valid = true;
    }
    a.eqm_flag = flag;
    a.eqm_mass = mass;
    const size_t smem = sizeof(double4) * SY_WARPS * 64 + sizeof(double) * SY_WARPS * SY_ROW * 3;
    a.pad_in_last_group = (c->n % 32 != 0) ? 1 : 0;
    if (!c->sym_attr_set) {
        GB_CUDA(cudaFuncSetAttribute(direct_sum_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->sym_attr_set = true;
    }
    if (a.U > 0) {
        direct_sum_sym_kernel<<<ctas, SY_THREADS, smem, c->stream>>>(a);
        GB_LAUNCH_CHECK();
        count_launch();
    }
    direct_sum_sym_finish_kernel<<<a.NR, SY_ROW, 0, c->stream>>>(a, ctas);
    GB_LAUNCH_CHECK();
    count_launch();
    c->sym_priv_clean = true;
    if (c->world > 1) GB_TRY(comm_allreduce_sum(c, a.acc, 3 * c->n));
    return GRAV_B200_OK;
}

}  // namespace gb

// ---- test hooks (host code only: the decomposition the kernels use, so that it can be checked without a GPU) ------------
extern "C" {

// Segments of CTA `cta` of `ctas_total` for a system of n particles: up to max_segments triples (row, first group, end group)
// with absolute group indices; returns the number of segments, or -1 for a system without two rows.
int grav_b200_debug_pair_once_segments(int n, int ctas_total, int cta, int max_segments, int *rows, int *group_begin, int *group_end)
{
    using namespace gb;
    const int NR = (n + SY_ROW - 1) / SY_ROW, NG = (n + 31) / 32;
    if (NR < 2) return -1;
    const long long U = sym_row_start(NR - 1, NG);
    long long u = sym_unit_begin(cta, U, ctas_total);
    const long long u1 = sym_unit_begin(cta + 1LL, U, ctas_total);
    int count = 0;
    while (u < u1) {
        const int A = sym_row_of(u, NG, NR);
        const long long rs = sym_row_start(A, NG);
        const int L = NG - SY_GPR * (A + 1);
        const int ka = (int)(u - rs);
        const long long rem = u1 - rs;
        const int kb = rem < (long long)L ? (int)rem : L;
        if (count < max_segments) { rows[count] = A; group_begin[count] = SY_GPR * (A + 1) + ka; group_end[count] = SY_GPR * (A + 1) + kb; }
        count++;
        u = rs + kb;
    }
    return count;
}

// Whether the finishing kernel reads CTA `cta`'s private array for block `block` (SY_ROW particles): sym_cta_touches().
int grav_b200_debug_pair_once_touches(int n, int ctas_total, int cta, int block)
{
    using namespace gb;
    SymArgs a{};
    a.n = n;
    a.NR = (n + SY_ROW - 1) / SY_ROW;
    a.NG = (n + 31) / 32;
    if (a.NR < 2) return -1;
    a.U = sym_row_start(a.NR - 1, a.NG);
    a.cta0 = 0;
    a.ctas_total = ctas_total;
    return sym_cta_touches(a, cta, block) ? 1 : 0;
}

int grav_b200_debug_pair_once_row(void) { return gb::SY_ROW; }

}  // extern "C"
