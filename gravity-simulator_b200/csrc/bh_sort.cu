// Barnes-Hut stage BH-1..BH-3 on the device: bounding box, 63-bit Morton keys, stable LSD radix sort.
//
// Reference: calculate_bounding_box (src/linear_octree.c:113-146), compute_3d_particle_morton_indices_
// deepest_level (:159-202), radix_sort_particles_morton_index (:216-326).  The reference sorts with 7
// passes of 9 bits; any stable sort by key gives the same permutation, so we use 8 passes of 8 bits with
// warp-ballot ranking (match.any) -- integer work, bit-exact by construction.
#include "internal.cuh"

namespace gb {

// ---- bounding box ---------------------------------------------------------------------------------
// fmin/fmax are order independent, so a parallel reduction reproduces the serial loop exactly.  Doubles
// are mapped to order-preserving int64 so the cross-block step can use integer atomics.
__device__ __forceinline__ long long f64_to_ordered(double v)
{
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double ordered_to_f64(long long o)
{
    return __longlong_as_double(o >= 0 ? o : (o ^ 0x7fffffffffffffffLL));
}

__global__ void bbox_init_kernel(long long *mm)
{
    if (threadIdx.x < 3) mm[threadIdx.x] = 0x7fffffffffffffffLL;           // running minima
    else if (threadIdx.x < 6) mm[threadIdx.x] = (long long)0x8000000000000000ULL;  // running maxima
}

__global__ void __launch_bounds__(256) bbox_reduce_kernel(const double4 *__restrict__ posm, int n, long long *mm)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double4 q = posm[i];
        lo[0] = fmin(lo[0], q.x); hi[0] = fmax(hi[0], q.x);
        lo[1] = fmin(lo[1], q.y); hi[1] = fmax(hi[1], q.y);
        lo[2] = fmin(lo[2], q.z); hi[2] = fmax(hi[2], q.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&mm[k], f64_to_ordered(lo[k]));
            atomicMax(&mm[3 + k], f64_to_ordered(hi[k]));
        }
    }
}

// box[0..2] = center, box[3] = width; same expressions as src/linear_octree.c:137-145
__global__ void bbox_finalize_kernel(const long long *mm, double *box)
{
    double lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = ordered_to_f64(mm[k]); hi[k] = ordered_to_f64(mm[3 + k]); }
    for (int k = 0; k < 3; k++) box[k] = __ddiv_rn(__dadd_rn(hi[k], lo[k]), 2.0);
    const double wx = __dsub_rn(hi[0], lo[0]), wy = __dsub_rn(hi[1], lo[1]), wz = __dsub_rn(hi[2], lo[2]);
    box[3] = fmax(fmax(wx, wy), wz);
}

// ---- Morton keys ------------------------------------------------------------------------------------
__device__ __forceinline__ long long spread3(long long v)
{
    v &= 0x1fffffLL;
    v = (v | v << 32) & 0x1f00000000ffffLL;
    v = (v | v << 16) & 0x1f0000ff0000ffLL;
    v = (v | v << 8) & 0x100f00f00f00f00fLL;
    v = (v | v << 4) & 0x10c30c30c30c30c3LL;
    v = (v | v << 2) & 0x1249249249249249LL;
    return v;
}

// u = (x - c)/w + 0.5 with IEEE division (no reciprocal, no FMA), n = (int64)(u * 2^21) & 0x1fffff.
// The cast is cvt.rzi.s64.f64; NaN (w == 0) converts to 0 here and to INT64_MIN on x86, both masked to 0.
// Out-of-range values (user-supplied box smaller than the data) saturate here but give INT64_MIN on x86:
// emulate that so the masked result agrees.
__device__ __forceinline__ long long cell_index(double x, double c, double w)
{
    const double u = __dadd_rn(__ddiv_rn(__dsub_rn(x, c), w), 0.5);
    const double s = __dmul_rn(u, 2097152.0);
    long long n;
    if (!(s > -9.2233720368547758e18 && s < 9.2233720368547758e18)) n = (long long)0x8000000000000000ULL;
    else n = (long long)s;
    return n;
}

__global__ void __launch_bounds__(256) morton_kernel(const double4 *__restrict__ posm, int n, const double *__restrict__ box,
                                                     long long *__restrict__ keys_unsorted, long long *__restrict__ keys,
                                                     int *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cx = box[0], cy = box[1], cz = box[2], w = box[3];
    const double4 q = posm[i];
    const long long k = spread3(cell_index(q.x, cx, w)) | (spread3(cell_index(q.y, cy, w)) << 1) |
                        (spread3(cell_index(q.z, cz, w)) << 2);
    if (keys_unsorted) keys_unsorted[i] = k;
    keys[i] = k;
    perm[i] = i;
}

// ---- stable LSD radix sort of (key, index) pairs ------------------------------------------------------
// 8 passes of 8 bits.  A CTA of 256 threads owns a tile of 4096 consecutive pairs; warp w owns the contiguous
// sub-chunk [512 w, 512 (w+1)) and walks it 32 pairs at a time, in order.
//   pass A (sort_hist_kernel)    per tile: digit counts                       -> hist[digit][tile]
//   scan                         exclusive prefix sum in (digit, tile) order  =  first output slot of each group
//   pass B (sort_scatter_kernel) per tile: (1) per-warp digit counts (match.any, warp-private rows, no atomics),
//                                (2) 256 threads turn them into local slots: exclusive over warps, then over digits,
//                                (3) every pair is placed in shared memory at  slot = warp_start[w][d] + rank
//                                    (rank = popc of lower lanes with the same digit) -- the tile is now sorted by
//                                    digit, stably -- (4) thread i streams slot i to  gbase[d] + i : consecutive
//                                    threads write consecutive addresses inside each digit run (avg 16 pairs).
// Output order inside a digit value = tile order, then warp order, then position order: stable.
int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp);
constexpr int SORT_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                              // pairs per thread: 4096-pair tiles (large n)
constexpr int SORT_ROUNDS_SMALL = 4;                         // 1024-pair tiles, n <= SORT_SMALL_MAX_N
constexpr int SORT_SMALL_MAX_N = 128 * SORT_THREADS * SORT_ROUNDS_SMALL;   // SMALL_MAX_TILES tiles   // <= 128 tiles: every CTA scans the histogram itself
// Small-n variant (FUSED): a 1e5-particle sort has only 25 of the big tiles -- a sixth of the SMs -- and its three scan
// launches per pass cost more than the sort kernels.  With 1024-pair tiles the work spreads over ~100 CTAs, the histogram
// is laid out [tile][digit] and each scatter CTA derives its own output offsets from it (<= 128 x 256 counters, read
// coalesced from L2), so a pass is two launches instead of five.

__device__ __forceinline__ unsigned match_digit(int d, bool valid)
{
#ifdef SORT_BALLOT_MATCH
    unsigned same = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < SORT_BITS; b++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
        same &= ((d >> b) & 1) ? bal : ~bal;
    }
    return same;
#else
    (void)valid;
    return __match_any_sync(0xffffffffu, d);
#endif
}
__device__ __forceinline__ int digit_of(long long k, int shift) { return (int)((unsigned long long)k >> shift) & (SORT_RADIX - 1); }

template <int ROUNDS, bool FUSED>
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const long long *__restrict__ keys, int n, int shift,
                                                                 int num_tiles, int *__restrict__ hist)
{
    constexpr int SORT_TILE = SORT_THREADS * ROUNDS;
    __shared__ int cnt[SORT_RADIX];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int r = 0; r < ROUNDS; r++) {
        const int p = base + r * SORT_THREADS + threadIdx.x;      // order is irrelevant for counting: coalesced
        const bool valid = p < n;
        const int d = valid ? digit_of(keys[p], shift) : SORT_RADIX;
        const unsigned same = match_digit(d, valid);
        if (valid && (same >> lane) == 1u) atomicAdd(&cnt[d], __popc(same));   // one atomic per distinct digit per warp
    }
    __syncthreads();
    if (FUSED) hist[(size_t)blockIdx.x * SORT_RADIX + threadIdx.x] = cnt[threadIdx.x];
    else hist[(size_t)threadIdx.x * num_tiles + blockIdx.x] = cnt[threadIdx.x];
}

template <int ROUNDS, bool FUSED>
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const long long *__restrict__ keys_in,
                                                                    const int *__restrict__ vals_in, int n, int shift,
                                                                    int num_tiles, const int *__restrict__ offs,
                                                                    long long *__restrict__ keys_out,
                                                                    int *__restrict__ vals_out)
{
    constexpr int SORT_ROUNDS = ROUNDS;
    constexpr int SORT_TILE = SORT_THREADS * ROUNDS;
    constexpr int SORT_WCHUNK = 32 * ROUNDS;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    long long *skeys = reinterpret_cast<long long *>(sort_smem);
    int *svals = reinterpret_cast<int *>(sort_smem + (size_t)SORT_TILE * sizeof(long long));
    __shared__ int wcnt[SORT_WARPS][SORT_RADIX];
    __shared__ int gbase[SORT_RADIX];
    __shared__ int wsum[SORT_WARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int base = blockIdx.x * SORT_TILE;
    const int tile_n = min(SORT_TILE, n - base);
    for (int d = lane; d < SORT_RADIX; d += 32) wcnt[warp][d] = 0;

    // (1) load this warp's sub-chunk in order and count digits
    long long k[SORT_ROUNDS];
    int v[SORT_ROUNDS];
    const int wbase = warp * SORT_WCHUNK;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        k[r] = valid ? keys_in[base + li] : 0;
        v[r] = valid ? vals_in[base + li] : 0;
        const int d = valid ? digit_of(k[r], shift) : SORT_RADIX;
        const unsigned same = match_digit(d, valid);
        if (valid && (same >> lane) == 1u) wcnt[warp][d] += __popc(same);
        __syncwarp();
    }
    __syncthreads();

    // (2) local slots: thread d owns digit d
    {
        const int d = tid;
        int run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const int t = wcnt[w][d];
            wcnt[w][d] = run;
            run += t;
        }
        // exclusive scan of the 256 digit totals
        int inc = run;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, s);
            if (lane >= s) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int woff = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) woff += (w < warp) ? wsum[w] : 0;
        const int dstart = woff + inc - run;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) wcnt[w][d] += dstart;
        if (!FUSED) {
            gbase[d] = offs[(size_t)d * num_tiles + blockIdx.x] - dstart;
        } else {
            // offs is the raw [tile][digit] histogram: first output slot of (d, this tile) = pairs with a smaller digit
            // anywhere + pairs with this digit in earlier tiles
            int tot = 0, before = 0;
            for (int t = 0; t < num_tiles; t++) {
                const int h = offs[(size_t)t * SORT_RADIX + d];
                tot += h;
                if (t < (int)blockIdx.x) before += h;
            }
            int ginc = tot;
#pragma unroll
            for (int s2 = 1; s2 < 32; s2 <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ginc, s2);
                if (lane >= s2) ginc += t;
            }
            __syncthreads();                   // wsum is reused
            if (lane == 31) wsum[warp] = ginc;
            __syncthreads();
            int goff = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) goff += (w < warp) ? wsum[w] : 0;
            gbase[d] = goff + ginc - tot + before - dstart;
        }
    }
    __syncthreads();

    // (3) place every pair at its slot of the digit-sorted tile
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        const int d = valid ? digit_of(k[r], shift) : SORT_RADIX;
        const unsigned same = match_digit(d, valid);
        if (valid) {
            const int slot = wcnt[warp][d] + __popc(same & lt);
            skeys[slot] = k[r];
            svals[slot] = v[r];
        }
        __syncwarp();
        if (valid && (same >> lane) == 1u) wcnt[warp][d] += __popc(same);
        __syncwarp();
    }
    __syncthreads();

    // (4) stream the sorted tile out
    for (int i = tid; i < tile_n; i += SORT_THREADS) {
        const long long kk = skeys[i];
        const int pos = gbase[digit_of(kk, shift)] + i;
        keys_out[pos] = kk;
        vals_out[pos] = svals[i];
    }
}


// ---- large n: one kernel per pass ("onesweep": chained scan with decoupled look-back) ---------------------------------
// ncu on the three-kernel pass at N = 2^24 (profiles/r1_sort_n16m.txt): both sort kernels sit on one saturated pipe
// (sm__throughput 91 % / 72 % with 9-11 % issue activity) -- MATCH.ANY retires about one warp instruction per 64 cycles
// per SM, and a pass executed three of them per 32 pairs (histogram kernel, counting and placement in the scatter
// kernel).  This version needs ONE: the peer mask of every round is kept in a register, the warp-local rank of a pair is
// fixed while counting (the leader lane hands out the running count of its digit), and the per-tile digit counts are
// published for the tiles behind it instead of being computed by a separate histogram kernel + scan:
//   upfront   sort_global_hist_kernel: digit totals of all 8 passes from one read of the keys (they do not depend on the
//             order of the pairs) -> first output slot of every digit value
//   per pass  tile id from an atomic ticket (so every tile only waits for tiles that already started), counts, then
//             thread d publishes (count | AGGREGATE) for digit d, walks back over earlier tiles adding their words until
//             it meets an inclusive PREFIX, publishes its own (prefix | PREFIX), and the tile is scattered as before.
// A status word carries flag and value together (2 + 30 bits), so no fence is needed.
constexpr unsigned OSW_FLAG_AGG = 1u << 30, OSW_FLAG_PREFIX = 2u << 30, OSW_VALUE_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(SORT_THREADS) sort_global_hist_kernel(const long long *__restrict__ keys, int n, int *__restrict__ ghist)
{
    __shared__ int h[8][SORT_RADIX];
    for (int i = threadIdx.x; i < 8 * SORT_RADIX; i += SORT_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int base = blockIdx.x * SORT_THREADS; base < n; base += gridDim.x * SORT_THREADS) {
        const int i = base + threadIdx.x;
        const bool valid = i < n;
        const unsigned long long k = valid ? (unsigned long long)keys[i] : 0ull;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const int d = (int)(k >> (8 * p)) & (SORT_RADIX - 1);
            // the high digits of sorted-ish keys are equal across the warp: one add instead of a 32-way conflict
            const int d0 = __shfl_sync(0xffffffffu, d, __ffs(act) - 1);
            if (__all_sync(0xffffffffu, !valid || d == d0)) {
                if (lane == __ffs(act) - 1) atomicAdd(&h[p][d0], __popc(act));
            } else if (valid) {
                atomicAdd(&h[p][d], 1);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * SORT_RADIX; i += SORT_THREADS) {
        const int c = (&h[0][0])[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

__device__ __forceinline__ unsigned ld_status(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int ROUNDS>
__global__ void __launch_bounds__(SORT_THREADS) sort_onesweep_kernel(const long long *__restrict__ keys_in,
                                                                     const int *__restrict__ vals_in, int n, int shift,
                                                                     const int *__restrict__ ghist_pass, unsigned *__restrict__ status,
                                                                     int *__restrict__ ticket, long long *__restrict__ keys_out,
                                                                     int *__restrict__ vals_out)
{
    constexpr int TILE = SORT_THREADS * ROUNDS;
    constexpr int WCHUNK = 32 * ROUNDS;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    long long *skeys = reinterpret_cast<long long *>(sort_smem);
    int *svals = reinterpret_cast<int *>(sort_smem + (size_t)TILE * sizeof(long long));
    __shared__ int wcnt[SORT_WARPS][SORT_RADIX];
    __shared__ int gbase[SORT_RADIX];
    __shared__ int wsum[SORT_WARPS], wsum2[SORT_WARPS];
    __shared__ int s_tile;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) s_tile = atomicAdd(ticket, 1);
    for (int d = lane; d < SORT_RADIX; d += 32) wcnt[warp][d] = 0;
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * TILE;
    const int tile_n = min(TILE, n - base);

    // (1) load this warp's sub-chunk in order, count digits, fix the warp-local rank of every pair
    long long k[ROUNDS];
    int v[ROUNDS];
    unsigned short rk[ROUNDS];
    const int wbase = warp * WCHUNK;
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        k[r] = valid ? keys_in[base + li] : 0;
        v[r] = valid ? vals_in[base + li] : 0;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        const int d = valid ? digit_of(k[r], shift) : SORT_RADIX;
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const int leader = 31 - __clz(same);
        int prev = 0;
        if (valid && lane == leader) {
            prev = wcnt[warp][d];
            wcnt[warp][d] = prev + __popc(same);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rk[r] = (unsigned short)(prev + __popc(same & lt));
        __syncwarp();
    }
    __syncthreads();

    // (2) thread d owns digit d: local slots, publish / look back, global base
    {
        const int d = tid;
        int run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const int t = wcnt[w][d];
            wcnt[w][d] = run;
            run += t;
        }
        unsigned *my = status + (size_t)tile * SORT_RADIX + d;
        st_status(my, (unsigned)run | (tile == 0 ? OSW_FLAG_PREFIX : OSW_FLAG_AGG));
        // exclusive scans over the digits: this tile's counts (local slots) and the global totals (first slot per digit)
        const int tot = ghist_pass[d];
        int inc = run, ginc = tot;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, s2);
            const int g = __shfl_up_sync(0xffffffffu, ginc, s2);
            if (lane >= s2) { inc += t; ginc += g; }
        }
        if (lane == 31) { wsum[warp] = inc; wsum2[warp] = ginc; }
        __syncthreads();
        int woff = 0, goff = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            woff += (w < warp) ? wsum[w] : 0;
            goff += (w < warp) ? wsum2[w] : 0;
        }
        const int dstart = woff + inc - run;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) wcnt[w][d] += dstart;
        int excl = 0;
        for (int t = tile - 1; t >= 0; t--) {
            const unsigned *p = status + (size_t)t * SORT_RADIX + d;
            unsigned sv;
            do { sv = ld_status(p); } while ((sv >> 30) == 0u);
            excl += (int)(sv & OSW_VALUE_MASK);
            if (sv & OSW_FLAG_PREFIX) break;
        }
        if (tile > 0) st_status(my, (unsigned)(excl + run) | OSW_FLAG_PREFIX);
        gbase[d] = (goff + ginc - tot) + excl - dstart;
    }
    __syncthreads();

    // (3) place every pair at its slot of the digit-sorted tile
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        if (li < tile_n) {
            const int slot = wcnt[warp][digit_of(k[r], shift)] + rk[r];
            skeys[slot] = k[r];
            svals[slot] = v[r];
        }
    }
    __syncthreads();

    // (4) stream the sorted tile out
    for (int i = tid; i < tile_n; i += SORT_THREADS) {
        const long long kk = skeys[i];
        const int pos = gbase[digit_of(kk, shift)] + i;
        keys_out[pos] = kk;
        vals_out[pos] = svals[i];
    }
}

// ---- small n (<= 128 tiles of 1024 pairs): one kernel per pass, every CTA reads every tile's digit counts -----------
// All CTAs of the grid are resident at once (<= 128 CTAs of 256 threads), so a tile can simply wait until the digit
// counts of ALL tiles have been published: status[tile][d] = count | VALID.  Thread d then knows the pairs with digit d
// in earlier tiles and in total; a block scan over the totals gives the digit's first output slot.  One launch per pass
// (the two-launch version above needs the histogram kernel to finish first); the waits are batched so the ~100 loads
// per thread overlap.
constexpr unsigned SMALL_VALID = 0x80000000u;
constexpr int SMALL_MAX_TILES = 128;

// One pass of one tile; shared by the one-pass kernel and the all-passes kernel below.  CROSS: the input was written by
// other CTAs earlier in this launch, so it is read past L1 (__ldcg).
template <int ROUNDS, bool CROSS>
__device__ __forceinline__ void sort_small_tile_pass(const long long *keys_in, const int *vals_in, int n, int shift, int num_tiles,
                                                     unsigned *status, long long *keys_out, int *vals_out)
{
    constexpr int TILE = SORT_THREADS * ROUNDS;
    constexpr int WCHUNK = 32 * ROUNDS;
    __shared__ long long skeys[TILE];
    __shared__ int svals[TILE];
    __shared__ int wcnt[SORT_WARPS][SORT_RADIX];
    __shared__ int gbase[SORT_RADIX];
    __shared__ int wsum[SORT_WARPS], wsum2[SORT_WARPS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int tile = blockIdx.x;
    const int base = tile * TILE;
    const int tile_n = min(TILE, n - base);
    for (int d = lane; d < SORT_RADIX; d += 32) wcnt[warp][d] = 0;
    __syncwarp();

    long long k[ROUNDS];
    int v[ROUNDS];
    unsigned short rk[ROUNDS];
    const int wbase = warp * WCHUNK;
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        k[r] = valid ? (CROSS ? __ldcg(keys_in + base + li) : keys_in[base + li]) : 0;
        v[r] = valid ? (CROSS ? __ldcg(vals_in + base + li) : vals_in[base + li]) : 0;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        const bool valid = li < tile_n;
        const int d = valid ? digit_of(k[r], shift) : SORT_RADIX;
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const int leader = 31 - __clz(same);
        int prev = 0;
        if (valid && lane == leader) {
            prev = wcnt[warp][d];
            wcnt[warp][d] = prev + __popc(same);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rk[r] = (unsigned short)(prev + __popc(same & lt));
        __syncwarp();
    }
    __syncthreads();

    {
        const int d = tid;
        int run = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const int t = wcnt[w][d];
            wcnt[w][d] = run;
            run += t;
        }
        st_status(status + (size_t)tile * SORT_RADIX + d, (unsigned)run | SMALL_VALID);
        // every tile's count of digit d ([tile][digit] layout: a warp reads 128 contiguous bytes per tile): batches of 32
        // independent loads (one L2 round trip each batch), then spin only on the stragglers
        int tot = 0, before = 0;
        for (int t0 = 0; t0 < num_tiles; t0 += 32) {
            unsigned w16[32];
#pragma unroll
            for (int j = 0; j < 32; j++) w16[j] = (t0 + j < num_tiles) ? ld_status(status + (size_t)(t0 + j) * SORT_RADIX + d) : SMALL_VALID;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                unsigned sv = w16[j];
                while (!(sv & SMALL_VALID)) sv = ld_status(status + (size_t)(t0 + j) * SORT_RADIX + d);
                const int h = (int)(sv & ~SMALL_VALID);
                tot += h;
                if (t0 + j < tile) before += h;
            }
        }
        int inc = run, ginc = tot;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, s2);
            const int g = __shfl_up_sync(0xffffffffu, ginc, s2);
            if (lane >= s2) { inc += t; ginc += g; }
        }
        if (lane == 31) { wsum[warp] = inc; wsum2[warp] = ginc; }
        __syncthreads();
        int woff = 0, goff = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            woff += (w < warp) ? wsum[w] : 0;
            goff += (w < warp) ? wsum2[w] : 0;
        }
        const int dstart = woff + inc - run;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) wcnt[w][d] += dstart;
        gbase[d] = (goff + ginc - tot) + before - dstart;
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        const int li = wbase + r * 32 + lane;
        if (li < tile_n) {
            const int slot = wcnt[warp][digit_of(k[r], shift)] + rk[r];
            skeys[slot] = k[r];
            svals[slot] = v[r];
        }
    }
    __syncthreads();
    for (int i = tid; i < tile_n; i += SORT_THREADS) {
        const long long kk = skeys[i];
        const int pos = gbase[digit_of(kk, shift)] + i;
        keys_out[pos] = kk;
        vals_out[pos] = svals[i];
    }
}

template <int ROUNDS>
__global__ void __launch_bounds__(SORT_THREADS) sort_small_pass_kernel(const long long *keys_in, const int *vals_in, int n, int shift,
                                                                       int num_tiles, unsigned *status, long long *keys_out,
                                                                       int *vals_out)
{
    sort_small_tile_pass<ROUNDS, false>(keys_in, vals_in, n, shift, num_tiles, status, keys_out, vals_out);
}

// The same eight passes inside ONE kernel: the wait for all tiles' counts already is a grid-wide rendezvous, a second
// one (atomic arrival counter) after the scatter lets the next pass start without a kernel boundary.  All CTAs are
// resident (<= 128; launched cooperatively outside graph captures), so the barriers cannot deadlock.  A graph node costs
// ~6 us on this path, a barrier ~2 us.
template <int ROUNDS>
__global__ void __launch_bounds__(SORT_THREADS) sort_small_all_kernel(long long *ka, int *pa, long long *kb, int *pb, int n, int num_tiles, unsigned *status_all,
                                                                      unsigned *bar)
{
    for (int pass = 0; pass < 8; pass++) {
        // buffers alternate
        sort_small_tile_pass<ROUNDS, true>((pass & 1) ? kb : ka, (pass & 1) ? pb : pa, n, pass * SORT_BITS, num_tiles,
                                           status_all + (size_t)pass * SMALL_MAX_TILES * SORT_RADIX, (pass & 1) ? ka : kb,
                                           (pass & 1) ? pa : pb);
        // grid barrier: every tile's output must be in memory before anybody reads it as the next pass's input
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(bar + pass, 1u);
            while (ld_status(bar + pass) < (unsigned)num_tiles) { }
        }
        __syncthreads();
    }
}

int radix_pass(grav_b200_ctx *c, const long long *kin, const int *vin, long long *kout, int *vout, int n, int shift);

// All 8 passes of a small sort (n <= SORT_SMALL_MAX_N): one memset + 8 launches; ka/pa hold the result.
int radix_sort_small(grav_b200_ctx *c, long long *ka, int *pa, long long *kb, int *pb, int n)
{
    DevTree &t = c->tree;
    constexpr int TILE = SORT_THREADS * SORT_ROUNDS_SMALL;
    const int num_tiles = (n + TILE - 1) / TILE;
    // Both kernels below wait for every tile of the grid, so the whole grid has to be resident at once.  They are
    // launched COOPERATIVELY: the runtime then either places all CTAs together or refuses the launch, and it never
    // interleaves two cooperative grids half-resident (two contexts or streams sorting on the same GPU at the same time
    // cannot starve each other).  A refusal, or a device too small for the grid (a MIG slice), falls back to the
    // histogram + scatter pair of launches per pass, which has no inter-CTA waits.  The residency estimate is kept per
    // context: it depends on the device.
    if (c->sort_resident_ctas < 0) {
        int per_sm_all = 0, per_sm_pass = 0;
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_all, sort_small_all_kernel<SORT_ROUNDS_SMALL>, SORT_THREADS, 0));
        GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_pass, sort_small_pass_kernel<SORT_ROUNDS_SMALL>, SORT_THREADS, 0));
        c->sort_resident_ctas = c->sm_count * (per_sm_all < per_sm_pass ? per_sm_all : per_sm_pass);
    }
    auto wait_free_passes = [&]() -> int {
        for (int pass = 0; pass < 8; pass++) {
            GB_TRY(radix_pass(c, ka, pa, kb, pb, n, pass * SORT_BITS));
            long long *tk = ka; ka = kb; kb = tk;
            int *tp = pa; pa = pb; pb = tp;
        }
        return GRAV_B200_OK;
    };
    if (num_tiles > c->sort_resident_ctas) return wait_free_passes();
    const size_t words = (size_t)8 * SMALL_MAX_TILES * SORT_RADIX + 8;      // 1 MiB of status words + 8 barrier counters
    GB_TRY(t.hist.reserve(sizeof(int) * words));
    unsigned *status = t.hist.as<unsigned>();
    GB_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * words, c->stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)num_tiles);
    cfg.blockDim = dim3(SORT_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c->stream;
    cudaLaunchAttribute coop[1];
    coop[0].id = cudaLaunchAttributeCooperative;
    coop[0].val.cooperative = 1;
    static const bool coop_off = getenv("GRAV_B200_SORT_COOP") && atoi(getenv("GRAV_B200_SORT_COOP")) == 0;   // A/B only
    // Not inside a stream capture: graphs with cooperative kernel nodes made every new WHFast context slower than the last
    // (launch_simulation_python of config 3: 87, 242, 639 ms for the same 400 steps; 80 ms flat without).  A captured launch is
    // replayed by the graph on this context's stream like the plain launch it was in round 1.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    GB_CUDA(cudaStreamIsCapturing(c->stream, &cap));
    cfg.attrs = coop;
    cfg.numAttrs = (coop_off || cap != cudaStreamCaptureStatusNone) ? 0 : 1;
    static const bool per_pass = getenv("GRAV_B200_SORT_SMALL_PER_PASS") && atoi(getenv("GRAV_B200_SORT_SMALL_PER_PASS")) != 0;
    if (!per_pass) {
        unsigned *barrier = status + words - 8;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, sort_small_all_kernel<SORT_ROUNDS_SMALL>, ka, pa, kb, pb, n, num_tiles, status, barrier);
        if (e == cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); return wait_free_passes(); }
        GB_CUDA(e);
        count_launch();
        return GRAV_B200_OK;        // even number of passes: the result is in ka / pa
    }
    for (int pass = 0; pass < 8; pass++) {
        unsigned *st = status + (size_t)pass * SMALL_MAX_TILES * SORT_RADIX;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, sort_small_pass_kernel<SORT_ROUNDS_SMALL>, (const long long *)ka, (const int *)pa, n,
                                                 pass * SORT_BITS, num_tiles, st, kb, pb);
        if (e == cudaErrorCooperativeLaunchTooLarge && pass == 0) { cudaGetLastError(); return wait_free_passes(); }
        GB_CUDA(e);
        count_launch();
        long long *tk = ka; ka = kb; kb = tk;
        int *tp = pa; pa = pb; pb = tp;
    }
    return GRAV_B200_OK;
}


// One stable pass on an 8-bit digit of arbitrary (key, value) arrays; scratch: t.hist / t.scan_tmp.
int radix_pass(grav_b200_ctx *c, const long long *kin, const int *vin, long long *kout, int *vout, int n, int shift)
{
    DevTree &t = c->tree;
    constexpr int TILE_BIG = SORT_THREADS * SORT_ROUNDS, TILE_SMALL = SORT_THREADS * SORT_ROUNDS_SMALL;
    constexpr size_t SMEM_BIG = (size_t)TILE_BIG * (sizeof(long long) + sizeof(int));       // staging: 48 KiB
    constexpr size_t SMEM_SMALL = (size_t)TILE_SMALL * (sizeof(long long) + sizeof(int));   // 12 KiB
    if (!c->sort_attr_scatter) {    // per device, hence per context
        GB_CUDA(cudaFuncSetAttribute(sort_scatter_kernel<SORT_ROUNDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BIG));
        c->sort_attr_scatter = true;
    }
    if (n <= SORT_SMALL_MAX_N) {
        const int num_tiles = (n + TILE_SMALL - 1) / TILE_SMALL;
        GB_TRY(t.hist.reserve(sizeof(int) * (size_t)SORT_RADIX * num_tiles));
        sort_hist_kernel<SORT_ROUNDS_SMALL, true><<<num_tiles, SORT_THREADS, 0, c->stream>>>(kin, n, shift, num_tiles, t.hist.as<int>());
        GB_LAUNCH_CHECK();
        sort_scatter_kernel<SORT_ROUNDS_SMALL, true><<<num_tiles, SORT_THREADS, SMEM_SMALL, c->stream>>>(kin, vin, n, shift, num_tiles,
                                                                                                      t.hist.as<int>(), kout, vout);
        GB_LAUNCH_CHECK();
        count_launch(2);
        return GRAV_B200_OK;
    }
    const int num_tiles = (n + TILE_BIG - 1) / TILE_BIG;
    GB_TRY(t.hist.reserve(sizeof(int) * (size_t)SORT_RADIX * num_tiles));
    sort_hist_kernel<SORT_ROUNDS, false><<<num_tiles, SORT_THREADS, 0, c->stream>>>(kin, n, shift, num_tiles, t.hist.as<int>());
    GB_LAUNCH_CHECK();
    count_launch();
    GB_TRY(exclusive_scan_int(c, t.hist.as<int>(), t.hist.as<int>(), SORT_RADIX * num_tiles, t.scan_tmp));
    sort_scatter_kernel<SORT_ROUNDS, false><<<num_tiles, SORT_THREADS, SMEM_BIG, c->stream>>>(kin, vin, n, shift, num_tiles,
                                                                                             t.hist.as<int>(), kout, vout);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

// All 8 passes of a large sort with the one-kernel-per-pass scheme above; ka/pa hold the result (even pass count).
template <int ROUNDS>
static int onesweep_sort_t(grav_b200_ctx *c, long long *ka, int *pa, long long *kb, int *pb, int n)
{
    DevTree &t = c->tree;
    constexpr int TILE = SORT_THREADS * ROUNDS;
    constexpr size_t SMEM = (size_t)TILE * (sizeof(long long) + sizeof(int));
    if (!c->sort_attr_onesweep) {   // per device, hence per context
        GB_CUDA(cudaFuncSetAttribute(sort_onesweep_kernel<ROUNDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        c->sort_attr_onesweep = true;
    }
    const int num_tiles = (n + TILE - 1) / TILE;
    // scratch layout (ints): [0, 2048) digit totals of the 8 passes, [2048, 2056) tickets, then 8 x num_tiles x 256 status words
    const size_t head = 8 * SORT_RADIX + 8;
    const size_t words = head + (size_t)8 * num_tiles * SORT_RADIX;
    GB_TRY(t.hist.reserve(sizeof(int) * words));
    int *scratch = t.hist.as<int>();
    GB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * words, c->stream));
    int gblocks = (n + SORT_THREADS * 8 - 1) / (SORT_THREADS * 8);
    if (gblocks > c->sm_count * 8) gblocks = c->sm_count * 8;
    sort_global_hist_kernel<<<gblocks, SORT_THREADS, 0, c->stream>>>(ka, n, scratch);
    GB_LAUNCH_CHECK();
    count_launch();
    for (int pass = 0; pass < 8; pass++) {
        unsigned *status = reinterpret_cast<unsigned *>(scratch + head) + (size_t)pass * num_tiles * SORT_RADIX;
        sort_onesweep_kernel<ROUNDS><<<num_tiles, SORT_THREADS, SMEM, c->stream>>>(ka, pa, n, pass * SORT_BITS, scratch + pass * SORT_RADIX,
                                                                                 status, scratch + 8 * SORT_RADIX + pass, kb, pb);
        GB_LAUNCH_CHECK();
        count_launch();
        long long *tk = ka; ka = kb; kb = tk;
        int *tp = pa; pa = pb; pb = tp;
    }
    return GRAV_B200_OK;
}
static int onesweep_sort(grav_b200_ctx *c, long long *ka, int *pa, long long *kb, int *pb, int n)
{
    // 4096-pair tiles: 2048- and 3072-pair tiles measured 24 % and 11 % slower at N = 2^24
    return onesweep_sort_t<SORT_ROUNDS>(c, ka, pa, kb, pb, n);
}

// Stable sort of n (key, value) pairs held in ka / pa with kb / pb as scratch; the result ends in ka / pa.
int radix_sort_buffers(grav_b200_ctx *c, long long *ka, int *pa, long long *kb, int *pb, int n)
{
    if (n <= SORT_SMALL_MAX_N) return radix_sort_small(c, ka, pa, kb, pb, n);
    return onesweep_sort(c, ka, pa, kb, pb, n);
}

// keys/perm sorted in place (8 passes ping-pong through keys_tmp/perm_tmp)
int radix_sort_pairs(grav_b200_ctx *c)
{
    DevTree &t = c->tree;
    const int n = t.n;
    GB_TRY(t.keys_tmp.reserve(sizeof(long long) * (size_t)n));
    GB_TRY(t.perm_tmp.reserve(sizeof(int) * (size_t)n));
    long long *ka = t.keys.as<long long>(), *kb = t.keys_tmp.as<long long>();
    int *pa = t.perm.as<int>(), *pb = t.perm_tmp.as<int>();
    static const bool three_kernel = getenv("GRAV_B200_SORT_THREE_KERNEL") && atoi(getenv("GRAV_B200_SORT_THREE_KERNEL")) != 0;
    if (n > SORT_SMALL_MAX_N && !three_kernel) return onesweep_sort(c, ka, pa, kb, pb, n);
    static const bool two_launch = getenv("GRAV_B200_SORT_SMALL_TWO_LAUNCH") && atoi(getenv("GRAV_B200_SORT_SMALL_TWO_LAUNCH")) != 0;
    if (n <= SORT_SMALL_MAX_N && !two_launch) return radix_sort_small(c, ka, pa, kb, pb, n);
    for (int pass = 0; pass < 8; pass++) {
        GB_TRY(radix_pass(c, ka, pa, kb, pb, n, pass * SORT_BITS));
        long long *tk = ka; ka = kb; kb = tk;
        int *tp = pa; pa = pb; pb = tp;
    }
    return GRAV_B200_OK;   // even number of passes: result is back in keys / perm
}

int bh_keys(grav_b200_ctx *c, const double *box_center, double box_width, bool want_unsorted)
{
    DevTree &t = c->tree;
    const int n = c->n;
    t.n = n;
    GB_TRY(t.bbox.reserve(sizeof(double) * 16));
    GB_TRY(t.keys.reserve(sizeof(long long) * (size_t)n));
    GB_TRY(t.perm.reserve(sizeof(int) * (size_t)n));
    if (want_unsorted) GB_TRY(t.keys_unsorted.reserve(sizeof(long long) * (size_t)n));
    long long *mm = t.bbox.as<long long>();
    double *box = t.bbox.as<double>() + 8;
    if (!box_center || box_width <= 0.0) {
        bbox_init_kernel<<<1, 32, 0, c->stream>>>(mm);
        GB_LAUNCH_CHECK();
        int blocks = (n + 255) / 256;
        if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
        bbox_reduce_kernel<<<blocks, 256, 0, c->stream>>>(c->posm.as<double4>(), n, mm);
        GB_LAUNCH_CHECK();
        bbox_finalize_kernel<<<1, 1, 0, c->stream>>>(mm, box);
        GB_LAUNCH_CHECK();
        count_launch(3);
    } else {
        const double h[4] = {box_center[0], box_center[1], box_center[2], box_width};
        GB_CUDA(cudaMemcpyAsync(box, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));   // h is a stack array
    }
    morton_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->posm.as<double4>(), n, box,
                                                          want_unsorted ? t.keys_unsorted.as<long long>() : nullptr,
                                                          t.keys.as<long long>(), t.perm.as<int>());
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

}  // namespace gb
