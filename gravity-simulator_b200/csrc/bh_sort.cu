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
// Unit of work = one warp and a contiguous chunk of SORT_CHUNK pairs, processed 32 at a time in order.
//   pass 1 (histogram): per chunk, count the 256 digit values          -> hist[digit][chunk]
//   scan              : exclusive prefix sum over hist in (digit, chunk) order = first output slot of
//                       every (digit, chunk) group
//   pass 2 (scatter)  : per chunk, walk the pairs again in order; rank inside a group of 32 comes from
//                       match.any + popc of lower lanes, the running count per digit lives in shared memory.
// Within a digit value output order = chunk order, then position order: stable.
constexpr int SORT_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_BITS;
constexpr int SORT_CHUNK = 2048;        // pairs per warp
constexpr int SORT_WARPS = 8;           // warps per CTA

__global__ void __launch_bounds__(SORT_WARPS * 32) sort_hist_kernel(const long long *__restrict__ keys, int n, int shift,
                                                                    int num_chunks, int *__restrict__ hist)
{
    __shared__ int cnt[SORT_WARPS][SORT_RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * SORT_WARPS + warp;
    for (int d = lane; d < SORT_RADIX; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (chunk < num_chunks) {
        const int base = chunk * SORT_CHUNK;
        const int end = min(base + SORT_CHUNK, n);
        for (int p0 = base; p0 < end; p0 += 32) {
            const int p = p0 + lane;
            const bool valid = p < end;
            const int d = valid ? ((int)((unsigned long long)keys[p] >> shift) & (SORT_RADIX - 1)) : SORT_RADIX;
            const unsigned same = __match_any_sync(0xffffffffu, d);
            if (valid && (same >> lane) == 1u) cnt[warp][d] += __popc(same);   // one lane per distinct digit
            __syncwarp();
        }
        for (int d = lane; d < SORT_RADIX; d += 32) hist[(size_t)d * num_chunks + chunk] = cnt[warp][d];
    }
}

__global__ void __launch_bounds__(SORT_WARPS * 32) sort_scatter_kernel(const long long *__restrict__ keys_in,
                                                                       const int *__restrict__ perm_in, int n, int shift,
                                                                       int num_chunks, const int *__restrict__ offs,
                                                                       long long *__restrict__ keys_out,
                                                                       int *__restrict__ perm_out)
{
    __shared__ int pos[SORT_WARPS][SORT_RADIX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * SORT_WARPS + warp;
    if (chunk >= num_chunks) return;
    for (int d = lane; d < SORT_RADIX; d += 32) pos[warp][d] = offs[(size_t)d * num_chunks + chunk];
    __syncwarp();
    const int base = chunk * SORT_CHUNK;
    const int end = min(base + SORT_CHUNK, n);
    const unsigned lt = (1u << lane) - 1u;
    for (int p0 = base; p0 < end; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < end;
        long long k = 0;
        int v = 0;
        if (valid) { k = keys_in[p]; v = perm_in[p]; }
        const int d = valid ? ((int)((unsigned long long)k >> shift) & (SORT_RADIX - 1)) : SORT_RADIX;   // invalid lanes never match
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        const unsigned same = __match_any_sync(0xffffffffu, d) & act;
        int dst = 0;
        if (valid) dst = pos[warp][d] + __popc(same & lt);
        __syncwarp();
        if (valid && (same >> lane) == 1u) pos[warp][d] += __popc(same);   // highest lane of each group advances the counter
        __syncwarp();
        if (valid) { keys_out[dst] = k; perm_out[dst] = v; }
    }
}

int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp);

// keys/perm sorted in place (8 passes ping-pong through keys_tmp/perm_tmp)
int radix_sort_pairs(grav_b200_ctx *c)
{
    DevTree &t = c->tree;
    const int n = t.n;
    const int num_chunks = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    GB_TRY(t.keys_tmp.reserve(sizeof(long long) * (size_t)n));
    GB_TRY(t.perm_tmp.reserve(sizeof(int) * (size_t)n));
    GB_TRY(t.hist.reserve(sizeof(int) * (size_t)SORT_RADIX * num_chunks));
    long long *ka = t.keys.as<long long>(), *kb = t.keys_tmp.as<long long>();
    int *pa = t.perm.as<int>(), *pb = t.perm_tmp.as<int>();
    const int blocks = (num_chunks + SORT_WARPS - 1) / SORT_WARPS;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = pass * SORT_BITS;
        sort_hist_kernel<<<blocks, SORT_WARPS * 32, 0, c->stream>>>(ka, n, shift, num_chunks, t.hist.as<int>());
        GB_LAUNCH_CHECK();
        count_launch();
        GB_TRY(exclusive_scan_int(c, t.hist.as<int>(), t.hist.as<int>(), SORT_RADIX * num_chunks, t.scan_tmp));
        sort_scatter_kernel<<<blocks, SORT_WARPS * 32, 0, c->stream>>>(ka, pa, n, shift, num_chunks, t.hist.as<int>(), kb, pb);
        GB_LAUNCH_CHECK();
        count_launch();
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
