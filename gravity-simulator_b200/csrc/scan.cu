// Exclusive prefix sum of int32 arrays (helper for stream compaction and the octree numbering).
// Three-phase block scan: per-block exclusive scan + block totals, recursive scan of the totals,
// uniform add.  HBM traffic 4 bytes read + 4 written per element per phase; this is never hot.
#include "internal.cuh"

namespace gb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_kernel(const int *in, int *out, int *block_sums, int n)
{
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    // warp inclusive scan of the per-thread sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    int run = inc - sum + (warp > 0 ? warp_tot[warp - 1] : 0);   // exclusive prefix of this thread
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && block_sums) block_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int *__restrict__ out, const int *__restrict__ block_off, int n)
{
    const int off = block_off[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += off;
}

static int scan_rec(grav_b200_ctx *c, const int *d_in, int *d_out, int n, int *tmp, size_t tmp_ints)
{
    const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb <= 1) {
        scan_block_kernel<<<1, SCAN_THREADS, 0, c->stream>>>(d_in, d_out, nullptr, n);
        GB_LAUNCH_CHECK();
        count_launch();
        return GRAV_B200_OK;
    }
    if ((size_t)nb > tmp_ints) { set_error("scan: scratch too small"); return GRAV_B200_EINVAL; }
    scan_block_kernel<<<nb, SCAN_THREADS, 0, c->stream>>>(d_in, d_out, tmp, n);
    GB_LAUNCH_CHECK();
    count_launch();
    GB_TRY(scan_rec(c, tmp, tmp, nb, tmp + nb, tmp_ints - nb));
    scan_add_kernel<<<nb, SCAN_THREADS, 0, c->stream>>>(d_out, tmp, n);
    GB_LAUNCH_CHECK();
    count_launch();
    return GRAV_B200_OK;
}

// d_in may alias d_out.  n >= 1.
int exclusive_scan_int(grav_b200_ctx *c, const int *d_in, int *d_out, int n, DevBuf &tmp)
{
    if (n <= 0) return GRAV_B200_OK;
    size_t need = 0;
    for (int m = (n + SCAN_TILE - 1) / SCAN_TILE; m > 1; m = (m + SCAN_TILE - 1) / SCAN_TILE) need += m;
    need += 16;
    GB_TRY(tmp.reserve(need * sizeof(int)));
    return scan_rec(c, d_in, d_out, n, tmp.as<int>(), need);
}

}  // namespace gb
