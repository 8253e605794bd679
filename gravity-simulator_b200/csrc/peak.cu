// Measures the FP64 FMA peak of the device with a register-resident DFMA loop: the roofline
// denominator for the direct sum (MEASURED_PEAKS.json has HBM and bf16 only).
#include "internal.cuh"

namespace gb {

constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, double a, double b, long long *clk)
{
    double v[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) v[k] = (double)(threadIdx.x + k);
    const long long t0 = clock64();
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int k = 0; k < PEAK_CHAINS; k++) v[k] = fma(v[k], a, b);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) s += v[k];
    if (s == 12345.678) out[0] = s;   // keep the chains alive
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}

}  // namespace gb

extern "C" int grav_b200_measure_fp64_peak(int device, double *tflops, double *sm_mhz)
{
    using namespace gb;
    if (!tflops) { set_error("NULL out pointer"); return GRAV_B200_EINVAL; }
    GB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    double *d_out;
    long long *d_clk;
    GB_CUDA(cudaMalloc(&d_out, 64));
    GB_CUDA(cudaMalloc(&d_clk, 64));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 8;   // 2048 threads per SM
    float best = 1e30f;
    long long cyc = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<grid, 256>>>(d_out, 0.999999, 1e-9, d_clk);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) return cuda_fail(e, "dfma_peak_kernel", __FILE__, __LINE__);
        count_launch();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) {
            best = ms;
            cudaMemcpy(&cyc, d_clk, sizeof(cyc), cudaMemcpyDeviceToHost);
        }
    }
    const double fmas = (double)grid * 256.0 * PEAK_CHAINS * 4.0 * PEAK_ITERS;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    // equivalent SM clock if every SM retires 64 DFMA per cycle (clock64() does not tick at the SM clock here)
    (void)cyc;
    if (sm_mhz) *sm_mhz = *tflops * 1e12 / (2.0 * 64.0 * prop.multiProcessorCount) / 1e6;
    cudaFree(d_out);
    cudaFree(d_clk);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return GRAV_B200_OK;
}
