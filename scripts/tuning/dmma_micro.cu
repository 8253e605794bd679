// Does the FP64 tensor instruction (DMMA, mma.sync m8n8k4 f64) run beside the FP64 vector pipe on sm_100a, or on it?
// Measures: DFMA alone, DMMA alone, and DFMA + DMMA interleaved in the same warps (K DFMA per DMMA per thread).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_micro dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// NF DFMAs (two distinct registers, the full-rate form) and NM DMMAs per loop trip, all independent chains
template<int NF, int NM, int SHAPE> __global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
    double v[8], c[8][4], fa[4], fb[2];
    for (int i = 0; i < 8; i++) { v[i] = threadIdx.x + i; for (int j = 0; j < 4; j++) c[i][j] = i + j; }
    for (int j = 0; j < 4; j++) fa[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    fb[0] = 1e-9 * threadIdx.x; fb[1] = a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NF; i++) v[i % 8] = fma(v[i % 8], a, b);
#pragma unroll
        for (int i = 0; i < NM; i++) {
            if (SHAPE == 0) dmma884(c[i % 8][0], c[i % 8][1], fa[0], fb[0]);
            else dmma1688(c[i % 8], fa, fb);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) { s += v[i]; for (int j = 0; j < 4; j++) s += c[i][j]; }
    if (s == 1.2345) out[0] = s;
}

template<int NF, int NM, int SHAPE> void run(const char* name, int bps) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * bps; float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<NF, NM, SHAPE><<<grid, 256>>>(d, 0.999999, 1e-9, ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    double thr = (double)grid * 256 * ITERS;
    double fma_flop = thr * NF * 2;
    double mma_flop = (double)grid * 8 * ITERS * NM * (SHAPE == 0 ? 8 * 8 * 4 * 2 : 16 * 8 * 8 * 2);   // per warp
    printf("%-34s CTAs/SM=%d  %.3f ms   DFMA %.2f TF/s   DMMA %.2f TF/s   sum %.2f\n", name, bps, best, fma_flop / best / 1e9, mma_flop / best / 1e9,
           (fma_flop + mma_flop) / best / 1e9);
    cudaFree(d);
}

int main() {
    for (int b : {2, 4}) {
        run<16, 0, 0>("16 DFMA", b);
        run<0, 8, 0>("8 DMMA m8n8k4", b);
        run<0, 8, 1>("8 DMMA m16n8k8", b);
        run<16, 1, 0>("16 DFMA + 1 DMMA m8n8k4", b);
        run<16, 2, 0>("16 DFMA + 2 DMMA m8n8k4", b);
        run<16, 4, 0>("16 DFMA + 4 DMMA m8n8k4", b);
        run<13, 1, 0>("13 DFMA + 1 DMMA m8n8k4", b);
        run<16, 1, 1>("16 DFMA + 1 DMMA m16n8k8", b);
        run<32, 1, 1>("32 DFMA + 1 DMMA m16n8k8", b);
    }
    return 0;
}
