// FP64 pipe microbenchmarks: DFMA throughput vs number of distinct register operands.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 8
template<int MODE> __global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
    double v[CH], u[CH], w[CH];
    for (int i=0;i<CH;i++){ v[i]=threadIdx.x+i; u[i]=1.0+1e-9*(threadIdx.x+i); w[i]=1e-9*(threadIdx.x*3+i+1)+a; }
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int r=0;r<4;r++)
        #pragma unroll
        for (int i=0;i<CH;i++) {
            if (MODE==1) v[i]=fma(v[i],a,b);
            if (MODE==2) v[i]=fma(v[i],w[i],b);
            if (MODE==3) v[i]=fma(u[i],w[i],v[i]);
            if (MODE==4) v[i]=fma(u[i],w[(i+1)%CH],v[i]);   // different pairing
            if (MODE==5) v[i]=v[i]*w[i];                    // DMUL 2 regs
            if (MODE==6) v[i]=v[i]+w[i];                    // DADD 2 regs
            if (MODE==7) v[i]=fma(u[i],u[i],v[i]);          // 2 distinct (square)
        }
        if (MODE==8) {   // accumulate pattern: groups of 3 FMAs sharing one operand (s), 8 accumulators + ...
            #pragma unroll
            for (int r=0;r<4;r++) {
                #pragma unroll
                for (int g=0; g<2; g++) {
                    double s = w[g*3+r%2];
                    v[g*3+0]=fma(u[g*3+0],s,v[g*3+0]);
                    v[g*3+1]=fma(u[g*3+1],s,v[g*3+1]);
                    v[g*3+2]=fma(u[g*3+2],s,v[g*3+2]);
                }
                v[6]=fma(v[6],a,b); v[7]=fma(v[7],a,b);
            }
        }
    }
    double s=0; for (int i=0;i<CH;i++) s+=v[i]+u[i]+w[i];
    if (s==1.2345) out[0]=s;
}
template<int MODE> void run(const char* name, int blocks_per_sm) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid=148*blocks_per_sm; float best=1e9;
    for (int rep=0;rep<4;rep++){ cudaEventRecord(e0); k<MODE><<<grid,256>>>(d,0.999999,1e-9,ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep&&ms<best)best=ms; }
    double ops=(double)grid*256*CH*4*ITERS;
    printf("%-28s blocks/SM=%d  %.2f T ops/s  (%.1f%% of 18.6)\n", name, blocks_per_sm, ops/(best*1e-3)/1e12, 100*ops/(best*1e-3)/1e12/(148*64*1.965e-3));
    cudaFree(d);
}
int main(){
    for (int b : {2,4,8}) {
        run<1>("dfma 1 reg (v,a,b const)", b);
        run<2>("dfma 2 regs", b);
        run<3>("dfma 3 regs", b);
        run<4>("dfma 3 regs alt pairing", b);
        run<5>("dmul 2 regs", b);
        run<6>("dadd 2 regs", b);
        run<7>("dfma u*u+v", b);
        run<8>("6x(3reg shared s)+2x1reg", b);
    }
    return 0;
}
