// Direct-sum inner loop: can part of the r^-3 correction leave the FP64 pipe?  (stand-alone experiment)
//   V0  baseline, 16 FP64-pipe instructions per interaction (csrc/direct_sum.cu)
//   V1  q = e(3/2 + 15/8 e) entirely in FP32 through F2F conversions     -> 14 FP64  (q carries ~2^-24 relative error)
//   V2  as V1, conversions done with integer bit moves instead of F2F     -> 14 FP64
//   V3  only the 15/8 e^2 term in FP32 (F2F), 3/2 e stays FP64            -> 15 FP64  (full accuracy)
//   V4  as V3 with integer bit moves                                      -> 15 FP64
// Prints rate and the max relative deviation of the accelerations from V0.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq(double x){ double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

__device__ __forceinline__ float d2f_bits(double e){        // |e| < 2^-10, sign kept, 23 mantissa bits (truncated)
    int hi=__double2hiint(e), lo=__double2loint(e);
    int a=(hi&0x7fffffff)-0x38000000; a=max(a,0);
    unsigned fb=__funnelshift_l((unsigned)lo,(unsigned)a,3);
    fb|=(unsigned)hi&0x80000000u;
    return __uint_as_float(fb);
}
__device__ __forceinline__ float d2f_abs_bits(double e){    // |e| only, 20 mantissa bits
    int a=(__double2hiint(e)&0x7fffffff)-0x38000000; a=max(a,0);
    return __uint_as_float((unsigned)a<<3);
}
__device__ __forceinline__ double f2d_bits(float q){        // normal floats (or 0 -> 2^-126-ish, harmless)
    unsigned fb=__float_as_uint(q);
    unsigned s=fb&0x80000000u, m=fb&0x7fffffffu;
    return __hiloint2double((int)(((m>>3)+0x38000000u)|s),(int)(m<<29));
}

template<int V> __device__ __forceinline__ void inter(double4 pj,double xi,double yi,double zi,double eps2,double&ax,double&ay,double&az){
    double dx=pj.x-xi, dy=pj.y-yi, dz=pj.z-zi;
    double r2=fma(dx,dx,eps2); r2=fma(dy,dy,r2); r2=fma(dz,dz,r2);
    double y=rsq(r2); double t=y*y; double e=fma(-r2,t,1.0); double my=pj.w*y; double y3m=my*t;
    double s;
    if(V==0){ double p=fma(1.875,e,1.5); double q=e*p; s=fma(q,y3m,y3m); }
    if(V==1){ float ef=(float)e; float qf=ef*fmaf(1.875f,ef,1.5f); s=fma((double)qf,y3m,y3m); }
    if(V==2){ float ef=d2f_bits(e); float qf=ef*fmaf(1.875f,ef,1.5f); s=fma(f2d_bits(qf),y3m,y3m); }
    if(V==3){ float ef=(float)e; float q2=1.875f*ef*ef; double q=fma(e,1.5,(double)q2); s=fma(q,y3m,y3m); }
    if(V==4){ float ef=d2f_abs_bits(e); float q2=1.875f*ef*ef; double q=fma(e,1.5,f2d_bits(q2)); s=fma(q,y3m,y3m); }
    ax=fma(s,dx,ax); ay=fma(s,dy,ay); az=fma(s,dz,az);
}
template<int TI,int BLOCK,int MINB,int UNROLL,int V>
__global__ void __launch_bounds__(BLOCK,MINB) k(const double4* __restrict__ src,int n,double eps2,double* __restrict__ acc){
    __shared__ double4 tile[256];
    double xi[TI],yi[TI],zi[TI],ax[TI],ay[TI],az[TI];
    const int ib=blockIdx.x, tid=threadIdx.x;
    #pragma unroll
    for(int t=0;t<TI;t++){ double4 q=src[(ib*TI+t)*BLOCK+tid]; xi[t]=q.x;yi[t]=q.y;zi[t]=q.z;ax[t]=ay[t]=az[t]=0; }
    for(int j0=0;j0<n;j0+=256){
        __syncthreads();
        for(int k2=tid;k2<256;k2+=BLOCK) tile[k2]=src[j0+k2];
        __syncthreads();
        #pragma unroll UNROLL
        for(int j=0;j<256;j++){
            double4 pj=tile[j];
            #pragma unroll
            for(int t=0;t<TI;t++) inter<V>(pj,xi[t],yi[t],zi[t],eps2,ax[t],ay[t],az[t]);
        }
    }
    #pragma unroll
    for(int t=0;t<TI;t++){ int i=(ib*TI+t)*BLOCK+tid; acc[3*i]=ax[t];acc[3*i+1]=ay[t];acc[3*i+2]=az[t]; }
}
static std::vector<double> ref;
template<int TI,int BLOCK,int MINB,int UNROLL,int V> void run(const double4* d,double* acc,int fp64ops){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k<TI,BLOCK,MINB,UNROLL,V>,BLOCK,0);
    int grid=148*occ; const int nsrc=131072; float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); k<TI,BLOCK,MINB,UNROLL,V><<<grid,BLOCK>>>(d,nsrc,1e-4,acc); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<TI,BLOCK,MINB,UNROLL,V>);
    const int ncheck=4096;
    std::vector<double> h(3*ncheck); cudaMemcpy(h.data(),acc,h.size()*8,cudaMemcpyDeviceToHost);
    double worst=0;
    if(V==0&&ref.empty()) ref=h;
    for(int i=0;i<ncheck;i++){ double dn=0,rn=0; for(int c=0;c<3;c++){ double dd=h[3*i+c]-ref[3*i+c]; dn+=dd*dd; rn+=ref[3*i+c]*ref[3*i+c]; } worst=fmax(worst,sqrt(dn/rn)); }
    double rate=(double)grid*TI*BLOCK*(double)nsrc/(best*1e-3);
    printf("V=%d TI=%d MINB=%d UNR=%d regs=%d occ=%d  %.2f ms  %.1f G/s  pipe-instr util(%d)=%.1f%%  max rel dev vs V0 %.2e\n",V,TI,MINB,UNROLL,fa.numRegs,occ,best,rate/1e9,fp64ops,rate*fp64ops/(148*64*1.965e9)*100,worst);
}
int main(){
    const int n=148*256*16;
    std::vector<double4> h(n); srand(1);
    for(auto&p:h){p.x=rand()/(double)RAND_MAX;p.y=rand()/(double)RAND_MAX;p.z=rand()/(double)RAND_MAX;p.w=(0.5+rand()/(double)RAND_MAX)/n;}
    double4* d; double* acc; cudaMalloc(&d,n*sizeof(double4)); cudaMalloc(&acc,3*n*sizeof(double)); cudaMemcpy(d,h.data(),n*sizeof(double4),cudaMemcpyHostToDevice);
    run<4,256,2,2,0>(d,acc,16);
    run<4,256,2,2,1>(d,acc,14);
    run<4,256,2,2,2>(d,acc,14);
    run<4,256,2,2,3>(d,acc,15);
    run<4,256,2,2,4>(d,acc,15);
    run<4,256,2,1,1>(d,acc,14);
    run<4,256,2,1,2>(d,acc,14);
    run<4,256,2,1,3>(d,acc,15);
    run<4,256,2,1,4>(d,acc,15);
    run<4,256,2,4,1>(d,acc,14);
    run<4,256,2,4,3>(d,acc,15);
    run<3,256,2,2,1>(d,acc,14);
    run<3,256,2,2,3>(d,acc,15);
    run<2,256,4,2,1>(d,acc,14);
    run<2,256,4,2,3>(d,acc,15);
    printf("%s\n",cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
