// Tuning harness for the direct-sum inner loop (stand-alone; winner gets ported to csrc/direct_sum.cu)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq(double x){ double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

// SHAPE 0: baseline   1: diag - acc uses 2 distinct regs (wrong numerics)   2: acc via one asm block  3: whole body asm-ordered
template<int SHAPE> __device__ __forceinline__ void inter(double4 pj, double xi,double yi,double zi,double eps2,double&ax,double&ay,double&az){
    double dx=pj.x-xi, dy=pj.y-yi, dz=pj.z-zi;
    double r2=fma(dx,dx,eps2); r2=fma(dy,dy,r2); r2=fma(dz,dz,r2);
    double y=rsq(r2); double t=y*y; double e=fma(-r2,t,1.0); double my=pj.w*y; double y3m=my*t;
    double p=fma(1.875,e,1.5); double q=e*p; double s=fma(q,y3m,y3m);
    if (SHAPE==0){ ax=fma(s,dx,ax); ay=fma(s,dy,ay); az=fma(s,dz,az); }
    if (SHAPE==1){ ax=fma(s,s,ax); ay=fma(dy,dy,ay); az=fma(dz,dz,az); }
    if (SHAPE==2){ asm volatile("fma.rn.f64 %0, %4, %3, %0;\n\tfma.rn.f64 %1, %5, %3, %1;\n\tfma.rn.f64 %2, %6, %3, %2;" : "+d"(ax),"+d"(ay),"+d"(az) : "d"(s),"d"(dx),"d"(dy),"d"(dz)); }
}

template<int TI,int BLOCK,int MINB,int UNROLL,int SHAPE>
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
            for(int t=0;t<TI;t++) inter<SHAPE>(pj,xi[t],yi[t],zi[t],eps2,ax[t],ay[t],az[t]);
        }
    }
    #pragma unroll
    for(int t=0;t<TI;t++){ int i=(ib*TI+t)*BLOCK+tid; acc[3*i]=ax[t];acc[3*i+1]=ay[t];acc[3*i+2]=az[t]; }
}
template<int TI,int BLOCK,int MINB,int UNROLL,int PF>
__global__ void __launch_bounds__(BLOCK,MINB) kg(const double4* __restrict__ src,int n,double eps2,double* __restrict__ acc){
    double xi[TI],yi[TI],zi[TI],ax[TI],ay[TI],az[TI];
    const int ib=blockIdx.x, tid=threadIdx.x;
    #pragma unroll
    for(int t=0;t<TI;t++){ double4 q=src[(ib*TI+t)*BLOCK+tid]; xi[t]=q.x;yi[t]=q.y;zi[t]=q.z;ax[t]=ay[t]=az[t]=0; }
    double4 buf[PF];
    #pragma unroll
    for(int k2=0;k2<PF;k2++) buf[k2]=src[k2];
    for(int j0=0;j0<n;j0+=PF){
        #pragma unroll
        for(int k2=0;k2<PF;k2++){
            double4 pj=buf[k2];
            int jn=j0+PF+k2; if(jn>=n) jn=n-1;
            buf[k2]=src[jn];
            #pragma unroll
            for(int t=0;t<TI;t++) inter<0>(pj,xi[t],yi[t],zi[t],eps2,ax[t],ay[t],az[t]);
        }
    }
    #pragma unroll
    for(int t=0;t<TI;t++){ int i=(ib*TI+t)*BLOCK+tid; acc[3*i]=ax[t];acc[3*i+1]=ay[t];acc[3*i+2]=az[t]; }
}
// stage-major ordering over the TI targets
template<int TI,int BLOCK,int MINB,int UNROLL,int ASM>
__global__ void __launch_bounds__(BLOCK,MINB) ks(const double4* __restrict__ src,int n,double eps2,double* __restrict__ acc){
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
            const double4 pj=tile[j];
            double dx[TI],dy[TI],dz[TI],r2[TI],y[TI],tt[TI],e[TI],my[TI],y3[TI],p[TI],q[TI],s[TI];
            #pragma unroll
            for(int t=0;t<TI;t++){ dx[t]=pj.x-xi[t]; dy[t]=pj.y-yi[t]; dz[t]=pj.z-zi[t]; }
            #pragma unroll
            for(int t=0;t<TI;t++) r2[t]=fma(dx[t],dx[t],eps2);
            #pragma unroll
            for(int t=0;t<TI;t++) r2[t]=fma(dy[t],dy[t],r2[t]);
            #pragma unroll
            for(int t=0;t<TI;t++) r2[t]=fma(dz[t],dz[t],r2[t]);
            #pragma unroll
            for(int t=0;t<TI;t++) y[t]=rsq(r2[t]);
            #pragma unroll
            for(int t=0;t<TI;t++){ tt[t]=y[t]*y[t]; my[t]=pj.w*y[t]; }
            #pragma unroll
            for(int t=0;t<TI;t++){ e[t]=fma(-r2[t],tt[t],1.0); y3[t]=my[t]*tt[t]; }
            #pragma unroll
            for(int t=0;t<TI;t++) p[t]=fma(1.875,e[t],1.5);
            #pragma unroll
            for(int t=0;t<TI;t++) q[t]=e[t]*p[t];
            #pragma unroll
            for(int t=0;t<TI;t++) s[t]=fma(q[t],y3[t],y3[t]);
            #pragma unroll
            for(int t=0;t<TI;t++){
                if(ASM){ asm volatile("fma.rn.f64 %0, %4, %3, %0;\n\tfma.rn.f64 %1, %5, %3, %1;\n\tfma.rn.f64 %2, %6, %3, %2;" : "+d"(ax[t]),"+d"(ay[t]),"+d"(az[t]) : "d"(s[t]),"d"(dx[t]),"d"(dy[t]),"d"(dz[t])); }
                else { ax[t]=fma(dx[t],s[t],ax[t]); ay[t]=fma(dy[t],s[t],ay[t]); az[t]=fma(dz[t],s[t],az[t]); }
            }
        }
    }
    #pragma unroll
    for(int t=0;t<TI;t++){ int i=(ib*TI+t)*BLOCK+tid; acc[3*i]=ax[t];acc[3*i+1]=ay[t];acc[3*i+2]=az[t]; }
}
template<int TI,int BLOCK,int MINB,int UNROLL,int ASM> void runs(const double4* d,int n,double* acc){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0,ks<TI,BLOCK,MINB,UNROLL,ASM>,BLOCK,0);
    int grid=148*occ0; const int nsrc=131072; float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); ks<TI,BLOCK,MINB,UNROLL,ASM><<<grid,BLOCK>>>(d,nsrc,1e-4,acc); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,ks<TI,BLOCK,MINB,UNROLL,ASM>);
    double rate=(double)grid*TI*BLOCK*(double)nsrc/(best*1e-3);
    printf("STAGED TI=%d BLOCK=%d MINB=%d UNR=%d ASM=%d regs=%d occ=%d grid=%d  %.2f ms  %.1f G/s  util16=%.1f%%\n",TI,BLOCK,MINB,UNROLL,ASM,fa.numRegs,occ0,grid,best,rate/1e9,rate*16/(148*64*1.965e9)*100);
}
// double-buffered shared tiles, one barrier per tile
template<int TI,int BLOCK,int MINB,int UNROLL,int TILE>
__global__ void __launch_bounds__(BLOCK,MINB) kd(const double4* __restrict__ src,int n,double eps2,double* __restrict__ acc){
    __shared__ double4 tile[2][TILE];
    double xi[TI],yi[TI],zi[TI],ax[TI],ay[TI],az[TI];
    const int ib=blockIdx.x, tid=threadIdx.x;
    #pragma unroll
    for(int t=0;t<TI;t++){ double4 q=src[(ib*TI+t)*BLOCK+tid]; xi[t]=q.x;yi[t]=q.y;zi[t]=q.z;ax[t]=ay[t]=az[t]=0; }
    constexpr int PER=TILE/BLOCK;
    double4 nx[PER];
    #pragma unroll
    for(int k2=0;k2<PER;k2++) nx[k2]=src[k2*BLOCK+tid];
    #pragma unroll
    for(int k2=0;k2<PER;k2++) tile[0][k2*BLOCK+tid]=nx[k2];
    __syncthreads();
    int cur=0;
    for(int j0=0;j0<n;j0+=TILE){
        const bool more = j0+TILE<n;
        if(more){
            #pragma unroll
            for(int k2=0;k2<PER;k2++) nx[k2]=src[j0+TILE+k2*BLOCK+tid];
        }
        #pragma unroll UNROLL
        for(int j=0;j<TILE;j++){
            double4 pj=tile[cur][j];
            #pragma unroll
            for(int t=0;t<TI;t++) inter<0>(pj,xi[t],yi[t],zi[t],eps2,ax[t],ay[t],az[t]);
        }
        if(more){
            #pragma unroll
            for(int k2=0;k2<PER;k2++) tile[cur^1][k2*BLOCK+tid]=nx[k2];
        }
        __syncthreads();
        cur^=1;
    }
    #pragma unroll
    for(int t=0;t<TI;t++){ int i=(ib*TI+t)*BLOCK+tid; acc[3*i]=ax[t];acc[3*i+1]=ay[t];acc[3*i+2]=az[t]; }
}
template<int TI,int BLOCK,int MINB,int UNROLL,int TILE> void rund(const double4* d,int n,double* acc){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0,kd<TI,BLOCK,MINB,UNROLL,TILE>,BLOCK,0);
    int grid=148*occ0; const int nsrc=131072; float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); kd<TI,BLOCK,MINB,UNROLL,TILE><<<grid,BLOCK>>>(d,nsrc,1e-4,acc); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,kd<TI,BLOCK,MINB,UNROLL,TILE>);
    double rate=(double)grid*TI*BLOCK*(double)nsrc/(best*1e-3);
    printf("DB TI=%d BLOCK=%d MINB=%d UNR=%d TILE=%d regs=%d occ=%d grid=%d  %.2f ms  %.1f G/s  util16=%.1f%%\n",TI,BLOCK,MINB,UNROLL,TILE,fa.numRegs,occ0,grid,best,rate/1e9,rate*16/(148*64*1.965e9)*100);
}
template<int TI,int BLOCK,int MINB,int UNROLL,int PF> void rung(const double4* d,int n,double* acc){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0,kg<TI,BLOCK,MINB,UNROLL,PF>,BLOCK,0);
    int grid=148*occ0; const int nsrc=131072; float best=1e9;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); kg<TI,BLOCK,MINB,UNROLL,PF><<<grid,BLOCK>>>(d,nsrc,1e-4,acc); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,kg<TI,BLOCK,MINB,UNROLL,PF>);
    double rate=(double)grid*TI*BLOCK*(double)nsrc/(best*1e-3);
    printf("LDG TI=%d BLOCK=%d MINB=%d PF=%d regs=%d occ=%d grid=%d  %.2f ms  %.1f G/s  util16=%.1f%%\n",TI,BLOCK,MINB,PF,fa.numRegs,occ0,grid,best,rate/1e9,rate*16/(148*64*1.965e9)*100);
}
template<int TI,int BLOCK,int MINB,int UNROLL,int SHAPE> void run(const double4* d,int n,double* acc){
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0,k<TI,BLOCK,MINB,UNROLL,SHAPE>,BLOCK,0);
    int grid=148*occ0; const int nsrc=131072; float best=1e9; n=nsrc;
    for(int r=0;r<3;r++){ cudaEventRecord(e0); k<TI,BLOCK,MINB,UNROLL,SHAPE><<<grid,BLOCK>>>(d,n,1e-4,acc); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<TI,BLOCK,MINB,UNROLL,SHAPE>);
    int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k<TI,BLOCK,MINB,UNROLL,SHAPE>,BLOCK,0);
    double rate=(double)grid*TI*BLOCK*(double)nsrc/(best*1e-3);
    printf("TI=%d BLOCK=%d MINB=%d UNR=%d SHAPE=%d regs=%d occ=%d grid=%d  %.2f ms  %.1f G/s  util16=%.1f%%\n",TI,BLOCK,MINB,UNROLL,SHAPE,fa.numRegs,occ,grid,best,rate/1e9,rate*16/(148*64*1.965e9)*100);
}
int main(){
    const int n=148*256*16;   // divisible by 256*{1,2,3,4,5,6} and 128*8, 512*4... (568320)   // 303104: exactly one wave for TI=2,BLOCK=256,4/SM
    std::vector<double4> h(n); srand(1);
    for(auto&p:h){p.x=rand()/(double)RAND_MAX;p.y=rand()/(double)RAND_MAX;p.z=rand()/(double)RAND_MAX;p.w=1.0/n;}
    double4* d; double* acc; cudaMalloc(&d,n*sizeof(double4)); cudaMalloc(&acc,3*n*sizeof(double)); cudaMemcpy(d,h.data(),n*sizeof(double4),cudaMemcpyHostToDevice);
    run<4,256,2,2,0>(d,n,acc);
    runs<4,256,2,1,0>(d,n,acc);
    runs<4,256,2,1,1>(d,n,acc);
    runs<4,256,2,2,0>(d,n,acc);
    runs<4,256,2,2,1>(d,n,acc);
    runs<4,256,2,4,0>(d,n,acc);
    runs<3,256,2,1,0>(d,n,acc);
    runs<3,256,2,2,0>(d,n,acc);
    runs<2,256,4,2,0>(d,n,acc);
    runs<2,256,4,4,0>(d,n,acc);
    runs<5,256,2,1,0>(d,n,acc);
    runs<6,256,1,1,0>(d,n,acc);
    runs<4,128,4,2,0>(d,n,acc);
    runs<4,512,1,2,0>(d,n,acc);
    return 0;
}
