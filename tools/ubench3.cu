// tools/ubench3.cu -- IDP.2A (dp2a) next to IMAD and SHF: rate and pipe on sm_100a (not part of the product)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define ILP 8
#define IMAD(a,b,c)  asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define SHF(a)       asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define IDPLO(a,w,c) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(w))
#define IDPHI(a,w,c) asm volatile("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(w))
#define IDP4(a,w,c)  asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(c) : "r"(a), "r"(w))
template<int MODE> __global__ void __launch_bounds__(256) k(int* out, int b, int c, int w) {
    int a[ILP], e[ILP];
    #pragma unroll
    for (int i=0;i<ILP;i++){ a[i]=threadIdx.x+i; e[i]=threadIdx.x*3+i; }
    for (int it=0; it<ITERS; it++) {
        #pragma unroll
        for (int i=0;i<ILP;i++) {
            if (MODE==0) { IDPLO(b,w,a[i]); }
            if (MODE==1) { IDPLO(e[i],w,a[i]); SHF(e[i]); }
            if (MODE==2) { IDPLO(e[i],w,a[i]); IDPHI(a[i],w,e[i]); SHF(a[i]); SHF(e[i]); }
            if (MODE==3) { IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); }
            if (MODE==4) { IDP4(b,w,a[i]); }
            if (MODE==5) { IDPLO(e[i],w,a[i]); IMAD(e[i],b,c); }
            if (MODE==6) { IDPLO(e[i],w,a[i]); IDPHI(a[i],w,e[i]); SHF(a[i]); SHF(e[i]); IMAD(a[i],b,c);}
        }
    }
    int s=0;
    #pragma unroll
    for (int i=0;i<ILP;i++){ s+=a[i]+e[i]; }
    if (s==c) out[0]=s;
}
static const char* names[] = {"IDP.2A","IDP.2A+SHF","2IDP.2A+2SHF","2IMAD+2SHF","IDP.4A","IDP.2A+IMAD","2IDP+2SHF+IMAD"};
static const int ninstr[] = {1,2,4,4,1,2,5};
template<int MODE> void run(int* d, int sms, double clk_ghz) {
    dim3 grid(sms*8), block(256);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid,block>>>(d,3,5,0x000100FF); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid,block>>>(d,3,5,0x000100FF); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double winstr = (double)grid.x*(256/32)*(double)ITERS*ILP*ninstr[MODE];
    printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM\n", names[MODE], ms, winstr/sms/(ms*1e-3)/(clk_ghz*1e9));
}
template<int M> struct All { static void go(int*d,int sms,double c){ All<M-1>::go(d,sms,c); run<M>(d,sms,c);} };
template<> struct All<-1> { static void go(int*,int,double){} };
int main(){ cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int clk=0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int* d; cudaMalloc(&d,4); All<6>::go(d,p.multiProcessorCount,clk/1e6); return 0; }
