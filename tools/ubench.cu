// tools/ubench.cu -- sm_100a pipe-throughput microbenchmarks used to size the CORDIC
// kernels (see DESIGN.md "Machine model").  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

#define IMAD(a,b,c)  asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define IMADHI(a,b,c) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define SHF(a)       asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define LOP(a,b)     asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a) : "r"(b), "r"(c))
#define IADD(a,b,c)  asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(a) : "r"(b), "r"(c))
#define PRMT(a,b)    asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(a) : "r"(b))
#define FFMA(a,b,c)  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c))
#define FFMAI(a)     asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3F000000;" : "+f"(a))
#define FFMARM(a,c)  asm volatile("fma.rm.f32 %0, %0, 0f3E000000, %1;" : "+f"(a) : "f"(c))
#define FFMA2(a,b,c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c))
#define FFMA2RM(a,b,c) asm volatile("fma.rm.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c))

template<int MODE> __global__ void __launch_bounds__(256) k(int* out, int b, int c, float fb, float fc, unsigned long long lb, unsigned long long lc) {
    int a[ILP]; float f[ILP]; unsigned long long l[ILP];
    #pragma unroll
    for (int i=0;i<ILP;i++){ a[i]=threadIdx.x+i; f[i]=1.0f+i; l[i]=lb+i; }
    for (int it=0; it<ITERS; it++) {
        #pragma unroll
        for (int i=0;i<ILP;i++) {
            if (MODE==0) { IMAD(a[i],b,c); }
            if (MODE==1) { SHF(a[i]); }
            if (MODE==2) { LOP(a[i],b); }
            if (MODE==3) { IADD(a[i],b,c); }
            if (MODE==4) { PRMT(a[i],b); }
            if (MODE==5) { FFMA(f[i],fb,fc); }
            if (MODE==6) { FFMAI(f[i]); }
            if (MODE==7) { FFMA2(l[i],lb,lc); }
            if (MODE==8) { IMADHI(a[i],b,c); }
            if (MODE==9) { IMAD(a[i],b,c); SHF(a[i]); }                 // 1 FMA : 1 ALU
            if (MODE==10){ IMAD(a[i],b,c); LOP(a[i],b); SHF(a[i]); }    // 1 : 2
            if (MODE==11){ IMAD(a[i],b,c); IMAD(a[i],c,b); SHF(a[i]); } // 2 : 1
            if (MODE==12){ FFMA(f[i],fb,fc); SHF(a[i]); }
            if (MODE==13){ FFMA2(l[i],lb,lc); SHF(a[i]); }
            if (MODE==14){ FFMA2(l[i],lb,lc); SHF(a[i]); LOP(a[i],b); }
            if (MODE==15){ FFMARM(f[i],fc); }
            if (MODE==16){ FFMA2RM(l[i],lb,lc); }
            if (MODE==17){ FFMA(f[i],fb,fc); IMAD(a[i],b,c); }          // both on FMA pipe?
            if (MODE==18){ FFMA2(l[i],lb,lc); IMAD(a[i],b,c); }
            if (MODE==19){ FFMA2(l[i],lb,lc); IMAD(a[i],b,c); SHF(a[i]); }
            if (MODE==20){ IMADHI(a[i],b,c); SHF(a[i]); }
            if (MODE==21){ PRMT(a[i],b); IMAD(a[i],b,c); }
            if (MODE==22){ SHF(a[i]); LOP(a[i],b); }                     // 2 ALU
            if (MODE==23){ FFMA2(l[i],lb,lc); FFMA(f[i],fb,fc); }
        }
    }
    int s=0; float fs=0; unsigned long long ls=0;
    #pragma unroll
    for (int i=0;i<ILP;i++){ s+=a[i]; fs+=f[i]; ls+=l[i]; }
    int r = s ^ __float_as_int(fs) ^ (int)ls ^ (int)(ls>>32); if (r==c) out[0]=r;
}

static const char* names[] = {"IMAD","SHF","LOP3","IADD3x2","PRMT","FFMA","FFMA.imm","FFMA2","IMAD.HI",
 "IMAD+SHF","IMAD+LOP3+SHF","2IMAD+SHF","FFMA+SHF","FFMA2+SHF","FFMA2+SHF+LOP3","FFMA.RM","FFMA2.RM",
 "FFMA+IMAD","FFMA2+IMAD","FFMA2+IMAD+SHF","IMAD.HI+SHF","PRMT+IMAD","SHF+LOP3","FFMA2+FFMA"};
static const int ninstr[] = {1,1,1,2,1,1,1,1,1,2,3,3,2,2,3,1,1,2,2,3,2,2,2,2};

template<int MODE> void run(int* d, int sms, double clk_ghz) {
    dim3 grid(sms*8), block(256);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid,block>>>(d,3,5,1.0000001f,0.5f,0x3f8000013f800001ull,0x3f0000003f000000ull);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid,block>>>(d,3,5,1.0000001f,0.5f,0x3f8000013f800001ull,0x3f0000003f000000ull);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double winstr = (double)grid.x*(256/32)*(double)ITERS*ILP*ninstr[MODE];
    double per_sm_per_s = winstr/sms/(ms*1e-3);
    printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM (at %.3f GHz)  %7.2f Gwinst/s/SM\n", names[MODE], ms, per_sm_per_s/(clk_ghz*1e9), clk_ghz, per_sm_per_s/1e9);
    cudaError_t err = cudaGetLastError(); if (err) printf("  err %s\n", cudaGetErrorString(err));
}

template<int M> struct All { static void go(int*d,int sms,double c){ All<M-1>::go(d,sms,c); run<M>(d,sms,c);} };
template<> struct All<-1> { static void go(int*,int,double){} };

int main(){
    cudaDeviceProp p; cudaGetDeviceProperties(&p,0);
    int clk_khz=0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s SMs=%d clock attr %.3f GHz\n", p.name, p.multiProcessorCount, clk_khz/1e6);
    int* d; cudaMalloc(&d,4);
    All<23>::go(d,p.multiProcessorCount,clk_khz/1e6);
    return 0;
}
