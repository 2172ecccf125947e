// tools/ubench2.cu -- which pipe do negations / adds / FP adds land on next to IMAD and SHF? (sm_100a; not part of the product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu && ./ubench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define ILP 8
#define IMAD(a,b,c)  asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define SHF(a)       asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c))
#define NEG(a)       asm volatile("neg.s32 %0, %0;" : "+r"(a))
#define SUB(a,z)     asm volatile("sub.s32 %0, %1, %0;" : "+r"(a) : "r"(z))
#define ADD(a,b)     asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b))
#define XORI(a)      asm volatile("xor.b32 %0, %0, -2;" : "+r"(a))
#define FFMA(a,b,c)  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c))
#define FADD(a,c)    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c))
#define MOV(a,b)     asm volatile("mov.b32 %0, %1;" : "=r"(a) : "r"(b))
template<int MODE> __global__ void __launch_bounds__(256) k(int* out, int b, int c, int z, float fb, float fc) {
    int a[ILP], e[ILP]; float f[ILP];
    #pragma unroll
    for (int i=0;i<ILP;i++){ a[i]=threadIdx.x+i; e[i]=threadIdx.x*3+i; f[i]=1.0f+i; }
    for (int it=0; it<ITERS; it++) {
        #pragma unroll
        for (int i=0;i<ILP;i++) {
            if (MODE==0) { NEG(a[i]); }
            if (MODE==1) { IMAD(a[i],b,c); NEG(e[i]); }
            if (MODE==2) { SHF(a[i]); NEG(e[i]); }
            if (MODE==3) { IMAD(a[i],b,c); SHF(e[i]); NEG(a[i]); }
            if (MODE==4) { SUB(a[i],z); }
            if (MODE==5) { IMAD(a[i],b,c); SUB(e[i],z); }
            if (MODE==6) { SHF(a[i]); SUB(e[i],z); }
            if (MODE==7) { ADD(a[i],b); }
            if (MODE==8) { IMAD(a[i],b,c); ADD(e[i],b); }
            if (MODE==9) { SHF(a[i]); ADD(e[i],b); }
            if (MODE==10){ IMAD(a[i],b,c); SHF(e[i]); FFMA(f[i],fb,fc); }
            if (MODE==11){ IMAD(a[i],b,c); SHF(e[i]); FADD(f[i],fc); }
            if (MODE==12){ IMAD(a[i],b,c); SHF(e[i]); ADD(a[i],b); }
            if (MODE==13){ IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); NEG(a[i]); }   // the seeded stage mix
            if (MODE==14){ IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); XORI(a[i]); }
            if (MODE==15){ IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); FADD(f[i],fc); }
            if (MODE==16){ IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); }
            if (MODE==17){ IMAD(a[i],b,c); IMAD(e[i],c,b); SHF(a[i]); SHF(e[i]); SUB(a[i],z); }
            if (MODE==18){ FADD(f[i],fc); }
            if (MODE==19){ IMAD(a[i],b,c); FADD(f[i],fc); }
        }
    }
    int s=0; float fs=0;
    #pragma unroll
    for (int i=0;i<ILP;i++){ s+=a[i]+e[i]; fs+=f[i]; }
    int r = s ^ __float_as_int(fs); if (r==c) out[0]=r;
}
static const char* names[] = {"NEG","IMAD+NEG","SHF+NEG","IMAD+SHF+NEG","SUB(z-a)","IMAD+SUB","SHF+SUB","ADD","IMAD+ADD","SHF+ADD",
 "IMAD+SHF+FFMA","IMAD+SHF+FADD","IMAD+SHF+ADD","2IMAD+2SHF+NEG","2IMAD+2SHF+XOR","2IMAD+2SHF+FADD","2IMAD+2SHF","2IMAD+2SHF+SUB","FADD","IMAD+FADD"};
static const int ninstr[] = {1,2,2,3,1,2,2,1,2,2,3,3,3,5,5,5,4,5,1,2};
template<int MODE> void run(int* d, int sms, double clk_ghz) {
    dim3 grid(sms*8), block(256);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid,block>>>(d,3,5,0,1.0000001f,0.5f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid,block>>>(d,3,5,0,1.0000001f,0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double winstr = (double)grid.x*(256/32)*(double)ITERS*ILP*ninstr[MODE];
    printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM\n", names[MODE], ms, winstr/sms/(ms*1e-3)/(clk_ghz*1e9));
}
template<int M> struct All { static void go(int*d,int sms,double c){ All<M-1>::go(d,sms,c); run<M>(d,sms,c);} };
template<> struct All<-1> { static void go(int*,int,double){} };
int main(){ cudaDeviceProp p; cudaGetDeviceProperties(&p,0); int clk=0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int* d; cudaMalloc(&d,4); All<19>::go(d,p.multiProcessorCount,clk/1e6); return 0; }
