// Microbenchmark 2: issue rate of individual integer SASS ops on sm_100a (inline PTX keeps the op).
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(S) S(a0,a1) S(a1,a2) S(a2,a3) S(a3,a4) S(a4,a5) S(a5,a6) S(a6,a7) S(a7,a0)
#define LOP(x,y)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(m));
#define SHF(x,y)  asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(m));
#define PRM(x,y)  asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(m));
#define ADD(x,y)  asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
#define MAD(x,y)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
#define MNX(x,y)  asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
#define SHL(x,y)  asm volatile("shl.b32 %0, %0, 1;" : "+r"(x));
#define BFE(x,y)  asm volatile("bfe.u32 %0, %0, 3, 9;" : "+r"(x));
#define DP4(x,y)  asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(y), "r"(m));
#define SETSEL(x,y) asm volatile("{ .reg .pred p; setp.le.s32 p, %0, 0; selp.u32 %0, %1, %0, p; }" : "+r"(x) : "r"(y));
#define POPC(x,y) asm volatile("popc.b32 %0, %0;" : "+r"(x));
#define VMX(x,y)  asm volatile("vmax2.s32.s32.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(m));
template <int MODE>
__global__ void k(unsigned *out, int iters, unsigned seed)
{
    unsigned a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    unsigned a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    unsigned m = seed | 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) { REP8(LOP) }
            if (MODE == 1) { REP8(SHF) }
            if (MODE == 2) { REP8(PRM) }
            if (MODE == 3) { REP8(ADD) }
            if (MODE == 4) { REP8(MAD) }
            if (MODE == 5) { REP8(MNX) }
            if (MODE == 6) { LOP(a0,a1) MAD(a1,a2) LOP(a2,a3) MAD(a3,a4) LOP(a4,a5) MAD(a5,a6) LOP(a6,a7) MAD(a7,a0) }
            if (MODE == 7) { LOP(a0,a1) ADD(a1,a2) SHF(a2,a3) PRM(a3,a4) LOP(a4,a5) ADD(a5,a6) SHF(a6,a7) PRM(a7,a0) }
            if (MODE == 8) { REP8(SHL) }
            if (MODE == 10) { REP8(DP4) }
            if (MODE == 11) { LOP(a0,a1) DP4(a1,a2) LOP(a2,a3) DP4(a3,a4) LOP(a4,a5) DP4(a5,a6) LOP(a6,a7) DP4(a7,a0) }
            if (MODE == 12) { MAD(a0,a1) DP4(a1,a2) MAD(a2,a3) DP4(a3,a4) MAD(a4,a5) DP4(a5,a6) MAD(a6,a7) DP4(a7,a0) }
            if (MODE == 13) { REP8(SETSEL) }
            if (MODE == 14) { REP8(POPC) }
            if (MODE == 9) { LOP(a0,a1) LOP(a1,a2) MAD(a2,a3) LOP(a3,a4) LOP(a4,a5) MAD(a5,a6) LOP(a6,a7) LOP(a7,a0) }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
template <int MODE> void run(const char *name)
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(d, 100, 1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)148 * 8 * 8 * iters * 64;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %.3f warp-inst/clk/SMSP\n", name, ms, winst / (ms * 1e-3) / (clk * 1e3) / (148 * 4));
    cudaFree(d);
}
int main()
{
    run<0>("LOP3"); run<1>("SHF"); run<2>("PRMT"); run<3>("add.u32"); run<4>("mad.lo (IMAD)"); run<5>("max.s32");
    run<6>("LOP3:IMAD 1:1"); run<7>("LOP3/ADD/SHF/PRMT"); run<8>("shl imm"); run<9>("LOP3:IMAD 3:1");
    run<10>("dp4a (IDP.4A)"); run<11>("LOP3:dp4a 1:1"); run<12>("IMAD:dp4a 1:1"); run<13>("setp+selp pair (counted as 1)"); run<14>("popc");
    return 0;
}
