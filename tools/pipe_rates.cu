// Microbenchmark: issue rate of ALU-pipe vs FMA-pipe integer instructions on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned *out, int iters, unsigned seed)
{
    unsigned a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
    unsigned a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    unsigned m = seed | 3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {        // ALU: LOP3 / IADD3
                a0 = (a0 ^ m) + a1; a1 = (a1 & m) ^ a2; a2 = (a2 | m) + a3; a3 = (a3 ^ a0) + m;
                a4 = (a4 ^ m) + a5; a5 = (a5 & m) ^ a6; a6 = (a6 | m) + a7; a7 = (a7 ^ a4) + m;
            } else if (MODE == 1) { // FMA pipe: IMAD
                a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a0;
                a4 = a4 * m + a5; a5 = a5 * m + a6; a6 = a6 * m + a7; a7 = a7 * m + a4;
            } else if (MODE == 2) { // mixed 1:1
                a0 = a0 * m + a1; a1 = (a1 ^ m) + a2; a2 = a2 * m + a3; a3 = (a3 ^ a0) + m;
                a4 = a4 * m + a5; a5 = (a5 ^ m) + a6; a6 = a6 * m + a7; a7 = (a7 ^ a4) + m;
            } else if (MODE == 3) { // PRMT + SHF
                a0 = __byte_perm(a0, a1, m); a1 = __funnelshift_r(a1, a2, m); a2 = __byte_perm(a2, a3, m); a3 = __funnelshift_r(a3, a0, m);
                a4 = __byte_perm(a4, a5, m); a5 = __funnelshift_r(a5, a6, m); a6 = __byte_perm(a6, a7, m); a7 = __funnelshift_r(a7, a4, m);
            } else if (MODE == 4) { // mixed 3 ALU : 1 IMAD
                a0 = a0 * m + a1; a1 = (a1 ^ m) + a2; a2 = (a2 | m) ^ a3; a3 = (a3 ^ a0) + m;
                a4 = a4 * m + a5; a5 = (a5 ^ m) + a6; a6 = (a6 | m) ^ a7; a7 = (a7 ^ a4) + m;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
template <int MODE> void run(const char *name, int per_iter)
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(d, 100, 1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)148 * 8 * 8 * iters * 16 * per_iter;   // warp-instructions (nominal)
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %.3f nominal warp-inst/clk/SMSP (at %d MHz)\n", name, ms, winst / (ms * 1e-3) / (clk * 1e3) / (148 * 4), clk / 1000);
    cudaFree(d);
}
int main()
{
    run<0>("ALU (LOP3+IADD3) 16/iter", 16);
    run<1>("IMAD 8/iter", 8);
    run<2>("mixed 4 IMAD + 8 ALU", 12);
    run<3>("PRMT+SHF 8/iter", 8);
    run<4>("mixed 2 IMAD + 12 ALU", 14);
    return 0;
}
