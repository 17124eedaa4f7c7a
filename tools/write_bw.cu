// Microbenchmark: write-only HBM bandwidth on B200, plain 16-byte stores vs TMA bulk stores of
// shared-memory strips laid out like K2's output (17 280-byte rows of a 622 080-byte frame).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void fill16(uint4 *out, size_t n)
{
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void copy16(const uint4 *in, uint4 *out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void __launch_bounds__(128, 8) bulk_rows(uint8_t *out, int rows_per_frame)
{
    extern __shared__ __align__(128) uint8_t tile[];
    const int f = blockIdx.y, my = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 17280 / 16; i += 128) reinterpret_cast<uint4 *>(tile)[i] = make_uint4(i, f, my, 7);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        uint8_t *fr = out + (size_t)f * 622080;
        uint8_t *oy = fr + (size_t)my * 11520, *ou = fr + 414720 + (size_t)my * 2880, *ov = ou + 103680;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(oy), "r"((unsigned)__cvta_generic_to_shared(tile)), "r"(11520u) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(ou), "r"((unsigned)__cvta_generic_to_shared(tile + 11520)), "r"(2880u) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(ov), "r"((unsigned)__cvta_generic_to_shared(tile + 14400)), "r"(2880u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
int main()
{
    const int F = 4096;
    const size_t bytes = (size_t)F * 622080;
    uint8_t *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) fill16<<<148 * 16, 256>>>((uint4 *)b, bytes / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("fill, 16-byte stores        %7.3f ms  %7.1f GB/s written\n", ms / 10, bytes / (ms / 10 * 1e-3) / 1e9);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) copy16<<<148 * 16, 256>>>((const uint4 *)a, (uint4 *)b, bytes / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("copy, 16-byte loads/stores  %7.3f ms  %7.1f GB/s read+written\n", ms / 10, 2.0 * bytes / (ms / 10 * 1e-3) / 1e9);
        cudaFuncSetAttribute(bulk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 17280);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) bulk_rows<<<dim3(36, F), 128, 17280>>>(b, 36);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("K2-shaped TMA bulk stores   %7.3f ms  %7.1f GB/s written\n", ms / 10, bytes / (ms / 10 * 1e-3) / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
