// ubench_imma.cu — throughput of the legacy int8 tensor-core path (mma.sync m16n8k32) on B200, to decide whether a
// Toeplitz-GEMM form of the q15 FIRs can beat the IMAD kernel.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k(int *out, int seed)
{
    int a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 ^ 0x55, b1 = a0 ^ 0x33;
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        a0 += 1;
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main()
{
    int *d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(int));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k<<<148 * 8, 256>>>(d, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)148 * 8 * 8 * ITERS * 8;            // warps * iters * 8
    printf("mma.sync m16n8k32 s8: %.3f ms, %.1f G mma/s, %.1f int8 TOPS, %.2f mma/clk/SM\n", ms, mmas / ms / 1e6,
           mmas * 16 * 8 * 32 * 2 / ms / 1e9, mmas / (ms * 1e-3) / 148 / 1.965e9);
    return 0;
}
