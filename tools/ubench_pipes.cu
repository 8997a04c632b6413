// ubench_pipes.cu — instruction throughput of the integer / FP32 pipes on B200 (sm_100a), to choose the arithmetic
// of the q15 FIR kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNR 16

template <int OP>
__global__ void k(int *out, int a0, int b0)
{
    int acc[UNR];
    float facc[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) { acc[i] = threadIdx.x + i; facc[i] = (float)(threadIdx.x + i); }
    int a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
    float fa = (float)a, fb = (float)b * 1e-3f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) acc[i] = acc[i] + a * b;                                                    // IMAD
            if (OP == 1) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));   // IDP.2A
            if (OP == 2) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));      // IDP.4A
            if (OP == 3) facc[i] = fmaf(fa, fb, facc[i]);                                            // FFMA
            if (OP == 4) acc[i] = __mulhi(acc[i], a) + b;                                            // IMAD.HI
            if (OP == 5) acc[i] = __funnelshift_r(acc[i], a, 16) ^ b;                                // SHF + LOP3
            if (OP == 6) acc[i] = __byte_perm(acc[i], a, 0x5410) + b;                                // PRMT + IADD
            if (OP == 7) acc[i] = max(min(acc[i] + a, 32767), -32768);                               // add + clamp
            if (OP == 8) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a), "r"(b));
        }
        a += 1;
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) r += acc[i] + (int)facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char *name, int *d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<OP><<<grid, block>>>(d, 3, 5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<grid, block>>>(d, 3, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * block * ITERS * UNR;
    printf("%-22s %8.3f ms  %8.2f Tera lane-ops/s  = %6.1f lanes/clk/SM at 1.965 GHz\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    int *d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(int));
    run<0>("IMAD", d);
    run<1>("IDP.2A s32.s32", d);
    run<8>("IDP.2A s32.u32", d);
    run<2>("IDP.4A", d);
    run<3>("FFMA", d);
    run<4>("IMAD.HI + IADD", d);
    run<5>("SHF + LOP3", d);
    run<6>("PRMT + IADD", d);
    run<7>("IADD + 2x IMNMX", d);
    return 0;
}
