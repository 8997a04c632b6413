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

// the FIR register tile: 8 accumulators x 16 taps from 24 samples, everything in registers (no memory in the loop)
template <int MODE>
__global__ void k_tile(int *out, const int *in)
{
    int s[24], tp[16];
    unsigned acc[8];
    float fs[24], ftp[16], facc[8];
#pragma unroll
    for (int i = 0; i < 24; i++) { s[i] = in[threadIdx.x + i]; fs[i] = (float)s[i]; }
#pragma unroll
    for (int i = 0; i < 16; i++) { tp[i] = in[64 + i]; ftp[i] = (float)tp[i]; }
#pragma unroll
    for (int j = 0; j < 8; j++) { acc[j] = 0; facc[j] = 0.f; }
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (MODE == 0) acc[j] += (unsigned)(tp[kk] * s[16 + j - kk]);
                else facc[j] = fmaf(ftp[kk], fs[16 + j - kk], facc[j]);
            }
        s[it & 7] += 1; fs[it & 7] += 1.f;     // keep the loop from being hoisted
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r += acc[j] + (unsigned)facc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (int)r;
}

template <int MODE>
void run_tile(const char *name, int *d, int *din)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 4, block = 256;
    k_tile<MODE><<<grid, block>>>(d, din);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_tile<MODE><<<grid, block>>>(d, din);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * block * (ITERS / 8) * 128;
    printf("%-22s %8.3f ms  %8.2f Tera lane-ops/s  = %6.1f lanes/clk/SM at 1.965 GHz\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
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
    int *din;
    cudaMalloc(&din, 4096 * sizeof(int));
    cudaMemset(din, 1, 4096 * sizeof(int));
    run_tile<0>("FIR tile IMAD 8x16", d, din);
    run_tile<1>("FIR tile FFMA 8x16", d, din);
    return 0;
}
