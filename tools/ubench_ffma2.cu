// ubench_ffma2.cu — does the packed f32x2 FMA of sm_100 (FFMA2, __ffma2_rn) buy issue slots on B200?
// Per SM and clock: lanes of f32 FMA retired with (a) FFMA, (b) FFMA2, and both mixed with independent integer adds
// (ALU pipe) — a kernel that is issue bound on a mix gains from FFMA2 only if (d) beats (c).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float x, float y, int z)
{
    float2 a[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); q[i] = threadIdx.x + i; }
    const float2 xx = make_float2(x, x * 1.01f), yy = make_float2(y, y * 0.99f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) { a[i].x = fmaf(a[i].x, xx.x, yy.x); a[i].y = fmaf(a[i].y, xx.y, yy.y); }
            else a[i] = __ffma2_rn(a[i], xx, yy);
            if (MODE >= 2) { q[i] = (q[i] + z) ^ it; q[i] = (q[i] + it) ^ z; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *d)
{
    const int iters = 8192, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, iters, 0.999f, 0.001f, 3);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 0.999f, 0.001f, 3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int mhz = 0; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double clk = ms * 1e-3 * mhz * 1e3;
    const double fma_lanes = (double)grid * 256 * iters * 16;
    printf("%-34s %7.3f ms  %6.1f f32 FMA lanes / clk / SM (at %d MHz nominal)\n", name, ms, fma_lanes / clk / 148, mhz / 1000);
}

int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FFMA + 4 ALU per pair", d); run<3>("FFMA2 + 4 ALU per pair", d);
    return 0;
}
