// ubench_lat.cu — dependent-issue latency (clk per op of ONE warp running a dependent chain) of the instructions the NLMS
// recurrence is made of: FFMA, FFMA2, FADD, FADD2, FMUL2, SHFL.BFLY, MUFU.RCP, LDS.128 — and their single-warp throughput
// with 8 independent chains.  B200: nvcc -arch=sm_100a -O3 -o ubench_lat ubench_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CHAINS>
__global__ void k(float *out, long long *clk, int iters, float x, float y)
{
    __shared__ float4 sm[64];
    float2 a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f + 1.0f);
    sm[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f); sm[threadIdx.x + 32] = sm[threadIdx.x];
    const float2 xx = make_float2(x, x * 1.01f), yy = make_float2(y, y * 0.99f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (OP == 0) a[i].x = fmaf(a[i].x, xx.x, yy.x);
                if (OP == 1) a[i] = __ffma2_rn(a[i], xx, yy);
                if (OP == 2) a[i].x = __fadd_rn(a[i].x, yy.x);
                if (OP == 3) a[i] = __fadd2_rn(a[i], yy);
                if (OP == 4) a[i] = __fmul2_rn(a[i], xx);
                if (OP == 5) a[i].x = __shfl_xor_sync(0xffffffffu, a[i].x, 8);
                if (OP == 6) { float rr; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(a[i].x)); a[i].x = rr; }
                if (OP == 7) { const float4 v = sm[(threadIdx.x + (__float_as_int(a[i].x) & 1)) & 63]; a[i].x = v.x + 1.0f; }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += a[i].x + a[i].y;
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *clk = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char *name, float *d, long long *dc)
{
    const int iters = 4096;
    k<OP, CHAINS><<<1, 32>>>(d, dc, iters, 0.999f, 0.001f);
    k<OP, CHAINS><<<1, 32>>>(d, dc, iters, 0.999f, 0.001f);
    long long c = 0;
    cudaMemcpy(&c, dc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-10s %d chain(s): %6.2f clk per instruction (one warp)\n", name, CHAINS, (double)c / ((double)iters * 8 * CHAINS));
}

int main()
{
    float *d; long long *dc;
    cudaMalloc(&d, 1024); cudaMalloc(&dc, 8);
    run<0, 1>("FFMA", d, dc); run<0, 8>("FFMA", d, dc);
    run<1, 1>("FFMA2", d, dc); run<1, 2>("FFMA2", d, dc); run<1, 4>("FFMA2", d, dc); run<1, 8>("FFMA2", d, dc);
    run<2, 1>("FADD", d, dc); run<3, 1>("FADD2", d, dc); run<3, 8>("FADD2", d, dc);
    run<4, 1>("FMUL2", d, dc); run<4, 8>("FMUL2", d, dc);
    run<5, 1>("SHFL", d, dc); run<5, 8>("SHFL", d, dc);
    run<6, 1>("MUFU.RCP", d, dc); run<6, 8>("MUFU.RCP", d, dc);
    run<7, 1>("LDS.128", d, dc); run<7, 8>("LDS.128", d, dc);
    return 0;
}
