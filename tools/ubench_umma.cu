// ubench_umma.cu — cycles per tcgen05.mma kind::i8 instruction on B200 as a function of N, with the A operand in
// shared memory (.ss) or in tensor memory (.ts): which MMA shape should the Toeplitz FIR use?
// One CTA per SM, one thread issues ITERS MMAs back to back (operands are whatever the memory holds), commit, wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes)
{
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

template <int N, bool TS, int KIND /*0 i8, 1 f8f6f4*/, int ND /* distinct accumulators in rotation */, int MM = 128>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tptr;
    const uint32_t idesc = (KIND == 0 ? (2u << 4) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);   // S32 / F32 accum, M = MM
    if (threadIdx.x == 0) {
        const uint64_t adesc = make_desc(smem_u32(smem), 2048);
        const uint64_t bdesc = make_desc(smem_u32(smem) + 16384, N * 16);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tmem + 256 + (i % ND) * N;
            if (TS) {
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(d), "r"(tmem + (i & 7) * 8), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(d), "r"(tmem + (i & 7) * 8), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
            } else {
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n"
                     :: "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// NW warps issue their own MMA streams (own accumulator each) at the same time: is the ~52 clk per instruction a limit
// of the issuing thread or of the tensor pipe?
template <int N, int NW>
__global__ void __launch_bounds__(128, 1) k_multi(long long *out, int iters)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; w++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[w])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tptr;
    const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    long long t0 = clock64();
    if (lane == 0 && warp < NW) {
        const uint64_t adesc = make_desc(smem_u32(smem) + warp * 4096, 2048);
        const uint64_t bdesc = make_desc(smem_u32(smem) + 16384 + warp * 4096, N * 16);
        for (int i = 0; i < iters; i++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(tmem + warp * 64), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[warp])) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n"
                     :: "r"(smem_u32(&bar[warp])) : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
template <int N, int NW>
void run_multi(long long *d_out)
{
    const int iters = 4096;
    cudaFuncSetAttribute(k_multi<N, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; rep++) k_multi<N, NW><<<148, 128, 48 * 1024>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0;
    cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
    printf("i8 N=%3d, %d issuing warps: %7.1f clk per MMA of the SM (%d MMAs)  (%s)\n", N, NW, (double)clk / (iters * NW), iters * NW, cudaGetErrorString(e));
}

template <int N, bool TS, int KIND, int ND = 1, int MM = 128>
void run(const char *name, long long *d_out)
{
    const int iters = 4096;
    cudaFuncSetAttribute(k<N, TS, KIND, ND, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; rep++) k<N, TS, KIND, ND, MM><<<148, 128, 48 * 1024>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0;
    cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-8s M=%3d N=%3d %s, %d accumulator(s) in rotation: %7.1f clk / MMA  (%s)\n", name, MM, N, TS ? "A in TMEM" : "A in smem", ND, (double)clk / iters, cudaGetErrorString(e));
}

int main()
{
    long long *d_out;
    cudaMalloc(&d_out, 64);
    run<32, false, 0>("i8", d_out);  run<64, false, 0>("i8", d_out);  run<128, false, 0>("i8", d_out); run<256, false, 0>("i8", d_out);
    run<32, true, 0>("i8", d_out);   run<64, true, 0>("i8", d_out);   run<128, true, 0>("i8", d_out);  run<256, true, 0>("i8", d_out);
    run<16, true, 0>("i8", d_out);   run<16, false, 0>("i8", d_out);
    run<32, false, 0, 2>("i8", d_out); run<32, false, 0, 4>("i8", d_out); run<32, false, 0, 8>("i8", d_out);
    run<32, true, 0, 4>("i8", d_out); run<32, true, 0, 8>("i8", d_out); run<16, true, 0, 8>("i8", d_out); run<64, true, 0, 4>("i8", d_out);
    run<32, false, 1>("f8f6f4", d_out); run<32, true, 1>("f8f6f4", d_out); run<256, false, 1>("f8f6f4", d_out);
    run<32, false, 0, 1, 64>("i8", d_out); run<64, false, 0, 1, 64>("i8", d_out); run<32, false, 0, 4, 64>("i8", d_out);   // M = 64: is a half-height MMA cheaper?
    run_multi<32, 1>(d_out); run_multi<32, 2>(d_out); run_multi<32, 4>(d_out); run_multi<64, 4>(d_out);
    return 0;
}
