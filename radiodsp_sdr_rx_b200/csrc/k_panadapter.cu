// k_panadapter.cu — K11: panadapter trace smoothing, S-meter level and waterfall history for every channel.
//
// Replaces the pre-processing loop of Update_Panadapter (RDSP_display.h:260-280: 5-tap frequency
// smoothing 0.7/0.3/0.15 in double, 0.7*2*sqrt(|avg|*5) + 0.3*old in float, truncated to uint16) and
// Update_smeter (RDSP_display.h:366-374: sum of bins 75..85, /5).  One thread per bin.
// Waterfall (RDSP_display.h:30,282-319): a ring [50][128] of u16 per channel; the new line SpectrumView[2x] lands in the
// slot the host hands out (older lines "move down" by moving the head, not the data); k_waterfall_read unrolls the ring
// into row order (row 0 newest) and classifies every cell by the sketch's colour thresholds.  The reference's shift loop
// also reads row -1 (out of bounds); here row 0 simply is the new line.
#include "rdsp_common.cuh"
#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) k_panadapter(PanArgs a)
{
    __shared__ uint16_t s_o[256];
    const int ch = a.ch_first + blockIdx.x;
    const int x = threadIdx.x;
    s_o[x] = a.spec[(size_t)ch * 256 + x];
    __syncthreads();
    float avg;
    if (x > 1 && x < 254)
        avg = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)s_o[x], 0.7), __dmul_rn((double)s_o[x - 1], 0.3)),
                                                             __dmul_rn((double)s_o[x - 2], 0.15)),
                                                   __dmul_rn((double)s_o[x + 1], 0.3)),
                                         __dmul_rn((double)s_o[x + 2], 0.15));
    else
        avg = (float)s_o[x];
    const float LPF = 0.7f;
    const float old = (float)a.view[(size_t)ch * 256 + x];
    const float val = __fadd_rn(__fmul_rn(__fmul_rn(LPF, 2.0f), sqrtf(__fmul_rn(fabsf(avg), 5.0f))), __fmul_rn(1.0f - LPF, old));
    a.view[(size_t)ch * 256 + x] = (uint16_t)val;
    if (a.waterfall && (x & 1) == 0)
        a.waterfall[((size_t)ch * 50 + a.wf_head[blockIdx.x]) * 128 + (x >> 1)] = (uint16_t)val;
    if (x == 0) {
        float s = 0.0f;
        for (int m = 75; m <= 85; m++) s = s + (float)s_o[m];
        a.smeter[ch] = fabsf(s / 5.0f);
    }
}

__global__ void __launch_bounds__(128) k_waterfall_read(WaterfallArgs a)
{
    const int ch = a.ch_first + blockIdx.x, col = threadIdx.x;
    const int head = a.wf_head[blockIdx.x];
    for (int row = 0; row < 50; row++) {
        const uint16_t v = a.ring[((size_t)ch * 50 + (head + row) % 50) * 128 + col];
        const size_t o = ((size_t)blockIdx.x * 50 + row) * 128 + col;
        a.rows[o] = v;
        if (a.colour) a.colour[o] = v >= 75 ? 6 : v >= 50 ? 5 : v >= 40 ? 4 : v >= 25 ? 3 : v >= 15 ? 2 : v >= 5 ? 1 : 0;
    }
}

}  // namespace

void launch_waterfall_read(const WaterfallArgs &a, cudaStream_t st)
{
    if (a.ch_count <= 0) return;
    k_waterfall_read<<<a.ch_count, 128, 0, st>>>(a);
}

void launch_panadapter(const PanArgs &a, cudaStream_t st)
{
    if (a.ch_count <= 0) return;
    k_panadapter<<<a.ch_count, 256, 0, st>>>(a);
}
