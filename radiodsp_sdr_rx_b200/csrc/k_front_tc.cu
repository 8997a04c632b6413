// k_front_tc.cu — K0+K1+K2 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Same arithmetic as k_front.cu (the CUDA-core version, kept as cross-check): AudioMixer4 gain / IQ balance, the
// +-45 degree Hilbert FIR pair, sideband sum or AM envelope, audio band-pass FIR — all q15 with the 32-bit wrap-around
// accumulator of arm_fir_fast_q15 and SSAT(acc >> 15, 16); bit-exact against oracle/rdsp_oracle.c:stage_frontend.
// Replaces the AudioSDRpreProcessor -> AudioSDR front half of the graph wired at RadioDSP_SDR_RX.ino:71-72,81-82.
//
// Toeplitz formulation (north_star: "only if it measurably wins on ncu" — it does, see profiles/):
//   * one MMA row = one channel, K = time.  A[row][k] is simply the channel's delay line (K-major = the natural
//     row layout), B[k][n] = h[128 + n - k] is a constant banded Toeplitz image of the 129 taps (built on the host,
//     once per tap row), D[row][n] = 32 consecutive outputs.  K = 160 = 128 + 32: 24 % padding, no other waste.
//   * q15 x q15 in Z/2^32 from int8 tensor-core products: x = 256 xh + xl (xh signed, xl unsigned), same for h:
//         x h = 65536 xh hh + 256 (xh hl + xl hh) + xl hl
//     Three TMEM accumulators (ll, mid, hh), four MMAs per 32-byte K step with the four signedness combinations of
//     kind::i8.  No partial sum exceeds 2^24, the recombination wraps in 32-bit registers like the CMSIS accumulator.
//   * r02, "paired" images: with the taps split into SIGNED digits (h = 256 hh + hl, hl in [-128, 127]; possible for every
//     tap below 32640) the two digit planes of a tap row have the same signedness and sit side by side as ONE B operand of
//     N = 64 columns [hl | hh].  A K step is then two MMAs instead of four — xl x [hl | hh] into the columns [ll | mid] and
//     xh x [hl | hh] into the columns [mid | hh], the second one simply 32 columns further right, so that its hl half
//     accumulates onto the lh half of the first — 11 MMAs per FIR and chunk instead of 20, and 12 KB instead of 20 KB of
//     operand fetch per K step (the kernel is bound by the per-instruction cost and the shared-memory operand traffic of
//     its small MMAs).  Accumulator columns, epilogues and results are unchanged.  A tile whose tap rows
//     hold a tap >= 32640 keeps the classic four-product path (both images of every row are in the table).
//   * delay lines live in shared memory as two byte planes per line (I', Q', demodulated), each a ring of six
//     32-sample slices in the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes): slice =
//     [2 k-groups][128 rows][16 B].  An MMA K step reads one slice, so the ring never moves data: chunk c reads slices
//     c .. c+4 (mod 6) and the loader fills slice c+5 meanwhile.
//   * one CTA = one tile of 128 channels that share their three tap rows (host-built tile table; channels of a tile
//     need not be contiguous) and one time segment of the call's T blocks, walked as chunks of 32 samples through a
//     warp-specialised pipeline (loader / MMA issue / two epilogues, see k_front_tc below).
//   * the loop bodies are deliberately NOT unrolled beyond 8 outputs: four roles run four different instruction streams
//     on every SM sub-partition, and the first fully unrolled version spent 30 % of its issue slots waiting for
//     instruction fetch (ncu: stall_no_inst) with 134 KB of SASS.
#include "rdsp_common.cuh"
#include "kernels.h"
#include <vector>
#include <map>
#include <array>
#include <cstring>

#ifndef RDSP_TC_A_TMEM
#define RDSP_TC_A_TMEM 0
#endif
#define RDSP_TC_A_TMEM_EARLY RDSP_TC_A_TMEM

namespace {

constexpr int ROWS = 128;                     // channels per tile = MMA M
constexpr int NOUT = 32;                      // outputs per chunk = MMA N
constexpr int KSTEPS = 5;                     // 160 bytes of K, 32 per MMA
constexpr int SLICES = 6;
constexpr int SLICE_B = 2 * ROWS * 16;        // 4096: [2 k-groups][128 rows][16 bytes]
constexpr int PLANE_B = SLICES * SLICE_B;     // 24576
constexpr int TOEP_PLANE_B = 10 * NOUT * 16;  // 5120: [10 k-groups][32 outputs][16 bytes]
constexpr int TOEP_SET_B = 2 * TOEP_PLANE_B;  // lo plane, hi plane

constexpr int NLD = RDSP_TC_A_TMEM_EARLY ? 4 : 8;             // loader warps (the TMEM-operand variant maps them on lane quadrants)
constexpr int ITEMS = 32 / NLD;                               // (row group, k-group) items per loader warp and chunk
// warps 0-3: epilogue 1, outputs 0..15 of a chunk; W_E1B..+3: epilogue 1, outputs 16..31 (same TMEM lane quadrants)
constexpr int W_E2 = 4, W_LD = 8, W_E1B = W_LD + NLD, W_MMA = W_E1B + 4, W_MMA2 = W_MMA + 1, W_MMA3 = W_MMA2 + 1, NWARPS = W_MMA3 + 1;
static_assert(W_E1B % 4 == 0, "the second epilogue-1 group must sit on the lane quadrants of its warp indices");
constexpr int NTHREADS = NWARPS * 32;

constexpr int OFF_RING = 0;                                   // [line 0..2][plane 0..1][PLANE_B]
constexpr int OFF_TAPS = 6 * PLANE_B;                         // [set 0..2][plane 0..1][TOEP_PLANE_B]
constexpr int OFF_STAGE = OFF_TAPS + 3 * TOEP_SET_B;          // loader staging [2][4 warps][8 items][32 lanes][16 B]
constexpr int STAGE_B = 4 * 8 * 32 * 16;                      // 16384 per buffer
constexpr int OFF_META = OFF_STAGE + 2 * STAGE_B;             // per-row parameters, barriers, sqrt seed table
constexpr int META_B = ROWS * 16 + 128 + 80 + ROWS * 16;
constexpr int SMEM_B = OFF_META + META_B;
static_assert(SMEM_B <= 227 * 1024, "shared memory budget");

// TMEM columns (480 of 512): two Hilbert accumulator sets [I ll|mid|hh][Q ll|mid|hh] x 32 @0 / @192, band-pass
// [ll|mid|hh] x 32 @384.  With -DRDSP_TC_A_TMEM=1 the I' / Q' byte planes are ALSO kept in TMEM ([0,192): plane p at
// 48 p, slice s at + 8 s, lane = row) and are the A operand of the Hilbert MMAs (.ts form), with one accumulator set
// @192.  Built, bit-exact, and NOT faster (93 vs 80 us): tools/ubench_umma.cu shows that a tcgen05.mma of M = 128,
// K = 32 costs ~52 clk for every N <= 64 whether A comes from shared memory or TMEM (64 clk at N = 128, 128 at
// N = 256; ~40 clk with several issuing warps) — the N = 32 MMAs of this kernel pay a fixed per-instruction cost, not
// operand fetch, and the .ts form only loses the second accumulator set.  Kept as an option for the record.
constexpr int TM_COLS = 512;
constexpr int TM_RING = 0, TM_PLANE = 48;
constexpr int TM_ACC1B = 192, TM_ACC2P = 384;

// mbarriers (index = chunk parity unless single):
constexpr int B_IN_FULL = 0, B_M1_DONE = 2, B_E1_DONE = 4, B_M2_DONE = 6, B_E2_DONE = 8, B_TAPS = 9;
constexpr int TOEP_ROW_B = 256, TOEP_ROWS_PER_SET = TOEP_SET_B / TOEP_ROW_B;   // the table as a 2-D byte tensor for TMA

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 @4, a_format @7, b_format @10 (1 = signed),
// K-major A and B, N >> 3 @17, M >> 4 @24
constexpr uint32_t idesc(int n, bool a_signed, bool b_signed)
{
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24) | (a_signed ? (1u << 7) : 0u) | (b_signed ? (1u << 10) : 0u);
}
constexpr uint32_t IDESC_UU = idesc(NOUT, false, false);
constexpr uint32_t IDESC_SU = idesc(NOUT, true, false);
constexpr uint32_t IDESC_US = idesc(NOUT, false, true);
constexpr uint32_t IDESC_SS = idesc(NOUT, true, true);
// paired images: B = [hl | hh], both signed, N = 64 (or one half of it, N = 32)
constexpr uint32_t IDESC_P_US64 = idesc(2 * NOUT, false, true), IDESC_P_SS64 = idesc(2 * NOUT, true, true), IDESC_P_SS32 = idesc(NOUT, true, true);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: start >> 4 @0, LBO >> 4 @16 (between the two 16-byte K groups),
// SBO >> 4 @32 (between 8-row groups), version 1 @46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes)
{
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
// bounded wait for the one-shot barriers of the prologue: a descriptor or byte-count mistake must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// TMA: one box of a 2-D tensor, global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const int4 a, const int4 b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(saddr), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// AudioMixer4 gain: sat16((mult * x) >> 16).  mult = 65536 mh + ml (ml unsigned): the shift distributes exactly,
// (mult x) >> 16 = mh x + ((ml x) >> 16), and |ml x| < 2^31 — three 32-bit operations, unity gain included.
__device__ __forceinline__ int32_t mix_gain_tc(int32_t x, int32_t mh, int32_t ml)
{
    return sat16(mh * x + ((ml * x) >> 16));
}

__device__ __forceinline__ int32_t recombine(uint32_t ll, uint32_t mid, uint32_t hh)
{
    const uint32_t acc = ll + (mid << 8) + (hh << 16);                   // wraps like the CMSIS fast-FIR accumulator
    return sat16(((int32_t)acc) >> 15);
}

struct TcSmem {
    uint8_t *base;
    __device__ __forceinline__ uint8_t *ring(int line, int plane) const { return base + OFF_RING + (line * 2 + plane) * PLANE_B; }
    __device__ __forceinline__ uint8_t *taps(int set, int plane) const { return base + OFF_TAPS + (set * 2 + plane) * TOEP_PLANE_B; }
    __device__ __forceinline__ uint8_t *stage(int buf, int lw) const { return base + OFF_STAGE + buf * STAGE_B + lw * (STAGE_B / NLD); }
    __device__ __forceinline__ int *row_ch() const { return reinterpret_cast<int *>(base + OFF_META); }
    __device__ __forceinline__ int *row_mi() const { return reinterpret_cast<int *>(base + OFF_META + ROWS * 4); }
    __device__ __forceinline__ int *row_mq() const { return reinterpret_cast<int *>(base + OFF_META + ROWS * 8); }
    __device__ __forceinline__ int *row_usb() const { return reinterpret_cast<int *>(base + OFF_META + ROWS * 12); }
    __device__ __forceinline__ uint64_t *bars() const { return reinterpret_cast<uint64_t *>(base + OFF_META + ROWS * 16); }
    __device__ __forceinline__ uint32_t *tmem_ptr() const { return reinterpret_cast<uint32_t *>(base + OFF_META + ROWS * 16 + 96); }
    __device__ __forceinline__ uint16_t *sqrt_guess() const { return reinterpret_cast<uint16_t *>(base + OFF_META + ROWS * 16 + 128); }
    // noise blanker, per row: Q8 threshold factor (0 = off), running magnitude, threshold of the current chunk, chunk sum
    __device__ __forceinline__ uint32_t *nb_mult() const { return reinterpret_cast<uint32_t *>(base + OFF_META + ROWS * 16 + 208); }
    __device__ __forceinline__ int32_t *nb_ref() const { return reinterpret_cast<int32_t *>(base + OFF_META + ROWS * 20 + 208); }
    __device__ __forceinline__ uint32_t *nb_thr() const { return reinterpret_cast<uint32_t *>(base + OFF_META + ROWS * 24 + 208); }
    __device__ __forceinline__ uint32_t *nb_sum() const { return reinterpret_cast<uint32_t *>(base + OFF_META + ROWS * 28 + 208); }
};

// ---- loader -------------------------------------------------------------------------------------------------------
// 32 IQ frames of every row -> gain -> byte planes of the I' and Q' rings.  A warp instruction covers 8 rows x 64 bytes
// (4 lanes x 16 B per row): full 32-byte sectors from HBM, and each of the four 32-bit plane stores of a lane lands in a
// 128-byte contiguous run (8 rows x 16 B): conflict free.  Four warps share the 32 (row group, k-group) items of a
// chunk; the 8 x 16 bytes of a lane go through a private cp.async staging slot, one chunk period ahead of their use.
__device__ __forceinline__ void fetch_chunk(const TcSmem &s, const FrontArgs &a, int t, int cq, int buf, int lw, int lane)
{
    const int rr = lane >> 2, q = lane & 3;
    const uint32_t st = smem_u32(s.stage(buf, lw)) + lane * 16;
#pragma unroll 1
    for (int k = 0; k < ITEMS; k++) {
        const int i = lw + NLD * k, g = i & 1, row = (i >> 1) * 8 + rr;
        const int ch = s.row_ch()[row];
        const int16_t *src = a.iq + (((size_t)t * a.C + (ch >= 0 ? ch : 0)) * RDSP_BLK + cq * NOUT + g * 16 + q * 4) * 2;
        cp_async16(st + k * 512, src, ch >= 0 ? 16u : 0u);               // padding rows: zero fill
    }
    cp_async_commit();
}

template <bool NB>
__device__ __forceinline__ void split_chunk(const TcSmem &s, int buf, int slice, int lw, int lane)
{
    const int rr = lane >> 2, q = lane & 3;
    const uint8_t *st = s.stage(buf, lw) + lane * 16;
#pragma unroll 1
    for (int k = 0; k < ITEMS; k++) {
        const int i = lw + NLD * k, g = i & 1, row = (i >> 1) * 8 + rr;
        const int4 v = *reinterpret_cast<const int4 *>(st + k * 512);
        uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
        const int mi = s.row_mi()[row], mq = s.row_mq()[row];
        const int mih = mi >> 16, mil = mi & 0xFFFF, mqh = mq >> 16, mql = mq & 0xFFFF;
        if (NB) {
            // noise blanker (oracle/rdsp_oracle.c:stage_frontend): zero the frames above the chunk's threshold, feed
            // the clipped magnitudes into the row's chunk sum
            const uint32_t thr = s.nb_thr()[row];
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int32_t xi = mix_gain_tc(lo16(w[j]), mih, mil), xq = mix_gain_tc(hi16(w[j]), mqh, mql);
                const uint32_t mag = (uint32_t)abs(xi) + (uint32_t)abs(xq);
                const bool hit = mag > thr;
                sum += hit ? thr : mag;
                w[j] = hit ? 0u : mk16(xi, xq);
            }
            if (s.nb_mult()[row]) atomicAdd(&s.nb_sum()[row], sum);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = mk16(mix_gain_tc(lo16(w[j]), mih, mil), mix_gain_tc(hi16(w[j]), mqh, mql));
        }
        // word = [I lo, I hi, Q lo, Q hi]; 4 x 4 byte transpose
        const uint32_t i01 = prmt(w[0], w[1], 0x5140), q01 = prmt(w[0], w[1], 0x7362);
        const uint32_t i23 = prmt(w[2], w[3], 0x5140), q23 = prmt(w[2], w[3], 0x7362);
        const int off = slice * SLICE_B + g * (ROWS * 16) + row * 16 + q * 4;
        *reinterpret_cast<uint32_t *>(s.ring(0, 0) + off) = prmt(i01, i23, 0x5410);
        *reinterpret_cast<uint32_t *>(s.ring(0, 1) + off) = prmt(i01, i23, 0x7632);
        *reinterpret_cast<uint32_t *>(s.ring(1, 0) + off) = prmt(q01, q23, 0x5410);
        *reinterpret_cast<uint32_t *>(s.ring(1, 1) + off) = prmt(q01, q23, 0x7632);
    }
}

// after every loader thread has added its frames of a chunk: thread r < 128 closes the chunk of row r (running
// magnitude, next threshold); two named barriers among the loader threads fence the exchange
__device__ __forceinline__ void nb_close_chunk(const TcSmem &s, int r)
{
    asm volatile("bar.sync 1, %0;" :: "n"(NLD * 32) : "memory");
    const uint32_t mult = r < ROWS ? s.nb_mult()[r] : 0u;
    if (mult) {
        int32_t ref = s.nb_ref()[r];
        const int32_t cm = (int32_t)(s.nb_sum()[r] >> 5);
        ref = ref > 0 ? ref + ((cm - ref) >> 3) : cm;
        s.nb_ref()[r] = ref;
        s.nb_thr()[r] = ref > 0 ? (uint32_t)(((uint32_t)ref * mult) >> 8) : 0xFFFFFFFFu;
        s.nb_sum()[r] = 0u;
    }
    asm volatile("bar.sync 1, %0;" :: "n"(NLD * 32) : "memory");
}

// ---- delay-line state <-> ring, all warps -------------------------------------------------------------------------
// One warp item = 8 rows x 64 bytes of one line (lane = k-group select, row, 8-sample half): 64-byte runs in HBM,
// 128-byte runs in shared memory.  192 items per tile.
template <bool STORE>
__device__ __forceinline__ void state_io(const TcSmem &s, int16_t *hist, int first_slice, int warp, int lane)
{
    const int kgsel = lane >> 4, row8 = (lane >> 1) & 7, half = lane & 1;
#pragma unroll 1
    for (int item = warp; item < 3 * 16 * 4; item += NWARPS) {
        const int line = item >> 6, rg = (item >> 2) & 15, kg = (item & 3) * 2 + kgsel;
        const int row = rg * 8 + row8, ch = s.row_ch()[row];
        const int off = ((first_slice + (kg >> 1)) % SLICES) * SLICE_B + (kg & 1) * (ROWS * 16) + row * 16 + half * 8;
        int2 *plo = reinterpret_cast<int2 *>(s.ring(line, 0) + off), *phi = reinterpret_cast<int2 *>(s.ring(line, 1) + off);
        int4 *g = ch >= 0 ? reinterpret_cast<int4 *>(hist + ((size_t)ch * 3 + line) * RDSP_BLK + (kg * 2 + half) * 8) : nullptr;
        if (STORE) {
            if (g) {
                const int2 lo = *plo, hi = *phi;
                *g = make_int4((int)prmt((uint32_t)lo.x, (uint32_t)hi.x, 0x5140), (int)prmt((uint32_t)lo.x, (uint32_t)hi.x, 0x7362),
                               (int)prmt((uint32_t)lo.y, (uint32_t)hi.y, 0x5140), (int)prmt((uint32_t)lo.y, (uint32_t)hi.y, 0x7362));
            }
        } else {
            int4 v = make_int4(0, 0, 0, 0);
            if (g) v = *g;
            *plo = make_int2((int)prmt((uint32_t)v.x, (uint32_t)v.y, 0x6420), (int)prmt((uint32_t)v.z, (uint32_t)v.w, 0x6420));
            *phi = make_int2((int)prmt((uint32_t)v.x, (uint32_t)v.y, 0x7531), (int)prmt((uint32_t)v.z, (uint32_t)v.w, 0x7531));
        }
    }
}

// ---- MMA issue ----------------------------------------------------------------------------------------------------
// one 129-tap FIR over 128 rows x 32 outputs: 5 K steps x 4 byte-plane products into (ll, mid, hh) at tmem_d
// PART 0: all three accumulators; 1: ll and hh; 2: mid — two issuing threads may share one FIR, each owning whole
// accumulators (the order between MMAs of different threads is not defined, the first write of an accumulator must
// clear it)
template <int PART = 0>
__device__ __forceinline__ void issue_fir(uint32_t ring_lo, uint32_t ring_hi, uint32_t taps_lo, uint32_t taps_hi, uint32_t tmem_d, int c)
{
    int sl = c % SLICES;
#pragma unroll 1
    for (int ks = 0; ks < KSTEPS; ks++) {
        const uint32_t so = (uint32_t)(sl * SLICE_B);
        const uint64_t a_lo = make_desc(ring_lo + so, ROWS * 16), a_hi = make_desc(ring_hi + so, ROWS * 16);
        const uint64_t b_lo = make_desc(taps_lo + ks * 2 * NOUT * 16, NOUT * 16), b_hi = make_desc(taps_hi + ks * 2 * NOUT * 16, NOUT * 16);
        if (PART != 2) mma_i8(tmem_d + 0 * NOUT, a_lo, b_lo, IDESC_UU, ks > 0);
        if (PART != 1) mma_i8(tmem_d + 1 * NOUT, a_hi, b_lo, IDESC_SU, ks > 0);
        if (PART != 1) mma_i8(tmem_d + 1 * NOUT, a_lo, b_hi, IDESC_US, 1);
        if (PART != 2) mma_i8(tmem_d + 2 * NOUT, a_hi, b_hi, IDESC_SS, ks > 0);
        sl = sl + 1 == SLICES ? 0 : sl + 1;
    }
}

// the same FIR from a PAIRED image (taps = [10 k-groups][64 columns: hl of output n | hh of output n][16 B], signed digits):
// two N = 64 MMAs per K step; the first K step clears the accumulators with three (the hh columns are cleared by an N = 32
// MMA of their own, because the N = 64 MMA that covers them must accumulate onto mid)
__device__ __forceinline__ void issue_fir_paired(uint32_t ring_lo, uint32_t ring_hi, uint32_t taps, uint32_t tmem_d, int c)
{
    int sl = c % SLICES;
#pragma unroll 1
    for (int ks = 0; ks < KSTEPS; ks++) {
        const uint32_t so = (uint32_t)(sl * SLICE_B);
        const uint64_t a_lo = make_desc(ring_lo + so, ROWS * 16), a_hi = make_desc(ring_hi + so, ROWS * 16);
        const uint32_t tb = taps + ks * 2 * (2 * NOUT) * 16;
        const uint64_t b_all = make_desc(tb, 2 * NOUT * 16);
        if (ks == 0) {
            mma_i8(tmem_d + 0 * NOUT, a_lo, b_all, IDESC_P_US64, 0);                                   // [ll | mid] = xl x [hl | hh]
            mma_i8(tmem_d + 2 * NOUT, a_hi, make_desc(tb + NOUT * 16, 2 * NOUT * 16), IDESC_P_SS32, 0);   // hh = xh x hh
            mma_i8(tmem_d + 1 * NOUT, a_hi, b_all, IDESC_P_SS32, 1);                                   // mid += xh x hl
        } else {
            mma_i8(tmem_d + 0 * NOUT, a_lo, b_all, IDESC_P_US64, 1);
            mma_i8(tmem_d + 1 * NOUT, a_hi, b_all, IDESC_P_SS64, 1);                                   // [mid | hh] += xh x [hl | hh]
        }
        sl = sl + 1 == SLICES ? 0 : sl + 1;
    }
}

// the same FIR with the delay-line planes as TMEM A operand: ring_lo / ring_hi = TMEM column of slice 0 of the plane
__device__ __forceinline__ void issue_fir_ts(uint32_t ring_lo, uint32_t ring_hi, uint32_t taps_lo, uint32_t taps_hi, uint32_t tmem_d, int c)
{
    int sl = c % SLICES;
#pragma unroll 1
    for (int ks = 0; ks < KSTEPS; ks++) {
        const uint32_t a_lo = ring_lo + 8u * (uint32_t)sl, a_hi = ring_hi + 8u * (uint32_t)sl;
        const uint64_t b_lo = make_desc(taps_lo + ks * 2 * NOUT * 16, NOUT * 16), b_hi = make_desc(taps_hi + ks * 2 * NOUT * 16, NOUT * 16);
        mma_i8_ts(tmem_d + 0 * NOUT, a_lo, b_lo, IDESC_UU, ks > 0);
        mma_i8_ts(tmem_d + 1 * NOUT, a_hi, b_lo, IDESC_SU, ks > 0);
        mma_i8_ts(tmem_d + 1 * NOUT, a_lo, b_hi, IDESC_US, 1);
        mma_i8_ts(tmem_d + 2 * NOUT, a_hi, b_hi, IDESC_SS, ks > 0);
        sl = sl + 1 == SLICES ? 0 : sl + 1;
    }
}

// one 32-sample slice of the four I' / Q' byte planes, shared memory -> TMEM; the thread of row r moves its row
// (tmem_lane = TMEM base with the lane offset of this warp's quadrant)
__device__ __forceinline__ void slice_to_tmem(const TcSmem &s, uint32_t tmem_lane, int slice, int r)
{
#pragma unroll
    for (int pl = 0; pl < 4; pl++) {
        const uint8_t *src = s.ring(pl >> 1, pl & 1) + slice * SLICE_B + r * 16;
        tmem_st8(tmem_lane + TM_RING + pl * TM_PLANE + slice * 8, *reinterpret_cast<const int4 *>(src),
                 *reinterpret_cast<const int4 *>(src + ROWS * 16));
    }
}

// ---- epilogues ----------------------------------------------------------------------------------------------------
// epilogue 1: I' and Q' accumulators -> q15 -> sideband sum / envelope / synchronous detection -> D slice (byte planes),
// 8 outputs per round.  MODE 0: QADD16 / QSUB16, 1: AM envelope, 2: SAM (carrier PLL, sequential in time: the thread
// of a row walks its samples in order, chunk after chunk, with the loop state in registers).
struct SamState { float phi, omega, dc; };

template <int MODE>
__device__ __forceinline__ void epilogue1(const TcSmem &s, uint32_t tmem_row, int row, int slice, bool usb, SamState &sam, int it0, int it1)
{
#pragma unroll 1
    for (int it = it0; it < it1; it++) {
        uint32_t v[6][8];
#pragma unroll
        for (int k = 0; k < 6; k++) tmem_ld8(tmem_row + k * NOUT + it * 8, v[k]);
        tmem_ld_wait();
        int32_t d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int32_t ya = recombine(v[0][j], v[1][j], v[2][j]);
            const int32_t yb = recombine(v[3][j], v[4][j], v[5][j]);
            if (MODE == 1) {
                const uint32_t e = sqrt_u32_approx_fast((uint32_t)(ya * ya) + (uint32_t)(yb * yb), s.sqrt_guess());
                d[j] = (int32_t)min(e, 32767u);
            } else if (MODE == 2) {
                // oracle/rdsp_oracle.c:stage_frontend, RDSP_DEMOD_SAM
                float sn, cs;
                __sincosf(sam.phi, &sn, &cs);                               // |phi| <= pi: 2^-21 absolute, no slow path / stack
                const float fa = (float)ya, fb = (float)yb;
                const float re = fa * cs + fb * sn, im = fb * cs - fa * sn;
                const float err = (re == 0.0f && im == 0.0f) ? 0.0f : atan2f(im, re);
                sam.omega += RDSP_SAM_K2 * err;
                sam.omega = fminf(fmaxf(sam.omega, -RDSP_SAM_WMAX), RDSP_SAM_WMAX);
                sam.phi += sam.omega + RDSP_SAM_K1 * err;
                if (sam.phi >= RDSP_SAM_PI) sam.phi -= 2.0f * RDSP_SAM_PI;
                if (sam.phi < -RDSP_SAM_PI) sam.phi += 2.0f * RDSP_SAM_PI;
                sam.dc += (re - sam.dc) * RDSP_SAM_ADC;
                const float o = fminf(fmaxf(re - sam.dc, -32768.0f), 32767.0f);
                d[j] = (int32_t)o;                                           // truncation toward zero
            } else {
                d[j] = sat16(usb ? ya - yb : ya + yb);
            }
        }
        const uint32_t p01 = prmt((uint32_t)d[0], (uint32_t)d[1], 0x5140), p23 = prmt((uint32_t)d[2], (uint32_t)d[3], 0x5140);   // [l0 l1 h0 h1]
        const uint32_t p45 = prmt((uint32_t)d[4], (uint32_t)d[5], 0x5140), p67 = prmt((uint32_t)d[6], (uint32_t)d[7], 0x5140);
        const int off = slice * SLICE_B + (it >> 1) * (ROWS * 16) + row * 16 + (it & 1) * 8;
        *reinterpret_cast<int2 *>(s.ring(2, 0) + off) = make_int2((int)prmt(p01, p23, 0x5410), (int)prmt(p45, p67, 0x5410));
        *reinterpret_cast<int2 *>(s.ring(2, 1) + off) = make_int2((int)prmt(p01, p23, 0x7632), (int)prmt(p45, p67, 0x7632));
    }
}

// epilogue 2: band-pass accumulators -> q15 -> audio rows.  The one band-pass accumulator set is the shortest cycle of the
// chunk pipeline (band-pass MMAs of chunk c+1 wait for this drain; phase timers r02: issue 1.6 k + drain-and-store 2.0 k clk of a
// 3.9 k clk chunk period), so the drain is separated from the stores: both halves of the 32 outputs are pulled out of TMEM with
// six loads in flight each and recombined into 16 packed registers, the accumulators are handed back (e2_done), and only then
// do the global stores go out.
__device__ __forceinline__ void epilogue2(const FrontArgs &a, uint32_t tmem_row, int ch, int t, int cq, uint32_t bar_e2_done)
{
    uint32_t pk[NOUT / 2];                                                   // outputs 2i, 2i+1 as a q15 pair
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t v[3][16];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            tmem_ld8(tmem_row + k * NOUT + half * 16, v[k]);
            tmem_ld8(tmem_row + k * NOUT + half * 16 + 8, v[k] + 8);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; j++)
            pk[half * 8 + j] = mk16(recombine(v[0][2 * j], v[1][2 * j], v[2][2 * j]), recombine(v[0][2 * j + 1], v[1][2 * j + 1], v[2][2 * j + 1]));
    }
    tc_fence_before();
    mbar_arrive(bar_e2_done);                                                // acc2 is free: the stores below no longer hold the tensor core up
    if (ch < 0) return;
    const size_t n0 = ((size_t)t * a.C + ch) * RDSP_BLK + cq * NOUT;
    if (a.out_mono) {
#pragma unroll
        for (int q = 0; q < 4; q++) st_stream16(a.out_mono + n0 + 8 * q, make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]));
    }
    if (a.out_stereo) {
#pragma unroll
        for (int q = 0; q < 8; q++) {                                        // 4 frames (L = R) per 16-byte store
            const uint32_t p0 = pk[2 * q], p1 = pk[2 * q + 1];
            st_stream16(a.out_stereo + (n0 + 4 * q) * 2, make_int4((int)prmt(p0, p0, 0x1010), (int)prmt(p0, p0, 0x3232), (int)prmt(p1, p1, 0x1010), (int)prmt(p1, p1, 0x3232)));
        }
    }
    if (a.dbg) {
        float4 *dp = reinterpret_cast<float4 *>(a.dbg + n0 * 2);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const float f0 = (float)lo16(pk[q]) / 32768.0f, f1 = (float)hi16(pk[q]) / 32768.0f;
            dp[q] = make_float4(f0, f0, f1, f1);
        }
    }
}

#ifdef RDSP_TC_PROF
__device__ unsigned long long g_tc_prof[16];
__device__ unsigned long long g_tc_cta[4096][2];
#define TCP_BEGIN long long tp_ = clock64()
#define TCP(i) do { if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0) { const long long n_ = clock64(); g_tc_prof[i] += n_ - tp_; tp_ = n_; } } while (0)
#define TCQ(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { const long long n_ = clock64(); g_tc_prof[i] += n_ - tq_; tq_ = n_; } } while (0)
#else
#define TCP_BEGIN do { } while (0)
#define TCP(i) do { } while (0)
#define TCQ(i) do { } while (0)
#endif

// ---- the kernel: warp-specialised pipeline over the chunks of one tile -----------------------------------------------
//   warps 0-3, 16-19  epilogue 1 (TMEM lane quadrant = warp % 4) warps 8-15  loader (HBM -> gain -> byte planes)
//   warps 4-7         epilogue 2 (quadrant = warp - 4)           warps 20-22 MMA issue (one lane each: I' FIR, Q' FIR, band-pass FIR)
//   in_full[2]   loader -> MMA      chunk c's I'/Q' slice is in shared memory                      (128 arrivals)
//   m1_done[2]   MMA -> E1, loader  Hilbert-pair MMAs of chunk c retired: acc1[c&1] valid, slice c%6 free (commit)
//   e1_done[2]   E1 -> MMA          acc1[c&1] drained and the D slice of chunk c written           (256 arrivals)
//   m2_done[2]   MMA -> E2, E1      band-pass MMAs of chunk c retired: acc2 valid, D slice c%6 free     (commit)
//   e2_done      E2 -> MMA          acc2 drained                                                    (128 arrivals)
// Each FIR has an issuing thread of its own: the band-pass MMAs of chunk c run beside the Hilbert MMAs of chunk c + 1.
// WITH_SAM: the instantiation that carries the SAM detector (atan2f and the loop state cost 15 registers and a stack
// frame; with them in the common kernel every step of cfg5 was 17 % slower although no channel used SAM).
template <bool WITH_SAM>
__global__ void __launch_bounds__(NTHREADS, 1) k_front_tc(FrontArgs a, const __grid_constant__ FrontTcTables tb)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    TcSmem s{s_raw};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, seg = blockIdx.y;
    const int4 rows = tb.tile_rows[tile];                                 // x, y, z = Toeplitz images; w = AM flag
    const int dmode = rows.w & 0xFF;                                      // 0 sideband sum, 1 AM envelope, 2 SAM
    const bool paired = (rows.w >> 8) & 1;                                // the tile's three tap rows have paired images (CTA-uniform)
#ifdef RDSP_TC_PROF
    unsigned long long gt0_;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0_));
    long long tq_ = clock64();
#endif

    // this CTA's share of the T blocks: outputs of blocks [t_store, t_end); a later segment first rebuilds its
    // delay lines from the input itself (block t0-1 as I'/Q' history, block t0 = t_store-1 as warm-up for the D line)
    const int kind_seg = dmode == 1 ? 1 : 0;                              // AM tiles are cut into more segments (launch_front_tc)
    const int n_seg = tb.n_seg[kind_seg];
    if (seg >= n_seg) return;                                              // (whole CTA, before any barrier exists)
    const int t_store = tb.seg_bounds[kind_seg][seg], t_end = tb.seg_bounds[kind_seg][seg + 1];
    const int t0 = seg ? t_store - 1 : 0;
    const int nch = 4 * (t_end - t0);

    // ---- prologue (all warps) ----
    if (threadIdx.x < ROWS) {
        const int row = threadIdx.x;
        const int ch = tb.tile_ch[tile * ROWS + row];
        RdspChanParams p{};
        if (ch >= 0) p = a.par[ch];
        s.row_ch()[row] = ch;
        s.row_mi()[row] = ch >= 0 ? p.mult_i : 65536;
        s.row_mq()[row] = ch >= 0 ? p.mult_q : 65536;
        s.row_usb()[row] = (p.demod == 1 || p.demod == 3) ? 1 : 0;
        if (row < 33) s.sqrt_guess()[row] = c_sqrt_guess[row];
        const uint32_t nbm = ch >= 0 ? p.nb_mult_q8 : 0u;
        const int32_t nbr = nbm ? a.nb_ref[ch] : 0;
        s.nb_mult()[row] = nbm;
        s.nb_ref()[row] = nbr;
        s.nb_thr()[row] = (nbm && nbr > 0) ? (uint32_t)(((uint32_t)nbr * nbm) >> 8) : 0xFFFFFFFFu;
        s.nb_sum()[row] = 0u;
    }
    const uint32_t bar0 = smem_u32(s.bars());
    auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    if (threadIdx.x == 0) {
        mbar_init(bar(B_IN_FULL), NLD * 32); mbar_init(bar(B_IN_FULL + 1), NLD * 32);
        mbar_init(bar(B_M1_DONE), RDSP_TC_A_TMEM ? 1 : 2); mbar_init(bar(B_M1_DONE + 1), RDSP_TC_A_TMEM ? 1 : 2);   // two issuing threads commit (I' FIR, Q' FIR)
        mbar_init(bar(B_E1_DONE), 2 * ROWS); mbar_init(bar(B_E1_DONE + 1), 2 * ROWS);     // both epilogue-1 groups arrive
        mbar_init(bar(B_M2_DONE), 1); mbar_init(bar(B_M2_DONE + 1), 1);                   // the band-pass FIR has an issuer of its own
        mbar_init(bar(B_E2_DONE), ROWS);
        mbar_init(bar(B_TAPS), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the three Toeplitz images of this tile (I', Q', band-pass: 2 byte planes each, 30 KB) arrive by TMA — three
        // cp.async.bulk.tensor boxes issued by this one thread, completion counted in bytes on B_TAPS — while the other 703
        // threads load row parameters and delay-line state; only the MMA issuers wait for them
        mbar_expect_tx(bar(B_TAPS), 3 * TOEP_SET_B);
        const int img = paired ? 1 : 0;                                    // the table holds both images of every tap row
        tma_load_2d(smem_u32(s.taps(0, 0)), &tb.toep_map, 0, (rows.x * 2 + img) * TOEP_ROWS_PER_SET, bar(B_TAPS));
        tma_load_2d(smem_u32(s.taps(1, 0)), &tb.toep_map, 0, (rows.y * 2 + img) * TOEP_ROWS_PER_SET, bar(B_TAPS));
        tma_load_2d(smem_u32(s.taps(2, 0)), &tb.toep_map, 0, (rows.z * 2 + img) * TOEP_ROWS_PER_SET, bar(B_TAPS));
    }
    __syncthreads();                                                       // row parameters visible
    if (seg == 0) {
        state_io<false>(s, a.hist, 0, warp, lane);                         // delay lines of the previous call -> slices 0..3
    } else {
        if (warp >= W_LD && warp < W_LD + NLD) {
#pragma unroll 1
            for (int cq = 0; cq < 4; cq++) {
                fetch_chunk(s, a, t0 - 1, cq, cq & 1, warp - W_LD, lane);
                cp_async_wait<0>();
                split_chunk<false>(s, cq & 1, cq, warp - W_LD, lane);      // (a call with blanked channels is one segment)
            }
        } else if (threadIdx.x < ROWS) {
            // demodulated line: zero history (its warm-up block only has to flush the band-pass through)
            const int4 z = make_int4(0, 0, 0, 0);
#pragma unroll 1
            for (int kg = 0; kg < 8; kg++) {
                const int off = (kg >> 1) * SLICE_B + (kg & 1) * (ROWS * 16) + threadIdx.x * 16;
                *reinterpret_cast<int4 *>(s.ring(2, 0) + off) = z;
                *reinterpret_cast<int4 *>(s.ring(2, 1) + off) = z;
            }
        }
    }
    if (warp == W_MMA) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s.tmem_ptr())), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s.tmem_ptr();
#if RDSP_TC_A_TMEM
    if (warp >= W_LD && warp < W_LD + NLD) {
        const int r = (warp - W_LD) * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)((warp - W_LD) * 32) << 16);
#pragma unroll 1
        for (int sl = 0; sl < 4; sl++) slice_to_tmem(s, tl, sl, r);        // the history of the I' / Q' lines
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#endif
    TCQ(13);

    if (warp >= W_LD && warp < W_LD + NLD) {
        // ===== loader =====
        const int lw = warp - W_LD;
        bool tile_nb = false;                                              // any blanked row in this tile (CTA-uniform)
        for (int r = 0; r < ROWS; r++) tile_nb |= s.nb_mult()[r] != 0u;
        fetch_chunk(s, a, t0, 0, 0, lw, lane);
        TCP_BEGIN;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
            if (c + 1 < nch) {
                fetch_chunk(s, a, t0 + ((c + 1) >> 2), (c + 1) & 3, (c + 1) & 1, lw, lane);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            // slice (c+4)%6 was the oldest slice of chunk c-2
            if (c >= 2) mbar_wait(bar(B_M1_DONE + (c & 1)), ((c - 2) >> 1) & 1);
            if (lw == 0) TCP(0);
            if (tile_nb) {
                split_chunk<true>(s, c & 1, (c + 4) % SLICES, lw, lane);
                nb_close_chunk(s, lw * 32 + lane);
            } else {
                split_chunk<false>(s, c & 1, (c + 4) % SLICES, lw, lane);
            }
#if RDSP_TC_A_TMEM
            asm volatile("bar.sync 2, 128;" ::: "memory");               // the rows of this slice were written by all four loader warps
            slice_to_tmem(s, tmem + ((uint32_t)(lw * 32) << 16), (c + 4) % SLICES, lw * 32 + lane);
            tmem_st_wait();
            tc_fence_before();
#else
            fence_async_smem();
#endif
            mbar_arrive(bar(B_IN_FULL + (c & 1)));
            if (lw == 0) TCP(1);
        }
        if (tile_nb) {
            const int r = lw * 32 + lane, ch = r < ROWS ? s.row_ch()[r] : -1;
            if (ch >= 0 && s.nb_mult()[r]) a.nb_ref[ch] = s.nb_ref()[r];
        }
    } else if (warp == W_MMA) {
        // ===== MMA issue =====
        if (lane == 0) {
            const uint32_t rI0 = smem_u32(s.ring(0, 0)), rI1 = smem_u32(s.ring(0, 1)), rQ0 = smem_u32(s.ring(1, 0)), rQ1 = smem_u32(s.ring(1, 1));
            const uint32_t rD0 = smem_u32(s.ring(2, 0)), rD1 = smem_u32(s.ring(2, 1));
            const uint32_t tA0 = smem_u32(s.taps(0, 0)), tA1 = smem_u32(s.taps(0, 1)), tB0 = smem_u32(s.taps(1, 0)), tB1 = smem_u32(s.taps(1, 1));
            const uint32_t tM0 = smem_u32(s.taps(2, 0)), tM1 = smem_u32(s.taps(2, 1));
            mbar_wait_bounded(bar(B_TAPS), 0);                             // the Toeplitz images have landed (TMA)
            TCP_BEGIN;
#pragma unroll 1
            for (int c = 0; c < nch; c++) {
                {
                    mbar_wait(bar(B_IN_FULL + (c & 1)), (c >> 1) & 1);
                    TCP(2);
#if RDSP_TC_A_TMEM
                    if (c >= 1) mbar_wait(bar(B_E1_DONE + ((c - 1) & 1)), ((c - 1) >> 1) & 1);   // the one accumulator set is drained
                    TCP(3);
                    tc_fence_after();
                    const uint32_t acc1 = tmem + TM_ACC1B;
                    issue_fir_ts(tmem + TM_RING + 0 * TM_PLANE, tmem + TM_RING + 1 * TM_PLANE, tA0, tA1, acc1, c);
                    issue_fir_ts(tmem + TM_RING + 2 * TM_PLANE, tmem + TM_RING + 3 * TM_PLANE, tB0, tB1, acc1 + 3 * NOUT, c);
#else
                    if (c >= 2) mbar_wait(bar(B_E1_DONE + (c & 1)), ((c - 2) >> 1) & 1);
                    TCP(3);
                    tc_fence_after();
                    const uint32_t acc1 = tmem + (c & 1) * TM_ACC1B;
                    if (paired) issue_fir_paired(rI0, rI1, tA0, acc1, c);
                    else issue_fir<0>(rI0, rI1, tA0, tA1, acc1, c);         // the Q' FIR is issued by the second issuer warp
#endif
                    mma_commit(bar(B_M1_DONE + (c & 1)));
                    TCP(4);
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA2) {
        // ===== second MMA issuer: the Q' FIR of every chunk (several issuing threads keep more MMAs in flight: ~40 clk per
        // N = 32 MMA instead of ~52, tools/ubench_umma.cu); same waits as the first issuer, own commit on m1_done =====
        if (lane == 0 && !RDSP_TC_A_TMEM) {
            const uint32_t rQ0 = smem_u32(s.ring(1, 0)), rQ1 = smem_u32(s.ring(1, 1));
            const uint32_t tB0 = smem_u32(s.taps(1, 0)), tB1 = smem_u32(s.taps(1, 1));
            mbar_wait_bounded(bar(B_TAPS), 0);
#pragma unroll 1
            for (int c = 0; c < nch; c++) {
                mbar_wait(bar(B_IN_FULL + (c & 1)), (c >> 1) & 1);
                if (c >= 2) mbar_wait(bar(B_E1_DONE + (c & 1)), ((c - 2) >> 1) & 1);
                tc_fence_after();
                if (paired) issue_fir_paired(rQ0, rQ1, tB0, tmem + (c & 1) * TM_ACC1B + 3 * NOUT, c);
                else issue_fir<0>(rQ0, rQ1, tB0, tB1, tmem + (c & 1) * TM_ACC1B + 3 * NOUT, c);
                mma_commit(bar(B_M1_DONE + (c & 1)));
            }
        }
        __syncwarp();
    } else if (warp == W_MMA3) {
        // ===== third MMA issuer (r02): the band-pass FIR of every chunk.  The phase timers showed the two Hilbert issuers
        // blocked ~170 clk per MMA inside their own issue (the instruction waits for a slot in the tensor queue) and then
        // again on e1_done / e2_done before they could issue the band-pass FIR of the chunk before — one thread cannot do
        // both without serialising them.  With its own issuer the band-pass FIR of chunk c runs beside the Hilbert FIRs of
        // chunk c + 1, and the Hilbert issuers never wait for an epilogue of the other stage. =====
        if (lane == 0) {
            const uint32_t rD0 = smem_u32(s.ring(2, 0)), rD1 = smem_u32(s.ring(2, 1));
            const uint32_t tM0 = smem_u32(s.taps(2, 0)), tM1 = smem_u32(s.taps(2, 1));
            mbar_wait_bounded(bar(B_TAPS), 0);
            TCP_BEGIN;
#pragma unroll 1
            for (int cc = 0; cc < nch; cc++) {
                mbar_wait(bar(B_E1_DONE + (cc & 1)), (cc >> 1) & 1);       // the D slice of chunk cc is written
                TCP(5);
                if (cc >= 1) mbar_wait(bar(B_E2_DONE), (cc - 1) & 1);      // the one band-pass accumulator set is drained
                TCP(6);
                tc_fence_after();
                if (paired) issue_fir_paired(rD0, rD1, tM0, tmem + TM_ACC2P, cc);
                else issue_fir<0>(rD0, rD1, tM0, tM1, tmem + TM_ACC2P, cc);
                mma_commit(bar(B_M2_DONE + (cc & 1)));
                TCP(7);
            }
        }
        __syncwarp();
    } else if (warp < W_E2 || (warp >= W_E1B && warp < W_E1B + 4)) {
        // ===== epilogue 1: two groups of four warps share the 32 outputs of a chunk (the SAM loop is sequential over
        // the samples of a row: there the first group takes all of them and the second only keeps the barriers moving) =====
        const bool second = warp >= W_E1B;
        const int quad = second ? warp - W_E1B : warp;
        const int row = quad * 32 + lane;
        const bool usb = s.row_usb()[row] != 0;
        const uint32_t tmem_row = tmem + ((uint32_t)(quad * 32) << 16);
        const int ch1 = s.row_ch()[row];
        const bool sam_tile = WITH_SAM && dmode == 2;
        const int it0 = sam_tile ? (second ? 4 : 0) : (second ? 2 : 0), it1 = sam_tile ? 4 : (second ? 4 : 2);
        SamState sam{0.f, 0.f, 0.f};
        if (sam_tile && !second && ch1 >= 0) {
            const float4 st = *reinterpret_cast<const float4 *>(a.sam_state + (size_t)ch1 * 4);
            sam.phi = st.x; sam.omega = st.y; sam.dc = st.z;
        }
        TCP_BEGIN;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
            mbar_wait(bar(B_M1_DONE + (c & 1)), (c >> 1) & 1);
            if (warp == 0) TCP(8);
            // D slice (c+4)%6 was the oldest slice of the band-pass MMAs of chunk c-2
            if (c >= 2) mbar_wait(bar(B_M2_DONE + (c & 1)), ((c - 2) >> 1) & 1);
            if (warp == 0) TCP(9);
            tc_fence_after();
            const uint32_t acc1 = tmem_row + (RDSP_TC_A_TMEM ? TM_ACC1B : (c & 1) * TM_ACC1B);
            if (dmode == 1) epilogue1<1>(s, acc1, row, (c + 4) % SLICES, usb, sam, it0, it1);
            else if (WITH_SAM && dmode == 2) epilogue1<2>(s, acc1, row, (c + 4) % SLICES, usb, sam, it0, it1);
            else epilogue1<0>(s, acc1, row, (c + 4) % SLICES, usb, sam, it0, it1);
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(bar(B_E1_DONE + (c & 1)));
            if (warp == 0) TCP(10);
        }
        if (sam_tile && !second && ch1 >= 0) *reinterpret_cast<float4 *>(a.sam_state + (size_t)ch1 * 4) = make_float4(sam.phi, sam.omega, sam.dc, 0.f);
    } else {
        // ===== epilogue 2 =====
        const int row = threadIdx.x - W_E2 * 32;
        const int ch = s.row_ch()[row];
        const uint32_t tmem_row = tmem + ((uint32_t)((warp - W_E2) * 32) << 16) + TM_ACC2P;
        TCP_BEGIN;
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
            const int t = t0 + (c >> 2);
            mbar_wait(bar(B_M2_DONE + (c & 1)), (c >> 1) & 1);
            if (warp == W_E2) TCP(11);
            tc_fence_after();
            epilogue2(a, tmem_row, t >= t_store ? ch : -1, t, c & 3, bar(B_E2_DONE));
            if (warp == W_E2) TCP(12);
        }
    }

    // ---- state for the next call: the last 128 samples of each line = slices nch .. nch+3 ----
    tc_fence_before();
    __syncthreads();
    TCQ(14);
    if (seg == n_seg - 1) state_io<true>(s, a.hist_out, nch % SLICES, warp, lane);
    if (warp == W_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TM_COLS) : "memory");
    }
    TCQ(15);
#ifdef RDSP_TC_PROF
    if (threadIdx.x == 0) {
        unsigned long long gt1_;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1_));
        const int id = blockIdx.y * gridDim.x + blockIdx.x;
        if (id < 4096) { g_tc_cta[id][0] = gt0_; g_tc_cta[id][1] = gt1_; }
    }
#endif
}

}  // namespace

// ---- host side: Toeplitz images and the tile table ---------------------------------------------------------------

// every tap of the row has signed digits (hh = (h - (int8)h) / 256 fits int8): a paired image exists
static bool row_pairable(const int16_t *row)
{
    for (int k = 0; k < RDSP_NTAPS; k++) if (row[k] >= 32640) return false;
    return true;
}

// Two images per tap row, [15][2][TOEP_SET_B]:
//   image 0 (classic)  [plane: low byte (unsigned) | high byte (signed)][10 k-groups][32 outputs][16 B]
//   image 1 (paired)   [10 k-groups][64 columns: signed low digit of output n | high digit of output n][16 B]
void front_tc_build_toeplitz(const int16_t *taps /*[15][stride]*/, int stride, uint8_t *out /*[15][2][TOEP_SET_B]*/)
{
    for (int r = 0; r < 15; r++) {
        uint8_t *classic = out + (size_t)(2 * r) * TOEP_SET_B, *pairimg = out + (size_t)(2 * r + 1) * TOEP_SET_B;
        const bool ok = row_pairable(taps + (size_t)r * stride);
        for (int kg = 0; kg < 10; kg++)
            for (int n = 0; n < NOUT; n++)
                for (int b = 0; b < 16; b++) {
                    const int k = kg * 16 + b, tap = 128 + n - k;
                    const int h = (tap >= 0 && tap <= 128) ? taps[(size_t)r * stride + tap] : 0;
                    classic[0 * TOEP_PLANE_B + (kg * NOUT + n) * 16 + b] = (uint8_t)((uint16_t)h & 0xFF);
                    classic[1 * TOEP_PLANE_B + (kg * NOUT + n) * 16 + b] = (uint8_t)((uint16_t)h >> 8);
                    const int hl = (int8_t)((uint16_t)h & 0xFF), hh = (h - hl) / 256;        // h = 256 hh + hl, hl in [-128, 127]
                    pairimg[(kg * 2 * NOUT + n) * 16 + b] = ok ? (uint8_t)(int8_t)hl : 0;
                    pairimg[(kg * 2 * NOUT + NOUT + n) * 16 + b] = ok ? (uint8_t)(int8_t)hh : 0;
                }
    }
}

// Channels that share their three tap rows (by content) and the detector (sideband sum / envelope / SAM) form a class;
// classes are cut into tiles of 128.
int front_tc_build_tiles(const RdspChanParams *par, int C, const int16_t *taps, int stride,
                         std::vector<int> &tile_ch, std::vector<int4> &tile_rows)
{
    int canon[15];
    for (int r = 0; r < 15; r++) {
        canon[r] = r;
        for (int q = 0; q < r; q++)
            if (!memcmp(taps + (size_t)q * stride, taps + (size_t)r * stride, RDSP_NTAPS * sizeof(int16_t))) { canon[r] = q; break; }
    }
    std::map<std::array<int, 4>, std::vector<int>> classes;
    for (int ch = 0; ch < C; ch++) {
        const RdspChanParams &p = par[ch];
        const int td = p.demod == RDSP_DEMOD_SAM_ ? RDSP_DEMOD_AM_ : p.demod;      // SAM filters its arms with the AM rows
        const int kind = p.demod == RDSP_DEMOD_AM_ ? 1 : (p.demod == RDSP_DEMOD_SAM_ ? 2 : 0);
        const std::array<int, 4> key = {canon[td], canon[RDSP_N_DEMOD + td], canon[2 * RDSP_N_DEMOD + p.filter], kind};
        classes[key].push_back(ch);
    }
    tile_ch.clear();
    tile_rows.clear();
    for (auto &kv : classes) {
        const std::vector<int> &v = kv.second;
        for (size_t i = 0; i < v.size(); i += ROWS) {
            // w = detector kind | paired-image flag << 8 (all three tap rows must have one; RDSP_FRONT_PAIRED=0 forces the classic path)
            static const bool allow = [] { const char *e = getenv("RDSP_FRONT_PAIRED"); return !(e && e[0] == '0'); }();
            const bool pr = allow && row_pairable(taps + (size_t)kv.first[0] * stride) && row_pairable(taps + (size_t)kv.first[1] * stride) &&
                            row_pairable(taps + (size_t)kv.first[2] * stride);
            tile_rows.push_back(make_int4(kv.first[0], kv.first[1], kv.first[2], kv.first[3] | (pr ? 256 : 0)));
            for (int r = 0; r < ROWS; r++) tile_ch.push_back(i + r < v.size() ? v[i + r] : -1);
        }
    }
    return (int)tile_rows.size();
}

#ifdef RDSP_TC_PROF
void front_tc_read_prof(unsigned long long *out, bool reset)
{
    cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(g_tc_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z)); }
}
void front_tc_read_cta(unsigned long long *out) { cudaMemcpyFromSymbol(out, g_tc_cta, sizeof(g_tc_cta)); }
#endif

size_t front_tc_toeplitz_bytes() { return (size_t)15 * 2 * TOEP_SET_B; }

// The Toeplitz table as a 2-D byte tensor [15 * 2 * 40 rows][256 bytes]; one box = 40 rows = one image of one tap row,
// dense in shared memory (no swizzle, no interleave) = the layout the UMMA descriptors of issue_fir expect.
int front_tc_make_tensor_map(const uint8_t *d_toep, CUtensorMap *map)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)TOEP_ROW_B, (cuuint64_t)15 * 2 * TOEP_ROWS_PER_SET};
    const cuuint64_t strides[1] = {(cuuint64_t)TOEP_ROW_B};
    const cuuint32_t box[2] = {(cuuint32_t)TOEP_ROW_B, (cuuint32_t)TOEP_ROWS_PER_SET};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((EncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_toep), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// Segments: with fewer tiles than SMs the T blocks of a tile are cut into up to 8 time segments, one CTA each.  A later
// segment pays about 1.25 blocks of warm-up (one block of loads, one of Hilbert MMAs), so the cut points balance
// n_0 against n_s + 1.25.  S = wanted number of segments; fewer are planned when T is too short.
void front_tc_plan_segments(int S, int T, int *bounds, int *n_seg)
{
    if (S > 8) S = 8;
    while (S > 1) {
        // segment s > 0 must start at block >= 2 and hold at least one block
        const double target = (T + 1.25 * (S - 1)) / S;
        int n0 = (int)(target + 0.5);
        if (n0 < 2) n0 = 2;
        if (T - n0 >= S - 1 && target - 1.25 >= 0.5) {
            bounds[0] = 0; bounds[1] = n0;
            int left = T - n0;
            for (int k = 1; k < S; k++) {
                int n = (left + (S - k) - 1) / (S - k);
                bounds[k + 1] = bounds[k] + n;
                left -= n;
            }
            break;
        }
        S--;
    }
    if (S <= 1) { S = 1; bounds[0] = 0; bounds[1] = T; }
    *n_seg = S;
}

// How many segments for the sideband tiles and for the AM tiles?  An AM tile's chunk takes about 1.5 x as long (its epilogue 1
// computes the integer square root of the envelope; phase timers), so with one segment count for all the AM tiles finish last
// (cfg5: 48 + 16 tiles x 2 segments = 128 CTAs, 20 SMs idle, and the 16 AM tiles set the kernel's time).  Pick the pair that
// minimises the slowest CTA while everything still runs as one wave.
static void front_tc_choose_segments(int n_sum, int n_am, int T, int n_sm, int *s_sum, int *s_am)
{
    auto cost = [&](double w, int n, int S) { return n ? w * (T + 1.25 * (S - 1)) / S : 0.0; };
    double best = 1e30;
    *s_sum = 1; *s_am = 1;
    for (int a = 1; a <= 8; a++)
        for (int b = 1; b <= 8; b++) {
            if (n_sum * a + n_am * b > n_sm && !(a == 1 && b == 1)) continue;
            if ((!n_sum && a > 1) || (!n_am && b > 1)) continue;
            const double c = std::max(cost(1.0, n_sum, a), cost(1.5, n_am, b)) + 1e-3 * (a + b);   // ties: fewer CTAs
            if (c < best) { best = c; *s_sum = a; *s_am = b; }
        }
}

void launch_front_tc(const FrontArgs &a_in, const FrontTcTables &tb_in, cudaStream_t st)
{
    // the opt-in to > 48 KB of dynamic shared memory is per (kernel, device): a process with one handle per GPU needs it on each
    static std::atomic<int> n_sm_dev[RDSP_MAX_DEVICES];
    const int dev = rdsp_current_device();
    int n_sm = n_sm_dev[dev].load(std::memory_order_acquire);
    if (!n_sm) {
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_front_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_B);
        cudaFuncSetAttribute(k_front_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_B);
        n_sm_dev[dev].store(n_sm, std::memory_order_release);
    }
    FrontArgs a = a_in;
    FrontTcTables tb = tb_in;
    if (!a.hist_out) a.hist_out = a.hist;
    int want[2] = {1, 1};                                                   // per detector kind: 0 sideband sum (and SAM), 1 AM envelope
    if (a.hist_out != a.hist && !tb.any_sam)                                // in-place state / SAM loop / noise blanker: one segment
        front_tc_choose_segments(tb.n_tiles - tb.n_am_tiles, tb.n_am_tiles, a.T, n_sm, &want[0], &want[1]);
    int S = 1;
    for (int k = 0; k < 2; k++) {
        front_tc_plan_segments(want[k], a.T, tb.seg_bounds[k], &tb.n_seg[k]);
        if (tb.n_seg[k] > S && (k == 0 ? tb.n_tiles > tb.n_am_tiles : tb.n_am_tiles > 0)) S = tb.n_seg[k];
    }
    if (tb.sam_tiles) k_front_tc<true><<<dim3(tb.n_tiles, S), NTHREADS, SMEM_B, st>>>(a, tb);
    else k_front_tc<false><<<dim3(tb.n_tiles, S), NTHREADS, SMEM_B, st>>>(a, tb);
}
