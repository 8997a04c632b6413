// front_tc_check.cu — development harness: the tcgen05 front end (k_front_tc.cu) against the CUDA-core front end
// (k_front.cu, itself bit-exact against the oracle in tests/test_gpu_parity.py) on random channels, taps and history.
// Build: make -C tools front_tc_check     Run (GPU box): tools/front_tc_check [C] [T] [reps]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../radiodsp_sdr_rx_b200/csrc/kernels.h"

#ifdef RDSP_TC_PROF
void front_tc_read_prof(unsigned long long *out, bool reset);
void front_tc_read_cta(unsigned long long *out);
#endif
#define CKX(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static uint64_t rng_s = 0x5D5DB200ull;
static uint32_t rnd() { rng_s += 0x9E3779B97F4A7C15ull; uint64_t z = rng_s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return (uint32_t)((z ^ (z >> 31)) >> 16); }

int main(int argc, char **argv)
{
    const int C = argc > 1 ? atoi(argv[1]) : 300;
    const int T = argc > 2 ? atoi(argv[2]) : 3;
    const int reps = argc > 3 ? atoi(argv[3]) : 0;
    const int extreme = argc > 4 ? atoi(argv[4]) : 0;
    const int modes = argc > 5 ? atoi(argv[5]) : 0;      // 0: random demod / filter, 1: no AM, 2: one class, 3: all AM

    // taps: random q15 rows, some large so that the 32-bit accumulator wraps; rows 0 and 1 identical (LSB / USB)
    std::vector<int16_t> taps16(15 * RDSP_NTAPS);
    for (int r = 0; r < 15; r++)
        for (int k = 0; k < RDSP_NTAPS; k++) {
            int v = (int)(rnd() % 65536) - 32768;
            if (!extreme) v = (r % 3 == 0) ? v / 64 : (r % 3 == 1 ? v / 8 : v);
            taps16[r * RDSP_NTAPS + k] = (int16_t)v;
        }
    memcpy(&taps16[1 * RDSP_NTAPS], &taps16[0], RDSP_NTAPS * 2);
    memcpy(&taps16[6 * RDSP_NTAPS], &taps16[5 * RDSP_NTAPS], RDSP_NTAPS * 2);
    std::vector<int32_t> taps32(15 * RDSP_TAPS_PAD, 0);
    for (int r = 0; r < 15; r++) for (int k = 0; k < RDSP_NTAPS; k++) taps32[r * RDSP_TAPS_PAD + k] = taps16[r * RDSP_NTAPS + k];

    std::vector<RdspChanParams> par(C);
    for (int c = 0; c < C; c++) {
        memset(&par[c], 0, sizeof(RdspChanParams));
        par[c].demod = rnd() % 5;
        par[c].filter = rnd() % 5;
        if (modes == 1) par[c].demod %= 4;
        if (modes == 2) { par[c].demod = c & 1; par[c].filter = 2; }
        if (modes == 3) { par[c].demod = 4; par[c].filter = 4; }
        par[c].mult_i = (rnd() % 4 == 0) ? 65536 : (int32_t)(rnd() % 200000);
        par[c].mult_q = (rnd() % 4 == 0) ? 65536 : (int32_t)(rnd() % 200000) - 50000;
    }
    const size_t n_iq = (size_t)T * C * RDSP_BLK * 2;
    std::vector<int16_t> iq(n_iq), hist((size_t)C * 3 * RDSP_BLK);
    for (auto &v : iq) v = (int16_t)((int)(rnd() % 65536) - 32768);
    for (auto &v : hist) v = (int16_t)((int)(rnd() % 65536) - 32768);
    if (extreme) for (size_t i = 0; i < n_iq; i += 7) iq[i] = (i & 8) ? 32767 : -32768;

    std::vector<int> tile_ch; std::vector<int4> tile_rows;
    const int n_tiles = front_tc_build_tiles(par.data(), C, taps16.data(), RDSP_NTAPS, tile_ch, tile_rows);
    std::vector<uint8_t> toep(front_tc_toeplitz_bytes());
    front_tc_build_toeplitz(taps16.data(), RDSP_NTAPS, toep.data());
    printf("C=%d T=%d tiles=%d\n", C, T, n_tiles);

    int16_t *d_iq, *d_hist_a, *d_hist_b, *d_hist_c, *d_mono_a, *d_mono_b, *d_st_a, *d_st_b; int32_t *d_taps; RdspChanParams *d_par;
    int *d_tile_ch; int4 *d_tile_rows; uint8_t *d_toep;
    CKX(cudaMalloc(&d_iq, n_iq * 2)); CKX(cudaMalloc(&d_hist_a, hist.size() * 2)); CKX(cudaMalloc(&d_hist_b, hist.size() * 2)); CKX(cudaMalloc(&d_hist_c, hist.size() * 2));
    CKX(cudaMalloc(&d_mono_a, n_iq)); CKX(cudaMalloc(&d_mono_b, n_iq)); CKX(cudaMalloc(&d_st_a, n_iq * 2)); CKX(cudaMalloc(&d_st_b, n_iq * 2));
    CKX(cudaMalloc(&d_taps, taps32.size() * 4)); CKX(cudaMalloc(&d_par, C * sizeof(RdspChanParams)));
    CKX(cudaMalloc(&d_tile_ch, tile_ch.size() * 4)); CKX(cudaMalloc(&d_tile_rows, tile_rows.size() * sizeof(int4))); CKX(cudaMalloc(&d_toep, toep.size()));
    CKX(cudaMemcpy(d_iq, iq.data(), n_iq * 2, cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_hist_a, hist.data(), hist.size() * 2, cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_hist_b, hist.data(), hist.size() * 2, cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_taps, taps32.data(), taps32.size() * 4, cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_par, par.data(), C * sizeof(RdspChanParams), cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_tile_ch, tile_ch.data(), tile_ch.size() * 4, cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_tile_rows, tile_rows.data(), tile_rows.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CKX(cudaMemcpy(d_toep, toep.data(), toep.size(), cudaMemcpyHostToDevice));
    CKX(cudaMemset(d_mono_a, 0x11, n_iq)); CKX(cudaMemset(d_mono_b, 0x22, n_iq)); CKX(cudaMemset(d_st_a, 0x11, n_iq * 2)); CKX(cudaMemset(d_st_b, 0x22, n_iq * 2));

    FrontArgs a{};
    a.iq = d_iq; a.par = d_par; a.taps = d_taps; a.C = C; a.T = T;
    FrontTcTables tb{};
    tb.tile_ch = d_tile_ch; tb.tile_rows = d_tile_rows; tb.toep = d_toep; tb.n_tiles = n_tiles;
    if (front_tc_make_tensor_map(d_toep, &tb.toep_map) != 0) { printf("tensor map failed\n"); return 1; }
    for (const int4 &r : tile_rows) if ((r.w & 0xFF) == 1) tb.n_am_tiles++;

    int bad = 0;
    for (int pass = 0; pass < 2; pass++) {          // pass 0: mono output, pass 1: stereo output (state carries over)
        FrontArgs x = a, y = a;
        x.hist = d_hist_a; y.hist = pass == 0 ? d_hist_b : d_hist_c; y.hist_out = pass == 0 ? d_hist_c : d_hist_b;
        if (pass == 0) { x.out_mono = d_mono_a; y.out_mono = d_mono_b; } else { x.out_stereo = d_st_a; y.out_stereo = d_st_b; }
        launch_front(x, 0);
        CKX(cudaDeviceSynchronize());
        launch_front_tc(y, tb, 0);
        CKX(cudaDeviceSynchronize());
        const size_t n_out = pass == 0 ? n_iq / 2 : n_iq;
        std::vector<int16_t> oa(n_out), ob(n_out), ha(hist.size()), hb(hist.size());
        CKX(cudaMemcpy(oa.data(), pass == 0 ? d_mono_a : d_st_a, n_out * 2, cudaMemcpyDeviceToHost));
        CKX(cudaMemcpy(ob.data(), pass == 0 ? d_mono_b : d_st_b, n_out * 2, cudaMemcpyDeviceToHost));
        CKX(cudaMemcpy(ha.data(), d_hist_a, hist.size() * 2, cudaMemcpyDeviceToHost));
        CKX(cudaMemcpy(hb.data(), y.hist_out, hist.size() * 2, cudaMemcpyDeviceToHost));
        size_t nd = 0, first = (size_t)-1;
        for (size_t i = 0; i < n_out; i++) if (oa[i] != ob[i]) { if (first == (size_t)-1) first = i; nd++; }
        size_t nh = 0, firsth = (size_t)-1;
        for (size_t i = 0; i < hist.size(); i++) if (ha[i] != hb[i]) { if (firsth == (size_t)-1) firsth = i; nh++; }
        printf("pass %d: output mismatches %zu / %zu, state mismatches %zu / %zu\n", pass, nd, n_out, nh, hist.size());
        if (nd) {
            const size_t per = pass == 0 ? RDSP_BLK : 2 * RDSP_BLK;
            const size_t cbk = first / per;
            printf("  first output mismatch at %zu: block %zu channel %zu sample %zu: ref %d tc %d (demod %d filter %d)\n", first, cbk / C, cbk % C,
                   first % per, oa[first], ob[first], par[cbk % C].demod, par[cbk % C].filter);
            for (size_t i = first; i < first + 8 && i < n_out; i++) printf("    [%zu] ref %6d tc %6d\n", i, oa[i], ob[i]);
        }
        if (nh) printf("  first state mismatch at %zu: channel %zu line %zu sample %zu: ref %d tc %d\n", firsth, firsth / 384, (firsth / 128) % 3, firsth % 128, ha[firsth], hb[firsth]);
        bad += (nd != 0) + (nh != 0);
    }

    if (reps > 0) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        FrontArgs x = a; x.hist = d_hist_a; x.out_mono = d_mono_a;
        FrontArgs y = x; y.hist_out = d_hist_c;
        float ms_ref = 0.f, ms_tc = 0.f;
        for (int w = 0; w < 2; w++) {
            cudaEventRecord(e0); for (int r = 0; r < reps; r++) launch_front(x, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_ref, e0, e1);
            cudaEventRecord(e0); for (int r = 0; r < reps; r++) launch_front_tc(y, tb, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_tc, e0, e1);
        }
        CKX(cudaDeviceSynchronize());
        printf("timing (L2-warm, back to back): cuda-core %.3f us / launch, tcgen05 %.3f us / launch  (%.2fx)\n", 1e3 * ms_ref / reps, 1e3 * ms_tc / reps, ms_ref / ms_tc);
    }
#ifdef RDSP_TC_PROF
    {
        unsigned long long pr[16];
        front_tc_read_prof(pr, true);
        FrontArgs y = a; y.hist = d_hist_a; y.out_mono = d_mono_a; y.hist_out = d_hist_c;
        launch_front_tc(y, tb, 0);
        CKX(cudaDeviceSynchronize());
        front_tc_read_prof(pr, true);
        const char *nm[16] = {"LD wait m1_done", "LD split+fetch", "MMA wait in_full", "MMA wait e1(c-2)", "MMA issue M1", "MMA wait e1(c-1)", "MMA wait e2",
                              "MMA issue M2", "E1 wait m1_done", "E1 wait m2_done", "E1 work", "E2 wait m2_done", "E2 work", "prologue (thread 0)", "main loop (thread 0)", "state store"};
        {
            std::vector<unsigned long long> ct(4096 * 2);
            front_tc_read_cta(ct.data());
            unsigned long long t0 = ~0ull, t1 = 0;
            for (int i = 0; i < n_tiles * 8 && i < 4096; i++) if (ct[2 * i + 1]) { if (ct[2 * i] < t0) t0 = ct[2 * i]; if (ct[2 * i + 1] > t1) t1 = ct[2 * i + 1]; }
            printf("  kernel span %.1f us; per CTA (tile, seg): start, duration us\n", (t1 - t0) * 1e-3);
            for (int i = 0; i < n_tiles * 8 && i < 4096; i++)
                if (ct[2 * i + 1] && (i % 16 == 0 || ct[2 * i + 1] == t1)) printf("    cta %4d (tile %3d seg %d): start %7.1f dur %7.1f\n", i, i % n_tiles, i / n_tiles, (ct[2 * i] - t0) * 1e-3, (ct[2 * i + 1] - ct[2 * i]) * 1e-3);
        }
        for (int i = 0; i < 16; i++) printf("  %-18s %10.0f clk total (CTA 0)\n", nm[i], (double)pr[i]);
    }
#endif
    printf(bad ? "FAIL\n" : "PASS\n");
    return bad ? 1 : 0;
}
