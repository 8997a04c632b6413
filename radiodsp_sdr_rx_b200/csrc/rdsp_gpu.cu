// rdsp_gpu.cu — the C ABI of include/rdsp_gpu.h: handle, per-channel state in HBM, parameter
// bookkeeping and the per-tick kernel pipeline.  Host code only; the kernels are in k_*.cu.
//
// Pipeline of one rdsp_gpu_process_blocks(T) call (reference graph order, RadioDSP_SDR_RX.ino:71-89):
//   spectrum path : k_spec256 (HP biquads + 256-pt IQ spectrum)                       on raw IQ
//   audio path    : k_front (K0-K2) -> k_nlms<notch> (K3, listed channels) -> k_agc (K4)
//                   -> k_fftfilt (K5/K8/K7) -> k_nlms<dnr> (K6, listed channels) -> k_spec1024 (K10)
// Intermediates between kernels are int16 mono / f32 rows in handle-owned scratch (L2 resident at
// the batch sizes of BASELINE.json).  Every launch covers all T blocks of the call, so per-channel state makes
// one HBM round trip per call.  Behind the front end the call forks into up to three streams (the channels whose
// notch runs, the channels that bypass it, the spectrum branch) joined by events on the handle's stream.
#include "../../include/rdsp_gpu.h"
#include "host_design.h"
#include "kernels.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#define RDSP_VERSION "rdsp-b200 0.2 (sm_100a)"

namespace {

std::string g_create_error;

enum KernelKind { KK_FRONT = 0, KK_NOTCH, KK_AGC, KK_FFTFILT, KK_DNR, KK_BIQUAD, KK_SPEC256, KK_SPEC1024, KK_PAN, KK_COUNT };
const char *const kKernelNames[KK_COUNT] = {"k_front", "k_nlms_notch", "k_agc", "k_fftfilt", "k_nlms_dnr",
                                            "k_biquad", "k_spec256", "k_spec1024", "k_panadapter"};

struct ProfRec { int kind; cudaEvent_t e0, e1; };

// one captured call shape (run_call)
struct GraphEntry {
    int T; const void *iq; void *audio; int fe_cur, par; cudaStream_t st;
    int frame;                                 // one-block calls: does this tick complete an audio-spectrum frame (else 0)
    cudaGraphExec_t exec;                      // nullptr: seen once, not captured yet
    uint64_t launches;                         // kernels in the graph
    unsigned long long last_use;
};

constexpr int kMaxGroups = 8;
constexpr int kStreams = 3 * kMaxGroups + 1;   // per channel group: main chain, spectrum branch, side branch; + the front end

}  // namespace

struct rdsp_gpu {
    rdsp_gpu_config_t cfg;
    int C = 0, maxT = 1;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // one stream per stage of the graph + one event per (stage, chunk): the wavefront of process_blocks
    cudaStream_t stage_stream[kStreams] = {nullptr};
    cudaEvent_t ev_group[kMaxGroups][3] = {{nullptr}};
    cudaEvent_t ev_mark[kMaxGroups][4] = {{nullptr}};      // after the notch / AGC / FFT filter / DNR of the chain that runs the notch
    cudaEvent_t ev_fork = nullptr, ev_front = nullptr;
    // IO_HOST: copies run on their own streams over double-buffered staging, so that the H2D of call n+1 and the
    // D2H of call n-1 overlap the kernels of call n (async handles)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    // host I/O: a ring of n_stage device staging buffers (H2D of call n+1, kernels of call n, D2H of call n-1 overlap;
    // RDSP_HOST_STAGES = 2..4.  Measured end to end, cfg5: with 8 blocks per call 11.5 - 11.7 GS/s for 2, 3 and 4 alike (a
    // call's copies and kernels are 0.7 / 0.4 ms: two buffers overlap them), so two it is there; with ONE block per call the
    // three phases of a call take ~ 90 / 100 / 90 us and two buffers leave them waiting for each other: 2 / 3 / 4 buffers =
    // 7.6 - 8.4 / 9.1 - 9.7 / 9.8 GS/s (0.84 / 0.92 - 0.97 / 0.97 of the copy ceiling measured in the same run), so handles
    // for short calls (max_blocks_per_call <= 4) get four)
    static constexpr int kMaxStage = 4;
    int n_stage = 2;
    cudaEvent_t ev_h2d[kMaxStage] = {}, ev_comp[kMaxStage] = {}, ev_d2h[kMaxStage] = {};
    int16_t *d_in_stage2[kMaxStage] = {}, *d_out_stage2[kMaxStage] = {};
    unsigned long long host_calls = 0;
    std::string err;

    // parameters
    std::vector<rdsp_chan_params_t> par;       // as set by the caller
    std::vector<RdspChanParams> dpar;          // as the kernels see them
    std::vector<int> dnr_old_level, notch_old_level;
    bool par_dirty = true;
    RdspChanParams *d_par = nullptr;
    int *d_list_notch = nullptr, *d_list_plain = nullptr, *d_list_dnr = nullptr;
    int *d_list_dnr_p = nullptr, *d_list_dnr_n = nullptr;      // DNR channels that bypass / run the notch
    int n_notch = 0, n_plain = 0, n_dnr = 0;
    std::vector<int> l_notch, l_plain, l_dnr, l_dnr_p, l_dnr_n;   // host copies (ascending): launches take sub-ranges

    // coefficient tables
    int16_t taps[15][RDSP_FIR_TAPS];
    bool taps_dirty = true;
    int32_t *d_taps = nullptr;
    std::map<std::pair<float, float>, int> mask_ids;     // masks designed from (pbt_lo, pbt_hi)
    std::map<std::vector<float>, int> custom_mask_ids;   // masks installed with rdsp_gpu_set_mask, by content
    std::vector<int> custom_mask;              // [C] id of the installed mask, -1 = the designed one; survives set_mode
                                               // until the channel's pbt cut-offs change (reInitializeFilter redesigns)
    std::vector<float> masks;                  // [n][512]
    int masks_uploaded = 0, mask_cap = 0;
    float2 *d_masks = nullptr;
    int2 *d_tw = nullptr;
    int2 tw3_256[4][3] = {};
    int16_t *d_win256 = nullptr, *d_win1024 = nullptr;
    float2 *d_tw256 = nullptr;
    float *d_sin512 = nullptr;
    int32_t bq[5];
    float agc_alpha_a = 0.f, agc_alpha_d[4] = {0.f, 0.f, 0.f, 0.f};

    // per-channel state
    int16_t *d_fe_hist = nullptr;              // delay lines of the front end; k_front_tc ping-pongs between the two
    int16_t *d_fe_hist2 = nullptr;             // buffers so that a call's blocks can run as concurrent time segments
    int fe_hist_cur = 0;
    bool front_tc = true;                      // RDSP_FRONT_IMPL=cuda-core selects k_front.cu (cross-check)
    int pdl_max_T = 2;                         // calls of up to this many blocks chain their kernels by programmatic dependent launch (RDSP_PDL_MAX_T)
    int spec_after = 1;                        // where the spectrum branch starts (enqueue_call); RDSP_SPEC_AFTER=0..5, experiments
    bool nlms_direct = false;                  // RDSP_NLMS_IMPL=direct selects k_nlms_direct.cu (cross-check); read at create
    uint8_t *d_toep = nullptr;                 // Toeplitz byte planes of the 15 tap rows
    CUtensorMap toep_map;                      // TMA descriptor of d_toep (k_front_tc loads its three images with it)
    float *d_sam_state = nullptr;              // [C][4] SAM carrier loop
    int32_t *d_nb_ref = nullptr;               // [C] noise blanker running magnitude
    int any_sam = 0, sam_tiles = 0, n_am_tiles = 0;
    int *d_tile_ch = nullptr; int4 *d_tile_rows = nullptr; int n_tiles = 0, tile_cap = 0;
    float *d_nc_coeff = nullptr, *d_nc_prev = nullptr, *d_nc_energy = nullptr; uint8_t *d_nc_first = nullptr;
    float *d_dn_coeff = nullptr, *d_dn_prev = nullptr, *d_dn_energy = nullptr; uint8_t *d_dn_first = nullptr;
    float *d_agc_env = nullptr;
    int16_t *d_conv_last = nullptr; float *d_nfloor = nullptr;
    int32_t *d_bq_state = nullptr; int16_t *d_spec_prev = nullptr; uint32_t *d_spec_sum = nullptr; uint16_t *d_spec_out = nullptr;
    int16_t *d_ring = nullptr; uint16_t *d_spec1024_out = nullptr;
    uint16_t *d_view = nullptr; float *d_smeter = nullptr;
    uint16_t *d_waterfall = nullptr; int *d_wf_head = nullptr; std::vector<int> wf_head;   // [C][50][128] ring + newest slot
    uint16_t *d_wf_rows = nullptr; uint8_t *d_wf_col = nullptr; size_t wf_scratch_ch = 0;  // read-out scratch, grown on demand

    // scratch
    int16_t *d_mid_a = nullptr, *d_mid_b = nullptr;
    float *d_scr = nullptr, *d_dbg = nullptr;
    int16_t *d_hp_iq = nullptr;                // high-passed IQ between k_biquad and k_spec256

    // tick bookkeeping (uniform over channels): the kernels read it from the device copy d_tick[tick_par] and write the
    // advanced values into d_tick[tick_par ^ 1]; the host keeps a mirror for the ready flags
    int spec_have_prev = 0, spec_count = 0;
    unsigned long long tick = 0;
    RdspTick *d_tick = nullptr;
    int tick_par = 0;
    // CUDA graphs of the call shapes seen so far (run_call); dropped whenever tables, lists or tiles change
    static constexpr size_t kMaxGraphs = 32;
    std::vector<GraphEntry> graphs;
    bool use_graphs = true;
    unsigned long long graph_clock = 0, graph_replays = 0;
    std::vector<uint8_t> spec_ready, spec1024_ready;

    // instrumentation
    uint64_t launches = 0;
    bool profiling = false;
    bool timeline = false;                     // RDSP_TIMELINE=1: event-bracket every launch IN the wavefront and print when each ran
    std::vector<ProfRec> prof_pending;
    double prof_ms[KK_COUNT] = {0};
    uint64_t prof_n[KK_COUNT] = {0};
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            h->err = b_;                                                                           \
            return RDSP_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

template <typename T>
cudaError_t dalloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemset(*p, 0, n * sizeof(T));
}

bool has(const rdsp_gpu *h, uint32_t st) { return (h->cfg.stage_mask & st) != 0; }

int validate_params(const rdsp_chan_params_t *p, std::string &why)
{
    if (p->demod < 0 || p->demod >= RDSP_DEMOD_MODES) { why = "demod out of range"; return RDSP_ERR_RANGE; }
    if (p->als_peak < 0 || p->als_peak > 1) { why = "als_peak must be 0 or 1"; return RDSP_ERR_RANGE; }
    if (p->nb_on < 0 || p->nb_on > 1) { why = "nb_on must be 0 or 1"; return RDSP_ERR_RANGE; }
    if (!(p->nb_threshold_db >= 0.0f && p->nb_threshold_db <= 40.0f)) { why = "nb_threshold_db out of range (0 .. 40)"; return RDSP_ERR_RANGE; }
    if (p->audio_filter < 0 || p->audio_filter >= RDSP_FILTER_COUNT) { why = "audio_filter out of range"; return RDSP_ERR_RANGE; }
    if (p->agc_mode < 0 || p->agc_mode >= RDSP_AGC_COUNT) { why = "agc_mode out of range"; return RDSP_ERR_RANGE; }
    if (p->notch_on < 0 || p->notch_on > 1) { why = "notch_on must be 0 or 1"; return RDSP_ERR_RANGE; }
    if (p->notch_level < 0 || p->notch_level > 100) { why = "notch_level out of range"; return RDSP_ERR_RANGE; }
    if (p->nr_kind < RDSP_NR_OFF || p->nr_kind > RDSP_NR_SPECTRAL) { why = "nr_kind out of range"; return RDSP_ERR_RANGE; }
    if (p->nr_level < 0 || p->nr_level > 100) { why = "nr_level out of range"; return RDSP_ERR_RANGE; }
    if (!(p->pbt_lo_hz >= -22050.0f && p->pbt_hi_hz <= 22050.0f && p->pbt_hi_hz > p->pbt_lo_hz)) { why = "pbt cut-offs invalid"; return RDSP_ERR_RANGE; }
    if (!(p->in_gain >= 0.0f && p->in_gain <= 16.0f)) { why = "in_gain out of range"; return RDSP_ERR_RANGE; }
    if (!(p->out_gain >= 0.0f && p->out_gain <= 16.0f)) { why = "out_gain out of range"; return RDSP_ERR_RANGE; }
    if (!(p->iq_balance > 0.0f && p->iq_balance <= 4.0f)) { why = "iq_balance out of range"; return RDSP_ERR_RANGE; }
    return RDSP_OK;
}

int mask_id_for(rdsp_gpu *h, float lo, float hi)
{
    auto key = std::make_pair(lo, hi);
    auto it = h->mask_ids.find(key);
    if (it != h->mask_ids.end()) return it->second;
    const int id = (int)(h->masks.size() / 512);
    h->masks.resize(h->masks.size() + 512);
    rdsp_host::design_mask((double)lo, (double)hi, &h->masks[(size_t)id * 512]);
    h->mask_ids[key] = id;
    return id;
}

void derive_params(rdsp_gpu *h, int ch)
{
    const rdsp_chan_params_t &p = h->par[ch];
    RdspChanParams &d = h->dpar[ch];
    memset(&d, 0, sizeof(d));
    d.mult_i = (int32_t)((double)p.in_gain * 65536.0);
    d.mult_q = (int32_t)((double)p.in_gain * (double)p.iq_balance * 65536.0);
    d.out_gain = p.out_gain;
    d.mu_notch = rdsp_host::lms_mu(p.notch_level);
    d.mu_dnr = rdsp_host::lms_mu(p.nr_level);
    d.agc_alpha_d = h->agc_alpha_d[p.agc_mode];
    d.nr_spec_level = (float)p.nr_level;
    d.demod = (uint8_t)p.demod;
    d.filter = (uint8_t)p.audio_filter;
    d.agc_mode = (uint8_t)p.agc_mode;
    d.notch_on = (uint8_t)p.notch_on;
    d.als_peak = (uint8_t)p.als_peak;
    d.nb_mult_q8 = p.nb_on ? (uint32_t)(pow(10.0, (double)p.nb_threshold_db / 20.0) * 256.0 + 0.5) : 0u;
    d.nr_kind = (uint8_t)(p.nr_level > 0 ? p.nr_kind : RDSP_NR_OFF);
}

int upload_masks(rdsp_gpu *h)
{
    const int n = (int)(h->masks.size() / 512);
    if (n == h->masks_uploaded) return RDSP_OK;
    if (n > h->mask_cap) {
        int cap = h->mask_cap ? h->mask_cap : 16;
        while (cap < n) cap *= 2;
        CK(cudaStreamSynchronize(h->stream));
        if (h->d_masks) CK(cudaFree(h->d_masks));
        CK(cudaMalloc((void **)&h->d_masks, (size_t)cap * 512 * sizeof(float)));
        h->mask_cap = cap;
        h->masks_uploaded = 0;
    }
    CK(cudaMemcpyAsync(h->d_masks + (size_t)h->masks_uploaded * 256, &h->masks[(size_t)h->masks_uploaded * 512],
                       (size_t)(n - h->masks_uploaded) * 512 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    // the host vector may be reallocated by a later set_mode before the copy has run
    CK(cudaStreamSynchronize(h->stream));
    h->masks_uploaded = n;
    return RDSP_OK;
}

// zero [first, first+count) rows of an array with `row` elements per channel, merging contiguous runs
template <typename T>
cudaError_t zero_rows(T *base, size_t row, const std::vector<int> &chs, cudaStream_t st)
{
    size_t i = 0;
    while (i < chs.size()) {
        size_t j = i + 1;
        while (j < chs.size() && chs[j] == chs[j - 1] + 1) j++;
        cudaError_t e = cudaMemsetAsync(base + (size_t)chs[i] * row, 0, (j - i) * row * sizeof(T), st);
        if (e != cudaSuccess) return e;
        i = j;
    }
    return cudaSuccess;
}

void drop_graphs(rdsp_gpu *h);

// bring the device view of parameters, work lists, taps and masks up to date (start of a process call)
int sync_tables(rdsp_gpu *h)
{
    if (h->taps_dirty || h->par_dirty) drop_graphs(h);     // launch shapes and table pointers may change
    if (h->taps_dirty) {
        std::vector<int32_t> t(15 * RDSP_TAPS_PAD, 0);
        for (int r = 0; r < 15; r++)
            for (int k = 0; k < RDSP_FIR_TAPS; k++) t[(size_t)r * RDSP_TAPS_PAD + k] = h->taps[r][k];
        CK(cudaMemcpyAsync(h->d_taps, t.data(), t.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        std::vector<uint8_t> toep(front_tc_toeplitz_bytes());
        if (h->d_toep) {
            front_tc_build_toeplitz(&h->taps[0][0], RDSP_FIR_TAPS, toep.data());
            CK(cudaMemcpyAsync(h->d_toep, toep.data(), toep.size(), cudaMemcpyHostToDevice, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        h->taps_dirty = false;
        h->par_dirty = true;                   // tiles group channels by tap-row content
    }
    if (!h->par_dirty) return RDSP_OK;

    const bool notch_stage = has(h, RDSP_STAGE_NOTCH), nr_stage = has(h, RDSP_STAGE_NR);
    std::vector<int> l_notch, l_plain, l_dnr, l_dnr_p, l_dnr_n, re_notch, re_dnr;
    for (int ch = 0; ch < h->C; ch++) {
        const rdsp_chan_params_t &p = h->par[ch];
        // Init_LMS_NR on a level change: clears ring/state/energy, keeps the coefficients
        // (RDSP_convolutional.h:327-330, RDSP_noise_reduction.h:35-64)
        if (notch_stage && p.notch_on) {
            l_notch.push_back(ch);
            if (p.notch_level != h->notch_old_level[ch]) { re_notch.push_back(ch); h->notch_old_level[ch] = p.notch_level; }
        } else if (notch_stage) {
            l_plain.push_back(ch);
        }
        if (nr_stage && p.nr_kind == RDSP_NR_LMS && p.nr_level > 0) {
            l_dnr.push_back(ch);
            if (notch_stage && p.notch_on) l_dnr_n.push_back(ch); else l_dnr_p.push_back(ch);
            if (p.nr_level != h->dnr_old_level[ch]) { re_dnr.push_back(ch); h->dnr_old_level[ch] = p.nr_level; }
        }
    }
    if (h->d_nc_prev) {
        CK(zero_rows(h->d_nc_prev, RDSP_BLK, re_notch, h->stream));
        CK(zero_rows(h->d_nc_energy, 1, re_notch, h->stream));
    }
    if (h->d_dn_prev) {
        CK(zero_rows(h->d_dn_prev, RDSP_BLK, re_dnr, h->stream));
        CK(zero_rows(h->d_dn_energy, 1, re_dnr, h->stream));
    }
    int rc = upload_masks(h);
    if (rc != RDSP_OK) return rc;
    CK(cudaMemcpyAsync(h->d_par, h->dpar.data(), (size_t)h->C * sizeof(RdspChanParams), cudaMemcpyHostToDevice, h->stream));
    std::vector<int> tile_ch; std::vector<int4> tile_rows;
    if (has(h, RDSP_STAGE_FRONTEND)) {
        // k_front_tc: channels that share their tap rows, in tiles of 128 MMA rows
        h->n_tiles = front_tc_build_tiles(h->dpar.data(), h->C, &h->taps[0][0], RDSP_FIR_TAPS, tile_ch, tile_rows);
        h->any_sam = 0; h->sam_tiles = 0; h->n_am_tiles = 0;
        for (const int4 &r : tile_rows) if ((r.w & 0xFF) == 2) { h->any_sam = 1; h->sam_tiles = 1; }
        for (const int4 &r : tile_rows) if ((r.w & 0xFF) == 1) h->n_am_tiles++;
        for (int ch = 0; ch < h->C; ch++) if (h->dpar[ch].nb_mult_q8) h->any_sam = 1;   // blanker state is sequential too
        if (h->n_tiles > h->tile_cap) {
            if (h->d_tile_ch) cudaFree(h->d_tile_ch);
            if (h->d_tile_rows) cudaFree(h->d_tile_rows);
            h->d_tile_ch = nullptr; h->d_tile_rows = nullptr;
            h->tile_cap = h->n_tiles + 32;
            CK(cudaMalloc((void **)&h->d_tile_ch, (size_t)h->tile_cap * 128 * sizeof(int)));
            CK(cudaMalloc((void **)&h->d_tile_rows, (size_t)h->tile_cap * sizeof(int4)));
        }
        CK(cudaMemcpyAsync(h->d_tile_ch, tile_ch.data(), tile_ch.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_tile_rows, tile_rows.data(), tile_rows.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    }
    h->n_notch = (int)l_notch.size();
    h->n_plain = (int)l_plain.size();
    h->n_dnr = (int)l_dnr.size();
    if (h->n_plain) CK(cudaMemcpyAsync(h->d_list_plain, l_plain.data(), l_plain.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (h->n_notch) CK(cudaMemcpyAsync(h->d_list_notch, l_notch.data(), l_notch.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (h->n_dnr) CK(cudaMemcpyAsync(h->d_list_dnr, l_dnr.data(), l_dnr.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!l_dnr_p.empty()) CK(cudaMemcpyAsync(h->d_list_dnr_p, l_dnr_p.data(), l_dnr_p.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!l_dnr_n.empty()) CK(cudaMemcpyAsync(h->d_list_dnr_n, l_dnr_n.data(), l_dnr_n.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));      // the host staging vectors go out of scope
    h->l_notch.swap(l_notch); h->l_plain.swap(l_plain); h->l_dnr.swap(l_dnr); h->l_dnr_p.swap(l_dnr_p); h->l_dnr_n.swap(l_dnr_n);
    h->par_dirty = false;
    return RDSP_OK;
}

struct Prof {
    rdsp_gpu *h; int kind; ProfRec r; bool on; cudaStream_t st;
    Prof(rdsp_gpu *h_, int kind_, cudaStream_t st_ = nullptr) : h(h_), kind(kind_), on(h_->profiling || h_->timeline), st(st_ ? st_ : h_->stream) {
        if (on) {
            r.kind = kind;
            cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
            cudaEventRecord(r.e0, st);
        }
    }
    ~Prof() {
        h->launches++;
        if (on) { cudaEventRecord(r.e1, st); h->prof_pending.push_back(r); }
    }
};

void prof_collect(rdsp_gpu *h)
{
    if (h->timeline && !h->prof_pending.empty()) {
        // development aid: start / end of every launch of the last call relative to its fork event, in launch order
        cudaStreamSynchronize(h->stream);
        fprintf(stderr, "[rdsp timeline] kernel            start us   end us\n");
        for (auto &r : h->prof_pending) {
            float a = 0.f, b = 0.f;
            cudaEventSynchronize(r.e1);
            cudaEventElapsedTime(&a, h->ev_fork, r.e0);
            cudaEventElapsedTime(&b, h->ev_fork, r.e1);
            fprintf(stderr, "[rdsp timeline] %-14s %10.1f %8.1f\n", kKernelNames[r.kind], a * 1e3f, b * 1e3f);
        }
    }
    for (auto &r : h->prof_pending) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { h->prof_ms[r.kind] += ms; h->prof_n[r.kind]++; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    h->prof_pending.clear();
}

void free_all(rdsp_gpu *h)
{
    void *ptrs[] = {h->d_fe_hist2, h->d_toep, h->d_tile_ch, h->d_tile_rows, h->d_sam_state, h->d_nb_ref,
                    h->d_par, h->d_tick, h->d_list_notch, h->d_list_plain, h->d_list_dnr, h->d_list_dnr_p, h->d_list_dnr_n, h->d_taps, h->d_masks, h->d_tw, h->d_win256, h->d_win1024,
                    h->d_tw256, h->d_sin512, h->d_fe_hist, h->d_nc_coeff, h->d_nc_prev, h->d_nc_energy, h->d_nc_first, h->d_dn_coeff,
                    h->d_dn_prev, h->d_dn_energy, h->d_dn_first, h->d_agc_env, h->d_conv_last, h->d_nfloor, h->d_bq_state,
                    h->d_spec_prev, h->d_spec_sum, h->d_spec_out, h->d_ring, h->d_spec1024_out, h->d_view, h->d_smeter,
                    h->d_waterfall, h->d_wf_head, h->d_wf_rows, h->d_wf_col,
                    h->d_mid_a, h->d_mid_b, h->d_scr, h->d_dbg, h->d_hp_iq};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < rdsp_gpu::kMaxStage; i++) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
        if (h->d_in_stage2[i]) cudaFree(h->d_in_stage2[i]);
        if (h->d_out_stage2[i]) cudaFree(h->d_out_stage2[i]);
    }
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int g = 0; g < kMaxGroups; g++)
        for (int k = 0; k < 3; k++) if (h->ev_group[g][k]) cudaEventDestroy(h->ev_group[g][k]);
    for (int g = 0; g < kMaxGroups; g++)
        for (int k = 0; k < 4; k++) if (h->ev_mark[g][k]) cudaEventDestroy(h->ev_mark[g][k]);
    if (h->ev_front) cudaEventDestroy(h->ev_front);
    for (int s = 0; s < kStreams; s++) {
        if (h->stage_stream[s]) cudaStreamDestroy(h->stage_stream[s]);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
}

// ---- one call on the device ---------------------------------------------------------------------------------
// Everything a process call enqueues between the fork on the handle's stream `st` and the joins back onto it.
// Nothing in here depends on how many calls came before except through (fe_cur, par) — the ping-pong index of the
// front end's delay lines and of the device-resident tick state — so the whole fork/join DAG can be captured into
// a CUDA graph once per (T, buffers, fe_cur, par, table version) and replayed (graph_launch below).
//
// Channel groups.  Channels are independent, and every kernel after the front end is bound by the latency of a
// per-channel recurrence rather than by throughput.  The call is therefore cut ACROSS channels: the front end takes
// all channels in one launch (tiles x time segments fill the SMs), then G channel groups walk the rest of the graph
// on streams of their own — notch -> AGC -> FFT filter -> DNR -> audio spectrum on one, high-pass -> IQ spectrum
// on another — with no dependency between groups, so the hardware overlaps G latency-bound chains.  Every launch
// still covers all T blocks of the call: state makes one round trip and the launch latency is paid once.
// (The first version cut the call along TIME instead; its timeline, tools/diag_timeline.py, showed every stage
// paying its launch + state latency per chunk and the DNR stage of 4 chunks taking twice its one-launch time.)
// While per-kernel profiling is on, everything runs as one group on one stream so that each kernel's time is its own.
// does the next call, if it covers ONE block, complete a frame of the audio spectrum?  (host mirror of the tick counter: the
// cadence of AudioAnalyzeFFT1024 — a frame every 4 ticks from tick 7 on — is uniform over the channels)
int one_block_frame(const rdsp_gpu *h, int T) { return (T == 1 && h->tick >= 7 && ((h->tick - 7) & 3) == 0) ? 1 : 0; }

int enqueue_call(rdsp_gpu *h, int T, const int16_t *iq, int16_t *audio, cudaStream_t st, bool piped, int fe_cur, int par, int frame)
{
    const bool fe = has(h, RDSP_STAGE_FRONTEND), notch = has(h, RDSP_STAGE_NOTCH), agc = has(h, RDSP_STAGE_AGC);
    const bool ff = has(h, RDSP_STAGE_FFTFILT), nr = has(h, RDSP_STAGE_NR);
    const int C = h->C;
    const bool mono_out = h->cfg.audio_layout == RDSP_AUDIO_MONO;      // audio = [T][C][128] (L only) instead of [T][C][128][2]
    const RdspTick *tick_in = h->d_tick + par;
    RdspTick *tick_out = h->d_tick + (par ^ 1);
    int G = 1;
    if (piped) {
        // measured (cfg5, 8192 channels, T = 8): 1 group 0.71 ms, 2 groups 0.90, 4 groups 0.83, 8 groups 1.11 — small
        // concurrent launches of the same kernel slow each other more than they overlap, so the default is ONE group
        // (main chain + spectrum branch + the AGC of the channels that bypass the notch on three streams)
        G = h->cfg.pipeline_chunks ? (int)h->cfg.pipeline_chunks : 1;
        if (G > kMaxGroups) G = kMaxGroups;
        while (G > 1 && C / G < 256) G--;
        CK(cudaEventRecord(h->ev_fork, st));                             // the input (and earlier calls) are in place
    }
    cudaStream_t s_front = piped ? h->stage_stream[3 * kMaxGroups] : st;
    const int naverage = (int)h->cfg.spec256_naverage;
    int lgn = 0; while ((1u << lgn) < h->cfg.spec256_naverage) lgn++;
    float *dbg = h->d_dbg;

    // ---- K0+K1+K2, all channels
    bool front_last = false;
    if (fe) {
        if (piped) CK(cudaStreamWaitEvent(s_front, h->ev_fork, 0));
        front_last = !(notch || agc || ff);
        FrontArgs a{};
        a.iq = iq; a.out_mono = front_last ? (mono_out ? audio : nullptr) : h->d_mid_a; a.out_stereo = (front_last && !mono_out) ? audio : nullptr;
        a.dbg = front_last ? dbg : nullptr; a.par = h->d_par; a.taps = h->d_taps; a.C = C; a.T = T;
        Prof pr(h, KK_FRONT, s_front);
        if (h->front_tc) {
            a.hist = fe_cur ? h->d_fe_hist2 : h->d_fe_hist;
            a.hist_out = fe_cur ? h->d_fe_hist : h->d_fe_hist2;
            FrontTcTables tb{};
            tb.tile_ch = h->d_tile_ch; tb.tile_rows = h->d_tile_rows; tb.toep = h->d_toep; tb.toep_map = h->toep_map; tb.n_tiles = h->n_tiles;
            tb.any_sam = h->any_sam; tb.sam_tiles = h->sam_tiles; tb.n_am_tiles = h->n_am_tiles;
            a.sam_state = h->d_sam_state; a.nb_ref = h->d_nb_ref;
            launch_front_tc(a, tb, s_front);
        } else {
            if (h->any_sam) { h->err = "SAM and the noise blanker are only built in the tensor-core front end (unset RDSP_FRONT_IMPL)"; return RDSP_ERR_STATE; }
            a.hist = fe_cur ? h->d_fe_hist2 : h->d_fe_hist;
            launch_front(a, s_front);
        }
    }
    if (piped && fe) CK(cudaEventRecord(h->ev_front, s_front));

    // sub-range of a sorted channel list that falls into [c0, c1)
    auto sub = [](const std::vector<int> &l, int c0, int c1, int &first, int &count) {
        first = (int)(std::lower_bound(l.begin(), l.end(), c0) - l.begin());
        count = (int)(std::lower_bound(l.begin(), l.end(), c1) - l.begin()) - first;
    };

    // one chain of the audio graph on stream `cs` for the channels [c0, c1) of class `cls`:
    //   0 = every channel (no split), 1 = the channels that bypass the notch, 2 = the channels whose notch runs.
    // With classes 1 and 2 on two streams the latency-bound notch -> AGC -> ... chain of the (few) notched channels runs
    // beside the wide kernels of the others instead of in front of them.
    // the spectrum branches keep the issue slots contended while the NLMS kernels run (k_nlms.cu, launch_nlms)
    const int nlms_contended = (has(h, RDSP_STAGE_SPEC256) || has(h, RDSP_STAGE_SPEC1024)) ? 1 : 0;
    // short calls: every kernel of a chain is a programmatic dependent of the one before it (kernels.h); `first` marks the kernel
    // that follows an event wait (no kernel in front of it on its stream)
    const int pdl = (piped && T <= h->pdl_max_T) ? 1 : 0;
    // one-block calls (the sketch's calling pattern): the kernel that emits a channel's audio — the DNR, else the FFT filter — also
    // appends the row to the ring of the audio spectrum, whose own kernel then has nothing to do on the three ticks out of four that
    // complete no frame (it is launched as ONE CTA that advances the tick counter, instead of a CTA per channel to copy 256 bytes).  Only with one
    // block per call: a frame reads the eight newest rows, and rows appended ahead of it would overwrite the oldest of them.
    const bool fused_append = T == 1 && ff && has(h, RDSP_STAGE_SPEC1024) && !h->nlms_direct;
    auto run_chain = [&](cudaStream_t cs, int cls, int c0, int c1, cudaEvent_t *marks, int *n_marks) -> int {
        auto mark = [&](int k) { if (marks && piped && h->spec_after >= 2) { cudaEventRecord(marks[k], cs); if (*n_marks < k + 1) *n_marks = k + 1; } };
        bool first = true;                                   // the chain's first kernel has no kernel before it on `cs`
        auto dep = [&]() { const int d = (pdl && !first) ? 1 : 0; first = false; return d; };
        const int nc = c1 - c0;
        int f_notch = 0, n_notch = 0, f_plain = 0, n_plain = 0;
        if (notch) { sub(h->l_notch, c0, c1, f_notch, n_notch); sub(h->l_plain, c0, c1, f_plain, n_plain); }
        const int *cls_list = cls == 1 ? h->d_list_plain + f_plain : (cls == 2 ? h->d_list_notch + f_notch : nullptr);
        const int cls_n = cls == 1 ? n_plain : (cls == 2 ? n_notch : nc);
        if (cls_n <= 0) return RDSP_OK;
        const int16_t *mono = nullptr;
        if (fe) {
            mono = h->d_mid_a;
            if (notch || agc) {
                AgcArgs ag{};
                ag.out_mono = ff ? h->d_mid_b : (mono_out ? audio : nullptr); ag.out_stereo = (ff || mono_out) ? nullptr : audio;
                ag.dbg = ff ? nullptr : dbg; ag.env = h->d_agc_env; ag.par = h->d_par; ag.C = C; ag.T = T;
                ag.agc_stage = agc ? 1 : 0;
                ag.target = h->cfg.agc_target; ag.max_gain = h->cfg.agc_max_gain; ag.alpha_a = h->agc_alpha_a;
                if (cls != 2) {
                    // channels that bypass the notch read the front end's q15 rows ...
                    ag.list = notch ? h->d_list_plain + f_plain : nullptr; ag.n_list = notch ? n_plain : nc; ag.ch0 = c0;
                    ag.in_q15 = h->d_mid_a; ag.in_f32 = nullptr;
                    if (ag.n_list > 0) { ag.pdl = dep(); Prof pr(h, KK_AGC, cs); launch_agc(ag, cs); }
                }
                if (cls != 1 && notch && n_notch > 0) {
                    NlmsArgs n{};
                    n.list = h->d_list_notch + f_notch; n.n_list = n_notch; n.C = C; n.T = T;
                    n.in_q15 = h->d_mid_a; n.out_f32 = h->d_scr;
                    n.coeff = h->d_nc_coeff; n.prev = h->d_nc_prev; n.energy = h->d_nc_energy; n.first = h->d_nc_first;
                    n.par = h->d_par; n.mode = 0; n.contended = nlms_contended; n.direct = h->nlms_direct; n.pdl = dep();
                    { Prof pr(h, KK_NOTCH, cs); launch_nlms(n, cs); }
                    mark(0);
                    // ... the others read the notch's f32 error signal
                    ag.list = h->d_list_notch + f_notch; ag.n_list = n_notch; ag.in_q15 = nullptr; ag.in_f32 = h->d_scr; ag.pdl = dep();
                    { Prof pr(h, KK_AGC, cs); launch_agc(ag, cs); }
                    mark(1);
                }
                mono = h->d_mid_b;
            }
        }
        if (ff) {
            FftFiltArgs f{};
            f.in_mono = fe ? mono : nullptr; f.in_stereo = fe ? nullptr : iq; f.out_stereo = mono_out ? nullptr : audio; f.out_mono = mono_out ? audio : nullptr; f.out_f32_L = h->d_scr;
            f.dbg = dbg; f.last = h->d_conv_last; f.nfloor = h->d_nfloor; f.masks = h->d_masks; f.tw256 = h->d_tw256; f.sin512 = h->d_sin512;
            f.par = h->d_par; f.C = C; f.T = T; f.list = cls_list; f.ch0 = c0; f.n = cls_n; f.nr_stage = nr ? 1 : 0; f.pdl = dep();
            if (fused_append) { f.ring = h->d_ring; f.tick_in = tick_in; }
            { Prof pr(h, KK_FFTFILT, cs); launch_fftfilt(f, cs); }
            if (cls != 1 && notch) mark(2);
            if (nr) {
                const std::vector<int> &ld = cls == 1 ? h->l_dnr_p : (cls == 2 ? h->l_dnr_n : h->l_dnr);
                const int *dl = cls == 1 ? h->d_list_dnr_p : (cls == 2 ? h->d_list_dnr_n : h->d_list_dnr);
                int f_dnr = 0, n_dnr = 0;
                sub(ld, c0, c1, f_dnr, n_dnr);
                if (n_dnr > 0) {
                    NlmsArgs n{};
                    n.list = dl + f_dnr; n.n_list = n_dnr; n.C = C; n.T = T;
                    n.in_f32 = h->d_scr; n.out_stereo = mono_out ? nullptr : audio; n.out_mono = mono_out ? audio : nullptr; n.dbg = dbg;
                    n.coeff = h->d_dn_coeff; n.prev = h->d_dn_prev; n.energy = h->d_dn_energy; n.first = h->d_dn_first;
                    n.par = h->d_par; n.mode = 1; n.contended = nlms_contended; n.direct = h->nlms_direct; n.pdl = dep();
                    if (fused_append) { n.ring = h->d_ring; n.tick_in = tick_in; }
                    { Prof pr(h, KK_DNR, cs); launch_nlms(n, cs); }
                    if (cls != 1 && notch) mark(3);
                }
            }
        }
        if (has(h, RDSP_STAGE_SPEC1024)) {
            Spec1024Args s1{};
            s1.audio = audio; s1.audio_mono = mono_out ? 1 : 0; s1.ring = h->d_ring; s1.output = h->d_spec1024_out; s1.C = C; s1.T = T;
            s1.list = cls_list; s1.ch0 = c0; s1.n = cls_n;
            s1.tick_in = tick_in; s1.tick_out = tick_out; s1.tw = h->d_tw; s1.win = h->d_win1024; s1.pdl = dep();
            s1.appended = fused_append ? 1 : 0;
            // ... and on the three ticks out of four that complete no frame all that is left of this kernel is the tick counter: ONE
            // CTA (the host mirrors the cadence, and a call shape is captured per `frame`)
            if (fused_append && !frame) s1.n = 1;
            { Prof pr(h, KK_SPEC1024, cs); launch_spec1024(s1, cs); }
        }
        return RDSP_OK;
    };

    for (int g = 0; g < G; g++) {
        const int c0 = (int)((long long)C * g / G), c1 = (int)((long long)C * (g + 1) / G), nc = c1 - c0;
        cudaStream_t s_main = piped ? h->stage_stream[g] : st;
        cudaStream_t s_spec = piped ? h->stage_stream[kMaxGroups + g] : st;
        cudaStream_t s_side = piped ? h->stage_stream[2 * kMaxGroups + g] : st;
        if (piped) CK(cudaStreamWaitEvent(s_main, fe ? h->ev_front : h->ev_fork, 0));

        // the channels whose notch runs form their own chain on the side stream (when both classes exist)
        int fn = 0, nn = 0, fp = 0, np = 0;
        if (fe && notch) { sub(h->l_notch, c0, c1, fn, nn); sub(h->l_plain, c0, c1, fp, np); }
        static const bool no_split = [] { const char *e = getenv("RDSP_NO_SPLIT"); return e && e[0] == '1'; }();   // experiments
        const bool split = piped && fe && notch && nn > 0 && np > 0 && !no_split;
        int n_marks = 0;
        if (split) {
            CK(cudaStreamWaitEvent(s_side, h->ev_front, 0));
            int rc2 = run_chain(s_side, 2, c0, c1, h->ev_mark[g], &n_marks);
            if (rc2 != RDSP_OK) return rc2;
            rc2 = run_chain(s_main, 1, c0, c1, nullptr, nullptr);
            if (rc2 != RDSP_OK) return rc2;
            CK(cudaEventRecord(h->ev_group[g][2], s_side));
            CK(cudaStreamWaitEvent(st, h->ev_group[g][2], 0));
        } else {
            const int rc2 = run_chain(s_main, 0, c0, c1, h->ev_mark[g], &n_marks);
            if (rc2 != RDSP_OK) return rc2;
        }

        // The spectrum branch (high-pass biquads -> IQ spectrum) needs only the input and nobody waits for it, so WHEN it
        // starts is a pure scheduling choice.  spec_after: 0 = with the front end, 1 = behind the front end, 2.. = behind
        // the notch / AGC / FFT filter / DNR of the chain that runs the notch (the longest dependent chain of the call: its
        // latency-bound kernels stretch in proportion to what shares their SMs — tools/diag_timeline.py).
        if (has(h, RDSP_STAGE_SPEC256)) {
            if (piped) {
                cudaEvent_t after = fe ? h->ev_front : h->ev_fork;
                if (h->spec_after == 0) after = h->ev_fork;
                else if (h->spec_after >= 2 && n_marks > 0) after = h->ev_mark[g][std::min(h->spec_after - 2, n_marks - 1)];
                CK(cudaStreamWaitEvent(s_spec, after, 0));
            }
            BiquadArgs b{};
            b.iq = iq; b.out = h->d_hp_iq; b.state = h->d_bq_state; b.C = C; b.T = T; b.ch0 = c0; b.n = nc;
            b.b0 = h->bq[0]; b.b1 = h->bq[1]; b.b2 = h->bq[2]; b.a1 = h->bq[3]; b.a2 = h->bq[4];
            { Prof pr(h, KK_BIQUAD, s_spec); launch_biquad(b, s_spec); }
            Spec256Args a{};
            a.iq = h->d_hp_iq; a.prev = h->d_spec_prev; a.sum = h->d_spec_sum; a.output = h->d_spec_out;
            a.C = C; a.T = T; a.ch0 = c0; a.n = nc; a.tick_in = tick_in; a.tick_out = tick_out; a.naverage = naverage;
            a.div_shift = 32 + lgn;
            a.div_magic = ((1ull << a.div_shift) + h->cfg.spec256_naverage - 1) / h->cfg.spec256_naverage;
            a.tw = h->d_tw; a.win = h->d_win256;
            for (int d0 = 0; d0 < 4; d0++) for (int k = 0; k < 3; k++) a.tw3[d0][k] = h->tw3_256[d0][k];
            { Prof pr(h, KK_SPEC256, s_spec); launch_spec256(a, s_spec); }
        } else if (piped) {
            CK(cudaStreamWaitEvent(s_spec, h->ev_fork, 0));
        }
        if (piped) {
            // join: the call is complete on the handle's stream when every group has finished all of its streams
            CK(cudaEventRecord(h->ev_group[g][0], s_main));
            CK(cudaEventRecord(h->ev_group[g][1], s_spec));
            CK(cudaStreamWaitEvent(st, h->ev_group[g][0], 0));
            CK(cudaStreamWaitEvent(st, h->ev_group[g][1], 0));
        }
    }
    if (piped && fe && front_last) CK(cudaStreamWaitEvent(st, h->ev_front, 0));
    CK(cudaGetLastError());
    return RDSP_OK;
}

// host mirror of the uniform counters (analyze_fft256iq.cpp:73-77,99-113; AudioAnalyzeFFT1024 frame cadence): which
// read-outs became available, where the ping-pong buffers stand after the call
void advance_host_state(rdsp_gpu *h, int T)
{
    const int naverage = (int)h->cfg.spec256_naverage;
    if (has(h, RDSP_STAGE_SPEC256)) {
        const int updates = T - (h->spec_have_prev ? 0 : 1);
        h->spec_have_prev = 1;
        const int total = h->spec_count + updates;
        if (total / naverage > 0) std::fill(h->spec_ready.begin(), h->spec_ready.end(), (uint8_t)1);
        h->spec_count = total % naverage;
    }
    if (has(h, RDSP_STAGE_SPEC1024)) {
        int n_fft_frames = 0;
        for (int t = 0; t < T; t++) {
            const unsigned long long tk = h->tick + t;
            if (tk >= 7 && ((tk - 7) & 3) == 0) n_fft_frames++;
        }
        if (n_fft_frames) std::fill(h->spec1024_ready.begin(), h->spec1024_ready.end(), (uint8_t)1);
    }
    h->tick += T;
    h->tick_par ^= 1;
    if (has(h, RDSP_STAGE_FRONTEND) && h->front_tc) h->fe_hist_cur ^= 1;
}

void drop_graphs(rdsp_gpu *h)
{
    for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// The device side of one call through the graph cache.  The sketch's calling pattern is one tick per 128 samples
// (RadioDSP_SDR_RX.ino:198): about a dozen launches on three streams plus ~30 event calls for 130 us of GPU work.  A call
// shape is captured the second time it is seen (same T, same buffers, same ping-pong phase, same tables) and replayed with
// ONE cudaGraphLaunch from then on; the first sight runs the plain path (it also performs the one-time per-device kernel
// attribute set-up, which must not happen inside a capture).
int run_call(rdsp_gpu *h, int T, const int16_t *iq, int16_t *audio, cudaStream_t st)
{
    const bool piped = !h->profiling;
    const int fe_cur = h->fe_hist_cur, par = h->tick_par, frame = one_block_frame(h, T);
    if (!h->use_graphs || h->profiling || h->timeline) return enqueue_call(h, T, iq, audio, st, piped, fe_cur, par, frame);

    GraphEntry *e = nullptr;
    for (auto &g : h->graphs)
        if (g.T == T && g.iq == iq && g.audio == audio && g.fe_cur == fe_cur && g.par == par && g.st == st && g.frame == frame) { e = &g; break; }
    h->graph_clock++;
    if (e && e->exec) {
        e->last_use = h->graph_clock;
        CK(cudaGraphLaunch(e->exec, st));
        h->launches += e->launches;
        h->graph_replays++;
        return RDSP_OK;
    }
    if (!e) {
        // first sight: remember the shape, run it plainly
        if (h->graphs.size() >= rdsp_gpu::kMaxGraphs) {
            size_t victim = 0;
            for (size_t i = 1; i < h->graphs.size(); i++) if (h->graphs[i].last_use < h->graphs[victim].last_use) victim = i;
            if (h->graphs[victim].exec) cudaGraphExecDestroy(h->graphs[victim].exec);
            h->graphs.erase(h->graphs.begin() + (long)victim);
        }
        GraphEntry n{};
        n.T = T; n.iq = iq; n.audio = audio; n.fe_cur = fe_cur; n.par = par; n.st = st; n.frame = frame; n.last_use = h->graph_clock;
        h->graphs.push_back(n);
        return enqueue_call(h, T, iq, audio, st, piped, fe_cur, par, frame);
    }
    // second sight: capture, instantiate, launch
    e->last_use = h->graph_clock;
    const uint64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_call(h, T, iq, audio, st, piped, fe_cur, par, frame);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const uint64_t n_launch = h->launches - l0;
    h->launches = l0;
    if (rc != RDSP_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc != RDSP_OK) return rc;
        h->use_graphs = false;                                   // capture not possible here (e.g. a caller stream that is itself capturing)
        drop_graphs(h);
        return enqueue_call(h, T, iq, audio, st, piped, fe_cur, par, frame);
    }
    if (getenv("RDSP_GRAPH_DEBUG")) {
        // development aid: what the capture recorded (node count, kernel priorities)
        size_t nn = 0;
        cudaGraphGetNodes(graph, nullptr, &nn);
        std::vector<cudaGraphNode_t> nodes(nn);
        cudaGraphGetNodes(graph, nodes.data(), &nn);
        fprintf(stderr, "[rdsp graph] T=%d: %zu nodes, %llu kernels;", T, nn, (unsigned long long)n_launch);
        for (auto nd : nodes) {
            cudaGraphNodeType ty;
            cudaGraphNodeGetType(nd, &ty);
            if (ty != cudaGraphNodeTypeKernel) continue;
            cudaLaunchAttributeValue v{};
            if (cudaGraphKernelNodeGetAttribute(nd, cudaLaunchAttributePriority, &v) == cudaSuccess) fprintf(stderr, " prio %d", v.priority);
        }
        fprintf(stderr, "\n");
        cudaGetLastError();
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        cudaGetLastError();
        h->use_graphs = false;
        drop_graphs(h);
        return enqueue_call(h, T, iq, audio, st, piped, fe_cur, par, frame);
    }
    e->exec = exec;
    e->launches = n_launch;
    CK(cudaGraphLaunch(exec, st));
    h->launches += n_launch;
    h->graph_replays++;
    return RDSP_OK;
}

}  // namespace

extern "C" {

const char *rdsp_gpu_version(void) { return RDSP_VERSION; }

void rdsp_gpu_default_config(rdsp_gpu_config_t *cfg)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(*cfg);
    cfg->n_channels = 1;
    cfg->device = 0;
    cfg->stage_mask = RDSP_STAGE_ALL;
    cfg->max_blocks_per_call = 1;
    cfg->io_location = RDSP_IO_DEVICE;
    cfg->spec256_naverage = 30;                  // FFT.averageTogether(30), RadioDSP_SDR_RX.ino:145
    cfg->agc_target = 0.25f;
    cfg->agc_max_gain = 1000.0f;
    cfg->agc_attack_ms = 5.0f;
    cfg->agc_decay_ms[RDSP_AGC_FAST] = 100.0f;
    cfg->agc_decay_ms[RDSP_AGC_MEDIUM] = 500.0f;
    cfg->agc_decay_ms[RDSP_AGC_SLOW] = 2000.0f;
}

void rdsp_gpu_default_params(rdsp_chan_params_t *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->demod = RDSP_DEMOD_LSB;                   // RadioDSP_SDR_RX.ino:139
    p->audio_filter = RDSP_FILTER_2700;          // :138
    p->agc_mode = RDSP_AGC_MEDIUM;               // :121
    p->notch_on = 0;                             // :125
    p->notch_level = 20;
    p->nr_kind = RDSP_NR_OFF;                    // RDSP_general_includes.h:111
    p->nr_level = 0;
    p->pbt_lo_hz = 300.0f;                       // RadioDSP_SDR_RX.ino:183
    p->pbt_hi_hz = 4000.0f;
    p->in_gain = 1.0f;                           // :133
    p->out_gain = 0.5f;                          // :134
    p->iq_balance = 1.020f;                      // :135
    p->als_peak = 0;                             // SDR.setALSfilterNotch(), RDSP_controls.h:258
    p->nb_on = 0;                                // SDR.disableNoiseBlanker(), :131
    p->nb_threshold_db = 20.0f;                  // :130
}

const char *rdsp_gpu_last_error(const rdsp_gpu_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rdsp_gpu_create(const rdsp_gpu_config_t *cfg, rdsp_gpu_t **out)
{
    if (!cfg || !out) { g_create_error = "NULL argument"; return RDSP_ERR_INVALID; }
    *out = nullptr;
    if (cfg->struct_size != sizeof(rdsp_gpu_config_t)) { g_create_error = "config struct_size mismatch"; return RDSP_ERR_INVALID; }
    if (cfg->n_channels == 0 || cfg->n_channels > (1u << 24)) { g_create_error = "n_channels out of range"; return RDSP_ERR_RANGE; }
    if (cfg->max_blocks_per_call == 0 || cfg->max_blocks_per_call > 4096) { g_create_error = "max_blocks_per_call out of range"; return RDSP_ERR_RANGE; }
    const uint32_t sm = cfg->stage_mask;
    if ((sm & ~RDSP_STAGE_ALL) || sm == 0) { g_create_error = "stage_mask invalid"; return RDSP_ERR_RANGE; }
    if ((sm & (RDSP_STAGE_NOTCH | RDSP_STAGE_AGC)) && !(sm & RDSP_STAGE_FRONTEND)) { g_create_error = "NOTCH/AGC need FRONTEND"; return RDSP_ERR_STATE; }
    if ((sm & RDSP_STAGE_NR) && !(sm & RDSP_STAGE_FFTFILT)) { g_create_error = "NR needs FFTFILT"; return RDSP_ERR_STATE; }
    if ((sm & RDSP_STAGE_SPEC1024) && !(sm & (RDSP_STAGE_FRONTEND | RDSP_STAGE_FFTFILT))) { g_create_error = "SPEC1024 needs an audio path"; return RDSP_ERR_STATE; }
    if (cfg->spec256_naverage == 0 || cfg->spec256_naverage > 255) { g_create_error = "spec256_naverage must be 1..255"; return RDSP_ERR_RANGE; }
    if (cfg->io_location > RDSP_IO_HOST) { g_create_error = "io_location invalid"; return RDSP_ERR_RANGE; }
    if (cfg->pipeline_chunks > (uint32_t)kMaxGroups) { g_create_error = "pipeline_chunks must be 0 (auto) .. 8"; return RDSP_ERR_RANGE; }
    if (cfg->graph_mode > RDSP_GRAPH_OFF) { g_create_error = "graph_mode invalid"; return RDSP_ERR_RANGE; }
    if (cfg->audio_layout > RDSP_AUDIO_MONO) { g_create_error = "audio_layout invalid"; return RDSP_ERR_RANGE; }
    if (!(cfg->agc_target > 0.f) || !(cfg->agc_max_gain > 0.f) || !(cfg->agc_attack_ms > 0.f)) { g_create_error = "AGC constants invalid"; return RDSP_ERR_RANGE; }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return RDSP_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return RDSP_ERR_RANGE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not sm_100-class (kernels are built for sm_100a only)";
        return RDSP_ERR_CUDA;
    }

    rdsp_gpu *h = new (std::nothrow) rdsp_gpu();
    if (!h) { g_create_error = "out of host memory"; return RDSP_ERR_NOMEM; }
    h->cfg = *cfg;
    h->C = (int)cfg->n_channels;
    h->maxT = (int)cfg->max_blocks_per_call;
    const size_t C = (size_t)h->C, T = (size_t)h->maxT;

    auto fail = [&](cudaError_t ce, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        free_all(h);
        delete h;
        return ce == cudaErrorMemoryAllocation ? RDSP_ERR_NOMEM : RDSP_ERR_CUDA;
    };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(e_, #call); } while (0)

    CKC(cudaSetDevice(cfg->device));
    CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CKC(cudaEventCreate(&h->ev_fork));                            // timing enabled: origin of the RDSP_TIMELINE printout
    CKC(cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming));
    for (int g = 0; g < kMaxGroups; g++)
        for (int k = 0; k < 3; k++) CKC(cudaEventCreateWithFlags(&h->ev_group[g][k], cudaEventDisableTiming));
    for (int g = 0; g < kMaxGroups; g++)
        for (int k = 0; k < 4; k++) CKC(cudaEventCreateWithFlags(&h->ev_mark[g][k], cudaEventDisableTiming));
    {
        // the main chain (and the front end it waits for) outranks the spectrum and side branches: when an SM slot frees
        // up, a block of the latency-critical notch / DNR launch goes first, the wide FFT grids fill what is left
        int prio_lo = 0, prio_hi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int s = 0; s < kStreams; s++) {
            static const bool no_prio = [] { const char *e = getenv("RDSP_NO_PRIO"); return e && e[0] == '1'; }();   // experiments
            const bool critical = no_prio || s < kMaxGroups || s >= 2 * kMaxGroups;      // main chains, notch-class chains, front end
            CKC(cudaStreamCreateWithPriority(&h->stage_stream[s], cudaStreamNonBlocking, critical ? prio_hi : prio_lo));
        }
    }

    // host-side tables
    for (int m = 0; m < RDSP_DEMOD_COUNT; m++) rdsp_host::design_hilbert_pair(m, h->taps[m], h->taps[RDSP_DEMOD_COUNT + m]);
    for (int f = 0; f < RDSP_FILTER_COUNT; f++) rdsp_host::design_bandpass(f, h->taps[2 * RDSP_DEMOD_COUNT + f]);
    h->agc_alpha_a = rdsp_host::agc_alpha(cfg->agc_attack_ms);
    for (int m = 1; m < 4; m++) h->agc_alpha_d[m] = cfg->agc_decay_ms[m] > 0.f ? rdsp_host::agc_alpha(cfg->agc_decay_ms[m]) : 0.f;
    rdsp_host::biquad_highpass_q30(500.0f, 0.5f, h->bq);           // RadioDSP_SDR_RX.ino:155-156

    rdsp_chan_params_t defp;
    rdsp_gpu_default_params(&defp);
    h->par.assign(C, defp);
    h->dpar.resize(C);
    h->dnr_old_level.assign(C, 15);                                 // oldNRLevel = 15 + Init_LMS_NR(15): RDSP_convolutional.h:80, .ino:172
    h->notch_old_level.assign(C, -1);
    h->custom_mask.assign(C, -1);
    const int mid = mask_id_for(h, defp.pbt_lo_hz, defp.pbt_hi_hz);
    for (size_t ch = 0; ch < C; ch++) { derive_params(h, (int)ch); h->dpar[ch].mask_id = mid; }
    h->spec_ready.assign(C, 0);
    h->spec1024_ready.assign(C, 0);

    CKC(dalloc(&h->d_par, C));
    CKC(dalloc(&h->d_tick, (size_t)2));
    h->use_graphs = cfg->graph_mode == RDSP_GRAPH_AUTO;
    if (const char *e = getenv("RDSP_GRAPH")) h->use_graphs = e[0] != '0';
    CKC(dalloc(&h->d_taps, (size_t)15 * RDSP_TAPS_PAD));
    if (sm & RDSP_STAGE_FRONTEND) {
        CKC(dalloc(&h->d_fe_hist, C * 3 * RDSP_BLK));
        CKC(dalloc(&h->d_fe_hist2, C * 3 * RDSP_BLK));
        CKC(dalloc(&h->d_toep, front_tc_toeplitz_bytes()));
        if (front_tc_make_tensor_map(h->d_toep, &h->toep_map) != 0) return fail(cudaErrorUnknown, "cuTensorMapEncodeTiled (TMA descriptor of the Toeplitz table)");
        CKC(dalloc(&h->d_sam_state, C * 4));
        CKC(dalloc(&h->d_nb_ref, C));
        if (const char *e = getenv("RDSP_FRONT_IMPL")) h->front_tc = !(e[0] == 'c' || e[0] == 'C');
        if (const char *e = getenv("RDSP_TIMELINE")) h->timeline = e[0] == '1';
    }
    if (const char *e = getenv("RDSP_NLMS_IMPL")) h->nlms_direct = e[0] == 'd';
    if (const char *e = getenv("RDSP_PDL_MAX_T")) h->pdl_max_T = atoi(e);
    if (const char *e = getenv("RDSP_SPEC_AFTER")) h->spec_after = std::max(0, std::min(5, atoi(e)));
    if (const char *e = getenv("RDSP_SPEC_WITH_FRONT")) { if (e[0] == '1') h->spec_after = 0; }
    if (sm & RDSP_STAGE_FRONTEND) {
        CKC(dalloc(&h->d_mid_a, T * C * RDSP_BLK));
    }
    if (sm & RDSP_STAGE_NOTCH) {
        CKC(dalloc(&h->d_list_notch, C));
        CKC(dalloc(&h->d_list_plain, C));
        CKC(dalloc(&h->d_nc_coeff, C * RDSP_LMS_NTAPS));
        CKC(dalloc(&h->d_nc_prev, C * RDSP_BLK));
        CKC(dalloc(&h->d_nc_energy, C));
        CKC(cudaMalloc((void **)&h->d_nc_first, C));
        CKC(cudaMemset(h->d_nc_first, 1, C));
    }
    if (sm & (RDSP_STAGE_NOTCH | RDSP_STAGE_AGC)) {
        CKC(dalloc(&h->d_agc_env, C));
        CKC(dalloc(&h->d_mid_b, T * C * RDSP_BLK));
    }
    if (sm & (RDSP_STAGE_NOTCH | RDSP_STAGE_NR)) CKC(dalloc(&h->d_scr, T * C * RDSP_BLK));
    if (sm & RDSP_STAGE_FFTFILT) {
        CKC(dalloc(&h->d_conv_last, C * 2 * RDSP_BLK));
        CKC(dalloc(&h->d_nfloor, C));
        CKC(dalloc(&h->d_tw256, (size_t)256));
        float cs[512];
        rdsp_host::make_twiddle_256_f32(cs);
        CKC(cudaMemcpy(h->d_tw256, cs, sizeof(cs), cudaMemcpyHostToDevice));
        CKC(dalloc(&h->d_sin512, (size_t)513));
        float st[513];
        rdsp_host::make_sin512_f32(st);
        CKC(cudaMemcpy(h->d_sin512, st, sizeof(st), cudaMemcpyHostToDevice));
    }
    if (sm & RDSP_STAGE_NR) {
        CKC(dalloc(&h->d_list_dnr, C));
        CKC(dalloc(&h->d_list_dnr_p, C));
        CKC(dalloc(&h->d_list_dnr_n, C));
        CKC(dalloc(&h->d_dn_coeff, C * RDSP_LMS_NTAPS));
        CKC(dalloc(&h->d_dn_prev, C * RDSP_BLK));
        CKC(dalloc(&h->d_dn_energy, C));
        CKC(cudaMalloc((void **)&h->d_dn_first, C));
        CKC(cudaMemset(h->d_dn_first, 1, C));
    }
    if (sm & (RDSP_STAGE_SPEC256 | RDSP_STAGE_SPEC1024)) {
        CKC(dalloc(&h->d_tw, (size_t)3072));
        std::vector<uint32_t> tw(3072);
        rdsp_host::make_twiddle_4096_q15(tw.data());
        std::vector<int2> tw2(3072);
        for (int k = 0; k < 3072; k++) tw2[k] = make_int2((int16_t)(tw[k] & 0xFFFFu), (int16_t)(tw[k] >> 16));
        CKC(cudaMemcpy(h->d_tw, tw2.data(), tw2.size() * sizeof(int2), cudaMemcpyHostToDevice));
        for (int d0 = 0; d0 < 4; d0++)
            for (int k = 1; k <= 3; k++) h->tw3_256[d0][k - 1] = tw2[256 * d0 * k];       // stage 3 of the 256-point transform (k_spec256)
    }
    if (sm & RDSP_STAGE_SPEC256) {
        CKC(dalloc(&h->d_bq_state, C * 8));
        CKC(dalloc(&h->d_hp_iq, T * C * 2 * RDSP_BLK));
        CKC(dalloc(&h->d_spec_prev, C * 2 * RDSP_BLK));
        CKC(dalloc(&h->d_spec_sum, C * 256));
        CKC(dalloc(&h->d_spec_out, C * 256));
        CKC(dalloc(&h->d_win256, (size_t)256));
        CKC(dalloc(&h->d_view, C * 256));
        CKC(dalloc(&h->d_smeter, C));
        CKC(dalloc(&h->d_waterfall, C * 50 * 128));
        CKC(dalloc(&h->d_wf_head, C));
        h->wf_head.assign(C, 0);
        int16_t w[256];
        rdsp_host::make_hann_q15(w, 256);
        CKC(cudaMemcpy(h->d_win256, w, sizeof(w), cudaMemcpyHostToDevice));
        std::vector<uint16_t> v0(C * 256, 0);
        for (size_t ch = 0; ch < C; ch++) v0[ch * 256] = 1;        // SpectrumViewOld[512] = {1}, RDSP_display.h:32
        CKC(cudaMemcpy(h->d_view, v0.data(), v0.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    if (sm & RDSP_STAGE_SPEC1024) {
        CKC(dalloc(&h->d_ring, C * 8 * RDSP_BLK));
        CKC(dalloc(&h->d_spec1024_out, C * 512));
        CKC(dalloc(&h->d_win1024, (size_t)1024));
        int16_t w[1024];
        rdsp_host::make_hann_q15(w, 1024);
        CKC(cudaMemcpy(h->d_win1024, w, sizeof(w), cudaMemcpyHostToDevice));
    }
    if (cfg->debug_f32) CKC(dalloc(&h->d_dbg, T * C * 2 * RDSP_BLK));
    if (cfg->io_location == RDSP_IO_HOST) {
        CKC(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
        CKC(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        h->n_stage = h->maxT <= 4 ? rdsp_gpu::kMaxStage : 2;
        if (const char *e = getenv("RDSP_HOST_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= rdsp_gpu::kMaxStage) h->n_stage = v; }
        for (int i = 0; i < h->n_stage; i++) {
            CKC(dalloc(&h->d_in_stage2[i], T * C * 2 * RDSP_BLK));
            CKC(dalloc(&h->d_out_stage2[i], T * C * 2 * RDSP_BLK));
            CKC(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
        }
    }
#undef CKC
    h->par_dirty = true;
    h->taps_dirty = true;
    *out = h;
    return RDSP_OK;
}

void rdsp_gpu_destroy(rdsp_gpu_t *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    prof_collect(h);
    drop_graphs(h);
    free_all(h);
    delete h;
}

int rdsp_gpu_set_mode(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, const rdsp_chan_params_t *p)
{
    if (!h || !p) return RDSP_ERR_INVALID;
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    std::string why;
    int rc = validate_params(p, why);
    if (rc != RDSP_OK) { h->err = why; return rc; }
    const int mid = mask_id_for(h, p->pbt_lo_hz, p->pbt_hi_hz);
    bool changed = false;
    for (uint32_t ch = ch_first; ch < ch_first + ch_count; ch++) {
        if (memcmp(&h->par[ch], p, sizeof(*p)) == 0) continue;           // the sketch re-sends its settings every tick
        if (h->par[ch].pbt_lo_hz != p->pbt_lo_hz || h->par[ch].pbt_hi_hz != p->pbt_hi_hz) h->custom_mask[ch] = -1;
        h->par[ch] = *p;
        derive_params(h, (int)ch);
        h->dpar[ch].mask_id = h->custom_mask[ch] >= 0 ? h->custom_mask[ch] : mid;
        changed = true;
    }
    if (changed) h->par_dirty = true;
    return RDSP_OK;
}

int rdsp_gpu_get_mode(rdsp_gpu_t *h, uint32_t ch, rdsp_chan_params_t *p)
{
    if (!h || !p) return RDSP_ERR_INVALID;
    if (ch >= (uint32_t)h->C) { h->err = "channel out of bounds"; return RDSP_ERR_RANGE; }
    *p = h->par[ch];
    return RDSP_OK;
}

int rdsp_gpu_set_stream(rdsp_gpu_t *h, void *cuda_stream)
{
    if (!h) return RDSP_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    drop_graphs(h);                                          // captured shapes belong to the stream they were captured on
    return RDSP_OK;
}

int rdsp_gpu_synchronize(rdsp_gpu_t *h)
{
    if (!h) return RDSP_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d2h_stream) CK(cudaStreamSynchronize(h->d2h_stream));
    if (h->h2d_stream) CK(cudaStreamSynchronize(h->h2d_stream));
    return RDSP_OK;
}

int rdsp_gpu_stream_join(rdsp_gpu_t *h)
{
    if (!h) return RDSP_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    if (h->d2h_stream) {
        for (int i = 0; i < h->n_stage; i++) CK(cudaStreamWaitEvent(h->stream, h->ev_d2h[i], 0));
    }
    return RDSP_OK;
}

int rdsp_gpu_process_blocks(rdsp_gpu_t *h, uint32_t n_blocks, const int16_t *iq_in, int16_t *audio_out)
{
    if (!h || !iq_in) return RDSP_ERR_INVALID;
    const bool audio_path = has(h, RDSP_STAGE_FRONTEND) || has(h, RDSP_STAGE_FFTFILT);
    if (audio_path && !audio_out) { h->err = "audio_out is NULL"; return RDSP_ERR_INVALID; }
    if (n_blocks == 0) return RDSP_OK;
    if (n_blocks > (uint32_t)h->maxT) { h->err = "n_blocks exceeds max_blocks_per_call"; return RDSP_ERR_RANGE; }
    if (((uintptr_t)iq_in & 15) || ((uintptr_t)audio_out & 15)) { h->err = "buffers must be 16-byte aligned"; return RDSP_ERR_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    int rc = sync_tables(h);
    if (rc != RDSP_OK) return rc;

    const int C = h->C, T = (int)n_blocks;
    const size_t io_bytes = (size_t)T * C * 2 * RDSP_BLK * sizeof(int16_t);
    const size_t out_bytes = h->cfg.audio_layout == RDSP_AUDIO_MONO ? io_bytes / 2 : io_bytes;
    const int16_t *iq = iq_in;
    int16_t *audio = audio_out;
    const bool host_io = h->cfg.io_location == RDSP_IO_HOST;
    const int hb = (int)(h->host_calls % (unsigned)h->n_stage);
    cudaStream_t st = h->stream;
    if (host_io) {
        // staging buffer hb was last read by the kernels of call n - n_stage and last drained by the D2H of that call
        CK(cudaStreamWaitEvent(h->h2d_stream, h->ev_comp[hb], 0));
        CK(cudaMemcpyAsync(h->d_in_stage2[hb], iq_in, io_bytes, cudaMemcpyHostToDevice, h->h2d_stream));
        CK(cudaEventRecord(h->ev_h2d[hb], h->h2d_stream));
        CK(cudaStreamWaitEvent(st, h->ev_h2d[hb], 0));
        CK(cudaStreamWaitEvent(st, h->ev_d2h[hb], 0));
        iq = h->d_in_stage2[hb];
        audio = h->d_out_stage2[hb];
    }

    rc = run_call(h, T, iq, audio, st);
    if (rc != RDSP_OK) return rc;
    advance_host_state(h, T);

    if (host_io) {
        CK(cudaEventRecord(h->ev_comp[hb], st));
        if (audio_path) {
            CK(cudaStreamWaitEvent(h->d2h_stream, h->ev_comp[hb], 0));
            CK(cudaMemcpyAsync(audio_out, h->d_out_stage2[hb], out_bytes, cudaMemcpyDeviceToHost, h->d2h_stream));
            CK(cudaEventRecord(h->ev_d2h[hb], h->d2h_stream));
        }
        h->host_calls++;
    }
    if (!h->cfg.async) {
        CK(cudaStreamSynchronize(h->stream));
        if (host_io) CK(cudaStreamSynchronize(h->d2h_stream));
    }
    if (h->timeline) prof_collect(h);
    return RDSP_OK;
}

int rdsp_gpu_process_block(rdsp_gpu_t *h, const int16_t *iq_in, int16_t *audio_out)
{
    return rdsp_gpu_process_blocks(h, 1, iq_in, audio_out);
}

static int read_rows(rdsp_gpu_t *h, const uint16_t *dsrc, size_t row, std::vector<uint8_t> &flags,
                     uint32_t ch_first, uint32_t ch_count, uint16_t *out, uint8_t *ready)
{
    if (!out) return RDSP_ERR_INVALID;
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, dsrc + (size_t)ch_first * row, (size_t)ch_count * row * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < ch_count; i++) {
        if (ready) ready[i] = flags[ch_first + i];
        flags[ch_first + i] = 0;                 // available() clears the flag, analyze_fft256iq.h:61-67
    }
    return RDSP_OK;
}

int rdsp_gpu_read_spectrum(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *out, uint8_t *ready)
{
    if (!h) return RDSP_ERR_INVALID;
    if (!has(h, RDSP_STAGE_SPEC256)) { h->err = "handle has no SPEC256 stage"; return RDSP_ERR_STATE; }
    return read_rows(h, h->d_spec_out, 256, h->spec_ready, ch_first, ch_count, out, ready);
}

int rdsp_gpu_read_audio_spectrum(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *out, uint8_t *ready)
{
    if (!h) return RDSP_ERR_INVALID;
    if (!has(h, RDSP_STAGE_SPEC1024)) { h->err = "handle has no SPEC1024 stage"; return RDSP_ERR_STATE; }
    return read_rows(h, h->d_spec1024_out, 512, h->spec1024_ready, ch_first, ch_count, out, ready);
}

int rdsp_gpu_read_panadapter(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *trace, float *smeter)
{
    if (!h || !trace || !smeter) return RDSP_ERR_INVALID;
    if (!has(h, RDSP_STAGE_SPEC256)) { h->err = "handle has no SPEC256 stage"; return RDSP_ERR_STATE; }
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    CK(cudaSetDevice(h->cfg.device));
    // older waterfall lines "move one row down": the head of each channel's ring steps back and takes the new line
    for (uint32_t i = 0; i < ch_count; i++) h->wf_head[ch_first + i] = (h->wf_head[ch_first + i] + 49) % 50;
    CK(cudaMemcpyAsync(h->d_wf_head, &h->wf_head[ch_first], ch_count * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    PanArgs a{};
    a.spec = h->d_spec_out; a.view = h->d_view; a.smeter = h->d_smeter; a.ch_first = (int)ch_first; a.ch_count = (int)ch_count;
    a.waterfall = h->d_waterfall; a.wf_head = h->d_wf_head;
    { Prof pr(h, KK_PAN); launch_panadapter(a, h->stream); }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(trace, h->d_view + (size_t)ch_first * 256, (size_t)ch_count * 256 * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(smeter, h->d_smeter + ch_first, (size_t)ch_count * sizeof(float), cudaMemcpyDeviceToHost));
    return RDSP_OK;
}

int rdsp_gpu_read_waterfall(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, uint16_t *rows, uint8_t *colour)
{
    if (!h || !rows) return RDSP_ERR_INVALID;
    if (!has(h, RDSP_STAGE_SPEC256)) { h->err = "handle has no SPEC256 stage"; return RDSP_ERR_STATE; }
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    CK(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)ch_count * 50 * 128;
    if (ch_count > h->wf_scratch_ch) {                  // read-out scratch lives in the handle: no allocation per call
        CK(cudaStreamSynchronize(h->stream));
        if (h->d_wf_rows) cudaFree(h->d_wf_rows);
        if (h->d_wf_col) cudaFree(h->d_wf_col);
        h->d_wf_rows = nullptr; h->d_wf_col = nullptr; h->wf_scratch_ch = 0;
        if (cudaMalloc((void **)&h->d_wf_rows, n * sizeof(uint16_t)) != cudaSuccess ||
            cudaMalloc((void **)&h->d_wf_col, n) != cudaSuccess) {
            if (h->d_wf_rows) { cudaFree(h->d_wf_rows); h->d_wf_rows = nullptr; }
            h->err = "out of device memory"; return RDSP_ERR_NOMEM;
        }
        h->wf_scratch_ch = ch_count;
    }
    uint16_t *d_rows = h->d_wf_rows;
    uint8_t *d_col = colour ? h->d_wf_col : nullptr;
    cudaError_t e = cudaMemcpyAsync(h->d_wf_head, &h->wf_head[ch_first], ch_count * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        WaterfallArgs w{};
        w.ring = h->d_waterfall; w.wf_head = h->d_wf_head; w.rows = d_rows; w.colour = d_col; w.ch_first = (int)ch_first; w.ch_count = (int)ch_count;
        { Prof pr(h, KK_PAN); launch_waterfall_read(w, h->stream); }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(rows, d_rows, n * sizeof(uint16_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && colour) e = cudaMemcpy(colour, d_col, n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { h->err = std::string("read_waterfall: ") + cudaGetErrorString(e); return RDSP_ERR_CUDA; }
    return RDSP_OK;
}

static int16_t *taps_row(rdsp_gpu_t *h, int kind, int index)
{
    if (kind == RDSP_TAPS_HILBERT_I && index >= 0 && index < RDSP_DEMOD_COUNT) return h->taps[index];
    if (kind == RDSP_TAPS_HILBERT_Q && index >= 0 && index < RDSP_DEMOD_COUNT) return h->taps[RDSP_DEMOD_COUNT + index];
    if (kind == RDSP_TAPS_BANDPASS && index >= 0 && index < RDSP_FILTER_COUNT) return h->taps[2 * RDSP_DEMOD_COUNT + index];
    return nullptr;
}

int rdsp_gpu_set_taps(rdsp_gpu_t *h, int kind, int index, const int16_t *taps, uint32_t n_taps)
{
    if (!h || !taps) return RDSP_ERR_INVALID;
    int16_t *row = taps_row(h, kind, index);
    if (!row || n_taps != RDSP_FIR_TAPS) { h->err = "tap table kind/index/length invalid"; return RDSP_ERR_RANGE; }
    memcpy(row, taps, RDSP_FIR_TAPS * sizeof(int16_t));
    h->taps_dirty = true;
    return RDSP_OK;
}

int rdsp_gpu_design_bandpass(float lo_hz, float hi_hz, int16_t *taps, uint32_t n_taps)
{
    if (!taps || n_taps != RDSP_FIR_TAPS) return RDSP_ERR_INVALID;
    if (!(lo_hz >= 0.0f && hi_hz > lo_hz && hi_hz <= 22050.0f)) return RDSP_ERR_RANGE;
    rdsp_host::design_bandpass_hz((double)lo_hz, (double)hi_hz, taps);
    return RDSP_OK;
}

int rdsp_gpu_get_taps(rdsp_gpu_t *h, int kind, int index, int16_t *taps, uint32_t n_taps)
{
    if (!h || !taps) return RDSP_ERR_INVALID;
    int16_t *row = taps_row(h, kind, index);
    if (!row || n_taps != RDSP_FIR_TAPS) { h->err = "tap table kind/index/length invalid"; return RDSP_ERR_RANGE; }
    memcpy(taps, row, RDSP_FIR_TAPS * sizeof(int16_t));
    return RDSP_OK;
}

int rdsp_gpu_set_mask(rdsp_gpu_t *h, uint32_t ch_first, uint32_t ch_count, const float *mask512)
{
    if (!h || !mask512) return RDSP_ERR_INVALID;
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    // rows are immutable and shared: the same table installed twice (or on many ranges) is stored once
    std::vector<float> key(mask512, mask512 + 512);
    auto it = h->custom_mask_ids.find(key);
    int id;
    if (it != h->custom_mask_ids.end()) id = it->second;
    else {
        id = (int)(h->masks.size() / 512);
        h->masks.insert(h->masks.end(), mask512, mask512 + 512);
        h->custom_mask_ids.emplace(std::move(key), id);
    }
    for (uint32_t ch = ch_first; ch < ch_first + ch_count; ch++) { h->custom_mask[ch] = id; h->dpar[ch].mask_id = id; }
    h->par_dirty = true;
    return RDSP_OK;
}

int rdsp_gpu_get_mask(rdsp_gpu_t *h, uint32_t ch, float *mask512)
{
    if (!h || !mask512) return RDSP_ERR_INVALID;
    if (ch >= (uint32_t)h->C) { h->err = "channel out of bounds"; return RDSP_ERR_RANGE; }
    memcpy(mask512, &h->masks[(size_t)h->dpar[ch].mask_id * 512], 512 * sizeof(float));
    return RDSP_OK;
}

int rdsp_gpu_read_debug_f32(rdsp_gpu_t *h, uint32_t n_blocks, uint32_t ch_first, uint32_t ch_count, float *out)
{
    if (!h || !out) return RDSP_ERR_INVALID;
    if (!h->d_dbg) { h->err = "handle was created without debug_f32"; return RDSP_ERR_STATE; }
    if (n_blocks == 0 || n_blocks > (uint32_t)h->maxT) { h->err = "n_blocks out of range"; return RDSP_ERR_RANGE; }
    if (ch_count == 0 || (uint64_t)ch_first + ch_count > (uint64_t)h->C) { h->err = "channel range out of bounds"; return RDSP_ERR_RANGE; }
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    const size_t row = 2 * RDSP_BLK;
    CK(cudaMemcpy2D(out, (size_t)ch_count * row * sizeof(float), h->d_dbg + (size_t)ch_first * row,
                    (size_t)h->C * row * sizeof(float), (size_t)ch_count * row * sizeof(float), n_blocks, cudaMemcpyDeviceToHost));
    return RDSP_OK;
}

uint64_t rdsp_gpu_kernel_launches(const rdsp_gpu_t *h) { return h ? h->launches : 0; }
uint64_t rdsp_gpu_graph_replays(const rdsp_gpu_t *h) { return h ? h->graph_replays : 0; }

int rdsp_gpu_profile(rdsp_gpu_t *h, int enable)
{
    if (!h) return RDSP_ERR_INVALID;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    prof_collect(h);
    if (enable && !h->profiling) {
        for (int i = 0; i < KK_COUNT; i++) { h->prof_ms[i] = 0.0; h->prof_n[i] = 0; }
    }
    h->profiling = enable != 0;
    return RDSP_OK;
}

int rdsp_gpu_profile_read(rdsp_gpu_t *h, int max, const char **names, double *ms, uint64_t *launches)
{
    if (!h) return RDSP_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    prof_collect(h);
    for (int i = 0; i < KK_COUNT && i < max; i++) {
        if (names) names[i] = kKernelNames[i];
        if (ms) ms[i] = h->prof_ms[i];
        if (launches) launches[i] = h->prof_n[i];
    }
    return KK_COUNT;
}

}  // extern "C"
