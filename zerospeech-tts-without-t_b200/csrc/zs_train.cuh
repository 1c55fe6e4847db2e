// Host side of the pretrain_AE step (trainer.py:321-332): training-mode forward passes that keep what the backward
// needs, and the backward passes (data-gradient GEMMs through conv_gemm_kernel with transposed weights,
// weight-gradient GEMMs through wgrad_gemm_kernel, element-wise tails in train_kernels.cuh).
// Included at the end of zs_ae.cu (same translation unit: shares Layer / Buf / run_layer / launchers).
#pragma once

// -------------------------------------------------------------------------------------------------
// weight-gradient GEMM launcher
// -------------------------------------------------------------------------------------------------
// K splits fill this many QUARTER waves of SMs.  Measured at B = 32 (pretrain_AE step, graph replay, in-order launches): 8 quarter
// waves 6.42 ms, 4: 6.14, 2: 5.82 - the fp32 atomic adds of the split partial sums cost more than the idle SMs; with the
// launches on the side streams of zs_wgrad_async 1 quarter wave is best (5.26 ms) because concurrent launches fill the rest.
static int g_wgrad_waves = 1;
static int launch_wgrad(const zs_wgrad_desc* d, cudaStream_t st) {
    { static const int w = env_int("ZS_WGRAD_QWAVES", 1); g_wgrad_waves = w < 1 ? 1 : w; }
    ZS_TRY(ensure_device());
    if (!d->dy || !d->x || !d->grad) return fail(ZS_ERR_ARG, "wgrad: null pointer");
    if (d->dy_pitch % 8 || d->x_pitch % 8) return fail(ZS_ERR_ARG, "wgrad: buffer pitches must be multiples of 8");
    if (d->stride != 1 && d->stride != 2) return fail(ZS_ERR_ARG, "wgrad: stride %d", d->stride);
    if (d->stride == 2 && (d->x_rows & 1)) return fail(ZS_ERR_ARG, "wgrad: stride 2 needs an even x_rows");
    if (d->B < 1 || d->T < 1 || d->taps < 1 || d->c_out < 1 || d->c_in < 1) return fail(ZS_ERR_ARG, "wgrad: empty problem");
    const int rows_ps = std::min(d->T, WG_KROWS);
    if (WG_KROWS % rows_ps || d->T % rows_ps) return fail(ZS_ERR_ARG, "wgrad: T %d must be a power of two <= 64 or a multiple of 64", d->T);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.rows_ps = rows_ps; p.nb = WG_KROWS / rows_ps; p.seg_steps = d->T / rows_ps; p.n_groups = (d->B + p.nb - 1) / p.nb;
    p.m_tiles = (d->c_in + 127) / 128; p.n_tiles = (d->c_out + 255) / 256; p.taps = d->taps;     // M = input channels
    p.n_blk_last = (d->c_out - (p.n_tiles - 1) * 256 + 63) / 64;
    const int k_stages = p.n_groups * p.seg_steps, items0 = p.m_tiles * p.n_tiles * p.taps;
    int ksplit = (g_wgrad_waves * g_num_sms / 4 + items0 - 1) / items0;
    // large batches (not measured above): no CTA reduces more than 256 stages while SMs are left to split over
    const int ksplit_deep = std::min((k_stages + 255) / 256, std::max(1, 2 * g_num_sms / items0));
    ksplit = std::max(ksplit, ksplit_deep);
    ksplit = std::max(1, std::min(ksplit, std::max(1, k_stages / 4)));
    // At small batches the kernel is bound by its fp32 atomic adds (one per weight per K split), not by the MMAs:
    // when the gradient is known to be zero and the items fill at least half the SMs on their own, keep the reduction
    // whole and store (measured at B = 32 on one box: 2.35 ms of weight-gradient GEMMs per step -> 2.10 ms).
    static const int direct_mode = env_int("ZS_WGRAD_DIRECT", 8);   // 0 off, n: items >= SMs / n (8 measured best at B = 32: 5.25 ms vs 5.29 at 2)
    p.direct = (d->grad_is_zero && direct_mode > 0 && direct_mode * items0 >= g_num_sms && ksplit_deep == 1) ? 1 : 0;
    if (p.direct) ksplit = 1;
    p.ksplit = ksplit;
    p.grad = d->grad; p.c_in = d->c_in; p.c_in_total = d->c_in_total; p.ci_off = d->ci_off; p.c_out = d->c_out; p.k = d->k; p.tap0 = d->tap0;
    p.a_ch0 = d->dy_ch0; p.a_row0 = d->dy_row0; p.b_ch0 = d->x_ch0; p.b_row0 = d->x_row0; p.stride = d->stride;
    p.ps_c = d->ps_c; p.scale = d->scale;
    {
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->dy_channels), static_cast<cuuint64_t>(d->dy_rows), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->dy_pitch) * 2, static_cast<cuuint64_t>(d->dy_rows) * d->dy_pitch * 2};
        cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows_ps), static_cast<cuuint32_t>(p.nb)};
        ZS_TRY(make_map(&p.tmA, ZS_OPERAND_FP16, const_cast<void*>(d->dy), 3, dims, strides, box));
    }
    if (d->stride == 1) {
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->x_channels), static_cast<cuuint64_t>(d->x_rows), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->x_pitch) * 2, static_cast<cuuint64_t>(d->x_rows) * d->x_pitch * 2};
        cuuint32_t box[3] = {64, static_cast<cuuint32_t>(rows_ps), static_cast<cuuint32_t>(p.nb)};
        ZS_TRY(make_map(&p.tmB, ZS_OPERAND_FP16, const_cast<void*>(d->x), 3, dims, strides, box));
    } else {
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->x_channels), 2, static_cast<cuuint64_t>(d->x_rows / 2), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->x_pitch) * 2, static_cast<cuuint64_t>(d->x_pitch) * 4,
                                 static_cast<cuuint64_t>(d->x_rows) * d->x_pitch * 2};
        cuuint32_t box[4] = {64, 1, static_cast<cuuint32_t>(rows_ps), static_cast<cuuint32_t>(p.nb)};
        ZS_TRY(make_map(&p.tmB, ZS_OPERAND_FP16, const_cast<void*>(d->x), 4, dims, strides, box));
    }
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(wgrad_gemm_kernel), WG_SMEM_BYTES));
    const int total = items0 * ksplit;
    {
        LaunchScope scope(st, KC_GEMM, 2.0 * d->c_out * d->c_in * d->taps * static_cast<double>(d->B) * d->T, "wgrad_gemm_kernel");
        wgrad_gemm_kernel<<<std::min(total, g_num_sms), WG_THREADS, WG_SMEM_BYTES, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_wgrad_cl(const zs_wgrad_desc* d, void* stream) { return launch_wgrad(d, static_cast<cudaStream_t>(stream)); }

// -------------------------------------------------------------------------------------------------
// small helpers
// -------------------------------------------------------------------------------------------------
static ClView cl_view(const Buf& b, int choff = 0) { return ClView{static_cast<__half*>(b.p), b.rows, b.pitch, b.halo, choff}; }
static ClView cl_none() { return ClView{nullptr, 0, 0, 0, 0}; }
static GradSrc gsrc_none() { return GradSrc{nullptr, 0, 0, 0, 0, GS_NONE}; }
static GradSrc gsrc(const Buf& b, int mode, int pad = 0, int choff = 0) {
    return GradSrc{static_cast<const __half*>(b.p), b.rows, b.pitch, choff, pad, mode};
}
static DropSpec drop_spec(float p, uint64_t seed, const uint64_t* seed_dev, int layer, const uint8_t* const* keep, int C, int T) {
    DropSpec d;
    d.p = p; d.seed = seed; d.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev); d.layer = layer;
    d.keep = (p > 0.f && keep) ? keep[layer] : nullptr; d.C = C; d.T = T;
    return d;
}
static DropSpec drop_none() { return DropSpec{0.f, 0ull, nullptr, 0, nullptr, 0, 0}; }

static int launch_combine(CombineParams& p, cudaStream_t st) {
    if (p.C % 2) return fail(ZS_ERR_ARG, "combine: odd channel count %d", p.C);
    const size_t n = static_cast<size_t>(p.B) * p.T * (p.C / 2);
    LaunchScope scope(st, KC_OTHER, 0.0, "combine_fwd_kernel");
    combine_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
static int launch_act_bwd(ActBwdParams& p, cudaStream_t st) {
    if (p.C % 2) return fail(ZS_ERR_ARG, "act_bwd: odd channel count %d", p.C);
    dim3 grid((p.C + 63) / 64, p.B);
    LaunchScope scope(st, KC_OTHER, 0.0, "act_bwd_kernel");
    act_bwd_kernel<<<grid, 256, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
// bias gradient: out[c] += scale * sum over all rows of all segments (halo rows are zero)
static int launch_colsum(const Buf& b, int B, int choff, int C, float scale, float* out, int ps_c, cudaStream_t st, int pitch_override = 0,
                         long long rows_override = 0) {
    const long long n_rows = rows_override ? rows_override : static_cast<long long>(B) * b.rows;
    const int pitch = pitch_override ? pitch_override : b.pitch;
    dim3 grid((C + 63) / 64, static_cast<unsigned>(std::min<long long>(64, (n_rows + 63) / 64)));
    LaunchScope scope(st, KC_OTHER, 0.0, "colsum_kernel");
    colsum_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(b.p), n_rows, pitch, choff, C, scale, out, ps_c);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

// data-gradient GEMM of layer L: reads the zero-padded gradient buffer `dpre` from row 0, writes T_out rows
static int run_dgrad(const Layer& L, const Buf& dpre, int B, int T_out, const Buf* out, float* out_nct, int ps, float ns, cudaStream_t st) {
    if (!L.wt) return fail(ZS_ERR_ARG, "train: handle was not packed with cfg.train = 1");
    Layer v;
    v.w = L.wt; v.m_rows = L.t_rows; v.m_valid = L.t_valid; v.taps = v.w_taps = L.t_taps; v.c_in_pad = L.t_kpad; v.c_in_valid = L.t_kvalid;
    v.bias = nullptr; v.per_spk = 0; v.n_tab = 1;
    ConvExtras ex; ex.no_sat = 1;
    ConvOpts o; o.lrelu = 0; o.in_row0 = 0; o.ex = &ex;
    o.out_mode = out_nct ? OUT_NCT32 : (ps ? OUT_PS : OUT_CL);
    return run_layer(v, ZS_OPERAND_FP16, ns, dpre, B, T_out, out, out_nct, 0, 0, o, st);
}

// weight (+ bias) gradient of layer L: dy = `dpre` (frames at rows dy_halo..), x = the layer's forward input buffer
struct WgradOpts {
    int taps = 1, k = 1, tap0 = 0, stride = 1, x_ch0 = 0, c_in = 0, c_in_total = 0, ci_off = 0, dy_ch0 = 0, c_out = 0, ps_c = 0;
    int pad_left = -1;    // frames of left padding the forward conv saw (default k / 2)
    int x_shift = 0;      // extra row offset (GRU: h_{t-1} / h_{t+1})
};
// Weight (and bias) gradients are leaves of the backward pass: nothing downstream of the data-gradient chain reads them
// before the optimiser.  zs_wgrad_async(1) moves them onto a per-device side stream that forks from the caller's stream
// after the kernel that produced `dpre` (an event record / wait pair, capturable into a CUDA graph), so the small-batch
// step overlaps the latency-bound weight-gradient GEMMs with the data-gradient / element-wise / GRU chain.  Every gradient
// buffer of a backward pass is written once and forward activations are read-only there, so the fork needs no
// write-after-read edges.  The caller joins with zs_wgrad_join(stream) before it reads the gradients.
constexpr int WGRAD_STREAMS_MAX = 4;
struct WgradSide { cudaStream_t s[WGRAD_STREAMS_MAX] = {}; cudaEvent_t fork[WGRAD_STREAMS_MAX] = {}, join[WGRAD_STREAMS_MAX] = {}; };
static WgradSide g_wside[MAX_DEV];
static thread_local int t_wgrad_async = 0, t_wgrad_next = 0;
static thread_local unsigned t_wgrad_pending = 0;          // bit i: side stream i holds unjoined work
static int wgrad_streams() {     // concurrent weight-gradient launches: each fills at most half the SMs (see launch_wgrad)
    static const int n = std::max(1, std::min(WGRAD_STREAMS_MAX, env_int("ZS_WGRAD_STREAMS", 3)));
    return n;
}
extern "C" int zs_wgrad_async(int on) {
    ZS_TRY(ensure_device());
    if (t_wgrad_pending) return fail(ZS_ERR_ARG, "wgrad_async: weight gradients still in flight - call zs_wgrad_join first");
    if (on) {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        WgradSide& w = g_wside[t_dev];
        for (int i = 0; i < WGRAD_STREAMS_MAX && !w.s[WGRAD_STREAMS_MAX - 1]; ++i) {
            CUDA_TRY(cudaStreamCreateWithFlags(&w.s[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&w.fork[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&w.join[i], cudaEventDisableTiming));
        }
    }
    t_wgrad_async = on ? 1 : 0;
    t_wgrad_next = 0;
    return ZS_OK;
}
extern "C" int zs_wgrad_join(void* stream) {
    if (!t_wgrad_pending) return ZS_OK;
    ZS_TRY(ensure_device());
    WgradSide& w = g_wside[t_dev];
    for (int i = 0; i < WGRAD_STREAMS_MAX; ++i)
        if (t_wgrad_pending & (1u << i)) {
            CUDA_TRY(cudaEventRecord(w.join[i], w.s[i]));
            CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), w.join[i], 0));
        }
    t_wgrad_pending = 0;
    return ZS_OK;
}

static int run_wgrad(const Buf& dpre, const Buf& x, int B, int T, float* grad_w, float* grad_b, float inv_scale, const WgradOpts& o,
                     cudaStream_t st) {
    if (t_wgrad_async && !g_prof_on) {       // (per-launch profiling keeps one stream so that its times add up)
        ZS_TRY(ensure_device());
        WgradSide& w = g_wside[t_dev];
        const int i = t_wgrad_next;
        t_wgrad_next = (i + 1) % wgrad_streams();
        if (!w.s[i]) return fail(ZS_ERR_ARG, "wgrad: zs_wgrad_async(1) was called on another device");
        CUDA_TRY(cudaEventRecord(w.fork[i], st));
        CUDA_TRY(cudaStreamWaitEvent(w.s[i], w.fork[i], 0));
        st = w.s[i];
        t_wgrad_pending |= 1u << i;
    }
    zs_wgrad_desc d;
    memset(&d, 0, sizeof(d));
    d.dy = dpre.p; d.dy_rows = dpre.rows; d.dy_pitch = dpre.pitch; d.dy_channels = dpre.pitch; d.dy_ch0 = o.dy_ch0; d.dy_row0 = dpre.halo; d.c_out = o.c_out;
    d.x = x.p; d.x_rows = x.rows; d.x_pitch = x.pitch; d.x_channels = x.pitch; d.x_ch0 = o.x_ch0; d.x_row0 = x.halo - (o.pad_left >= 0 ? o.pad_left : o.k / 2) + o.x_shift; d.c_in = o.c_in; d.stride = o.stride;
    d.B = B; d.T = T; d.taps = o.taps; d.grad = grad_w; d.c_in_total = o.c_in_total ? o.c_in_total : o.c_in; d.ci_off = o.ci_off; d.k = o.k; d.tap0 = o.tap0;
    d.ps_c = o.ps_c; d.scale = inv_scale;
    d.grad_is_zero = 1;       // every weight tensor is written by exactly one call per backward, into zeroed gradients
    ZS_TRY(launch_wgrad(&d, st));
    if (grad_b) ZS_TRY(launch_colsum(dpre, B, o.dy_ch0, o.c_out, inv_scale, grad_b, o.ps_c, st));
    return ZS_OK;
}

// -------------------------------------------------------------------------------------------------
// GRU backward (CUDA-core BPTT) launcher
// -------------------------------------------------------------------------------------------------
static int launch_gru_bptt_cluster(const void* whhT_img, const Buf& gates, const Buf& hbuf, int h_choff, const Buf& dout, int do_choff, int B,
                                   int T, int H, const Buf& dgx, const Buf& dgh, cudaStream_t st) {
    GruBpttParams p;
    p.w_img = whhT_img; p.gates = static_cast<const __half*>(gates.p);
    p.hbuf = static_cast<const __half*>(hbuf.p); p.h_rows = hbuf.rows; p.h_pitch = hbuf.pitch; p.h_choff = h_choff;
    p.dout = static_cast<const __half*>(dout.p); p.do_rows = dout.rows; p.do_pitch = dout.pitch; p.do_choff = do_choff;
    p.dgx = static_cast<__half*>(dgx.p); p.dgh = static_cast<__half*>(dgh.p); p.B = B; p.T = T; p.H = H;
    const int NC = H / GRU_UNITS, n_groups = (B + GRU_NSEQ - 1) / GRU_NSEQ, smem = gb_smem_bytes(H);
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(gru_bptt_cluster_kernel), smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * n_groups * NC);
    cfg.blockDim = dim3(GB_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    LaunchScope scope(st, KC_GRU, 2.0 * 2 * B * static_cast<double>(T) * 3 * H * H, "gru_bptt_cluster_kernel");
    CUDA_TRY(cudaLaunchKernelEx(&cfg, gru_bptt_cluster_kernel, p));
    return ZS_OK;
}

static int launch_gru_bptt(const Buf& gates, const Buf& hbuf, int h_choff, const Buf& dout, int do_choff, const float* const* w_hh,
                           int B, int T, int H, const Buf& dgx, const Buf& dgh, cudaStream_t st, const void* whhT_img = nullptr) {
    if (whhT_img && !env_int("ZS_GRU_BPTT_SIMPLE", 0)) return launch_gru_bptt_cluster(whhT_img, gates, hbuf, h_choff, dout, do_choff, B, T, H, dgx, dgh, st);
    if (H > 1024) return fail(ZS_ERR_ARG, "gru bptt: H %d > 1024", H);
    constexpr int NBG = 4;
    dim3 grid((B + NBG - 1) / NBG, 2);
    const size_t smem = static_cast<size_t>(NBG) * 3 * H * 4;
    LaunchScope scope(st, KC_GRU, 2.0 * 2 * B * static_cast<double>(T) * 3 * H * H, "gru_bptt_simple_kernel");
    gru_bptt_simple_kernel<NBG><<<grid, H, smem, st>>>(static_cast<const __half*>(gates.p), static_cast<const __half*>(hbuf.p), hbuf.rows, hbuf.pitch,
                                                       h_choff, static_cast<const __half*>(dout.p), dout.rows, dout.pitch, do_choff, w_hh[0], w_hh[1],
                                                       B, T, H, static_cast<__half*>(dgx.p), static_cast<__half*>(dgh.p));
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

// forward GRU of a training pass (saves the gates)
static int run_gru_train(void* whh_img, const float* whhT, const float* bhh, const Buf& gx, int B, int T, int H, const Buf& out, int choff,
                         const Buf& gates, cudaStream_t st) {
    if (whh_img) return launch_gru_cluster(whh_img, bhh, gx.p, B, T, H, out.p, out.rows, out.pitch, 0, choff, ZS_OPERAND_FP16, st, gates.p);
    return launch_gru(gx.p, whhT, bhh, B, T, H, out.p, out.rows, out.pitch, 0, choff, ZS_OPERAND_FP16, st, gates.p);
}

// shared GRU backward tail: BPTT, then W_ih / W_hh / bias gradients.  `xin` = the GRU's input buffer (c_in channels from 0),
// `hbuf` holds h_t at channels [h_choff + dir*H, ...), `dout` the gradient of that same region.
static int gru_backward(const Buf& gates, const Buf& xin, int c_in, const Buf& hbuf, int h_choff, const Buf& dout, const float* const* w_hh,
                        int B, int T, int H, const Buf& dgx, const Buf& dgh, float* const* g_w_ih, float* const* g_w_hh,
                        float* const* g_b_ih, float* const* g_b_hh, float inv_scale, cudaStream_t st, const void* whhT_img) {
    ZS_TRY(launch_gru_bptt(gates, hbuf, h_choff, dout, h_choff, w_hh, B, T, H, dgx, dgh, st, whhT_img));
    for (int d = 0; d < 2; ++d) {
        WgradOpts wi; wi.c_out = 3 * H; wi.dy_ch0 = d * 3 * H; wi.c_in = c_in;
        ZS_TRY(run_wgrad(dgx, xin, B, T, g_w_ih[d], g_b_ih[d], inv_scale, wi, st));
        WgradOpts wh; wh.c_out = 3 * H; wh.dy_ch0 = d * 3 * H; wh.c_in = H; wh.x_ch0 = h_choff + d * H; wh.x_shift = d ? 1 : -1;   // h_{t-1} / h_{t+1}
        ZS_TRY(run_wgrad(dgh, hbuf, B, T, g_w_hh[d], g_b_hh[d], inv_scale, wh, st));
    }
    return ZS_OK;
}

static int check_train_T(int T) {
    if (T != 64 && T != 128) return fail(ZS_ERR_ARG, "train: T %d must be 64 or 128 frames (hps seg_len)", T);
    return ZS_OK;
}

// =================================================================================================
// Decoder
// =================================================================================================
struct DecTrainWs {
    Buf actp, x0, x0e, p[3], xh[3], y[3], ye[3], d0, xhA, yA, yAe, d2, xhB, catr, yBe, gx, gates, d5;
    float* stats[5];
    // backward
    Buf dlin, G_d5, dpre5, G2, dgx, dgh, G3, dpre_d4, gsumB, G_d2, dpre_d3, G_yAe, dpre_d2, gsumA, G_d0, dpre_d1, G_x3e;
    Buf dpre_c2[3], gsum_c[3], Gp_p[3], dpre_c1[3], Gp_xe[3], dpre_x0;
    size_t bytes;
};
static DecTrainWs carve_decoder_train(const zs_decoder* h, void* ws, int B, int T8) {
    DecTrainWs w;
    Carver c(ws);
    const int ch = h->cfg.c_h, Tf = 8 * T8;
    w.actp = c.act(B, T8, 0, h->cfg.c_in);
    w.x0 = c.act(B, T8, 1, ch);
    w.x0e = c.act(B, T8, 1, ch);
    for (int i = 0; i < 3; ++i) {
        const int To = T8 << (i + 1);
        w.p[i] = c.act(B, To, 1, ch);
        w.xh[i] = c.act(B, To, 0, ch);
        w.y[i] = c.act(B, To, 1, ch);
        w.ye[i] = c.act(B, To, 1, ch);
    }
    Buf* full[] = {&w.d0, &w.xhA, &w.yA, &w.yAe, &w.d2, &w.xhB, &w.yBe, &w.d5, &w.G_d5, &w.dpre5, &w.G3, &w.dpre_d4, &w.gsumB, &w.G_d2,
                   &w.dpre_d3, &w.G_yAe, &w.dpre_d2, &w.gsumA, &w.G_d0, &w.dpre_d1, &w.G_x3e};
    for (Buf* b : full) *b = c.act(B, Tf, 0, ch);
    w.catr = c.act(B, Tf, 0, 3 * ch);
    w.gx = c.act(B, Tf, 0, 3 * ch, true);
    w.gates = c.act(B, Tf, 0, 4 * ch, true);
    w.G2 = c.act(B, Tf, 0, 3 * ch);
    w.dgx = c.act(B, Tf, 0, 3 * ch, true);
    w.dgh = c.act(B, Tf, 0, 3 * ch, true);
    w.dlin = c.act(B, Tf, 0, h->cfg.c_out);
    for (int i = 0; i < 3; ++i) {
        const int Ti = T8 << i, To = 2 * Ti;
        w.dpre_c2[i] = c.act(B, To, 2, ch);        // zero halo k-1 = 2
        w.gsum_c[i] = c.act(B, To, 0, ch);
        w.Gp_p[i] = c.act(B, To + 2, 0, ch);       // gradient over the padded input of conv2/4/6
        w.dpre_c1[i] = c.act(B, To, 4, ch);        // pixel-shuffle space; viewed as [Ti + 4][2 ch] (zero halo 2) by the GEMMs
        w.Gp_xe[i] = c.act(B, Ti + 2, 0, ch);      // gradient over the padded input of conv1/3/5
    }
    w.dpre_x0 = c.act(B, T8, 0, ch);
    for (int i = 0; i < 5; ++i) w.stats[i] = static_cast<float*>(c.take(static_cast<size_t>(B) * round_up(ch, BM) * 2 * 4));
    w.bytes = c.off;
    return w;
}
extern "C" size_t zs_decoder_train_workspace_bytes(const zs_decoder* h, int B, int T8) {
    if (!h || B < 1 || T8 < 1) return 0;
    return carve_decoder_train(h, nullptr, B, T8).bytes;
}
static Buf ps_view(const Buf& b) {     // [B][rows][pitch] in pixel-shuffle space -> [B][rows/2][2*pitch]
    Buf v = b;
    v.rows = b.rows / 2; v.pitch = 2 * b.pitch; v.halo = b.halo / 2; v.T = b.T / 2;
    return v;
}

extern "C" int zs_decoder_forward_train(zs_decoder* h, const float* enc_act, const int64_t* spk, int B, int T8, float* spec,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    t_zero_pad = 0;          // the training path is reflect-mode only (checked at pack time)
    if (!h || !enc_act || !spk || !spec) return fail(ZS_ERR_ARG, "decoder_forward_train: null argument");
    if (!h->cfg.train) return fail(ZS_ERR_ARG, "decoder_forward_train: handle was packed without cfg.train");
    if (B < 1) return fail(ZS_ERR_ARG, "decoder_forward_train: B %d", B);
    ZS_TRY(check_train_T(8 * T8));
    const zs_decoder_cfg& g = h->cfg;
    DecTrainWs w = carve_decoder_train(h, workspace, B, T8);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "decoder_forward_train: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int op = ZS_OPERAND_FP16, ch = g.c_h, ca = g.c_a;
    const float ns = g.ns;
    auto emb_add = [&](const Buf& x, const Buf* res, int res_mode, const Buf* y, int y_choff, const Buf* ye, const float* emb, const Buf* bc,
                       int bc_choff, int T) -> int {
        CombineParams p;
        memset(&p, 0, sizeof(p));
        p.x = cl_view(x); p.res = res ? cl_view(*res) : cl_none(); p.res_mode = res_mode;
        p.y = y ? cl_view(*y, y_choff) : cl_none();
        p.ye = ye ? cl_view(*ye) : cl_none(); p.emb = emb; p.spk = reinterpret_cast<const long long*>(spk); p.emb_pitch = ch; p.n_spk = ca;
        p.bc = bc ? cl_view(*bc, bc_choff) : cl_none();
        p.drop = drop_none(); p.B = B; p.T = T; p.C = ch;
        return launch_combine(p, st);
    };
    // model/model.py:346 input_emb on the dense activations
    ZS_TRY(launch_pack_nct(enc_act, B, g.c_in, T8, w.actp.p, w.actp.rows, w.actp.pitch, 0, 0, 0, ns, op, 0, st));
    { ConvOpts o; o.lrelu = 0; ZS_TRY(run_layer(h->input_emb, op, ns, w.actp, B, T8, &w.x0, nullptr, 0, 0, o, st)); }
    ZS_TRY(emb_add(w.x0, nullptr, RES_NONE, nullptr, 0, &w.x0e, h->emb[0], nullptr, 0, T8));
    const Buf* xin = &w.x0;
    const Buf* xe = &w.x0e;
    for (int blk = 0; blk < 3; ++blk) {   // :317-331
        const int Ti = T8 << blk;
        ConvExtras e1; e1.post_emb = h->emb[blk]; e1.post_spk = spk; e1.post_pitch = ch; e1.post_n = ca;
        ConvOpts o1; o1.out_mode = OUT_PS; o1.ex = &e1;
        ZS_TRY(run_layer(h->conv[2 * blk], op, ns, *xe, B, Ti, &w.p[blk], nullptr, 0, 0, o1, st));
        ConvExtras e2; e2.stats = w.stats[blk];
        ConvOpts o2; o2.inorm = 1; o2.ex = &e2;
        ZS_TRY(run_layer(h->conv[2 * blk + 1], op, ns, w.p[blk], B, 2 * Ti, &w.xh[blk], nullptr, 0, 0, o2, st));
        ZS_TRY(emb_add(w.xh[blk], xin, RES_UP2, &w.y[blk], 0, &w.ye[blk], h->emb[blk < 2 ? blk + 1 : 3], nullptr, 0, 2 * Ti));
        xin = &w.y[blk];
        xe = &w.ye[blk];
    }
    const int Tf = 8 * T8;
    {   // :350-351 two dense blocks, both conditioned on emb4
        ConvExtras e; e.post_emb = h->emb[3]; e.post_spk = spk; e.post_pitch = ch; e.post_n = ca;
        ConvOpts o; o.ex = &e;
        ConvExtras sA; sA.stats = w.stats[3];
        ConvOpts oA; oA.inorm = 1; oA.ex = &sA;
        ConvExtras sB; sB.stats = w.stats[4];
        ConvOpts oB; oB.inorm = 1; oB.ex = &sB;
        ZS_TRY(run_layer(h->dense[0], op, ns, w.ye[2], B, Tf, &w.d0, nullptr, 0, 0, o, st));
        ZS_TRY(run_layer(h->dense[1], op, ns, w.d0, B, Tf, &w.xhA, nullptr, 0, 0, oA, st));
        ZS_TRY(emb_add(w.xhA, &w.y[2], RES_SAME, &w.yA, 0, &w.yAe, h->emb[3], nullptr, 0, Tf));
        ZS_TRY(run_layer(h->dense[2], op, ns, w.yAe, B, Tf, &w.d2, nullptr, 0, 0, o, st));
        ZS_TRY(run_layer(h->dense[3], op, ns, w.d2, B, Tf, &w.xhB, nullptr, 0, 0, oB, st));
        // out -> catr[:, 0:ch]; out + emb5 -> GRU input; emb5 broadcast -> catr[:, 2ch:3ch] (append_emb, :356)
        ZS_TRY(emb_add(w.xhB, &w.yA, RES_SAME, &w.catr, 0, &w.yBe, h->emb[4], &w.catr, 2 * ch, Tf));
    }
    {   // :352-355
        ConvOpts o; o.lrelu = 0;
        ZS_TRY(run_layer(h->gru_ih, op, ns, w.yBe, B, Tf, &w.gx, nullptr, 0, 0, o, st));
        ZS_TRY(run_gru_train(h->whh_img, h->whhT, h->bhh, w.gx, B, Tf, ch / 2, w.catr, ch, w.gates, st));
    }
    {   // :356-364
        ConvOpts o;
        ZS_TRY(run_layer(h->dense5, op, ns, w.catr, B, Tf, &w.d5, nullptr, 0, 0, o, st));
        ConvOpts f; f.lrelu = 0; f.act = g.output_mask ? ACT_TANH : ACT_SIGMOID; f.out_mode = OUT_NCT32;
        ZS_TRY(run_layer(h->linear, op, ns, w.d5, B, Tf, nullptr, spec, 0, 0, f, st));
    }
    return ZS_OK;
}

extern "C" int zs_decoder_backward(zs_decoder* h, const float* spec, const float* target, const float* d_spec, const int64_t* spk,
                                   int B, int T8, float loss_scale, float* loss, const zs_decoder_weights* grads, float* d_act,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    t_zero_pad = 0;          // the training path is reflect-mode only (checked at pack time)
    if (!h || !spec || !spk || !grads || (!target && !d_spec)) return fail(ZS_ERR_ARG, "decoder_backward: null argument");
    if (target && !loss) return fail(ZS_ERR_ARG, "decoder_backward: the fused L1 loss needs a `loss` output");
    if (!h->cfg.train) return fail(ZS_ERR_ARG, "decoder_backward: handle was packed without cfg.train");
    if (!(loss_scale > 0.f)) return fail(ZS_ERR_ARG, "decoder_backward: loss_scale must be positive");
    ZS_TRY(check_train_T(8 * T8));
    const zs_decoder_cfg& g = h->cfg;
    DecTrainWs w = carve_decoder_train(h, workspace, B, T8);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "decoder_backward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ch = g.c_h, ca = g.c_a, Tf = 8 * T8, H = ch / 2;
    const float ns = g.ns, inv = 1.f / loss_scale;
    auto G = [](const float* p) { return const_cast<float*>(p); };
    const long long* spk_ll = reinterpret_cast<const long long*>(spk);

    // element-wise tail of a layer; see act_bwd_kernel
    auto tail = [&](GradSrc a, GradSrc b, GradSrc r, const Buf* fwd, const float* stats, int lrelu, const float* post_emb, const Buf* dpre,
                    const Buf* gsum, float* demb, int demb_from, int T, int C) -> int {
        ActBwdParams p;
        memset(&p, 0, sizeof(p));
        p.a = a; p.b = b; p.r = r;
        if (fwd) { p.fwd = static_cast<const __half*>(fwd->p); p.f_rows = fwd->rows; p.f_pitch = fwd->pitch; p.f_halo = fwd->halo; }
        p.stats = stats; p.stat_pitch = round_up(ch, BM); p.lrelu = lrelu; p.ns = ns;
        p.post_emb = post_emb; p.spk = spk_ll; p.emb_pitch = ch; p.n_spk = ca;
        p.drop = drop_none();
        if (dpre) { p.dpre = static_cast<__half*>(dpre->p); p.d_rows = dpre->rows; p.d_pitch = dpre->pitch; p.d_halo = dpre->halo; }
        if (gsum) { p.gsum = static_cast<__half*>(gsum->p); p.g_rows = gsum->rows; p.g_pitch = gsum->pitch; }
        p.demb = demb; p.demb_from = demb ? demb_from : 0; p.demb_scale = inv;
        p.B = B; p.T = T; p.C = C;
        return launch_act_bwd(p, st);
    };

    {   // loss + gradient through sigmoid / tanh  ->  dlin (channels-last, c_out channels)
        dim3 grid((Tf + 31) / 32, (w.dlin.pitch + 31) / 32, B), block(32, 8);
        const double n = static_cast<double>(B) * g.c_out * Tf;
        LaunchScope scope(st, KC_OTHER, 0.0, "l1_loss_bwd_kernel");
        if (target)
            l1_loss_bwd_kernel<<<grid, block, 0, st>>>(spec, target, g.c_out, Tf, static_cast<__half*>(w.dlin.p), w.dlin.rows, w.dlin.pitch,
                                                       static_cast<float>(loss_scale / n), static_cast<float>(1.0 / n), g.output_mask, loss);
        else
            dspec_bwd_kernel<<<grid, block, 0, st>>>(spec, d_spec, g.c_out, Tf, static_cast<__half*>(w.dlin.p), w.dlin.rows, w.dlin.pitch, loss_scale,
                                                     g.output_mask);
        CUDA_TRY(cudaGetLastError());
    }
    {   // linear (:360)
        WgradOpts o; o.c_out = g.c_out; o.c_in = ch;
        ZS_TRY(run_wgrad(w.dlin, w.d5, B, Tf, G(grads->linear_w), G(grads->linear_b), inv, o, st));
        ZS_TRY(run_dgrad(h->linear, w.dlin, B, Tf, &w.G_d5, nullptr, 0, ns, st));
    }
    {   // dense5 + lrelu (:358-359) over cat([out, rnn, emb5])
        ZS_TRY(tail(gsrc(w.G_d5, GS_PADDED), gsrc_none(), gsrc_none(), &w.d5, nullptr, 1, nullptr, &w.dpre5, nullptr, nullptr, 0, Tf, ch));
        WgradOpts o; o.c_out = ch; o.c_in = 3 * ch;
        ZS_TRY(run_wgrad(w.dpre5, w.catr, B, Tf, G(grads->dense5_w), G(grads->dense5_b), inv, o, st));
        ZS_TRY(run_dgrad(h->dense5, w.dpre5, B, Tf, &w.G2, nullptr, 0, ns, st));
        // appended emb5 channels: d emb5 += sum_t G2[:, 2ch:3ch]
        ZS_TRY(tail(gsrc(w.G2, GS_PADDED, 0, 2 * ch), gsrc_none(), gsrc_none(), nullptr, nullptr, 0, nullptr, nullptr, nullptr, G(grads->emb[4]), 1, Tf, ch));
    }
    {   // bi-GRU on out + emb5 (:352-355)
        float* gwi[2] = {G(grads->gru_w_ih[0]), G(grads->gru_w_ih[1])};
        float* gwh[2] = {G(grads->gru_w_hh[0]), G(grads->gru_w_hh[1])};
        float* gbi[2] = {G(grads->gru_b_ih[0]), G(grads->gru_b_ih[1])};
        float* gbh[2] = {G(grads->gru_b_hh[0]), G(grads->gru_b_hh[1])};
        ZS_TRY(gru_backward(w.gates, w.yBe, ch, w.catr, ch, w.G2, h->w_hh, B, Tf, H, w.dgx, w.dgh, gwi, gwh, gbi, gbh, inv, st, h->whhT_img));
        ZS_TRY(run_dgrad(h->gru_ih, w.dgx, B, Tf, &w.G3, nullptr, 0, ns, st));
    }
    {   // dense block 2 (:351): out = IN(lrelu(dense4(lrelu(dense3(yA + e4)) + e4))) + yA
        ZS_TRY(tail(gsrc(w.G2, GS_PADDED, 0, 0), gsrc(w.G3, GS_PADDED), gsrc_none(), &w.xhB, w.stats[4], 1, nullptr, &w.dpre_d4, &w.gsumB,
                    G(grads->emb[4]), 2, Tf, ch));
        WgradOpts o; o.c_out = ch; o.c_in = ch;
        ZS_TRY(run_wgrad(w.dpre_d4, w.d2, B, Tf, G(grads->dense_w[3]), G(grads->dense_b[3]), inv, o, st));
        ZS_TRY(run_dgrad(h->dense[3], w.dpre_d4, B, Tf, &w.G_d2, nullptr, 0, ns, st));
        ZS_TRY(tail(gsrc(w.G_d2, GS_PADDED), gsrc_none(), gsrc_none(), &w.d2, nullptr, 1, h->emb[3], &w.dpre_d3, nullptr, G(grads->emb[3]), 3, Tf, ch));
        ZS_TRY(run_wgrad(w.dpre_d3, w.yAe, B, Tf, G(grads->dense_w[2]), G(grads->dense_b[2]), inv, o, st));
        ZS_TRY(run_dgrad(h->dense[2], w.dpre_d3, B, Tf, &w.G_yAe, nullptr, 0, ns, st));
    }
    {   // dense block 1 (:350)
        ZS_TRY(tail(gsrc(w.G_yAe, GS_PADDED), gsrc_none(), gsrc(w.gsumB, GS_SAME), &w.xhA, w.stats[3], 1, nullptr, &w.dpre_d2, &w.gsumA,
                    G(grads->emb[3]), 1, Tf, ch));
        WgradOpts o; o.c_out = ch; o.c_in = ch;
        ZS_TRY(run_wgrad(w.dpre_d2, w.d0, B, Tf, G(grads->dense_w[1]), G(grads->dense_b[1]), inv, o, st));
        ZS_TRY(run_dgrad(h->dense[1], w.dpre_d2, B, Tf, &w.G_d0, nullptr, 0, ns, st));
        ZS_TRY(tail(gsrc(w.G_d0, GS_PADDED), gsrc_none(), gsrc_none(), &w.d0, nullptr, 1, h->emb[3], &w.dpre_d1, nullptr, G(grads->emb[3]), 3, Tf, ch));
        ZS_TRY(run_wgrad(w.dpre_d1, w.ye[2], B, Tf, G(grads->dense_w[0]), G(grads->dense_b[0]), inv, o, st));
        ZS_TRY(run_dgrad(h->dense[0], w.dpre_d1, B, Tf, &w.G_x3e, nullptr, 0, ns, st));
    }
    // conv blocks (:317-331), last to first.  Block output y feeds (y + e_next) -> next layer [source A] and the next
    // block's / dense block's residual [source R].
    GradSrc next_a = gsrc(w.G_x3e, GS_PADDED, 0), next_r = gsrc(w.gsumA, GS_SAME);
    for (int blk = 2; blk >= 0; --blk) {
        const int Ti = T8 << blk, To = 2 * Ti;
        const float* e_next = grads->emb[blk < 2 ? blk + 1 : 3];
        // conv2/4/6: IN + lrelu, k = 3 on p (pixel-shuffle output + e)
        ZS_TRY(tail(next_a, gsrc_none(), next_r, &w.xh[blk], w.stats[blk], 1, nullptr, &w.dpre_c2[blk], &w.gsum_c[blk], G(e_next), 1, To, ch));
        WgradOpts o2; o2.c_out = ch; o2.c_in = ch; o2.taps = 3; o2.k = 3;
        ZS_TRY(run_wgrad(w.dpre_c2[blk], w.p[blk], B, To, G(grads->conv_w[2 * blk + 1]), G(grads->conv_b[2 * blk + 1]), inv, o2, st));
        ZS_TRY(run_dgrad(h->conv[2 * blk + 1], w.dpre_c2[blk], B, To + 2, &w.Gp_p[blk], nullptr, 0, ns, st));
        // conv1/3/5: lrelu, pixel shuffle, + e: the tail runs in pixel-shuffle space (To frames x ch channels)
        ZS_TRY(tail(gsrc(w.Gp_p[blk], GS_PADDED, 1), gsrc_none(), gsrc_none(), &w.p[blk], nullptr, 1, h->emb[blk], &w.dpre_c1[blk], nullptr,
                    G(grads->emb[blk]), 3, To, ch));
        const Buf dv = ps_view(w.dpre_c1[blk]);   // [Ti + 4][2 ch], zero halo 2
        const Buf& xe = blk == 0 ? w.x0e : w.ye[blk - 1];
        WgradOpts o1; o1.c_out = 2 * ch; o1.c_in = ch; o1.taps = 3; o1.k = 3; o1.ps_c = ch;
        ZS_TRY(run_wgrad(dv, xe, B, Ti, G(grads->conv_w[2 * blk]), G(grads->conv_b[2 * blk]), inv, o1, st));
        ZS_TRY(run_dgrad(h->conv[2 * blk], dv, B, Ti + 2, &w.Gp_xe[blk], nullptr, 0, ns, st));
        next_a = gsrc(w.Gp_xe[blk], GS_PADDED, 1);
        next_r = gsrc(w.gsum_c[blk], GS_UP2);
    }
    {   // x0 = input_emb(enc_act) (:346): x0 + e1 -> conv1 [A], x0 up-sampled -> block 1 residual [R]
        ZS_TRY(tail(next_a, gsrc_none(), next_r, nullptr, nullptr, 0, nullptr, &w.dpre_x0, nullptr, G(grads->emb[0]), 1, T8, ch));
        WgradOpts o; o.c_out = ch; o.c_in = g.c_in;
        ZS_TRY(run_wgrad(w.dpre_x0, w.actp, B, T8, G(grads->input_emb_w), G(grads->input_emb_b), inv, o, st));
        if (d_act) ZS_TRY(run_dgrad(h->input_emb, w.dpre_x0, B, T8, nullptr, d_act, 0, ns, st));
    }
    return ZS_OK;
}

// =================================================================================================
// Encoder
// =================================================================================================
struct EncTrainWs {
    Buf xp, cat, xh[4], a[7], d0, xhA, dA, d2, xhB, catr, gx, gates;
    int T[4];
    float* stats[6];
    Buf dlog, G_catr, dgx, dgh, G3, dpre_d4, gsumB, G_d2, dpre_d3, G_dA, dpre_d2, gsumA, G_d0, dpre_d1, G_a6;
    Buf dpre_s2[3], gs_a[3], Gp_odd[3], dpre_c[3], Gp_even[3], dpre_c2, G_cat, dpre_bank;
    float* dact_scaled;
    size_t bytes;
};
static EncTrainWs carve_encoder_train(const zs_encoder* h, void* ws, int B, int T) {
    EncTrainWs w;
    Carver c(ws);
    const zs_encoder_cfg& g = h->cfg;
    const int h2 = g.c_h2;
    w.T[0] = T; w.T[1] = T / 2; w.T[2] = T / 4; w.T[3] = T / 8;
    const int T8 = w.T[3];
    w.xp = c.act(B, T, 3, g.c_in);
    w.cat = c.act(B, T, 0, 7 * g.c_h1 + g.c_in);
    for (int i = 0; i < 4; ++i) w.xh[i] = c.act(B, w.T[i], 0, h2);
    w.a[0] = c.act(B, w.T[0], 2, h2);
    w.a[1] = c.act(B, w.T[0], 2, h2);
    w.a[2] = c.act(B, w.T[1], 2, h2);
    w.a[3] = c.act(B, w.T[1], 2, h2);
    w.a[4] = c.act(B, w.T[2], 2, h2);
    w.a[5] = c.act(B, w.T[2], 2, h2);
    w.a[6] = c.act(B, w.T[3], 0, h2);
    Buf* small[] = {&w.d0, &w.xhA, &w.dA, &w.d2, &w.xhB, &w.G3, &w.dpre_d4, &w.gsumB, &w.G_d2, &w.dpre_d3, &w.G_dA, &w.dpre_d2, &w.gsumA,
                    &w.G_d0, &w.dpre_d1, &w.G_a6};
    for (Buf* b : small) *b = c.act(B, T8, 0, h2);
    w.catr = c.act(B, T8, 0, h2 + 2 * g.c_h3);
    w.gx = c.act(B, T8, 0, 6 * g.c_h3, true);
    w.gates = c.act(B, T8, 0, 8 * g.c_h3, true);
    w.dlog = c.act(B, T8, 0, h->n_out);
    w.G_catr = c.act(B, T8, 0, h2 + 2 * g.c_h3);
    w.dgx = c.act(B, T8, 0, 6 * g.c_h3, true);
    w.dgh = c.act(B, T8, 0, 6 * g.c_h3, true);
    for (int j = 0; j < 3; ++j) {
        w.dpre_s2[j] = c.act(B, w.T[j + 1], 2, h2);     // stride-2 conv output gradient, zero halo p = 2
        w.gs_a[j] = c.act(B, w.T[j + 1], 0, h2);        // total gradient of block output a[2j+2]
        w.Gp_odd[j] = c.act(B, w.T[j] + 4, 0, h2);      // gradient over the padded input of the stride-2 conv
        w.dpre_c[j] = c.act(B, w.T[j], 4, h2);          // k = 5 conv output gradient, zero halo k-1 = 4
        w.Gp_even[j] = c.act(B, w.T[j] + 4, 0, h2);     // gradient over the padded input of conv3/5/7
    }
    w.dpre_c2 = c.act(B, T, 0, h2);
    w.G_cat = c.act(B, T, 0, 7 * g.c_h1);
    w.dpre_bank = c.act(B, T, 0, 7 * g.c_h1);
    for (int i = 0; i < 6; ++i) w.stats[i] = static_cast<float*>(c.take(static_cast<size_t>(B) * round_up(h2, BM) * 2 * 4));
    w.bytes = c.off;
    return w;
}
extern "C" size_t zs_encoder_train_workspace_bytes(const zs_encoder* h, int B, int T) {
    if (!h || B < 1 || T < 1) return 0;
    return carve_encoder_train(h, nullptr, B, T).bytes;
}

extern "C" int zs_encoder_forward_train(zs_encoder* h, const float* x, int B, int T, const float* gumbel_noise, float dropout_p,
                                        uint64_t dropout_seed, const uint64_t* dropout_seed_dev, const uint8_t* const* keep_masks, float* logits, float* act,
                                        int32_t* unit_ids, void* workspace, size_t workspace_bytes, void* stream) {
    t_zero_pad = 0;          // the training path is reflect-mode only (checked at pack time)
    if (!h || !x || !logits || !gumbel_noise) return fail(ZS_ERR_ARG, "encoder_forward_train: null argument");
    if (!h->cfg.train) return fail(ZS_ERR_ARG, "encoder_forward_train: handle was packed without cfg.train");
    if (B < 1) return fail(ZS_ERR_ARG, "encoder_forward_train: B %d", B);
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail(ZS_ERR_ARG, "encoder_forward_train: dropout_p %f", dropout_p);
    ZS_TRY(check_train_T(T));
    const zs_encoder_cfg& g = h->cfg;
    EncTrainWs w = carve_encoder_train(h, workspace, B, T);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "encoder_forward_train: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int op = ZS_OPERAND_FP16, h2 = g.c_h2;
    const float ns = g.ns;
    auto drop_res = [&](const Buf& xh, int layer, const Buf* res, int res_mode, const Buf& y, int Tl) -> int {
        CombineParams p;
        memset(&p, 0, sizeof(p));
        p.x = cl_view(xh); p.res = res ? cl_view(*res) : cl_none(); p.res_mode = res_mode;
        p.y = cl_view(y); p.ye = cl_none(); p.bc = cl_none();
        p.drop = drop_spec(dropout_p, dropout_seed, dropout_seed_dev, layer, keep_masks, h2, Tl);
        p.B = B; p.T = Tl; p.C = h2;
        return launch_combine(p, st);
    };
    ZS_TRY(launch_pack_x_dual(x, ZS_X_F32, ZS_X_NCT, B, g.c_in, T, w.xp.p, w.xp.rows, w.xp.pitch, 3, w.cat.p, w.cat.rows, w.cat.pitch, 7 * g.c_h1, ns, op, st));
    if (h->bank_merged) {
        ConvOpts o; o.bank = 1;
        ZS_TRY(run_layer(h->bank[0], op, ns, w.xp, B, T, &w.cat, nullptr, 0, 0, o, st));
    } else {
        for (int i = 0; i < 7; ++i) {
            ConvOpts o; o.out_choff = i * g.c_h1;
            ZS_TRY(run_layer(h->bank[i], op, ns, w.xp, B, T, &w.cat, nullptr, 0, 0, o, st));
        }
    }
    {   // :447 conv2 -> lrelu -> IN -> dropout
        ConvExtras e; e.stats = w.stats[0];
        ConvOpts o; o.inorm = 1; o.ex = &e;
        ZS_TRY(run_layer(h->conv[0], op, ns, w.cat, B, T, &w.xh[0], nullptr, 0, 0, o, st));
        ZS_TRY(drop_res(w.xh[0], 0, nullptr, RES_NONE, w.a[0], T));
    }
    for (int blk = 0; blk < 3; ++blk) {   // :448-450
        const Buf& xin = w.a[2 * blk];
        ConvOpts o1;
        ZS_TRY(run_layer(h->conv[1 + 2 * blk], op, ns, xin, B, w.T[blk], &w.a[2 * blk + 1], nullptr, 0, 0, o1, st));
        ConvExtras e; e.stats = w.stats[blk + 1];
        ConvOpts o2; o2.stride = 2; o2.inorm = 1; o2.ex = &e;
        ZS_TRY(run_layer(h->conv[2 + 2 * blk], op, ns, w.a[2 * blk + 1], B, w.T[blk + 1], &w.xh[blk + 1], nullptr, 0, 0, o2, st));
        ZS_TRY(drop_res(w.xh[blk + 1], blk + 1, &xin, RES_AVG2, w.a[2 * blk + 2], w.T[blk + 1]));
    }
    const int T8 = w.T[3];
    {   // :452-453
        ConvOpts o;
        ConvExtras eA; eA.stats = w.stats[4];
        ConvOpts oA; oA.inorm = 1; oA.ex = &eA;
        ConvExtras eB; eB.stats = w.stats[5];
        ConvOpts oB; oB.inorm = 1; oB.ex = &eB;
        ZS_TRY(run_layer(h->dense[0], op, ns, w.a[6], B, T8, &w.d0, nullptr, 0, 0, o, st));
        ZS_TRY(run_layer(h->dense[1], op, ns, w.d0, B, T8, &w.xhA, nullptr, 0, 0, oA, st));
        ZS_TRY(drop_res(w.xhA, 4, &w.a[6], RES_SAME, w.dA, T8));
        ZS_TRY(run_layer(h->dense[2], op, ns, w.dA, B, T8, &w.d2, nullptr, 0, 0, o, st));
        ZS_TRY(run_layer(h->dense[3], op, ns, w.d2, B, T8, &w.xhB, nullptr, 0, 0, oB, st));
        ZS_TRY(drop_res(w.xhB, 5, &w.dA, RES_SAME, w.catr, T8));
    }
    {   // :454-455
        ConvOpts o; o.lrelu = 0; o.c_in_valid = h2;
        ZS_TRY(run_layer(h->gru_ih, op, ns, w.catr, B, T8, &w.gx, nullptr, 0, 0, o, st));
        ZS_TRY(run_gru_train(h->whh_img, h->whhT, h->bhh, w.gx, B, T8, g.c_h3, w.catr, h2, w.gates, st));
    }
    {
        ConvOpts o; o.lrelu = 0; o.out_mode = OUT_NCT32;
        ZS_TRY(run_layer(h->linear, op, ns, w.catr, B, T8, nullptr, logits, 0, 0, o, st));
    }
    ZS_TRY(launch_onehot(logits, gumbel_noise, nullptr, B, g.enc_size, T8, act, unit_ids, st));
    return ZS_OK;
}

extern "C" int zs_encoder_backward(zs_encoder* h, const float* d_act, float d_act_scale, const float* gumbel_noise, const float* logits, int B,
                                   int T, float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev, const uint8_t* const* keep_masks, float loss_scale,
                                   const zs_encoder_weights* grads, void* workspace, size_t workspace_bytes, void* stream) {
    t_zero_pad = 0;          // the training path is reflect-mode only (checked at pack time)
    if (!h || !d_act || !gumbel_noise || !logits || !grads) return fail(ZS_ERR_ARG, "encoder_backward: null argument");
    if (!h->cfg.train) return fail(ZS_ERR_ARG, "encoder_backward: handle was packed without cfg.train");
    if (!(loss_scale > 0.f) || !(d_act_scale > 0.f)) return fail(ZS_ERR_ARG, "encoder_backward: scales must be positive");
    ZS_TRY(check_train_T(T));
    const zs_encoder_cfg& g = h->cfg;
    EncTrainWs w = carve_encoder_train(h, workspace, B, T);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "encoder_backward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int h2 = g.c_h2, H = g.c_h3, T8 = w.T[3];
    const float ns = g.ns, inv = 1.f / loss_scale;
    auto G = [](const float* p) { return const_cast<float*>(p); };
    auto tail = [&](GradSrc a, GradSrc b, GradSrc r, const Buf* fwd, const float* stats, int drop_layer, const Buf* dpre, const Buf* gsum, int Tl,
                    int C, int f_choff = 0) -> int {
        ActBwdParams p;
        memset(&p, 0, sizeof(p));
        p.a = a; p.b = b; p.r = r;
        if (fwd) { p.fwd = static_cast<const __half*>(fwd->p); p.f_rows = fwd->rows; p.f_pitch = fwd->pitch; p.f_halo = fwd->halo; p.f_choff = f_choff; }
        p.stats = stats; p.stat_pitch = round_up(h2, BM); p.lrelu = 1; p.ns = ns;
        p.drop = drop_layer >= 0 ? drop_spec(dropout_p, dropout_seed, dropout_seed_dev, drop_layer, keep_masks, h2, Tl) : drop_none();
        if (dpre) { p.dpre = static_cast<__half*>(dpre->p); p.d_rows = dpre->rows; p.d_pitch = dpre->pitch; p.d_halo = dpre->halo; }
        if (gsum) { p.gsum = static_cast<__half*>(gsum->p); p.g_rows = gsum->rows; p.g_pitch = gsum->pitch; }
        p.B = B; p.T = Tl; p.C = C;
        return launch_act_bwd(p, st);
    };
    {   // straight-through Gumbel softmax (:461-464) -> gradient of the logits, channels-last
        const int C = g.enc_size;
        const size_t smem = static_cast<size_t>(2) * C * (T8 + 1) * 4;
        if (smem > 200 * 1024) return fail(ZS_ERR_ARG, "encoder_backward: enc_size %d x T8 %d does not fit shared memory", C, T8);
        if (smem > 48 * 1024) ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(gumbel_st_bwd_kernel), 200 * 1024));
        LaunchScope scope(st, KC_OTHER, 0.0, "gumbel_st_bwd_kernel");
        gumbel_st_bwd_kernel<<<B, 512, smem, st>>>(logits, gumbel_noise, d_act, C, T8, 10.f /* 1 / temperature 0.1 */,
                                                   10.f * (loss_scale / d_act_scale), static_cast<__half*>(w.dlog.p), w.dlog.rows, w.dlog.pitch);
        CUDA_TRY(cudaGetLastError());
    }
    {   // linear on cat([out, rnn])
        WgradOpts o; o.c_out = h->n_out; o.c_in = h2 + 2 * H;
        ZS_TRY(run_wgrad(w.dlog, w.catr, B, T8, G(grads->linear_w), G(grads->linear_b), inv, o, st));
        ZS_TRY(run_dgrad(h->linear, w.dlog, B, T8, &w.G_catr, nullptr, 0, ns, st));
    }
    {   // bi-GRU
        float* gwi[2] = {G(grads->gru_w_ih[0]), G(grads->gru_w_ih[1])};
        float* gwh[2] = {G(grads->gru_w_hh[0]), G(grads->gru_w_hh[1])};
        float* gbi[2] = {G(grads->gru_b_ih[0]), G(grads->gru_b_ih[1])};
        float* gbh[2] = {G(grads->gru_b_hh[0]), G(grads->gru_b_hh[1])};
        ZS_TRY(gru_backward(w.gates, w.catr, h2, w.catr, h2, w.G_catr, h->w_hh, B, T8, H, w.dgx, w.dgh, gwi, gwh, gbi, gbh, inv, st, h->whhT_img));
        ZS_TRY(run_dgrad(h->gru_ih, w.dgx, B, T8, &w.G3, nullptr, 0, ns, st));
    }
    WgradOpts od; od.c_out = h2; od.c_in = h2;
    {   // dense block 2: catr[:, :h2] = drop6(IN(lrelu(dense4(lrelu(dense3(dA)))))) + dA
        ZS_TRY(tail(gsrc(w.G_catr, GS_PADDED, 0, 0), gsrc(w.G3, GS_PADDED), gsrc_none(), &w.xhB, w.stats[5], 5, &w.dpre_d4, &w.gsumB, T8, h2));
        ZS_TRY(run_wgrad(w.dpre_d4, w.d2, B, T8, G(grads->dense_w[3]), G(grads->dense_b[3]), inv, od, st));
        ZS_TRY(run_dgrad(h->dense[3], w.dpre_d4, B, T8, &w.G_d2, nullptr, 0, ns, st));
        ZS_TRY(tail(gsrc(w.G_d2, GS_PADDED), gsrc_none(), gsrc_none(), &w.d2, nullptr, -1, &w.dpre_d3, nullptr, T8, h2));
        ZS_TRY(run_wgrad(w.dpre_d3, w.dA, B, T8, G(grads->dense_w[2]), G(grads->dense_b[2]), inv, od, st));
        ZS_TRY(run_dgrad(h->dense[2], w.dpre_d3, B, T8, &w.G_dA, nullptr, 0, ns, st));
    }
    {   // dense block 1
        ZS_TRY(tail(gsrc(w.G_dA, GS_PADDED), gsrc_none(), gsrc(w.gsumB, GS_SAME), &w.xhA, w.stats[4], 4, &w.dpre_d2, &w.gsumA, T8, h2));
        ZS_TRY(run_wgrad(w.dpre_d2, w.d0, B, T8, G(grads->dense_w[1]), G(grads->dense_b[1]), inv, od, st));
        ZS_TRY(run_dgrad(h->dense[1], w.dpre_d2, B, T8, &w.G_d0, nullptr, 0, ns, st));
        ZS_TRY(tail(gsrc(w.G_d0, GS_PADDED), gsrc_none(), gsrc_none(), &w.d0, nullptr, -1, &w.dpre_d1, nullptr, T8, h2));
        ZS_TRY(run_wgrad(w.dpre_d1, w.a[6], B, T8, G(grads->dense_w[0]), G(grads->dense_b[0]), inv, od, st));
        ZS_TRY(run_dgrad(h->dense[0], w.dpre_d1, B, T8, &w.G_a6, nullptr, 0, ns, st));
    }
    // conv blocks, last to first: a[2j+2] = drop(IN(lrelu(conv_s2(a[2j+1])))) + avgpool(a[2j]),  a[2j+1] = lrelu(conv(a[2j]))
    GradSrc next_a = gsrc(w.G_a6, GS_PADDED, 0), next_r = gsrc(w.gsumA, GS_SAME);
    for (int j = 2; j >= 0; --j) {
        const int Ti = w.T[j], To = w.T[j + 1];
        ZS_TRY(tail(next_a, gsrc_none(), next_r, &w.xh[j + 1], w.stats[j + 1], j + 1, &w.dpre_s2[j], &w.gs_a[j], To, h2));
        WgradOpts o2; o2.c_out = h2; o2.c_in = h2; o2.taps = 5; o2.k = 5; o2.stride = 2;
        ZS_TRY(run_wgrad(w.dpre_s2[j], w.a[2 * j + 1], B, To, G(grads->conv_w[2 + 2 * j]), G(grads->conv_b[2 + 2 * j]), inv, o2, st));
        ZS_TRY(run_dgrad(h->conv[2 + 2 * j], w.dpre_s2[j], B, To + 2, &w.Gp_odd[j], nullptr, 1, ns, st));   // rows 2w'+r of the padded input
        ZS_TRY(tail(gsrc(w.Gp_odd[j], GS_PADDED, 2), gsrc_none(), gsrc_none(), &w.a[2 * j + 1], nullptr, -1, &w.dpre_c[j], nullptr, Ti, h2));
        WgradOpts o1; o1.c_out = h2; o1.c_in = h2; o1.taps = 5; o1.k = 5;
        ZS_TRY(run_wgrad(w.dpre_c[j], w.a[2 * j], B, Ti, G(grads->conv_w[1 + 2 * j]), G(grads->conv_b[1 + 2 * j]), inv, o1, st));
        ZS_TRY(run_dgrad(h->conv[1 + 2 * j], w.dpre_c[j], B, Ti + 4, &w.Gp_even[j], nullptr, 0, ns, st));
        next_a = gsrc(w.Gp_even[j], GS_PADDED, 2);
        next_r = gsrc(w.gs_a[j], GS_AVG2);
    }
    {   // conv2 (k = 1 on cat([bank, x])) -> IN -> drop1 ; x takes no gradient, so only the bank channels are propagated
        ZS_TRY(tail(next_a, gsrc_none(), next_r, &w.xh[0], w.stats[0], 0, &w.dpre_c2, nullptr, T, h2));
        WgradOpts o; o.c_out = h2; o.c_in = 7 * g.c_h1 + g.c_in;
        ZS_TRY(run_wgrad(w.dpre_c2, w.cat, B, T, G(grads->conv_w[0]), G(grads->conv_b[0]), inv, o, st));
        ZS_TRY(run_dgrad(h->conv[0], w.dpre_c2, B, T, &w.G_cat, nullptr, 0, ns, st));
    }
    {   // conv bank (:441-446): cat[:, :7 c_h1] = lrelu(conv_k(x))
        ZS_TRY(tail(gsrc(w.G_cat, GS_PADDED), gsrc_none(), gsrc_none(), &w.cat, nullptr, -1, &w.dpre_bank, nullptr, T, 7 * g.c_h1));
        for (int i = 0; i < 7; ++i) {
            const int k = i + 1;
            WgradOpts o; o.c_out = g.c_h1; o.dy_ch0 = i * g.c_h1; o.c_in = g.c_in; o.taps = k; o.k = k;   // xp has halo 3; kernel k pads k/2 on the left
            ZS_TRY(run_wgrad(w.dpre_bank, w.xp, B, T, G(grads->conv1s_w[i]), G(grads->conv1s_b[i]), inv, o, st));
        }
    }
    return ZS_OK;
}

// =================================================================================================
// optimiser
// =================================================================================================
extern "C" int zs_grad_sqnorm(const float* g, size_t n, float* out, void* stream) {
    if (!g || !out) return fail(ZS_ERR_ARG, "grad_sqnorm: null argument");
    ZS_TRY(ensure_device());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "sqnorm_kernel");
    sqnorm_kernel<<<std::min<size_t>(4 * g_num_sms, (n + 1023) / 1024 + 1), 1024, 0, st>>>(g, n, out);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, const float* sqnorm,
                            float grad_mult, float max_norm, float lr, float beta1, float beta2, float eps, int step, const float* bias_corr_dev,
                            int* skipped, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !sqnorm) return fail(ZS_ERR_ARG, "adam_step: null argument");
    if (step < 1) return fail(ZS_ERR_ARG, "adam_step: step %d must be >= 1", step);
    ZS_TRY(ensure_device());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step)), bc2 = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    LaunchScope scope(st, KC_OTHER, 0.0, "adam_kernel");
    adam_kernel<<<std::min<size_t>(8 * g_num_sms, (n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, sqnorm, grad_mult, max_norm, lr,
                                                                                  beta1, beta2, eps, bc1, bc2, bias_corr_dev, skipped);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

extern "C" int zs_train_meta_begin(void* meta, uint64_t seed_salt, void* stream) {
    if (!meta) return fail(ZS_ERR_ARG, "train_meta_begin: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "train_meta_begin_kernel");
    train_meta_begin_kernel<<<1, 1, 0, st>>>(static_cast<unsigned long long*>(meta), seed_salt);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_train_meta_commit(void* meta, const float* sqnorm_a, const float* sqnorm_b, float beta1, float beta2, int* skipped,
                                    void* stream) {
    if (!meta || !sqnorm_a) return fail(ZS_ERR_ARG, "train_meta_commit: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "train_meta_commit_kernel");
    train_meta_commit_kernel<<<1, 1, 0, st>>>(static_cast<float*>(meta), sqnorm_a, sqnorm_b, beta1, beta2, skipped);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
