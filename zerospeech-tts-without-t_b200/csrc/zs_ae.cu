// libzsae.so - C ABI (include/zs_ae.h) over the sm_100a kernels of the autoencoder hot path.
// Host side: weight packing, workspace carving, TMA descriptors, layer sequencing.
#include "../../include/zs_ae.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "conv_gemm.cuh"
#include "kernels.cuh"
#include "gru_cluster.cuh"
#include "gru_wide.cuh"
#include "train_kernels.cuh"
#include "gru_bptt_cluster.cuh"
#include "wgrad_gemm.cuh"

using namespace zs;

// -------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) return fail(ZS_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define ZS_TRY(expr)              \
    do {                          \
        int r_ = (expr);          \
        if (r_ != ZS_OK) return r_; \
    } while (0)

extern "C" const char* zs_last_error(void) { return g_err; }
extern "C" int zs_version(void) { return 200; }

// -------------------------------------------------------------------------------------------------
// driver entry point for TMA descriptors (no -lcuda link dependency)
// -------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

// Per-device state: SM count, the fp16 saturation counter the GEMM epilogues raise, and which kernels already carry
// their dynamic-shared-memory opt-in (cudaFuncSetAttribute is per device, not per process).
constexpr int MAX_DEV = 64;
struct DevState {
    int num_sms = 0;
    unsigned int* sat = nullptr;          // device word: clamped-to-+-65504 events since the last reset
};
static DevState g_dev[MAX_DEV];
static std::mutex g_dev_mu;
static std::map<std::pair<int, const void*>, int> g_smem_attr;
static thread_local int t_dev = 0;        // device of the call in flight on this host thread (set by ensure_device)
#define g_num_sms (g_dev[t_dev].num_sms)

static int ensure_device() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail(ZS_ERR_CUDA, "device ordinal %d outside [0, %d)", dev, MAX_DEV);
    t_dev = dev;
    if (g_encode && g_dev[dev].num_sms) return ZS_OK;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(ZS_ERR_CUDA, "libzsae needs an sm_100-class GPU (tcgen05/TMEM); found sm_%d%d", prop.major,
                    prop.minor);
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ZS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    if (!g_dev[dev].sat) {
        CUDA_TRY(cudaMalloc(&g_dev[dev].sat, sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(g_dev[dev].sat, 0, sizeof(unsigned int)));
    }
    g_dev[dev].num_sms = prop.multiProcessorCount;
    return ZS_OK;
}
extern "C" int zs_device_check(void) { return ensure_device(); }

// dynamic shared memory above 48 KB needs an opt-in per (device, kernel)
static std::map<std::pair<int, int>, int> g_pair_clusters;      // (device, operand) -> resident CTA pairs of the PAIR GEMM kernel
static int set_smem_attr(const void* fn, int bytes) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    const auto key = std::make_pair(t_dev, fn);
    auto it = g_smem_attr.find(key);
    if (it != g_smem_attr.end() && it->second >= bytes) return ZS_OK;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_smem_attr[key] = bytes;
    return ZS_OK;
}

// fp16 range: the GEMM epilogues clamp to +-65504 (cvt.rn.satfinite) and count every thread that had to; nothing on the
// reference's value ranges gets near it, so a non-zero count means the checkpoint needs operand = bf16.
extern "C" int zs_saturation_count(void* stream, unsigned long long* count, int reset) {
    if (!count) return fail(ZS_ERR_ARG, "saturation_count: null argument");
    ZS_TRY(ensure_device());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned int v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, g_dev[t_dev].sat, sizeof(v), cudaMemcpyDeviceToHost, st));
    if (reset) CUDA_TRY(cudaMemsetAsync(g_dev[t_dev].sat, 0, sizeof(unsigned int), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *count = v;
    return ZS_OK;
}

// Experiment knobs (ZS_GEMM_DEBUG, ZS_GRU_DEBUG, ZS_PDL, ZS_GRU_*, ZS_WGRAD_DIRECT ...) exist only in builds made with
// -DZS_EXPERIMENTS (ZS_BUILD_EXPERIMENTS=1 for _lib.build); the shipped library never reads the environment.
static inline int env_int(const char* name, int dflt) {
#ifdef ZS_EXPERIMENTS
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// -------------------------------------------------------------------------------------------------
// launch accounting + optional per-kernel-class CUDA-event timing (zs_profile_begin / zs_profile_end)
// -------------------------------------------------------------------------------------------------
enum { KC_GEMM = 0, KC_GRU = 1, KC_OTHER = 2, KC_COUNT = 3 };
struct ProfSpan { cudaEvent_t a, b; int cls; double flops; const char* name; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_prof;
static long long g_launches[KC_COUNT] = {0, 0, 0};
static double g_flops[KC_COUNT] = {0, 0, 0};

struct LaunchScope {     // brackets one kernel launch on `st`
    cudaStream_t st; int cls; double flops; const char* name; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(cudaStream_t s, int c, double f = 0.0, const char* nm = "") : st(s), cls(c), flops(f), name(nm) {
        g_launches[c]++; g_flops[c] += f;
        if (g_prof_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~LaunchScope() {
        if (a) { cudaEventRecord(b, st); g_prof.push_back({a, b, cls, flops, name}); }
    }
};

extern "C" void zs_profile_begin(void) {
    for (auto& s : g_prof) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    g_prof.clear();
    for (int i = 0; i < KC_COUNT; ++i) { g_launches[i] = 0; g_flops[i] = 0; }
    g_prof_on = true;
}
extern "C" int zs_profile_end(double* ms, double* flops, long long* launches) {
    g_prof_on = false;
    for (int i = 0; i < KC_COUNT; ++i) { ms[i] = 0; flops[i] = g_flops[i]; launches[i] = g_launches[i]; }
    for (auto& s : g_prof) {
        CUDA_TRY(cudaEventSynchronize(s.b));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.cls] += t;
        cudaEventDestroy(s.a); cudaEventDestroy(s.b);
    }
    g_prof.clear();
    return ZS_OK;
}
extern "C" int zs_profile_detail(double* ms, double* flops, int* cls, int max) {
    int n = 0;
    for (auto& s : g_prof) {
        if (n < max) {
            if (cudaEventSynchronize(s.b) != cudaSuccess) return -1;
            float t = 0.f;
            if (cudaEventElapsedTime(&t, s.a, s.b) != cudaSuccess) return -1;
            ms[n] = t; flops[n] = s.flops; cls[n] = s.cls;
        }
        ++n;
    }
    return n;
}
extern "C" const char* zs_profile_name(int i) {
    return (i >= 0 && i < static_cast<int>(g_prof.size())) ? g_prof[i].name : "";
}
extern "C" void zs_launch_counts(long long* launches) {
    for (int i = 0; i < KC_COUNT; ++i) launches[i] = g_launches[i];
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static inline int buf_rows(int T, int halo) { return round_up(T + 2 * halo, 2); }

// -------------------------------------------------------------------------------------------------
// one conv / linear layer
// -------------------------------------------------------------------------------------------------
static int make_map(CUtensorMap* m, int operand, void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides, const cuuint32_t* box, bool swizzle = true) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(m, operand == ZS_OPERAND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                          rank, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ZS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)r, rank);
    return ZS_OK;
}

struct ConvExtras {        // training-path additions to a layer launch (not part of the public descriptor)
    float* stats = nullptr;
    const float* post_emb = nullptr;
    const int64_t* post_spk = nullptr;
    int post_pitch = 0, post_n = 0, no_sat = 0;
    // zero-padding mode (seg_len < 64): zero halo rows + the edge corrections of a folded speaker embedding
    int zero_halo = 0;
    const float* edge_lo = nullptr;
    const float* edge_hi = nullptr;
};
// padding mode of the forward pass being issued on this host thread (set by zs_*_forward from cfg.seg_len)
static thread_local int t_zero_pad = 0;

// 1 (default) = layers that qualify run as CTA pairs (cta_group::2), 0 = one CTA per tile everywhere (the A/B reference of the tests)
static int g_gemm_pair_mode = 1;
static int g_gemm_reuse_mode = 1;        // bit 10 turns the tap-reusing main loop off (A/B)
static int g_gemm_deep_mode = 1;         // bit 8 of zs_set_gemm_pair_mode's argument turns the deeper single-CTA ring off (A/B)
extern "C" void zs_set_gemm_pair_mode(int mode) {
    g_gemm_pair_mode = mode & 0xff;
    g_gemm_deep_mode = (mode & 0x100) ? 0 : ((mode & 0x200) ? 2 : 1);
    g_gemm_reuse_mode = (mode & 0x400) ? 0 : ((mode & 0x800) ? 2 : 1);
}

static int launch_conv(const zs_conv_desc* d, cudaStream_t stream, const ConvExtras* ex = nullptr) {
    ZS_TRY(ensure_device());
    if (!d->w || !d->in || !d->out) return fail(ZS_ERR_ARG, "conv: null operand pointer");
    if (d->m_rows % BM || d->m_rows <= 0 || d->m_valid > d->m_rows) return fail(ZS_ERR_ARG, "conv: m_rows %d must be a multiple of 128 >= m_valid %d", d->m_rows, d->m_valid);
    if (d->c_in_pad % BK || d->c_in_pad <= 0) return fail(ZS_ERR_ARG, "conv: c_in_pad %d must be a multiple of 64", d->c_in_pad);
    if (d->in_pitch % 8 || d->c_in_valid > d->in_pitch) return fail(ZS_ERR_ARG, "conv: in_pitch %d must be a multiple of 8 and >= c_in_valid %d", d->in_pitch, d->c_in_valid);
    if (d->stride != 1 && d->stride != 2) return fail(ZS_ERR_ARG, "conv: stride %d", d->stride);
    if (d->stride == 2 && (d->in_rows & 1)) return fail(ZS_ERR_ARG, "conv: stride 2 needs an even in_rows");
    if (d->T_out < 1 || d->T_out > MAX_BN) return fail(ZS_ERR_ARG, "conv: T_out %d outside [1, 256] (segments are at most 2*seg_len-1 frames)", d->T_out);
    if (d->B < 1) return fail(ZS_ERR_ARG, "conv: B %d", d->B);
    if (d->bank && (d->w_taps != 7 || d->m_rows != 7 * BM)) return fail(ZS_ERR_ARG, "conv: bank mode needs 7 x 128 rows of 7 taps");
    if (d->out_mode == OUT_PS && (d->m_rows != d->m_valid)) return fail(ZS_ERR_ARG, "conv: pixel-shuffle needs m_valid == m_rows");
    if (d->res_mode != RES_NONE && d->out_mode != OUT_CL) return fail(ZS_ERR_ARG, "conv: a residual needs the channels-last operand output mode");
    if (d->res_mode != RES_NONE && !d->res) return fail(ZS_ERR_ARG, "conv: null residual buffer");
    if (reinterpret_cast<uintptr_t>(d->w) % 16 || reinterpret_cast<uintptr_t>(d->in) % 16) return fail(ZS_ERR_ARG, "conv: operand pointers must be 16-byte aligned");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    const bool train_ex = ex && (ex->stats || ex->post_emb || ex->no_sat);
    if (d->out_f16 && d->out_mode != OUT_NCT32) return fail(ZS_ERR_ARG, "conv: out_f16 applies to the (B, C, T) output mode");
    const int Tt = round_up(d->T_out, 16);
    const int m_tiles = d->m_rows / BM;
    int nb = d->nb_hint;
    if (nb <= 0) {
        // minimise waves x tile cost: a tile costs ~ (columns + fixed overhead) per k-step
        const int max_nb = std::max(1, std::min(MAX_BN / Tt, d->B));
        long best_cost = -1;
        for (int cand = 1; cand <= max_nb; ++cand) {
            const long tiles = static_cast<long>(m_tiles) * ((d->B + cand - 1) / cand);
            const long waves = (tiles + g_num_sms - 1) / g_num_sms;
            const long cost = waves * (cand * Tt + 48);
            if (best_cost < 0 || cost <= best_cost) {
                best_cost = cost;
                nb = cand;
            }
        }
    }
    if (nb * Tt > MAX_BN) return fail(ZS_ERR_ARG, "conv: nb %d x Tt %d exceeds 256 columns", nb, Tt);
    const int n_tiles = (d->B + nb - 1) / nb;
    const long long k_total = static_cast<long long>(d->w_taps) * d->c_in_pad;
    // CTA pairs (tcgen05 cta_group::2, M = 256): two adjacent channel tiles share the columns - each CTA stages half of them.
    // Every inference layer with an even number of channel tiles and of segments per tile.  (Before the halo-row path of the epilogue
    // was slimmed, the 16- / 32-frame layers lost as pairs - d.conv3 355 -> 391 us - and were excluded; measured again afterwards,
    // tools/layer_profile.py 960 10 5 1 | 2: d.conv2 189 -> 166 us, d.conv3 283 -> 271, the 16-frame layers unchanged.)
    const bool pair = g_gemm_pair_mode != 0 && !train_ex && !(ex && (ex->zero_halo || ex->edge_lo || ex->edge_hi)) && !d->bank &&
                      m_tiles % 2 == 0 && nb % 2 == 0 && g_num_sms >= 2;
    // tap-reusing main loop (conv_gemm.cuh, REUSE): stride-1 layers with several taps whose per-segment MMAs stay >= 128 columns wide
    // (single CTA: 128-frame segments; pairs: from 64 frames) and whose B stage with the taps' extra rows fits its slot
    const int reuse_rb = Tt + (d->bank ? d->w_taps : d->taps) - 1;
    const int reuse_g = pair ? nb / 2 : nb, reuse_nm = pair ? 2 * Tt : Tt;
    // Measured per layer (tools/layer_profile.py, mode 0x202 vs 0x602): the conv bank gains 14 % (561 -> 481 us; its main loop is bound
    // by the bytes staged per FLOP), conv2 / conv3 (one channel tile) are neutral, and the CTA-PAIR layers lose - they already stage
    // half the columns per CTA and run near the tensor pipe's rate (d.conv6, one MMA per step: 527 -> 543 us; d.conv5 / d.conv4 /
    // conv5 with two M = 256, N = 128 MMAs per step: 514 -> 584 us).  Default: single-CTA layers only; mode bit 0x800 also reuses in pairs.
    const bool reuse = g_gemm_reuse_mode != 0 && (!pair || g_gemm_reuse_mode == 2) && !train_ex && !(ex && (ex->zero_halo || ex->edge_lo || ex->edge_hi)) && d->stride == 1 &&
                       (d->taps > 1 || d->bank) && reuse_nm >= 128 && reuse_rb <= 256 &&
                       reuse_g * reuse_rb * 128 <= (pair ? REUSE_PAIR_B_BYTES : REUSE_B_BYTES);
    const int nb_box = reuse ? (pair ? 1 : nb) : (pair ? nb / 2 : nb);
    const int rows_box = reuse ? reuse_rb : Tt;
    // the deeper single-CTA ring (four stages, one output tile per epilogue set) is possible for inference layers without a residual
    // whose output is channels-last; measured per layer (tools/layer_profile.py 960 10 5 1 | 0x101) it pays on the pixel-shuffle
    // up-convs on 16 / 32 frames (d.conv3 310 -> 280 us, d.conv1 172 -> 166) and costs 1-3 % on bank / conv2 / conv3 / the encoder's
    // dense layers (their main loop is bound by L2 -> shared-memory bandwidth, not latency): used for the pixel-shuffle layers only
    const bool deep = g_gemm_deep_mode != 0 && !pair && !reuse && !train_ex && !(ex && (ex->zero_halo || ex->edge_lo || ex->edge_hi)) &&
                      d->res_mode == RES_NONE && (d->out_mode == OUT_PS || (g_gemm_deep_mode == 2 && d->out_mode != OUT_NCT32));

    {   // A: weights [m_rows][k_total]
        cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(d->m_rows)};
        cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * 2};
        cuuint32_t box[2] = {BK, BM};
        ZS_TRY(make_map(&p.tmA, d->operand, const_cast<void*>(d->w), 2, dims, strides, box));
    }
    if (d->stride == 1) {   // B: (channel, row, segment)
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->c_in_valid), static_cast<cuuint64_t>(d->in_rows), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->in_pitch) * 2, static_cast<cuuint64_t>(d->in_rows) * d->in_pitch * 2};
        cuuint32_t box[3] = {BK, static_cast<cuuint32_t>(rows_box), static_cast<cuuint32_t>(nb_box)};
        ZS_TRY(make_map(&p.tmB, d->operand, const_cast<void*>(d->in), 3, dims, strides, box));
    } else {                // B: (channel, row parity, row pair, segment)
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->c_in_valid), 2, static_cast<cuuint64_t>(d->in_rows / 2), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->in_pitch) * 2, static_cast<cuuint64_t>(d->in_pitch) * 4,
                                 static_cast<cuuint64_t>(d->in_rows) * d->in_pitch * 2};
        cuuint32_t box[4] = {BK, 1, static_cast<cuuint32_t>(Tt), static_cast<cuuint32_t>(nb_box)};
        ZS_TRY(make_map(&p.tmB, d->operand, const_cast<void*>(d->in), 4, dims, strides, box));
    }
    if (d->out_mode != OUT_NCT32) {   // epilogue rounds + the TMA store map of the channels-last output
        const bool ps = d->out_mode == OUT_PS;
        // a round fills one 16 KB staging tile of its epilogue set: 64 frames x 128 channels; avg-pool residual tiles
        // have twice the rows of their output, so those layers run 32-frame rounds
        const int rf = d->res_mode == RES_AVG2 ? 32 : 64;
        p.rnd_frames = rf;
        p.rnd_rows = std::min(Tt, rf);
        p.rnd_sub = (Tt + rf - 1) / rf;
        // segments per round must DIVIDE the segments per tile: a round's TMA store box always covers rnd_ns
        // segments, so a partial last round would spill stale staging rows into the next tile's segments
        p.rnd_ns = 1;
        if (Tt <= rf)
            for (int c = std::min(rf / Tt, nb); c >= 1; --c)
                if (nb % c == 0) { p.rnd_ns = c; break; }
        if (d->out_halo < 0 || d->out_halo > 3) return fail(ZS_ERR_ARG, "conv: out_halo %d outside [0, 3] (kernel sizes up to 7)", d->out_halo);
        if (d->out_choff % 8 || d->out_pitch % 8) return fail(ZS_ERR_ARG, "conv: out_choff %d / out_pitch %d must be multiples of 8", d->out_choff, d->out_pitch);
        if (reinterpret_cast<uintptr_t>(d->out) % 16) return fail(ZS_ERR_ARG, "conv: output pointer must be 16-byte aligned");
        const int T_rows = d->out_halo + (ps ? 2 * d->T_out : d->T_out);
        if (T_rows > d->out_rows) return fail(ZS_ERR_ARG, "conv: output buffer has %d rows, needs %d", d->out_rows, T_rows);
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(ps ? d->m_valid / 2 : d->m_valid), static_cast<cuuint64_t>(T_rows), static_cast<cuuint64_t>(d->B)};
        cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->out_pitch) * 2, static_cast<cuuint64_t>(d->out_rows) * d->out_pitch * 2};
        cuuint32_t box[3] = {static_cast<cuuint32_t>(ps ? 64 : 128), static_cast<cuuint32_t>(ps ? 2 * p.rnd_rows : p.rnd_rows), static_cast<cuuint32_t>(p.rnd_ns)};
        void* base = static_cast<uint8_t*>(d->out) + static_cast<size_t>(d->out_choff) * 2;
        ZS_TRY(make_map(&p.tmOut, d->operand, base, 3, dims, strides, box, false));
        if (d->res_mode != RES_NONE) {   // residual tile: same channels, rows scaled by the mode
            if (d->res_pitch % 8 || reinterpret_cast<uintptr_t>(d->res) % 16) return fail(ZS_ERR_ARG, "conv: residual buffer must be 16-byte aligned with a pitch multiple of 8");
            const int rr = d->res_mode == RES_UP2 ? p.rnd_rows / 2 : (d->res_mode == RES_AVG2 ? 2 * p.rnd_rows : p.rnd_rows);
            if (rr < 1 || rr > 256) return fail(ZS_ERR_ARG, "conv: residual box rows %d", rr);
            cuuint64_t rdims[3] = {static_cast<cuuint64_t>(std::min(d->m_valid, d->res_pitch)), static_cast<cuuint64_t>(d->res_rows), static_cast<cuuint64_t>(d->B)};
            cuuint64_t rstrides[2] = {static_cast<cuuint64_t>(d->res_pitch) * 2, static_cast<cuuint64_t>(d->res_rows) * d->res_pitch * 2};
            cuuint32_t rbox[3] = {128u, static_cast<cuuint32_t>(rr), static_cast<cuuint32_t>(p.rnd_ns)};
            ZS_TRY(make_map(&p.tmRes, d->operand, const_cast<void*>(d->res), 3, rdims, rstrides, rbox, false));
        }
    }
    else if (!d->accumulate && !(ex && ex->zero_halo)) {
        // (B, C, T) output staged through shared memory: each epilogue warp stores {one 128-byte row of frames} x 32 channels boxes
        // (128-byte swizzle -> conflict-free 16-byte shared stores).  Needs whole rows; anything else takes the direct-store path.
        // rows of 128 bytes with the 128-byte swizzle, or (16-frame segments in fp32) plain 64-byte rows
        const int es = d->out_f16 ? 2 : 4, fpr = std::min(128 / es, d->T_out), rb = fpr * es;
        if ((rb == 128 || rb == 64) && d->T_out % fpr == 0 && reinterpret_cast<uintptr_t>(d->out) % 16 == 0) {
            cuuint64_t dims[3] = {static_cast<cuuint64_t>(d->T_out), static_cast<cuuint64_t>(d->m_valid), static_cast<cuuint64_t>(d->B)};
            cuuint64_t strides[2] = {static_cast<cuuint64_t>(d->T_out) * es, static_cast<cuuint64_t>(d->m_valid) * d->T_out * es};
            cuuint32_t box[3] = {static_cast<cuuint32_t>(fpr), 32u, 1u};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = g_encode(&p.tmOut, d->out_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d->out, dims, strides,
                                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(ZS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the (B, C, T) output", (int)r);
            p.nct_tma = rb;
        }
    }
    p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.nb = nb; p.Tt = Tt; p.T = d->T_out; p.B = d->B; p.N = nb * Tt;
    p.kc = d->c_in_pad / BK; p.taps = d->taps; p.bank = d->bank; p.stride = d->stride; p.in_row0 = d->in_row0;
    p.c_in_pad = d->c_in_pad; p.m_valid = d->m_valid;
    {   // MMAs of a tap's last chunk that still cover valid channels (c_in_valid need not be a multiple of 64)
        const int tail = d->c_in_valid - (p.kc - 1) * BK;
        p.last_mmas = tail >= BK ? BK / 16 : std::max(1, (tail + 15) / 16);
    }
    p.bias = d->bias; p.spk = reinterpret_cast<const long long*>(d->spk); p.bias_stride = d->m_rows; p.n_spk = d->n_spk > 0 ? d->n_spk : 1;
    p.lrelu = d->lrelu; p.ns = d->ns; p.inorm = d->inorm;
    p.res_mode = d->res_mode; p.res = d->res; p.res_rows = d->res_rows; p.res_pitch = d->res_pitch; p.res_halo = d->res_halo;
    p.act = d->act; p.out_mode = d->out_mode; p.out = d->out; p.out_rows = d->out_rows; p.out_pitch = d->out_pitch;
    p.out_halo = d->out_halo; p.out_choff = d->out_choff; p.accumulate = d->accumulate; p.out_f16 = d->out_f16;
    p.idesc = umma_idesc_f16(d->operand == ZS_OPERAND_BF16 ? 1 : 0, reuse ? reuse_nm : p.N, pair ? 256 : 128);
    p.pair = pair ? 1 : 0;
    p.reuse_rb = reuse_rb; p.reuse_g = reuse_g; p.reuse_nm = reuse_nm;
    p.debug = env_int("ZS_GEMM_DEBUG", 0);
    p.sat_count = g_dev[t_dev].sat;
    if (ex) {
        p.stats = ex->stats; p.post_emb = ex->post_emb; p.post_spk = reinterpret_cast<const long long*>(ex->post_spk);
        p.post_pitch = ex->post_pitch; p.post_n = ex->post_n > 0 ? ex->post_n : 1; p.no_sat = ex->no_sat;
        p.zero_halo = ex->zero_halo; p.edge_lo = ex->edge_lo; p.edge_hi = ex->edge_hi;
        if (p.post_emb && !p.post_spk) return fail(ZS_ERR_ARG, "conv: post-add embedding needs speaker ids");
    }

    int max_pairs = g_num_sms / 2;
    if (pair) {   // CTA pairs the device can keep resident at once (a TPC with one SM fused off holds none): asked once per device and kernel
        std::lock_guard<std::mutex> lk(g_dev_mu);
        const auto key = std::make_pair(t_dev, static_cast<int>(d->operand == ZS_OPERAND_BF16));
        auto it = g_pair_clusters.find(key);
        if (it == g_pair_clusters.end()) {
            using KT = void (*)(const GemmParams);
            KT k2 = d->operand == ZS_OPERAND_BF16 ? conv_gemm_kernel<__nv_bfloat16, false, false, true> : conv_gemm_kernel<__half, false, false, true>;
            CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
            cudaLaunchConfig_t qc;
            memset(&qc, 0, sizeof(qc));
            qc.gridDim = dim3(2 * max_pairs); qc.blockDim = dim3(GEMM_THREADS); qc.dynamicSmemBytes = PAIR_SMEM_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k2, &qc) != cudaSuccess || n < 1) { cudaGetLastError(); n = max_pairs; }
            it = g_pair_clusters.emplace(key, std::min(n, max_pairs)).first;
        }
        max_pairs = it->second;
    }
    const int grid = pair ? 2 * std::min((m_tiles / 2) * n_tiles, max_pairs) : std::min(m_tiles * n_tiles, g_num_sms);
    const int which = d->operand == ZS_OPERAND_BF16 ? 1 : 0;
    const int zp = p.zero_halo ? 1 : 0;
    using KernelT = void (*)(const GemmParams);
    if (zp && train_ex) return fail(ZS_ERR_ARG, "conv: the zero-padding mode is inference only");
    KernelT kern = zp ? (which ? conv_gemm_kernel<__nv_bfloat16, true, false, false> : conv_gemm_kernel<__half, true, false, false>)
                 : train_ex ? (which ? conv_gemm_kernel<__nv_bfloat16, false, true, false> : conv_gemm_kernel<__half, false, true, false>)
                 : (pair && reuse) ? (which ? conv_gemm_kernel<__nv_bfloat16, false, false, true, false, true> : conv_gemm_kernel<__half, false, false, true, false, true>)
                 : reuse ? (which ? conv_gemm_kernel<__nv_bfloat16, false, false, false, false, true> : conv_gemm_kernel<__half, false, false, false, false, true>)
                 : pair ? (which ? conv_gemm_kernel<__nv_bfloat16, false, false, true> : conv_gemm_kernel<__half, false, false, true>)
                 : deep ? (which ? conv_gemm_kernel<__nv_bfloat16, false, false, false, true> : conv_gemm_kernel<__half, false, false, false, true>)
                        : (which ? conv_gemm_kernel<__nv_bfloat16, false, false, false> : conv_gemm_kernel<__half, false, false, false>);
    const int smem_bytes = reuse ? (pair ? REUSE_PAIR_SMEM_BYTES : REUSE_SMEM_BYTES) : (pair ? PAIR_SMEM_BYTES : (deep ? DEEP_SMEM_BYTES : GEMM_SMEM_BYTES));
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(kern), smem_bytes));
    {   // algorithmic FLOPs: 2 * valid out channels * true taps * true in channels * valid frames
        double taps_sum = d->bank ? 28.0 / 7.0 : static_cast<double>(d->taps);
        const double flops = 2.0 * d->m_valid * taps_sum * d->c_in_valid * static_cast<double>(d->B) * d->T_out;
        LaunchScope scope(stream, KC_GEMM, flops, d->stride == 2 ? "conv_gemm s2" : (d->taps > 1 ? "conv_gemm" : "conv_gemm k1"));
        // programmatic dependent launch (ZS_PDL=1): measured neutral on this path (9.505 vs 9.514 ms per step) - the
        // kernels run back to back without host gaps and every CTA needs a whole SM, so only prologues could overlap
        static const int use_pdl = env_int("ZS_PDL", 0);
        p.pdl = use_pdl;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (use_pdl) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (pair) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr; cfg.numAttrs = na;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    }
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_conv1d_cl(const zs_conv_desc* d, void* stream) { return launch_conv(d, static_cast<cudaStream_t>(stream)); }

// -------------------------------------------------------------------------------------------------
// small launchers
// -------------------------------------------------------------------------------------------------
static int launch_pack_nct(const float* x, int B, int C, int T, void* out, int rows, int pitch, int halo, int choff,
                           int lrelu, float ns, int operand, int zero_pad, cudaStream_t st) {
    if (halo >= T && halo > 0) return fail(ZS_ERR_ARG, "pack: halo %d needs more than %d frames", halo, T);
    const int c_fill = zero_pad ? pitch - choff : C;
    dim3 grid((T + 31) / 32, (c_fill + 31) / 32, B), block(32, 8);
    LaunchScope scope(st, KC_OTHER, 0.0, "pack_nct_kernel");
    if (operand == ZS_OPERAND_BF16)
        pack_nct_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(x, static_cast<__nv_bfloat16*>(out), C, T, rows, pitch, halo, choff, c_fill, lrelu, ns, t_zero_pad);
    else
        pack_nct_kernel<__half><<<grid, block, 0, st>>>(x, static_cast<__half*>(out), C, T, rows, pitch, halo, choff, c_fill, lrelu, ns, t_zero_pad);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
// x -> (bank input with halo, leaky-relu'd copy inside the concat buffer) in one pass
// x_dtype: ZS_X_F32 | ZS_X_F16; x_layout: ZS_X_NCT (B, C, T) | ZS_X_NTC (B, T, C)
template <typename OT>
static void pack_x_dual_dispatch(const void* x, int x_dtype, int x_layout, dim3 grid, cudaStream_t st, int C, int T, OT* bank_p, int bank_rows,
                                 int bank_pitch, int bank_halo, OT* cat_p, int cat_rows, int cat_pitch, int cat_choff, float ns) {
#define ZS_PXD(IT, NTC) pack_x_dual_kernel<OT, IT, NTC><<<grid, 256, 0, st>>>(static_cast<const IT*>(x), C, T, bank_p, bank_rows, bank_pitch, \
                                                                              bank_halo, cat_p, cat_rows, cat_pitch, cat_choff, C, ns, t_zero_pad)
    if (x_dtype == ZS_X_F16) { if (x_layout == ZS_X_NTC) ZS_PXD(__half, true); else ZS_PXD(__half, false); }
    else { if (x_layout == ZS_X_NTC) ZS_PXD(float, true); else ZS_PXD(float, false); }
#undef ZS_PXD
}
static int launch_pack_x_dual(const void* x, int x_dtype, int x_layout, int B, int C, int T, void* bank_p, int bank_rows, int bank_pitch,
                              int bank_halo, void* cat_p, int cat_rows, int cat_pitch, int cat_choff, float ns, int operand, cudaStream_t st) {
    if (bank_halo >= T) return fail(ZS_ERR_ARG, "pack: halo %d needs more than %d frames", bank_halo, T);
    if ((cat_choff & 1) || (bank_pitch & 1)) return fail(ZS_ERR_ARG, "pack: channel offsets / pitches must be even");
    if (x_dtype != ZS_X_F32 && x_dtype != ZS_X_F16) return fail(ZS_ERR_ARG, "pack: x_dtype %d", x_dtype);
    if (x_layout != ZS_X_NCT && x_layout != ZS_X_NTC) return fail(ZS_ERR_ARG, "pack: x_layout %d", x_layout);
    if (x_dtype == ZS_X_F16 && operand != ZS_OPERAND_FP16)
        return fail(ZS_ERR_ARG, "an fp16 input is rounded once more by bf16 operands: upload fp32 with operand = bf16");
    dim3 grid((T + 31) / 32, (C + 63) / 64, B);
    LaunchScope scope(st, KC_OTHER, 0.0, "pack_x_dual_kernel");
    // (B, C, T) input with 16-byte aligned rows on both sides: whole 512-byte channel rows in, 16-byte operand stores out
    const bool wide = x_layout == ZS_X_NCT && T % 4 == 0 && bank_pitch % 8 == 0 && cat_pitch % 8 == 0 && cat_choff % 8 == 0 &&
                      reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(bank_p) % 16 == 0 && reinterpret_cast<uintptr_t>(cat_p) % 16 == 0;
    if (wide) {
        const dim3 wgrid((T + 127) / 128, (C + 63) / 64, B);
#define ZS_PXW(OT_, IT_) pack_x_dual_wide_kernel<OT_, IT_><<<wgrid, 256, 0, st>>>(static_cast<const IT_*>(x), C, T, static_cast<OT_*>(bank_p), bank_rows, \
                                                                                 bank_pitch, bank_halo, static_cast<OT_*>(cat_p), cat_rows, cat_pitch, cat_choff, C, ns, t_zero_pad)
        if (operand == ZS_OPERAND_BF16) ZS_PXW(__nv_bfloat16, float);
        else if (x_dtype == ZS_X_F16) ZS_PXW(__half, __half);
        else ZS_PXW(__half, float);
#undef ZS_PXW
        CUDA_TRY(cudaGetLastError());
        return ZS_OK;
    }
    if (operand == ZS_OPERAND_BF16)
        pack_x_dual_dispatch<__nv_bfloat16>(x, x_dtype, x_layout, grid, st, C, T, static_cast<__nv_bfloat16*>(bank_p), bank_rows, bank_pitch, bank_halo,
                                            static_cast<__nv_bfloat16*>(cat_p), cat_rows, cat_pitch, cat_choff, ns);
    else
        pack_x_dual_dispatch<__half>(x, x_dtype, x_layout, grid, st, C, T, static_cast<__half*>(bank_p), bank_rows, bank_pitch, bank_halo,
                                     static_cast<__half*>(cat_p), cat_rows, cat_pitch, cat_choff, ns);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_pack_nct(const float* x, int B, int C, int T, void* out, int rows, int pitch, int halo, int choff,
                           int lrelu, float ns, int operand, int zero_pad_channels, void* stream) {
    t_zero_pad = 0;
    return launch_pack_nct(x, B, C, T, out, rows, pitch, halo, choff, lrelu, ns, operand, zero_pad_channels, static_cast<cudaStream_t>(stream));
}

static int launch_onehot(const float* logits, const float* noise, const uint64_t* seg_seeds, int B, int C, int T8, float* act, int32_t* ids, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(C) * (T8 + 1) * 4 + static_cast<size_t>(T8) * 4;
    if (smem > 200 * 1024) return fail(ZS_ERR_ARG, "bottleneck: C %d x T8 %d does not fit shared memory", C, T8);
    ZS_TRY(ensure_device());
    if (smem > 48 * 1024) ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(bottleneck_onehot_kernel), 200 * 1024));
    LaunchScope scope(st, KC_OTHER, 0.0, "bottleneck_onehot_kernel");
    if (!noise && !seg_seeds) return fail(ZS_ERR_ARG, "bottleneck: neither a noise tensor nor per-segment seeds");
    bottleneck_onehot_kernel<<<B, 512, smem, st>>>(logits, noise, reinterpret_cast<const unsigned long long*>(seg_seeds), C, T8, act, ids);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
extern "C" int zs_bottleneck_one_hot(const float* logits, const float* noise, int B, int C, int T8, float* act,
                                     int32_t* unit_ids, void* stream) {
    if (!logits || !noise) return fail(ZS_ERR_ARG, "bottleneck: null logits/noise");
    return launch_onehot(logits, noise, nullptr, B, C, T8, act, unit_ids, static_cast<cudaStream_t>(stream));
}

static int launch_gru(const void* gx, const float* whhT, const float* bhh, int B, int T, int H, void* out, int rows,
                      int pitch, int halo, int choff, int operand, cudaStream_t st, void* gates = nullptr) {
    if (H < 1 || H > 1024) return fail(ZS_ERR_ARG, "gru: hidden size %d outside [1, 1024]", H);
    constexpr int NBG = 4;
    dim3 grid((B + NBG - 1) / NBG, 2);
    const size_t smem = static_cast<size_t>(NBG) * H * 4;
    LaunchScope scope(st, KC_GRU, 2.0 * 2 * B * static_cast<double>(T) * 3 * H * H, "gru_simple_kernel");
    if (operand == ZS_OPERAND_BF16)
        gru_simple_kernel<__nv_bfloat16, NBG><<<grid, H, smem, st>>>(static_cast<const __nv_bfloat16*>(gx), whhT, bhh, B, T, H, static_cast<__nv_bfloat16*>(out), rows, pitch, halo, choff, static_cast<__nv_bfloat16*>(gates));
    else
        gru_simple_kernel<__half, NBG><<<grid, H, smem, st>>>(static_cast<const __half*>(gx), whhT, bhh, B, T, H, static_cast<__half*>(out), rows, pitch, halo, choff, static_cast<__half*>(gates));
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

static bool gru_cluster_ok(int H) { return H % GRU_UNITS == 0 && H / GRU_UNITS >= 1 && H / GRU_UNITS <= 8; }

// GRU weight images of both directions: the swizzled shared-memory images of the cluster kernels, followed by the plain
// r|z rows the wide kernel loads into TMEM
static size_t gru_images_bytes(int H) {
    const size_t NC = H / GRU_UNITS;
    return 2 * NC * (static_cast<size_t>(gru_w_image_bytes(H)) + static_cast<size_t>(128) * H * 2);
}
static const void* gru_wrz_part(const void* img, int H) {
    return static_cast<const uint8_t*>(img) + static_cast<size_t>(2) * (H / GRU_UNITS) * gru_w_image_bytes(H);
}
static int launch_gru_cluster(const void* w_img, const float* bhh, const void* gx, int B, int T, int H, void* out, int rows,
                              int pitch, int halo, int choff, int operand, cudaStream_t st, void* gates = nullptr, void* xchg = nullptr, size_t xchg_bytes = 0) {
    ZS_TRY(ensure_device());
    GruParams p;
    p.gates = gates;
    // gates through tanh.approx.f32 (one MUFU each; 2^-11 relative error, below the fp16 rounding of the state that
    // feeds the next step's MMA) on the inference path; the training forward keeps the exp/rcp form
    p.fast_act = env_int("ZS_GRU_FAST_ACT", gates == nullptr ? 1 : 0);
    p.w_img = w_img; p.bhh = bhh; p.gx = gx; p.out = out; p.B = B; p.T = T; p.H = H;
    p.out_rows = rows; p.out_pitch = pitch; p.out_halo = halo; p.out_choff = choff;
    p.fmt = operand == ZS_OPERAND_BF16 ? 1 : 0;
    p.debug = 0;
    p.dbg = nullptr;
#ifdef ZS_EXPERIMENTS
    {   // timing experiments only (results are wrong with any bit set): 1 = no state exchange, 2 = no gx loads, 4 = no stores
        p.debug = env_int("ZS_GRU_DEBUG", 0);
        if (p.debug & 8) {
            static long long* dbuf = nullptr;
            if (!dbuf) cudaMalloc(&dbuf, 8 * 512 * sizeof(long long));
            p.dbg = dbuf;
            if (T <= 512) {   // dump the previous call's stamps (host-synchronous)
                static bool first = true;
                if (!first) {
                    std::vector<long long> hbuf(8 * T);
                    cudaMemcpy(hbuf.data(), dbuf, hbuf.size() * 8, cudaMemcpyDeviceToHost);
                    const int t = T / 2;
                    const long long* r = &hbuf[t * 8];
                    const long long* rn = &hbuf[(t + 1) * 8];
                    fprintf(stderr, "gru step %d: mma_issue %lld | gate: wait_mma %lld ld %lld math %lld fence+bar %lld copy_issue %lld | ctrl next-step start +%lld (step period %lld)\n",
                            t, r[1] - r[0], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], rn[0] - r[7], rn[0] - r[0]);
                }
                first = false;
            }
        }
    }
#endif
    if (static_cast<long long>(B) * T * 6 * H >= (1ll << 31) || static_cast<long long>(B) * rows * pitch >= (1ll << 31))
        return fail(ZS_ERR_ARG, "gru: %d sequences x %d steps exceed the 32-bit element offsets of the recurrence kernel", B, T);
    {   // 64 sequences per cluster with the r|z rows of W_hh in TMEM (gru_wide.cuh): the shape for batches that need more than
        // one wave of 32-sequence clusters anyway.  ZS_GRU_WIDE=0 disables, =1 forces (where its preconditions hold).
        static const int wide_mode = env_int("ZS_GRU_WIDE", 2);
        const int NCw = H / GRU_UNITS;
        // 64 sequences per cluster (2 gate passes), or 128 (4 passes) when the 64-sequence clusters would need a second wave and
        // the state of 128 sequences fits (H <= 512: 64 KB of n rows + 128 KB of state; H/2 + 256 TMEM columns)
        int npass = 2;
        if (gru_wide_smem_bytes(H, 4) <= 232448 && (H >> 1) + 256 <= 512) {
            // waves of resident clusters x cycles per step (MMAs + exchange ~5 000, ~1 900 per gate pass - measured)
            const int per_wave = std::max(1, 120 / NCw);
            auto cost = [&](int np) { const int cls = 2 * ((B + 32 * np - 1) / (32 * np)); return ((cls + per_wave - 1) / per_wave) * (5000 + 1900 * np); };
            if (cost(4) < cost(2)) npass = 4;
        }
        { const int e = env_int("ZS_GRU_NPASS", 0); if (e == 2 || e == 4) npass = e; }
        const int nseq_w = 32 * npass, groups64 = (B + nseq_w - 1) / nseq_w;
        const size_t need = static_cast<size_t>(2) * groups64 * NCw * nseq_w * 128;
        const bool can = !gates && xchg && need <= xchg_bytes && NCw >= 2 && NCw <= 8 && H % 64 == 0;
        const bool want = wide_mode == 1 || (wide_mode == 2 && 2 * ((B + GRU_FWD_NSEQ - 1) / GRU_FWD_NSEQ) * NCw > 120);
        if (can && want && wide_mode != 0) {
            GruWideParams wp;
            memset(&wp, 0, sizeof(wp));
            wp.wrz = gru_wrz_part(w_img, H); wp.w_img = w_img; wp.bhh = bhh; wp.gx = gx; wp.out = out; wp.xchg = static_cast<uint8_t*>(xchg);
            wp.B = B; wp.T = T; wp.H = H; wp.out_rows = rows; wp.out_pitch = pitch; wp.out_halo = halo; wp.out_choff = choff;
            wp.fmt = operand == ZS_OPERAND_BF16 ? 1 : 0;
            const int smem_w = gru_wide_smem_bytes(H, npass);
            using WideT = void (*)(const GruWideParams);
            // 16 gate warps: a step's passes run side by side (the gate math is latency-bound at two warps per scheduler)
            const int gw = env_int("ZS_GRU_GW", 16) == 8 ? 8 : 16;
            WideT wk = gw == 8 ? (npass == 4 ? (wp.fmt ? gru_wide_kernel<__nv_bfloat16, 4, 8> : gru_wide_kernel<__half, 4, 8>)
                                             : (wp.fmt ? gru_wide_kernel<__nv_bfloat16, 2, 8> : gru_wide_kernel<__half, 2, 8>))
                               : (npass == 4 ? (wp.fmt ? gru_wide_kernel<__nv_bfloat16, 4, 16> : gru_wide_kernel<__half, 4, 16>)
                                             : (wp.fmt ? gru_wide_kernel<__nv_bfloat16, 2, 16> : gru_wide_kernel<__half, 2, 16>));
            ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(wk), smem_w));
            cudaLaunchConfig_t wc;
            memset(&wc, 0, sizeof(wc));
            wc.gridDim = dim3(2 * groups64 * NCw); wc.blockDim = dim3(32 * (gw + 1)); wc.dynamicSmemBytes = smem_w; wc.stream = st;
            cudaLaunchAttribute wa[1];
            wa[0].id = cudaLaunchAttributeClusterDimension;
            wa[0].val.clusterDim.x = NCw; wa[0].val.clusterDim.y = 1; wa[0].val.clusterDim.z = 1;
            wc.attrs = wa; wc.numAttrs = 1;
            LaunchScope scope(st, KC_GRU, 2.0 * 2 * B * static_cast<double>(T) * 3 * H * H, "gru_wide_kernel");
            CUDA_TRY(cudaLaunchKernelEx(&wc, wk, wp));
            return ZS_OK;
        }
    }
    // 32 sequences per cluster is the throughput shape; when the batch fits one wave of 16-sequence clusters (15 eight-CTA
    // clusters are resident on a B200) the smaller shape halves the per-step exchange and gate math: lower latency
    const int NC = H / GRU_UNITS;
    int nseq = (2 * ((B + GRU_FWD_NSEQ_SMALL - 1) / GRU_FWD_NSEQ_SMALL) * NC <= 120) ? GRU_FWD_NSEQ_SMALL : GRU_FWD_NSEQ;
    { const int e = env_int("ZS_GRU_NSEQ", 0); if (e == 16 || e == 32) nseq = e; }
    const int n_groups = (B + nseq - 1) / nseq;
    const int smem = gru_smem_bytes(H, nseq);
    using KernelT = void (*)(const GruParams);
    const int which = p.fmt;
    KernelT kern = nseq == 16 ? (which ? gru_cluster_kernel<__nv_bfloat16, 16> : gru_cluster_kernel<__half, 16>)
                              : (which ? gru_cluster_kernel<__nv_bfloat16, 32> : gru_cluster_kernel<__half, 32>);
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(kern), smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * n_groups * NC);
    cfg.blockDim = dim3(GRU_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (ZS_DBG(p) & 16) {
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
        fprintf(stderr, "gru: max active clusters of %d CTAs with %d B smem: %d (%s); launching %d clusters of %d sequences\n", NC, smem, ncl, cudaGetErrorString(e), 2 * n_groups, nseq);
    }
    {   // state exchange through L2 + multicast (ZS_GRU_XCHG=0: direct SM-to-SM bulk copies)
        static const int xmode = env_int("ZS_GRU_XCHG", 1);
        p.xchg = nullptr;
        const size_t need = static_cast<size_t>(2) * n_groups * NC * nseq * 128;     // one slice per CTA of every cluster
        if (xmode && NC > 1 && xchg && need <= xchg_bytes) p.xchg = static_cast<uint8_t*>(xchg);
    }
    LaunchScope scope(st, KC_GRU, 2.0 * 2 * B * static_cast<double>(T) * 3 * H * H, "gru_cluster_kernel");
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    return ZS_OK;
}

static int pack_gru_image(void* img, const float* const* w_hh_dirs, const float* w_hh_packed, int H, int operand, cudaStream_t st) {
    const int NC = H / GRU_UNITS;
    for (int dir = 0; dir < 2; ++dir) {
        const float* W = w_hh_dirs ? w_hh_dirs[dir] : w_hh_packed + static_cast<size_t>(dir) * 3 * H * H;
        void* dst = static_cast<uint8_t*>(img) + static_cast<size_t>(dir) * NC * gru_w_image_bytes(H);
        const int blocks = (3 * H * H + 255) / 256;
        if (operand == ZS_OPERAND_BF16) gru_pack_whh_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(dst), H);
        else gru_pack_whh_kernel<__half><<<blocks, 256, 0, st>>>(W, static_cast<__half*>(dst), H);
        void* plain = const_cast<uint8_t*>(static_cast<const uint8_t*>(gru_wrz_part(img, H))) + static_cast<size_t>(dir) * NC * 128 * H * 2;
        const int blocks2 = (2 * H * H + 255) / 256;
        if (operand == ZS_OPERAND_BF16) gru_pack_wrz_plain_kernel<__nv_bfloat16><<<blocks2, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(plain), H);
        else gru_pack_wrz_plain_kernel<__half><<<blocks2, 256, 0, st>>>(W, static_cast<__half*>(plain), H);
    }
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

extern "C" int zs_gru_recurrence(const float* gx, const float* w_hh, const float* b_hh, int B, int T, int H, void* out,
                                 int out_rows, int out_pitch, int out_halo, int out_choff, int operand, int impl, void* stream) {
    // test entry: packs w_hh and the fp32 projections into temporary (stream-ordered) buffers, then runs the recurrence
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* gx_ot = nullptr;
    const size_t n_gx = static_cast<size_t>(B) * T * 6 * H;
    CUDA_TRY(cudaMallocAsync(&gx_ot, n_gx * 2, st));
    if (operand == ZS_OPERAND_BF16) cast_to_ot_kernel<__nv_bfloat16><<<static_cast<unsigned>((n_gx + 255) / 256), 256, 0, st>>>(gx, static_cast<__nv_bfloat16*>(gx_ot), n_gx);
    else cast_to_ot_kernel<__half><<<static_cast<unsigned>((n_gx + 255) / 256), 256, 0, st>>>(gx, static_cast<__half*>(gx_ot), n_gx);
    struct FreeLater { void* p; cudaStream_t s; ~FreeLater() { cudaFreeAsync(p, s); } } free_gx{gx_ot, st};
    if (impl == 2 && !gru_cluster_ok(H)) return fail(ZS_ERR_ARG, "gru: the cluster kernel needs H %% 64 == 0 and H <= 512 (H = %d)", H);
    if (impl == 2 || (impl == 0 && gru_cluster_ok(H))) {
        void* img = nullptr;
        const size_t bytes = gru_images_bytes(H);
        CUDA_TRY(cudaMallocAsync(&img, bytes, st));
        int r = pack_gru_image(img, nullptr, w_hh, H, operand, st);
        void* xch = nullptr;
        const size_t xch_bytes = static_cast<size_t>(2) * round_up(B, 128) * H * 2;      // state-exchange scratch (L2 multicast path)
        CUDA_TRY(cudaMallocAsync(&xch, xch_bytes, st));
        if (r == ZS_OK) r = launch_gru_cluster(img, b_hh, gx_ot, B, T, H, out, out_rows, out_pitch, out_halo, out_choff, operand, st, nullptr, xch, xch_bytes);
        cudaFreeAsync(xch, st);
        cudaFreeAsync(img, st);
        return r;
    }
    float* wt = nullptr;
    CUDA_TRY(cudaMallocAsync(&wt, static_cast<size_t>(2) * 3 * H * H * 4, st));
    for (int dir = 0; dir < 2; ++dir)
        transpose_whh_kernel<<<(3 * H * H + 255) / 256, 256, 0, st>>>(w_hh + static_cast<size_t>(dir) * 3 * H * H, wt + static_cast<size_t>(dir) * 3 * H * H, H);
    int r = launch_gru(gx_ot, wt, b_hh, B, T, H, out, out_rows, out_pitch, out_halo, out_choff, operand, st);
    cudaFreeAsync(wt, st);
    return r;
}

// -------------------------------------------------------------------------------------------------
// packed weights
// -------------------------------------------------------------------------------------------------
struct Layer {          // one GEMM's worth of packed weights
    void* w = nullptr;      // operand type [m_rows][w_taps * c_in_pad]
    float* bias = nullptr;  // [n_tab][m_rows]
    float* edge_lo = nullptr, *edge_hi = nullptr;   // zero-padding mode, folded embedding, k > 1: [n_tab][m_rows] = W_0 . e, W_{k-1} . e
    int m_rows = 0, m_valid = 0, taps = 1, w_taps = 1, c_in_pad = 0, c_in_valid = 0, per_spk = 0, ps = 0, n_tab = 1;
    // training: transposed weights for the data-gradient GEMM, [t_rows][t_taps * t_kpad] fp16 (pack_weight_T_kernel)
    void* wt = nullptr;
    int t_rows = 0, t_valid = 0, t_taps = 1, t_kpad = 0, t_kvalid = 0;
};

struct DevPool {        // owns every device allocation of a handle
    std::vector<void*> ptrs;
    int alloc(void** p, size_t bytes, cudaStream_t st) {
        if (*p) return ZS_OK;        // re-pack into the existing allocation
        CUDA_TRY(cudaMalloc(p, bytes));
        CUDA_TRY(cudaMemsetAsync(*p, 0, bytes, st));
        ptrs.push_back(*p);
        return ZS_OK;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

// conv / linear weight W (C_out, C_in, k) -> Layer; only input channels [ci_lo, ci_lo + ci_n) enter the GEMM.
// emb != null folds  sum_{j, ci} W[co][ci_emb_lo + ci][j] * emb[s][ci]  into a per-speaker bias table.
static int pack_layer(DevPool& pool, Layer& L, int operand, const float* W, const float* b, int C_out, int C_in, int k,
                      int ci_lo, int ci_n, int ps, const float* emb, int emb_ci_lo, int C_e, int n_spk,
                      cudaStream_t st, int zero_pad = 0) {
    L.m_rows = round_up(C_out, BM); L.m_valid = C_out; L.taps = k; L.w_taps = k;
    L.c_in_pad = round_up(ci_n, BK); L.c_in_valid = ci_n; L.ps = ps; L.per_spk = emb ? 1 : 0;
    const long long k_total = static_cast<long long>(k) * L.c_in_pad;
    ZS_TRY(pool.alloc(&L.w, static_cast<size_t>(L.m_rows) * k_total * 2, st));
    if (operand == ZS_OPERAND_BF16)
        CUDA_TRY(launch_pack_weight(W, static_cast<__nv_bfloat16*>(L.w), C_out, C_in, k, ci_lo, ci_n, k_total, L.c_in_pad, 0, 0, ps, st));
    else
        CUDA_TRY(launch_pack_weight(W, static_cast<__half*>(L.w), C_out, C_in, k, ci_lo, ci_n, k_total, L.c_in_pad, 0, 0, ps, st));
    const int n_tab = emb ? n_spk : 1;
    L.n_tab = n_tab;
    ZS_TRY(pool.alloc(reinterpret_cast<void**>(&L.bias), static_cast<size_t>(n_tab) * L.m_rows * 4, st));
    const long long warps = static_cast<long long>(n_tab) * C_out;
    fold_bias_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, st>>>(W, b, emb, L.bias, C_out, C_in, k, emb_ci_lo, emb ? C_e : 0, n_tab, L.m_rows, 0, ps);
    CUDA_TRY(cudaGetLastError());
    if (zero_pad && emb && k > 1) {
        // zero padding is applied AFTER the embedding was added (model/model.py:319-325: pad_layer(x + emb)), so the
        // taps that fall into the padding at the two ends of a segment must not see it: frame 0 loses W_0 . e, frame
        // T-1 loses W_{k-1} . e (k = 3 on this path)
        if (k != 3) return fail(ZS_ERR_ARG, "zero-padding mode: a folded speaker embedding is implemented for kernel size 3 (got %d)", k);
        ZS_TRY(pool.alloc(reinterpret_cast<void**>(&L.edge_lo), static_cast<size_t>(n_tab) * L.m_rows * 4, st));
        ZS_TRY(pool.alloc(reinterpret_cast<void**>(&L.edge_hi), static_cast<size_t>(n_tab) * L.m_rows * 4, st));
        fold_bias_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, st>>>(W, nullptr, emb, L.edge_lo, C_out, C_in, k, emb_ci_lo, C_e, n_tab, L.m_rows, 0, ps, 0);
        fold_bias_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, st>>>(W, nullptr, emb, L.edge_hi, C_out, C_in, k, emb_ci_lo, C_e, n_tab, L.m_rows, 0, ps, k - 1);
        CUDA_TRY(cudaGetLastError());
    }
    return ZS_OK;
}

// transposed weights of layer L for its data-gradient GEMM (training handles only).
// mode 0: stride-1 conv / linear over input channels [0, ci_n); mode 1: stride-2 conv (pixel-shuffle-by-parity rows)
static int pack_layer_T(DevPool& pool, Layer& L, const float* W, int C_out, int C_in, int k, int ci_n, int mode, int ps_c,
                        cudaStream_t st) {
    L.t_valid = mode == 1 ? 2 * ci_n : ci_n;
    L.t_rows = round_up(L.t_valid, BM);
    L.t_taps = mode == 1 ? k / 2 + 1 : k;
    L.t_kpad = round_up(C_out, BK);
    L.t_kvalid = C_out;
    if (mode == 1 && ci_n % 64) return fail(ZS_ERR_ARG, "train: a stride-2 layer needs c_in %% 64 == 0 (got %d)", ci_n);
    const long long k_total = static_cast<long long>(L.t_taps) * L.t_kpad;
    ZS_TRY(pool.alloc(&L.wt, static_cast<size_t>(L.t_rows) * k_total * 2, st));
    CUDA_TRY(launch_pack_weight_T(W, static_cast<__half*>(L.wt), C_out, C_in, k, ci_n, k_total, L.t_kpad, mode, ps_c, st));
    return ZS_OK;
}

struct zs_encoder {
    zs_encoder_cfg cfg;
    DevPool pool;
    int n_out = 0;          // linear rows: enc_size or 2 * enc_size
    bool bank_merged = false;
    Layer bank[7];          // merged: bank[0] holds all 7 kernels
    Layer conv[7];          // conv2..conv8
    Layer dense[4];
    Layer gru_ih;           // both directions stacked: rows [0, 3H) forward, [3H, 6H) reverse
    float* whhT = nullptr;  // [2][H][3H] fp32
    float* bhh = nullptr;   // [2][3H]
    void* whh_img = nullptr;  // cluster-kernel shared-memory images (H % 64 == 0)
    Layer linear;
    const float* w_hh[2] = {nullptr, nullptr};   // training: the caller's fp32 W_hh (read by the CUDA-core BPTT kernel)
    void* whhT_img = nullptr;                    // training: streamed W_hh^T images of the cluster BPTT kernel
};

struct zs_decoder {
    zs_decoder_cfg cfg;
    DevPool pool;
    Layer input_emb;
    void* emb_table = nullptr;   // [c_in][c_h] operand type, for the unit-id gather
    Layer conv[6];
    Layer dense[4];
    Layer gru_ih;
    float* whhT = nullptr;
    float* bhh = nullptr;
    void* whh_img = nullptr;
    Layer dense5;
    Layer linear;
    const float* w_hh[2] = {nullptr, nullptr};   // training: the caller's fp32 parameters read directly by kernels
    void* whhT_img = nullptr;
    const float* emb[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

static int pack_gru(DevPool& pool, Layer& ih, float** whhT, float** bhh, void** whh_img, int operand, const float* const* w_ih,
                    const float* const* w_hh, const float* const* b_ih, const float* const* b_hh, int C, int H,
                    const float* emb, int n_spk, cudaStream_t st) {
    ih.m_rows = round_up(6 * H, BM); ih.m_valid = 6 * H; ih.taps = ih.w_taps = 1;
    ih.c_in_pad = round_up(C, BK); ih.c_in_valid = C; ih.per_spk = emb ? 1 : 0;
    ZS_TRY(pool.alloc(&ih.w, static_cast<size_t>(ih.m_rows) * ih.c_in_pad * 2, st));
    const int n_tab = emb ? n_spk : 1;
    ih.n_tab = n_tab;
    ZS_TRY(pool.alloc(reinterpret_cast<void**>(&ih.bias), static_cast<size_t>(n_tab) * ih.m_rows * 4, st));
    ZS_TRY(pool.alloc(reinterpret_cast<void**>(whhT), static_cast<size_t>(2) * 3 * H * H * 4, st));
    ZS_TRY(pool.alloc(reinterpret_cast<void**>(bhh), static_cast<size_t>(2) * 3 * H * 4, st));
    for (int dir = 0; dir < 2; ++dir) {
        if (operand == ZS_OPERAND_BF16)
            CUDA_TRY(launch_pack_weight(w_ih[dir], static_cast<__nv_bfloat16*>(ih.w), 3 * H, C, 1, 0, C, ih.c_in_pad, ih.c_in_pad, 0, dir * 3 * H, 0, st));
        else
            CUDA_TRY(launch_pack_weight(w_ih[dir], static_cast<__half*>(ih.w), 3 * H, C, 1, 0, C, ih.c_in_pad, ih.c_in_pad, 0, dir * 3 * H, 0, st));
        const long long warps = static_cast<long long>(n_tab) * 3 * H;
        fold_bias_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, st>>>(w_ih[dir], b_ih[dir], emb, ih.bias, 3 * H, C, 1, 0, emb ? C : 0, n_tab, ih.m_rows, dir * 3 * H, 0);
        transpose_whh_kernel<<<(3 * H * H + 255) / 256, 256, 0, st>>>(w_hh[dir], *whhT + static_cast<size_t>(dir) * 3 * H * H, H);
        CUDA_TRY(cudaMemcpyAsync(*bhh + static_cast<size_t>(dir) * 3 * H, b_hh[dir], static_cast<size_t>(3) * H * 4, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(cudaGetLastError());
    if (gru_cluster_ok(H)) {
        ZS_TRY(pool.alloc(whh_img, gru_images_bytes(H), st));
        ZS_TRY(pack_gru_image(*whh_img, w_hh, nullptr, H, operand, st));
    }
    return ZS_OK;
}

// streamed W_hh^T images of the cluster BPTT kernel (both directions), only when the cluster kernels apply
static int pack_gru_bptt_image(DevPool& pool, void** img, const float* const* w_hh, int H, cudaStream_t st) {
    if (!gru_cluster_ok(H)) return ZS_OK;
    const size_t per_dir = static_cast<size_t>(H / 64) * gb_w_image_bytes(H);
    ZS_TRY(pool.alloc(img, 2 * per_dir, st));
    for (int dir = 0; dir < 2; ++dir)
        gru_pack_whhT_kernel<<<(3 * H * H + 255) / 256, 256, 0, st>>>(w_hh[dir], reinterpret_cast<__half*>(static_cast<uint8_t*>(*img) + dir * per_dir), H);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

// W_ih of both directions, transposed: [C rows][6H] (K index = dir * 3H + gate row), for the GRU input's data gradient
static int pack_gru_T(DevPool& pool, Layer& ih, const float* const* w_ih, int C, int H, cudaStream_t st) {
    ih.t_valid = C; ih.t_rows = round_up(C, BM); ih.t_taps = 1; ih.t_kpad = round_up(6 * H, BK); ih.t_kvalid = 6 * H;
    ZS_TRY(pool.alloc(&ih.wt, static_cast<size_t>(ih.t_rows) * ih.t_kpad * 2, st));
    for (int dir = 0; dir < 2; ++dir) {
        // mode 0, k = 1: dst[ci][kk] with kk = co; shift the destination by dir * 3H columns
        CUDA_TRY(launch_pack_weight_T(w_ih[dir], static_cast<__half*>(ih.wt) + dir * 3 * H, 3 * H, C, 1, C, ih.t_kpad, ih.t_kpad, 0, 0, st));
    }
    return ZS_OK;
}

static int encoder_fill(zs_encoder* h, const zs_encoder_weights* w, cudaStream_t st) {
    const zs_encoder_cfg* cfg = &h->cfg;
    const int op = cfg->operand, c_in = cfg->c_in, h1 = cfg->c_h1, h2 = cfg->c_h2, h3 = cfg->c_h3;
    h->bank_merged = (h1 == BM);
    if (h->bank_merged) {   // 7 kernels in one [896][7 taps][c_in_pad] matrix, kernel k at tap 3 - k/2
        Layer& L = h->bank[0];
        L.m_rows = 7 * BM; L.m_valid = 7 * BM; L.taps = 7; L.w_taps = 7; L.c_in_pad = round_up(c_in, BK); L.c_in_valid = c_in;
        const long long k_total = 7LL * L.c_in_pad;
        ZS_TRY(h->pool.alloc(&L.w, static_cast<size_t>(L.m_rows) * k_total * 2, st));
        ZS_TRY(h->pool.alloc(reinterpret_cast<void**>(&L.bias), static_cast<size_t>(L.m_rows) * 4, st));
        for (int i = 0; i < 7; ++i) {
            const int k = i + 1;
            if (op == ZS_OPERAND_BF16)
                CUDA_TRY(launch_pack_weight(w->conv1s_w[i], static_cast<__nv_bfloat16*>(L.w), h1, c_in, k, 0, c_in, k_total, L.c_in_pad, 3 - k / 2, i * BM, 0, st));
            else
                CUDA_TRY(launch_pack_weight(w->conv1s_w[i], static_cast<__half*>(L.w), h1, c_in, k, 0, c_in, k_total, L.c_in_pad, 3 - k / 2, i * BM, 0, st));
            fold_bias_kernel<<<(h1 * 32 + 255) / 256, 256, 0, st>>>(w->conv1s_w[i], w->conv1s_b[i], nullptr, L.bias, h1, c_in, k, 0, 0, 1, L.m_rows, i * BM, 0);
        }
        CUDA_TRY(cudaGetLastError());
    } else {
        for (int i = 0; i < 7; ++i)
            ZS_TRY(pack_layer(h->pool, h->bank[i], op, w->conv1s_w[i], w->conv1s_b[i], h1, c_in, i + 1, 0, c_in, 0, nullptr, 0, 0, 1, st));
    }
    ZS_TRY(pack_layer(h->pool, h->conv[0], op, w->conv_w[0], w->conv_b[0], h2, 7 * h1 + c_in, 1, 0, 7 * h1 + c_in, 0, nullptr, 0, 0, 1, st));
    for (int i = 1; i < 7; ++i)
        ZS_TRY(pack_layer(h->pool, h->conv[i], op, w->conv_w[i], w->conv_b[i], h2, h2, 5, 0, h2, 0, nullptr, 0, 0, 1, st));
    for (int i = 0; i < 4; ++i)
        ZS_TRY(pack_layer(h->pool, h->dense[i], op, w->dense_w[i], w->dense_b[i], h2, h2, 1, 0, h2, 0, nullptr, 0, 0, 1, st));
    ZS_TRY(pack_gru(h->pool, h->gru_ih, &h->whhT, &h->bhh, &h->whh_img, op, w->gru_w_ih, w->gru_w_hh, w->gru_b_ih, w->gru_b_hh, h2, h3, nullptr, 1, st));
    ZS_TRY(pack_layer(h->pool, h->linear, op, w->linear_w, w->linear_b, h->n_out, h2 + 2 * h3, 1, 0, h2 + 2 * h3, 0, nullptr, 0, 0, 1, st));
    if (cfg->train) {   // data-gradient operands; the conv bank needs none (x takes no gradient)
        ZS_TRY(pack_layer_T(h->pool, h->conv[0], w->conv_w[0], h2, 7 * h1 + c_in, 1, 7 * h1, 0, 0, st));   // bank channels only
        for (int i = 1; i < 7; ++i)
            ZS_TRY(pack_layer_T(h->pool, h->conv[i], w->conv_w[i], h2, h2, 5, h2, (i % 2 == 0) ? 1 : 0, 0, st));   // conv4/6/8 are stride 2
        for (int i = 0; i < 4; ++i)
            ZS_TRY(pack_layer_T(h->pool, h->dense[i], w->dense_w[i], h2, h2, 1, h2, 0, 0, st));
        ZS_TRY(pack_gru_T(h->pool, h->gru_ih, w->gru_w_ih, h2, h3, st));
        ZS_TRY(pack_gru_bptt_image(h->pool, &h->whhT_img, w->gru_w_hh, h3, st));
        ZS_TRY(pack_layer_T(h->pool, h->linear, w->linear_w, h->n_out, h2 + 2 * h3, 1, h2 + 2 * h3, 0, 0, st));
        for (int d = 0; d < 2; ++d) h->w_hh[d] = w->gru_w_hh[d];
    }
    return ZS_OK;
}

extern "C" int zs_encoder_pack(const zs_encoder_cfg* cfg, const zs_encoder_weights* w, void* stream, zs_encoder** out) {
    if (!cfg || !w || !out) return fail(ZS_ERR_ARG, "encoder_pack: null argument");
    ZS_TRY(ensure_device());
    if (cfg->seg_len < 64 && cfg->train) return fail(ZS_ERR_ARG, "encoder: seg_len %d < 64 selects zero padding (model/model.py:38); the training path implements the reflect mode", cfg->seg_len);
    if (cfg->enc_mode < 0 || cfg->enc_mode > 4) return fail(ZS_ERR_ARG, "encoder: enc_mode %d not supported", cfg->enc_mode);
    if (cfg->enc_mode == ZS_ENC_BINARY && cfg->enc_size > 128) return fail(ZS_ERR_ARG, "encoder: enc_mode 'binary' projects to enc_size^2 channels; enc_size %d > 128", cfg->enc_size);
    if (cfg->c_h2 % 8 || cfg->c_h1 % 8 || cfg->c_h3 < 1) return fail(ZS_ERR_ARG, "encoder: c_h1/c_h2 must be multiples of 8");
    if (cfg->train && cfg->operand != ZS_OPERAND_FP16) return fail(ZS_ERR_ARG, "encoder: the training path computes in fp16 operands (loss-scaled gradients)");
    if (cfg->train && cfg->enc_mode != ZS_ENC_ONE_HOT) return fail(ZS_ERR_ARG, "encoder: the training path implements enc_mode 'one_hot' only");
    if (cfg->train && (cfg->c_h2 % 64 || cfg->c_h3 % 8)) return fail(ZS_ERR_ARG, "encoder: training needs c_h2 %% 64 == 0 and c_h3 %% 8 == 0");
    zs_encoder* h = new zs_encoder();
    h->cfg = *cfg;
    h->n_out = cfg->enc_mode == ZS_ENC_MULTILABEL_BINARY ? 2 * cfg->enc_size : (cfg->enc_mode == ZS_ENC_BINARY ? cfg->enc_size * cfg->enc_size : cfg->enc_size);
    const int rc = encoder_fill(h, w, static_cast<cudaStream_t>(stream));
    if (rc != ZS_OK) {
        h->pool.release();
        delete h;
        return rc;
    }
    *out = h;
    return ZS_OK;
}

/* refresh every packed buffer from the (updated) fp32 parameters; same pointers semantics as zs_encoder_pack */
extern "C" int zs_encoder_repack(zs_encoder* h, const zs_encoder_weights* w, void* stream) {
    if (!h || !w) return fail(ZS_ERR_ARG, "encoder_repack: null argument");
    return encoder_fill(h, w, static_cast<cudaStream_t>(stream));
}

extern "C" void zs_encoder_free(zs_encoder* h) {
    if (!h) return;
    h->pool.release();
    delete h;
}

static int decoder_fill(zs_decoder* h, const zs_decoder_weights* w, cudaStream_t st) {
    const zs_decoder_cfg* cfg = &h->cfg;
    const int op = cfg->operand, ch = cfg->c_h, ca = cfg->c_a;
    const bool tr = cfg->train != 0;     // training: speaker embeddings are added to the activations, not folded
    ZS_TRY(pack_layer(h->pool, h->input_emb, op, w->input_emb_w, w->input_emb_b, ch, cfg->c_in, 1, 0, cfg->c_in, 0, nullptr, 0, 0, 1, st));
    ZS_TRY(h->pool.alloc(&h->emb_table, static_cast<size_t>(cfg->c_in) * ch * 2, st));
    {
        const long long total = static_cast<long long>(ch) * cfg->c_in;
        const int blocks = static_cast<int>((total + 255) / 256);
        if (op == ZS_OPERAND_BF16) transpose_emb_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w->input_emb_w, static_cast<__nv_bfloat16*>(h->emb_table), ch, cfg->c_in);
        else transpose_emb_kernel<__half><<<blocks, 256, 0, st>>>(w->input_emb_w, static_cast<__half*>(h->emb_table), ch, cfg->c_in);
        CUDA_TRY(cudaGetLastError());
    }
    for (int i = 0; i < 6; ++i) {   // conv1,3,5: 2*c_h rows pixel-shuffle-permuted; block b uses emb[b] for both convs
        const bool up = (i % 2 == 0);
        ZS_TRY(pack_layer(h->pool, h->conv[i], op, w->conv_w[i], w->conv_b[i], up ? 2 * ch : ch, ch, 3, 0, ch, up ? 1 : 0, tr ? nullptr : w->emb[i / 2], 0, ch, ca, st, cfg->seg_len < 64));
    }
    for (int i = 0; i < 4; ++i)     // emb4 conditions all four dense layers (model/model.py:350-351)
        ZS_TRY(pack_layer(h->pool, h->dense[i], op, w->dense_w[i], w->dense_b[i], ch, ch, 1, 0, ch, 0, tr ? nullptr : w->emb[3], 0, ch, ca, st));
    ZS_TRY(pack_gru(h->pool, h->gru_ih, &h->whhT, &h->bhh, &h->whh_img, op, w->gru_w_ih, w->gru_w_hh, w->gru_b_ih, w->gru_b_hh, ch, ch / 2, tr ? nullptr : w->emb[4], ca, st));
    // dense5 sees cat([out, rnn, emb5]): inference sends the first 2*c_h inputs through the GEMM and folds the last c_h
    // into the bias; training keeps all 3*c_h as real input channels (append_emb materialised)
    if (tr) ZS_TRY(pack_layer(h->pool, h->dense5, op, w->dense5_w, w->dense5_b, ch, 3 * ch, 1, 0, 3 * ch, 0, nullptr, 0, 0, 1, st));
    else ZS_TRY(pack_layer(h->pool, h->dense5, op, w->dense5_w, w->dense5_b, ch, 3 * ch, 1, 0, 2 * ch, 0, w->emb[4], 2 * ch, ch, ca, st));
    ZS_TRY(pack_layer(h->pool, h->linear, op, w->linear_w, w->linear_b, cfg->c_out, ch, 1, 0, ch, 0, nullptr, 0, 0, 1, st));
    if (tr) {
        ZS_TRY(pack_layer_T(h->pool, h->input_emb, w->input_emb_w, ch, cfg->c_in, 1, cfg->c_in, 0, 0, st));
        for (int i = 0; i < 6; ++i) {
            const bool up = (i % 2 == 0);
            ZS_TRY(pack_layer_T(h->pool, h->conv[i], w->conv_w[i], up ? 2 * ch : ch, ch, 3, ch, 0, up ? ch : 0, st));
        }
        for (int i = 0; i < 4; ++i) ZS_TRY(pack_layer_T(h->pool, h->dense[i], w->dense_w[i], ch, ch, 1, ch, 0, 0, st));
        ZS_TRY(pack_gru_T(h->pool, h->gru_ih, w->gru_w_ih, ch, ch / 2, st));
        ZS_TRY(pack_gru_bptt_image(h->pool, &h->whhT_img, w->gru_w_hh, ch / 2, st));
        ZS_TRY(pack_layer_T(h->pool, h->dense5, w->dense5_w, ch, 3 * ch, 1, 3 * ch, 0, 0, st));
        ZS_TRY(pack_layer_T(h->pool, h->linear, w->linear_w, cfg->c_out, ch, 1, ch, 0, 0, st));
        for (int d = 0; d < 2; ++d) h->w_hh[d] = w->gru_w_hh[d];
        for (int i = 0; i < 5; ++i) h->emb[i] = w->emb[i];
    }
    return ZS_OK;
}

extern "C" int zs_decoder_pack(const zs_decoder_cfg* cfg, const zs_decoder_weights* w, void* stream, zs_decoder** out) {
    if (!cfg || !w || !out) return fail(ZS_ERR_ARG, "decoder_pack: null argument");
    ZS_TRY(ensure_device());
    if (cfg->seg_len < 64 && cfg->train) return fail(ZS_ERR_ARG, "decoder: seg_len %d < 64 selects zero padding (model/model.py:38); the training path implements the reflect mode", cfg->seg_len);
    if (cfg->c_h % 64) return fail(ZS_ERR_ARG, "decoder: c_h %d must be a multiple of 64 (pixel-shuffle tile permutation)", cfg->c_h);
    if (cfg->train && cfg->operand != ZS_OPERAND_FP16) return fail(ZS_ERR_ARG, "decoder: the training path computes in fp16 operands (loss-scaled gradients)");
    if (cfg->train && cfg->c_in % 8) return fail(ZS_ERR_ARG, "decoder: training needs c_in %% 8 == 0");
    zs_decoder* h = new zs_decoder();
    h->cfg = *cfg;
    const int rc = decoder_fill(h, w, static_cast<cudaStream_t>(stream));
    if (rc != ZS_OK) {
        h->pool.release();
        delete h;
        return rc;
    }
    *out = h;
    return ZS_OK;
}

extern "C" int zs_decoder_repack(zs_decoder* h, const zs_decoder_weights* w, void* stream) {
    if (!h || !w) return fail(ZS_ERR_ARG, "decoder_repack: null argument");
    return decoder_fill(h, w, static_cast<cudaStream_t>(stream));
}

extern "C" void zs_decoder_free(zs_decoder* h) {
    if (!h) return;
    h->pool.release();
    delete h;
}

// -------------------------------------------------------------------------------------------------
// workspace carving
// -------------------------------------------------------------------------------------------------
struct Buf {            // channels-last activation buffer [B][rows][pitch]
    void* p = nullptr;
    int rows = 0, pitch = 0, halo = 0, T = 0;
};
struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
    void* take(size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    }
    Buf act(int B, int T, int halo, int channels, bool exact_rows = false) {
        Buf b;
        b.rows = exact_rows ? T + 2 * halo : buf_rows(T, halo); b.pitch = round_up(channels, 8); b.halo = halo; b.T = T;
        b.p = take(static_cast<size_t>(B) * b.rows * b.pitch * 2);
        return b;
    }
};

struct EncWs {
    Buf xp, cat, a[7], d[3], catr, gx, xch;
    float* logits = nullptr;        // (B, n_out, T8) fp32 scratch for callers that do not want the logits back
    int T[4];
    size_t bytes;
};
static EncWs carve_encoder(const zs_encoder* h, void* ws, int B, int T) {
    EncWs w;
    Carver c(ws);
    const zs_encoder_cfg& g = h->cfg;
    w.T[0] = T; w.T[1] = (T + 1) / 2; w.T[2] = (w.T[1] + 1) / 2; w.T[3] = (w.T[2] + 1) / 2;
    w.xp = c.act(B, T, 3, g.c_in);
    w.cat = c.act(B, T, 0, 7 * g.c_h1 + g.c_in);
    w.a[0] = c.act(B, w.T[0], 2, g.c_h2);   // conv2 out
    w.a[1] = c.act(B, w.T[0], 2, g.c_h2);   // conv3 out
    w.a[2] = c.act(B, w.T[1], 2, g.c_h2);   // conv4 out
    w.a[3] = c.act(B, w.T[1], 2, g.c_h2);   // conv5 out
    w.a[4] = c.act(B, w.T[2], 2, g.c_h2);   // conv6 out
    w.a[5] = c.act(B, w.T[2], 2, g.c_h2);   // conv7 out
    w.a[6] = c.act(B, w.T[3], 0, g.c_h2);   // conv8 out
    for (int i = 0; i < 3; ++i) w.d[i] = c.act(B, w.T[3], 0, g.c_h2);
    w.catr = c.act(B, w.T[3], 0, g.c_h2 + 2 * g.c_h3);
    w.gx = c.act(B, w.T[3], 0, 6 * g.c_h3, true);   // the recurrence indexes it as a dense [B][T][2][3H] array
    w.xch = c.act(round_up(B, 128), 2, 0, g.c_h3);   // GRU state exchange scratch: one H-wide fp16 row per (direction, sequence)
    w.logits = static_cast<float*>(c.take(static_cast<size_t>(B) * h->n_out * w.T[3] * sizeof(float)));
    w.bytes = c.off;
    return w;
}
extern "C" size_t zs_encoder_workspace_bytes(const zs_encoder* h, int B, int T) {
    if (!h || B < 1 || T < 1) return 0;
    return carve_encoder(h, nullptr, B, T).bytes;
}

struct DecWs {
    Buf actp, x0, p[3], y[3], d[3], catr, d5, gx, xch;
    size_t bytes;
};
static DecWs carve_decoder(const zs_decoder* h, void* ws, int B, int T8) {
    DecWs w;
    Carver c(ws);
    const int ch = h->cfg.c_h;
    w.actp = c.act(B, T8, 0, h->cfg.c_in);
    w.x0 = c.act(B, T8, 1, ch);
    for (int i = 0; i < 3; ++i) {
        const int To = T8 << (i + 1);
        w.p[i] = c.act(B, To, 1, ch);
        w.y[i] = c.act(B, To, i == 2 ? 0 : 1, ch);
    }
    const int Tf = 8 * T8;
    for (int i = 0; i < 3; ++i) w.d[i] = c.act(B, Tf, 0, ch);
    w.catr = c.act(B, Tf, 0, 2 * ch);
    w.d5 = c.act(B, Tf, 0, ch);
    w.gx = c.act(B, Tf, 0, 3 * ch, true);
    w.xch = c.act(round_up(B, 128), 2, 0, ch / 2);   // GRU state exchange scratch
    w.bytes = c.off;
    return w;
}
extern "C" size_t zs_decoder_workspace_bytes(const zs_decoder* h, int B, int T8) {
    if (!h || B < 1 || T8 < 1) return 0;
    return carve_decoder(h, nullptr, B, T8).bytes;
}

// -------------------------------------------------------------------------------------------------
// layer sequencing
// -------------------------------------------------------------------------------------------------
struct ConvOpts {
    int stride = 1, lrelu = 1, inorm = 0, res_mode = RES_NONE, act = ACT_NONE, out_mode = OUT_CL, out_choff = 0,
        accumulate = 0, bank = 0, c_in_valid = -1, out_f16 = 0;
    const Buf* res = nullptr;
    const int64_t* spk = nullptr;
    int in_row0 = -1000;           // override (data-gradient GEMMs read zero-padded gradient buffers from row 0)
    const ConvExtras* ex = nullptr;
};
// runs layer L on `in`, writing T_out frames per segment into `out` (a Buf, or raw fp32 (B, C, T) for NCT32)
static int run_layer(const Layer& L, int operand, float ns, const Buf& in, int B, int T_out, const Buf* out, void* out_raw,
                     int out_raw_rows, int out_raw_pitch, const ConvOpts& o, cudaStream_t st) {
    zs_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.w = L.w; d.m_rows = L.m_rows; d.m_valid = L.m_valid; d.taps = L.taps; d.c_in_pad = L.c_in_pad; d.w_taps = L.w_taps;
    d.bank = o.bank;
    d.in = in.p; d.in_rows = in.rows; d.in_pitch = in.pitch;
    const int pad_left = o.bank ? 3 : L.taps / 2;
    d.in_row0 = o.in_row0 != -1000 ? o.in_row0 : in.halo - pad_left;
    if (d.in_row0 < 0) return fail(ZS_ERR_ARG, "layer: input halo %d < pad %d", in.halo, pad_left);
    d.c_in_valid = o.c_in_valid >= 0 ? o.c_in_valid : L.c_in_valid;
    d.stride = o.stride; d.B = B; d.T_out = T_out;
    d.bias = L.bias; d.spk = L.per_spk ? o.spk : nullptr; d.n_spk = L.n_tab;
    if (L.per_spk && !o.spk) return fail(ZS_ERR_ARG, "layer: speaker ids required");
    d.lrelu = o.lrelu; d.ns = ns; d.inorm = o.inorm;
    d.res_mode = o.res_mode;
    if (o.res) { d.res = o.res->p; d.res_rows = o.res->rows; d.res_pitch = o.res->pitch; d.res_halo = o.res->halo; }
    d.act = o.act; d.out_mode = o.out_mode;
    if (out) { d.out = out->p; d.out_rows = out->rows; d.out_pitch = out->pitch; d.out_halo = out->halo; }
    else { d.out = out_raw; d.out_rows = out_raw_rows; d.out_pitch = out_raw_pitch; d.out_halo = 0; }
    d.out_choff = o.out_choff; d.accumulate = o.accumulate; d.operand = operand; d.nb_hint = 0; d.out_f16 = o.out_f16;
    if (!t_zero_pad) return launch_conv(&d, st, o.ex);
    ConvExtras ex = o.ex ? *o.ex : ConvExtras();
    ex.zero_halo = 1; ex.edge_lo = L.edge_lo; ex.edge_hi = L.edge_hi;
    return launch_conv(&d, st, &ex);
}

// -------------------------------------------------------------------------------------------------
// The recurrent tail shared by the Decoder (model/model.py:350-364) and the Spectrogram_Patcher (model/model.py:536-549):
// two dense blocks conditioned on one embedding, bi-GRU on out + a second embedding, dense5 on cat([out, rnn, emb]) ->
// lrelu -> linear -> sigmoid | tanh.  All speaker terms are folded into per-speaker bias tables at pack time.
// -------------------------------------------------------------------------------------------------
struct TailNet {
    const Layer* dense;      // [4]
    const Layer* gru_ih;
    const float* whhT; const float* bhh; const void* whh_img;
    const Layer* dense5; const Layer* linear;
    int ch, output_mask, op; float ns;
};
struct TailBufs { const Buf* in; Buf* d; Buf* catr; Buf* d5; Buf* gx; Buf* xch; };
static int run_tail(const TailNet& n, const TailBufs& w, const int64_t* spk, int B, int Tf, void* spec, int accumulate, cudaStream_t st, int spec_f16 = 0) {
    const int op = n.op, ch = n.ch;
    const float ns = n.ns;
    {   // two dense blocks, both conditioned on the same embedding (Decoder: emb4 twice, model/model.py:350-351)
        ConvOpts o; o.spk = spk;
        ZS_TRY(run_layer(n.dense[0], op, ns, *w.in, B, Tf, &w.d[0], nullptr, 0, 0, o, st));
        ConvOpts r; r.inorm = 1; r.res_mode = RES_SAME; r.res = w.in; r.spk = spk;
        ZS_TRY(run_layer(n.dense[1], op, ns, w.d[0], B, Tf, &w.d[1], nullptr, 0, 0, r, st));
        ZS_TRY(run_layer(n.dense[2], op, ns, w.d[1], B, Tf, &w.d[2], nullptr, 0, 0, o, st));
        ConvOpts r2; r2.inorm = 1; r2.res_mode = RES_SAME; r2.res = &w.d[1]; r2.spk = spk;
        ZS_TRY(run_layer(n.dense[3], op, ns, w.d[2], B, Tf, w.catr, nullptr, 0, 0, r2, st));
    }
    {   // bi-GRU on out + emb (model/model.py:352-355)
        ConvOpts o; o.lrelu = 0; o.c_in_valid = ch; o.spk = spk;
        ZS_TRY(run_layer(*n.gru_ih, op, ns, *w.catr, B, Tf, w.gx, nullptr, 0, 0, o, st));
        if (n.whh_img) ZS_TRY(launch_gru_cluster(n.whh_img, n.bhh, w.gx->p, B, Tf, ch / 2, w.catr->p, w.catr->rows, w.catr->pitch, 0, ch, op, st, nullptr, w.xch->p, static_cast<size_t>(w.xch->rows) * w.xch->pitch * 2 * round_up(B, 128)));
        else ZS_TRY(launch_gru(w.gx->p, n.whhT, n.bhh, B, Tf, ch / 2, w.catr->p, w.catr->rows, w.catr->pitch, 0, ch, op, st));
    }
    {   // dense5 on cat([out, rnn, emb]) -> lrelu -> linear -> sigmoid | tanh (model/model.py:356-364)
        ConvOpts o; o.spk = spk;
        ZS_TRY(run_layer(*n.dense5, op, ns, *w.catr, B, Tf, w.d5, nullptr, 0, 0, o, st));
        ConvOpts f; f.lrelu = 0; f.act = n.output_mask ? ACT_TANH : ACT_SIGMOID; f.out_mode = OUT_NCT32; f.accumulate = accumulate; f.out_f16 = spec_f16;
        ZS_TRY(run_layer(*n.linear, op, ns, *w.d5, B, Tf, nullptr, spec, 0, 0, f, st));
    }
    return ZS_OK;
}

extern "C" int zs_encoder_forward(zs_encoder* h, const float* x, int B, int T, const float* gumbel_noise, float* logits,
                                  float* act, int32_t* unit_ids, void* workspace, size_t workspace_bytes, void* stream) {
    if (!logits) return fail(ZS_ERR_ARG, "encoder_forward: null argument");
    if (h && h->cfg.enc_mode != ZS_ENC_CONTINUES && !gumbel_noise) return fail(ZS_ERR_ARG, "encoder_forward: enc_mode %d needs the Gumbel noise tensor", h->cfg.enc_mode);
    return zs_encoder_forward_x(h, x, ZS_X_F32, ZS_X_NCT, B, T, gumbel_noise, nullptr, logits, act, unit_ids, workspace, workspace_bytes, stream);
}

extern "C" int zs_encoder_forward_x(zs_encoder* h, const void* x, int x_dtype, int x_layout, int B, int T, const float* gumbel_noise,
                                    const uint64_t* noise_seeds, float* logits, float* act, int32_t* unit_ids, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (!h || !x) return fail(ZS_ERR_ARG, "encoder_forward: null argument");
    if (B < 1) return fail(ZS_ERR_ARG, "encoder_forward: B %d", B);
    if (T < 9 || T > 256) return fail(ZS_ERR_ARG, "encoder_forward: T %d outside [9, 256] (convert.py MIN_LEN=9; segments are < 2*seg_len frames)", T);
    const zs_encoder_cfg& g = h->cfg;
    if (g.enc_mode != ZS_ENC_CONTINUES && !gumbel_noise && !(g.enc_mode == ZS_ENC_ONE_HOT && noise_seeds))
        return fail(ZS_ERR_ARG, "encoder_forward: enc_mode %d needs the Gumbel noise tensor (device-generated noise: one_hot with noise_seeds)", g.enc_mode);
    if (g.enc_mode != ZS_ENC_ONE_HOT && !logits && !act) return fail(ZS_ERR_ARG, "encoder_forward: no output requested");
    t_zero_pad = g.seg_len < 64;           // model/model.py:36-38: 'constant' padding below seg_len 64
    EncWs w = carve_encoder(h, workspace, B, T);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "encoder_forward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int op = g.operand;
    const float ns = g.ns;

    // model/model.py:441-446: conv bank on x, concatenated with x, leaky-relu
    if (!logits) logits = w.logits;
    ZS_TRY(launch_pack_x_dual(x, x_dtype, x_layout, B, g.c_in, T, w.xp.p, w.xp.rows, w.xp.pitch, 3, w.cat.p, w.cat.rows, w.cat.pitch, 7 * g.c_h1, ns, op, st));
    if (h->bank_merged) {
        ConvOpts o; o.bank = 1;
        ZS_TRY(run_layer(h->bank[0], op, ns, w.xp, B, T, &w.cat, nullptr, 0, 0, o, st));
    } else {
        for (int i = 0; i < 7; ++i) {
            ConvOpts o; o.out_choff = i * g.c_h1;
            // every kernel reads the halo-3 buffer; its own left pad is (i+1)/2
            ZS_TRY(run_layer(h->bank[i], op, ns, w.xp, B, T, &w.cat, nullptr, 0, 0, o, st));
        }
    }
    {   // :447 conv2 -> lrelu -> IN (dropout is identity in eval)
        ConvOpts o; o.inorm = 1;
        ZS_TRY(run_layer(h->conv[0], op, ns, w.cat, B, T, &w.a[0], nullptr, 0, 0, o, st));
    }
    for (int blk = 0; blk < 3; ++blk) {   // :448-450 three (conv k5, conv k5 stride 2) blocks with avg-pool residual
        const Buf& xin = w.a[2 * blk];
        ConvOpts o1;
        ZS_TRY(run_layer(h->conv[1 + 2 * blk], op, ns, xin, B, w.T[blk], &w.a[2 * blk + 1], nullptr, 0, 0, o1, st));
        ConvOpts o2; o2.stride = 2; o2.inorm = 1; o2.res_mode = RES_AVG2; o2.res = &xin;
        ZS_TRY(run_layer(h->conv[2 + 2 * blk], op, ns, w.a[2 * blk + 1], B, w.T[blk + 1], &w.a[2 * blk + 2], nullptr, 0, 0, o2, st));
    }
    const int T8 = w.T[3];
    {   // :452-453 two dense blocks
        ConvOpts o;
        ZS_TRY(run_layer(h->dense[0], op, ns, w.a[6], B, T8, &w.d[0], nullptr, 0, 0, o, st));
        ConvOpts r; r.inorm = 1; r.res_mode = RES_SAME; r.res = &w.a[6];
        ZS_TRY(run_layer(h->dense[1], op, ns, w.d[0], B, T8, &w.d[1], nullptr, 0, 0, r, st));
        ZS_TRY(run_layer(h->dense[2], op, ns, w.d[1], B, T8, &w.d[2], nullptr, 0, 0, o, st));
        ConvOpts r2; r2.inorm = 1; r2.res_mode = RES_SAME; r2.res = &w.d[1];
        ZS_TRY(run_layer(h->dense[3], op, ns, w.d[2], B, T8, &w.catr, nullptr, 0, 0, r2, st));
    }
    {   // :454-455 bi-GRU: input projection on tensor cores, then the recurrence
        ConvOpts o; o.lrelu = 0; o.c_in_valid = g.c_h2;
        ZS_TRY(run_layer(h->gru_ih, op, ns, w.catr, B, T8, &w.gx, nullptr, 0, 0, o, st));
        if (h->whh_img) ZS_TRY(launch_gru_cluster(h->whh_img, h->bhh, w.gx.p, B, T8, g.c_h3, w.catr.p, w.catr.rows, w.catr.pitch, 0, g.c_h2, op, st, nullptr, w.xch.p, static_cast<size_t>(w.xch.rows) * w.xch.pitch * 2 * round_up(B, 128)));
        else ZS_TRY(launch_gru(w.gx.p, h->whhT, h->bhh, B, T8, g.c_h3, w.catr.p, w.catr.rows, w.catr.pitch, 0, g.c_h2, op, st));
    }
    {   // linear -> logits in the reference's (B, n_out, T8) fp32 layout
        ConvOpts o; o.lrelu = 0; o.out_mode = OUT_NCT32;
        ZS_TRY(run_layer(h->linear, op, ns, w.catr, B, T8, nullptr, logits, 0, 0, o, st));
    }
    if (g.enc_mode == ZS_ENC_ONE_HOT) {
        ZS_TRY(launch_onehot(logits, gumbel_noise, noise_seeds, B, g.enc_size, T8, act, unit_ids, st));
    } else if (act) {
        const size_t n = g.enc_mode == ZS_ENC_GUMBEL_T ? static_cast<size_t>(B) * g.enc_size : static_cast<size_t>(B) * g.enc_size * T8;
        if (g.enc_mode == ZS_ENC_BINARY) CUDA_TRY(cudaMemsetAsync(act, 0, n * sizeof(float), st));
        LaunchScope scope(st, KC_OTHER, 0.0, "bottleneck_misc_kernel");
        bottleneck_misc_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(logits, gumbel_noise, g.enc_mode, B, g.enc_size, T8, ns, act);
        CUDA_TRY(cudaGetLastError());
    }
    return ZS_OK;
}

extern "C" int zs_decoder_forward(zs_decoder* h, const float* enc_act, const int32_t* unit_ids, const int64_t* spk, int B,
                                  int T8, float* spec, int accumulate, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    return zs_decoder_forward_x(h, enc_act, unit_ids, spk, B, T8, spec, ZS_X_F32, accumulate, workspace, workspace_bytes, stream);
}

extern "C" int zs_decoder_forward_x(zs_decoder* h, const float* enc_act, const int32_t* unit_ids, const int64_t* spk, int B,
                                    int T8, void* spec, int spec_dtype, int accumulate, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    if (!h || !spk || !spec || (!enc_act && !unit_ids)) return fail(ZS_ERR_ARG, "decoder_forward: null argument");
    if (spec_dtype != ZS_X_F32 && spec_dtype != ZS_X_F16) return fail(ZS_ERR_ARG, "decoder_forward: spec_dtype %d", spec_dtype);
    if (B < 1 || T8 < 2 || T8 > 32) return fail(ZS_ERR_ARG, "decoder_forward: B %d, T8 %d (T8 must be in [2, 32])", B, T8);
    if (accumulate < 0 || accumulate > 2) return fail(ZS_ERR_ARG, "decoder_forward: accumulate %d", accumulate);
    const zs_decoder_cfg& g = h->cfg;
    t_zero_pad = g.seg_len < 64;           // model/model.py:36-38: 'constant' padding below seg_len 64
    DecWs w = carve_decoder(h, workspace, B, T8);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "decoder_forward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int op = g.operand, ch = g.c_h;
    const float ns = g.ns;

    if (unit_ids) {   // one-hot input: input_emb is a column gather (model/model.py:346)
        // 128 threads x 8 channels cover a 1024-channel row; 4 unit frames per CTA
        const dim3 blk(128, 4), grid((T8 + 3) / 4, B);
        LaunchScope scope(st, KC_OTHER, 0.0, "unit_gather_kernel");
        if (op == ZS_OPERAND_BF16)
            unit_gather_kernel<__nv_bfloat16><<<grid, blk, 0, st>>>(unit_ids, static_cast<const __nv_bfloat16*>(h->emb_table), h->input_emb.bias, static_cast<__nv_bfloat16*>(w.x0.p), T8, ch, w.x0.rows, w.x0.pitch, 1, g.c_in, t_zero_pad);
        else
            unit_gather_kernel<__half><<<grid, blk, 0, st>>>(unit_ids, static_cast<const __half*>(h->emb_table), h->input_emb.bias, static_cast<__half*>(w.x0.p), T8, ch, w.x0.rows, w.x0.pitch, 1, g.c_in, t_zero_pad);
        CUDA_TRY(cudaGetLastError());
    } else {
        ZS_TRY(launch_pack_nct(enc_act, B, g.c_in, T8, w.actp.p, w.actp.rows, w.actp.pitch, 0, 0, 0, ns, op, 0, st));
        ConvOpts o; o.lrelu = 0;
        ZS_TRY(run_layer(h->input_emb, op, ns, w.actp, B, T8, &w.x0, nullptr, 0, 0, o, st));
    }
    const Buf* xin = &w.x0;
    for (int blk = 0; blk < 3; ++blk) {   // model/model.py:317-331, speaker embedding folded into the bias tables
        const int Ti = T8 << blk;
        ConvOpts o1; o1.out_mode = OUT_PS; o1.spk = spk;
        ZS_TRY(run_layer(h->conv[2 * blk], op, ns, *xin, B, Ti, &w.p[blk], nullptr, 0, 0, o1, st));
        ConvOpts o2; o2.inorm = 1; o2.res_mode = RES_UP2; o2.res = xin; o2.spk = spk;
        ZS_TRY(run_layer(h->conv[2 * blk + 1], op, ns, w.p[blk], B, 2 * Ti, &w.y[blk], nullptr, 0, 0, o2, st));
        xin = &w.y[blk];
    }
    const int Tf = 8 * T8;
    TailNet net{h->dense, &h->gru_ih, h->whhT, h->bhh, h->whh_img, &h->dense5, &h->linear, ch, g.output_mask, op, ns};
    TailBufs tb{&w.y[2], w.d, &w.catr, &w.d5, &w.gx, &w.xch};
    return run_tail(net, tb, spk, B, Tf, spec, accumulate, st, spec_dtype == ZS_X_F16);
}

// -------------------------------------------------------------------------------------------------
// Spectrogram_Patcher (model/model.py:503-549; Trainer g_mode 'spectrogram', trainer.py:78-79, 212-213)
// -------------------------------------------------------------------------------------------------
struct zs_patcher {
    zs_patcher_cfg cfg;
    DevPool pool;
    Layer input, dense[4], gru_ih, dense5, linear;
    float* whhT = nullptr;
    float* bhh = nullptr;
    void* whh_img = nullptr;
};

extern "C" int zs_patcher_pack(const zs_patcher_cfg* cfg, const zs_patcher_weights* w, void* stream, zs_patcher** out) {
    if (!cfg || !w || !out) return fail(ZS_ERR_ARG, "patcher_pack: null argument");
    ZS_TRY(ensure_device());
    if (cfg->c_h % 64) return fail(ZS_ERR_ARG, "patcher: c_h %d must be a multiple of 64", cfg->c_h);
    zs_patcher* h = new zs_patcher();
    h->cfg = *cfg;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int op = cfg->operand, ch = cfg->c_h, ca = cfg->c_a;
    int rc = pack_layer(h->pool, h->input, op, w->input_w, w->input_b, ch, cfg->c_in, 1, 0, cfg->c_in, 0, nullptr, 0, 0, 1, st);
    for (int i = 0; i < 4 && rc == ZS_OK; ++i)     // emb1 conditions all four dense layers (model/model.py:538-539)
        rc = pack_layer(h->pool, h->dense[i], op, w->dense_w[i], w->dense_b[i], ch, ch, 1, 0, ch, 0, w->emb[0], 0, ch, ca, st);
    if (rc == ZS_OK) rc = pack_gru(h->pool, h->gru_ih, &h->whhT, &h->bhh, &h->whh_img, op, w->gru_w_ih, w->gru_w_hh, w->gru_b_ih, w->gru_b_hh, ch, ch / 2, w->emb[1], ca, st);
    if (rc == ZS_OK) rc = pack_layer(h->pool, h->dense5, op, w->dense5_w, w->dense5_b, ch, 3 * ch, 1, 0, 2 * ch, 0, w->emb[1], 2 * ch, ch, ca, st);
    if (rc == ZS_OK) rc = pack_layer(h->pool, h->linear, op, w->linear_w, w->linear_b, cfg->c_out, ch, 1, 0, ch, 0, nullptr, 0, 0, 1, st);
    if (rc != ZS_OK) {
        h->pool.release();
        delete h;
        return rc;
    }
    *out = h;
    return ZS_OK;
}
extern "C" void zs_patcher_free(zs_patcher* h) {
    if (!h) return;
    h->pool.release();
    delete h;
}

struct PatWs { Buf xp, x0, d[3], catr, d5, gx, xch; size_t bytes; };
static PatWs carve_patcher(const zs_patcher* h, void* ws, int B, int T) {
    PatWs w;
    Carver c(ws);
    const int ch = h->cfg.c_h;
    w.xp = c.act(B, T, 0, h->cfg.c_in);
    w.x0 = c.act(B, T, 0, ch);
    for (int i = 0; i < 3; ++i) w.d[i] = c.act(B, T, 0, ch);
    w.catr = c.act(B, T, 0, 2 * ch);
    w.d5 = c.act(B, T, 0, ch);
    w.gx = c.act(B, T, 0, 3 * ch, true);
    w.xch = c.act(round_up(B, 128), 2, 0, ch / 2);
    w.bytes = c.off;
    return w;
}
extern "C" size_t zs_patcher_workspace_bytes(const zs_patcher* h, int B, int T) {
    if (!h || B < 1 || T < 1) return 0;
    return carve_patcher(h, nullptr, B, T).bytes;
}

extern "C" int zs_patcher_forward(zs_patcher* h, const float* x, const int64_t* spk, int B, int T, float* spec, int accumulate,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !x || !spk || !spec) return fail(ZS_ERR_ARG, "patcher_forward: null argument");
    if (B < 1 || T < 1 || T > 256) return fail(ZS_ERR_ARG, "patcher_forward: B %d, T %d (T must be in [1, 256])", B, T);
    if (accumulate < 0 || accumulate > 2) return fail(ZS_ERR_ARG, "patcher_forward: accumulate %d", accumulate);
    const zs_patcher_cfg& g = h->cfg;
    t_zero_pad = 0;                        // every layer is per-frame (k = 1): pad_layer never pads here
    PatWs w = carve_patcher(h, workspace, B, T);
    if (!workspace || workspace_bytes < w.bytes) return fail(ZS_ERR_WORKSPACE, "patcher_forward: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // model/model.py:533-534: out = linear(x, input_layer) - no activation
    ZS_TRY(launch_pack_nct(x, B, g.c_in, T, w.xp.p, w.xp.rows, w.xp.pitch, 0, 0, 0, g.ns, g.operand, 1, st));
    ConvOpts o; o.lrelu = 0;
    ZS_TRY(run_layer(h->input, g.operand, g.ns, w.xp, B, T, &w.x0, nullptr, 0, 0, o, st));
    TailNet net{h->dense, &h->gru_ih, h->whhT, h->bhh, h->whh_img, &h->dense5, &h->linear, g.c_h, 0, g.operand, g.ns};
    TailBufs tb{&w.x0, w.d, &w.catr, &w.d5, &w.gx, &w.xch};
    return run_tail(net, tb, spk, B, T, spec, accumulate, st);
}

#include "zs_train.cuh"
#include "zs_dsp.cuh"
#include "critic.cuh"
