// Fused conv1d / per-frame-linear layer as an implicit GEMM on tcgen05 tensor cores.
//
//   D[m = out channel, n = (segment, frame)] = sum_{tap, ci} W[m][tap][ci] * X[segment][row0 + frame*stride + tap][ci]
//
// A (weights, [m_rows][w_taps * c_in_pad], K-major) and B (activations, channels-last with halo rows,
// [B][rows][pitch]) are both fetched by TMA with the 128-byte swizzle; a tap shift is a row offset of the
// B box, a stride-2 conv reads the buffer through a (channel, row parity, row pair, segment) view so the
// box stays dense.
// One CTA per SM, persistent over output tiles of 128 channels x (nb segments x Tt frames <= 256 columns).
// Warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-3 idle (they only exist so that the driver
// warps form a warp group of their own that hands its registers to the epilogue: setmaxnreg 56 / 224), warps 4..11 =
// epilogue in two SETS of four (one warp per TMEM lane quadrant in each set); the sets take alternate segment groups of a tile, each with its
// own staging tiles, residual barrier and TMA stores, so two epilogue warps per scheduler hide each other's
// TMEM-load / shared-memory latency.  Two fp32 accumulators of 256 TMEM columns each let the epilogue of tile i
// overlap the main loop of tile i+1.
//
// Epilogue (one thread = one output channel = one TMEM lane; the frames of a segment are its columns, so
// InstanceNorm statistics never leave the thread): + bias[speaker] -> leaky-relu -> InstanceNorm ->
// + residual (same frame / avg-pool-2 / nearest-up-2) -> sigmoid|tanh -> store (channels-last with
// reflected halo rows for the next conv, pixel-shuffled channels-last, or the reference's (B, C, T) fp32).
//
// Variants of the main loop (template arguments of conv_gemm_kernel; launch_conv in zs_ae.cu picks per layer from measurements, every
// variant accumulates chunk-major with the taps inside so that all of them give bit-identical results):
//   PAIR   clusters of two CTAs share one tcgen05.mma.cta_group::2 (M = 256), each CTA stages half of the tile's columns, 4 stages
//   REUSE  the B stage carries the taps' extra rows and is staged once per 64-channel chunk; tap = a row offset of the MMA's descriptor
//   DEEP   four 48 KB stages and one output tile per epilogue set (pixel-shuffle up-convs on 16 / 32 frames)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace zs {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int MAX_BN = 256;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;
constexpr int EPI_SETS = 2;
constexpr int GEMM_THREADS = 128 + EPI_SETS * 128;   // warp group 0: TMA producer, MMA issuer, two idle warps; warp groups 1, 2: the epilogue sets
constexpr int REGS_DRIVER = 56, REGS_EPILOGUE = 224;   // setmaxnreg split of the 168 x 384 register budget (40 + 2 x 232 = 504 <= 512 per scheduler)
constexpr int TMEM_COLS = 512;
constexpr int STAGING_BYTES = 65536;   // per epilogue set 2 staging tiles of 64 columns x 128 channels (or 128 x 64) x 2 B:
                                       // output double buffer, or output + residual tile
constexpr int SET_STAGING_BYTES = STAGING_BYTES / EPI_SETS;
constexpr int STG_TILE_BYTES = SET_STAGING_BYTES / 2;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/;
// CTA-pair variant (tcgen05 cta_group::2, M = 256): each CTA stages its own 128 weight rows and HALF of the tile's columns,
// 32 KB per stage instead of 48 KB -> four stages in less shared memory, a third less L2 -> shared-memory traffic per FLOP
constexpr int PAIR_STAGES = 4;
constexpr int PAIR_B_STAGE_BYTES = B_STAGE_BYTES / 2;
constexpr int MAX_STAGES = 4;
// ... and the shared memory that frees holds a SECOND residual tile per epilogue set: the residual of round r + 1 is in flight
// while round r is processed (with one tile its TMA load latency is exposed once per round: 17 % of the IN + residual epilogue)
constexpr int PAIR_SET_STAGING_BYTES = SET_STAGING_BYTES + STG_TILE_BYTES;
// Single-CTA layers WITHOUT a residual and with a channels-last output (conv bank, conv2/3/7, the 16- / 32-frame up-convs ...) run
// a deeper ring instead: four 48 KB stages and ONE 16 KB output tile per epilogue set (template DEEP) - their main loop waits on TMA
// latency with three stages in flight (ncu: producer and MMA warp both > 55 % in their barrier waits on `bank`)
constexpr int DEEP_STAGES = 4;
constexpr int DEEP_SET_STAGING_BYTES = STG_TILE_BYTES;
constexpr int DEEP_SMEM_BYTES = DEEP_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_SETS * DEEP_SET_STAGING_BYTES + 1024 + 256;
// Tap-reusing main loop (template REUSE; stride-1 convs with more than one tap whose MMAs stay >= 128 columns wide): the B stage
// holds the frames of a tile WITH the taps' extra rows ({64 channels, Tt + taps - 1 rows} per segment) and every tap's MMA reads
// it through a descriptor that starts `tap` rows further - a K-major 128-byte-swizzled operand may start at any 128-byte row of a
// 1024-byte-aligned tile with base offset 0 (tools/desc_offset_test.cu: the swizzle is a function of the absolute address).  B is
// then staged once per 64-channel chunk instead of once per (tap, chunk): 43 % fewer staged bytes for k = 3, 49 % for the bank.
// The rows of consecutive segments are (taps - 1) apart from contiguous, so the tile takes one MMA per segment (N = Tt; CTA pairs:
// N = 2 Tt over segment 2g of the leader and 2g + 1 of the peer).  A and B travel through separate rings.
constexpr int REUSE_A_STAGES = 5, REUSE_B_STAGES = 2;
constexpr int REUSE_B_BYTES = 35 * 1024;             // 2 segments x (128 + 6) rows x 128 B = 34 304
constexpr int REUSE_PAIR_A_STAGES = 4, REUSE_PAIR_B_STAGES = 3;
constexpr int REUSE_PAIR_B_BYTES = 17 * 1024;        // per CTA: (128 + 2) rows, or 2 segments x (64 + 2) rows
constexpr int REUSE_SMEM_BYTES = REUSE_A_STAGES * A_STAGE_BYTES + REUSE_B_STAGES * REUSE_B_BYTES + STAGING_BYTES + 1024 + 256;
constexpr int REUSE_PAIR_SMEM_BYTES = REUSE_PAIR_A_STAGES * A_STAGE_BYTES + REUSE_PAIR_B_STAGES * REUSE_PAIR_B_BYTES +
                                      EPI_SETS * PAIR_SET_STAGING_BYTES + 1024 + 256;
constexpr int MAX_A_STAGES = 5, MAX_B_STAGES = 3;
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * (A_STAGE_BYTES + PAIR_B_STAGE_BYTES) + EPI_SETS * PAIR_SET_STAGING_BYTES + 1024 + 256;
constexpr float IN_EPS = 1e-5f;

// timing-experiment bits (results are wrong with any bit set) exist only in -DZS_EXPERIMENTS builds
#ifdef ZS_EXPERIMENTS
#define ZS_DBG(p) ((p).debug)
#else
#define ZS_DBG(p) 0
#endif

enum { RES_NONE = 0, RES_SAME = 1, RES_AVG2 = 2, RES_UP2 = 3 };
enum { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_TANH = 2 };
enum { OUT_CL = 0, OUT_PS = 1, OUT_NCT32 = 2 };

struct alignas(64) GemmParams {
    CUtensorMap tmA;
    CUtensorMap tmB;
    CUtensorMap tmOut;   // channels-last output (OUT_CL / OUT_PS): box {128|64 channels, rows, segments}
    CUtensorMap tmRes;   // residual tile (RES_SAME: same box; RES_UP2: half the rows; RES_AVG2: twice the rows)
    // epilogue rounds: segments per round, box rows (frames) per segment, sub-rounds per segment, frames per sub-round
    int rnd_ns, rnd_rows, rnd_sub, rnd_frames;
    int m_tiles, n_tiles, nb, Tt, T, B, N;
    int kc, taps, bank, stride, in_row0, c_in_pad;
    int pair;            // CTA-pair launch (cluster of 2): tmB's box holds nb / 2 segments, idesc has M = 256
    // tap-reusing main loop: tmB's box = {64 channels, reuse_rb rows, all nb segments (single CTA) | 1 segment (pair)}; reuse_g MMAs of
    // reuse_nm columns per K = 16 step (idesc is built for reuse_nm), the B operand of group g and tap j starts (g reuse_rb + j) rows in
    int reuse_rb, reuse_g, reuse_nm;
    int last_mmas;       // K = 16 MMAs of the LAST 64-channel chunk of a tap that hold valid input channels (1..4)
    int pdl;             // launched with programmatic stream serialization: wait for the producer grid before touching its data
    int m_valid;
    const float* bias;
    const long long* spk;
    int bias_stride, n_spk;
    int lrelu;
    float ns;
    int inorm;
    int res_mode;
    const void* res;
    int res_rows, res_pitch, res_halo;
    int act, out_mode;
    void* out;
    int out_rows, out_pitch, out_halo, out_choff, accumulate;
    uint32_t idesc;
    int debug;   // -DZS_EXPERIMENTS builds only (results wrong): 1 = no TMA loads, 2 = no MMA issue, 4 = no epilogue stores/loads
    // training extras
    float* stats;             // InstanceNorm (mean, rstd) per (segment, channel): [B][bias_stride][2], or null
    const float* post_emb;    // added after everything: post_emb[spk[b]][out channel] (speaker embedding of the NEXT layer's input)
    const long long* post_spk;
    int post_pitch, post_n;
    int no_sat;               // gradient outputs: let fp16 overflow to inf (the loss-scale logic detects it) instead of clamping
    int nct_tma;              // OUT_NCT32 only, != 0: tmOut describes the (frames, channels, segments) output; the epilogue stages rows of this many bytes (128 swizzled / 64 plain) in shared memory and TMA-stores them
    int out_f16;              // OUT_NCT32 only: write the (B, C, T) output as fp16 instead of fp32 (halves the D2H bytes of the spectrograms)
    unsigned int* sat_count;  // device word, += 1 per epilogue thread that clamped an fp16 output to +-65504 (never silent)
    // zero-padding mode (model/model.py:36-38, seg_len < 64): halo rows are zeros, and a layer whose speaker embedding is
    // folded into the bias loses the taps that fall into the padding: frame 0 gets -edge_lo, frame T-1 gets -edge_hi
    int zero_halo;
    const float* edge_lo;     // [n_spk][bias_stride] = W_0 . e   (null: reflect mode or no folded embedding)
    const float* edge_hi;     // [n_spk][bias_stride] = W_{k-1} . e
};

template <typename OT>
__device__ __forceinline__ float ot_to_float(OT v);
template <>
__device__ __forceinline__ float ot_to_float<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float ot_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename OT>
__device__ __forceinline__ OT float_to_ot(float v);
template <>
__device__ __forceinline__ __half float_to_ot<__half>(float v) {
    // saturate instead of overflowing to inf (one F2FP.SATFINITE): fp16 operands carry tf32-class mantissa but less
    // range; the GEMM epilogue counts the values it had to clamp (GemmParams::sat_count -> zs_saturation_count)
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return __ushort_as_half(h);
}
template <>
__device__ __forceinline__ __nv_bfloat16 float_to_ot<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// MUFU.RCP (1 ulp): the IEEE-rounded reciprocal compiles to a branchy slow-path call that serialises the gates
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_f(float v) { return rcp_approx(1.f + __expf(-v)); }
__device__ __forceinline__ float tanh_f(float v) {
    float e = __expf(-2.f * fabsf(v));
    float r = (1.f - e) * rcp_approx(1.f + e);
    return copysignf(r, v);
}

// ---- epilogue building blocks ---------------------------------------------------------------------
// One thread = one output channel = one TMEM lane; the frames of a segment are that lane's columns.

struct ChanNorm {      // y = act(lrelu(acc + bias)) * scale + shift (+ residual) + post   (InstanceNorm folded into scale/shift)
    float bias, scale, shift, post;
    int edge_off;             // zero-padding mode: index into the edge tables (subtracted from frame 0 / frame T-1), -1 = none
};
// zero-padding mode: the accumulators of a 16-frame chunk starting at frame c0, corrected at the segment's two edges
// (the two table reads happen twice per segment and pass, so they are not kept in registers)
__device__ __forceinline__ void apply_edges(const GemmParams& p, uint32_t (&v)[16], int c0, int T, const ChanNorm& cn) {
    if (c0 == 0) v[0] = __float_as_uint(__uint_as_float(v[0]) - p.edge_lo[cn.edge_off]);
    if (T - 1 >= c0 && T - 1 < c0 + 16) {
        const float e = p.edge_hi[cn.edge_off];
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i == T - 1) v[i] = __float_as_uint(__uint_as_float(v[i]) - e);
    }
}
template <typename OT> struct IS_BF16 { static constexpr bool value = false; };
template <> struct IS_BF16<__nv_bfloat16> { static constexpr bool value = true; };
template <typename OT>
__device__ __forceinline__ OT float_to_ot_nosat(float v);
template <>
__device__ __forceinline__ __half float_to_ot_nosat<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 float_to_ot_nosat<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-bit operand helpers of the packed epilogue
template <typename OT> __device__ __forceinline__ float res_add(uint16_t r, float x);            // x + float(r), one FHADD
template <> __device__ __forceinline__ float res_add<__half>(uint16_t r, float x) { return fhadd_f16(r, x); }
template <> __device__ __forceinline__ float res_add<__nv_bfloat16>(uint16_t r, float x) { return fhadd_bf16(r, x); }
// two frames -> one packed 16-bit pair (lo = first frame).  fp16 saturates at +-65504 unless NOSAT (gradient outputs)
template <typename OT, bool NOSAT> __device__ __forceinline__ uint32_t cvt_pair(float lo, float hi);
template <> __device__ __forceinline__ uint32_t cvt_pair<__half, false>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t cvt_pair<__half, true>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t cvt_pair<__nv_bfloat16, false>(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <> __device__ __forceinline__ uint32_t cvt_pair<__nv_bfloat16, true>(float lo, float hi) { return cvt_pair<__nv_bfloat16, false>(lo, hi); }
// running max of |fp16 pair| (HMNMX2 with the abs modifier): the fp16 range check costs half an instruction per value
__device__ __forceinline__ uint32_t absmax_h2(uint32_t pair, uint32_t m) {
    const __half2 r = __hmax2(__habs2(*reinterpret_cast<const __half2*>(&pair)), *reinterpret_cast<const __half2*>(&m));
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ bool h2_at_limit(uint32_t m) {       // either half reached the largest finite fp16 (what the clamp produces)
    return (m & 0xffffu) >= 0x7bffu || (m >> 16) >= 0x7bffu;
}

// The epilogue's affine + leaky-relu chain in packed form.  With s = InstanceNorm scale (> 0) and ns the leaky slope,
//   lrelu(a + bias) * s + shift (+ post)  =  max(a * s + t1, a * (ns s) + t2),   t1 = bias s + shift (+ post), t2 = ns bias s + shift (+ post)
// so two FFMA2 and two FMNMX serve two frames (was add, mul, max, fma per frame).
struct PackedAffine {
    uint64_t S1, T1, S2, T2;
};
__device__ __forceinline__ PackedAffine packed_affine(const ChanNorm& cn, bool lrelu, float ns, bool with_post) {
    const float post = with_post ? cn.post : 0.f;
    const float t1 = fmaf(cn.bias, cn.scale, cn.shift) + post;
    const float s2 = lrelu ? ns * cn.scale : cn.scale;
    const float t2 = lrelu ? fmaf(ns * cn.bias, cn.scale, cn.shift) + post : t1;
    PackedAffine pa;
    pa.S1 = pk2(cn.scale, cn.scale);
    pa.T1 = pk2(t1, t1);
    pa.S2 = pk2(s2, s2);
    pa.T2 = pk2(t2, t2);
    return pa;
}

// explicit shared-window accesses of the staging tiles (32-bit shared addresses: [base + immediate] forms, nothing for the
// compiler to re-derive from a generic pointer inside the hot loops)
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)));
}
__device__ __forceinline__ uint16_t lds16(uint32_t addr) {
    uint16_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
    return r;
}

// InstanceNorm statistics of one (segment, channel) in ONE pass over TMEM: sums are taken relative to the
// first frame's value so the variance does not cancel catastrophically.  Packed math (two frames per instruction,
// d = lrelu(a + bias) - x0 = max(a + c1, ns a + c2)), two independent accumulator pairs, two TMEM buffers in ping-pong
// (the next chunk's load is in flight while this one is summed; no register copies).
template <bool ZP>
__device__ __forceinline__ void chan_stats(const GemmParams& p, uint32_t t_seg, int T, float bias, bool lrelu, float ns,
                                           const ChanNorm& cn, float& mean, float& rstd) {
    uint32_t va[16], vb[16];
    tmem_ld16(t_seg, va);
    tmem_ld_wait();
    if (T > 16) tmem_ld16(t_seg + 16, vb);
    if (ZP && cn.edge_off >= 0) apply_edges(p, va, 0, T, cn);
    float x0 = __uint_as_float(va[0]) + bias;
    if (lrelu) x0 = fmaxf(x0, x0 * ns);
    const float nsv = lrelu ? ns : 1.f;
    const float c1 = bias - x0, c2 = fmaf(nsv, bias, -x0);
    const uint64_t C1 = pk2(c1, c1), C2 = pk2(c2, c2), NS = pk2(nsv, nsv);
    uint64_t Sa = pk2(0.f, 0.f), Sb = Sa, Qa = Sa, Qb = Sa;
    auto chunk = [&](uint32_t (&v)[16], int c0) {
        if (c0 + 16 <= T) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const uint64_t a0 = pk2u(v[i], v[i + 1]), a1 = pk2u(v[i + 2], v[i + 3]);
                float p0, p1, q0, q1, p2, p3, q2, q3;
                upk2(fadd2(a0, C1), p0, p1);
                upk2(ffma2(a0, NS, C2), q0, q1);
                upk2(fadd2(a1, C1), p2, p3);
                upk2(ffma2(a1, NS, C2), q2, q3);
                const uint64_t d0 = pk2(fmaxf(p0, q0), fmaxf(p1, q1)), d1 = pk2(fmaxf(p2, q2), fmaxf(p3, q3));
                Sa = fadd2(Sa, d0);
                Qa = ffma2(d0, d0, Qa);
                Sb = fadd2(Sb, d1);
                Qb = ffma2(d1, d1, Qb);
            }
        } else {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float a = __uint_as_float(v[i]);
                const float d = (c0 + i < T) ? fmaxf(a + c1, fmaf(a, nsv, c2)) : 0.f;
                s1 += d;
                s2 = fmaf(d, d, s2);
            }
            Sa = fadd2(Sa, pk2(s1, 0.f));
            Qa = fadd2(Qa, pk2(s2, 0.f));
        }
    };
    for (int c0 = 0; c0 < T; c0 += 32) {          // va holds chunk c0 (waited for, edges applied); vb's load (chunk c0 + 16) is in flight
        chunk(va, c0);
        if (c0 + 16 >= T) break;
        tmem_ld_wait();
        if (c0 + 32 < T) tmem_ld16(t_seg + c0 + 32, va);
        if (ZP && cn.edge_off >= 0) apply_edges(p, vb, c0 + 16, T, cn);
        chunk(vb, c0 + 16);
        if (c0 + 32 >= T) break;
        tmem_ld_wait();
        if (c0 + 48 < T) tmem_ld16(t_seg + c0 + 48, vb);
        if (ZP && cn.edge_off >= 0) apply_edges(p, va, c0 + 32, T, cn);
    }
    float sa, sb, qa, qb;
    upk2(fadd2(Sa, Sb), sa, sb);
    upk2(fadd2(Qa, Qb), qa, qb);
    const float s1 = sa + sb, s2 = qa + qb;
    const float inv_T = 1.f / static_cast<float>(T);
    const float m1 = s1 * inv_T;
    mean = x0 + m1;
    rstd = rsqrtf(fmaxf(s2 * inv_T - m1 * m1, 0.f) + IN_EPS);
}

// Frames [f_lo, f_hi) of one (segment, channel) -> shared-memory staging tile (channels-last rows that a TMA
// store then writes out), plus the reflected halo rows written directly.  RES / PS are compile time.
//   stg_a    : shared address of the staging slot of (frame f_lo, this thread's channel); a frame is ROWB bytes further
//   res_a    : shared address of the residual tile's slot (TMA-loaded) of (first residual row of this sub-round, channel)
//   out_s    : output buffer at (this segment, row 0, this thread's output channel) - for the halo rows only
// TRAIN: the training extras (post-added embedding, unsaturated gradient outputs) exist - compile-time so that the
// inference layers carry no instruction for what they do not use.  LR = false: the layer has no leaky-relu (GRU input
// projections) - one FFMA2 per pair and no max.
template <typename OT, int RES, bool PS, bool ZP, bool TRAIN, bool LR>
__device__ __forceinline__ void frames_to_staging(const GemmParams& p, uint32_t t_seg, int f_lo, int f_hi, int T,
                                                  const ChanNorm& cn, bool lrelu, float ns, uint32_t stg_a,
                                                  uint32_t res_a, OT* __restrict__ out_s, int ps_r,
                                                  bool ch_ok, bool& sat) {
    constexpr int ROWB = 256;                         // bytes per input frame in the staging tile: 128 channels, or 2 rows of 64 (pixel shuffle)
    constexpr bool F16 = !IS_BF16<OT>::value;
    const int T_out = PS ? 2 * T : T;
    const int halo = p.out_halo;
    const PackedAffine pa = packed_affine(cn, LR && lrelu, ns, TRAIN);
    const bool nosat = TRAIN && p.no_sat;
    uint32_t satm = 0;                                // running |max| of the fp16 pairs this thread stored
    auto chunk = [&](uint32_t (&v)[16], int c0) {
        if (ZP && cn.edge_off >= 0) apply_edges(p, v, c0, T, cn);
        const int rel = c0 - f_lo;
        const uint32_t ra = RES == RES_UP2 ? res_a + (rel >> 1) * 256 : (RES == RES_AVG2 ? res_a + rel * 512 : res_a + rel * 256);
        uint32_t yp[8];                                 // 16 outputs as 8 packed 16-bit pairs
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const uint64_t a = pk2u(v[i], v[i + 1]);
            float x0, x1;
            upk2(ffma2(a, pa.S1, pa.T1), x0, x1);
            if (LR) {
                float q0, q1;
                upk2(ffma2(a, pa.S2, pa.T2), q0, q1);
                x0 = fmaxf(x0, q0);
                x1 = fmaxf(x1, q1);
            }
            if (RES == RES_SAME) {
                x0 = res_add<OT>(lds16(ra + i * 256), x0);
                x1 = res_add<OT>(lds16(ra + (i + 1) * 256), x1);
            } else if (RES == RES_UP2) {
                const uint16_t r = lds16(ra + (i >> 1) * 256);
                x0 = res_add<OT>(r, x0);
                x1 = res_add<OT>(r, x1);
            } else if (RES == RES_AVG2) {
                const uint16_t r0 = lds16(ra + i * 512), r1 = lds16(ra + i * 512 + 256), r2 = lds16(ra + i * 512 + 512), r3 = lds16(ra + i * 512 + 768);
                const OT o1 = *reinterpret_cast<const OT*>(&r1), o3 = *reinterpret_cast<const OT*>(&r3);
                x0 += 0.5f * res_add<OT>(r0, ot_to_float<OT>(o1));
                x1 += 0.5f * res_add<OT>(r2, ot_to_float<OT>(o3));
            }
            yp[i >> 1] = nosat ? cvt_pair<OT, true>(x0, x1) : cvt_pair<OT, false>(x0, x1);
        }
        // fp16 range check covers what is STORED: frames below T (padding columns hold garbage by design)
        if (F16 && !nosat) {
            if (c0 + 16 <= T) {
#pragma unroll
                for (int i = 0; i < 8; ++i) satm = absmax_h2(yp[i], satm);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t m = (c0 + 2 * i + 1 < T) ? 0xffffffffu : (c0 + 2 * i < T ? 0xffffu : 0u);
                    satm = absmax_h2(yp[i] & m, satm);
                }
            }
        }
        const uint32_t sp = stg_a + rel * ROWB;
#pragma unroll
        for (int i = 0; i < 8; ++i) {                     // columns >= T are clipped by the TMA store
            sts16(sp + (2 * i) * ROWB, yp[i]);
            sts16(sp + (2 * i + 1) * ROWB, yp[i] >> 16);
        }
        // reflected halo rows of the output buffer (read by the next conv's outer taps): only output rows 1..halo and
        // T_out-1-halo..T_out-2 are mirrored, i.e. (halo <= 3, checked at launch) at most the first and the last four frames - read
        // back from the staging row this thread just wrote instead of selecting among the packed registers
        if (halo > 0 && (c0 == 0 || c0 + 16 + 3 >= T)) {     // warp-uniform condition: ch_ok (per lane) must not guard the __syncwarp below
            uint16_t* out16 = reinterpret_cast<uint16_t*>(out_s);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = j < 4 ? j : T - 1 - (j - 4);        // candidates: frames 0..3, then T-1..T-4 (those >= 4: no frame twice)
                if (t >= c0 && t < c0 + 16 && t < T && (j < 4 || t >= 4) && ch_ok) {
                    const int f = PS ? 2 * t + ps_r : t;
                    const bool top = f >= 1 && f <= halo, bot = f >= T_out - 1 - halo && f <= T_out - 2;
                    if (top || bot) {
                        const uint16_t hv = ZP ? uint16_t(0) : lds16(sp + (t - c0) * ROWB);
                        if (top) out16[(halo - f) * p.out_pitch] = hv;
                        if (bot) out16[(halo + 2 * (T_out - 1) - f) * p.out_pitch] = hv;
                    }
                }
            }
            __syncwarp();
        }
    };
    uint32_t va[16], vb[16];                              // ping-pong: the next chunk's accumulators load while this one is processed
    tmem_ld16(t_seg + f_lo, va);
    for (int c0 = f_lo; c0 < f_hi; c0 += 32) {
        tmem_ld_wait();
        const bool more = c0 + 16 < f_hi;
        if (more) tmem_ld16(t_seg + c0 + 16, vb);
        chunk(va, c0);
        if (!more) break;
        tmem_ld_wait();
        if (c0 + 32 < f_hi) tmem_ld16(t_seg + c0 + 32, va);
        chunk(vb, c0 + 16);
    }
    if (F16 && ch_ok) sat |= h2_at_limit(satm);
}

// Frames of one (segment, channel) -> the reference's (B, C, T) layout, 16 contiguous values per chunk: fp32 (what
// Decoder.forward returns), or fp16 when p.out_f16 (the streaming front-end's byte-saving download format).
template <typename OT, bool ZP>
__device__ __forceinline__ void frames_to_nct(const GemmParams& p, uint32_t t_seg, int T, const ChanNorm& cn,
                                              bool lrelu, float ns, void* __restrict__ nct_v, bool ch_ok) {
    float* nct = reinterpret_cast<float*>(nct_v);
    __half* nct16 = reinterpret_cast<__half*>(nct_v);
    const bool f16 = p.out_f16 != 0;
    const bool vec4 = !f16 && (T & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    const bool vec8 = f16 && (T & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int c0 = 0; c0 < T; c0 += 16) {
        __syncwarp();
        uint32_t v[16];
        tmem_ld16(t_seg + c0, v);
        float r[16];
        if (p.accumulate) {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = (ch_ok && c0 + i < T) ? (f16 ? __half2float(nct16[c0 + i]) : nct[c0 + i]) : 0.f;
        }
        tmem_ld_wait();
        if (ZP && cn.edge_off >= 0) apply_edges(p, v, c0, T, cn);
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float y = __uint_as_float(v[i]) + cn.bias;
            if (lrelu) y = fmaxf(y, y * ns);
            y = fmaf(y, cn.scale, cn.shift);
            if (p.act == ACT_SIGMOID) y = sigmoid_f(y);
            else if (p.act == ACT_TANH) y = tanh_f(y);
            if (p.accumulate == 1) y = r[i] + y;
            else if (p.accumulate == 2) y = fmaf(r[i], y, r[i]);
            x[i] = y;
        }
        if (!ch_ok) continue;
        if (f16) {
            if (vec8 && c0 + 16 <= T) {
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
                    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                *reinterpret_cast<uint4*>(nct16 + c0) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(nct16 + c0 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < T) nct16[c0 + i] = __float2half_rn(x[i]);
            }
        } else if (vec4 && c0 + 16 <= T) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(nct + c0 + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < T) nct[c0 + i] = x[i];
        }
    }
}

// Same result through shared memory: the warp's 32 channels x one 128-byte row of frames (32 fp32 / 64 fp16) are staged with
// the 128-byte swizzle (16-byte chunk j of row r sits at chunk j ^ (r & 7): the 32 lanes' 16-byte stores hit distinct banks) and
// written by a per-warp TMA store, double buffered - no 64-byte strided global stores, no barrier wider than the warp.  Channels past
// m_valid are clipped by the tensor map.  T is a multiple of the row's frames (launch_conv checks).
template <typename OT>
__device__ __forceinline__ void frames_to_nct_tma(const GemmParams& p, uint32_t t_seg, int T, const ChanNorm& cn, bool lrelu, float ns,
                                                  uint8_t* __restrict__ wstage, int lane, int ch0, int b, int& grp) {
    const PackedAffine pa = packed_affine(cn, lrelu, ns, false);
    const bool f16 = p.out_f16 != 0;
    const int rb = p.nct_tma;                            // bytes per staging row: 128 (swizzled) or 64
    const int fpr = f16 ? rb >> 1 : rb >> 2;             // frames per staging row
    const int sw = rb == 128 ? (lane & 7) : 0;
    uint32_t v[16], vn[16];
    tmem_ld16(t_seg, v);
    for (int c0 = 0; c0 < T; c0 += 16) {
        tmem_ld_wait();
        const bool more = c0 + 16 < T;
        if (more) tmem_ld16(t_seg + c0 + 16, vn);
        if (c0 % fpr == 0) {                              // starting a row group: the store that read this buffer two groups ago is done
            if (elect_one()) tma_store_wait_read1();
            __syncwarp();
        }
        uint8_t* rowp = wstage + (grp & 1) * 4096 + lane * rb;
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const uint64_t a = pk2u(v[i], v[i + 1]);
            float x0, x1, q0, q1;
            upk2(ffma2(a, pa.S1, pa.T1), x0, x1);
            upk2(ffma2(a, pa.S2, pa.T2), q0, q1);
            x[i] = fmaxf(x0, q0);
            x[i + 1] = fmaxf(x1, q1);
        }
        if (p.act == ACT_SIGMOID) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = sigmoid_f(x[i]);
        } else if (p.act == ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = tanh_f(x[i]);
        }
        if (f16) {
            const int j0 = (c0 & (fpr - 1)) >> 3;
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
                h[i] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(rowp + (((j0) ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(rowp + (((j0 + 1) ^ sw) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
        } else {
            const int j0 = (c0 & (fpr - 1)) >> 2;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(rowp + (((j0 + q) ^ sw) << 4)) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        }
        if ((c0 + 16) % fpr == 0) {
            fence_proxy_async();                          // staged rows -> visible to the TMA engine
            __syncwarp();
            if (elect_one()) {
                tma_store_3d(&p.tmOut, wstage + (grp & 1) * 4096, c0 + 16 - fpr, ch0, b);
                tma_store_commit();
            }
            ++grp;
        }
        if (more) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = vn[i];
        }
    }
}

// ZP: zero-padding mode (seg_len < 64) - a separate instantiation so the reflect-mode kernel carries none of its code.
// TRAIN: training extras compiled in (InstanceNorm statistics output, post-added embedding, unsaturated outputs).
// PAIR: launched as clusters of two CTAs that share one tcgen05.mma.cta_group::2 (M = 256 = two adjacent 128-channel tiles, the same
// N columns).  Each CTA loads its own weight tile and half of the activation columns; the pair leader (cluster rank 0) collects both
// CTAs' TMA bytes on its `full` barriers, issues the MMAs for both, and its commits are multicast to both CTAs' `empty` / `tfull`
// barriers; each CTA runs the epilogue of its own 128 channels out of its own TMEM; both epilogues hand the accumulator back on the
// leader's `tempty`.
template <typename OT, bool ZP, bool TRAIN, bool PAIR, bool DEEP = false, bool REUSE = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    static_assert(!(PAIR && DEEP), "the pair kernel already runs four stages");
    static_assert(!REUSE || (!ZP && !TRAIN && !DEEP), "the tap-reusing loop exists for the inference kernels");
    // NSTG = stages of the A ring (= of the joint A + B stages without REUSE), NSTG_B = stages of REUSE's separate B ring
    constexpr int NSTG = REUSE ? (PAIR ? REUSE_PAIR_A_STAGES : REUSE_A_STAGES) : (PAIR ? PAIR_STAGES : (DEEP ? DEEP_STAGES : STAGES));
    constexpr int NSTG_B = REUSE ? (PAIR ? REUSE_PAIR_B_STAGES : REUSE_B_STAGES) : NSTG;
    constexpr int B_BYTES = REUSE ? (PAIR ? REUSE_PAIR_B_BYTES : REUSE_B_BYTES) : (PAIR ? PAIR_B_STAGE_BYTES : B_STAGE_BYTES);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + NSTG * A_STAGE_BYTES;
    constexpr int SET_BYTES = PAIR ? PAIR_SET_STAGING_BYTES : (DEEP ? DEEP_SET_STAGING_BYTES : SET_STAGING_BYTES);
    uint8_t* sStage = smem + NSTG * A_STAGE_BYTES + NSTG_B * B_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sStage + EPI_SETS * SET_BYTES);      // A ring (A + B stages without REUSE)
    uint64_t* empty = full + MAX_A_STAGES;
    uint64_t* fullB = empty + MAX_A_STAGES;                                           // REUSE: the B ring
    uint64_t* emptyB = fullB + MAX_B_STAGES;
    uint64_t* tfull = emptyB + MAX_B_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* rbar = tempty + 2;   // [EPI_SETS][2] residual tile (slot 0 / 1) landed in the set's staging buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 2 * EPI_SETS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // 0 = the pair's leader

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        if (REUSE)
            for (int i = 0; i < NSTG_B; ++i) {
                mbar_init(&fullB[i], 1);
                mbar_init(&emptyB[i], 1);
            }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], (PAIR ? 2 : 1) * 4 * EPI_SETS);
        }
        for (int i = 0; i < 2 * EPI_SETS; ++i) mbar_init(&rbar[i], 1);
        fence_barrier_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        if (p.out_mode != OUT_NCT32 || p.nct_tma) tma_prefetch_desc(&p.tmOut);
        if (p.res_mode != RES_NONE) tma_prefetch_desc(&p.tmRes);
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
        else tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the peer's barriers are initialised and its TMEM allocated before anything is signalled across
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (p.pdl) {
        // programmatic dependent launch: the next kernel of the stream may be scheduled as SMs free up (its prologue
        // then overlaps this grid's tail); everything above touched no data of the producer grid - from here on it does
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    // a pair walks (m-tile pair, n-tile) units: CTA `rank` takes m-tile 2 mu + rank
    const int m_units = PAIR ? p.m_tiles >> 1 : p.m_tiles;
    const int total_tiles = m_units * p.n_tiles;
    const int tile0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, tile_step = PAIR ? gridDim.x >> 1 : gridDim.x;
    const uint32_t stage_tx = PAIR ? 2u * (A_STAGE_BYTES + static_cast<uint32_t>(p.N >> 1) * (BK * 2))
                                   : A_STAGE_BYTES + static_cast<uint32_t>(p.N) * (BK * 2);

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_DRIVER));
    if (warp == 0) {
        // ------------------------------ TMA producer (whole warp, one elected lane issues) ------
        {
            int stage = 0, stage_b = 0;
            uint32_t phase = 0, phase_b = 0;
            const int seg_off = PAIR ? static_cast<int>(rank) * (p.nb >> 1) : 0;      // this CTA's half of the tile's segments
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int mu = tile % m_units, nt = tile / m_units;
                const int mt = PAIR ? 2 * mu + static_cast<int>(rank) : mu;
                int tap_lo = 0, ntaps = p.taps;
                if (!PAIR && p.bank) {
                    ntaps = mt + 1;
                    tap_lo = 3 - ntaps / 2;
                }
                // chunk-major, taps inside (every variant of the kernel accumulates in this order: results do not depend on
                // which variant ran a layer)
                for (int c = 0; c < p.kc; ++c) {
                    if (REUSE) {
                        // ---- B(c): the tile's frames with the taps' extra rows, once per chunk ----
                        mbar_wait(&emptyB[stage_b], phase_b ^ 1);
                        if (elect_one()) {
                            uint8_t* dB = sB + stage_b * B_BYTES;
                            if (PAIR) {
                                if (rank == 0) mbar_expect_tx(&fullB[stage_b], 2u * p.reuse_g * p.reuse_rb * 128u);
                                const uint32_t fb = mapa_shared(smem_u32(&fullB[stage_b]), 0);
                                for (int g = 0; g < p.reuse_g; ++g)      // segment 2g (leader) / 2g + 1 (peer): MMA g spans both
                                    tma_load_3d_pair(&p.tmB, dB + g * p.reuse_rb * 128, fb, c * BK, p.in_row0, nt * p.nb + 2 * g + static_cast<int>(rank));
                            } else {
                                mbar_expect_tx(&fullB[stage_b], static_cast<uint32_t>(p.nb) * p.reuse_rb * 128u);
                                tma_load_3d(&p.tmB, dB, &fullB[stage_b], c * BK, p.in_row0, nt * p.nb);
                            }
                        }
                        __syncwarp();
                        if (++stage_b == NSTG_B) {
                            stage_b = 0;
                            phase_b ^= 1;
                        }
                    }
                    for (int j = 0; j < ntaps; ++j) {
                        const int tap = tap_lo + j;
                        const int row_b = p.in_row0 + tap;
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (ZS_DBG(p) & 1) {
                            if (elect_one()) mbar_arrive(&full[stage]);
                        } else if (elect_one()) {
                            uint8_t* dA = sA + stage * A_STAGE_BYTES;
                            uint8_t* dB = sB + stage * B_BYTES;
                            if (REUSE) {
                                if (PAIR) {
                                    if (rank == 0) mbar_expect_tx(&full[stage], 2u * A_STAGE_BYTES);
                                    tma_load_2d_pair(&p.tmA, dA, mapa_shared(smem_u32(&full[stage]), 0), tap * p.c_in_pad + c * BK, mt * BM);
                                } else {
                                    mbar_expect_tx(&full[stage], A_STAGE_BYTES);
                                    tma_load_2d(&p.tmA, dA, &full[stage], tap * p.c_in_pad + c * BK, mt * BM);
                                }
                            } else if (PAIR) {
                                // both CTAs' bytes complete on the LEADER's barrier; only the leader arms it
                                if (rank == 0) mbar_expect_tx(&full[stage], stage_tx);
                                const uint32_t fb = mapa_shared(smem_u32(&full[stage]), 0);
                                tma_load_2d_pair(&p.tmA, dA, fb, tap * p.c_in_pad + c * BK, mt * BM);
                                if (p.stride == 2) tma_load_4d_pair(&p.tmB, dB, fb, c * BK, row_b & 1, row_b >> 1, nt * p.nb + seg_off);
                                else tma_load_3d_pair(&p.tmB, dB, fb, c * BK, row_b, nt * p.nb + seg_off);
                            } else {
                                mbar_expect_tx(&full[stage], stage_tx);
                                tma_load_2d(&p.tmA, dA, &full[stage], tap * p.c_in_pad + c * BK, mt * BM);
                                if (p.stride == 2)  // buffer viewed as (channel, row parity, row pair, segment)
                                    tma_load_4d(&p.tmB, dB, &full[stage], c * BK, row_b & 1, row_b >> 1, nt * p.nb);
                                else
                                    tma_load_3d(&p.tmB, dB, &full[stage], c * BK, row_b, nt * p.nb);
                            }
                        }
                        __syncwarp();
                        if (++stage == NSTG) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ------------------------------ MMA issuer (whole warp, one elected lane issues; a pair's leader issues for both CTAs) --------
        {
            int stage = 0, stage_b = 0;
            uint32_t phase = 0, phase_b = 0;
            int it = 0;
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const int mt = tile % m_units;          // (bank mode is never paired: mt is the m-tile there)
                int tap_lo = 0, ntaps = p.taps;
                if (!PAIR && p.bank) {
                    ntaps = mt + 1;
                    tap_lo = 3 - ntaps / 2;
                }
                const int as = it & 1;
                mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * MAX_BN;
                for (int c = 0; c < p.kc; ++c) {
                    // the zero-padded tail of the input channels (513 -> 576, 1409 -> 1472) needs no MMAs
                    const int n_mma = (c + 1 == p.kc) ? p.last_mmas : BK / 16;
                    uint64_t db_c = 0;
                    if (REUSE) {
                        mbar_wait(&fullB[stage_b], phase_b);
                        db_c = umma_desc_sw128(b0 + stage_b * B_BYTES);
                    }
                    for (int j = 0; j < ntaps; ++j) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint64_t da = umma_desc_sw128(a0 + stage * A_STAGE_BYTES);
                        const uint32_t first = (c | j) == 0 ? 0u : 1u;      // 0: the tile's first MMAs overwrite the accumulator
                        if (elect_one()) {
                            if (!(ZS_DBG(p) & 2)) {
                                if (REUSE) {
                                    // one MMA per segment (pair: per segment pair); tap = a row offset of the operand: +8 per row in the >>4 field
                                    for (int g = 0; g < p.reuse_g; ++g) {
                                        const uint64_t db = db_c + static_cast<uint64_t>((g * p.reuse_rb + tap_lo + j) * 8);
                                        const uint32_t dg = d_tmem + g * p.reuse_nm;
#pragma unroll
                                        for (int k = 0; k < BK / 16; ++k) {
                                            if (k < n_mma) {
                                                if (PAIR) umma_f16_pair(dg, da + 2 * k, db + 2 * k, p.idesc, first | static_cast<uint32_t>(k != 0));
                                                else umma_f16(dg, da + 2 * k, db + 2 * k, p.idesc, first | static_cast<uint32_t>(k != 0));
                                            }
                                        }
                                    }
                                } else {
                                    const uint64_t db = umma_desc_sw128(b0 + stage * B_BYTES);
#pragma unroll
                                    for (int k = 0; k < BK / 16; ++k) {
                                        // advance 16 elements (32 B) along K inside the swizzle row: +2 in the >>4 address field
                                        if (k < n_mma) {
                                            if (PAIR) umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, p.idesc, first | static_cast<uint32_t>(k != 0));
                                            else umma_f16(d_tmem, da + 2 * k, db + 2 * k, p.idesc, first | static_cast<uint32_t>(k != 0));
                                        }
                                    }
                                }
                            }
                            if (PAIR) umma_commit_pair(&empty[stage]);      // frees this stage in BOTH CTAs
                            else umma_commit(&empty[stage]);
                        }
                        __syncwarp();
                        if (++stage == NSTG) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (REUSE) {
                        if (elect_one()) {
                            if (PAIR) umma_commit_pair(&emptyB[stage_b]);
                            else umma_commit(&emptyB[stage_b]);
                        }
                        __syncwarp();
                        if (++stage_b == NSTG_B) {
                            stage_b = 0;
                            phase_b ^= 1;
                        }
                    }
                }
                if (elect_one()) {
                    if (PAIR) umma_commit_pair(&tfull[as]);
                    else umma_commit(&tfull[as]);
                }
                __syncwarp();
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPILOGUE));
        // ------------------------------ epilogue ----------------------------------
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        const int eset = (warp - 4) >> 2;                    // epilogue set: warps 4..7 / 8..11
        const bool set_lead = (warp & 3) == 0;               // the set's TMA-issuing warp
        const int set_bar = 1 + eset;                        // named barrier of the set's 128 threads
        uint64_t* my_rbar = &rbar[2 * eset];                  // [2]: one per residual slot (single-CTA kernels use slot 0 only)
        const int row = quad * 32 + lane;
        const bool lrelu = p.lrelu != 0;
        const float ns = p.ns;
        const int T = p.T;
        uint8_t* set_stage = sStage + eset * SET_BYTES;
        OT* stage = reinterpret_cast<OT*>(set_stage);
        int it = 0, rnd = 0;
        uint32_t res_phase = 0;     // bit k = phase of residual slot k's barrier
        int rres = 0;               // residual rounds this set has consumed (PAIR: slot = rres & 1)
        bool sat = false;           // this thread clamped an fp16 output (reported once per thread and launch)
        // both epilogues of a pair hand the accumulator back on the LEADER's barrier
        const uint32_t tempty_leader = PAIR ? mapa_shared(smem_u32(&tempty[0]), 0) : 0u;
        auto release_acc = [&](int as_) {
            if (lane == 0) {
                if (PAIR && rank != 0) mbar_arrive_cluster(tempty_leader + 8u * as_);
                else mbar_arrive(&tempty[as_]);
            }
        };
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
            const int mu = tile % m_units, nt = tile / m_units;
            const int mt = PAIR ? 2 * mu + static_cast<int>(rank) : mu;
            const int as = it & 1;
            const int ch = mt * BM + row;  // weight row == bias index
            const bool ch_ok = ch < p.m_valid;
            // the (speaker id -> bias table row -> bias) chain of the first segment this set touches is two dependent global
            // loads: issue them now, in the shadow of the wait for the tile's MMAs
            const int b_pf = nt * p.nb + eset * (p.out_mode == OUT_NCT32 ? 1 : p.rnd_ns);
            float bias_pf = 0.f;
            size_t off_pf = 0;
            if (p.bias != nullptr && b_pf < p.B) {
                if (p.spk) {
                    long long sp = p.spk[b_pf];
                    sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
                    off_pf = static_cast<size_t>(sp) * p.bias_stride;
                }
                bias_pf = p.bias[off_pf + ch];
            }
            mbar_wait_relaxed(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * MAX_BN;
            auto chan_norm = [&](int b, uint32_t t_seg) {
                ChanNorm cn;
                cn.bias = 0.f;
                cn.edge_off = -1;
                if (p.bias != nullptr) {  // tables are padded to m_tiles * 128 rows
                    size_t off = off_pf;
                    if (b == b_pf) {
                        cn.bias = bias_pf;
                    } else {
                        off = 0;
                        if (p.spk) {
                            long long sp = p.spk[b];
                            sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
                            off = static_cast<size_t>(sp) * p.bias_stride;
                        }
                        cn.bias = p.bias[off + ch];
                    }
                    if (ZP && p.edge_lo != nullptr) cn.edge_off = static_cast<int>(off) + ch;
                }
                cn.scale = 1.f;
                cn.shift = 0.f;
                cn.post = 0.f;
                if (p.inorm) {
                    float mean, rstd;
                    chan_stats<ZP>(p, t_seg, T, cn.bias, lrelu, ns, cn, mean, rstd);
                    cn.scale = rstd;
                    cn.shift = -mean * rstd;
                    if (TRAIN && p.stats != nullptr && ch_ok)
                        *reinterpret_cast<float2*>(p.stats + (static_cast<size_t>(b) * p.bias_stride + ch) * 2) = make_float2(mean, rstd);
                }
                if (TRAIN && p.post_emb != nullptr) {
                    long long sp = p.post_spk[b];
                    sp = sp < 0 ? 0 : (sp >= p.post_n ? p.post_n - 1 : sp);
                    const int oc = p.out_mode == OUT_PS ? (mt * 64 + (row & 63)) : ch;
                    if (oc < p.post_pitch) cn.post = p.post_emb[static_cast<size_t>(sp) * p.post_pitch + oc];
                }
                return cn;
            };
            if (ZS_DBG(p) & 4) {            // timing experiment: main loop only
                tc_fence_before();
                __syncwarp();
                release_acc(as);
            } else if (p.out_mode == OUT_NCT32) {
                for (int s = eset; s < p.nb; s += EPI_SETS) {
                    const int b = nt * p.nb + s;
                    if (b >= p.B) break;
                    const uint32_t t_seg = t_lane + s * p.Tt;
                    const ChanNorm cn = chan_norm(b, t_seg);
                    if (!ZP && p.nct_tma) {
                        frames_to_nct_tma<OT>(p, t_seg, T, cn, lrelu, ns, set_stage + quad * 8192, lane, mt * BM + quad * 32, b, rnd);
                        continue;
                    }
                    const size_t el = (static_cast<size_t>(b) * p.m_valid + ch) * T;
                    frames_to_nct<OT, ZP>(p, t_seg, T, cn, lrelu, ns,
                                          p.out_f16 ? static_cast<void*>(reinterpret_cast<__half*>(p.out) + el) : static_cast<void*>(reinterpret_cast<float*>(p.out) + el), ch_ok);
                }
                tc_fence_before();
                __syncwarp();
                release_acc(as);
            } else {
                // pixel shuffle: weight rows are packed so rows [0,64) of a tile hold r = 0, [64,128) r = 1
                const bool ps = p.out_mode == OUT_PS;
                const bool has_res = p.res_mode != RES_NONE;
                const int ps_r = row >> 6;
                const int out_ch = ps ? (mt * 64 + (row & 63)) : ch;
                const int fstep = ps ? 2 : 1;
                const int stg_row = ps ? 64 : 128;                    // elements per staging row
                const int stg_ch = ps ? (ps_r * 64 + (row & 63)) : row;
                // residual rows per output frame: SAME 1, UP2 1/2, AVG2 2
                const int res_rows_per_seg = p.res_mode == RES_UP2 ? p.rnd_rows / 2 : (p.res_mode == RES_AVG2 ? 2 * p.rnd_rows : p.rnd_rows);
                ChanNorm cn_keep;
                // segment groups alternate between the two sets; the last group of THIS set hands the accumulator back
                const int n_groups = (p.nb + p.rnd_ns - 1) / p.rnd_ns;
                const int my_last = ((n_groups - 1 - eset) & ~1) + eset;      // last group index with my parity (< 0: none)
                if (my_last < 0 || eset >= n_groups) {
                    tc_fence_before();
                    __syncwarp();
                    release_acc(as);
                }
                for (int s0 = eset * p.rnd_ns; s0 < p.nb; s0 += EPI_SETS * p.rnd_ns) {
                    for (int h = 0; h < p.rnd_sub; ++h, ++rnd) {
                        const int f_lo = h * p.rnd_frames, f_hi = min(T, f_lo + p.rnd_frames);
                        // staging: with a residual, tile 0 = output and tile 1 = residual; otherwise the two tiles
                        // alternate as output buffers so a store's shared-memory read overlaps the next round
                        OT* stage_out = stage + ((has_res || DEEP) ? 0 : (rnd & 1) * (STG_TILE_BYTES / 2));
                        const int rslot = PAIR ? (rres & 1) : 0;
                        const OT* stage_res = stage + (1 + rslot) * (STG_TILE_BYTES / 2);
                        const bool live = nt * p.nb + s0 < p.B;
                        // the residual tile of round (s0_, h_) -> slot; the slot's previous readers passed an end-of-round barrier
                        auto issue_res = [&](int s0_, int h_, int slot) {
                            mbar_expect_tx(&my_rbar[slot], static_cast<uint32_t>(p.rnd_ns) * res_rows_per_seg * 256);
                            const int fl = h_ * p.rnd_frames;
                            const int r_row = p.res_halo + (p.res_mode == RES_UP2 ? fl / 2 : (p.res_mode == RES_AVG2 ? 2 * fl : fl));
                            tma_load_3d(&p.tmRes, set_stage + (1 + slot) * STG_TILE_BYTES, &my_rbar[slot], mt * BM, r_row, nt * p.nb + s0_);
                        };
                        if (has_res) {
                            // single-CTA kernels: one slot, loaded at the start of its round.  PAIR: the first round of a tile loads its own
                            // tile here; every later round's tile was requested one round ahead (below)
                            const bool first = s0 == eset * p.rnd_ns && h == 0;
                            if ((!PAIR || first) && set_lead && live && elect_one()) issue_res(s0, h, rslot);
                            __syncwarp();
                        }
                        bool waited = false;
                        for (int s = s0; s < min(s0 + p.rnd_ns, p.nb); ++s) {
                            const int b = nt * p.nb + s;
                            if (b >= p.B) break;
                            const uint32_t t_seg = t_lane + s * p.Tt;
                            if (h == 0) cn_keep = chan_norm(b, t_seg);
                            if (has_res && !waited) {
                                mbar_wait(&my_rbar[rslot], (res_phase >> rslot) & 1);
                                waited = true;
                                if (PAIR) {     // request the NEXT round's residual into the other slot while this round is processed
                                    int s0n = s0, hn = h + 1;
                                    if (hn == p.rnd_sub) {
                                        hn = 0;
                                        s0n = s0 + EPI_SETS * p.rnd_ns;
                                    }
                                    if (s0n < p.nb && nt * p.nb + s0n < p.B && set_lead && elect_one()) issue_res(s0n, hn, rslot ^ 1);
                                    __syncwarp();
                                }
                            }
                            const uint32_t stg = smem_u32(stage_out + static_cast<size_t>(s - s0) * p.rnd_rows * fstep * stg_row + stg_ch);
                            const uint32_t res_stg = smem_u32(stage_res + static_cast<size_t>(s - s0) * res_rows_per_seg * 128 + row);
                            OT* out_s = reinterpret_cast<OT*>(p.out) + static_cast<size_t>(b) * p.out_rows * p.out_pitch +
                                        p.out_choff + out_ch;
#define ZS_F2S(RES_, PS_, LR_) frames_to_staging<OT, RES_, PS_, ZP, TRAIN, LR_>(p, t_seg, f_lo, f_hi, T, cn_keep, lrelu, ns, stg, res_stg, out_s, PS_ ? ps_r : 0, ch_ok, sat)
                            if (ps) ZS_F2S(RES_NONE, true, true);
                            else if (p.res_mode == RES_NONE) {
                                if (lrelu) ZS_F2S(RES_NONE, false, true);
                                else ZS_F2S(RES_NONE, false, false);
                            } else if (p.res_mode == RES_SAME) ZS_F2S(RES_SAME, false, true);
                            else if (p.res_mode == RES_UP2) ZS_F2S(RES_UP2, false, true);
                            else ZS_F2S(RES_AVG2, false, true);
#undef ZS_F2S
                        }
                        if (has_res && live) {
                            res_phase ^= 1u << rslot;
                            ++rres;
                        }
                        const bool last = (s0 / p.rnd_ns == my_last) && (h + 1 == p.rnd_sub);
                        if (last) {   // every TMEM read of this tile is done: hand the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            release_acc(as);
                        }
                        fence_proxy_async();                              // staging writes -> visible to the TMA engine
                        asm volatile("bar.sync %0, 128;" ::"r"(set_bar) : "memory");
                        if (set_lead && live && elect_one()) {
                            tma_store_3d(&p.tmOut, stage_out, ps ? mt * 64 : mt * BM, p.out_halo + fstep * f_lo, nt * p.nb + s0);
                            tma_store_commit();
                            if (has_res || DEEP) tma_store_wait_read();    // single output tile: it must be free next round
                            else tma_store_wait_read1();                   // the other tile's store (2 rounds ago) is done
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(set_bar) : "memory");
                    }
                }
            }
        }
        if (sat && !p.no_sat && p.sat_count != nullptr) atomicAdd(p.sat_count, 1u);
        if (elect_one()) tma_store_wait_read();        // no bulk store may still be reading the staging tiles when the CTA retires
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();      // the leader's MMAs read the peer's shared memory, the peer arrives on the leader's barriers: leave together
    if (warp == 1) {
        if (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
        else tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace zs
