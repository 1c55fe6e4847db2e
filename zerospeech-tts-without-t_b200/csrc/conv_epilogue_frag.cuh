// Fragment-layout epilogue of conv_gemm_kernel (inference path).
//
// The lane-per-thread epilogue (conv_gemm.cuh: one thread = one channel, tcgen05.ld.32x32b) has to transpose the
// accumulator tile into the channels-last output through 2-byte shared-memory stores - one STS.U16 per element - and
// that, not the MMA main loop, bounds the layers with K <= 1024 (ncu r01: tensor pipe 43 % on the InstanceNorm + residual
// dense layers, 36 % on the fp32 output layer).  Here the accumulators are read with tcgen05.ld.16x256b: a thread gets, for
// 8-column groups, the element pairs (row = lane/4 [+8], columns 2*(lane%4), +1) - the mma C-fragment layout - so that
//   * two frames of one channel pack into one b16x2 register (one F2FP per two elements),
//   * stmatrix.x4.trans writes 32 channels x 8 frames per instruction into the [frame][channel] staging tile
//     (the TMA store's 128-byte swizzle keeps the eight 16-byte rows of a matrix on different banks),
//   * a same-frame residual tile comes back in the same fragment through ldmatrix.x4.trans,
//   * the fp32 (B, C, T) output mode stores float2 (four threads = 32 contiguous bytes of one channel row).
// InstanceNorm statistics: each thread accumulates its 2-of-8 columns of its four channels, the four threads of a quad
// combine with two shuffles.  Layout facts verified on hardware by tools/frag_probe.cu.
// (included by conv_gemm.cuh between its epilogue building blocks and the kernel)
#pragma once

namespace zs {

// 16 lanes x 32 columns: reg 4g + 2rh + e = (lane base + lane/4 + 8 rh, column 8g + 2 (lane%4) + e)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 16 columns (the tail of a segment whose padded length is 16 mod 32): regs 0..7 as above
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// one lane half (16 TMEM lanes) of this warp's quadrant, `ng` (2 or 4) column groups of 8 starting at t_addr
__device__ __forceinline__ void frag_load(uint32_t t_addr, int ng, uint32_t (&v)[16]) {
    if (ng >= 4) tmem_ld_16x256b_x4(t_addr, v);
    else tmem_ld_16x256b_x2(t_addr, v);
    tmem_ld_wait();
}
__device__ __forceinline__ void stmatrix_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
    asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}

template <typename OT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi, bool nosat);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi, bool nosat) {
    uint32_t r;
    if (nosat) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi, bool) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <typename OT>
__device__ __forceinline__ float2 unpack2(uint32_t v);
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

// byte offset of (row, 16-byte chunk) inside a [rows][64 channels] b16 tile written / read by TMA with SWIZZLE_128B
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) { return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4)); }

struct FragCh {            // the four channels of this thread: slot = 2 * lane-half + row-half -> tile row 32 quad + 8 slot + lane / 4
    float bias[4], scale[4], shift[4];
    bool ok[4];
};

// Per-(segment) parameters of the thread's four channels: bias (+ folded speaker term), then InstanceNorm scale / shift
// from one pass over the segment's accumulator columns (one lane half at a time: 16 live accumulator registers).
template <typename OT>
__device__ __forceinline__ void frag_chan_norm(const GemmParams& p, FragCh& fc, uint32_t t_q, int T, int Tt, int b, int ch0, int lane,
                                               bool lrelu, float ns) {
    size_t off = 0;
    if (p.spk) {
        long long sp = p.spk[b];
        sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
        off = static_cast<size_t>(sp) * p.bias_stride;
    }
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
        const int ch = ch0 + 8 * sl;
        fc.ok[sl] = ch < p.m_valid;
        fc.bias[sl] = p.bias != nullptr ? p.bias[off + ch] : 0.f;          // tables are padded to m_tiles * 128 rows
        fc.scale[sl] = 1.f;
        fc.shift[sl] = 0.f;
    }
    if (!p.inorm) return;
    const int fcol = 2 * (lane & 3);
    const float inv_T = 1.f / static_cast<float>(T);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f}, x0[2] = {0.f, 0.f};
        for (int c0 = 0; c0 < Tt; c0 += 32) {
            uint32_t v[16];
            const int ng = min(4, (Tt - c0) >> 3);         // Tt is a multiple of 16: the last chunk may hold two groups only
            frag_load(t_q + (static_cast<uint32_t>(16 * hf) << 16) + c0, ng, v);
            if (c0 == 0) {    // shift the sums by the channel's first value (held by the quad's first thread): no cancellation
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    float x = __uint_as_float(v[2 * rh]) + fc.bias[2 * hf + rh];
                    if (lrelu) x = fmaxf(x, x * ns);
                    x0[rh] = __shfl_sync(0xffffffffu, x, lane & ~3);
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g >= ng) break;
                const int f = c0 + 8 * g + fcol;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float x = __uint_as_float(v[4 * g + 2 * rh + e]) + fc.bias[2 * hf + rh];
                        if (lrelu) x = fmaxf(x, x * ns);
                        const float d = (f + e < T) ? x - x0[rh] : 0.f;
                        s1[rh] += d;
                        s2[rh] = fmaf(d, d, s2[rh]);
                    }
                }
            }
        }
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            float a = s1[rh], q = s2[rh];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            q += __shfl_xor_sync(0xffffffffu, q, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            q += __shfl_xor_sync(0xffffffffu, q, 2);
            const float m1 = a * inv_T;
            const float mean = x0[rh] + m1;
            const float rstd = rsqrtf(fmaxf(q * inv_T - m1 * m1, 0.f) + IN_EPS);
            fc.scale[2 * hf + rh] = rstd;
            fc.shift[2 * hf + rh] = -mean * rstd;
        }
    }
}

// Frames [f_lo, f_hi) of one segment -> the set's swizzled staging tile (rows = frames, 64-channel halves), plus the
// reflected halo rows written directly.  seg_row0 = first staging row of this segment in the round.
//   RES_SAME : residual tile (same geometry, TMA-loaded) read through ldmatrix.trans
//   RES_UP2  : residual tile holds half the rows; out frames (2m, 2m+1) share row m
//   RES_AVG2 : residual tile holds twice the rows; out frame f averages rows 2f, 2f+1
template <typename OT, int RES, bool PS>
__device__ __forceinline__ void frag_frames_to_staging(const GemmParams& p, const FragCh& fc, uint32_t t_q, int f_lo, int f_hi, int T, int Tt,
                                                       bool lrelu, float ns, uint32_t stg_base, uint32_t res_base, int seg_row0,
                                                       int res_row0, OT* __restrict__ out_s, int quad, int lane, bool& sat) {
    const int fcol = 2 * (lane & 3);
    const int mi = (lane >> 3) & 1, mj = lane & 7;               // stmatrix / ldmatrix .x2: thread 8 i + j addresses row j of matrix i
    const uint32_t half_off = PS ? 0u : static_cast<uint32_t>(quad >> 1) * 8192u;
    const char* res_gen = reinterpret_cast<const char*>(__cvta_shared_to_generic(res_base + half_off));
    const int ps_r = quad >> 1;
    const int T_out = PS ? 2 * T : T;
    const int halo = p.out_halo;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int chunk0 = 4 * (quad & 1) + 2 * hf;               // 16-byte chunk (8 channels) of slot 2 hf inside its 64-channel half
        for (int c0 = f_lo; c0 < f_hi; c0 += 32) {
            uint32_t v[16];
            const int ng = min(4, (Tt - c0) >> 3);
            frag_load(t_q + (static_cast<uint32_t>(16 * hf) << 16) + c0, ng, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g >= ng || c0 + 8 * g >= f_hi) break;
                const int fr = c0 + 8 * g - f_lo;                // first frame of the group, relative to the round
                uint32_t rr[2];
                if (RES == RES_SAME) ldmatrix_x2_trans(res_base + half_off + swz_off(res_row0 + fr + mj, chunk0 + mi), rr);
                uint32_t pk[2];
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int sl = 2 * hf + rh;
                    float2 r2 = make_float2(0.f, 0.f);
                    if (RES == RES_SAME) r2 = unpack2<OT>(rr[rh]);
                    else if (RES == RES_UP2) {
                        const int row = res_row0 + ((fr + fcol) >> 1);
                        r2.x = r2.y = ot_to_float<OT>(*reinterpret_cast<const OT*>(res_gen + swz_off(row, chunk0 + rh) + 2 * (lane >> 2)));
                    } else if (RES == RES_AVG2) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int row = res_row0 + 2 * (fr + fcol + e);
                            const float a0 = ot_to_float<OT>(*reinterpret_cast<const OT*>(res_gen + swz_off(row, chunk0 + rh) + 2 * (lane >> 2)));
                            const float a1 = ot_to_float<OT>(*reinterpret_cast<const OT*>(res_gen + swz_off(row + 1, chunk0 + rh) + 2 * (lane >> 2)));
                            (e ? r2.y : r2.x) = 0.5f * (a0 + a1);
                        }
                    }
                    float y[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float x = __uint_as_float(v[4 * g + 2 * rh + e]) + fc.bias[sl];
                        if (lrelu) x = fmaxf(x, x * ns);
                        if (p.inorm) x = fmaf(x, fc.scale[sl], fc.shift[sl]);
                        if (RES != RES_NONE) x += e ? r2.y : r2.x;
                        if (!IS_BF16<OT>::value) sat |= fc.ok[sl] && (c0 + 8 * g + fcol + e < T) && fabsf(x) > 65504.f;
                        y[e] = x;
                    }
                    pk[rh] = pack2<OT>(y[0], y[1], p.no_sat != 0);
                }
                // memory row of frame fr + mj: PS interleaves the two pixel-shuffle phases (out frame 2 t + r)
                const int srow = PS ? seg_row0 + 2 * (fr + mj) + ps_r : seg_row0 + fr + mj;
                stmatrix_x2_trans(stg_base + half_off + swz_off(srow, chunk0 + mi), pk[0], pk[1]);
                // reflected halo rows of the output buffer (read by the next conv's outer taps), straight to global memory
                if (halo > 0 && (c0 + 8 * g <= halo || c0 + 8 * g + 8 + halo + 1 >= T)) {
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) {
                        const int sl = 2 * hf + rh;
                        if (!fc.ok[sl]) continue;
                        const int oc = (PS ? 32 * (quad & 1) : 0) + 8 * sl + (lane >> 2);      // channel offset inside out_s
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int t = c0 + 8 * g + fcol + e;
                            if (t >= T) continue;
                            const int f = PS ? 2 * t + ps_r : t;
                            const unsigned short bits = static_cast<unsigned short>(e ? (pk[rh] >> 16) : (pk[rh] & 0xffffu));
                            const OT hv = *reinterpret_cast<const OT*>(&bits);
                            if (f >= 1 && f <= halo) out_s[static_cast<size_t>(halo - f) * p.out_pitch + oc] = hv;
                            if (f >= T_out - 1 - halo && f <= T_out - 2) out_s[static_cast<size_t>(halo + 2 * (T_out - 1) - f) * p.out_pitch + oc] = hv;
                        }
                    }
                }
            }
        }
    }
}

// Frames of one segment -> the reference's (B, C, T) fp32 layout: float2 per (channel, frame pair)
template <typename OT>
__device__ __forceinline__ void frag_frames_to_nct(const GemmParams& p, const FragCh& fc, uint32_t t_q, int T, int Tt, bool lrelu, float ns,
                                                   float* __restrict__ nct_seg, int ch0, int lane) {
    const int fcol = 2 * (lane & 3);
    const bool vec2 = (T & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        for (int c0 = 0; c0 < Tt; c0 += 32) {
            if (c0 >= T) break;
            uint32_t v[16];
            const int ng = min(4, (Tt - c0) >> 3);
            frag_load(t_q + (static_cast<uint32_t>(16 * hf) << 16) + c0, ng, v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int f = c0 + 8 * g + fcol;
                if (g >= ng || f >= T) continue;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const int sl = 2 * hf + rh;
                    if (!fc.ok[sl]) continue;
                    float* dst = nct_seg + static_cast<size_t>(ch0 + 8 * sl) * T + f;
                    float y[2], r[2] = {0.f, 0.f};
                    const bool two = f + 1 < T;
                    if (p.accumulate) {
                        if (vec2) { const float2 t2 = *reinterpret_cast<const float2*>(dst); r[0] = t2.x; r[1] = t2.y; }
                        else { r[0] = dst[0]; if (two) r[1] = dst[1]; }
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float x = __uint_as_float(v[4 * g + 2 * rh + e]) + fc.bias[sl];
                        if (lrelu) x = fmaxf(x, x * ns);
                        if (p.inorm) x = fmaf(x, fc.scale[sl], fc.shift[sl]);
                        if (p.act == ACT_SIGMOID) x = sigmoid_f(x);
                        else if (p.act == ACT_TANH) x = tanh_f(x);
                        if (p.accumulate == 1) x = r[e] + x;
                        else if (p.accumulate == 2) x = fmaf(r[e], x, r[e]);
                        y[e] = x;
                    }
                    if (vec2) *reinterpret_cast<float2*>(dst) = make_float2(y[0], y[1]);
                    else { dst[0] = y[0]; if (two) dst[1] = y[1]; }
                }
            }
        }
    }
}

}  // namespace zs
