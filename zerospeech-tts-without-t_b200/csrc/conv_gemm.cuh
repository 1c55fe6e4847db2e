// Fused conv1d / per-frame-linear layer as an implicit GEMM on tcgen05 tensor cores.
//
//   D[m = out channel, n = (segment, frame)] = sum_{tap, ci} W[m][tap][ci] * X[segment][row0 + frame*stride + tap][ci]
//
// A (weights, [m_rows][w_taps * c_in_pad], K-major) and B (activations, channels-last with halo rows,
// [B][rows][pitch]) are both fetched by TMA with the 128-byte swizzle; a tap shift is a row offset of the
// B box, a stride-2 conv reads the buffer through a (channel, row parity, row pair, segment) view so the
// box stays dense.
// One CTA per SM, persistent over output tiles of 128 channels x (nb segments x Tt frames <= 256 columns).
// Warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..5 = epilogue.  Two fp32
// accumulators of 256 TMEM columns each let the epilogue of tile i overlap the main loop of tile i+1.
//
// Epilogue (one thread = one output channel = one TMEM lane; the frames of a segment are its columns, so
// InstanceNorm statistics never leave the thread): + bias[speaker] -> leaky-relu -> InstanceNorm ->
// + residual (same frame / avg-pool-2 / nearest-up-2) -> sigmoid|tanh -> store (channels-last with
// reflected halo rows for the next conv, pixel-shuffled channels-last, or the reference's (B, C, T) fp32).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace zs {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int MAX_BN = 256;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
constexpr float IN_EPS = 1e-5f;

enum { RES_NONE = 0, RES_SAME = 1, RES_AVG2 = 2, RES_UP2 = 3 };
enum { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_TANH = 2 };
enum { OUT_CL = 0, OUT_PS = 1, OUT_NCT32 = 2, OUT_CL32 = 3 };

struct alignas(64) GemmParams {
    CUtensorMap tmA;
    CUtensorMap tmB;
    int m_tiles, n_tiles, nb, Tt, T, B, N;
    int kc, taps, bank, stride, in_row0, c_in_pad;
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
    // saturate instead of overflowing to inf: fp16 operands carry tf32-class mantissa but less range
    return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
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

// Epilogue of one (segment, channel): all T frames of the segment are columns [t_seg, t_seg + T) of this
// thread's TMEM lane.  RES / OUT are compile-time so the per-element code has no mode branches and the
// residual loads of a 16-frame chunk are issued together, ahead of the TMEM load they overlap with.
template <typename OT, int RES, int OUT>
__device__ __forceinline__ void epilogue_segment(const GemmParams& p, uint32_t t_seg, int b, int ch, bool ch_ok,
                                                 int out_ch, int ps_r, float bias) {
    const OT* __restrict__ res = reinterpret_cast<const OT*>(p.res);
    OT* __restrict__ out_cl = reinterpret_cast<OT*>(p.out);
    float* __restrict__ out_f = reinterpret_cast<float*>(p.out);
    const int T = p.T;
    const float inv_T = 1.f / static_cast<float>(T);
    const bool lrelu = p.lrelu != 0;
    const float ns = p.ns;
    float mean = 0.f, rstd = 1.f;
    if (p.inorm) {
        float sum = 0.f;
        for (int c0 = 0; c0 < T; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_seg + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x = __uint_as_float(v[i]) + bias;
                if (lrelu) x = fmaxf(x, x * ns);
                if (c0 + i < T) sum += x;
            }
        }
        mean = sum * inv_T;
        float sq = 0.f;
        for (int c0 = 0; c0 < T; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_seg + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x = __uint_as_float(v[i]) + bias;
                if (lrelu) x = fmaxf(x, x * ns);
                const float d = x - mean;
                if (c0 + i < T) sq += d * d;
            }
        }
        rstd = rsqrtf(sq * inv_T + IN_EPS);
    }
    const OT* res_ch = res + static_cast<size_t>(b) * p.res_rows * p.res_pitch +
                       static_cast<size_t>(p.res_halo) * p.res_pitch + ch;
    const int T_out = (OUT == OUT_PS) ? 2 * T : T;
    const size_t out_base = static_cast<size_t>(b) * p.out_rows * p.out_pitch + p.out_choff + out_ch;
    float* nct = out_f + (static_cast<size_t>(b) * p.m_valid + ch) * T;
    const bool vec4 = (T & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int c0 = 0; c0 < T; c0 += 16) {
        __syncwarp();  // stores below are predicated per lane; re-converge for the aligned TMEM load
        uint32_t v[16];
        tmem_ld16(t_seg + c0, v);
        float r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0.f;
        if (ch_ok) {
            if (RES == RES_SAME) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < T) r[i] = ot_to_float<OT>(res_ch[static_cast<size_t>(c0 + i) * p.res_pitch]);
            } else if (RES == RES_UP2) {
#pragma unroll
                for (int i = 0; i < 16; i += 2)
                    if (c0 + i < T) r[i] = r[i + 1] = ot_to_float<OT>(res_ch[static_cast<size_t>((c0 + i) >> 1) * p.res_pitch]);
            } else if (RES == RES_AVG2) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < T) {
                        const OT* q = res_ch + static_cast<size_t>(2 * (c0 + i)) * p.res_pitch;
                        r[i] = 0.5f * (ot_to_float<OT>(q[0]) + ot_to_float<OT>(q[p.res_pitch]));
                    }
            }
            if (OUT == OUT_NCT32 && p.accumulate) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < T) r[i] = nct[c0 + i];
            }
        }
        tmem_ld_wait();
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float y = __uint_as_float(v[i]) + bias;
            if (lrelu) y = fmaxf(y, y * ns);
            y = (y - mean) * rstd;
            if (RES != RES_NONE) y += r[i];
            if (p.act == ACT_SIGMOID) y = sigmoid_f(y);
            else if (p.act == ACT_TANH) y = tanh_f(y);
            if (OUT == OUT_NCT32) {
                if (p.accumulate == 1) y = r[i] + y;
                else if (p.accumulate == 2) y = r[i] + r[i] * y;
            }
            x[i] = y;
        }
        if (!ch_ok) continue;
        if (OUT == OUT_NCT32) {
            if (vec4 && c0 + 16 <= T) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(nct + c0 + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < T) nct[c0 + i] = x[i];
            }
        } else if (OUT == OUT_CL32) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < T) out_f[out_base + static_cast<size_t>(p.out_halo + c0 + i) * p.out_pitch] = x[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int t = c0 + i;
                if (t < T) {
                    const int f = (OUT == OUT_PS) ? 2 * t + ps_r : t;
                    const OT y = float_to_ot<OT>(x[i]);
                    out_cl[out_base + static_cast<size_t>(p.out_halo + f) * p.out_pitch] = y;
                    if (p.out_halo > 0) {  // reflected halo rows for the next conv's taps
                        if (f >= 1 && f <= p.out_halo)
                            out_cl[out_base + static_cast<size_t>(p.out_halo - f) * p.out_pitch] = y;
                        if (f >= T_out - 1 - p.out_halo && f <= T_out - 2)
                            out_cl[out_base + static_cast<size_t>(p.out_halo + 2 * (T_out - 1) - f) * p.out_pitch] = y;
                    }
                }
            }
        }
    }
}

template <typename OT>
__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_gemm_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.m_tiles * p.n_tiles;
    const uint32_t stage_tx = A_STAGE_BYTES + static_cast<uint32_t>(p.N) * (BK * 2);

    if (warp == 0) {
        // ------------------------------ TMA producer (whole warp, one elected lane issues) ------
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
                int tap_lo = 0, ntaps = p.taps;
                if (p.bank) {
                    ntaps = mt + 1;
                    tap_lo = 3 - ntaps / 2;
                }
                for (int j = 0; j < ntaps; ++j) {
                    const int tap = tap_lo + j;
                    const int row_b = p.in_row0 + tap;
                    for (int c = 0; c < p.kc; ++c) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full[stage], stage_tx);
                            tma_load_2d(&p.tmA, sA + stage * A_STAGE_BYTES, &full[stage], tap * p.c_in_pad + c * BK,
                                        mt * BM);
                            if (p.stride == 2)  // buffer viewed as (channel, row parity, row pair, segment)
                                tma_load_4d(&p.tmB, sB + stage * B_STAGE_BYTES, &full[stage], c * BK, row_b & 1,
                                            row_b >> 1, nt * p.nb);
                            else
                                tma_load_3d(&p.tmB, sB + stage * B_STAGE_BYTES, &full[stage], c * BK, row_b,
                                            nt * p.nb);
                        }
                        __syncwarp();
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (whole warp, one elected lane issues) --------
        {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int mt = tile % p.m_tiles;
                const int ksteps = (p.bank ? mt + 1 : p.taps) * p.kc;
                const int as = it & 1;
                mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * MAX_BN;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(a0 + stage * A_STAGE_BYTES);
                    const uint64_t db = umma_desc_sw128(b0 + stage * B_STAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 elements (32 B) along K inside the swizzle row: +2 in the >>4 address field
                            umma_f16(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (ks | k) != 0);
                        }
                        umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit(&tfull[as]);
                __syncwarp();
            }
        }
    } else {
        // ------------------------------ epilogue ----------------------------------
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
            const int as = it & 1;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const int ch = mt * BM + row;  // weight row == bias index
            const bool ch_ok = ch < p.m_valid;
            // pixel shuffle: weight rows are packed so rows [0,64) of a tile hold r = 0, [64,128) r = 1
            const int ps_r = row >> 6;
            const int out_ch = (p.out_mode == OUT_PS) ? (mt * 64 + (row & 63)) : ch;
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * MAX_BN;
            for (int s = 0; s < p.nb; ++s) {
                const int b = nt * p.nb + s;
                if (b >= p.B) break;
                float bias = 0.f;
                if (p.bias != nullptr) {  // tables are padded to m_tiles * 128 rows
                    size_t off = 0;
                    if (p.spk) {
                        long long sp = p.spk[b];
                        sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
                        off = static_cast<size_t>(sp) * p.bias_stride;
                    }
                    bias = p.bias[off + ch];
                }
                const uint32_t t_seg = t_lane + s * p.Tt;
                if (p.out_mode == OUT_CL) {
                    if (p.res_mode == RES_NONE) epilogue_segment<OT, RES_NONE, OUT_CL>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                    else if (p.res_mode == RES_SAME) epilogue_segment<OT, RES_SAME, OUT_CL>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                    else if (p.res_mode == RES_AVG2) epilogue_segment<OT, RES_AVG2, OUT_CL>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                    else epilogue_segment<OT, RES_UP2, OUT_CL>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                } else if (p.out_mode == OUT_PS) {
                    epilogue_segment<OT, RES_NONE, OUT_PS>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                } else if (p.out_mode == OUT_NCT32) {
                    epilogue_segment<OT, RES_NONE, OUT_NCT32>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                } else {
                    epilogue_segment<OT, RES_NONE, OUT_CL32>(p, t_seg, b, ch, ch_ok, out_ch, ps_r, bias);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace zs
