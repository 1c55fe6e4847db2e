// HBM-bound companion kernels of the pretrain_AE step (trainer.py:321-332): everything that is not a GEMM.
//
// Gradients of activations travel as fp16 channels-last buffers multiplied by the loss scale S (fp16 x bf16
// operand mixes are an illegal tcgen05 instruction, so the weight-gradient GEMM needs both operands in the
// activations' format); weight gradients leave the GEMMs as unscaled fp32.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace zs {

// ---- dropout mask: counter-based, reproducible in the backward pass ------------------------------------
// keep(seed, layer, b, c, t) with P(keep) = 1 - p; or an explicit reference-layout (B, C, T) byte mask
__device__ __forceinline__ uint32_t mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint32_t>(z >> 32);
}
struct DropSpec {
    float p;                          // 0 = no dropout
    unsigned long long seed;          // step seed (value) ...
    const unsigned long long* seed_dev;   // ... or read from device memory (CUDA-graph replays change it without re-capture)
    int layer;                        // Dropout layer index, mixed into the seed
    const uint8_t* keep;              // explicit mask (B, C, T) or null
    int C, T;
};
__device__ __forceinline__ unsigned long long drop_layer_seed(const DropSpec& d) {
    unsigned long long z = (d.seed_dev ? *d.seed_dev : d.seed) + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(d.layer + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// `ls` = drop_layer_seed(d), hoisted by the caller
__device__ __forceinline__ float drop_scale(const DropSpec& d, unsigned long long ls, int b, int c, int t) {
    if (d.p <= 0.f) return 1.f;
    const size_t idx = (static_cast<size_t>(b) * d.C + c) * d.T + t;
    bool keep;
    if (d.keep) keep = d.keep[idx] != 0;
    else keep = (mix64(ls ^ (idx * 0x2545F4914F6CDD1Dull)) >> 8) * (1.f / 16777216.f) >= d.p;
    return keep ? 1.f / (1.f - d.p) : 0.f;
}

// a channels-last fp16 view [B][rows][pitch] whose frame t lives at row halo + t, channels [choff, choff + C)
struct ClView {
    __half* p;
    int rows, pitch, halo, choff;
};
__device__ __forceinline__ __half* cl_at(const ClView& v, int b, int row, int c) {
    return v.p + (static_cast<size_t>(b) * v.rows + row) * v.pitch + v.choff + c;
}
// write frame t (and its reflected halo copies) of a buffer with `halo` reflected rows
__device__ __forceinline__ void cl_store_reflect(const ClView& v, int b, int t, int T, int c, __half2 val) {
    *reinterpret_cast<__half2*>(cl_at(v, b, v.halo + t, c)) = val;
    if (v.halo > 0) {
        if (t >= 1 && t <= v.halo) *reinterpret_cast<__half2*>(cl_at(v, b, v.halo - t, c)) = val;
        if (t >= T - 1 - v.halo && t <= T - 2) *reinterpret_cast<__half2*>(cl_at(v, b, v.halo + 2 * (T - 1) - t, c)) = val;
    }
}

// ---------------------------------------------------------------------------------------------
// Forward glue of the training path:  y = dropout(xhat) + residual,  ye = y + emb[spk]   (both with reflected halos)
//   residual: none / same frame / nearest-up-2 (frame t/2) / avg-pool-2 (frames 2t, 2t+1)
// One thread = one (segment, frame, channel pair).
// ---------------------------------------------------------------------------------------------
struct CombineParams {
    ClView x;                 // xhat (InstanceNorm output) or any plain activation
    ClView res; int res_mode; // RES_*
    ClView y;                 // p == null: not written
    ClView ye; const float* emb; const long long* spk; int emb_pitch, n_spk;   // p == null: not written
    ClView bc;                // p != null: bc[b][t][c] = emb[spk[b]][c] (append_emb, model/model.py:81-85)
    DropSpec drop;
    int B, T, C;
};
__global__ void combine_fwd_kernel(const CombineParams p) {
    const int c2 = p.C >> 1;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<size_t>(p.B) * p.T * c2) return;
    const int c = static_cast<int>(i % c2) * 2, t = static_cast<int>((i / c2) % p.T), b = static_cast<int>(i / (static_cast<size_t>(c2) * p.T));
    float2 v = __half22float2(*reinterpret_cast<const __half2*>(cl_at(p.x, b, p.x.halo + t, c)));
    if (p.drop.p > 0.f) {
        const unsigned long long ls = drop_layer_seed(p.drop);
        v.x *= drop_scale(p.drop, ls, b, c, t);
        v.y *= drop_scale(p.drop, ls, b, c + 1, t);
    }
    if (p.res_mode == RES_SAME) {
        const float2 r = __half22float2(*reinterpret_cast<const __half2*>(cl_at(p.res, b, p.res.halo + t, c)));
        v.x += r.x; v.y += r.y;
    } else if (p.res_mode == RES_UP2) {
        const float2 r = __half22float2(*reinterpret_cast<const __half2*>(cl_at(p.res, b, p.res.halo + (t >> 1), c)));
        v.x += r.x; v.y += r.y;
    } else if (p.res_mode == RES_AVG2) {
        const float2 r0 = __half22float2(*reinterpret_cast<const __half2*>(cl_at(p.res, b, p.res.halo + 2 * t, c)));
        const float2 r1 = __half22float2(*reinterpret_cast<const __half2*>(cl_at(p.res, b, p.res.halo + 2 * t + 1, c)));
        v.x += 0.5f * (r0.x + r1.x); v.y += 0.5f * (r0.y + r1.y);
    }
    if (p.y.p) cl_store_reflect(p.y, b, t, p.T, c, __floats2half2_rn(v.x, v.y));
    if (p.ye.p || p.bc.p) {
        long long sp = p.spk[b];
        sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
        const float2 e = *reinterpret_cast<const float2*>(p.emb + static_cast<size_t>(sp) * p.emb_pitch + c);
        if (p.ye.p) cl_store_reflect(p.ye, b, t, p.T, c, __floats2half2_rn(v.x + e.x, v.y + e.y));
        if (p.bc.p) *reinterpret_cast<__half2*>(cl_at(p.bc, b, p.bc.halo + t, c)) = __floats2half2_rn(e.x, e.y);
    }
}

// ---------------------------------------------------------------------------------------------
// L1 loss + its gradient through the output non-linearity (trainer.py:327; model/model.py:361-364):
//   loss += sum |spec - x| / N ;  dpre[b][t][c] = S * sign(spec - x) / N * act'(spec)   (channels-last fp16)
// 32x32 tile transpose: reads (B, C, T) fp32 coalesced along T, writes channels-last coalesced along C.
// ---------------------------------------------------------------------------------------------
__global__ void l1_loss_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ x, int C, int T,
                                   __half* __restrict__ dpre, int rows, int pitch, float g_scale /* S / N */,
                                   float inv_n, int tanh_out, float* __restrict__ loss) {
    __shared__ float tile[32][33];
    __shared__ float red[8];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, t = t0 + tx;
        float g = 0.f;
        if (c < C && t < T) {
            const size_t idx = (static_cast<size_t>(b) * C + c) * T + t;
            const float s = spec[idx], d = s - x[idx];
            part += fabsf(d);
            const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            g = sg * g_scale * (tanh_out ? (1.f - s * s) : s * (1.f - s));
        }
        tile[ty + 8 * i][tx] = g;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (tx == 0) red[ty] = part;
    __syncthreads();
    if (tx == 0 && ty == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss, s * inv_n);
    }
    __half* ob = dpre + static_cast<size_t>(b) * rows * pitch;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + 8 * i, c = c0 + tx;
        if (t < T && c < pitch) ob[static_cast<size_t>(t) * pitch + c] = __float2half_rn(c < C ? tile[tx][ty + 8 * i] : 0.f);
    }
}

// generic upstream gradient instead of the fused L1 loss: dpre = S * d_spec * act'(spec)
__global__ void dspec_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ d_spec, int C, int T,
                                 __half* __restrict__ dpre, int rows, int pitch, float scale, int tanh_out) {
    __shared__ float tile[32][33];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, t = t0 + tx;
        float g = 0.f;
        if (c < C && t < T) {
            const size_t idx = (static_cast<size_t>(b) * C + c) * T + t;
            const float s = spec[idx];
            g = d_spec[idx] * scale * (tanh_out ? (1.f - s * s) : s * (1.f - s));
        }
        tile[ty + 8 * i][tx] = g;
    }
    __syncthreads();
    __half* ob = dpre + static_cast<size_t>(b) * rows * pitch;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + 8 * i, c = c0 + tx;
        if (t < T && c < pitch) ob[static_cast<size_t>(t) * pitch + c] = __float2half_rn(c < C ? tile[tx][ty + 8 * i] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// Straight-through Gumbel-softmax backward (model/model.py:93-110): forward value is the one-hot, the gradient is
// the softmax's:  y = softmax((l + g) / tau),  dl = y * (da - sum_c y da) / tau.
//   logits (B, C, T8) fp32, noise (B, T8, C) fp32, dact (B, C, T8) fp32 (loss-scaled)  ->  dl channels-last fp16
// One CTA per segment; logits and dact tiles staged in shared memory so every global access is coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void gumbel_st_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                     const float* __restrict__ dact, int C, int T8, float inv_tau, float out_scale,
                                     __half* __restrict__ dl, int rows, int pitch) {
    extern __shared__ float s_buf[];          // logits [C][T8+1], dact [C][T8+1]
    float* s_log = s_buf;
    float* s_da = s_buf + static_cast<size_t>(C) * (T8 + 1);
    const int b = blockIdx.x, n = C * T8;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_log[(i / T8) * (T8 + 1) + (i % T8)] = logits[static_cast<size_t>(b) * n + i];
        s_da[(i / T8) * (T8 + 1) + (i % T8)] = dact[static_cast<size_t>(b) * n + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t = warp; t < T8; t += nwarps) {
        const float* nz = noise + (static_cast<size_t>(b) * T8 + t) * C;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, (s_log[c * (T8 + 1) + t] + nz[c]) * inv_tau);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float se = 0.f, sd = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float e = __expf((s_log[c * (T8 + 1) + t] + nz[c]) * inv_tau - mx);
            se += e;
            sd = fmaf(e, s_da[c * (T8 + 1) + t], sd);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, off);
            sd += __shfl_xor_sync(0xffffffffu, sd, off);
        }
        const float inv = 1.f / se, dot = sd * inv;
        __half* o = dl + (static_cast<size_t>(b) * rows + t) * pitch;
        for (int c = lane; c < C; c += 32) {
            const float y = __expf((s_log[c * (T8 + 1) + t] + nz[c]) * inv_tau - mx) * inv;
            o[c] = __float2half_rn(y * (s_da[c * (T8 + 1) + t] - dot) * out_scale);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward of one layer's element-wise tail.  The layer computed (training forward)
//     pre = conv(...) + bias ;  u = lrelu(pre) ;  xhat = InstanceNorm(u) ;  out = dropout(xhat) + residual      (IN layers)
//     out = lrelu(pre) + post_emb                                                                              (plain layers)
// and its output gradient arrives from up to three places:
//     A, B : zero-extended data-gradient GEMM outputs over the PADDED input of a consumer conv (pad rows each side,
//            reflect padding folded back here), B optional (two consumers)
//     R    : the materialised output gradient of a later layer that used this output as its residual
//            (same frame / it up-sampled us x2 / it avg-pooled us x2)
// Outputs: dpre (gradient of the GEMM accumulator, channels-last fp16 with ZERO halo rows, the B operand of the
// data-gradient GEMM and the A operand of the weight-gradient GEMM), optionally gsum (= total output gradient,
// for the residual consumer) and the speaker-embedding gradient (sum over frames of A, B or the total).
// One CTA = one segment x 64 channels; 32 channel pairs x 8 frame lanes.
// ---------------------------------------------------------------------------------------------
enum { GS_NONE = 0, GS_PADDED = 1, GS_SAME = 2, GS_UP2 = 3, GS_AVG2 = 4 };
struct GradSrc {
    const __half* p;
    int rows, pitch, choff, pad, mode;
};
struct ActBwdParams {
    GradSrc a, b, r;
    const __half* fwd; int f_rows, f_pitch, f_halo, f_choff;   // xhat (IN layers) or the stored output (plain layers)
    const float* stats; int stat_pitch;                         // (mean, rstd) per (segment, channel); null = no InstanceNorm
    int lrelu; float ns;
    const float* post_emb; const long long* spk; int emb_pitch, n_spk;   // plain layers stored out = lrelu(pre) + emb
    DropSpec drop;
    __half* dpre; int d_rows, d_pitch, d_halo, d_choff;
    __half* gsum; int g_rows, g_pitch;
    float* demb; int demb_from; float demb_scale;               // demb[spk[b]][c] += demb_scale * sum_t {1: A, 2: B, 3: total}
    int B, T, C;
};
__device__ __forceinline__ float2 ld_h2(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
__device__ __forceinline__ float2 grad_src_at(const GradSrc& s, int b, int t, int T, int c) {
    if (s.mode == GS_NONE) return make_float2(0.f, 0.f);
    const __half* base = s.p + static_cast<size_t>(b) * s.rows * s.pitch + s.choff + c;
    if (s.mode == GS_PADDED) {
        float2 v = ld_h2(base + static_cast<size_t>(s.pad + t) * s.pitch);
        if (t >= 1 && t <= s.pad) {              // padded row pad - t mirrors frame t
            const float2 w = ld_h2(base + static_cast<size_t>(s.pad - t) * s.pitch);
            v.x += w.x; v.y += w.y;
        }
        if (t >= T - 1 - s.pad && t <= T - 2) {  // padded row pad + 2(T-1) - t mirrors frame t
            const float2 w = ld_h2(base + static_cast<size_t>(s.pad + 2 * (T - 1) - t) * s.pitch);
            v.x += w.x; v.y += w.y;
        }
        return v;
    }
    if (s.mode == GS_SAME) return ld_h2(base + static_cast<size_t>(t) * s.pitch);
    if (s.mode == GS_UP2) {
        const float2 v0 = ld_h2(base + static_cast<size_t>(2 * t) * s.pitch), v1 = ld_h2(base + static_cast<size_t>(2 * t + 1) * s.pitch);
        return make_float2(v0.x + v1.x, v0.y + v1.y);
    }
    const float2 v = ld_h2(base + static_cast<size_t>(t >> 1) * s.pitch);   // GS_AVG2
    return make_float2(0.5f * v.x, 0.5f * v.y);
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const ActBwdParams p) {
    __shared__ float red[8][32][6];
    const int b = blockIdx.y, lane = threadIdx.x & 31, tl = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + lane * 2;
    const bool c_ok = c < p.C;
    const int T = p.T;
    float mean0 = 0.f, rstd0 = 1.f, mean1 = 0.f, rstd1 = 1.f, e0 = 0.f, e1 = 0.f;
    if (c_ok && p.stats) {
        const float4 st = *reinterpret_cast<const float4*>(p.stats + (static_cast<size_t>(b) * p.stat_pitch + c) * 2);
        mean0 = st.x; rstd0 = st.y; mean1 = st.z; rstd1 = st.w;
    }
    if (c_ok && p.post_emb) {
        long long sp = p.spk[b];
        sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
        const float2 e = *reinterpret_cast<const float2*>(p.post_emb + static_cast<size_t>(sp) * p.emb_pitch + c);
        e0 = e.x; e1 = e.y;
    }
    const __half* fwd = p.fwd ? p.fwd + static_cast<size_t>(b) * p.f_rows * p.f_pitch + p.f_choff + c : nullptr;
    const bool need_pass1 = p.stats != nullptr || p.demb != nullptr;
    const unsigned long long ls = p.drop.p > 0.f ? drop_layer_seed(p.drop) : 0ull;
    float s1x = 0.f, s1y = 0.f, s2x = 0.f, s2y = 0.f, ex = 0.f, ey = 0.f;
    if (need_pass1 && c_ok) {
        for (int t = tl; t < T; t += 8) {
            const float2 ga = grad_src_at(p.a, b, t, T, c), gb = grad_src_at(p.b, b, t, T, c), gr = grad_src_at(p.r, b, t, T, c);
            float gx = ga.x + gb.x + gr.x, gy = ga.y + gb.y + gr.y;
            if (p.demb_from == 1) { ex += ga.x; ey += ga.y; }
            else if (p.demb_from == 2) { ex += gb.x; ey += gb.y; }
            else if (p.demb_from == 3) { ex += gx; ey += gy; }
            if (p.stats) {
                const float2 xh = ld_h2(fwd + static_cast<size_t>(p.f_halo + t) * p.f_pitch);
                if (p.drop.p > 0.f) { gx *= drop_scale(p.drop, ls, b, c, t); gy *= drop_scale(p.drop, ls, b, c + 1, t); }
                s1x += gx; s1y += gy;
                s2x = fmaf(gx, xh.x, s2x); s2y = fmaf(gy, xh.y, s2y);
            }
        }
    }
    if (need_pass1) {
        red[tl][lane][0] = s1x; red[tl][lane][1] = s1y; red[tl][lane][2] = s2x; red[tl][lane][3] = s2y;
        red[tl][lane][4] = ex; red[tl][lane][5] = ey;
        __syncthreads();
        s1x = s1y = s2x = s2y = ex = ey = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s1x += red[i][lane][0]; s1y += red[i][lane][1]; s2x += red[i][lane][2]; s2y += red[i][lane][3];
            ex += red[i][lane][4]; ey += red[i][lane][5];
        }
        if (p.demb && c_ok && tl == 0) {
            long long sp = p.spk[b];
            sp = sp < 0 ? 0 : (sp >= p.n_spk ? p.n_spk - 1 : sp);
            atomicAdd(p.demb + static_cast<size_t>(sp) * p.emb_pitch + c, ex * p.demb_scale);
            atomicAdd(p.demb + static_cast<size_t>(sp) * p.emb_pitch + c + 1, ey * p.demb_scale);
        }
    }
    if (!c_ok) return;
    const float inv_T = 1.f / static_cast<float>(T);
    const float m1x = s1x * inv_T, m1y = s1y * inv_T, m2x = s2x * inv_T, m2y = s2y * inv_T;
    __half* dp = p.dpre ? p.dpre + static_cast<size_t>(b) * p.d_rows * p.d_pitch + p.d_choff + c : nullptr;
    __half* gs = p.gsum ? p.gsum + static_cast<size_t>(b) * p.g_rows * p.g_pitch + c : nullptr;
    for (int t = tl; t < T; t += 8) {
        const float2 ga = grad_src_at(p.a, b, t, T, c), gb = grad_src_at(p.b, b, t, T, c), gr = grad_src_at(p.r, b, t, T, c);
        float gx = ga.x + gb.x + gr.x, gy = ga.y + gb.y + gr.y;
        if (gs) *reinterpret_cast<__half2*>(gs + static_cast<size_t>(t) * p.g_pitch) = __floats2half2_rn(gx, gy);
        if (!dp) continue;
        float2 f = make_float2(1.f, 1.f);
        if (fwd) f = ld_h2(fwd + static_cast<size_t>(p.f_halo + t) * p.f_pitch);
        float ux, uy;   // sign carriers of the pre-activation
        if (p.stats) {
            if (p.drop.p > 0.f) { gx *= drop_scale(p.drop, ls, b, c, t); gy *= drop_scale(p.drop, ls, b, c + 1, t); }
            gx = rstd0 * (gx - m1x - f.x * m2x);
            gy = rstd1 * (gy - m1y - f.y * m2y);
            ux = f.x + mean0 * rstd0;     // u = xhat / rstd + mean has the sign of xhat + mean * rstd
            uy = f.y + mean1 * rstd1;
        } else {
            ux = f.x - e0;
            uy = f.y - e1;
        }
        if (p.lrelu) {
            if (ux < 0.f) gx *= p.ns;
            if (uy < 0.f) gy *= p.ns;
        }
        *reinterpret_cast<__half2*>(dp + static_cast<size_t>(p.d_halo + t) * p.d_pitch) = __floats2half2_rn(gx, gy);
    }
    if (dp && p.d_halo > 0) {   // the GEMMs read these rows as zero padding
        for (int h = tl; h < p.d_halo; h += 8) {
            *reinterpret_cast<__half2*>(dp + static_cast<size_t>(h) * p.d_pitch) = __floats2half2_rn(0.f, 0.f);
            *reinterpret_cast<__half2*>(dp + static_cast<size_t>(p.d_halo + T + h) * p.d_pitch) = __floats2half2_rn(0.f, 0.f);
        }
    }
}

// out[c] += scale * sum_{rows} buf[row][choff + c]   (bias gradients; zero halo rows contribute nothing)
__global__ void colsum_kernel(const __half* __restrict__ buf, long long n_rows, int pitch, int choff, int C, float scale,
                              float* __restrict__ out, int ps_c) {
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), rl = threadIdx.x >> 6;   // 256 threads: 64 channels x 4 row lanes
    __shared__ float red[4][64];
    float s = 0.f;
    if (c < C) {
        const long long per = (n_rows + gridDim.y - 1) / gridDim.y;
        const long long r0 = blockIdx.y * per, r1 = min(n_rows, r0 + per);
        for (long long r = r0 + rl; r < r1; r += 4) s += __half2float(buf[r * pitch + choff + c]);
    }
    red[rl][threadIdx.x & 63] = s;
    __syncthreads();
    if (rl == 0 && c < C) {
        s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
        // pixel-shuffle layers keep channel m = r * ps_c + cc for conv output channel 2 cc + r
        const int oc = ps_c > 0 ? 2 * (c % ps_c) + c / ps_c : c;
        atomicAdd(out + oc, s * scale);
    }
}

// ---------------------------------------------------------------------------------------------
// GRU backward through time, CUDA-core version (one CTA per direction x group of NBG sequences, one thread per
// hidden unit).  Forward saved the gates r, z, n and hn = W_hn h + b_hn per step.
//   dh_t   = dOut_t + carry ;  dn = dh (1 - z) ;  dz = dh (h_prev - n) ;  carry_direct = dh z
//   da_n   = dn (1 - n^2) ;  da_z = dz z (1 - z) ;  da_r = da_n hn r (1 - r)
//   dgx_t  = (da_r, da_z, da_n)            (gradient of the input projection  -> W_ih, b_ih, input)
//   dgh_t  = (da_r, da_z, da_n r)          (gradient of W_hh h + b_hh         -> W_hh, b_hh)
//   carry  = carry_direct + W_hh^T dgh_t
// ---------------------------------------------------------------------------------------------
template <int NBG>
__global__ void gru_bptt_simple_kernel(const __half* __restrict__ gates /* [B][T][2][4][H] r,z,n,hn */,
                                       const __half* __restrict__ hbuf, int h_rows, int h_pitch, int h_choff,
                                       const __half* __restrict__ dout, int do_rows, int do_pitch, int do_choff,
                                       const float* __restrict__ w_hh0, const float* __restrict__ w_hh1, int B, int T, int H,
                                       __half* __restrict__ dgx, __half* __restrict__ dgh /* [B][T][2][3H] */) {
    extern __shared__ float s_g[];   // [NBG][3H]
    const int dir = blockIdx.y, b0 = blockIdx.x * NBG, j = threadIdx.x;
    const float* W = dir ? w_hh1 : w_hh0;     // (3H, H) row-major: column j is read coalesced across threads
    float carry[NBG];
#pragma unroll
    for (int s = 0; s < NBG; ++s) carry[s] = 0.f;
    for (int step = T - 1; step >= 0; --step) {
        const int t = dir ? T - 1 - step : step;           // time index processed at this step of the direction
        const int tp = dir ? t + 1 : t - 1;                // time index of h_prev
        float keep[NBG];
#pragma unroll
        for (int s = 0; s < NBG; ++s) {
            const int b = b0 + s;
            float ar = 0.f, az = 0.f, an = 0.f, anr = 0.f;
            keep[s] = 0.f;
            if (b < B) {
                const __half* g = gates + ((static_cast<size_t>(b) * T + t) * 2 + dir) * 4 * H + j;
                const float r = __half2float(g[0]), z = __half2float(g[H]), n = __half2float(g[2 * H]), hn = __half2float(g[3 * H]);
                const float hp = step > 0 ? __half2float(hbuf[(static_cast<size_t>(b) * h_rows + tp) * h_pitch + h_choff + dir * H + j]) : 0.f;
                const float dh = carry[s] + __half2float(dout[(static_cast<size_t>(b) * do_rows + t) * do_pitch + do_choff + dir * H + j]);
                const float dn = dh * (1.f - z), dz = dh * (hp - n);
                keep[s] = dh * z;
                an = dn * (1.f - n * n);
                az = dz * z * (1.f - z);
                ar = an * hn * r * (1.f - r);
                anr = an * r;
                const size_t o = ((static_cast<size_t>(b) * T + t) * 2 + dir) * 3 * H + j;
                dgx[o] = __float2half_rn(ar); dgx[o + H] = __float2half_rn(az); dgx[o + 2 * H] = __float2half_rn(an);
                dgh[o] = __float2half_rn(ar); dgh[o + H] = __float2half_rn(az); dgh[o + 2 * H] = __float2half_rn(anr);
            }
            s_g[s * 3 * H + j] = ar; s_g[s * 3 * H + H + j] = az; s_g[s * 3 * H + 2 * H + j] = anr;
        }
        __syncthreads();
        float acc[NBG];
#pragma unroll
        for (int s = 0; s < NBG; ++s) acc[s] = 0.f;
        for (int g = 0; g < 3 * H; ++g) {
            const float w = W[static_cast<size_t>(g) * H + j];
#pragma unroll
            for (int s = 0; s < NBG; ++s) acc[s] = fmaf(w, s_g[s * 3 * H + g], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < NBG; ++s) carry[s] = keep[s] + acc[s];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Weights for the data-gradient GEMM (run by conv_gemm_kernel): W (C_out, C_in, k) fp32 ->
//   dst[row(ci)][tap' * c_out_pad + kk(co)]  fp16, K-major over (tap', output channel)
// mode 0 (stride-1 conv / linear): row = ci, tap' = k-1-j  (dXpad[t'] = sum_j W_j^T dpre[t' - j])
// mode 1 (stride-2 conv, odd k = 2p+1): rows pixel-shuffle-permuted by the parity r = j & 1 of the padded output
//        row 2w'+r, tap' = p - (j - r)/2  (p+1 taps; dXpad[2w'+r] = sum_{j = r mod 2} W_j^T dpre[w' - (j-r)/2])
// ps_c > 0: the forward conv was pixel-shuffled: its gradient buffer keeps channel kk = r*ps_c + c for co = 2c + r
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int ps_row_ci(int ci, int r) { return (ci >> 6) * 128 + r * 64 + (ci & 63); }
// One block per tile of PT_CO output channels x PT_CI input channels: the (co) <-> (ci) transpose goes through shared
// memory, reads are runs of PT_CI * k floats per output channel, writes runs of up to PT_CO halves per (ci, tap).
constexpr int PT_CO = 64, PT_CI = 16;
__global__ void pack_weight_T_kernel(const float* __restrict__ W, __half* __restrict__ dst, int C_out, int C_in, int k,
                                     int ci_n, long long k_total, int c_out_pad, int mode, int ps_c) {
    extern __shared__ float pt_sw[];                      // [PT_CO][PT_CI * k + 1]
    const int co0 = blockIdx.y * PT_CO, ci0 = blockIdx.x * PT_CI;
    const int n_co = min(PT_CO, C_out - co0), n_ci = min(PT_CI, ci_n - ci0);
    const int rowlen = PT_CI * k + 1, run = n_ci * k;
    for (int i = threadIdx.x; i < n_co * run; i += blockDim.x) {
        const int co = i / run, r = i - co * run;
        pt_sw[co * rowlen + r] = W[(static_cast<long long>(co0 + co) * C_in + ci0) * k + r];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PT_CO * run; i += blockDim.x) {
        const int col = i & (PT_CO - 1), r = i / PT_CO;  // r = ci * k + j
        // consecutive threads write consecutive K columns kk: with ps_c the even / odd output channels form two runs
        const int co = ps_c > 0 ? ((col & 31) << 1 | (col >> 5)) : col;
        if (co >= n_co) continue;
        const int ci = ci0 + r / k, j = r % k;
        const int gco = co0 + co;
        const int kk = ps_c > 0 ? (gco & 1) * ps_c + (gco >> 1) : gco;
        int row, tap;
        if (mode == 0) { row = ci; tap = k - 1 - j; }
        else { const int rr = j & 1; row = ps_row_ci(ci, rr); tap = (k >> 1) - ((j - rr) >> 1); }
        dst[row * k_total + static_cast<long long>(tap) * c_out_pad + kk] = __float2half_rn(pt_sw[co * rowlen + r]);
    }
}
inline cudaError_t launch_pack_weight_T(const float* W, __half* dst, int C_out, int C_in, int k, int ci_n, long long k_total,
                                        int c_out_pad, int mode, int ps_c, cudaStream_t st) {
    dim3 grid((ci_n + PT_CI - 1) / PT_CI, (C_out + PT_CO - 1) / PT_CO);
    pack_weight_T_kernel<<<grid, 256, PT_CO * (PT_CI * k + 1) * sizeof(float), st>>>(W, dst, C_out, C_in, k, ci_n, k_total, c_out_pad, mode, ps_c);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Optimiser (trainer.py:64-66, 330-332; utils.py:53-55) on flat fp32 buffers
// ---------------------------------------------------------------------------------------------
// out[0] += sum g^2 ; non-finite values make it non-finite, which the step kernel treats as "skip"
__global__ void sqnorm_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
    float s = 0.f;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = g[i];
        s = fmaf(v, v, s);
    }
    __shared__ float red[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}
// clip_grad_norm_(max_norm) folded into Adam: g *= min(1, max_norm / (norm + 1e-6)); torch.optim.Adam update.
// `sq` = squared gradient norm of THIS network (device scalar); `skip` (device int, may be null) is set when
// the norm is not finite - the step is then a no-op (dynamic loss scaling).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, const float* __restrict__ sq, float grad_mult, float max_norm,
                            float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt, const float* __restrict__ bc_dev,
                            int* __restrict__ skip) {
    if (bc_dev) {     // the step state lives in device memory (train_meta_*_kernel): {bc1, bc2_sqrt, apply}
        if (reinterpret_cast<const int*>(bc_dev)[2] == 0) return;      // some network's gradient overflowed: nobody steps
        bc1 = bc_dev[0];
        bc2_sqrt = bc_dev[1];
    }
    const float norm = sqrtf(*sq) * grad_mult;
    if (!isfinite(norm)) {
        if (skip && blockIdx.x == 0 && threadIdx.x == 0) *skip = 1;
        return;
    }
    const float coef = fminf(1.f, max_norm / (norm + 1e-6f)) * grad_mult;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
}

// ---------------------------------------------------------------------------------------------
// Device-resident step state of the pretrain_AE iteration, so that a CUDA-graph replay (or an eager step) advances it on
// the stream instead of through host writes that a still-running earlier replay could read too late.
// 8 x 32-bit words: [0,1] dropout seed (u64) | [2] 1 - beta1^n | [3] sqrt(1 - beta2^n) | [4] apply (int) |
//                   [5] n = optimiser steps actually applied | [6,7] iterations started (u64)
// ---------------------------------------------------------------------------------------------
__global__ void train_meta_begin_kernel(unsigned long long* __restrict__ meta, unsigned long long salt) {
    const unsigned long long it = meta[3] + 1;
    meta[3] = it;
    unsigned long long z = it * 0x9E3779B97F4A7C15ull + salt;           // splitmix64 of (iteration, rank salt)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    meta[0] = z ^ (z >> 31);
}
// after the gradient norms are known: either EVERY network steps (bias corrections of the next applied step) or none does
__global__ void train_meta_commit_kernel(float* __restrict__ meta, const float* __restrict__ sq_a, const float* __restrict__ sq_b,
                                         float beta1, float beta2, int* __restrict__ skipped) {
    int* mi = reinterpret_cast<int*>(meta);
    const bool ok = isfinite(*sq_a) && (sq_b == nullptr || isfinite(*sq_b));
    if (ok) {
        const int n = mi[5] + 1;
        mi[5] = n;
        meta[2] = 1.f - powf(beta1, static_cast<float>(n));
        meta[3] = sqrtf(1.f - powf(beta2, static_cast<float>(n)));
        mi[4] = 1;
    } else {
        mi[4] = 0;
        if (skipped) *skipped = 1;
    }
}

}  // namespace zs
