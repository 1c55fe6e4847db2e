// HBM-bound companion kernels of the autoencoder path: layout packing, the discrete bottleneck,
// the GRU recurrence, weight packing and speaker-embedding bias folding.
#pragma once
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace zs {

// ---------------------------------------------------------------------------------------------
// (B, C, T) fp32  ->  channels-last operand buffer [B][rows][pitch] with reflected halo rows (zero_halo: zero rows,
// the 'constant' padding mode model/model.py:36-38 selects for seg_len < 64).
// 32x32 tile transpose through shared memory: reads coalesced along T, writes coalesced along C.
// ---------------------------------------------------------------------------------------------
template <typename OT>
__global__ void pack_nct_kernel(const float* __restrict__ x, OT* __restrict__ out, int C, int T, int rows, int pitch,
                                int halo, int choff, int c_fill, int lrelu, float ns, int zero_halo) {
    __shared__ float tile[32][33];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, t = t0 + tx;
        float v = 0.f;
        if (c < C && t < T) v = x[(static_cast<size_t>(b) * C + c) * T + t];
        tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
    OT* ob = out + static_cast<size_t>(b) * rows * pitch + choff;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + 8 * i, c = c0 + tx;
        if (t >= T || c >= c_fill) continue;
        float v = tile[tx][ty + 8 * i];
        if (lrelu) v = fmaxf(v, v * ns);
        const OT y = float_to_ot<OT>(v);
        ob[static_cast<size_t>(halo + t) * pitch + c] = y;
        if (halo > 0) {
            const OT hv = zero_halo ? float_to_ot<OT>(0.f) : y;
            if (t >= 1 && t <= halo) ob[static_cast<size_t>(halo - t) * pitch + c] = hv;
            if (t >= T - 1 - halo && t <= T - 2) ob[static_cast<size_t>(halo + 2 * (T - 1) - t) * pitch + c] = hv;
        }
    }
}

// The encoder's two views of its input in ONE pass over x (model/model.py:441-446): the conv bank reads x with a
// 3-frame halo (`bank`, no activation), and conv2 reads cat([bank outputs, x]) after a leaky-relu (`cat`, channel offset
// `cat_choff`, no halo).  IT = float | __half (an fp16 upload halves the PCIe bytes and rounds exactly like the fp32 path
// does here).  NTC = false: x is (B, C, T) as Encoder.forward receives it - tiles of 64 channels x 32 frames, 128-byte reads
// along T, 128-byte half2 writes along C.  NTC = true: x is (B, T, C), the layout Trainer.test_step is handed before its
// permute (trainer.py:196) - already channels-last, so the tile is read along C and no transpose is needed.
__device__ __forceinline__ float in_to_float(float v) { return v; }
__device__ __forceinline__ float in_to_float(__half v) { return __half2float(v); }

template <typename OT, typename IT, bool NTC>
__global__ void __launch_bounds__(256) pack_x_dual_kernel(const IT* __restrict__ x, int C, int T,
                                                          OT* __restrict__ bank, int bank_rows, int bank_pitch, int bank_halo,
                                                          OT* __restrict__ cat, int cat_rows, int cat_pitch, int cat_choff, int cat_fill,
                                                          float ns, int zero_halo) {
    __shared__ float tile[64][33];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 64, b = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!NTC) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + ty + 8 * i, t = t0 + tx;
            tile[ty + 8 * i][tx] = (c < C && t < T) ? in_to_float(x[(static_cast<size_t>(b) * C + c) * T + t]) : 0.f;
        }
        __syncthreads();
    }
    OT* bb = bank + static_cast<size_t>(b) * bank_rows * bank_pitch;
    OT* cb = cat + static_cast<size_t>(b) * cat_rows * cat_pitch + cat_choff;
    const int c = c0 + 2 * tx;                       // this lane's channel pair
    using OT2 = typename std::conditional<std::is_same<OT, __half>::value, __half2, __nv_bfloat162>::type;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty + 8 * i;
        if (t >= T) continue;
        float v0, v1;
        if (NTC) {
            const IT* xr = x + (static_cast<size_t>(b) * T + t) * C;
            v0 = c < C ? in_to_float(xr[c]) : 0.f;
            v1 = c + 1 < C ? in_to_float(xr[c + 1]) : 0.f;
        } else {
            v0 = tile[2 * tx][ty + 8 * i];
            v1 = tile[2 * tx + 1][ty + 8 * i];
        }
        if (c < bank_pitch) {                        // channels >= C are the zero padding of the K dimension
            OT2 y; y.x = float_to_ot<OT>(v0); y.y = float_to_ot<OT>(v1);
            OT2 hv = y;
            if (zero_halo) { hv.x = float_to_ot<OT>(0.f); hv.y = hv.x; }
            *reinterpret_cast<OT2*>(bb + static_cast<size_t>(bank_halo + t) * bank_pitch + c) = y;
            if (t >= 1 && t <= bank_halo) *reinterpret_cast<OT2*>(bb + static_cast<size_t>(bank_halo - t) * bank_pitch + c) = hv;
            if (t >= T - 1 - bank_halo && t <= T - 2)
                *reinterpret_cast<OT2*>(bb + static_cast<size_t>(bank_halo + 2 * (T - 1) - t) * bank_pitch + c) = hv;
        }
        if (c < cat_fill) {
            const float l0 = fmaxf(v0, v0 * ns), l1 = fmaxf(v1, v1 * ns);
            OT* dst = cb + static_cast<size_t>(t) * cat_pitch + c;
            if (((cat_choff + c) & 1) == 0 && c + 1 < cat_fill) {
                OT2 y; y.x = float_to_ot<OT>(l0); y.y = float_to_ot<OT>(l1);
                *reinterpret_cast<OT2*>(dst) = y;
            } else {
                dst[0] = float_to_ot<OT>(l0);
                if (c + 1 < cat_fill) dst[1] = float_to_ot<OT>(l1);
            }
        }
    }
}

// The same two views for the (B, C, T) layout with 16-byte traffic on both sides: a CTA takes 64 channels x 128 frames, a warp
// reads a WHOLE 512-byte channel row per request (one float4 per lane; the 32-frame tiles of the kernel above read 128-byte pieces
// of 64 different rows), the tile sits in shared memory as float4 with the float4 index XOR-ed by f(ch) = (ch >> 3) ^ (ch & 7)
// (conflict-free both for the row-wise writes and for the channel-wise reads below), and a thread writes 8 consecutive channels of
// one frame (16 bytes; 8 lanes cover a 128-byte piece of an operand row).  Needs T % 4 == 0 (16-byte aligned rows) and 16-byte
// aligned operand rows (pitches and the concat offset multiples of 8); everything else takes the kernel above.
template <typename OT, typename IT>
__global__ void __launch_bounds__(256) pack_x_dual_wide_kernel(const IT* __restrict__ x, int C, int T,
                                                               OT* __restrict__ bank, int bank_rows, int bank_pitch, int bank_halo,
                                                               OT* __restrict__ cat, int cat_rows, int cat_pitch, int cat_choff, int cat_fill,
                                                               float ns, int zero_halo) {
    __shared__ float4 tile[64][32];
    const int t0 = blockIdx.x * 128, c0 = blockIdx.y * 64, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ch = warp + 8 * i, c = c0 + ch, t = t0 + 4 * lane;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C && t < T) {                         // T % 4 == 0: a float4 is inside or outside as a whole
            const IT* src = x + (static_cast<size_t>(b) * C + c) * T + t;
            if (sizeof(IT) == 4) {
                v = *reinterpret_cast<const float4*>(src);
            } else {
                const uint2 h = *reinterpret_cast<const uint2*>(src);
                const __half2 h0 = *reinterpret_cast<const __half2*>(&h.x), h1 = *reinterpret_cast<const __half2*>(&h.y);
                v = make_float4(__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1));
            }
        }
        tile[ch][lane ^ (((ch >> 3) ^ ch) & 7)] = v;
    }
    __syncthreads();
    OT* bb = bank + static_cast<size_t>(b) * bank_rows * bank_pitch;
    OT* cb = cat + static_cast<size_t>(b) * cat_rows * cat_pitch + cat_choff;
    const int g = lane & 7, ts = lane >> 3;           // 8 channels x one of 4 consecutive frames
    const int c = c0 + 8 * g;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int tl = 4 * (warp + 8 * i) + ts, t = t0 + tl;
        if (t >= T) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = 8 * g + j;
            const float4 q = tile[ch][(tl >> 2) ^ (((ch >> 3) ^ ch) & 7)];
            v[j] = (tl & 3) == 0 ? q.x : ((tl & 3) == 1 ? q.y : ((tl & 3) == 2 ? q.z : q.w));
        }
        if (c < bank_pitch) {                         // channels >= C are the zero padding of the K dimension (v = 0 there)
            uint4 y;
            OT* yv = reinterpret_cast<OT*>(&y);
#pragma unroll
            for (int j = 0; j < 8; ++j) yv[j] = float_to_ot<OT>(v[j]);
            const uint4 hv = zero_halo ? make_uint4(0, 0, 0, 0) : y;
            *reinterpret_cast<uint4*>(bb + static_cast<size_t>(bank_halo + t) * bank_pitch + c) = y;
            if (t >= 1 && t <= bank_halo) *reinterpret_cast<uint4*>(bb + static_cast<size_t>(bank_halo - t) * bank_pitch + c) = hv;
            if (t >= T - 1 - bank_halo && t <= T - 2)
                *reinterpret_cast<uint4*>(bb + static_cast<size_t>(bank_halo + 2 * (T - 1) - t) * bank_pitch + c) = hv;
        }
        if (c < cat_fill) {
            OT* dst = cb + static_cast<size_t>(t) * cat_pitch + c;
            if (c + 8 <= cat_fill) {
                uint4 y;
                OT* yv = reinterpret_cast<OT*>(&y);
#pragma unroll
                for (int j = 0; j < 8; ++j) yv[j] = float_to_ot<OT>(fmaxf(v[j], v[j] * ns));
                *reinterpret_cast<uint4*>(dst) = y;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (c + j < cat_fill) dst[j] = float_to_ot<OT>(fmaxf(v[j], v[j] * ns));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Discrete bottleneck, one_hot mode: ids[b,t] = argmax_c(logits[b,c,t] + noise[b,t,c]) with first-index
// tie-break (torch.max), act = one-hot (skipped when the caller passes no `act`: the batched front-end only
// wants the ids).  One CTA per segment; the (C x T8) logits tile is staged in shared memory (16-byte global loads) so
// both the logits (time-fastest) and the noise (unit-fastest, 16-byte loads) are read coalesced.
// noise == nullptr: the Gumbel noise is generated in the kernel from a counter-based generator (splitmix64 of the
// segment's seed + (t, c) -> 24-bit uniform like torch.rand -> -log(-log(u + eps) + eps), model/model.py:95-98): the
// throughput mode of the streaming front-end - same distribution, not the reference's CPU generator stream.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gumbel_from_counter(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = static_cast<float>(static_cast<unsigned int>(z >> 40)) * (1.0f / 16777216.0f);
    return -logf(-logf(u + 1e-20f) + 1e-20f);
}

__global__ void __launch_bounds__(512) bottleneck_onehot_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                                                const unsigned long long* __restrict__ seg_seeds, int C, int T8,
                                                                float* __restrict__ act, int* __restrict__ ids) {
    extern __shared__ float s_log[];  // [C][T8 + 1]
    int* s_id = reinterpret_cast<int*>(s_log + static_cast<size_t>(C) * (T8 + 1));
    const int b = blockIdx.x;
    const int n = C * T8;
    const float* lb = logits + static_cast<size_t>(b) * n;
    if ((T8 & 3) == 0 && (reinterpret_cast<uintptr_t>(lb) & 15) == 0) {
        const int q = T8 >> 2;                      // float4 per channel row
        for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
            const float4 v = reinterpret_cast<const float4*>(lb)[i];
            float* d = s_log + (i / q) * (T8 + 1) + ((i % q) << 2);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_log[(i / T8) * (T8 + 1) + (i % T8)] = lb[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t = warp; t < T8; t += nwarps) {
        const size_t row = (static_cast<size_t>(b) * T8 + t) * C;
        const float* nz = noise ? noise + row : nullptr;
        const unsigned long long seed = nz ? 0ull : seg_seeds[b];      // per-segment stream: independent of the batch it rides in
        float best = -INFINITY;
        int bi = 0x7fffffff;
        auto take = [&](int c, float g) {
            const float v = s_log[c * (T8 + 1) + t] + g;
            if (v > best || bi == 0x7fffffff) {     // strictly greater: the first index wins ties (torch.max)
                best = v;
                bi = c;
            }
        };
        if (nz != nullptr && (C & 127) == 0 && (reinterpret_cast<uintptr_t>(nz) & 15) == 0) {
            for (int c4 = lane * 4; c4 < C; c4 += 128) {         // lane visits its channels in increasing order
                const float4 g = *reinterpret_cast<const float4*>(nz + c4);
                take(c4, g.x); take(c4 + 1, g.y); take(c4 + 2, g.z); take(c4 + 3, g.w);
            }
        } else {
            for (int c = lane; c < C; c += 32) take(c, nz ? nz[c] : gumbel_from_counter(seed, static_cast<unsigned long long>(t) * C + c));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            s_id[t] = bi;
            if (ids) ids[b * T8 + t] = bi;
        }
    }
    if (act == nullptr) return;
    __syncthreads();
    float* ab = act + static_cast<size_t>(b) * n;
    if ((T8 & 3) == 0 && (reinterpret_cast<uintptr_t>(ab) & 15) == 0) {
        const int q = T8 >> 2;
        for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
            const int c = i / q, t = (i % q) << 2;
            reinterpret_cast<float4*>(ab)[i] = make_float4(c == s_id[t] ? 1.f : 0.f, c == s_id[t + 1] ? 1.f : 0.f,
                                                           c == s_id[t + 2] ? 1.f : 0.f, c == s_id[t + 3] ? 1.f : 0.f);
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) ab[i] = (i / T8 == s_id[i % T8]) ? 1.f : 0.f;
    }
}

// continues / multilabel_binary / gumbel_t / binary epilogues of Encoder.forward (model/model.py:457-484)
__global__ void bottleneck_misc_kernel(const float* __restrict__ logits, const float* __restrict__ noise, int mode,
                                       int B, int E, int T8, float ns, float* __restrict__ act) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (mode == 0) {  // continues: leaky-relu of the logits
        if (i < static_cast<size_t>(B) * E * T8) {
            const float v = logits[i];
            act[i] = fmaxf(v, v * ns);
        }
    } else if (mode == 2) {  // multilabel_binary: bit = (argmax over the pair == 0); ties -> index 0
        if (i < static_cast<size_t>(B) * E * T8) {
            const int t = i % T8, e = (i / T8) % E, b = i / (static_cast<size_t>(T8) * E);
            const float* lb = logits + static_cast<size_t>(b) * 2 * E * T8;
            const float* nz = noise + ((static_cast<size_t>(b) * T8 + t) * E + e) * 2;
            const float v0 = lb[(2 * e) * T8 + t] + nz[0], v1 = lb[(2 * e + 1) * T8 + t] + nz[1];
            act[i] = (v0 >= v1) ? 1.f : 0.f;
        }
    } else if (mode == 4) {  // binary (model/model.py:466-472): channel r*E + c; every row r picks one column c (argmax of
                             // logits + noise, first index on ties), act[c] = 1 if ANY row picked c.  act is pre-zeroed.
        if (i < static_cast<size_t>(B) * T8 * E) {
            const int r = i % E, t = (i / E) % T8, b = i / (static_cast<size_t>(E) * T8);
            const float* lb = logits + (static_cast<size_t>(b) * E * E + static_cast<size_t>(r) * E) * T8 + t;
            const float* nz = noise + ((static_cast<size_t>(b) * T8 + t) * E + r) * E;
            float best = lb[0] + nz[0];
            int bi = 0;
            for (int c = 1; c < E; ++c) {
                const float v = lb[static_cast<size_t>(c) * T8] + nz[c];
                if (v > best) {
                    best = v;
                    bi = c;
                }
            }
            act[(static_cast<size_t>(b) * E + bi) * T8 + t] = 1.f;
        }
    } else {  // gumbel_t: one-hot over the TIME axis of each (segment, unit) row
        if (i < static_cast<size_t>(B) * E) {
            const float* lb = logits + i * T8;
            const float* nz = noise + i * T8;
            float best = lb[0] + nz[0];
            int bi = 0;
            for (int t = 1; t < T8; ++t) {
                const float v = lb[t] + nz[t];
                if (v > best) {
                    best = v;
                    bi = t;
                }
            }
            for (int t = 0; t < T8; ++t) act[i * T8 + t] = (t == bi) ? 1.f : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bidirectional GRU recurrence, zero initial state (model/model.py:59-66).  Straightforward CUDA-core
// version: one CTA per (direction, group of NBG segments), one thread per hidden unit, W_hh^T streamed
// from L2 each step.  gx = W_ih x + b_ih (+ folded speaker term) comes from the tensor-core GEMM.
// ---------------------------------------------------------------------------------------------
template <typename OT, int NBG>
__global__ void gru_simple_kernel(const OT* __restrict__ gx, const float* __restrict__ whhT,
                                  const float* __restrict__ bhh, int B, int T, int H, OT* __restrict__ out,
                                  int rows, int pitch, int halo, int choff, OT* __restrict__ gates) {
    extern __shared__ float s_h[];  // [NBG][H]
    const int dir = blockIdx.y, b0 = blockIdx.x * NBG, j = threadIdx.x;
    const float* W = whhT + static_cast<size_t>(dir) * H * 3 * H;
    const float br = bhh[dir * 3 * H + j], bz = bhh[dir * 3 * H + H + j], bn = bhh[dir * 3 * H + 2 * H + j];
    float h[NBG];
#pragma unroll
    for (int s = 0; s < NBG; ++s) {
        h[s] = 0.f;
        s_h[s * H + j] = 0.f;
    }
    __syncthreads();
    for (int step = 0; step < T; ++step) {
        const int t = dir ? T - 1 - step : step;
        float ar[NBG], az[NBG], an[NBG];
#pragma unroll
        for (int s = 0; s < NBG; ++s) ar[s] = az[s] = an[s] = 0.f;
        for (int k = 0; k < H; ++k) {
            const float wr = W[static_cast<size_t>(k) * 3 * H + j], wz = W[static_cast<size_t>(k) * 3 * H + H + j],
                        wn = W[static_cast<size_t>(k) * 3 * H + 2 * H + j];
#pragma unroll
            for (int s = 0; s < NBG; ++s) {
                const float hk = s_h[s * H + k];
                ar[s] = fmaf(wr, hk, ar[s]);
                az[s] = fmaf(wz, hk, az[s]);
                an[s] = fmaf(wn, hk, an[s]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NBG; ++s) {
            const int b = b0 + s;
            if (b < B) {
                const OT* g = gx + ((static_cast<size_t>(b) * T + t) * 2 + dir) * 3 * H;
                const float r = 1.f / (1.f + expf(-(ot_to_float<OT>(g[j]) + ar[s] + br)));
                const float z = 1.f / (1.f + expf(-(ot_to_float<OT>(g[H + j]) + az[s] + bz)));
                const float n = tanhf(ot_to_float<OT>(g[2 * H + j]) + r * (an[s] + bn));
                h[s] = (1.f - z) * n + z * h[s];
                s_h[s * H + j] = h[s];
                if (gates != nullptr) {   // training: r, z, n, hn per step [B][T][2][4][H]
                    OT* gs = gates + ((static_cast<size_t>(b) * T + t) * 2 + dir) * 4 * H + j;
                    gs[0] = float_to_ot<OT>(r); gs[H] = float_to_ot<OT>(z); gs[2 * H] = float_to_ot<OT>(n); gs[3 * H] = float_to_ot<OT>(an[s] + bn);
                }
                out[(static_cast<size_t>(b) * rows + halo + t) * pitch + choff + dir * H + j] = float_to_ot<OT>(h[s]);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Weight packing (runs once per load_state_dict)
// ---------------------------------------------------------------------------------------------
// pixel-shuffle row permutation: conv output channel co = 2c + r  ->  row tile(c/64)*128 + r*64 + c%64
__host__ __device__ inline int ps_row(int co) {
    const int c = co >> 1, r = co & 1;
    return (c >> 6) * 128 + r * 64 + (c & 63);
}

// W (C_out, C_in, k) fp32 -> dst[row_off + row(co)][(tap_off + j) * c_in_pad + ci] operand type,
// for input channels ci_lo <= ci_lo + ci < ci_lo + ci_n.  dst is pre-zeroed.
// One block per (output channel, chunk of PACK_CI input channels): the (ci, j) -> (j, ci) transpose goes through
// shared memory so both the fp32 reads and the operand writes are contiguous runs.
constexpr int PACK_CI = 256;
template <typename OT>
__global__ void pack_weight_kernel(const float* __restrict__ W, OT* __restrict__ dst, int C_out, int C_in, int k,
                                   int ci_lo, int ci_n, long long k_total, int c_in_pad, int tap_off, int row_off,
                                   int ps) {
    extern __shared__ float pack_sw[];                    // [PACK_CI * k]
    const int co = blockIdx.y, c0 = blockIdx.x * PACK_CI;
    const int n = min(PACK_CI, ci_n - c0);
    const float* src = W + (static_cast<long long>(co) * C_in + ci_lo + c0) * k;
    for (int i = threadIdx.x; i < n * k; i += blockDim.x) pack_sw[i] = src[i];
    __syncthreads();
    const int row = row_off + (ps ? ps_row(co) : co);
    OT* d = dst + row * k_total + static_cast<long long>(tap_off) * c_in_pad + c0;
    for (int i = threadIdx.x; i < n * k; i += blockDim.x) {
        const int j = i / n, ci = i - j * n;
        d[static_cast<long long>(j) * c_in_pad + ci] = float_to_ot<OT>(pack_sw[ci * k + j]);
    }
    (void)C_out;
}
template <typename OT>
inline cudaError_t launch_pack_weight(const float* W, OT* dst, int C_out, int C_in, int k, int ci_lo, int ci_n,
                                      long long k_total, int c_in_pad, int tap_off, int row_off, int ps, cudaStream_t st) {
    dim3 grid((ci_n + PACK_CI - 1) / PACK_CI, C_out);
    pack_weight_kernel<OT><<<grid, 256, PACK_CI * k * sizeof(float), st>>>(W, dst, C_out, C_in, k, ci_lo, ci_n, k_total, c_in_pad,
                                                                          tap_off, row_off, ps);
    return cudaGetLastError();
}

// tab[s][row_off + row(co)] = (b ? b[co] : 0) + sum_{ci < C_e, j < k} W[co][ci_lo + ci][j] * emb[s][ci]
// one warp per (speaker, out channel).  C_e = 0 gives the plain (padded, permuted) bias vector.
// j_only >= 0: only tap j_only enters the sum (edge-correction tables of the zero-padding mode).
__global__ void fold_bias_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                 const float* __restrict__ emb, float* __restrict__ tab, int C_out, int C_in, int k,
                                 int ci_lo, int C_e, int n_spk, int m_rows, int row_off, int ps, int j_only = -1) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<long long>(n_spk) * C_out) return;
    const int co = w % C_out, s = w / C_out;
    float acc = 0.f;
    const float* wr = W + (static_cast<long long>(co) * C_in + ci_lo) * k;
    const float* e = emb + static_cast<long long>(s) * C_e;
    for (int i = lane; i < C_e * k; i += 32)
        if (j_only < 0 || i % k == j_only) acc = fmaf(wr[i], e[i / k], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) tab[static_cast<long long>(s) * m_rows + row_off + (ps ? ps_row(co) : co)] = acc + (b ? b[co] : 0.f);
}

template <typename OT>
__global__ void cast_to_ot_kernel(const float* __restrict__ x, OT* __restrict__ y, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) y[i] = float_to_ot<OT>(x[i]);
}

// W_hh (3H, H) fp32 -> W^T [H][3H] fp32 for the CUDA-core recurrence
__global__ void transpose_whh_kernel(const float* __restrict__ W, float* __restrict__ WT, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * H * H) return;
    const int k = i % H, g = i / H;
    WT[static_cast<size_t>(k) * 3 * H + g] = W[i];
}

// input_emb (c_h, c_in) -> table [c_in][c_h] operand type (rounded like the GEMM operand would be)
template <typename OT>
__global__ void transpose_emb_kernel(const float* __restrict__ W, OT* __restrict__ WT, int c_h, int c_in) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(c_h) * c_in) return;
    const int u = i % c_in, c = i / c_in;
    WT[static_cast<long long>(u) * c_h + c] = float_to_ot<OT>(W[i]);
}

// Decoder input from unit ids: x0[b][halo + t][c] = input_emb.weight[c][id] + bias[c], reflected halo rows.
// One thread = 8 consecutive channels of one unit frame (16-byte loads of the transposed table row and 16-byte stores of the
// operand row; c_h and the pitch are multiples of 8), blockDim.y unit frames per CTA; the scalar tail covers c_h % 8.
template <typename OT>
__global__ void unit_gather_kernel(const int* __restrict__ ids, const OT* __restrict__ WT,
                                   const float* __restrict__ bias, OT* __restrict__ out, int T8, int c_h, int rows,
                                   int pitch, int halo, int n_units, int zero_halo) {
    const int t = blockIdx.x * blockDim.y + threadIdx.y, b = blockIdx.y;
    if (t >= T8) return;
    int id = ids[b * T8 + t];
    id = min(max(id, 0), n_units - 1);
    OT* ob = out + static_cast<size_t>(b) * rows * pitch;
    const OT* wr = WT + static_cast<size_t>(id) * c_h;
    const bool top = halo > 0 && t >= 1 && t <= halo, bot = halo > 0 && t >= T8 - 1 - halo && t <= T8 - 2;
    const bool vec = (c_h & 7) == 0 && (pitch & 7) == 0 && ((reinterpret_cast<uintptr_t>(WT) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
    const int c_vec = vec ? c_h : 0;
    for (int c = threadIdx.x * 8; c < c_vec; c += blockDim.x * 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(wr + c);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c), b1 = *reinterpret_cast<const float4*>(bias + c + 4);
        const OT* wv = reinterpret_cast<const OT*>(&w);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint4 y;
        OT* yv = reinterpret_cast<OT*>(&y);
#pragma unroll
        for (int i = 0; i < 8; ++i) yv[i] = float_to_ot<OT>(ot_to_float<OT>(wv[i]) + bb[i]);
        *reinterpret_cast<uint4*>(ob + static_cast<size_t>(halo + t) * pitch + c) = y;
        const uint4 hv = zero_halo ? make_uint4(0, 0, 0, 0) : y;
        if (top) *reinterpret_cast<uint4*>(ob + static_cast<size_t>(halo - t) * pitch + c) = hv;
        if (bot) *reinterpret_cast<uint4*>(ob + static_cast<size_t>(halo + 2 * (T8 - 1) - t) * pitch + c) = hv;
    }
    for (int c = c_vec + threadIdx.x; c < c_h; c += blockDim.x) {
        const OT y = float_to_ot<OT>(ot_to_float<OT>(wr[c]) + bias[c]);
        ob[static_cast<size_t>(halo + t) * pitch + c] = y;
        const OT hv = zero_halo ? float_to_ot<OT>(0.f) : y;
        if (top) ob[static_cast<size_t>(halo - t) * pitch + c] = hv;
        if (bot) ob[static_cast<size_t>(halo + 2 * (T8 - 1) - t) * pitch + c] = hv;
    }
}

}  // namespace zs
