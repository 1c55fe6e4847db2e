// STFT-domain kernels either side of the autoencoder path:
//   * Griffin-Lim vocoder   convert.py:39-62  (300 x { istft, stft, phase projection }, then the final istft)
//   * featurisation         preprocess.py:231-256 (pre-emphasis, stft, |.|, dB, normalise)
// Constants: hps/hps.py:22-33 (n_fft 1024, hop 200, Hann window of 800 centred in the 1024-sample frame).
//
// One Griffin-Lim iteration is ONE kernel: a CTA owns a tile of GL_F consecutive frames of one utterance, rebuilds the
// waveform under them in shared memory (inverse real FFTs of the GL_F + 6 frames whose windows reach into the tile,
// overlap-added in a fixed order and divided by the window sum-of-squares exactly as librosa.istft does), then
// re-analyses it (reflect padding at the utterance ends, window, forward real FFT) and writes mag * est / max(1e-8, |est|).
// The waveform never exists in HBM between the two transforms; per iteration and frame the kernel reads ~1.23 spectra
// (the halo frames come from L2) and writes one.
//
// FFT: a real 1024-point transform is a 512-point complex FFT (one warp, data in 4 KB of shared memory, three radix-8
// Stockham passes, twiddles from a table computed in double on the host) plus the even/odd split pass.  fp32 throughout -
// the reference runs numpy's FFT in float64 and rounds every spectrum / signal to complex64 / float32 (librosa).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace zs {

constexpr int ST_NFFT = 1024, ST_HOP = 200, ST_WIN = 800, ST_NBIN = 513;
constexpr int ST_WPAD = (ST_NFFT - ST_WIN) / 2;      // 112 zero samples each side of the window inside a frame
constexpr int GL_F = 26;                             // frames a tile re-analyses
constexpr int GL_INV = GL_F + 6;                     // frames whose synthesis windows reach into the tile's samples: 32 = 4 phases x 8 warps
constexpr int GL_WARPS = 8, GL_THREADS = GL_WARPS * 32;
constexpr int GL_YT = (GL_F - 1) * ST_HOP + ST_WIN;  // 5800 waveform samples under a tile
constexpr int GL_ZBUF = 512;                         // one warp's FFT buffer: 512 complex values
constexpr int GL_WTAB = 1024;                        // twiddle table
constexpr int GL_SMEM_BYTES = GL_WTAB * 8 + ST_WIN * 4 + GL_YT * 4 + GL_WARPS * GL_ZBUF * 8;   // twiddles + window + waveform tile + FFT buffers

// Index maps of the FFT buffer / twiddle table.  A skew (i + i / 8, i + i / 16) removes the 8-way bank conflicts of the first
// radix-8 pass's stores, but was measured SLOWER on a B200 (Griffin-Lim x 300 on 32 768 frames: 116 ms vs 98 ms): the kernel
// is bound by instruction issue and latency, not by shared-memory wavefronts, and the skew costs two integer
// instructions per access.  Kept as the identity so the experiment is one line away.
__device__ __forceinline__ int ZIDX(int i) { return i; }
__device__ __forceinline__ int WIDX(int i) { return i; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 8-point forward DFT in registers, natural order in and out
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
    const float r = 0.70710678118654752440f;
    float2 s[4], d[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        s[t] = make_float2(v[t].x + v[t + 4].x, v[t].y + v[t + 4].y);
        d[t] = make_float2(v[t].x - v[t + 4].x, v[t].y - v[t + 4].y);
    }
    d[1] = make_float2((d[1].x + d[1].y) * r, (d[1].y - d[1].x) * r);        // * (1 - i) / sqrt 2
    d[2] = make_float2(d[2].y, -d[2].x);                                      // * -i
    d[3] = make_float2((d[3].y - d[3].x) * r, -(d[3].x + d[3].y) * r);       // * (-1 - i) / sqrt 2
    auto dft4 = [](const float2 (&u)[4], float2& y0, float2& y1, float2& y2, float2& y3) {
        const float2 e0 = make_float2(u[0].x + u[2].x, u[0].y + u[2].y), e1 = make_float2(u[0].x - u[2].x, u[0].y - u[2].y);
        const float2 o0 = make_float2(u[1].x + u[3].x, u[1].y + u[3].y);
        const float2 o1 = make_float2(u[1].y - u[3].y, -(u[1].x - u[3].x));   // (u1 - u3) * -i
        y0 = make_float2(e0.x + o0.x, e0.y + o0.y);
        y1 = make_float2(e1.x + o1.x, e1.y + o1.y);
        y2 = make_float2(e0.x - o0.x, e0.y - o0.y);
        y3 = make_float2(e1.x - o1.x, e1.y - o1.y);
    };
    dft4(s, v[0], v[2], v[4], v[6]);
    dft4(d, v[1], v[3], v[5], v[7]);
}

// 512-point forward complex FFT of z (shared memory, this warp's buffer) in place; w = exp(-2 pi i m / 1024) table in
// shared memory.  Stockham autosort: natural order in, natural order out.
__device__ __forceinline__ void fft512_warp(float2* z, const float2* w, int lane) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const int Ns = pass == 0 ? 1 : (pass == 1 ? 8 : 64);
        float2 v[2][8];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lane + 32 * b, k = j & (Ns - 1);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[b][t] = z[ZIDX(j + 64 * t)];
            if (pass > 0) {
                const int step = k * (128 / Ns);              // twiddle exp(-2 pi i t k / (8 Ns)) = w[t * k * 1024 / (8 Ns)]
#pragma unroll
                for (int t = 1; t < 8; ++t) v[b][t] = cmul(v[b][t], w[WIDX((t * step) & 1023)]);
            }
            fft8(v[b]);
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int j = lane + 32 * b, k = j & (Ns - 1);
            const int j0 = ((j - k) << 3) + k;
#pragma unroll
            for (int t = 0; t < 8; ++t) z[ZIDX(j0 + t * Ns)] = v[b][t];
        }
        __syncwarp();
    }
}

// spectrum X[0..512] (global, one frame) -> the 1024 real samples of irfft(X) * 1024, left interleaved in z
// (sample 2n = z[n].x, sample 2n+1 = -z[n].y).  `real_in`: X is a real row (the first Griffin-Lim iteration starts from
// the magnitudes with zero phase).
__device__ __forceinline__ void irfft1024_warp(const float2* __restrict__ X, const float* __restrict__ Xreal, float2* z, const float2* w,
                                               int lane) {
#pragma unroll 4
    for (int k = lane; k < 512; k += 32) {
        float2 a, b;
        if (Xreal != nullptr) {
            a = make_float2(Xreal[k], 0.f);
            b = make_float2(Xreal[512 - k], 0.f);
        } else {
            a = X[k];
            b = X[512 - k];
        }
        if (k == 0) { a.y = 0.f; b.y = 0.f; }               // irfft ignores the imaginary parts of the DC and Nyquist bins
        const float2 xe = make_float2(a.x + b.x, a.y - b.y);  // X[k] + conj X[512-k]
        const float2 xd = make_float2(a.x - b.x, a.y + b.y);  // X[k] - conj X[512-k]
        const float2 wk = w[WIDX(k)];
        const float2 tw = make_float2(wk.x, -wk.y);           // exp(+2 pi i k / 1024)
        const float2 xo = cmul(xd, tw);
        // Z = xe + i xo; the inverse transform is conj(FFT(conj Z))
        z[ZIDX(k)] = make_float2(xe.x - xo.y, -(xe.y + xo.x));
    }
    __syncwarp();
    fft512_warp(z, w, lane);
}

// z holds the forward FFT of the packed frame (z[n] = x[2n] + i x[2n+1]); returns bin k of the 1024-point real transform
__device__ __forceinline__ float2 rfft_bin(const float2* z, const float2* w, int k) {
    const float2 zk = z[ZIDX(k & 511)], zm = z[ZIDX((512 - k) & 511)];
    const float2 ze = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));      // (Z[k] + conj Z[512-k]) / 2
    const float2 zd = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));      // (Z[k] - conj Z[512-k]) / 2
    const float2 t = cmul(zd, w[WIDX(k)]);                                           // * exp(-2 pi i k / 1024)
    return make_float2(ze.x + t.y, ze.y - t.x);                                      // ze - i t
}

struct StftMeta {            // per-utterance prefix sums, device int32 arrays of U + 1 entries each
    const int* frame_start;
    const int* sample_start;
    const int* tile_start;
};

__device__ __forceinline__ int find_utt(const int* tile_start, int U, int tile) {
    int lo = 0, hi = U - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct GlParams {
    const float2* x_in;      // [frames][513] complex spectra of the previous iteration (unused when first != 0)
    const float* mag;        // [frames][513] linear magnitudes (convert.py:57-58)
    float2* x_out;           // [frames][513] next spectra (unused when final != 0)
    float* wav;              // final != 0: [samples] waveform, utterance u at sample_start[u], 200 * (n_frames - 1) samples
    StftMeta meta;
    int U;
    const float2* w1024;     // exp(-2 pi i m / 1024)
    const float* win;        // Hann(800), periodic
    int first, final;
};

// mode bits in p: first (input = magnitudes, zero phase), final (write the waveform instead of re-analysing)
__global__ void __launch_bounds__(GL_THREADS) gl_iter_kernel(const GlParams p) {
    extern __shared__ __align__(16) uint8_t gl_smem[];
    float2* s_w = reinterpret_cast<float2*>(gl_smem);
    float* s_win = reinterpret_cast<float*>(gl_smem + GL_WTAB * 8);
    float* s_y = s_win + ST_WIN;
    float2* s_z = reinterpret_cast<float2*>(s_y + GL_YT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += GL_THREADS) s_w[WIDX(i)] = p.w1024[i];
    for (int i = threadIdx.x; i < ST_WIN; i += GL_THREADS) s_win[i] = p.win[i];
    for (int i = threadIdx.x; i < GL_YT; i += GL_THREADS) s_y[i] = 0.f;

    const int u = find_utt(p.meta.tile_start, p.U, blockIdx.x);
    const int n = p.meta.frame_start[u + 1] - p.meta.frame_start[u];        // frames of this utterance
    const int L = ST_HOP * (n - 1);                                          // samples of its waveform
    int f0 = (blockIdx.x - p.meta.tile_start[u]) * GL_F;
    if (f0 == n - 1 && n >= 2) f0 = n - 2;       // a one-frame last tile would need one more sample below its range (reflection)
    const int q0 = ST_HOP * f0 - 400;            // waveform index of s_y[0]
    const size_t frow = static_cast<size_t>(p.meta.frame_start[u]);
    float2* z = s_z + warp * GL_ZBUF;
    const float* zf = reinterpret_cast<const float*>(z);
    __syncthreads();

    // ---- synthesis: frames f0-3 .. f0+F+2, overlap-added in four phases (frames 4 apart never touch the same sample) ----
    for (int ph = 0; ph < 4; ++ph) {
        const int i = f0 - 3 + warp * 4 + ph;
        const bool live = i >= 0 && i < n;
        if (live) {
            const size_t row = (frow + i) * ST_NBIN;
            irfft1024_warp(p.first ? nullptr : p.x_in + row, p.first ? p.mag + row : nullptr, z, s_w, lane);
            for (int m = lane; m < ST_WIN; m += 32) {
                const int idx = ST_HOP * (i - f0) + m;          // q - q0: window sample m of frame i sits at q = 200 i - 400 + m
                if (idx >= 0 && idx < GL_YT) {
                    const int s = m + ST_WPAD;
                    const float zv = zf[2 * ZIDX(s >> 1) + (s & 1)];
                    const float x = (s & 1) ? -zv : zv;
                    s_y[idx] += x * (1.0f / 1024.0f) * s_win[m];
                }
            }
        }
        __syncthreads();
    }
    // ---- divide by the window sum-of-squares of the frames that exist (librosa.istft / window_sumsquare) ----
    for (int ty = threadIdx.x; ty < GL_YT; ty += GL_THREADS) {
        const int q = q0 + ty;
        if (q < 0 || q >= L) continue;
        const int ih = (q + 400) / ST_HOP;                      // newest frame whose window covers q
        const int m0 = (q + 400) - ST_HOP * ih;
        float wss = 0.f;
#pragma unroll
        for (int jj = 3; jj >= 0; --jj) {                       // oldest frame first, like the reference's accumulation
            const int i = ih - jj;
            if (i >= 0 && i < n) {
                const float wv = s_win[m0 + ST_HOP * jj];
                wss += wv * wv;
            }
        }
        if (wss > 1.17549435e-38f) s_y[ty] = __fdividef(s_y[ty], wss);
    }
    __syncthreads();

    if (p.final) {   // the tile owns the samples under its first F hops
        const int k = blockIdx.x - p.meta.tile_start[u];
        const int qa = ST_HOP * GL_F * k, qb = min(L, qa + ST_HOP * GL_F);
        float* out = p.wav + p.meta.sample_start[u];
        for (int q = qa + threadIdx.x; q < qb; q += GL_THREADS) out[q] = s_y[q - q0];
        return;
    }

    // ---- analysis of frames f0 .. f0+F-1 and the phase projection ----
    for (int j = f0 + warp; j < min(n, f0 + GL_F); j += GL_WARPS) {
        const int qj = ST_HOP * j - 400;                         // waveform index of window sample 0
        const bool interior = qj >= 0 && qj + ST_WIN <= L;       // no reflection needed (warp-uniform)
        const float2* y2 = reinterpret_cast<const float2*>(s_y + (qj - q0));      // even offset: 8-byte aligned pairs
        const float2* w2 = reinterpret_cast<const float2*>(s_win);
        for (int nn = lane; nn < 512; nn += 32) {
            float2 val = make_float2(0.f, 0.f);
            const int m = 2 * nn - ST_WPAD;                      // even: the pair (m, m + 1) is inside or outside the window together
            if (m >= 0 && m < ST_WIN) {
                if (interior) {
                    const float2 a = y2[m >> 1], b = w2[m >> 1];
                    val = make_float2(a.x * b.x, a.y * b.y);
                } else {
                    float v[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int q = qj + m + h;
                        if (q < 0) q = -q;                          // np.pad(..., mode='reflect')
                        if (q >= L) q = 2 * (L - 1) - q;
                        v[h] = s_y[q - q0] * s_win[m + h];
                    }
                    val = make_float2(v[0], v[1]);
                }
            }
            z[ZIDX(nn)] = val;
        }
        __syncwarp();
        fft512_warp(z, s_w, lane);
        const size_t row = (frow + j) * ST_NBIN;
        for (int k = lane; k <= 512; k += 32) {
            const float2 est = rfft_bin(z, s_w, k);
            // convert.py:48-49: mag * est / max(1e-8, |est|); one MUFU.RSQ instead of a square root and an IEEE division
            // (the division's FCHK slow-path check alone took a quarter of the kernel's stall samples)
            const float sc = p.mag[row + k] * rsqrtf(fmaxf(1e-16f, est.x * est.x + est.y * est.y));
            p.x_out[row + k] = make_float2(est.x * sc, est.y * sc);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// convert.py:57-58: normalised (T, 513) rows -> linear magnitudes  10^((clip(m, 0, 1) * 100 - 100 + 20) / 20)
// ---------------------------------------------------------------------------------------------
__global__ void denormalise_kernel(const float* __restrict__ spec, float* __restrict__ mag, size_t n, float max_db, float ref_db) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float db = fminf(fmaxf(spec[i], 0.f), 1.f) * max_db - max_db + ref_db;
    mag[i] = exp10f(db * 0.05f);
}

// convert.py:60: scipy.signal.lfilter([1], [1, -a]) = y[t] = x[t] + a y[t-1], per utterance.  Each thread produces 64 output
// samples; it starts the recurrence 1536 samples earlier from zero (a^1536 = 5e-21 for a = 0.97: below fp32 resolution of
// any carried state), so chunks are independent.  In place is NOT allowed (x != y).
__global__ void deemphasis_kernel(const float* __restrict__ x, float* __restrict__ y, const int* __restrict__ sample_start, int U,
                                  float a) {
    const int u = blockIdx.y;
    if (u >= U) return;
    const int s0 = sample_start[u], L = sample_start[u + 1] - s0;
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 64;
    if (c0 >= L) return;
    const float* xs = x + s0;
    float acc = 0.f;
    for (int t = max(0, c0 - 1536); t < c0; ++t) acc = fmaf(a, acc, xs[t]);
    float* ys = y + s0;
    const int c1 = min(L, c0 + 64);
    for (int t = c0; t < c1; ++t) {
        acc = fmaf(a, acc, xs[t]);
        ys[t] = acc;
    }
}

// librosa.feature.rms frames for librosa.effects.trim (convert.py:61, frame_length 2048, hop 512, centred with reflect
// padding): power[u][f] = mean(y_pad[512 f : 512 f + 2048]^2).  One warp per frame.
__global__ void frame_power_kernel(const float* __restrict__ y, const int* __restrict__ sample_start, const int* __restrict__ pframe_start,
                                   int U, float* __restrict__ power) {
    const int u = blockIdx.y;
    const int s0 = sample_start[u], L = sample_start[u + 1] - s0;
    const int nf = pframe_start[u + 1] - pframe_start[u];
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= nf || L <= 0) return;
    const float* ys = y + s0;
    float acc = 0.f;
    for (int t = lane; t < 2048; t += 32) {
        int q = 512 * f - 1024 + t;
        if (q < 0) q = -q;
        if (q >= L) q = 2 * (L - 1) - q;
        q = min(max(q, 0), L - 1);
        const float v = ys[q];
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) power[pframe_start[u] + f] = acc * (1.0f / 2048.0f);
}

// ---------------------------------------------------------------------------------------------
// Featurisation (preprocess.py:233-256): wav -> pre-emphasis -> stft -> |.| -> 20 log10(max(1e-5, .)) ->
// clip((db - ref_db + max_db) / max_db, 1e-8, 1), written as (T, 513) rows in fp32 and / or fp16 (the encoder's
// frames-major input format).  Tile = GL_F frames; the pre-emphasised, reflect-padded samples under it sit in shared memory.
// ---------------------------------------------------------------------------------------------
struct SpecParams {
    const float* wav;        // utterance u at sample_start[u]
    float* spec32;           // [frames][513] or null
    __half* spec16;          // [frames][513] or null
    StftMeta meta;
    int U;
    const float2* w1024;
    const float* win;
    float preemph, max_db, ref_db;
};

__global__ void __launch_bounds__(GL_THREADS) spec_kernel(const SpecParams p) {
    extern __shared__ __align__(16) uint8_t gl_smem[];
    float2* s_w = reinterpret_cast<float2*>(gl_smem);
    float* s_win = reinterpret_cast<float*>(gl_smem + GL_WTAB * 8);
    float* s_y = s_win + ST_WIN;
    float2* s_z = reinterpret_cast<float2*>(s_y + GL_YT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = find_utt(p.meta.tile_start, p.U, blockIdx.x);
    const int n = p.meta.frame_start[u + 1] - p.meta.frame_start[u];
    const int L = p.meta.sample_start[u + 1] - p.meta.sample_start[u];
    const int f0 = (blockIdx.x - p.meta.tile_start[u]) * GL_F;
    const int q0 = ST_HOP * f0 - 400;
    const float* xs = p.wav + p.meta.sample_start[u];
    for (int i = threadIdx.x; i < 1024; i += GL_THREADS) s_w[WIDX(i)] = p.w1024[i];
    for (int i = threadIdx.x; i < ST_WIN; i += GL_THREADS) s_win[i] = p.win[i];
    for (int ty = threadIdx.x; ty < GL_YT; ty += GL_THREADS) {
        int q = q0 + ty;
        if (q < 0) q = -q;                                      // reflect padding of the PRE-EMPHASISED signal
        if (q >= L) q = 2 * (L - 1) - q;
        float v = 0.f;
        if (q >= 0 && q < L) v = q == 0 ? xs[0] : xs[q] - p.preemph * xs[q - 1];          // preprocess.py:233
        s_y[ty] = v;
    }
    __syncthreads();
    float2* z = s_z + warp * GL_ZBUF;
    const size_t frow = static_cast<size_t>(p.meta.frame_start[u]);
    for (int j = f0 + warp; j < min(n, f0 + GL_F); j += GL_WARPS) {
        for (int nn = lane; nn < 512; nn += 32) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = 2 * nn + h - ST_WPAD;
                v[h] = (m >= 0 && m < ST_WIN) ? s_y[ST_HOP * (j - f0) + m] * s_win[m] : 0.f;
            }
            z[ZIDX(nn)] = make_float2(v[0], v[1]);
        }
        __syncwarp();
        fft512_warp(z, s_w, lane);
        const size_t row = (frow + j) * ST_NBIN;
        for (int k = lane; k <= 512; k += 32) {
            const float2 X = rfft_bin(z, s_w, k);
            const float a = sqrtf(X.x * X.x + X.y * X.y);
            const float db = 20.f * log10f(fmaxf(1e-5f, a));
            const float v = fminf(fmaxf(__fdividef(db - p.ref_db + p.max_db, p.max_db), 1e-8f), 1.f);
            if (p.spec32) p.spec32[row + k] = v;
            if (p.spec16) p.spec16[row + k] = __float2half_rn(v);
        }
        __syncwarp();
    }
}

}  // namespace zs
