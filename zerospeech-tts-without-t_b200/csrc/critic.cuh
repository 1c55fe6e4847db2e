// 2-D critic / classifier forward (model/model.py:113-226: PatchDiscriminator, TargetClassifier) on the 1-D conv GEMM.
//
// A 5x5 stride-2 Conv2d over (H = frequency, W = time) is run as a 1-D stride-2 convolution along W whose "channels" are
// the five kernel rows: out[b, :, ho, wo] = sum_kw sum_(kh, ci) W[:, ci, kh, kw] * Xpad[b, ci, 2 ho + kh, 2 wo + kw].
// conv2d_gather_kernel lays the rows 2 ho + kh - 2 (reflected, model/model.py:31-38) of the channels-last activation next to
// each other - segment = (b, ho), frame = padded w, channel = kh * C + ci, a 5x expansion instead of im2col's 25x - and
// applies the PREVIOUS layer's InstanceNorm2d on the way (its statistics span every segment of a sample, so they cannot
// be fused into the GEMM epilogue the way the 1-D InstanceNorm is).  conv_gemm_kernel (taps = 5, stride 2, leaky-relu
// epilogue) does the rest; instnorm2d_stats_kernel reduces its output per (sample, channel).
#pragma once

struct Gather2dParams {
    const void* src;          // fp32 (B, H, W) [layer 1: one input channel] or fp16 channels-last [(b*H + h)*W + w][src_pitch]
    int B, H, W, C, src_pitch;
    int KH, stride_h, pad_h, pad_w, Ho, Wp;
    const longlong2* stats;   // [B][C] (sum, sum of squares) x 2^20 of the source over its H*W positions (fixed point), or nullptr
    float inv_count;
    __half* out;              // [(b*Ho + ho)][Wp][out_pitch], channel kh*C + ci
    int out_pitch;
};

// InstanceNorm2d statistics travel as 2^20-scaled 64-bit integers: the row splits combine with integer atomics, whose sum does not
// depend on the order the CTAs arrive in (fp32 atomics made the forward pass differ from run to run in the fourth digit)
constexpr float STATS_FIX = 1048576.f;

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// layer 1: C = 1, out_pitch = 8 -> one 16-byte store per (segment, padded frame)
__global__ void conv2d_gather_f32_kernel(Gather2dParams p) {
    const long long n = static_cast<long long>(p.B) * p.Ho * p.Wp;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int wp = static_cast<int>(i % p.Wp);
    const long long seg = i / p.Wp;
    const int ho = static_cast<int>(seg % p.Ho), b = static_cast<int>(seg / p.Ho);
    const int w = reflect_idx(wp - p.pad_w, p.W);
    const float* x = static_cast<const float*>(p.src) + static_cast<long long>(b) * p.H * p.W;
    __align__(16) __half v[8];
#pragma unroll
    for (int kh = 0; kh < 8; ++kh) {
        float f = 0.f;
        if (kh < p.KH) f = x[static_cast<long long>(reflect_idx(p.stride_h * ho + kh - p.pad_h, p.H)) * p.W + w];
        v[kh] = __float2half_rn(f);
    }
    *reinterpret_cast<uint4*>(p.out + i * 8) = *reinterpret_cast<const uint4*>(v);
}

// C a multiple of 8: one thread moves 8 channels of one (segment, padded frame, kernel row)
__global__ void conv2d_gather_f16_kernel(Gather2dParams p) {
    const int c8n = p.C / 8;
    const long long n = static_cast<long long>(p.B) * p.Ho * p.Wp * p.KH * c8n;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const int c8 = static_cast<int>(i % c8n);
    long long r = i / c8n;
    const int kh = static_cast<int>(r % p.KH); r /= p.KH;
    const int wp = static_cast<int>(r % p.Wp);
    const long long seg = r / p.Wp;
    const int ho = static_cast<int>(seg % p.Ho), b = static_cast<int>(seg / p.Ho);
    const int h = reflect_idx(p.stride_h * ho + kh - p.pad_h, p.H), w = reflect_idx(wp - p.pad_w, p.W);
    const __half* s = static_cast<const __half*>(p.src) + ((static_cast<long long>(b) * p.H + h) * p.W + w) * p.src_pitch + c8 * 8;
    uint4 raw = *reinterpret_cast<const uint4*>(s);
    if (p.stats) {
        __half2* h2 = reinterpret_cast<__half2*>(&raw);
        const longlong2* st = p.stats + static_cast<long long>(b) * p.C + c8 * 8;
        const float sc = p.inv_count * (1.f / STATS_FIX);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 v = __half22float2(h2[j]);
            const longlong2 s0 = st[2 * j], s1 = st[2 * j + 1];
            const float m0 = static_cast<float>(s0.x) * sc, m1 = static_cast<float>(s1.x) * sc;
            const float r0 = rsqrtf(fmaxf(static_cast<float>(s0.y) * sc - m0 * m0, 0.f) + 1e-5f);      // nn.InstanceNorm2d: biased variance, eps 1e-5
            const float r1 = rsqrtf(fmaxf(static_cast<float>(s1.y) * sc - m1 * m1, 0.f) + 1e-5f);
            v.x = (v.x - m0) * r0; v.y = (v.y - m1) * r1;
            h2[j] = __floats2half2_rn(v.x, v.y);
        }
    }
    __half* o = p.out + (seg * p.Wp + wp) * p.out_pitch + kh * p.C + c8 * 8;
    *reinterpret_cast<uint4*>(o) = raw;
}

// stats[b][c] += 2^20 x (sum, sum of squares) over the P rows of sample b; grid (C / 64, B, row splits), 256 threads = 4 row lanes x 64 channels
__global__ void instnorm2d_stats_kernel(const __half* y, long long P, int C, int pitch, longlong2* stats) {
    __shared__ float2 red[4][64];
    const int cl = threadIdx.x & 63, rl = threadIdx.x >> 6, c = blockIdx.x * 64 + cl, b = blockIdx.y;
    const long long per = (P + gridDim.z - 1) / gridDim.z, r0 = blockIdx.z * per, r1 = r0 + per < P ? r0 + per : P;
    float s = 0.f, ss = 0.f;
    if (c < C) {
        const __half* base = y + static_cast<long long>(b) * P * pitch + c;
        for (long long r = r0 + rl; r < r1; r += 4) {
            const float v = __half2float(base[r * pitch]);
            s += v; ss += v * v;
        }
    }
    red[rl][cl] = make_float2(s, ss);
    __syncthreads();
    if (rl == 0 && c < C) {
        for (int k = 1; k < 4; ++k) { s += red[k][cl].x; ss += red[k][cl].y; }
        unsigned long long* o = reinterpret_cast<unsigned long long*>(stats + static_cast<long long>(b) * C + c);
        atomicAdd(o, static_cast<unsigned long long>(__float2ll_rn(s * STATS_FIX)));          // two's complement: negative sums wrap correctly
        atomicAdd(o + 1, static_cast<unsigned long long>(__float2ll_rn(ss * STATS_FIX)));
    }
}

// conv7 / conv_classify (kernel = the whole 17 x W/32 map, model/model.py:123-131): out[b][j] = bias[j] + <y[b], w[j]> over P*C values
__global__ void critic_head_kernel(const __half* y, int P, int C, int pitch, const float* w, const float* bias, int J, float* out) {
    __shared__ float red[8];
    const int j = blockIdx.x, b = blockIdx.y, n = P * C;
    const __half* yb = y + static_cast<long long>(b) * P * pitch;
    const float* wj = w + static_cast<long long>(j) * n;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __half2float(yb[(i / C) * pitch + i % C]) * wj[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = bias[j];
        for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) t += red[k];
        out[static_cast<long long>(b) * J + j] = t;
    }
}

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" int zs_conv2d_gather(const void* src, int src_is_f32, int B, int H, int W, int C, int src_pitch, int KH, int stride_h, int pad_h,
                                int pad_w, int Ho, int Wp, const void* stats, float inv_count, void* out, int out_pitch, void* stream) {
    ZS_TRY(ensure_device());
    if (!src || !out) return fail(ZS_ERR_ARG, "conv2d_gather: null pointer");
    if (B < 1 || H < 1 || W < 1 || C < 1 || KH < 1 || KH > 8 || stride_h < 1 || Ho < 1 || Wp < 1)
        return fail(ZS_ERR_ARG, "conv2d_gather: bad shape (B %d, H %d, W %d, C %d, KH %d, stride %d, Ho %d, Wp %d)", B, H, W, C, KH, stride_h, Ho, Wp);
    if (pad_h >= H || pad_w >= W) return fail(ZS_ERR_ARG, "conv2d_gather: reflect padding (%d, %d) needs a larger map than %d x %d", pad_h, pad_w, H, W);
    if (stride_h * (Ho - 1) + KH - 1 - pad_h > 2 * (H - 1) || Wp - 1 - pad_w > 2 * (W - 1)) return fail(ZS_ERR_ARG, "conv2d_gather: window leaves the reflected map");
    if (reinterpret_cast<uintptr_t>(out) % 16 || reinterpret_cast<uintptr_t>(src) % 16) return fail(ZS_ERR_ARG, "conv2d_gather: pointers must be 16-byte aligned");
    Gather2dParams p;
    p.src = src; p.B = B; p.H = H; p.W = W; p.C = C; p.src_pitch = src_pitch; p.KH = KH; p.stride_h = stride_h; p.pad_h = pad_h; p.pad_w = pad_w;
    p.Ho = Ho; p.Wp = Wp; p.stats = reinterpret_cast<const longlong2*>(stats); p.inv_count = inv_count; p.out = static_cast<__half*>(out); p.out_pitch = out_pitch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "conv2d_gather_kernel");
    if (src_is_f32) {
        if (C != 1 || out_pitch != 8 || stats) return fail(ZS_ERR_ARG, "conv2d_gather: the fp32 source is the single-channel network input (out_pitch 8, no statistics)");
        const long long n = static_cast<long long>(B) * Ho * Wp;
        conv2d_gather_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p);
    } else {
        if (C % 8 || src_pitch % 8 || out_pitch % 8 || out_pitch < KH * C) return fail(ZS_ERR_ARG, "conv2d_gather: C %d / pitches %d, %d must be multiples of 8, out_pitch >= KH * C", C, src_pitch, out_pitch);
        const long long n = static_cast<long long>(B) * Ho * Wp * KH * (C / 8);
        if ((n + 255) / 256 > 0x7fffffffLL) return fail(ZS_ERR_ARG, "conv2d_gather: batch too large");
        conv2d_gather_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

extern "C" int zs_instnorm2d_stats(const void* y, int B, long long P, int C, int pitch, void* stats, void* stream) {
    ZS_TRY(ensure_device());
    if (!y || !stats || B < 1 || P < 1 || C < 1 || pitch < C) return fail(ZS_ERR_ARG, "instnorm2d_stats: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(stats, 0, static_cast<size_t>(B) * C * sizeof(longlong2), st));
    const int splits = static_cast<int>(std::max<long long>(1, std::min<long long>(64, P / 256)));
    dim3 grid((C + 63) / 64, B, splits);
    LaunchScope scope(st, KC_OTHER, 0.0, "instnorm2d_stats_kernel");
    instnorm2d_stats_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(y), P, C, pitch, reinterpret_cast<longlong2*>(stats));
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

extern "C" int zs_critic_head(const void* y, int B, int P, int C, int pitch, const float* w, const float* bias, int J, float* out, void* stream) {
    ZS_TRY(ensure_device());
    if (!y || !w || !bias || !out || B < 1 || P < 1 || C < 1 || J < 1 || pitch < C) return fail(ZS_ERR_ARG, "critic_head: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "critic_head_kernel");
    critic_head_kernel<<<dim3(J, B), 256, 0, st>>>(static_cast<const __half*>(y), P, C, pitch, w, bias, J, out);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
