// Weight-gradient GEMM of a conv1d / per-frame-linear layer on tcgen05 tensor cores.
//
//   dW[co][ci][j] += scale * sum_{b, t} dY[b][t][co] * X[b][row0 + stride*t + j][ci]
//
// Both operands are channels-last activation buffers, so the reduction index n = (segment, frame) is their OUTER
// dimension: they enter the MMA as MN-major tiles (instruction-descriptor bits 15/16), which is exactly what a TMA
// box of {64 channels, 64 rows} with the 128-byte swizzle leaves in shared memory - no transposed copies of the
// activations are ever made.  A tap is a row offset of the X box, a stride-2 conv reads X through the same
// (channel, row parity, row pair, segment) view the forward kernel uses.
//
// Work item = (128 INPUT channels = the M rows of the MMA) x (<= 256 output channels) x (one tap) x (one K split);
// one CTA per SM walks the items.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue: fp32
// accumulators leave TMEM through red.global.add.f32 into the gradient in the reference's (C_out, C_in, k) layout.
// The input channel is the TMEM lane on purpose: the 32 lanes of an epilogue warp then add to 32 CONSECUTIVE
// input channels of one output channel (stride k floats - one or a few cache lines per warp instruction) instead
// of 32 different output-channel rows (32 lines per instruction).
#pragma once
#include <cuda_fp16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace zs {

constexpr int WG_STAGES = 4;
constexpr int WG_KROWS = 64;                               // reduction rows per pipeline stage
constexpr int WG_BLK_BYTES = WG_KROWS * 128;               // one {64 channels, 64 rows} box
constexpr int WG_A_BYTES = 2 * WG_BLK_BYTES;               // 128 input channels (X)
constexpr int WG_B_BYTES = 4 * WG_BLK_BYTES;               // up to 256 output channels (dY)
constexpr int WG_SMEM_BYTES = WG_STAGES * (WG_A_BYTES + WG_B_BYTES) + 1024 + 256;
constexpr int WG_THREADS = 192;

struct alignas(64) WgradParams {
    CUtensorMap tmA;      // dY: (channel, row, segment), box {64, rows_ps, nb}
    CUtensorMap tmB;      // X: stride 1 (channel, row, segment) box {64, rows_ps, nb}; stride 2 (channel, parity, pair, segment) box {64, 1, rows_ps, nb}
    float* grad;          // (c_out, c_in_total, k) fp32
    int m_tiles, n_tiles, taps, ksplit;
    int n_blk_last;       // 64-channel blocks of the LAST n (output-channel) tile (others have 4)
    int c_in, c_in_total, ci_off, c_out, k, tap0;   // grad index ((co * c_in_total + ci_off + ci) * k + tap0 + tap)
    int a_ch0, a_row0;    // A: first channel / buffer row of frame 0
    int b_ch0, b_row0;    // B: first channel / buffer row read by tap 0 of frame 0
    int stride;
    int rows_ps, nb, seg_steps, n_groups;   // stage = nb segments x rows_ps rows (= 64); seg_steps stages per segment group
    int ps_c;             // > 0: A channel m = r * ps_c + c is conv output channel 2c + r (pixel-shuffled layer)
    float scale;
    int direct;           // 1: ksplit == 1 and the gradient is zero on entry - store instead of red.add
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(WG_BLK_BYTES >> 4) << 16;     // LBO: next 64-channel block
    d |= static_cast<uint64_t>(1024 >> 4) << 32;             // SBO: next group of 8 reduction rows
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ inline uint32_t umma_idesc_f16_mn(int n) {   // fp16 x fp16 -> fp32, both operands MN-major, M = 128
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 1u << 15;
    d |= 1u << 16;
    d |= static_cast<uint32_t>(n >> 3) << 17;
    d |= static_cast<uint32_t>(128 >> 4) << 24;
    return d;
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + WG_STAGES * WG_A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * (WG_A_BYTES + WG_B_BYTES));
    uint64_t* empty = full + WG_STAGES;
    uint64_t* tfull = empty + WG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total = p.m_tiles * p.n_tiles * p.taps * p.ksplit;
    const int k_stages = p.n_groups * p.seg_steps;             // pipeline stages of the whole reduction
    // item -> (mt fastest, then nt, tap, ks): neighbouring CTAs share the X boxes of an (nt, tap, ks)
    auto decode = [&](int item, int& mt, int& nt, int& tap, int& s0, int& s1) {
        mt = item % p.m_tiles; item /= p.m_tiles;
        nt = item % p.n_tiles; item /= p.n_tiles;
        tap = item % p.taps;
        const int ks = item / p.taps;
        s0 = static_cast<int>(static_cast<long long>(k_stages) * ks / p.ksplit);
        s1 = static_cast<int>(static_cast<long long>(k_stages) * (ks + 1) / p.ksplit);
    };

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            int mt, nt, tap, s0, s1;
            decode(item, mt, nt, tap, s0, s1);
            const int n_blk = nt == p.n_tiles - 1 ? p.n_blk_last : 4;
            const uint32_t tx = static_cast<uint32_t>(2 + n_blk) * WG_BLK_BYTES;
            for (int s = s0; s < s1; ++s) {
                const int g = s / p.seg_steps, sub = s % p.seg_steps;
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], tx);
                    uint8_t* a = sA + stage * WG_A_BYTES;
                    uint8_t* bb = sB + stage * WG_B_BYTES;
                    for (int j = 0; j < 2; ++j) {               // A: 128 input channels of X, shifted by the tap
                        const int ch = p.b_ch0 + mt * 128 + j * 64;
                        if (p.stride == 2) {
                            const int r = p.b_row0 + tap;       // buffer row of frame 0; frame t is row r + 2t
                            tma_load_4d(&p.tmB, a + j * WG_BLK_BYTES, &full[stage], ch, r & 1, (r >> 1) + sub * p.rows_ps, g * p.nb);
                        } else {
                            tma_load_3d(&p.tmB, a + j * WG_BLK_BYTES, &full[stage], ch, p.b_row0 + tap + sub * p.rows_ps, g * p.nb);
                        }
                    }
                    const int a_row = p.a_row0 + sub * p.rows_ps;
                    for (int j = 0; j < n_blk; ++j)             // B: up to 256 output channels of dY
                        tma_load_3d(&p.tmA, bb + j * WG_BLK_BYTES, &full[stage], p.a_ch0 + nt * 256 + j * 64, a_row, g * p.nb);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0, it = 0;
        uint32_t phase = 0;
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
            int mt, nt, tap, s0, s1;
            decode(item, mt, nt, tap, s0, s1);
            const int n_blk = nt == p.n_tiles - 1 ? p.n_blk_last : 4;
            const uint32_t idesc = umma_idesc_f16_mn(n_blk * 64);
            const int as = it & 1;
            mbar_wait(&tempty[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * MAX_BN;
            for (int s = s0; s < s1; ++s) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < WG_KROWS / 16; ++k) {     // 16 reduction rows = 2048 B further
                        const uint64_t da = umma_desc_mn_sw128(a0 + stage * WG_A_BYTES + k * 2048);
                        const uint64_t db = umma_desc_mn_sw128(b0 + stage * WG_B_BYTES + k * 2048);
                        umma_f16(d_tmem, da, db, idesc, (s > s0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull[as]);
            __syncwarp();
        }
    } else {
        const int quad = warp & 3, row = quad * 32 + lane;
        int it = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
            int mt, nt, tap, s0, s1;
            decode(item, mt, nt, tap, s0, s1);
            const int n_blk = nt == p.n_tiles - 1 ? p.n_blk_last : 4;
            const int as = it & 1;
            mbar_wait(&tfull[as], (it >> 1) & 1);
            tc_fence_after();
            const int ci = mt * 128 + row;                         // input channel (relative to b_ch0) = TMEM lane
            const bool ci_ok = ci < p.c_in;
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * MAX_BN;
            float* gr = p.grad + (static_cast<size_t>(p.ci_off) + (ci_ok ? ci : 0)) * p.k + p.tap0 + tap;
            const size_t co_stride = static_cast<size_t>(p.c_in_total) * p.k;
            if (s1 > s0) {
                for (int c0 = 0; c0 < n_blk * 64; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + c0, v);
                    tmem_ld_wait();
                    const int m0 = nt * 256 + c0;                  // channel of the dY buffer (relative to a_ch0)
                    if (ci_ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int m = m0 + i;
                            if (m < p.c_out) {
                                const int co = p.ps_c > 0 ? 2 * (m % p.ps_c) + m / p.ps_c : m;
                                if (p.direct) gr[co * co_stride] = __uint_as_float(v[i]) * p.scale;
                                else red_add_f32(gr + co * co_stride, __uint_as_float(v[i]) * p.scale);
                            }
                        }
                    }
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
