// GRU backward through time on tcgen05 tensor cores (mirror of gru_cluster.cuh).
//
// One thread-block CLUSTER of NC = H/64 CTAs walks the T steps of one (direction, group of 16 sequences) in reverse.
// CTA r owns hidden units [64r, 64r+64).  Per step every CTA needs   carry[j] = sum_g W_hh[g][j] * dgh[g]   for its
// 64 units j over ALL 3H gate rows g:
//   A = W_hh^T slice [64 units (M = 64)][3H] fp16, K-major.  192*H*2 bytes do not fit next to the operand buffers,
//       so it is STREAMED: a producer warp loops over its 3*NC chunks of [64 rows][64 k] (8 KB, pre-swizzled image in
//       global memory, L2 resident) with cp.async.bulk into an 8-stage ring, independent of the recurrence;
//   B = dgh of the 16 sequences [16 rows][3H] fp16, K-major, double buffered; K index = (3*src_cta + gate)*64 + unit, so
//       the slice a CTA produces (its 64 units x 3 gates) is ONE contiguous 6 KB region that it publishes to every
//       peer with a single bulk shared->shared::cluster copy completing on the peer's mbarrier;
//   D = [64 x 16] fp32 in TMEM.
// The four gate warps (thread = one unit x 8 sequences, like the forward kernel) read D, add the direct path and the
// output gradient, apply the gate derivatives with the saved r, z, n, hn, write dgx / dgh rows for the weight-gradient
// GEMMs and the next step's B slice.
#pragma once
#include <cuda_fp16.h>

#include "gru_cluster.cuh"

namespace zs {

constexpr int GB_STAGES = 8;
constexpr int GB_CHUNK_BYTES = 64 * 128;                    // [64 rows][64 k] fp16
constexpr int GB_THREADS = 192;                             // warps 0-3 gate math, warp 4 MMA issue, warp 5 weight stream
constexpr int GB_TMEM_COLS = 32;

__host__ __device__ inline int gb_bbuf_bytes(int H) { return GRU_NSEQ * 3 * H * 2; }
__host__ __device__ inline int gb_smem_bytes(int H) { return GB_STAGES * GB_CHUNK_BYTES + 2 * gb_bbuf_bytes(H) + 1024 + 256; }
__host__ __device__ inline int gb_w_image_bytes(int H) { return 3 * (H / 64) * GB_CHUNK_BYTES; }   // per CTA = 192 * H * 2

// W_hh (3H, H) fp32 of one direction -> per-CTA streamed images [NC][3*NC chunks][64 rows j][64 k] fp16, 128-byte swizzle
__global__ void gru_pack_whhT_kernel(const float* __restrict__ W, __half* __restrict__ img, int H) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(3) * H * H) return;
    const int j = i % H, grow = i / H;
    const int gate = grow / H, unit = grow % H;
    const int src = unit >> 6, u = unit & 63;
    const int cta = j >> 6, row = j & 63;
    const int chunk = 3 * src + gate;
    const long long off = static_cast<long long>(cta) * gb_w_image_bytes(H) + static_cast<long long>(chunk) * GB_CHUNK_BYTES + row * 128 +
                          ((((u >> 3) ^ (row & 7)) << 4) | ((u & 7) << 1));
    *reinterpret_cast<__half*>(reinterpret_cast<char*>(img) + off) = __float2half_rn(W[i]);
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct GruBpttParams {
    const void* w_img;        // [2 dirs][NC][3*NC chunks][8 KB]
    const __half* gates;      // [B][T][2][4][H] r, z, n, hn
    const __half* hbuf; int h_rows, h_pitch, h_choff;       // h_t at hbuf[b][t][h_choff + dir*H + j]
    const __half* dout; int do_rows, do_pitch, do_choff;    // dL/dh_t (loss-scaled)
    __half* dgx;              // [B][T][2][3H]
    __half* dgh;              // [B][T][2][3H]
    int B, T, H;
};

__global__ void __launch_bounds__(GB_THREADS, 1) gru_bptt_cluster_kernel(const GruBpttParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    const int H = p.H, NC = H >> 6, NCH = 3 * NC;            // chunks of 64 along K = 3H
    uint8_t* sA = smem;                                       // ring of GB_STAGES chunks
    uint8_t* sB = smem + GB_STAGES * GB_CHUNK_BYTES;          // 2 x [NCH chunks][16 rows][128 B]
    const int bbuf = gb_bbuf_bytes(H);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * bbuf);
    uint64_t* full = bars;                    // [GB_STAGES]
    uint64_t* empty = bars + GB_STAGES;       // [GB_STAGES]
    uint64_t* b_full = bars + 2 * GB_STAGES;  // [2]
    uint64_t* mma_done = b_full + 2;          // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cl = cluster_id_x();
    const int n_groups = (p.B + GRU_NSEQ - 1) / GRU_NSEQ;
    const int dir = cl / n_groups, b0 = (cl % n_groups) * GRU_NSEQ;
    const int T = p.T;

    if (threadIdx.x == 0) {
        for (int i = 0; i < GB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&b_full[0], 2);
        mbar_init(&b_full[1], 2);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc<GB_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();

    const uint32_t slice_bytes = 3 * GRU_NSEQ * 128;          // my 64 units x 3 gates: 3 chunks of [16 rows][128 B]
    const uint32_t peer_tx = (NC - 1) * slice_bytes;
    // the matmul of step s (s = 0 .. T-2) turns dgh of step s into the carry of step s + 1; buffer parity = s & 1

    if (warp == 5) {
        // ------------------------------ W_hh^T stream (independent of the recurrence) ------------------
        const uint8_t* src = static_cast<const uint8_t*>(p.w_img) + (static_cast<size_t>(dir) * NC + rank) * gb_w_image_bytes(H);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s + 1 < T; ++s) {
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[stage], GB_CHUNK_BYTES);
                    bulk_g2s(sA + stage * GB_CHUNK_BYTES, src + static_cast<size_t>(c) * GB_CHUNK_BYTES, GB_CHUNK_BYTES, &full[stage]);
                }
                __syncwarp();
                if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 4) {
        // ------------------------------ MMA issue ------------------------------------------------------
        const uint32_t idesc = umma_idesc_f16_m(0, 64, GRU_NSEQ);
        const uint32_t a0 = smem_u32(sA), bb = smem_u32(sB);
        if (elect_one()) {                                     // arm the first use of each buffer
            if (T > 1) { if (NC > 1) mbar_expect_tx(&b_full[0], peer_tx); else mbar_arrive(&b_full[0]); }
            if (T > 2) { if (NC > 1) mbar_expect_tx(&b_full[1], peer_tx); else mbar_arrive(&b_full[1]); }
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s + 1 < T; ++s) {
            const int pb = s & 1;
            mbar_wait(&b_full[pb], (s >> 1) & 1);
            if (s + 2 < T - 1 + 0 && elect_one()) {            // re-arm this buffer for step s + 2 (if that step publishes)
                if (NC > 1) mbar_expect_tx(&b_full[pb], peer_tx); else mbar_arrive(&b_full[pb]);
            }
            __syncwarp();
            tc_fence_after();
            const uint32_t bcur = bb + pb * bbuf;
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_sw128(a0 + stage * GB_CHUNK_BYTES);
                    const uint64_t db = umma_desc_sw128(bcur + c * (GRU_NSEQ * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (c | k) != 0);
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(mma_done);
            __syncwarp();
        }
    } else {
        // ------------------------------ gate math (warps 0..3) -----------------------------------------
        const int q = warp, l = lane & 15, hi = lane >> 4;
        const int u_loc = 16 * q + l;
        const int unit = rank * GRU_UNITS + u_loc;
        const int seq0 = b0 + 8 * hi;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
        float keep[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) keep[i] = 0.f;
        // saved forward values of the step after next are prefetched raw (their latency must stay off the critical path)
        __half gr[8], gz[8], gn[8], ghn[8], ghp[8], gdo[8], pr[8], pz[8], pn[8], phn[8], php[8], pdo[8];
        auto load_step = [&](int s, __half (&xr)[8], __half (&xz)[8], __half (&xn)[8], __half (&xhn)[8], __half (&xhp)[8], __half (&xdo)[8]) {
            // step s of the backward walk handles forward step T-1-s of this direction
            const int fs = T - 1 - s;                                  // forward step index
            const int t = dir ? T - 1 - fs : fs;                       // time index
            const int tp = dir ? t + 1 : t - 1;                        // time index of h_prev
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = seq0 + i;
                if (b < p.B && s < T) {
                    const __half* g = p.gates + ((static_cast<size_t>(b) * T + t) * 2 + dir) * 4 * H + unit;
                    xr[i] = g[0]; xz[i] = g[H]; xn[i] = g[2 * H]; xhn[i] = g[3 * H];
                    xhp[i] = fs > 0 ? p.hbuf[(static_cast<size_t>(b) * p.h_rows + tp) * p.h_pitch + p.h_choff + dir * H + unit] : __float2half_rn(0.f);
                    xdo[i] = p.dout[(static_cast<size_t>(b) * p.do_rows + t) * p.do_pitch + p.do_choff + dir * H + unit];
                } else {
                    xr[i] = xz[i] = xn[i] = xhn[i] = xhp[i] = xdo[i] = __float2half_rn(0.f);
                }
            }
        };
        load_step(0, pr, pz, pn, phn, php, pdo);
        for (int s = 0; s < T; ++s) {
            const int fs = T - 1 - s;
            const int t = dir ? T - 1 - fs : fs;
            const int pb = s & 1;                                       // B buffer this step's dgh goes into
#pragma unroll
            for (int i = 0; i < 8; ++i) { gr[i] = pr[i]; gz[i] = pz[i]; gn[i] = pn[i]; ghn[i] = phn[i]; ghp[i] = php[i]; gdo[i] = pdo[i]; }
            load_step(s + 1, pr, pz, pn, phn, php, pdo);
            float acc[8];
            if (s > 0) {
                mbar_wait(mma_done, (s - 1) & 1);
                tc_fence_after();
                uint32_t d[16];
                tmem_ld16(t_addr, d);            // lanes 0-15 of the quadrant: D row 16q + l, 16 sequences
                tmem_ld_wait();
                tc_fence_before();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, d[8 + i], 16);   // upper lanes take columns 8..15 of row l
                    acc[i] = __uint_as_float(hi ? got : d[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
            uint8_t* bnext = sB + pb * bbuf + (3 * rank) * (GRU_NSEQ * 128);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = __half2float(gr[i]), z = __half2float(gz[i]), n = __half2float(gn[i]), hn = __half2float(ghn[i]);
                const float dh = keep[i] + acc[i] + __half2float(gdo[i]);
                const float dn = dh * (1.f - z), dz = dh * (__half2float(ghp[i]) - n);
                keep[i] = dh * z;
                const float an = dn * (1.f - n * n), az = dz * z * (1.f - z), ar = an * hn * r * (1.f - r), anr = an * r;
                const __half har = __float2half_rn(ar), haz = __float2half_rn(az), han = __float2half_rn(an), hanr = __float2half_rn(anr);
                const int b = seq0 + i;
                if (b < p.B) {
                    const size_t o = ((static_cast<size_t>(b) * T + t) * 2 + dir) * 3 * H + unit;
                    p.dgx[o] = har; p.dgx[o + H] = haz; p.dgx[o + 2 * H] = han;
                    p.dgh[o] = har; p.dgh[o + H] = haz; p.dgh[o + 2 * H] = hanr;
                }
                const int row = 8 * hi + i;
                const int sw = row * 128 + ((((u_loc >> 3) ^ (row & 7)) << 4) | ((u_loc & 7) << 1));
                *reinterpret_cast<__half*>(bnext + sw) = har;
                *reinterpret_cast<__half*>(bnext + GRU_NSEQ * 128 + sw) = haz;
                *reinterpret_cast<__half*>(bnext + 2 * GRU_NSEQ * 128 + sw) = hanr;
            }
            if (s + 1 < T) {
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (elect_one()) {
                    const uint32_t src = smem_u32(bnext);
                    const uint32_t bar_local = smem_u32(&b_full[pb]);
                    for (uint32_t dd = 1 + warp; dd < static_cast<uint32_t>(NC); dd += 4) {
                        const uint32_t peer = (rank + dd) % NC;
                        dsmem_bulk_copy(mapa_shared(src, peer), src, slice_bytes, mapa_shared(bar_local, peer));
                    }
                    if (warp == 0) mbar_arrive(&b_full[pb]);
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 4) tmem_dealloc<GB_TMEM_COLS>(tmem_base);
}

}  // namespace zs
