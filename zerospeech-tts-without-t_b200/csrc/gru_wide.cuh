// Bidirectional GRU recurrence, throughput shape: 64 sequences per cluster, the r|z slice of W_hh resident in TMEM.
//
// Same decomposition as gru_cluster.cuh (cluster of NC = H/64 CTAs per (direction, sequence group), CTA r owns hidden units
// [64r, 64r+64), per-step state exchange through L2 with one bulk store + one multicast bulk load, `consumed` handshake by a
// multicast tcgen05.commit).  What changes is where W_hh lives: its 128 r|z rows are the MMA's A operand FROM TENSOR MEMORY
// (lane = row, column c = the fp16 pair (k = 2c, 2c+1); tcgen05.mma with [a_tmem] runs at the same 33 cycles per MMA as a
// shared-memory A - tools/ts_mma_test.cu), H/2 of the 512 columns; only the 64 n-gate rows (64 KB at H = 512) stay in shared
// memory.  That frees the shared memory for the state of 64 sequences (64 KB), so the ~2 100 cycles of MMAs and the L2 round
// trip of the exchange are paid once per 64 sequence-steps instead of once per 32.  The 8 gate warps walk the 64 columns in two
// passes of 32 (warp w: TMEM quadrant w & 3, 16 columns per pass; thread: one unit x 8 sequences per pass, fp32 state in shared
// memory).  Inference only (no gate save), tanh.approx gates.
//
// Measured and rejected (round 2): running the two 32-sequence halves as independent, software-pipelined recurrences (own
// accumulators / barriers / 4 KB exchanges, no block barrier).  With ONE control warp walking the groups in order the decoder GRU
// went from 1.17 to 1.68 ms per 960 segments (waiting for group 0's late slices blocks the issue of group 1's ready MMAs); with an
// MMA-issue warp and an exchange warp PER GROUP (20 warps, 96 registers, bit-identical results), staggered or not, it measures
// 1.15 - 1.20 ms: exactly the one-group kernel.  The step is a latency chain (MMAs -> gate math -> exchange through L2), two groups
// have the same chain as one, and throughput only grows with more independent groups per cluster than a B200 can hold at 960
// segments (128 sequences per cluster are 16 clusters against the 15 resident ones).
#pragma once
#include "gru_cluster.cuh"

namespace zs {

constexpr int GRU_WIDE_NSEQ = 64;       // sequences per cluster with 2 gate passes (NPASS = 4: 128)

__host__ __device__ inline int gru_wide_wn_bytes(int H) { return 64 * H * 2; }               // n-gate rows, swizzled image
__host__ __device__ inline int gru_wide_state_bytes(int H, int npass) { return 32 * npass * H * 2; }
__host__ __device__ inline int gru_wide_smem_bytes(int H, int npass) {
    return gru_wide_wn_bytes(H) + gru_wide_state_bytes(H, npass) + npass * 8 * 256 * 4 /* fp32 state of every pass */ + 1024 + 128;
}

// W_hh (3H, H) fp32 of one direction -> plain r|z rows [NC][128 rows][H] operand type, rows ordered like the TMEM lanes of the
// r|z tile: lane 32q + 16*gate + l = gate (0 r, 1 z) of unit 16q + l of the CTA's 64.
template <typename OT>
__global__ void gru_pack_wrz_plain_kernel(const float* __restrict__ W, OT* __restrict__ out, int H) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(2) * H * H) return;
    const int k = i % H, grow = i / H;           // grow = gate * H + unit, gate in {0, 1}
    const int gate = grow / H, unit = grow % H;
    const int cta = unit / GRU_UNITS, u = unit % GRU_UNITS;
    const int row = 32 * (u >> 4) + 16 * gate + (u & 15);
    out[(static_cast<long long>(cta) * 128 + row) * H + k] = float_to_ot<OT>(W[i]);
}

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

struct GruWideParams {
    const void* wrz;        // [2 dirs][NC][128][H] operand type (gru_pack_wrz_plain_kernel)
    const void* w_img;      // [2 dirs][NC][192 * H] swizzled images of gru_pack_whh_kernel: only the n part (last 64 rows) is used
    const float* bhh;       // [2][3H]
    const void* gx;         // [B][T][2][3H] operand type
    void* out;              // operand type [B][out_rows][out_pitch]
    uint8_t* xchg;          // [clusters][NC][64 x 128 B] exchange scratch
    int B, T, H, out_rows, out_pitch, out_halo, out_choff, fmt;
};

// GW gate warps (8 or 16): with 16, a 32-sequence pass belongs to ONE set of eight warps (warp w: TMEM quadrant w & 3, 16-column half
// (w >> 2) & 1, passes [PPW (w >> 3), +PPW)), so the passes of a step run side by side instead of back to back - the gate math is
// latency-bound at two warps per scheduler.
template <typename OT, int NPASS, int GW>
__global__ void __launch_bounds__(32 * (GW + 1), 1) gru_wide_kernel(const GruWideParams p) {
    constexpr int NTHR = 32 * (GW + 1);
    constexpr int PPW = NPASS * 8 / GW;            // passes per gate warp
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    constexpr int NSEQ = 32 * NPASS;
    const int H = p.H, KCH = H >> 6, NC = KCH;
    uint8_t* sWn = smem;                                        // [KCH chunks][64 rows][128 B]
    uint8_t* sH = smem + gru_wide_wn_bytes(H);                  // [KCH chunks][NSEQ rows][128 B]
    float* sState = reinterpret_cast<float*>(sH + gru_wide_state_bytes(H, NPASS));     // [NPASS][8][256 threads]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sState) + NPASS * 8 * 256 * 4);
    uint64_t* h_chunk = bars;         // [8]
    uint64_t* rz_done = bars + 8;
    uint64_t* mma_done = bars + 9;
    uint64_t* consumed = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cl = cluster_id_x();
    const int n_groups = (p.B + NSEQ - 1) / NSEQ;
    const int dir = cl / n_groups, b0 = (cl % n_groups) * NSEQ;

    // ---- one-time setup ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.w_img) +
                                                          (static_cast<size_t>(dir) * KCH + rank) * gru_w_image_bytes(H) + 128 * H * 2);
        uint4* dst = reinterpret_cast<uint4*>(sWn);
        const int n16 = gru_wide_wn_bytes(H) / 16;
        for (int i = threadIdx.x; i < n16; i += NTHR) dst[i] = src[i];
        uint4* hz = reinterpret_cast<uint4*>(sH);
        for (int i = threadIdx.x; i < (gru_wide_state_bytes(H, NPASS) + NPASS * 8 * 256 * 4) / 16; i += NTHR) hz[i] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 8; ++c) mbar_init(&h_chunk[c], 1);
        mbar_init(rz_done, 1);
        mbar_init(mma_done, 1);
        mbar_init(consumed, NC);
        fence_barrier_init();
    }
    if (warp == GW) tmem_alloc<512>(tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_wrz = tmem_base;                          // columns [0, H/2): the r|z rows of W_hh
    const uint32_t t_drz = tmem_base + (H >> 1);               // NSEQ columns
    const uint32_t t_dn = t_drz + NSEQ;                        // NSEQ columns (H/2 + 2 NSEQ <= 512)
    if (warp < 8) {
        // W_rz -> TMEM: this warp's quadrant (lane = row), its half of the K columns
        const int q = warp & 3, part = warp >> 2;
        const int row = 32 * q + lane;
        const OT* wr = reinterpret_cast<const OT*>(p.wrz) + ((static_cast<size_t>(dir) * NC + rank) * 128 + row) * H;
        const int c_lo = part * (H >> 2), c_hi = c_lo + (H >> 2);        // 32-bit columns: 2 fp16 each
        for (int c = c_lo; c < c_hi; c += 8) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(wr + 2 * c), v1 = *reinterpret_cast<const uint4*>(wr + 2 * c + 8);
            const uint32_t v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            tmem_st8(t_wrz + (static_cast<uint32_t>(32 * q) << 16) + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();

    const uint32_t slice_bytes = NSEQ * 128;                   // [NSEQ rows][128 B]

    if (warp == GW) {
        // ------------------------------ control / MMA issue ------------------------------
        const uint32_t idesc_rz = umma_idesc_f16_m(p.fmt, 128, NSEQ);
        const uint32_t idesc_n = umma_idesc_f16_m(p.fmt, 64, NSEQ);
        const uint32_t w_n = smem_u32(sWn), hb = smem_u32(sH);
        const uint16_t all_ctas = static_cast<uint16_t>((1u << NC) - 1u);
        if (p.T > 1 && elect_one())
            for (int c = 0; c < NC; ++c)
                if (c != static_cast<int>(rank)) mbar_expect_tx(&h_chunk[c], slice_bytes);
        __syncwarp();
        for (int t = 0; t < p.T; ++t) {
            for (int j = 0; j < KCH; ++j) {
                const int c = (static_cast<int>(rank) + KCH - j) % KCH;
                if (t > 0) {
                    mbar_wait(&h_chunk[c], (t - 1) & 1);
                    if (c != static_cast<int>(rank) && t + 1 < p.T && elect_one()) mbar_expect_tx(&h_chunk[c], slice_bytes);
                    __syncwarp();
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t db = umma_desc_sw128(hb + c * slice_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ts(t_drz, t_wrz + c * 32 + k * 8, db + 2 * k, idesc_rz, (j | k) != 0);
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(rz_done);
                uint64_t dn = umma_desc_sw128(w_n), db = umma_desc_sw128(hb);
                for (int c = 0; c < KCH; ++c) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(t_dn, dn + 2 * k, db + 2 * k, idesc_n, (c | k) != 0);
                    dn += (64 * 128) >> 4;
                    db += slice_bytes >> 4;
                }
                umma_commit(mma_done);
                if (t + 1 < p.T) umma_commit_multicast(consumed, all_ctas);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------ gate math (warps 0..7), two passes of 32 columns ------------------------------
        const int q = warp & 3, half = (warp >> 2) & 1, pbase = (warp >> 3) * PPW;
        const int l = lane & 15, hi = lane >> 4;
        const int u_loc = 16 * q + l;
        const int unit = rank * GRU_UNITS + u_loc;
        const float* bh = p.bhh + static_cast<size_t>(dir) * 3 * H;
        const float b_r = bh[unit], b_z = bh[H + unit], b_n = bh[2 * H + unit];
        const float hb_r = 0.5f * b_r, hb_z = 0.5f * b_z;
        OT* out = reinterpret_cast<OT*>(p.out);
        const uint32_t lane_addr = static_cast<uint32_t>(32 * q) << 16;
        float* my_state = sState + (threadIdx.x & 255);        // [pass][i] at (pass * 8 + i) * 256
        const OT* gx_base = reinterpret_cast<const OT*>(p.gx);
        const int gx_step = (dir ? -1 : 1) * 2 * 3 * H;
        const int out_step = (dir ? -1 : 1) * p.out_pitch;
        const int t_first = dir ? p.T - 1 : 0;
        const int gx_seq = p.T * 6 * H, out_seq = p.out_rows * p.out_pitch;
        // rows of the state tile / sequences of this thread: pass ps -> rows 32 ps + 16 half + 8 hi + i
        int row0[PPW], gx_off[PPW], out_off[PPW];
        uint32_t live[PPW];
#pragma unroll
        for (int ps = 0; ps < PPW; ++ps) {
            row0[ps] = 32 * (pbase + ps) + 16 * half + 8 * hi;
            const int seq0 = b0 + row0[ps];
            live[ps] = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (seq0 + i < p.B) live[ps] |= 1u << i;
            gx_off[ps] = ((seq0 * p.T + t_first) * 2 + dir) * 3 * H + unit;
            out_off[ps] = (seq0 * p.out_rows + p.out_halo + t_first) * p.out_pitch + p.out_choff + dir * H + unit;
        }
        OT gr[8], gz[8], gn[8], pr[8], pz[8], pn[8];
        auto load_gx = [&](int ps, bool ok, OT (&xr)[8], OT (&xz)[8], OT (&xn)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (ok && (live[ps] >> i & 1)) {
                    const OT* g = gx_base + (gx_off[ps] + i * gx_seq);
                    xr[i] = g[0]; xz[i] = g[H]; xn[i] = g[2 * H];
                } else {
                    xr[i] = xz[i] = xn[i] = float_to_ot<OT>(0.f);
                }
            }
            gx_off[ps] += gx_step;
        };
        load_gx(0, true, pr, pz, pn);
        uint8_t* hnext = sH + rank * slice_bytes;
        for (int t = 0; t < p.T; ++t) {
#pragma unroll
            for (int ps = 0; ps < PPW; ++ps) {
                const int gp = pbase + ps;                      // the pass (32-sequence column block) this iteration works on
#pragma unroll
                for (int i = 0; i < 8; ++i) { gr[i] = pr[i]; gz[i] = pz[i]; gn[i] = pn[i]; }
                // prefetch the next pass: of this step, or this warp's first pass of the next one
                if (ps + 1 < PPW) load_gx(ps + 1, true, pr, pz, pn);
                else load_gx(0, t + 1 < p.T, pr, pz, pn);
                if (ps == 0) {
                    mbar_wait(rz_done, t & 1);
                    tc_fence_after();
                }
                float r[8], z[8];
                {
                    uint32_t a[16];
                    tmem_ld16(t_drz + lane_addr + 32 * gp + 16 * half, a);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t got = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[8 + i], 16);
                        const float hr = __uint_as_float(hi ? got : a[i]);
                        const float hz = __uint_as_float(hi ? a[8 + i] : got);
                        // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5 with x = gx + W h + b: one mixed-precision add, two FMAs, one MUFU
                        // (the same expression as gru_cluster.cuh: the kernels stay bit-identical to each other)
                        r[i] = fmaf(0.5f, tanh_mufu(fmaf(0.5f, res_add<OT>(ot_bits(gr[i]), hr), hb_r)), 0.5f);
                        z[i] = fmaf(0.5f, tanh_mufu(fmaf(0.5f, res_add<OT>(ot_bits(gz[i]), hz), hb_z)), 0.5f);
                    }
                }
                if (ps == 0) {
                    mbar_wait(mma_done, t & 1);
                    tc_fence_after();
                }
                OT y[8];
                {
                    uint32_t nn[16];
                    tmem_ld16(t_dn + lane_addr + 32 * gp + 16 * half, nn);
                    tmem_ld_wait();
                    if (ps == PPW - 1) tc_fence_before();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t gotn = __shfl_xor_sync(0xffffffffu, nn[8 + i], 16);
                        const float hn = __uint_as_float(hi ? gotn : nn[i]) + b_n;
                        const float n = tanh_mufu(fmaf(r[i], hn, ot_to_float<OT>(gn[i])));
                        const float hp = my_state[(gp * 8 + i) * 256];
                        const float hv = fmaf(z[i], hp - n, n);
                        my_state[(gp * 8 + i) * 256] = hv;
                        y[i] = float_to_ot<OT>(hv);
                    }
                }
                if (t + 1 < p.T) {
                    if (ps == 0) mbar_wait(consumed, t & 1);       // every CTA's MMAs of this step have read the state
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int s = row0[ps] + i;
                        *reinterpret_cast<OT*>(hnext + s * 128 + ((((u_loc >> 3) ^ (s & 7)) << 4) | ((u_loc & 7) << 1))) = y[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (live[ps] >> i & 1) out[out_off[ps] + i * out_seq] = y[i];
                out_off[ps] += out_step;
            }
            if (t + 1 < p.T) {
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * GW) : "memory");
                if (warp == 0 && elect_one()) {
                    mbar_arrive(&h_chunk[rank]);
                    if (NC > 1) {
                        const uint32_t src = smem_u32(hnext);
                        uint8_t* g = p.xchg + (static_cast<size_t>(cl) * KCH + rank) * slice_bytes;
                        bulk_store_smem_to_global(g, src, slice_bytes);
                        const uint16_t peers = static_cast<uint16_t>(((1u << NC) - 1u) & ~(1u << rank));
                        bulk_load_multicast(src, g, slice_bytes, smem_u32(&h_chunk[rank]), peers);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == GW) tmem_dealloc<512>(tmem_base);
}

}  // namespace zs
