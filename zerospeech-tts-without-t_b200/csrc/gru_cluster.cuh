// Bidirectional GRU recurrence (model/model.py:59-66, zero initial state) on tcgen05 tensor cores.
//
// One thread-block CLUSTER of NC = H/64 CTAs runs the T sequential steps of one (direction, group of 32
// sequences).  CTA r owns hidden units [64r, 64r+64): its 192 rows of W_hh (r|z|n gates of those units, all
// K = H columns, fp16/bf16) stay resident in shared memory for the whole kernel as the A operand; the hidden
// state of the 32 sequences is the B operand ([32 rows][H], K-major, 128-byte swizzled) in ONE buffer (W_hh
// leaves room for no more at H = 512).
// Per step:  D_rz[128 x 32] = W_rz h^T (M = 128),  D_n[64 x 32] = W_n h^T (M = 64)  -> TMEM;
// the 8 gate warps (two per TMEM lane quadrant, 16 sequences each) read D, add the precomputed input
// projections gx (from the GEMM kernel) and b_hh, apply the gates, keep h in fp32 registers, write h_t to the
// output buffer and publish their 64-unit slice of the new state to every CTA of the cluster: one bulk store to an
// L2-resident scratch and one MULTICAST bulk load back into the same shared-memory offset of all peers, completing on
// each peer's per-source mbarrier (the direct alternative - one shared->shared::cluster bulk copy per peer - is bound by
// the ~17 B/clk SM-to-SM port: 28 KB out + 28 KB in per CTA and step; through L2 the CTA sends 4 KB).  No cluster-wide
// barrier inside the time loop.  The single state buffer may only be overwritten once EVERY CTA's MMAs of the step have read
// it: each CTA's tcgen05.commit is multicast to a `consumed` mbarrier in all CTAs of the cluster, and the gate
// warps wait for it (normally long complete - the gate math sits in between) before they store or send h_{t+1}.
//
// Row order of the A tiles is chosen so a unit's three gate pre-activations land in the same warp:
//   tile RZ (M = 128): TMEM lane 32q + l  = r-gate of unit 16q + l (l < 16), z-gate of unit 16q + l - 16 (l >= 16)
//   tile N  (M = 64):  TMEM lane 32q + l  = n-gate of unit 16q + l (l < 16)      [M = 64 uses 16 lanes per quadrant]
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace zs {

constexpr int GRU_NSEQ = 16;        // sequences per cluster of the BPTT kernel (gru_bptt_cluster.cuh)
constexpr int GRU_FWD_NSEQ = 32;    // sequences per cluster of the forward recurrence (UMMA N): throughput shape
constexpr int GRU_FWD_NSEQ_SMALL = 16;   // latency shape for batches whose 16-sequence clusters fit one wave
constexpr int GRU_UNITS = 64;       // hidden units per CTA
constexpr int GRU_GATE_WARPS = 8;   // warp w: TMEM quadrant w & 3, sequences (NSEQ/2) * (w >> 2) .. + NSEQ/2
constexpr int GRU_THREADS = 32 * (GRU_GATE_WARPS + 1);    // + warp 8: MMA issue / control
constexpr int GRU_TMEM_COLS = 64;

__host__ __device__ inline int gru_w_image_bytes(int H) { return 192 * H * 2; }
__host__ __device__ inline int gru_smem_bytes(int H, int nseq = GRU_FWD_NSEQ) {
    return gru_w_image_bytes(H) + nseq * H * 2 + 1024 /*align*/ + 128 /*barriers*/;
}

// byte offset of element (row, k) inside a K-major tile of `rows` rows stored as K/64 consecutive chunks of
// [rows][64] elements, each row 128 bytes with the 128-byte swizzle (16-byte unit index XOR (row % 8))
__host__ __device__ inline int sw128_offset(int rows, int row, int k) {
    const int chunk = k >> 6, kk = k & 63;
    return chunk * rows * 128 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1));
}

// W_hh (3H, H) fp32 of one direction -> per-CTA shared-memory images [NC][192 * H] operand type:
// tile RZ (128 rows) followed by tile N (64 rows), rows ordered as described above.
template <typename OT>
__global__ void gru_pack_whh_kernel(const float* __restrict__ W, OT* __restrict__ img, int H) {
    const int NC = H / GRU_UNITS;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(3) * H * H) return;
    const int k = i % H, grow = i / H;          // grow: row of W_hh = gate * H + unit
    const int gate = grow / H, unit = grow % H;
    const int cta = unit / GRU_UNITS, u = unit % GRU_UNITS;
    const int q = u >> 4, l = u & 15;
    int off;
    if (gate < 2) off = sw128_offset(128, 32 * q + 16 * gate + l, k);
    else off = 128 * H * 2 + sw128_offset(64, u, k);
    (void)NC;
    OT* dst = reinterpret_cast<OT*>(reinterpret_cast<char*>(img) + static_cast<long long>(cta) * gru_w_image_bytes(H) + off);
    *dst = float_to_ot<OT>(W[i]);
}

// local shared -> peer CTA shared, completion (bytes) signalled on the PEER's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                uint32_t mbar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dst_cluster_addr),
        "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
        : "memory");
}
// The same slice to EVERY peer through L2: one bulk store shared -> global scratch, then one multicast bulk load
// global -> the same shared-memory offset of every CTA in `cta_mask`, completing (bytes) on each destination's mbarrier.
// Per step a CTA then moves 4 KB out + 28 KB in through the L2 path instead of 28 KB out + 28 KB in through the
// SM-to-SM port (measured ~17 B/clk per SM, the bound of the direct exchange).
__device__ __forceinline__ void bulk_store_smem_to_global(void* gdst, uint32_t src_cta_addr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src_cta_addr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // the write is complete (not only the source read)
}
__device__ __forceinline__ void bulk_load_multicast(uint32_t dst_cta_addr, const void* gsrc, uint32_t bytes, uint32_t mbar_cta_addr,
                                                    uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            dst_cta_addr),
        "l"(gsrc), "r"(bytes), "r"(mbar_cta_addr), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline uint32_t umma_idesc_f16_m(int fmt, int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= static_cast<uint32_t>(fmt) << 7;
    d |= static_cast<uint32_t>(fmt) << 10;
    d |= static_cast<uint32_t>(n >> 3) << 17;
    d |= static_cast<uint32_t>(m >> 4) << 24;
    return d;
}

struct GruParams {
    const void* w_img;      // [2 dirs][NC][192 * H] operand type (gru_pack_whh_kernel)
    const float* bhh;       // [2][3H]
    const void* gx;         // [B][T][2][3H] operand type, b_ih (+ speaker term) folded in
    void* out;              // operand type [B][out_rows][out_pitch]
    int B, T, H, out_rows, out_pitch, out_halo, out_choff, fmt, debug;
    long long* dbg;        // debug bit 3: per-step clock64 stamps of cluster 0 / CTA 0 ([T][8])
    int fast_act;          // 1: sigmoid/tanh through tanh.approx.f32 (one MUFU each, 2^-11 relative error)
    uint8_t* xchg;         // non-null: exchange the state slices through this L2-resident scratch ([clusters][NC][slice]) by multicast
    void* gates;           // training: r, z, n, hn = W_hn h + b_hn per step, operand type [B][T][2][4][H]; null = not saved
};

__device__ __forceinline__ uint16_t ot_bits(__half v) { return __half_as_ushort(v); }
__device__ __forceinline__ uint16_t ot_bits(__nv_bfloat16 v) { return __bfloat16_as_ushort(v); }
__device__ __forceinline__ float tanh_mufu(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

template <typename OT, int NSEQ>
__global__ void __launch_bounds__(GRU_THREADS, 1) gru_cluster_kernel(const GruParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    constexpr int CPW = NSEQ / 2;      // TMEM columns (sequences) per gate warp
    constexpr int SPT = NSEQ / 4;      // sequences per gate thread
    const int H = p.H, KCH = H >> 6;
    const int NC = (ZS_DBG(p) & 1) ? 1 : KCH;      // debug bit 0: pretend to be alone (no exchange, no peer waits)
    uint8_t* sW = smem;                                   // [192 * H * 2]
    uint8_t* sH = smem + gru_w_image_bytes(H);            // [KCH chunks][32 rows][128 B]
    const int hbuf_bytes = NSEQ * H * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sH + hbuf_bytes);
    uint64_t* h_chunk = bars;         // [8] chunk c of the next state (the 64 units of CTA c) is in place
    uint64_t* rz_done = bars + 8;     // my r|z-tile MMAs of this step are complete
    uint64_t* mma_done = bars + 9;    // all my MMAs of this step are complete
    uint64_t* consumed = bars + 10;   // all MMAs of this step are complete in EVERY CTA of the cluster
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cl = cluster_id_x();
    const int n_groups = (p.B + NSEQ - 1) / NSEQ;
    const int dir = cl / n_groups, b0 = (cl % n_groups) * NSEQ;

    // ---- one-time setup: W_hh slice -> smem, zero state, barriers, TMEM ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.w_img) +
                                                          (static_cast<size_t>(dir) * KCH + rank) * gru_w_image_bytes(H));
        uint4* dst = reinterpret_cast<uint4*>(sW);
        const int n16 = gru_w_image_bytes(H) / 16;
        for (int i = threadIdx.x; i < n16; i += GRU_THREADS) dst[i] = src[i];
        uint4* hz = reinterpret_cast<uint4*>(sH);
        for (int i = threadIdx.x; i < hbuf_bytes / 16; i += GRU_THREADS) hz[i] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 8; ++c) mbar_init(&h_chunk[c], 1);
        mbar_init(rz_done, 1);
        mbar_init(mma_done, 1);
        mbar_init(consumed, NC);
        fence_barrier_init();
    }
    if (warp == GRU_GATE_WARPS) tmem_alloc<GRU_TMEM_COLS>(tmem_slot);
    fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();                // every CTA's barriers are initialised before any peer signals them

    const uint32_t slice_bytes = NSEQ * 128;              // one 64-unit chunk of the state: [32 rows][128 B]

    if (warp == GRU_GATE_WARPS) {
        // ------------------------------ control / MMA issue (whole warp, elected lane issues) ----
        const uint32_t idesc_rz = umma_idesc_f16_m(p.fmt, 128, NSEQ);
        const uint32_t idesc_n = umma_idesc_f16_m(p.fmt, 64, NSEQ);
        const uint32_t w_rz = smem_u32(sW), w_n = smem_u32(sW) + 128 * H * 2, hb = smem_u32(sH);
        const uint16_t all_ctas = NC > 1 ? static_cast<uint16_t>((1u << NC) - 1u) : static_cast<uint16_t>(1u << rank);
        if (p.T > 1 && elect_one())                      // arm the arrival of the peers' slices of h_1
            for (int c = 0; c < NC; ++c)
                if (c != static_cast<int>(rank)) mbar_expect_tx(&h_chunk[c], slice_bytes);
        __syncwarp();
        for (int t = 0; t < p.T; ++t) {
            const bool rec = (ZS_DBG(p) & 8) && p.dbg && blockIdx.x == 0 && lane == 0;
            // r|z tile first, chunk by chunk in the order the slices land (mine, then the peers by ring distance):
            // the MMAs of the early chunks run while the late ones are still in flight
            if ((ZS_DBG(p) & 32) && t > 0 && NC > 1)            // experiment: no MMA before every slice has landed
                for (int c = 0; c < KCH; ++c) mbar_wait(&h_chunk[c], (t - 1) & 1);
            for (int j = 0; j < KCH; ++j) {
                const int c = (static_cast<int>(rank) + KCH - j) % KCH;
                if (t > 0 && (NC > 1 || j == 0)) {
                    mbar_wait(&h_chunk[c], (t - 1) & 1);
                    // re-arm for h_{t+1}: no peer sends it before it has seen my commit of this step
                    if (c != static_cast<int>(rank) && t + 1 < p.T && elect_one()) mbar_expect_tx(&h_chunk[c], slice_bytes);
                    __syncwarp();
                }
                tc_fence_after();
                if (rec && j == 0) p.dbg[t * 8 + 0] = clock64();
                if (elect_one()) {
                    const uint64_t da = umma_desc_sw128(w_rz + c * (128 * 128)), db = umma_desc_sw128(hb + c * slice_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc_rz, (j | k) != 0);
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(rz_done);
                uint64_t dn = umma_desc_sw128(w_n), db = umma_desc_sw128(hb);
                for (int c = 0; c < KCH; ++c) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_base + NSEQ, dn + 2 * k, db + 2 * k, idesc_n, (c | k) != 0);
                    dn += (64 * 128) >> 4;       // next 64-wide K chunk of each operand (address field is >> 4)
                    db += slice_bytes >> 4;
                }
                umma_commit(mma_done);
                if (t + 1 < p.T) umma_commit_multicast(consumed, all_ctas);
            }
            __syncwarp();
            if (rec) p.dbg[t * 8 + 1] = clock64();
        }
    } else {
        // ------------------------------ gate math (warps 0..7) ------------------------------
        const int q = warp & 3, half = warp >> 2;
        const int l = lane & 15, hi = lane >> 4;               // hi: which half of this warp's CPW sequences
        const int u_loc = 16 * q + l;                          // unit within this CTA's 64
        const int unit = rank * GRU_UNITS + u_loc;
        const float* bh = p.bhh + static_cast<size_t>(dir) * 3 * H;
        const float b_r = bh[unit], b_z = bh[H + unit], b_n = bh[2 * H + unit];
        const bool fast = p.fast_act != 0;
        float h[SPT];
#pragma unroll
        for (int i = 0; i < SPT; ++i) h[i] = 0.f;
        OT* out = reinterpret_cast<OT*>(p.out);
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + CPW * half;
        const int row0 = CPW * half + SPT * hi;                   // first row of the state tile this thread writes
        const int seq0 = b0 + row0;
        // input projections are prefetched one full step ahead (their HBM latency would otherwise sit on the
        // critical path of every step); they stay RAW in registers - converting them here would wait for the load.
        // Per-sequence element offsets are kept in registers and stepped by a constant (no 64-bit address
        // arithmetic inside the time loop); sequences past the batch are masked once.
        OT gr[SPT], gz[SPT], gn[SPT], pr[SPT], pz[SPT], pn[SPT];
        const OT* gx_base = reinterpret_cast<const OT*>(p.gx);
        const int gx_step = (dir ? -1 : 1) * 2 * 3 * H;              // elements between consecutive steps of a sequence
        const int out_step = (dir ? -1 : 1) * p.out_pitch;
        const int t_first = dir ? p.T - 1 : 0;
        uint32_t live = 0;
#pragma unroll
        for (int i = 0; i < SPT; ++i)
            if (seq0 + i < p.B) live |= 1u << i;
        // element offsets of this thread's FIRST sequence at the current step (the launcher checks the 32-bit range);
        // its other sequences follow at a constant stride
        const int gx_seq = p.T * 6 * H, out_seq = p.out_rows * p.out_pitch;
        int gx_off = ((seq0 * p.T + t_first) * 2 + dir) * 3 * H + unit;
        int out_off = (seq0 * p.out_rows + p.out_halo + t_first) * p.out_pitch + p.out_choff + dir * H + unit;
        const bool no_gx = (ZS_DBG(p) & 2) != 0;
        auto load_gx = [&](int step, OT (&xr)[SPT], OT (&xz)[SPT], OT (&xn)[SPT]) {      // `step` must be the NEXT unread step
            const bool ok = step < p.T && !no_gx;
#pragma unroll
            for (int i = 0; i < SPT; ++i) {
                if (ok && (live >> i & 1)) {
                    const OT* g = gx_base + (gx_off + i * gx_seq);
                    xr[i] = g[0]; xz[i] = g[H]; xn[i] = g[2 * H];
                } else {
                    xr[i] = xz[i] = xn[i] = float_to_ot<OT>(0.f);
                }
            }
            gx_off += gx_step;
        };
        load_gx(0, pr, pz, pn);
        for (int t = 0; t < p.T; ++t) {
            const int tt = dir ? p.T - 1 - t : t;
#pragma unroll
            for (int i = 0; i < SPT; ++i) { gr[i] = pr[i]; gz[i] = pz[i]; gn[i] = pn[i]; }
            load_gx(t + 1, pr, pz, pn);
            const bool rec = (ZS_DBG(p) & 8) && p.dbg && blockIdx.x == 0 && threadIdx.x == 0;
            if (rec) p.dbg[t * 8 + 2] = clock64();
            // ---- r and z while the n-tile MMAs still run ----
            mbar_wait(rz_done, t & 1);
            tc_fence_after();
            if (rec) p.dbg[t * 8 + 3] = clock64();
            float r[SPT], z[SPT];
            {
                uint32_t a[CPW];
                tmem_ld_cols(t_addr, a);        // lanes 0-15: W_hr h, lanes 16-31: W_hz h   (this warp's 16 sequences)
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    const uint32_t got = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[SPT + i], 16);
                    const float hr = __uint_as_float(hi ? got : a[i]);
                    const float hz = __uint_as_float(hi ? a[SPT + i] : got);
                    if (fast) {
                        // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5 with x = gx + W h + b: one mixed-precision add, two FMAs, one MUFU
                        // (the same expression as gru_wide.cuh: the kernels stay bit-identical to each other)
                        r[i] = fmaf(0.5f, tanh_mufu(fmaf(0.5f, res_add<OT>(ot_bits(gr[i]), hr), 0.5f * b_r)), 0.5f);
                        z[i] = fmaf(0.5f, tanh_mufu(fmaf(0.5f, res_add<OT>(ot_bits(gz[i]), hz), 0.5f * b_z)), 0.5f);
                    } else {
                        const float xr = ot_to_float<OT>(gr[i]) + hr + b_r, xz = ot_to_float<OT>(gz[i]) + hz + b_z;
                        r[i] = sigmoid_f(xr);
                        z[i] = sigmoid_f(xz);
                    }
                }
            }
            if (rec) p.dbg[t * 8 + 4] = clock64();
            // ---- n and the new state ----
            mbar_wait(mma_done, t & 1);
            tc_fence_after();
            OT y[SPT];
            {
                uint32_t nn[CPW];
                tmem_ld_cols(t_addr + NSEQ, nn);    // lanes 0-15: W_hn h
                tmem_ld_wait();
                tc_fence_before();
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    const uint32_t gotn = __shfl_xor_sync(0xffffffffu, nn[SPT + i], 16);
                    const float hn = __uint_as_float(hi ? gotn : nn[i]) + b_n;
                    const float xn = fmaf(r[i], hn, ot_to_float<OT>(gn[i]));
                    const float n = fast ? tanh_mufu(xn) : tanh_f(xn);
                    h[i] = fmaf(z[i], h[i] - n, n);          // (1 - z) n + z h
                    y[i] = float_to_ot<OT>(h[i]);
                    const int b = seq0 + i;
                    if (p.gates != nullptr && b < p.B) {
                        OT* g = reinterpret_cast<OT*>(p.gates) + ((static_cast<size_t>(b) * p.T + tt) * 2 + dir) * 4 * H + unit;
                        g[0] = float_to_ot<OT>(r[i]); g[H] = float_to_ot<OT>(z[i]); g[2 * H] = float_to_ot<OT>(n); g[3 * H] = float_to_ot<OT>(hn);
                    }
                }
            }
            if (rec) p.dbg[t * 8 + 5] = clock64();
            if ((ZS_DBG(p) & 64) && !(ZS_DBG(p) & 4)) {           // experiment: output stores before the exchange
#pragma unroll
                for (int i = 0; i < SPT; ++i)
                    if (live >> i & 1) out[out_off + i * out_seq] = y[i];
            }
            if (t + 1 < p.T) {
                // every CTA's MMAs of this step have read the state (and with them my previous outgoing copies have
                // long landed): the buffer may now take h_{t+1}
                mbar_wait(consumed, t & 1);
                uint8_t* hnext = sH + rank * slice_bytes;
#pragma unroll
                for (int i = 0; i < SPT; ++i) {
                    const int s = row0 + i;      // row of the state tile
                    *reinterpret_cast<OT*>(hnext + s * 128 + ((((u_loc >> 3) ^ (s & 7)) << 4) | ((u_loc & 7) << 1))) = y[i];
                }
                fence_proxy_async_smem();                       // my slice -> visible to the bulk-copy engine / MMA
                asm volatile("bar.sync 1, %0;" ::"n"(32 * GRU_GATE_WARPS) : "memory");   // the gate warps only
                if (rec) p.dbg[t * 8 + 6] = clock64();
                if (elect_one()) {      // each gate warp publishes the slice to its share of the peers
                    const uint32_t src = smem_u32(hnext);
                    const uint32_t bar_local = smem_u32(&h_chunk[rank]);
                    if (warp == 0) mbar_arrive(&h_chunk[rank]);   // my own slice is in place
                    if (p.xchg != nullptr) {
                        if (warp == 0 && NC > 1) {
                            uint8_t* g = p.xchg + (static_cast<size_t>(cl) * KCH + rank) * slice_bytes;
                            bulk_store_smem_to_global(g, src, slice_bytes);
                            const uint16_t peers = static_cast<uint16_t>(((1u << NC) - 1u) & ~(1u << rank));
                            bulk_load_multicast(src, g, slice_bytes, bar_local, peers);
                        }
                    } else {
                        for (uint32_t d = 1 + warp; d < static_cast<uint32_t>(NC); d += GRU_GATE_WARPS) {
                            const uint32_t peer = (rank + d) % NC;
                            dsmem_bulk_copy(mapa_shared(src, peer), src, slice_bytes, mapa_shared(bar_local, peer));
                        }
                    }
                }
                __syncwarp();
                if (rec) p.dbg[t * 8 + 7] = clock64();
            }
            // h_t to the output buffer, off the exchange's critical path
            if (!(ZS_DBG(p) & (4 | 64))) {
#pragma unroll
                for (int i = 0; i < SPT; ++i)
                    if (live >> i & 1) out[out_off + i * out_seq] = y[i];
            }
            out_off += out_step;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                // nobody leaves while a peer may still be reading its slices
    if (warp == GRU_GATE_WARPS) tmem_dealloc<GRU_TMEM_COLS>(tmem_base);
}

}  // namespace zs
