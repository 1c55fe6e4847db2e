// Probe: tcgen05.mma.cta_group::2 (CTA pair) semantics on sm_100a - where the halves of A / B come from and where D lands.
//   D[m][n] = m + 256 n  (A[m][0] = m, A[m][1] = 1, B[n][0] = 1, B[n][1] = 256 n, everything else 0: exact in fp16/fp32),
// so every accumulator element names the (m, n) it holds.  Each CTA of the pair owns HALF the A rows and HALF the B rows
// in its own shared memory at identical offsets; CTA 0 issues the MMAs for both.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cta2_mma_test tools/cta2_mma_test.cu && ./cta2_mma_test
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../zerospeech-tts-without-t_b200/csrc/ptx.cuh"
using namespace zs;

__host__ __device__ inline uint32_t idesc_f16(int m, int n) {
    uint32_t d = 0; d |= 1u << 4; d |= (uint32_t)(n >> 3) << 17; d |= (uint32_t)(m >> 4) << 24; return d;
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int K = 64, N = 64;

// M_total = 256 (128 A rows per CTA) or 128 (64 A rows per CTA); each CTA holds N/2 = 32 B rows
__global__ void __launch_bounds__(128, 1) test(float* D_out, int M_total) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t rank = ctarank();
    uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sA = smem_raw + (base - smem_u32(smem_raw));     // [128 rows][64 k] K-major, 128-byte swizzle (16 KB)
    uint8_t* sB = sA + 16384;                                 // [32 rows][64 k]
    const int m_per = M_total / 2;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        const int m = rank * m_per + r;                       // global A row held by this CTA's row r
        float v = 0.f;
        if (r < m_per) v = k == 0 ? (float)m : (k == 1 ? 1.f : 0.f);
        *reinterpret_cast<__half*>(sA + r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1))) = __float2half(v);
    }
    for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        const int n = rank * 32 + r;                          // global B row held by this CTA's row r
        const float v = k == 0 ? 1.f : (k == 1 ? 256.f * n : 0.f);
        *reinterpret_cast<__half*>(sB + r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1))) = __float2half(v);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    cluster_sync();                                           // both CTAs' operands, barriers and TMEM are ready
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const uint32_t id = idesc_f16(M_total, N);
            const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_2cta(tm, da + 2 * k, db + 2 * k, id, k != 0 ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {                      // raw dump: [cta][lane 0..127][col 0..63]
        uint32_t v[16];
        tmem_ld16(tm + (uint32_t(32 * warp) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D_out[(rank * 128 + threadIdx.x) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// issue cost: `n_mma` back-to-back MMAs of the given shape on the same operands, leader-side clock from first issue to commit arrival
__global__ void __launch_bounds__(128, 1) bench2(long long* cyc, int M_total, int n_mma) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t rank = ctarank();
    uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    cluster_sync();
    const uint32_t tm = slot;
    if (rank == 0 && threadIdx.x < 32) {
        long long t0 = clock64();
        if (elect_one()) {
            const uint32_t id = idesc_f16(M_total, N);
            const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
            for (int i = 0; i < n_mma; ++i) umma_f16_2cta(tm + 64 * (i & 3), da + 2 * (i & 3), db + 2 * (i & 3), id, i >= 4 ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    } else {
        mbar_wait(&bar, 0);
    }
    tc_fence_before(); __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
    {
        long long* c; cudaMalloc(&c, 16);
        cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        for (int M_total : {256, 128}) for (int n_mma : {32, 64, 96}) {
            cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, bench2, c, M_total, n_mma);
            cudaError_t e = cudaDeviceSynchronize();
            long long hc[2]; cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
            printf("cta_group::2 M=%d N=%d: %d MMAs: issue %lld cycles, complete %lld cycles (%.1f / MMA) %s\n", M_total, N, n_mma, hc[0], hc[1], (double)hc[1] / n_mma, cudaGetErrorString(e));
        }
    }
    float* d;
    cudaMalloc(&d, 2 * 128 * N * 4);
    cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int M_total : {256, 128}) {
        cudaMemset(d, 0xff, 2 * 128 * N * 4);
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, test, d, M_total);
        cudaError_t e = cudaDeviceSynchronize();
        static float h[2 * 128 * N];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M_total=%d: launch %s, sync %s\n", M_total, cudaGetErrorString(le), cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        for (int cta = 0; cta < 2; ++cta) {
            // decode: value = m + 256 n
            int ok = 0, junk = 0;
            printf("  cta %d: lane -> m (col 0):", cta);
            for (int l = 0; l < 128; l += 8) { float v = h[(cta * 128 + l) * N]; printf(" %d:%d", l, (v == v && v >= 0 && v < 65536) ? ((int)v) % 256 : -1); }
            printf("\n  cta %d: col -> n (lane 0):", cta);
            for (int c = 0; c < N; c += 4) { float v = h[(cta * 128) * N + c]; printf(" %d:%d", c, (v == v && v >= 0 && v < 65536) ? ((int)v) / 256 : -1); }
            printf("\n");
            const int m_per = M_total / 2;
            for (int l = 0; l < 128; ++l) for (int c = 0; c < N; ++c) {
                float v = h[(cta * 128 + l) * N + c];
                if (!(v == v) || v < 0 || v > 65536) { ++junk; continue; }
                if (M_total == 256) { if ((int)v == cta * 128 + l + 256 * c) ++ok; }
                else if (l < 64 && (int)v == cta * m_per + l + 256 * c) ++ok;     // hypothesis: rows on lanes 0..63
            }
            printf("  cta %d: %d elements match the natural layout (lane = local row, col = n), %d untouched/junk\n", cta, ok, junk);
            if (M_total == 128) {
                for (int l : {0, 20, 40, 63, 64, 84, 104, 127}) {
                    printf("    lane %3d (m,n) at cols 0,9,18,..:", l);
                    for (int c = 0; c < N; c += 9) { int v = (int)h[(cta * 128 + l) * N + c]; printf(" (%d,%d)", v % 256, v / 256); }
                    printf("\n");
                }
            }
        }
    }
    return 0;
}
