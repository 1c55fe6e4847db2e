// Test: tcgen05.mma with the A operand in TMEM (kind::f16, M=128): layout check + issue cost.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../zerospeech-tts-without-t_b200/csrc/ptx.cuh"
using namespace zs;

__host__ __device__ inline uint32_t idesc_mn(int m, int n) {
    uint32_t d = 0; d |= 1u << 4; d |= (uint32_t)(n >> 3) << 17; d |= (uint32_t)(m >> 4) << 24; return d;
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

constexpr int K = 64, N = 64;   // one 64-wide K chunk (4 MMAs of K=16)

// A[m][k] = (m % 7) - 3 + (k % 5); B[n][k] = (n % 3) - 1 + (k % 4)  (exact in fp16)
__global__ void __launch_bounds__(128, 1) test(float* D_out, long long* cyc, int iters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sB = smem_raw + (base - smem_u32(smem_raw));
    // B tile [N rows][64 k] K-major, 128B swizzle
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        int n = i / K, k = i % K;
        int off = n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
        *reinterpret_cast<__half*>(sB + off) = __float2half((float)((n % 3) - 1 + (k % 4)));
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
    // A in TMEM at columns [256, 256 + K/2): lane m, column c holds (A[m][2c], A[m][2c+1]) low half = even k
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) {
            int k = 2 * (c0 + j);
            __half2 h = __floats2half2_rn((float)((m % 7) - 3 + (k % 5)), (float)((m % 7) - 3 + ((k + 1) % 5)));
            v[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st8(tm + (uint32_t(32 * warp) << 16) + 256 + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) {
        const uint32_t id = idesc_mn(128, N);
        const uint64_t db = umma_desc_sw128(base);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ts(tm, tm + 256 + k * 8, db + 2 * k, id, (k | it) != 0 ? 1u : 0u);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
    }
    __syncthreads(); tc_fence_after();
    // read D (first N columns), one row per thread
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tm + (uint32_t(32 * warp) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D_out[m * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
    float* d; long long* c;
    cudaMalloc(&d, 128 * N * 4); cudaMalloc(&c, 16);
    cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int iters : {1, 64}) {
        test<<<1, 128, 64 * 1024>>>(d, c, iters);
        cudaError_t e = cudaDeviceSynchronize();
        static float h[128 * N]; long long hc[2];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += ((m % 7) - 3 + (k % 5)) * ((n % 3) - 1 + (k % 4));
            ref *= iters;
            double err = fabs(h[m * N + n] - ref);
            if (err > 1e-3 * (1 + fabs(ref))) ++bad;
            if (err > maxerr) maxerr = err;
        }
        printf("iters=%d: %s, mismatches %d / %d (max err %.3g), D[0][0]=%g D[5][3]=%g; issue %.1f cyc/mma, complete %.1f cyc/mma\n", iters,
               cudaGetErrorString(e), bad, 128 * N, maxerr, h[0], h[5 * N + 3], hc[0] / (4.0 * iters), hc[1] / (4.0 * iters));
    }
    return 0;
}
