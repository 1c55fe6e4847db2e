// Test: tcgen05.mma kind::f16 with BOTH operands MN-major (128-byte swizzle, as TMA writes a {64 channels, K rows}
// box) and MIXED formats (A = bf16, B = fp16).  This is the operand form of the weight-gradient GEMM
//   dW[co][ci] = sum_n dY[n][co] * X[n][ci]      (n = (segment, frame) is the reduction dimension and the OUTER
// dimension of both channels-last buffers).
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../zerospeech-tts-without-t_b200/csrc/ptx.cuh"
using namespace zs;

constexpr int M = 128, N = 128, K = 64;

// MN-major SW128 descriptor: LBO = byte distance between 64-element MN atoms, SBO = distance between 8-row K groups
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ inline uint32_t idesc_mixed(int afmt, int bfmt, int amn, int bmn, int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= static_cast<uint32_t>(afmt) << 7;
    d |= static_cast<uint32_t>(bfmt) << 10;
    d |= static_cast<uint32_t>(amn) << 15;
    d |= static_cast<uint32_t>(bmn) << 16;
    d |= static_cast<uint32_t>(n >> 3) << 17;
    d |= static_cast<uint32_t>(m >> 4) << 24;
    return d;
}
__host__ __device__ inline float aval(int m, int k) { return static_cast<float>((m % 7) - 3 + (k % 5)); }
__host__ __device__ inline float bval(int n, int k) { return static_cast<float>((n % 3) - 1 + (k % 4)) * 0.5f; }

__global__ void __launch_bounds__(128, 1) test(float* D_out, int afmt, int bfmt, int amn, int bmn) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sA = smem_raw + (base - smem_u32(smem_raw));   // [M/64 blocks][K rows][128 B]
    uint8_t* sB = sA + (M / 64) * K * 128;                  // [N/64 blocks][K rows][128 B]
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
        const int m = i % M, k = i / M;
        const int off = amn ? (m / 64) * K * 128 + k * 128 + (((((m % 64) >> 3) ^ (k & 7)) << 4) | ((m & 7) << 1))
                            : m * 128 + ((((k >> 3) ^ (m & 7)) << 4) | ((k & 7) << 1));
        if (afmt) *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(aval(m, k));
        else *reinterpret_cast<__half*>(sA + off) = __float2half(aval(m, k));
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i % N, k = i / N;
        const int off = bmn ? (n / 64) * K * 128 + k * 128 + (((((n % 64) >> 3) ^ (k & 7)) << 4) | ((n & 7) << 1))
                            : n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
        if (bfmt) *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(bval(n, k));
        else *reinterpret_cast<__half*>(sB + off) = __float2half(bval(n, k));
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5, m = threadIdx.x;
    if (warp == 0) {
        const uint32_t id = idesc_mixed(afmt, bfmt, amn, bmn, M, N);
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < K / 16; ++k) {
                // 16 K rows further = 2048 B
                const uint32_t bB = base + (M / 64) * K * 128;
                const uint64_t da = amn ? desc_mn_sw128(base + k * 2048, K * 128) : umma_desc_sw128(base) + 2 * k;
                const uint64_t db = bmn ? desc_mn_sw128(bB + k * 2048, K * 128) : umma_desc_sw128(bB) + 2 * k;
                umma_f16(tm, da, db, id, k != 0);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    __syncthreads(); tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tm + (uint32_t(32 * warp) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D_out[m * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<256>(tm);
}

int main(int argc, char** argv) {
    float* d;
    cudaMalloc(&d, M * N * 4);
    cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int afmt = atoi(argv[1]), bfmt = atoi(argv[2]), amn = atoi(argv[3]), bmn = atoi(argv[4]);
    test<<<1, 128, 64 * 1024>>>(d, afmt, bfmt, amn, bmn);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[M * N];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += aval(m, k) * bval(n, k);
        const double err = fabs(h[m * N + n] - ref);
        if (err > 1e-3 * (1 + fabs(ref))) ++bad;
        if (err > maxerr) maxerr = err;
    }
    printf("afmt %d bfmt %d a_mn %d b_mn %d: %s, mismatches %d / %d (max err %.3g), D[0][0]=%g D[5][3]=%g D[127][127]=%g\n",
           afmt, bfmt, amn, bmn, cudaGetErrorString(e), bad, M * N, maxerr, h[0], h[5 * N + 3], h[127 * N + 127]);
    return bad != 0;
}
