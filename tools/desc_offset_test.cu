// Probe: may a K-major, 128-byte-swizzled UMMA operand start at an arbitrary 128-byte ROW of a 1024-byte-aligned tile, and which
// "matrix base offset" (descriptor bits 49-51) does it need?  The B tile holds 160 rows x 64 fp16 written in the layout TMA
// SWIZZLE_128B produces (16-byte chunk c of row r at chunk c ^ (r & 7)); A selects one k per output row, so
//   D[m][n] = B[n + r0][m % 64]     for a descriptor that starts r0 rows into the tile.
// Prints, for r0 = 0..9 and base_offset in {0, r0 & 7}, how many of the 128 x 128 outputs match.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o desc_offset_test tools/desc_offset_test.cu && ./desc_offset_test
#include <cstdio>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../zerospeech-tts-without-t_b200/csrc/ptx.cuh"
using namespace zs;

constexpr int ROWS_B = 160, N = 128;

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t addr, uint32_t base_off) {
    return umma_desc_sw128(addr) | (static_cast<uint64_t>(base_off & 7) << 49);
}

__global__ void __launch_bounds__(128, 1) test(float* D_out, int r0, int use_bo) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sA = smem_raw + (base - smem_u32(smem_raw));       // [128][64] K-major SW128 (16 KB)
    uint8_t* sB = sA + 16384;                                   // [160][64]
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int r = i / 64, k = i % 64;
        const float v = (k == (r % 64)) ? 1.f : 0.f;
        *reinterpret_cast<__half*>(sA + r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1))) = __float2half(v);
    }
    for (int i = threadIdx.x; i < ROWS_B * 64; i += blockDim.x) {
        const int r = i / 64, k = i % 64;
        const float v = static_cast<float>((r % 32) * 64 + k);  // < 2048: exact in fp16
        *reinterpret_cast<__half*>(sB + r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1))) = __float2half(v);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<128>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t id = umma_idesc_f16(0, N);
            const uint32_t b_start = base + 16384 + r0 * 128;
            const uint64_t da = umma_desc_sw128(base), db = desc_sw128_bo(b_start, use_bo ? (r0 & 7) : 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tm, da + 2 * k, db + 2 * k, id, k != 0 ? 1u : 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tm + (uint32_t(32 * warp) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D_out[threadIdx.x * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<128>(tm);
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * N * 4);
    cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    static float h[128 * N];
    for (int r0 = 0; r0 < 10; ++r0)
        for (int bo = 0; bo < 2; ++bo) {
            cudaMemset(d, 0xff, sizeof(h));
            test<<<1, 128, 64 * 1024>>>(d, r0, bo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("r0=%d bo=%d: %s\n", r0, bo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            int ok = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) ok += h[m * N + n] == static_cast<float>(((n + r0) % 32) * 64 + (m % 64));
            printf("start row %d, base_offset %d: %5d / %d match%s\n", r0, bo ? (r0 & 7) : 0, ok, 128 * N, ok == 128 * N ? "  <- exact" : "");
        }
    return 0;
}
