// Microbenchmark: cycles per tcgen05.mma (kind::f16, SS) as a function of (M, N), back-to-back issue by one thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zerospeech-tts-without-t_b200/csrc/ptx.cuh"
using namespace zs;

__host__ __device__ inline uint32_t idesc_mn(int m, int n) {
    uint32_t d = 0; d |= 1u << 4; d |= (uint32_t)(n >> 3) << 17; d |= (uint32_t)(m >> 4) << 24; return d;
}

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

template <int ts>
__global__ void __launch_bounds__(128, 1) bench(int M, int N, int iters, int distinct, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tm = slot;
    if (threadIdx.x < 32) {
        uint32_t id = idesc_mn(M, N);
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                // walk over `distinct` different 16 KB A tiles so the operand read is not trivially cached
                uint32_t a = base + (i % distinct) * 16384, b = base + 131072;
                uint64_t da = umma_desc_sw128(a), db = umma_desc_sw128(b);
#pragma unroll
                for (int k = 0; k < 4; ++k) if (elect_one()) {
                    if (ts) umma_f16_ts(tm, tm + 256 + ((i % distinct) * 4 + k) * 8, db + 2 * k, id, 1);
                    else umma_f16(tm, da + 2 * k, db + 2 * k, id, 1);
                }
            }
            if (elect_one()) umma_commit(&bar);
            long long t1 = clock64();
            mbar_wait(&bar, rep & 1);
            long long t2 = clock64();
            if (rep == 2 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int shapes[][2] = {{128, 16}, {128, 32}, {128, 64}, {128, 128}, {128, 256}, {64, 8}, {64, 16}, {64, 64}, {64, 256}};
    for (auto& s : shapes) {
        for (int ts : {0, 1}) {
            int iters = 64, distinct = 8;
            if (ts) bench<1><<<1, 128, 200 * 1024>>>(s[0], s[1], iters, distinct, d); else bench<0><<<1, 128, 200 * 1024>>>(s[0], s[1], iters, distinct, d);
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("M=%3d N=%3d A-from-TMEM=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (%s)\n", s[0], s[1], ts,
                   h[0] / (4.0 * iters), h[1] / (4.0 * iters), cudaGetErrorString(e));
        }
    }
    return 0;
}
