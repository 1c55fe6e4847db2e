// Probe of the register layouts the fragment epilogue relies on (run on a B200):
//   tcgen05.ld.16x256b.x4 - which (lane, column) each thread's 16 registers hold
//   stmatrix.x4.trans / ldmatrix.x4.trans - that the fragment lands in shared memory as [frame][channel]
// TMEM is filled through tcgen05.st.32x32b (thread = lane, register = column) with value = lane * 1000 + column.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(float* regs_out, unsigned short* smem_out, unsigned* ld_out) {
    __shared__ uint32_t slot;
    __shared__ __align__(1024) unsigned short tile[64 * 64];      // [row = frame][64 channels] b16
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    // every warp w writes its lane quadrant: lanes 32w..32w+31, columns 0..31
    uint32_t taddr = base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(static_cast<float>((warp * 32 + lane) * 1000 + c0 + i));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {          // quadrant 1: lanes 32..63; read its upper half (lanes 48..63), columns 0..31
        uint32_t r[16];
        const uint32_t a = base + (static_cast<uint32_t>(32 + 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(a) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) regs_out[lane * 16 + i] = __uint_as_float(r[i]);
        // pack pairs (r[4g], r[4g+1]) = channel row t/4 and (r[4g+2], r[4g+3]) = row t/4 + 8 as b16 pairs holding small ints:
        // encode value = (lane_rel * 64 + col) so it fits fp16 exactly: lane_rel in 0..15, col 0..31
        uint32_t p[8];
        for (int g = 0; g < 4; ++g)
            for (int h = 0; h < 2; ++h) {
                const float f0 = __uint_as_float(r[4 * g + 2 * h]), f1 = __uint_as_float(r[4 * g + 2 * h + 1]);
                const int l0 = static_cast<int>(f0) / 1000 - 48, c0 = static_cast<int>(f0) % 1000;
                const int l1 = static_cast<int>(f1) / 1000 - 48, c1 = static_cast<int>(f1) % 1000;
                const __half2 hh = __floats2half2_rn(static_cast<float>(l0 * 64 + c0), static_cast<float>(l1 * 64 + c1));   // .x = low half
                p[2 * g + h] = *reinterpret_cast<const uint32_t*>(&hh);
            }
        for (int i = lane; i < 64 * 64; i += 32) tile[i] = 0xFFFF;
        __syncwarp();
        // frame groups g = 0,1 (frames 0..15): matrices (g0,h0) (g0,h1) (g1,h0) (g1,h1); memory row = frame, 8 channels per 16 B
        for (int gp = 0; gp < 2; ++gp) {
            const int i = lane >> 3, j = lane & 7;           // thread 8i + j: address of memory row j of matrix i
            const int g = 2 * gp + (i >> 1), h = i & 1;
            const int row = 8 * g + j, chunk = h;            // channels 8h..8h+7 -> 16-byte chunk h (no swizzle in the probe)
            const uint32_t addr = smem_u32(&tile[row * 64 + chunk * 8]);
            asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(p[2 * (2 * gp) + 0]),
                         "r"(p[2 * (2 * gp) + 1]), "r"(p[2 * (2 * gp + 1) + 0]), "r"(p[2 * (2 * gp + 1) + 1]) : "memory");
        }
        __syncwarp();
        for (int i = lane; i < 32 * 16; i += 32) smem_out[i] = tile[(i / 16) * 64 + (i % 16)];      // rows 0..31 x channels 0..15
        // ldmatrix.trans back: should reproduce p[0..3] of gp = 0
        {
            const int i = lane >> 3, j = lane & 7;
            const int g = (i >> 1), h = i & 1;
            const uint32_t addr = smem_u32(&tile[(8 * g + j) * 64 + h * 8]);
            uint32_t q0, q1, q2, q3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(addr));
            ld_out[lane * 8 + 0] = q0; ld_out[lane * 8 + 1] = q1; ld_out[lane * 8 + 2] = q2; ld_out[lane * 8 + 3] = q3;
            ld_out[lane * 8 + 4] = p[0]; ld_out[lane * 8 + 5] = p[1]; ld_out[lane * 8 + 6] = p[2]; ld_out[lane * 8 + 7] = p[3];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
}

int main() {
    float* d_regs; unsigned short* d_sm; unsigned* d_ld;
    cudaMalloc(&d_regs, 32 * 16 * 4); cudaMalloc(&d_sm, 32 * 16 * 2); cudaMalloc(&d_ld, 32 * 8 * 4);
    probe<<<1, 128>>>(d_regs, d_sm, d_ld);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    float regs[32 * 16]; unsigned short sm[32 * 16]; unsigned ld[32 * 8];
    cudaMemcpy(regs, d_regs, sizeof(regs), cudaMemcpyDeviceToHost);
    cudaMemcpy(sm, d_sm, sizeof(sm), cudaMemcpyDeviceToHost);
    cudaMemcpy(ld, d_ld, sizeof(ld), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 16; ++i) {
            const int g = i >> 2, rh = (i >> 1) & 1, e2 = i & 1;
            const int want_lane = 48 + t / 4 + 8 * rh, want_col = 8 * g + 2 * (t % 4) + e2;
            if (regs[t * 16 + i] != want_lane * 1000 + want_col) {
                if (bad < 12) printf("thread %d reg %d: got lane %d col %d, expected lane %d col %d\n", t, i, (int)regs[t * 16 + i] / 1000, (int)regs[t * 16 + i] % 1000, want_lane, want_col);
                ++bad;
            }
        }
    printf("tcgen05.ld.16x256b.x4 layout mismatches: %d / 512\n", bad);
    int bad2 = 0;
    for (int row = 0; row < 32; ++row)
        for (int ch = 0; ch < 16; ++ch) {
            const float v = __half2float(*reinterpret_cast<__half*>(&sm[row * 16 + ch]));
            const float want = ch * 64 + row;          // memory[frame = row][channel = ch] = (lane_rel = ch, col = row)
            if (v != want) { if (bad2 < 12) printf("smem row %d ch %d: got %g expected %g\n", row, ch, v, want); ++bad2; }
        }
    printf("stmatrix.trans [frame][channel] mismatches: %d / 512\n", bad2);
    int bad3 = 0;
    for (int t = 0; t < 32; ++t) for (int i = 0; i < 4; ++i) if (ld[t * 8 + i] != ld[t * 8 + 4 + i]) ++bad3;
    printf("ldmatrix.trans round-trip mismatches: %d / 128\n", bad3);
    return 0;
}
