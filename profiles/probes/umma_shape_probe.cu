// Microbenchmark: tcgen05.mma cta_group::2 kind::f16 issue-to-completion time per instruction for the
// operand shapes the pair kernel could use.  Operand contents are irrelevant (zeroed shared memory).
//   mode 0: M=256 N=128, A K-major (weights, 128 rows per CTA), B MN-major (state, 64 frames per CTA)  [current main chain]
//   mode 1: M=128 N=256, A K-major (state, 64 frames per CTA),  B K-major (weights, 128 rows per CTA) [swapped roles]
//   mode 2: M=256 N=16,  A MN-major state, B K-major                                                  [old readout]
//   mode 3: M=128 N=16,  A MN-major state, B K-major                                                  [current readout]
//   mode 4: M=128 N=16,  A K-major state,  B K-major                                                  [swapped readout]
//   mode 5: M=256 N=256, A K-major, B K-major
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o profiles/probes/umma_shape_probe profiles/probes/umma_shape_probe.cu
#include <cstdio>
#include "../../esn-ofdm-mimo_b200/csrc/recurrence_tc.cu"

namespace {
__device__ __forceinline__ uint32_t umma_idesc_major(int M, int N, int a_mn, int b_mn) {
    return umma_idesc(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
// MN-major SWIZZLE_128B descriptor, low word: start address, LBO = 0
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr) { return (saddr >> 4) & 0x3FFF; }
}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int mode, int n, long long *out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (rank == 0 && tid == 0) {
        const uint32_t w = desc_lo(smem_u32(base)), xk = desc_lo(smem_u32(base + 65536)), xmn = desc_lo_mn(smem_u32(base + 65536));
        uint32_t idesc, a, b, astep, bstep;
        switch (mode) {
            case 0: idesc = umma_idesc_major(256, 128, 0, 1); a = w; b = xmn; astep = 2; bstep = 128; break;
            case 1: idesc = umma_idesc_major(128, 256, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            case 2: idesc = umma_idesc_major(256, 16, 1, 0); a = xmn; b = w; astep = 128; bstep = 2; break;
            case 3: idesc = umma_idesc_major(128, 16, 1, 0); a = xmn; b = w; astep = 128; bstep = 2; break;
            case 4: idesc = umma_idesc_major(128, 16, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            case 5: idesc = umma_idesc_major(256, 256, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            case 6: idesc = umma_idesc_major(128, 144, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            case 7: idesc = umma_idesc_major(128, 128, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            case 8: idesc = umma_idesc_major(128, 64, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
            default: idesc = umma_idesc_major(128, 32, 0, 0); a = xk; b = w; astep = 2; bstep = 2; break;
        }
        const long long t0 = clock64();
        const uint32_t a1 = a + astep, a2 = a + 2 * astep, a3 = a + 3 * astep, b1 = b + bstep, b2 = b + 2 * bstep, b3 = b + 3 * bstep;
        for (int i = 0; i < n; i += 8) {           // descriptors precomputed: pure issue rate
            umma2_f16(tmem, a, b, idesc, 1u);  umma2_f16(tmem, a1, b1, idesc, 1u);
            umma2_f16(tmem, a2, b2, idesc, 1u); umma2_f16(tmem, a3, b3, idesc, 1u);
            umma2_f16(tmem, a, b1, idesc, 1u); umma2_f16(tmem, a1, b2, idesc, 1u);
            umma2_f16(tmem, a2, b3, idesc, 1u); umma2_f16(tmem, a3, b, idesc, 1u);
        }
        const long long t1 = clock64();
        umma2_commit_pair(&bar);
        mbar_wait<false>(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    } else if (tid == 0) {
        mbar_wait<false>(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const size_t smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[] = {"M256 N128 A=K B=MN (main now)", "M128 N256 A=K B=K (swapped)", "M256 N16 A=MN (old readout)",
                           "M128 N16 A=MN (readout now)", "M128 N16 A=K (swapped readout)", "M256 N256 A=K B=K",
                           "M128 N144 A=K B=K", "M128 N128 A=K B=K", "M128 N64 A=K B=K", "M128 N32 A=K B=K"};
    for (int mode = 0; mode < 10; ++mode)
        for (int n : {64, 512}) {
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<2, 128, smem>>>(mode, n, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-34s n=%4d  issue %7lld cyc  done %7lld cyc  -> %.1f cyc/MMA\n", names[mode], n, h[0], h[1], (double)h[1] / n);
        }
    return 0;
}
