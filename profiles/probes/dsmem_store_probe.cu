// Microbenchmark: cost of the T2 epilogue's state exchange -- 16 warps each writing a
// [64 frames x 32 neurons] block of fp16 into a SWIZZLE_128B tile that lives either in
// the local CTA or in the peer CTA of a 2-CTA cluster (st.shared::cluster).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u16(uint32_t a, unsigned short v) {
    asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t a, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>   // 0: local u16, 1: remote u16, 2: remote u32 (two neurons packed), 3: local u32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) probe(long long *out, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = warp * 32 + lane;                     // neuron 0..511
    const int k = n & 63;
    const uint32_t dst_rank = (MODE == 1 || MODE == 2) ? (rank ^ 1) : rank;
    const uint32_t tile = mapa((uint32_t)__cvta_generic_to_shared(smem), dst_rank) + (n >> 6) * 8192;
    cluster_sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE < 2) {
#pragma unroll 8
            for (int f = 0; f < 64; ++f) {
                uint32_t a = tile + (f >> 3) * 1024 + (f & 7) * 128 + ((((k >> 3) & 7) ^ (f & 7)) << 4) + (k & 7) * 2;
                st_cluster_u16(a, (unsigned short)(it + f));
                st_cluster_u16(a + 73728, (unsigned short)(it - f));
            }
        } else {
            // lane pair packs neurons (k, k+1): even lane writes even frames, odd lane odd frames
            const int kk = k & ~1;
#pragma unroll 8
            for (int f0 = 0; f0 < 64; f0 += 2) {
                const int f = f0 + (lane & 1);
                uint32_t a = tile + (f >> 3) * 1024 + (f & 7) * 128 + ((((kk >> 3) & 7) ^ (f & 7)) << 4) + (kk & 7) * 2;
                st_cluster_u32(a, (uint32_t)(it + f));
                st_cluster_u32(a + 73728, (uint32_t)(it - f));
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    cluster_sync();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int smem = 160 * 1024, iters = 200;
    const char *names[] = {"local u16", "remote u16", "remote u32 packed", "local u32 packed"};
    for (int mode = 0; mode < 4; ++mode) {
        void (*k)(long long *, int) = mode == 0 ? probe<0> : mode == 1 ? probe<1> : mode == 2 ? probe<2> : probe<3>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<<<2, 512, smem>>>(d, iters);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-20s: %lld / %lld cycles per step (64 frames x 512 neurons x hi+lo) [%s]\n", names[mode], h[0], h[1],
               cudaGetErrorString(e));
    }
    return 0;
}
