// Microbenchmark: L2 -> shared memory streaming rate of one SM (and of all 148) through a ring of slots filled by
// bulk copies, with a consumer that frees a slot as soon as its data has landed -- the weight / state stream of
// the tensor-core recurrence kernels without the MMAs.
//   mode 0: tensor map, uint16 elements, SWIZZLE_NONE, box 256 x 32 rows of 512 B (16 KB; what esn_predict_tc2 / tcs use)
//   mode 1: tensor map, uint16, SWIZZLE_128B, box 64 x 128 rows of 128 B (16 KB; the GEMM-style box)
//   mode 2: plain 1-D bulk copies of `req` bytes (4 KB .. 64 KB), P producer threads (one warp each) taking turns
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -lcuda -o profiles/probes/tma_stream_probe profiles/probes/tma_stream_probe.cu
#include <cstdio>
#include <vector>
#include "../../esn-ofdm-mimo_b200/csrc/tc_common.cuh"

namespace {
constexpr int MAXS = 48;

__device__ __forceinline__ void tma_g2s_local(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

__global__ void __launch_bounds__(256, 1)
probe(int mode, int S, int req, int P, int n, size_t region, int private_region, const unsigned char *src,
      const __grid_constant__ CUtensorMap map_lin, const __grid_constant__ CUtensorMap map_sw, long long *out) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[MAXS], empty[MAXS];
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < MAXS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t off0 = private_region ? (size_t)blockIdx.x * 262144 : 0;      // private: 256 KB per CTA
    const size_t span = private_region ? 262144 : region;
    long long t0 = clock64();
    if (warp < P && elect_one()) {                        // producers: warp w takes items w, w + P, ...
        for (int it = warp; it < n; it += P) {
            const int slot = it % S;
            mbar_wait<false>(&empty[slot], ((it / S) & 1) ^ 1);
            const uint32_t dst = smem_u32(base) + slot * req;
            const size_t off = off0 + ((size_t)it * req) % span;
            mbar_expect_tx(&full[slot], (uint32_t)req);
            if (mode == 0) tma_g2s_local(dst, &map_lin, 0, (int)(off / 512), smem_u32(&full[slot]));
            else if (mode == 1) tma_g2s_local(dst, &map_sw, 0, (int)(off / 128), smem_u32(&full[slot]));
            else bulk_g2s(dst, src + off, (uint32_t)req, smem_u32(&full[slot]));
        }
    } else if (warp == 7 && elect_one()) {                // consumer: free the slot as soon as the data is there
        for (int it = 0; it < n; ++it) {
            const int slot = it % S;
            mbar_wait<false>(&full[slot], (it / S) & 1);
            mbar_arrive(&empty[slot]);
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

bool make_map(CUtensorMap *m, void *base, size_t bytes, unsigned inner, unsigned rows, CUtensorMapSwizzle sw) {
    const cuuint64_t dims[2] = {inner, (cuuint64_t)(bytes / (inner * 2))};
    const cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
    const cuuint32_t box[2] = {inner, rows}, estr[2] = {1, 1};
    return cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

int main() {
    const size_t region = 72 * 16384;                      // 1.15 MB: the cfg3 weight image, shared by all CTAs
    const size_t bytes = (size_t)148 * 262144 + region;
    unsigned char *src;
    cudaMalloc(&src, bytes);
    cudaMemset(src, 1, bytes);
    CUtensorMap m_lin, m_sw;
    if (!make_map(&m_lin, src, bytes, 256, 32, CU_TENSOR_MAP_SWIZZLE_NONE) || !make_map(&m_sw, src, bytes, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B)) {
        printf("tensor map failed\n");
        return 1;
    }
    long long *out;
    cudaMalloc(&out, sizeof(long long) * 148);
    const size_t smem = 196608 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("bytes / cycle / SM streamed from L2 into a ring of shared-memory slots\n");
    printf("%-34s %6s %5s %5s %5s | %8s %8s\n", "mode", "req", "slots", "prod", "grid", "shared", "private");
    struct Cfg { int mode, req, S, P; };
    std::vector<Cfg> cfgs = {{0, 16384, 4, 1}, {0, 16384, 12, 1}, {1, 16384, 4, 1}, {1, 16384, 12, 1},
                             {2, 4096, 12, 1}, {2, 4096, 48, 1}, {2, 4096, 48, 4}, {2, 8192, 24, 1}, {2, 8192, 24, 2},
                             {2, 16384, 4, 1}, {2, 16384, 12, 1}, {2, 16384, 12, 2}, {2, 16384, 12, 4},
                             {2, 32768, 6, 1}, {2, 32768, 6, 2}, {2, 65536, 3, 1}, {2, 65536, 3, 3}};
    const char *names[3] = {"0 tensor u16 no swizzle 512 B rows", "1 tensor u16 SWIZZLE_128B rows", "2 plain bulk"};
    for (const Cfg &c : cfgs)
        for (int grid : {1, 148}) {
            double res[2];
            for (int priv = 0; priv < 2; ++priv) {
                const int n = (int)(64u * 1024 * 1024 / c.req);
                for (int rep = 0; rep < 2; ++rep)
                    probe<<<grid, 256, smem>>>(c.mode, c.S, c.req, c.P, n, region, priv, src, m_lin, m_sw, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s (mode %d)\n", cudaGetErrorString(e), c.mode); return 1; }
                std::vector<long long> h(grid);
                cudaMemcpy(h.data(), out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
                double worst = 0;
                for (int b = 0; b < grid; ++b) worst = std::max(worst, (double)h[b]);
                res[priv] = (double)n * c.req / worst;
            }
            printf("%-34s %6d %5d %5d %5d | %8.1f %8.1f\n", names[c.mode], c.req, c.S, c.P, grid, res[0], res[1]);
        }
    return 0;
}
