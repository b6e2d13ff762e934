// fp64 throughput of one B200: independent DFMA chains and DMMA (mma.sync m8n8k4 f64) chains, all SMs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak_probe fp64_peak_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double seed) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = seed + i + threadIdx.x;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double seed) {
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0);
        dfma_kernel<8><<<sms * 8, 256>>>(out, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  8 chains, %d CTAs x 256: %.2f ms -> %.1f TFLOP/s\n", sms * 8, ms, 2.0 * sms * 8 * 256 * 8 * (double)iters / ms / 1e9);
        cudaEventRecord(e0);
        dmma_kernel<8><<<sms * 8, 256>>>(out, iters, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  8 chains, %d CTAs x 256: %.2f ms -> %.1f TFLOP/s\n", sms * 8, ms, 2.0 * 256 * (double)sms * 8 * 8 * 8 * (double)iters / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
