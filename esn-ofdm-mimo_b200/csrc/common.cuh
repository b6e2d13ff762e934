// Shared device/host helpers for the esn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "esn_b200.h"

#define ESN_CUDA_TRY(expr)                      \
    do {                                        \
        cudaError_t _e = (expr);                \
        if (_e != cudaSuccess) return (int)_e;  \
    } while (0)

static inline int esn_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// Counter-hash state noise.  One key per (frame, time step); one 32-bit hash serves the two
// neurons of a pair (neuron >> 1): the even neuron takes the low 16 bits, the odd neuron the high
// 16, so a thread that owns one frame and walks over neurons pays one hash per two values.
// uniform = bits * 2^-16 in [0,1).  The same integer recipe is restated in esn_b200/noise.py so
// tests can feed the oracle the identical stream.
__host__ __device__ __forceinline__ uint32_t esn_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t esn_noise_key(uint64_t seed, uint32_t frame, uint32_t row) {
    uint32_t k = esn_mix32((uint32_t)seed + 0x9E3779B9U * frame);
    return esn_mix32(k ^ (row * 0x85EBCA6BU + (uint32_t)(seed >> 32)));
}
// 32 noise bits of a neuron pair from the (well mixed) key plus the pair index: one xorshift, one
// 32 x 32 -> 64 bit multiply, high word folded onto the low word -- 5 instructions per two values.
__host__ __device__ __forceinline__ uint32_t esn_fold32(uint32_t x) {
    x ^= x >> 16;
    const uint64_t m = (uint64_t)x * 0x9E3779B1U;
    return (uint32_t)m ^ (uint32_t)(m >> 32);
}
__host__ __device__ __forceinline__ uint32_t esn_noise_bits(uint32_t key, uint32_t neuron) {
    return esn_fold32(key + (neuron >> 1) * 0xC2B2AE35U);   // low half: even neuron, high half: odd neuron
}
__host__ __device__ __forceinline__ float esn_noise_uniform(uint32_t key, uint32_t neuron) {
    const uint32_t h = esn_noise_bits(key, neuron);
    return (float)((neuron & 1u) ? (h >> 16) : (h & 0xFFFFu)) * (1.0f / 65536.0f);
}

template <typename T> struct esn_cplx { T re, im; };

template <typename T> __device__ __forceinline__ T esn_tanh(T x);
template <> __device__ __forceinline__ float esn_tanh<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double esn_tanh<double>(double x) { return tanh(x); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
