// numpy's legacy generator on the device: the state noise of pyESN comes from `random_state_.rand(N_res)` per
// time step (reference libs/pyESN.py:125), i.e. from MT19937 through numpy's `random_sample`
// (double = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53 from two consecutive 32-bit outputs).  Drawing the 267 K
// doubles of one cfg3 frame on the host costs 1.3 ms -- as much as the recurrence kernel itself on the one-frame
// path.  The twist of a 624-word block only looks sequential: word i needs old[i], old[i+1] and old[i+397] for
// i < 227, and new[i-227] afterwards, so a block is three dependent phases of <= 227 independent words.  One
// CTA steps the generator block by block from the host generator's state (uploaded: 624 words + position),
// writes the tempered words, and returns the advanced state, which the host puts back with set_state(): the
// stream the caller's generator sees is exactly what the reference would have left behind.
#include <cstdint>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MT_A = 0x9908b0dfu, MT_UP = 0x80000000u, MT_LO = 0x7fffffffu;

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & MT_UP) | (nxt & MT_LO);
    return far ^ (y >> 1) ^ ((y & 1u) ? MT_A : 0u);
}

// state: [624 key words][1 position]; words: [n_words] tempered outputs in generation order
__global__ void __launch_bounds__(256) mt19937_words_kernel(uint32_t *__restrict__ state, long long n_words,
                                                            uint32_t *__restrict__ words) {
    __shared__ uint32_t bufA[MT_N], bufB[MT_N];
    uint32_t *cur = bufA, *nxt = bufB;
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += blockDim.x) cur[i] = state[i];
    int pos = (int)state[MT_N];
    __syncthreads();
    long long done = 0;
    while (done < n_words) {
        if (pos >= MT_N) {                                             // twist: three phases, then swap buffers
            for (int i = tid; i < MT_N - MT_M; i += blockDim.x) nxt[i] = mt_mix(cur[i], cur[i + 1], cur[i + MT_M]);
            __syncthreads();
            for (int i = MT_N - MT_M + tid; i < 2 * (MT_N - MT_M); i += blockDim.x)
                nxt[i] = mt_mix(cur[i], cur[i + 1], nxt[i - (MT_N - MT_M)]);
            __syncthreads();
            for (int i = 2 * (MT_N - MT_M) + tid; i < MT_N; i += blockDim.x)
                nxt[i] = mt_mix(cur[i], i + 1 < MT_N ? cur[i + 1] : nxt[0], nxt[i - (MT_N - MT_M)]);
            __syncthreads();
            uint32_t *t = cur; cur = nxt; nxt = t;
            pos = 0;
        }
        const int n = (int)min((long long)(MT_N - pos), n_words - done);
        for (int j = tid; j < n; j += blockDim.x) {
            uint32_t y = cur[pos + j];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            words[done + j] = y;
        }
        pos += n;
        done += n;
        __syncthreads();
    }
    for (int i = tid; i < MT_N; i += blockDim.x) state[i] = cur[i];
    if (tid == 0) state[MT_N] = (uint32_t)pos;
}

template <typename T>
__global__ void mt19937_doubles_kernel(const uint32_t *__restrict__ words, long long count, T *__restrict__ out) {
    for (long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x; d < count; d += (long long)gridDim.x * blockDim.x) {
        const uint2 w = reinterpret_cast<const uint2 *>(words)[d];
        out[d] = (T)(((double)(w.x >> 5) * 67108864.0 + (double)(w.y >> 6)) / 9007199254740992.0);
    }
}

}  // namespace

extern "C" int esn_mt19937_uniforms(uint32_t *state_dev, long long count, int dtype, uint32_t *words_scratch,
                                    void *out, void *stream) {
    if (!state_dev || !words_scratch || !out || count <= 0) return ESN_E_BADARG;
    if (dtype != ESN_F32 && dtype != ESN_F64) return ESN_E_BADARG;
    if ((uintptr_t)words_scratch % 8) return ESN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    mt19937_words_kernel<<<1, 256, 0, st>>>(state_dev, 2 * count, words_scratch);
    const int blocks = (int)std::min<long long>((count + 255) / 256, 148 * 8);
    if (dtype == ESN_F64) mt19937_doubles_kernel<double><<<blocks, 256, 0, st>>>(words_scratch, count, (double *)out);
    else mt19937_doubles_kernel<float><<<blocks, 256, 0, st>>>(words_scratch, count, (float *)out);
    return esn_launch_status();
}
