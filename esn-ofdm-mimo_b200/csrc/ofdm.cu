// OFDM-side kernels of the detection path (north-star subsystem 3): batched
// shared-memory FFTs fused with their neighbours, pilot LS channel estimation
// with time-domain MMSE shrinkage, per-subcarrier ZF/MMSE equalisation, hard QAM
// demapping and bit-error counting.  All HBM-bound: one pass over the data, one
// CTA per (frame, antenna) FFT.  Reference sites are cited per entry point in
// include/esn_b200.h (system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py).
#include <algorithm>
#include <cstdint>
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int MAX_FFT = 4096;
constexpr int MAXA = 8;   // max antennas per side in the equaliser

__device__ __forceinline__ void sincospi_t(float x, float *s, float *c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double *s, double *c) { sincospi(x, s, c); }
__device__ __forceinline__ float log_t(float x) { return logf(x); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }

__device__ __forceinline__ int bitrev(int x, int logn) { return (int)(__brev((unsigned)x) >> (32 - logn)); }

// twiddle table tw[k] = exp(-2 pi i k / N), k < N/2
template <typename T>
__device__ void fft_make_twiddles(T *twr, T *twi, int N) {
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
        T s, c;
        sincospi_t((T)(-2.0) * (T)k / (T)N, &s, &c);
        twr[k] = c; twi[k] = s;
    }
}

// In-place radix-2 DIT on data already stored in bit-reversed order.  inverse:
// conjugate twiddles (no scaling here).  Ends with a __syncthreads().
template <typename T>
__device__ void fft_inplace(T *re, T *im, const T *twr, const T *twi, int N, int logn, bool inverse) {
    __syncthreads();
    for (int s = 1; s <= logn; ++s) {
        const int half = 1 << (s - 1);
        const int tstride = N >> s;
        for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
            const T wr = twr[pos * tstride];
            const T wi = inverse ? -twi[pos * tstride] : twi[pos * tstride];
            const T xr = re[i1], xi = im[i1];
            const T tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
            const T ur = re[i0], ui = im[i0];
            re[i0] = ur + tr; im[i0] = ui + ti;
            re[i1] = ur - tr; im[i1] = ui - ti;
        }
        __syncthreads();
    }
}

// Skewed shared-memory index: one pad word per 32 keeps the power-of-two strides of the FFT (and the
// bit-reversed scatter of its input) off a single bank.
__device__ __forceinline__ int skew(int i) { return i + (i >> 5); }

// r radix-2 layers (stages s0+1 .. s0+r of the DIT on bit-reversed data) fused in registers: each work
// item loads 2^r points at stride 2^s0, runs the r butterfly layers and stores them back, so a 512-point
// FFT takes 3 passes over shared memory instead of 9.  `nb` arrays of N points side by side (skewed).
template <typename T, int r>
__device__ __forceinline__ void fft_pass(T *re, T *im, const T *twr, const T *twi, int N, int logn, int s0, int nb) {
    constexpr int R = 1 << r;
    const int h = 1 << s0, per = N >> r, items = per * nb;
    for (int q = threadIdx.x; q < items; q += blockDim.x) {
        const int a = q >> (logn - r), j = q & (per - 1);
        const int pos = j & (h - 1), base = a * N + ((j >> s0) << (s0 + r)) + pos;
        T xr[R], xi[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { xr[k] = re[skew(base + k * h)]; xi[k] = im[skew(base + k * h)]; }
#pragma unroll
        for (int l = 0; l < r; ++l) {
            const int d = 1 << l, tstride = N >> (s0 + l + 1);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if ((k & d) == 0) {
                    const int pw = (pos + (k & (d - 1)) * h) * tstride;
                    const T wr = twr[pw], wi = twi[pw];
                    const T tr = wr * xr[k + d] - wi * xi[k + d], ti = wr * xi[k + d] + wi * xr[k + d];
                    xr[k + d] = xr[k] - tr; xi[k + d] = xi[k] - ti;
                    xr[k] += tr; xi[k] += ti;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k) { re[skew(base + k * h)] = xr[k]; im[skew(base + k * h)] = xi[k]; }
    }
    __syncthreads();
}
// forward FFT of nb arrays already in bit-reversed order (caller has synchronised)
template <typename T>
__device__ void fft_batched_radix8(T *re, T *im, const T *twr, const T *twi, int N, int logn, int nb) {
    int s0 = 0;
    const int rem = logn % 3;
    if (rem == 1) { fft_pass<T, 1>(re, im, twr, twi, N, logn, 0, nb); s0 = 1; }
    if (rem == 2) { fft_pass<T, 2>(re, im, twr, twi, N, logn, 0, nb); s0 = 2; }
    for (; s0 < logn; s0 += 3) fft_pass<T, 3>(re, im, twr, twi, N, logn, s0, nb);
}

__device__ __forceinline__ int ilog2(int n) { return 31 - __clz(n); }

template <typename T>
struct Slicer {
    int L, bits;
    T s;
    __device__ Slicer(int qam_bits) {
        bits = qam_bits;
        L = 1 << (qam_bits / 2);
        s = sqrt_t((T)(2.0 * (L * L - 1) / 3.0));
    }
    __device__ int level(T v) const {
        int i = (int)floor((v * s + (T)L) * (T)0.5);
        return min(max(i, 0), L - 1);
    }
    __device__ int index(T re, T im) const { return L * level(re) + level(im); }
    // distance (constellation units) to the nearest decision boundary on either axis.  In units of half the point
    // spacing (u = v s) the boundaries are the even integers -(L-2) .. L-2: the nearest one is the clamped
    // rounding of u to an even integer (closed form instead of a loop with a division per boundary).
    __device__ T axis_dist(T v) const {
        const T u = v * s, lim = (T)(L - 2);
        const T nb = fmin(fmax((T)2 * rint(u * (T)0.5), -lim), lim);
        return fabs(u - nb);
    }
    __device__ T boundary_dist(T re, T im) const {
        if (L < 2) return (T)1e30;
        return fmin(axis_dist(re), axis_dist(im)) / s;
    }
};

__device__ void block_add_counts(unsigned long long errs, unsigned long long near,
                                 unsigned long long *out) {
    __shared__ unsigned long long s_acc[2];
    if (threadIdx.x == 0) { s_acc[0] = 0; s_acc[1] = 0; }
    __syncthreads();
    for (int s = 16; s > 0; s >>= 1) {
        errs += __shfl_xor_sync(0xffffffffu, errs, s);
        near += __shfl_xor_sync(0xffffffffu, near, s);
    }
    if ((threadIdx.x & 31) == 0) {
        if (errs) atomicAdd(&s_acc[0], errs);
        if (near) atomicAdd(&s_acc[1], near);
    }
    __syncthreads();
    if (threadIdx.x == 0 && out) {
        if (s_acc[0]) atomicAdd(&out[0], s_acc[0]);
        if (s_acc[1]) atomicAdd(&out[1], s_acc[1]);
    }
}

// ---- ESN output -> complex -> FFT -> scale -> slicer -> errors -------------
template <typename T>
__global__ void unpack_fft_demap_kernel(const T *__restrict__ y, int rows, int N, int N_t,
                                        const T *__restrict__ Pi, int pi_stride, int qam_bits,
                                        T *__restrict__ X_hat, uint8_t *__restrict__ idx,
                                        const uint8_t *__restrict__ tx_idx, T eps,
                                        unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char sm[];
    T *re = reinterpret_cast<T *>(sm), *im = re + N, *twr = im + N, *twi = twr + N / 2;
    const int tx = blockIdx.x, b = blockIdx.y, logn = ilog2(N);
    fft_make_twiddles(twr, twi, N);
    const T *yb = y + (size_t)b * rows * 2 * N_t;
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        const int r = bitrev(t, logn);
        re[r] = yb[(size_t)t * 2 * N_t + 2 * tx];
        im[r] = yb[(size_t)t * 2 * N_t + 2 * tx + 1];
    }
    fft_inplace(re, im, twr, twi, N, logn, false);
    const T scale = (T)1 / ((T)N * sqrt_t(Pi[(size_t)b * pi_stride]));
    const Slicer<T> sl(qam_bits);
    unsigned long long errs = 0, near = 0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const T xr = re[k] * scale, xi = im[k] * scale;
        const size_t o = ((size_t)b * N + k) * N_t + tx;
        if (X_hat) { X_hat[2 * o] = xr; X_hat[2 * o + 1] = xi; }
        const int id = sl.index(xr, xi);
        if (idx) idx[o] = (uint8_t)id;
        if (tx_idx) errs += __popc((unsigned)(id ^ (int)tx_idx[o]));
        if (eps > (T)0 && sl.boundary_dist(xr, xi) < eps) near += 1;
    }
    block_add_counts(errs, near, counts);
}

// One CTA per frame, all N_t streams at once: the frame's N x 2 N_t block of ESN outputs is read with
// fully coalesced loads, the N_t FFTs run side by side in shared memory, and symbol indices / X_hat are
// written as contiguous rows.  Algorithmic traffic: N (2 N_t) reals in, N N_t bytes (+ N N_t for tx_idx)
// out -- one pass over HBM.
template <typename T>
__global__ void __launch_bounds__(512)
unpack_fft_demap_frame_kernel(const T *__restrict__ y, int rows, int N, int N_t, const T *__restrict__ Pi,
                              int pi_stride, int qam_bits, T *__restrict__ X_hat, uint8_t *__restrict__ idx,
                              const uint8_t *__restrict__ tx_idx, T eps, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tot = N * N_t, tots = tot + (tot >> 5) + 1;             // skewed array length
    T *re = reinterpret_cast<T *>(sm), *im = re + tots, *twr = im + tots, *twi = twr + N / 2;
    const int b = blockIdx.x, logn = ilog2(N), W2 = 2 * N_t;
    const bool w2pow2 = (W2 & (W2 - 1)) == 0;                         // shifts instead of integer divisions
    const int lw2 = ilog2(W2);
    fft_make_twiddles(twr, twi, N);
    const T *yb = y + (size_t)b * rows * W2;
    for (int e0 = 0; e0 < N * W2; e0 += 8 * blockDim.x) {             // contiguous, coalesced, 8 loads in flight
        T v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = e0 + i * blockDim.x + threadIdx.x;
            v[i] = e < N * W2 ? yb[e] : (T)0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = e0 + i * blockDim.x + threadIdx.x;
            if (e < N * W2) {
                const int t = w2pow2 ? (e >> lw2) : e / W2, c = e - t * W2;
                ((c & 1) ? im : re)[skew((c >> 1) * N + bitrev(t, logn))] = v[i];
            }
        }
    }
    __syncthreads();
    fft_batched_radix8(re, im, twr, twi, N, logn, N_t);              // N_t FFTs side by side, 3 layers per pass
    const T scale = (T)1 / ((T)N * sqrt_t(Pi[(size_t)b * pi_stride]));
    const Slicer<T> sl(qam_bits);
    unsigned long long errs = 0, near = 0;
    for (int e = threadIdx.x; e < N * N_t; e += blockDim.x) {          // e = k N_t + tx: contiguous outputs
        const int k = w2pow2 ? (e >> (lw2 - 1)) : e / N_t, tx = e - k * N_t;
        const T xr = re[skew(tx * N + k)] * scale, xi = im[skew(tx * N + k)] * scale;
        const size_t o = (size_t)b * N * N_t + e;
        if (X_hat) { X_hat[2 * o] = xr; X_hat[2 * o + 1] = xi; }
        const int id = sl.index(xr, xi);
        if (idx) idx[o] = (uint8_t)id;
        if (tx_idx) errs += __popc((unsigned)(id ^ (int)tx_idx[o]));
        if (eps > (T)0 && sl.boundary_dist(xr, xi) < eps) near += 1;
    }
    block_add_counts(errs, near, counts);
}

// ---- CP removal + FFT of the received samples -------------------------------
template <typename T>
__global__ void rx_fft_kernel(const T *__restrict__ y_cp, int N, int cp, int N_r, T *__restrict__ Y) {
    extern __shared__ __align__(16) unsigned char sm[];
    T *re = reinterpret_cast<T *>(sm), *im = re + N, *twr = im + N, *twi = twr + N / 2;
    const int rx = blockIdx.x, b = blockIdx.y, logn = ilog2(N);
    fft_make_twiddles(twr, twi, N);
    const T *src = y_cp + (size_t)b * (N + cp) * N_r * 2;
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        const int r = bitrev(t, logn);
        re[r] = src[((size_t)(cp + t) * N_r + rx) * 2];
        im[r] = src[((size_t)(cp + t) * N_r + rx) * 2 + 1];
    }
    fft_inplace(re, im, twr, twi, N, logn, false);
    const T scale = (T)1 / (T)N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const size_t o = ((size_t)b * N + k) * N_r + rx;
        Y[2 * o] = re[k] * scale; Y[2 * o + 1] = im[k] * scale;
    }
}

// Whole-frame version: one CTA per frame reads the N rows behind the CP as one contiguous block, runs the N_r
// FFTs side by side (three radix-2 layers per shared-memory pass) and writes Y [N][N_r] as contiguous rows.
template <typename T>
__global__ void __launch_bounds__(512)
rx_fft_frame_kernel(const T *__restrict__ y_cp, int N, int cp, int N_r, T *__restrict__ Y) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tot = N * N_r, tots = tot + (tot >> 5) + 1;
    T *re = reinterpret_cast<T *>(sm), *im = re + tots, *twr = im + tots, *twi = twr + N / 2;
    const int b = blockIdx.x, logn = ilog2(N), W2 = 2 * N_r;
    fft_make_twiddles(twr, twi, N);
    const T *src = y_cp + ((size_t)b * (N + cp) + cp) * W2;
    for (int e0 = 0; e0 < N * W2; e0 += 8 * blockDim.x) {             // contiguous, coalesced, 8 loads in flight
        T v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = e0 + i * blockDim.x + threadIdx.x;
            v[i] = e < N * W2 ? src[e] : (T)0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = e0 + i * blockDim.x + threadIdx.x;
            if (e < N * W2) {
                const int t = e / W2, c = e - t * W2;
                ((c & 1) ? im : re)[skew((c >> 1) * N + bitrev(t, logn))] = v[i];
            }
        }
    }
    __syncthreads();
    fft_batched_radix8(re, im, twr, twi, N, logn, N_r);
    const T scale = (T)1 / (T)N;
    T *dst = Y + (size_t)b * N * W2;
    for (int e = threadIdx.x; e < N * W2; e += blockDim.x) {           // e = (k N_r + rx) 2 + {re, im}: contiguous
        const int k = e / W2, c = e - k * W2;
        dst[e] = ((c & 1) ? im : re)[skew((c >> 1) * N + k)] * scale;
    }
}


// ---- 512-point FFT as three register-resident radix-8 stages (Stockham autosort) ---------------------------
// The generic path above runs nine radix-2 layers, three per shared-memory pass, on bit-reversed data; for the
// demos' N = 512 = 8^3 the whole transform is three passes of one 8-point DFT per work item with natural-order
// input AND output (no bit-reversed scatter), complex pairs kept together (8-byte shared-memory accesses) and
// the twiddles from a 512-entry table: ~1.8x fewer instructions per transform.  `nb` transforms side by side;
// input in `a` (index stream * F512_STR + f512_skew(t)), scratch `b`; the result is in `b`.
template <typename T> struct Cx { T re, im; };
constexpr int F512_STR = 512 + 32 + 9;                 // skewed length of one stream; = 9 mod 16, so that the streams of one
                                                       // sample (the frame's load and output order) fall into different banks
// one pad slot per 16 elements (8-byte elements, 16 bank pairs): aligned runs of 16 consecutive elements -- every stage's
// reads and the last stage's writes -- stay conflict-free, as do the stride-8 writes of the first stage; only the
// second stage's writes (two runs of 8 elements 64 apart) are 2-way.  (A slot per 8 elements made every consecutive
// access 2-way: a run of 16 then spans 17-18 slots.)
__device__ __forceinline__ int f512_skew(int i) { return i + (i >> 4); }

template <typename T>
__device__ __forceinline__ void f512_dft8(Cx<T> (&v)[8]) {
    const T h = (T)0.70710678118654752440;
#define F512_BF(x, y) { const Cx<T> t_ = v[x]; v[x].re = t_.re + v[y].re; v[x].im = t_.im + v[y].im; v[y].re = t_.re - v[y].re; v[y].im = t_.im - v[y].im; }
    F512_BF(0, 4) F512_BF(1, 5) F512_BF(2, 6) F512_BF(3, 7)
    { const Cx<T> t = v[5]; v[5].re = (t.re + t.im) * h; v[5].im = (t.im - t.re) * h; }           // x W8
    { const Cx<T> t = v[6]; v[6].re = t.im; v[6].im = -t.re; }                                    // x (-i)
    { const Cx<T> t = v[7]; v[7].re = (t.im - t.re) * h; v[7].im = -(t.re + t.im) * h; }          // x W8^3
    F512_BF(0, 2) F512_BF(1, 3) F512_BF(4, 6) F512_BF(5, 7)
    { const Cx<T> t = v[3]; v[3].re = t.im; v[3].im = -t.re; }
    { const Cx<T> t = v[7]; v[7].re = t.im; v[7].im = -t.re; }
    F512_BF(0, 1) F512_BF(2, 3) F512_BF(4, 5) F512_BF(6, 7)
#undef F512_BF
    // v now holds X[0], X[4], X[2], X[6], X[1], X[5], X[3], X[7]
}

// W^k = exp(-2 pi i k / 512) as exp(-2 pi i 16 a / 512) exp(-2 pi i b / 512), k = 16 a + b: 48 sincospi per CTA
// instead of 512 (ends with a __syncthreads())
template <typename T>
__device__ void fft512_make_table(Cx<T> *tw) {
    __shared__ Cx<T> base[48];
    if (threadIdx.x < 48) {
        const int k = threadIdx.x < 32 ? 16 * threadIdx.x : threadIdx.x - 32;
        T sn, cs;
        sincospi_t((T)(-2.0) * (T)k / (T)512, &sn, &cs);
        base[threadIdx.x].re = cs; base[threadIdx.x].im = sn;
    }
    __syncthreads();
    // laid out as the stages read it: entry (r - 1) * 8 + k = W^(8 k r) for the second stage (k = j & 7), entry
    // 56 + (r - 1) * 64 + j = W^(j r) for the third -- the lanes of a warp read consecutive entries (a table indexed
    // by the exponent put W^(8 k r), r even, of all eight k into one or two banks: l1tex 95 % busy, 37 M bank
    // conflicts per launch in profiles/r2_ncu_chain_kernels.txt)
    for (int e = threadIdx.x; e < 504; e += blockDim.x) {
        const int k = e < 56 ? 8 * (e & 7) * ((e >> 3) + 1) : ((e - 56) & 63) * (((e - 56) >> 6) + 1);
        const Cx<T> a = base[k >> 4], b = base[32 + (k & 15)];
        tw[e].re = a.re * b.re - a.im * b.im;
        tw[e].im = a.re * b.im + a.im * b.re;
    }
    __syncthreads();
}

// one stage: Ns = 1, 8, 64 (LOG_NS = 0, 3, 6)
template <typename T, int LOG_NS>
__device__ __forceinline__ void fft512_stage(const Cx<T> *in, Cx<T> *out, const Cx<T> *tw, int nb) {
    constexpr int NS = 1 << LOG_NS;
    for (int q = threadIdx.x; q < nb * 64; q += blockDim.x) {
        const int a = q >> 6, j = q & 63, k = j & (NS - 1);
        const Cx<T> *src = in + a * F512_STR;
        Cx<T> v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = src[f512_skew(j + 64 * r)];
        if (LOG_NS > 0) {
#pragma unroll
            for (int r = 1; r < 8; ++r) {
                const Cx<T> w = LOG_NS == 3 ? tw[(r - 1) * 8 + k] : tw[56 + (r - 1) * 64 + k];
                const T xr = v[r].re, xi = v[r].im;
                v[r].re = xr * w.re - xi * w.im;
                v[r].im = xr * w.im + xi * w.re;
            }
        }
        f512_dft8(v);
        Cx<T> *dst = out + a * F512_STR;
        const int j0 = ((j - k) << 3) + k;
        constexpr int ORD[8] = {0, 4, 2, 6, 1, 5, 3, 7};
#pragma unroll
        for (int r = 0; r < 8; ++r) dst[f512_skew(j0 + ORD[r] * NS)] = v[r];
    }
    __syncthreads();
}

// forward FFT-512 of nb streams: a -> (b) -> (a) -> b.  The caller has synchronised after filling a and tw.
template <typename T>
__device__ void fft512_forward(Cx<T> *a, Cx<T> *b, const Cx<T> *tw, int nb) {
    fft512_stage<T, 0>(a, b, tw, nb);
    fft512_stage<T, 3>(b, a, tw, nb);
    fft512_stage<T, 6>(a, b, tw, nb);
}

// The two N = 512 kernels below are persistent: grid = (CTAs that fit an SM) x SMs, every CTA builds its twiddle
// table once and walks frames b = blockIdx.x, + gridDim.x, ...; the global loads of its next frame are issued
// into registers before the current frame's three FFT stages and land in `a` (free after the last stage) behind
// the output loop, so the load latency hides behind the transform instead of heading every frame.
// NV = elements per thread and frame: 512 x nb streams / 256 threads (8 up to four streams, 16 up to eight)
#define F512_LOAD(v, src, ne)                                                                                    \
    _Pragma("unroll") for (int i = 0; i < NV; ++i) {                                                        \
        const int e = i * 256 + (int)threadIdx.x;                                                                \
        if (e < (ne)) v[i] = (src)[e];                                                                           \
    }
#define F512_STORE(v, A, ne, nb, p2, lg)                                                                         \
    _Pragma("unroll") for (int i = 0; i < NV; ++i) {                                                        \
        const int e = i * 256 + (int)threadIdx.x;                                                                \
        if (e < (ne)) {                                                                                          \
            const int t = p2 ? (e >> lg) : e / (nb), s_ = e - t * (nb);                                          \
            A[s_ * F512_STR + f512_skew(t)] = v[i];                                                              \
        }                                                                                                        \
    }

// ESN output -> FFT -> slicer -> error count for N = 512 (see unpack_fft_demap_frame_kernel for the semantics)
template <typename T, int NV>
__global__ void __launch_bounds__(256)
unpack_fft_demap_frame512_kernel(const T *__restrict__ y, int B, int rows, int N_t, const T *__restrict__ Pi,
                                 int pi_stride, int qam_bits, T *__restrict__ X_hat, uint8_t *__restrict__ idx,
                                 const uint8_t *__restrict__ tx_idx, T eps, unsigned long long *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int N = 512;
    Cx<T> *A = reinterpret_cast<Cx<T> *>(sm), *Bf = A + N_t * F512_STR, *tw = Bf + N_t * F512_STR;
    const int ne = N * N_t;
    const bool p2 = (N_t & (N_t - 1)) == 0;
    const int lg = ilog2(N_t);
    const Slicer<T> sl(qam_bits);
    unsigned long long errs = 0, near = 0;
    Cx<T> v[NV];
    int b = blockIdx.x;
    {
        const Cx<T> *yb = reinterpret_cast<const Cx<T> *>(y + (size_t)b * rows * 2 * N_t);   // [t][tx] complex pairs
        F512_LOAD(v, yb, ne)
        fft512_make_table(tw);                                     // built while the first frame is in flight
        F512_STORE(v, A, ne, N_t, p2, lg)
    }
    __syncthreads();
    for (; b < B; b += gridDim.x) {
        const int bn = b + gridDim.x;
        if (bn < B) {
            const Cx<T> *yb = reinterpret_cast<const Cx<T> *>(y + (size_t)bn * rows * 2 * N_t);
            F512_LOAD(v, yb, ne)
        }
        fft512_forward(A, Bf, tw, N_t);
        const T scale = (T)1 / ((T)N * sqrt_t(Pi[(size_t)b * pi_stride]));
        for (int e = threadIdx.x; e < ne; e += blockDim.x) {               // e = k N_t + tx: contiguous outputs
            const int k = p2 ? (e >> lg) : e / N_t, tx = e - k * N_t;
            const Cx<T> w = Bf[tx * F512_STR + f512_skew(k)];
            const T xr = w.re * scale, xi = w.im * scale;
            const size_t o = (size_t)b * ne + e;
            if (X_hat) { X_hat[2 * o] = xr; X_hat[2 * o + 1] = xi; }
            const int id = sl.index(xr, xi);
            if (idx) idx[o] = (uint8_t)id;
            if (tx_idx) errs += __popc((unsigned)(id ^ (int)tx_idx[o]));
            if (eps > (T)0 && sl.boundary_dist(xr, xi) < eps) near += 1;
        }
        if (bn < B) F512_STORE(v, A, ne, N_t, p2, lg)
        __syncthreads();                                           // next frame in `a`; every read of `b` is done
    }
    block_add_counts(errs, near, counts);
}

// CP strip + FFT of the received samples for N = 512 (see rx_fft_frame_kernel)
template <typename T, int NV>
__global__ void __launch_bounds__(256)
rx_fft_frame512_kernel(const T *__restrict__ y_cp, int B, int cp, int N_r, T *__restrict__ Y) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int N = 512;
    Cx<T> *A = reinterpret_cast<Cx<T> *>(sm), *Bf = A + N_r * F512_STR, *tw = Bf + N_r * F512_STR;
    const int ne = N * N_r;
    const bool p2 = (N_r & (N_r - 1)) == 0;
    const int lg = ilog2(N_r);
    const T scale = (T)1 / (T)N;
    Cx<T> v[NV];
    int b = blockIdx.x;
    {
        const Cx<T> *src = reinterpret_cast<const Cx<T> *>(y_cp + ((size_t)b * (N + cp) + cp) * 2 * N_r);
        F512_LOAD(v, src, ne)
        fft512_make_table(tw);
        F512_STORE(v, A, ne, N_r, p2, lg)
    }
    __syncthreads();
    for (; b < B; b += gridDim.x) {
        const int bn = b + gridDim.x;
        if (bn < B) {
            const Cx<T> *src = reinterpret_cast<const Cx<T> *>(y_cp + ((size_t)bn * (N + cp) + cp) * 2 * N_r);
            F512_LOAD(v, src, ne)
        }
        fft512_forward(A, Bf, tw, N_r);
        Cx<T> *dst = reinterpret_cast<Cx<T> *>(Y + (size_t)b * N * 2 * N_r);
        for (int e = threadIdx.x; e < ne; e += blockDim.x) {               // e = k N_r + rx: contiguous
            const int k = p2 ? (e >> lg) : e / N_r, rx = e - k * N_r;
            Cx<T> w = Bf[rx * F512_STR + f512_skew(k)];
            w.re *= scale; w.im *= scale;
            dst[e] = w;
        }
        if (bn < B) F512_STORE(v, A, ne, N_r, p2, lg)
        __syncthreads();
    }
}

// ---- pilot LS + linear inter/extrapolation + time-domain MMSE ---------------
template <typename T>
__global__ void chanest_kernel(const T *__restrict__ Y_LS, const T *__restrict__ X_LS, int N, int N_r,
                               int N_t, const T *__restrict__ Pi, const T *__restrict__ mag, int taps,
                               T No, T *__restrict__ H_LS, T *__restrict__ H_MMSE) {
    extern __shared__ __align__(16) unsigned char sm[];
    T *re = reinterpret_cast<T *>(sm), *im = re + N, *twr = im + N, *twi = twr + N / 2;
    T *hr = twi + N / 2, *hi = hr + N;                     // comb estimates, then full-band estimate
    const int tx = blockIdx.x, nr = blockIdx.y, b = blockIdx.z, logn = ilog2(N);
    fft_make_twiddles(twr, twi, N);
    const T pi_b = Pi[b], sp = sqrt_t(pi_b);
    const int M = (N - tx + N_t - 1) / N_t;
    const T *Yb = Y_LS + (size_t)b * N * N_r * 2;
    const T *Xb = X_LS + (size_t)b * N * N_t * 2;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const int k = tx + j * N_t;
        const T yr = Yb[((size_t)k * N_r + nr) * 2], yi = Yb[((size_t)k * N_r + nr) * 2 + 1];
        const T dr = Xb[((size_t)k * N_t + tx) * 2] * sp + (T)1e-12, di = Xb[((size_t)k * N_t + tx) * 2 + 1] * sp;
        const T den = dr * dr + di * di;
        hr[j] = (yr * dr + yi * di) / den;
        hi[j] = (yi * dr - yr * di) / den;
    }
    __syncthreads();
    // interpolate into (re, im) in natural order, emit H_LS, then bit-reverse for the IFFT
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        int j = (k >= tx) ? (k - tx) / N_t : 0;
        j = min(max(j, 0), max(M - 2, 0));
        T vr, vi;
        if (M >= 2) {
            const T t = (T)(k - (tx + j * N_t)) / (T)N_t;
            vr = hr[j] + t * (hr[j + 1] - hr[j]);
            vi = hi[j] + t * (hi[j + 1] - hi[j]);
        } else { vr = hr[0]; vi = hi[0]; }
        const size_t o = (((size_t)b * N + k) * N_r + nr) * N_t + tx;
        H_LS[2 * o] = vr; H_LS[2 * o + 1] = vi;
        const int r = bitrev(k, logn);
        re[r] = vr; im[r] = vi;
    }
    fft_inplace(re, im, twr, twi, N, logn, true);          // c_LS * N
    // shrink the first `taps` taps, zero the rest, forward FFT
    const T mmse_scaler = (No / pi_b) / ((T)N * (T)0.5);
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        T vr = (T)0, vi = (T)0;
        if (t < taps) {
            const T g = (T)1 / ((T)N * (mmse_scaler / mag[t] + (T)1));
            vr = re[t] * g; vi = im[t] * g;
        }
        hr[t] = vr; hi[t] = vi;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        const int r = bitrev(t, logn);
        re[r] = hr[t]; im[r] = hi[t];
    }
    fft_inplace(re, im, twr, twi, N, logn, false);
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const size_t o = (((size_t)b * N + k) * N_r + nr) * N_t + tx;
        H_MMSE[2 * o] = re[k]; H_MMSE[2 * o + 1] = im[k];
    }
}

// ---- per-subcarrier solve(H^H H + reg I, H^H Y) / power_scale ---------------
template <typename T>
__global__ void __launch_bounds__(128)
equalize_kernel(const T *__restrict__ Y, const T *__restrict__ H, const int *__restrict__ h_index, int B,
                int N, int N_r, int N_t, const T *__restrict__ reg, int reg_stride,
                const T *__restrict__ ps, int ps_stride, T *__restrict__ X_hat) {
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * N) return;
    const int b = (int)(gid / N), k = (int)(gid % N);
    const int hb = h_index ? h_index[b] : b;
    const T *Hk = H + (((size_t)hb * N + k) * N_r) * N_t * 2;
    const T *Yk = Y + ((size_t)b * N + k) * N_r * 2;
    T Gr[MAXA][MAXA], Gi[MAXA][MAXA], br[MAXA], bi[MAXA];
    for (int i = 0; i < N_t; ++i) {
        T sr = 0, si = 0;
        for (int r = 0; r < N_r; ++r) {                    // (H^H Y)_i = sum_r conj(H[r][i]) Y[r]
            const T hr = Hk[(r * N_t + i) * 2], hi = Hk[(r * N_t + i) * 2 + 1];
            const T yr = Yk[2 * r], yi = Yk[2 * r + 1];
            sr += hr * yr + hi * yi; si += hr * yi - hi * yr;
        }
        br[i] = sr; bi[i] = si;
        for (int j = 0; j < N_t; ++j) {                    // G_ij = sum_r conj(H[r][i]) H[r][j]
            T gr = 0, gi = 0;
            for (int r = 0; r < N_r; ++r) {
                const T ar = Hk[(r * N_t + i) * 2], ai = Hk[(r * N_t + i) * 2 + 1];
                const T cr = Hk[(r * N_t + j) * 2], ci = Hk[(r * N_t + j) * 2 + 1];
                gr += ar * cr + ai * ci; gi += ar * ci - ai * cr;
            }
            Gr[i][j] = gr; Gi[i][j] = gi;
        }
        Gr[i][i] += reg[(size_t)b * reg_stride];
    }
    // Gaussian elimination with partial pivoting (what LAPACK gesv does)
    for (int c = 0; c < N_t; ++c) {
        int piv = c; T best = Gr[c][c] * Gr[c][c] + Gi[c][c] * Gi[c][c];
        for (int r = c + 1; r < N_t; ++r) {
            const T v = Gr[r][c] * Gr[r][c] + Gi[r][c] * Gi[r][c];
            if (v > best) { best = v; piv = r; }
        }
        if (piv != c) {
            for (int j = 0; j < N_t; ++j) {
                T t = Gr[c][j]; Gr[c][j] = Gr[piv][j]; Gr[piv][j] = t;
                t = Gi[c][j]; Gi[c][j] = Gi[piv][j]; Gi[piv][j] = t;
            }
            T t = br[c]; br[c] = br[piv]; br[piv] = t;
            t = bi[c]; bi[c] = bi[piv]; bi[piv] = t;
        }
        const T inv = (T)1 / best, pr = Gr[c][c] * inv, pi_ = -Gi[c][c] * inv;   // 1/pivot
        for (int r = c + 1; r < N_t; ++r) {
            const T fr = Gr[r][c] * pr - Gi[r][c] * pi_, fi = Gr[r][c] * pi_ + Gi[r][c] * pr;
            for (int j = c; j < N_t; ++j) {
                Gr[r][j] -= fr * Gr[c][j] - fi * Gi[c][j];
                Gi[r][j] -= fr * Gi[c][j] + fi * Gr[c][j];
            }
            br[r] -= fr * br[c] - fi * bi[c];
            bi[r] -= fr * bi[c] + fi * br[c];
        }
    }
    const T inv_ps = (T)1 / ps[(size_t)b * ps_stride];
    T xr[MAXA], xi[MAXA];
    for (int c = N_t - 1; c >= 0; --c) {
        T sr = br[c], si = bi[c];
        for (int j = c + 1; j < N_t; ++j) {
            sr -= Gr[c][j] * xr[j] - Gi[c][j] * xi[j];
            si -= Gr[c][j] * xi[j] + Gi[c][j] * xr[j];
        }
        const T den = Gr[c][c] * Gr[c][c] + Gi[c][c] * Gi[c][c];
        xr[c] = (sr * Gr[c][c] + si * Gi[c][c]) / den;
        xi[c] = (si * Gr[c][c] - sr * Gi[c][c]) / den;
    }
    T *out = X_hat + ((size_t)b * N + k) * N_t * 2;
    for (int c = 0; c < N_t; ++c) { out[2 * c] = xr[c] * inv_ps; out[2 * c + 1] = xi[c] * inv_ps; }
}

// The same solve for the antenna counts the demos use, everything in registers: thread = (frame, subcarrier);
// its N_r x N_t block of H (64 contiguous numbers at 4x8) and its N_r entries of Y are read with 16-byte
// loads (adjacent threads read adjacent blocks, so every 128-byte line is consumed within the warp and the
// frames of a coherence block share H through L1 / L2), the Gram matrix is built from its upper triangle,
// and the row exchanges of the partial pivoting are predicated swaps -- no runtime-indexed array, so
// nothing spills to local memory (the generic kernel above ran 10.7 ms for 9472 frames of the 4x8 link,
// 0.7 % of the HBM rate; this one is bound by the loads).
template <typename T, int NT, int NR>
__global__ void __launch_bounds__(128)
equalize_fixed_kernel(const T *__restrict__ Y, const T *__restrict__ H, const int *__restrict__ h_index, int B,
                      int N, const T *__restrict__ reg, int reg_stride, const T *__restrict__ ps, int ps_stride,
                      T *__restrict__ X_hat) {
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (gid >= (size_t)B * N) return;
    const int b = (int)(gid / N), k = (int)(gid % N);
    const int hb = h_index ? h_index[b] : b;
    constexpr int HN = NR * NT * 2, YN = NR * 2, XN = NT * 2;
    constexpr int VEC = 16 / (int)sizeof(T);                       // numbers per 16-byte load
    T h[HN], y[YN];
    const T *Hk = H + ((size_t)hb * N + k) * HN;
    const T *Yk = Y + ((size_t)b * N + k) * YN;
    if constexpr (HN % VEC == 0) {
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
#pragma unroll
        for (int i = 0; i < HN / VEC; ++i) {
            const V v = __ldg(reinterpret_cast<const V *>(Hk) + i);
            if constexpr (sizeof(T) == 4) { h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = ((const float4 &)v).z; h[4 * i + 3] = ((const float4 &)v).w; }
            else { h[2 * i] = v.x; h[2 * i + 1] = v.y; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < HN; ++i) h[i] = Hk[i];
    }
    if constexpr (YN % VEC == 0) {
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
#pragma unroll
        for (int i = 0; i < YN / VEC; ++i) {
            const V v = __ldg(reinterpret_cast<const V *>(Yk) + i);
            if constexpr (sizeof(T) == 4) { y[4 * i] = v.x; y[4 * i + 1] = v.y; y[4 * i + 2] = ((const float4 &)v).z; y[4 * i + 3] = ((const float4 &)v).w; }
            else { y[2 * i] = v.x; y[2 * i + 1] = v.y; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < YN; ++i) y[i] = Yk[i];
    }
    T Gr[NT][NT], Gi[NT][NT], br[NT], bi[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        T sr = 0, si = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {                             // (H^H Y)_i = sum_r conj(H[r][i]) Y[r]
            const T hr = h[(r * NT + i) * 2], hi = h[(r * NT + i) * 2 + 1];
            sr += hr * y[2 * r] + hi * y[2 * r + 1]; si += hr * y[2 * r + 1] - hi * y[2 * r];
        }
        br[i] = sr; bi[i] = si;
#pragma unroll
        for (int j = i; j < NT; ++j) {                             // G_ij = sum_r conj(H[r][i]) H[r][j], G_ji = conj(G_ij)
            T gr = 0, gi = 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const T ar = h[(r * NT + i) * 2], ai = h[(r * NT + i) * 2 + 1];
                const T cr = h[(r * NT + j) * 2], ci = h[(r * NT + j) * 2 + 1];
                gr += ar * cr + ai * ci; gi += ar * ci - ai * cr;
            }
            Gr[i][j] = gr; Gi[i][j] = gi;
            Gr[j][i] = gr; Gi[j][i] = -gi;
        }
        Gi[i][i] = 0;
        Gr[i][i] += reg[(size_t)b * reg_stride];
    }
    // Gaussian elimination with partial pivoting (what LAPACK gesv does); exchanges as predicated swaps
#pragma unroll
    for (int c = 0; c < NT; ++c) {
        int piv = c; T best = Gr[c][c] * Gr[c][c] + Gi[c][c] * Gi[c][c];
#pragma unroll
        for (int r = c + 1; r < NT; ++r) {
            const T v = Gr[r][c] * Gr[r][c] + Gi[r][c] * Gi[r][c];
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = c + 1; r < NT; ++r) {
            const bool sw = piv == r;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const T a = Gr[c][j], d = Gr[r][j], e = Gi[c][j], f = Gi[r][j];
                Gr[c][j] = sw ? d : a; Gr[r][j] = sw ? a : d;
                Gi[c][j] = sw ? f : e; Gi[r][j] = sw ? e : f;
            }
            const T a = br[c], d = br[r], e = bi[c], f = bi[r];
            br[c] = sw ? d : a; br[r] = sw ? a : d;
            bi[c] = sw ? f : e; bi[r] = sw ? e : f;
        }
        const T inv = (T)1 / best, pr = Gr[c][c] * inv, pi_ = -Gi[c][c] * inv;   // 1/pivot
#pragma unroll
        for (int r = c + 1; r < NT; ++r) {
            const T fr = Gr[r][c] * pr - Gi[r][c] * pi_, fi = Gr[r][c] * pi_ + Gi[r][c] * pr;
#pragma unroll
            for (int j = c; j < NT; ++j) {
                Gr[r][j] -= fr * Gr[c][j] - fi * Gi[c][j];
                Gi[r][j] -= fr * Gi[c][j] + fi * Gr[c][j];
            }
            br[r] -= fr * br[c] - fi * bi[c];
            bi[r] -= fr * bi[c] + fi * br[c];
        }
    }
    const T inv_ps = (T)1 / ps[(size_t)b * ps_stride];
    T xr[NT], xi[NT], out[XN];
#pragma unroll
    for (int c = NT - 1; c >= 0; --c) {
        T sr = br[c], si = bi[c];
#pragma unroll
        for (int j = c + 1; j < NT; ++j) {
            sr -= Gr[c][j] * xr[j] - Gi[c][j] * xi[j];
            si -= Gr[c][j] * xi[j] + Gi[c][j] * xr[j];
        }
        const T den = Gr[c][c] * Gr[c][c] + Gi[c][c] * Gi[c][c];
        xr[c] = (sr * Gr[c][c] + si * Gi[c][c]) / den;
        xi[c] = (si * Gr[c][c] - sr * Gi[c][c]) / den;
        out[2 * c] = xr[c] * inv_ps; out[2 * c + 1] = xi[c] * inv_ps;
    }
    T *dst = X_hat + ((size_t)b * N + k) * XN;
    if constexpr (XN % VEC == 0) {
        using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
#pragma unroll
        for (int i = 0; i < XN / VEC; ++i) {
            V v;
            if constexpr (sizeof(T) == 4) { float4 t = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]); v = (V &)t; }
            else { double2 t = make_double2(out[2 * i], out[2 * i + 1]); v = (V &)t; }
            reinterpret_cast<V *>(dst)[i] = v;
        }
    } else {
#pragma unroll
        for (int i = 0; i < XN; ++i) dst[i] = out[i];
    }
}

// The fp32 solve when consecutive frames share their channel estimate (the frames of a coherence block):
// thread = (run of FR consecutive frames, subcarrier).  The Gram matrix and its LU factors (same arithmetic,
// same pivoting as above) are rebuilt only when the estimate or the regulariser changes from one frame to
// the next; every frame then costs its 64 bytes of Y, the H^H y product, the replay of the row exchanges and
// multipliers on the right-hand side, and the back substitution -- the results equal the per-frame solve bit
// for bit.  The next frame's Y is in flight while the current one is solved.
template <int NT, int NR, int FR>
__global__ void __launch_bounds__(128)
equalize_run_kernel(const float *__restrict__ Y, const float *__restrict__ H, const int *__restrict__ h_index, int B,
                    int N, const float *__restrict__ reg, int reg_stride, const float *__restrict__ ps,
                    int ps_stride, float *__restrict__ X_hat) {
    const size_t gid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int runs = (B + FR - 1) / FR;
    if (gid >= (size_t)runs * N) return;
    const int b0 = (int)(gid / N) * FR, k = (int)(gid % N);
    constexpr int HN = NR * NT * 2, YN = NR * 2, XN = NT * 2;
    static_assert(HN % 4 == 0 && YN % 4 == 0 && XN % 4 == 0, "16-byte rows");
    float h[HN], Ur[NT][NT], Ui[NT][NT], Mr[NT][NT], Mi[NT][NT];
    int pv[NT];
    float4 yn[YN / 4], yn2[YN / 4];                               // the next two frames' Y, in flight
    const int nfr = min(FR, B - b0);
#pragma unroll
    for (int i = 0; i < YN / 4; ++i) {
        yn[i] = __ldcs(reinterpret_cast<const float4 *>(Y + ((size_t)b0 * N + k) * YN) + i);
        yn2[i] = __ldcs(reinterpret_cast<const float4 *>(Y + ((size_t)(b0 + (nfr > 1)) * N + k) * YN) + i);
    }
    int hb_prev = -1;
    float rg_prev = 0.f;
    for (int fi = 0; fi < nfr; ++fi) {
        const int b = b0 + fi;
        float y[YN];
#pragma unroll
        for (int i = 0; i < YN / 4; ++i) {
            y[4 * i] = yn[i].x; y[4 * i + 1] = yn[i].y; y[4 * i + 2] = yn[i].z; y[4 * i + 3] = yn[i].w;
            yn[i] = yn2[i];
        }
        if (fi + 2 < nfr) {
#pragma unroll
            for (int i = 0; i < YN / 4; ++i)
                yn2[i] = __ldcs(reinterpret_cast<const float4 *>(Y + ((size_t)(b + 2) * N + k) * YN) + i);
        }
        const int hb = h_index ? h_index[b] : b;
        const float rg = reg[(size_t)b * reg_stride];
        if (fi == 0 || hb != hb_prev || rg != rg_prev) {
            hb_prev = hb; rg_prev = rg;
            const float4 *Hk = reinterpret_cast<const float4 *>(H + ((size_t)hb * N + k) * HN);
#pragma unroll
            for (int i = 0; i < HN / 4; ++i) {
                const float4 v = __ldg(Hk + i);
                h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < NT; ++i) {
#pragma unroll
                for (int j = i; j < NT; ++j) {                     // G_ij = sum_r conj(H[r][i]) H[r][j], G_ji = conj(G_ij)
                    float gr = 0, gi = 0;
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const float ar = h[(r * NT + i) * 2], ai = h[(r * NT + i) * 2 + 1];
                        const float cr = h[(r * NT + j) * 2], ci = h[(r * NT + j) * 2 + 1];
                        gr += ar * cr + ai * ci; gi += ar * ci - ai * cr;
                    }
                    Ur[i][j] = gr; Ui[i][j] = gi;
                    Ur[j][i] = gr; Ui[j][i] = -gi;
                }
                Ui[i][i] = 0;
                Ur[i][i] += rg;
            }
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                int piv = c; float best = Ur[c][c] * Ur[c][c] + Ui[c][c] * Ui[c][c];
#pragma unroll
                for (int r = c + 1; r < NT; ++r) {
                    const float v = Ur[r][c] * Ur[r][c] + Ui[r][c] * Ui[r][c];
                    if (v > best) { best = v; piv = r; }
                }
                pv[c] = piv;
#pragma unroll
                for (int r = c + 1; r < NT; ++r) {
                    const bool sw = piv == r;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const float a = Ur[c][j], d = Ur[r][j], e = Ui[c][j], f = Ui[r][j];
                        Ur[c][j] = sw ? d : a; Ur[r][j] = sw ? a : d;
                        Ui[c][j] = sw ? f : e; Ui[r][j] = sw ? e : f;
                    }
                }
                const float inv = 1.0f / best, pr = Ur[c][c] * inv, pi_ = -Ui[c][c] * inv;   // 1/pivot
#pragma unroll
                for (int r = c + 1; r < NT; ++r) {
                    const float fr = Ur[r][c] * pr - Ui[r][c] * pi_, fi_ = Ur[r][c] * pi_ + Ui[r][c] * pr;
                    Mr[r][c] = fr; Mi[r][c] = fi_;
#pragma unroll
                    for (int j = c; j < NT; ++j) {
                        Ur[r][j] -= fr * Ur[c][j] - fi_ * Ui[c][j];
                        Ui[r][j] -= fr * Ui[c][j] + fi_ * Ur[c][j];
                    }
                }
            }
        }
        float br[NT], bi[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            float sr = 0, si = 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {                         // (H^H Y)_i = sum_r conj(H[r][i]) Y[r]
                const float hr = h[(r * NT + i) * 2], hi = h[(r * NT + i) * 2 + 1];
                sr += hr * y[2 * r] + hi * y[2 * r + 1]; si += hr * y[2 * r + 1] - hi * y[2 * r];
            }
            br[i] = sr; bi[i] = si;
        }
#pragma unroll
        for (int c = 0; c < NT; ++c) {                             // the exchanges and multipliers of column c
#pragma unroll
            for (int r = c + 1; r < NT; ++r) {
                const bool sw = pv[c] == r;
                const float a = br[c], d = br[r], e = bi[c], f = bi[r];
                br[c] = sw ? d : a; br[r] = sw ? a : d;
                bi[c] = sw ? f : e; bi[r] = sw ? e : f;
            }
#pragma unroll
            for (int r = c + 1; r < NT; ++r) {
                const float fr = Mr[r][c], fi_ = Mi[r][c];
                const float t = br[r] - (fr * br[c] - fi_ * bi[c]);
                bi[r] -= fr * bi[c] + fi_ * br[c];
                br[r] = t;
            }
        }
        const float inv_ps = 1.0f / ps[(size_t)b * ps_stride];
        float xr[NT], xi[NT], out[XN];
#pragma unroll
        for (int c = NT - 1; c >= 0; --c) {
            float sr = br[c], si = bi[c];
#pragma unroll
            for (int j = c + 1; j < NT; ++j) {
                sr -= Ur[c][j] * xr[j] - Ui[c][j] * xi[j];
                si -= Ur[c][j] * xi[j] + Ui[c][j] * xr[j];
            }
            const float den = Ur[c][c] * Ur[c][c] + Ui[c][c] * Ui[c][c];
            xr[c] = (sr * Ur[c][c] + si * Ui[c][c]) / den;
            xi[c] = (si * Ur[c][c] - sr * Ui[c][c]) / den;
            out[2 * c] = xr[c] * inv_ps; out[2 * c + 1] = xi[c] * inv_ps;
        }
        float4 *dst = reinterpret_cast<float4 *>(X_hat + ((size_t)b * N + k) * XN);
#pragma unroll
        for (int i = 0; i < XN / 4; ++i)
            __stcs(dst + i, make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]));
    }
}

template <typename T>
static bool launch_equalize_fixed(const T *Y, const T *H, const int *h_index, int B, int N, int N_r, int N_t,
                                  const T *reg, int reg_stride, const T *ps, int ps_stride, T *X, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        // a frame -> estimate table (frames of a coherence block share H): runs of 16 frames per thread
        constexpr int FR = 16;
        const int rblocks = (int)(((size_t)((B + FR - 1) / FR) * N + 127) / 128);
#define ESN_EQ_RUN(NT_, NR_)                                                                                      \
        if (N_t == NT_ && N_r == NR_) {                                                                           \
            equalize_run_kernel<NT_, NR_, FR><<<rblocks, 128, 0, st>>>(Y, H, h_index, B, N, reg, reg_stride, ps,  \
                                                                       ps_stride, X);                             \
            return true;                                                                                          \
        }
        if (h_index && B >= 4 * FR) { ESN_EQ_RUN(2, 2) ESN_EQ_RUN(2, 4) ESN_EQ_RUN(4, 4) ESN_EQ_RUN(4, 8) }
#undef ESN_EQ_RUN
    }
    const int blocks = (int)(((size_t)B * N + 127) / 128);
#define ESN_EQ_CASE(NT_, NR_)                                                                                     \
    if (N_t == NT_ && N_r == NR_) {                                                                               \
        equalize_fixed_kernel<T, NT_, NR_><<<blocks, 128, 0, st>>>(Y, H, h_index, B, N, reg, reg_stride, ps,      \
                                                                   ps_stride, X);                                 \
        return true;                                                                                              \
    }
    ESN_EQ_CASE(1, 1) ESN_EQ_CASE(1, 2) ESN_EQ_CASE(2, 2) ESN_EQ_CASE(2, 4) ESN_EQ_CASE(4, 4) ESN_EQ_CASE(4, 8)
#undef ESN_EQ_CASE
    return false;
}

template <typename T>
__global__ void demap_count_kernel(const T *__restrict__ X_hat, size_t total, int qam_bits,
                                   uint8_t *__restrict__ idx, const uint8_t *__restrict__ tx_idx, T eps,
                                   unsigned long long *__restrict__ counts) {
    const Slicer<T> sl(qam_bits);
    unsigned long long errs = 0, near = 0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const T xr = X_hat[2 * e], xi = X_hat[2 * e + 1];
        const int id = sl.index(xr, xi);
        if (idx) idx[e] = (uint8_t)id;
        if (tx_idx) errs += __popc((unsigned)(id ^ (int)tx_idx[e]));
        if (eps > (T)0 && sl.boundary_dist(xr, xi) < eps) near += 1;
    }
    block_add_counts(errs, near, counts);
}

// Vector version for 16-byte aligned buffers: four symbols per thread and iteration (32 bytes of X_hat in two
// 16-byte loads, the four reference indices as one word, the four decisions stored as one word), two
// iterations in flight.  The scalar kernel above keeps 2-3 small loads in flight per thread and runs at a
// third of the HBM rate.
template <typename T>
__global__ void __launch_bounds__(256)
demap_count_vec_kernel(const T *__restrict__ X_hat, size_t total, int qam_bits, uint8_t *__restrict__ idx,
                       const uint8_t *__restrict__ tx_idx, T eps, unsigned long long *__restrict__ counts) {
    const Slicer<T> sl(qam_bits);
    unsigned long long errs = 0, near = 0;
    const size_t n4 = total / 4, stride = (size_t)gridDim.x * blockDim.x;
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    constexpr int NV = 8 * (int)sizeof(T) / 16;                        // 16-byte loads per four symbols
    const V *xv = reinterpret_cast<const V *>(X_hat);
    const uint32_t *tv = reinterpret_cast<const uint32_t *>(tx_idx);
    uint32_t *iv = reinterpret_cast<uint32_t *>(idx);
#pragma unroll 2
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n4; g += stride) {
        V raw[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) raw[i] = __ldg(xv + g * NV + i);
        const uint32_t t4 = tx_idx ? __ldg(tv + g) : 0u;
        const T *v = reinterpret_cast<const T *>(raw);
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int id = sl.index(v[2 * j], v[2 * j + 1]);
            out |= (uint32_t)id << (8 * j);
            if (eps > (T)0 && sl.boundary_dist(v[2 * j], v[2 * j + 1]) < eps) near += 1;
        }
        if (idx) iv[g] = out;
        if (tx_idx) errs += __popc(out ^ t4);
    }
    if (blockIdx.x == 0)                                               // up to three symbols left over
        for (size_t e = n4 * 4 + threadIdx.x; e < total; e += blockDim.x) {
            const T xr = X_hat[2 * e], xi = X_hat[2 * e + 1];
            const int id = sl.index(xr, xi);
            if (idx) idx[e] = (uint8_t)id;
            if (tx_idx) errs += __popc((unsigned)(id ^ (int)tx_idx[e]));
            if (eps > (T)0 && sl.boundary_dist(xr, xi) < eps) near += 1;
        }
    block_add_counts(errs, near, counts);
}

// ---- workload generation on the device ------------------
// One CTA per frame: symbol indices -> QAM -> N*IFFT -> CP -> *sqrt(Pi) -> soft PA clip
// x/sqrt(1+(|x|/A)^2) -> per-link FIR channel -> + AWGN -> received samples y_cp, and the
// real-valued ESN input rows (re/im interleaved per antenna, `delay` trailing zero rows).
// Reference: system_model_2/OFDM_MIMO_2-2_NBF_LDPC.py:402-426, :430-433.  Gaussian noise comes
// from `noise` [B][N+CP][N_r] complex standard normals when given (parity runs), else from the
// counter hash (Box-Muller) keyed by (seed, frame, sample, antenna).
template <typename T>
__global__ void synth_frames_kernel(const uint8_t *__restrict__ tx_idx, const T *__restrict__ taps,
                                    const int *__restrict__ chan_index, const T *__restrict__ Pi,
                                    const T *__restrict__ A_clip, const T *__restrict__ noise, T noise_std,
                                    unsigned long long seed, int N, int cp, int N_t, int N_r, int ntaps,
                                    int qam_bits, int delay, T *__restrict__ x_cp_out, T *__restrict__ y_cp,
                                    T *__restrict__ esn_in) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int L = N + cp, b = blockIdx.x, logn = ilog2(N);
    T *re = reinterpret_cast<T *>(sm), *im = re + N, *twr = im + N, *twi = twr + N / 2;
    T *xr = twi + N / 2, *xi = xr + (size_t)N_t * L;          // clipped Tx samples [N_t][L]
    fft_make_twiddles(twr, twi, N);
    const int side = 1 << (qam_bits / 2);
    const T cs = (T)1 / sqrt_t((T)(2.0 * (side * side - 1) / 3.0));
    const T sp = sqrt_t(Pi[b]), A = A_clip[b];
    for (int tx = 0; tx < N_t; ++tx) {
        __syncthreads();
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            const int id = tx_idx[((size_t)b * N + k) * N_t + tx];
            const int r = bitrev(k, logn);
            // index = side*i_re + i_im; 255 = empty subcarrier (the comb pilot of the LS estimator)
            re[r] = id == 255 ? (T)0 : (T)(2 * (id / side) - (side - 1)) * cs;
            im[r] = id == 255 ? (T)0 : (T)(2 * (id % side) - (side - 1)) * cs;
        }
        fft_inplace(re, im, twr, twi, N, logn, true);             // = N * ifft(X)
        for (int t = threadIdx.x; t < L; t += blockDim.x) {
            const int src = t < cp ? N - cp + t : t - cp;
            const T vr = re[src] * sp, vi = im[src] * sp;
            if (x_cp_out) {
                const size_t o = (((size_t)b * L + t) * N_t + tx) * 2;
                x_cp_out[o] = vr; x_cp_out[o + 1] = vi;
            }
            const T m2 = (vr * vr + vi * vi) / (A * A);
            const T g = (T)1 / sqrt_t((T)1 + m2);
            xr[tx * L + t] = vr * g; xi[tx * L + t] = vi * g;
        }
    }
    __syncthreads();
    const T *cb = taps + (size_t)(chan_index ? chan_index[b] : b) * N_r * N_t * ntaps * 2;
    const int rows = L + delay;
    for (int e = threadIdx.x; e < rows * N_r; e += blockDim.x) {
        const int t = e / N_r, rx = e - t * N_r;
        T yr = 0, yi = 0;
        if (t < L) {
            for (int tx = 0; tx < N_t; ++tx) {
                const T *c = cb + ((size_t)(rx * N_t + tx) * ntaps) * 2;
                for (int k = 0; k < ntaps && k <= t; ++k) {
                    const T ar = c[2 * k], ai = c[2 * k + 1];
                    const T br = xr[tx * L + t - k], bi = xi[tx * L + t - k];
                    yr += ar * br - ai * bi; yi += ar * bi + ai * br;
                }
            }
            T nr, ni;
            if (noise) {
                const size_t o = (((size_t)b * L + t) * N_r + rx) * 2;
                nr = noise[o]; ni = noise[o + 1];
            } else {
                const uint32_t key = esn_noise_key(seed ^ 0xA5A5A5A5ULL, (uint32_t)b, (uint32_t)t);
                const uint32_t h1 = esn_mix32(key + (uint32_t)rx * 0xC2B2AE35U);
                const uint32_t h2 = esn_mix32(h1 ^ 0x68E31DA4U);
                const T u1 = ((T)(h1 >> 8) + (T)0.5) * (T)(1.0 / 16777216.0);
                const T u2 = (T)(h2 >> 8) * (T)(1.0 / 16777216.0);
                const T rad = sqrt_t((T)-2 * log_t(u1));
                T sn, cn;
                sincospi_t((T)2 * u2, &sn, &cn);
                nr = rad * cn; ni = rad * sn;
            }
            yr += noise_std * nr; yi += noise_std * ni;
            if (y_cp) {
                const size_t o = (((size_t)b * L + t) * N_r + rx) * 2;
                y_cp[o] = yr; y_cp[o + 1] = yi;
            }
        }
        if (esn_in) {
            const size_t o = ((size_t)b * rows + t) * 2 * N_r + 2 * rx;
            esn_in[o] = yr; esn_in[o + 1] = yi;
        }
    }
}

// Whole-frame version (used whenever the N_t streams fit in shared memory together, i.e. at every demo
// size): the frame's symbol indices are read as one contiguous block, the N_t IFFTs run side by side with
// three radix-2 layers per shared-memory pass (IFFT(X) = conj(FFT(conj(X))), so the forward passes of the
// unpack kernel are reused), the channel taps sit in shared memory, and a thread filters TB consecutive
// samples of one Rx antenna so that a tap is read once per TB products.  ncu on the first version: 44 % of
// the stall samples in the nine-barrier radix-2 IFFTs, 42 % in the FIR fetching its taps from global memory.
template <typename T>
__global__ void __launch_bounds__(256)
synth_frames_frame_kernel(const uint8_t *__restrict__ tx_idx, const T *__restrict__ taps,
                          const int *__restrict__ chan_index, const T *__restrict__ Pi,
                          const T *__restrict__ A_clip, const T *__restrict__ noise, T noise_std,
                          unsigned long long seed, int N, int cp, int N_t, int N_r, int ntaps,
                          int qam_bits, int delay, T *__restrict__ x_cp_out, T *__restrict__ y_cp,
                          T *__restrict__ esn_in) {
    extern __shared__ __align__(16) unsigned char sm[];
    constexpr int TB = 4;
    const int L = N + cp, b = blockIdx.x, logn = ilog2(N), rows = L + delay;
    const int tot = N * N_t, tots = tot + (tot >> 5) + 1, ntap_all = N_r * N_t * ntaps;
    T *re = reinterpret_cast<T *>(sm), *im = re + tots, *twr = im + tots, *twi = twr + N / 2;
    T *xr = twi + N / 2, *xi = xr + (size_t)N_t * L;          // clipped Tx samples [N_t][L]
    // channel taps [N_r][N_t][tstr]: an odd row length, so that the rows of the N_r antennas a warp reads in the FIR fall
    // into different banks (rows of 8 taps x 4 Tx put all of them into one: 8-way conflicts on every tap load, 71 % of the
    // kernel's shared-memory wavefronts in profiles/r2_ncu_chain_kernels.txt)
    const int tstr = ntaps | 1;
    T *tr = xi + (size_t)N_t * L, *ti = tr + N_r * N_t * tstr;
    fft_make_twiddles(twr, twi, N);
    const T *cb = taps + (size_t)(chan_index ? chan_index[b] : b) * ntap_all * 2;
    for (int e = threadIdx.x; e < ntap_all; e += blockDim.x) {
        const int o = (e / ntaps) * tstr + e % ntaps;
        tr[o] = cb[2 * e]; ti[o] = cb[2 * e + 1];
    }
    const int side = 1 << (qam_bits / 2);
    const T cs = (T)1 / sqrt_t((T)(2.0 * (side * side - 1) / 3.0));
    const T sp = sqrt_t(Pi[b]), A = A_clip[b];
    const uint8_t *ib = tx_idx + (size_t)b * tot;
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {              // e = k N_t + tx: contiguous
        const int k = e / N_t, tx = e - k * N_t, id = ib[e];
        const int o = skew(tx * N + bitrev(k, logn));
        // index = side*i_re + i_im; 255 = empty subcarrier (the comb pilot of the LS estimator); conjugated
        re[o] = id == 255 ? (T)0 : (T)(2 * (id / side) - (side - 1)) * cs;
        im[o] = id == 255 ? (T)0 : -(T)(2 * (id % side) - (side - 1)) * cs;
    }
    __syncthreads();
    fft_batched_radix8(re, im, twr, twi, N, logn, N_t);               // conj(result) = N * ifft(X)
    const T inv_A2 = (T)1 / (A * A);
    for (int e = threadIdx.x; e < N_t * L; e += blockDim.x) {          // e = t N_t + tx: contiguous x_cp rows
        const int t = e / N_t, tx = e - t * N_t;
        const int src = t < cp ? N - cp + t : t - cp;
        const T vr = re[skew(tx * N + src)] * sp, vi = -im[skew(tx * N + src)] * sp;
        if (x_cp_out) {
            const size_t o = ((size_t)b * L * N_t + e) * 2;
            x_cp_out[o] = vr; x_cp_out[o + 1] = vi;
        }
        const T g = (T)1 / sqrt_t((T)1 + (vr * vr + vi * vi) * inv_A2);
        xr[tx * L + t] = vr * g; xi[tx * L + t] = vi * g;
    }
    __syncthreads();
    const int nblk = (L + TB - 1) / TB;
    for (int e = threadIdx.x; e < nblk * N_r; e += blockDim.x) {
        const int tb = e / N_r, rx = e - tb * N_r, t0 = tb * TB;
        T yr[TB], yi[TB];
#pragma unroll
        for (int j = 0; j < TB; ++j) { yr[j] = 0; yi[j] = 0; }
        for (int tx = 0; tx < N_t; ++tx) {
            const T *c_r = tr + (rx * N_t + tx) * tstr, *c_i = ti + (rx * N_t + tx) * tstr;
            const T *x_r = xr + tx * L, *x_i = xi + tx * L;
            if (ntaps <= 8) {
                // sliding window in registers: the TB + 7 samples the block needs are loaded once per Tx stream
                constexpr int WIN = TB + 7;
                T wr[WIN], wi[WIN];
#pragma unroll
                for (int q = 0; q < WIN; ++q) {
                    const int t = t0 - 7 + q;
                    const bool ok = t >= 0 && t < L;
                    wr[q] = ok ? x_r[t] : (T)0; wi[q] = ok ? x_i[t] : (T)0;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < ntaps) {
                        const T ar = c_r[k], ai = c_i[k];
#pragma unroll
                        for (int j = 0; j < TB; ++j) {
                            const T br = wr[7 + j - k], bi = wi[7 + j - k];
                            yr[j] += ar * br - ai * bi; yi[j] += ar * bi + ai * br;
                        }
                    }
                }
            } else {
                for (int k = 0; k < ntaps; ++k) {
                    const T ar = c_r[k], ai = c_i[k];
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        const int t = t0 + j - k;
                        if (t >= 0 && t0 + j < L) {
                            const T br = x_r[t], bi = x_i[t];
                            yr[j] += ar * br - ai * bi; yi[j] += ar * bi + ai * br;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            const int t = t0 + j;
            if (t >= L) break;
            T nr, ni;
            if (noise) {
                const size_t o = (((size_t)b * L + t) * N_r + rx) * 2;
                nr = noise[o]; ni = noise[o + 1];
            } else {
                const uint32_t key = esn_noise_key(seed ^ 0xA5A5A5A5ULL, (uint32_t)b, (uint32_t)t);
                const uint32_t h1 = esn_mix32(key + (uint32_t)rx * 0xC2B2AE35U);
                const uint32_t h2 = esn_mix32(h1 ^ 0x68E31DA4U);
                const T u1 = ((T)(h1 >> 8) + (T)0.5) * (T)(1.0 / 16777216.0);
                const T u2 = (T)(h2 >> 8) * (T)(1.0 / 16777216.0);
                const T rad = sqrt_t((T)-2 * log_t(u1));
                T sn, cn;
                sincospi_t((T)2 * u2, &sn, &cn);
                nr = rad * cn; ni = rad * sn;
            }
            const T vr = yr[j] + noise_std * nr, vi = yi[j] + noise_std * ni;
            if (y_cp) {
                const size_t o = (((size_t)b * L + t) * N_r + rx) * 2;
                y_cp[o] = vr; y_cp[o + 1] = vi;
            }
            if (esn_in) {
                const size_t o = ((size_t)b * rows + t) * 2 * N_r + 2 * rx;
                esn_in[o] = vr; esn_in[o + 1] = vi;
            }
        }
    }
    if (esn_in)                                                        // the `delay` trailing zero rows
        for (int e = threadIdx.x; e < delay * 2 * N_r; e += blockDim.x)
            esn_in[((size_t)b * rows + L) * 2 * N_r + e] = (T)0;
}

template <typename K>
inline int allow_smem(K kern, size_t smem) {
    if (smem > 48 * 1024)
        ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

inline bool pow2_ok(int N) { return N >= 2 && N <= MAX_FFT && (N & (N - 1)) == 0; }
inline int fft_threads(int N) { int t = N / 2; if (t < 32) t = 32; if (t > 512) t = 512; return t; }

}  // namespace

// grid of the persistent N = 512 kernels: the CTAs that stay resident (registers, shared memory) x SMs
template <typename K>
static int f512_grid(K kernel, int B, size_t smem) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return std::min(B, sms * per_sm);
}

template <typename T, int NV>
static int launch_unpack512(const void *y, int B, int rows, int N_t, const void *Pi, int pi_stride, int qam_bits,
                            void *X_hat, uint8_t *idx, const uint8_t *tx_idx, double eps, unsigned long long *counts,
                            size_t smem, cudaStream_t st) {
    auto kernel = unpack_fft_demap_frame512_kernel<T, NV>;
    if (int rc = allow_smem(kernel, smem)) return rc;
    kernel<<<f512_grid(kernel, B, smem), 256, smem, st>>>((const T *)y, B, rows, N_t, (const T *)Pi, pi_stride, qam_bits,
                                                         (T *)X_hat, idx, tx_idx, (T)eps, counts);
    return 0;
}

template <typename T, int NV>
static int launch_rxfft512(const void *y_cp, int B, int cp, int N_r, void *Y, size_t smem, cudaStream_t st) {
    auto kernel = rx_fft_frame512_kernel<T, NV>;
    if (int rc = allow_smem(kernel, smem)) return rc;
    kernel<<<f512_grid(kernel, B, smem), 256, smem, st>>>((const T *)y_cp, B, cp, N_r, (T *)Y);
    return 0;
}

extern "C" int ofdm_unpack_fft_demap(int dtype, const void *y, int B, int rows, int N, int N_t, const void *Pi,
                                     int pi_stride, int qam_bits, void *X_hat, uint8_t *idx,
                                     const uint8_t *tx_idx, double boundary_eps, unsigned long long *counts,
                                     void *stream) {
    if (!y || !Pi || B <= 0 || N_t <= 0 || rows < N || !pow2_ok(N)) return ESN_E_BADARG;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return ESN_E_BADARG;
    dim3 grid(N_t, B);
    cudaStream_t st = (cudaStream_t)stream;
    // whole-frame kernel when the N_t streams fit in shared memory together (always at the demo sizes)
    const size_t el = dtype == ESN_F64 ? sizeof(double) : sizeof(float);
    const size_t frame_smem = ((size_t)2 * ((size_t)N * N_t + (size_t)N * N_t / 32 + 1) + N) * el;
    if (N == 512 && N_t <= 8 && (dtype == ESN_F32 || dtype == ESN_F64)) {       // three radix-8 stages in registers
        const size_t smem512 = ((size_t)2 * N_t * F512_STR + 512) * 2 * el;
        int rc;
        if (dtype == ESN_F32)
            rc = N_t <= 4 ? launch_unpack512<float, 8>(y, B, rows, N_t, Pi, pi_stride, qam_bits, X_hat, idx, tx_idx, boundary_eps, counts, smem512, st)
                          : launch_unpack512<float, 16>(y, B, rows, N_t, Pi, pi_stride, qam_bits, X_hat, idx, tx_idx, boundary_eps, counts, smem512, st);
        else
            rc = N_t <= 4 ? launch_unpack512<double, 8>(y, B, rows, N_t, Pi, pi_stride, qam_bits, X_hat, idx, tx_idx, boundary_eps, counts, smem512, st)
                          : launch_unpack512<double, 16>(y, B, rows, N_t, Pi, pi_stride, qam_bits, X_hat, idx, tx_idx, boundary_eps, counts, smem512, st);
        if (rc) return rc;
        return esn_launch_status();
    }
    if (frame_smem <= 160 * 1024 && (dtype == ESN_F32 || dtype == ESN_F64)) {
        const int threads = std::min(256, std::max(64, (N / 8) * N_t));
        if (dtype == ESN_F32) {
            if (int rc = allow_smem(unpack_fft_demap_frame_kernel<float>, frame_smem)) return rc;
            unpack_fft_demap_frame_kernel<float><<<B, threads, frame_smem, st>>>(
                (const float *)y, rows, N, N_t, (const float *)Pi, pi_stride, qam_bits, (float *)X_hat, idx, tx_idx,
                (float)boundary_eps, counts);
        } else {
            if (int rc = allow_smem(unpack_fft_demap_frame_kernel<double>, frame_smem)) return rc;
            unpack_fft_demap_frame_kernel<double><<<B, threads, frame_smem, st>>>(
                (const double *)y, rows, N, N_t, (const double *)Pi, pi_stride, qam_bits, (double *)X_hat, idx, tx_idx,
                boundary_eps, counts);
        }
        return esn_launch_status();
    }
    if (int rc = dtype == ESN_F64 ? allow_smem(unpack_fft_demap_kernel<double>, 3 * (size_t)N * sizeof(double)) : 0) return rc;
    if (dtype == ESN_F32)
        unpack_fft_demap_kernel<float><<<grid, fft_threads(N), 3 * N * sizeof(float), st>>>(
            (const float *)y, rows, N, N_t, (const float *)Pi, pi_stride, qam_bits, (float *)X_hat, idx, tx_idx,
            (float)boundary_eps, counts);
    else if (dtype == ESN_F64)
        unpack_fft_demap_kernel<double><<<grid, fft_threads(N), 3 * N * sizeof(double), st>>>(
            (const double *)y, rows, N, N_t, (const double *)Pi, pi_stride, qam_bits, (double *)X_hat, idx, tx_idx,
            boundary_eps, counts);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_rx_fft(int dtype, const void *y_cp, int B, int N, int cp, int N_r, void *Y, void *stream) {
    if (!y_cp || !Y || B <= 0 || N_r <= 0 || cp < 0 || !pow2_ok(N)) return ESN_E_BADARG;
    dim3 grid(N_r, B);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tot = (size_t)N * N_r, esz = dtype == ESN_F64 ? sizeof(double) : sizeof(float);
    const size_t frame_smem = (2 * (tot + tot / 32 + 1) + N) * esz;
    if (N == 512 && N_r <= 8 && (dtype == ESN_F32 || dtype == ESN_F64)) {       // three radix-8 stages in registers
        const size_t smem512 = ((size_t)2 * N_r * F512_STR + 512) * 2 * esz;
        int rc;
        if (dtype == ESN_F32)
            rc = N_r <= 4 ? launch_rxfft512<float, 8>(y_cp, B, cp, N_r, Y, smem512, st) : launch_rxfft512<float, 16>(y_cp, B, cp, N_r, Y, smem512, st);
        else
            rc = N_r <= 4 ? launch_rxfft512<double, 8>(y_cp, B, cp, N_r, Y, smem512, st) : launch_rxfft512<double, 16>(y_cp, B, cp, N_r, Y, smem512, st);
        if (rc) return rc;
        return esn_launch_status();
    }
    if (frame_smem <= 160 * 1024 && (dtype == ESN_F32 || dtype == ESN_F64)) {
        const int threads = (int)std::min<size_t>(512, std::max<size_t>(64, tot / 8));
        if (dtype == ESN_F32) {
            if (int rc = allow_smem(rx_fft_frame_kernel<float>, frame_smem)) return rc;
            rx_fft_frame_kernel<float><<<B, threads, frame_smem, st>>>((const float *)y_cp, N, cp, N_r, (float *)Y);
        } else {
            if (int rc = allow_smem(rx_fft_frame_kernel<double>, frame_smem)) return rc;
            rx_fft_frame_kernel<double><<<B, threads, frame_smem, st>>>((const double *)y_cp, N, cp, N_r, (double *)Y);
        }
        return esn_launch_status();
    }
    if (int rc = dtype == ESN_F64 ? allow_smem(rx_fft_kernel<double>, 3 * (size_t)N * sizeof(double)) : 0) return rc;
    if (dtype == ESN_F32)
        rx_fft_kernel<float><<<grid, fft_threads(N), 3 * N * sizeof(float), st>>>((const float *)y_cp, N, cp, N_r, (float *)Y);
    else if (dtype == ESN_F64)
        rx_fft_kernel<double><<<grid, fft_threads(N), 3 * N * sizeof(double), st>>>((const double *)y_cp, N, cp, N_r, (double *)Y);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_chanest(int dtype, const void *Y_LS, const void *X_LS, int B, int N, int N_r, int N_t,
                            const void *Pi, const void *isi_magnitude, int taps, double No, void *H_LS,
                            void *H_MMSE, void *stream) {
    if (!Y_LS || !X_LS || !Pi || !isi_magnitude || !H_LS || !H_MMSE) return ESN_E_BADARG;
    if (B <= 0 || N_r <= 0 || N_t <= 0 || taps <= 0 || taps > N || !pow2_ok(N)) return ESN_E_BADARG;
    dim3 grid(N_t, N_r, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = dtype == ESN_F32 ? allow_smem(chanest_kernel<float>, 5 * (size_t)N * sizeof(float)) : 0) return rc;
    if (dtype == ESN_F32)
        chanest_kernel<float><<<grid, fft_threads(N), 5 * N * sizeof(float), st>>>(
            (const float *)Y_LS, (const float *)X_LS, N, N_r, N_t, (const float *)Pi, (const float *)isi_magnitude,
            taps, (float)No, (float *)H_LS, (float *)H_MMSE);
    else if (dtype == ESN_F64) {
        auto kern = chanest_kernel<double>;
        size_t smem = 5 * (size_t)N * sizeof(double);
        if (smem > 48 * 1024) ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, fft_threads(N), smem, st>>>(
            (const double *)Y_LS, (const double *)X_LS, N, N_r, N_t, (const double *)Pi,
            (const double *)isi_magnitude, taps, No, (double *)H_LS, (double *)H_MMSE);
    } else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_equalize(int dtype, const void *Y, const void *H, const int32_t *h_index, int B, int N,
                             int N_r, int N_t, const void *reg, int reg_stride, const void *power_scale,
                             int ps_stride, void *X_hat, void *stream) {
    if (!Y || !H || !reg || !power_scale || !X_hat) return ESN_E_BADARG;
    if (B <= 0 || N <= 0 || N_r <= 0 || N_t <= 0 || N_r > MAXA || N_t > MAXA) return ESN_E_BADARG;
    const size_t total = (size_t)B * N;
    const int blocks = (int)((total + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ESN_F32 && launch_equalize_fixed<float>((const float *)Y, (const float *)H, h_index, B, N, N_r, N_t,
                                                         (const float *)reg, reg_stride, (const float *)power_scale,
                                                         ps_stride, (float *)X_hat, st))
        return esn_launch_status();
    if (dtype == ESN_F64 && launch_equalize_fixed<double>((const double *)Y, (const double *)H, h_index, B, N, N_r, N_t,
                                                          (const double *)reg, reg_stride, (const double *)power_scale,
                                                          ps_stride, (double *)X_hat, st))
        return esn_launch_status();
    if (dtype == ESN_F32)
        equalize_kernel<float><<<blocks, 128, 0, st>>>((const float *)Y, (const float *)H, h_index, B, N, N_r, N_t,
                                                       (const float *)reg, reg_stride, (const float *)power_scale,
                                                       ps_stride, (float *)X_hat);
    else if (dtype == ESN_F64)
        equalize_kernel<double><<<blocks, 128, 0, st>>>((const double *)Y, (const double *)H, h_index, B, N, N_r, N_t,
                                                        (const double *)reg, reg_stride, (const double *)power_scale,
                                                        ps_stride, (double *)X_hat);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

// ---- soft outputs: sigma2 estimate + max-log LLRs (+ logistic calibration) ----------------
// One CTA per frame.  Pass 1: sigma2 = mean |X - nearest point|^2 + 1e-12 over the frame (the mean of
// the reference's per-Tx means).  Pass 2: for the square constellation with index L i_re + i_im and
// LSB-first binary labels the max-log LLR of a bit separates per axis: bits [0, m/2) belong to the
// imaginary level, bits [m/2, m) to the real level, and the other axis cancels in d1 - d0.
template <typename T>
__global__ void __launch_bounds__(256)
soft_demap_kernel(const T *__restrict__ X, int N, int N_t, int qam_bits, const double *__restrict__ cal_a,
                  const double *__restrict__ cal_b, T clip, T *__restrict__ sigma2, T *__restrict__ llr) {
    const Slicer<T> sl(qam_bits);
    const int b = blockIdx.x, n_sym = N * N_t, hb = qam_bits / 2;
    const T *Xf = X + (size_t)b * n_sym * 2;
    __shared__ double s_red[8];
    __shared__ double s_sigma;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_sym; i += blockDim.x) {
        const T re = Xf[2 * i], im = Xf[2 * i + 1];
        const T dr = re - (T)(2 * sl.level(re) - (sl.L - 1)) / sl.s, di = im - (T)(2 * sl.level(im) - (sl.L - 1)) / sl.s;
        acc += (double)(dr * dr + di * di);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
        s_sigma = t / n_sym + 1e-12;
        if (sigma2) sigma2[b] = (T)s_sigma;
    }
    __syncthreads();
    if (!llr) return;
    const T inv = (T)(1.0 / fmax(s_sigma, 1e-12));
    T *Lf = llr + (size_t)b * n_sym * qam_bits;
    for (int i = threadIdx.x; i < n_sym; i += blockDim.x) {
        const int n = i / N_t, tx = i - n * N_t;
        const T v[2] = {Xf[2 * i + 1], Xf[2 * i]};          // axis 0 = imaginary (low bits), axis 1 = real
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {
            for (int k = 0; k < hb; ++k) {
                T d0 = (T)1e30, d1 = (T)1e30;
                for (int j = 0; j < sl.L; ++j) {
                    const T e = v[ax] - (T)(2 * j - (sl.L - 1)) / sl.s, e2 = e * e;
                    if ((j >> k) & 1) d1 = fmin(d1, e2); else d0 = fmin(d0, e2);
                }
                const int bit = ax * hb + k;
                T l = (d1 - d0) * inv;
                if (cal_a) l = fmin(fmax(-((T)cal_a[bit] * l + (T)cal_b[bit]), -clip), clip);
                Lf[((size_t)n * qam_bits + bit) * N_t + tx] = l;
            }
        }
    }
}

// The same for fp32 frames of four streams (the 4x8 link): a thread owns a subcarrier -- its four symbols are two
// 16-byte loads, its LLR rows [bit][tx] are qam_bits contiguous 16-byte stores -- and the level values are divided
// out once per thread instead of once per candidate (the scalar version is issue-bound at 0.21 of the HBM rate).
__global__ void __launch_bounds__(256)
soft_demap_nt4_f32_kernel(const float *__restrict__ X, int N, int qam_bits, const double *__restrict__ cal_a,
                          const double *__restrict__ cal_b, float clip, float *__restrict__ sigma2, float *__restrict__ llr) {
    const Slicer<float> sl(qam_bits);
    const int b = blockIdx.x, n_sym = N * 4, hb = qam_bits / 2;
    const float4 *Xf = reinterpret_cast<const float4 *>(X + (size_t)b * n_sym * 2);
    __shared__ double s_red[8];
    __shared__ double s_sigma;
    float lv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lv[j] = (float)(2 * j - (sl.L - 1)) / sl.s;
    double acc = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float4 a = Xf[2 * n], c = Xf[2 * n + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float d = v[i] - (float)(2 * sl.level(v[i]) - (sl.L - 1)) / sl.s;
            t += d * d;
            if (i & 1) { acc += (double)t; t = 0.f; }            // per symbol, as the scalar kernel sums
        }
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
        s_sigma = t / n_sym + 1e-12;
        if (sigma2) sigma2[b] = (float)s_sigma;
    }
    __syncthreads();
    if (!llr) return;
    const float inv = (float)(1.0 / fmax(s_sigma, 1e-12));
    float ca[6], cb[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { ca[i] = (cal_a && i < qam_bits) ? (float)cal_a[i] : 0.f; cb[i] = (cal_b && i < qam_bits) ? (float)cal_b[i] : 0.f; }
    float4 *Lf = reinterpret_cast<float4 *>(llr + (size_t)b * n_sym * qam_bits);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float4 a = Xf[2 * n], c = Xf[2 * n + 1];
        const float re[4] = {a.x, a.z, c.x, c.z}, im[4] = {a.y, a.w, c.y, c.w};
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {                           // axis 0 = imaginary (low bits), axis 1 = real
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < hb) {
                    float out[4];
#pragma unroll
                    for (int tx = 0; tx < 4; ++tx) {
                        const float v = ax ? re[tx] : im[tx];
                        float d0 = 1e30f, d1 = 1e30f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j < sl.L) {
                                const float e = v - lv[j], e2 = e * e;
                                if ((j >> k) & 1) d1 = fminf(d1, e2); else d0 = fminf(d0, e2);
                            }
                        }
                        float l = (d1 - d0) * inv;
                        const int bit = ax * hb + k;
                        if (cal_a) l = fminf(fmaxf(-(ca[bit] * l + cb[bit]), -clip), clip);
                        out[tx] = l;
                    }
                    Lf[(size_t)n * qam_bits + ax * hb + k] = make_float4(out[0], out[1], out[2], out[3]);
                }
            }
        }
    }
}

// LLR calibration: per bit position, maxiter steps of full-batch gradient descent on the 1-D logistic
// regression p(y = 1 | x) = sigmoid(a x + b), x = llr[.., bit, ..], y = transmitted bit.  One CTA per
// bit position, fp64 throughout.
template <typename T>
__global__ void __launch_bounds__(1024)
llr_calibrate_kernel(const T *__restrict__ llr, const uint8_t *__restrict__ tx_idx, long long n_sym_total, int N_t,
                     int qam_bits, int maxiter, double lr, double l2, double *__restrict__ ab) {
    const int bit = blockIdx.x;
    __shared__ double s_ga[32], s_gb[32], s_a, s_b;
    if (threadIdx.x == 0) { s_a = 1.0; s_b = 0.0; }
    __syncthreads();
    for (int it = 0; it < maxiter; ++it) {
        const double a = s_a, bb = s_b;
        double ga = 0.0, gb = 0.0;
        for (long long i = threadIdx.x; i < n_sym_total; i += blockDim.x) {
            const long long n = i / N_t;                   // (frame, subcarrier) index
            const int tx = (int)(i - n * N_t);
            const double x = (double)llr[(n * qam_bits + bit) * N_t + tx];
            const double y = (double)((tx_idx[i] >> bit) & 1);
            const double p = 1.0 / (1.0 + exp(-(a * x + bb)));
            ga += (p - y) * x;
            gb += (p - y);
        }
        for (int s = 16; s > 0; s >>= 1) {
            ga += __shfl_xor_sync(0xffffffffu, ga, s);
            gb += __shfl_xor_sync(0xffffffffu, gb, s);
        }
        if ((threadIdx.x & 31) == 0) { s_ga[threadIdx.x >> 5] = ga; s_gb[threadIdx.x >> 5] = gb; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ta = 0.0, tb = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ta += s_ga[w]; tb += s_gb[w]; }
            s_a = a - lr * (ta / (double)n_sym_total + l2 * a);
            s_b = bb - lr * (tb / (double)n_sym_total);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { ab[2 * bit] = s_a; ab[2 * bit + 1] = s_b; }
}

extern "C" int ofdm_soft_demap(int dtype, const void *X_hat, int B, int N, int N_t, int qam_bits, const double *cal_a,
                               const double *cal_b, double clip, void *sigma2, void *llr, void *stream) {
    if (!X_hat || B <= 0 || N <= 0 || N_t <= 0 || (!sigma2 && !llr) || ((cal_a == nullptr) != (cal_b == nullptr)))
        return ESN_E_BADARG;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return ESN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ESN_F32)
        if (N_t == 4)
            soft_demap_nt4_f32_kernel<<<B, 256, 0, st>>>((const float *)X_hat, N, qam_bits, cal_a, cal_b, (float)clip, (float *)sigma2, (float *)llr);
        else
            soft_demap_kernel<float><<<B, 256, 0, st>>>((const float *)X_hat, N, N_t, qam_bits, cal_a, cal_b, (float)clip, (float *)sigma2, (float *)llr);
    else if (dtype == ESN_F64)
        soft_demap_kernel<double><<<B, 256, 0, st>>>((const double *)X_hat, N, N_t, qam_bits, cal_a, cal_b, clip, (double *)sigma2, (double *)llr);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_llr_calibrate(int dtype, const void *llr, const uint8_t *tx_idx, int B, int N, int N_t, int qam_bits,
                                  int maxiter, double lr, double l2, double *ab, void *stream) {
    if (!llr || !tx_idx || !ab || B <= 0 || N <= 0 || N_t <= 0 || maxiter < 0) return ESN_E_BADARG;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return ESN_E_BADARG;
    const long long n = (long long)B * N * N_t;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ESN_F32)
        llr_calibrate_kernel<float><<<qam_bits, 1024, 0, st>>>((const float *)llr, tx_idx, n, N_t, qam_bits, maxiter, lr, l2, ab);
    else if (dtype == ESN_F64)
        llr_calibrate_kernel<double><<<qam_bits, 1024, 0, st>>>((const double *)llr, tx_idx, n, N_t, qam_bits, maxiter, lr, l2, ab);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_demap_count(int dtype, const void *X_hat, int B, int N, int N_t, int qam_bits, uint8_t *idx,
                                const uint8_t *tx_idx, double boundary_eps, unsigned long long *counts,
                                void *stream) {
    if (!X_hat || B <= 0 || N <= 0 || N_t <= 0) return ESN_E_BADARG;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return ESN_E_BADARG;
    const size_t total = (size_t)B * N * N_t;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)(148 * 8));
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = ((uintptr_t)X_hat % 16 == 0) && ((uintptr_t)tx_idx % 4 == 0) && ((uintptr_t)idx % 4 == 0);
    if (aligned && total >= 4 && (dtype == ESN_F32 || dtype == ESN_F64)) {
        const int vb = (int)std::min<size_t>((total / 4 + 255) / 256, (size_t)(148 * 8));
        if (dtype == ESN_F32)
            demap_count_vec_kernel<float><<<vb, 256, 0, st>>>((const float *)X_hat, total, qam_bits, idx, tx_idx, (float)boundary_eps, counts);
        else
            demap_count_vec_kernel<double><<<vb, 256, 0, st>>>((const double *)X_hat, total, qam_bits, idx, tx_idx, boundary_eps, counts);
        return esn_launch_status();
    }
    if (dtype == ESN_F32)
        demap_count_kernel<float><<<blocks, 256, 0, st>>>((const float *)X_hat, total, qam_bits, idx, tx_idx, (float)boundary_eps, counts);
    else if (dtype == ESN_F64)
        demap_count_kernel<double><<<blocks, 256, 0, st>>>((const double *)X_hat, total, qam_bits, idx, tx_idx, boundary_eps, counts);
    else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int ofdm_synth_frames(int dtype, const uint8_t *tx_idx, const void *taps, const int32_t *chan_index,
                                 const void *Pi, const void *A_clip, const void *noise, double noise_std,
                                 unsigned long long seed, int B, int N, int cp, int N_t, int N_r, int ntaps,
                                 int qam_bits, int delay, void *x_cp, void *y_cp, void *esn_in, void *stream) {
    if (!tx_idx || !taps || !Pi || !A_clip || (!y_cp && !esn_in)) return ESN_E_BADARG;
    if (B <= 0 || N_t <= 0 || N_r <= 0 || ntaps <= 0 || cp < 0 || cp >= N || delay < 0 || !pow2_ok(N)) return ESN_E_BADARG;
    if (qam_bits != 2 && qam_bits != 4 && qam_bits != 6) return ESN_E_BADARG;
    const size_t el = 3 * (size_t)N + 2 * (size_t)N_t * (N + cp);
    cudaStream_t st = (cudaStream_t)stream;
    // whole-frame kernel when the N_t streams, the clipped samples and the taps fit in shared memory together
    const size_t tot = (size_t)N * N_t;
    const size_t el_frame = 2 * (tot + tot / 32 + 1) + N + 2 * (size_t)N_t * (N + cp) + 2 * (size_t)N_r * N_t * (ntaps | 1);
    const size_t esz = dtype == ESN_F64 ? sizeof(double) : sizeof(float);
    if (el_frame * esz <= 160 * 1024 && (dtype == ESN_F32 || dtype == ESN_F64)) {
        if (dtype == ESN_F32) {
            if (int rc = allow_smem(synth_frames_frame_kernel<float>, el_frame * esz)) return rc;
            synth_frames_frame_kernel<float><<<B, 256, el_frame * esz, st>>>(
                tx_idx, (const float *)taps, chan_index, (const float *)Pi, (const float *)A_clip, (const float *)noise,
                (float)noise_std, seed, N, cp, N_t, N_r, ntaps, qam_bits, delay, (float *)x_cp, (float *)y_cp, (float *)esn_in);
        } else {
            if (int rc = allow_smem(synth_frames_frame_kernel<double>, el_frame * esz)) return rc;
            synth_frames_frame_kernel<double><<<B, 256, el_frame * esz, st>>>(
                tx_idx, (const double *)taps, chan_index, (const double *)Pi, (const double *)A_clip, (const double *)noise,
                noise_std, seed, N, cp, N_t, N_r, ntaps, qam_bits, delay, (double *)x_cp, (double *)y_cp, (double *)esn_in);
        }
        return esn_launch_status();
    }
    if (dtype == ESN_F32) {
        if (int rc = allow_smem(synth_frames_kernel<float>, el * sizeof(float))) return rc;
        synth_frames_kernel<float><<<B, fft_threads(N), el * sizeof(float), st>>>(
            tx_idx, (const float *)taps, chan_index, (const float *)Pi, (const float *)A_clip, (const float *)noise,
            (float)noise_std, seed, N, cp, N_t, N_r, ntaps, qam_bits, delay, (float *)x_cp, (float *)y_cp, (float *)esn_in);
    } else if (dtype == ESN_F64) {
        if (int rc = allow_smem(synth_frames_kernel<double>, el * sizeof(double))) return rc;
        synth_frames_kernel<double><<<B, fft_threads(N), el * sizeof(double), st>>>(
            tx_idx, (const double *)taps, chan_index, (const double *)Pi, (const double *)A_clip, (const double *)noise,
            noise_std, seed, N, cp, N_t, N_r, ntaps, qam_bits, delay, (double *)x_cp, (double *)y_cp, (double *)esn_in);
    } else return ESN_E_BADARG;
    return esn_launch_status();
}

extern "C" int esn_version(void) { return 1; }

extern "C" int esn_device_info(char *name, int n, int *major, int *minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ESN_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return ESN_E_NODEVICE;
    if (name && n > 0) {
        int i = 0;
        for (; i < n - 1 && prop.name[i]; ++i) name[i] = prop.name[i];
        name[i] = 0;
    }
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    return prop.multiProcessorCount;
}

// Host-callable copy of the device noise stream (tests pin the numpy restatement to it).
extern "C" float esn_noise_uniform_host(unsigned long long seed, unsigned frame, unsigned row, unsigned neuron) {
    return esn_noise_uniform(esn_noise_key(seed, frame, row), neuron);
}
