// Reservoir recurrence, SIMT path (fp32 FFMA / fp64 DFMA), any reservoir size.
//
// Replaces ESN._update and the Python time loops of ESN.fit / ESN.predict
// (reference libs/pyESN.py:111-125, :179-182, :243-253).  One CTA owns a tile of
// BT independent frames and steps them through all T time steps without leaving
// the SM: the augmented state [x; u; feedback] of the tile lives in shared
// memory, the augmented weight matrix Wt_aug = [W^T; W_in^T; W_fb^T] is streamed
// from L2 in KC-row chunks with cp.async (it is shared by every CTA, so it stays
// L2-resident), each thread keeps an 8-frame x 4-neuron register tile.  A step
// needs the whole old state until its last slab is done, so the new state is
// parked in a small global workspace (L2-resident, B*N elements) and pulled back
// into shared memory between steps; that keeps ONE state buffer on chip and
// doubles the frame tile that fits.  The readout y_n = W_out [x_n; u_n] that the
// free-running mode feeds back is fused between steps.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace esn_simt {

struct RecParams {
    int mode, B, T, N, n_in, n_out, N_pad, K_aug_pad, transient, feedback, n_groups;
    int stage_wout;
    double noise_amp;
    unsigned long long seed;
    const void *Wt_aug, *in, *in_scale, *in_shift, *teacher, *t_scale, *t_shift, *W_out;
    const int *group_ids;
    const void *x0, *y0, *noise;
    void *ext_out, *y_out, *workspace;
};

constexpr int NPT = 4;     // neurons per thread (lane + 32*i)
constexpr int SLABW = 128; // neurons per warp column

// FR = frames per thread (register tile FR x 4); a smaller FR gives more warps per frame tile,
// which is what hides the DFMA / LDS latency of the fp64 path at small tiles.
template <typename T, int BT, int WN, int KC, int FR_ = 8, int ST_ = 2>
struct RecCfg {
    static constexpr int ST = ST_;                        // cp.async stages of the weight stream
    static constexpr int FR = FR_;
    static constexpr int THREADS = (BT / FR) * WN * 32;
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int NS = SLABW * WN;                 // neurons per slab pass
    static constexpr int PADF = 16 / (int)sizeof(T);      // row padding (16 bytes)
    static constexpr int RS = BT + PADF;                  // row stride of xs in elements
    static constexpr int FL = BT < 32 ? BT : 32;          // frames across lanes in the readout
    static constexpr int KS = 32 / FL;                    // k sub-slices per warp in the readout
    static constexpr int NPART = NWARPS * KS;
};

template <typename T, int BT, int WN, int KC, int FR_, int ST_>
__global__ void __launch_bounds__(RecCfg<T, BT, WN, KC, FR_, ST_>::THREADS, 1)
esn_recurrence_simt(const RecParams p) {
    using C = RecCfg<T, BT, WN, KC, FR_, ST_>;
    constexpr int RS = C::RS, NS = C::NS, THREADS = C::THREADS, FR = C::FR, ST = C::ST;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int N = p.N, n_in = p.n_in, n_out = p.n_out;
    const int P = N + n_in;
    const int Kp = p.K_aug_pad;
    const int nchunks = Kp / KC;
    const int nslabs = (p.N_pad + NS - 1) / NS;
    const int tiles_per_step = nslabs * nchunks;

    T *xs = reinterpret_cast<T *>(smem_raw);               // [Kp][RS]  rows: x | u | feedback | 0
    T *wbuf = xs + (size_t)Kp * RS;                        // [ST][KC][NS]
    T *red = wbuf + ST * KC * NS;                          // [NPART][BT][n_out]
    T *wout_s = red + C::NPART * BT * n_out;               // [n_out][P] (group-uniform tiles)
    __shared__ int s_group[BT];
    __shared__ int s_uniform;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fg = warp / WN, wn = warp % WN;
    const int tile0 = blockIdx.x * BT;                     // first frame of this tile

    const T *Wt = static_cast<const T *>(p.Wt_aug);
    const T *gin = static_cast<const T *>(p.in);
    const T *in_scale = static_cast<const T *>(p.in_scale);
    const T *in_shift = static_cast<const T *>(p.in_shift);
    const T *t_scale = static_cast<const T *>(p.t_scale);
    const T *t_shift = static_cast<const T *>(p.t_shift);
    const T *teacher = static_cast<const T *>(p.teacher);
    const T *gW_out = static_cast<const T *>(p.W_out);
    const T *noise = static_cast<const T *>(p.noise);
    T *ext = static_cast<T *>(p.ext_out);
    T *yout = static_cast<T *>(p.y_out);
    T *ws = static_cast<T *>(p.workspace);                 // [B][N] parked new state
    const bool predict = p.mode == ESN_MODE_PREDICT;
    const int s0 = predict ? 0 : 1;
    const int noise_rows = predict ? p.T : p.T - 1;
    const T namp = (T)p.noise_amp;
    const bool use_noise = p.noise_amp != 0.0;

    // ---- one-time setup -------------------------------------------------
    for (int i = tid; i < Kp * RS; i += THREADS) xs[i] = (T)0;
    if (tid < BT) {
        int b = tile0 + tid;
        s_group[tid] = (predict && p.group_ids && b < p.B) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;   // clamped: a bad id must not read outside W_out
    }
    __syncthreads();
    if (tid == 0) {
        int u = p.stage_wout;
        for (int f = 1; f < BT; ++f)
            if (tile0 + f < p.B && s_group[f] != s_group[0]) u = 0;
        s_uniform = u;
    }
    __syncthreads();
    const bool uniform = s_uniform != 0;
    if (predict && uniform) {
        const T *src = gW_out + (size_t)s_group[0] * n_out * P;
        for (int i = tid; i < n_out * P; i += THREADS) wout_s[i] = src[i];
    }

    // stage the scaled inputs of time step `row` into the u rows (and into E)
    auto stage_inputs = [&](int row, bool to_smem) {
        for (int i = tid; i < BT * n_in; i += THREADS) {
            int f = i / n_in, j = i - f * n_in, b = tile0 + f;
            T v = (T)0;
            if (b < p.B && row < p.T) {
                v = gin[((size_t)b * p.T + row) * n_in + j] * in_scale[j] + in_shift[j];
                if (ext) ext[((size_t)b * p.T + row) * P + N + j] = v;
            }
            if (to_smem) xs[(N + j) * RS + f] = v;
        }
    };
    auto stage_teacher = [&](int row) {   // HARVEST feedback d_row
        for (int i = tid; i < BT * n_out; i += THREADS) {
            int f = i / n_out, o = i - f * n_out, b = tile0 + f;
            T v = (T)0;
            if (p.feedback && b < p.B && row < p.T)
                v = teacher[((size_t)b * p.T + row) * n_out + o] * t_scale[o] + t_shift[o];
            xs[(P + o) * RS + f] = v;
        }
    };

    if (predict) {
        if (p.x0) {
            const T *x0 = static_cast<const T *>(p.x0);
            for (int i = tid; i < BT * N; i += THREADS) {
                int f = i / N, k = i - f * N, b = tile0 + f;
                if (b < p.B) xs[k * RS + f] = x0[(size_t)b * N + k];
            }
        }
        if (p.y0 && p.feedback) {
            const T *y0 = static_cast<const T *>(p.y0);
            for (int i = tid; i < BT * n_out; i += THREADS) {
                int f = i / n_out, o = i - f * n_out, b = tile0 + f;
                if (b < p.B) xs[(P + o) * RS + f] = y0[(size_t)b * n_out + o];
            }
        }
        stage_inputs(0, true);
    } else {
        // E row 0 = [0, u_0]; x_0 = 0
        for (int i = tid; i < BT * N; i += THREADS) {
            int f = i / N, k = i - f * N, b = tile0 + f;
            if (b < p.B) ext[((size_t)b * p.T) * P + k] = (T)0;
        }
        stage_inputs(0, false);
        stage_inputs(1, true);
        stage_teacher(0);
    }

    // ---- W chunk pipeline: tile t -> (slab, chunk), endless across steps ----
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int PIECES_PER_ROW = NS / VEC;
    const int total_tiles = (p.T - s0) * tiles_per_step;
    // one commit group per call, empty past the end, so that wait_group<ST-2> always means "tile t landed"
    auto prefetch = [&](int t) {
        if (t >= total_tiles) { cp_async_commit(); return; }
        int tt = t % tiles_per_step;
        int slab = tt / nchunks, c = tt - slab * nchunks;
        T *dst = wbuf + (t % ST) * KC * NS;
        const T *src = Wt + (size_t)(c * KC) * p.N_pad + slab * NS;
        int cols_left = p.N_pad - slab * NS;                  // multiple of 128, may be < NS
        for (int i = tid; i < KC * PIECES_PER_ROW; i += THREADS) {
            int r = i / PIECES_PER_ROW, q = i - r * PIECES_PER_ROW;
            if (q * VEC < cols_left) cp_async16(dst + r * NS + q * VEC, src + (size_t)r * p.N_pad + q * VEC);
        }
        cp_async_commit();
    };

    for (int i = 0; i < ST - 1; ++i) prefetch(i);
    int t = 0;

    for (int n = s0; n < p.T; ++n) {
        const int nrow = predict ? n : n - 1;                 // noise row of this step
        for (int slab = 0; slab < nslabs; ++slab) {
            T acc[FR][NPT];
#pragma unroll
            for (int f = 0; f < FR; ++f)
#pragma unroll
                for (int i = 0; i < NPT; ++i) acc[f][i] = (T)0;
            const bool warp_active = slab * NS + wn * SLABW < p.N_pad;

            for (int c = 0; c < nchunks; ++c, ++t) {
                cp_async_wait<ST - 2>();
                __syncthreads();        // chunk t landed for everyone; everyone is done with chunk t-1
                prefetch(t + ST - 1);   // into the buffer chunk t-1 just left
                if (warp_active) {
                    const T *wb = wbuf + (t % ST) * KC * NS + wn * SLABW + lane;
                    const T *xb = xs + (size_t)(c * KC) * RS + fg * FR;
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) {
                        T xv[FR], wv[NPT];
#pragma unroll
                        for (int f = 0; f < FR; ++f) xv[f] = xb[kk * RS + f];
#pragma unroll
                        for (int i = 0; i < NPT; ++i) wv[i] = wb[kk * NS + 32 * i];
#pragma unroll
                        for (int f = 0; f < FR; ++f)
#pragma unroll
                            for (int i = 0; i < NPT; ++i) acc[f][i] = fma(xv[f], wv[i], acc[f][i]);
                    }
                }
            }
            // ---- slab epilogue: tanh + noise; park x_n in the workspace, E row n ----
            if (warp_active) {
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    const int nn = slab * NS + wn * SLABW + lane + 32 * i;
                    if (nn < N) {
#pragma unroll
                        for (int f = 0; f < FR; ++f) {
                            const int b = tile0 + fg * FR + f;
                            if (b < p.B) {
                                T x = esn_tanh<T>(acc[f][i]);
                                if (use_noise) {
                                    T u;
                                    if (noise) u = noise[((size_t)b * noise_rows + nrow) * N + nn];
                                    else u = (T)esn_noise_uniform(esn_noise_key(p.seed, (uint32_t)b, (uint32_t)nrow), (uint32_t)nn);
                                    x += namp * (u - (T)0.5);
                                }
                                if (ext) ext[((size_t)b * p.T + n) * P + nn] = x;
                                ws[(size_t)b * N + nn] = x;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                  // every slab is done with x_{n-1}; x_n is parked (block-visible)

        // ---- pull x_n back into shared memory ----
        for (int i = tid; i < BT * N; i += THREADS) {
            int f = i / N, k = i - f * N, b = tile0 + f;
            if (b < p.B) xs[k * RS + f] = __ldcg(&ws[(size_t)b * N + k]);
        }
        __syncthreads();

        if (!predict) {
            stage_teacher(n);
            stage_inputs(n + 1, true);
        } else {
            // y_n[f][o] = sum_k W_out[g][o][k] * [x_n ; u_n][k]
            constexpr int FL = C::FL, KS = C::KS;
            const int fl = lane % FL, ksub = lane / FL;
            const int part = warp * KS + ksub;
            const int kper = (P + C::NPART - 1) / C::NPART;
            const int kbeg = part * kper, kend = min(P, kbeg + kper);
#pragma unroll 1
            for (int f0 = 0; f0 < BT; f0 += FL) {
                const int f = f0 + fl;
                T a[ESN_MAX_OUT];
#pragma unroll
                for (int o = 0; o < ESN_MAX_OUT; ++o) a[o] = (T)0;
                const T *wsrc = uniform ? wout_s : gW_out + (size_t)s_group[f] * n_out * P;
                for (int k = kbeg; k < kend; ++k) {
                    const T xv = xs[k * RS + f];
#pragma unroll
                    for (int o = 0; o < ESN_MAX_OUT; ++o)
                        if (o < n_out) a[o] = fma(wsrc[(size_t)o * P + k], xv, a[o]);
                }
#pragma unroll
                for (int o = 0; o < ESN_MAX_OUT; ++o)
                    if (o < n_out) red[(part * BT + f) * n_out + o] = a[o];
            }
            __syncthreads();              // partial sums ready; nobody reads u_n any more
            for (int i = tid; i < BT * n_out; i += THREADS) {
                const int f = i / n_out, o = i - f * n_out, b = tile0 + f;
                T y = (T)0;
                for (int q = 0; q < C::NPART; ++q) y += red[(q * BT + f) * n_out + o];
                xs[(P + o) * RS + f] = p.feedback ? y : (T)0;
                if (b < p.B && n >= p.transient)
                    yout[((size_t)b * (p.T - p.transient) + (n - p.transient)) * n_out + o] =
                        (y - t_shift[o]) / t_scale[o];
            }
            stage_inputs(n + 1, true);
        }
        // the __syncthreads at the top of the next chunk publishes the staged rows
    }
}

template <typename T, int BT, int WN, int KC, int FR_, int ST_>
size_t rec_smem_bytes(int Kp, int n_out, int P, bool stage_wout) {
    using C = RecCfg<T, BT, WN, KC, FR_, ST_>;
    size_t el = (size_t)Kp * C::RS + (size_t)ST_ * KC * C::NS + (size_t)C::NPART * BT * n_out +
                (stage_wout ? (size_t)n_out * P : 0);
    return el * sizeof(T);
}

constexpr size_t kSmemLimit = 227 * 1024 - 2048;   // static __shared__ + slack

template <typename T, int BT, int WN, int KC, int FR_ = 8, int ST_ = 2>
int launch_cfg(RecParams p, cudaStream_t st, bool dry) {
    using C = RecCfg<T, BT, WN, KC, FR_, ST_>;
    const int P = p.N + p.n_in;
    bool stage = p.mode == ESN_MODE_PREDICT;
    size_t smem = rec_smem_bytes<T, BT, WN, KC, FR_, ST_>(p.K_aug_pad, p.n_out, P, stage);
    if (smem > kSmemLimit && stage) {
        stage = false;
        smem = rec_smem_bytes<T, BT, WN, KC, FR_, ST_>(p.K_aug_pad, p.n_out, P, false);
    }
    if (smem > kSmemLimit) return ESN_E_TOOLARGE;
    if (dry) return 0;
    p.stage_wout = stage ? 1 : 0;
    auto kern = esn_recurrence_simt<T, BT, WN, KC, FR_, ST_>;
    ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (p.B + BT - 1) / BT;
    kern<<<grid, C::THREADS, smem, st>>>(p);
    return esn_launch_status();
}

// Pick the largest frame tile that fits shared memory and still gives every SM
// at least one CTA; small batches get the smallest tile.
template <typename T>
int launch_any(const RecParams &p, cudaStream_t st) {
    const int slabs128 = p.N_pad / 128;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // tuning aid: ESN_SIMT_CFG="BT,WN,KC" restricts the choice to that tile (if it is instantiated)
    int fb = 0, fw = 0, fk = 0;
    if (const char *e = getenv("ESN_SIMT_CFG")) sscanf(e, "%d,%d,%d", &fb, &fw, &fk);
#define FORCED_OUT(BT_, WN_, KC_) (fb != 0 && !(fb == BT_ && fw == WN_ && fk == KC_))
#define TRY(BT_, WN_, KC_)                                                              \
    if (!FORCED_OUT(BT_, WN_, KC_) && ((p.B + BT_ - 1) / BT_ >= sms || fb != 0) &&          \
        launch_cfg<T, BT_, WN_, KC_>(p, st, true) == 0)                                    \
        return launch_cfg<T, BT_, WN_, KC_>(p, st, false);
#define LAST(BT_, WN_, KC_)                                                             \
    if (!FORCED_OUT(BT_, WN_, KC_) && launch_cfg<T, BT_, WN_, KC_>(p, st, true) == 0)       \
        return launch_cfg<T, BT_, WN_, KC_>(p, st, false);
    if (slabs128 == 1) {
        TRY(64, 1, 16) TRY(32, 1, 16) TRY(16, 1, 16) LAST(8, 1, 16)
        return ESN_E_TOOLARGE;
    }
    if constexpr (sizeof(T) == 8) {
        // fp64: measured on B200 (profiles/fit_bench.py, 1184 frames, N=512): (8,4,8) 46 ms,
        // (8,4,16) 113 ms, (16,4,8) 62 ms, (32,2,8) 96 ms -> small tiles with KC=8 first
        TRY(32, 2, 8) TRY(16, 4, 8) LAST(8, 4, 8) LAST(8, 2, 8) LAST(8, 4, 4)
    } else {
        TRY(64, 2, 8) TRY(32, 2, 16) TRY(32, 2, 8) TRY(16, 4, 16) TRY(16, 4, 8)
        LAST(8, 4, 16) LAST(8, 4, 8) LAST(8, 2, 8)
    }
#undef TRY
#undef LAST
    return ESN_E_TOOLARGE;
}

}  // namespace esn_simt
