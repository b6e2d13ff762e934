// Reservoir recurrence in fp64 on the fp64 tensor cores (mma.sync.m8n8k4.f64, DMMA), sm_100a.
//
// The reference's own arithmetic class for whole batches: the teacher-forced harvest of ESN.fit (reference
// libs/pyESN.py:179-182: states[n] = tanh(W states[n-1] + W_in u[n] + W_fb d[n-1]) + noise (rand - 0.5)) -- the
// parity-grade half of readout training -- and the free-running loop of ESN.predict (:243-253), for reservoirs up to
// 1024 neurons.  The streaming SIMT kernel (recurrence_simt.cuh) issues one DFMA per MAC with two shared-memory
// operands per 4 x 8 register tile and sits at 0.39 of the fp64 rate; here a CTA steps a tile of 32 (16, 8) frames
// with the MACs on the DMMA pipe (4736 pilots of cfg3: 94 -> 47 ms, 0.78 of the fp64 rate; 9472 frames predicted:
// 215 -> 117 ms; profiles/r2_fp64_harvest.txt, r2_fp64_predict.txt).
//
// Layout.  One CTA = 8 FB frames (FB = 4, 2, 1), 16 warps.  The augmented state [x | u | d | 0] of the tile lives in
// shared memory as xs[frame][k] (row stride K_aug_pad + 4 doubles: the 8 x 4 A fragments of a half-warp fall into
// distinct banks).  Warp w owns neurons 32 w .. 32 w + 31 for all frames: FB x 4 accumulator fragments.  The
// augmented weights Wt_aug [K_aug_pad][N_pad] (the SIMT kernel's image; shared by every CTA, L2-resident) reach the
// warp through its own cp.async ring in shared memory (see below).  A step = 136 k-steps of (FB A loads, 4 B loads,
// 4 FB DMMAs) per warp, one barrier, the epilogue (tanh, noise, E row n, new state into xs in place -- every warp
// has finished reading x_{n-1}), the teacher / input rows of the next step, one barrier.  Same noise stream, same
// E layout as the SIMT kernel.  (First version: B fragments straight from global memory, 32-byte rows -- the
// 32-frame tile ran at the same speed, the 8- and 16-frame tiles at 42 / 49 us per step against 24.5 now.)
#include "recurrence_simt.cuh"

namespace {

constexpr int DH_THREADS = 512;

__device__ __forceinline__ void dh_dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// FB = 8-frame blocks per CTA (tiles of 8, 16 or 32 frames), NSTG = weight-ring stages, PREDICT = free-running mode
// (ESN.predict, libs/pyESN.py:243-253: the readout y_n = W_out[g(b)] [x_n; u_n] is fed back and emitted; tiles of 8 or
// 16 frames: one warp (or two) per frame in the readout).
// CF = 8-neuron fragments per warp: 16 warps x 8 CF neurons cover the reservoir in one pass (4: up to 512 neurons,
// 5 / 6 / 8: the 600-neuron demo, 768, 1024), so the state never has to be parked.
template <int FB, int NSTG, bool PREDICT, int CF>
__global__ void __launch_bounds__(DH_THREADS, 1)
esn_harvest_dmma_kernel(const esn_simt::RecParams p) {
    constexpr int DH_BT = 8 * FB;
    extern __shared__ __align__(16) double dh_xs[];
    const int N = p.N, n_in = p.n_in, n_out = p.n_out, P = N + n_in, Kp = p.K_aug_pad, NP = p.N_pad;
    const int RS = Kp + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, fk = lane & 3, fi = lane >> 2;
    const int tile0 = blockIdx.x * DH_BT;
    const int n0 = warp * 8 * CF;
    const bool wact = n0 < N;
    const double *Wt = static_cast<const double *>(p.Wt_aug);
    const double *gin = static_cast<const double *>(p.in);
    const double *in_scale = static_cast<const double *>(p.in_scale), *in_shift = static_cast<const double *>(p.in_shift);
    const double *t_scale = static_cast<const double *>(p.t_scale), *t_shift = static_cast<const double *>(p.t_shift);
    const double *teacher = static_cast<const double *>(p.teacher);
    const double *noise = static_cast<const double *>(p.noise);
    double *ext = static_cast<double *>(p.ext_out);
    double *xs = dh_xs;
    const bool use_noise = p.noise_amp != 0.0;
    const double namp = p.noise_amp;
    const int noise_rows = PREDICT ? p.T : p.T - 1;
    const int s0 = PREDICT ? 0 : 1;
    const double *gW_out = static_cast<const double *>(p.W_out);
    double *yout = static_cast<double *>(p.y_out);
    __shared__ int s_group[DH_BT];

    for (int i = tid; i < DH_BT * RS; i += DH_THREADS) xs[i] = 0.0;
    if (!PREDICT) {
        // E row 0 = [0, u_0]
        for (int i = tid; i < DH_BT * N; i += DH_THREADS) {
            const int f = i / N, k = i - f * N, b = tile0 + f;
            if (b < p.B) ext[((size_t)b * p.T) * P + k] = 0.0;
        }
    } else if (tid < DH_BT) {
        const int b = tile0 + tid;                          // clamped: a bad id must not read outside W_out
        s_group[tid] = (p.group_ids && b < p.B) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;
    }
    __syncthreads();
    // scaled inputs of time step `row` -> E (and the u columns of xs); scaled teacher of `row` -> the d columns
#define DH_STAGE_INPUTS(row, to_smem)                                                                             \
    for (int i = tid; i < DH_BT * n_in; i += DH_THREADS) {                                                        \
        const int f = i / n_in, j = i - f * n_in, b = tile0 + f;                                                  \
        double v = 0.0;                                                                                           \
        if (b < p.B && (row) < p.T) {                                                                             \
            v = gin[((size_t)b * p.T + (row)) * n_in + j] * in_scale[j] + in_shift[j];                            \
            if (ext) ext[((size_t)b * p.T + (row)) * P + N + j] = v;                                              \
        }                                                                                                         \
        if (to_smem) xs[f * RS + N + j] = v;                                                                      \
    }
#define DH_STAGE_TEACHER(row)                                                                                     \
    for (int i = tid; i < DH_BT * n_out; i += DH_THREADS) {                                                       \
        const int f = i / n_out, o = i - f * n_out, b = tile0 + f;                                                \
        double v = 0.0;                                                                                           \
        if (p.feedback && b < p.B && (row) < p.T)                                                                 \
            v = teacher[((size_t)b * p.T + (row)) * n_out + o] * t_scale[o] + t_shift[o];                         \
        xs[f * RS + P + o] = v;                                                                                   \
    }
    if (!PREDICT) {
        DH_STAGE_INPUTS(0, false)
        DH_STAGE_INPUTS(1, true)
        DH_STAGE_TEACHER(0)
    } else {
        if (p.x0) {                                         // continuation state / last output of the fit
            const double *x0 = static_cast<const double *>(p.x0);
            for (int i = tid; i < DH_BT * N; i += DH_THREADS) {
                const int f = i / N, k = i - f * N, b = tile0 + f;
                if (b < p.B) xs[f * RS + k] = x0[(size_t)b * N + k];
            }
        }
        if (p.y0 && p.feedback) {
            const double *y0 = static_cast<const double *>(p.y0);
            for (int i = tid; i < DH_BT * n_out; i += DH_THREADS) {
                const int f = i / n_out, o = i - f * n_out, b = tile0 + f;
                if (b < p.B) xs[f * RS + P + o] = y0[(size_t)b * n_out + o];
            }
        }
        DH_STAGE_INPUTS(0, true)
    }
    __syncthreads();

    // The weight stream: warp w needs Wt_aug[k][32 w .. 32 w + 31] for every k, 256 contiguous bytes per k row.  It
    // copies them itself, 8 k rows (2 KB) per cp.async group, into its own ring of NSTG stages (row stride 36 doubles:
    // conflict-free B fragments) -- full 128-byte lines instead of 32-byte fragment rows, no register staging, and
    // only warp-level synchronisation; the ring runs ahead across time steps (the weights do not change).
    constexpr int KR = 8, WRS = (8 * CF + 15) / 16 * 16 + 4;  // row stride = 4 mod 16 doubles: conflict-free B fragments
    constexpr int PPR = 4 * CF;                             // 16-byte pieces per k row of the warp's columns
    const int nch = Kp / KR;                                // K_aug_pad is a multiple of 16
    double *wst = xs + DH_BT * RS + (size_t)warp * NSTG * KR * WRS;
    int left = (p.T - s0) * nch, ic = 0, is = 0;             // chunks still to issue, next chunk of the image, next stage
    auto issue = [&]() {                                    // one commit group per call, empty past the end
        if (wact && left > 0) {
            const double *src = Wt + n0 + (size_t)(KR * ic) * NP;
            double *dst = wst + is * KR * WRS;
#pragma unroll
            for (int q = 0; q < CF; ++q) {                  // KR * PPR = 32 CF pieces per chunk
                const int piece = lane + 32 * q, row = piece / PPR, col = 2 * (piece % PPR);
                cp_async16(dst + row * WRS + col, src + (size_t)row * NP + col);
            }
            --left;
            if (++ic == nch) ic = 0;
            if (++is == NSTG) is = 0;
        }
        cp_async_commit();
    };
    for (int i = 0; i < NSTG - 1; ++i) issue();
    int ds = 0;                                             // stage of the chunk in use
    const double *xa = xs + fi * RS + fk;                   // A fragment r of k-step k4: xa[8 r RS + 4 k4]
    for (int n = s0; n < p.T; ++n) {
        const int nrow = PREDICT ? n : n - 1;               // noise row of this step
        double acc[FB][CF][2];
#pragma unroll
        for (int r = 0; r < FB; ++r)
#pragma unroll
            for (int c = 0; c < CF; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
        if (wact) {
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
                cp_async_wait<NSTG - 2>();
                __syncwarp();                               // this chunk has landed for every lane; the previous stage is free
                issue();
                const double *wb = wst + ds * KR * WRS + fk * WRS + fi;
                if (++ds == NSTG) ds = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double bu[CF], a[FB];
#pragma unroll
                    for (int cc = 0; cc < CF; ++cc) bu[cc] = wb[4 * j * WRS + 8 * cc];
#pragma unroll
                    for (int r = 0; r < FB; ++r) a[r] = xa[8 * r * RS + 4 * (2 * c + j)];
#pragma unroll
                    for (int r = 0; r < FB; ++r)
#pragma unroll
                        for (int cc = 0; cc < CF; ++cc) dh_dmma(acc[r][cc], a[r], bu[cc]);
                }
            }
        }
        __syncthreads();                                    // every warp is through x_{n-1}
        if (wact) {
#pragma unroll
            for (int r = 0; r < FB; ++r) {
                const int f = 8 * r + fi, b = tile0 + f;
                if (b < p.B) {
#pragma unroll
                    for (int c = 0; c < CF; ++c) {
                        const int nn = n0 + 8 * c + 2 * fk;
                        double x[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            x[h] = 0.0;
                            if (nn + h < N) {
                                x[h] = tanh(acc[r][c][h]);
                                if (use_noise) {
                                    double u;
                                    if (noise) u = noise[((size_t)b * noise_rows + nrow) * N + nn + h];
                                    else u = (double)esn_noise_uniform(esn_noise_key(p.seed, (uint32_t)b, (uint32_t)nrow), (uint32_t)(nn + h));
                                    x[h] += namp * (u - 0.5);
                                }
                                if (ext) ext[((size_t)b * p.T + n) * P + nn + h] = x[h];
                                xs[f * RS + nn + h] = x[h];
                            }
                        }
                    }
                }
            }
        }
        if constexpr (!PREDICT) {
            DH_STAGE_TEACHER(n)
            DH_STAGE_INPUTS(n + 1, true)
            __syncthreads();
        } else {
            __syncthreads();                                // x_n is complete in xs
            // the next input row's global loads are in flight during the readout
            constexpr int UPT = (DH_BT * ESN_MAX_IN + DH_THREADS - 1) / DH_THREADS;
            double un[UPT];
#pragma unroll
            for (int q = 0; q < UPT; ++q) {
                const int i = tid + q * DH_THREADS, f = i / n_in, j = i - f * n_in, b = tile0 + f;
                un[q] = 0.0;
                if (i < DH_BT * n_in && b < p.B && n + 1 < p.T)
                    un[q] = gin[((size_t)b * p.T + n + 1) * n_in + j] * in_scale[j] + in_shift[j];
            }
            // y_n[f][o] = sum_k W_out[g(f)][o][k] [x_n; u_n][k]: a warp owns a frame (two warps with half of the
            // outputs each for 8-frame tiles), its lanes stride k -- coalesced readout rows, conflict-free state
            // reads -- and a butterfly leaves every sum in every lane: no partial sums through shared memory
            {
                constexpr int WPF = (DH_THREADS / 32) / DH_BT;      // warps per frame: 1 or 2
                const int f = warp % DH_BT, oh = warp / DH_BT, b = tile0 + f;
                const double *wsrc = gW_out + (size_t)s_group[f] * n_out * P;
                const double *xr = xs + f * RS;
                double a[ESN_MAX_OUT / WPF];
#pragma unroll
                for (int q = 0; q < ESN_MAX_OUT / WPF; ++q) a[q] = 0.0;
#pragma unroll 6
                for (int k = lane; k < P; k += 32) {          // (unrolled: the loads of several rounds in flight)
                    const double xv = xr[k];
#pragma unroll
                    for (int q = 0; q < ESN_MAX_OUT / WPF; ++q) {
                        const int o = q * WPF + oh;
                        if (o < n_out) a[q] = fma(__ldg(wsrc + (size_t)o * P + k), xv, a[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < ESN_MAX_OUT / WPF; ++q) {
                    if (q * WPF + oh < n_out) {             // (warp-uniform)
#pragma unroll
                        for (int sh = 16; sh > 0; sh >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], sh);
                    }
                }
#pragma unroll
                for (int q = 0; q < ESN_MAX_OUT / WPF; ++q) {
                    const int o = q * WPF + oh;
                    if (o < n_out && lane == (q & 31)) {
                        xs[f * RS + P + o] = p.feedback ? a[q] : 0.0;       // the d columns are not read by any readout
                        if (b < p.B && n >= p.transient)
                            yout[((size_t)b * (p.T - p.transient) + (n - p.transient)) * n_out + o] = (a[q] - t_shift[o]) / t_scale[o];
                    }
                }
            }
            __syncthreads();                                // nobody reads u_n any more
#pragma unroll
            for (int q = 0; q < UPT; ++q) {
                const int i = tid + q * DH_THREADS, f = i / n_in, j = i - f * n_in, b = tile0 + f;
                if (i < DH_BT * n_in) {
                    xs[f * RS + N + j] = un[q];
                    if (ext && b < p.B && n + 1 < p.T) ext[((size_t)b * p.T + n + 1) * P + N + j] = un[q];
                }
            }
            __syncthreads();
        }
    }
#undef DH_STAGE_INPUTS
#undef DH_STAGE_TEACHER
}

}  // namespace

template <int FB, int NSTG, bool PREDICT, int CF>
static int dh_launch(const esn_simt::RecParams &p, cudaStream_t st) {
    const int bt = 8 * FB, ctas = (p.B + bt - 1) / bt;
    constexpr int WRS = (8 * CF + 15) / 16 * 16 + 4;
    const size_t smem = ((size_t)bt * (p.K_aug_pad + 4) + (size_t)(DH_THREADS / 32) * NSTG * 8 * WRS) * sizeof(double);
    if (smem > 226 * 1024) return ESN_E_UNSUPPORTED;
    auto kernel = esn_harvest_dmma_kernel<FB, NSTG, PREDICT, CF>;
    ESN_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<ctas, DH_THREADS, smem, st>>>(p);
    return esn_launch_status();
}

// wider reservoirs (640 / 768 / 1024 padded neurons): tiles of 16 or 8 frames, two ring stages
template <bool PREDICT, int CF>
static int dh_launch_wide(const esn_simt::RecParams &p, int bt, cudaStream_t st) {
    if constexpr (CF <= 6) {
        if (bt >= 16) {
            const int rc = dh_launch<2, 2, PREDICT, CF>(p, st);
            if (rc != ESN_E_UNSUPPORTED) return rc;
        }
    }
    return dh_launch<1, 2, PREDICT, CF>(p, st);
}

// Takes the fp64 harvests and predictions of reservoirs up to 1024 neurons that the cluster kernel (a few frames) has
// left: the largest tile of 32 (harvest only) / 16 / 8 frames that still gives every SM a CTA.  ESN_HARVEST_DMMA=0
// switches it off (the streaming SIMT kernel as the cross-check), =8 / 16 / 32 pins the tile.
static int dh_mode() {
    static const int mode = [] { const char *e = getenv("ESN_HARVEST_DMMA"); return e ? atoi(e) : -1; }();
    return mode;
}
bool esn_dmma_enabled() { return dh_mode() != 0; }

int esn_dmma_launch(const esn_simt::RecParams &p, cudaStream_t st) {
    const int mode = dh_mode();
    if (mode == 0 || p.N_pad > 1024) return ESN_E_UNSUPPORTED;
    const bool predict = p.mode == ESN_MODE_PREDICT;
    if (!predict && (p.mode != ESN_MODE_HARVEST || !p.ext_out)) return ESN_E_UNSUPPORTED;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    int bt = mode > 0 ? mode : ((p.B + 31) / 32 >= sms ? 32 : (p.B + 15) / 16 >= sms ? 16 : 8);
    if (p.N_pad > 512) {
        const int cf = p.N_pad / 128;                       // 5 .. 8
        if (cf == 5) return predict ? dh_launch_wide<true, 5>(p, bt, st) : dh_launch_wide<false, 5>(p, bt, st);
        if (cf == 6) return predict ? dh_launch_wide<true, 6>(p, bt, st) : dh_launch_wide<false, 6>(p, bt, st);
        return predict ? dh_launch_wide<true, 8>(p, bt, st) : dh_launch_wide<false, 8>(p, bt, st);   // 896 and 1024
    }
    if (predict) {
        if (bt > 16) bt = 16;
        return bt == 16 ? dh_launch<2, 4, true, 4>(p, st) : dh_launch<1, 4, true, 4>(p, st);
    }
    if (bt == 32) return dh_launch<4, 2, false, 4>(p, st);
    if (bt == 16) return dh_launch<2, 4, false, 4>(p, st);
    return dh_launch<1, 4, false, 4>(p, st);
}
