// Reservoir recurrence for SMALL batches: one thread-block cluster per group of frames, weights resident.
//
// Replaces ESN._update and the Python time loops of ESN.fit / ESN.predict (reference
// libs/pyESN.py:111-125, :179-182, :243-253) when only a few frames are in flight -- the call shape of
// the unmodified demo scripts (one OFDM symbol per predict(), one pilot per fit()).  There the step is a
// latency-bound mat-vec, so instead of streaming the 2 MB (fp64) weight matrix from L2 every step, the
// cluster keeps it in shared memory for the whole kernel: CTA r of the cluster owns the 32 neurons
// [32 r, 32 r + 32) and holds their rows of [W | W_in | W_fb] (K_aug x 32 elements, 139 KB in fp64 at
// N = 512 -> cluster of 16).  Every CTA keeps a full copy of the augmented state [x; u; y] of its frames;
// per step it computes its 32 neurons (8 warps split the K range, shared-memory reduction), applies
// tanh + noise, and writes the 32 new values -- and, in predict mode, its partial sums of the readout --
// into the next-step buffers of ALL CTAs through distributed shared memory; one cluster barrier per step
// publishes them.  State buffers are double-buffered, so no second barrier is needed.
#include <algorithm>
#include "common.cuh"
#include "recurrence_simt.cuh"        // RecParams

namespace {

using esn_simt::RecParams;

constexpr int CL_NS = 32;             // neurons per CTA (= one warp lane per neuron)
constexpr int CL_THREADS = 256;       // 8 warps split the K range
constexpr int CL_WARPS = CL_THREADS / 32;

__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_mapa(const void *p, uint32_t rank) {
    uint32_t r, a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_store(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cl_store(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared-memory plan (elements of T): Ws [Kp][32] | v [2][FB][Kp] | yp [2][CS][FB*n_out] | red [8][FB][32]
// | wos [FB][n_out][32] | wou [FB][n_out][n_in] | part [FB*n_out] | yu [FB*n_out]
template <typename T>
size_t cl_smem_bytes(int Kp, int FB, int CS, int n_in, int n_out) {
    size_t el = (size_t)Kp * CL_NS + 2 * (size_t)Kp * FB + 2 * (size_t)CS * FB * n_out +
                (size_t)CL_WARPS * FB * CL_NS + (size_t)FB * n_out * CL_NS + (size_t)FB * n_out * n_in +
                2 * (size_t)FB * n_out;
    return el * sizeof(T) + 64;
}

// four consecutive elements from 16-byte aligned shared memory (one LDS.128 in fp32, two in fp64)
__device__ __forceinline__ void cl_load4(const float *p, float (&o)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void cl_load4(const double *p, double (&o)[4]) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

template <typename T, int FB>
__global__ void __launch_bounds__(CL_THREADS, 1) esn_recurrence_cluster(const RecParams p, int CS) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.N, n_in = p.n_in, n_out = p.n_out, P = N + n_in, Kp = p.K_aug_pad;
    const int FO = FB * n_out;
    T *Ws = reinterpret_cast<T *>(smem_raw);
    T *v = Ws + (size_t)Kp * CL_NS;                        // [2][FB][Kp]: frame-major, so that the 32 neurons a
    T *yp = v + 2 * (size_t)Kp * FB;                       // CTA publishes are contiguous in every peer's copy
    T *red = yp + 2 * (size_t)CS * FO;
    T *wos = red + (size_t)CL_WARPS * FB * CL_NS;
    T *wou = wos + (size_t)FO * CL_NS;
    T *part = wou + (size_t)FO * n_in;
    T *yu = part + FO;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cl_rank();
    const int f0 = (blockIdx.x / CS) * FB;                 // first frame of this cluster
    const int nn = rank * CL_NS + lane;                    // neuron of this lane
    const bool predict = p.mode == ESN_MODE_PREDICT;
    const int s0 = predict ? 0 : 1;
    const int noise_rows = predict ? p.T : p.T - 1;
    const T namp = (T)p.noise_amp;
    const bool use_noise = p.noise_amp != 0.0;

    const T *Wt = static_cast<const T *>(p.Wt_aug);
    const T *gin = static_cast<const T *>(p.in);
    const T *in_scale = static_cast<const T *>(p.in_scale), *in_shift = static_cast<const T *>(p.in_shift);
    const T *t_scale = static_cast<const T *>(p.t_scale), *t_shift = static_cast<const T *>(p.t_shift);
    const T *teacher = static_cast<const T *>(p.teacher);
    const T *gW_out = static_cast<const T *>(p.W_out);
    const T *noise = static_cast<const T *>(p.noise);
    T *ext = static_cast<T *>(p.ext_out);
    T *yout = static_cast<T *>(p.y_out);

    // ---- one-time setup: weight slice, readout slices, zero state ----
    for (int i = tid; i < Kp * CL_NS; i += CL_THREADS) {
        const int k = i / CL_NS, c = i - k * CL_NS;
        Ws[i] = Wt[(size_t)k * p.N_pad + rank * CL_NS + c];
    }
    for (int i = tid; i < 2 * Kp * FB; i += CL_THREADS) v[i] = (T)0;
    for (int i = tid; i < 2 * CS * FO; i += CL_THREADS) yp[i] = (T)0;
    for (int i = tid; i < 2 * FO; i += CL_THREADS) part[i] = (T)0;
    if (predict) {
        for (int i = tid; i < FO * CL_NS; i += CL_THREADS) {
            const int f = i / (n_out * CL_NS), o = (i / CL_NS) % n_out, c = i % CL_NS, b = f0 + f;
            const int g = (p.group_ids && b < p.B) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;
            const int n = rank * CL_NS + c;
            wos[i] = (b < p.B && n < N) ? gW_out[((size_t)g * n_out + o) * P + n] : (T)0;
        }
        for (int i = tid; i < FO * n_in; i += CL_THREADS) {
            const int f = i / (n_out * n_in), o = (i / n_in) % n_out, j = i % n_in, b = f0 + f;
            const int g = (p.group_ids && b < p.B) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;
            wou[i] = b < p.B ? gW_out[((size_t)g * n_out + o) * P + N + j] : (T)0;
        }
    }
    __syncthreads();
    auto V = [&](int buf, int k, int f) -> T & { return v[((size_t)buf * FB + f) * Kp + k]; };
    // scaled inputs of time step `row` -> u rows of buffer `buf` (every CTA stages its own copy; CTA 0 writes E)
    auto stage_inputs = [&](int buf, int row, bool to_smem) {
        for (int i = tid; i < FB * n_in; i += CL_THREADS) {
            const int f = i / n_in, j = i - f * n_in, b = f0 + f;
            T val = (T)0;
            if (b < p.B && row < p.T) {
                val = gin[((size_t)b * p.T + row) * n_in + j] * in_scale[j] + in_shift[j];
                if (ext && rank == 0) ext[((size_t)b * p.T + row) * P + N + j] = val;
            }
            if (to_smem) V(buf, N + j, f) = val;
        }
    };
    auto stage_teacher = [&](int buf, int row) {
        for (int i = tid; i < FO; i += CL_THREADS) {
            const int f = i / n_out, o = i - f * n_out, b = f0 + f;
            T val = (T)0;
            if (p.feedback && b < p.B && row < p.T)
                val = teacher[((size_t)b * p.T + row) * n_out + o] * t_scale[o] + t_shift[o];
            V(buf, P + o, f) = val;
        }
    };
    int cur = 0;
    if (predict) {
        if (p.x0) {
            const T *x0 = static_cast<const T *>(p.x0);
            for (int i = tid; i < FB * N; i += CL_THREADS) {
                const int f = i / N, k = i - f * N, b = f0 + f;
                if (b < p.B) V(0, k, f) = x0[(size_t)b * N + k];
            }
        }
        if (p.y0 && p.feedback) {
            const T *y0 = static_cast<const T *>(p.y0);
            for (int i = tid; i < FO; i += CL_THREADS) {
                const int f = i / n_out, o = i - f * n_out, b = f0 + f;
                if (b < p.B) V(0, P + o, f) = y0[(size_t)b * n_out + o];
            }
        }
        stage_inputs(0, 0, true);
    } else {
        if (rank == 0)                                     // E row 0 = [0, u_0]; x_0 = 0
            for (int i = tid; i < FB * N; i += CL_THREADS) {
                const int f = i / N, k = i - f * N, b = f0 + f;
                if (b < p.B) ext[((size_t)b * p.T) * P + k] = (T)0;
            }
        stage_inputs(0, 0, false);
        stage_inputs(0, 1, true);
        stage_teacher(0, 0);
    }
    __syncthreads();
    cl_sync();                                             // every CTA's buffers exist before remote stores

    // K range of this warp: a multiple of 4 rows, so the state is read 16 bytes at a time
    const int kper = ((Kp + CL_WARPS - 1) / CL_WARPS + 3) / 4 * 4, kb = min(Kp, warp * kper), ke = min(Kp, kb + kper);
    // tanh stage: warp -> (frame, share of the outputs); the warps of one frame recompute the same state
    const int fw = warp % FB, qw = warp / FB;
    constexpr int NSUB = CL_WARPS / FB;
    const int bw = f0 + fw;
    for (int n = s0; n < p.T; ++n) {
        const int nrow = predict ? n : n - 1, nxt = cur ^ 1;
        // ---- 0. loads that do not depend on this step, issued ahead of the mat-vec and parked in registers:
        //         the next step's inputs (and teacher row), plus the input part of this step's readout ----
        T u_raw = (T)0, u_sc = (T)0, u_sh = (T)0, t_raw = (T)0, t_sc = (T)0, t_sh = (T)0;
        const bool has_u = tid < FB * n_in, has_t = !predict && tid < FO;
        const int uf = tid / n_in, uj = tid - uf * n_in, ub = f0 + uf;
        const int tf = tid / n_out, to = tid - tf * n_out, tb = f0 + tf;
        const bool u_live = has_u && ub < p.B && n + 1 < p.T;
        if (u_live) {
            u_raw = gin[((size_t)ub * p.T + n + 1) * n_in + uj];
            u_sc = in_scale[uj];
            u_sh = in_shift[uj];
        }
        if (has_t && p.feedback && tb < p.B) {
            t_raw = teacher[((size_t)tb * p.T + n) * n_out + to];
            t_sc = t_scale[to];
            t_sh = t_shift[to];
        }
        if (predict) {
            for (int i = tid; i < FO; i += CL_THREADS) {
                const int f = i / n_out;
                T y = (T)0;
                for (int j = 0; j < n_in; ++j) y = fma(wou[i * n_in + j], V(cur, N + j, f), y);
                yu[i] = y;
            }
        }
        // ---- 1. partial mat-vec of this warp's K slice, all frames ----
        T acc[FB];
#pragma unroll
        for (int f = 0; f < FB; ++f) acc[f] = (T)0;
        const T *vc = v + (size_t)cur * FB * Kp;
#pragma unroll 2
        for (int k = kb; k < ke; k += 4) {
            T w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = Ws[(k + i) * CL_NS + lane];
#pragma unroll
            for (int f = 0; f < FB; ++f) {
                T x4[4];
                cl_load4(vc + (size_t)f * Kp + k, x4);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[f] = fma(w[i], x4[i], acc[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < FB; ++f) red[(warp * FB + f) * CL_NS + lane] = acc[f];
        if (has_u) {                                       // the prefetched rows go into the NEXT buffer
            const T val = u_live ? u_raw * u_sc + u_sh : (T)0;
            if (u_live && ext && rank == 0) ext[((size_t)ub * p.T + n + 1) * P + N + uj] = val;
            V(nxt, N + uj, uf) = val;
        }
        if (has_t) V(nxt, P + to, tf) = t_raw * t_sc + t_sh;
        __syncthreads();
        // ---- 2a. finish: sum the partials of the warp's frame, tanh + noise, new state into the LOCAL buffer;
        //          in predict mode also this CTA's share of y_n = W_out [x_n; u_n] for the warp's outputs ----
        {
            T pre = (T)0;
#pragma unroll
            for (int w = 0; w < CL_WARPS; ++w) pre += red[(w * FB + fw) * CL_NS + lane];
            T x = (T)0;
            if (nn < N && bw < p.B) {
                x = esn_tanh<T>(pre);
                if (use_noise) {
                    T u;
                    if (noise) u = noise[((size_t)bw * noise_rows + nrow) * N + nn];
                    else u = (T)esn_noise_uniform(esn_noise_key(p.seed, (uint32_t)bw, (uint32_t)nrow), (uint32_t)nn);
                    x += namp * (u - (T)0.5);
                }
                if (qw == 0) {
                    if (ext) ext[((size_t)bw * p.T + n) * P + nn] = x;
                    V(nxt, nn, fw) = x;                    // pad lanes must not touch the u / y rows behind the state
                }
            }
            if (predict) {
                for (int o = qw; o < n_out; o += 2 * NSUB) {   // two outputs at a time for shuffle ILP
                    const int o2 = o + NSUB;
                    T sa = wos[(fw * n_out + o) * CL_NS + lane] * x;
                    T sb = o2 < n_out ? wos[(fw * n_out + o2) * CL_NS + lane] * x : (T)0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        sa += __shfl_xor_sync(0xffffffffu, sa, d);
                        sb += __shfl_xor_sync(0xffffffffu, sb, d);
                    }
                    if (lane == 0) {
                        part[fw * n_out + o] = sa;
                        if (o2 < n_out) part[fw * n_out + o2] = sb;
                    }
                }
            }
        }
        __syncthreads();
        // ---- 2b. publish, all warps: warp w serves the peer CTAs w, w + 8: 32 contiguous neurons per frame
        //          and (predict) the block of FB * n_out readout shares ----
        {
            T xs[FB];
#pragma unroll
            for (int f = 0; f < FB; ++f) xs[f] = nn < N ? V(nxt, nn, f) : (T)0;
            T ps[2];
            ps[0] = (predict && lane < FO) ? part[lane] : (T)0;
            ps[1] = (predict && lane + 32 < FO) ? part[lane + 32] : (T)0;
            for (int r = warp; r < CS; r += CL_WARPS) {
                if (r == rank) {
                    if (predict) {
                        T *ys = &yp[((size_t)nxt * CS + rank) * FO];
                        if (lane < FO) ys[lane] = ps[0];
                        if (lane + 32 < FO) ys[lane + 32] = ps[1];
                        for (int i = lane + 64; i < FO; i += 32) ys[i] = part[i];
                    }
                    continue;
                }
                if (nn < N) {
                    const uint32_t base = cl_mapa(&V(nxt, nn, 0), (uint32_t)r);
#pragma unroll
                    for (int f = 0; f < FB; ++f) cl_store(base + (uint32_t)(f * Kp * sizeof(T)), xs[f]);
                }
                if (predict) {
                    const uint32_t base = cl_mapa(&yp[((size_t)nxt * CS + rank) * FO], (uint32_t)r);
                    if (lane < FO) cl_store(base + (uint32_t)(lane * sizeof(T)), ps[0]);
                    if (lane + 32 < FO) cl_store(base + (uint32_t)((lane + 32) * sizeof(T)), ps[1]);
                    for (int i = lane + 64; i < FO; i += 32) cl_store(base + (uint32_t)(i * sizeof(T)), part[i]);
                }
            }
        }
        cl_sync();                                         // x_n (and the readout shares) are everywhere
        // ---- 3. y_n: sum of the CTAs' shares + the input part; fed back through the next buffer ----
        if (predict) {
            for (int i = tid; i < FO; i += CL_THREADS) {
                const int f = i / n_out, o = i - f * n_out, b = f0 + f;
                T y = yu[i];
                for (int r = 0; r < CS; ++r) y += yp[((size_t)nxt * CS + r) * FO + i];
                V(nxt, P + o, f) = p.feedback ? y : (T)0;
                if (rank == 0 && b < p.B && n >= p.transient)
                    yout[((size_t)b * (p.T - p.transient) + (n - p.transient)) * n_out + o] = (y - t_shift[o]) / t_scale[o];
            }
        }
        __syncthreads();
        cur = nxt;
    }
    cl_sync();                                             // nobody leaves while a peer may still store into it
}

template <typename T, int FB>
int cl_launch(const RecParams &p, int CS, cudaStream_t st) {
    const size_t smem = cl_smem_bytes<T>(p.K_aug_pad, FB, CS, p.n_in, p.n_out);
    if (smem > 227 * 1024 - 1024) return ESN_E_TOOLARGE;
    if (FB * p.n_in > CL_THREADS || FB * p.n_out > CL_THREADS) return ESN_E_UNSUPPORTED;   // one prefetch per thread
    auto kern = esn_recurrence_cluster<T, FB>;
    ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * ((p.B + FB - 1) / FB)), 1, 1);
    cfg.blockDim = dim3(CL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ESN_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p, CS));
    return esn_launch_status();
}

// clusters of CS CTAs that can be resident at once (a cluster of 16 needs a GPC with 16 free SMs)
template <typename T, int FB>
int cl_resident_clusters(const RecParams &p, int CS) {
    const size_t smem = cl_smem_bytes<T>(p.K_aug_pad, FB, CS, p.n_in, p.n_out);
    if (smem > 227 * 1024 - 1024) return 0;
    auto kern = esn_recurrence_cluster<T, FB>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * 64), 1, 1);
    cfg.blockDim = dim3(CL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cl_cluster_size(int N) {
    int CS = (N + CL_NS - 1) / CL_NS;                      // CTAs per cluster: 1, or an even count up to 16
    if (CS > 1 && (CS & 1)) ++CS;
    return CS;
}

}  // namespace

// Largest batch the cluster kernel should take (esn_recurrence_run's automatic choice).  One wave of
// resident clusters costs the same whatever its fill, and the streaming kernel is latency-bound (one
// CTA pulls the whole weight matrix through L2 every step) until its own grid fills the GPU: measured on
// B200 (profiles/r1_small_batch_crossover.txt) the streaming kernel takes as long as 11 (100 neurons) to 19
// (512 neurons) cluster waves in fp64 and 8 to 11 in fp32; the cluster kernel keeps batches up to ~0.8 of that.
int esn_cluster_auto_limit(const esn_simt::RecParams &p, int dtype) {
    if (p.N > 512) return 0;
    const int CS = cl_cluster_size(p.N);
    static int cache[2][17];                               // resident clusters + 1; 0 = not asked yet
    int &c = cache[dtype == ESN_F64 ? 1 : 0][CS];
    if (c == 0) c = 1 + (dtype == ESN_F64 ? cl_resident_clusters<double, 4>(p, CS) : cl_resident_clusters<float, 8>(p, CS));
    const int resident = c - 1;
    return dtype == ESN_F64 ? 4 * resident * (9 + CS / 3) : 8 * resident * (6 + CS / 5);
}

// Small-batch dispatch (called by esn_recurrence_run): returns ESN_E_UNSUPPORTED when the shape does not
// suit the cluster kernel (more than 512 neurons, or the weight slice does not fit).
int esn_cluster_launch(const esn_simt::RecParams &p, int dtype, cudaStream_t st) {
    if (p.N > 512) return ESN_E_UNSUPPORTED;
    const int CS = cl_cluster_size(p.N);
    if (dtype == ESN_F64) {
        if (p.B >= 3) return cl_launch<double, 4>(p, CS, st);
        if (p.B == 2) return cl_launch<double, 2>(p, CS, st);
        return cl_launch<double, 1>(p, CS, st);
    }
    if (p.B >= 5) return cl_launch<float, 8>(p, CS, st);
    if (p.B >= 3) return cl_launch<float, 4>(p, CS, st);
    if (p.B == 2) return cl_launch<float, 2>(p, CS, st);
    return cl_launch<float, 1>(p, CS, st);
}
