// Readout training and application in fp64.
//
// Replaces `np.linalg.pinv(extended_states[transient:]) @ teachers` of ESN.fit
// (reference libs/pyESN.py:189-192) by normal equations with lambda = 0 --
// primal Gram E^T E when rows >= cols, dual Gram E E^T when rows < cols -- a
// batched blocked Cholesky and the triangular solves, and the train-set
// prediction E W_out^T (libs/pyESN.py:212-213).  Everything accumulates in
// fp64: cond(E)^2 reaches 1e9..1e12.
#include <algorithm>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ SYRK ----
// C[i][j] = sum_k A(k,i) A(k,j) over one or more frames, lower-triangular tiles
// only, mirrored on store.  PRIMAL: A(k,i) = E[row k][col i]; DUAL: A(k,i) =
// E[row i][col k].  64x64 tile per CTA, 256 threads, 4x4 doubles per thread.
constexpr int TS = 64, TK = 16;

template <typename TE, bool DUAL>
__global__ void __launch_bounds__(256)
syrk_f64_kernel(const TE *__restrict__ ext, int T, int p, int transient, int frames_per_cta,
                int B, int shared, int accumulate, double *__restrict__ G) {
    const int m = T - transient;
    const int n = DUAL ? m : p;              // order of G
    const int kdim = DUAL ? p : m;           // contraction length
    // decode lower-triangular tile index
    int tl = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
    const int tj = tl - ti * (ti + 1) / 2;
    const int i0 = ti * TS, j0 = tj * TS;

    __shared__ double As[TK][TS + 2], Bs[TK][TS + 2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double acc[4][4] = {};
    const int fbeg = blockIdx.y * frames_per_cta;
    const int fend = min(B, fbeg + frames_per_cta);

    for (int b = fbeg; b < fend; ++b) {
        const TE *E = ext + ((size_t)b * T + transient) * p;   // [m][p]
        for (int k0 = 0; k0 < kdim; k0 += TK) {
            // load TK x TS panels of A for the i-tile and the j-tile
            for (int e = tid; e < TK * TS; e += 256) {
                int kk, ii;
                if (DUAL) { kk = e % TK; ii = e / TK; } else { ii = e % TS; kk = e / TS; }
                const int k = k0 + kk;
                double va = 0.0, vb = 0.0;
                if (k < kdim) {
                    const int ia = i0 + ii, ib = j0 + ii;
                    if (ia < n) va = (double)(DUAL ? E[(size_t)ia * p + k] : E[(size_t)k * p + ia]);
                    if (ib < n) vb = (double)(DUAL ? E[(size_t)ib * p + k] : E[(size_t)k * p + ib]);
                }
                As[kk][ii] = va;
                Bs[kk][ii] = vb;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                double a[4], bb[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) { a[r] = As[kk][ty + 16 * r]; bb[r] = Bs[kk][tx + 16 * r]; }
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], bb[c], acc[r][c]);
            }
            __syncthreads();
        }
    }
    double *Gb = G + (shared ? 0 : (size_t)fbeg * n * n);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = i0 + ty + 16 * r, j = j0 + tx + 16 * c;
            if (i < n && j < n && j <= i) {
                if (shared) {
                    atomicAdd(&Gb[(size_t)i * n + j], acc[r][c]);
                    if (i != j) atomicAdd(&Gb[(size_t)j * n + i], acc[r][c]);
                } else if (accumulate) {
                    Gb[(size_t)i * n + j] += acc[r][c];
                    if (i != j) Gb[(size_t)j * n + i] += acc[r][c];
                } else {
                    Gb[(size_t)i * n + j] = acc[r][c];
                    Gb[(size_t)j * n + i] = acc[r][c];
                }
            }
        }
}

// The same Gram tile on the fp64 tensor cores (mma.sync m8n8k4, DMMA): 8 warps, each a 16 x 32 sub-tile =
// 2 x 4 accumulator fragments; per k4 step a warp loads 2 A and 4 B fragments (one double per lane each)
// for 8 MMAs, i.e. 6 shared-memory loads per 2048 FMAs where the FMA kernel above needs 8 per 512 (it is
// bound by shared-memory bandwidth at ~40 % of the fp64 rate).  Fragment layouts (PTX ISA, m8n8k4 .f64):
// A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4], C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
// Panels are stored so that the global-memory fetch order is also the conflict-free store order: [k][i]
// with a row stride of TS + 4 doubles in the primal case (lanes walk i), [i][k] with a row stride of DK + 4
// in the dual case (lanes walk k; a [k][i] store there is an 8-way bank conflict that costs as much as the
// MMAs).  With either stride the 4 x 4 doubles a half-warp reads per fragment fall into distinct banks.
// The next panel is fetched into registers while this one is multiplied.
constexpr int DK = 32, DLD = TS + 4, DLT = DK + 4;
constexpr int DPANEL = DK * DLD > TS * DLT ? DK * DLD : TS * DLT;

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(a), "l"(gmem));
}
__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <typename TE, bool DUAL>
__global__ void __launch_bounds__(256, 3)        // 80 registers, no spills: three CTAs per SM (11.5 -> 7.5 ms per 1184 pilots)
syrk_dmma_kernel(const TE *__restrict__ ext, int T, int p, int transient, int frames_per_cta,
                 int B, int shared, int accumulate, double *__restrict__ G) {
    const int m = T - transient;
    const int n = DUAL ? m : p;              // order of G
    const int kdim = DUAL ? p : m;           // contraction length
    int tl = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
    const int tj = tl - ti * (ti + 1) / 2;
    const int i0 = ti * TS, j0 = tj * TS;
    const bool diag = ti == tj;              // the j panel IS the i panel

    __shared__ double As[DPANEL], Bs[DPANEL];
    auto at = [](int k, int i) { return DUAL ? i * DLT + k : k * DLD + i; };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;            // sub-tile of this warp
    const int fk = lane & 3, fi = lane >> 2;
    double acc[2][4][2] = {};
    const int fbeg = blockIdx.y * frames_per_cta;
    const int fend = min(B, fbeg + frames_per_cta);
    constexpr int PER = DK * TS / 256;                               // panel elements per thread
    double va[PER], vb[PER];

    auto fetch = [&](const TE *E, int k0) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * 256;
            int kk, ii;
            if (DUAL) { kk = e % DK; ii = e / DK; } else { ii = e % TS; kk = e / TS; }
            const int k = k0 + kk, ia = i0 + ii, ib = j0 + ii;
            va[q] = (k < kdim && ia < n) ? (double)(DUAL ? E[(size_t)ia * p + k] : E[(size_t)k * p + ia]) : 0.0;
            vb[q] = (!diag && k < kdim && ib < n) ? (double)(DUAL ? E[(size_t)ib * p + k] : E[(size_t)k * p + ib]) : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * 256;
            int kk, ii;
            if (DUAL) { kk = e % DK; ii = e / DK; } else { ii = e % TS; kk = e / TS; }
            As[at(kk, ii)] = va[q];
            if (!diag) Bs[at(kk, ii)] = vb[q];
        }
    };
    const double *Bp = diag ? As : Bs;
    const int nchunk = (kdim + DK - 1) / DK;
    for (int b = fbeg; b < fend; ++b) {
        const TE *E = ext + ((size_t)b * T + transient) * p;   // [m][p]
        fetch(E, 0);
        for (int ch = 0; ch < nchunk; ++ch) {
            __syncthreads();                                         // previous panel fully consumed
            stash();
            __syncthreads();
            if (ch + 1 < nchunk) fetch(E, (ch + 1) * DK);            // in flight during the MMAs
#pragma unroll
            for (int k4 = 0; k4 < DK; k4 += 4) {
                double a[2], bb[4];
#pragma unroll
                for (int r = 0; r < 2; ++r) a[r] = As[at(k4 + fk, r0 + 8 * r + fi)];
#pragma unroll
                for (int c = 0; c < 4; ++c) bb[c] = Bp[at(k4 + fk, c0 + 8 * c + fi)];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) dmma8x8x4(acc[r][c], a[r], bb[c]);
            }
        }
    }
    double *Gb = G + (shared ? 0 : (size_t)fbeg * n * n);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + r0 + 8 * r + fi, j = j0 + c0 + 8 * c + 2 * fk + h;
                const double v = acc[r][c][h];
                if (i < n && j < n && j <= i) {
                    if (shared) {
                        atomicAdd(&Gb[(size_t)i * n + j], v);
                        if (i != j) atomicAdd(&Gb[(size_t)j * n + i], v);
                    } else if (accumulate) {
                        Gb[(size_t)i * n + j] += v;
                        if (i != j) Gb[(size_t)j * n + i] += v;
                    } else {
                        Gb[(size_t)i * n + j] = v;
                        Gb[(size_t)j * n + i] = v;
                    }
                }
            }
}

// ------------------------------------------------------------------- X^T Y --
// out(i, o) = sum_r E[r][i] * Y(r, o), r over the m kept rows of each frame.
// Y is either a scaled teacher (TY raw, scale/shift applied) or a plain fp64
// matrix [m][n_out] (the dual solution A).  Output strides let the caller write
// R [p][n_out] or W_out [n_out][p] directly.
template <typename TE, typename TY>
__global__ void __launch_bounds__(128)
xty_f64_kernel(const TE *__restrict__ ext, const TY *__restrict__ Y, int y_is_teacher,
               const double *__restrict__ t_scale, const double *__restrict__ t_shift,
               int T, int p, int n_out, int transient, int frames_per_cta, int B, int shared,
               int accumulate, double *__restrict__ out, int out_si, int out_so, size_t out_sb) {
    const int m = T - transient;
    const int i = blockIdx.x * 128 + threadIdx.x;
    __shared__ double ys[32][ESN_MAX_OUT];
    double acc[ESN_MAX_OUT] = {};
    const int fbeg = blockIdx.y * frames_per_cta, fend = min(B, fbeg + frames_per_cta);
    for (int b = fbeg; b < fend; ++b) {
        const TE *E = ext + ((size_t)b * T + transient) * p;
        for (int r0 = 0; r0 < m; r0 += 32) {
            for (int e = threadIdx.x; e < 32 * n_out; e += 128) {
                int rr = e / n_out, o = e - rr * n_out, r = r0 + rr;
                double v = 0.0;
                if (r < m) {
                    if (y_is_teacher)
                        v = (double)Y[((size_t)b * T + transient + r) * n_out + o] * t_scale[o] + t_shift[o];
                    else
                        v = (double)Y[((size_t)b * m + r) * n_out + o];
                }
                ys[rr][o] = v;
            }
            __syncthreads();
            if (i < p) {
                const int rmax = min(32, m - r0);
                for (int rr = 0; rr < rmax; ++rr) {
                    const double e = (double)E[(size_t)(r0 + rr) * p + i];
#pragma unroll
                    for (int o = 0; o < ESN_MAX_OUT; ++o)
                        if (o < n_out) acc[o] = fma(e, ys[rr][o], acc[o]);
                }
            }
            __syncthreads();
        }
    }
    if (i < p) {
        double *ob = out + (shared ? 0 : (size_t)fbeg * out_sb);
#pragma unroll
        for (int o = 0; o < ESN_MAX_OUT; ++o)
            if (o < n_out) {
                double *dst = &ob[(size_t)i * out_si + (size_t)o * out_so];
                if (shared) atomicAdd(dst, acc[o]);
                else if (accumulate) *dst += acc[o];
                else *dst = acc[o];
            }
    }
}

template <typename TY>
__global__ void scaled_teacher_rows_kernel(const TY *__restrict__ teacher, const double *__restrict__ t_scale,
                                           const double *__restrict__ t_shift, int B, int T, int n_out,
                                           int transient, double *__restrict__ rhs) {
    const int m = T - transient;
    const size_t total = (size_t)B * m * n_out;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int o = (int)(e % n_out);
        size_t br = e / n_out;
        int r = (int)(br % m);
        size_t b = br / m;
        rhs[e] = (double)teacher[(b * T + transient + r) * n_out + o] * t_scale[o] + t_shift[o];
    }
}

// -------------------------------------------------------------- Cholesky ----
// One CTA per problem, matrix in global memory, LEFT-looking blocked factorisation with NB = 32: panel k first
// receives the contributions of all previous panels,
//     A[r][k0 + c] -= sum_{q < k0} L[r][q] L[k0 + c][q],   r >= k0,
// on the fp64 tensor cores (mma.sync.m8n8k4.f64): a warp owns a 16 x 32 tile of the panel, the 32 pivot rows
// L[k0 .. k0+31][q] are staged in shared memory in chunks of 32 columns (shared by all warps), the tile's own rows
// L[r][q] go straight from L2 / HBM into A fragments (each element is used by exactly one warp), prefetched one
// chunk ahead.  Then the diagonal block is factored in shared memory and the rows below are solved against it, one
// row per thread.  Every element of L is written once and read ~k/2 times as a tile row: 6.6 MB of traffic per
// 512 x 512 problem, where the right-looking update this replaces rewrote the whole trailing matrix per panel
// (32 MB per problem; with ~450 problems in flight that is HBM traffic, not L2).  Forward and backward substitution
// for n_rhs <= ESN_MAX_OUT right-hand sides follow.  Lower triangle is used/written.
// NG groups of 128 threads share a problem (more warps over the same tiles): NG = 1 with four CTAs per SM when
// there are many problems, NG = 2 when the batch does not fill the GPU (the drop-in ESN.fit is a batch of one).
constexpr int NB = 32;
#ifndef CH_GROUP_T
#define CH_GROUP_T 128
#endif
#ifndef CH_MINB
#define CH_MINB 4
#endif
constexpr int CH_GROUP = CH_GROUP_T;
constexpr int CKK = 32;                       // K chunk of the left-looking update
constexpr int PSTR = CKK + 4;                 // row stride of the staged pivot rows: half-warp LDS.64 conflict-free

template <int NG>
__global__ void __launch_bounds__(CH_GROUP * NG, NG == 1 ? CH_MINB : CH_MINB / 2)
cholesky_solve_f64_kernel(double *__restrict__ Gall, double *__restrict__ rhs_all, int n, int n_rhs,
                          int *__restrict__ info_all, double *__restrict__ pivots) {
    constexpr int CH_THREADS = CH_GROUP * NG;
    constexpr int NWARP = CH_THREADS / 32;
    double *A = Gall + (size_t)blockIdx.x * n * n;
    double *Bm = rhs_all + (size_t)blockIdx.x * n * n_rhs;
    __shared__ double D[NB][NB + 1];          // diagonal block / its factor
    __shared__ double Pc[2][NB][PSTR];        // pivot rows of the current / next K chunk
    __shared__ double Z[NB][ESN_MAX_OUT];     // substitution: the block's solution
    __shared__ double Rd[NB];                 // reciprocals of the current block's pivots l_cc
    __shared__ int s_info;
    __shared__ double s_pmin, s_pmax;         // smallest / largest pivot d_jj = l_jj^2: max / min bounds cond(G) from below
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { s_info = 0; s_pmin = INFINITY; s_pmax = 0.0; }
    for (int e = tid; e < NB * ESN_MAX_OUT; e += CH_THREADS) Z[e / ESN_MAX_OUT][e % ESN_MAX_OUT] = 0.0;   // (rows / columns a short block leaves unwritten)
    __syncthreads();

    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        // 0. left-looking update of the panel (rows k0 .. n-1, columns k0 .. k0+nb-1) from the panels before it
        if (k0 > 0) {
            const int ntile = (n - k0 + 15) / 16, nchunk = k0 / CKK;       // k0 is a multiple of NB = CKK
            const int fr = lane >> 2, fk = lane & 3;                         // fragment row / k (A, B); C: row fr, cols 2 fk, 2 fk + 1
            constexpr int NST = NB * CKK / CH_THREADS;                       // staged pivot-row elements per thread and chunk
            // pivot rows x chunk c -> Pc[buf] by cp.async (no registers; in flight during the DMMAs of the chunk before)
            auto stage = [&](int c, int buf) {
#pragma unroll
                for (int i = 0; i < NST; ++i) {
                    const int e = tid + i * CH_THREADS, r = e / CKK, q = e % CKK;
                    if (r < nb) cp_async8(&Pc[buf][r][q], &A[(size_t)(k0 + r) * n + c * CKK + q]);
                    else Pc[buf][r][q] = 0.0;
                }
                cp_async_commit();
            };
            for (int t0 = 0; t0 < ntile; t0 += NWARP) {                      // a pass: one tile per warp
                const int tile = t0 + warp;
                const bool have = tile < ntile;
                const int r0 = k0 + tile * 16;
                // rows of the two 8-row fragments (clamped: loads stay inside the matrix, stores are masked)
                const double *ra = A + (size_t)min(r0 + fr, n - 1) * n + fk;
                const double *rb = A + (size_t)min(r0 + 8 + fr, n - 1) * n + fk;
                double acc[2][4][2] = {};
                double a_cur[2][CKK / 4], a_nxt[2][CKK / 4];
                if (have) {
#pragma unroll
                    for (int j = 0; j < CKK / 4; ++j) { a_cur[0][j] = ra[4 * j]; a_cur[1][j] = rb[4 * j]; }
                }
                stage(0, 0);
                cp_async_wait<0>();
                __syncthreads();
                for (int c = 0; c < nchunk; ++c) {
                    const int buf = c & 1;
                    const bool more = c + 1 < nchunk;
                    if (more) {
                        if (have) {
#pragma unroll
                            for (int j = 0; j < CKK / 4; ++j) {
                                a_nxt[0][j] = ra[(c + 1) * CKK + 4 * j];
                                a_nxt[1][j] = rb[(c + 1) * CKK + 4 * j];
                            }
                        }
                        stage(c + 1, buf ^ 1);
                    }
                    if (have) {
#pragma unroll
                        for (int j = 0; j < CKK / 4; ++j) {
                            double bf[4];
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) bf[nt] = Pc[buf][nt * 8 + fr][4 * j + fk];
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) {
                                dmma8x8x4(acc[0][nt], a_cur[0][j], bf[nt]);
                                dmma8x8x4(acc[1][nt], a_cur[1][j], bf[nt]);
                            }
                        }
                    }
                    if (more) {
                        cp_async_wait<0>();
                        if (have) {
#pragma unroll
                            for (int j = 0; j < CKK / 4; ++j) { a_cur[0][j] = a_nxt[0][j]; a_cur[1][j] = a_nxt[1][j]; }
                        }
                    }
                    __syncthreads();                                         // Pc[buf] consumed, Pc[buf ^ 1] staged
                }
                if (have) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = r0 + 8 * h + fr;
                        if (r < n) {
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt) {
                                const int cc = nt * 8 + 2 * fk;
                                double *dst = A + (size_t)r * n + k0 + cc;
                                if (cc + 1 < nb && !(n & 1)) {
                                    double2 v = *reinterpret_cast<double2 *>(dst);
                                    v.x -= acc[h][nt][0];
                                    v.y -= acc[h][nt][1];
                                    *reinterpret_cast<double2 *>(dst) = v;
                                } else {
                                    if (cc < nb) dst[0] -= acc[h][nt][0];
                                    if (cc + 1 < nb) dst[1] -= acc[h][nt][1];
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        // 1. diagonal block -> shared, unblocked Cholesky
        for (int e = tid; e < NB * NB; e += CH_THREADS) {
            int r = e / NB, c = e % NB;
            D[r][c] = (r < nb && c < nb && c <= r) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
        __syncthreads();
        {
            // all warps: thread = (row r, column class cg); per column j two block barriers -- the column is scaled by
            // l_jj by class 0, every class subtracts l_rj l_cj from its own columns c > j
            const int r = lane, cg = warp;
            double pmin = INFINITY, pmax = 0.0;
            int bad = 0;
            for (int j = 0; j < nb; ++j) {
                double d = D[j][j];
                pmin = fmin(pmin, d);
                pmax = fmax(pmax, d);
                if (!(d > 0.0)) {
                    if (bad == 0) bad = k0 + j + 1;
                    d = nan("");
                }
                const double sj = sqrt(d);
                const double lr = (r > j && r < nb) ? D[r][j] / sj : 0.0;
                double lc[(NB + NWARP - 1) / NWARP];                 // l_cj of this class's columns: lane c of the warp holds it
#pragma unroll
                for (int q = 0; q < (NB + NWARP - 1) / NWARP; ++q) {
                    const int c = j + 1 + cg + q * NWARP;
                    lc[q] = __shfl_sync(0xffffffffu, lr, c & 31);
                }
                __syncthreads();                                     // column j has been read by everyone
                if (cg == 0) {
                    if (r == j) D[j][j] = sj;
                    else if (r > j && r < nb) D[r][j] = lr;
                }
#pragma unroll
                for (int q = 0; q < (NB + NWARP - 1) / NWARP; ++q) {
                    const int c = j + 1 + cg + q * NWARP;
                    if (c < nb && r >= c && r < nb) D[r][c] -= lr * lc[q];
                }
                __syncthreads();                                     // column j + 1 is final
            }
            if (tid == 0) {
                s_pmin = fmin(s_pmin, pmin);
                s_pmax = fmax(s_pmax, pmax);
                if (bad && s_info == 0) s_info = bad;
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += CH_THREADS) {
            int r = e / nb, c = e % nb;
            if (c <= r) A[(size_t)(k0 + r) * n + k0 + c] = D[r][c];
        }
        // 2. panel: rows below the block, A[i, k0:k0+nb] <- A[i, k0:k0+nb] L_kk^-T (one row per thread)
        const int below = n - (k0 + nb);
        // (32 divisions per row by the same 32 pivots: reciprocal, one residual correction -- the quotient the divide
        // instruction sequence itself would return, without its special-case handling)
        if (tid < NB) Rd[tid] = tid < nb ? 1.0 / D[tid][tid] : 0.0;
        __syncthreads();
        for (int i = tid; i < below; i += CH_THREADS) {
            double *row = A + (size_t)(k0 + nb + i) * n + k0;
            double x[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) x[c] = c < nb ? row[c] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < nb) {
                    double s = x[c];
#pragma unroll
                    for (int q = 0; q < NB; ++q)
                        if (q < c) s -= x[q] * D[c][q];
                    const double dcc = D[c][c], rc = Rd[c];
                    double qv = s * rc;
                    qv = fma(fma(-qv, dcc, s), rc, qv);
                    x[c] = fma(fma(-qv, dcc, s), rc, qv);
                }
            }
#pragma unroll
            for (int c = 0; c < NB; ++c)
                if (c < nb) row[c] = x[c];
        }
        __syncthreads();
    }

    // ---- forward substitution L Z = B (blocked by NB) ----
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        for (int e = tid; e < NB * NB; e += CH_THREADS) {
            int r = e / NB, c = e % NB;
            D[r][c] = (r < nb && c < nb && c <= r) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
        __syncthreads();
        // the nb x nb block: a warp per right-hand side, lane = row, column-oriented -- z_c is final once the columns
        // before it have been subtracted; lane c broadcasts it and the rows below take their share
        for (int o = warp; o < n_rhs; o += NWARP) {
            double bv = lane < nb ? Bm[(size_t)(k0 + lane) * n_rhs + o] : 0.0;
            const double dii = lane < nb ? D[lane][lane] : 1.0;
            double zr = 0.0;
            for (int c = 0; c < nb; ++c) {
                const double zc = __shfl_sync(0xffffffffu, bv / dii, c);
                if (lane == c) zr = zc;
                if (lane > c && lane < nb) bv -= D[lane][c] * zc;
            }
            if (lane < nb) {
                Z[lane][o] = zr;
                Bm[(size_t)(k0 + lane) * n_rhs + o] = zr;
            }
        }
        __syncthreads();
        // update the rows below: B[i] -= L[i, k0:k0+nb] z.  One row per thread (its 32 entries of L are loaded once, in
        // flight together, and meet all right-hand sides), eight right-hand sides at a time in registers.
        const int below = n - (k0 + nb);
        for (int i = tid; i < below; i += CH_THREADS) {
            const double *Lrow = A + (size_t)(k0 + nb + i) * n + k0;
            double l[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) l[c] = c < nb ? Lrow[c] : 0.0;
            double *brow = Bm + (size_t)(k0 + nb + i) * n_rhs;
            for (int o0 = 0; o0 < n_rhs; o0 += 8) {
                double sacc[8] = {};
#pragma unroll
                for (int c = 0; c < NB; ++c)
#pragma unroll
                    for (int o = 0; o < 8; ++o) sacc[o] = fma(l[c], Z[c][o0 + o], sacc[o]);
#pragma unroll
                for (int o = 0; o < 8; ++o)
                    if (o0 + o < n_rhs) brow[o0 + o] -= sacc[o];
            }
        }
        __syncthreads();
    }
    // ---- backward substitution L^T X = Z ----
    const int nblk = (n + NB - 1) / NB;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int k0 = kb * NB, nb = min(NB, n - k0);
        for (int e = tid; e < NB * NB; e += CH_THREADS) {
            int r = e / NB, c = e % NB;
            D[r][c] = (r < nb && c < nb && c <= r) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
        __syncthreads();
        for (int o = warp; o < n_rhs; o += NWARP) {               // L_kk^T x = z: the same, from the last column up
            double bv = lane < nb ? Bm[(size_t)(k0 + lane) * n_rhs + o] : 0.0;
            const double dii = lane < nb ? D[lane][lane] : 1.0;
            double xr = 0.0;
            for (int c = nb - 1; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, bv / dii, c);
                if (lane == c) xr = xc;
                if (lane < c) bv -= D[c][lane] * xc;
            }
            if (lane < nb) {
                Z[lane][o] = xr;
                Bm[(size_t)(k0 + lane) * n_rhs + o] = xr;
            }
        }
        __syncthreads();
        // rows above: B[i] -= sum_c L[k0+c][i] x[c], i < k0: one row i per thread (the loads of a column block are
        // coalesced over i), eight right-hand sides at a time
        for (int i = tid; i < k0; i += CH_THREADS) {
            double l[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) l[c] = c < nb ? A[(size_t)(k0 + c) * n + i] : 0.0;
            double *brow = Bm + (size_t)i * n_rhs;
            for (int o0 = 0; o0 < n_rhs; o0 += 8) {
                double sacc[8] = {};
#pragma unroll
                for (int c = 0; c < NB; ++c)
#pragma unroll
                    for (int o = 0; o < 8; ++o) sacc[o] = fma(l[c], Z[c][o0 + o], sacc[o]);
#pragma unroll
                for (int o = 0; o < 8; ++o)
                    if (o0 + o < n_rhs) brow[o0 + o] -= sacc[o];
            }
        }
        __syncthreads();
    }
    if (tid == 0 && info_all) info_all[blockIdx.x] = s_info;
    if (tid == 0 && pivots) { pivots[2 * blockIdx.x] = s_pmin; pivots[2 * blockIdx.x + 1] = s_pmax; }
}

// ---------------------------------------------------------- apply readout ---
template <typename T>
__global__ void __launch_bounds__(256)
apply_readout_kernel(const T *__restrict__ ext, const T *__restrict__ W_out, const int *__restrict__ group_ids,
                     const T *__restrict__ t_scale, const T *__restrict__ t_shift, size_t rows, int Trows, int p,
                     int n_out, T *__restrict__ pred) {
    const int lane = threadIdx.x & 31;
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / Trows);
    const int g = group_ids ? group_ids[b] : 0;
    const T *e = ext + row * p;
    const T *w = W_out + (size_t)g * n_out * p;
    T acc[ESN_MAX_OUT];
#pragma unroll
    for (int o = 0; o < ESN_MAX_OUT; ++o) acc[o] = (T)0;
    for (int k = lane; k < p; k += 32) {
        const T ev = e[k];
#pragma unroll
        for (int o = 0; o < ESN_MAX_OUT; ++o)
            if (o < n_out) acc[o] = fma(w[(size_t)o * p + k], ev, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < ESN_MAX_OUT; ++o) {
        if (o < n_out) {
            T v = acc[o];
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if (lane == 0) pred[row * n_out + o] = (v - t_shift[o]) / t_scale[o];
        }
    }
}

__global__ void transpose_rhs_kernel(const double *__restrict__ rhs, int B, int p, int n_out,
                                     double *__restrict__ W_out) {
    const size_t total = (size_t)B * p * n_out;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(e % p);
        size_t bo = e / p;
        int o = (int)(bo % n_out);
        size_t b = bo / n_out;
        W_out[e] = rhs[(b * p + i) * n_out + o];
    }
}

}  // namespace

extern "C" int esn_gram_f64(const void *ext, int ext_dtype, const void *teacher, int teacher_dtype,
                            const double *t_scale, const double *t_shift, int B, int T, int p,
                            int n_out, int transient, int dual, int shared, int accumulate,
                            double *G, double *rhs, void *stream) {
    if (!ext || !teacher || !t_scale || !t_shift || !G || !rhs) return ESN_E_BADARG;
    if (B <= 0 || p <= 0 || n_out <= 0 || n_out > ESN_MAX_OUT || transient < 0 || transient >= T)
        return ESN_E_BADARG;
    if (dual && (shared || accumulate)) return ESN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int m = T - transient;
    const int n = dual ? m : p;
    const int nt = (n + TS - 1) / TS;
    // shared readout: every CTA sums 8 frames in registers before it adds its tile to G -- with 45 tiles and three
    // resident CTAs per SM that is >= 10 full waves from a few hundred pilots on (a grid of 1.3 waves, 92 frames per
    // CTA, took 30 ms per 1184 pilots where the per-pilot Gram takes 13), and the tile CTAs of the same frames run
    // side by side, so the frames are read from HBM once
    const int fpc = shared ? 8 : 1;
    dim3 grid(nt * (nt + 1) / 2, (B + fpc - 1) / fpc);
    const bool f32 = ext_dtype == ESN_F32;
    if (shared && !accumulate) {
        ESN_CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)n * n, st));
        ESN_CUDA_TRY(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)p * n_out, st));
    }
    // fp64 tensor cores (DMMA); ESN_GRAM_FMA=1 selects the plain FMA kernel (kept as the cross-check)
    static const bool use_fma = getenv("ESN_GRAM_FMA") != nullptr;
#define SYRK(TE, D) (use_fma ? syrk_f64_kernel<TE, D><<<grid, 256, 0, st>>>((const TE *)ext, T, p, transient, fpc, B, shared, accumulate, G) : syrk_dmma_kernel<TE, D><<<grid, 256, 0, st>>>((const TE *)ext, T, p, transient, fpc, B, shared, accumulate, G))
#define SYRK_OLD(TE, D) syrk_f64_kernel<TE, D><<<grid, 256, 0, st>>>((const TE *)ext, T, p, transient, fpc, B, shared, accumulate, G)
    if (dual) { if (f32) SYRK(float, true); else SYRK(double, true); }
    else      { if (f32) SYRK(float, false); else SYRK(double, false); }
#undef SYRK
#undef SYRK_OLD
    int rc = esn_launch_status();
    if (rc) return rc;
    if (dual) {
        const size_t total = (size_t)B * m * n_out;
        int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)(4096));
        if (teacher_dtype == ESN_F32)
            scaled_teacher_rows_kernel<float><<<blocks, 256, 0, st>>>((const float *)teacher, t_scale, t_shift, B, T, n_out, transient, rhs);
        else
            scaled_teacher_rows_kernel<double><<<blocks, 256, 0, st>>>((const double *)teacher, t_scale, t_shift, B, T, n_out, transient, rhs);
    } else {
        dim3 g2((p + 127) / 128, (B + fpc - 1) / fpc);
#define XTY(TE, TY) xty_f64_kernel<TE, TY><<<g2, 128, 0, st>>>((const TE *)ext, (const TY *)teacher, 1, t_scale, t_shift, T, p, n_out, transient, fpc, B, shared, accumulate, rhs, n_out, 1, (size_t)p * n_out)
        if (f32) { if (teacher_dtype == ESN_F32) XTY(float, float); else XTY(float, double); }
        else     { if (teacher_dtype == ESN_F32) XTY(double, float); else XTY(double, double); }
#undef XTY
    }
    return esn_launch_status();
}

extern "C" int esn_cholesky_solve_f64(double *G, double *rhs, int batch, int n, int n_rhs, int32_t *info,
                                      void *stream) {
    return esn_cholesky_solve_piv_f64(G, rhs, batch, n, n_rhs, info, nullptr, stream);
}

extern "C" int esn_cholesky_solve_piv_f64(double *G, double *rhs, int batch, int n, int n_rhs, int32_t *info,
                                          double *pivots, void *stream) {
    if (!G || !rhs || batch <= 0 || n <= 0 || n_rhs <= 0 || n_rhs > ESN_MAX_OUT) return ESN_E_BADARG;
    // many problems: 128 threads each, four CTAs per SM (the diagonal factor, the row solves and the substitutions are
    // latency-bound and overlap across problems: 8.3 ms per 1184 problems of 512 x 512 against 11.7 ms with 256
    // threads x 2; five CTAs of 96 registers spill: 11.4 ms); no more problems than SMs: two groups per problem
    cudaStream_t st = (cudaStream_t)stream;
    if (batch <= 148) cholesky_solve_f64_kernel<2><<<batch, CH_GROUP * 2, 0, st>>>(G, rhs, n, n_rhs, info, pivots);
    else cholesky_solve_f64_kernel<1><<<batch, CH_GROUP, 0, st>>>(G, rhs, n, n_rhs, info, pivots);
    return esn_launch_status();
}

extern "C" int esn_readout_from_dual_f64(const void *ext, int ext_dtype, const double *A, int B, int T,
                                         int p, int n_out, int transient, double *W_out, void *stream) {
    if (!ext || !A || !W_out || B <= 0 || n_out <= 0 || n_out > ESN_MAX_OUT) return ESN_E_BADARG;
    dim3 g2((p + 127) / 128, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (ext_dtype == ESN_F32)
        xty_f64_kernel<float, double><<<g2, 128, 0, st>>>((const float *)ext, A, 0, nullptr, nullptr, T, p, n_out, transient, 1, B, 0, 0, W_out, 1, p, (size_t)p * n_out);
    else
        xty_f64_kernel<double, double><<<g2, 128, 0, st>>>((const double *)ext, A, 0, nullptr, nullptr, T, p, n_out, transient, 1, B, 0, 0, W_out, 1, p, (size_t)p * n_out);
    return esn_launch_status();
}

extern "C" int esn_transpose_rhs_f64(const double *rhs, int B, int p, int n_out, double *W_out, void *stream) {
    if (!rhs || !W_out || B <= 0) return ESN_E_BADARG;
    const size_t total = (size_t)B * p * n_out;
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)(4096));
    transpose_rhs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rhs, B, p, n_out, W_out);
    return esn_launch_status();
}

extern "C" int esn_apply_readout(int dtype, const void *ext, const void *W_out, const int32_t *group_ids,
                                 const void *t_scale, const void *t_shift, int B, int T, int p, int n_out,
                                 void *pred, void *stream) {
    if (!ext || !W_out || !t_scale || !t_shift || !pred || B <= 0 || T <= 0 || n_out > ESN_MAX_OUT)
        return ESN_E_BADARG;
    const size_t rows = (size_t)B * T;
    const int blocks = (int)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ESN_F32)
        apply_readout_kernel<float><<<blocks, 256, 0, st>>>((const float *)ext, (const float *)W_out, group_ids, (const float *)t_scale, (const float *)t_shift, rows, T, p, n_out, (float *)pred);
    else if (dtype == ESN_F64)
        apply_readout_kernel<double><<<blocks, 256, 0, st>>>((const double *)ext, (const double *)W_out, group_ids, (const double *)t_scale, (const double *)t_shift, rows, T, p, n_out, (double *)pred);
    else
        return ESN_E_BADARG;
    return esn_launch_status();
}
