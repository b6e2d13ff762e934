// Reservoir recurrence on the tensor cores with the STATE STREAMED through L2 (tcgen05 + TMEM), sm_100a.
//
// Same arithmetic as recurrence_tc.cu (free-running ESN.predict, reference libs/pyESN.py:243-255, and the
// teacher-forced harvest of ESN.fit, :179-182; fp16 hi/lo operand split, three kind::f16 MMAs per product,
// fp32 accumulation in TMEM, frames on the M side of a cta_group::2 pair), for the two cases the resident
// kernel cannot take:
//   * reservoirs of more than 512 neurons (the 4x8 fast demo's 600, the sweep's 1024 / 2048:
//     system_model_2/Demo_MIMO_4x8_ChannelRank_TrainSNR_LDPC_fast.py:142, BASELINE.json configs[3]) -- the
//     state of 64 frames no longer fits one SM next to a weight ring;
//   * a different readout for every frame (the demos train a fresh readout every L = 19 symbols,
//     OFDM_MIMO_2-2_NBF_LDPC.py:151-153,270: 18 data frames per W_out, which never aligns with 64-frame tiles).
//
// Layout.  A pair of CTAs owns 128 frames, 64 per CTA.  The state x_{t-1} of a CTA's frames lives in GLOBAL
// memory (L2-resident) as UMMA-ready tiles -- per 64-neuron chunk one 16 KB item [hi 8 KB | lo 8 KB], rows =
// frames, K-major SWIZZLE_128B, the byte image a tensor-map copy can drop into shared memory -- in two
// buffers (step t reads buffer (t+1)&1, its epilogue writes buffer t&1).  Per step the neurons are worked
// off in PASSES of one 256-neuron group (one N = 256 MMA shape, 128 TMEM columns per accumulator): every pass
// streams all state chunks (ring A) against the group's weight tiles (ring B, CTA r fetches the slab 2j + r of
// the shared image), then the resident aug chunk [u_t | y_{t-1}].
//
// Accumulation.  The tensor core adds each K = 16 product block to the fp32 accumulator with truncation, so a
// chain of n MMAs into one accumulator loses ~n^1.5 half-ulps, all in the same direction: 105 MMAs per step at
// 512 neurons are the 6e-6 per step of the resident kernel, and the chain grows with the reservoir.  Here a
// pass owns NACC accumulators: the small correction products (lo*hi, hi*lo: 2^-11 of the main term) go to
// their own accumulator, where the truncation is 2^-11 smaller, and the main products hi*hi of chunk c go to
// accumulator c mod (NACC - 1); the epilogue adds the NACC partial sums in fp32 round-to-nearest.  NACC = 2
// with two TMEM buffers (the epilogue's TMEM reads of a pass overlap the next pass) up to 768 neurons, NACC = 4
// with one buffer above: main chains of at most 4 (Cx + 1) / 3 MMAs.
// The epilogue (thread = frame x 32 neurons) turns accumulators into tanh + noise, splits to fp16 hi/lo and
// writes the granules straight into the next state buffer (generic stores + fence.proxy.async, read back by
// the producer's tensor-map copies once the pass is published).  Only 16 KB of the state (the aug chunk) is
// shared-memory resident, so the rings are three times as deep as in the resident kernel.
//
// The readout y_t = W_out[g(b)] [x_t; u_t] runs on the CUDA cores inside the epilogue: a thread owns one
// frame, so it multiplies ITS frame's readout rows (fp32 table [pair of neurons][output][2], read through L1)
// with the fp32 states it has just produced, two neurons per FFMA2; the eight partial sums per (frame,
// output) go through shared memory to the frame warps, which add the input block, emit y and feed it back.
// Any frame -> readout map is allowed.
#include "tc_common.cuh"

namespace {

constexpr int TCS_THREADS = 672;           // 21 warps (80 registers per thread: the register file is split over four schedulers)
constexpr int TCS_MAXA = 4, TCS_MAXB = 6, TCS_MAXPASS = 16;
constexpr int ATILE = 2 * STILE;           // ring-A item: hi | lo tile of one 64-neuron state chunk (16 KB)
constexpr int BSLOT = 2 * SLOT;            // ring-B item: hi | lo weight tile of one (slab, chunk) (32 KB, ONE bulk copy:
                                           // a bulk-copy request costs its issuing thread ~680 cycles whatever its size,
                                           // profiles/r2_tma_stream_probe.txt)
constexpr int WSTAGE_SLOTS = 3;            // readouts staged per epilogue warp and pass (32 frames of 18-frame blocks: <= 3)

struct TcsParams {
    int B, T, N, n_in, n_out, transient, feedback, su, sy, n_groups;
    int NG, Cx, nacc, nbuf;                // 256-neuron groups (= passes per step), state chunks, accumulators per pass, TMEM buffers
    int nA, nB;                            // ring depths
    float noise_amp;
    unsigned long long seed;
    const float *wo_x, *wo_u;              // readout tables [G][N_pad/2][NOP][2], [G][NOP][24]
    const float *in, *in_scale, *in_shift, *t_scale, *t_shift;
    const int *group_ids;
    const float *x0, *y0, *noise;
    float *ext_out, *y_out;
    const float *teacher;
    unsigned char *state;                  // [CTA][2][Cx][ATILE]
    int steps, row0;
    float acc_gain;                        // 1 + truncation bias of a main accumulate chain (esn_tc_set_acc_k0)
    long long *timeline;                   // [steps][16] SM-clock stamps of CTA 0 (issuer 0-8, epilogue warp 9-14, frame warp 15), or null
};

// fp64 W_out [G][n_out][N + n_in] -> fp32 tables of the CUDA-core readout
__global__ void tcs_prepare_readout_kernel(const double *__restrict__ W_out, int N, int N_pad, int n_in, int n_out,
                                           int NOP, float *__restrict__ wo_x, float *__restrict__ wo_u) {
    const int g = blockIdx.x, P = N + n_in;
    const double *w = W_out + (size_t)g * n_out * P;
    float *ox = wo_x + (size_t)g * N_pad * NOP;
    for (int e = threadIdx.x; e < N_pad * NOP; e += blockDim.x) {
        const int pr = e / (2 * NOP), r = e % (2 * NOP), o = r >> 1, n = 2 * pr + (r & 1);
        ox[e] = (o < n_out && n < N) ? (float)w[(size_t)o * P + n] : 0.f;
    }
    float *ou = wo_u + (size_t)g * NOP * 24;
    for (int e = threadIdx.x; e < NOP * 24; e += blockDim.x) {
        const int o = e / 24, i = e % 24;
        ou[e] = (o < n_out && i < n_in) ? (float)w[(size_t)o * P + N + i] : 0.f;
    }
}

struct TcsEpi {
    uint32_t key;
    float dsc, amp16s, ampoffs, ampf;
};

// One 32-neuron block of an epilogue thread: accumulators of ITS frame for neurons n0 .. n0+31 ->
// [7/6] Pade tanh -> noise -> x 2^SX -> fp16 hi/lo granules stored into the frame's row of the NEXT state
// image in global memory, and the frame's share of the readout (acc[o] += w[o][n] x[n], neuron pairs packed).
template <bool DBG, bool FIX, bool PAD, int NOP, bool RO>
__device__ __forceinline__ void tcs_block32(const TcsParams &p, const uint32_t (&v)[32], int it, int n0, int b, bool live,
                                            unsigned char *grow, int fx, int P, const TcsEpi &es,
                                            const float *__restrict__ wo, uint64_t (&acc)[NOP]) {
    constexpr float XS = (float)(1 << SX);
    const bool use_noise = p.noise_amp != 0.f;
    const uint64_t dsc2 = pk2(es.dsc, es.dsc);
    const uint64_t n0c = pk2(XS, XS), n1 = pk2(378.0f * XS, 378.0f * XS), n2 = pk2(17325.0f * XS, 17325.0f * XS),
                   n3 = pk2(135135.0f * XS, 135135.0f * XS);
    const uint64_t d0 = pk2(28.0f, 28.0f), d1 = pk2(3150.0f, 3150.0f), d2 = pk2(62370.0f, 62370.0f),
                   d3 = pk2(135135.0f, 135135.0f);
    const uint64_t amp2 = pk2(es.amp16s, es.amp16s), off2 = pk2(-es.ampoffs, -es.ampoffs);
    const uint32_t hkey = es.key + (uint32_t)(n0 >> 1) * 0xC2B2AE35U;
    const int g0 = (n0 & 63) >> 3;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {                      // granule of 8 neurons
        uint32_t hi2[4], lo2[4];
        float ex[8];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {                  // pair of neurons
            const int jj = g8 * 8 + pr * 2, n = n0 + jj;
            float4 wv[NOP / 2];
            if (RO) {
                const float4 *w4 = reinterpret_cast<const float4 *>(wo + (size_t)(jj >> 1) * (2 * NOP));
#pragma unroll
                for (int q = 0; q < NOP / 2; ++q) wv[q] = w4[q];        // staged in shared memory (or global: generic load)
            }
            const uint64_t z = mul2(pk2u(v[jj], v[jj + 1]), dsc2);
            const uint64_t z2 = mul2(z, z);
            uint64_t num = fma2(z2, n0c, n1);
            num = fma2(num, z2, n2);
            num = fma2(num, z2, n3);
            uint64_t den = fma2(d0, z2, d1);
            den = fma2(den, z2, d2);
            den = fma2(den, z2, d3);
            float za, zb, da, db;
            un2(z, za, zb);
            un2(den, da, db);
            uint64_t nt = pk2(0.f, 0.f);                  // noise term, already x 2^SX
            if (use_noise) {
                if (DBG && p.noise) {
                    float u[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        u[e] = (live && n + e < p.N) ? p.noise[((size_t)b * p.steps + it) * p.N + n + e] : 0.5f;
                    nt = pk2(fmaf(u[0], es.ampf, -es.ampoffs), fmaf(u[1], es.ampf, -es.ampoffs));
                } else {
                    const uint32_t hb = esn_fold32(hkey + (uint32_t)(jj >> 1) * 0xC2B2AE35U);
                    nt = fma2(pk2((float)(hb & 0xFFFFu), (float)(hb >> 16)), amp2, off2);
                }
            }
            uint64_t xs = fma2(mul2(num, z), pk2(rcp_approx(da), rcp_approx(db)), nt);
            if (FIX || PAD) {
                float xa, xb, na, nb;
                un2(xs, xa, xb);
                un2(nt, na, nb);
                if (FIX && fabsf(za) > 3.0f) xa = fmaf(tanh_large(za), XS, na);
                if (FIX && fabsf(zb) > 3.0f) xb = fmaf(tanh_large(zb), XS, nb);
                if (PAD && n >= p.N) xa = 0.f;
                if (PAD && n + 1 >= p.N) xb = 0.f;
                xs = pk2(xa, xb);
            }
            if (DBG) un2(xs, ex[2 * pr], ex[2 * pr + 1]);
            if (RO) {
#pragma unroll
                for (int q = 0; q < NOP / 2; ++q) {
                    acc[2 * q] = fma2(pk2(wv[q].x, wv[q].y), xs, acc[2 * q]);
                    acc[2 * q + 1] = fma2(pk2(wv[q].z, wv[q].w), xs, acc[2 * q + 1]);
                }
            }
            split_pair(xs, hi2[pr], lo2[pr]);
        }
        if (DBG && p.ext_out && live) {                   // this frame's 8 new states: 32 contiguous bytes of E
            const int n = n0 + g8 * 8;
            float *dst = p.ext_out + ((size_t)b * p.T + it + p.row0) * P + n;
            if (!PAD && (P & 3) == 0) {
                reinterpret_cast<float4 *>(dst)[0] = make_float4(ex[0] * (1.0f / XS), ex[1] * (1.0f / XS), ex[2] * (1.0f / XS), ex[3] * (1.0f / XS));
                reinterpret_cast<float4 *>(dst)[1] = make_float4(ex[4] * (1.0f / XS), ex[5] * (1.0f / XS), ex[6] * (1.0f / XS), ex[7] * (1.0f / XS));
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (n + i < p.N) dst[i] = ex[i] * (1.0f / XS);
            }
        }
        unsigned char *a = grow + ((uint32_t)((g0 + g8) ^ fx) << 4);
        *reinterpret_cast<uint4 *>(a) = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
        *reinterpret_cast<uint4 *>(a + STILE) = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
    }
}

__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Warps (672 threads): 0-1 frame warps (thread = frame: inputs, feedback, readout assembly), 2 weight producer
// (ring B), 3 MMA issuer (CTA 0), 4-19 epilogue (quadrant q = warp & 3: TMEM lanes 32 q ..; cq = (warp - 4) >> 2:
// columns 32 cq .. of every group), 20 state producer (ring A).  One producer thread per ring because a bulk-copy
// request occupies its issuing thread for ~680 cycles on an idle SM whatever its size
// (profiles/r2_tma_stream_probe.txt): one thread cannot feed one chunk (776 cycles of MMAs) with two requests.
// (Measured and dropped: two threads per ring -- 1.67 K instead of 1.49 K cycles per chunk; the readout as a second
// sweep after the state is published -- the aug chunk of the next step then waits for y.  What bounds a chunk is
// shared-memory bandwidth: 48 KB of bulk-copy writes plus 72 KB of operand reads by its twelve MMAs = 940 cycles
// at 128 B/cycle, next to the epilogue's own traffic.)
template <bool DBG, int NOP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TCS_THREADS, 1)
esn_predict_tcs(const TcsParams p, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t fullA[TCS_MAXA], emptyA[TCS_MAXA], fullB[TCS_MAXB], emptyB[TCS_MAXB];
    __shared__ __align__(8) uint64_t bar_pass[2], bar_tfree[2], bar_xready[TCS_MAXPASS], bar_aug, bar_ypart, bar_step;
    __shared__ uint32_t s_tmem;

    const TcGeom gm = tc_geom(p.N, p.n_in);
    const int C = gm.C, Cx = p.Cx, NG = p.NG, NP = p.NG, nacc = p.nacc, nbuf = p.nbuf;
    const int Cr = (p.N + 63) >> 6;        // state chunks that hold real neurons (600 neurons: 10 of the 12 of N_pad = 768)
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    unsigned char *aug = base, *ringA = base + ATILE, *ringB = ringA + (size_t)p.nA * ATILE;
    float *ypart = reinterpret_cast<float *>(ringB + (size_t)p.nB * BSLOT);         // [8][NOP][64]
    unsigned char *wstage = reinterpret_cast<unsigned char *>(ypart) + (size_t)8 * NOP * FT * sizeof(float);   // [16 warps][3][1 KB] (NOP == 8)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int tile0 = (blockIdx.x >> 1) * (2 * FT) + (int)rank * FT;     // first frame owned by this CTA
    const int P = p.N + p.n_in, N_pad = NG * 256;
    const bool harvest = p.teacher != nullptr;
    const int nst = p.steps;
    unsigned char *st_cta = p.state + (size_t)blockIdx.x * 2 * Cx * ATILE;           // this CTA's two state buffers
    const int xrow_cta = (int)(((size_t)blockIdx.x * 2 * Cx * ATILE) / 512);         // the same, in tensor-map rows

    if (tid == 0) {
        for (int i = 0; i < TCS_MAXA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < TCS_MAXB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_pass[i], 1); mbar_init(&bar_tfree[i], 2 * 16); }
        for (int i = 0; i < TCS_MAXPASS; ++i) mbar_init(&bar_xready[i], 16);
        mbar_init(&bar_aug, 2 * 2);
        mbar_init(&bar_ypart, 16);
        mbar_init(&bar_step, 1);                             // every MMA of a time step has completed (harvest mode)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // aug tile starts as zeros; x_{-1} (state buffer 1) = 0 or x0
    for (int i = tid; i < ATILE / 16; i += TCS_THREADS) reinterpret_cast<uint4 *>(aug)[i] = make_uint4(0, 0, 0, 0);
    {
        uint4 *b1 = reinterpret_cast<uint4 *>(st_cta + (size_t)Cx * ATILE);
        for (int i = tid; i < Cx * ATILE / 16; i += TCS_THREADS) b1[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (p.x0) {
        const float xscale = ldexpf(1.0f, SX);
        unsigned char *b1 = st_cta + (size_t)Cx * ATILE;
        for (int i = tid; i < FT * p.N; i += TCS_THREADS) {
            const int f = i / p.N, n = i - f * p.N, b = tile0 + f;
            if (b < p.B) {
                const float xs = p.x0[(size_t)b * p.N + n] * xscale;
                const __half h = __float2half_rn(xs);
                const __half l = __float2half_rn(xs - __half2float(h));
                unsigned char *a = b1 + (size_t)(n >> 6) * ATILE + sw128_off(f, n & 63);
                *reinterpret_cast<__half *>(a) = h;
                *reinterpret_cast<__half *>(a + STILE) = l;
            }
        }
    }
    fence_async_global();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t r_tfree[2] = {mapa_u32(smem_u32(&bar_tfree[0]), 0), mapa_u32(smem_u32(&bar_tfree[1]), 0)};
    const uint32_t r_aug = mapa_u32(smem_u32(&bar_aug), 0);
    auto arrive0 = [&](uint64_t *local, uint32_t remote) {
        if (rank == 0) mbar_arrive(local);
        else mbar_arrive_cluster_relaxed(remote);
    };

    if (warp < 2) {
        // ============ frame warps: thread = own frame; inputs, readout assembly, feedback ============
        const int f = warp * 32 + lane, b = tile0 + f;
        const bool live = b < p.B;
        const uint32_t row = smem_u32(aug) + (f >> 3) * 1024 + (f & 7) * 128;
        const int fx = f & 7, ng = gm.UW >> 3, yg = gm.YO >> 3;
        const float su = ldexpf(1.0f, p.su), sy = ldexpf(1.0f, p.sy);
        const int g = (live && p.group_ids) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;
        const float *wu = harvest ? nullptr : p.wo_u + (size_t)g * NOP * 24;
        int tl_it = -1;                   // step whose feedback is being published (timeline stamps)
        float cur[24], nxt[24];           // u_it, u_{it+1}: scaled inputs (reference units)
#pragma unroll
        for (int j = 0; j < 24; ++j) { cur[j] = 0.f; nxt[j] = 0.f; }
        auto load_row = [&](int r) {
#pragma unroll
            for (int j = 0; j < 24; ++j) {
                float v = 0.f;
                if (j < p.n_in && live && r < p.T) {
                    v = p.in[((size_t)b * p.T + r) * p.n_in + j] * p.in_scale[j] + p.in_shift[j];
                    if (DBG && p.ext_out) p.ext_out[((size_t)b * p.T + r) * P + p.N + j] = v;
                }
                nxt[j] = v;
            }
        };
        auto store8 = [&](int gi, const float *v8, float sc) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pair(pk2(v8[2 * e] * sc, v8[2 * e + 1] * sc), h[e], l[e]);
            const uint32_t a = row + ((uint32_t)(gi ^ fx) << 4);
            sts_v4(a, h[0], h[1], h[2], h[3]);
            sts_v4(a + STILE, l[0], l[1], l[2], l[3]);
        };
        // aug tile of the coming step: columns [0, UW) <- nxt 2^su, columns [YO, YO + NOP) <- y 2^sy
        auto publish = [&](const float *y) {
#pragma unroll
            for (int gi = 0; gi < 3; ++gi)
                if (gi < ng) store8(gi, nxt + gi * 8, su);
            store8(yg, y, sy);
            if (NOP == 16) store8(yg + 1, y + 8, sy);
#pragma unroll
            for (int j = 0; j < 24; ++j) cur[j] = nxt[j];
            fence_async_smem();
            __syncwarp();
            if (lane == 0) arrive0(&bar_aug, r_aug);
            if (p.timeline && blockIdx.x == 0 && tid == 0 && tl_it >= 0) p.timeline[tl_it * 16 + 15] = clock64();
        };
        float y[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) y[o] = 0.f;
        if (harvest) {
            // teacher-forced (ESN.fit, libs/pyESN.py:179-182): step it computes states[it + 1] from input row
            // it + 1 and teacher row it.  ext row 0 = [0, u_0] (state part zeroed by the epilogue warps).
            auto load_teacher = [&](int r) {
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    float v = 0.f;
                    if (o < p.n_out && live && p.feedback)
                        v = p.teacher[((size_t)b * p.T + r) * p.n_out + o] * p.t_scale[o] + p.t_shift[o];
                    y[o] = v;
                }
            };
            load_row(0);
            load_row(1);
            load_teacher(0);
            publish(y);
            for (int it = 0; it < nst; ++it) {
                const bool more = it + 1 < nst;
                if (more) { load_row(it + 2); load_teacher(it + 1); }
                mbar_wait<true>(&bar_step, it & 1);                    // every MMA of step it is done with the aug tile
                if (more) publish(y);
            }
        } else {
            load_row(0);
#pragma unroll
            for (int o = 0; o < 16; ++o)
                y[o] = (p.y0 && live && o < p.n_out && p.feedback) ? p.y0[(size_t)b * p.n_out + o] : 0.f;
            publish(y);
            const float inv_xs = ldexpf(1.0f, -SX);
            for (int it = 0; it < nst; ++it) {
                if (it + 1 < p.T) load_row(it + 1);
                // input block of y_it = W_out_u u_it
#pragma unroll
                for (int o = 0; o < NOP; ++o) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < 24; ++i)
                        if (i < gm.UW) a = fmaf(__ldg(wu + o * 24 + i), cur[i], a);
                    y[o] = a;
                }
                mbar_wait<true>(&bar_ypart, it & 1);
#pragma unroll
                for (int o = 0; o < NOP; ++o) {
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) s += ypart[(i * NOP + o) * FT + f];
                    y[o] = (o < p.n_out) ? fmaf(s, inv_xs, y[o]) : 0.f;
                }
                if (it + 1 < nst) {                                   // feedback first: the issuer will wait for it
                    tl_it = it;
                    float yf[16];
#pragma unroll
                    for (int o = 0; o < 16; ++o) yf[o] = (p.feedback && o < NOP) ? y[o] : 0.f;
                    publish(yf);
                }
                if (it >= p.transient && live) {
                    float *dst = p.y_out + ((size_t)b * (p.T - p.transient) + (it - p.transient)) * p.n_out;
#pragma unroll
                    for (int o = 0; o < NOP; ++o)
                        if (o < p.n_out) dst[o] = (y[o] - p.t_shift[o]) / p.t_scale[o];
                }
            }
        }
    } else if (warp == 2) {
        // ============ weight producers: this CTA's slab of every pass, hi | lo tile of a chunk in ONE 32 KB copy ============
        if (elect_one()) {
            const uint32_t ringB_s = smem_u32(ringB);
            uint32_t r_fullB[TCS_MAXB];
#pragma unroll
            for (int i = 0; i < TCS_MAXB; ++i) r_fullB[i] = mapa_u32(smem_u32(&fullB[i]), 0);
            uint32_t itemB = 0;
            int slotB = 0;
            for (int it = 0; it < nst; ++it)
                for (int ps = 0; ps < NP; ++ps) {
                    const int slab = 2 * ps + (int)rank;
                    for (int c = 0; c <= Cx; ++c) {                  // state chunks, then the aug chunk
                        if (c >= Cr && c < Cx) continue;             // padding chunk: the state there is identically zero
                        mbar_wait<false>(&emptyB[slotB], ((itemB / p.nB) & 1) ^ 1);
                        if (rank == 0) mbar_expect_tx(&fullB[slotB], 2u * BSLOT);
                        tma2_g2s(ringB_s + (uint32_t)slotB * BSLOT, &map_w, 0, ((slab * C + c) * 2) * (SLOT / 512), r_fullB[slotB]);
                        ++itemB;
                        if (++slotB == p.nB) slotB = 0;
                    }
                }
        }
    } else if (warp == 20) {
        // ============ state producers: this CTA's state chunks of x_{it-1}, once per pass ============
        if (elect_one()) {
            const uint32_t ringA_s = smem_u32(ringA);
            uint32_t r_fullA[TCS_MAXA];
#pragma unroll
            for (int i = 0; i < TCS_MAXA; ++i) r_fullA[i] = mapa_u32(smem_u32(&fullA[i]), 0);
            uint32_t itemA = 0;
            int slotA = 0;
            for (int it = 0; it < nst; ++it) {
                const int rb = (it + 1) & 1;                           // buffer that holds x_{it-1}
                for (int ps = 0; ps < NP; ++ps)
                    for (int c = 0; c < Cr; ++c) {
                        // chunk c was written by the epilogue of pass c / 4 of the previous step
                        if (ps == 0 && it > 0 && (c & 3) == 0) mbar_wait<false>(&bar_xready[c >> 2], (it - 1) & 1);
                        mbar_wait<false>(&emptyA[slotA], ((itemA / p.nA) & 1) ^ 1);
                        if (rank == 0) mbar_expect_tx(&fullA[slotA], 2u * ATILE);
                        tma2_g2s(ringA_s + (uint32_t)slotA * ATILE, &map_x, 0, xrow_cta + (rb * Cx + c) * (ATILE / 512), r_fullA[slotA]);
                        ++itemA;
                        if (++slotA == p.nA) slotA = 0;
                    }
            }
        }
    } else if (warp == 3 && rank == 1) {
        // (CTA 1 has no issuer: its tensor core is driven from CTA 0)
    } else if (warp == 3) {
        // ============ MMA issuer (CTA 0): one thread drives both tensor cores ============
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(128, 256);
            const uint32_t ringA0 = desc_lo(smem_u32(ringA)), ringB0 = desc_lo(smem_u32(ringB)), aug0 = desc_lo(smem_u32(aug));
            const uint32_t lod = STILE >> 4;
            const int ku = (gm.UW + 15) / 16, ky = gm.YO / 16;
            uint32_t itemA = 0, itemB = 0;
            int slotA = 0, slotB = 0, s = 0;
            // one state chunk (descriptor low word a_hi) against the pass's slab pair: one ring-B item [hi | lo].  Main
            // products hi*hi -> accumulator dm, corrections lo*hi and hi*lo -> accumulator dc.
            long long *trace = nullptr;                               // per-chunk stamps of one step (profiling aid)
            int tr_i = 0;
            auto chunk_items = [&](uint32_t a_hi, uint32_t dm, uint32_t dc, bool first_main, bool first_corr, bool is_aug) {
                const int ks = is_aug ? ku + 1 : 4;
                mbar_wait<false>(&fullB[slotB], (itemB / p.nB) & 1);
                if (trace && tr_i < 64) trace[tr_i * 4 + 2] = clock64();
                tc_fence_after();
                const uint32_t w = ringB0 + slotB * (BSLOT >> 4), wl = w + (SLOT >> 4);
#pragma unroll 4
                for (int kk = 0; kk < ks; ++kk) {
                    const uint32_t ko = (uint32_t)((is_aug && kk == ku) ? ky : kk) * 2;
                    umma2_f16(dm, a_hi + ko, w + ko, idesc, (first_main && kk == 0) ? 0u : 1u);
                    umma2_f16(dc, a_hi + lod + ko, w + ko, idesc, (first_corr && kk == 0) ? 0u : 1u);
                    umma2_f16(dc, a_hi + ko, wl + ko, idesc, 1u);
                }
                umma2_commit_pair(&emptyB[slotB]);
                if (trace && tr_i < 64) { trace[tr_i * 4 + 3] = clock64(); ++tr_i; }
                ++itemB;
                if (++slotB == p.nB) slotB = 0;
            };
            const int nmain = nacc - 1;
            for (int it = 0; it < nst; ++it) {
                if (p.timeline) p.timeline[it * 16] = clock64();
                trace = (p.timeline && it == 200) ? p.timeline + (size_t)nst * 16 : nullptr;
                tr_i = 0;
                for (int ps = 0; ps < NP; ++ps, ++s) {
                    const int buf = nbuf == 2 ? (s & 1) : 0, use = nbuf == 2 ? (s >> 1) : s;
                    const uint32_t dbase = tmem + buf * nacc * 128, dc = dbase + nmain * 128;
                    mbar_wait_cluster<false>(&bar_tfree[buf], use & 1);          // both CTAs have drained this TMEM buffer
                    tc_fence_after();
                    if (p.timeline && ps < 2) p.timeline[it * 16 + 1 + 4 * ps] = clock64();
                    for (int c = 0; c < Cr; ++c) {                   // (chunks >= Cr hold padding neurons only: skipped)
                        if (trace && tr_i < 64) trace[tr_i * 4] = clock64();
                        mbar_wait<false>(&fullA[slotA], (itemA / p.nA) & 1);
                        if (trace && tr_i < 64) trace[tr_i * 4 + 1] = clock64();
                        tc_fence_after();
                        const uint32_t a = ringA0 + slotA * (ATILE >> 4);
                        chunk_items(a, dbase + (c % nmain) * 128, dc, c < nmain, c == 0, false);
                        umma2_commit_pair(&emptyA[slotA]);
                        ++itemA;
                        if (++slotA == p.nA) slotA = 0;
                    }
                    if (p.timeline && ps < 2) p.timeline[it * 16 + 2 + 4 * ps] = clock64();
                    if (ps == 0) {
                        mbar_wait_cluster<false>(&bar_aug, it & 1);              // [u_it | y_{it-1}] is in place in both CTAs
                        tc_fence_after();
                        if (p.timeline) p.timeline[it * 16 + 3] = clock64();
                    }
                    chunk_items(aug0, dbase + (Cx % nmain) * 128, dc, false, false, true);
                    umma2_commit_pair(&bar_pass[buf]);
                    if (p.timeline && ps < 2) p.timeline[it * 16 + 4 + 4 * ps] = clock64();
                }
                if (harvest) umma2_commit_pair(&bar_step);
            }
        }
    } else if (warp >= 4 && warp < 20) {
        // ============ epilogue: TMEM -> tanh -> noise -> fp16 hi/lo -> next state image; readout shares ============
        const int e = warp - 4, q = warp & 3, cq = e >> 2, hl = q >> 1;
        const int f = 32 * (q & 1) + lane, b = tile0 + f;
        const bool live = b < p.B;
        const int fx = f & 7;
        const uint32_t lane_tm = tmem + ((uint32_t)(q * 32) << 16);
        const int g = (live && p.group_ids) ? min(max(p.group_ids[b], 0), p.n_groups - 1) : 0;
        const float *wo_g = harvest ? nullptr : p.wo_x + (size_t)g * N_pad * NOP;
        const size_t frow = (size_t)(f >> 3) * 1024 + fx * 128;
        // The readout rows this warp needs in a pass -- [32 neurons][NOP outputs] of each distinct readout among its 32
        // frames, 1 KB per readout -- are staged in shared memory by cp.async while the warp waits for the pass's MMAs
        // (from L2 they cost ~600 cycles per dependent batch: 11 K cycles per block instead of 3 K).
        constexpr bool STAGE = NOP == 8;
        int wslot = WSTAGE_SLOTS, nslots = 0, g_slot[WSTAGE_SLOTS] = {0, 0, 0};
        unsigned char *wst = wstage + (size_t)e * WSTAGE_SLOTS * 1024;
        if (STAGE && !harvest) {
            const unsigned mm = __match_any_sync(0xffffffffu, g);
            const int leader = __ffs(mm) - 1;
            const unsigned leaders = __ballot_sync(0xffffffffu, lane == leader);
            wslot = __popc(leaders & ((1u << leader) - 1));
            nslots = min(__popc(leaders), WSTAGE_SLOTS);
#pragma unroll
            for (int sl = 0; sl < WSTAGE_SLOTS; ++sl) {
                const unsigned src_lane = __fns(leaders, 0, sl + 1);
                g_slot[sl] = __shfl_sync(0xffffffffu, g, src_lane & 31);
            }
        }
        TcsEpi es;
        es.dsc = ldexpf(1.0f, -(SX + SW)) * p.acc_gain;
        es.ampf = p.noise_amp * (float)(1 << SX);
        es.amp16s = es.ampf * (1.0f / 65536.0f);
        es.ampoffs = 0.5f * es.ampf;
        __syncwarp();
        if (lane == 0) { arrive0(&bar_tfree[0], r_tfree[0]); arrive0(&bar_tfree[1], r_tfree[1]); }   // the TMEM buffers start free
        if (harvest && DBG && p.ext_out && live) {         // ext row 0: the state before the first step is zero
            float *row = p.ext_out + (size_t)b * p.T * P;
            for (int j = 0; j < NG; ++j)
                for (int i = 0; i < 32; ++i) {
                    const int n = 256 * j + 128 * hl + 32 * cq + i;
                    if (n < p.N) row[n] = 0.f;
                }
        }
        int s = 0;
        for (int it = 0; it < nst; ++it) {
            es.key = esn_noise_key(p.seed, (uint32_t)b, (uint32_t)it);
            unsigned char *nbuf_ptr = st_cta + (size_t)(it & 1) * Cx * ATILE;        // x_it goes to buffer it & 1
            uint64_t acc[NOP];
#pragma unroll
            for (int o = 0; o < NOP; ++o) acc[o] = pk2(0.f, 0.f);
            for (int ps = 0; ps < NP; ++ps, ++s) {
                const int buf = nbuf == 2 ? (s & 1) : 0, use = nbuf == 2 ? (s >> 1) : s;
                const int n0 = 256 * ps + 128 * hl + 32 * cq;
                if (STAGE && !harvest) {
                    __syncwarp();                                      // the previous block's reads of the staging slots are done
#pragma unroll
                    for (int sl = 0; sl < WSTAGE_SLOTS; ++sl)
                        if (sl < nslots) {
                            const unsigned char *src = reinterpret_cast<const unsigned char *>(
                                p.wo_x + (size_t)g_slot[sl] * N_pad * NOP + (size_t)(n0 >> 1) * (2 * NOP));
                            cp_async16(wst + sl * 1024 + lane * 16, src + lane * 16);
                            cp_async16(wst + sl * 1024 + 512 + lane * 16, src + 512 + lane * 16);
                        }
                    cp_async_commit();
                }
                mbar_wait<true>(&bar_pass[buf], use & 1);
                tc_fence_after();
                const bool st4 = p.timeline && blockIdx.x == 0 && warp == 4 && lane == 0 && ps < 2;
                if (st4) p.timeline[it * 16 + 9 + 3 * ps] = clock64();
                uint32_t v[32];
                {
                    // the pass's partial sums (main accumulators, then the corrections), added in fp32 RN
                    const uint32_t col = lane_tm + (uint32_t)(buf * nacc * 128 + 32 * cq);
                    tmem_ld32(col, v);
                    tmem_ld_wait();
                    for (int a = 1; a < nacc; ++a) {
                        uint32_t t[32];
                        tmem_ld32(col + (uint32_t)(a * 128), t);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(t[i]));
                    }
                }
                tc_fence_before();                                     // this warp is done with the TMEM buffer
                __syncwarp();
                if (lane == 0) arrive0(&bar_tfree[buf], r_tfree[buf]);
                if (st4) p.timeline[it * 16 + 10 + 3 * ps] = clock64();
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) m = fmaxf(m, fmaxf(fabsf(__uint_as_float(v[i])), fabsf(__uint_as_float(v[i + 1]))));
                const bool big = __any_sync(0xffffffffu, m * es.dsc > 3.0f);
                unsigned char *grow = nbuf_ptr + (size_t)(n0 >> 6) * ATILE + frow;
                const float *wo = harvest ? nullptr : wo_g + (size_t)(n0 >> 1) * (2 * NOP);
                if (STAGE && !harvest) {
                    cp_async_wait<0>();
                    __syncwarp();
                    if (wslot < WSTAGE_SLOTS) wo = reinterpret_cast<const float *>(wst + wslot * 1024);
                }
                if (harvest) {
                    if (n0 + 32 <= p.N && !big) tcs_block32<DBG, false, false, NOP, false>(p, v, it, n0, b, live, grow, fx, P, es, wo, acc);
                    else tcs_block32<DBG, true, true, NOP, false>(p, v, it, n0, b, live, grow, fx, P, es, wo, acc);
                } else {
                    if (n0 + 32 <= p.N && !big) tcs_block32<DBG, false, false, NOP, true>(p, v, it, n0, b, live, grow, fx, P, es, wo, acc);
                    else tcs_block32<DBG, true, true, NOP, true>(p, v, it, n0, b, live, grow, fx, P, es, wo, acc);
                }
                // publish the chunks of this pass: generic stores -> visible to the producers' tensor-map copies
                fence_async_global();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_xready[ps]);
                if (st4) p.timeline[it * 16 + 11 + 3 * ps] = clock64();
            }
            if (!harvest) {
#pragma unroll
                for (int o = 0; o < NOP; ++o) {
                    float lo, hi;
                    un2(acc[o], lo, hi);
                    ypart[((hl * 4 + cq) * NOP + o) * FT + f] = lo + hi;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_ypart);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                   // nobody leaves while the peer may still touch its SMEM / TMEM
    if (warp == 3) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

struct TcsGeom { int NG, N_pad, Cx, NOP; };
inline TcsGeom tcs_geom(int N, int n_out) {
    TcsGeom g;
    g.NG = (N + 255) / 256;
    g.N_pad = g.NG * 256;
    g.Cx = g.N_pad / 64;
    g.NOP = n_out <= 8 ? 8 : 16;
    return g;
}

}  // namespace

extern "C" int esn_tcs_supported(int N, int n_in, int n_out) {
    return (N > 0 && N <= 4096 && n_in > 0 && n_in <= 24 && n_out > 0 && n_out <= 16) ? 1 : 0;
}

extern "C" long long esn_tcs_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    const TcsGeom g = tcs_geom(N, 8);
    return (long long)(2 * ((B + 2 * FT - 1) / (2 * FT))) * 2 * g.Cx * ATILE;
}

extern "C" long long esn_tcs_readout_floats(int N, int n_out, long long *wo_u_floats_host) {
    const TcsGeom g = tcs_geom(N, n_out);
    if (wo_u_floats_host) *wo_u_floats_host = (long long)g.NOP * 24;
    return (long long)g.N_pad * g.NOP;
}

extern "C" int esn_tcs_prepare_readout(const double *W_out, int N, int n_in, int n_out, int n_groups, float *wo_x,
                                       float *wo_u, void *stream) {
    if (!W_out || !wo_x || !wo_u || n_groups <= 0) return ESN_E_BADARG;
    if (!esn_tcs_supported(N, n_in, n_out)) return ESN_E_UNSUPPORTED;
    const TcsGeom g = tcs_geom(N, n_out);
    tcs_prepare_readout_kernel<<<n_groups, 256, 0, (cudaStream_t)stream>>>(W_out, N, g.N_pad, n_in, n_out, g.NOP, wo_x, wo_u);
    return esn_launch_status();
}

extern "C" int esn_tcs_run(const esn_tcs_args *a, void *stream) {
    if (!a) return ESN_E_BADARG;
    if (a->B <= 0 || a->T <= 0 || a->transient < 0 || a->transient >= a->T) return ESN_E_BADARG;
    if (!esn_tcs_supported(a->N, a->n_in, a->n_out)) return ESN_E_UNSUPPORTED;
    const bool harvest = a->teacher != nullptr;
    if (!a->weights || !a->in || !a->in_scale || !a->in_shift || !a->t_scale || !a->t_shift || !a->workspace) return ESN_E_BADARG;
    if (harvest ? (!a->ext_out || a->T < 2) : (!a->wo_x || !a->wo_u || !a->y_out || a->n_groups <= 0)) return ESN_E_BADARG;
    const TcsGeom g = tcs_geom(a->N, a->n_out);
    const TcGeom gm = tc_geom(a->N, a->n_in);
    if (gm.S != 2 * g.NG) return ESN_E_BADARG;
    TcsParams p;
    p.B = a->B; p.T = a->T; p.N = a->N; p.n_in = a->n_in; p.n_out = a->n_out; p.transient = a->transient;
    p.feedback = a->feedback; p.su = a->su_exp; p.sy = a->sy_exp; p.n_groups = a->n_groups > 0 ? a->n_groups : 1;
    p.NG = g.NG; p.Cx = g.Cx;
    p.nacc = a->accumulators == 2 || a->accumulators == 4 ? a->accumulators : (g.NG <= 3 ? 2 : 4);
    p.nbuf = 4 / p.nacc;
    if (p.NG > TCS_MAXPASS) return ESN_E_TOOLARGE;
    // main products only (the corrections have their own accumulator, 2^-11 smaller): chains of 1 / (nacc - 1) of the k-steps
    p.acc_gain = (float)(1.0 + esn_tc_acc_k0() * ((a->N + 15) / 16 + (gm.UW + 15) / 16 + 1) / (double)(p.nacc - 1));
    p.noise_amp = (float)a->noise_amp; p.seed = a->seed;
    p.wo_x = a->wo_x; p.wo_u = a->wo_u;
    p.in = a->in; p.in_scale = a->in_scale; p.in_shift = a->in_shift; p.t_scale = a->t_scale; p.t_shift = a->t_shift;
    p.group_ids = a->group_ids; p.x0 = a->x0; p.y0 = a->y0; p.noise = a->noise_uniforms;
    p.ext_out = a->ext_out; p.y_out = a->y_out; p.teacher = a->teacher;
    p.state = (unsigned char *)a->workspace;
    p.steps = harvest ? a->T - 1 : a->T;
    p.row0 = harvest ? 1 : 0;
    p.timeline = (long long *)a->timeline;
    // shared memory: aug tile + rings + readout partial sums, as deep as 227 KB allow
    const size_t ypart = (size_t)8 * g.NOP * FT * sizeof(float);
    const size_t wstage = g.NOP == 8 ? (size_t)16 * WSTAGE_SLOTS * 1024 : 0;
    const size_t budget = 227 * 1024 - 1024 - 1024 - ATILE - ypart - wstage;
    p.nA = a->ring_a > 0 ? std::min(a->ring_a, TCS_MAXA) : 3;
    int nb = (int)((budget - (size_t)p.nA * ATILE) / BSLOT);
    nb = std::min(nb, TCS_MAXB);
    if (a->ring_b > 0) nb = std::min(nb, a->ring_b);
    if (nb < 2) return ESN_E_TOOLARGE;
    p.nB = nb;
    const size_t smem = 1024 + ATILE + (size_t)p.nA * ATILE + (size_t)p.nB * BSLOT + ypart + wstage;
    const int grid = 2 * ((a->B + 2 * FT - 1) / (2 * FT));
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ESN_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return ESN_E_UNSUPPORTED;
        encode = (encode_fn)fn;
    }
    auto make_map = [&](CUtensorMap *m, const void *base, size_t bytes, unsigned box_bytes) -> bool {
        const cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 512)};
        const cuuint64_t strides[1] = {512};
        const cuuint32_t box[2] = {256, box_bytes / 512}, estr[2] = {1, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    CUtensorMap map_w, map_x;
    if (!make_map(&map_w, a->weights, gm.weight_bytes, BSLOT)) return ESN_E_BADARG;
    if (!make_map(&map_x, a->workspace, (size_t)grid * 2 * g.Cx * ATILE, ATILE)) return ESN_E_BADARG;
    const bool dbg = a->noise_uniforms || a->ext_out;
    auto launch = [&](auto kern) -> int {
        ESN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, TCS_THREADS, smem, (cudaStream_t)stream>>>(p, map_w, map_x);
        return 0;
    };
    int rc;
    if (g.NOP == 8) rc = dbg ? launch(esn_predict_tcs<true, 8>) : launch(esn_predict_tcs<false, 8>);
    else rc = dbg ? launch(esn_predict_tcs<true, 16>) : launch(esn_predict_tcs<false, 16>);
    if (rc) return rc;
    return esn_launch_status();
}
