// Shared pieces of the tensor-core (tcgen05 + TMEM) recurrence kernels: tile geometry of the pre-swizzled
// fp16 hi/lo images, PTX wrappers (mbarrier, tcgen05, bulk tensor copies) and the packed fp32x2 helpers.
// Included by recurrence_tc.cu (state resident in shared memory, N <= 512) and recurrence_tcs.cu (state
// streamed through L2, any N, per-frame readouts).
#pragma once
#include <algorithm>
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through the runtime)
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int FT = 64;                 // frames per CTA (UMMA N)
constexpr int NST = 4;                 // ring slots
constexpr int SLOT = 16384;            // bytes per ring slot: [128 rows x 64 k] fp16
constexpr int STILE = 8192;            // bytes per state tile: [64 rows x 64 k] fp16
constexpr int YTILE = 2048;            // bytes per readout tile: [16 rows x 64 k] fp16
constexpr int SX = 8, SW = 8;          // power-of-two pre-scales of state and weights
constexpr int TMEM_COLS = 512;
#define ESN_TC_ACC_K0_DEFAULT 1.45e-8  // relative truncation loss per MMA of an accumulate chain (esn_tc_set_acc_k0; profiles/r2_tc_acc_bias.txt)

struct TcGeom {
    int S, C, UW, YO, ca, kaug;        // slabs, 64-wide K chunks, input block width, y column offset, aug chunk, k-steps in aug chunk
    size_t weight_bytes, readout_tile_bytes, readout_bytes;   // readout image = UMMA tiles + fp32 input-block table [16][24]
};

__host__ __device__ inline TcGeom tc_geom(int N, int n_in) {
    TcGeom g;
    g.S = ((N + 255) / 256) * 2;       // 128-neuron slabs, an even number: each CTA of the pair streams every other one
    g.UW = (n_in + 7) / 8 * 8;
    g.YO = (2 * g.UW + 15) / 16 * 16;
    g.C = 2 * g.S + 1;
    g.ca = 2 * g.S;
    g.kaug = g.YO / 16 + 1;
    g.weight_bytes = (size_t)g.S * g.C * 2 * SLOT;
    g.readout_tile_bytes = (size_t)g.C * 2 * YTILE;
    g.readout_bytes = g.readout_tile_bytes + 16 * 24 * sizeof(float);
    return g;
}

// byte offset of element (row r, column k) inside a SWIZZLE_128B K-major tile of fp16
__host__ __device__ inline int sw128_off(int r, int k) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) & 7) ^ (r & 7)) << 4) + (k & 7) * 2;
}

__device__ inline void store_split(unsigned char *tile_hi, unsigned char *tile_lo, int off, double v) {
    const __half hi = __double2half(v);
    const __half lo = __double2half(v - (double)__half2float(hi));
    *reinterpret_cast<__half *>(tile_hi + off) = hi;
    *reinterpret_cast<__half *>(tile_lo + off) = lo;
}

// ---------------------------------------------------------- PTX helpers -----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spin with a watchdog: a protocol bug traps instead of hanging the GPU.  SLEEP backs
// off between polls so that the many waiting warps do not steal issue slots from the
// single MMA-issuing thread and the producer.
template <bool SLEEP>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done) {
            if (SLEEP) __nanosleep(64);
            if (spin > (1u << 24)) __trap();
        }
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor (K-major, SWIZZLE_128B, 8-row groups 1024 B apart,
// descriptor version 1 of sm_100).  Only the low word depends on the address, so the
// issuing thread keeps low words in registers and advances them by adds.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | (1u << 16); }
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    // D = F32, A = B = F16, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void sts_u16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// x (already pre-scaled) -> fp16 hi at `addr`, fp16 lo at `addr + lo_delta`
__device__ __forceinline__ void split_sts(uint32_t addr, uint32_t lo_delta, float xs) {
    const __half h = __float2half_rn(xs);
    const __half l = __float2half_rn(xs - __half2float(h));
    sts_u16(addr, __half_as_ushort(h));
    sts_u16(addr + lo_delta, __half_as_ushort(l));
}


__device__ __forceinline__ float tanh_large(float z) {      // |z| > 3
    const float e = __expf(2.0f * fabsf(z));
    return copysignf(1.0f - __fdividef(2.0f, e + 1.0f), z);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Signal-only remote arrive (no data published by this thread): relaxed, so no
// MEMBAR.ALL.GPU is emitted (release.cluster costs one per arrive).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Wait on a barrier that collects arrivals from both CTAs (cluster-scope acquire).  Measured:
// polling with acquire.cluster is ~1.5 K cycles per step cheaper than a CTA-scoped poll followed
// by one fence.acq_rel.cluster.
template <bool SLEEP>
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done) {
            if (SLEEP) __nanosleep(64);
            if (spin > (1u << 24)) __trap();
        }
    }
}
// The same wait without acquire semantics.  ptxas follows every successful acquire.cluster wait by CCTL.IVALL, which
// invalidates the SM's whole L1D: five times per time step in the issuer's SM, which a kernel whose other warps
// read tables through L1 (the CUDA-core readout of esn_recur_tcr) cannot afford.  The issuing thread reads nothing
// the arrivals publish: the arriving warps have fenced their shared-memory writes towards the async proxy
// (fence.proxy.async) before they arrive, and the MMAs issued after the wait read them through that proxy, behind
// tcgen05.fence::after_thread_sync.
template <bool SLEEP>
__device__ __forceinline__ void mbar_wait_cluster_relaxed(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done) {
            if (SLEEP) __nanosleep(64);
            if (spin > (1u << 24)) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint64_t *bar) {     // arrives on `bar` in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {   // -> {lo half, hi half}
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two lanes per instruction ----
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void un2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 2-D tiled bulk tensor copy global -> this CTA's shared memory whose completion (complete_tx) is
// signalled on an mbarrier of EITHER CTA of the pair (cta_group::2): both halves of a ring item report
// to the issuer's barrier in CTA 0, no relay hop.
__device__ __forceinline__ void tma2_g2s(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t cluster_mbar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(cluster_mbar) : "memory");
}

// (a, b) pre-scaled -> packed fp16 hi pair and lo pair
__device__ __forceinline__ void split_pair(uint64_t xs, uint32_t &h, uint32_t &l) {
    float xa, xb, la, lb;
    un2(xs, xa, xb);
    h = pack_h2(xa, xb);
    const float2 back = __half22float2(*reinterpret_cast<const __half2 *>(&h));
    un2(sub2(xs, pk2(back.x, back.y)), la, lb);
    l = pack_h2(la, lb);
}

}  // namespace
