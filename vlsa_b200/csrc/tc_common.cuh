// tcgen05 / TMEM / UMMA-descriptor helpers (sm_100a).  Raw PTX; layouts follow the PTX ISA canonical
// shared-memory layouts for tcgen05.mma (128-byte swizzle) — see DESIGN.md §4 (tensor-core variant).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdio.h>

#include "common.cuh"

namespace vlsa {

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on an mbarrier once every tcgen05.mma previously issued by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], 16-bit inputs (formats in the instruction descriptor), fp32 accumulate; ONE thread issues
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand resident in TMEM (lane = row of A, two 16-bit K elements per 32-bit column)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    tc_mma_ss(tmem_d, adesc, bdesc, idesc, accumulate);
}

// one lane of the (converged) warp: the compiler treats an elect.sync branch as uniform, so tcgen05.mma / commit
// inside it compile to straight-line UTCHMMA (a `lane == 0` test makes it emit an ELECT loop around every MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// instruction descriptor, kind::f16 with f32 accumulator
//   [4,6) c_format=1 (f32) | [7,10) a_format | [10,13) b_format (0 = f16, 1 = bf16) | 15 a_major (1 = MN) | 16 b_major |
//   [17,23) N>>3 | [24,29) M>>4
#define UMMA_F16 0u
#define UMMA_BF16 1u
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t a_fmt, uint32_t b_fmt, int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
           (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return umma_idesc(UMMA_BF16, UMMA_BF16, M, N, a_mn_major, b_mn_major);
}
// shared-memory matrix descriptor, 128-byte swizzle: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
// the same descriptor with the start address advanced by `bytes` (no carry out of the 14-bit field: smem < 256 KB)
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }
// byte offset of element (row, 16-byte chunk c, byte b) inside a [rows x 128 B] tile with 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16, uint32_t byte_in_chunk) {
    return row * 128u + (((chunk16 ^ (row & 7u)) & 7u) << 4) + byte_in_chunk;
}

// TMEM -> registers: thread i of warp w reads lane 32*(w%4)+i, N consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// streaming 128-bit global load: read-only path, no L1 allocation, L2 evict-first (X is read exactly once)
__device__ __forceinline__ float4 ldg_stream_f4(const float* p, uint64_t policy) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}
// asynchronous bulk prefetch of a contiguous span into L2 (bytes % 16 == 0)
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// mbarrier wait with a watchdog: a protocol bug traps instead of hanging the GPU.  try_wait carries a
// suspend-time hint, so a waiting warp sleeps in hardware instead of burning issue slots in a spin loop.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// Fast path first: one try_wait (it sleeps in hardware up to its suspend hint) decides almost every wait.  The watchdog
// loop stays INLINE: moving it into a __noinline__ function made the backward of the TMA-fed tcgen05 kernel of round 2 fault intermittently
// (illegal address, only without compute-sanitizer) — a call from single-elected-lane / tcgen05 code is not worth it.
#ifdef VLSA_WD_DEBUG
// Development build: a wait that times out (~50 ms) records (line, block, thread, barrier address | parity) and RETURNS, so the
// kernel ends (with garbage) and the host can read who was stuck where (vlsa_debug_read_wd)
__device__ unsigned int g_wd[4 + 4 * 1000];
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int line) {
    if (mbar_try_wait_hint(bar, parity, 4000u)) return;
    uint32_t spins = 0;
    do {
        if (++spins > (1u << 14)) {
            const unsigned am = __activemask();
            const unsigned slot = (threadIdx.x & 31u) == unsigned(__ffs(am) - 1) ? atomicAdd(&g_wd[0], 1u) : 1000u;
            if (slot < 1000u) {
                g_wd[4 + 4 * slot] = unsigned(line); g_wd[5 + 4 * slot] = blockIdx.x; g_wd[6 + 4 * slot] = threadIdx.x;
                g_wd[7 + 4 * slot] = (smem_u32(bar) << 1) | parity;
            }
            return;
        }
    } while (!mbar_try_wait_hint(bar, parity, 4000u));
}
#define mbar_wait_wd(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
#else
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_hint(bar, parity, 4000u)) return;
    uint32_t spins = 0;
    do {
        // brkpt, not trap / printf: an exit edge or a call inside the kernel makes ptxas ignore setmaxnreg when it allocates
        // registers (every role is then held to the launch bound and the producers of agg_tc_kernel spill)
        if (++spins > (1u << 22)) { asm volatile("brkpt;"); spins = 0; }
    } while (!mbar_try_wait_hint(bar, parity, 4000u));
}
#endif

// x = hi + lo with hi, lo bf16 (16 significant bits in total); packs (a, b) -> bf16x2 with a in the low half
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// x = t0 + t1 + t2 with bf16 terms (24 significant bits); packs (a, b) -> bf16x2 with a in the low half
__device__ __forceinline__ void split_bf16x3(float a, float b, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t0) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(t0 << 16), rb = b - __uint_as_float(t0 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(rb), "f"(ra));
    const float sa = ra - __uint_as_float(t1 << 16), sb = rb - __uint_as_float(t1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(sb), "f"(sa));
}

// x = hi + lo with hi, lo fp16 (packed pairs, first element in the low half); SASS: 2 F2FP + 2 HADD2.F32 + 1 FADD2
__device__ __forceinline__ void split_f16x2(float2 a, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __float22half2_rn(a);
    const float2 hf = __half22float2(h);
    const __half2 l = __float22half2_rn(__fadd2_rn(a, make_float2(-hf.x, -hf.y)));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    split_f16x2(make_float2(a, b), hi, lo);
}

// barrier among `nthreads` threads that also ORs a predicate across them (bar.red.or)
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred q, r;\n\tsetp.ne.u32 q, %3, 0;\n\t"
        "bar.red.or.pred r, %1, %2, q;\n\tselp.u32 %0, 1, 0, r;\n\t}"
        : "=r"(out) : "r"(id), "r"(nthreads), "r"(uint32_t(pred)) : "memory");
    return out != 0;
}

// 2-D tiled TMA load global -> shared (SASS: UTMALDG), completion on an mbarrier, L2 evict-first
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Development build (-DVLSA_TMA_PROF): block 0 of a tcgen05 kernel accumulates, per role, the cycles spent in every kind
// of wait (read back with vlsa_debug_read_prof, scripts/dev_tma_prof.py).  Compiled out of the shipped library.
#ifdef VLSA_TMA_PROF
__device__ long long g_tma_prof[32];
#define PROF_DECL long long prof_t0 = 0, prof_acc[6] = {0, 0, 0, 0, 0, 0}; const long long prof_start = clock64(); (void)prof_t0;
#define PROF_BEGIN() prof_t0 = clock64()
#define PROF_END(k) prof_acc[k] += clock64() - prof_t0
#define PROF_FLUSH(base, n, who)                                                                 \
    if (blockIdx.x == 0 && (who)) {                                                              \
        for (int k_ = 0; k_ < (n); ++k_) g_tma_prof[(base) + k_] = prof_acc[k_];                 \
        g_tma_prof[(base) + (n)] = clock64() - prof_start;                                       \
    }
#else
#define PROF_DECL
#define PROF_BEGIN()
#define PROF_END(k)
#define PROF_FLUSH(base, n, who)
#endif

}  // namespace vlsa

// ---------------------------------------------------------------- thread-block cluster / DSMEM --
namespace vlsa {
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t dsmem_map(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void dsmem_st_v4(uint32_t caddr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(caddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster_wd(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) { printf("vlsa: cluster mbarrier watchdog (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
// exp(x) for x <= ~0 with fp32-grade accuracy: 2^(x*log2e) with a two-term product (x can reach -200)
__device__ __forceinline__ float exp_accurate_fast(float x) {
    const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.92596299112661746e-8f;
    const float t = x * L2E_HI;
    const float e = fmaf(x, L2E_HI, -t) + x * L2E_LO;     // rounding error of t + low part
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return fmaf(r, e * 0.693147180559945f, r);             // 2^(t+e) ~= 2^t (1 + e ln2)
}
}  // namespace vlsa
