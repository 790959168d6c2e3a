// Fused language-guided aggregation on the 5th-generation tensor cores (tcgen05 + TMEM), forward and backward.
// Same contract as agg_simt_kernel<P,BWD,float>: ONE read of X, per-chunk partials — forward: online-softmax
// (m, l, O[P,D]); backward: partial dQn[P,D] — with both skinny contractions on tcgen05.mma (kind::f16, fp32
// accumulators in TMEM):
//
//   GEMM1  S^T[128, 64]  = Qn'[128, 512] (A, resident in TMEM) . [X_hi ; X_lo][64, 512]^T (B, K-major smem)
//   GEMM2  O^T[512 d, 32 | 16] += X_hi^T | X_lo^T [512, 32] (A, MN-major, the SAME smem bytes) . W[32 | 16, 32]^T
//
// Precision design (the tensor core truncates its fp32 accumulator toward zero after every instruction —
// measured with scripts/dev_tc_unit.cu — so the number of accumulation steps per accumulator is kept small):
//   * every row of X is split into fp16 (hi, lo) planes: 22 significant bits at 4 bytes of shared memory per
//     element.  Rows whose norm surely lies in [2, 2^14] (every CONCH-like row) are split as they are; any other row
//     is first scaled by a power of two (largest |x| -> [1,2)) that the softmax warps undo exactly;
//     Qn is split the same way;
//   * A-operand row (TMEM lane) 32 w + j holds prototype p = 4 w + (j & 3), part (j >> 2) & 1 (hi / lo) of Qn
//     restricted to the feature range 128 (j >> 3) .. +127 (zeros elsewhere): every accumulator only sees
//     8 non-zero steps, and the 16 partial sums of a score (2 parts x 4 ranges x 2 planes) are added in fp32
//     registers.  TMEM quadrant q therefore owns prototypes 4 q .. 4 q + 3 completely;
//   * the per-row weights go to the tensor core as two fp16 terms scaled 1 and 2^11 (22 bits); products with
//     the hi and the lo plane of X accumulate in separate TMEM columns; tcgen05.mma needs A and B in the same
//     16-bit format (fp16 x bf16 traps), hence fp16 weights with the lazy-rescale range control below.
//
// Warp roles (20 warps, 1 persistent CTA / SM, static round-robin over chunks):
//   warps 0-7   weights : warp w reads TMEM quadrant w & 3 (prototypes) for tile rows 16 (w >> 2) .. +15:
//               scores -> online softmax with lazy rescale (forward) | A (u - delta) (backward) -> fp16 weight
//               terms to smem; drain of the accumulators per chunk
//   warp  8     GEMM1 issuer (one thread) + TMEM allocation
//   warp  9     GEMM2 issuer (one thread)
//   warp  10    L2 bulk prefetch, PF tiles ahead of the producers (warp 11 idles)
//   warps 12-19 producers: LDG.128 of whole rows (one tile in flight in registers)
//               -> row norm (packed f32x2 FMAs) -> fp16 split -> swizzled STS; backward: also u = dv . x / P
// Ring: 3 tile buffers x 64 KB, a tile = 32 rows = 8 slots of 64 columns x (hi 4 KB | lo 4 KB), 128-byte swizzle.
#pragma once
#include <cuda_fp16.h>
#include <stdio.h>

#include "agg_simt.cuh"
#include "tc_common.cuh"

namespace vlsa {

struct TcCfg {
    static constexpr int D = VLSA_D;
    static constexpr int NP = 16;                 // prototypes padded to 16
    static constexpr int TR = 32;                 // rows per tile (GEMM1 N = 2 TR, GEMM2 K = TR)
    static constexpr int KC = 64;                 // feature columns per slot (128 B of fp16)
    static constexpr int NSLOT = D / KC;          // 8 slots per tile
    static constexpr int PLANE = TR * 128;        // 4 KB
    static constexpr int SLOT = 2 * PLANE;        // hi | lo
    static constexpr int TILE = NSLOT * SLOT;     // 64 KB
    static constexpr int NBUF = 3;
    static constexpr int WROWS = 2 * NP;          // weight operand rows: term 0 | term 1
    static constexpr int WBUF = WROWS * 128;      // 4 KB per weight buffer (rows of 128 B, 64 B used)
    static constexpr int OFF_W = NBUF * TILE;
    static constexpr int OFF_F = OFF_W + 2 * WBUF;
    // floats: rowinfo[NBUF][TR][4] | alpha[16] | cand[2][16] | lsum[16]
    static constexpr int NFLOAT = NBUF * TR * 4 + 16 + 32 + 16;
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 2 * NBUF + 8;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int NSOFT = 8;               // weight ("softmax") warps
#ifndef VLSA_TC_NPROD
#define VLSA_TC_NPROD 8
#endif
    static constexpr int NPROD = VLSA_TC_NPROD;   // producer warps (8 or 16)
    static constexpr int RPW = TR / NPROD;        // tile rows (= register sets) per producer warp
    // register budget per warp (setmaxnreg; the pool is what the launch bound grants the CTA):
    //   8 producers : launch 96 -> weights 96, issuers 48, producers 120   (8 x 96 + 4 x 48 + 8 x 120 = 20 x 96)
    //   16 producers: launch 72 -> weights 96, issuers 48, producers 64    (8 x 96 + 4 x 48 + 16 x 64 <= 28 x 72)
    static constexpr int REG_PROD = NPROD == 8 ? 120 : 64;
    static constexpr int REG_SOFT = 96;
    // rows of the next tile whose loads are issued while the current tile is still being converted; the others are
    // issued after the proxy fence (which waits for every load in flight): tile period = load latency +
    // max(EARLY, RPW - EARLY) row conversions
#ifndef VLSA_TC_EARLY
#define VLSA_TC_EARLY (VLSA_TC_NPROD == 8 ? 3 : 1)
#endif
    static constexpr int EARLY = VLSA_TC_EARLY;
    static constexpr int W_G1 = NSOFT;            // GEMM1 issuer warp
    static constexpr int W_G2 = NSOFT + 1;        // GEMM2 issuer warp
    static constexpr int W_PF = NSOFT + 2;        // L2 prefetch warp (warp NSOFT + 3 idles: warps are allocated in fours)
    static constexpr int W_PROD = NSOFT + 4;      // first producer warp
    static constexpr int NWARPS = NSOFT + 4 + NPROD;
    static constexpr int THREADS = NWARPS * 32;
#ifndef VLSA_TC_PF
#define VLSA_TC_PF 0
#endif
    // L2 prefetch distance in tiles; 0 = off (default).  Measured (scripts/dev_readbw.cu): the register path alone
    // streams at 7.3 TB/s, an L2 prefetch ahead of it only costs bandwidth (6.1 TB/s at 2 tiles, 4.8 at 8).
    static constexpr int PF = VLSA_TC_PF;
    static constexpr int QPITCH = D + 1;          // prologue staging of Qn (aliases the ring)
    // TMEM columns: Qn operand | O^T accumulators: 4 blocks of 128 d x (hi.t0 16 | hi.t1 16 | lo.t0 16) | scores
    static constexpr int TM_Q = 0;
    static constexpr int D2W = 3 * NP;            // 48 columns per 128-feature block
    static constexpr int TM_D2 = 256;
    static constexpr int TM_D1 = TM_D2 + 4 * D2W;   // 448: scores [hi plane rows 0..31 | lo plane rows 0..31]
    static constexpr int TMEM_COLS = 512;
    // forward, online softmax with lazy rescaling: on (re)set the reference is the running maximum + HEADROOM;
    // the TMEM accumulators are rescaled only when a tile maximum exceeds the reference by more than MARGIN, so
    // the fp16 weight terms stay within (0, e^MARGIN] (e^10 = 22026 < 65504) with an absolute floor of 2^-35.
    static constexpr float HEADROOM = 6.f;
    static constexpr float MARGIN = 10.f;
    // backward, the same idea on a power-of-two normaliser H_p: weights / H_p are kept <= 2^BWD_MAXE, (re)set
    // so that the largest weight of the triggering tile becomes 2^BWD_SETE
    static constexpr int BWD_MAXE = 14;
    static constexpr int BWD_SETE = 6;
    // rows with |x|^2 in [FAST_LO, FAST_HI] are split unscaled
    static constexpr float FAST_LO = 4.f;
    static constexpr float FAST_HI = 268435456.f;   // 2^28
};

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

template <bool BWD>
__global__ void __launch_bounds__(TcCfg::THREADS, 1) agg_tc_kernel(const AggParams prm, const int P) {
    using C = TcCfg;
    constexpr int D = C::D, NP = C::NP, TR = C::TR;
    if (int(blockIdx.x) >= prm.total_chunks) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* ring = sm;
    unsigned char* wt = sm + C::OFF_W;
    float* s_rowinfo = reinterpret_cast<float*>(sm + C::OFF_F);      // [NBUF][TR] x (score factor, 2^e, u, -)
    float* s_alpha = s_rowinfo + C::NBUF * TR * 4;                   // [16] rescale factors (rare path)
    float* s_cand = s_alpha + 16;                                    // [2][16] per-half reference candidates
    float* s_lsum = s_cand + 32;                                     // [16] softmax sums of the upper half
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* full = bars;                       // [NBUF] producers (8 warps)      -> GEMM1, weight warps (row info)
    uint64_t* empty = bars + C::NBUF;            // [NBUF] GEMM2 commit             -> producers
    uint64_t* s_ready = bars + 2 * C::NBUF;      //        GEMM1 commit             -> weight warps
    uint64_t* s_free = s_ready + 1;              //        weight warps (8)         -> GEMM1
    uint64_t* w_ready = s_ready + 2;             // [2]    weight warps (8)         -> GEMM2
    uint64_t* w_free = s_ready + 4;              // [2]    GEMM2 commit             -> weight warps
    uint64_t* d2_done = s_ready + 6;             //        last GEMM2 of a chunk    -> weight warps (drain)
    uint64_t* d2_free = s_ready + 7;             //        weight warps (8)         -> GEMM2 of the next chunk
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    uint32_t* s_prog = tmem_ptr + 1;             // tiles filled so far (paces the L2 prefetcher)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < C::NBUF; ++s) { mbar_init(full + s, C::NPROD); mbar_init(empty + s, 1); }
        mbar_init(s_ready, 1); mbar_init(s_free, C::NSOFT);
        for (int s = 0; s < 2; ++s) { mbar_init(w_ready + s, C::NSOFT); mbar_init(w_free + s, 1); }
        mbar_init(d2_done, 1); mbar_init(d2_free, C::NSOFT);
        *s_prog = 0u;
        mbar_fence_init();
    }
    if (warp == C::W_G1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    // ---- prologue: Qn = Q / max(|Q|, eps) staged as fp32 in the (still unused) ring, rows >= P are zero
    {
        float* qn = reinterpret_cast<float*>(ring);
        for (int p = warp; p < NP; p += C::NWARPS) {
            float inv = 0.f;
            if (p < P) {
                float ss = 0.f;
                for (int d = lane; d < D; d += 32) { const float v = __ldg(prm.Q + size_t(p) * D + d); ss += v * v; }
                ss = warp_sum(ss);
                inv = prm.q_prenorm ? 1.f : 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
            }
            for (int d = lane; d < D; d += 32) qn[p * C::QPITCH + d] = p < P ? __ldg(prm.Q + size_t(p) * D + d) * inv : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (warp < 4) {
        // TMEM lane 32 warp + lane: prototype 4 warp + (lane & 3), part (lane >> 2) & 1, feature range lane >> 3
        const float* qrow = reinterpret_cast<const float*>(ring) + (4 * warp + (lane & 3)) * C::QPITCH;
        const bool lo_part = (lane >> 2) & 1;
        const int range = lane >> 3;
        const uint32_t tq = tmem + (uint32_t(32 * warp) << 16) + C::TM_Q;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {                   // 32 columns = 64 features per batch
            uint32_t v[32];
            const bool mine = (cb >> 1) == range;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                uint32_t hi, lo;
                split_f16x2(qrow[cb * 64 + 2 * i], qrow[cb * 64 + 2 * i + 1], hi, lo);
                v[i] = mine ? (lo_part ? lo : hi) : 0u;
            }
            tmem_st32(tq + 32 * cb, v);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();                                       // Qn is in TMEM; the ring may be overwritten from here on
    tc_fence_after();

    // register budget: the launch bound gives every warp 96; the issuer warpgroup (8-11) hands most of its
    // registers back and the two producer warpgroups (12-19) grow to 120 (the pool is per CTA: 8 x 96 + 4 x 48 + 8 x 120 = 20 x 96)
    if (warp >= C::W_PROD) {
        // =========================================================================== producers
        if (C::NPROD == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_PROD));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_PROD));
        const int pw = warp - C::W_PROD;
        const uint64_t policy = make_evict_first_policy();
        // cursor: tile = (pointer to its first row, already offset by this lane's 4 columns; rows left in the chunk)
        int cc = blockIdx.x, bag = 0, rows_left = 0;
        const float* tptr = nullptr;
        bool valid = cc < prm.total_chunks;
        if (valid) {
            long long r0, r1;
            chunk_info(prm, cc, bag, r0, r1);
            tptr = reinterpret_cast<const float*>(prm.X) + r0 * D + 4 * lane;
            rows_left = int(r1 - r0);
        }
        // four register sets of one row each: set j = tile row 4 pw + j.  Rows past the end of the chunk re-read its
        // last row (their weights are masked by the weight warps), so no load is predicated.
        float4 buf[C::RPW][4];
        auto issue_row = [&](int j, const float* tp, int rl) {
            const int lr = min(C::RPW * pw + j, rl - 1);
            const float* src = tp + lr * D;
#pragma unroll
#ifdef VLSA_TC_PLAINLDG
            for (int i = 0; i < 4; ++i) buf[j][i] = __ldg(reinterpret_cast<const float4*>(src + 128 * i));
#else
            for (int i = 0; i < 4; ++i) buf[j][i] = ldg_stream_f4(src + 128 * i, policy);
#endif
        };
        if (valid) {
#pragma unroll
            for (int j = 0; j < C::RPW; ++j) issue_row(j, tptr, rows_left);
        }
        // swizzled byte offset of this lane's 8-byte store inside a slot plane, per row j of the warp:
        // columns 128 i + 4 lane .. +3 -> slot 2 i + (lane >> 4), 16-byte chunk (lane & 15) >> 1, half (lane & 1)
        uint32_t soff[C::RPW];
#pragma unroll
        for (int j = 0; j < C::RPW; ++j)
            soff[j] = (lane >> 4) * C::SLOT + sw128_offset(C::RPW * pw + j, (lane & 15) >> 1, (lane & 1) * 8);
        // backward: dv / P of the bag whose tile sits in the registers (16 columns per lane, as buf)
        float4 dvr[BWD ? 4 : 1];
        int dv_bag = -1;
        uint32_t tt = 0;
        while (valid) {
            const int cur_bag = bag;
            // the tile after this one
            bool nvalid = true;
            const float* nptr = tptr + TR * D;
            int nrows = rows_left - TR;
            if (nrows <= 0) {
                cc += gridDim.x;
                nvalid = cc < prm.total_chunks;
                if (nvalid) {
                    long long r0, r1;
                    chunk_info(prm, cc, bag, r0, r1);
                    nptr = reinterpret_cast<const float*>(prm.X) + r0 * D + 4 * lane;
                    nrows = int(r1 - r0);
                }
            }
            if (BWD && cur_bag != dv_bag) {
                dv_bag = cur_bag;
                const float invP = 1.f / float(P);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t4 = __ldg(reinterpret_cast<const float4*>(prm.dv + size_t(cur_bag) * D + 128 * i + 4 * lane));
                    dvr[i] = make_float4(t4.x * invP, t4.y * invP, t4.z * invP, t4.w * invP);
                }
            }
            const uint32_t b = tt % C::NBUF, u = tt / C::NBUF;
            unsigned char* tile = ring + b * C::TILE;
            // one row: per-lane partial norm (and u = dv . x / P) on the raw values; a warp vote on the partials decides
            // whether the row is split AS IT IS (sufficient for FAST_LO <= |x|^2 <= FAST_HI: every CONCH-like row) or
            // first scaled by a power of two; the five-level shuffle reduction of the norm is interleaved by hand
            // with the quarters of the conversion (ptxas keeps the order), then lane 0 writes the row info.
            auto convert_part = [&](int j, int i) {
#ifdef VLSA_TC_NOCONV
                return;
#endif
                unsigned char* dst = tile + soff[j];
                const float4 x4 = buf[j][i];
                uint32_t h0, l0, h1, l1;
                split_f16x2(make_float2(x4.x, x4.y), h0, l0);
                split_f16x2(make_float2(x4.z, x4.w), h1, l1);
                *reinterpret_cast<uint2*>(dst + 2 * i * C::SLOT) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(dst + 2 * i * C::SLOT + C::PLANE) = make_uint2(l0, l1);
            };
            auto process_row = [&](int j) {
                float2 a2 = make_float2(0.f, 0.f), u2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 x4 = buf[j][i];
                    const float2 xy = make_float2(x4.x, x4.y), zw = make_float2(x4.z, x4.w);
                    a2 = __ffma2_rn(xy, xy, a2);
                    a2 = __ffma2_rn(zw, zw, a2);
                    if (BWD) {
                        u2 = __ffma2_rn(xy, make_float2(dvr[BWD ? i : 0].x, dvr[BWD ? i : 0].y), u2);
                        u2 = __ffma2_rn(zw, make_float2(dvr[BWD ? i : 0].z, dvr[BWD ? i : 0].w), u2);
                    }
                }
                float ssq = a2.x + a2.y, ud = u2.x + u2.y;
                uint32_t ex = 127u;
                const bool fast = __all_sync(0xffffffffu, ssq <= C::FAST_HI * (1.f / 32.f)) &&
                                  __any_sync(0xffffffffu, ssq >= C::FAST_LO);
                if (!fast) {
                    // general path (warp-uniform, rare): power-of-two scale, largest |x| -> [1, 2); zero / denormal /
                    // non-finite rows keep scale 1
                    float mx = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(buf[j][i].x), fabsf(buf[j][i].y))),
                                   fmaxf(fabsf(buf[j][i].z), fabsf(buf[j][i].w)));
                    mx = warp_max(mx);
                    ex = __float_as_uint(mx) >> 23;
                    if (ex == 0u || ex >= 255u) ex = 127u;
                    if (ex > 253u) ex = 253u;
                    const float sc = __uint_as_float((254u - ex) << 23);
                    ssq = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4& x4 = buf[j][i];
                        x4.x *= sc; x4.y *= sc; x4.z *= sc; x4.w *= sc;
                        ssq += x4.x * x4.x + x4.y * x4.y + x4.z * x4.z + x4.w * x4.w;
                    }
                }
                auto level = [&](int o) {
                    ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
                    if (BWD) ud += __shfl_xor_sync(0xffffffffu, ud, o);
                };
                level(16);
                convert_part(j, 0);
                level(8);
                convert_part(j, 1);
                level(4);
                convert_part(j, 2);
                level(2);
                convert_part(j, 3);
                level(1);
                if (lane == 0) {
                    // score = info.x * (Qn . x~), info.x = scale / max(|x~|, eps 2^-e);  x = 2^e x~ with 2^e = info.y;
                    // info.z = dv . x / P (backward).  1 / |x~| = rsqrt + one Newton step (<= 2 ulp).
                    float4 info;
                    info.y = __uint_as_float(ex << 23);
                    float y = rsqrtf(ssq);
                    y = y * fmaf(-0.5f * ssq * y, y, 1.5f);
                    y = fminf(y, info.y * (1.f / VLSA_NORM_EPS));          // also catches ssq == 0 (NaN -> cap)
                    info.x = prm.scale * y;
                    info.z = BWD ? ud : 0.f;
                    info.w = 0.f;
                    *reinterpret_cast<float4*>(s_rowinfo + (b * TR + C::RPW * pw + j) * 4) = info;
                }
            };
            // schedule: the proxy fence below waits for every outstanding load of the thread; with one row per
            // register set only rows 0-2 of the next tile are in flight at the fence (the youngest a quarter of the
            // conversion old), row 3 is issued right after it: tile period = load latency + a quarter conversion
            mbar_wait_wd(empty + b, (u & 1u) ^ 1u);
#pragma unroll
            for (int j = 0; j < C::RPW; ++j) {
                process_row(j);
                if (j < C::EARLY && nvalid) issue_row(j, nptr, nrows);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + b);
            if (nvalid) {
#pragma unroll
                for (int j = C::EARLY; j < C::RPW; ++j) issue_row(j, nptr, nrows);
            }
            ++tt;
            if (pw == 0 && lane == 0) *reinterpret_cast<volatile uint32_t*>(s_prog) = tt;
            valid = nvalid; tptr = nptr; rows_left = nrows;
        }
    } else if (warp >= C::NSOFT) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
      if (warp == C::W_PF) {
        // =========================================================================== L2 prefetcher
        // one bulk prefetch per tile, PF tiles ahead of the tile the producers are filling: DRAM latency is absorbed
        // by the 126 MB L2, so one tile of loads in flight per SM (64 KB of registers) streams at full rate
        if (C::PF > 0) {
            const float* X = reinterpret_cast<const float*>(prm.X);
            int fcc = blockIdx.x; long long frow = 0, fr1 = 0; bool fvalid = fcc < prm.total_chunks;
            if (fvalid) { int fb; chunk_info(prm, fcc, fb, frow, fr1); }
            auto prefetch_tile = [&]() {
                if (!fvalid) return;
                const long long n = fr1 - frow < TR ? fr1 - frow : TR;
#ifdef VLSA_TC_LANEPF
                // one 128-byte line per lane and instruction: 16 lines per row
                for (int l = lane; l < int(n) * 16; l += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(X + frow * D + l * 32));
#else
                if (lane == 0) l2_prefetch_bulk(X + frow * D, uint32_t(n) * D * 4u);
#endif
                frow += TR;
                if (frow >= fr1) {
                    fcc += gridDim.x;
                    fvalid = fcc < prm.total_chunks;
                    if (fvalid) { int fb; chunk_info(prm, fcc, fb, frow, fr1); }
                }
            };
            // paced by the producers' progress counter (monotonic: no parity aliasing if this warp lags)
            volatile uint32_t* prog = s_prog;
            uint32_t done = 0;
            while (fvalid) {
                if (done < *prog + uint32_t(C::PF + 1)) { prefetch_tile(); ++done; }
                else __nanosleep(128);
            }
        }
        __syncwarp();
    } else if (warp == C::W_G1) {
        // =========================================================================== GEMM1 issuer
        if (elect_one()) {
            constexpr uint32_t idesc1 = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * TR, false, false);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(ring), 16, 1024);
            const uint32_t d1 = tmem + C::TM_D1, tq0 = tmem + C::TM_Q;
            uint32_t tt = 0;
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t b = tt % C::NBUF, u = tt / C::NBUF;
                    mbar_wait_wd(s_free, (tt & 1u) ^ 1u);           // scores of the previous tile have been read
                    mbar_wait_wd(full + b, u & 1u);
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(desc0, b * C::TILE);
#pragma unroll
                    for (int s = 0; s < C::NSLOT; ++s) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)              // B = 64 rows: hi plane | lo plane of the slot
                            tc_mma_ts(d1, tq0 + (s * 4 + ks) * 8, umma_desc_advance(tb, s * C::SLOT + ks * 32), idesc1,
                                      (s | ks) != 0);
                    }
                    tc_commit(s_ready);
                }
            }
        }
        __syncwarp();
      } else if (warp == C::W_G2) {
        // =========================================================================== GEMM2 issuer
        if (elect_one()) {
            constexpr uint32_t idesc_hi = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * NP, true, false);
            constexpr uint32_t idesc_lo = umma_idesc(UMMA_F16, UMMA_F16, 128, NP, true, false);
            const uint64_t a0 = umma_desc_sw128(smem_u32(ring), C::SLOT, 1024);
            const uint64_t w0 = umma_desc_sw128(smem_u32(wt), 16, 1024);
            uint32_t tt = 0, cc = 0;
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t i = tt & 1u, v = tt >> 1, b = tt % C::NBUF;
                    mbar_wait_wd(w_ready + i, v & 1u);
                    if (t == 0) mbar_wait_wd(d2_free, (cc & 1u) ^ 1u);   // previous chunk's accumulators drained
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(a0, b * C::TILE), wb = umma_desc_advance(w0, i * C::WBUF);
                    const uint32_t acc0 = t != 0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t d2 = tmem + C::TM_D2 + g * C::D2W;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {              // 16 tile rows per step
                            const uint64_t ah = umma_desc_advance(tb, (2 * g) * C::SLOT + ks * 2048);
                            const uint64_t bd = umma_desc_advance(wb, ks * 32);
                            tc_mma_ss(d2, ah, bd, idesc_hi, ks ? 1u : acc0);
                            tc_mma_ss(d2 + 2 * NP, umma_desc_advance(ah, C::PLANE), bd, idesc_lo, ks ? 1u : acc0);
                        }
                    }
                    tc_commit(empty + b);
                    tc_commit(w_free + i);
                    if (t == ntiles - 1) tc_commit(d2_done);
                }
            }
        }
        __syncwarp();
      }
    } else {
        // =========================================================================== weights / drain
        if (C::NPROD != 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_SOFT));
        // this thread: TMEM quadrant q (prototype p = 4 q + (lane & 3)), tile rows n0, n0 + 1
        const int q = warp & 3, half = warp >> 2;
        const int pl = lane & 3, p = 4 * q + pl, grp = lane >> 2;
        const int n0 = 16 * half + 2 * grp;
        const bool pvalid = p < P;
        const uint32_t tq = tmem + (uint32_t(32 * q) << 16);
        constexpr int NT = C::NSOFT * 32;
        uint32_t tt = 0, cc = 0;
        for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
            int bag; long long r0, r1;
            chunk_info(prm, c, bag, r0, r1);
            const int chunk_nrows = int(r1 - r0);
            const int ntiles = (chunk_nrows + TR - 1) / TR;
            // forward: m_ref = softmax reference, lsum = running sum, exE = chunk reference exponent E
            //          (accumulators hold 2^-E O)
            // backward: m_ref = log2 of the normaliser H_p (accumulators hold dQn_p / H_p)
            float m_ref = -INFINITY, lsum = 0.f;
            int exE = 127;
            float bw_m = 0.f, bw_il = 0.f, bw_delta = 0.f;
            if (BWD && pvalid) {
                bw_m = __ldg(prm.ml + (size_t(bag) * P + p) * 2);
                bw_il = 1.f / __ldg(prm.ml + (size_t(bag) * P + p) * 2 + 1);
                bw_delta = __ldg(prm.delta + size_t(bag) * P + p);
            }
            for (int t = 0; t < ntiles; ++t, ++tt) {
                const uint32_t i = tt & 1u, v = tt >> 1, b = tt % C::NBUF, u = tt / C::NBUF;
                const int nvalid = min(TR, chunk_nrows - t * TR);
                mbar_wait_wd(s_ready, tt & 1u);
                tc_fence_after();
                float sc2[2];
                {
                    // 32 partial scores of this lane's (prototype, part, range): rows 16 half .. +15 x (hi | lo plane)
                    uint32_t sa[16], sb[16];
                    tmem_ld16(tq + C::TM_D1 + 16 * half, sa);
                    tmem_ld16(tq + C::TM_D1 + 32 + 16 * half, sb);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                    // add the planes, then a transposed butterfly over the 8 lanes (part, range) of this prototype:
                    // lane bits 4, 3, 2 select which half of the rows a lane keeps -> rows n0, n0 + 1
                    float a8[8], a4[4];
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const float lo_half = __uint_as_float(sa[n]) + __uint_as_float(sb[n]);
                        const float hi_half = __uint_as_float(sa[8 + n]) + __uint_as_float(sb[8 + n]);
                        const bool up = lane & 16;
                        const float keep = up ? hi_half : lo_half, send = up ? lo_half : hi_half;
                        a8[n] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const bool up = lane & 8;
                        const float keep = up ? a8[4 + n] : a8[n], send = up ? a8[n] : a8[4 + n];
                        a4[n] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int n = 0; n < 2; ++n) {
                        const bool up = lane & 4;
                        const float keep = up ? a4[2 + n] : a4[n], send = up ? a4[n] : a4[2 + n];
                        sc2[n] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                mbar_wait_wd(full + b, u & 1u);                        // acquire the producers' row info
                float4 info[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) info[k] = *reinterpret_cast<const float4*>(s_rowinfo + (b * TR + n0 + k) * 4);
                float w[2];                                            // weights fed to GEMM2 (before the fp16 split)
                if (!BWD) {
                    if (t == 0) exE = int(__float_as_uint(s_rowinfo[(b * TR) * 4 + 1]) >> 23);
                    // ts = score + (e_row - E) ln 2: the weight is exp(ts - m_ref) = A-weight x 2^(e_row - E)
                    float ts[2], unscale[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        int de = int(__float_as_uint(info[k].y) >> 23) - exE;
                        de = de < -100 ? -100 : (de > 100 ? 100 : de);
                        ts[k] = (n0 + k < nvalid) ? fmaf(float(de), 0.693147180559945f, sc2[k] * info[k].x) : -INFINITY;
                        unscale[k] = __uint_as_float(uint32_t(127 - de) << 23);       // 2^-(e_row - E)
                    }
                    float mt = fmaxf(ts[0], ts[1]);
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
                    const bool grow = pvalid && (mt > m_ref + C::MARGIN);             // true on the first tile
                    if (named_bar_or(1, NT, grow)) {
                        // rare: both halves agree on the new reference of every prototype; from the second tile
                        // on the TMEM accumulators are rescaled once GEMM2 of the previous tile has completed
                        if (lane < 4) s_cand[16 * half + p] = grow ? mt + C::HEADROOM : m_ref;
                        named_bar_sync(2, NT);
                        const float m_new = fmaxf(s_cand[p], s_cand[16 + p]);
                        if (t > 0) {
                            const float alpha = (pvalid && m_new > m_ref) ? expf(m_ref - m_new) : 1.f;
                            if (half == 0 && lane < 4) s_alpha[p] = alpha;
                            mbar_wait_wd(w_free + ((tt - 1) & 1u), ((tt - 1) >> 1) & 1u);
                            tc_fence_after();
                            named_bar_sync(3, NT);
                            float al[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) al[j] = s_alpha[j];
#pragma unroll 1
                            for (int k = 6 * half; k < 6 * half + 6; ++k) {
                                uint32_t o[16];
                                tmem_ld16(tq + C::TM_D2 + 16 * k, o);
                                tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * al[j]);
                                tmem_st16(tq + C::TM_D2 + 16 * k, o);
                            }
                            tmem_wait_st();
                            tc_fence_before();
                            lsum *= alpha;
                        }
                        m_ref = m_new;
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        w[k] = (pvalid && n0 + k < nvalid) ? expf(ts[k] - m_ref) : 0.f;
                        lsum = fmaf(w[k], unscale[k], lsum);
                    }
                } else {
                    // c = scale A (u - delta) / |x| = A (u - delta) info.x 2^-e; the weight on x~ = 2^-e x is c 2^e
                    float cw[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float a = expf(sc2[k] * info[k].x - bw_m) * bw_il;        // A_pn (deepmil.py:198)
                        cw[k] = (pvalid && n0 + k < nvalid) ? a * (info[k].z - bw_delta) * info[k].x : 0.f;
                    }
                    float mt = fmaxf(fabsf(cw[0]), fabsf(cw[1]));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
                    // binary exponent of the tile maximum (zero / denormal -> very small, non-finite -> very large)
                    int et = int((__float_as_uint(mt) >> 23) & 0xffu) - 127;
                    et = et < -100 ? -100 : (et > 100 ? 100 : et);
                    const bool grow = pvalid && (t == 0 || float(et) > m_ref + float(C::BWD_MAXE));
                    if (named_bar_or(1, NT, grow)) {
                        if (lane < 4) s_cand[16 * half + p] = grow ? float(et - C::BWD_SETE) : m_ref;
                        named_bar_sync(2, NT);
                        const float m_new = fmaxf(s_cand[p], s_cand[16 + p]);
                        if (t > 0) {
                            // exact: power-of-two ratio H_old / H_new
                            const float alpha = (pvalid && m_new > m_ref) ? exp2f(m_ref - m_new) : 1.f;
                            if (half == 0 && lane < 4) s_alpha[p] = alpha;
                            mbar_wait_wd(w_free + ((tt - 1) & 1u), ((tt - 1) >> 1) & 1u);
                            tc_fence_after();
                            named_bar_sync(3, NT);
                            float al[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) al[j] = s_alpha[j];
#pragma unroll 1
                            for (int k = 6 * half; k < 6 * half + 6; ++k) {
                                uint32_t o[16];
                                tmem_ld16(tq + C::TM_D2 + 16 * k, o);
                                tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * al[j]);
                                tmem_st16(tq + C::TM_D2 + 16 * k, o);
                            }
                            tmem_wait_st();
                            tc_fence_before();
                        }
                        m_ref = m_new;
                    }
                    const float inv_h = pvalid ? __uint_as_float(uint32_t(127 - int(m_ref)) << 23) : 0.f;   // 1 / H_p
#pragma unroll
                    for (int k = 0; k < 2; ++k) w[k] = cw[k] * inv_h;
                }
                // weights of this tile as two fp16 terms (w = t0 + 2^-11 t1); B operand row (term * 16 + p), K = tile row
                uint32_t b0[2], b1[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const __half h0 = __float2half_rn(w[k]);
                    const __half h1 = __float2half_rn((w[k] - __half2float(h0)) * 2048.f);
                    b0[k] = __half_as_ushort(h0); b1[k] = __half_as_ushort(h1);
                }
                mbar_wait_wd(w_free + i, (v & 1u) ^ 1u);               // GEMM2 of tile tt-2 has read this buffer
                unsigned char* wb = wt + i * C::WBUF;
                *reinterpret_cast<uint32_t*>(wb + sw128_offset(p, n0 >> 3, (2 * n0) & 15)) = b0[0] | (b0[1] << 16);
                *reinterpret_cast<uint32_t*>(wb + sw128_offset(NP + p, n0 >> 3, (2 * n0) & 15)) = b1[0] | (b1[1] << 16);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(w_ready + i);
            }
            // ---- chunk end: drain O^T (lane = feature within a 128-block, columns = hi.t0 | hi.t1 | lo.t0 per prototype)
            if (BWD) {
                // per-prototype normalisers for the drain (every weight warp holds its own four)
                if (half == 0 && lane < 4) s_alpha[p] = pvalid ? __uint_as_float(uint32_t(127 + int(m_ref)) << 23) : 0.f;
                named_bar_sync(2, NT);
            }
            mbar_wait_wd(d2_done, cc & 1u);
            tc_fence_after();
            float* po = prm.part_O + size_t(c) * P * D;
            float mul[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) mul[j] = BWD ? s_alpha[j] : __uint_as_float(uint32_t(exE) << 23);
#pragma unroll 1
            for (int g = 2 * half; g < 2 * half + 2; ++g) {
                uint32_t o0[16], o1[16], o2[16];
                tmem_ld16(tq + C::TM_D2 + g * C::D2W, o0);
                tmem_ld16(tq + C::TM_D2 + g * C::D2W + 16, o1);
                tmem_ld16(tq + C::TM_D2 + g * C::D2W + 32, o2);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < P) po[size_t(j) * D + 128 * g + 32 * q + lane] =
                        mul[j] * (fmaf(__uint_as_float(o1[j]), 0x1p-11f, __uint_as_float(o2[j])) + __uint_as_float(o0[j]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_free);
            if (!BWD) {
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
                if (half == 1 && lane < 4) s_lsum[p] = lsum;
                named_bar_sync(2, NT);
                if (half == 0 && lane < 4 && pvalid) {
                    prm.part_l[size_t(c) * P + p] = lsum + s_lsum[p];
                    prm.part_m[size_t(c) * P + p] = m_ref;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::W_G1) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace vlsa
