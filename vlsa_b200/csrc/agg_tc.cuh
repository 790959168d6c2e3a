// Fused language-guided aggregation on the 5th-generation tensor cores (tcgen05 + TMEM), forward pass.
// Same contract as agg_simt_kernel<P,false,float>: ONE read of X, per-chunk online-softmax partials
// (m, l, O[P,D]) — with both skinny contractions on tcgen05.mma (kind::f16, fp32 accumulators in TMEM):
//
//   GEMM1  S^T[128, 64]  = Qn'[128, 512] (A, resident in TMEM) . [X_hi ; X_lo][64, 512]^T (B, K-major smem)
//   GEMM2  O^T[512 d, 32 | 16] += X_hi^T | X_lo^T [512, 32] (A, MN-major, the SAME smem bytes) . W[32 | 16, 32]^T
//
// Precision design (the tensor core truncates its fp32 accumulator toward zero after every instruction —
// measured with scripts/dev_tc_unit.cu — so the number of accumulation steps per accumulator is kept small):
//   * every row of X is scaled by a power of two (largest |x| -> [1,2)) and split into fp16 (hi, lo) planes:
//     22 significant bits at 4 bytes of shared memory per element; Qn is split the same way;
//   * A-operand row (TMEM lane) 32 w + j holds prototype p = 4 w + (j & 3), part (j >> 2) & 1 (hi / lo) of Qn
//     restricted to the feature range 128 (j >> 3) .. +127 (zeros elsewhere): every accumulator only sees
//     8 non-zero steps, and the 16 partial sums of a score (2 parts x 4 ranges x 2 planes) are added in fp32
//     registers.  Softmax warp w therefore owns prototypes 4 w .. 4 w + 3 completely (no cross-warp traffic);
//   * the softmax weights go to the tensor core as two fp16 terms scaled 1 and 2^11 (22 bits); products with
//     the hi and the lo plane of X accumulate in separate TMEM columns; tcgen05.mma needs A and B in the same
//     16-bit format (fp16 x bf16 traps), hence fp16 weights with the lazy-rescale range control below.
//
// Warp roles (14 warps, 1 persistent CTA / SM, static round-robin over chunks):
//   warps 0-3   softmax : TMEM scores -> online softmax (lazy rescale) -> weights to smem; drain O per chunk
//   warp  4     GEMM1 issuer (one thread) + TMEM allocation
//   warp  5     GEMM2 issuer (one thread)
//   warps 6-13  producers: L2 bulk prefetch ahead; LDG.128 of whole rows (one tile in flight in registers)
//               -> power-of-two scale / fp16 split -> swizzled STS
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
    // floats: rowinfo[NBUF][TR][2] | alpha[16]
    static constexpr int NFLOAT = NBUF * TR * 2 + 16;
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 2 * NBUF + 8;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int NWARPS = 14;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int NPROD = 8;
    static constexpr int PF = 3;                  // L2 prefetch distance in tiles (3 x 64 KB x 148 SMs = 28 MB of L2)
    static constexpr int QPITCH = D + 1;          // prologue staging of Qn (aliases the ring)
    // TMEM columns: Qn operand | O^T accumulators: 4 blocks of 128 d x (hi.t0 16 | hi.t1 16 | lo.t0 16) | scores
    static constexpr int TM_Q = 0;
    static constexpr int D2W = 3 * NP;            // 48 columns per 128-feature block
    static constexpr int TM_D2 = 256;
    static constexpr int TM_D1 = TM_D2 + 4 * D2W;   // 448: scores [hi plane rows 0..31 | lo plane rows 0..31]
    static constexpr int TMEM_COLS = 512;
    // online softmax with lazy rescaling: on (re)set the reference is the running maximum + HEADROOM; the TMEM
    // accumulators are rescaled only when a tile maximum exceeds the reference by more than MARGIN, so the
    // fp16 weight terms stay within (0, e^MARGIN] (e^10 = 22026 < 65504) with an absolute floor of 2^-35.
    static constexpr float HEADROOM = 6.f;
    static constexpr float MARGIN = 10.f;
};

// x = hi + lo with hi, lo fp16 (packed pairs, first element in the low half)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
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

__global__ void __launch_bounds__(TcCfg::THREADS, 1) agg_tc_kernel(const AggParams prm, const int P) {
    using C = TcCfg;
    constexpr int D = C::D, NP = C::NP, TR = C::TR;
    if (int(blockIdx.x) >= prm.total_chunks) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* ring = sm;
    unsigned char* wt = sm + C::OFF_W;
    float* s_rowinfo = reinterpret_cast<float*>(sm + C::OFF_F);      // [NBUF][TR] x (score factor, 2^e)
    float* s_alpha = s_rowinfo + C::NBUF * TR * 2;                   // [16] rescale factors (rare path)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* full = bars;                       // [NBUF] producers (8 warps)      -> GEMM1, softmax (row info)
    uint64_t* empty = bars + C::NBUF;            // [NBUF] GEMM2 commit             -> producers
    uint64_t* s_ready = bars + 2 * C::NBUF;      //        GEMM1 commit             -> softmax
    uint64_t* s_free = s_ready + 1;              //        softmax (4 warps)        -> GEMM1
    uint64_t* w_ready = s_ready + 2;             // [2]    softmax (4 warps)        -> GEMM2
    uint64_t* w_free = s_ready + 4;              // [2]    GEMM2 commit             -> softmax
    uint64_t* d2_done = s_ready + 6;             //        last GEMM2 of a chunk    -> softmax (drain)
    uint64_t* d2_free = s_ready + 7;             //        softmax (4 warps)        -> GEMM2 of the next chunk
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < C::NBUF; ++s) { mbar_init(full + s, C::NPROD); mbar_init(empty + s, 1); }
        mbar_init(s_ready, 1); mbar_init(s_free, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(w_ready + s, 4); mbar_init(w_free + s, 1); }
        mbar_init(d2_done, 1); mbar_init(d2_free, 4);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    // ---- prologue: Qn = Q / max(|Q|, eps) staged as fp32 in the (still unused) ring, rows >= P are zero
    {
        float* qn = reinterpret_cast<float*>(ring);
        for (int p = warp; p < NP; p += C::NWARPS) {
            float inv = 0.f;
            if (p < P) {
                float ss = 0.f;
                for (int d = lane; d < D; d += 32) { const float v = __ldg(prm.Q + size_t(p) * D + d); ss += v * v; }
                ss = warp_sum(ss);
                inv = 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
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

    if (warp >= 6) {
        // =========================================================================== producers
        const int pw = warp - 6;
        const float* X = reinterpret_cast<const float*>(prm.X);
        const uint64_t policy = make_evict_first_policy();
        int cc = blockIdx.x; long long row = 0, r1 = 0; bool valid = cc < prm.total_chunks;
        if (valid) { int bag; chunk_info(prm, cc, bag, row, r1); }
        float4 buf[4][4];
        auto issue_row = [&](int j) {
            const long long r = row + 4 * pw + j;
            if (valid && r < r1) {
                const float* src = X + r * D + 4 * lane;
#pragma unroll
                for (int i = 0; i < 4; ++i) buf[j][i] = ldg_stream_f4(src + 128 * i, policy);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) buf[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
#pragma unroll
        for (int j = 0; j < 4; ++j) issue_row(j);
        // L2 prefetch cursor, PF tiles ahead of the register loads: DRAM latency is absorbed by the 126 MB L2,
        // so one tile of loads in flight per SM (64 KB of registers) is enough to stream at full rate
        int fcc = blockIdx.x; long long frow = 0, fr1 = 0; bool fvalid = fcc < prm.total_chunks;
        if (fvalid) { int bag; chunk_info(prm, fcc, bag, frow, fr1); }
        auto prefetch_tile = [&]() {
            if (!fvalid) return;
            const long long r = frow + 4 * pw;
            if (lane == 0 && r < fr1) {
                const long long n = fr1 - r < 4 ? fr1 - r : 4;
                l2_prefetch_bulk(X + r * D, uint32_t(n) * D * 4u);
            }
            frow += TR;
            if (frow >= fr1) {
                fcc += gridDim.x;
                fvalid = fcc < prm.total_chunks;
                if (fvalid) { int bag; chunk_info(prm, fcc, bag, frow, fr1); }
            }
        };
#pragma unroll 1
        for (int k = 0; k < C::PF; ++k) prefetch_tile();
        uint32_t tt = 0;
        while (valid) {
            prefetch_tile();
            // advance the load cursor to the next tile of this CTA
            row += TR;
            if (row >= r1) {
                cc += gridDim.x;
                valid = cc < prm.total_chunks;
                if (valid) { int bag; chunk_info(prm, cc, bag, row, r1); }
            }
            const uint32_t b = tt % C::NBUF, u = tt / C::NBUF;
            // per-row power-of-two scale: largest |x| -> [1, 2); the four rows reduce as independent shuffle chains
            float mx[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                mx[j] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    mx[j] = fmaxf(fmaxf(mx[j], fmaxf(fabsf(buf[j][i].x), fabsf(buf[j][i].y))),
                                  fmaxf(fabsf(buf[j][i].z), fabsf(buf[j][i].w)));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
            }
            mbar_wait_wd(empty + b, (u & 1u) ^ 1u);
            unsigned char* tile = ring + b * C::TILE;
            float ssq[4];
            uint32_t exs[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * pw + j;
                // zero / denormal / non-finite rows keep scale 1
                uint32_t ex = __float_as_uint(mx[j]) >> 23;
                if (ex == 0u || ex >= 255u) ex = 127u;
                if (ex > 253u) ex = 253u;
                exs[j] = ex;
                const float sc = __uint_as_float((254u - ex) << 23);
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a0 = buf[j][i].x * sc, a1 = buf[j][i].y * sc, a2 = buf[j][i].z * sc, a3 = buf[j][i].w * sc;
                    acc += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
                    uint32_t h0, l0, h1, l1;
                    split_f16x2(a0, a1, h0, l0);
                    split_f16x2(a2, a3, h1, l1);
                    // columns 128 i + 4 lane .. +3: slot 2 i + (lane >> 4), 16-byte chunk (lane & 15) >> 1, half (lane & 1)
                    unsigned char* hi = tile + (2 * i + (lane >> 4)) * C::SLOT + sw128_offset(r, (lane & 15) >> 1, (lane & 1) * 8);
                    *reinterpret_cast<uint2*>(hi) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(hi + C::PLANE) = make_uint2(l0, l1);
                }
                ssq[j] = acc;
                issue_row(j);                               // same register slot, next tile
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
            }
            if (lane < 4) {
                const float sq = lane == 0 ? ssq[0] : (lane == 1 ? ssq[1] : (lane == 2 ? ssq[2] : ssq[3]));
                const uint32_t ex = lane == 0 ? exs[0] : (lane == 1 ? exs[1] : (lane == 2 ? exs[2] : exs[3]));
                float2 info;
                // score = info.x * (Qn . x~);  x = 2^e x~ with 2^e = info.y
                info.x = prm.scale / fmaxf(sqrtf(sq), VLSA_NORM_EPS * __uint_as_float((254u - ex) << 23));
                info.y = __uint_as_float(ex << 23);
                *reinterpret_cast<float2*>(s_rowinfo + (b * TR + 4 * pw + lane) * 2) = info;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + b);
            ++tt;
        }
    } else if (warp == 4) {
        // =========================================================================== GEMM1 issuer
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * TR, false, false);
            const uint32_t ring_a = smem_u32(ring);
            const uint32_t d1 = tmem + C::TM_D1;
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
                    const uint32_t tb = ring_a + b * C::TILE;
#pragma unroll 1
                    for (int s = 0; s < C::NSLOT; ++s) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)              // B = 64 rows: hi plane | lo plane of the slot
                            tc_mma_ts(d1, tmem + C::TM_Q + (s * 4 + ks) * 8,
                                      umma_desc_sw128(tb + s * C::SLOT + ks * 32, 16, 1024), idesc1, (s | ks) != 0);
                    }
                    tc_commit(s_ready);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // =========================================================================== GEMM2 issuer
        if (lane == 0) {
            constexpr uint32_t idesc_hi = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * NP, true, false);
            constexpr uint32_t idesc_lo = umma_idesc(UMMA_F16, UMMA_F16, 128, NP, true, false);
            const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wt);
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
                    const uint32_t tb = ring_a + b * C::TILE, wb = w_a + i * C::WBUF;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t d2 = tmem + C::TM_D2 + g * C::D2W;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {              // 16 tile rows per step
                            const uint32_t ah = tb + (2 * g) * C::SLOT + ks * 2048;
                            const uint64_t bd = umma_desc_sw128(wb + ks * 32, 16, 1024);
                            tc_mma_ss(d2, umma_desc_sw128(ah, C::SLOT, 1024), bd, idesc_hi, (t | ks) != 0);
                            tc_mma_ss(d2 + 2 * NP, umma_desc_sw128(ah + C::PLANE, C::SLOT, 1024), bd, idesc_lo, (t | ks) != 0);
                        }
                    }
                    tc_commit(empty + b);
                    tc_commit(w_free + i);
                    if (t == ntiles - 1) tc_commit(d2_done);
                }
            }
        }
        __syncwarp();
    } else {
        // =========================================================================== softmax / drain
        const int pl = lane & 3, p = 4 * warp + pl, grp = lane >> 2;       // this thread: prototype p, tile rows 4 grp .. +3
        const bool pvalid = p < P;
        const uint32_t tq = tmem + (uint32_t(32 * warp) << 16);
        uint32_t tt = 0, cc = 0;
        for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
            int bag; long long r0, r1;
            chunk_info(prm, c, bag, r0, r1);
            const int ntiles = int((r1 - r0 + TR - 1) / TR);
            float m_ref = -INFINITY, lsum = 0.f;
            int exE = 127;                       // chunk reference exponent E: accumulators hold 2^-E O
            for (int t = 0; t < ntiles; ++t, ++tt) {
                const uint32_t i = tt & 1u, v = tt >> 1, b = tt % C::NBUF, u = tt / C::NBUF;
                const long long left = r1 - (r0 + (long long)t * TR);
                const int nvalid = left < TR ? int(left) : TR;
                mbar_wait_wd(s_ready, tt & 1u);
                tc_fence_after();
                float sc4[4];
                {
                    // 64 partial scores of this lane's (prototype, part, range): rows 0..31 x (hi | lo plane)
                    uint32_t sv[64];
                    tmem_ld32(tq + C::TM_D1, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
                    tmem_ld32(tq + C::TM_D1 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                    // add the planes, then a transposed butterfly over the 8 lanes (part, range) of this prototype:
                    // lane bits 4, 3, 2 select which half of the rows a lane keeps -> rows 4 grp .. 4 grp + 3
                    float a16[16], a8[8];
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        const float lo_half = __uint_as_float(sv[n]) + __uint_as_float(sv[32 + n]);
                        const float hi_half = __uint_as_float(sv[16 + n]) + __uint_as_float(sv[48 + n]);
                        const bool up = lane & 16;
                        const float keep = up ? hi_half : lo_half, send = up ? lo_half : hi_half;
                        a16[n] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const bool up = lane & 8;
                        const float keep = up ? a16[8 + n] : a16[n], send = up ? a16[n] : a16[8 + n];
                        a8[n] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const bool up = lane & 4;
                        const float keep = up ? a8[4 + n] : a8[n], send = up ? a8[n] : a8[4 + n];
                        sc4[n] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                mbar_wait_wd(full + b, u & 1u);                        // acquire the producers' row info
                if (t == 0) exE = int(__float_as_uint(s_rowinfo[(b * TR) * 2 + 1]) >> 23);
                // ts = score + (e_row - E) ln 2: the weight fed to GEMM2 is exp(ts - m_ref) = A-weight x 2^(e_row - E)
                const int n0 = 4 * grp;
                float ts[4], unscale[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 info = *reinterpret_cast<const float2*>(s_rowinfo + (b * TR + n0 + k) * 2);
                    int de = int(__float_as_uint(info.y) >> 23) - exE;
                    de = de < -100 ? -100 : (de > 100 ? 100 : de);
                    ts[k] = (n0 + k < nvalid) ? fmaf(float(de), 0.693147180559945f, sc4[k] * info.x) : -INFINITY;
                    unscale[k] = __uint_as_float(uint32_t(127 - de) << 23);       // 2^-(e_row - E)
                }
                float mt = fmaxf(fmaxf(ts[0], ts[1]), fmaxf(ts[2], ts[3]));
                mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
                mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
                mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
                const bool grow = pvalid && (mt > m_ref + C::MARGIN);             // always true on the first tile
                const bool any_grow = named_bar_or(1, 128, grow);
                if (any_grow) {
                    const float m_new = grow ? mt + C::HEADROOM : m_ref;
                    if (t > 0) {
                        // a later tile beats the reference by more than the margin: rescale the TMEM accumulators
                        // once GEMM2 of the previous tile has completed (every warp owns its 32 TMEM lanes)
                        const float alpha = grow ? expf(m_ref - m_new) : 1.f;
                        if (lane < 4) s_alpha[p] = alpha;
                        mbar_wait_wd(w_free + ((tt - 1) & 1u), ((tt - 1) >> 1) & 1u);
                        tc_fence_after();
                        named_bar_sync(2, 128);
                        float al[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) al[q] = s_alpha[q];
#pragma unroll 1
                        for (int k = 0; k < 4 * C::D2W / 16; ++k) {
                            uint32_t o[16];
                            tmem_ld16(tq + C::TM_D2 + 16 * k, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int q = 0; q < 16; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * al[q]);
                            tmem_st16(tq + C::TM_D2 + 16 * k, o);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        lsum *= alpha;
                    }
                    m_ref = m_new;
                }
                // weights of this tile as two fp16 terms (w = t0 + 2^-11 t1); B operand row (term * 16 + p), K = tile row
                uint32_t b0[4], b1[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float w = (pvalid && n0 + k < nvalid) ? expf(ts[k] - m_ref) : 0.f;
                    lsum = fmaf(w, unscale[k], lsum);
                    const __half h0 = __float2half_rn(w);
                    const __half h1 = __float2half_rn((w - __half2float(h0)) * 2048.f);
                    b0[k] = __half_as_ushort(h0); b1[k] = __half_as_ushort(h1);
                }
                mbar_wait_wd(w_free + i, (v & 1u) ^ 1u);               // GEMM2 of tile tt-2 has read this buffer
                unsigned char* wb = wt + i * C::WBUF;
                *reinterpret_cast<uint2*>(wb + sw128_offset(p, grp >> 1, 8 * (grp & 1))) = make_uint2(b0[0] | (b0[1] << 16), b0[2] | (b0[3] << 16));
                *reinterpret_cast<uint2*>(wb + sw128_offset(NP + p, grp >> 1, 8 * (grp & 1))) = make_uint2(b1[0] | (b1[1] << 16), b1[2] | (b1[3] << 16));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(w_ready + i);
            }
            // ---- chunk end: drain O^T (lane = feature within a 128-block, columns = hi.t0 | hi.t1 | lo.t0 per prototype)
            mbar_wait_wd(d2_done, cc & 1u);
            tc_fence_after();
            float* po = prm.part_O + size_t(c) * P * D;
            const float pwE = __uint_as_float(uint32_t(exE) << 23);
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                uint32_t o0[16], o1[16], o2[16];
                tmem_ld16(tq + C::TM_D2 + g * C::D2W, o0);
                tmem_ld16(tq + C::TM_D2 + g * C::D2W + 16, o1);
                tmem_ld16(tq + C::TM_D2 + g * C::D2W + 32, o2);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q < P) po[size_t(q) * D + 128 * g + 32 * warp + lane] =
                        pwE * (fmaf(__uint_as_float(o1[q]), 0x1p-11f, __uint_as_float(o2[q])) + __uint_as_float(o0[q]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_free);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
            lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
            if (lane < 4 && pvalid) {
                prm.part_l[size_t(c) * P + p] = lsum;
                prm.part_m[size_t(c) * P + p] = m_ref;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace vlsa
