// Fused language-guided aggregation on the 5th-generation tensor cores (tcgen05 + TMEM), forward and backward
// (fp32 rows, P > 5).  Same contract as agg_simt_kernel<P,BWD,float>: ONE read of X, per-chunk partials — forward:
// online-softmax (m, l, O[P,D]); backward: partial dQn[P,D] — with both skinny contractions on tcgen05.mma (kind::f16,
// fp32 accumulators in TMEM), issued by one elected thread per GEMM:
//
//   GEMM1  S^T[128, 32]  = Qn'[128, 512] (A, resident in TMEM) . [X_hi ; X_lo][32, 512]^T (B, K-major smem)
//   GEMM2  O^T[512 d, 32 | 16] += X_hi^T | X_lo^T [512, 16] (A, MN-major, the SAME smem bytes) . W[32 | 16, 16]^T
//
// Precision design (the tensor core truncates its fp32 accumulator toward zero after every instruction, so the number of
// accumulation steps per accumulator is kept small):
//   * every row of X is split into fp16 (hi, lo) planes: 22 significant bits at 4 bytes of shared memory per element; rows
//     with |x| in [2, 2^14) are split as they are (every CONCH-like row), any other row is first scaled by a power of two
//     from its own norm (|x~| in [1, 2)) that the weight warps undo exactly; Qn is split the same way;
//   * A-operand row (TMEM lane) 32 w + j holds prototype p = 4 w + (j & 3), part (j >> 2) & 1 (hi / lo) of Qn restricted
//     to the feature range 128 (j >> 3) .. +127 (zeros elsewhere): every accumulator only sees 8 non-zero steps, and the
//     16 partial sums of a score (2 parts x 4 ranges x 2 planes) are added in fp32 registers.  TMEM quadrant q therefore
//     owns prototypes 4 q .. 4 q + 3 completely;
//   * the per-row weights go to the tensor core as two fp16 terms scaled 1 and 2^11 (22 bits); products with the hi and
//     the lo plane of X accumulate in separate TMEM columns (a merged accumulator costs the backward three orders of
//     magnitude of gradient accuracy: tiny lo-plane products are absorbed by a large, cancelling accumulator);
//     tcgen05.mma needs A and B in the same 16-bit format, hence fp16 weights kept in range by a lazily rescaled softmax
//     reference (forward) / a lazily grown power-of-two normaliser per prototype (backward, exact rescale).
//
// What bounds this pass on a B200 is SHARED-MEMORY BANDWIDTH, not HBM and not the tensor pipe (ncu of the TMA-fed kernel
// this one replaced, profiles/agg_tc_r02_ncu_summary.md: LSU + tensor-core shared-memory wavefronts = 96 % of the cycles).  Every
// byte of X crosses shared memory once per GEMM (two operand reads) plus whatever it takes to get the fp16 planes there:
//   TMA-fed, converted in place : TMA write + LDS + STS + 2 operand reads = 5 passes  -> 0.70-0.77 of the HBM roofline
//   this kernel                 : STS of the planes     + 2 operand reads = 3 passes
// Rows therefore come in through REGISTERS (LDG.128, read-only path, no L1 allocation, L2 evict-first): a producer warp
// owns two rows of every 16-row tile (lane = 16 columns of each), keeps R tiles in flight in R register sets, and on
// arrival takes the row norm by warp shuffle, splits and stores the planes (STS.64, conflict-free) — no cross-warp
// synchronisation.  Two things make the register path stream at full rate:
//   * NO proxy fence in the producers: fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR
//     waits for every global load the thread has in flight (the tiles being prefetched).  The single thread that issues
//     the MMAs fences instead, after it has acquired the producers' mbarrier (it has no loads in flight);
//   * <= ~160 KB of shared memory per CTA: the loads in flight live in the L1 carve-out; above ~200 KB of shared memory
//     the register path is capped at 6.3 TB/s (profiles/readbw_r01.txt).
//
// Tile = 16 rows: slot s (64 features) at s * 4096; row group g (8 rows) at + g * 2048;
// hi atom at + 0, lo atom at + 1024 (128-byte swizzle).
//
// Warp roles (20 warps, 1 persistent CTA / SM, static round-robin over chunks):
//   warps 0-7   weights (two alternating sets of four, one warp per TMEM quadrant)
//   warp  8     GEMM1 issuer + TMEM allocation      warp 9   GEMM2 issuer      warps 10-11 idle (24 registers)
//   warps 12-19 producers: rows 2 (w - 12), 2 (w - 12) + 1 of every tile
#pragma once
#include "agg_simt.cuh"
#include "tc_common.cuh"

namespace vlsa {

#ifndef VLSA_TC_NBUF
#define VLSA_TC_NBUF 4
#endif
#ifndef VLSA_TC_R
#define VLSA_TC_R 2
#endif
#ifndef VLSA_TC_PF
#define VLSA_TC_PF 0             // L2 prefetch distance in tiles ahead of the producers' loads (0 = off), warp 10
#endif

struct TcCfg {
    static constexpr int D = VLSA_D;
    static constexpr int NP = 16;
    static constexpr int TR = 16;                 // rows per tile
    static constexpr int NSLOT = 8;               // 64-feature slots
    static constexpr int GRP = 2048;              // bytes of one 8-row group of a slot (hi | lo atoms)
    static constexpr int SLOT = 2 * GRP;          // 4 KB
    static constexpr int TILE = NSLOT * SLOT;     // 32 KB
    static constexpr int NBUF = VLSA_TC_NBUF;
    static constexpr int R = VLSA_TC_R;           // tiles in flight per producer warp (register sets of 32 registers)
    static constexpr int WBUF = 2 * NP * 128;     // weight operand: 32 rows (term, prototype) x 128 B (32 B used)
    static constexpr int OFF_W = NBUF * TILE;
    static constexpr int OFF_F = OFF_W + 2 * WBUF;
    // floats: rowinfo[NBUF][TR][4] | alpha[16] | mref[16] | lsum[2][16] | exE[4]
    static constexpr int NFLOAT = NBUF * TR * 4 + 16 + 16 + 32 + 4;
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 2 * NBUF + 12;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int NSOFT = 8, NPROD = 8;
    static constexpr int W_G1 = NSOFT, W_G2 = NSOFT + 1, W_PROD = NSOFT + 4;
    static constexpr int NWARPS = NSOFT + 4 + NPROD;
    static constexpr int THREADS = NWARPS * 32;
    // register budget (setmaxnreg, per warpgroup of four warps): the launch bound grants 96 to each of the 640 threads;
    // 8 x 32 x REG_SOFT + 4 x 32 x REG_ISSUE + 8 x 32 x REG_PROD <= 640 x 96
    static constexpr int REG_SOFT = 80, REG_ISSUE = 24, REG_PROD = 144;
    static_assert(256 * REG_SOFT + 128 * REG_ISSUE + 256 * REG_PROD <= THREADS * 96, "register pool");
    static constexpr int QPITCH = D + 1;
    static constexpr int TM_Q = 0;
    static constexpr int D2W = 3 * NP;
    static constexpr int TM_D2 = 256;
    static constexpr int TM_D1 = TM_D2 + 4 * D2W;   // 448: two score buffers of 32 columns
    static constexpr int TMEM_COLS = 512;
    static constexpr float HEADROOM = 6.f, MARGIN = 10.f;
    static constexpr int BWD_MAXE = 14, BWD_SETE = 6;
};

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
    float* s_alpha = s_rowinfo + C::NBUF * TR * 4;                   // [16] rescale factors (rare path) / drain normalisers
    float* s_mref = s_alpha + 16;                                    // [16] current softmax reference (fwd) | log2 H_p (bwd)
    float* s_lsum = s_mref + 16;                                     // [2][16] per-set softmax sums at a chunk end
    int* s_exE = reinterpret_cast<int*>(s_lsum + 32);                // [4] reference row-scale exponent of the chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* full = bars;                       // [NBUF] producers (8 warps)       -> GEMM1, weight warps (row info)
    uint64_t* empty = bars + C::NBUF;            // [NBUF] GEMM2 commit              -> producers
    uint64_t* s_ready = bars + 2 * C::NBUF;      // [2]    GEMM1 commit              -> weight warps
    uint64_t* s_free = s_ready + 2;              // [2]    weight set s (4 warps)    -> GEMM1
    uint64_t* w_ready = s_ready + 4;             // [2]    weight set s (4 warps)    -> GEMM2
    uint64_t* w_free = s_ready + 6;              // [2]    GEMM2 commit              -> weight warps
    uint64_t* d2_done = s_ready + 8;             //        last GEMM2 of a chunk     -> weight warps (drain)
    uint64_t* d2_free = s_ready + 9;             //        weight warps (8)          -> GEMM2 of the next chunk
    uint64_t* decided = s_ready + 10;            // [2]    weight set s (4 warps): softmax reference settled for its tile -> other set
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    volatile uint32_t* s_prog = tmem_ptr + 1;    // tiles whose loads the producers have issued (L2 prefetcher only)

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for ptxas

    if (tid == 0) {
        for (int s = 0; s < C::NBUF; ++s) { mbar_init(full + s, C::NPROD); mbar_init(empty + s, 1); }
        *s_prog = 0u;
        for (int s = 0; s < 2; ++s) {
            mbar_init(s_ready + s, 1); mbar_init(s_free + s, C::NSOFT / 2);
            mbar_init(w_ready + s, C::NSOFT / 2); mbar_init(w_free + s, 1);
            mbar_init(decided + s, C::NSOFT / 2);
        }
        mbar_init(d2_done, 1); mbar_init(d2_free, C::NSOFT);
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
        for (int cb = 0; cb < 8; ++cb) {
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

    if (warp >= C::W_PROD) {
        // =========================================================================== producers
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_PROD));
        constexpr int R = C::R;
        const int pw = warp - C::W_PROD;                   // tile rows 2 pw, 2 pw + 1
        const uint64_t policy = make_evict_first_policy();
        const float* Xf = reinterpret_cast<const float*>(prm.X);
        // a cursor walks this CTA's tiles: pointer to the tile's first row (already offset by the lane's 4 columns), rows
        // left in its chunk.  Two cursors: `ld` (loads being issued) runs R tiles ahead of `cv` (tile being converted)
        struct Cursor { const float* ptr; int rows_left, cc, bag; bool valid; };
        auto cur_open = [&](Cursor& k) {
            k.valid = k.cc < prm.total_chunks;
            if (k.valid) {
                long long r0, r1;
                chunk_info(prm, k.cc, k.bag, r0, r1);
                k.ptr = Xf + r0 * D + 4 * lane;
                k.rows_left = int(r1 - r0);
            }
        };
        auto cur_next = [&](Cursor& k) {
            k.ptr += TR * D;
            k.rows_left -= TR;
            if (k.rows_left <= 0) { k.cc += gridDim.x; cur_open(k); }
        };
        Cursor ld, cv;
        ld.cc = cv.cc = blockIdx.x; ld.ptr = cv.ptr = nullptr; ld.rows_left = cv.rows_left = 0; ld.bag = cv.bag = 0;
        cur_open(ld);
        cur_open(cv);
        // R register sets of one tile share each: set s, row j, columns 128 i + 4 lane .. + 3.  Rows past the end of the
        // chunk re-read its last row (their weights are masked by the weight warps), so no load is predicated
        float4 buf[R][2][4];
        // swizzled byte offset of this lane's 8-byte store inside a tile buffer, per row j (+ 2 i slots, + 1024 for lo):
        // columns 128 i + 4 lane .. + 3 -> slot 2 i + (lane >> 4), 16-byte chunk (lane & 15) >> 1, half (lane & 1)
        uint32_t soff[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int rr = 2 * pw + j;
            soff[j] = (lane >> 4) * C::SLOT + (rr >> 3) * C::GRP + sw128_offset(rr & 7, (lane & 15) >> 1, (lane & 1) * 8);
        }
        float4 dvr[BWD ? 4 : 1];                           // dv / P at this lane's columns (backward)
        int dv_bag = -1;
        uint32_t b = 0, par = 1;                           // ring position, parity to wait for on empty[b]
        uint32_t n_issued = 0;
        PROF_DECL
#define VLSA_TC_ISSUE(S)                                                                                     \
        if (ld.valid) {                                                                                      \
            _Pragma("unroll") for (int j = 0; j < 2; ++j) {                                                  \
                const float* src = ld.ptr + min(2 * pw + j, ld.rows_left - 1) * D;                           \
                _Pragma("unroll") for (int i = 0; i < 4; ++i) buf[S][j][i] = ldg_stream_f4(src + 128 * i, policy); \
            }                                                                                                \
            cur_next(ld);                                                                                    \
            if (VLSA_TC_PF > 0 && pw == 0 && lane == 0) *s_prog = ++n_issued;                                \
        }
#pragma unroll
        for (int s = 0; s < R; ++s) { VLSA_TC_ISSUE(s) }
        while (cv.valid) {
#pragma unroll
            for (int s = 0; s < R; ++s) {
                if (!cv.valid) break;
                if (BWD && cv.bag != dv_bag) {
                    dv_bag = cv.bag;
                    const float invP = 1.f / float(P);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(prm.dv + size_t(dv_bag) * D + 128 * i + 4 * lane));
                        dvr[BWD ? i : 0] = make_float4(t4.x * invP, t4.y * invP, t4.z * invP, t4.w * invP);
                    }
                }
                // per-lane partial |x|^2 (and dv . x / P) of both rows, then one interleaved five-level butterfly
                float ssq[2], ud[2];
                PROF_BEGIN();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float2 a2 = make_float2(0.f, 0.f), u2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 x4 = buf[s][j][i];
                        const float2 xy = make_float2(x4.x, x4.y), zw = make_float2(x4.z, x4.w);
                        a2 = __ffma2_rn(xy, xy, a2);
                        a2 = __ffma2_rn(zw, zw, a2);
                        if (BWD) {
                            const float4 d4 = dvr[BWD ? i : 0];
                            u2 = __ffma2_rn(xy, make_float2(d4.x, d4.y), u2);
                            u2 = __ffma2_rn(zw, make_float2(d4.z, d4.w), u2);
                        }
                    }
                    ssq[j] = a2.x + a2.y;
                    ud[j] = u2.x + u2.y;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
                        if (BWD) ud[j] += __shfl_xor_sync(0xffffffffu, ud[j], o);
                    }
                }
                PROF_END(1);
                // rows with |x| in [2, 2^14) are split as they are (every CONCH-like row); any other row is first scaled
                // by a power of two (|x~| in [1, 2)) that the weight warps undo exactly.  Warp-uniform.
                int e[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t ex = __float_as_uint(ssq[j]) >> 23;               // biased exponent (ssq >= 0)
                    if (ex == 0u || ex >= 255u) ex = 127u;                     // zero / denormal / non-finite rows: scale 1
                    e[j] = (int(ex) - 127) >> 1;                               // x = 2^e x~
                    if (e[j] >= 1 && e[j] <= 13) e[j] = 0;
                    if (e[j] != 0) {                                           // rare
                        const float sc = __uint_as_float(uint32_t(127 - e[j]) << 23);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { buf[s][j][i].x *= sc; buf[s][j][i].y *= sc; buf[s][j][i].z *= sc; buf[s][j][i].w *= sc; }
                    }
                }
                PROF_BEGIN();
                mbar_wait_wd(empty + b, par);                  // GEMM2 of the tile that used this buffer is done
                PROF_END(0);
                unsigned char* tile = ring + b * C::TILE;
                PROF_BEGIN();
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t h0, l0, h1, l1;
                        split_f16x2(make_float2(buf[s][j][i].x, buf[s][j][i].y), h0, l0);
                        split_f16x2(make_float2(buf[s][j][i].z, buf[s][j][i].w), h1, l1);
                        unsigned char* dst = tile + soff[j] + 2 * i * C::SLOT;
                        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(dst + 1024) = make_uint2(l0, l1);
                    }
                if (lane < 2) {
                    // row info: score = info.x (Qn . x~), info.x = scale / max(|x~|, eps 2^-e), 2^e = info.y,
                    // info.z = dv . x / P (backward).  1 / |x~| = rsqrt + one Newton step.
                    const int ej = lane ? e[1] : e[0];
                    const float sj = lane ? ssq[1] : ssq[0];
                    const float sc = __uint_as_float(uint32_t(127 - ej) << 23);
                    const float st = sj * sc * sc;                             // |x~|^2, exact scaling
                    float4 info;
                    info.y = __uint_as_float(uint32_t(127 + ej) << 23);
                    float y = rsqrtf(st);
                    y = y * fmaf(-0.5f * st * y, y, 1.5f);
                    y = fminf(y, info.y * (1.f / VLSA_NORM_EPS));              // also catches ssq == 0 (NaN -> cap)
                    info.x = prm.scale * y;
                    info.z = BWD ? (lane ? ud[1] : ud[0]) : 0.f;
                    info.w = 0.f;
                    *reinterpret_cast<float4*>(s_rowinfo + (b * TR + 2 * pw + lane) * 4) = info;
                }
                PROF_END(2);
                // the register set is free: the tile R ahead goes in flight before this one is published
                VLSA_TC_ISSUE(s)
                __syncwarp();
                mbar_arrive_if(full + b, lane == 0);
                cur_next(cv);
                if (++b == uint32_t(C::NBUF)) { b = 0; par ^= 1u; }
            }
        }
#undef VLSA_TC_ISSUE
        PROF_FLUSH(0, 3, pw == 0 && lane == 0)
    } else if (warp >= C::NSOFT) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_ISSUE));
      if (VLSA_TC_PF > 0 && warp == C::W_G1 + 2) {
        // =========================================================================== L2 prefetcher (development option)
        // one bulk prefetch per tile, VLSA_TC_PF tiles ahead of the producers' loads: a producer's first use of a register
        // set waits for EVERY load its warp has in flight (ptxas puts them on one scoreboard), i.e. for a full memory
        // latency per rotation of its registers; loads that hit L2 shorten that wait
        if (lane == 0) {
            const float* Xf = reinterpret_cast<const float*>(prm.X);
            uint32_t done = 0;
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                for (long long r = r0; r < r1; r += TR, ++done) {
                    while (int(done) >= int(*s_prog) + VLSA_TC_PF) __nanosleep(64);
                    const long long n = r1 - r < TR ? r1 - r : TR;
                    l2_prefetch_bulk(Xf + r * D, uint32_t(n) * D * 4u);
                }
            }
        }
        __syncwarp();
      } else if (warp == C::W_G1) {
        // =========================================================================== GEMM1 issuer
        if (elect_one()) {
            constexpr uint32_t idesc1 = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * TR, false, false);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(ring), 16, 1024);
            const uint32_t tq0 = tmem + C::TM_Q;
            uint32_t tt = 0;
            PROF_DECL
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t b = tt % C::NBUF, u = tt / C::NBUF, par = tt & 1u, v = tt >> 1;
                    PROF_BEGIN();
                    mbar_wait_wd(s_free + par, (v & 1u) ^ 1u);         // scores of tile tt - 2 have been read
                    PROF_END(0);
                    PROF_BEGIN();
                    mbar_wait_wd(full + b, u & 1u);
                    PROF_END(1);
                    // the producers' plain shared-memory stores (generic proxy) become visible to the tensor core's
                    // operand reads (async proxy) through THIS fence: it sits on the causality path producers ->
                    // mbarrier `full` -> issuer -> tcgen05.mma.  The producers themselves cannot fence: a proxy fence
                    // waits for every global load its thread has in flight, i.e. for the tiles they are prefetching.
                    fence_proxy_async_smem();
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(desc0, b * C::TILE);
                    const uint32_t d1 = tmem + C::TM_D1 + 32 * par;
#pragma unroll
                    for (int s = 0; s < C::NSLOT; ++s) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // B = 32 rows: [g0 hi | g0 lo | g1 hi | g1 lo] of the slot
                            tc_mma_ts(d1, tq0 + (s * 4 + ks) * 8, umma_desc_advance(tb, s * C::SLOT + ks * 32), idesc1,
                                      (s | ks) != 0);
                    }
                    tc_commit(s_ready + par);
                }
            }
            PROF_FLUSH(6, 2, true)
        }
        __syncwarp();
      } else if (warp == C::W_G2) {
        // =========================================================================== GEMM2 issuer
        if (elect_one()) {
            constexpr uint32_t idesc_hi = umma_idesc(UMMA_F16, UMMA_F16, 128, 2 * NP, true, false);
            constexpr uint32_t idesc_lo = umma_idesc(UMMA_F16, UMMA_F16, 128, NP, true, false);
            const uint64_t a0 = umma_desc_sw128(smem_u32(ring), C::SLOT, C::GRP);   // M atoms: next slot; K atoms: next row group
            const uint64_t w0 = umma_desc_sw128(smem_u32(wt), 16, 1024);
            uint32_t tt = 0, cc = 0;
            PROF_DECL
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t i = tt & 1u, v = tt >> 1, b = tt % C::NBUF;
                    PROF_BEGIN();
                    mbar_wait_wd(w_ready + i, v & 1u);
                    PROF_END(0);
                    PROF_BEGIN();
                    if (t == 0) mbar_wait_wd(d2_free, (cc & 1u) ^ 1u);   // previous chunk's accumulators drained
                    PROF_END(1);
                    fence_proxy_async_smem();                            // as in the GEMM1 issuer: the tile buffer and the weight
                                                                         // operand the weight warps wrote with plain stores
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(a0, b * C::TILE), wb = umma_desc_advance(w0, i * C::WBUF);
                    const uint32_t acc0 = t != 0;
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const uint32_t d2 = tmem + C::TM_D2 + gg * C::D2W;
                        const uint64_t ah = umma_desc_advance(tb, (2 * gg) * C::SLOT);
                        tc_mma_ss(d2, ah, wb, idesc_hi, acc0);
                        tc_mma_ss(d2 + 2 * NP, umma_desc_advance(ah, 1024), wb, idesc_lo, acc0);
                    }
                    tc_commit(empty + b);
                    tc_commit(w_free + i);
                    if (t == ntiles - 1) tc_commit(d2_done);
                }
            }
            PROF_FLUSH(9, 2, true)
        }
        __syncwarp();
      }
    } else {
        // =========================================================================== weights / drain
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_SOFT));
        // Two sets of four warps (one per TMEM quadrant) take the tiles alternately: tile tt belongs to set tt & 1, which
        // also owns score buffer tt & 1 and weight buffer tt & 1 — two tiles are in this stage at any time.  A thread owns
        // (prototype p = 4 q + (lane & 3)) x (tile rows 2 rj, 2 rj + 1), so a warp sees all 16 rows of its four prototypes and
        // settles their softmax reference on its own.  What the sets share is the reference itself (s_mref): set s may
        // only decide tile tt after the other set has decided tile tt - 1 (mbarrier `decided`), and every thread folds a
        // reference it finds changed into its running sum before going on.
        const int q = warp & 3, set = warp >> 2;
        const int pl = lane & 3, p = 4 * q + pl, rj = lane >> 2;
        const bool pvalid = p < P;
        const uint32_t tq = tmem + (uint32_t(32 * q) << 16);
        constexpr int NT = C::NSOFT * 32, NTS = NT / 2;
        // B operand of GEMM2: row (term * 16 + p), K = tile row: rows 2 rj, 2 rj + 1 -> 16-byte chunk rj >> 2, bytes 4 (rj & 3)
        const uint32_t w_off = sw128_offset(p, rj >> 2, 4 * (rj & 3));
        uint32_t tt = 0, cc = 0;
        PROF_DECL
        for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
            int bag; long long r0, r1;
            chunk_info(prm, c, bag, r0, r1);
            const int chunk_nrows = int(r1 - r0);
            const int ntiles = (chunk_nrows + TR - 1) / TR;
            // forward: m_loc = softmax reference this thread's sum refers to, lsum = running sum, exE = chunk reference
            //          exponent E (accumulators hold 2^-E O)
            // backward: m_loc = log2 of the normaliser H_p (accumulators hold dQn_p / H_p)
            float m_loc = -INFINITY, lsum = 0.f;
            int exE = 127;
            float bw_m = 0.f, bw_il = 0.f, bw_delta = 0.f;
            if (BWD && pvalid) {
                bw_m = __ldg(prm.ml + (size_t(bag) * P + p) * 2);
                bw_il = 1.f / __ldg(prm.ml + (size_t(bag) * P + p) * 2 + 1);
                bw_delta = __ldg(prm.delta + size_t(bag) * P + p);
            }
            // own tiles of this chunk: t = t_first, t_first + 2, ...; tt0 = index of the chunk's first tile in the CTA's sequence
            const uint32_t tt0 = tt;
            const int t_first = int((uint32_t(set) ^ tt0) & 1u);
            uint32_t b = (tt0 + t_first) % C::NBUF, ph = ((tt0 + t_first) / C::NBUF) & 1u;
            tt = tt0 + uint32_t(ntiles);                           // for the next chunk
            for (int t = t_first; t < ntiles; t += 2, b += 2u) {
                if (b >= uint32_t(C::NBUF)) { b -= C::NBUF; ph ^= 1u; }
                const uint32_t tt = tt0 + uint32_t(t), v = tt >> 1;
                const int nvalid = min(TR, chunk_nrows - t * TR);
                PROF_BEGIN();
                mbar_wait_wd(s_ready + set, v & 1u);
                PROF_END(0);
                tc_fence_after();
                float sc2[2];
                {
                    // 32 partial scores of this lane's (prototype, part, range): [g0 hi | g0 lo | g1 hi | g1 lo] x 8 rows
                    uint32_t sa[32];
                    tmem_ld32(tq + C::TM_D1 + 32 * set, sa);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    mbar_arrive_if(s_free + set, lane == 0);
                    // add the planes; the 8 (part, range) partial sums of a (row, prototype) then sit in the lanes that differ
                    // in bits 2-4: a transposed butterfly adds them in a fixed order and halves the rows a lane keeps at every
                    // level (14 shuffles, no shared memory): lane bit 4 -> row bit 3, bit 3 -> row bit 2, bit 2 -> row bit 1
                    float v8[8], v4[4];
                    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float lo = __uint_as_float(sa[i]) + __uint_as_float(sa[8 + i]);            // rows 0 .. 7
                        const float hi = __uint_as_float(sa[16 + i]) + __uint_as_float(sa[24 + i]);      // rows 8 .. 15
                        v8[i] = (u16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, u16 ? lo : hi, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) v4[i] = (u8 ? v8[4 + i] : v8[i]) + __shfl_xor_sync(0xffffffffu, u8 ? v8[i] : v8[4 + i], 8);
#pragma unroll
                    for (int i = 0; i < 2; ++i) sc2[i] = (u4 ? v4[2 + i] : v4[i]) + __shfl_xor_sync(0xffffffffu, u4 ? v4[i] : v4[2 + i], 4);
                }
                PROF_BEGIN();
                mbar_wait_wd(full + b, ph);                            // acquire the converters' row info
                PROF_END(1);
                float4 info[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) info[k] = *reinterpret_cast<const float4*>(s_rowinfo + (b * TR + 2 * rj + k) * 4);
                // ---- in tile order from here: the other set has settled tile tt - 1
                PROF_BEGIN();
                if (tt > 0) mbar_wait_wd(decided + (set ^ 1), ((tt - 1) >> 1) & 1u);
                PROF_END(2);
                if (t == 0) {
                    exE = int(__float_as_uint(s_rowinfo[(b * TR) * 4 + 1]) >> 23);
                    if (warp == 4 * set && lane == 0) s_exE[0] = exE;
                } else {
                    exE = s_exE[0];
                    const float m_sh = s_mref[p];                      // a reference (normaliser) the other set has moved
                    if (pvalid && m_sh > m_loc) {
                        if (!BWD) lsum *= expf(m_loc - m_sh);
                        m_loc = m_sh;
                    }
                }
                float w[2];                                            // weights fed to GEMM2 (before the fp16 split)
                float ts[2], unscale[2], cw[2];
                bool grow;
                if (!BWD) {
                    // ts = score + (e_row - E) ln 2: the weight is exp(ts - m) = A-weight x 2^(e_row - E)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        int de = int(__float_as_uint(info[k].y) >> 23) - exE;
                        de = de < -100 ? -100 : (de > 100 ? 100 : de);
                        ts[k] = (2 * rj + k < nvalid) ? fmaf(float(de), 0.693147180559945f, sc2[k] * info[k].x) : -INFINITY;
                        unscale[k] = __uint_as_float(uint32_t(127 - de) << 23);       // 2^-(e_row - E)
                    }
                    grow = pvalid && (fmaxf(ts[0], ts[1]) > m_loc + C::MARGIN);        // true on the first tile
                } else {
                    // c = scale A (u - delta) / |x| = A (u - delta) info.x 2^-e; the weight on x~ = 2^-e x is c 2^e
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float a = expf(sc2[k] * info[k].x - bw_m) * bw_il;       // A_pn (deepmil.py:198)
                        cw[k] = (pvalid && 2 * rj + k < nvalid) ? a * (info[k].z - bw_delta) * info[k].x : 0.f;
                    }
                    // binary exponent of the larger |cw| (zero / denormal -> very small, non-finite -> very large)
                    int et = int((__float_as_uint(fmaxf(fabsf(cw[0]), fabsf(cw[1]))) >> 23) & 0xffu) - 127;
                    et = et < -100 ? -100 : (et > 100 ? 100 : et);
                    ts[0] = float(et);
                    grow = pvalid && (t == 0 || ts[0] > m_loc + float(C::BWD_MAXE));
                }
                PROF_BEGIN();
                const bool any_grow = named_bar_or(1 + set, NTS, grow);
                PROF_END(3);
                if (any_grow) {
                    // rare (always on the first tile of a chunk): the warp settles the new reference of its four prototypes
                    // from all 16 rows; from the second tile on the TMEM accumulators are rescaled once GEMM2 of the
                    // previous tile has completed (every warp of the set needs every prototype's factor: s_alpha)
                    float mt = BWD ? ts[0] : fmaxf(ts[0], ts[1]);
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
                    float m_new = m_loc;
                    if (!BWD) { if (pvalid && mt > m_loc + C::MARGIN) m_new = mt + C::HEADROOM; }
                    else { if (pvalid && (t == 0 || mt > m_loc + float(C::BWD_MAXE))) m_new = mt - float(C::BWD_SETE); }
                    if (t > 0) {
                        const float alpha = (pvalid && m_new > m_loc) ? (BWD ? exp2f(m_loc - m_new) : expf(m_loc - m_new)) : 1.f;
                        if (lane < 4) s_alpha[p] = alpha;
                        mbar_wait_wd(w_free + (set ^ 1), ((tt - 1) >> 1) & 1u);   // GEMM2 of tile tt - 1 has completed
                        tc_fence_after();
                        named_bar_sync(3 + set, NTS);
                        float al[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) al[j] = s_alpha[j];
#pragma unroll 1
                        for (int k = 0; k < 12; ++k) {
                            uint32_t o[16];
                            tmem_ld16(tq + C::TM_D2 + 16 * k, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * al[j]);
                            tmem_st16(tq + C::TM_D2 + 16 * k, o);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        if (!BWD) lsum *= alpha;
                        named_bar_sync(3 + set, NTS);                  // s_alpha may be rewritten by the next rare event
                    }
                    m_loc = m_new;
                    if (lane < 4) s_mref[p] = m_new;
                }
                __syncwarp();
                mbar_arrive_if(decided + set, lane == 0);             // releases s_mref / s_exE to the other set
                if (!BWD) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        w[k] = (pvalid && 2 * rj + k < nvalid) ? expf(ts[k] - m_loc) : 0.f;
                        lsum = fmaf(w[k], unscale[k], lsum);
                    }
                } else {
                    const float inv_h = pvalid ? __uint_as_float(uint32_t(127 - int(m_loc)) << 23) : 0.f;   // 1 / H_p
#pragma unroll
                    for (int k = 0; k < 2; ++k) w[k] = cw[k] * inv_h;
                }
                // weights as two fp16 terms (w = t0 + 2^-11 t1); B operand row (term * 16 + p), K = tile row (2 rj, 2 rj + 1)
                unsigned short b0[2], b1[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const __half h0 = __float2half_rn(w[k]);
                    const __half h1 = __float2half_rn((w[k] - __half2float(h0)) * 2048.f);
                    b0[k] = __half_as_ushort(h0); b1[k] = __half_as_ushort(h1);
                }
                PROF_BEGIN();
                mbar_wait_wd(w_free + set, (v & 1u) ^ 1u);             // GEMM2 of tile tt - 2 has read this buffer
                PROF_END(4);
                // tile rows 2 rj, 2 rj + 1 are neighbours along K: one 32-bit store per term
                unsigned char* wb = wt + set * C::WBUF + w_off;
                *reinterpret_cast<uint32_t*>(wb) = uint32_t(b0[0]) | (uint32_t(b0[1]) << 16);
                *reinterpret_cast<uint32_t*>(wb + NP * 128) = uint32_t(b1[0]) | (uint32_t(b1[1]) << 16);
                __syncwarp();                                          // (the GEMM2 issuer fences for the async proxy)
                mbar_arrive_if(w_ready + set, lane == 0);
            }
            // ---- chunk end: both sets meet, agree on the final reference, write (m, l), drain O^T
            named_bar_sync(7, NT);
            {
                const float m_sh = s_mref[p];
                if (pvalid && m_sh > m_loc) {
                    if (!BWD) lsum *= expf(m_loc - m_sh);
                    m_loc = m_sh;
                }
                exE = s_exE[0];
            }
            if (BWD) {
                if (set == 0 && lane < 4) s_alpha[p] = pvalid ? __uint_as_float(uint32_t(127 + int(m_loc)) << 23) : 0.f;
            } else {
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 4);
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
                lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
                if (lane < 4) s_lsum[16 * set + p] = lsum;
            }
            named_bar_sync(7, NT);
            if (!BWD && set == 0 && lane < 4 && pvalid) {
                prm.part_l[size_t(c) * P + p] = s_lsum[p] + s_lsum[16 + p];
                prm.part_m[size_t(c) * P + p] = m_loc;
            }
            PROF_BEGIN();
            mbar_wait_wd(d2_done, cc & 1u);
            PROF_END(5);
            tc_fence_after();
            float* po = prm.part_O + size_t(c) * P * D;
            float mul[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) mul[j] = BWD ? s_alpha[j] : __uint_as_float(uint32_t(exE) << 23);
#pragma unroll 1
            for (int gg = 2 * set; gg < 2 * set + 2; ++gg) {
                // lane = feature within a 128-block, columns = hi.t0 | hi.t1 | lo.t0 per prototype
                uint32_t o0[16], o1[16], o2[16];
                tmem_ld16(tq + C::TM_D2 + gg * C::D2W, o0);
                tmem_ld16(tq + C::TM_D2 + gg * C::D2W + 16, o1);
                tmem_ld16(tq + C::TM_D2 + gg * C::D2W + 32, o2);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < P) po[size_t(j) * D + 128 * gg + 32 * q + lane] =
                        mul[j] * (fmaf(__uint_as_float(o1[j]), 0x1p-11f, __uint_as_float(o2[j])) + __uint_as_float(o0[j]));
            }
            tc_fence_before();
            __syncwarp();
            mbar_arrive_if(d2_free, lane == 0);
            named_bar_sync(7, NT);                                     // s_alpha / s_lsum / s_mref / s_exE are free again
        }
        PROF_FLUSH(12, 6, warp == 0 && lane == 0)
#ifdef VLSA_TMA_PROF
        if (blockIdx.x == 0 && warp == 0 && lane == 0) g_tma_prof[19] = tt;
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::W_G1) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace vlsa
