// Fused language-guided aggregation on tcgen05 for rows stored as PRE-SPLIT TILE IMAGES (VLSA_DTYPE_SPLIT16), forward and
// backward, every P <= 16.  The fp16 (hi, lo) planes the tensor core multiplies for fp32 rows (agg_tc.cuh) cost that
// kernel its HBM roofline: the split needs the values in registers, fed from shared memory (two more passes) or from
// global memory (one scoreboard).  A DEVICE-RESIDENT COHORT does not have to pay that every epoch: its bags are uploaded
// once per training run (dataset/cohort.py; the reference re-uploads every bag every epoch, runner/vlsa_handler.py:205),
// and `vlsa_split16_pack` stores them at upload time as the exact shared-memory image of agg_tc_kernel's tile buffer:
//
//   record of 16 rows (32 896 B): 8 slots (64 features) x 2 row groups x (hi atom 1 KB | lo atom 1 KB), 128-byte swizzle,
//                                 then 16 x (1 / max(|x~|, eps 2^-e), 2^e) — the same 4 bytes per element as fp32 rows
//
// The pass is then what it is for bf16 rows (agg_bf16.cuh): ONE bulk copy per tile (cp.async.bulk, SASS UBLKCP) lands an
// operand-ready tile, GEMM1 starts on the TMA's mbarrier, nobody converts anything; a byte of X crosses shared memory three
// times (TMA write + two operand reads; backward: + one LDS pass of the norm warps for u = dv . x / P).  Same planes, same
// Qn layout, same weight terms, same summation orders as agg_tc_kernel: the results are bit-identical to that kernel's on
// the fp32 rows the cohort was packed from (tests/test_gpu_cohort.py).
//
//   GEMM1  S^T[128, 32]  = Qn'[128, 512] (A, TMEM) . [X_hi ; X_lo][32, 512]^T (B, K-major smem)
//   GEMM2  O^T[512 d, 32 | 16] += X_hi^T | X_lo^T [512, 16] (A, MN-major, the SAME smem bytes) . W[32 | 16, 16]^T
//
// Bags start at record boundaries (rows 16 k of the padded row space the plan's ranges live in), chunks are multiples of
// 32 rows, so a tile of a chunk is always one whole record.
//
// Warp roles (20 warps, 1 persistent CTA / SM): warps 0-7 weights (two alternating sets of four), warp 8 GEMM1 issuer +
// TMEM allocation, warp 9 GEMM2 issuer, warp 10 TMA issuer, warp 11 idle, warps 12-19 norm warps (backward only: two
// alternating sets of four, u = dv . x / P from the landed planes; they exit at once in the forward).
#pragma once
#include "agg_simt.cuh"
#include "tc_common.cuh"

namespace vlsa {

struct SplitCfg {
    static constexpr int D = VLSA_D;
    static constexpr int NP = 16;
    static constexpr int TR = 16;                 // rows per tile = rows per record
    static constexpr int NSLOT = 8;
    static constexpr int GRP = 2048;              // bytes of one 8-row group of a slot (hi | lo atoms)
    static constexpr int SLOT = 2 * GRP;          // 4 KB
    static constexpr int TILE = NSLOT * SLOT;     // 32 KB of planes
    static constexpr int META = TR * 8;           // 16 x float2
    static constexpr int REC = TILE + META;       // 32 896 B per record in HBM (2 056 B per row)
    static constexpr int ROW_BYTES = REC / TR;
    static constexpr int BUF = TILE + 1024;       // shared-memory stride of a tile buffer (atoms need 1 KB alignment)
    static constexpr int NBUF = 6;
    static constexpr int WBUF = 2 * NP * 128;
    static constexpr int OFF_W = NBUF * BUF;
    static constexpr int OFF_F = OFF_W + 2 * WBUF;
    // floats: u[NBUF][TR] | alpha[16] | mref[16] | lsum[2][16] | exE[4]
    static constexpr int NFLOAT = NBUF * TR + 16 + 16 + 32 + 4;
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 3 * NBUF + 12;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int NSOFT = 8, NCONV = 8;
    static constexpr int W_G1 = NSOFT, W_G2 = NSOFT + 1, W_TMA = NSOFT + 2, W_CONV = NSOFT + 4;
    static constexpr int NWARPS = NSOFT + 4 + NCONV;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int REG_SOFT = 96;           // no setmaxnreg needed: every role fits the launch bound
    static constexpr int QPITCH = D + 1;
    static constexpr int TM_Q = 0;
    static constexpr int D2W = 3 * NP;
    static constexpr int TM_D2 = 256;
    static constexpr int TM_D1 = TM_D2 + 4 * D2W;   // 448: two score buffers of 32 columns
    static constexpr int TMEM_COLS = 512;
    static constexpr float HEADROOM = 6.f, MARGIN = 10.f;
    static constexpr int BWD_MAXE = 14, BWD_SETE = 6;
};

// Pack rows [0, n_rows) of X fp32 [n_rows, 512] into the records starting at padded row `first_row` (a multiple of 16) of
// `image`: one CTA per record, one warp per row.  The arithmetic (row norm by FFMA2 chain + five-level butterfly, power-of-
// two row scale, fp16 split) is agg_tc_kernel's producer code, so the planes are bit-identical to what that kernel computes.
__global__ void __launch_bounds__(512) split16_pack_kernel(const float* __restrict__ X, long long n_rows, unsigned char* __restrict__ image,
                                                          long long first_row) {
    using C = SplitCfg;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;           // w = row of the record
    const long long row = (long long)blockIdx.x * C::TR + w;
    unsigned char* rec = image + (first_row / C::TR + blockIdx.x) * (long long)C::REC;
    float4 x4[4];
    const bool live = row < n_rows;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        x4[i] = live ? __ldg(reinterpret_cast<const float4*>(X + row * C::D + 128 * i + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a2 = __ffma2_rn(make_float2(x4[i].x, x4[i].y), make_float2(x4[i].x, x4[i].y), a2);
        a2 = __ffma2_rn(make_float2(x4[i].z, x4[i].w), make_float2(x4[i].z, x4[i].w), a2);
    }
    float ssq = a2.x + a2.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    uint32_t ex = __float_as_uint(ssq) >> 23;
    if (ex == 0u || ex >= 255u) ex = 127u;
    int e = (int(ex) - 127) >> 1;
    if (e >= 1 && e <= 13) e = 0;
    const float sc = __uint_as_float(uint32_t(127 - e) << 23);
    if (e != 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { x4[i].x *= sc; x4[i].y *= sc; x4[i].z *= sc; x4[i].w *= sc; }
    }
    const uint32_t soff = (lane >> 4) * C::SLOT + (w >> 3) * C::GRP + sw128_offset(w & 7, (lane & 15) >> 1, (lane & 1) * 8);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t h0, l0, h1, l1;
        split_f16x2(make_float2(x4[i].x, x4[i].y), h0, l0);
        split_f16x2(make_float2(x4[i].z, x4[i].w), h1, l1);
        unsigned char* dst = rec + soff + 2 * i * C::SLOT;
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(dst + 1024) = make_uint2(l0, l1);
    }
    if (lane == 0) {
        const float st = ssq * sc * sc;
        const float p2e = __uint_as_float(uint32_t(127 + e) << 23);
        float y = rsqrtf(st);
        y = y * fmaf(-0.5f * st * y, y, 1.5f);
        y = fminf(y, p2e * (1.f / VLSA_NORM_EPS));
        reinterpret_cast<float2*>(rec + C::TILE)[w] = live ? make_float2(y, p2e) : make_float2(0.f, 1.f);
    }
}

template <bool BWD>
__global__ void __launch_bounds__(SplitCfg::THREADS, 1) agg_split_kernel(const AggParams prm, const int P) {
    using C = SplitCfg;
    constexpr int D = C::D, NP = C::NP, TR = C::TR;
    if (int(blockIdx.x) >= prm.total_chunks) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* ring = sm;
    unsigned char* wt = sm + C::OFF_W;
    float* s_u = reinterpret_cast<float*>(sm + C::OFF_F);            // [NBUF][TR] dv . x / P per row (backward)
    float* s_alpha = s_u + C::NBUF * TR;                             // [16] rescale factors (rare path) / drain normalisers
    float* s_mref = s_alpha + 16;                                    // [16] current softmax reference (fwd) | log2 H_p (bwd)
    float* s_lsum = s_mref + 16;                                     // [2][16] per-set softmax sums at a chunk end
    int* s_exE = reinterpret_cast<int*>(s_lsum + 32);                // [4] reference row-scale exponent of the chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* landed = bars;                     // [NBUF] TMA (expect_tx)           -> GEMM1, weight warps (metadata), norm warps
    uint64_t* full = bars + C::NBUF;             // [NBUF] norm warps (4, backward)  -> weight warps (u)
    uint64_t* empty = bars + 2 * C::NBUF;        // [NBUF] GEMM2 commit              -> TMA issuer
    uint64_t* s_ready = bars + 3 * C::NBUF;      // [2]    GEMM1 commit              -> weight warps
    uint64_t* s_free = s_ready + 2;              // [2]    weight set s (4 warps)    -> GEMM1
    uint64_t* w_ready = s_ready + 4;             // [2]    weight set s (4 warps)    -> GEMM2
    uint64_t* w_free = s_ready + 6;              // [2]    GEMM2 commit              -> weight warps
    uint64_t* d2_done = s_ready + 8;             //        last GEMM2 of a chunk     -> weight warps (drain)
    uint64_t* d2_free = s_ready + 9;             //        weight warps (8)          -> GEMM2 of the next chunk
    uint64_t* decided = s_ready + 10;            // [2]    weight set s (4 warps): softmax reference settled for its tile -> other set
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned char* image = reinterpret_cast<const unsigned char*>(prm.X);

    if (tid == 0) {
        for (int s = 0; s < C::NBUF; ++s) { mbar_init(landed + s, 1); mbar_init(full + s, 4); mbar_init(empty + s, 1); }
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
    // the async proxy (TMA) writes the ring next: order the generic-proxy staging reads / writes before it
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= C::W_CONV) {
        // =========================================================================== norm warps (backward: u = dv . x / P)
        // Two sets of four warps take the tiles alternately.  Warp cw4 of a set: tile rows 4 cw4 + 2 h (lanes 0-15) and
        // 4 cw4 + 2 h + 1 (lanes 16-31), h = 0, 1; a lane reads the 16-byte chunk at PHYSICAL position (lane & 7) ^ 2 h of
        // its row in the hi and the lo atom of slots 2 it + ((lane >> 3) & 1), it = 0 .. 3 (conflict-free LDS.128: the
        // eight lanes of a quarter warp cover one 128-byte swizzled row segment); the LOGICAL chunk = position ^ (row & 7)
        // is the same for both h, so one set of dv registers serves both rows.
        if (BWD) {
            const int cw = warp - C::W_CONV, nset = cw >> 2, cw4 = cw & 3;
            const int row = 4 * cw4 + (lane >> 4), sub = (lane >> 3) & 1, pos = lane & 7;     // row of h = 0 (h = 1: + 2)
            const int chunk = pos ^ (row & 7);
            const uint32_t ld_off = sub * C::SLOT + (row >> 3) * C::GRP + (row & 7) * 128;
            float dvr[BWD ? 32 : 1];                           // dv / P at this lane's 32 features
            int dv_bag = -1;
            uint32_t tt = 0;
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    if (int(tt & 1u) != nset) continue;
                    if (bag != dv_bag) {
                        dv_bag = bag;
                        const float invP = 1.f / float(P);
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const float* src = prm.dv + size_t(bag) * D + (2 * it + sub) * 64 + chunk * 8;
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(src)), b4 = __ldg(reinterpret_cast<const float4*>(src + 4));
                            float* d8 = dvr + (BWD ? 8 * it : 0);
                            d8[0] = a4.x * invP; d8[1] = a4.y * invP; d8[2] = a4.z * invP; d8[3] = a4.w * invP;
                            d8[4] = b4.x * invP; d8[5] = b4.y * invP; d8[6] = b4.z * invP; d8[7] = b4.w * invP;
                        }
                    }
                    const uint32_t b = tt % C::NBUF, ph = (tt / C::NBUF) & 1u;
                    const unsigned char* tile = ring + b * C::BUF + ld_off;
                    mbar_wait_wd(landed + b, ph);
                    float uu[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float2 u2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const unsigned char* src = tile + h * 256 + ((pos ^ (2 * h)) << 4) + 2 * it * C::SLOT;
                            const uint4 hi4 = *reinterpret_cast<const uint4*>(src), lo4 = *reinterpret_cast<const uint4*>(src + 1024);
                            const uint32_t hw[4] = {hi4.x, hi4.y, hi4.z, hi4.w}, lw[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 xh = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
                                const float2 xl = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
                                const float2 d2 = make_float2(dvr[BWD ? 8 * it + 2 * k : 0], dvr[BWD ? 8 * it + 2 * k + 1 : 0]);
                                u2[0] = __ffma2_rn(xh, d2, u2[0]);
                                u2[1] = __ffma2_rn(xl, d2, u2[1]);
                            }
                        }
                        uu[h] = (u2[0].x + u2[0].y) + (u2[1].x + u2[1].y);
                    }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                        for (int h = 0; h < 2; ++h) uu[h] += __shfl_xor_sync(0xffffffffu, uu[h], o);
                    if ((lane & 7) == 0) {
                        // lanes 0 / 16: rows of h = 0, lanes 8 / 24: rows of h = 1; the planes hold x~ = 2^-e x
                        const int h = (lane >> 3) & 1, r = row + 2 * h;
                        const float p2e = reinterpret_cast<const float2*>(ring + b * C::BUF + C::TILE)[r].y;
                        s_u[b * TR + r] = (h ? uu[1] : uu[0]) * p2e;
                    }
                    __syncwarp();
                    mbar_arrive_if(full + b, lane == 0);
                }
            }
        }
    } else if (warp >= C::NSOFT) {
      if (warp == C::W_TMA) {
        // =========================================================================== TMA issuer
        if (elect_one()) {
            const uint64_t policy = make_evict_first_policy();
            uint32_t tt = 0;
            PROF_DECL
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t b = tt % C::NBUF, u = tt / C::NBUF;
                    PROF_BEGIN();
                    mbar_wait_wd(empty + b, (u & 1u) ^ 1u);            // GEMM2 of the tile that used this buffer is done
                    PROF_END(0);
                    mbar_expect_tx(landed + b, C::REC);
                    bulk_g2s_evict_first(ring + b * C::BUF, image + (r0 / TR + t) * (long long)C::REC, C::REC, landed + b, policy);
                }
            }
            PROF_FLUSH(4, 1, true)
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
                    mbar_wait_wd(landed + b, u & 1u);             // operand-ready as it lands (TMA -> tensor core: one proxy)
                    PROF_END(1);
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(desc0, b * C::BUF);
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
                    fence_proxy_async_smem();                            // the weight operand was written with plain stores
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(a0, b * C::BUF), wb = umma_desc_advance(w0, i * C::WBUF);
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
                mbar_wait_wd(landed + b, ph);                          // acquire the record's row metadata (written by the TMA)
                if (BWD) mbar_wait_wd(full + b, ph);                   // ... and u = dv . x / P from the norm warps
                PROF_END(1);
                // metadata of a row: (1 / max(|x~|, eps 2^-e), 2^e), written once when the cohort was packed
                const float2* meta = reinterpret_cast<const float2*>(ring + b * C::BUF + C::TILE);
                float4 info[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float2 m2 = meta[2 * rj + k];
                    info[k] = make_float4(prm.scale * m2.x, m2.y, BWD ? s_u[b * TR + 2 * rj + k] : 0.f, 0.f);
                }
                // ---- in tile order from here: the other set has settled tile tt - 1
                PROF_BEGIN();
                if (tt > 0) mbar_wait_wd(decided + (set ^ 1), ((tt - 1) >> 1) & 1u);
                PROF_END(2);
                if (t == 0) {
                    exE = int(__float_as_uint(meta[0].y) >> 23);
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
