// Fused language-guided aggregation for bf16-STORED rows on the 5th-generation tensor cores (tcgen05 + TMEM), forward and
// backward, every P <= 16 (BASELINE configs[4]: bf16 storage halves the bytes of the pass; the reference holds fp32).
// Same contract as agg_simt_kernel<P,BWD,__nv_bfloat16>: ONE read of X, per-chunk partials.
//
//   GEMM1  S^T[128, 16]  = Qn'[128, 512] (A, resident in TMEM, bf16) . X[16, 512]^T (B, K-major smem, bf16 AS STORED)
//   GEMM2  O^T[512 d, 32] += X^T [512, 16] (A, MN-major, the SAME smem bytes) . W[32, 16]^T (two bf16 weight terms)
//
// A bf16 row needs no conversion: the TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) lands a box of 16 rows x 64 columns as
// the two 8-row K-major atoms the tensor core reads, so a byte of X crosses shared memory four times at HALF the bytes of
// the fp32 kernels: TMA write, one LDS pass for the row norms (fp32 FMAs on the unpacked values; backward: also
// u = dv . x / P), and the two operand reads.  The values the tensor core multiplies are exactly the stored ones; the
// bf16 terms of Qn (three: 24 significant bits, one accumulator per term and 256-feature range) and of the weights (two)
// carry the rest of the precision, accumulation is fp32 in TMEM.  bf16 has the exponent range of fp32: no row
// scaling, no weight-term scaling (the lazily rescaled softmax reference is still what keeps the ACCUMULATORS in range).
//
// Tile = 16 rows = 16 KB: slot s (64 features) at s * 2048; row group g (8 rows) at + g * 1024 (128-byte swizzle).  Ring
// of 8 tiles (128 KB), all of it available as TMA prefetch depth.
//
// Warp roles (28 warps, 1 persistent CTA / SM, static round-robin over chunks).  With 16 KB per tile the HBM time of a
// tile is ~730 cycles, shorter than the latency chain of any role (mbarrier waits, TMEM / shared-memory round trips), so
// every stage runs in several alternating sets:
//   warps 0-15  weights: FOUR sets of four (one warp per TMEM quadrant), tile tt -> set tt % 4 (own score / weight buffer);
//               the 8 partial sums of a score are added by a transposed shuffle butterfly (no shared memory)
//   warp  16    GEMM1 issuer + TMEM allocation      warp 17  GEMM2 issuer      warp 18  TMA issuer      warp 19  idle
//   warps 20-27 row norms: TWO sets of four, tile tt -> set tt & 1; warp: four rows of its tiles
// Registers (setmaxnreg): weights 80, issuers 24, norm warps 80 of the 72 x 896 the launch bound grants.
#pragma once
#include "agg_simt.cuh"
#include "tc_common.cuh"

namespace vlsa {

#ifndef VLSA_BF16_TR
#define VLSA_BF16_TR 32          // rows per tile: 16 or 32
#endif
#ifndef VLSA_BF16_NSET
#define VLSA_BF16_NSET (VLSA_BF16_TR == 16 ? 4 : 2)
#endif

struct Bf16Cfg {
    static constexpr int D = VLSA_D;
    static constexpr int NP = 16;
    static constexpr int TR = VLSA_BF16_TR;       // rows per tile
    static constexpr int RPT = TR / 8;            // tile rows per weight thread
    static constexpr int NSLOT = 8;               // 64-feature slots
    static constexpr int GRP = 1024;              // bytes of one 8-row group of a slot (one swizzled atom)
    static constexpr int SLOT = (TR / 8) * GRP;   // 2 KB | 4 KB
    static constexpr int TILE = NSLOT * SLOT;     // 16 KB | 32 KB
#ifndef VLSA_BF16_NBUF
#define VLSA_BF16_NBUF (VLSA_BF16_TR == 16 ? 8 : 6)
#endif
    static constexpr int NBUF = VLSA_BF16_NBUF;
    static constexpr int NSET = VLSA_BF16_NSET;   // sets of four weight warps; tile tt belongs to set tt % NSET
    static constexpr int NSOFT = 4 * NSET, NCONV = 8;   // norm warps: two sets of four (tile tt: set tt & 1)
    // register budget (setmaxnreg, per warpgroup): the launch bound grants 72 to each of the 896 threads
    static constexpr int REG_SOFT = 80, REG_ISSUE = 24, REG_NORM = 80;
    static_assert(NSOFT * 32 * REG_SOFT + 128 * REG_ISSUE + NCONV * 32 * REG_NORM <= (NSOFT + 4 + NCONV) * 32 * 72, "register pool");
    static constexpr int WBUF = 2 * NP * 128;     // weight operand: 32 rows (term, prototype) x 128 B (32 B used)
    static constexpr int OFF_W = NBUF * TILE;
    static constexpr int OFF_F = OFF_W + NSET * WBUF;
    // floats: rowinfo[NBUF][TR][4] | alpha[16] | mref[16] | lsum[NSET][16] | exE[4]
    static constexpr int NFLOAT = NBUF * TR * 4 + 16 + 16 + NSET * 16 + 4;
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 3 * NBUF + 5 * NSET + 2;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int W_G1 = NSOFT, W_G2 = NSOFT + 1, W_TMA = NSOFT + 2, W_CONV = NSOFT + 4;
    static constexpr int NWARPS = NSOFT + 4 + NCONV;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int QPITCH = D + 1;
    static constexpr int TM_Q = 0;
    static constexpr int D2W = 2 * NP;            // per 128-feature block: t0 (16 prototypes) | t1
    static constexpr int TM_D2 = 256;
    static constexpr int TM_D1 = TM_D2 + 4 * D2W;   // 384: NSET score buffers of 16 columns
    // NBUF is a multiple of the number of norm-warp sets (2) and of NSET: a buffer is always served by the SAME set, whose warps
    // take their tiles in order — with an odd ring a set could reach a buffer's next phase before the other set had seen the
    // current one land, and an mbarrier parity wait that is a phase ahead returns at once (measured: an intermittent hang)
    static_assert((TR == 16 || TR == 32) && TM_D1 + NSET * TR <= 512 && NBUF % NSET == 0 && NBUF % 2 == 0 && 4 % NSET == 0,
                  "TMEM columns / ring");
    static constexpr int TMEM_COLS = 512;
    static constexpr float HEADROOM = 6.f, MARGIN = 10.f;
    static constexpr int BWD_MAXE = 14, BWD_SETE = 6;
};

template <bool BWD>
__global__ void __launch_bounds__(Bf16Cfg::THREADS, 1) agg_bf16_kernel(const AggParams prm, const int P,
                                                                       const __grid_constant__ CUtensorMap tmap) {
    using C = Bf16Cfg;
    constexpr int D = C::D, NP = C::NP, TR = C::TR;
    if (int(blockIdx.x) >= prm.total_chunks) return;
#ifdef VLSA_TMA_PROF
    unsigned long long prof_gt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_gt0));
#endif

    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* ring = sm;
    unsigned char* wt = sm + C::OFF_W;
    float* s_rowinfo = reinterpret_cast<float*>(sm + C::OFF_F);      // [NBUF][TR] x (score factor, 1, u, -)
    float* s_alpha = s_rowinfo + C::NBUF * TR * 4;                   // [16] rescale factors (rare path) / drain normalisers
    float* s_mref = s_alpha + 16;                                    // [16] current softmax reference (fwd) | log2 H_p (bwd)
    float* s_lsum = s_mref + 16;                                     // [NSET][16] per-set softmax sums at a chunk end
    int* s_exE = reinterpret_cast<int*>(s_lsum + C::NSET * 16);                // [4] reference row-scale exponent of the chunk (127 here)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* landed = bars;                     // [NBUF] TMA (expect_tx)           -> GEMM1, norm warps
    uint64_t* full = bars + C::NBUF;             // [NBUF] norm warps (8)            -> weight warps (row info)
    uint64_t* empty = bars + 2 * C::NBUF;        // [NBUF] GEMM2 commit              -> TMA issuer
    uint64_t* s_ready = bars + 3 * C::NBUF;      // [NSET] GEMM1 commit              -> weight warps
    uint64_t* s_free = s_ready + C::NSET;        // [NSET] weight set s (4 warps)    -> GEMM1
    uint64_t* w_ready = s_ready + 2 * C::NSET;   // [NSET] weight set s (4 warps)    -> GEMM2
    uint64_t* w_free = s_ready + 3 * C::NSET;    // [NSET] GEMM2 commit              -> weight warps
    uint64_t* decided = s_ready + 4 * C::NSET;   // [NSET] weight set s (4 warps): softmax reference settled for its tile -> next set
    uint64_t* d2_done = s_ready + 5 * C::NSET;   //        last GEMM2 of a chunk     -> weight warps (drain)
    uint64_t* d2_free = d2_done + 1;             //        weight warps (all)        -> GEMM2 of the next chunk
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for ptxas

#ifdef VLSA_WD_DEBUG
    if (tid == 0 && blockIdx.x == 0) g_wd[1] = smem_u32(bars);
#endif
    if (tid == 0) {
        for (int s = 0; s < C::NBUF; ++s) { mbar_init(landed + s, 1); mbar_init(full + s, 4); mbar_init(empty + s, 1); }
        for (int s = 0; s < C::NSET; ++s) {
            mbar_init(s_ready + s, 1); mbar_init(s_free + s, 4);
            mbar_init(w_ready + s, 4); mbar_init(w_free + s, 1);
            mbar_init(decided + s, 4);
        }
        mbar_init(d2_done, 1); mbar_init(d2_free, C::NSOFT);
        mbar_fence_init();
    }
    if (warp == C::W_G1) tmem_alloc(tmem_ptr, C::TMEM_COLS);
#ifndef VLSA_NO_TMAP_PREFETCH
    if (warp == C::W_TMA && lane == 0) tma_prefetch_desc(&tmap);
#endif
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
#ifdef VLSA_TMA_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_tma_prof[22] = (long long)(t1 - prof_gt0);          // ns until Qn is staged in shared memory and TMEM is allocated
    }
#endif
    const uint32_t tmem = *tmem_ptr;
    if (warp < 4) {
        // TMEM lane 32 warp + lane: prototype 4 warp + (lane & 3); c = lane >> 2: bf16 term c % 3 of Qn (three terms: 24
        // significant bits) restricted to the feature range 256 (c / 3) .. + 255 for c < 6, a zero row for c = 6, 7
        const float* qrow = reinterpret_cast<const float*>(ring) + (4 * warp + (lane & 3)) * C::QPITCH;
        const int c8 = lane >> 2, term = c8 % 3, range = c8 / 3;
        const uint32_t tq = tmem + (uint32_t(32 * warp) << 16) + C::TM_Q;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
            uint32_t v[32];
            const bool mine = c8 < 6 && (cb >> 2) == range;
            // per-lane masks instead of a per-element ternary: the lanes of a warp differ in (mine, term), and the compiler turned
            // the nested select into a divergent branch with a reconvergence point PER ELEMENT (256 of them: 20 us of every
            // launch, measured with the prologue timestamps of the -DVLSA_TMA_PROF build)
            const uint32_t m0 = (mine && term == 0) ? 0xffffffffu : 0u, m1 = (mine && term == 1) ? 0xffffffffu : 0u,
                           m2 = (mine && term == 2) ? 0xffffffffu : 0u;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                uint32_t t0, t1, t2;
                split_bf16x3(qrow[cb * 64 + 2 * i], qrow[cb * 64 + 2 * i + 1], t0, t1, t2);
                v[i] = (t0 & m0) | (t1 & m1) | (t2 & m2);
            }
            tmem_st32(tq + 32 * cb, v);
        }
        tmem_wait_st();
#ifdef VLSA_TMA_PROF
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            g_tma_prof[23] = (long long)(t1 - prof_gt0);      // ns until warp 0 has staged its quadrant of Qn in TMEM
        }
#endif
    }
    // the async proxy (TMA) writes the ring next: order the generic-proxy staging reads / writes before it
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#ifdef VLSA_TMA_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_tma_prof[21] = (long long)(t1 - prof_gt0);          // ns spent in the prologue
    }
#endif

    if (warp >= C::W_CONV) {
        // =========================================================================== row norms
        // Two sets of four warps take the tiles alternately (set = tile index & 1): the per-tile work is a latency chain
        // (mbarrier -> LDS -> FMA chain -> shuffles -> STS -> mbarrier) longer than the HBM time of a 16 KB tile.
        // Warp cw4 of a set: tile rows (TR / 4) cw4 + 2 h (lanes 0-15) and + 2 h + 1 (lanes 16-31), h = 0 .. TR / 8 - 1.  A lane
        // reads the 16-byte chunk at PHYSICAL position (lane & 7) ^ 2 h of its row in slots 2 it + ((lane >> 3) & 1), it =
        // 0 .. 3: the eight lanes of a quarter warp cover one 128-byte swizzled row segment (conflict-free LDS.128), and the
        // LOGICAL chunk = position ^ (row & 7) is the same for every h (the rows of a warp lie in one 8-row group and differ
        // by 2 h in row & 7), so one set of dv registers serves all of them in the backward.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_NORM));
        const int cw = warp - C::W_CONV, nset = cw >> 2, cw4 = cw & 3;
        constexpr int NH = TR / 8;                                                 // row pairs of a warp per tile
        const int row = (TR / 4) * cw4 + (lane >> 4), sub = (lane >> 3) & 1, pos = lane & 7;   // row of h = 0 (h: + 2 h)
        const int chunk = pos ^ (row & 7);
        const uint32_t ld_off = sub * C::SLOT + (row >> 3) * C::GRP + (row & 7) * 128;
        float dvr[BWD ? 32 : 1];                               // dv / P at this lane's 32 features (backward)
        int dv_bag = -1;
        uint32_t tt = 0;
        PROF_DECL
        for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
            int bag; long long r0, r1;
            chunk_info(prm, c, bag, r0, r1);
            const int ntiles = int((r1 - r0 + TR - 1) / TR);
            for (int t = 0; t < ntiles; ++t, ++tt) {
                if (int(tt & 1u) != nset) continue;
                if (BWD && bag != dv_bag) {
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
                const unsigned char* tile = ring + b * C::TILE + ld_off;
                PROF_BEGIN();
                mbar_wait_wd(landed + b, ph);
                PROF_END(0);
                float ss[NH], uu[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    uint4 raw[4];
#pragma unroll
                    for (int it = 0; it < 4; ++it)
                        raw[it] = *reinterpret_cast<const uint4*>(tile + h * 256 + ((pos ^ (2 * h)) << 4) + 2 * it * C::SLOT);
                    float2 a2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, u2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const uint32_t w4[4] = {raw[it].x, raw[it].y, raw[it].z, raw[it].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 xy = make_float2(bf16_lo(w4[k]), bf16_hi(w4[k]));
                            a2[k & 1] = __ffma2_rn(xy, xy, a2[k & 1]);
                            if (BWD) u2[k & 1] = __ffma2_rn(xy, make_float2(dvr[BWD ? 8 * it + 2 * k : 0], dvr[BWD ? 8 * it + 2 * k + 1 : 0]), u2[k & 1]);
                        }
                    }
                    ss[h] = (a2[0].x + a2[0].y) + (a2[1].x + a2[1].y); uu[h] = (u2[0].x + u2[0].y) + (u2[1].x + u2[1].y);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        ss[h] += __shfl_xor_sync(0xffffffffu, ss[h], o);
                        if (BWD) uu[h] += __shfl_xor_sync(0xffffffffu, uu[h], o);
                    }
                if ((lane & 15) < NH) {
                    // lane h of each half warp writes the row of pair h.  Row info: score = info.x (Qn . x), info.x = scale /
                    // max(|x|, eps); info.y = 1 (no row scaling); info.z = dv . x / P (backward).
                    // 1 / |x| = rsqrt + one Newton step.
                    const int h = lane & 15;
                    float sq = ss[0], uq = uu[0];
#pragma unroll
                    for (int i = 1; i < NH; ++i) { sq = h == i ? ss[i] : sq; uq = h == i ? uu[i] : uq; }
                    float4 info;
                    float y = rsqrtf(sq);
                    y = y * fmaf(-0.5f * sq * y, y, 1.5f);
                    y = fminf(y, 1.f / VLSA_NORM_EPS);                         // also catches ss == 0 (NaN -> cap)
                    info.x = prm.scale * y;
                    info.y = 1.f;
                    info.z = BWD ? uq : 0.f;
                    info.w = 0.f;
                    *reinterpret_cast<float4*>(s_rowinfo + (b * TR + row + 2 * h) * 4) = info;
                }
                __syncwarp();
                mbar_arrive_if(full + b, lane == 0);
            }
        }
        PROF_FLUSH(0, 3, cw == 0 && lane == 0)
    } else if (warp >= C::NSOFT) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_ISSUE));
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
                    mbar_expect_tx(landed + b, C::TILE);
                    unsigned char* dst = ring + b * C::TILE;
                    const int row = int(r0) + t * TR;                  // rows past the packed X arrive as zeros
#pragma unroll
                    for (int s = 0; s < C::NSLOT; ++s) tma_load_2d(dst + s * C::SLOT, &tmap, 64 * s, row, landed + b, policy);
                }
            }
            PROF_FLUSH(4, 1, true)
        }
        __syncwarp();
      } else if (warp == C::W_G1) {
        // =========================================================================== GEMM1 issuer
        if (elect_one()) {
            constexpr uint32_t idesc1 = umma_idesc(UMMA_BF16, UMMA_BF16, 128, TR, false, false);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(ring), 16, 1024);
            const uint32_t tq0 = tmem + C::TM_Q;
            uint32_t tt = 0;
            PROF_DECL
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t b = tt % C::NBUF, u = tt / C::NBUF, par = tt % C::NSET, v = tt / C::NSET;
                    PROF_BEGIN();
                    mbar_wait_wd(s_free + par, (v & 1u) ^ 1u);         // scores of tile tt - NSET have been read
                    PROF_END(0);
                    PROF_BEGIN();
                    mbar_wait_wd(landed + b, u & 1u);             // operand-ready as it lands: no conversion
                    PROF_END(1);
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(desc0, b * C::TILE);
                    const uint32_t d1 = tmem + C::TM_D1 + TR * par;
#pragma unroll
                    for (int s = 0; s < C::NSLOT; ++s) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)      // B = TR rows: the 8-row groups of the slot
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
            constexpr uint32_t idesc2 = umma_idesc(UMMA_BF16, UMMA_BF16, 128, 2 * NP, true, false);
            const uint64_t a0 = umma_desc_sw128(smem_u32(ring), C::SLOT, C::GRP);   // M atoms: next slot; K atoms: next row group
            const uint64_t w0 = umma_desc_sw128(smem_u32(wt), 16, 1024);
            uint32_t tt = 0, cc = 0;
            PROF_DECL
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x, ++cc) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + TR - 1) / TR);
                for (int t = 0; t < ntiles; ++t, ++tt) {
                    const uint32_t i = tt % C::NSET, v = tt / C::NSET, b = tt % C::NBUF;
                    PROF_BEGIN();
                    mbar_wait_wd(w_ready + i, v & 1u);
                    PROF_END(0);
                    PROF_BEGIN();
                    if (t == 0) mbar_wait_wd(d2_free, (cc & 1u) ^ 1u);   // previous chunk's accumulators drained
                    PROF_END(1);
                    // the weight warps' plain stores of the weight operand (generic proxy) become visible to the tensor core
                    // (async proxy) through THIS fence, on the causality path weight warps -> mbarrier w_ready -> issuer:
                    // the MEMBAR + FENCE pair stays out of the weight warps' per-tile latency chain
                    fence_proxy_async_smem();
                    tc_fence_after();
                    const uint64_t tb = umma_desc_advance(a0, b * C::TILE), wb = umma_desc_advance(w0, i * C::WBUF);
                    const uint32_t acc0 = t != 0;
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        const uint32_t d2 = tmem + C::TM_D2 + gg * C::D2W;
                        const uint64_t ah = umma_desc_advance(tb, (2 * gg) * C::SLOT);
#pragma unroll
                        for (int k2 = 0; k2 < TR / 16; ++k2)             // 16 tile rows per instruction: two 8-row groups
                            tc_mma_ss(d2, umma_desc_advance(ah, k2 * 2 * C::GRP), umma_desc_advance(wb, k2 * 32), idesc2, k2 == 0 ? acc0 : 1u);
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
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_SOFT));
        // NSET sets of four warps (one per TMEM quadrant) take the tiles in turn: tile tt belongs to set tt % NSET, which
        // also owns score buffer and weight buffer tt % NSET — NSET tiles are in this stage at any time (the stage is a
        // chain of mbarrier / TMEM / shared-memory round trips: latency, not issue slots, is what a set spends per tile).  A thread owns
        // (prototype p = 4 q + (lane & 3)) x (tile rows 2 rj, 2 rj + 1), so a warp sees all 16 rows of its four prototypes and
        // settles their softmax reference on its own.  What the sets share is the reference itself (s_mref): set s may
        // only decide tile tt after the other set has decided tile tt - 1 (mbarrier `decided`), and every thread folds a
        // reference it finds changed into its running sum before going on.
        const int q = warp & 3, set = warp >> 2, prev_set = (set + C::NSET - 1) % C::NSET;
        const int pl = lane & 3, p = 4 * q + pl, rj = lane >> 2;
        const bool pvalid = p < P;
        const uint32_t tq = tmem + (uint32_t(32 * q) << 16);
        constexpr int NT = C::NSOFT * 32, NTS = 128;
        // B operand of GEMM2: row (term * 16 + p), K = tile row: rows RPT rj .. + RPT - 1 -> 16-byte chunk (RPT rj) >> 3, bytes 2 ((RPT rj) & 7)
        const uint32_t w_off = sw128_offset(p, (C::RPT * rj) >> 3, 2 * ((C::RPT * rj) & 7));
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
            const int t_first = int((uint32_t(set) + C::NSET - tt0 % C::NSET) % C::NSET);
            uint32_t b = (tt0 + t_first) % C::NBUF, ph = ((tt0 + t_first) / C::NBUF) & 1u;
            tt = tt0 + uint32_t(ntiles);                           // for the next chunk
            for (int t = t_first; t < ntiles; t += C::NSET, b += uint32_t(C::NSET)) {
                if (b >= uint32_t(C::NBUF)) { b -= C::NBUF; ph ^= 1u; }
                const uint32_t tt = tt0 + uint32_t(t), v = tt / C::NSET;
                const int nvalid = min(TR, chunk_nrows - t * TR);
                PROF_BEGIN();
                mbar_wait_wd(s_ready + set, v & 1u);
                PROF_END(0);
                tc_fence_after();
                float sc2[C::RPT];
                {
                    // TR partial scores of this lane's (prototype, term, range): the tile's rows (zeros in the two spare lanes)
                    uint32_t sa[TR];
                    if constexpr (TR == 32) tmem_ld32(tq + C::TM_D1 + TR * set, *reinterpret_cast<uint32_t(*)[32]>(sa));
                    else tmem_ld16(tq + C::TM_D1 + TR * set, *reinterpret_cast<uint32_t(*)[16]>(sa));
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    mbar_arrive_if(s_free + set, lane == 0);
                    // the 6 (term, range) partial sums of a (row, prototype) (and two zeros) sit in the lanes that differ in bits 2-4: a
                    // transposed butterfly adds them in a fixed order and halves the rows a lane keeps at every level
                    // (7 TR / 8 shuffles, no shared memory): lane bit 4 -> the top row bit, bit 3 -> the next, bit 2 -> the next; a
                    // thread ends up with rows RPT (lane >> 2) .. + RPT - 1 of prototype lane & 3
                    constexpr int H1 = TR / 2, H2 = TR / 4, H3 = TR / 8;
                    float v1[H1], v2[H2];
                    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
                    for (int i = 0; i < H1; ++i) {
                        const float lo = __uint_as_float(sa[i]), hi = __uint_as_float(sa[H1 + i]);
                        v1[i] = (u16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, u16 ? lo : hi, 16);
                    }
#pragma unroll
                    for (int i = 0; i < H2; ++i) v2[i] = (u8 ? v1[H2 + i] : v1[i]) + __shfl_xor_sync(0xffffffffu, u8 ? v1[i] : v1[H2 + i], 8);
#pragma unroll
                    for (int i = 0; i < H3; ++i) sc2[i] = (u4 ? v2[H3 + i] : v2[i]) + __shfl_xor_sync(0xffffffffu, u4 ? v2[i] : v2[H3 + i], 4);
                }
                PROF_BEGIN();
                mbar_wait_wd(full + b, ph);                            // acquire the norm warps' row info
                PROF_END(1);
                float4 info[C::RPT];
#pragma unroll
                for (int k = 0; k < C::RPT; ++k) info[k] = *reinterpret_cast<const float4*>(s_rowinfo + (b * TR + C::RPT * rj + k) * 4);
                // ---- in tile order from here: the other set has settled tile tt - 1
                PROF_BEGIN();
                if (tt > 0) mbar_wait_wd(decided + prev_set, ((tt - 1) / C::NSET) & 1u);
                PROF_END(2);
                if (t == 0) {
                    exE = int(__float_as_uint(s_rowinfo[(b * TR) * 4 + 1]) >> 23);
                    if (q == 0 && lane == 0) s_exE[0] = exE;
                } else {
                    exE = s_exE[0];
                    const float m_sh = s_mref[p];                      // a reference (normaliser) the other set has moved
                    if (pvalid && m_sh > m_loc) {
                        if (!BWD) lsum *= expf(m_loc - m_sh);
                        m_loc = m_sh;
                    }
                }
                float w[C::RPT];                                       // weights fed to GEMM2 (before the bf16 split)
                float ts[C::RPT], unscale[C::RPT], cw[C::RPT];
                bool grow;
                if (!BWD) {
                    // ts = score + (e_row - E) ln 2: the weight is exp(ts - m) = A-weight x 2^(e_row - E)
#pragma unroll
                    for (int k = 0; k < C::RPT; ++k) {
                        int de = int(__float_as_uint(info[k].y) >> 23) - exE;
                        de = de < -100 ? -100 : (de > 100 ? 100 : de);
                        ts[k] = (C::RPT * rj + k < nvalid) ? fmaf(float(de), 0.693147180559945f, sc2[k] * info[k].x) : -INFINITY;
                        unscale[k] = __uint_as_float(uint32_t(127 - de) << 23);       // 2^-(e_row - E)
                    }
                    float tmax = ts[0];
#pragma unroll
                    for (int k = 1; k < C::RPT; ++k) tmax = fmaxf(tmax, ts[k]);
                    grow = pvalid && (tmax > m_loc + C::MARGIN);                       // true on the first tile
                } else {
                    // c = scale A (u - delta) / |x| = A (u - delta) info.x 2^-e; the weight on x~ = 2^-e x is c 2^e
#pragma unroll
                    for (int k = 0; k < C::RPT; ++k) {
                        const float a = expf(sc2[k] * info[k].x - bw_m) * bw_il;       // A_pn (deepmil.py:198)
                        cw[k] = (pvalid && C::RPT * rj + k < nvalid) ? a * (info[k].z - bw_delta) * info[k].x : 0.f;
                    }
                    // binary exponent of the larger |cw| (zero / denormal -> very small, non-finite -> very large)
                    float cmax = fabsf(cw[0]);
#pragma unroll
                    for (int k = 1; k < C::RPT; ++k) cmax = fmaxf(cmax, fabsf(cw[k]));
                    int et = int((__float_as_uint(cmax) >> 23) & 0xffu) - 127;
                    et = et < -100 ? -100 : (et > 100 ? 100 : et);
                    ts[0] = float(et);
                    grow = pvalid && (t == 0 || ts[0] > m_loc + float(C::BWD_MAXE));
                }
                PROF_BEGIN();
                const bool any_grow = named_bar_or(1 + set, NTS, grow);            // ids 1 .. NSET
                PROF_END(3);
                if (any_grow) {
                    // rare (always on the first tile of a chunk): the warp settles the new reference of its four prototypes
                    // from all 16 rows; from the second tile on the TMEM accumulators are rescaled once GEMM2 of the
                    // previous tile has completed (every warp of the set needs every prototype's factor: s_alpha)
                    float mt = ts[0];
                    if (!BWD) {
#pragma unroll
                        for (int k = 1; k < C::RPT; ++k) mt = fmaxf(mt, ts[k]);
                    }
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 4));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 8));
                    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 16));
                    float m_new = m_loc;
                    if (!BWD) { if (pvalid && mt > m_loc + C::MARGIN) m_new = mt + C::HEADROOM; }
                    else { if (pvalid && (t == 0 || mt > m_loc + float(C::BWD_MAXE))) m_new = mt - float(C::BWD_SETE); }
                    if (t > 0) {
                        const float alpha = (pvalid && m_new > m_loc) ? (BWD ? exp2f(m_loc - m_new) : expf(m_loc - m_new)) : 1.f;
                        if (lane < 4) s_alpha[p] = alpha;
                        mbar_wait_wd(w_free + prev_set, ((tt - 1) / C::NSET) & 1u);   // GEMM2 of tile tt - 1 (and of every earlier tile) has completed
                        tc_fence_after();
                        named_bar_sync(1 + C::NSET + set, NTS);
                        float al[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) al[j] = s_alpha[j];
#pragma unroll 1
                        for (int k = 0; k < C::D2W * 4 / 16; ++k) {
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
                        named_bar_sync(1 + C::NSET + set, NTS);                  // s_alpha may be rewritten by the next rare event
                    }
                    m_loc = m_new;
                    if (lane < 4) s_mref[p] = m_new;
                }
                __syncwarp();
                mbar_arrive_if(decided + set, lane == 0);             // releases s_mref / s_exE to the other set
                if (!BWD) {
#pragma unroll
                    for (int k = 0; k < C::RPT; ++k) {
                        w[k] = (pvalid && C::RPT * rj + k < nvalid) ? expf(ts[k] - m_loc) : 0.f;
                        lsum = fmaf(w[k], unscale[k], lsum);
                    }
                } else {
                    const float inv_h = pvalid ? __uint_as_float(uint32_t(127 - int(m_loc)) << 23) : 0.f;   // 1 / H_p
#pragma unroll
                    for (int k = 0; k < C::RPT; ++k) w[k] = cw[k] * inv_h;
                }
                // weights as two bf16 terms (w = t0 + t1, 16 significant bits; bf16 has the exponent range of fp32, no scaling);
                // B operand row (term * 16 + p), K = tile row (2 rj, 2 rj + 1)
                unsigned short b0[C::RPT], b1[C::RPT];
#pragma unroll
                for (int k = 0; k < C::RPT; ++k) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(w[k]);
                    const __nv_bfloat16 h1 = __float2bfloat16_rn(w[k] - __bfloat162float(h0));
                    b0[k] = __bfloat16_as_ushort(h0); b1[k] = __bfloat16_as_ushort(h1);
                }
                PROF_BEGIN();
                mbar_wait_wd(w_free + set, (v & 1u) ^ 1u);             // GEMM2 of tile tt - 2 has read this buffer
                PROF_END(4);
                // the thread's RPT tile rows are neighbours along K: one 4- or 8-byte store per term
                unsigned char* wb = wt + set * C::WBUF + w_off;
                if constexpr (C::RPT == 4) {
                    *reinterpret_cast<uint2*>(wb) = make_uint2(uint32_t(b0[0]) | (uint32_t(b0[1]) << 16), uint32_t(b0[2]) | (uint32_t(b0[3]) << 16));
                    *reinterpret_cast<uint2*>(wb + NP * 128) = make_uint2(uint32_t(b1[0]) | (uint32_t(b1[1]) << 16), uint32_t(b1[2]) | (uint32_t(b1[3]) << 16));
                } else {
                    *reinterpret_cast<uint32_t*>(wb) = uint32_t(b0[0]) | (uint32_t(b0[1]) << 16);
                    *reinterpret_cast<uint32_t*>(wb + NP * 128) = uint32_t(b1[0]) | (uint32_t(b1[1]) << 16);
                }
                __syncwarp();                                          // (the GEMM2 issuer fences for the async proxy)
                mbar_arrive_if(w_ready + set, lane == 0);
            }
            // ---- chunk end: both sets meet, agree on the final reference, write (m, l), drain O^T
            named_bar_sync(1 + 2 * C::NSET, NT);
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
            named_bar_sync(1 + 2 * C::NSET, NT);
            if (!BWD && set == 0 && lane < 4 && pvalid) {
                float lt = s_lsum[p];
#pragma unroll
                for (int k = 1; k < C::NSET; ++k) lt += s_lsum[16 * k + p];
                prm.part_l[size_t(c) * P + p] = lt;
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
            for (int gg = set * (4 / C::NSET); gg < (set + 1) * (4 / C::NSET); ++gg) {
                // lane = feature within a 128-block, columns = t0 | t1 per prototype
                uint32_t o0[16], o1[16];
                tmem_ld16(tq + C::TM_D2 + gg * C::D2W, o0);
                tmem_ld16(tq + C::TM_D2 + gg * C::D2W + 16, o1);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < P) po[size_t(j) * D + 128 * gg + 32 * q + lane] = mul[j] * (__uint_as_float(o0[j]) + __uint_as_float(o1[j]));
            }
            tc_fence_before();
            __syncwarp();
            mbar_arrive_if(d2_free, lane == 0);
            named_bar_sync(1 + 2 * C::NSET, NT);                                     // s_alpha / s_lsum / s_mref / s_exE are free again
        }
        PROF_FLUSH(12, 6, warp == 0 && lane == 0)
#ifdef VLSA_TMA_PROF
        if (blockIdx.x == 0 && warp == 0 && lane == 0) g_tma_prof[19] = tt;
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::W_G1) tmem_dealloc(tmem, C::TMEM_COLS);
#ifdef VLSA_TMA_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        g_tma_prof[20] = (long long)(t1 - prof_gt0);          // ns from kernel entry to exit, block 0
    }
#endif
}

}  // namespace vlsa
