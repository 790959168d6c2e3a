// Fused language-guided aggregation, streaming pass over X (replaces model/deepmil.py:187-200 of
// liupei101/VLSA).  One read of X; per chunk of rows the kernel produces an
// online-softmax partial (m, l, O[P,D]) (forward) or a partial dQn[P,D] (backward).
//
// CUDA-core (fp32 FFMA) variant: TMA bulk copies stage [TN,D] tiles of X in shared memory behind an
// mbarrier ring; phase A = P+1 dot products per row with a transposed-butterfly warp reduction,
// phase S = per-tile softmax bookkeeping, phase B = P-way weighted accumulation with one thread per
// two feature columns.  Deterministic: a chunk's partial depends only on the chunk.
#pragma once
#include "common.cuh"

namespace vlsa {

struct AggParams {
    const void* X;              // [total_rows, D] fp32 or bf16
    const long long* cu_rows;   // [B+1] row offsets of bags (device) | row_ranges: [2B] (first row, one past the last row) per bag
    int row_ranges;             // 1: the bags lie anywhere inside X, in any order (a step drawn from a device-resident cohort)
    const int* chunk_start;     // [B+1] first chunk id of each bag (device)
    int B;
    int chunk_rows;             // rows per chunk (multiple of TN)
    int total_chunks;
    const float* Q;             // [P, D] un-normalised queries
    int q_prenorm;              // 1: rows of Q are the score directions as they are (gated queries: the difference
                                //    of two unit vectors, deepmil.py:192-195); 0: normalise them (deepmil.py:187)
    float scale;                // exp(fp32(log 100)), deepmil.py:122
    // forward outputs: per-chunk partials
    float* part_m;              // [chunks, P]
    float* part_l;              // [chunks, P]
    float* part_O;              // [chunks, P, D]  (backward: partial dQn)
    // backward inputs (per bag)
    const float* dv;            // MODE 1: [B, D] d loss / d pooled (divided by P inside the kernel)
                                // MODE 2: [B, P, D] d loss / d O_p, one gradient row per prototype
    const float* ml;            // [B, P, 2] (max, sum) from forward
    const float* delta;         // [B, P]   dO_p . O_p
    int p_stride;               // MODE 2: prototypes per bag in dv / ml / delta (>= P: a launch may serve a sub-range of
                                //         the prototypes, the pointers then start at its first one)
};

#ifndef VLSA_AGG_WARPS
#define VLSA_AGG_WARPS 4
#endif
#ifndef VLSA_AGG_STAGES
#define VLSA_AGG_STAGES 2
#endif

// MODE 0: forward.  MODE 1: backward of the mean-pooled path (one gradient row dv / P shared by all prototypes).
// MODE 2: backward with a gradient row per prototype (any pooling over the P outputs: max, weight, attention).
template <int P, int MODE, typename XT>
struct AggCfg {
    static constexpr bool BWD = MODE != 0, GEN = MODE == 2;
    static constexpr int D = VLSA_D;
    static constexpr int NW = VLSA_AGG_WARPS;          // warps per CTA (several CTAs share an SM)
    static constexpr int THREADS = 32 * NW;
    static constexpr int TN = 4 * NW;                  // rows per tile (4 rows per warp in phase A)
    static constexpr int CPT = D / THREADS;            // feature columns per thread in phase B
    static constexpr int STAGES = VLSA_AGG_STAGES;
    static constexpr int NQ = GEN ? 2 * P : (BWD ? P + 1 : P);   // query (+ gradient) rows resident in smem
    static constexpr int NRED = NQ + 1;                // + sum of squares
    static constexpr int PP = (P + 3) & ~3;            // weights per row, float4-padded
    static constexpr int NV = 4 * NRED;                // values reduced per warp per tile
    static constexpr size_t XS_BYTES = size_t(STAGES) * TN * D * sizeof(XT);
    static constexpr size_t QS_BYTES = size_t(NQ) * D * sizeof(float);
    static constexpr size_t RED_BYTES = size_t(TN) * NRED * sizeof(float);
    static constexpr size_t WT_BYTES = size_t(TN) * PP * sizeof(float);
    static constexpr size_t SC_BYTES = 6 * VLSA_MAX_P * sizeof(float);
    static constexpr size_t SMEM = XS_BYTES + QS_BYTES + RED_BYTES + WT_BYTES + SC_BYTES + STAGES * 8 + 128;
};

// chunk id -> (bag, first row, one-past-last row); chunk_start is monotone, bags with 0 rows have 0 chunks
__device__ __forceinline__ void chunk_info(const AggParams& p, int c, int& bag, long long& r0, long long& r1) {
    int lo = 0, hi = p.B;            // find bag with chunk_start[bag] <= c < chunk_start[bag+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.chunk_start + mid) <= c) lo = mid; else hi = mid;
    }
    bag = lo;
    const long long b0 = __ldg(p.cu_rows + (p.row_ranges ? 2 * lo : lo)), b1 = __ldg(p.cu_rows + (p.row_ranges ? 2 * lo + 1 : lo + 1));
    r0 = b0 + (long long)(c - __ldg(p.chunk_start + lo)) * p.chunk_rows;
    r1 = r0 + p.chunk_rows < b1 ? r0 + p.chunk_rows : b1;
}

template <int P, int MODE, typename XT>
__global__ void __launch_bounds__(AggCfg<P, MODE, XT>::THREADS) agg_simt_kernel(const AggParams prm) {
    using C = AggCfg<P, MODE, XT>;
    constexpr bool BWD = C::BWD, GEN = C::GEN;
    constexpr int D = C::D, TN = C::TN, STAGES = C::STAGES, NQ = C::NQ, NRED = C::NRED, PP = C::PP, NV = C::NV;
    constexpr int NW = C::NW, CPT = C::CPT;
    static_assert(TN <= 32 && (CPT == 2 || CPT == 4), "phase S maps rows to lanes; phase B loads 8 or 16 bytes");
    constexpr bool BF16 = sizeof(XT) == 2;
#ifndef VLSA_SIMT_UNROLL_B
#define VLSA_SIMT_UNROLL_B 4
#endif
    constexpr int UNROLL_B = VLSA_SIMT_UNROLL_B;   // rows of phase B in flight per thread
#ifdef VLSA_SIMT_NOPACK
    constexpr bool PACKED_B = false;
#else
    constexpr bool PACKED_B = true;
#endif
#ifdef VLSA_SIMT_NOPACK
    constexpr bool PACKED = false;
#else
    constexpr bool PACKED = NV <= 40;          // packed phase-A accumulators double their register count: P <= 8
#endif

    extern __shared__ __align__(128) unsigned char smem_raw[];
    XT* xs = reinterpret_cast<XT*>(smem_raw);
    float* qs = reinterpret_cast<float*>(smem_raw + C::XS_BYTES);
    float* red = reinterpret_cast<float*>(smem_raw + C::XS_BYTES + C::QS_BYTES);
    float* wt = reinterpret_cast<float*>(smem_raw + C::XS_BYTES + C::QS_BYTES + C::RED_BYTES);
    float* sc = reinterpret_cast<float*>(smem_raw + C::XS_BYTES + C::QS_BYTES + C::RED_BYTES + C::WT_BYTES);
    float* s_m = sc;                       // running max        (fwd) | saved max        (bwd)
    float* s_l = sc + VLSA_MAX_P;          // running sum        (fwd) | 1 / saved sum    (bwd)
    float* s_alpha = sc + 2 * VLSA_MAX_P;  // per-tile rescale   (fwd) | delta_p          (bwd)
    float* s_lw = sc + 3 * VLSA_MAX_P;     // [NW][P] per-warp softmax sums at a chunk end (fused forward)
    int* s_flag = reinterpret_cast<int*>(sc + 5 * VLSA_MAX_P);   // "a row beat the lazy reference" (fused forward)
    uint64_t* full = reinterpret_cast<uint64_t*>(
        (reinterpret_cast<uintptr_t>(sc + 6 * VLSA_MAX_P) + 7) & ~uintptr_t(7));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (int(blockIdx.x) >= prm.total_chunks) return;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    // Qn = Q / max(||Q||, eps)  (deepmil.py:187): warp w normalises rows w, w+NW, ...
    for (int p = warp; p < P; p += NW) {
        const float* q = prm.Q + size_t(p) * D;
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = __ldg(q + d); ss += v * v; }
        ss = warp_sum(ss);
        const float inv = prm.q_prenorm ? 1.f : 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
        for (int d = lane; d < D; d += 32) qs[p * D + d] = __ldg(q + d) * inv;
    }
    if (tid < P) { s_m[tid] = -INFINITY; s_l[tid] = 0.f; s_alpha[tid] = 0.f; }
    if (tid == 0) *s_flag = 0;
    __syncthreads();
#ifndef VLSA_SIMT_QREG_J
#define VLSA_SIMT_QREG_J 2
#endif
#ifndef VLSA_SIMT_QREG_J_BWD
#define VLSA_SIMT_QREG_J_BWD 2
#endif
    constexpr int QREG_J = (!BF16 && PACKED && P <= 4 && !GEN) ? (BWD ? VLSA_SIMT_QREG_J_BWD : VLSA_SIMT_QREG_J) : 0;
#ifndef VLSA_SIMT_QREG_JB
#define VLSA_SIMT_QREG_JB 2
#endif
    // bf16 storage: a lane owns 8 consecutive columns per 256-column block, i.e. two float4 of every query row
    constexpr int QREG_JB = (BF16 && PACKED && P <= 4 && !GEN) ? VLSA_SIMT_QREG_JB : 0;
    float4 qregb[QREG_JB > 0 ? QREG_JB * 2 * NQ : 1];
    auto load_qregb = [&](int q) {
#pragma unroll
        for (int j = 0; j < QREG_JB; ++j) {
            qregb[(j * NQ + q) * 2 + 0] = *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8);
            qregb[(j * NQ + q) * 2 + 1] = *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8 + 4);
        }
    };
    if (QREG_JB > 0) {
#pragma unroll
        for (int q = 0; q < P; ++q) load_qregb(q);
    }
    float4 qreg[QREG_J > 0 ? QREG_J * NQ : 1];
    if (QREG_J > 0) {
#pragma unroll
        for (int j = 0; j < QREG_J; ++j)
#pragma unroll
            for (int q = 0; q < P; ++q) qreg[j * NQ + q] = *reinterpret_cast<const float4*>(qs + q * D + j * 128 + lane * 4);
    }

    const uint64_t policy = make_evict_first_policy();

    // ---- producer cursor (thread 0) and consumer cursor (all threads, uniform) ---------------
    int pc = blockIdx.x; long long pr0 = 0, pr1 = 0; int pbag = 0; long long prow = 0;
    bool phas = pc < prm.total_chunks;
    if (phas) { chunk_info(prm, pc, pbag, pr0, pr1); prow = pr0; }
    int issued = 0;
    auto produce = [&]() {      // thread 0 only
        if (!phas) return;
        const int stage = issued % STAGES;
        const long long left = pr1 - prow;
        const int n = left < TN ? int(left) : TN;
        const uint32_t bytes = uint32_t(n) * D * sizeof(XT);
        mbar_expect_tx(full + stage, bytes);
        bulk_g2s_evict_first(xs + size_t(stage) * TN * D, reinterpret_cast<const XT*>(prm.X) + prow * D, bytes,
                             full + stage, policy);
        ++issued;
        prow += n;
        if (prow >= pr1) {
            pc += gridDim.x;
            phas = pc < prm.total_chunks;
            if (phas) { chunk_info(prm, pc, pbag, pr0, pr1); prow = pr0; }
        }
    };
    if (tid == 0) for (int s = 0; s < STAGES - 1; ++s) produce();

    constexpr float LAZY_MARGIN = 16.f;
    // forward with all 4 (P + 1) per-warp values in one butterfly group (P <= 7): the weights of a warp's own rows are
    // computed straight from the butterfly result against the current lazy reference — no separate softmax phase, no
    // barrier before it; a tile whose rows beat the reference takes the rare path below
    constexpr bool FUSED_S = !BWD && NV <= 32;
    // backward with one butterfly group (P <= 6): the weight of (row, prototype) needs only that row's values and the
    // saved per-bag (m, 1/l, delta): computed by the lane that holds the score, no softmax phase and no barrier for it
    constexpr bool FUSED_W = BWD && NV <= 32;
    float lpart_f = 0.f;                       // fused forward: this lane's (row, prototype) share of the sum l
    float lpart[(P + NW - 1) / NW];
#pragma unroll
    for (int k = 0; k < (P + NW - 1) / NW; ++k) lpart[k] = 0.f;
    float acc2[P][CPT];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int k = 0; k < CPT; ++k) acc2[p][k] = 0.f;

    int it = 0;
    for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
        int bag; long long r0, r1;
        chunk_info(prm, c, bag, r0, r1);
        if (BWD) {
            // per-bag operands: extra query row(s) dv_b / P | dO_b, saved (m, 1/l), delta_p = (dv . O_p) / P | dO_p . O_p
            __syncthreads();
            const size_t bp = size_t(bag) * (GEN ? prm.p_stride : P);       // first prototype of the bag
            if (GEN) {
                for (int i = tid; i < P * D; i += C::THREADS) qs[P * D + i] = __ldg(prm.dv + bp * D + i);
            } else {
                const float invP = 1.f / float(P);
                for (int d = tid; d < D; d += C::THREADS) qs[P * D + d] = __ldg(prm.dv + size_t(bag) * D + d) * invP;
            }
            if (tid < P) {
                s_m[tid] = __ldg(prm.ml + (bp + tid) * 2);
                s_l[tid] = 1.f / __ldg(prm.ml + (bp + tid) * 2 + 1);
                s_alpha[tid] = __ldg(prm.delta + bp + tid);
            }
            __syncthreads();
            if (QREG_J > 0) {                   // the bag's extra query row dv / P joins the register-resident Qn
#pragma unroll
                for (int j = 0; j < QREG_J; ++j)
                    qreg[j * NQ + P] = *reinterpret_cast<const float4*>(qs + P * D + j * 128 + lane * 4);
            }
            if (QREG_JB > 0) load_qregb(P);
        }
        for (long long row = r0; row < r1; row += TN, ++it) {
            const int stage = it % STAGES;
            const uint32_t parity = (it / STAGES) & 1;
            if (tid == 0) produce();          // tile it+STAGES-1 -> stage freed by the previous trailing barrier
            const int nvalid = (r1 - row) < TN ? int(r1 - row) : TN;
            const XT* xt = xs + size_t(stage) * TN * D;
            mbar_wait(full + stage, parity);

            // ---------------- phase A: NQ dots + sum of squares for rows 4*warp .. 4*warp+3 --------
            int fr = 0, fq = 0; bool fact = false; float fs = 0.f, fw = 0.f;     // fused forward: this lane's (row, prototype)
            float acc[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.f;
            if (!BF16 && PACKED) {
                // packed fp32x2 FMAs (sm_100 FFMA2): two partial sums per value, half the FMA instructions
                float2 acc2p[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) acc2p[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 xv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        xv[r] = *reinterpret_cast<const float4*>(
                            reinterpret_cast<const float*>(xt) + size_t(4 * warp + r) * D + j * 128 + lane * 4);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float2 lo = make_float2(xv[r].x, xv[r].y), hi = make_float2(xv[r].z, xv[r].w);
                        acc2p[r * NRED + NQ] = __ffma2_rn(lo, lo, acc2p[r * NRED + NQ]);
                        acc2p[r * NRED + NQ] = __ffma2_rn(hi, hi, acc2p[r * NRED + NQ]);
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        // the first QREG_J column blocks of the query rows live in registers (small P): the per-warp,
                        // per-tile re-read of Qn is a third of this kernel's shared-memory traffic
                        const float4 qv = (QREG_J > 0 && j < QREG_J) ? qreg[(j < QREG_J ? j : 0) * NQ + q]
                                        : *reinterpret_cast<const float4*>(qs + q * D + j * 128 + lane * 4);
                        const float2 qlo = make_float2(qv.x, qv.y), qhi = make_float2(qv.z, qv.w);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            acc2p[r * NRED + q] = __ffma2_rn(qlo, make_float2(xv[r].x, xv[r].y), acc2p[r * NRED + q]);
                            acc2p[r * NRED + q] = __ffma2_rn(qhi, make_float2(xv[r].z, xv[r].w), acc2p[r * NRED + q]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] = acc2p[i].x + acc2p[i].y;
            } else if (!BF16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 xv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        xv[r] = *reinterpret_cast<const float4*>(
                            reinterpret_cast<const float*>(xt) + size_t(4 * warp + r) * D + j * 128 + lane * 4);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        acc[r * NRED + NQ] += xv[r].x * xv[r].x + xv[r].y * xv[r].y + xv[r].z * xv[r].z + xv[r].w * xv[r].w;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 qv = *reinterpret_cast<const float4*>(qs + q * D + j * 128 + lane * 4);
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            acc[r * NRED + q] += qv.x * xv[r].x + qv.y * xv[r].y + qv.z * xv[r].z + qv.w * xv[r].w;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float xf[4][8];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint4 u = *reinterpret_cast<const uint4*>(
                            reinterpret_cast<const __nv_bfloat16*>(xt) + size_t(4 * warp + r) * D + j * 256 + lane * 8);
                        xf[r][0] = bf16_lo(u.x); xf[r][1] = bf16_hi(u.x); xf[r][2] = bf16_lo(u.y); xf[r][3] = bf16_hi(u.y);
                        xf[r][4] = bf16_lo(u.z); xf[r][5] = bf16_hi(u.z); xf[r][6] = bf16_lo(u.w); xf[r][7] = bf16_hi(u.w);
                    }
                    if (PACKED) {
                        // packed fp32x2 FMAs; the two halves of every pair are added once per 256-column block
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            float2 a = make_float2(0.f, 0.f);
#pragma unroll
                            for (int k = 0; k < 8; k += 2) {
                                const float2 x2 = make_float2(xf[r][k], xf[r][k + 1]);
                                a = __ffma2_rn(x2, x2, a);
                            }
                            acc[r * NRED + NQ] += a.x + a.y;
                        }
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const bool in_reg = QREG_JB > 0 && j < QREG_JB;
                            const float4 qa = in_reg ? qregb[((j < QREG_JB ? j : 0) * NQ + q) * 2 + 0]
                                                     : *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8);
                            const float4 qb = in_reg ? qregb[((j < QREG_JB ? j : 0) * NQ + q) * 2 + 1]
                                                     : *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8 + 4);
                            const float2 q2[4] = {make_float2(qa.x, qa.y), make_float2(qa.z, qa.w), make_float2(qb.x, qb.y),
                                                  make_float2(qb.z, qb.w)};
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                                for (int k = 0; k < 4; ++k) a = __ffma2_rn(q2[k], make_float2(xf[r][2 * k], xf[r][2 * k + 1]), a);
                                acc[r * NRED + q] += a.x + a.y;
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int k = 0; k < 8; ++k) acc[r * NRED + NQ] += xf[r][k] * xf[r][k];
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const float4 qa = *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8);
                            const float4 qb = *reinterpret_cast<const float4*>(qs + q * D + j * 256 + lane * 8 + 4);
#pragma unroll
                            for (int r = 0; r < 4; ++r)
                                acc[r * NRED + q] += qa.x * xf[r][0] + qa.y * xf[r][1] + qa.z * xf[r][2] + qa.w * xf[r][3] +
                                                     qb.x * xf[r][4] + qb.y * xf[r][5] + qb.z * xf[r][6] + qb.w * xf[r][7];
                        }
                    }
                }
            }
            // transposed butterfly in groups of 32 values; lane L of group g ends with value 32g+L
#pragma unroll
            for (int g = 0; g * 32 < NV; ++g) {
                constexpr int dummy = 0; (void)dummy;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (g * 32 + i < NV) ? acc[(g * 32 + i < NV) ? g * 32 + i : 0] : 0.f;
                float tot;
                if (NV - g * 32 >= 32) tot = warp_reduce_transpose<32>(v);
                else tot = warp_reduce_transpose<(NV % 32 == 0 ? 32 : NV % 32)>(v);
                const int idx = g * 32 + lane;
                if (FUSED_W) {
                    // lane i < NV: (row r = i / NRED, column q = i % NRED); q < P: Qn_q . x, q == P (+ q): gradient row . x,
                    // q == NQ: |x|^2
                    const int r = lane / NRED, q = lane % NRED;
                    const float ssv = __shfl_sync(0xffffffffu, tot, (r * NRED + NQ) & 31);
                    const float uv = __shfl_sync(0xffffffffu, tot, (r * NRED + P + (GEN && q < P ? q : 0)) & 31);
                    const int rowl = 4 * warp + r;
                    if (lane < NV && q < P) {
                        const float nrm = fmaxf(sqrtf(ssv), VLSA_NORM_EPS);
                        const float sv = prm.scale * (tot / nrm);
                        const float a = expf(sv - s_m[q]) * s_l[q];                      // A_pn (deepmil.py:198)
                        wt[rowl * PP + q] = rowl < nvalid ? prm.scale * a * (uv - s_alpha[q]) / nrm : 0.f;
                    }
                } else if (!FUSED_S) {
                    if (idx < NV) red[(4 * warp + idx / NRED) * NRED + idx % NRED] = tot;
                } else {
                    // lane i < NV holds value (row r = i / NRED of this warp, column q = i % NRED; q == NQ: |x_r|^2)
                    fr = lane / NRED; fq = lane % NRED;
                    const float ssv = __shfl_sync(0xffffffffu, tot, (fr * NRED + NQ) & 31);
                    fact = lane < NV && fq < P;
                    const int rowl = 4 * warp + fr;
                    const float nrm = fmaxf(sqrtf(ssv), VLSA_NORM_EPS);
                    fs = (fact && rowl < nvalid) ? prm.scale * (tot / nrm) : -INFINITY;
                    const float mo = fact ? s_m[fq < P ? fq : 0] : 0.f;
                    fw = fact ? expf(fs - mo) : 0.f;                 // tentative: right unless the reference moves
                    if (fact) { wt[rowl * PP + fq] = fw; red[rowl * NRED + fq] = fs; }
                    if (__any_sync(0xffffffffu, fact && fs > mo + LAZY_MARGIN) && lane == 0) *s_flag = 1;
                }
            }
            __syncthreads();
            bool grew = false;
            if (FUSED_S) {
                grew = *reinterpret_cast<volatile int*>(s_flag) != 0;      // block-uniform
                if (grew) {
                    // rare (always on the first tile of a chunk): move the reference of the prototypes that were beaten,
                    // rewrite the tile's weights; one warp per prototype, lane = row
                    for (int p = warp; p < P; p += NW) {
                        const float sv = lane < TN ? red[lane * NRED + p] : -INFINITY;
                        const float tmax = warp_max(sv);
                        const float mo = s_m[p];
                        const float mn = tmax > mo + LAZY_MARGIN ? tmax : mo;
                        if (lane < TN) wt[lane * PP + p] = sv == -INFINITY ? 0.f : expf(sv - mn);
                        __syncwarp();                                   // every lane has read s_m[p]
                        if (lane == 0) { s_alpha[p] = expf(mo - mn); s_m[p] = mn; }
                    }
                    __syncthreads();
                    if (tid == 0) *s_flag = 0;
                    if (fact) {
                        const int qq = fq < P ? fq : 0;
                        fw = fs == -INFINITY ? 0.f : expf(fs - s_m[qq]);
                        lpart_f = fmaf(lpart_f, s_alpha[qq], fw);
                    }
                } else if (fact) {
                    lpart_f += fw;
                }
            }

            // ---------------- phase S: per-(row, p) weights; lane = row, warp handles p = warp, warp+NW ----
            for (int p = warp; p < P && !FUSED_S && !FUSED_W; p += NW) {
                const int rl = lane < TN ? lane : 0;
                const float dot = red[rl * NRED + p], ss = red[rl * NRED + NQ];
                const float nrm = fmaxf(sqrtf(ss), VLSA_NORM_EPS);
                const bool live = lane < nvalid;
                const float s = live ? prm.scale * (dot / nrm) : -INFINITY;
                if (!BWD) {
                    // online softmax with a LAZY reference: it only moves (one warp_max, accumulators rescaled) when a
                    // row beats it by more than LAZY_MARGIN; otherwise a tile costs one vote and no shuffle chain.  Any
                    // reference gives the same (m, l, O) triple up to the common factor the merge removes; weights stay
                    // <= e^LAZY_MARGIN.  The per-prototype sum l is kept per lane and reduced once per chunk.
                    const float mo = s_m[p];
                    float mn = mo, a = 1.f;
                    if (__any_sync(0xffffffffu, s > mo + LAZY_MARGIN)) {        // always on the first tile (mo = -inf)
                        mn = warp_max(s);
                        a = expf(mo - mn);
                    }
                    const float w = expf(s - mn);
                    float& lp = lpart[(p - warp) / NW];
                    lp = fmaf(lp, a, w);
                    if (lane < TN) wt[lane * PP + p] = w;
                    __syncwarp();                                       // every lane has read s_m[p]
                    if (lane == 0) { s_alpha[p] = a; s_m[p] = mn; }
                } else {
                    const float a = expf(s - s_m[p]) * s_l[p];                 // A_pn (deepmil.py:198)
                    const float u = red[rl * NRED + P + (GEN ? p : 0)];         // (dv . x_n) / P | dO_p . x_n
                    if (lane < TN) wt[lane * PP + p] = live ? prm.scale * a * (u - s_alpha[p]) / nrm : 0.f;
                }
            }
            if (!FUSED_S && !FUSED_W) __syncthreads();

            // ---------------- phase B: acc2[p][:] (+)= sum_r w[r][p] * x[r][CPT*tid .. CPT*tid+CPT-1] ---------
            if (!BWD && (!FUSED_S || grew)) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float a = s_alpha[p];
#pragma unroll
                    for (int k = 0; k < CPT; ++k) acc2[p][k] *= a;
                }
            }
#pragma unroll UNROLL_B
            for (int r = 0; r < nvalid; ++r) {
                float xv[CPT];
                if (!BF16) {
                    const float* src = reinterpret_cast<const float*>(xt) + size_t(r) * D + CPT * tid;
                    if (CPT == 4) {
                        const float4 v4 = *reinterpret_cast<const float4*>(src);
                        xv[0] = v4.x; xv[1] = v4.y; xv[CPT - 2] = v4.z; xv[CPT - 1] = v4.w;
                    } else {
                        const float2 v2 = *reinterpret_cast<const float2*>(src);
                        xv[0] = v2.x; xv[1] = v2.y;
                    }
                } else {
                    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(xt) + size_t(r) * D + CPT * tid;
                    if (CPT == 4) {
                        const uint2 u = *reinterpret_cast<const uint2*>(src);
                        xv[0] = bf16_lo(u.x); xv[1] = bf16_hi(u.x); xv[CPT - 2] = bf16_lo(u.y); xv[CPT - 1] = bf16_hi(u.y);
                    } else {
                        const uint32_t u = *reinterpret_cast<const uint32_t*>(src);
                        xv[0] = bf16_lo(u); xv[1] = bf16_hi(u);
                    }
                }
#pragma unroll
                for (int k4 = 0; k4 < PP / 4; ++k4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wt + r * PP + 4 * k4);
                    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (4 * k4 + j < P) {
                            float* a = acc2[(4 * k4 + j < P) ? 4 * k4 + j : 0];
                            if (CPT == 4 && PACKED_B) {
                                // FFMA2 with the weight as the broadcast scalar operand
                                const float2 ww = make_float2(wv[j], wv[j]);
                                const float2 r0 = __ffma2_rn(ww, make_float2(xv[0], xv[1]), make_float2(a[0], a[1]));
                                const float2 r1 = __ffma2_rn(ww, make_float2(xv[CPT - 2], xv[CPT - 1]), make_float2(a[CPT - 2], a[CPT - 1]));
                                a[0] = r0.x; a[1] = r0.y; a[CPT - 2] = r1.x; a[CPT - 1] = r1.y;
                            } else {
#pragma unroll
                                for (int k = 0; k < CPT; ++k) a[k] += wv[j] * xv[k];
                            }
                        }
                    }
                }
            }

            if (row + TN >= r1) {
                // ---------------- chunk done: write the partial and reset ------------------------------
                float* o = prm.part_O + size_t(c) * P * D;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (CPT == 4)
                        *reinterpret_cast<float4*>(o + size_t(p) * D + CPT * tid) =
                            make_float4(acc2[p][0], acc2[p][1], acc2[p][CPT - 2], acc2[p][CPT - 1]);
                    else
                        *reinterpret_cast<float2*>(o + size_t(p) * D + CPT * tid) = make_float2(acc2[p][0], acc2[p][1]);
#pragma unroll
                    for (int k = 0; k < CPT; ++k) acc2[p][k] = 0.f;
                }
                if (FUSED_S) {
                    // l_p = sum over the 4 rows of every warp: lanes (r, q) -> lane q by two shuffles, warps via smem
                    float v = lpart_f;
                    v += __shfl_down_sync(0xffffffffu, v, 2 * NRED);
                    v += __shfl_down_sync(0xffffffffu, v, NRED);
                    if (lane < P) s_lw[warp * P + lane] = v;
                    lpart_f = 0.f;
                    __syncthreads();
                    if (tid < P) {
                        float l = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; ++w) l += s_lw[w * P + tid];
                        prm.part_m[size_t(c) * P + tid] = s_m[tid];
                        prm.part_l[size_t(c) * P + tid] = l;
                        s_m[tid] = -INFINITY;
                    }
                } else if (!BWD) {
#pragma unroll
                    for (int k = 0; k < (P + NW - 1) / NW; ++k) {
                        const int p = warp + k * NW;                    // the prototypes phase S of this warp owns
                        if (p < P) {
                            const float l = warp_sum(lpart[k]);
                            if (lane == 0) {
                                prm.part_m[size_t(c) * P + p] = s_m[p];
                                prm.part_l[size_t(c) * P + p] = l;
                                s_m[p] = -INFINITY;
                            }
                        }
                        lpart[k] = 0.f;
                    }
                }
            }
            __syncthreads();    // stage `stage`, red and wt are free again
        }
    }
}

}  // namespace vlsa
