// Fused language-guided aggregation on the 5th-generation tensor cores (tcgen05 + TMEM), forward pass.
// Same contract as agg_simt_kernel<P,false,float>: one read of X, per-chunk online-softmax partials
// (m, l, O[P,D]) — but both skinny contractions run as tcgen05.mma:
//   GEMM1  S[64 rows, p]   = X_tile[64, 512] . Qn[p, 512]^T        (A = X planes, K-major;  M = 64)
//   GEMM2  O[512 d, p]    += X_tile^T[512, 64] . W[p, 64]^T         (A = same planes, MN-major; M = 128 x 4)
// fp32 inputs are split on the fly into bf16 (hi, lo) planes (16 significant bits); Qn and the softmax
// weights are split the same way, so  S = (x_hi + x_lo).(q0 + q1),  O = (x_hi + x_lo).(w0 + w1)  with fp32
// accumulation in TMEM: fp32-grade results at 3x (not 7x) shared-memory traffic per byte of X.
//
// Warp roles (13 warps, 1 CTA / SM, persistent over chunks):
//   warps 0-3  softmax / epilogue : TMEM -> registers (lane quadrant = warp), weights -> smem, partial out
//   warp  4    MMA issuer (one elected thread) + TMEM allocator
//   warps 5-12 producers          : LDG.128 (2 items in flight per thread) -> split -> swizzled STS
// Ring of 10 slots x (8 KB hi + 8 KB lo); a tile = 64 rows = 8 slots (64 feature columns each).
#pragma once
#include <stdio.h>

#include "agg_simt.cuh"
#include "tc_common.cuh"

namespace vlsa {

template <int NP>
struct TcCfg {
    static constexpr int D = VLSA_D;
    static constexpr int TM = 64;                 // rows per tile (UMMA M of GEMM1)
    static constexpr int KC = 64;                 // feature columns per slot (128 B of bf16)
    static constexpr int NCH = D / KC;            // 8 slots per tile
    static constexpr int SLOTS = 10;              // even, so a GEMM2 slot pair never straddles the wrap
    static constexpr int PLANE = TM * 128;        // 8 KB
    static constexpr int SLOT = 2 * PLANE;        // hi | lo
    static constexpr int NB = 2 * NP;             // B rows: part 0 | part 1 of the bf16 split
    static constexpr int QCH = NB * 128;          // bytes of the Q operand per slot
    static constexpr int OFF_Q = SLOTS * SLOT;
    static constexpr int OFF_W = OFF_Q + NCH * QCH;
    static constexpr int OFF_F = OFF_W + NB * 128;
    static constexpr int NFLOAT = 64 + 2 * 64 + 64;     // ss[64] | red[2][4][16] | lred[4][16]
    static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
    static constexpr int NBAR = 2 * SLOTS + 5;
    static constexpr int SMEM = OFF_BAR + NBAR * 8 + 16 + 1024;
    static constexpr int NWARPS = 13;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int NPROD = 8;
    static constexpr int TMEM_COLS = (5 * NB <= 128) ? 128 : 256;   // D1: NB columns, D2: 4 x NB
    static constexpr float RESCALE_MARGIN = 20.f;
};

struct TileCursor {
    int c; long long r1, row; int kc; bool valid;
};

template <int NP>
__global__ void __launch_bounds__(TcCfg<NP>::THREADS, 1) agg_tc_kernel(const AggParams prm, const int P) {
    using C = TcCfg<NP>;
    constexpr int D = C::D, NB = C::NB, SLOTS = C::SLOTS, NCH = C::NCH;
    if (int(blockIdx.x) >= prm.total_chunks) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* ring = sm;
    unsigned char* qt = sm + C::OFF_Q;
    unsigned char* wt = sm + C::OFF_W;
    float* s_ss = reinterpret_cast<float*>(sm + C::OFF_F);
    float* s_red = s_ss + 64;            // [2][4][16]
    float* s_lred = s_red + 128;         // [4][16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + C::OFF_BAR);
    uint64_t* full = bars;               // [SLOTS] producers -> MMA
    uint64_t* empty = bars + SLOTS;      // [SLOTS] MMA (commit) -> producers
    uint64_t* s_ready = bars + 2 * SLOTS;      // GEMM1 of a tile done          (commit)   -> epilogue
    uint64_t* w_ready = s_ready + 1;           // weights of a tile in smem     (4 warps)  -> MMA
    uint64_t* ss_ready = s_ready + 2;          // row sums of squares in smem   (8 warps)  -> epilogue
    uint64_t* d2_done = s_ready + 3;           // GEMM2 of a chunk done         (commit)   -> epilogue
    uint64_t* d2_free = s_ready + 4;           // D2 drained                    (4 warps)  -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + C::NBAR);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < SLOTS; ++s) { mbar_init(full + s, C::NPROD); mbar_init(empty + s, 1); }
        mbar_init(s_ready, 1); mbar_init(w_ready, 4); mbar_init(ss_ready, C::NPROD); mbar_init(d2_done, 1);
        mbar_init(d2_free, 4);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, C::TMEM_COLS);
    // ---- Q operand: Qn = Q / max(|Q|, eps) split into bf16 (q0 | q1), K-major SW128 per 64-column slot
    for (int p = warp; p < NP; p += C::NWARPS) {
        float inv = 0.f;
        if (p < P) {
            float ss = 0.f;
            for (int d = lane; d < D; d += 32) { const float v = __ldg(prm.Q + size_t(p) * D + d); ss += v * v; }
            ss = warp_sum(ss);
            inv = 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
        }
        for (int d = lane; d < D; d += 32) {
            const float q = p < P ? __ldg(prm.Q + size_t(p) * D + d) * inv : 0.f;
            const __nv_bfloat16 q0 = __float2bfloat16_rn(q);
            const __nv_bfloat16 q1 = __float2bfloat16_rn(q - __bfloat162float(q0));
            const int kc = d >> 6, j = d & 63;
            unsigned char* base = qt + kc * C::QCH;
            *reinterpret_cast<__nv_bfloat16*>(base + sw128_offset(p, j >> 3, (j & 7) * 2)) = q0;
            *reinterpret_cast<__nv_bfloat16*>(base + sw128_offset(NP + p, j >> 3, (j & 7) * 2)) = q1;
        }
    }
    for (int i = tid; i < NB * 128 / 4; i += C::THREADS) reinterpret_cast<uint32_t*>(wt)[i] = 0u;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp >= 5) {
        // =========================================================================== producers
        const int pw = warp - 5, half = lane >> 4, c4 = lane & 15;
        const float* X = reinterpret_cast<const float*>(prm.X);
        auto start_cursor = [&](TileCursor& t, int c) {
            t.c = c; t.valid = c < prm.total_chunks; t.kc = 0;
            if (t.valid) { int bag; long long r0; chunk_info(prm, c, bag, r0, t.r1); t.row = r0; }
        };
        auto advance = [&](TileCursor& t) {
            if (++t.kc == NCH) {
                t.kc = 0; t.row += C::TM;
                if (t.row >= t.r1) start_cursor(t, t.c + gridDim.x);
            }
        };
        auto issue = [&](const TileCursor& t, float4 (&v)[4]) {
            const long long left = t.r1 - t.row;
            const int nvalid = left < C::TM ? int(left) : C::TM;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int r = 8 * pw + 2 * jj + half;
                if (r < nvalid) {
                    const float* src = X + (t.row + r) * D + t.kc * C::KC + c4 * 4;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[jj].x), "=f"(v[jj].y), "=f"(v[jj].z), "=f"(v[jj].w) : "l"(src));
                } else {
                    v[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        // software pipeline, PD items (slots) of loads in flight per thread: 8 warps x 32 lanes x PD x 64 B
        constexpr int PD = 3;
        TileCursor nxt;
        start_cursor(nxt, blockIdx.x);
        float4 buf[PD][4];
        bool have[PD];
#pragma unroll
        for (int u = 0; u < PD; ++u) {
            have[u] = nxt.valid;
            if (nxt.valid) { issue(nxt, buf[u]); advance(nxt); }
        }
        float ssq[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t q = 0;
        bool running = have[0];
        while (running) {
#pragma unroll
            for (int u = 0; u < PD; ++u) {
                if (!have[u]) { running = false; break; }
                const uint32_t slot = q % SLOTS;
                mbar_wait_wd(empty + slot, ((q / SLOTS) & 1u) ^ 1u);
                unsigned char* hi = ring + slot * C::SLOT;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int r = 8 * pw + 2 * jj + half;
                    const float4 v = buf[u][jj];
                    uint32_t h0, l0, h1, l1;
                    split_bf16x2(v.x, v.y, h0, l0);
                    split_bf16x2(v.z, v.w, h1, l1);
                    const uint32_t off = sw128_offset(r, c4 >> 1, (c4 & 1) * 8);
                    *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(hi + C::PLANE + off) = make_uint2(l0, l1);
                    ssq[jj] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                }
                // refill this register buffer with the item PD ahead before signalling
                have[u] = nxt.valid;
                if (nxt.valid) { issue(nxt, buf[u]); advance(nxt); }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + slot);
                if ((q & (NCH - 1)) == NCH - 1) {          // last slot of a tile: publish the row sums of squares
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float s = ssq[jj];
                        s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
                        if (c4 == 0) s_ss[8 * pw + 2 * jj + half] = s;
                        ssq[jj] = 0.f;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ss_ready);
                }
                ++q;
            }
        }
    } else if (warp == 4) {
        // =========================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc_bf16(64, NB, false, false);
            constexpr uint32_t idesc2 = umma_idesc_bf16(128, NB, true, false);
            const uint32_t ring_a = smem_u32(ring), q_a = smem_u32(qt), w_a = smem_u32(wt);
            uint32_t q = 0, tile_ctr = 0, chunk_ctr = 0;
            for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
                int bag; long long r0, r1;
                chunk_info(prm, c, bag, r0, r1);
                const int ntiles = int((r1 - r0 + C::TM - 1) / C::TM);
                for (int t = 0; t < ntiles; ++t) {
                    // ---- GEMM1: scores of this tile, slot by slot as the producers deliver
                    for (int kc = 0; kc < NCH; ++kc) {
                        const uint32_t qq = q + kc, slot = qq % SLOTS;
                        mbar_wait_wd(full + slot, (qq / SLOTS) & 1u);
                        tc_fence_after();
                        const uint32_t a_hi = ring_a + slot * C::SLOT, b0 = q_a + kc * C::QCH;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_bf16(tmem, umma_desc_sw128(a_hi + ks * 32, 16, 1024),
                                        umma_desc_sw128(b0 + ks * 32, 16, 1024), idesc1, (kc | ks) != 0);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_bf16(tmem, umma_desc_sw128(a_hi + C::PLANE + ks * 32, 16, 1024),
                                        umma_desc_sw128(b0 + ks * 32, 16, 1024), idesc1, 1u);
                    }
                    tc_commit(s_ready);
                    // ---- GEMM2: O += X^T . W once the softmax warps have published the weights
                    mbar_wait_wd(w_ready, tile_ctr & 1u);
                    if (t == 0) mbar_wait_wd(d2_free, (chunk_ctr & 1u) ^ 1u);   // previous chunk's D2 drained
                    tc_fence_after();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t slot = (q + 2 * g) % SLOTS;
                        const uint32_t a_hi = ring_a + slot * C::SLOT;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_bf16(tmem + C::NB + g * NB, umma_desc_sw128(a_hi + ks * 2048, C::SLOT, 1024),
                                        umma_desc_sw128(w_a + ks * 32, 16, 1024), idesc2, (t | ks) != 0);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_bf16(tmem + C::NB + g * NB, umma_desc_sw128(a_hi + C::PLANE + ks * 2048, C::SLOT, 1024),
                                        umma_desc_sw128(w_a + ks * 32, 16, 1024), idesc2, 1u);
                        tc_commit(empty + slot);
                        tc_commit(empty + slot + 1);
                    }
                    q += NCH;
                    ++tile_ctr;
                }
                tc_commit(d2_done);
                ++chunk_ctr;
            }
        }
        __syncwarp();
    } else {
        // =========================================================================== softmax / epilogue
        const bool rowlane = lane < 16;
        const int row = 16 * warp + (lane & 15);
        const uint32_t tq = tmem + (uint32_t(32 * warp) << 16);
        uint32_t tile_ctr = 0, chunk_ctr = 0;
        float m_ref[NP], lsum[NP];
        for (int c = blockIdx.x; c < prm.total_chunks; c += gridDim.x) {
            int bag; long long r0, r1;
            chunk_info(prm, c, bag, r0, r1);
            const int ntiles = int((r1 - r0 + C::TM - 1) / C::TM);
#pragma unroll
            for (int p = 0; p < NP; ++p) { m_ref[p] = -INFINITY; lsum[p] = 0.f; }
            for (int t = 0; t < ntiles; ++t) {
                const long long left = r1 - (r0 + (long long)t * C::TM);
                const int nvalid = left < C::TM ? int(left) : C::TM;
                mbar_wait_wd(s_ready, tile_ctr & 1u);
                tc_fence_after();
                uint32_t sv[NB];
#pragma unroll
                for (int k = 0; k < NB / 16; ++k) tmem_ld16(tq + 16 * k, *reinterpret_cast<uint32_t(*)[16]>(&sv[16 * k]));
                tmem_wait_ld();
                mbar_wait_wd(ss_ready, tile_ctr & 1u);
                const bool live = rowlane && row < nvalid;
                const float inv = prm.scale / fmaxf(sqrtf(s_ss[row]), VLSA_NORM_EPS);
                float s[NP];
                float* red = s_red + (tile_ctr & 1u) * 64;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    s[p] = live ? (__uint_as_float(sv[p]) + __uint_as_float(sv[NP + p])) * inv : -INFINITY;
                    const float mx = warp_max(s[p]);
                    if (lane == 0) red[warp * 16 + p] = mx;
                }
                named_bar_sync(1, 128);
                float tmax[NP];
                bool grow = false;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    tmax[p] = fmaxf(fmaxf(red[p], red[16 + p]), fmaxf(red[32 + p], red[48 + p]));
                    grow |= (p < P) && (tmax[p] > m_ref[p] + C::RESCALE_MARGIN);
                }
                if (t == 0) {
#pragma unroll
                    for (int p = 0; p < NP; ++p) m_ref[p] = tmax[p];
                } else if (grow) {
                    // rare: a later tile beats the reference max by > e^20 -> rescale the TMEM accumulators
                    float alpha[NP];
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        const float mn = fmaxf(m_ref[p], tmax[p]);
                        alpha[p] = (p < P) ? expf(m_ref[p] - mn) : 1.f;
                        m_ref[p] = mn;
                        lsum[p] *= alpha[p];
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int k = 0; k < NB / 16; ++k) {
                            uint32_t o[16];
                            tmem_ld16(tq + C::NB + g * NB + 16 * k, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha[(16 * k + i) % NP]);
                            tmem_st16(tq + C::NB + g * NB + 16 * k, o);
                        }
                    }
                    tmem_wait_st();
                }
                // weights: w = exp(s - m_ref) = w0 + w1 (bf16 pair); B operand row p | NP+p, K index = tile row
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float w = (live && p < P) ? expf(s[p] - m_ref[p]) : 0.f;
                    const __nv_bfloat16 w0 = __float2bfloat16_rn(w);
                    const __nv_bfloat16 w1 = __float2bfloat16_rn(w - __bfloat162float(w0));
                    lsum[p] += __bfloat162float(w0) + __bfloat162float(w1);
                    if (rowlane) {
                        *reinterpret_cast<__nv_bfloat16*>(wt + sw128_offset(p, row >> 3, (row & 7) * 2)) = w0;
                        *reinterpret_cast<__nv_bfloat16*>(wt + sw128_offset(NP + p, row >> 3, (row & 7) * 2)) = w1;
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(w_ready);
                ++tile_ctr;
            }
            // ---- chunk end: drain D2, publish the partial
            mbar_wait_wd(d2_done, chunk_ctr & 1u);
            tc_fence_after();
            float* po = prm.part_O + size_t(c) * P * D;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t o[NB];
#pragma unroll
                for (int k = 0; k < NB / 16; ++k) tmem_ld16(tq + C::NB + g * NB + 16 * k, *reinterpret_cast<uint32_t(*)[16]>(&o[16 * k]));
                tmem_wait_ld();
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    if (p < P) po[size_t(p) * D + 128 * g + 32 * warp + lane] = __uint_as_float(o[p]) + __uint_as_float(o[NP + p]);
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float v = warp_sum(rowlane ? lsum[p] : 0.f);
                if (lane == 0) s_lred[warp * 16 + p] = v;
            }
            tc_fence_before();
            named_bar_sync(1, 128);
            if (tid < P) {
                prm.part_l[size_t(c) * P + tid] = (s_lred[tid] + s_lred[16 + tid]) + (s_lred[32 + tid] + s_lred[48 + tid]);
                float mr = -INFINITY;
#pragma unroll
                for (int p = 0; p < NP; ++p) if (p == tid) mr = m_ref[p];
                prm.part_m[size_t(c) * P + tid] = mr;
            }
            named_bar_sync(1, 128);          // s_lred is reused by the next chunk
            if (lane == 0) mbar_arrive(d2_free);
            ++chunk_ctr;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace vlsa
