// Epilogue kernels of the aggregation path: split-N merge, visual adapter (nn.Linear 512x512),
// cosine head against the ordinal prompt embeddings and the incidence softmax.
// Reference arithmetic: model/deepmil.py:136,200-204, model/vlsa.py:185-192, utils/func.py:44.
#pragma once
#include "common.cuh"

namespace vlsa {

// ------------------------------------------------------------------------------------------------
// merge of the per-chunk online-softmax partials of one bag (flash-decoding style, fixed order):
//   m = max_c m_c ; l = sum_c l_c e^{m_c-m} ; O_p = sum_c e^{m_c-m} O_c,p / l ; v = mean_p O_p
// grid (B, D/64), block (64, P): thread (x, y) = (feature column, prototype) folds the bag's entries in order — the loads of
// different entries are independent, so a thread keeps many in flight; the mean over P goes through shared memory in
// prototype order.  (The first version gave a thread all P prototypes of its column on a (B, 4) grid: one bag per call left
// 4 CTAs walking P x S dependent-latency loads, 20-40 us for work that is 3 us here.)
// With S > 0 the inputs are the level-1 partials of merge_fwd_split_kernel: bag b owns entries [b S, b S + S)
// (unused entries carry m = -inf, l = 0, O = 0); chunk_start then only tells whether the bag is empty.
__global__ void __launch_bounds__(64 * VLSA_MAX_P) merge_fwd_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                                     const float* __restrict__ part_O, const int* __restrict__ chunk_start,
                                                                     int P, int S, float* __restrict__ out_ml, float* __restrict__ out_O,
                                                                     float* __restrict__ out_v) {
    constexpr int D = VLSA_D;
    __shared__ float s_o[VLSA_MAX_P][64];
    const int b = blockIdx.x, col = threadIdx.x, p = threadIdx.y;
    const int d = blockIdx.y * 64 + col;
    const bool bag_empty = chunk_start[b + 1] == chunk_start[b];
    const int c0 = S > 0 ? b * S : chunk_start[b], c1 = S > 0 ? c0 + (bag_empty ? 0 : S) : chunk_start[b + 1];

    // a warp = 32 consecutive columns of ONE prototype: its lanes share the per-entry scalars (m, l, the scale factor), so each
    // lane fetches / computes them for one entry in 32 and they travel by shuffle; the O loads of eight entries are issued
    // together, the fold itself stays in entry order (bit-stable, and the same bits as a plain sequential loop)
    const int lane = col & 31;
    float mx = -INFINITY;
    for (int c = c0 + lane; c < c1; c += 32) mx = fmaxf(mx, __ldg(part_m + size_t(c) * P + p));
    mx = warp_max(mx);
    float o = 0.f, l = 0.f;
    for (int cb = c0; cb < c1; cb += 32) {
        const int c = cb + lane;
        float sf_mine = 0.f, l_mine = 0.f;
        if (c < c1) {
            const float mc = __ldg(part_m + size_t(c) * P + p);
            sf_mine = mc == -INFINITY ? 0.f : expf(mc - mx);
            l_mine = __ldg(part_l + size_t(c) * P + p);
        }
        const int nb = (c1 - cb) < 32 ? (c1 - cb) : 32;
        for (int k0 = 0; k0 < nb; k0 += 8) {
            float ov[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) ov[k] = (k0 + k < nb) ? part_O[(size_t(cb + k0 + k) * P + p) * D + d] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float sf = __shfl_sync(0xffffffffu, sf_mine, k0 + k), lv = __shfl_sync(0xffffffffu, l_mine, k0 + k);
                o = fmaf(sf, ov[k], o);
                l = fmaf(sf, lv, l);
            }
        }
    }
    const float op = (c1 > c0) ? o / l : 0.f;                 // empty bag: matmul over an empty N gives zeros
    if (out_O) out_O[(size_t(b) * P + p) * D + d] = op;
    s_o[p][col] = op;
    __syncthreads();
    if (p == 0 && out_v) {
        float vs = 0.f;
        for (int q = 0; q < P; ++q) vs += s_o[q][col];
        out_v[size_t(b) * D + d] = vs / float(P);             // torch.mean over P (deepmil.py:136)
    }
    if (blockIdx.y == 0 && col == 0) {
        out_ml[(size_t(b) * P + p) * 2 + 0] = mx;
        out_ml[(size_t(b) * P + p) * 2 + 1] = l;
    }
}

// ------------------------------------------------------------------------------------------------
// level 1 of the forward merge: CTA (split z, prototype p, bag b) folds chunks [c0 + z cps, c0 + (z + 1) cps) of its
// bag into ONE partial of the same (m, l, O) format, entry (b S + z) of the level-2 arrays.  Fixed order.
// grid (S, P, B), 128 threads (thread = 4 feature columns).
__global__ void __launch_bounds__(128) merge_fwd_split_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                              const float* __restrict__ part_O,
                                                              const int* __restrict__ chunk_start, int P, int S,
                                                              float* __restrict__ out_m, float* __restrict__ out_l,
                                                              float* __restrict__ out_O) {
    constexpr int D = VLSA_D;
    const int z = blockIdx.x, p = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
    const int c0 = chunk_start[b], c1 = chunk_start[b + 1];
    const int cps = (c1 - c0 + S - 1) / S;
    const int lo = c0 + z * cps, hi = (lo + cps) < c1 ? (lo + cps) : c1;
    float mx = -INFINITY;
    for (int c = lo; c < hi; ++c) mx = fmaxf(mx, __ldg(part_m + size_t(c) * P + p));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
#pragma unroll 4
    for (int c = lo; c < hi; ++c) {
        const float mc = __ldg(part_m + size_t(c) * P + p);
        const float sf = mc == -INFINITY ? 0.f : expf(mc - mx);
        const float4 o = *reinterpret_cast<const float4*>(part_O + (size_t(c) * P + p) * D + 4 * tid);
        acc.x = fmaf(sf, o.x, acc.x); acc.y = fmaf(sf, o.y, acc.y); acc.z = fmaf(sf, o.z, acc.z); acc.w = fmaf(sf, o.w, acc.w);
        l = fmaf(sf, __ldg(part_l + size_t(c) * P + p), l);
    }
    const size_t e = (size_t(b) * S + z) * P + p;
    *reinterpret_cast<float4*>(out_O + e * D + 4 * tid) = acc;
    if (tid == 0) { out_m[e] = mx; out_l[e] = l; }
}

// level 1 of the backward merge: out[z][p][:] = sum of part[c][p][:] over chunks c in split z (fixed order).
// grid (S, P), 128 threads (thread = 4 feature columns).
__global__ void __launch_bounds__(128) merge_bwd_split_kernel(const float* __restrict__ part, int total_chunks, int P, int S,
                                                              float* __restrict__ out) {
    constexpr int D = VLSA_D;
    const int z = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    const int cps = (total_chunks + S - 1) / S;
    const int lo = z * cps, hi = (lo + cps) < total_chunks ? (lo + cps) : total_chunks;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int c = lo;
    for (; c + 2 <= hi; c += 2) {
        const float4 o0 = *reinterpret_cast<const float4*>(part + (size_t(c) * P + p) * D + 4 * tid);
        const float4 o1 = *reinterpret_cast<const float4*>(part + (size_t(c + 1) * P + p) * D + 4 * tid);
        a0.x += o0.x; a0.y += o0.y; a0.z += o0.z; a0.w += o0.w;
        a1.x += o1.x; a1.y += o1.y; a1.z += o1.z; a1.w += o1.w;
    }
    if (c < hi) {
        const float4 o0 = *reinterpret_cast<const float4*>(part + (size_t(c) * P + p) * D + 4 * tid);
        a0.x += o0.x; a0.y += o0.y; a0.z += o0.z; a0.w += o0.w;
    }
    *reinterpret_cast<float4*>(out + (size_t(z) * P + p) * D + 4 * tid) =
        make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
}

// ------------------------------------------------------------------------------------------------
// f[b][o] = sum_i W[o][i] v[b][i] + bias[o]   (nn.Linear, deepmil.py:117,204)
// grid D/4, 128 threads: warp w owns output row o = 4*blockIdx.x + w for every bag; W row in registers,
// v staged through shared memory 8 bags at a time.
__global__ void __launch_bounds__(128) adapter_fwd_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                          const float* __restrict__ v, int B, float* __restrict__ f) {
    constexpr int D = VLSA_D, BT = 8;
    __shared__ __align__(16) float s_v[BT][D];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = blockIdx.x * 4 + warp;
    float4 w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(W + size_t(o) * D + j * 128 + lane * 4);
    const float bo = bias[o];
    for (int b0 = 0; b0 < B; b0 += BT) {
        const int nb = (B - b0) < BT ? (B - b0) : BT;
        __syncthreads();
        for (int i = tid; i < nb * D / 4; i += 128)
            reinterpret_cast<float4*>(&s_v[0][0])[i] = reinterpret_cast<const float4*>(v + size_t(b0) * D)[i];
        __syncthreads();
        for (int bb = 0; bb < nb; ++bb) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 x = *reinterpret_cast<const float4*>(&s_v[bb][j * 128 + lane * 4]);
                acc += w[j].x * x.x + w[j].y * x.y + w[j].z * x.z + w[j].w * x.w;
            }
            acc = warp_sum(acc);
            if (lane == 0) f[size_t(b0 + bb) * D + o] = acc + bo;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// g = f/||f|| ; Tn = T/||T|| ; logits = (exp(logit_scale) * g) @ Tn^T ; IF = softmax(logits)
// (model/vlsa.py:185-192, utils/func.py:44).  grid B, 256 threads.
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ f, const float* __restrict__ T, int R,
                                                       const float* __restrict__ logit_scale, float* __restrict__ out_g,
                                                       float* __restrict__ out_logits, float* __restrict__ out_if,
                                                       float* __restrict__ out_Tn) {
    constexpr int D = VLSA_D;
    __shared__ float s_g[D];
    __shared__ float s_red[32];
    __shared__ float s_logit[VLSA_MAX_R];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float ls = expf(*logit_scale);
    const float f0 = f[size_t(b) * D + tid], f1 = f[size_t(b) * D + 256 + tid];
    const float ss = block_sum(f0 * f0 + f1 * f1, s_red);
    const float inv = 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
    const float g0 = f0 * inv, g1 = f1 * inv;
    s_g[tid] = g0; s_g[256 + tid] = g1;
    out_g[size_t(b) * D + tid] = g0; out_g[size_t(b) * D + 256 + tid] = g1;
    __syncthreads();
    for (int r = warp; r < R; r += 8) {
        float tt = 0.f, dot = 0.f;
        float tv[D / 32];
#pragma unroll
        for (int k = 0; k < D / 32; ++k) { tv[k] = T[size_t(r) * D + k * 32 + lane]; tt += tv[k] * tv[k]; }
        tt = warp_sum(tt);
        const float tinv = 1.f / fmaxf(sqrtf(tt), VLSA_NORM_EPS);
#pragma unroll
        for (int k = 0; k < D / 32; ++k) {
            const float tn = tv[k] * tinv;
            dot += (ls * s_g[k * 32 + lane]) * tn;
            if (b == 0 && out_Tn) out_Tn[size_t(r) * D + k * 32 + lane] = tn;
        }
        dot = warp_sum(dot);
        if (lane == 0) { s_logit[r] = dot; out_logits[size_t(b) * R + r] = dot; }
    }
    __syncthreads();
    if (warp == 0 && out_if) {
        const float x = lane < R ? s_logit[lane] : -INFINITY;
        const float mx = warp_max(x);
        const float e = lane < R ? expf(x - mx) : 0.f;
        const float sum = warp_sum(e);
        if (lane < R) out_if[size_t(b) * R + lane] = e / sum;
    }
}

}  // namespace vlsa

// ================================================================================================
// Backward of the head (autograd of model/vlsa.py:185-192 + deepmil.py:117,136,204).
// ================================================================================================
namespace vlsa {

// per bag: dg = ls * dlogits @ Tn (+ external dg); df = (dg - g (g.dg)) / |f|; dls_part = sum_r dlogits_r logits_r
// grid B, 256 threads.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ f, const float* __restrict__ g,
                                                       const float* __restrict__ T, int R,
                                                       const float* __restrict__ logit_scale,
                                                       const float* __restrict__ logits,
                                                       const float* __restrict__ d_logits,
                                                       const float* __restrict__ d_g_ext,
                                                       const float* __restrict__ d_f_ext, float* __restrict__ df,
                                                       float* __restrict__ dls_part) {
    constexpr int D = VLSA_D;
    __shared__ float s_tinv[VLSA_MAX_R];
    __shared__ float s_dl[VLSA_MAX_R];
    __shared__ float s_red[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float ls = expf(*logit_scale);
    for (int r = warp; r < R; r += 8) {
        float tt = 0.f;
        for (int k = lane; k < D; k += 32) { const float t = T[size_t(r) * D + k]; tt += t * t; }
        tt = warp_sum(tt);
        if (lane == 0) s_tinv[r] = 1.f / fmaxf(sqrtf(tt), VLSA_NORM_EPS);
    }
    if (tid < R) s_dl[tid] = d_logits[size_t(b) * R + tid];
    __syncthreads();
    float dg[2], gv[2], fv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int d = tid + h * 256;
        float a = 0.f;
        for (int r = 0; r < R; ++r) a += s_dl[r] * (T[size_t(r) * D + d] * s_tinv[r]);
        dg[h] = ls * a + (d_g_ext ? d_g_ext[size_t(b) * D + d] : 0.f);
        gv[h] = g[size_t(b) * D + d];
        fv[h] = f[size_t(b) * D + d];
    }
    const float gdg = block_sum(gv[0] * dg[0] + gv[1] * dg[1], s_red);
    const float ff = block_sum(fv[0] * fv[0] + fv[1] * fv[1], s_red);
    const float inv = 1.f / fmaxf(sqrtf(ff), VLSA_NORM_EPS);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const size_t at = size_t(b) * D + tid + h * 256;
        df[at] = (dg[h] - gv[h] * gdg) * inv + (d_f_ext ? d_f_ext[at] : 0.f);
    }
    if (warp == 0) {
        float a = lane < R ? s_dl[lane] * logits[size_t(b) * R + lane] : 0.f;
        a = warp_sum(a);
        if (lane == 0) dls_part[b] = a;
    }
}

// dW[o][i] = sum_b df[b][o] v[b][i] ; db[o] = sum_b df[b][o].  grid D/8, 512 threads (thread = column i).
__global__ void __launch_bounds__(512) adapter_bwd_dw_kernel(const float* __restrict__ df, const float* __restrict__ v,
                                                             int B, float* __restrict__ dW, float* __restrict__ db) {
    constexpr int D = VLSA_D, OT = 8, BT = 32;
    __shared__ float s_df[BT][OT];
    const int i = threadIdx.x, o0 = blockIdx.x * OT;
    float acc[OT], accb = 0.f;
#pragma unroll
    for (int k = 0; k < OT; ++k) acc[k] = 0.f;
    for (int b0 = 0; b0 < B; b0 += BT) {
        const int nb = (B - b0) < BT ? (B - b0) : BT;
        __syncthreads();
        if (i < nb * OT) s_df[i / OT][i % OT] = df[size_t(b0 + i / OT) * D + o0 + i % OT];
        __syncthreads();
        for (int bb = 0; bb < nb; ++bb) {
            const float vv = v[size_t(b0 + bb) * D + i];
#pragma unroll
            for (int k = 0; k < OT; ++k) acc[k] += s_df[bb][k] * vv;
            if (i < OT) accb += s_df[bb][i];
        }
    }
#pragma unroll
    for (int k = 0; k < OT; ++k) dW[size_t(o0 + k) * D + i] = acc[k];
    if (i < OT) db[o0 + i] = accb;
}

// dv[b][i] = sum_o df[b][o] W[o][i].  grid D/4, 128 threads: a CTA owns 4 output columns i0 .. i0+3 for every
// bag; thread t holds W[t + 128 k][i0 .. i0+3] (k < 4) in registers, bags go 8 at a time through a transposed
// butterfly (32 values = 8 bags x 4 columns) and a 4-warp shared-memory sum.
__global__ void __launch_bounds__(128) adapter_bwd_dv_kernel(const float* __restrict__ df, const float* __restrict__ W,
                                                             int B, float* __restrict__ dv) {
    constexpr int D = VLSA_D, BT = 8;
    __shared__ float s_red[4][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, i0 = blockIdx.x * 4;
    float4 w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(W + size_t(tid + 128 * k) * D + i0);
    for (int b0 = 0; b0 < B; b0 += BT) {
        float v[32];
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + bb < B) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = __ldg(df + size_t(b0 + bb) * D + tid + 128 * k);
                    a.x = fmaf(d, w[k].x, a.x); a.y = fmaf(d, w[k].y, a.y); a.z = fmaf(d, w[k].z, a.z); a.w = fmaf(d, w[k].w, a.w);
                }
            }
            v[4 * bb + 0] = a.x; v[4 * bb + 1] = a.y; v[4 * bb + 2] = a.z; v[4 * bb + 3] = a.w;
        }
        const float tot = warp_reduce_transpose<32>(v);       // lane L: sum over the warp of value L
        __syncthreads();
        s_red[warp][lane] = tot;
        __syncthreads();
        if (warp == 0) {
            const float sum = (s_red[0][lane] + s_red[1][lane]) + (s_red[2][lane] + s_red[3][lane]);
            const int bb = lane >> 2, c = lane & 3;
            if (b0 + bb < B) dv[size_t(b0 + bb) * D + i0 + c] = sum;
        }
    }
}

// delta[b][p] = (dv_b . O_b,p) / P  (= dO_p . O_p with dO_p = dv / P).  grid B, 256 threads.
__global__ void __launch_bounds__(256) delta_kernel(const float* __restrict__ dv, const float* __restrict__ O, int P,
                                                    float* __restrict__ delta) {
    constexpr int D = VLSA_D;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = warp; p < P; p += 8) {
        float a = 0.f;
        for (int k = lane; k < D; k += 32) a += dv[size_t(b) * D + k] * O[(size_t(b) * P + p) * D + k];
        a = warp_sum(a);
        if (lane == 0) delta[size_t(b) * P + p] = a / float(P);
    }
}

// delta[b][p] = dO[b][p] . O[b][p] for a gradient row per prototype (any pooling over the P outputs).  grid B.
__global__ void __launch_bounds__(256) delta_gen_kernel(const float* __restrict__ dO, const float* __restrict__ O, int P,
                                                        float* __restrict__ delta) {
    constexpr int D = VLSA_D;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p = warp; p < P; p += 8) {
        const size_t at = (size_t(b) * P + p) * D;
        float a = 0.f;
        for (int k = lane; k < D; k += 32) a += dO[at + k] * O[at + k];
        a = warp_sum(a);
        if (lane == 0) delta[size_t(b) * P + p] = a;
    }
}

// dTn[r] = ls * sum_b dlogits[b][r] g[b] ; dT[r] = (dTn - Tn (Tn.dTn)) / |T_r| ; block 0 also reduces dls.
// grid R, 256 threads.
__global__ void __launch_bounds__(256) text_bwd_kernel(const float* __restrict__ T, int R, const float* __restrict__ g,
                                                       const float* __restrict__ d_logits, int B,
                                                       const float* __restrict__ logit_scale,
                                                       const float* __restrict__ dls_part, float* __restrict__ dT,
                                                       float* __restrict__ dls) {
    constexpr int D = VLSA_D;
    __shared__ float s_red[32];
    const int r = blockIdx.x, tid = threadIdx.x;
    const float ls = expf(*logit_scale);
    float acc[2] = {0.f, 0.f};
    for (int b = 0; b < B; ++b) {
        const float dl = d_logits[size_t(b) * R + r];
        acc[0] += dl * g[size_t(b) * D + tid];
        acc[1] += dl * g[size_t(b) * D + 256 + tid];
    }
    acc[0] *= ls; acc[1] *= ls;
    const float t0 = T[size_t(r) * D + tid], t1 = T[size_t(r) * D + 256 + tid];
    const float tt = block_sum(t0 * t0 + t1 * t1, s_red);
    const float inv = 1.f / fmaxf(sqrtf(tt), VLSA_NORM_EPS);
    const float tn0 = t0 * inv, tn1 = t1 * inv;
    const float proj = block_sum(tn0 * acc[0] + tn1 * acc[1], s_red);
    dT[size_t(r) * D + tid] = (acc[0] - tn0 * proj) * inv;
    dT[size_t(r) * D + 256 + tid] = (acc[1] - tn1 * proj) * inv;
    if (r == 0) {
        float a = 0.f;
        for (int b = tid; b < B; b += 256) a += dls_part[b];
        a = block_sum(a, s_red);
        if (tid == 0) *dls = a;
    }
}

// dQn[p] = sum over all chunks of the partials (fixed order); dQ[p] = (dQn - Qn (Qn.dQn)) / |Q_p|, or dQn itself
// when the rows of Q entered the scores as they are (prenorm).  grid P, 512 threads.
__global__ void __launch_bounds__(512) merge_bwd_kernel(const float* __restrict__ part, int total_chunks, int P,
                                                        const float* __restrict__ Q, float* __restrict__ dQ,
                                                        int prenorm = 0) {
    constexpr int D = VLSA_D;
    __shared__ float s_red[32];
    const int p = blockIdx.x, d = threadIdx.x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int c = 0;
    for (; c + 4 <= total_chunks; c += 4) {
        a0 += part[(size_t(c + 0) * P + p) * D + d];
        a1 += part[(size_t(c + 1) * P + p) * D + d];
        a2 += part[(size_t(c + 2) * P + p) * D + d];
        a3 += part[(size_t(c + 3) * P + p) * D + d];
    }
    for (; c < total_chunks; ++c) a0 += part[(size_t(c) * P + p) * D + d];
    const float dqn = (a0 + a1) + (a2 + a3);
    if (prenorm) { dQ[size_t(p) * D + d] = dqn; return; }      // block-uniform
    const float q = Q[size_t(p) * D + d];
    const float qq = block_sum(q * q, s_red);
    const float inv = 1.f / fmaxf(sqrtf(qq), VLSA_NORM_EPS);
    const float qn = q * inv;
    const float proj = block_sum(qn * dqn, s_red);
    dQ[size_t(p) * D + d] = (dqn - qn * proj) * inv;
}

}  // namespace vlsa
