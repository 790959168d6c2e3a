// Secondary consumers of the path:
//  * attention read-out A[P,N] (ret_with_attn=True, model/deepmil.py:206-213; used by
//    utils/model_inference.py:118-131),
//  * the zero-shot arm: per-patch logits + logit pooling (model/vlsa.py:189-196, model/deepmil.py:16-37).
#pragma once
#include "common.cuh"

namespace vlsa {

template <typename XT>
__device__ __forceinline__ void load_row16(const XT* __restrict__ row, int lane, float (&x)[16]);

// lane holds x[j*128 + lane*4 + k], j=0..3, k=0..3  (fp32)
template <>
__device__ __forceinline__ void load_row16<float>(const float* __restrict__ row, int lane, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + j * 128 + lane * 4));
        x[4 * j + 0] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
    }
}
// lane holds x[j*256 + lane*8 + k], j=0..1, k=0..7  (bf16)
template <>
__device__ __forceinline__ void load_row16<__nv_bfloat16>(const __nv_bfloat16* __restrict__ row, int lane, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(row + j * 256 + lane * 8));
        x[8 * j + 0] = bf16_lo(u.x); x[8 * j + 1] = bf16_hi(u.x); x[8 * j + 2] = bf16_lo(u.y); x[8 * j + 3] = bf16_hi(u.y);
        x[8 * j + 4] = bf16_lo(u.z); x[8 * j + 5] = bf16_hi(u.z); x[8 * j + 6] = bf16_lo(u.w); x[8 * j + 7] = bf16_hi(u.w);
    }
}
// feature index of register slot i for the two layouts above
template <typename XT>
__device__ __forceinline__ int slot_col(int lane, int i) {
    return sizeof(XT) == 4 ? (i >> 2) * 128 + lane * 4 + (i & 3) : (i >> 3) * 256 + lane * 8 + (i & 7);
}

// Normalise `nq` rows of `src` [nq, D] into shared memory `dst` (row-major), whole block cooperates.
__device__ __forceinline__ void load_normalized_rows(const float* __restrict__ src, int nq, float* dst,
                                                     bool prenorm = false) {
    constexpr int D = VLSA_D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < nq; r += nw) {
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = __ldg(src + size_t(r) * D + d); ss += v * v; }
        ss = warp_sum(ss);
        const float inv = prenorm ? 1.f : 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
        for (int d = lane; d < D; d += 32) dst[r * D + d] = __ldg(src + size_t(r) * D + d) * inv;
    }
}

// cos[n][q] * mult for one bag; block = 256 threads = 8 warps x 4 rows = 32 rows per block.
// MODE 0: A[q][n] = exp(scale*cos - m_q) / l_q          (attention read-out, out is [nq, N])
// MODE 1: out[n][q] = mult * cos                        (per-patch logits, out is [N, nq])
// MODE 2: A[q][n] = softmax over q of scale*cos          (utils/model_inference.py:104-113, axis_softmax='L')
template <typename XT, int MODE>
__global__ void __launch_bounds__(256) row_cosine_kernel(const XT* __restrict__ X, long long N, const float* __restrict__ Qsrc,
                                                         int nq, float mult, const float* __restrict__ mult_log,
                                                         const float* __restrict__ ml, float* __restrict__ out,
                                                         int q_prenorm = 0) {
    constexpr int D = VLSA_D;
    extern __shared__ __align__(16) float s_q[];            // [nq][D] + [32][nq+1] staging
    float* s_out = s_q + size_t(nq) * D;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_normalized_rows(Qsrc, nq, s_q, q_prenorm != 0);
    __syncthreads();
    if (mult_log) mult = expf(*mult_log);
    const long long row0 = (long long)blockIdx.x * 32;
    for (int rr = 0; rr < 4; ++rr) {
        const long long n = row0 + warp * 4 + rr;
        if (n >= N) break;                                   // warp-uniform
        float x[16];
        load_row16<XT>(X + n * D, lane, x);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) ss += x[i] * x[i];
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), VLSA_NORM_EPS);
        for (int q = 0; q < nq; ++q) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) a += x[i] * s_q[q * D + slot_col<XT>(lane, i)];
            a = warp_sum(a);
            if (lane == 0) s_out[(warp * 4 + rr) * (nq + 1) + q] = a * inv;
        }
    }
    __syncthreads();
    const int nrows = (N - row0) < 32 ? int(N - row0) : 32;
    if (MODE == 0) {
        for (int i = tid; i < nq * 32; i += 256) {
            const int q = i >> 5, r = i & 31;
            if (r < nrows)
                out[size_t(q) * N + row0 + r] = expf(mult * s_out[r * (nq + 1) + q] - ml[q * 2]) / ml[q * 2 + 1];
        }
    } else if (MODE == 2) {
        for (int r = tid; r < nrows; r += 256) {
            float mx = -INFINITY, sum = 0.f;
            for (int q = 0; q < nq; ++q) mx = fmaxf(mx, mult * s_out[r * (nq + 1) + q]);
            for (int q = 0; q < nq; ++q) sum += expf(mult * s_out[r * (nq + 1) + q] - mx);
            for (int q = 0; q < nq; ++q) out[size_t(q) * N + row0 + r] = expf(mult * s_out[r * (nq + 1) + q] - mx) / sum;
        }
    } else {
        for (int i = tid; i < nrows * nq; i += 256) {
            const int r = i / nq, q = i % nq;
            out[size_t(row0 + r) * nq + q] = (mult * s_out[r * (nq + 1) + q]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Interpretation path (utils/model_inference.py:115-131), without a second pass over X: the attention rows sum
// to one, so  cottn_score @ ((visual_adapter(X) / |f|) @ Tn^T)  =  (W O_p + b) . Tn_r / |f|  with the per-prototype
// pooled features O_p the forward already produced.  grid B*P, 256 threads; out_sim is [B, P, R].
__global__ void __launch_bounds__(256) interp_sim_kernel(const float* __restrict__ O, const float* __restrict__ W,
                                                         const float* __restrict__ bias, const float* __restrict__ T, int R,
                                                         const float* __restrict__ f, int P, float* __restrict__ out_sim) {
    constexpr int D = VLSA_D;
    __shared__ __align__(16) float s_o[D];
    __shared__ float s_e[D];
    __shared__ float s_red[32];
    const int bp = blockIdx.x, b = bp / P, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int d = tid; d < D; d += 256) s_o[d] = O[size_t(bp) * D + d];
    float ff = 0.f;
    for (int d = tid; d < D; d += 256) { const float v = f[size_t(b) * D + d]; ff += v * v; }
    ff = block_sum(ff, s_red);                                   // also orders the s_o writes
    const float invL = 1.f / sqrtf(ff);                          // image_feature / L_image_feature: no eps in the reference
    for (int o = warp; o < D; o += 8) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + size_t(o) * D + j * 128 + lane * 4));
            const float4 x = *reinterpret_cast<const float4*>(s_o + j * 128 + lane * 4);
            a += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
        }
        a = warp_sum(a);
        if (lane == 0) s_e[o] = a + bias[o];
    }
    __syncthreads();
    for (int r = warp; r < R; r += 8) {
        float dot = 0.f, tt = 0.f;
        for (int d = lane; d < D; d += 32) { const float t = __ldg(T + size_t(r) * D + d); dot += t * s_e[d]; tt += t * t; }
        dot = warp_sum(dot); tt = warp_sum(tt);
        if (lane == 0) out_sim[size_t(bp) * R + r] = dot / fmaxf(sqrtf(tt), VLSA_NORM_EPS) * invL;
    }
}

// decoupled_imp = softmax over P of ls*sim ; probs_2 = softmax over R of ls*mean_P(sim).  grid B, 32*ceil(R/32) threads.
__global__ void __launch_bounds__(VLSA_MAX_R) interp_softmax_kernel(const float* __restrict__ sim, int P, int R,
                                                                    const float* __restrict__ logit_scale,
                                                                    float* __restrict__ out_imp, float* __restrict__ out_probs) {
    __shared__ float s_mean[VLSA_MAX_R];
    const int b = blockIdx.x, r = threadIdx.x;
    const float ls = expf(*logit_scale);
    const float* s = sim + size_t(b) * P * R;
    if (r < R) {
        float mx = -INFINITY, sum = 0.f, mean = 0.f;
        for (int p = 0; p < P; ++p) { mx = fmaxf(mx, ls * s[p * R + r]); mean += s[p * R + r]; }
        for (int p = 0; p < P; ++p) sum += expf(ls * s[p * R + r] - mx);
        for (int p = 0; p < P; ++p) out_imp[(size_t(b) * P + p) * R + r] = expf(ls * s[p * R + r] - mx) / sum;
        s_mean[r] = ls * (mean / float(P));
    }
    __syncthreads();
    if (r < R) {
        float mx = -INFINITY, sum = 0.f;
        for (int k = 0; k < R; ++k) mx = fmaxf(mx, s_mean[k]);
        for (int k = 0; k < R; ++k) sum += expf(s_mean[k] - mx);
        out_probs[size_t(b) * R + r] = expf(s_mean[r] - mx) / sum;
    }
}

// ------------------------------------------------------------------------------------------------
// logit pooling over N per class (deepmil.py:16-37).  grid R, 1024 threads.
//   mode 0: mean over N;  mode 1: mean of the top-min(k,N) values (values only: ties are harmless).
// Then (last block done) preds = argmax_r pooled (first maximal index, torch.argmax semantics).
__global__ void __launch_bounds__(1024) logit_pool_kernel(const float* __restrict__ logits, long long N, int R, int mode,
                                                          int k, float* __restrict__ pooled, long long* __restrict__ pred,
                                                          unsigned int* __restrict__ done_counter) {
    __shared__ float s_val[32];
    __shared__ long long s_idx[32];
    __shared__ float s_top[64];
    __shared__ long long s_sel[64];
    __shared__ bool s_last;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float result;
    if (mode == 0) {
        float a = 0.f;
        for (long long n = tid; n < N; n += 1024) a += logits[n * R + r];
        a = block_sum(a, s_val);
        result = a / float(N);
    } else {
        const int kk = (long long)k < N ? k : int(N);
        for (int j = 0; j < kk; ++j) {
            float best = -INFINITY; long long bi = -1;
            for (long long n = tid; n < N; n += 1024) {
                bool taken = false;
                for (int q = 0; q < j; ++q) taken |= (s_sel[q] == n);
                const float v = logits[n * R + r];
                if (!taken && (v > best || bi < 0)) { best = v; bi = n; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
            }
            __syncthreads();
            if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
            __syncthreads();
            if (warp == 0) {
                best = s_val[lane]; bi = s_idx[lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
                }
                if (lane == 0) { s_top[j] = best; s_sel[j] = bi; }
            }
            __syncthreads();
        }
        float a = 0.f;
        for (int j = 0; j < kk; ++j) a += s_top[j];        // descending order, like values.mean(dim=0)
        result = a / float(kk);
    }
    if (tid == 0) {
        pooled[r] = result;
        __threadfence();
        const unsigned int prev = atomicAdd(done_counter, 1u);
        s_last = (prev == unsigned(R - 1));
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        int best = 0; float bv = *((volatile float*)pooled);
        for (int q = 1; q < R; ++q) { const float v = ((volatile float*)pooled)[q]; if (v > bv) { bv = v; best = q; } }
        *pred = best;
        *done_counter = 0;          // re-arm for the next call on this workspace
    }
}

// ------------------------------------------------------------------------------------------------
// Zero-shot arm, the remaining branches of FeatMIL.forward + VLSA.forward (model/deepmil.py:51-67, model/vlsa.py:188-198):
//  * pooling 'mean' | 'max': column-wise mean / max over the N rows -> one [512] vector that then goes through the cosine
//    head (head_fwd_kernel).  Two fixed-order levels: `G` CTAs each fold rows blockIdx.x, blockIdx.x + G, ... (thread =
//    4 columns, sequential over rows), then one CTA folds the G partials in index order.  Bit-stable.
//  * image_features of the logit-pooling modes: the N normalised patches (F.normalize, model/vlsa.py:189).
template <typename XT>
__device__ __forceinline__ float4 load_x4(const XT* p);
template <> __device__ __forceinline__ float4 load_x4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 load_x4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}

template <typename XT, int MODE /* 0 mean (sum), 1 max */>
__global__ void __launch_bounds__(128) feat_pool_partial_kernel(const XT* __restrict__ X, long long N, float* __restrict__ part) {
    constexpr int D = VLSA_D;
    const int c = threadIdx.x * 4;
    float4 a = MODE == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (long long n = blockIdx.x; n < N; n += gridDim.x) {
        const float4 v = load_x4<XT>(X + n * D + c);
        if (MODE == 0) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        else { a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w); }
    }
    *reinterpret_cast<float4*>(part + size_t(blockIdx.x) * D + c) = a;
}
template <int MODE>
__global__ void __launch_bounds__(128) feat_pool_final_kernel(const float* __restrict__ part, int G, long long N, float* __restrict__ out) {
    constexpr int D = VLSA_D;
    const int c = threadIdx.x * 4;
    float4 a = *reinterpret_cast<const float4*>(part + c);
    for (int g = 1; g < G; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(part + size_t(g) * D + c);
        if (MODE == 0) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        else { a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w); }
    }
    if (MODE == 0) { const float inv = 1.f / float(N); a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv; }
    *reinterpret_cast<float4*>(out + c) = a;
}
// out[n] = x_n / max(|x_n|, eps) in fp32; one warp per row, 8 rows per CTA
template <typename XT>
__global__ void __launch_bounds__(256) row_normalize_kernel(const XT* __restrict__ X, long long N, float* __restrict__ out) {
    constexpr int D = VLSA_D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * 8 + warp;
    if (n >= N) return;
    float x[16];
    load_row16<XT>(X + n * D, lane, x);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) ss += x[i] * x[i];
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), VLSA_NORM_EPS);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[n * D + slot_col<XT>(lane, i)] = x[i] / nrm;
}

// dX of the pooled aggregation for callers that train something IN FRONT of it (VLFAN's feat_proj, a Linear + LayerNorm
// over all patch rows, model/layers.py:65-82 + model/deepmil.py:176-179).  With O_p = sum_n A_pn x_n and
// s_pn = scale qdir_p . x_n / |x_n|:
//   dx_n = sum_p ( A_pn dO_p + (scale dS_pn / |x_n|) qdir_p ) - (sum_p dS_pn s_pn) x_n / |x_n|^2,   dS_pn = A_pn (dO_p . x_n - dO_p . O_p)
// (no projection term for rows whose norm is clamped at eps, as F.normalize's backward).  The 2 P rows (qdir, dO_b) of
// the bag sit in shared memory; a warp works on FOUR rows at a time (lane = 16 columns of each) so that every 16-byte
// read of a qdir / dO row serves four patch rows in both of its uses (dot product, then accumulation) — with one row
// per warp the kernel re-read 2 x 2 P x 2 KB of shared memory per 2 KB row and was bound by that.  Packed fp32x2 FMAs.
// grid (ceil(max_rows / 128), B), 128 threads.
__global__ void __launch_bounds__(128) agg_dx_kernel(const float* __restrict__ X, const long long* __restrict__ cu_rows,
                                                     const float* __restrict__ Q, int P, int q_prenorm, float scale,
                                                     const float* __restrict__ ml, const float* __restrict__ O,
                                                     const float* __restrict__ dO, float* __restrict__ dX) {
    constexpr int D = VLSA_D, ROWS = 128, RW = 4;
    extern __shared__ __align__(16) float s_dx[];            // [P][D] qdir | [P][D] dO_b | [P] m | [P] 1/l | [P] delta
    float* s_qd = s_dx;
    float* s_g = s_dx + size_t(P) * D;
    float* s_mm = s_g + size_t(P) * D;
    float* s_il = s_mm + P;
    float* s_dl = s_il + P;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = cu_rows[b] + (long long)blockIdx.x * ROWS, rend = cu_rows[b + 1];
    if (r0 >= rend) return;                                    // block-uniform
    load_normalized_rows(Q, P, s_qd, q_prenorm != 0);
    for (int i = tid; i < P * D; i += 128) s_g[i] = __ldg(dO + size_t(b) * P * D + i);
    if (tid < P) {
        s_mm[tid] = ml[(size_t(b) * P + tid) * 2];
        s_il[tid] = 1.f / ml[(size_t(b) * P + tid) * 2 + 1];
    }
    for (int p = warp; p < P; p += 4) {                        // delta_p = dO_p . O_p
        const size_t at = (size_t(b) * P + p) * D;
        float a = 0.f;
        for (int k = lane; k < D; k += 32) a += __ldg(dO + at + k) * __ldg(O + at + k);
        a = warp_sum(a);
        if (lane == 0) s_dl[p] = a;
    }
    __syncthreads();
    const long long r1 = r0 + ROWS < rend ? r0 + ROWS : rend;
    for (long long n0 = r0 + warp * RW; n0 < r1; n0 += 4 * RW) {
        float2 x[RW][8], dx[RW][8];
        float inv[RW], tsum[RW];
        bool clamped[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const bool live = n0 + r < r1;                     // warp-uniform
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = live ? __ldg(reinterpret_cast<const float4*>(X + (n0 + r) * D + j * 128 + lane * 4))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                x[r][2 * j] = make_float2(v.x, v.y);
                x[r][2 * j + 1] = make_float2(v.z, v.w);
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s2 = __ffma2_rn(x[r][i], x[r][i], s2); dx[r][i] = make_float2(0.f, 0.f); }
            const float nr = sqrtf(warp_sum(s2.x + s2.y));
            clamped[r] = nr < VLSA_NORM_EPS;
            inv[r] = 1.f / fmaxf(nr, VLSA_NORM_EPS);
            tsum[r] = 0.f;
        }
        for (int p = 0; p < P; ++p) {
            float2 q2[8], g2[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 q4 = *reinterpret_cast<const float4*>(s_qd + p * D + j * 128 + lane * 4);
                const float4 g4 = *reinterpret_cast<const float4*>(s_g + p * D + j * 128 + lane * 4);
                q2[2 * j] = make_float2(q4.x, q4.y); q2[2 * j + 1] = make_float2(q4.z, q4.w);
                g2[2 * j] = make_float2(g4.x, g4.y); g2[2 * j + 1] = make_float2(g4.z, g4.w);
            }
            float qa[RW], ga[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                float2 aq = make_float2(0.f, 0.f), ag = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) { aq = __ffma2_rn(q2[i], x[r][i], aq); ag = __ffma2_rn(g2[i], x[r][i], ag); }
                qa[r] = aq.x + aq.y;
                ga[r] = ag.x + ag.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {                 // eight interleaved warp reductions
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    qa[r] += __shfl_xor_sync(0xffffffffu, qa[r], o);
                    ga[r] += __shfl_xor_sync(0xffffffffu, ga[r], o);
                }
            }
            const float mp = s_mm[p], ilp = s_il[p], dlp = s_dl[p];
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                const float sv = scale * (qa[r] * inv[r]);
                const float a = expf(sv - mp) * ilp;                           // A_pn (deepmil.py:198)
                const float ds = a * (ga[r] - dlp);
                const float c = scale * ds * inv[r];
                tsum[r] += ds * sv;
                const float2 a2 = make_float2(a, a), c2 = make_float2(c, c);
#pragma unroll
                for (int i = 0; i < 8; ++i) dx[r][i] = __ffma2_rn(c2, q2[i], __ffma2_rn(a2, g2[i], dx[r][i]));
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            if (n0 + r >= r1) break;                           // warp-uniform
            const float k = clamped[r] ? 0.f : -tsum[r] * inv[r] * inv[r];
            const float2 k2 = make_float2(k, k);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 lo = __ffma2_rn(k2, x[r][2 * j], dx[r][2 * j]), hi = __ffma2_rn(k2, x[r][2 * j + 1], dx[r][2 * j + 1]);
                *reinterpret_cast<float4*>(dX + (n0 + r) * D + j * 128 + lane * 4) = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        }
    }
}

}  // namespace vlsa
