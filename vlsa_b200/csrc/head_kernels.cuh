// Epilogue kernels of the aggregation path: split-N merge, visual adapter (nn.Linear 512x512),
// cosine head against the ordinal prompt embeddings and the incidence softmax.
// Reference arithmetic: model/deepmil.py:136,200-204, model/vlsa.py:185-192, utils/func.py:44.
#pragma once
#include "common.cuh"

namespace vlsa {

// ------------------------------------------------------------------------------------------------
// merge of the per-chunk online-softmax partials of one bag (flash-decoding style, fixed order):
//   m = max_c m_c ; l = sum_c l_c e^{m_c-m} ; O_p = sum_c e^{m_c-m} O_c,p / l ; v = mean_p O_p
// grid (B, D/128), 128 threads, thread = one feature column.
template <int P>
__global__ void __launch_bounds__(128) merge_fwd_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                        const float* __restrict__ part_O, const int* __restrict__ chunk_start,
                                                        float* __restrict__ out_ml, float* __restrict__ out_O,
                                                        float* __restrict__ out_v) {
    constexpr int D = VLSA_D, CB = 32;                 // chunks per batch
    __shared__ float s_mx[P];
    __shared__ float s_sf[CB][P];
    __shared__ float s_red[4][P];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.y * 128 + tid;
    const int c0 = chunk_start[b], c1 = chunk_start[b + 1];

    // pass 1: global max per p over the bag's chunks
    float mx[P];
#pragma unroll
    for (int p = 0; p < P; ++p) mx[p] = -INFINITY;
    for (int c = c0 + tid; c < c1; c += 128)
#pragma unroll
        for (int p = 0; p < P; ++p) mx[p] = fmaxf(mx[p], part_m[size_t(c) * P + p]);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const float v = warp_max(mx[p]);
        if (lane == 0) s_red[warp][p] = v;
    }
    __syncthreads();
    if (tid < P) s_mx[tid] = fmaxf(fmaxf(s_red[0][tid], s_red[1][tid]), fmaxf(s_red[2][tid], s_red[3][tid]));
    __syncthreads();

    float o[P], l[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { o[p] = 0.f; l[p] = 0.f; }
    for (int cb = c0; cb < c1; cb += CB) {
        const int nb = (c1 - cb) < CB ? (c1 - cb) : CB;
        __syncthreads();
        for (int i = tid; i < nb * P; i += 128) {
            const int c = i / P, p = i % P;
            s_sf[c][p] = expf(part_m[size_t(cb + c) * P + p] - s_mx[p]);
        }
        __syncthreads();
        for (int c = 0; c < nb; ++c) {
            const float* po = part_O + size_t(cb + c) * P * D + d;
            const float* pl = part_l + size_t(cb + c) * P;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float sf = s_sf[c][p];
                o[p] += sf * po[size_t(p) * D];
                l[p] += sf * pl[p];
            }
        }
    }
    float vs = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const float op = (c1 > c0) ? o[p] / l[p] : 0.f;      // empty bag: matmul over an empty N gives zeros
        if (out_O) out_O[(size_t(b) * P + p) * D + d] = op;
        vs += op;
    }
    out_v[size_t(b) * D + d] = vs / float(P);               // torch.mean over P (deepmil.py:136)
    if (blockIdx.y == 0 && tid < P) {
        out_ml[(size_t(b) * P + tid) * 2 + 0] = s_mx[tid];
        float lt = 0.f;
        for (int p = 0; p < P; ++p) if (p == tid) lt = l[p];
        out_ml[(size_t(b) * P + tid) * 2 + 1] = lt;
    }
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
