// Fused survival objective on [B,R] logits: softmax -> SurvIFMLE + SurvEMD, value and d/dlogits.
// Reference: utils/func.py:44, loss/loss_surv.py:144-169, loss/loss_surv_ext.py:13-109,
// runner/vlsa_handler.py:241-258.  One warp per sample, lane = time bin.
#pragma once
#include "common.cuh"

namespace vlsa {

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
// suffix sum: sum over lanes >= lane
__device__ __forceinline__ float warp_suffix_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += n;
    }
    return v;
}

__global__ void __launch_bounds__(128) surv_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ t_,
                                                        const long long* __restrict__ e_, int B, int R,
                                                        const float* __restrict__ logit_scale, float w_ifmle, float w_emd,
                                                        float alpha, float eps, float inv_norm, int input_is_prob,
                                                        float* __restrict__ out_if, float* __restrict__ out_dlogits,
                                                        float* __restrict__ per_sample) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + warp;
    if (i >= B) return;
    const bool live = lane < R;
    const float ls = expf(*logit_scale);                      // get_logit_scale().detach()
    const int t = int(t_[i]);
    const float e = float(e_[i]);
    const float c = 1.f - e;

    // incidence = softmax(logits)
    const float x = live ? logits[size_t(i) * R + lane] : -INFINITY;
    const float mx = warp_max(x);
    const float ex = live ? expf(x - mx) : 0.f;
    const float sm = ex / warp_sum(ex);
    // input_is_prob: the caller already applied the output converter (reference loss-module signature)
    const float p = input_is_prob ? (live ? x : 0.f) : sm;
    if (out_if && live) out_if[size_t(i) * R + lane] = p;

    // ---- SurvIFMLE (loss_surv.py:153-160)
    const float cif = warp_incl_scan(p, lane);
    const float pt = __shfl_sync(0xffffffffu, p, t);
    // 1 - CIF(t): with a softmax input sum(p) = 1, so the tail sum over r > t is the same number without the
    // cancellation of 1 - cumsum (closer to the fp64 run of the reference when CIF(t) -> 1)
    const float tail = warp_sum((live && lane > t) ? p : 0.f);
    const float surv_t = input_is_prob ? 1.f - __shfl_sync(0xffffffffu, cif, t) : tail;
    const float unc = -(1.f - c) * logf(fmaxf(pt, eps));
    const float cen = -c * logf(fmaxf(surv_t, eps));
    const float l1 = (1.f - alpha) * (cen + unc) + alpha * unc;
    float dp1 = 0.f;                                          // d l1 / d p_lane
    if (live) {
        if (lane == t && pt >= eps) dp1 += -(1.f - c) / pt;                     // coefficient (1-a)+a = 1
        if (lane <= t && surv_t >= eps) dp1 += (1.f - alpha) * c / surv_t;
    }

    // ---- SurvEMD (loss_surv_ext.py:42-55, 81-102), p=2 raw distance
    const float target = (lane == t) ? 1.f : ((lane > t) ? (1.f - e) : 0.f);
    const float tl = live ? (2.f * target - 1.f) * ls : -INFINITY;
    const float tmx = warp_max(tl);
    const float tex = live ? expf(tl - tmx) : 0.f;
    const float tdist = tex / warp_sum(tex);
    const float mix = (1.f - e) * (1.f - target) + e;         // d pred / d p
    const float pred = live ? ((1.f - e) * ((1.f - target) * p + target * ls) + e * p) : -INFINITY;
    const float pmx = warp_max(pred);
    const float pex = live ? expf(pred - pmx) : 0.f;
    const float pdist = pex / warp_sum(pex);
    const float cdf_p = warp_incl_scan(pdist, lane), cdf_t = warp_incl_scan(tdist, lane);   // all lanes shuffle
    const float diff = live ? (cdf_p - cdf_t) : 0.f;
    const float l2 = warp_sum(diff * diff);
    const float G = warp_suffix_sum(2.f * diff, lane);        // d l2 / d pdist_lane
    const float gp = warp_sum(pdist * G);
    const float dpred = pdist * (G - gp);
    const float dp2 = live ? dpred * mix : 0.f;

    // ---- back through the incidence softmax, mean reduction folded in (inv_norm)
    const float dp = w_ifmle * dp1 + w_emd * dp2;
    const float pd = warp_sum(p * dp);
    if (out_dlogits && live)
        out_dlogits[size_t(i) * R + lane] = (input_is_prob ? dp : p * (dp - pd)) * inv_norm;
    if (lane == 0) { per_sample[size_t(i) * 2] = l1; per_sample[size_t(i) * 2 + 1] = l2; }
}

// out_loss = (w1*sum(l1) + w2*sum(l2), sum(l1), sum(l2)) * inv_norm, fixed summation order.  1 block, 256 threads.
__global__ void __launch_bounds__(256) surv_loss_reduce_kernel(const float* __restrict__ per_sample, int B, float w_ifmle,
                                                               float w_emd, float inv_norm, float* __restrict__ out_loss) {
    __shared__ float s_red[32];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) { a += per_sample[size_t(i) * 2]; b += per_sample[size_t(i) * 2 + 1]; }
    a = block_sum(a, s_red);
    b = block_sum(b, s_red);
    if (threadIdx.x == 0) {
        out_loss[1] = a * inv_norm;
        out_loss[2] = b * inv_norm;
        out_loss[0] = w_ifmle * (a * inv_norm) + w_emd * (b * inv_norm);
    }
}

}  // namespace vlsa
