// One Adam step over every trainable tensor of the path in ONE launch (runner/vlsa_handler.py:283 `optimizer.step()` with the
// optimizer optim/optim_factory.py:25-37 builds for cfg_vlsa_conch.yaml:111-118: torch.optim.Adam, L2 weight decay on the
// matrices only).  The gradients live in the flat all-reduce bucket (runner/dist.py FlatBucket), the moments in two buffers of
// the same layout; a parameter nobody's backward reached this step (reduced flag == 0) is skipped entirely — moments, step
// count and weight decay untouched, which is what `grad = None` means to torch.optim.Adam — decided on the DEVICE, so the
// step needs no device -> host read.
#pragma once
#include "common.cuh"

namespace vlsa {

struct AdamSeg {            // one parameter tensor (device array, 32 bytes)
    float* param;
    long long offset;       // first float of its gradient / moments in the flat buffers
    long long n;
    float weight_decay;
    float lr;
};

// grid (ceil(max n / (256 * 4)), S) x 256.  step_in[s]: optimizer steps parameter s has taken so far (float, on the device);
// block x = 0 of every segment writes the new count to step_out[s] — a DIFFERENT array (the caller ping-pongs two), so no
// block of this launch can read a count another block has already bumped.
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamSeg* __restrict__ segs, const float* __restrict__ grads,
                                                        float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                        const float* __restrict__ step_in, float* __restrict__ step_out,
                                                        const float* __restrict__ flags, float beta1, float beta2, float eps) {
    const int s = blockIdx.y;
    const bool skip = flags && flags[s] == 0.f;
    const float* step_count = step_in;
    if (blockIdx.x == 0 && threadIdx.x == 0) step_out[s] = step_in[s] + (skip ? 0.f : 1.f);
    if (skip) return;
    const AdamSeg sg = segs[s];
    const long long i0 = (long long)(blockIdx.x) * 1024 + threadIdx.x * 4;
    if (i0 >= sg.n) return;
    const double step = double(step_count[s]) + 1.0;
    const float bc1 = float(1.0 - pow(double(beta1), step));
    const float bc2_sqrt = sqrtf(float(1.0 - pow(double(beta2), step)));
    const float step_size = sg.lr / bc1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = i0 + k;
        if (i >= sg.n) break;
        float p = sg.param[i];
        float g = grads[sg.offset + i];
        if (sg.weight_decay != 0.f) g = fmaf(p, sg.weight_decay, g);          // grad = grad + param * weight_decay
        float m = exp_avg[sg.offset + i], v = exp_avg_sq[sg.offset + i];
        m = m + (g - m) * (1.f - beta1);                                       // exp_avg.lerp_(grad, 1 - beta1)
        v = beta2 * v + (1.f - beta2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p -= step_size * (m / denom);
        sg.param[i] = p;
        exp_avg[sg.offset + i] = m;
        exp_avg_sq[sg.offset + i] = v;
    }
}

}  // namespace vlsa
