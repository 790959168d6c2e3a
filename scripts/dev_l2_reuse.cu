// Dev harness (GPU box): can the backward of a bag group re-read its rows from the 126 MB L2 if it runs right after
// the group's forward?  Streams a 3.2 GB buffer in groups of S MB: pass A over the group (L2 policy: normal or
// evict_last), then pass B over the same group (evict_first), next group.  Reports the time of the whole A+B sweep
// against two full sweeps from HBM (S = whole buffer).  Launch overhead of the 2 launches per group is included.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/dev_l2_reuse scripts/dev_l2_reuse.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../vlsa_b200/csrc/common.cuh"
using namespace vlsa;

__device__ __forceinline__ uint64_t make_policy(int kind) {
    uint64_t pol;
    if (kind == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// persistent CTAs, 32 KB tiles dealt round-robin, 2-stage ring (the shape of agg_simt_kernel's load path)
template <int TILE, int S>
__global__ void __launch_bounds__(128) k_stream(const char* __restrict__ x, size_t bytes, int policy_kind, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[S];
    const long long ntiles = bytes / TILE;
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(full + s, 1); mbar_fence_init(); }
    __syncthreads();
    const uint64_t pol = make_policy(policy_kind);
    float acc = 0.f;
    long long tp = blockIdx.x; int issued = 0;
    auto produce = [&]() {
        if (tp >= ntiles) return;
        const int s = issued % S;
        mbar_expect_tx(full + s, TILE);
        bulk_g2s_evict_first(sm + s * TILE, x + size_t(tp) * TILE, TILE, full + s, pol);
        ++issued; tp += gridDim.x;
    };
    if (threadIdx.x == 0) for (int s = 0; s < S - 1; ++s) produce();
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        if (threadIdx.x == 0) produce();
        mbar_wait(full + (it % S), (it / S) & 1);
        acc += reinterpret_cast<const float*>(sm + (it % S) * TILE)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    const size_t total = size_t(3200) << 20;
    char* x; float* out;
    cudaMalloc(&x, total); cudaMalloc(&out, 4);
    cudaMemset(x, 1, total);
    constexpr int TILE = 32768, S = 2;
    auto kern = k_stream<TILE, S>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * S);
    const int grid = 148 * 3;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto sweep = [&](size_t group, int polA) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            for (size_t off = 0; off < total; off += group) {
                const size_t n = off + group <= total ? group : total - off;
                kern<<<grid, 128, TILE * S>>>(x + off, n, polA, out);
                kern<<<grid, 128, TILE * S>>>(x + off, n, 0, out);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        return best;
    };
    const float base = sweep(total, 0);
    printf("[l2_reuse] two full sweeps from HBM (3.2 GB each): %.1f us  = %.0f GB/s\n", base * 1e3, 2.0 * total / base / 1e6);
    for (int polA : {1, 2})
        for (int mb : {16, 24, 32, 48, 64, 80, 96, 112, 128, 192}) {
            const float ms = sweep(size_t(mb) << 20, polA);
            printf("[l2_reuse] group %3d MB, pass A policy %s: A+B sweep %.1f us = %.2f x of two HBM sweeps (%.0f launches)\n", mb,
                   polA == 1 ? "evict_normal" : "evict_last  ", ms * 1e3, ms / base, 2.0 * ((total + (size_t(mb) << 20) - 1) / (size_t(mb) << 20)));
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("[l2_reuse] %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
