// Dev harness (GPU box): read-only streaming bandwidth of this B200 for the access patterns the aggregation
// kernels use.  Establishes what "HBM-bound" can mean for a kernel that only READS X.
//   ldg   : grid-stride LDG.128 (evict-first), U loads in flight per thread
//   bulk  : persistent CTAs, cp.async.bulk tiles of TILE bytes into an S-stage smem ring, contiguous chunk per CTA
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../vlsa_b200/csrc/tc_common.cuh"
using namespace vlsa;

template <int U>
__global__ void __launch_bounds__(512) k_ldg(const float4* __restrict__ x, size_t n4, float* out) {
    const uint64_t pol = make_evict_first_policy();
    float acc = 0.f;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream_f4(reinterpret_cast<const float*>(x + i + u * stride), pol);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// each CTA streams a contiguous span with an S-stage ring of TILE-byte bulk copies; consumers just touch one word
template <int TILE, int S>
__global__ void __launch_bounds__(128) k_bulk(const char* __restrict__ x, size_t bytes, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[S];
    const size_t per = (bytes / gridDim.x) / TILE * TILE;
    const char* base = x + size_t(blockIdx.x) * per;
    const int ntiles = int(per / TILE);
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(full + s, 1); mbar_fence_init(); }
    __syncthreads();
    const uint64_t pol = make_evict_first_policy();
    float acc = 0.f;
    if (threadIdx.x == 0)
        for (int s = 0; s < S - 1 && s < ntiles; ++s) { mbar_expect_tx(full + s, TILE); bulk_g2s_evict_first(sm + s * TILE, base + size_t(s) * TILE, TILE, full + s, pol); }
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        if (threadIdx.x == 0 && t + S - 1 < ntiles) {
            const int s2 = (t + S - 1) % S;
            mbar_expect_tx(full + s2, TILE);
            bulk_g2s_evict_first(sm + s2 * TILE, base + size_t(t + S - 1) * TILE, TILE, full + s2, pol);
        }
        mbar_wait(full + s, (t / S) & 1);
        acc += reinterpret_cast<const float*>(sm + s * TILE)[threadIdx.x];
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
}

// register-path model of the tensor-core producers: 1 CTA/SM, 8 warps, each warp loads 4 rows (2 KB each) of a
// 32-row tile with 16 LDG.128 per lane, consumes them, __syncthreads-free; MODE 0 = no prefetch,
// 1 = one 64 KB cp.async.bulk.prefetch.L2 per tile issued PF tiles ahead by a 9th warp (paced by a smem counter),
// 2 = same with per-line prefetch.global.L2; FENCE = 1 adds the proxy fence after each tile (as the real kernel)
template <int KIND>
__device__ __forceinline__ float4 ld_kind(const float* p, uint64_t pol) {
    float4 v;
    if (KIND == 0) return ldg_stream_f4(p, pol);
    if (KIND == 1) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (KIND == 2) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (KIND == 3) asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (KIND == 4) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (KIND == 5) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (KIND == 6) asm volatile("ld.global.relaxed.gpu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int MODE, int FENCE, int KIND = 0>
__global__ void __launch_bounds__(288) k_regpath(const float* __restrict__ x, size_t bytes, int PF, float* out) {
    __shared__ volatile uint32_t prog;
    __shared__ float sink[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per = (bytes / gridDim.x) / 65536 * 65536;
    const float* base = x + size_t(blockIdx.x) * per / 4;
    const int ntiles = int(per / 65536);
    if (threadIdx.x == 0) prog = 0;
    __syncthreads();
    if (warp == 8) {
        if (MODE == 0) return;
        int done = 0;
        while (done < ntiles) {
            if (done < int(prog) + PF + 1) {
                if (MODE == 1) { if (lane == 0) l2_prefetch_bulk(base + size_t(done) * 16384, 65536); }
                else for (int l = lane; l < 512; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + size_t(done) * 16384 + l * 32));
                ++done;
            } else __nanosleep(128);
        }
        return;
    }
    const uint64_t pol = make_evict_first_policy();
    float acc = 0.f;
    float4 buf[16];
    auto issue = [&](int t) {
        const float* src = base + size_t(t) * 16384 + (4 * warp) * 512 + 4 * lane;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) buf[4 * j + i] = ld_kind<KIND>(src + j * 512 + 128 * i, pol);
    };
    issue(0);
    for (int t = 0; t < ntiles; ++t) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += buf[k].x * buf[k].y + buf[k].z * buf[k].w;
        if (t + 1 < ntiles) issue(t + 1);
        sink[threadIdx.x] = acc;
        if (FENCE) fence_proxy_async_smem();
        if (warp == 0 && lane == 0) prog = t + 1;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static float time_ms(F f, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(0); f(1); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f(i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main() {
    const size_t bytes = size_t(32) * 50000 * 2048;   // 3.28 GB, as one bench step
    char* buf[2]; float* out;
    cudaMalloc(&buf[0], bytes); cudaMalloc(&buf[1], bytes); cudaMalloc(&out, 4);
    cudaMemset(buf[0], 1, bytes); cudaMemset(buf[1], 2, bytes);
    auto report = [&](const char* name, float ms) { printf("%-34s %8.1f us  %7.0f GB/s\n", name, ms * 1e3, bytes / ms * 1e-6); fflush(stdout); };
    for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
        char nm[64];
        snprintf(nm, 64, "ldg U=4  %d CTAs x 512", ctas);
        report(nm, time_ms([&](int i) { k_ldg<4><<<ctas, 512>>>((const float4*)buf[i & 1], bytes / 16, out); }, 10));
        snprintf(nm, 64, "ldg U=8  %d CTAs x 512", ctas);
        report(nm, time_ms([&](int i) { k_ldg<8><<<ctas, 512>>>((const float4*)buf[i & 1], bytes / 16, out); }, 10));
    }
    {
        auto k = k_bulk<32768, 2>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
        report("bulk 32K x2 stages, 3 CTAs/SM", time_ms([&](int i) { k<<<148 * 3, 128, 65536>>>(buf[i & 1], bytes, out); }, 10));
    }
    {
        auto k = k_bulk<32768, 6>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768);
        report("bulk 32K x6 stages, 1 CTA/SM", time_ms([&](int i) { k<<<148, 128, 6 * 32768>>>(buf[i & 1], bytes, out); }, 10));
    }
    {
        auto k = k_bulk<65536, 3>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536);
        report("bulk 64K x3 stages, 1 CTA/SM", time_ms([&](int i) { k<<<148, 128, 3 * 65536>>>(buf[i & 1], bytes, out); }, 10));
    }
    {
        auto k = k_bulk<16384, 4>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384);
        report("bulk 16K x4 stages, 3 CTAs/SM", time_ms([&](int i) { k<<<148 * 3, 128, 4 * 16384>>>(buf[i & 1], bytes, out); }, 10));
    }
    {
        auto k = k_bulk<8192, 3>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 8192);
        report("bulk 8K x3 stages, 8 CTAs/SM", time_ms([&](int i) { k<<<148 * 8, 128, 3 * 8192>>>(buf[i & 1], bytes, out); }, 10));
    }
    for (int pf : {2, 4, 8}) {
        char nm[64];
        snprintf(nm, 64, "regpath nofence bulkpf PF=%d", pf);
        report(nm, time_ms([&](int i) { k_regpath<1, 0><<<148, 288>>>((const float*)buf[i & 1], bytes, pf, out); }, 10));
        snprintf(nm, 64, "regpath fence   bulkpf PF=%d", pf);
        report(nm, time_ms([&](int i) { k_regpath<1, 1><<<148, 288>>>((const float*)buf[i & 1], bytes, pf, out); }, 10));
        snprintf(nm, 64, "regpath fence   linepf PF=%d", pf);
        report(nm, time_ms([&](int i) { k_regpath<2, 1><<<148, 288>>>((const float*)buf[i & 1], bytes, pf, out); }, 10));
    }
    report("regpath nofence nopf", time_ms([&](int i) { k_regpath<0, 0><<<148, 288>>>((const float*)buf[i & 1], bytes, 0, out); }, 10));
    // the same with a large dynamic shared-memory allocation: the L1 carve-out shrinks to ~20 KB
    {
        auto k = k_regpath<0, 0>;
        for (int kb : {64, 128, 160, 200}) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
            char nm[64]; snprintf(nm, 64, "regpath nopf, %d KB dyn smem", kb);
            report(nm, time_ms([&](int i) { k<<<148, 288, kb * 1024>>>((const float*)buf[i & 1], bytes, 0, out); }, 10));
        }
    }
    {
        const int kb = 205;
        auto run = [&](auto k, const char* nm) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
            report(nm, time_ms([&](int i) { k<<<148, 288, kb * 1024>>>((const float*)buf[i & 1], bytes, 0, out); }, 10));
        };
        run(k_regpath<0, 0, 0>, "205KB smem: nc.no_allocate.evict_first");
        run(k_regpath<0, 0, 1>, "205KB smem: ld.cg");
        run(k_regpath<0, 0, 2>, "205KB smem: ld.cv");
        run(k_regpath<0, 0, 3>, "205KB smem: ld.ca");
        run(k_regpath<0, 0, 4>, "205KB smem: ld.nc");
        run(k_regpath<0, 0, 5>, "205KB smem: ld.cs");
        run(k_regpath<0, 0, 6>, "205KB smem: ld.relaxed.gpu");
    }
    report("regpath fence   nopf", time_ms([&](int i) { k_regpath<0, 1><<<148, 288>>>((const float*)buf[i & 1], bytes, 0, out); }, 10));
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
