// Dev harness (GPU box): tensor-pipe time of ONE 32-row tile of agg_tc_kernel's MMA stream —
//   GEMM1: 32 x (A in TMEM, M=128, N=64, B K-major smem)   GEMM2: 8 x (A MN-major smem, M=128, N=32) + 8 x (.., N=16)
// — issued (a) by one thread, kinds in sequence, (b) by two threads of two warps concurrently (as the kernel does),
// and (b) again while the other warps of the CTA (c) stream STS.64 into shared memory, (d) read TMEM with tcgen05.ld,
// (e) both.  Answers whether the 3 000 cycles per tile seen in situ are the pipe itself or contention.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/dev_mma_mix scripts/dev_mma_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../vlsa_b200/csrc/tc_common.cuh"
using namespace vlsa;

constexpr int NTILES = 64;
constexpr int TILE_BYTES = 65536, SLOT = 8192, PLANE = 4096;

__device__ __forceinline__ void issue_gemm1(uint32_t tm, uint64_t tb) {
    constexpr uint32_t idesc1 = umma_idesc(UMMA_F16, UMMA_F16, 128, 64, false, false);
#pragma unroll
    for (int s = 0; s < 8; ++s)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            tc_mma_ts(tm + 448, tm + (s * 4 + ks) * 8, umma_desc_advance(tb, s * SLOT + ks * 32), idesc1, (s | ks) != 0);
}
__device__ __forceinline__ void issue_gemm2(uint32_t tm, uint64_t ta, uint64_t wb) {
    constexpr uint32_t idesc_hi = umma_idesc(UMMA_F16, UMMA_F16, 128, 32, true, false);
    constexpr uint32_t idesc_lo = umma_idesc(UMMA_F16, UMMA_F16, 128, 16, true, false);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t d2 = tm + 256 + g * 48;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ah = umma_desc_advance(ta, (2 * g) * SLOT + ks * 2048);
            const uint64_t bd = umma_desc_advance(wb, ks * 32);
            tc_mma_ss(d2, ah, bd, idesc_hi, 1u);
            tc_mma_ss(d2 + 32, umma_desc_advance(ah, PLANE), bd, idesc_lo, 1u);
        }
    }
}

// mode bit 0: two issuers (else one), bit 1: STS stream from warps 4..11, bit 2: TMEM reads from warps 12..15 (+ 4..7)
__global__ void __launch_bounds__(640) mix_kernel(int mode, int pace_ns, long long* cyc, float* sink) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));       // 3 tile buffers (192 KB) | weights 8 KB | scratch 16 KB
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (3 * TILE_BYTES + 8192 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_fence_init(); stop = 0; }
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp < 4) {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = 0u;
        for (int c = 0; c < 512; c += 32) tmem_st32(tm + c + (uint32_t(32 * warp) << 16), v);
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint64_t ring_k = umma_desc_sw128(smem_u32(sm), 16, 1024);          // K-major view (GEMM1 B)
    const uint64_t ring_mn = umma_desc_sw128(smem_u32(sm), SLOT, 1024);       // MN-major view (GEMM2 A)
    const uint64_t wdesc = umma_desc_sw128(smem_u32(sm) + 3 * TILE_BYTES, 16, 1024);
    const bool two = mode & 1;
    if (warp == 0) {
        if (elect_one()) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int t = 0; t < NTILES; ++t) {
                const uint32_t b = t % 3;
                issue_gemm1(tm, umma_desc_advance(ring_k, b * TILE_BYTES));
                if (!two) issue_gemm2(tm, umma_desc_advance(ring_mn, b * TILE_BYTES), wdesc);
            }
            tc_commit(bar);
            mbar_wait_wd(bar, 0);
            cyc[0] = clock64() - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp == 1) {
        if (two && elect_one()) {
            const long long t0 = clock64();
#pragma unroll 1
            for (int t = 0; t < NTILES; ++t) {
                const uint32_t b = (t + 1) % 3;
                issue_gemm2(tm, umma_desc_advance(ring_mn, b * TILE_BYTES), wdesc);
            }
            tc_commit(bar + 1);
            mbar_wait_wd(bar + 1, 0);
            cyc[1] = clock64() - t0;
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12 && (mode & 2)) {
        // producers' store stream: STS.64 at swizzled offsets of a scratch tile (16 KB), until told to stop
        unsigned char* scr = sm + 3 * TILE_BYTES + 8192;
        uint32_t n = 0;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<uint2*>(scr + ((warp - 4) & 1) * 8192 + sw128_offset((i * 4 + (lane >> 3)) & 63, lane & 7, 0) + (i & 1) * 8) = make_uint2(n, lane);
            ++n;
            if (pace_ns) __nanosleep(pace_ns);
        }
        if (n == 0xffffffffu) sink[0] = 1.f;
    } else if (warp >= 12 && (mode & 4)) {
        // weight warps' TMEM reads (scores), until told to stop
        float acc = 0.f;
        const uint32_t tq = tm + (uint32_t(32 * (warp & 3)) << 16);
        while (!stop) {
            uint32_t a[16], b[16];
            tmem_ld16(tq + 448, a);
            tmem_ld16(tq + 480, b);
            tmem_wait_ld();
            acc += __uint_as_float(a[3]) + __uint_as_float(b[5]);
            __nanosleep(200);
        }
        if (acc == 123.f) sink[1] = acc;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* dC; float* dS;
    cudaMalloc(&dC, 16); cudaMalloc(&dS, 8);
    const int smem = 3 * TILE_BYTES + 8192 + 16384 + 1024;
    cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[8] = {"one issuer, GEMM1 then GEMM2 per tile", "two issuers", "one issuer + STS stream", "two issuers + STS stream",
                            "one issuer + TMEM reads", "two issuers + TMEM reads", "one issuer + STS + TMEM reads", "two issuers + STS + TMEM reads"};
    for (int run = 0; run < 12; ++run) {
        const int mode = run < 8 ? run : (run - 8) * 2 + (run >= 10 ? 1 - 2 * (run - 10) + 2 * (run - 10) : 0);
        const int pace = run < 8 ? 300 : 0;             // 300 ns between bursts of 8 STS.64 per warp ~ the kernel's 64 KB per tile period; 0 = saturating
        if (run >= 8 && !(run == 8 || run == 9 || run == 10 || run == 11)) continue;
        cudaMemset(dC, 0, 16);
        for (int rep = 0; rep < 2; ++rep) {
            mix_kernel<<<1, 640, smem>>>(run < 8 ? mode : (run == 8 ? 2 : run == 9 ? 3 : run == 10 ? 6 : 7), pace, dC, dS);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("[mma_mix] mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        long long C[2];
        cudaMemcpy(C, dC, 16, cudaMemcpyDeviceToHost);
        const int m_ = run < 8 ? mode : (run == 8 ? 2 : run == 9 ? 3 : run == 10 ? 6 : 7);
        printf("[mma_mix] %s %-34s: GEMM1 issuer %7.1f cycles / tile, GEMM2 issuer %7.1f   (standalone pipe model: 32 x 32.8 + 16 x 40.9 = 1704)\n",
               pace ? "paced STS    " : "saturated STS", names[m_], C[0] / double(NTILES), C[1] / double(NTILES));
    }
    return 0;
}
