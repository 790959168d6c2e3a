// Dev harness (GPU box): validates the UMMA descriptor / TMEM layout assumptions of the tensor-core kernel.
//  test 1: D[64 x N]  = A[64 x 64] (K-major, SW128) * B[N x 64]^T (K-major, SW128)           (scores GEMM form)
//  test 2: D[128 x N] = A^T, A = two [64(K) x 64(M)] planes (MN-major, SW128, LBO = plane stride) * B[N x 64]^T
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../vlsa_b200/csrc/tc_common.cuh"
using namespace vlsa;

constexpr int N = 32;

__global__ void __launch_bounds__(128) umma_test(const __nv_bfloat16* A1, const __nv_bfloat16* B1, float* D1,
                                                 const __nv_bfloat16* X2, const __nv_bfloat16* B2, float* D2) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));
    unsigned char* sA = sm;                 // 8 KB  (test 1 A) / 2 x 8 KB planes (test 2)
    unsigned char* sB = sm + 16384;         // 4 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;

    // ---------------- test 1
    for (int i = tid; i < 64 * 64; i += 128) {        // A1[row][k]
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128_offset(r, k >> 3, (k & 7) * 2)) = A1[i];
    }
    for (int i = tid; i < N * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k >> 3, (k & 7) * 2)) = B1[i];
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_bf16(64, N, false, false);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = umma_desc_sw128(smem_u32(sA) + ks * 32, 16, 1024);
            const uint64_t bd = umma_desc_sw128(smem_u32(sB) + ks * 32, 16, 1024);
            tc_mma_bf16(tm, ad, bd, idesc, ks > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait_wd(&bar, 0);
    tc_fence_after();
    {
        uint32_t r0[16], r1[16];
        const uint32_t ta = tm + (uint32_t(warp * 32) << 16);
        tmem_ld16(ta, r0);
        tmem_ld16(ta + 16, r1);
        tmem_wait_ld();
        // dump every lane: host decides which lanes hold which rows
        for (int c = 0; c < 16; ++c) { D1[(warp * 32 + lane) * N + c] = __uint_as_float(r0[c]); D1[(warp * 32 + lane) * N + 16 + c] = __uint_as_float(r1[c]); }
    }
    tc_fence_before();
    __syncthreads();

    // ---------------- test 2: X2[plane p][k row][m] : m = p*64 + mm
    for (int i = tid; i < 2 * 64 * 64; i += 128) {
        const int p = i / 4096, k = (i / 64) % 64, mm = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sA + p * 8192 + sw128_offset(k, mm >> 3, (mm & 7) * 2)) = X2[i];
    }
    for (int i = tid; i < N * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(r, k >> 3, (k & 7) * 2)) = B2[i];
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, true, false);
        for (int ks = 0; ks < 4; ++ks) {          // 16 K rows per step = 2 atoms of 8 rows
            const uint64_t ad = umma_desc_sw128(smem_u32(sA) + ks * 2048, 8192, 1024);
            const uint64_t bd = umma_desc_sw128(smem_u32(sB) + ks * 32, 16, 1024);
            tc_mma_bf16(tm + 32, ad, bd, idesc, ks > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait_wd(&bar, 1);
    tc_fence_after();
    {
        uint32_t r0[16], r1[16];
        const uint32_t ta = tm + 32 + (uint32_t(warp * 32) << 16);
        tmem_ld16(ta, r0);
        tmem_ld16(ta + 16, r1);
        tmem_wait_ld();
        for (int c = 0; c < 16; ++c) { D2[(warp * 32 + lane) * N + c] = __uint_as_float(r0[c]); D2[(warp * 32 + lane) * N + 16 + c] = __uint_as_float(r1[c]); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

static float bf(float x) { __nv_bfloat16 b = __float2bfloat16(x); return __bfloat162float(b); }

int main() {
    std::vector<__nv_bfloat16> A1(64 * 64), B1(N * 64), X2(2 * 64 * 64), B2(N * 64);
    std::vector<float> fA1(64 * 64), fB1(N * 64), fX2(2 * 64 * 64), fB2(N * 64);
    srand(1);
    auto rnd = [] { return float(rand() % 2001 - 1000) / 500.f; };
    for (size_t i = 0; i < A1.size(); ++i) { fA1[i] = bf(rnd()); A1[i] = __float2bfloat16(fA1[i]); }
    for (size_t i = 0; i < B1.size(); ++i) { fB1[i] = bf(rnd()); B1[i] = __float2bfloat16(fB1[i]); }
    for (size_t i = 0; i < X2.size(); ++i) { fX2[i] = bf(rnd()); X2[i] = __float2bfloat16(fX2[i]); }
    for (size_t i = 0; i < B2.size(); ++i) { fB2[i] = bf(rnd()); B2[i] = __float2bfloat16(fB2[i]); }
    __nv_bfloat16 *dA1, *dB1, *dX2, *dB2; float *dD1, *dD2;
    cudaMalloc(&dA1, A1.size() * 2); cudaMalloc(&dB1, B1.size() * 2); cudaMalloc(&dX2, X2.size() * 2); cudaMalloc(&dB2, B2.size() * 2);
    cudaMalloc(&dD1, 128 * N * 4); cudaMalloc(&dD2, 128 * N * 4);
    cudaMemcpy(dA1, A1.data(), A1.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB1, B1.data(), B1.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dX2, X2.data(), X2.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB2, B2.data(), B2.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD1, 0, 128 * N * 4); cudaMemset(dD2, 0, 128 * N * 4);
    const int smem = 16384 + 4096 + 1024;
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_test<<<1, 128, smem>>>(dA1, dB1, dD1, dX2, dB2, dD2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> D1(128 * N), D2(128 * N);
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    // test 1: expected row m at lane (m%16) + 32*(m/16)
    double err1 = 0, err1_alt = 0;
    for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < 64; ++k) ref += double(fA1[m * 64 + k]) * fB1[n * 64 + k];
        const int lane = (m % 16) + 32 * (m / 16);
        err1 = fmax(err1, fabs(D1[lane * N + n] - ref));
        err1_alt = fmax(err1_alt, fabs(D1[m * N + n] - ref));     // alternative: row m at lane m
    }
    printf("test1 (M=64 K-major): max err with lane=(m%%16)+32*(m/16): %.3e ; with lane=m: %.3e\n", err1, err1_alt);
    double err2 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < 64; ++k) ref += double(fX2[(m / 64) * 4096 + k * 64 + (m % 64)]) * fB2[n * 64 + k];
        err2 = fmax(err2, fabs(D2[m * N + n] - ref));
    }
    printf("test2 (M=128 MN-major A, LBO=plane stride): max err %.3e\n", err2);
    printf("sample D1 lane0: %f %f ; D2 lane0: %f %f\n", D1[0], D1[1], D2[0], D2[1]);
    return (err1 < 1e-2 && err2 < 1e-2) ? 0 : 2;
}
