// Dev harness (GPU box): cycles per tcgen05.mma (kind::f16, K = 16) as a function of shape and operand source.
// One CTA, one issuing thread, NREP back-to-back accumulating MMAs, clock64 around issue .. commit-barrier.
// Operand contents are zeros (timing only).   usage: dev_mma_sweep [ws]
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/dev_mma_sweep scripts/dev_mma_sweep.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../vlsa_b200/csrc/tc_common.cuh"
using namespace vlsa;

struct Case {
    int a_src;        // 0 = TMEM, 1 = smem K-major, 2 = smem MN-major
    int b_mn;         // 0 = K-major B, 1 = MN-major B
    int M, N;
    int nacc;         // accumulators cycled through (1 = same D every time)
    int ws;           // 1 = tcgen05.mma.ws
    int a_step;       // advance of the A operand between MMAs (bytes for smem, columns for TMEM), cycled over 4
    int b_step;       // same for B (bytes)
};

__device__ __forceinline__ void tc_mma_ws_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ws_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.ws.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

constexpr int NREP = 256;

__global__ void __launch_bounds__(128) sweep_kernel(const Case* cases, int ncases, long long* cyc) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* sm = raw + (base - smem_u32(raw));            // 64 KB A region | 64 KB B region, zeros
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_base;
    {   // zero all of TMEM (A operands and accumulators)
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = 0u;
        for (int c = 0; c < 512; c += 32) tmem_st32(tm + c + (uint32_t(32 * warp) << 16), v);
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t phase = 0;
    for (int ci = 0; ci < ncases; ++ci) {
        const Case c = cases[ci];
        if (warp == 0 && elect_one()) {
            const uint32_t idesc = umma_idesc(UMMA_F16, UMMA_F16, c.M, c.N, c.a_src == 2, c.b_mn != 0);
            const uint64_t adesc = c.a_src == 2 ? umma_desc_sw128(smem_u32(sm), 8192, 1024) : umma_desc_sw128(smem_u32(sm), 16, 1024);
            const uint64_t bdesc = c.b_mn ? umma_desc_sw128(smem_u32(sm) + 65536, 8192, 1024)
                                          : umma_desc_sw128(smem_u32(sm) + 65536, 16, 1024);
            // accumulators from column 128 on, N columns each; A (TMEM) in columns 0..31.  Everything the loop needs is
            // precomputed: the timed loop is one MMA per iteration with register operands (a single thread issues
            // dependent integer code at ~6 cycles per instruction, which would otherwise dominate)
            uint32_t dd[4], ta[4]; uint64_t ad[4], bd[4];
            for (int j = 0; j < 4; ++j) {
                dd[j] = tm + 128 + c.N * (j % c.nacc);
                ta[j] = tm + uint32_t(c.a_step) * j;
                ad[j] = umma_desc_advance(adesc, uint32_t(c.a_step) * j);
                bd[j] = umma_desc_advance(bdesc, uint32_t(c.b_step) * j);
            }
            const int kind = (c.a_src == 0 ? 0 : 1) + 2 * c.ws;
            const long long t0 = clock64();
            if (kind == 0) {
#pragma unroll 1
                for (int i = 0; i < NREP; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tc_mma_ts(dd[j], ta[j], bd[j], idesc, 1);
                }
            } else if (kind == 1) {
#pragma unroll 1
                for (int i = 0; i < NREP; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tc_mma_ss(dd[j], ad[j], bd[j], idesc, 1);
                }
            } else if (kind == 2) {
#pragma unroll 1
                for (int i = 0; i < NREP; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tc_mma_ws_ts(dd[j], ta[j], bd[j], idesc, 1);
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < NREP; i += 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tc_mma_ws_ss(dd[j], ad[j], bd[j], idesc, 1);
                }
            }
            tc_commit(&bar);
            mbar_wait_wd(&bar, phase);
            cyc[ci] = clock64() - t0;
        }
        phase ^= 1;
        __syncthreads();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
    const bool ws = argc > 1 && !strcmp(argv[1], "ws");
    std::vector<Case> cs;
    std::vector<const char*> notes;
    auto add = [&](int a_src, int b_mn, int M, int N, int nacc, int w, int as, int bs) { cs.push_back({a_src, b_mn, M, N, nacc, w, as, bs}); };
    if (!ws) {
        const int Ns[] = {8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256};
        for (int M : {128, 64}) {
            auto legal = [&](int N) { return M == 128 ? N % 16 == 0 : N % 8 == 0; };
            for (int N : Ns) if (legal(N)) add(0, 0, M, N, 1, 0, 8, 32);           // A TMEM, B K-major (GEMM1 today: M128 N64)
            for (int N : Ns) if (legal(N)) add(1, 0, M, N, 1, 0, 32, 32);          // A smem K-major, B K-major (swapped GEMM1)
            for (int N : Ns) if (legal(N)) add(2, 0, M, N, 1, 0, 2048, 32);        // A smem MN-major, B K-major (GEMM2 today: N32 / N16)
            for (int N : {64, 128, 256}) add(1, 1, M, N, 1, 0, 32, 2048);   // A K-major, B MN-major (GEMM2 with d on N)
            for (int N : {64, 128, 256}) add(0, 1, M, N, 1, 0, 8, 2048);    // A TMEM, B MN-major
        }
        // several accumulators (does switching D cost?)
        for (int nacc : {2, 4}) { add(0, 0, 128, 64, nacc, 0, 8, 32); add(2, 0, 128, 32, nacc, 0, 2048, 32); add(1, 0, 128, 32, nacc, 0, 32, 32); }
        // same operands every time (no descriptor advance): is the cost the operand fetch?
        add(0, 0, 128, 64, 1, 0, 0, 0); add(2, 0, 128, 32, 1, 0, 0, 0); add(1, 0, 128, 32, 1, 0, 0, 0); add(1, 0, 64, 32, 1, 0, 0, 0);
    } else {
        for (int M : {128, 64, 32})
            for (int N : {64, 128, 256}) {
                add(1, 0, M, N, 1, 1, 32, 32);
                add(0, 0, M, N, 1, 1, 8, 32);
                add(1, 1, M, N, 1, 1, 32, 2048);
            }
    }
    Case* dC; long long* dT;
    cudaMalloc(&dC, cs.size() * sizeof(Case)); cudaMalloc(&dT, cs.size() * 8);
    cudaMemcpy(dC, cs.data(), cs.size() * sizeof(Case), cudaMemcpyHostToDevice);
    cudaMemset(dT, 0, cs.size() * 8);
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    for (int rep = 0; rep < 2; ++rep) {
        sweep_kernel<<<1, 128, 131072 + 1024>>>(dC, int(cs.size()), dT);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[mma_sweep%s] kernel: %s\n", ws ? " ws" : "", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> T(cs.size());
    cudaMemcpy(T.data(), dT, cs.size() * 8, cudaMemcpyDeviceToHost);
    const char* an[3] = {"A tmem", "A smem K-major", "A smem MN-major"};
    printf("[mma_sweep%s] cycles per MMA (%d back-to-back, K=16, fp16)\n", ws ? " ws" : "", NREP);
    for (size_t i = 0; i < cs.size(); ++i)
        printf("  %s%-16s B %-8s M=%3d N=%3d nacc=%d astep=%4d bstep=%4d : %7.1f\n", cs[i].ws ? "ws " : "", an[cs[i].a_src],
               cs[i].b_mn ? "MN-major" : "K-major", cs[i].M, cs[i].N, cs[i].nacc, cs[i].a_step, cs[i].b_step, T[i] / double(NREP));
    return 0;
}
