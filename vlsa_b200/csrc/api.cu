// C-ABI entry points (include/vlsa_b200.h).  Host code only validates arguments, carves the
// caller's workspace and launches kernels on the caller's stream.
#include "../../include/vlsa_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>

#include "agg_simt.cuh"
#include "head_kernels.cuh"

using namespace vlsa;

#define VLSA_CUDA(expr)                                   \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return static_cast<int>(_e); \
    } while (0)

// switch over the compile-time prototype count
#define VLSA_DISPATCH_P(P_, ...)                                   \
    switch (P_) {                                                  \
        case 1: { constexpr int kP = 1; __VA_ARGS__; break; }      \
        case 2: { constexpr int kP = 2; __VA_ARGS__; break; }      \
        case 3: { constexpr int kP = 3; __VA_ARGS__; break; }      \
        case 4: { constexpr int kP = 4; __VA_ARGS__; break; }      \
        case 5: { constexpr int kP = 5; __VA_ARGS__; break; }      \
        case 6: { constexpr int kP = 6; __VA_ARGS__; break; }      \
        case 7: { constexpr int kP = 7; __VA_ARGS__; break; }      \
        case 8: { constexpr int kP = 8; __VA_ARGS__; break; }      \
        case 9: { constexpr int kP = 9; __VA_ARGS__; break; }      \
        case 10: { constexpr int kP = 10; __VA_ARGS__; break; }    \
        case 11: { constexpr int kP = 11; __VA_ARGS__; break; }    \
        case 12: { constexpr int kP = 12; __VA_ARGS__; break; }    \
        case 13: { constexpr int kP = 13; __VA_ARGS__; break; }    \
        case 14: { constexpr int kP = 14; __VA_ARGS__; break; }    \
        case 15: { constexpr int kP = 15; __VA_ARGS__; break; }    \
        case 16: { constexpr int kP = 16; __VA_ARGS__; break; }    \
        default: return VLSA_EINVAL;                               \
    }

static constexpr int kRowTile = 32;   // AggCfg::TN

static int device_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    return sms;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct AggWorkspace {
    float* part_m;
    float* part_l;
    float* part_O;
    float* delta;      // [B, P]   backward
    float* dgf;        // [2, B, D] backward: dg then df
    size_t bytes;
};

static AggWorkspace carve(void* base, int total_chunks, int B, int P) {
    AggWorkspace w;
    size_t off = 0;
    auto take = [&](size_t nfloat) {
        float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
        off += align_up(nfloat * sizeof(float), 256);
        return p;
    };
    w.part_m = take(size_t(total_chunks) * P);
    w.part_l = take(size_t(total_chunks) * P);
    w.part_O = take(size_t(total_chunks) * P * VLSA_D);
    w.delta = take(size_t(B) * P);
    w.dgf = take(size_t(2) * B * VLSA_D);
    w.bytes = off;
    return w;
}

template <int P, bool BWD, typename XT>
static int launch_agg(const AggParams& prm, cudaStream_t st) {
    using C = AggCfg<P, BWD, XT>;
    auto kern = agg_simt_kernel<P, BWD, XT>;
    VLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::SMEM)));
    const int sms = device_sm_count();
    const int grid = prm.total_chunks < sms ? prm.total_chunks : sms;
    if (grid <= 0) return 0;
    kern<<<grid, C::THREADS, C::SMEM, st>>>(prm);
    return static_cast<int>(cudaGetLastError());
}

extern "C" {

int vlsa_version(void) { return 100; }

const char* vlsa_error_string(int code) {
    if (code == 0) return "success";
    if (code == VLSA_EINVAL) return "vlsa: invalid argument";
    if (code == VLSA_EWORKSPACE) return "vlsa: workspace too small";
    if (code == VLSA_EUNSUPPORTED) return "vlsa: unsupported configuration";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "vlsa: unknown error";
}

int vlsa_agg_plan(const int64_t* cu_rows_host, int B, int sm_count, int* chunk_rows_out, int32_t* chunk_start_host) {
    if (!cu_rows_host || !chunk_rows_out || !chunk_start_host || B < 0) return VLSA_EINVAL;
    if (sm_count <= 0) sm_count = device_sm_count();
    long long total_tiles = 0;
    for (int b = 0; b < B; ++b) {
        const long long n = cu_rows_host[b + 1] - cu_rows_host[b];
        if (n < 0) return VLSA_EINVAL;
        total_tiles += (n + kRowTile - 1) / kRowTile;
    }
    // aim at ~4 chunks per persistent CTA (tail balance) while keeping a partial (P*2 KB) small next
    // to the chunk it summarises (>= 8 tiles = 512 KB of fp32 rows once the batch is large enough)
    long long tiles_per_chunk = (total_tiles + 4LL * sm_count - 1) / (4LL * sm_count);
    if (tiles_per_chunk < 1) tiles_per_chunk = 1;
    if (tiles_per_chunk > 64) tiles_per_chunk = 64;
    const int chunk_rows = int(tiles_per_chunk) * kRowTile;
    long long c = 0;
    for (int b = 0; b < B; ++b) {
        chunk_start_host[b] = int32_t(c);
        const long long n = cu_rows_host[b + 1] - cu_rows_host[b];
        c += (n + chunk_rows - 1) / chunk_rows;
        if (c > 0x7fffffffLL) return VLSA_EINVAL;
    }
    chunk_start_host[B] = int32_t(c);
    *chunk_rows_out = chunk_rows;
    return 0;
}

size_t vlsa_agg_workspace_bytes(int total_chunks, int B, int P) {
    if (total_chunks < 0 || B < 0 || P < 1 || P > VLSA_MAX_P) return 0;
    return carve(nullptr, total_chunks, B, P).bytes + 256;
}

int vlsa_agg_fwd(const void* X, int x_dtype, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                 int chunk_rows, int total_chunks, const float* Q, int P, float coattn_scale, const float* W,
                 const float* bias, const float* T, int R, const float* logit_scale, void* workspace,
                 size_t workspace_bytes, float* out_v, float* out_f, float* out_g, float* out_logits,
                 float* out_if, float* out_ml, float* out_O, float* out_Tn, void* stream) {
    if (B == 0) return 0;
    if (!cu_rows || !chunk_start || !Q || !W || !bias || !T || !logit_scale || !out_v || !out_f || !out_g ||
        !out_logits || !out_ml)
        return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || R < 1 || R > VLSA_MAX_R || chunk_rows <= 0 || chunk_rows % kRowTile ||
        total_chunks < 0)
        return VLSA_EINVAL;
    if (x_dtype != VLSA_DTYPE_F32 && x_dtype != VLSA_DTYPE_BF16) return VLSA_EUNSUPPORTED;
    if (total_chunks > 0 && (!X || !workspace)) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if (total_chunks > 0 && (base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes)
        return VLSA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.scale = coattn_scale;
    prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;

    int rc = 0;
    VLSA_DISPATCH_P(P, {
        if (x_dtype == VLSA_DTYPE_F32) rc = launch_agg<kP, false, float>(prm, st);
        else rc = launch_agg<kP, false, __nv_bfloat16>(prm, st);
        if (rc) return rc;
        merge_fwd_kernel<kP><<<dim3(B, VLSA_D / 128), 128, 0, st>>>(ws.part_m, ws.part_l, ws.part_O, chunk_start,
                                                                     out_ml, out_O, out_v);
    });
    VLSA_CUDA(cudaGetLastError());
    adapter_fwd_kernel<<<VLSA_D / 4, 128, 0, st>>>(W, bias, out_v, B, out_f);
    VLSA_CUDA(cudaGetLastError());
    head_fwd_kernel<<<B, 256, 0, st>>>(out_f, T, R, logit_scale, out_g, out_logits, out_if, out_Tn);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
