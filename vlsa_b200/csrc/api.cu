// C-ABI entry points (include/vlsa_b200.h).  Host code only validates arguments, carves the
// caller's workspace and launches kernels on the caller's stream.
#include "../../include/vlsa_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "agg_simt.cuh"
#include "agg_tc.cuh"
#include "agg_bf16.cuh"
#include "agg_split.cuh"
#include "head_kernels.cuh"
#include "loss_kernels.cuh"
#include "aux_kernels.cuh"
#include "optim_kernels.cuh"

using namespace vlsa;

#define VLSA_CUDA(expr)                                   \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return static_cast<int>(_e); \
    } while (0)

// switch over the compile-time prototype count (VLSA_DEV_FEWP: development builds with P in {4, 12} only)
#ifdef VLSA_DEV_FEWP
#define VLSA_DISPATCH_P(P_, ...)                                   \
    switch (P_) {                                                  \
        case 4: { constexpr int kP = 4; __VA_ARGS__; break; }      \
        case 12: { constexpr int kP = 12; __VA_ARGS__; break; }    \
        default: return VLSA_EINVAL;                               \
    }
#else
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
#endif

#ifndef VLSA_BF16_TC_MIN_P
#define VLSA_BF16_TC_MIN_P 1     // bf16 rows: tensor-core kernel from this P on (measured: 315 vs 472 us at P = 4, 332 vs 1226 us at P = 12)
#endif
static constexpr int kRowTile = 32;   // a multiple of the tcgen05 kernels' tile (16 rows) and of the CUDA-core kernel's AggCfg::TN
static_assert(kRowTile % (4 * VLSA_AGG_WARPS) == 0 && kRowTile % TcCfg::TR == 0 && kRowTile % Bf16Cfg::TR == 0,
              "chunk_rows must suit every streaming kernel");

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
    float* df;         // [B, D]   backward
    float* dv;         // [B, D]   backward
    float* dls_part;   // [B]      backward
    float* l2_m;       // level-2 merge partials: [B S, P] | backward: unused
    float* l2_l;       // [B S, P]
    float* l2_O;       // [B S, P, D]  | backward: [S, P, D]
    size_t bytes;
};

// Two-level merge of the per-chunk partials.  Level 1 folds runs of chunks with many small CTAs (about 8 per
// SM), level 2 finishes per bag (forward) / per prototype (backward).  Splits depend only on the arguments below,
// so the workspace size and the summation order are fixed by the plan.
static int merge_fwd_splits(int total_chunks, int B, int P) {
    if (B <= 0 || total_chunks < 8 * B) return 0;          // short bags: one level is enough
    int S = 1184 / (B * P);
    const int avg = total_chunks / B;
    if (S > avg / 4) S = avg / 4;
    if (S > 64) S = 64;
    return S < 1 ? 1 : S;
}
// Level 1 (S x P CTAs) folds total_chunks / S partials per thread two at a time, level 2 (P CTAs) folds S entries four at a
// time: both are chains of dependent-latency rounds, shortest in sum at S ~ sqrt(2 total_chunks) (the first version maximised
// the CTAs of level 1 — S = 1184 / P — and left level 2 walking 296 entries on 4 CTAs at P = 4: 25 us of the step).
static int merge_bwd_splits(int total_chunks, int P) {
    if (total_chunks < 16) return 0;
    int S = int(sqrtf(2.f * float(total_chunks)) + 0.5f);
    if (S > 1184 / P) S = 1184 / P;
    if (S > total_chunks / 4) S = total_chunks / 4;
    return S < 1 ? 1 : S;
}

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
    w.df = take(size_t(B) * VLSA_D);
    w.dv = take(size_t(B) * VLSA_D);
    w.dls_part = take(size_t(B));
    const int sf = merge_fwd_splits(total_chunks, B, P), sb = merge_bwd_splits(total_chunks, P);
    const size_t e = size_t(sf) * B > size_t(sb) ? size_t(sf) * B : size_t(sb);
    w.l2_m = take(size_t(sf) * B * P);
    w.l2_l = take(size_t(sf) * B * P);
    w.l2_O = take(e * P * VLSA_D);
    w.bytes = off;
    return w;
}

// Per-(kernel, device) launch attributes, set / queried once: cudaFuncSetAttribute and the occupancy query cost more
// host time than the launch itself on the one-bag-per-call path.  Once-initialised caches of device facts, not state:
// a value is written once and is the same for every caller.
static constexpr int kMaxDevices = 64;
template <typename K>
static int kernel_slots(K kern, int threads, int smem, std::atomic<int>* cache) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    int v = cache[dev].load(std::memory_order_acquire);
    if (v > 0) return v;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) return -1;
    if (occ < 1) occ = 1;
    v = device_sm_count() * occ;
    cache[dev].store(v, std::memory_order_release);
    return v;
}

template <int P, int MODE, typename XT>
static int launch_agg(const AggParams& prm, cudaStream_t st) {
    using C = AggCfg<P, MODE, XT>;
    auto kern = agg_simt_kernel<P, MODE, XT>;
    static std::atomic<int> cache[kMaxDevices];
    const int slots = kernel_slots(kern, C::THREADS, int(C::SMEM), cache);   // persistent CTAs: every resident slot, no second wave
    if (slots <= 0) return static_cast<int>(cudaGetLastError());
    const int grid = prm.total_chunks < slots ? prm.total_chunks : slots;
    if (grid <= 0) return 0;
    kern<<<grid, C::THREADS, C::SMEM, st>>>(prm);
    return static_cast<int>(cudaGetLastError());
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// bf16-stored rows on the tensor cores (agg_bf16.cuh): X [total_rows, 512] bf16 as a 2-D tensor, box = 16 rows x 64 columns
// landing as two 128-byte-swizzled K-major atoms
template <bool BWD>
static int launch_agg_bf16(const AggParams& prm, int P, long long total_rows, cudaStream_t st) {
    using C = Bf16Cfg;
    if (total_rows <= 0 || total_rows > 0x7fffffffLL) return VLSA_EINVAL;
    if (reinterpret_cast<uintptr_t>(prm.X) & 15u) return VLSA_EINVAL;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return VLSA_EUNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {cuuint64_t(VLSA_D), cuuint64_t(total_rows)};
    const cuuint64_t gstr[1] = {cuuint64_t(VLSA_D) * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {64u, cuuint32_t(C::TR)};
    const cuuint32_t estr[2] = {1u, 1u};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(prm.X), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return VLSA_EINVAL;
    auto kern = agg_bf16_kernel<BWD>;
    static std::atomic<int> cache[kMaxDevices];
    if (kernel_slots(kern, C::THREADS, int(C::SMEM), cache) <= 0) return static_cast<int>(cudaGetLastError());
    const int sms = device_sm_count();
    const int grid = prm.total_chunks < sms ? prm.total_chunks : sms;
    if (grid <= 0) return 0;
    kern<<<grid, C::THREADS, C::SMEM, st>>>(prm, P, tmap);
    return static_cast<int>(cudaGetLastError());
}

// pre-split tile images of a device cohort (agg_split.cuh): one bulk copy per 16-row record
template <bool BWD>
static int launch_agg_split(const AggParams& prm, int P, cudaStream_t st) {
    using C = SplitCfg;
    if (reinterpret_cast<uintptr_t>(prm.X) & 15u) return VLSA_EINVAL;
    if (!prm.row_ranges) return VLSA_EINVAL;                 // bags start at record boundaries: ranges in the padded row space
    auto kern = agg_split_kernel<BWD>;
    static std::atomic<int> cache[kMaxDevices];
    if (kernel_slots(kern, C::THREADS, int(C::SMEM), cache) <= 0) return static_cast<int>(cudaGetLastError());
    const int sms = device_sm_count();
    const int grid = prm.total_chunks < sms ? prm.total_chunks : sms;
    if (grid <= 0) return 0;
    kern<<<grid, C::THREADS, C::SMEM, st>>>(prm, P);
    return static_cast<int>(cudaGetLastError());
}

// register-staged tcgen05 kernel (agg_tc.cuh): rows through LDG, three shared-memory passes per byte
template <bool BWD>
static int launch_agg_tc(const AggParams& prm, int P, cudaStream_t st) {
    using C = TcCfg;
    if (reinterpret_cast<uintptr_t>(prm.X) & 15u) return VLSA_EINVAL;
    auto kern = agg_tc_kernel<BWD>;
    static std::atomic<int> cache[kMaxDevices];
    if (kernel_slots(kern, C::THREADS, int(C::SMEM), cache) <= 0) return static_cast<int>(cudaGetLastError());
    const int sms = device_sm_count();
    const int grid = prm.total_chunks < sms ? prm.total_chunks : sms;
    if (grid <= 0) return 0;
    kern<<<grid, C::THREADS, C::SMEM, st>>>(prm, P);
    return static_cast<int>(cudaGetLastError());
}

// Which streaming kernel serves a pass.  Default: fp32 rows — the register-staged tcgen05 kernel for P > 5 (the CUDA-core
// kernel is at the HBM roofline for P <= 5); bf16 rows — the TMA-fed tcgen05 kernel for every P.  The caller can force a
// kernel per call with the VLSA_KERNEL_* bits of x_dtype (cross-checks in the parity tests): no process-wide switch.
enum AggKernel { kAggSimt = 0, kAggTc = 2, kAggBf16 = 3, kAggSplit = 4 };
static AggKernel agg_kernel_choice(int P, int x_dtype_flags) {
    const int dtype = x_dtype_flags & VLSA_DTYPE_MASK;
    if (dtype == VLSA_DTYPE_SPLIT16) return kAggSplit;
    if (x_dtype_flags & VLSA_KERNEL_SIMT) return kAggSimt;
    if (dtype == VLSA_DTYPE_BF16) return ((x_dtype_flags & VLSA_KERNEL_TC) || P > VLSA_BF16_TC_MIN_P - 1) ? kAggBf16 : kAggSimt;
    if (x_dtype_flags & VLSA_KERNEL_TC) return kAggTc;
    return P > 5 ? kAggTc : kAggSimt;
}
static bool dtype_ok(int x_dtype_flags) {
    const int dtype = x_dtype_flags & VLSA_DTYPE_MASK;
    if (dtype == VLSA_DTYPE_SPLIT16) return (x_dtype_flags & ~(VLSA_DTYPE_MASK | VLSA_KERNEL_TC)) == VLSA_ROWS_RANGES;
    return (dtype == VLSA_DTYPE_F32 || dtype == VLSA_DTYPE_BF16) &&
           (x_dtype_flags & ~(VLSA_DTYPE_MASK | VLSA_KERNEL_SIMT | VLSA_KERNEL_TC | VLSA_ROWS_RANGES)) == 0;
}

// prototypes per launch of the per-prototype-gradient backward (measured best, profiles/variant_time_r01.json)
static constexpr int kGenGroup = 8;

static int launch_agg_fwd(const AggParams& prm, int P, int x_dtype, long long total_rows, cudaStream_t st) {
    const AggKernel k = agg_kernel_choice(P, x_dtype);
    if (k == kAggTc) return launch_agg_tc<false>(prm, P, st);
    if (k == kAggBf16) return launch_agg_bf16<false>(prm, P, total_rows, st);
    if (k == kAggSplit) return launch_agg_split<false>(prm, P, st);
    int rc = 0;
    VLSA_DISPATCH_P(P, {
        if ((x_dtype & VLSA_DTYPE_MASK) == VLSA_DTYPE_F32) rc = launch_agg<kP, 0, float>(prm, st);
        else rc = launch_agg<kP, 0, __nv_bfloat16>(prm, st);
    });
    return rc;
}

extern "C" {

int vlsa_version(void) { return 200; }

#ifdef VLSA_TMA_PROF
int vlsa_debug_read_prof(long long* out_host) {          // development builds only
    return static_cast<int>(cudaMemcpyFromSymbol(out_host, vlsa::g_tma_prof, sizeof(long long) * 32));
}
#endif

size_t vlsa_adam_segment_bytes(void) { return sizeof(AdamSeg); }

int vlsa_adam_step(const void* segments, int S, int64_t max_n, const float* grads_flat, float* exp_avg, float* exp_avg_sq,
                   const float* step_count_in, float* step_count_out, const float* flags, float beta1, float beta2, float eps,
                   void* stream) {
    if (S < 0 || max_n < 0 || !segments || !grads_flat || !exp_avg || !exp_avg_sq || !step_count_in || !step_count_out ||
        step_count_in == step_count_out)
        return VLSA_EINVAL;
    if (S == 0 || max_n == 0) return 0;
    if (S > 65535) return VLSA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(unsigned((max_n + 1023) / 1024), unsigned(S));
    adam_step_kernel<<<grid, 256, 0, st>>>(static_cast<const AdamSeg*>(segments), grads_flat, exp_avg, exp_avg_sq, step_count_in,
                                           step_count_out, flags, beta1, beta2, eps);
    return static_cast<int>(cudaGetLastError());
}

#ifdef VLSA_WD_DEBUG
int vlsa_debug_read_wd(unsigned int* out_host, int reset) {          // development builds only
    int rc = static_cast<int>(cudaMemcpyFromSymbol(out_host, vlsa::g_wd, sizeof(unsigned int) * (4 + 4 * 1000)));
    if (rc == 0 && reset) { static unsigned int zero[4 + 4 * 1000]; rc = static_cast<int>(cudaMemcpyToSymbol(vlsa::g_wd, zero, sizeof(zero))); }
    return rc;
}
#endif

const char* vlsa_error_string(int code) {
    if (code == 0) return "success";
    if (code == VLSA_EINVAL) return "vlsa: invalid argument";
    if (code == VLSA_EWORKSPACE) return "vlsa: workspace too small";
    if (code == VLSA_EUNSUPPORTED) return "vlsa: unsupported configuration";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "vlsa: unknown error";
}

size_t vlsa_split16_row_bytes(void) { return size_t(SplitCfg::ROW_BYTES); }

int vlsa_split16_pack(const float* X, int64_t n_rows, void* image, int64_t first_row, void* stream) {
    if (n_rows == 0) return 0;
    if (!X || !image || n_rows < 0 || first_row < 0 || first_row % SplitCfg::TR) return VLSA_EINVAL;
    if ((reinterpret_cast<uintptr_t>(X) & 15u) || (reinterpret_cast<uintptr_t>(image) & 15u)) return VLSA_EINVAL;
    const long long records = (n_rows + SplitCfg::TR - 1) / SplitCfg::TR;
    if (records > 0x7fffffffLL) return VLSA_EINVAL;
    split16_pack_kernel<<<unsigned(records), 512, 0, static_cast<cudaStream_t>(stream)>>>(X, n_rows, static_cast<unsigned char*>(image), first_row);
    return static_cast<int>(cudaGetLastError());
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
    // ~12 chunks per SM (several persistent CTAs share an SM; a few chunks each for tail balance), but keep
    // a chunk >= 8 row tiles so that its partial (P x 2 KB) stays small next to the rows it summarises;
    // tiny problems fall back to one chunk per SM.
    long long tiles_per_chunk = (total_tiles + 12LL * sm_count - 1) / (12LL * sm_count);
    if (tiles_per_chunk < 8) {
        const long long per_sm = total_tiles / sm_count;
        tiles_per_chunk = per_sm < 8 ? (per_sm < 1 ? 1 : per_sm) : 8;
    }
    if (tiles_per_chunk > 128) tiles_per_chunk = 128;
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

int vlsa_agg_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                 int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale, const float* W,
                 const float* bias, const float* T, int R, const float* logit_scale, void* workspace,
                 size_t workspace_bytes, float* out_v, float* out_f, float* out_g, float* out_logits,
                 float* out_if, float* out_ml, float* out_O, float* out_Tn, void* stream) {
    if (B == 0) return 0;
    if (!cu_rows || !chunk_start || !Q || !W || !bias || !out_v || !out_f || !out_ml) return VLSA_EINVAL;
    if (T && (!logit_scale || !out_g || !out_logits || R < 1 || R > VLSA_MAX_R)) return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || chunk_rows <= 0 || chunk_rows % kRowTile || total_chunks < 0)
        return VLSA_EINVAL;
    if (!dtype_ok(x_dtype)) return VLSA_EUNSUPPORTED;
    if (q_prenorm != 0 && q_prenorm != 1) return VLSA_EINVAL;
    if (total_chunks > 0 && (!X || !workspace || total_rows <= 0)) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if (total_chunks > 0 && (base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes)
        return VLSA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.row_ranges = (x_dtype & VLSA_ROWS_RANGES) ? 1 : 0;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.scale = coattn_scale;
    prm.q_prenorm = q_prenorm;
    prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;

    int rc = launch_agg_fwd(prm, P, x_dtype, total_rows, st);
    if (rc) return rc;
    const int S = merge_fwd_splits(total_chunks, B, P);
    if (S > 0) {
        merge_fwd_split_kernel<<<dim3(S, P, B), 128, 0, st>>>(ws.part_m, ws.part_l, ws.part_O, chunk_start, P, S,
                                                                ws.l2_m, ws.l2_l, ws.l2_O);
        VLSA_CUDA(cudaGetLastError());
    }
    merge_fwd_kernel<<<dim3(B, VLSA_D / 64), dim3(64, P), 0, st>>>(S > 0 ? ws.l2_m : ws.part_m, S > 0 ? ws.l2_l : ws.part_l,
                                                                   S > 0 ? ws.l2_O : ws.part_O, chunk_start, P, S,
                                                                   out_ml, out_O, out_v);
    VLSA_CUDA(cudaGetLastError());
    adapter_fwd_kernel<<<VLSA_D / 4, 128, 0, st>>>(W, bias, out_v, B, out_f);
    VLSA_CUDA(cudaGetLastError());
    if (T) {
        head_fwd_kernel<<<B, 256, 0, st>>>(out_f, T, R, logit_scale, out_g, out_logits, out_if, out_Tn);
        VLSA_CUDA(cudaGetLastError());
    }
    return 0;
}

int vlsa_agg_partial_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                         int chunk_rows, int total_chunks, const float* Q, int P, float coattn_scale,
                         void* workspace, size_t workspace_bytes, void* stream) {
    if (B == 0 || total_chunks == 0) return 0;
    if (!X || !cu_rows || !chunk_start || !Q || !workspace) return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || chunk_rows <= 0 || chunk_rows % kRowTile || total_chunks < 0)
        return VLSA_EINVAL;
    if (!dtype_ok(x_dtype) || total_rows <= 0) return VLSA_EUNSUPPORTED;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes) return VLSA_EWORKSPACE;
    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.row_ranges = (x_dtype & VLSA_ROWS_RANGES) ? 1 : 0;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.scale = coattn_scale;
    prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return launch_agg_fwd(prm, P, x_dtype, total_rows, st);
}

int vlsa_agg_bwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                 int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale, const float* W,
                 const float* T, int R, const float* logit_scale, const float* v, const float* f, const float* g,
                 const float* logits, const float* ml, const float* O, const float* d_logits, const float* d_g,
                 const float* d_f, void* workspace, size_t workspace_bytes, float* dQ, float* dW, float* db,
                 float* dT, float* dlogit_scale, void* stream) {
    if (!cu_rows || !chunk_start || !Q || !W || !v || !ml || !O || !dQ || !dW || !db || !workspace) return VLSA_EINVAL;
    if (T && (!logit_scale || !f || !g || !logits || !d_logits || !dT || !dlogit_scale || R < 1 || R > VLSA_MAX_R))
        return VLSA_EINVAL;
    if (!T && !d_f) return VLSA_EINVAL;
    if (B < 1 || P < 1 || P > VLSA_MAX_P || chunk_rows <= 0 || chunk_rows % kRowTile || total_chunks < 0)
        return VLSA_EINVAL;
    if (!dtype_ok(x_dtype)) return VLSA_EUNSUPPORTED;
    if (q_prenorm != 0 && q_prenorm != 1) return VLSA_EINVAL;
    if (total_chunks > 0 && (!X || total_rows <= 0)) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes) return VLSA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const float* df = d_f;
    if (T) {
        head_bwd_kernel<<<B, 256, 0, st>>>(f, g, T, R, logit_scale, logits, d_logits, d_g, d_f, ws.df, ws.dls_part);
        VLSA_CUDA(cudaGetLastError());
        text_bwd_kernel<<<R, 256, 0, st>>>(T, R, g, d_logits, B, logit_scale, ws.dls_part, dT, dlogit_scale);
        VLSA_CUDA(cudaGetLastError());
        df = ws.df;
    }
    adapter_bwd_dw_kernel<<<VLSA_D / 8, 512, 0, st>>>(df, v, B, dW, db);
    VLSA_CUDA(cudaGetLastError());
    adapter_bwd_dv_kernel<<<VLSA_D / 4, 128, 0, st>>>(df, W, B, ws.dv);
    VLSA_CUDA(cudaGetLastError());
    delta_kernel<<<B, 256, 0, st>>>(ws.dv, O, P, ws.delta);
    VLSA_CUDA(cudaGetLastError());

    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.row_ranges = (x_dtype & VLSA_ROWS_RANGES) ? 1 : 0;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.scale = coattn_scale;
    prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;
    prm.dv = ws.dv; prm.ml = ml; prm.delta = ws.delta; prm.q_prenorm = q_prenorm;
    int rc = 0;
    const AggKernel kern = total_chunks > 0 ? agg_kernel_choice(P, x_dtype) : kAggSimt;
    if (kern == kAggTc) {
        rc = launch_agg_tc<true>(prm, P, st);
        if (rc) return rc;
    } else if (kern == kAggBf16) {
        rc = launch_agg_bf16<true>(prm, P, total_rows, st);
        if (rc) return rc;
    } else if (kern == kAggSplit) {
        rc = launch_agg_split<true>(prm, P, st);
        if (rc) return rc;
    } else {
        VLSA_DISPATCH_P(P, {
            if ((x_dtype & VLSA_DTYPE_MASK) == VLSA_DTYPE_F32) rc = launch_agg<kP, 1, float>(prm, st);
            else rc = launch_agg<kP, 1, __nv_bfloat16>(prm, st);
            if (rc) return rc;
        });
    }
    const int Sb = merge_bwd_splits(total_chunks, P);
    if (Sb > 0) {
        merge_bwd_split_kernel<<<dim3(Sb, P), 128, 0, st>>>(ws.part_O, total_chunks, P, Sb, ws.l2_O);
        VLSA_CUDA(cudaGetLastError());
        merge_bwd_kernel<<<P, 512, 0, st>>>(ws.l2_O, Sb, P, Q, dQ, q_prenorm);
    } else {
        merge_bwd_kernel<<<P, 512, 0, st>>>(ws.part_O, total_chunks, P, Q, dQ, q_prenorm);
    }
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_agg_pooled_fwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                        int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale,
                        void* workspace, size_t workspace_bytes, float* out_ml, float* out_O, void* stream) {
    if (B == 0) return 0;
    if (!cu_rows || !chunk_start || !Q || !out_ml || !out_O) return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || chunk_rows <= 0 || chunk_rows % kRowTile || total_chunks < 0)
        return VLSA_EINVAL;
    if (q_prenorm != 0 && q_prenorm != 1) return VLSA_EINVAL;
    if (!dtype_ok(x_dtype)) return VLSA_EUNSUPPORTED;
    if (total_chunks > 0 && (!X || !workspace || total_rows <= 0)) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if (total_chunks > 0 && (base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes)
        return VLSA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.row_ranges = (x_dtype & VLSA_ROWS_RANGES) ? 1 : 0;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.q_prenorm = q_prenorm;
    prm.scale = coattn_scale; prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;
    int rc = launch_agg_fwd(prm, P, x_dtype, total_rows, st);
    if (rc) return rc;
    const int S = merge_fwd_splits(total_chunks, B, P);
    if (S > 0) {
        merge_fwd_split_kernel<<<dim3(S, P, B), 128, 0, st>>>(ws.part_m, ws.part_l, ws.part_O, chunk_start, P, S,
                                                                ws.l2_m, ws.l2_l, ws.l2_O);
        VLSA_CUDA(cudaGetLastError());
    }
    merge_fwd_kernel<<<dim3(B, VLSA_D / 64), dim3(64, P), 0, st>>>(S > 0 ? ws.l2_m : ws.part_m, S > 0 ? ws.l2_l : ws.part_l,
                                                                   S > 0 ? ws.l2_O : ws.part_O, chunk_start, P, S,
                                                                   out_ml, out_O, nullptr);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_agg_pooled_bwd(const void* X, int x_dtype, int64_t total_rows, const int64_t* cu_rows, const int32_t* chunk_start, int B,
                        int chunk_rows, int total_chunks, const float* Q, int P, int q_prenorm, float coattn_scale,
                        const float* ml, const float* O, const float* d_O, void* workspace, size_t workspace_bytes,
                        float* dQ, void* stream) {
    if (!cu_rows || !chunk_start || !Q || !ml || !O || !d_O || !dQ || !workspace) return VLSA_EINVAL;
    if (B < 1 || P < 1 || P > VLSA_MAX_P || chunk_rows <= 0 || chunk_rows % kRowTile || total_chunks < 0)
        return VLSA_EINVAL;
    if (q_prenorm != 0 && q_prenorm != 1) return VLSA_EINVAL;
    if (!dtype_ok(x_dtype)) return VLSA_EUNSUPPORTED;
    (void)total_rows;                                          // this pass runs on the CUDA-core kernel for every P
    if ((x_dtype & VLSA_DTYPE_MASK) == VLSA_DTYPE_SPLIT16) return VLSA_EUNSUPPORTED;   // ... which reads plain rows
    if (total_chunks > 0 && !X) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    AggWorkspace ws = carve(reinterpret_cast<void*>(base), total_chunks, B, P);
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes) return VLSA_EWORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    delta_gen_kernel<<<B, 256, 0, st>>>(d_O, O, P, ws.delta);
    VLSA_CUDA(cudaGetLastError());
    AggParams prm{};
    prm.X = X; prm.cu_rows = reinterpret_cast<const long long*>(cu_rows); prm.chunk_start = chunk_start;
    prm.row_ranges = (x_dtype & VLSA_ROWS_RANGES) ? 1 : 0;
    prm.B = B; prm.chunk_rows = chunk_rows; prm.total_chunks = total_chunks; prm.Q = Q; prm.q_prenorm = q_prenorm;
    prm.scale = coattn_scale; prm.part_m = ws.part_m; prm.part_l = ws.part_l; prm.part_O = ws.part_O;
    prm.p_stride = P;
    // 2 P + 1 dots per row and 2 P resident rows outgrow registers and shared memory beyond P = 8 (one launch at
    // P = 12: 4.0 ms for 32 x 50k rows, against 1.3 ms at P = 8), and prototypes are independent in this pass: serve
    // them in balanced groups of <= 8, one launch and one read of X per group, each with its own slice of the partials.
    const int group = kGenGroup;
    const int ngroups = (P + group - 1) / group;
    for (int gi = 0, p0 = 0; gi < ngroups; ++gi) {
        const int pg = (P - p0 + (ngroups - gi) - 1) / (ngroups - gi);       // balanced: 12 -> 6+6, 9 -> 5+4
        float* part = ws.part_O + size_t(total_chunks) * p0 * VLSA_D;
        float* l2 = ws.l2_O + size_t(merge_bwd_splits(total_chunks, P)) * p0 * VLSA_D;
        prm.Q = Q + size_t(p0) * VLSA_D; prm.dv = d_O + size_t(p0) * VLSA_D; prm.ml = ml + size_t(p0) * 2;
        prm.delta = ws.delta + p0; prm.part_O = part;
        int rc = 0;
        VLSA_DISPATCH_P(pg, {
            if ((x_dtype & VLSA_DTYPE_MASK) == VLSA_DTYPE_F32) rc = launch_agg<kP, 2, float>(prm, st);
            else rc = launch_agg<kP, 2, __nv_bfloat16>(prm, st);
            if (rc) return rc;
        });
        // the split count is the one the workspace was carved for (P prototypes); any split is a valid fixed order
        const int Sb = merge_bwd_splits(total_chunks, P);
        if (Sb > 0) {
            merge_bwd_split_kernel<<<dim3(Sb, pg), 128, 0, st>>>(part, total_chunks, pg, Sb, l2);
            VLSA_CUDA(cudaGetLastError());
            merge_bwd_kernel<<<pg, 512, 0, st>>>(l2, Sb, pg, prm.Q, dQ + size_t(p0) * VLSA_D, q_prenorm);
        } else {
            merge_bwd_kernel<<<pg, 512, 0, st>>>(part, total_chunks, pg, prm.Q, dQ + size_t(p0) * VLSA_D, q_prenorm);
        }
        VLSA_CUDA(cudaGetLastError());
        p0 += pg;
    }
    return 0;
}

int vlsa_agg_pooled_bwd_dx(const float* X, const int64_t* cu_rows, int B, int64_t max_rows, const float* Q, int P,
                           int q_prenorm, float coattn_scale, const float* ml, const float* O, const float* d_O,
                           float* out_dX, void* stream) {
    if (B == 0 || max_rows == 0) return 0;
    if (!X || !cu_rows || !Q || !ml || !O || !d_O || !out_dX) return VLSA_EINVAL;
    if (B < 0 || max_rows < 0 || P < 1 || P > VLSA_MAX_P || (q_prenorm != 0 && q_prenorm != 1)) return VLSA_EINVAL;
    const long long tiles = (max_rows + 127) / 128;
    if (tiles > 0x7fffffffLL || B > 65535) return VLSA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t(2 * P) * VLSA_D + 3 * P) * sizeof(float);
    VLSA_CUDA(cudaFuncSetAttribute(agg_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    agg_dx_kernel<<<dim3(unsigned(tiles), unsigned(B)), 128, smem, st>>>(X, reinterpret_cast<const long long*>(cu_rows), Q, P,
                                                                         q_prenorm, coattn_scale, ml, O, d_O, out_dX);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_attn_fwd(const void* X, int x_dtype, int64_t N, const float* Q, int P, int q_prenorm, float coattn_scale,
                  const float* ml, float* out_A, void* stream) {
    if (N == 0) return 0;
    if (!X || !Q || !out_A || N < 0 || P < 1 || P > VLSA_MAX_P) return VLSA_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t(P) * VLSA_D + 32 * (P + 1)) * sizeof(float);
    const unsigned grid = unsigned((N + 31) / 32);
    if (x_dtype != VLSA_DTYPE_F32 && x_dtype != VLSA_DTYPE_BF16) return VLSA_EUNSUPPORTED;
    const bool f32 = x_dtype == VLSA_DTYPE_F32;
    if (ml) {       // softmax over the N patches, normalisers from the forward
        if (f32) row_cosine_kernel<float, 0><<<grid, 256, smem, st>>>(static_cast<const float*>(X), N, Q, P, coattn_scale, nullptr, ml, out_A, q_prenorm);
        else row_cosine_kernel<__nv_bfloat16, 0><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(X), N, Q, P, coattn_scale, nullptr, ml, out_A, q_prenorm);
    } else {        // softmax over the P prototypes per patch
        if (f32) row_cosine_kernel<float, 2><<<grid, 256, smem, st>>>(static_cast<const float*>(X), N, Q, P, coattn_scale, nullptr, nullptr, out_A, q_prenorm);
        else row_cosine_kernel<__nv_bfloat16, 2><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(X), N, Q, P, coattn_scale, nullptr, nullptr, out_A, q_prenorm);
    }
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_interp_fwd(const float* O, const float* W, const float* bias, const float* T, int R, const float* f,
                    const float* logit_scale, int B, int P, float* out_sim, float* out_imp, float* out_probs,
                    void* stream) {
    if (B == 0) return 0;
    if (!O || !W || !bias || !T || !f || !logit_scale || !out_sim || !out_imp || !out_probs) return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || R < 1 || R > VLSA_MAX_R) return VLSA_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    interp_sim_kernel<<<B * P, 256, 0, st>>>(O, W, bias, T, R, f, P, out_sim);
    VLSA_CUDA(cudaGetLastError());
    interp_softmax_kernel<<<B, VLSA_MAX_R, 0, st>>>(out_sim, P, R, logit_scale, out_imp, out_probs);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_surv_loss_fwd_bwd(const float* logits, const int64_t* t, const int64_t* e, int B, int R,
                           const float* logit_scale, float w_ifmle, float w_emd, float alpha, float eps,
                           float inv_norm, int input_is_prob, float* out_loss, float* out_if, float* out_dlogits,
                           float* out_per_sample, void* stream) {
    if (!logits || !t || !e || !logit_scale || !out_loss || !out_per_sample) return VLSA_EINVAL;
    if (B < 1 || R < 1 || R > VLSA_MAX_R) return VLSA_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    surv_loss_kernel<<<(B + 3) / 4, 128, 0, st>>>(logits, reinterpret_cast<const long long*>(t),
                                                  reinterpret_cast<const long long*>(e), B, R, logit_scale, w_ifmle,
                                                  w_emd, alpha, eps, inv_norm, input_is_prob, out_if, out_dlogits,
                                                  out_per_sample);
    VLSA_CUDA(cudaGetLastError());
    surv_loss_reduce_kernel<<<1, 256, 0, st>>>(out_per_sample, B, w_ifmle, w_emd, inv_norm, out_loss);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

size_t vlsa_logit_pool_workspace_bytes(int64_t N, int R, int k) {
    if (N < 0 || R < 1 || R > VLSA_MAX_R || k < 0) return 0;
    return align_up(size_t(N) * R * sizeof(float), 256) + 512;
}

int vlsa_logit_pool_fwd(const void* X, int x_dtype, int64_t N, const float* T, int R, const float* logit_scale,
                        int mode, int k, void* workspace, size_t workspace_bytes, float* out_logits,
                        int64_t* out_pred, void* stream) {
    if (!X || !T || !logit_scale || !workspace || !out_logits || !out_pred) return VLSA_EINVAL;
    if (N < 1 || R < 1 || R > VLSA_MAX_R) return VLSA_EINVAL;
    if (mode != VLSA_POOL_MEAN && mode != VLSA_POOL_TOPK) return VLSA_EINVAL;
    if (mode == VLSA_POOL_TOPK && (k < 1 || k > 64)) return VLSA_EUNSUPPORTED;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    const size_t logits_bytes = align_up(size_t(N) * R * sizeof(float), 256);
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + logits_bytes + 4 > workspace_bytes) return VLSA_EWORKSPACE;
    float* patch_logits = reinterpret_cast<float*>(base);
    unsigned int* counter = reinterpret_cast<unsigned int*>(base + logits_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VLSA_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const size_t smem = (size_t(R) * VLSA_D + 32 * (R + 1)) * sizeof(float);
    const unsigned grid = unsigned((N + 31) / 32);
    if (x_dtype == VLSA_DTYPE_F32) {
        auto kern = row_cosine_kernel<float, 1>;
        VLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<grid, 256, smem, st>>>(static_cast<const float*>(X), N, T, R, 0.f, logit_scale, nullptr, patch_logits, 0);
    } else if (x_dtype == VLSA_DTYPE_BF16) {
        auto kern = row_cosine_kernel<__nv_bfloat16, 1>;
        VLSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        kern<<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(X), N, T, R, 0.f, logit_scale, nullptr,
                                      patch_logits, 0);
    } else {
        return VLSA_EUNSUPPORTED;
    }
    VLSA_CUDA(cudaGetLastError());
    logit_pool_kernel<<<R, 1024, 0, st>>>(patch_logits, N, R, mode, k, out_logits,
                                          reinterpret_cast<long long*>(out_pred), counter);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

size_t vlsa_feat_pool_workspace_bytes(void) { return size_t(296) * VLSA_D * sizeof(float) + 512; }

int vlsa_feat_pool_fwd(const void* X, int x_dtype, int64_t N, int mode, const float* T, int R, const float* logit_scale,
                       void* workspace, size_t workspace_bytes, float* out_f, float* out_g, float* out_logits,
                       float* out_Tn, void* stream) {
    if (!X || !T || !logit_scale || !workspace || !out_f || !out_g || !out_logits) return VLSA_EINVAL;
    if (N < 1 || R < 1 || R > VLSA_MAX_R || (mode != VLSA_POOL_MEAN && mode != VLSA_POOL_MAX)) return VLSA_EINVAL;
    if (x_dtype != VLSA_DTYPE_F32 && x_dtype != VLSA_DTYPE_BF16) return VLSA_EUNSUPPORTED;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    const int G = int(N < 296 ? N : 296);                      // two CTAs per SM; fixed by N alone (bit-stable)
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + size_t(G) * VLSA_D * sizeof(float) > workspace_bytes)
        return VLSA_EWORKSPACE;
    float* part = reinterpret_cast<float*>(base);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (x_dtype == VLSA_DTYPE_F32) {
        if (mode == VLSA_POOL_MEAN) feat_pool_partial_kernel<float, 0><<<G, 128, 0, st>>>(static_cast<const float*>(X), N, part);
        else feat_pool_partial_kernel<float, 1><<<G, 128, 0, st>>>(static_cast<const float*>(X), N, part);
    } else {
        if (mode == VLSA_POOL_MEAN) feat_pool_partial_kernel<__nv_bfloat16, 0><<<G, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(X), N, part);
        else feat_pool_partial_kernel<__nv_bfloat16, 1><<<G, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(X), N, part);
    }
    VLSA_CUDA(cudaGetLastError());
    if (mode == VLSA_POOL_MEAN) feat_pool_final_kernel<0><<<1, 128, 0, st>>>(part, G, N, out_f);
    else feat_pool_final_kernel<1><<<1, 128, 0, st>>>(part, G, N, out_f);
    VLSA_CUDA(cudaGetLastError());
    head_fwd_kernel<<<1, 256, 0, st>>>(out_f, T, R, logit_scale, out_g, out_logits, nullptr, out_Tn);
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

int vlsa_row_normalize(const void* X, int x_dtype, int64_t N, float* out, void* stream) {
    if (N == 0) return 0;
    if (!X || !out || N < 0) return VLSA_EINVAL;
    const long long blocks = (N + 7) / 8;
    if (blocks > 0x7fffffffLL) return VLSA_EUNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (x_dtype == VLSA_DTYPE_F32) row_normalize_kernel<float><<<unsigned(blocks), 256, 0, st>>>(static_cast<const float*>(X), N, out);
    else if (x_dtype == VLSA_DTYPE_BF16) row_normalize_kernel<__nv_bfloat16><<<unsigned(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(X), N, out);
    else return VLSA_EUNSUPPORTED;
    VLSA_CUDA(cudaGetLastError());
    return 0;
}

// ---- host-buffer entry ---------------------------------------------------------------------------
struct HostWs {
    void* X; long long* cu_rows; int* chunk_start; void* agg; size_t agg_bytes;
    float *v, *f, *g, *logits, *inc, *ml; size_t bytes;
};
static int max_chunks_bound(int B) { return 16 * device_sm_count() + B + 8; }
static HostWs carve_host(void* base, int64_t total_rows, int B, int P, int esize) {
    HostWs w; size_t off = 0;
    auto take = [&](size_t nbytes) {
        void* p = base ? static_cast<void*>(static_cast<char*>(base) + off) : nullptr;
        off += align_up(nbytes ? nbytes : 1, 256);
        return p;
    };
    w.X = take(size_t(total_rows) * VLSA_D * esize);
    w.cu_rows = static_cast<long long*>(take(size_t(B + 1) * sizeof(long long)));
    w.chunk_start = static_cast<int*>(take(size_t(B + 1) * sizeof(int)));
    w.agg_bytes = carve(nullptr, max_chunks_bound(B), B, P).bytes + 256;
    w.agg = take(w.agg_bytes);
    w.v = static_cast<float*>(take(size_t(B) * VLSA_D * 4));
    w.f = static_cast<float*>(take(size_t(B) * VLSA_D * 4));
    w.g = static_cast<float*>(take(size_t(B) * VLSA_D * 4));
    w.logits = static_cast<float*>(take(size_t(B) * VLSA_MAX_R * 4));
    w.inc = static_cast<float*>(take(size_t(B) * VLSA_MAX_R * 4));
    w.ml = static_cast<float*>(take(size_t(B) * VLSA_MAX_P * 2 * 4));
    w.bytes = off;
    return w;
}

size_t vlsa_forward_host_workspace_bytes(int64_t total_rows, int B, int P, int x_dtype) {
    if (total_rows < 0 || B < 0 || P < 1 || P > VLSA_MAX_P) return 0;
    if (x_dtype != VLSA_DTYPE_F32 && x_dtype != VLSA_DTYPE_BF16) return 0;
    return carve_host(nullptr, total_rows, B, P, x_dtype == VLSA_DTYPE_F32 ? 4 : 2).bytes + 256;
}

int vlsa_forward_host(const void* X_host, int x_dtype, const int64_t* cu_rows_host, int B, const float* Q, int P,
                      int q_prenorm, float coattn_scale, const float* W, const float* bias, const float* T, int R,
                      const float* logit_scale, void* workspace, size_t workspace_bytes, float* out_if_host,
                      float* out_logits_host, void* stream_compute, void* stream_copy) {
    if (B == 0) return 0;
    if (!cu_rows_host || !Q || !W || !bias || !T || !logit_scale || !workspace || !out_if_host) return VLSA_EINVAL;
    if (B < 0 || P < 1 || P > VLSA_MAX_P || R < 1 || R > VLSA_MAX_R || (q_prenorm != 0 && q_prenorm != 1)) return VLSA_EINVAL;
    if (x_dtype != VLSA_DTYPE_F32 && x_dtype != VLSA_DTYPE_BF16) return VLSA_EUNSUPPORTED;
    const int esize = x_dtype == VLSA_DTYPE_F32 ? 4 : 2;
    const int64_t total_rows = cu_rows_host[B];
    if (total_rows < 0 || (total_rows > 0 && !X_host)) return VLSA_EINVAL;
    uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
    HostWs ws = carve_host(reinterpret_cast<void*>(base), total_rows, B, P, esize);
    if ((base - reinterpret_cast<uintptr_t>(workspace)) + ws.bytes > workspace_bytes) return VLSA_EWORKSPACE;

    int chunk_rows = 0;
    int32_t cs_stack[1025];                      // a step is 32 bags; larger batches take the heap
    int32_t* cs_host = B + 1 <= 1025 ? cs_stack : static_cast<int32_t*>(malloc(size_t(B + 1) * sizeof(int32_t)));
    if (!cs_host) return VLSA_EINVAL;
    int rc = vlsa_agg_plan(cu_rows_host, B, 0, &chunk_rows, cs_host);
    const int total_chunks = rc == 0 ? cs_host[B] : 0;
    if (rc == 0 && total_chunks > max_chunks_bound(B)) rc = VLSA_EWORKSPACE;
    cudaStream_t sc = static_cast<cudaStream_t>(stream_compute), sx = static_cast<cudaStream_t>(stream_copy);
    // Two streams: the copies must not overwrite a workspace the previous call's kernels are still reading
    // (compute -> copy), and the kernels must wait for the copies (copy -> compute).  One stream needs neither.
    cudaEvent_t ev = nullptr;
    if (rc == 0 && sx != sc) {
        rc = static_cast<int>(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        if (rc == 0) rc = static_cast<int>(cudaEventRecord(ev, sc));
        if (rc == 0) rc = static_cast<int>(cudaStreamWaitEvent(sx, ev, 0));
    }
    if (rc == 0 && total_rows > 0)
        rc = static_cast<int>(cudaMemcpyAsync(ws.X, X_host, size_t(total_rows) * VLSA_D * esize, cudaMemcpyHostToDevice, sx));
    // small pageable arrays: the runtime stages them before returning, so cs_host may go out of scope below
    if (rc == 0) rc = static_cast<int>(cudaMemcpyAsync(ws.cu_rows, cu_rows_host, size_t(B + 1) * 8, cudaMemcpyHostToDevice, sx));
    if (rc == 0) rc = static_cast<int>(cudaMemcpyAsync(ws.chunk_start, cs_host, size_t(B + 1) * 4, cudaMemcpyHostToDevice, sx));
    if (rc == 0 && ev) {
        rc = static_cast<int>(cudaEventRecord(ev, sx));
        if (rc == 0) rc = static_cast<int>(cudaStreamWaitEvent(sc, ev, 0));
    }
    if (rc == 0)
        rc = vlsa_agg_fwd(ws.X, x_dtype, total_rows, reinterpret_cast<const int64_t*>(ws.cu_rows), ws.chunk_start, B,
                          chunk_rows, total_chunks, Q, P, q_prenorm, coattn_scale, W, bias, T, R, logit_scale, ws.agg,
                          ws.agg_bytes, ws.v, ws.f, ws.g, ws.logits, ws.inc, ws.ml, nullptr, nullptr, sc);
    if (rc == 0) rc = static_cast<int>(cudaMemcpyAsync(out_if_host, ws.inc, size_t(B) * R * 4, cudaMemcpyDeviceToHost, sc));
    if (rc == 0 && out_logits_host)
        rc = static_cast<int>(cudaMemcpyAsync(out_logits_host, ws.logits, size_t(B) * R * 4, cudaMemcpyDeviceToHost, sc));
    if (ev) cudaEventDestroy(ev);               // deferred by the runtime until the event has completed
    if (cs_host != cs_stack) free(cs_host);
    return rc;
}

}  // extern "C"
