// tvq_api.cu — the C ABI declared in include/tvq.h: argument checks, kernel selection, launches.
// Built for sm_100a only (see __graft_entry__.build); no torch, no C++ exceptions across the ABI.
#include "../../include/tvq.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>
#include <string.h>

#include "tvq_aux.cuh"
#include "tvq_common.cuh"
#include "tvq_frontend.cuh"
#include "tvq_fwd_simt.cuh"
#include "tvq_fwd_stream.cuh"
#include "tvq_fwd_umma.cuh"
#include "tvq_maskgit.cuh"
#include "tvq_snake.cuh"

using namespace tvq;

namespace {

// NVTX range around every launching ABI call (the tracing equivalent of the reference's MLflow step
// logging): an Nsight Systems / ncu --nvtx timeline shows which call of the reference-shaped API each kernel belongs to.
// nvtx3 is header-only and resolves the profiler's injection library lazily: without a profiler a push/pop is a load and
// a predictable branch.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define TVQ_RANGE(name) NvtxRange tvq_nvtx_range_(name)

constexpr int kMaxDevices = 64;
struct DeviceInfo {
    int checked = 0;
    int ok = 0;
    int index = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
};
DeviceInfo g_dev[kMaxDevices];

// cudaFuncSetAttribute / occupancy are per DEVICE: every host-side cache of them is indexed by the device the call runs on
// (one process may drive several GPUs).  Benign if raced: the worst case is a repeated, idempotent runtime call.
struct PerDeviceInt {
    int v[kMaxDevices];
    explicit PerDeviceInt(int init) { for (int i = 0; i < kMaxDevices; ++i) v[i] = init; }
    int& operator[](int dev) { return v[dev]; }
};

template <typename Kern>
int ensure_dynamic_smem(Kern kern, PerDeviceInt& cache, int dev, size_t bytes) {
    if ((long long)bytes > (long long)cache[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
        cache[dev] = (int)bytes;
    }
    return TVQ_OK;
}

int device_info(DeviceInfo** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= kMaxDevices) return TVQ_ERR_DEVICE;
    DeviceInfo& di = g_dev[dev];
    if (!di.checked) {
        di.index = dev;
        int major = 0;
        if ((e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&di.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        di.ok = (major == 10);
        di.checked = 1;
    }
    *out = &di;
    return di.ok ? TVQ_OK : TVQ_ERR_DEVICE;
}

// How long a data-parallel kernel waits for a peer's statistics before it reports the step in the exchange buffer's error
// word and carries on (tvq_set_peer_timeout; default 30 min — longer than torch's NCCL watchdog, so a slow rank is never
// turned into an error by this library first).  Process-wide, passed by value with every launch.
unsigned long long g_peer_timeout_ns = 1800ull * 1000000000ull;

// One-shot launch hint of the calling thread (tvq_hint_max_ctas): the NEXT resident-codebook forward launched from this
// thread uses at most that many CTAs, then the hint is cleared.  The kernel is persistent with a dynamic tile scheduler, so
// any grid size computes the same result; a caller that runs two independent quantisers on two streams (the LF and HF
// codebooks of a stage-1 step) gives each a share of the SMs so that the two launches are resident TOGETHER instead of one
// after the other (each CTA needs a whole SM's shared memory).
thread_local int t_hint_max_ctas = 0;
// One-shot hint (tvq_hint_defer_exchange): the next data-parallel fused train step of this thread only publishes its
// statistics; the caller completes the step with tvq_ema_finalize_dp.
thread_local int t_hint_defer_exchange = 0;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int pad_dim(int d) { return d <= 32 ? 32 : d <= 64 ? 64 : d <= 128 ? 128 : 256; }

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TVQ_OK : (int)e;
}

template <int DP, bool TRAIN>
int launch_fwd_simt(const FwdParams& p, const SmemPlan& pl, const DeviceInfo& di, cudaStream_t stream) {
    auto kern = fwd_simt_kernel<DP, TRAIN>;
    static PerDeviceInt configured_smem(-1);   // per instantiation and device
    static PerDeviceInt occ_smem(-1), occ_val(0);
    if (int rc = ensure_dynamic_smem(kern, configured_smem, di.index, pl.total)) return rc;
    if (occ_smem[di.index] != pl.total) {
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, pl.total);
        if (e != cudaSuccess) return (int)e;
        occ_val[di.index] = o;
        occ_smem[di.index] = pl.total;
    }
    const int occ = occ_val[di.index];
    if (occ < 1) return TVQ_ERR_UNSUPPORTED;
    long long grid = (long long)occ * di.sm_count;
    if (grid > p.num_tiles) grid = p.num_tiles;
    kern<<<(unsigned)grid, kThreads, pl.total, stream>>>(p);
    return launch_status();
}

template <bool TRAIN>
int dispatch_fwd_simt(int dp, const FwdParams& p, const SmemPlan& pl, const DeviceInfo& di, cudaStream_t s) {
    switch (dp) {
        case 32: return launch_fwd_simt<32, TRAIN>(p, pl, di, s);
        case 64: return launch_fwd_simt<64, TRAIN>(p, pl, di, s);
        case 128: return launch_fwd_simt<128, TRAIN>(p, pl, di, s);
        case 256: return launch_fwd_simt<256, TRAIN>(p, pl, di, s);
    }
    return TVQ_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 path: TMA tensor map over x, stage count from the shared-memory budget, launch.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// x viewed as a [n, d] fp32 tensor; one box = 32 floats (128 bytes, the swizzle span) x `rows` latents.
int make_x_tensor_map(CUtensorMap* tm, const float* x, int64_t n, int d, int rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return TVQ_ERR_DEVICE;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t gstride[1] = {(cuuint64_t)d * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TVQ_OK : TVQ_ERR_BAD_ARG;
}

template <int DP, int KP, bool TRAIN, bool FULLD, bool XCF = false>
int launch_fwd_umma_impl(FwdParams p, const DeviceInfo& di, cudaStream_t stream) {
    auto kern = fwd_umma_kernel<DP, KP, TRAIN, FULLD, XCF>;
    const UmmaPlan fixed = make_umma_plan(DP, KP, 0);
    int stages = (di.max_smem_optin - fixed.total) / (kUM * DP * 4);
    if (stages > kUMaxStages) stages = kUMaxStages;
    if (stages < 2) return TVQ_ERR_UNSUPPORTED;
    if (TRAIN && stages * (kUM * DP * 4) < 4 * KP * 32 * 16) return TVQ_ERR_UNSUPPORTED;   // end-of-kernel dump area
    const UmmaPlan pl = make_umma_plan(DP, KP, stages);
    static PerDeviceInt configured_smem(-1);
    if (int rc = ensure_dynamic_smem(kern, configured_smem, di.index, pl.total)) return rc;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));                       // channels-first x: loaded with cp.async, no tensor map
    if (!XCF) {
        int rc = make_x_tensor_map(&tm, p.x, p.n, p.d, kUM);
        if (rc != TVQ_OK) return rc;
    }
    p.num_tiles = (int)((p.n + kUM - 1) / kUM);
    // Data-parallel: the last CTA of a launch holds its SM while it waits for the peers' statistics.  One SM is left out
    // of the grid so that a second exchange kernel running next to it (the other codebook of the step, on another
    // stream) can always get ALL of its CTAs resident and publish its own statistics — whatever order the ranks
    // happen to start the two kernels in, nobody waits for an SM held by a waiting CTA.
    int cap = di.sm_count - (p.dp_world > 1 && di.sm_count > 1 ? 1 : 0);
    if (t_hint_max_ctas > 0 && t_hint_max_ctas < cap) cap = t_hint_max_ctas;
    t_hint_max_ctas = 0;
    int grid = p.num_tiles < cap ? p.num_tiles : cap;
    kern<<<grid, kUThreads, pl.total, stream>>>(tm, p, stages);
    return launch_status();
}

template <int DP, int KP, bool TRAIN>
int launch_fwd_umma(const FwdParams& p, const DeviceInfo& di, cudaStream_t stream) {
    if (p.x_hw > 0) return launch_fwd_umma_impl<DP, KP, TRAIN, false, true>(p, di, stream);
    if (DP == 128 && p.d == 128) return launch_fwd_umma_impl<DP, KP, TRAIN, (DP == 128)>(p, di, stream);
    return launch_fwd_umma_impl<DP, KP, TRAIN, false>(p, di, stream);
}

template <bool TRAIN>
int dispatch_fwd_umma(int dp, int kp, const FwdParams& p, const DeviceInfo& di, cudaStream_t s) {
    if (dp == 64) {
        if (kp == 16) return launch_fwd_umma<64, 16, TRAIN>(p, di, s);
        if (kp == 32) return launch_fwd_umma<64, 32, TRAIN>(p, di, s);
        if constexpr (!TRAIN) { if (kp == 64) return launch_fwd_umma<64, 64, false>(p, di, s); }
    } else if (dp == 128) {
        if (kp == 16) return launch_fwd_umma<128, 16, TRAIN>(p, di, s);
        if (kp == 32) return launch_fwd_umma<128, 32, TRAIN>(p, di, s);
        if constexpr (!TRAIN) { if (kp == 64) return launch_fwd_umma<128, 64, false>(p, di, s); }
    }
    return TVQ_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// Streamed-codebook tcgen05 path (any k): TMA tensor map over the bf16 codebook copy in the workspace.
inline int stream_dp(int d) { return d <= 64 ? 64 : d <= 128 ? 128 : 256; }
// |e|^2 table length: padded to a multiple of 256 (pad = +BIG) so that the streamed path can bulk-copy whole slices
inline size_t e2_len(int k) { return ((size_t)(k > 0 ? k : 0) + 255) & ~(size_t)255; }
inline size_t ws_bf16_offset(int k, int d) {
    const size_t kk = (size_t)k, dd = (size_t)d;
    size_t o = sizeof(WsHeader) + e2_len(k) * sizeof(float) + (size_t)TVQ_STATS_LEN(kk, dd) * sizeof(float);
    return (o + 255) & ~(size_t)255;
}

inline size_t ws_e2h_offset(int k, int d) {
    const size_t o = ws_bf16_offset(k, d) + (size_t)k * (size_t)stream_dp(d) * 2;
    return (o + 255) & ~(size_t)255;
}

// |e|^2 / 2 pieces [roundup(k, 256), 16] bf16: one box = 16 elements (32 bytes, the SWIZZLE_32B span) x nt codes
int make_e2_tensor_map(CUtensorMap* tm, const void* e2h, int k, int nt) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return TVQ_ERR_DEVICE;
    cuuint64_t gdim[2] = {16u, (cuuint64_t)e2_len(k)};
    cuuint64_t gstride[1] = {32u};
    cuuint32_t box[2] = {16u, (cuuint32_t)nt};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(e2h), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TVQ_OK : TVQ_ERR_BAD_ARG;
}

int make_cb_tensor_map(CUtensorMap* tm, const void* cbh, int k, int dp, int nt) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return TVQ_ERR_DEVICE;
    cuuint64_t gdim[2] = {(cuuint64_t)dp, (cuuint64_t)k};
    cuuint64_t gstride[1] = {(cuuint64_t)dp * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)nt};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(cbh), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TVQ_OK : TVQ_ERR_BAD_ARG;
}

template <int DP, int NT, bool TRAIN, int CG, int SP, int SETS = 1>
int launch_fwd_stream_impl(FwdParams p, const void* cbh, const void* e2h, const DeviceInfo& di, cudaStream_t stream) {
    auto kern = fwd_stream_kernel<DP, NT, TRAIN, CG, SP, SETS>;
    constexpr int kSThreads = stream_threads(SP * SETS);
    // staged x blocks for the converter (d <= 128): 3 blocks of 8 KB per converter warp at d <= 64, 2 at d <= 128;
    // TVQ_STREAM_XD=0..4 in the environment overrides (experiments)
    static int forced_xd = -2;
    if (forced_xd == -2) {
        const char* e = getenv("TVQ_STREAM_XD");
        forced_xd = e ? atoi(e) : -1;
    }
    int xdepth = DP == 64 ? 3 : DP == 128 ? 2 : 0;
    if (forced_xd >= 0 && forced_xd <= kSXMaxDepth && DP <= 128) xdepth = forced_xd;
    if (DP <= 128 && xdepth < 2) xdepth = 2;     // the staged converter path is compiled in for d <= 128
    const StreamPlan fixed = make_stream_plan(DP, NT, 0, CG, SP * SETS, xdepth);
    int stages = (di.max_smem_optin - fixed.total) / ((NT / CG) * 128);
    if (stages > kSMaxStages) stages = kSMaxStages;
    if (stages < 2) return TVQ_ERR_UNSUPPORTED;
    const StreamPlan pl = make_stream_plan(DP, NT, stages, CG, SP * SETS, xdepth);
    static PerDeviceInt configured_smem(-1);
    if (int rc = ensure_dynamic_smem(kern, configured_smem, di.index, pl.total)) return rc;
    CUtensorMap tm, tm2;
    int rc = make_cb_tensor_map(&tm, cbh, p.k, DP, NT / CG);
    if (rc != TVQ_OK) return rc;
    if ((rc = make_e2_tensor_map(&tm2, e2h, p.k, NT / CG)) != TVQ_OK) return rc;
    p.num_tiles = (int)((p.n + kSM - 1) / kSM);
    const __nv_bfloat16* e2h_bf = reinterpret_cast<const __nv_bfloat16*>(e2h);
    const int groups = (p.num_tiles + CG - 1) / CG, units = di.sm_count / CG;
    const int grid = CG * (groups < units ? groups : units);
    if (CG == 1) {
        kern<<<grid, kSThreads, pl.total, stream>>>(tm, tm2, p, stages, e2h_bf, xdepth);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kSThreads);
        cfg.dynamicSmemBytes = (size_t)pl.total;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm, tm2, p, stages, e2h_bf, xdepth);
        if (e != cudaSuccess) return (int)e;
    }
    return launch_status();
}

// CTA pairs pay off where the code stream is long enough to matter (k >= 1024) and there are at least two row tiles
// per pair of SMs; TVQ_STREAM_CG=1|2 in the environment overrides the choice (experiments).
inline int stream_cg(const FwdParams& p) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("TVQ_STREAM_CG");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 1 || forced == 2) return forced;
    return (p.k >= 512 && p.n > kSM) ? 2 : 1;
}

// Scan parts per quadrant (tvq_fwd_stream.cuh): 4 (20 warps) for d <= 64, and for d <= 128 from k = 2048 up; 2 (12 warps)
// otherwise — measured at 2^20 latents: 512 x 64 1.03 -> 0.79 ms, 4096 x 64 2.25 -> 1.88, 16384 x 64 6.0 -> 5.1, 2048 x 128
// 1.69 -> 1.54, 4096 x 128 2.28 -> 2.21, but 1024 x 128 1.12 -> 1.48 (at d = 128 the two converter warps, left with 64
// registers, are the floor for short code streams).  TVQ_STREAM_SP=2|4 in the environment overrides the choice for d <= 128.
inline int stream_sp(const FwdParams& p) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("TVQ_STREAM_SP");
        forced = e ? atoi(e) : 0;
    }
    if (p.d > 128) return 2;
    if (forced == 2 || forced == 4) return forced;
    return 4;
}

template <bool TRAIN>
int dispatch_fwd_stream(const FwdParams& p, const void* cbh, const void* e2h, const DeviceInfo& di, cudaStream_t s) {
    const int cg = stream_cg(p), sp = stream_sp(p);
    switch (stream_dp(p.d)) {
        case 64:
            if (sp == 4) return cg == 2 ? launch_fwd_stream_impl<64, 256, TRAIN, 2, 4>(p, cbh, e2h, di, s) : launch_fwd_stream_impl<64, 256, TRAIN, 1, 4>(p, cbh, e2h, di, s);
            return cg == 2 ? launch_fwd_stream_impl<64, 256, TRAIN, 2, 2>(p, cbh, e2h, di, s) : launch_fwd_stream_impl<64, 256, TRAIN, 1, 2>(p, cbh, e2h, di, s);
        case 128:
            if (sp == 4) return cg == 2 ? launch_fwd_stream_impl<128, 256, TRAIN, 2, 4>(p, cbh, e2h, di, s) : launch_fwd_stream_impl<128, 256, TRAIN, 1, 4>(p, cbh, e2h, di, s);
            return cg == 2 ? launch_fwd_stream_impl<128, 256, TRAIN, 2, 2>(p, cbh, e2h, di, s) : launch_fwd_stream_impl<128, 256, TRAIN, 1, 2>(p, cbh, e2h, di, s);
        case 256: return cg == 2 ? launch_fwd_stream_impl<256, 256, TRAIN, 2, 2>(p, cbh, e2h, di, s) : launch_fwd_stream_impl<256, 128, TRAIN, 1, 2>(p, cbh, e2h, di, s);
    }
    return TVQ_ERR_UNSUPPORTED;
}

int launch_prep(const float* cb, int k, int d, float* e2, WsHeader* hdr, float* stats, int64_t stats_len, void* cbh, int dp,
                void* e2h, const DeviceInfo& di, cudaStream_t stream) {
    int64_t work = stats_len / 4 > (int64_t)e2_len(k) * 32 ? stats_len / 4 : (int64_t)e2_len(k) * 32;
    if (cbh && (int64_t)k * (dp / 4) > work) work = (int64_t)k * (dp / 4);
    int blocks = (int)((work + 255) / 256);
    if (blocks > 4 * di.sm_count) blocks = 4 * di.sm_count;
    if (blocks < 1) blocks = 1;
    prep_kernel<<<blocks, 256, 0, stream>>>(cb, k, d, e2, hdr, stats, stats_len, reinterpret_cast<__nv_bfloat16*>(cbh), dp,
                                            reinterpret_cast<__nv_bfloat16*>(e2h));
    return launch_status();
}

}  // namespace

extern "C" {

#ifdef TVQ_STREAM_PROF
__attribute__((visibility("default"))) int tvq_debug_stream_prof(unsigned long long* out64) {
    return (int)cudaMemcpyFromSymbol(out64, g_sprof, sizeof(unsigned long long) * 128);
}
#endif
#ifdef TVQ_PROFILE_PHASES
__attribute__((visibility("default"))) int tvq_debug_phases(unsigned long long* out32) {
    return (int)cudaMemcpyFromSymbol(out32, g_phase_clk, sizeof(unsigned long long) * 32);
}
__attribute__((visibility("default"))) int tvq_debug_phases_all(unsigned long long* out24) {
    return (int)cudaMemcpyFromSymbol(out24, g_phase_all, sizeof(unsigned long long) * 24);
}
__attribute__((visibility("default"))) int tvq_debug_tiles(unsigned long long* out32) {
    return (int)cudaMemcpyFromSymbol(out32, g_tile_clk, sizeof(unsigned long long) * 32);
}
__attribute__((visibility("default"))) int tvq_debug_gt(unsigned long long* out4, int reset) {
    if (reset) {
        unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
        return (int)cudaMemcpyToSymbol(g_gt, init, sizeof(init));
    }
    return (int)cudaMemcpyFromSymbol(out4, g_gt, sizeof(unsigned long long) * 4);
}
#endif

int tvq_abi_version(void) { return 2; }

int tvq_hint_max_ctas(int max_ctas) {
    t_hint_max_ctas = max_ctas > 0 ? max_ctas : 0;
    return TVQ_OK;
}

int tvq_hint_defer_exchange(int defer) {
    t_hint_defer_exchange = (defer == 1 || defer == 2) ? defer : 0;
    return TVQ_OK;
}

int tvq_set_peer_timeout(double seconds) {
    if (!(seconds >= 0.0) || seconds > 1e9) return TVQ_ERR_BAD_ARG;
    g_peer_timeout_ns = (unsigned long long)(seconds * 1e9);
    return TVQ_OK;
}

const char* tvq_error_string(int code) {
    switch (code) {
        case TVQ_OK: return "ok";
        case TVQ_ERR_UNSUPPORTED: return "tvq: unsupported shape (need 1 <= d <= 256, d % 4 == 0, k >= 1, k*d < 2^31)";
        case TVQ_ERR_BAD_ARG: return "tvq: bad argument (null or misaligned pointer, or workspace too small)";
        case TVQ_ERR_DEVICE: return "tvq: current device is not an sm_100 (B200) GPU";
    }
    return code > 0 ? cudaGetErrorString((cudaError_t)code) : "tvq: unknown error";
}

int tvq_device_check(int device, int* sm_count) {
    int major = 0, sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (sm_count) *sm_count = sms;
    return major == 10 ? TVQ_OK : TVQ_ERR_DEVICE;
}

size_t tvq_workspace_bytes(int64_t n, int k, int d) {
    (void)n;
    const size_t kk = k > 0 ? (size_t)k : 0, dd = d > 0 ? (size_t)d : 0;
    // header | |e|^2 table | private statistics scratch of tvq_train_step | bf16 codebook copy (streamed tcgen05 path)
    if (kk == 0 || dd == 0) return sizeof(WsHeader) + 256;
    // ... | bf16 codebook copy | |e|^2 / 2 pieces (both operands of the streamed tcgen05 path)
    return ws_e2h_offset((int)kk, (int)dd) + e2_len((int)kk) * 32 + 256;
}

}  // extern "C"

namespace {
int forward_impl(const float* x, const float* codebook, int64_t n, int k, int d, unsigned flags,
                 float commitment_weight, int64_t* idx, float* q, float* stats, float* scalars, void* workspace,
                 size_t workspace_bytes, void* stream_, int q_hw, int x_hw = 0) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || d > 256 || (d & 3) || n < 0 || (int64_t)k * d >= (int64_t(1) << 31)) return TVQ_ERR_UNSUPPORTED;
    if (!x || !codebook || !idx || !stats || !scalars || !workspace) return TVQ_ERR_BAD_ARG;
    if (!aligned16(x) || !aligned16(codebook) || !aligned16(stats) || !aligned16(workspace) || (q && !aligned16(q)))
        return TVQ_ERR_BAD_ARG;
    if ((flags & TVQ_F_WRITE_Q) && !q) return TVQ_ERR_BAD_ARG;
    if (workspace_bytes < tvq_workspace_bytes(n, k, d)) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;

    WsHeader* hdr = reinterpret_cast<WsHeader*>(workspace);
    float* e2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + sizeof(WsHeader));
    const bool train = flags & TVQ_F_TRAIN;
    const int64_t stats_len = TVQ_STATS_LEN(k, d);

    // Streamed-codebook tcgen05 path: every shape the resident-codebook path does not take.
    const bool resident = k <= (train ? 32 : 64) && d <= 128;
    if (q_hw > 0 && (!resident || (flags & (TVQ_F_EXACT | TVQ_F_GIVEN_IDX | TVQ_F_NO_UMMA)) || n % q_hw != 0 || n >= (int64_t(1) << 31) - 64))
        return TVQ_ERR_UNSUPPORTED;      // the channels-first q store exists in the resident-codebook kernel only
    if (x_hw > 0 && (!resident || (flags & (TVQ_F_EXACT | TVQ_F_GIVEN_IDX | TVQ_F_NO_UMMA)) || n % x_hw != 0 || n >= (int64_t(1) << 31) - 64 ||
                     (q_hw > 0 && q_hw != x_hw) || (q_hw == 0 && (flags & TVQ_F_WRITE_Q))))
        return TVQ_ERR_UNSUPPORTED;      // so does the channels-first x load (q, if written, in the same layout)
    const bool use_stream = !(flags & (TVQ_F_EXACT | TVQ_F_GIVEN_IDX | TVQ_F_NO_UMMA)) && !resident && n > 0 &&
                            n < (int64_t(1) << 31) - 256;
    void* cbh = use_stream ? reinterpret_cast<unsigned char*>(workspace) + ws_bf16_offset(k, d) : nullptr;
    void* e2h = use_stream ? reinterpret_cast<unsigned char*>(workspace) + ws_e2h_offset(k, d) : nullptr;
    // per-call preparation: |e|^2, zero statistics / loss (+ the bf16 operands of the streamed path)
    if ((rc = launch_prep(codebook, k, d, e2, hdr, stats, stats_len, cbh, stream_dp(d), e2h, *di, stream)) != TVQ_OK) return rc;
    if (n == 0) return TVQ_OK;   // nothing to assign; scalars are left to the caller (reference yields NaN)

    FwdParams p;
    p.x = x; p.cb = codebook; p.n = n; p.k = k; p.d = d;
    p.idx = idx; p.q = (flags & TVQ_F_WRITE_Q) ? q : nullptr; p.stats = stats; p.scalars = scalars;
    p.hdr = hdr; p.e2 = e2; p.commitment_weight = commitment_weight;
    p.commit_out = nullptr; p.weighted_out = nullptr; p.fuse_ema = 0;
    p.peers = nullptr; p.dp_rank = 0; p.dp_world = 1; p.dp_defer = 0; p.dp_timeout_ns = 0; p.q_hw = q_hw; p.x_hw = x_hw;
    p.cluster_size = nullptr; p.embed_avg = nullptr; p.embed = nullptr; p.embed_prev = nullptr;
    p.decay = p.one_minus_decay = p.eps = p.k_eps = 0.f;
    p.num_tiles = (int)((n + kBM - 1) / kBM);
    p.exact = (flags & TVQ_F_EXACT) ? 1 : 0;
    p.given_idx = (flags & TVQ_F_GIVEN_IDX) ? 1 : 0;
    // Resident-codebook tcgen05 path: train k <= 32 (per-warp TMEM accumulators), eval k <= 64; d <= 128
    // (the configs/config.yaml regime).
    if (!(flags & (TVQ_F_EXACT | TVQ_F_GIVEN_IDX | TVQ_F_NO_UMMA)) && k <= (train ? 32 : 64) && d <= 128 &&
        n < (int64_t(1) << 31) - 64) {
        const int udp = d <= 64 ? 64 : 128;
        const int ukp = k <= 16 ? 16 : k <= 32 ? 32 : 64;
        p.use_hist = 1;
        p.stats_mode = train ? kStatsSmall : kStatsNone;
        return train ? dispatch_fwd_umma<true>(udp, ukp, p, *di, stream) : dispatch_fwd_umma<false>(udp, ukp, p, *di, stream);
    }
    if (use_stream) {
        p.use_hist = 0;
        p.stats_mode = train ? kStatsLarge : kStatsNone;
        return train ? dispatch_fwd_stream<true>(p, cbh, e2h, *di, stream) : dispatch_fwd_stream<false>(p, cbh, e2h, *di, stream);
    }
    const int dp = pad_dim(d);
    p.use_hist = k <= 2048;
    p.stats_mode = kStatsNone;
    if (train) p.stats_mode = ((int64_t)k * dp <= 8192 && k <= 512) ? kStatsSmall : kStatsLarge;
    SmemPlan pl = make_smem_plan(dp, k, p.stats_mode, p.use_hist, kBN);
    if (pl.total > di->max_smem_optin) return TVQ_ERR_UNSUPPORTED;
    return train ? dispatch_fwd_simt<true>(dp, p, pl, *di, stream) : dispatch_fwd_simt<false>(dp, p, pl, *di, stream);
}

int train_step_impl(const float* x, float* embed, float* cluster_size, float* embed_avg, float* embed_prev, int64_t n, int k,
                    int d, float commitment_weight, double decay, double eps, int64_t* idx, float* q, float* scalars,
                    float* commit_out, float* weighted_out, void* workspace, size_t workspace_bytes, void* stream_,
                    void* const* peers, int dp_rank, int dp_world, int q_hw = 0, int x_hw = 0) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || d > 256 || (d & 3) || n < 0 || (int64_t)k * d >= (int64_t(1) << 31)) return TVQ_ERR_UNSUPPORTED;
    if (!embed || !cluster_size || !embed_avg || !scalars || !workspace || (n > 0 && (!x || !idx || !q))) return TVQ_ERR_BAD_ARG;
    if ((x && !aligned16(x)) || !aligned16(embed) || !aligned16(embed_avg) || (q && !aligned16(q)) || !aligned16(workspace) ||
        (embed_prev && !aligned16(embed_prev)))
        return TVQ_ERR_BAD_ARG;
    if (workspace_bytes < tvq_workspace_bytes(n, k, d)) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    WsHeader* hdr = reinterpret_cast<WsHeader*>(workspace);
    float* e2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + sizeof(WsHeader));
    float* scratch = e2 + e2_len(k);                          // private statistics: zero on entry, zero on exit
    const bool umma = n > 0 && k <= 32 && d <= 128 && n < (int64_t(1) << 31) - 64;
    if (dp_world > 1 && !umma) return TVQ_ERR_UNSUPPORTED;   // the fused data-parallel step exists for the resident-codebook kernel only
    if (q_hw > 0 && (!umma || n % q_hw != 0)) return TVQ_ERR_UNSUPPORTED;   // so does the channels-first q store
    if (x_hw > 0 && (!umma || x_hw != q_hw)) return TVQ_ERR_UNSUPPORTED;    // and the channels-first x load
    const bool use_stream = !umma && n > 0 && n < (int64_t(1) << 31) - 256;
    void* cbh = use_stream ? reinterpret_cast<unsigned char*>(workspace) + ws_bf16_offset(k, d) : nullptr;
    void* e2h = use_stream ? reinterpret_cast<unsigned char*>(workspace) + ws_e2h_offset(k, d) : nullptr;
    if (!umma) {
        // generic composition: zero + |e|^2 (+ bf16 copy), streamed tcgen05 forward, EMA kernel (three launches)
        if ((rc = launch_prep(embed, k, d, e2, hdr, scratch, TVQ_STATS_LEN(k, d), cbh, stream_dp(d), e2h, *di, stream)) != TVQ_OK) return rc;
    }
    FwdParams p;
    p.x = x; p.cb = embed; p.n = n; p.k = k; p.d = d;
    p.idx = idx; p.q = q; p.stats = scratch; p.scalars = scalars;
    p.hdr = hdr; p.e2 = e2; p.commitment_weight = commitment_weight;
    p.commit_out = commit_out; p.weighted_out = weighted_out;
    p.cluster_size = cluster_size; p.embed_avg = embed_avg; p.embed = embed; p.embed_prev = embed_prev;
    p.decay = (float)decay; p.one_minus_decay = (float)(1.0 - decay); p.eps = (float)eps; p.k_eps = (float)((double)k * eps);
    p.peers = peers; p.dp_rank = dp_rank; p.dp_world = dp_world; p.dp_timeout_ns = g_peer_timeout_ns; p.q_hw = q_hw; p.x_hw = x_hw;
    p.dp_defer = dp_world > 1 ? t_hint_defer_exchange : 0;
    t_hint_defer_exchange = 0;
    p.num_tiles = (int)((n + kBM - 1) / kBM);
    p.exact = 0; p.given_idx = 0; p.use_hist = k <= 2048;
    if (umma) {
        p.fuse_ema = 1;
        p.use_hist = 1;
        p.stats_mode = kStatsSmall;
        return dispatch_fwd_umma<true>(d <= 64 ? 64 : 128, k <= 16 ? 16 : 32, p, *di, stream);
    }
    p.fuse_ema = 0;
    if (use_stream) {
        p.use_hist = 0;
        p.stats_mode = kStatsLarge;
        if ((rc = dispatch_fwd_stream<true>(p, cbh, e2h, *di, stream)) != TVQ_OK) return rc;
    } else if (n > 0) {
        const int dp = pad_dim(d);
        p.stats_mode = ((int64_t)k * dp <= 8192 && k <= 512) ? kStatsSmall : kStatsLarge;
        SmemPlan pl = make_smem_plan(dp, k, p.stats_mode, p.use_hist, kBN);
        if (pl.total > di->max_smem_optin) return TVQ_ERR_UNSUPPORTED;
        if ((rc = dispatch_fwd_simt<true>(dp, p, pl, *di, stream)) != TVQ_OK) return rc;
    }
    return tvq_ema_update(scratch, cluster_size, embed_avg, embed, embed_prev, k, d, decay, eps, workspace, workspace_bytes, stream_);
}
}  // namespace

extern "C" {

int tvq_forward(const float* x, const float* codebook, int64_t n, int k, int d, unsigned flags,
                float commitment_weight, int64_t* idx, float* q, float* stats, float* scalars, void* workspace,
                size_t workspace_bytes, void* stream_) {
    TVQ_RANGE("tvq_forward");
    return forward_impl(x, codebook, n, k, d, flags, commitment_weight, idx, q, stats, scalars, workspace, workspace_bytes, stream_, 0);
}

int tvq_forward_qcf(const float* x, const float* codebook, int64_t n, int k, int d, unsigned flags,
                    float commitment_weight, int64_t* idx, float* q, float* stats, float* scalars, void* workspace,
                    size_t workspace_bytes, int q_hw, void* stream_) {
    TVQ_RANGE("tvq_forward_qcf");
    if (q_hw < 1 || !q || !(flags & TVQ_F_WRITE_Q)) return TVQ_ERR_BAD_ARG;
    return forward_impl(x, codebook, n, k, d, flags, commitment_weight, idx, q, stats, scalars, workspace, workspace_bytes, stream_, q_hw);
}

int tvq_train_step_qcf(const float* x, float* embed, float* cluster_size, float* embed_avg, float* embed_prev, int64_t n, int k,
                       int d, float commitment_weight, double decay, double eps, int64_t* idx, float* q, float* scalars,
                       float* commit_out, float* weighted_out, void* workspace, size_t workspace_bytes, void* const* peer_bufs,
                       int rank, int world, int q_hw, void* stream_) {
    TVQ_RANGE("tvq_train_step_qcf");
    if (q_hw < 1 || n < 1 || world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && !peer_bufs)) return TVQ_ERR_BAD_ARG;
    return train_step_impl(x, embed, cluster_size, embed_avg, embed_prev, n, k, d, commitment_weight, decay, eps, idx, q, scalars,
                           commit_out, weighted_out, workspace, workspace_bytes, stream_, world > 1 ? peer_bufs : nullptr, rank, world,
                           q_hw);
}

int tvq_forward_cf(const float* z, const float* codebook, int64_t b, int hw, int k, int d, unsigned flags, float commitment_weight,
                   int64_t* idx, float* q, float* stats, float* scalars, void* workspace, size_t workspace_bytes, void* stream_) {
    TVQ_RANGE("tvq_forward_cf");
    if (hw < 1 || b < 1 || b * (int64_t)hw >= (int64_t(1) << 31) - 64) return TVQ_ERR_BAD_ARG;
    if (((flags & TVQ_F_WRITE_Q) != 0) != (q != nullptr)) return TVQ_ERR_BAD_ARG;
    return forward_impl(z, codebook, b * hw, k, d, flags, commitment_weight, idx, q, stats, scalars, workspace, workspace_bytes, stream_,
                        q ? hw : 0, hw);
}

int tvq_train_step_cf(const float* z, float* embed, float* cluster_size, float* embed_avg, float* embed_prev, int64_t b, int hw, int k,
                      int d, float commitment_weight, double decay, double eps, int64_t* idx, float* q, float* scalars,
                      float* commit_out, float* weighted_out, void* workspace, size_t workspace_bytes, void* const* peer_bufs,
                      int rank, int world, void* stream_) {
    TVQ_RANGE("tvq_train_step_cf");
    if (hw < 1 || b < 1 || b * (int64_t)hw >= (int64_t(1) << 31) - 64 || world < 1 || world > 64 || rank < 0 || rank >= world ||
        (world > 1 && !peer_bufs))
        return TVQ_ERR_BAD_ARG;
    return train_step_impl(z, embed, cluster_size, embed_avg, embed_prev, b * hw, k, d, commitment_weight, decay, eps, idx, q, scalars,
                           commit_out, weighted_out, workspace, workspace_bytes, stream_, world > 1 ? peer_bufs : nullptr, rank, world,
                           hw, hw);
}

int tvq_backward_cfx(const float* g_zq, const float* g_commit, const float* g_weighted, const float* z, const int64_t* idx,
                     const float* codebook, int64_t b, int hw, int k, int d, float commitment_weight, float* g_z, void* stream_) {
    TVQ_RANGE("tvq_backward_cfx");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d < 1 || hw < 1 || b < 0 || k < 1) return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    if (!z || !idx || !codebook || !g_z) return TVQ_ERR_BAD_ARG;
    const size_t smem = ((size_t)k * (d + 1) + (size_t)hw) * sizeof(float);
    const bool vec = ((int64_t)hw * d) % 4 == 0 && aligned16(z) && aligned16(g_z) && (!g_zq || aligned16(g_zq));
    if (smem > 100 * 1024) return TVQ_ERR_UNSUPPORTED;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) {      // idempotent and cheap next to the launch; no cache, so nothing to key by device
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        const float scale = (float)(2.0 / ((double)b * (double)hw * (double)d));
        // every CTA copies the codebook into shared memory once, so a CTA should serve several batch elements — the
        // same number each (b = 1024: 512 CTAs x 2)
        // (measured at 1024 x 75: 4 CTAs per SM 21.1 us, 8 per SM 23.9 us, 2 per SM 30.4 us)
        const int64_t cap = 4LL * di->sm_count;
        const int64_t per = (b + cap - 1) / cap;
        const int64_t grid = (b + per - 1) / per;
        kern<<<(unsigned)grid, 256, smem, stream>>>(g_zq, g_commit, g_weighted, z, idx, codebook, b, hw, k, d, commitment_weight, scale, g_z);
        return TVQ_OK;
    };
    if (k <= 256) rc = vec ? launch(backward_cfx_kernel<true, unsigned char>) : launch(backward_cfx_kernel<false, unsigned char>);
    else rc = vec ? launch(backward_cfx_kernel<true, int>) : launch(backward_cfx_kernel<false, int>);
    if (rc != TVQ_OK) return rc;
    return launch_status();
}

int tvq_backward_cf(const float* g_zq, const float* g_commit, const float* g_weighted, const float* x, const int64_t* idx,
                    const float* codebook, int64_t b, int hw, int k, int d, float commitment_weight, float* g_z, void* stream_) {
    TVQ_RANGE("tvq_backward_cf");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (d < 1 || hw < 1 || b < 0) return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    if (!x || !idx || !codebook || !g_z) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    const float scale = (float)(2.0 / ((double)b * (double)hw * (double)d));
    const size_t slab = ((size_t)hw * (d + 1) + (size_t)k * (d + 1) + (size_t)hw) * sizeof(float);
    if (k >= 1 && (d & 3) == 0 && aligned16(x) && aligned16(codebook) && slab <= 100 * 1024) {
        static PerDeviceInt configured(48 * 1024);
        if ((rc = ensure_dynamic_smem(backward_cf_slab_kernel, configured, di->index, slab)) != TVQ_OK) return rc;
        int64_t grid = b < 8LL * di->sm_count ? b : 8LL * di->sm_count;
        backward_cf_slab_kernel<<<(unsigned)grid, 256, slab, stream>>>(g_zq, g_commit, g_weighted, x, idx, codebook, b, hw, k, d,
                                                                      commitment_weight, scale, g_z);
        return launch_status();
    }
    int64_t tiles = b * ((hw + 31) / 32) * ((d + 31) / 32);
    if (tiles > 32LL * di->sm_count) tiles = 32LL * di->sm_count;
    backward_cf_kernel<<<(unsigned)tiles, 256, 0, stream>>>(g_zq, g_commit, g_weighted, x, idx, codebook, b, hw, d, commitment_weight,
                                                             scale, g_z);
    return launch_status();
}

int tvq_train_step(const float* x, float* embed, float* cluster_size, float* embed_avg, float* embed_prev, int64_t n, int k,
                   int d, float commitment_weight, double decay, double eps, int64_t* idx, float* q, float* scalars,
                   float* commit_out, float* weighted_out, void* workspace, size_t workspace_bytes, void* stream_) {
    TVQ_RANGE("tvq_train_step");
    return train_step_impl(x, embed, cluster_size, embed_avg, embed_prev, n, k, d, commitment_weight, decay, eps, idx, q, scalars,
                           commit_out, weighted_out, workspace, workspace_bytes, stream_, nullptr, 0, 1);
}

int tvq_train_step_dp(const float* x, float* embed, float* cluster_size, float* embed_avg, float* embed_prev, int64_t n, int k,
                      int d, float commitment_weight, double decay, double eps, int64_t* idx, float* q, float* scalars,
                      float* commit_out, float* weighted_out, void* workspace, size_t workspace_bytes, void* const* peer_bufs,
                      int rank, int world, void* stream_) {
    TVQ_RANGE("tvq_train_step_dp");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || (world > 1 && !peer_bufs)) return TVQ_ERR_BAD_ARG;
    if (n < 1) return TVQ_ERR_UNSUPPORTED;      // every rank must launch (the exchange is collective)
    return train_step_impl(x, embed, cluster_size, embed_avg, embed_prev, n, k, d, commitment_weight, decay, eps, idx, q, scalars,
                           commit_out, weighted_out, workspace, workspace_bytes, stream_, peer_bufs, rank, world);
}

int tvq_ema_update(const float* stats, float* cluster_size, float* embed_avg, float* embed, float* embed_prev,
                   int k, int d, double decay, double eps, void* workspace, size_t workspace_bytes, void* stream_) {
    TVQ_RANGE("tvq_ema_update");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || (d & 3)) return TVQ_ERR_UNSUPPORTED;
    if (!stats || !cluster_size || !embed_avg || !embed || !workspace || workspace_bytes < sizeof(WsHeader)) return TVQ_ERR_BAD_ARG;
    if (!aligned16(stats) || !aligned16(embed_avg) || !aligned16(embed) || (embed_prev && !aligned16(embed_prev))) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    EmaParams p;
    p.stats = stats; p.cluster_size = cluster_size; p.embed_avg = embed_avg; p.embed = embed; p.embed_prev = embed_prev;
    p.k = k; p.d = d;
    p.decay = (float)decay;                 // python double -> tensor dtype, as mul_(decay) does
    p.one_minus_decay = (float)(1.0 - decay);
    p.eps = (float)eps;
    p.k_eps = (float)((double)k * eps);
    p.hdr = reinterpret_cast<WsHeader*>(workspace);
    int64_t work = (int64_t)k * (d / 4);
    int blocks = (int)((work + 255) / 256);
    if (blocks > di->sm_count) blocks = di->sm_count;
    if (blocks < 1) blocks = 1;
    ema_kernel<<<blocks, 256, 0, stream>>>(p);
    return launch_status();
}

size_t tvq_exchange_bytes(int k, int d, int world) {
    if (k < 1 || d < 4 || world < 1) return 0;
    const size_t len4 = ((size_t)TVQ_STATS_LEN(k, d) + 3) / 4;
    return 64 + (((size_t)2 * world * 4 + 63) & ~(size_t)63) + (size_t)2 * world * len4 * 16;
}

namespace {
int ema_dp_launch(const float* stats, void* const* peer_bufs, int rank, int world, float* cluster_size, float* embed_avg,
                  float* embed, float* embed_prev, int k, int d, double decay, double eps, int finalize, float* consume,
                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || (d & 3) || world < 1 || world > 64 || rank < 0 || rank >= world) return TVQ_ERR_UNSUPPORTED;
    if (TVQ_STATS_LEN(k, d) > (int64_t(1) << 16)) return TVQ_ERR_UNSUPPORTED;   // one CTA: small statistics only
    if ((!finalize && !stats) || !peer_bufs || !cluster_size || !embed_avg || !embed) return TVQ_ERR_BAD_ARG;
    if ((stats && !aligned16(stats)) || !aligned16(embed_avg) || !aligned16(embed) || (embed_prev && !aligned16(embed_prev))) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    EmaDpParams p;
    p.e.stats = stats; p.e.cluster_size = cluster_size; p.e.embed_avg = embed_avg; p.e.embed = embed; p.e.embed_prev = embed_prev;
    p.e.k = k; p.e.d = d;
    p.e.decay = (float)decay;
    p.e.one_minus_decay = (float)(1.0 - decay);
    p.e.eps = (float)eps;
    p.e.k_eps = (float)((double)k * eps);
    p.e.hdr = nullptr;
    p.peers = peer_bufs; p.rank = rank; p.world = world;
    p.len4 = (TVQ_STATS_LEN(k, d) + 3) / 4;
    p.timeout_ns = g_peer_timeout_ns;
    p.finalize = finalize;
    p.consume = consume;
    ema_dp_kernel<<<1, 1024, 0, stream>>>(p);
    return launch_status();
}
}  // namespace

int tvq_ema_update_dp(const float* stats, void* const* peer_bufs, int rank, int world, float* cluster_size, float* embed_avg,
                      float* embed, float* embed_prev, int k, int d, double decay, double eps, void* stream_) {
    TVQ_RANGE("tvq_ema_update_dp");
    return ema_dp_launch(stats, peer_bufs, rank, world, cluster_size, embed_avg, embed, embed_prev, k, d, decay, eps, 0, nullptr, stream_);
}

int tvq_ema_finalize_dp(int published, void* workspace, size_t workspace_bytes, void* const* peer_bufs, int rank, int world,
                        float* cluster_size, float* embed_avg, float* embed, int k, int d, double decay, double eps, void* stream_) {
    TVQ_RANGE("tvq_ema_finalize_dp");
    if (published)
        return ema_dp_launch(nullptr, peer_bufs, rank, world, cluster_size, embed_avg, embed, nullptr, k, d, decay, eps, 1, nullptr, stream_);
    // the step left its statistics in the private scratch of the workspace (zero on entry, zero again after this call)
    if (!workspace || !aligned16(workspace) || k < 1 || d < 4 || workspace_bytes < tvq_workspace_bytes(0, k, d)) return TVQ_ERR_BAD_ARG;
    float* scratch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + sizeof(WsHeader)) + e2_len(k);
    return ema_dp_launch(scratch, peer_bufs, rank, world, cluster_size, embed_avg, embed, nullptr, k, d, decay, eps, 0, scratch, stream_);
}

int tvq_backward(const float* g_q, const float* g_commit, const float* g_weighted, const float* x, const int64_t* idx,
                 const float* codebook, int64_t n, int k, int d, float commitment_weight, float* g_x, void* stream_) {
    TVQ_RANGE("tvq_backward");
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)k;
    if (d < 4 || (d & 3) || n < 0) return TVQ_ERR_UNSUPPORTED;
    if (n == 0) return TVQ_OK;
    if (!x || !idx || !codebook || !g_x) return TVQ_ERR_BAD_ARG;
    if ((g_q && !aligned16(g_q)) || !aligned16(x) || !aligned16(codebook) || !aligned16(g_x)) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    const float scale = (float)(2.0 / ((double)n * (double)d));
    int64_t work = n * (d / 4);
    int64_t blocks = (work + 255) / 256;
    if (blocks > 8LL * di->sm_count) blocks = 8LL * di->sm_count;
    backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g_q, g_commit, g_weighted, x, idx, codebook, n, d, commitment_weight, scale, g_x);
    return launch_status();
}

int tvq_gather(const int64_t* tokens, const float* codebook, int64_t b, int64_t t, int k, int d, int layout,
               float* out, void* stream_) {
    return tvq_gather_checked(tokens, codebook, b, t, k, d, layout, out, nullptr, stream_);
}

int tvq_gather_checked(const int64_t* tokens, const float* codebook, int64_t b, int64_t t, int k, int d, int layout,
                       float* out, unsigned* bad_count, void* stream_) {
    TVQ_RANGE("tvq_gather_checked");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 1 || b < 0 || t < 0 || (layout != 0 && layout != 1)) return TVQ_ERR_UNSUPPORTED;
    if (b * t == 0) return TVQ_OK;
    if (!tokens || !codebook || !out) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    if (layout == 0) {
        if ((d & 3) || !aligned16(codebook) || !aligned16(out)) return TVQ_ERR_UNSUPPORTED;
        int64_t blocks = (b * t + 7) / 8;
        if (blocks > 16LL * di->sm_count) blocks = 16LL * di->sm_count;
        gather_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tokens, codebook, b * t, k, d, out, bad_count);
    } else {
        int64_t tiles = b * ((t + 31) / 32) * ((d + 31) / 32);
        if (tiles > 16LL * di->sm_count) tiles = 16LL * di->sm_count;
        gather_transposed_kernel<<<(unsigned)tiles, 256, 0, stream>>>(tokens, codebook, b, t, k, d, out, bad_count);
    }
    return launch_status();
}

int tvq_neg_dist(const float* x, const float* codebook, int64_t n, int k, int d, float* dist, void* stream_) {
    TVQ_RANGE("tvq_neg_dist");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || (d & 3) || n < 0) return TVQ_ERR_UNSUPPORTED;
    if (n == 0) return TVQ_OK;
    if (!x || !codebook || !dist || !aligned16(x) || !aligned16(codebook)) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 16LL * di->sm_count) blocks = 16LL * di->sm_count;
    neg_dist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, codebook, n, k, d, dist);
    return launch_status();
}

int tvq_frontend(const float* x, int64_t b, int c, int l, int n_fft, float* xf, float* enc_in_l, float* enc_in_h, float* x_l,
                 float* x_h, void* stream_) {
    TVQ_RANGE("tvq_frontend");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b < 0 || c < 1 || n_fft < 4 || n_fft > 64 || (n_fft & 3) || l <= n_fft / 2 || l / (n_fft / 4) < 1) return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    if (!x) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    const size_t smem = frontend_smem_bytes(l, n_fft);
    if (smem > (size_t)di->max_smem_optin) return TVQ_ERR_UNSUPPORTED;
    static PerDeviceInt configured(48 * 1024);
    if ((rc = ensure_dynamic_smem(frontend_kernel, configured, di->index, smem)) != TVQ_OK) return rc;
    FrontendParams p;
    p.x = x; p.rows = b * c; p.c = c; p.l = l; p.n_fft = n_fft;
    p.xf = xf; p.enc_in_l = enc_in_l; p.enc_in_h = enc_in_h; p.x_l = x_l; p.x_h = x_h;
    int64_t grid = p.rows;                                   // one CTA per row: CTAs retire and start independently
    if (grid > 64LL * di->sm_count) grid = 64LL * di->sm_count;
    frontend_kernel<<<(unsigned)grid, 128, smem, stream>>>(p);
    return launch_status();
}

}  // extern "C"

namespace {
template <bool BACKWARD>
int launch_band_istft(const BandIstftParams& p, int64_t b, int c, cudaStream_t stream) {
    if (b < 0 || c < 1 || p.n_fft < 4 || p.n_fft > 64 || (p.n_fft & 3) || p.l < 1 || p.t < 2 || p.band < 0 || p.band > 2)
        return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    const size_t smem = band_istft_smem_bytes(p.t, p.n_fft);
    if (smem > (size_t)di->max_smem_optin) return TVQ_ERR_UNSUPPORTED;
    static PerDeviceInt configured(48 * 1024);
    if ((rc = ensure_dynamic_smem(band_istft_kernel<BACKWARD>, configured, di->index, smem)) != TVQ_OK) return rc;
    int64_t grid = p.rows;
    if (grid > 64LL * di->sm_count) grid = 64LL * di->sm_count;
    band_istft_kernel<BACKWARD><<<(unsigned)grid, 128, smem, stream>>>(p);
    return launch_status();
}
}  // namespace

extern "C" {

int tvq_band_istft_frames(const float* u, int64_t b, int c, int t, int l, int n_fft, int band, float* y, void* stream_) {
    TVQ_RANGE("tvq_band_istft_frames");
    if (b > 0 && (!u || !y)) return TVQ_ERR_BAD_ARG;
    BandIstftParams p;
    p.u = u; p.g_y = nullptr; p.y = y; p.g_u = nullptr; p.rows = b * c; p.l = l; p.n_fft = n_fft; p.band = band; p.t = t;
    return launch_band_istft<false>(p, b, c, (cudaStream_t)stream_);
}

int tvq_band_istft_frames_backward(const float* g_y, int64_t b, int c, int t, int l, int n_fft, int band, float* g_u, void* stream_) {
    TVQ_RANGE("tvq_band_istft_frames_backward");
    if (b > 0 && (!g_y || !g_u)) return TVQ_ERR_BAD_ARG;
    BandIstftParams p;
    p.u = nullptr; p.g_y = g_y; p.y = nullptr; p.g_u = g_u; p.rows = b * c; p.l = l; p.n_fft = n_fft; p.band = band; p.t = t;
    return launch_band_istft<true>(p, b, c, (cudaStream_t)stream_);
}

int tvq_band_istft(const float* u, int64_t b, int c, int l, int n_fft, int band, float* y, void* stream_) {
    if (n_fft < 4 || l <= n_fft / 2) return TVQ_ERR_UNSUPPORTED;
    return tvq_band_istft_frames(u, b, c, l / (n_fft / 4) + 1, l, n_fft, band, y, stream_);
}

int tvq_band_istft_backward(const float* g_y, int64_t b, int c, int l, int n_fft, int band, float* g_u, void* stream_) {
    if (n_fft < 4 || l <= n_fft / 2) return TVQ_ERR_UNSUPPORTED;
    return tvq_band_istft_frames_backward(g_y, b, c, l / (n_fft / 4) + 1, l, n_fft, band, g_u, stream_);
}

int tvq_maskgit_step(const float* logits, const int64_t* s, const float* q, const float* u, int64_t b, int n, int k,
                     int64_t mask_token_id, int mask_len, float temperature, int64_t* s_new, int64_t* sampled, uint8_t* masking,
                     void* stream_) {
    TVQ_RANGE("tvq_maskgit_step");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b < 0 || n < 1 || k < 1 || n > 8192 || mask_len < 0) return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    if (!logits || !s || !q || !u || !s_new) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    MaskgitParams p;
    p.logits = logits; p.s = s; p.q = q; p.u = u; p.b = b; p.n = n; p.k = k; p.mask_token_id = mask_token_id;
    p.mask_len = mask_len; p.temperature = temperature; p.s_new = s_new; p.sampled = sampled; p.masking = masking;
    int64_t grid = b;
    if (grid > 32LL * di->sm_count) grid = 32LL * di->sm_count;
    maskgit_step_kernel<<<(unsigned)grid, 128, (size_t)n * 8, stream>>>(p);
    return launch_status();
}

int tvq_transpose(const float* in, int64_t b, int r, int s, float* out, void* stream_) {
    TVQ_RANGE("tvq_transpose");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b < 0 || r < 1 || s < 1) return TVQ_ERR_UNSUPPORTED;
    if (b == 0) return TVQ_OK;
    if (!in || !out) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    const size_t slab = (size_t)r * (size_t)(s | 1) * sizeof(float);
    // one batch element per CTA pass where the 32 x 32 tiling would leave most of a tile empty (LF: 18 positions); measured
    // at 1024 x 128 x 18: 6.4 vs 8.0 us, but at x 75 the tiled kernel wins (20 vs 27 us: fewer, wider instructions)
    if (slab <= 64 * 1024 && (int64_t)r * s >= 256 && (r < 32 || s < 32)) {
        if (slab > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(slab_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab);
            if (e != cudaSuccess) return (int)e;
        }
        int64_t grid = b < 16LL * di->sm_count ? b : 16LL * di->sm_count;
        slab_transpose_kernel<<<(unsigned)grid, 256, slab, stream>>>(in, out, b, r, s);
        return launch_status();
    }
    int64_t tiles = b * ((r + 31) / 32) * ((s + 31) / 32);
    if (tiles > 32LL * di->sm_count) tiles = 32LL * di->sm_count;
    batched_transpose_kernel<<<(unsigned)tiles, 256, 0, stream>>>(in, out, b, r, s);
    return launch_status();
}

int tvq_snake_forward(const float* x, const float* a, int64_t n, int c, int64_t s, int channels_last, float* y, void* stream_) {
    TVQ_RANGE("tvq_snake_forward");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 1 || s < 1 || c > kSnakeMaxC || s >= (int64_t(1) << 31)) return TVQ_ERR_UNSUPPORTED;
    const int64_t total = n * c * s;
    if (total == 0) return TVQ_OK;
    if (!x || !a || !y) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    int64_t blocks = (total + 1023) / 1024;
    if (blocks > 16LL * di->sm_count) blocks = 16LL * di->sm_count;
    if (channels_last) snake_fwd_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(x, a, total, c, (int)s, y);
    else snake_fwd_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(x, a, total, c, (int)s, y);
    return launch_status();
}

int tvq_snake_backward(const float* g, const float* x, const float* a, int64_t n, int c, int64_t s, int channels_last, float* g_x,
                       float* g_a, void* stream_) {
    TVQ_RANGE("tvq_snake_backward");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || c < 1 || s < 1 || c > kSnakeMaxC || s >= (int64_t(1) << 31)) return TVQ_ERR_UNSUPPORTED;
    const int64_t total = n * c * s;
    if (total == 0) return TVQ_OK;
    if (!g || !x || !a || !g_x || !g_a) return TVQ_ERR_BAD_ARG;
    DeviceInfo* di = nullptr;
    int rc = device_info(&di);
    if (rc != TVQ_OK) return rc;
    int64_t blocks = (total + 2047) / 2048;
    if (blocks > 8LL * di->sm_count) blocks = 8LL * di->sm_count;
    const size_t smem = (size_t)c * sizeof(float);
    if (channels_last) snake_bwd_kernel<true><<<(unsigned)blocks, 256, smem, stream>>>(g, x, a, total, c, (int)s, g_x, g_a);
    else snake_bwd_kernel<false><<<(unsigned)blocks, 256, smem, stream>>>(g, x, a, total, c, (int)s, g_x, g_a);
    return launch_status();
}

int tvq_reseed(const float* x, const int64_t* rows, const float* cluster_size, float threshold, float* embed,
               int64_t n, int k, int d, void* stream_) {
    TVQ_RANGE("tvq_reseed");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || d < 4 || (d & 3) || n < 1) return TVQ_ERR_UNSUPPORTED;
    if (!x || !rows || !cluster_size || !embed || !aligned16(x) || !aligned16(embed)) return TVQ_ERR_BAD_ARG;
    int blocks = (k + 7) / 8;
    reseed_kernel<<<blocks, 256, 0, stream>>>(x, rows, cluster_size, threshold, embed, n, k, d);
    return launch_status();
}

}  // extern "C"
