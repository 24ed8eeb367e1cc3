// pmg_abi.cu — extern "C" entry points declared in include/pmg_b200.h.
// No torch types, no allocation, no global mutable state besides a thread-local error string.
#include <stdio.h>
#include <string.h>

#include "pmg_common.cuh"

namespace pmg {

static thread_local char g_cuda_err[256] = "";

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    (void)cudaGetLastError();   // clear the sticky launch error so the next call starts clean
}

int gram_batched_generic(const void* Fm, int f_dtype, int B, int C, int HW, float scale, float* G, cudaStream_t st);
int gram_batched_fast(const void* Fm, int f_dtype, int B, int C, int HW, float scale, float* G, cudaStream_t st);

// pmg_forward_profile only: keeps the GPU busy for ~0.3 ms so that the host-side work of the profiled forward
// (tensor-map encoding, launches) is enqueued ahead of the device and the per-stage events measure device time only
__global__ void profile_delay_kernel(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}

static inline bool misaligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

static int check_arch() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PMG_ERR_CUDA;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PMG_ERR_CUDA;
    return major == 10 ? PMG_OK : PMG_ERR_ARCH;
}

}  // namespace pmg

using namespace pmg;

extern "C" {

int pmg_abi_version(void) { return PMG_ABI_VERSION; }

const char* pmg_status_string(int s) {
    switch (s) {
        case PMG_OK: return "ok";
        case PMG_ERR_BAD_ARG: return "bad argument";
        case PMG_ERR_BAD_SHAPE: return "bad shape";
        case PMG_ERR_WORKSPACE: return "workspace / packed buffer too small";
        case PMG_ERR_ARCH: return "device is not sm_100 (B200)";
        case PMG_ERR_CUDA: return "CUDA error";
        case PMG_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

const char* pmg_last_cuda_error(void) { return g_cuda_err; }

int pmg_param_count(const pmg_config* cfg) {
    const Plan p = make_plan(cfg);
    if (!p.ok) return -1;
    return (p.norm == PMG_NORM_BATCH ? 5 : 3) * p.n_layers + (p.se ? 2 : 0) + 3 + 12 * p.nl + 3 * p.ntask;
}

size_t pmg_packed_bytes(const pmg_config* cfg) {
    const Plan p = make_plan(cfg);
    if (!p.ok || p.nl > 8) return 0;
    return make_layout(p).total;
}

int pmg_logits_per_image(const pmg_config* cfg) {
    const Plan p = make_plan(cfg);
    return p.ok ? p.ncls_total : -1;
}

int pmg_pack_weights(const pmg_config* cfg, const float* const* params, int n_params, void* packed,
                     size_t packed_bytes, pmg_stream_t stream) {
    const Plan p = make_plan(cfg);
    if (!p.ok || !params || !packed) return PMG_ERR_BAD_ARG;
    if (p.nl > 8) return PMG_ERR_UNSUPPORTED;
    if (n_params != pmg_param_count(cfg)) return PMG_ERR_BAD_ARG;
    for (int i = 0; i < n_params; ++i)
        if (!params[i]) return PMG_ERR_BAD_ARG;
    const PackedLayout L = make_layout(p);
    if (packed_bytes < L.total) return PMG_ERR_WORKSPACE;
    int rc = check_arch();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    PMG_CUDA_OK(cudaMemsetAsync(packed, 0, 256, st), "pack header");
    if ((rc = pack_generic(p, L, params, reinterpret_cast<char*>(packed), st))) return rc;
    return pack_fast(p, L, params, reinterpret_cast<char*>(packed), st);
}

static bool use_fast(const Plan& p, const Shape& s, unsigned flags) {
    return !(flags & PMG_FLAG_FORCE_GENERIC) && fast_supported(p, s);
}

size_t pmg_workspace_bytes(const pmg_config* cfg, int B, int H, int W) {
    const Plan p = make_plan(cfg);
    if (!p.ok || B < 1) return 0;
    const Shape s = make_shape(p, H, W);
    if (!s.ok) return 0;
    size_t g = generic_workspace_bytes(p, s, B);
    if (fast_supported(p, s)) {
        const size_t f = fast_workspace_bytes(p, s, B);
        g = f > g ? f : g;
    }
    return g;
}

int pmg_forward_launch_count(const pmg_config* cfg, int B, int H, int W, unsigned flags) {
    const Plan p = make_plan(cfg);
    if (!p.ok || B < 1) return -1;
    const Shape s = make_shape(p, H, W);
    if (!s.ok) return -1;
    return use_fast(p, s, flags) ? fast_launch_count(p, s, B, flags) : generic_launch_count(p);
}

static int forward_impl(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                        float* logits, float* gram, const pmg_taps* taps, void* workspace, size_t workspace_bytes,
                        unsigned flags, pmg_stream_t stream, StageTimer* tm, const LogitPeers* peers = nullptr) {
    const Plan p = make_plan(cfg);
    if (!p.ok || !packed || !x || !logits || !workspace) return PMG_ERR_BAD_ARG;
    if (x_dtype != PMG_F32 && x_dtype != PMG_BF16) return PMG_ERR_BAD_ARG;
    // vector loads / TMA tensor maps: the image batch, workspace and packed weights must be 16-byte aligned, outputs 4-byte
    if (misaligned(x, 16) || misaligned(workspace, 16) || misaligned(packed, 16) || misaligned(logits, 4) || misaligned(gram, 4))
        return PMG_ERR_BAD_ARG;
    if (B < 1) return PMG_ERR_BAD_SHAPE;
    const Shape s = make_shape(p, H, W);
    if (!s.ok) return PMG_ERR_BAD_SHAPE;
    int rc = check_arch();
    if (rc) return rc;
    const PackedLayout L = make_layout(p);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool fast = use_fast(p, s, flags);
    if (!fast && (flags & PMG_FLAG_FORCE_FAST)) return PMG_ERR_UNSUPPORTED;
    const size_t need = fast ? fast_workspace_bytes(p, s, B) : generic_workspace_bytes(p, s, B);
    if (workspace_bytes < need) return PMG_ERR_WORKSPACE;
    if (fast)
        return fast_forward(p, s, L, reinterpret_cast<const char*>(packed), x, x_dtype, B, logits, gram, taps,
                            reinterpret_cast<char*>(workspace), st, tm, flags, peers);
    rc = generic_forward(p, s, L, reinterpret_cast<const char*>(packed), x, x_dtype, B, logits, gram, taps,
                         reinterpret_cast<char*>(workspace), st, tm);
    return rc ? rc : copy_logits_to_peers(logits, (size_t)B * p.ncls_total * sizeof(float), peers, st);
}

int pmg_forward(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                float* logits, float* gram, const pmg_taps* taps, void* workspace, size_t workspace_bytes,
                unsigned flags, pmg_stream_t stream) {
    return forward_impl(cfg, packed, x, x_dtype, B, H, W, logits, gram, taps, workspace, workspace_bytes, flags,
                        stream, nullptr);
}

int pmg_forward_bcast(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                      float* logits, float* const* peer_logits, int n_peers, void* workspace, size_t workspace_bytes,
                      unsigned flags, pmg_stream_t stream) {
    if (n_peers < 0 || n_peers > PMG_MAX_PEERS || (n_peers > 0 && !peer_logits)) return PMG_ERR_BAD_ARG;
    LogitPeers peers;
    peers.n = n_peers;
    for (int d = 0; d < n_peers; ++d) {
        if (!peer_logits[d] || misaligned(peer_logits[d], 4)) return PMG_ERR_BAD_ARG;
        peers.dst[d] = peer_logits[d];
    }
    return forward_impl(cfg, packed, x, x_dtype, B, H, W, logits, nullptr, nullptr, workspace, workspace_bytes, flags, stream,
                        nullptr, &peers);
}

int pmg_forward_profile(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                        float* logits, void* workspace, size_t workspace_bytes, unsigned flags, pmg_stream_t stream,
                        pmg_stage_time* stages, int max_stages, int* n_stages) {
    if (!stages || !n_stages || max_stages < 1) return PMG_ERR_BAD_ARG;
    StageTimer tm;
    tm.st = reinterpret_cast<cudaStream_t>(stream);
    profile_delay_kernel<<<1, 1, 0, tm.st>>>(600000);
    int rc = forward_impl(cfg, packed, x, x_dtype, B, H, W, logits, nullptr, nullptr, workspace, workspace_bytes,
                          flags, stream, &tm);
    *n_stages = 0;
    if (rc == PMG_OK && cudaStreamSynchronize(tm.st) != cudaSuccess) rc = PMG_ERR_CUDA;
    for (int i = 1; i < tm.n && rc == PMG_OK; ++i) {
        if (*n_stages >= max_stages) break;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, tm.ev[i - 1], tm.ev[i]) != cudaSuccess) { rc = PMG_ERR_CUDA; break; }
        stages[*n_stages].name = tm.name[i];
        stages[*n_stages].ms = ms;
        ++*n_stages;
    }
    for (int i = 0; i < tm.n; ++i) cudaEventDestroy(tm.ev[i]);
    return rc;
}

int pmg_style_gram_shape(const pmg_config* cfg, int H, int W, int* n_patches, int* C) {
    const Plan p = make_plan(cfg);
    if (!p.ok || !n_patches || !C) return PMG_ERR_BAD_ARG;
    const Shape s = make_shape(p, H, W);
    if (!s.ok) return PMG_ERR_BAD_SHAPE;
    const int ph = s.Hc / p.pdiv, pw = s.Wc / p.pdiv;
    if (ph < 1 || pw < 1) return PMG_ERR_BAD_SHAPE;
    *n_patches = (s.Hc / ph) * (s.Wc / pw);
    *C = p.C;
    return PMG_OK;
}

int pmg_style_grams(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                    float* grams, void* workspace, size_t workspace_bytes, pmg_stream_t stream) {
    const Plan p = make_plan(cfg);
    if (!p.ok || !packed || !x || !grams || !workspace) return PMG_ERR_BAD_ARG;
    if (x_dtype != PMG_F32 && x_dtype != PMG_BF16) return PMG_ERR_BAD_ARG;
    if (misaligned(x, 16) || misaligned(workspace, 16) || misaligned(packed, 16) || misaligned(grams, 4)) return PMG_ERR_BAD_ARG;
    if (B < 1) return PMG_ERR_BAD_SHAPE;
    const Shape s = make_shape(p, H, W);
    if (!s.ok) return PMG_ERR_BAD_SHAPE;
    int rc = check_arch();
    if (rc) return rc;
    if (workspace_bytes < generic_workspace_bytes(p, s, B)) return PMG_ERR_WORKSPACE;
    return style_grams_forward(p, s, make_layout(p), reinterpret_cast<const char*>(packed), x, x_dtype, B, grams,
                               reinterpret_cast<char*>(workspace), reinterpret_cast<cudaStream_t>(stream));
}

int pmg_gram_batched(const void* F, int f_dtype, int B, int C, int HW, float scale, float* G, pmg_stream_t stream) {
    if (!F || !G) return PMG_ERR_BAD_ARG;
    if (f_dtype != PMG_F32 && f_dtype != PMG_BF16) return PMG_ERR_BAD_ARG;
    if (misaligned(F, f_dtype == PMG_F32 ? 4 : 2) || misaligned(G, 4)) return PMG_ERR_BAD_ARG;
    if (B < 1 || C < 1 || HW < 1 || B > 65535) return PMG_ERR_BAD_SHAPE;
    int rc = check_arch();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    rc = gram_batched_fast(F, f_dtype, B, C, HW, scale, G, st);
    if (rc != PMG_ERR_UNSUPPORTED) return rc;
    return gram_batched_generic(F, f_dtype, B, C, HW, scale, G, st);
}

}  // extern "C"
