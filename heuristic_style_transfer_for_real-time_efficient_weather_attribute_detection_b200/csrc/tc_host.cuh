// tc_host.cuh — host helpers for the tcgen05 GEMM: TMA tensor-map encoding (driver entry point fetched
// through the runtime, so the library does not link libcuda) and the templated launcher.
#pragma once

#include <atomic>
#include <cstring>

#include "pmg_common.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "tc_igemm.cuh"

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// A tensor map is a pure function of (base pointer, dims, strides, box, element strides, swizzle): memoise it.  One forward
// needs 8 maps and the driver's encode call costs several microseconds each — more than the launches themselves at B = 1.
// The cache is thread-local (no locking, re-entrant per thread) and direct-mapped; a miss just re-encodes.
struct TmapKey {
    uint64_t base, dims[4], strides[3];
    uint32_t box[4], estr[4], rank, swizzle;
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct alignas(64) TmapSlot { CUtensorMap map; TmapKey key; bool valid; };
constexpr int kTmapSlots = 256;

inline int encode_tmap_cached(CUtensorMap* m, uint32_t rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                              const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle swizzle, const char* what,
                              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    thread_local TmapSlot* slots = nullptr;
    if (!slots) slots = new TmapSlot[kTmapSlots]();      // lives as long as the thread (24 KB)
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.base = reinterpret_cast<uint64_t>(base); k.rank = rank; k.swizzle = (uint32_t)swizzle | ((uint32_t)dtype << 8);
    for (uint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; }
    for (uint32_t i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
    uint64_t h = 0xcbf29ce484222325ull;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 0x100000001b3ull; h ^= h >> 29; }
    TmapSlot& s = slots[h % kTmapSlots];
    if (s.valid && s.key == k) { *m = s.map; return PMG_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) { pmg::set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point"); return PMG_ERR_CUDA; }
    CUresult r = fn(m, dtype, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pmg::set_cuda_error(cudaErrorInvalidValue, what); return PMG_ERR_CUDA; }
    s.map = *m; s.key = k; s.valid = true;
    return PMG_OK;
}

// launch with programmatic dependent launch enabled (the kernel must call tc::pdl_wait() before touching
// data produced by earlier kernels in the stream)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, as clusters of 2 CTAs (CTA pairs for tcgen05 cta_group::2)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Kernel attributes (cudaFuncSetAttribute) and the SM count are PER DEVICE: a process that drives several GPUs through this
// library must set / query them once on each.  One bit per device ordinal, lock-free (a racing thread at worst repeats
// the idempotent driver call).
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}
struct PerDeviceFlag {
    std::atomic<unsigned long long> bits{0};
    bool test() const { return (bits.load(std::memory_order_acquire) >> current_device()) & 1ull; }
    void set() { bits.fetch_or(1ull << current_device(), std::memory_order_release); }
};

inline int num_sms() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        int v = 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(n = v, std::memory_order_relaxed);
    }
    return n;
}

// bf16 row-major [rows, cols] (cols contiguous, row pitch = pitch_elems), box = box_rows × 64 cols,
// 128-byte swizzle; out-of-bounds rows/cols read as zero.
inline int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                          uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return encode_tmap_cached(m, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "cuTensorMapEncodeTiled");
}

// 4-D view of a bf16 NCHW image batch (B,3,224,224) for the fused front end (tc_front.cuh):
//   dims (innermost first)  x = 224 | py = 14 | y16 = 16 | b·c = 3B          (y = 16·py + y16)
//   box                     224 × 8 × 4 × 1  = for 4 kernel rows, 8 whole image rows (448 contiguous bytes each; the
//                           tile uses 7, an 8th keeps every kernel-row slab a 256-byte multiple) = the 16 input columns
//                           of 98 conv1 pixels for one channel; NO swizzle, so TMA moves 448-byte
//                           runs (a swizzled 16-column inner box made it move 98 separate 32-byte rows per box,
//                           which bounded the whole kernel).  The UMMA descriptor still reads the slab with the
//                           32-byte swizzle; tc_front.cuh explains how the resulting half swap is absorbed.
inline int make_tmap_image_rows(CUtensorMap* m, const void* base, int B) {
    cuuint64_t dims[4] = {224, 14, 16, (cuuint64_t)3 * B};
    cuuint64_t strides[3] = {7168, 448, 100352};
    cuuint32_t box[4] = {224, 8, 4, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode_tmap_cached(m, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE, "cuTensorMapEncodeTiled(image rows)");
}

// 4-D view of an NHWC bf16 activation plane (B, H, W, C) for the implicit-GEMM convolution (tc_igemm.cuh): one box =
// 64 channels × Wo pixels (every `stride`-th) × bh rows (every `stride`-th) × nb images, 128-byte swizzle, zero OOB fill.
inline int make_tmap_act4d(CUtensorMap* m, const void* base, int B, int H, int W, int C, int stride, int Wo, int bh, int nb) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(stride * Wo), (cuuint32_t)(stride * bh), (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return encode_tmap_cached(m, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "cuTensorMapEncodeTiled(act 4d)");
}

template <int BN>
int launch_conv_igemm_t(const void* act_hi, const void* act_lo, const void* wgt, int B, int H, int W, const IgemmParams& p, cudaStream_t st) {
    CUtensorMap ta_hi, ta_lo, tb;
    int rc;
    if ((rc = make_tmap_act4d(&ta_hi, act_hi, B, H, W, p.C, p.stride, p.Wo, p.bh, p.nb))) return rc;
    if ((rc = make_tmap_act4d(&ta_lo, act_lo, B, H, W, p.C, p.stride, p.Wo, p.bh, p.nb))) return rc;
    if ((rc = make_tmap_bf16(&tb, wgt, p.N, (uint64_t)16 * p.C, (uint64_t)16 * p.C, BN))) return rc;
    constexpr int smem = ig_smem_bytes<BN>();
    static_assert(smem <= 232448, "implicit-GEMM kernel exceeds the 227 KB shared-memory limit");
    static PerDeviceFlag attr_set;
    if (!attr_set.test()) {
        PMG_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "igemm smem attr");
        attr_set.set();
    }
    const long long items = (long long)((p.B + p.nb - 1) / p.nb) * p.tiles_y * (p.N / BN);
    const int grid = (int)(items < num_sms() ? items : num_sms());
    conv_igemm_kernel<BN><<<grid, ig_threads<BN>(), smem, st>>>(ta_hi, ta_lo, tb, p);
    PMG_LAUNCH_OK("conv_igemm");
    return PMG_OK;
}

// conv k4 p1 stride s of NHWC hi/lo planes (B,H,W,C) with weights [N][16·C] ((ky,kx,c) order) → out fp32 [B·Ho·Wo, N] + bias
// narrow = keep the 128-wide N tile even where 256 divides N (A/B measurements)
inline int launch_conv_igemm(const void* act_hi, const void* act_lo, const void* wgt, const float* bias, int B, int H, int W, int C,
                             int Ho, int Wo, int stride, int N, float* out, cudaStream_t st, bool narrow = false,
                             float2* stats = nullptr, int* stats_tiles_y = nullptr, int* stats_rows_full = nullptr) {
    if (C % 64 || N % 128 || N > IG_MAX_N || Wo < 1 || Wo > 128 || stride * Wo > 256) return PMG_ERR_UNSUPPORTED;
    // M tile = nb images × bh output rows × Wo ≤ 128 rows: several rows of one image, or one row of several images,
    // whichever leaves fewer of the 128 MMA rows empty over the whole layer
    const int per = 128 / Wo;
    int bh = per < Ho ? per : Ho, nb = 1;
    const double eff_rows = (double)Ho * Wo / (((Ho + bh - 1) / bh) * 128.0);
    const double eff_imgs = (double)B * Wo / (((B + per - 1) / per) * 128.0);
    if (eff_imgs > eff_rows + 1e-9) { bh = 1; nb = per < B ? per : B; }
    if (stride * bh > 256 || nb > 256) return PMG_ERR_UNSUPPORTED;
    IgemmParams p{};
    p.B = B; p.Ho = Ho; p.Wo = Wo; p.bh = bh; p.nb = nb; p.tiles_y = (Ho + bh - 1) / bh; p.C = C; p.N = N; p.stride = stride;
    p.out = out; p.bias = bias;
    // fused InstanceNorm statistics need every M tile inside one image; the caller learns the tile geometry (or that it
    // has to take its statistics from the output itself: *stats_tiles_y = 0)
    p.stats = nb == 1 ? stats : nullptr;
    if (stats_tiles_y) *stats_tiles_y = p.stats ? p.tiles_y : 0;
    if (stats_rows_full) *stats_rows_full = bh * Wo;
    return (N % 256 == 0 && !narrow) ? launch_conv_igemm_t<256>(act_hi, act_lo, wgt, B, H, W, p, st)
                                     : launch_conv_igemm_t<128>(act_hi, act_lo, wgt, B, H, W, p, st);
}

// A: [M, K] bf16 planes (lo may be null ⇒ single plane), pitch lda; B: [N, K] bf16, pitch ldb.
// k_cols = columns that really exist in memory (≤ K); TMA zero-fills [k_cols, K) — used for conv0's K = 48.
// SEG2: a second K segment (A2 [M, K2] one plane, pitch lda; B2 [N, K2], pitch K2) accumulated into the same tiles.
template <int BN, int NA, int EPI, int ACT, bool SEG2 = false>
int launch_gemm_t(const void* A_hi, const void* A_lo, int lda, const void* B, int ldb, GemmParams p, int splits,
                  int k_cols, cudaStream_t st, const void* A2 = nullptr, const void* B2 = nullptr, int K2 = 0) {
    if (p.K % GEMM_BK || p.N % BN || p.M < 1 || splits < 1 || k_cols > p.K || k_cols < 1 || p.N > GEMM_MAX_N)
        return PMG_ERR_BAD_SHAPE;
    if (SEG2 && (!A2 || !B2 || K2 < GEMM_BK || K2 % GEMM_BK)) return PMG_ERR_BAD_SHAPE;
    CUtensorMap ta_hi, ta_lo, tb, ta2, tb2;
    int rc;
    if ((rc = make_tmap_bf16(&ta_hi, A_hi, p.M, k_cols, lda, GEMM_BM))) return rc;
    if ((rc = make_tmap_bf16(&ta_lo, NA == 2 ? A_lo : A_hi, p.M, k_cols, lda, GEMM_BM))) return rc;
    if ((rc = make_tmap_bf16(&tb, B, p.N, k_cols, ldb, BN))) return rc;
    if (SEG2) {
        if ((rc = make_tmap_bf16(&ta2, A2, p.M, K2, lda, GEMM_BM))) return rc;
        if ((rc = make_tmap_bf16(&tb2, B2, p.N, K2, K2, BN))) return rc;
    } else {
        ta2 = ta_hi; tb2 = tb;
    }
    const int total_kb = p.K / GEMM_BK;
    p.kblocks_per_split = (total_kb + splits - 1) / splits;
    int eff_splits = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
    p.K2 = 0; p.kblocks2_per_split = 1;
    if (SEG2) {
        const int total_kb2 = K2 / GEMM_BK;
        p.K2 = K2;
        p.kblocks2_per_split = (total_kb2 + splits - 1) / splits;
        const int eff2 = (total_kb2 + p.kblocks2_per_split - 1) / p.kblocks2_per_split;
        if (eff2 > eff_splits) eff_splits = eff2;
    }
    constexpr int smem = gemm_smem_bytes<BN, NA>();
    static PerDeviceFlag attr_set;
    if (!attr_set.test()) {
        PMG_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, NA, EPI, ACT, SEG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "gemm smem attr");
        attr_set.set();
    }
    const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int items = ((tiles_m + gemm_mt<BN>() - 1) / gemm_mt<BN>()) * eff_splits * (p.N / BN);
    const int grid = items < num_sms() ? items : num_sms();
    PMG_CUDA_OK(launch_pdl(gemm_kernel<BN, NA, EPI, ACT, SEG2>, dim3(grid), dim3(GEMM_THREADS), smem, st, ta_hi, ta_lo, tb, ta2, tb2, p), "tc gemm");
    return PMG_OK;
}

// split-K GEMM on CTA pairs (tc_gemm2.cuh): out[s][M,192] partial sums; A planes [M,K] pitch lda, B [192,K] pitch ldb
template <int NA>
int launch_gemm2_t(const void* A_hi, const void* A_lo, int lda, const void* B, int ldb, int M, int K, int splits, float* out,
                   size_t partial_stride, cudaStream_t st) {
    if (K % GEMM_BK || M < 1 || splits < 1) return PMG_ERR_BAD_SHAPE;
    CUtensorMap ta_hi, ta_lo, tb;
    int rc;
    if ((rc = make_tmap_bf16(&ta_hi, A_hi, M, K, lda, 128))) return rc;
    if ((rc = make_tmap_bf16(&ta_lo, NA == 2 ? A_lo : A_hi, M, K, lda, 128))) return rc;
    if ((rc = make_tmap_bf16(&tb, B, 192, K, ldb, 96))) return rc;
    Gemm2Params p{};
    p.M = M; p.K = K; p.out = out; p.partial_stride = partial_stride;
    const int total_kb = K / GEMM_BK;
    p.kblocks_per_split = (total_kb + splits - 1) / splits;
    const int eff_splits = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
    constexpr int smem = g2_smem_bytes<NA>();
    static PerDeviceFlag attr_set;
    if (!attr_set.test()) {
        PMG_CUDA_OK(cudaFuncSetAttribute(gemm2_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "gemm2 smem attr");
        PMG_CUDA_OK(cudaFuncSetAttribute(gemm2_kernel<NA>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0), "gemm2 cluster attr");
        attr_set.set();
    }
    const int items = ((M + 255) / 256) * eff_splits;
    const int max_pairs = num_sms() / 2;
    const int pairs = items < max_pairs ? items : max_pairs;
    PMG_CUDA_OK(launch_pdl_cluster2(gemm2_kernel<NA>, dim3(2 * pairs), dim3(G2_THREADS), smem, st, ta_hi, ta_lo, tb, p), "tc gemm2");
    return PMG_OK;
}

// number of split-K partial buffers launch_gemm_t will actually write for (K, splits) [and a second segment K2]
inline int effective_splits(int K, int splits, int K2 = 0) {
    const int total_kb = K / GEMM_BK;
    const int per = (total_kb + splits - 1) / splits;
    int eff = (total_kb + per - 1) / per;
    if (K2 > 0) {
        const int kb2 = K2 / GEMM_BK, per2 = (kb2 + splits - 1) / splits, eff2 = (kb2 + per2 - 1) / per2;
        if (eff2 > eff) eff = eff2;
    }
    return eff;
}

}  // namespace tc
