// tc_debug.cu — hardware-semantics probes used while developing the kernels (not a product path).
//   pmg_debug_tmem_a: D[128×64] = A[128×64]·B[64×64]ᵀ with the A operand read from TENSOR MEMORY
//   (tcgen05.mma [d], [a_tmem], b_desc …).  Each thread (= row = TMEM lane) packs its bf16 row two elements
//   per 32-bit column and tcgen05.st's it; tools/tmem_a_probe.py checks the result against torch and thereby the
//   assumed TMEM layout of a 16-bit A operand (column j ↔ K elements 2j, 2j+1; low half = even element).
#include "pmg_common.cuh"
#include "tc_prims.cuh"

namespace tc {

__global__ void __launch_bounds__(128, 1) tmem_a_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                              float* __restrict__ D) {
    __shared__ __align__(1024) uint8_t sB[64 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 256);
    // B [64 n][64 k] bf16 → K-major SWIZZLE_128B tile
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + j * 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    // A row of this thread: 64 bf16 = 32 packed columns at TMEM columns [64, 96)
    uint32_t a[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint4 t = *reinterpret_cast<const uint4*>(A + tid * 64 + j * 8);
        a[4 * j] = t.x; a[4 * j + 1] = t.y; a[4 * j + 2] = t.z; a[4 * j + 3] = t.w;
    }
    {
        uint32_t r0[16], r1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { r0[i] = a[i]; r1[i] = a[16 + i]; }
        tmem_st_32x32b_x16(tb + lane_addr + 64, r0);
        tmem_st_32x32b_x16(tb + lane_addr + 80, r1);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tb, tb + 64 + 8 * k, make_desc_k_sw128(smem_u32(sB) + k * 32), idesc, k > 0 ? 1u : 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32b_x32(tb + lane_addr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[tid * 64 + i] = __uint_as_float(r[i]);
    tmem_ld_32x32b_x32(tb + lane_addr + 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[tid * 64 + 32 + i] = __uint_as_float(r[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

}  // namespace tc

extern "C" int pmg_debug_tmem_a(const void* A, const void* B, float* D, pmg_stream_t stream) {
    if (!A || !B || !D) return PMG_ERR_BAD_ARG;
    tc::tmem_a_probe_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(A), reinterpret_cast<const __nv_bfloat16*>(B), D);
    PMG_LAUNCH_OK("tmem_a_probe");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// pmg_debug_load_probe: how fast can one CTA per SM stream a (B,3,224,224) bf16 image batch into shared memory?
// No compute: a producer thread keeps `stages` loads of 14,336 B in flight, a consumer thread frees a stage as soon as it has
// landed.  mode 0: the front kernel's 4-D tensor-map box (224 x 8 x 4 x 1, 448-byte rows 16 image rows apart);
// mode 1: the same bytes as ONE contiguous 14,336-byte cp.async.bulk (32 consecutive image rows);  mode 2: a 2-D tensor map
// over the batch as rows of 448 B, box = 32 consecutive rows;  mode 3 = mode 0 without the L2 evict-first hint.
// Returns the kernel's device time; tools/load_probe.py turns it into bytes per clock per SM.
// ------------------------------------------------------------------------------------------------
#include "tc_host.cuh"
#include "tc_front.cuh"

namespace tc {

__device__ __forceinline__ void tma_load_4d_nohint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(64, 1) load_probe_kernel(const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tm2,
                                                           const __nv_bfloat16* __restrict__ x, int B, int mode, int stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 12 * 14336);
    uint64_t* empty = full + 12;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int n_tiles = 2 * B;
    if (threadIdx.x == 0) {
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int b = tile >> 1, py0 = (tile & 1) * 7;
            for (int sy = 0; sy < 4; ++sy)
                for (int c = 0; c < 3; ++c, ++it) {
                    const int s = it % stages;
                    mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
                    mbar_expect_tx(&full[s], 14336);
                    uint8_t* dst = smem + s * 14336;
                    if (mode == 0) tma_load_4d(dst, &tm4, &full[s], 0, py0, sy * 4, b * 3 + c);
                    else if (mode == 3) tma_load_4d_nohint(dst, &tm4, &full[s], 0, py0, sy * 4, b * 3 + c);
                    else if (mode == 1) fr_bulk_g2s(dst, x + ((size_t)(b * 3 + c) * 224 + (tile & 1) * 112 + sy * 28) * 224, 14336, &full[s]);
                    else tma_load_2d(dst, &tm2, &full[s], 0, (b * 3 + c) * 224 + (tile & 1) * 112 + sy * 28);
                }
        }
    } else if (threadIdx.x == 32) {
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int j = 0; j < 12; ++j, ++it) {
                const int s = it % stages;
                mbar_wait(&full[s], (it / stages) & 1);
                mbar_arrive(&empty[s]);
            }
    }
}

}  // namespace tc

extern "C" int pmg_debug_load_probe(const void* x, int B, int mode, int stages, float* ms_out, pmg_stream_t stream) {
    if (!x || !ms_out || B < 1 || mode < 0 || mode > 3 || stages < 1 || stages > 12) return PMG_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap tm4, tm2;
    int rc;
    if ((rc = tc::make_tmap_image_rows(&tm4, x, B))) return rc;
    {
        tc::EncodeTiledFn fn = tc::encode_fn();
        if (!fn) return PMG_ERR_CUDA;
        cuuint64_t dims[2] = {224, (cuuint64_t)B * 3 * 224};
        cuuint64_t strides[1] = {448};
        cuuint32_t box[2] = {224, 32};
        cuuint32_t estr[2] = {1, 1};
        if (fn(&tm2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PMG_ERR_CUDA;
    }
    const int smem = 12 * 14336 + 256 + 1024;
    PMG_CUDA_OK(cudaFuncSetAttribute(tc::load_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "probe attr");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 2 * B < tc::num_sms() ? 2 * B : tc::num_sms();
    tc::load_probe_kernel<<<grid, 64, smem, st>>>(tm4, tm2, reinterpret_cast<const __nv_bfloat16*>(x), B, mode, stages);   // warm-up
    cudaEventRecord(e0, st);
    tc::load_probe_kernel<<<grid, 64, smem, st>>>(tm4, tm2, reinterpret_cast<const __nv_bfloat16*>(x), B, mode, stages);
    cudaEventRecord(e1, st);
    PMG_LAUNCH_OK("load_probe");
    PMG_CUDA_OK(cudaEventSynchronize(e1), "probe sync");
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return PMG_OK;
}
