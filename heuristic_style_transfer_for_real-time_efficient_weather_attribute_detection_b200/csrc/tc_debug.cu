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
