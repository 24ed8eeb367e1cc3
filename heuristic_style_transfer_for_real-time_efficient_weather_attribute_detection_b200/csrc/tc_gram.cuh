// tc_gram.cuh — per-patch Gram matrices of the PMG path on the tcgen05 tensor cores.
//
// Reference: models_PMG.py:157-174 (replicate-pad 14→15, 3×3 disjoint 5×5 patches) and :188
//            G_k = P_k·P_kᵀ / 25,  P_k ∈ (96 channels, 25 pixels).
//
// One persistent CTA loops over images.  For each image
//   * one thread streams the whole pixel-major fp32 chan_proj map (196×96×4 = 75 264 B, contiguous) into smem
//     with a 1-D bulk TMA copy (cp.async.bulk), double-buffered across images;
//   * for each of the 9 patches, 96 loader threads (thread = channel) read the patch's 25 pixels from smem
//     (replicate padding = index clamp), split every value into two bf16 planes (hi, lo) and write them
//     TRANSPOSED into a K-major, 128B-swizzled operand tile [channel row][pixel], pixels 25..31 zero;
//   * one thread issues 6 tcgen05.mma (M=128 rows incl. 32 pad rows, N=96, K=2×16):
//         D = Ph·Phᵀ + Ph·Plᵀ + Pl·Phᵀ            (fp32 in TMEM; Pl·Plᵀ ≈ 2⁻¹⁸ relative is dropped)
//     the SAME smem tile serves as A and B operand;
//   * 3 epilogue warps pull D out of TMEM, scale by 1/25, split into hi/lo bf16 and store the token's
//     9216-entry row that token_proj consumes (192-byte contiguous runs per thread), plus fp32 G on request.
// Image buffers (2), operand stages (2) and TMEM accumulators (2) are multi-buffered, so the map load, the
// transposes, the MMAs and the 36 KB/token store stream overlap.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int GRAM_THREADS = 224;      // warps 0-2 loaders, 3 MMA, 4-6 epilogue
constexpr int GRAM_STAGES = 2;
constexpr int GRAM_PLANE = 128 * 128;  // bytes: 128 rows × 128 B (64 pixel slots, 32 used)
constexpr int GRAM_IMG_BYTES = 196 * 96 * 4;
constexpr int GRAM_IMG_SLOT = 76 * 1024;
constexpr int GRAM_SMEM = GRAM_STAGES * 2 * GRAM_PLANE + 2 * GRAM_IMG_SLOT + 1024 + 256;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram96_tc_kernel(const float* __restrict__ chan, __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                 float* __restrict__ g_f32, int n_images) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* img_buf = smem + GRAM_STAGES * 2 * GRAM_PLANE;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(img_buf + 2 * GRAM_IMG_SLOT);
    uint64_t* empty_bar = full_bar + GRAM_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GRAM_STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint64_t* img_full_bar = tmem_empty_bar + 2;         // [2]
    uint64_t* img_empty_bar = img_full_bar + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(img_empty_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GRAM_STAGES; ++s) { mbar_init(&full_bar[s], 96); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 3);
            mbar_init(&img_full_bar[a], 1); mbar_init(&img_empty_bar[a], 96);
        }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc(tmem_slot, 256);
    // pad rows 96..127 of every plane are never written by the loaders: clear them once so no NaN
    // patterns feed the (ignored) accumulator rows
    for (int i = threadIdx.x; i < GRAM_STAGES * 2 * (32 * 128 / 16); i += GRAM_THREADS) {
        const int plane = i / 256, off = i % 256;
        reinterpret_cast<uint4*>(smem + plane * GRAM_PLANE + 96 * 128)[off] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 3) {
        // ===== loaders: thread = channel =====
        const int ch = threadIdx.x;
        auto issue_image = [&](int img, uint32_t li) {          // thread 0 only
            const uint32_t ib = li & 1, iph = (li >> 1) & 1;
            mbar_wait(&img_empty_bar[ib], iph ^ 1);
            mbar_expect_tx(&img_full_bar[ib], GRAM_IMG_BYTES);
            bulk_load_1d(img_buf + ib * GRAM_IMG_SLOT, chan + (size_t)img * (196 * 96), GRAM_IMG_BYTES, &img_full_bar[ib]);
        };
        if (ch == 0 && (int)blockIdx.x < n_images) issue_image(blockIdx.x, 0);
        uint32_t it = 0, li = 0;
        for (int img = blockIdx.x; img < n_images; img += gridDim.x, ++li) {
            const uint32_t ib = li & 1, iph = (li >> 1) & 1;
            if (ch == 0 && img + (int)gridDim.x < n_images) issue_image(img + gridDim.x, li + 1);   // prefetch next map
            mbar_wait(&img_full_bar[ib], iph);
            const float* simg = reinterpret_cast<const float*>(img_buf + ib * GRAM_IMG_SLOT) + ch;
#pragma unroll 1
            for (int k = 0; k < 9; ++k, ++it) {
                const int py = k / 3, px = k % 3;
                const int s = it % GRAM_STAGES;
                const uint32_t ph = (it / GRAM_STAGES) & 1;
                float v[25];
#pragma unroll
                for (int n = 0; n < 25; ++n) {
                    const int y = min(py * 5 + n / 5, 13), x = min(px * 5 + n % 5, 13);
                    v[n] = simg[(y * 14 + x) * 96];
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 12; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                split_bf16x2(v[24], 0.f, hi[12], lo[12]);
                hi[13] = hi[14] = hi[15] = 0; lo[13] = lo[14] = lo[15] = 0;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* ph_base = smem + s * 2 * GRAM_PLANE + ch * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int sw = (c ^ (ch & 7)) * 16;
                    *reinterpret_cast<uint4*>(ph_base + sw) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    *reinterpret_cast<uint4*>(ph_base + GRAM_PLANE + sw) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
                fence_proxy_async();                           // generic-proxy smem writes → visible to the tensor core
                mbar_arrive(&full_bar[s]);
            }
            mbar_arrive(&img_empty_bar[ib]);                   // this thread is done reading the map
        }
    } else if (warp == 3) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
            uint32_t it = 0;
            for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
                for (int k = 0; k < 9; ++k, ++it) {
                    const int s = it % GRAM_STAGES;
                    const uint32_t ph = (it / GRAM_STAGES) & 1;
                    const uint32_t as = it & 1, aph = (it >> 1) & 1;
                    mbar_wait(&tmem_empty_bar[as], aph ^ 1);
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t p_hi = smem_u32(smem + s * 2 * GRAM_PLANE), p_lo = p_hi + GRAM_PLANE;
                    const uint32_t d = tmem_base + as * 128;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t dh = make_desc_k_sw128(p_hi + kk * 32), dl = make_desc_k_sw128(p_lo + kk * 32);
                        umma_bf16(d, dh, dh, idesc, kk > 0 ? 1u : 0u);
                        umma_bf16(d, dh, dl, idesc, 1u);
                        umma_bf16(d, dl, dh, idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);
                    umma_commit(&tmem_full_bar[as]);
                }
            }
        }
    } else {
        // ===== epilogue: warps 4,5,6 ↔ TMEM lane quarters 0,1,2 (rows 0..95) =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t it = 0;
        for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
            for (int k = 0; k < 9; ++k, ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                mbar_wait(&tmem_full_bar[as], aph);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + as * 128 + ((uint32_t)(q * 32) << 16);
                const size_t o = ((size_t)img * 9 + k) * 9216 + (size_t)row * 96;
#pragma unroll 1
                for (int c0 = 0; c0 < 96; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
                    tmem_ld_wait();
                    if (c0 == 64) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        split_bf16x2(__uint_as_float(r[2 * j]) * 0.04f, __uint_as_float(r[2 * j + 1]) * 0.04f, hi[j], lo[j]);
                    uint4* dh = reinterpret_cast<uint4*>(g_hi + o + c0);
                    uint4* dl = reinterpret_cast<uint4*>(g_lo + o + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    if (g_f32) {
                        float4* df = reinterpret_cast<float4*>(g_f32 + o + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            df[j] = make_float4(__uint_as_float(r[4 * j]) * 0.04f, __uint_as_float(r[4 * j + 1]) * 0.04f,
                                                __uint_as_float(r[4 * j + 2]) * 0.04f, __uint_as_float(r[4 * j + 3]) * 0.04f);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc
