// tc_gram.cuh — per-patch Gram matrices of the PMG path on the tcgen05 tensor cores.
//
// Reference: models_PMG.py:157-174 (replicate-pad 14→15, 3×3 disjoint 5×5 patches) and :188
//            G_k = P_k·P_kᵀ / 25,  P_k ∈ (96 channels, 25 pixels).
//
// One persistent CTA loops over tokens (image, patch).  For each token
//   * 96 loader threads (thread = channel) read the patch's 25 pixels from the pixel-major fp32 chan_proj map
//     (coalesced over channels), split every value into two bf16 planes (hi, lo) and write them TRANSPOSED
//     into a K-major, 128B-swizzled smem operand tile [channel row][pixel], pixels 25..31 zero;
//   * one thread issues 6 tcgen05.mma (M=128 rows incl. 32 pad rows, N=96, K=2×16):
//         D = Ph·Phᵀ + Ph·Plᵀ + Pl·Phᵀ            (fp32 in TMEM; Pl·Plᵀ ≈ 2⁻¹⁸ relative is dropped)
//     the SAME smem tile serves as A and B operand, so P is read from global once;
//   * 9 epilogue warps (one per 32×32 block) pull D out of TMEM, scale by 1/25, split into hi/lo bf16 and store the
//     6 upper blocks (G is symmetric) as the token's 6144-entry row that token_proj consumes, plus fp32 G on request.
// smem operand stages (3) and TMEM accumulators (2) are multi-buffered, so loads, MMAs and the 36 KB/token
// store stream overlap.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int GRAM_THREADS = 512;     // warps 0-2 loaders, 3 MMA, 4-15 epilogue: (lane quarter w%4 < 3) × (32-col chunk (w-4)/4)
constexpr int GRAM_STAGES = 3;
constexpr int GRAM_PLANE = 128 * 128; // bytes: 128 rows × 128 B (64 pixel slots, 32 used)
constexpr int GRAM_SMEM = GRAM_STAGES * 2 * GRAM_PLANE + 1024 + 256 + 12 * 4096 /*epilogue staging*/;

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram96_tc_kernel(const float* __restrict__ chan, __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                 float* __restrict__ g_f32, int n_tokens) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GRAM_STAGES * 2 * GRAM_PLANE);
    uint64_t* empty_bar = full_bar + GRAM_STAGES;
    uint64_t* tmem_full_bar = empty_bar + GRAM_STAGES;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* s_stage = smem + GRAM_STAGES * 2 * GRAM_PLANE + 256;             // 12 × 4 KB
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform

    if (threadIdx.x == 0) {
        for (int s = 0; s < GRAM_STAGES; ++s) { mbar_init(&full_bar[s], 96); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 9); }
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
    pdl_trigger();
    pdl_wait();          // chan comes from the previous kernel

    const bool loader_a = warp < 3;                               // tokens with even local index
    const bool loader_b = warp >= 4 && (warp & 3) == 3;          // warps 7, 11, 15 (no TMEM rows to drain): odd tokens
    if (loader_a || loader_b) {
        // ===== loaders: two groups of 96 threads (thread = channel) alternate tokens, so two tokens' global
        // loads are in flight per CTA =====
        const int grp = loader_a ? 0 : 1;
        const int ch = loader_a ? threadIdx.x : ((warp - 7) >> 2) * 32 + lane;
        float v[25];
        auto load_token = [&](int tok) {
            const int b = tok / 9, k = tok % 9, py = k / 3, px = k % 3;
            const float* src = chan + (size_t)b * (196 * 96) + ch;
#pragma unroll
            for (int n = 0; n < 25; ++n) {
                const int y = min(py * 5 + n / 5, 13), x = min(px * 5 + n % 5, 13);
                v[n] = __ldg(src + (y * 14 + x) * 96);
            }
        };
        const int step = 2 * gridDim.x;
        int tok = blockIdx.x + grp * gridDim.x;
        if (tok < n_tokens) load_token(tok);
        for (uint32_t it = grp; tok < n_tokens; tok += step, it += 2) {
            const int s = it % GRAM_STAGES;
            const uint32_t ph = (it / GRAM_STAGES) & 1;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 12; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            split_bf16x2(v[24], 0.f, hi[12], lo[12]);
            hi[13] = hi[14] = hi[15] = 0; lo[13] = lo[14] = lo[15] = 0;
            if (tok + step < n_tokens) load_token(tok + step);   // prefetch this group's next token
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
    } else if (warp == 3) {
        // ===== MMA issuer =====
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas (no per-MMA ELECT loops)
            constexpr uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
            uint32_t it = 0;
            for (int tok = blockIdx.x; tok < n_tokens; tok += gridDim.x, ++it) {
                const int s = it % GRAM_STAGES;
                const uint32_t ph = (it / GRAM_STAGES) & 1;
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aph ^ 1);
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t p_hi = smem_u32(smem + s * 2 * GRAM_PLANE), p_lo = p_hi + GRAM_PLANE;
                const uint32_t d = tmem_base + as * 128;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t dh = make_desc_k_sw128(p_hi + k * 32), dl = make_desc_k_sw128(p_lo + k * 32);
                    umma_bf16(d, dh, dh, idesc, k > 0 ? 1u : 0u);
                    umma_bf16(d, dh, dl, idesc, 1u);
                    umma_bf16(d, dl, dh, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        // ===== epilogue: 9 warps, warp w ↔ 32×32 block (q, ch) of G: TMEM lane quarter q = w%4 (rows 32q..32q+31,
        // q < 3) and 32-column chunk ch = (w-4)/4.  G is symmetric, so only the 6 upper blocks (ch ≥ q) are stored for
        // token_proj (which runs on weights folded accordingly, pmg_fast.cu: pack_wt_folded_kernel): per token
        // 6144 instead of 9216 entries, block order (0,0) (1,1) (2,2) | (0,1) (0,2) (1,2), row-major inside a block.
        const int q = warp & 3;
        {
            const int ch = (warp - 4) >> 2;
            const int c0 = ch * 32;
            const bool upper = ch >= q;
            const int blk = ch == q ? q : q + ch + 2;
            uint8_t* wbuf = s_stage + (warp - 4) * 4096;
            const int sr = lane >> 3, sc = lane & 7;
            uint32_t it = 0;
            for (int tok = blockIdx.x; tok < n_tokens; tok += gridDim.x, ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                mbar_wait(&tmem_full_bar[as], aph);
                tc_fence_after();
                uint32_t r[32];
                if (upper || g_f32) {
                    tmem_ld_32x32b_x32(tmem_base + as * 128 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                if (upper) {
                    uint32_t w[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        split_bf16x2(__uint_as_float(r[2 * j]) * 0.04f, __uint_as_float(r[2 * j + 1]) * 0.04f, w[j], w[16 + j]);
                    // row slice [hi 64 B | lo 64 B] → warp-private swizzled staging → 8 lanes per row: one store
                    // instruction covers 4 rows = 256 contiguous bytes of the block in each plane
                    __syncwarp();
                    stage_write_row(wbuf, lane, w);
                    __syncwarp();
                    const size_t tbase = (size_t)tok * 6144 + (size_t)blk * 1024 + (size_t)sr * 32 + (sc & 3) * 8;
                    __nv_bfloat16* dst = (sc < 4 ? g_hi : g_lo) + tbase;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + sr;
                        *reinterpret_cast<uint4*>(dst + (size_t)(4 * i) * 32) =
                            *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                    }
                }
                if (g_f32) {
                    float4* df = reinterpret_cast<float4*>(g_f32 + (size_t)tok * 9216 + (size_t)(q * 32 + lane) * 96 + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        df[j] = make_float4(__uint_as_float(r[4 * j]) * 0.04f, __uint_as_float(r[4 * j + 1]) * 0.04f,
                                            __uint_as_float(r[4 * j + 2]) * 0.04f, __uint_as_float(r[4 * j + 3]) * 0.04f);
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
