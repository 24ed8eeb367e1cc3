// tc_mid.cuh — conv2 (1×1, 64→128) + BN + LeakyReLU + chan_proj (1×1, 128→96) fused in ONE persistent tcgen05
// kernel: the 128-channel activation never leaves the SM.
// Reference: models_PMG.py:124-128 (final 1×1 of feature_extractor), :136,182 (chan_proj).
//
//   act1 hi/lo [P, 64] bf16 (P = B·196 pixels)  ──TMA──►  conv2 MMA (M128 N128, K 64, hi+lo)  ──►  TMEM D2
//   epilogue A (16 warps): D2 → BN scale/shift → LeakyReLU → hi/lo split → smem A3 (K-major SW128, [128 × 128] ×2)
//   chan MMA (M128 N96, K 128, hi+lo) ──► TMEM D3 ──► epilogue B (4 warps): smem-staged, fully coalesced fp32 stores
//   chan [P, 96] fp32 pixel-major (the Gram kernel's input).
// Both weight matrices (40 KB) stay resident in smem; D2 and D3 are double-buffered in TMEM so tile t+1's conv2
// runs under tile t's epilogues; A3 is single-buffered (epilogue A of tile t+1 waits for the chan MMAs of tile t).
// Algorithmic HBM bytes per pixel: 256 in + 384 out (the two separate GEMMs moved 256 + 512 + 512 + 384).
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int MID_THREADS = 704;                 // warp 0 TMA, 1 MMA, 2-17 epilogue A (lane quarter × 32-column chunk), 18-21 epilogue B
constexpr int MID_STAGES = 3;
constexpr int MID_A_STAGE = 2 * 128 * 128;       // hi + lo [128 rows × 64 K]
constexpr int MID_W2_BYTES = 128 * 128;          // [128 n × 64 k]
constexpr int MID_W3_BYTES = 2 * 96 * 128;       // 2 k-blocks × [96 n × 64 k]
constexpr int MID_A3_BYTES = 2 * 2 * 128 * 128;  // [plane][kb] × [128 rows × 64 K]
constexpr int MID_OFF_W3 = MID_W2_BYTES;
constexpr int MID_OFF_A = MID_OFF_W3 + MID_W3_BYTES;                 // 40,960 (1024-aligned)
constexpr int MID_OFF_A3 = MID_OFF_A + MID_STAGES * MID_A_STAGE;
constexpr int MID_OFF_STG = MID_OFF_A3 + MID_A3_BYTES;               // 4 × 4 KB epilogue-B staging
constexpr int MID_OFF_BAR = MID_OFF_STG + 4 * 4096;
constexpr int MID_OFF_PRM = MID_OFF_BAR + 256;
constexpr int MID_SMEM = MID_OFF_PRM + 2 * 128 * 4 + 1024;
static_assert(MID_SMEM <= 232448, "mid kernel exceeds the shared-memory limit");

struct MidParams {
    const float* scale2; const float* shift2;    // [128] BN(eval) of conv2 folded
    float* chan;                                 // [P, 96] fp32
    int P;
};

__global__ void __launch_bounds__(MID_THREADS, 1)
mid_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const MidParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW2 = smem;
    uint8_t* sW3 = smem + MID_OFF_W3;
    uint8_t* sA = smem + MID_OFF_A;
    uint8_t* sA3 = smem + MID_OFF_A3;
    uint8_t* sStg = smem + MID_OFF_STG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MID_OFF_BAR);
    uint64_t* a_full = bars;             // [MID_STAGES ≤ 4]
    uint64_t* a_empty = bars + 4;        // [MID_STAGES ≤ 4]
    uint64_t* d2_full = bars + 8;        // [2]
    uint64_t* d2_empty = bars + 10;      // [2]
    uint64_t* a3_ready = bars + 12;
    uint64_t* a3_free = bars + 13;
    uint64_t* d3_full = bars + 14;       // [2]
    uint64_t* d3_empty = bars + 16;      // [2]
    uint64_t* w_bar = bars + 18;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
    float* s2 = reinterpret_cast<float*>(smem + MID_OFF_PRM);
    float* h2 = s2 + 128;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_tiles = (p.P + 127) / 128;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmW2); prefetch_tmap(&tmW3);
        for (int i = 0; i < MID_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 16);
            mbar_init(&d3_full[i], 1); mbar_init(&d3_empty[i], 4);
        }
        mbar_init(a3_ready, 16); mbar_init(a3_free, 1); mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 128) { s2[threadIdx.x] = p.scale2[threadIdx.x]; h2[threadIdx.x] = p.shift2[threadIdx.x]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD2 = tmem_base, tD3 = tmem_base + 256;      // D2: 2 × 128 cols, D3: 2 × 96 cols (128 apart)
    pdl_trigger();

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(w_bar, MID_W2_BYTES + MID_W3_BYTES);
            tma_load_2d_hint(sW2, &tmW2, w_bar, 0, 0, kEvictLast);
            tma_load_2d_hint(sW3, &tmW3, w_bar, 0, 0, kEvictLast);
            tma_load_2d_hint(sW3 + 96 * 128, &tmW3, w_bar, 64, 0, kEvictLast);
            pdl_wait();                  // act1 comes from the front kernel
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int s = it % MID_STAGES;
                mbar_wait(&a_empty[s], ((it / MID_STAGES) & 1) ^ 1);
                mbar_expect_tx(&a_full[s], MID_A_STAGE);
                tma_load_2d_hint(sA + s * MID_A_STAGE, &tmA_hi, &a_full[s], 0, tile * 128, kEvictFirst);
                tma_load_2d_hint(sA + s * MID_A_STAGE + 128 * 128, &tmA_lo, &a_full[s], 0, tile * 128, kEvictFirst);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t id96 = make_idesc_bf16(128, 96, 0, 0);
            const uint32_t aW2 = smem_u32(sW2), aW3 = smem_u32(sW3), aA = smem_u32(sA), aA3 = smem_u32(sA3);
            mbar_wait(w_bar, 0);
            auto chan_mma = [&](uint32_t t) {                       // D3[t&1] = A3(hi+lo) · W3ᵀ for local tile t
                const uint32_t b = t & 1;
                mbar_wait(a3_ready, t & 1);
                mbar_wait(&d3_empty[b], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = make_desc_k_sw128(aW3 + kb * (96 * 128) + k * 32);
                        umma_bf16(tD3 + 128 * b, make_desc_k_sw128(aA3 + kb * 16384 + k * 32), bd, id96, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(tD3 + 128 * b, make_desc_k_sw128(aA3 + 32768 + kb * 16384 + k * 32), bd, id96, 1u);
                    }
                umma_commit(a3_free);
                umma_commit(&d3_full[b]);
            };
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % MID_STAGES, b = it & 1;
                mbar_wait(&d2_empty[b], ((it >> 1) & 1) ^ 1);
                mbar_wait(&a_full[s], (it / MID_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = aA + s * MID_A_STAGE, a_lo = a_hi + 128 * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = make_desc_k_sw128(aW2 + k * 32);
                    umma_bf16(tD2 + 128 * b, make_desc_k_sw128(a_hi + k * 32), bd, id128, k > 0 ? 1u : 0u);
                    umma_bf16(tD2 + 128 * b, make_desc_k_sw128(a_lo + k * 32), bd, id128, 1u);
                }
                umma_commit(&a_empty[s]);
                umma_commit(&d2_full[b]);
                if (it >= 1) chan_mma(it - 1);
            }
            if (it >= 1) chan_mma(it - 1);
        }
    } else if (warp < 18) {
        // ===================== epilogue A: conv2 → BN + LeakyReLU → hi/lo → A3 (chan_proj's operand) =====================
        const int q = warp & 3, cq = (warp - 2) >> 2;         // lane quarter, 32-column chunk (K block cq/2 of A3, chunks 4(cq%2)..+3)
        const int m = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint8_t* a3row = sA3 + (cq >> 1) * 16384 + m * 128;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1;
            mbar_wait(&d2_full[b], (it >> 1) & 1);
            tc_fence_after();
            uint32_t r[32], H[16], L[16];
            tmem_ld_32x32b_x32(tD2 + 128 * b + lane_addr + (uint32_t)(cq * 32), r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[b]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 sc = *reinterpret_cast<const float4*>(s2 + cq * 32 + 4 * j);
                const float4 sh = *reinterpret_cast<const float4*>(h2 + cq * 32 + 4 * j);
                float v0 = fmaf(__uint_as_float(r[4 * j]), sc.x, sh.x), v1 = fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y);
                float v2 = fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z), v3 = fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w);
                v0 = fmaxf(v0, 0.2f * v0); v1 = fmaxf(v1, 0.2f * v1); v2 = fmaxf(v2, 0.2f * v2); v3 = fmaxf(v3, 0.2f * v3);
                split_bf16x2(v0, v1, H[2 * j], L[2 * j]);
                split_bf16x2(v2, v3, H[2 * j + 1], L[2 * j + 1]);
            }
            if (it > 0) mbar_wait(a3_free, (it - 1) & 1);      // chan MMAs of the previous tile have read A3
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sw = ((4 * (cq & 1) + j) ^ (m & 7)) << 4;
                *reinterpret_cast<uint4*>(a3row + sw) = make_uint4(H[4 * j], H[4 * j + 1], H[4 * j + 2], H[4 * j + 3]);
                *reinterpret_cast<uint4*>(a3row + 32768 + sw) = make_uint4(L[4 * j], L[4 * j + 1], L[4 * j + 2], L[4 * j + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a3_ready);
        }
    } else {
        // ===================== epilogue B: chan_proj accumulator → fp32 [P, 96], smem-staged coalesced stores =====================
        const int q = warp & 3;                                // lane quarter (warps 18..21 → 2, 3, 0, 1)
        uint8_t* wbuf = sStg + (warp - 18) * 4096;
        const int sr = lane >> 3, sc = lane & 7;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t b = it & 1;
            mbar_wait(&d3_full[b], (it >> 1) & 1);
            tc_fence_after();
            const int row0 = tile * 128 + q * 32 + sr;
            const int rows_left = p.P - row0;
#pragma unroll 1
            for (int c0 = 0; c0 < 96; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tD3 + 128 * b + lane_addr + (uint32_t)c0, r);
                tmem_ld_wait();
                if (c0 == 64) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d3_empty[b]);
                }
                __syncwarp();
                stage_write_row(wbuf, lane, r);
                __syncwarp();
                float* dst = p.chan + (size_t)row0 * 96 + c0 + sc * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = 4 * i + sr;
                    if (4 * i < rows_left)
                        *reinterpret_cast<uint4*>(dst + (size_t)(4 * i) * 96) = *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
