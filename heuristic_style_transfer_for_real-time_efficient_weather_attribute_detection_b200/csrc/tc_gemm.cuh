// tc_gemm.cuh — the tcgen05 + TMA GEMM building block of the PMG fast path.
//
//   D[M, N] = (A_hi[M, K] + A_lo[M, K]) · B[N, K]ᵀ        bf16 operands, fp32 accumulate in TMEM
//
// * A and B are K-major (row-major, K contiguous) bf16 in global memory, fetched by TMA into 128B-swizzled
//   shared-memory stages (BLOCK_K = 64 elements = one 128-byte swizzle row).
// * "Split activations": the activation operand may be given as two bf16 planes (hi = bf16(v),
//   lo = bf16(v − hi)).  Both planes are multiplied by the same B tile into the same accumulator, which
//   carries ~16 significant bits of the fp32 activation through the bf16 tensor cores.
// * Warp-specialised, one 128×BN output tile per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM
//   allocation), warps 2..5 = epilogue (TMEM → registers → global).  Optional split-K over gridDim.y.
#pragma once

#include "tc_prims.cuh"

namespace tc {

enum GemmEpi : int {
    EPI_F32 = 0,        // C fp32 [M, ldc]: act(acc*scale[n] + shift[n]) (+ resid)
    EPI_SPLIT = 1,      // two bf16 planes [M, ldc]: hi/lo split of act(acc*scale[n] + shift[n])
    EPI_PARTIAL = 2     // raw fp32 partial sums [split][M, ldc] (split-K; reduced by the consumer)
};
enum GemmAct : int { GACT_NONE = 0, GACT_LEAKY = 1, GACT_RELU = 2 };

struct GemmParams {
    int M, N, K;              // N = full output width; this launch covers BN columns per blockIdx.z
    int kblocks_per_split;    // BLOCK_K blocks handled by one blockIdx.y
    int epi, act;
    int ldc;
    float* out_f32;           // EPI_F32 / EPI_PARTIAL
    __nv_bfloat16* out_hi;    // EPI_SPLIT
    __nv_bfloat16* out_lo;
    const float* scale;       // [N] or null
    const float* shift;       // [N] or null
    const float* resid;       // [M, ldc] fp32 or null (EPI_F32)
    size_t partial_stride;    // elements between split-K partial buffers
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN>
__host__ __device__ constexpr int gemm_tmem_cols() { return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256; }

template <int BN, int NA>
__host__ __device__ constexpr int gemm_stage_bytes() { return NA * GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2; }

template <int BN, int NA>
__host__ __device__ constexpr int gemm_stages() {
    constexpr int budget = 200 * 1024;
    constexpr int s = budget / gemm_stage_bytes<BN, NA>();
    return s > 8 ? 8 : s;
}

template <int BN, int NA>
__host__ __device__ constexpr int gemm_smem_bytes() { return gemm_stages<BN, NA>() * gemm_stage_bytes<BN, NA>() + 1024 /*align*/ + 256 /*barriers*/; }

template <int BN, int NA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
            const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int STAGES = gemm_stages<BN, NA>();
    constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    constexpr int B_BYTES = BN * GEMM_BK * 2;
    constexpr int STAGE_BYTES = gemm_stage_bytes<BN, NA>();
    constexpr int TMEM_COLS = gemm_tmem_cols<BN>();
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
    static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for the 128B swizzle");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GEMM_BM;
    const int n0 = blockIdx.z * BN;
    const int kb0 = blockIdx.y * p.kblocks_per_split;
    const int total_kb = p.K / GEMM_BK;
    const int kb1 = min(total_kb, kb0 + p.kblocks_per_split);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        if (NA == 2) prefetch_tmap(&tmA_lo);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                const int kc = (kb0 + i) * GEMM_BK;
                tma_load_2d_hint(st, &tmA_hi, &full_bar[s], kc, m0, kEvictFirst);
                if (NA == 2) tma_load_2d_hint(st + A_BYTES, &tmA_lo, &full_bar[s], kc, m0, kEvictFirst);
                tma_load_2d_hint(st + NA * A_BYTES, &tmB, &full_bar[s], kc, n0, kEvictLast);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t b_ad = a_hi + NA * A_BYTES;
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    const uint64_t bd = make_desc_k_sw128(b_ad + k * 32);
                    umma_bf16(tmem_base, make_desc_k_sw128(a_hi + k * 32), bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    if (NA == 2) umma_bf16(tmem_base, make_desc_k_sw128(a_hi + A_BYTES + k * 32), bd, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);          // frees the stage when these MMAs have read it
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const bool in_m = row < p.M;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (!in_m) continue;
            const int nb = n0 + c0;
            if (p.epi == EPI_PARTIAL) {
                float* dst = p.out_f32 + (size_t)blockIdx.y * p.partial_stride + (size_t)row * p.ldc + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (c0 + j < BN)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                          __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                continue;
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(r[j]);
                if (c0 + j < BN) {
                    if (p.scale) x *= __ldg(p.scale + nb + j);
                    if (p.shift) x += __ldg(p.shift + nb + j);
                    if (p.act == GACT_LEAKY) x = x >= 0.f ? x : 0.2f * x;
                    else if (p.act == GACT_RELU) x = fmaxf(x, 0.f);
                }
                v[j] = x;
            }
            if (p.epi == EPI_F32) {
                float* dst = p.out_f32 + (size_t)row * p.ldc + nb;
                const float* rs = p.resid ? p.resid + (size_t)row * p.ldc + nb : nullptr;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (c0 + j >= BN) break;
                    float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (rs) { const float4 t = *reinterpret_cast<const float4*>(rs + j); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                    *reinterpret_cast<float4*>(dst + j) = o;
                }
            } else {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                uint4* dh = reinterpret_cast<uint4*>(p.out_hi + (size_t)row * p.ldc + nb);
                uint4* dl = reinterpret_cast<uint4*>(p.out_lo + (size_t)row * p.ldc + nb);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c0 + j * 8 >= BN) break;
                    dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
