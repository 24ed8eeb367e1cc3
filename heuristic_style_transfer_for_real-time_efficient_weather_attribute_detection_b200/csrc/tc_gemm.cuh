// tc_gemm.cuh — the tcgen05 + TMA GEMM building block of the PMG fast path.
//
//   D[M, N] = (A_hi[M, K] + A_lo[M, K]) · B[N, K]ᵀ        bf16 operands, fp32 accumulate in TMEM
//
// * A and B are K-major (row-major, K contiguous) bf16 in global memory, fetched by TMA into 128B-swizzled
//   shared-memory stages (BLOCK_K = 64 elements = one 128-byte swizzle row).
// * "Split activations": the activation operand may be given as two bf16 planes (hi = bf16(v),
//   lo = bf16(v − hi)).  Both planes are multiplied by the same B tile into the same accumulator, which
//   carries ~16 significant bits of the fp32 activation through the bf16 tensor cores.
// * Warp-specialised, persistent over 128×BN output tiles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM
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
    // SEG2 kernels only: a second K segment  D += A2[M, K2] · B2[N, K2]ᵀ  (ONE activation plane) accumulated into the same
    // TMEM tile, split like the first one — token_proj's lo-weight term rides in the main GEMM instead of being a second
    // launch with its own partial buffers
    int K2, kblocks2_per_split;
    float2* stats;            // optional (EPI_F32): per (M tile, epilogue warp q) column statistics [tiles_m·4][N] (mean, M2) of the stored values
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 = two epilogue groups
constexpr int GEMM_MAX_N = 576;     // scale/shift staged in smem for the whole output width

// Work item = MT consecutive 128-row M-tiles sharing one TMEM accumulator stage (MT·BN columns): for narrow
// outputs (conv0: N = 32, K = 48) this amortises the per-tile barrier round trips over 4 tiles.
template <int BN>
__host__ __device__ constexpr int gemm_mt() { return BN <= 32 ? 4 : BN <= 64 ? 2 : 1; }

template <int BN>
__host__ __device__ constexpr int gemm_acc_cols() {
    constexpr int c = BN * gemm_mt<BN>();
    return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : 256;
}

template <int BN, int NA>
__host__ __device__ constexpr int gemm_stage_bytes() { return NA * GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2; }

template <int BN, int NA>
__host__ __device__ constexpr int gemm_stages() {
    constexpr int budget = 180 * 1024;
    constexpr int s = budget / gemm_stage_bytes<BN, NA>();
    return s > 8 ? 8 : s;
}

template <int BN, int NA>
__host__ __device__ constexpr int gemm_smem_bytes() {
    return gemm_stages<BN, NA>() * gemm_stage_bytes<BN, NA>() + 1024 /*align*/ + 256 /*barriers*/ + 2 * GEMM_MAX_N * 4 +
           8 * 4096 /*epilogue staging, 4 KB per epilogue warp*/;
}

// Persistent: CTA c processes work items c, c+gridDim.x, …; item = (split, n_tile, m_group), m fastest.
// Two TMEM accumulator stages let the epilogue of item i overlap the MMAs of item i+1.
template <int BN, int NA, int EPI, int ACT, bool SEG2 = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
    constexpr int STAGES = gemm_stages<BN, NA>();
    constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    constexpr int B_BYTES = BN * GEMM_BK * 2;
    constexpr int STAGE_BYTES = gemm_stage_bytes<BN, NA>();
    constexpr int MT = gemm_mt<BN>();
    constexpr int ACC_COLS = gemm_acc_cols<BN>();
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N / epilogue chunking");
    static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment for the 128B swizzle");

    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic on the __shared__ array (not an integer round-trip) keeps the shared address space,
    // so the epilogue's scale/shift reads compile to LDS instead of generic LD
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_scale = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);
    float* s_shift = s_scale + GEMM_MAX_N;
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_shift + GEMM_MAX_N);      // 8 × 4 KB, 128-byte aligned

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int groups_m = (tiles_m + MT - 1) / MT;
    const int tiles_n = p.N / BN;
    const int total_kb = p.K / GEMM_BK;
    const int total_kb2 = SEG2 ? p.K2 / GEMM_BK : 0;
    const int nsplit1 = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
    const int nsplit2 = SEG2 ? (total_kb2 + p.kblocks2_per_split - 1) / p.kblocks2_per_split : 0;
    const int nsplit = nsplit1 > nsplit2 ? nsplit1 : nsplit2;
    const int total_items = groups_m * tiles_n * nsplit;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        if (NA == 2) prefetch_tmap(&tmA_lo);
        prefetch_tmap(&tmB);
        if (SEG2) { prefetch_tmap(&tmA2); prefetch_tmap(&tmB2); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    if (EPI != EPI_PARTIAL) {
        for (int i = threadIdx.x; i < p.N; i += GEMM_THREADS) {
            s_scale[i] = p.scale ? p.scale[i] : 1.f;
            s_shift[i] = p.shift ? p.shift[i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();          // A (and resid) come from the previous kernel

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas (no per-MMA ELECT loops)
            uint32_t it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const int mg = item % groups_m, nt = (item / groups_m) % tiles_n, sp = item / (groups_m * tiles_n);
                const int kb0 = min(total_kb, sp * p.kblocks_per_split), kb1 = min(total_kb, kb0 + p.kblocks_per_split);
                for (int j = 0; j < MT; ++j) {
                    const int mt = mg * MT + j;
                    if (mt >= tiles_m) break;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* st = smem + s * STAGE_BYTES;
                        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                        tma_load_2d_hint(st, &tmA_hi, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM, kEvictFirst);
                        if (NA == 2) tma_load_2d_hint(st + A_BYTES, &tmA_lo, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM, kEvictFirst);
                        tma_load_2d_hint(st + NA * A_BYTES, &tmB, &full_bar[s], kb * GEMM_BK, nt * BN, kEvictLast);
                    }
                    if (SEG2) {
                        const int kc0 = min(total_kb2, sp * p.kblocks2_per_split), kc1 = min(total_kb2, kc0 + p.kblocks2_per_split);
                        for (int kb = kc0; kb < kc1; ++kb, ++it) {
                            const int s = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            uint8_t* st = smem + s * STAGE_BYTES;
                            mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
                            tma_load_2d_hint(st, &tmA2, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM, kEvictFirst);
                            tma_load_2d_hint(st + NA * A_BYTES, &tmB2, &full_bar[s], kb * GEMM_BK, nt * BN, kEvictLast);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas (no per-MMA ELECT loops)
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0, 0);
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
                const int mg = item % groups_m, sp = item / (groups_m * tiles_n);
                const int kb0 = min(total_kb, sp * p.kblocks_per_split), kb1 = min(total_kb, kb0 + p.kblocks_per_split);
                const uint32_t as = li & 1, aph = (li >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aph ^ 1);     // epilogue has drained this accumulator stage
                tc_fence_after();
                for (int j = 0; j < MT; ++j) {
                    if (mg * MT + j >= tiles_m) break;
                    const uint32_t d_tmem = tmem_base + as * ACC_COLS + j * BN;
                    uint32_t acc = 0;                           // 0 for the very first MMA of the tile
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                        const uint32_t b_ad = a_hi + NA * A_BYTES;
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t bd = make_desc_k_sw128(b_ad + k * 32);
                            umma_bf16(d_tmem, make_desc_k_sw128(a_hi + k * 32), bd, idesc, acc);
                            acc = 1u;
                            if (NA == 2) umma_bf16(d_tmem, make_desc_k_sw128(a_hi + A_BYTES + k * 32), bd, idesc, 1u);
                        }
                        umma_commit(&empty_bar[s]);          // frees the smem stage when these MMAs have read it
                    }
                    if (SEG2) {
                        const int kc0 = min(total_kb2, sp * p.kblocks2_per_split), kc1 = min(total_kb2, kc0 + p.kblocks2_per_split);
                        for (int kb = kc0; kb < kc1; ++kb, ++it) {
                            const int s = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            mbar_wait(&full_bar[s], ph);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                            const uint32_t b_ad = a_hi + NA * A_BYTES;
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; ++k) {
                                umma_bf16(d_tmem, make_desc_k_sw128(a_hi + k * 32), make_desc_k_sw128(b_ad + k * 32), idesc, acc);
                                acc = 1u;
                            }
                            umma_commit(&empty_bar[s]);
                        }
                    }
                }
                umma_commit(&tmem_full_bar[as]);             // accumulator stage complete
            }
        }
    } else {
        // ===== epilogue: group g = (warp-2)/4 drains accumulator stage g (items li ≡ g mod 2), so two
        // items' epilogues are in flight per CTA; TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        uint8_t* wbuf = s_stage + (warp - 2) * 4096;
        // staged-store geometry is per-lane constant: 8 lanes per row, lane&7 = 16-byte chunk of the 128-byte slice
        const int sr = lane >> 3, sc = lane & 7;
        uint32_t li = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
            if ((li & 1) != grp) continue;
            const int mg = item % groups_m, nt = (item / groups_m) % tiles_n, sp = item / (groups_m * tiles_n);
            const uint32_t aph = (li >> 1) & 1;
            mbar_wait(&tmem_full_bar[grp], aph);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + grp * ACC_COLS + ((uint32_t)(q * 32) << 16);
            const int mt_cnt = min(MT, tiles_m - mg * MT);
#pragma unroll 1
            for (int j = 0; j < mt_cnt; ++j) {
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_addr + (uint32_t)(j * BN + c0), r);
                    tmem_ld_wait();
                    if (j == mt_cnt - 1 && c0 + 32 >= BN) {   // last read of this stage: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[grp]);
                    }
                    const int nb = nt * BN + c0;
                    uint32_t w[32];
                    if (EPI == EPI_PARTIAL) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) w[i] = r[i];
                    } else {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 sc4 = *reinterpret_cast<const float4*>(s_scale + nb + i);
                            const float4 sh4 = *reinterpret_cast<const float4*>(s_shift + nb + i);
                            v[i] = fmaf(__uint_as_float(r[i]), sc4.x, sh4.x);
                            v[i + 1] = fmaf(__uint_as_float(r[i + 1]), sc4.y, sh4.y);
                            v[i + 2] = fmaf(__uint_as_float(r[i + 2]), sc4.z, sh4.z);
                            v[i + 3] = fmaf(__uint_as_float(r[i + 3]), sc4.w, sh4.w);
                        }
                        if (ACT == GACT_LEAKY) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);     // LeakyReLU(0.2) = max(x, 0.2x)
                        } else if (ACT == GACT_RELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (EPI == EPI_F32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(v[i]);
                        } else {                                 // row slice = [hi 64 B | lo 64 B]
#pragma unroll
                            for (int i = 0; i < 16; ++i) split_bf16x2(v[2 * i], v[2 * i + 1], w[i], w[16 + i]);
                        }
                    }
                    // park the row slice in the warp's swizzled staging tile, then store with 8 lanes per row
                    __syncwarp();
                    stage_write_row(wbuf, lane, w);
                    __syncwarp();
                    if (EPI == EPI_F32 && p.stats)
                        p.stats[((size_t)(mg * MT + j) * 4 + q) * p.N + nb + lane] =
                            stage_col_stats(wbuf, lane, min(32, max(0, p.M - ((mg * MT + j) * GEMM_BM + q * 32))));
                    const int row0 = (mg * MT + j) * GEMM_BM + q * 32 + sr;    // + 4·i per store instruction
                    const int rows_left = p.M - row0;                          // row 4i valid ⇔ 4i < rows_left
                    if (EPI == EPI_SPLIT) {
                        __nv_bfloat16* dst = (sc < 4 ? p.out_hi : p.out_lo) + (size_t)row0 * p.ldc + nb + (sc & 3) * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = 4 * i + sr;
                            const uint4 d = *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                            if (4 * i < rows_left) *reinterpret_cast<uint4*>(dst + (size_t)(4 * i) * p.ldc) = d;
                        }
                    } else {
                        const size_t off = (size_t)row0 * p.ldc + nb + sc * 4;
                        float* dst = p.out_f32 + (EPI == EPI_PARTIAL ? (size_t)sp * p.partial_stride : 0) + off;
                        const float* rs = (EPI == EPI_F32 && p.resid) ? p.resid + off : nullptr;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int rr = 4 * i + sr;
                            uint4 d = *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                            if (4 * i < rows_left) {
                                if (rs) {
                                    const float4 t = *reinterpret_cast<const float4*>(rs + (size_t)(4 * i) * p.ldc);
                                    d.x = __float_as_uint(__uint_as_float(d.x) + t.x); d.y = __float_as_uint(__uint_as_float(d.y) + t.y);
                                    d.z = __float_as_uint(__uint_as_float(d.z) + t.z); d.w = __float_as_uint(__uint_as_float(d.w) + t.w);
                                }
                                *reinterpret_cast<uint4*>(dst + (size_t)(4 * i) * p.ldc) = d;
                            }
                        }
                    }
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
