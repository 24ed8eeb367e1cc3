// tc_midgram.cuh — conv2 (1×1, 64→128) + BN + LeakyReLU + chan_proj (1×1, 128→96) + replicate-pad + the nine
// per-patch Gram matrices in ONE persistent tcgen05 kernel: neither the 128-channel activation nor the 96-channel
// chan_proj map (the "feature map" the Gram is taken of) ever leaves the SM.
// Reference: models_PMG.py:124-128 (final 1×1 of feature_extractor), :136,182 (chan_proj), :157-174 (pad 14→15,
//            3×3 disjoint 5×5 patches), :188 (G_k = P_k·P_kᵀ / 25).
//
// Work unit = one BAND of an image: the 5 image rows of one patch row py (70 pixels; 56 for py = 2, whose 5th row is
// the replicate padding), i.e. exactly the pixels of the three patches (py, 0..2).
//
//   act1 hi/lo rows [70·py, +72) ──TMA──► conv2 MMA (M128 N128, K 64, hi+lo) ──► TMEM D2 [pixel lanes][128 ch]
//   epilogue A (6 warps, thread = pixel): BN + LeakyReLU → hi/lo → smem A3, written at ROW px·32 + n of the patch-slot
//       order (n = 5·(y%5) + x%5; the replicate padding is one or three extra copies of the row; slots 25..31 stay zero)
//   chan MMA with the operand roles swapped (A = W3 [96 ch × 128], B = A3 [96 slots × 128], hi+lo):
//       TMEM D3ᵀ [channel lanes][slot columns] — chan_proj comes out transposed AND already split into patches
//   epilogue B (3 warps, thread = channel): 32 slot columns of one patch → hi/lo → one 128-byte row of the K-major,
//       128B-swizzled Gram operand P (hi in chunks 0-3, lo in chunks 4-7 of the same row)
//   Gram MMA (second issuer thread): 6 × (M128 N96 K16): Ph·Phᵀ + Ph·Plᵀ + Pl·Phᵀ ──► TMEM G (double-buffered)
//   epilogue C (9 warps, one per 32×32 block): ·1/25, hi/lo, the 6 upper blocks → G' row of token_proj (as tc_gram.cuh)
//
// TMEM (512 columns): D2 [0,128) | D3ᵀ 2 × 96 at 128, 224 | G 2 × 96 at 320, 416.
// Operand tiles whose row count is below the UMMA M = 128 (72-row act1 boxes, 96-row W3 / P) let the MMA read on into the
// neighbouring smem buffer: those accumulator lanes are never read.
// Algorithmic HBM bytes per image: 50,176 in (act1 hi/lo) + 221,184 out (G' hi/lo) — the separate mid + Gram kernels
// moved 50,176 + 75,264 (chan write) + 75,264 (chan read) + 221,184.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int MG_THREADS = 768;                   // 24 warps, roles below
constexpr int MG_STAGES = 3;
constexpr int MG_BOX_ROWS = 72;                   // TMA box: 70 band pixels rounded up to whole 8-row swizzle atoms
constexpr int MG_A_PLANE = MG_BOX_ROWS * 128;     // 9,216
constexpr int MG_A_STAGE = 2 * MG_A_PLANE;        // hi + lo
constexpr int MG_W2_BYTES = 128 * 128;            // [128 n × 64 k]
constexpr int MG_W3_BYTES = 2 * 96 * 128;         // 2 k-blocks × [96 ch × 64 k]
constexpr int MG_A3_BLK = 96 * 128;               // [96 slots × 64 k]
constexpr int MG_A3_BYTES = 4 * MG_A3_BLK;        // [plane][kb]
constexpr int MG_P_SLOT = 96 * 128;               // [96 ch × (32 px hi | 32 px lo)]
constexpr int MG_P_SLOTS = 4;
constexpr int MG_OFF_W3 = MG_W2_BYTES;
constexpr int MG_OFF_A = MG_OFF_W3 + MG_W3_BYTES;                     //  40,960
constexpr int MG_OFF_A3 = MG_OFF_A + MG_STAGES * MG_A_STAGE;          //  96,256
constexpr int MG_OFF_P = MG_OFF_A3 + MG_A3_BYTES;                     // 145,408
constexpr int MG_OFF_STG = MG_OFF_P + MG_P_SLOTS * MG_P_SLOT;         // 194,560: 6 × 4 KB epilogue-C staging
constexpr int MG_OFF_BAR = MG_OFF_STG + 6 * 4096;
constexpr int MG_OFF_PRM = MG_OFF_BAR + 256;
constexpr int MG_SMEM = MG_OFF_PRM + 2 * 128 * 4 + 1024;
static_assert(MG_SMEM <= 232448, "midgram kernel exceeds the shared-memory limit");
static_assert(MG_OFF_A % 1024 == 0 && MG_OFF_A3 % 1024 == 0 && MG_OFF_P % 1024 == 0 && MG_A_PLANE % 1024 == 0, "SW128 tiles need 1024-byte alignment");

struct MidGramParams {
    const float* scale2; const float* shift2;     // [128] BN(eval) of conv2 folded
    __nv_bfloat16* g_hi; __nv_bfloat16* g_lo;     // [B·9, 6144] upper Gram blocks (token_proj's operand)
    float* g_f32;                                 // optional [B·9, 96, 96]
    float* chan;                                  // optional [B·196, 96] fp32 (parity taps)
    int B;
};

__global__ void __launch_bounds__(MG_THREADS, 1)
midgram_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const MidGramParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW2 = smem;
    uint8_t* sW3 = smem + MG_OFF_W3;
    uint8_t* sA = smem + MG_OFF_A;
    uint8_t* sA3 = smem + MG_OFF_A3;
    uint8_t* sP = smem + MG_OFF_P;
    uint8_t* sStg = smem + MG_OFF_STG;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MG_OFF_BAR);
    uint64_t* a_full = bars;             // [3]
    uint64_t* a_empty = bars + 3;        // [3]
    uint64_t* d2_full = bars + 6;
    uint64_t* d2_empty = bars + 7;
    uint64_t* a3_ready = bars + 8;
    uint64_t* a3_free = bars + 9;
    uint64_t* d3_full = bars + 10;       // [2]
    uint64_t* d3_empty = bars + 12;      // [2]
    uint64_t* p_full = bars + 14;        // [4]
    uint64_t* p_empty = bars + 18;       // [4]
    uint64_t* g_full = bars + 22;        // [2]
    uint64_t* g_empty = bars + 24;       // [2]
    uint64_t* w_bar = bars + 26;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);
    float* s2 = reinterpret_cast<float*>(smem + MG_OFF_PRM);
    float* h2 = s2 + 128;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_img = p.B > (int)blockIdx.x ? (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_bands = 3 * n_img;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmW2); prefetch_tmap(&tmW3);
        for (int i = 0; i < MG_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(d2_full, 1); mbar_init(d2_empty, 6);
        mbar_init(a3_ready, 6); mbar_init(a3_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d3_full[i], 1); mbar_init(&d3_empty[i], 3);
            mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 9);
        }
        for (int i = 0; i < MG_P_SLOTS; ++i) { mbar_init(&p_full[i], 3); mbar_init(&p_empty[i], 1); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 15) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 128) { s2[threadIdx.x] = p.scale2[threadIdx.x]; h2[threadIdx.x] = p.shift2[threadIdx.x]; }
    // A3: the pad slots 25..31 of every patch must read as zero rows for the whole kernel (they become the zero
    // pixels 25..31 of P); P: clear once so the never-written lo/hi bytes of no slot hold NaN patterns
    for (int i = threadIdx.x; i < (MG_A3_BYTES + MG_P_SLOTS * MG_P_SLOT) / 16; i += MG_THREADS)
        reinterpret_cast<uint4*>(sA3)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD2 = tmem_base, tD3 = tmem_base + 128, tG = tmem_base + 320;
    pdl_trigger();

    if (warp == 3) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(w_bar, MG_W2_BYTES + MG_W3_BYTES);
            tma_load_2d_hint(sW2, &tmW2, w_bar, 0, 0, kEvictLast);
            tma_load_2d_hint(sW3, &tmW3, w_bar, 0, 0, kEvictLast);
            tma_load_2d_hint(sW3 + 96 * 128, &tmW3, w_bar, 64, 0, kEvictLast);
            pdl_wait();                  // act1 comes from the front kernel
            for (int i = 0; i < n_bands; ++i) {
                const int s = i % MG_STAGES;
                const int img = (int)blockIdx.x + (i / 3) * (int)gridDim.x, py = i % 3;
                mbar_wait(&a_empty[s], ((i / MG_STAGES) & 1) ^ 1);
                mbar_expect_tx(&a_full[s], MG_A_STAGE);
                tma_load_2d_hint(sA + s * MG_A_STAGE, &tmA_hi, &a_full[s], 0, img * 196 + 70 * py, kEvictFirst);
                tma_load_2d_hint(sA + s * MG_A_STAGE + MG_A_PLANE, &tmA_lo, &a_full[s], 0, img * 196 + 70 * py, kEvictFirst);
            }
        }
    } else if (warp == 7) {
        // ===================== MMA issuer 1: conv2 (band i) then chan_proj (band i-1) =====================
        if (elect_one()) {
            constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t id96 = make_idesc_bf16(128, 96, 0, 0);
            const uint32_t aW2 = smem_u32(sW2), aW3 = smem_u32(sW3), aA = smem_u32(sA), aA3 = smem_u32(sA3);
            mbar_wait(w_bar, 0);
            auto chan_mma = [&](uint32_t t) {                       // D3ᵀ[t&1] = W3 · A3(hi+lo)ᵀ for local band t
                const uint32_t b = t & 1;
                mbar_wait(a3_ready, t & 1);
                mbar_wait(&d3_empty[b], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t wd = make_desc_k_sw128(aW3 + kb * MG_A3_BLK + k * 32);
                        umma_bf16(tD3 + 96 * b, wd, make_desc_k_sw128(aA3 + kb * MG_A3_BLK + k * 32), id96, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(tD3 + 96 * b, wd, make_desc_k_sw128(aA3 + 2 * MG_A3_BLK + kb * MG_A3_BLK + k * 32), id96, 1u);
                    }
                umma_commit(a3_free);
                umma_commit(&d3_full[b]);
            };
            for (int i = 0; i < n_bands; ++i) {
                const uint32_t s = i % MG_STAGES;
                mbar_wait(d2_empty, (i & 1) ^ 1);
                mbar_wait(&a_full[s], (i / MG_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = aA + s * MG_A_STAGE, a_lo = a_hi + MG_A_PLANE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = make_desc_k_sw128(aW2 + k * 32);
                    umma_bf16(tD2, make_desc_k_sw128(a_hi + k * 32), bd, id128, k > 0 ? 1u : 0u);
                    umma_bf16(tD2, make_desc_k_sw128(a_lo + k * 32), bd, id128, 1u);
                }
                umma_commit(&a_empty[s]);
                umma_commit(d2_full);
                if (i >= 1) chan_mma(i - 1);
            }
            if (n_bands >= 1) chan_mma(n_bands - 1);
        }
    } else if (warp == 11) {
        // ===================== MMA issuer 2: Gram of patch t (slot t%4 of P) =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
            const uint32_t aP = smem_u32(sP);
            const int n_patches = 3 * n_bands;
            for (int t = 0; t < n_patches; ++t) {
                const uint32_t slot = t & 3, as = t & 1;
                mbar_wait(&g_empty[as], ((t >> 1) & 1) ^ 1);
                mbar_wait(&p_full[slot], (t >> 2) & 1);
                tc_fence_after();
                const uint32_t ph = aP + slot * MG_P_SLOT, pl = ph + 64;
                const uint32_t d = tG + 96 * as;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t dh = make_desc_k_sw128(ph + k * 32), dl = make_desc_k_sw128(pl + k * 32);
                    umma_bf16(d, dh, dh, idesc, k > 0 ? 1u : 0u);
                    umma_bf16(d, dh, dl, idesc, 1u);
                    umma_bf16(d, dl, dh, idesc, 1u);
                }
                umma_commit(&p_empty[slot]);
                umma_commit(&g_full[as]);
            }
        }
    } else if ((warp & 3) == 3) {
        // warps 15, 19, 23: no role (a warp with warp%4 == 3 can only reach TMEM lanes 96..127, which hold nothing here)
    } else if (warp < 8) {
        // ===================== epilogue A: conv2 → BN + LeakyReLU → hi/lo → A3 rows in patch-slot order =====================
        const int q = warp & 3, half = warp >> 2;             // lane quarter (pixels 32q..32q+31), 64-column half
        const int m = q * 32 + lane;                          // pixel of the band: image row 5py + m/14, column m%14
        const int r = m / 14, x = m % 14;
        const int R0 = (x / 5) * 32 + r * 5 + x % 5;          // A3 row: patch column px = x/5, slot n = 5r + x%5
        const bool dupx = x == 13;                            // column 13 is also the padded column 14 (slot + 1)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (int i = 0; i < n_bands; ++i) {
            const int py = i % 3;
            const bool valid = m < (py == 2 ? 56 : 70);
            const bool dupy = py == 2 && r == 3;              // image row 13 is also the padded row 14 (slot + 5)
            mbar_wait(d2_full, i & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cq = half * 2 + c;                  // 32-column chunk: K block cq/2 of A3, 16-byte chunks 4(cq%2)..+3
                uint32_t rr[32], H[16], L[16];
                tmem_ld_32x32b_x32(tD2 + lane_addr + (uint32_t)(cq * 32), rr);
                tmem_ld_wait();
                if (c == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d2_empty);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 sc = *reinterpret_cast<const float4*>(s2 + cq * 32 + 4 * j);
                    const float4 sh = *reinterpret_cast<const float4*>(h2 + cq * 32 + 4 * j);
                    float v0 = fmaf(__uint_as_float(rr[4 * j]), sc.x, sh.x), v1 = fmaf(__uint_as_float(rr[4 * j + 1]), sc.y, sh.y);
                    float v2 = fmaf(__uint_as_float(rr[4 * j + 2]), sc.z, sh.z), v3 = fmaf(__uint_as_float(rr[4 * j + 3]), sc.w, sh.w);
                    v0 = fmaxf(v0, 0.2f * v0); v1 = fmaxf(v1, 0.2f * v1); v2 = fmaxf(v2, 0.2f * v2); v3 = fmaxf(v3, 0.2f * v3);
                    split_bf16x2(v0, v1, H[2 * j], L[2 * j]);
                    split_bf16x2(v2, v3, H[2 * j + 1], L[2 * j + 1]);
                }
                if (c == 0 && i > 0) mbar_wait(a3_free, (i - 1) & 1);      // chan MMAs of the previous band have read A3
                if (valid) {
                    uint8_t* blk = sA3 + (cq >> 1) * MG_A3_BLK;
                    auto put = [&](int R) {
                        uint8_t* row = blk + R * 128;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int sw = ((4 * (cq & 1) + j) ^ (R & 7)) << 4;
                            *reinterpret_cast<uint4*>(row + sw) = make_uint4(H[4 * j], H[4 * j + 1], H[4 * j + 2], H[4 * j + 3]);
                            *reinterpret_cast<uint4*>(row + 2 * MG_A3_BLK + sw) = make_uint4(L[4 * j], L[4 * j + 1], L[4 * j + 2], L[4 * j + 3]);
                        }
                    };
                    put(R0);
                    if (dupx) put(R0 + 1);
                    if (dupy) put(R0 + 5);
                    if (dupx && dupy) put(R0 + 6);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a3_ready);
        }
        tc_fence_before();
    } else if (warp < 12) {
        // ===================== epilogue B: D3ᵀ (thread = channel, columns = the 32 slots of a patch) → hi/lo → P row =====================
        const int q = warp & 3;
        const int ch = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint8_t* prow = sP + ch * 128;
        for (int i = 0; i < n_bands; ++i) {
            const uint32_t b = i & 1;
            const int img = (int)blockIdx.x + (i / 3) * (int)gridDim.x, py = i % 3;
            mbar_wait(&d3_full[b], (i >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                const uint32_t t = 3 * i + j, slot = t & 3;
                uint32_t v[32], hi[16], lo[16];
                tmem_ld_32x32b_x32(tD3 + 96 * b + lane_addr + (uint32_t)(32 * j), v);
                tmem_ld_wait();
                if (j == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d3_empty[b]);
                }
                if (p.chan) {                                 // parity tap: the chan_proj map, pixel-major fp32
                    float* dst = p.chan + (size_t)img * (196 * 96) + ch;
#pragma unroll
                    for (int n = 0; n < 25; ++n) {
                        const int y = min(py * 5 + n / 5, 13), xx = min(j * 5 + n % 5, 13);
                        dst[(y * 14 + xx) * 96] = __uint_as_float(v[n]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) split_bf16x2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]), hi[e], lo[e]);
                mbar_wait(&p_empty[slot], ((t >> 2) & 1) ^ 1);
                uint8_t* row = prow + slot * MG_P_SLOT;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<uint4*>(row + ((c ^ (ch & 7)) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    *reinterpret_cast<uint4*>(row + (((4 + c) ^ (ch & 7)) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[slot]);
            }
        }
        tc_fence_before();
    } else {
        // ===================== epilogue C: Gram accumulator → ·1/25 → hi/lo → the 6 upper 32×32 blocks of G' =====================
        // warp ↔ block (q, cb): TMEM lane quarter q (rows 32q..32q+31) and 32-column chunk cb; block order in the
        // 6144-entry token row: (0,0) (1,1) (2,2) | (0,1) (0,2) (1,2), row-major inside a block (tc_gram.cuh,
        // pmg_fast.cu: pack_wt_folded_kernel)
        const int q = warp & 3, cb = (warp - 12) >> 2;
        const int c0 = cb * 32;
        const bool upper = cb >= q;
        const int blk = cb == q ? q : q + cb + 2;
        uint8_t* wbuf = sStg + (upper ? blk : 0) * 4096;
        const int sr = lane >> 3, sc = lane & 7;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int n_patches = 3 * n_bands;
        for (int t = 0; t < n_patches; ++t) {
            const uint32_t as = t & 1;
            const int i = t / 3, j = t % 3;
            const int tok = ((int)blockIdx.x + (i / 3) * (int)gridDim.x) * 9 + (i % 3) * 3 + j;
            mbar_wait(&g_full[as], (t >> 1) & 1);
            tc_fence_after();
            uint32_t r[32];
            if (upper || p.g_f32) {
                tmem_ld_32x32b_x32(tG + 96 * as + lane_addr + (uint32_t)c0, r);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[as]);
            if (upper) {
                uint32_t w[32];
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    split_bf16x2(__uint_as_float(r[2 * e]) * 0.04f, __uint_as_float(r[2 * e + 1]) * 0.04f, w[e], w[16 + e]);
                __syncwarp();
                stage_write_row(wbuf, lane, w);
                __syncwarp();
                const size_t tbase = (size_t)tok * 6144 + (size_t)blk * 1024 + (size_t)sr * 32 + (sc & 3) * 8;
                __nv_bfloat16* dst = (sc < 4 ? p.g_hi : p.g_lo) + tbase;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int rw = 4 * e + sr;
                    *reinterpret_cast<uint4*>(dst + (size_t)(4 * e) * 32) =
                        *reinterpret_cast<const uint4*>(wbuf + rw * 128 + ((sc ^ (rw & 7)) << 4));
                }
            }
            if (p.g_f32) {
                float4* df = reinterpret_cast<float4*>(p.g_f32 + (size_t)tok * 9216 + (size_t)(q * 32 + lane) * 96 + c0);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    df[e] = make_float4(__uint_as_float(r[4 * e]) * 0.04f, __uint_as_float(r[4 * e + 1]) * 0.04f,
                                        __uint_as_float(r[4 * e + 2]) * 0.04f, __uint_as_float(r[4 * e + 3]) * 0.04f);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 15) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
