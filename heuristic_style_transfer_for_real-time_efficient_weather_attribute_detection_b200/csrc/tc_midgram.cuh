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
//   epilogue A (12 warps, thread = pixel): BN + LeakyReLU → hi/lo → smem A3, written at ROW px·32 + n of the patch-slot
//       order (n = 5·(y%5) + x%5; the replicate padding is one or three extra copies of the row; slots 25..31 stay zero)
//   chan MMA with the operand roles swapped (A = W3 [96 ch × 128], B = A3 [96 slots × 128], hi+lo):
//       TMEM D3ᵀ [channel lanes][slot columns] — chan_proj comes out transposed AND already split into patches
//   epilogue B (3 warps, thread = channel): 32 slot columns of one patch → hi/lo → one 128-byte row of the K-major,
//       128B-swizzled Gram operand P (hi in chunks 0-3, lo in chunks 4-7 of the same row)
//   Gram MMA (second issuer thread): 6 × (M128 N96 K16): Ph·Phᵀ + Ph·Plᵀ + Pl·Phᵀ ──► TMEM G (double-buffered)
//   epilogue C (6 warps, one per upper 32×32 block): ·1/25, hi/lo → G' row of token_proj (layout of tc_gram.cuh),
//       straight from registers as 256-bit stores (thread = Gram row = 64 contiguous bytes per plane)
//
// Work balance over the four SM sub-partitions.  A warp can only read the TMEM lane quarter warp%4 — which is also its
// sub-partition — and every accumulator here has 96 real rows, so a straightforward mapping leaves one sub-partition
// idle and gives quarter 0 (Gram rows 0..31: three of the six upper blocks) twice the work of the others; the kernel
// then runs at the pace of that one instruction-issue port (role trace: tools/mg_trace.py).  So the unused M rows
// 96..127 of two operands carry copies: W3 rows 96..127 = channels 32..63 (D3ᵀ lanes 96..127 = channels 32..63: their
// epilogue B runs on quarter 3) and P rows 96..127 = channels 0..31 (G lanes 96..127 = Gram rows 0..31: blocks (0,1) and
// (0,2) drain on quarter 3).  Per quarter: 0 → EA×4, EB, EC(0,0) | 1 → EA×4, EC(1,1), EC(1,2) | 2 → EA×4, EB, EC(2,2) |
// 3 → EB, EC(0,1), EC(0,2); the three single-thread roles (TMA, two MMA issuers) ride on quarters 0..2.
//
// TMEM (512 columns): D2 [0,128) | D3ᵀ 2 × 96 at 128, 224 | G 2 × 96 at 320, 416.
// The 72-row act1 boxes are shorter than the UMMA M = 128: the MMA reads on into the neighbouring smem buffer and
// those accumulator lanes are never read.
// Algorithmic HBM bytes per image: 50,176 in (act1 hi/lo) + 221,184 out (G' hi/lo) — the separate mid + Gram kernels
// moved 50,176 + 75,264 (chan write) + 75,264 (chan read) + 221,184.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int MG_THREADS = 896;                   // 28 warps: warp w ↔ (quarter w%4, slot w/4), roles in mg_role()
constexpr int MG_STAGES = 2;
constexpr int MG_BOX_ROWS = 72;                   // TMA box: 70 band pixels rounded up to whole 8-row swizzle atoms
constexpr int MG_A_PLANE = MG_BOX_ROWS * 128;     // 9,216
constexpr int MG_A_STAGE = 2 * MG_A_PLANE;        // hi + lo
constexpr int MG_W2_BYTES = 128 * 128;            // [128 n × 64 k]
constexpr int MG_W3_BLK = 128 * 128;              // one k-block: [96 ch + copy of ch 32..63 × 64 k]
constexpr int MG_W3_BYTES = 2 * MG_W3_BLK;
constexpr int MG_A3_BLK = 96 * 128;               // [96 slots × 64 k]
constexpr int MG_A3_BYTES = 4 * MG_A3_BLK;        // [plane][kb]; two such buffers (bands alternate)
constexpr int MG_P_SLOT = 128 * 128;              // [96 ch + copy of ch 0..31 × (32 px hi | 32 px lo)]
constexpr int MG_P_SLOTS = 2;
constexpr int MG_OFF_W3 = MG_W2_BYTES;
constexpr int MG_OFF_A = MG_OFF_W3 + MG_W3_BYTES;                     //  49,152
constexpr int MG_OFF_A3 = MG_OFF_A + MG_STAGES * MG_A_STAGE;          //  86,016
constexpr int MG_OFF_P = MG_OFF_A3 + 2 * MG_A3_BYTES;                 // 184,320
constexpr int MG_OFF_BAR = MG_OFF_P + MG_P_SLOTS * MG_P_SLOT;         // 217,088
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
    long long* trace;                             // optional debug stamps (pmg_debug_set_trace), slots 384..503: bands 6..8 of CTA 0
};
#define MG_T(band, slot) do { if (p.trace && blockIdx.x == 0 && (band) >= 6 && (band) < 9) p.trace[384 + ((band) - 6) * 40 + (slot)] = clock64(); } while (0)

enum : int { MG_IDLE = 0, MG_TMA, MG_MMA1, MG_MMA2, MG_EA, MG_EB, MG_EC };
// role of warp (quarter q, slot k) and its argument: EA → 32-column chunk, EC → column block cb (row block = q, or 0 on quarter 3)
__device__ __forceinline__ int mg_role(int q, int k, int& arg) {
    arg = 0;
    if (q < 3 && k < 4) { arg = k; return MG_EA; }
    if (q == 0) return k == 4 ? MG_EB : k == 5 ? MG_EC : MG_TMA;                                  // EC (0,0)
    if (q == 1) { arg = k == 4 ? 1 : 2; return k < 6 ? MG_EC : MG_MMA1; }                        // EC (1,1), (1,2)
    if (q == 2) { arg = 2; return k == 4 ? MG_EB : k == 5 ? MG_EC : MG_MMA2; }                   // EC (2,2)
    if (k == 0) return MG_EB;
    if (k <= 2) { arg = k; return MG_EC; }                                                        // EC (0,1), (0,2) through the row copy
    return MG_IDLE;
}

__global__ void __launch_bounds__(MG_THREADS, 1)
midgram_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
               const __grid_constant__ CUtensorMap tmW3d, const MidGramParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW2 = smem;
    uint8_t* sW3 = smem + MG_OFF_W3;
    uint8_t* sA = smem + MG_OFF_A;
    uint8_t* sA3 = smem + MG_OFF_A3;
    uint8_t* sP = smem + MG_OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MG_OFF_BAR);
    uint64_t* a_full = bars;             // [2]
    uint64_t* a_empty = bars + 3;        // [2]
    uint64_t* d2_full = bars + 6;
    uint64_t* d2_empty = bars + 7;
    uint64_t* a3_ready = bars + 8;       // [2]
    uint64_t* a3_free = bars + 28;       // [2]
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
    const int q = warp & 3;
    int role_arg;
    const int role = mg_role(q, warp >> 2, role_arg);
    // bands (image, py) are dealt round-robin: global band gb = blockIdx.x + i·gridDim.x, so a small batch spreads the three
    // bands of an image over three CTAs instead of chaining them on one
    const int total_bands = 3 * p.B;
    const int n_bands = total_bands > (int)blockIdx.x ? (total_bands - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;     // this warp's TMEM lane quarter

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmW2); prefetch_tmap(&tmW3); prefetch_tmap(&tmW3d);
        for (int i = 0; i < MG_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        mbar_init(d2_full, 1); mbar_init(d2_empty, 12);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a3_ready[i], 12); mbar_init(&a3_free[i], 1);
            mbar_init(&d3_full[i], 1); mbar_init(&d3_empty[i], 3);
            mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 6);
        }
        for (int i = 0; i < MG_P_SLOTS; ++i) { mbar_init(&p_full[i], 3); mbar_init(&p_empty[i], 1); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 27) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 128) { s2[threadIdx.x] = p.scale2[threadIdx.x]; h2[threadIdx.x] = p.shift2[threadIdx.x]; }
    // A3: the pad slots 25..31 of every patch must read as zero rows for the whole kernel (they become the zero
    // pixels 25..31 of P); P: clear once so the never-written lo/hi bytes of no slot hold NaN patterns
    for (int i = threadIdx.x; i < (2 * MG_A3_BYTES + MG_P_SLOTS * MG_P_SLOT) / 16; i += MG_THREADS)
        reinterpret_cast<uint4*>(sA3)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD2 = tmem_base, tD3 = tmem_base + 128, tG = tmem_base + 320;
    pdl_trigger();

    if (role == MG_TMA) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(w_bar, MG_W2_BYTES + MG_W3_BYTES);
            tma_load_2d_hint(sW2, &tmW2, w_bar, 0, 0, kEvictLast);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d_hint(sW3 + kb * MG_W3_BLK, &tmW3, w_bar, 64 * kb, 0, kEvictLast);                 // channels 0..95
                tma_load_2d_hint(sW3 + kb * MG_W3_BLK + 96 * 128, &tmW3d, w_bar, 64 * kb, 32, kEvictLast);    // rows 96..127 = channels 32..63
            }
            pdl_wait();                  // act1 comes from the front kernel
            for (int i = 0; i < n_bands; ++i) {
                const int s = i % MG_STAGES;
                const int gb = (int)blockIdx.x + i * (int)gridDim.x, img = gb / 3, py = gb - 3 * img;
                mbar_wait(&a_empty[s], ((i / MG_STAGES) & 1) ^ 1);
                MG_T(i, 0);
                mbar_expect_tx(&a_full[s], MG_A_STAGE);
                tma_load_2d_hint(sA + s * MG_A_STAGE, &tmA_hi, &a_full[s], 0, img * 196 + 70 * py, kEvictFirst);
                tma_load_2d_hint(sA + s * MG_A_STAGE + MG_A_PLANE, &tmA_lo, &a_full[s], 0, img * 196 + 70 * py, kEvictFirst);
            }
        }
    } else if (role == MG_MMA1) {
        // ===================== MMA issuer 1: conv2 (band i) then chan_proj (band i-1) =====================
        if (elect_one()) {
            constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t id96 = make_idesc_bf16(128, 96, 0, 0);
            const uint32_t aW2 = smem_u32(sW2), aW3 = smem_u32(sW3), aA = smem_u32(sA), aA3 = smem_u32(sA3);
            mbar_wait(w_bar, 0);
            auto chan_mma = [&](uint32_t t) {                       // D3ᵀ[t&1] = W3 · A3(hi+lo)ᵀ for local band t
                const uint32_t b = t & 1, a3 = aA3 + b * MG_A3_BYTES;
                mbar_wait(&a3_ready[b], (t >> 1) & 1);
                mbar_wait(&d3_empty[b], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                MG_T((int)t, 3);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t wd = make_desc_k_sw128(aW3 + kb * MG_W3_BLK + k * 32);
                        umma_bf16(tD3 + 96 * b, wd, make_desc_k_sw128(a3 + kb * MG_A3_BLK + k * 32), id96, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(tD3 + 96 * b, wd, make_desc_k_sw128(a3 + 2 * MG_A3_BLK + kb * MG_A3_BLK + k * 32), id96, 1u);
                    }
                umma_commit(&a3_free[b]);
                umma_commit(&d3_full[b]);
                MG_T((int)t, 4);
            };
            for (int i = 0; i < n_bands; ++i) {
                const uint32_t s = i % MG_STAGES;
                mbar_wait(d2_empty, (i & 1) ^ 1);
                mbar_wait(&a_full[s], (i / MG_STAGES) & 1);
                tc_fence_after();
                MG_T(i, 1);
                const uint32_t a_hi = aA + s * MG_A_STAGE, a_lo = a_hi + MG_A_PLANE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = make_desc_k_sw128(aW2 + k * 32);
                    umma_bf16(tD2, make_desc_k_sw128(a_hi + k * 32), bd, id128, k > 0 ? 1u : 0u);
                    umma_bf16(tD2, make_desc_k_sw128(a_lo + k * 32), bd, id128, 1u);
                }
                umma_commit(&a_empty[s]);
                umma_commit(d2_full);
                MG_T(i, 2);
                if (i >= 1) chan_mma(i - 1);
            }
            if (n_bands >= 1) chan_mma(n_bands - 1);
        }
    } else if (role == MG_MMA2) {
        // ===================== MMA issuer 2: Gram of patch t (slot t%4 of P) =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
            const uint32_t aP = smem_u32(sP);
            const int n_patches = 3 * n_bands;
            for (int t = 0; t < n_patches; ++t) {
                const uint32_t slot = t & 1, as = t & 1;
                mbar_wait(&g_empty[as], ((t >> 1) & 1) ^ 1);
                mbar_wait(&p_full[slot], (t >> 1) & 1);
                tc_fence_after();
                MG_T(t / 3, 19 + 2 * (t % 3));
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
                MG_T(t / 3, 20 + 2 * (t % 3));
            }
        }
    } else if (role == MG_EA) {
        // ===================== epilogue A: conv2 → BN + LeakyReLU → hi/lo → A3 rows in patch-slot order =====================
        // 12 warps (3 lane quarters × 4 column chunks).  A3 is double-buffered, so this band's arithmetic and stores run
        // under the previous band's chan MMAs
        const int cq = role_arg;                              // 32-column chunk: K block cq/2 of A3, 16-byte chunks 4(cq%2)..+3
        const int m = q * 32 + lane;                          // pixel of the band: image row 5py + m/14, column m%14
        const int r = m / 14, x = m % 14;
        const int R0 = (x / 5) * 32 + r * 5 + x % 5;          // A3 row: patch column px = x/5, slot n = 5r + x%5
        const bool dupx = x == 13;                            // column 13 is also the padded column 14 (slot + 1)
        for (int i = 0; i < n_bands; ++i) {
            const int py = ((int)blockIdx.x + i * (int)gridDim.x) % 3;
            const int nrows = py == 2 ? 56 : 70;
            const bool valid = m < nrows;
            const bool dupy = py == 2 && r == 3;              // image row 13 is also the padded row 14 (slot + 5)
            mbar_wait(d2_full, i & 1);
            tc_fence_after();
            if (threadIdx.x == 0) MG_T(i, 5);
            if (q * 32 >= nrows) {
                // none of this warp's 32 pixel rows exists in the band (quarter 2 of the short band): keep the barrier
                // protocol (both arrivals after the phases they belong to) and skip the work
                __syncwarp();
                if (lane == 0) mbar_arrive(d2_empty);
                mbar_wait(&a3_free[i & 1], ((i >> 1) & 1) ^ 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&a3_ready[i & 1]);
                continue;
            }
            uint32_t rr[32], H[16], L[16];
            tmem_ld_32x32b_x32(tD2 + lane_addr + (uint32_t)(cq * 32), rr);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 sc = *reinterpret_cast<const float4*>(s2 + cq * 32 + 4 * j);
                const float4 sh = *reinterpret_cast<const float4*>(h2 + cq * 32 + 4 * j);
                bn_lrelu_split_x2(rr[4 * j], rr[4 * j + 1], sc.x, sc.y, sh.x, sh.y, H[2 * j], L[2 * j]);
                bn_lrelu_split_x2(rr[4 * j + 2], rr[4 * j + 3], sc.z, sc.w, sh.z, sh.w, H[2 * j + 1], L[2 * j + 1]);
            }
            if (threadIdx.x == 0) MG_T(i, 6);
            mbar_wait(&a3_free[i & 1], ((i >> 1) & 1) ^ 1);   // the chan MMAs of band i-2 have read this A3 buffer
            if (threadIdx.x == 0) MG_T(i, 7);
            if (valid) {
                uint8_t* blk = sA3 + (i & 1) * MG_A3_BYTES + (cq >> 1) * MG_A3_BLK;
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
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a3_ready[i & 1]);
            if (threadIdx.x == 0) MG_T(i, 8);
        }
        tc_fence_before();
    } else if (role == MG_EB) {
        // ===================== epilogue B: D3ᵀ (thread = channel, columns = the 32 slots of a patch) → hi/lo → P row =====================
        // quarter 3 sees the copy of channels 32..63 in lanes 96..127; quarter 0 also writes the copy of channels 0..31
        // that puts Gram rows 0..31 into lanes 96..127 of G
        const int ch = (q == 3 ? 32 : q * 32) + lane;
        const bool tracer = q == 0 && lane == 0;
        uint8_t* prow = sP + ch * 128;
        for (int i = 0; i < n_bands; ++i) {
            const uint32_t b = i & 1;
            const int gb = (int)blockIdx.x + i * (int)gridDim.x, img = gb / 3, py = gb - 3 * img;
            mbar_wait(&d3_full[b], (i >> 1) & 1);
            tc_fence_after();
            if (tracer) MG_T(i, 9);
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                const uint32_t t = 3 * i + j, slot = t & 1;
                uint32_t v[32], hi[16], lo[16];
                tmem_ld_32x32b_x32(tD3 + 96 * b + lane_addr + (uint32_t)(32 * j), v);
                tmem_ld_wait();
                if (tracer) MG_T(i, 10 + 3 * j);
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
                mbar_wait(&p_empty[slot], ((t >> 1) & 1) ^ 1);
                if (tracer) MG_T(i, 11 + 3 * j);
                uint8_t* row = prow + slot * MG_P_SLOT;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 h4 = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    const uint4 l4 = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                    const int oh = (c ^ (ch & 7)) << 4, ol = ((4 + c) ^ (ch & 7)) << 4;
                    *reinterpret_cast<uint4*>(row + oh) = h4;
                    *reinterpret_cast<uint4*>(row + ol) = l4;
                    if (q == 0) {                             // rows 96..127 (96 ≡ 0 mod 8: same swizzle phase)
                        *reinterpret_cast<uint4*>(row + 96 * 128 + oh) = h4;
                        *reinterpret_cast<uint4*>(row + 96 * 128 + ol) = l4;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[slot]);
                if (tracer) MG_T(i, 12 + 3 * j);
            }
        }
        tc_fence_before();
    } else if (role == MG_EC) {
        // ===================== epilogue C: Gram accumulator → ·1/25 → hi/lo → one upper 32×32 block of G' =====================
        // block (rb, cb): Gram rows 32rb.. (TMEM lane quarter rb, or quarter 3 = the copy of rows 0..31) and columns 32cb..;
        // block order in the 6144-entry token row: (0,0) (1,1) (2,2) | (0,1) (0,2) (1,2), row-major inside a block
        // (tc_gram.cuh, pmg_fast.cu: pack_wt_folded_kernel)
        const int rb = q == 3 ? 0 : q, cb = role_arg;
        const int c0 = cb * 32;
        const int blk = cb == rb ? rb : rb + cb + 2;
        const bool tracer = q == 0 && lane == 0;                // EC (0,0)
        int i = 0, j = 0;
        const int n_patches = 3 * n_bands;
        for (int t = 0; t < n_patches; ++t) {
            const uint32_t as = t & 1;
            const int tok = ((int)blockIdx.x + i * (int)gridDim.x) * 3 + j;      // (img·3 + py)·3 + px
            mbar_wait(&g_full[as], (t >> 1) & 1);
            tc_fence_after();
            if (tracer) MG_T(i, 25 + 3 * j);
            uint32_t r[32];
            tmem_ld_32x32b_x32(tG + 96 * as + lane_addr + (uint32_t)c0, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[as]);
            if (tracer) MG_T(i, 26 + 3 * j);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                float g0, g1;
                unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), pack_f32x2(0.04f, 0.04f)), g0, g1);
                r[2 * e] = __float_as_uint(g0); r[2 * e + 1] = __float_as_uint(g1);
            }
            if (p.g_f32) {                                    // full fp32 Gram (tests, forward_with_gram): the block and its mirror image
                float* g = p.g_f32 + (size_t)tok * 9216;
                float4* df = reinterpret_cast<float4*>(g + (size_t)(rb * 32 + lane) * 96 + c0);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    df[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
                if (cb != rb) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) g[(size_t)(c0 + e) * 96 + rb * 32 + lane] = __uint_as_float(r[e]);
                }
            }
            {
                // thread = Gram row: its 32 columns are 64 contiguous bytes in each plane and the warp's 32 rows are
                // contiguous (block is row-major), so four 256-bit stores per thread write whole 32-byte sectors
                uint32_t w[32];
#pragma unroll
                for (int e = 0; e < 16; ++e) split_bf16x2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]), w[e], w[16 + e]);
                const size_t off = (size_t)tok * 6144 + (size_t)blk * 1024 + (size_t)lane * 32;
                st_global_v8(p.g_hi + off, w);
                st_global_v8(p.g_hi + off + 16, w + 8);
                st_global_v8(p.g_lo + off, w + 16);
                st_global_v8(p.g_lo + off + 16, w + 24);
            }
            if (tracer) MG_T(i, 27 + 3 * j);
            if (++j == 3) { j = 0; ++i; }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 27) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace tc
