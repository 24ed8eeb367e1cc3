// tc_front.cuh — fused PatchGAN front end: conv0 (k4 s4, 3→32) + BN + LeakyReLU + conv1 (k4 s4, 32→64) + BN +
// LeakyReLU in ONE persistent tcgen05 kernel; the 56×56×32 activation never leaves the SM.
// Reference: models_PMG.py:113-121 (the two strided layers of feature_extractor).
//
// Tile = 98 conv1 pixels (7 rows × 14 cols of the 14×14 map) of one image, 2 tiles per image; UMMA M = 128
// (rows 98..127 are don't-care).  For each of the 4 row offsets sy of conv1's 4×4 window:
//
//   * TMA (4-D tensor map over the NCHW bf16 image, dims x|py|y16|b·c, box 224×8×4×1, no swizzle) drops, per
//     channel c, four slabs (kernel rows ky) of 8 whole image rows each straight into smem: no im2col, no register
//     staging, 448-byte runs, ONE bulk tensor copy per (sy, c) step (the 8th image row is the neighbouring tile's /
//     zero fill: don't-care MMA rows).  fp32 input arrives as two bf16 planes (hi, lo): 3 more steps per sy against
//     the same weights (template parameter NPLANES).
//   * conv0 as a GEMM whose M rows are the 16-BYTE UNITS of a slab (8 image columns = two conv0 pixels xp = 0, 1 of
//     one conv1 pixel; 28 units per image row, 196 per slab = 2 M tiles): the slab IS a K-major operand in the
//     no-swizzle core-matrix layout (8 units = 128 contiguous bytes, SBO = 128 B) and the two 16-byte K chunks of a
//     K = 16 step are the same 8 columns in slabs ky and ky+1 (LBO = slab pitch):
//         D0[unit, (xp, co)] = Σ_{c, ky} A[unit, 8 cols] · W0u[(xp, co), (c, ky, 8 cols)]
//     W0u is block-sparse over xp (each pixel sees its 4 of the 8 columns): N = 64, 6 K steps × 2 M tiles = 12 MMAs
//     (M128 N64 K16) per sy, 12 KB of weights.  (The first formulation used 32-byte rows = one conv1 pixel's 16
//     columns, N = 128 block-sparse over 4 pixels: twice the MMA time, 48 KB of weights, 33 % more operand bytes.)
//   * epilogue 0 (16 warps): TMEM → BN scale/shift → LeakyReLU → hi/lo bf16 split → written as conv1's A operand
//     (K-major, 128-byte swizzle): unit u ↔ conv1 pixel u>>1, K block u&1, 64-byte half xp.
//   * conv1 accumulates that K slice:  D1[pixel, 64] += A1(hi) · W1[sy] + A1(lo) · W1[sy]   (16 MMAs, M128 N64 K16).
//     A1 is DOUBLE-BUFFERED (sub-tile pitch 104 rows: the shared memory the smaller W0u freed), so epilogue 0 writes
//     slice sy+1 while conv1 reads slice sy.
//
// After the 4th slice, epilogue 1 (its own 4 warps, D1 double-buffered) applies BN + LeakyReLU and stores act1 as hi/lo
// bf16 planes [B·196, 64] while the other warps already work on the next tile.  conv0 and conv1 MMAs are issued by two
// different threads, so neither stream is blocked behind a barrier wait of the other.  Both weight tensors stay
// resident in smem for the life of the CTA; D0 is double-buffered in TMEM.
// What bounds it (role trace + tools/load_probe.py): shared-memory bandwidth — per slice the MMAs read 72 + 96 KB of
// operands, epilogue 0 writes 50 KB, TMA lands 42 KB; the image stream alone runs at 6.4 TB/s with the same boxes.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int FRONT_THREADS = 736;                 // warp 0 TMA, 1 conv0 MMA issuer, 2-17 epilogue 0 (lane quarter × 32-column group), 18 conv1 MMA issuer, 19-22 epilogue 1
constexpr int FR_W0_KSTEP = 2 * 64 * 16;           // one conv0 K step (c, ky pair): 2 K chunks × [64 rows × 16 B] as 8×16-byte core matrices
constexpr int FR_W0_BYTES = 6 * FR_W0_KSTEP;       // 12 KB (the block-sparse [128 × 192] formulation took 48 KB)
constexpr int FR_W1_BYTES = 8 * 64 * 128;          // 64 KB: (sy, kb) × [64 rows × 128 B]
constexpr int FR_A0_SLOT = 3584;                   // one kernel row ky: 8 image rows × 448 B (7 used by the tile; 256-byte multiple)
constexpr int FR_A0_STAGE = 4 * FR_A0_SLOT;        // one channel: 4 kernel rows = ONE TMA box
constexpr int FR_A0_STAGES = 3;                    // (6 stages measured: no gain — with no compute the same boxes stream at 6.4 TB/s, 22 B/clk
                                                   // per SM, tools/load_probe.py; inside the kernel the loads are slowed by the shared-memory
                                                   // traffic of the MMAs and epilogue 0, not by ring depth)
constexpr int FR_A1_SUB = 104 * 128;               // one [rows × 64 K] sub-tile of conv1's A operand: 98 real rows, pitch 13 × 1024 B; the
                                                   // M = 128 MMA reads on into the next sub-tile — accumulator rows nobody looks at
constexpr int FR_A1_BUF = 4 * FR_A1_SUB;           // [plane hi/lo][kb]
constexpr int FR_A1_BYTES = 2 * FR_A1_BUF;         // double-buffered: epilogue 0 fills slice sy+1 while conv1 reads slice sy
constexpr int FR_BOX_BYTES = FR_A0_STAGE;        // box 224 × 8 × 4 × 1 bf16 (out-of-range image rows are zero-filled but counted)
constexpr int FR_OFF_W1 = FR_W0_BYTES;
constexpr int FR_OFF_A0 = FR_OFF_W1 + FR_W1_BYTES;
constexpr int FR_OFF_A1 = FR_OFF_A0 + FR_A0_STAGES * FR_A0_STAGE;
constexpr int FR_OFF_BAR = FR_OFF_A1 + FR_A1_BYTES;
constexpr int FR_OFF_PRM = FR_OFF_BAR + 384;
// the M = 128 MMA on the last A1 sub-tile reads 3072 bytes past the end of A1: barriers and parameters live there, the rest is padding
constexpr int FRONT_SMEM = FR_OFF_BAR + 3072 + 1024;
static_assert(FR_OFF_PRM + (32 + 32 + 64 + 64) * 4 <= FR_OFF_BAR + 3072, "barriers + parameters must fit the over-read pad");
static_assert(FRONT_SMEM <= 232448, "front kernel exceeds the 227 KB shared-memory limit");
static_assert(FR_OFF_W1 % 1024 == 0 && FR_OFF_A0 % 1024 == 0 && FR_OFF_A1 % 1024 == 0, "SW128 tiles need 1024-byte alignment");

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "l"(kEvictFirst)
        : "memory");
}

// L2 prefetch of one image-row box: the smem ring holds ~38 KB per SM, far less than HBM latency × bandwidth needs
// in flight (≈ 90 KB per SM), so the producer asks L2 for the boxes of the NEXT tile while it loads this one
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// K-major operand WITHOUT swizzle: 8 rows × 16 bytes core matrices (128 contiguous bytes); `sbo` = distance between
// consecutive 8-row groups, `lbo` = distance between the two 16-byte K chunks of one K = 16 step
__device__ __forceinline__ uint64_t make_desc_k_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) { return make_smem_desc(saddr, lbo, sbo, LAYOUT_NONE); }
__device__ __forceinline__ void fr_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FrontParams {
    const float* scale0; const float* shift0;    // [32]  BN(eval) of conv0 folded
    const float* scale1; const float* shift1;    // [64]
    const uint8_t* w0u;                          // conv0 weights in the core-matrix layout of pack_w0u_kernel (FR_W0_BYTES)
    __nv_bfloat16* act1_hi; __nv_bfloat16* act1_lo;   // [B·196, 64]
    int B;
    // template parameter NPLANES — 1: bf16 image; 2: fp32 image carried as bf16 hi + lo planes (tmXlo): conv0 is linear in the
    // image, so the lo plane is just three more "channels" against the same weights
    long long* trace;                            // optional debug stamps (pmg_debug_set_trace), slots 256..383
};
#define FR_T(slot) do { if (p.trace && blockIdx.x == 0 && lt == 2) p.trace[256 + (slot)] = clock64(); } while (0)

template <int NPLANES>
__global__ void __launch_bounds__(FRONT_THREADS, 1)
front_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXlo,
             const __grid_constant__ CUtensorMap tmW1, const FrontParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW0 = smem;
    uint8_t* sW1 = smem + FR_OFF_W1;
    uint8_t* sA0 = smem + FR_OFF_A0;
    uint8_t* sA1 = smem + FR_OFF_A1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FR_OFF_BAR);
    uint64_t* a0_full = bars + 24;       // [FR_A0_STAGES]
    uint64_t* a0_empty = bars + 32;      // [FR_A0_STAGES]
    uint64_t* d0_full = bars + 6;        // [2]
    uint64_t* d0_empty = bars + 8;       // [2]
    uint64_t* a1_ready = bars + 10;      // [2] A1 buffer written (both planes)
    uint64_t* a1_free = bars + 18;       // [2] A1 buffer read by conv1
    uint64_t* d1_full = bars + 12;       // [2]
    uint64_t* d1_empty = bars + 14;      // [2]
    uint64_t* w_bar = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    float* s0 = reinterpret_cast<float*>(smem + FR_OFF_PRM);
    float* h0 = s0 + 32;
    float* s1 = h0 + 32;
    float* h1 = s1 + 64;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int n_tiles = 2 * p.B;
    constexpr int nc = 3 * NPLANES;                // (plane, channel) steps per conv1 kernel row: compile-time — the producer thread
                                                   // sits on the critical path of the 3-stage ring, and divisions by a run-time nc
                                                   // in its step loop cost the bf16 path 9 % of the kernel (0.109 -> 0.118 ms)

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmXlo); prefetch_tmap(&tmW1);
        for (int i = 0; i < FR_A0_STAGES; ++i) { mbar_init(&a0_full[i], 1); mbar_init(&a0_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d0_full[i], 1); mbar_init(&d0_empty[i], 16); }
        for (int i = 0; i < 2; ++i) { mbar_init(&a1_ready[i], 16); mbar_init(&a1_free[i], 1); }
        mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 32) { s0[threadIdx.x] = p.scale0[threadIdx.x]; h0[threadIdx.x] = p.shift0[threadIdx.x]; }
    if (threadIdx.x >= 64 && threadIdx.x < 128) { s1[threadIdx.x - 64] = p.scale1[threadIdx.x - 64]; h1[threadIdx.x - 64] = p.shift1[threadIdx.x - 64]; }
    // A0 slot tails (rows 98..127) and A1 are read by the MMAs before/without ever being written for the
    // don't-care rows: clear once so no NaN bit patterns circulate
    for (int i = threadIdx.x; i < (FR_A0_STAGES * FR_A0_STAGE + FR_A1_BYTES) / 16; i += FRONT_THREADS)
        reinterpret_cast<uint4*>(sA0)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD0 = tmem_base, tD1 = tmem_base + 256;
    pdl_trigger();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas (no per-MMA ELECT loops)
            mbar_expect_tx(w_bar, FR_W0_BYTES + FR_W1_BYTES);
            fr_bulk_g2s(sW0, p.w0u, FR_W0_BYTES, w_bar);
            for (int i = 0; i < 8; ++i) tma_load_2d(sW1 + i * 8192, &tmW1, w_bar, 0, i * 64);
            pdl_wait();      // weights are constants; everything below may depend on earlier kernels
            constexpr uint32_t PF = 4 * nc;              // prefetch distance in (sy, c) steps = one tile
            auto prefetch = [&](uint32_t j) {             // j-th (tile, sy, c) step of this CTA
                const int tile = blockIdx.x + (int)(j / PF) * gridDim.x;
                if (tile >= n_tiles) return;
                const int r = j % PF, sy = r / nc, c = r - nc * sy;
                tma_prefetch_4d(c >= 3 ? &tmXlo : &tmX, 0, (tile & 1) * 7, sy * 4, (tile >> 1) * 3 + (c >= 3 ? c - 3 : c));
            };
            for (uint32_t j = 0; j < PF; ++j) prefetch(j);
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const int b = tile >> 1, py0 = (tile & 1) * 7;
                for (int sy = 0; sy < 4; ++sy)
                    for (int c = 0; c < nc; ++c, ++it) {
                        const int s = it % FR_A0_STAGES;
                        const uint32_t ph = (it / FR_A0_STAGES) & 1;
                        prefetch(it + PF);
                        mbar_wait(&a0_empty[s], ph ^ 1);
                        if (c < 3) FR_T(sy * 3 + c);
                        mbar_expect_tx(&a0_full[s], FR_BOX_BYTES);
                        tma_load_4d(sA0 + s * FR_A0_STAGE, c >= 3 ? &tmXlo : &tmX, &a0_full[s], 0, py0, sy * 4, b * 3 + (c >= 3 ? c - 3 : c));
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== conv0 MMA issuer =====================
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas (no per-MMA ELECT loops)
            constexpr uint32_t idesc0 = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t aW0 = smem_u32(sW0), aA0 = smem_u32(sA0);
            mbar_wait(w_bar, 0);
            uint32_t it = 0, g = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                for (int sy = 0; sy < 4; ++sy, ++g) {
                    const uint32_t as = g & 1, aph = (g >> 1) & 1;
                    mbar_wait(&d0_empty[as], aph ^ 1);
                    tc_fence_after();
                    FR_T(20 + 4 * sy);
                    for (int c = 0; c < nc; ++c, ++it) {
                        const int s = it % FR_A0_STAGES;
                        const uint32_t ph = (it / FR_A0_STAGES) & 1;
                        mbar_wait(&a0_full[s], ph);
                        tc_fence_after();
                        if (c == nc - 1) FR_T(21 + 4 * sy);
                        // K step = (channel, kernel-row pair): the two 16-byte K chunks of an M row are the same 8 image columns
                        // in slabs ky and ky+1 (LBO = slab pitch); M tile t = 16-byte units 128t.. of the slab (SBO = 128 B)
                        const uint32_t wj = aW0 + (c >= 3 ? c - 3 : c) * 2 * FR_W0_KSTEP;      // the lo plane meets the same weights as the hi plane
#pragma unroll
                        for (int kp = 0; kp < 2; ++kp)
#pragma unroll
                            for (int t = 0; t < 2; ++t)
                                umma_bf16(tD0 + as * 128 + t * 64,
                                          make_desc_k_none(aA0 + s * FR_A0_STAGE + 2 * kp * FR_A0_SLOT + t * 2048, FR_A0_SLOT, 128),
                                          make_desc_k_none(wj + kp * FR_W0_KSTEP, 1024, 128), idesc0, (c > 0 || kp > 0) ? 1u : 0u);
                        umma_commit(&a0_empty[s]);
                    }
                    umma_commit(&d0_full[as]);
                    FR_T(22 + 4 * sy);
                }
            }
        }
    } else if (warp == 18) {
        // ===================== conv1 MMA issuer (its own thread: conv0 of later slices is never stuck behind a wait
        // for this slice's A1 operand, and vice versa) =====================
        if (elect_one()) {
            constexpr uint32_t idesc1 = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t aW1 = smem_u32(sW1), aA1 = smem_u32(sA1);
            mbar_wait(w_bar, 0);
            uint32_t g = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const uint32_t db = lt & 1;
                for (uint32_t sy = 0; sy < 4; ++sy, ++g) {
                    const uint32_t ab = g & 1;
                    mbar_wait(&a1_ready[ab], (g >> 1) & 1);
                    if (sy == 0) mbar_wait(&d1_empty[db], ((lt >> 1) & 1) ^ 1);
                    tc_fence_after();
#pragma unroll
                    for (int plane = 0; plane < 2; ++plane)
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tD1 + 64 * db, make_desc_k_sw128(aA1 + ab * FR_A1_BUF + (plane * 2 + kb) * FR_A1_SUB + k * 32),
                                          make_desc_k_sw128(aW1 + (sy * 2 + kb) * 8192 + k * 32), idesc1, (sy > 0 || plane > 0 || kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&a1_free[ab]);
                    FR_T(23 + 4 * sy);
                }
                umma_commit(&d1_full[db]);
                FR_T(36);
            }
        }
    } else if (warp >= 19) {
        // ===================== epilogue 1 (4 warps, one per TMEM lane quarter): act1 = LeakyReLU(BN(D1)), all 64 channels
        // of this thread's pixel → 128 contiguous bytes per plane; runs under the next tile's slices =====================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t lt = 0;
        pdl_wait();          // act1 planes are still being read by the previous forward's conv2 until it completes
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t db = lt & 1;
            mbar_wait(&d1_full[db], (lt >> 1) & 1);
            tc_fence_after();
            if (warp == 19 && lane == 0) FR_T(60);
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(tD1 + 64 * db + lane_addr, r0);
            tmem_ld_32x32b_x32(tD1 + 64 * db + lane_addr + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[db]);
            if (m < 98) {
                const size_t o = ((size_t)(tile >> 1) * 196 + (tile & 1) * 98 + m) * 64;
                uint4* dh = reinterpret_cast<uint4*>(p.act1_hi + o);
                uint4* dl = reinterpret_cast<uint4*>(p.act1_lo + o);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = 8 * j4 + 2 * j, c = half * 32 + i;
                            const uint32_t a0 = half ? r1[i] : r0[i], a1 = half ? r1[i + 1] : r0[i + 1];
                            bn_lrelu_split_x2(a0, a1, s1[c], s1[c + 1], h1[c], h1[c + 1], hi[j], lo[j]);
                        }
                        dh[half * 4 + j4] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        dl[half * 4 + j4] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
            if (warp == 19 && lane == 0) FR_T(61);
        }
        tc_fence_before();
    } else if (warp >= 2) {
        // ===================== epilogue 0 warps =====================
        // 16 warps: TMEM lane quarter q = warp % 4; (warp-2)/4 = 2·t + xp: accumulator tile t (units 128t..128t+127) and pixel xp
        // of the unit.  D0 row = one 16-byte unit u of the slab = image columns 8·(u%28).. of slab row u/28 = conv0 pixels
        // (sx = 2h, 2h+1) of conv1 pixel p = u >> 1 with h = u & 1; D0 column = xp·32 + co.
        const int q = warp & 3;
        const int tcol = (warp - 2) >> 2;             // 0..3
        const int t = tcol >> 1, xp = tcol & 1;
        const int u = t * 128 + q * 32 + lane;        // unit index inside the slab
        const bool live = u < 196;                    // 7 image rows × 28 units; the rest are don't-care MMA rows
        const int pix = u >> 1, h = u & 1;            // conv1 pixel (A1 row), K block of conv1's operand (sx >> 1 = h)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint8_t* a1row = sA1 + h * FR_A1_SUB + pix * 128;   // conv1 K slice (sy, sx = 2h + xp, c): 64-byte half xp of the row
        uint32_t g = 0, lt = 0;
        pdl_wait();          // act1 planes are still being read by the previous forward's conv2 until it completes
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            for (int sy = 0; sy < 4; ++sy, ++g) {
                const uint32_t as = g & 1, aph = (g >> 1) & 1;
                mbar_wait(&d0_full[as], aph);
                tc_fence_after();
                if (warp == 2 && lane == 0) FR_T(40 + 4 * sy);
                uint32_t r[32], H[16], L[16];
                tmem_ld_32x32b_x32(tD0 + as * 128 + lane_addr + (uint32_t)(tcol * 32), r);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d0_empty[as]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 sc = *reinterpret_cast<const float4*>(s0 + 4 * j);
                    const float4 sh = *reinterpret_cast<const float4*>(h0 + 4 * j);
                    bn_lrelu_split_x2(r[4 * j], r[4 * j + 1], sc.x, sc.y, sh.x, sh.y, H[2 * j], L[2 * j]);
                    bn_lrelu_split_x2(r[4 * j + 2], r[4 * j + 3], sc.z, sc.w, sh.z, sh.w, H[2 * j + 1], L[2 * j + 1]);
                }
                if (warp == 2 && lane == 0) FR_T(41 + 4 * sy);
                const uint32_t ab = g & 1;
                mbar_wait(&a1_free[ab], ((g >> 1) & 1) ^ 1);        // conv1 of slice g-2 has read this A1 buffer
                if (warp == 2 && lane == 0) FR_T(42 + 4 * sy);
                if (live) {
                    uint8_t* dst = a1row + ab * FR_A1_BUF;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sw = ((xp * 4 + j) ^ (pix & 7)) << 4;
                        *reinterpret_cast<uint4*>(dst + sw) = make_uint4(H[4 * j], H[4 * j + 1], H[4 * j + 2], H[4 * j + 3]);
                        *reinterpret_cast<uint4*>(dst + 2 * FR_A1_SUB + sw) = make_uint4(L[4 * j], L[4 * j + 1], L[4 * j + 2], L[4 * j + 3]);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a1_ready[ab]);
                if (warp == 2 && lane == 0) FR_T(43 + 4 * sy);
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
