// tc_front.cuh — fused PatchGAN front end: conv0 (k4 s4, 3→32) + BN + LeakyReLU + conv1 (k4 s4, 32→64) + BN +
// LeakyReLU in ONE persistent tcgen05 kernel; the 56×56×32 activation never leaves the SM.
// Reference: models_PMG.py:113-121 (the two strided layers of feature_extractor).
//
// Tile = 98 conv1 pixels (7 rows × 14 cols of the 14×14 map) of one image, 2 tiles per image; UMMA M = 128
// (rows 98..127 are don't-care).  For each of the 4 row offsets sy of conv1's 4×4 window:
//
//   * TMA (4-D tensor map over the NCHW bf16 image, dims x|py|y16|b·c, box 224×8×4×1, no swizzle) drops, per
//     channel c, four slabs (kernel rows ky) of 8 whole image rows each = [98(+14) pixels × 16 input columns]
//     straight into smem as rows of 32 bytes: no im2col, no register staging, 448-byte runs, ONE bulk tensor
//     copy per (sy, c) step (the 8th image row is the neighbouring tile's / zero fill: don't-care MMA rows).
//     The MMA reads the slab as a K-major SWIZZLE_32B operand.  The data are NOT swizzled, so for slab rows with
//     bit 2 set (r & 4) the hardware's chunk XOR hands the MMA the two 16-byte halves (input columns 0-7 / 8-15)
//     exchanged.  W0' is block-sparse over the column offset sx (4 columns each), so for those pixels the
//     accumulator column group cg simply holds the conv0 output of column offset sx = cg ^ 2; epilogue 0
//     (thread = pixel row) files it under the right K slice of conv1.  Nothing is moved twice.
//   * conv0 as a GEMM over those slabs:  D0[pixel, (sx,co)] = Σ_{c,ky} A[pixel, 16 cols] · W0'[(sx,co), (c,ky,16 cols)]
//     with W0' the conv0 weight laid out block-sparse over the 4 column offsets sx (each output column group sx
//     only sees its own 4 of the 16 input columns).  12 MMAs (M128 N128 K16) per sy; the 4× zero padding is
//     free: the tensor pipe is far from being the bottleneck of this layer.
//   * epilogue 0 (16 warps): TMEM → BN scale/shift → LeakyReLU → hi/lo bf16 split → written as conv1's A operand
//     (K-major, 128-byte swizzle) for the K slice (sy, sx, c).
//   * conv1 accumulates that K slice:  D1[pixel, 64] += A1(hi) · W1[sy] + A1(lo) · W1[sy]   (16 MMAs, M128 N64 K16).
//
// After the 4th slice, epilogue 1 (its own 4 warps, D1 double-buffered) applies BN + LeakyReLU and stores act1 as hi/lo
// bf16 planes [B·196, 64] while the other warps already work on the next tile.  conv0 and conv1 MMAs are issued by two
// different threads, so neither stream is blocked behind a barrier wait of the other.
// Both weight tensors stay resident in smem for the life of the CTA.  D0 is double-buffered in TMEM so the
// conv0 MMAs of slice sy+1 run under epilogue 0 of slice sy.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int FRONT_THREADS = 736;                 // warp 0 TMA, 1 conv0 MMA issuer, 2-17 epilogue 0 (lane quarter × 32-column group), 18 conv1 MMA issuer, 19-22 epilogue 1
constexpr int FR_W0_BYTES = 3 * 128 * 128;         // 48 KB: 3 K-blocks × [128 rows × 128 B]
constexpr int FR_W1_BYTES = 8 * 64 * 128;          // 64 KB: (sy, kb) × [64 rows × 128 B]
constexpr int FR_A0_SLOT = 3584;                   // one kernel row ky: 8 image rows × 448 B (7 used by the tile; 256-byte multiple)
constexpr int FR_A0_STAGE = 4 * FR_A0_SLOT;        // one channel: 4 kernel rows = ONE TMA box
constexpr int FR_A0_STAGES = 3;
constexpr int FR_A1_BYTES = 2 * 2 * 128 * 128;     // 64 KB: [plane hi/lo][kb][128 rows × 128 B]
constexpr int FR_BOX_BYTES = FR_A0_STAGE;        // box 224 × 8 × 4 × 1 bf16 (out-of-range image rows are zero-filled but counted)
constexpr int FR_OFF_W1 = FR_W0_BYTES;
constexpr int FR_OFF_A0 = FR_OFF_W1 + FR_W1_BYTES;
constexpr int FR_OFF_A1 = FR_OFF_A0 + FR_A0_STAGES * FR_A0_STAGE;
constexpr int FR_OFF_BAR = FR_OFF_A1 + FR_A1_BYTES;
constexpr int FR_OFF_PRM = FR_OFF_BAR + 256;
constexpr int FRONT_SMEM = FR_OFF_PRM + (32 + 32 + 64 + 64) * 4 + 1024;

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

// K-major operand whose rows are 32 bytes (16 bf16 = one UMMA K step), 32-byte swizzle, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t make_desc_k_sw32(uint32_t saddr) { return make_smem_desc(saddr, 16, 256, LAYOUT_SW32); }

struct FrontParams {
    const float* scale0; const float* shift0;    // [32]  BN(eval) of conv0 folded
    const float* scale1; const float* shift1;    // [64]
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
             const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1, const FrontParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW0 = smem;
    uint8_t* sW1 = smem + FR_OFF_W1;
    uint8_t* sA0 = smem + FR_OFF_A0;
    uint8_t* sA1 = smem + FR_OFF_A1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FR_OFF_BAR);
    uint64_t* a0_full = bars;            // [3]
    uint64_t* a0_empty = bars + 3;       // [3]
    uint64_t* d0_full = bars + 6;        // [2]
    uint64_t* d0_empty = bars + 8;       // [2]
    uint64_t* a1_ready = bars + 10;      // hi plane of A1 written
    uint64_t* a1_free = bars + 11;       // hi plane of A1 read by conv1
    uint64_t* a1l_ready = bars + 18;     // same for the lo plane
    uint64_t* a1l_free = bars + 19;
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
        prefetch_tmap(&tmX); prefetch_tmap(&tmXlo); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1);
        for (int i = 0; i < 3; ++i) { mbar_init(&a0_full[i], 1); mbar_init(&a0_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d0_full[i], 1); mbar_init(&d0_empty[i], 16); }
        mbar_init(a1_ready, 16); mbar_init(a1_free, 1); mbar_init(w_bar, 1);
        mbar_init(a1l_ready, 16); mbar_init(a1l_free, 1);
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
            for (int i = 0; i < 3; ++i) tma_load_2d(sW0 + i * 16384, &tmW0, w_bar, i * 64, 0);
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
            constexpr uint32_t idesc0 = make_idesc_bf16(128, 128, 0, 0);
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
#pragma unroll
                        for (int ky = 0; ky < 4; ++ky) {
                            const int kk = (c >= 3 ? c - 3 : c) * 4 + ky;     // the lo plane meets the same weights as the hi plane
                            umma_bf16(tD0 + as * 128, make_desc_k_sw32(aA0 + s * FR_A0_STAGE + ky * FR_A0_SLOT),
                                      make_desc_k_sw128(aW0 + (kk >> 2) * 16384 + (kk & 3) * 32), idesc0, (c > 0 || ky > 0) ? 1u : 0u);
                        }
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
                    mbar_wait(a1_ready, g & 1);
                    if (sy == 0) mbar_wait(&d1_empty[db], ((lt >> 1) & 1) ^ 1);
                    tc_fence_after();
                    // the hi plane first, then the lo plane, each with its own ready / free pair: epilogue 0 refills the hi
                    // plane for the next slice while the lo MMAs of this one run (and vice versa), so the single A1 buffer
                    // no longer serialises "write A1 → 16 MMAs → write A1"
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tD1 + 64 * db, make_desc_k_sw128(aA1 + kb * 16384 + k * 32),
                                      make_desc_k_sw128(aW1 + (sy * 2 + kb) * 8192 + k * 32), idesc1, (sy > 0 || kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(a1_free);
                    mbar_wait(a1l_ready, g & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tD1 + 64 * db, make_desc_k_sw128(aA1 + 32768 + kb * 16384 + k * 32),
                                      make_desc_k_sw128(aW1 + (sy * 2 + kb) * 8192 + k * 32), idesc1, 1u);
                    umma_commit(a1l_free);
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
        // 16 warps: TMEM lane quarter q = warp % 4, 32-column group cg = (warp-2)/4 = conv0 column offset sx
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;               // D0 columns cg*32 .. +31
        const int m = q * 32 + lane;                  // tile row = conv1 pixel index inside the tile
        const int sx = cg ^ ((m >> 1) & 2);           // conv0 column offset those columns hold for this row (header: half swap)
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint8_t* a1row = sA1 + (sx >> 1) * 16384 + m * 128;   // conv1 K slice (sy, sx, c): K block sx/2, 64-byte half sx%2
        uint32_t g = 0, lt = 0;
        pdl_wait();          // act1 planes are still being read by the previous forward's conv2 until it completes
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            for (int sy = 0; sy < 4; ++sy, ++g) {
                const uint32_t as = g & 1, aph = (g >> 1) & 1;
                mbar_wait(&d0_full[as], aph);
                tc_fence_after();
                if (warp == 2 && lane == 0) FR_T(40 + 4 * sy);
                uint32_t r[32], H[16], L[16];
                tmem_ld_32x32b_x32(tD0 + as * 128 + lane_addr + (uint32_t)(cg * 32), r);
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
                if (g > 0) mbar_wait(a1_free, (g - 1) & 1);        // conv1's hi MMAs of the previous slice have read the hi plane
                if (warp == 2 && lane == 0) FR_T(42 + 4 * sy);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sw = (((cg & 1) * 4 + j) ^ (m & 7)) << 4;
                    *reinterpret_cast<uint4*>(a1row + sw) = make_uint4(H[4 * j], H[4 * j + 1], H[4 * j + 2], H[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a1_ready);
                if (g > 0) mbar_wait(a1l_free, (g - 1) & 1);       // ... and the lo MMAs the lo plane
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int sw = (((cg & 1) * 4 + j) ^ (m & 7)) << 4;
                    *reinterpret_cast<uint4*>(a1row + 32768 + sw) = make_uint4(L[4 * j], L[4 * j + 1], L[4 * j + 2], L[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a1l_ready);
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
