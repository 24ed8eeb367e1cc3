// tc_igemm.cuh — implicit-GEMM convolution (k 4×4, pad 1, stride s) on tcgen05 for the PM trunk: no im2col planes.
// Reference: models_PM.py:121-132 (Conv2d(k4, s2, p1) layers of the trunk).
//
//   conv[b, oy, ox, n] = bias[n] + Σ_{ky,kx,c} (act_hi + act_lo)[b, s·oy+ky-1, s·ox+kx-1, c] · W[n, (ky·4+kx)·C + c]
//
// The activation stays where the InstanceNorm kernel left it: two bf16 planes, NHWC.  A K block of the GEMM is one filter
// tap (ky, kx) × 64 channels, and the A tile of an M tile (bh output rows × Wo pixels ≤ 128 rows) for that K block is ONE
// TMA box of a 4-D tensor map over the plane — dims (C, W, H, B), box (64, s·Wo, s·bh, 1) walked with element strides
// (1, s, s, 1) from (c0, kx-1, s·oy0+ky-1, b): the box lands as [bh·Wo rows × 128 B], 128-byte swizzled, exactly the
// K-major operand tile the MMA wants; the zero padding is TMA's out-of-bounds fill.  Everything else (B tiles, smem ring,
// single-thread MMA issue, double-buffered TMEM accumulator, smem-staged coalesced epilogue) is tc_gemm.cuh's structure.
// An M tile is nb images × bh output rows × Wo pixels (≤ 128 rows, the box's 4th dimension is the image): wide layers use
// nb = 1 and several rows, the small late layers (14×14, 13×13) one output row of 9 images, whichever fills the 128 MMA rows
// better.  Tile rows that do not exist (short last tile, batch tail) read TMA's zero fill and are never stored.
#pragma once

#include "tc_prims.cuh"

namespace tc {

struct IgemmParams {
    int B, Ho, Wo, bh, nb, tiles_y;     // M tile = nb images × bh output rows; tiles_y = ceil(Ho / bh) tiles per image group
    int C, N, stride;                   // input channels (multiple of 64), output channels (multiple of BN), conv stride
    float* out;                         // [B·Ho·Wo, N] fp32
    const float* bias;                  // [N]
    float2* stats;                      // optional (nb == 1 only): per (M tile, epilogue warp q) column statistics [tiles_m·4][N] (mean, M2)
};

constexpr int IG_STAGES = 3;
constexpr int IG_MAX_N = 1024;
// N tile 128: two epilogue groups (warps 2-5 / 6-9), one per accumulator stage.  N tile 256 (layers with ≥ 256 output
// channels): a k-block stage is A hi + A lo + B = 16 + 16 + 32 KB for 16 MMAs of 128 cycles — 32 B/clk of shared-memory fill
// instead of the 94 B/clk a 128-wide tile needs to keep the tensor pipe busy (measured: 58-72 % busy, the TMA fill being the
// limit), and the A boxes are re-fetched for half as many N tiles.  Three 64 KB stages leave room for ONE epilogue group,
// which drains the two accumulator stages alternately.
template <int BN>
__host__ __device__ constexpr int ig_groups() { return BN > 128 ? 1 : 2; }
template <int BN>
__host__ __device__ constexpr int ig_threads() { return 64 + 128 * ig_groups<BN>(); }   // warp 0 TMA, warp 1 MMA, then the epilogue groups

__device__ __forceinline__ void tma_load_4d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                                 int32_t c2, int32_t c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "l"(policy)
        : "memory");
}

template <int BN>
__host__ __device__ constexpr int ig_stage_bytes() { return 2 * 128 * 128 + BN * 128; }
template <int BN>
__host__ __device__ constexpr int ig_smem_bytes() {
    return IG_STAGES * ig_stage_bytes<BN>() + 1024 + 256 + IG_MAX_N * 4 + ig_groups<BN>() * 4 * 4096;
}

template <int BN>
__global__ void __launch_bounds__(ig_threads<BN>(), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB, const IgemmParams p) {
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = ig_stage_bytes<BN>();
    static_assert(BN == 256 || BN == 128 || BN == 64, "accumulator stages are 128 (256) columns apart");
    constexpr int IG_THREADS = ig_threads<BN>();
    constexpr int NG = ig_groups<BN>();
    constexpr uint32_t ACC_STRIDE = BN > 128 ? 256 : 128;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + IG_STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + IG_STAGES;
    uint64_t* tmem_full_bar = empty_bar + IG_STAGES;  // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    float* s_bias = reinterpret_cast<float*>(smem + IG_STAGES * STAGE_BYTES + 256);
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + IG_MAX_N);          // 4 KB per epilogue warp

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int tiles_m = ((p.B + p.nb - 1) / p.nb) * p.tiles_y;
    const int tiles_n = p.N / BN;
    const int cblocks = p.C / 64;                     // K blocks per filter tap
    const int total_kb = 16 * cblocks;
    const int total_items = tiles_m * tiles_n;
    const int rows_img = p.bh * p.Wo;                 // tile rows per image
    const uint32_t a_box_bytes = (uint32_t)(p.nb * rows_img) * 128u;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB);
        for (int s = 0; s < IG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * ACC_STRIDE);
    for (int i = threadIdx.x; i < p.N; i += IG_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    // rows of an A stage beyond the box are read by the MMA (M = 128) without ever being written: clear once
    for (int i = threadIdx.x; i < IG_STAGES * STAGE_BYTES / 16; i += IG_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const int mt = item % tiles_m, nt = item / tiles_m;
                const int g = mt / p.tiles_y, b = g * p.nb, oy0 = (mt - g * p.tiles_y) * p.bh;
                for (int kb = 0; kb < total_kb; ++kb, ++it) {
                    const int s = it % IG_STAGES;
                    const uint32_t ph = (it / IG_STAGES) & 1;
                    const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * 64, ky = tap >> 2, kx = tap & 3;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], 2 * a_box_bytes + B_BYTES);
                    tma_load_4d_hint(st, &tmA_hi, &full_bar[s], c0, kx - 1, p.stride * oy0 + ky - 1, b, kEvictLast);
                    tma_load_4d_hint(st + A_BYTES, &tmA_lo, &full_bar[s], c0, kx - 1, p.stride * oy0 + ky - 1, b, kEvictLast);
                    tma_load_2d_hint(st + 2 * A_BYTES, &tmB, &full_bar[s], kb * 64, nt * BN, kEvictLast);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
                const uint32_t as = li & 1, aph = (li >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
                for (int kb = 0; kb < total_kb; ++kb, ++it) {
                    const int s = it % IG_STAGES;
                    const uint32_t ph = (it / IG_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t b_ad = a_hi + 2 * A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = make_desc_k_sw128(b_ad + k * 32);
                        umma_bf16(d_tmem, make_desc_k_sw128(a_hi + k * 32), bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_bf16(d_tmem, make_desc_k_sw128(a_hi + A_BYTES + k * 32), bd, idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        // ===== epilogue: group g = (warp-2)/4 drains accumulator stage g (items li ≡ g mod 2); a single group drains both =====
        const int q = warp & 3;
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        uint8_t* wbuf = s_stage + (warp - 2) * 4096;
        const int sr = lane >> 3, sc = lane & 7;
        // the 8 tile rows this lane stores (q·32 + sr + 4i) as (image of the tile, row inside that image's part): fixed per thread
        int r_img[8], r_rem[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = q * 32 + sr + 4 * i;
            r_img[i] = r / rows_img;
            r_rem[i] = r - r_img[i] * rows_img;
        }
        uint32_t li = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
            if (NG == 2 && (li & 1) != grp) continue;
            const uint32_t as = NG == 2 ? grp : (li & 1);
            const int mt = item % tiles_m, nt = item / tiles_m;
            const int g = mt / p.tiles_y, b = g * p.nb, oy0 = (mt - g * p.tiles_y) * p.bh;
            const int valid_img = min(p.nb, p.B - b);                        // images of this tile that exist
            const int valid_rem = min(p.bh, p.Ho - oy0) * p.Wo;              // rows per image that exist
            const size_t out_row0 = ((size_t)b * p.Ho + oy0) * p.Wo;
            mbar_wait(&tmem_full_bar[as], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
                tmem_ld_wait();
                if (c0 + 32 >= BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                }
                const int nb = nt * BN + c0;
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nb + i);
                    w[i] = __float_as_uint(__uint_as_float(r[i]) + b4.x); w[i + 1] = __float_as_uint(__uint_as_float(r[i + 1]) + b4.y);
                    w[i + 2] = __float_as_uint(__uint_as_float(r[i + 2]) + b4.z); w[i + 3] = __float_as_uint(__uint_as_float(r[i + 3]) + b4.w);
                }
                __syncwarp();
                stage_write_row(wbuf, lane, w);
                __syncwarp();
                if (p.stats)      // rows 32q.. of the tile that exist (nb == 1: the tile lies inside one image)
                    p.stats[((size_t)mt * 4 + q) * p.N + nb + lane] = stage_col_stats(wbuf, lane, min(32, max(0, valid_rem - q * 32)));
                float* dst = p.out + out_row0 * (size_t)p.N + nb + sc * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = 4 * i + sr;
                    if (r_img[i] < valid_img && r_rem[i] < valid_rem)
                        *reinterpret_cast<uint4*>(dst + ((size_t)r_img[i] * p.Ho * p.Wo + r_rem[i]) * p.N) =
                            *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * ACC_STRIDE);
    }
}

}  // namespace tc
