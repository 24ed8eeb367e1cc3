// tc_gramb.cuh — generic batched Gram matrices on the tcgen05 tensor cores (BASELINE.json configs[1]):
//     G[b] = scale · F[b] · F[b]ᵀ,   F (B, C, HW) bf16 row-major (HW contiguous),  G (B, C, C) fp32.
// Reference op: torch.matmul(patches, patches.transpose(2,3)) / N — Models/models_PMG.py:188 and
// utils/PMG/PMG_style_transfer.py:44-45.
//
// F[b] is K-major for BOTH operands of D = A·Bᵀ (rows = channels, K = pixels), so one 3-D tensor map over
// (HW, C, B) feeds A (128-row boxes) and B (BN-row boxes); TMA zero-fills channel rows ≥ C and the pixel tail
// of the last 64-wide k-block, so any C (multiple of 16) and any HW (multiple of 8: 16-byte row pitch) work.
// Persistent CTAs walk work items (batch, m-tile, n-tile); warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..9 = two epilogue groups draining the double-buffered TMEM accumulator with coalesced staged stores.
// The full C×C matrix is computed (no symmetry shortcut), FLOPs counted as 2·C²·HW.
//
// LSU = true: the same kernel for inputs TMA cannot address — HW not a multiple of 8 (row pitch not a multiple of 16
// bytes: the PMG patch shapes HW = 25 and the 14×14 map HW = 196 of models_PMG.py:188 / PMG_style_transfer.py:44) or a base
// pointer that is not 16-byte aligned.  Eight loader warps replace the TMA producer: one warp-wide load moves the 64
// pixels (128 B) of one channel row of a k-block (4-byte lanes when every row start is 4-byte aligned, 2-byte lanes
// otherwise), the lanes store them straight into the K-major 128B-swizzled operand tile, pixel tail and rows ≥ C zero.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int GB_THREADS = 320;
constexpr int GB_LOADER_WARPS = 8;
constexpr int GB_THREADS_LSU = GB_THREADS + 32 * GB_LOADER_WARPS;

struct GramBParams {
    float* G;
    int B, C, HW;
    float scale;
    const __nv_bfloat16* F;      // LSU variant only
};

// LSU staging of one k-block (pixels k0 .. k0+w-1 of every row, w ≤ 64).  The (128 + BN) row pieces are walked as ONE linear
// index space of VEC-pixel slots (VEC = 4 / 2 / 1 pixels = 8 / 4 / 2-byte loads, the widest the row pitch and the base pointer
// allow), so a warp-wide load is fully used whatever HW is — for a single k-block (HW ≤ 64: the PMG patch shape HW = 25) the
// slots of consecutive rows are contiguous in memory.  GB_LOADS_IN_FLIGHT loads are issued before the first store, so every
// warp keeps that many independent global loads in flight.
constexpr int GB_LOADS_IN_FLIGHT = 8;

template <int VEC, int BN>
__device__ __forceinline__ void gb_stage_kblock(uint8_t* st, const __nv_bfloat16* Fb, int mt, int nt, int C, int HW, int k0, int lw, int lane, bool same) {
    constexpr int A_BYTES = 128 * 128;
    const int w = min(64, HW - k0);                         // valid pixels of this k-block
    // a ring stage is reused by k-blocks of different width when HW > 64: then the whole 64-pixel block is written (zeros past
    // w); with a single k-block every item has the same footprint and the tail stays as cleared at kernel start
    const int wv = ((HW > 64 ? 64 : w) + VEC - 1) / VEC;    // slots per row
    const int S = (same ? 128 : 128 + BN) * wv;            // same: the B tile is the head of the A tile (diagonal item)
    for (int base = lw * 32; base < S; base += GB_LOADER_WARPS * 32 * GB_LOADS_IN_FLIGHT) {
        uint32_t v[GB_LOADS_IN_FLIGHT][VEC == 4 ? 2 : 1];
        int rr[GB_LOADS_IN_FLIGHT], ee[GB_LOADS_IN_FLIGHT];
#pragma unroll
        for (int j = 0; j < GB_LOADS_IN_FLIGHT; ++j) {
            const int i = base + j * GB_LOADER_WARPS * 32 + lane;
            const int r = i / wv, e = (i - r * wv) * VEC;    // tile row (A rows 0..127, then B rows), first pixel of the slot
            rr[j] = i < S ? r : -1; ee[j] = e;
            const int row = r >= 128 ? nt * BN + r - 128 : mt * 128 + r;
            v[j][0] = 0;
            if (VEC == 4) v[j][1] = 0;
            if (i < S && row < C && e < w) {
                const __nv_bfloat16* src = Fb + (size_t)row * HW + k0 + e;
                if (VEC == 4) {                              // HW % 4 == 0: the slot is whole
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(src));
                    v[j][0] = t.x; v[j][1] = t.y;
                } else if (VEC == 2) {                       // HW % 2 == 0: the slot is whole
                    v[j][0] = __ldg(reinterpret_cast<const unsigned int*>(src));
                } else {
                    v[j][0] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(src));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < GB_LOADS_IN_FLIGHT; ++j) {
            if (rr[j] < 0) continue;
            const int rl = rr[j] >= 128 ? rr[j] - 128 : rr[j], e = ee[j];
            uint8_t* dst = st + (rr[j] >= 128 ? A_BYTES : 0) + rl * 128 + (((e >> 3) ^ (rl & 7)) << 4) + (e & 7) * 2;
            if (VEC == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(v[j][0], v[j][1]);
            else if (VEC == 2) *reinterpret_cast<uint32_t*>(dst) = v[j][0];
            else *reinterpret_cast<unsigned short*>(dst) = (unsigned short)v[j][0];
        }
    }
}

__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

template <int BN>
__host__ __device__ constexpr int gb_stage_bytes() { return 128 * 128 + BN * 128; }
template <int BN>
__host__ __device__ constexpr int gb_stages() { return (160 * 1024) / gb_stage_bytes<BN>() > 8 ? 8 : (160 * 1024) / gb_stage_bytes<BN>(); }
template <int BN>
__host__ __device__ constexpr int gb_smem_bytes() { return gb_stages<BN>() * gb_stage_bytes<BN>() + 1024 + 256 + 8 * 4096; }

template <int BN, bool LSU = false>
__global__ void __launch_bounds__(LSU ? GB_THREADS_LSU : GB_THREADS, 1)
gram_batched_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GramBParams p) {
    constexpr int STAGES = gb_stages<BN>();
    constexpr int STAGE_BYTES = gb_stage_bytes<BN>();
    constexpr int A_BYTES = 128 * 128;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 128, "UMMA N");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* s_stage = smem + STAGES * STAGE_BYTES + 256;      // 8 × 4 KB epilogue staging

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int tiles_m = (p.C + 127) / 128, tiles_n = (p.C + BN - 1) / BN;
    const int kblocks = (p.HW + 63) / 64;
    const int total_items = p.B * tiles_m * tiles_n;

    if (threadIdx.x == 0) {
        if (!LSU) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], LSU ? GB_LOADER_WARPS : 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    if (LSU) {      // the loaders never write the pixel tail of a k-block: clear the operand stages once (see gb_stage_kblock)
        for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (LSU && warp >= GB_THREADS / 32) {
        // ===== loader warps (LSU variant): warp lw stages rows lw, lw+8, … of the A tile and of the B tile =====
        const int lw = warp - GB_THREADS / 32;
        const uintptr_t fa = reinterpret_cast<uintptr_t>(p.F);
        const int vec = ((p.HW & 3) == 0 && (fa & 7) == 0) ? 4 : ((p.HW & 1) == 0 && (fa & 3) == 0) ? 2 : 1;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const int nt = item % tiles_n, mt = (item / tiles_n) % tiles_m, b = item / (tiles_n * tiles_m);
            const __nv_bfloat16* Fb = p.F + (size_t)b * p.C * p.HW;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                const bool same = nt * BN == mt * 128;
                if (vec == 4) gb_stage_kblock<4, BN>(st, Fb, mt, nt, p.C, p.HW, kb * 64, lw, lane, same);
                else if (vec == 2) gb_stage_kblock<2, BN>(st, Fb, mt, nt, p.C, p.HW, kb * 64, lw, lane, same);
                else gb_stage_kblock<1, BN>(st, Fb, mt, nt, p.C, p.HW, kb * 64, lw, lane, same);
                fence_proxy_async();                   // generic-proxy smem writes → visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
            }
        }
    } else if (warp == 0) {
        if (!LSU && elect_one()) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const int nt = item % tiles_n, mt = (item / tiles_n) % tiles_m, b = item / (tiles_n * tiles_m);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    // diagonal item: the B rows are the first BN rows of the A tile — one load, one operand tile
                    const bool same = nt * BN == mt * 128;
                    mbar_expect_tx(&full_bar[s], same ? A_BYTES : STAGE_BYTES);
                    tma_load_3d_hint(st, &tmA, &full_bar[s], kb * 64, mt * 128, b, kEvictFirst);
                    if (!same) tma_load_3d_hint(st + A_BYTES, &tmB, &full_bar[s], kb * 64, nt * BN, b, kEvictFirst);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
                const uint32_t as = li & 1;
                const int nt = item % tiles_n, mt = (item / tiles_n) % tiles_m;
                const bool same = nt * BN == mt * 128;
                mbar_wait(&tmem_empty_bar[as], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 128;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_ad = smem_u32(smem + s * STAGE_BYTES), b_ad = same ? a_ad : a_ad + A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, make_desc_k_sw128(a_ad + k * 32), make_desc_k_sw128(b_ad + k * 32), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full_bar[as]);
            }
        }
    } else {
        const int q = warp & 3;
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        uint8_t* wbuf = s_stage + (warp - 2) * 4096;
        const int sr = lane >> 3, sc = lane & 7;
        uint32_t li = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++li) {
            if ((li & 1) != grp) continue;
            const int nt = item % tiles_n, mt = (item / tiles_n) % tiles_m, b = item / (tiles_n * tiles_m);
            mbar_wait(&tmem_full_bar[grp], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + grp * 128 + ((uint32_t)(q * 32) << 16);
            float* gout = p.G + (size_t)b * p.C * p.C;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                if (BN - c0 >= 32 || BN % 32 == 0) {
                    tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
                } else {                                           // BN % 32 == 16: last chunk has 16 columns
                    uint32_t r16[16];
                    tmem_ld_32x32b_x16(t_addr + (uint32_t)c0, r16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0; }
                }
                tmem_ld_wait();
                if (c0 + 32 >= BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[grp]);
                }
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(r[i]) * p.scale);
                __syncwarp();
                stage_write_row(wbuf, lane, w);
                __syncwarp();
                const int row0 = mt * 128 + q * 32 + sr;
                const int col = nt * BN + c0 + sc * 4;
                if (col < p.C && col < nt * BN + BN) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + sr, row = row0 + 4 * i;
                        if (row < p.C)
                            *reinterpret_cast<uint4*>(gout + (size_t)row * p.C + col) = *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((sc ^ (rr & 7)) << 4));
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc
