// tc_gemm2.cuh — split-K GEMM on a CTA PAIR (tcgen05 cta_group::2), used for token_proj:
//
//   partial[s][M, 192] = (A_hi[M, Ks] + A_lo[M, Ks]) · B[192, Ks]ᵀ     bf16 operands, fp32 accumulate, raw fp32 partial sums
//
// Two CTAs of a cluster (two SMs of one TPC) share one 256 × 192 output tile: CTA r loads A rows [128r, 128r+128) and
// HALF of the B tile (rows [96r, 96r+96)) into its own shared memory; the leader CTA (rank 0) issues ONE
// tcgen05.mma.cta_group::2 per K step for both (M = 256), each SM's tensor core reading its own A rows and both B
// halves, and each CTA's TMEM receives its 128 accumulator rows.  Per SM that halves the B-operand shared-memory
// traffic and footprint (cta_group::1 MMAs of this shape run at ≈65 % of the rate).
//
// Protocol (all barriers that the MMA thread waits on live in the leader; everything it signals is multicast):
//   full[s]   (leader)  : leader's producer arms 2×stage bytes; BOTH CTAs' TMA loads complete_tx on it
//   empty[s]  (each CTA): tcgen05.commit.multicast from the leader → each producer may refill its own stage
//   tfull[a]  (each CTA): commit.multicast → each CTA's epilogue drains its own TMEM half
//   tempty[a] (leader)  : 4 local + 4 remote epilogue-warp arrivals
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int G2_THREADS = 192;     // warp 0 TMA, warp 1 MMA (leader only) + TMEM alloc, warps 2-5 epilogue
constexpr int G2_BN = 192;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the leader CTA (rank 0 of a pair)

template <int NA>
__host__ __device__ constexpr int g2_stage_bytes() { return NA * 128 * 128 + 96 * 128; }
template <int NA>
__host__ __device__ constexpr int g2_stages() { return (200 * 1024) / g2_stage_bytes<NA>() > 8 ? 8 : (200 * 1024) / g2_stage_bytes<NA>(); }
template <int NA>
__host__ __device__ constexpr int g2_smem_bytes() { return g2_stages<NA>() * g2_stage_bytes<NA>() + 1024 + 256; }

struct Gemm2Params {
    int M, K;                  // K % 64 == 0
    int kblocks_per_split;
    float* out;                // [nsplit][M, 192]
    size_t partial_stride;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` (same offset) in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 2-D tile load issued by either CTA of the pair; the transaction bytes are signalled on the LEADER's mbarrier
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

template <int NA>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
             const __grid_constant__ CUtensorMap tmB, const Gemm2Params p) {
    constexpr int STAGES = g2_stages<NA>();
    constexpr int STAGE_BYTES = g2_stage_bytes<NA>();
    constexpr int A_BYTES = 128 * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_m = (p.M + 255) / 256;
    const int total_kb = p.K / 64;
    const int nsplit = (total_kb + p.kblocks_per_split - 1) / p.kblocks_per_split;
    const int total_items = tiles_m * nsplit;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA_hi);
        if (NA == 2) prefetch_tmap(&tmA_lo);
        prefetch_tmap(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // barriers of both CTAs are initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            for (int item = pair; item < total_items; item += n_pairs) {
                const int mt = item % tiles_m, sp = item / tiles_m;
                const int kb0 = sp * p.kblocks_per_split, kb1 = min(total_kb, kb0 + p.kblocks_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    const int row = mt * 256 + (int)rank * 128;
                    tma2_load_2d(st, &tmA_hi, &full_bar[s], kb * 64, row, kEvictFirst);
                    if (NA == 2) tma2_load_2d(st + A_BYTES, &tmA_lo, &full_bar[s], kb * 64, row, kEvictFirst);
                    tma2_load_2d(st + NA * A_BYTES, &tmB, &full_bar[s], kb * 64, (int)rank * 96, kEvictLast);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(256, G2_BN, 0, 0);
            uint32_t it = 0, li = 0;
            for (int item = pair; item < total_items; item += n_pairs, ++li) {
                const int sp = item / tiles_m;
                const int kb0 = sp * p.kblocks_per_split, kb1 = min(total_kb, kb0 + p.kblocks_per_split);
                const uint32_t as = li & 1;
                mbar_wait(&tempty_bar[as], ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t b_ad = a_hi + NA * A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = make_desc_k_sw128(b_ad + k * 32);
                        umma2_bf16(d_tmem, make_desc_k_sw128(a_hi + k * 32), bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        if (NA == 2) umma2_bf16(d_tmem, make_desc_k_sw128(a_hi + A_BYTES + k * 32), bd, idesc, 1u);
                    }
                    umma2_commit_mc(&empty_bar[s]);
                }
                umma2_commit_mc(&tfull_bar[as]);
            }
        }
    } else {
        // ===== epilogue: 4 warps, each CTA drains its own 128 accumulator rows; raw fp32 partial sums, row slices of
        // 32 columns (128 B) per thread =====
        const int q = warp & 3;
        uint32_t li = 0;
        for (int item = pair; item < total_items; item += n_pairs, ++li) {
            const int mt = item % tiles_m, sp = item / tiles_m;
            const uint32_t as = li & 1;
            mbar_wait(&tfull_bar[as], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + as * 256 + ((uint32_t)(q * 32) << 16);
            const int row = mt * 256 + (int)rank * 128 + q * 32 + lane;
            float* dst = p.out + (size_t)sp * p.partial_stride + (size_t)row * G2_BN;
#pragma unroll 1
            for (int c0 = 0; c0 < G2_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
                tmem_ld_wait();
                if (c0 + 32 >= G2_BN) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(&tempty_bar[as]);
                        else mbar_arrive_remote(&tempty_bar[as], 0);
                    }
                }
                if (row < p.M) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<uint4*>(dst + c0 + 4 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();                 // the peer may still be reading this CTA's smem / signalling its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

}  // namespace tc
