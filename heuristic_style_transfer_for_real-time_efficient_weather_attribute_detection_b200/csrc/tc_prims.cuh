// tc_prims.cuh — thin inline-PTX wrappers for the sm_100a primitives the PMG fast path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences), UMMA descriptors.
// Hand-written; bit layouts follow the PTX ISA "tcgen05" descriptor tables.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every fast-path kernel is launched with programmaticStreamSerialization: its prologue (barrier init, TMEM
// allocation, tensor-map prefetch, weight staging) overlaps the tail of the previous kernel; pdl_wait() blocks
// until the previous grid has completed and its writes are visible, so it must precede the first read of any
// activation.  pdl_trigger() lets the next kernel in the stream start ITS prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (A suspend-time hint on try_wait — `…, parity, 0x989680` as CUTLASS's ClusterBarrier::wait uses — was measured: no change
// in any kernel, the default try_wait already parks the warp; without it the bounded spin below stays a usable deadlock trap.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must not hang the box (gpurun strikes) — trap instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 23)) __trap();
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates innermost-first (c0 = contiguous dim).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;   // createpolicy L2::evict_first (fraction 1.0)
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;    // createpolicy L2::evict_last

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] · B[smem]ᵀ, bf16 in / fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem]ᵀ: A operand from tensor memory (lane = row, two bf16 per 32-bit column, 8 columns per K = 16)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM → registers: this warp's 32 lanes × 32 consecutive fp32 columns (thread t ↔ lane base+t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// registers → TMEM: this warp's 32 lanes × 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (64 bit):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100) | [49,52) base offset | [61,64) layout (0 none, 2 SW128, 4 SW64, 6 SW32)
enum : uint64_t { LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6 };

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
// K-major operand tile stored as rows of 128 bytes (64 bf16) with the 128B swizzle (what TMA SWIZZLE_128B
// writes): 8-row groups are 1024 B apart; a K=16 step advances the start address by 32 B.
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024, LAYOUT_SW128); }

// Instruction descriptor (32 bit) for kind::f16 with bf16 A/B and fp32 accumulate:
//   [4,6) D format (1 = f32) | [7,10) A format (1 = bf16) | [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN) | [16] B major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- packed fp32 (sm_100: FFMA2 / FMUL2 / FADD2)
// Two fp32 operations per instruction on a 64-bit register pair.  The BN / LeakyReLU / hi-lo epilogues of this library
// are bound by instruction issue (≈6 instructions per accumulator value on 16 warps), not by the fp32 pipe, so halving
// the arithmetic instructions is a direct gain; results are bit-identical to the scalar forms (same .rn operations).
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---------------------------------------------------------------- split-precision helpers
// v ≈ hi + lo with hi = bf16(v), lo = bf16(v - hi): 16 significant bits through two bf16 MMAs.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<uint32_t*>(&h);
    float ra, rb;
    unpack_f32x2(sub_f32x2(pack_f32x2(a, b), pack_f32x2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u))), ra, rb);
    __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
// LeakyReLU_0.2(x·scale + shift) of two accumulator values (folded eval BatchNorm), then the hi/lo split
__device__ __forceinline__ void bn_lrelu_split_x2(uint32_t x0, uint32_t x1, float s0, float s1, float h0, float h1, uint32_t& hi, uint32_t& lo) {
    const uint64_t v = fma_f32x2(pack_f32x2(__uint_as_float(x0), __uint_as_float(x1)), pack_f32x2(s0, s1), pack_f32x2(h0, h1));
    const uint64_t w = mul_f32x2(v, pack_f32x2(0.2f, 0.2f));
    float v0, v1, w0, w1;
    unpack_f32x2(v, v0, v1);
    unpack_f32x2(w, w0, w1);
    split_bf16x2(fmaxf(v0, w0), fmaxf(v1, w1), hi, lo);
}

// 256-bit global store (sm_100: STG.E.ENL2.256): one whole 32-byte sector per thread
__device__ __forceinline__ void st_global_v8(void* gptr, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(gptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}

// ---------------------------------------------------------------- coalescing epilogue staging
// A warp owns 32 accumulator rows (thread = row).  Writing each thread's row straight to global gives 32
// scattered 16-byte pieces per store instruction (half-filled sectors, LSU-bound).  Instead each thread parks
// its 128-byte row slice in a warp-private, chunk-swizzled 4 KB smem tile (conflict-free STS.128), and the warp
// re-reads it so that 8 consecutive lanes hold one row's 128 contiguous bytes (conflict-free LDS.128):
// instruction i covers rows 4i..4i+3.
__device__ __forceinline__ void stage_write_row(uint8_t* wbuf, int lane, const uint32_t (&w)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(wbuf + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}
// Per-column statistics of the staged 32-row × 32-column fp32 tile (lane = column; conflict-free thanks to the chunk swizzle):
// mean and Σ(x − mean)² of the first nvalid rows, two passes over values held in registers.  InstanceNorm of the PM trunk
// merges these per-warp partials (Chan et al.) instead of re-reading the whole convolution output for its statistics.
__device__ __forceinline__ float2 stage_col_stats(const uint8_t* wbuf, int lane, int nvalid) {
    float v[32];
#pragma unroll
    for (int rr = 0; rr < 32; ++rr)
        v[rr] = *reinterpret_cast<const float*>(wbuf + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int rr = 0; rr < 32; rr += 4) {
        s0 += rr < nvalid ? v[rr] : 0.f; s1 += rr + 1 < nvalid ? v[rr + 1] : 0.f;
        s2 += rr + 2 < nvalid ? v[rr + 2] : 0.f; s3 += rr + 3 < nvalid ? v[rr + 3] : 0.f;
    }
    const float mean = nvalid > 0 ? ((s0 + s1) + (s2 + s3)) / (float)nvalid : 0.f;
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int rr = 0; rr < 32; rr += 4) {
        const float d0 = rr < nvalid ? v[rr] - mean : 0.f, d1 = rr + 1 < nvalid ? v[rr + 1] - mean : 0.f;
        const float d2 = rr + 2 < nvalid ? v[rr + 2] - mean : 0.f, d3 = rr + 3 < nvalid ? v[rr + 3] - mean : 0.f;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
    }
    return make_float2(mean, (q0 + q1) + (q2 + q3));
}
__device__ __forceinline__ uint4 stage_read(const uint8_t* wbuf, int lane, int i, int& r, int& c) {
    r = 4 * i + (lane >> 3);
    c = lane & 7;
    return *reinterpret_cast<const uint4*>(wbuf + r * 128 + ((c ^ (r & 7)) << 4));
}

}  // namespace tc
