// tc_encoder.cuh — the whole token stage of the PMG forward in ONE persistent tcgen05 kernel:
//   token_proj split-K reduce + bias  →  nl (≤ 2) × pre-LN TransformerEncoderLayer (d 192, 4 heads × 48, FFN 384, ReLU)
//   →  task attention pooling + 12 linear heads  →  logits.
// Reference: models_PMG.py:60-72,144-147,194 (encoder), :47-57 (TaskAttentionPooling), :153-155,198 (heads).
//
// Tile = 7 images = 63 token rows (row 63 is a don't-care); B = 1024 gives 147 tiles for the 148 SMs.  The
// residual stream x of the tile lives in TMEM (columns 0..191, fp32) for the whole kernel: the out-proj and
// FFN2 MMAs accumulate straight onto it, so the residual adds cost nothing.  Activations that feed an MMA are
// written by the epilogue warps as two bf16 planes (hi + lo, DESIGN.md "precision") into K-major 128B-swizzled
// smem operand tiles; all layer weights stream through a 3-stage ring of 24 KB blocks with cp.async.bulk from a
// pre-swizzled copy (enc_pack_kernel), so one bulk copy lands a ready-to-use UMMA B tile.  (A 4th stage — paid for by
// letting O share A's shared memory — was built and measured in round 2: no change, the MMA thread still waits ≈18-20 k
// cycles per tile for weights.  148 SMs pulling the same 1.2 MB at once is an aggregate L2 rate of ≈7 TB/s: the supply is
// bandwidth-bound at the L2, not latency-bound at the ring.)  Every small parameter
// (biases, LayerNorm affine, task map) is staged in smem once per CTA, before griddepcontrol.wait, so no global
// load sits on the per-phase critical path.
//
// UMMA M is 128 but only 64 rows are real: operand k-blocks are packed 8 KB apart (64 rows × 128 B), so rows
// 64..127 of k-block j alias rows 0..63 of k-block j+1 — the MMA computes throw-away rows in TMEM lanes 64..127
// (rows of D depend only on the same row of A) and smem holds only the real rows.
//
// Warp roles (18 warps): 0 = weight producer, 1 = MMA issuer (+ TMEM alloc), warps w with w%4 ∈ {0,1}, w ≥ 4
// = 8 "compute" warps (TMEM lane quarter q = w%4 holds rows 32q..32q+31; column group g = (w-4)/4), warps with
// w%4 ∈ {2,3} = 8 "aux" warps that run the 9-token attention core out of shared memory.
//
// Per layer:  LN1 → A tile │ per head h: MMA [q_h|k_h|v_h] (N 144, TMEM double-buffered) → compute warps drain
// (+bias, q·log2e/√48) to smem fp32 → aux warps: scores, softmax, P·V → O tile (hi/lo) │ out-proj MMA onto x │
// +bias, LN2 → A tile │ FFN in 3 chunks of 128 hidden units: FFN1 MMA (N 128, TMEM double-buffered) → +bias,
// ReLU, hi/lo → H tile → FFN2 MMA onto x (chunk c+1's FFN1 runs under chunk c's epilogue) │ +bias.
// After the last layer: x hi/lo → A tile → one MMA against [queries | classifier rows] (N 64) → softmax-weighted
// sum over the 9 tokens (the heads are linear in the pooled token) → logits.
#pragma once

#include "tc_prims.cuh"

namespace tc {

constexpr int ENC_ATTN_WARPS = 8;
constexpr int ENC_THREADS = 576, ENC_W_PROD = 0, ENC_W_MMA = 1, ENC_W_CMP0 = 4;
constexpr int ENC_IPT = 7;                   // images per tile
constexpr int ENC_ROWS = 63;                 // token rows per tile
constexpr int ENC_MAX_LAYERS = 2;            // parameters of every layer stay resident in smem
constexpr int ENC_STAGE = 24576;             // weight ring stage
constexpr int ENC_NSTAGE = 3;
constexpr int ENC_QKV_LD = 148;              // floats per row of the fp32 [q|k|v] head buffer (144 + 4: conflict-free 16-byte row accesses)
constexpr int ENC_KB = 8192;                 // operand k-block: 64 rows × 128 B
constexpr int ENC_PLANE = 3 * ENC_KB;        // [64 rows × 192 K] bf16
constexpr int ENC_OFF_QKV = ENC_NSTAGE * ENC_STAGE;                 // fp32 head buffer; aliased by the FFN hidden tile and the heads result
constexpr int ENC_QKV_BYTES = 64 * ENC_QKV_LD * 4;                  // 37,888 = 37 KB
constexpr int ENC_OFF_A = ENC_OFF_QKV + ENC_QKV_BYTES;              // LN output / final tokens, hi then lo plane
constexpr int ENC_OFF_O = ENC_OFF_A + 2 * ENC_PLANE;                // attention output, hi then lo plane
constexpr int ENC_SC_LD = 65;                                       // [64][65] fp32 heads-GEMM result (inside the QKV region)
// per-layer parameter block (floats): b_in 576 | b_o 192 | b_1 384 | b_2 192 | ln1w | ln1b | ln2w | ln2b (192 each)
constexpr int ENC_P_BIN = 0, ENC_P_BO = 576, ENC_P_B1 = 768, ENC_P_B2 = 1152, ENC_P_LN1W = 1344, ENC_P_LN1B = 1536,
              ENC_P_LN2W = 1728, ENC_P_LN2B = 1920, ENC_P_LAYER = 2112;
constexpr int ENC_P_TOKB = ENC_MAX_LAYERS * ENC_P_LAYER, ENC_P_CLSB = ENC_P_TOKB + 192, ENC_P_FLOATS = ENC_P_CLSB + 64;
constexpr int ENC_OFF_PRM = ENC_OFF_O + 2 * ENC_PLANE;              // also the ≥ 8 KB over-read slack behind the O tile
constexpr int ENC_OFF_TASK = ENC_OFF_PRM + ENC_P_FLOATS * 4;        // task_of[64] bytes
constexpr int ENC_OFF_RED = ENC_OFF_TASK + 64;                      // LayerNorm partial sums: 2 × [4][64] fp32
constexpr int ENC_OFF_BAR = ENC_OFF_RED + 2 * 4 * 64 * 4;
constexpr int ENC_SMEM = ENC_OFF_BAR + 256 + 1024;
static_assert(ENC_SMEM <= 232448, "token-stage kernel exceeds the 227 KB shared-memory limit");
static_assert(ENC_OFF_QKV % 1024 == 0 && ENC_OFF_A % 1024 == 0 && ENC_OFF_O % 1024 == 0 && ENC_OFF_PRM % 16 == 0, "operand tiles need 1024-byte alignment");

// weight stream geometry (bytes), per layer, in the order the MMA warp consumes it
constexpr int ENC_QKV_BLK = 144 * 128;                               // one k-block of one head's [q|k|v] rows
constexpr int ENC_F1_BLK = 128 * 128;                                // one k-block of one FFN1 chunk (128 hidden units)
constexpr int ENC_L_QKV = 0;
constexpr int ENC_L_OUT = 12 * ENC_QKV_BLK;                          // 221,184
constexpr int ENC_L_FFN = ENC_L_OUT + 3 * ENC_STAGE;                 // 294,912: F1(0) F1(1) F2(0) F1(2) F2(1) F2(2)
constexpr int ENC_LAYER_BYTES = ENC_L_FFN + 9 * ENC_F1_BLK + 6 * ENC_STAGE;   // 589,824

struct EncLayerPtrs {
    const float *b_in, *b_o, *b_1, *b_2, *ln1w, *ln1b, *ln2w, *ln2b;
};
struct EncParams {
    const uint8_t* wstream;          // nl × ENC_LAYER_BYTES + ENC_STAGE (heads), pre-swizzled bf16
    const float* partial;            // token_proj split-K partial sums [nsplit][T][192]
    size_t pstride;
    const float* tok_bias;           // [192]
    const float* cls_bias;           // [ncls_total]
    float* logits;                   // [B][ncls_total]
    float* peer_logits[15];          // optional copies of the logits (peer-mapped memory of other GPUs): stored from the same epilogue
    int n_peers;
    float* tok_in;                   // optional taps [T][192]
    float* tok_out;
    EncLayerPtrs layer[ENC_MAX_LAYERS];
    int nsplit, nl, B, ntask, ncls_total;
    unsigned char task_of[64];
    long long* trace;                // optional (pmg_debug_set_trace): clock64 stamps of CTA 0's roles, 512 slots
};
#define ENC_T(slot) do { if (p.trace && blockIdx.x == 0) p.trace[slot] = clock64(); } while (0)

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictLast)
        : "memory");
}
// 16-byte asynchronous copy global → shared; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 48 fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld48(uint32_t taddr, float (&v)[48]) {
    uint32_t r0[16], r1[16], r2[16];
    tmem_ld_32x32b_x16(taddr, r0);
    tmem_ld_32x32b_x16(taddr + 16, r1);
    tmem_ld_32x32b_x16(taddr + 32, r2);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r0[i]); v[16 + i] = __uint_as_float(r1[i]); v[32 + i] = __uint_as_float(r2[i]); }
}
__device__ __forceinline__ void tmem_st48(uint32_t taddr, const float (&v)[48]) {
    uint32_t r[16];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[16 * b + i]);
        tmem_st_32x32b_x16(taddr + 16 * b, r);
    }
}

// write 48 values (columns 48g..48g+47 of row m) as hi/lo bf16 into a [64 × 192] K-major SW128 operand tile
__device__ __forceinline__ void write_operand48(uint8_t* tile_hi, int m, int g, const float (&y)[48]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16x2(y[8 * c + 2 * i], y[8 * c + 2 * i + 1], hi[i], lo[i]);
        const int chunk = 6 * g + c;                       // 16-byte chunk index inside the 384-byte row
        uint8_t* dst = tile_hi + (chunk >> 3) * ENC_KB + m * 128 + (((chunk & 7) ^ (m & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(dst + ENC_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// v += bias (48 floats from shared memory)
__device__ __forceinline__ void add_bias48(float (&v)[48], const float* b) {
#pragma unroll
    for (int i = 0; i < 48; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(b + i);
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
    }
}

// LayerNorm (in place) over 192 columns held by the 4 column-group threads of a row (48 each); two-pass, fp32;
// w / b point at this thread's 48-column slice in shared memory.
__device__ __forceinline__ void layer_norm48(float (&v)[48], float* red, int m, int g, const float* w, const float* b) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 48; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
    red[g * 64 + m] = (s0 + s1) + (s2 + s3);
    named_bar_sync(1, 256);
    const float mu = ((red[m] + red[64 + m]) + (red[128 + m] + red[192 + m])) * (1.f / 192.f);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 48; i += 4) {
        const float d0 = v[i] - mu, d1 = v[i + 1] - mu, d2 = v[i + 2] - mu, d3 = v[i + 3] - mu;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
    }
    red[256 + g * 64 + m] = (q0 + q1) + (q2 + q3);
    named_bar_sync(1, 256);
    const float var = ((red[256 + m] + red[320 + m]) + (red[384 + m] + red[448 + m])) * (1.f / 192.f);
    const float rs = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int i = 0; i < 48; i += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + i);
        const float4 b4 = *reinterpret_cast<const float4*>(b + i);
        v[i] = fmaf((v[i] - mu) * rs, w4.x, b4.x);
        v[i + 1] = fmaf((v[i + 1] - mu) * rs, w4.y, b4.y);
        v[i + 2] = fmaf((v[i + 2] - mu) * rs, w4.z, b4.z);
        v[i + 3] = fmaf((v[i + 3] - mu) * rs, w4.w, b4.w);
    }
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ------------------------------------------------------------------------------------------------
// 9-token attention core of one head out of the fp32 [q|k|v] head buffer (models_PMG.py:64: nn.MultiheadAttention,
// 4 heads × 48, scale 1/sqrt(48) folded into q together with log2 e) on the eight aux warps: a thread owns ONE query row
// × 12 of the head's 48 dims (4 lanes per row, 8 rows per warp); the 4 lanes of a row sum their partial q·k with two
// xor-shuffles: 2.85 k cycles per head.  Measured alternatives: three rows per thread on three warps (round 1; two of the
// warps share a sub-partition) 3.1 k; three rows per thread with packed FFMA2 on four warps (a third of the LDS traffic, 16
// warps and 128 registers for the kernel) 3.45 k — two warps per sub-partition cannot hide the dependent LDS → FMA → shuffle
// → exp chains.
// ------------------------------------------------------------------------------------------------
struct AttnCtx1 {
    const float* qp; const float* kp; const float* vp;
    int row, part;
    bool active;
};
__device__ __forceinline__ AttnCtx1 attn_ctx1(const float* sQKV, int ai, int lane) {
    AttnCtx1 c;
    const int r = ai * 8 + (lane >> 2);
    c.active = r < ENC_ROWS;
    c.row = c.active ? r : ENC_ROWS - 1;
    const int img = c.row / 9;
    c.part = lane & 3;
    c.qp = sQKV + c.row * ENC_QKV_LD + 12 * c.part;
    c.kp = sQKV + img * 9 * ENC_QKV_LD + 48 + 12 * c.part;
    c.vp = c.kp + 48;
    return c;
}
__device__ __forceinline__ void attention_rows(const AttnCtx1& ac, float (&o)[12], int lane) {
    float4 qv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) qv[c] = *reinterpret_cast<const float4*>(ac.qp + 4 * c);
    float sc[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;                // three independent chains
        const float4 k0 = *reinterpret_cast<const float4*>(ac.kp + j * ENC_QKV_LD);
        const float4 k1 = *reinterpret_cast<const float4*>(ac.kp + j * ENC_QKV_LD + 4);
        const float4 k2 = *reinterpret_cast<const float4*>(ac.kp + j * ENC_QKV_LD + 8);
        a0 = fmaf(qv[0].x, k0.x, fmaf(qv[0].y, k0.y, fmaf(qv[0].z, k0.z, qv[0].w * k0.w)));
        a1 = fmaf(qv[1].x, k1.x, fmaf(qv[1].y, k1.y, fmaf(qv[1].z, k1.z, qv[1].w * k1.w)));
        a2 = fmaf(qv[2].x, k2.x, fmaf(qv[2].y, k2.y, fmaf(qv[2].z, k2.z, qv[2].w * k2.w)));
        sc[j] = (a0 + a1) + a2;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], 1);
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 9; ++j) { sc[j] += __shfl_xor_sync(0xffffffffu, sc[j], 2); mx = fmaxf(mx, sc[j]); }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) { sc[j] = fast_exp2(sc[j] - mx); den += sc[j]; }
    const float inv = fast_rcp(den);
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const float pj = sc[j] * inv;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v4 = *reinterpret_cast<const float4*>(ac.vp + j * ENC_QKV_LD + 4 * c);
            o[4 * c] = fmaf(pj, v4.x, o[4 * c]); o[4 * c + 1] = fmaf(pj, v4.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(pj, v4.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pj, v4.w, o[4 * c + 3]);
        }
    }
}
// this thread's 12 output dims of head h → O tile (hi/lo planes, 8-byte pieces of the swizzled 128-byte rows)
__device__ __forceinline__ void store_o_rows(const AttnCtx1& ac, uint8_t* sO, int h, const float (&o)[12]) {
    if (!ac.active) return;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int e = 48 * h + 12 * ac.part + 4 * t;         // element column, multiple of 4
        const int ce = e & 63;
        uint32_t h0, l0, h1, l1;
        split_bf16x2(o[4 * t], o[4 * t + 1], h0, l0);
        split_bf16x2(o[4 * t + 2], o[4 * t + 3], h1, l1);
        uint8_t* dst = sO + (e >> 6) * ENC_KB + ((ce & 7) << 1) + ac.row * 128 + (((ce >> 3) ^ (ac.row & 7)) << 4);
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(dst + ENC_PLANE) = make_uint2(l0, l1);
    }
}

__global__ void __launch_bounds__(ENC_THREADS, 1) encoder_kernel(const EncParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;
    float* sQKV = reinterpret_cast<float*>(smem + ENC_OFF_QKV);
    // FFN hidden tile [plane hi/lo][kb 2] × 8 KB, DOUBLE-BUFFERED (FFN phase only): buffer 0 in the head-buffer region, buffer 1 in
    // the O tile — both idle during the FFN — so the epilogue of chunk c+1 writes while FFN2 still reads chunk c (with one buffer
    // the two were serialised: ≈3 k cycles per chunk instead of the MMAs' 3 k for FFN1 + FFN2 together)
    uint8_t* sH0 = smem + ENC_OFF_QKV;
    uint8_t* sH1 = smem + ENC_OFF_O;
    float* sSC = reinterpret_cast<float*>(smem + ENC_OFF_QKV);   // heads result (after the last layer only)
    uint8_t* sA = smem + ENC_OFF_A;
    uint8_t* sO = smem + ENC_OFF_O;
    float* sPrm = reinterpret_cast<float*>(smem + ENC_OFF_PRM);
    unsigned char* sTask = smem + ENC_OFF_TASK;
    float* sRed = reinterpret_cast<float*>(smem + ENC_OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ENC_OFF_BAR);
    uint64_t* w_full = bars;                        // [ENC_NSTAGE]
    uint64_t* w_empty = bars + ENC_NSTAGE;          // [ENC_NSTAGE]
    uint64_t* bb = bars + 2 * ENC_NSTAGE;
    uint64_t* a_ready = bb;
    uint64_t* dq_full = bb + 1;         // [2]
    uint64_t* dq_empty = bb + 3;        // [2]
    uint64_t* qkv_full = bb + 5;
    uint64_t* qkv_empty = bb + 6;
    uint64_t* o_ready = bb + 7;
    uint64_t* x_full = bb + 8;
    uint64_t* dh_full = bb + 9;         // [2]
    uint64_t* dh_empty = bb + 11;       // [2]
    uint64_t* h_ready = bb + 13;        // [2]
    uint64_t* h_free = bb + 16;         // [2]
    uint64_t* sc_full = bb + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bb + 18);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int wq = warp & 3;
    const int n_tiles = (p.B + ENC_IPT - 1) / ENC_IPT;
    const int T = p.B * 9;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ENC_NSTAGE; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        mbar_init(a_ready, 8);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&dq_full[i], 1); mbar_init(&dq_empty[i], 6);
            mbar_init(&dh_full[i], 1); mbar_init(&dh_empty[i], 8);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&h_ready[i], 8); mbar_init(&h_free[i], 1); }
        mbar_init(qkv_full, 6); mbar_init(qkv_empty, ENC_ATTN_WARPS); mbar_init(o_ready, 4 * ENC_ATTN_WARPS); mbar_init(x_full, 1); mbar_init(sc_full, 1);
        fence_barrier_init();
    }
    if (warp == ENC_W_MMA) tmem_alloc(tmem_slot, 512);
    // parameters are constants of the model: stage them before waiting for the previous kernel
    for (int l = 0; l < p.nl; ++l) {
        const EncLayerPtrs& lp = p.layer[l];
        float* d = sPrm + l * ENC_P_LAYER;
        for (int i = threadIdx.x; i < 576; i += ENC_THREADS) d[ENC_P_BIN + i] = __ldg(lp.b_in + i);
        for (int i = threadIdx.x; i < 384; i += ENC_THREADS) d[ENC_P_B1 + i] = __ldg(lp.b_1 + i);
        for (int i = threadIdx.x; i < 192; i += ENC_THREADS) {
            d[ENC_P_BO + i] = __ldg(lp.b_o + i); d[ENC_P_B2 + i] = __ldg(lp.b_2 + i);
            d[ENC_P_LN1W + i] = __ldg(lp.ln1w + i); d[ENC_P_LN1B + i] = __ldg(lp.ln1b + i);
            d[ENC_P_LN2W + i] = __ldg(lp.ln2w + i); d[ENC_P_LN2B + i] = __ldg(lp.ln2b + i);
        }
    }
    for (int i = threadIdx.x; i < 192; i += ENC_THREADS) sPrm[ENC_P_TOKB + i] = __ldg(p.tok_bias + i);
    if (threadIdx.x < 64) {
        sPrm[ENC_P_CLSB + threadIdx.x] = threadIdx.x < p.ncls_total ? __ldg(p.cls_bias + threadIdx.x) : 0.f;
        sTask[threadIdx.x] = p.task_of[threadIdx.x];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == ENC_W_PROD) {
        // ===================== weight producer =====================
        if (elect_one()) {      // one elected thread; elect.sync keeps the branch warp-uniform for ptxas
            uint32_t wi = 0;
            auto load = [&](size_t off, uint32_t bytes) {
                const int s = wi % ENC_NSTAGE;
                const uint32_t ph = (wi / ENC_NSTAGE) & 1;
                mbar_wait(&w_empty[s], ph ^ 1);
                mbar_expect_tx(&w_full[s], bytes);          // the whole block: this CTA's half + the peer's
                bulk_g2s(sW + s * ENC_STAGE, p.wstream + off, bytes, &w_full[s]);     // (issuing a block as 4 or 8 smaller bulk copies: measured, no change)
                ++wi;
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < p.nl; ++l) {
                    const size_t base = (size_t)l * ENC_LAYER_BYTES;
                    for (int i = 0; i < 12; ++i) load(base + ENC_L_QKV + (size_t)i * ENC_QKV_BLK, ENC_QKV_BLK);
                    for (int i = 0; i < 3; ++i) load(base + ENC_L_OUT + (size_t)i * ENC_STAGE, ENC_STAGE);
                    size_t off = base + ENC_L_FFN;
                    auto f1 = [&]() { for (int i = 0; i < 3; ++i, off += ENC_F1_BLK) load(off, ENC_F1_BLK); };
                    auto f2 = [&]() { for (int i = 0; i < 2; ++i, off += ENC_STAGE) load(off, ENC_STAGE); };
                    f1(); f1(); f2(); f1(); f2(); f2();
                }
                load((size_t)p.nl * ENC_LAYER_BYTES, ENC_STAGE);
            }
        }
    } else if (warp == ENC_W_MMA) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t id144 = make_idesc_bf16(128, 144, 0, 0);
            constexpr uint32_t id192 = make_idesc_bf16(128, 192, 0, 0);
            constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t id64 = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t aW = smem_u32(sW), aA = smem_u32(sA), aO = smem_u32(sO), aH0 = smem_u32(sH0), aH1 = smem_u32(sH1);
            const uint32_t tX = tmem_base, tS = tmem_base + 192;
            uint32_t wi = 0, na = 0, hq = 0, no = 0, hc = 0;
            // one weight k-block against one [64 × 64] hi/lo operand k-block: D (+)= (A_hi + A_lo) · Wᵀ
            auto mma_stage = [&](uint32_t a_hi, uint32_t a_lo, uint32_t d, uint32_t idesc, uint32_t w_off, bool first) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = make_desc_k_sw128(w_off + k * 32);
                    umma_bf16(d, make_desc_k_sw128(a_hi + k * 32), bd, idesc, (first && k == 0) ? 0u : 1u);
                    umma_bf16(d, make_desc_k_sw128(a_lo + k * 32), bd, idesc, 1u);
                }
            };
            long long w_wait = 0;                        // trace only: cycles this thread spent waiting for weight blocks
            auto wait_w = [&]() -> uint32_t {
                const int s = wi % ENC_NSTAGE;
                const long long t0 = p.trace ? clock64() : 0;
                mbar_wait(&w_full[s], (wi / ENC_NSTAGE) & 1);
                if (p.trace) w_wait += clock64() - t0;
                tc_fence_after();
                return aW + s * ENC_STAGE;
            };
            auto free_w = [&]() { umma_commit(&w_empty[wi % ENC_NSTAGE]); ++wi; };
            auto ffn2 = [&](uint32_t pc) {               // x += H(chunk pc) · W2[:, 128·chunk ..]ᵀ
                const uint32_t hb = pc & 1, aH = hb ? aH1 : aH0;
                mbar_wait(&h_ready[hb], (pc >> 1) & 1);
                tc_fence_after();
                for (int kb = 0; kb < 2; ++kb) {
                    const uint32_t w = wait_w();
                    mma_stage(aH + kb * ENC_KB, aH + 2 * ENC_KB + kb * ENC_KB, tX, id192, w, false);
                    free_w();
                }
                umma_commit(&h_free[hb]);
            };
            ENC_T(0);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < p.nl; ++l) {
                    mbar_wait(a_ready, na & 1); ++na;
                    tc_fence_after();
                    ENC_T(1 + 40 * l);
                    for (int h = 0; h < 4; ++h, ++hq) {                       // [q_h | k_h | v_h] = LN1(x) · W_hᵀ
                        const uint32_t b = hq & 1;
                        mbar_wait(&dq_empty[b], ((hq >> 1) & 1) ^ 1);
                        tc_fence_after();
                        for (int kb = 0; kb < 3; ++kb) {
                            const uint32_t w = wait_w();
                            mma_stage(aA + kb * ENC_KB, aA + ENC_PLANE + kb * ENC_KB, tS + 144 * b, id144, w, kb == 0);
                            free_w();
                        }
                        umma_commit(&dq_full[b]);
                        ENC_T(2 + 40 * l + h);
                    }
                    mbar_wait(o_ready, no & 1); ++no;                          // x += O · Woᵀ
                    tc_fence_after();
                    ENC_T(6 + 40 * l);
                    for (int kb = 0; kb < 3; ++kb) {
                        const uint32_t w = wait_w();
                        mma_stage(aO + kb * ENC_KB, aO + ENC_PLANE + kb * ENC_KB, tX, id192, w, false);
                        free_w();
                    }
                    umma_commit(x_full);
                    ENC_T(7 + 40 * l);
                    mbar_wait(a_ready, na & 1); ++na;                          // LN2(x) in the A tile
                    tc_fence_after();
                    ENC_T(8 + 40 * l);
                    for (int c = 0; c < 3; ++c, ++hc) {                        // hidden chunk c = LN2(x) · W1[128c..]ᵀ
                        const uint32_t b = hc & 1;
                        mbar_wait(&dh_empty[b], ((hc >> 1) & 1) ^ 1);
                        tc_fence_after();
                        for (int kb = 0; kb < 3; ++kb) {
                            const uint32_t w = wait_w();
                            mma_stage(aA + kb * ENC_KB, aA + ENC_PLANE + kb * ENC_KB, tS + 128 * b, id128, w, kb == 0);
                            free_w();
                        }
                        umma_commit(&dh_full[b]);
                        ENC_T(9 + 40 * l + c);
                        if (c >= 1) { ffn2(hc - 1); ENC_T(15 + 40 * l + c - 1); }
                    }
                    ffn2(hc - 1);
                    umma_commit(x_full);
                    ENC_T(20 + 40 * l);
                }
                // heads: [scores | class projections] = x · [Q | W_cls]ᵀ
                mbar_wait(a_ready, na & 1); ++na;
                tc_fence_after();
                {
                    const uint32_t w = wait_w();
#pragma unroll
                    for (int kb = 0; kb < 3; ++kb)
                        mma_stage(aA + kb * ENC_KB, aA + ENC_PLANE + kb * ENC_KB, tS, id64, w + kb * ENC_KB, kb == 0);
                    free_w();
                }
                umma_commit(sc_full);
            }
            if (p.trace && blockIdx.x == 0) p.trace[250] = w_wait;
        }
    } else if (warp >= ENC_W_CMP0 && wq < 2) {
        // ===================== compute warps: TMEM <-> registers, LayerNorm, epilogues =====================
        const int q = wq, g = (warp - ENC_W_CMP0) >> 2, m = q * 32 + lane;
        const int ct = g * 64 + m;                                  // 0..255
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t tX = tmem_base + lane_addr + 48 * g;         // this thread's slice of the residual stream
        const uint32_t tS = tmem_base + lane_addr + 192;
        uint32_t hq = 0, nx = 0, hc = 0, nsc = 0;
        float v[48];
        pdl_wait();                                                 // partial sums come from the token_proj kernel
        const bool tr = (warp == ENC_W_CMP0 && lane == 0);
#define ENC_TC(slot) do { if (tr) ENC_T(slot); } while (0)
        ENC_TC(100);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int grow = tile * ENC_ROWS + m;
            const bool valid = m < ENC_ROWS && grow < T;
            // ---- x0 = Σ_s partial[s] + bias (token_proj epilogue, models_PMG.py:193) ----
            // the 256 compute threads pull the tile's [63 × 192] fp32 block of two split-K partials at a time with
            // 16-byte cp.async into the (idle) O and A regions, chunk-swizzled per row; then every thread adds up
            // its own 48-column slice.  One global round trip, no registers held across it.
            {
#pragma unroll
                for (int i = 0; i < 48; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sPrm + ENC_P_TOKB + 48 * g + i);
                    v[i] = b4.x; v[i + 1] = b4.y; v[i + 2] = b4.z; v[i + 3] = b4.w;
                }
                for (int s0 = 0; s0 < p.nsplit; s0 += 2) {
                    const int nb = min(2, p.nsplit - s0);
#pragma unroll
                    for (int i = 0; i < 12; ++i) {
                        const int f = ct + 256 * i, r = f / 48, c4 = f - r * 48;
                        const int gr = tile * ENC_ROWS + r;
                        const bool ok = r < ENC_ROWS && gr < T;
                        const uint32_t so = (uint32_t)(r * 48 + (c4 ^ (r & 7))) * 16;
                        const float* src = p.partial + (size_t)s0 * p.pstride + (ok ? (size_t)gr * 192 + 4 * c4 : 0);
                        cp_async16(sO + so, src, ok ? 16u : 0u);
                        if (nb == 2) cp_async16(sA + so, src + p.pstride, ok ? 16u : 0u);
                    }
                    cp_async_wait_all();
                    named_bar_sync(1, 256);
#pragma unroll
                    for (int j = 0; j < 12; ++j) {
                        const uint32_t so = (uint32_t)(m * 48 + ((12 * g + j) ^ (m & 7))) * 16;
                        const float4 a = *reinterpret_cast<const float4*>(sO + so);
                        v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        if (nb == 2) {
                            const float4 c = *reinterpret_cast<const float4*>(sA + so);
                            v[4 * j] += c.x; v[4 * j + 1] += c.y; v[4 * j + 2] += c.z; v[4 * j + 3] += c.w;
                        }
                    }
                    if (s0 + 2 < p.nsplit) named_bar_sync(1, 256);   // staging is reused by the next pair of splits
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 48; ++i) v[i] = 0.f;
                } else if (p.tok_in) {
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        *reinterpret_cast<float4*>(p.tok_in + (size_t)grow * 192 + 48 * g + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            ENC_TC(101);
            if (p.nl > 0) tmem_st48(tX, v);
            for (int l = 0; l < p.nl; ++l) {
                const float* lp = sPrm + l * ENC_P_LAYER;
                // ---- LN1 → A tile ----
                layer_norm48(v, sRed, m, g, lp + ENC_P_LN1W + 48 * g, lp + ENC_P_LN1B + 48 * g);      // x itself stays in TMEM
                write_operand48(sA, m, g, v);
                tmem_st_wait();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
                ENC_TC(102 + 40 * l);
                // ---- per head: drain [q|k|v] (+bias) to the fp32 head buffer; g = 0/1/2 ↔ q/k/v ----
                for (int h = 0; h < 4; ++h, ++hq) {
                    if (g < 3) {
                        const uint32_t b = hq & 1;
                        mbar_wait(&dq_full[b], (hq >> 1) & 1);
                        tc_fence_after();
                        tmem_ld48(tS + 144 * b + 48 * g, v);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&dq_empty[b]);
                        const float* bias = lp + ENC_P_BIN + 192 * g + 48 * h;
                        // softmax runs in the exp2 domain: fold log2(e)/sqrt(48) into q
                        const float scl = (g == 0) ? 0.20823509638153327f : 1.f;
                        mbar_wait(qkv_empty, (hq & 1) ^ 1);                           // attention of the previous head is done reading
                        float* dst = sQKV + m * ENC_QKV_LD + 48 * g;
#pragma unroll
                        for (int j = 0; j < 12; ++j) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j);
                            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4((v[4 * j] + b4.x) * scl, (v[4 * j + 1] + b4.y) * scl,
                                                                                  (v[4 * j + 2] + b4.z) * scl, (v[4 * j + 3] + b4.w) * scl);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(qkv_full);
                        ENC_TC(103 + 40 * l + h);
                    }
                }
                // ---- x += attention block done: + out-proj bias, LN2 → A tile ----
                mbar_wait(x_full, nx & 1); ++nx;
                tc_fence_after();
                ENC_TC(107 + 40 * l);
                tmem_ld48(tX, v);
                add_bias48(v, lp + ENC_P_BO + 48 * g);
                tmem_st48(tX, v);
                layer_norm48(v, sRed, m, g, lp + ENC_P_LN2W + 48 * g, lp + ENC_P_LN2B + 48 * g);
                write_operand48(sA, m, g, v);
                tmem_st_wait();
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
                ENC_TC(108 + 40 * l);
                // ---- FFN hidden chunks (128 units each): + bias, ReLU, hi/lo → H tile ----
                for (int c = 0; c < 3; ++c, ++hc) {
                    const uint32_t b = hc & 1;
                    mbar_wait(&dh_full[b], (hc >> 1) & 1);
                    tc_fence_after();
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tS + 128 * b + 32 * g, r);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&dh_empty[b]);
                    uint32_t hi[16], lo[16];
                    const float* b1 = lp + ENC_P_B1 + 128 * c + 32 * g;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(b1 + i);
                        const float h0 = fmaxf(__uint_as_float(r[i]) + b4.x, 0.f), h1 = fmaxf(__uint_as_float(r[i + 1]) + b4.y, 0.f);
                        const float h2 = fmaxf(__uint_as_float(r[i + 2]) + b4.z, 0.f), h3 = fmaxf(__uint_as_float(r[i + 3]) + b4.w, 0.f);
                        split_bf16x2(h0, h1, hi[i >> 1], lo[i >> 1]);
                        split_bf16x2(h2, h3, hi[(i >> 1) + 1], lo[(i >> 1) + 1]);
                    }
                    mbar_wait(&h_free[b], ((hc >> 1) & 1) ^ 1);                        // FFN2 of the chunk two back has read this H buffer
                    uint8_t* hrow = (b ? sH1 : sH0) + (g >> 1) * ENC_KB + m * 128;     // hidden columns 32g.. → k-block g/2, chunks 4(g%2)..+3
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sw = ((4 * (g & 1) + j) ^ (m & 7)) << 4;
                        *reinterpret_cast<uint4*>(hrow + sw) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<uint4*>(hrow + 2 * ENC_KB + sw) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_ready[b]);
                    ENC_TC(109 + 40 * l + c);
                }
                // ---- x += FFN done: + bias ----
                mbar_wait(x_full, nx & 1); ++nx;
                tc_fence_after();
                ENC_TC(115 + 40 * l);
                tmem_ld48(tX, v);
                add_bias48(v, lp + ENC_P_B2 + 48 * g);
                if (l + 1 < p.nl) tmem_st48(tX, v);
            }
            // ---- final tokens → heads operand ----
            if (p.tok_out && valid) {
#pragma unroll
                for (int j = 0; j < 12; ++j)
                    *reinterpret_cast<float4*>(p.tok_out + (size_t)grow * 192 + 48 * g + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            write_operand48(sA, m, g, v);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            mbar_wait(sc_full, nsc & 1); ++nsc;
            tc_fence_after();
            ENC_TC(190);
            {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tS + 16 * g, r);
                tmem_ld_wait();
                tc_fence_before();
#pragma unroll
                for (int i = 0; i < 16; ++i) sSC[m * ENC_SC_LD + 16 * g + i] = __uint_as_float(r[i]);
            }
            named_bar_sync(1, 256);
            // ---- task attention pooling + heads (models_PMG.py:47-57,198): one (image, class) per thread ----
            for (int idx = ct; idx < ENC_IPT * p.ncls_total; idx += 256) {
                const int bi = idx / p.ncls_total, c = idx - bi * p.ncls_total, t = sTask[c];
                const int img = tile * ENC_IPT + bi;
                if (img >= p.B) continue;
                const float* s = sSC + bi * 9 * ENC_SC_LD;
                float sv[9], mx = -INFINITY;
#pragma unroll
                for (int n = 0; n < 9; ++n) { sv[n] = s[n * ENC_SC_LD + t]; mx = fmaxf(mx, sv[n]); }
                float den = 0.f, acc = 0.f;
#pragma unroll
                for (int n = 0; n < 9; ++n) {
                    const float e = expf(sv[n] - mx);
                    den += e;
                    acc = fmaf(e, s[n * ENC_SC_LD + p.ntask + c], acc);
                }
                const float lg = acc / den + sPrm[ENC_P_CLSB + c];
                const size_t li = (size_t)img * p.ncls_total + c;
                p.logits[li] = lg;
                for (int d = 0; d < p.n_peers; ++d) p.peer_logits[d][li] = lg;      // posted stores over NVLink: the exchange rides in this kernel
            }
            named_bar_sync(1, 256);
            ENC_TC(191);
        }
        tc_fence_before();
    } else if (warp >= 2 && wq >= 2) {
        // ===================== all eight aux warps: the attention core =====================
        const AttnCtx1 ac = attn_ctx1(sQKV, (warp >> 2) * 2 + (wq - 2), lane);
        const bool tr = warp == 2 && lane == 0;
        uint32_t hq = 0;
        float o[12];
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int l = 0; l < p.nl; ++l)
                for (int h = 0; h < 4; ++h, ++hq) {
                    mbar_wait(qkv_full, hq & 1);
                    if (tr) ENC_T(200 + 40 * l + 2 * h);
                    attention_rows(ac, o, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(qkv_empty);        // every lane of this warp has finished reading the head buffer
                    store_o_rows(ac, sO, h, o);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_ready);
                    if (tr) ENC_T(201 + 40 * l + 2 * h);
                }
    }
    __syncthreads();
    if (warp == ENC_W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// weight stream packing: fp32 (or bf16) [rows][ld] matrices → bf16 blocks of [nrows × 64 K], rows 128 bytes,
// 16-byte chunks XOR-swizzled with (row & 7) exactly as the UMMA SWIZZLE_128B K-major layout expects.
// ------------------------------------------------------------------------------------------------
struct EncPackDesc {
    const void* src;
    int src_bf16;        // 0: fp32 source, 1: bf16 source
    int ld;
    int row0[3];         // source row of segment i = block row / seg
    int seg;
    int nrows;
    int k0;
    unsigned dst;        // byte offset in the stream
};
struct EncPackArgs {
    EncPackDesc d[40];
};

__global__ void __launch_bounds__(256) enc_pack_kernel(const EncPackArgs a, uint8_t* __restrict__ stream) {
    const EncPackDesc& d = a.d[blockIdx.x];
    for (int i = threadIdx.x; i < d.nrows * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        const size_t so = (size_t)(d.row0[r / d.seg] + r % d.seg) * d.ld + d.k0 + 8 * j;
        uint32_t w[4];
        if (d.src_bf16) {
            const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.src) + so);
            w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        } else {
            const float* s = reinterpret_cast<const float*>(d.src) + so;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(s[2 * e], s[2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h);
            }
        }
        *reinterpret_cast<uint4*>(stream + d.dst + r * 128 + ((j ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace tc
