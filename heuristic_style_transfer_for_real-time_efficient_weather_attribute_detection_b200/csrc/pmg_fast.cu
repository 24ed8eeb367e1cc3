// pmg_fast.cu — tcgen05 / TMA fast path of the PMG forward for the committed configuration
// (hyperparameters_PMG.json: ndf 32, patch_size 64, patch_div 3, gram_channels 96, d_model 192, 4 heads)
// at 3×224×224.  Reference: /root/reference/Models/models_PMG.py:176-199.
//
// Every GEMM-shaped op runs on the 5th-gen tensor cores through tc::gemm_kernel (TMA → 128B-swizzled
// smem → tcgen05.mma → TMEM → epilogue):
//   conv0  [B·3136, 48]  × [48, 32]    (k=s=4 conv as space-to-depth GEMM; im2col rows ordered so that the
//                                        output IS conv1's A operand)                         models_PMG.py:115
//   conv1  [B·196, 512]  × [512, 64]   + BN + LeakyReLU                                       :115-117
//   conv2  [B·196, 64]   × [64, 128]   + BN + LeakyReLU                                       :125-127
//   chan   [B·196, 128]  × [128, 96]                                                          :136,182
//   token  [B·9, 6144]   × [6144, 192] split-K (9216 folded to the 6 upper 32×32 Gram blocks) :141,193
//   encoder linears (QKV, out-proj, FFN)                                                      :64-69,194
// Activations travel between kernels as two bf16 planes (hi + lo, ~16 significant bits) so the result
// stays within the 1e-3 parity gate with wide margin (DESIGN.md "precision").  Per-patch Gram, LayerNorm,
// the 9-token attention core and the task pooling + heads are small CUDA-core kernels.
#include "pmg_common.cuh"
#include "tc_host.cuh"
#include "tc_gram.cuh"
#include "tc_front.cuh"
#include "tc_encoder.cuh"
#include "tc_mid.cuh"
#include "tc_midgram.cuh"
#include "tc_gramb.cuh"

namespace pmg {

long long* g_enc_trace = nullptr;          // debug only (pmg_debug_set_trace); pm_path.cu reads it too

int launch_pool_heads(const Plan& p, const PackedLayout& L, const char* packed, const float* tok, int B, int N,
                      float* logits, cudaStream_t st);

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------------
// layout of the fast section of the packed weights (bf16, [N][K] K-contiguous = UMMA "K-major" B operand)
// ------------------------------------------------------------------------------------------------
struct FastW {
    size_t w0, w1, w2, w3;              // conv0 [32][64] (K 48 zero-padded to 64), conv1 [64][512] (K = (sy,sx,c)), conv2 [128][64], chan [96][128]
    size_t wth, wtl;                     // token_proj folded over the Gram's symmetry: hi [192][6144], lo [192][3072] (off-diagonal blocks)
    size_t win[8], wo[8], wf1[8], wf2[8];
    size_t w0x, w1x;                     // fused front end: conv0 for 8-column units (pack_w0u_kernel, 12 KB), conv1 as (sy,kb) blocks [512][64]
    size_t wq;                           // [64][192]: rows 0..ntask-1 task queries, then the ncls_total classifier rows, zero pad
    size_t wenc;                         // fused token stage: pre-swizzled weight stream (tc_encoder.cuh), nl layers + heads block
    size_t total;
};

static FastW fast_w(const Plan& p) {
    FastW f{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o = align256(o + n * sizeof(bf16)); return r; };
    f.w0 = take(32 * 64); f.w1 = take(64 * 512); f.w2 = take(128 * 64); f.w3 = take(96 * 128);
    f.wth = take((size_t)192 * 6144); f.wtl = take((size_t)192 * 3072);
    for (int l = 0; l < p.nl && l < 8; ++l) {
        f.win[l] = take(576 * 192); f.wo[l] = take(192 * 192); f.wf1[l] = take(384 * 192); f.wf2[l] = take(192 * 384);
    }
    f.wq = take(64 * 192);
    f.w0x = take(6 * 2 * 64 * 8); f.w1x = take(512 * 64);
    f.wenc = take(((size_t)(p.nl <= tc::ENC_MAX_LAYERS ? p.nl : 0) * tc::ENC_LAYER_BYTES + tc::ENC_STAGE) / sizeof(bf16));
    f.total = o;
    return f;
}

static bool is_committed(const Plan& p) {
    return p.ok && p.norm == PMG_NORM_BATCH && !p.se && p.n_layers == 3 && p.layer[0].cin == 3 && p.layer[0].cout == 32 && p.layer[0].k == 4 &&
           p.layer[1].cout == 64 && p.layer[1].k == 4 && p.layer[2].cout == 128 && p.layer[2].k == 1 && p.Cr == 96 &&
           p.d == 192 && p.nh == 4 && p.pdiv == 3 && p.nl <= 8 && p.ntask + p.ncls_total <= 64;
}

size_t fast_section_bytes(const Plan& p) { return is_committed(p) ? fast_w(p).total : 0; }

bool fast_supported(const Plan& p, const Shape& s) { return is_committed(p) && s.h[0] == 224 && s.w[0] == 224; }

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
__global__ void cvt_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
// conv1 weight (64, 32, 4, 4) → [64][(ky*4+kx)*32 + c]: conv1's kernel offset (ky,kx) is the conv0 sub-position
__global__ void pack_w1_kernel(const float* __restrict__ in, bf16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 512) return;
    const int co = i / 512, k = i % 512, s = k / 32, c = k % 32;
    out[i] = __float2bfloat16_rn(in[(co * 32 + c) * 16 + s]);
}

// conv0 weight (32, 3, 4, 4) → [32][64]: k = c*16+ky*4+kx for k < 48, zero for the 16 pad columns
__global__ void pack_w0_kernel(const float* __restrict__ in, bf16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * 64) return;
    const int co = i / 64, k = i % 64;
    out[i] = __float2bfloat16_rn(k < 48 ? in[co * 48 + k] : 0.f);
}

// fused front end, conv0 weight for 8-column units (tc_front.cuh): an MMA row is a 16-byte unit of an image row = two conv0
// pixels xp = 0, 1; N = 64 = (xp, co), block-sparse over xp.  Stored exactly as the no-swizzle K-major operand the kernel's
// descriptors walk: K step j = (c, ky pair) | K chunk ck (ky = 2·(j&1) + ck) | 8-row core matrices of 8 rows × 16 bytes:
//   element e of row n = xp·32 + co  =  (e >> 2 == xp) ? w0[co][c][ky][e & 3] : 0
__global__ void pack_w0u_kernel(const float* __restrict__ in, bf16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * 2 * 64 * 8) return;
    const int e = i & 7, n = (i >> 3) & 63, ck = (i >> 9) & 1, j = i >> 10;
    const int c = j >> 1, ky = 2 * (j & 1) + ck, xp = n >> 5, co = n & 31;
    out[i] = __float2bfloat16_rn((e >> 2) == xp ? in[((co * 3 + c) * 4 + ky) * 4 + (e & 3)] : 0.f);
}
// fused front end, conv1 weight as 8 blocks (sy, kb): [blk*64 + n][col = (sx&1)*32 + c] = w1[n][c][sy][sx], sx = kb*2 + col/32
__global__ void pack_w1x_kernel(const float* __restrict__ in, bf16* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 512 * 64) return;
    const int row = i >> 6, col = i & 63, blk = row >> 6, n = row & 63, sy = blk >> 1, kb = blk & 1;
    const int sx = kb * 2 + (col >> 5), c = col & 31;
    out[i] = __float2bfloat16_rn(in[((n * 32 + c) * 4 + sy) * 4 + sx]);
}

// token_proj weight (192, 96·96) folded over G = Gᵀ.  The Gram kernel stores only the 6 upper 32×32 blocks of each
// 96×96 matrix, K' = blk·1024 + il·32 + jl with blk order (0,0) (1,1) (2,2) | (0,1) (0,2) (1,2):
//   diagonal block (b,b):      W'[n][K'] = W[n][(32b+il)·96 + 32b+jl]                                   (exact in bf16)
//   off-diagonal (bi < bj):    W'[n][K'] = W[n][(32bi+il)·96 + 32bj+jl] + W[n][(32bj+jl)·96 + 32bi+il]   (hi + lo planes)
__global__ void pack_wt_folded_kernel(const float* __restrict__ in, bf16* __restrict__ hi, bf16* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 192 * 6144) return;
    const int n = i / 6144, k = i % 6144, blk = k >> 10, il = (k >> 5) & 31, jl = k & 31;
    const float* w = in + (size_t)n * 9216;
    if (blk < 3) {
        hi[i] = __float2bfloat16_rn(w[(32 * blk + il) * 96 + 32 * blk + jl]);
    } else {
        const int bi = blk == 5 ? 1 : 0, bj = blk == 3 ? 1 : 2;
        const float v = w[(32 * bi + il) * 96 + 32 * bj + jl] + w[(32 * bj + jl) * 96 + 32 * bi + il];
        const bf16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[(size_t)n * 3072 + (k - 3072)] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

int pack_fast(const Plan& p, const PackedLayout& L, const float* const* prm, char* packed, cudaStream_t st) {
    if (!is_committed(p)) return PMG_OK;
    const FastW f = fast_w(p);
    char* base = packed + L.fast;
    auto Bp = [&](size_t off) { return reinterpret_cast<bf16*>(base + off); };
    auto cvt = [&](const float* src, size_t off, size_t n) {
        size_t g = (n + 255) / 256;
        cvt_bf16_kernel<<<(int)(g > 2048 ? 2048 : g), 256, 0, st>>>(src, Bp(off), n);
    };
    pack_w0_kernel<<<8, 256, 0, st>>>(prm[0], Bp(f.w0));           // (32,3,4,4) == [32][c*16+ky*4+kx], K padded to 64
    pack_w1_kernel<<<(64 * 512 + 255) / 256, 256, 0, st>>>(prm[5], Bp(f.w1));
    pack_w0u_kernel<<<(6144 + 255) / 256, 256, 0, st>>>(prm[0], Bp(f.w0x));
    pack_w1x_kernel<<<(512 * 64 + 255) / 256, 256, 0, st>>>(prm[5], Bp(f.w1x));
    cvt(prm[10], f.w2, 128 * 64);
    cvt(prm[15], f.w3, 96 * 128);
    pack_wt_folded_kernel<<<(192 * 6144 + 255) / 256, 256, 0, st>>>(prm[16], Bp(f.wth), Bp(f.wtl));
    int i = 18;
    for (int l = 0; l < p.nl; ++l) {
        cvt(prm[i], f.win[l], 576 * 192);
        cvt(prm[i + 2], f.wo[l], 192 * 192);
        cvt(prm[i + 4], f.wf1[l], 384 * 192);
        cvt(prm[i + 6], f.wf2[l], 192 * 384);
        i += 12;
    }
    // heads as one GEMM operand: [queries | classifier rows] (models_PMG.py:45,153-155)
    PMG_CUDA_OK(cudaMemsetAsync(Bp(f.wq), 0, 64 * 192 * sizeof(bf16), st), "wq memset");
    for (int t = 0; t < p.ntask; ++t) cvt(prm[i + t], f.wq + (size_t)t * 192 * sizeof(bf16), 192);
    i += p.ntask;
    for (int t = 0; t < p.ntask; ++t) {
        const int n = p.cls_off[t + 1] - p.cls_off[t];
        cvt(prm[i], f.wq + (size_t)(p.ntask + p.cls_off[t]) * 192 * sizeof(bf16), (size_t)n * 192);
        i += 2;
    }
    // fused token stage: every encoder weight as pre-swizzled [rows × 64 K] blocks in the order the kernel's
    // MMA warp consumes them (tc_encoder.cuh), then the heads operand
    if (p.nl <= tc::ENC_MAX_LAYERS) {
        uint8_t* stream = reinterpret_cast<uint8_t*>(base + f.wenc);
        static const int ffn_order[6][2] = {{1, 0}, {1, 1}, {2, 0}, {1, 2}, {2, 1}, {2, 2}};   // FFN1 / FFN2 of 128-unit chunk c
        int pi = 18;
        for (int l = 0; l < p.nl; ++l, pi += 12) {
            tc::EncPackArgs a{};
            int n = 0;
            const unsigned lb = (unsigned)((size_t)l * tc::ENC_LAYER_BYTES);
            auto add = [&](const void* src, int bf, int ld, int r0, int r1, int r2, int seg, int nrows, int k0, unsigned dst) {
                tc::EncPackDesc& d = a.d[n++];
                d.src = src; d.src_bf16 = bf; d.ld = ld; d.row0[0] = r0; d.row0[1] = r1; d.row0[2] = r2; d.seg = seg;
                d.nrows = nrows; d.k0 = k0; d.dst = dst;
            };
            for (int h = 0; h < 4; ++h)
                for (int kb = 0; kb < 3; ++kb)
                    add(prm[pi], 0, 192, 48 * h, 192 + 48 * h, 384 + 48 * h, 48, 144, 64 * kb, lb + tc::ENC_L_QKV + (h * 3 + kb) * tc::ENC_QKV_BLK);
            for (int kb = 0; kb < 3; ++kb) add(prm[pi + 2], 0, 192, 0, 0, 0, 192, 192, 64 * kb, lb + tc::ENC_L_OUT + kb * tc::ENC_STAGE);
            unsigned off = lb + tc::ENC_L_FFN;
            for (int i = 0; i < 6; ++i) {
                const int c = ffn_order[i][1];
                if (ffn_order[i][0] == 1)
                    for (int kb = 0; kb < 3; ++kb, off += tc::ENC_F1_BLK) add(prm[pi + 4], 0, 192, 128 * c, 0, 0, 128, 128, 64 * kb, off);
                else
                    for (int kb = 0; kb < 2; ++kb, off += tc::ENC_STAGE) add(prm[pi + 6], 0, 384, 0, 0, 0, 192, 192, 128 * c + 64 * kb, off);
            }
            tc::enc_pack_kernel<<<n, 256, 0, st>>>(a, stream);
        }
        tc::EncPackArgs a{};
        for (int kb = 0; kb < 3; ++kb) {
            tc::EncPackDesc& d = a.d[kb];
            d.src = Bp(f.wq); d.src_bf16 = 1; d.ld = 192; d.row0[0] = d.row0[1] = d.row0[2] = 0; d.seg = 64; d.nrows = 64; d.k0 = 64 * kb;
            d.dst = (unsigned)((size_t)p.nl * tc::ENC_LAYER_BYTES) + kb * tc::ENC_KB;
        }
        tc::enc_pack_kernel<<<3, 256, 0, st>>>(a, stream);
    }
    PMG_LAUNCH_OK("pack_fast");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// im2col for conv0: x (B,3,224,224) NCHW → A0 [B·3136][64] bf16 (cols 48..63 zero), row r = (b·196 + py·14 + px)·16 + sy·4 + sx,
// col = c·16 + ky·4 + kx  (pixel y = 16py+4sy+ky, x = 16px+4sx+kx).  One thread = one 16-byte chunk
// (8 bf16 = two kernel rows ky, ky+1 of one channel), so a warp writes 512 contiguous bytes.
// fp32 input is split into hi/lo planes; bf16 input is copied.
// ------------------------------------------------------------------------------------------------
template <bool F32IN>
__global__ void __launch_bounds__(256) im2col0_kernel(const void* __restrict__ x, bf16* __restrict__ a_hi,
                                                      bf16* __restrict__ a_lo, int B) {
    tc::pdl_wait();
    tc::pdl_trigger();
    const size_t total = (size_t)B * 3136 * 8;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t r = t >> 3;
        const int j = (int)(t & 7);
        if (j >= 6) {                                   // K padding 48 → 64: TMA rows stay 128-byte, fully in bounds
            *reinterpret_cast<uint4*>(a_hi + r * 64 + j * 8) = make_uint4(0, 0, 0, 0);
            if (F32IN) *reinterpret_cast<uint4*>(a_lo + r * 64 + j * 8) = make_uint4(0, 0, 0, 0);
            continue;
        }
        const int s = (int)(r & 15);
        const size_t bp = r >> 4;
        const int pix = (int)(bp % 196);
        const size_t b = bp / 196;
        const int py = pix / 14, px = pix % 14, sy = s >> 2, sx = s & 3;
        const int c = j >> 1, ky = (j & 1) * 2;
        const size_t src = ((b * 3 + c) * 224 + (16 * py + 4 * sy + ky)) * 224 + 16 * px + 4 * sx;
        uint32_t hi[4], lo[4];
        if (F32IN) {
            const float* xf = reinterpret_cast<const float*>(x);
            const float4 r0 = *reinterpret_cast<const float4*>(xf + src);
            const float4 r1 = *reinterpret_cast<const float4*>(xf + src + 224);
            tc::split_bf16x2(r0.x, r0.y, hi[0], lo[0]); tc::split_bf16x2(r0.z, r0.w, hi[1], lo[1]);
            tc::split_bf16x2(r1.x, r1.y, hi[2], lo[2]); tc::split_bf16x2(r1.z, r1.w, hi[3], lo[3]);
            *reinterpret_cast<uint4*>(a_lo + r * 64 + j * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        } else {
            const bf16* xb = reinterpret_cast<const bf16*>(x);
            const uint2 r0 = *reinterpret_cast<const uint2*>(xb + src);
            const uint2 r1 = *reinterpret_cast<const uint2*>(xb + src + 224);
            hi[0] = r0.x; hi[1] = r0.y; hi[2] = r1.x; hi[3] = r1.y;
        }
        *reinterpret_cast<uint4*>(a_hi + r * 64 + j * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 NCHW image (what the reference's DataLoader delivers, Functions/PMG_evaluation.py:126) → two bf16 NCHW planes
// hi = bf16(x), lo = bf16(x − hi): the fused front kernel reads them as 3 + 3 channels against the same conv0 weights
// (conv0 is linear in the image), so the fp32 entry format keeps ~16 significant bits and never touches im2col.
// One thread = 8 pixels: 32 B in, 16 B + 16 B out.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_image_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, size_t n8) {
    tc::pdl_wait();
    tc::pdl_trigger();
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n8; t += (size_t)gridDim.x * blockDim.x) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(x) + 2 * t);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(x) + 2 * t + 1);
        uint32_t h[4], l[4];
        tc::split_bf16x2(a.x, a.y, h[0], l[0]); tc::split_bf16x2(a.z, a.w, h[1], l[1]);
        tc::split_bf16x2(b.x, b.y, h[2], l[2]); tc::split_bf16x2(b.z, b.w, h[3], l[3]);
        reinterpret_cast<uint4*>(hi)[t] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(lo)[t] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// token reduce (+bias) and LayerNorm → hi/lo planes.  One warp per token, d = 192 (6 values per lane).
//   nsplit > 0: tok = Σ_s partial[s] + bias is written to tok_out first (token_proj epilogue)
//   nsplit = 0: tok_out already holds the residual stream
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tok_ln_kernel(const float* __restrict__ partial, int nsplit, size_t pstride,
                                                     const float* __restrict__ bias, float* __restrict__ tok,
                                                     const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                     bf16* __restrict__ h_hi, bf16* __restrict__ h_lo, int T, int do_ln) {
    tc::pdl_wait();
    tc::pdl_trigger();
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= T) return;
    float v[6];
    if (nsplit > 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = 2 * lane + 64 * j;
            float2 a = make_float2(bias[c], bias[c + 1]);
            for (int s = 0; s < nsplit; ++s) {
                const float2 t = *reinterpret_cast<const float2*>(partial + (size_t)s * pstride + (size_t)row * 192 + c);
                a.x += t.x; a.y += t.y;
            }
            v[2 * j] = a.x; v[2 * j + 1] = a.y;
            *reinterpret_cast<float2*>(tok + (size_t)row * 192 + c) = a;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float2 t = *reinterpret_cast<const float2*>(tok + (size_t)row * 192 + 2 * lane + 64 * j);
            v[2 * j] = t.x; v[2 * j + 1] = t.y;
        }
    }
    if (!do_ln) return;
    if (do_ln == 2) {                       // raw residual stream as hi/lo planes (operand of the heads GEMM)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = 2 * lane + 64 * j;
            uint32_t hi, lo;
            tc::split_bf16x2(v[2 * j], v[2 * j + 1], hi, lo);
            *reinterpret_cast<uint32_t*>(h_hi + (size_t)row * 192 + c) = hi;
            *reinterpret_cast<uint32_t*>(h_lo + (size_t)row * 192 + c) = lo;
        }
        return;
    }
    float s = v[0] + v[1] + v[2] + v[3] + v[4] + v[5];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * (1.f / 192.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) { const float d = v[j] - mu; q += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.f / sqrtf(q * (1.f / 192.f) + kLnEps);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = 2 * lane + 64 * j;
        const float y0 = (v[2 * j] - mu) * rs * lnw[c] + lnb[c];
        const float y1 = (v[2 * j + 1] - mu) * rs * lnw[c + 1] + lnb[c + 1];
        uint32_t hi, lo;
        tc::split_bf16x2(y0, y1, hi, lo);
        *reinterpret_cast<uint32_t*>(h_hi + (size_t)row * 192 + c) = hi;
        *reinterpret_cast<uint32_t*>(h_lo + (size_t)row * 192 + c) = lo;
    }
}

// ------------------------------------------------------------------------------------------------
// Attention core for 9 tokens × 4 heads × 48: softmax(q·kᵀ/√48)·v, one block per image (4 warps = heads).
// qkv fp32 (B·9, 576) → out hi/lo (B·9, 192).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn9_kernel(const float* __restrict__ qkv, bf16* __restrict__ o_hi,
                                                    bf16* __restrict__ o_lo) {
    constexpr int LD = 580;                 // 576 + 4: token rows land in different banks (576 ≡ 0 mod 32)
    __shared__ __align__(16) float s[9 * LD];
    __shared__ float sp[4][9][12];
    tc::pdl_wait();
    tc::pdl_trigger();
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < 9 * 144; e += 128) {
        const int t = e / 144, c4 = e % 144;
        *reinterpret_cast<float4*>(s + t * LD + c4 * 4) = reinterpret_cast<const float4*>(qkv + ((size_t)b * 9 + t) * 576)[c4];
    }
    __syncthreads();
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // scores: 81 (i,j) pairs per head over 32 lanes
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        const float* q = s + i * LD + h * 48;
        const float* kk = s + j * LD + 192 + h * 48;
        float a = 0.f;
#pragma unroll 8
        for (int c = 0; c < 48; ++c) a = fmaf(q[c], kk[c], a);
        sp[h][i][j] = a * 0.14433756729740643f;   // 1/sqrt(48)
    }
    __syncwarp();
    if (lane < 9) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 9; ++j) mx = fmaxf(mx, sp[h][lane][j]);
        float e[9], den = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) { e[j] = expf(sp[h][lane][j] - mx); den += e[j]; }
        const float inv = 1.f / den;
#pragma unroll
        for (int j = 0; j < 9; ++j) sp[h][lane][j] = e[j] * inv;
    }
    __syncwarp();
    // out[i][c] for c pairs: 9 × 24 pairs per head
    for (int e = lane; e < 9 * 24; e += 32) {
        const int i = e / 24, c = (e % 24) * 2;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float w = sp[h][i][j];
            const float2 v = *reinterpret_cast<const float2*>(s + j * LD + 384 + h * 48 + c);
            a0 = fmaf(w, v.x, a0); a1 = fmaf(w, v.y, a1);
        }
        uint32_t hi, lo;
        tc::split_bf16x2(a0, a1, hi, lo);
        const size_t o = ((size_t)b * 9 + i) * 192 + h * 48 + c;
        *reinterpret_cast<uint32_t*>(o_hi + o) = hi;
        *reinterpret_cast<uint32_t*>(o_lo + o) = lo;
    }
}

// ------------------------------------------------------------------------------------------------
// Task pooling + heads, finishing step (models_PMG.py:47-57,198).  The heads GEMM produced, per token n,
//   sc[n][t]          = tok_n · q_t                 (t < ntask)         — the pooling scores
//   sc[n][ntask + c]  = W_c · tok_n                 (c < ncls_total)    — every class projected per token
// and  logits[c] = Σ_n softmax_n(sc[:,t(c)]) · sc[n][ntask+c] + b_c  because the head is linear in the pooled token.
// One thread per (image, class); 9 tokens.
// ------------------------------------------------------------------------------------------------
struct FinishArgs {
    const float* sc; const float* bias; float* logits;
    int B, ntask, ncls_total;
    unsigned char task_of[64];
};

__global__ void __launch_bounds__(256) pool_finish_kernel(const FinishArgs a) {
    tc::pdl_wait();
    tc::pdl_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.ncls_total) return;
    const int b = idx / a.ncls_total, c = idx % a.ncls_total, t = a.task_of[c];
    const float* s = a.sc + (size_t)b * 9 * 64;
    float sv[9], mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < 9; ++n) { sv[n] = s[n * 64 + t]; mx = fmaxf(mx, sv[n]); }
    float den = 0.f, acc = 0.f;
#pragma unroll
    for (int n = 0; n < 9; ++n) {
        const float e = expf(sv[n] - mx);
        den += e;
        acc = fmaf(e, s[n * 64 + a.ntask + c], acc);
    }
    a.logits[idx] = acc / den + a.bias[c];
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct FastWs {
    size_t a0_hi, a0_lo, act0_hi, act0_lo, act1_hi, act1_lo, act2_hi, act2_lo, chan, g_hi, g_lo, partial, tok,
        h_hi, h_lo, qkv, ffh_hi, ffh_lo, sc, total;
    int splits;                          // split-K factor of token_proj (hi weights K 6144 + lo-weight segment K 3072 in the same launch)
};

static int token_splits(int T) {
    const int tiles = (T + 127) / 128;
    int s = 148 / tiles;
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    return s;
}

static FastWs fast_ws(int B) {
    FastWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    const size_t P = (size_t)B * 196, T = (size_t)B * 9;
    w.a0_hi = take(P * 16 * 64 * 2); w.a0_lo = take(P * 16 * 64 * 2);
    w.act0_hi = take(P * 512 * 2); w.act0_lo = take(P * 512 * 2);
    w.act1_hi = take(P * 64 * 2); w.act1_lo = take(P * 64 * 2);
    w.act2_hi = take(P * 128 * 2); w.act2_lo = take(P * 128 * 2);
    w.chan = take(P * 96 * 4);
    w.g_hi = take(T * 6144 * 2); w.g_lo = take(T * 6144 * 2);
    w.splits = tc::effective_splits(6144, token_splits((int)T), 3072);
    w.partial = take((size_t)w.splits * T * 192 * 4);
    w.tok = take(T * 192 * 4);
    w.h_hi = take(T * 192 * 2); w.h_lo = take(T * 192 * 2);
    w.qkv = take(T * 576 * 4);
    w.ffh_hi = take(T * 384 * 2); w.ffh_lo = take(T * 384 * 2);
    w.sc = take(T * 64 * 4);
    w.total = o;
    return w;
}

size_t fast_workspace_bytes(const Plan&, const Shape&, int B) { return fast_ws(B).total; }

// bf16 input: fused front kernel (1 launch) instead of im2col + conv0 + conv1 (3 launches)
// The fused conv2 + chan_proj + Gram kernel chains an image band's three MMA stages on one CTA; below ≈100 images the
// two-kernel path (mid kernel, then a Gram kernel with one CTA per token) has the shorter critical path (B = 1: 109 vs 116 us
// per forward, B = 64: 0.111 vs 0.122 ms; equal at 128, fused ahead from 256 up), so small batches keep it.
static bool use_midgram(int B, unsigned flags) {
    if (flags & (PMG_FLAG_NO_FUSED_MIDGRAM | PMG_FLAG_NO_FUSED_MID)) return false;
    return B >= 96 || (flags & PMG_FLAG_ALWAYS_FUSED_MIDGRAM);
}

int fast_launch_count(const Plan& p, const Shape&, int B, unsigned flags) {
    // fused front kernel (+ the hi/lo plane split in front of it for fp32 input), or im2col + conv0 GEMM + conv1 GEMM
    const int front = (flags & PMG_FLAG_NO_FUSED_FRONT) ? 3 : (flags & PMG_FLAG_COUNT_F32_INPUT) ? 2 : 1;
    // conv2 + chan_proj + Gram: one kernel (tc_midgram.cuh), or the fused mid kernel + Gram, or two GEMMs + Gram
    const int mid = use_midgram(B, flags) ? 1 : (flags & PMG_FLAG_NO_FUSED_MID) ? 3 : 2;
    if (!(flags & PMG_FLAG_NO_FUSED_ENCODER) && p.nl <= tc::ENC_MAX_LAYERS) return front + mid + 1 + 1;   // + token_proj, token stage
    return front + mid + 1 + 1 + 6 * p.nl + (p.nl > 1 ? p.nl - 1 : 0) + 3;
}

template <int BN, int NA, int EPI, int ACT>
static int gemm(const void* a_hi, const void* a_lo, int lda, const void* b, int M, int N, int K, int ldc, float* out_f32,
                bf16* out_hi, bf16* out_lo, const float* scale, const float* shift, const float* resid, int splits,
                size_t pstride, cudaStream_t st, int k_cols = 0) {
    tc::GemmParams g{};
    g.M = M; g.N = N; g.K = K; g.epi = EPI; g.act = ACT; g.ldc = ldc; g.out_f32 = out_f32; g.out_hi = out_hi; g.out_lo = out_lo;
    g.scale = scale; g.shift = shift; g.resid = resid; g.partial_stride = pstride;
    const int kc = k_cols > 0 ? k_cols : K;
    return tc::launch_gemm_t<BN, NA, EPI, ACT>(a_hi, a_lo, lda, b, kc, g, splits, kc, st);
}

int copy_logits_to_peers(const float* logits, size_t bytes, const LogitPeers* peers, cudaStream_t st) {
    for (int d = 0; peers && d < peers->n; ++d)
        PMG_CUDA_OK(cudaMemcpyAsync(peers->dst[d], logits, bytes, cudaMemcpyDefault, st), "logits -> peer");
    return PMG_OK;
}

int fast_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x, int x_dtype,
                 int B, float* logits, float* gram, const pmg_taps* taps, char* ws, cudaStream_t st, StageTimer* tm,
                 unsigned flags, const LogitPeers* peers) {
    const FastWs w = fast_ws(B);
    const FastW f = fast_w(p);
    const char* fb = packed + L.fast;
    auto Wb = [&](size_t off) { return reinterpret_cast<const bf16*>(fb + off); };
    auto Pf = [&](size_t off) { return reinterpret_cast<const float*>(packed + off); };
    auto Hb = [&](size_t off) { return reinterpret_cast<bf16*>(ws + off); };
    auto Ff = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const int P = B * 196, T = B * 9;
    int rc;
    PMG_MARK(tm, "start");
    const bool fused_front = !(flags & PMG_FLAG_NO_FUSED_FRONT);
    if (fused_front) {
        // conv0 + conv1 (+ both BN + LeakyReLU) in one TMA-fed tcgen05 kernel (tc_front.cuh)
        CUtensorMap tx, txl, tw1;
        const void* x_hi = x;
        const void* x_lo = x;
        if (x_dtype == PMG_F32) {
            // the planes live in the im2col buffers of the unfused path (401,408 B per image each, 301,056 needed)
            const size_t n8 = (size_t)B * 3 * 224 * 224 / 8;
            const int grid = (int)((n8 + 255) / 256 > (size_t)tc::num_sms() * 16 ? (size_t)tc::num_sms() * 16 : (n8 + 255) / 256);
            PMG_CUDA_OK(tc::launch_pdl(split_image_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float*>(x), Hb(w.a0_hi), Hb(w.a0_lo), n8), "split_image");
            PMG_MARK(tm, "split_image");
            x_hi = Hb(w.a0_hi); x_lo = Hb(w.a0_lo);
        }
        if ((rc = tc::make_tmap_image_rows(&tx, x_hi, B))) return rc;
        if ((rc = tc::make_tmap_image_rows(&txl, x_lo, B))) return rc;
        if ((rc = tc::make_tmap_bf16(&tw1, Wb(f.w1x), 512, 64, 64, 64))) return rc;
        static tc::PerDeviceFlag attr_set;
        if (!attr_set.test()) {
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::front_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FRONT_SMEM), "front smem attr");
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::front_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FRONT_SMEM), "front smem attr");
            attr_set.set();
        }
        tc::FrontParams fp;
        fp.scale0 = Pf(L.conv_scale[0]); fp.shift0 = Pf(L.conv_shift[0]); fp.scale1 = Pf(L.conv_scale[1]); fp.shift1 = Pf(L.conv_shift[1]);
        fp.act1_hi = Hb(w.act1_hi); fp.act1_lo = Hb(w.act1_lo); fp.B = B; fp.trace = g_enc_trace; fp.w0u = reinterpret_cast<const uint8_t*>(Wb(f.w0x));
        const int grid = 2 * B < tc::num_sms() ? 2 * B : tc::num_sms();
        if (x_dtype == PMG_F32)
            PMG_CUDA_OK(tc::launch_pdl(tc::front_kernel<2>, dim3(grid), dim3(tc::FRONT_THREADS), tc::FRONT_SMEM, st, tx, txl, tw1, fp), "front_kernel");
        else
            PMG_CUDA_OK(tc::launch_pdl(tc::front_kernel<1>, dim3(grid), dim3(tc::FRONT_THREADS), tc::FRONT_SMEM, st, tx, txl, tw1, fp), "front_kernel");
        PMG_MARK(tm, "front_conv0_conv1");
    } else {
    // conv0 (:115-117) = im2col + GEMM [B·3136,48]×[48,32], BN+LeakyReLU epilogue → act0 planes == conv1's A
    {
        const size_t chunks = (size_t)B * 3136 * 8;
        const int grid = (int)((chunks + 255) / 256 > 148 * 16 ? 148 * 16 : (chunks + 255) / 256);
        if (x_dtype == PMG_F32)
            PMG_CUDA_OK(tc::launch_pdl(im2col0_kernel<true>, dim3(grid), dim3(256), 0, st, x, Hb(w.a0_hi), Hb(w.a0_lo), B), "im2col0");
        else
            PMG_CUDA_OK(tc::launch_pdl(im2col0_kernel<false>, dim3(grid), dim3(256), 0, st, x, Hb(w.a0_hi), (bf16*)nullptr, B), "im2col0");
        PMG_MARK(tm, "im2col0");
        if (x_dtype == PMG_F32)
            rc = gemm<32, 2, tc::EPI_SPLIT, tc::GACT_LEAKY>(Hb(w.a0_hi), Hb(w.a0_lo), 64, Wb(f.w0), P * 16, 32, 64, 32, nullptr,
                                                            Hb(w.act0_hi), Hb(w.act0_lo), Pf(L.conv_scale[0]),
                                                            Pf(L.conv_shift[0]), nullptr, 1, 0, st);
        else
            rc = gemm<32, 1, tc::EPI_SPLIT, tc::GACT_LEAKY>(Hb(w.a0_hi), nullptr, 64, Wb(f.w0), P * 16, 32, 64, 32, nullptr,
                                                            Hb(w.act0_hi), Hb(w.act0_lo), Pf(L.conv_scale[0]),
                                                            Pf(L.conv_shift[0]), nullptr, 1, 0, st);
        if (rc) return rc;
        PMG_MARK(tm, "conv0");
    }
    // conv1 (:115-117, second strided layer)
    if ((rc = gemm<64, 2, tc::EPI_SPLIT, tc::GACT_LEAKY>(Hb(w.act0_hi), Hb(w.act0_lo), 512, Wb(f.w1), P, 64, 512, 64, nullptr,
                   Hb(w.act1_hi), Hb(w.act1_lo), Pf(L.conv_scale[1]), Pf(L.conv_shift[1]), nullptr, 1, 0, st)))
        return rc;
    PMG_MARK(tm, "conv1");
    }
    const bool fused_midgram = use_midgram(B, flags);
    if (fused_midgram) {
        // conv2 (:124-128) + chan_proj (:136,182) + pad/patches (:157-174) + Gram (:188) in one kernel: neither the
        // 128-channel activation nor the chan_proj map touch HBM (tc_midgram.cuh)
        CUtensorMap ta_hi, ta_lo, tw2, tw3, tw3d;
        if ((rc = tc::make_tmap_bf16(&tw3d, Wb(f.w3), 96, 128, 128, 32))) return rc;       // 32-row box: the copy of channels 32..63
        if ((rc = tc::make_tmap_bf16(&ta_hi, Hb(w.act1_hi), P, 64, 64, tc::MG_BOX_ROWS))) return rc;
        if ((rc = tc::make_tmap_bf16(&ta_lo, Hb(w.act1_lo), P, 64, 64, tc::MG_BOX_ROWS))) return rc;
        if ((rc = tc::make_tmap_bf16(&tw2, Wb(f.w2), 128, 64, 64, 128))) return rc;
        if ((rc = tc::make_tmap_bf16(&tw3, Wb(f.w3), 96, 128, 128, 96))) return rc;
        static tc::PerDeviceFlag attr_set;
        if (!attr_set.test()) {
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::midgram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::MG_SMEM), "midgram smem attr");
            attr_set.set();
        }
        tc::MidGramParams mp;
        mp.scale2 = Pf(L.conv_scale[2]); mp.shift2 = Pf(L.conv_shift[2]); mp.g_hi = Hb(w.g_hi); mp.g_lo = Hb(w.g_lo);
        mp.g_f32 = gram; mp.chan = taps ? taps->chan : nullptr; mp.B = B; mp.trace = g_enc_trace;
        const int grid = 3 * B < tc::num_sms() ? 3 * B : tc::num_sms();
        PMG_CUDA_OK(tc::launch_pdl(tc::midgram_kernel, dim3(grid), dim3(tc::MG_THREADS), tc::MG_SMEM, st, ta_hi, ta_lo, tw2, tw3, tw3d, mp), "midgram_kernel");
        PMG_MARK(tm, "conv2_chan_gram");
    } else {
    if (!(flags & PMG_FLAG_NO_FUSED_MID)) {
        // conv2 (:124-128) + chan_proj (:136,182) in one kernel: the 128-channel activation stays on chip (tc_mid.cuh)
        CUtensorMap ta_hi, ta_lo, tw2, tw3;
        if ((rc = tc::make_tmap_bf16(&ta_hi, Hb(w.act1_hi), P, 64, 64, 128))) return rc;
        if ((rc = tc::make_tmap_bf16(&ta_lo, Hb(w.act1_lo), P, 64, 64, 128))) return rc;
        if ((rc = tc::make_tmap_bf16(&tw2, Wb(f.w2), 128, 64, 64, 128))) return rc;
        if ((rc = tc::make_tmap_bf16(&tw3, Wb(f.w3), 96, 128, 128, 96))) return rc;
        static tc::PerDeviceFlag attr_set;
        if (!attr_set.test()) {
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::MID_SMEM), "mid smem attr");
            attr_set.set();
        }
        tc::MidParams mp;
        mp.scale2 = Pf(L.conv_scale[2]); mp.shift2 = Pf(L.conv_shift[2]); mp.chan = Ff(w.chan); mp.P = P;
        const int tiles = (P + 127) / 128;
        const int grid = tiles < tc::num_sms() ? tiles : tc::num_sms();
        PMG_CUDA_OK(tc::launch_pdl(tc::mid_kernel, dim3(grid), dim3(tc::MID_THREADS), tc::MID_SMEM, st, ta_hi, ta_lo, tw2, tw3, mp), "mid_kernel");
        PMG_MARK(tm, "conv2_chan_proj");
    } else {
    // conv2: final 1×1 (:124-128)
    if ((rc = gemm<128, 2, tc::EPI_SPLIT, tc::GACT_LEAKY>(Hb(w.act1_hi), Hb(w.act1_lo), 64, Wb(f.w2), P, 128, 64, 128, nullptr,
                   Hb(w.act2_hi), Hb(w.act2_lo), Pf(L.conv_scale[2]), Pf(L.conv_shift[2]), nullptr, 1, 0, st)))
        return rc;
    PMG_MARK(tm, "conv2");
    // chan_proj (:136,182) → fp32 pixel-major (B,14,14,96)
    if ((rc = gemm<96, 2, tc::EPI_F32, tc::GACT_NONE>(Hb(w.act2_hi), Hb(w.act2_lo), 128, Wb(f.w3), P, 96, 128, 96, Ff(w.chan),
                   nullptr, nullptr, nullptr, nullptr, nullptr, 1, 0, st)))
        return rc;
    PMG_MARK(tm, "chan_proj");
    }
    if (taps && taps->chan)
        PMG_CUDA_OK(cudaMemcpyAsync(taps->chan, Ff(w.chan), (size_t)P * 96 * 4, cudaMemcpyDeviceToDevice, st), "tap chan");
    // patches + Gram (:185-188)
    {
        static tc::PerDeviceFlag attr_set;
        if (!attr_set.test()) {
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::gram96_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GRAM_SMEM), "gram smem attr");
            attr_set.set();
        }
        const int grid = T < tc::num_sms() ? T : tc::num_sms();
        PMG_CUDA_OK(tc::launch_pdl(tc::gram96_tc_kernel, dim3(grid), dim3(tc::GRAM_THREADS), tc::GRAM_SMEM, st,
                                   (const float*)Ff(w.chan), Hb(w.g_hi), Hb(w.g_lo), gram, T), "gram96_tc");
    }
    PMG_MARK(tm, "gram");
    }
    // token_proj (:192-193): split-K partial sums, reduced (+bias) by tok_ln_kernel
    const size_t pstride = (size_t)T * 192;
    // G' (upper blocks, hi+lo) · W'hᵀ  +  the off-diagonal part of G'hi · W'lᵀ as a second K segment of the SAME launch: one
    // accumulator, one set of split-K partials (it was a second GEMM with its own partial buffers: 4 buffers for the token
    // stage to reduce instead of 2, and a second kernel tail).
    // (the CTA-pair variant tc_gemm2.cuh — tcgen05 cta_group::2 — is parity-tested and measured 64 us here vs 60 us
    // for the cta_group::1 launches: the op is tensor-bound at ~56 % of peak either way, so it is not the default)
    {
        tc::GemmParams g{};
        g.M = T; g.N = 192; g.K = 6144; g.epi = tc::EPI_PARTIAL; g.act = tc::GACT_NONE; g.ldc = 192; g.out_f32 = Ff(w.partial);
        g.partial_stride = pstride;
        if ((rc = tc::launch_gemm_t<192, 2, tc::EPI_PARTIAL, tc::GACT_NONE, true>(Hb(w.g_hi), Hb(w.g_lo), 6144, Wb(f.wth), 6144, g,
                       token_splits(T), 6144, st, Hb(w.g_hi) + 3072, Wb(f.wtl), 3072)))
            return rc;
    }
    PMG_MARK(tm, "token_proj");
    if (!(flags & PMG_FLAG_NO_FUSED_ENCODER) && p.nl <= tc::ENC_MAX_LAYERS) {
        // split-K reduce + bias, the whole transformer, task pooling and the heads: one kernel (tc_encoder.cuh)
        static tc::PerDeviceFlag attr_set;
        if (!attr_set.test()) {
            PMG_CUDA_OK(cudaFuncSetAttribute(tc::encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::ENC_SMEM), "encoder smem attr");
            attr_set.set();
        }
        tc::EncParams e{};
        e.wstream = reinterpret_cast<const uint8_t*>(fb + f.wenc);
        e.partial = Ff(w.partial); e.pstride = pstride; e.nsplit = w.splits; e.tok_bias = Pf(L.tok_b);
        e.cls_bias = Pf(L.cls_b); e.logits = logits;
        e.tok_in = taps ? taps->tokens_in : nullptr; e.tok_out = taps ? taps->tokens_out : nullptr;
        e.nl = p.nl; e.B = B; e.ntask = p.ntask; e.ncls_total = p.ncls_total; e.trace = g_enc_trace;
        e.n_peers = peers ? peers->n : 0;
        for (int d = 0; d < e.n_peers; ++d) e.peer_logits[d] = peers->dst[d];
        for (int l = 0; l < p.nl; ++l) {
            const TfLayerOff& t = L.tf[l];
            e.layer[l] = {Pf(t.b_in), Pf(t.b_o), Pf(t.b_1), Pf(t.b_2), Pf(t.ln1w), Pf(t.ln1b), Pf(t.ln2w), Pf(t.ln2b)};
        }
        for (int t = 0; t < p.ntask; ++t)
            for (int c = p.cls_off[t]; c < p.cls_off[t + 1]; ++c) e.task_of[c] = (unsigned char)t;
        const int tiles = (B + tc::ENC_IPT - 1) / tc::ENC_IPT;
        const int grid = tiles < tc::num_sms() ? tiles : tc::num_sms();
        PMG_CUDA_OK(tc::launch_pdl(tc::encoder_kernel, dim3(grid), dim3(tc::ENC_THREADS), tc::ENC_SMEM, st, e), "encoder_kernel");
        PMG_MARK(tm, "encoder_heads");
        return PMG_OK;
    }
    {
        const TfLayerOff* t0 = p.nl > 0 ? &L.tf[0] : nullptr;
        PMG_CUDA_OK(tc::launch_pdl(tok_ln_kernel, dim3((T + 3) / 4), dim3(128), 0, st, (const float*)Ff(w.partial), w.splits, pstride,
                                   Pf(L.tok_b), Ff(w.tok), t0 ? Pf(t0->ln1w) : (const float*)nullptr,
                                   t0 ? Pf(t0->ln1b) : (const float*)nullptr, Hb(w.h_hi), Hb(w.h_lo), T, t0 ? 1 : 0), "tok_ln");
        PMG_MARK(tm, "tok_reduce_ln");
    }
    if (taps && taps->tokens_in)
        PMG_CUDA_OK(cudaMemcpyAsync(taps->tokens_in, Ff(w.tok), (size_t)T * 192 * 4, cudaMemcpyDeviceToDevice, st), "tap tok");
    // transformer (:194)
    for (int l = 0; l < p.nl; ++l) {
        const TfLayerOff& t = L.tf[l];
        if (l > 0) {
            PMG_CUDA_OK(tc::launch_pdl(tok_ln_kernel, dim3((T + 3) / 4), dim3(128), 0, st, (const float*)nullptr, 0, (size_t)0,
                                       (const float*)nullptr, Ff(w.tok), Pf(t.ln1w), Pf(t.ln1b), Hb(w.h_hi), Hb(w.h_lo), T, 1), "ln1");
        }
        if ((rc = gemm<192, 2, tc::EPI_F32, tc::GACT_NONE>(Hb(w.h_hi), Hb(w.h_lo), 192, Wb(f.win[l]), T, 576, 192, 576, Ff(w.qkv),
                       nullptr, nullptr, nullptr, Pf(t.b_in), nullptr, 1, 0, st)))
            return rc;
        PMG_CUDA_OK(tc::launch_pdl(attn9_kernel, dim3(B), dim3(128), 0, st, (const float*)Ff(w.qkv), Hb(w.h_hi), Hb(w.h_lo)), "attn9");
        if ((rc = gemm<192, 2, tc::EPI_F32, tc::GACT_NONE>(Hb(w.h_hi), Hb(w.h_lo), 192, Wb(f.wo[l]), T, 192, 192, 192, Ff(w.tok),
                       nullptr, nullptr, nullptr, Pf(t.b_o), Ff(w.tok), 1, 0, st)))
            return rc;
        PMG_CUDA_OK(tc::launch_pdl(tok_ln_kernel, dim3((T + 3) / 4), dim3(128), 0, st, (const float*)nullptr, 0, (size_t)0,
                                   (const float*)nullptr, Ff(w.tok), Pf(t.ln2w), Pf(t.ln2b), Hb(w.h_hi), Hb(w.h_lo), T, 1), "ln2");
        if ((rc = gemm<192, 2, tc::EPI_SPLIT, tc::GACT_RELU>(Hb(w.h_hi), Hb(w.h_lo), 192, Wb(f.wf1[l]), T, 384, 192, 384, nullptr,
                       Hb(w.ffh_hi), Hb(w.ffh_lo), nullptr, Pf(t.b_1), nullptr, 1, 0, st)))
            return rc;
        if ((rc = gemm<192, 2, tc::EPI_F32, tc::GACT_NONE>(Hb(w.ffh_hi), Hb(w.ffh_lo), 384, Wb(f.wf2[l]), T, 192, 384, 192,
                       Ff(w.tok), nullptr, nullptr, nullptr, Pf(t.b_2), Ff(w.tok), 1, 0, st)))
            return rc;
        PMG_MARK(tm, "encoder_layer");
    }
    if (taps && taps->tokens_out)
        PMG_CUDA_OK(cudaMemcpyAsync(taps->tokens_out, Ff(w.tok), (size_t)T * 192 * 4, cudaMemcpyDeviceToDevice, st), "tap tok out");
    // task attention pooling + heads (:197-198): one [T,192]×[192,64] tensor-core GEMM + a finishing kernel
    PMG_CUDA_OK(tc::launch_pdl(tok_ln_kernel, dim3((T + 3) / 4), dim3(128), 0, st, (const float*)nullptr, 0, (size_t)0,
                               (const float*)nullptr, Ff(w.tok), (const float*)nullptr, (const float*)nullptr, Hb(w.h_hi),
                               Hb(w.h_lo), T, 2), "tok split");
    if ((rc = gemm<64, 2, tc::EPI_F32, tc::GACT_NONE>(Hb(w.h_hi), Hb(w.h_lo), 192, Wb(f.wq), T, 64, 192, 64, Ff(w.sc), nullptr,
                                                      nullptr, nullptr, nullptr, nullptr, 1, 0, st)))
        return rc;
    {
        FinishArgs a;
        a.sc = Ff(w.sc); a.bias = Pf(L.cls_b); a.logits = logits; a.B = B; a.ntask = p.ntask; a.ncls_total = p.ncls_total;
        for (int t = 0; t < p.ntask; ++t)
            for (int c = p.cls_off[t]; c < p.cls_off[t + 1]; ++c) a.task_of[c] = (unsigned char)t;
        PMG_CUDA_OK(tc::launch_pdl(pool_finish_kernel, dim3((B * p.ncls_total + 255) / 256), dim3(256), 0, st, a), "pool_finish");
    }
    PMG_MARK(tm, "pool_heads");
    return copy_logits_to_peers(logits, (size_t)B * p.ncls_total * sizeof(float), peers, st);      // unfused token stage: copies after the fact
}

// ------------------------------------------------------------------------------------------------
// test / microbenchmark hooks
// ------------------------------------------------------------------------------------------------
static int make_tmap_gram3d(CUtensorMap* m, const void* base, int B, int C, int HW, uint32_t box_rows) {
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return tc::encode_tmap_cached(m, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B, "cuTensorMapEncodeTiled(gram 3d)");
}

template <int BN, bool LSU>
static int launch_gram_tc(const void* F, int B, int C, int HW, float scale, float* G, cudaStream_t st) {
    CUtensorMap ta{}, tb{};
    int rc;
    if (!LSU) {
        if ((rc = make_tmap_gram3d(&ta, F, B, C, HW, 128))) return rc;
        if ((rc = make_tmap_gram3d(&tb, F, B, C, HW, BN))) return rc;
    }
    constexpr int smem = tc::gb_smem_bytes<BN>();
    static tc::PerDeviceFlag attr_set;
    if (!attr_set.test()) {
        PMG_CUDA_OK(cudaFuncSetAttribute(tc::gram_batched_tc_kernel<BN, LSU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "gram tc smem attr");
        attr_set.set();
    }
    tc::GramBParams p{G, B, C, HW, scale, reinterpret_cast<const bf16*>(F)};
    const long long items = (long long)B * ((C + 127) / 128) * ((C + BN - 1) / BN);
    const int grid = (int)(items < tc::num_sms() ? items : tc::num_sms());
    tc::gram_batched_tc_kernel<BN, LSU><<<grid, LSU ? tc::GB_THREADS_LSU : tc::GB_THREADS, smem, st>>>(ta, tb, p);
    PMG_LAUNCH_OK("gram_batched_tc");
    return PMG_OK;
}

// tcgen05 batched Gram (tc_gramb.cuh), bf16 input: rows TMA can address (HW % 8 == 0: 16-byte pitch, 16-byte aligned base)
// go through the TMA producer; every other bf16 shape (HW = 25, 196, …) through the loader-warp variant of the same kernel.
// Long rows that TMA cannot take stay on the CUDA-core kernel: a warp-load per 128-byte row piece is the small-map design.
int gram_batched_fast(const void* F, int f_dtype, int B, int C, int HW, float scale, float* G, cudaStream_t st) {
    if (f_dtype != PMG_BF16 || C % 4 != 0 || C < 16) return PMG_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(G) & 15) return PMG_ERR_UNSUPPORTED;
    const bool tma_ok = HW % 8 == 0 && !(reinterpret_cast<uintptr_t>(F) & 15);
    if (!tma_ok && HW > 1024) return PMG_ERR_UNSUPPORTED;
#define PMG_GRAM_TC(BNV) return tma_ok ? launch_gram_tc<BNV, false>(F, B, C, HW, scale, G, st) : launch_gram_tc<BNV, true>(F, B, C, HW, scale, G, st)
    if (C <= 32) PMG_GRAM_TC(32);
    if (C <= 64) PMG_GRAM_TC(64);
    if (C <= 96) PMG_GRAM_TC(96);
    PMG_GRAM_TC(128);
#undef PMG_GRAM_TC
}

}  // namespace pmg

extern "C" void pmg_debug_set_trace(void* dev_buf_512_i64) { pmg::g_enc_trace = reinterpret_cast<long long*>(dev_buf_512_i64); }

extern "C" int pmg_tc_gemm2(const void* a_hi, const void* a_lo, const void* b, int M, int K, int splits, float* out, pmg_stream_t stream) {
    if (!a_hi || !b || !out) return PMG_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return a_lo ? tc::launch_gemm2_t<2>(a_hi, a_lo, K, b, K, M, K, splits, out, (size_t)M * 192, st)
                : tc::launch_gemm2_t<1>(a_hi, a_lo, K, b, K, M, K, splits, out, (size_t)M * 192, st);
}

extern "C" int pmg_tc_gemm(const void* a_hi, const void* a_lo, const void* b, int M, int N, int K, int bn, int splits,
                           float* out, pmg_stream_t stream) {
    if (!a_hi || !b || !out) return PMG_ERR_BAD_ARG;
    tc::GemmParams g{};
    g.M = M; g.N = N; g.K = K; g.ldc = N; g.out_f32 = out;
    g.epi = splits > 1 ? tc::EPI_PARTIAL : tc::EPI_F32;
    g.partial_stride = (size_t)M * N;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define PMG_TG(BNV)                                                                                                        \
    if (bn == BNV) {                                                                                                       \
        if (splits > 1)                                                                                                    \
            return a_lo ? tc::launch_gemm_t<BNV, 2, tc::EPI_PARTIAL, tc::GACT_NONE>(a_hi, a_lo, K, b, K, g, splits, K, st) \
                        : tc::launch_gemm_t<BNV, 1, tc::EPI_PARTIAL, tc::GACT_NONE>(a_hi, a_lo, K, b, K, g, splits, K, st); \
        return a_lo ? tc::launch_gemm_t<BNV, 2, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, K, b, K, g, splits, K, st)         \
                    : tc::launch_gemm_t<BNV, 1, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, K, b, K, g, splits, K, st);        \
    }
    PMG_TG(32) PMG_TG(64) PMG_TG(96) PMG_TG(128) PMG_TG(192)
#undef PMG_TG
    return PMG_ERR_UNSUPPORTED;
}
