// pm_path.cu — PM model (MultiTaskPatchGAN) eval forward on B200.
// Reference: /root/reference/Models/models_PM.py:91-194 (trunk :114-141, heads :27-76, forward :159-194).
//
// Every trunk layer is  Conv2d(k4, s2|s1, p1, bias) → InstanceNorm2d(affine, per-sample stats) → LeakyReLU(0.2):
//   im2col (hi/lo bf16 planes, zero padding folded in)  →  tc::gemm_kernel (tcgen05, fp32 accumulate, + bias)
//   →  instnorm kernel (fp32 two-pass statistics per (sample, channel), affine, LeakyReLU, hi/lo split)
// The 12 task heads (attn_conv 1×1 → spatial softmax(a/τ) → cls_conv 1×1 → attention-weighted sum) are one
// tensor-core GEMM  feats[B·HW, C] × [attn rows | cls rows]ᵀ  (N = 64) plus a warp-shuffle finishing kernel
// (softmax over the HW positions, weighted sum of the per-position class projections): cls_conv is linear, so
// Σ_p A_p (W·f_p) == W·(Σ_p A_p f_p) and no (B, K, H, W) map is materialised.
#include "pmg_common.cuh"
#include "tc_host.cuh"
#include "pm_conv0.cuh"

#include <mutex>

namespace pmg {

using bf16 = __nv_bfloat16;
extern long long* g_enc_trace;            // pmg_fast.cu (pmg_debug_set_trace)
constexpr int PM_MAX_LAYERS = 8;
// Images per internal pass (bounds the workspace) and the number of passes in flight.  A pass is a chain of tensor-bound
// kernels (implicit-GEMM convolutions: persistent, one CTA per SM) and HBM-bound ones (InstanceNorm, layer-0 im2col: small
// blocks that fit beside a GEMM CTA).  Two passes on two streams let the HBM-bound kernels of one pass run under the
// convolutions of the other instead of leaving the tensor pipe idle a third of the time.
static int g_pm_chunk = 128, g_pm_lanes = 2;
constexpr float kInEps = 1e-5f;        // nn.InstanceNorm2d default

struct PmLayer { int cin, cout, stride, K, Kp; };   // K = cin*16, Kp = K rounded up to 64

struct PmPlan {
    int n_layers = 0;
    PmLayer layer[PM_MAX_LAYERS];
    int C = 0, ntask = 0, ncls_total = 0;
    int cls_off[PMG_MAX_TASKS + 1];
    bool ok = false;
};

static PmPlan pm_plan(const pm_config* c) {
    PmPlan p;
    if (!c || c->input_nc < 1 || c->ndf < 32 || c->ndf % 32 || c->n_tasks < 1 || c->n_tasks > PMG_MAX_TASKS || !(c->attn_tau > 0.f))
        return p;
    int cin = c->input_nc, nf = c->ndf;
    float rf = c->patch_size;
    auto add = [&](int ci, int co, int st) {
        PmLayer& l = p.layer[p.n_layers++];
        l.cin = ci; l.cout = co; l.stride = st; l.K = ci * 16; l.Kp = (l.K + 63) / 64 * 64;
    };
    while (rf > 4.f && nf <= 512) {                       // models_PM.py:121-132
        if (p.n_layers >= PM_MAX_LAYERS - 1) return p;
        add(cin, nf, 2);
        cin = nf; nf *= 2; rf /= 2.f;
    }
    add(cin, nf, 1);                                      // :134-139
    p.C = nf;
    p.ntask = c->n_tasks;
    int off = 0;
    for (int t = 0; t < c->n_tasks; ++t) {
        if (c->ncls[t] < 1 || c->ncls[t] > 8) return p;     // finishing kernel keeps ≤ 8 class accumulators per task
        p.cls_off[t] = off; off += c->ncls[t];
    }
    p.cls_off[c->n_tasks] = off;
    p.ncls_total = off;
    if (p.ntask + p.ncls_total > 64) return p;            // heads GEMM is one N = 64 tile
    p.ok = true;
    return p;
}

struct PmShape { int h[PM_MAX_LAYERS + 1], w[PM_MAX_LAYERS + 1]; bool ok = false; };

static PmShape pm_shape(const PmPlan& p, int H, int W) {
    PmShape s;
    s.h[0] = H; s.w[0] = W;
    for (int l = 0; l < p.n_layers; ++l) {
        s.h[l + 1] = (s.h[l] + 2 - 4) / p.layer[l].stride + 1;
        s.w[l + 1] = (s.w[l] + 2 - 4) / p.layer[l].stride + 1;
        if (s.h[l] + 2 < 4 || s.w[l] + 2 < 4 || s.h[l + 1] < 1 || s.w[l + 1] < 1) return s;
    }
    s.ok = true;
    return s;
}

struct PmPacked {
    size_t w[PM_MAX_LAYERS], bias[PM_MAX_LAYERS], gamma[PM_MAX_LAYERS], beta[PM_MAX_LAYERS];
    size_t wh, bh;      // heads: [64][C] bf16, bias [64] fp32 (attn biases in the first ntask slots)
    size_t total;
};

static PmPacked pm_packed(const PmPlan& p) {
    PmPacked L{};
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    for (int l = 0; l < p.n_layers; ++l) {
        L.w[l] = take((size_t)p.layer[l].cout * p.layer[l].Kp * 2);
        L.bias[l] = take(p.layer[l].cout * 4); L.gamma[l] = take(p.layer[l].cout * 4); L.beta[l] = take(p.layer[l].cout * 4);
    }
    L.wh = take((size_t)64 * p.C * 2);
    L.bh = take(64 * 4);
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------------ packing
// layer 0: [co][k = c*16+ky*4+kx] (torch order), zero-padded to Kp;  layers ≥ 1: [co][k = (ky*4+kx)*cin + c]
__global__ void pm_pack_conv_kernel(const float* __restrict__ in, bf16* __restrict__ out, int cout, int cin, int Kp, int nhwc) {
    const size_t total = (size_t)cout * Kp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(i / Kp), k = (int)(i % Kp);
        float v = 0.f;
        if (k < cin * 16) {
            if (nhwc) { const int kk = k / cin, c = k % cin; v = in[((size_t)co * cin + c) * 16 + kk]; }
            else v = in[(size_t)co * cin * 16 + k];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}
__global__ void pm_copy_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void pm_cvt_kernel(const float* __restrict__ in, bf16* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

// ------------------------------------------------------------------------------------------------ im2col
// first layer: x (B,Cin,H,W) NCHW fp32|bf16 → A planes [B·Ho·Wo][Kp], k = c*16 + ky*4 + kx, pad 1
template <bool F32IN>
__global__ void __launch_bounds__(256) pm_im2col_first_kernel(const void* __restrict__ x, bf16* __restrict__ a_hi,
                                                              bf16* __restrict__ a_lo, int B, int Cin, int H, int W, int Ho,
                                                              int Wo, int stride, int Kp) {
    const int chunks = Kp / 8;
    const size_t total = (size_t)B * Ho * Wo * chunks;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t r = t / chunks;
        const int j = (int)(t % chunks);
        const int xo = (int)(r % Wo), yo = (int)((r / Wo) % Ho);
        const size_t b = r / ((size_t)Wo * Ho);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = j * 8 + e, c = k >> 4, ky = (k >> 2) & 3, kx = k & 3;
            const int y = yo * stride - 1 + ky, xx = xo * stride - 1 + kx;
            float val = 0.f;
            if (c < Cin && y >= 0 && y < H && xx >= 0 && xx < W) {
                const size_t idx = ((b * Cin + c) * H + y) * W + xx;
                val = F32IN ? reinterpret_cast<const float*>(x)[idx] : __bfloat162float(reinterpret_cast<const bf16*>(x)[idx]);
            }
            v[e] = val;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) tc::split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        *reinterpret_cast<uint4*>(a_hi + r * Kp + j * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (F32IN) *reinterpret_cast<uint4*>(a_lo + r * Kp + j * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);   // bf16 input: lo ≡ 0, no plane
    }
}

// Same operand planes, one block per output row (b, yo): the 4·Cin image rows that row needs are read once, coalesced, into
// shared memory (zero padding included) and the [Wo][Kp] operand rows are written from there as 16-byte chunks.  The
// per-element gather above re-reads every pixel 4 times with 2-byte loads and ran at ≈1.3 TB/s of operand writes.
template <bool F32IN>
__global__ void __launch_bounds__(256) pm_im2col_first_rows_kernel(const void* __restrict__ x, bf16* __restrict__ a_hi,
                                                                   bf16* __restrict__ a_lo, int Cin, int H, int W, int Ho,
                                                                   int Wo, int stride, int Kp) {
    extern __shared__ float s_rows[];                       // [Cin·4][W + 2]: column 0 is x = -1
    const int b = blockIdx.x / Ho, yo = blockIdx.x % Ho;
    const int ld = W + 2, nrows = Cin * 4;
    for (int i = threadIdx.x; i < nrows * ld; i += 256) {
        const int rr = i / ld, col = i - rr * ld, c = rr >> 2, ky = rr & 3;
        const int y = yo * stride - 1 + ky, xx = col - 1;
        float v = 0.f;
        if (y >= 0 && y < H && xx >= 0 && xx < W) {
            const size_t idx = (((size_t)b * Cin + c) * H + y) * W + xx;
            v = F32IN ? reinterpret_cast<const float*>(x)[idx] : __bfloat162float(reinterpret_cast<const bf16*>(x)[idx]);
        }
        s_rows[i] = v;
    }
    __syncthreads();
    const int chunks = Kp / 8;
    const size_t row0 = ((size_t)b * Ho + yo) * Wo;
    for (int t = threadIdx.x; t < Wo * chunks; t += 256) {
        const int xo = t / chunks, j = t - xo * chunks;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = j * 8 + e, rr = k >> 2, kx = k & 3;         // rr = c·4 + ky
            v[e] = rr < nrows ? s_rows[rr * ld + xo * stride + kx] : 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) tc::split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        const size_t o = (row0 + xo) * Kp + j * 8;
        *reinterpret_cast<uint4*>(a_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (F32IN) *reinterpret_cast<uint4*>(a_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// deeper layers: activation planes (B,H,W,C) bf16 → A planes [B·Ho·Wo][16·C], k = (ky*4+kx)*C + c; 16-byte copies
__global__ void __launch_bounds__(256) pm_im2col_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo,
                                                        bf16* __restrict__ a_hi, bf16* __restrict__ a_lo, int B, int C, int H,
                                                        int W, int Ho, int Wo, int stride) {
    const int c8 = C / 8;
    const size_t total = (size_t)B * Ho * Wo * 16 * c8;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(t % c8);
        const int kk = (int)((t / c8) & 15);
        const size_t r = t / ((size_t)c8 * 16);
        const int xo = (int)(r % Wo), yo = (int)((r / Wo) % Ho);
        const size_t b = r / ((size_t)Wo * Ho);
        const int y = yo * stride - 1 + (kk >> 2), xx = xo * stride - 1 + (kk & 3);
        uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
        if (y >= 0 && y < H && xx >= 0 && xx < W) {
            const size_t src = ((b * H + y) * W + xx) * C + cc * 8;
            vh = *reinterpret_cast<const uint4*>(in_hi + src);
            vl = *reinterpret_cast<const uint4*>(in_lo + src);
        }
        const size_t dst = r * (size_t)(16 * C) + (size_t)kk * C + cc * 8;
        *reinterpret_cast<uint4*>(a_hi + dst) = vh;
        *reinterpret_cast<uint4*>(a_lo + dst) = vl;
    }
}

// ------------------------------------------------------------------------------------------------ InstanceNorm
// normalise + LeakyReLU + hi/lo split of one (sample, 32-channel) slice: a thread owns 4 channels of every 32nd pixel
__device__ __forceinline__ void pm_normalise_pass(const float* __restrict__ src, const float4 sc, const float4 sh, bf16* __restrict__ o_hi,
                                                  bf16* __restrict__ o_lo, float* __restrict__ o_f32, size_t obase, int pr, int HW, int C) {
    auto emit = [&](int pp, const float4& v) {
        float y0 = fmaf(v.x, sc.x, sh.x), y1 = fmaf(v.y, sc.y, sh.y), y2 = fmaf(v.z, sc.z, sh.z), y3 = fmaf(v.w, sc.w, sh.w);
        y0 = fmaxf(y0, 0.2f * y0); y1 = fmaxf(y1, 0.2f * y1); y2 = fmaxf(y2, 0.2f * y2); y3 = fmaxf(y3, 0.2f * y3);
        uint32_t h0, l0, h1, l1;
        tc::split_bf16x2(y0, y1, h0, l0);
        tc::split_bf16x2(y2, y3, h1, l1);
        const size_t o = obase + (size_t)pp * C;
        *reinterpret_cast<uint2*>(o_hi + o) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(o_lo + o) = make_uint2(l0, l1);
        if (o_f32) *reinterpret_cast<float4*>(o_f32 + o) = make_float4(y0, y1, y2, y3);
    };
    int p = pr;
    for (; p + 7 * 32 < HW; p += 8 * 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(src + (size_t)(p + 32 * u) * C);
#pragma unroll
        for (int u = 0; u < 8; ++u) emit(p + 32 * u, v[u]);
    }
    for (; p < HW; p += 32) emit(p, *reinterpret_cast<const float4*>(src + (size_t)p * C));
}

// conv (B, HW, C) fp32 → LeakyReLU(IN(conv)) as hi/lo planes (and fp32 on request).  Block = (sample, 32 channels).
// A thread owns 4 channels (float4) of every 32nd pixel: 8 lanes cover the block's 128-byte channel slice of one pixel,
// a warp load covers 4 pixels, and 8 independent loads per thread are in flight — the first version (one 4-byte load per
// lane per iteration) kept ≈1 KB per block in flight and ran at a fraction of the HBM rate.  Statistics in ONE pass over
// values shifted by the sample's first pixel (Σd, Σd² with d = x − x₀: no catastrophic cancellation when |mean| ≫ std),
// then one normalise pass (the slice is re-read, mostly from L2).
__global__ void __launch_bounds__(256) pm_instnorm_kernel(const float* __restrict__ conv, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, bf16* __restrict__ o_hi,
                                                          bf16* __restrict__ o_lo, float* __restrict__ o_f32, int HW, int C) {
    __shared__ float4 red_s[32][8], red_q[32][8];
    __shared__ float4 s_scale[8], s_shift[8];
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int cl = (threadIdx.x & 7) * 4;                    // this thread's 4 channels inside the block's 32
    const int pr = threadIdx.x >> 3;                          // pixel phase 0..31
    const float* src = conv + (size_t)b * HW * C + c0 + cl;
    const float4 x0 = *reinterpret_cast<const float4*>(src);  // pixel 0 of the sample: the shift
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    int p = pr;
    for (; p + 7 * 32 < HW; p += 8 * 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(src + (size_t)(p + 32 * u) * C);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float dx = v[u].x - x0.x, dy = v[u].y - x0.y, dz = v[u].z - x0.z, dw = v[u].w - x0.w;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
        }
    }
    for (; p < HW; p += 32) {
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)p * C);
        const float dx = v.x - x0.x, dy = v.y - x0.y, dz = v.z - x0.z, dw = v.w - x0.w;
        s.x += dx; s.y += dy; s.z += dz; s.w += dw;
        q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
    }
    red_s[pr][threadIdx.x & 7] = s;
    red_q[pr][threadIdx.x & 7] = q;
    __syncthreads();
    if (threadIdx.x < 8) {
        float4 ts = make_float4(0.f, 0.f, 0.f, 0.f), tq = ts;
        for (int i = 0; i < 32; ++i) {
            const float4 a = red_s[i][threadIdx.x], c = red_q[i][threadIdx.x];
            ts.x += a.x; ts.y += a.y; ts.z += a.z; ts.w += a.w;
            tq.x += c.x; tq.y += c.y; tq.z += c.z; tq.w += c.w;
        }
        const float inv = 1.f / (float)HW;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c0 + cl), be = *reinterpret_cast<const float4*>(beta + c0 + cl);
        float4 sc, sh;
        auto fin = [&](float sd, float sq, float xs, float gg, float bb, float& o_sc, float& o_sh) {
            const float md = sd * inv;                                   // mean of the shifted values
            const float var = fmaxf(sq * inv - md * md, 0.f);            // biased variance (InstanceNorm2d)
            o_sc = gg / sqrtf(var + kInEps);
            o_sh = bb - (xs + md) * o_sc;
        };
        fin(ts.x, tq.x, x0.x, g.x, be.x, sc.x, sh.x); fin(ts.y, tq.y, x0.y, g.y, be.y, sc.y, sh.y);
        fin(ts.z, tq.z, x0.z, g.z, be.z, sc.z, sh.z); fin(ts.w, tq.w, x0.w, g.w, be.w, sc.w, sh.w);
        s_scale[threadIdx.x] = sc; s_shift[threadIdx.x] = sh;
    }
    __syncthreads();
    pm_normalise_pass(src, s_scale[threadIdx.x & 7], s_shift[threadIdx.x & 7], o_hi, o_lo, o_f32, (size_t)b * HW * C + c0 + cl, pr, HW, C);
}

// Block of 256 threads = (sample b, channels c0..c0+31): merge the sample's per-(tile, epilogue warp) partials (mean, M2) in a
// fixed order — warp w takes partials w, w+8, …, then thread c < 32 folds the 8 results (Chan et al.: deterministic, no
// cancellation) — and turn them into InstanceNorm's scale / shift (y = v·scale + shift for the values v the statistics were
// taken of).  Returns true in the 32 threads that hold a result.
template <bool RAW_SUMS = false>
__device__ __forceinline__ bool pm_merge_tile_stats(const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, int b, int c0, int C, int HW, int tiles,
                                                    int rows_full, int rows_last, float& sc, float& sh) {
    __shared__ float m_n[8][32], m_mean[8][32], m_m2[8][32];
    {
        const int c = threadIdx.x & 31, pw = threadIdx.x >> 5;            // channel, partial phase (one warp per phase)
        const int n_part = tiles * 4;
        const float2* src = stats + ((size_t)b * n_part) * C + c0 + c;
        float n = 0.f, mean = 0.f, m2 = 0.f;
        for (int i0 = pw; i0 < n_part; i0 += 8 * 8) {                     // 8 loads in flight, then the (dependent) merges
            float2 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 8 * u;
                t[u] = i < n_part ? src[(size_t)i * C] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 8 * u;
                const int ty = i >> 2, q = i & 3;
                const int rows = ty < tiles - 1 ? rows_full : rows_last;
                const int cnt = i < n_part ? min(32, max(0, rows - 32 * q)) : 0;      // warp-uniform
                if (cnt == 0) continue;
                const float nb = (float)cnt, nn = n + nb;
                // RAW_SUMS: the partial is (Σ v, Σ v²) of values already shifted near zero (pm_conv0.cuh) → (mean, M2)
                const float pm = RAW_SUMS ? t[u].x / nb : t[u].x;
                const float pm2 = RAW_SUMS ? fmaxf(t[u].y - t[u].x * pm, 0.f) : t[u].y;
                const float d = pm - mean;
                mean += d * (nb / nn);
                m2 += pm2 + d * d * (n * nb / nn);
                n = nn;
            }
        }
        m_n[pw][c] = n; m_mean[pw][c] = mean; m_m2[pw][c] = m2;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return false;
    const int c = threadIdx.x;
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int w = 0; w < 8; ++w) {
        const float nb = m_n[w][c];
        if (nb == 0.f) continue;
        const float nn = n + nb, d = m_mean[w][c] - mean;
        mean += d * (nb / nn);
        m2 += m_m2[w][c] + d * d * (n * nb / nn);
        n = nn;
    }
    const float var = fmaxf(m2 / (float)HW, 0.f);                         // biased variance (InstanceNorm2d)
    sc = gamma[c0 + c] / sqrtf(var + kInEps);
    sh = beta[c0 + c] - mean * sc;
    return true;
}

// layer 0 (pm_conv0.cuh): partials of pass 1 → (scale, shift) [B][C] for pass 2; grid (C/32, B)
__global__ void __launch_bounds__(256) pm_conv0_finalize_kernel(const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float2* __restrict__ scsh, int HW,
                                                                int C, int tiles, int rows) {
    float sc, sh;
    if (pm_merge_tile_stats<true>(stats, gamma, beta, blockIdx.y, blockIdx.x * 32, C, HW, tiles, rows, rows, sc, sh))
        scsh[(size_t)blockIdx.y * C + blockIdx.x * 32 + threadIdx.x] = make_float2(sc, sh);
}

// Same layer with the statistics taken from the convolution's own epilogue (tc_gemm.cuh / tc_igemm.cuh `stats`): every
// epilogue warp left (mean, Σ(x − mean)²) of its ≤ 32 rows per channel, the tile lying inside one sample.  The block merges
// the sample's partials in a fixed order (Chan et al. pairwise update: deterministic, no cancellation) and then makes the
// ONE pass over the fp32 output that normalises it — the separate statistics pass (a full extra read of the output, from
// HBM for the early layers) is gone.  tiles = M tiles per sample, rows_full / rows_last = rows of a full / the last tile.
__global__ void __launch_bounds__(256) pm_instnorm_fused_kernel(const float* __restrict__ conv, const float2* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                bf16* __restrict__ o_hi, bf16* __restrict__ o_lo,
                                                                float* __restrict__ o_f32, int HW, int C, int tiles, int rows_full,
                                                                int rows_last) {
    __shared__ float4 s_scale[8], s_shift[8];
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    float sc, sh;
    if (pm_merge_tile_stats(stats, gamma, beta, b, c0, C, HW, tiles, rows_full, rows_last, sc, sh)) {
        reinterpret_cast<float*>(s_scale)[threadIdx.x] = sc;
        reinterpret_cast<float*>(s_shift)[threadIdx.x] = sh;
    }
    __syncthreads();
    const int cl = (threadIdx.x & 7) * 4, pr = threadIdx.x >> 3;
    pm_normalise_pass(conv + (size_t)b * HW * C + c0 + cl, s_scale[threadIdx.x & 7], s_shift[threadIdx.x & 7], o_hi, o_lo, o_f32,
                      (size_t)b * HW * C + c0 + cl, pr, HW, C);
}

// ------------------------------------------------------------------------------------------------ heads finish
// sc (B·HW, 64): col t < ntask = attention logit a_t (bias added), col ntask + c = cls_conv projection of class c.
// One warp per (sample, task): A = softmax_p(a/τ) (or uniform), logits_c = Σ_p A_p·proj_c(p) / (ΣA + 1e-6).
struct PmFinishArgs {
    const float* sc; float* logits; float* attn;
    int B, HW, ntask, ncls_total, ablate;
    float inv_tau;
    int cls_off[PMG_MAX_TASKS + 1];
};

__global__ void __launch_bounds__(128) pm_heads_finish_kernel(const PmFinishArgs a) {
    const int wid = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (wid >= a.B * a.ntask) return;
    const int b = wid / a.ntask, t = wid % a.ntask;
    const float* s = a.sc + (size_t)b * a.HW * 64;
    float mx = -INFINITY;
    if (!a.ablate)
        for (int p = lane; p < a.HW; p += 32) mx = fmaxf(mx, s[(size_t)p * 64 + t] * a.inv_tau);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int c0 = a.cls_off[t], nc = a.cls_off[t + 1] - c0;
    float den = 0.f, acc[8] = {};
    for (int p = lane; p < a.HW; p += 32) {
        const float e = a.ablate ? 1.f : expf(s[(size_t)p * 64 + t] * a.inv_tau - mx);
        den += e;
        for (int k = 0; k < nc && k < 8; ++k) acc[k] = fmaf(e, s[(size_t)p * 64 + a.ntask + c0 + k], acc[k]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        den += __shfl_xor_sync(0xffffffffu, den, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    // A = e/den sums to 1; the reference divides Σ(M·A) by (ΣA + 1e-6) for the softmax variant (models_PM.py:73-75)
    // and uses plain Σ(M·A) with A = 1/HW for the ablation (:60-63)
    const float norm = a.ablate ? 1.f / (float)a.HW : 1.f / (den * (1.f + 1e-6f));
    if (lane == 0)
        for (int k = 0; k < nc && k < 8; ++k) a.logits[(size_t)b * a.ncls_total + c0 + k] = acc[k] * norm;
    if (a.attn) {
        float* dst = a.attn + ((size_t)b * a.ntask + t) * a.HW;
        for (int p = lane; p < a.HW; p += 32)
            dst[p] = a.ablate ? 1.f / (float)a.HW : expf(s[(size_t)p * 64 + t] * a.inv_tau - mx) / den;
    }
}

// ------------------------------------------------------------------------------------------------ workspace
struct PmWs { size_t a_hi, a_lo, conv, act_hi[2], act_lo[2], sc, stats, scsh, feats, total; };

static PmWs pm_ws(const PmPlan& p, const PmShape& s, int Bc) {
    PmWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align256(o + bytes); return r; };
    size_t amax = 0, cmax = 0, smax = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        const size_t rows = (size_t)Bc * s.h[l + 1] * s.w[l + 1];
        amax = max(amax, rows * p.layer[l].Kp);
        cmax = max(cmax, rows * p.layer[l].cout);
        // fused InstanceNorm statistics: ≤ 4 (mean, M2) pairs per channel per M tile; an implicit-GEMM tile holds ≥ one
        // output row of an image, a plain GEMM tile 128 rows
        const size_t tiles = max((rows + 127) / 128, (size_t)Bc * s.h[l + 1]);
        smax = max(smax, tiles * 4 * p.layer[l].cout * sizeof(float2));
    }
    w.a_hi = take(amax * 2); w.a_lo = take(amax * 2);
    w.conv = take(cmax * 4);
    for (int i = 0; i < 2; ++i) { w.act_hi[i] = take(cmax * 2); w.act_lo[i] = take(cmax * 2); }
    w.sc = take((size_t)Bc * s.h[p.n_layers] * s.w[p.n_layers] * 64 * 4);
    w.stats = take(smax);
    w.scsh = take((size_t)Bc * p.layer[0].cout * sizeof(float2));
    w.total = o;
    return w;
}

template <int BN>
static int pm_gemm(const void* a_hi, const void* a_lo, int Kp, const void* wgt, int M, int N, float* out, const float* bias,
                   cudaStream_t st) {
    tc::GemmParams g{};
    g.M = M; g.N = N; g.K = Kp; g.epi = tc::EPI_F32; g.act = tc::GACT_NONE; g.ldc = N; g.out_f32 = out; g.shift = bias;
    return tc::launch_gemm_t<BN, 2, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, Kp, wgt, Kp, g, 1, Kp, st);
}

// N may exceed GEMM_MAX_N (scale/shift staging): run column panels of ≤ 512
template <int NA>
static int pm_conv_gemm(const void* a_hi, const void* a_lo, int Kp, const bf16* wgt, int M, int N, float* out, const float* bias,
                        cudaStream_t st, float2* stats = nullptr) {
    for (int n0 = 0; n0 < N; n0 += 512) {
        const int nn = min(512, N - n0);
        tc::GemmParams g{};
        g.M = M; g.N = nn; g.K = Kp; g.epi = tc::EPI_F32; g.act = tc::GACT_NONE; g.ldc = N; g.out_f32 = out + n0; g.shift = bias + n0;
        g.stats = N <= 512 ? stats : nullptr;          // statistics rows are N wide: single column panel only
        int rc;
        if (nn % 128 == 0) rc = tc::launch_gemm_t<128, NA, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, Kp, wgt + (size_t)n0 * Kp, Kp, g, 1, Kp, st);
        else if (nn % 64 == 0) rc = tc::launch_gemm_t<64, NA, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, Kp, wgt + (size_t)n0 * Kp, Kp, g, 1, Kp, st);
        else if (nn % 32 == 0) rc = tc::launch_gemm_t<32, NA, tc::EPI_F32, tc::GACT_NONE>(a_hi, a_lo, Kp, wgt + (size_t)n0 * Kp, Kp, g, 1, Kp, st);
        else rc = PMG_ERR_UNSUPPORTED;
        if (rc) return rc;
    }
    return PMG_OK;
}

// the side stream of the second pass: one per device, created on first use, never destroyed
static cudaStream_t pm_side_stream() {
    static std::mutex mu;
    static cudaStream_t side[tc::kMaxDevices] = {};
    const int dev = tc::current_device();
    std::lock_guard<std::mutex> lk(mu);
    if (!side[dev] && cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking) != cudaSuccess) side[dev] = nullptr;
    return side[dev];
}

// layer 0 fused (pm_conv0.cuh): two convolution passes around a statistics merge; writes the NHWC hi/lo planes of layer 1's input
static bool pm_conv0_supported(const PmPlan& p, const PmShape& s, int x_dtype) {
    const PmLayer& y = p.layer[0];
    const int es = x_dtype == PMG_F32 ? 4 : 2;
    return p.n_layers > 1 && y.stride == 2 && y.cin <= 4 && y.cout % 32 == 0 && y.cout <= 128 && s.w[1] <= 128 && s.w[0] == 2 * s.w[1] &&
           s.h[0] == 2 * s.h[1] && (s.w[0] * es) % 16 == 0 && 4 * y.cin * (s.w[0] * es + 32) <= tc::C0_ROWBUF_MAX && s.w[0] + 32 / es <= 256;
}

template <bool F32IN>
static int pm_conv0_run(const tc::Conv0Params& cp, const float* gamma, const float* beta, float2* scsh, cudaStream_t st) {
    constexpr int smem = tc::c0_smem_bytes<F32IN>();
    static tc::PerDeviceFlag attr_set;
    if (!attr_set.test()) {
        PMG_CUDA_OK(cudaFuncSetAttribute(tc::pm_conv0_kernel<F32IN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "conv0 smem attr");
        PMG_CUDA_OK(cudaFuncSetAttribute(tc::pm_conv0_kernel<F32IN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "conv0 smem attr");
        attr_set.set();
    }
    // image as (x, y, b·Cin + c); a box = (W + 32 bytes of pad) × 4 rows × Cin planes, out-of-bounds = zero = the conv padding
    constexpr int ES = F32IN ? 4 : 2;
    CUtensorMap tmx;
    {
        cuuint64_t dims[3] = {(cuuint64_t)cp.W, (cuuint64_t)cp.H, (cuuint64_t)cp.B * cp.Cin};
        cuuint64_t strides[2] = {(cuuint64_t)cp.W * ES, (cuuint64_t)cp.H * cp.W * ES};
        cuuint32_t box[3] = {(cuuint32_t)(cp.W + 32 / ES), 4, (cuuint32_t)cp.Cin};
        cuuint32_t estr[3] = {1, 1, 1};
        int rc = tc::encode_tmap_cached(&tmx, 3, cp.x, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE, "cuTensorMapEncodeTiled(pm conv0 image)",
                                        F32IN ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (rc) return rc;
    }
    const int tiles = cp.B * cp.Ho;
    const int grid = tiles < tc::num_sms() ? tiles : tc::num_sms();
    tc::pm_conv0_kernel<F32IN, 1><<<grid, tc::C0_THREADS, smem, st>>>(tmx, cp);
    PMG_LAUNCH_OK("pm conv0 pass 1");
    pm_conv0_finalize_kernel<<<dim3(cp.N / 32, cp.B), 256, 0, st>>>(cp.stats, gamma, beta, scsh, cp.Ho * cp.Wo, cp.N, cp.Ho, cp.Wo);
    PMG_LAUNCH_OK("pm conv0 finalize");
    tc::pm_conv0_kernel<F32IN, 2><<<grid, tc::C0_THREADS, smem, st>>>(tmx, cp);
    PMG_LAUNCH_OK("pm conv0 pass 2");
    return PMG_OK;
}

static int grid1d(size_t n) { size_t g = (n + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

}  // namespace pmg

using namespace pmg;

extern "C" {

int pm_param_count(const pm_config* cfg) {
    const PmPlan p = pm_plan(cfg);
    return p.ok ? 4 * p.n_layers + 3 * p.ntask : -1;
}

size_t pm_packed_bytes(const pm_config* cfg) {
    const PmPlan p = pm_plan(cfg);
    return p.ok ? pm_packed(p).total : 0;
}

int pm_feature_shape(const pm_config* cfg, int H, int W, int* Hf, int* Wf, int* C) {
    const PmPlan p = pm_plan(cfg);
    if (!p.ok || !Hf || !Wf || !C) return PMG_ERR_BAD_ARG;
    const PmShape s = pm_shape(p, H, W);
    if (!s.ok) return PMG_ERR_BAD_SHAPE;
    *Hf = s.h[p.n_layers]; *Wf = s.w[p.n_layers]; *C = p.C;
    return PMG_OK;
}

int pm_pack_weights(const pm_config* cfg, const float* const* prm, int n_params, void* packed, size_t packed_bytes,
                    pmg_stream_t stream) {
    const PmPlan p = pm_plan(cfg);
    if (!p.ok || !prm || !packed) return PMG_ERR_BAD_ARG;
    if (n_params != 4 * p.n_layers + 3 * p.ntask) return PMG_ERR_BAD_ARG;
    for (int i = 0; i < n_params; ++i)
        if (!prm[i]) return PMG_ERR_BAD_ARG;
    const PmPacked L = pm_packed(p);
    if (packed_bytes < L.total) return PMG_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char* base = reinterpret_cast<char*>(packed);
    int i = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        const PmLayer& y = p.layer[l];
        pm_pack_conv_kernel<<<grid1d((size_t)y.cout * y.Kp), 256, 0, st>>>(prm[i], reinterpret_cast<bf16*>(base + L.w[l]), y.cout, y.cin, y.Kp, l > 0);
        pm_copy_kernel<<<(y.cout + 255) / 256, 256, 0, st>>>(prm[i + 1], reinterpret_cast<float*>(base + L.bias[l]), y.cout);
        pm_copy_kernel<<<(y.cout + 255) / 256, 256, 0, st>>>(prm[i + 2], reinterpret_cast<float*>(base + L.gamma[l]), y.cout);
        pm_copy_kernel<<<(y.cout + 255) / 256, 256, 0, st>>>(prm[i + 3], reinterpret_cast<float*>(base + L.beta[l]), y.cout);
        i += 4;
    }
    PMG_CUDA_OK(cudaMemsetAsync(base + L.wh, 0, (size_t)64 * p.C * 2, st), "pm wh memset");
    PMG_CUDA_OK(cudaMemsetAsync(base + L.bh, 0, 64 * 4, st), "pm bh memset");
    bf16* wh = reinterpret_cast<bf16*>(base + L.wh);
    float* bh = reinterpret_cast<float*>(base + L.bh);
    for (int t = 0; t < p.ntask; ++t) {
        const int n = p.cls_off[t + 1] - p.cls_off[t];
        pm_cvt_kernel<<<(p.C + 255) / 256, 256, 0, st>>>(prm[i], wh + (size_t)t * p.C, p.C);
        pm_copy_kernel<<<1, 32, 0, st>>>(prm[i + 1], bh + t, 1);
        pm_cvt_kernel<<<(n * p.C + 255) / 256, 256, 0, st>>>(prm[i + 2], wh + (size_t)(p.ntask + p.cls_off[t]) * p.C, n * p.C);
        i += 3;
    }
    PMG_LAUNCH_OK("pm_pack");
    return PMG_OK;
}

size_t pm_workspace_bytes(const pm_config* cfg, int B, int H, int W) {
    const PmPlan p = pm_plan(cfg);
    if (!p.ok || B < 1) return 0;
    const PmShape s = pm_shape(p, H, W);
    if (!s.ok) return 0;
    const int Bc = B < g_pm_chunk ? B : g_pm_chunk;
    const int lanes = (B > Bc && g_pm_lanes > 1) ? 2 : 1;
    return lanes * align256(pm_ws(p, s, Bc).total);
}

// tuning hook (tools/): images per internal pass and passes in flight (1 or 2); affects pm_workspace_bytes
void pm_debug_set_chunking(int chunk, int lanes) {
    if (chunk >= 1) g_pm_chunk = chunk;
    if (lanes >= 1) g_pm_lanes = lanes > 2 ? 2 : lanes;
}

int pm_forward(const pm_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W, float* logits,
               float* attn, float* feats, void* workspace, size_t workspace_bytes, unsigned flags, pmg_stream_t stream) {
    const PmPlan p = pm_plan(cfg);
    if (!p.ok || !packed || !x || !logits || !workspace) return PMG_ERR_BAD_ARG;
    if (x_dtype != PMG_F32 && x_dtype != PMG_BF16) return PMG_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(packed)) & 15) return PMG_ERR_BAD_ARG;
    if (B < 1) return PMG_ERR_BAD_SHAPE;
    const PmShape s = pm_shape(p, H, W);
    if (!s.ok) return PMG_ERR_BAD_SHAPE;
    {
        int dev = 0, major = 0;
        PMG_CUDA_OK(cudaGetDevice(&dev), "getdevice");
        PMG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev), "attr");
        if (major != 10) return PMG_ERR_ARCH;
    }
    const int chunk = g_pm_chunk;
    const int Bc_max = B < chunk ? B : chunk;
    const PmWs w = pm_ws(p, s, Bc_max);
    const int lanes = (B > Bc_max && g_pm_lanes > 1) ? 2 : 1;
    const size_t lane_bytes = align256(w.total);
    if (workspace_bytes < lanes * lane_bytes) return PMG_ERR_WORKSPACE;
    const PmPacked L = pm_packed(p);
    const char* pk = reinterpret_cast<const char*>(packed);
    cudaStream_t st0 = reinterpret_cast<cudaStream_t>(stream);
    // second pass lane: an internal stream forked from / joined to the caller's stream with events (capturable)
    cudaStream_t lane_st[2] = {st0, st0};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (lanes == 2) {
        cudaStream_t side = pm_side_stream();
        if (!side) { set_cuda_error(cudaErrorUnknown, "pm side stream"); return PMG_ERR_CUDA; }
        PMG_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming), "pm fork event");
        PMG_CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming), "pm join event");
        PMG_CUDA_OK(cudaEventRecord(ev_fork, st0), "pm fork record");
        PMG_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0), "pm fork wait");
        lane_st[1] = side;
    }
    const int Hf = s.h[p.n_layers], Wf = s.w[p.n_layers], HWf = Hf * Wf;
    const size_t img_elems = (size_t)cfg->input_nc * H * W;
    // the passes: a lambda, so that the side stream is joined to the caller's on EVERY exit path (an error return that left
    // the fork un-joined would leave the caller's stream — or a graph capture — depending on a dangling branch)
    auto run_passes = [&]() -> int {
    int ci = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++ci) {
        const int Bc = min(chunk, B - b0);
        char* ws = reinterpret_cast<char*>(workspace) + (size_t)(ci % lanes) * lane_bytes;
        cudaStream_t st = lane_st[ci % lanes];
        auto Hb = [&](size_t off) { return reinterpret_cast<bf16*>(ws + off); };
        const char* xb = reinterpret_cast<const char*>(x) + (size_t)b0 * img_elems * (x_dtype == PMG_F32 ? 4 : 2);
        int rc;
        for (int l = 0; l < p.n_layers; ++l) {
            const PmLayer& y = p.layer[l];
            const int Ho = s.h[l + 1], Wo = s.w[l + 1];
            const int rows = Bc * Ho * Wo;
            if (l == 0 && !(flags & PM_FLAG_NO_FUSED_CONV0) && pm_conv0_supported(p, s, x_dtype)) {
                tc::Conv0Params cp{};
                cp.x = xb; cp.wgt = reinterpret_cast<const bf16*>(pk + L.w[0]);      // the conv bias cancels in InstanceNorm
                cp.stats = reinterpret_cast<float2*>(ws + w.stats); cp.scsh = reinterpret_cast<const float2*>(ws + w.scsh);
                cp.o_hi = Hb(w.act_hi[0]); cp.o_lo = Hb(w.act_lo[0]);
                cp.trace = g_enc_trace; cp.B = Bc; cp.Cin = y.cin; cp.H = H; cp.W = W; cp.Ho = Ho; cp.Wo = Wo; cp.N = y.cout;
                const float* ga = reinterpret_cast<const float*>(pk + L.gamma[0]);
                const float* be = reinterpret_cast<const float*>(pk + L.beta[0]);
                rc = x_dtype == PMG_F32 ? pm_conv0_run<true>(cp, ga, be, reinterpret_cast<float2*>(ws + w.scsh), st)
                                        : pm_conv0_run<false>(cp, ga, be, reinterpret_cast<float2*>(ws + w.scsh), st);
                if (rc) return rc;
                continue;
            }
            float* conv = reinterpret_cast<float*>(ws + w.conv);
            // InstanceNorm statistics from the convolution's own epilogue wherever every M tile lies inside one sample
            float2* stats = (flags & PM_FLAG_NO_FUSED_STATS) ? nullptr : reinterpret_cast<float2*>(ws + w.stats);
            int st_tiles = 0, st_rows_full = 128, st_rows_last = 128;
            // layers with ≥ 64 input channels: implicit GEMM straight from the previous layer's NHWC hi/lo planes
            // (tc_igemm.cuh) — no im2col planes are written or re-read.  Layer 0 (3 input channels, NCHW image) keeps
            // im2col + GEMM.
            const bool igemm = l > 0 && y.cin % 64 == 0 && y.cout % 128 == 0 && y.cout <= tc::IG_MAX_N &&
                               Wo <= 128 && !(flags & PM_FLAG_NO_IGEMM);
            if (igemm) {
                if ((rc = tc::launch_conv_igemm(Hb(w.act_hi[(l - 1) & 1]), Hb(w.act_lo[(l - 1) & 1]), pk + L.w[l],
                                                reinterpret_cast<const float*>(pk + L.bias[l]), Bc, s.h[l], s.w[l], y.cin, Ho, Wo,
                                                y.stride, y.cout, conv, st, (flags & PM_FLAG_IGEMM_N128) != 0, stats, &st_tiles,
                                                &st_rows_full)))
                    return rc;
                st_rows_last = Ho * Wo - (st_tiles - 1) * st_rows_full;
            } else {
            const size_t row_smem = (size_t)y.cin * 4 * (s.w[l] + 2) * sizeof(float);
            if (l == 0 && row_smem <= 48 * 1024) {
                if (x_dtype == PMG_F32)
                    pm_im2col_first_rows_kernel<true><<<Bc * Ho, 256, row_smem, st>>>(xb, Hb(w.a_hi), Hb(w.a_lo), y.cin, H, W, Ho, Wo, y.stride, y.Kp);
                else
                    pm_im2col_first_rows_kernel<false><<<Bc * Ho, 256, row_smem, st>>>(xb, Hb(w.a_hi), Hb(w.a_lo), y.cin, H, W, Ho, Wo, y.stride, y.Kp);
            } else if (l == 0) {
                const size_t n = (size_t)rows * (y.Kp / 8);
                if (x_dtype == PMG_F32)
                    pm_im2col_first_kernel<true><<<grid1d(n), 256, 0, st>>>(xb, Hb(w.a_hi), Hb(w.a_lo), Bc, y.cin, H, W, Ho, Wo, y.stride, y.Kp);
                else
                    pm_im2col_first_kernel<false><<<grid1d(n), 256, 0, st>>>(xb, Hb(w.a_hi), Hb(w.a_lo), Bc, y.cin, H, W, Ho, Wo, y.stride, y.Kp);
            } else {
                const size_t n = (size_t)rows * 16 * (y.cin / 8);
                pm_im2col_kernel<<<grid1d(n), 256, 0, st>>>(Hb(w.act_hi[(l - 1) & 1]), Hb(w.act_lo[(l - 1) & 1]), Hb(w.a_hi), Hb(w.a_lo), Bc,
                                                            y.cin, s.h[l], s.w[l], Ho, Wo, y.stride);
            }
            PMG_LAUNCH_OK("pm im2col");
            // plain GEMM tiles are 128 consecutive rows of [B·Ho·Wo]: inside one sample iff 128 divides Ho·Wo
            float2* gstats = (stats && (Ho * Wo) % 128 == 0 && y.cout <= 512) ? stats : nullptr;
            if (gstats) st_tiles = Ho * Wo / 128;
            // a bf16 image is exact in one plane: layer 0 then runs without the (all-zero) lo plane
            if (l == 0 && x_dtype == PMG_BF16)
                rc = pm_conv_gemm<1>(Hb(w.a_hi), nullptr, y.Kp, reinterpret_cast<const bf16*>(pk + L.w[l]), rows, y.cout, conv,
                                     reinterpret_cast<const float*>(pk + L.bias[l]), st, gstats);
            else
                rc = pm_conv_gemm<2>(Hb(w.a_hi), Hb(w.a_lo), y.Kp, reinterpret_cast<const bf16*>(pk + L.w[l]), rows, y.cout, conv,
                                     reinterpret_cast<const float*>(pk + L.bias[l]), st, gstats);
            if (rc) return rc;
            }
            const bool last = l == p.n_layers - 1;
            float* f32_out = (last && feats) ? feats + (size_t)b0 * HWf * p.C : nullptr;
            if (st_tiles > 0)
                pm_instnorm_fused_kernel<<<dim3(y.cout / 32, Bc), 256, 0, st>>>(conv, stats, reinterpret_cast<const float*>(pk + L.gamma[l]),
                                                                                reinterpret_cast<const float*>(pk + L.beta[l]),
                                                                                Hb(w.act_hi[l & 1]), Hb(w.act_lo[l & 1]), f32_out, Ho * Wo,
                                                                                y.cout, st_tiles, st_rows_full, st_rows_last);
            else
                pm_instnorm_kernel<<<dim3(y.cout / 32, Bc), 256, 0, st>>>(conv, reinterpret_cast<const float*>(pk + L.gamma[l]),
                                                                          reinterpret_cast<const float*>(pk + L.beta[l]), Hb(w.act_hi[l & 1]),
                                                                          Hb(w.act_lo[l & 1]), f32_out, Ho * Wo, y.cout);
            PMG_LAUNCH_OK("pm instnorm");
        }
        const int lf = (p.n_layers - 1) & 1;
        float* sc = reinterpret_cast<float*>(ws + w.sc);
        if ((rc = pm_gemm<64>(Hb(w.act_hi[lf]), Hb(w.act_lo[lf]), p.C, pk + L.wh, Bc * HWf, 64, sc,
                              reinterpret_cast<const float*>(pk + L.bh), st)))
            return rc;
        PmFinishArgs a;
        a.sc = sc; a.logits = logits + (size_t)b0 * p.ncls_total; a.attn = attn ? attn + (size_t)b0 * p.ntask * HWf : nullptr;
        a.B = Bc; a.HW = HWf; a.ntask = p.ntask; a.ncls_total = p.ncls_total; a.ablate = cfg->ablate_attention;
        a.inv_tau = 1.f / cfg->attn_tau;
        for (int t = 0; t <= p.ntask; ++t) a.cls_off[t] = p.cls_off[t];
        pm_heads_finish_kernel<<<(Bc * p.ntask + 3) / 4, 128, 0, st>>>(a);
        PMG_LAUNCH_OK("pm heads finish");
    }
    return PMG_OK;
    };
    const int rc_passes = run_passes();
    if (lanes == 2) {
        const cudaError_t e1 = cudaEventRecord(ev_join, lane_st[1]);
        const cudaError_t e2 = cudaStreamWaitEvent(st0, ev_join, 0);
        cudaEventDestroy(ev_fork);          // released by the runtime once the recorded work has completed
        cudaEventDestroy(ev_join);
        if (rc_passes == PMG_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
            set_cuda_error(e1 != cudaSuccess ? e1 : e2, "pm join");
            return PMG_ERR_CUDA;
        }
    }
    return rc_passes;
}

}  // extern "C"
