// pmg_generic.cu — shape-generic fp32 CUDA path of the PMG forward (any ndf / patch_size / patch_div /
// gram_channels / d_model / layers / heads / tasks / H×W the reference ctor accepts with norm="batch").
// It is the correctness anchor on the device and the route for shapes the tcgen05 fast path does not
// cover.  Everything is CUDA-core fp32 (fp32 accumulate, same operation order class as the reference),
// activations pixel-major (NHWC) so every conv is one GEMM over pixels.
//
// Reference: /root/reference/Models/models_PMG.py:176-199 (forward), line cites per kernel.
#include "pmg_common.cuh"

namespace pmg {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
// out[k'][n] = in[n][k]; mode 0: k' = k = (c,kk);  mode 1: k' = (kk,c)  (NHWC im2col order)
__global__ void pack_wt_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C, int KK, int mode) {
    const int K = C * KK;
    const size_t total = (size_t)K * N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kp = (int)(i / N), n = (int)(i % N);
        const int k = mode == 0 ? kp : ((kp % C) * KK + kp / C);
        out[i] = in[(size_t)n * K + k];
    }
}

// BatchNorm2d(eval) → y = x*scale + shift  (models_PMG.py:116 with running stats)
__global__ void bn_fold_kernel(const float* g, const float* b, const float* m, const float* v, float* scale,
                               float* shift, int n, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float s = g[i] / sqrtf(v[i] + eps);
        scale[i] = s;
        shift[i] = b[i] - m[i] * s;
    }
}

__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

static inline int grid_for(size_t n, int block = 256) {
    size_t g = (n + block - 1) / block;
    return (int)(g > 4096 ? 4096 : (g < 1 ? 1 : g));
}

int pack_generic(const Plan& p, const PackedLayout& L, const float* const* prm, char* packed, cudaStream_t st) {
    int i = 0;
    auto F = [&](size_t off) { return reinterpret_cast<float*>(packed + off); };
    for (int l = 0; l < p.n_layers; ++l) {
        const ConvLayer& c = p.layer[l];
        const int KK = c.k * c.k;
        pack_wt_kernel<<<grid_for((size_t)c.cin * KK * c.cout), 256, 0, st>>>(prm[i], F(L.conv_w[l]), c.cout, c.cin, KK,
                                                                             (l > 0 && c.k == 4) ? 1 : 0);
        if (p.norm == PMG_NORM_BATCH) {
            bn_fold_kernel<<<(c.cout + 127) / 128, 128, 0, st>>>(prm[i + 1], prm[i + 2], prm[i + 3], prm[i + 4],
                                                                 F(L.conv_scale[l]), F(L.conv_shift[l]), c.cout, kBnEps);
            i += 5;
        } else {                                  // instance / group norm: per-sample statistics at run time; keep gamma, beta
            copy_f32_kernel<<<grid_for(c.cout), 256, 0, st>>>(prm[i + 1], F(L.conv_scale[l]), c.cout);
            copy_f32_kernel<<<grid_for(c.cout), 256, 0, st>>>(prm[i + 2], F(L.conv_shift[l]), c.cout);
            i += 3;
        }
    }
    if (p.se) {
        copy_f32_kernel<<<grid_for((size_t)p.se_hidden * p.C), 256, 0, st>>>(prm[i], F(L.se_w1), (size_t)p.se_hidden * p.C);
        copy_f32_kernel<<<grid_for((size_t)p.se_hidden * p.C), 256, 0, st>>>(prm[i + 1], F(L.se_w2), (size_t)p.se_hidden * p.C);
        i += 2;
    }
    pack_wt_kernel<<<grid_for((size_t)p.C * p.Cr), 256, 0, st>>>(prm[i], F(L.chan_w), p.Cr, p.C, 1, 0); i += 1;
    pack_wt_kernel<<<grid_for((size_t)p.Cr * p.Cr * p.d), 256, 0, st>>>(prm[i], F(L.tok_w), p.d, p.Cr * p.Cr, 1, 0); i += 1;
    copy_f32_kernel<<<1, 256, 0, st>>>(prm[i], F(L.tok_b), p.d); i += 1;
    for (int l = 0; l < p.nl; ++l) {
        const TfLayerOff& t = L.tf[l];
        const int d = p.d;
        pack_wt_kernel<<<grid_for((size_t)3 * d * d), 256, 0, st>>>(prm[i], F(t.w_in), 3 * d, d, 1, 0);
        copy_f32_kernel<<<2, 256, 0, st>>>(prm[i + 1], F(t.b_in), 3 * d);
        pack_wt_kernel<<<grid_for((size_t)d * d), 256, 0, st>>>(prm[i + 2], F(t.w_o), d, d, 1, 0);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 3], F(t.b_o), d);
        pack_wt_kernel<<<grid_for((size_t)2 * d * d), 256, 0, st>>>(prm[i + 4], F(t.w_1), 2 * d, d, 1, 0);
        copy_f32_kernel<<<2, 256, 0, st>>>(prm[i + 5], F(t.b_1), 2 * d);
        pack_wt_kernel<<<grid_for((size_t)2 * d * d), 256, 0, st>>>(prm[i + 6], F(t.w_2), d, 2 * d, 1, 0);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 7], F(t.b_2), d);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 8], F(t.ln1w), d);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 9], F(t.ln1b), d);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 10], F(t.ln2w), d);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 11], F(t.ln2b), d);
        i += 12;
    }
    for (int t = 0; t < p.ntask; ++t) copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + t], F(L.q) + (size_t)t * p.d, p.d);
    i += p.ntask;
    for (int t = 0; t < p.ntask; ++t) {
        const int n = p.cls_off[t + 1] - p.cls_off[t];
        copy_f32_kernel<<<grid_for((size_t)n * p.d), 256, 0, st>>>(prm[i], F(L.cls_w) + (size_t)p.cls_off[t] * p.d, (size_t)n * p.d);
        copy_f32_kernel<<<1, 256, 0, st>>>(prm[i + 1], F(L.cls_b) + p.cls_off[t], n);
        i += 2;
    }
    PMG_LAUNCH_OK("pack_generic");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// GEMM over pixels / tokens:  C[M,N] = act((A[M,K] · Wt[K,N]) * scale[n] + shift[n]) (+ resid)
// A is read through a gather that implements the k=s=4 space-to-depth view of a conv input
// (models_PMG.py:115: Conv2d(k=4,s=4,p=0,bias=False) == GEMM over non-overlapping 4×4 windows).
// ------------------------------------------------------------------------------------------------
enum AMode { A_PLAIN = 0, A_S2D_NCHW_F32 = 1, A_S2D_NCHW_BF16 = 2, A_S2D_NHWC = 3 };
enum Act { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };

struct GemmArgs {
    const void* A; const float* Wt; float* C;
    const float* scale; const float* shift; const float* resid;
    int M, N, K, lda, ldc, act;
    int Cin, Hi, Wi, Ho, Wo;   // gather geometry (input map Hi×Wi, output map Ho×Wo)
};

constexpr int GBM = 64, GBN = 64, GBK = 16;

template <int MODE>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
    __shared__ float As[GBK][GBM + 4];
    __shared__ float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    // A loader: this thread owns row (tid/4) and 4 consecutive k at (tid%4)*4
    const int am = tid >> 2, ak = (tid & 3) * 4;
    const int gm = m0 + am;
    size_t abase = 0;
    if (gm < g.M) {
        if (MODE == A_PLAIN) abase = (size_t)gm * g.lda;
        else {
            const int hw = g.Ho * g.Wo;
            const int b = gm / hw, r = gm % hw, yo = r / g.Wo, xo = r % g.Wo;
            if (MODE == A_S2D_NHWC) abase = (((size_t)b * g.Hi + 4 * yo) * g.Wi + 4 * xo) * g.Cin;
            else abase = (((size_t)b * g.Cin) * g.Hi + 4 * yo) * g.Wi + 4 * xo;
        }
    }
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < g.K; k0 += GBK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + ak + j;
            float v = 0.f;
            if (gm < g.M && k < g.K) {
                if (MODE == A_PLAIN) v = reinterpret_cast<const float*>(g.A)[abase + k];
                else if (MODE == A_S2D_NHWC) {
                    const int kk = k / g.Cin, c = k - kk * g.Cin;
                    v = reinterpret_cast<const float*>(g.A)[abase + ((size_t)(kk >> 2) * g.Wi + (kk & 3)) * g.Cin + c];
                } else {
                    const int c = k >> 4, ky = (k >> 2) & 3, kx = k & 3;
                    const size_t idx = abase + ((size_t)c * g.Hi + ky) * g.Wi + kx;
                    if (MODE == A_S2D_NCHW_F32) v = reinterpret_cast<const float*>(g.A)[idx];
                    else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g.A)[idx]);
                }
            }
            As[ak + j][am] = v;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + bk, n = n0 + bn + j;
            Bs[bk][bn + j] = (k < g.K && n < g.N) ? g.Wt[(size_t)k * g.N + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.scale) v *= g.scale[n];
            if (g.shift) v += g.shift[n];
            if (g.act == ACT_LEAKY) v = v >= 0.f ? v : v * kLeaky;
            else if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
            if (g.resid) v += g.resid[(size_t)m * g.ldc + n];
            g.C[(size_t)m * g.ldc + n] = v;
        }
    }
}

static int launch_sgemm(int mode, const GemmArgs& g, cudaStream_t st) {
    dim3 grid((g.M + GBM - 1) / GBM, (g.N + GBN - 1) / GBN);
    switch (mode) {
        case A_PLAIN: sgemm_kernel<A_PLAIN><<<grid, 256, 0, st>>>(g); break;
        case A_S2D_NCHW_F32: sgemm_kernel<A_S2D_NCHW_F32><<<grid, 256, 0, st>>>(g); break;
        case A_S2D_NCHW_BF16: sgemm_kernel<A_S2D_NCHW_BF16><<<grid, 256, 0, st>>>(g); break;
        default: sgemm_kernel<A_S2D_NHWC><<<grid, 256, 0, st>>>(g); break;
    }
    PMG_LAUNCH_OK("sgemm");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// Per-patch Gram: G[b,k] = P·Pᵀ / N  with the replicate padding folded in as an index clamp
// (models_PMG.py:157-174 pad + unfold, :188 Gram).  chan is pixel-major (B,Hc,Wc,Cr).
// One block per (image, patch); each thread owns 4×4 micro-tiles of the Cr×Cr output.
// ------------------------------------------------------------------------------------------------
constexpr int GRAM_CHUNK = 32;

__global__ void __launch_bounds__(256) gram_patch_kernel(const float* __restrict__ chan, float* __restrict__ G, int Hc,
                                                         int Wc, int Cr, int npy, int npx, int ph, int pw, int flat_w) {
    extern __shared__ float sP[];   // [GRAM_CHUNK][Cr+1]
    const int np = npy * npx, n = ph * pw;
    const int b = blockIdx.x / np, k = blockIdx.x % np;
    const int py = k / npx, px = k % npx;
    const int ld = Cr + 1;
    const int tiles1 = (Cr + 3) / 4, ntiles = tiles1 * tiles1;
    const float inv = 1.f / (float)n;
    float* Gout = G + ((size_t)b * np + k) * Cr * Cr;
    for (int pass = 0; pass * (int)blockDim.x < ntiles; ++pass) {
        const int tile = pass * blockDim.x + threadIdx.x;
        const int ti = (tile / tiles1) * 4, tj = (tile % tiles1) * 4;
        float acc[4][4] = {};
        for (int c0 = 0; c0 < n; c0 += GRAM_CHUNK) {
            const int cn = min(GRAM_CHUNK, n - c0);
            __syncthreads();
            for (int e = threadIdx.x; e < cn * Cr; e += blockDim.x) {
                const int pi = c0 + e / Cr, ch = e % Cr;
                int y, x;
                if (flat_w > 0) {       // style-transfer export: "patch" k = pixels [k·n, (k+1)·n) of the cropped map, row-major
                    const int f = k * n + pi;
                    y = f / flat_w; x = f - y * flat_w;
                } else {
                    y = min(py * ph + pi / pw, Hc - 1); x = min(px * pw + pi % pw, Wc - 1);
                }
                sP[(e / Cr) * ld + ch] = chan[(((size_t)b * Hc + y) * Wc + x) * Cr + ch];
            }
            __syncthreads();
            if (tile < ntiles) {
                for (int q = 0; q < cn; ++q) {
                    float a[4], c[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        a[i] = (ti + i < Cr) ? sP[q * ld + ti + i] : 0.f;
                        c[i] = (tj + i < Cr) ? sP[q * ld + tj + i] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], c[j], acc[i][j]);
                }
            }
        }
        if (tile < ntiles) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (ti + i < Cr && tj + j < Cr) Gout[(size_t)(ti + i) * Cr + tj + j] = acc[i][j] * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row (TransformerEncoderLayer norm1/norm2, eps 1e-5)
// ------------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ y, int rows, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * d;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s += xr[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / d;
    float v = 0.f;
    for (int i = lane; i < d; i += 32) { const float t = xr[i] - mu; v += t * t; }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rs = 1.f / sqrtf(v / d + eps);
    for (int i = lane; i < d; i += 32) y[(size_t)row * d + i] = (xr[i] - mu) * rs * w[i] + b[i];
}

// ------------------------------------------------------------------------------------------------
// Multi-head self-attention core over the Np tokens of one image (nn.MultiheadAttention inside
// TransformerEncoderLayer, models_PMG.py:64-68): softmax(q·kᵀ/√hd)·v.  Block = (image, head).
// qkv is (B*N, 3d) [q | k | v]; out is (B*N, d).
// ------------------------------------------------------------------------------------------------
constexpr int ATT_MAX_HD = 128;

__global__ void attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int d, int nh) {
    extern __shared__ float sm[];   // q,k,v [N][hd] each
    const int b = blockIdx.x, h = blockIdx.y, hd = d / nh;
    float* sq = sm; float* sk = sq + N * hd; float* sv = sk + N * hd;
    for (int e = threadIdx.x; e < N * hd; e += blockDim.x) {
        const int t = e / hd, c = e % hd;
        const float* r = qkv + ((size_t)b * N + t) * 3 * d + h * hd + c;
        sq[e] = r[0]; sk[e] = r[d]; sv[e] = r[2 * d];
    }
    __syncthreads();
    const float sc = 1.f / sqrtf((float)hd);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float mx = -INFINITY;
        for (int j = 0; j < N; ++j) {
            float s = 0.f;
            for (int c = 0; c < hd; ++c) s = fmaf(sq[i * hd + c], sk[j * hd + c], s);
            mx = fmaxf(mx, s * sc);
        }
        float o[ATT_MAX_HD];
        for (int c = 0; c < hd; ++c) o[c] = 0.f;
        float den = 0.f;
        for (int j = 0; j < N; ++j) {
            float s = 0.f;
            for (int c = 0; c < hd; ++c) s = fmaf(sq[i * hd + c], sk[j * hd + c], s);
            const float e = expf(s * sc - mx);
            den += e;
            for (int c = 0; c < hd; ++c) o[c] = fmaf(e, sv[j * hd + c], o[c]);
        }
        const float inv = 1.f / den;
        float* orow = out + ((size_t)b * N + i) * d + h * hd;
        for (int c = 0; c < hd; ++c) orow[c] = o[c] * inv;
    }
}

// ------------------------------------------------------------------------------------------------
// TaskAttentionPooling + classifiers (models_PMG.py:47-57 and :198).  Block = image, warp = task.
//   s_n = tok_n·q_t (no 1/√d),  w = softmax_n(s),  pooled = Σ w_n tok_n,  logits = W_t pooled + b_t
// ------------------------------------------------------------------------------------------------
struct PoolArgs {
    const float* tok; const float* q; const float* cw; const float* cb; float* logits;
    int N, d, ntask, ncls_total;
    int cls_off[PMG_MAX_TASKS + 1];
};

__global__ void pool_heads_kernel(const PoolArgs a) {
    extern __shared__ float sm[];
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stok = sm;                          // [N][d]
    float* spool = stok + a.N * a.d;           // [nw][d]
    float* ssc = spool + nw * a.d;             // [nw][N]
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < a.N * a.d; e += blockDim.x) stok[e] = a.tok[(size_t)b * a.N * a.d + e];
    __syncthreads();
    for (int t = warp; t < a.ntask; t += nw) {
        const float* q = a.q + (size_t)t * a.d;
        float mx = -INFINITY;
        for (int n = 0; n < a.N; ++n) {
            float s = 0.f;
            for (int i = lane; i < a.d; i += 32) s = fmaf(stok[n * a.d + i], q[i], s);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) ssc[warp * a.N + n] = s;
            mx = fmaxf(mx, s);
        }
        __syncwarp();
        float den = 0.f;
        for (int n = 0; n < a.N; ++n) den += expf(ssc[warp * a.N + n] - mx);
        const float inv = 1.f / den;
        for (int i = lane; i < a.d; i += 32) {
            float pv = 0.f;
            for (int n = 0; n < a.N; ++n) pv = fmaf(expf(ssc[warp * a.N + n] - mx) * inv, stok[n * a.d + i], pv);
            spool[warp * a.d + i] = pv;
        }
        __syncwarp();
        for (int c = a.cls_off[t]; c < a.cls_off[t + 1]; ++c) {
            float s = 0.f;
            for (int i = lane; i < a.d; i += 32) s = fmaf(spool[warp * a.d + i], a.cw[(size_t)c * a.d + i], s);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) a.logits[(size_t)b * a.ncls_total + c] = s + a.cb[c];
        }
        __syncwarp();
    }
}

int launch_pool_heads(const Plan& p, const PackedLayout& L, const char* packed, const float* tok, int B, int N,
                      float* logits, cudaStream_t st) {
    PoolArgs a;
    a.tok = tok; a.q = reinterpret_cast<const float*>(packed + L.q);
    a.cw = reinterpret_cast<const float*>(packed + L.cls_w); a.cb = reinterpret_cast<const float*>(packed + L.cls_b);
    a.logits = logits; a.N = N; a.d = p.d; a.ntask = p.ntask; a.ncls_total = p.ncls_total;
    for (int t = 0; t <= p.ntask; ++t) a.cls_off[t] = p.cls_off[t];
    const int threads = 128, nw = threads / 32;
    const size_t smem = ((size_t)N * p.d + (size_t)nw * p.d + (size_t)nw * N) * sizeof(float);
    if (smem > 200 * 1024) return PMG_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        PMG_CUDA_OK(cudaFuncSetAttribute(pool_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "pool attr");
    pool_heads_kernel<<<B, threads, smem, st>>>(a);
    PMG_LAUNCH_OK("pool_heads");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// generic forward
// ------------------------------------------------------------------------------------------------
struct GenericWs {
    size_t act[2], chan, gram, tok, h, qkv, att, ffh, total;
};

static GenericWs generic_ws(const Plan& p, const Shape& s, int B) {
    GenericWs w{};
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o = align256(o + n * sizeof(float)); return r; };
    size_t amax = 0;
    for (int l = 0; l < p.n_layers; ++l) {
        const size_t n = (size_t)B * s.h[l + 1] * s.w[l + 1] * p.layer[l].cout;
        amax = n > amax ? n : amax;
    }
    w.act[0] = take(amax); w.act[1] = take(amax);
    w.chan = take((size_t)B * s.Hc * s.Wc * p.Cr);
    w.gram = take((size_t)B * s.np * p.Cr * p.Cr);
    const size_t T = (size_t)B * s.np;
    w.tok = take(T * p.d); w.h = take(T * p.d); w.qkv = take(T * 3 * p.d); w.att = take(T * p.d); w.ffh = take(T * 2 * p.d);
    w.total = o;
    return w;
}

size_t generic_workspace_bytes(const Plan& p, const Shape& s, int B) { return generic_ws(p, s, B).total; }

int generic_launch_count(const Plan& p) {
    return p.n_layers * (p.norm == PMG_NORM_BATCH ? 1 : 2) + (p.se ? 1 : 0) + 3 + 7 * p.nl + 1;
}

// ------------------------------------------------------------------------------------------------
// norm = "instance" (models_PMG.py:100-101, InstanceNorm2d(affine=True)): per-sample statistics over HW (x cg channels;
// cg = 1 here), biased variance, eps 1e-5, per-channel affine, then LeakyReLU(0.2); in place on the pixel-major
// (B, HW, C) fp32 conv output.  One block per (sample, channel group).  Two-pass statistics
// (mean, then centred sum of squares) in fp32 with a block reduction — a correctness path, not a tuned one.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(256) sample_norm_lrelu_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int HW, int C, int cg) {
    __shared__ float red[8];
    const int groups = C / cg, b = blockIdx.x / groups, g = blockIdx.x % groups;
    float* base = x + (size_t)b * HW * C + g * cg;
    const int n = HW * cg;
    float s = 0.f;
    for (int e = threadIdx.x; e < n; e += 256) s += base[(size_t)(e / cg) * C + e % cg];
    const float mean = block_sum_256(s, red) / (float)n;
    float q = 0.f;
    for (int e = threadIdx.x; e < n; e += 256) { const float d = base[(size_t)(e / cg) * C + e % cg] - mean; q = fmaf(d, d, q); }
    const float rstd = 1.f / sqrtf(block_sum_256(q, red) / (float)n + 1e-5f);
    for (int e = threadIdx.x; e < n; e += 256) {
        const int c = e % cg;
        float* px = base + (size_t)(e / cg) * C + c;
        const float v = (*px - mean) * rstd * gamma[g * cg + c] + beta[g * cg + c];
        *px = v >= 0.f ? v : v * kLeaky;
    }
}

// SEBlock (models_PMG.py:15-30): w = sigmoid(W2 · relu(W1 · mean_hw(x)));  x *= w.  One block per sample.
__global__ void __launch_bounds__(256) se_scale_kernel(float* __restrict__ x, const float* __restrict__ w1,
                                                       const float* __restrict__ w2, int HW, int C, int hid) {
    extern __shared__ float s_se[];          // pooled[C] | hidden[hid] | gate[C]
    float* pooled = s_se; float* hidden = s_se + C; float* gate = hidden + hid;
    float* base = x + (size_t)blockIdx.x * HW * C;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += base[(size_t)p * C + c];
        pooled[c] = s / (float)HW;
    }
    __syncthreads();
    for (int h = threadIdx.x; h < hid; h += 256) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)h * C + c], pooled[c], s);
        hidden[h] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int h = 0; h < hid; ++h) s = fmaf(w2[(size_t)c * hid + h], hidden[h], s);
        gate[c] = 1.f / (1.f + expf(-s));
    }
    __syncthreads();
    for (size_t e = threadIdx.x; e < (size_t)HW * C; e += 256) base[e] *= gate[e % C];
}

// feature_extractor (models_PMG.py:113-128,178): conv k4 s4 / 1x1 as GEMMs with the folded eval-BatchNorm + LeakyReLU
// epilogue; *out = the last activation, pixel-major (B, Hc, Wc, C) fp32, inside the workspace
static int run_backbone(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x, int x_dtype,
                        int B, char* ws, const GenericWs& w, cudaStream_t st, StageTimer* tm, const float** out, bool apply_se) {
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    auto P = [&](size_t off) { return reinterpret_cast<const float*>(packed + off); };
    static const char* kConvName[PMG_MAX_BACKBONE] = {"conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7"};
    const float* cur = nullptr;
    int rc;
    for (int l = 0; l < p.n_layers; ++l) {
        const ConvLayer& c = p.layer[l];
        GemmArgs g{};
        const bool folded = p.norm == PMG_NORM_BATCH;
        g.Wt = P(L.conv_w[l]); g.scale = folded ? P(L.conv_scale[l]) : nullptr; g.shift = folded ? P(L.conv_shift[l]) : nullptr;
        g.resid = nullptr;
        g.C = F(w.act[l & 1]);
        g.M = B * s.h[l + 1] * s.w[l + 1]; g.N = c.cout; g.K = c.cin * c.k * c.k; g.ldc = c.cout; g.act = folded ? ACT_LEAKY : ACT_NONE;
        g.Cin = c.cin; g.Hi = s.h[l]; g.Wi = s.w[l]; g.Ho = s.h[l + 1]; g.Wo = s.w[l + 1];
        int mode;
        if (l == 0) { g.A = x; mode = x_dtype == PMG_BF16 ? A_S2D_NCHW_BF16 : A_S2D_NCHW_F32; }
        else { g.A = cur; g.lda = c.cin; mode = c.k == 4 ? A_S2D_NHWC : A_PLAIN; }
        if ((rc = launch_sgemm(mode, g, st))) return rc;
        if (!folded) {
            const int cg = 1;
            sample_norm_lrelu_kernel<<<B * (c.cout / cg), 256, 0, st>>>(g.C, P(L.conv_scale[l]), P(L.conv_shift[l]),
                                                                        s.h[l + 1] * s.w[l + 1], c.cout, cg);
            PMG_LAUNCH_OK("sample_norm");
        }
        PMG_MARK(tm, kConvName[l]);
        cur = g.C;
    }
    if (p.se) {                                              // models_PMG.py:179 (style-transfer export: feature_extractor only)
        const size_t smem = (size_t)(2 * p.C + p.se_hidden) * sizeof(float);
        if (smem > 48 * 1024) return PMG_ERR_UNSUPPORTED;
        if (apply_se) {
            se_scale_kernel<<<B, 256, smem, st>>>(const_cast<float*>(cur), P(L.se_w1), P(L.se_w2), s.Hc * s.Wc, p.C, p.se_hidden);
            PMG_LAUNCH_OK("se_scale");
            PMG_MARK(tm, "se");
        }
    }
    *out = cur;
    return PMG_OK;
}

// Style-transfer Gram export (utils/PMG/PMG_style_transfer.py:11-48): Gram over the feature_extractor output (C channels),
// UNPADDED, (Hc/ph)·(Wc/pw) "patches" of ph·pw pixels each.  The reference's permute + reshape (:37) makes patch k the run
// of ph·pw consecutive pixels of the cropped map in row-major order (not the k-th spatial window); that behaviour is the
// contract.  Shape-generic fp32 kernels (the reference calls it one image at a time).
int style_grams_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x, int x_dtype,
                        int B, float* grams, char* ws, cudaStream_t st) {
    if (p.layer[0].k != 4) return PMG_ERR_UNSUPPORTED;
    const GenericWs w = generic_ws(p, s, B);
    const float* feats = nullptr;
    int rc;
    if ((rc = run_backbone(p, s, L, packed, x, x_dtype, B, ws, w, st, nullptr, &feats, false))) return rc;
    const int ph = s.Hc / p.pdiv, pw = s.Wc / p.pdiv;
    if (ph < 1 || pw < 1) return PMG_ERR_BAD_SHAPE;
    const int npy = s.Hc / ph, npx = s.Wc / pw;
    const size_t smem = (size_t)GRAM_CHUNK * (p.C + 1) * sizeof(float);
    if (smem > 48 * 1024) return PMG_ERR_UNSUPPORTED;
    gram_patch_kernel<<<B * npy * npx, 256, smem, st>>>(feats, grams, s.Hc, s.Wc, p.C, npy, npx, ph, pw, npx * pw);
    PMG_LAUNCH_OK("style gram_patch");
    return PMG_OK;
}

int generic_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x,
                    int x_dtype, int B, float* logits, float* gram, const pmg_taps* taps, char* ws, cudaStream_t st,
                    StageTimer* tm) {
    if (p.layer[0].k != 4) return PMG_ERR_UNSUPPORTED;   // patch_size <= 4: no strided layer (never configured)
    if (p.nl > 8) return PMG_ERR_UNSUPPORTED;
    if (p.nl > 0 && p.d / p.nh > ATT_MAX_HD) return PMG_ERR_UNSUPPORTED;
    const GenericWs w = generic_ws(p, s, B);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    auto P = [&](size_t off) { return reinterpret_cast<const float*>(packed + off); };
    int rc;
    PMG_MARK(tm, "start");
    // backbone (models_PMG.py:178)
    const float* cur = nullptr;
    if ((rc = run_backbone(p, s, L, packed, x, x_dtype, B, ws, w, st, tm, &cur, true))) return rc;
    // chan_proj (:182)
    {
        GemmArgs g{};
        g.A = cur; g.lda = p.C; g.Wt = P(L.chan_w); g.C = F(w.chan);
        g.M = B * s.Hc * s.Wc; g.N = p.Cr; g.K = p.C; g.ldc = p.Cr; g.act = ACT_NONE;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        PMG_MARK(tm, "chan_proj");
        if (taps && taps->chan)
            PMG_CUDA_OK(cudaMemcpyAsync(taps->chan, g.C, (size_t)g.M * p.Cr * sizeof(float), cudaMemcpyDeviceToDevice, st), "tap chan");
    }
    // patches + Gram (:185-188)
    float* G = gram ? gram : F(w.gram);
    {
        const size_t smem = (size_t)GRAM_CHUNK * (p.Cr + 1) * sizeof(float);
        if (smem > 48 * 1024) return PMG_ERR_UNSUPPORTED;
        gram_patch_kernel<<<B * s.np, 256, smem, st>>>(F(w.chan), G, s.Hc, s.Wc, p.Cr, p.pdiv, p.pdiv, s.ph, s.pw, 0);
        PMG_LAUNCH_OK("gram_patch");
        PMG_MARK(tm, "gram");
    }
    // token_proj (:192-193)
    const int T = B * s.np;
    {
        GemmArgs g{};
        g.A = G; g.lda = p.Cr * p.Cr; g.Wt = P(L.tok_w); g.shift = P(L.tok_b); g.C = F(w.tok);
        g.M = T; g.N = p.d; g.K = p.Cr * p.Cr; g.ldc = p.d; g.act = ACT_NONE;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        PMG_MARK(tm, "token_proj");
        if (taps && taps->tokens_in)
            PMG_CUDA_OK(cudaMemcpyAsync(taps->tokens_in, g.C, (size_t)T * p.d * sizeof(float), cudaMemcpyDeviceToDevice, st), "tap tok");
    }
    // transformer (:194) — pre-LN encoder layers, eval (dropout inert)
    for (int l = 0; l < p.nl; ++l) {
        const TfLayerOff& t = L.tf[l];
        const int d = p.d;
        layernorm_kernel<<<(T + 3) / 4, 128, 0, st>>>(F(w.tok), P(t.ln1w), P(t.ln1b), F(w.h), T, d, kLnEps);
        PMG_LAUNCH_OK("ln1");
        GemmArgs g{};
        g.A = F(w.h); g.lda = d; g.Wt = P(t.w_in); g.shift = P(t.b_in); g.C = F(w.qkv); g.M = T; g.N = 3 * d; g.K = d; g.ldc = 3 * d;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        {
            const size_t smem = (size_t)3 * s.np * (d / p.nh) * sizeof(float);
            if (smem > 48 * 1024) return PMG_ERR_UNSUPPORTED;
            attention_kernel<<<dim3(B, p.nh), 32, smem, st>>>(F(w.qkv), F(w.att), s.np, d, p.nh);
            PMG_LAUNCH_OK("attention");
        }
        g = GemmArgs{};
        g.A = F(w.att); g.lda = d; g.Wt = P(t.w_o); g.shift = P(t.b_o); g.resid = F(w.tok); g.C = F(w.tok); g.M = T; g.N = d; g.K = d; g.ldc = d;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        layernorm_kernel<<<(T + 3) / 4, 128, 0, st>>>(F(w.tok), P(t.ln2w), P(t.ln2b), F(w.h), T, d, kLnEps);
        PMG_LAUNCH_OK("ln2");
        g = GemmArgs{};
        g.A = F(w.h); g.lda = d; g.Wt = P(t.w_1); g.shift = P(t.b_1); g.C = F(w.ffh); g.M = T; g.N = 2 * d; g.K = d; g.ldc = 2 * d; g.act = ACT_RELU;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        g = GemmArgs{};
        g.A = F(w.ffh); g.lda = 2 * d; g.Wt = P(t.w_2); g.shift = P(t.b_2); g.resid = F(w.tok); g.C = F(w.tok); g.M = T; g.N = d; g.K = 2 * d; g.ldc = d;
        if ((rc = launch_sgemm(A_PLAIN, g, st))) return rc;
        PMG_MARK(tm, "encoder_layer");
    }
    if (taps && taps->tokens_out)
        PMG_CUDA_OK(cudaMemcpyAsync(taps->tokens_out, F(w.tok), (size_t)T * p.d * sizeof(float), cudaMemcpyDeviceToDevice, st), "tap tok out");
    // task pooling + heads (:197-198)
    rc = launch_pool_heads(p, L, packed, F(w.tok), B, s.np, logits, st);
    PMG_MARK(tm, "pool_heads");
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Generic batched Gram for the microbenchmark shapes:  G[b] = scale · F[b]·F[b]ᵀ, F (B,C,HW) row-major.
// (CUDA-core fp32 version; the tcgen05 version lives in pmg_fast.cu and is preferred when it applies.)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gram_cm_kernel(const T* __restrict__ Fm, float* __restrict__ G, int C, int HW, float scale) {
    // block computes a 64×64 tile of G[b]; K = HW streamed in chunks of 16
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int b = blockIdx.z, i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    const T* Fb = Fm + (size_t)b * C * HW;
    const int tid = threadIdx.x, lr = tid >> 2, lk = (tid & 3) * 4, ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < HW; k0 += 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + lk + j;
            As[lk + j][lr] = (i0 + lr < C && k < HW) ? (float)Fb[(size_t)(i0 + lr) * HW + k] : 0.f;
            Bs[lk + j][lr] = (j0 + lr < C && k < HW) ? (float)Fb[(size_t)(j0 + lr) * HW + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 c = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gi = i0 + ty * 4 + i, gj = j0 + tx * 4 + j;
            if (gi < C && gj < C) G[((size_t)b * C + gi) * C + gj] = acc[i][j] * scale;
        }
}

int gram_batched_generic(const void* Fm, int f_dtype, int B, int C, int HW, float scale, float* G, cudaStream_t st) {
    dim3 grid((C + 63) / 64, (C + 63) / 64, B);
    if (f_dtype == PMG_BF16)
        gram_cm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(Fm), G, C, HW, scale);
    else
        gram_cm_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(Fm), G, C, HW, scale);
    PMG_LAUNCH_OK("gram_cm");
    return PMG_OK;
}

}  // namespace pmg
