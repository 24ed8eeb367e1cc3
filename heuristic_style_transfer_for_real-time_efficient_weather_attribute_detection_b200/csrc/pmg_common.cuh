// pmg_common.cuh — shared host-side plan / packed-weight layout for the PMG forward path.
// Reference semantics: /root/reference/Models/models_PMG.py:78-199 (cited per item below).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/pmg_b200.h"

namespace pmg {

constexpr float kBnEps = 1e-5f;   // nn.BatchNorm2d default (models_PMG.py:105)
constexpr float kLnEps = 1e-5f;   // TransformerEncoderLayer layer_norm_eps default (models_PMG.py:64)
constexpr float kLeaky = 0.2f;    // models_PMG.py:117,127

struct ConvLayer {
    int cin, cout, k;             // k = 4 (stride 4, pad 0) or 1
};

// Backbone plan = the ctor's layer-count loop (models_PMG.py:107-128).
struct Plan {
    int n_layers = 0;
    ConvLayer layer[PMG_MAX_BACKBONE];
    int C = 0;                    // channels after the final 1x1 (self.C)
    int Cr = 0, d = 0, nl = 0, nh = 0, ntask = 0, ncls_total = 0, pdiv = 0;
    int norm = PMG_NORM_BATCH;    // models_PMG.py:100-105
    int se = 0, se_hidden = 0;    // SEBlock (models_PMG.py:15-30): C -> max(1, C/16) -> C
    int cls_off[PMG_MAX_TASKS + 1];
    bool ok = false;
};

inline Plan make_plan(const pmg_config* c) {
    Plan p;
    if (!c || c->input_nc < 1 || c->ndf < 1 || c->patch_div < 1 || c->gram_channels < 1 || c->d_model < 1 ||
        c->transformer_layers < 0 || c->n_tasks < 1 || c->n_tasks > PMG_MAX_TASKS)
        return p;
    if (c->transformer_layers > 0 && (c->transformer_heads < 1 || c->d_model % c->transformer_heads)) return p;
    int cin = c->input_nc, nf = c->ndf;
    float rf = (float)c->patch_size;
    while (rf > 4.f && nf <= 512) {
        if (p.n_layers >= PMG_MAX_BACKBONE - 1) return p;
        p.layer[p.n_layers++] = {cin, nf, 4};
        cin = nf; nf *= 2; rf /= 4.f;
    }
    p.layer[p.n_layers++] = {cin, nf, 1};
    p.C = nf;
    if (c->norm != PMG_NORM_BATCH && c->norm != PMG_NORM_INSTANCE) return p;
    p.norm = c->norm;
    p.se = c->use_channel_attention ? 1 : 0;
    p.se_hidden = p.se ? (p.C / 16 > 1 ? p.C / 16 : 1) : 0;
    p.Cr = c->gram_channels; p.d = c->d_model; p.nl = c->transformer_layers; p.nh = c->transformer_heads;
    p.ntask = c->n_tasks; p.pdiv = c->patch_div;
    int off = 0;
    for (int t = 0; t < c->n_tasks; ++t) {
        if (c->ncls[t] < 1) return p;
        p.cls_off[t] = off; off += c->ncls[t];
    }
    p.cls_off[c->n_tasks] = off;
    p.ncls_total = off;
    p.ok = true;
    return p;
}

// Spatial sizes through the backbone + patch split (models_PMG.py:157-174).
struct Shape {
    int h[PMG_MAX_BACKBONE + 1], w[PMG_MAX_BACKBONE + 1];   // h[0],w[0] = input
    int Hc, Wc;        // chan_proj map
    int ph, pw;        // patch height / width after replicate padding
    int np, n;         // patches per image, pixels per patch
    bool ok = false;
};

inline Shape make_shape(const Plan& p, int H, int W) {
    Shape s;
    s.h[0] = H; s.w[0] = W;
    for (int l = 0; l < p.n_layers; ++l) {
        if (p.layer[l].k == 4) { s.h[l + 1] = s.h[l] / 4; s.w[l + 1] = s.w[l] / 4; }
        else { s.h[l + 1] = s.h[l]; s.w[l + 1] = s.w[l]; }
        if (s.h[l + 1] < 1 || s.w[l + 1] < 1) return s;
    }
    s.Hc = s.h[p.n_layers]; s.Wc = s.w[p.n_layers];
    const int Hp = s.Hc + ((p.pdiv - s.Hc % p.pdiv) % p.pdiv);
    const int Wp = s.Wc + ((p.pdiv - s.Wc % p.pdiv) % p.pdiv);
    s.ph = Hp / p.pdiv; s.pw = Wp / p.pdiv;
    if (s.ph < 1 || s.pw < 1) return s;
    s.np = p.pdiv * p.pdiv; s.n = s.ph * s.pw;
    s.ok = true;
    return s;
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// ---------------------------------------------------------------------------------------------
// Packed weights, generic (fp32) section.  All offsets in bytes from the start of the buffer.
// ---------------------------------------------------------------------------------------------
struct TfLayerOff {
    size_t w_in, b_in, w_o, b_o, w_1, b_1, w_2, b_2, ln1w, ln1b, ln2w, ln2b;
};

struct PackedLayout {
    size_t conv_w[PMG_MAX_BACKBONE];   // Wt [K][Cout] fp32; K order (c,ky,kx) for layer 0, (ky,kx,c) after
    size_t conv_scale[PMG_MAX_BACKBONE], conv_shift[PMG_MAX_BACKBONE];
    size_t se_w1, se_w2;               // SEBlock fc.0.weight [hid][C], fc.2.weight [C][hid] (torch layouts)
    size_t chan_w;                     // Wt [C][Cr]
    size_t tok_w, tok_b;               // Wt [Cr*Cr][d], b [d]
    TfLayerOff tf[8];
    size_t q;                          // [ntask][d]  (Tasks.json order)
    size_t cls_w, cls_b;               // [ncls_total][d], [ncls_total]
    size_t fast;                       // start of the tcgen05 fast-path section (0 bytes if not applicable)
    size_t total;
};

size_t fast_section_bytes(const Plan& p);   // defined in pmg_fast.cu

inline PackedLayout make_layout(const Plan& p) {
    PackedLayout L{};
    size_t o = 256;  // header
    auto take = [&](size_t n_floats) { size_t r = o; o = align256(o + n_floats * sizeof(float)); return r; };
    for (int l = 0; l < p.n_layers; ++l) {
        const int K = p.layer[l].cin * p.layer[l].k * p.layer[l].k;
        L.conv_w[l] = take((size_t)K * p.layer[l].cout);
        L.conv_scale[l] = take(p.layer[l].cout);
        L.conv_shift[l] = take(p.layer[l].cout);
    }
    if (p.se) { L.se_w1 = take((size_t)p.se_hidden * p.C); L.se_w2 = take((size_t)p.C * p.se_hidden); }
    L.chan_w = take((size_t)p.C * p.Cr);
    L.tok_w = take((size_t)p.Cr * p.Cr * p.d);
    L.tok_b = take(p.d);
    for (int l = 0; l < p.nl && l < 8; ++l) {
        TfLayerOff& t = L.tf[l];
        t.w_in = take((size_t)p.d * 3 * p.d); t.b_in = take(3 * p.d);
        t.w_o = take((size_t)p.d * p.d); t.b_o = take(p.d);
        t.w_1 = take((size_t)p.d * 2 * p.d); t.b_1 = take(2 * p.d);
        t.w_2 = take((size_t)2 * p.d * p.d); t.b_2 = take(p.d);
        t.ln1w = take(p.d); t.ln1b = take(p.d); t.ln2w = take(p.d); t.ln2b = take(p.d);
    }
    L.q = take((size_t)p.ntask * p.d);
    L.cls_w = take((size_t)p.ncls_total * p.d);
    L.cls_b = take(p.ncls_total);
    L.fast = o;
    o = align256(o + fast_section_bytes(p));
    L.total = o;
    return L;
}

// thread-local last CUDA error text
void set_cuda_error(cudaError_t e, const char* where);

#define PMG_CUDA_OK(expr, where)                                        \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) { ::pmg::set_cuda_error(_e, where); return PMG_ERR_CUDA; } \
    } while (0)

#define PMG_LAUNCH_OK(where) PMG_CUDA_OK(cudaPeekAtLastError(), where)

// Optional per-stage device timing (pmg_forward_profile): one event after every launch.
struct StageTimer {
    static constexpr int kMax = 64;
    cudaEvent_t ev[kMax];
    const char* name[kMax];
    int n = 0;
    cudaStream_t st = nullptr;
    void mark(const char* nm) {
        if (n < kMax && cudaEventCreate(&ev[n]) == cudaSuccess) { cudaEventRecord(ev[n], st); name[n] = nm; ++n; }
    }
};
#define PMG_MARK(tm, nm) do { if (tm) (tm)->mark(nm); } while (0)

// generic path entry points (pmg_generic.cu)
size_t generic_workspace_bytes(const Plan& p, const Shape& s, int B);
int generic_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x,
                    int x_dtype, int B, float* logits, float* gram, const pmg_taps* taps, char* ws,
                    cudaStream_t st, StageTimer* tm);
int generic_launch_count(const Plan& p);
int style_grams_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x, int x_dtype,
                        int B, float* grams, char* ws, cudaStream_t st);
int pack_generic(const Plan& p, const PackedLayout& L, const float* const* prm, char* packed, cudaStream_t st);

// Extra destinations of the packed logits (pmg_forward_bcast): device pointers, possibly peer-mapped memory of other GPUs.
// The fused token-stage kernel stores every logit to all of them from its epilogue (the multi-GPU exchange costs no extra
// launch, copy or SM); every other path copies after the fact.
constexpr int PMG_MAX_PEERS = 15;
struct LogitPeers {
    float* dst[PMG_MAX_PEERS];
    int n = 0;
};
int copy_logits_to_peers(const float* logits, size_t bytes, const LogitPeers* peers, cudaStream_t st);

// fast path entry points (pmg_fast.cu)
bool fast_supported(const Plan& p, const Shape& s);
size_t fast_workspace_bytes(const Plan& p, const Shape& s, int B);
int fast_forward(const Plan& p, const Shape& s, const PackedLayout& L, const char* packed, const void* x,
                 int x_dtype, int B, float* logits, float* gram, const pmg_taps* taps, char* ws,
                 cudaStream_t st, StageTimer* tm, unsigned flags, const LogitPeers* peers = nullptr);
int fast_launch_count(const Plan& p, const Shape& s, int B, unsigned flags);
int pack_fast(const Plan& p, const PackedLayout& L, const float* const* prm, char* packed, cudaStream_t st);

}  // namespace pmg
