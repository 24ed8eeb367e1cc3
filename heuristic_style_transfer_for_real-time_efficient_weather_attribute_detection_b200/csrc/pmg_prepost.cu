// pmg_prepost.cu — the callers' pre- and post-processing either side of the forward, on the GPU.
//
// Pre  : transforms.ToTensor() + transforms.Normalize(mean, std) of utils/datasets_utils.py:212-214 on decoded
//        uint8 images (HWC as PIL / OpenCV deliver them, or CHW), written as the NCHW bf16 / fp32 batch the
//        forward consumes.  Shipping uint8 over PCIe instead of the normalised float tensor cuts the host→device
//        bytes per image from 602 KB (fp32) / 301 KB (bf16) to 150 KB.
// Post : per-task softmax, max probability, arg-max and the prob_threshold → -1 rule of
//        Functions/PMG_evaluation.py:144-146 (and utils/benchmark_utils.py:73-75) on the packed logits, so the
//        caller does one device→host copy instead of 12 per batch.
// Both are pure HBM-streaming kernels (no tensor-core work).
#include "pmg_common.cuh"

namespace pmg {

struct PreArgs {
    const unsigned char* src;
    void* dst;
    int B, H, W;
    float mean[3], stdv[3];
};

__device__ __forceinline__ float norm_px(unsigned char u, float mean, float stdv) {
    // ToTensor: float(u) / 255 (IEEE division, as torch's .div(255)); Normalize: (x - mean) / std
    return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), mean), stdv);
}

// one thread = 4 consecutive pixels of one image row, all 3 channels
template <bool HWC, bool OUT_BF16>
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const PreArgs a) {
    const int wq = (a.W + 3) >> 2;
    const size_t total = (size_t)a.B * a.H * wq;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int xq = (int)(t % wq);
        const size_t by = t / wq;
        const int y = (int)(by % a.H);
        const size_t b = by / a.H;
        const int x0 = xq * 4, n = min(4, a.W - x0);
        unsigned char px[3][4];
        if (HWC) {
            const unsigned char* s = a.src + ((b * a.H + y) * a.W + x0) * 3;
            if (n == 4 && ((reinterpret_cast<uintptr_t>(s) & 3) == 0)) {
                const uint32_t w0 = reinterpret_cast<const uint32_t*>(s)[0], w1 = reinterpret_cast<const uint32_t*>(s)[1],
                               w2 = reinterpret_cast<const uint32_t*>(s)[2];
                const unsigned char raw[12] = {
                    (unsigned char)w0, (unsigned char)(w0 >> 8), (unsigned char)(w0 >> 16), (unsigned char)(w0 >> 24),
                    (unsigned char)w1, (unsigned char)(w1 >> 8), (unsigned char)(w1 >> 16), (unsigned char)(w1 >> 24),
                    (unsigned char)w2, (unsigned char)(w2 >> 8), (unsigned char)(w2 >> 16), (unsigned char)(w2 >> 24)};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) px[c][i] = raw[3 * i + c];
            } else {
                for (int i = 0; i < 4; ++i)
                    for (int c = 0; c < 3; ++c) px[c][i] = i < n ? s[3 * i + c] : 0;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const unsigned char* s = a.src + ((b * 3 + c) * a.H + y) * a.W + x0;
                if (n == 4 && ((reinterpret_cast<uintptr_t>(s) & 3) == 0)) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(s);
                    px[c][0] = (unsigned char)w; px[c][1] = (unsigned char)(w >> 8); px[c][2] = (unsigned char)(w >> 16); px[c][3] = (unsigned char)(w >> 24);
                } else {
                    for (int i = 0; i < 4; ++i) px[c][i] = i < n ? s[i] : 0;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = norm_px(px[c][i], a.mean[c], a.stdv[c]);
            const size_t o = ((b * 3 + c) * a.H + y) * a.W + x0;
            if (OUT_BF16) {
                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(a.dst) + o;
                if (n == 4 && ((reinterpret_cast<uintptr_t>(d) & 7) == 0)) {
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                    *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
                } else {
                    for (int i = 0; i < n; ++i) d[i] = __float2bfloat16_rn(v[i]);
                }
            } else {
                float* d = reinterpret_cast<float*>(a.dst) + o;
                if (n == 4 && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
                else for (int i = 0; i < n; ++i) d[i] = v[i];
            }
        }
    }
}

struct PostArgs {
    const float* logits;
    float* probs;        // (B, L) or null
    int* preds;          // (B, ntask) or null
    float* maxp;         // (B, ntask) or null
    int B, ntask, L;
    float thr;
    int off[PMG_MAX_TASKS + 1];
};

// one thread = one (image, task): ncls ≤ a few tens
__global__ void __launch_bounds__(256) postprocess_kernel(const PostArgs a) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.B * a.ntask) return;
    const int b = idx / a.ntask, t = idx - b * a.ntask;
    const int o0 = a.off[t], n = a.off[t + 1] - o0;
    const float* lg = a.logits + (size_t)b * a.L + o0;
    float mx = -INFINITY;
    for (int c = 0; c < n; ++c) mx = fmaxf(mx, lg[c]);
    float den = 0.f;
    for (int c = 0; c < n; ++c) den += expf(lg[c] - mx);
    float best = -1.f;
    int arg = 0;
    for (int c = 0; c < n; ++c) {
        const float p = expf(lg[c] - mx) / den;
        if (a.probs) a.probs[(size_t)b * a.L + o0 + c] = p;
        if (p > best) { best = p; arg = c; }      // first maximum, as torch.max
    }
    if (a.maxp) a.maxp[idx] = best;
    if (a.preds) a.preds[idx] = best < a.thr ? -1 : arg;
}

}  // namespace pmg

using namespace pmg;

extern "C" int pmg_preprocess_u8(const void* src, int layout, int B, int H, int W, const float* mean3, const float* std3,
                                 int out_dtype, void* dst, pmg_stream_t stream) {
    if (!src || !dst || !mean3 || !std3) return PMG_ERR_BAD_ARG;
    if (layout != PMG_LAYOUT_NCHW && layout != PMG_LAYOUT_NHWC) return PMG_ERR_BAD_ARG;
    if (out_dtype != PMG_F32 && out_dtype != PMG_BF16) return PMG_ERR_BAD_ARG;
    if (B < 1 || H < 1 || W < 1) return PMG_ERR_BAD_SHAPE;
    PreArgs a{};
    a.src = reinterpret_cast<const unsigned char*>(src); a.dst = dst; a.B = B; a.H = H; a.W = W;
    for (int c = 0; c < 3; ++c) {
        if (!(std3[c] > 0.f)) return PMG_ERR_BAD_ARG;
        a.mean[c] = mean3[c]; a.stdv[c] = std3[c];
    }
    const size_t total = (size_t)B * H * ((W + 3) / 4);
    const size_t want = (total + 255) / 256;
    const int grid = (int)(want > 148 * 32 ? 148 * 32 : want);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (layout == PMG_LAYOUT_NHWC) {
        if (out_dtype == PMG_BF16) preprocess_u8_kernel<true, true><<<grid, 256, 0, st>>>(a);
        else preprocess_u8_kernel<true, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (out_dtype == PMG_BF16) preprocess_u8_kernel<false, true><<<grid, 256, 0, st>>>(a);
        else preprocess_u8_kernel<false, false><<<grid, 256, 0, st>>>(a);
    }
    PMG_LAUNCH_OK("preprocess_u8");
    return PMG_OK;
}

extern "C" int pmg_postprocess(const float* logits, int B, int n_tasks, const int32_t* ncls, float prob_threshold,
                               float* probs, int32_t* preds, float* max_probs, pmg_stream_t stream) {
    if (!logits || !ncls || (!probs && !preds && !max_probs)) return PMG_ERR_BAD_ARG;
    if (B < 1 || n_tasks < 1 || n_tasks > PMG_MAX_TASKS) return PMG_ERR_BAD_SHAPE;
    PostArgs a{};
    a.logits = logits; a.probs = probs; a.preds = preds; a.maxp = max_probs; a.B = B; a.ntask = n_tasks; a.thr = prob_threshold;
    int o = 0;
    for (int t = 0; t < n_tasks; ++t) {
        if (ncls[t] < 1) return PMG_ERR_BAD_ARG;
        a.off[t] = o; o += ncls[t];
    }
    a.off[n_tasks] = o; a.L = o;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    postprocess_kernel<<<(B * n_tasks + 255) / 256, 256, 0, st>>>(a);
    PMG_LAUNCH_OK("postprocess");
    return PMG_OK;
}

// ------------------------------------------------------------------------------------------------
// transforms.Resize(size) + transforms.CenterCrop(crop) (utils/datasets_utils.py:209-210) on decoded uint8 HWC images,
// bit-identical to what the reference's transform computes on the CPU:
//   * output size and crop offsets as torchvision (_compute_resized_output_size: the shorter edge becomes `size`, the
//     longer int(size·long/short); center_crop: int(round((H' − crop) / 2.0)));
//   * pixel values as Pillow's Image.resize(…, BILINEAR) (src/libImaging/Resample.c): tap tables from
//     precompute_coeffs (support = max(scale, 1), taps normalised in double) turned into 22-bit fixed point by
//     normalize_coeffs_8bpc; a horizontal 8-bit pass, THEN a vertical 8-bit pass, each Σ pixel·tap from 1 << 21, >> 22,
//     clamped to uint8.  A pass whose axis keeps its size is skipped, as Pillow does.
// The tap tables (only for the crop window) are computed on the host with the same double arithmetic and uploaded into
// the caller's scratch; one thread = one output pixel, the horizontal pass recomputed for each of the vertical taps
// (≤ (2·ceil(scale)+1)² byte loads per channel — integer work, HBM / L2 bound).  Optionally fused with ToTensor +
// Normalize (norm_px above) into the NCHW bf16 / fp32 batch the forward consumes.
// ------------------------------------------------------------------------------------------------
namespace pmg {

constexpr int kResizeBits = 32 - 8 - 2;

struct ResizeArgs {
    const unsigned char* src;
    void* dst;
    const int* bh; const int* kh;     // [crop][2] first input index / tap count, [crop][ksh] taps (horizontal)
    const int* bv; const int* kv;
    int B, H, W, crop, ksh, ksv, skip_h, skip_v;
    float mean[3], stdv[3];
};

__device__ __forceinline__ int clip8i(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// OUT: 0 = uint8 HWC, 1 = normalised bf16 NCHW, 2 = normalised fp32 NCHW
template <int OUT>
__global__ void __launch_bounds__(256) resize_crop_kernel(const ResizeArgs a) {
    const size_t total = (size_t)a.B * a.crop * a.crop;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % a.crop), y = (int)((t / a.crop) % a.crop);
        const size_t b = t / ((size_t)a.crop * a.crop);
        // an axis that keeps its size is not resampled (Pillow skips the pass): its source index is the tap window's first
        const int x0 = a.bh[2 * x], nx = a.bh[2 * x + 1], y0 = a.bv[2 * y], ny = a.skip_v ? 1 : a.bv[2 * y + 1];
        const int* kx = a.kh + (size_t)x * a.ksh;
        const int* ky = a.kv + (size_t)y * a.ksv;
        const unsigned char* img = a.src + b * (size_t)a.H * a.W * 3;
        int out[3];
        int av[3] = {1 << (kResizeBits - 1), 1 << (kResizeBits - 1), 1 << (kResizeBits - 1)};
        for (int j = 0; j < ny; ++j) {
            const unsigned char* row = img + ((size_t)(y0 + j) * a.W + x0) * 3;
            int hv[3];
            if (a.skip_h) {
                hv[0] = row[0]; hv[1] = row[1]; hv[2] = row[2];
            } else {
                int ah[3] = {1 << (kResizeBits - 1), 1 << (kResizeBits - 1), 1 << (kResizeBits - 1)};
                for (int i = 0; i < nx; ++i) {
                    const int k = kx[i];
                    ah[0] += row[3 * i] * k; ah[1] += row[3 * i + 1] * k; ah[2] += row[3 * i + 2] * k;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) hv[c] = clip8i(ah[c] >> kResizeBits);
            }
            if (a.skip_v) {
#pragma unroll
                for (int c = 0; c < 3; ++c) out[c] = hv[c];
            } else {
                const int k = ky[j];
#pragma unroll
                for (int c = 0; c < 3; ++c) av[c] += hv[c] * k;
            }
        }
        if (!a.skip_v) {
#pragma unroll
            for (int c = 0; c < 3; ++c) out[c] = clip8i(av[c] >> kResizeBits);
        }
        if (OUT == 0) {
            unsigned char* d = reinterpret_cast<unsigned char*>(a.dst) + t * 3;
            d[0] = (unsigned char)out[0]; d[1] = (unsigned char)out[1]; d[2] = (unsigned char)out[2];
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = norm_px((unsigned char)out[c], a.mean[c], a.stdv[c]);
                const size_t o = ((b * 3 + c) * a.crop + y) * a.crop + x;
                if (OUT == 1) reinterpret_cast<__nv_bfloat16*>(a.dst)[o] = __float2bfloat16_rn(v);
                else reinterpret_cast<float*>(a.dst)[o] = v;
            }
        }
    }
}

// torchvision _compute_resized_output_size for an int size
static void resized_hw(int H, int W, int size, int* nh, int* nw) {
    const int sh = W <= H ? W : H, lg = W <= H ? H : W;
    const int nl = (int)((double)size * lg / sh);          // Python: int(size * long / short) — true division, truncation
    if (W <= H) { *nw = size; *nh = nl; } else { *nh = size; *nw = nl; }
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc (BILINEAR) for output indices [o0, o0 + n) of an axis in_size → out_size
static int tap_count(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(scale) * 2 + 1;
}
static void make_taps(int in_size, int out_size, int o0, int n, int ksize, int* bounds, int* kk) {
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale, ss = 1.0 / filterscale;
    double* w = new double[ksize];
    for (int q = 0; q < n; ++q) {
        const int xx = o0 + q;
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double v = (x + xmin - center + 0.5) * ss;
            if (v < 0.0) v = -v;
            w[x] = v < 1.0 ? 1.0 - v : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) w[x] /= ww;
        for (int x = xmax; x < ksize; ++x) w[x] = 0.0;
        for (int x = 0; x < ksize; ++x)
            kk[(size_t)q * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << kResizeBits)) : (int)(0.5 + w[x] * (1 << kResizeBits));
        bounds[2 * q] = xmin; bounds[2 * q + 1] = xmax;
    }
    delete[] w;
}

}  // namespace pmg

extern "C" size_t pmg_resize_scratch_bytes(int H, int W, int size, int crop) {
    if (H < 1 || W < 1 || size < 1 || crop < 1) return 0;
    int nh, nw;
    resized_hw(H, W, size, &nh, &nw);
    if (nh < crop || nw < crop) return 0;
    return (size_t)crop * (4 + tap_count(W, nw) + tap_count(H, nh)) * sizeof(int) + 256;
}

extern "C" int pmg_resize_crop_u8(const void* src, int B, int H, int W, int size, int crop, const float* mean3, const float* std3,
                                  int out_mode, void* dst, void* scratch, size_t scratch_bytes, pmg_stream_t stream) {
    if (!src || !dst || !scratch) return PMG_ERR_BAD_ARG;
    if (out_mode < 0 || out_mode > 2 || (out_mode > 0 && (!mean3 || !std3))) return PMG_ERR_BAD_ARG;
    if (B < 1 || H < 1 || W < 1 || size < 1 || crop < 1 || H > 32768 || W > 32768) return PMG_ERR_BAD_SHAPE;
    int nh, nw;
    resized_hw(H, W, size, &nh, &nw);
    if (nh < crop || nw < crop) return PMG_ERR_UNSUPPORTED;       // CenterCrop would pad (crop > resized image): not implemented
    const size_t need = pmg_resize_scratch_bytes(H, W, size, crop);
    if (scratch_bytes < need || (reinterpret_cast<uintptr_t>(scratch) & 3)) return PMG_ERR_WORKSPACE;
    const int top = (int)nearbyint((nh - crop) / 2.0), left = (int)nearbyint((nw - crop) / 2.0);   // Python round(): half to even
    ResizeArgs a{};
    a.ksh = tap_count(W, nw); a.ksv = tap_count(H, nh);
    const size_t n_int = (size_t)crop * (4 + a.ksh + a.ksv);
    int* host = new int[n_int];
    int* hbh = host, *hbv = host + 2 * crop, *hkh = host + 4 * crop, *hkv = hkh + (size_t)crop * a.ksh;
    make_taps(W, nw, left, crop, a.ksh, hbh, hkh);
    make_taps(H, nh, top, crop, a.ksv, hbv, hkv);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // pageable source: the runtime stages the bytes before cudaMemcpyAsync returns, so `host` may be freed right after
    cudaError_t e = cudaMemcpyAsync(scratch, host, n_int * sizeof(int), cudaMemcpyHostToDevice, st);
    delete[] host;
    PMG_CUDA_OK(e, "resize tap upload");
    int* dev = reinterpret_cast<int*>(scratch);
    a.bh = dev; a.bv = dev + 2 * crop; a.kh = dev + 4 * crop; a.kv = a.kh + (size_t)crop * a.ksh;
    a.src = reinterpret_cast<const unsigned char*>(src); a.dst = dst;
    a.B = B; a.H = H; a.W = W; a.crop = crop; a.skip_h = nw == W; a.skip_v = nh == H;
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean3 ? mean3[c] : 0.f; a.stdv[c] = std3 ? std3[c] : 1.f; }
    if (out_mode > 0)
        for (int c = 0; c < 3; ++c)
            if (!(a.stdv[c] > 0.f)) return PMG_ERR_BAD_ARG;
    const size_t total = (size_t)B * crop * crop;
    const int grid = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    if (out_mode == 0) resize_crop_kernel<0><<<grid, 256, 0, st>>>(a);
    else if (out_mode == 1) resize_crop_kernel<1><<<grid, 256, 0, st>>>(a);
    else resize_crop_kernel<2><<<grid, 256, 0, st>>>(a);
    PMG_LAUNCH_OK("resize_crop");
    return PMG_OK;
}
