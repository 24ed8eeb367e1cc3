/*
 * pmg_b200.h — C ABI of the B200-native PMG (PatchGAN-MultiTasks-Gram) forward path.
 *
 * The reference (Hamedkiri/Heuristic_Style_Transfer_for_Real-Time_Efficient_Weather_Attribute_Detection)
 * has no FFI / plugin layer: its boundary for this path is the Python nn.Module
 *   Models/models_PMG.py:77-199  class MultiTaskPatchGANGramModelNonOverlapV2
 * whose forward() is pure eager PyTorch.  This header is the boundary that sits INSIDE that
 * forward in the replacement module (package file models_PMG.py): plain pointers and sizes,
 * no torch types, no allocation, no global state (re-entrant per {stream, workspace}).
 * Every function returns a pmg_status (0 = ok) and never throws across the ABI.
 *
 * All pointers are DEVICE pointers unless stated otherwise.  The library is built for sm_100a
 * only; on any other device pmg_forward returns PMG_ERR_ARCH (there is no CPU fallback).
 */
#ifndef PMG_B200_H
#define PMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMG_ABI_VERSION 2
#define PMG_MAX_TASKS 32
#define PMG_MAX_BACKBONE 8

typedef struct CUstream_st* pmg_stream_t; /* == cudaStream_t */

typedef enum pmg_status {
    PMG_OK = 0,
    PMG_ERR_BAD_ARG = 1,      /* null pointer, bad dtype code, inconsistent config            */
    PMG_ERR_BAD_SHAPE = 2,    /* B<1, H or W too small for the backbone / patch split          */
    PMG_ERR_WORKSPACE = 3,    /* workspace / packed buffer smaller than *_bytes() reports      */
    PMG_ERR_ARCH = 4,         /* current device is not compute capability 10.x                 */
    PMG_ERR_CUDA = 5,         /* a CUDA call or kernel launch failed (see pmg_last_cuda_error) */
    PMG_ERR_UNSUPPORTED = 6   /* configuration the kernels do not implement                    */
} pmg_status;

typedef enum pmg_dtype { PMG_F32 = 0, PMG_BF16 = 1 } pmg_dtype;

/* Constructor hyper-parameters — mirrors Models/models_PMG.py:78-92 and the "hparams" block of
 * Model_weights/PMG/hyperparameters_PMG.json:8-33.  The committed configuration (norm "batch", no channel attention) at
 * 224x224 takes the tcgen05 fast path; norm "instance" (models_PMG.py:100-101) and use_channel_attention (SEBlock,
 * :15-30,133) run on the shape-generic fp32 kernels.  norm "group" cannot be constructed in the reference either
 * (:102-103 binds num_groups by keyword and then passes the channel count positionally: TypeError) and is rejected. */
#define PMG_NORM_BATCH 0     /* BatchNorm2d(affine), eval: running statistics folded into the conv epilogue */
#define PMG_NORM_INSTANCE 1  /* InstanceNorm2d(affine=True): per-sample, per-channel statistics */
typedef struct pmg_config {
    int32_t input_nc;            /* 3 */
    int32_t ndf;                 /* 32 */
    int32_t patch_size;          /* 64: drives the layer-count loop only (models_PMG.py:113-121) */
    int32_t patch_div;           /* 3 */
    int32_t gram_channels;       /* 96 */
    int32_t d_model;             /* 192 */
    int32_t transformer_layers;  /* 2 (0 = Identity, models_PMG.py:147) */
    int32_t transformer_heads;   /* 4 */
    int32_t n_tasks;             /* 12 */
    int32_t ncls[PMG_MAX_TASKS]; /* classes per task, Tasks.json order */
    int32_t norm;                /* PMG_NORM_BATCH or PMG_NORM_INSTANCE */
    int32_t use_channel_attention; /* 1 = SEBlock between feature_extractor and chan_proj (models_PMG.py:133,179) */
} pmg_config;

/* pmg_forward flags */
#define PMG_FLAG_FORCE_GENERIC 1u  /* skip the tcgen05 fast path even when the config/shape allow it */
#define PMG_FLAG_FORCE_FAST 2u     /* fail with PMG_ERR_UNSUPPORTED instead of using the generic kernels */
#define PMG_FLAG_NO_FUSED_FRONT 4u /* fast path: run conv0/conv1 as im2col + two GEMMs instead of the fused TMA-fed kernel (fp32 input: the fused kernel reads it as bf16 hi + lo planes) */
#define PMG_FLAG_NO_FUSED_MID 16u /* fast path: run conv2 and chan_proj as two GEMMs instead of the fused kernel */
#define PMG_FLAG_NO_FUSED_MIDGRAM 32u /* fast path: run conv2 + chan_proj and the Gram as two kernels (chan map through HBM) instead of the fused kernel */
#define PMG_FLAG_ALWAYS_FUSED_MIDGRAM 64u /* fast path: use the fused conv2+chan_proj+Gram kernel at every batch size (default: from 96 images up) */
#define PMG_FLAG_COUNT_F32_INPUT 128u /* pmg_forward_launch_count only: count for fp32 input (one more launch: the hi/lo plane split) */
#define PMG_FLAG_NO_FUSED_ENCODER 8u /* fast path: run the transformer / pooling / heads as separate kernels instead of the fused token-stage kernel */

/* Optional taps for parity tests: any non-null member receives that intermediate in fp32.
 *   chan       (B, Hc, Wc, gram_channels)  chan_proj output, pixel-major  (models_PMG.py:182)
 *   tokens_in  (B, Np, d_model)            token_proj output               (models_PMG.py:193)
 *   tokens_out (B, Np, d_model)            transformer output              (models_PMG.py:194) */
typedef struct pmg_taps {
    float* chan;
    float* tokens_in;
    float* tokens_out;
} pmg_taps;

int pmg_abi_version(void);
const char* pmg_status_string(int status);
/* last CUDA error string seen by the calling thread inside this library ("" if none) */
const char* pmg_last_cuda_error(void);

/* Number of parameter tensors pmg_pack_weights expects, in this order (all fp32, contiguous, torch
 * layouts; = the reference state_dict order of hyperparameters_PMG.json:53-135 minus the int64
 * num_batches_tracked entries, task queries in Tasks.json order):
 *   per backbone conv l:  conv.weight (Cout,Cin,k,k), norm.weight, norm.bias [, bn.running_mean, bn.running_var — norm "batch" only]
 *   [se.fc.0.weight (max(1,C/16), C), se.fc.2.weight (C, max(1,C/16)) — use_channel_attention only]
 *   chan_proj.weight (Cr,C,1,1);  token_proj.weight (d,Cr*Cr), token_proj.bias (d)
 *   per encoder layer:    in_proj_weight (3d,d), in_proj_bias, out_proj.weight (d,d), out_proj.bias,
 *                         linear1.weight (2d,d), linear1.bias, linear2.weight (d,2d), linear2.bias,
 *                         norm1.weight, norm1.bias, norm2.weight, norm2.bias
 *   per task: task_pool.query (d);   per task: classifier.weight (ncls,d), classifier.bias (ncls) */
int pmg_param_count(const pmg_config* cfg);

/* Bytes of the packed-weights buffer for this configuration (256-byte aligned sections). */
size_t pmg_packed_bytes(const pmg_config* cfg);

/* Fold BatchNorm(eval) into per-channel scale/shift, transpose / re-order / convert the weights
 * into the kernels' layouts.  `params` is a HOST array of n_params DEVICE pointers.
 * Replaces nothing in the reference (it keeps torch Parameters); must be re-run after every
 * load_state_dict / .to(device) (utils/PMG/PMG_model_utils.py:185-199). */
int pmg_pack_weights(const pmg_config* cfg, const float* const* params, int n_params,
                     void* packed, size_t packed_bytes, pmg_stream_t stream);

/* Bytes of scratch pmg_forward needs for a (B,input_nc,H,W) batch. */
size_t pmg_workspace_bytes(const pmg_config* cfg, int B, int H, int W);

/* Total number of logits per image = sum(ncls). */
int pmg_logits_per_image(const pmg_config* cfg);

/* Number of kernels one pmg_forward call launches for this shape/flags (for bench accounting). */
int pmg_forward_launch_count(const pmg_config* cfg, int B, int H, int W, unsigned flags);

/* The forward pass — replaces the body of MultiTaskPatchGANGramModelNonOverlapV2.forward
 * (Models/models_PMG.py:176-199).
 *   x       (B, input_nc, H, W) NCHW contiguous, fp32 or bf16 (x_dtype)
 *   logits  (B, sum(ncls)) fp32, task-major columns in Tasks.json order (the dict of :198, packed)
 *   gram    optional (B, patch_div^2, Cr, Cr) fp32 — G of :188; null = not materialised for the caller
 *   taps    optional, tests only */
int pmg_forward(const pmg_config* cfg, const void* packed, const void* x, int x_dtype,
                int B, int H, int W, float* logits, float* gram, const pmg_taps* taps,
                void* workspace, size_t workspace_bytes, unsigned flags, pmg_stream_t stream);

/* pmg_forward whose packed logits are ALSO stored to n_peers (<= 15) further destinations — `peer_logits` is a HOST array of
 * DEVICE pointers, each to a (B, sum(ncls)) fp32 buffer that may live in the peer-mapped memory of another GPU.  This is the
 * data-parallel exchange of the path (north_star: "a final gather of per-task logits over NVLink"; reference counterpart: none,
 * the reference is single-device) fused into the compute: the token-stage kernel stores every logit to all destinations from
 * its own epilogue as posted NVLink writes — no collective kernel, no copy-engine transfer, no SM beyond the forward's own.
 * Paths without that kernel (generic, PMG_FLAG_NO_FUSED_ENCODER) copy after the fact.  The caller fences (a barrier across
 * the ranks after this stream's work) before anyone reads a destination; pmg_b200.runner.PushBucketedGather does. */
int pmg_forward_bcast(const pmg_config* cfg, const void* packed, const void* x, int x_dtype,
                      int B, int H, int W, float* logits, float* const* peer_logits, int n_peers,
                      void* workspace, size_t workspace_bytes, unsigned flags, pmg_stream_t stream);

/* Same forward, with a CUDA event recorded on `stream` after every kernel; synchronises the stream and
 * returns the device time of each stage (names are static strings owned by the library).  Used by
 * bench.py for the per-kernel roofline; not a product path. */
typedef struct pmg_stage_time {
    const char* name;
    float ms;
} pmg_stage_time;
int pmg_forward_profile(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H,
                        int W, float* logits, void* workspace, size_t workspace_bytes, unsigned flags,
                        pmg_stream_t stream, pmg_stage_time* stages, int max_stages, int* n_stages);

/* Generic batched Gram microbenchmark kernel (BASELINE.json configs[1]):
 *   G[b] = scale * F[b]·F[b]^T,  F (B, C, HW) row-major bf16 or fp32, G (B, C, C) fp32.
 * Mirrors torch.matmul(patches, patches.transpose(2,3)) / N of models_PMG.py:188 and
 * utils/PMG/PMG_style_transfer.py:44-45. */
int pmg_gram_batched(const void* F, int f_dtype, int B, int C, int HW, float scale, float* G,
                     pmg_stream_t stream);

/* Forward-only Gram export of the patch style transfer — replaces extract_patch_grams_for_image
 * (utils/PMG/PMG_style_transfer.py:11-48) for a whole batch: features = feature_extractor(x) (the conv / BN / LeakyReLU
 * stack, C = last conv width; no chan_proj), disjoint ph x pw windows with ph = Hf / patch_div, pw = Wf / patch_div
 * (feats.unfold: NO padding, a remainder is dropped), G_k = F_k·F_k^T / (ph·pw).  As in the reference (:37, permute +
 * reshape), "patch" k is the run of ph·pw consecutive pixels [k·N, (k+1)·N) of the cropped map flattened row-major.
 *   grams (B, n_patches, C, C) fp32; pmg_style_gram_shape gives n_patches = (Hf/ph)·(Wf/pw) and C.
 * Workspace: pmg_workspace_bytes(cfg, B, H, W). */
int pmg_style_gram_shape(const pmg_config* cfg, int H, int W, int* n_patches, int* C);
int pmg_style_grams(const pmg_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
                    float* grams, void* workspace, size_t workspace_bytes, pmg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Callers' pre/post-processing on the GPU (the steps either side of model(x) in the reference's loops).
 * ------------------------------------------------------------------------------------------------ */
#define PMG_LAYOUT_NCHW 0
#define PMG_LAYOUT_NHWC 1
/* transforms.ToTensor() + transforms.Normalize(mean, std) (utils/datasets_utils.py:212-214) on decoded uint8
 * images: dst[b,c,y,x] = (src/255 - mean[c]) / std[c], fp32 arithmetic in torchvision's operation order
 * (bit-identical in fp32; rounded to nearest-even for bf16).  src (B,3,H,W) or (B,H,W,3) uint8; dst (B,3,H,W).
 * mean3 / std3 are HOST pointers to 3 floats. */
int pmg_preprocess_u8(const void* src, int layout, int B, int H, int W, const float* mean3, const float* std3,
                      int out_dtype, void* dst, pmg_stream_t stream);
/* transforms.Resize(size) + transforms.CenterCrop(crop) (utils/datasets_utils.py:209-210) on decoded uint8 HWC images
 * src (B,H,W,3), bit-identical to the reference's CPU transform: output size / crop offsets as torchvision, pixel values as
 * Pillow's 8-bit BILINEAR resample (22-bit fixed-point taps, horizontal pass then vertical pass, each rounded to uint8).
 * out_mode 0: dst (B,crop,crop,3) uint8 HWC;  1 / 2: ToTensor + Normalize fused, dst (B,3,crop,crop) bf16 / fp32 NCHW (mean3,
 * std3 HOST pointers, may be NULL for out_mode 0).  scratch: pmg_resize_scratch_bytes() device bytes (tap tables).
 * Returns PMG_ERR_UNSUPPORTED when the resized image is smaller than the crop (CenterCrop would pad). */
size_t pmg_resize_scratch_bytes(int H, int W, int size, int crop);
int pmg_resize_crop_u8(const void* src, int B, int H, int W, int size, int crop, const float* mean3, const float* std3,
                       int out_mode, void* dst, void* scratch, size_t scratch_bytes, pmg_stream_t stream);
/* Per-task softmax / max / arg-max / threshold of Functions/PMG_evaluation.py:144-146 on the packed logits
 * (B, sum(ncls)): probs (B, sum(ncls)), preds (B, n_tasks) int32 with -1 where max prob < prob_threshold,
 * max_probs (B, n_tasks); any output may be NULL (not all).  ncls is a HOST array. */
int pmg_postprocess(const float* logits, int B, int n_tasks, const int32_t* ncls, float prob_threshold,
                    float* probs, int32_t* preds, float* max_probs, pmg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * PM model (PatchGAN multi-task, no Gram branch) — BASELINE.json configs[3].
 * Reference boundary: Models/models_PM.py:79-194 class MultiTaskPatchGAN (forward :159-194,
 * TaskHeadImproved.forward :54-76); built by test_PM.py:153-164 with ndf=64, norm="instance".
 * Supported: norm="instance", attn_use_se=false, attn_softmax_spatial=true (the committed
 * hyperparameters_PM.json); anything else is rejected by the host module at construction.
 * ------------------------------------------------------------------------------------------------ */
typedef struct pm_config {
    int32_t input_nc;            /* 3 */
    int32_t ndf;                 /* 64; must be a multiple of 32 */
    float patch_size;            /* 73: drives the layer-count loop (models_PM.py:121-132) */
    int32_t n_tasks;
    int32_t ncls[PMG_MAX_TASKS]; /* classes per task, tasks_dict order */
    float attn_tau;              /* 0.7 */
    int32_t ablate_attention;    /* 1 = uniform attention (GAP), models_PM.py:59-63 */
} pm_config;

/* Parameter order for pm_pack_weights (fp32, torch layouts, = reference state_dict order):
 *   per trunk conv: conv.weight (Cout,Cin,4,4), conv.bias, instnorm.weight, instnorm.bias
 *   per task:       attn_conv.weight (1,C,1,1), attn_conv.bias (1), cls_conv.weight (ncls,C,1,1) */
int pm_param_count(const pm_config* cfg);
size_t pm_packed_bytes(const pm_config* cfg);
int pm_pack_weights(const pm_config* cfg, const float* const* params, int n_params, void* packed,
                    size_t packed_bytes, pmg_stream_t stream);
size_t pm_workspace_bytes(const pm_config* cfg, int B, int H, int W);
/* spatial size (Hf, Wf) and channel count C of the trunk output for an H×W input; returns pmg_status */
int pm_feature_shape(const pm_config* cfg, int H, int W, int* Hf, int* Wf, int* C);
#define PM_FLAG_NO_IGEMM 1u /* pm_forward: run every trunk layer as im2col + GEMM (default: implicit GEMM for the layers with >= 64 input channels) */
#define PM_FLAG_NO_FUSED_STATS 4u /* pm_forward: InstanceNorm statistics by a separate pass over the convolution output (default: from the convolution epilogue where every M tile lies inside one sample) */
#define PM_FLAG_NO_FUSED_CONV0 8u /* pm_forward: first layer as im2col + GEMM + InstanceNorm kernels (default: pm_conv0.cuh — the convolution computed twice around a statistics merge, straight from the NCHW image) */
#define PM_FLAG_IGEMM_N128 2u /* pm_forward: keep the 128-wide N tile in the implicit GEMMs (default: 256 where 256 divides the output channels) */

/* The forward pass — replaces MultiTaskPatchGAN.forward (models_PM.py:159-194).
 *   logits (B, sum(ncls)) fp32, task-major columns; attn optional (B, n_tasks, Hf, Wf) fp32 = the per-task
 *   spatial attention A of return_full=True; feats optional (B, Hf, Wf, C) fp32 pixel-major trunk output. */
int pm_forward(const pm_config* cfg, const void* packed, const void* x, int x_dtype, int B, int H, int W,
               float* logits, float* attn, float* feats, void* workspace, size_t workspace_bytes,
               unsigned flags, pmg_stream_t stream);
/* Tuning hook (tools/bench_extra.py): images per internal pass of pm_forward (default 128) and passes in flight (1 or 2,
 * default 2: the second pass runs on an internal stream forked from / joined to `stream` with events, so the HBM-bound
 * InstanceNorm kernels of one pass overlap the tensor-bound convolutions of the other).  Changes what
 * pm_workspace_bytes returns.  Process-global; not thread safe. */
void pm_debug_set_chunking(int chunk, int lanes);

/* Building block of the fast path, exported for unit tests and micro-benchmarks: one tcgen05 + TMA GEMM
 *   out[M,N] (fp32) = (a_hi[M,K] + a_lo[M,K]) · b[N,K]^T      bf16 row-major operands, a_lo may be null
 * bn = N-tile (32, 64, 96, 128 or 192; N % bn == 0), K % 64 == 0.  splits > 1 writes split-K partial sums
 * out[s][M,N] (s < effective splits) that the caller adds up. */
int pmg_tc_gemm(const void* a_hi, const void* a_lo, const void* b, int M, int N, int K, int bn, int splits,
                float* out, pmg_stream_t stream);

/* Debug / tuning hook, not a product path: when set to a device buffer of 512 int64, CTA 0 of the fused token-stage
 * kernel records clock64() stamps of its warp roles there (scratch/enc_trace.py decodes them).  NULL switches it off.
 * Process-global; not thread safe. */
void pmg_debug_set_trace(void* dev_buf_512_i64);
/* Hardware probe (tools/tmem_a_probe.py): D[128x64] fp32 = A[128x64] · B[64x64]^T, bf16 row-major inputs, with the A
 * operand of the MMA read from tensor memory. */
int pmg_debug_tmem_a(const void* A, const void* B, float* D, pmg_stream_t stream);

/* Hardware probe (tools/load_probe.py): device time of streaming a (B,3,224,224) bf16 batch into shared memory with one CTA per
 * SM and `stages` 14,336-byte loads in flight; mode 0 = the front kernel's 4-D tensor-map box, 1 = contiguous cp.async.bulk,
 * 2 = 2-D tensor map of 32 consecutive rows, 3 = mode 0 without the L2 evict-first hint.  ms_out is a HOST pointer. */
int pmg_debug_load_probe(const void* x, int B, int mode, int stages, float* ms_out, pmg_stream_t stream);

/* Same contract for N = 192 on CTA PAIRS (tcgen05 cta_group::2, clusters of 2): out[s][M,192] split-K partial sums
 * (s < effective splits, always written as partials, also for splits == 1). */
int pmg_tc_gemm2(const void* a_hi, const void* a_lo, const void* b, int M, int K, int splits, float* out, pmg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PMG_B200_H */
