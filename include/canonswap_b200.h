/*
 * canonswap_b200 -- C ABI of the B200-native CanonSwap per-frame generator hot path.
 *
 * The reference (Pixel-Talk/CanonSwap) has no FFI: its "operator API" for this path is a set of
 * nn.Module attributes / wrapper methods on `can_swapper` (reference src/can_swap_e2e.py:39-348)
 * called by the per-frame loop of `CanSwapPipeline.execute` (src/can_swap_pipeline_e2e.py:223-283).
 * Each entry point below replaces one of those calls; the reference interface it stands in for is
 * cited next to it.  Signatures are plain C: opaque context, raw device/host pointers, sizes, a
 * CUDA stream passed as void* (cudaStream_t).  No torch types.
 *
 * Conventions
 *   - All tensor arguments are DEVICE pointers to contiguous fp32 in the reference's own layouts
 *     (NCHW / NCDHW) unless stated otherwise; cs_load_weights takes HOST pointers.
 *   - h = net_h/4, w = net_w/4 is the feature-volume resolution; the volume is [B,32,16,h,w].
 *   - Every call enqueues work on `stream` and returns without synchronising.
 *   - Return value: 0 on success, negative cs_status otherwise; cs_last_error() gives the text.
 *   - The library never falls back to a CPU or ATen/cuDNN path.
 */
#ifndef CANONSWAP_B200_H
#define CANONSWAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CS_API __attribute__((visibility("default")))
#else
#define CS_API
#endif

typedef struct cs_ctx cs_ctx;

enum cs_status {
  CS_OK = 0,
  CS_ERR_INVALID = -1,      /* bad argument / unsupported shape            */
  CS_ERR_CUDA = -2,         /* CUDA runtime or driver error                */
  CS_ERR_WEIGHTS = -3,      /* missing / mis-shaped tensor in the state dict */
  CS_ERR_STATE = -4,        /* call order (weights or identity not set)    */
  CS_ERR_NOMEM = -5
};

enum cs_dtype { CS_F32 = 0, CS_I64 = 1, CS_U8 = 2 };

/* One entry of a reference state_dict, e.g. name = "warping_module.third.conv.weight"
 * (net prefix = key of combined_weights.pth, reference can_swap_e2e.py:93-98). data: HOST pointer. */
typedef struct cs_tensor_desc {
  const char* name;
  const void* data;
  int32_t dtype;            /* cs_dtype */
  int32_t ndim;
  int64_t shape[6];
} cs_tensor_desc;

/* flags of cs_frame */
#define CS_FRAME_IN_U8_HWC   1   /* frames are [B,net_h,net_w,3] u8 (else [B,3,net_h,net_w] fp32)   */
#define CS_FRAME_DEBUG_DECODES 2 /* also run the two debug decodes of pipeline_e2e.py:248,257 (results discarded) */
#define CS_FRAME_MOTION      8   /* derive kp_t / kp_can from the frames themselves with the motion extractor (needs the
                                    'motion_extractor' tensors): x_t = transform_keypoint(M(I)), x_can = scale * kp
                                    (can_swap_pipeline_e2e.py:112-125,231-243); the kp_t / kp_can arguments may be null */
#define CS_FRAME_V2I         4   /* video-to-image per-frame body (can_swap_pipeline_v2i.py:308-309) instead of the e2e one:
                                    out = warp_decode(extract_feature_3d(frames), kp_source = kp_t, kp_driving = kp_can);
                                    no identity needed (the swap ran once per source, outside the loop) */
#define CS_FRAME_V2I_FEATURE 16  /* with CS_FRAME_V2I: `frames` is not images but ONE appearance volume [1,32,16,h,w] fp32 (NCDHW),
                                    shared by the B samples -- the v2i loop extracts the features of the SAME swapped canonical image
                                    every frame (can_swap_pipeline_v2i.py:308), so they are computed once per source and stay
                                    resident: out = G(W.forward(feature, kp_driving = kp_can, kp_source = kp_t)) */

/* options of cs_set_option */
#define CS_OPT_CONV_IMPL 1       /* 0 = auto (tcgen05 where eligible), 1 = force SIMT fp32 convs (debug) */
#define CS_OPT_TC_PASSES 3       /* split-fp16 MMA passes of the tcgen05 conv: 3 (default, fp32-grade) | 2 | 1 (measurement only) */
#define CS_OPT_TC_SETS 4         /* cap on TMEM accumulator sets per tile (0 = automatic; 1 = single accumulator, measurement only) */
#define CS_OPT_TC_COMP 5         /* tensor-core accumulate-truncation compensation per chained MMA, units of 1e-10 (default 170, 0 = off) */
#define CS_OPT_TC_PAIR 6         /* tcgen05 pair mode (cta_group::2 over 2-CTA clusters) for wide N tiles: 0 = off, 1 = on (default),
                                    n > 1 = only for convs with at least n K iterations */
#define CS_OPT_TC_STACKED3 7     /* 1 (default) = depth-stacked kernel for the 32->32 3x3x3 volume convs, 0 = generic implicit GEMM */
#define CS_OPT_TC_DOUBLE_BUFFER 8 /* 1 (default) = two TMEM accumulator buffers where they fit: a tile's epilogue overlaps the next tile's MMAs */
#define CS_OPT_TC_BN_MAX 9       /* cap on the N tile of convs packed AFTER the call (0 = automatic; experiments) */
#define CS_OPT_TC_SINGLE_CHAIN 11 /* convs whose whole MMA chain is at most this long accumulate in ONE TMEM set (0 = never) */
#define CS_OPT_TC_CHAIN_MAX 12   /* longest hi*hi MMA chain per TMEM accumulator set for convs packed AFTER the call: the N tile is halved
                                    until it holds (0 = default 256) */
#define CS_OPT_WINOGRAD 13       /* (2 = only the adaptive convs in Winograd form: measured 152.7 against 155.7 frames/s with everything) 1 (default) = the wide 3x3 2-D convs (adaptive convs of the swap module, SPADE conv_0 / conv_1, refine ResBlock2d) in
                                    Winograd F(2x2,3x3) form, 0 = direct implicit GEMM */
#define CS_OPT_TC_POSCOMP 14     /* position-dependent pre-compensation of the tensor core's accumulate truncation, folded into the packed
                                    weights: units of 1e-10 per truncation event (default 330, 0 = off: the epilogue then applies the
                                    constant CS_OPT_TC_COMP factor).  Options 3, 4, 8, 9, 11, 12 and 14 shape the packed weights and
                                    must be set before cs_load_weights (CS_ERR_STATE afterwards). */
#define CS_OPT_TEST_AMUL 15      /* cs_test_conv only: log2 of the activation pre-scale applied to the test conv's operand (default 0) */
#define CS_OPT_LANES 10         /* 1 | 2 (default) | 4: a graph-captured cs_frame runs as this many concurrent sub-batches (forked streams) */
#define CS_OPT_USE_GRAPH 2       /* 1 = capture cs_frame into a CUDA graph per batch size (default 0)  */

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces can_swapper.__init__ module construction (can_swap_e2e.py:60-68) for the five hot-path
 * networks. net_h, net_w: generator input size (multiple of 128; 256 for 512-px frames). */
CS_API int cs_create(cs_ctx** out, int device, int max_batch, int net_h, int net_w);
CS_API void cs_destroy(cs_ctx* ctx);
/* Text of the last error on this ctx (ctx may be NULL for cs_create failures). */
CS_API const char* cs_last_error(const cs_ctx* ctx);
CS_API int cs_set_option(cs_ctx* ctx, int option, int value);
/* Number of kernel launches issued by this ctx so far (for bench.py's gpu_launches). */
CS_API int64_t cs_launch_count(const cs_ctx* ctx);
/* Bytes of device workspace + packed weights owned by the ctx. */
CS_API size_t cs_workspace_bytes(const cs_ctx* ctx);

/* Replaces can_swapper.load_cpk (can_swap_e2e.py:87-100): ingests the raw reference state_dicts,
 * folds eval-mode BatchNorm and spectral-norm sigma, permutes and re-lays weights for the kernels. */
CS_API int cs_load_weights(cs_ctx* ctx, const cs_tensor_desc* table, int n);

/* Per-source state: derives the 14 style vectors and demodulated weight sets of transfer_model2
 * from the ArcFace identity (adaptive_modulate.py:148-155). id_dev: device [512] fp32. */
CS_API int cs_set_identity(cs_ctx* ctx, const float* id_dev, void* stream);

/* ---- per-stage entry points (reference module forward()s) ---------------------------------- */
/* AppearanceFeatureExtractor.forward (appearance_feature_extractor.py:38-48):
 * img [B,3,net_h,net_w] -> f3d [B,32,16,h,w] */
CS_API int cs_appearance(cs_ctx* ctx, const float* img, float* f3d, int B, void* stream);

/* WarpingNetwork.warp (warping_network.py:49-62): -> out3d [B,32,16,h,w], occ [B,1,h,w];
 * deformation [B,16,h,w,3] optional (may be NULL). kp_* are [B,21,3]. */
CS_API int cs_warp(cs_ctx* ctx, const float* f3d, const float* kp_source, const float* kp_driving,
            float* out3d, float* occ, float* deformation, int B, void* stream);

/* WarpingNetwork.warp_out (warping_network.py:64-71): f3d [B,32,16,h,w], occ [B,1,h,w] or NULL
 * -> out [B,256,h,w] */
CS_API int cs_warp_out(cs_ctx* ctx, const float* f3d, const float* occ, float* out, int B, void* stream);

/* WarpingNetwork.forward (warping_network.py:83-111) -> out [B,256,h,w], occ [B,1,h,w],
 * deformation [B,16,h,w,3] (occ/deformation may be NULL). */
CS_API int cs_warp_forward(cs_ctx* ctx, const float* f3d, const float* kp_driving, const float* kp_source,
                    float* out, float* occ, float* deformation, int B, void* stream);

/* transfer_model2.forward (adaptive_modulate.py:522-554) with the identity given to cs_set_identity.
 * masks: optional [7,B,1,h,w] (return_mask=True), may be NULL. */
CS_API int cs_swap(cs_ctx* ctx, const float* f3d, float* out3d, float* masks, int B, void* stream);

/* G3d.forward (adaptive_modulate.py:721-733) */
CS_API int cs_refine(cs_ctx* ctx, const float* f3d, float* out3d, int B, void* stream);

/* SPADEDecoder.forward (spade_generator.py:41-59): feat [B,256,h,w] -> img [B,3,8h,8w] in [0,1];
 * img_u8 optional [B,8h,8w,3] u8 with can_swapper.parse_output semantics (can_swap_e2e.py:314-322). */
CS_API int cs_spade(cs_ctx* ctx, const float* feat, float* img, uint8_t* img_u8, int B, void* stream);

/* ---- the fused per-frame loop body (pipeline_e2e.py:242-267) -------------------------------- */
/* frames: see CS_FRAME_IN_U8_HWC; kp_t, kp_can [B,21,3] (x_t = x_t_info['x_s'], x_can = scale*kp);
 * out_f32 [B,3,2*net_h,2*net_w] and/or out_u8 [B,2*net_h,2*net_w,3] (either may be NULL). */
CS_API int cs_frame(cs_ctx* ctx, const void* frames, const float* kp_t, const float* kp_can,
             float* out_f32, uint8_t* out_u8, int B, int flags, void* stream);

/* ---- motion extractor M + keypoint transform ------------------- */
#define CS_MOTION_HEADS 328      /* kp 63 | scale 1 | pitch 66 | yaw 66 | roll 66 | t 3 | exp 63 (registration order of the heads) */
/* can_swapper.motion_extractor(x) (src/can_swap_e2e.py:64, src/modules/motion_extractor.py:33-35 -> convnextv2.py:110-144):
 * img [B,3,net_h,net_w] fp32 in [0,1] -> heads [B,CS_MOTION_HEADS], the seven raw Linear outputs concatenated. */
CS_API int cs_motion(cs_ctx* ctx, const float* img, float* heads, int B, void* stream);
/* can_swapper.transform_keypoint (src/can_swap_e2e.py:226-254) with headpose_pred_to_degree / get_rotation_matrix
 * (src/utils/camera.py:14-73): heads -> x_s [B,21,3]; optional x_can = scale * kp [B,21,3] (can_swap_pipeline_e2e.py:242),
 * R [B,3,3], deg [B,3] = pitch, yaw, roll in degrees (get_kp_info, src/can_swap_e2e.py:191-197). */
CS_API int cs_keypoints(cs_ctx* ctx, const float* heads, float* x_s, float* x_can, float* R, float* deg, int B, void* stream);

/* ---- paste-back of the swapped crop into the full frame ------- */
#define CS_PASTE_MAX_BATCH 16
/* prepare_paste_back(mask, M_c2o, dsize, if_float=True) + paste_back(img_crop, M_c2o, img_ori, mask_ori), reference
 * src/utils/crop.py:515-529 (two cv2.warpAffine INTER_LINEAR + float blend per frame, src/can_swap_pipeline_e2e.py:277-282),
 * fused, bit-exact with OpenCV's fixed-point arithmetic.  img_crop [B,hc,wc,3] u8, mask_crop [B,hc,wc] f32 (the soft mask; the
 * reference stacks it to 3 equal channels), M_c2o HOST pointer [B][6] doubles (rows 0..1 of the crop->original matrix),
 * img_ori / out [B,H,W,3] u8 (out may alias img_ori).  Needs no weights. */
CS_API int cs_paste_back(cs_ctx* ctx, const uint8_t* img_crop, const float* mask_crop, const double* M_c2o, const uint8_t* img_ori,
                         uint8_t* out, int B, int hc, int wc, int H, int W, void* stream);

/* SoftErosion.forward (reference src/utils/crop.py:21-47; pipeline_e2e.py:42,275 uses kernel_size 21, threshold 0.9, iterations 3):
 * mask [B,H,W] f32 -> out [B,H,W] f32 soft mask, hard [B,H,W] u8 = (x >= threshold) or NULL.  kernel: DEVICE pointer to the module's
 * normalised [kernel_size^2] weight buffer (the binding builds it exactly as the reference's __init__ does).
 * Normalisation: the reference divides the below-threshold values by their maximum over the WHOLE tensor (crop.py:44); it only ever
 * passes one image (pipeline_e2e.py:275).  Here every image of the batch is normalised by ITS OWN maximum, i.e. a batch of B equals
 * B reference calls of one image each -- not one reference call on a [B,1,H,W] tensor. */
CS_API int cs_soft_erosion(cs_ctx* ctx, const float* mask, const float* kernel, float* out, uint8_t* hard, int B, int H, int W,
                           int kernel_size, float threshold, int iterations, void* stream);

/* Activation-scale calibration.  The tcgen05 convs compute on split-fp16 operands (value = hi + lo, both fp16): full fp32-grade precision
 * needs |activation| within about [0.06, 65504] per tensor.  Weights are pre-scaled at load; activations get one power-of-two scale per
 * conv, chosen from the largest |input| seen while a representative batch runs:
 *     cs_calibrate(ctx, 1, NULL, 0);  cs_frame(...) [one or more batches];  cs_calibrate(ctx, 0, maxima, cap);
 * phase 2 resets every scale to 1.  Scaling by a power of two is exact, so results change only where operands were leaving fp16's
 * normal range.  maxima (may be NULL): the measured max |input| per conv, in the library's internal conv order. */
CS_API int cs_calibrate(cs_ctx* ctx, int phase, float* maxima, int cap);

/* Face-parsing post-processing, the step before the path (reference src/can_swap_pipeline_e2e.py:183-190; the Segformer itself is an
 * external HF model and stays outside): logits [B,C,h,w] fp32 -> F.interpolate(size=(H,W), bilinear, align_corners=False) -> argmax over C
 * -> isin(valid) fused in one kernel.  valid_classes: bit c set = class c belongs to valid_list (:48: {1,2,4,5,6,7,10,11,12}).
 * mask [B,H,W] f32 (1.0 / 0.0: the input of cs_soft_erosion, no host round trip), labels [B,H,W] i32 or NULL. */
CS_API int cs_parse_mask(cs_ctx* ctx, const float* logits, int B, int C, int h, int w, int H, int W, uint64_t valid_classes, float* mask,
                         int32_t* labels, void* stream);

/* ---- per-kernel-family timing (measurement only) -------------------------------------------- */
/* enable != 0: bracket every kernel launch of this ctx with CUDA events on the launching stream. */
CS_API int cs_profile(cs_ctx* ctx, int enable);
/* Synchronises and writes out[6][4] = per family {tcgen05 conv, SIMT conv, prep, stats, sampling, other}:
 * {total ms, algorithmic flops, algorithmic bytes, launches}; clears the records. */
CS_API int cs_profile_read(cs_ctx* ctx, double* out);
/* Synchronises and writes one CSV line per recorded launch ("index,family,ms,flops,bytes,description") into buf
 * (NUL-terminated, truncated at cap); does not clear the records. */
CS_API int cs_profile_dump(cs_ctx* ctx, char* buf, int cap);

/* ---- kernel-level entry points (unit tests / profiling) ------------------------------------- */
/* Generic "same"-style convolution on channels-last fp32 through the library's conv kernels.
 * x [B,D,H,W,Cin] -> y [B,Do,Ho,Wo,Cout]; w in PyTorch layout [Cout,Cin,KD,KH,KW] (device), bias
 * [Cout] or NULL. impl: 0 auto, 1 SIMT fp32, 2 tcgen05 split-fp16, 3 tcgen05 depth-stacked 7x7x7 kernel,
 * 4 tcgen05 depth-stacked 32->32 3x3x3 kernel, 5 Winograd F(2x2,3x3) form of a 3x3 2-D conv (Cin % 32 == 0, Cout % 256 == 0),
 * 6 phase form: the 3x3 / 3x3x3 conv of nearest-upsample(x, (1,2,2)) computed on x itself, y [B,D,2H,2W,Cout] (Cout % 16 == 0, <= 256).
 * act: 0 none 1 relu 2 lrelu 3 sigmoid */
CS_API int cs_test_conv(cs_ctx* ctx, const float* x, const float* w, const float* bias, float* y,
                 int B, int D, int H, int W, int Cin, int Cout, int KD, int KH, int KW,
                 int PD, int PH, int PW, int act, float slope, int impl, void* stream);
/* F.grid_sample(inp, grid, align_corners=False) 5-D trilinear zeros-padding on NCDHW fp32:
 * inp [B,C,D,H,W] (C multiple of 4), grid [B,D,H,W,3] -> out [B,C,D,H,W] */
CS_API int cs_test_grid_sample3d(cs_ctx* ctx, const float* inp, const float* grid, float* out,
                          int B, int C, int D, int H, int W, void* stream);
/* per-(b,c) mean / rstd (biased variance, eps) of NCHW-style [B,C,S] fp32 */
CS_API int cs_test_instance_stats(cs_ctx* ctx, const float* x, float* mean, float* rstd,
                           int B, int C, int S, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CANONSWAP_B200_H */
