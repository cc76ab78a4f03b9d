/*
 * dvsloss.h -- C ABI of libdvsloss.so: the B200 (sm_100a) implementation of the
 * Monodepth2-style view-synthesis loss that dominates Deep-Visual-SLAM's VO training step.
 *
 * The reference has NO native boundary on this path: it is eager PyTorch
 * (vo/learner_new.py:60-74,132-258 composing vo/learner_func.py:16-207 == model/layers.py:16-248).
 * The functions below are what a ctypes / pybind shim inside the reference's Python modules binds
 * instead of those ATen op sequences (INTEGRATION.md shows the stub).  Each entry point names the
 * reference code it replaces.
 *
 * Conventions
 *   - plain C: no torch types, no exceptions, no exit(); return 0 (DVS_OK) or a negative DVS_E* code.
 *   - every pointer is a DEVICE pointer to contiguous memory (fp32 unless stated), owned by the caller,
 *     never retained after the call returns; arrays of pointers (disp[], src[], ...) are HOST arrays.
 *   - every call is asynchronous on the given cudaStream_t (passed as void*; NULL = legacy default stream).
 *   - no hidden allocation: scratch memory is a caller-provided workspace sized by
 *     dvs_loss_workspace_bytes(); distinct (stream, workspace) pairs may run concurrently.
 *   - tensor layouts are the reference's: images [B,3,H,W], disparity [B,1,h_s,w_s], matrices [B,4,4]
 *     row-major, sampling grids [B,H,W,2], point clouds [B,4,H*W].
 */
#ifndef DVSLOSS_H_
#define DVSLOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVS_MAX_SCALES 4
#define DVS_MAX_SOURCES 4

enum {
  DVS_OK = 0,
  DVS_EINVAL = -1,   /* bad shape / null pointer / unsupported N or S                */
  DVS_ECUDA = -2,    /* a CUDA runtime call failed (see dvs_last_cuda_error())        */
  DVS_EWORKSPACE = -3 /* workspace pointer null or misaligned                         */
};

/* Problem shape.  dh/dw: size of the disparity map of each scale (the reference: H>>s, W>>s,
 * model/depthnet.py:87-88); they are bilinearly up-sampled to HxW (vo/learner_new.py:136-140). */
typedef struct DvsShape {
  int32_t B, H, W;
  int32_t N;                     /* source frames, 1..DVS_MAX_SOURCES (reference: 2)  */
  int32_t S;                     /* scales, 1..DVS_MAX_SCALES (reference: 4)          */
  int32_t dh[DVS_MAX_SCALES];
  int32_t dw[DVS_MAX_SCALES];
} DvsShape;

/* Hyper-parameters = config['Train'] keys read at vo/learner_new.py:31-41. */
typedef struct DvsParams {
  float min_depth;               /* 0.1   */
  float max_depth;               /* 10.0  */
  float ssim_ratio;              /* 0.85  */
  float smoothness_ratio;        /* 1e-3  */
  float eps;                     /* Project3D eps, 1e-7 (vo/learner_func.py:140)      */
  int32_t auto_mask;             /* 1     */
} DvsParams;

int dvs_version(void);
const char* dvs_error_string(int code);
/* cudaError_t of the last failing CUDA call on this host thread (0 if none). */
int dvs_last_cuda_error(void);

/* Measurement hook (bench.py): when enabled, dvs_photometric_forward brackets its dominant kernel (the fused
 * tile kernel) with CUDA events on the caller's stream; dvs_last_tile_kernel_ms waits for the most recent
 * one and returns its device duration.  Process-wide, not thread-safe; off by default. */
int dvs_set_profiling(int enabled);
int dvs_last_tile_kernel_ms(float* ms);

/* In-kernel automask noise under CUDA-graph replay: register a device-resident uint64 step counter for `device` (NULL to
 * clear).  Every dvs_photometric_forward* call on that device that draws its noise in the kernel (noise == NULL) and is not
 * given its own offset_dev adds *counter to `offset`; a caller that increments the counter on the stream before each call
 * gets fresh draws on every replay of a captured graph.  The pointer is not owned; it must outlive the calls. */
int dvs_set_noise_counter(int device, const uint64_t* counter);

/* Bytes of scratch the fused loss needs for this shape (>=256-byte aligned pointer expected). */
int dvs_loss_workspace_bytes(const DvsShape* shape, size_t* bytes);

/* ------------------------------------------------------------------------------------------
 * Fused view-synthesis loss.  Replaces MonodepthTrainer._generate_images_pred
 * (vo/learner_new.py:132-172) + _compute_losses (:175-258) for all S scales and N sources:
 * disparity up-sampling, disp_to_depth, BackprojectDepth, Project3D, border grid_sample,
 * SSIM+L1, identity automask, min over sources, edge-aware smoothness, per-scale reduction.
 *
 *   disp[s]      [B,1,dh[s],dw[s]]        target, src[i]  [B,3,H,W]
 *   K, inv_K     [B,4,4] (scale-0 intrinsics, vo/learner_new.py:155-156)      T[i] [B,4,4] cam_T_cam
 *   noise[s]     [B,N,H,W] standard-normal draws of vo/learner_new.py:228 (scaled by 1e-5 inside);
 *                noise == NULL -> an in-kernel counter-based generator keyed by (seed, offset, s, pixel)
 *                is used instead (same distribution, not the same stream); ignored if !auto_mask.
 *   loss_per_scale [S]  = losses["loss/s"]          loss_total [1] = losses["loss"]
 *   sel[s]       optional uint8 [B,H,W]: argmin channel over [identity_0..N-1, reproj_0..N-1]
 *                (identity_selection/s of the reference == sel >= N);  sel or sel[s] may be NULL.
 *   ugrad_disp[s], ugrad_T  optional.  When non-NULL the same pass also produces the gradient of
 *                each loss/s w.r.t. disp[s] (ugrad_disp[s], same shape as disp[s]) and w.r.t. T
 *                (ugrad_T, [S,N,B,4,4]), so that backward is a scaling (dvs_photometric_backward).
 *                Both or neither must be given.
 * ------------------------------------------------------------------------------------------ */
int dvs_photometric_forward(const DvsShape* shape, const DvsParams* params,
                            const float* const* disp, const float* target, const float* const* src,
                            const float* K, const float* inv_K, const float* const* T,
                            const float* const* noise, uint64_t seed, uint64_t offset,
                            float* loss_per_scale, float* loss_total,
                            uint8_t* const* sel,
                            float* const* ugrad_disp, float* ugrad_T,
                            void* workspace, void* stream);

/* The same with reduced-precision INPUTS read directly by the kernel (two sources only, N == 2):
 *   disp_dtype  DVS_DTYPE_F32 | DVS_DTYPE_BF16   disp[s] as DepthNet emits them under bf16 autocast (vo/train.py:177-181)
 *   image_dtype DVS_DTYPE_F32 | DVS_DTYPE_U8     target / src[i] as the loader decodes them, before ToTensor
 *                                                (vo/dataset/common.py:39-46,77); x/255 is formed in-register, correctly
 *                                                rounded, so results are bit-identical to the fp32-expanded tensors
 * All arithmetic, the unit gradients and every output stay fp32. */
enum { DVS_DTYPE_F32 = 0, DVS_DTYPE_BF16 = 1, DVS_DTYPE_U8 = 2 };
int dvs_photometric_forward_ex(const DvsShape* shape, const DvsParams* params,
                               const void* const* disp, int disp_dtype,
                               const void* target, const void* const* src, int image_dtype,
                               const float* K, const float* inv_K, const float* const* T,
                               const float* const* noise, uint64_t seed, uint64_t offset,
                               float* loss_per_scale, float* loss_total,
                               uint8_t* const* sel,
                               float* const* ugrad_disp, float* ugrad_T,
                               void* workspace, void* stream);
/* dvs_photometric_backward writing grad_disp[s] in grad_dtype (DVS_DTYPE_F32 | DVS_DTYPE_BF16, round-to-nearest-even):
 * what autograd hands back to a bf16 disparity tensor. */
int dvs_photometric_backward_ex(const DvsShape* shape, const float* grad_per_scale,
                                const float* const* ugrad_disp, const float* ugrad_T,
                                void* const* grad_disp, int grad_dtype, float* const* grad_T, void* stream);

/* The same with the POSE PARAMETERS as inputs instead of matrices (no separate pose launches, nothing 4x4 crosses HBM):
 * axisangle[i], translation[i] [B,3] and invert[i] (host ints) are what the reference feeds
 * transformation_from_parameters (vo/learner_func.py:29-46; vo/learner_new.py:124-127: invert for the left frame); the
 * pre-pass kernel builds T_i, the post-pass kernel chains d loss / d T_i back to the six numbers.
 *   offset_dev   optional device counter added to `offset` for the in-kernel noise generator: a caller that increments
 *                it on the stream gets fresh draws on every replay of a captured CUDA graph (NULL: `offset` alone).
 *   ugrad_pose   [S,N,B,6] (d loss/s / d axisangle, d loss/s / d translation) followed by [S,B] floats of internal
 *                coefficients; dvs_photometric_backward_pose combines it with the upstream gradients. */
int dvs_photometric_forward_pose(const DvsShape* shape, const DvsParams* params,
                                 const void* const* disp, int disp_dtype,
                                 const void* target, const void* const* src, int image_dtype,
                                 const float* K, const float* inv_K,
                                 const float* const* axisangle, const float* const* translation, const int32_t* invert,
                                 const float* const* noise, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                 float* loss_per_scale, float* loss_total,
                                 uint8_t* const* sel,
                                 float* const* ugrad_disp, float* ugrad_pose,
                                 void* workspace, void* stream);
int dvs_photometric_backward_pose(const DvsShape* shape, const float* grad_per_scale,
                                  const float* const* ugrad_disp, const float* ugrad_pose,
                                  void* const* grad_disp, int grad_dtype,
                                  float* const* grad_axisangle, float* const* grad_translation, void* stream);

/* Backward of the above given what forward stored: for upstream gradients
 * grad_per_scale[s] = d objective / d loss/s (device, [S]; add grad_total/S to each if the total is used)
 *   grad_disp[s] = grad_per_scale[s] * ugrad_disp[s]            (in place allowed: grad_disp[s]==ugrad_disp[s])
 *   grad_T[i]    = sum_s grad_per_scale[s] * ugrad_T[s,i]       ([B,4,4] each)
 * Equals what autograd produces through vo/learner_new.py:132-258 for d/d outputs[("disp",s)]
 * and d/d outputs[("cam_T_cam",0,f)]. */
int dvs_photometric_backward(const DvsShape* shape, const float* grad_per_scale,
                             const float* const* ugrad_disp, const float* ugrad_T,
                             float* const* grad_disp, float* const* grad_T, void* stream);

/* Stand-alone backward that recomputes the forward from the inputs (nothing saved between passes
 * except nothing at all): same inputs as forward plus grad_per_scale; writes grad_disp[s], grad_T[i]. */
int dvs_photometric_backward_recompute(const DvsShape* shape, const DvsParams* params,
                                       const float* const* disp, const float* target, const float* const* src,
                                       const float* K, const float* inv_K, const float* const* T,
                                       const float* const* noise, uint64_t seed, uint64_t offset,
                                       const float* grad_per_scale,
                                       float* const* grad_disp, float* const* grad_T,
                                       void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Granular operators: the reference's public primitives, one kernel each, forward and backward.
 * ------------------------------------------------------------------------------------------ */

/* disp_to_depth (vo/learner_func.py:16-26): scaled = 1/max + (1/min-1/max)*disp, depth = 1/scaled. n elements. */
int dvs_disp_to_depth_fwd(const float* disp, float* scaled_disp, float* depth, int64_t n,
                          float min_depth, float max_depth, void* stream);
/* grad_disp = (grad_scaled - grad_depth*depth^2) * (1/min-1/max); either upstream may be NULL. */
int dvs_disp_to_depth_bwd(const float* depth, const float* grad_scaled, const float* grad_depth,
                          float* grad_disp, int64_t n, float min_depth, float max_depth, void* stream);

/* F.interpolate(disp, [H,W], mode="bilinear", align_corners=False) (vo/learner_new.py:136-140); C channels. */
int dvs_upsample_bilinear_fwd(const float* in, float* out, int B, int C, int h, int w, int H, int W, void* stream);
int dvs_upsample_bilinear_bwd(const float* grad_out, float* grad_in, int B, int C, int h, int w, int H, int W,
                              void* stream);

/* BackprojectDepth.forward (vo/learner_func.py:130-135): depth [B,1,H,W], inv_K [B,4,4] -> cam_points [B,4,H*W]. */
int dvs_backproject_fwd(const float* depth, const float* inv_K, float* cam_points, int B, int H, int W, void* stream);
/* grad_depth [B,1,H,W] from grad_cam [B,4,H*W] (inv_K receives no gradient in the reference's use). */
int dvs_backproject_bwd(const float* grad_cam, const float* inv_K, float* grad_depth, int B, int H, int W,
                        void* stream);

/* Project3D.forward (vo/learner_func.py:148-159): points [B,4,HW], K,T [B,4,4] -> normalised grid [B,H,W,2]. */
int dvs_project3d_fwd(const float* points, const float* K, const float* T, float* pix, int B, int H, int W,
                      float eps, void* stream);
/* grad_points [B,4,HW] and grad_T [B,4,4] (either may be NULL) from grad_pix [B,H,W,2];
 * workspace: dvs_project3d_bwd_workspace_bytes(). */
int dvs_project3d_bwd_workspace_bytes(int B, int H, int W, size_t* bytes);
int dvs_project3d_bwd(const float* grad_pix, const float* points, const float* K, const float* T,
                      float* grad_points, float* grad_T, int B, int H, int W, float eps,
                      void* workspace, void* stream);

/* F.grid_sample(src, grid, padding_mode="border", align_corners=True), bilinear (vo/learner_new.py:165-170).
 * src [B,C,H,W], grid [B,Ho,Wo,2] -> out [B,C,Ho,Wo]; backward w.r.t. the grid only (images are data). */
int dvs_grid_sample_border_fwd(const float* src, const float* grid, float* out, int B, int C, int H, int W,
                               int Ho, int Wo, void* stream);
int dvs_grid_sample_border_bwd(const float* grad_out, const float* src, const float* grid, float* grad_grid,
                               int B, int C, int H, int W, int Ho, int Wo, void* stream);

/* SSIM.forward (vo/learner_func.py:190-207): x,y [B,C,H,W] -> clamp((1-SSIM)/2,0,1) [B,C,H,W]. */
int dvs_ssim_fwd(const float* x, const float* y, float* out, int B, int C, int H, int W, void* stream);
/* grad_x and/or grad_y (either may be NULL) from grad_out. */
int dvs_ssim_bwd(const float* grad_out, const float* x, const float* y, float* grad_x, float* grad_y,
                 int B, int C, int H, int W, void* stream);

/* compute_reprojection_loss (vo/learner_new.py:60-74): pred,target [B,C,H,W] -> [B,1,H,W]. */
int dvs_reprojection_loss_fwd(const float* pred, const float* target, float* out, int B, int C, int H, int W,
                              float ssim_ratio, void* stream);
int dvs_reprojection_loss_bwd(const float* grad_out, const float* pred, const float* target, float* grad_pred,
                              int B, int C, int H, int W, float ssim_ratio, void* stream);

/* get_smooth_loss (vo/learner_func.py:161-174): disp [B,1,H,W], img [B,C,H,W] -> scalar out[1].
 * workspace: dvs_smooth_loss_workspace_bytes(). */
int dvs_smooth_loss_workspace_bytes(int B, int H, int W, size_t* bytes);
int dvs_smooth_loss_fwd(const float* disp, const float* img, float* out, int B, int C, int H, int W,
                        void* workspace, void* stream);
/* grad_disp [B,1,H,W] = grad_out[0] * d loss / d disp (grad_out is a device scalar). */
int dvs_smooth_loss_bwd(const float* grad_out, const float* disp, const float* img, float* grad_disp,
                        int B, int C, int H, int W, void* stream);

/* transformation_from_parameters (vo/learner_func.py:29-104): axisangle, translation [B,3] -> M [B,4,4]. */
int dvs_pose_matrix_fwd(const float* axisangle, const float* translation, float* M, int B, int invert,
                        void* stream);
int dvs_pose_matrix_bwd(const float* grad_M, const float* axisangle, const float* translation,
                        float* grad_axisangle, float* grad_translation, int B, int invert, void* stream);

/* ToTensor of the reference's loader (vo/dataset/common.py:77): dst[i] = (float)src[i] / 255 for n bytes; exact
 * (IEEE division), so uint8 image batches can cross PCIe as bytes and be expanded on the device. */
int dvs_u8_to_f32(const uint8_t* src, float* dst, int64_t n, void* stream);

/* Disparity head of DepthNet, fused: ReflectionPad2d(1) + Conv2d(C, 1, 3) + Sigmoid (the ("dispconv", s) blocks,
 * model/depthnet.py:57-58,87-88; Conv3x3 = model/layers.py:120-136) in one pass, writing disp_s in the dtype the loss
 * kernel reads.  x: the decoder activation, CHANNELS-LAST [B,H,W,C] (C a multiple of 8, <= 128; 16-byte aligned), fp32 or
 * bf16; weight: the Conv2d weight [1,C,3,3] fp32 as stored; bias [1] fp32 (may be NULL); disp [B,1,H,W] fp32 or bf16.
 * x_pad = 1: x carries the reflected one-pixel ring written by dvs_bias_elu_fwd(pad = 1) ([B,H+2,W+2,C]; only its interior is
 * read); needs C in {8, 16, 32, 64, 128}. */
int dvs_disp_head_fwd(const void* x, int x_dtype, int x_pad, const float* weight, const float* bias, void* disp, int disp_dtype,
                      int B, int C, int H, int W, void* stream);
/* Backward: grad_disp and disp (the saved forward output) [B,1,H,W] in disp_dtype; grad_x channels-last in x_dtype;
 * grad_weight [1,C,3,3] and grad_bias [1] (may be NULL) fp32, fixed-order reductions (bit-reproducible).  With x_pad = 1 grad_x
 * has x's padded layout; its ring is NOT written (the caller zeroes it). */
int dvs_disp_head_bwd_workspace_bytes(int B, int C, int H, int W, size_t* bytes);
int dvs_disp_head_bwd(const void* grad_disp, const void* disp, int disp_dtype, const void* x, int x_dtype, int x_pad,
                      const float* weight, void* grad_x, float* grad_weight, float* grad_bias, int B, int C, int H, int W,
                      void* workspace, void* stream);

/* Decoder glue of DepthNet, fused (model/depthnet.py:77-84: ConvBlock's ELU, model/layers.py:106-117; `upsample`,
 * model/layers.py:196-199; torch.cat with the encoder skip): out = cat([nearest_up2(ELU(x + bias)), skip], channels), one pass.
 * CHANNELS-LAST tensors of one dtype (fp32 or bf16), 16-byte aligned: x [B,h,w,C1] (the convolution output BEFORE bias and
 * ELU; bias fp32 [C1] or NULL: the Conv2d bias, model/layers.py:131, folded in so that the convolution runs without its
 * separate bias pass), skip [B,2h,2w,C2] (NULL iff C2 == 0), out [B,2h,2w,C1+C2]; C1, C2 multiples of 8 (bf16) / 4 (fp32).
 * pad = 1: out is [B,2h+2,2w+2,C1+C2] with the reflected one-pixel ring of nn.ReflectionPad2d(1) (model/layers.py:126) filled
 * in, so that the next Conv3x3 runs as an un-padded stock convolution on it. */
int dvs_elu_up2_cat_fwd(const void* x, const void* skip, const float* bias, void* out, int dtype, int B, int C1, int C2,
                        int h, int w, int pad, void* stream);
/* Backward: grad_out [B,2h,2w,C1+C2] -> grad_x [B,h,w,C1] = ELU'(x + bias) * (2x2 sum of grad_out, fp32, fixed order),
 * grad_skip [B,2h,2w,C2] (NULL iff C2 == 0) and, if grad_bias != NULL, grad_bias [C1] fp32 = the per-channel sums of grad_x
 * (two-stage fixed-order reduction; needs `workspace` of dvs_glue_workspace_bytes(C1) bytes, 256-byte aligned, and C1 / 8
 * (bf16) or C1 / 4 (fp32) a power of two <= 256).  pad = 1: grad_out has the padded layout and the gradients of the ring
 * positions are folded onto the pixels they mirror. */
int dvs_glue_workspace_bytes(int C, size_t* bytes);
int dvs_elu_up2_cat_bwd(const void* x, const void* grad_out, const float* bias, void* grad_x, void* grad_skip, float* grad_bias,
                        int dtype, int B, int C1, int C2, int h, int w, int pad, void* workspace, void* stream);
/* ConvBlock's bias + ELU (model/layers.py:106-117 with the Conv2d bias of :131) in one pass: y = ELU(x + bias[c]),
 * channels-last [B,H,W,C], y may alias x.  Backward from the OUTPUT, as nn.ELU(inplace=True) does: grad_x = grad_y *
 * (y <= 0 ? y + 1 : 1); grad_bias [C] fp32 (optional) as above.  pad = 1: y and grad_y are [B,H+2,W+2,C] with the reflected ring
 * (written / folded as for dvs_elu_up2_cat_*); x and grad_x stay [B,H,W,C]. */
int dvs_bias_elu_fwd(const void* x, const float* bias, void* y, int dtype, int B, int C, int H, int W, int pad, void* stream);
int dvs_bias_elu_bwd(const void* y, const void* grad_y, void* grad_x, float* grad_bias, int dtype, int B, int C, int H, int W,
                     int pad, void* workspace, void* stream);

/* Supervised-depth path (depth/depth_learner.py).  SILog loss (:75-95): over the n elements with valid[e] != 0,
 * d = log(max(pred, 1e-6)) - log(target), loss = sqrt(mean(d^2) - variance_focus * mean(d)^2).  stats [4] receives
 * {loss, mean(d), count, mean(d^2)} (kept for the backward); sums in double, fixed order.  The multi-scale loss around it
 * (:97-117) re-uses dvs_upsample_bilinear_* and dvs_smooth_loss_*. */
int dvs_silog_workspace_bytes(int64_t n, size_t* bytes);
int dvs_silog_fwd(const float* pred, const float* target, const uint8_t* valid, int64_t n, float variance_focus,
                  float* stats, void* workspace, void* stream);
int dvs_silog_bwd(const float* grad_out, const float* stats, const float* pred, const float* target, const uint8_t* valid,
                  int64_t n, float variance_focus, float* grad_pred, void* stream);

/* EvalTrajectory.depth_to_pointcloud (vo/eval_traj.py:85-128) / the point cloud of vo/predict.py:81-95: depth [H,W],
 * inv_K [4,4], T [4,4] (camera-to-world) -> points [H*W,3] = (T [depth * inv_K (u,v,1); 1])[:3]; valid [H*W] = depth > 0
 * (may be NULL). */
int dvs_depth_to_pointcloud(const float* depth, const float* inv_K, const float* T, float* points, uint8_t* valid,
                            int H, int W, void* stream);

/* Device side of MonoDataset.__getitem__ + the DataLoader's collation (vo/dataset/common.py:48-92) for frames already
 * resident on the device as decoded, resized uint8 RGB (what _read_image returns; decoding and resizing stay on the host):
 *   frames  [T,H,W,3] (frames_hwc != 0, PIL / OpenCV order) or [T,3,H,W]
 *   idx     [B,3] int32 device array: frame numbers of (source_left, target_image, source_right) of every sample
 *   out_*   [B,3,H,W] each, out_dtype DVS_DTYPE_F32 (ToTensor applied: x / 255, exact) or DVS_DTYPE_U8 (bytes kept; feed
 *           dvs_photometric_forward_ex with image_dtype = DVS_DTYPE_U8) */
int dvs_gather_triplets_u8(const uint8_t* frames, int frames_hwc, const int32_t* idx, void* out_left, void* out_target,
                           void* out_right, int out_dtype, int B, int H, int W, void* stream);

/* Network inputs of a training step in one pass: from the NCHW fp32 frames of the sample dict (vo/dataset/common.py:77) to
 * what the first convolutions of DepthNet / PoseNet read -- the channels-last target [B,H,W,3] and, per source k, the
 * concatenated pose pair [B,H,W,6] ([source, target] if bit k of src_first_mask is set, as for the frames before the target,
 * else [target, source]; vo/learner_new.py:110-123), optionally normalised as the encoders do, (x - 0.45) / 0.225
 * (model/resnet_encoder.py forward), in out_dtype (fp32, or bf16 = autocast's cast at conv1).  H * W must be a multiple of 8;
 * inputs 16-byte aligned.  Same arithmetic as the stock sequence on CUDA (fp32 subtract, times the fp32 reciprocal of 0.225 as
 * ATen's tensor / scalar does, one rounding to bf16). */
int dvs_pack_net_inputs(const float* target, const float* const* sources, int num_sources, unsigned int src_first_mask,
                        int normalize, void* out_target, void* const* out_pairs, int out_dtype, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVSLOSS_H_ */
