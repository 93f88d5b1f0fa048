/*
 * b200pose.h -- C ABI of libb200pose.so: the RNNPose recurrent pose-refinement inner loop as
 * hand-written sm_100a CUDA kernels.
 *
 * The reference (DecaYale/RNNPose) has no native interface on this path: every operation below is a
 * chain of PyTorch calls inside PoseRefiner.forward (reference model/PoseRefiner.py:315-365).  Its
 * only FFI precedent is thirdparty/nn/src/ext.h:1-10 (plain C function, raw pointers + ints).  This
 * header follows that shape and SURVEY.md section 8(b):
 *   - plain pointers and sizes only (no torch types), every pointer is a DEVICE pointer unless its
 *     name ends in _host;
 *   - every function returns int: 0 = ok, negative = invalid argument (B200POSE_E_*), positive = the
 *     cudaError_t of a failed launch / API call;  nothing exits, throws or allocates;
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*), no device-wide sync,
 *     graph-capturable (except where noted); scratch memory comes from the caller
 *     (b200pose_workspace_bytes);
 *   - process-wide state is limited to (1) the kernel-selection OPTIONS below (read once from B200POSE_* environment
 *     variables at the first call, then changed only by b200pose_set_option), (2) a bounded cache of encoded TMA
 *     descriptors keyed by (device address, extents, box) -- a descriptor depends on nothing else, so a hit is always
 *     valid -- and (3) per-device attributes queried once (SM count, shared-memory opt-in).  No data, no handles.
 *
 * Internal activation layout ("PXC"): pixel-major, channels contiguous: [B*h*w][C] float32, where
 * h = H/8, w = W/8 is the 1/8-resolution grid.  Boundary tensors keep the reference's NCHW layout.
 * The per-operator entry points use PXC for the low-resolution maps; the fused entry point
 * b200pose_refine_iters consumes exactly what the reference's inner loop consumes.
 */
#ifndef B200POSE_H
#define B200POSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200POSE_VERSION 3

/* error codes (negative); positive return values are cudaError_t */
#define B200POSE_OK            0
#define B200POSE_E_NULL       (-1)   /* a required pointer is NULL */
#define B200POSE_E_SHAPE      (-2)   /* unsupported shape (see each function) */
#define B200POSE_E_WORKSPACE  (-3)   /* workspace too small / misaligned */
#define B200POSE_E_ARG        (-4)   /* bad scalar argument */

#define B200POSE_CORR_LEVELS   4     /* reference model/CFNet.py:59 */
#define B200POSE_CORR_RADIUS   4     /* reference model/CFNet.py:60 */
#define B200POSE_CORR_CH       324   /* 4 * 9 * 9 lookup channels */
#define B200POSE_CORR_PITCH    328   /* PXC pitch of the lookup output (pad channels are zero) */
#define B200POSE_NUM_WEIGHT_TENSORS 30

/* `flags` bits of b200pose_update_block / b200pose_refine_iters[_host] */
#define B200POSE_FLAG_EXACT_FP32    0   /* CUDA-core FFMA convolutions, fp32 throughout (LM step fp64) */
#define B200POSE_FLAG_TENSOR_CORES  1   /* tcgen05 convolutions, fp16 hi/lo split operands (x ~= hi+lo, 3 MMAs,
                                           fp32 TMEM accumulate): fp32-level accuracy, see DESIGN.md section 5 */
#define B200POSE_FLAG_GEO2_CHANNELS_LAST 2 /* b200pose_refine_iters only: geofea2 is [B,H*W,32] (pixel-major, the layout
                                           b200pose_zoom_crop writes) instead of NCHW; needs C_geo == 32 and the
                                           foreground pipeline (options fg_list, fg_pipeline), else B200POSE_E_ARG */
#define B200POSE_FLAG_CONTEXT_TEXELS 4  /* b200pose_refine_iters only: `context` holds, instead of the [B,256,H,W] map, only the
                                           four texels F.interpolate(scale 1/8, align_corners=True) (CFNet.py:129) reads per
                                           low-resolution sample: [B,256,(H/8)*(W/8),4] = (v00, v01, v10, v11) with rows
                                           y0 = min(int(y*(H-1)/(H/8-1)), H-1), y1 = min(y0+1, H-1), columns likewise (float
                                           arithmetic).  1/16 of the bytes; what b200pose_refine_iters_host2 sends over PCIe */

int b200pose_version(void);
const char* b200pose_error_string(int code);

/* ---- options ------------------------------------------------------------------------------------
 * Integer options that select between kernel variants (all variants meet the parity bar; they exist for A/B timing and
 * for the tests, which run every variant).  Names (initial value <- environment variable, read once):
 *   conv_mode (B200POSE_CONV_MODE, 19) tensor-core convolution: bit0 CTA pairs, bit1 vertical-tap reuse, bit2 force the
 *                                      second-generation kernel, bit4 the eleven convolutions of an update-block pass in one
 *                                      persistent launch (machine-filling batches; else layer by layer); 0 = first generation
 *   fg_list (B200POSE_FG_LIST, 1)      LM steps over the per-call foreground list
 *   fg_pipeline (B200POSE_FG_PIPELINE, 2)  channels-last target / weight kernels + cluster LM kernel: 0 off, 1 on, 2 only when
 *                                      geofea2 arrives channels-last (B200POSE_FLAG_GEO2_CHANNELS_LAST)
 *   fg_upsample, fg_blocks, sparse_g1, sparse_g2, g2_margin, host_gather, host_gather_planes, upsample_variant, enc_stem, enc_chunk,
 *   tail_min_n, lookup_mode, pool_mode, pdl_off,
 *   chain_rings, chain_dynamic, chain_xmajor, chain_merge, lm_cluster, conv_debug, lm_debug: see csrc/options.cu
 * Returns 0, B200POSE_E_NULL or B200POSE_E_ARG (unknown name).  Not synchronised against launches in flight on other
 * threads.                                                                                         */
int b200pose_set_option(const char* name, int value);
int b200pose_get_option(const char* name, int* value);
int b200pose_option_count(void);
const char* b200pose_option_name(int index);

/* ---- weights ----------------------------------------------------------------------------------
 * Replaces: nn.Conv2d parameter storage of BasicUpdateBlock (reference thirdparty/raft/update.py:
 * 164-177), loaded from weights/gru_update.pth at model/CFNet.py:71-74.
 * `tensors` = 30 device pointers in state-dict order, reference layouts
 * ([Cout,Cin,KH,KW] weight, [Cout] bias):
 *   encoder.convc1, encoder.convc2, encoder.convf1, encoder.convf2, encoder.conv,
 *   gru.convz1, gru.convr1, gru.convq1, gru.convz2, gru.convr2, gru.convq2,
 *   flow_head.conv1, flow_head.conv2, mask.0, mask.2          (each: weight then bias)
 * `packed` = caller-owned device buffer of b200pose_packed_weights_bytes() bytes (256-B aligned)
 * that receives the kernel-side layouts (fp32 GEMM layouts + fp16 hi/lo planes); it stays valid as long as the caller keeps it.          */
size_t b200pose_packed_weights_bytes(void);
int b200pose_pack_weights(const float* const* tensors_host /* host array of 30 device ptrs */,
                          void* packed, void* stream);

/* ---- a1: correlation volume + pyramid ---------------------------------------------------------
 * Replaces CorrBlock.__init__ / CorrBlock.corr (reference thirdparty/raft/corr.py:13-34,60-67).
 * fmap1,fmap2: [B,D,h,w] NCHW.  pyramid: b200pose_pyramid_floats(B,h,w) floats; level l holds
 * [B][h*w][hl*wl] with hl = h>>l, wl = w>>l (floor), levels stored back to back.
 * Requires h>>3 >= 2 and w>>3 >= 2 (the reference divides by (wl-1) in its sampler).             */
size_t b200pose_pyramid_floats(int B, int h, int w);
int b200pose_corr_pyramid(const float* fmap1, const float* fmap2, int B, int D, int h, int w,
                          float* pyramid, void* stream);
/* Tensor-core variant (what the fused loop uses with B200POSE_FLAG_TENSOR_CORES): D = 256 only, h*w must have a
 * divisor n <= 240 with n % 16 == 0; the volume GEMM runs on tcgen05 with fp16 hi/lo split operands.          */
size_t b200pose_corr_pyramid_tc_workspace_bytes(int B, int h, int w);
int b200pose_corr_pyramid_tc(const float* fmap1, const float* fmap2, int B, int h, int w,
                             float* pyramid, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2: pyramid lookup -----------------------------------------------------------------------
 * Replaces CorrBlock.__call__ + bilinear_sampler (corr.py:36-57, utils/utils.py:57-71).
 * coords: [B*h*w][2] (x,y) PXC.  out: [B*h*w][328] PXC; channel l*81 + i*9 + j samples
 * (x/2^l + i-4, y/2^l + j-4); channels 324..327 are written as 0.                               */
int b200pose_corr_lookup(const float* pyramid, const float* coords, int B, int h, int w,
                         float* out, void* stream);

/* ---- a6 (part): hidden state / context init ---------------------------------------------------
 * Replaces the 1/8 bilinear (align_corners=True) resample + tanh/relu split, model/CFNet.py:124-133.
 * context: [B,256,H,W] NCHW (already multiplied by 0.1, PoseRefiner.py:283).
 * net: [B*h*w][128] = tanh(ctx[:128]);  xbuf: [B*h*w][256], channels 0..127 = relu(ctx[128:]).   */
int b200pose_context_init(const float* context, int B, int H, int W, float* net, float* xbuf, void* stream);

/* ---- a8 + a6: reprojection flow-init at 1/8 resolution ----------------------------------------
 * Replaces SE3.transform + flow_init (PoseRefiner.py:324-328, transformation.py:184-198) and the
 * "/8 then 1/8 bilinear align_corners=True" of model/CFNet.py:138-144.
 * depth: [B,H,W] (syn_depth, WITHOUT the +1e-5; added inside), K: [B,3,3], G: [B,4,4] row-major.
 * coords1: [B*h*w][2] = coords0 + resampled flow_init/8;  flow: [B*h*w][2] = coords1 - coords0.  */
int b200pose_flow_init(const float* depth, const float* K, const float* G, int B, int H, int W,
                       float* coords1, float* flow, void* stream);

/* ---- a3-a5: update block ----------------------------------------------------------------------
 * Replaces BasicUpdateBlock.forward (thirdparty/raft/update.py:179-188).
 * net: [P][128] in/out (hidden state);  xbuf: [P][256], channels 0..127 = inp (input), channels
 * 128..255 are scratch (motion features);  corr: [P][328];  coords1: [P][2] in/out
 * (coords1 += delta_flow, CFNet.py:157);  flow: [P][2] in = coords1-coords0 (CFNet.py:151),
 * out = new coords1-coords0 (CFNet.py:166);  mask: [P][576] out (already x0.25).
 * workspace: b200pose_update_workspace_bytes(B,h,w).                                             */
size_t b200pose_update_workspace_bytes(int B, int h, int w);
int b200pose_update_block(const void* packed_weights, float* net, float* xbuf, const float* corr,
                          float* coords1, float* flow, float* mask, float* dflow_out /* [P][2] or NULL */,
                          int B, int h, int w, int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3-a5 (single layer; test / micro-benchmark entry) ---------------------------------------
 * One convolution of the update block (nn.Conv2d of thirdparty/raft/update.py) WITHOUT activation:
 * out[P][pitch] = conv(in) + bias, pitch = roundup(cout,4).  layer: 0 convc1, 1 convc2, 2 convf1 (as a
 * 1x1 over the 98-wide im2col), 3 convf2, 4 conv, 5 convz1|convr1, 6 convq1, 7 convz2|convr2, 8 convq2,
 * 9 flow_head.conv1|mask.0, 10 mask.2.  in0/in1: PXC fp32 inputs (in1 only for the GRU layers: in0 = h or
 * r*h [P][128], in1 = x [P][256]).  b200pose_conv_layer_info reports the channel counts.            */
size_t b200pose_conv_layer_workspace_bytes(int B, int h, int w);
int b200pose_conv_layer_info(int layer, int* cin0, int* cin1, int* cout, int* kh, int* kw);
int b200pose_conv_layer(const void* packed_weights, int layer, const float* in0, int pitch0,
                        const float* in1, int pitch1, float* out, int B, int h, int w, int flags,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- a7 + a9: convex upsampling fused with the correspondence weight --------------------------
 * Replaces GRU_CFUpdator.upsample_flow (model/CFNet.py:95-106), target = flow + grid
 * (PoseRefiner.py:335-338) and the descriptor-similarity weight (PoseRefiner.py:342-345).
 * flow: [P][2] low-res flow, mask: [P][576]; geofea1, geofea2: [B,C,H,W] NCHW (may both be NULL
 * together with weight: then only flow_up/target are produced); depth: [B,H,W] (syn_depth).
 * Outputs (each may be NULL): flow_up [B,2,H,W] NCHW; target [B,H,W,2]; weight [B,H,W].          */
int b200pose_upsample_weight(const float* flow, const float* mask, const float* geofea1,
                             const float* geofea2, const float* depth, float sigma,
                             int B, int C, int H, int W,
                             float* flow_up, float* target, float* weight, void* stream);

/* ---- a10-a12: Levenberg-Marquardt steps --------------------------------------------------------
 * Replaces SE3Sequence.reprojction_optim (geometry/transformation.py:265-316): residual/Jacobian,
 * fp64 H = sum v w J^T J, b = sum v w J^T r, damping H += ep*I + lm*diag(H), geometry/cholesky.py
 * solve (NaN -> 0, clamp +-1), se3 exponential retraction G <- exp(delta) G (geometry/se3.py).
 * depth: [B,H,W]; depth_offset is added to it before use: pass 1e-5f with the raw syn_depth, 0 with the
 * reference's `depths = syn_depth + EPS` (PoseRefiner.py:313); target: [B,H,W,2]; weight: [B,H,W]; K: [B,3,3];
 * G: [B,4,4] updated IN PLACE n_steps times.  Optional taps (may be NULL), each written per step:
 * H_out [n_steps][B][36] fp64 (un-damped), b_out [n_steps][B][6] fp64, delta_out [n_steps][B][6].
 * workspace: b200pose_lm_workspace_bytes(B,H,W).                                                 */
size_t b200pose_lm_workspace_bytes(int B, int H, int W);
int b200pose_lm_solve(const float* depth, const float* target, const float* weight, const float* K,
                      float* G, int B, int H, int W, float depth_offset, int n_steps, double ep_lmbda, double lm_lmbda,
                      double* H_out, double* b_out, float* delta_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- f4: backward of one LM step ------------------------------------------------------------------
 * What autograd computes in the reference for SE3Sequence.reprojction_optim(num_iters = 1) (geometry/transformation.py:
 * 265-316; the shipped OPTIM_ITER_COUNT is 1) through the custom Cholesky backward of geometry/cholesky.py:19-28: given
 * grad_delta [B,6] = dL/d(delta) of the step's clamped fp32 update, the gradients with respect to target [B,H,W,2] and
 * weight [B,H,W].  G [B,4,4] is the pose ENTERING the step (a constant of the step: PoseRefiner.py:320 detaches it);
 * depth / depth_offset / K / ep_lmbda / lm_lmbda as in b200pose_lm_solve.  The gradient through NaN -> 0 and the +-1 clamp is
 * zero for the affected components, as torch.where / torch.clamp.  workspace: b200pose_lm_backward_workspace_bytes.      */
size_t b200pose_lm_backward_workspace_bytes(int B, int H, int W);
int b200pose_lm_backward(const float* depth, const float* target, const float* weight, const float* K, const float* G,
                         const float* grad_delta, int B, int H, int W, float depth_offset, double ep_lmbda, double lm_lmbda,
                         float* grad_target, float* grad_weight, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11, a12 in isolation (the LM kernels above run them inline; these entries let a caller -- and the tests -- reach
 * the two small pieces directly) -------------------------------------------------------------------
 * b200pose_cholesky_solve: geometry/cholesky.py:32-50 `solve` for 6x6 systems: x = H^-1 b by Cholesky in fp64, NaN -> 0,
 *   clamp to [-1, 1], cast to fp32.  H [B,6,6] fp64 symmetric (no damping is added), b [B,6] fp64, x [B,6] fp32.
 * b200pose_se3_retract: SE3.increment (geometry/transformation.py:110-115): G <- exp(delta) G with the fp32 exponential of
 *   geometry/se3.py:228-306 (Taylor branch below theta = 1e-4); delta [B,6] = (upsilon, omega), G [B,4,4] in place.   */
int b200pose_cholesky_solve(const double* H, const double* b, float* x, int B, void* stream);
int b200pose_se3_retract(const float* delta, float* G, int B, void* stream);

/* ---- f3: per-object pose metrics ----------------------------------------------------------------
 * Replaces the evaluator's per-object arithmetic (utils/eval_metric.py:306-339 evaluate_rnnpose): add_metric /
 * add2_metric / add5_metric (:120-179) with the brute-force nearest neighbour of thirdparty/nn (nn_utils.py:6-22,
 * src/nearest_neighborhood.cu:48-80) for symmetric objects, projection_2d (:102-110, :23-35), cm_degree_5_metric
 * (:181-192) and utils/geometric.py:36-40 (rotation_angle).  T_pred,T_gt: [B,4,4] row-major; pts: [B,n_pts,3] model
 * points; diameter: [B] (same unit as pts); K: [B,3,3] projection intrinsics (the reference uses linemod_K).
 * out: [B,B200POSE_METRIC_COLS] =
 *   0 ADD            mean_x |T_pred x - T_gt x|
 *   1 ADD-S          mean over the GROUND-TRUTH points of the distance to the nearest PREDICTED point (the reference's
 *                    direction: find_nearest_point_idx(model_pred, model_targets), eval_metric.py:167-171)
 *   2 rotation error, chordal form 2 asin(|R_gt - R_pred|_F / sqrt 8), degrees     3 translation error |t_pred - t_gt|
 *   4 mean 2-D projection error (pixels)     5 rotation error from the trace, acos((tr(R_pred R_gt^T) - 1) / 2), degrees
 *   6 ADD < 0.1 d   7 ADD-S < 0.1 d   8 ADD < 0.02 d   9 ADD-S < 0.02 d   10 ADD < 0.05 d   11 ADD-S < 0.05 d
 *   12 projection error < 5 px     13 translation * 100 < 5 and trace angle < 5 (5 cm 5 deg, pts in metres)
 *   14 0 (slot for the caller's object index)     15 d
 * workspace: b200pose_pose_metrics_workspace_bytes.                                                              */
#define B200POSE_METRIC_COLS 16
size_t b200pose_pose_metrics_workspace_bytes(int B, int n_pts);
int b200pose_pose_metrics(const float* T_pred, const float* T_gt, const float* pts, const float* diameter,
                          const float* K, int B, int n_pts, float* out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- f1: online zoom-crop -------------------------------------------------------------------------
 * Replaces, per render iteration, PoseRefiner.gen_zoom_crop_grids / get_affine_transformation (model/PoseRefiner.py:145-218:
 * numpy nonzero + cv2.getAffineTransform on the host, per sample) and the two F.grid_sample crops of :287 and :292.
 * pc_depth [B,H,W]: rendered point-cloud depth (foreground = depth > 0, :259-263); K [B,3,3] full-image intrinsics;
 * T [B,4,4] current pose (the crop is centred on the projection of its translation, :204-205); image [B,Ci,H,W] and
 * geofea [B,Cg,H,W] are the maps to crop (either may be NULL with its output).  margin_ratio: 0.4 in the reference.
 * Outputs: image_crop [B,Ci,Hc,Wc]; geofea_crop [B,Cg,Hc,Wc], or with B200POSE_ZOOM_GEO_CHANNELS_LAST in `flags` (needs
 * Cg == 32) [B,Hc*Wc,32] = what b200pose_refine_iters takes with B200POSE_FLAG_GEO2_CHANNELS_LAST; K_crop [B,3,3] (:211);
 * theta [B,2,3] (optional, the F.affine_grid matrix).  Bilinear, zeros outside, align_corners=False (torch defaults, as the
 * reference).  An empty foreground yields the reference's zero box.  workspace: b200pose_zoom_crop_workspace_bytes(B).   */
#define B200POSE_ZOOM_GEO_CHANNELS_LAST 1
size_t b200pose_zoom_crop_workspace_bytes(int B);
int b200pose_zoom_crop(const float* pc_depth, const float* K, const float* T, const float* image, const float* geofea,
                       int B, int Ci, int Cg, int H, int W, int Hc, int Wc, float margin_ratio, int flags,
                       float* image_crop, float* geofea_crop, float* K_crop, float* theta,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- f2: RAFT feature encoder ---------------------------------------------------------------------
 * Replaces ImageFeaEncoder.forward (model/CFNet.py:26-49) = BasicEncoder(output_dim 256, norm_fn 'instance') of
 * thirdparty/raft/extractor.py:118-232 on both images at once: 7x7 s2 stem, six residual blocks with InstanceNorm + ReLU
 * (two of them stride 2 with a 1x1 down-sampling branch), 1x1 output convolution; inputs are normalised as
 * 2 (x / 255) - 1 inside (CFNet.py:42-43).  fp32-equivalent arithmetic (the stem in fp32 FFMA, the other fifteen
 * convolutions on tcgen05 with fp16 hi/lo split operands).
 * `tensors` = 32 device pointers in state-dict order of weights/img_fea_enc.pth (fnet.conv1, fnet.layer1.0.conv1,
 * .conv2, fnet.layer1.1.*, fnet.layer2.0.conv1, .conv2, .downsample.0, fnet.layer2.1.*, fnet.layer3.0.*, fnet.layer3.1.*,
 * fnet.conv2; each weight then bias).  image1, image2: [B,3,H,W] (H, W multiples of 8, >= 16); fmap1, fmap2: [B,256,H/8,W/8].  */
#define B200POSE_NUM_ENCODER_TENSORS 32
size_t b200pose_encoder_packed_weights_bytes(void);
int b200pose_encoder_pack_weights(const float* const* tensors_host, void* packed, void* stream);
size_t b200pose_encoder_workspace_bytes(int B, int H, int W);
int b200pose_image_encoder(const void* packed_weights, const float* image1, const float* image2, int B, int H, int W,
                           float* fmap1, float* fmap2, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a14: the fused inner loop ----------------------------------------------------------------
 * Replaces the body of `for i in range(cfg.ITER_COUNT)` in PoseRefiner.forward
 * (model/PoseRefiner.py:315-362) for one render iteration, natively batched.
 * Inputs (device, fp32): fmap1,fmap2 [B,256,h,w]; context [B,256,H,W] (x0.1 applied);
 * geofea1,geofea2 [B,C,H,W]; depth [B,H,W]; K [B,3,3]; G [B,4,4] = Tij at loop entry, overwritten
 * with Tij at loop exit.  Optional outputs: flow_first [B,2,H,W] (full-res flow of recurrent
 * iteration 0, the "flow" entry of the reference's return dict), flow_last [B,2,H,W],
 * weight_last [B,H,W].  n_iters = ITER_COUNT, n_lm = OPTIM_ITER_COUNT.                            */
size_t b200pose_refine_workspace_bytes(int B, int H, int W);
int b200pose_refine_iters(const void* packed_weights,
                          const float* fmap1, const float* fmap2, const float* context,
                          const float* geofea1, const float* geofea2, const float* depth,
                          const float* K, float* G, float sigma,
                          int B, int C_geo, int H, int W, int n_iters, int n_lm,
                          double ep_lmbda, double lm_lmbda, int flags,
                          float* flow_first, float* flow_last, float* weight_last,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (pinned or pageable): uploads the inputs, runs the loop, downloads G.
 * This is the end-to-end entry a caller without device-resident tensors uses (bench.py "e2e").
 * device_scratch must hold b200pose_refine_host_scratch_bytes(B,C_geo,H,W) bytes.                 */
size_t b200pose_refine_host_scratch_bytes(int B, int C_geo, int H, int W);
int b200pose_refine_iters_host(const void* packed_weights,
                               const float* fmap1_host, const float* fmap2_host, const float* context_host,
                               const float* geofea1_host, const float* geofea2_host, const float* depth_host,
                               const float* K_host, float* G_host, float sigma,
                               int B, int C_geo, int H, int W, int n_iters, int n_lm,
                               double ep_lmbda, double lm_lmbda, int flags,
                               void* device_scratch, size_t device_scratch_bytes, void* stream);

/* The same, with the context map gathered on the host: n_host_threads worker threads (<= 0: half of the CPUs the process
 * may run on, at most 8; they live for the call) copy the texels the 1/8 resample reads (B200POSE_FLAG_CONTEXT_TEXELS layout,
 * 1/16 of the map) into host_staging, one sub-batch ahead of the transfers, and only those cross PCIe.  host_staging: 16-byte
 * aligned, b200pose_refine_host_staging_bytes(B,H,W) bytes, pinned for full copy speed; NULL = exactly
 * b200pose_refine_iters_host.  Results are bit-identical to the device-buffer entry.                  */
size_t b200pose_refine_host_staging_bytes(int B, int H, int W);
/* Host-only helper (no CUDA call): the gather by itself, context_host [B,256,H,W] -> texels_host [B,256,(H/8)*(W/8),4]
 * (16-byte aligned), for callers that upload the texels themselves and pass B200POSE_FLAG_CONTEXT_TEXELS. */
int b200pose_context_gather_texels(const float* context_host, int B, int H, int W, float* texels_host, int n_host_threads);
int b200pose_refine_iters_host2(const void* packed_weights,
                                const float* fmap1_host, const float* fmap2_host, const float* context_host,
                                const float* geofea1_host, const float* geofea2_host, const float* depth_host,
                                const float* K_host, float* G_host, float sigma,
                                int B, int C_geo, int H, int W, int n_iters, int n_lm,
                                double ep_lmbda, double lm_lmbda, int flags,
                                void* device_scratch, size_t device_scratch_bytes,
                                void* host_staging, size_t host_staging_bytes, int n_host_threads, void* stream);

/* Measurement hook (bench.py "roofline"): the NEXT tensor-core update-block pass (b200pose_update_block or an iteration of
 * b200pose_refine_iters) records the two cudaEvent_t on its stream immediately before and after its convolution launch(es)
 * -- the chained launch, or the eleven layer launches -- and forgets them.  Not thread-safe.                        */
int b200pose_debug_set_conv_events(void* ev_start, void* ev_stop);

/* number of kernel launches b200pose_refine_iters enqueues for this shape with the tensor-core flag, C_geo = 32 and the
 * current options (for bench.py's gpu_launches); 0 for an unsupported shape */
int b200pose_refine_launch_count(int B, int H, int W, int n_iters, int n_lm);

#ifdef __cplusplus
}
#endif
#endif /* B200POSE_H */
