/*
 * dff_b200.h — C-ABI of the B200-native (sm_100a) depth-from-focus hot path.
 *
 * This is the drop-in boundary: plain C types only (pointers, sizes, ints), no torch / ATen types.  Every entry
 * point replaces a piece of the reference's Python hot path (paths relative to the reference repository,
 * wcy199705/DfFintheWild):
 *
 *   dff_forward            <- DFF_net.forward        train_codes/Depth_Estimation_Network.py:77-137
 *                                                    Depth_Estimation_Test/Depth_Estimation_Network.py:74-127
 *   dff_forward_host       <- the eval loop's `.cuda()` uploads + forward + `.cpu()` reads, Depth_Estimation_Test/test.py:115-121
 *   dff_pack_weights       <- nn.Module parameter/buffer storage read by every conv/BN call
 *                             (convbn_3d, train_codes/Depth_Estimation_Network.py:352-355)
 *   dff_conv3d             <- one nn.Conv3d / nn.ConvTranspose3d (+BatchNorm3d eval +ReLU +residual) call site,
 *                             e.g. train_codes/Depth_Estimation_Network.py:144-148, 278-301
 *   dff_srd_attention      <- the channel-attention branch of Feature_Extraction / SRD, train_codes/Depth_Estimation_Network.py:399-407
 *   dff_depth_head         <- upsample + softplus-normalise + expected focus distance,
 *                             train_codes/Depth_Estimation_Network.py:92-98, 118-136
 *   dff_conv3d_dgrad, dff_conv3d_wgrad, dff_bn_train_forward/backward, dff_pool3d(_backward), dff_add,
 *   dff_depth_head_backward <- autograd of DFF_net.forward in train mode (invoked at train_codes/train_code_Defocus.py:167;
 *                             BatchNorm3d with batch statistics, train_codes/Depth_Estimation_Network.py:355); PyTorch's autograd
 *                             stays the tape (dffinthewild_b200/train.py), every computation is one of these calls
 *   dff_fov_warp(_cl)      <- FlowNetwork.FOV_warp    End_to_End/End_to_End.py:106-134
 *   dff_pair_volume, dff_spatial_mean_accum (+ dff_conv3d)
 *                          <- FlowNetwork.forward     End_to_End/End_to_End.py:63-104 (composed in dffinthewild_b200/End_to_End.py)
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.  The library never allocates or frees
 *     device memory and keeps no pointer past the call: inputs, outputs, packed weights and workspace belong to
 *     the caller (PyTorch tensors in the Python binding).
 *   - Every launch goes to `stream` (a cudaStream_t passed as void*) on CUDA device `device`.
 *   - Return value: 0 on success, a negative DFF_E_* code on failure; `dff_last_error` returns the thread-local
 *     message.  There is no CPU fallback: an unsupported shape, precision or device is an error.
 *   - Thread-safe / re-entrant: no mutable global state (nn.DataParallel may call from one thread per GPU).
 *   - Tensors at the boundary use the reference's layouts: FS (B,3,S,H,W) fp32 contiguous; focus_dists addressed
 *     through four element strides (0 = broadcast) so (B,S,H,W), (B,S,1,1) ... are accepted; outputs (B,H,W) fp32.
 *     Internally activations are channels-last (B,S,H,W,C).
 */
#ifndef DFF_B200_H_
#define DFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFF_ABI_VERSION 1

/* error codes */
#define DFF_OK 0
#define DFF_E_ARG (-1)       /* bad argument / unsupported shape (H or W not a multiple of 32, S < 1 ...) */
#define DFF_E_WORKSPACE (-2) /* workspace too small */
#define DFF_E_CUDA (-3)      /* CUDA runtime / launch error */
#define DFF_E_DEVICE (-4)    /* not an sm_100 device */
#define DFF_E_UNSUPPORTED (-5)

/* precision / mode flags (bit-or) */
#define DFF_FP32 0  /* fp32 activations, FFMA kernels: parity mode (<= 1e-4 relative per pixel) */
#define DFF_BF16 1  /* bf16 activations, tcgen05/TMEM implicit-GEMM kernels, fp32 accumulate */
#define DFF_TRAIN 2 /* BatchNorm in batch-statistics mode; reserved: train mode runs through the operator calls below */
#define DFF_NO_TC 4 /* debugging aid with DFF_BF16: bf16 storage but FFMA kernels instead of tcgen05 */
#define DFF_NO_SLAB 8 /* debugging aid with DFF_BF16: only the per-tap TMA tcgen05 kernel, never the slab kernel */
#define DFF_OUT_F32 16 /* with DFF_BF16 as the `elem` of dff_conv3d: store the output as fp32 (C -> 1 cost volumes) */

/* which parameter set a call refers to */
#define DFF_NET_DFF 0  /* DFF_net (384-key state_dict)                                */
#define DFF_NET_FLOW 1 /* End_to_End FlowNetwork (`optical_flow_aggregation.*` keys) */

/* ---- introspection / handshake ---------------------------------------------------------------------------- */
int dff_abi_version(void);
/* message of the last failure on this thread */
const char *dff_last_error(void);
/* 0 if `device` is a usable sm_100 GPU */
int dff_check_device(int device);

/* The library declares the parameters it expects, by reference state_dict key (without the `DFF_net.` /
 * `optical_flow_aggregation.` prefix).  The caller concatenates them as fp32 in this order into one flat device
 * buffer ("raw parameters"); `num_batches_tracked` is not part of it. */
int dff_param_count(int net);
const char *dff_param_name(int net, int index);
int64_t dff_param_numel(int net, int index);
int64_t dff_param_offset(int net, int index); /* element offset inside the raw buffer */
int64_t dff_raw_numel(int net);               /* total fp32 elements of the raw buffer */

/* ---- weights ---------------------------------------------------------------------------------------------- */
/* bytes of the packed-weight buffer for `net` */
size_t dff_packed_bytes(int net);
/* raw fp32 parameters -> packed kernels layouts ([tap][cin][cout] fp32, [tap][cout][cin] bf16) and, for eval,
 * per-channel BatchNorm scale/shift (gamma/sqrt(var+eps), beta-mean*scale).  Runs on `stream`. */
int dff_pack_weights(int net, const float *raw, void *packed, int device, void *stream);

/* ---- whole-network forward -------------------------------------------------------------------------------- */
size_t dff_workspace_bytes(int B, int S, int H, int W, int mode);
/* FS (B,3,S,H,W) fp32; fd + 4 element strides; out4[0..3] = mid_out, pred1, pred2, pred3, each (B,H,W) fp32.
 * cost4 (optional, may be NULL): pre-softplus costs (B,S,H/8,W/8), (B,S,H/4,W/4), (B,S,H/2,W/2), (B,S,H,W). */
int dff_forward(const void *packed, const float *FS, const float *fd, const int64_t fd_strides[4], int B, int S,
                int H, int W, float *const out4[4], float *const cost4[4], void *workspace, size_t workspace_bytes,
                int mode, int device, void *stream);

/* Profiled variant (bench / roofline): CUDA events around every operator on `stream`, then a stream synchronise.
 * Reports, per operator in schedule order: device ms, algorithmic FLOPs (2*MACs, no padding or zero-tap MACs),
 * compulsory bytes, kernel launches and a 64-byte name slot.  With packed == NULL nothing runs and only the plan
 * (names / flops / bytes / launches) is reported. */
int dff_forward_profiled(const void *packed, const float *FS, const float *fd, const int64_t fd_strides[4], int B, int S,
                         int H, int W, float *const out4[4], void *workspace, size_t workspace_bytes, int mode, int device,
                         void *stream, int max_ops, float *op_ms_host, double *op_flops_host, double *op_bytes_host,
                         int *op_launches_host, char *op_names_host, int *n_ops);

/* Same computation with HOST buffers (pinned for full speed; pageable works): B stacks are processed in micro-batches of
 * `micro_batch` through a three-stage pipeline — the host->device copies of chunks i+1, i+2 and the device->host reads of chunk i-1
 * overlap chunk i's kernels (two private copy streams per calling thread and device).  Returns after every map is in host
 * memory.  FS_host (B,3,S,H,W); fd_host + strides as in dff_forward; out4_host[j] (B,H,W) or NULL.  `dev_io` is caller-owned
 * device scratch of dff_host_io_bytes(micro_batch, ...) bytes; `workspace` >= dff_workspace_bytes(micro_batch, ...).
 * This is the call behind `model(FS.cuda(), fd.cuda())[3].cpu()` of Depth_Estimation_Test/test.py:115-121. */
size_t dff_host_io_bytes(int micro_batch, int S, int H, int W);
int dff_forward_host(const void *packed, const float *FS_host, const float *fd_host, const int64_t fd_strides[4], int B,
                     int micro_batch, int S, int H, int W, float *const out4_host[4], void *dev_io, void *workspace,
                     size_t workspace_bytes, int mode, int device, void *stream);

/* ---- uint8 input staging (reference Depth_Estimation_Test/test_Dataloader.py:122-147, 105-113) ---------------
 * The datasets store focal stacks as uint8 (S, H0, W0, 3); the reference's dataloader turns them into fp32 `x/127.5 - 1.0`, pads
 * H, W to multiples of 32 with -1, transposes to (3, S, H, W) and tiles the S focus distances to (S, H, W).  These entry points
 * take the stacks as stored — FS_u8 (B, S, H0, W0, 3) uint8, H0 <= H, W0 <= W with H, W the padded extent — and do the
 * normalisation (IEEE fp32, bit-identical to numpy's), the padding and the layout change in the first kernel of the forward;
 * focus_dists may be the S scalars per stack (strides {S, 1, 0, 0}).  Results are bit-identical to dff_forward on the fp32 tensor
 * the dataloader would have produced.  A quarter of the bytes cross PCIe (6.3 instead of 35.4 MB per DDFF stack). */
int dff_forward_u8(const void *packed, const uint8_t *FS_u8, int H0, int W0, const float *fd, const int64_t fd_strides[4], int B,
                   int S, int H, int W, float *const out4[4], float *const cost4[4], void *workspace, size_t workspace_bytes,
                   int mode, int device, void *stream);
/* host-buffer variant: same pipeline as dff_forward_host; dev_io >= dff_host_io_bytes_u8(...) */
size_t dff_host_io_bytes_u8(int micro_batch, int S, int H0, int W0, int H, int W, const int64_t fd_strides[4]);
int dff_forward_host_u8(const void *packed, const uint8_t *FS_u8_host, int H0, int W0, const float *fd_host,
                        const int64_t fd_strides[4], int B, int micro_batch, int S, int H, int W, float *const out4_host[4],
                        void *dev_io, void *workspace, size_t workspace_bytes, int mode, int device, void *stream);
/* Double-buffered variant for a dataloader loop that prefetches: the call only QUEUES the uploads, kernels and reads of its B
 * stacks and returns; dff_forward_host_wait(device, ticket) blocks until the maps of the call that used `ticket` (0 or 1) are in
 * host memory.  Two calls may be in flight per calling thread and device when they use different tickets AND different `dev_io`,
 * `workspace` and output buffers: call i+1's uploads then run during call i's kernels and its first kernels fill call i's drain
 * (a synchronous call exposes its first chunk's upload and its last chunk's read: ~6 % of a 64-stack call).  Usage:
 *   async(batch 0, ticket 0); for i = 1..: async(batch i, ticket i & 1); wait(ticket (i-1) & 1); consume batch i-1. */
int dff_forward_host_u8_async(const void *packed, const uint8_t *FS_u8_host, int H0, int W0, const float *fd_host,
                              const int64_t fd_strides[4], int B, int micro_batch, int S, int H, int W,
                              float *const out4_host[4], void *dev_io, void *workspace, size_t workspace_bytes, int mode,
                              int device, void *stream, int ticket);
int dff_forward_host_wait(int device, int ticket);
/* the dataloader tail alone: FS_u8 (B,S,H0,W0,3) -> FS (B,3,S,H,W) fp32, normalised and -1 padded (the tensor the reference feeds) */
int dff_stage_u8(const uint8_t *FS_u8, int H0, int W0, int B, int S, int H, int W, float *FS, int device, void *stream);

/* ---- single operators (unit-parity surface; also what dff_forward is made of) ----------------------------- */
/* Generic 3-D convolution on channels-last activations.
 *   in0 (B,S,IH,IW,C0) [+ in1 (B,S,IH,IW,C1): virtual channel concat]  ->  out (B,S,OH,OW,Cout)
 *   weight: reference layout, fp32: conv (Cout,Cin,kd,kh,kw); transposed (Cin,Cout,3,3,3) with stride (1,2,2).
 *   y = acc*scale[c] + shift[c] (+ res_pre) ; ReLU if relu ; (+ res_post).  scale/shift may be NULL (1 / 0).
 *   elem: DFF_FP32 or DFF_BF16 storage of in/out/res tensors.  `scratch` >= dff_conv3d_scratch_bytes(). */
size_t dff_conv3d_scratch_bytes(int Cin, int Cout, int kd, int kh, int kw);
int dff_conv3d(const void *in0, int C0, const void *in1, int C1, int B, int S, int IH, int IW, const float *weight,
               int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, const float *scale,
               const float *shift, const void *res_pre, const void *res_post, int relu, void *out, int elem,
               int use_tensor_cores, void *scratch, int device, void *stream);

/* The same operator with the forward's own kernel selection and fusions, so that every form dff_forward can pick is testable in
 * isolation.  `plan`: 0 FFMA; 1 best tensor-core kernel for the plain form; 2 per-tap TMA kernel; 3 slab kernel; 4 = exactly what
 * dff_forward runs for a layer of this shape — x-folded small-Cout layers, x-folded transposed convolutions, the row-folded
 * pair-packed first layer (`pair_input`: in0 is the (B,S,H,W+2,8) bf16 tensor of dff_to_pair_packed and `weight` the (Cout,3,1,9,9)
 * layer), a second output `aux_out = out + aux_add` (the `x + out` of train_codes/Depth_Estimation_Network.py:104,110) and a fused
 * 1x1x1 classifier `proj_out[pixel] = sum_c proj_w[c] * bf16(v[c])` (reference :53-57,105,111,116) on the stored value
 * (proj_on_aux 0) or on the second output (1); skip_out: `out` itself is not written.  aux / proj / pair_input need plan 4. */
int dff_conv3d_ex(const void *in0, int C0, const void *in1, int C1, int B, int S, int IH, int IW, const float *weight, int Cout,
                  int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, const float *scale, const float *shift,
                  const void *res_pre, const void *res_post, int relu, void *out, int elem, int plan, int pair_input,
                  const void *aux_add, void *aux_out, const float *proj_w, float *proj_out, int proj_on_aux, int skip_out,
                  void *scratch, int device, void *stream);
/* FS (B,3,S,H,W) fp32 -> the first layer's tensor-core input (B,S,H,W+2,8) bf16: column c = [RGB(c-2) | RGB(c) | 0 0] */
int dff_to_pair_packed(const float *FS, int B, int S, int H, int W, void *dst, int device, void *stream);

/* SRD / Feature_Extraction channel-attention branch (reference train_codes/Depth_Estimation_Network.py:399-407), one pass:
 *   out = F + relu( conv1x1x1( relu( conv3x1x1(F; w_a) ); w_b ) )      (no BatchNorm, no bias)
 * F, out: channels-last (B,S,H,W,C) bf16, C in {8,16,32}, H*W % 16 == 0; w_a (C,C,3,1,1), w_b (C,C,1,1,1) fp32 in the
 * reference layout; scratch >= 4*C*C floats.  The intermediate is rounded to bf16 (as every activation of the bf16 mode). */
int dff_srd_attention(const void *F, int B, int S, int H, int W, int C, const float *w_a, const float *w_b, void *out,
                      void *scratch, int device, void *stream);

/* cost (B,S,h,w) fp32 with H % h == 0 -> depth (B,H,W) fp32 */
int dff_depth_head(const float *cost, int h, int w, const float *fd, const int64_t fd_strides[4], int B, int S, int H,
                   int W, float *depth, int device, void *stream);

/* The four heads of DFF_net in one launch (what dff_forward runs last): cost4 = pre-softplus costs at 1/8, 1/4, 1/2, 1/1 resolution,
 * each (B,S,h,w) fp32; depth4 = mid_out, pred1, pred2, pred3, each (B,H,W).  fast != 0: the bf16 mode's kernel (SFU exp/log). */
int dff_depth_heads4(const float *const cost4[4], const float *fd, const int64_t fd_strides[4], int B, int S, int H, int W,
                     float *const depth4[4], int fast, int device, void *stream);

/* ---- evaluation tail on the device ------------------------------------------------------------------------------
 * The masked depth metrics of metrics.py:90-133 for B maps at once: est (B,H,W) as the forward returns it (padded), metrics over the
 * [:Hc,:Wc] crop (test.py:125) against gt (B,Hc,Wc), mask (B,Hc,Wc) bytes or NULL (all valid), conf (B,Hc,Wc) or NULL.
 * out12 (B,12) fp32: abs_rel, sq_rel, mse, mae, rmse, rmse_log, accuracy_1, accuracy_2, accuracy_3, mse_w_conf, mae_w_conf, valid count.
 * fp64 accumulation, fixed reduction order.  scratch >= dff_depth_metrics_scratch_bytes(B). */
size_t dff_depth_metrics_scratch_bytes(int B);
int dff_depth_metrics(const float *est, const float *gt, const uint8_t *mask, const float *conf, int B, int H, int W, int Hc, int Wc,
                      float *out12, void *scratch, int device, void *stream);
/* test.py:133-140: est[:Hc,:Wc] -> (est - lo)/(hi - lo) -> matplotlib 'jet' (256-entry table) -> rgb (B,Hc,Wc,3) uint8.
 * scratch >= 768 bytes. */
int dff_depth_to_jet(const float *est, int B, int H, int W, int Hc, int Wc, float lo, float hi, uint8_t *rgb, void *scratch, int device,
                     void *stream);

/* x (B,C,S,H,W) fp32 reference layout; alpha (B,3,S) [a0,a1,a2 per slice] or NULL (zeros); fov (B,S);
 * out (B,C,S,H,W); flow (B,2,S,H,W) or NULL.  Bug-compatible with the reference only for B == 1 (B > 1 uses
 * sample 0's scale correction exactly as End_to_End.py:112-118 does). */
int dff_fov_warp(const float *x, const float *alpha, const float *fov, int B, int C, int S, int H, int W, float *out,
                 float *flow, int device, void *stream);

/* layout helpers: reference (B,C,S,H,W) fp32 <-> channels-last (B,S,H,W,Cp) of `elem` type (Cp >= C, zero pad) */
int dff_to_channels_last(const float *src, int B, int C, int S, int H, int W, void *dst, int Cp, int elem, int device,
                         void *stream);
int dff_from_channels_last(const void *src, int B, int C, int S, int H, int W, int Cp, int elem, float *dst,
                           int device, void *stream);

/* ---- train-mode building blocks ----------------------------------------------------------------------
 * The train path keeps PyTorch's autograd as the tape (dffinthewild_b200/train.py: one autograd.Function per fused operator);
 * every forward and backward computation is one of the calls below.  Tensors are channels-last (B,S,H,W,C) of `elem` type;
 * parameter gradients, statistics and reductions are fp32.  Replaces ATen's convolution_backward, native_batch_norm(+_backward),
 * threshold_backward, max/avg_pool3d_backward, upsample_bilinear2d_backward and softplus_backward behind `Total.backward()`
 * (train_codes/train_code_Defocus.py:167) and nn.BatchNorm3d in batch-statistics mode (train_codes/Depth_Estimation_Network.py:355). */

/* dx[.., ci0:ci0+nci] of a conv / transposed-conv layer (weight in the reference layout) from dy (B,S,OH,OW,CoS).
 * dx is (B,S,IH,IW,nci): the gradient of one source of a (virtually concatenated) input. */
size_t dff_conv3d_dgrad_scratch_bytes(int Cin, int Cout, int kd, int kh, int kw);
int dff_conv3d_dgrad(const void *dy, int CoS, int B, int S, int OH, int OW, const float *weight, int Cin, int Cout, int kd,
                     int kh, int kw, int stride_hw, int dil_hw, int transposed, int ci0, int nci, void *dx, int elem,
                     void *scratch, int device, void *stream);
/* dw (fp32, reference layout, overwritten) from the layer input in0 [+ in1] (B,S,IH,IW,C0[+C1]) and dy (.., CoS). */
int dff_conv3d_wgrad(const void *in0, int C0, const void *in1, int C1, int B, int S, int IH, int IW, const void *dy, int CoS,
                     int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, float *dw, int elem,
                     int device, void *stream);
/* Same, ADDED to dw (no clearing): lets a training loop accumulate every layer's gradient straight into its slot of one flat,
 * once-per-step-zeroed gradient bucket (dffinthewild_b200/train_step.py) instead of a temporary + an add per parameter. */
int dff_conv3d_wgrad_acc(const void *in0, int C0, const void *in1, int C1, int B, int S, int IH, int IW, const void *dy, int CoS,
                         int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int dil_hw, int transposed, float *dw, int elem,
                         int device, void *stream);
/* out = [relu]( BN_batchstats(x) + res_pre ) + res_post ; saves mean / invstd, updates running statistics in place
 * (momentum, unbiased variance) when given.  gamma == NULL: no BatchNorm (activation / adds only).
 * scale_shift: 2*C floats of scratch; scratch: dff_bn_scratch_bytes(C). */
size_t dff_bn_scratch_bytes(int C);
int dff_bn_train_forward(const void *x, int64_t npix, int C, int elem, const float *gamma, const float *beta,
                         float *running_mean, float *running_var, float momentum, float eps, const void *res_pre,
                         const void *res_post, int relu, void *out, float *save_mean, float *save_invstd, float *scale_shift,
                         void *scratch, int device, void *stream);
/* backward of the above w.r.t. x (dx), res_pre (dres = masked dy), gamma, beta.  y_relu: the stored output BEFORE res_post
 * (ReLU mask) or NULL when there was no ReLU; save_mean == NULL: no BatchNorm. */
int dff_bn_train_backward(const void *dy, const void *y_relu, const void *x, const float *save_mean, const float *save_invstd,
                          const float *gamma, int64_t npix, int C, int elem, void *dx, void *dres, float *dgamma, float *dbeta,
                          void *scratch, int device, void *stream);
/* The same operator with the module's RUNNING statistics (a BatchNorm3d in eval mode inside a differentiable forward: frozen-BN
 * fine-tuning, input gradients of an eval-mode network).  Nothing is updated; save_mean / save_invstd receive running_mean and
 * 1/sqrt(running_var + eps) for the backward, which has no batch-mean terms: dx = gamma * invstd * g. */
int dff_bn_eval_forward(const void *x, int64_t npix, int C, int elem, const float *gamma, const float *beta,
                        const float *running_mean, const float *running_var, float eps, const void *res_pre, const void *res_post,
                        int relu, void *out, float *save_mean, float *save_invstd, float *scale_shift, int device, void *stream);
int dff_bn_eval_backward(const void *dy, const void *y_relu, const void *x, const float *save_mean, const float *save_invstd,
                         const float *gamma, int64_t npix, int C, int elem, void *dx, void *dres, float *dgamma, float *dbeta,
                         void *scratch, int device, void *stream);
int dff_add(const void *a, const void *b, int64_t n, int elem, void *out, int device, void *stream);

/* ---- loss and optimizer of the reference's training loop (train_codes/train_code_Defocus.py:17-19,67,160-168) -----
 * Total = sum_k weights[k] * mean over valid pixels of (pred4[k] - gt)^2 for the four heads (n elements each; mask: 1 byte per pixel).
 * grad4[k] receives d Total / d pred4[k].  stats (8 floats): [0] valid count, [1] Total, [2..5] per-head masked MSE, [6] 1/count.
 * scratch: 8192 doubles.  Replaces four boolean-mask gathers (`pred[mask]`, `gt[mask]`) + nn.MSELoss and their backward. */
int dff_masked_mse(const float *const pred4[4], const float *gt, const uint8_t *mask, int64_t n, const float weights[4],
                   float *const grad4[4], float *stats, void *scratch, int device, void *stream);
/* One Adam step (no weight decay, no amsgrad) over flat fp32 buffers of n elements; `step` counts from 1; grad_scale (optional,
 * device scalar) multiplies the gradients first.  Element-wise arithmetic, order and rounding of torch.optim.Adam's CUDA
 * default, so states and checkpoints stay interchangeable with the reference's optimizer (train_code_Defocus.py:67,168). */
int dff_adam_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int step, const float *grad_scale, int device, void *stream);
/* (1,k,k) pooling of (BS,H,W,C) channels-last volumes, forward and backward (max: first maximum in row-major order). */
int dff_pool3d(const void *x, int BS, int H, int W, int C, int k, int is_max, int elem, void *out, int device, void *stream);
/* hourglassup's three average pools (AvgPool3d (1,2,2), (1,4,4), (1,8,8) of the same volume, train_codes/Depth_Estimation_Network.py:
 * 183-187, 248-250) in one pass: x (BS,H,W,C) bf16 channels-last, C, H, W multiples of 8 -> out2, out4, out8 (bf16). */
int dff_avgpool_pyramid(const void *x, int BS, int H, int W, int C, void *out2, void *out4, void *out8, int device, void *stream);
int dff_pool3d_backward(const void *x, const void *dy, int BS, int H, int W, int C, int k, int is_max, int elem, void *dx,
                        int device, void *stream);
/* d cost (B,S,h,w) from d depth (B,H,W) */
int dff_depth_head_backward(const float *cost, int h, int w, const float *fd, const int64_t fd_strides[4], int B, int S, int H,
                            int W, const float *ddepth, float *dcost, int device, void *stream);

/* ---- End-to-End alignment network (FlowNetwork.forward, End_to_End/End_to_End.py:63-104) --------------------------------------
 * dff_flow_forward runs the whole alignment network in one call: six resnet_block_2d_OF blocks, three alignment heads (BatchNorm
 * folded once by dff_pack_weights(DFF_NET_FLOW)), the feature warps, the pairwise volumes, the per-slice spatial means and the final
 * FOV_warp of the focal stack.  FS (B,3,S,H,W) fp32, fov (B,S) fp32 -> FS_out (B,3,S,H,W) fp32; alpha_out (B,3,S) optional: the
 * estimated (scale correction, x shift, y shift) per slice.  mode: DFF_FP32 (FFMA parity path) or DFF_BF16 (tcgen05 kernels, bf16
 * feature volumes; alpha, the final warp and the stack stay fp32).  H, W multiples of 4.  The three single operators below are its
 * building blocks (unit-parity surface). */
size_t dff_flow_workspace_bytes(int B, int S, int H, int W, int mode);
int dff_flow_forward(const void *packed_flow, const float *FS, const float *fov, int B, int S, int H, int W, float *FS_out,
                     float *alpha_out, void *workspace, size_t workspace_bytes, int mode, int device, void *stream);
/* FOV_warp (End_to_End.py:106-134) of a channels-last volume (B,S,H,W,C), C % 4 == 0; alpha (B,3,S) or NULL, fov (B,S) */
int dff_fov_warp_cl(const void *x, const float *alpha, const float *fov, int B, int C, int S, int H, int W, void *out, int elem,
                    int device, void *stream);
/* input of an alignment head (End_to_End.py:71-76): out (B,S,H,W,2C+8) = [ feat[b,S-1] | feat[b,s] | flow_x, flow_y, 0 x 6 ] with the
 * flow field of the current (alpha, fov) computed analytically */
int dff_pair_volume(const void *feat, const float *alpha, const float *fov, int B, int C, int S, int H, int W, void *out, int elem,
                    int device, void *stream);
/* AdaptiveAvgPool3d((S,1,1)) of the head output x (B,S,H,W,Cs) fp32 + scaling + running sum (End_to_End.py:78-79, 88-90):
 * alpha_out[b][c][s] = alpha_in[b][c][s] + s_c * mean_{y,x} x[b,s,y,x,c], c < 3 (alpha_in may be NULL) */
int dff_spatial_mean_accum(const float *x, int Cs, int B, int S, int H, int W, const float *alpha_in, float s0, float s1, float s2,
                           float *alpha_out, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DFF_B200_H_ */
