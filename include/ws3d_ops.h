/*
 * ws3d_ops.h -- C ABI of libws3d_ops.so: the B200 (sm_100a) implementation of the
 * WS3D PointNet++ set-abstraction / feature-propagation ops, roipool3d and iou3d.
 *
 * This is the drop-in boundary.  Every entry point replaces one launcher that the
 * reference's pybind wrappers call (file:line under the reference tree is cited on
 * each declaration); arguments keep the reference's order and meaning, with two
 * systematic differences:
 *   - a trailing stream argument everywhere (the reference's iou3d/roipool3d
 *     launchers use the legacy default stream; pass NULL for that behaviour);
 *   - an int return value: 0 on success, otherwise a cudaError_t code (the
 *     reference prints and calls exit(); see ws3d_last_error()).
 *
 * All pointers are DEVICE pointers to contiguous row-major arrays unless a
 * parameter is explicitly named *_host.  float = IEEE binary32, int = int32.
 * The caller owns and pre-initialises every output exactly as the reference's
 * Python wrappers do (1e10-filled temp for FPS, zero-filled grads / pooled
 * features / flags; ball_query's idx needs no fill here).  No entry point allocates device
 * memory except the *_host convenience wrappers and ws3d_roipool3d / ws3d_nms*
 * when called with workspace == NULL (they then use a cached per-device scratch
 * buffer that is grown on demand and never shrunk).
 *
 * No torch types, no C++ types: bindable from ctypes / cgo / JNI / pybind alike.
 */
#ifndef WS3D_OPS_H_
#define WS3D_OPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same object as cudaStream_t (struct CUstream_st *) */
typedef struct CUstream_st *ws3d_stream_t;

#define WS3D_ABI_VERSION 1

/* Library identity / diagnostics ------------------------------------------- */
int ws3d_abi_version(void);
/* Human-readable text of the last failure on the calling thread ("" if none). */
const char *ws3d_last_error(void);
/* Number of kernels this library has launched since load (all threads). */
uint64_t ws3d_launch_count(void);
/* Scratch arena (0..7) used by the calling thread's subsequent launches; returns the previous one.
 * The cached scratch buffers (cell grids, NMS masks) are per (device, arena): forward passes that are
 * in flight at the same time -- e.g. two CUDA graphs replayed on different streams -- must be issued
 * (or captured) under different arenas.  The reference has no counterpart (it cudaMallocs per call,
 * iou3d.cpp:87, roipool3d_kernel.cu:214). */
int ws3d_set_workspace_arena(int arena);
/* Number of scratch arenas (valid arguments of ws3d_set_workspace_arena are 0 .. ws3d_num_arenas() - 1). */
int ws3d_num_arenas(void);
/* Bytes of cached scratch on the current device: retired (outgrown, kept alive because a captured CUDA graph or a
 * queued launch may still hold the address) plus, unless retired_only, the live buffers of every arena. */
size_t ws3d_scratch_bytes(int retired_only);
/* Frees the retired scratch buffers of the current device (all = 0) or every cached buffer (all = 1).  Synchronises
 * the device first.  The CALLER guarantees that no captured graph that will be replayed again was captured while a
 * buffer that is freed here was live (destroy such graphs first; ws3d_b200.graphs runners do so in close()).  The
 * reference frees its scratch after every call (iou3d.cpp:118, roipool3d_kernel.cu:235-236). */
int ws3d_release_scratch(int all);
/* Upper bound on the SMs a persistent kernel (ws3d_mlp_layer) spreads over; 0 = all 148.  Used when a
 * latency-bound kernel of another batch (FPS) is meant to run beside it.  Returns the previous value. */
int ws3d_set_sm_budget(int sms);
/* How the calling thread's subsequent furthest-point-sampling launches trade latency for SMs:
 * 0 = automatic (default: thread-block clusters, 4 SMs per cloud, unless the batch leaves a large cloud fewer
 * than that), 1 = throughput (for 2048 <= n <= 16384: Morton buckets with exact culling, running distances in
 * shared memory, two samples per traversal of the latency chain when the second is provably the next one,
 * SEVERAL CLOUDS PER SM -- two at 16384 points, eight at 4096: 1.25x the latency at an eighth of the SM time --
 * for samplers that run beside other work, see ws3d_b200.graphs.StreamedBackboneRunner),
 * 2 = latency (never that kernel).  Results are bit-identical in every mode.  Returns the previous mode. */
int ws3d_set_fps_mode(int mode);
/* Clouds that share one CTA (= one SM) when the throughput sampler runs a batch of b clouds of n points under the calling
 * thread's mode (diagnostics: SM-time accounting of a pipelined step). */
int ws3d_fps_clouds_per_cta(int b, int n);

/* ---- pointnet2_cuda -------------------------------------------------------- */

/* Replaces furthest_point_sampling_kernel_launcher
 * (pointnet2_lib/pointnet2/src/sampling_gpu.h:26-27, sampling_gpu.cu:211-253).
 * xyz (B,N,3), temp (B,N) running min-distance scratch (read at entry, final
 * values written back; may be NULL = start from 1e10, nothing written),
 * idx (B,M) int32.  Bit-exact index parity with the reference, including its
 * block-size dependent tie-break. */
int ws3d_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                 ws3d_stream_t stream);

/* Extension (fusion of sampling_gpu.cu:211 + :26): FPS that also emits the
 * sampled coordinates new_xyz (B,M,3) from its epilogue, replacing the
 * transpose + gather_points + transpose sequence of pointnet2_modules.py:30-35.
 * new_xyz may be NULL. */
int ws3d_furthest_point_sampling_gather(int b, int n, int m, const float *xyz, float *temp,
                                        int *idx, float *new_xyz, ws3d_stream_t stream);

/* Replaces gather_points_kernel_launcher_fast (sampling_gpu.h:12-13, sampling_gpu.cu:26-44).
 * points (B,C,N), idx (B,M) -> out (B,C,M). */
int ws3d_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                       float *out, ws3d_stream_t stream);

/* Replaces gather_points_grad_kernel_launcher_fast (sampling_gpu.h:19-20, sampling_gpu.cu:65-84).
 * grad_out (B,C,M), idx (B,M) -> grad_points (B,C,N) += (caller zero-fills). */
int ws3d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                            const int *idx, float *grad_points, ws3d_stream_t stream);

/* Replaces ball_query_kernel_launcher_fast (ball_query_gpu.h:12-13, ball_query_gpu.cu:48-67).
 * NOTE the argument order: new_xyz (B,M,3) comes BEFORE xyz (B,N,3), as in the
 * reference wrapper (ball_query.cpp:14-25).  idx (B,M,nsample): every row is written.  A row
 * with no neighbour is ZERO-FILLED by the kernel: the reference leaves it untouched and relies
 * on the caller's zero fill (pointnet2_utils.py:218) -- same result, and callers of this library
 * may pass uninitialised memory. */
int ws3d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int *idx, ws3d_stream_t stream);

/* Extension: ball_query for TWO radii over the same centres in one scan (the multi-scale
 * grouping of PointnetSAModuleMSG, pointnet2_modules.py:37-38, queries the same new_xyz/xyz once
 * per scale).  idx0 (B,M,nsample0), idx1 (B,M,nsample1); each is identical to what
 * ws3d_ball_query returns for its radius (no-neighbour rows zero-filled by the kernel). */
int ws3d_ball_query2(int b, int n, int m, float radius0, int nsample0, float radius1, int nsample1,
                     const float *new_xyz, const float *xyz, int *idx0, int *idx1, ws3d_stream_t stream);

/* Replaces group_points_kernel_launcher_fast (group_points_gpu.h:13-14, group_points_gpu.cu:69-86).
 * points (B,C,N), idx (B,npoints,nsample) -> out (B,C,npoints,nsample). */
int ws3d_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int *idx, float *out, ws3d_stream_t stream);

/* Replaces group_points_grad_kernel_launcher_fast (group_points_gpu.h:19-20, group_points_gpu.cu:27-45). */
int ws3d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                           const int *idx, float *grad_points, ws3d_stream_t stream);

/* Extension: gradient of ws3d_group_concat with respect to `features` (the autograd mirror of
 * QueryAndGroup.forward, pointnet2_utils.py:241-264, where GroupingOperation.backward :187-197 receives
 * a contiguous copy of grad[:, 3:]).  grad_out (B, 3*use_xyz + C, M, nsample) is read in place -- the
 * coordinate channels are skipped by addressing, not by a strided copy; grad_features (B,C,N) is
 * accumulated into (caller zero-fills), exactly like ws3d_group_points_grad. */
int ws3d_group_concat_grad(int b, int n, int m, int c, int nsample, int use_xyz, const float *grad_out,
                           const int *idx, float *grad_features, ws3d_stream_t stream);

/* Extension: QueryAndGroup.forward (pointnet2_utils.py:241-264) as one pass per
 * scale: ball_query -> gather xyz -> subtract centre -> gather features -> concat,
 * written once.  xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) or NULL,
 * out (B, 3*use_xyz + C, M, nsample), idx_out (B,M,nsample) zero-filled or NULL.
 * Results are identical to the unfused sequence. */
int ws3d_query_and_group(int b, int n, int m, int c, float radius, int nsample, int use_xyz,
                         const float *xyz, const float *new_xyz, const float *features,
                         float *out, int *idx_out, ws3d_stream_t stream);

/* The grouping half of the above for a precomputed idx (B,M,nsample): gathers xyz (minus the
 * centre) and every feature channel into out (B, 3*use_xyz + C, M, nsample) in one pass
 * (pointnet2_utils.py:250-257). */
int ws3d_group_concat(int b, int n, int m, int c, int nsample, int use_xyz, const float *xyz,
                      const float *new_xyz, const float *features, const int *idx, float *out,
                      ws3d_stream_t stream);

/* Replaces three_nn_kernel_launcher_fast (interpolate_gpu.h:13-14, interpolate_gpu.cu:55-74).
 * unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) SQUARED distances, idx (B,N,3). */
int ws3d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx, ws3d_stream_t stream);

/* Extension: three_nn that also emits the normalised inverse-distance interpolation weights
 * PointnetFPModule.forward derives from its result with five elementwise torch kernels
 * (pointnet2_modules.py:139-144): dist = sqrt(dist2); r = 1 / (dist + 1e-8); weight = r / sum_k r_k,
 * each step one IEEE float32 operation.  weight (B,N,3); dist2 may be NULL (not written). */
int ws3d_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                          int *idx, float *weight, ws3d_stream_t stream);

/* Replaces three_interpolate_kernel_launcher_fast (interpolate_gpu.h:20-21, interpolate_gpu.cu:99-117).
 * points (B,C,M), idx (B,N,3), weight (B,N,3) -> out (B,C,N). */
int ws3d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                           const float *weight, float *out, ws3d_stream_t stream);

/* Replaces three_interpolate_grad_kernel_launcher_fast (interpolate_gpu.h:27-28, interpolate_gpu.cu:144-160).
 * grad_out (B,C,N) -> grad_points (B,C,M) += (caller zero-fills). */
int ws3d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                const float *weight, float *grad_points, ws3d_stream_t stream);

/* Extension: grouping with an affine epilogue,
 *   out[b,c,j,s] = act( P[b,c,i] + wx[c,:] . (xyz[b,i,:] - new_xyz[b,j,:]) + shift[c] ),  i = idx[b,j,s];
 * flags bit 0 = ReLU, bit 1 = round to TF32.  P (B,c,n), xyz (B,n,3), new_xyz (B,m,3), wx (c,3), shift (c),
 * idx (B,m,nsample) -> out (B,c,m,nsample).  It is the first layer of a set-abstraction MLP
 * (pointnet2_modules.py:37-44) with the feature part of the 1x1 convolution moved in front of QueryAndGroup:
 * W [xyz[i] - centre ; f[i]] = (W_f f)[i] + W_x (xyz[i] - centre): P = W_f f is formed on the n source points
 * instead of the m * nsample grouped columns, the coordinate channels are applied in FP32, and the
 * (3 + C)-channel grouped tensor is never written. */
int ws3d_group_affine(int b, int n, int m, int c, int nsample, const float *P, const float *xyz,
                      const float *new_xyz, const float *wx, const float *shift, const int *idx, int flags,
                      float *out, ws3d_stream_t stream);

/* Extension: three_interpolate with an affine epilogue,
 *   out[b,c,i] = act( sum_k weight[b,i,k] * points[b,c,idx[b,i,k]] + scale1[c] * row1[b,i] + shift[c] ),
 * scale1 (c) and row1 (B,n) optional (both or neither), shift (c) optional; flags bit 0 = ReLU, bit 1 =
 * round the result to TF32.  The stencil weights do not depend on the channel, so the interpolated part
 * of a feature-propagation layer's first 1x1 convolution (pointnet2_modules.py:139-154) can be applied
 * to the KNOWN points (m, typically n / 4 columns) and its product interpolated: interp(W f) = W interp(f).
 * Shapes served: m <= 8192, n >= 256, c % 4 == 0 (others: invalid argument). */
int ws3d_three_interpolate_affine(int b, int c, int m, int n, const float *points, const int *idx,
                                  const float *weight, const float *scale1, const float *row1,
                                  const float *shift, int flags, float *out, ws3d_stream_t stream);

/* Extension: one shared-MLP layer -- 1x1 conv with BatchNorm(eval) folded
 * in, optional ReLU, optional max-pool over runs of `pool` consecutive columns -- on the tcgen05 tensor
 * cores (TF32 inputs, FP32 accumulate), replacing the conv / BN / ReLU / max-pool kernel sequence of
 * pointnet2_modules.py:40-44,154 + pytorch_utils.py:5-32 in inference.
 *   w (c_out_pad, 32*(ceil(c1/32)+ceil(c2/32))) row-major, zero padded, c_out_pad % 128 == 0;
 *   x1 (B, c1, cols), x2 (B, c2, cols) or NULL (second K range, e.g. skip features); shift (c_out_pad);
 *   out (B, c_out, cols) or, when pool > 0, (B, c_out, cols / pool).  cols % 4 == 0; pool is a power of two <= 128 dividing cols (and 128 / r, see relu bits 4-5).
 *   relu: bit 0 = apply ReLU; bit 1 = round the stored output to the nearest TF32 value (use for every layer
 *   whose output feeds another ws3d_mlp_layer: the tensor core truncates FP32 operands to TF32);
 *   bits 4-5 = log2(r), r in {1, 2, 4}: rows [k*128/r, (k+1)*128/r) of w (and shift) each hold a copy of the
 *   c_out <= 128/r real rows (c_out_pad must be 128) -- lets all four epilogue warps work on narrow layers. */
int ws3d_mlp_layer(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w,
                   const float *shift, const float *x1, const float *x2, float *out, int relu, int pool,
                   ws3d_stream_t stream);
/* Same, writing channels [out_coff, out_coff + c_out) of an (B, out_ctot, cols or cols / pool) tensor: the slot of one
 * scale in the concatenated multi-scale output (pointnet2_modules.py:55) without a torch.cat pass. */
int ws3d_mlp_layer_into(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w,
                        const float *shift, const float *x1, const float *x2, float *out, int out_ctot,
                        int out_coff, int relu, int pool, ws3d_stream_t stream);

/* ---- training-mode shared MLP (BASELINE config 3) -------------------
 * conv1x1 -> BatchNorm(batch statistics) -> ReLU [-> max-pool over nsample], forward and backward, replacing the cuDNN
 * conv / BN / ReLU / max-pool kernels and their autograd mirrors behind pytorch_utils.py:5-32 and
 * pointnet2_modules.py:40-44 in training.  All activations are (B, C, cols) channel-major float32; cols % 4 == 0. */

/* Forward GEMM with statistics: y (B,c_out,cols) = W [x1 ; x2] (weights laid out as for ws3d_mlp_layer, zero_shift = c_out_pad
 * zeros) on the tcgen05 tensor cores, and stats[0..c_out) += sum_y, stats[c_out..2c_out) += sum_y^2 (double, caller zeroes). */
int ws3d_mlp_layer_stats(int b, int c_out, int c_out_pad, int c1, int c2, int cols, const float *w, const float *zero_shift,
                         const float *x1, const float *x2, float *y, double *stats, ws3d_stream_t stream);
/* Batch statistics -> per-channel affine: scale = gamma * invstd, shift = beta - mean * scale (gamma / beta may be NULL = 1 / 0);
 * running_mean / running_var (may be NULL) updated as torch.nn.BatchNorm does (momentum, unbiased variance). count = B * cols. */
int ws3d_bn_finalize(int c, double count, const double *stats, const float *gamma, const float *beta, float eps, float momentum,
                     float *running_mean, float *running_var, float *scale, float *shift, float *mean, float *invstd,
                     ws3d_stream_t stream);
/* z = act(y * scale[c] + shift[c]); flags: 1 ReLU, 2 round z to TF32.  pool > 0 (power of two in [4,128] dividing cols):
 * z is (B,c,cols/pool) = max over every run of `pool` columns and arg (B,c,cols/pool) uint8 its first arg-max. */
int ws3d_bn_relu_apply(int b, int c, int cols, int pool, const float *y, const float *scale, const float *shift, int flags,
                       float *z, unsigned char *arg, ws3d_stream_t stream);
/* Backward reductions: sums[0..c) += sum dA, sums[c..2c) += sum dA * xhat (double, caller zeroes), dA = dz where the
 * activation passed, xhat = (y - mean) * invstd; dz is (B,c,cols) or, with pool > 0, (B,c,cols/pool) routed through arg. */
int ws3d_bn_relu_bwd_reduce(int b, int c, int cols, int pool, const float *y, const float *dz, const unsigned char *arg,
                            const float *scale, const float *shift, const float *mean, const float *invstd, int flags,
                            double *sums, ws3d_stream_t stream);
/* dy (B,c,cols) = scale * (dA - sums[c] / count - xhat * sums[C + c] / count), rounded to TF32; count <= 0: dy = scale * dA
 * (a layer without batch norm).  dgamma = sums[c..2c), dbeta = sums[0..c). */
int ws3d_bn_relu_bwd_apply(int b, int c, int cols, int pool, const float *y, const float *dz, const unsigned char *arg,
                           const float *scale, const float *shift, const float *mean, const float *invstd, int flags,
                           const double *sums, double count, float *dy, ws3d_stream_t stream);
/* Weight gradient on the tensor cores: dw (c_out x c_in, row stride ldw floats) += sum_b dy[b] (c_out x cols) x[b]^T; TF32
 * operands, FP32 accumulation, split over the contraction axis with FP32 atomics (caller zeroes dw). */
int ws3d_mlp_wgrad(int b, int c_out, int c_in, int cols, const float *dy, const float *x, float *dw, int ldw,
                   ws3d_stream_t stream);

/* Extension: the `_break_up_pc` step of lib/net/pointnet2_msg.py:52-60 in one launch:
 * pc (B,N,3+C) -> xyz (B,N,3) and features (B,C,N) (channel-major; NULL when C == 0). */
int ws3d_split_pointcloud(int b, int n, int c, const float *pc, float *xyz, float *features,
                          ws3d_stream_t stream);

/* Extension: the opposite packing -- xyz (B,N,3) and channel-major features (B,C,N) (NULL when C == 0) -> point-major
 * rows (B,N,ld) = [xyz | features | zeros], ld >= 3 + C: the operand layout ws3d_sa_mlp_fused_rows gathers from
 * (what `pts[..., 3:].transpose(1, 2)` in lib/net/rcnn_net.py undoes for the pooled Stage-2 input). */
int ws3d_pack_rows(int b, int n, int c, int ld, const float *xyz, const float *features, float *rows,
                   ws3d_stream_t stream);

/* Extension (complete form): ONE set-abstraction scale in one kernel --
 * QueryAndGroup's grouping (pointnet2_utils.py:241-264, use_xyz = True), the three SharedMLP layers
 * (pytorch_utils.py:5-32: conv1x1 + BatchNorm(eval, folded) + ReLU) and the max-pool over nsample
 * (pointnet2_modules.py:40-44), without materialising the grouped tensor or the intermediate
 * activations (they stay in tensor memory).  TF32 inputs, FP32 accumulation.
 *   xyz (B,n,3), new_xyz (B,m,3), features (B,c_feat,n) or NULL when c_feat == 0, idx (B,m,nsample);
 *   k0 = roundup(3 + c_feat, 8), n_l = roundup(c_l, 16);
 *   w1 (n1, 32*ceil(k0/32)) with columns [dx,dy,dz,features...], w2 (n2, 32*ceil(n1/32)),
 *   w3 (n3, 32*ceil(n2/32)): row-major, zero padded, rounded to TF32; shift_l (n_l) zero padded;
 *   out (B, out_ctot, m): channels [out_coff, out_coff + c3) are written (the scale's slot of the
 *   concatenated multi-scale output, pointnet2_modules.py:55).
 * ws3d_sa_mlp_fused_supported: 1 when the shape fits (nsample a power of two <= 128, weights resident
 * in shared memory, <= 512 tensor-memory columns), else 0 -- callers then use ws3d_mlp_layer. */
int ws3d_sa_mlp_fused_supported(int c_feat, int nsample, int c1, int c2, int c3);
int ws3d_sa_mlp_fused(int b, int n, int m, int nsample, int c_feat, const float *xyz,
                      const float *new_xyz, const float *features, const int *idx, int c1, int c2,
                      int c3, const float *w1, const float *shift1, const float *w2,
                      const float *shift2, const float *w3, const float *shift3, float *out,
                      int out_ctot, int out_coff, ws3d_stream_t stream);
/* Same, gathering the grouped points from POINT-MAJOR rows (B, n, ld) = [x, y, z, features[0..c_feat), zeros] (ld % 4 == 0,
 * 16-byte aligned; e.g. the (B,N,4) input cloud itself for the first level): one contiguous row per grouped point instead
 * of 3 + c_feat scattered 4-byte reads of a channel-major tensor.  out_pm (B, m, ld_pm) or NULL: this scale's pooled
 * channels also written in that layout at columns [3 + out_coff, 3 + out_coff + c3) for the next level; the scale called
 * with pm_xyz != 0 also writes the centre coordinates (columns 0..2) and zeroes columns [3 + out_ctot, ld_pm).
 * Results are bit-identical to ws3d_sa_mlp_fused on the equivalent channel-major inputs. */
int ws3d_sa_mlp_fused_rows(int b, int n, int m, int nsample, int c_feat, const float *rows, int ld,
                           const float *new_xyz, const int *idx, int c1, int c2, int c3, const float *w1,
                           const float *shift1, const float *w2, const float *shift2, const float *w3,
                           const float *shift3, float *out, int out_ctot, int out_coff, float *out_pm, int ld_pm,
                           int pm_xyz, ws3d_stream_t stream);

/* ---- iou3d_cuda ------------------------------------------------------------ */

/* Replaces boxesoverlapLauncher (lib/utils/iou3d/src/iou3d.cpp:26, iou3d_kernel.cu:354-363).
 * boxes_a (Na,5), boxes_b (Nb,5) [x1,y1,x2,y2,ry] -> ans (Na,Nb) overlap area. */
int ws3d_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b,
                           float *ans, ws3d_stream_t stream);

/* Replaces boxesioubevLauncher (iou3d.cpp:27, iou3d_kernel.cu:365-372). */
int ws3d_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b,
                       float *ans, ws3d_stream_t stream);

/* Bytes of device scratch ws3d_nms / ws3d_nms_normal need for boxes_num boxes. */
size_t ws3d_nms_workspace_bytes(int boxes_num);

/* Replaces nmsLauncher + the host greedy scan (iou3d.cpp:73-121, iou3d_kernel.cu:374-380)
 * with an all-device pipeline.  boxes (N,5) sorted by descending score.
 * keep (N) int64 DEVICE buffer receives the kept indices in ascending order,
 * num_keep (1) int32 DEVICE receives their count.  workspace: device scratch of
 * ws3d_nms_workspace_bytes(N) bytes, or NULL to use the library's cached scratch. */
int ws3d_nms(const float *boxes, int boxes_num, float nms_overlap_thresh, int64_t *keep,
             int *num_keep, void *workspace, ws3d_stream_t stream);

/* Same for nmsNormalLauncher (iou3d.cpp:123-171, iou3d_kernel.cu:382-388): axis-aligned IoU. */
int ws3d_nms_normal(const float *boxes, int boxes_num, float nms_overlap_thresh, int64_t *keep,
                    int *num_keep, void *workspace, ws3d_stream_t stream);

/* Reference-signature convenience: nms_gpu(boxes, keep_cpu_int64, thresh) -> num kept
 * (iou3d.cpp:73).  keep_host is HOST memory (N int64).  Synchronises the stream.
 * Returns the number kept (>= 0) or -(cudaError_t). */
int ws3d_nms_host(const float *boxes, int boxes_num, float nms_overlap_thresh,
                  int64_t *keep_host, ws3d_stream_t stream);
int ws3d_nms_normal_host(const float *boxes, int boxes_num, float nms_overlap_thresh,
                         int64_t *keep_host, ws3d_stream_t stream);

/* Extension: the DIAGONAL of boxes_iou3d_gpu
 * (lib/utils/iou3d/iou3d_utils.py:21-56) for n aligned pairs -- all that the Stage-2 losses keep of
 * their fg x fg matrices (lib/net/train_functions.py:258-260, :287-289).  boxes_a, boxes_b (n,7)
 * [x,y,z,h,w,l,ry]; iou2d, iou3d (n) (either may be NULL).  BEV conversion (kitti_utils.py:134-147),
 * rotated overlap and the height / volume arithmetic in one launch; bit-identical to the diagonal. */
int ws3d_boxes_iou3d_aligned(int n, const float *boxes_a, const float *boxes_b, float *iou2d,
                             float *iou3d, ws3d_stream_t stream);

/* Extension (row f3): greedy "radius NMS" of tools/eval_auto.py:263-279.  centers (n,2) BEV (x,z)
 * ALREADY sorted by descending score; candidate i is kept iff its distance (lib/utils/distance.py:3,
 * float32) to every centre kept before it is > radius.  keep (n) int64 device indices into the sorted
 * order, num_keep (1) int32 device.  workspace: ws3d_nms_workspace_bytes(n) bytes or NULL. */
int ws3d_radius_nms(const float *centers, int n, float radius, int64_t *keep, int *num_keep,
                    void *workspace, ws3d_stream_t stream);

/* Extension (row f3): cylinder crop of tools/eval_auto.py:289-291,:327-343.  pts (n,3), centers (m,2)
 * BEV (x,z).  Point i belongs to centre c iff distance_2(centre, (x_i,z_i)) < radius.  idx (m,cap)
 * gets the first `cap` members of every centre in index order (other slots untouched), cnt (m) the
 * full member count, any (n) bytes (caller-zeroed; may be NULL) is set to 1 for members of any centre. */
int ws3d_cylinder_query(int n, int m, int cap, float radius, const float *pts, const float *centers,
                        int *idx, int *cnt, unsigned char *any, ws3d_stream_t stream);

/* Extension (row f4): Stage-1 training labels, KittiRCNNDataset.generate_gaussian_training_labels
 * (lib/datasets/kitti_rcnn_dataset.py:529-573), batched.  pts (B,n,3) rect-camera points, gt_boxes3d
 * (B,g_max,7) [x,y,z,h,w,l,ry] padded, num_gt (B) valid boxes per scene (NULL = g_max everywhere;
 * g_max <= 512).  Per point: d_k = sqrt((x-bx)^2 + (y*gauss_height)^2 + (z-bz)^2) in float32 as numpy
 * evaluates it; cls_label (B,n) = exp(-m^2 / (2 gauss_cov)) with m = min_k clip(d_k - gauss_status,
 * 0, 100) (the scipy Gaussian of :563-564, float64 inside, stored as float32); reg_label (B,n,3) =
 * (bx - x, 0, bz - z) of the nearest box where min_k d_k < fg_radius (4.0 there), else 0.  Scenes
 * without boxes get zeros.  cfg defaults: gauss_height 0.707, gauss_status 0.7, gauss_cov 1.5. */
int ws3d_gaussian_rpn_labels(int b, int n, int g_max, const float *pts, const float *gt_boxes3d,
                             const int *num_gt, float gauss_height, float gauss_status,
                             float gauss_cov, float fg_radius, float *cls_label, float *reg_label,
                             ws3d_stream_t stream);

/* ---- loss-side box math and loader subsampling ---------------- */

/* Replaces boxes3d_to_corners3d_torch (lib/utils/kitti_utils.py:104-131, ~15 eager torch kernels per call).
 * boxes3d (N,7) [x, y(bottom), z, h, w, l, ry] -> corners (N,8,3); flip adds pi to ry (:113-114). */
int ws3d_boxes3d_to_corners3d(int n, const float *boxes3d, int flip, float *corners, ws3d_stream_t stream);

/* The corner distance of the Stage-2 corner loss (lib/net/train_functions.py:266-271): for aligned (N,7) box pairs,
 * dist (N,8) = min(|P - G|, |P - G_flipped|) over the eight corners P of the predicted box, G of the ground truth and
 * G_flipped of the ground truth turned by pi -- three corner computations, two norms and a min in one launch. */
int ws3d_corner_distance(int n, const float *pred_boxes3d, const float *gt_boxes3d, float *dist, ws3d_stream_t stream);
/* Its gradient with respect to the predicted boxes: grad_pred (N,7) from grad_dist (N,8) (the ground truth carries
 * no gradient in the reference either).  Subgradients as torch's autograd takes them. */
int ws3d_corner_distance_grad(int n, const float *pred_boxes3d, const float *gt_boxes3d, const float *grad_dist,
                              float *grad_pred, ws3d_stream_t stream);

/* The loader's subsampling to a fixed point count (lib/datasets/kitti_rcnn_dataset.py:424-452) on the device.
 * pts (n,c) rows [x,y,z,features...], depth (n) (used when n > npoints).  The random draws are the HOST's, exactly
 * those numpy makes there, so the result equals the reference's sample:
 *   n > npoints : perm = np.random.permutation(n_near) (what np.random.choice(near_idxs, k, replace=False) draws; its
 *                 first k = npoints - n_far entries are used), n_near = #(depth < near_depth);
 *   n <= npoints: perm = np.random.permutation(n * ceil(npoints / n)) (first npoints entries are used);
 *   order       = np.random.shuffle applied to arange(npoints).
 * out (npoints,c) = pts[choice[order]] with `sub_last` subtracted from the last channel when c > 3 (the intensity shift
 * of :444), choice (npoints) the selected source indices (may be NULL), status (1) int: 0, or 1 when n_near does not
 * match the depths (nothing is written then). */
int ws3d_subsample_points(int n, int c, int npoints, int n_near, float near_depth, float sub_last, const float *pts,
                          const float *depth, const int *perm, const int *order, float *out, int *choice,
                          int *status, ws3d_stream_t stream);

/* ---- roipool3d_cuda -------------------------------------------------------- */

/* Replaces roipool3dLauncher (lib/utils/roipool3d/src/roipool3d.cpp:12-13,
 * roipool3d_kernel.cu:209-237) -- and roipool3dLauncher_slow (:197-207), whose
 * results are identical.  xyz (B,N,3), boxes3d (B,M,7) [x,y,z,h,w,l,ry],
 * pts_feature (B,N,C) -> pooled_features (B,M,S,3+C) and pooled_empty_flag (B,M),
 * both zero-filled by the caller.  No (B,N,M) flag tensor is materialised. */
int ws3d_roipool3d(int batch_size, int pts_num, int boxes_num, int feature_in_len,
                   int sampled_pts_num, const float *xyz, const float *boxes3d,
                   const float *pts_feature, float *pooled_features, int *pooled_empty_flag,
                   ws3d_stream_t stream);

/* Host (CPU) entry points of the reference module, HOST pointers throughout.
 * Replaces pts_in_boxes3d_cpu (roipool3d.cpp:97-125): pts_flag (M,N) int64 <- 0/1. */
int ws3d_pts_in_boxes3d_cpu(int64_t *pts_flag_host, const float *pts_host, const float *boxes3d_host,
                            int boxes_num, int pts_num);

/* Replaces roipool3d_cpu (roipool3d.cpp:127-197): pooled_pts (M,S,3), pooled_features (M,S,C),
 * pooled_empty_flag (M) int64 (cleared here, like the reference's memset). */
int ws3d_roipool3d_cpu(const float *pts_host, const float *boxes3d_host, const float *pts_feature_host,
                       float *pooled_pts_host, float *pooled_features_host,
                       int64_t *pooled_empty_flag_host, int boxes_num, int pts_num, int feature_len,
                       int sampled_pts_num);

#ifdef __cplusplus
}
#endif
#endif /* WS3D_OPS_H_ */
