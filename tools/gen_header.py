"""Write include/emd_b200.h from the `extern "C"` definitions in emd_b200/csrc/*.cu plus the
hand-written per-function documentation below (what each entry point replaces in the reference,
file:line).  `python tools/gen_header.py --write`; tests/test_cpu_abi.py checks the header is current."""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

GROUPS = [
    ("Library plumbing", ["emd_abi_version", "emd_device_check", "emd_last_error_string", "emd_launch_count",
                          "emd_kernel_id_count", "emd_kernel_name", "emd_profile_enable", "emd_profile_collect",
                          "emd_fp32_probe", "emd_fp32_probe_lane_instructions"]),
    ("K1a  EMD motion-embedding deformation, rigid nodes", ["emd_rigid_chunk_size", "emd_rigid_param_count",
                                                          "emd_rigid_deform_fwd", "emd_rigid_deform_bwd"]),
    ("K1c  EMD motion-embedding deformation, SMPL nodes", ["emd_smpl_param_count", "emd_smpl_max_chunks",
                                                         "emd_smpl_reduce_width", "emd_smpl_deform_fwd",
                                                         "emd_smpl_deform_bwd"]),
    ("K1b  spherical harmonics + node activations", ["emd_sh_fwd", "emd_sh_bwd", "emd_activate_fwd", "emd_activate_bwd"]),
    ("K1d  S3Gaussian EMD deformation MLP", ["emd_linear_bwd_workspace_bytes", "emd_linear_fwd", "emd_linear_fwd_tc", "emd_linear_bwd", "emd_linear_bwd_tc", "emd_temb_fwd", "emd_temb_bwd"]),
    ("K1e  HexPlane feature gather (input of the S3Gaussian EMD MLP)", ["emd_hexplane_fwd", "emd_hexplane_fwd_ld", "emd_hexplane_bwd_workspace_bytes",
                                                                       "emd_hexplane_bwd", "emd_hexplane_bwd_ld"]),
    ("K1d' residual application of the S3Gaussian deformation + regulariser sums",
     ["emd_s3g_apply_blocks", "emd_s3g_apply_fwd", "emd_s3g_apply_bwd"]),
    ("Next (SURVEY 8f-2): fused Adam step + densification statistics", ["emd_adam_max_tensors", "emd_adam_step", "emd_densify_stats"]),
    ("Next (SURVEY 8f-3): fused image losses between the rasterizer forward and backward",
     ["emd_image_loss_partials_floats", "emd_image_loss_fwd", "emd_image_loss_bwd"]),
    ("Next (SURVEY 8f-4): voxel LBS weights of the SMPL nodes (K1f)",
     ["emd_voxel_lbs_fwd", "emd_voxel_lbs_bwd", "emd_smpl_weight_grad"]),
    ("Next (SURVEY 8f-4): DeformableNodes deformation network (K1g)",
     ["emd_deform_input_fwd", "emd_dense_fwd", "emd_dense_bwd_workspace_bytes", "emd_dense_bwd", "emd_dense_tc_enabled",
      "emd_dense_set_tc", "emd_deform_apply_fwd",
      "emd_deform_apply_bwd", "emd_deform_embed_grad_workspace_bytes", "emd_deform_embed_grad"]),
    ("K2   projection", ["emd_projection_fwd", "emd_projection_bwd", "emd_dg_preprocess_fwd", "emd_dg_preprocess_bwd"]),
    ("K3   tile intersection", ["emd_scan_workspace_bytes", "emd_cumsum_i32_i64", "emd_exclusive_scan_u32", "emd_exclusive_scan_u8_u32",
                                "emd_exclusive_scan_u8_u32", "emd_isect_emit", "emd_dg_isect_emit"]),
    ("K4   radix sort", ["emd_radix_sort_workspace_bytes", "emd_radix_sort_pairs"]),
    ("K5   tile ranges", ["emd_isect_offsets", "emd_tile_order", "emd_raster_segment_size", "emd_raster_checkpoint_floats",
                        "emd_raster_segout_floats", "emd_raster_segment_slots", "emd_raster_max_ctas"]),
    ("K6/K7 rasterization", ["emd_raster_pack", "emd_raster_sort_records", "emd_rasterize_fwd", "emd_rasterize_bwd_workspace_bytes",
                             "emd_rasterize_bwd", "emd_raster_set_counters"]),
]

DOC = {
    "emd_abi_version": "ABI revision of this header (bumped on any signature change).",
    "emd_device_check": "0 iff the current CUDA device can run this library (compute capability 10.x).",
    "emd_last_error_string": "Text of the last error raised on the calling thread (thread-local).",
    "emd_launch_count": "Number of kernels the library has launched in this process (bench.py's gpu_launches).",
    "emd_kernel_id_count": "Number of kernel ids the profiler distinguishes.",
    "emd_kernel_name": "Name of kernel id `id`.",
    "emd_profile_enable": "When on, every kernel launch is bracketed by CUDA events on its own stream.",
    "emd_profile_collect": "Synchronise recorded events; ADD durations (ms) / launch counts into the [emd_kernel_id_count()] arrays.",
    "emd_raster_set_counters": "Measurement aid: while a non-NULL DEVICE pointer to 8 zero-initialised uint64 is set, emd_rasterize_bwd "
                               "runs the counting build of its kernel: [0] (warp, candidate Gaussian) evaluations, [1] those in which "
                               "a lane blended, [2] blended (pixel, Gaussian) pairs, [3] staged (tile, Gaussian) pairs.  NULL restores "
                               "the product kernel.  Process-global.",
    "emd_fp32_probe": "Measurement aid (no reference counterpart): register-only throughput probe of one instruction class "
                      "(kind 0 FFMA, 1 FFMA2, 2 FADD, 3 FADD2, 4 SHFL.BFLY, 5 MUFU.EX2) on 148 x 8 CTAs; the caller times it with CUDA "
                      "events -> the measured FP32-pipe ceiling bench.py reports beside the nominal one.",
    "emd_fp32_probe_lane_instructions": "Lane-level instructions one emd_fp32_probe(kind, iters) launch executes.",
    "emd_rigid_chunk_size": "Points per (instance, chunk) block of the rigid kernels (sizes the scratch buffers).",
    "emd_rigid_param_count": "Floats in the flat track_* gradient: 2*(d+g+1) + 2*(3*(d+g)+3).",
    "emd_rigid_deform_fwd": "Replaces RigidNodes.transform_means + transform_quats incl. the per-instance Python loops "
                            "(OmniRe/models/nodes/rigid.py:478-568; embedding_track_{rot,trans}_offset :203-246; "
                            "get_temporal_embed/query_time :150-201).  heads = HOST array of 8 DEVICE pointers "
                            "{rot_c_w, rot_c_b, rot_f_w, rot_f_b, trans_c_w, trans_c_b, trans_f_w, trans_f_b}.  "
                            "order/seg_start: points stably sorted by instance + the I+1 boundaries.",
    "emd_rigid_deform_bwd": "VJP of emd_rigid_deform_fwd (what torch.autograd derives from rigid.py:478-568).  "
                            "v_table must be zero-filled by the caller.",
    "emd_smpl_param_count": "Floats in the flat track_smpl_{c,f} gradient: 2*(24*(d+g)+24).",
    "emd_smpl_max_chunks": "Chunks per SMPL instance of V points.",
    "emd_smpl_reduce_width": "Floats per (instance, chunk) backward partial (24x12 + 3).",
    "emd_smpl_deform_fwd": "Replaces SMPLNodes.transform_means_and_quats (OmniRe/models/nodes/smpl.py:438-532; "
                           "embedding_track_smpl_offset :401-436) + SMPLTemplate.forward's chain "
                           "(OmniRe/models/human_body.py:158-172) + pytorch3d matrix_to_quaternion (smpl.py:522).  "
                           "heads = HOST array of 4 DEVICE pointers {smpl_c_w, smpl_c_b, smpl_f_w, smpl_f_b}.",
    "emd_smpl_deform_bwd": "VJP of emd_smpl_deform_fwd.  v_table must be zero-filled by the caller.",
    "emd_sh_fwd": "gsplat.cuda._wrapper.spherical_harmonics (OmniRe/models/gaussians/basics.py:16; calls vanilla.py:388, "
                  "rigid.py:584, smpl.py:555): dirs[N,3] (normalised inside), coeffs[N,K,3] -> out[N,3].",
    "emd_sh_bwd": "VJP of emd_sh_fwd w.r.t. the coefficients (the reference passes detached directions).",
    "emd_activate_fwd": "Tail of {VanillaGaussians,RigidNodes,SMPLNodes}.get_gaussians (vanilla.py:378-414, rigid.py:578-603, "
                        "smpl.py:549-576): view dir, SH -> clamp(+0.5,0,1) for C cameras, sigmoid(opacity) x frame-valid "
                        "mask, exp(scale), normalize(quat).  cam_pos_host = HOST pointer to C x 3 floats.",
    "emd_activate_bwd": "VJP of emd_activate_fwd.",
    "emd_projection_fwd": "gsplat fully_fused_projection (pinhole, quats+scales) as reached from "
                          "OmniRe/models/trainers/base.py:393; also writes tiles_per_gauss (first pass of isect_tiles).",
    "emd_projection_bwd": "VJP of emd_projection_fwd w.r.t. means, quats, scales.",
    "emd_dg_preprocess_fwd": "diff_gauss preprocessCUDA (Inria-derived) as called from "
                             "S3Gaussian/gaussian_renderer/__init__.py:145 with the settings of :49-62: frustum cull, "
                             "cov3D/cov2D, conic, radius, getRect tile count, SH -> RGB.  viewmatrix/projmatrix/campos are "
                             "HOST pointers (row-vector convention of S3Gaussian/scene/cameras.py:55-66).  K = SH bases "
                             "stored per Gaussian, 0 when colours are precomputed.",
    "emd_dg_preprocess_bwd": "VJP of emd_dg_preprocess_fwd w.r.t. means3D, scales, rotations, shs (incl. the view-direction "
                             "term of the SH colour); v_means2d is the gradient w.r.t. the pixel-space mean.",
    "emd_dg_isect_emit": "diff_gauss duplicateWithKeys: key = tile << 32 | float_bits(view depth), value = Gaussian index.",
    "emd_linear_bwd_workspace_bytes": "Workspace bytes of emd_linear_bwd.",
    "emd_linear_fwd": "One Linear layer of the S3Gaussian EMD deformation network with its ReLUs fused "
                      "(S3Gaussian/scene/deformation.py:100-185, 339-386): Y = act_out(act_in(X) W^T + b); X[M,K], W[Nout,K].",
    "emd_linear_fwd_tc": "emd_linear_fwd on the tensor cores: tcgen05.mma kind::tf32 with every operand split hi/lo (3xTF32, "
                         "fp32-class accuracy), accumulator in TMEM.  K % 4 == 0, K <= 136, Nout <= 64.",
    "emd_linear_bwd": "VJP of emd_linear_fwd: dX (may be NULL), dW, db; fixed-order reductions.",
    "emd_linear_bwd_tc": "emd_linear_bwd on the tensor cores (3xTF32): data gradient as a K-major GEMM against W^T, weight "
                         "gradient + bias gradient accumulated in TMEM over each CTA's row tiles from MN-major operands, "
                         "fixed-order reduction of the per-CTA partials.  Same workspace as emd_linear_bwd.",
    "emd_temb_fwd": "get_temporal_embed (deformation.py:208-221 / rigid.py:150-164): resample table[E,d] to `cur` rows and "
                    "sample at time t -> emb[d].  t is a DEVICE scalar (time + learnable time_offset, deformation.py:325-328).",
    "emd_temb_bwd": "VJP of emd_temb_fwd w.r.t. the table (ADDS into v_table) and t.",
    "emd_hexplane_fwd": "Replaces HexPlaneField.forward -> interpolate_ms_features -> 24 x F.grid_sample(bilinear, border, "
                        "align_corners=True) (S3Gaussian/scene/hexplane.py:73-106, 165-187; call deformation.py:187-199): "
                        "feat[N, S*F] = concat over scales of the product over the 6 planes (xy,xz,xt,yz,yt,zt).  planes: ONE flat "
                        "device buffer, plane (s,p) stored feature-last [H][W][F] at float offset plane_offsets[s*6+p] (HOST "
                        "array); reso: HOST int[S*4] grid size per coordinate (x,y,z,t) and scale; aabb: HOST float[6] = "
                        "{aabb[0], aabb[1]} (hexplane.py:19-20); t: DEVICE, one shared value (t_stride 0) or one per point (1).",
    "emd_hexplane_fwd_ld": "emd_hexplane_fwd with a row pitch: row n of the features starts at feat + n*ld floats (ld >= S*F, multiple "
                           "of 4), so the gather writes straight into the left columns of the deformation MLP's input "
                           "[N, S*F + E] -- the concatenation of deformation.py:205 without a copy pass.",
    "emd_hexplane_bwd_ld": "emd_hexplane_bwd reading v_feat with a row pitch of ld floats (the gradient of the MLP input, in place).",
    "emd_s3g_apply_blocks": "Rows of the `partial` array emd_s3g_apply_fwd writes for N Gaussians.",
    "emd_s3g_apply_fwd": "Residual application of deform_network.forward (S3Gaussian/scene/deformation.py:439-481, 484-527, flag set "
                         "--no_ds --no_dr): means = point + dx_c + dx_f, opac = opacity + do_c + do_f, shs = cat(dc, rest) + dshs_c + "
                         "dshs_f -- fused with GaussianModel.get_features' concatenation (scene/gaussian_model.py) and with the sums "
                         "the trainer's L1 regularisers need (train.py:240-305): partial[emd_s3g_apply_blocks(N)][6] = per-block "
                         "sums of |dx_c| |dx_f| |do_c| |do_f| |dshs_c| |dshs_f| (fixed order; the caller adds the rows).",
    "emd_s3g_apply_bwd": "VJP of emd_s3g_apply_fwd: coef is a DEVICE float[6], d loss / d sum_j; v_x = cotangent + coef_j * sign(x) "
                         "for the six residuals, v_dc / v_rest = the split of v_shs (the gradients of point and opacity are v_means "
                         "and v_opac themselves).  v_means / v_opac / v_shs may be NULL.",
    "emd_hexplane_bwd_workspace_bytes": "Workspace bytes of emd_hexplane_bwd (per-block partials of the shared-time gradient).",
    "emd_hexplane_bwd": "VJP of emd_hexplane_fwd: v_planes (layout of planes, ADDED into, caller zero-fills; 16-byte vector "
                        "reductions), v_pts[N,3] (may be NULL), v_t (N values written for t_stride 1, one value ADDED into for "
                        "t_stride 0 in a fixed order; may be NULL).",
    "emd_adam_max_tensors": "Tensors one emd_adam_step call can take (the table travels in the kernel parameter space).",
    "emd_adam_step": "torch.optim.Adam(groups, lr=0.0, eps=1e-15).step() as the reference configures it "
                     "(OmniRe/models/trainers/base.py:190-226, S3Gaussian/scene/gaussian_model.py:186-200): ONE launch over up to "
                     "emd_adam_max_tensors() fp32 tensors.  All arrays are HOST arrays [n_tensors]; params/grads/exp_avg/exp_avg_sq "
                     "hold DEVICE pointers; step[i] is the 1-based count after this update; grad_scale multiplies the gradients "
                     "(1/world_size after a sum all-reduce); L2 weight decay as in torch (g += wd * p).",
    "emd_densify_stats": "BasicTrainer.postprocess_per_train_step + VanillaGaussians.after_train (OmniRe/models/trainers/base.py:279-297, "
                         "OmniRe/models/gaussians/vanilla.py:163-191) for all Gaussian classes and the C cameras of a step in one "
                         "launch: xys_grad_norm += |grad * (W/2, H/2)|, vis_counts += 1, max_2Dsize = max(., radius / last_size) "
                         "where radius > 0; `first` reproduces the reference's first-call initialisation.",
    "emd_image_loss_partials_floats": "Floats of the `partials` scratch buffer of emd_image_loss_fwd.",
    "emd_image_loss_fwd": "Replaces the image terms of BasicTrainer.compute_losses (OmniRe/models/trainers/base.py:518-587: rgb L1, "
                          "pytorch_msssim SSIM, sky-opacity BCE / SafeBCE of models/losses.py:33-83, DepthLoss :91-172, opacity "
                          "entropy, kornia inverse-depth smoothness) together with render_fn's clamp (:415) and forward()'s sky "
                          "blend (:486-493); with blend=1 / ssim_pad=1 the S3Gaussian flavour (gaussian_renderer/__init__.py:299-300, "
                          "train.py:226, 348-363, utils/loss_utils.py:24-96).  rgb / depth / gt / sky are addressed through the "
                          "strides in cfg (gsplat renders[C,H,W,4]: rgb = base, depth = base + 3, pixel stride 4; diff_gauss CHW: "
                          "pixel stride 1, channel stride H*W); alpha, valid_mask (1 - egocar mask), sky_mask (1 = sky), lidar "
                          "are dense [C,H,W]; depth, sky, valid_mask, sky_mask, lidar may be NULL.  cfg and window (the 11 "
                          "normalised Gaussian taps) are HOST pointers.  Outputs: ssim_maps [C,3,3,H,W] and sums "
                          "[C,EMD_LOSS_SUMS] (kept for the backward), partials (scratch), terms [C,EMD_LOSS_TERMS] = weighted "
                          "loss terms per view; fixed-order reductions.",
    "emd_image_loss_bwd": "VJP of emd_image_loss_fwd: v_rgb (rgb strides), v_depth (depth strides; NULL iff depth is), v_alpha "
                          "[C,H,W], v_sky (sky strides; may be NULL) from v_terms [C,EMD_LOSS_TERMS] on the DEVICE -- the "
                          "cotangents the rasterizer backward consumes, every pixel written.",
    "emd_voxel_lbs_fwd": "Replaces VoxelDeformer.forward (OmniRe/models/modules.py:612-625: normalize :627-632 + 5-D F.grid_sample, "
                         "trilinear / border / align_corners=True) and get_voxel_weight's full-volume add (:575-582), queried by "
                         "SMPLTemplate.forward (OmniRe/models/human_body.py:174-179, `use_voxel_deformer: true`): out[B,V,J] from "
                         "xc[B,V,3].  base / corr: DEVICE volumes stored channel-LAST [B,D,H,W,J] (corr may be NULL); offset[B,3], "
                         "scale[B] DEVICE; ratio_dim: 0 = x, 1 = y, 2 = z (the reference's -1 - short_dim_dhw, modulo 3).",
    "emd_voxel_lbs_bwd": "VJP of emd_voxel_lbs_fwd: v_corr (layout of corr, ADDED into, caller zero-fills; 16-byte vector "
                         "reductions; may be NULL), v_xc[B,V,3] (may be NULL).",
    "emd_smpl_weight_grad": "d loss / d W[I,V,24] of emd_smpl_deform_fwd -- needed only when the LBS weights come from the voxel "
                            "deformer (human_body.py:174-179).  A[I,24,12]: the skinning matrices the forward wrote.",
    "emd_deform_input_fwd": "Input of ConditionalDeformNetwork as DeformableNodes.get_deformation builds it "
                            "(OmniRe/models/nodes/deformable.py:40-46, OmniRe/models/modules.py:341-366, 436-438): "
                            "x = means / instances_size[id, 2] * 2; columns [x, sin/cos(2^f x) ..., t, sin/cos(2^f t) ..., "
                            "instances_embedding[id]] written to out0 (row stride ld0) and, when out1 != NULL, to out1 (row "
                            "stride ld1: the head of the skip layer's operand).",
    "emd_dense_fwd": "One nn.Linear (+ F.relu) of ConditionalDeformNetwork.forward (modules.py:438-455) on strided operands: "
                     "Y[M,0:Nout] (row stride ldy) = act(X[M,0:K] (row stride ldx) W[Nout,K]^T + b).  fp32 SIMT GEMM.",
    "emd_dense_bwd_workspace_bytes": "Workspace bytes of emd_dense_bwd (split-K weight-gradient partials + bias partials).",
    "emd_dense_bwd": "VJP of emd_dense_fwd given dZ = dL/d(pre-activation): dX on the column window [col0, col0+ncols) of the "
                     "operand, multiplied by (mask > 0) where mask is the operand's producer's ReLU output (may be NULL); "
                     "dW[Nout,K], db[Nout] by fixed-order reductions.  dX, dW, db may each be NULL.",
    "emd_dense_tc_enabled": "1 when the EXPERIMENTAL tcgen05 (3xTF32) path of emd_dense_fwd / emd_dense_bwd's data gradient is "
                            "selected (default 0; EMD_DENSE_TC=1 or emd_dense_set_tc).  Written in round 1 after the GPU budget "
                            "was spent: compiled, not yet run on hardware (csrc/deform_net_tc.cu).",
    "emd_dense_set_tc": "Select (1) / deselect (0) the experimental tensor-core path; process-wide, for bring-up and measurement.",
    "emd_deform_apply_fwd": "deformable.py:57-68: means + d_xyz and get_quats + delta_quat (get_quats = quats / |quats|, "
                            "vanilla.py:142-146) from the heads' output d[N,dcols] (3, or 7 with the quaternion head).",
    "emd_deform_apply_bwd": "VJP of emd_deform_apply_fwd: v_d[N,dcols], v_means (NULL when stop_optimizing_canonical_xyz), v_quats.",
    "emd_deform_embed_grad_workspace_bytes": "Workspace bytes of emd_deform_embed_grad (per-chunk partial sums).",
    "emd_deform_embed_grad": "VJP of the instances_embedding[point_ids] gather (deformable.py:40): per-instance fixed-order sum of "
                             "the embedding-column gradients g0 (+ g1) over the instance-sorted index (order, seg_start); "
                             "max_points_per_instance is any upper bound on the largest instance (sizes the grid).",
    "emd_scan_workspace_bytes": "Workspace bytes of the scans for n elements.",
    "emd_cumsum_i32_i64": "Inclusive cumulative sum (torch.cumsum of tiles_per_gauss in gsplat's isect_tiles); total -> device scalar.",
    "emd_exclusive_scan_u32": "Exclusive scan (radix-sort tables); in-place allowed.",
    "emd_exclusive_scan_u8_u32": "Exclusive scan uint8 -> uint32 with the total written to *total_out (device; out + n makes `out` "
                                 "an (n+1)-long offsets array): per-pair gradient-entry counts -> entry_base of the raster backward.",
    "emd_isect_emit": "gsplat isect_tiles second pass: key = cam << (32+tile_n_bits) | tile << 32 | float_bits(depth), value = cam*N+gid.",
    "emd_radix_sort_workspace_bytes": "Workspace bytes of emd_radix_sort_pairs for n pairs.",
    "emd_radix_sort_pairs": "cub::DeviceRadixSort::SortPairs as gsplat/diff_gauss call it: stable, ascending, bits [begin,end). "
                            "*result_buffer = 0/1 tells which buffer holds the result.",
    "emd_isect_offsets": "gsplat isect_offset_encode / diff_gauss identifyTileRanges: first sorted index of every (camera, tile).",
    "emd_tile_order": "Tile ids ordered longest-list-first (scheduling only; results do not depend on it) plus the segment "
                      "bookkeeping of the segment-parallel forward / backward: seg_prefix (inclusive #segments along that order) "
                      "ckpt_base (first checkpoint slot per multi-segment tile) and cta_map ([emd_raster_max_ctas] x (tile id, "
                      "segment | segments << 16): the (tile, segment) each compositing CTA works on).",
    "emd_raster_max_ctas": "Upper bound on the compositing CTAs (tile segments) for P intersections: length of cta_map.",
    "emd_raster_segment_size": "Gaussians per segment of a tile's sorted list (the backward runs one CTA per segment).",
    "emd_raster_checkpoint_floats": "Floats per forward checkpoint (T and 4 accumulators of a tile's 256 pixels).",
    "emd_raster_segout_floats": "Floats per segment-output record of the segment-parallel forward (scratch).",
    "emd_raster_segment_slots": "Upper bound on the checkpoint / segment-output slots needed for P intersections.",
    "emd_raster_pack": "Packs per-(camera,Gaussian) mean2d/conic/opacity/colour(+depth) + the half extents of its alpha >= 1/255 box "
                       "into three float4 records.",
    "emd_raster_sort_records": "Writes the depth-sorted, per-tile-contiguous record stream the compositing kernels stage with bulk "
                               "asynchronous copies (cp.async.bulk + mbarrier): one 48-byte record per sorted (tile, Gaussian) "
                               "intersection -- mean, opacity, conic pre-scaled for exp2, four channels, the mask of the tile's 8x4 "
                               "pixel blocks its alpha >= 1/255 box reaches, and its gradient slot (emission-order index).  "
                               "isect_ids / flatten_ids are the SORTED keys / values; flavour 0 gsplat, 1 diff_gauss tile rectangles.",
    "emd_rasterize_fwd": "gsplat rasterize_to_pixels forward (RGB / +depth / expected depth, alpha, last_ids) "
                         "(reference call OmniRe/models/trainers/base.py:393-408).  Tiles longer than one segment are "
                         "composited one CTA per segment (transmittance pass, compositing pass, combine); ckpt "
                         "[emd_raster_segment_slots(P) x emd_raster_checkpoint_floats()] is kept for the backward, seg_out "
                         "[slots x emd_raster_segout_floats()] is scratch.  srecs = emd_raster_sort_records' stream.",
    "emd_rasterize_bwd_workspace_bytes": "Workspace bytes of emd_rasterize_bwd for P intersections.",
    "emd_rasterize_bwd": "gsplat rasterize_to_pixels backward: v_means2d (+abs), v_conics, v_colors, v_depths, v_opacities; "
                         "back to front from the forward's checkpoints over the same record stream; per warp a pixel-major pass "
                         "(alpha, T, two scalars per pair) and a Gaussian-major pass (each lane sums one Gaussian's 12 components "
                         "over the block's pixels); deterministic (no float atomics, no shuffle reductions per pair).",
}


def prototypes():
    out = {}
    for f in sorted((ROOT / "emd_b200" / "csrc").glob("*.cu")):
        src = f.read_text()
        for m in re.finditer(r'^extern "C" ([^{;]*?)\s*\{', src, re.S | re.M):
            p = " ".join(m.group(1).split())
            name = re.search(r"(\w+)\s*\(", p).group(1)
            out[name] = (p, f.name)
    return out


def abi_types():
    """Struct / constant definitions shared with the sources: the text between the ABI-types markers of the .cuh files."""
    out = []
    for f in sorted((ROOT / "emd_b200" / "csrc").glob("*.cuh")):
        m = re.search(r"// >>> ABI types[^\n]*\n(.*?)// <<< ABI types", f.read_text(), re.S)
        if m:
            out.append(f"/* ---- types (from emd_b200/csrc/{f.name}) ---- */\n" + m.group(1))
    return "\n".join(out)


def wrap(text, width=100, indent=" * "):
    words, lines, cur = text.split(), [], ""
    for w in words:
        if len(cur) + len(w) + 1 > width:
            lines.append(cur)
            cur = w
        else:
            cur = (cur + " " + w).strip()
    if cur:
        lines.append(cur)
    return "\n".join(indent + ln for ln in lines)


def render():
    protos = prototypes()
    out = ['''/* emd_b200.h -- C ABI of libemd_b200.so (sm_100a).  GENERATED by tools/gen_header.py from the
 * extern "C" definitions in emd_b200/csrc; do not edit the prototypes by hand.
 *
 * Boundary contract (SURVEY.md section 8b):
 *  - plain pointers and sizes only; every buffer (inputs, outputs, workspace) is device memory owned by the
 *    caller (PyTorch in the reference-side bindings); the library never allocates, frees or retains pointers;
 *  - every function returns 0 on success or a negative code (EMD_ERR_*); the text is emd_last_error_string();
 *    no exception crosses the boundary;
 *  - re-entrant, no implicit device synchronisation; kernels are enqueued on the cudaStream_t passed last
 *    (backward runs on autograd's thread -- pass torch.cuda.current_stream().cuda_stream);
 *  - tensors are dense row-major fp32 unless stated; [N,4] quaternion arrays must be 16-byte aligned;
 *  - there is no CPU implementation behind these symbols.
 */
#ifndef EMD_B200_H
#define EMD_B200_H
#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime_api.h>

#define EMD_OK 0
#define EMD_ERR_BAD_ARG -1
#define EMD_ERR_ALIGN -2
#define EMD_ERR_WORKSPACE -3
#define EMD_ERR_CUDA -4
#define EMD_ERR_UNSUPPORTED -5

''' + abi_types() + '''
#ifdef __cplusplus
extern "C" {
#endif
''']
    seen = set()
    for title, names in GROUPS:
        have = [n for n in names if n in protos]
        if not have:
            continue
        out.append(f"/* ---- {title} {'-' * max(4, 90 - len(title))} */\n")
        for n in have:
            seen.add(n)
            if n not in DOC:
                raise SystemExit(f"tools/gen_header.py: no DOC entry for {n}")
            out.append("/*\n" + wrap(DOC[n]) + f"\n * (defined in emd_b200/csrc/{protos[n][1]})\n */")
            out.append(protos[n][0] + ";\n")
    missing = sorted(set(protos) - seen)
    if missing:
        raise SystemExit(f"tools/gen_header.py: exported but not listed in GROUPS: {missing}")
    out.append("#ifdef __cplusplus\n}\n#endif\n#endif /* EMD_B200_H */\n")
    return "\n".join(out)


if __name__ == "__main__":
    text = render()
    if "--write" in sys.argv:
        (ROOT / "include").mkdir(exist_ok=True)
        (ROOT / "include" / "emd_b200.h").write_text(text)
        print("wrote include/emd_b200.h")
    else:
        print(text)
