/*
 * fiery_b200 -- C ABI of the Blackwell-native (sm_100a) camera->BEV lift.
 *
 * The reference (wayveai/fiery) has no FFI: its "operator API" for this path is three Python call sites
 * (SURVEY.md section 8b).  Each entry point below replaces one of them and is what a binding for the path would
 * bind (ctypes stub: fiery_b200/_lib.py; reference-side patch: INTEGRATION.md):
 *
 *   fiery_lift_plan + fiery_lift_forward / fiery_lift_backward
 *       replace Fiery.get_geometry              fiery/models/fiery.py:193-208     (the plan: geometry once per batch)
 *             + Encoder.forward tail            fiery/models/encoder.py:98-102   (softmax x context outer product)
 *             + Fiery.projection_to_birds_eye_view   fiery/models/fiery.py:221-273
 *       i.e. the body of Fiery.calculate_birds_eye_view_features (fiery.py:275-286) after depth_layer.
 *   fiery_voxels_summing_forward / _backward
 *       replace VoxelsSumming.forward/backward  fiery/utils/geometry.py:283-314  (call site fiery.py:261)
 *   fiery_lift_point_indices
 *       exposes the integer voxel coordinates the reference computes at fiery.py:236-256 (for parity checks)
 *   fiery_compose_calibration
 *       exposes combined = R @ inverse(K), translation  (fiery.py:196,203)
 *   fiery_depth_layer_forward
 *       replaces Encoder.depth_layer (the head tensor's producer)  fiery/models/encoder.py:36,96   [SURVEY.md section 8f, next-3]
 *   fiery_bev_first_conv_forward
 *       replaces Decoder.first_conv (+ bn1 + relu in eval mode)  fiery/models/decoder.py:11,59-61   [SURVEY.md section 8f, next-2]
 *   fiery_warp_features_forward / _backward, fiery_warp_theta
 *       replace affine_grid + grid_sample inside warp_features  fiery/utils/geometry.py:219-220 (called from
 *       cumulative_warp_features geometry.py:225-253, call site fiery.py:143)   [SURVEY.md section 8f, next-1]
 *
 * Conventions: every pointer is a DEVICE pointer on the current CUDA device unless its name starts with
 * `host_`; tensors are dense row-major with the shapes given; `stream` is a cudaStream_t passed as void*
 * (NULL = default stream).  Calls enqueue work and return without synchronising unless stated.  Return value:
 * 0 on success, a negative FIERY_E_* code otherwise; fiery_last_error() gives the message for the calling thread.
 * There is no CPU implementation behind this ABI.
 */
#ifndef FIERY_B200_H_
#define FIERY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FIERY_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define FIERY_API __attribute__((visibility("default")))
#else
#define FIERY_API
#endif

enum {
    FIERY_OK = 0,
    FIERY_E_INVALID = -1,     /* bad argument / unsupported shape (message says which) */
    FIERY_E_CUDA = -2,        /* a CUDA runtime/driver call failed */
    FIERY_E_UNSUPPORTED = -3  /* valid request this build does not implement */
};

enum { FIERY_DTYPE_F32 = 0, FIERY_DTYPE_F16 = 1 };

/* How the camera calibration is supplied (fiery.py:193-205). */
enum {
    FIERY_CALIB_RAW = 0,       /* calib_a = intrinsics (B',n,3,3), calib_b = extrinsics (B',n,4,4); R @ K^-1 is
                                  composed on the device (LU with partial pivoting + solve, explicit fp32 order) */
    FIERY_CALIB_COMPOSED = 1   /* calib_a = combined (B',n,3,3) = R @ K^-1, calib_b = translation (B',n,3) */
};

/* Memory layout of the BEV tensor produced / consumed. Logical shape is always (B', C, X, Y) (fiery.py:225). */
enum {
    FIERY_BEV_NCHW = 0,        /* contiguous (B', C, X, Y) -- what the reference returns */
    FIERY_BEV_NHWC = 1         /* physical (B', X, Y, C): torch "channels_last" for the same logical tensor */
};

typedef struct fiery_lift_desc {
    int32_t n_frames;          /* B' = batch x time receptive field (fiery.py:278) */
    int32_t n_cameras;         /* n */
    int32_t depth_bins;        /* D, len(arange(*LIFT.D_BOUND)) (fiery.py:115) */
    int32_t channels;          /* C, MODEL.ENCODER.OUT_CHANNELS */
    int32_t feat_h, feat_w;    /* h, w = FINAL_DIM // DOWNSAMPLE (fiery.py:112) */
    int32_t bev_x, bev_y, bev_z;   /* bev_dimension (geometry.py:55); bev_z must be 1 (fiery.py:269) */
    float bev_offset[3];       /* bev_start_position - bev_resolution / 2, evaluated in fp32 (fiery.py:236) */
    float bev_resolution[3];   /* geometry.py:53 */
    float z_valid_lo, z_valid_hi; /* closed fp32 interval of (z - bev_offset[2]) for which
                                     0 <= trunc((z - bev_offset[2]) / bev_resolution[2]) < bev_z; computed
                                     exactly on the host (fiery_b200/geometry.py) */
    int32_t use_depth_distribution;  /* encoder.py:98: 1 = softmax x context, 0 = uniform depth (head has C channels) */
    int32_t head_dtype;        /* FIERY_DTYPE_* of the head tensor */
    int32_t calib_mode;        /* FIERY_CALIB_* */
    int32_t bev_layout;        /* FIERY_BEV_* */
} fiery_lift_desc_t;

FIERY_API int fiery_abi_version(void);
FIERY_API const char* fiery_last_error(void);

/*
 * Geometry plan.  Where every frustum point lands -- get_geometry (fiery.py:193-208) and the voxel index, mask and rank of
 * projection_to_birds_eye_view (fiery.py:236-256) -- depends on the calibration, the frustum and the BEV grid only, not on the head
 * tensor.  fiery_lift_plan evaluates it once for a batch of calibrations (the reference's exact fp32 operation order) and stores it
 * as pillar runs per (camera, feature column, depth), in the orders the forward and the backward kernel consume, plus one "receives
 * a point" byte per pillar.  fiery_lift_forward / fiery_lift_backward take `plan`:
 *   NULL      the geometry is evaluated inside the call (forward: in the tile kernel, under the latency of its loads; backward: by
 *             the plan kernel into the workspace);
 *   non-NULL  a buffer of fiery_lift_plan_bytes(desc) bytes filled by fiery_lift_plan with the SAME descriptor shape and the
 *             calibration of this batch -- the plan of a training step shared by its forward and backward, or one plan reused by
 *             every call while the camera rig is static.  It is only read.
 */
FIERY_API size_t fiery_lift_plan_bytes(const fiery_lift_desc_t* desc);
FIERY_API int fiery_lift_plan(const fiery_lift_desc_t* desc, const float* calib_a, const float* calib_b, const float* frustum_u,
                              const float* frustum_v, const float* frustum_d, void* plan_out, void* stream);

/* Bytes of zero-initialised device scratch fiery_lift_forward needs for FIERY_BEV_NCHW output (0 for NHWC): a
 * channel-last fp32 accumulator (chunk, X*Y, C) followed by one mark byte per pillar, where chunk <= B' is the number
 * of frames processed per pass (all of them unless the accumulator would exceed 1 GiB).
 * Invariant: the scratch must be all zero on entry; it is all zero again when the call's work completes. */
FIERY_API size_t fiery_lift_scratch_bytes(const fiery_lift_desc_t* desc);

/*
 * Forward lift.  head: (B'*n, D+C, h, w) [C channels if !use_depth_distribution], dtype head_dtype (FIERY_DTYPE_F32, or
 * FIERY_DTYPE_F16 for AMP heads: values are converted exactly to fp32 and all arithmetic is fp32, as autocast does to the
 * reference's softmax and outer product, encoder.py:99-100).
 * frustum_u (w), frustum_v (h), frustum_d (D): the separable factors of Fiery.frustum (fiery.py:109-128), fp32.
 * bev_out: (B',C,X,Y) fp32 in bev_layout.  For FIERY_BEV_NHWC the caller must pass bev_out zero-filled (the kernel
 * accumulates into it) and scratch may be NULL.
 */
FIERY_API int fiery_lift_forward(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                       const float* frustum_u, const float* frustum_v, const float* frustum_d,
                       float* bev_out, void* scratch, const void* plan, void* stream);

/*
 * The lift followed by cumulative_warp_features (fiery/models/fiery.py:140-146, fiery/utils/geometry.py:225-253) in one chain: same
 * arguments as fiery_lift_forward (bev_layout must be FIERY_BEV_NCHW), plus the sampling maps of fiery_warp_theta for the B' frames:
 * theta (B', 2, 3) fp32 and copy_mask (B') bytes (1 = the present frame of its sequence: written as is, geometry.py:243).  Frame f of
 * bev_out is the bilinear sample (affine_grid + grid_sample, zero padding, align_corners=False) of frame f's lifted features --
 * gathered straight from the channel-last accumulator by the layout pass, so the unwarped BEV is never written or re-read
 * (3 launches per frame group instead of 2, and none of the standalone warp's).  [SURVEY.md section 8f, next-1]
 */
FIERY_API int fiery_lift_forward_warped(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                                        const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev_out,
                                        void* scratch, const void* plan, const float* theta, const uint8_t* copy_mask, void* stream);

/* Number of kernel launches one fiery_lift_forward call with this descriptor issues (NHWC: the tile kernel; NCHW: tile kernel
 * + layout pass per frame group; groups of frames run as concurrent chains on internal streams that are forked from and
 * joined back into `stream` with events, so the call behaves like work queued on `stream` and can be captured in a graph). */
FIERY_API int fiery_lift_forward_launches(const fiery_lift_desc_t* desc);

/* fiery_lift_forward with every kernel launch bracketed by an event pair on its own stream (profiling / bench.py's roofline): runs
 * the call, synchronises `stream`, and writes per launch the duration in milliseconds and the kind (1 tile kernel, 2 layout pass)
 * into host arrays of max_launches entries; *host_n_launches receives the count.  Launches of different chains overlap, so the
 * durations are what each kernel took while the others were running -- the launches the step really runs. */
FIERY_API int fiery_lift_forward_timed(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                                       const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev_out,
                                       void* scratch, const void* plan, void* stream, int32_t max_launches, float* host_ms,
                                       int32_t* host_kind, int32_t* host_n_launches);

/* Test hook: caps the frames per pass (0 = default: as many as fit 1 GiB of scratch), so the multi-pass path can be exercised on
 * small batches.  Process-wide; changes fiery_lift_scratch_bytes accordingly. */
FIERY_API void fiery_lift_set_max_chunk_frames(int32_t n);

/* Bytes of device workspace fiery_lift_backward needs: the channel-last re-layout of an FIERY_BEV_NCHW grad_bev (0 for NHWC) plus
 * room for the plan records (used when plan is NULL).  Contents on entry/exit are irrelevant. */
FIERY_API size_t fiery_lift_workspace_bytes(const fiery_lift_desc_t* desc);

/*
 * Backward of the lift w.r.t. the head tensor (head_dtype must be FIERY_DTYPE_F32 in this build: widen an fp16 head first).
 * grad_bev: (B',C,X,Y) fp32 in bev_layout; grad_head: same shape and dtype as head, fully overwritten.  Calibration gets no gradient (geometry is integer, geometry.py:300).
 */
FIERY_API int fiery_lift_backward(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                        const float* frustum_u, const float* frustum_v, const float* frustum_d,
                        const float* grad_bev, void* grad_head, float* workspace, const void* plan, void* stream);


/*
 * Integer voxel coordinates of all N = n*D*h*w points per frame, in the reference's point order
 * (camera, depth, row, column) (fiery.py:233).  idx_out: (B',N,3) int64 = trunc((p - offset)/res) (fiery.py:236-237);
 * valid_out: (B',N) uint8 (fiery.py:240-247); pillar_out: (B',N) int32 = rank (fiery.py:252-256) or -1 -- taken
 * from the same device function the lift kernels use.  Any output pointer may be NULL.
 */
FIERY_API int fiery_lift_point_indices(const fiery_lift_desc_t* desc, const float* calib_a, const float* calib_b,
                             const float* frustum_u, const float* frustum_v, const float* frustum_d,
                             int64_t* idx_out, uint8_t* valid_out, int32_t* pillar_out, void* stream);

/* combined (B'*n,3,3) and translation (B'*n,3) from intrinsics (B'*n,3,3) and extrinsics (B'*n,4,4). */
FIERY_API int fiery_compose_calibration(int32_t n_matrices, const float* intrinsics, const float* extrinsics,
                              float* combined_out, float* translation_out, void* stream);

/*
 * VoxelsSumming.forward (geometry.py:286-302).  feats (Nm,C) fp32 with row stride feat_stride elements,
 * coords (Nm,3) int64, ranks (Nm) int64 sorted ascending.
 * Step 1 -- fiery_voxels_summing_plan: writes segment_of_row (Nm) int32 (the index into the output each row sums
 * into) and *host_n_segments = U.  Synchronises `stream` (the reference's boolean indexing at geometry.py:295 has
 * the same host sync).  Step 2 -- fiery_voxels_summing_forward: sums_out (U,C) fp32, coords_out (U,3) int64 (coords
 * of the last row of each run, geometry.py:295).
 * Backward (geometry.py:305-314): grad_feats[i] = grad_sums[segment_of_row[i]].
 */
FIERY_API int fiery_voxels_summing_plan(int64_t n_rows, const int64_t* ranks, int32_t* segment_of_row,
                              int64_t* host_n_segments, void* stream);
FIERY_API int fiery_voxels_summing_forward(int64_t n_rows, int32_t channels, int64_t feat_stride, const float* feats,
                                 const int64_t* coords, const int32_t* segment_of_row, int64_t n_segments,
                                 float* sums_out, int64_t* coords_out, void* stream);
FIERY_API int fiery_voxels_summing_backward(int64_t n_rows, int32_t channels, const float* grad_sums,
                                  const int32_t* segment_of_row, float* grad_feats, void* stream);

/*
 * BEV feature warping -- the heavy part of warp_features / cumulative_warp_features (fiery/utils/geometry.py:181-253, call
 * site fiery/models/fiery.py:143-146): affine_grid + grid_sample (align_corners=False, zero padding; bilinear, or nearest
 * if `nearest` != 0) of n_maps feature maps (C, H, W) fp32 under the affine maps theta (n_maps, 2, 3).  Map m starts at
 * x + m * x_map_stride (elements); channel planes are dense (H*W).  copy_mask (n_maps bytes, may be NULL): maps with a
 * non-zero byte are copied unchanged -- the present frame of a sequence (geometry.py:243).
 * backward: grad_x[m] = adjoint of the sampling applied to grad_out[m].  grad_x is OVERWRITTEN (no zero-fill needed): the adjoint
 * runs as a gather over the output pixels that sampled each source pixel (deterministic, no atomics); for maps that are no near-rigid
 * transforms (|det| < 1/4, strong scaling, non-finite) the search covers the whole image: exact, but slow.
 */
FIERY_API int fiery_warp_features_forward(int32_t n_maps, int32_t channels, int32_t height, int32_t width, const float* x,
                                          int64_t x_map_stride, const float* theta, const uint8_t* copy_mask, float* out,
                                          int64_t out_map_stride, int32_t nearest, void* stream);
FIERY_API int fiery_warp_features_backward(int32_t n_maps, int32_t channels, int32_t height, int32_t width, const float* grad_out,
                                           int64_t grad_out_map_stride, const float* theta, const uint8_t* copy_mask,
                                           float* grad_x, int64_t grad_x_map_stride, int32_t nearest, void* stream);

/*
 * The pose algebra in front of the sampling: flow (6-DoF vectors tx,ty,tz,rx,ry,rz) -> theta (., 2, 3) for the calls above.
 * cumulative != 0 replaces the loop of cumulative_warp_features (geometry.py:241-251: pose_vec2mat :145-160, the running
 * product flow[t] @ ... @ flow[T-2], mat2pose_vec :82-107, then the theta of warp_features :197-219): flow is
 * (n_sequences, T, 6), theta (n_sequences*T, 2, 3) and copy_mask (n_sequences*T bytes, required) are written; the last frame
 * of every sequence is flagged "copy".  cumulative == 0 is the theta of a plain warp_features call: flow (n_sequences, 6),
 * T ignored, copy_mask may be NULL.  spatial_extent_x/y as in geometry.py:205-206.
 */
FIERY_API int fiery_warp_theta(int32_t n_sequences, int32_t T, int32_t cumulative, const float* flow, float spatial_extent_x,
                               float spatial_extent_y, float* theta, uint8_t* copy_mask, void* stream);

/*
 * First BEV convolution on the tensor cores (tcgen05, TF32 operands, fp32 accumulation in tensor memory) -- Decoder.first_conv
 * (fiery/models/decoder.py:11,59): Conv2d(64, 64, kernel_size=7, stride=2, padding=3, bias=False), optionally followed by a per-channel
 * affine (bn1 folded for inference, decoder.py:60) and relu (decoder.py:61).  [SURVEY.md section 8f, next-2]
 * It consumes the lift's channel-last result directly: x_nhwc (B', H, W, 64) fp32 = FIERY_BEV_NHWC output of fiery_lift_forward;
 * y_nhwc (B', Ho, Wo, 64) fp32 with Ho = (H - 1) / 2 + 1, Wo likewise.  packed_weight: (49, 64, 64) = (tap r*7+s, out, in), made from
 * the module's (64, 64, 7, 7) weight by fiery_bev_conv_pack_weights.  scale / shift: 64 floats each, or both NULL.
 */
FIERY_API int fiery_bev_conv_pack_weights(const float* weight_oihw, float* packed_out, void* stream);
FIERY_API int fiery_bev_first_conv_forward(int32_t n_frames, int32_t height, int32_t width, const float* x_nhwc, const float* packed_weight,
                                           const float* scale, const float* shift, int32_t relu, float* y_nhwc, void* stream);

/*
 * Encoder.depth_layer on the tensor cores -- the 1x1 convolution 128 -> D + C that produces the head tensor
 * (fiery/models/encoder.py:36,96): head_out (n_images, n_out, pixels) fp32 = weight @ feat + bias, computed by tcgen05 (fp16 / bf16
 * operands under AMP, TF32 for fp32 features; fp32 accumulation).  feat: (n_images, 128, pixels) with pixels = h*w, dtype 0 fp32 /
 * 1 fp16 / 2 bf16; weight_padded: (128, 128) row-major in the SAME dtype, rows >= n_out zero; bias: n_out floats or NULL.  Writing the
 * fp32 head directly removes the widening pass an AMP step otherwise needs in front of the lift.  [SURVEY.md section 8f, next-3]
 */
FIERY_API int fiery_depth_layer_forward(int32_t n_images, int32_t pixels, int32_t n_out, const void* feat, int32_t dtype,
                                        const void* weight_padded, const float* bias, float* head_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FIERY_B200_H_ */
