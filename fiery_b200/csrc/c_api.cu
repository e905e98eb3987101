// extern "C" boundary of libfiery_b200.so (declared in include/fiery_b200.h).  Argument validation, TMA descriptor
// creation and launch dispatch; no torch types, no host<->device copies except where the header says so.
#include <stdarg.h>
#include <string.h>

#include "lift_plan.cuh"

namespace fiery {

static thread_local char g_last_error[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

// ---- TMA descriptor: head tensor viewed as (channels_total = images*head_channels, h, w), innermost w ------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
    // cuTensorMapEncodeTiled is a driver call and needs a current context; a thread that has made no runtime call yet
    // (e.g. the autograd engine's worker on its first backward) has none, so bind the primary context once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    static encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<encode_tiled_fn>(sym);
    return fn;
}

// One 4-D view (column, row, channel-within-image, image) of `channels` channels starting at channel `ch0` of every image.
static int encode_view(CUtensorMap* map, const void* head, long long n_images, int head_channels, int ch0, int channels,
                       int hh, int ww) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const size_t es = 4;
    const char* base = static_cast<const char*>(head) + static_cast<size_t>(ch0) * hh * ww * es;
    FIERY_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "head tensor (or its context slice) is not 16-byte aligned");
    FIERY_REQUIRE(hh <= 256, "feat_h=%d exceeds the TMA box limit of 256", hh);
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(ww), static_cast<cuuint64_t>(hh), static_cast<cuuint64_t>(channels),
                          static_cast<cuuint64_t>(n_images)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(ww) * es, static_cast<cuuint64_t>(ww) * hh * es,
                             static_cast<cuuint64_t>(ww) * hh * head_channels * es};
    cuuint32_t box[4] = {WT, static_cast<cuuint32_t>(hh), CH_BOX, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<char*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FIERY_OK;
}

int encode_head_maps(HeadMaps* maps, const void* head, int dtype, const LiftParams& P) {
    FIERY_REQUIRE(dtype == FIERY_DTYPE_F32, "TMA map: only fp32 head tensors are supported");
    const long long n_images = static_cast<long long>(P.frame0 + P.n_frames) * P.n_cameras;
    int rc = FIERY_OK;
    if (P.use_depth) {
        rc = encode_view(&maps->depth, head, n_images, P.head_channels, 0, P.D, P.hh, P.ww);
        if (rc != FIERY_OK) return rc;
    } else {
        memset(&maps->depth, 0, sizeof(CUtensorMap));
    }
    return encode_view(&maps->ctx, head, n_images, P.head_channels, P.use_depth ? P.D : 0, P.C, P.hh, P.ww);
}

// Tensor maps of the column-packed forward kernel (lift_fwd_cols.cu): the tile arrives as prob[row][depth][col4] and
// ctx[row][k][cl][col4] with channel = CPL*cl + k -- the dimension order of the maps is the shared-memory order, the strides do
// the permutation.
struct HeadMapsCols {
    CUtensorMap depth;
    CUtensorMap ctx;
};

int encode_head_maps_cols(HeadMapsCols* maps, const void* head, const LiftParams& P, int channels_per_lane) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const size_t es = 4;
    // the image coordinate of a tile is absolute (frame0 * n_cameras + ...): the map spans every image up to this launch's last
    const cuuint64_t ww = P.ww, hh = P.hh, n_images = static_cast<cuuint64_t>(P.frame0 + P.n_frames) * P.n_cameras;
    const cuuint64_t plane = ww * hh * es, image = plane * P.head_channels;
    FIERY_REQUIRE((reinterpret_cast<uintptr_t>(head) & 15) == 0 && (plane * (P.use_depth ? P.D : 0)) % 16 == 0,
                  "head tensor (or its context slice) is not 16-byte aligned");
    const int cpl = channels_per_lane;
    FIERY_REQUIRE(cpl >= 1 && P.C % cpl == 0 && P.C / cpl <= 256 && hh <= 256, "column kernel: unsupported head shape");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (P.use_depth) {
        cuuint64_t dims[4] = {ww, static_cast<cuuint64_t>(P.D), hh, n_images};
        cuuint64_t strides[3] = {plane, ww * es, image};
        cuuint32_t box[4] = {WT, 48, static_cast<cuuint32_t>(hh), 1};
        CUresult r = fn(&maps->depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(head), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled (depth, 4-D) failed with CUresult %d", (int)r);
    } else {
        memset(&maps->depth, 0, sizeof(CUtensorMap));
    }
    const char* ctx_base = static_cast<const char*>(head) + plane * (P.use_depth ? P.D : 0);
    cuuint64_t dims[5] = {ww, static_cast<cuuint64_t>(P.C / cpl), static_cast<cuuint64_t>(cpl), hh, n_images};
    cuuint64_t strides[4] = {cpl * plane, plane, ww * es, image};
    cuuint32_t box[5] = {WT, static_cast<cuuint32_t>(P.C / cpl), static_cast<cuuint32_t>(cpl), static_cast<cuuint32_t>(hh), 1};
    CUresult r = fn(&maps->ctx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<char*>(ctx_base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled (context, 5-D) failed with CUresult %d", (int)r);
    return FIERY_OK;
}

// NCHW output (frames, C, X*Y) as a 3-D map, innermost the pillar axis; the layout pass stores (box_pillars x C) blocks into it
int encode_bev_map(CUtensorMap* map, float* bev, long long pillars, int channels, int n_frames, int box_pillars) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    FIERY_REQUIRE((reinterpret_cast<uintptr_t>(bev) & 15) == 0 && pillars % 4 == 0, "BEV output is not 16-byte aligned / pitched");
    FIERY_REQUIRE(channels <= 256 && box_pillars <= 256, "BEV map: box too large");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(pillars), static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(n_frames)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(pillars) * 4, static_cast<cuuint64_t>(pillars) * channels * 4};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_pillars), static_cast<cuuint32_t>(channels), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, bev, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FIERY_E_CUDA, "cuTensorMapEncodeTiled (BEV output) failed with CUresult %d", (int)r);
    return FIERY_OK;
}

// launchers defined next to their kernels
int launch_lift_forward(const LiftParams& P, const void* head, int head_dtype, float* bev_out, void* scratch, const void* plan,
                        const float* warp_theta, const unsigned char* warp_copy,
                        cudaStream_t);
int launch_lift_plan(const LiftParams& P, unsigned char* tiles, unsigned char* touched, int want_streams, cudaStream_t stream);
int lift_chunk_frames(const LiftParams& P);
size_t lift_scratch_bytes(const LiftParams& P);
void lift_set_max_chunk_frames(int n);
void lift_set_timer(LaunchTimer* t);
int lift_forward_launches(const LiftParams& P);
size_t lift_backward_relayout_bytes(const LiftParams& P);
int launch_lift_backward(const LiftParams& P, const void* head, int head_dtype, float* workspace, const void* plan, cudaStream_t);
int launch_point_indices(const LiftParams& P, int64_t* idx_out, uint8_t* valid_out, int32_t* pillar_out, cudaStream_t);
int launch_compose(int n, const float* K, const float* E, float* combined, float* translation, cudaStream_t);
int launch_warp(int forward, int n_maps, int C, int H, int W, const float* a, long long a_stride, const float* theta,
                const unsigned char* copy_mask, float* b, long long b_stride, int nearest, cudaStream_t stream);
int launch_warp_theta(int n_seq, int T, int cumulative, const float* flow, float ex, float ey, float* theta,
                      unsigned char* copy_mask, cudaStream_t stream);
int launch_pack_conv_weights(const float* w_oihw, float* packed, cudaStream_t stream);
int launch_bev_conv(int n_frames, int H, int W, const float* x_nhwc, const float* w_packed, const float* scale, const float* shift,
                    int relu, float* y_nhwc, cudaStream_t stream);
int launch_depth_layer(int n_images, int pixels, int n_out, const void* feat, int dtype, const void* weight_padded, const float* bias,
                       float* head, cudaStream_t stream);
int vs_plan(int64_t n_rows, const int64_t* ranks, int32_t* seg, int64_t* host_n, cudaStream_t);
int vs_forward(int64_t n_rows, int channels, int64_t feat_stride, const float* feats, const int64_t* coords,
               const int32_t* seg, int64_t n_seg, float* sums, int64_t* coords_out, cudaStream_t);
int vs_backward(int64_t n_rows, int channels, const float* grad_sums, const int32_t* seg, float* grad_feats, cudaStream_t);

static int make_params(const fiery_lift_desc_t* d, const float* calib_a, const float* calib_b, const float* fu,
                       const float* fv, const float* fd, LiftParams& P) {
    FIERY_REQUIRE(d != nullptr, "desc is NULL");
    FIERY_REQUIRE(d->n_frames >= 0 && d->n_cameras >= 1, "bad n_frames=%d / n_cameras=%d", d->n_frames, d->n_cameras);
    FIERY_REQUIRE(d->depth_bins >= 1 && d->channels >= 1 && d->feat_h >= 1 && d->feat_w >= 1,
                  "bad head shape D=%d C=%d h=%d w=%d", d->depth_bins, d->channels, d->feat_h, d->feat_w);
    FIERY_REQUIRE(d->bev_x >= 1 && d->bev_y >= 1, "bad BEV size %dx%d", d->bev_x, d->bev_y);
    // the reference squeezes the Z axis and assigns into (C, X, Y) (fiery.py:268-271): only one height cell works
    FIERY_REQUIRE(d->bev_z == 1, "bev_z=%d: the reference path only supports a single height cell (fiery.py:269)", d->bev_z);
    FIERY_REQUIRE(static_cast<long long>(d->bev_x) * d->bev_y < (1ll << 31), "BEV grid too large");
    FIERY_REQUIRE(d->calib_mode == FIERY_CALIB_RAW || d->calib_mode == FIERY_CALIB_COMPOSED, "bad calib_mode %d", d->calib_mode);
    FIERY_REQUIRE(d->bev_layout == FIERY_BEV_NCHW || d->bev_layout == FIERY_BEV_NHWC, "bad bev_layout %d", d->bev_layout);
    FIERY_REQUIRE(d->n_frames == 0 || (calib_a && calib_b), "calibration pointer is NULL");
    FIERY_REQUIRE(fu && fv && fd, "frustum pointer is NULL");
    for (int a = 0; a < 3; ++a) FIERY_REQUIRE(d->bev_resolution[a] > 0.f, "bev_resolution[%d] must be positive", a);
    P.n_frames = d->n_frames; P.n_cameras = d->n_cameras; P.frame0 = 0;
    P.D = d->depth_bins; P.C = d->channels; P.hh = d->feat_h; P.ww = d->feat_w;
    P.n_wtiles = (d->feat_w + WT - 1) / WT;
    P.use_depth = d->use_depth_distribution ? 1 : 0;
    P.head_channels = d->channels + (P.use_depth ? d->depth_bins : 0);
    P.calib_mode = d->calib_mode;
    P.calib_a = calib_a; P.calib_b = calib_b; P.fu = fu; P.fv = fv; P.fd = fd;
    P.accum = nullptr; P.touched = nullptr; P.plan_tiles = nullptr; P.grad_bev = nullptr; P.grad_head = nullptr; P.head_f16 = nullptr;
    P.bev_layout = d->bev_layout;
    P.pillars = static_cast<long long>(d->bev_x) * d->bev_y;
    P.grid = make_grid_params(*d);
    return FIERY_OK;
}

}  // namespace fiery

using namespace fiery;

extern "C" {

FIERY_API int fiery_abi_version(void) { return FIERY_B200_ABI_VERSION; }

FIERY_API const char* fiery_last_error(void) { return g_last_error; }

// shape-only parameters for the size queries (no pointers)
static bool shape_params(const fiery_lift_desc_t* d, LiftParams& P) {
    if (!d || d->n_frames < 0 || d->n_cameras < 1 || d->feat_w < 1 || d->bev_x < 1 || d->bev_y < 1) return false;
    P = LiftParams{};
    P.n_frames = d->n_frames; P.n_cameras = d->n_cameras; P.C = d->channels; P.D = d->depth_bins;
    P.hh = d->feat_h; P.ww = d->feat_w;
    P.n_wtiles = (d->feat_w + WT - 1) / WT;
    P.bev_layout = d->bev_layout;
    P.pillars = static_cast<long long>(d->bev_x) * d->bev_y;
    return true;
}

FIERY_API size_t fiery_lift_plan_bytes(const fiery_lift_desc_t* d) {
    LiftParams P;
    if (!shape_params(d, P) || P.n_frames == 0) return 0;
    return plan_bytes(P.n_frames, P.n_cameras, P.n_wtiles, P.pillars);
}

FIERY_API int fiery_lift_plan(const fiery_lift_desc_t* desc, const float* calib_a, const float* calib_b, const float* frustum_u,
                              const float* frustum_v, const float* frustum_d, void* plan_out, void* stream) {
    LiftParams P;
    int rc = make_params(desc, calib_a, calib_b, frustum_u, frustum_v, frustum_d, P);
    if (rc != FIERY_OK) return rc;
    if (P.n_frames == 0) return FIERY_OK;
    FIERY_REQUIRE(plan_out != nullptr, "plan_out is NULL");
    const PlanView v = plan_view(plan_out, P.n_frames, P.n_cameras, P.n_wtiles, P.pillars, 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FIERY_CUDA_CHECK(cudaMemsetAsync(const_cast<unsigned char*>(v.touched), 0, static_cast<size_t>(P.n_frames) * P.pillars, st));
    return launch_lift_plan(P, const_cast<unsigned char*>(v.tiles), const_cast<unsigned char*>(v.touched), 1, st);
}

FIERY_API size_t fiery_lift_scratch_bytes(const fiery_lift_desc_t* d) {
    LiftParams P;
    if (!shape_params(d, P) || P.n_frames == 0) return 0;
    return lift_scratch_bytes(P);
}

FIERY_API void fiery_lift_set_max_chunk_frames(int32_t n) { lift_set_max_chunk_frames(n); }

FIERY_API int fiery_lift_forward_launches(const fiery_lift_desc_t* d) {
    LiftParams P;
    if (!shape_params(d, P) || P.n_frames == 0) return 0;
    return lift_forward_launches(P);
}

FIERY_API size_t fiery_lift_workspace_bytes(const fiery_lift_desc_t* d) {
    LiftParams P;
    if (!shape_params(d, P) || P.n_frames == 0) return 0;
    const size_t relayout = (lift_backward_relayout_bytes(P) + 127) & ~static_cast<size_t>(127);
    return relayout + static_cast<size_t>(P.n_frames) * P.n_cameras * P.n_wtiles * PLAN_TILE_BYTES;
}

FIERY_API int fiery_lift_forward(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                       const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev_out,
                       void* scratch, const void* plan, void* stream) {
    LiftParams P;
    int rc = make_params(desc, calib_a, calib_b, frustum_u, frustum_v, frustum_d, P);
    if (rc != FIERY_OK) return rc;
    if (P.n_frames == 0) return FIERY_OK;
    FIERY_REQUIRE(head && bev_out, "head / bev_out is NULL");
    return launch_lift_forward(P, head, desc->head_dtype, bev_out, scratch, plan, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_lift_forward_warped(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                                        const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev_out,
                                        void* scratch, const void* plan, const float* theta, const uint8_t* copy_mask, void* stream) {
    LiftParams P;
    int rc = make_params(desc, calib_a, calib_b, frustum_u, frustum_v, frustum_d, P);
    if (rc != FIERY_OK) return rc;
    if (P.n_frames == 0) return FIERY_OK;
    FIERY_REQUIRE(head && bev_out && theta && copy_mask, "head / bev_out / theta / copy_mask is NULL");
    FIERY_REQUIRE(desc->bev_layout == FIERY_BEV_NCHW, "the warped lift writes the NCHW layout only");
    return launch_lift_forward(P, head, desc->head_dtype, bev_out, scratch, plan, theta, copy_mask, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_lift_forward_timed(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                                       const float* frustum_u, const float* frustum_v, const float* frustum_d, float* bev_out,
                                       void* scratch, const void* plan, void* stream, int32_t max_launches, float* host_ms,
                                       int32_t* host_kind, int32_t* host_n_launches) {
    FIERY_REQUIRE(host_ms && host_kind && host_n_launches && max_launches >= 1, "timed forward: NULL output / no room");
    LaunchTimer t;
    t.cap = max_launches < LaunchTimer::MAX ? max_launches : LaunchTimer::MAX;
    for (int i = 0; i < 2 * t.cap; ++i) FIERY_CUDA_CHECK(cudaEventCreate(&t.ev[i]));
    lift_set_timer(&t);
    const int rc = fiery_lift_forward(desc, head, calib_a, calib_b, frustum_u, frustum_v, frustum_d, bev_out, scratch, plan, stream);
    lift_set_timer(nullptr);
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    *host_n_launches = t.n;
    for (int i = 0; i < t.n && rc == FIERY_OK && e == cudaSuccess; ++i) {
        host_kind[i] = t.kind[i];
        e = cudaEventElapsedTime(&host_ms[i], t.ev[2 * i], t.ev[2 * i + 1]);
    }
    for (int i = 0; i < 2 * t.cap; ++i) cudaEventDestroy(t.ev[i]);
    if (rc != FIERY_OK) return rc;
    FIERY_CUDA_CHECK(e);
    return FIERY_OK;
}

FIERY_API int fiery_lift_backward(const fiery_lift_desc_t* desc, const void* head, const float* calib_a, const float* calib_b,
                        const float* frustum_u, const float* frustum_v, const float* frustum_d, const float* grad_bev,
                        void* grad_head, float* workspace, const void* plan, void* stream) {
    LiftParams P;
    int rc = make_params(desc, calib_a, calib_b, frustum_u, frustum_v, frustum_d, P);
    if (rc != FIERY_OK) return rc;
    if (P.n_frames == 0) return FIERY_OK;
    FIERY_REQUIRE(head && grad_bev && grad_head, "head / grad_bev / grad_head is NULL");
    P.grad_bev = grad_bev;
    P.grad_head = static_cast<float*>(grad_head);
    FIERY_REQUIRE(workspace != nullptr || (desc->bev_layout == FIERY_BEV_NHWC && plan != nullptr),
                  "backward needs the workspace of fiery_lift_workspace_bytes() (NCHW grad_bev re-layout and/or the geometry plan)");
    return launch_lift_backward(P, head, desc->head_dtype, workspace, plan, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_lift_point_indices(const fiery_lift_desc_t* desc, const float* calib_a, const float* calib_b,
                             const float* frustum_u, const float* frustum_v, const float* frustum_d, int64_t* idx_out,
                             uint8_t* valid_out, int32_t* pillar_out, void* stream) {
    LiftParams P;
    int rc = make_params(desc, calib_a, calib_b, frustum_u, frustum_v, frustum_d, P);
    if (rc != FIERY_OK) return rc;
    return launch_point_indices(P, idx_out, valid_out, pillar_out, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_compose_calibration(int32_t n, const float* intrinsics, const float* extrinsics, float* combined_out,
                              float* translation_out, void* stream) {
    FIERY_REQUIRE(n >= 0, "n_matrices=%d", n);
    FIERY_REQUIRE(n == 0 || (intrinsics && extrinsics && combined_out && translation_out), "NULL pointer");
    return launch_compose(n, intrinsics, extrinsics, combined_out, translation_out, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_voxels_summing_plan(int64_t n_rows, const int64_t* ranks, int32_t* segment_of_row, int64_t* host_n_segments,
                              void* stream) {
    FIERY_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "n_rows=%lld out of range", (long long)n_rows);
    FIERY_REQUIRE(host_n_segments != nullptr, "host_n_segments is NULL");
    if (n_rows == 0) { *host_n_segments = 0; return FIERY_OK; }
    FIERY_REQUIRE(ranks && segment_of_row, "NULL pointer");
    return vs_plan(n_rows, ranks, segment_of_row, host_n_segments, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_voxels_summing_forward(int64_t n_rows, int32_t channels, int64_t feat_stride, const float* feats,
                                 const int64_t* coords, const int32_t* segment_of_row, int64_t n_segments,
                                 float* sums_out, int64_t* coords_out, void* stream) {
    FIERY_REQUIRE(n_rows >= 0 && channels >= 1 && feat_stride >= channels, "bad shape n_rows=%lld C=%d stride=%lld",
                  (long long)n_rows, channels, (long long)feat_stride);
    if (n_rows == 0) return FIERY_OK;
    FIERY_REQUIRE(feats && coords && segment_of_row && sums_out && coords_out, "NULL pointer");
    return vs_forward(n_rows, channels, feat_stride, feats, coords, segment_of_row, n_segments, sums_out, coords_out,
                      static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_voxels_summing_backward(int64_t n_rows, int32_t channels, const float* grad_sums, const int32_t* segment_of_row,
                                  float* grad_feats, void* stream) {
    FIERY_REQUIRE(n_rows >= 0 && channels >= 1, "bad shape");
    if (n_rows == 0) return FIERY_OK;
    FIERY_REQUIRE(grad_sums && segment_of_row && grad_feats, "NULL pointer");
    return vs_backward(n_rows, channels, grad_sums, segment_of_row, grad_feats, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_warp_features_forward(int32_t n_maps, int32_t channels, int32_t height, int32_t width, const float* x,
                                          int64_t x_map_stride, const float* theta, const uint8_t* copy_mask, float* out,
                                          int64_t out_map_stride, int32_t nearest, void* stream) {
    FIERY_REQUIRE(n_maps >= 0 && channels >= 1 && height >= 1 && width >= 1, "warp: bad shape");
    FIERY_REQUIRE(n_maps == 0 || (x && theta && out), "warp: NULL pointer");
    return launch_warp(1, n_maps, channels, height, width, x, x_map_stride, theta, copy_mask, out, out_map_stride, nearest ? 1 : 0,
                       static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_warp_features_backward(int32_t n_maps, int32_t channels, int32_t height, int32_t width, const float* grad_out,
                                           int64_t grad_out_map_stride, const float* theta, const uint8_t* copy_mask,
                                           float* grad_x, int64_t grad_x_map_stride, int32_t nearest, void* stream) {
    FIERY_REQUIRE(n_maps >= 0 && channels >= 1 && height >= 1 && width >= 1, "warp: bad shape");
    FIERY_REQUIRE(n_maps == 0 || (grad_out && theta && grad_x), "warp: NULL pointer");
    return launch_warp(0, n_maps, channels, height, width, grad_out, grad_out_map_stride, theta, copy_mask, grad_x, grad_x_map_stride,
                       nearest ? 1 : 0, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_warp_theta(int32_t n_sequences, int32_t T, int32_t cumulative, const float* flow, float spatial_extent_x,
                               float spatial_extent_y, float* theta, uint8_t* copy_mask, void* stream) {
    FIERY_REQUIRE(n_sequences >= 0 && (!cumulative || T >= 1), "warp_theta: bad shape");
    FIERY_REQUIRE(n_sequences == 0 || (flow && theta && (copy_mask || !cumulative)), "warp_theta: NULL pointer");
    return launch_warp_theta(n_sequences, T, cumulative ? 1 : 0, flow, spatial_extent_x, spatial_extent_y, theta, copy_mask,
                             static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_bev_conv_pack_weights(const float* weight_oihw, float* packed_out, void* stream) {
    FIERY_REQUIRE(weight_oihw && packed_out, "bev conv: NULL weight pointer");
    return launch_pack_conv_weights(weight_oihw, packed_out, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_bev_first_conv_forward(int32_t n_frames, int32_t height, int32_t width, const float* x_nhwc, const float* packed_weight,
                                           const float* scale, const float* shift, int32_t relu, float* y_nhwc, void* stream) {
    FIERY_REQUIRE(n_frames == 0 || (x_nhwc && packed_weight && y_nhwc), "bev conv: NULL pointer");
    return launch_bev_conv(n_frames, height, width, x_nhwc, packed_weight, scale, shift, relu ? 1 : 0, y_nhwc, static_cast<cudaStream_t>(stream));
}

FIERY_API int fiery_depth_layer_forward(int32_t n_images, int32_t pixels, int32_t n_out, const void* feat, int32_t dtype,
                                        const void* weight_padded, const float* bias, float* head_out, void* stream) {
    FIERY_REQUIRE(n_images == 0 || (feat && weight_padded && head_out), "depth layer: NULL pointer");
    return launch_depth_layer(n_images, pixels, n_out, feat, dtype, weight_padded, bias, head_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
