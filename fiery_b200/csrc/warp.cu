// BEV feature warping: the step right after the lift.
//
// Replaces the heavy part of warp_features / cumulative_warp_features (fiery/utils/geometry.py:181-253, call site
// fiery/models/fiery.py:143-146): torch.nn.functional.affine_grid + grid_sample (bilinear or nearest, zero padding,
// align_corners=False) of a (C, H, W) feature map under a 2x3 affine map theta.  The 6-DoF pose algebra that produces theta
// (pose_vec2mat, cumulative products, mat2pose_vec: a few 4x4 matrices per call) stays on the host side with the reference's
// own torch calls (fiery_b200/warp.py).
//
// HBM-bound gather: algorithmic bytes per map = read C*H*W*4 + write C*H*W*4.  One thread per output pixel and channel
// group; the sample position and the four weights are computed once per pixel, a warp covers 32 consecutive columns so the
// stores are full 128-byte lines and the four gathered rows are near-contiguous for the small rotations of ego motion.
#include "common.cuh"

namespace fiery {

constexpr int WARP_THREADS = 256;
constexpr int WARP_CH = 16;       // channels per thread

struct SamplePos {
    int off[4];      // element offsets of the 4 neighbours inside one channel plane (valid ones only)
    float w[4];      // bilinear weights, 0 for out-of-range neighbours (zero padding)
};

// affine_grid (align_corners=False): normalised pixel centres x_i = (2i+1)/W - 1; grid = theta @ (x, y, 1)
// grid_sample unnormalise (align_corners=False): ix = ((gx + 1) * W - 1) / 2
__device__ __forceinline__ void sample_coords(const float* __restrict__ th, int i, int j, int W, int H, float& ix, float& iy) {
    const float xs = (2.0f * i + 1.0f) / W - 1.0f;
    const float ys = (2.0f * j + 1.0f) / H - 1.0f;
    const float gx = fmaf(th[0], xs, fmaf(th[1], ys, th[2]));
    const float gy = fmaf(th[3], xs, fmaf(th[4], ys, th[5]));
    ix = ((gx + 1.0f) * W - 1.0f) * 0.5f;
    iy = ((gy + 1.0f) * H - 1.0f) * 0.5f;
}

__device__ __forceinline__ SamplePos make_sample(const float* __restrict__ th, int i, int j, int W, int H, int nearest) {
    float ix, iy;
    sample_coords(th, i, j, W, H, ix, iy);
    SamplePos s;
    if (nearest) {
        const float rx = nearbyintf(ix), ry = nearbyintf(iy);          // round half to even, like grid_sample 'nearest'
        const bool ok = rx >= 0.f && rx < W && ry >= 0.f && ry < H;
        s.off[0] = ok ? static_cast<int>(ry) * W + static_cast<int>(rx) : 0;
        s.w[0] = ok ? 1.f : 0.f;
        s.off[1] = s.off[2] = s.off[3] = 0;
        s.w[1] = s.w[2] = s.w[3] = 0.f;
        return s;
    }
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float fx = ix - x0f, fy = iy - y0f;
    const float wx[2] = {1.f - fx, fx}, wy[2] = {1.f - fy, fy};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float xf = x0f + (k & 1), yf = y0f + (k >> 1);
        const bool ok = xf >= 0.f && xf < W && yf >= 0.f && yf < H;        // also false for NaN / huge coordinates
        s.off[k] = ok ? static_cast<int>(yf) * W + static_cast<int>(xf) : 0;
        s.w[k] = ok ? wx[k & 1] * wy[k >> 1] : 0.f;
    }
    return s;
}

__global__ void __launch_bounds__(WARP_THREADS)
warp_forward_kernel(int C, int H, int W, const float* __restrict__ x, long long x_stride, const float* __restrict__ theta,
                    const unsigned char* __restrict__ copy_mask, float* __restrict__ out, long long out_stride, int nearest) {
    const int pix = blockIdx.x * WARP_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int map = blockIdx.z, c0 = blockIdx.y * WARP_CH;
    SamplePos s;
    if (copy_mask && copy_mask[map]) {          // the present frame of a sequence passes through unchanged (geometry.py:243)
        s.off[0] = pix; s.w[0] = 1.f;
        s.off[1] = s.off[2] = s.off[3] = 0; s.w[1] = s.w[2] = s.w[3] = 0.f;
    } else {
        s = make_sample(theta + map * 6, pix % W, pix / W, W, H, nearest);
    }
    const float* src = x + map * x_stride + static_cast<long long>(c0) * H * W;
    float* dst = out + map * out_stride + static_cast<long long>(c0) * H * W + pix;
    const int plane = H * W;
    const int nc = min(WARP_CH, C - c0);
#pragma unroll 4
    for (int c = 0; c < nc; ++c, src += plane, dst += plane) {
        float v = s.w[0] * __ldg(src + s.off[0]);
        v = fmaf(s.w[1], __ldg(src + s.off[1]), v);
        v = fmaf(s.w[2], __ldg(src + s.off[2]), v);
        v = fmaf(s.w[3], __ldg(src + s.off[3]), v);
        *dst = v;
    }
}

// adjoint: grad_x[neighbour] += w * grad_out[pixel]; grad_x is accumulated into (caller zero-fills)
__global__ void __launch_bounds__(WARP_THREADS)
warp_backward_kernel(int C, int H, int W, const float* __restrict__ gout, long long gout_stride, const float* __restrict__ theta,
                     const unsigned char* __restrict__ copy_mask, float* __restrict__ gx, long long gx_stride, int nearest) {
    const int pix = blockIdx.x * WARP_THREADS + threadIdx.x;
    if (pix >= H * W) return;
    const int map = blockIdx.z, c0 = blockIdx.y * WARP_CH;
    SamplePos s;
    if (copy_mask && copy_mask[map]) {
        s.off[0] = pix; s.w[0] = 1.f;
        s.off[1] = s.off[2] = s.off[3] = 0; s.w[1] = s.w[2] = s.w[3] = 0.f;
    } else {
        s = make_sample(theta + map * 6, pix % W, pix / W, W, H, nearest);
    }
    const float* g = gout + map * gout_stride + static_cast<long long>(c0) * H * W + pix;
    float* dst = gx + map * gx_stride + static_cast<long long>(c0) * H * W;
    const int plane = H * W;
    const int nc = min(WARP_CH, C - c0);
#pragma unroll 4
    for (int c = 0; c < nc; ++c, g += plane, dst += plane) {
        const float v = __ldg(g);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (s.w[k] != 0.f) atomicAdd(dst + s.off[k], s.w[k] * v);
    }
}

int launch_warp(int forward, int n_maps, int C, int H, int W, const float* a, long long a_stride, const float* theta,
                const unsigned char* copy_mask, float* b, long long b_stride, int nearest, cudaStream_t stream) {
    if (n_maps == 0) return FIERY_OK;
    FIERY_REQUIRE(static_cast<long long>(H) * W < (1ll << 31) && n_maps <= 65535, "warp: map too large / too many maps");
    const dim3 grid((H * W + WARP_THREADS - 1) / WARP_THREADS, (C + WARP_CH - 1) / WARP_CH, n_maps);
    if (forward) warp_forward_kernel<<<grid, WARP_THREADS, 0, stream>>>(C, H, W, a, a_stride, theta, copy_mask, b, b_stride, nearest);
    else warp_backward_kernel<<<grid, WARP_THREADS, 0, stream>>>(C, H, W, a, a_stride, theta, copy_mask, b, b_stride, nearest);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

}  // namespace fiery
