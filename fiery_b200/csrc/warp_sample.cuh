// Sample positions of warp_features (fiery/utils/geometry.py:181-222): torch.nn.functional.affine_grid + grid_sample (bilinear or
// nearest, zero padding, align_corners=False) of one output pixel under a 2x3 affine map.  Shared by the standalone warp kernels
// (warp.cu) and the warp epilogue of the lift (lift_fwd.cu: finalize_warp_kernel), so both sample with the same arithmetic.
#pragma once
#include "common.cuh"

namespace fiery {

struct SamplePos {
    int off[4];      // element offsets of the 4 neighbours inside one channel plane (0 for out-of-range ones)
    float w[4];      // bilinear weights
    bool ok[4];      // neighbour inside the map (zero padding otherwise: never loaded)
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

__device__ __forceinline__ SamplePos make_sample(const float* __restrict__ theta, const unsigned char* __restrict__ copy_mask,
                                                 int map, int pix, int W, int H, int nearest) {
    SamplePos s;
#pragma unroll
    for (int k = 0; k < 4; ++k) { s.off[k] = 0; s.w[k] = 0.f; s.ok[k] = false; }
    if (copy_mask && copy_mask[map]) {          // the present frame of a sequence passes through unchanged (geometry.py:243)
        s.off[0] = pix; s.w[0] = 1.f; s.ok[0] = true;
        return s;
    }
    float ix, iy;
    sample_coords(theta + map * 6, pix % W, pix / W, W, H, ix, iy);
    if (nearest) {
        const float rx = nearbyintf(ix), ry = nearbyintf(iy);          // round half to even, like grid_sample 'nearest'
        const bool ok = rx >= 0.f && rx < W && ry >= 0.f && ry < H;
        s.off[0] = ok ? static_cast<int>(ry) * W + static_cast<int>(rx) : 0;
        s.w[0] = 1.f; s.ok[0] = ok;
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
        s.w[k] = wx[k & 1] * wy[k >> 1];
        s.ok[k] = ok;
    }
    return s;
}

}  // namespace fiery
