/*
 * Plain-C restatement of the integer/geometry half of the reference's Lift-Splat path.  TEST INFRASTRUCTURE ONLY:
 * loaded (ctypes) by tests/ to cross-check the numpy/torch oracle and the CUDA kernels; never linked into the product.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -- no FMA contraction, every fp32 operation rounds separately)
 *
 * Each function cites the reference lines (wayveai/fiery @ fd03f16) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* fiery/models/fiery.py:199-205: p = (R @ inverse(K)) @ (u*d, v*d, d) + t, with `combined` = R @ inverse(K) supplied.
 * torch-CPU evaluates the batched 3x3 @ 3x1 as individually rounded fp32 mul/add in k = 0,1,2 order (pinned by
 * oracle/gen_golden.py); that order is written out here.
 * fiery.py:236-237: idx = ((p - (start - res/2)) / res).long()  -- true division, truncation toward zero.
 * fiery.py:240-247: keep = 0 <= idx < dim on all axes.
 * Layout of the N = n*D*h*w points per frame: (camera, depth, row, column), as produced by fiery.py:233. */
void oracle_voxel_indices(int n_frames, int n_cams, int D, int h, int w,
                          const float* u, const float* v, const float* depth,      /* frustum factors, fiery.py:115-122 */
                          const float* combined /* (B',n,3,3) */, const float* translation /* (B',n,3) */,
                          const float* offset /* start - res/2, fp32 */, const float* res, const int64_t* dim,
                          int64_t* idx_out /* (B',N,3) */, uint8_t* keep_out /* (B',N) */) {
    const int64_t N = (int64_t)n_cams * D * h * w;
    for (int f = 0; f < n_frames; ++f)
        for (int c = 0; c < n_cams; ++c) {
            const float* M = combined + ((int64_t)f * n_cams + c) * 9;
            const float* t = translation + ((int64_t)f * n_cams + c) * 3;
            for (int d = 0; d < D; ++d)
                for (int r = 0; r < h; ++r)
                    for (int col = 0; col < w; ++col) {
                        const float ud = u[col] * depth[d];                 /* fiery.py:202 */
                        const float vd = v[r] * depth[d];
                        const int64_t i = (int64_t)f * N + (((int64_t)c * D + d) * h + r) * w + col;
                        int ok = 1;
                        for (int a = 0; a < 3; ++a) {
                            float acc = M[a * 3 + 0] * ud;
                            acc = acc + M[a * 3 + 1] * vd;
                            acc = acc + M[a * 3 + 2] * depth[d];
                            const float p = acc + t[a];                     /* fiery.py:205 */
                            const float s = (p - offset[a]) / res[a];       /* fiery.py:236 */
                            int64_t k;
                            if (isnan(s) || s >= 9.2233720368547758e18f || s < -9.2233720368547758e18f)
                                k = INT64_MIN;                              /* what x86 cvttss2si gives torch's .long() */
                            else
                                k = (int64_t)s;                             /* truncation toward zero, fiery.py:237 */
                            idx_out[i * 3 + a] = k;
                            ok &= (k >= 0 && k < dim[a]);                   /* fiery.py:240-247 */
                        }
                        keep_out[i] = (uint8_t)ok;
                    }
        }
}

/* Direct pooling in double precision (oracle variant O3): the numerical ground truth for fiery.py:252-265 /
 * geometry.py:283-302 -- per-voxel sums of prob[d] * ctx[c] over the kept points; no sort, no prefix sums.
 * prob (B'n, D, h, w) already softmaxed (encoder.py:99), ctx (B'n, C, h, w); bev (B', C, X, Y) zero-initialised by the caller. */
void oracle_pool_exact(int n_frames, int n_cams, int D, int C, int h, int w, const double* prob, const double* ctx,
                       const int64_t* idx, const uint8_t* keep, int X, int Y, double* bev) {
    const int64_t N = (int64_t)n_cams * D * h * w, hw = (int64_t)h * w;
    for (int f = 0; f < n_frames; ++f)
        for (int c = 0; c < n_cams; ++c)
            for (int d = 0; d < D; ++d)
                for (int64_t px = 0; px < hw; ++px) {
                    const int64_t i = (int64_t)f * N + ((int64_t)c * D + d) * hw + px;
                    if (!keep[i]) continue;
                    const int64_t cell = idx[i * 3 + 0] * Y + idx[i * 3 + 1];
                    const int64_t img = (int64_t)f * n_cams + c;
                    const double p = prob[(img * D + d) * hw + px];
                    double* out = bev + (int64_t)f * C * X * Y + cell;
                    const double* cx = ctx + img * C * hw + px;
                    for (int ch = 0; ch < C; ++ch) out[(int64_t)ch * X * Y] += p * cx[(int64_t)ch * hw];
                }
}
