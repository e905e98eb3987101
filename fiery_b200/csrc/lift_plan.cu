// Geometry plan kernel (layout and rationale: lift_plan.cuh).
//
// get_geometry (fiery/models/fiery.py:193-208) + voxel index / mask / rank of every frustum point (fiery.py:236-256), evaluated with
// the reference's exact fp32 operation order (geometry.cuh) and reduced on the fly to pillar runs.  One CTA per tile (camera image x
// 4 feature-map columns), one thread per (depth, column) pair walking the image rows.
#include "lift_plan.cuh"

namespace fiery {

template <bool POW2>
__global__ void __launch_bounds__(PLAN_PAIRS)
lift_plan_kernel(const LiftParams P, unsigned char* __restrict__ tiles, unsigned char* __restrict__ touched) {
    __shared__ float s_cam[12];
    __shared__ float s_u[WT];
    __shared__ float s_v[PLAN_MAX_ROWS];
    __shared__ float s_d[48];
    __shared__ unsigned s_mask[PLAN_PAIRS];
    __shared__ int s_warp_sum[PLAN_PAIRS / 32];
    __shared__ int s_len[PLAN_STREAMS];
    __shared__ int s_soff[PLAN_STREAMS + 1];
    __shared__ int s_tmp[PLAN_MAX_ROWS * PLAN_PAIRS];     // [k][pair]: pillar of the pair's k-th run

    const int tid = threadIdx.x;
    const int wtile = blockIdx.x % P.n_wtiles;
    const int img_local = blockIdx.x / P.n_wtiles;      // (frame, camera) within this launch
    const int img = P.frame0 * P.n_cameras + img_local; // absolute: indexes the calibration
    const int frame = img_local / P.n_cameras;          // launch-local: indexes the touched map
    const int w0 = wtile * WT;
    const int hh = P.hh;
    unsigned char* rec = tiles + static_cast<size_t>(blockIdx.x) * PLAN_TILE_BYTES;

    if (tid < WT) s_u[tid] = (w0 + tid < P.ww) ? P.fu[w0 + tid] : 0.f;
    if (tid >= 32 && tid < 64) s_v[tid - 32] = P.fv[min(tid - 32, hh - 1)];
    if (tid >= 64 && tid < 64 + 48) s_d[tid - 64] = (tid - 64 < P.D) ? P.fd[tid - 64] : 0.f;
    if (tid == PLAN_PAIRS - 1) {                        // one lane composes R @ K^-1 (fiery.py:203)
        CameraTransform T;
        load_camera(P.calib_mode, P.calib_a, P.calib_b, img, T);
#pragma unroll
        for (int i = 0; i < 9; ++i) s_cam[i] = T.m[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) s_cam[9 + i] = T.t[i];
    }
    __syncthreads();

    // ---- runs of my pair ---------------------------------------------------------------------------------------------------
    const int pair = tid, d = pair >> 2, col = pair & 3;
    const bool dead = d >= P.D || w0 + col >= P.ww;
    unsigned mask = 0;
    int n = 1;
    if (dead) {
        s_tmp[pair] = -1;
    } else {
        CameraTransform T;
#pragma unroll
        for (int i = 0; i < 9; ++i) T.m[i] = s_cam[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) T.t[i] = s_cam[9 + i];
        const float offx = P.grid.off[0], offy = P.grid.off[1], offz = P.grid.off[2];
        const float kx = POW2 ? P.grid.inv_res[0] : P.grid.res[0], ky = POW2 ? P.grid.inv_res[1] : P.grid.res[1];
        const float Xf = static_cast<float>(P.grid.X), Yf = static_cast<float>(P.grid.Y);
        const float z_lo = P.grid.z_lo, z_hi = P.grid.z_hi;
        const int Y = P.grid.Y;
        unsigned char* tmap = touched ? touched + static_cast<size_t>(frame) * P.pillars : nullptr;
        const float depth = s_d[d];
        const ColumnTerms ct = column_terms(T, s_u[col], depth);
        int prev = 0;
        n = 0;
#pragma unroll 4
        for (int h = 0; h < hh; ++h) {
            float p[3];
            ego_point(T, ct, s_v[h], depth, p);                               // fiery.py:199-205
            const float ax = __fsub_rn(p[0], offx), ay = __fsub_rn(p[1], offy), az = __fsub_rn(p[2], offz);
            const float sx = POW2 ? __fmul_rn(ax, kx) : __fdiv_rn(ax, kx);    // fiery.py:236 (the scale is exact when res is 2^k)
            const float sy = POW2 ? __fmul_rn(ay, ky) : __fdiv_rn(ay, ky);
            const int rank = static_cast<int>(sx) * Y + static_cast<int>(sy); // truncation, fiery.py:237,252-256
            const int cur = select_pillar(sx, sy, az, Xf, Yf, z_lo, z_hi, rank);   // mask, fiery.py:240-247
            if (h == 0 || cur != prev) {
                if (h) mask |= 1u << h;
                s_tmp[n * PLAN_PAIRS + pair] = cur;
                ++n;
                if (tmap && cur >= 0) tmap[cur] = 1;
            }
            prev = cur;
        }
    }
    s_mask[pair] = mask;

    // ---- exclusive scan of the run counts over the pairs: offsets into runs[] -------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < PLAN_PAIRS / 32; ++w) base += (w < warp) ? s_warp_sum[w] : 0;
    const int off = base + incl - n;
    reinterpret_cast<unsigned*>(rec + PLAN_OFF_MASK)[pair] = mask;
    reinterpret_cast<unsigned short*>(rec + PLAN_OFF_OFF)[pair] = static_cast<unsigned short>(off);
    {
        int* runs = reinterpret_cast<int*>(rec + PLAN_OFF_RUNS) + off;
        for (int k = 0; k < n; ++k) runs[k] = s_tmp[k * PLAN_PAIRS + pair];
    }

    // ---- backward streams: per (row group, column, slot j) the runs of depths j, 4 + j, 8 + j, ... clipped to the row group ---------
    const int s = tid;
    const int rg = s >> 4, scol = (s >> 2) & 3, sj = s & 3;
    const int r_lo = plan_group_row(hh, rg & (PLAN_RG - 1)), r_hi = plan_group_row(hh, (rg & (PLAN_RG - 1)) + 1);
    const unsigned upto_lo = (2u << r_lo) - 1u;                                 // rows 0 .. r_lo
    const unsigned upto_hi = r_hi >= 32 ? 0xffffffffu : ((1u << r_hi) - 1u);    // rows 0 .. r_hi - 1
    if (s < PLAN_STREAMS) {
        int len = 2;                                                            // two pad entries
        for (int g = 0; g < 48 / PLAN_ND; ++g) {
            const unsigned m = s_mask[((g * PLAN_ND + sj) << 2) + scol];
            len += 1 + __popc(m & upto_hi & ~upto_lo);
        }
        s_len[s] = len;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < PLAN_STREAMS; ++i) { s_soff[i] = acc; acc += s_len[i]; }
        s_soff[PLAN_STREAMS] = acc;
    }
    if (tid == PLAN_PAIRS - 1) reinterpret_cast<unsigned*>(rec + PLAN_OFF_COUNTS)[0] = static_cast<unsigned>(base + incl);   // n_runs
    __syncthreads();
    if (s < PLAN_STREAMS) {
        reinterpret_cast<unsigned short*>(rec + PLAN_OFF_SOFF)[s] = static_cast<unsigned short>(s_soff[s]);
        int* out = reinterpret_cast<int*>(rec + PLAN_OFF_STREAMS) + s_soff[s];
        for (int g = 0; g < 48 / PLAN_ND; ++g) {
            const int p = ((g * PLAN_ND + sj) << 2) + scol;
            const unsigned m = s_mask[p];
            const int k0 = __popc(m & upto_lo);                                 // the run that contains row r_lo
            const int c = __popc(m & upto_hi & ~upto_lo);                       // runs that start inside the group
            for (int k = k0; k <= k0 + c; ++k) *out++ = s_tmp[k * PLAN_PAIRS + p];
        }
        out[0] = -1;
        out[1] = -1;
    }
    if (tid == 0) {
        unsigned* counts = reinterpret_cast<unsigned*>(rec + PLAN_OFF_COUNTS);
        counts[1] = static_cast<unsigned>(s_soff[PLAN_STREAMS]);                // n_stream
    }
}

int launch_lift_plan(const LiftParams& P, unsigned char* tiles, unsigned char* touched, cudaStream_t stream) {
    FIERY_REQUIRE(P.hh >= 1 && P.hh <= PLAN_MAX_ROWS, "feat_h=%d not supported by this build (<= %d)", P.hh, PLAN_MAX_ROWS);
    FIERY_REQUIRE(P.D >= 1 && P.D <= 48, "depth_bins=%d not supported by this build (1..48)", P.D);
    const long long n_tiles = static_cast<long long>(P.n_frames) * P.n_cameras * P.n_wtiles;
    if (n_tiles == 0) return FIERY_OK;
    if (P.grid.pow2[0] && P.grid.pow2[1])
        lift_plan_kernel<true><<<static_cast<unsigned>(n_tiles), PLAN_PAIRS, 0, stream>>>(P, tiles, touched);
    else
        lift_plan_kernel<false><<<static_cast<unsigned>(n_tiles), PLAN_PAIRS, 0, stream>>>(P, tiles, touched);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

}  // namespace fiery
