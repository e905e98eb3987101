// Forward camera->BEV lift for sm_100a, host side and the passes around the tile kernel (lift_fwd_cols.cu): the NCHW layout
// pass, the integer index dump and the calibration composition used by the parity checks, and the launchers.
//
// Replaces, per call: Fiery.get_geometry (fiery/models/fiery.py:193-208), the tail of Encoder.forward
// (fiery/models/encoder.py:98-102), and Fiery.projection_to_birds_eye_view incl. VoxelsSumming
// (fiery/models/fiery.py:221-273, fiery/utils/geometry.py:283-314).
#include "lift_tile.cuh"

namespace fiery {

// ---------------------------------------------------------------------------------------------------------------------
// Finalize for NCHW output: accum (B', X*Y, C) -> bev (B', C, X*Y).  One thread per pillar: a lane reads its pillar's 256-byte
// accumulator row as 16 independent 16-byte loads (its own two cache lines, so the sectors are fully used through L1),
// and the warp then writes one channel of 32 consecutive pillars per store instruction -- a full 128-byte line.  No
// shared-memory transpose (measured: the transposing variants are bound by 16-byte-per-lane scattered stores or by
// bank conflicts, tools/microbench/finalize_variants.cu).  Only ~1/3-1/2 of the pillars receive any point, and the lift
// kernel marks those in a byte map: unmarked pillars are written as zeros without touching the accumulator; marked
// rows and their marks are re-zeroed on the way, which restores the scratch invariant of include/fiery_b200.h.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
__global__ void __launch_bounds__(FIN_THREADS)
finalize_nchw_kernel(float* __restrict__ accum, unsigned char* __restrict__ flags, float* __restrict__ bev,
                     long long pillars, int blocks_per_frame) {
    constexpr int C = 64;
    const int frame = blockIdx.x / blocks_per_frame;
    const long long pl = static_cast<long long>(blockIdx.x % blocks_per_frame) * FIN_THREADS + threadIdx.x;
    if (pl >= pillars) return;
    unsigned char* f = flags + static_cast<size_t>(frame) * pillars + pl;
    float4 v[C / 4];
    if (*f) {
        float4* row = reinterpret_cast<float4*>(accum + (static_cast<size_t>(frame) * pillars + pl) * C);
#pragma unroll
        for (int q = 0; q < C / 4; ++q) v[q] = row[q];
#pragma unroll
        for (int q = 0; q < C / 4; ++q) row[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        *f = 0;
    } else {
#pragma unroll
        for (int q = 0; q < C / 4; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* dst = bev + static_cast<size_t>(frame) * C * pillars + pl;
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        dst[static_cast<size_t>(4 * q + 0) * pillars] = v[q].x;
        dst[static_cast<size_t>(4 * q + 1) * pillars] = v[q].y;
        dst[static_cast<size_t>(4 * q + 2) * pillars] = v[q].z;
        dst[static_cast<size_t>(4 * q + 3) * pillars] = v[q].w;
    }
}

// 32-byte global accesses (sm_100: LDG/STG.256): one full L2 sector per access
__device__ __forceinline__ void ldcg_256(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void st_zero_256(float* p) {
    asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" :: "l"(p), "f"(0.0f) : "memory");
}

// The layout pass and the re-zeroing of the scratch in ONE kernel.  A thread takes four consecutive pillars x one quarter
// of the channels (16): it loads its 4 x 64 bytes up front, transposes the 4 x 16 block in registers and writes sixteen
// 16-byte streaming stores; a warp covers 128 pillars, i.e. 512 contiguous bytes per channel row.  The 64 bytes a thread
// owns of every pillar row are moved as two full 32-byte sectors, so the zeroes that restore the scratch invariant are
// whole-sector stores right behind the loads (16-byte stores to the same place doubled the kernel's time: partial-sector
// writes; two separate re-zeroing kernels cost 14 us where this costs 11).  The touched byte of a pillar carries one bit
// per channel quarter so the four quarter-blocks clear their bit independently.  Needs X*Y to be a multiple of 4.
__global__ void __launch_bounds__(FIN_THREADS)
finalize_clear_nchw_kernel(float* __restrict__ accum, unsigned* __restrict__ flags32, float* __restrict__ bev,
                           long long pillars, int blocks_per_frame) {
    constexpr int C = 64;
    const int q = blockIdx.x & 3;                      // channel quarter
    const int b = blockIdx.x >> 2;
    const int frame = b / blocks_per_frame;
    const long long p0 = (static_cast<long long>(b % blocks_per_frame) * FIN_THREADS + threadIdx.x) * 4;
    if (p0 >= pillars) return;
    unsigned* fw = flags32 + (static_cast<size_t>(frame) * pillars + p0) / 4;
    const unsigned mine = __ldcg(fw) & (0x01010101u << q);     // bit q of each byte: this quarter still holds data there
    float* row = accum + (static_cast<size_t>(frame) * pillars + p0) * C + q * 16;      // pillar rows are 64 floats apart
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (mine & (0xffu << (8 * i))) {
            ldcg_256(row + i * C, v[i][0], v[i][1]);
            ldcg_256(row + i * C + 8, v[i][2], v[i][3]);
        } else {
            v[i][0] = v[i][1] = v[i][2] = v[i][3] = z4;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (mine & (0xffu << (8 * i))) {
            st_zero_256(row + i * C);
            st_zero_256(row + i * C + 8);
        }
    }
    if (mine) atomicAnd(fw, ~mine);
    float* dst = bev + (static_cast<size_t>(frame) * C + q * 16) * pillars + p0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float* d = dst + static_cast<size_t>(4 * k) * pillars;
        __stcs(reinterpret_cast<float4*>(d), make_float4(v[0][k].x, v[1][k].x, v[2][k].x, v[3][k].x));
        __stcs(reinterpret_cast<float4*>(d + pillars), make_float4(v[0][k].y, v[1][k].y, v[2][k].y, v[3][k].y));
        __stcs(reinterpret_cast<float4*>(d + 2 * pillars), make_float4(v[0][k].z, v[1][k].z, v[2][k].z, v[3][k].z));
        __stcs(reinterpret_cast<float4*>(d + 3 * pillars), make_float4(v[0][k].w, v[1][k].w, v[2][k].w, v[3][k].w));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Integer index dump (fiery.py:236-256) for parity checks; one thread per (frame, camera, depth, row, column) point.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void point_indices_kernel(const LiftParams P, int64_t* __restrict__ idx_out, uint8_t* __restrict__ valid_out,
                                     int32_t* __restrict__ pillar_out, long long n_points_total) {
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= n_points_total) return;
    const int w = static_cast<int>(gid % P.ww);
    long long r = gid / P.ww;
    const int h = static_cast<int>(r % P.hh); r /= P.hh;
    const int d = static_cast<int>(r % P.D); r /= P.D;
    const int cam_flat = static_cast<int>(r);
    CameraTransform T;
    load_camera(P.calib_mode, P.calib_a, P.calib_b, cam_flat, T);
    const float depth = P.fd[d];
    const ColumnTerms ct = column_terms(T, P.fu[w], depth);
    float p[3];
    ego_point(T, ct, P.fv[h], depth, p);
    const int pl = pillar_of(P.grid, p);
    if (pillar_out) pillar_out[gid] = pl;
    if (valid_out) valid_out[gid] = pl >= 0 ? 1 : 0;
    if (idx_out) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            // the reference's own expression: true division, then .long() (fiery.py:236-237)
            const float s = __fdiv_rn(__fsub_rn(p[a], P.grid.off[a]), P.grid.res[a]);
            idx_out[gid * 3 + a] = static_cast<int64_t>(s);
        }
    }
}

__global__ void compose_calibration_kernel(int n, const float* __restrict__ K, const float* __restrict__ E,
                                           float* __restrict__ combined, float* __restrict__ translation) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    CameraTransform T;
    compose_camera(K + i * 9, E + i * 16, T);
    for (int k = 0; k < 9; ++k) combined[i * 9 + k] = T.m[k];
    for (int k = 0; k < 3; ++k) translation[i * 3 + k] = T.t[k];
}

// ---------------------------------------------------------------------------------------------------------------------
// host launchers (called from c_api.cu)
// ---------------------------------------------------------------------------------------------------------------------
// Frames per launch for NCHW output.  Measured on B200 (profiles/r01_notes.md): chunks small enough to keep the accumulator
// L2-resident (3 frames, 31 MB) are slower end to end (152.9 us vs 122.7 us for 9 frames) -- the extra launches and the
// single-wave grids cost more than the saved HBM traffic -- so the chunk only bounds the scratch footprint (1 GiB).
int lift_chunk_frames(int n_frames, long long pillars, int channels) {
    const long long per_frame = pillars * channels * 4 + pillars;
    long long c = (1ll << 30) / (per_frame > 0 ? per_frame : 1);
    if (c < 1) c = 1;
    if (c > n_frames) c = n_frames;
    return static_cast<int>(c < 1 ? 1 : c);
}

int launch_forward_cols(const LiftParams& P, const void* head, cudaStream_t stream);

int launch_lift_forward(const LiftParams& P, const void* head, int head_dtype, float* bev_out, float* scratch,
                        cudaStream_t stream) {
    FIERY_REQUIRE(head_dtype == FIERY_DTYPE_F32, "head dtype %d not supported by this build (fp32 only)", head_dtype);
    FIERY_REQUIRE(P.C == 64, "channels=%d not supported by this build (C must be 64)", P.C);
    FIERY_REQUIRE(P.D >= 1 && P.D <= 48, "depth_bins=%d not supported by this build (1..48)", P.D);
    FIERY_REQUIRE(P.ww % 4 == 0, "feat_w=%d must be a multiple of 4 (TMA row pitch must be 16-byte aligned)", P.ww);
    int rc = FIERY_OK;
    LiftParams Q = P;
    if (P.bev_layout == FIERY_BEV_NHWC) {          // the caller's zero-filled channel-last tensor is the accumulator
        Q.accum = bev_out;
        Q.touched = nullptr;
        Q.frame0 = 0;
        return launch_forward_cols(Q, head, stream);
    }
    // NCHW: lift into the channel-last accumulator, then the layout pass; chunked only to bound the scratch footprint
    const int chunk = lift_chunk_frames(P.n_frames, P.pillars, P.C);
    Q.accum = scratch;                             // [accumulator floats of one chunk][one "touched" byte per pillar]
    Q.touched = reinterpret_cast<unsigned char*>(scratch + static_cast<size_t>(chunk) * P.pillars * P.C);
    const int bpf = static_cast<int>((P.pillars + FIN_THREADS - 1) / FIN_THREADS);
    for (int f0 = 0; f0 < P.n_frames; f0 += chunk) {
        Q.frame0 = f0;
        Q.n_frames = (P.n_frames - f0 < chunk) ? P.n_frames - f0 : chunk;
        rc = launch_forward_cols(Q, head, stream);
        if (rc != FIERY_OK) return rc;
        if (P.pillars % 4 == 0) {
            const int bpf4 = static_cast<int>((P.pillars / 4 + FIN_THREADS - 1) / FIN_THREADS);
            finalize_clear_nchw_kernel<<<static_cast<unsigned>(bpf4) * Q.n_frames * 4, FIN_THREADS, 0, stream>>>(
                Q.accum, reinterpret_cast<unsigned*>(Q.touched), bev_out + static_cast<size_t>(f0) * P.C * P.pillars, P.pillars, bpf4);
        } else {
            finalize_nchw_kernel<<<static_cast<unsigned>(bpf) * Q.n_frames, FIN_THREADS, 0, stream>>>(
                Q.accum, Q.touched, bev_out + static_cast<size_t>(f0) * P.C * P.pillars, P.pillars, bpf);
        }
        FIERY_CUDA_CHECK(cudaGetLastError());
    }
    return FIERY_OK;
}

int launch_point_indices(const LiftParams& P, int64_t* idx_out, uint8_t* valid_out, int32_t* pillar_out, cudaStream_t stream) {
    const long long total = static_cast<long long>(P.n_frames) * P.n_cameras * P.D * P.hh * P.ww;
    if (total == 0) return FIERY_OK;
    const int threads = 256;
    point_indices_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
        P, idx_out, valid_out, pillar_out, total);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

int launch_compose(int n, const float* K, const float* E, float* combined, float* translation, cudaStream_t stream) {
    if (n == 0) return FIERY_OK;
    compose_calibration_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, K, E, combined, translation);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

}  // namespace fiery
