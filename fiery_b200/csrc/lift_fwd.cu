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
// Layout pass built on the copy engine.  A CTA takes FT_P consecutive pillars of one frame:
//   1. one thread per pillar reads the pillar's touched byte and, if set, fetches the pillar's 256-byte accumulator row with
//      a 1-D bulk copy (cp.async.bulk, completion on an mbarrier) -- only rows that received points are read, each as one
//      contiguous 256-byte burst;
//   2. the (pillar, channel) block is transposed shared -> shared with an XOR-swizzled lane mapping (reads hit bank
//      (c ^ lane) mod 32, writes bank lane: both conflict free); rows that were not fetched read as zero;
//   3. one tiled TMA store writes the (64 channels x FT_P pillars) block into the NCHW output (256 contiguous bytes per
//      channel row), and a 256-byte bulk copy of zeros per fetched row plus a byte store per mark restore the scratch
//      invariant of include/fiery_b200.h.
// Needs X*Y to be a multiple of 4 (16-byte row pitch of the output map).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FT_P = 64;
constexpr int FT_THREADS = 256;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(FT_THREADS)
finalize_tma_kernel(const __grid_constant__ CUtensorMap bev_map, float* __restrict__ accum, unsigned char* __restrict__ touched,
                    long long pillars, int tiles_per_frame, int frame_out0) {
    constexpr int C = 64;
    __shared__ __align__(128) float s_in[FT_P * C];      // [pillar][channel]: bulk-copy destination
    __shared__ __align__(128) float s_out[C * FT_P];     // [channel][pillar]: TMA store source
    __shared__ __align__(16) float s_zero[C];
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned char s_flag[FT_P];
    const int tid = threadIdx.x;
    const int frame = blockIdx.x / tiles_per_frame;
    const long long p0 = static_cast<long long>(blockIdx.x % tiles_per_frame) * FT_P;
    if (tid == 0) {
        tma_prefetch_desc(&bev_map);
        mbar_init(&bar, FT_P);
        fence_mbar_init();
    }
    if (tid >= FT_THREADS - C) s_zero[tid - (FT_THREADS - C)] = 0.f;
    __syncthreads();
    float* row = nullptr;
    unsigned char* mark = nullptr;
    if (tid < FT_P) {
        const long long p = p0 + tid;
        unsigned char f = 0;
        if (p < pillars) {
            mark = touched + static_cast<size_t>(frame) * pillars + p;
            f = *mark;
        }
        s_flag[tid] = f;
        if (f) {
            row = accum + (static_cast<size_t>(frame) * pillars + p) * C;
            mbar_arrive_expect_tx(&bar, C * 4);
            bulk_load_1d(s_in + tid * C, row, C * 4, &bar);
        } else {
            mbar_arrive(&bar);
        }
    }
    __syncthreads();                                    // s_flag
    mbar_wait(&bar, 0);                                 // every fetched row has landed
    {
        const int lane = tid & 31, w = tid >> 5;
        const int pl = (w & 1) * 32 + lane;             // pillar of this lane
        const int c_hi = (w >> 2) * 32;                 // 32-channel block; (w >> 1) & 1 picks its half
        const int c_lo = ((w >> 1) & 1) * 16;
        const bool have = s_flag[pl] != 0;
        const float* src = s_in + pl * C + c_hi;
        float* dst = s_out + c_hi * FT_P + pl;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = (c_lo + i) ^ lane;
            dst[c * FT_P] = have ? src[c] : 0.f;
        }
    }
    fence_proxy_async();                                // generic-proxy writes (s_out, s_zero) -> visible to the copy engine
    __syncthreads();
    if (tid == 0) tma_store_3d(&bev_map, s_out, static_cast<int>(p0), 0, frame_out0 + frame);
    if (row) {                                          // restore the all-zero scratch
        bulk_store_1d(row, s_zero, C * 4);
        *mark = 0;
    }
    if (tid == 0 || row) tma_store_commit_and_wait();   // the shared sources must outlive the copies
}

// ---------------------------------------------------------------------------------------------------------------------
// Streaming layout pass: persistent CTAs walk the (frame, 64-pillar) tiles with a three-deep software pipeline in registers:
//   touched bytes of tile i+2  ->  accumulator rows of tile i+1 (only where touched; 32-byte loads, lane = pillar, warp =
//   channel octet)  ->  tile i: 16 conflict-free STS.32 per thread build the (channel, pillar) block in shared memory (bank =
//   lane), the rows just read are re-zeroed with 32-byte stores, and ONE tiled TMA store writes the (64 x 64) block into the
//   NCHW output.  The output block is double buffered: the copy engine drains tile i-1 while tile i is assembled, and a CTA
//   never waits for DRAM latency in its loop because every load was issued one iteration earlier.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FS_P = 64;
constexpr int FS_THREADS = 256;

__global__ void __launch_bounds__(FS_THREADS)
finalize_stream_kernel(const __grid_constant__ CUtensorMap bev_map, float* __restrict__ accum, unsigned char* __restrict__ touched,
                       long long pillars, int tiles_per_frame, int n_tiles, int frame_out0) {
    constexpr int C = 64;
    __shared__ __align__(128) float s_out[2][C * FS_P];          // [channel][pillar], TMA store source
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;   // w: channels 8w .. 8w+7
    if (tid == 0) tma_prefetch_desc(&bev_map);

    // pillar index (frame * pillars + p) of this lane's two pillars of tile t, or -1 past the end of the frame / the tiles
    auto pillar_of_tile = [&](int t, int g) -> long long {
        if (t >= n_tiles) return -1;
        const int frame = t / tiles_per_frame;
        const long long p = static_cast<long long>(t - frame * tiles_per_frame) * FS_P + g * 32 + lane;
        return p < pillars ? static_cast<long long>(frame) * pillars + p : -1;
    };
    auto load_flags = [&](int t) -> unsigned {
        unsigned f = 0;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const long long gp = pillar_of_tile(t, g);
            if (gp >= 0) f |= static_cast<unsigned>(__ldcg(touched + gp)) << (8 * g);
        }
        return f;
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_rows = [&](int t, unsigned f, float4 (&v)[2][2]) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (f & (0xffu << (8 * g))) ldcg_256(accum + pillar_of_tile(t, g) * C + w * 8, v[g][0], v[g][1]);
            else v[g][0] = v[g][1] = z4;
        }
    };

    auto emit_tile = [&](int t, int buf, unsigned f, const float4 (&v)[2][2]) {
        float* dst = s_out[buf] + (w * 8) * FS_P + lane;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            dst[g * 32 + 0 * FS_P] = v[g][0].x; dst[g * 32 + 1 * FS_P] = v[g][0].y;
            dst[g * 32 + 2 * FS_P] = v[g][0].z; dst[g * 32 + 3 * FS_P] = v[g][0].w;
            dst[g * 32 + 4 * FS_P] = v[g][1].x; dst[g * 32 + 5 * FS_P] = v[g][1].y;
            dst[g * 32 + 6 * FS_P] = v[g][1].z; dst[g * 32 + 7 * FS_P] = v[g][1].w;
            if (f & (0xffu << (8 * g))) {                         // restore the all-zero scratch
                const long long gp = pillar_of_tile(t, g);
                st_zero_256(accum + gp * C + w * 8);
                if (w == 0) touched[gp] = 0;
            }
        }
        fence_proxy_async();                                      // my STS -> visible to the copy engine
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile t-1 has left the other buffer
        __syncthreads();
        if (tid == 0) {
            const int frame = t / tiles_per_frame;
            tma_store_3d(&bev_map, s_out[buf], (t - frame * tiles_per_frame) * FS_P, 0, frame_out0 + frame);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    };

    // The pipeline registers rotate with period 3 (flags) and 2 (rows): six tiles per trip keep every index a compile-time
    // constant, so nothing in flight is ever copied (a register move would wait for the load).
    const int step = gridDim.x;
    int t = blockIdx.x;
    unsigned fl[3];
    float4 rows[2][2][2];
    fl[0] = load_flags(t);
    fl[1] = load_flags(t + step);
    load_rows(t, fl[0], rows[0]);
    bool more = t < n_tiles;
#pragma unroll 1
    while (more) {
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (more) {
                fl[(u + 2) % 3] = load_flags(t + 2 * step);                   // consumed two tiles from now
                load_rows(t + step, fl[(u + 1) % 3], rows[(u + 1) % 2]);      // consumed by the next tile
                emit_tile(t, u & 1, fl[u % 3], rows[u % 2]);
                t += step;
                more = t < n_tiles;
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Wide streaming layout pass: like finalize_stream_kernel, but a tile is 256 consecutive pillars x ONE QUARTER of the channels,
// so the TMA store writes 16 channel rows of 1 KB each (instead of 64 rows of 256 B): four times longer DRAM bursts for
// the 10 MB/frame the pass has to write.  thread = pillar, 64-byte row quarter (two 32-byte loads, only where touched), 16
// conflict-free STS.32 (bank = lane).  The touched bytes are read-only here (four quarter tiles share them); the launcher
// clears the map afterwards with a memset node.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FW_P = 256;
constexpr int FW_C = 16;
constexpr int FW_THREADS = 256;

__global__ void __launch_bounds__(FW_THREADS)
finalize_wide_kernel(const __grid_constant__ CUtensorMap bev_map, float* __restrict__ accum, const unsigned char* __restrict__ touched,
                     long long pillars, int blocks_per_frame, int n_tiles, int frame_out0) {
    constexpr int C = 64;
    __shared__ __align__(128) float s_out[2][FW_C * FW_P];        // [channel][pillar], TMA store source
    const int tid = threadIdx.x;
    if (tid == 0) tma_prefetch_desc(&bev_map);

    // tile t = (frame, pillar block, quarter), quarter fastest: the four quarters of a block run back to back on neighbouring
    // CTAs, so the 256-byte rows are fetched from DRAM once
    auto row_of_tile = [&](int t) -> long long {                  // element offset of this thread's 64 bytes, or -1
        if (t >= n_tiles) return -1;
        const int q = t & 3, b = t >> 2;
        const int frame = b / blocks_per_frame;
        const long long p = static_cast<long long>(b - frame * blocks_per_frame) * FW_P + tid;
        return p < pillars ? (static_cast<long long>(frame) * pillars + p) * C + q * FW_C : -1;
    };
    auto load_flag = [&](int t) -> unsigned {
        const long long r = row_of_tile(t);
        return r >= 0 ? static_cast<unsigned>(__ldcg(touched + r / C)) : 0u;
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_row = [&](int t, unsigned f, float4 (&v)[4]) {
        if (f) {
            const float* src = accum + row_of_tile(t);
            ldcg_256(src, v[0], v[1]);
            ldcg_256(src + 8, v[2], v[3]);
        } else {
            v[0] = v[1] = v[2] = v[3] = z4;
        }
    };
    auto emit_tile = [&](int t, int buf, unsigned f, const float4 (&v)[4]) {
        float* dst = s_out[buf] + tid;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dst[(4 * k + 0) * FW_P] = v[k].x; dst[(4 * k + 1) * FW_P] = v[k].y;
            dst[(4 * k + 2) * FW_P] = v[k].z; dst[(4 * k + 3) * FW_P] = v[k].w;
        }
        if (f) {                                                  // restore the all-zero scratch
            float* row = accum + row_of_tile(t);
            st_zero_256(row);
            st_zero_256(row + 8);
        }
        fence_proxy_async();
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // tile t-1 has left the other buffer
        __syncthreads();
        if (tid == 0) {
            const int q = t & 3, b = t >> 2;
            const int frame = b / blocks_per_frame;
            tma_store_3d(&bev_map, s_out[buf], (b - frame * blocks_per_frame) * FW_P, q * FW_C, frame_out0 + frame);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    };

    const int step = gridDim.x;
    int t = blockIdx.x;
    unsigned fl[3];
    float4 rows[2][4];
    fl[0] = load_flag(t);
    fl[1] = load_flag(t + step);
    load_row(t, fl[0], rows[0]);
    bool more = t < n_tiles;
#pragma unroll 1
    while (more) {
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (more) {
                fl[(u + 2) % 3] = load_flag(t + 2 * step);
                load_row(t + step, fl[(u + 1) % 3], rows[(u + 1) % 2]);
                emit_tile(t, u & 1, fl[u % 3], rows[u % 2]);
                t += step;
                more = t < n_tiles;
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
int encode_bev_map(CUtensorMap* map, float* bev, long long pillars, int channels, int n_frames, int box_pillars, int box_channels);

// Side streams of the forward chains, created once per device (non-blocking).  Concurrent callers may share them: every call
// orders its work with its own events, so sharing only serialises.
constexpr int MAX_CHAINS = 4;
static int side_streams(int n, cudaStream_t* out) {
    static cudaStream_t pool[16][MAX_CHAINS - 1] = {};
    int dev = 0;
    FIERY_CUDA_CHECK(cudaGetDevice(&dev));
    FIERY_REQUIRE(dev >= 0 && dev < 16 && n <= MAX_CHAINS - 1, "side streams: device %d / n=%d out of range", dev, n);
    for (int i = 0; i < n; ++i) {
        if (!pool[dev][i]) FIERY_CUDA_CHECK(cudaStreamCreateWithFlags(&pool[dev][i], cudaStreamNonBlocking));
        out[i] = pool[dev][i];
    }
    return FIERY_OK;
}

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
    // NCHW: lift into the channel-last accumulator, then the layout pass; chunked only to bound the scratch footprint.
    // Within a chunk the frames are cut into up to `chains` groups, each a (tile kernel -> layout pass) chain on its own
    // stream: the layout pass of one group (DRAM-bound) then runs under the tile kernel of the next (issue-bound), and only
    // the last group's pass is exposed.  Frames are independent and every group owns its slice of the scratch, so the chains
    // share nothing; they are forked from and joined back into the caller's stream with events (capturable in a CUDA graph).
    const int chunk = lift_chunk_frames(P.n_frames, P.pillars, P.C);
    float* accum = scratch;                        // [accumulator floats of one chunk][one "touched" byte per pillar]
    unsigned char* touched = reinterpret_cast<unsigned char*>(scratch + static_cast<size_t>(chunk) * P.pillars * P.C);
    const int bpf = static_cast<int>((P.pillars + FIN_THREADS - 1) / FIN_THREADS);
    // 3 streaming pass (TMA store), 2 bulk-copy pass, 1 register-transposing pass, 0 one thread per pillar
    int pass = P.pillars % 4 == 0 ? 3 : 0;
    int ctas_per_sm = 6;
    int chains = 2;
    int min_tiles_per_sm = 2;
#ifdef FIERY_COLS_AB
    if (const char* e = getenv("FIERY_FINALIZE")) pass = P.pillars % 4 == 0 ? atoi(e) : 0;
    if (const char* e = getenv("FIERY_FINALIZE_CTAS")) ctas_per_sm = atoi(e);
    if (const char* e = getenv("FIERY_CHAINS")) chains = atoi(e);
    if (const char* e = getenv("FIERY_CHAIN_MIN_TILES")) min_tiles_per_sm = atoi(e);
#endif
    if (chains > MAX_CHAINS) chains = MAX_CHAINS;
    CUtensorMap bev_map;
    if (pass >= 2) {
        rc = pass == 4 ? encode_bev_map(&bev_map, bev_out, P.pillars, P.C, P.n_frames, FW_P, FW_C)
                       : encode_bev_map(&bev_map, bev_out, P.pillars, P.C, P.n_frames, FT_P, P.C);
        if (rc != FIERY_OK) return rc;
    }
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        FIERY_CUDA_CHECK(cudaGetDevice(&dev));
        FIERY_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    for (int f0 = 0; f0 < P.n_frames; f0 += chunk) {
        const int nf = (P.n_frames - f0 < chunk) ? P.n_frames - f0 : chunk;
        // a group should fill the machine at least once with tiles (2 per SM), or the split only adds launches
        int groups = chains;
        while (groups > 1 && static_cast<long long>(nf / groups) * P.n_cameras * P.n_wtiles < static_cast<long long>(min_tiles_per_sm) * n_sm) --groups;
        cudaStream_t side[MAX_CHAINS] = {stream};
        cudaEvent_t fork = nullptr;
        if (groups > 1) {
            rc = side_streams(groups - 1, side + 1);
            if (rc != FIERY_OK) return rc;
            FIERY_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
            FIERY_CUDA_CHECK(cudaEventRecord(fork, stream));
        }
        for (int g = 0; g < groups; ++g) {
            const int s0 = static_cast<int>(static_cast<long long>(nf) * g / groups);
            const int s1 = static_cast<int>(static_cast<long long>(nf) * (g + 1) / groups);
            cudaStream_t st = side[g];
            if (g > 0) FIERY_CUDA_CHECK(cudaStreamWaitEvent(st, fork, 0));
            Q.frame0 = f0 + s0;
            Q.n_frames = s1 - s0;
            Q.accum = accum + static_cast<size_t>(s0) * P.pillars * P.C;
            Q.touched = touched + static_cast<size_t>(s0) * P.pillars;
            float* out_g = bev_out + static_cast<size_t>(Q.frame0) * P.C * P.pillars;
            rc = launch_forward_cols(Q, head, st);
            if (rc != FIERY_OK) return rc;
            if (pass == 4) {
                const int bpf256 = static_cast<int>((P.pillars + FW_P - 1) / FW_P);
                const long long n_tiles = 4ll * bpf256 * Q.n_frames;
                FIERY_REQUIRE(n_tiles < (1ll << 30), "layout pass: too many tiles");
                const long long cap = static_cast<long long>(n_sm) * ctas_per_sm;
                finalize_wide_kernel<<<static_cast<unsigned>(n_tiles < cap ? n_tiles : cap), FW_THREADS, 0, st>>>(
                    bev_map, Q.accum, Q.touched, P.pillars, bpf256, static_cast<int>(n_tiles), Q.frame0);
                FIERY_CUDA_CHECK(cudaMemsetAsync(Q.touched, 0, static_cast<size_t>(Q.n_frames) * P.pillars, st));
            } else if (pass == 3) {
                const int tpf = static_cast<int>((P.pillars + FS_P - 1) / FS_P);
                const long long n_tiles = static_cast<long long>(tpf) * Q.n_frames;
                FIERY_REQUIRE(n_tiles < (1ll << 30), "layout pass: too many tiles");
                const long long cap = static_cast<long long>(n_sm) * ctas_per_sm;
                finalize_stream_kernel<<<static_cast<unsigned>(n_tiles < cap ? n_tiles : cap), FS_THREADS, 0, st>>>(
                    bev_map, Q.accum, Q.touched, P.pillars, tpf, static_cast<int>(n_tiles), Q.frame0);
            } else if (pass == 2) {
                const int tpf = static_cast<int>((P.pillars + FT_P - 1) / FT_P);
                finalize_tma_kernel<<<static_cast<unsigned>(tpf) * Q.n_frames, FT_THREADS, 0, st>>>(bev_map, Q.accum, Q.touched,
                                                                                                 P.pillars, tpf, Q.frame0);
            } else if (pass == 1) {
                const int bpf4 = static_cast<int>((P.pillars / 4 + FIN_THREADS - 1) / FIN_THREADS);
                finalize_clear_nchw_kernel<<<static_cast<unsigned>(bpf4) * Q.n_frames * 4, FIN_THREADS, 0, st>>>(
                    Q.accum, reinterpret_cast<unsigned*>(Q.touched), out_g, P.pillars, bpf4);
            } else {
                finalize_nchw_kernel<<<static_cast<unsigned>(bpf) * Q.n_frames, FIN_THREADS, 0, st>>>(Q.accum, Q.touched, out_g,
                                                                                                    P.pillars, bpf);
            }
            FIERY_CUDA_CHECK(cudaGetLastError());
            if (g > 0) {                                            // join the chain back into the caller's stream
                cudaEvent_t done;
                FIERY_CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
                FIERY_CUDA_CHECK(cudaEventRecord(done, st));
                FIERY_CUDA_CHECK(cudaStreamWaitEvent(stream, done, 0));
                FIERY_CUDA_CHECK(cudaEventDestroy(done));           // released once the work it marks has completed
            }
        }
        if (fork) FIERY_CUDA_CHECK(cudaEventDestroy(fork));
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
