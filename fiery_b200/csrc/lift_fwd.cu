// Forward camera->BEV lift for sm_100a: geometry + depth softmax + depth x context outer product + pillar pooling in
// one kernel, the frustum volume (124 MB/frame in the reference, fiery/models/encoder.py:100) never leaves the SM.
//
// Replaces, per call: Fiery.get_geometry (fiery/models/fiery.py:193-208), the tail of Encoder.forward
// (fiery/models/encoder.py:98-102), and Fiery.projection_to_birds_eye_view incl. VoxelsSumming
// (fiery/models/fiery.py:221-273, fiery/utils/geometry.py:283-314).
//
// Observation the kernel is built on: at fixed (camera, column, depth) the h image rows of a column fall into one
// BEV pillar, or a handful, because Z is collapsed (Z_BOUND has one cell) and cameras are close to level.  So the
// reference's global argsort + cumsum (fiery.py:257, geometry.py:289) becomes a register-resident *segmented* sum along
// the image column: a thread owns 8 depths x 4 channels of one column, walks the rows, and only when the pillar changes
// (a precomputed change bit, rare) does it flush its partial sum with one 16-byte vector reduction
// (red.global.add.v4.f32) into a channel-last BEV accumulator.  Sixteen lanes of a half-warp cover the 64 channels of
// a pillar, so each flush is two full 128-byte lines.  ~17k column segments per frame reach L2 instead of 453k points.
#include <cstdlib>
#include <cstring>

#include "lift_tile.cuh"

namespace fiery {

// packed fp32x2 FMA (SASS FFMA2): d.lo += a * b.lo, d.hi += a * b.hi with the scalar a broadcast to both halves
__device__ __forceinline__ void ffma2_bcast(unsigned long long& acc, float a, unsigned long long b) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(b));
}

// acc = 0 where `bit` is set, as predicated moves: straight-line code for the compiler (a C++ conditional assignment
// makes ptxas keep two copies of all accumulator registers across the hot loop)
__device__ __forceinline__ void clear_if(unsigned long long& a, unsigned long long& b, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p mov.b64 %0, 0;\n\t@p mov.b64 %1, 0;\n\t}" : "+l"(a), "+l"(b) : "r"(bit));
}

__device__ __forceinline__ void flush_pair(float* dst, unsigned long long lo, unsigned long long hi) {
    float a, b, c, d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(lo));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c), "=f"(d) : "l"(hi));
    red_add_v4(dst, a, b, c, d);
}

// the same under a predicate (straight-line code: run ends are tested warp-uniformly, the half-warp that owns the run flushes)
__device__ __forceinline__ void flush_pair_if(char* dst, unsigned long long lo, unsigned long long hi, unsigned bit) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .f32 a, b, c, d;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 {a, b}, %1;\n\tmov.b64 {c, d}, %2;\n\t"
                 "@p red.global.add.v4.f32 [%0], {a, b, c, d};\n\t}" :: "l"(dst), "l"(lo), "l"(hi), "r"(bit) : "memory");
}

template <int DBLKS>
__global__ void __launch_bounds__(64 * DBLKS, 3)
lift_forward_kernel(const __grid_constant__ HeadMaps head_maps, const LiftParams P) {
    using TL = TileLayout<DBLKS>;
    constexpr int DPAD = TL::DPAD;
    constexpr int PS = TL::PS;
    extern __shared__ __align__(128) unsigned char smem[];
    const TL L(P.hh, P.C);

    // tile coordinates: blockIdx.x = (chunk-local frame * n + camera) * n_wtiles + wtile
    const int wtile = blockIdx.x % P.n_wtiles;
    const int img_local = blockIdx.x / P.n_wtiles;    // (frame, camera) within this launch's chunk of frames
    const int img = P.frame0 * P.n_cameras + img_local;
    const int frame = img_local / P.n_cameras;        // chunk-local: indexes the accumulator
    const int w0 = wtile * WT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    if (tid == 0) {
        tma_prefetch_desc(&head_maps.depth);
        tma_prefetch_desc(&head_maps.ctx);
        mbar_init(bar, 1);
        fence_mbar_init();
        issue_tile_loads<DBLKS>(P, L, smem, &head_maps, img, w0);
    }
    stage_constants<DBLKS>(P, L, smem, img, w0);
    __syncthreads();                                  // mbarrier init + constants visible
    // one lane composes R @ K^-1 while the TMA is in flight (its latency is longer than the composition); the result
    // is first read after the barriers inside transform_tile
    if (tid == 64 * DBLKS - 1) stage_camera<DBLKS>(P, L, smem, img);
    mbar_wait(bar, 0);                                // head tile has landed
    transform_tile<DBLKS>(P, L, smem);                // softmax + transposes (two barriers inside; camera visible after)
    stage_pillars<DBLKS>(P, L, smem, w0);
    __syncthreads();
    stage_events<DBLKS>(L, smem, P.touched ? P.touched + static_cast<size_t>(frame) * P.pillars : nullptr);
    __syncthreads();

    // ---- pooling: thread = (column wt, depth block dblk of 8, channel group cg of 4) ----------------------------------
    const int half = lane >> 4;
    const int unit = warp * 2 + half;
    const int wt = unit / DBLKS, dblk = unit % DBLKS;
    const int cg = lane & 15;
    const int hh = L.hh;
    const float* prob = reinterpret_cast<const float*>(smem + L.off_prob) + (wt * hh) * PS + dblk * 8;
    const float* ctx = reinterpret_cast<const float*>(smem + L.off_ctx) + (wt * hh) * L.C + cg * 4;
    const int* pillar = reinterpret_cast<const int*>(smem + L.off_pillar) + (wt * hh) * DPAD + dblk * 8;
    char* out = reinterpret_cast<char*>(P.accum + static_cast<size_t>(frame) * P.pillars * P.C + cg * 4);
    // lane r keeps the warp's event word of row r; every row broadcasts its word, so all run-end branches below are
    // warp-uniform (the two half-warps own different depths and would otherwise diverge on every event)
    const unsigned ev_mine = reinterpret_cast<const unsigned*>(smem + L.off_ev)[warp * 32 + lane];
    const unsigned own_shift = 8 * half;

    unsigned long long acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0ull;

    const float* pp = prob;
    const float* cp = ctx;
    const int* plp = pillar - DPAD;                  // row h-1
#pragma unroll 2
    for (int h = 0; h < hh; ++h, pp += PS, cp += L.C, plp += DPAD) {
        const unsigned ev = __shfl_sync(0xffffffffu, ev_mine, h);
        const unsigned mw = (ev | (ev >> 8)) & 0xffu;                // depth slots that end a run in either half-warp
        if (mw) {
            const unsigned own = ev >> own_shift;                     // bits 0-7: my runs that end, bits 16-23: ... and flush
#pragma unroll
            for (int nib = 0; nib < 2; ++nib) {
                if (mw & (0xfu << (4 * nib))) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * nib + jj;
                        if (mw & (1u << j)) {
                            const unsigned pl = static_cast<unsigned>(plp[j]);
                            flush_pair_if(out + static_cast<size_t>(pl) * (64 * 4), acc[j][0], acc[j][1], own & (0x10000u << j));
                            clear_if(acc[j][0], acc[j][1], own & (1u << j));
                        }
                    }
                }
            }
        }
        const float4 p0 = *reinterpret_cast<const float4*>(pp);
        const float4 p1 = *reinterpret_cast<const float4*>(pp + 4);
        const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(cp);
        // depth x context outer product (encoder.py:100), summed along the column
        ffma2_bcast(acc[0][0], p0.x, c.x); ffma2_bcast(acc[0][1], p0.x, c.y);
        ffma2_bcast(acc[1][0], p0.y, c.x); ffma2_bcast(acc[1][1], p0.y, c.y);
        ffma2_bcast(acc[2][0], p0.z, c.x); ffma2_bcast(acc[2][1], p0.z, c.y);
        ffma2_bcast(acc[3][0], p0.w, c.x); ffma2_bcast(acc[3][1], p0.w, c.y);
        ffma2_bcast(acc[4][0], p1.x, c.x); ffma2_bcast(acc[4][1], p1.x, c.y);
        ffma2_bcast(acc[5][0], p1.y, c.x); ffma2_bcast(acc[5][1], p1.y, c.y);
        ffma2_bcast(acc[6][0], p1.z, c.x); ffma2_bcast(acc[6][1], p1.z, c.y);
        ffma2_bcast(acc[7][0], p1.w, c.x); ffma2_bcast(acc[7][1], p1.w, c.y);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pl = plp[j];                        // plp now points at the last row
        flush_pair_if(out + static_cast<size_t>(static_cast<unsigned>(pl)) * (64 * 4), acc[j][0], acc[j][1], pl >= 0 ? 1u : 0u);
    }
}

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

// Same pass with 16-byte stores: a thread takes four consecutive pillars x one quarter of the channels (16), loads its
// sixteen 16-byte pieces up front, transposes the 4 x 16 block in registers and writes 16 stores of 16 bytes; a warp covers
// 128 pillars, i.e. 512 contiguous bytes per channel row.  Half the LSU instructions of the per-pillar kernel (which is
// limited by its 64 four-byte stores per thread: 59 % "lg throttle" stalls) at the same thread count.  The touched byte
// of a pillar carries one bit per channel quarter so the four quarter-blocks can clear their bit independently.
// Needs X*Y to be a multiple of 4.
__global__ void __launch_bounds__(FIN_THREADS)
finalize_nchw_q_kernel(float* __restrict__ accum, unsigned* __restrict__ flags32, float* __restrict__ bev,
                       long long pillars, int blocks_per_frame) {
    constexpr int C = 64;
    const int q = blockIdx.x & 3;                      // channel quarter
    const int b = blockIdx.x >> 2;
    const int frame = b / blocks_per_frame;
    const long long p0 = (static_cast<long long>(b % blocks_per_frame) * FIN_THREADS + threadIdx.x) * 4;
    if (p0 >= pillars) return;
    unsigned* fw = flags32 + (static_cast<size_t>(frame) * pillars + p0) / 4;
    const unsigned word = __ldcg(fw) >> q;             // bit 0 of each byte: this quarter still holds data for that pillar
    const bool t0 = word & 0x1u, t1 = word & 0x100u, t2 = word & 0x10000u, t3 = word & 0x1000000u;
    float4* row = reinterpret_cast<float4*>(accum + (static_cast<size_t>(frame) * pillars + p0) * C) + q * 4;   // rows are 16 float4 apart
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[0][k] = t0 ? __ldcg(row + 0 * 16 + k) : z4;
        v[1][k] = t1 ? __ldcg(row + 1 * 16 + k) : z4;
        v[2][k] = t2 ? __ldcg(row + 2 * 16 + k) : z4;
        v[3][k] = t3 ? __ldcg(row + 3 * 16 + k) : z4;
    }
    float* dst = bev + (static_cast<size_t>(frame) * C + q * 16) * pillars + p0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float* d = dst + static_cast<size_t>(4 * k) * pillars;
        // streaming stores: the BEV is written once and not re-read here; keep L2 for the accumulator rows
        __stcs(reinterpret_cast<float4*>(d), make_float4(v[0][k].x, v[1][k].x, v[2][k].x, v[3][k].x));
        __stcs(reinterpret_cast<float4*>(d + pillars), make_float4(v[0][k].y, v[1][k].y, v[2][k].y, v[3][k].y));
        __stcs(reinterpret_cast<float4*>(d + 2 * pillars), make_float4(v[0][k].z, v[1][k].z, v[2][k].z, v[3][k].z));
        __stcs(reinterpret_cast<float4*>(d + 3 * pillars), make_float4(v[0][k].w, v[1][k].w, v[2][k].w, v[3][k].w));
    }
}

// Re-zero the accumulator rows and the marks of the touched pillars (scratch invariant of include/fiery_b200.h).  Kept out
// of the layout pass: interleaving these scattered stores with the output stream doubles that kernel's time
// (tools/microbench: 64.7 us with, 31.8 us without, 9 frames), while on their own they are cheap.
__global__ void __launch_bounds__(FIN_THREADS)
clear_touched_kernel(float* __restrict__ accum, unsigned char* __restrict__ flags, long long total_pillars) {
    const long long i = static_cast<long long>(blockIdx.x) * FIN_THREADS + threadIdx.x;     // (pillar, 16-byte piece)
    const long long pl = i >> 4;
    if (pl >= total_pillars) return;
    if (__ldcg(flags + pl)) reinterpret_cast<float4*>(accum)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the mark is cleared by a second launch of this kernel's sibling below, after every piece has been zeroed
}

__global__ void __launch_bounds__(FIN_THREADS)
clear_marks_kernel(unsigned* __restrict__ flags32, long long n_words) {
    const long long i = static_cast<long long>(blockIdx.x) * FIN_THREADS + threadIdx.x;
    if (i < n_words && __ldcg(flags32 + i)) flags32[i] = 0u;
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
int encode_head_maps(HeadMaps* maps, const void* head, int dtype, const LiftParams& P);

template <int DBLKS>
static int launch_forward_t(const HeadMaps& map, const LiftParams& P, cudaStream_t stream) {
    const TileLayout<DBLKS> L(P.hh, P.C);
    const int n_pblk = (L.PX + 31) / 32;
    FIERY_REQUIRE(P.hh <= 32, "feat_h=%d not supported by this build (<= 32)", P.hh);
    FIERY_REQUIRE(n_pblk * (1 + P.C / 32) <= TileLayout<DBLKS>::NWARPS,
                  "feature map too tall for this build: h=%d needs %d staging warps, kernel has %d", P.hh,
                  n_pblk * (1 + P.C / 32), TileLayout<DBLKS>::NWARPS);
    static bool configured = false;
    if (!configured) {
        FIERY_CUDA_CHECK(cudaFuncSetAttribute(lift_forward_kernel<DBLKS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              227 * 1024));
        // three tiles per SM (3 x 74 KB for the reference shape): ask for the full shared-memory carve-out
        FIERY_CUDA_CHECK(cudaFuncSetAttribute(lift_forward_kernel<DBLKS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    FIERY_REQUIRE(L.total <= 227 * 1024, "tile needs %d bytes of shared memory", L.total);
    const long long n_tiles = static_cast<long long>(P.n_frames) * P.n_cameras * P.n_wtiles;
    lift_forward_kernel<DBLKS><<<static_cast<unsigned>(n_tiles), 64 * DBLKS, L.total, stream>>>(map, P);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

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

int launch_forward_cols(const LiftParams& P, const void* head, int variant, cudaStream_t stream);

// FIERY_LIFT_FORWARD=rows selects the row-major tile kernel of this file (kept as the A/B partner of lift_fwd_cols.cu)
static int use_cols_kernel() {        // 0: rows kernel, else the column kernel variant (see launch_forward_cols)
    static const int cols = [] {
        const char* e = getenv("FIERY_LIFT_FORWARD");
        if (e && strcmp(e, "rows") == 0) return 0;
        if (e && strcmp(e, "cols3") == 0) return 43;
        if (e && strcmp(e, "cols2") == 0) return 42;
        return 22;
    }();
    return cols;
}

int launch_lift_forward(const LiftParams& P, const void* head, int head_dtype, float* bev_out, float* scratch,
                        cudaStream_t stream) {
    FIERY_REQUIRE(head_dtype == FIERY_DTYPE_F32, "head dtype %d not supported by this build (fp32 only)", head_dtype);
    FIERY_REQUIRE(P.C == 64, "channels=%d not supported by this build (C must be 64)", P.C);
    FIERY_REQUIRE(P.D >= 1 && P.D <= 48, "depth_bins=%d not supported by this build (1..48)", P.D);
    FIERY_REQUIRE(P.ww % 4 == 0, "feat_w=%d must be a multiple of 4 (TMA row pitch must be 16-byte aligned)", P.ww);
    const int cols = use_cols_kernel();
    HeadMaps map;
    int rc = cols ? FIERY_OK : encode_head_maps(&map, head, head_dtype, P);
    if (rc != FIERY_OK) return rc;
    LiftParams Q = P;
    if (P.bev_layout == FIERY_BEV_NHWC) {          // the caller's zero-filled channel-last tensor is the accumulator
        Q.accum = bev_out;
        Q.touched = nullptr;
        Q.frame0 = 0;
        return cols ? launch_forward_cols(Q, head, cols, stream) : launch_forward_t<6>(map, Q, stream);
    }
    // NCHW: lift into the channel-last accumulator, then the layout pass; chunked only to bound the scratch footprint
    const int chunk = lift_chunk_frames(P.n_frames, P.pillars, P.C);
    Q.accum = scratch;                             // [accumulator floats of one chunk][one "touched" byte per pillar]
    Q.touched = reinterpret_cast<unsigned char*>(scratch + static_cast<size_t>(chunk) * P.pillars * P.C);
    const int bpf = static_cast<int>((P.pillars + FIN_THREADS - 1) / FIN_THREADS);
    for (int f0 = 0; f0 < P.n_frames; f0 += chunk) {
        Q.frame0 = f0;
        Q.n_frames = (P.n_frames - f0 < chunk) ? P.n_frames - f0 : chunk;
        rc = cols ? launch_forward_cols(Q, head, cols, stream) : launch_forward_t<6>(map, Q, stream);
        if (rc != FIERY_OK) return rc;
        if (P.pillars % 4 == 0) {
            const int bpf4 = static_cast<int>((P.pillars / 4 + FIN_THREADS - 1) / FIN_THREADS);
            finalize_nchw_q_kernel<<<static_cast<unsigned>(bpf4) * Q.n_frames * 4, FIN_THREADS, 0, stream>>>(
                Q.accum, reinterpret_cast<unsigned*>(Q.touched), bev_out + static_cast<size_t>(f0) * P.C * P.pillars, P.pillars, bpf4);
            const long long tp = static_cast<long long>(Q.n_frames) * P.pillars;
            clear_touched_kernel<<<static_cast<unsigned>((tp * 16 + FIN_THREADS - 1) / FIN_THREADS), FIN_THREADS, 0, stream>>>(Q.accum, Q.touched, tp);
            const long long nw = (tp + 3) / 4;
            clear_marks_kernel<<<static_cast<unsigned>((nw + FIN_THREADS - 1) / FIN_THREADS), FIN_THREADS, 0, stream>>>(
                reinterpret_cast<unsigned*>(Q.touched), nw);
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
