// Forward camera->BEV lift for sm_100a, host side and the passes around the tile kernel (lift_fwd_cols.cu): the NCHW layout
// pass, the integer index dump and the calibration composition used by the parity checks, and the launchers.
//
// Replaces, per call: Fiery.get_geometry (fiery/models/fiery.py:193-208), the tail of Encoder.forward
// (fiery/models/encoder.py:98-102), and Fiery.projection_to_birds_eye_view incl. VoxelsSumming
// (fiery/models/fiery.py:221-273, fiery/utils/geometry.py:283-314).
#include <atomic>
#include <mutex>

#include "lift_plan.cuh"
#include "warp_sample.cuh"

namespace fiery {

// ---------------------------------------------------------------------------------------------------------------------
// Fallback layout pass for NCHW output when X*Y is not a multiple of 4 (the TMA pass below needs a 16-byte row pitch):
// accum (B', X*Y, C) -> bev (B', C, X*Y).  One thread per pillar: a lane reads its pillar's 256-byte accumulator row as 16
// independent 16-byte loads (its own two cache lines, so the sectors are fully used through L1), and the warp then writes one
// channel of 32 consecutive pillars per store instruction -- a full 128-byte line.  Only ~1/3-1/2 of the pillars receive any
// point, and the tile kernel marks those in a byte map: unmarked pillars are written as zeros without touching the
// accumulator; marked rows (and, when the marks live in the scratch, the marks) are re-zeroed on the way (scratch invariant of
// include/fiery_b200.h).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
__global__ void __launch_bounds__(FIN_THREADS)
finalize_nchw_kernel(float* __restrict__ accum, unsigned char* __restrict__ flags, float* __restrict__ bev,
                     long long pillars, int blocks_per_frame, int clear_marks) {
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
        if (clear_marks) *f = 0;
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

// ---------------------------------------------------------------------------------------------------------------------
// Layout pass built on the copy engine.  A CTA takes FT_P consecutive pillars of one frame:
//   1. one thread per pillar reads the pillar's touched byte and, if set, fetches the pillar's 256-byte accumulator row with
//      a 1-D bulk copy (cp.async.bulk, completion on an mbarrier) -- only rows that received points are read, each as one
//      contiguous 256-byte burst;
//   2. the (pillar, channel) block is transposed shared -> shared: 16-byte reads in an XOR-rotated chunk order (every
//      quarter-warp phase covers all 32 banks), 4-byte writes with bank = lane; rows that were not fetched read as zero;
//   3. one tiled TMA store writes the (64 channels x FT_P pillars) block into the NCHW output (256 contiguous bytes per
//      channel row), and a 256-byte bulk copy of zeros per fetched row plus a byte store per mark restore the scratch
//      invariant of include/fiery_b200.h.
// Needs X*Y to be a multiple of 4 (16-byte row pitch of the output map).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FT_P = 64;
constexpr int FT_THREADS = 256;

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
                    long long pillars, int tiles_per_frame, int frame_out0, int clear_marks) {
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
        // 16-byte reads: a quarter-warp (8 lanes = 8 pillars, rows 256 B apart) reads the 8 different 16-byte chunks
        // chunk0 + (i ^ (lane & 7)) of a 32-channel block, so every LDS.128 phase covers all 32 banks; the four values go to
        // channel rows 4*chunk .. 4*chunk+3 of this lane's pillar column (bank = lane)
        const int lane = tid & 31, w = tid >> 5;
        const int pl = (w & 1) * 32 + lane;             // pillar of this lane
        const int chunk0 = ((w >> 1) & 1) * 8;          // 32-channel block = 8 chunks of 4 channels
        const int i0 = (w >> 2) * 4;                    // this warp's four of the eight rotations
        const bool have = s_flag[pl] != 0;
        const float4* src = reinterpret_cast<const float4*>(s_in + pl * C);
        float* dst = s_out + pl;
        const int l7 = lane & 7;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int chunk = chunk0 + ((i0 + ii) ^ l7);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have) v = src[chunk];
            float* d = dst + chunk * (4 * FT_P);
            d[0] = v.x; d[FT_P] = v.y; d[2 * FT_P] = v.z; d[3 * FT_P] = v.w;
        }
    }
    fence_proxy_async();                                // generic-proxy writes (s_out, s_zero) -> visible to the copy engine
    __syncthreads();
    if (tid == 0) tma_store_3d(&bev_map, s_out, static_cast<int>(p0), 0, frame_out0 + frame);
    if (row) {                                          // restore the all-zero scratch
        bulk_store_1d(row, s_zero, C * 4);
        if (clear_marks) *mark = 0;             // marks of a caller-owned plan stay: the plan is reused
    }
    if (tid == 0 || row) tma_store_commit_and_wait();   // the shared sources must outlive the copies
}

// ---------------------------------------------------------------------------------------------------------------------
// Layout pass with the warp of cumulative_warp_features folded in (fiery/models/fiery.py:143-146,
// fiery/utils/geometry.py:181-253): instead of transposing the channel-last accumulator of a frame into the NCHW output and letting
// a second kernel re-read it, every OUTPUT pixel of the frame gathers its (up to) four bilinear neighbours straight from the
// accumulator -- whose 256-byte channel rows are exactly what a gather wants -- and the blended pixel goes to the NCHW output.
// Present frames (copy flag) take their own pillar with weight one: bit-identical to the plain layout pass.  Sample positions
// and the blend order are those of warp_forward_kernel (warp_sample.cuh).  A quarter-warp serves one output pixel (lane = two
// 16-byte pieces of the channel row: one full 128-byte line per quarter and load instruction), four neighbours are fetched
// before the first use; pillars that received no point (mark byte clear) are not read.  The (64 pixels x 64 channels) block is
// turned through padded shared memory (conflict-free both ways) so that every store instruction writes 128 contiguous bytes of a
// channel plane.  A source pillar is read by several CTAs, so the accumulator cannot be cleared here: clear_touched_kernel
// restores the all-zero scratch afterwards.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FW_P = 64;
constexpr int FW_THREADS = 256;
__global__ void __launch_bounds__(FW_THREADS)
finalize_warp_kernel(const float* __restrict__ accum, const unsigned char* __restrict__ touched, float* __restrict__ bev,
                     long long pillars, int H, int W, int tiles_per_frame, int frame_out0, const float* __restrict__ theta,
                     const unsigned char* __restrict__ copy_mask) {
    constexpr int C = 64;
    __shared__ float s_out[C][FW_P + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.x / tiles_per_frame;
    const long long p0 = static_cast<long long>(blockIdx.x % tiles_per_frame) * FW_P;
    const int gframe = frame_out0 + frame;
    const float* acc_f = accum + static_cast<size_t>(frame) * pillars * C;
    const unsigned char* tch = touched + static_cast<size_t>(frame) * pillars;
    const int l = lane & 7, q = lane >> 3;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int pxl = pass * 32 + warp * 4 + q;
        const long long pix = p0 + pxl;
        float4 a[4], b[4];
        float w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            b[k] = a[k];
            w[k] = 0.f;
        }
        if (pix < pillars) {
            const SamplePos s = make_sample(theta, copy_mask, gframe, static_cast<int>(pix), W, H, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                w[k] = s.w[k];
                if (s.ok[k] && tch[s.off[k]]) {
                    const float4* row = reinterpret_cast<const float4*>(acc_f + static_cast<size_t>(s.off[k]) * C);
                    a[k] = __ldg(row + l);
                    b[k] = __ldg(row + 8 + l);
                }
            }
        }
        // the blend of warp_forward_kernel: w0*v0, then fma with neighbours 1..3
        float4 ra, rb;
#define FIERY_BLEND(f) \
        ra.f = fmaf(w[3], a[3].f, fmaf(w[2], a[2].f, fmaf(w[1], a[1].f, w[0] * a[0].f))); \
        rb.f = fmaf(w[3], b[3].f, fmaf(w[2], b[2].f, fmaf(w[1], b[1].f, w[0] * b[0].f)));
        FIERY_BLEND(x) FIERY_BLEND(y) FIERY_BLEND(z) FIERY_BLEND(w)
#undef FIERY_BLEND
        s_out[4 * l + 0][pxl] = ra.x; s_out[4 * l + 1][pxl] = ra.y; s_out[4 * l + 2][pxl] = ra.z; s_out[4 * l + 3][pxl] = ra.w;
        s_out[32 + 4 * l + 0][pxl] = rb.x; s_out[32 + 4 * l + 1][pxl] = rb.y; s_out[32 + 4 * l + 2][pxl] = rb.z; s_out[32 + 4 * l + 3][pxl] = rb.w;
    }
    __syncthreads();
    float* dst = bev + static_cast<size_t>(gframe) * C * pillars + p0;
#pragma unroll
    for (int cc = 0; cc < C / (FW_THREADS / 32); ++cc) {
        const int c = warp * (C / (FW_THREADS / 32)) + cc;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int px = half * 32 + lane;
            if (p0 + px < pillars) __stcs(dst + static_cast<size_t>(c) * pillars + px, s_out[c][px]);
        }
    }
}

// Restores the all-zero scratch after finalize_warp_kernel: a warp looks at 32 marks and zeroes the 256-byte accumulator row of
// every marked pillar with one 8-byte store per lane; marks that live in the scratch are cleared too (a caller's plan keeps its own).
__global__ void __launch_bounds__(256)
clear_touched_kernel(float* __restrict__ accum, unsigned char* __restrict__ touched, long long n_pillars, int clear_marks) {
    const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned char f = p < n_pillars ? touched[p] : 0;
    unsigned m = __ballot_sync(0xffffffffu, f != 0);
    const long long base = p - lane;
    while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        reinterpret_cast<float2*>(accum + static_cast<size_t>(base + i) * 64)[lane] = make_float2(0.f, 0.f);
    }
    if (f && clear_marks) touched[p] = 0;
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
// Frames per launch.  Measured on B200 (profiles/r01_notes.md): chunks small enough to keep the accumulator L2-resident (3 frames,
// 31 MB) are slower end to end (152.9 us vs 122.7 us for 9 frames) -- the extra launches and the single-wave grids cost more than the
// saved HBM traffic -- so the chunk only bounds the scratch footprint (1 GiB: accumulator + marks of a chunk).
static std::atomic<int> g_max_chunk_frames{0};       // fiery_lift_set_max_chunk_frames (test hook: forces the multi-pass path)
void lift_set_max_chunk_frames(int n) { g_max_chunk_frames.store(n > 0 ? n : 0); }

int lift_chunk_frames(const LiftParams& P) {
    const long long per_frame = P.pillars * P.C * 4 + P.pillars;
    long long c = (1ll << 30) / (per_frame > 0 ? per_frame : 1);
    const int forced = g_max_chunk_frames.load();
    if (forced > 0 && forced < c) c = forced;
    if (c > P.n_frames) c = P.n_frames;
    return static_cast<int>(c < 1 ? 1 : c);
}

// scratch of one pass (NCHW output): [accumulator (chunk, X*Y, C) fp32][marks (chunk, X*Y) bytes, padded to 128]; all zero between calls
size_t lift_scratch_bytes(const LiftParams& P) {
    if (P.bev_layout != FIERY_BEV_NCHW || P.n_frames <= 0) return 0;
    const size_t chunk = static_cast<size_t>(lift_chunk_frames(P));
    return chunk * P.pillars * P.C * 4 + ((chunk * P.pillars + 127) & ~static_cast<size_t>(127));
}

int launch_forward_cols(const LiftParams& P, const void* head, cudaStream_t stream);
int encode_bev_map(CUtensorMap* map, float* bev, long long pillars, int channels, int n_frames, int box_pillars);

// Side streams and fork/join events of the forward chains: created once per host thread and device (thread_local, so concurrent
// callers never share or race on them), reused by every call -- nothing is created or destroyed on the launch path.
constexpr int MAX_CHAINS = 4;
struct ChainResources {
    cudaStream_t side[MAX_CHAINS - 1] = {};
    cudaEvent_t fork = nullptr;
    cudaEvent_t done[MAX_CHAINS - 1] = {};
    bool ready = false;
};
static int chain_resources(ChainResources** out) {
    static thread_local ChainResources pool[16];
    int dev = 0;
    FIERY_CUDA_CHECK(cudaGetDevice(&dev));
    FIERY_REQUIRE(dev >= 0 && dev < 16, "device %d out of range for the chain resources", dev);
    ChainResources& r = pool[dev];
    if (!r.ready) {
        for (int i = 0; i < MAX_CHAINS - 1; ++i) {
            FIERY_CUDA_CHECK(cudaStreamCreateWithFlags(&r.side[i], cudaStreamNonBlocking));
            FIERY_CUDA_CHECK(cudaEventCreateWithFlags(&r.done[i], cudaEventDisableTiming));
        }
        FIERY_CUDA_CHECK(cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming));
        r.ready = true;
    }
    *out = &r;
    return FIERY_OK;
}

// NCHW output: the frames of a chunk are cut into groups, each a (plan kernel ->) tile kernel -> layout pass chain on its own stream.
// The layout pass of one group (DRAM-bound) runs under the tile kernels of the others (issue-bound); only the last pass is
// exposed.  Measured on B200 (profiles/r01_notes.md), 8 frames: 1 chain 84.7 us, 2 chains 78.3 us, 4 chains 73.3 us; 8 chains of
// one frame (90 tiles) each fall back to 79.6 us, so a group keeps at least one tile per SM (148).
#ifdef FIERY_COLS_AB
static int g_max_chains = MAX_CHAINS;      // FIERY_CHAINS (A/B builds)
static int g_chain_min_tiles = 148;        // FIERY_CHAIN_MIN_TILES (A/B builds)
#else
constexpr int g_max_chains = MAX_CHAINS;
constexpr int g_chain_min_tiles = 148;
#endif
int lift_forward_groups(const LiftParams& P, int frames_in_chunk) {
    if (P.bev_layout == FIERY_BEV_NHWC) return 1;          // no layout pass to hide
    int groups = frames_in_chunk < g_max_chains ? frames_in_chunk : g_max_chains;
    const long long tiles_per_frame = static_cast<long long>(P.n_cameras) * P.n_wtiles;
    while (groups > 1 && (frames_in_chunk / groups) * tiles_per_frame < g_chain_min_tiles) --groups;
    return groups < 1 ? 1 : groups;
}

// kernel launches of one forward call (include/fiery_b200.h: fiery_lift_forward_launches)
int lift_forward_launches(const LiftParams& P) {
    if (P.n_frames <= 0) return 0;
    const int per_group = 1 + (P.bev_layout == FIERY_BEV_NCHW ? 1 : 0);
    const int chunk = lift_chunk_frames(P);
    int n = 0;
    for (int f0 = 0; f0 < P.n_frames; f0 += chunk)
        n += per_group * lift_forward_groups(P, (P.n_frames - f0 < chunk) ? P.n_frames - f0 : chunk);
    return n;
}

// Per-launch timing (bench / profiling): when set, every kernel launch of the next forward call on this host thread is bracketed by
// a pair of events on its own stream; see fiery_lift_forward_timed in c_api.cu.
static thread_local LaunchTimer* g_timer = nullptr;
void lift_set_timer(LaunchTimer* t) { g_timer = t; }
static inline void timer_begin(cudaStream_t st, int kind) {
    if (g_timer && g_timer->n < g_timer->cap) {
        g_timer->kind[g_timer->n] = kind;
        cudaEventRecord(g_timer->ev[2 * g_timer->n], st);
    }
}
static inline void timer_end(cudaStream_t st) {
    if (g_timer && g_timer->n < g_timer->cap) {
        cudaEventRecord(g_timer->ev[2 * g_timer->n + 1], st);
        ++g_timer->n;
    }
}

// warp_theta != NULL: the layout pass samples every frame under its (2, 3) affine map (frames flagged in warp_copy pass through) --
// the lift followed by cumulative_warp_features in one chain; NCHW output only
int launch_lift_forward(const LiftParams& P, const void* head, int head_dtype, float* bev_out, void* scratch, const void* plan,
                        const float* warp_theta, const unsigned char* warp_copy, cudaStream_t stream) {
    FIERY_REQUIRE(head_dtype == FIERY_DTYPE_F32 || head_dtype == FIERY_DTYPE_F16, "head dtype %d not supported (fp32 / fp16)", head_dtype);
    FIERY_REQUIRE(P.C == 64, "channels=%d not supported by this build (C must be 64)", P.C);
    FIERY_REQUIRE(P.D >= 1 && P.D <= 48, "depth_bins=%d not supported by this build (1..48)", P.D);
    FIERY_REQUIRE(P.ww % 4 == 0, "feat_w=%d must be a multiple of 4 (TMA row pitch must be 16-byte aligned)", P.ww);
    FIERY_REQUIRE(P.hh <= PLAN_MAX_ROWS, "feat_h=%d not supported by this build (<= %d)", P.hh, PLAN_MAX_ROWS);
    const bool nchw = P.bev_layout == FIERY_BEV_NCHW;
    FIERY_REQUIRE(!warp_theta || nchw, "the warped lift writes the NCHW layout only");
    FIERY_REQUIRE(scratch != nullptr || !nchw, "NCHW output needs the zeroed scratch buffer of fiery_lift_scratch_bytes()");
    int rc = FIERY_OK;
    LiftParams Q = P;
    Q.head_f16 = head_dtype == FIERY_DTYPE_F16 ? head : nullptr;
#ifdef FIERY_COLS_AB
    if (const char* e = getenv("FIERY_CHAINS")) g_max_chains = atoi(e) < MAX_CHAINS ? atoi(e) : MAX_CHAINS;
    if (const char* e = getenv("FIERY_CHAIN_MIN_TILES")) g_chain_min_tiles = atoi(e);
#endif
    // lift into a channel-last accumulator (NHWC: the caller's zero-filled output itself), then the layout pass for NCHW; several
    // passes only to bound the scratch footprint
    const int chunk = lift_chunk_frames(P);
    float* accum = static_cast<float*>(scratch);    // [accumulator floats of one pass][one mark byte per pillar]
    unsigned char* scratch_marks = nchw ? reinterpret_cast<unsigned char*>(accum + static_cast<size_t>(chunk) * P.pillars * P.C) : nullptr;
    const bool tma_pass = P.pillars % 4 == 0;      // the output map needs a 16-byte row pitch
    CUtensorMap bev_map;
    if (nchw && tma_pass && !warp_theta) {
        rc = encode_bev_map(&bev_map, bev_out, P.pillars, P.C, P.n_frames, FT_P);
        if (rc != FIERY_OK) return rc;
    }
    ChainResources* res = nullptr;
    for (int f0 = 0; f0 < P.n_frames; f0 += chunk) {
        const int nf = (P.n_frames - f0 < chunk) ? P.n_frames - f0 : chunk;
        const int groups = lift_forward_groups(P, nf);
        cudaStream_t chain[MAX_CHAINS] = {stream};
        if (groups > 1) {                           // fork: the side streams start behind everything queued on the caller's
            if (!res) {
                rc = chain_resources(&res);
                if (rc != FIERY_OK) return rc;
            }
            for (int g = 1; g < groups; ++g) chain[g] = res->side[g - 1];
            FIERY_CUDA_CHECK(cudaEventRecord(res->fork, stream));
        }
        for (int g = 0; g < groups; ++g) {
            const int s0 = static_cast<int>(static_cast<long long>(nf) * g / groups);
            const int s1 = static_cast<int>(static_cast<long long>(nf) * (g + 1) / groups);
            cudaStream_t st = chain[g];
            if (g > 0) FIERY_CUDA_CHECK(cudaStreamWaitEvent(st, res->fork, 0));
            Q.frame0 = f0 + s0;
            Q.n_frames = s1 - s0;
            unsigned char* marks = nullptr;         // "pillar receives a point" bytes the layout pass reads
            if (plan) {                             // caller-owned plan of the whole batch: read only, marks included
                const PlanView v = plan_view(plan, P.n_frames, P.n_cameras, P.n_wtiles, P.pillars, Q.frame0);
                Q.plan_tiles = v.tiles;
                Q.touched = nullptr;
                marks = const_cast<unsigned char*>(v.touched);
            } else {                                // the tile kernel evaluates the geometry itself and marks into the scratch
                Q.plan_tiles = nullptr;
                marks = nchw ? scratch_marks + static_cast<size_t>(s0) * P.pillars : nullptr;
                Q.touched = marks;
            }
            Q.accum = nchw ? accum + static_cast<size_t>(s0) * P.pillars * P.C
                           : bev_out + static_cast<size_t>(Q.frame0) * P.pillars * P.C;
            timer_begin(st, 1);
            rc = launch_forward_cols(Q, head, st);
            timer_end(st);
            if (rc != FIERY_OK) return rc;
            if (nchw) {
                timer_begin(st, 2);
                if (warp_theta) {
                    const int tpf = static_cast<int>((P.pillars + FW_P - 1) / FW_P);
                    finalize_warp_kernel<<<static_cast<unsigned>(tpf) * Q.n_frames, FW_THREADS, 0, st>>>(
                        Q.accum, marks, bev_out, P.pillars, P.grid.X, P.grid.Y, tpf, Q.frame0, warp_theta, warp_copy);
                    const long long n = P.pillars * Q.n_frames;
                    clear_touched_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(Q.accum, marks, n, plan ? 0 : 1);
                } else if (tma_pass) {
                    const int tpf = static_cast<int>((P.pillars + FT_P - 1) / FT_P);
                    finalize_tma_kernel<<<static_cast<unsigned>(tpf) * Q.n_frames, FT_THREADS, 0, st>>>(bev_map, Q.accum, marks, P.pillars,
                                                                                                     tpf, Q.frame0, plan ? 0 : 1);
                } else {
                    const int bpf = static_cast<int>((P.pillars + FIN_THREADS - 1) / FIN_THREADS);
                    finalize_nchw_kernel<<<static_cast<unsigned>(bpf) * Q.n_frames, FIN_THREADS, 0, st>>>(
                        Q.accum, marks, bev_out + static_cast<size_t>(Q.frame0) * P.C * P.pillars, P.pillars, bpf, plan ? 0 : 1);
                }
                timer_end(st);
                FIERY_CUDA_CHECK(cudaGetLastError());
            }
            if (g > 0) {                            // join the chain back into the caller's stream
                FIERY_CUDA_CHECK(cudaEventRecord(res->done[g - 1], st));
                FIERY_CUDA_CHECK(cudaStreamWaitEvent(stream, res->done[g - 1], 0));
            }
        }
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
