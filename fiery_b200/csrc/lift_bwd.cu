// Backward of the camera->BEV lift for sm_100a: gradient of the BEV features w.r.t. the head tensor
// (depth logits + context), i.e. autograd through fiery/models/encoder.py:99-100 (softmax, outer product) and
// fiery/utils/geometry.py:305-314 (VoxelsSumming.backward = "send the voxel's gradient to every point summed into it")
// without ever materialising the (N, C) point gradient the reference builds.
//
// Per point (pixel p = (camera, row, column), depth d) with pillar pi(p, d) and G = grad_bev[:, pi]:
//     g_ctx[p][c]  = sum_d prob[p][d] * G[pi(p,d)][c]
//     g_prob[p][d] = sum_c ctx[p][c]  * G[pi(p,d)][c]
//     g_logit[p][d] = prob[p][d] * (g_prob[p][d] - sum_d' prob[p][d'] g_prob[p][d'])          (softmax backward)
// The tile staging (TMA, softmax, transposes, pillar ranks, change bits) lives in lift_tile.cuh.  A thread
// owns one column, a group of <= MAXR consecutive rows and 4 channels; it loops over the depth blocks keeping the
// 8 x 4 gradient values G of the current pillars in registers (reloaded from the channel-last grad_bev only where a
// change bit says the pillar changed), so g_ctx needs no cross-thread reduction; g_prob is reduced over the 16 channel
// lanes with a transposing shuffle butterfly.  The result is transposed back to NCHW in shared memory and written with
// TMA stores.
#include "lift_tile.cuh"

namespace fiery {

constexpr int MAXR = 5;   // rows per thread; host checks ceil(h / (threads/64)) <= MAXR (h <= 30)

// packed fp32x2 helpers (SASS FFMA2 / FMUL2)
__device__ __forceinline__ void fma2_bcast(unsigned long long& acc, float a, unsigned long long b) {
    unsigned long long aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(aa), "l"(b));
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float pair_sum(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
// In-place predicated updates of a register pair (straight-line for the compiler: a conditional C++ assignment inside the
// row loop makes ptxas shuffle the whole register set every iteration, see profiles/r01_notes.md).
__device__ __forceinline__ void clear_pair_if(unsigned long long& a, unsigned long long& b, unsigned pred) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p mov.b64 %0, 0;\n\t@p mov.b64 %1, 0;\n\t}" : "+l"(a), "+l"(b) : "r"(pred));
}
__device__ __forceinline__ void load_pair_if(unsigned long long& a, unsigned long long& b, const float* ptr, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.v2.b64 {%0, %1}, [%2];\n\t}"
                 : "+l"(a), "+l"(b) : "l"(ptr), "r"(static_cast<unsigned>(pred)) : "memory");
}

template <int DBLKS>
__global__ void __launch_bounds__(64 * DBLKS, 2)
lift_backward_kernel(const __grid_constant__ HeadMaps head_maps, const __grid_constant__ HeadMaps grad_maps,
                     const LiftParams P) {
    using TL = TileLayout<DBLKS>;
    constexpr int DPAD = TL::DPAD;
    constexpr int PS = TL::PS;
    constexpr int HQ = DBLKS;                        // row groups per column: threads = WT * HQ * 16
    extern __shared__ __align__(128) unsigned char smem[];
    const TL L(P.hh, P.C);
    float* s_gprob = reinterpret_cast<float*>(smem + L.total);     // [pix][DPAD], appended to the forward layout

    const int wtile = blockIdx.x % P.n_wtiles;
    const int img = blockIdx.x / P.n_wtiles;
    const int frame = img / P.n_cameras;
    const int w0 = wtile * WT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hh = L.hh, PX = L.PX;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    if (tid == 0) {
        tma_prefetch_desc(&head_maps.depth);
        tma_prefetch_desc(&head_maps.ctx);
        mbar_init(bar, 1);
        fence_mbar_init();
        issue_tile_loads<DBLKS>(P, L, smem, &head_maps, img, w0);
    }
    stage_constants<DBLKS>(P, L, smem, img, w0);
    __syncthreads();
    if (tid == 64 * DBLKS - 1) stage_camera<DBLKS>(P, L, smem, img);     // overlaps the TMA latency
    mbar_wait(bar, 0);
    transform_tile<DBLKS>(P, L, smem);
    stage_pillars<DBLKS>(P, L, smem, w0);
    __syncthreads();
    stage_change_bits<DBLKS>(L, smem);
    __syncthreads();

    // ---- main loop ----------------------------------------------------------------------------------------------------
    float* s_prob = reinterpret_cast<float*>(smem + L.off_prob);
    float* s_ctx = reinterpret_cast<float*>(smem + L.off_ctx);
    const int* s_pillar = reinterpret_cast<const int*>(smem + L.off_pillar);
    const unsigned char* s_chg = smem + L.off_chg;

    const int unit = warp * 2 + (lane >> 4);
    const int wt = unit / HQ, hq = unit % HQ;
    const int cg = lane & 15;
    const unsigned half_mask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
    const int h_lo = (hh * hq) / HQ, h_hi = (hh * (hq + 1)) / HQ;
    const float* gbev = P.grad_bev + static_cast<size_t>(frame) * P.pillars * P.C + cg * 4;   // channel-last

    // accumulators and gradient vectors as fp32x2 pairs: (c0,c1) and (c2,c3) of this lane's 4 channels
    unsigned long long gc[MAXR][2];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) gc[r][0] = gc[r][1] = 0ull;

    const bool b3 = cg & 8, b2 = cg & 4, b1 = cg & 2;
    const int jsel = (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);   // the depth (within the block) this lane ends up owning

    for (int dblk = 0; dblk < DBLKS; ++dblk) {
        unsigned long long G[8][2];
        {
            const int* pl = s_pillar + (wt * hh + h_lo) * DPAD + dblk * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = (h_lo < h_hi) ? pl[j] : -1;
                G[j][0] = G[j][1] = 0ull;
                load_pair_if(G[j][0], G[j][1], gbev + static_cast<size_t>(static_cast<unsigned>(p < 0 ? 0 : p)) * P.C, p >= 0);
            }
        }
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            const int h = h_lo + r;
            if (h < h_hi) {
                const int pix = wt * hh + h;
                if (r > 0) {
                    const unsigned m = s_chg[pix * DBLKS + dblk];
                    if (m) {     // some depth of this block enters another pillar at this row: fetch its gradient vector
                        const int* pl_row = s_pillar + pix * DPAD + dblk * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const unsigned bit = m & (1u << j);
                            const int p = bit ? pl_row[j] : -1;
                            clear_pair_if(G[j][0], G[j][1], bit);
                            load_pair_if(G[j][0], G[j][1], gbev + static_cast<size_t>(static_cast<unsigned>(p < 0 ? 0 : p)) * P.C, p >= 0);
                        }
                    }
                }
                const float4 p0 = *reinterpret_cast<const float4*>(s_prob + pix * PS + dblk * 8);
                const float4 p1 = *reinterpret_cast<const float4*>(s_prob + pix * PS + dblk * 8 + 4);
                const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(s_ctx + pix * L.C + cg * 4);
                const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                float gp[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    fma2_bcast(gc[r][0], pv[j], G[j][0]);                       // g_ctx += prob * G
                    fma2_bcast(gc[r][1], pv[j], G[j][1]);
                    const unsigned long long t = fma2(c.y, G[j][1], mul2(c.x, G[j][0]));   // ctx . G over this lane's 4 channels
                    gp[j] = pair_sum(t);
                }
                if (P.use_depth) {
                    // transposing butterfly over the 16 channel lanes: 8 values -> 1 per lane, summed over all 16 lanes
                    float a4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float send = b3 ? gp[i] : gp[i + 4];
                        const float keep = b3 ? gp[i + 4] : gp[i];
                        a4[i] = keep + __shfl_xor_sync(half_mask, send, 8);
                    }
                    float a2[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = b2 ? a4[i] : a4[i + 2];
                        const float keep = b2 ? a4[i + 2] : a4[i];
                        a2[i] = keep + __shfl_xor_sync(half_mask, send, 4);
                    }
                    const float send = b1 ? a2[0] : a2[1];
                    const float keep = b1 ? a2[1] : a2[0];
                    float a1 = keep + __shfl_xor_sync(half_mask, send, 2);
                    a1 += __shfl_xor_sync(half_mask, a1, 1);
                    if (!(cg & 1)) s_gprob[pix * DPAD + dblk * 8 + jsel] = a1;
                }
            }
        }
    }
    __syncthreads();    // prob / ctx fully consumed, g_prob complete

    // ---- g_ctx registers -> ctx region in the transposed layout [pix][c] ------------------------------------------------
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        const int h = h_lo + r;
        if (h < h_hi)
            *reinterpret_cast<ulonglong2*>(s_ctx + (wt * hh + h) * L.C + cg * 4) = make_ulonglong2(gc[r][0], gc[r][1]);
    }
    __syncthreads();

    // ---- softmax backward + transposes back to the raw [channel][row][col] layout (mirror of transform_tile) -----------
    const int n_pblk = (PX + 31) >> 5;
    const int n_cblk = L.C >> 5;
    const bool depth_unit = warp < n_pblk;
    const bool ctx_unit = !depth_unit && warp < n_pblk * (1 + n_cblk);
    const int pblk = depth_unit ? warp : (warp - n_pblk) % n_pblk;
    const int c0 = ctx_unit ? ((warp - n_pblk) / n_pblk) * 32 : 0;
    const int pixr = pblk * 32 + lane;                          // raw pixel index row*WT + col
    const bool active = (depth_unit || ctx_unit) && pixr < PX;
    const int pixT = active ? (pixr % WT) * hh + pixr / WT : 0;

    float v[DPAD > 32 ? DPAD : 32];
    if (depth_unit && active && P.use_depth) {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < DPAD; ++k) {
            int d = lane + k;
            d = (d >= DPAD) ? d - DPAD : d;
            const float pr = s_prob[pixT * PS + d];
            const float gp = (d < P.D) ? s_gprob[pixT * DPAD + d] : 0.f;
            dot = fmaf(pr, gp, dot);
            v[k] = gp;
        }
#pragma unroll
        for (int k = 0; k < DPAD; ++k) {
            int d = lane + k;
            d = (d >= DPAD) ? d - DPAD : d;
            v[k] = s_prob[pixT * PS + d] * (v[k] - dot);
        }
    } else if (ctx_unit && active) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = s_ctx[pixT * L.C + c0 + ((lane + k) & 31)];
    }
    __syncthreads();
    if (depth_unit && active && P.use_depth) {
#pragma unroll
        for (int k = 0; k < DPAD; ++k) {
            int d = lane + k;
            d = (d >= DPAD) ? d - DPAD : d;
            s_prob[d * PX + pixr] = v[k];
        }
    } else if (ctx_unit && active) {
#pragma unroll
        for (int k = 0; k < 32; ++k) s_ctx[(c0 + ((lane + k) & 31)) * PX + pixr] = v[k];
    }
    fence_proxy_async();       // generic-proxy writes -> visible to the TMA (async proxy)
    __syncthreads();
    if (tid == 0) {
        const int box_floats = CH_BOX * PX;
        if (P.use_depth)
            for (int i = 0; i < DBLKS; ++i)
                if (i * CH_BOX < P.D) tma_store_4d(&grad_maps.depth, s_prob + i * box_floats, w0, 0, i * CH_BOX, img);
        for (int i = 0; i < L.C / CH_BOX; ++i) tma_store_4d(&grad_maps.ctx, s_ctx + i * box_floats, w0, 0, i * CH_BOX, img);
        tma_store_commit_and_wait();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// grad_bev (B', C, X*Y) -> channel-last workspace (B', X*Y, C); mirror of finalize_nchw_kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TR_PILLARS = 64;
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long pillars, int blocks_per_frame) {
    __shared__ float tile[TR_PILLARS][65];
    const int frame = blockIdx.x / blocks_per_frame;
    const long long p0 = static_cast<long long>(blockIdx.x % blocks_per_frame) * TR_PILLARS;
    const int n_here = static_cast<int>(min(static_cast<long long>(TR_PILLARS), pillars - p0));
    const float* s = src + static_cast<size_t>(frame) * C * pillars + p0;
    for (int i = threadIdx.x; i < C * TR_PILLARS; i += 256) {
        const int c = i / TR_PILLARS, pl = i % TR_PILLARS;
        tile[pl][c] = (pl < n_here) ? s[static_cast<size_t>(c) * pillars + pl] : 0.f;
    }
    __syncthreads();
    float* d = dst + (static_cast<size_t>(frame) * pillars + p0) * C;
    for (int i = threadIdx.x; i < TR_PILLARS * 16; i += 256) {
        const int pl = i >> 4, q = i & 15;
        if (pl < n_here)
            reinterpret_cast<float4*>(d + static_cast<size_t>(pl) * C)[q] =
                make_float4(tile[pl][q * 4 + 0], tile[pl][q * 4 + 1], tile[pl][q * 4 + 2], tile[pl][q * 4 + 3]);
    }
}

int encode_head_maps(HeadMaps* maps, const void* head, int dtype, const LiftParams& P);

template <int DBLKS>
static int launch_backward_t(const HeadMaps& hm, const HeadMaps& gm, const LiftParams& P, cudaStream_t stream) {
    const TileLayout<DBLKS> L(P.hh, P.C);
    const int n_pblk = (L.PX + 31) / 32;
    FIERY_REQUIRE(n_pblk * (1 + P.C / 32) <= TileLayout<DBLKS>::NWARPS, "feature map too tall for this build: h=%d", P.hh);
    FIERY_REQUIRE((P.hh + DBLKS - 1) / DBLKS <= MAXR && P.hh <= 32, "feature map too tall for this build: h=%d", P.hh);
    const int smem = L.total + L.PX * TileLayout<DBLKS>::DPAD * 4;
    FIERY_REQUIRE(smem <= 227 * 1024, "tile needs %d bytes of shared memory", smem);
    static int smem_configured_on[64] = {};           // function attributes are per device
    int dev_id = 0;
    FIERY_CUDA_CHECK(cudaGetDevice(&dev_id));
    int& smem_configured = smem_configured_on[dev_id & 63];
    if (smem > smem_configured) {
        FIERY_CUDA_CHECK(cudaFuncSetAttribute(lift_backward_kernel<DBLKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_configured = smem;
    }
    const long long n_tiles = static_cast<long long>(P.n_frames) * P.n_cameras * P.n_wtiles;
    lift_backward_kernel<DBLKS><<<static_cast<unsigned>(n_tiles), 64 * DBLKS, smem, stream>>>(hm, gm, P);
    FIERY_CUDA_CHECK(cudaGetLastError());
    return FIERY_OK;
}

int launch_lift_backward(const LiftParams& P, const void* head, int head_dtype, float* workspace, cudaStream_t stream) {
    FIERY_REQUIRE(head_dtype == FIERY_DTYPE_F32, "head dtype %d not supported by this build (fp32 only)", head_dtype);
    FIERY_REQUIRE(P.C == 64, "channels=%d not supported by this build (C must be 64)", P.C);
    FIERY_REQUIRE(P.D >= 1 && P.D <= 48, "depth_bins=%d not supported by this build (1..48)", P.D);
    FIERY_REQUIRE(P.ww % 4 == 0, "feat_w=%d must be a multiple of 4 (TMA row pitch must be 16-byte aligned)", P.ww);
    HeadMaps hm, gm;
    int rc = encode_head_maps(&hm, head, head_dtype, P);
    if (rc != FIERY_OK) return rc;
    rc = encode_head_maps(&gm, P.grad_head, head_dtype, P);
    if (rc != FIERY_OK) return rc;
    LiftParams Q = P;
    if (P.bev_layout == FIERY_BEV_NCHW) {
        const int bpf = static_cast<int>((P.pillars + TR_PILLARS - 1) / TR_PILLARS);
        nchw_to_nhwc_kernel<<<static_cast<unsigned>(bpf) * P.n_frames, 256, 0, stream>>>(P.grad_bev, workspace, P.C, P.pillars, bpf);
        FIERY_CUDA_CHECK(cudaGetLastError());
        Q.grad_bev = workspace;
    }
    return launch_backward_t<6>(hm, gm, Q, stream);
}

}  // namespace fiery
