// Foreground pipeline of the fused loop: convex upsampling + target + descriptor-similarity weight over the per-call
// foreground list, on CHANNELS-LAST descriptors, producing one float4 (target x, target y, weight, depth) per listed pixel.
//   reference model/CFNet.py:95-106 (upsample_flow), model/PoseRefiner.py:335-345 (target, grid_sample warp, weight),
//   geometry/projective_ops.py:11-23; the arithmetic per pixel is that of upsample_weight.cu (same order for the flow).
//
// Why a second layout.  upsample_weight_kernel reads the two descriptor maps as 2 x 32 NCHW planes: one pixel touches
// 32 + 4 x 32 different 128-byte lines, each in a different plane (300 KB apart), and ncu put it at 27 % of HBM peak with
// long-scoreboard stalls (profiles/r1c_summary.md) -- the traffic equals the algorithmic bytes, it is the request
// pattern that is slow.  Here a pixel's 32 channels are ONE 128-byte line:
//   * geofea2 is transposed once per call to [B][H*W][32] (it does not change over the recurrent iterations,
//     PoseRefiner.py:292) -- or arrives that way from the zoom-crop kernel (zoom_crop.cu);
//   * geofea1 is only ever read at the foreground pixels, always the same ones: it is gathered once per call into list
//     order, [B][k][32], so every later read is a pure stream;
//   * 8 lanes share a pixel (4 channels = one 128-bit load each): a warp request is 4 full lines (4 L1 wavefronts, every
//     sector fully used), the similarity is an 8-lane shuffle reduction.
// The LM kernel (lm.cu, lm_cluster_kernel) then streams the float4 records.
#include "common.cuh"

namespace {

constexpr int GC = 32;                       // descriptor channels of the fast path
constexpr int TR_PX = 64;                    // pixels (or list entries) per transpose block

// NCHW [B][32][N] -> channels-last.  INDIRECT = false: dst[b][r][c] for every pixel r.  INDIRECT = true: dst[b][k][c] =
// src[b][c][fg_idx[b][k]] for k < fg_count[b] (list order).  src may be a mapped host pointer (reads are 128-byte coalesced).
template <bool INDIRECT>
__global__ void __launch_bounds__(256) geo_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int N,
                                                        const int* __restrict__ fg_idx, const int* __restrict__ fg_count) {
    __shared__ float tile[GC][TR_PX + 1];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y, k0 = blockIdx.x * TR_PX;
    const int count = INDIRECT ? fg_count[b] : N;
    if (k0 >= count) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int r[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = k0 + lane + 32 * u;
        r[u] = k < count ? (INDIRECT ? __ldg(fg_idx + (size_t)b * N + k) : k) : -1;
    }
    const float* sb = src + (size_t)b * GC * N;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int c = warp * 4 + cc;
#pragma unroll
        for (int u = 0; u < 2; ++u) tile[c][lane + 32 * u] = r[u] >= 0 ? __ldg(sb + (size_t)c * N + r[u]) : 0.f;
    }
    __syncthreads();
    const int cg = threadIdx.x & 7;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int t = (threadIdx.x >> 3) + 32 * u;
        if (k0 + t >= count) continue;
        const float4 v = make_float4(tile[4 * cg][t], tile[4 * cg + 1][t], tile[4 * cg + 2][t], tile[4 * cg + 3][t]);
        reinterpret_cast<float4*>(dst + ((size_t)b * N + k0 + t) * GC)[cg] = v;
    }
}

// Two kernels per recurrent iteration (one fused kernel was measured first: 175 us at B=32 -- 8 lanes per pixel leave only
// 4 pixels in flight per warp behind a three-deep dependent chain index -> mask/flow -> descriptor gather; profiles/r2b):
//
// (A) fg_target_kernel: one THREAD per listed pixel.  Convex upsampling (CFNet.py:95-106) and target = flow + grid
//     (PoseRefiner.py:335-338); needs only the mask and the low-resolution flow (both L2-resident, just written by the update
//     block).  rec[b][k] = (target x, target y, 0, depth).
__global__ void __launch_bounds__(256) fg_target_kernel(const float* __restrict__ flow, const float* __restrict__ mask,
                                                        const float* __restrict__ depth, int H, int W, const int* __restrict__ fg_idx,
                                                        const int* __restrict__ fg_count, float4* __restrict__ rec) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int count = fg_count[b];
    const int h = H >> 3, w = W >> 3, N = H * W;
    const int* idx = fg_idx + (size_t)b * N;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const int r = __ldg(idx + k);
        const int Y = r / W, X = r - Y * W;
        const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
        const size_t p = ((size_t)b * h + y) * w + x;
        const float dz = __ldg(depth + (size_t)b * N + r);
        // softmax over the 9 taps of mask[p][t*64 + i*8 + j], in the order of upsample_weight_kernel
        const float* mp = mask + p * 576 + i * 8 + j;
        float mk[9];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 9; ++t) { mk[t] = __ldg(mp + t * 64); mx = fmaxf(mx, mk[t]); }
        float2 fl[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int ny = y + t / 3 - 1, nx = x + t % 3 - 1;
            const int cy = min(max(ny, 0), h - 1), cx = min(max(nx, 0), w - 1);
            const float2 f = __ldg(reinterpret_cast<const float2*>(flow + (((size_t)b * h + cy) * w + cx) * 2));
            const bool in = ny >= 0 && ny < h && nx >= 0 && nx < w;
            fl[t] = in ? f : make_float2(0.f, 0.f);
        }
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) { mk[t] = expf(mk[t] - mx); den += mk[t]; }
        float ux = 0.f, uy = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float sm = mk[t] / den;
            ux += sm * (8.f * fl[t].x);
            uy += sm * (8.f * fl[t].y);
        }
        rec[(size_t)b * N + k] = make_float4(ux + (float)X, uy + (float)Y, 0.f, dz);
    }
}

// (B) fg_weight_kernel: one lane GROUP (8 lanes, 4 channels each) per listed pixel, FGW_U pixels per group and trip so that a
//     warp keeps 4 * FGW_U pixels x 5 lines in flight.  Descriptor warp at the target (normalize_coords_grid + grid_sample with
//     align_corners=False, zeros; PoseRefiner.py:343, projective_ops.py:11-23), similarity and weight exp(-|1 - s| / sigma)
//     (:344-345).  Writes rec[b][k].z and, on request, scatters the weight into a dense [B][H*W] map.
constexpr int FGW_U = 2;

__global__ void __launch_bounds__(256, 5) fg_weight_kernel(const float* __restrict__ g1c, const float* __restrict__ g2cl, float sigma, int H,
                                                           int W, const int* __restrict__ fg_idx, const int* __restrict__ fg_count,
                                                           float4* __restrict__ rec, float* __restrict__ weight_dense) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int count = fg_count[b];
    const int N = H * W;
    const int grp = threadIdx.x >> 3, cg = threadIdx.x & 7;            // 32 groups per block
    const float* g2b = g2cl + (size_t)b * N * GC;
    float4* recb = rec + (size_t)b * N;
    // warp-uniform trip count (the shuffles below need all 32 lanes): a group past the end repeats the last entry and does
    // not store.  Per trip a block covers 32 * FGW_U consecutive entries.
    const int stride = gridDim.x * 32 * FGW_U;
    for (int k0 = blockIdx.x * 32 * FGW_U + (grp & ~3) * FGW_U; k0 < count; k0 += stride) {
        int kk[FGW_U]; bool live[FGW_U];
        float4 t[FGW_U], a[FGW_U];
#pragma unroll
        for (int u = 0; u < FGW_U; ++u) {
            const int k = k0 + (grp & 3) * FGW_U + u;
            live[u] = k < count;
            kk[u] = live[u] ? k : count - 1;
            t[u] = recb[kk[u]];                                         // written by fg_target_kernel in the previous launch
            a[u] = __ldg(reinterpret_cast<const float4*>(g1c + ((size_t)b * N + kk[u]) * GC) + cg);
        }
        float4 c00[FGW_U], c01[FGW_U], c10[FGW_U], c11[FGW_U];
        float w00[FGW_U], w01[FGW_U], w10[FGW_U], w11[FGW_U];
        bool k00[FGW_U], k01[FGW_U], k10[FGW_U], k11[FGW_U];
#pragma unroll
        for (int u = 0; u < FGW_U; ++u) {
            const float tx = t[u].x, ty = t[u].y;
            const float gx = 2.f * tx / (float)(W - 1) - 1.f;
            const float gy = 2.f * ty / (float)(H - 1) - 1.f;
            const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
            const float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            // ix may be NaN / inf for a degenerate flow: every comparison false -> zero sample, like grid_sample
            const bool fin = isfinite(ix) && isfinite(iy);
            const int x0 = fin ? (int)fmaxf(fminf(fx0, 1e7f), -1e7f) : -100, y0 = fin ? (int)fmaxf(fminf(fy0, 1e7f), -1e7f) : -100;
            const float wnw = (fx0 + 1.f - ix) * (fy0 + 1.f - iy), wne = (ix - fx0) * (fy0 + 1.f - iy);
            const float wsw = (fx0 + 1.f - ix) * (iy - fy0), wse = (ix - fx0) * (iy - fy0);
            const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
            const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
            const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
            const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
            k00[u] = fin && ya && xa; k01[u] = fin && ya && xb; k10[u] = fin && yb && xa; k11[u] = fin && yb && xb;
            w00[u] = k00[u] ? wnw : 0.f; w01[u] = k01[u] ? wne : 0.f; w10[u] = k10[u] ? wsw : 0.f; w11[u] = k11[u] ? wse : 0.f;
            // four corners: one 128-byte line each, this lane's 4 channels
            c00[u] = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc0 * W + xc0) * GC) + cg);
            c01[u] = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc0 * W + xc1) * GC) + cg);
            c10[u] = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc1 * W + xc0) * GC) + cg);
            c11[u] = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc1 * W + xc1) * GC) + cg);
        }
#pragma unroll
        for (int u = 0; u < FGW_U; ++u) {
            // per channel: v = ((t00 w00 + t01 w01) + t10 w10) + t11 w11 as in upsample_weight_kernel (a discarded corner's
            // value is replaced by 0 so that a non-finite texel outside the sampled set cannot leak in)
            auto blend = [&](float q00, float q01, float q10, float q11) {
                float v = 0.f;
                v += (k00[u] ? q00 : 0.f) * w00[u];
                v += (k01[u] ? q01 : 0.f) * w01[u];
                v += (k10[u] ? q10 : 0.f) * w10[u];
                v += (k11[u] ? q11 : 0.f) * w11[u];
                return v;
            };
            float s = 0.f;
            s += a[u].x * blend(c00[u].x, c01[u].x, c10[u].x, c11[u].x);
            s += a[u].y * blend(c00[u].y, c01[u].y, c10[u].y, c11[u].y);
            s += a[u].z * blend(c00[u].z, c01[u].z, c10[u].z, c11[u].z);
            s += a[u].w * blend(c00[u].w, c01[u].w, c10[u].w, c11[u].w);
            // the 8 lanes of a group are consecutive lanes: xor 1, 2, 4 stay inside it
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const float wgt = t[u].w > 0.f ? expf(-fabsf(1.f - s) / sigma) : 0.f;
            if (cg == 0 && live[u]) {
                reinterpret_cast<float*>(recb + kk[u])[2] = wgt;
                if (weight_dense) weight_dense[(size_t)b * N + __ldg(fg_idx + (size_t)b * N + kk[u])] = wgt;
            }
        }
    }
}

}  // namespace

size_t b2p_fgpipe_ws_bytes(int B, int H, int W) {
    const size_t N = (size_t)B * H * W;
    return 2 * align_up(N * GC * sizeof(float), 1024) + align_up(N * sizeof(float4), 1024);
}

static inline void fgpipe_split(void* ws, int B, int H, int W, float** g2cl, float** g1c, float4** rec) {
    const size_t N = (size_t)B * H * W;
    char* p = reinterpret_cast<char*>(ws);
    *g2cl = reinterpret_cast<float*>(p); p += align_up(N * GC * sizeof(float), 1024);
    *g1c = reinterpret_cast<float*>(p); p += align_up(N * GC * sizeof(float), 1024);
    *rec = reinterpret_cast<float4*>(p);
}

const float4* b2p_fgpipe_records(const void* ws, int B, int H, int W) {
    float *a, *b; float4* r;
    fgpipe_split(const_cast<void*>(ws), B, H, W, &a, &b, &r);
    return r;
}

// once per call: geofea2 -> channels-last (skipped when g2_is_cl: the caller's buffer already is [B][H*W][32]), geofea1 ->
// list order.  fg_ws: the foreground list of b2p_fg_build.
int b2p_fgpipe_prepare(const float* g1, const float* g2, int g2_is_cl, int B, int H, int W, const void* fg_ws, void* ws, cudaStream_t s) {
    float *g2cl, *g1c; float4* rec;
    fgpipe_split(ws, B, H, W, &g2cl, &g1c, &rec);
    const int N = H * W;
    const dim3 grid((unsigned)ceil_div(N, TR_PX), (unsigned)B);
    if (!g2_is_cl) {
        B2P_CUDA(b2p_launch_pdl(geo_to_cl_kernel<false>, grid, dim3(256), 0, s, g2, g2cl, N, (const int*)nullptr, (const int*)nullptr));
        B2P_LAUNCH_CHECK();
    }
    B2P_CUDA(b2p_launch_pdl(geo_to_cl_kernel<true>, grid, dim3(256), 0, s, g1, g1c, N, b2p_fg_idx(fg_ws), b2p_fg_count(fg_ws, B, H, W)));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_fgpipe_upsample_weight(const float* flow, const float* mask, const float* g2_cl_or_null, const float* depth, float sigma, int B,
                               int H, int W, const void* fg_ws, void* ws, float* weight_dense, cudaStream_t s) {
    float *g2cl, *g1c; float4* rec;
    fgpipe_split(ws, B, H, W, &g2cl, &g1c, &rec);
    int dev = 0, sms = 0;
    B2P_CUDA(cudaGetDevice(&dev));
    B2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int* fg_idx = b2p_fg_idx(fg_ws);
    const int* fg_count = b2p_fg_count(fg_ws, B, H, W);
    // resident blocks only (each strides over its sample's list), shared by the samples
    int nA = (sms * 6) / B;
    const int usefulA = ceil_div(H * W, 256);
    nA = nA > usefulA ? usefulA : (nA < 1 ? 1 : nA);
    B2P_CUDA(b2p_launch_pdl(fg_target_kernel, dim3((unsigned)nA, (unsigned)B), dim3(256), 0, s, flow, mask, depth, H, W, fg_idx, fg_count, rec));
    B2P_LAUNCH_CHECK();
    int nB = (sms * 5) / B;
    const int usefulB = ceil_div(H * W, 32 * FGW_U);
    nB = nB > usefulB ? usefulB : (nB < 1 ? 1 : nB);
    B2P_CUDA(b2p_launch_pdl(fg_weight_kernel, dim3((unsigned)nB, (unsigned)B), dim3(256), 0, s, (const float*)g1c,
                            g2_cl_or_null ? g2_cl_or_null : (const float*)g2cl, sigma, H, W, fg_idx, fg_count, rec, weight_dense));
    B2P_LAUNCH_CHECK();
    return 0;
}
