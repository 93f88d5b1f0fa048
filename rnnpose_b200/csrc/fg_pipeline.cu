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

// One lane group (8 lanes) per listed pixel, 4 pixels per warp; blocks stride over the sample's list.
//   rec[b][k] = (target x, target y, weight, depth) of the k-th foreground pixel.
__global__ void __launch_bounds__(256, 4) upsample_weight_cl_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1c, const float* __restrict__ g2cl,
    const float* __restrict__ depth, float sigma, int H, int W, const int* __restrict__ fg_idx, const int* __restrict__ fg_count,
    float4* __restrict__ rec, float* __restrict__ weight_dense /* optional [B][H*W]: scatter of the weights (background untouched) */) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int count = fg_count[b];
    const int h = H >> 3, w = W >> 3, N = H * W;
    const int grp = threadIdx.x >> 3, cg = threadIdx.x & 7;            // 32 groups per block
    const int* idx = fg_idx + (size_t)b * N;
    const float* g2b = g2cl + (size_t)b * N * GC;
    const float* dep = depth + (size_t)b * N;
    // warp-uniform trip count (the shuffles below need all 32 lanes): a group past the end repeats the last entry and
    // does not store.  The next entry's pixel index is fetched one trip ahead.
    const int stride = gridDim.x * 32;
    const int k_first = blockIdx.x * 32 + (grp & ~3);                  // first entry of this warp's four groups
    int r_next = (count > 0) ? __ldg(idx + min(k_first + (grp & 3), count - 1)) : 0;
    for (int k0 = k_first; k0 < count; k0 += stride) {
        const int k = k0 + (grp & 3);
        const bool live = k < count;
        const int kk = live ? k : count - 1;
        const int r = r_next;
        if (k0 + stride < count) r_next = __ldg(idx + min(k + stride, count - 1));
        // descriptor of the rendered view at this pixel: independent of everything below, issued first
        const float4 a = __ldg(reinterpret_cast<const float4*>(g1c + ((size_t)b * N + kk) * GC) + cg);
        const int Y = r / W, X = r - Y * W;
        const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
        const size_t p = ((size_t)b * h + y) * w + x;
        const float dz = __ldg(dep + r);
        // convex upsampling (CFNet.py:95-106): softmax over the 9 taps of mask[p][k*64 + i*8 + j]; the 8 lanes of the group
        // compute it redundantly (same addresses: one request), in the order of upsample_weight_kernel
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
        const float tx = ux + (float)X, ty = uy + (float)Y;
        // normalize_coords_grid then grid_sample's align_corners=False un-normalisation (PoseRefiner.py:343)
        const float gx = 2.f * tx / (float)(W - 1) - 1.f;
        const float gy = 2.f * ty / (float)(H - 1) - 1.f;
        const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
        const float iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float wnw = (fx0 + 1.f - ix) * (fy0 + 1.f - iy);
        const float wne = (ix - fx0) * (fy0 + 1.f - iy);
        const float wsw = (fx0 + 1.f - ix) * (iy - fy0);
        const float wse = (ix - fx0) * (iy - fy0);
        const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
        const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
        const bool fin = isfinite(ix) && isfinite(iy);
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        const bool k00 = fin && ya && xa, k01 = fin && ya && xb, k10 = fin && yb && xa, k11 = fin && yb && xb;
        const float w00 = k00 ? wnw : 0.f, w01 = k01 ? wne : 0.f, w10 = k10 ? wsw : 0.f, w11 = k11 ? wse : 0.f;
        // four corners: one 128-byte line each, this lane's 4 channels
        const float4 t00 = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc0 * W + xc0) * GC) + cg);
        const float4 t01 = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc0 * W + xc1) * GC) + cg);
        const float4 t10 = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc1 * W + xc0) * GC) + cg);
        const float4 t11 = __ldg(reinterpret_cast<const float4*>(g2b + (size_t)(yc1 * W + xc1) * GC) + cg);
        // per channel: v = ((t00 w00 + t01 w01) + t10 w10) + t11 w11 as in upsample_weight_kernel (a discarded corner's
        // value is replaced by 0 so that a non-finite texel outside the sampled set cannot leak in)
        auto blend = [&](float c00, float c01, float c10, float c11) {
            float v = 0.f;
            v += (k00 ? c00 : 0.f) * w00;
            v += (k01 ? c01 : 0.f) * w01;
            v += (k10 ? c10 : 0.f) * w10;
            v += (k11 ? c11 : 0.f) * w11;
            return v;
        };
        float s = 0.f;
        s += a.x * blend(t00.x, t01.x, t10.x, t11.x);
        s += a.y * blend(t00.y, t01.y, t10.y, t11.y);
        s += a.z * blend(t00.z, t01.z, t10.z, t11.z);
        s += a.w * blend(t00.w, t01.w, t10.w, t11.w);
        // the 8 lanes of a group are consecutive lanes: xor 1, 2, 4 stay inside it
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float wgt = dz > 0.f ? expf(-fabsf(1.f - s) / sigma) : 0.f;
        if (cg == 0 && live) {
            rec[(size_t)b * N + k] = make_float4(tx, ty, wgt, dz);
            if (weight_dense) weight_dense[(size_t)b * N + r] = wgt;
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
    // resident blocks only (each strides over its sample's list): 4 blocks of 256 threads per SM shared by the samples
    int per_sample = (sms * 4) / B;
    const int useful = ceil_div(H * W, 32);
    if (per_sample > useful) per_sample = useful;
    if (per_sample < 1) per_sample = 1;
    B2P_CUDA(b2p_launch_pdl(upsample_weight_cl_kernel, dim3((unsigned)per_sample, (unsigned)B), dim3(256), 0, s, flow, mask, (const float*)g1c,
                            g2_cl_or_null ? g2_cl_or_null : (const float*)g2cl, depth, sigma, H, W, b2p_fg_idx(fg_ws),
                            b2p_fg_count(fg_ws, B, H, W), rec, weight_dense));
    B2P_LAUNCH_CHECK();
    return 0;
}
