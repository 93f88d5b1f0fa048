// Convex 8x upsampling of the low-resolution flow fused with target = flow + grid and the
// descriptor-similarity correspondence weight.
//   reference model/CFNet.py:95-106 (upsample_flow), model/PoseRefiner.py:335-345,
//   geometry/projective_ops.py:11-23 (normalize_coords_grid), F.grid_sample default
//   (align_corners=False, zeros padding) -- SURVEY Appendix A.5 / A.7.
// Four consecutive lanes share one full-resolution pixel (x fastest): each takes a quarter of the descriptor
// channels (a 4-tap gather per channel) and the partial dot products meet in two shuffles.  The kernel is
// latency-bound (dependent mask -> target -> gather chain), so the 4x thread count is what buys memory-level
// parallelism; mask reads stay 32-B-sector exact, descriptor planes (NCHW) are read coalesced.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) upsample_weight_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int B, int C, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background) {
    const int h = H >> 3, w = W >> 3;
    const size_t N = (size_t)H * W;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t idx = tid >> 2;                       // pixel
    const int sub = (int)(tid & 3);                    // channel quarter
    const bool in_range = idx < (size_t)B * N;
    const size_t idc = in_range ? idx : 0;
    const int b = (int)(idc / N);
    const int r = (int)(idc - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;
    const float dz = depth ? __ldg(depth + idc) : 1.f;

    // Fused-loop shortcut: a background pixel (syn_depth <= 0) has weight exactly 0, so the LM step ignores its target
    // (any finite value contributes 0 * finite = 0, as in the reference).  When the up-sampled flow itself is not an
    // output of this iteration, skip the mask softmax and the descriptor warp for it.
    const bool lazy = lazy_background && !flow_up && dz <= 0.f;
    float tx = (float)X, ty = (float)Y, ux = 0.f, uy = 0.f;
    if (in_range && !lazy) {
        // softmax over the 9 taps of mask[p][k*64 + i*8 + j]
        const float* mp = mask + p * 576 + i * 8 + j;
        float mk[9];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mk[k] = __ldg(mp + k * 64); mx = fmaxf(mx, mk[k]); }
        float den = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mk[k] = expf(mk[k] - mx); den += mk[k]; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
            float2 f = make_float2(0.f, 0.f);
            if (ny >= 0 && ny < h && nx >= 0 && nx < w)
                f = __ldg(reinterpret_cast<const float2*>(flow + (((size_t)b * h + ny) * w + nx) * 2));
            const float sm = mk[k] / den;
            ux += sm * (8.f * f.x);
            uy += sm * (8.f * f.y);
        }
        tx = ux + (float)X; ty = uy + (float)Y;
    }
    if (in_range && sub == 0) {
        if (flow_up) {
            flow_up[((size_t)b * 2 + 0) * N + r] = ux;
            flow_up[((size_t)b * 2 + 1) * N + r] = uy;
        }
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2(tx, ty);
    }
    if (!weight) return;                                // uniform over the grid

    float s = 0.f;
    const bool fg = in_range && !lazy && dz > 0.f;
    if (fg) {
        // normalize_coords_grid then grid_sample's align_corners=False un-normalisation
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
        // ix may be NaN/inf for degenerate flow: no corner is in bounds -> zero sample, like grid_sample
        const bool fin = isfinite(ix) && isfinite(iy);
        const bool xa = fin && x0 >= 0 && x0 < W, xb = fin && x0 + 1 >= 0 && x0 + 1 < W;
        const bool ya = fin && y0 >= 0 && y0 < H, yb = fin && y0 + 1 >= 0 && y0 + 1 < H;
        const size_t o00 = (size_t)y0 * W + x0;
        const float* g1p = g1 + (size_t)b * C * N + r;
        const float* g2p = g2 + (size_t)b * C * N;
#pragma unroll 8
        for (int c = sub; c < C; c += 4) {
            const float* pl = g2p + (size_t)c * N;
            float v = 0.f;
            if (ya && xa) v += __ldg(pl + o00) * wnw;
            if (ya && xb) v += __ldg(pl + o00 + 1) * wne;
            if (yb && xa) v += __ldg(pl + o00 + W) * wsw;
            if (yb && xb) v += __ldg(pl + o00 + W + 1) * wse;
            s += __ldg(g1p + (size_t)c * N) * v;
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (in_range && sub == 0) weight[idx] = fg ? expf(-fabsf(1.f - s) / sigma) : 0.f;
}


// ------------------------------------------------------------------------------------------------------------------
// Channels-last variant used by the fused loop.  The NCHW descriptor planes make every channel of a gather land in a
// different DRAM page (40 scattered 4-byte requests per pixel); with the descriptors transposed once per call to
// [B][H][W][C] a bilinear corner is ONE 128-byte line (C = 32) read by 8 lanes as float4.
// ------------------------------------------------------------------------------------------------------------------
// NCHW [B][C][N] -> NHWC [B][N][C]; with only_fg the background pixels (depth <= 0) are skipped (never read later).
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, const float* __restrict__ depth,
                                                           int C, size_t N, int only_fg, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const size_t n0 = (size_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t n = n0 + tx;
    for (int cc = ty; cc < 32; cc += 8)
        tile[cc][tx] = (n < N && c0 + cc < C) ? __ldg(src + ((size_t)b * C + c0 + cc) * N + n) : 0.f;
    __syncthreads();
    for (int pp = ty; pp < 32; pp += 8) {
        const size_t pn = n0 + pp;
        if (pn >= N || c0 + tx >= C) continue;
        if (only_fg && !(__ldg(depth + (size_t)b * N + pn) > 0.f)) continue;
        dst[((size_t)b * N + pn) * C + c0 + tx] = tile[tx][pp];
    }
}

// 8 lanes per pixel, 4 channels (one float4) per lane; requires C == 32.
__global__ void __launch_bounds__(256) upsample_weight_nhwc_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1t,
    const float* __restrict__ g2t, const float* __restrict__ depth, float sigma, int B, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background) {
    const int h = H >> 3, w = W >> 3;
    const size_t N = (size_t)H * W;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t idx = tid >> 3;                       // pixel
    const int sub = (int)(tid & 7);                    // channels 4*sub .. 4*sub+3
    const bool in_range = idx < (size_t)B * N;
    const size_t idc = in_range ? idx : 0;
    const int b = (int)(idc / N);
    const int r = (int)(idc - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;
    const float dz = __ldg(depth + idc);
    const bool lazy = lazy_background && !flow_up && dz <= 0.f;
    float tx = (float)X, ty = (float)Y, ux = 0.f, uy = 0.f;
    if (in_range && !lazy) {
        const float* mp = mask + p * 576 + i * 8 + j;
        float mk[9];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mk[k] = __ldg(mp + k * 64); mx = fmaxf(mx, mk[k]); }
        float den = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { mk[k] = expf(mk[k] - mx); den += mk[k]; }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
            float2 f = make_float2(0.f, 0.f);
            if (ny >= 0 && ny < h && nx >= 0 && nx < w)
                f = __ldg(reinterpret_cast<const float2*>(flow + (((size_t)b * h + ny) * w + nx) * 2));
            const float sm = mk[k] / den;
            ux += sm * (8.f * f.x);
            uy += sm * (8.f * f.y);
        }
        tx = ux + (float)X; ty = uy + (float)Y;
    }
    if (in_range && sub == 0) {
        if (flow_up) {
            flow_up[((size_t)b * 2 + 0) * N + r] = ux;
            flow_up[((size_t)b * 2 + 1) * N + r] = uy;
        }
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2(tx, ty);
    }
    float s = 0.f;
    const bool fg = in_range && !lazy && dz > 0.f;
    if (fg) {
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
        const bool fin = isfinite(ix) && isfinite(iy);
        const bool xa = fin && x0 >= 0 && x0 < W, xb = fin && x0 + 1 >= 0 && x0 + 1 < W;
        const bool ya = fin && y0 >= 0 && y0 < H, yb = fin && y0 + 1 >= 0 && y0 + 1 < H;
        const float4* q = reinterpret_cast<const float4*>(g2t + ((size_t)b * N + (size_t)y0 * W + x0) * 32) + sub;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = __ldg(reinterpret_cast<const float4*>(g1t + ((size_t)b * N + r) * 32) + sub);
        const float4 cnw = (ya && xa) ? __ldg(q) : zero;
        const float4 cne = (ya && xb) ? __ldg(q + 8) : zero;                   // next pixel = +32 floats = +8 float4
        const float4 csw = (yb && xa) ? __ldg(q + (size_t)W * 8) : zero;
        const float4 cse = (yb && xb) ? __ldg(q + (size_t)W * 8 + 8) : zero;
        // same operation order per channel as the NCHW kernel: v = nw*w + ne*w + sw*w + se*w ; s += g1 * v
        float v;
        v = 0.f; if (ya && xa) v += cnw.x * wnw; if (ya && xb) v += cne.x * wne; if (yb && xa) v += csw.x * wsw; if (yb && xb) v += cse.x * wse; s += a.x * v;
        v = 0.f; if (ya && xa) v += cnw.y * wnw; if (ya && xb) v += cne.y * wne; if (yb && xa) v += csw.y * wsw; if (yb && xb) v += cse.y * wse; s += a.y * v;
        v = 0.f; if (ya && xa) v += cnw.z * wnw; if (ya && xb) v += cne.z * wne; if (yb && xa) v += csw.z * wsw; if (yb && xb) v += cse.z * wse; s += a.z * v;
        v = 0.f; if (ya && xa) v += cnw.w * wnw; if (ya && xb) v += cne.w * wne; if (yb && xa) v += csw.w * wsw; if (yb && xb) v += cse.w * wse; s += a.w * v;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (in_range && sub == 0) weight[idx] = fg ? expf(-fabsf(1.f - s) / sigma) : 0.f;
}

}  // namespace

int b2p_upsample_weight(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth,
                        float sigma, int B, int C, int H, int W, float* flow_up, float* target, float* weight,
                        int lazy_background, cudaStream_t s) {
    const size_t total = (size_t)B * H * W * 4;
    upsample_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(flow, mask, g1, g2, depth, sigma, B, C, H, W,
                                                                           flow_up, target, weight, lazy_background);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_nchw_to_nhwc(const float* src, const float* depth, int B, int C, int H, int W, int only_fg, float* dst, cudaStream_t s) {
    const size_t N = (size_t)H * W;
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((C + 31) / 32), B);
    nchw_to_nhwc_kernel<<<grid, 256, 0, s>>>(src, depth, C, N, only_fg, dst);
    B2P_LAUNCH_CHECK();
    return 0;
}

// g1t, g2t: channels-last descriptors [B][H][W][32] (b2p_nchw_to_nhwc); depth and weight are required.
int b2p_upsample_weight_nhwc(const float* flow, const float* mask, const float* g1t, const float* g2t, const float* depth,
                             float sigma, int B, int H, int W, float* flow_up, float* target, float* weight,
                             int lazy_background, cudaStream_t s) {
    const size_t total = (size_t)B * H * W * 8;
    upsample_weight_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(flow, mask, g1t, g2t, depth, sigma, B, H, W,
                                                                                flow_up, target, weight, lazy_background);
    B2P_LAUNCH_CHECK();
    return 0;
}
