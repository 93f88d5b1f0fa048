// Convex 8x upsampling of the low-resolution flow fused with target = flow + grid and the
// descriptor-similarity correspondence weight.
//   reference model/CFNet.py:95-106 (upsample_flow), model/PoseRefiner.py:335-345,
//   geometry/projective_ops.py:11-23 (normalize_coords_grid), F.grid_sample default
//   (align_corners=False, zeros padding) -- SURVEY Appendix A.5 / A.7.
//
// Two phases per warp (32 consecutive full-resolution pixels, x fastest):
//   1. one lane per pixel: softmax over the 9 mask taps, convex combination of the 3x3 low-res flows, target.
//      (Doing this once per pixel matters: it is ~450 instructions, and replicating it over the lanes that share a
//      pixel in phase 2 made an earlier version instruction-bound.)
//   2. eight lanes per pixel, four pixels per step, eight steps: each lane gathers the 4 bilinear corners of C/8
//      descriptor channels of the owner pixel's target (broadcast by shuffle) and the partial dot products are
//      reduced with three shuffles.  Steps without a foreground pixel are skipped (warp-uniform).
//      (One lane per pixel serialises 32 channels x 5 dependent-latency loads; eight lanes keep them in flight.)
// Mask reads are 32-B-sector exact, descriptor planes (NCHW) are read with per-plane locality.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) upsample_weight_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int B, int C, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background) {
    const int h = H >> 3, w = W >> 3;
    const size_t N = (size_t)H * W;
    const size_t total = (size_t)B * N;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t warp_base = idx - lane;
    const bool in_range = idx < total;
    const size_t idc = in_range ? idx : 0;
    const int b = (int)(idc / N);
    const int r = (int)(idc - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;
    const float dz = depth ? __ldg(depth + idc) : 1.f;

    // ---- phase 1.  Fused-loop shortcut: a background pixel (syn_depth <= 0) has weight exactly 0, so the LM step
    // ignores its target (any finite value contributes 0 * finite = 0, as in the reference).  When the up-sampled flow
    // itself is not an output of this iteration, skip the mask softmax and the descriptor warp for it.
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
    if (in_range) {
        if (flow_up) {
            flow_up[((size_t)b * 2 + 0) * N + r] = ux;
            flow_up[((size_t)b * 2 + 1) * N + r] = uy;
        }
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2(tx, ty);
    }
    if (!weight) return;                                // uniform over the grid

    // ---- phase 2
    const bool fg = in_range && !lazy && dz > 0.f;
    if (in_range && !fg) weight[idx] = 0.f;
    const unsigned fgmask = __ballot_sync(0xffffffffu, fg);
    if (fgmask == 0u) return;                           // warp-uniform
    const int sub = lane & 7, grp = lane >> 3;
#pragma unroll 1
    for (int step = 0; step < 8; ++step) {
        if (((fgmask >> (step * 4)) & 0xFu) == 0u) continue;     // warp-uniform
        const int owner = step * 4 + grp;
        const float otx = __shfl_sync(0xffffffffu, tx, owner);
        const float oty = __shfl_sync(0xffffffffu, ty, owner);
        const bool ofg = (fgmask >> owner) & 1u;
        float s = 0.f;
        if (ofg) {
            const size_t oidx = warp_base + owner;
            const int ob = (int)(oidx / N);
            const size_t orr = oidx - (size_t)ob * N;
            // normalize_coords_grid then grid_sample's align_corners=False un-normalisation
            const float gx = 2.f * otx / (float)(W - 1) - 1.f;
            const float gy = 2.f * oty / (float)(H - 1) - 1.f;
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
            const float* g1p = g1 + (size_t)ob * C * N + orr;
            const float* g2p = g2 + (size_t)ob * C * N;
#pragma unroll 4
            for (int c = sub; c < C; c += 8) {
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
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ofg && sub == 0) weight[warp_base + owner] = expf(-fabsf(1.f - s) / sigma);
    }
}

}  // namespace

int b2p_upsample_weight(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth,
                        float sigma, int B, int C, int H, int W, float* flow_up, float* target, float* weight,
                        int lazy_background, cudaStream_t s) {
    const size_t total = (size_t)B * H * W;
    upsample_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(flow, mask, g1, g2, depth, sigma, B, C, H, W,
                                                                           flow_up, target, weight, lazy_background);
    B2P_LAUNCH_CHECK();
    return 0;
}
