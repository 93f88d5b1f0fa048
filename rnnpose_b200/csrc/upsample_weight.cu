// Convex 8x upsampling of the low-resolution flow fused with target = flow + grid and the
// descriptor-similarity correspondence weight.
//   reference model/CFNet.py:95-106 (upsample_flow), model/PoseRefiner.py:335-345,
//   geometry/projective_ops.py:11-23 (normalize_coords_grid), F.grid_sample default
//   (align_corners=False, zeros padding) -- SURVEY Appendix A.5 / A.7.
// One thread per full-resolution pixel, x fastest: every descriptor request of a warp is 32 consecutive floats of ONE
// plane (1-2 L1 wavefronts; sharing a pixel between lanes multiplies the wavefronts and made earlier versions
// L1-bound, ncu: 8.7 sectors per request).  The kernel is then latency-bound (depth -> mask/flow -> target -> 4 batches
// of 8 channels x 5 loads), so blocks are only 2 warps: background warps exit at once (lazy shortcut) and a small block
// frees its slot as soon as its own warps finish, which roughly doubles the resident foreground warps.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(64) upsample_weight_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int B, int C, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background) {
    pdl_trigger();
    pdl_wait();
    const int h = H >> 3, w = W >> 3;
    const size_t N = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * N) return;
    const int b = (int)(idx / N);
    const int r = (int)(idx - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;

    // Fused-loop shortcut: a background pixel (syn_depth <= 0) has weight exactly 0, so the LM step ignores its target
    // (any finite value contributes 0 * finite = 0, as in the reference).  When the up-sampled flow itself is not an
    // output of this iteration, skip the mask softmax and the descriptor warp for it.
    if (lazy_background && !flow_up && depth[idx] <= 0.f) {
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2((float)X, (float)Y);
        if (weight) weight[idx] = 0.f;
        return;
    }

    // softmax over the 9 taps of mask[p][k*64 + i*8 + j]
    const float* mp = mask + p * 576 + i * 8 + j;
    float mk[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mk[k] = __ldg(mp + k * 64); mx = fmaxf(mx, mk[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mk[k] = expf(mk[k] - mx); den += mk[k]; }
    float ux = 0.f, uy = 0.f;
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
    if (flow_up) {
        flow_up[((size_t)b * 2 + 0) * N + r] = ux;
        flow_up[((size_t)b * 2 + 1) * N + r] = uy;
    }
    const float tx = ux + (float)X, ty = uy + (float)Y;
    if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2(tx, ty);
    if (!weight) return;

    const float dz = depth[idx];
    float wgt = 0.f;
    if (dz > 0.f) {
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
        const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
        const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
        // ix may be NaN/inf for degenerate flow: all comparisons false -> zero sample, like grid_sample
        const bool fin = isfinite(ix) && isfinite(iy);
        const size_t o00 = (size_t)y0 * W + x0;
        float s = 0.f;
        const float* g1p = g1 + (size_t)b * C * N + r;
        const float* g2p = g2 + (size_t)b * C * N;
#pragma unroll 8
        for (int c = 0; c < C; ++c) {
            const float* pl = g2p + (size_t)c * N;
            float v = 0.f;
            if (fin) {
                if (ya && xa) v += __ldg(pl + o00) * wnw;
                if (ya && xb) v += __ldg(pl + o00 + 1) * wne;
                if (yb && xa) v += __ldg(pl + o00 + W) * wsw;
                if (yb && xb) v += __ldg(pl + o00 + W + 1) * wse;
            }
            s += __ldg(g1p + (size_t)c * N) * v;
        }
        wgt = expf(-fabsf(1.f - s) / sigma);
    }
    weight[idx] = wgt;
}

}  // namespace

int b2p_upsample_weight(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth,
                        float sigma, int B, int C, int H, int W, float* flow_up, float* target, float* weight,
                        int lazy_background, cudaStream_t s) {
    const size_t total = (size_t)B * H * W;
    B2P_CUDA(b2p_launch_pdl(upsample_weight_kernel, dim3((unsigned)((total + 63) / 64)), dim3(64), 0, s, flow, mask, g1, g2, depth, sigma, B,
                            C, H, W, flow_up, target, weight, lazy_background));
    B2P_LAUNCH_CHECK();
    return 0;
}
