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

#include <stdlib.h>

namespace {

// A read-only load the compiler may not reorder against its siblings or sink to its use (volatile asm): the staged variants
// below want all loads of a channel batch issued back to back.
__device__ __forceinline__ float ldg_ordered(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// One channel of the descriptor similarity: s + a * (bilinear sample).  The operation sequence is pinned with intrinsics so that
// every build of the kernel (and the list-driven one) rounds identically, whatever the compiler would contract or reorder.
__device__ __forceinline__ float corner_dot(float s, float a, float t00, float t01, float t10, float t11, float w00, float w01,
                                            float w10, float w11) {
    float v = __fmul_rn(t00, w00);
    v = __fmaf_rn(t01, w01, v);
    v = __fmaf_rn(t10, w10, v);
    v = __fmaf_rn(t11, w11, v);
    return __fmaf_rn(a, v, s);
}

// Everything for one full-resolution pixel (b, Y, X): convex upsampling, target, descriptor similarity weight.
// NPLANE: H*W when known at compile time (the plane stride of the descriptor loads becomes an immediate offset: 5 loads per
// channel with no address arithmetic), 0 = run-time size.
// BATCH: 0 = one loop over the channels (the compiler picks the load schedule), 8 / 16 = channels are loaded BATCH at a time into
// registers before any of them is used (5 * BATCH independent loads in flight per thread).
template <int NPLANE = 0, int BATCH = 0>
__device__ __forceinline__ void upsample_weight_pixel(const float* __restrict__ flow, const float* __restrict__ mask,
                                                      const float* __restrict__ g1, const float* __restrict__ g2,
                                                      const float* __restrict__ depth, float sigma, int b, int Y, int X, int C, int H,
                                                      int W, float* __restrict__ flow_up, float* __restrict__ target,
                                                      float* __restrict__ weight) {
    const int h = H >> 3, w = W >> 3;
    const size_t N = NPLANE ? (size_t)NPLANE : (size_t)H * W;
    const int r = Y * W + X;
    const size_t idx = (size_t)b * N + r;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;
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
    // (loads are unconditional from clamped addresses and the out-of-image taps are zeroed by a select: a load behind a
    //  branch cannot be hoisted, and the kernel was serialising on one memory round trip per tap)
    float2 fl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
        const int cy = min(max(ny, 0), h - 1), cx = min(max(nx, 0), w - 1);
        const float2 f = __ldg(reinterpret_cast<const float2*>(flow + (((size_t)b * h + cy) * w + cx) * 2));
        const bool in = ny >= 0 && ny < h && nx >= 0 && nx < w;
        fl[k] = in ? f : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float sm = mk[k] / den;
        ux += sm * (8.f * fl[k].x);
        uy += sm * (8.f * fl[k].y);
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
        // Branch-free gather: every corner is loaded from an in-image (clamped) address and discarded by a select when
        // the reference would not sample it (outside the image, or non-finite coordinates -> all four).  Same arithmetic
        // and order as the conditional form; the 5 x C loads of a pixel are independent and can all be in flight.
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        const int o00 = yc0 * W + xc0, o01 = yc0 * W + xc1, o10 = yc1 * W + xc0, o11 = yc1 * W + xc1;
        const bool k00 = fin && ya && xa, k01 = fin && ya && xb, k10 = fin && yb && xa, k11 = fin && yb && xb;
        const float w00 = k00 ? wnw : 0.f, w01 = k01 ? wne : 0.f, w10 = k10 ? wsw : 0.f, w11 = k11 ? wse : 0.f;
        float s = 0.f;
        const float* g1p = g1 + (size_t)b * C * N + r;
        const float* g2p = g2 + (size_t)b * C * N;
        int c = 0;
        if (BATCH > 0) {
            constexpr int NB = BATCH > 0 ? BATCH : 1;
            for (; c + NB <= C; c += NB) {
                float t[NB][4], a[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const float* pl = g2p + (size_t)(c + k) * N;
                    t[k][0] = ldg_ordered(pl + o00); t[k][1] = ldg_ordered(pl + o01); t[k][2] = ldg_ordered(pl + o10); t[k][3] = ldg_ordered(pl + o11);
                    a[k] = ldg_ordered(g1p + (size_t)(c + k) * N);
                }
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    s = corner_dot(s, a[k], k00 ? t[k][0] : 0.f, k01 ? t[k][1] : 0.f, k10 ? t[k][2] : 0.f, k11 ? t[k][3] : 0.f, w00, w01, w10, w11);
            }
        }
#pragma unroll 8
        for (; c < C; ++c) {
            const float* pl = g2p + (size_t)c * N;
            const float t00 = __ldg(pl + o00), t01 = __ldg(pl + o01), t10 = __ldg(pl + o10), t11 = __ldg(pl + o11);
            const float a = __ldg(g1p + (size_t)c * N);
            s = corner_dot(s, a, k00 ? t00 : 0.f, k01 ? t01 : 0.f, k10 ? t10 : 0.f, k11 ? t11 : 0.f, w00, w01, w10, w11);
        }
        wgt = expf(-fabsf(1.f - s) / sigma);
    }
    weight[idx] = wgt;
}

template <int NPLANE, int BATCH>
__global__ void __launch_bounds__(64) upsample_weight_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int B, int C, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background) {
    pdl_trigger();
    pdl_wait();
    const size_t N = NPLANE ? (size_t)NPLANE : (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * N) return;
    const int b = (int)(idx / N);
    const int r = (int)(idx - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    // Fused-loop shortcut: a background pixel (syn_depth <= 0) has weight exactly 0, so the LM step ignores its target
    // (any finite value contributes 0 * finite = 0, as in the reference).  When the up-sampled flow itself is not an
    // output of this iteration, skip the mask softmax and the descriptor warp for it.
    if (lazy_background && !flow_up && depth[idx] <= 0.f) {
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2((float)X, (float)Y);
        if (weight) weight[idx] = 0.f;
        return;
    }
    upsample_weight_pixel<NPLANE, BATCH>(flow, mask, g1, g2, depth, sigma, b, Y, X, C, H, W, flow_up, target, weight);
}

// The same pixel routine for the host entry's windowed second descriptor map.  It is a separate copy on purpose: with the
// window logic as a template flag of upsample_weight_pixel, the DEFAULT build of the kernel came out 20 us slower (108 vs 88 us,
// same source after dead-code elimination, a different ptxas schedule; A/B on one box in profiles/r2w).
// NPLANE: H*W when known at compile time (the plane stride of the descriptor loads becomes an immediate offset: 5 loads per
// channel with no address arithmetic), 0 = run-time size.
// BATCH: 0 = one loop over the channels (the compiler picks the load schedule), 8 / 16 = channels are loaded BATCH at a time into
// registers before any of them is used (5 * BATCH independent loads in flight per thread).
// WINDOW: g2 holds valid data only inside the per-sample window win[b] = (y0, y1, x0, x1) (what the host entry copied over PCIe:
// the foreground box plus a margin); a corner that is sampled outside it is read from g2_far, the same map in mapped host memory.
template <int NPLANE = 0, int BATCH = 0, bool WINDOW = true>
__device__ __forceinline__ void upsample_weight_pixel_win(const float* __restrict__ flow, const float* __restrict__ mask,
                                                      const float* __restrict__ g1, const float* __restrict__ g2,
                                                      const float* __restrict__ depth, float sigma, int b, int Y, int X, int C, int H,
                                                      int W, float* __restrict__ flow_up, float* __restrict__ target,
                                                      float* __restrict__ weight, const float* __restrict__ g2_far = nullptr,
                                                      const int* __restrict__ win = nullptr) {
    const int h = H >> 3, w = W >> 3;
    const size_t N = NPLANE ? (size_t)NPLANE : (size_t)H * W;
    const int r = Y * W + X;
    const size_t idx = (size_t)b * N + r;
    const int y = Y >> 3, i = Y & 7, x = X >> 3, j = X & 7;
    const size_t p = ((size_t)b * h + y) * w + x;
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
    // (loads are unconditional from clamped addresses and the out-of-image taps are zeroed by a select: a load behind a
    //  branch cannot be hoisted, and the kernel was serialising on one memory round trip per tap)
    float2 fl[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
        const int cy = min(max(ny, 0), h - 1), cx = min(max(nx, 0), w - 1);
        const float2 f = __ldg(reinterpret_cast<const float2*>(flow + (((size_t)b * h + cy) * w + cx) * 2));
        const bool in = ny >= 0 && ny < h && nx >= 0 && nx < w;
        fl[k] = in ? f : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float sm = mk[k] / den;
        ux += sm * (8.f * fl[k].x);
        uy += sm * (8.f * fl[k].y);
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
        // Branch-free gather: every corner is loaded from an in-image (clamped) address and discarded by a select when
        // the reference would not sample it (outside the image, or non-finite coordinates -> all four).  Same arithmetic
        // and order as the conditional form; the 5 x C loads of a pixel are independent and can all be in flight.
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        const int o00 = yc0 * W + xc0, o01 = yc0 * W + xc1, o10 = yc1 * W + xc0, o11 = yc1 * W + xc1;
        const bool k00 = fin && ya && xa, k01 = fin && ya && xb, k10 = fin && yb && xa, k11 = fin && yb && xb;
        const float w00 = k00 ? wnw : 0.f, w01 = k01 ? wne : 0.f, w10 = k10 ? wsw : 0.f, w11 = k11 ? wse : 0.f;
        float s = 0.f;
        const float* g1p = g1 + (size_t)b * C * N + r;
        const float* g2p = g2 + (size_t)b * C * N;
        // per-corner plane-0 addresses; the channel loop only adds c * N
        const float *q00 = g2p + o00, *q01 = g2p + o01, *q10 = g2p + o10, *q11 = g2p + o11;
        if (WINDOW) {
            const int4 wn = __ldg(reinterpret_cast<const int4*>(win) + b);
            const float* far = g2_far + (size_t)b * C * N;
            const bool ra = yc0 >= wn.x && yc0 < wn.y, rb = yc1 >= wn.x && yc1 < wn.y;
            const bool ca = xc0 >= wn.z && xc0 < wn.w, cb = xc1 >= wn.z && xc1 < wn.w;
            // a corner the reference would not sample is discarded by a select below: leave it on the device buffer
            if (k00 && !(ra && ca)) q00 = far + o00;
            if (k01 && !(ra && cb)) q01 = far + o01;
            if (k10 && !(rb && ca)) q10 = far + o10;
            if (k11 && !(rb && cb)) q11 = far + o11;
        }
        int c = 0;
        if (BATCH > 0) {
            constexpr int NB = BATCH > 0 ? BATCH : 1;
            for (; c + NB <= C; c += NB) {
                float t[NB][4], a[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const size_t co = (size_t)(c + k) * N;
                    if (WINDOW) {
                        t[k][0] = ldg_ordered(q00 + co); t[k][1] = ldg_ordered(q01 + co); t[k][2] = ldg_ordered(q10 + co); t[k][3] = ldg_ordered(q11 + co);
                    } else {                     // one plane pointer + four 32-bit offsets: the form the compiler schedules best (88 vs 96 us)
                        const float* pl = g2p + co;
                        t[k][0] = ldg_ordered(pl + o00); t[k][1] = ldg_ordered(pl + o01); t[k][2] = ldg_ordered(pl + o10); t[k][3] = ldg_ordered(pl + o11);
                    }
                    a[k] = ldg_ordered(g1p + co);
                }
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    s = corner_dot(s, a[k], k00 ? t[k][0] : 0.f, k01 ? t[k][1] : 0.f, k10 ? t[k][2] : 0.f, k11 ? t[k][3] : 0.f, w00, w01, w10, w11);
            }
        }
#pragma unroll 8
        for (; c < C; ++c) {
            const size_t co = (size_t)c * N;
            float t00, t01, t10, t11;
            if (WINDOW) {
                t00 = __ldg(q00 + co); t01 = __ldg(q01 + co); t10 = __ldg(q10 + co); t11 = __ldg(q11 + co);
            } else {
                const float* pl = g2p + co;
                t00 = __ldg(pl + o00); t01 = __ldg(pl + o01); t10 = __ldg(pl + o10); t11 = __ldg(pl + o11);
            }
            const float a = __ldg(g1p + co);
            s = corner_dot(s, a, k00 ? t00 : 0.f, k01 ? t01 : 0.f, k10 ? t10 : 0.f, k11 ? t11 : 0.f, w00, w01, w10, w11);
        }
        wgt = expf(-fabsf(1.f - s) / sigma);
    }
    weight[idx] = wgt;
}

template <int NPLANE, int BATCH, bool WINDOW = true>
__global__ void __launch_bounds__(64) upsample_weight_win_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int B, int C, int H, int W,
    float* __restrict__ flow_up, float* __restrict__ target, float* __restrict__ weight, int lazy_background,
    const float* __restrict__ g2_far, const int* __restrict__ win) {
    pdl_trigger();
    pdl_wait();
    const size_t N = NPLANE ? (size_t)NPLANE : (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * N) return;
    const int b = (int)(idx / N);
    const int r = (int)(idx - (size_t)b * N);
    const int Y = r / W, X = r - Y * W;
    // Fused-loop shortcut: a background pixel (syn_depth <= 0) has weight exactly 0, so the LM step ignores its target
    // (any finite value contributes 0 * finite = 0, as in the reference).  When the up-sampled flow itself is not an
    // output of this iteration, skip the mask softmax and the descriptor warp for it.
    if (lazy_background && !flow_up && depth[idx] <= 0.f) {
        if (target) *reinterpret_cast<float2*>(target + idx * 2) = make_float2((float)X, (float)Y);
        if (weight) weight[idx] = 0.f;
        return;
    }
    upsample_weight_pixel_win<NPLANE, BATCH, WINDOW>(flow, mask, g1, g2, depth, sigma, b, Y, X, C, H, W, flow_up, target, weight, g2_far, win);
}

// ------------------------------------------------------------------------------------------------ foreground list
// The rendered depth does not change over the recurrent iterations of a call, so the set of pixels that can carry a
// non-zero weight (depth > 0; non-finite depths are kept so that they poison the LM sums exactly as before) is compacted
// ONCE into a per-sample index list, in raster order (deterministic).  The LM kernel then runs over the list only (every
// unlisted pixel has weight exactly 0 and finite inputs: upsample_weight_kernel's background shortcut).  A list-driven
// variant of upsample_weight_kernel was measured too: no faster than the dense kernel (173 us both; profiles/r1c_summary.md).
//   row_count [B][H]  foreground pixels per image row      (fg_rows_kernel)
//   row_start [B][H]  exclusive prefix over the rows, fg_count[B] the total   (fg_scan_kernel)
//   fg_idx    [B][N]  pixel index Y*W+X of the k-th foreground pixel   (fg_fill_kernel)
__device__ __forceinline__ bool is_fg(float d) { return !(d <= 0.f); }

__global__ void __launch_bounds__(128) fg_rows_kernel(const float* __restrict__ depth, int H, int W, int* __restrict__ row_count) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;                    // b * H + Y
    const float* d = depth + (size_t)row * W;
    int c = 0;
    for (int X = threadIdx.x; X < W; X += blockDim.x) c += is_fg(d[X]) ? 1 : 0;
    __shared__ int red[4];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) row_count[row] = red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(256) fg_scan_kernel(const int* __restrict__ row_count, int H, int* __restrict__ row_start,
                                                      int* __restrict__ fg_count) {
    pdl_trigger();
    pdl_wait();
    // one block per sample, serial over chunks of 256 rows (H is a few hundred)
    const int b = blockIdx.x;
    __shared__ int sh[256];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < H; base += 256) {
        const int Y = base + threadIdx.x;
        const int v = Y < H ? row_count[b * H + Y] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                      // Hillis-Steele inclusive scan
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (Y < H) row_start[b * H + Y] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 255) carry += sh[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) fg_count[b] = carry;
}

__global__ void __launch_bounds__(128) fg_fill_kernel(const float* __restrict__ depth, int H, int W, const int* __restrict__ row_start,
                                                      int* __restrict__ fg_idx, float* __restrict__ target,
                                                      float* __restrict__ weight) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;                    // b * H + Y
    const int b = row / H, Y = row - b * H;
    const size_t N = (size_t)H * W;
    const float* d = depth + (size_t)row * W;
    __shared__ int wsum[4];
    __shared__ int base;
    if (threadIdx.x == 0) base = row_start[row];
    __syncthreads();
    for (int X0 = 0; X0 < W; X0 += blockDim.x) {
        const int X = X0 + threadIdx.x;
        const bool fg = X < W && is_fg(d[X]);
        const unsigned bal = __ballot_sync(0xffffffffu, fg);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int k = 0; k < wid; ++k) off += wsum[k];
        const int rank = off + __popc(bal & ((1u << lane) - 1u));
        if (fg) {
            fg_idx[(size_t)b * N + rank] = Y * W + X;
        } else if (X < W && weight) {                   // background: weight 0 and a finite target, once per call
            const size_t idx = (size_t)b * N + (size_t)Y * W + X;
            *reinterpret_cast<float2*>(target + idx * 2) = make_float2((float)X, (float)Y);
            weight[idx] = 0.f;
        }
        __syncthreads();
        if (threadIdx.x == 0) base += wsum[0] + wsum[1] + wsum[2] + wsum[3];
        __syncthreads();
    }
}

// upsample + weight over the foreground list, PERSISTENT: gridDim.x blocks per sample stride over that sample's list.
// The dense kernel launches 38 400 two-warp blocks of which 60 % exit at once; ncu shows it latency-bound at 27 % of
// HBM peak with half the warp slots idle, and a list-driven kernel with the same block count ran no faster: the block
// launch rate (~260 blocks per SM) bounds both.  Here every resident warp works on foreground pixels until the list ends.
template <int BLOCKS_PER_SM>
__global__ void __launch_bounds__(128, BLOCKS_PER_SM) upsample_weight_fg_kernel(
    const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ g1,
    const float* __restrict__ g2, const float* __restrict__ depth, float sigma, int C, int H, int W,
    const int* __restrict__ fg_idx, const int* __restrict__ fg_count, float* __restrict__ target, float* __restrict__ weight) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    const int count = fg_count[b];
    const int* idx = fg_idx + (size_t)b * H * W;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const int r = __ldg(idx + k);
        const int Y = r / W, X = r - Y * W;
        upsample_weight_pixel(flow, mask, g1, g2, depth, sigma, b, Y, X, C, H, W, nullptr, target, weight);
    }
}

// Host-entry helper: upload of the first descriptor map.  upsample_weight_pixel reads geofea1 only where depth > 0, so when
// the caller's buffer is pinned (device-accessible) only those pixels are fetched over PCIe: src is the mapped HOST pointer,
// depth / dst are device memory.  Grid (pixel chunks, channel groups of 8, samples); 8 independent loads per thread.
__global__ void __launch_bounds__(256) gather_fg_planes_kernel(const float* __restrict__ depth, const float* __restrict__ src,
                                                               float* __restrict__ dst, int C, int N) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z, c0 = blockIdx.y * 8;
    if (r >= N || !(depth[(size_t)b * N + r] > 0.f)) return;
    const size_t base = ((size_t)b * C + c0) * N + r;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = c0 + k < C ? src[base + (size_t)k * N] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (c0 + k < C) dst[base + (size_t)k * N] = v[k];
}

}  // namespace

int b2p_upsample_weight(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth,
                        float sigma, int B, int C, int H, int W, float* flow_up, float* target, float* weight,
                        int lazy_background, cudaStream_t s, const float* g2_far, const int* g2_window) {
    const size_t total = (size_t)B * H * W;
    const dim3 grid((unsigned)((total + 63) / 64));
    // upsample_variant (A/B): 0 run-time plane size, compiler-scheduled loads; 1 / 2 / 3 the reference's crop size (ZOOM_CROP_SIZE
    // 240x320, config/default.py) as a compile-time plane stride (immediate load offsets) with the plain loop / 8 / 16 channels
    // staged (5: 32); 4 / 6 run-time size, 8 / 16 staged.  Measured (profiles/r2l): 97 / 124 / 120 / 88 / 97 us -> default 3;
    // other crop sizes run variant 0
    const int var = b2p_options().upsample_variant;
    const bool cs = H * W == 240 * 320;
    b2p_pdl_next_allowed() = b2p_pdl_allowed(4);
#define UPW_LAUNCH(NP, BT) B2P_CUDA(b2p_launch_pdl(upsample_weight_kernel<NP, BT>, grid, dim3(64), 0, s, flow, mask, g1, g2, depth, sigma, B, C, H, W, flow_up, target, weight, lazy_background))
#define UPW_LAUNCH_WIN(NP, BT) B2P_CUDA(b2p_launch_pdl(upsample_weight_win_kernel<NP, BT>, grid, dim3(64), 0, s, flow, mask, g1, g2, depth, sigma, B, C, H, W, flow_up, target, weight, lazy_background, g2_far, g2_window))
    if (g2_far && g2_window && g2) {            // host entry: geofea2 is only valid inside a per-sample window
        if (var != 0 && cs) UPW_LAUNCH_WIN(240 * 320, 16);
        else UPW_LAUNCH_WIN(0, 0);
    } else if (var == 1 && cs) UPW_LAUNCH(240 * 320, 0);
    else if (var == 2 && cs) UPW_LAUNCH(240 * 320, 8);
    else if (var == 3 && cs) UPW_LAUNCH(240 * 320, 16);
    else if (var == 5 && cs) UPW_LAUNCH(240 * 320, 32);
    else if (var == 6) UPW_LAUNCH(0, 16);
    else if (var == 4) UPW_LAUNCH(0, 8);
    else UPW_LAUNCH(0, 0);
#undef UPW_LAUNCH
#undef UPW_LAUNCH_WIN
    B2P_LAUNCH_CHECK();
    return 0;
}

size_t b2p_fg_ws_bytes(int B, int H, int W) {
    return align_up((size_t)B * H * W * sizeof(int), 256) + 2 * align_up((size_t)B * H * sizeof(int), 256) + align_up((size_t)B * sizeof(int), 256);
}

static inline void fg_ws_split(void* ws, int B, int H, int W, int** fg_idx, int** row_count, int** row_start, int** fg_count) {
    char* p = reinterpret_cast<char*>(ws);
    *fg_idx = reinterpret_cast<int*>(p); p += align_up((size_t)B * H * W * sizeof(int), 256);
    *row_count = reinterpret_cast<int*>(p); p += align_up((size_t)B * H * sizeof(int), 256);
    *row_start = reinterpret_cast<int*>(p); p += align_up((size_t)B * H * sizeof(int), 256);
    *fg_count = reinterpret_cast<int*>(p);
}

const int* b2p_fg_idx(const void* ws) { return reinterpret_cast<const int*>(ws); }
const int* b2p_fg_count(const void* ws, int B, int H, int W) {
    return reinterpret_cast<const int*>(reinterpret_cast<const char*>(ws) + align_up((size_t)B * H * W * sizeof(int), 256) +
                                        2 * align_up((size_t)B * H * sizeof(int), 256));
}

// once per call: the foreground list of `depth`; target / weight (optional): their background pixels are set here
// (finite target, weight 0) for the list-driven upsample + weight kernel, which never touches them
int b2p_fg_build(const float* depth, int B, int H, int W, void* ws, float* target, float* weight, cudaStream_t s) {
    int *fg_idx, *row_count, *row_start, *fg_count;
    fg_ws_split(ws, B, H, W, &fg_idx, &row_count, &row_start, &fg_count);
    B2P_CUDA(b2p_launch_pdl(fg_rows_kernel, dim3(B * H), dim3(128), 0, s, depth, H, W, row_count));
    B2P_LAUNCH_CHECK();
    B2P_CUDA(b2p_launch_pdl(fg_scan_kernel, dim3(B), dim3(256), 0, s, (const int*)row_count, H, row_start, fg_count));
    B2P_LAUNCH_CHECK();
    B2P_CUDA(b2p_launch_pdl(fg_fill_kernel, dim3(B * H), dim3(128), 0, s, depth, H, W, (const int*)row_start, fg_idx, target, weight));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_upsample_weight_fg(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth, float sigma,
                           int B, int C, int H, int W, const void* fg_ws, float* target, float* weight, cudaStream_t s) {
    int dev = 0, sms = 0;
    B2P_CUDA(cudaGetDevice(&dev));
    B2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // all blocks resident; B200POSE_FG_BLOCKS picks the register / occupancy point (8 blocks of 128 threads per SM at 64
    // registers, or 6 at 80)
    const int bps = b2p_options().fg_blocks == 6 ? 6 : 8;
    int per_sample = (sms * bps) / B;
    const int useful = ceil_div(H * W, 128);
    if (per_sample > useful) per_sample = useful;
    if (per_sample < 1) per_sample = 1;
    const dim3 grid((unsigned)per_sample, (unsigned)B);
    if (bps == 6)
        B2P_CUDA(b2p_launch_pdl(upsample_weight_fg_kernel<6>, grid, dim3(128), 0, s, flow, mask, g1, g2, depth, sigma, C, H, W,
                                b2p_fg_idx(fg_ws), b2p_fg_count(fg_ws, B, H, W), target, weight));
    else
        B2P_CUDA(b2p_launch_pdl(upsample_weight_fg_kernel<8>, grid, dim3(128), 0, s, flow, mask, g1, g2, depth, sigma, C, H, W,
                                b2p_fg_idx(fg_ws), b2p_fg_count(fg_ws, B, H, W), target, weight));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_gather_fg_planes(const float* depth_dev, const float* src_mapped, float* dst_dev, int B, int C, int H, int W, cudaStream_t s) {
    const int N = H * W;
    gather_fg_planes_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)ceil_div(C, 8), (unsigned)B), 256, 0, s>>>(depth_dev, src_mapped, dst_dev, C, N);
    B2P_LAUNCH_CHECK();
    return 0;
}
