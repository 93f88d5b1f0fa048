// Correlation pyramid pooling, 9x9x4 bilinear lookup, context init and reprojection flow-init.
//   reference thirdparty/raft/corr.py:28-57, thirdparty/raft/utils/utils.py:57-71,
//   model/CFNet.py:124-144, model/PoseRefiner.py:324-328, geometry/transformation.py:184-198,
//   geometry/projective_ops.py:68-114.
#include "common.cuh"

namespace {

// 2x2 average pooling with floor (odd trailing row/col dropped), torch order ((a+b)+c)+d then /4.
__global__ void corr_pool_kernel(const float* __restrict__ src, int hs, int ws, int hd, int wd,
                                 float* __restrict__ dst, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % wd);
    size_t t = i / wd;
    const int y = (int)(t % hd);
    const size_t n = t / hd;
    const float* s = src + n * (size_t)hs * ws + (size_t)(2 * y) * ws + 2 * x;
    float sum = s[0];
    sum += s[1];
    sum += s[ws];
    sum += s[ws + 1];
    dst[i] = sum * 0.25f;
}

// NCHW feature map [B][D][P] -> PXC fp16 hi/lo planes [B*P][D] (operands of the tensor-core correlation GEMM).
__global__ void __launch_bounds__(256) fmap_to_pxc_half_kernel(const float* __restrict__ f, int D, int P,
                                                               __half* __restrict__ hi, __half* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int dd = ty; dd < 32; dd += 8) {
        const int p = p0 + tx, d = d0 + dd;
        tile[dd][tx] = (p < P && d < D) ? __ldg(f + ((size_t)b * D + d) * P + p) : 0.f;
    }
    __syncthreads();
    for (int pp = ty; pp < 32; pp += 8) {
        const int p = p0 + pp, d = d0 + tx;
        if (p < P && d < D) {
            const size_t o = ((size_t)b * P + p) * D + d;
            b2p_split_half(tile[tx][pp], hi[o], lo[o]);
        }
    }
}

// One warp per low-resolution pixel; each lane produces ~10 of the 324 samples.
// Output channel l*81 + i*9 + j samples (cx/2^l + i - 4, cy/2^l + j - 4): the slow window index moves x
// (reference corr.py:44-50 stacks meshgrid(dy,dx) into the (x,y) slots).
__global__ void __launch_bounds__(256) corr_lookup_kernel(const float* __restrict__ pyr, const float* __restrict__ coords,
                                                          int B, int h, int w, float* __restrict__ out,
                                                          __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    pdl_trigger();
    pdl_wait();
    const int P = h * w;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B * P) return;
    const int b = warp / P, p = warp - b * P;
    const float cx = coords[(size_t)warp * 2 + 0];
    const float cy = coords[(size_t)warp * 2 + 1];
    float* o = out ? out + (size_t)warp * B200POSE_CORR_PITCH : nullptr;
    __half* oh = out_hi ? out_hi + (size_t)warp * B200POSE_CORR_PITCH : nullptr;
    __half* ol = out_hi ? out_lo + (size_t)warp * B200POSE_CORR_PITCH : nullptr;
    size_t lvl_off = 0;
    int hl = h, wl = w;
    float inv = 1.0f;
#pragma unroll 1
    for (int l = 0; l < B200POSE_CORR_LEVELS; ++l) {
        const float* img = pyr + lvl_off + ((size_t)b * P + p) * (size_t)(hl * wl);
        const float x0c = cx * inv, y0c = cy * inv;      // cx / 2^l (exact: power of two)
        for (int k = lane; k < 81; k += 32) {
            const int i = k / 9, j = k - i * 9;
            const float xs = x0c + (float)(i - 4);
            const float ys = y0c + (float)(j - 4);
            const float xf = floorf(xs), yf = floorf(ys);
            const float fx = xs - xf, fy = ys - yf;
            const int xi = (int)xf, yi = (int)yf;
            float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
            const bool x0ok = xi >= 0 && xi < wl, x1ok = xi + 1 >= 0 && xi + 1 < wl;
            const bool y0ok = yi >= 0 && yi < hl, y1ok = yi + 1 >= 0 && yi + 1 < hl;
            if (y0ok && x0ok) v00 = __ldg(img + yi * wl + xi);
            if (y0ok && x1ok) v01 = __ldg(img + yi * wl + xi + 1);
            if (y1ok && x0ok) v10 = __ldg(img + (yi + 1) * wl + xi);
            if (y1ok && x1ok) v11 = __ldg(img + (yi + 1) * wl + xi + 1);
            const float val = v00 * (1.f - fx) * (1.f - fy) + v01 * fx * (1.f - fy) + v10 * (1.f - fx) * fy + v11 * fx * fy;
            if (o) o[l * 81 + k] = val;
            if (oh) b2p_split_half(val, oh[l * 81 + k], ol[l * 81 + k]);
        }
        lvl_off += (size_t)B * P * (size_t)(hl * wl);
        hl >>= 1; wl >>= 1; inv *= 0.5f;
    }
    if (lane < B200POSE_CORR_PITCH - B200POSE_CORR_CH) {
        if (o) o[B200POSE_CORR_CH + lane] = 0.f;
        if (oh) { oh[B200POSE_CORR_CH + lane] = __float2half(0.f); ol[B200POSE_CORR_CH + lane] = __float2half(0.f); }
    }
}

// ------------------------------------------------------------------------------------------------ lookup, second version
// The 81 samples of one level share ONE fractional offset (the window offsets are integers), so they are the four-corner
// blend of a 10 x 10 texel window with constant weights.  One warp per low-resolution pixel:
//   1. the four windows (4 x 100 texels) are staged into shared memory with branch-free loads (clamped address + select:
//      zeros outside the level, utils/utils.py:57-65 grid_sample zero padding), 400 loads instead of 1296;
//   2. every lane blends its samples from shared memory (same products and order as the first kernel) into a 328-wide
//      staging row; 3. the row leaves as 128-bit stores: fp32 (float4) and/or the fp16 hi / lo operand planes (8 halves).
// Channel order: l*81 + i*9 + j samples (x + i - 4, y + j - 4): the slow window index moves x (corr.py:44-50).
constexpr int LK_WARPS = 8;
constexpr int LK_WIN = 104;                     // 100 texels per level window, padded

__global__ void __launch_bounds__(LK_WARPS * 32) corr_lookup_win_kernel(const float* __restrict__ pyr, const float* __restrict__ coords,
                                                                         int B, int h, int w, float* __restrict__ out,
                                                                         __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    __shared__ __align__(16) float win[LK_WARPS][B200POSE_CORR_LEVELS][LK_WIN];
    __shared__ __align__(16) float row[LK_WARPS][B200POSE_CORR_PITCH];
    pdl_trigger();
    pdl_wait();
    const int P = h * w;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * LK_WARPS + wib;
    if (warp >= B * P) return;                                   // warp-uniform; no block-wide barrier below
    const int b = warp / P, p = warp - b * P;
    const float cx = coords[(size_t)warp * 2 + 0];
    const float cy = coords[(size_t)warp * 2 + 1];
    float fxs[B200POSE_CORR_LEVELS], fys[B200POSE_CORR_LEVELS];
    size_t lvl_off = 0;
    int hl = h, wl = w;
    float inv = 1.0f;
#pragma unroll
    for (int l = 0; l < B200POSE_CORR_LEVELS; ++l) {
        const float* img = pyr + lvl_off + ((size_t)b * P + p) * (size_t)(hl * wl);
        // sample (i, j) sits at (x0c + i - 4, y0c + j - 4); floor(x0c + k) == floor(x0c) + k exactly unless |x0c| is so
        // large that the window lies outside every level anyway (then the integer conversion saturates and all texels
        // fail the bounds test: zeros, like the reference)
        const float x0c = cx * inv, y0c = cy * inv;              // cx / 2^l (exact: power of two)
        const float xb = floorf(x0c - 4.0f), yb = floorf(y0c - 4.0f);
        const bool fin = isfinite(x0c) && isfinite(y0c);
        fxs[l] = fin ? (x0c - 4.0f) - xb : 0.f; fys[l] = fin ? (y0c - 4.0f) - yb : 0.f;     // non-finite centre: exact zeros
        const int xi0 = fin ? (int)fmaxf(fminf(xb, 1e6f), -1e6f) : -1000000;
        const int yi0 = fin ? (int)fmaxf(fminf(yb, 1e6f), -1e6f) : -1000000;
#pragma unroll
        for (int t = lane; t < 100; t += 32) {
            const int ry = t / 10, rx = t - ry * 10;
            const int yy = yi0 + ry, xx = xi0 + rx;
            const bool ok = yy >= 0 && yy < hl && xx >= 0 && xx < wl;
            const int yc = min(max(yy, 0), hl - 1), xc = min(max(xx, 0), wl - 1);
            const float v = __ldg(img + yc * wl + xc);
            win[wib][l][t] = ok ? v : 0.f;
        }
        lvl_off += (size_t)B * P * (size_t)(hl * wl);
        hl >>= 1; wl >>= 1; inv *= 0.5f;
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < B200POSE_CORR_LEVELS; ++l) {
        const float fx = fxs[l], fy = fys[l];
        // the first kernel evaluates floor() per sample: xs = x0c + (i - 4), fx = xs - floor(xs).  For |x0c| < 2^22 both
        // give the same fraction up to the rounding of the addition; keep its weights' form (1-fx)(1-fy) etc.
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        for (int k = lane; k < 81; k += 32) {
            const int i = k / 9, j = k - i * 9;                  // i moves x, j moves y
            const float* wp = &win[wib][l][j * 10 + i];
            row[wib][l * 81 + k] = wp[0] * w00 + wp[1] * w01 + wp[10] * w10 + wp[11] * w11;
        }
    }
    if (lane < B200POSE_CORR_PITCH - B200POSE_CORR_CH) row[wib][B200POSE_CORR_CH + lane] = 0.f;
    __syncwarp();
    if (out) {
        float4* o4 = reinterpret_cast<float4*>(out + (size_t)warp * B200POSE_CORR_PITCH);
        const float4* r4 = reinterpret_cast<const float4*>(row[wib]);
        for (int c = lane; c < B200POSE_CORR_PITCH / 4; c += 32) o4[c] = r4[c];
    }
    if (out_hi) {
        uint4* oh = reinterpret_cast<uint4*>(out_hi + (size_t)warp * B200POSE_CORR_PITCH);
        uint4* ol = reinterpret_cast<uint4*>(out_lo + (size_t)warp * B200POSE_CORR_PITCH);
        for (int c = lane; c < B200POSE_CORR_PITCH / 8; c += 32) {
            uint32_t wh[4], wlw[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                __half h0, l0, h1, l1;
                b2p_split_half(row[wib][c * 8 + 2 * t], h0, l0);
                b2p_split_half(row[wib][c * 8 + 2 * t + 1], h1, l1);
                wh[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                wlw[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            oh[c] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            ol[c] = make_uint4(wlw[0], wlw[1], wlw[2], wlw[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ pyramid levels 1..3 in one pass
// One block per (sample, source pixel) image of level 0 (hs x ws floats): the image is read once into shared memory and the
// three 2x2 floor poolings are chained there (corr.py:32-34; the first kernel reads level l to write level l+1: 3 passes).
// Summation order as corr_pool_kernel: ((a + b) + c) + d, then * 0.25.
__global__ void __launch_bounds__(128) corr_pool3_kernel(const float* __restrict__ l0, int hs, int ws, float* __restrict__ l1,
                                                         float* __restrict__ l2, float* __restrict__ l3) {
    extern __shared__ float sm[];
    const size_t n = blockIdx.x;
    const int n0 = hs * ws, h1 = hs >> 1, w1 = ws >> 1, h2 = h1 >> 1, w2 = w1 >> 1, h3 = h2 >> 1, w3 = w2 >> 1;
    float* s0 = sm; float* s1 = s0 + n0; float* s2 = s1 + h1 * w1;
    const float* src = l0 + n * n0;
    if ((n0 & 3) == 0) {
        for (int i = threadIdx.x; i < n0 / 4; i += blockDim.x) reinterpret_cast<float4*>(s0)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int i = threadIdx.x; i < n0; i += blockDim.x) s0[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h1 * w1; i += blockDim.x) {
        const int y = i / w1, x = i - y * w1;
        const float* s = s0 + (2 * y) * ws + 2 * x;
        float sum = s[0]; sum += s[1]; sum += s[ws]; sum += s[ws + 1];
        const float v = sum * 0.25f;
        s1[i] = v; l1[n * (size_t)(h1 * w1) + i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h2 * w2; i += blockDim.x) {
        const int y = i / w2, x = i - y * w2;
        const float* s = s1 + (2 * y) * w1 + 2 * x;
        float sum = s[0]; sum += s[1]; sum += s[w1]; sum += s[w1 + 1];
        const float v = sum * 0.25f;
        s2[i] = v; l2[n * (size_t)(h2 * w2) + i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h3 * w3; i += blockDim.x) {
        const int y = i / w3, x = i - y * w3;
        const float* s = s2 + (2 * y) * w2 + 2 * x;
        float sum = s[0]; sum += s[1]; sum += s[w2]; sum += s[w2 + 1];
        l3[n * (size_t)(h3 * w3) + i] = sum * 0.25f;
    }
}

// context [B,256,H,W] --(1/8 bilinear, align_corners=True)--> net = tanh(ch 0..127) [P][128],
// xbuf[:, 0:128] = relu(ch 128..255).   32 low-res pixels x 32 channels per block, smem transpose.
// PACKED: ctx holds only the four texels of each low-res sample, [B,256,P] float4 = (v00, v01, v10, v11), as gathered on the
// host by b200pose_refine_iters_host2 (api.cu: gather_context_texels) -- same arithmetic, 1/16 of the bytes.
// MIXED (the host entry's split between host-thread gather and in-place reads): channels below c_split come from `tex`
// ([B][c_split][P] float4), the others from the full map `ctx`; a block handles one 32-channel tile, so the choice is block-uniform.
template <bool PACKED>
__global__ void __launch_bounds__(256) context_init_kernel(const float* __restrict__ ctx, const float* __restrict__ tex, int c_split,
                                                           int B, int H, int W, int h, int w,
                                                           float sy, float sx, float* __restrict__ net,
                                                           float* __restrict__ xbuf, __half* __restrict__ net_hi,
                                                           __half* __restrict__ net_lo, __half* __restrict__ x_hi,
                                                           __half* __restrict__ x_lo) {
    __shared__ float tile[32][33];
    const int P = h * w;
    const int p0 = blockIdx.x * 32;      // pixel tile within sample
    const int c0 = blockIdx.y * 32;      // channel tile
    const int b = blockIdx.z;
    const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;   // 8 rows of 32
    // read phase: tx -> pixel, rows -> channels
    const int p = p0 + tx;
    if (p < P) {
        const int y = p / w, x = p - y * w;
        const float fy = sy * (float)y, fxx = sx * (float)x;
        int y0 = (int)fy, x0 = (int)fxx;
        y0 = min(y0, H - 1); x0 = min(x0, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = fy - (float)y0, lx = fxx - (float)x0;
        for (int cc = tyy; cc < 32; cc += 8) {
            float v00, v01, v10, v11;
            if (PACKED && c0 < c_split) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(tex) + ((size_t)b * c_split + c0 + cc) * (size_t)P + p);
                v00 = t.x; v01 = t.y; v10 = t.z; v11 = t.w;
            } else {
                const float* pl = ctx + ((size_t)b * 256 + c0 + cc) * (size_t)H * W;
                v00 = __ldg(pl + (size_t)y0 * W + x0); v01 = __ldg(pl + (size_t)y0 * W + x1);
                v10 = __ldg(pl + (size_t)y1 * W + x0); v11 = __ldg(pl + (size_t)y1 * W + x1);
            }
            const float top = v00 * (1.f - lx) + v01 * lx;
            const float bot = v10 * (1.f - lx) + v11 * lx;
            tile[cc][tx] = top * (1.f - ly) + bot * ly;
        }
    }
    __syncthreads();
    // write phase: tx -> channel, rows -> pixels
    for (int pp = tyy; pp < 32; pp += 8) {
        const int pw = p0 + pp;
        if (pw >= P) continue;
        const float v = tile[tx][pp];
        const int c = c0 + tx;
        const size_t pix = (size_t)b * P + pw;
        if (c < 128) {
            const float t = tanhf(v);
            net[pix * 128 + c] = t;
            if (net_hi) b2p_split_half(t, net_hi[pix * 128 + c], net_lo[pix * 128 + c]);
        } else {
            const float t = fmaxf(v, 0.f);
            if (xbuf) xbuf[pix * 256 + (c - 128)] = t;
            if (x_hi) b2p_split_half(t, x_hi[pix * 256 + (c - 128)], x_lo[pix * 256 + (c - 128)]);
        }
    }
}

// Reprojection of one full-resolution pixel: flow_init / 8 (PoseRefiner.py:324-328, CFNet.py:140).
__device__ __forceinline__ float2 reproj_flow8(const float* __restrict__ depth, int W, int u, int v, float fx, float fy,
                                               float cx, float cy, const float* G) {
    const float Z = depth[(size_t)v * W + u] + 1e-5f;
    const float X = Z * ((float)u - cx) / fx;
    const float Y = Z * ((float)v - cy) / fy;
    const float X1 = G[0] * X + G[1] * Y + G[2] * Z + G[3];
    const float Y1 = G[4] * X + G[5] * Y + G[6] * Z + G[7];
    const float Z1 = G[8] * X + G[9] * Y + G[10] * Z + G[11];
    const float Zc = fmaxf(Z1, 0.01f);
    const float x1 = fx * (X1 / Zc) + cx;
    const float y1 = fy * (Y1 / Zc) + cy;
    const float msk = (Z > 1e-5f) ? 1.f : 0.f;
    return make_float2(((x1 - (float)u) * msk) / 8.f, ((y1 - (float)v) * msk) / 8.f);
}

__global__ void __launch_bounds__(256) flow_init_kernel(const float* __restrict__ depth, const float* __restrict__ K,
                                                        const float* __restrict__ G, int B, int H, int W, int h, int w,
                                                        float sy, float sx, float* __restrict__ coords1,
                                                        float* __restrict__ flow) {
    pdl_trigger();
    pdl_wait();
    const int P = h * w;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * P) return;
    const int b = idx / P, p = idx - b * P;
    const int y = p / w, x = p - y * w;
    const float* Kb = K + b * 9;
    const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
    float Gm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Gm[i] = G[b * 16 + i];
    const float* d = depth + (size_t)b * H * W;
    const float fyy = sy * (float)y, fxx = sx * (float)x;
    int y0 = min((int)fyy, H - 1), x0 = min((int)fxx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = fyy - (float)y0, lx = fxx - (float)x0;
    const float2 f00 = reproj_flow8(d, W, x0, y0, fx, fy, cx, cy, Gm);
    const float2 f01 = reproj_flow8(d, W, x1, y0, fx, fy, cx, cy, Gm);
    const float2 f10 = reproj_flow8(d, W, x0, y1, fx, fy, cx, cy, Gm);
    const float2 f11 = reproj_flow8(d, W, x1, y1, fx, fy, cx, cy, Gm);
    const float flx = (f00.x * (1.f - lx) + f01.x * lx) * (1.f - ly) + (f10.x * (1.f - lx) + f11.x * lx) * ly;
    const float fly = (f00.y * (1.f - lx) + f01.y * lx) * (1.f - ly) + (f10.y * (1.f - lx) + f11.y * lx) * ly;
    const float c1x = (float)x + flx, c1y = (float)y + fly;       // coords1 = coords0 + flow_init (CFNet.py:144)
    coords1[(size_t)idx * 2 + 0] = c1x;
    coords1[(size_t)idx * 2 + 1] = c1y;
    flow[(size_t)idx * 2 + 0] = c1x - (float)x;                   // flow = coords1 - coords0 (CFNet.py:151)
    flow[(size_t)idx * 2 + 1] = c1y - (float)y;
}

}  // namespace

int b2p_fmap_to_pxc_half(const float* f, int B, int D, int P, __half* hi, __half* lo, cudaStream_t s) {
    dim3 grid(ceil_div(P, 32), ceil_div(D, 32), B);
    fmap_to_pxc_half_kernel<<<grid, 256, 0, s>>>(f, D, P, hi, lo);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_corr_pool(const float* src, int NP, int hs, int ws, float* dst, cudaStream_t s) {
    const int hd = hs / 2, wd = ws / 2;
    const size_t total = (size_t)NP * hd * wd;
    if (total == 0) return 0;
    corr_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, hs, ws, hd, wd, dst, total);
    B2P_LAUNCH_CHECK();
    return 0;
}

// lookup_mode 2: one thread per (pixel, level, window row j).  The thread loads the two texel rows j and j + 1 of its level's 10x10
// window straight from the pyramid (2 x 10 contiguous floats: clamped addresses + selects, zeros outside the level like the
// reference's grid_sample) and blends the 9 samples of its row from registers -- 20 loads per 9 outputs instead of the window
// kernel's 100 staged texels + 4 shared-memory reads per output; same weights, same blend order.  The 328-wide row of a pixel is
// assembled in shared memory and leaves as 128-bit stores as before.  36 threads per pixel, LR_PIX pixels per block.
template <int LR_PIX>
__global__ void __launch_bounds__(LR_PIX * 36) corr_lookup_row_kernel(const float* __restrict__ pyr, const float* __restrict__ coords,
                                                                       int B, int h, int w, float* __restrict__ out,
                                                                       __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    __shared__ __align__(16) float row[LR_PIX][B200POSE_CORR_PITCH];
    pdl_trigger();
    pdl_wait();
    const int P = h * w;
    const int total = B * P;
    const int pl = threadIdx.x / 36, r = threadIdx.x - pl * 36;
    const int l = r / 9, j = r - l * 9;
    const int pix = blockIdx.x * LR_PIX + pl;
    if (pix < total) {
        const int b = pix / P, p = pix - b * P;
        size_t lvl_off = 0;
        int hl = h, wl = w;
        float inv = 1.f;
        for (int k = 0; k < l; ++k) { lvl_off += (size_t)B * P * (size_t)(hl * wl); hl >>= 1; wl >>= 1; inv *= 0.5f; }
        const float* img = pyr + lvl_off + ((size_t)b * P + p) * (size_t)(hl * wl);
        const float x0c = coords[(size_t)pix * 2 + 0] * inv, y0c = coords[(size_t)pix * 2 + 1] * inv;    // / 2^l, exact
        const float xb = floorf(x0c - 4.0f), yb = floorf(y0c - 4.0f);
        const bool fin = isfinite(x0c) && isfinite(y0c);
        const float fx = fin ? (x0c - 4.0f) - xb : 0.f, fy = fin ? (y0c - 4.0f) - yb : 0.f;          // non-finite centre: exact zeros
        const int xi0 = fin ? (int)fmaxf(fminf(xb, 1e6f), -1e6f) : -1000000;
        const int yi0 = fin ? (int)fmaxf(fminf(yb, 1e6f), -1e6f) : -1000000;
        const int ya = yi0 + j, yn = ya + 1;
        const bool oka = ya >= 0 && ya < hl, okn = yn >= 0 && yn < hl;
        const float* ra = img + min(max(ya, 0), hl - 1) * wl;
        const float* rn = img + min(max(yn, 0), hl - 1) * wl;
        float va[10], vn[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const int xx = xi0 + k;
            const bool okx = xx >= 0 && xx < wl;
            const int xc = min(max(xx, 0), wl - 1);
            const float a = __ldg(ra + xc), c = __ldg(rn + xc);
            va[k] = (okx && oka) ? a : 0.f;
            vn[k] = (okx && okn) ? c : 0.f;
        }
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
        float* dst = &row[pl][l * 81 + j];                        // channel = l*81 + i*9 + j: i moves x, j moves y
#pragma unroll
        for (int i = 0; i < 9; ++i)
            dst[i * 9] = __fmaf_rn(vn[i + 1], w11, __fmaf_rn(vn[i], w10, __fmaf_rn(va[i + 1], w01, __fmul_rn(va[i], w00))));
        if (r < B200POSE_CORR_PITCH - B200POSE_CORR_CH) row[pl][B200POSE_CORR_CH + r] = 0.f;
    }
    __syncthreads();
    const int npix = min(LR_PIX, total - blockIdx.x * LR_PIX);
    const size_t base = (size_t)blockIdx.x * LR_PIX * B200POSE_CORR_PITCH;
    const float* flat = &row[0][0];
    if (out) {
        float4* o4 = reinterpret_cast<float4*>(out + base);
        const float4* r4 = reinterpret_cast<const float4*>(flat);
        for (int c = threadIdx.x; c < npix * (B200POSE_CORR_PITCH / 4); c += LR_PIX * 36) o4[c] = r4[c];
    }
    if (out_hi) {
        uint4* oh = reinterpret_cast<uint4*>(out_hi + base);
        uint4* ol = reinterpret_cast<uint4*>(out_lo + base);
        for (int c = threadIdx.x; c < npix * (B200POSE_CORR_PITCH / 8); c += LR_PIX * 36) {
            uint32_t wh[4], wlw[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                __half h0, l0, h1, l1;
                b2p_split_half(flat[c * 8 + 2 * t], h0, l0);
                b2p_split_half(flat[c * 8 + 2 * t + 1], h1, l1);
                wh[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                wlw[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            oh[c] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
            ol[c] = make_uint4(wlw[0], wlw[1], wlw[2], wlw[3]);
        }
    }
}

// levels 1..3 from level 0 in one pass; -1 if the level-0 image does not fit the shared-memory budget (caller falls back)
int b2p_corr_pool3(float* pyramid, int B, int h, int w, cudaStream_t s) {
    const int P = h * w;
    const size_t n0 = (size_t)h * w, n1 = (size_t)(h >> 1) * (w >> 1), n2 = (size_t)(h >> 2) * (w >> 2), n3 = (size_t)(h >> 3) * (w >> 3);
    const size_t smem = (n0 + n1 + n2) * sizeof(float);
    if (smem > 40 * 1024 || b2p_options().pool_mode != 1) return -1;
    float* l0 = pyramid; float* l1 = l0 + (size_t)B * P * n0; float* l2 = l1 + (size_t)B * P * n1; float* l3 = l2 + (size_t)B * P * n2;
    (void)n3;
    corr_pool3_kernel<<<(unsigned)((size_t)B * P), 128, smem, s>>>(l0, h, w, l1, l2, l3);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_corr_lookup(const float* pyramid, const float* coords, int B, int h, int w, float* out, __half* out_hi,
                    __half* out_lo, cudaStream_t s) {
    const int warps = B * h * w;
    const int lm = b2p_options().lookup_mode;
    // 2 (default): the row kernel as a plain launch.  Launched with the programmatic-dependent-launch attribute (mode 3) the same
    // kernel makes the step 0.15 ms SLOWER although it is 17 us faster than the window kernel in isolation: its 4800 blocks become
    // resident behind the early triggers of the two tiny kernels in front of it and sit on the SMs while the LM kernel of the
    // previous iteration is still running (profiles/r3e: 3.535 ms with, 3.388 ms without the attribute; window kernel 3.455).
    if (lm == 2) {
        corr_lookup_row_kernel<8><<<ceil_div(warps, 8), 8 * 36, 0, s>>>(pyramid, coords, B, h, w, out, out_hi, out_lo);
        B2P_LAUNCH_CHECK();
        return 0;
    }
    if (lm == 3) {
        B2P_CUDA(b2p_launch_pdl(corr_lookup_row_kernel<8>, dim3(ceil_div(warps, 8)), dim3(8 * 36), 0, s, pyramid, coords, B, h, w,
                                out, out_hi, out_lo));
        B2P_LAUNCH_CHECK();
        return 0;
    }
    if (b2p_options().lookup_mode == 1) {
        B2P_CUDA(b2p_launch_pdl(corr_lookup_win_kernel, dim3(ceil_div(warps, LK_WARPS)), dim3(LK_WARPS * 32), 0, s, pyramid, coords, B, h, w,
                                out, out_hi, out_lo));
        B2P_LAUNCH_CHECK();
        return 0;
    }
    B2P_CUDA(b2p_launch_pdl(corr_lookup_kernel, dim3(ceil_div(warps, 8)), dim3(256), 0, s, pyramid, coords, B, h, w, out, out_hi, out_lo));
    B2P_LAUNCH_CHECK();
    return 0;
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int b2p_context_init(const float* ctx, int B, int H, int W, float* net, float* xbuf, __half* net_hi, __half* net_lo,
                     __half* x_hi, __half* x_lo, cudaStream_t s, const float* texels, int c_split) {
    const int h = H / 8, w = W / 8;
    dim3 grid(ceil_div(h * w, 32), 8, B);
    if (texels && c_split > 0)             // c_split (a multiple of 32) channels from the texel layout, the rest from the map
        context_init_kernel<true><<<grid, 256, 0, s>>>(ctx, texels, c_split, B, H, W, h, w, ac_scale(H, h), ac_scale(W, w), net, xbuf,
                                                         net_hi, net_lo, x_hi, x_lo);
    else
        context_init_kernel<false><<<grid, 256, 0, s>>>(ctx, nullptr, 0, B, H, W, h, w, ac_scale(H, h), ac_scale(W, w), net, xbuf,
                                                          net_hi, net_lo, x_hi, x_lo);
    B2P_LAUNCH_CHECK();
    return 0;
}

// The sample positions of the 1/8 resample (F.interpolate(align_corners=True), CFNet.py:129) exactly as context_init_kernel
// computes them; used by the host-side texel gather so that both sides pick the same texels.
void b2p_context_sample_taps(int in, int out, int* i0, int* i1) {
    const float sc = ac_scale(in, out);
    for (int o = 0; o < out; ++o) {
        const float f = sc * (float)o;
        int a = (int)f;
        a = a < in - 1 ? a : in - 1;
        i0[o] = a;
        i1[o] = a + 1 < in - 1 ? a + 1 : in - 1;
    }
}

int b2p_flow_init(const float* depth, const float* K, const float* G, int B, int H, int W, float* coords1, float* flow,
                  cudaStream_t s) {
    const int h = H / 8, w = W / 8;
    b2p_pdl_next_allowed() = b2p_pdl_allowed(0);
    B2P_CUDA(b2p_launch_pdl(flow_init_kernel, dim3(ceil_div(B * h * w, 256)), dim3(256), 0, s, depth, K, G, B, H, W, h, w, ac_scale(H, h),
                            ac_scale(W, w), coords1, flow));
    B2P_LAUNCH_CHECK();
    return 0;
}
