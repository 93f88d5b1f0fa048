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

// context [B,256,H,W] --(1/8 bilinear, align_corners=True)--> net = tanh(ch 0..127) [P][128],
// xbuf[:, 0:128] = relu(ch 128..255).   32 low-res pixels x 32 channels per block, smem transpose.
__global__ void __launch_bounds__(256) context_init_kernel(const float* __restrict__ ctx, int B, int H, int W, int h, int w,
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
            const float* pl = ctx + ((size_t)b * 256 + c0 + cc) * (size_t)H * W;
            const float v00 = __ldg(pl + (size_t)y0 * W + x0), v01 = __ldg(pl + (size_t)y0 * W + x1);
            const float v10 = __ldg(pl + (size_t)y1 * W + x0), v11 = __ldg(pl + (size_t)y1 * W + x1);
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

int b2p_corr_lookup(const float* pyramid, const float* coords, int B, int h, int w, float* out, __half* out_hi,
                    __half* out_lo, cudaStream_t s) {
    const int warps = B * h * w;
    B2P_CUDA(b2p_launch_pdl(corr_lookup_kernel, dim3(ceil_div(warps, 8)), dim3(256), 0, s, pyramid, coords, B, h, w, out, out_hi, out_lo));
    B2P_LAUNCH_CHECK();
    return 0;
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int b2p_context_init(const float* ctx, int B, int H, int W, float* net, float* xbuf, __half* net_hi, __half* net_lo,
                     __half* x_hi, __half* x_lo, cudaStream_t s) {
    const int h = H / 8, w = W / 8;
    dim3 grid(ceil_div(h * w, 32), 8, B);
    context_init_kernel<<<grid, 256, 0, s>>>(ctx, B, H, W, h, w, ac_scale(H, h), ac_scale(W, w), net, xbuf, net_hi,
                                               net_lo, x_hi, x_lo);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_flow_init(const float* depth, const float* K, const float* G, int B, int H, int W, float* coords1, float* flow,
                  cudaStream_t s) {
    const int h = H / 8, w = W / 8;
    B2P_CUDA(b2p_launch_pdl(flow_init_kernel, dim3(ceil_div(B * h * w, 256)), dim3(256), 0, s, depth, K, G, B, H, W, h, w, ac_scale(H, h),
                            ac_scale(W, w), coords1, flow));
    B2P_LAUNCH_CHECK();
    return 0;
}
