// Online zoom-crop on device: the step directly in front of the inner loop.
// Replaces, per render iteration, reference model/PoseRefiner.py:145-218 (get_affine_transformation: numpy nonzero + cv2 on the
// host, per sample; gen_zoom_crop_grids: projected model centre, F.affine_grid, inverse of the crop transform) and the two
// F.grid_sample calls of :287,:292 that crop the observed image and the dense 2-D descriptors:
//   1. zc_bbox_kernel     bounding box of the rendered foreground (pc_depth > 0) by integer atomicMin/Max (order-free)
//   2. zc_resample_kernel every block derives its sample's crop box (closed form of the axis-aligned affine map that the
//                         reference obtains from cv2.getAffineTransform), writes theta / K_crop once per sample, and
//                         resamples 64 output pixels x all channels: bilinear, zeros outside, align_corners=False for both
//                         affine_grid and grid_sample (torch defaults, as the reference).  The descriptors can leave
//                         CHANNELS-LAST [B][Hc*Wc][32] (through a shared-memory transpose), which is the layout the
//                         foreground pipeline of the loop wants (fg_pipeline.cu): the transposition costs nothing here.
// No host round trip: the reference synchronises on mask.cpu().numpy() per sample and render iteration.
#include "common.cuh"

namespace {

constexpr int ZC_PX = 64;          // output pixels per block
constexpr int ZC_BIG = 1 << 30;

__global__ void zc_init_kernel(int* __restrict__ bbox, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { bbox[i * 4 + 0] = ZC_BIG; bbox[i * 4 + 1] = -ZC_BIG; bbox[i * 4 + 2] = ZC_BIG; bbox[i * 4 + 3] = -ZC_BIG; }
}

// grid (row chunks, B): x_min, x_max, y_min, y_max of depth > 0 (np.nonzero of the mask, PoseRefiner.py:154-164)
__global__ void __launch_bounds__(256) zc_bbox_kernel(const float* __restrict__ depth, int H, int W, int* __restrict__ bbox) {
    const int b = blockIdx.y;
    const float* d = depth + (size_t)b * H * W;
    int xmin = ZC_BIG, xmax = -ZC_BIG, ymin = ZC_BIG, ymax = -ZC_BIG;
    for (int y = blockIdx.x; y < H; y += gridDim.x)
        for (int x = threadIdx.x; x < W; x += blockDim.x)
            if (d[(size_t)y * W + x] > 0.f) { xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0 && xmax >= 0) {
        atomicMin(&bbox[b * 4 + 0], xmin); atomicMax(&bbox[b * 4 + 1], xmax);
        atomicMin(&bbox[b * 4 + 2], ymin); atomicMax(&bbox[b * 4 + 3], ymax);
    }
}

struct CropBox { float x1, x2, y1, y2; };

// PoseRefiner.py:166-176 (crop box around the projected model centre) with :204-205 (centre = K T[:3,3]).
__device__ __forceinline__ CropBox crop_box(const int* bbox, const float* K, const float* T, int H, int W, float margin_ratio) {
    const float tx = T[3], ty = T[7], tz = T[11];
    const float cx = K[0] * tx + K[1] * ty + K[2] * tz, cy = K[3] * tx + K[4] * ty + K[5] * tz, cz = K[6] * tx + K[7] * ty + K[8] * tz;
    const float zx = cx / cz, zy = cy / cz;
    const bool has = bbox[1] >= 0;                       // empty mask: the reference uses a zero box (:160-164)
    const float x_min = has ? (float)bbox[0] : 0.f, x_max = has ? (float)bbox[1] : 0.f;
    const float y_min = has ? (float)bbox[2] : 0.f, y_max = has ? (float)bbox[3] : 0.f;
    const float ratio = (float)H / (float)W;
    const float left = zx - x_min, right = x_max - zx, up = zy - y_min, down = y_max - zy;
    const float crop_h = fmaxf(fmaxf(ratio * right, ratio * left), fmaxf(up, down)) * 2.f * (1.f + margin_ratio);
    const float crop_w = crop_h / ratio;
    CropBox c;
    c.x1 = zx - crop_w / 2.f; c.x2 = zx + crop_w / 2.f; c.y1 = zy - crop_h / 2.f; c.y2 = zy + crop_h / 2.f;
    return c;
}

// bilinear sample of one plane at pixel position (ix, iy), zeros outside (F.grid_sample default)
__device__ __forceinline__ float sample_plane(const float* __restrict__ pl, int H, int W, int x0, int y0, float wnw, float wne,
                                              float wsw, float wse, bool k00, bool k01, bool k10, bool k11, int o00, int o01, int o10,
                                              int o11) {
    const float t00 = __ldg(pl + o00), t01 = __ldg(pl + o01), t10 = __ldg(pl + o10), t11 = __ldg(pl + o11);
    float v = 0.f;
    v += (k00 ? t00 : 0.f) * wnw;
    v += (k01 ? t01 : 0.f) * wne;
    v += (k10 ? t10 : 0.f) * wsw;
    v += (k11 ? t11 : 0.f) * wse;
    return v;
}

// grid (ceil(Hc*Wc / 64), B), 256 threads.  Channels are processed 32 at a time (warp w: channels 4w..4w+3 of the group,
// lanes over the 64 pixels), staged in shared memory when the destination is channels-last.
__global__ void __launch_bounds__(256) zc_resample_kernel(const int* __restrict__ bbox, const float* __restrict__ K,
                                                          const float* __restrict__ T, const float* __restrict__ image,
                                                          const float* __restrict__ geo, int Ci, int Cg, int H, int W, int Hc, int Wc,
                                                          float margin_ratio, int geo_channels_last, float* __restrict__ image_crop,
                                                          float* __restrict__ geo_crop, float* __restrict__ K_crop,
                                                          float* __restrict__ theta_out) {
    __shared__ float tile[32][ZC_PX + 1];
    const int b = blockIdx.y, k0 = blockIdx.x * ZC_PX, Nc = Hc * Wc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CropBox cb = crop_box(bbox + b * 4, K + b * 9, T + b * 16, H, W, margin_ratio);
    // affine_grid theta: [-1,1]^2 of the output -> normalised input coordinates (the cv2.getAffineTransform of an axis-aligned
    // box is this diagonal map, PoseRefiner.py:178-186)
    const float nx1 = cb.x1 * 2.f / (float)W - 1.f, nx2 = cb.x2 * 2.f / (float)W - 1.f;
    const float ny1 = cb.y1 * 2.f / (float)H - 1.f, ny2 = cb.y2 * 2.f / (float)H - 1.f;
    const float t00 = (nx2 - nx1) / 2.f, t02 = (nx2 + nx1) / 2.f, t11 = (ny2 - ny1) / 2.f, t12 = (ny2 + ny1) / 2.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (theta_out) {
            float* th = theta_out + b * 6;
            th[0] = t00; th[1] = 0.f; th[2] = t02; th[3] = 0.f; th[4] = t11; th[5] = t12;
        }
        if (K_crop) {
            // crop pixel (0..Wc-1, 0..Hc-1) -> input pixel: A = [[sx,0,x1],[0,sy,y1],[0,0,1]]; K_crop = inv(A) K (:188-198, :211)
            const float sx = (cb.x2 - cb.x1) / (float)(Wc - 1), sy = (cb.y2 - cb.y1) / (float)(Hc - 1);
            const float* Kb = K + b * 9;
            float* o = K_crop + b * 9;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                o[j] = (Kb[j] - cb.x1 * Kb[6 + j]) / sx;
                o[3 + j] = (Kb[3 + j] - cb.y1 * Kb[6 + j]) / sy;
                o[6 + j] = Kb[6 + j];
            }
        }
    }
    // this thread's two output pixels: sampling position, corner offsets and weights (shared by every channel)
    int o00[2], o01[2], o10[2], o11[2];
    float wnw[2], wne[2], wsw[2], wse[2];
    bool k00[2], k01[2], k10[2], k11[2], live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k = k0 + lane + 32 * u;
        live[u] = k < Nc;
        const int kk = live[u] ? k : Nc - 1;
        const int yo = kk / Wc, xo = kk - yo * Wc;
        // F.affine_grid, align_corners=False: base coordinates (2 j + 1) / n - 1
        const float bx = (2.f * (float)xo + 1.f) / (float)Wc - 1.f, by = (2.f * (float)yo + 1.f) / (float)Hc - 1.f;
        const float gx = bx * t00 + t02, gy = by * t11 + t12;
        // F.grid_sample, align_corners=False
        const float ix = ((gx + 1.f) * (float)W - 1.f) / 2.f, iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const bool fin = isfinite(ix) && isfinite(iy);
        const int x0 = fin ? (int)fmaxf(fminf(fx0, 1e7f), -1e7f) : -100, y0 = fin ? (int)fmaxf(fminf(fy0, 1e7f), -1e7f) : -100;
        wnw[u] = (fx0 + 1.f - ix) * (fy0 + 1.f - iy); wne[u] = (ix - fx0) * (fy0 + 1.f - iy);
        wsw[u] = (fx0 + 1.f - ix) * (iy - fy0);       wse[u] = (ix - fx0) * (iy - fy0);
        const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
        k00[u] = fin && ya && xa; k01[u] = fin && ya && xb; k10[u] = fin && yb && xa; k11[u] = fin && yb && xb;
        const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1), yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        o00[u] = yc0 * W + xc0; o01[u] = yc0 * W + xc1; o10[u] = yc1 * W + xc0; o11[u] = yc1 * W + xc1;
    }
    const size_t HW = (size_t)H * W;
    // ---- image channels: NCHW out, direct coalesced writes
    if (image && image_crop) {
        for (int c = warp; c < Ci; c += 8) {
            const float* pl = image + ((size_t)b * Ci + c) * HW;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float v = sample_plane(pl, H, W, 0, 0, wnw[u], wne[u], wsw[u], wse[u], k00[u], k01[u], k10[u], k11[u], o00[u], o01[u],
                                             o10[u], o11[u]);
                if (live[u]) image_crop[((size_t)b * Ci + c) * Nc + k0 + lane + 32 * u] = v;
            }
        }
    }
    if (!geo || !geo_crop) return;
    // ---- descriptor channels, 32 at a time
    for (int c0 = 0; c0 < Cg; c0 += 32) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c0 + warp * 4 + cc;
            if (c >= Cg) continue;
            const float* pl = geo + ((size_t)b * Cg + c) * HW;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float v = sample_plane(pl, H, W, 0, 0, wnw[u], wne[u], wsw[u], wse[u], k00[u], k01[u], k10[u], k11[u], o00[u], o01[u],
                                             o10[u], o11[u]);
                if (geo_channels_last) tile[warp * 4 + cc][lane + 32 * u] = v;
                else if (live[u]) geo_crop[((size_t)b * Cg + c) * Nc + k0 + lane + 32 * u] = v;
            }
        }
        if (geo_channels_last) {                       // Cg == 32 (checked on the host): one 128-byte line per pixel
            __syncthreads();
            const int cg = threadIdx.x & 7;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int t = (threadIdx.x >> 3) + 32 * u;
                if (k0 + t >= Nc) continue;
                reinterpret_cast<float4*>(geo_crop + ((size_t)b * Nc + k0 + t) * 32)[cg] =
                    make_float4(tile[4 * cg][t], tile[4 * cg + 1][t], tile[4 * cg + 2][t], tile[4 * cg + 3][t]);
            }
            __syncthreads();
        }
    }
}

}  // namespace

size_t b2p_zoom_crop_ws_bytes(int B) { return align_up((size_t)B * 4 * sizeof(int), 256); }

int b2p_zoom_crop(const float* pc_depth, const float* K, const float* T, const float* image, const float* geo, int B, int Ci, int Cg,
                  int H, int W, int Hc, int Wc, float margin_ratio, int geo_channels_last, float* image_crop, float* geo_crop,
                  float* K_crop, float* theta, void* ws, cudaStream_t s) {
    int* bbox = reinterpret_cast<int*>(ws);
    zc_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(bbox, B);
    B2P_LAUNCH_CHECK();
    zc_bbox_kernel<<<dim3((unsigned)(H < 64 ? H : 64), (unsigned)B), 256, 0, s>>>(pc_depth, H, W, bbox);
    B2P_LAUNCH_CHECK();
    zc_resample_kernel<<<dim3((unsigned)ceil_div(Hc * Wc, ZC_PX), (unsigned)B), 256, 0, s>>>(bbox, K, T, image, geo, Ci, Cg, H, W, Hc, Wc,
                                                                                             margin_ratio, geo_channels_last, image_crop,
                                                                                             geo_crop, K_crop, theta);
    B2P_LAUNCH_CHECK();
    return 0;
}
