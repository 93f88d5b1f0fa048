// Weight packing and the two "skinny" layers of the update block that do not fit the GEMM tile:
//   encoder.convf1 (7x7, Cin=2): lowered to a 98-wide im2col + 1x1 GEMM,
//   flow_head.conv2 (3x3, Cout=2): warp-per-pixel dot products, fused with coords1 += delta_flow.
// reference thirdparty/raft/update.py:13-14,83-97 ; model/CFNet.py:157.
#include "common.cuh"

#include <algorithm>

namespace {

// dst[tap][c][n_off + n] = src[n][c][ky][kx]     (tap = ky*kw + kx)
__global__ void pack_conv_kernel(const float* __restrict__ src, int cout, int cin, int kh, int kw,
                                 float* __restrict__ dst, int cin_pad, int cout_pad, int n_off, int flatten_taps) {
    const size_t total = (size_t)cout * cin * kh * kw;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int kx = (int)(i % kw); size_t t = i / kw;
    const int ky = (int)(t % kh); t /= kh;
    const int c = (int)(t % cin);
    const int n = (int)(t / cin);
    const int tap = ky * kw + kx;
    size_t o;
    if (flatten_taps) o = ((size_t)(tap * cin + c)) * cout_pad + n_off + n;            // 1x1 over im2col, k = tap*cin + c
    else o = ((size_t)tap * cin_pad + c) * cout_pad + n_off + n;
    dst[o] = src[i];
}

__global__ void pack_vec_kernel(const float* __restrict__ src, int n, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// fp16 hi/lo planes: dst[tap][n_off + n][c] = split(src[n][c][ky][kx])   (K-major rows of cin_pad halves)
__global__ void pack_conv_half_kernel(const float* __restrict__ src, int cout, int cin, int kh, int kw,
                                      __half* __restrict__ dst_hi, __half* __restrict__ dst_lo, int cin_pad, int cout_pad,
                                      int n_off, int flatten_taps) {
    const size_t total = (size_t)cout * cin * kh * kw;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int kx = (int)(i % kw); size_t t = i / kw;
    const int ky = (int)(t % kh); t /= kh;
    const int c = (int)(t % cin);
    const int n = (int)(t / cin);
    const int tap = ky * kw + kx;
    size_t o;
    if (flatten_taps) o = (size_t)(n_off + n) * cin_pad + (tap * cin + c);           // single "tap", k = tap*cin + c
    else o = ((size_t)tap * cout_pad + n_off + n) * cin_pad + c;
    __half hi, lo;
    b2p_split_half(src[i], hi, lo);
    dst_hi[o] = hi;
    dst_lo[o] = lo;
}

__global__ void split_planes_kernel(const float* __restrict__ src, int pitch_in, int C, size_t P, __half* __restrict__ hi,
                                    __half* __restrict__ lo, int pitch_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * (size_t)C) return;
    const size_t p = i / C; const int c = (int)(i - p * C);
    b2p_split_half(src[p * pitch_in + c], hi[p * pitch_out + c], lo[p * pitch_out + c]);
}

// flow_head.conv2 [2][256][3][3] -> [2][9][256]
__global__ void pack_fh2_kernel(const float* __restrict__ src, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * 256 * 9) return;
    const int tap = i % 9; int t = i / 9;
    const int c = t % 256; const int o = t / 256;
    dst[(o * 9 + tap) * 256 + c] = src[i];
}

__global__ void __launch_bounds__(256) im2col_f1_kernel(const float* __restrict__ flow, int B, int h, int w,
                                                        float* __restrict__ col, float* __restrict__ xbuf,
                                                        __half* __restrict__ col_hi, __half* __restrict__ col_lo,
                                                        __half* __restrict__ x_hi, __half* __restrict__ x_lo) {
    pdl_trigger();
    pdl_wait();
    const size_t total = (size_t)B * h * w * 56;          // 56 float2 slots per pixel (49 taps + 7 zero pads)
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int slot = (int)(i % 56);
    const size_t pix = i / 56;
    const int hw = h * w;
    const int b = (int)(pix / hw); const int r = (int)(pix - (size_t)b * hw);
    const int y = r / w, x = r - y * w;
    float2 v = make_float2(0.f, 0.f);
    if (slot < 49) {
        const int ky = slot / 7, kx = slot - ky * 7;
        const int sy = y + ky - 3, sx = x + kx - 3;
        if (sy >= 0 && sy < h && sx >= 0 && sx < w)
            v = *reinterpret_cast<const float2*>(flow + ((size_t)(b * h + sy) * w + sx) * 2);
        if (slot == 24) {                                                          // centre tap = flow itself -> cat[out, flow]
            if (xbuf) *reinterpret_cast<float2*>(xbuf + pix * 256 + 254) = v;
            if (x_hi) {
                b2p_split_half(v.x, x_hi[pix * 256 + 254], x_lo[pix * 256 + 254]);
                b2p_split_half(v.y, x_hi[pix * 256 + 255], x_lo[pix * 256 + 255]);
            }
        }
    }
    if (col) *reinterpret_cast<float2*>(col + pix * 112 + slot * 2) = v;
    if (col_hi) {
        b2p_split_half(v.x, col_hi[pix * 112 + slot * 2], col_lo[pix * 112 + slot * 2]);
        b2p_split_half(v.y, col_hi[pix * 112 + slot * 2 + 1], col_lo[pix * 112 + slot * 2 + 1]);
    }
}

// flow_head.conv2 (3x3, 256 -> 2, update.py:13-14) in two steps that read every feature vector once instead of nine times:
//   (1) partial: for every SOURCE pixel the 18 dot products <features(src), W2[o][tap]> (4 lanes per pixel, 64 channels each,
//       weights broadcast from shared memory), part[src][tap*2 + o];
//   (2) gather: delta(y,x)[o] = b2[o] + sum_tap part[(y+ky-1, x+kx-1)][tap*2+o]  (zero padding = skipped taps), then
//       coords1 += delta, flow = coords1 - coords0  (model/CFNet.py:157,166).
constexpr int FH2_PITCH = 20;      // 18 partial sums per pixel, rows padded to 16-byte multiples

constexpr int FH2_PX = 2;          // pixels per lane quad: every weight vector read from shared memory serves both

__global__ void __launch_bounds__(256) flow_head2_partial_kernel(const float* __restrict__ hm, const __half* __restrict__ hm_hi,
                                                                 const __half* __restrict__ hm_lo, const float* __restrict__ w2,
                                                                 float* __restrict__ part, int npix) {
    pdl_trigger();
    __shared__ __align__(16) float ws[18][256];            // [tap*2+o][c]
    for (int i = threadIdx.x; i < 18 * 64; i += blockDim.x) {          // float4 copies, all independent
        const int k = i >> 6, c4 = i & 63;
        *reinterpret_cast<float4*>(&ws[k][c4 * 4]) =
            __ldg(reinterpret_cast<const float4*>(w2 + ((k & 1) * 9 + (k >> 1)) * 256) + c4);      // weights: not written by the previous kernel
    }
    __syncthreads();
    pdl_wait();
    const int q = threadIdx.x & 3;
    // persistent blocks (the weights are staged once per block): lane quads walk the pixel pairs
    for (int quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; quad * FH2_PX < npix; quad += (gridDim.x * blockDim.x) >> 2) {
    const int pix0 = quad * FH2_PX;
    float acc[FH2_PX][18];
#pragma unroll
    for (int u = 0; u < FH2_PX; ++u)
#pragma unroll
        for (int k = 0; k < 18; ++k) acc[u][k] = 0.f;
    {
#pragma unroll 2
        for (int it = 0; it < 8; ++it) {
            const int c = (it * 4 + q) * 8;                // the quad reads 64 contiguous bytes of each fp16 plane
            float a[FH2_PX][8];
#pragma unroll
            for (int u = 0; u < FH2_PX; ++u) {
                const size_t so = (size_t)min(pix0 + u, npix - 1) * 512;
                if (hm) {
                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(hm + so + c));
                    const float4 a1 = __ldg(reinterpret_cast<const float4*>(hm + so + c + 4));
                    a[u][0] = a0.x; a[u][1] = a0.y; a[u][2] = a0.z; a[u][3] = a0.w;
                    a[u][4] = a1.x; a[u][5] = a1.y; a[u][6] = a1.z; a[u][7] = a1.w;
                } else {
                    const uint4 uh = __ldg(reinterpret_cast<const uint4*>(hm_hi + so + c));
                    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(hm_lo + so + c));
                    const __half2* hh = reinterpret_cast<const __half2*>(&uh);
                    const __half2* ll = reinterpret_cast<const __half2*>(&ul);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        a[u][2 * j] = __low2float(hh[j]) + __low2float(ll[j]);
                        a[u][2 * j + 1] = __high2float(hh[j]) + __high2float(ll[j]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 18; ++k) {
                const float4 u0 = *reinterpret_cast<const float4*>(&ws[k][c]);
                const float4 u1 = *reinterpret_cast<const float4*>(&ws[k][c + 4]);
#pragma unroll
                for (int u = 0; u < FH2_PX; ++u)
                    acc[u][k] += a[u][0] * u0.x + a[u][1] * u0.y + a[u][2] * u0.z + a[u][3] * u0.w + a[u][4] * u1.x + a[u][5] * u1.y +
                                 a[u][6] * u1.z + a[u][7] * u1.w;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < FH2_PX; ++u)
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], 1);
            acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], 2);
        }
    if (q == 0) {
#pragma unroll
        for (int u = 0; u < FH2_PX; ++u) {
            if (pix0 + u >= npix) continue;
            float4* d = reinterpret_cast<float4*>(part + (size_t)(pix0 + u) * FH2_PITCH);
            d[0] = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
            d[1] = make_float4(acc[u][4], acc[u][5], acc[u][6], acc[u][7]);
            d[2] = make_float4(acc[u][8], acc[u][9], acc[u][10], acc[u][11]);
            d[3] = make_float4(acc[u][12], acc[u][13], acc[u][14], acc[u][15]);
            d[4] = make_float4(acc[u][16], acc[u][17], 0.f, 0.f);
        }
    }
    }   // quad loop (uniform per warp up to whole quads: the shuffles above see every lane of a live quad)
}

__global__ void __launch_bounds__(128) flow_head2_gather_kernel(const float* __restrict__ part, const float* __restrict__ b2,
                                                                float* __restrict__ coords1, float* __restrict__ flow,
                                                                float* __restrict__ dflow_out, int B, int h, int w) {
    pdl_trigger();
    pdl_wait();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int hw = h * w;
    if (pix >= B * hw) return;
    const int b = pix / hw, r = pix - b * hw;
    const int y = r / w, x = r - y * w;
    float s0 = b2[0], s1 = b2[1];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int sy = y + tap / 3 - 1, sx = x + tap % 3 - 1;
        if (sy < 0 || sy >= h || sx < 0 || sx >= w) continue;
        const float2 v = __ldg(reinterpret_cast<const float2*>(part + ((size_t)(b * h + sy) * w + sx) * FH2_PITCH + tap * 2));
        s0 += v.x; s1 += v.y;
    }
    if (dflow_out) { dflow_out[(size_t)pix * 2] = s0; dflow_out[(size_t)pix * 2 + 1] = s1; }
    float2 c1 = *reinterpret_cast<float2*>(coords1 + (size_t)pix * 2);
    c1.x += s0; c1.y += s1;
    *reinterpret_cast<float2*>(coords1 + (size_t)pix * 2) = c1;
    *reinterpret_cast<float2*>(flow + (size_t)pix * 2) = make_float2(c1.x - (float)x, c1.y - (float)y);
}


// fp32 [P][C] pixel-major <-> the tiled side-buffer layout (common.cuh: b2p_tiled_index); one thread per float4
// xmajor: the 128 pixel slots of a tile are ordered x * 16 + y (the x-major layers of the chained launch) instead of y * 8 + x
template <bool TO_TILED>
__global__ void __launch_bounds__(256) tiled_convert_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int h,
                                                            int w, int C, int xmajor) {
    pdl_trigger();
    pdl_wait();
    const int tiles_x = (w + B2P_TILE_COLS - 1) / B2P_TILE_COLS, tiles_y = (h + B2P_TILE_ROWS - 1) / B2P_TILE_ROWS;
    const size_t total = (size_t)B * tiles_x * tiles_y * (C >> 2) * 128;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int m = (int)(i & 127);
    const size_t t2 = i >> 7;
    const int c4 = (int)(t2 % (size_t)(C >> 2));
    const int tile = (int)(t2 / (size_t)(C >> 2));
    const int b = tile / (tiles_x * tiles_y), tr = tile - b * tiles_x * tiles_y;
    const int y = (tr / tiles_x) * B2P_TILE_ROWS + (xmajor ? (m & 15) : (m >> 3)), x = (tr % tiles_x) * B2P_TILE_COLS + (xmajor ? (m >> 4) : (m & 7));
    const bool in = y < h && x < w;
    const size_t pxc = (((size_t)b * h + y) * w + x) * C + c4 * 4;
    if (TO_TILED) {
        reinterpret_cast<float4*>(dst)[i] = in ? *reinterpret_cast<const float4*>(src + pxc) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (in) {
        *reinterpret_cast<float4*>(dst + pxc) = reinterpret_cast<const float4*>(src)[i];
    }
}

B2PWeightLayout make_layout() {
    B2PWeightLayout L;
    auto set = [&](int id, int kh, int kw, int cin, int cout) {
        B2PConvDesc& d = L.cv[id];
        d.kh = kh; d.kw = kw; d.cin = cin; d.cout = cout;
        d.cin_pad = (cin + 15) / 16 * 16; d.cout_pad = (cout + 63) / 64 * 64;
    };
    set(CV_C1, 1, 1, 324, 256);
    set(CV_C2, 3, 3, 256, 192);
    set(CV_F1, 1, 1, 98, 128);
    set(CV_F2, 3, 3, 128, 64);
    set(CV_ENC, 3, 3, 256, 126);
    set(CV_ZR1, 1, 5, 384, 256);
    set(CV_Q1, 1, 5, 384, 128);
    set(CV_ZR2, 5, 1, 384, 256);
    set(CV_Q2, 5, 1, 384, 128);
    set(CV_HEADS, 3, 3, 128, 512);
    set(CV_MASK2, 1, 1, 256, 576);
    size_t off = 0;
    for (int i = 0; i < CV_COUNT; ++i) {
        B2PConvDesc& d = L.cv[i];
        d.w_off = off; off += (size_t)d.kh * d.kw * d.cin_pad * d.cout_pad;
        d.b_off = off; off += d.cout_pad;
        off = (off + 63) / 64 * 64;
    }
    L.fh2_w_off = off; off += 2 * 9 * 256;
    L.fh2_b_off = off; off += 64;
    L.total_floats = off;
    return L;
}

}  // namespace

const B2PWeightLayout& b2p_weight_layout() {
    static const B2PWeightLayout L = make_layout();
    return L;
}

// tensor order (state-dict order):
//  0,1 convc1 | 2,3 convc2 | 4,5 convf1 | 6,7 convf2 | 8,9 conv | 10,11 convz1 | 12,13 convr1 | 14,15 convq1
//  16,17 convz2 | 18,19 convr2 | 20,21 convq2 | 22,23 flow_head.conv1 | 24,25 flow_head.conv2 | 26,27 mask.0 | 28,29 mask.2
int b2p_pack_weights(const float* const* t, float* packed, cudaStream_t s) {
    const B2PWeightLayout& L = b2p_weight_layout();
    B2P_CUDA(cudaMemsetAsync(packed, 0, L.total_floats * sizeof(float), s));
    auto pack = [&](int id, const float* w, const float* b, int cout_src, int cin_src, int kh, int kw, int n_off,
                    int flatten) -> int {
        const B2PConvDesc& d = L.cv[id];
        const size_t total = (size_t)cout_src * cin_src * kh * kw;
        pack_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, cout_src, cin_src, kh, kw, packed + d.w_off,
                                                                         d.cin_pad, d.cout_pad, n_off, flatten);
        pack_vec_kernel<<<ceil_div(cout_src, 256), 256, 0, s>>>(b, cout_src, packed + d.b_off + n_off);
        B2P_LAUNCH_CHECK();
        return 0;
    };
    int rc = 0;
    if ((rc = pack(CV_C1, t[0], t[1], 256, 324, 1, 1, 0, 0))) return rc;
    if ((rc = pack(CV_C2, t[2], t[3], 192, 256, 3, 3, 0, 0))) return rc;
    if ((rc = pack(CV_F1, t[4], t[5], 128, 2, 7, 7, 0, 1))) return rc;
    if ((rc = pack(CV_F2, t[6], t[7], 64, 128, 3, 3, 0, 0))) return rc;
    if ((rc = pack(CV_ENC, t[8], t[9], 126, 256, 3, 3, 0, 0))) return rc;
    if ((rc = pack(CV_ZR1, t[10], t[11], 128, 384, 1, 5, 0, 0))) return rc;
    if ((rc = pack(CV_ZR1, t[12], t[13], 128, 384, 1, 5, 128, 0))) return rc;
    if ((rc = pack(CV_Q1, t[14], t[15], 128, 384, 1, 5, 0, 0))) return rc;
    if ((rc = pack(CV_ZR2, t[16], t[17], 128, 384, 5, 1, 0, 0))) return rc;
    if ((rc = pack(CV_ZR2, t[18], t[19], 128, 384, 5, 1, 128, 0))) return rc;
    if ((rc = pack(CV_Q2, t[20], t[21], 128, 384, 5, 1, 0, 0))) return rc;
    if ((rc = pack(CV_HEADS, t[22], t[23], 256, 128, 3, 3, 0, 0))) return rc;
    if ((rc = pack(CV_HEADS, t[26], t[27], 256, 128, 3, 3, 256, 0))) return rc;
    if ((rc = pack(CV_MASK2, t[28], t[29], 576, 256, 1, 1, 0, 0))) return rc;
    pack_fh2_kernel<<<ceil_div(2 * 256 * 9, 256), 256, 0, s>>>(t[24], packed + L.fh2_w_off);
    pack_vec_kernel<<<1, 256, 0, s>>>(t[25], 2, packed + L.fh2_b_off);
    B2P_LAUNCH_CHECK();

    // ---- fp16 hi/lo section for the tensor-core path: W[tap][cout_pad][cin_pad] (K-major)
    const B2PHalfLayout& HL = b2p_half_layout();
    __half* hbase = reinterpret_cast<__half*>(reinterpret_cast<char*>(packed) + b2p_half_section_offset_bytes());
    B2P_CUDA(cudaMemsetAsync(hbase, 0, HL.total_halves * sizeof(__half), s));
    auto packh = [&](int id, const float* w, int cout_src, int cin_src, int kh, int kw, int n_off, int flatten) -> int {
        const B2PHalfConvDesc& d = HL.cv[id];
        const size_t total = (size_t)cout_src * cin_src * kh * kw;
        pack_conv_half_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, cout_src, cin_src, kh, kw, hbase + d.hi_off,
                                                                              hbase + d.lo_off, d.cin_pad, d.cout_pad, n_off,
                                                                              flatten);
        B2P_LAUNCH_CHECK();
        return 0;
    };
    if ((rc = packh(CV_C1, t[0], 256, 324, 1, 1, 0, 0))) return rc;
    if ((rc = packh(CV_C2, t[2], 192, 256, 3, 3, 0, 0))) return rc;
    if ((rc = packh(CV_F1, t[4], 128, 2, 7, 7, 0, 1))) return rc;
    if ((rc = packh(CV_F2, t[6], 64, 128, 3, 3, 0, 0))) return rc;
    if ((rc = packh(CV_ENC, t[8], 126, 256, 3, 3, 0, 0))) return rc;
    if ((rc = packh(CV_ZR1, t[10], 128, 384, 1, 5, 0, 0))) return rc;
    if ((rc = packh(CV_ZR1, t[12], 128, 384, 1, 5, 128, 0))) return rc;
    if ((rc = packh(CV_Q1, t[14], 128, 384, 1, 5, 0, 0))) return rc;
    if ((rc = packh(CV_ZR2, t[16], 128, 384, 5, 1, 0, 0))) return rc;
    if ((rc = packh(CV_ZR2, t[18], 128, 384, 5, 1, 128, 0))) return rc;
    if ((rc = packh(CV_Q2, t[20], 128, 384, 5, 1, 0, 0))) return rc;
    if ((rc = packh(CV_HEADS, t[22], 256, 128, 3, 3, 0, 0))) return rc;
    if ((rc = packh(CV_HEADS, t[26], 256, 128, 3, 3, 256, 0))) return rc;
    if ((rc = packh(CV_MASK2, t[28], 576, 256, 1, 1, 0, 0))) return rc;
    {   // flow_head.conv2 for the chained launch: rows 2..31 of every tap stay zero
        const B2PHalfConvDesc& d = HL.fh2;
        pack_conv_half_kernel<<<(unsigned)((2 * 256 * 9 + 255) / 256), 256, 0, s>>>(t[24], 2, 256, 3, 3, hbase + d.hi_off, hbase + d.lo_off,
                                                                                  d.cin_pad, d.cout_pad, 0, 0);
        B2P_LAUNCH_CHECK();
    }
    return 0;
}

const B2PHalfLayout& b2p_half_layout() {
    static const B2PHalfLayout HL = []() {
        B2PHalfLayout H;
        const B2PWeightLayout& L = b2p_weight_layout();
        const int ntile[CV_COUNT] = {128, 192, 128, 64, 128, 128, 128, 128, 128, 128, 192};
        size_t off = 0;
        for (int i = 0; i < CV_COUNT; ++i) {
            B2PHalfConvDesc& d = H.cv[i];
            d.kh = L.cv[i].kh; d.kw = L.cv[i].kw; d.cin = L.cv[i].cin; d.cout = L.cv[i].cout;
            d.n_tile = ntile[i];
            d.cin_pad = (d.cin + 63) / 64 * 64;
            d.cout_pad = (d.cout + d.n_tile - 1) / d.n_tile * d.n_tile;
            const size_t n = (size_t)d.kh * d.kw * d.cout_pad * d.cin_pad;
            d.hi_off = off; off += n;
            d.lo_off = off; off += n;
            off = (off + 127) / 128 * 128;
        }
        {
            B2PHalfConvDesc& d = H.fh2;
            d.kh = d.kw = 3; d.cin = 256; d.cout = 2; d.n_tile = 32; d.cin_pad = 256; d.cout_pad = 32;
            const size_t n = (size_t)9 * d.cout_pad * d.cin_pad;
            d.hi_off = off; off += n;
            d.lo_off = off; off += n;
            off = (off + 127) / 128 * 128;
        }
        H.total_halves = off;
        return H;
    }();
    return HL;
}

int b2p_split_planes(const float* src, int pitch_in, int C, size_t P, __half* hi, __half* lo, int pitch_out, cudaStream_t s) {
    const size_t total = P * (size_t)C;
    split_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, pitch_in, C, P, hi, lo, pitch_out);
    B2P_LAUNCH_CHECK();
    return 0;
}

size_t b2p_half_section_offset_bytes() { return align_up(b2p_weight_layout().total_floats * sizeof(float), 1024); }

int b2p_im2col_f1(const float* flow, int B, int h, int w, float* col, float* xbuf, __half* col_hi, __half* col_lo,
                  __half* x_hi, __half* x_lo, cudaStream_t s) {
    const size_t total = (size_t)B * h * w * 56;
    b2p_pdl_next_allowed() = b2p_pdl_allowed(2);
    B2P_CUDA(b2p_launch_pdl(im2col_f1_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, flow, B, h, w, col, xbuf, col_hi, col_lo, x_hi, x_lo));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_flow_head2(const float* hm, const __half* hm_hi, const __half* hm_lo, const float* w2, const float* b2,
                   float* coords1, float* flow, float* dflow_out, float* part, int B, int h, int w, cudaStream_t s) {
    const int npix = B * h * w;
    B2P_CUDA(b2p_launch_pdl(flow_head2_partial_kernel, dim3(std::min(ceil_div(ceil_div(npix, FH2_PX) * 4, 256), 2 * 148)), dim3(256), 0, s, hm, hm_hi, hm_lo, w2, part, npix));
    B2P_LAUNCH_CHECK();
    B2P_CUDA(b2p_launch_pdl(flow_head2_gather_kernel, dim3(ceil_div(npix, 128)), dim3(128), 0, s, (const float*)part, b2, coords1, flow,
                            dflow_out, B, h, w));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_pxc_to_tiled(const float* src, float* dst, int B, int h, int w, int C, cudaStream_t s, int xmajor) {
    const size_t total = b2p_tiled_pixels(B, h, w) * (size_t)(C >> 2);
    B2P_CUDA(b2p_launch_pdl(tiled_convert_kernel<true>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, src, dst, B, h, w, C, xmajor));
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_tiled_to_pxc(const float* src, float* dst, int B, int h, int w, int C, cudaStream_t s) {
    const size_t total = b2p_tiled_pixels(B, h, w) * (size_t)(C >> 2);
    B2P_CUDA(b2p_launch_pdl(tiled_convert_kernel<false>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, src, dst, B, h, w, C, 0));
    B2P_LAUNCH_CHECK();
    return 0;
}
