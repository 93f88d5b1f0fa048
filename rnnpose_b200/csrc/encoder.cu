// RAFT BasicEncoder (the feature extractor in front of the loop) on the tcgen05 convolution engine.
// Replaces ImageFeaEncoder.forward (reference model/CFNet.py:26-49: both images normalised as 2 (x / 255) - 1, batched through
// one BasicEncoder) and BasicEncoder / ResidualBlock with norm_fn = 'instance' (thirdparty/raft/extractor.py:6-57,118-232):
//   conv1 7x7 s2 3->64, InstanceNorm, ReLU;  layer1: 2 residual blocks 64->64;  layer2: 64->96 (first block stride 2 with a
//   1x1 s2 down-sampling branch + InstanceNorm);  layer3: 96->128 likewise;  conv2 1x1 128->256.
//   ResidualBlock: y = relu(norm1(conv1 x)); y = relu(norm2(conv2 y)); x = downsample(x) if any; out = relu(x + y).
//   InstanceNorm2d defaults: no affine parameters, biased variance, eps = 1e-5.
// The reference runs this under fp16 autocast on its GPU path and in fp32 on CPU; the oracle is the CPU fp32 result (SURVEY
// Appendix D10).  Here: the 7x7 stem re-indexed into a 4x1 convolution of gathered channels (enc_stem_im2col_kernel; or an exact fp32 FFMA kernel), the fifteen
// other convolutions on tcgen05 with fp16 hi/lo split operands (conv_umma.cu; stride 2 through the TMA box's element strides),
// InstanceNorm as a deterministic two-stage fp64 reduction + one fused normalise / ReLU / residual / operand-split pass.
#include "common.cuh"

namespace {

enum { EC_STEM = 0, EC_COUNT = 16 };        // state-dict order: conv1, layer1.{0,1}.conv{1,2}, layer2.0.{conv1,conv2,downsample.0}, ...

struct EncDesc { int cin, cout, k, stride; };
const EncDesc kEnc[EC_COUNT] = {
    {3, 64, 7, 2},                                                      // 0  conv1 (stem)
    {64, 64, 3, 1}, {64, 64, 3, 1}, {64, 64, 3, 1}, {64, 64, 3, 1},     // 1-4   layer1.0.conv1/2, layer1.1.conv1/2
    {64, 96, 3, 2}, {96, 96, 3, 1}, {64, 96, 1, 2},                     // 5-7   layer2.0.conv1, conv2, downsample.0
    {96, 96, 3, 1}, {96, 96, 3, 1},                                     // 8-9   layer2.1.conv1/2
    {96, 128, 3, 2}, {128, 128, 3, 1}, {96, 128, 1, 2},                 // 10-12 layer3.0.conv1, conv2, downsample.0
    {128, 128, 3, 1}, {128, 128, 3, 1},                                 // 13-14 layer3.1.conv1/2
    {128, 256, 1, 1},                                                   // 15 conv2
};

struct EncLayout {
    size_t stem_w, stem_b;                 // float offsets: stem weights [147][64] (k = c*49 + ky*7 + kx), bias [64]
    size_t bias[EC_COUNT];                 // float offsets of the (padded) biases of the tensor-core layers
    size_t f32_floats;
    int cin_pad[EC_COUNT], cout_pad[EC_COUNT], n_tile[EC_COUNT];
    size_t hi[EC_COUNT], lo[EC_COUNT];     // half offsets inside the fp16 section
    size_t halves;
};

const EncLayout& enc_layout() {
    static const EncLayout L = []() {
        EncLayout l;
        memset(&l, 0, sizeof(l));
        size_t off = 0;
        l.stem_w = off; off += 147 * 64;
        l.stem_b = off; off += 64;
        size_t h = 0;
        // the stem as a tensor-core layer (enc_stem_im2col_kernel): 4 vertical taps x 64 (48 used) gathered channels -> 64
        l.cin_pad[0] = 64; l.n_tile[0] = 64; l.cout_pad[0] = 64; l.bias[0] = l.stem_b;
        l.hi[0] = h; h += 4 * 64 * 64;
        l.lo[0] = h; h += 4 * 64 * 64;
        for (int i = 1; i < EC_COUNT; ++i) {
            l.cin_pad[i] = (kEnc[i].cin + 63) / 64 * 64;
            l.n_tile[i] = kEnc[i].cout == 256 ? 128 : kEnc[i].cout;       // 64, 96, 128 (pairs load 256-row tiles for conv2)
            l.cout_pad[i] = (kEnc[i].cout + l.n_tile[i] - 1) / l.n_tile[i] * l.n_tile[i];
            l.bias[i] = off; off += (size_t)l.cout_pad[i];
            const size_t n = (size_t)kEnc[i].k * kEnc[i].k * l.cout_pad[i] * l.cin_pad[i];
            l.hi[i] = h; h += n;
            l.lo[i] = h; h += n;
            h = (h + 127) / 128 * 128;
        }
        l.f32_floats = (off + 255) / 256 * 256;
        l.halves = h;
        return l;
    }();
    return L;
}

size_t enc_half_offset_bytes() { return align_up(enc_layout().f32_floats * sizeof(float), 1024); }

// ---------------------------------------------------------------------------------------------- weight packing
__global__ void enc_pack_stem_kernel(const float* __restrict__ w /*[64][3][7][7]*/, const float* __restrict__ b, float* __restrict__ dw,
                                     float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64 * 147) { const int n = i / 147, k = i - n * 147; dw[k * 64 + n] = w[i]; }
    if (i < 64) db[i] = b[i];
}

// The stem on the tensor cores.  A 7x7 stride-2 convolution of 3 channels is a 4x4 stride-1 convolution of the 2x2 pixel-unshuffled
// image (12 channels; input row 2 oy + ky - 3 = 2 (oy + t - 2) + py with t = 0..3, py = 0,1, ky = 2 t - 1 + py, and ky = -1 gets a
// zero weight).  The four horizontal taps are gathered into the channel dimension by enc_stem_im2col_kernel, so the engine sees a
// 4x1 convolution of 64 (48 used) channels: channel = j*12 + py*6 + px*3 + c, j = horizontal tap.  K = 256 instead of 147.
__device__ __forceinline__ void stem_channel(int ch, int& j, int& py, int& px, int& c) {
    j = ch / 12; const int r = ch - j * 12; py = r / 6; px = (r - py * 6) / 3; c = r % 3;
}

__global__ void enc_pack_stem_tc_kernel(const float* __restrict__ w /*[64][3][7][7]*/, __half* __restrict__ hi, __half* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // [tap t][n][ch]
    if (i >= 4 * 64 * 64) return;
    const int ch = i & 63, n = (i >> 6) & 63, t = i >> 12;
    float v = 0.f;
    if (ch < 48) {
        int j, py, px, c;
        stem_channel(ch, j, py, px, c);
        const int ky = 2 * t - 1 + py, kx = 2 * j - 1 + px;
        if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((n * 3 + c) * 7 + ky) * 7 + kx];
    }
    b2p_split_half(v, hi[i], lo[i]);
}

// images [B][3][H][W] x 2 -> operand planes [NI][H/2][W/2][64]: normalised 2 (x / 255) - 1 (CFNet.py:42-43), zero outside the image
// (the convolution pads the NORMALISED image) and in the 16 pad channels.  One thread per output pixel: for each (py, colour) the
// eight consecutive image columns 2 (ox - 2) .. 2 (ox + 1) + 1 are the (j, px) pairs in order, so every index below is a
// compile-time constant.  The 2 x 128 bytes of a pixel go through shared memory (16-byte chunks XOR-swizzled by the pixel
// index) so that the block's global stores are linear.
constexpr int SI_PIX = 128;

__global__ void __launch_bounds__(SI_PIX) enc_stem_im2col_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int B, int H,
                                                                 int W, int H1, int W1, size_t total_px, __half* __restrict__ hi,
                                                                 __half* __restrict__ lo) {
    __shared__ uint4 sh[2][SI_PIX * 8];
    const size_t pix0 = (size_t)blockIdx.x * SI_PIX;
    const size_t pix = pix0 + threadIdx.x;
    if (pix < total_px) {
        const int ox = (int)(pix % W1);
        const size_t t = pix / W1;
        const int qy = (int)(t % H1), n = (int)(t / H1);
        const float* img = (n < B ? img1 + (size_t)n * 3 * H * W : img2 + (size_t)(n - B) * 3 * H * W);
        const int ix0 = 2 * (ox - 2);
        __half vh[64], vl[64];
#pragma unroll
        for (int ch = 48; ch < 64; ++ch) { vh[ch] = __float2half_rn(0.f); vl[ch] = __float2half_rn(0.f); }
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* row = img + ((size_t)c * H + (2 * qy + py)) * W;        // 2 qy + py < H always (H even)
#pragma unroll
                for (int k = 0; k < 8; k += 2) {                                       // k = 2 j + px; ix0 is even: aligned pairs
                    const int ix = ix0 + k;
                    float2 v = make_float2(0.f, 0.f);
                    if (ix >= 0 && ix + 1 < W) {
                        v = __ldg(reinterpret_cast<const float2*>(row + ix));
                        v.x = 2.f * (v.x / 255.f) - 1.f; v.y = 2.f * (v.y / 255.f) - 1.f;
                    }
                    const int j = k >> 1;
                    b2p_split_half(v.x, vh[j * 12 + py * 6 + c], vl[j * 12 + py * 6 + c]);
                    b2p_split_half(v.y, vh[j * 12 + py * 6 + 3 + c], vl[j * 12 + py * 6 + 3 + c]);
                }
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e] = (uint32_t)__half_as_ushort(vh[k * 8 + 2 * e]) | ((uint32_t)__half_as_ushort(vh[k * 8 + 2 * e + 1]) << 16);
                b[e] = (uint32_t)__half_as_ushort(vl[k * 8 + 2 * e]) | ((uint32_t)__half_as_ushort(vl[k * 8 + 2 * e + 1]) << 16);
            }
            const int slot = threadIdx.x * 8 + (k ^ (threadIdx.x & 7));
            sh[0][slot] = make_uint4(a[0], a[1], a[2], a[3]);
            sh[1][slot] = make_uint4(b[0], b[1], b[2], b[3]);
        }
    }
    __syncthreads();
    const size_t left = total_px - pix0 < (size_t)SI_PIX ? total_px - pix0 : (size_t)SI_PIX;
    uint4* oh = reinterpret_cast<uint4*>(hi) + pix0 * 8;
    uint4* ol = reinterpret_cast<uint4*>(lo) + pix0 * 8;
    for (int i = threadIdx.x; i < (int)left * 8; i += SI_PIX) {
        const int p = i >> 3, k = i & 7;
        const int slot = p * 8 + (k ^ (p & 7));
        oh[i] = sh[0][slot];
        ol[i] = sh[1][slot];
    }
}

// dst[tap][n][c] = split(src[n][c][ky][kx]), rows of cin_pad halves (K-major); bias copied to its padded slot
__global__ void enc_pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, int cout, int cin, int k,
                                     __half* __restrict__ hi, __half* __restrict__ lo, int cin_pad, int cout_pad, float* __restrict__ db) {
    const size_t total = (size_t)cout * cin * k * k;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)cout) db[i] = b[i];
    if (i >= total) return;
    const int kx = (int)(i % k); size_t t = i / k;
    const int ky = (int)(t % k); t /= k;
    const int c = (int)(t % cin);
    const int n = (int)(t / cin);
    const size_t o = ((size_t)(ky * k + kx) * cout_pad + n) * cin_pad + c;
    b2p_split_half(w[i], hi[o], lo[o]);
}

// ---------------------------------------------------------------------------------------------- stem: 7x7 stride 2, 3 -> 64, fp32
// Block = 16 x 16 output pixels x 64 channels (256 threads: pixel pair = tid & 127 -> rows ty and ty + 8, channel half = tid >> 7).
// The input patch (37 x 37 x 3, normalised 2 (x / 255) - 1, zero outside the image: the convolution pads the NORMALISED image)
// and all weights live in shared memory; a warp shares one channel half, so every weight read is a broadcast, and each
// 128-bit weight read feeds 8 FFMAs (two pixels): 10 shared-memory instructions per 64 FFMAs (round 2 first version: one
// pixel per thread, 9 per 32, 856 us for 64 images; profiles/r2m).
constexpr int ST_TH = 16, ST_TW = 16, ST_PH = ST_TH * 2 + 5, ST_PW = ST_TW * 2 + 5;
constexpr size_t ST_SMEM_BYTES = (147 * 64 + 3 * ST_PH * (ST_PW + 1)) * sizeof(float);

__global__ void __launch_bounds__(256) enc_stem_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int B, int H, int W,
                                                       int H1, int W1, const float* __restrict__ wk /*[147][64]*/,
                                                       const float* __restrict__ bias, float* __restrict__ out /*[NI*H1*W1][64]*/) {
    extern __shared__ __align__(16) float st_smem[];               // ST_SMEM_BYTES (dynamic: more than the 48 KB static limit)
    float* ws = st_smem;
    float (*patch)[ST_PH][ST_PW + 1] = reinterpret_cast<float (*)[ST_PH][ST_PW + 1]>(st_smem + 147 * 64);
    const int n = blockIdx.z;
    const float* img = (n < B ? img1 + (size_t)n * 3 * H * W : img2 + (size_t)(n - B) * 3 * H * W);
    const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
    for (int i = threadIdx.x; i < 147 * 16; i += 256) reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);
    for (int i = threadIdx.x; i < 3 * ST_PH * ST_PW; i += 256) {
        const int c = i / (ST_PH * ST_PW), r = i - c * (ST_PH * ST_PW);
        const int py = r / ST_PW, px = r - py * ST_PW;
        const int iy = oy0 * 2 - 3 + py, ix = ox0 * 2 - 3 + px;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = 2.f * (__ldg(img + ((size_t)c * H + iy) * W + ix) / 255.f) - 1.f;   // CFNet.py:42-43
        patch[c][py][px] = v;
    }
    __syncthreads();
    const int pix = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int ty = pix / ST_TW, tx = pix - ty * ST_TW;          // ty 0..7; the second pixel is 8 rows below
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc0[j] = acc1[j] = __ldg(bias + half * 32 + j);
    for (int c = 0; c < 3; ++c)
        for (int ky = 0; ky < 7; ++ky)
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float v0 = patch[c][ty * 2 + ky][tx * 2 + kx];
                const float v1 = patch[c][ty * 2 + 16 + ky][tx * 2 + kx];
                const float4* w4 = reinterpret_cast<const float4*>(ws + ((c * 7 + ky) * 7 + kx) * 64 + half * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = w4[j];
                    acc0[4 * j] += v0 * q.x; acc0[4 * j + 1] += v0 * q.y; acc0[4 * j + 2] += v0 * q.z; acc0[4 * j + 3] += v0 * q.w;
                    acc1[4 * j] += v1 * q.x; acc1[4 * j + 1] += v1 * q.y; acc1[4 * j + 2] += v1 * q.z; acc1[4 * j + 3] += v1 * q.w;
                }
            }
    const int ox = ox0 + tx;
    if (ox < W1) {
        const int oya = oy0 + ty, oyb = oya + 8;
        if (oya < H1) {
            float4* o = reinterpret_cast<float4*>(out + (((size_t)n * H1 + oya) * W1 + ox) * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(acc0[4 * j], acc0[4 * j + 1], acc0[4 * j + 2], acc0[4 * j + 3]);
        }
        if (oyb < H1) {
            float4* o = reinterpret_cast<float4*>(out + (((size_t)n * H1 + oyb) * W1 + ox) * 64 + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(acc1[4 * j], acc1[4 * j + 1], acc1[4 * j + 2], acc1[4 * j + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- InstanceNorm statistics
// x: [NI][P][C] fp32.  Grid (chunks, NI); a thread owns 4 channels of every (256 / (C/4))-th pixel of its chunk and
// accumulates sum / sum of squares in fp64; the pixel lanes of the block are then summed in a fixed order.
// The block that finishes an image last (ticket counter, reset for the next use) folds the partials in chunk order -- the same
// sums whichever block that is -- into mean and 1 / sqrt(var + eps): biased variance, eps = 1e-5 (nn.InstanceNorm2d defaults).
__global__ void __launch_bounds__(256) in_stats1_kernel(const float* __restrict__ x, int P, int C, int chunk_px,
                                                        double* part /*[NI][chunks][C][2]*/, int* counter /*[NI], zero*/,
                                                        float2* __restrict__ stats /*[NI][C]*/) {
    extern __shared__ double sm[];                 // [lanes][C][2]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int c4n = C >> 2, lanes = 256 / c4n;
    const int pl = threadIdx.x / c4n, c4 = threadIdx.x - pl * c4n;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (pl < lanes) {
        const int p0 = chunk * chunk_px, p1 = min(P, p0 + chunk_px);
        const float4* xb = reinterpret_cast<const float4*>(x + (size_t)n * P * C) + c4;
        for (int p = p0 + pl; p < p1; p += lanes) {
            const float4 v = __ldg(xb + (size_t)p * c4n);
            s[0] += v.x; q[0] += (double)v.x * v.x; s[1] += v.y; q[1] += (double)v.y * v.y;
            s[2] += v.z; q[2] += (double)v.z * v.z; s[3] += v.w; q[3] += (double)v.w * v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { sm[((size_t)pl * C + c4 * 4 + j) * 2] = s[j]; sm[((size_t)pl * C + c4 * 4 + j) * 2 + 1] = q[j]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        double a = 0, b = 0;
        for (int l = 0; l < lanes; ++l) { a += sm[((size_t)l * C + c) * 2]; b += sm[((size_t)l * C + c) * 2 + 1]; }
        double* o = part + (((size_t)n * gridDim.x + chunk) * C + c) * 2;
        o[0] = a; o[1] = b;
    }
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(counter + n, 1);
        is_last = (t == (int)gridDim.x - 1);
        if (is_last) counter[n] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int chunks = gridDim.x;
    for (int c = threadIdx.x; c < C; c += 256) {
        double a = 0, b = 0;
        for (int k = 0; k < chunks; ++k) {
            const double* o = part + (((size_t)n * chunks + k) * C + c) * 2;
            a += __ldcg(o); b += __ldcg(o + 1);
        }
        const double mean = a / P, var = fmax(b / P - mean * mean, 0.0);
        stats[(size_t)n * C + c] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
    }
}

// ---------------------------------------------------------------------------------------------- normalise / ReLU / residual / split
// v = relu((y - mean_y) rstd_y).  mode 0: out = v.  mode 1: out = relu(res + v) with res an fp32 map of the same shape (the
// block input).  mode 2: out = relu((d - mean_d) rstd_d + v) (the down-sampling branch, extractor.py:40-43,51-54).
// Writes the fp32 map (optional; it is the next block's residual) and the fp16 hi / lo operand planes of the next convolution.
__global__ void __launch_bounds__(256) in_apply_kernel(const float* __restrict__ y, const float2* __restrict__ st_y, int mode,
                                                       const float* res, const float2* __restrict__ st_d, int P, int C,
                                                       size_t total4, float* out_f32 /* may alias res (in place) */, __half* __restrict__ out_hi,
                                                       __half* __restrict__ out_lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4n = C >> 2;
    const int c = (int)(i % c4n) * 4;
    const size_t pix = i / c4n;
    const int n = (int)(pix / P);
    const float4 v = __ldg(reinterpret_cast<const float4*>(y) + i);
    const float2* sy = st_y + (size_t)n * C + c;
    float o[4] = {fmaxf((v.x - sy[0].x) * sy[0].y, 0.f), fmaxf((v.y - sy[1].x) * sy[1].y, 0.f), fmaxf((v.z - sy[2].x) * sy[2].y, 0.f),
                  fmaxf((v.w - sy[3].x) * sy[3].y, 0.f)};
    if (mode == 1) {
        const float4 r = reinterpret_cast<const float4*>(res)[i];
        o[0] = fmaxf(r.x + o[0], 0.f); o[1] = fmaxf(r.y + o[1], 0.f); o[2] = fmaxf(r.z + o[2], 0.f); o[3] = fmaxf(r.w + o[3], 0.f);
    } else if (mode == 2) {
        const float4 r = reinterpret_cast<const float4*>(res)[i];
        const float2* sd = st_d + (size_t)n * C + c;
        o[0] = fmaxf((r.x - sd[0].x) * sd[0].y + o[0], 0.f); o[1] = fmaxf((r.y - sd[1].x) * sd[1].y + o[1], 0.f);
        o[2] = fmaxf((r.z - sd[2].x) * sd[2].y + o[2], 0.f); o[3] = fmaxf((r.w - sd[3].x) * sd[3].y + o[3], 0.f);
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (out_hi) {
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b2p_split_half(o[j], h[j], l[j]);
        uint2 wh, wl;
        wh.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
        wh.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        wl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
        wl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
        reinterpret_cast<uint2*>(out_hi)[i] = wh;
        reinterpret_cast<uint2*>(out_lo)[i] = wl;
    }
}

// [NI][P][C] pixel-major -> fmap1 / fmap2 [B][C][P] (NCHW), images 0..B-1 and B..2B-1
__global__ void __launch_bounds__(256) enc_out_nchw_kernel(const float* __restrict__ x, int B, int P, int C, float* __restrict__ f1,
                                                           float* __restrict__ f2) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int pp = ty; pp < 32; pp += 8) {
        const int p = p0 + pp, c = c0 + tx;
        tile[pp][tx] = (p < P && c < C) ? __ldg(x + ((size_t)n * P + p) * C + c) : 0.f;
    }
    __syncthreads();
    float* dst = n < B ? f1 + (size_t)n * C * P : f2 + (size_t)(n - B) * C * P;
    for (int cc = ty; cc < 32; cc += 8) {
        const int p = p0 + tx, c = c0 + cc;
        if (p < P && c < C) dst[(size_t)c * P + p] = tile[tx][cc];
    }
}

struct EncWs {
    float *r, *y, *d;            // fp32 maps: block input / residual, raw convolution output, raw down-sampling branch
    __half *xh[2], *yh[2];       // operand planes of the block input and of relu(norm1(conv1 x))
    double* part; float2 *st_y, *st_d; int* counter;
    float* zero_bias_unused;
};
constexpr int IN_CHUNKS = 32;

inline int down2(int v) { return (v - 1) / 2 + 1; }

size_t enc_ws_layout(int NI, int H, int W, void* ws, EncWs* out) {
    const int H1 = down2(H), W1 = down2(W);
    const size_t e = (size_t)NI * H1 * W1 * 64;          // the largest map (layer1); later stages are smaller
    char* base = reinterpret_cast<char*>(ws);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 1024); void* p = base ? base + off : nullptr; off += bytes; return p; };
    EncWs w;
    w.r = (float*)take(e * 4); w.y = (float*)take(e * 4); w.d = (float*)take(e * 4);
    for (int k = 0; k < 2; ++k) { w.xh[k] = (__half*)take(e * 2); w.yh[k] = (__half*)take(e * 2); }
    w.part = (double*)take((size_t)NI * IN_CHUNKS * 256 * 2 * sizeof(double));
    w.st_y = (float2*)take((size_t)NI * 256 * sizeof(float2));
    w.st_d = (float2*)take((size_t)NI * 256 * sizeof(float2));
    w.counter = (int*)take((size_t)NI * sizeof(int));
    w.zero_bias_unused = nullptr;
    if (out) *out = w;
    return align_up(off, 1024);
}

int in_stats(const float* x, int NI, int P, int C, const EncWs& w, float2* stats, cudaStream_t s) {
    const int chunk_px = ceil_div(P, IN_CHUNKS);
    const int chunks = ceil_div(P, chunk_px);
    const int lanes = 256 / (C / 4);
    in_stats1_kernel<<<dim3((unsigned)chunks, (unsigned)NI), 256, (size_t)lanes * C * 2 * sizeof(double), s>>>(x, P, C, chunk_px, w.part,
                                                                                                                 w.counter, stats);
    B2P_LAUNCH_CHECK();
    return 0;
}

int in_apply(const float* y, const float2* st_y, int mode, const float* res, const float2* st_d, int NI, int P, int C, float* out_f32,
             __half* hi, __half* lo, cudaStream_t s) {
    const size_t total4 = (size_t)NI * P * (C / 4);
    in_apply_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(y, st_y, mode, res, st_d, P, C, total4, out_f32, hi, lo);
    B2P_LAUNCH_CHECK();
    return 0;
}

// one tensor-core convolution of the encoder: planes [NI][in_h*in_w][cin] -> raw output + bias, fp32 [NI][h*w][cout]
int enc_conv(const float* packed, int id, __half* const* in, int NI, int in_h, int in_w, int h, int w, float* out, cudaStream_t s) {
    const EncLayout& L = enc_layout();
    const __half* hbase = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(packed) + enc_half_offset_bytes());
    UmmaConvArgs a;
    memset(&a, 0, sizeof(a));
    const bool stem = id == EC_STEM;                      // the gathered 4x1 form, see enc_stem_im2col_kernel
    a.seg_hi[0] = in[0]; a.seg_lo[0] = in[1]; a.seg_c[0] = stem ? 64 : kEnc[id].cin; a.seg_pitch[0] = stem ? 64 : kEnc[id].cin;
    a.w_hi = hbase + L.hi[id]; a.w_lo = hbase + L.lo[id]; a.bias = packed + L.bias[id];
    a.cin_pad = L.cin_pad[id]; a.cout_pad = L.cout_pad[id]; a.cout = kEnc[id].cout; a.n_tile = L.n_tile[id];
    a.kh = a.kw = kEnc[id].k;
    a.B = NI; a.h = h; a.w = w; a.stride = kEnc[id].stride; a.in_h = in_h; a.in_w = in_w;
    if (stem) { a.kh = 4; a.kw = 1; a.stride = 1; }
    a.epi = EPI_SCALE; a.scale = 1.f; a.out_f32 = out; a.out_f32_pitch = kEnc[id].cout;
    a.layer_id = -1;
    return b2p_launch_conv_umma(a, s);
}

}  // namespace

size_t b2p_encoder_packed_bytes() { return enc_half_offset_bytes() + enc_layout().halves * sizeof(__half); }

int b2p_encoder_pack(const float* const* t /*32 device pointers, state-dict order*/, void* packed, cudaStream_t s) {
    const EncLayout& L = enc_layout();
    float* f = reinterpret_cast<float*>(packed);
    __half* hbase = reinterpret_cast<__half*>(reinterpret_cast<char*>(packed) + enc_half_offset_bytes());
    B2P_CUDA(cudaMemsetAsync(packed, 0, b2p_encoder_packed_bytes(), s));
    enc_pack_stem_kernel<<<ceil_div(64 * 147, 256), 256, 0, s>>>(t[0], t[1], f + L.stem_w, f + L.stem_b);
    B2P_LAUNCH_CHECK();
    enc_pack_stem_tc_kernel<<<ceil_div(4 * 64 * 64, 256), 256, 0, s>>>(t[0], hbase + L.hi[0], hbase + L.lo[0]);
    B2P_LAUNCH_CHECK();
    for (int i = 1; i < EC_COUNT; ++i) {
        const size_t total = (size_t)kEnc[i].cout * kEnc[i].cin * kEnc[i].k * kEnc[i].k;
        enc_pack_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(t[2 * i], t[2 * i + 1], kEnc[i].cout, kEnc[i].cin, kEnc[i].k,
                                                                             hbase + L.hi[i], hbase + L.lo[i], L.cin_pad[i], L.cout_pad[i],
                                                                             f + L.bias[i]);
        B2P_LAUNCH_CHECK();
    }
    return 0;
}

size_t b2p_encoder_ws_bytes(int B, int H, int W) { return enc_ws_layout(2 * B, H, W, nullptr, nullptr); }

static int encoder_pass(const void* packed_v, const float* image1, const float* image2, int B, int H, int W, float* fmap1, float* fmap2,
                        void* ws, cudaStream_t s);

// The network is HBM-bound on its normalisation passes (every convolution output is written as fp32, read for the statistics, read
// again to be normalised).  Running it over `enc_chunk` pairs at a time keeps a layer's maps small enough that the statistics and
// normalisation passes find the convolution output still in the 126 MB L2; InstanceNorm is per image, so chunking is exact.
int b2p_image_encoder(const void* packed_v, const float* image1, const float* image2, int B, int H, int W, float* fmap1, float* fmap2,
                      void* ws, cudaStream_t s) {
    int chunk = b2p_options().enc_chunk;
    if (chunk <= 0 || chunk > B) chunk = B;
    const size_t img = (size_t)3 * H * W, fm = (size_t)256 * (H / 8) * (W / 8);
    for (int p0 = 0; p0 < B; p0 += chunk) {
        const int nb = B - p0 < chunk ? B - p0 : chunk;
        const int rc = encoder_pass(packed_v, image1 + p0 * img, image2 + p0 * img, nb, H, W, fmap1 + p0 * fm, fmap2 + p0 * fm, ws, s);
        if (rc) return rc;
    }
    return 0;
}

static int encoder_pass(const void* packed_v, const float* image1, const float* image2, int B, int H, int W, float* fmap1, float* fmap2,
                        void* ws, cudaStream_t s) {
    const float* packed = reinterpret_cast<const float*>(packed_v);
    const EncLayout& L = enc_layout();
    const int NI = 2 * B;
    EncWs w;
    enc_ws_layout(NI, H, W, ws, &w);
    int rc;
    const int H1 = down2(H), W1 = down2(W), H2 = down2(H1), W2 = down2(W1), H3 = down2(H2), W3 = down2(W2);
    const int P1 = H1 * W1, P2 = H2 * W2, P3 = H3 * W3;
    // conv1 + norm1 + relu (extractor.py:200-202)
    B2P_CUDA(cudaMemsetAsync(w.counter, 0, (size_t)NI * sizeof(int), s));       // ticket counters of the statistics kernel
    if (b2p_options().enc_stem != 0) {            // tensor cores: gather, then a 4x1 convolution of 64 channels through the engine
        const size_t total_px = (size_t)NI * P1;
        enc_stem_im2col_kernel<<<(unsigned)((total_px + SI_PIX - 1) / SI_PIX), SI_PIX, 0, s>>>(image1, image2, B, H, W, H1, W1, total_px, w.xh[0],
                                                                                              w.xh[1]);
        B2P_LAUNCH_CHECK();
        if ((rc = enc_conv(packed, EC_STEM, w.xh, NI, H1, W1, H1, W1, w.y, s))) return rc;
    } else {                                      // exact fp32 FFMA kernel
        static const cudaError_t st_attr = cudaFuncSetAttribute(enc_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM_BYTES);
        B2P_CUDA(st_attr);
        enc_stem_kernel<<<dim3((unsigned)ceil_div(W1, ST_TW), (unsigned)ceil_div(H1, ST_TH), (unsigned)NI), 256, ST_SMEM_BYTES, s>>>(
            image1, image2, B, H, W, H1, W1, packed + L.stem_w, packed + L.stem_b, w.y);
        B2P_LAUNCH_CHECK();
    }
    if ((rc = in_stats(w.y, NI, P1, 64, w, w.st_y, s))) return rc;
    if ((rc = in_apply(w.y, w.st_y, 0, nullptr, nullptr, NI, P1, 64, w.r, w.xh[0], w.xh[1], s))) return rc;
    // residual blocks: (first conv id, input dims, output dims, channels, has down-sampling branch)
    struct Blk { int id, ih, iw, oh, ow, cout, ds; };
    const Blk blks[6] = {{1, H1, W1, H1, W1, 64, 0}, {3, H1, W1, H1, W1, 64, 0}, {5, H1, W1, H2, W2, 96, 1}, {8, H2, W2, H2, W2, 96, 0},
                         {10, H2, W2, H3, W3, 128, 1}, {13, H3, W3, H3, W3, 128, 0}};
    for (int bi = 0; bi < 6; ++bi) {
        const Blk& b = blks[bi];
        const int P = b.oh * b.ow;
        // y = relu(norm1(conv1 x))
        if ((rc = enc_conv(packed, b.id, w.xh, NI, b.ih, b.iw, b.oh, b.ow, w.y, s))) return rc;
        if ((rc = in_stats(w.y, NI, P, b.cout, w, w.st_y, s))) return rc;
        if ((rc = in_apply(w.y, w.st_y, 0, nullptr, nullptr, NI, P, b.cout, nullptr, w.yh[0], w.yh[1], s))) return rc;
        // y = relu(norm2(conv2 y)); x = downsample(x); x = relu(x + y)
        if ((rc = enc_conv(packed, b.id + 1, w.yh, NI, b.oh, b.ow, b.oh, b.ow, w.y, s))) return rc;
        if ((rc = in_stats(w.y, NI, P, b.cout, w, w.st_y, s))) return rc;
        if (b.ds) {
            if ((rc = enc_conv(packed, b.id + 2, w.xh, NI, b.ih, b.iw, b.oh, b.ow, w.d, s))) return rc;
            if ((rc = in_stats(w.d, NI, P, b.cout, w, w.st_d, s))) return rc;
            if ((rc = in_apply(w.y, w.st_y, 2, w.d, w.st_d, NI, P, b.cout, w.r, w.xh[0], w.xh[1], s))) return rc;
        } else {
            if ((rc = in_apply(w.y, w.st_y, 1, w.r, nullptr, NI, P, b.cout, w.r, w.xh[0], w.xh[1], s))) return rc;
        }
    }
    // conv2 1x1 128 -> 256 (extractor.py:217), then to the boundary layout [B,256,h,w] x 2
    if ((rc = enc_conv(packed, 15, w.xh, NI, H3, W3, H3, W3, w.y, s))) return rc;
    enc_out_nchw_kernel<<<dim3((unsigned)ceil_div(P3, 32), 8, (unsigned)NI), 256, 0, s>>>(w.y, B, P3, 256, fmap1, fmap2);
    B2P_LAUNCH_CHECK();
    (void)P2;
    return 0;
}
