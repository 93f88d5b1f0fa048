// Exact-fp32 implicit-GEMM convolution over PXC activations (CUDA-core FFMA path) and the all-pairs
// correlation GEMM.  One CTA computes a 128 (pixels) x BN (output channels) tile with 256 threads,
// 8 x (BN/16) accumulators per thread, K consumed in chunks of 16 with register prefetch and a
// double-buffered shared-memory stage.  Replaces the cuDNN/cuBLAS calls issued by
//   reference thirdparty/raft/update.py:89-97 (motion encoder), :45-60 (SepConvGRU), :13-14,172-176 (heads)
//   reference thirdparty/raft/corr.py:60-67 (torch.matmul correlation volume).
// This is the bit-for-bit "fp32 everywhere" baseline; the tcgen05 path (conv_umma.cu) is gated
// against it by the parity tests.
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int BN>
__device__ __forceinline__ void tile_fma(const float (*As)[BM], const float (*Bs)[BN], int ty, int tx,
                                         float (&acc)[8][BN / 16]) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        float a[8], b[BN / 16];
        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        if (BN == 128) *reinterpret_cast<float4*>(&b[BN / 16 - 4]) = *reinterpret_cast<const float4*>(&Bs[k][BN / 2 + tx * 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < BN / 16; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

template <int BN>
__global__ void __launch_bounds__(256, 2) conv_gemm_kernel(const ConvParams p) {
    constexpr int TN = BN / 16;
    constexpr int NB4 = BN / 64;   // float4 weight loads per thread per chunk
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int hw = p.h * p.w;
    const int M = p.B * hw;

    // A-tile load mapping: one pixel row per thread, 8 consecutive channels (one 32-B sector)
    const int am = tid & 127, akh = tid >> 7;
    const int m = m0 + am;
    const bool mvalid = m < M;
    int pb = 0, py = 0, px = 0;
    if (mvalid) { pb = m / hw; int r = m - pb * hw; py = r / p.w; px = r - py * p.w; }

    const int chunks_per_tap = p.cin_pad / BK;
    const int nchunks = p.kh * p.kw * chunks_per_tap;
    const int ph = p.kh >> 1, pw = p.kw >> 1;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 a_reg[2];
    float4 b_reg[NB4];

    auto load_global = [&](int kc) {
        const int tap = kc / chunks_per_tap;
        const int cc = (kc - tap * chunks_per_tap) * BK;
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        const int sy = py + ky - ph, sx = px + kx - pw;
        const bool ok = mvalid && sy >= 0 && sy < p.h && sx >= 0 && sx < p.w;
        const size_t pix = (size_t)(pb * p.h + sy) * p.w + sx;
        const float* base; int cl, climit;
        if (cc < p.c0) { base = p.src0 + pix * p.pitch0; cl = cc + akh * 8; climit = p.c0; }
        else { base = p.src1 + pix * p.pitch1; cl = cc - p.c0 + akh * 8; climit = p.c1; }
        a_reg[0] = (ok && cl < climit) ? ldg4(base + cl) : make_float4(0.f, 0.f, 0.f, 0.f);
        a_reg[1] = (ok && cl + 4 < climit) ? ldg4(base + cl + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* wb = p.wgt + ((size_t)tap * p.cin_pad + cc) * p.cout_pad + n0;
#pragma unroll
        for (int j = 0; j < NB4; ++j) {
            const int idx = tid + j * 256;
            const int kk = idx / (BN / 4), n4 = idx - kk * (BN / 4);
            b_reg[j] = ldg4(wb + (size_t)kk * p.cout_pad + n4 * 4);
        }
    };
    auto store_smem = [&](int buf) {
        float* a = &As[buf][akh * 8][am];
        a[0 * BM] = a_reg[0].x; a[1 * BM] = a_reg[0].y; a[2 * BM] = a_reg[0].z; a[3 * BM] = a_reg[0].w;
        a[4 * BM] = a_reg[1].x; a[5 * BM] = a_reg[1].y; a[6 * BM] = a_reg[1].z; a[7 * BM] = a_reg[1].w;
#pragma unroll
        for (int j = 0; j < NB4; ++j) {
            const int idx = tid + j * 256;
            const int kk = idx / (BN / 4), n4 = idx - kk * (BN / 4);
            *reinterpret_cast<float4*>(&Bs[buf][kk][n4 * 4]) = b_reg[j];
        }
    };

    load_global(0);
    store_smem(0);
    __syncthreads();
    int buf = 0;
    for (int kc = 0; kc < nchunks; ++kc) {
        const bool more = kc + 1 < nchunks;
        if (more) load_global(kc + 1);
        tile_fma<BN>(As[buf], Bs[buf], ty, tx, acc);
        if (more) store_smem(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---------------------------------------------------------------- epilogue
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int g = 0; g < TN / 4; ++g) {
            const int n = n0 + (g == 0 ? tx * 4 : BN / 2 + tx * 4);
            if (n >= p.cout) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][g * 4 + j] + __ldg(p.bias + n + j);
            if (p.epi == EPI_RELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (p.epi == EPI_SCALE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= p.scale;
            } else if (p.epi == EPI_GRU_ZR) {
                // channels [0,128): z gate;  [128,256): r gate -> r*h
                if (n < 128) {
                    float4 z = make_float4(sigmoidf_(v[0]), sigmoidf_(v[1]), sigmoidf_(v[2]), sigmoidf_(v[3]));
                    *reinterpret_cast<float4*>(p.zbuf + (size_t)r * 128 + n) = z;
                } else {
                    const int nn = n - 128;
                    const float4 hh = *reinterpret_cast<const float4*>(p.hbuf + (size_t)r * 128 + nn);
                    float4 o = make_float4(sigmoidf_(v[0]) * hh.x, sigmoidf_(v[1]) * hh.y, sigmoidf_(v[2]) * hh.z,
                                           sigmoidf_(v[3]) * hh.w);
                    *reinterpret_cast<float4*>(p.rhbuf + (size_t)r * 128 + nn) = o;
                }
                continue;
            } else if (p.epi == EPI_GRU_Q) {
                const float4 z = *reinterpret_cast<const float4*>(p.zbuf + (size_t)r * 128 + n);
                float4* hp = reinterpret_cast<float4*>(p.hbuf + (size_t)r * 128 + n);
                const float4 hh = *hp;
                float4 o;
                o.x = (1.f - z.x) * hh.x + z.x * tanhf(v[0]);
                o.y = (1.f - z.y) * hh.y + z.y * tanhf(v[1]);
                o.z = (1.f - z.z) * hh.z + z.z * tanhf(v[2]);
                o.w = (1.f - z.w) * hh.w + z.w * tanhf(v[3]);
                *hp = o;
                continue;
            }
            float* d = p.dst + (size_t)r * p.dst_pitch + n;
            if (n + 3 < p.cout) {
                *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.cout) d[j] = v[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Correlation volume: C[b][p][q] = (sum_d f1[b][d][p] * f2[b][d][q]) / sqrt(D)
// Both operands are "pixel-contiguous" (NCHW feature maps), so the A tile is loaded along m.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) corr_volume_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                            int D, int P, float inv_div_is_div, float* __restrict__ out) {
    constexpr int BN = 128;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, b = blockIdx.z;
    const float* A = f1 + (size_t)b * D * P;
    const float* Bm = f2 + (size_t)b * D * P;
    const bool vec = (P & 3) == 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 a_reg[2], b_reg[2];
    auto ld_row4 = [&](const float* base, int k, int col) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= D) return v;
        const float* r = base + (size_t)k * P + col;
        if (vec && col + 3 < P) return ldg4(r);
        if (col < P) v.x = __ldg(r);
        if (col + 1 < P) v.y = __ldg(r + 1);
        if (col + 2 < P) v.z = __ldg(r + 2);
        if (col + 3 < P) v.w = __ldg(r + 3);
        return v;
    };
    auto load_global = [&](int kc) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = tid + j * 256;
            const int kk = idx >> 5, c4 = idx & 31;
            a_reg[j] = ld_row4(A, kc * BK + kk, m0 + c4 * 4);
            b_reg[j] = ld_row4(Bm, kc * BK + kk, n0 + c4 * 4);
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int idx = tid + j * 256;
            const int kk = idx >> 5, c4 = idx & 31;
            *reinterpret_cast<float4*>(&As[buf][kk][c4 * 4]) = a_reg[j];
            *reinterpret_cast<float4*>(&Bs[buf][kk][c4 * 4]) = b_reg[j];
        }
    };
    const int nchunks = (D + BK - 1) / BK;
    load_global(0);
    store_smem(0);
    __syncthreads();
    int buf = 0;
    for (int kc = 0; kc < nchunks; ++kc) {
        const bool more = kc + 1 < nchunks;
        if (more) load_global(kc + 1);
        tile_fma<BN>(As[buf], Bs[buf], ty, tx, acc);
        if (more) store_smem(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    float* C = out + (size_t)b * P * P;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= P) continue;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int n = n0 + (g == 0 ? tx * 4 : 64 + tx * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < P) C[(size_t)r * P + n + j] = acc[i][g * 4 + j] / inv_div_is_div;
        }
    }
}

}  // namespace

int b2p_launch_conv(const ConvParams& p, cudaStream_t s) {
    const int M = p.B * p.h * p.w;
    const bool bn128 = (p.cout_pad % 128 == 0);
    dim3 grid(ceil_div(M, BM), p.cout_pad / (bn128 ? 128 : 64));
    if (bn128) conv_gemm_kernel<128><<<grid, 256, 0, s>>>(p);
    else conv_gemm_kernel<64><<<grid, 256, 0, s>>>(p);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_corr_volume(const float* f1, const float* f2, int B, int D, int P, float* level0, cudaStream_t s) {
    dim3 grid(ceil_div(P, BM), ceil_div(P, 128), B);
    corr_volume_kernel<<<grid, 256, 0, s>>>(f1, f2, D, P, sqrtf((float)D), level0);
    B2P_LAUNCH_CHECK();
    return 0;
}
