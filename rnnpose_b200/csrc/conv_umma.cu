// Tensor-core implicit-GEMM convolution for the update block: tcgen05.mma (kind::f16) with fp32 accumulators
// in TMEM, operands staged by TMA into 128B-swizzled shared memory, warp-specialised
// (warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2-9 = epilogue).
// Two kernels: conv_umma_kernel (one smem ring, single CTA; correlation volume and small problems) and
// conv_umma2_kernel<NCTA> (separate activation / weight rings, vertical-tap reuse, CTA pairs with
// tcgen05.mma.cta_group::2; the default for the update block, B200POSE_CONV_MODE).  DESIGN.md section 6.
// Diagnostics compiled in (off unless B200POSE_V2_DEBUG is set): bits 1/2/4 drop the TMA loads / MMAs / epilogue for
// timing experiments (results garbage), bit 16 records per-CTA clock counters and a launch timeline
// (b200pose_debug_conv_counters / b200pose_debug_conv_log, read by tools/conv_counters.py).
//
// Precision scheme ("fp16x3"): every fp32 operand x is carried as two fp16 planes hi = fp16(x),
// lo = fp16(x - hi) (22 significant bits together); a product is accumulated as
//   a_lo*b_hi + a_hi*b_lo + a_hi*b_hi     (three MMAs into the same fp32 TMEM accumulator),
// dropping only the 2^-22 lo*lo term.  Measured on the reference itself this keeps the
// final SE(3) within 3e-7 of the fp32 path, whereas single-pass TF32/FP16 sits at 4-9e-5 (too close to the
// 1e-4 parity bar) and BF16 fails it.
//
// GEMM view: D[M = 128 pixels (a 16x8 spatial patch of one sample)][N = n_tile output channels]
//            += A[M][K] * B[N][K]^T,  K = taps x input channels, consumed 64 channels (one 128-byte row) at a
// time.  The A tile of filter tap (ky,kx) is the same 16x8x64 box of the PXC activation tensor shifted by
// (ky-ph, kx-pw): TMA tiled mode zero-fills everything outside the image, which implements the "same"
// padding of reference thirdparty/raft/update.py (nn.Conv2d(padding=k//2)) without an im2col buffer.
#include "common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <unordered_map>

namespace {

constexpr int UM_THREADS = 320;                        // warps: 0 TMA, 1 MMA, 2-9 epilogue (two per TMEM lane quarter)
constexpr int TILE_ROWS = 16, TILE_COLS = 8;          // 128-pixel M tile
constexpr int BKC = 64;                                // channels per K chunk (128 B of fp16)
constexpr int A_TILE_BYTES = 128 * BKC * 2;            // 16 KB
constexpr int TMEM_COLS = 512;                          // two accumulator buffers
constexpr int TMEM_BUF_COLS = 256;

struct UmmaConvParams {
    CUtensorMap a_hi[2], a_lo[2];     // activation segments (PXC fp16 planes), rank 4: (C, w, h, B)
    CUtensorMap b_hi, b_lo;           // weights, rank 3: (Cin_pad, Cout_pad, taps)
    CUtensorMap bs_hi, bs_lo;         // same tensors with an n_sub-row box, for the split tiles of the last round
    int seg0_chunks, chunks_per_tap, kh, kw;
    int B, h, w, tiles_x, tiles_y;
    int n_tile, cout, stages;
    int n_active, chunk_list[8];           // channel chunks (of 64) visited per tap; the others are skipped
    const float* pre; int pre_pitch;        // optional fp32 partial sums added before the activation ([P][pre_pitch])
    int m_tiles, total_tiles, b_batched;   // b_batched: the weight map's 3rd coordinate is the sample index (correlation volume)
    // Work units.  Units [0, full_units) are whole tiles (n_tile output channels).  The tiles of the last, partially
    // filled round of the persistent grid are each cut into `split` units of n_sub channels so that the round keeps
    // (almost) every SM busy for 1/split of a tile time instead of a few SMs for a whole one.
    int full_units, total_units, split, n_sub;
    const float* bias;
    int epi; float scale;
    float* out_f32; int out_f32_pitch;
    __half* out_hi; __half* out_lo; int out_h_pitch;
    float* zbuf; float* hbuf;         // GRU side buffers, fp32 [P][128]
    float* fl_coords1; float* fl_flow; float* fl_dflow;   // EPI_FLOW (flow head conv2, chained launch only)
    // conv_umma2_kernel only (separate activation / weight rings, optional CTA pair):
    int a_taps;                        // vertical taps served by one activation box (kh when the box carries the halo rows, else 1)
    int a_rows;                        // rows of the activation box = 16 + a_taps - 1
    int ring_a, ring_b;                // ring depths
    int m_groups;                      // M tiles / CTAs per cluster, rounded up (a unit = one group x one N tile)
    int side_tiled;                    // z / h / pre-sum side buffers use the tiled layout (b2p_tiled_index)
    int out_tiled;                     // EPI_SCALE output is such a side buffer (the GRU pre-sum GEMMs)
    int dbg_layer;                     // row of g_conv_dbg (layer id + 1)
    int n_reverse;                     // conv_chain_kernel: list the layer's N tiles last-to-first
    int stride;                        // convolution stride (1, or 2 for the encoder's down-sampling layers; gen-1 / gen-2 without tap reuse)
    // conv_chain_kernel, 1 x kw layers with horizontal-tap reuse: the activation box is ordered [x][y][C] (16 rows x 8+kw-1
    // columns, x slow), one box serves all kw taps (tap kx = the same box 16 pixels = 2 swizzle atoms further on), and the
    // accumulator row of a pixel is m = x * 16 + y instead of y * 8 + x
    int xmajor;
    float* hbuf_alt;                   // EPI_GRU_Q: second copy of the hidden state, kept in the OTHER pixel order (see api.cu)
    int debug;                         // timing experiments only (results are garbage): 1 = no TMA loads, 2 = no MMAs, 4 = no epilogue
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 26)) __trap();      // a protocol bug must fail the launch, never hang the GPU
    }
}
// same, for the long waits of the epilogue warps: back off between polls so that the eight waiting warps do not
// compete with the single TMA / MMA issuing threads for issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 = 1 | [32,46) SBO>>4 = 1024/16 | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// the "+r" operands tie the loaded registers to the wait so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// one lane of a converged warp (the compiler keeps the surrounding loop in uniform registers and issues the tcgen05 / TMA
// instructions directly, without the per-thread serialisation loop it emits inside an `if (lane == 0)` region)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred;
}
// Branch-free gate functions of the tensor-core epilogue (ex2.approx + rcp.approx, ~3e-7 relative error).  The IEEE
// expf / division / tanhf sequences carry a slow-path branch per element, which serialises the 32 elements of a column
// group (~130 dependent clocks each) and made the GRU epilogues as long as their main loops.  The exact-fp32 path
// (conv_gemm.cu) keeps the IEEE functions.
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
    const float t = __expf(-2.0f * fabsf(x));               // in (0, 1]
    return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// 256-bit global stores (sm_100: STG.E.256): a thread's 32-byte row segment in one instruction and one L1 wavefront
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// Epilogue of one tile for one warp: TMEM -> registers -> bias / activation / GRU blend -> global (fp32 side outputs and the
// fp16 hi/lo planes the next convolution reads).  `valid` = this thread's pixel lies inside the image.
//
// Every global access here is one pixel row segment per thread (thread = pixel), i.e. 32 different cache lines per warp
// instruction: the L1 takes one wavefront per line, so the epilogue cost is (instructions x 32) wavefronts and the GRU
// epilogues (24 loads + 16 stores per 32-channel group) took as long as their main loops.  Two remedies:
//  * the fp32 side buffers that only the epilogues touch (z gate, hidden state h, GRU pre-sums) use a TILED layout
//    [pixel tile][C/4][128 pixels][4] (b2p_tiled_index): the 32 pixels of a warp are contiguous, 4 lines per instruction;
//  * the PXC rows that must stay pixel-major (TMA operand planes, mask) are written with 256-bit stores.
// COHERENT: the side buffers may have been written by another SM earlier in the SAME launch (conv_chain_kernel): read them
// through L2 (ld.global.cg) instead of the L1 / non-coherent path.
template <bool COHERENT = false>
__device__ __forceinline__ void epilogue_columns(const UmmaConvParams& p, uint32_t tmem_base, int warp, int q, int buf, int n0,
                                                 int n_cnt, bool valid, size_t pix, int tile, int mrow, int mrow_alt = -1) {
    // float4 slot of channel c (multiple of 4) of this thread's pixel in a side buffer with C channels; step between
    // consecutive float4 slots: 1 (PXC) or 128 (tiled)
    const int sstep = p.side_tiled ? 128 : 1;
    auto side4 = [&](const float* base, int C, int c) -> const float4* {
        return reinterpret_cast<const float4*>(base) +
               (p.side_tiled ? ((size_t)tile * (C >> 2) + (c >> 2)) * 128 + mrow : (pix * C + c) >> 2);
    };
    // the two warps of a lane quarter take alternate 32-column groups
    for (int col0 = ((warp - 2) >> 2) * 32; col0 < n_cnt; col0 += 64) {
        uint32_t r[32];
        tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS + col0), r);
        const int nb = n0 + col0;
        const bool live = valid && nb < p.cout;
        // side inputs of the GRU epilogues: all loads of the 32-channel group are issued back to back (one
        // dependent-load round trip per group instead of eight) while the TMEM load is still in flight
        float4 hh[8], zz[8];
        const bool need_h = live && ((p.epi == EPI_GRU_ZR && nb >= 128) || p.epi == EPI_GRU_Q);
        const bool need_z = live && p.epi == EPI_GRU_Q;
        if (need_h) {
            const float4* hp4 = side4(p.hbuf, 128, nb & 127);
#pragma unroll
            for (int j = 0; j < 8; ++j) hh[j] = COHERENT ? __ldcg(hp4 + j * sstep) : hp4[j * sstep];
        }
        if (need_z) {
            const float4* zp4 = side4(p.zbuf, 128, nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) zz[j] = COHERENT ? __ldcg(zp4 + j * sstep) : __ldg(zp4 + j * sstep);
        }
        float4 pre4[8];
        if (live && p.pre) {
            const float4* pp = side4(p.pre, p.pre_pitch, nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) pre4[j] = __ldg(pp + j * sstep);
        }
        tmem_ld_wait(r);
        if (!live) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
            v[j] = __uint_as_float(r[j]) + b4.x; v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + b4.z; v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
        }
        if (p.pre) {                                       // contribution of the iteration-invariant input channels
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[4 * j] += pre4[j].x; v[4 * j + 1] += pre4[j].y; v[4 * j + 2] += pre4[j].z; v[4 * j + 3] += pre4[j].w;
            }
        }
        // channels of this 32-group that exist: bounded by the layer (cout) and by the tile (n_tile need not be a
        // multiple of 32, e.g. 240 for the correlation volume)
        const int lim = min(p.cout - nb, n_cnt - col0);
        if (p.epi == EPI_FLOW) {
            // flow head conv2 (update.py:13-14) fused with coords1 += delta_flow; flow = coords1 - coords0 (CFNet.py:157,166):
            // this thread's pixel, output channels 0 (x) and 1 (y)
            const int px = (int)(pix % (size_t)p.w), py = (int)((pix / (size_t)p.w) % (size_t)p.h);
            if (p.fl_dflow) *reinterpret_cast<float2*>(p.fl_dflow + pix * 2) = make_float2(v[0], v[1]);
            float2 c1 = *reinterpret_cast<const float2*>(p.fl_coords1 + pix * 2);
            c1.x += v[0]; c1.y += v[1];
            *reinterpret_cast<float2*>(p.fl_coords1 + pix * 2) = c1;
            *reinterpret_cast<float2*>(p.fl_flow + pix * 2) = make_float2(c1.x - (float)px, c1.y - (float)py);
            continue;
        }
        if (p.epi == EPI_SCALE) {
            if (p.out_tiled) {                             // GRU pre-sums: a side buffer of out_f32_pitch channels
                float4* d = const_cast<float4*>(side4(p.out_f32, p.out_f32_pitch, nb));
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (4 * j + 3 < lim) d[j * sstep] = make_float4(v[4 * j] * p.scale, v[4 * j + 1] * p.scale, v[4 * j + 2] * p.scale, v[4 * j + 3] * p.scale);
                continue;
            }
            float* d = p.out_f32 + pix * p.out_f32_pitch + nb;
            const bool v8 = (p.out_f32_pitch & 7) == 0;    // 32-byte aligned row segments
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                if (v8 && j + 7 < lim) {
                    uint32_t w8[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) w8[t] = __float_as_uint(v[j + t] * p.scale);
                    st_global_v8(d + j, w8);
                } else {
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (j + t < lim) d[j + t] = v[j + t] * p.scale;
                }
            }
            continue;
        }
        if (p.epi == EPI_GRU_ZR && nb < 128) {            // z gate, kept in fp32 for the blend
            float4* d = const_cast<float4*>(side4(p.zbuf, 128, nb));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                d[j * sstep] = make_float4(sigm(v[4 * j]), sigm(v[4 * j + 1]), sigm(v[4 * j + 2]), sigm(v[4 * j + 3]));
            continue;
        }
        int oc = nb;                                       // output channel of v[0]
        if (p.epi == EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (p.epi == EPI_GRU_ZR) {                  // r gate -> r * h
            oc = nb - 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[4 * j] = sigm(v[4 * j]) * hh[j].x; v[4 * j + 1] = sigm(v[4 * j + 1]) * hh[j].y;
                v[4 * j + 2] = sigm(v[4 * j + 2]) * hh[j].z; v[4 * j + 3] = sigm(v[4 * j + 3]) * hh[j].w;
            }
        } else if (p.epi == EPI_GRU_Q) {                   // h <- (1-z) h + z tanh(q)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[4 * j] = (1.f - zz[j].x) * hh[j].x + zz[j].x * tanh_fast(v[4 * j]);
                v[4 * j + 1] = (1.f - zz[j].y) * hh[j].y + zz[j].y * tanh_fast(v[4 * j + 1]);
                v[4 * j + 2] = (1.f - zz[j].z) * hh[j].z + zz[j].z * tanh_fast(v[4 * j + 2]);
                v[4 * j + 3] = (1.f - zz[j].w) * hh[j].w + zz[j].w * tanh_fast(v[4 * j + 3]);
            }
            float4* hp4 = const_cast<float4*>(side4(p.hbuf, 128, nb));
#pragma unroll
            for (int j = 0; j < 8; ++j) hp4[j * sstep] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (p.hbuf_alt && mrow_alt >= 0) {             // the copy in the other pixel order (tiled buffers only)
                float4* ha = reinterpret_cast<float4*>(p.hbuf_alt) + ((size_t)tile * 32 + (nb >> 2)) * 128 + mrow_alt;
#pragma unroll
                for (int j = 0; j < 8; ++j) ha[j * 128] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        // fp16 hi/lo planes for the next convolution: 16 channels = 32 bytes per store
        __half* dh = p.out_hi + pix * p.out_h_pitch + oc;
        __half* dl = p.out_lo + pix * p.out_h_pitch + oc;
        const int cvalid = lim;
        const bool h16 = (p.out_h_pitch & 15) == 0 && (oc & 15) == 0;
#pragma unroll
        for (int j = 0; j < 32; j += 16) {
            __half hi16[16], lo16[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) b2p_split_half(v[j + t], hi16[t], lo16[t]);
            if (h16 && j + 15 < cvalid) {
                uint32_t wh[8], wl[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    wh[t] = (uint32_t)__half_as_ushort(hi16[2 * t]) | ((uint32_t)__half_as_ushort(hi16[2 * t + 1]) << 16);
                    wl[t] = (uint32_t)__half_as_ushort(lo16[2 * t]) | ((uint32_t)__half_as_ushort(lo16[2 * t + 1]) << 16);
                }
                st_global_v8(dh + j, wh);
                st_global_v8(dl + j, wl);
            } else {
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    if (j + t < cvalid) { dh[j + t] = hi16[t]; dl[j + t] = lo16[t]; }
            }
        }
    }
}

// Timing experiment (debug bit 16): per-CTA clock counters of the last launch.
//   [0] MMA warp: clocks from first to last instruction of its loop   [1] ... spent waiting for full barriers
//   [2] ... waiting for a drained TMEM buffer   [3] producer: waiting for empty slots   [4] epilogue warp 2: waiting for
//   tmem_full   [5] epilogue warp 2: working   [6] tiles of this CTA   [7] stages of this CTA
__device__ unsigned long long g_conv_dbg[12][160][8];      // [layer id + 1][CTA][counter], accumulated over launches
// launch log of CTA 0 (debug bit 16): {globaltimer at entry, after griddepcontrol.wait, at exit, layer row}
__device__ unsigned long long g_conv_log[512][4];
__device__ unsigned int g_conv_log_n;
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, bool timed, unsigned long long& accum) {
    if (!timed) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    accum += (unsigned long long)(clock64() - t0);
}

// ---------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(UM_THREADS, 1) conv_umma_kernel(const __grid_constant__ UmmaConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long t_entry = (p.debug & 16) ? gtime_ns() : 0ull;
    pdl_trigger();                       // the next kernel's blocks may take this SM as soon as this CTA retires
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_tile_bytes = (uint32_t)p.n_tile * 128u;
    const uint32_t stage_bytes = 2u * A_TILE_BYTES + 2u * b_tile_bytes;
    const uint32_t bars = smem_base + (uint32_t)p.stages * stage_bytes;     // full[S], empty[S], tmem_full, tmem_slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    // two TMEM accumulator buffers: the epilogue of tile i overlaps the main loop of tile i+1
    auto tmem_full_bar = [&](int b) { return bars + 16u * p.stages + 8u * b; };
    auto tmem_empty_bar = [&](int b) { return bars + 16u * p.stages + 16u + 8u * b; };
    const uint32_t tmem_slot = bars + 16u * p.stages + 32u;

    // persistent CTA: tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; t = n_idx * m_tiles + m_idx so that the
    // CTAs of one round share the weight tile (L2) and walk over different pixel tiles
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int nk = p.kh * p.kw * p.n_active;
    auto decode = [&](int u, int& bimg, int& y0, int& x0, int& n0, int& n_cnt, int& m_idx) {
        int t = u, sub = 0;
        n_cnt = p.n_tile;
        if (u >= p.full_units) {
            const int v = u - p.full_units;
            t = p.full_units + v / p.split; sub = v - (v / p.split) * p.split;
            n_cnt = p.n_sub;
        }
        const int n_idx = t / p.m_tiles;
        m_idx = t - n_idx * p.m_tiles;
        bimg = m_idx / tiles_per_img;
        const int trem = m_idx - bimg * tiles_per_img;
        y0 = (trem / p.tiles_x) * TILE_ROWS; x0 = (trem % p.tiles_x) * TILE_COLS;
        n0 = n_idx * p.n_tile + sub * p.n_sub;
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 256); }
        mbar_init(tmem_slot + 16u, 1);           // scratch barrier of the commit-cost experiment (debug bit 8)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();                          // barriers and TMEM are set up; from here on the previous kernel's output is read
    const unsigned long long t_dep = (p.debug & 16) ? gtime_ns() : 0ull;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (runs ahead across tile boundaries): warp-uniform loop,
        // one elected lane issues.  Ring slot and phase are carried incrementally and the K loop is nested (tap, channel
        // chunk): no integer divisions of a flat K index on the issue path.
        int s = 0; uint32_t ph = 0;
        const bool timed = (p.debug & 16) != 0;
        unsigned long long w_empty = 0;
        for (int t = blockIdx.x; t < p.total_units; t += gridDim.x) {
            int bimg, y0, x0, n0, n_cnt, m_idx;
            decode(t, bimg, y0, x0, n0, n_cnt, m_idx);
            const bool whole = n_cnt == p.n_tile;
            const CUtensorMap* bh = whole ? &p.b_hi : &p.bs_hi;
            const CUtensorMap* bl = whole ? &p.b_lo : &p.bs_lo;
            const uint32_t tx_bytes = 2u * A_TILE_BYTES + 2u * (uint32_t)n_cnt * 128u;
            int tap = 0;
            for (int ky = 0; ky < p.kh; ++ky) {
                const int ys = y0 * p.stride + ky - (p.kh >> 1);
                for (int kx = 0; kx < p.kw; ++kx, ++tap) {
                    const int xs = x0 * p.stride + kx - (p.kw >> 1);
                    const int b2 = p.b_batched ? bimg : tap;
                    for (int a = 0; a < p.n_active; ++a) {
                        const int cc = p.chunk_list[a];
                        const int seg = cc >= p.seg0_chunks ? 1 : 0;
                        const int c0 = (seg ? cc - p.seg0_chunks : cc) * BKC;
                        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
                        mbar_wait_timed(empty_bar(s), ph ^ 1u, timed, w_empty);
                        if (elect_one()) {
                            if (p.debug & 1) {
                                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_bar(s)) : "memory");
                            } else {
                                mbar_expect_tx(full_bar(s), tx_bytes);
                                tma_load_4d(&p.a_hi[seg], sa, full_bar(s), c0, xs, ys, bimg);
                                tma_load_4d(&p.a_lo[seg], sa + A_TILE_BYTES, full_bar(s), c0, xs, ys, bimg);
                                tma_load_3d(bh, sa + 2 * A_TILE_BYTES, full_bar(s), cc * BKC, n0, b2);
                                tma_load_3d(bl, sa + 2 * A_TILE_BYTES + b_tile_bytes, full_bar(s), cc * BKC, n0, b2);
                            }
                        }
                        __syncwarp();
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
        if (timed && lane == 0) g_conv_dbg[p.dbg_layer][blockIdx.x][3] += w_empty;
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer: the whole warp walks the loop (uniform control flow and
        // uniform-register descriptors), one elected lane issues
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=B=F16 (0), K-major both,
        // N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc_whole = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc_sub = (1u << 4) | ((uint32_t)(p.n_sub >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t tile_iter = 0;
        int s = 0; uint32_t ph = 0;
        const bool timed = (p.debug & 16) != 0;
        unsigned long long w_full = 0, w_tmem = 0, n_stage = 0;
        const long long t_begin = clock64();
        for (int t = blockIdx.x; t < p.total_units; t += gridDim.x, ++tile_iter) {
            const uint32_t idesc = t < p.full_units ? idesc_whole : idesc_sub;
            const int buf = tile_iter & 1;
            const uint32_t acc = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
            mbar_wait_timed(tmem_empty_bar(buf), ((tile_iter >> 1) & 1) ^ 1, timed, w_tmem);     // epilogue has drained this buffer
            tc_fence_after();
            for (int kc = 0; kc < nk; ++kc) {
                mbar_wait_timed(full_bar(s), ph, timed, w_full);
                ++n_stage;
                tc_fence_after();
                const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
                if (elect_one()) {
                    const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + A_TILE_BYTES);
                    const uint64_t b_hi = umma_desc_sw128(sa + 2 * A_TILE_BYTES);
                    const uint64_t b_lo = umma_desc_sw128(sa + 2 * A_TILE_BYTES + b_tile_bytes);
                    if (!(p.debug & 2))
#pragma unroll
                    for (int k = 0; k < BKC / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);   // 16 halves = 32 B = 2 x 16 B along K inside the swizzle atom
                        tc_mma_f16(acc, a_lo + adv, b_hi + adv, idesc, (kc | k) != 0 ? 1u : 0u);
                        tc_mma_f16(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                        tc_mma_f16(acc, a_hi + adv, b_hi + adv, idesc, 1u);
                    }
                    if (p.debug & 8) tc_commit(tmem_slot + 16u);     // experiment: cost of one more commit per stage
                    tc_commit(empty_bar(s));      // frees the smem stage when these MMAs have read it
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) tc_commit(tmem_full_bar(buf));    // accumulator of this tile complete
            __syncwarp();
        }
        if (timed && lane == 0) {
            unsigned long long* d = g_conv_dbg[p.dbg_layer][blockIdx.x];
            d[0] += (unsigned long long)(clock64() - t_begin); d[1] += w_full; d[2] += w_tmem; d[6] += tile_iter; d[7] += n_stage;
        }
    } else {
        // ---------------------------------------------------- epilogue: TMEM -> registers -> global
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int mrow = q * 32 + lane;
        uint32_t tile_iter = 0;
        const bool timed = (p.debug & 16) != 0 && warp == 2;
        unsigned long long w_tfull = 0, busy = 0;
        for (int t = blockIdx.x; t < p.total_units; t += gridDim.x, ++tile_iter) {
        int bimg, y0, x0, n0, n_cnt, m_idx;
        decode(t, bimg, y0, x0, n0, n_cnt, m_idx);
        const int buf = tile_iter & 1;
        const long long e0 = timed ? clock64() : 0;
        const int yy = y0 + (mrow >> 3), xx = x0 + (mrow & 7);
        const bool valid = yy < p.h && xx < p.w;
        const size_t pix = ((size_t)bimg * p.h + yy) * p.w + xx;
        mbar_wait_backoff(tmem_full_bar(buf), (tile_iter >> 1) & 1);
        const long long e1 = timed ? clock64() : 0;
        tc_fence_after();
        if (!(p.debug & 4)) epilogue_columns(p, tmem_base, warp, q, buf, n0, n_cnt, valid, pix, m_idx, mrow);
        if (timed) { w_tfull += (unsigned long long)(e1 - e0); busy += (unsigned long long)(clock64() - e1); }
        // every TMEM read of this tile has completed (tcgen05.wait::ld above): hand the buffer back to the MMA warp
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(buf)) : "memory");
        }   // tile loop
        if (timed && lane == 0) { g_conv_dbg[p.dbg_layer][blockIdx.x][4] += w_tfull; g_conv_dbg[p.dbg_layer][blockIdx.x][5] += busy; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if ((p.debug & 16) && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned i = atomicAdd(&g_conv_log_n, 1u) & 511u;
        g_conv_log[i][0] = t_entry; g_conv_log[i][1] = t_dep; g_conv_log[i][2] = gtime_ns(); g_conv_log[i][3] = (unsigned long long)p.dbg_layer;
    }
}


// ---------------------------------------------------------------------------------------------- kernel, second generation
// The first kernel is bound by the L2 -> shared-memory feed (profiles/r1_final_summary.md: 6.9 GB per update-block pass at
// 8-11 TB/s, the LTS cap is ~12 TB/s), not by the tensor pipe.  This one moves fewer bytes for the same MMAs:
//  * NCTA = 2: a CTA pair (cluster of two SMs of one TPC) computes two M tiles against ONE weight tile with
//    tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128 pixels and only HALF of the weight rows, the MMA reads
//    both halves.  Weight traffic per pixel tile halves.  The leader CTA (rank 0) issues the MMAs; both CTAs run a
//    producer and an epilogue.  Full barriers live in the leader (the peer's TMA completes on them remotely), the
//    MMA-completion commits are multicast to the empty barriers of both CTAs.
//  * vertical-tap reuse: for a kh x kw filter the activation box is loaded once per (kx, channel chunk) with kh-1 halo rows
//    (16+kh-1 rows of 8 pixels = whole 1024-byte swizzle atoms); the A descriptor of vertical tap ky starts ky atoms
//    further down.  Activation traffic of the 3x3 layers drops 9 -> 3.4 boxes, of the 5x1 layers 5 -> 1.25.
//    Activations and weights therefore travel in two rings (one A slot serves a_taps B slots).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in the cluster's CTA `rank` (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
// pair variants: destination in the executing CTA, completion on the (possibly remote) barrier `bar`
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {     // arrives on `bar`'s offset in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int NCTA>
__global__ void __launch_bounds__(UM_THREADS, 1) conv_umma2_kernel(const __grid_constant__ UmmaConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long t_entry = (p.debug & 16) ? gtime_ns() : 0ull;
    pdl_trigger();
    // identical offsets in both CTAs of a pair: the MMA applies one descriptor to the shared memory of both
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
    const uint32_t a_plane = (uint32_t)p.a_rows * 1024u;                 // one fp16 plane of an activation box
    const uint32_t a_slot = 2u * a_plane;
    const uint32_t b_plane = (uint32_t)(p.n_tile / NCTA) * 128u;          // this CTA's share of the weight rows, one plane
    const uint32_t b_slot = 2u * b_plane;
    const uint32_t a_ring = smem_base, b_ring = smem_base + (uint32_t)p.ring_a * a_slot;
    const uint32_t bars = b_ring + (uint32_t)p.ring_b * b_slot;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto empty_a = [&](int s) { return bars + 8u * (p.ring_a + s); };
    const uint32_t bars_b = bars + 16u * p.ring_a;
    auto full_b = [&](int s) { return bars_b + 8u * s; };
    auto empty_b = [&](int s) { return bars_b + 8u * (p.ring_b + s); };
    const uint32_t bars_t = bars_b + 16u * p.ring_b;
    auto tmem_full_bar = [&](int b) { return bars_t + 8u * b; };
    auto tmem_empty_bar = [&](int b) { return bars_t + 16u + 8u * b; };
    const uint32_t tmem_slot = bars_t + 32u;

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    // a unit = (N tile or sub-tile) x (group of NCTA consecutive M tiles); CTA `rank` of the cluster owns M tile group*NCTA+rank.
    // A group that runs past the last M tile repeats it and discards the result (`real` = false).
    auto decode = [&](int u, int& bimg, int& y0, int& x0, int& n0, int& n_cnt, bool& real, int& m_idx) {
        int t = u, sub = 0;
        n_cnt = p.n_tile;
        if (u >= p.full_units) {
            const int v = u - p.full_units;
            t = p.full_units + v / p.split; sub = v - (v / p.split) * p.split;
            n_cnt = p.n_sub;
        }
        const int n_idx = t / p.m_groups;
        m_idx = (t - n_idx * p.m_groups) * NCTA + (int)rank;
        real = m_idx < p.m_tiles;
        if (!real) m_idx = p.m_tiles - 1;
        bimg = m_idx / tiles_per_img;
        const int trem = m_idx - bimg * tiles_per_img;
        y0 = (trem / p.tiles_x) * TILE_ROWS; x0 = (trem % p.tiles_x) * TILE_COLS;
        n0 = n_idx * p.n_tile + sub * p.n_sub;
    };
    const int unit0 = (int)blockIdx.x / NCTA, unit_step = (int)gridDim.x / NCTA;
    const int outer_taps = p.a_taps == p.kh ? 1 : p.kh;       // vertical taps that need an activation box of their own

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.ring_a; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < p.ring_b; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 256 * NCTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();      // the peer's barriers exist before anything signals them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    const unsigned long long t_dep = (p.debug & 16) ? gtime_ns() : 0ull;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (every CTA; the leader arms the full barriers).
        // Warp-uniform loop, one elected lane issues.
        int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
        const uint32_t a_tx = (uint32_t)NCTA * a_slot;
        const bool timed = (p.debug & 16) != 0;
        unsigned long long w_empty = 0;
        for (int u = unit0; u < p.total_units; u += unit_step) {
            int bimg, y0, x0, n0, n_cnt, m_idx; bool real;
            decode(u, bimg, y0, x0, n0, n_cnt, real, m_idx);
            const bool whole = n_cnt == p.n_tile;
            const CUtensorMap* bh = whole ? &p.b_hi : &p.bs_hi;
            const CUtensorMap* bl = whole ? &p.b_lo : &p.bs_lo;
            const int b_rows = n_cnt / NCTA;
            const uint32_t b_tx = (uint32_t)NCTA * 2u * (uint32_t)b_rows * 128u;
            const int nb0 = n0 + (int)rank * b_rows;
            for (int kyo = 0; kyo < outer_taps; ++kyo) {
                const int ys = y0 * p.stride + kyo - (p.kh >> 1);
                for (int kx = 0; kx < p.kw; ++kx) {
                    const int xs = x0 * p.stride + kx - (p.kw >> 1);
                    for (int a = 0; a < p.n_active; ++a) {
                        const int cc = p.chunk_list[a];
                        const int seg = cc >= p.seg0_chunks ? 1 : 0;
                        const int c0 = (seg ? cc - p.seg0_chunks : cc) * BKC;
                        const uint32_t da = a_ring + (uint32_t)sa * a_slot;
                        mbar_wait_timed(empty_a(sa), pha ^ 1u, timed, w_empty);
                        if (elect_one()) {
                            if (p.debug & 1) {
                                if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_a(sa)) : "memory");
                            } else if (NCTA == 2) {
                                const uint32_t fb = mapa_rank(full_a(sa), 0);
                                if (rank == 0) mbar_expect_tx(full_a(sa), a_tx);
                                tma_load_4d_pair(&p.a_hi[seg], da, fb, c0, xs, ys, bimg);
                                tma_load_4d_pair(&p.a_lo[seg], da + a_plane, fb, c0, xs, ys, bimg);
                            } else {
                                mbar_expect_tx(full_a(sa), a_tx);
                                tma_load_4d(&p.a_hi[seg], da, full_a(sa), c0, xs, ys, bimg);
                                tma_load_4d(&p.a_lo[seg], da + a_plane, full_a(sa), c0, xs, ys, bimg);
                            }
                        }
                        __syncwarp();
                        if (++sa == p.ring_a) { sa = 0; pha ^= 1u; }
                        for (int j = 0; j < p.a_taps; ++j) {
                            const int tap = (kyo + j) * p.kw + kx;
                            const uint32_t db = b_ring + (uint32_t)sb * b_slot;
                            mbar_wait_timed(empty_b(sb), phb ^ 1u, timed, w_empty);
                            if (elect_one()) {
                                if (p.debug & 1) {
                                    if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_b(sb)) : "memory");
                                } else if (NCTA == 2) {
                                    const uint32_t fb = mapa_rank(full_b(sb), 0);
                                    if (rank == 0) mbar_expect_tx(full_b(sb), b_tx);
                                    tma_load_3d_pair(bh, db, fb, cc * BKC, nb0, tap);
                                    tma_load_3d_pair(bl, db + b_plane, fb, cc * BKC, nb0, tap);
                                } else {
                                    mbar_expect_tx(full_b(sb), b_tx);
                                    tma_load_3d(bh, db, full_b(sb), cc * BKC, nb0, tap);
                                    tma_load_3d(bl, db + b_plane, full_b(sb), cc * BKC, nb0, tap);
                                }
                            }
                            __syncwarp();
                            if (++sb == p.ring_b) { sb = 0; phb ^= 1u; }
                        }
                    }
                }
            }
        }
        if (timed && lane == 0) g_conv_dbg[p.dbg_layer][blockIdx.x][3] += w_empty;
        // drain: every commit aimed at this CTA's empty barriers has landed before the CTA may exit
        for (int i = 0; i < p.ring_a; ++i) { mbar_wait(empty_a(sa), pha ^ 1u); if (++sa == p.ring_a) { sa = 0; pha ^= 1u; } }
        for (int i = 0; i < p.ring_b; ++i) { mbar_wait(empty_b(sb), phb ^ 1u); if (++sb == p.ring_b) { sb = 0; phb ^= 1u; } }
    } else if (warp == 1) {
        if (rank == 0) {
            // ------------------------------------------------ MMA issuer (leader CTA only): warp-uniform loop, one elected
            // lane issues
            const uint32_t m_field = ((128u * NCTA) >> 4) << 24;
            const uint32_t idesc_whole = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | m_field;
            const uint32_t idesc_sub = (1u << 4) | ((uint32_t)(p.n_sub >> 3) << 17) | m_field;
            const int a_items = outer_taps * p.kw * p.n_active;
            uint32_t tile_iter = 0;
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            const bool timed = (p.debug & 16) != 0;
            unsigned long long w_full = 0, w_tmem = 0, n_stage = 0;
            const long long t_begin = clock64();
            for (int u = unit0; u < p.total_units; u += unit_step, ++tile_iter) {
                const uint32_t idesc = u < p.full_units ? idesc_whole : idesc_sub;
                const int buf = tile_iter & 1;
                const uint32_t acc = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
                mbar_wait_timed(tmem_empty_bar(buf), ((tile_iter >> 1) & 1) ^ 1, timed, w_tmem);     // the epilogues (both CTAs) drained this buffer
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int ai = 0; ai < a_items; ++ai) {
                    mbar_wait_timed(full_a(sa), pha, timed, w_full);
                    const uint32_t da = a_ring + (uint32_t)sa * a_slot;
                    for (int j = 0; j < p.a_taps; ++j) {
                        mbar_wait_timed(full_b(sb), phb, timed, w_full);
                        ++n_stage;
                        tc_fence_after();
                        const uint32_t db = b_ring + (uint32_t)sb * b_slot;
                        if (elect_one()) {
                            const uint64_t a_hi = umma_desc_sw128(da + (uint32_t)j * 1024u);
                            const uint64_t a_lo = umma_desc_sw128(da + a_plane + (uint32_t)j * 1024u);
                            const uint64_t b_hi = umma_desc_sw128(db), b_lo = umma_desc_sw128(db + b_plane);
                            if (!(p.debug & 2))
#pragma unroll
                            for (int k = 0; k < BKC / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                const uint32_t acc_flag = k == 0 ? accumulate : 1u;
                                if (NCTA == 2) {
                                    tc_mma_f16_pair(acc, a_lo + adv, b_hi + adv, idesc, acc_flag);
                                    tc_mma_f16_pair(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                                    tc_mma_f16_pair(acc, a_hi + adv, b_hi + adv, idesc, 1u);
                                } else {
                                    tc_mma_f16(acc, a_lo + adv, b_hi + adv, idesc, acc_flag);
                                    tc_mma_f16(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                                    tc_mma_f16(acc, a_hi + adv, b_hi + adv, idesc, 1u);
                                }
                            }
                            if (NCTA == 2) tc_commit_pair(empty_b(sb)); else tc_commit(empty_b(sb));
                            if (j == p.a_taps - 1) { if (NCTA == 2) tc_commit_pair(empty_a(sa)); else tc_commit(empty_a(sa)); }
                        }
                        __syncwarp();
                        accumulate = 1u;
                        if (++sb == p.ring_b) { sb = 0; phb ^= 1u; }
                    }
                    if (++sa == p.ring_a) { sa = 0; pha ^= 1u; }
                }
                if (elect_one()) { if (NCTA == 2) tc_commit_pair(tmem_full_bar(buf)); else tc_commit(tmem_full_bar(buf)); }
                __syncwarp();
            }
            if (timed && lane == 0) {
                unsigned long long* d = g_conv_dbg[p.dbg_layer][blockIdx.x];
                d[0] += (unsigned long long)(clock64() - t_begin); d[1] += w_full; d[2] += w_tmem; d[6] += tile_iter; d[7] += n_stage;
            }
        }
    } else {
        // ---------------------------------------------------- epilogue (every CTA, its own 128 accumulator lanes)
        const int q = warp & 3;
        const int mrow = q * 32 + lane;
        uint32_t tile_iter = 0;
        const bool timed = (p.debug & 16) != 0 && warp == 2;
        unsigned long long w_tfull = 0, busy = 0;
        for (int u = unit0; u < p.total_units; u += unit_step, ++tile_iter) {
            int bimg, y0, x0, n0, n_cnt, m_idx; bool real;
            decode(u, bimg, y0, x0, n0, n_cnt, real, m_idx);
            const int buf = tile_iter & 1;
            const long long e0 = timed ? clock64() : 0;
            const int yy = y0 + (mrow >> 3), xx = x0 + (mrow & 7);
            const bool valid = real && yy < p.h && xx < p.w;
            const size_t pix = ((size_t)bimg * p.h + yy) * p.w + xx;
            mbar_wait_backoff(tmem_full_bar(buf), (tile_iter >> 1) & 1);
            const long long e1 = timed ? clock64() : 0;
            tc_fence_after();
            if (!(p.debug & 4)) epilogue_columns(p, tmem_base, warp, q, buf, n0, n_cnt, valid, pix, m_idx, mrow);
            if (timed) { w_tfull += (unsigned long long)(e1 - e0); busy += (unsigned long long)(clock64() - e1); }
            tc_fence_before();
            if (NCTA == 2) mbar_arrive_cluster(mapa_rank(tmem_empty_bar(buf), 0));
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(buf)) : "memory");
        }
        if (timed && lane == 0) { g_conv_dbg[p.dbg_layer][blockIdx.x][4] += w_tfull; g_conv_dbg[p.dbg_layer][blockIdx.x][5] += busy; }
    }
    tc_fence_before();
    // nobody leaves while the peer may still read this CTA's shared memory or signal its barriers
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    if ((p.debug & 16) && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned i = atomicAdd(&g_conv_log_n, 1u) & 511u;
        g_conv_log[i][0] = t_entry; g_conv_log[i][1] = t_dep; g_conv_log[i][2] = gtime_ns(); g_conv_log[i][3] = (unsigned long long)p.dbg_layer;
    }
}


// ---------------------------------------------------------------------------------------------- kernel, chained layers
// (conv_mode bit 4; first run on hardware in round 2: profiles/r2a, r2b.)
// The eleven convolutions of an update-block pass in ONE persistent launch of CTA pairs.  Each separate launch spends
// 15-20 us outside its MMA loop (pipeline fill, the last unit's epilogue, the grid-wide griddepcontrol.wait; see
// profiles/r1c_conv_counters_timeline.txt), although a tile of layer L+1 only needs its 3x3 tile neighbourhood of layer L.
// Here the units of all layers form one list in topological order (layer by layer); clusters take them round-robin; the
// epilogue bumps done[layer][tile] once per warp (release), and before the first activation load of a unit the producer warp
// polls the counters of the tile neighbourhood in the unit's source layers (acquire), then fences the async proxy (the TMA
// reads what other SMs wrote with ordinary stores).  The same neighbourhood wait also covers every write-after-read hazard of
// the pass (net_h, rh_h, zbuf are rewritten by later layers): a layer that rewrites tile T waits for exactly the units that
// read T.  Every cluster processes its units in list order and only waits for earlier units, so with all clusters
// co-resident there is no deadlock.  Side buffers written earlier in the same launch are read through L2
// (epilogue_columns<true>).  Shared memory: fixed-size ring slots for the largest layer.
constexpr int CH_MAX_LAYERS = 12;
constexpr int CH_MAX_NSUB = 3;                            // N units per tile of a layer (MASK2: 576 = 3 x 192)
// activation slot: hi + lo planes of the largest box of the launch -- 20 atoms (5x1 with halo rows) or 24 (1x5 with halo
// columns, x-major) of 1024 bytes each; ChainParams::a_slot
constexpr uint32_t CH_B_SLOT = 2u * 128u * 128u;          // hi + lo planes of 128 weight rows per CTA (256-channel tiles)

struct ChainDep {
    int n_src, src[2];                 // source layers
    int n_first[2], n_cnt[2];          // which N units of the source this layer reads (HEADS -> MASK2: the mask half only)
    int halo;                          // 1: wait for the 3x3 tile neighbourhood, 0: the same tile only (1x1 layers)
};
struct ChainParams {
    int n_layers, m_tiles;
    int ring_a, ring_b;                     // ring depths (option chain_rings)
    uint32_t a_slot;                        // bytes of one activation ring slot
    // The unit list is a sequence of SEGMENTS.  A segment is one layer (its units N-major: all pixel tiles of N tile 0, then of
    // N tile 1, ...) or two layers interleaved per pixel-tile pair: for M group m, c0 units of layer A (its N tiles) followed by
    // c1 units of layer B.  Interleaving lets an epilogue-bound layer (MASK2, F1) and an MMA-bound one (flow head, C1) run
    // concurrently in the two halves of the pipeline instead of one after the other.
    int n_seg;
    int seg_start[CH_MAX_LAYERS + 1];      // prefix sums of the segments' unit counts
    int seg_layer[CH_MAX_LAYERS][2];       // layer indices (second = -1 for a plain segment)
    int seg_cnt[CH_MAX_LAYERS][2];         // units per M group of each of the two layers (plain segment: {0, 0})
    int* done;                              // [n_layers][CH_MAX_NSUB][m_tiles] epilogue-warp arrivals, zeroed before the launch
    int* next_unit;                         // the unit queue's head (dynamic scheduling), zeroed before the launch
    int dynamic;                            // 1: clusters take units from the queue as they become free; 0: static round robin
    ChainDep dep[CH_MAX_LAYERS];
    UmmaConvParams L[CH_MAX_LAYERS];
};
static_assert(sizeof(ChainParams) <= 32000, "kernel parameters are limited to 32764 bytes (CUDA >= 12.1, sm_70+)");

// cluster-scope barrier operations for the unit queue (the leader CTA hands unit numbers to its peer through distributed
// shared memory: the peer must see the number the remote store wrote once it has seen the remote arrival)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
constexpr int CH_SCHED = 3;                               // unit-queue slots per cluster (producer runs <= 2 units ahead of the epilogue)

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(UM_THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long t_entry = (cp.L[0].debug & 16) ? gtime_ns() : 0ull;
    pdl_trigger();
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int RA = cp.ring_a, RB = cp.ring_b;
    const uint32_t CH_A_SLOT = cp.a_slot;
    const uint32_t a_ring = smem_base, b_ring = smem_base + (uint32_t)RA * CH_A_SLOT;
    const uint32_t bars = b_ring + (uint32_t)RB * CH_B_SLOT;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto empty_a = [&](int s) { return bars + 8u * (RA + s); };
    const uint32_t bars_b = bars + 16u * RA;
    auto full_b = [&](int s) { return bars_b + 8u * s; };
    auto empty_b = [&](int s) { return bars_b + 8u * (RB + s); };
    const uint32_t bars_t = bars_b + 16u * RB;
    auto tmem_full_bar = [&](int b) { return bars_t + 8u * b; };
    auto tmem_empty_bar = [&](int b) { return bars_t + 16u + 8u * b; };
    const uint32_t tmem_slot = bars_t + 32u;
    // unit queue: sched_full[s] in every CTA (1 arrival: the leader's producer), sched_empty[s] in the leader (18 arrivals: its
    // MMA warp + 8 epilogue warps, the peer's producer + 8 epilogue warps), sched_id[s] in every CTA
    auto sched_full = [&](int s) { return bars_t + 64u + 8u * s; };
    auto sched_empty = [&](int s) { return bars_t + 64u + 8u * (CH_SCHED + s); };
    auto sched_id = [&](int s) { return bars_t + 64u + 16u * CH_SCHED + 4u * s; };

    // unit g of the chain -> layer l (g is monotonic per role, so l only ever advances) and the unit inside the layer;
    // a unit = one N tile x two consecutive M tiles, CTA `rank` owns M tile 2 * group + rank (a missing second tile is
    // replaced by the last tile and its result discarded)
    const int total_units = cp.seg_start[cp.n_seg];
    // unit g -> (layer l, unit u inside the layer in N-major numbering); `sidx` only ever advances
    auto locate = [&](int g, int& sidx, int& l, int& u) {
        while (g >= cp.seg_start[sidx + 1]) ++sidx;
        const int off = g - cp.seg_start[sidx];
        const int c0 = cp.seg_cnt[sidx][0], c1 = cp.seg_cnt[sidx][1];
        if (c1 == 0) { l = cp.seg_layer[sidx][0]; u = off; return; }
        // position j of M group mg holds item (j + mg) mod P: the rotation makes the static round-robin (stride = number of
        // clusters, which shares factors with small P) hand every cluster a mix of both layers
        const int P = c0 + c1, mg = off / P, j = (off - mg * P + mg) % P;
        if (j < c0) { l = cp.seg_layer[sidx][0]; u = j * cp.L[l].m_groups + mg; }
        else { l = cp.seg_layer[sidx][1]; u = (j - c0) * cp.L[l].m_groups + mg; }
    };
    const int unit0 = (int)blockIdx.x >> 1, unit_step = (int)gridDim.x >> 1;
    auto decode = [&](const UmmaConvParams& p, int u, int& bimg, int& y0, int& x0, int& n0, bool& real, int& m_idx) {
        const int n_ord = u / p.m_groups;                       // position of the unit's N tile in the list order
        m_idx = (u - n_ord * p.m_groups) * 2 + (int)rank;
        real = m_idx < p.m_tiles;
        if (!real) m_idx = p.m_tiles - 1;
        const int tiles_per_img = p.tiles_x * p.tiles_y;
        bimg = m_idx / tiles_per_img;
        const int trem = m_idx - bimg * tiles_per_img;
        y0 = (trem / p.tiles_x) * TILE_ROWS; x0 = (trem % p.tiles_x) * TILE_COLS;
        // n_reverse: the layer's N tiles are listed last-to-first (HEADS: the mask half, which MASK2 waits for, goes first)
        n0 = (p.n_reverse ? (p.total_tiles / p.m_groups - 1 - n_ord) : n_ord) * p.n_tile;
    };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < RA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < RB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 512); }
        for (int s = 0; s < CH_SCHED; ++s) { mbar_init(sched_full(s), 1); mbar_init(sched_empty(s), 18); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();
    const unsigned long long t_dep = (cp.L[0].debug & 16) ? gtime_ns() : 0ull;

    // Unit sequence of this cluster.  Static: unit0, unit0 + unit_step, ...  Dynamic: the leader's producer warp takes the
    // next unit of the list from a global counter when it is ready to load it and publishes the number to both CTAs; every
    // other role reads it from its CTA's queue slot.  Units are handed out in list order, and a unit only ever waits for
    // units EARLIER in the list, each of which is already held by a resident cluster: no deadlock.
    int q_slot = 0; uint32_t q_phase = 0; int g_static = unit0;
    auto q_advance = [&]() { if (++q_slot == CH_SCHED) { q_slot = 0; q_phase ^= 1u; } };
    // consumer side (whole warp calls it): next unit or -1
    auto next_unit_consumer = [&]() -> int {
        if (!cp.dynamic) { const int g = g_static; g_static += unit_step; return g < total_units ? g : -1; }
        mbar_wait_acquire_cluster(sched_full(q_slot), q_phase);
        const int g = (int)ld_shared_u32(sched_id(q_slot));
        __syncwarp();
        if (lane == 0) mbar_arrive_release_cluster(mapa_rank(sched_empty(q_slot), 0));
        q_advance();
        return g;
    };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (both CTAs)
        int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
        int l = 0, sidx = 0;
        const bool timed = (cp.L[0].debug & 16) != 0;
        int sched_prefetched = 0;
        if (cp.dynamic && rank == 0 && lane == 0) sched_prefetched = atomicAdd(cp.next_unit, 1);
        while (true) {
            int g;
            if (cp.dynamic && rank == 0) {
                // scheduler: wait until every reader has taken the number that occupied this slot, publish the unit fetched
                // one trip ago, and fetch the following one now (the atomic's round trip hides behind this unit's loads)
                mbar_wait(sched_empty(q_slot), q_phase ^ 1u);
                int gq = 0;
                if (lane == 0) {
                    gq = sched_prefetched;                                   // first trip: fetched before the loop
                    if (gq >= total_units) gq = -1;
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sched_id(q_slot)), "r"((uint32_t)gq) : "memory");
                    st_cluster_u32(mapa_rank(sched_id(q_slot), 1), (uint32_t)gq);
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sched_full(q_slot)) : "memory");
                    mbar_arrive_release_cluster(mapa_rank(sched_full(q_slot), 1));
                    if (gq >= 0) sched_prefetched = atomicAdd(cp.next_unit, 1);
                }
                g = __shfl_sync(0xffffffffu, gq, 0);
                q_advance();
            } else {
                g = next_unit_consumer();
            }
            if (g < 0) break;
            int u_in;
            locate(g, sidx, l, u_in);
            const UmmaConvParams& p = cp.L[l];
            int bimg, y0, x0, n0, m_idx; bool real;
            decode(p, u_in, bimg, y0, x0, n0, real, m_idx);
            // ---- dependencies: lanes 0..8 each watch one tile of the 3x3 neighbourhood (lane 4 = the tile itself)
            const ChainDep& d = cp.dep[l];
            const long long td0 = timed ? clock64() : 0;
            if (d.n_src > 0 && lane < 9 && (d.halo || lane == 4)) {
                const int tiles_per_img = p.tiles_x * p.tiles_y;
                const int trem = m_idx - bimg * tiles_per_img;
                const int ty = trem / p.tiles_x + lane / 3 - 1, tx = trem % p.tiles_x + lane % 3 - 1;
                if (ty >= 0 && ty < p.tiles_y && tx >= 0 && tx < p.tiles_x) {
                    const int t = bimg * tiles_per_img + ty * p.tiles_x + tx;
                    for (int k = 0; k < d.n_src; ++k) {
                        for (int n = d.n_first[k]; n < d.n_first[k] + d.n_cnt[k]; ++n) {
                            const int* c = cp.done + ((size_t)d.src[k] * CH_MAX_NSUB + n) * cp.m_tiles + t;
                            uint32_t spins = 0;
                            while (ld_acquire_gpu(c) < 8) {                // the 8 epilogue warps of the CTA that owned the tile
                                __nanosleep(64);
                                if (++spins > (1u << 24)) __trap();      // a scheduling bug must fail the launch, never hang
                            }
                        }
                    }
                }
            }
            __syncwarp();
            asm volatile("fence.proxy.async.global;" ::: "memory");        // what those stores wrote, as seen by the TMA unit
            if (timed && lane == 0) atomicAdd(&g_conv_dbg[p.dbg_layer][blockIdx.x][7], (unsigned long long)(clock64() - td0));
            unsigned long long w_empty = 0;
            const uint32_t a_plane = (uint32_t)p.a_rows * 1024u;
            const int b_rows = p.n_tile >> 1;
            const uint32_t b_plane = (uint32_t)b_rows * 128u;
            const uint32_t a_tx = 2u * 2u * a_plane, b_tx = 2u * 2u * b_plane;
            const int nb0 = n0 + (int)rank * b_rows;
            const int outer_taps = (p.xmajor || p.a_taps == p.kh) ? 1 : p.kh;
            const int kxn = p.xmajor ? 1 : p.kw;                 // x-major: one box serves every horizontal tap
            for (int kyo = 0; kyo < outer_taps; ++kyo) {
                const int ys = y0 + kyo - (p.kh >> 1);
                for (int kx = 0; kx < kxn; ++kx) {
                    const int xs = x0 + kx - (p.kw >> 1);
                    for (int a = 0; a < p.n_active; ++a) {
                        const int cc = p.chunk_list[a];
                        const int seg = cc >= p.seg0_chunks ? 1 : 0;
                        const int c0 = (seg ? cc - p.seg0_chunks : cc) * BKC;
                        const uint32_t da = a_ring + (uint32_t)sa * CH_A_SLOT;
                        mbar_wait_timed(empty_a(sa), pha ^ 1u, timed, w_empty);
                        if (elect_one()) {
                            const uint32_t fb = mapa_rank(full_a(sa), 0);
                            if (rank == 0) mbar_expect_tx(full_a(sa), a_tx);
                            // tensor-map dimension order: (C, x, y, B), or (C, y, x, B) for the x-major boxes
                            const int d1 = p.xmajor ? ys : xs, d2 = p.xmajor ? xs : ys;
                            tma_load_4d_pair(&p.a_hi[seg], da, fb, c0, d1, d2, bimg);
                            tma_load_4d_pair(&p.a_lo[seg], da + a_plane, fb, c0, d1, d2, bimg);
                        }
                        __syncwarp();
                        if (++sa == RA) { sa = 0; pha ^= 1u; }
                        for (int j = 0; j < p.a_taps; ++j) {
                            const int tap = p.xmajor ? j : (kyo + j) * p.kw + kx;
                            const uint32_t db = b_ring + (uint32_t)sb * CH_B_SLOT;
                            mbar_wait_timed(empty_b(sb), phb ^ 1u, timed, w_empty);
                            if (elect_one()) {
                                const uint32_t fb = mapa_rank(full_b(sb), 0);
                                if (rank == 0) mbar_expect_tx(full_b(sb), b_tx);
                                tma_load_3d_pair(&p.b_hi, db, fb, cc * BKC, nb0, tap);
                                tma_load_3d_pair(&p.b_lo, db + b_plane, fb, cc * BKC, nb0, tap);
                            }
                            __syncwarp();
                            if (++sb == RB) { sb = 0; phb ^= 1u; }
                        }
                    }
                }
            }
            if (timed && lane == 0) atomicAdd(&g_conv_dbg[p.dbg_layer][blockIdx.x][3], w_empty);
        }
        // drain: every commit aimed at this CTA's empty barriers has landed before the CTA may exit
        for (int i = 0; i < RA; ++i) { mbar_wait(empty_a(sa), pha ^ 1u); if (++sa == RA) { sa = 0; pha ^= 1u; } }
        for (int i = 0; i < RB; ++i) { mbar_wait(empty_b(sb), phb ^ 1u); if (++sb == RB) { sb = 0; phb ^= 1u; } }
    } else if (warp == 1) {
        if (rank == 0) {
            // ------------------------------------------------ MMA issuer (leader CTA)
            uint32_t tile_iter = 0;
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            int l = 0, sidx = 0;
            const bool timed = (cp.L[0].debug & 16) != 0;
            for (;; ++tile_iter) {
                const int g = next_unit_consumer();
                if (g < 0) break;
                int u_in;
                locate(g, sidx, l, u_in);
                unsigned long long w_full = 0, w_tmem = 0;
                const long long t_begin = timed ? clock64() : 0;
                const UmmaConvParams& p = cp.L[l];
                const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((256u >> 4) << 24);
                const uint32_t a_plane = (uint32_t)p.a_rows * 1024u, b_plane = (uint32_t)(p.n_tile >> 1) * 128u;
                const int outer_taps = (p.xmajor || p.a_taps == p.kh) ? 1 : p.kh;
                const int a_items = outer_taps * (p.xmajor ? 1 : p.kw) * p.n_active;
                const uint32_t tap_step = p.xmajor ? 2048u : 1024u;        // bytes between the A views of consecutive taps
                const int buf = tile_iter & 1;
                const uint32_t acc = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
                mbar_wait_timed(tmem_empty_bar(buf), ((tile_iter >> 1) & 1) ^ 1, timed, w_tmem);
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int ai = 0; ai < a_items; ++ai) {
                    mbar_wait_timed(full_a(sa), pha, timed, w_full);
                    const uint32_t da = a_ring + (uint32_t)sa * CH_A_SLOT;
                    for (int j = 0; j < p.a_taps; ++j) {
                        mbar_wait_timed(full_b(sb), phb, timed, w_full);
                        tc_fence_after();
                        const uint32_t db = b_ring + (uint32_t)sb * CH_B_SLOT;
                        if (elect_one()) {
                            const uint64_t a_hi = umma_desc_sw128(da + (uint32_t)j * tap_step);
                            const uint64_t a_lo = umma_desc_sw128(da + a_plane + (uint32_t)j * tap_step);
                            const uint64_t b_hi = umma_desc_sw128(db), b_lo = umma_desc_sw128(db + b_plane);
#pragma unroll
                            for (int k = 0; k < BKC / 16; ++k) {
                                const uint64_t adv = (uint64_t)(k * 2);
                                tc_mma_f16_pair(acc, a_lo + adv, b_hi + adv, idesc, k == 0 ? accumulate : 1u);
                                tc_mma_f16_pair(acc, a_hi + adv, b_lo + adv, idesc, 1u);
                                tc_mma_f16_pair(acc, a_hi + adv, b_hi + adv, idesc, 1u);
                            }
                            tc_commit_pair(empty_b(sb));
                            if (j == p.a_taps - 1) tc_commit_pair(empty_a(sa));
                        }
                        __syncwarp();
                        accumulate = 1u;
                        if (++sb == RB) { sb = 0; phb ^= 1u; }
                    }
                    if (++sa == RA) { sa = 0; pha ^= 1u; }
                }
                if (elect_one()) tc_commit_pair(tmem_full_bar(buf));
                __syncwarp();
                if (timed && lane == 0) {
                    unsigned long long* dd = g_conv_dbg[p.dbg_layer][blockIdx.x];
                    atomicAdd(&dd[0], (unsigned long long)(clock64() - t_begin)); atomicAdd(&dd[1], w_full); atomicAdd(&dd[2], w_tmem);
                    atomicAdd(&dd[6], 1ull);
                }
            }
        }
    } else {
        // ---------------------------------------------------- epilogue (both CTAs), then the completion signal of the tile
        const int q = warp & 3;
        const int mrow = q * 32 + lane;
        uint32_t tile_iter = 0;
        int l = 0, sidx = 0;
        const bool timed = (cp.L[0].debug & 16) != 0 && warp == 2;
        for (;; ++tile_iter) {
            const int g = next_unit_consumer();
            if (g < 0) break;
            int u_in;
            locate(g, sidx, l, u_in);
            const UmmaConvParams& p = cp.L[l];
            int bimg, y0, x0, n0, m_idx; bool real;
            decode(p, u_in, bimg, y0, x0, n0, real, m_idx);
            const int buf = tile_iter & 1;
            // accumulator row -> pixel of the tile: y * 8 + x, or x * 16 + y for the x-major layers
            const int ty = p.xmajor ? (mrow & 15) : (mrow >> 3), tx = p.xmajor ? (mrow >> 4) : (mrow & 7);
            const int mrow_alt = p.xmajor ? (ty * 8 + tx) : (tx * 16 + ty);       // the same pixel's row in the other order
            const int yy = y0 + ty, xx = x0 + tx;
            const bool valid = real && yy < p.h && xx < p.w;
            const size_t pix = ((size_t)bimg * p.h + yy) * p.w + xx;
            const long long e0 = timed ? clock64() : 0;
            mbar_wait_backoff(tmem_full_bar(buf), (tile_iter >> 1) & 1);
            const long long e1 = timed ? clock64() : 0;
            tc_fence_after();
            epilogue_columns<true>(p, tmem_base, warp, q, buf, n0, p.n_tile, valid, pix, m_idx, mrow, mrow_alt);
            tc_fence_before();
            mbar_arrive_cluster(mapa_rank(tmem_empty_bar(buf), 0));
            if (timed && lane == 0) {
                atomicAdd(&g_conv_dbg[p.dbg_layer][blockIdx.x][4], (unsigned long long)(e1 - e0));
                atomicAdd(&g_conv_dbg[p.dbg_layer][blockIdx.x][5], (unsigned long long)(clock64() - e1));
            }
            // publish this warp's share of the tile: its stores become visible device-wide (and to the async proxy of the
            // SMs that will TMA-load them) before the counter moves
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncwarp();
            if (lane == 0 && real)
                asm volatile("red.release.gpu.global.add.s32 [%0], 1;"
                             ::"l"(cp.done + ((size_t)l * CH_MAX_NSUB + n0 / p.n_tile) * cp.m_tiles + m_idx) : "memory");
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if ((cp.L[0].debug & 16) && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned i = atomicAdd(&g_conv_log_n, 1u) & 511u;
        g_conv_log[i][0] = t_entry; g_conv_log[i][1] = t_dep; g_conv_log[i][2] = gtime_ns(); g_conv_log[i][3] = 0ull;
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// Encoded tensor maps depend only on (base, extents, pitches, box), never on the data: keep them in a small cache so
// that the ~90 launches of one refine call do not re-encode ~500 descriptors (matters for small batches, which are
// bound by host launch cost).
struct MapKey {
    const void* base; int d[6];
    bool operator==(const MapKey& o) const { return base == o.base && !memcmp(d, o.d, sizeof(d)); }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.base);
        for (int i = 0; i < 6; ++i) h = h * 1000003u ^ (size_t)(unsigned)k.d[i];
        return h;
    }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

template <class Make>
int cached_map(CUtensorMap* m, const MapKey& key, Make make) {
    {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) { *m = it->second; return 0; }
    }
    int rc = make(m);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_map_mutex);
    if (g_map_cache.size() > 8192) g_map_cache.clear();
    g_map_cache.emplace(key, *m);
    return 0;
}

// h, w: dimensions of the INPUT map; stride > 1: the box covers stride x as many input pixels per dimension and TMA keeps
// every stride-th one (elementStrides), so the tile in shared memory is always box_rows x 8 pixels.
int make_act_map(CUtensorMap* m, const __half* base, int C, int pitch, int B, int h, int w, int box_rows = TILE_ROWS, int stride = 1) {
    return cached_map(m, MapKey{base, {C, pitch, B, h, w, -box_rows - 1000 * (stride - 1)}}, [&](CUtensorMap* out) -> int {
        EncodeTiledFn enc = get_encode();
        if (!enc) return (int)cudaErrorNotSupported;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)w * pitch * 2, (cuuint64_t)h * w * pitch * 2};
        cuuint32_t box[4] = {BKC, (cuuint32_t)(TILE_COLS * stride), (cuuint32_t)(box_rows * stride), 1};
        cuuint32_t est[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
    });
}

// x-major activation box for the 1 x kw layers of the chained launch: dimensions ordered (C, y, x, B) so that shared memory
// receives [x][y][C]; box = 16 rows x box_cols columns.
int make_act_map_x(CUtensorMap* m, const __half* base, int C, int pitch, int B, int h, int w, int box_cols) {
    return cached_map(m, MapKey{base, {C, pitch, B, h, w, -5000 - box_cols}}, [&](CUtensorMap* out) -> int {
        EncodeTiledFn enc = get_encode();
        if (!enc) return (int)cudaErrorNotSupported;
        cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)h, (cuuint64_t)w, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)w * pitch * 2, (cuuint64_t)pitch * 2, (cuuint64_t)h * w * pitch * 2};
        cuuint32_t box[4] = {BKC, TILE_ROWS, (cuuint32_t)box_cols, 1};
        cuuint32_t est[4] = {1, 1, 1, 1};
        CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
    });
}

int make_wgt_map(CUtensorMap* m, const __half* base, int cin_pad, int cout_pad, int taps, int n_tile) {
    return cached_map(m, MapKey{base, {cin_pad, cout_pad, taps, n_tile, -2, -2}}, [&](CUtensorMap* out) -> int {
        EncodeTiledFn enc = get_encode();
        if (!enc) return (int)cudaErrorNotSupported;
        cuuint64_t gdim[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)taps};
        cuuint64_t gstr[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cout_pad * cin_pad * 2};
        cuuint32_t box[3] = {BKC, (cuuint32_t)n_tile, 1};
        cuuint32_t est[3] = {1, 1, 1};
        CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
    });
}

// SM count and the shared-memory opt-in, once per device
int device_sms() {
    static int sms_of[64];
    static bool done[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!done[dev]) {
        std::lock_guard<std::mutex> lk(g_map_mutex);
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(conv_umma2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(conv_umma2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
        if (cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return -1;
        sms_of[dev] = sms; done[dev] = true;
    }
    return sms_of[dev];
}

// Kernel selection: option "conv_mode" (options.cu; initial value from B200POSE_CONV_MODE, read once): bit 0 = CTA pairs
// (tcgen05.mma.cta_group::2), bit 1 = vertical-tap reuse of the activation box; 0 = first-generation kernel.
// "tail_min_n": smallest channel count of a split unit (0 disables splitting).
int conv_mode() { return b2p_options().conv_mode; }

template <typename... KArgs>
cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), int grid, int cluster, size_t smem, cudaStream_t s, const UmmaConvParams& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(UM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 1 ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

// Second-generation launch (conv_umma2_kernel).  ncta = 2: units are CTA pairs.
int launch_conv_umma2(const UmmaConvArgs& a, UmmaConvParams& p, int ncta, bool reuse_v, int sms, cudaStream_t s) {
    int rc;
    const int taps = a.kh * a.kw;
    // rings: one activation slot serves a_taps weight slots; without reuse the two rings advance together
    const int b_slot = 2 * (a.n_tile / ncta) * 128, budget = 222 * 1024;
    const int st = a.stride > 1 ? a.stride : 1;
    const int in_h = a.in_h > 0 ? a.in_h : a.h, in_w = a.in_w > 0 ? a.in_w : a.w;
    p.stride = st;
    if (st > 1) reuse_v = false;                       // the halo-box trick needs consecutive input rows
    p.a_taps = reuse_v ? a.kh : 1;
    p.a_rows = TILE_ROWS + p.a_taps - 1;
    p.ring_a = 2;
    p.ring_b = (budget - p.ring_a * 2 * p.a_rows * 1024) / b_slot;
    if (p.a_taps > 1 && p.ring_b < 2) { p.a_taps = 1; p.a_rows = TILE_ROWS; }     // the halo box does not fit: plain boxes
    if (p.a_taps == 1) {
        int st = budget / (2 * A_TILE_BYTES + b_slot);
        p.ring_a = p.ring_b = st > 6 ? 6 : st;
    } else if (p.ring_b > 6) {
        p.ring_b = 6;
    }
    if (p.ring_a < 2 || p.ring_b < 2) return -1;
    for (int g = 0; g < 2; ++g) {
        if (a.seg_c[g] == 0) { p.a_hi[g] = p.a_hi[0]; p.a_lo[g] = p.a_lo[0]; continue; }
        if ((rc = make_act_map(&p.a_hi[g], a.seg_hi[g], a.seg_c[g], a.seg_pitch[g], a.B, in_h, in_w, p.a_rows, st))) return rc;
        if ((rc = make_act_map(&p.a_lo[g], a.seg_lo[g], a.seg_c[g], a.seg_pitch[g], a.B, in_h, in_w, p.a_rows, st))) return rc;
    }
    if ((rc = make_wgt_map(&p.b_hi, a.w_hi, a.cin_pad, a.cout_pad, taps, a.n_tile / ncta))) return rc;
    if ((rc = make_wgt_map(&p.b_lo, a.w_lo, a.cin_pad, a.cout_pad, taps, a.n_tile / ncta))) return rc;
    const size_t smem = (size_t)p.ring_a * (2 * p.a_rows * 1024) + (size_t)p.ring_b * (2 * (a.n_tile / ncta) * 128) + 1024 +
                        16 * (p.ring_a + p.ring_b) + 64;
    p.debug = b2p_options().conv_debug;
    p.m_groups = ceil_div(p.m_tiles, ncta);
    p.total_tiles = p.m_groups * (a.cout_pad / a.n_tile);
    const int slots = sms / ncta;                                    // clusters resident at once
    int nclusters = p.total_tiles < slots ? p.total_tiles : slots;
    p.full_units = p.total_tiles; p.total_units = p.total_tiles; p.split = 1; p.n_sub = a.n_tile;
    const int tail = p.total_tiles < slots ? p.total_tiles : p.total_tiles % slots;
    const int tail_min_n = b2p_options().tail_min_n;
    if (tail && tail_min_n > 0) {
        for (int sp = a.n_tile / tail_min_n; sp >= 2; --sp) {
            if (a.n_tile % sp || (a.n_tile / sp) % 32 || tail * sp > slots) continue;
            p.split = sp; p.n_sub = a.n_tile / sp;
            p.full_units = p.total_tiles - tail; p.total_units = p.full_units + tail * sp;
            if (p.total_units < slots) nclusters = p.total_units;
            break;
        }
    }
    p.bs_hi = p.b_hi; p.bs_lo = p.b_lo;
    if (p.split > 1) {
        if ((rc = make_wgt_map(&p.bs_hi, a.w_hi, a.cin_pad, a.cout_pad, taps, p.n_sub / ncta))) return rc;
        if ((rc = make_wgt_map(&p.bs_lo, a.w_lo, a.cin_pad, a.cout_pad, taps, p.n_sub / ncta))) return rc;
    }
    if (ncta == 2) B2P_CUDA(launch_pdl_cluster(conv_umma2_kernel<2>, nclusters * 2, 2, smem, s, p));
    else B2P_CUDA(launch_pdl_cluster(conv_umma2_kernel<1>, nclusters, 1, smem, s, p));
    B2P_LAUNCH_CHECK();
    return 0;
}


// One layer of a chained launch: the same planning as launch_conv_umma2 for a CTA pair with vertical-tap reuse, without
// rings (fixed) and without tail splitting (the next layer's units fill the tail).
int fill_chain_layer(const UmmaConvArgs& a, UmmaConvParams& p) {
    memset(&p, 0, sizeof(p));
    int rc;
    const int taps = a.kh * a.kw;
    const int n_tile = (a.n_tile == 128 && a.cout_pad % 256 == 0) ? 256 : a.n_tile;
    if (a.b_batched || n_tile % 32 || n_tile / 2 > 128 || a.kh > 5 || a.cout_pad % n_tile) return -1;
    // horizontal-tap reuse for the 1 x kw layers (option chain_xmajor): x-major boxes, see UmmaConvParams::xmajor
    p.xmajor = (b2p_options().chain_xmajor != 0 && a.kh == 1 && a.kw > 1 && a.kw <= 5) ? 1 : 0;
    if (p.xmajor) {
        const int box_cols = TILE_COLS + a.kw - 1;
        p.a_taps = a.kw;
        p.a_rows = 2 * box_cols;                       // 1024-byte atoms of one plane: box_cols columns x 16 rows x 128 B
        for (int g = 0; g < 2; ++g) {
            if (a.seg_c[g] == 0) { p.a_hi[g] = p.a_hi[0]; p.a_lo[g] = p.a_lo[0]; continue; }
            if ((rc = make_act_map_x(&p.a_hi[g], a.seg_hi[g], a.seg_c[g], a.seg_pitch[g], a.B, a.h, a.w, box_cols))) return rc;
            if ((rc = make_act_map_x(&p.a_lo[g], a.seg_lo[g], a.seg_c[g], a.seg_pitch[g], a.B, a.h, a.w, box_cols))) return rc;
        }
    } else {
        p.a_taps = a.kh;
        p.a_rows = TILE_ROWS + a.kh - 1;
        for (int g = 0; g < 2; ++g) {
            if (a.seg_c[g] == 0) { p.a_hi[g] = p.a_hi[0]; p.a_lo[g] = p.a_lo[0]; continue; }
            if ((rc = make_act_map(&p.a_hi[g], a.seg_hi[g], a.seg_c[g], a.seg_pitch[g], a.B, a.h, a.w, p.a_rows))) return rc;
            if ((rc = make_act_map(&p.a_lo[g], a.seg_lo[g], a.seg_c[g], a.seg_pitch[g], a.B, a.h, a.w, p.a_rows))) return rc;
        }
    }
    if ((rc = make_wgt_map(&p.b_hi, a.w_hi, a.cin_pad, a.cout_pad, taps, n_tile / 2))) return rc;
    if ((rc = make_wgt_map(&p.b_lo, a.w_lo, a.cin_pad, a.cout_pad, taps, n_tile / 2))) return rc;
    p.bs_hi = p.b_hi; p.bs_lo = p.b_lo;
    p.seg0_chunks = (a.seg_c[0] + BKC - 1) / BKC;
    p.chunks_per_tap = a.cin_pad / BKC;
    for (int cc = 0; cc < p.chunks_per_tap && cc < 8; ++cc)
        if (a.chunk_mask == 0 || ((a.chunk_mask >> cc) & 1u)) p.chunk_list[p.n_active++] = cc;
    p.pre = a.pre; p.pre_pitch = a.pre_pitch;
    p.kh = a.kh; p.kw = a.kw; p.B = a.B; p.h = a.h; p.w = a.w;
    p.tiles_x = ceil_div(a.w, TILE_COLS); p.tiles_y = ceil_div(a.h, TILE_ROWS);
    p.n_tile = n_tile; p.cout = a.cout;
    p.bias = a.bias; p.epi = a.epi; p.scale = a.scale;
    p.out_f32 = a.out_f32; p.out_f32_pitch = a.out_f32_pitch;
    p.out_hi = a.out_hi; p.out_lo = a.out_lo; p.out_h_pitch = a.out_h_pitch;
    p.zbuf = a.zbuf; p.hbuf = a.hbuf;
    // (two hidden-state copies, one per pixel order: an x-major layer reads / writes the x-major copy as its own)
    if (p.xmajor && a.hbuf_x) { p.hbuf = a.hbuf_x; p.hbuf_alt = a.hbuf; } else { p.hbuf_alt = a.hbuf_x; }
    p.fl_coords1 = a.fl_coords1; p.fl_flow = a.fl_flow; p.fl_dflow = a.fl_dflow;
    p.side_tiled = a.side_tiled; p.out_tiled = a.out_tiled;
    p.stride = 1;
    p.debug = b2p_options().conv_debug & 16;             // clock counters only; the drop-a-stage experiments are gen-2 only
    p.dbg_layer = a.layer_id >= 0 && a.layer_id < 11 ? a.layer_id + 1 : 0;
    p.m_tiles = a.B * p.tiles_x * p.tiles_y;
    p.m_groups = ceil_div(p.m_tiles, 2);
    p.total_tiles = p.m_groups * (a.cout_pad / n_tile);
    p.full_units = p.total_units = p.total_tiles; p.split = 1; p.n_sub = n_tile;
    return 0;
}

}  // namespace

int b2p_launch_conv_umma(const UmmaConvArgs& a, cudaStream_t s) {
    UmmaConvParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    const int st = a.stride > 1 ? a.stride : 1;
    const int in_h = a.in_h > 0 ? a.in_h : a.h, in_w = a.in_w > 0 ? a.in_w : a.w;
    for (int g = 0; g < 2; ++g) {
        if (a.seg_c[g] == 0) { p.a_hi[g] = p.a_hi[0]; p.a_lo[g] = p.a_lo[0]; continue; }
        if ((rc = make_act_map(&p.a_hi[g], a.seg_hi[g], a.seg_c[g], a.seg_pitch[g], a.B, in_h, in_w, TILE_ROWS, st))) return rc;
        if ((rc = make_act_map(&p.a_lo[g], a.seg_lo[g], a.seg_c[g], a.seg_pitch[g], a.B, in_h, in_w, TILE_ROWS, st))) return rc;
    }
    p.stride = st;
    const int taps = a.b_batched ? a.b_batched : a.kh * a.kw;     // 3rd weight-map dimension: tap, or sample
    if ((rc = make_wgt_map(&p.b_hi, a.w_hi, a.cin_pad, a.cout_pad, taps, a.n_tile))) return rc;
    if ((rc = make_wgt_map(&p.b_lo, a.w_lo, a.cin_pad, a.cout_pad, taps, a.n_tile))) return rc;
    p.seg0_chunks = (a.seg_c[0] + BKC - 1) / BKC;
    p.chunks_per_tap = a.cin_pad / BKC;
    p.n_active = 0;
    for (int cc = 0; cc < p.chunks_per_tap && cc < 8; ++cc)
        if (a.chunk_mask == 0 || ((a.chunk_mask >> cc) & 1u)) p.chunk_list[p.n_active++] = cc;
    p.pre = a.pre; p.pre_pitch = a.pre_pitch;
    p.kh = a.kh; p.kw = a.kw; p.B = a.B; p.h = a.h; p.w = a.w;
    p.tiles_x = ceil_div(a.w, TILE_COLS); p.tiles_y = ceil_div(a.h, TILE_ROWS);
    p.n_tile = a.n_tile; p.cout = a.cout;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * a.n_tile * 128;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 6) p.stages = 6;
    p.bias = a.bias; p.epi = a.epi; p.scale = a.scale;
    p.out_f32 = a.out_f32; p.out_f32_pitch = a.out_f32_pitch;
    p.out_hi = a.out_hi; p.out_lo = a.out_lo; p.out_h_pitch = a.out_h_pitch;
    p.zbuf = a.zbuf; p.hbuf = a.hbuf;
    p.side_tiled = a.side_tiled; p.out_tiled = a.out_tiled;
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
    const int sms = device_sms();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    p.m_tiles = a.B * p.tiles_x * p.tiles_y;
    p.total_tiles = p.m_tiles * (a.cout_pad / a.n_tile);
    p.b_batched = a.b_batched;
    p.debug = a.b_batched ? 0 : b2p_options().conv_debug;
    p.dbg_layer = a.layer_id >= 0 && a.layer_id < 11 ? a.layer_id + 1 : 0;
    // second-generation kernel (see conv_mode): not for the batched-weight volume GEMM (a pair of M tiles may straddle two
    // samples); CTA pairs only when the problem fills the machine (they halve the number of schedulable units)
    int mode = a.b_batched ? 0 : conv_mode();
    if (mode) {
        const bool pair = (mode & 1) && p.total_tiles >= sms && (a.n_tile % 32) == 0;
        const bool reuse_v = (mode & 2) && a.kh > 1;
        if (pair && a.n_tile == 128 && a.cout_pad % 256 == 0) {
            // A pair holds a 256-row weight tile in the shared memory of two SMs (128 rows each): one pass over the
            // activations instead of two, and M=256 x N=256 MMAs read half as many operand bytes per SM and flop.
            UmmaConvArgs a2 = a;
            a2.n_tile = 256;
            p.n_tile = 256;
            p.total_tiles = p.m_tiles * (a.cout_pad / 256);
            return launch_conv_umma2(a2, p, 2, reuse_v, sms, s);
        }
        if (pair || reuse_v || (mode & 4)) return launch_conv_umma2(a, p, pair ? 2 : 1, reuse_v, sms, s);
    }
    int grid = p.total_tiles < sms ? p.total_tiles : sms;            // persistent: one CTA per SM
    // split the tiles of the last partial round (see UmmaConvParams): the largest split whose units still fit one round.
    // A problem with fewer tiles than SMs (small batches) is one partial round: all of its tiles are split.
    p.full_units = p.total_tiles; p.total_units = p.total_tiles; p.split = 1; p.n_sub = a.n_tile;
    const int tail = p.total_tiles < sms ? p.total_tiles : p.total_tiles % sms;
    const int tail_min_n = b2p_options().tail_min_n;
    if (tail && tail_min_n > 0) {
        for (int sp = a.n_tile / tail_min_n; sp >= 2; --sp) {
            if (a.n_tile % sp || (a.n_tile / sp) % 32 || tail * sp > sms) continue;
            p.split = sp; p.n_sub = a.n_tile / sp;
            p.full_units = p.total_tiles - tail; p.total_units = p.full_units + tail * sp;
            if (p.total_units < sms) grid = p.total_units;
            break;
        }
    }
    p.bs_hi = p.b_hi; p.bs_lo = p.b_lo;
    if (p.split > 1) {
        if ((rc = make_wgt_map(&p.bs_hi, a.w_hi, a.cin_pad, a.cout_pad, taps, p.n_sub))) return rc;
        if ((rc = make_wgt_map(&p.bs_lo, a.w_lo, a.cin_pad, a.cout_pad, taps, p.n_sub))) return rc;
    }
    B2P_CUDA(b2p_launch_pdl(conv_umma_kernel, dim3(grid), dim3(UM_THREADS), smem, s, p));
    B2P_LAUNCH_CHECK();
    return 0;
}

// Experiment hook (not part of include/b200pose.h): per-CTA clock counters of the last gen-1 launch made with
// B200POSE_V2_DEBUG bit 16.  host_out: 12 x 160 x 8 unsigned 64-bit values.
extern "C" int b200pose_debug_conv_counters(unsigned long long* host_out, int reset) {
    B2P_CUDA(cudaDeviceSynchronize());
    if (host_out) B2P_CUDA(cudaMemcpyFromSymbol(host_out, g_conv_dbg, sizeof(unsigned long long) * 12 * 160 * 8));
    if (reset) {
        void* d = nullptr;
        B2P_CUDA(cudaGetSymbolAddress(&d, g_conv_dbg));
        B2P_CUDA(cudaMemset(d, 0, sizeof(unsigned long long) * 12 * 160 * 8));
        B2P_CUDA(cudaGetSymbolAddress(&d, g_conv_log_n));
        B2P_CUDA(cudaMemset(d, 0, sizeof(unsigned int)));
    }
    return 0;
}
// launch log of CTA 0: host_out 512 x 4 values {entry ns, dependency-resolved ns, exit ns, layer row}; returns the count
extern "C" int b200pose_debug_conv_log(unsigned long long* host_out) {
    B2P_CUDA(cudaDeviceSynchronize());
    unsigned int n = 0;
    B2P_CUDA(cudaMemcpyFromSymbol(&n, g_conv_log_n, sizeof(n)));
    B2P_CUDA(cudaMemcpyFromSymbol(host_out, g_conv_log, sizeof(unsigned long long) * 512 * 4));
    return (int)(n > 512 ? 512 : n);
}

// Chained launch (conv_chain_kernel): n layers in list order; deps[l] names the layers whose output layer l reads (and which
// of their N units), halo = 0 for 1x1 layers; n_reverse[l]: list layer l's N tiles last-to-first.  done_ws:
// b2p_conv_chain_done_ints() ints of device scratch.
// Returns -1 when the layer set does not fit the fixed ring geometry (the caller then launches the layers one by one).
bool b2p_conv_chain_enabled() { return (conv_mode() & 16) != 0; }

int b2p_launch_conv_chain(const UmmaConvArgs* args, int n, const B2PChainDep* deps, const int* n_reverse, const int* merge_next,
                          int* done_ws, cudaStream_t s) {
    if (n < 1 || n > CH_MAX_LAYERS) return -1;
    const int sms = device_sms();
    if (sms <= 0) return (int)cudaErrorInvalidDevice;
    static ChainParams cp;                       // large: keep it off the stack; filled and consumed under the lock
    static std::mutex cp_mutex;
    std::lock_guard<std::mutex> lk(cp_mutex);
    memset(&cp, 0, sizeof(cp));
    cp.n_layers = n;
    // ring depths: option chain_rings = 10 * A + B (activation slots of 40 KB, weight slots of 32 KB; A * 40 + B * 32 <= 222)
    const int rings = b2p_options().chain_rings;
    cp.ring_a = rings / 10; cp.ring_b = rings % 10;
    int rc;
    int max_atoms = TILE_ROWS;
    for (int l = 0; l < n; ++l) {
        if ((rc = fill_chain_layer(args[l], cp.L[l]))) return rc;
        if (cp.L[l].a_rows > max_atoms) max_atoms = cp.L[l].a_rows;
        cp.L[l].n_reverse = n_reverse ? n_reverse[l] : 0;
        if (cp.L[l].m_tiles != cp.L[0].m_tiles) return -1;
        if (cp.L[l].total_tiles / cp.L[l].m_groups > CH_MAX_NSUB) return -1;
    }
    cp.m_tiles = cp.L[0].m_tiles;
    // segments: layer l alone, or interleaved with layer l + 1 (merge_next[l], option chain_merge)
    for (int l = 0; l < n;) {
        const int sg = cp.n_seg++;
        const bool merge = merge_next && merge_next[l] && l + 1 < n && b2p_options().chain_merge != 0 &&
                           cp.L[l].m_groups == cp.L[l + 1].m_groups;
        cp.seg_layer[sg][0] = l; cp.seg_layer[sg][1] = merge ? l + 1 : -1;
        cp.seg_cnt[sg][0] = merge ? cp.L[l].total_units / cp.L[l].m_groups : 0;
        cp.seg_cnt[sg][1] = merge ? cp.L[l + 1].total_units / cp.L[l + 1].m_groups : 0;
        cp.seg_start[sg + 1] = cp.seg_start[sg] + cp.L[l].total_units + (merge ? cp.L[l + 1].total_units : 0);
        l += merge ? 2 : 1;
    }
    cp.a_slot = 2u * (uint32_t)max_atoms * 1024u;
    auto ring_bytes = [&](int ra, int rb) { return (size_t)ra * cp.a_slot + (size_t)rb * CH_B_SLOT + 1024 + 16 * (ra + rb) + 64 + 256; };
    if (cp.ring_a < 2 || cp.ring_b < 2 || cp.ring_a > 9 || cp.ring_b > 9 || ring_bytes(cp.ring_a, cp.ring_b) > 227 * 1024) { cp.ring_a = 2; cp.ring_b = 4; }
    if (ring_bytes(cp.ring_a, cp.ring_b) > 227 * 1024) return -1;
    for (int l = 0; l < n; ++l) {
        cp.dep[l].n_src = deps[l].n_src; cp.dep[l].halo = deps[l].halo;
        for (int k = 0; k < deps[l].n_src; ++k) {
            const int sl = deps[l].src[k];
            if (sl < 0 || sl >= l) return -1;                       // topological order
            const int units = cp.L[sl].total_tiles / cp.L[sl].m_groups;      // N units per tile of the source layer
            cp.dep[l].src[k] = sl;
            cp.dep[l].n_first[k] = deps[l].n_cnt[k] > 0 ? deps[l].n_first[k] : 0;
            cp.dep[l].n_cnt[k] = deps[l].n_cnt[k] > 0 ? deps[l].n_cnt[k] : units;
            if (cp.dep[l].n_first[k] < 0 || cp.dep[l].n_first[k] + cp.dep[l].n_cnt[k] > units) return -1;
        }
    }
    cp.done = done_ws;
    cp.next_unit = done_ws + (size_t)n * CH_MAX_NSUB * cp.m_tiles;
    cp.dynamic = b2p_options().chain_dynamic != 0;
    B2P_CUDA(cudaMemsetAsync(done_ws, 0, ((size_t)n * CH_MAX_NSUB * cp.m_tiles + 1) * sizeof(int), s));
    int nclusters = sms / 2;
    const size_t smem = ring_bytes(cp.ring_a, cp.ring_b);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * 2); cfg.blockDim = dim3(UM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = b2p_pdl_allowed(3);
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    // The units wait for each other, so every cluster of the grid must be resident at once: ask the driver how many CTA
    // pairs it can co-schedule (an SM without a free partner holds none) and launch no more than that.
    {
        static int max_clusters[64][10][10][2];
        int dev = 0;
        B2P_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64) return -1;
        int& cap = max_clusters[dev][cp.ring_a][cp.ring_b][cp.a_slot > 40960u ? 1 : 0];
        if (cap == 0) {
            int nmax = 0;
            B2P_CUDA(cudaOccupancyMaxActiveClusters(&nmax, conv_chain_kernel, &cfg));
            cap = nmax > 0 ? nmax : -1;
        }
        if (cap < 1) return -1;
        if (nclusters > cap) nclusters = cap;
        cfg.gridDim = dim3(nclusters * 2);
    }
    B2P_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, cp));
    B2P_LAUNCH_CHECK();
    return 0;
}

// ints of device scratch b2p_launch_conv_chain needs for n layers over m_tiles pixel tiles
size_t b2p_conv_chain_done_ints(int n, int m_tiles) { return (size_t)n * CH_MAX_NSUB * m_tiles + 1; }
