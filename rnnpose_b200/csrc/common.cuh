// Shared declarations for the libb200pose.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/b200pose.h"

#define B2P_LAUNCH_CHECK()                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

// Programmatic dependent launch (PDL): a kernel launched through b2p_launch_pdl may be scheduled while the previous
// kernel of the stream is still draining; it must execute pdl_wait() before it touches global memory (the wait returns
// when the previous grid has completed and its writes are visible).  pdl_trigger() lets the NEXT kernel of the stream be
// scheduled as soon as every block of this one has started.  Both are no-ops for a normally launched kernel.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline int& b2p_pdl_next_allowed() { static thread_local int v = 1; return v; }     // consumed by the next b2p_launch_pdl
template <typename... KArgs, typename... Args>
inline cudaError_t b2p_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = b2p_pdl_next_allowed();
    b2p_pdl_next_allowed() = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define B2P_CUDA(call)                                       \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

// Process-wide kernel-selection options (options.cu): read once from B200POSE_* environment variables at the first call
// into the library, changed afterwards only through b200pose_set_option.
struct B2POptions {
    int conv_mode, fg_list, fg_pipeline, fg_upsample, sparse_g1, fg_blocks, tail_min_n, conv_debug, lookup_mode, pool_mode,
        lm_debug, chain_rings, chain_dynamic, chain_xmajor, lm_cluster, chain_merge, upsample_variant, host_gather, sparse_g2, g2_margin, enc_chunk, enc_stem, host_gather_planes, pdl_off;
};
B2POptions& b2p_options();
// A/B switch (option pdl_off, bit mask): launch the tagged kernel WITHOUT the programmatic-dependent-launch attribute.
// tags: 0 flow_init, 1 lookup, 2 im2col_f1, 3 chained convolution launch, 4 upsample + weight, 5 LM
inline int b2p_pdl_allowed(int tag) { return (b2p_options().pdl_off >> tag) & 1 ? 0 : 1; }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// Packed weight blob (device).  All conv weights are stored GEMM-ready:
//   W[tap][cin_pad][cout_pad]  (cout contiguous), cin_pad = roundup(cin,16), cout_pad = roundup(cout,64)
// plus bias[cout_pad].  Fused layers: GRU z|r (cout 256), flow_head.conv1|mask.0 (cout 512).
// ------------------------------------------------------------------------------------------------
enum B2PConvId {
    CV_C1 = 0,   // encoder.convc1 1x1  324 -> 256
    CV_C2,       // encoder.convc2 3x3  256 -> 192
    CV_F1,       // encoder.convf1 7x7  2 -> 128, stored as 1x1 over a 98(+14)-wide im2col
    CV_F2,       // encoder.convf2 3x3  128 -> 64
    CV_ENC,      // encoder.conv   3x3  256 -> 126
    CV_ZR1,      // gru.convz1|convr1 1x5 384 -> 256
    CV_Q1,       // gru.convq1 1x5 384 -> 128
    CV_ZR2,      // gru.convz2|convr2 5x1
    CV_Q2,       // gru.convq2 5x1
    CV_HEADS,    // flow_head.conv1|mask.0 3x3 128 -> 512
    CV_MASK2,    // mask.2 1x1 256 -> 576
    CV_COUNT
};

struct B2PConvDesc {
    int kh, kw, cin, cout, cin_pad, cout_pad;
    size_t w_off, b_off;   // float offsets into the packed blob
};

struct B2PWeightLayout {
    B2PConvDesc cv[CV_COUNT];
    size_t fh2_w_off;      // flow_head.conv2 as [2][9][256]
    size_t fh2_b_off;      // [2]
    size_t total_floats;
};

const B2PWeightLayout& b2p_weight_layout();

// conv epilogues
enum { EPI_NONE = 0, EPI_RELU = 1, EPI_SCALE = 2, EPI_GRU_ZR = 3, EPI_GRU_Q = 4, EPI_FLOW = 5 };

struct ConvParams {
    const float* src0; int pitch0; int c0;   // input segment 0 (PXC): pointer at first channel, pixel pitch, channels
    const float* src1; int pitch1; int c1;   // optional segment 1 (c1 = 0 when unused)
    const float* wgt;                         // [taps][cin_pad][cout_pad]
    const float* bias;                        // [cout_pad]
    float* dst; int dst_pitch;                // output (pointer at first channel), pixel pitch
    int cout, cout_pad, cin_pad;
    int B, h, w, kh, kw;
    int epi; float scale;
    float* zbuf;                              // GRU: z gate [P][128]
    float* rhbuf;                             // GRU: r*h    [P][128]
    float* hbuf;                              // GRU: hidden state [P][128] (read; written by EPI_GRU_Q)
};

int b2p_launch_conv(const ConvParams& p, cudaStream_t s);

// ------------------------------------------------------------------------------------------------
// Tensor-core path (conv_umma.cu): operands are fp16 hi/lo plane pairs, x ~= hi + lo.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void b2p_split_half(float v, __half& hi, __half& lo) {
#ifdef __CUDA_ARCH__
    unsigned short h, l;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));          // saturate instead of +-inf
    hi = __ushort_as_half(h);
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(v - __half2float(hi)));
    lo = __ushort_as_half(l);
#else
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
#endif
}

// Tiled layout of the fp32 buffers that only the tensor-core epilogues touch (z gate, hidden state, GRU pre-sums):
// [pixel tile (16 rows x 8 pixels, the conv_umma M tile)][C/4][128 pixels of the tile, row-major][4 channels].
// The 32 pixels a warp owns are then contiguous for every float4 of channels.
constexpr int B2P_TILE_ROWS = 16, B2P_TILE_COLS = 8;
__host__ __device__ __forceinline__ size_t b2p_tiled_index(int tile, int m, int c, int C) {
    return (((size_t)tile * (C >> 2) + (c >> 2)) * 128 + m) * 4 + (c & 3);
}
static inline size_t b2p_tiled_pixels(int B, int h, int w) {       // pixel slots of a tiled buffer (>= B*h*w)
    return (size_t)B * ((h + B2P_TILE_ROWS - 1) / B2P_TILE_ROWS) * ((w + B2P_TILE_COLS - 1) / B2P_TILE_COLS) * 128;
}
// fp32 [P][C] pixel-major <-> tiled
int b2p_pxc_to_tiled(const float* src, float* dst, int B, int h, int w, int C, cudaStream_t s, int xmajor = 0);
int b2p_tiled_to_pxc(const float* src, float* dst, int B, int h, int w, int C, cudaStream_t s);

// fp16 weight planes: W[tap][cout_pad][cin_pad] (cin contiguous = K-major), cin_pad multiple of 64,
// cout_pad multiple of n_tile.
struct B2PHalfConvDesc {
    int kh, kw, cin, cout, cin_pad, cout_pad, n_tile;
    size_t hi_off, lo_off;     // offsets in halves from the start of the fp16 section
};
struct B2PHalfLayout {
    B2PHalfConvDesc cv[CV_COUNT];
    B2PHalfConvDesc fh2;       // flow_head.conv2 (3x3, 256 -> 2) padded to 32 output channels: the last layer of the chained launch
    size_t total_halves;
};
const B2PHalfLayout& b2p_half_layout();
// byte offset of the fp16 section inside the packed blob (after the fp32 section)
size_t b2p_half_section_offset_bytes();

struct UmmaConvArgs {
    const __half* seg_hi[2]; const __half* seg_lo[2]; int seg_c[2]; int seg_pitch[2];
    const __half* w_hi; const __half* w_lo; const float* bias;
    int cin_pad, cout_pad, cout, n_tile, kh, kw, B, h, w;
    int epi; float scale;
    float* out_f32; int out_f32_pitch;
    __half* out_hi; __half* out_lo; int out_h_pitch;
    float* zbuf; float* hbuf;
    float* hbuf_x;             // chained launch with x-major 1 x kw layers: the hidden state's copy in x-major tile order (or nullptr)
    float* fl_coords1; float* fl_flow; float* fl_dflow;   // EPI_FLOW: coords1 [P][2] in/out, flow [P][2] out, delta [P][2] out or nullptr
    unsigned chunk_mask;       // bit cc set = visit 64-channel chunk cc of every tap (0 = all chunks)
    const float* pre; int pre_pitch;   // fp32 [P][pre_pitch] partial sums added in the epilogue (or nullptr)
    int layer_id;              // B2PConvId of an update-block layer, or -1
    int side_tiled;            // zbuf / hbuf / pre use the tiled side-buffer layout (b2p_tiled_index)
    int out_tiled;             // EPI_SCALE: out_f32 is a tiled side buffer with out_f32_pitch channels
    int b_batched;             // weights differ per sample: 3rd weight-map coordinate = sample index (1x1 only)
    int stride, in_h, in_w;    // stride 2 (encoder): h, w are the OUTPUT dimensions, in_h x in_w the input map (0: same as h, w)
};
int b2p_launch_conv_umma(const UmmaConvArgs& a, cudaStream_t s);
// several layers in one persistent launch with tile-level dependencies (conv_chain_kernel, conv_umma.cu)
struct B2PChainDep {
    int n_src, src[2];                 // indices (in list order) of the layers whose output this layer reads
    int n_first[2], n_cnt[2];          // N units of the source that are read (n_cnt = 0: all of them)
    int halo;                          // 1: 3x3 tile neighbourhood, 0: the same tile only
};
bool b2p_conv_chain_enabled();
int b2p_launch_conv_chain(const UmmaConvArgs* args, int n, const B2PChainDep* deps, const int* n_reverse, const int* merge_next,
                          int* done_ws, cudaStream_t s);
size_t b2p_conv_chain_done_ints(int n, int m_tiles);
// fp32 [P][pitch_in] -> fp16 hi/lo planes [P][pitch_out] (first C channels); used by the per-operator entry
int b2p_split_planes(const float* src, int pitch_in, int C, size_t P, __half* hi, __half* lo, int pitch_out, cudaStream_t s);

// kernels implemented across the .cu files (host launchers; all return 0 / cudaError_t)
int b2p_corr_volume(const float* f1, const float* f2, int B, int D, int P, float* level0, cudaStream_t s);
int b2p_corr_pool(const float* src, int NP, int hs, int ws, float* dst, cudaStream_t s);
int b2p_corr_pool3(float* pyramid, int B, int h, int w, cudaStream_t s);   // levels 1..3 in one pass, or -1 (not applicable)
int b2p_fmap_to_pxc_half(const float* f, int B, int D, int P, __half* hi, __half* lo, cudaStream_t s);
// hi/lo != nullptr: additionally (or instead, when out == nullptr) write fp16 hi/lo planes with the same pitch
int b2p_corr_lookup(const float* pyramid, const float* coords, int B, int h, int w, float* out, __half* out_hi,
                    __half* out_lo, cudaStream_t s);
int b2p_context_init(const float* ctx, int B, int H, int W, float* net, float* xbuf, __half* net_hi, __half* net_lo,
                     __half* x_hi, __half* x_lo, cudaStream_t s, const float* texels = nullptr, int c_split = 0);
void b2p_context_sample_taps(int in, int out, int* i0, int* i1);
int b2p_flow_init(const float* depth, const float* K, const float* G, int B, int H, int W,
                  float* coords1, float* flow, cudaStream_t s);
int b2p_im2col_f1(const float* flow, int B, int h, int w, float* col /*[P][112]*/, float* xbuf /*[P][256] ch 254,255*/,
                  __half* col_hi, __half* col_lo, __half* x_hi, __half* x_lo, cudaStream_t s);
int b2p_flow_head2(const float* hm /*[P][512] fp32, first 256 = flow-head features; or nullptr*/, const __half* hm_hi,
                   const __half* hm_lo, const float* w2, const float* b2, float* coords1 /*[P][2] in/out*/,
                   float* flow /*[P][2] out = coords1 - coords0*/, float* dflow_out, float* part /*scratch [P][20]*/,
                   int B, int h, int w, cudaStream_t s);
int b2p_upsample_weight(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth,
                        float sigma, int B, int C, int H, int W, float* flow_up, float* target, float* weight,
                        int lazy_background, cudaStream_t s, const float* g2_far = nullptr, const int* g2_window = nullptr);
// foreground list (depth > 0 or non-finite) of a call, built once; see upsample_weight.cu
size_t b2p_fg_ws_bytes(int B, int H, int W);
int b2p_fg_build(const float* depth, int B, int H, int W, void* fg_ws, float* target, float* weight, cudaStream_t s);
int b2p_upsample_weight_fg(const float* flow, const float* mask, const float* g1, const float* g2, const float* depth, float sigma,
                           int B, int C, int H, int W, const void* fg_ws, float* target, float* weight, cudaStream_t s);
const int* b2p_fg_idx(const void* fg_ws);
const int* b2p_fg_count(const void* fg_ws, int B, int H, int W);
// dst[b][c][r] = src[b][c][r] for the pixels with depth[b][r] > 0 (src may be a mapped host pointer)
int b2p_gather_fg_planes(const float* depth_dev, const float* src_mapped, float* dst_dev, int B, int C, int H, int W, cudaStream_t s);
int b2p_lm_step(const float* depth, const float* target, const float* weight, const float* K, float* G,
                int B, int H, int W, float depth_add, double ep, double lm, double* H_out, double* b_out,
                float* delta_out, void* ws, cudaStream_t s);
// fg_idx / fg_count != nullptr: visit only the listed pixels (every other pixel must have weight 0 and finite inputs)
int b2p_lm_steps(const float* depth, const float* target, const float* weight, const float* K, float* G,
                 int B, int H, int W, float depth_add, double ep, double lm, int n_steps, void* ws, cudaStream_t s,
                 const int* fg_idx = nullptr, const int* fg_count = nullptr);
size_t b2p_lm_ws_bytes(int B, int H, int W);
size_t b2p_lm_bwd_ws_bytes(int B, int H, int W);
int b2p_lm_backward(const float* depth, const float* target, const float* weight, const float* K, const float* G, const float* grad_delta,
                    int B, int H, int W, float depth_add, double ep, double lm, float* grad_target, float* grad_weight, void* ws,
                    cudaStream_t s);
int b2p_se3_retract(const float* delta, float* G, int B, cudaStream_t s);
int b2p_chol_solve(const double* H, const double* b, float* x, int B, cudaStream_t s);
int b2p_lm_reset(void* ws, int B, int H, int W, cudaStream_t s);   // once before the first b2p_lm_step on a workspace
int b2p_lm_cluster(const float4* rec, const float* depth, const float* target, const float* weight, const int* fg_idx, const int* fg_count,
                   const float* K, float* G, int B, int H, int W, float depth_add, double ep, double lm, int n_steps, cudaStream_t s);
// foreground pipeline (fg_pipeline.cu): channels-last descriptors, float4 records per listed pixel
size_t b2p_fgpipe_ws_bytes(int B, int H, int W);
const float4* b2p_fgpipe_records(const void* ws, int B, int H, int W);
int b2p_fgpipe_prepare(const float* g1, const float* g2, int g2_is_cl, int B, int H, int W, const void* fg_ws, void* ws, cudaStream_t s);
int b2p_fgpipe_upsample_weight(const float* flow, const float* mask, const float* g2_cl_or_null, const float* depth, float sigma, int B,
                               int H, int W, const void* fg_ws, void* ws, float* weight_dense, cudaStream_t s);
size_t b2p_encoder_packed_bytes();
int b2p_encoder_pack(const float* const* t, void* packed, cudaStream_t s);
size_t b2p_encoder_ws_bytes(int B, int H, int W);
int b2p_image_encoder(const void* packed, const float* image1, const float* image2, int B, int H, int W, float* fmap1, float* fmap2,
                      void* ws, cudaStream_t s);
size_t b2p_zoom_crop_ws_bytes(int B);
int b2p_zoom_crop(const float* pc_depth, const float* K, const float* T, const float* image, const float* geo, int B, int Ci, int Cg,
                  int H, int W, int Hc, int Wc, float margin_ratio, int geo_channels_last, float* image_crop, float* geo_crop,
                  float* K_crop, float* theta, void* ws, cudaStream_t s);
size_t b2p_pose_metrics_ws_bytes(int B, int n);
int b2p_pose_metrics(const float* T_pred, const float* T_gt, const float* pts, const float* diameter, const float* K, int B,
                     int n, float* out, void* ws, cudaStream_t s);
int b2p_pack_weights(const float* const* t, float* packed, cudaStream_t s);   // fp32 section + fp16 hi/lo section
