// C-ABI entry points of libb200pose.so (declared in include/b200pose.h) and the host-side sequencing of
// the inner loop (reference model/PoseRefiner.py:315-362, model/CFNet.py:109-173).
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <sched.h>
#include <xmmintrin.h>
#include <atomic>
#include <memory>
#include <thread>
#include <vector>

namespace {

struct Carver {
    char* base; size_t off; size_t cap;
    Carver(void* p, size_t c) : base(reinterpret_cast<char*>(p)), off(0), cap(c) {}
    template <typename T> T* take(size_t n) {
        off = align_up(off, 1024);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
};

// scratch of one update-block pass.  fp32 buffers serve the exact FFMA path; the __half hi/lo plane pairs
// ([0] = hi, [1] = lo) serve the tensor-core path.
struct UpdateWs {
    float *col, *c1, *corflo, *f1o, *zbuf, *rhbuf, *hm;
    float* hbuf_x;            // tensor-core path: the hidden state's second tiled copy, x-major tile order (chained launch)
    float* pre[4];            // GRU partial sums of the iteration-invariant `inp` channels: zr1 [P][256], q1 [P][128], zr2, q2
    float* zero_bias;         // 1024 zeros
    __half *corr_h[2], *col_h[2], *c1_h[2], *corflo_h[2], *f1o_h[2], *x_h[2], *net_h[2], *rh_h[2], *hm_h[2];
};

size_t update_ws_layout(int B, int h, int w, void* ws, size_t cap, UpdateWs* out) {
    const size_t P = (size_t)B * h * w;
    const size_t Pt = b2p_tiled_pixels(B, h, w);       // pixel slots of the tiled side buffers (z, h, pre-sums), >= P
    Carver c(ws, cap);
    UpdateWs u;
    u.col = c.take<float>(P * 112);
    u.c1 = c.take<float>(P * 256);
    u.corflo = c.take<float>(P * 256);
    u.f1o = c.take<float>(P * 128);
    u.zbuf = c.take<float>(Pt * 128);
    u.rhbuf = c.take<float>(Pt * 128);             // exact path: r*h (PXC); tensor-core path: the hidden state h (tiled)
    u.hm = c.take<float>(P * 512);
    u.hbuf_x = c.take<float>(Pt * 128);
    u.pre[0] = c.take<float>(Pt * 256); u.pre[1] = c.take<float>(Pt * 128);
    u.pre[2] = c.take<float>(Pt * 256); u.pre[3] = c.take<float>(Pt * 128);
    u.zero_bias = c.take<float>(1024);
    for (int k = 0; k < 2; ++k) {
        u.corr_h[k] = c.take<__half>(P * B200POSE_CORR_PITCH);
        u.col_h[k] = c.take<__half>(P * 112);
        u.c1_h[k] = c.take<__half>(P * 256);
        u.corflo_h[k] = c.take<__half>(P * 256);
        u.f1o_h[k] = c.take<__half>(P * 128);
        u.x_h[k] = c.take<__half>(P * 256);
        u.net_h[k] = c.take<__half>(P * 128);
        u.rh_h[k] = c.take<__half>(P * 128);
        u.hm_h[k] = c.take<__half>(P * 512);
    }
    if (out) *out = u;
    return align_up(c.off, 1024);
}

inline bool fg_list_enabled() { return b2p_options().fg_list != 0; }
inline bool fg_upsample_enabled() { return b2p_options().fg_upsample == 1; }   // round-1 list-driven upsample + weight kernel (opt-in)

inline bool shape_ok(int B, int H, int W) {
    return B >= 1 && H >= 128 && W >= 128 && (H % 8) == 0 && (W % 8) == 0;   // (H/8)>>3 >= 2: the reference's
}                                                                             // sampler divides by (w_l - 1)

// ------------------------------------------------------------------------------------------------
// exact fp32 path (CUDA-core FFMA implicit GEMM)
// ------------------------------------------------------------------------------------------------
int run_update_block(const float* wts, float* net, float* xbuf, const float* corr, float* coords1, float* flow,
                     float* mask, float* dflow_out, int B, int h, int w, const UpdateWs& u, cudaStream_t s) {
    const B2PWeightLayout& L = b2p_weight_layout();
    int rc;
    auto conv = [&](int id, const float* s0, int p0, int c0, const float* s1, int p1, int c1n, float* dst, int dpitch,
                    int epi, float scale) -> int {
        const B2PConvDesc& d = L.cv[id];
        ConvParams p;
        p.src0 = s0; p.pitch0 = p0; p.c0 = c0; p.src1 = s1; p.pitch1 = p1; p.c1 = c1n;
        p.wgt = wts + d.w_off; p.bias = wts + d.b_off; p.dst = dst; p.dst_pitch = dpitch;
        p.cout = d.cout; p.cout_pad = d.cout_pad; p.cin_pad = d.cin_pad;
        p.B = B; p.h = h; p.w = w; p.kh = d.kh; p.kw = d.kw; p.epi = epi; p.scale = scale;
        p.zbuf = u.zbuf; p.rhbuf = u.rhbuf; p.hbuf = net;
        return b2p_launch_conv(p, s);
    };
    // motion encoder (update.py:89-97)
    if ((rc = b2p_im2col_f1(flow, B, h, w, u.col, xbuf, nullptr, nullptr, nullptr, nullptr, s))) return rc;   // + cat[out, flow]
    if ((rc = conv(CV_C1, corr, B200POSE_CORR_PITCH, 324, nullptr, 0, 0, u.c1, 256, EPI_RELU, 1.f))) return rc;
    if ((rc = conv(CV_C2, u.c1, 256, 256, nullptr, 0, 0, u.corflo, 256, EPI_RELU, 1.f))) return rc;    // cor -> [0,192)
    if ((rc = conv(CV_F1, u.col, 112, 98, nullptr, 0, 0, u.f1o, 128, EPI_RELU, 1.f))) return rc;
    if ((rc = conv(CV_F2, u.f1o, 128, 128, nullptr, 0, 0, u.corflo + 192, 256, EPI_RELU, 1.f))) return rc;  // flo -> [192,256)
    if ((rc = conv(CV_ENC, u.corflo, 256, 256, nullptr, 0, 0, xbuf + 128, 256, EPI_RELU, 1.f))) return rc;  // out -> x[128,254)
    // SepConvGRU (update.py:45-60): hx = [h | inp | motion]
    if ((rc = conv(CV_ZR1, net, 128, 128, xbuf, 256, 256, nullptr, 0, EPI_GRU_ZR, 1.f))) return rc;
    if ((rc = conv(CV_Q1, u.rhbuf, 128, 128, xbuf, 256, 256, nullptr, 0, EPI_GRU_Q, 1.f))) return rc;
    if ((rc = conv(CV_ZR2, net, 128, 128, xbuf, 256, 256, nullptr, 0, EPI_GRU_ZR, 1.f))) return rc;
    if ((rc = conv(CV_Q2, u.rhbuf, 128, 128, xbuf, 256, 256, nullptr, 0, EPI_GRU_Q, 1.f))) return rc;
    // heads (update.py:13-14, 172-176, 187)
    if ((rc = conv(CV_HEADS, net, 128, 128, nullptr, 0, 0, u.hm, 512, EPI_RELU, 1.f))) return rc;
    if ((rc = b2p_flow_head2(u.hm, nullptr, nullptr, wts + L.fh2_w_off, wts + L.fh2_b_off, coords1, flow, dflow_out, u.col, B, h, w, s))) return rc;
    if ((rc = conv(CV_MASK2, u.hm + 256, 512, 256, nullptr, 0, 0, mask, 576, EPI_SCALE, 0.25f))) return rc;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// tensor-core path (tcgen05, fp16 hi/lo operands).  Expects u.corr_h, u.net_h and u.x_h[:, 0:128] filled.
// ------------------------------------------------------------------------------------------------
// use_pre: the `inp` channels (chunks 2,3 of the 384-wide GRU input [h | inp | motion]) do not change between the
// recurrent iterations of one render iteration (CFNet.py:124-133 computes inp once): their contribution to the six GRU
// convolutions is computed once (run_gru_precompute) and added in the epilogue, so the per-iteration GEMMs skip 1/3 of K.
int run_update_block_tc_chain(const float* wts, float* net, float* coords1, float* flow, float* mask, float* dflow_out,
                              int B, int h, int w, const UpdateWs& u, bool use_pre, cudaStream_t s);

// Timing hook (b200pose_debug_set_conv_events): the next tensor-core update-block pass records these two events on its stream
// immediately before the first and after the last convolution launch (the chained launch, or the eleven layer launches).
cudaEvent_t g_conv_ev[2] = {nullptr, nullptr};

int run_update_block_tc(const float* wts, float* net, float* coords1, float* flow, float* mask, float* dflow_out,
                        int B, int h, int w, const UpdateWs& u, bool use_pre, cudaStream_t s) {
    if (b2p_conv_chain_enabled() && B * ceil_div(h, B2P_TILE_ROWS) * ceil_div(w, B2P_TILE_COLS) >= 296) {     // (also in b200pose_refine_launch_count)
        // experimental single-launch variant; -1 = not applicable, fall through to the layer-by-layer pass
        const int rcc = run_update_block_tc_chain(wts, net, coords1, flow, mask, dflow_out, B, h, w, u, use_pre, s);
        if (rcc != -1) return rcc;
    }
    const B2PWeightLayout& L = b2p_weight_layout();
    const B2PHalfLayout& HL = b2p_half_layout();
    const __half* hbase = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(wts) + b2p_half_section_offset_bytes());
    int rc;
    auto conv = [&](int id, __half* const* s0, int off0, int c0, int p0, __half* const* s1, int c1n, int p1,
                    __half* const* dst, int doff, int dpitch, int epi, float scale, float* out_f32, int f32_pitch,
                    const float* pre = nullptr, int pre_pitch = 0) -> int {
        const B2PHalfConvDesc& d = HL.cv[id];
        UmmaConvArgs a;
        memset(&a, 0, sizeof(a));
        a.layer_id = id;
        if (pre) { a.pre = pre; a.pre_pitch = pre_pitch; a.chunk_mask = 0x33u; }     // chunks {0,1,4,5}: h and motion
        a.seg_hi[0] = s0[0] + off0; a.seg_lo[0] = s0[1] + off0; a.seg_c[0] = c0; a.seg_pitch[0] = p0;
        if (s1) { a.seg_hi[1] = s1[0]; a.seg_lo[1] = s1[1]; a.seg_c[1] = c1n; a.seg_pitch[1] = p1; }
        a.w_hi = hbase + d.hi_off; a.w_lo = hbase + d.lo_off; a.bias = wts + L.cv[id].b_off;
        a.cin_pad = d.cin_pad; a.cout_pad = d.cout_pad; a.cout = d.cout; a.n_tile = d.n_tile; a.kh = d.kh; a.kw = d.kw;
        a.B = B; a.h = h; a.w = w; a.epi = epi; a.scale = scale;
        a.out_f32 = out_f32; a.out_f32_pitch = f32_pitch;
        if (dst) { a.out_hi = dst[0] + doff; a.out_lo = dst[1] + doff; a.out_h_pitch = dpitch; }
        a.zbuf = u.zbuf; a.hbuf = u.rhbuf; a.side_tiled = 1;       // h: tiled fp32 copy of the hidden state (see callers)
        return b2p_launch_conv_umma(a, s);
    };
    if ((rc = b2p_im2col_f1(flow, B, h, w, nullptr, nullptr, u.col_h[0], u.col_h[1], u.x_h[0], u.x_h[1], s))) return rc;
    if (g_conv_ev[0]) B2P_CUDA(cudaEventRecord(g_conv_ev[0], s));
    if ((rc = conv(CV_C1, u.corr_h, 0, B200POSE_CORR_PITCH, B200POSE_CORR_PITCH, nullptr, 0, 0, u.c1_h, 0, 256, EPI_RELU, 1.f, nullptr, 0))) return rc;
    if ((rc = conv(CV_C2, u.c1_h, 0, 256, 256, nullptr, 0, 0, u.corflo_h, 0, 256, EPI_RELU, 1.f, nullptr, 0))) return rc;
    if ((rc = conv(CV_F1, u.col_h, 0, 112, 112, nullptr, 0, 0, u.f1o_h, 0, 128, EPI_RELU, 1.f, nullptr, 0))) return rc;
    if ((rc = conv(CV_F2, u.f1o_h, 0, 128, 128, nullptr, 0, 0, u.corflo_h, 192, 256, EPI_RELU, 1.f, nullptr, 0))) return rc;
    if ((rc = conv(CV_ENC, u.corflo_h, 0, 256, 256, nullptr, 0, 0, u.x_h, 128, 256, EPI_RELU, 1.f, nullptr, 0))) return rc;
    const float* pz1 = use_pre ? u.pre[0] : nullptr; const float* pq1 = use_pre ? u.pre[1] : nullptr;
    const float* pz2 = use_pre ? u.pre[2] : nullptr; const float* pq2 = use_pre ? u.pre[3] : nullptr;
    if ((rc = conv(CV_ZR1, u.net_h, 0, 128, 128, u.x_h, 256, 256, u.rh_h, 0, 128, EPI_GRU_ZR, 1.f, nullptr, 0, pz1, 256))) return rc;
    if ((rc = conv(CV_Q1, u.rh_h, 0, 128, 128, u.x_h, 256, 256, u.net_h, 0, 128, EPI_GRU_Q, 1.f, nullptr, 0, pq1, 128))) return rc;
    if ((rc = conv(CV_ZR2, u.net_h, 0, 128, 128, u.x_h, 256, 256, u.rh_h, 0, 128, EPI_GRU_ZR, 1.f, nullptr, 0, pz2, 256))) return rc;
    if ((rc = conv(CV_Q2, u.rh_h, 0, 128, 128, u.x_h, 256, 256, u.net_h, 0, 128, EPI_GRU_Q, 1.f, nullptr, 0, pq2, 128))) return rc;
    if ((rc = conv(CV_HEADS, u.net_h, 0, 128, 128, nullptr, 0, 0, u.hm_h, 0, 512, EPI_RELU, 1.f, nullptr, 0))) return rc;
    if ((rc = conv(CV_MASK2, u.hm_h, 256, 256, 512, nullptr, 0, 0, nullptr, 0, 0, EPI_SCALE, 0.25f, mask, 576))) return rc;
    if (g_conv_ev[1]) { B2P_CUDA(cudaEventRecord(g_conv_ev[1], s)); g_conv_ev[0] = g_conv_ev[1] = nullptr; }
    if ((rc = b2p_flow_head2(nullptr, u.hm_h[0], u.hm_h[1], wts + L.fh2_w_off, wts + L.fh2_b_off, coords1, flow, dflow_out, u.col, B, h, w, s))) return rc;
    return 0;
}

// conv_mode bit 4: the same pass with the eleven convolutions in one persistent launch (conv_chain_kernel).  Returns -1 when the chain cannot take this problem; the caller then runs the
// layer-by-layer version above.
int run_update_block_tc_chain(const float* wts, float* net, float* coords1, float* flow, float* mask, float* dflow_out,
                              int B, int h, int w, const UpdateWs& u, bool use_pre, cudaStream_t s) {
    (void)net;
    const B2PWeightLayout& L = b2p_weight_layout();
    const B2PHalfLayout& HL = b2p_half_layout();
    const __half* hbase = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(wts) + b2p_half_section_offset_bytes());
    UmmaConvArgs args[12];
    int n = 0;
    auto add = [&](int id, __half* const* s0, int off0, int c0, int p0, __half* const* s1, int c1n, int p1,
                   __half* const* dst, int doff, int dpitch, int epi, float scale, float* out_f32, int f32_pitch,
                   const float* pre = nullptr, int pre_pitch = 0) {
        const B2PHalfConvDesc& d = HL.cv[id];
        UmmaConvArgs& a = args[n++];
        memset(&a, 0, sizeof(a));
        a.layer_id = id;
        if (pre) { a.pre = pre; a.pre_pitch = pre_pitch; a.chunk_mask = 0x33u; }
        a.seg_hi[0] = s0[0] + off0; a.seg_lo[0] = s0[1] + off0; a.seg_c[0] = c0; a.seg_pitch[0] = p0;
        if (s1) { a.seg_hi[1] = s1[0]; a.seg_lo[1] = s1[1]; a.seg_c[1] = c1n; a.seg_pitch[1] = p1; }
        a.w_hi = hbase + d.hi_off; a.w_lo = hbase + d.lo_off; a.bias = wts + L.cv[id].b_off;
        a.cin_pad = d.cin_pad; a.cout_pad = d.cout_pad; a.cout = d.cout; a.n_tile = d.n_tile; a.kh = d.kh; a.kw = d.kw;
        a.B = B; a.h = h; a.w = w; a.epi = epi; a.scale = scale;
        a.out_f32 = out_f32; a.out_f32_pitch = f32_pitch;
        if (dst) { a.out_hi = dst[0] + doff; a.out_lo = dst[1] + doff; a.out_h_pitch = dpitch; }
        a.zbuf = u.zbuf; a.hbuf = u.rhbuf; a.hbuf_x = u.hbuf_x; a.side_tiled = 1;
    };
    const float* pz1 = use_pre ? u.pre[0] : nullptr; const float* pq1 = use_pre ? u.pre[1] : nullptr;
    const float* pz2 = use_pre ? u.pre[2] : nullptr; const float* pq2 = use_pre ? u.pre[3] : nullptr;
    // List order: a layer's units follow its sources' units by at least one whole layer where the graph allows it (F1 before
    // C2 so that F2 does not wait on the F1 units issued just before it; the mask half of HEADS first, see n_reverse).
    //  0 C1   1 F1   2 C2   3 F2   4 ENC   5 ZR1   6 Q1   7 ZR2   8 Q2   9 HEADS   10 MASK2
    add(CV_C1, u.corr_h, 0, B200POSE_CORR_PITCH, B200POSE_CORR_PITCH, nullptr, 0, 0, u.c1_h, 0, 256, EPI_RELU, 1.f, nullptr, 0);
    add(CV_F1, u.col_h, 0, 112, 112, nullptr, 0, 0, u.f1o_h, 0, 128, EPI_RELU, 1.f, nullptr, 0);
    add(CV_C2, u.c1_h, 0, 256, 256, nullptr, 0, 0, u.corflo_h, 0, 256, EPI_RELU, 1.f, nullptr, 0);
    add(CV_F2, u.f1o_h, 0, 128, 128, nullptr, 0, 0, u.corflo_h, 192, 256, EPI_RELU, 1.f, nullptr, 0);
    add(CV_ENC, u.corflo_h, 0, 256, 256, nullptr, 0, 0, u.x_h, 128, 256, EPI_RELU, 1.f, nullptr, 0);
    add(CV_ZR1, u.net_h, 0, 128, 128, u.x_h, 256, 256, u.rh_h, 0, 128, EPI_GRU_ZR, 1.f, nullptr, 0, pz1, 256);
    add(CV_Q1, u.rh_h, 0, 128, 128, u.x_h, 256, 256, u.net_h, 0, 128, EPI_GRU_Q, 1.f, nullptr, 0, pq1, 128);
    add(CV_ZR2, u.net_h, 0, 128, 128, u.x_h, 256, 256, u.rh_h, 0, 128, EPI_GRU_ZR, 1.f, nullptr, 0, pz2, 256);
    add(CV_Q2, u.rh_h, 0, 128, 128, u.x_h, 256, 256, u.net_h, 0, 128, EPI_GRU_Q, 1.f, nullptr, 0, pq2, 128);
    add(CV_HEADS, u.net_h, 0, 128, 128, nullptr, 0, 0, u.hm_h, 0, 512, EPI_RELU, 1.f, nullptr, 0);
    add(CV_MASK2, u.hm_h, 256, 256, 512, nullptr, 0, 0, nullptr, 0, 0, EPI_SCALE, 0.25f, mask, 576);
    {   // 11 FH2: flow_head.conv2 (3x3, 256 -> 2, padded to 32 columns) on the flow half of HEADS; its epilogue applies
        // coords1 += delta and flow = coords1 - coords0 (the two flow_head2 FFMA launches of the layer-by-layer path)
        const B2PHalfConvDesc& d = HL.fh2;
        UmmaConvArgs& a = args[n++];
        memset(&a, 0, sizeof(a));
        a.layer_id = -1;
        a.seg_hi[0] = u.hm_h[0]; a.seg_lo[0] = u.hm_h[1]; a.seg_c[0] = 256; a.seg_pitch[0] = 512;
        a.w_hi = hbase + d.hi_off; a.w_lo = hbase + d.lo_off; a.bias = wts + L.fh2_b_off;       // 2 biases followed by zeros
        a.cin_pad = d.cin_pad; a.cout_pad = d.cout_pad; a.cout = d.cout; a.n_tile = d.n_tile; a.kh = 3; a.kw = 3;
        a.B = B; a.h = h; a.w = w; a.epi = EPI_FLOW; a.scale = 1.f;
        a.fl_coords1 = coords1; a.fl_flow = flow; a.fl_dflow = dflow_out;
    }
    // which earlier layers each layer reads (also the layers whose readers it must not overtake, see conv_chain_kernel).
    // HEADS lists its N unit 1 (channels 256..511, the mask half) first: MASK2 reads only that unit, FH2 only unit 0.
    B2PChainDep deps[12];
    memset(deps, 0, sizeof(deps));
    auto dep = [&](int l, int halo, int s0, int s1 = -1) {
        deps[l].halo = halo; deps[l].n_src = s0 < 0 ? 0 : (s1 < 0 ? 1 : 2);
        deps[l].src[0] = s0 < 0 ? 0 : s0; deps[l].src[1] = s1 < 0 ? 0 : s1;
    };
    dep(0, 0, -1); dep(1, 0, -1); dep(2, 1, 0); dep(3, 1, 1); dep(4, 1, 2, 3); dep(5, 1, 4); dep(6, 1, 5); dep(7, 1, 6);
    dep(8, 1, 7); dep(9, 1, 8); dep(10, 0, 9); dep(11, 1, 9);
    deps[10].n_first[0] = 1; deps[10].n_cnt[0] = 1;
    deps[11].n_first[0] = 0; deps[11].n_cnt[0] = 1;
    int n_reverse[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0};
    // interleaved pairs: C1 | F1 (F1's unit is all epilogue, C1's waits for operands) and MASK2 | flow head (MASK2 is
    // epilogue-bound, the 32-column flow head MMA-bound)
    const int merge_next[12] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0};
    int rc;
    if ((rc = b2p_im2col_f1(flow, B, h, w, nullptr, nullptr, u.col_h[0], u.col_h[1], u.x_h[0], u.x_h[1], s))) return rc;
    // completion counters: the fp32 scratch of the exact path (unused here)
    const int m_tiles = B * ceil_div(h, B2P_TILE_ROWS) * ceil_div(w, B2P_TILE_COLS);
    if (b2p_conv_chain_done_ints(n, m_tiles) * sizeof(int) > (size_t)B * h * w * 256 * sizeof(float)) return -1;
    if (g_conv_ev[0]) B2P_CUDA(cudaEventRecord(g_conv_ev[0], s));
    if ((rc = b2p_launch_conv_chain(args, n, deps, n_reverse, merge_next, reinterpret_cast<int*>(u.c1), s))) return rc;
    if (g_conv_ev[1]) { B2P_CUDA(cudaEventRecord(g_conv_ev[1], s)); g_conv_ev[0] = g_conv_ev[1] = nullptr; }
    return 0;
}

// GRU partial sums over the `inp` channels only (chunks 2,3), no bias, no activation: pre[k][P][cout] fp32.
int run_gru_precompute(const float* wts, int B, int h, int w, const UpdateWs& u, cudaStream_t s) {
    const B2PHalfLayout& HL = b2p_half_layout();
    const __half* hbase = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(wts) + b2p_half_section_offset_bytes());
    B2P_CUDA(cudaMemsetAsync(u.zero_bias, 0, 1024 * sizeof(float), s));
    const int ids[4] = {CV_ZR1, CV_Q1, CV_ZR2, CV_Q2};
    UmmaConvArgs args[4];
    for (int k = 0; k < 4; ++k) {
        const B2PHalfConvDesc& d = HL.cv[ids[k]];
        UmmaConvArgs& a = args[k];
        memset(&a, 0, sizeof(a));
        // segment 0 is only a placeholder here (its chunks 0,1 are masked out); segment 1 = x = [inp | motion]
        a.seg_hi[0] = u.net_h[0]; a.seg_lo[0] = u.net_h[1]; a.seg_c[0] = 128; a.seg_pitch[0] = 128;
        a.seg_hi[1] = u.x_h[0]; a.seg_lo[1] = u.x_h[1]; a.seg_c[1] = 256; a.seg_pitch[1] = 256;
        a.w_hi = hbase + d.hi_off; a.w_lo = hbase + d.lo_off; a.bias = u.zero_bias;
        a.cin_pad = d.cin_pad; a.cout_pad = d.cout_pad; a.cout = d.cout; a.n_tile = d.n_tile; a.kh = d.kh; a.kw = d.kw;
        a.B = B; a.h = h; a.w = w; a.epi = EPI_SCALE; a.scale = 1.f;
        a.out_f32 = u.pre[k]; a.out_f32_pitch = d.cout;
        a.out_tiled = 1; a.side_tiled = 1;                        // the consumers read the pre-sums as a tiled side buffer
        a.chunk_mask = 0x0Cu;                                     // chunks {2,3}: the inp channels
        a.layer_id = ids[k];
    }
    // the four GEMMs are independent: one chained launch (no dependencies) when the batch fills the machine
    const int m_tiles = B * ceil_div(h, B2P_TILE_ROWS) * ceil_div(w, B2P_TILE_COLS);
    if (b2p_conv_chain_enabled() && m_tiles >= 296 &&
        b2p_conv_chain_done_ints(4, m_tiles) * sizeof(int) <= (size_t)B * h * w * 256 * sizeof(float)) {
        B2PChainDep deps[4];
        memset(deps, 0, sizeof(deps));
        const int rcc = b2p_launch_conv_chain(args, 4, deps, nullptr, nullptr, reinterpret_cast<int*>(u.c1), s);
        if (rcc != -1) return rcc;
    }
    for (int k = 0; k < 4; ++k) {
        int rc = b2p_launch_conv_umma(args[k], s);
        if (rc) return rc;
    }
    return 0;
}

// levels 1..3 of the pyramid from level 0 (2x2 floor average pooling, corr.py:32-34)
int pool_pyramid(float* pyramid, int B, int h, int w, cudaStream_t s) {
    const int P = h * w;
    {
        const int rc3 = b2p_corr_pool3(pyramid, B, h, w, s);       // one pass over level 0 when an image fits shared memory
        if (rc3 != -1) return rc3;
    }
    float* src = pyramid;
    int hl = h, wl = w, rc;
    for (int l = 1; l < B200POSE_CORR_LEVELS; ++l) {
        float* dst = src + (size_t)B * P * hl * wl;
        if ((rc = b2p_corr_pool(src, B * P, hl, wl, dst, s))) return rc;
        src = dst; hl >>= 1; wl >>= 1;
    }
    return 0;
}

struct VolumeWs {                 // operands of the tensor-core correlation GEMM
    __half* fm_h[4];              // f1 hi, f1 lo, f2 hi, f2 lo as PXC [B*P][256]
    float* zero_bias;             // the shared epilogue adds a bias; the volume has none
};

size_t volume_ws_layout(int B, int h, int w, void* ws, size_t cap, VolumeWs* out) {
    const size_t P = (size_t)B * h * w;
    Carver c(ws, cap);
    VolumeWs v;
    for (int k = 0; k < 4; ++k) v.fm_h[k] = c.take<__half>(P * 256);
    v.zero_bias = c.take<float>((size_t)h * w + 64);
    if (out) *out = v;
    return align_up(c.off, 1024);
}

// largest MMA N (multiple of 16, <= 240) that divides P; 0 if there is none
inline int volume_ntile(int P) {
    for (int n = 240; n >= 16; n -= 16)
        if (P % n == 0) return n;
    return 0;
}

// C[b][p][q] = <f1[b,:,p], f2[b,:,q]> / 16 on the tensor cores (D = 256): the conv_umma kernel run as a 1x1 "conv"
// whose weights are the second feature map of the same sample (b_batched).
int run_corr_volume_tc(const float* fmap1, const float* fmap2, int B, int h, int w, float* level0, const VolumeWs& v,
                       cudaStream_t s) {
    const int P = h * w;
    const int nt = volume_ntile(P);
    int rc;
    if ((rc = b2p_fmap_to_pxc_half(fmap1, B, 256, P, v.fm_h[0], v.fm_h[1], s))) return rc;
    if ((rc = b2p_fmap_to_pxc_half(fmap2, B, 256, P, v.fm_h[2], v.fm_h[3], s))) return rc;
    B2P_CUDA(cudaMemsetAsync(v.zero_bias, 0, ((size_t)P + 64) * sizeof(float), s));
    UmmaConvArgs a;
    memset(&a, 0, sizeof(a));
    a.seg_hi[0] = v.fm_h[0]; a.seg_lo[0] = v.fm_h[1]; a.seg_c[0] = 256; a.seg_pitch[0] = 256;
    a.w_hi = v.fm_h[2]; a.w_lo = v.fm_h[3]; a.bias = v.zero_bias; a.b_batched = B; a.layer_id = -1;
    a.cin_pad = 256; a.cout_pad = P; a.cout = P; a.n_tile = nt; a.kh = 1; a.kw = 1;
    a.B = B; a.h = h; a.w = w; a.epi = EPI_SCALE; a.scale = 0.0625f;     // 1/sqrt(256), exact
    a.out_f32 = level0; a.out_f32_pitch = P;
    return b2p_launch_conv_umma(a, s);
}

struct RefineWs {
    float *pyr, *net, *xbuf, *corr, *coords1, *flow, *mask, *target, *weight;
    void* lm;
    void* fg;                     // foreground list of the call (b2p_fg_build)
    void* fgp;                    // foreground pipeline: channels-last descriptors + per-pixel records (fg_pipeline.cu)
    VolumeWs vol;
    UpdateWs u;
};

size_t refine_ws_layout(int B, int H, int W, void* ws, size_t cap, RefineWs* out) {
    const int h = H / 8, w = W / 8;
    const size_t P = (size_t)B * h * w, N = (size_t)B * H * W;
    Carver c(ws, cap);
    RefineWs r;
    r.pyr = c.take<float>(b200pose_pyramid_floats(B, h, w));
    r.net = c.take<float>(P * 128);
    r.xbuf = c.take<float>(P * 256);
    r.corr = c.take<float>(P * B200POSE_CORR_PITCH);
    r.coords1 = c.take<float>(P * 2);
    r.flow = c.take<float>(P * 2);
    r.mask = c.take<float>(P * 576);
    r.target = c.take<float>(N * 2);
    r.weight = c.take<float>(N);
    r.lm = c.take<char>(b2p_lm_ws_bytes(B, H, W));
    r.fg = c.take<char>(b2p_fg_ws_bytes(B, H, W));
    r.fgp = c.take<char>(b2p_fgpipe_ws_bytes(B, H, W));
    c.off = align_up(c.off, 1024);
    c.off += volume_ws_layout(B, h, w, ws ? c.base + c.off : nullptr, 0, &r.vol);
    const size_t used = update_ws_layout(B, h, w, ws ? c.base + c.off : nullptr, cap > c.off ? cap - c.off : 0, &r.u);
    if (out) *out = r;
    return c.off + used;
}

}  // namespace

extern "C" {

int b200pose_version(void) { return B200POSE_VERSION; }

const char* b200pose_error_string(int code) {
    switch (code) {
        case B200POSE_OK: return "ok";
        case B200POSE_E_NULL: return "b200pose: required pointer is NULL";
        case B200POSE_E_SHAPE: return "b200pose: unsupported shape (need B>=1, H,W multiples of 8 and >= 128)";
        case B200POSE_E_WORKSPACE: return "b200pose: workspace too small or misaligned";
        case B200POSE_E_ARG: return "b200pose: bad scalar argument";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "b200pose: unknown error";
    }
}

size_t b200pose_packed_weights_bytes(void) {
    return b2p_half_section_offset_bytes() + b2p_half_layout().total_halves * sizeof(__half);
}

int b200pose_pack_weights(const float* const* tensors_host, void* packed, void* stream) {
    if (!tensors_host || !packed) return B200POSE_E_NULL;
    if ((uintptr_t)packed & 255) return B200POSE_E_WORKSPACE;
    for (int i = 0; i < B200POSE_NUM_WEIGHT_TENSORS; ++i)
        if (!tensors_host[i]) return B200POSE_E_NULL;
    return b2p_pack_weights(tensors_host, reinterpret_cast<float*>(packed), (cudaStream_t)stream);
}

size_t b200pose_pyramid_floats(int B, int h, int w) {
    size_t per = 0;
    int hl = h, wl = w;
    for (int l = 0; l < B200POSE_CORR_LEVELS; ++l) { per += (size_t)hl * wl; hl >>= 1; wl >>= 1; }
    return (size_t)B * h * w * per;
}

int b200pose_corr_pyramid(const float* fmap1, const float* fmap2, int B, int D, int h, int w, float* pyramid,
                          void* stream) {
    if (!fmap1 || !fmap2 || !pyramid) return B200POSE_E_NULL;
    if (B < 1 || D < 1 || (h >> 3) < 2 || (w >> 3) < 2) return B200POSE_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = b2p_corr_volume(fmap1, fmap2, B, D, h * w, pyramid, s))) return rc;
    return pool_pyramid(pyramid, B, h, w, s);
}

size_t b200pose_corr_pyramid_tc_workspace_bytes(int B, int h, int w) { return volume_ws_layout(B, h, w, nullptr, 0, nullptr); }

int b200pose_corr_pyramid_tc(const float* fmap1, const float* fmap2, int B, int h, int w, float* pyramid, void* workspace,
                             size_t workspace_bytes, void* stream) {
    if (!fmap1 || !fmap2 || !pyramid || !workspace) return B200POSE_E_NULL;
    if (B < 1 || (h >> 3) < 2 || (w >> 3) < 2 || volume_ntile(h * w) == 0) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 1023) || workspace_bytes < b200pose_corr_pyramid_tc_workspace_bytes(B, h, w)) return B200POSE_E_WORKSPACE;
    VolumeWs v;
    volume_ws_layout(B, h, w, workspace, workspace_bytes, &v);
    int rc;
    if ((rc = run_corr_volume_tc(fmap1, fmap2, B, h, w, pyramid, v, (cudaStream_t)stream))) return rc;
    return pool_pyramid(pyramid, B, h, w, (cudaStream_t)stream);
}

int b200pose_corr_lookup(const float* pyramid, const float* coords, int B, int h, int w, float* out, void* stream) {
    if (!pyramid || !coords || !out) return B200POSE_E_NULL;
    if (B < 1 || (h >> 3) < 2 || (w >> 3) < 2) return B200POSE_E_SHAPE;
    return b2p_corr_lookup(pyramid, coords, B, h, w, out, nullptr, nullptr, (cudaStream_t)stream);
}

int b200pose_context_init(const float* context, int B, int H, int W, float* net, float* xbuf, void* stream) {
    if (!context || !net || !xbuf) return B200POSE_E_NULL;
    if (B < 1 || H < 16 || W < 16 || (H % 8) || (W % 8)) return B200POSE_E_SHAPE;
    return b2p_context_init(context, B, H, W, net, xbuf, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int b200pose_flow_init(const float* depth, const float* K, const float* G, int B, int H, int W, float* coords1,
                       float* flow, void* stream) {
    if (!depth || !K || !G || !coords1 || !flow) return B200POSE_E_NULL;
    if (B < 1 || H < 16 || W < 16 || (H % 8) || (W % 8)) return B200POSE_E_SHAPE;
    return b2p_flow_init(depth, K, G, B, H, W, coords1, flow, (cudaStream_t)stream);
}

size_t b200pose_update_workspace_bytes(int B, int h, int w) { return update_ws_layout(B, h, w, nullptr, 0, nullptr); }

int b200pose_update_block(const void* packed_weights, float* net, float* xbuf, const float* corr, float* coords1,
                          float* flow, float* mask, float* dflow_out, int B, int h, int w, int flags, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (!packed_weights || !net || !xbuf || !corr || !coords1 || !flow || !mask || !workspace) return B200POSE_E_NULL;
    if (B < 1 || h < 1 || w < 1) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 1023) || workspace_bytes < b200pose_update_workspace_bytes(B, h, w)) return B200POSE_E_WORKSPACE;
    UpdateWs u;
    update_ws_layout(B, h, w, workspace, workspace_bytes, &u);
    cudaStream_t s = (cudaStream_t)stream;
    const float* wts = reinterpret_cast<const float*>(packed_weights);
    if (!(flags & B200POSE_FLAG_TENSOR_CORES))
        return run_update_block(wts, net, xbuf, corr, coords1, flow, mask, dflow_out, B, h, w, u, s);
    const size_t P = (size_t)B * h * w;
    int rc;
    if ((rc = b2p_split_planes(corr, B200POSE_CORR_PITCH, B200POSE_CORR_PITCH, P, u.corr_h[0], u.corr_h[1], B200POSE_CORR_PITCH, s))) return rc;
    if ((rc = b2p_split_planes(net, 128, 128, P, u.net_h[0], u.net_h[1], 128, s))) return rc;
    if ((rc = b2p_split_planes(xbuf, 256, 128, P, u.x_h[0], u.x_h[1], 256, s))) return rc;
    // the tensor-core epilogues keep the fp32 hidden state in the tiled side-buffer layout (u.rhbuf)
    if ((rc = b2p_pxc_to_tiled(net, u.rhbuf, B, h, w, 128, s))) return rc;
    if ((rc = b2p_pxc_to_tiled(net, u.hbuf_x, B, h, w, 128, s, 1))) return rc;
    if ((rc = run_update_block_tc(wts, net, coords1, flow, mask, dflow_out, B, h, w, u, false, s))) return rc;
    return b2p_tiled_to_pxc(u.rhbuf, net, B, h, w, 128, s);
}

size_t b200pose_conv_layer_workspace_bytes(int B, int h, int w) { return (size_t)B * h * w * 384 * 2 * sizeof(__half) + 4096; }

int b200pose_conv_layer_info(int layer, int* cin0, int* cin1, int* cout, int* kh, int* kw) {
    if (layer < 0 || layer >= CV_COUNT || !cin0 || !cin1 || !cout || !kh || !kw) return B200POSE_E_ARG;
    const B2PConvDesc& d = b2p_weight_layout().cv[layer];
    const bool two = (layer == CV_ZR1 || layer == CV_Q1 || layer == CV_ZR2 || layer == CV_Q2);
    *cin0 = two ? 128 : d.cin; *cin1 = two ? 256 : 0; *cout = d.cout; *kh = d.kh; *kw = d.kw;
    return 0;
}

int b200pose_conv_layer(const void* packed_weights, int layer, const float* in0, int pitch0, const float* in1, int pitch1,
                        float* out, int B, int h, int w, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    if (!packed_weights || !in0 || !out || !workspace) return B200POSE_E_NULL;
    if (layer < 0 || layer >= CV_COUNT) return B200POSE_E_ARG;
    if (B < 1 || h < 1 || w < 1) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 1023) || workspace_bytes < b200pose_conv_layer_workspace_bytes(B, h, w)) return B200POSE_E_WORKSPACE;
    int c0, c1, cout, kh, kw;
    b200pose_conv_layer_info(layer, &c0, &c1, &cout, &kh, &kw);
    if (c1 > 0 && !in1) return B200POSE_E_NULL;
    if (pitch0 < c0 || (pitch0 & 3) || (c1 > 0 && (pitch1 < c1 || (pitch1 & 3)))) return B200POSE_E_ARG;   // float4 rows
    cudaStream_t s = (cudaStream_t)stream;
    const float* wts = reinterpret_cast<const float*>(packed_weights);
    const B2PWeightLayout& L = b2p_weight_layout();
    const B2PConvDesc& d = L.cv[layer];
    if (!(flags & B200POSE_FLAG_TENSOR_CORES)) {
        ConvParams p;
        memset(&p, 0, sizeof(p));
        p.src0 = in0; p.pitch0 = pitch0; p.c0 = c0; p.src1 = in1; p.pitch1 = pitch1; p.c1 = c1;
        p.wgt = wts + d.w_off; p.bias = wts + d.b_off; p.dst = out; p.dst_pitch = (cout + 3) / 4 * 4;
        p.cout = d.cout; p.cout_pad = d.cout_pad; p.cin_pad = d.cin_pad;
        p.B = B; p.h = h; p.w = w; p.kh = d.kh; p.kw = d.kw; p.epi = EPI_NONE; p.scale = 1.f;
        return b2p_launch_conv(p, s);
    }
    const size_t P = (size_t)B * h * w;
    const int c0p = (c0 + 7) / 8 * 8, c1p = (c1 + 7) / 8 * 8;       // 16-byte row pitch for the TMA map
    __half* base = reinterpret_cast<__half*>(workspace);
    __half* h0[2] = {base, base + P * c0p};
    __half* h1[2] = {base + 2 * P * c0p, base + 2 * P * c0p + P * c1p};
    int rc;
    B2P_CUDA(cudaMemsetAsync(base, 0, (2 * P * c0p + 2 * P * c1p) * sizeof(__half), s));
    if ((rc = b2p_split_planes(in0, pitch0, c0, P, h0[0], h0[1], c0p, s))) return rc;
    if (c1 && (rc = b2p_split_planes(in1, pitch1, c1, P, h1[0], h1[1], c1p, s))) return rc;
    const B2PHalfConvDesc& hd = b2p_half_layout().cv[layer];
    const __half* hbase = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(wts) + b2p_half_section_offset_bytes());
    UmmaConvArgs a;
    memset(&a, 0, sizeof(a));
    a.seg_hi[0] = h0[0]; a.seg_lo[0] = h0[1]; a.seg_c[0] = c0p; a.seg_pitch[0] = c0p;
    if (c1) { a.seg_hi[1] = h1[0]; a.seg_lo[1] = h1[1]; a.seg_c[1] = c1p; a.seg_pitch[1] = c1p; }
    a.w_hi = hbase + hd.hi_off; a.w_lo = hbase + hd.lo_off; a.bias = wts + d.b_off;
    a.cin_pad = hd.cin_pad; a.cout_pad = hd.cout_pad; a.cout = hd.cout; a.n_tile = hd.n_tile; a.kh = hd.kh; a.kw = hd.kw;
    a.B = B; a.h = h; a.w = w; a.epi = EPI_SCALE; a.scale = 1.f; a.out_f32 = out; a.out_f32_pitch = (cout + 3) / 4 * 4;
    a.layer_id = layer;
    return b2p_launch_conv_umma(a, s);
}

int b200pose_upsample_weight(const float* flow, const float* mask, const float* geofea1, const float* geofea2,
                             const float* depth, float sigma, int B, int C, int H, int W, float* flow_up, float* target,
                             float* weight, void* stream) {
    if (!flow || !mask) return B200POSE_E_NULL;
    if (weight && (!geofea1 || !geofea2 || !depth)) return B200POSE_E_NULL;
    if (B < 1 || H < 8 || W < 8 || (H % 8) || (W % 8) || (weight && C < 1)) return B200POSE_E_SHAPE;
    return b2p_upsample_weight(flow, mask, geofea1, geofea2, depth, sigma, B, C, H, W, flow_up, target, weight, 0,
                               (cudaStream_t)stream);
}

size_t b200pose_encoder_packed_weights_bytes(void) { return b2p_encoder_packed_bytes(); }

int b200pose_encoder_pack_weights(const float* const* tensors_host, void* packed, void* stream) {
    if (!tensors_host || !packed) return B200POSE_E_NULL;
    if ((uintptr_t)packed & 1023) return B200POSE_E_WORKSPACE;
    for (int i = 0; i < B200POSE_NUM_ENCODER_TENSORS; ++i)
        if (!tensors_host[i]) return B200POSE_E_NULL;
    return b2p_encoder_pack(tensors_host, packed, (cudaStream_t)stream);
}

size_t b200pose_encoder_workspace_bytes(int B, int H, int W) { return b2p_encoder_ws_bytes(B, H, W); }

int b200pose_image_encoder(const void* packed_weights, const float* image1, const float* image2, int B, int H, int W, float* fmap1,
                           float* fmap2, void* workspace, size_t workspace_bytes, void* stream) {
    if (!packed_weights || !image1 || !image2 || !fmap1 || !fmap2 || !workspace) return B200POSE_E_NULL;
    if (B < 1 || H < 16 || W < 16 || (H % 8) || (W % 8)) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 1023) || workspace_bytes < b2p_encoder_ws_bytes(B, H, W)) return B200POSE_E_WORKSPACE;
    return b2p_image_encoder(packed_weights, image1, image2, B, H, W, fmap1, fmap2, workspace, (cudaStream_t)stream);
}

size_t b200pose_zoom_crop_workspace_bytes(int B) { return b2p_zoom_crop_ws_bytes(B); }

int b200pose_zoom_crop(const float* pc_depth, const float* K, const float* T, const float* image, const float* geofea, int B, int Ci,
                       int Cg, int H, int W, int Hc, int Wc, float margin_ratio, int flags, float* image_crop, float* geofea_crop,
                       float* K_crop, float* theta, void* workspace, size_t workspace_bytes, void* stream) {
    if (!pc_depth || !K || !T || !workspace) return B200POSE_E_NULL;
    if ((image != nullptr) != (image_crop != nullptr) || (geofea != nullptr) != (geofea_crop != nullptr)) return B200POSE_E_NULL;
    if (B < 1 || H < 2 || W < 2 || Hc < 2 || Wc < 2 || (image && Ci < 1) || (geofea && Cg < 1)) return B200POSE_E_SHAPE;
    const int cl = (flags & B200POSE_ZOOM_GEO_CHANNELS_LAST) ? 1 : 0;
    if (cl && geofea && Cg != 32) return B200POSE_E_SHAPE;
    if (!(margin_ratio >= 0.f)) return B200POSE_E_ARG;
    if (((uintptr_t)workspace & 255) || workspace_bytes < b2p_zoom_crop_ws_bytes(B)) return B200POSE_E_WORKSPACE;
    return b2p_zoom_crop(pc_depth, K, T, image, geofea, B, Ci, Cg, H, W, Hc, Wc, margin_ratio, cl, image_crop, geofea_crop, K_crop,
                         theta, workspace, (cudaStream_t)stream);
}

size_t b200pose_pose_metrics_workspace_bytes(int B, int n_pts) { return b2p_pose_metrics_ws_bytes(B, n_pts); }

int b200pose_pose_metrics(const float* T_pred, const float* T_gt, const float* pts, const float* diameter, const float* K,
                          int B, int n_pts, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!T_pred || !T_gt || !pts || !diameter || !K || !out || !workspace) return B200POSE_E_NULL;
    if (B < 1 || B > 65535 || n_pts < 1) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 255) || workspace_bytes < b2p_pose_metrics_ws_bytes(B, n_pts)) return B200POSE_E_WORKSPACE;
    return b2p_pose_metrics(T_pred, T_gt, pts, diameter, K, B, n_pts, out, workspace, (cudaStream_t)stream);
}

size_t b200pose_lm_workspace_bytes(int B, int H, int W) { return b2p_lm_ws_bytes(B, H, W); }

int b200pose_lm_solve(const float* depth, const float* target, const float* weight, const float* K, float* G, int B,
                      int H, int W, float depth_offset, int n_steps, double ep_lmbda, double lm_lmbda, double* H_out,
                      double* b_out, float* delta_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!depth || !target || !weight || !K || !G || !workspace) return B200POSE_E_NULL;
    if (B < 1 || H < 1 || W < 1) return B200POSE_E_SHAPE;
    if (n_steps < 0) return B200POSE_E_ARG;
    if (((uintptr_t)workspace & 255) || workspace_bytes < b2p_lm_ws_bytes(B, H, W)) return B200POSE_E_WORKSPACE;
    { int rc0 = b2p_lm_reset(workspace, B, H, W, (cudaStream_t)stream); if (rc0) return rc0; }
    if (!H_out && !b_out && !delta_out)      // no taps requested: all steps in one launch
        return b2p_lm_steps(depth, target, weight, K, G, B, H, W, depth_offset, ep_lmbda, lm_lmbda, n_steps, workspace,
                            (cudaStream_t)stream);
    for (int i = 0; i < n_steps; ++i) {
        int rc = b2p_lm_step(depth, target, weight, K, G, B, H, W, depth_offset, ep_lmbda, lm_lmbda,
                             H_out ? H_out + (size_t)i * B * 36 : nullptr, b_out ? b_out + (size_t)i * B * 6 : nullptr,
                             delta_out ? delta_out + (size_t)i * B * 6 : nullptr, workspace, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

size_t b200pose_lm_backward_workspace_bytes(int B, int H, int W) { return b2p_lm_bwd_ws_bytes(B, H, W); }

int b200pose_lm_backward(const float* depth, const float* target, const float* weight, const float* K, const float* G,
                         const float* grad_delta, int B, int H, int W, float depth_offset, double ep_lmbda, double lm_lmbda,
                         float* grad_target, float* grad_weight, void* workspace, size_t workspace_bytes, void* stream) {
    if (!depth || !target || !weight || !K || !G || !grad_delta || !grad_target || !grad_weight || !workspace) return B200POSE_E_NULL;
    if (B < 1 || H < 1 || W < 1) return B200POSE_E_SHAPE;
    if (((uintptr_t)workspace & 255) || workspace_bytes < b2p_lm_bwd_ws_bytes(B, H, W)) return B200POSE_E_WORKSPACE;
    return b2p_lm_backward(depth, target, weight, K, G, grad_delta, B, H, W, depth_offset, ep_lmbda, lm_lmbda, grad_target, grad_weight,
                           workspace, (cudaStream_t)stream);
}

int b200pose_debug_set_conv_events(void* ev_start, void* ev_stop) {
    g_conv_ev[0] = (cudaEvent_t)ev_start; g_conv_ev[1] = (cudaEvent_t)ev_stop;
    return 0;
}

int b200pose_cholesky_solve(const double* H, const double* b, float* x, int B, void* stream) {
    if (!H || !b || !x) return B200POSE_E_NULL;
    if (B < 1) return B200POSE_E_SHAPE;
    return b2p_chol_solve(H, b, x, B, (cudaStream_t)stream);
}

int b200pose_se3_retract(const float* delta, float* G, int B, void* stream) {
    if (!delta || !G) return B200POSE_E_NULL;
    if (B < 1) return B200POSE_E_SHAPE;
    return b2p_se3_retract(delta, G, B, (cudaStream_t)stream);
}

size_t b200pose_refine_workspace_bytes(int B, int H, int W) { return refine_ws_layout(B, H, W, nullptr, 0, nullptr); }

int b200pose_refine_launch_count(int B, int H, int W, int n_iters, int n_lm) {
    // Kernels b200pose_refine_iters enqueues on the tensor-core path with C_geo = 32 and the current options:
    //   per call: LM counter reset, 2 feature-map transposes, volume GEMM, pooling (1 pass, or 3), context init, hidden state to
    //   the two tiled layouts; with n_iters > 0: 3 for the foreground list (+ 2 for the pipeline's channels-last descriptors);
    //   with n_iters > 1: the 4 GRU partial-sum GEMMs (1 chained launch, or 4);  per recurrent iteration: flow_init, lookup, im2col, the convolutions
    //   (1 chained launch incl. the flow head, or 11 + 2), upsample/target/weight (2 with the pipeline, else 1), one launch
    //   for all LM steps
    if (!shape_ok(B, H, W)) return 0;
    const int h = H / 8, w = W / 8;
    const B2POptions& o = b2p_options();
    const bool chain = (o.conv_mode & 16) && B * ceil_div(h, B2P_TILE_ROWS) * ceil_div(w, B2P_TILE_COLS) >= 296;
    const bool pool3 = o.pool_mode == 1 && (size_t)(h * w + (h >> 1) * (w >> 1) + (h >> 2) * (w >> 2)) * sizeof(float) <= 40 * 1024;
    const bool pipe = o.fg_list != 0 && o.fg_pipeline == 1;      // (2 = only with channels-last geofea2, which this count does not assume)
    const int per_call = 4 + (pool3 ? 1 : 3) + 3 + (n_iters > 0 && o.fg_list ? 3 + (pipe ? 2 : 0) : 0) + (n_iters > 1 ? (chain ? 1 : 4) : 0);
    const int per_iter = 2 + 1 + (chain ? 1 : 13) + (pipe ? 2 : 1) + (n_lm > 0 ? 1 : 0);
    return per_call + n_iters * per_iter;
}

}  // extern "C"

// g2_far / g2_window (host entry only): geofea2 holds valid data only inside the per-sample window [B][4] = (y0, y1, x0, x1);
// samples outside it are read from g2_far, the same map in mapped host memory (upsample_weight_pixel<WINDOW>).
static int refine_iters_impl(const void* packed_weights, const float* fmap1, const float* fmap2, const float* context,
                             const float* geofea1, const float* geofea2, const float* depth, const float* K, float* G,
                             float sigma, int B, int C_geo, int H, int W, int n_iters, int n_lm, double ep_lmbda,
                             double lm_lmbda, int flags, float* flow_first, float* flow_last, float* weight_last,
                             void* workspace, size_t workspace_bytes, void* stream, const float* g2_far, const int* g2_window,
                             const float* ctx_texels, int ctx_split);

extern "C" {

int b200pose_refine_iters(const void* packed_weights, const float* fmap1, const float* fmap2, const float* context,
                          const float* geofea1, const float* geofea2, const float* depth, const float* K, float* G,
                          float sigma, int B, int C_geo, int H, int W, int n_iters, int n_lm, double ep_lmbda,
                          double lm_lmbda, int flags, float* flow_first, float* flow_last, float* weight_last,
                          void* workspace, size_t workspace_bytes, void* stream) {
    return refine_iters_impl(packed_weights, fmap1, fmap2, context, geofea1, geofea2, depth, K, G, sigma, B, C_geo, H, W, n_iters,
                             n_lm, ep_lmbda, lm_lmbda, flags, flow_first, flow_last, weight_last, workspace, workspace_bytes, stream,
                             nullptr, nullptr, (flags & B200POSE_FLAG_CONTEXT_TEXELS) ? context : nullptr,
                             (flags & B200POSE_FLAG_CONTEXT_TEXELS) ? 256 : 0);
}

}  // extern "C"

static int refine_iters_impl(const void* packed_weights, const float* fmap1, const float* fmap2, const float* context,
                             const float* geofea1, const float* geofea2, const float* depth, const float* K, float* G,
                             float sigma, int B, int C_geo, int H, int W, int n_iters, int n_lm, double ep_lmbda,
                             double lm_lmbda, int flags, float* flow_first, float* flow_last, float* weight_last,
                             void* workspace, size_t workspace_bytes, void* stream, const float* g2_far, const int* g2_window,
                             const float* ctx_texels, int ctx_split) {
    if (!packed_weights || !fmap1 || !fmap2 || !context || !geofea1 || !geofea2 || !depth || !K || !G || !workspace)
        return B200POSE_E_NULL;
    if (!shape_ok(B, H, W) || C_geo < 1) return B200POSE_E_SHAPE;
    if (n_iters < 0 || n_lm < 0 || !(sigma > 0.f)) return B200POSE_E_ARG;
    if (((uintptr_t)workspace & 1023) || workspace_bytes < b200pose_refine_workspace_bytes(B, H, W)) return B200POSE_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int h = H / 8, w = W / 8;
    const bool tc = (flags & B200POSE_FLAG_TENSOR_CORES) != 0;
    const float* wts = reinterpret_cast<const float*>(packed_weights);
    RefineWs r;
    refine_ws_layout(B, H, W, workspace, workspace_bytes, &r);
    const UpdateWs& u = r.u;
    int rc;
    // update_corr_fn == True part (CFNet.py:115-133): pyramid + hidden-state reset, once per render iteration
    if ((rc = b2p_lm_reset(r.lm, B, H, W, s))) return rc;
    if (tc && volume_ntile(h * w) > 0) {
        if ((rc = run_corr_volume_tc(fmap1, fmap2, B, h, w, r.pyr, r.vol, s))) return rc;
        if ((rc = pool_pyramid(r.pyr, B, h, w, s))) return rc;
    } else if ((rc = b200pose_corr_pyramid(fmap1, fmap2, B, 256, h, w, r.pyr, stream))) return rc;
    // ctx_texels / ctx_split: the first ctx_split context planes of every object come as texels (B200POSE_FLAG_CONTEXT_TEXELS:
    // all 256), the rest from the full map
    if (tc) rc = b2p_context_init(context, B, H, W, r.net, nullptr, u.net_h[0], u.net_h[1], u.x_h[0], u.x_h[1], s, ctx_texels, ctx_split);
    else rc = b2p_context_init(context, B, H, W, r.net, r.xbuf, nullptr, nullptr, nullptr, nullptr, s, ctx_texels, ctx_split);
    if (rc) return rc;
    if (tc && (rc = b2p_pxc_to_tiled(r.net, u.rhbuf, B, h, w, 128, s))) return rc;     // hidden state of the tensor-core epilogues
    if (tc && (rc = b2p_pxc_to_tiled(r.net, u.hbuf_x, B, h, w, 128, s, 1))) return rc;  // ... and its x-major copy (chained launch)
    if (tc && n_iters > 1 && (rc = run_gru_precompute(wts, B, h, w, u, s))) return rc;
    // The rendered depth is fixed over the recurrent iterations: compact its foreground once and run the LM steps and the
    // upsample + weight kernel over the list (option fg_list = 0: dense kernels).
    const bool use_fg = n_iters > 0 && fg_list_enabled();
    const bool g2_cl = (flags & B200POSE_FLAG_GEO2_CHANNELS_LAST) != 0;
    // foreground pipeline (fg_pipeline.cu + lm_cluster_kernel): channels-last descriptors, one record per listed pixel
    // option fg_pipeline: 0 off, 1 on, 2 (default) only when geofea2 arrives channels-last (zoom-crop output): with NCHW input the
    // two transposes (140 us per call at B=32) outweigh the 16 us per iteration the pipeline saves (profiles/r2d)
    const int pipe_opt = b2p_options().fg_pipeline;
    const bool windowed = g2_far != nullptr && g2_window != nullptr;      // only the dense kernel knows the window
    if (windowed && g2_cl) return B200POSE_E_ARG;
    const bool use_pipe = !windowed && use_fg && C_geo == 32 && (pipe_opt == 1 || (pipe_opt == 2 && g2_cl));
    if (g2_cl && !use_pipe) return B200POSE_E_ARG;          // only the pipeline reads channels-last descriptors
    const bool fg_up = !windowed && use_fg && !use_pipe && fg_upsample_enabled();
    if (use_fg && (rc = b2p_fg_build(depth, B, H, W, r.fg, fg_up ? r.target : nullptr, fg_up ? r.weight : nullptr, s))) return rc;
    if (use_pipe && (rc = b2p_fgpipe_prepare(geofea1, geofea2, g2_cl ? 1 : 0, B, H, W, r.fg, r.fgp, s))) return rc;
    const int* fg_idx = use_fg ? b2p_fg_idx(r.fg) : nullptr;
    const int* fg_count = use_fg ? b2p_fg_count(r.fg, B, H, W) : nullptr;
    for (int it = 0; it < n_iters; ++it) {
        if ((rc = b2p_flow_init(depth, K, G, B, H, W, r.coords1, r.flow, s))) return rc;
        if (tc) {
            if ((rc = b2p_corr_lookup(r.pyr, r.coords1, B, h, w, nullptr, u.corr_h[0], u.corr_h[1], s))) return rc;
            if ((rc = run_update_block_tc(wts, r.net, r.coords1, r.flow, r.mask, nullptr, B, h, w, u, n_iters > 1, s))) return rc;
        } else {
            if ((rc = b2p_corr_lookup(r.pyr, r.coords1, B, h, w, r.corr, nullptr, nullptr, s))) return rc;
            if ((rc = run_update_block(wts, r.net, r.xbuf, r.corr, r.coords1, r.flow, r.mask, nullptr, B, h, w, u, s))) return rc;
        }
        float* fu = (it == 0 && flow_first) ? flow_first : ((it == n_iters - 1) ? flow_last : nullptr);
        if (use_pipe) {
            // dense outputs on request: the up-sampled flow of every pixel from the dense kernel (no descriptors involved),
            // the weight map as a scatter of the records' weights over a zeroed map
            if (fu && (rc = b2p_upsample_weight(r.flow, r.mask, nullptr, nullptr, nullptr, sigma, B, C_geo, H, W, fu, nullptr, nullptr, 0, s)))
                return rc;
            float* wd = (weight_last && it == n_iters - 1) ? weight_last : nullptr;
            if (wd) B2P_CUDA(cudaMemsetAsync(wd, 0, (size_t)B * H * W * sizeof(float), s));
            if ((rc = b2p_fgpipe_upsample_weight(r.flow, r.mask, g2_cl ? geofea2 : nullptr, depth, sigma, B, H, W, r.fg, r.fgp, wd, s)))
                return rc;
        } else if (fg_up && !fu) {      // persistent kernel over the foreground list (the up-sampled flow is not an output here)
            if ((rc = b2p_upsample_weight_fg(r.flow, r.mask, geofea1, geofea2, depth, sigma, B, C_geo, H, W, r.fg, r.target, r.weight, s)))
                return rc;
        } else if ((rc = b2p_upsample_weight(r.flow, r.mask, geofea1, geofea2, depth, sigma, B, C_geo, H, W, fu, r.target,
                                             r.weight, 1, s, g2_far, g2_window))) return rc;
        if (it == 0 && it == n_iters - 1 && flow_first && flow_last)
            B2P_CUDA(cudaMemcpyAsync(flow_last, flow_first, (size_t)B * 2 * H * W * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (use_pipe) {
            if ((rc = b2p_lm_cluster(b2p_fgpipe_records(r.fgp, B, H, W), nullptr, nullptr, nullptr, fg_idx, fg_count, K, G, B, H, W, 1e-5f, ep_lmbda, lm_lmbda, n_lm, s)))
                return rc;
        } else if ((rc = b2p_lm_steps(depth, r.target, r.weight, K, G, B, H, W, 1e-5f, ep_lmbda, lm_lmbda, n_lm, r.lm, s, fg_idx, fg_count))) return rc;
    }
    if (use_pipe) return 0;                                 // weight_last was written by the last iteration's scatter
    if (weight_last && n_iters > 0)
        B2P_CUDA(cudaMemcpyAsync(weight_last, r.weight, (size_t)B * H * W * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

namespace {
// Device staging of the host entry.  The batch is processed in sub-batches so that the PCIe transfer of sub-batch
// k+1 (on an internal copy stream) overlaps the kernels of sub-batch k (on the caller's stream); per-sample results
// do not depend on how the batch is split (tests/test_gpu_refine.py checks this bit for bit).
struct HostStage {
    float *fmap1, *fmap2, *context, *geo1, *geo2, *depth, *K, *G;
    int* win;                 // [bs][4] geofea2 window per sample (y0, y1, x0, x1)
};
struct HostScratch {
    HostStage st[2];          // ping-pong input staging for sub-batches
    void* ws; size_t ws_bytes;
    int bs;                   // samples per sub-batch
};
// Host-side gather of the context texels the 1/8 resample reads (4 of every 64 floats), by a few worker threads that live
// for one call.  Pure data movement: the interpolation itself stays in context_init_kernel<PACKED>.  Sub-batch k's texels
// are complete when done[k] == T; the workers never wait (the staging buffer holds the whole batch).
struct TexelGather {
    const float* ctx = nullptr; float* out = nullptr;
    int B = 0, H = 0, W = 0, h = 0, w = 0, bs = 0, nsub = 0, T = 0, mode = 0, planes_per = 256;   // planes_per: gathered planes per object
    std::vector<int> x0, x1, y0, y1;
    std::unique_ptr<std::atomic<int>[]> done;
    std::vector<std::thread> workers;

    void start(const float* ctx_, float* out_, int B_, int H_, int W_, int bs_, int nsub_, int threads, int planes_per_ = 256) {
        ctx = ctx_; out = out_; B = B_; H = H_; W = W_; h = H / 8; w = W / 8; bs = bs_; nsub = nsub_; planes_per = planes_per_;
        mode = b2p_options().host_gather;
        if (threads <= 0) {                          // auto: half of the CPUs this process may run on, at most 8
            cpu_set_t set; CPU_ZERO(&set);
            int n = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            threads = n / 2;
        }
        T = threads < 1 ? 1 : (threads > 16 ? 16 : threads);
        x0.resize(w); x1.resize(w); y0.resize(h); y1.resize(h);
        b2p_context_sample_taps(W, w, x0.data(), x1.data());
        b2p_context_sample_taps(H, h, y0.data(), y1.data());
        done.reset(new std::atomic<int>[nsub]);
        for (int k = 0; k < nsub; ++k) done[k].store(0, std::memory_order_relaxed);
        for (int t = 0; t < T; ++t) {
            try {
                workers.emplace_back([this, t] { run(t); });
            } catch (...) {                          // no more threads to be had: the caller's thread does the remaining shares
                for (int tt = t; tt < T; ++tt) run(tt);
                break;
            }
        }
    }
    void run(int t) const {
        const size_t P = (size_t)h * w;
        for (int k = 0; k < nsub; ++k) {
            const int b0 = k * bs, nb = (B - b0 < bs) ? (B - b0) : bs;
            const long planes = (long)nb * planes_per, lo = planes * t / T, hi = planes * (t + 1) / T;
            for (long pl = lo; pl < hi; ++pl) {
                const long bi = pl / planes_per, ci = pl - bi * planes_per;
                const float* src = ctx + ((size_t)(b0 + bi) * 256 + ci) * (size_t)H * W;
                float* dst = out + ((size_t)b0 * planes_per + pl) * P * 4;
                for (int y = 0; y < h; ++y) {
                    const float* r0 = src + (size_t)y0[y] * W;
                    const float* r1 = src + (size_t)y1[y] * W;
                    if ((mode & 1) && (y + 1 < h || pl + 1 < hi)) {   // the two source rows of the next output row (or of the next plane)
                        const float* n0 = y + 1 < h ? src + (size_t)y0[y + 1] * W : src + (size_t)H * W + (size_t)y0[0] * W;
                        const float* n1 = y + 1 < h ? src + (size_t)y1[y + 1] * W : src + (size_t)H * W + (size_t)y1[0] * W;
                        for (int x = 0; x < W; x += 16) { _mm_prefetch((const char*)(n0 + x), _MM_HINT_T0); _mm_prefetch((const char*)(n1 + x), _MM_HINT_T0); }
                    }
                    if (mode & 2)
                        for (int x = 0; x < w; ++x, dst += 4)
                            _mm_store_ps(dst, _mm_set_ps(r1[x1[x]], r1[x0[x]], r0[x1[x]], r0[x0[x]]));
                    else
                        for (int x = 0; x < w; ++x, dst += 4)          // streaming store: the CPU never reads the texels back
                            _mm_stream_ps(dst, _mm_set_ps(r1[x1[x]], r1[x0[x]], r0[x1[x]], r0[x0[x]]));
                }
            }
            _mm_sfence();
            done[k].fetch_add(1, std::memory_order_release);
        }
    }
    void wait(int k) const {
        while (done[k].load(std::memory_order_acquire) < T) std::this_thread::yield();
    }
    void join() {
        for (auto& th : workers) th.join();
        workers.clear();
    }
};

// The second descriptor map is only sampled at flow targets of foreground pixels: the box of depth > 0 plus a margin, columns
// rounded to 32-byte groups.  Anything sampled outside is fetched from the mapped host buffer by the kernel, so the margin is a
// performance knob, not a correctness one.
void geo2_window(const float* depth, int H, int W, int margin, int* win /*y0,y1,x0,x1*/) {
    std::vector<float> colmax((size_t)W, 0.f);
    int y0 = H, y1 = -1;
    for (int y = 0; y < H; ++y) {
        const float* row = depth + (size_t)y * W;
        float m = 0.f;
        for (int x = 0; x < W; ++x) { const float v = row[x]; m = v > m ? v : m; colmax[x] = v > colmax[x] ? v : colmax[x]; }
        if (m > 0.f) { if (y < y0) y0 = y; y1 = y; }
    }
    if (y1 < 0) { win[0] = win[1] = win[2] = win[3] = 0; return; }
    int x0 = 0, x1 = W - 1;
    while (x0 < W && !(colmax[x0] > 0.f)) ++x0;
    while (x1 > x0 && !(colmax[x1] > 0.f)) --x1;
    win[0] = y0 - margin > 0 ? y0 - margin : 0;
    win[1] = y1 + 1 + margin < H ? y1 + 1 + margin : H;
    const int a = x0 - margin > 0 ? x0 - margin : 0, b = x1 + 1 + margin < W ? x1 + 1 + margin : W;
    win[2] = a & ~7;
    win[3] = ((b + 7) & ~7) < W ? ((b + 7) & ~7) : W;
}

inline int host_sub_batch(int B) { return B >= 8 ? (B + 3) / 4 : (B >= 2 ? (B + 1) / 2 : 1); }

// stage_context: 0 = read in place from mapped host memory, 1 = full map, 2 = the four texels per low-res sample
size_t host_scratch_layout(int B, int C, int H, int W, int stage_context, void* p, size_t cap, HostScratch* out) {
    const int h = H / 8, w = W / 8;
    const int bs = host_sub_batch(B);
    Carver c(p, cap);
    HostScratch hs;
    hs.bs = bs;
    for (int k = 0; k < 2; ++k) {
        HostStage& st = hs.st[k];
        st.fmap1 = c.take<float>((size_t)bs * 256 * h * w);
        st.fmap2 = c.take<float>((size_t)bs * 256 * h * w);
        st.context = stage_context == 1 ? c.take<float>((size_t)bs * 256 * H * W)
                   : stage_context == 2 ? c.take<float>((size_t)bs * 256 * h * w * 4) : nullptr;
        st.geo1 = c.take<float>((size_t)bs * C * H * W);
        st.geo2 = c.take<float>((size_t)bs * C * H * W);
        st.depth = c.take<float>((size_t)bs * H * W);
        st.K = c.take<float>((size_t)bs * 9);
        st.G = c.take<float>((size_t)bs * 16);
        st.win = c.take<int>((size_t)bs * 4);
    }
    hs.ws_bytes = refine_ws_layout(bs, H, W, nullptr, 0, nullptr);
    hs.ws = c.take<char>(hs.ws_bytes);
    if (out) *out = hs;
    return align_up(c.off, 1024);
}
}  // namespace

extern "C" {

size_t b200pose_refine_host_scratch_bytes(int B, int C_geo, int H, int W) {
    return host_scratch_layout(B, C_geo, H, W, 1, nullptr, 0, nullptr);     // worst case: context staged too
}

size_t b200pose_refine_host_staging_bytes(int B, int H, int W) {
    if (B < 1 || H < 8 || W < 8) return 0;
    return (size_t)B * 256 * (H / 8) * (W / 8) * 4 * sizeof(float);
}

int b200pose_context_gather_texels(const float* context_host, int B, int H, int W, float* texels_host, int n_host_threads) {
    if (!context_host || !texels_host) return B200POSE_E_NULL;
    if (B < 1 || H < 16 || W < 16 || ((uintptr_t)texels_host & 15)) return B200POSE_E_SHAPE;
    TexelGather tg;
    tg.start(context_host, texels_host, B, H, W, B, 1, n_host_threads);
    tg.wait(0);
    tg.join();
    return 0;
}

int b200pose_refine_iters_host(const void* packed_weights, const float* fmap1_host, const float* fmap2_host,
                               const float* context_host, const float* geofea1_host, const float* geofea2_host,
                               const float* depth_host, const float* K_host, float* G_host, float sigma, int B,
                               int C_geo, int H, int W, int n_iters, int n_lm, double ep_lmbda, double lm_lmbda,
                               int flags, void* device_scratch, size_t device_scratch_bytes, void* stream) {
    return b200pose_refine_iters_host2(packed_weights, fmap1_host, fmap2_host, context_host, geofea1_host, geofea2_host,
                                       depth_host, K_host, G_host, sigma, B, C_geo, H, W, n_iters, n_lm, ep_lmbda, lm_lmbda,
                                       flags, device_scratch, device_scratch_bytes, nullptr, 0, 0, stream);
}

int b200pose_refine_iters_host2(const void* packed_weights, const float* fmap1_host, const float* fmap2_host,
                                const float* context_host, const float* geofea1_host, const float* geofea2_host,
                                const float* depth_host, const float* K_host, float* G_host, float sigma, int B,
                                int C_geo, int H, int W, int n_iters, int n_lm, double ep_lmbda, double lm_lmbda,
                                int flags, void* device_scratch, size_t device_scratch_bytes, void* host_staging,
                                size_t host_staging_bytes, int n_host_threads, void* stream) {
    if (!packed_weights || !fmap1_host || !fmap2_host || !context_host || !geofea1_host || !geofea2_host ||
        !depth_host || !K_host || !G_host || !device_scratch)
        return B200POSE_E_NULL;
    if (!shape_ok(B, H, W) || C_geo < 1) return B200POSE_E_SHAPE;
    if (flags & B200POSE_FLAG_CONTEXT_TEXELS) return B200POSE_E_SHAPE;          // the host entry takes the full map
    if (((uintptr_t)device_scratch & 1023) || device_scratch_bytes < b200pose_refine_host_scratch_bytes(B, C_geo, H, W))
        return B200POSE_E_WORKSPACE;
    const bool gather = host_staging != nullptr;
    if (gather && (((uintptr_t)host_staging & 15) || host_staging_bytes < b200pose_refine_host_staging_bytes(B, H, W)))
        return B200POSE_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;

    // The loop only ever touches the context map at the 4 texels around each 1/8-resolution sample (CFNet.py:129): 1/16 of
    // its floats, 2 of every 8 rows.  Two ways not to copy all B*256*H*W floats:
    //  * host_staging given: worker threads gather those texels into the caller's (pinned) staging buffer, one sub-batch
    //    ahead of the copies, and only the texels cross PCIe (B200POSE_FLAG_CONTEXT_TEXELS layout);
    //  * else, when context_host is pinned (device-accessible through UVA), the context-init kernel reads the rows straight
    //    from host memory (every 32-byte sector of 2 rows in 8 is touched: 1/4 of the bytes).
    const float* ctx_mapped = nullptr;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, context_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            ctx_mapped = reinterpret_cast<const float*>(attr.devicePointer);
        else
            (void)cudaGetLastError();
    }
    // Likewise the first descriptor map is only read where the rendered depth is positive (upsample_weight.cu): from a
    // pinned buffer only those pixels are fetched (B200POSE_SPARSE_G1=0: plain copy).
    const float* g1_mapped = nullptr;
    {
        cudaPointerAttributes attr;
        if (b2p_options().sparse_g1 != 0 && cudaPointerGetAttributes(&attr, geofea1_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            g1_mapped = reinterpret_cast<const float*>(attr.devicePointer);
        else
            (void)cudaGetLastError();
    }
    // ... and the second one only inside the box its samples can fall into (geo2_window); the rest stays reachable through
    // the mapped pointer (option sparse_g2 = 0, or a pageable buffer: plain copy).
    const float* g2_mapped = nullptr;
    {
        cudaPointerAttributes attr;
        if (b2p_options().sparse_g2 != 0 && cudaPointerGetAttributes(&attr, geofea2_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            g2_mapped = reinterpret_cast<const float*>(attr.devicePointer);
        else
            (void)cudaGetLastError();
    }
    std::vector<int> win_host(g2_mapped ? (size_t)B * 4 : 0);
    const int g2_margin = b2p_options().g2_margin < 0 ? 0 : b2p_options().g2_margin;
    HostScratch hs;
    host_scratch_layout(B, C_geo, H, W, gather ? 2 : (ctx_mapped == nullptr ? 1 : 0), device_scratch, device_scratch_bytes, &hs);
    const int h = H / 8, w = W / 8, bs = hs.bs;
    const int nsub = (B + bs - 1) / bs;
    const size_t f = sizeof(float);
    // gathered planes per object: all 256, or (option host_gather_planes, needs the mapped pointer for the rest) the first c_split
    int c_split = 0;
    if (gather) {
        c_split = b2p_options().host_gather_planes;
        c_split = ctx_mapped ? (c_split < 32 ? 32 : (c_split > 256 ? 256 : (c_split / 32) * 32)) : 256;
    }

    TexelGather tg;
    if (gather) tg.start(context_host, reinterpret_cast<float*>(host_staging), B, H, W, bs, nsub, n_host_threads, c_split);

    cudaStream_t cs = nullptr;                     // internal copy stream + events, created and destroyed per call
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_start = nullptr;
    int rc = 0;
#define B2P_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = (int)e__; goto cleanup; } } while (0)
    B2P_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    B2P_TRY(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        B2P_TRY(cudaEventCreateWithFlags(&ev_copy[k], cudaEventDisableTiming));
        B2P_TRY(cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming));
    }
    B2P_TRY(cudaEventRecord(ev_start, s));          // the copy stream starts after the caller's prior work
    B2P_TRY(cudaStreamWaitEvent(cs, ev_start, 0));
    for (int k = 0; k < nsub; ++k) {
        const int b0 = k * bs, nb = (B - b0 < bs) ? (B - b0) : bs;
        const HostStage& st = hs.st[k & 1];
        if (k >= 2) B2P_TRY(cudaStreamWaitEvent(cs, ev_done[k & 1], 0));        // staging buffer free again
        B2P_TRY(cudaMemcpyAsync(st.fmap1, fmap1_host + (size_t)b0 * 256 * h * w, (size_t)nb * 256 * h * w * f, cudaMemcpyHostToDevice, cs));
        B2P_TRY(cudaMemcpyAsync(st.fmap2, fmap2_host + (size_t)b0 * 256 * h * w, (size_t)nb * 256 * h * w * f, cudaMemcpyHostToDevice, cs));
        if (!ctx_mapped && !gather)
            B2P_TRY(cudaMemcpyAsync(st.context, context_host + (size_t)b0 * 256 * H * W, (size_t)nb * 256 * H * W * f, cudaMemcpyHostToDevice, cs));
        B2P_TRY(cudaMemcpyAsync(st.depth, depth_host + (size_t)b0 * H * W, (size_t)nb * H * W * f, cudaMemcpyHostToDevice, cs));
        if (g1_mapped) {
            rc = b2p_gather_fg_planes(st.depth, g1_mapped + (size_t)b0 * C_geo * H * W, st.geo1, nb, C_geo, H, W, cs);
            if (rc) goto cleanup;
        } else {
            B2P_TRY(cudaMemcpyAsync(st.geo1, geofea1_host + (size_t)b0 * C_geo * H * W, (size_t)nb * C_geo * H * W * f, cudaMemcpyHostToDevice, cs));
        }
        if (g2_mapped) {
            for (int i = 0; i < nb; ++i) {
                int* wn = &win_host[(size_t)(b0 + i) * 4];
                geo2_window(depth_host + (size_t)(b0 + i) * H * W, H, W, g2_margin, wn);
                if (wn[1] <= wn[0] || wn[3] <= wn[2]) continue;
                cudaMemcpy3DParms cp;
                memset(&cp, 0, sizeof(cp));
                cp.srcPtr = make_cudaPitchedPtr(const_cast<float*>(geofea2_host + (size_t)(b0 + i) * C_geo * H * W), (size_t)W * f, W, H);
                cp.dstPtr = make_cudaPitchedPtr(st.geo2 + (size_t)i * C_geo * H * W, (size_t)W * f, W, H);
                cp.srcPos = cp.dstPos = make_cudaPos((size_t)wn[2] * f, wn[0], 0);
                cp.extent = make_cudaExtent((size_t)(wn[3] - wn[2]) * f, wn[1] - wn[0], C_geo);
                cp.kind = cudaMemcpyHostToDevice;
                B2P_TRY(cudaMemcpy3DAsync(&cp, cs));
            }
            B2P_TRY(cudaMemcpyAsync(st.win, &win_host[(size_t)b0 * 4], (size_t)nb * 4 * sizeof(int), cudaMemcpyHostToDevice, cs));
        } else {
            B2P_TRY(cudaMemcpyAsync(st.geo2, geofea2_host + (size_t)b0 * C_geo * H * W, (size_t)nb * C_geo * H * W * f, cudaMemcpyHostToDevice, cs));
        }
        B2P_TRY(cudaMemcpyAsync(st.K, K_host + (size_t)b0 * 9, (size_t)nb * 9 * f, cudaMemcpyHostToDevice, cs));
        B2P_TRY(cudaMemcpyAsync(st.G, G_host + (size_t)b0 * 16, (size_t)nb * 16 * f, cudaMemcpyHostToDevice, cs));
        if (gather) {                               // the workers ran ahead while the copies above were queued / in flight
            tg.wait(k);
            const size_t per = (size_t)c_split * h * w * 4;
            B2P_TRY(cudaMemcpyAsync(st.context, reinterpret_cast<const float*>(host_staging) + (size_t)b0 * per, (size_t)nb * per * f, cudaMemcpyHostToDevice, cs));
        }
        B2P_TRY(cudaEventRecord(ev_copy[k & 1], cs));
        B2P_TRY(cudaStreamWaitEvent(s, ev_copy[k & 1], 0));
        const float* ctx = ctx_mapped ? ctx_mapped + (size_t)b0 * 256 * H * W : st.context;      // (unused when all 256 planes are gathered)
        rc = refine_iters_impl(packed_weights, st.fmap1, st.fmap2, ctx, st.geo1, st.geo2, st.depth, st.K, st.G, sigma, nb,
                               C_geo, H, W, n_iters, n_lm, ep_lmbda, lm_lmbda, flags, nullptr, nullptr, nullptr, hs.ws,
                               hs.ws_bytes, stream, g2_mapped ? g2_mapped + (size_t)b0 * C_geo * H * W : nullptr,
                               g2_mapped ? st.win : nullptr, gather ? st.context : nullptr, c_split);
        if (rc) goto cleanup;
        B2P_TRY(cudaMemcpyAsync(G_host + (size_t)b0 * 16, st.G, (size_t)nb * 16 * f, cudaMemcpyDeviceToHost, s));
        B2P_TRY(cudaEventRecord(ev_done[k & 1], s));
    }
    B2P_TRY(cudaStreamSynchronize(s));
cleanup:
#undef B2P_TRY
    tg.join();
    if (rc) { if (cs) cudaStreamSynchronize(cs); cudaStreamSynchronize(s); }
    for (int k = 0; k < 2; ++k) { if (ev_copy[k]) cudaEventDestroy(ev_copy[k]); if (ev_done[k]) cudaEventDestroy(ev_done[k]); }
    if (ev_start) cudaEventDestroy(ev_start);
    if (cs) cudaStreamDestroy(cs);
    return rc;
}

}  // extern "C"
