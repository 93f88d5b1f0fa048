// Process-wide kernel-selection options of libb200pose.so.
// They used to be getenv() calls on the launch path (several per launch; an inherited variable could silently change which
// kernels -- and therefore which rounding -- a run used).  Now: read ONCE, at the first call into the library, from the
// B200POSE_* environment variables below; afterwards only b200pose_set_option changes them.  Setting an option while
// another thread is inside a launch is not synchronised (the caller's problem, like changing a stream's priority).
#include "common.cuh"

#include <mutex>
#include <stdlib.h>

namespace {

struct OptDesc { const char* name; const char* env; int B2POptions::*field; int dflt; };

const OptDesc kOpts[] = {
    // tensor-core convolution kernel: bit 0 CTA pairs (cta_group::2), bit 1 vertical-tap reuse, bit 2 force the
    // second-generation kernel, bit 4 the chained single-launch update block; 0 = first-generation kernel
    {"conv_mode", "B200POSE_CONV_MODE", &B2POptions::conv_mode, 19},
    {"fg_list", "B200POSE_FG_LIST", &B2POptions::fg_list, 1},               // LM over the per-call foreground list
    {"fg_pipeline", "B200POSE_FG_PIPELINE", &B2POptions::fg_pipeline, 2},   // channels-last target/weight kernels + cluster LM: 0 off, 1 on, 2 when geofea2 is channels-last
    {"fg_upsample", "B200POSE_FG_UPSAMPLE", &B2POptions::fg_upsample, 0},   // round-1 list-driven upsample kernel (NCHW planes)
    {"sparse_g1", "B200POSE_SPARSE_G1", &B2POptions::sparse_g1, 1},         // host entry: fetch geofea1 only where depth > 0
    {"fg_blocks", "B200POSE_FG_BLOCKS", &B2POptions::fg_blocks, 8},
    {"tail_min_n", "B200POSE_TAIL_MIN_N", &B2POptions::tail_min_n, 32},     // smallest channel count of a split tail unit
    {"conv_debug", "B200POSE_V2_DEBUG", &B2POptions::conv_debug, 0},        // timing experiments (results garbage unless 0 / 16)
    {"lookup_mode", "B200POSE_LOOKUP_MODE", &B2POptions::lookup_mode, 2},   // 2 = one thread per window row (3: the same with PDL), 1 = shared-memory window lookup, 0 = round-1 kernel
    {"pool_mode", "B200POSE_POOL_MODE", &B2POptions::pool_mode, 1},         // 1 = three pyramid levels in one pass
    {"lm_cluster", "B200POSE_LM_CLUSTER", &B2POptions::lm_cluster, 1},      // LM over the list: cluster kernel (1) or spin-barrier kernel (0)
    {"lm_debug", "B200POSE_LM_DEBUG", &B2POptions::lm_debug, 0},            // 1 = drop the fp64 contraction (timing A/B only)
    {"chain_rings", "B200POSE_CHAIN_RINGS", &B2POptions::chain_rings, 24},  // chained launch: 10 * activation slots + weight slots
    {"chain_xmajor", "B200POSE_CHAIN_XMAJOR", &B2POptions::chain_xmajor, 1},   // chained launch: horizontal-tap reuse for the 1x5 layers
    {"chain_merge", "B200POSE_CHAIN_MERGE", &B2POptions::chain_merge, 0},     // chained launch: interleave C1|F1 and MASK2|flow head units (measured: no gain)
    {"upsample_variant", "B200POSE_UPSAMPLE_VARIANT", &B2POptions::upsample_variant, 3},   // dense upsample+weight kernel build, see b2p_upsample_weight
    {"host_gather", "B200POSE_HOST_GATHER", &B2POptions::host_gather, 1},     // host texel gather: bit 0 software prefetch of the next rows, bit 1 plain (not streaming) stores
    // host entry: copy geofea2 only inside the foreground box + g2_margin, the rest is read from mapped host memory on demand.
    // Measured (profiles/r2r): +12 % end to end when the flows stay inside the margin, neutral on the bench scenes (their flows
    // do not); off by default because a diverged refinement would pull more over PCIe sector by sector than the copy it saves.
    {"sparse_g2", "B200POSE_SPARSE_G2", &B2POptions::sparse_g2, 0},
    {"g2_margin", "B200POSE_G2_MARGIN", &B2POptions::g2_margin, 24},
    {"enc_chunk", "B200POSE_ENC_CHUNK", &B2POptions::enc_chunk, 0},           // image encoder: crop pairs per pass (0 = the whole batch at once)
    {"enc_stem", "B200POSE_ENC_STEM", &B2POptions::enc_stem, 1},              // image encoder stem: 1 = tensor cores (gathered 4x1 form), 0 = fp32 FFMA kernel
    // host entry with a staging buffer: how many of the 256 context planes per object the host threads gather (a multiple of 32);
    // the kernel reads the rest in place from the mapped buffer.  256 when one GPU has the host to itself; fewer when several
    // ranks share the host's cores (bench.py tries both)
    {"host_gather_planes", "B200POSE_HOST_GATHER_PLANES", &B2POptions::host_gather_planes, 256},
    // bit mask of loop kernels launched WITHOUT the PDL attribute (tags in common.cuh).  Default 16 = the dense upsample + weight
    // kernel: launched early its 38 400 blocks sit on the SMs during the chained convolution launch (3.387 -> 3.307 ms per batch
    // without; LM, flow_init, im2col and the chained launch itself keep it: no gain or a loss without, profiles/r3f)
    {"pdl_off", "B200POSE_PDL_OFF", &B2POptions::pdl_off, 16},
    {"chain_dynamic", "B200POSE_CHAIN_DYNAMIC", &B2POptions::chain_dynamic, 0},   // chained launch: units from a global queue (1) or static round robin (0)
};
constexpr int kNumOpts = (int)(sizeof(kOpts) / sizeof(kOpts[0]));

B2POptions g_opts;
std::once_flag g_once;

void init_opts() {
    for (int i = 0; i < kNumOpts; ++i) {
        const char* e = getenv(kOpts[i].env);
        g_opts.*(kOpts[i].field) = (e && *e) ? atoi(e) : kOpts[i].dflt;
    }
}

}  // namespace

B2POptions& b2p_options() {
    std::call_once(g_once, init_opts);
    return g_opts;
}

extern "C" {

int b200pose_set_option(const char* name, int value) {
    if (!name) return B200POSE_E_NULL;
    B2POptions& o = b2p_options();
    for (int i = 0; i < kNumOpts; ++i)
        if (!strcmp(name, kOpts[i].name)) { o.*(kOpts[i].field) = value; return 0; }
    return B200POSE_E_ARG;
}

int b200pose_get_option(const char* name, int* value) {
    if (!name || !value) return B200POSE_E_NULL;
    B2POptions& o = b2p_options();
    for (int i = 0; i < kNumOpts; ++i)
        if (!strcmp(name, kOpts[i].name)) { *value = o.*(kOpts[i].field); return 0; }
    return B200POSE_E_ARG;
}

int b200pose_option_count(void) { return kNumOpts; }

const char* b200pose_option_name(int index) { return index >= 0 && index < kNumOpts ? kOpts[index].name : nullptr; }

}  // extern "C"
