// Per-object pose metrics on device: ADD, ADD-S (brute-force nearest neighbour fused with the mean), 2-D projection
// error, the two rotation angles and the translation error the reference's evaluator computes, and its pass/fail flags.
// SURVEY.md section 8(f)-3.  Replaces, per object:
//   utils/eval_metric.py:161-179 add_metric (and :120-158 add2 / add5: same distances, 0.02 d / 0.05 d thresholds);
//     syn=True: idxs = find_nearest_point_idx(model_pred, model_targets), i.e. for every GROUND-TRUTH point the nearest
//     PREDICTED point (thirdparty/nn/nn_utils.py:6-22: ref = model_pred, que = model_targets;
//     thirdparty/nn/src/nearest_neighborhood.cu:48-80: one thread per query, linear scan, first minimum), then
//     mean || model_pred[idxs] - model_targets || over the ground-truth points -- the distance to the nearest point does
//     not depend on which of several equidistant points the scan picks, so the mean of sqrt(min squared distance) is the
//     same number;
//   utils/eval_metric.py:102-110 projection_2d with :23-35 project (mean pixel distance of the projected model, < 5 px);
//   utils/eval_metric.py:181-192 cm_degree_5_metric (angle from the trace, translation in cm);
//   utils/geometric.py:36-40 rotation_angle (chordal form, the "ang_err" of evaluate_rnnpose, eval_metric.py:326-327).
// Grid (query chunks, B): each thread owns one model point transformed by the GROUND-TRUTH pose and scans all points
// transformed by the PREDICTED pose, staged through shared memory in tiles; per-block partial sums go to the workspace
// and a finalise kernel reduces them in a fixed order (deterministic, no float atomics).
#include "common.cuh"

namespace {

constexpr int PM_THREADS = 256;
constexpr int PM_PART = 3;       // partial sums per block: ADD, ADD-S, projected distance

__device__ __forceinline__ float3 xform(const float* T, float3 p) {
    return make_float3(T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
                       T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]);
}

__device__ __forceinline__ float2 project_px(const float* K, float3 p) {      // eval_metric.py:23-35
    const float x = K[0] * p.x + K[1] * p.y + K[2] * p.z;
    const float y = K[3] * p.x + K[4] * p.y + K[5] * p.z;
    const float z = K[6] * p.x + K[7] * p.y + K[8] * p.z;
    return make_float2(x / z, y / z);
}

__global__ void __launch_bounds__(PM_THREADS) pose_metric_partial_kernel(const float* __restrict__ T_pred,
                                                                        const float* __restrict__ T_gt,
                                                                        const float* __restrict__ pts,
                                                                        const float* __restrict__ K, int n,
                                                                        float* __restrict__ partial /*[B][chunks][3]*/) {
    __shared__ float3 tile[PM_THREADS];
    __shared__ float red[PM_PART][PM_THREADS / 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * PM_THREADS + tid;
    const float* P = pts + (size_t)b * n * 3;
    float Tp[12], Tg[12], Kb[9];
#pragma unroll
    for (int k = 0; k < 12; ++k) { Tp[k] = T_pred[b * 16 + k]; Tg[k] = T_gt[b * 16 + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) Kb[k] = K[b * 9 + k];
    const bool live = i < n;
    float3 qg = make_float3(0.f, 0.f, 0.f), qp = qg;     // this thread's model point under the ground-truth / predicted pose
    if (live) {
        const float3 m = make_float3(P[i * 3], P[i * 3 + 1], P[i * 3 + 2]);
        qg = xform(Tg, m); qp = xform(Tp, m);
    }
    float best = INFINITY;                                // query = ground-truth point, searched set = predicted points
    for (int j0 = 0; j0 < n; j0 += PM_THREADS) {
        const int j = j0 + tid;
        if (j < n) tile[tid] = xform(Tp, make_float3(P[j * 3], P[j * 3 + 1], P[j * 3 + 2]));
        __syncthreads();
        const int cnt = min(PM_THREADS, n - j0);
        for (int k = 0; k < cnt; ++k) {
            const float dx = tile[k].x - qg.x, dy = tile[k].y - qg.y, dz = tile[k].z - qg.z;
            best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
        __syncthreads();
    }
    float v[PM_PART] = {0.f, 0.f, 0.f};
    if (live) {
        const float dx = qp.x - qg.x, dy = qp.y - qg.y, dz = qp.z - qg.z;
        v[0] = sqrtf(dx * dx + dy * dy + dz * dz);
        v[1] = sqrtf(best);
        const float2 a = project_px(Kb, qp), g = project_px(Kb, qg);
        v[2] = sqrtf((a.x - g.x) * (a.x - g.x) + (a.y - g.y) * (a.y - g.y));
    }
#pragma unroll
    for (int c = 0; c < PM_PART; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_down_sync(0xffffffffu, v[c], o);
        if ((tid & 31) == 0) red[c][tid >> 5] = v[c];
    }
    __syncthreads();
    if (tid < PM_PART) {
        float s = 0.f;
        for (int wv = 0; wv < PM_THREADS / 32; ++wv) s += red[tid][wv];
        partial[((size_t)b * gridDim.x + blockIdx.x) * PM_PART + tid] = s;
    }
}

// out[b][16], see include/b200pose.h
__global__ void pose_metric_finalize_kernel(const float* __restrict__ T_pred, const float* __restrict__ T_gt,
                                            const float* __restrict__ diameter, const float* __restrict__ partial, int chunks,
                                            int n, int B, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float s[PM_PART] = {0.f, 0.f, 0.f};
    for (int c = 0; c < chunks; ++c)
#pragma unroll
        for (int k = 0; k < PM_PART; ++k) s[k] += partial[((size_t)b * chunks + c) * PM_PART + k];
    const float add = s[0] / (float)n, adds = s[1] / (float)n, proj = s[2] / (float)n;
    const float* Tp = T_pred + b * 16; const float* Tg = T_gt + b * 16;
    float fro = 0.f, trace = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = Tg[r * 4 + c] - Tp[r * 4 + c];
            fro += d * d;
            trace += Tp[r * 4 + c] * Tg[r * 4 + c];         // trace(R_pred R_gt^T), eval_metric.py:184-185
        }
    const float rad2deg = 180.f / 3.14159265358979323846f;
    const float ang = 2.f * asinf(fminf(sqrtf(fro) / sqrtf(8.f), 1.f)) * rad2deg;          // geometric.py:36-40 (in degrees)
    trace = trace <= 3.f ? trace : 3.f;                                                      // eval_metric.py:186
    const float ang_tr = acosf((trace - 1.f) / 2.f) * rad2deg;                               // :187 (NaN below -1, as numpy)
    const float tx = Tp[3] - Tg[3], ty = Tp[7] - Tg[7], tz = Tp[11] - Tg[11];
    const float trans = sqrtf(tx * tx + ty * ty + tz * tz);
    const float d = diameter[b];
    float* o = out + (size_t)b * B200POSE_METRIC_COLS;
    o[0] = add; o[1] = adds; o[2] = ang; o[3] = trans; o[4] = proj; o[5] = ang_tr;
    o[6] = add < d * 0.1f ? 1.f : 0.f;   o[7] = adds < d * 0.1f ? 1.f : 0.f;                 // eval_metric.py:161-179
    o[8] = add < d * 0.02f ? 1.f : 0.f;  o[9] = adds < d * 0.02f ? 1.f : 0.f;                // :120-138
    o[10] = add < d * 0.05f ? 1.f : 0.f; o[11] = adds < d * 0.05f ? 1.f : 0.f;               // :140-158
    o[12] = proj < 5.f ? 1.f : 0.f;                                                          // :102-110
    o[13] = (trans * 100.f < 5.f && ang_tr < 5.f) ? 1.f : 0.f;                               // :181-192
    o[14] = 0.f;                                                                             // object index (filled by the caller)
    o[15] = d;
}

}  // namespace

size_t b2p_pose_metrics_ws_bytes(int B, int n) { return align_up((size_t)B * ceil_div(n, PM_THREADS) * PM_PART * sizeof(float), 256); }

int b2p_pose_metrics(const float* T_pred, const float* T_gt, const float* pts, const float* diameter, const float* K, int B,
                     int n, float* out, void* ws, cudaStream_t s) {
    const int chunks = ceil_div(n, PM_THREADS);
    float* partial = reinterpret_cast<float*>(ws);
    dim3 grid(chunks, B);
    pose_metric_partial_kernel<<<grid, PM_THREADS, 0, s>>>(T_pred, T_gt, pts, K, n, partial);
    pose_metric_finalize_kernel<<<ceil_div(B, 128), 128, 0, s>>>(T_pred, T_gt, diameter, partial, chunks, n, B, out);
    B2P_LAUNCH_CHECK();
    return 0;
}
