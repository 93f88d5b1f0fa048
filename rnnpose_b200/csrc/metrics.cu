// Per-object pose metrics on device: ADD, ADD-S (brute-force nearest neighbour fused with the mean), rotation
// angle, translation error and the pass/fail flags the reference's evaluator accumulates.  SURVEY.md section 8(f)-3.
// Replaces: reference thirdparty/nn/src/nearest_neighborhood.cu:48-117 (sm_52 NN-index kernel + numpy mean),
//   utils/eval_metric.py:161-192 (add_metric, cm_degree_5_metric) and utils/geometric.py:36-40 (rotation_angle).
// Grid (query chunks, B): each thread owns one model point transformed by the predicted pose and scans all points
// transformed by the ground-truth pose, staged through shared memory in tiles; per-block partial sums are written to
// the workspace and reduced in a fixed order by a one-warp finalise kernel (deterministic, no float atomics).
#include "common.cuh"

namespace {

constexpr int PM_THREADS = 256;

__device__ __forceinline__ float3 xform(const float* T, float3 p) {
    return make_float3(T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3], T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7],
                       T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11]);
}

__global__ void __launch_bounds__(PM_THREADS) pose_metric_partial_kernel(const float* __restrict__ T_pred,
                                                                        const float* __restrict__ T_gt,
                                                                        const float* __restrict__ pts, int n,
                                                                        float* __restrict__ partial /*[B][chunks][2]*/) {
    __shared__ float3 tile[PM_THREADS];
    __shared__ float red[2][PM_THREADS / 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int i = blockIdx.x * PM_THREADS + tid;
    const float* P = pts + (size_t)b * n * 3;
    float Tp[12], Tg[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { Tp[k] = T_pred[b * 16 + k]; Tg[k] = T_gt[b * 16 + k]; }
    const bool live = i < n;
    float3 q = make_float3(0.f, 0.f, 0.f), qg = q;
    if (live) {
        const float3 m = make_float3(P[i * 3], P[i * 3 + 1], P[i * 3 + 2]);
        q = xform(Tp, m); qg = xform(Tg, m);
    }
    float best = INFINITY;
    for (int j0 = 0; j0 < n; j0 += PM_THREADS) {
        const int j = j0 + tid;
        if (j < n) tile[tid] = xform(Tg, make_float3(P[j * 3], P[j * 3 + 1], P[j * 3 + 2]));
        __syncthreads();
        const int cnt = min(PM_THREADS, n - j0);
        for (int k = 0; k < cnt; ++k) {
            const float dx = q.x - tile[k].x, dy = q.y - tile[k].y, dz = q.z - tile[k].z;
            best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
        __syncthreads();
    }
    float add = 0.f, adds = 0.f;
    if (live) {
        const float dx = q.x - qg.x, dy = q.y - qg.y, dz = q.z - qg.z;
        add = sqrtf(dx * dx + dy * dy + dz * dz);
        adds = sqrtf(best);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        add += __shfl_down_sync(0xffffffffu, add, o);
        adds += __shfl_down_sync(0xffffffffu, adds, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = add; red[1][tid >> 5] = adds; }
    __syncthreads();
    if (tid < 2) {
        float s = 0.f;
        for (int wv = 0; wv < PM_THREADS / 32; ++wv) s += red[tid][wv];
        partial[((size_t)b * gridDim.x + blockIdx.x) * 2 + tid] = s;
    }
}

// out[b] = {ADD, ADD-S, ang_err_deg, trans_err, ADD < 0.1 d, ADD-S < 0.1 d, 5cm & 5deg, 0}
__global__ void pose_metric_finalize_kernel(const float* __restrict__ T_pred, const float* __restrict__ T_gt,
                                            const float* __restrict__ diameter, const float* __restrict__ partial, int chunks,
                                            int n, int B, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float add = 0.f, adds = 0.f;
    for (int c = 0; c < chunks; ++c) { add += partial[((size_t)b * chunks + c) * 2]; adds += partial[((size_t)b * chunks + c) * 2 + 1]; }
    add /= (float)n; adds /= (float)n;
    const float* Tp = T_pred + b * 16; const float* Tg = T_gt + b * 16;
    float fro = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float d = Tg[r * 4 + c] - Tp[r * 4 + c]; fro += d * d; }
    const float ang = 2.f * asinf(fminf(sqrtf(fro) / sqrtf(8.f), 1.f)) * (180.f / 3.14159265358979323846f);   // geometric.py:36-40
    const float tx = Tp[3] - Tg[3], ty = Tp[7] - Tg[7], tz = Tp[11] - Tg[11];
    const float trans = sqrtf(tx * tx + ty * ty + tz * tz);
    const float thr = 0.1f * diameter[b];
    float* o = out + (size_t)b * 8;
    o[0] = add; o[1] = adds; o[2] = ang; o[3] = trans;
    o[4] = add < thr ? 1.f : 0.f; o[5] = adds < thr ? 1.f : 0.f;
    o[6] = (trans * 100.f < 5.f && ang < 5.f) ? 1.f : 0.f;                                    // eval_metric.py:181-192
    o[7] = 0.f;
}

}  // namespace

size_t b2p_pose_metrics_ws_bytes(int B, int n) { return align_up((size_t)B * ceil_div(n, PM_THREADS) * 2 * sizeof(float), 256); }

int b2p_pose_metrics(const float* T_pred, const float* T_gt, const float* pts, const float* diameter, int B, int n, float* out,
                     void* ws, cudaStream_t s) {
    const int chunks = ceil_div(n, PM_THREADS);
    float* partial = reinterpret_cast<float*>(ws);
    dim3 grid(chunks, B);
    pose_metric_partial_kernel<<<grid, PM_THREADS, 0, s>>>(T_pred, T_gt, pts, n, partial);
    pose_metric_finalize_kernel<<<ceil_div(B, 128), 128, 0, s>>>(T_pred, T_gt, diameter, partial, chunks, n, B, out);
    B2P_LAUNCH_CHECK();
    return 0;
}
