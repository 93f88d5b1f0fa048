// Levenberg-Marquardt steps of the reprojection objective, fully on device:
//   per-pixel residual + 2x6 SE(3) Jacobian (fp32 geometry, as the reference), fp64 accumulation of
//   H = sum v w J^T J and b = sum v w J^T r, damping, 6x6 Cholesky solve (NaN -> 0, clamp +-1),
//   se3 exponential and the left-multiplicative retraction G <- exp(delta) G.
// reference geometry/transformation.py:265-316 (reprojction_optim), :27-46 (jac_local_perturb),
//   geometry/projective_ops.py:68-131, geometry/cholesky.py:32-50, geometry/se3.py:228-306.
//
// Two kernels share the arithmetic:
//   lm_step_kernel   one step per launch; many small blocks, the last block of each sample (ticket counter) sums the
//                    per-block partials in a fixed order, solves and retracts.  Provides the H / b / delta taps.
//   lm_multi_kernel  all n steps in ONE launch: a few fat blocks per sample, all co-resident; after each step the
//                    blocks of a sample meet at a per-sample spin barrier, every block then reduces the same partials in
//                    the same order and solves redundantly (bit-identical), so no broadcast and no relaunch is needed.
// Both reductions are deterministic (no floating-point atomics).
#include "common.cuh"

namespace {

constexpr int LM_THREADS = 256;
constexpr int LM_PX_PER_THREAD = 4;
constexpr int LM_PX_PER_BLOCK = LM_THREADS * LM_PX_PER_THREAD;
constexpr int NACC = 27;   // 21 upper-triangular entries of H + 6 of b
constexpr int LM_MAX_BLK_PER_SAMPLE = 64;

__device__ __forceinline__ int tri_idx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

// fp32 se3 exponential, reference geometry/se3.py:228-281 (Taylor branch below MIN_THETA = 1e-4)
__device__ void se3_exp_f32(const float* xi, float* dG /*12: rows of [R|t]*/) {
    const float v0 = xi[0], v1 = xi[1], v2 = xi[2];
    const float w0 = xi[3], w1 = xi[4], w2 = xi[5];
    const float th2 = (w0 * w0 + w1 * w1) + w2 * w2;
    const float th = sqrtf(th2);
    const float th4 = th2 * th2;
    const float wx[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
    float wx2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) wx2[i * 3 + j] = wx[i * 3 + 0] * wx[0 * 3 + j] + wx[i * 3 + 1] * wx[1 * 3 + j] + wx[i * 3 + 2] * wx[2 * 3 + j];
    float ra, rb, va, vb;
    if (th < 1e-4f) {
        ra = 1.0f - (1.0f / 6.0f) * th2 + (1.0f / 120.0f) * th4;
        rb = 0.5f - (1.0f / 12.0f) * th2 + (1.0f / 720.0f) * th4;
        va = 0.5f - (1.0f / 24.0f) * th2 + (1.0f / 720.0f) * th4;
        vb = (1.0f / 6.0f) - (1.0f / 120.0f) * th2 + (1.0f / 5040.0f) * th4;
    } else {
        const float eps = 1e-12f;
        const float s = sinf(th), c = cosf(th);
        ra = s / (th + eps);
        rb = (1.f - c) / (th2 + eps);
        va = (1.f - c) / (th2 + eps);
        vb = (th - s) / (th2 * th + eps);
    }
    float R[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float I = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
        R[i] = I + ra * wx[i] + rb * wx2[i];
        V[i] = I + va * wx[i] + vb * wx2[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dG[i * 4 + 0] = R[i * 3 + 0]; dG[i * 4 + 1] = R[i * 3 + 1]; dG[i * 4 + 2] = R[i * 3 + 2];
        dG[i * 4 + 3] = V[i * 3 + 0] * v0 + V[i * 3 + 1] * v1 + V[i * 3 + 2] * v2;
    }
}

// Contribution of one pixel to the 27 accumulators.
__device__ __forceinline__ void lm_pixel(double (&acc)[NACC], int u, int v, float Zraw, float2 tg, float wv, float depth_add,
                                         float fx, float fy, float cx, float cy, const float (&Gm)[12]) {
    const float Z = Zraw + depth_add;
    // A pixel whose weight is exactly 0 adds exactly 0 to H and b (all terms finite): skip its fp64 work.
    // 65-80 % of a crop is background (weight = ... * (depth > 0)).  Non-finite inputs still take the full path so
    // that they poison the sums exactly like the reference's arithmetic (NaN -> zero update downstream).
    if (wv == 0.f && isfinite(Z) && isfinite(tg.x) && isfinite(tg.y)) return;
    const float X = Z * ((float)u - cx) / fx;
    const float Y = Z * ((float)v - cy) / fy;
    const float X1 = Gm[0] * X + Gm[1] * Y + Gm[2] * Z + Gm[3];
    const float Y1 = Gm[4] * X + Gm[5] * Y + Gm[6] * Z + Gm[7];
    const float Z1 = Gm[8] * X + Gm[9] * Y + Gm[10] * Z + Gm[11];
    const double valid = (Z > 0.1f && Z1 > 0.1f) ? 1.0 : 0.0;
    const float Zc = fmaxf(Z1, 0.01f);
    const float x1 = fx * (X1 / Zc) + cx;
    const float y1 = fy * (Y1 / Zc) + cy;
    const bool cut = Zc <= 0.02f;
    const float zi1 = cut ? 0.f : 1.0f / Zc;
    const float zi2 = cut ? 0.f : 1.0f / (Zc * Zc);
    const double A = (double)(fx * zi1), C = (double)((-fx * X1) * zi2);
    const double Bq = (double)(fy * zi1), D = (double)((-fy * Y1) * zi2);
    const double dX = (double)X1, dY = (double)Y1, dZ = (double)Z1;
    double J0[6], J1[6];
    J0[0] = A;   J0[1] = 0.0; J0[2] = C; J0[3] = C * dY;              J0[4] = A * dZ + C * (-dX); J0[5] = A * (-dY);
    J1[0] = 0.0; J1[1] = Bq;  J1[2] = D; J1[3] = Bq * (-dZ) + D * dY; J1[4] = D * (-dX);          J1[5] = Bq * dX;
    const double r0 = (double)tg.x - (double)x1;
    const double r1 = (double)tg.y - (double)y1;
    const double vw = valid * (double)wv;
    double wJ0[6], wJ1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { wJ0[i] = vw * J0[i]; wJ1[i] = vw * J1[i]; }
    // J0[1] and J1[0] are structural zeros (jproj rows (fx/Z,0,.) and (0,fy/Z,.)): their products are dropped at
    // compile time, 30 + 10 FMAs instead of 42 + 12
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) {
            const bool u0 = (i != 1) && (j != 1), u1 = (i != 0) && (j != 0);
            if (u0 && u1) acc[tri_idx(i, j)] += wJ0[i] * J0[j] + wJ1[i] * J1[j];
            else if (u0) acc[tri_idx(i, j)] += wJ0[i] * J0[j];
            else if (u1) acc[tri_idx(i, j)] += wJ1[i] * J1[j];
        }
        if (i == 0) acc[21 + i] += wJ0[i] * r0;
        else if (i == 1) acc[21 + i] += wJ1[i] * r1;
        else acc[21 + i] += wJ0[i] * r0 + wJ1[i] * r1;
    }
}

// Block-level fixed-order reduction of the 27 accumulators; result for value i in out[i] (valid for tid < NACC).
__device__ __forceinline__ void lm_block_reduce(double (&acc)[NACC], double (*red)[NACC], int tid, double& out) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[wid][i] = v;
    }
    __syncthreads();
    out = 0.0;
    if (tid < NACC) {
#pragma unroll
        for (int wv = 0; wv < LM_THREADS / 32; ++wv) out += red[wv][tid];
    }
}

// 6x6 Cholesky solve in fp64, then NaN -> 0 and clamp to +-1 as fp32: geometry/cholesky.py:11-16,32-50
// (torch.cholesky + cholesky_solve; max_update = 1.0).  Hm is symmetric positive definite (or poisoned by NaN).
__device__ void chol_solve6(const double (&Hm)[6][6], const double (&bv)[6], float (&xi)[6]) {
    // one reciprocal per pivot (6 fp64 divisions on the dependent chain instead of 27; a product with the reciprocal differs
    // from the quotient by at most an fp64 ulp, far below the fp32 result)
    double L[6][6], rinv[6];
    for (int j = 0; j < 6; ++j) {
        double s = Hm[j][j];
        for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
        const double d = sqrt(s);
        L[j][j] = d;
        rinv[j] = 1.0 / d;
        for (int i = j + 1; i < 6; ++i) {
            double t = Hm[i][j];
            for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
            L[i][j] = t * rinv[j];
        }
    }
    double yv[6], xv[6];
    for (int i = 0; i < 6; ++i) {
        double t = bv[i];
        for (int k = 0; k < i; ++k) t -= L[i][k] * yv[k];
        yv[i] = t * rinv[i];
    }
    for (int i = 5; i >= 0; --i) {
        double t = yv[i];
        for (int k = i + 1; k < 6; ++k) t -= L[k][i] * xv[k];
        xv[i] = t * rinv[i];
    }
    for (int i = 0; i < 6; ++i) {
        double x = xv[i];
        if (x != x) x = 0.0;                       // NaN -> 0 (cholesky.py:42-43)
        x = fmin(fmax(x, -1.0), 1.0);               // clamp to +-max_update (cholesky.py:45)
        xi[i] = (float)x;
    }
}

// Damping, Cholesky solve, NaN -> 0, clamp, exp, retraction.  tot = 21 H entries + 6 b entries (un-damped).
__device__ void lm_solve_retract(const double* tot, const float (&Gm)[12], float (&Gn)[12], double ep, double lm,
                                 double* H_out, double* b_out, float* delta_out) {
    double Hm[6][6], bv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        bv[i] = tot[21 + i];
#pragma unroll
        for (int j = i; j < 6; ++j) { Hm[i][j] = tot[tri_idx(i, j)]; Hm[j][i] = Hm[i][j]; }
    }
    if (H_out)
        for (int i = 0; i < 36; ++i) H_out[i] = Hm[i / 6][i % 6];
    if (b_out)
        for (int i = 0; i < 6; ++i) b_out[i] = bv[i];
    // damping: H += ep*I + lm*H*I   (transformation.py:300)
#pragma unroll
    for (int i = 0; i < 6; ++i) Hm[i][i] = Hm[i][i] + (ep + lm * Hm[i][i]);
    float xi[6];
    chol_solve6(Hm, bv, xi);
    if (delta_out)
        for (int i = 0; i < 6; ++i) delta_out[i] = xi[i];
    float dG[12];
    se3_exp_f32(xi, dG);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = dG[i * 4 + 0] * Gm[0 * 4 + j] + dG[i * 4 + 1] * Gm[1 * 4 + j] + dG[i * 4 + 2] * Gm[2 * 4 + j];
            if (j == 3) s += dG[i * 4 + 3];       // last row of G is (0,0,0,1)
            Gn[i * 4 + j] = s;
        }
}

__device__ __forceinline__ void store_G(float* G, const float (&Gn)[12]) {
    for (int i = 0; i < 12; ++i) G[i] = Gn[i];
    G[12] = 0.f; G[13] = 0.f; G[14] = 0.f; G[15] = 1.f;
}

// ------------------------------------------------------------------------------------------------ one step per launch
__global__ void __launch_bounds__(LM_THREADS) lm_step_kernel(
    const float* __restrict__ depth, const float* __restrict__ target, const float* __restrict__ weight,
    const float* __restrict__ K, float* __restrict__ G, int B, int H, int W, float depth_add, double ep, double lm,
    double* __restrict__ partials, unsigned* __restrict__ counters, int nblk,
    double* __restrict__ H_out, double* __restrict__ b_out, float* __restrict__ delta_out) {
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int N = H * W;
    const float* Kb = K + b * 9;
    const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
    float Gm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Gm[i] = G[b * 16 + i];

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
    const float* dptr = depth + (size_t)b * N;
    const float2* tptr = reinterpret_cast<const float2*>(target) + (size_t)b * N;
    const float* wptr = weight + (size_t)b * N;

    // all loads of the thread's pixels are issued before any arithmetic (12 independent requests in flight)
    float Zs[LM_PX_PER_THREAD], ws_[LM_PX_PER_THREAD];
    float2 tgs[LM_PX_PER_THREAD];
#pragma unroll
    for (int it = 0; it < LM_PX_PER_THREAD; ++it) {
        const int px = blockIdx.x * LM_PX_PER_BLOCK + it * LM_THREADS + tid;
        const bool in = px < N;
        Zs[it] = in ? __ldg(dptr + px) : 0.f;
        tgs[it] = in ? __ldg(tptr + px) : make_float2(0.f, 0.f);
        ws_[it] = in ? __ldg(wptr + px) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < LM_PX_PER_THREAD; ++it) {
        const int px = blockIdx.x * LM_PX_PER_BLOCK + it * LM_THREADS + tid;
        if (px < N) lm_pixel(acc, px % W, px / W, Zs[it], tgs[it], ws_[it], depth_add, fx, fy, cx, cy, Gm);
    }

    __shared__ double red[LM_THREADS / 32][NACC];
    __shared__ double tot[NACC];
    __shared__ bool is_last;
    double mine;
    lm_block_reduce(acc, red, tid, mine);
    if (tid < NACC) partials[((size_t)b * nblk + blockIdx.x) * NACC + tid] = mine;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned t = atomicAdd(&counters[b], 1u);
        is_last = (t == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid < NACC) {
        double s = 0.0;
        const double* pp = partials + (size_t)b * nblk * NACC + tid;
        for (int k = 0; k < nblk; ++k) s += __ldcg(pp + (size_t)k * NACC);
        tot[tid] = s;
    }
    __syncthreads();
    if (tid != 0) return;
    counters[b] = 0;   // self-cleaning for the next step
    float Gn[12];
    lm_solve_retract(tot, Gm, Gn, ep, lm, H_out ? H_out + (size_t)b * 36 : nullptr, b_out ? b_out + (size_t)b * 6 : nullptr,
                     delta_out ? delta_out + (size_t)b * 6 : nullptr);
    store_G(G + b * 16, Gn);
}

// ------------------------------------------------------------------------------------------------ n steps per launch
// grid (nb, B) with nb * B blocks ALL co-resident (checked on the host).  counters: [B][2] = {arrivals, finished}.
__global__ void __launch_bounds__(LM_THREADS) lm_multi_kernel(
    const float* __restrict__ depth, const float* __restrict__ target, const float* __restrict__ weight,
    const float* __restrict__ K, float* __restrict__ G, int B, int H, int W, float depth_add, double ep, double lm,
    int n_steps, double* __restrict__ partials /*[2][B][nb][27]*/, unsigned* __restrict__ counters,
    const int* __restrict__ fg_idx, const int* __restrict__ fg_count) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y, nb = gridDim.x, tid = threadIdx.x;
    const int N = H * W;
    const float* Kb = K + b * 9;
    const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
    __shared__ double red[LM_THREADS / 32][NACC];
    __shared__ double tot[NACC];
    __shared__ float Gs[12];
    if (tid < 12) Gs[tid] = G[b * 16 + tid];
    __syncthreads();
    const float* dptr = depth + (size_t)b * N;
    const float2* tptr = reinterpret_cast<const float2*>(target) + (size_t)b * N;
    const float* wptr = weight + (size_t)b * N;
    // contiguous range of this block: of the pixels, or of the sample's foreground list (every unlisted pixel has weight
    // exactly 0 and finite inputs, i.e. contributes exactly 0)
    const int* fidx = fg_idx ? fg_idx + (size_t)b * N : nullptr;
    const int count = fg_idx ? fg_count[b] : N;
    const int per = (count + nb - 1) / nb;
    const int p_begin = blockIdx.x * per, p_end = min(count, p_begin + per);
    unsigned* arrive = counters + 2 * b;
    unsigned* finished = counters + 2 * b + 1;

    for (int step = 0; step < n_steps; ++step) {
        float Gm[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) Gm[i] = Gs[i];
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
        for (int base = p_begin; base < p_end; base += LM_PX_PER_BLOCK) {
            float Zs[LM_PX_PER_THREAD], ws_[LM_PX_PER_THREAD];
            float2 tgs[LM_PX_PER_THREAD];
            int rs[LM_PX_PER_THREAD];
#pragma unroll
            for (int it = 0; it < LM_PX_PER_THREAD; ++it) {
                const int px = base + it * LM_THREADS + tid;
                rs[it] = px < p_end ? (fidx ? __ldg(fidx + px) : px) : -1;
            }
#pragma unroll
            for (int it = 0; it < LM_PX_PER_THREAD; ++it) {
                const bool in = rs[it] >= 0;
                Zs[it] = in ? __ldg(dptr + rs[it]) : 0.f;
                tgs[it] = in ? tptr[rs[it]] : make_float2(0.f, 0.f);
                ws_[it] = in ? wptr[rs[it]] : 0.f;
            }
#pragma unroll
            for (int it = 0; it < LM_PX_PER_THREAD; ++it) {
                if (rs[it] >= 0) lm_pixel(acc, rs[it] % W, rs[it] / W, Zs[it], tgs[it], ws_[it], depth_add, fx, fy, cx, cy, Gm);
            }
        }
        double mine;
        lm_block_reduce(acc, red, tid, mine);
        double* mypart = partials + (((size_t)(step & 1) * B + b) * nb + blockIdx.x) * NACC;
        if (tid < NACC) mypart[tid] = mine;
        __threadfence();
        __syncthreads();
        // per-sample barrier: arrivals are monotonic over the steps of this launch
        if (tid == 0) {
            atomicAdd(arrive, 1u);
            const unsigned want = (unsigned)(step + 1) * (unsigned)nb;
            unsigned spins = 0;
            while (*reinterpret_cast<volatile unsigned*>(arrive) < want) {
                __nanosleep(32);
                if (++spins > (1u << 24)) __trap();          // co-residency violated: fail loudly, never hang
            }
            __threadfence();
        }
        __syncthreads();
        if (tid < NACC) {
            // same left-to-right order in every block; the loads of eight partials are in flight together
            double s = 0.0;
            const double* pp = partials + ((size_t)(step & 1) * B + b) * nb * NACC + tid;
            int k = 0;
            for (; k + 8 <= nb; k += 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldcg(pp + (size_t)(k + j) * NACC);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[j];
            }
            for (; k < nb; ++k) s += __ldcg(pp + (size_t)k * NACC);
            tot[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            float Gn[12];
            lm_solve_retract(tot, Gm, Gn, ep, lm, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < 12; ++i) Gs[i] = Gn[i];
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (blockIdx.x == 0) {
            float Gn[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) Gn[i] = Gs[i];
            store_G(G + b * 16, Gn);
        }
        // the last block to finish resets the counters (everybody has passed the final barrier by then)
        const unsigned f = atomicAdd(finished, 1u);
        if (f == (unsigned)nb - 1u) { *arrive = 0u; *finished = 0u; __threadfence(); }
    }
}

// ------------------------------------------------------------------------------------------------ n steps, one CLUSTER per sample
// Third kernel, used by the foreground pipeline (fg_pipeline.cu): the pixels arrive as float4 records (target x, target y,
// weight, depth) in list order, so the loads are a stream.  One thread-block cluster of LMC_CTAS CTAs per sample replaces
// the global spin barrier of lm_multi_kernel:
//   * the hardware co-schedules the CTAs of a cluster (no residency estimate, nothing to trap on);
//   * after each step every CTA writes its 27 partial sums into CTA 0's shared memory (distributed shared memory), one
//     barrier.cluster, then every CTA sums the LMC_CTAS partials in rank order and solves redundantly (bit-identical), so
//     nothing is broadcast; the partial slots are double-buffered by step parity, which makes one barrier per step enough;
//   * the reduction structure depends only on the sample's own list (not on the batch size): results are independent of
//     the batch position and of B.
// NOACC (option lm_debug = 1) drops the fp64 contraction (H, b stay ~0): a timing experiment that shows what a
// tensor-core J^T W J could save at most (profiles/r2_summary.md); results are meaningless with it.
constexpr int LMC_CTAS = 8;

__device__ __forceinline__ uint32_t lm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lm_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void lm_st_cluster_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lm_ld_cluster_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void lm_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t lm_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// RECORDS: the pixels' (target, weight, depth) come as float4 records in list order (foreground pipeline); otherwise they are
// gathered from the dense depth / target / weight maps through the list (the default NCHW path).
template <bool NOACC, bool RECORDS>
__global__ void __launch_bounds__(LM_THREADS) lm_cluster_kernel(
    const float4* __restrict__ rec, const float* __restrict__ depth, const float* __restrict__ target, const float* __restrict__ weight,
    const int* __restrict__ fg_idx, const int* __restrict__ fg_count, const float* __restrict__ K,
    float* __restrict__ G, int N, int W, float depth_add, double ep, double lm, int n_steps) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x / LMC_CTAS, tid = threadIdx.x;
    const uint32_t rank = lm_cluster_rank();
    __shared__ double red[LM_THREADS / 32][NACC];
    __shared__ double slots[2][LMC_CTAS][NACC];            // CTA 0's copy collects the cluster's partials
    __shared__ double tot[NACC];
    __shared__ float Gs[12];
    const float* Kb = K + b * 9;
    const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
    if (tid < 12) Gs[tid] = G[b * 16 + tid];
    __syncthreads();
    const int count = fg_count[b];
    const float4* rb = RECORDS ? rec + (size_t)b * N : nullptr;
    const int* ib = fg_idx + (size_t)b * N;
    const int first = (int)rank * LM_THREADS + tid, stride = LMC_CTAS * LM_THREADS;
    const uint32_t slot0 = lm_mapa(lm_smem_u32(&slots[0][0][0]), 0);      // slots[][][] of CTA 0, cluster address space

    for (int step = 0; step < n_steps; ++step) {
        float Gm[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) Gm[i] = Gs[i];
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
        for (int k = first; k < count; k += 4 * stride) {
            int rs[4]; float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * stride;
                rs[u] = kk < count ? __ldg(ib + kk) : -1;
                if (RECORDS) v[u] = kk < count ? rb[kk] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!RECORDS) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool in = rs[u] >= 0;
                    const size_t o = (size_t)b * N + (in ? rs[u] : 0);
                    const float2 tg = in ? reinterpret_cast<const float2*>(target)[o] : make_float2(0.f, 0.f);
                    v[u] = make_float4(tg.x, tg.y, in ? weight[o] : 0.f, in ? __ldg(depth + o) : 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (rs[u] < 0) continue;
                if (NOACC) {
                    double dummy[NACC];
#pragma unroll
                    for (int i = 0; i < NACC; ++i) dummy[i] = 0.0;
                    lm_pixel(dummy, rs[u] % W, rs[u] / W, v[u].w, make_float2(v[u].x, v[u].y), v[u].z, depth_add, fx, fy, cx, cy, Gm);
                    acc[0] += dummy[0];                     // keeps the loads and the Jacobian, drops 26 of the 27 sums
                } else {
                    lm_pixel(acc, rs[u] % W, rs[u] / W, v[u].w, make_float2(v[u].x, v[u].y), v[u].z, depth_add, fx, fy, cx, cy, Gm);
                }
            }
        }
        double mine;
        lm_block_reduce(acc, red, tid, mine);
        const uint32_t par = (uint32_t)(step & 1);
        if (tid < NACC) lm_st_cluster_f64(slot0 + ((par * LMC_CTAS + rank) * NACC + tid) * 8u, mine);
        lm_cluster_sync();                                  // every CTA's partials of this step are in CTA 0's slots
        if (tid < NACC) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < LMC_CTAS; ++r) s += lm_ld_cluster_f64(slot0 + ((par * LMC_CTAS + r) * NACC + tid) * 8u);
            tot[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            float Gn[12];
            lm_solve_retract(tot, Gm, Gn, ep, lm, nullptr, nullptr, nullptr);
#pragma unroll
            for (int i = 0; i < 12; ++i) Gs[i] = Gn[i];
        }
        __syncthreads();                                    // also orders red[] / tot[] reuse in the next step
    }
    if (rank == 0 && tid == 0) {
        float Gn[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) Gn[i] = Gs[i];
        store_G(G + b * 16, Gn);
    }
    lm_cluster_sync();                                      // CTA 0's shared memory stays alive until every CTA has read it
}

// Per-operator entries of the two small pieces of an LM step (tests pin their branches directly):
//   G <- exp(delta) G  (SE3.increment, geometry/transformation.py:110-115; se3.py:228-306 incl. the Taylor branch)
__global__ void se3_retract_kernel(const float* __restrict__ delta, float* __restrict__ G, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float xi[6], dG[12], Gm[12], Gn[12];
    for (int i = 0; i < 6; ++i) xi[i] = delta[b * 6 + i];
    for (int i = 0; i < 12; ++i) Gm[i] = G[b * 16 + i];
    se3_exp_f32(xi, dG);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = dG[i * 4 + 0] * Gm[0 * 4 + j] + dG[i * 4 + 1] * Gm[1 * 4 + j] + dG[i * 4 + 2] * Gm[2 * 4 + j];
            if (j == 3) s += dG[i * 4 + 3];
            Gn[i * 4 + j] = s;
        }
    store_G(G + b * 16, Gn);
}
//   x = clamp(nan_to_zero(H^-1 b))  (geometry/cholesky.py:32-50), H [B,6,6] fp64 (no damping added here)
__global__ void chol_solve_kernel(const double* __restrict__ H, const double* __restrict__ bvec, float* __restrict__ x, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double Hm[6][6], bv[6];
    for (int i = 0; i < 6; ++i) {
        bv[i] = bvec[b * 6 + i];
        for (int j = 0; j < 6; ++j) Hm[i][j] = H[(size_t)b * 36 + i * 6 + j];
    }
    float xi[6];
    chol_solve6(Hm, bv, xi);
    for (int i = 0; i < 6; ++i) x[b * 6 + i] = xi[i];
}

// ------------------------------------------------------------------------------------------------ backward of one LM step (f4)
// Gradient of a loss with respect to the correspondence target and weight through ONE damped Gauss-Newton step, i.e. what
// autograd computes in the reference for reprojction_optim(num_iters = 1) (the shipped OPTIM_ITER_COUNT,
// config/linemod/template_fw0.5.yml:81) given dL/d(delta):
//   geometry/cholesky.py:19-28  (OptNet backward of the solve): z = H_d^-1 dx, dL/dH_d = -x z^T, dL/db = z, with x the RAW
//     solution (before NaN -> 0 and the clamp, whose backward zeroes dx where x is NaN or outside [-1, 1], cholesky.py:42-45);
//   geometry/transformation.py:300: H_d = H + ep I + lm H (.) I  =>  dL/dH = dL/dH_d + lm diag(dL/dH_d);
//   :294-297: H = sum v w J^T J, b = sum v w J^T (target - x1)  =>  per pixel
//     dL/dw      = v sum_rows [ J_row dL/dH J_row^T + (J_row . dL/db) r_row ],   dL/dtarget_row = v w (J_row . dL/db).
// The pose entering the step is a constant here (PoseRefiner.py:320 detaches it: Tij.copy(stop_gradients=True)).
__global__ void lm_bwd_solve_kernel(const double* __restrict__ Hs, const double* __restrict__ bs, const float* __restrict__ grad_delta,
                                    double ep, double lm, int B, double* __restrict__ dHb /*[B][42]: dL/dH (36), dL/db (6)*/) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double Hm[6][6], bv[6];
    for (int i = 0; i < 6; ++i) {
        bv[i] = bs[b * 6 + i];
        for (int j = 0; j < 6; ++j) Hm[i][j] = Hs[(size_t)b * 36 + i * 6 + j];
    }
    for (int i = 0; i < 6; ++i) Hm[i][i] = Hm[i][i] + (ep + lm * Hm[i][i]);
    double L[6][6];
    for (int j = 0; j < 6; ++j) {
        double s = Hm[j][j];
        for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
        const double d = sqrt(s);
        L[j][j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double t = Hm[i][j];
            for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
            L[i][j] = t / d;
        }
    }
    auto solve = [&](const double* rhs, double* out) {
        double y[6];
        for (int i = 0; i < 6; ++i) { double t = rhs[i]; for (int k = 0; k < i; ++k) t -= L[i][k] * y[k]; y[i] = t / L[i][i]; }
        for (int i = 5; i >= 0; --i) { double t = y[i]; for (int k = i + 1; k < 6; ++k) t -= L[k][i] * out[k]; out[i] = t / L[i][i]; }
    };
    double x[6], dx[6], z[6];
    solve(bv, x);
    for (int i = 0; i < 6; ++i) {
        const bool pass = (x[i] == x[i]) && x[i] >= -1.0 && x[i] <= 1.0;       // where() and clamp() backward
        dx[i] = pass ? (double)grad_delta[b * 6 + i] : 0.0;
    }
    solve(dx, z);
    double* o = dHb + (size_t)b * 42;
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) {
            double g = -x[i] * z[j];
            if (i == j) g += lm * g;                 // d(H_d)/dH has (1 + lm) on the diagonal
            o[i * 6 + j] = g;
        }
        o[36 + i] = z[i];
    }
}

__global__ void __launch_bounds__(256) lm_bwd_pixel_kernel(const float* __restrict__ depth, const float* __restrict__ target,
                                                           const float* __restrict__ weight, const float* __restrict__ K,
                                                           const float* __restrict__ G, const double* __restrict__ dHb, int H, int W,
                                                           float depth_add, float* __restrict__ grad_target,
                                                           float* __restrict__ grad_weight) {
    const int b = blockIdx.y, N = H * W;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double sH[36], sb[6];
    if (threadIdx.x < 42) (threadIdx.x < 36 ? sH[threadIdx.x] : sb[threadIdx.x - 36]) = dHb[(size_t)b * 42 + threadIdx.x];
    __syncthreads();
    if (px >= N) return;
    const float* Kb = K + b * 9;
    const float fx = Kb[0], fy = Kb[4], cx = Kb[2], cy = Kb[5];
    float Gm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Gm[i] = G[b * 16 + i];
    const size_t idx = (size_t)b * N + px;
    const int u = px % W, v = px / W;
    const float Z = depth[idx] + depth_add;
    const float2 tg = reinterpret_cast<const float2*>(target)[idx];
    const float wv = weight[idx];
    // same geometry as lm_pixel (projective_ops.py:107-124, transformation.py:29-45,289)
    const float X = Z * ((float)u - cx) / fx, Y = Z * ((float)v - cy) / fy;
    const float X1 = Gm[0] * X + Gm[1] * Y + Gm[2] * Z + Gm[3];
    const float Y1 = Gm[4] * X + Gm[5] * Y + Gm[6] * Z + Gm[7];
    const float Z1 = Gm[8] * X + Gm[9] * Y + Gm[10] * Z + Gm[11];
    const double valid = (Z > 0.1f && Z1 > 0.1f) ? 1.0 : 0.0;
    const float Zc = fmaxf(Z1, 0.01f);
    const float x1 = fx * (X1 / Zc) + cx, y1 = fy * (Y1 / Zc) + cy;
    const bool cut = Zc <= 0.02f;
    const float zi1 = cut ? 0.f : 1.0f / Zc, zi2 = cut ? 0.f : 1.0f / (Zc * Zc);
    const double A = (double)(fx * zi1), C = (double)((-fx * X1) * zi2);
    const double Bq = (double)(fy * zi1), D = (double)((-fy * Y1) * zi2);
    const double dX = (double)X1, dY = (double)Y1, dZ = (double)Z1;
    double J[2][6];
    J[0][0] = A;   J[0][1] = 0.0; J[0][2] = C; J[0][3] = C * dY;              J[0][4] = A * dZ + C * (-dX); J[0][5] = A * (-dY);
    J[1][0] = 0.0; J[1][1] = Bq;  J[1][2] = D; J[1][3] = Bq * (-dZ) + D * dY; J[1][4] = D * (-dX);          J[1][5] = Bq * dX;
    const double r[2] = {(double)tg.x - (double)x1, (double)tg.y - (double)y1};
    double gw = 0.0, gt[2];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
        double jb = 0.0, q = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            jb += J[row][i] * sb[i];
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) t += sH[i * 6 + j] * J[row][j];
            q += J[row][i] * t;
        }
        gw += q + jb * r[row];
        gt[row] = valid * (double)wv * jb;
    }
    grad_weight[idx] = (float)(valid * gw);
    reinterpret_cast<float2*>(grad_target)[idx] = make_float2((float)gt[0], (float)gt[1]);
}

__global__ void zero_u32_kernel(unsigned* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

}  // namespace

static inline int lm_nblk(int H, int W) { return ceil_div(H * W, LM_PX_PER_BLOCK); }
static inline size_t lm_partials_bytes(int B, int H, int W) {
    const size_t a = (size_t)B * lm_nblk(H, W) * NACC * sizeof(double);
    const size_t m = (size_t)2 * B * LM_MAX_BLK_PER_SAMPLE * NACC * sizeof(double);
    return align_up(a > m ? a : m, 256);
}

size_t b2p_lm_ws_bytes(int B, int H, int W) {
    return lm_partials_bytes(B, H, W) + align_up((size_t)2 * B * sizeof(unsigned), 256);
}

// The per-sample counters must be zero before the first use of a workspace; every kernel leaves them zero again.
int b2p_lm_reset(void* ws, int B, int H, int W, cudaStream_t s) {
    unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + lm_partials_bytes(B, H, W));
    zero_u32_kernel<<<ceil_div(2 * B, 256), 256, 0, s>>>(counters, 2 * B);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_lm_step(const float* depth, const float* target, const float* weight, const float* K, float* G, int B, int H,
                int W, float depth_add, double ep, double lm, double* H_out, double* b_out, float* delta_out, void* ws,
                cudaStream_t s) {
    const int nblk = lm_nblk(H, W);
    double* partials = reinterpret_cast<double*>(ws);
    unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + lm_partials_bytes(B, H, W));
    dim3 grid(nblk, B);
    lm_step_kernel<<<grid, LM_THREADS, 0, s>>>(depth, target, weight, K, G, B, H, W, depth_add, ep, lm, partials, counters, nblk,
                                               H_out, b_out, delta_out);
    B2P_LAUNCH_CHECK();
    return 0;
}

// All n_steps in one launch when the blocks can be co-resident; otherwise n_steps single-step launches.
int b2p_lm_steps(const float* depth, const float* target, const float* weight, const float* K, float* G, int B, int H,
                 int W, float depth_add, double ep, double lm, int n_steps, void* ws, cudaStream_t s, const int* fg_idx,
                 const int* fg_count) {
    if (n_steps <= 0) return 0;
    // with a foreground list: the cluster kernel (hardware co-scheduled CTAs, barrier.cluster) instead of the grid-wide spin
    // barrier of lm_multi_kernel below (option lm_cluster = 0 keeps the latter)
    if (fg_idx && fg_count && b2p_options().lm_cluster != 0)
        return b2p_lm_cluster(nullptr, depth, target, weight, fg_idx, fg_count, K, G, B, H, W, depth_add, ep, lm, n_steps, s);
    int dev = 0, sms = 0, per_sm = 0;
    B2P_CUDA(cudaGetDevice(&dev));
    B2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lm_multi_kernel, LM_THREADS, 0));
    const int capacity = sms * per_sm;
    int nb = capacity / B;
    if (nb > LM_MAX_BLK_PER_SAMPLE) nb = LM_MAX_BLK_PER_SAMPLE;
    const int useful = ceil_div(H * W, LM_PX_PER_BLOCK);
    if (nb > useful) nb = useful;
    if (nb < 1 || n_steps == 1) {
        for (int i = 0; i < n_steps; ++i) {
            int rc = b2p_lm_step(depth, target, weight, K, G, B, H, W, depth_add, ep, lm, nullptr, nullptr, nullptr, ws, s);
            if (rc) return rc;
        }
        return 0;
    }
    double* partials = reinterpret_cast<double*>(ws);
    unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) + lm_partials_bytes(B, H, W));
    dim3 grid(nb, B);
    B2P_CUDA(b2p_launch_pdl(lm_multi_kernel, grid, dim3(LM_THREADS), 0, s, depth, target, weight, K, G, B, H, W, depth_add, ep, lm, n_steps,
                            partials, counters, fg_idx, fg_count));
    B2P_LAUNCH_CHECK();
    return 0;
}

// All n_steps of one recurrent iteration over the foreground list, one cluster per sample: on the records of the foreground
// pipeline (rec != nullptr), or gathering from the dense maps.
int b2p_lm_cluster(const float4* rec, const float* depth, const float* target, const float* weight, const int* fg_idx, const int* fg_count,
                   const float* K, float* G, int B, int H, int W, float depth_add, double ep, double lm, int n_steps, cudaStream_t s) {
    if (n_steps <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * LMC_CTAS)); cfg.blockDim = dim3(LM_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = b2p_pdl_allowed(5);
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = LMC_CTAS; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    const int N = H * W;
    const bool noacc = b2p_options().lm_debug == 1;
    if (rec) {
        if (noacc) B2P_CUDA(cudaLaunchKernelEx(&cfg, lm_cluster_kernel<true, true>, rec, depth, target, weight, fg_idx, fg_count, K, G, N, W, depth_add, ep, lm, n_steps));
        else B2P_CUDA(cudaLaunchKernelEx(&cfg, lm_cluster_kernel<false, true>, rec, depth, target, weight, fg_idx, fg_count, K, G, N, W, depth_add, ep, lm, n_steps));
    } else {
        if (noacc) B2P_CUDA(cudaLaunchKernelEx(&cfg, lm_cluster_kernel<true, false>, rec, depth, target, weight, fg_idx, fg_count, K, G, N, W, depth_add, ep, lm, n_steps));
        else B2P_CUDA(cudaLaunchKernelEx(&cfg, lm_cluster_kernel<false, false>, rec, depth, target, weight, fg_idx, fg_count, K, G, N, W, depth_add, ep, lm, n_steps));
    }
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_se3_retract(const float* delta, float* G, int B, cudaStream_t s) {
    se3_retract_kernel<<<ceil_div(B, 64), 64, 0, s>>>(delta, G, B);
    B2P_LAUNCH_CHECK();
    return 0;
}

int b2p_chol_solve(const double* H, const double* b, float* x, int B, cudaStream_t s) {
    chol_solve_kernel<<<ceil_div(B, 64), 64, 0, s>>>(H, b, x, B);
    B2P_LAUNCH_CHECK();
    return 0;
}

// f4: gradients of one LM step with respect to target and weight.  ws: b2p_lm_bwd_ws_bytes.
size_t b2p_lm_bwd_ws_bytes(int B, int H, int W) {
    return b2p_lm_ws_bytes(B, H, W) + align_up((size_t)B * (36 + 6 + 42) * sizeof(double), 256) + align_up((size_t)B * (16 + 6) * sizeof(float), 256);
}

int b2p_lm_backward(const float* depth, const float* target, const float* weight, const float* K, const float* G, const float* grad_delta,
                    int B, int H, int W, float depth_add, double ep, double lm, float* grad_target, float* grad_weight, void* ws,
                    cudaStream_t s) {
    char* p = reinterpret_cast<char*>(ws);
    void* lm_ws = p; p += b2p_lm_ws_bytes(B, H, W);
    double* Hs = reinterpret_cast<double*>(p); double* bs = Hs + (size_t)B * 36; double* dHb = bs + (size_t)B * 6;
    p += align_up((size_t)B * (36 + 6 + 42) * sizeof(double), 256);
    float* Gtmp = reinterpret_cast<float*>(p); float* dtmp = Gtmp + (size_t)B * 16;
    int rc;
    // forward normal equations of the step (un-damped H, b) on a scratch copy of the pose
    B2P_CUDA(cudaMemcpyAsync(Gtmp, G, (size_t)B * 16 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if ((rc = b2p_lm_reset(lm_ws, B, H, W, s))) return rc;
    if ((rc = b2p_lm_step(depth, target, weight, K, Gtmp, B, H, W, depth_add, ep, lm, Hs, bs, dtmp, lm_ws, s))) return rc;
    lm_bwd_solve_kernel<<<ceil_div(B, 64), 64, 0, s>>>(Hs, bs, grad_delta, ep, lm, B, dHb);
    B2P_LAUNCH_CHECK();
    lm_bwd_pixel_kernel<<<dim3((unsigned)ceil_div(H * W, 256), (unsigned)B), 256, 0, s>>>(depth, target, weight, K, G, dHb, H, W, depth_add,
                                                                                         grad_target, grad_weight);
    B2P_LAUNCH_CHECK();
    return 0;
}
