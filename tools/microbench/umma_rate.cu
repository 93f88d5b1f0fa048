// Microbenchmark: how fast does one SM execute back-to-back tcgen05.mma.kind::f16 (SS operands, K = 16) instructions?
// No TMA, no epilogue: operands are zero-filled shared memory, the issuing lane times n MMAs from first issue to the commit's
// arrival.  Variants: N, single CTA (M = 128) or CTA pair (M = 256, cta_group::2), and the operand pattern of conv_umma.cu
// (3 MMAs per K step over hi/lo planes, 4 K steps per 64-channel stage) or one fixed operand pair.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred;
}
template <int NCTA>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (NCTA == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int NCTA>
__global__ void __launch_bounds__(128, 1) rate_kernel(int stages, int N, int pattern, int ld_epilogue, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (NCTA == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t a_plane = 16384, b_plane = (uint32_t)(N / NCTA) * 128u;
    const uint32_t a0 = base, b0 = base + 2 * a_plane, bar = b0 + 2 * b_plane, slot = bar + 16;
    const bool rnd = pattern >= 10;
    pattern %= 10;
    for (uint32_t i = threadIdx.x * 16; i < 2 * a_plane + 2 * b_plane; i += blockDim.x * 16) {
        uint32_t v[4];
        for (int j = 0; j < 4; ++j) {
            uint32_t h = (i + j * 4 + blockIdx.x * 7919u) * 2654435761u;
            // two fp16 values in (-1, 1): sign | exponent 01110/01101 | random mantissa
            v[j] = rnd ? ((h & 0x83FF83FFu) | 0x38003400u) : 0u;
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + i), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (NCTA == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (((128u * NCTA) >> 4) << 24);
        const uint64_t ah = desc_sw128(a0), al = desc_sw128(a0 + a_plane), bh = desc_sw128(b0), bl = desc_sw128(b0 + b_plane);
        long long t0 = clock64();
        for (int s = 0; s < stages; ++s) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)(k * 2);
                    if (pattern == 1 || pattern == 3) {
                        mma<NCTA>(tmem, al + adv, bh + adv, idesc, (s | k) != 0);
                        mma<NCTA>(tmem, ah + adv, bl + adv, idesc, 1);
                        mma<NCTA>(tmem, ah + adv, bh + adv, idesc, 1);
                    } else if (pattern == 2) {                 // alternate two accumulators
                        mma<NCTA>(tmem, al + adv, bh + adv, idesc, (s | k) != 0);
                        mma<NCTA>(tmem + 256, ah + adv, bl + adv, idesc, (s | k) != 0);
                        mma<NCTA>(tmem, ah + adv, bh + adv, idesc, 1);
                    } else {
                        mma<NCTA>(tmem, ah, bh, idesc, (s | k) != 0);
                        mma<NCTA>(tmem, ah, bh, idesc, 1);
                        mma<NCTA>(tmem, ah, bh, idesc, 1);
                    }
                }
            }
            if (pattern == 3) {      // per-stage commit + wait for the previous stage's commit + fence, as the pipelined kernel does
                if (elect_one())
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8) : "memory");
                __syncwarp();
                if (s >= 1) {
                    uint32_t ok = 0, sp = 0;
                    while (!ok && ++sp < (1u << 22))
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(ok) : "r"(bar + 8), "r"((uint32_t)((s - 1) & 1)) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one()) {
            if (NCTA == 2)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(bar), "h"((uint16_t)1) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        uint32_t done = 0, spins = 0;
        while (!done && ++spins < (1u << 24))
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    } else if (ld_epilogue && warp != 1) {
        // concurrent TMEM reads like an epilogue would issue (other accumulator half), to see the interference
        uint32_t r[32];
        const int q = warp & 3;
        for (int it = 0; it < ld_epilogue; ++it) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)((it & 7) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (r[0] == 0x12345678u) out[0] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (NCTA == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    else __syncthreads();
    if (warp == 1) {
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int NCTA>
double run(int grid, int stages, int N, int pattern, int ld_epi) {
    unsigned long long* d;
    cudaMalloc(&d, grid * sizeof(unsigned long long));
    cudaMemset(d, 0, grid * sizeof(unsigned long long));
    const size_t smem = 2 * 16384 + 2 * (size_t)(N / NCTA) * 128 + 2048;
    cudaFuncSetAttribute(rate_kernel<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<NCTA>, stages, N, pattern, ld_epi, d);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -1; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return -1; }
    }
    std::vector<unsigned long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    double mx = 0;
    for (int i = 0; i < grid; i += NCTA) mx = h[i] > mx ? (double)h[i] : mx;
    return mx / (stages * 12.0);
}

int main() {
    const int stages = 400;
    printf("clocks per MMA (K=16), max over CTAs; ideal = N/2 per SM\n");
    printf("%-34s %8s %8s %8s %8s\n", "variant", "N=64", "N=128", "N=192", "N=256");
    const int Ns[4] = {64, 128, 192, 256};
    struct V { const char* name; int ncta, grid, pattern, ld; } vs[] = {
        {"1 CTA, M128, fixed operands", 1, 1, 0, 0},   {"148 CTAs, M128, fixed operands", 1, 148, 0, 0},
        {"148 CTAs, M128, hi/lo pattern", 1, 148, 1, 0}, {"148 CTAs, M128, two accumulators", 1, 148, 2, 0},
        {"148 CTAs, M128, hi/lo + tcgen05.ld", 1, 148, 1, 40000},
        {"148 CTAs, M128, hi/lo RANDOM data", 1, 148, 11, 0},
        {"148 CTAs, M128, commit+wait prev", 1, 148, 3, 0},
        {"148 CTAs, M128, same, RANDOM", 1, 148, 13, 0},
        {"2 CTAs pair, M256, hi/lo", 2, 2, 1, 0},     {"148 CTAs pairs, M256, hi/lo", 2, 148, 1, 0},
        {"148 CTAs pairs, M256, hi/lo RANDOM", 2, 148, 11, 0},
    };
    for (auto& v : vs) {
        printf("%-34s", v.name);
        for (int n : Ns) {
            double c = v.ncta == 2 ? run<2>(v.grid, stages, n, v.pattern, v.ld) : run<1>(v.grid, stages, n, v.pattern, v.ld);
            printf(" %8.1f", c);
        }
        printf("\n");
    }
    return 0;
}
