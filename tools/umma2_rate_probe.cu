// tcgen05.mma.cta_group::2 kind::tf32 rate (M = 256 over a CTA pair) for several N; SS mode, K-major SWIZZLE_NONE operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma2_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// N = total MMA N (each CTA supplies N/2 columns of B)
__global__ void __launch_bounds__(128) rate2(long long* out, int N, int n_mma, int n_acc, int a_distinct) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 8);
    float* A = reinterpret_cast<float*>(smem + 128);           // 8 x [128 x 8] tiles (4 KB each)
    float* B = A + 8 * 1024;                                   // 8 x [N/2 x 8] tiles
    uint32_t crank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    for (int i = threadIdx.x; i < 8 * 1024 + 8 * 128 * 8; i += 128) A[i] = 0.001f * (i % 13);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (threadIdx.x < 32) {
        long long t0 = clock64(), t1 = t0;
        if (crank == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t lbo_b = (N / 2 / 8) * 128;
            const uint64_t a0 = make_desc(smem_u32(A), 2048, 128);
            const uint64_t b0 = make_desc(smem_u32(B), lbo_b, 128);
            const bool leader = elect_one();
            for (int i = 0; i < n_mma; i += 6) {
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    const uint64_t a = a0 + (uint64_t)(a_distinct ? u * (4096 >> 4) : 0);
                    const uint64_t b = b0 + (uint64_t)(a_distinct ? u * ((2 * lbo_b) >> 4) : 0);
                    // n_acc 11: per 6 MMAs {0,0,1,0,0,2} (switch 4x); n_acc 12: {1,1,0,0,0,0} (switch 2x); otherwise round robin over n_acc sets
                    const int g6 = u;
                    const uint32_t set = n_acc == 11 ? (g6 == 2 ? 1u : (g6 == 5 ? 2u : 0u)) : (n_acc == 12 ? (g6 < 2 ? 1u : 0u) : (uint32_t)((n_acc >= 2) ? (u % n_acc) : 0));
                    const uint32_t d = tm + set * (uint32_t)N;
                    if (leader) mma2_tf32(d, a, b, idesc, i > 0);
                }
            }
            t1 = clock64();
            if (leader) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "h"((uint16_t)3) : "memory");
            __syncwarp();
        }
        asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(mbar)), "r"(0) : "memory");
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 16); long long h[2];
    const int smem = 128 + (8 * 1024 + 8 * 128 * 8) * 4;
    cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {2, 148})
        for (int N : {96, 160})
            for (int nacc : {1, 11, 12})
                for (int dist : {1}) {
                    if (3 * N > 512 && nacc > 1) continue;
                    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    cudaError_t e0 = cudaLaunchKernelEx(&cfg, rate2, d, N, 2046, nacc, dist);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("grid %3d N %3d n_acc %d distinct %d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)  %s %s\n", grid, N, nacc, dist, h[0] / 2048.0, h[1] / 2048.0,
                           128.0 * N / 256, cudaGetErrorString(e0), cudaGetErrorString(e));
                }
    return 0;
}
