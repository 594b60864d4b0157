// tcgen05.mma kind::tf32 issue/execute rate for several N (SS mode, K-major SWIZZLE_NONE operands in smem).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(128) rate(long long* out, int N, int n_mma, int n_acc, int a_distinct, int b_sw128) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 8);
    float* A = reinterpret_cast<float*>(smem + 1024);          // 8 x [128 x 8] tiles (4 KB each)
    float* B = reinterpret_cast<float*>(smem + 1024 + 8 * 4096);   // 8 x [N x 8] tiles (1024 B aligned for the swizzled variant)
    for (int i = threadIdx.x; i < 8 * 1024 + 8 * 256 * 8; i += 128) A[i] = 0.001f * (i % 13);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lbo_b = (N / 8) * 128;
        long long t0 = clock64();
        const uint64_t a0 = make_desc(smem_u32(A), 2048, 128);
        // b_sw128: B as K-major SWIZZLE_128B atoms (8 columns x 32 k = 1024 B), as the jet kernels keep their activation planes
        const uint64_t b0 = b_sw128 ? ((uint64_t)((smem_u32(B) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((1024 >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61))
                                    : make_desc(smem_u32(B), lbo_b, 128);
        const bool leader = elect_one();
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // descriptor low words advance by (bytes >> 4); operands rotate over 8 tiles when a_distinct
                const uint64_t a = a0 + (uint64_t)(a_distinct ? u * (4096 >> 4) : 0);
                const uint64_t b = b0 + (uint64_t)(b_sw128 ? ((u >> 2) * ((N / 8) * 1024) + (u & 3) * 32) >> 4 : (a_distinct ? u * ((2 * lbo_b) >> 4) : 0));
                const uint32_t d = tm + (uint32_t)(((n_acc == 2) ? (u & 1) : 0) * N);
                if (leader) mma_tf32(d, a, b, idesc, (i + u) >= n_acc);
            }
        }
        long long t1 = clock64();
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
        __syncwarp();
        asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(mbar)), "r"(0) : "memory");
        long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 16); long long h[2];
    const int smem = 1024 + (8 * 1024 + 8 * 256 * 8) * 4;
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {16, 32, 64, 80, 96, 128, 160, 256})
        for (int nacc : {1, 2})
            for (int dist : {1, 2}) {
                if (nacc * N > 512) continue;
                rate<<<1, 128, smem>>>(d, N, 2048, nacc, 1, dist == 2);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N %3d n_acc %d B layout %d (1 SWIZZLE_NONE, 2 SWIZZLE_128B): issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)  %s\n", N, nacc, dist, h[0] / 2048.0, h[1] / 2048.0, 128.0 * N / 256, cudaGetErrorString(e));
            }
    return 0;
}
