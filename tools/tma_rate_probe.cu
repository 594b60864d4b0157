// L2 -> shared-memory intake rate per SM: cp.async.bulk (1D) vs cp.async.bulk.tensor.2d (tensor map), 16 KB per copy, S stages
// in flight, every CTA streams the same 512 KB region (like the weight images).  Build: nvcc -arch=sm_100a -o tma_rate_probe ...
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr int kBytes = 16384;
__global__ void __launch_bounds__(128) probe(long long* out, const float* src, const __grid_constant__ CUtensorMap tmap, int mode, int stages, int iters, int issuers) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* buf = smem + 1024;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // issuer w (one lane of warp w) owns stages w, w + issuers, ...
    const int w = threadIdx.x >> 5;
    long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        auto issue = [&](int s, int it) {
            const int chunk = it & 31;     // 32 chunks of 16 KB = 512 KB
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(kBytes) : "memory");
            if (mode == 0) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + (size_t)s * kBytes)),
                             "l"(src + (size_t)chunk * (kBytes / 4)), "r"(kBytes), "r"(smem_u32(&bars[s])) : "memory");
            } else {
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(buf + (size_t)s * kBytes)),
                             "l"(&tmap), "r"(0), "r"(chunk * 64), "r"(smem_u32(&bars[s])) : "memory");
            }
        };
        int it = w;
        for (int s = w; s < stages; s += issuers, it += issuers) issue(s, it);
        uint32_t ph = 0;
        int s = w;
        for (int k = w; k < iters; k += issuers) {
            mbar_wait(&bars[s], ph);
            if (k + stages < iters) issue(s, k + stages);
            s += issuers;
            if (s >= stages) { s = w; ph ^= 1; }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    float* src; cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
    long long* d; cudaMalloc(&d, 16); long long h = 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    CUtensorMap tmap;
    cuuint64_t dims[2] = {64, (1 << 20) / 256}; cuuint64_t strides[1] = {256}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d\n", (int)r);
    const int iters = 4096;
    for (int grid : {1, 148})
        for (int mode : {0, 1})
            for (int stages : {2, 4, 8, 12})
                for (int issuers : {1, 2, 4}) {
                    if (stages % issuers) continue;
                    const int smem = 1024 + stages * kBytes;
                    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    probe<<<grid, 128, smem>>>(d, src, tmap, mode, stages, iters, issuers);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    printf("grid %3d %s stages %2d issuers %d: %.1f B/clk per SM  %s\n", grid, mode ? "tensor-2d" : "bulk-1d  ", stages, issuers,
                           (double)iters * kBytes / (double)h, cudaGetErrorString(e));
                }
    return 0;
}
