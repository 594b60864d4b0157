// tcgen05 bring-up probe: D[128 x N] = A[128 x K] (K-major, weights) * B[N x K]^T (MN-major, activations),
// kind::tf32, operands in shared memory (no swizzle, canonical core-matrix layouts), accumulator in TMEM.
// Checks descriptor encodings and the 3xTF32 split against a double-precision host reference.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 80, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

// mode 0: single TF32 (hi*hi); mode 1: 3xTF32
__global__ void __launch_bounds__(128) umma_probe(const float* __restrict__ Ag, const float* __restrict__ Bg,
                                                  float* __restrict__ Dg, int mode, int bmajor) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // canonical layouts
    float* Ahi = reinterpret_cast<float*>(smem);            // M*K floats
    float* Alo = Ahi + M * K;
    float* Bhi = Alo + M * K;                                // N*K floats
    float* Blo = Bhi + N * K;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Blo + N * K);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    // A (K-major): element (m,k) at (k/4)*LBO_A + (m/8)*SBO_A + (m%8)*16 + (k%4)*4 bytes
    constexpr uint32_t SBO_A = 128, LBO_A = (M / 8) * 128;
    // B (MN-major): element (n,k) at (k/8)*LBO_B + (n/4)*SBO_B + (k%8)*16 + (n%4)*4 bytes
    constexpr uint32_t SBO_B = 128, LBO_B = (N / 4) * 128;
    for (int i = tid; i < M * K; i += 128) {
        const int m = i / K, k = i % K;
        const float v = Ag[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int off = ((k / 4) * LBO_A + (m / 8) * SBO_A + (m % 8) * 16 + (k % 4) * 4) / 4;
        Ahi[off] = hi;
        Alo[off] = v - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float v = Bg[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        // K-major alternative for B: same core-matrix form as A with N rows
        const int off = bmajor ? ((k / 8) * LBO_B + (n / 4) * SBO_B + (k % 8) * 16 + (n % 4) * 4) / 4
                               : ((k / 4) * ((N / 8) * 128) + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4;
        Bhi[off] = hi;
        Blo[off] = v - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy operand writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    if (tid == 0) printf("tmem base 0x%08x  Ahi smem 0x%x  Bhi smem 0x%x\n", tmem_d, smem_u32(Ahi), smem_u32(Bhi));
    {   // TMEM store/load round trip on column 100..107
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) v[j] = 1000 * tid + j;
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + 100;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 8) {
            const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
            const uint32_t sv = __float_as_uint(7.0f);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(sv) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 37) printf("tmem roundtrip thread 37: %u %u .. %u (expect 37000 37001 .. 37007)\n", r[0], r[1], r[7]);
    }

    if (tid == 0) {
        // instruction descriptor: c=F32, a=b=TF32, A K-major, B MN-major, N, M
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)(bmajor ? 1 : 0) << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t a_hi = make_desc(smem_u32(Ahi) + ks * 2 * LBO_A, LBO_A, SBO_A);
            const uint64_t a_lo = make_desc(smem_u32(Alo) + ks * 2 * LBO_A, LBO_A, SBO_A);
            const uint32_t LBK = (N / 8) * 128;
            const uint64_t b_hi = bmajor ? (bmajor == 2 ? make_desc(smem_u32(Bhi) + ks * LBO_B, SBO_B, LBO_B) : make_desc(smem_u32(Bhi) + ks * LBO_B, LBO_B, SBO_B)) : make_desc(smem_u32(Bhi) + ks * 2 * LBK, LBK, 128);
            const uint64_t b_lo = bmajor ? (bmajor == 2 ? make_desc(smem_u32(Blo) + ks * LBO_B, SBO_B, LBO_B) : make_desc(smem_u32(Blo) + ks * LBO_B, LBO_B, SBO_B)) : make_desc(smem_u32(Blo) + ks * 2 * LBK, LBK, 128);
            if (mode == 1) {
                mma_tf32(tmem_d, a_lo, b_hi, idesc, acc);
                mma_tf32(tmem_d, a_hi, b_lo, idesc, 1);
                acc = 1;
            }
            mma_tf32(tmem_d, a_hi, b_hi, idesc, acc);
            acc = 1;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(mbar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w reads TMEM lanes [32w, 32w+32); thread = one row m of D
    const int m = tid;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t r[8];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) Dg[m * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 0.25f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (2 * M * K + 2 * N * K) * 4 + 64;
    cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 2; mode < 6; ++mode) {
        cudaMemset(dD, 0, D.size() * 4);
        umma_probe<<<1, 128, smem>>>(dA, dB, dD, mode & 1, mode >> 1);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0, maxerr = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
                const double d = D[m * N + n] - ref;
                num += d * d; den += ref * ref; maxerr = fmax(maxerr, fabs(d));
            }
        printf("bmajor %d mode %d (%s): relL2 %.3e  max|err| %.3e  D[0][0]=%f D[5][7]=%f D[127][79]=%f\n", mode >> 1, mode & 1, (mode & 1) ? "3xTF32" : "1xTF32",
               sqrt(num / den), maxerr, D[0], D[5 * N + 7], D[127 * N + 79]);
    }
    return 0;
}
