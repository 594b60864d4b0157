// FP32 peak probe for B200 (sm_100a): scalar FFMA vs packed fma.rn.f32x2.
// Prints achieved TFLOP/s so the roofline denominator for the FFMA jet kernels is measured, not assumed.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int ILP>
__global__ void __launch_bounds__(512) ffma_scalar(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// variant with three distinct register operands (a,b per-accumulator distinct) like a GEMM outer product
template <int RA, int RB>
__global__ void __launch_bounds__(512) ffma_outer(float* out, int iters, const float* __restrict__ in) {
    float acc[RA][RB];
    float av[RA], bv[RB];
#pragma unroll
    for (int i = 0; i < RA; ++i) av[i] = in[threadIdx.x % 32 + i];
#pragma unroll
    for (int j = 0; j < RB; ++j) bv[j] = in[threadIdx.x % 32 + 64 + j];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < RA; ++i)
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        // perturb operands cheaply so the loop is not hoisted
        av[it & (RA - 1)] += 1e-7f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint64_t pack2(float x, float y) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// packed outer product: acc[i][j/2] pairs over j; a broadcast pair (a_i,a_i); b natural pair
template <int RA, int RB2>
__global__ void __launch_bounds__(512) ffma2_outer(float* out, int iters, const float* __restrict__ in) {
    uint64_t acc[RA][RB2];
    uint64_t av[RA], bv[RB2];
#pragma unroll
    for (int i = 0; i < RA; ++i) { float v = in[threadIdx.x % 32 + i]; av[i] = pack2(v, v); }
#pragma unroll
    for (int j = 0; j < RB2; ++j) bv[j] = pack2(in[threadIdx.x % 32 + 64 + 2 * j], in[threadIdx.x % 32 + 65 + 2 * j]);
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB2; ++j) acc[i][j] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < RA; ++i)
#pragma unroll
            for (int j = 0; j < RB2; ++j) fma2(acc[i][j], av[i], bv[j]);
        av[it & (RA - 1)] ^= 1ull;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < RB2; ++j) { s += __uint_as_float((uint32_t)acc[i][j]) + __uint_as_float((uint32_t)(acc[i][j] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d\n", p.name, sms);
    float *out, *in; cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024); cudaMalloc(&in, 4096);
    cudaMemset(in, 0, 4096);
    const int iters = 20000;
    for (int threads : {256, 512}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 1024 && false) continue;
            int grid = sms * bps;
            {
                double ms = time_ms([&] { ffma_scalar<16><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
                double fl = 2.0 * 16 * iters * (double)threads * grid;
                printf("scalar ILP16 thr %d bps %d: %.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { ffma_outer<8, 8><<<grid, threads>>>(out, iters / 4, in); }, 5);
                double fl = 2.0 * 64 * (iters / 4) * (double)threads * grid;
                printf("outer 8x8 scalar thr %d bps %d: %.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { ffma2_outer<8, 4><<<grid, threads>>>(out, iters / 4, in); }, 5);
                double fl = 2.0 * 64 * (iters / 4) * (double)threads * grid;
                printf("outer 8x8 f32x2 thr %d bps %d: %.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
