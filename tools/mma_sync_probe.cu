// Throughput probe of the legacy warp-level tensor path on B200: mma.sync.m16n8k8 TF32 (and bf16 m16n8k16).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int MI, int NI>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
    float c[MI][NI][4];
    uint32_t a[MI][4], b[NI][2];
    for (int i = 0; i < MI; ++i) for (int r = 0; r < 4; ++r) a[i][r] = 0x3f800000u + threadIdx.x + i;
    for (int j = 0; j < NI; ++j) for (int r = 0; r < 2; ++r) b[j][r] = 0x3f000000u + threadIdx.x + j;
    for (int i = 0; i < MI; ++i) for (int j = 0; j < NI; ++j) for (int r = 0; r < 4; ++r) c[i][j][r] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                if (KIND == 0) mma_tf32(c[i][j], a[i], b[j]); else mma_bf16(c[i][j], a[i], b[j]);
            }
    }
    float s = 0.f;
    for (int i = 0; i < MI; ++i) for (int j = 0; j < NI; ++j) for (int r = 0; r < 4; ++r) s += c[i][j][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int MI, int NI>
void run(const char* name, float* out, int warps_per_sm) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    int threads = 256, grid = 148 * warps_per_sm / 8;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); probe<KIND, MI, NI><<<grid, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double macs = (KIND == 0 ? 1024.0 : 2048.0) * MI * NI * (double)iters * (threads / 32) * grid;
    printf("%-34s warps/SM %2d: %.1f TFLOP/s  (%s)\n", name, warps_per_sm, 2 * macs / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    for (int w : {8, 16, 32}) {
        run<0, 4, 10>("mma.sync tf32 m16n8k8 (4x10 tile)", out, w);
        run<0, 2, 4>("mma.sync tf32 m16n8k8 (2x4 tile)", out, w);
        run<1, 4, 10>("mma.sync bf16 m16n8k16 (4x10 tile)", out, w);
    }
    return 0;
}
