// Inner-loop probe: the register-tile GEMM of the jet kernels in isolation (operands from shared memory),
// scalar FFMA vs packed fma.rn.f32x2 variants.  Picks the inner-loop form with the best sustained rate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#ifndef UNR
#define UNR 2
#endif
constexpr int kUnr = UNR;
constexpr int W = 256, PITCH = 164, KC = 16, PJ = 20;

__device__ __forceinline__ uint64_t pack2(float x, float y) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

template <int V>
__global__ void __launch_bounds__(256, 1) probe(float* out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* A = sm;                    // [W][PITCH]
    float* Wc = sm + W * PITCH;       // [KC][W] (x2 when duplicated)
    const int tid = threadIdx.x, g = tid % 32, pg = tid / 32, col0 = pg * PJ;
    for (int i = tid; i < W * PITCH; i += 256) A[i] = (i % 7) * 0.01f;
    for (int i = tid; i < KC * W * 2; i += 256) Wc[i] = (i % 5) * 0.02f;
    __syncthreads();
    float res = 0.f;
    if (V == 0) {
        float acc[PJ][8];
#pragma unroll
        for (int i = 0; i < PJ; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it) {
            const float* Arow = A + (it % 16) * KC * PITCH + col0;
#pragma unroll (kUnr)
            for (int kk = 0; kk < KC; ++kk) {
                float a[PJ];
#pragma unroll
                for (int v = 0; v < PJ / 4; ++v) { float4 t = *(const float4*)(Arow + kk * PITCH + 4 * v); a[4*v]=t.x; a[4*v+1]=t.y; a[4*v+2]=t.z; a[4*v+3]=t.w; }
                float4 w0 = *(const float4*)(Wc + kk * W + 4 * g), w1 = *(const float4*)(Wc + kk * W + 128 + 4 * g);
                float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < PJ; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < PJ; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) res += acc[i][j];
    } else if (V == 1 || V == 2) {
        // pairs along rows: acc2[i][j] = (acc[2i][j], acc[2i+1][j]); a natural pairs; w duplicated
        uint64_t acc[PJ / 2][8];
#pragma unroll
        for (int i = 0; i < PJ / 2; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
        for (int it = 0; it < iters; ++it) {
            const float* Arow = A + (it % 16) * KC * PITCH + col0;
#pragma unroll (kUnr)
            for (int kk = 0; kk < KC; ++kk) {
                uint64_t a2[PJ / 2];
#pragma unroll
                for (int v = 0; v < PJ / 4; ++v) { ulonglong2 t = *(const ulonglong2*)(Arow + kk * PITCH + 4 * v); a2[2*v] = t.x; a2[2*v+1] = t.y; }
                uint64_t w2[8];
                if (V == 1) {
                    float4 w0 = *(const float4*)(Wc + kk * W + 4 * g), w1 = *(const float4*)(Wc + kk * W + 128 + 4 * g);
                    w2[0]=pack2(w0.x,w0.x); w2[1]=pack2(w0.y,w0.y); w2[2]=pack2(w0.z,w0.z); w2[3]=pack2(w0.w,w0.w);
                    w2[4]=pack2(w1.x,w1.x); w2[5]=pack2(w1.y,w1.y); w2[6]=pack2(w1.z,w1.z); w2[7]=pack2(w1.w,w1.w);
                } else {
                    // duplicated weights in smem: [kk][2W]: 4 x LDS.128
                    const float* wr = Wc + kk * 2 * W;
                    ulonglong2 t0 = *(const ulonglong2*)(wr + 8 * g), t1 = *(const ulonglong2*)(wr + 8 * g + 4);
                    ulonglong2 t2 = *(const ulonglong2*)(wr + 256 + 8 * g), t3 = *(const ulonglong2*)(wr + 256 + 8 * g + 4);
                    w2[0]=t0.x; w2[1]=t0.y; w2[2]=t1.x; w2[3]=t1.y; w2[4]=t2.x; w2[5]=t2.y; w2[6]=t3.x; w2[7]=t3.y;
                }
#pragma unroll
                for (int i = 0; i < PJ / 2; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) fma2(acc[i][j], a2[i], w2[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < PJ / 2; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { float x, y; unpack2(acc[i][j], x, y); res += x + y; }
    } else {
        // V3: pairs along neurons: acc2[i][j2] = (acc[i][2j2], acc[i][2j2+1]); w natural pairs; a duplicated
        uint64_t acc[PJ][4];
#pragma unroll
        for (int i = 0; i < PJ; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
        for (int it = 0; it < iters; ++it) {
            const float* Arow = A + (it % 16) * KC * PITCH + col0;
#pragma unroll (kUnr)
            for (int kk = 0; kk < KC; ++kk) {
                uint64_t a2[PJ];
#pragma unroll
                for (int v = 0; v < PJ / 4; ++v) { float4 t = *(const float4*)(Arow + kk * PITCH + 4 * v);
                    a2[4*v]=pack2(t.x,t.x); a2[4*v+1]=pack2(t.y,t.y); a2[4*v+2]=pack2(t.z,t.z); a2[4*v+3]=pack2(t.w,t.w); }
                ulonglong2 t0 = *(const ulonglong2*)(Wc + kk * W + 4 * g), t1 = *(const ulonglong2*)(Wc + kk * W + 128 + 4 * g);
                uint64_t w2[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
                for (int i = 0; i < PJ; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) fma2(acc[i][j], a2[i], w2[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < PJ; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { float x, y; unpack2(acc[i][j], x, y); res += x + y; }
    }
    out[blockIdx.x * 256 + tid] = res;
}

template <int V>
void run(const char* name, float* out) {
    const int smem = (W * PITCH + KC * W * 2) * 4;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); probe<V><<<148, 256, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double fl = 2.0 * PJ * 8 * KC * (double)iters * 256 * 148;
    printf("%-44s %.2f TFLOP/s (%s)\n", name, fl / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 256 * 4);
    run<0>("V0 scalar FFMA 20x8", out);
    run<1>("V1 f32x2 row pairs, w dup by mov", out);
    run<2>("V2 f32x2 row pairs, w pre-duplicated in smem", out);
    run<3>("V3 f32x2 neuron pairs, a dup by mov", out);
    return 0;
}
