// Fused Taylor-mode jet forward: x -> (value, first, second order channels) through every
// Linear+activation of the MLP, one persistent CTA per SM, activations resident in shared memory.
//
// Replaces AD_engine.autodiff's vmap(jacrev(jacrev(net))) (reference Engines/AD.py:70-90).
//
// Per CTA tile of TP = PG*PT points the jet block A[k][col] (k = neuron, col = pt*J + channel) lives
// in smem (K-major).  Hidden layer:  H = A^T-contract Wt, register tile PJ x 8 per thread (FFMA),
// weight rows streamed from L2 by a producer warp with cp.async.bulk + mbarrier (kKC rows per stage).
// Epilogue: bias, jet activation, stash (t | h_c) to HBM for the reverse sweep, write z back in place.
#pragma once
#include "jet_common.cuh"

namespace tde {

template <class S>
__device__ __forceinline__ void store_row_vec(float* dst, const float (&v)[S::PT][S::J]) {
    // PT*J contiguous floats -> NV float4
    float flat[S::PJ];
#pragma unroll
    for (int pt = 0; pt < S::PT; ++pt)
#pragma unroll
        for (int c = 0; c < S::J; ++c) flat[pt * S::J + c] = v[pt][c];
#pragma unroll
    for (int i = 0; i < S::NV; ++i)
        *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]);
}
// stash layout [layer][tile][jet column][neuron]: a warp's 32 consecutive neurons store one 128 B line per column
template <class S>
__device__ __forceinline__ void store_stash_cols(float* dst, int W, const float (&v)[S::PT][S::J]) {
#pragma unroll
    for (int pt = 0; pt < S::PT; ++pt)
#pragma unroll
        for (int c = 0; c < S::J; ++c) __stcs(dst + (size_t)(pt * S::J + c) * W, v[pt][c]);
}

template <int NF, int SMASK>
__global__ void __launch_bounds__(kThreads, 1) jet_forward_kernel(const __grid_constant__ JetParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J, PT = S::PT, PJ = S::PJ;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int W = p.W, NG = W / kNT, PG = p.PG, TP = PG * PT, pitch = p.pitch;
    const int n_stages = p.n_stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    float* A = reinterpret_cast<float*>(smem_raw + 128);  // [W][pitch]
    float* wbuf = A + (size_t)W * pitch;                   // [n_stages][kKC][W]
    float* wout_s = wbuf + (size_t)n_stages * kKC * W;     // [n_out][W]
    float* xs = wout_s + p.n_out * W;                      // [TP][d_in]
    float* ybuf = xs + TP * p.d_in;                        // [J*n_out][TP]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_chunks = W / kKC;
    const uint32_t chunk_bytes = (uint32_t)(kKC * W * sizeof(float));

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeThreads / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= kComputeThreads / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ===== producer warp: stream hidden-layer weight rows through the stage ring =====
        if (tid == kComputeThreads) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = 1; l < p.L; ++l) {
                    const float* src = p.wt_hidden + (size_t)(l - 1) * W * W;
                    for (int kc = 0; kc < n_chunks; ++kc) {
                        mbar_wait_backoff(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], chunk_bytes);
                        bulk_g2s(wbuf + (size_t)s * kKC * W, src + (size_t)kc * kKC * W, chunk_bytes, &full[s]);
                        if (++s == n_stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
        return;
    }

    // ===== compute warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    const int g = tid % NG;
    int pg = tid / NG;
    const bool active = pg < PG;
    if (!active) pg = PG - 1;  // idle threads shadow the last group (no stores) so barriers stay uniform
    const int col0 = pg * PJ;

    for (int i = tid; i < p.n_out * W; i += kComputeThreads) wout_s[i] = p.w_out[i];

    int cs = 0;        // consumer position in the stage ring
    uint32_t cph = 0;  // and its phase
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t pbase = tile * TP;
        // ---- stage the points of this tile ----
        for (int i = tid; i < TP * p.d_in; i += kComputeThreads) {
            const int64_t gp = pbase + i / p.d_in;
            xs[i] = gp < p.B ? p.x[gp * p.d_in + (i % p.d_in)] : 0.0f;
        }
        compute_sync();

        float* stash_tile = p.stash ? p.stash + (size_t)tile * W * (TP * J) : nullptr;
        const size_t stash_layer_stride = (size_t)p.stash_layer_stride;

        // ---- input layer (K = d_in): jets of an affine map are its columns ----
        {
#pragma unroll 1
            for (int j = 0; j < kNT; ++j) {
                const int n = g + NG * j;
                float wrow[TDE_MAX_IN];
#pragma unroll
                for (int i = 0; i < TDE_MAX_IN; ++i) wrow[i] = i < p.d_in ? __ldg(p.w_in + n * p.d_in + i) : 0.0f;
                const float bn = __ldg(p.b_in + n);
                float wf[NF > 0 ? NF : 1];
#pragma unroll
                for (int a = 0; a < NF; ++a) wf[a] = __ldg(p.w_in + n * p.d_in + p.first_in[a]);
                float zt[PT][J], st[PT][J];
#pragma unroll
                for (int pt = 0; pt < PT; ++pt) {
                    float h[J];
                    float h0 = bn;
#pragma unroll
                    for (int i = 0; i < TDE_MAX_IN; ++i)
                        if (i < p.d_in) h0 = fmaf(wrow[i], xs[(pg * PT + pt) * p.d_in + i], h0);
                    h[0] = h0;
#pragma unroll
                    for (int a = 0; a < NF; ++a) h[1 + a] = wf[a];
#pragma unroll
                    for (int q = 0; q < S::NS; ++q) h[1 + NF + q] = 0.0f;
                    act_forward<S>(layer_act(p.act, p.fourier_half, 0, n), h, zt[pt]);
#pragma unroll
                    for (int c = 0; c < J; ++c) st[pt][c] = h[c];
                }
                if (active) {
                    store_row_vec<S>(A + (size_t)n * pitch + col0, zt);
                    if (stash_tile) store_stash_cols<S>(stash_tile + (size_t)col0 * W + n, W, st);
                }
            }
        }
        compute_sync();

        // ---- hidden layers ----
        for (int l = 1; l < p.L; ++l) {
            AccTile<PJ> acc;
            acc.zero();
            // bias of this layer for the thread's 8 neurons: fetched now, consumed after the GEMM
            float bn[kNT];
#pragma unroll
            for (int j = 0; j < kNT; ++j) bn[j] = __ldg(p.b_hidden + (size_t)(l - 1) * W + g + NG * j);

            for (int kc = 0; kc < n_chunks; ++kc) {
                const int s = cs;
                mbar_wait(&full[s], cph);
                gemm_chunk<PJ>(acc, A + (size_t)(kc * kKC) * pitch + col0, wbuf + (size_t)s * kKC * W, pitch, W, g);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++cs == n_stages) {
                    cs = 0;
                    cph ^= 1;
                }
            }
            compute_sync();  // every warp is done reading A for this layer

            float* stash_l = stash_tile ? stash_tile + (size_t)l * stash_layer_stride : nullptr;
#pragma unroll
            for (int j = 0; j < kNT; ++j) {
                const int n = g + NG * j;
                float zt[PT][J], st[PT][J];
#pragma unroll
                for (int pt = 0; pt < PT; ++pt) {
                    float h[J];
#pragma unroll
                    for (int c = 0; c < J; ++c) h[c] = acc.get(pt * J + c, j);
                    h[0] += bn[j];
                    act_forward<S>(p.act, h, zt[pt]);
#pragma unroll
                    for (int c = 0; c < J; ++c) st[pt][c] = h[c];
                }
                if (active) {
                    store_row_vec<S>(A + (size_t)n * pitch + col0, zt);
                    if (stash_l) store_stash_cols<S>(stash_l + (size_t)col0 * W + n, W, st);
                }
            }
            compute_sync();
        }

        // ---- output layer (N = n_out): one thread per jet column ----
        for (int col = tid; col < TP * J; col += kComputeThreads) {
            float y[TDE_MAX_OUT];
#pragma unroll
            for (int o = 0; o < TDE_MAX_OUT; ++o) y[o] = 0.0f;
            for (int k = 0; k < W; ++k) {
                const float a = A[(size_t)k * pitch + col];
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o)
                    if (o < p.n_out) y[o] = fmaf(a, wout_s[o * W + k], y[o]);
            }
            const int pt = col / J, c = col - pt * J;
#pragma unroll
            for (int o = 0; o < TDE_MAX_OUT; ++o)
                if (o < p.n_out) ybuf[(c * p.n_out + o) * TP + pt] = y[o] + (c == 0 ? __ldg(p.b_out + o) : 0.0f);
        }
        compute_sync();
        for (int i = tid; i < J * p.n_out * TP; i += kComputeThreads) {
            const int r = i / TP, pt = i - r * TP;
            const int64_t gp = pbase + pt;
            if (gp < p.B) p.out[(size_t)r * p.ld_out + gp] = ybuf[i];
        }
        // next tile's xs/A writes are ordered behind this by the compute_sync after the xs load
        compute_sync();
    }
}

}  // namespace tde
