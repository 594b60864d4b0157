// Fused reverse sweep of the jet network: seeds dL/d(out jets) -> dL/d(weights).
//
// Replaces torch autograd's walk of the reverse-over-reverse graph in `loss.backward()`
// (reference call sites Tutorials/Spring.py:54, Lid_Driven_Cavity.py:95).
//
// The adjoint of a J-channel jet network is again a J-channel network:
//   activation adjoint (elementwise, needs the stash t | h_c)        -> Hbar
//   dW_l   += Hbar_l^T  Z_{l-1}        (reduction over the tile's jet columns)
//   Zbar_{l-1} = Hbar_l  W_l            (same register-tile GEMM as the forward, weights streamed)
// Per-CTA partial gradients are kept in a private slab (no atomics, fixed tile->CTA map), and a
// second-stage kernel reduces the slabs in a fixed order => bitwise run-to-run reproducible.
#pragma once
#include "jet_common.cuh"

namespace tde {

// Accumulate into the CTA-private gradient slab.  Each address has exactly one writer thread, so a
// fire-and-forget vector reduction (RED.128: no load latency to wait on) is still order-deterministic.
__device__ __forceinline__ void rmw4(float* dst, float4 v, bool first) {
    if (first) {
        *reinterpret_cast<float4*>(dst) = v;
    } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
}
__device__ __forceinline__ void rmw1(float* dst, float v, bool first) { *dst = first ? v : (*dst + v); }

template <class S>
__device__ __forceinline__ void load_row_vec(const float* src, float (&v)[S::PT][S::J]) {
    float flat[S::PJ];
#pragma unroll
    for (int i = 0; i < S::NV; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(src + 4 * i);
        flat[4 * i] = t.x;
        flat[4 * i + 1] = t.y;
        flat[4 * i + 2] = t.z;
        flat[4 * i + 3] = t.w;
    }
#pragma unroll
    for (int pt = 0; pt < S::PT; ++pt)
#pragma unroll
        for (int c = 0; c < S::J; ++c) v[pt][c] = flat[pt * S::J + c];
}
// stash layout [layer][tile][jet column][neuron]
template <class S>
__device__ __forceinline__ void load_stash_cols(const float* src, int W, float (&v)[S::PT][S::J]) {
#pragma unroll
    for (int pt = 0; pt < S::PT; ++pt)
#pragma unroll
        for (int c = 0; c < S::J; ++c) v[pt][c] = __ldcs(src + (size_t)(pt * S::J + c) * W);
}
template <class S>
__device__ __forceinline__ void store_row_vec_b(float* dst, const float (&v)[S::PT][S::J]) {
    float flat[S::PJ];
#pragma unroll
    for (int pt = 0; pt < S::PT; ++pt)
#pragma unroll
        for (int c = 0; c < S::J; ++c) flat[pt * S::J + c] = v[pt][c];
#pragma unroll
    for (int i = 0; i < S::NV; ++i)
        *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(flat[4 * i], flat[4 * i + 1], flat[4 * i + 2], flat[4 * i + 3]);
}

template <int NF, int SMASK>
__global__ void __launch_bounds__(kThreads, 1) jet_backward_kernel(const __grid_constant__ JetParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J, PT = S::PT, PJ = S::PJ;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int W = p.W, NG = W / kNT, PG = p.PG, TP = PG * PT, pitch = p.pitch;
    const int TPJ = TP * J;
    const int n_stages = p.n_stages;
    const int kblk = p.kblk, rsplit = p.rsplit;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    uint64_t* go = empty + kMaxStages;
    float* A = reinterpret_cast<float*>(smem_raw + 128);  // [W][pitch]   Zbar / Hbar
    float* ubuf = A + (size_t)W * pitch;                   // union: weight stages | Z block | reduction scratch
    size_t ubuf_floats = (size_t)n_stages * kKC * W;
    if ((size_t)kblk * pitch > ubuf_floats) ubuf_floats = (size_t)kblk * pitch;
    if ((size_t)PG * p.n_out * W > ubuf_floats) ubuf_floats = (size_t)PG * p.n_out * W;
    float* wout_s = ubuf + ubuf_floats;       // [n_out][W]
    float* xs = wout_s + p.n_out * W;         // [TP][d_in]
    float* seeds_s = xs + TP * p.d_in;        // [J*n_out][TP]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n_chunks = W / kKC;
    const uint32_t chunk_bytes = (uint32_t)(kKC * W * sizeof(float));

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeThreads / 32);
        }
        mbar_init(go, kComputeThreads / 32);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= kComputeThreads / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ===== producer: stream W_l rows (n-major) for the dX GEMM of every hidden layer =====
        if (tid == kComputeThreads) {
            int s = 0;
            uint32_t ph = 0, go_phase = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = p.L - 1; l >= 1; --l) {
                    mbar_wait_backoff(go, go_phase);  // the dW phase has released the aliased buffer
                    go_phase ^= 1;
                    const float* src = p.wn_hidden + (size_t)(l - 1) * W * W;
                    for (int kc = 0; kc < n_chunks; ++kc) {
                        mbar_wait_backoff(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], chunk_bytes);
                        bulk_g2s(ubuf + (size_t)s * kKC * W, src + (size_t)kc * kKC * W, chunk_bytes, &full[s]);
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
    if (!active) pg = PG - 1;
    const int col0 = pg * PJ;

    // dW-phase thread map: NI n-groups x KI k-groups x rsplit column splits
    const int NI = (W / 8 < 32) ? W / 8 : 32;
    const int KI = kblk / 8;
    const int ni = tid % NI;
    const int ki = (tid / NI) % KI;
    const int rs = tid / (NI * KI);
    const bool dw_active = rs < rsplit;
    const int n_pass = W / (8 * NI);
    const int cols_per_split = TPJ / rsplit;

    for (int i = tid; i < p.n_out * W; i += kComputeThreads) wout_s[i] = p.w_out[i];

    float* gp = p.gpart + (size_t)blockIdx.x * p.gp_stride;
    const size_t stash_layer_stride = (size_t)p.stash_layer_stride;

    int cs = 0;        // consumer position in the stage ring
    uint32_t cph = 0;  // and its phase
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const bool first = (tile == (int64_t)blockIdx.x);
        const int64_t pbase = tile * TP;
        for (int i = tid; i < TP * p.d_in; i += kComputeThreads) {
            const int64_t gpt = pbase + i / p.d_in;
            xs[i] = gpt < p.B ? p.x[gpt * p.d_in + (i % p.d_in)] : 0.0f;
        }
        for (int i = tid; i < J * p.n_out * TP; i += kComputeThreads) {
            const int r = i / TP, pt = i - r * TP;
            const int64_t gpt = pbase + pt;
            seeds_s[i] = gpt < p.B ? p.seeds[(size_t)r * p.ld_seeds + gpt] : 0.0f;
        }
        compute_sync();

        const float* stash_tile = p.stash + (size_t)tile * W * TPJ;

        // ---- output layer adjoint fused with the activation adjoint of layer L-1 ----
        {
            const float* stash_l = stash_tile + (size_t)(p.L - 1) * stash_layer_stride;
            float dwo[TDE_MAX_OUT][kNT];
#pragma unroll
            for (int o = 0; o < TDE_MAX_OUT; ++o)
#pragma unroll
                for (int j = 0; j < kNT; ++j) dwo[o][j] = 0.0f;
#pragma unroll
            for (int j = 0; j < kNT; ++j) {
                const int n = g + NG * j;
                float st[PT][J], hb[PT][J];
                load_stash_cols<S>(stash_l + (size_t)col0 * W + n, W, st);
                float wo[TDE_MAX_OUT];
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) wo[o] = o < p.n_out ? wout_s[o * W + n] : 0.0f;
                float dj[TDE_MAX_OUT];
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) dj[o] = 0.0f;
#pragma unroll
                for (int pt = 0; pt < PT; ++pt) {
                    float zb[J], z[J];
                    float sd[J][TDE_MAX_OUT];
#pragma unroll
                    for (int c = 0; c < J; ++c) {
                        float a = 0.0f;
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o) {
                            sd[c][o] = o < p.n_out ? seeds_s[(c * p.n_out + o) * TP + pg * PT + pt] : 0.0f;
                            a = fmaf(wo[o], sd[c][o], a);
                        }
                        zb[c] = a;
                    }
                    act_backward<S, true>(layer_act(p.act, p.fourier_half, p.L - 1, n), st[pt], zb, hb[pt], z);
#pragma unroll
                    for (int c = 0; c < J; ++c)
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o) dj[o] = fmaf(sd[c][o], z[c], dj[o]);
                }
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) dwo[o][j] = dj[o];
                if (active) store_row_vec_b<S>(A + (size_t)n * pitch + col0, hb);
            }
            // cross point-group reduction of dW_out through the scratch buffer (fixed order)
            if (active) {
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o)
                    if (o < p.n_out) {
#pragma unroll
                        for (int j = 0; j < kNT; ++j) ubuf[(size_t)(pg * p.n_out + o) * W + g + NG * j] = dwo[o][j];
                    }
            }
            compute_sync();
            for (int i = tid; i < p.n_out * W; i += kComputeThreads) {
                float s = 0.0f;
                for (int q = 0; q < PG; ++q) s += ubuf[(size_t)q * p.n_out * W + i];
                rmw1(gp + p.gp_off_w_out + i, s, first);
            }
            if (tid < p.n_out) {
                float s = 0.0f;
                for (int pt = 0; pt < TP; ++pt) s += seeds_s[(size_t)tid * TP + pt];  // channel 0 rows come first
                rmw1(gp + p.gp_off_b_out + tid, s, first);
            }
            compute_sync();
        }

        for (int l = p.L - 1; l >= 1; --l) {
            // ---- db_l: column sums of the value-channel adjoint ----
            for (int n = tid; n < W; n += kComputeThreads) {
                float s = 0.0f;
                for (int pt = 0; pt < TP; ++pt) s += A[(size_t)n * pitch + pt * J];
                rmw1(gp + p.gp_off_b_hidden + (size_t)(l - 1) * W + n, s, first);
            }

            // ---- dW_l[n][k] += sum_col Hbar[n][col] * Z_{l-1}[k][col] ----
            const float* stash_prev = stash_tile + (size_t)(l - 1) * stash_layer_stride;
            float* gw = gp + p.gp_off_hidden + ((size_t)(l - 1) * rsplit + rs) * W * W;
            for (int kb = 0; kb < W / kblk; ++kb) {
                // stage Z_{l-1} rows [kb*kblk, +kblk): recomputed from the stash
                for (int item = tid; item < kblk * PG; item += kComputeThreads) {
                    const int grp = item / kblk, r = item - grp * kblk;   // consecutive threads -> consecutive neurons
                    float st[PT][J], z[PT][J];
                    load_stash_cols<S>(stash_prev + (size_t)(grp * PJ) * W + kb * kblk + r, W, st);
#pragma unroll
                    for (int pt = 0; pt < PT; ++pt) act_recompute<S>(layer_act(p.act, p.fourier_half, l - 1, kb * kblk + r), st[pt], z[pt]);
                    store_row_vec_b<S>(ubuf + (size_t)r * pitch + grp * PJ, z);
                }
                compute_sync();
                for (int np = 0; dw_active && np < n_pass; ++np) {
                    // acc2[j][i] = (partial over even columns, partial over odd columns): operands are natural
                    // 64-bit pairs of the float4 loads, so the whole product runs on FFMA2 with no packing moves
                    uint64_t acc2[8][8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc2[j][i] = 0ull;
                    const float* Hrow = A + (size_t)(np * 8 * NI + ni) * pitch;
                    const float* Zrow = ubuf + (size_t)(8 * ki) * pitch;
                    const int cbeg = rs * cols_per_split, cend = cbeg + cols_per_split;
#pragma unroll 2
                    for (int c4 = cbeg; c4 < cend; c4 += 4) {
                        ulonglong2 hv[8], zv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hv[j] = *reinterpret_cast<const ulonglong2*>(Hrow + (size_t)(NI * j) * pitch + c4);
#pragma unroll
                        for (int i = 0; i < 8; ++i) zv[i] = *reinterpret_cast<const ulonglong2*>(Zrow + (size_t)i * pitch + c4);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                fma2(acc2[j][i], hv[j].x, zv[i].x);
                                fma2(acc2[j][i], hv[j].y, zv[i].y);
                            }
                    }
                    float acc[8][8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float lo, hi;
                            unpack2(acc2[j][i], lo, hi);
                            acc[j][i] = lo + hi;
                        }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int n = np * 8 * NI + ni + NI * j;
                        float* dst = gw + (size_t)n * W + kb * kblk + 8 * ki;
                        rmw4(dst, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]), first);
                        rmw4(dst + 4, make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]), first);
                    }
                }
                fence_proxy_async();  // generic-proxy accesses of the aliased buffer precede the next bulk copy
                compute_sync();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(go);

            // ---- Zbar_{l-1}[col][k] = sum_n Hbar_l[n][col] * W_l[n][k]  (k = g + NG*j) ----
            AccTile<PJ> acc;
            acc.zero();
            for (int kc = 0; kc < n_chunks; ++kc) {
                const int s = cs;
                mbar_wait(&full[s], cph);
                gemm_chunk<PJ>(acc, A + (size_t)(kc * kKC) * pitch + col0, ubuf + (size_t)s * kKC * W, pitch, W, g);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++cs == n_stages) {
                    cs = 0;
                    cph ^= 1;
                }
            }
            compute_sync();  // all reads of Hbar_l done; A can be overwritten

            // ---- activation adjoint of layer l-1 straight from the accumulators ----
            const float* stash_l1 = stash_prev;
#pragma unroll
            for (int j = 0; j < kNT; ++j) {
                const int n = g + NG * j;
                float st[PT][J], hb[PT][J];
                load_stash_cols<S>(stash_l1 + (size_t)col0 * W + n, W, st);
#pragma unroll
                for (int pt = 0; pt < PT; ++pt) {
                    float zb[J], zdummy[J];
#pragma unroll
                    for (int c = 0; c < J; ++c) zb[c] = acc.get(pt * J + c, j);
                    act_backward<S, false>(layer_act(p.act, p.fourier_half, l - 1, n), st[pt], zb, hb[pt], zdummy);
                }
                if (active) store_row_vec_b<S>(A + (size_t)n * pitch + col0, hb);
            }
            compute_sync();
        }

        // ---- input layer: h = W0 x + b0, h_a = W0[:, first_in[a]], second-order channels are zero ----
        for (int n = tid; n < W; n += kComputeThreads) {
            float db = 0.0f;
            float dw[TDE_MAX_IN];
#pragma unroll
            for (int i = 0; i < TDE_MAX_IN; ++i) dw[i] = 0.0f;
            for (int pt = 0; pt < TP; ++pt) {
                const float* hrow = A + (size_t)n * pitch + pt * J;
                const float h0 = hrow[0];
                db += h0;
#pragma unroll
                for (int i = 0; i < TDE_MAX_IN; ++i)
                    if (i < p.d_in) dw[i] = fmaf(h0, xs[pt * p.d_in + i], dw[i]);
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    const float ha = hrow[1 + a];
#pragma unroll
                    for (int i = 0; i < TDE_MAX_IN; ++i)
                        if (i == p.first_in[a]) dw[i] += ha;
                }
            }
            rmw1(gp + p.gp_off_b_in + n, db, first);
#pragma unroll
            for (int i = 0; i < TDE_MAX_IN; ++i)
                if (i < p.d_in) rmw1(gp + (size_t)n * p.d_in + i, dw[i], first);
        }
        compute_sync();
    }
}

}  // namespace tde
