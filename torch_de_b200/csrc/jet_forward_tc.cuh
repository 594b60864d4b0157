// tcgen05 forward for wide layers (W = 128 or 256): hidden-layer GEMMs on the 5th-gen tensor cores with a
// 3xTF32 split (hi*hi + hi*lo + lo*hi, FP32 accumulate in TMEM) to hold FP32 accuracy.
//
//   D[neuron][col] = sum_k W_l[neuron][k] * Z[k][col]        col = pt*J + channel, 16 points per CTA tile
//
// A operand = weights (M = 128 neurons per block, K-major): streamed by a producer warp with cp.async.bulk from a
//             global image that is already in UMMA core-matrix order and already split into hi/lo.
// B operand = activations (N = 16*J jet columns, K-major): resident in shared memory as two planes Z_hi, Z_lo,
//             written by the epilogue of the previous layer in UMMA core-matrix order (8 cols x 4 k per 128 B).
// D         = TMEM, lane = neuron => after tcgen05.ld every epilogue thread owns ALL jet channels of its neuron for
//             every point of the tile, which is exactly what the jet activation needs (no cross-lane exchange).
// Roles: warp 0 weight producer, warp 1 MMA issuer (one elected thread), warps 4.. epilogue (1 thread per neuron).
#pragma once
#include "jet_common.cuh"

namespace tde {

constexpr int kTcPoints = 16;      // points per tile
constexpr int kTcKC = 8;           // k rows per weight chunk (= one tf32 MMA K step)
constexpr int kTcThreads = 384;    // 4 control warps + 8 epilogue warps

struct TcParams {
    const float* x;
    int64_t B;
    int64_t n_tiles;       // 16-point tiles
    int d_in, n_out, L, W, act, MB;   // MB = W / 128
    int n_stages;
    int dbg;               // bit0: skip weight streaming (timing experiment only)
    int n_acc;             // TMEM accumulator sets: 3 = {small terms, main even k-steps, main odd k-steps}, 2, or 1
    int first_in[TDE_MAX_FIRST];
    const float* w_in;     // [W][d_in]
    const float* b_in;     // [W]
    const float* wc;       // [(L-1)][W/8] chunks, each MB*2*2*2048 floats... see pack_tc_weights_kernel
    const float* b_hidden; // [(L-1)][W]
    const float* w_out;    // [n_out][W]
    const float* b_out;
    float* out;
    int64_t ld_out;
    float* stash;          // [L][n_tiles_bw][TPbw*J columns][W]
    int64_t stash_layer_stride;
    int tiles_per_bw_tile; // TPbw / 16
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // K-major, SWIZZLE_NONE canonical layout; version = 1 (sm_100)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int NF, int SMASK>
__global__ void __launch_bounds__(kTcThreads, 1) jet_forward_tc_kernel(const __grid_constant__ TcParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J;
    constexpr int N = kTcPoints * J;             // jet columns per tile (multiple of 16)
    constexpr uint32_t LBZ = (N / 8) * 128 + 16;  // byte stride between 4-k groups of the activation planes (+16: bank skew)
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int W = p.W, MB = p.MB;
    const int n_stages = p.n_stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    uint64_t* act_ready = empty + kMaxStages;
    uint64_t* acc_ready = act_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
    unsigned char* zhi = smem_raw + 256;
    const uint32_t z_plane = (uint32_t)(W / 4) * LBZ;
    unsigned char* zlo = zhi + z_plane;
    const uint32_t chunk_bytes = (uint32_t)MB * 2u * 2u * 2048u;   // [hi|lo][mb][kgroup 2][16 row groups][8][4] floats*4
    unsigned char* wst = zlo + z_plane;                              // [n_stages][chunk_bytes]
    float* wout_s = reinterpret_cast<float*>(wst + (size_t)n_stages * chunk_bytes);   // [n_out][W]
    float* xs = wout_s + p.n_out * W;            // [16][d_in]
    float* ybuf = xs + kTcPoints * p.d_in;       // [J*n_out][16]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_epi_warps = 4 * MB;
    const int n_epi = 128 * MB;
    const int n_ks = W / kTcKC;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(act_ready, n_epi_warps);
        mbar_init(acc_ready, 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == 0) {
            // ===== producer: weight chunks (already hi/lo split, UMMA core-matrix order) =====
            int s = 0;
            uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = 1; l < p.L; ++l) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wc) + (size_t)(l - 1) * n_ks * chunk_bytes;
                    for (int ks = 0; ks < n_ks; ++ks) {
                        if (p.dbg & 1) continue;
                        mbar_wait(&empty[s], ph ^ 1);   // spin: a chunk is consumed in ~150 ns, the refill must be prompt
                        mbar_expect_tx(&full[s], chunk_bytes);
                        bulk_g2s(wst + (size_t)s * chunk_bytes, src + (size_t)ks * chunk_bytes, chunk_bytes, &full[s]);
                        if (++s == n_stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (tid == 32) {
            // ===== MMA issuer =====
            // instruction descriptor: D=F32, A=B=TF32, both K-major, N, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int s = 0;
            uint32_t ph = 0, act_ph = 0;
            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                for (int l = 1; l < p.L; ++l) {
                    mbar_wait(act_ready, act_ph);   // Z planes of layer l-1 are in smem, TMEM accumulators are free
                    act_ph ^= 1;
                    tc_fence_after();
                    for (int ks = 0; ks < n_ks; ++ks) {
                        if (!(p.dbg & 1)) mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t wbase = smem_u32(wst + (size_t)s * chunk_bytes);
                        const uint64_t b_hi = umma_desc(smem_u32(zhi) + ks * 2 * LBZ, LBZ, 128);
                        const uint64_t b_lo = umma_desc(smem_u32(zlo) + ks * 2 * LBZ, LBZ, 128);
                        for (int mb = 0; mb < MB && !(p.dbg & 2); ++mb) {
                            const uint64_t a_hi = umma_desc(wbase + (uint32_t)(0 * MB + mb) * 4096u, 2048, 128);
                            const uint64_t a_lo = umma_desc(wbase + (uint32_t)(1 * MB + mb) * 4096u, 2048, 128);
                            // The tensor core rounds its FP32 accumulator toward zero after every MMA: that bias, not the
                            // TF32 split, limits accuracy.  Keep the 2^-11-sized cross terms in their own accumulator
                            // and alternate the main term over two accumulators; the epilogue adds them in FP32 (RN).
                            const uint32_t d_small = tmem_base + (uint32_t)(mb * N);
                            uint32_t d_main = d_small, first_main = ks == 0, first_small = ks == 0;
                            if (p.n_acc == 3) {
                                d_main = tmem_base + (uint32_t)((MB * (1 + (ks & 1)) + mb) * N);
                                first_main = ks < 2;
                            } else if (p.n_acc == 2) {
                                d_main = tmem_base + (uint32_t)((MB + mb) * N);
                            }
                            umma_tf32(d_small, a_lo, b_hi, idesc, !first_small);
                            umma_tf32(d_small, a_hi, b_lo, idesc, 1);
                            umma_tf32(d_main, a_hi, b_hi, idesc, p.n_acc == 1 ? 1u : !first_main);
                        }
                        umma_commit(&empty[s]);      // frees the weight stage when these MMAs have read it
                        if (++s == n_stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                    umma_commit(acc_ready);          // accumulators of layer l complete
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
        const int e = warp - 4;                      // epilogue warp 0..7
        if (e < n_epi_warps) {
            const int quarter = warp & 3;            // TMEM lane quarter this warp may access
            const int mb = e >> 2;
            const int n = mb * 128 + quarter * 32 + lane;   // the neuron this thread owns
            const int et = e * 32 + lane;            // 0..n_epi-1 dense index for strided loops
            auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory"); };

            for (int i = et; i < p.n_out * W; i += n_epi) wout_s[i] = p.w_out[i];
            // byte offset of element (col, k = n) inside a Z plane
            const uint32_t zrow = (uint32_t)(n >> 2) * LBZ + (uint32_t)(n & 3) * 4u;
            uint32_t acc_ph = 0;

            for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int64_t pbase = tile * kTcPoints;
                for (int i = et; i < kTcPoints * p.d_in; i += n_epi) {
                    const int64_t gp = pbase + i / p.d_in;
                    xs[i] = gp < p.B ? p.x[gp * p.d_in + (i % p.d_in)] : 0.0f;
                }
                epi_sync();
                float* stash_row = nullptr;
                if (p.stash) {
                    const int64_t bw_tile = tile / p.tiles_per_bw_tile;
                    const int sub = (int)(tile - bw_tile * p.tiles_per_bw_tile);
                    // [layer][backward tile][jet column][neuron]: this thread's neuron, first column of its 16-point sub-tile
                    stash_row = p.stash + ((size_t)bw_tile * p.tiles_per_bw_tile * N + (size_t)sub * N) * W + n;
                }

                for (int l = 0; l < p.L; ++l) {
                    float hbuf[N];
                    if (l == 0) {
                        float wrow[TDE_MAX_IN];
#pragma unroll
                        for (int i = 0; i < TDE_MAX_IN; ++i) wrow[i] = i < p.d_in ? __ldg(p.w_in + n * p.d_in + i) : 0.0f;
                        const float bn = __ldg(p.b_in + n);
                        float wf[NF > 0 ? NF : 1];
#pragma unroll
                        for (int a = 0; a < NF; ++a) wf[a] = __ldg(p.w_in + n * p.d_in + p.first_in[a]);
#pragma unroll
                        for (int pt = 0; pt < kTcPoints; ++pt) {
                            float h0 = bn;
#pragma unroll
                            for (int i = 0; i < TDE_MAX_IN; ++i)
                                if (i < p.d_in) h0 = fmaf(wrow[i], xs[pt * p.d_in + i], h0);
                            hbuf[pt * J] = h0;
#pragma unroll
                            for (int a = 0; a < NF; ++a) hbuf[pt * J + 1 + a] = wf[a];
#pragma unroll
                            for (int q = 0; q < S::NS; ++q) hbuf[pt * J + 1 + NF + q] = 0.0f;
                        }
                    } else {
                        mbar_wait(acc_ready, acc_ph);
                        acc_ph ^= 1;
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * N);
#pragma unroll
                        for (int c0 = 0; c0 < N; c0 += 16) tmem_ld16(taddr + c0, hbuf + c0);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        for (int a = 1; a < p.n_acc; ++a) {
#pragma unroll
                            for (int c0 = 0; c0 < N; c0 += 16) {
                                float part[16];
                                tmem_ld16(taddr + (uint32_t)(a * MB * N) + c0, part);
                                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                for (int i = 0; i < 16; ++i) hbuf[c0 + i] += part[i];
                            }
                        }
                        const float bn = __ldg(p.b_hidden + (size_t)(l - 1) * W + n);
#pragma unroll
                        for (int pt = 0; pt < kTcPoints; ++pt) hbuf[pt * J] += bn;
                    }
                    // jet activation, stash, split and store the next B operand
                    float* srow = stash_row ? stash_row + (size_t)l * p.stash_layer_stride : nullptr;
#pragma unroll
                    for (int pt = 0; pt < kTcPoints; ++pt) {
                        float h[J], z[J];
#pragma unroll
                        for (int c = 0; c < J; ++c) h[c] = hbuf[pt * J + c];
                        act_forward<S>(p.act, h, z);
#pragma unroll
                        for (int c = 0; c < J; ++c) {
                            hbuf[pt * J + c] = h[c];     // (t | h_c) for the stash
                            const int col = pt * J + c;
                            const uint32_t off = zrow + (uint32_t)(col >> 3) * 128u + (uint32_t)(col & 7) * 16u;
                            float hi, lo;
                            split_tf32(z[c], hi, lo);
                            *reinterpret_cast<float*>(zhi + off) = hi;
                            *reinterpret_cast<float*>(zlo + off) = lo;
                        }
                    }
                    if (srow) {
#pragma unroll
                        for (int i = 0; i < N; ++i) __stcs(srow + (size_t)i * W, hbuf[i]);   // one 128 B line per warp store
                    }
                    if (l < p.L - 1) {
                        // publish: smem writes -> async proxy, TMEM reads retired, then one arrive per warp
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(act_ready);
                    }
                }
                epi_sync();   // every neuron's last-layer z is in the planes

                // ---- output layer: one thread per jet column, z = hi + lo ----
                for (int col = et; col < N; col += n_epi) {
                    float y[TDE_MAX_OUT];
#pragma unroll
                    for (int o = 0; o < TDE_MAX_OUT; ++o) y[o] = 0.0f;
                    const uint32_t cbase = (uint32_t)(col >> 3) * 128u + (uint32_t)(col & 7) * 16u;
                    for (int kg = 0; kg < W / 4; ++kg) {
                        const float4 a = *reinterpret_cast<const float4*>(zhi + (uint32_t)kg * LBZ + cbase);
                        const float4 b = *reinterpret_cast<const float4*>(zlo + (uint32_t)kg * LBZ + cbase);
                        const float zz[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o)
                            if (o < p.n_out) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) y[o] = fmaf(zz[i], wout_s[o * W + kg * 4 + i], y[o]);
                            }
                    }
                    const int pt = col / J, c = col - pt * J;
#pragma unroll
                    for (int o = 0; o < TDE_MAX_OUT; ++o)
                        if (o < p.n_out) ybuf[(c * p.n_out + o) * kTcPoints + pt] = y[o] + (c == 0 ? __ldg(p.b_out + o) : 0.0f);
                }
                epi_sync();
                for (int i = et; i < J * p.n_out * kTcPoints; i += n_epi) {
                    const int r = i / kTcPoints, pt = i - r * kTcPoints;
                    const int64_t gp = pbase + pt;
                    if (gp < p.B) p.out[(size_t)r * p.ld_out + gp] = ybuf[i];
                }
                epi_sync();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace tde
