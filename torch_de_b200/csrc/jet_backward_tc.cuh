// tcgen05 reverse sweep for wide layers (padded width W = 128, 256 or 512), two kernels:
//
//  jet_backward_tc_kernel : seeds -> Hbar jets of every layer.  The dX GEMMs  Zbar_{l-1}[k][col] = sum_n W_l[n][k] Hbar_l[n][col]
//      run on the tensor cores exactly like the forward (A = W_l^T image streamed + multicast, B = Hbar planes resident in
//      smem, D in TMEM with lane = neuron, 3xTF32, same issue loop).  An epilogue thread owns one neuron for 4 (W = 512: 8)
//      points of the tile, so the activation adjoint, the bias gradients, dW_out and dW_in need no cross-thread reduction (the
//      threads of a neuron add into separate copies of the small slab sections).  It also writes Hbar_l to HBM
//      ([layer][tile][column][neuron], coalesced) for the weight-gradient kernel.
//
//  jet_dw_tc_kernel       : dW_l[n][k] = sum_{points, channels} Hbar_l[n][col] * Z_{l-1}[k][col]  as a split-K tensor-core GEMM
//      (M = n, N = k = 256 TMEM columns per neuron block, K = jet columns; W = 512 as 2 x 2 quadrants of 256 x 256 spread over
//      the CTAs).  A producer thread streams Hbar_l and the stash
//      of layer l-1 with cp.async.bulk into raw stages; splitter threads turn the stash into Z on the fly, split both
//      3xTF32 and lay them out as K-major UMMA core matrices; the accumulator is flushed to the CTA's private gradient
//      slab every few tiles (FP32 RED, single writer) so the tensor core's round-toward-zero accumulation never runs long
//      enough to bias the gradient.
#pragma once
#include <type_traits>
#include "jet_forward_tc.cuh"

namespace tde {

struct TcBwdParams {
    const float* x;
    int64_t B;
    int64_t n_tiles;            // 16-point tiles
    int d_in, n_out, L, W, act, MB;
    int n_stages, n_acc, fourier_half;
    int pair;                   // W = 256: use the CTA-pair kernel
    int dbg;                    // bit 4: weight copies without the L2 evict-last hint (A/B measurement)
    int stage_dbuf;             // shared memory has room for a second (x | seeds) staging buffer: the next tile's is fetched early
    int first_in[TDE_MAX_FIRST];
    const float* wct;           // transposed-weight chunk images [(L-1)][W/8][chunk]
    const float* w_out;         // [n_out][W] packed
    const float* seeds;
    int64_t ld_seeds;
    const float* stash;         // forward stash (t | h_c)   [L][bw tiles][cols][W]
    float* hbar;                // Hbar_l                    same layout (layers 1..L-1 used)
    int64_t stash_layer_stride;
    int tiles_per_bw_tile;
    float* gpart;               // per-CTA slabs
    int64_t gp_stride, gp_off_b_in, gp_off_b_hidden, gp_off_w_out, gp_off_b_out;
    int64_t gp_dup_lo[4], gp_dup_hi[4];   // slab offsets of copy k of (w_in | b_in) and (b_hidden | w_out); [0] = 0
    long long* dbg_buf;         // optional cycle stamps
};

__device__ __forceinline__ void slab_add(float* dst, float v, bool first) {
    if (first) {
        *dst = v;
    } else {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
    }
}

template <int NF, int SMASK, int W, int PTS = tc_points(W, JetSpecT<NF, SMASK>::J)>
__global__ void __launch_bounds__(tc_threads(W, PTS), 1) jet_backward_tc_kernel(const __grid_constant__ TcBwdParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J;
    constexpr int N = PTS * J;
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;

    constexpr int MB = W / 128;
    constexpr int kAccSets = tc_acc_sets(MB, N);
    constexpr int KSC = tc_ksc(W, N);
    constexpr bool kZeroAcc = MB >= 4;           // W = 512: every MMA accumulates, the epilogue zeroes what it has read
    const int n_stages = p.n_stages;
    using P = TcPlanes<N>;
    unsigned char* zpl = smem_raw;                          // Hbar planes: [k-atom][hi | lo part][column group][8 x 128 B]
    constexpr uint32_t z_planes = (uint32_t)(W / 32) * P::kAtom;
    constexpr uint32_t chunk_bytes = (uint32_t)KSC * 8192u;
    unsigned char* wst = zpl + z_planes;
    // x and the seeds of a tile are staged twice over: the NEXT tile's are fetched while this tile's layers run
    float* const xs_buf = reinterpret_cast<float*>(wst + (size_t)n_stages * chunk_bytes);       // 2 x { x [PTS][d_in] | seeds [PTS points][J channels][4 outputs (padded)] }
    const int stage_floats = PTS * p.d_in + J * TDE_MAX_OUT * PTS;
    static_assert(TDE_MAX_OUT == 4, "seed rows are staged as float4 per channel");
    // barriers at the tail (same carve-up as the forward kernel: [planes][stages][2 x (x | seeds)][4 floats][barriers])
    uint64_t* full = reinterpret_cast<uint64_t*>(xs_buf + (p.stage_dbuf ? 2 : 1) * stage_floats + 4);
    uint64_t* empty = full + kTcMaxStages;
    uint64_t* act_ready = empty + kTcMaxStages;
    uint64_t* acc_ready = act_ready + kTcMaxBlocks;     // [neuron block]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + kTcMaxBlocks);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool kTurn = tc_turn(W);           // every epilogue warp serves the neuron blocks in turn (see tc_turn)
    constexpr int PPV = tc_visit_points(W, PTS);      // points per thread and block visit
    constexpr int n_epi_warps = tc_epi_warps(W, PTS);
    constexpr int n_epi = 32 * n_epi_warps;
    constexpr int kArrivals = kTurn ? n_epi_warps : 4 * (PTS / PPV);     // epilogue warps that publish a neuron block
    constexpr int n_ks = (int)((2u * W * W * 4u) / chunk_bytes);      // weight chunks per layer

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTcCluster);
        }
        for (int b = 0; b < MB; ++b) {
            mbar_init(&act_ready[b], kArrivals);     // the epilogue warps that write neuron block b's adjoints
            mbar_init(&acc_ready[b], 1);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = cluster_ctarank();
    const int64_t n_pairs = (p.n_tiles + kTcCluster - 1) / kTcCluster;
    const int64_t first_pair = blockIdx.x / kTcCluster, pair_step = gridDim.x / kTcCluster;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == 0) {
            // ===== producer: W_l^T chunk images, layers L-1 .. 1 =====
            int s = 0;
            uint32_t ph = 0;
            const uint32_t part = chunk_bytes / kTcCluster;
            const uint64_t pol = l2_policy_evict_last();
            const bool hint = !(p.dbg & 16);
            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                for (int l = p.L - 1; l >= 1; --l) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wct) + (size_t)(l - 1) * n_ks * chunk_bytes;
                    for (int ks = 0; ks < n_ks; ++ks) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_expect_tx(&full[s], chunk_bytes);
                        if (hint)
                            bulk_g2s_mc_hint(wst + (size_t)s * chunk_bytes + crank * part, src + (size_t)ks * chunk_bytes + crank * part, part,
                                             &full[s], (uint16_t)((1u << kTcCluster) - 1), pol);
                        else
                            bulk_g2s_mc(wst + (size_t)s * chunk_bytes + crank * part, src + (size_t)ks * chunk_bytes + crank * part, part,
                                        &full[s], (uint16_t)((1u << kTcCluster) - 1));
                        if (++s == n_stages) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer (warp converged, one elected lane) =====
            const bool leader = elect_one();
            const uint64_t b0 = umma_desc_sw128(smem_u32(zpl), 1024);
            const uint64_t a0 = umma_desc(smem_u32(wst), 2048, 128);
            const uint32_t stage_step = chunk_bytes >> 4;
            int s = 0;
            uint32_t ph = 0, act_ph = 0;
            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                for (int l = p.L - 1; l >= 1; --l) {
                    if constexpr (MB >= 4)
                        tc_issue_layer_wide<MB, N, kAccSets, KSC>(leader, tmem_base, a0, b0, stage_step, full, empty, act_ready, acc_ready,
                                                                  n_stages, s, ph, act_ph);
                    else
                        tc_issue_layer<MB, N, kAccSets>(leader, tmem_base, a0, b0, stage_step, full, empty, act_ready, acc_ready,
                                                        n_stages, s, ph, act_ph);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(tc_epi_regs(tc_threads(W, PTS))));
        const int e = warp - 4;
        if (e < n_epi_warps) {
            const int quarter = warp & 3;
            const int own_mb = kTurn ? 0 : (e >> 2) % MB;     // neuron block (turn mode: the first one visited)
            const int pg = kTurn ? (e >> 2) : (e >> 2) / MB;  // point group: this thread handles points PPV*pg .. PPV*pg+PPV-1 of every tile
            constexpr int n_visits = kTurn ? MB : 1;          // neuron blocks a thread serves per layer
            if constexpr (kZeroAcc) {                // all accumulator columns of my lane start at zero (every MMA accumulates)
                tmem_zero_cols<kAccSets * N>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(own_mb * kAccSets * N));
                tmem_wait_st();
            }
            const int n0 = own_mb * 128 + quarter * 32 + lane;   // the neuron this thread owns in every layer (turn mode: + 128 per visit)
            const int et = e * 32 + lane;
            auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory"); };

            const uint32_t hoff = P::template hi_off<kAccSets>((uint32_t)n0 >> 5);
            const uint32_t zrow = (uint32_t)(n0 >> 5) * P::kAtom + hoff + (uint32_t)(n0 & 3) * 4u;
            const uint32_t lo_delta = P::kHalf - 2u * hoff;
            const uint32_t kchunk = (uint32_t)(n0 >> 2) & 7u;
            // shared address of (column = first column of my point group + j, k = n0) in the hi part (see the forward kernel)
            uint32_t zaddr0[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t cb = (uint32_t)(pg * PPV * J + j);
                zaddr0[j] = smem_u32(zpl) + zrow + (cb >> 3) * 1024u + (cb & 7u) * 128u + ((kchunk ^ (cb & 7u)) << 4);
            }
            // the threads of a neuron (one per point group) add into separate copies of the small gradient sections
            float* gp = p.gpart + (size_t)blockIdx.x * p.gp_stride;
            float* gp_lo = gp + p.gp_dup_lo[pg];      // w_in | b_in
            float* gp_hi = gp + p.gp_dup_hi[pg];      // b_hidden | w_out
            uint32_t acc_ph = 0;

            constexpr int GC = kTcGroup * J;
            constexpr int NG2 = PPV / kTcGroup;       // 4-point steps per visit
            // global -> shared staging of one tile's points and seeds (a padding tile gets zeros)
            auto stage_tile = [&](const int64_t tile, float* xs_b) {
                const int64_t pbase = tile * PTS;
                float* seeds_b = xs_b + PTS * p.d_in;
                for (int i = et; i < PTS * p.d_in; i += n_epi) {
                    const int64_t gpt = pbase + i / p.d_in;
                    xs_b[i] = gpt < p.B ? p.x[gpt * p.d_in + (i % p.d_in)] : 0.0f;
                }
                for (int i = et; i < J * TDE_MAX_OUT * PTS; i += n_epi) {       // coalesced over the points of a seed row
                    const int r = i / PTS, pt = i - r * PTS;          // r = channel * 4 + output
                    const int c = r >> 2, o = r & 3;
                    const int64_t gpt = pbase + pt;
                    seeds_b[(pt * J + c) * 4 + o] = (o < p.n_out && gpt < p.B) ? p.seeds[(size_t)(c * p.n_out + o) * p.ld_seeds + gpt] : 0.0f;
                }
            };
            auto row_of = [&](const int64_t tile) { return ((size_t)tile * N + (size_t)pg * (PPV * J)) * W + n0; };   // [layer][tile][jet column][neuron]
            // all of this thread's stash values of a layer (every visit) are requested well before they are used, so their
            // latency hides under the MMAs (a later prefetch would share a scoreboard with younger loads and stall).  Streaming
            // loads (evict-first in L1 and L2: the weight images must stay L2-resident next to 40 GB of stash / Hbar traffic).
            float stv[n_visits][NG2][GC];
            auto load_stash = [&](const int l, const size_t row0) {
                const float* srow = p.stash + (size_t)l * p.stash_layer_stride + row0;
                asm volatile("" : "+l"(srow));
#pragma unroll
                for (int v = 0; v < n_visits; ++v)
#pragma unroll
                    for (int g = 0; g < NG2; ++g)
#pragma unroll
                        for (int i = 0; i < GC; ++i) stv[v][g][i] = __ldcs(srow + (size_t)(g * GC + i) * W + v * 128);
            };
            // my neurons' rows of w_out (read through L1; shared memory has no room for it next to a fourth weight stage)
            float wo_v[n_visits][TDE_MAX_OUT];
#pragma unroll
            for (int v = 0; v < n_visits; ++v)
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) wo_v[v][o] = o < p.n_out ? __ldg(p.w_out + o * W + n0 + v * 128) : 0.0f;
            int cur = 0;
            if (first_pair < n_pairs) {
                load_stash(p.L - 1, row_of(first_pair * kTcCluster + crank));
                stage_tile(first_pair * kTcCluster + crank, xs_buf);
            }
            epi_sync();

            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                const bool first = pair == first_pair;
                // the second tile of the last pair may be padding (>= n_tiles): x = 0 and seeds = 0 there and the streams have
                // rows for it, so it runs the same code and contributes exact zeros
                const int64_t tile = pair * kTcCluster + crank;
                const bool has_next = pair + pair_step < n_pairs;
                const int64_t next_tile = (pair + pair_step) * kTcCluster + crank;
                const bool tstamp = p.dbg_buf && blockIdx.x == 0 && et == 0 && pair == first_pair + 3 * pair_step;
                if (tstamp) p.dbg_buf[40] = clock64();      // tile start
                const float* xs = xs_buf + cur * stage_floats;
                const float* seeds_s = xs + PTS * p.d_in;
                if (tstamp) p.dbg_buf[41] = clock64();      // inputs staged (by the previous iteration)
                const size_t row0 = row_of(tile);

                // Epilogue as 2 x (2 steps of 4 points): 8 points are unrolled so every shared/global offset is an immediate (a
                // 16-point unroll is >200 KB of code and instruction-cache bound).  The stash values of the NEXT 4-point step
                // are requested before the current step is processed; TMEM is read where it is used (12-cycle latency).
                // Base addresses pass through empty asm statements: otherwise the compiler, short of registers, re-derives
                // every one of the 240 addresses of a layer from the kernel parameters (10 instructions per store).
                // The body is instantiated three times (last hidden layer: adjoint seeded from the output layer; middle layers;
                // input layer) so the hot middle-layer code carries none of the special-case registers or branches.
                auto layer_body = [&](auto kind_c, const int l) {
                    constexpr int KIND = decltype(kind_c)::value;     // 2 last, 1 middle, 0 input layer
                    constexpr bool last = KIND == 2;
                    const size_t loff = (size_t)l * p.stash_layer_stride + row0;
                    if (!last) load_stash(l, row0);      // (the last hidden layer's values were requested before the tile began)
                    static_for<n_visits>([&](auto v_c) {
                        constexpr int v = decltype(v_c)::value;
                        const int mb = own_mb + v, n = n0 + v * 128;
                        float* hrow_h = p.hbar + loff + v * 128;
                        uint32_t zaddr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) zaddr[j] = zaddr0[j] + (uint32_t)v * 4u * P::kAtom;
                        const bool stamp = p.dbg_buf && blockIdx.x == 0 && (n & 127) == 0 && pg == 0 && pair == first_pair + 3 * pair_step && ((l >= 3 && l <= 5) || KIND != 1);
                        // steady-state tile, two neuron blocks; slots 50.. / 58..: the last hidden layer (seeded from the output layer) / the input layer
                        long long* const dst = p.dbg_buf + (KIND == 1 ? 16 + ((5 - l) * 2 + (mb & 1)) * 4 : (KIND == 2 ? 50 : 58) + (mb & 1) * 4);
                        if (stamp) dst[0] = clock64();
                        float wo[TDE_MAX_OUT];
                        if (last) {
#pragma unroll
                            for (int o = 0; o < TDE_MAX_OUT; ++o) wo[o] = wo_v[v][o];
                        }
                        uint32_t taddr = 0;
                        if (!last) {
                            mbar_wait_park(&acc_ready[mb], acc_ph);
                            tc_fence_after();
                            taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * kAccSets * N + pg * PPV * J);
                        }
                        if (stamp) dst[1] = clock64();
                        float db = 0.0f;
                        float dwo[TDE_MAX_OUT], dw_in[TDE_MAX_IN], dh_first[NF > 0 ? NF : 1];
#pragma unroll
                        for (int a = 0; a < (NF > 0 ? NF : 1); ++a) dh_first[a] = 0.0f;
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o) dwo[o] = 0.0f;
#pragma unroll
                        for (int i = 0; i < TDE_MAX_IN; ++i) dw_in[i] = 0.0f;
                        const float* xs_h = xs + pg * PPV * p.d_in;
                        const float4* seeds_h = reinterpret_cast<const float4*>(seeds_s) + pg * PPV * J;     // [point][channel]
                        {
                            asm volatile("" : "+l"(hrow_h));
#pragma unroll
                            for (int j = 0; j < 8; ++j) asm volatile("" : "+r"(zaddr[j]));
#pragma unroll
                            for (int g2 = 0; g2 < NG2; ++g2) {
                                float zbg[GC];
                                const float (&stg)[GC] = stv[v][g2];
                                if (!last) {
                                    // one accumulator set at a time: at most 2 x GC registers in flight next to the stash values
                                    tmem_ld_cols<GC>(taddr + (uint32_t)(g2 * GC), zbg);
                                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                    for (int a = 1; a < kAccSets; ++a) {
                                        float cur[GC];
                                        tmem_ld_cols<GC>(taddr + (uint32_t)(a * N + g2 * GC), cur);
                                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                                        for (int i = 0; i < GC; ++i) zbg[i] += cur[i];
                                    }
                                    if constexpr (kZeroAcc) {
#pragma unroll
                                        for (int a = 0; a < kAccSets; ++a) tmem_zero_cols<GC>(taddr + (uint32_t)(a * N + g2 * GC));
                                    }
                                }
#pragma unroll
                                for (int q = 0; q < kTcGroup; ++q) {
                                    const int pt8 = g2 * kTcGroup + q;      // compile-time 0..PPV-1
                                    float st[J], zb[J], hb[J], z[J];
#pragma unroll
                                    for (int c = 0; c < J; ++c) st[c] = stg[q * J + c];
                                    float4 sd[last ? J : 1];       // seeds of this point: [channel] x 4 outputs (0 beyond n_out)
                                    if (last) {
#pragma unroll
                                        for (int c = 0; c < J; ++c) {
                                            sd[c] = seeds_h[pt8 * J + c];
                                            zb[c] = fmaf(wo[0], sd[c].x, fmaf(wo[1], sd[c].y, fmaf(wo[2], sd[c].z, wo[3] * sd[c].w)));
                                        }
                                    } else {
#pragma unroll
                                        for (int c = 0; c < J; ++c) zb[c] = zbg[q * J + c];
                                    }
                                    act_backward<S, last>(layer_act(p.act, p.fourier_half, l, n), st, zb, hb, z);     // Z_l was written by the forward pass
                                    db += hb[0];
                                    if (last) {
#pragma unroll
                                        for (int c = 0; c < J; ++c) {
                                            dwo[0] = fmaf(sd[c].x, z[c], dwo[0]);
                                            dwo[1] = fmaf(sd[c].y, z[c], dwo[1]);
                                            dwo[2] = fmaf(sd[c].z, z[c], dwo[2]);
                                            dwo[3] = fmaf(sd[c].w, z[c], dwo[3]);
                                        }
                                    }
                                    if constexpr (KIND >= 1) {
#pragma unroll
                                        for (int c = 0; c < J; ++c) {
                                            const int col8 = pt8 * J + c;
                                            __stcs(hrow_h + (size_t)col8 * W, hb[c]);
                                            float hi, lo;
                                            split_tf32(hb[c], hi, lo);
                                            const uint32_t a = zaddr[col8 & 7] + (uint32_t)(col8 >> 3) * 1024u;
                                            sts_f32(a, hi);
                                            sts_f32(a + lo_delta, lo);
                                        }
                                    } else {
                                        // input layer: h = W0 x + b0, h_a = W0[:, first_in[a]]
                                        const float h0 = hb[0];
#pragma unroll
                                        for (int i = 0; i < TDE_MAX_IN; ++i)
                                            if (i < p.d_in) dw_in[i] = fmaf(h0, xs_h[pt8 * p.d_in + i], dw_in[i]);
#pragma unroll
                                        for (int a = 0; a < NF; ++a) dh_first[a] += hb[1 + a];     // d/dW0[n][first_in[a]], added once below
                                    }
                                }
                            }
                        }
                        if (stamp) dst[2] = clock64();
                        if constexpr (kZeroAcc) {
                            if (!last) tmem_wait_st();
                        }
                        if constexpr (KIND >= 1) {
                            fence_proxy_async();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&act_ready[mb]);
                            if (stamp) dst[3] = clock64();
                            slab_add(gp_hi + p.gp_off_b_hidden + (size_t)(l - 1) * W + n, db, first);
                        } else {
                            slab_add(gp_lo + p.gp_off_b_in + n, db, first);
#pragma unroll
                            for (int a = 0; a < NF; ++a)
#pragma unroll
                                for (int i = 0; i < TDE_MAX_IN; ++i)
                                    if (i == p.first_in[a]) dw_in[i] += dh_first[a];
#pragma unroll
                            for (int i = 0; i < TDE_MAX_IN; ++i)
                                if (i < p.d_in) slab_add(gp_lo + (size_t)n * p.d_in + i, dw_in[i], first);
                            if (stamp) dst[3] = clock64();
                        }
                        if (last) {
#pragma unroll
                            for (int o = 0; o < TDE_MAX_OUT; ++o)
                                if (o < p.n_out) slab_add(gp_hi + p.gp_off_w_out + (size_t)o * W + n, dwo[o], first);
                        }
                    });
                    if (!last) acc_ph ^= 1;
                };
                layer_body(std::integral_constant<int, 2>{}, p.L - 1);
                if (tstamp) p.dbg_buf[45] = clock64();      // last hidden layer's adjoint (from the seeds) published
                // fetch the next tile's points and seeds now: the MMAs of the first layer have just started, this warp would
                // otherwise wait for them
                if (has_next && p.stage_dbuf) stage_tile(next_tile, xs_buf + (cur ^ 1) * stage_floats);
                for (int l = p.L - 2; l >= 1; --l) {
                    layer_body(std::integral_constant<int, 1>{}, l);
                    if (tstamp && l <= 2) p.dbg_buf[45 + l] = clock64();     // [46]: layer 1 done, [47]: layer 2 done
                }
                layer_body(std::integral_constant<int, 0>{}, 0);
                if (tstamp) p.dbg_buf[48] = clock64();      // input layer done
                if (et < p.n_out) {
                    float s = 0.0f;
                    for (int pt = 0; pt < PTS; ++pt) s += seeds_s[(pt * J) * 4 + et];   // channel 0, output et
                    slab_add(gp + p.gp_off_b_out + et, s, first);
                }
                if (has_next) load_stash(p.L - 1, row_of(next_tile));     // in flight across the barrier
                epi_sync();   // the next tile's staging is visible to every warp; this tile's buffer may be refilled
                if (p.stage_dbuf) {
                    cur ^= 1;
                } else if (has_next) {      // one buffer only: refill it now
                    stage_tile(next_tile, xs_buf);
                    epi_sync();
                }
                if (tstamp) p.dbg_buf[44] = clock64();      // tile done
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// dW_l = Hbar_l^T-contract Z_{l-1} on the tensor cores
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDwThreads = 640;       // warp 0: stream producer, warp 1: MMA issuer, warps 4..11: Hbar splitters + flushers, warps 12..19: Z splitters
constexpr int kDwStages = 4;          // operand stages (hi/lo split, UMMA core-matrix order), one 8-column k-step each
// raw FP32 stages filled by cp.async.bulk: 8 columns x W neurons of Hbar and of the stash.  The stash side turns 8 points
// (J k-steps) at a time into post-activation jets, so the ring must hold at least J stages.
__host__ __device__ constexpr int dw_raw_stages(int J) { return J > 5 ? J : 5; }
constexpr int kDwFlushTiles = 24;     // accumulator is flushed to the slab after this many 16-point tiles (720 accumulations per chain)

struct TcDwParams {
    int64_t n_tiles;                  // 16-point tiles
    int L, W, MB, J;                  // W: rows (n) and columns (k) of the dW block one CTA accumulates (<= 256: 512 TMEM columns)
    int Wr;                           // neurons per jet column in the streams (the layer width); Wr = n_split * W
    int n_split;                      // 1, or 2: W = 512 is worked as 2 x 2 quadrants (n half, k half) of 256 x 256, CTA b takes
                                      // quadrant b % 4 and the tiles (b / 4), (b / 4) + gridDim / 4, ...
    const float* hbar;                // [L][tile][cols][W]
    const float* stash;               // forward stash (t | h_c), same layout: Z_{l-1} is recomputed from it
    int act, fourier_half;
    int64_t stash_layer_stride;
    int tiles_per_bw_tile;
    float* gpart;                     // per-CTA slabs; hidden dW of layer l at gp_off_hidden + (l-1)*W*W, [n][k]
    int64_t gp_stride, gp_off_hidden;
};

// The forward pass keeps only the stash (t | h_c); the post-activation jets Z_{l-1} this GEMM contracts with are recomputed
// here by the splitter threads of the stash side (a few FMAs per point and neuron; writing them out in the forward and reading
// them back cost 21 GB of HBM traffic each way and half of the stash memory).
// Data path: one producer thread streams the two FP32 streams with cp.async.bulk (8 columns = one contiguous 8*W*4 B segment
// per stream and k-step) into a ring of raw stages.  (Register-buffered loads by the splitter threads were latency-bound: a
// warp has only six scoreboards, so waiting for the oldest of six load groups also waited for the newest.)  One splitter
// thread per neuron and stream reads its 8 values from the raw stage, splits them 3xTF32 and writes them as K-major UMMA core
// matrices into the operand ring.
template <int NF, int SMASK>
__global__ void __launch_bounds__(kDwThreads, 1) jet_dw_tc_kernel(const __grid_constant__ TcDwParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J;
    constexpr int kDwRaw = dw_raw_stages(J);
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;
    const int W = p.W, MB = p.MB;
    constexpr int N = 16 * J;                     // jet columns per tile
    constexpr int n_ks = N / 8;                   // k-steps per tile (2 J)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);       // [kDwStages] splitters -> MMA (count = splitter warps)
    uint64_t* empty = full + kDwStages;                            // MMA -> splitters (tcgen05.commit)
    uint64_t* raw_full = empty + kDwStages;                        // [kDwRaw] producer (complete_tx) -> splitters
    uint64_t* raw_empty = raw_full + kDwRaw;                       // splitters -> producer (count = splitter warps)
    uint64_t* acc_done = raw_empty + kDwRaw;                       // MMA -> flushers
    uint64_t* acc_free = acc_done + 1;                             // flushers -> MMA (count = Hbar splitter warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);
    // operand stage: A_hi | A_lo | B_hi | B_lo, each W rows x 8 k as K-major SWIZZLE_NONE core matrices: [kgroup 2][W/8 row groups][8][4]
    const uint32_t op_bytes = (uint32_t)W * 8u * 4u;
    const uint32_t stage_bytes = 4u * op_bytes;
    unsigned char* stages = smem_raw + 1024;
    // raw stage: Hbar [8 cols][W] | Z [8 cols][W]
    const uint32_t raw_bytes = 2u * op_bytes;
    unsigned char* raws = stages + (size_t)kDwStages * stage_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ld_warps = 4 * MB;          // per operand
    if (tid == 0) {
        for (int s = 0; s < kDwStages; ++s) {
            mbar_init(&full[s], 2 * n_ld_warps);
            mbar_init(&empty[s], 1);
        }
        for (int r = 0; r < kDwRaw; ++r) {
            mbar_init(&raw_full[r], 1);
            mbar_init(&raw_empty[r], 2 * n_ld_warps);
        }
        mbar_init(acc_done, 1);
        mbar_init(acc_free, n_ld_warps);
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

    // tiles of this CTA: group, group + n_groups, ...   (every CTA has at least one); quadrant (n half, k half) of the layer
    const int n_quads = p.n_split * p.n_split;
    const int quad = (int)blockIdx.x % n_quads, group = (int)blockIdx.x / n_quads, n_groups = (int)gridDim.x / n_quads;
    const int n0 = (quad / p.n_split) * W, k0 = (quad % p.n_split) * W;       // first Hbar neuron / first Z neuron of my quadrant
    const int Wr = p.Wr;
    const int64_t my_tiles = (p.n_tiles - group + n_groups - 1) / n_groups;

    if (warp < 4) {
        if (tid == 0) {
            // ===== producer: both streams, layer by layer, tile by tile, k-step by k-step =====
            int r = 0;
            uint32_t rph = 0;
            for (int l = 1; l < p.L; ++l) {
                const float* hsrc = p.hbar + (size_t)l * p.stash_layer_stride + n0;
                const float* zsrc = p.stash + (size_t)(l - 1) * p.stash_layer_stride + k0;
                for (int64_t t = 0; t < my_tiles; ++t) {
                    const int64_t tile = group + t * n_groups;
                    const size_t base = (size_t)tile * N * Wr;
                    for (int ks = 0; ks < n_ks; ++ks) {
                        mbar_wait(&raw_empty[r], rph ^ 1);
                        mbar_expect_tx(&raw_full[r], raw_bytes);
                        unsigned char* dst = raws + (size_t)r * raw_bytes;
                        if (Wr == W) {        // 8 columns of W neurons are one contiguous segment
                            bulk_g2s(dst, hsrc + base + (size_t)ks * 8 * W, op_bytes, &raw_full[r]);
                            bulk_g2s(dst + op_bytes, zsrc + base + (size_t)ks * 8 * W, op_bytes, &raw_full[r]);
                        } else {              // my W of the column's Wr neurons: one segment per column
                            const uint32_t seg = (uint32_t)W * 4u;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                bulk_g2s(dst + c * seg, hsrc + base + (size_t)(ks * 8 + c) * Wr, seg, &raw_full[r]);
                                bulk_g2s(dst + op_bytes + c * seg, zsrc + base + (size_t)(ks * 8 + c) * Wr, seg, &raw_full[r]);
                            }
                        }
                        if (++r == kDwRaw) {
                            r = 0;
                            rph ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            const bool leader = elect_one();
            // D = F32, A = B = TF32 K-major, N = W (k neurons), M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(W >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t lbo = (uint32_t)(W / 8) * 128u;        // bytes between the two k-groups
            const uint64_t d0 = umma_desc(smem_u32(stages), lbo, 128);
            int s = 0;
            uint32_t ph = 0, free_ph = 0;
            for (int l = 1; l < p.L; ++l) {
                for (int64_t t0 = 0; t0 < my_tiles; t0 += kDwFlushTiles) {
                    const int64_t t1 = t0 + kDwFlushTiles < my_tiles ? t0 + kDwFlushTiles : my_tiles;
                    mbar_wait(acc_free, free_ph ^ 1);    // the previous flush has drained TMEM (passes immediately the first time)
                    free_ph ^= 1;
                    tc_fence_after();
                    uint32_t accum = 0;
                    for (int64_t t = t0; t < t1; ++t) {
                        for (int ks = 0; ks < n_ks; ++ks) {
                            mbar_wait(&full[s], ph);
                            tc_fence_after();
                            if (leader) {
                                const uint64_t st = d0 + (uint64_t)((uint32_t)s * (stage_bytes >> 4));
                                const uint64_t a_hi = st, a_lo = st + (uint64_t)(op_bytes >> 4);
                                const uint64_t b_hi = st + (uint64_t)(2 * (op_bytes >> 4)), b_lo = st + (uint64_t)(3 * (op_bytes >> 4));
                                for (int mb = 0; mb < MB; ++mb) {
                                    const uint64_t aoff = (uint64_t)(mb * ((16u * 128u) >> 4));   // 128 rows = 16 row groups further
                                    const uint32_t d = tmem_base + (uint32_t)(mb * W);
                                    umma_tf32(d, a_lo + aoff, b_hi, idesc, accum);
                                    umma_tf32(d, a_hi + aoff, b_lo, idesc, 1);
                                    umma_tf32(d, a_hi + aoff, b_hi, idesc, 1);
                                }
                                umma_commit(&empty[s]);
                            }
                            accum = 1;
                            __syncwarp();
                            if (++s == kDwStages) {
                                s = 0;
                                ph ^= 1;
                            }
                        }
                    }
                    if (leader) umma_commit(acc_done);
                    __syncwarp();
                }
            }
        }
    } else {
        const int which = (warp - 4) >> 3;        // 0: Hbar splitter (+ flusher), 1: Z splitter
        const int e = (warp - 4) & 7;
        if (e < n_ld_warps) {
            const int quarter = warp & 3;
            const int mb = e >> 2;
            const int n = mb * 128 + quarter * 32 + lane;    // neuron handled by this thread; TMEM lane it flushes
            float* gp = p.gpart + (size_t)blockIdx.x * p.gp_stride + p.gp_off_hidden;
            // byte offset of row n inside an operand block: [kgroup][row group][row][4 k]
            const uint32_t rowoff = (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u;
            const uint32_t lbo = (uint32_t)(W / 8) * 128u;
            int s = 0, r = 0;
            uint32_t ph = 0, rph = 0, done_ph = 0;
            bool first_flush_of_layer;
            for (int l = 1; l < p.L; ++l) {
                first_flush_of_layer = true;
                for (int64_t t0 = 0; t0 < my_tiles; t0 += kDwFlushTiles) {
                    const int64_t t1 = t0 + kDwFlushTiles < my_tiles ? t0 + kDwFlushTiles : my_tiles;
                    const int total = (int)(t1 - t0) * n_ks;
                    auto read_raw = [&](float (&h8)[8]) {       // my 8 values of the next k-step from the raw stage
                        mbar_wait_park(&raw_full[r], rph);
                        const float* rawp = reinterpret_cast<const float*>(raws + (size_t)r * raw_bytes + (size_t)which * op_bytes) + n;
#pragma unroll
                        for (int c = 0; c < 8; ++c) h8[c] = rawp[c * W];
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&raw_empty[r]);
                        if (++r == kDwRaw) {
                            r = 0;
                            rph ^= 1;
                        }
                    };
                    auto emit = [&](const float (&h8)[8]) {     // split and lay out as operand of the next k-step
                        mbar_wait_park(&empty[s], ph ^ 1);
                        unsigned char* st = stages + (size_t)s * stage_bytes + (size_t)which * 2 * op_bytes;   // A_hi|A_lo or B_hi|B_lo
#pragma unroll
                        for (int kg = 0; kg < 2; ++kg) {
                            float4 hh, hl;
                            split_tf32(h8[4 * kg + 0], hh.x, hl.x);
                            split_tf32(h8[4 * kg + 1], hh.y, hl.y);
                            split_tf32(h8[4 * kg + 2], hh.z, hl.z);
                            split_tf32(h8[4 * kg + 3], hh.w, hl.w);
                            const uint32_t o = (uint32_t)kg * lbo + rowoff;
                            *reinterpret_cast<float4*>(st + o) = hh;
                            *reinterpret_cast<float4*>(st + op_bytes + o) = hl;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                        if (++s == kDwStages) {
                            s = 0;
                            ph ^= 1;
                        }
                    };
                    if (which == 0) {
                        for (int j = 0; j < total; ++j) {
                            float h8[8];
                            read_raw(h8);
                            emit(h8);
                        }
                    } else {
                        // stash side: 8 points = 8 J columns = J k-steps at a time: (t | h_c) -> Z jets of layer l-1, neuron n
                        const int actn = layer_act(p.act, p.fourier_half, l - 1, k0 + n);
                        for (int j = 0; j < total; j += J) {
                            // every output k-step is emitted as soon as the points that touch its 8 columns are converted:
                            // reads, conversions and emits interleave (all indices compile-time) instead of J reads, 8
                            // conversions, J emits in a row, which starved the MMAs between bursts
                            float v[J][8];
                            static_for<J>([&](auto ks_c) {
                                constexpr int ks = decltype(ks_c)::value;
                                constexpr int prev_pt = ks == 0 ? -1 : (8 * (ks - 1) + 7) / J;      // last point converted so far
                                constexpr int need_pt = (8 * ks + 7) / J;                            // last point touching k-step ks
                                constexpr int prev_raw = prev_pt < 0 ? -1 : ((prev_pt + 1) * J - 1) / 8;
                                constexpr int need_raw = ((need_pt + 1) * J - 1) / 8;                // last raw k-step that point needs
                                static_for<need_raw - prev_raw>([&](auto i_c) { read_raw(v[prev_raw + 1 + decltype(i_c)::value]); });
                                static_for<need_pt - prev_pt>([&](auto i_c) {
                                    constexpr int pt = prev_pt + 1 + decltype(i_c)::value;
                                    float st[J], z[J];
#pragma unroll
                                    for (int c = 0; c < J; ++c) st[c] = v[(pt * J + c) >> 3][(pt * J + c) & 7];
                                    act_recompute<S>(actn, st, z);
#pragma unroll
                                    for (int c = 0; c < J; ++c) v[(pt * J + c) >> 3][(pt * J + c) & 7] = z[c];
                                });
                                emit(v[ks]);
                            });
                        }
                    }
                    if (which) continue;      // Z splitters do not own TMEM lanes
                    // ---- flush the accumulator rows of this thread's neuron into the slab ----
                    mbar_wait(acc_done, done_ph);
                    done_ph ^= 1;
                    tc_fence_after();
                    float* dst = gp + (size_t)(l - 1) * W * W + (size_t)n * W;
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * W);
                    for (int c0 = 0; c0 < W; c0 += 16) {
                        float v[16];
                        tmem_ld16(taddr + c0, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            if (first_flush_of_layer) {
                                *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                            } else {
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(v[i]), "f"(v[i + 1]),
                                             "f"(v[i + 2]), "f"(v[i + 3])
                                             : "memory");
                            }
                        }
                    }
                    first_flush_of_layer = false;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_free);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace tde
