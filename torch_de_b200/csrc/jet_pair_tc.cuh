// tcgen05 CTA-pair kernels for W = 256 (J <= 5): forward pass and reverse sweep with `cta_group::2` MMAs.
//
// Two CTAs (one cluster, two SMs of a TPC) share every layer GEMM of a 32-point pair tile, M = 256 neurons over both CTAs:
//   * CTA r owns neurons 128r .. 128r+127: it streams only ITS rows of the weights (A operand) and its TMEM holds their
//     accumulators -> half the weight bytes and half the A-operand reads per point and SM of the single-CTA kernel;
//   * CTA r keeps the activation planes (B operand, K-major SWIZZLE_128B) of ITS 16 points; the MMA reads half of its N
//     columns from each CTA;
//   * the pair tile is worked as two SUB-TILES of 2 x 8 points with their own TMEM regions and barriers: while the epilogue
//     warps of one sub-tile run, the MMAs of the other one do (per sub-tile the chain MMA -> epilogue -> MMA is serial);
//   * per k-step and sub-tile two MMAs: A_hi x [B_hi | B_lo] (N = 32 J: each CTA supplies the 8J hi and the 8J lo columns of
//     its 8 points, stored next to each other) and A_lo x B_hi (N = 16 J).  TMEM region of a sub-tile (48 J columns):
//     [c0 main | c0 hi*lo | c1 main | c1 hi*lo][c0 lo*hi | c1 lo*hi]; the epilogue adds the three in FP32;
//   * epilogue thread = (neuron, 8 points): it writes the new activation row of its neuron into the planes of the CTA that
//     owns those points -- its own shared memory (st.shared) or the peer's (st.shared::cluster);
//   * only the leader (rank 0) issues MMAs; the peer's MMA warp relays "my weight stage is full" to the leader.  tcgen05.commit
//     with a multicast mask releases the weight stage and publishes the accumulators in both CTAs.
//
// The main term accumulates over all 32 k-steps in one accumulator (no even/odd alternation: 2 x 48 J = 480 of the 512 TMEM
// columns are taken): worst stream error 6.5e-6 instead of 3.7e-6 (bar: 1e-5), the same as the FFMA path.
#pragma once
#include "jet_backward_tc.cuh"

namespace tde {

constexpr int kPairThreads = 640;     // 4 control warps + 16 epilogue warps (4 TMEM lane quarters x 4 groups of 8 points)
constexpr int kPairEpiRegs = 112;     // (640*96 - 128*24) / 512 = 114 -> 112 (the setmaxnreg pool is what the CTA launched with)
constexpr int kPairChunkBytes = 16384;   // one weight chunk: 128 rows x 2 k-steps x (hi | lo)
// A cluster holds two CTA pairs (ranks 0,1 and 2,3).  A cp.async.bulk stream runs at ~21 B/clk per issuing SM, half of what the
// MMAs of a pair consume, so ranks r and r^2 (same neuron block, different pair) each load half of every weight chunk and
// multicast it to both: every SM takes in 42 B/clk and L2 is read once per two pairs.  The two pairs share the weight ring
// and therefore advance layer by layer in step.
#ifndef TDE_PAIR_CLUSTER
#define TDE_PAIR_CLUSTER 2
#endif
constexpr int kPairCluster = TDE_PAIR_CLUSTER;   // 2: one pair per cluster (each CTA loads its own weight rows); 4: two pairs, multicast
constexpr int kPairsPerCluster = kPairCluster / 2;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// arrive on a barrier of any CTA of the cluster (address from mapa), releasing this thread's writes cluster-wide
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// "my weight stage landed" relay: nothing of this thread's own needs publishing (the data was written by the bulk-copy engine
// and is read by the tensor core), and a release at cluster scope costs ~700 cycles per arrive -- more than the MMAs of a chunk
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_R:\n"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_R;\n"
        "bra LAB_WAIT_R;\n"
        "DONE_R:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// wait on a barrier of this CTA whose arrivals may come from the peer
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_C;\n"
        "bra LAB_WAIT_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// shared-memory header of both pair kernels
struct PairBars {
    uint64_t full[kTcMaxStages];        // producer -> MMA: my weight stage landed
    uint64_t empty[kTcMaxStages];       // MMA (commit, both CTAs) -> producer
    uint64_t peer_full[kTcMaxStages];   // leader only: the peer's stage landed (relayed by the peer's warp 1)
    uint64_t act_ready[2];              // [sub-tile] leader only: its 16 epilogue warps (both CTAs) wrote the planes and drained TMEM
    uint64_t acc_ready[2];              // [sub-tile] MMA (commit, both CTAs) -> epilogue
    uint64_t planes_done;               // forward: the 16 warps (8 per CTA) that write MY planes finished the last layer
    uint64_t planes_free[2];            // forward: CTA d finished reading its planes (output layer); one arrival from CTA d
    uint32_t tmem_slot;
};

// plane geometry of one CTA: [k-atom of 32 inputs][sub-tile 2][part: hi, lo][J column groups of 8][8 x 128 B]
template <int J>
struct PairPlanes {
    static constexpr uint32_t kPart = (uint32_t)J * 1024u;      // 8J columns of one part
    static constexpr uint32_t kSub = 2u * kPart;                // hi | lo of a sub-tile
    static constexpr uint32_t kAtom = 2u * kSub;
    static constexpr uint32_t kBytes = 8u * kAtom;              // K = 256 inputs
};

// MMA issue for one layer by the leader's converged warp 1: sub-tile 0 then sub-tile 1, each over all 32 k-steps (the weights
// stream twice per layer).  Fully unrolled; descriptors = constant high word + 32-bit low word (see tc_issue_layer).
template <int J>
__device__ __forceinline__ void pair_issue_layer(const bool leader, const uint32_t tmem_base, const uint64_t a0, const uint64_t b0,
                                                 PairBars* bars, const int n_stages, int& s, uint32_t& ph, uint32_t& act_ph,
                                                 const uint16_t pair_mask, const bool no_weights = false) {
    using P = PairPlanes<J>;
    constexpr int NS = 8 * J;              // columns of (CTA, sub-tile, part)
    // D = F32, A = B = TF32 K-major, M = 256 (128 rows per CTA); N = 4 NS (both parts, both CTAs) / 2 NS (hi part, both CTAs)
    constexpr uint32_t idesc_both = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((4 * NS) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    constexpr uint32_t idesc_hi = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NS) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t a_lo32 = (uint32_t)a0, a_hi32 = (uint32_t)(a0 >> 32), b_lo32 = (uint32_t)b0, b_hi32 = (uint32_t)(b0 >> 32);
    auto desc = [](const uint32_t lo, const uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    static_for<2>([&](auto t_c) {
        constexpr int t = decltype(t_c)::value;
        mbar_wait_cluster(&bars->act_ready[t], act_ph);
        tc_fence_after();
        const uint32_t d_both = tmem_base + (uint32_t)(t * 6 * NS), d_lohi = d_both + (uint32_t)(4 * NS);
        static_for<16>([&](auto c_c) {          // 16 chunks x 2 k-steps = K 256
            constexpr int c = decltype(c_c)::value;
            if (!no_weights) {      // (timing experiment: MMAs on whatever is in the stages)
                mbar_wait(&bars->full[s], ph);
                mbar_wait_relaxed(&bars->peer_full[s], ph);
            }
            tc_fence_after();
            if (leader) {
                const uint32_t a_stage32 = a_lo32 + (uint32_t)s * (uint32_t)(kPairChunkBytes >> 4);
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int ks = c * 2 + s2;
                    const uint32_t boff = (uint32_t)(ks >> 2) * P::kAtom + (uint32_t)t * P::kSub + (uint32_t)(ks & 3) * 32u;
                    const uint64_t b = desc(b_lo32 + (boff >> 4), b_hi32);                 // starts at the hi part of sub-tile t
                    const uint64_t a_hi = desc(a_stage32 + (uint32_t)(s2 * (8192 >> 4)), a_hi32);
                    const uint64_t a_lo = desc(a_stage32 + (uint32_t)(s2 * (8192 >> 4)) + (4096u >> 4), a_hi32);
                    umma2_tf32(d_both, a_hi, b, idesc_both, ks > 0);      // [main | hi*lo] of both CTAs' points
                    umma2_tf32(d_lohi, a_lo, b, idesc_hi, ks > 0);        // lo*hi
                }
                if (!no_weights) umma2_commit_mc(&bars->empty[s], (uint16_t)((1u << kPairCluster) - 1));
            }
            __syncwarp();
            if (++s == n_stages) {
                s = 0;
                ph ^= 1;
            }
        });
        if (leader) umma2_commit_mc(&bars->acc_ready[t], pair_mask);
        __syncwarp();
    });
    act_ph ^= 1;
}

// the peer's warp 1: tell the leader when each of my weight stages has landed
__device__ __forceinline__ void pair_relay_layer(PairBars* bars, const uint32_t leader_peer_full0, const int n_stages, int& s, uint32_t& ph) {
    for (int c = 0; c < 32; ++c) {          // two passes over the layer's 16 chunks (one per sub-tile)
        mbar_wait(&bars->full[s], ph);
        if ((threadIdx.x & 31) == 0) mbar_arrive_cluster_relaxed(leader_peer_full0 + (uint32_t)s * 8u);
        __syncwarp();
        if (++s == n_stages) {
            s = 0;
            ph ^= 1;
        }
    }
}

// producer (one thread): the 128 weight rows of my neuron block for every layer, in k order.  The image is the one the
// single-CTA kernel streams ([layer][quarter = (neuron block, k half)][8 chunks]); block r reads quarters (r, 0) and (r, 1).
// I load half of each chunk (by pair index) and multicast it to the CTA of the other pair that owns the same block.
__device__ __forceinline__ void pair_produce_layer(PairBars* bars, unsigned char* wst, const unsigned char* layer_img, const uint32_t crank4,
                                                   const int n_stages, int& s, uint32_t& ph) {
    const uint32_t role = crank4 & 1u, half = crank4 >> 1;
    const uint16_t mask = kPairsPerCluster == 2 ? (uint16_t)((1u << role) | (1u << (role + 2))) : (uint16_t)(1u << role);
    constexpr uint32_t part = kPairChunkBytes / kPairsPerCluster;
    for (int cc = 0; cc < 32; ++cc) {       // two passes over the layer's 16 chunks (one per sub-tile)
        const int c = cc & 15;
        mbar_wait(&bars->empty[s], ph ^ 1);
        mbar_expect_tx(&bars->full[s], kPairChunkBytes);
        const int ci = ((c >> 3) * 2 + (int)role) * 8 + (c & 7);
        const size_t off = (size_t)half * part;
        bulk_g2s_mc(wst + (size_t)s * kPairChunkBytes + off, layer_img + (size_t)ci * kPairChunkBytes + off, part, &bars->full[s], mask);
        if (++s == n_stages) {
            s = 0;
            ph ^= 1;
        }
    }
}

__device__ __forceinline__ void pair_setup(PairBars* bars, const int n_stages, const int warp, const int tid) {
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], kPairsPerCluster);     // one commit from each pair's leader
            mbar_init(&bars->peer_full[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bars->act_ready[t], 16);
            mbar_init(&bars->acc_ready[t], 1);
        }
        mbar_init(&bars->planes_done, 16);
        mbar_init(&bars->planes_free[0], 1);
        mbar_init(&bars->planes_free[1], 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
}
__device__ __forceinline__ void pair_teardown(const uint32_t tmem_base, const int warp) {
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int NF, int SMASK>
__global__ void __launch_bounds__(kPairThreads, 1) jet_forward_pair_kernel(const __grid_constant__ TcParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int W = 256;
    constexpr int J = S::J;
    constexpr int NT = kTcPoints * J;            // jet columns of one CTA's 16-point tile
    constexpr int NS = 8 * J;                    // jet columns of one CTA's half of a sub-tile (8 points)
    using P = PairPlanes<J>;
    static_assert(12 * NS <= 512, "two sub-tile regions of 6 x 8J TMEM columns");
    constexpr int kAccSets = 3;                  // main | hi*lo | lo*hi
    constexpr int n_epi = 512;
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;
    PairBars* bars = reinterpret_cast<PairBars*>(smem_raw);
    unsigned char* zpl = smem_raw + 1024;            // planes, see PairPlanes
    unsigned char* wst = zpl + P::kBytes;
    const int n_stages = p.n_stages;
    float* wout_s = reinterpret_cast<float*>(wst + (size_t)n_stages * kPairChunkBytes);   // [n_out][W]
    float* xs = wout_s + p.n_out * W;                // [32][d_in]: both tiles of the pair
    float* ybuf = xs + 2 * kTcPoints * p.d_in;       // [J*n_out][16]: my tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pair_setup(bars, n_stages, warp, tid);
    const uint32_t tmem_base = bars->tmem_slot;
    const uint32_t crank4 = cluster_ctarank();        // 0..3
    const uint32_t crank = crank4 & 1u;               // rank inside my pair = neuron block
    const uint32_t pbase = crank4 & 2u;               // cluster rank of my pair's leader
    const int64_t n_quads = (p.n_tiles + kPairCluster - 1) / kPairCluster;      // a cluster takes two tiles per pair and step
    const int64_t first_quad = blockIdx.x / kPairCluster, quad_step = gridDim.x / kPairCluster;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t quad = first_quad; quad < n_quads && !(p.dbg & 8); quad += quad_step)
                for (int l = 1; l < p.L; ++l)
                    pair_produce_layer(bars, wst, reinterpret_cast<const unsigned char*>(p.wc) + (size_t)(l - 1) * 32 * kPairChunkBytes, crank4,
                                       n_stages, s, ph);
        } else if (warp == 1) {
            int s = 0;
            uint32_t ph = 0, act_ph = 0;
            if (crank == 0) {
                const bool leader = elect_one();
                const uint64_t b0 = umma_desc_sw128(smem_u32(zpl), 1024);
                const uint64_t a0 = umma_desc(smem_u32(wst), 2048, 128);
                for (int64_t quad = first_quad; quad < n_quads; quad += quad_step)
                    for (int l = 1; l < p.L; ++l)
                        pair_issue_layer<J>(leader, tmem_base, a0, b0, bars, n_stages, s, ph, act_ph, (uint16_t)(3u << pbase), (p.dbg & 8) != 0);
            } else {
                const uint32_t leader_peer_full0 = mapa_u32(smem_u32(&bars->peer_full[0]), pbase);
                for (int64_t quad = first_quad; quad < n_quads && !(p.dbg & 8); quad += quad_step)
                    for (int l = 1; l < p.L; ++l) pair_relay_layer(bars, leader_peer_full0, n_stages, s, ph);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPairEpiRegs));
        const int e = warp - 4;                      // epilogue warp 0..15
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int pq = e >> 2;                       // points 8*pq .. 8*pq+7 of the pair's 32
        const uint32_t dest = (uint32_t)(pq >> 1);   // CTA whose tile (and planes) these points belong to
        const int ph8 = pq & 1;                      // first / second 8 points of that tile
        const int n = (int)crank * 128 + quarter * 32 + lane;    // the neuron this thread owns
        const int et = e * 32 + lane;
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(n_epi) : "memory"); };

        for (int i = et; i < p.n_out * W; i += n_epi) wout_s[i] = p.w_out[i];
        // byte offset of element (col, k = n) inside a plane, then its address in the destination CTA
        const uint32_t zrow = (uint32_t)(n >> 5) * P::kAtom + (uint32_t)ph8 * P::kSub + (uint32_t)(n & 3) * 4u;   // k-atom, sub-tile, hi part
        const uint32_t kchunk = (uint32_t)(n >> 2) & 7u;
        const uint32_t zhi_dest = mapa_u32(smem_u32(zpl), pbase + dest);
        const uint32_t act_ready_leader = mapa_u32(smem_u32(&bars->act_ready[ph8]), pbase);
        const uint32_t planes_done_dest = mapa_u32(smem_u32(&bars->planes_done), pbase + dest);
        const uint32_t planes_free_mine[2] = {mapa_u32(smem_u32(&bars->planes_free[crank]), pbase), mapa_u32(smem_u32(&bars->planes_free[crank]), pbase + 1)};
        uint32_t acc_ph = 0, tile_ph = 0;

        for (int64_t quad = first_quad; quad < n_quads; quad += quad_step) {
            const int64_t pair = quad * kPairsPerCluster + (crank4 >> 1);
            const int64_t pbase2 = pair * 2 * kTcPoints;            // first point of the pair
            const int64_t tile_d = pair * 2 + dest;                 // tile my points belong to (may be padding: x = 0)
            for (int i = et; i < 2 * kTcPoints * p.d_in; i += n_epi) {
                const int64_t gp = pbase2 + i / p.d_in;
                xs[i] = gp < p.B ? p.x[gp * p.d_in + (i % p.d_in)] : 0.0f;
            }
            epi_sync();
            // the planes I write (CTA `dest`) are free once that CTA has finished the previous pair's output layer
            mbar_wait_cluster(&bars->planes_free[dest], tile_ph ^ 1);
            float* const stash_row = p.stash ? p.stash + ((size_t)tile_d * NT + (size_t)ph8 * (8 * J)) * W + n : nullptr;

            constexpr int GC = kTcGroup * J;          // jet columns per 4-point step
            auto layer_body = [&](auto first_c, auto stash_c, const int l) {
                constexpr bool kFirst = decltype(first_c)::value;     // input layer (no GEMM) vs hidden layers
                constexpr bool kStash = decltype(stash_c)::value;
                float* srow_h = kStash ? stash_row + (size_t)l * p.stash_layer_stride : nullptr;
                uint32_t zaddr[8];      // cluster address of (column & 7 == j, k = n) in the destination hi plane
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    zaddr[j] = zhi_dest + zrow + ((kchunk ^ (uint32_t)j) << 4) + (uint32_t)j * 128u;
                float wrow[TDE_MAX_IN], wf[NF > 0 ? NF : 1];
                float bn;
                uint32_t taddr = 0;
                const bool stamp = p.dbg_buf && blockIdx.x == 0 && (n & 127) == 0 && pq == 0 && quad == first_quad + 3 * quad_step && l >= 2 && l <= 4;
                long long* const dst = p.dbg_buf + 16 + ((l - 2) * 2) * 4;
                if (stamp) dst[0] = clock64();
                if constexpr (kFirst) {
#pragma unroll
                    for (int i = 0; i < TDE_MAX_IN; ++i) wrow[i] = i < p.d_in ? __ldg(p.w_in + n * p.d_in + i) : 0.0f;
                    bn = __ldg(p.b_in + n);
#pragma unroll
                    for (int a = 0; a < NF; ++a) wf[a] = __ldg(p.w_in + n * p.d_in + p.first_in[a]);
                } else {
                    bn = __ldg(p.b_hidden + (size_t)(l - 1) * W + n);
                    mbar_wait_park(&bars->acc_ready[ph8], acc_ph);
                    acc_ph ^= 1;
                    tc_fence_after();
                    taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ph8 * 6 * NS);     // my sub-tile's TMEM region
                }
                const float* xs_h = xs + pq * 8 * p.d_in;
                if (stamp) dst[1] = clock64();
                if constexpr (kStash) asm volatile("" : "+l"(srow_h));
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("" : "+r"(zaddr[j]));
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) {
                    float hg[GC];
                    if constexpr (!kFirst) {
                        float cur[kAccSets][GC];
#pragma unroll
                        for (int a = 0; a < kAccSets; ++a)      // main, hi*lo, lo*hi of my 8 points (owner CTA `dest`)
                            tmem_ld_cols<GC>(taddr + (a < 2 ? dest * 2 * NS + a * NS : 4 * NS + dest * NS) + (uint32_t)(g2 * GC), cur[a]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < GC; ++i) {
                            float v = cur[0][i];
#pragma unroll
                            for (int a = 1; a < kAccSets; ++a) v += cur[a][i];
                            hg[i] = v;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kTcGroup; ++q) {
                        const int pt8 = g2 * kTcGroup + q;          // compile-time 0..7
                        float h[J], z[J];
                        if constexpr (kFirst) {
                            float h0 = bn;
#pragma unroll
                            for (int i = 0; i < TDE_MAX_IN; ++i)
                                if (i < p.d_in) h0 = fmaf(wrow[i], xs_h[pt8 * p.d_in + i], h0);
                            h[0] = h0;
#pragma unroll
                            for (int a = 0; a < NF; ++a) h[1 + a] = wf[a];
#pragma unroll
                            for (int qq = 0; qq < S::NS; ++qq) h[1 + NF + qq] = 0.0f;
                        } else {
#pragma unroll
                            for (int c = 0; c < J; ++c) h[c] = hg[q * J + c];
                            h[0] += bn;
                        }
                        act_forward<S>(layer_act(p.act, p.fourier_half, l, n), h, z);      // h[0] now holds the value to stash (t | h0)
#pragma unroll
                        for (int c = 0; c < J; ++c) {
                            const int col8 = pt8 * J + c;           // column within my 8 points (compile-time)
                            if constexpr (kStash) __stcs(srow_h + (size_t)col8 * W, h[c]);      // one 128 B line per warp store
                            float hi, lo;
                            split_tf32(z[c], hi, lo);
                            const uint32_t a = zaddr[col8 & 7] + (uint32_t)(col8 >> 3) * 1024u;
                            st_cluster_f32(a, hi);
                            st_cluster_f32(a + P::kPart, lo);
                        }
                    }
                }
                if (stamp) dst[2] = clock64();
                // publish: (remote) smem writes -> async proxy, TMEM reads retired, then one arrive per warp
                fence_proxy_async_all();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (l < p.L - 1) mbar_arrive_cluster(act_ready_leader);
                    else mbar_arrive_cluster(planes_done_dest);
                }
                if (stamp) dst[3] = clock64();
            };
            if (stash_row) {
                layer_body(std::true_type{}, std::true_type{}, 0);
                for (int l = 1; l < p.L; ++l) layer_body(std::false_type{}, std::true_type{}, l);
            } else {
                layer_body(std::true_type{}, std::false_type{}, 0);
                for (int l = 1; l < p.L; ++l) layer_body(std::false_type{}, std::false_type{}, l);
            }
            mbar_wait_cluster(&bars->planes_done, tile_ph);   // every neuron's last-layer z of MY tile is in my planes

            // ---- output layer of my tile: one thread per jet column, z = hi + lo ----
            const int64_t pt_base = (pair * 2 + crank) * kTcPoints;
            for (int col = et; col < NT; col += n_epi) {
                float y[TDE_MAX_OUT];
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) y[o] = 0.0f;
                const int sub = col / NS, cs = col - sub * NS;       // sub-tile, column inside it
                const uint32_t cbase = (uint32_t)sub * P::kSub + (uint32_t)(cs >> 3) * 1024u + (uint32_t)(cs & 7) * 128u;
                for (int kg = 0; kg < W / 4; ++kg) {
                    const uint32_t o4 = (uint32_t)(kg >> 3) * P::kAtom + cbase + ((((uint32_t)kg & 7u) ^ (uint32_t)(cs & 7)) << 4);
                    const float4 a = *reinterpret_cast<const float4*>(zpl + o4);
                    const float4 b = *reinterpret_cast<const float4*>(zpl + o4 + P::kPart);
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
            epi_sync();      // my planes have been read; ybuf is complete
            if (et == 0) {
                mbar_arrive_cluster(planes_free_mine[0]);
                mbar_arrive_cluster(planes_free_mine[1]);
            }
            for (int i = et; i < J * p.n_out * kTcPoints; i += n_epi) {
                const int r = i / kTcPoints, pt = i - r * kTcPoints;
                const int64_t gp = pt_base + pt;
                if (gp < p.B) p.out[(size_t)r * p.ld_out + gp] = ybuf[i];
            }
            epi_sync();      // xs / ybuf are rewritten by the next pair
            tile_ph ^= 1;
        }
    }
    pair_teardown(tmem_base, warp);
}

// ---------------------------------------------------------------------------------------------------------------------
// reverse sweep
// ---------------------------------------------------------------------------------------------------------------------
template <int NF, int SMASK>
__global__ void __launch_bounds__(kPairThreads, 1) jet_backward_pair_kernel(const __grid_constant__ TcBwdParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int W = 256;
    constexpr int J = S::J;
    constexpr int NT = kTcPoints * J;
    constexpr int NS = 8 * J;
    using P = PairPlanes<J>;
    static_assert(12 * NS <= 512, "two sub-tile regions of 6 x 8J TMEM columns");
    constexpr int kAccSets = 3;                  // main | hi*lo | lo*hi
    constexpr int n_epi = 512;
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;
    PairBars* bars = reinterpret_cast<PairBars*>(smem_raw);
    unsigned char* zpl = smem_raw + 1024;            // planes, see PairPlanes
    unsigned char* wst = zpl + P::kBytes;
    const int n_stages = p.n_stages;
    float* wout_s = reinterpret_cast<float*>(wst + (size_t)n_stages * kPairChunkBytes);   // [n_out][W]
    float* xs = wout_s + p.n_out * W;                       // [32][d_in]
    float* seeds_s = xs + 2 * kTcPoints * p.d_in;           // [J*n_out][32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pair_setup(bars, n_stages, warp, tid);
    const uint32_t tmem_base = bars->tmem_slot;
    const uint32_t crank4 = cluster_ctarank();        // 0..3
    const uint32_t crank = crank4 & 1u;               // rank inside my pair = neuron block
    const uint32_t pbase = crank4 & 2u;               // cluster rank of my pair's leader
    const int64_t n_quads = (p.n_tiles + kPairCluster - 1) / kPairCluster;      // a cluster takes two tiles per pair and step
    const int64_t first_quad = blockIdx.x / kPairCluster, quad_step = gridDim.x / kPairCluster;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t quad = first_quad; quad < n_quads; quad += quad_step)
                for (int l = p.L - 1; l >= 1; --l)
                    pair_produce_layer(bars, wst, reinterpret_cast<const unsigned char*>(p.wct) + (size_t)(l - 1) * 32 * kPairChunkBytes, crank4,
                                       n_stages, s, ph);
        } else if (warp == 1) {
            int s = 0;
            uint32_t ph = 0, act_ph = 0;
            if (crank == 0) {
                const bool leader = elect_one();
                const uint64_t b0 = umma_desc_sw128(smem_u32(zpl), 1024);
                const uint64_t a0 = umma_desc(smem_u32(wst), 2048, 128);
                for (int64_t quad = first_quad; quad < n_quads; quad += quad_step)
                    for (int l = p.L - 1; l >= 1; --l)
                        pair_issue_layer<J>(leader, tmem_base, a0, b0, bars, n_stages, s, ph, act_ph, (uint16_t)(3u << pbase));
            } else {
                const uint32_t leader_peer_full0 = mapa_u32(smem_u32(&bars->peer_full[0]), pbase);
                for (int64_t quad = first_quad; quad < n_quads; quad += quad_step)
                    for (int l = p.L - 1; l >= 1; --l) pair_relay_layer(bars, leader_peer_full0, n_stages, s, ph);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPairEpiRegs));
        const int e = warp - 4;
        const int quarter = warp & 3;
        const int pq = e >> 2;                       // points 8*pq .. 8*pq+7 of the pair's 32
        const uint32_t dest = (uint32_t)(pq >> 1);
        const int ph8 = pq & 1;
        const int n = (int)crank * 128 + quarter * 32 + lane;    // the neuron this thread owns (in every layer)
        const int et = e * 32 + lane;
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(n_epi) : "memory"); };

        for (int i = et; i < p.n_out * W; i += n_epi) wout_s[i] = p.w_out[i];
        const uint32_t zrow = (uint32_t)(n >> 5) * P::kAtom + (uint32_t)ph8 * P::kSub + (uint32_t)(n & 3) * 4u;   // k-atom, sub-tile, hi part
        const uint32_t kchunk = (uint32_t)(n >> 2) & 7u;
        const uint32_t zhi_dest = mapa_u32(smem_u32(zpl), pbase + dest);
        const uint32_t act_ready_leader = mapa_u32(smem_u32(&bars->act_ready[ph8]), pbase);
        // both CTAs of a pair add into the pair's slab (disjoint neurons); the four threads of a neuron (one per group of 8
        // points) into four copies of the small sections, which the reduction sums in a fixed order
        float* gp = p.gpart + (size_t)(blockIdx.x >> 1) * p.gp_stride;
        float* gp_lo = gp + p.gp_dup_lo[pq];      // w_in | b_in
        float* gp_hi = gp + p.gp_dup_hi[pq];      // b_hidden | w_out
        uint32_t acc_ph = 0;

        for (int64_t quad = first_quad; quad < n_quads; quad += quad_step) {
            const int64_t pair = quad * kPairsPerCluster + (crank4 >> 1);
            const bool first = quad == first_quad;
            const int64_t pbase2 = pair * 2 * kTcPoints;
            const int64_t tile_d = pair * 2 + dest;
            for (int i = et; i < 2 * kTcPoints * p.d_in; i += n_epi) {
                const int64_t gpt = pbase2 + i / p.d_in;
                xs[i] = gpt < p.B ? p.x[gpt * p.d_in + (i % p.d_in)] : 0.0f;
            }
            for (int i = et; i < J * p.n_out * 2 * kTcPoints; i += n_epi) {
                const int r = i / (2 * kTcPoints), pt = i - r * (2 * kTcPoints);
                const int64_t gpt = pbase2 + pt;
                seeds_s[i] = gpt < p.B ? p.seeds[(size_t)r * p.ld_seeds + gpt] : 0.0f;
            }
            epi_sync();
            const size_t row0 = ((size_t)tile_d * NT + (size_t)ph8 * (8 * J)) * W + n;     // [layer][tile][jet column][neuron]

            constexpr int GC = kTcGroup * J;
            float wo[TDE_MAX_OUT];
#pragma unroll
            for (int o = 0; o < TDE_MAX_OUT; ++o) wo[o] = o < p.n_out ? wout_s[o * W + n] : 0.0f;

            auto layer_body = [&](auto kind_c, const int l) {
                constexpr int KIND = decltype(kind_c)::value;     // 2 last, 1 middle, 0 input layer
                constexpr bool last = KIND == 2;
                const size_t loff = (size_t)l * p.stash_layer_stride + row0;
                const float* srow_h = p.stash + loff;
                float* hrow_h = p.hbar + loff;
                uint32_t zaddr[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    zaddr[j] = zhi_dest + zrow + ((kchunk ^ (uint32_t)j) << 4) + (uint32_t)j * 128u;
                const bool stamp = p.dbg_buf && blockIdx.x == 0 && (n & 127) == 0 && pq == 0 && quad == first_quad + 3 * quad_step && l >= 3 && l <= 5;
                long long* const dst = p.dbg_buf + 16 + ((5 - l) * 2) * 4;
                if (stamp) dst[0] = clock64();
                float st_next[GC];
#pragma unroll
                for (int i = 0; i < GC; ++i) st_next[i] = __ldcs(srow_h + (size_t)i * W);
                uint32_t taddr = 0;
                if (!last) {
                    mbar_wait_park(&bars->acc_ready[ph8], acc_ph);
                    acc_ph ^= 1;
                    tc_fence_after();
                    taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ph8 * 6 * NS);     // my sub-tile's TMEM region
                }
                if (stamp) dst[1] = clock64();
                float db = 0.0f;
                float dwo[TDE_MAX_OUT], dw_in[TDE_MAX_IN];
#pragma unroll
                for (int o = 0; o < TDE_MAX_OUT; ++o) dwo[o] = 0.0f;
#pragma unroll
                for (int i = 0; i < TDE_MAX_IN; ++i) dw_in[i] = 0.0f;
                const float* xs_h = xs + pq * 8 * p.d_in;
                const float* seeds_h = seeds_s + pq * 8;
                asm volatile("" : "+l"(srow_h), "+l"(hrow_h));
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("" : "+r"(zaddr[j]));
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) {
                    float stg[GC], zbg[GC];
#pragma unroll
                    for (int i = 0; i < GC; ++i) stg[i] = st_next[i];
                    if (g2 == 0) {
#pragma unroll
                        for (int i = 0; i < GC; ++i) st_next[i] = __ldcs(srow_h + (size_t)(GC + i) * W);
                    }
                    if (!last) {
                        float cur[kAccSets][GC];
#pragma unroll
                        for (int a = 0; a < kAccSets; ++a)      // main, hi*lo, lo*hi of my 8 points (owner CTA `dest`)
                            tmem_ld_cols<GC>(taddr + (a < 2 ? dest * 2 * NS + a * NS : 4 * NS + dest * NS) + (uint32_t)(g2 * GC), cur[a]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < GC; ++i) {
                            float v = cur[0][i];
#pragma unroll
                            for (int a = 1; a < kAccSets; ++a) v += cur[a][i];
                            zbg[i] = v;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kTcGroup; ++q) {
                        const int pt8 = g2 * kTcGroup + q;      // compile-time 0..7
                        float st[J], zb[J], hb[J], z[J];
#pragma unroll
                        for (int c = 0; c < J; ++c) st[c] = stg[q * J + c];
                        if (last) {
#pragma unroll
                            for (int c = 0; c < J; ++c) {
                                float a = 0.0f;
#pragma unroll
                                for (int o = 0; o < TDE_MAX_OUT; ++o)
                                    if (o < p.n_out) a = fmaf(wo[o], seeds_h[(c * p.n_out + o) * (2 * kTcPoints) + pt8], a);
                                zb[c] = a;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < J; ++c) zb[c] = zbg[q * J + c];
                        }
                        act_backward<S, last>(layer_act(p.act, p.fourier_half, l, n), st, zb, hb, z);     // Z_l was written by the forward pass
                        db += hb[0];
                        if (last) {
#pragma unroll
                            for (int c = 0; c < J; ++c)
#pragma unroll
                                for (int o = 0; o < TDE_MAX_OUT; ++o)
                                    if (o < p.n_out) dwo[o] = fmaf(seeds_h[(c * p.n_out + o) * (2 * kTcPoints) + pt8], z[c], dwo[o]);
                        }
                        if constexpr (KIND >= 1) {
#pragma unroll
                            for (int c = 0; c < J; ++c) {
                                const int col8 = pt8 * J + c;
                                __stcs(hrow_h + (size_t)col8 * W, hb[c]);
                                float hi, lo;
                                split_tf32(hb[c], hi, lo);
                                const uint32_t a = zaddr[col8 & 7] + (uint32_t)(col8 >> 3) * 1024u;
                                st_cluster_f32(a, hi);
                                st_cluster_f32(a + P::kPart, lo);
                            }
                        } else {
                            // input layer: h = W0 x + b0, h_a = W0[:, first_in[a]]
                            const float h0 = hb[0];
#pragma unroll
                            for (int i = 0; i < TDE_MAX_IN; ++i)
                                if (i < p.d_in) dw_in[i] = fmaf(h0, xs_h[pt8 * p.d_in + i], dw_in[i]);
#pragma unroll
                            for (int a = 0; a < NF; ++a) {
#pragma unroll
                                for (int i = 0; i < TDE_MAX_IN; ++i)
                                    if (i == p.first_in[a]) dw_in[i] += hb[1 + a];
                            }
                        }
                    }
                }
                if (stamp) dst[2] = clock64();
                if constexpr (KIND >= 1) {
                    fence_proxy_async_all();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(act_ready_leader);
                    if (stamp) dst[3] = clock64();
                    slab_add(gp_hi + p.gp_off_b_hidden + (size_t)(l - 1) * W + n, db, first);
                } else {
                    slab_add(gp_lo + p.gp_off_b_in + n, db, first);
#pragma unroll
                    for (int i = 0; i < TDE_MAX_IN; ++i)
                        if (i < p.d_in) slab_add(gp_lo + (size_t)n * p.d_in + i, dw_in[i], first);
                }
                if (last) {
#pragma unroll
                    for (int o = 0; o < TDE_MAX_OUT; ++o)
                        if (o < p.n_out) slab_add(gp_hi + p.gp_off_w_out + (size_t)o * W + n, dwo[o], first);
                }
            };
            layer_body(std::integral_constant<int, 2>{}, p.L - 1);
            for (int l = p.L - 2; l >= 1; --l) layer_body(std::integral_constant<int, 1>{}, l);
            layer_body(std::integral_constant<int, 0>{}, 0);
            if (crank == 0 && et < p.n_out) {
                float s = 0.0f;
                for (int pt = 0; pt < 2 * kTcPoints; ++pt) s += seeds_s[(size_t)et * (2 * kTcPoints) + pt];   // channel-0 rows come first
                slab_add(gp + p.gp_off_b_out + et, s, first);
            }
            epi_sync();   // xs / seeds_s are rewritten by the next pair
        }
    }
    pair_teardown(tmem_base, warp);
}

}  // namespace tde
