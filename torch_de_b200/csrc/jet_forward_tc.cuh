// tcgen05 forward for wide layers (padded width W = 128, 256 or 512): hidden-layer GEMMs on the 5th-gen tensor cores with a
// 3xTF32 split (hi*hi + hi*lo + lo*hi, FP32 accumulate in TMEM) to hold FP32 accuracy.
//
//   D[neuron][col] = sum_k W_l[neuron][k] * Z[k][col]        col = pt*J + channel, 16 points per CTA tile (W = 512: 8 or 4)
//
// A operand = weights (M = 128 neurons per block, K-major): streamed by a producer thread with cp.async.bulk (multicast to
//             the cluster peer) from a global image that is already in UMMA core-matrix order and already split into hi/lo.
// B operand = activations (N = 16*J jet columns, K-major SWIZZLE_128B): resident in shared memory, hi and lo column groups
//             side by side per k-atom (see TcPlanes / tc_issue_layer), written by the epilogue of the previous layer.
// D         = TMEM, lane = neuron => after tcgen05.ld every epilogue thread owns ALL jet channels of its neuron for its
//             points of the tile, which is exactly what the jet activation needs (no cross-lane exchange).
// Roles: warp 0 weight producer, warp 1 MMA issuer (one elected thread), warps 4..19 epilogue (W <= 256: every warp serves
// the neuron blocks in turn, one thread per neuron and 4 points; W = 512: one warp per block and TMEM lane quarter).
// Shared memory: [planes][weight stages][x][y][barriers]; w_out is read through L1.
#pragma once
#include <type_traits>
#include <utility>
#include "jet_common.cuh"

namespace tde {

constexpr int kTcPoints = 16;      // points per tile
constexpr int kTcKC = 8;           // k rows per weight chunk (= one tf32 MMA K step)
constexpr int kTcMaxStages = 8;
constexpr int kTcMaxBlocks = 4;     // 128-neuron blocks per layer (W <= 512)
constexpr int kTcTailBytes = 4 * 4 + (2 * kTcMaxStages + 2 * kTcMaxBlocks) * 8 + 16;   // b_out copy, barriers, TMEM slot   // weight stages of the tensor-core kernels (as many as shared memory allows)
// Epilogue warps.  A warp may only read the TMEM lanes 32*(warp%4)..+31, so every lane quarter gets the same number of warps.
//   W <= 256 ("turn" mode): 16 warps = 4 lane quarters x 4 groups of 4 points; EVERY warp serves the neuron blocks in turn
//            (block 0's accumulators, then block 1's), so a block's epilogue is done by all sixteen warps in half the time of
//            eight warps owning it -- the next layer's first MMAs wait for exactly that epilogue (period max(4Q, 3Q + E)).
//   W = 512: 16 warps = 4 blocks x 4 lane quarters, each owns one neuron block and the 8 points of the (8-point) tile.
__host__ __device__ constexpr bool tc_turn(int W) { return W <= 256; }
__host__ __device__ constexpr int tc_visit_points(int W, int pts) { return tc_turn(W) || pts < 8 ? 4 : 8; }     // points per thread and block visit
__host__ __device__ constexpr int tc_epi_warps(int W, int pts) {
    return tc_turn(W) ? 4 * (pts / tc_visit_points(W, pts)) : 4 * (W / 128) * (pts / tc_visit_points(W, pts));
}
// 4 control warps + the epilogue warps
__host__ __device__ constexpr int tc_threads(int W, int pts = 16) { return 128 + 32 * tc_epi_warps(W, pts); }
// registers per epilogue thread after setmaxnreg (control warps keep 24).  The pool is what the CTA was launched with:
// 640 threads x 96 registers, so (61440 - 128*24) / 512 = 114 -> 112; asking for more blocks forever
__host__ __device__ constexpr int tc_epi_regs(int threads) { return threads == 640 ? 112 : 232; }
// points per tile: the activation planes of a tile (W x 2 x PTS*J floats) must fit shared memory next to the weight stages
// (W = 512: 8 points; 4 when J >= 6, so that three accumulator sets of all four neuron blocks fit the 512 TMEM columns)
__host__ __device__ constexpr int tc_points(int W, int J) { return W >= 512 ? (J >= 6 ? 4 : 8) : 16; }
// k-steps (8 inputs) per streamed weight chunk of one neuron block: [hi 4 KB | lo 4 KB] each.  W = 512 with 48 jet columns
// has room for 8 KB stages only.
__host__ __device__ constexpr int tc_ksc(int W, int n) { return W == 128 ? 1 : (W == 512 && n > 40 ? 1 : 2); }
#ifndef TDE_TC_CLUSTER
#define TDE_TC_CLUSTER 2
#endif
constexpr int kTcCluster = TDE_TC_CLUSTER;      // CTAs per cluster: each loads 1/kTcCluster of every weight chunk and multicasts it to all

struct TcParams {
    const float* x;
    int64_t B;
    int64_t n_tiles;       // 16-point tiles
    int d_in, n_out, L, W, act, MB;   // MB = W / 128
    int n_stages;
    int fourier_half;      // > 0: layer 0 is a Fourier embedding (sin half | cos half)
    long long* dbg_buf;    // optional cycle stamps (TDE_TC_DBG bit 2)
    int dbg;               // bit 2: cycle stamps; bit 3: pair kernel without weight streaming (timing experiment only)
    int pair;              // W = 256: use the CTA-pair kernel
    int n_acc;             // informational: TMEM accumulator sets the kernel uses (tc_acc_sets; compile-time in the kernel)
    int first_in[TDE_MAX_FIRST];
    const float* w_in;     // [W][d_in]
    const float* b_in;     // [W]
    const float* wc;       // [(L-1)][W/8] chunks, each MB*2*2*2048 floats... see pack_tc_weights_kernel
    const float* b_hidden; // [(L-1)][W]
    const float* w_out;    // [n_out][W]
    const float* b_out;
    float* out;
    int64_t ld_out;
    float* stash;          // [L][tile][16*J columns][W]  (t | h_c), nullptr: no gradient wanted
    int64_t stash_layer_stride;
    int tiles_per_bw_tile; // FFMA tile / 16 (informational: the stash is addressed by 16-point tile)
};

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
    // K-major SWIZZLE_128B: rows of 128 B, 8-row groups sbo_bytes apart; layout_type = 2, version = 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// same copy with an L2 evict-last policy: the weight images (a few MB, re-read by every CTA for every tile) must stay in L2
// next to tens of GB of streamed stash / Hbar traffic
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_mc_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint16_t cta_mask,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
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

// LDW consecutive TMEM columns of this thread's lane -> registers (no wait: pair with tcgen05.wait::ld)
template <int LDW>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float (&v)[LDW]) {
    uint32_t r[LDW];
    if constexpr (LDW == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
    }
#pragma unroll
    for (int i = 0; i < LDW; ++i) v[i] = __uint_as_float(r[i]);
}

template <int NC>
__device__ __forceinline__ void tmem_ld_piece(uint32_t taddr, float* v) {
    uint32_t r[NC];
    if constexpr (NC == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
    } else if constexpr (NC == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    } else if constexpr (NC == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(taddr));
    } else if constexpr (NC == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r[i]);
}
// NC consecutive columns, NC a multiple of 4 (<= 40): split into power-of-two tcgen05.ld shapes.  No wait.
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
    if constexpr (NC >= 32) {
        tmem_ld_piece<16>(taddr, v);
        tmem_ld_piece<16>(taddr + 16, v + 16);
        if constexpr (NC > 32) tmem_ld_cols<NC - 32>(taddr + 32, v + 32);
    } else if constexpr (NC >= 16) {
        tmem_ld_piece<16>(taddr, v);
        if constexpr (NC > 16) tmem_ld_cols<NC - 16>(taddr + 16, v + 16);
    } else if constexpr (NC >= 8) {
        tmem_ld_piece<8>(taddr, v);
        if constexpr (NC > 8) tmem_ld_cols<NC - 8>(taddr + 8, v + 8);
    } else if constexpr (NC >= 4) {
        tmem_ld_piece<4>(taddr, v);
        if constexpr (NC > 4) tmem_ld_cols<NC - 4>(taddr + 4, v + 4);
    } else if constexpr (NC >= 2) {
        tmem_ld_piece<2>(taddr, v);
        if constexpr (NC > 2) tmem_ld_cols<NC - 2>(taddr + 2, v + 2);
    } else if constexpr (NC == 1) {
        tmem_ld_piece<1>(taddr, v);
    }
}
// zero NC consecutive TMEM columns of this thread's lane (no wait: pair with tcgen05.wait::st)
template <int NC>
__device__ __forceinline__ void tmem_zero_cols(uint32_t taddr) {
    const uint32_t z = 0u;
    if constexpr (NC >= 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
        if constexpr (NC > 16) tmem_zero_cols<NC - 16>(taddr + 16);
    } else if constexpr (NC >= 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
        if constexpr (NC > 8) tmem_zero_cols<NC - 8>(taddr + 8);
    } else if constexpr (NC >= 4) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
        if constexpr (NC > 4) tmem_zero_cols<NC - 4>(taddr + 4);
    } else if constexpr (NC >= 2) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%1};" ::"r"(taddr), "r"(z) : "memory");
        if constexpr (NC > 2) tmem_zero_cols<NC - 2>(taddr + 2);
    } else if constexpr (NC == 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(z) : "memory");
    }
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One layer's GEMM, issued by the converged MMA warp (one elected lane issues, the whole warp runs the control flow: a
// tcgen05.mma inside divergent code is wrapped by the compiler in an ELECT/branch loop of ~100 cycles each).
//
// Accuracy: the tensor core rounds its FP32 accumulator toward zero after every MMA; that bias, not the TF32 split, limits
// accuracy.  The 2^-11-sized cross terms get their own accumulator set and the main term alternates over two sets (by
// k-atom of 32 inputs); the epilogue adds them in FP32 (RN).  TMEM columns of neuron block mb: [main even | cross | main odd]
// (3 sets, N columns each) or [main | cross] (2 sets when 3 do not fit).
//
// Shared-memory traffic: these kernels are bound by the 128 B/clk of shared memory (MMA operand reads + weight-stage fills +
// epilogue stores = 15 k cycles per layer and tile with three MMAs per k-step).  A_hi x B_hi and A_hi x B_lo are therefore
// ONE MMA of N = 2 x 16J columns over the operand [B_hi | B_lo]: the activation planes keep, per k-atom, the 16J/8 column
// groups of hi and of lo next to each other (odd k-atoms in the order [lo | hi], so that the result lands on
// [cross | main odd]).  A_hi is read once instead of twice: 23.5 KB instead of 27.5 KB per k-step and neuron block.
//
// W = 256 (MB = 2): the layer is issued as four quarters (neuron block mb, input half kh) in the order
//   (0,0) (1,0) (0,1) (1,1), and each neuron block has its own act_ready / acc_ready barrier pair.  Quarter (0,0) needs only
// the new activations of neurons 0..127 and the drained accumulators of block 0, so it runs while the epilogue warps of
// block 1 still work on the previous layer; block 0's accumulators are complete after the third quarter, so its epilogue
// runs under the fourth.  The weight image is packed in exactly this order (one 16 KB chunk = one block, two k-steps).
template <int N>
struct TcPlanes {
    static constexpr uint32_t kHalf = (N / 8) * 1024;      // bytes of one part (hi or lo) of a k-atom: N/8 column groups
    static constexpr uint32_t kAtom = 2 * kHalf;           // bytes of a k-atom (32 inputs): both parts
    // byte offset of the hi part inside k-atom j; the lo part is the other half
    template <int kAccSets>
    __device__ static __forceinline__ uint32_t hi_off(const uint32_t j) { return kAccSets == 3 && (j & 1u) ? kHalf : 0u; }
};
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
// f(integral_constant<int, 0>) ... f(integral_constant<int, COUNT - 1>)
template <int COUNT, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, COUNT>{});
}
// The issue loop is fully unrolled (k-step, neuron block and k-atom parity are compile-time) and the descriptors are built
// from a constant high word and a 32-bit low word: the first version spent ~460 cycles per 16 KB chunk in 64-bit descriptor
// arithmetic on the uniform datapath -- more than the six MMAs of the chunk take -- and that, not the tensor pipe, set the
// length of the MMA phase.
template <int MB, int N, int kAccSets>
__device__ __forceinline__ void tc_issue_layer(const bool leader, const uint32_t tmem_base, const uint64_t a0, const uint64_t b0,
                                               const uint32_t stage_step, uint64_t* full, uint64_t* empty, uint64_t* act_ready,
                                               uint64_t* acc_ready, const int n_stages, int& s, uint32_t& ph, uint32_t& act_ph) {
    using P = TcPlanes<N>;
    static_assert(kAccSets == 2 || kAccSets == 3, "the merged hi*hi | hi*lo MMA needs a cross-term accumulator set");
    // instruction descriptors: D = F32, A = B = TF32 K-major, M = 128, N columns / both parts
    constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int n_ks = 16 * MB;           // k-steps of 8 inputs: W / 8
    const uint32_t a_lo32 = (uint32_t)a0, a_hi32 = (uint32_t)(a0 >> 32), b_lo32 = (uint32_t)b0, b_hi32 = (uint32_t)(b0 >> 32);
    auto desc = [](const uint32_t lo, const uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    // one k-step of neuron block mb; a_stage32: low descriptor word of the hi part of this k-step's weights (lo part: + 4 KB)
    auto kstep = [&](auto ks_c, auto mb_c, const uint32_t a_stage32) {
        constexpr int ks = decltype(ks_c)::value, mb = decltype(mb_c)::value;
        constexpr uint32_t j = (uint32_t)ks >> 2;                                  // k-atom, then 32 B per step inside the 128 B row
        constexpr uint32_t atom = j * P::kAtom + ((uint32_t)ks & 3u) * 32u;
        constexpr uint32_t hoff = (kAccSets == 3 && (j & 1u)) ? P::kHalf : 0u;
        const uint64_t a_hi = desc(a_stage32, a_hi32), a_lo = desc(a_stage32 + (4096u >> 4), a_hi32);
        const uint64_t b_both = desc(b_lo32 + (atom >> 4), b_hi32);                // [hi | lo], odd atoms: [lo | hi]
        const uint64_t b_hi = desc(b_lo32 + ((atom + hoff) >> 4), b_hi32);
        const uint32_t d = tmem_base + (uint32_t)(mb * kAccSets * N);
        if constexpr (kAccSets == 3 && hoff != 0) {
            if constexpr (ks == 4) {
                // first odd k-atom: "main odd" is not initialised yet, the cross set is -- two separate MMAs this once
                umma_tf32(d + 2 * N, a_hi, b_hi, idesc1, 0);
                umma_tf32(d + N, a_hi, b_both, idesc1, 1);       // N columns from the start of the atom = B_lo
            } else {
                umma_tf32(d + N, a_hi, b_both, idesc2, 1);       // [cross | main odd] += A_hi x [B_lo | B_hi]
            }
        } else {
            umma_tf32(d, a_hi, b_both, idesc2, ks > 0);          // [main (even) | cross] (+)= A_hi x [B_hi | B_lo]
        }
        umma_tf32(d + N, a_lo, b_hi, idesc1, 1);                 // cross += A_lo x B_hi
    };
    auto advance = [&]() {
        if (++s == n_stages) {
            s = 0;
            ph ^= 1;
        }
    };
    if constexpr (MB == 1) {
        mbar_wait(&act_ready[0], act_ph);   // planes of the previous layer are in smem, TMEM accumulators are drained
        tc_fence_after();
        static_for<n_ks>([&](auto ks_c) {      // one 8 KB chunk per k-step
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (leader) {
                kstep(ks_c, std::integral_constant<int, 0>{}, a_lo32 + (uint32_t)s * stage_step);
                umma_commit_mc(&empty[s], (uint16_t)((1u << kTcCluster) - 1));   // stage is free in BOTH CTAs' view
            }
            __syncwarp();
            advance();
        });
        if (leader) umma_commit(&acc_ready[0]);
        __syncwarp();
    } else {
        constexpr int ks_half = n_ks / 2;
        static_for<4>([&](auto q_c) {
            constexpr int q = decltype(q_c)::value, mb = q & 1, kh = q >> 1;
            if constexpr (q < 2) {
                mbar_wait(&act_ready[q], act_ph);
                tc_fence_after();
            }
            static_for<ks_half / 2>([&](auto cc_c) {      // 16 KB chunks: one block, two k-steps
                constexpr int cc = decltype(cc_c)::value;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_stage32 = a_lo32 + (uint32_t)s * stage_step;
                    kstep(std::integral_constant<int, kh * ks_half + cc * 2>{}, std::integral_constant<int, mb>{}, a_stage32);
                    kstep(std::integral_constant<int, kh * ks_half + cc * 2 + 1>{}, std::integral_constant<int, mb>{}, a_stage32 + (8192u >> 4));
                    umma_commit_mc(&empty[s], (uint16_t)((1u << kTcCluster) - 1));
                }
                __syncwarp();
                advance();
            });
            if constexpr (q >= 2) {
                if (leader) umma_commit(&acc_ready[q - 2]);
                __syncwarp();
            }
        });
    }
    act_ph ^= 1;
}

// accumulator sets that fit the 512 TMEM columns (3: cross terms | main even k-steps | main odd k-steps)
#ifndef TDE_MAX_ACC_SETS
#define TDE_MAX_ACC_SETS 3
#endif
__host__ __device__ constexpr int tc_acc_sets(int mb, int n) { return 512 / (mb * n) >= TDE_MAX_ACC_SETS ? TDE_MAX_ACC_SETS : 512 / (mb * n); }
// An M = 128 MMA takes N in multiples of 16.  A tile of 8 J columns with odd J (N = 8, 24, 40) therefore only has the merged
// 2N-wide MMA: both weight parts multiply [B_hi | B_lo] ("four products", the lo*lo term lands in the idle main set of the
// k-atom), which needs all three accumulator sets.
__host__ __device__ constexpr bool tc_four_products(int n) { return n % 16 != 0; }
// the single-CTA kernels need a cross-term set (>= 2) and a merged MMA of N = 2 x PTS J <= 256 columns
// W = 512 (mb = 4) always runs three sets: with two, every pre-activation is ONE chain of 64 round-toward-zero accumulations
// and ten layers of that put u_xx 1.0e-5 from the float64 reference (measured on the unsteady cylinder, J = 6, 8-point
// tiles); J = 6 therefore works 4-point tiles there (tc_points)
__host__ __device__ constexpr bool tc_shape_ok(int mb, int n) {
    if (mb >= 4) return n % 8 == 0 && 2 * n <= 256 && tc_acc_sets(mb, n) == 3;
    return 2 * n <= 256 && (tc_four_products(n) ? tc_acc_sets(mb, n) == 3 : tc_acc_sets(mb, n) >= 2);
}
// W = 512 (MB = 4 neuron blocks, 8- or 4-point tiles): the layer is issued as MB x MB units (k quarter kq = 128 inputs = the neurons
// of block kq, neuron block mb), kq-major: (0,0) (1,0) (2,0) (3,0) (0,1) ...  Unit (mb, 0) needs the previous layer's epilogue of
// block mb to have drained its accumulators and written its quarter of the planes -- one wait on act_ready[mb]; after the first
// MB units the whole plane is known to be complete.  Block mb's accumulators are complete after unit (mb, MB-1), so its epilogue
// runs under the remaining units of the layer and the first units of the next one: a window of MB-1 units (~3/16 of a layer).
// The order is forced by the in-place planes: block mb's epilogue overwrites the plane rows of k quarter mb, which no later
// unit of the layer reads (only k quarter MB-1 is still being read, and block MB-1 finishes last).  A block-major order would
// let every block's epilogue run under the following blocks, but those still read the rows it overwrites.
// Every MMA accumulates (the epilogue zeroes the TMEM columns it has read), so there is no first-k-step special case, and
// the k-step loop is unrolled per unit only (k-atom parity is the unit-local one because a unit starts on an even atom).
//   N % 16 == 0:  A_hi x [B_hi | B_lo] -> [main even | cross] resp. [cross | main odd];  A_lo x B_hi -> cross
//   N % 16 == 8:  an M = 128 MMA cannot be N wide, only 2N: A_lo multiplies [B_hi | B_lo] too; its lo*lo half -- a legitimate
//                 term of the exact product -- lands in the main set the k-atom does not use:
//                 even k-atom: A_hi -> [main even | cross], A_lo -> [cross | main odd];  odd: A_hi -> [cross | main odd], A_lo -> [main even | cross]
template <int MB, int N, int kAccSets, int KSC>
__device__ __forceinline__ void tc_issue_layer_wide(const bool leader, const uint32_t tmem_base, const uint64_t a0, const uint64_t b0,
                                                    const uint32_t stage_step, uint64_t* full, uint64_t* empty, uint64_t* act_ready,
                                                    uint64_t* acc_ready, const int n_stages, int& s, uint32_t& ph, uint32_t& act_ph) {
    using P = TcPlanes<N>;
    constexpr bool kFour = tc_four_products(N);
    static_assert(kAccSets == 3, "W = 512 runs three accumulator sets (tc_shape_ok)");
    constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lo32 = (uint32_t)a0, a_hi32 = (uint32_t)(a0 >> 32), b_lo32 = (uint32_t)b0, b_hi32 = (uint32_t)(b0 >> 32);
    auto desc = [](const uint32_t lo, const uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    for (int u = 0; u < MB * MB; ++u) {
        const int kq = u / MB, mb = u % MB;
        if (u < MB) {
            mbar_wait(&act_ready[u], act_ph);
            tc_fence_after();
        }
        const uint32_t d = tmem_base + (uint32_t)(mb * kAccSets * N);
        const uint32_t b_unit32 = b_lo32 + (((uint32_t)kq * 4u * P::kAtom) >> 4);       // 4 k-atoms of 32 inputs per unit
        static_for<16 / KSC>([&](auto cc_c) {
            constexpr int cc = decltype(cc_c)::value;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (leader) {
                const uint32_t a_stage32 = a_lo32 + (uint32_t)s * stage_step;
                static_for<KSC>([&](auto s2_c) {
                    constexpr int s2 = decltype(s2_c)::value, ks = cc * KSC + s2;          // k-step inside the unit
                    constexpr uint32_t j = (uint32_t)ks >> 2;
                    constexpr uint32_t atom = j * P::kAtom + ((uint32_t)ks & 3u) * 32u;
                    constexpr bool odd = (j & 1u) != 0;
                    const uint64_t a_hi = desc(a_stage32 + (uint32_t)(s2 * (8192 >> 4)), a_hi32);
                    const uint64_t a_lo = desc(a_stage32 + (uint32_t)(s2 * (8192 >> 4)) + (4096u >> 4), a_hi32);
                    const uint64_t b_both = desc(b_unit32 + (atom >> 4), b_hi32);        // [hi | lo], odd atoms: [lo | hi]
                    umma_tf32(d + (odd ? N : 0), a_hi, b_both, idesc2, 1);
                    if constexpr (kFour) {
                        umma_tf32(d + (odd ? 0 : N), a_lo, b_both, idesc2, 1);
                    } else {
                        const uint64_t b_hi = desc(b_unit32 + ((atom + (odd ? P::kHalf : 0u)) >> 4), b_hi32);
                        umma_tf32(d + N, a_lo, b_hi, idesc1, 1);
                    }
                });
                umma_commit_mc(&empty[s], (uint16_t)((1u << kTcCluster) - 1));
            }
            __syncwarp();
            if (++s == n_stages) {
                s = 0;
                ph ^= 1;
            }
        });
        if (kq == MB - 1) {
            if (leader) umma_commit(&acc_ready[mb]);
            __syncwarp();
        }
    }
    act_ph ^= 1;
}

constexpr int kTcGroup = 4;        // points handled per unrolled epilogue step (16 points = 4 steps)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

template <int NF, int SMASK, int W, int PTS = tc_points(W, JetSpecT<NF, SMASK>::J)>
__global__ void __launch_bounds__(tc_threads(W, PTS), 1) jet_forward_tc_kernel(const __grid_constant__ TcParams p) {
    using S = JetSpecT<NF, SMASK>;
    constexpr int J = S::J;
    constexpr int N = PTS * J;                   // jet columns per tile
    // activation planes: K-major SWIZZLE_128B atoms of 8 jet columns x 32 k (1024 B), ordered [k/32][col/8].  A warp's 32
    // consecutive neurons write one full 128 B row per column (conflict-free) and every MMA operand row is 128 B aligned.
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    unsigned char* const smem_raw = smem_tc;

    constexpr int MB = W / 128;
    constexpr int kAccSets = tc_acc_sets(MB, N);
    constexpr int KSC = tc_ksc(W, N);            // k-steps per weight chunk
    constexpr bool kZeroAcc = MB >= 4;           // W = 512: every MMA accumulates, the epilogue zeroes what it has read
    const int n_stages = p.n_stages;
    using P = TcPlanes<N>;
    unsigned char* zpl = smem_raw;                                   // activation planes: [k-atom][hi | lo part][column group][8 x 128 B]
    constexpr uint32_t z_planes = (uint32_t)(W / 32) * P::kAtom;
    constexpr uint32_t chunk_bytes = (uint32_t)KSC * 8192u;         // one neuron block, KSC k-steps: [hi|lo][kgroup 2][16 row groups][8][4] floats
    unsigned char* wst = zpl + z_planes;                             // [n_stages][chunk_bytes]
    // (w_out, 3 KB for three outputs, is read through L1 where it is needed: those 3 KB are the difference between three
    // and four 16 KB weight stages at W = 256, J = 5, and the MMA phase runs at the rate the stages in flight allow)
    float* xs = reinterpret_cast<float*>(wst + (size_t)n_stages * chunk_bytes);            // [PTS][d_in]
    float* ybuf = xs + PTS * p.d_in;             // [J*n_out][PTS]
    float* bout_s = ybuf + J * TDE_MAX_OUT * PTS; // [4]
    // barriers at the tail (kTcTailBytes): the planes need the 1024 B alignment of the base, a header would waste ~850 B
    uint64_t* full = reinterpret_cast<uint64_t*>(bout_s + 4);
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
    constexpr int n_chunks = (int)((2u * W * W * 4u) / chunk_bytes);      // weight chunks per layer

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTcCluster);   // one tcgen05.commit arrival from the MMA issuer of every CTA in the cluster
        }
        for (int b = 0; b < MB; ++b) {
            mbar_init(&act_ready[b], kArrivals);     // the epilogue warps that write neuron block b's activations
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
    cluster_sync_all();   // peers' barriers are initialised before any multicast copy / remote arrive can land
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t crank = cluster_ctarank();
    // tile schedule: cluster c handles tile pairs c, c + n_clusters, ...; rank r takes tile 2*pair + r.  Both CTAs run the
    // same number of iterations (a rank without a real tile computes a dummy one) so the shared weight pipeline stays in step.
    const int64_t n_pairs = (p.n_tiles + kTcCluster - 1) / kTcCluster;
    const int64_t first_pair = blockIdx.x / kTcCluster, pair_step = gridDim.x / kTcCluster;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == 0) {
            // ===== producer: weight chunks (already hi/lo split, UMMA core-matrix order) =====
            int s = 0;
            uint32_t ph = 0;
            const uint32_t part = chunk_bytes / kTcCluster;
            const uint64_t pol = l2_policy_evict_last();
            const bool hint = !(p.dbg & 16);
            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                for (int l = 1; l < p.L; ++l) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wc) + (size_t)(l - 1) * n_chunks * chunk_bytes;
                    for (int ks = 0; ks < n_chunks; ++ks) {
                        mbar_wait(&empty[s], ph ^ 1);   // both CTAs have finished reading this stage
                        mbar_expect_tx(&full[s], chunk_bytes);      // my part + the peer's part will land here
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
            // ===== MMA issuer: the whole warp runs the control flow converged, one elected lane issues =====
            // (a UTCHMMA inside divergent code is wrapped by the compiler in an ELECT/branch loop: ~100 cycles each)
            const bool leader = elect_one();
            const uint64_t b0 = umma_desc_sw128(smem_u32(zpl), 1024);
            const uint64_t a0 = umma_desc(smem_u32(wst), 2048, 128);
            const uint32_t stage_step = chunk_bytes >> 4;
            int s = 0;
            uint32_t ph = 0, act_ph = 0;
            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                const int64_t tile = pair * kTcCluster + crank;
                for (int l = 1; l < p.L; ++l) {
                    const bool mstamp = p.dbg_buf && blockIdx.x == 0 && tile == 0 && l == 2 && lane == 0;
                    if (mstamp) p.dbg_buf[8] = clock64();
                    if constexpr (MB >= 4)
                        tc_issue_layer_wide<MB, N, kAccSets, KSC>(leader, tmem_base, a0, b0, stage_step, full, empty, act_ready, acc_ready,
                                                                  n_stages, s, ph, act_ph);
                    else
                        tc_issue_layer<MB, N, kAccSets>(leader, tmem_base, a0, b0, stage_step, full, empty, act_ready, acc_ready,
                                                        n_stages, s, ph, act_ph);
                    if (mstamp) p.dbg_buf[10] = clock64();
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(tc_epi_regs(tc_threads(W, PTS))));
        const int e = warp - 4;                      // epilogue warp
        if (e < n_epi_warps) {
            const int quarter = warp & 3;            // TMEM lane quarter this warp may access
            const int own_mb = kTurn ? 0 : (e >> 2) % MB;     // neuron block (turn mode: the first one visited)
            const int pg = kTurn ? (e >> 2) : (e >> 2) / MB;  // point group: this thread handles points PPV*pg .. PPV*pg+PPV-1 of every tile
            constexpr int n_visits = kTurn ? MB : 1;          // neuron blocks a thread serves per layer
            if constexpr (kZeroAcc) {                // all accumulator columns of my lane start at zero (every MMA accumulates)
                tmem_zero_cols<kAccSets * N>(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(own_mb * kAccSets * N));
                tmem_wait_st();
            }
            const int n0 = own_mb * 128 + quarter * 32 + lane;   // the neuron this thread owns (turn mode: + 128 per visit)
            const int et = e * 32 + lane;            // 0..n_epi-1 dense index for strided loops
            auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory"); };

            if (et < p.n_out) bout_s[et] = p.b_out[et];
            // byte offset of element (col, k = n) inside a Z plane.  The k-atom parity (where the hi part sits), the byte inside the
            // 16 B chunk and the chunk index depend on the lane quarter and the lane only: a visit of block v adds 4 v k-atoms.
            const uint32_t hoff = P::template hi_off<kAccSets>((uint32_t)n0 >> 5);
            const uint32_t zrow = (uint32_t)(n0 >> 5) * P::kAtom + hoff + (uint32_t)(n0 & 3) * 4u;   // k-atom, hi part, byte within the 16 B chunk
            const uint32_t lo_delta = P::kHalf - 2u * hoff;                                       // hi -> lo part of the same k-atom (wraps for odd atoms)
            const uint32_t kchunk = (uint32_t)(n0 >> 2) & 7u;                                // 16 B chunk index before the swizzle
            // shared address of (column = first column of my point group + j, k = n0) in the hi part, j = 0..7: column c = 8 g + j of
            // my group is then zaddr0[j] + g * 1024 with compile-time g and j
            uint32_t zaddr0[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t cb = (uint32_t)(pg * PPV * J + j);
                zaddr0[j] = smem_u32(zpl) + zrow + (cb >> 3) * 1024u + (cb & 7u) * 128u + ((kchunk ^ (cb & 7u)) << 4);
            }
            uint32_t acc_ph = 0;

            for (int64_t pair = first_pair; pair < n_pairs; pair += pair_step) {
                const int64_t tile = pair * kTcCluster + crank;     // may be a padding tile (>= n_tiles, all points >= B)
                const int64_t pbase = tile * PTS;
                const bool tstamp = p.dbg_buf && blockIdx.x == 0 && et == 0 && pair == first_pair + 3 * pair_step;
                if (tstamp) p.dbg_buf[40] = clock64();      // tile start
                for (int i = et; i < PTS * p.d_in; i += n_epi) {
                    const int64_t gp = pbase + i / p.d_in;
                    xs[i] = gp < p.B ? p.x[gp * p.d_in + (i % p.d_in)] : 0.0f;
                }
                epi_sync();
                if (tstamp) p.dbg_buf[41] = clock64();      // inputs staged
                // [layer][tile][jet column][neuron]: this thread's neuron, first column of the tile.  A padding tile (the second
                // tile of the last pair when the tile count is odd) has stash rows of its own and x = 0.
                float* const stash_row = p.stash ? p.stash + ((size_t)tile * N + (size_t)pg * (PPV * J)) * W + n0 : nullptr;

                // Per layer the epilogue walks the 16 points as 2 x (2 steps of 4 points): 8 points are unrolled so every
                // shared/global offset is an immediate (a 16-point unroll is >80 KB of code and instruction-cache bound).
                // Base addresses pass through empty asm statements so the compiler keeps them in registers instead of
                // re-deriving each store address from the kernel parameters.
                constexpr int GC = kTcGroup * J;          // jet columns per 4-point step
                auto layer_body = [&](auto first_c, auto stash_c, const int l) {
                    constexpr bool kFirst = decltype(first_c)::value;     // input layer (no GEMM) vs hidden layers
                    constexpr bool kStash = decltype(stash_c)::value;
#pragma unroll 1
                    for (int v = 0; v < n_visits; ++v) {
                        const int mb = own_mb + v, n = n0 + v * 128;
                        float* srow_h = kStash ? stash_row + (size_t)l * p.stash_layer_stride + v * 128 : nullptr;
                        uint32_t zaddr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) zaddr[j] = zaddr0[j] + (uint32_t)v * 4u * P::kAtom;
                        float wrow[TDE_MAX_IN], wf[NF > 0 ? NF : 1];
                        float bn;
                        uint32_t taddr = 0;
                        const bool stamp = p.dbg_buf && blockIdx.x == 0 && (n & 127) == 0 && pg == 0 && pair == first_pair + 3 * pair_step && l >= 2 && l <= 4;
                        long long* const dst = p.dbg_buf + 16 + ((l - 2) * 2 + (mb & 1)) * 4;   // steady-state tile, two neuron blocks
                        if (stamp) dst[0] = clock64();
                        if constexpr (kFirst) {
#pragma unroll
                            for (int i = 0; i < TDE_MAX_IN; ++i) wrow[i] = i < p.d_in ? __ldg(p.w_in + n * p.d_in + i) : 0.0f;
                            bn = __ldg(p.b_in + n);
#pragma unroll
                            for (int a = 0; a < NF; ++a) wf[a] = __ldg(p.w_in + n * p.d_in + p.first_in[a]);
                        } else {
                            bn = __ldg(p.b_hidden + (size_t)(l - 1) * W + n);
                            mbar_wait_park(&acc_ready[mb], acc_ph);
                            tc_fence_after();
                            taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mb * kAccSets * N + pg * PPV * J);
                        }
                        const float* xs_h = xs + pg * PPV * p.d_in;
                        if (stamp) dst[1] = clock64();
                        {
                            if constexpr (kStash) asm volatile("" : "+l"(srow_h));
#pragma unroll
                            for (int j = 0; j < 8; ++j) asm volatile("" : "+r"(zaddr[j]));
#pragma unroll
                            for (int g2 = 0; g2 < PPV / kTcGroup; ++g2) {
                                float hg[GC];
                                if constexpr (!kFirst) {
                                    float cur[kAccSets][GC];
#pragma unroll
                                    for (int a = 0; a < kAccSets; ++a) tmem_ld_cols<GC>(taddr + (uint32_t)(a * N + g2 * GC), cur[a]);
                                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                                    if constexpr (kZeroAcc) {
#pragma unroll
                                        for (int a = 0; a < kAccSets; ++a) tmem_zero_cols<GC>(taddr + (uint32_t)(a * N + g2 * GC));
                                    }
#pragma unroll
                                    for (int i = 0; i < GC; ++i) {
                                        float vv = cur[0][i];
#pragma unroll
                                        for (int a = 1; a < kAccSets; ++a) vv += cur[a][i];
                                        hg[i] = vv;
                                    }
                                }
#pragma unroll
                                for (int q = 0; q < kTcGroup; ++q) {
                                    const int pt8 = g2 * kTcGroup + q;          // compile-time 0..PPV-1
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
                                        const int col8 = pt8 * J + c;           // column within my point group (compile-time)
                                        if constexpr (kStash) __stcs(srow_h + (size_t)col8 * W, h[c]);      // one 128 B line per warp store
                                        float hi, lo;
                                        split_tf32(z[c], hi, lo);
                                        const uint32_t a = zaddr[col8 & 7] + (uint32_t)(col8 >> 3) * 1024u;
                                        sts_f32(a, hi);
                                        sts_f32(a + lo_delta, lo);
                                    }
                                }
                            }
                        }
                        if (stamp) dst[2] = clock64();
                        if constexpr (kZeroAcc && !kFirst) tmem_wait_st();
                        if (l < p.L - 1) {
                            // publish: smem writes -> async proxy, TMEM reads retired, then one arrive per warp
                            fence_proxy_async();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&act_ready[mb]);
                        }
                        if (stamp) dst[3] = clock64();
                    }
                    if constexpr (!kFirst) acc_ph ^= 1;
                };
                if (stash_row) {
                    layer_body(std::true_type{}, std::true_type{}, 0);
                    for (int l = 1; l < p.L; ++l) layer_body(std::false_type{}, std::true_type{}, l);
                } else {
                    layer_body(std::true_type{}, std::false_type{}, 0);
                    for (int l = 1; l < p.L; ++l) layer_body(std::false_type{}, std::false_type{}, l);
                }
                epi_sync();   // every neuron's last-layer z is in the planes
                if (tstamp) p.dbg_buf[42] = clock64();      // last layer's epilogue done everywhere

                // ---- output layer, z = hi + lo.  A warp takes whole jet columns; its 32 lanes split the K = W reduction by 16 B
                // k-chunks (lane, lane + 32, ...), so a lane's rows of w_out are a handful of registers loaded once per tile and
                // the plane reads of a warp are four full 128 B rows per instruction (conflict-free).  w_out itself is read from
                // global memory (L1): shared memory has no room for it next to a fourth weight stage, and an inner loop that
                // fetched it per k-chunk doubled the time of this step (tensor pipe idle).
                {
                    constexpr int KPL = W / 128;         // k-chunks per lane
                    float4 wv[TDE_MAX_OUT][KPL];
#pragma unroll
                    for (int o = 0; o < TDE_MAX_OUT; ++o)
#pragma unroll
                        for (int j = 0; j < KPL; ++j)
                            wv[o][j] = o < p.n_out ? __ldg(reinterpret_cast<const float4*>(p.w_out + o * W + (lane + 32 * j) * 4))
                                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    for (int col = e; col < N; col += n_epi_warps) {       // whole warps: the shuffles below need all 32 lanes
                        float y[TDE_MAX_OUT];
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o) y[o] = 0.0f;
                        const uint32_t cbase = (uint32_t)(col >> 3) * 1024u + (uint32_t)(col & 7) * 128u;
#pragma unroll
                        for (int j = 0; j < KPL; ++j) {
                            const uint32_t kg = (uint32_t)(lane + 32 * j);
                            const uint32_t ho = P::template hi_off<kAccSets>(kg >> 3);
                            const uint32_t o4 = (kg >> 3) * P::kAtom + cbase + (((kg & 7u) ^ (uint32_t)(col & 7)) << 4);
                            const float4 a = *reinterpret_cast<const float4*>(zpl + o4 + ho);
                            const float4 b = *reinterpret_cast<const float4*>(zpl + o4 + (P::kHalf - ho));
                            const float zz[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
                            for (int o = 0; o < TDE_MAX_OUT; ++o)
                                y[o] = fmaf(zz[0], wv[o][j].x, fmaf(zz[1], wv[o][j].y, fmaf(zz[2], wv[o][j].z, fmaf(zz[3], wv[o][j].w, y[o]))));
                        }
#pragma unroll
                        for (int o = 0; o < TDE_MAX_OUT; ++o)
#pragma unroll
                            for (int sh = 16; sh > 0; sh >>= 1) y[o] += __shfl_xor_sync(0xffffffffu, y[o], sh);
                        if (lane == 0) {
                            const int pt = col / J, c = col - pt * J;
#pragma unroll
                            for (int o = 0; o < TDE_MAX_OUT; ++o)
                                if (o < p.n_out) ybuf[(c * p.n_out + o) * PTS + pt] = y[o] + (c == 0 ? bout_s[o] : 0.0f);
                        }
                    }
                }
                epi_sync();
                if (tstamp) p.dbg_buf[43] = clock64();      // output layer computed
                for (int i = et; i < J * p.n_out * PTS; i += n_epi) {
                    const int r = i / PTS, pt = i - r * PTS;
                    const int64_t gp = pbase + pt;
                    if (gp < p.B) p.out[(size_t)r * p.ld_out + gp] = ybuf[i];
                }
                epi_sync();
                if (tstamp) p.dbg_buf[44] = clock64();      // tile done
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA may exit while a peer can still multicast into it or arrive on its barriers
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace tde
