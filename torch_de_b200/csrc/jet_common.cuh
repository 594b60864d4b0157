// Shared pieces of the sm_100a jet kernels: compile-time jet spec, activation jets and their
// adjoints, mbarrier + bulk-async (TMA 1-D) helpers, and the kernel parameter block.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "tde_jet.h"

namespace tde {

constexpr int kComputeThreads = 256;            // 8 compute warps
constexpr int kThreads = kComputeThreads + 128;  // + 1 producer warpgroup (bulk-async weight streamer; donates registers)
constexpr int kNT = 8;                          // neurons per thread in the GEMM phases
constexpr int kKC = 16;                         // reduction rows per streamed weight chunk
constexpr int kMaxStages = 4;

// ---------------------------------------------------------------------------------------------
// Compile-time jet spec.  NF first-order channels; SMASK selects second-order pairs (a<=b over the
// first-order channels, enumerated (0,0),(0,1),..,(0,NF-1),(1,1),...).
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int popc_c(int m) { return m == 0 ? 0 : (m & 1) + popc_c(m >> 1); }
__host__ __device__ constexpr int pair_index(int nf, int a, int b) {
    // rank of (a,b), a<=b, in row-major upper-triangular order
    int idx = 0;
    for (int i = 0; i < a; ++i) idx += nf - i;
    return idx + (b - a);
}
__host__ __device__ constexpr int nth_set_bit(int mask, int n) {
    int pos = 0;
    while (true) {
        if (mask & 1) {
            if (n == 0) return pos;
            --n;
        }
        mask >>= 1;
        ++pos;
    }
}
__host__ __device__ constexpr int pair_a_of(int nf, int idx) {
    int a = 0;
    while (idx >= nf - a) {
        idx -= nf - a;
        ++a;
    }
    return a;
}
__host__ __device__ constexpr int pair_b_of(int nf, int idx) {
    int a = 0;
    while (idx >= nf - a) {
        idx -= nf - a;
        ++a;
    }
    return a + idx;
}

struct PairTables {
    int a[TDE_MAX_SECOND], b[TDE_MAX_SECOND];
};
__host__ __device__ constexpr PairTables make_pair_tables(int nf, int mask) {
    PairTables t{};
    int q = 0;
    for (int idx = 0; idx < 6 && q < TDE_MAX_SECOND; ++idx)
        if (mask >> idx & 1) {
            t.a[q] = pair_a_of(nf > 0 ? nf : 1, idx);
            t.b[q] = pair_b_of(nf > 0 ? nf : 1, idx);
            ++q;
        }
    return t;
}

template <int Q, int NF, int SMASK>
struct PairAt {
    static constexpr int idx = nth_set_bit(SMASK, Q);
    static constexpr int a = pair_a_of(NF, idx);
    static constexpr int b = pair_b_of(NF, idx);
};

template <int NF_, int SMASK_>
struct JetSpecT {
    static constexpr int NF = NF_;
    static constexpr int SMASK = SMASK_;
    static constexpr int NS = popc_c(SMASK_);
    static constexpr int J = 1 + NF_ + NS;
    static constexpr int PT = (J <= 5) ? 4 : 2;  // points per thread
    static constexpr int PJ = PT * J;            // jet columns per thread (multiple of 4)
    static constexpr int NV = PJ / 4;
    static_assert(PJ % 4 == 0, "thread jet columns must vectorise to float4");
};

// compile-time loop over the second-order pairs: f(integral_constant<q>, integral_constant<a>, integral_constant<b>)
template <class S, int Q = 0, class F>
__device__ __forceinline__ void for_each_pair(F&& f) {
    if constexpr (Q < S::NS) {
        f(Q, PairAt<Q, S::NF, S::SMASK>::a, PairAt<Q, S::NF, S::SMASK>::b);
        for_each_pair<S, Q + 1>(f);
    }
}

// ---------------------------------------------------------------------------------------------
// Kernel parameter block (passed by value, __grid_constant__)
// ---------------------------------------------------------------------------------------------
struct JetParams {
    const float* x;  // [B][d_in]
    int64_t B;
    int64_t n_tiles;
    int64_t stash_layer_stride;  // floats between layers of the stash (>= n_tiles * W * TP * J)
    int d_in, n_out, L;  // L = number of ACTIVATED layers (>= 1)
    int W;               // padded hidden width (multiple of 64)
    int act;
    int fourier_half;  // > 0: layer 0 is a Fourier embedding with this many frequencies (sin half, cos half)
    int PG;     // point groups per tile (tile points = PG * PT)
    int pitch;  // smem row pitch in floats (tile columns + pad, == 4 mod 8)
    int n_stages;
    int first_in[TDE_MAX_FIRST];
    // packed weights (workspace)
    const float* w_in;       // [W][d_in]
    const float* b_in;       // [W]
    const float* wt_hidden;  // [(L-1)][k=W][q=W]  forward operand:  Wt[k][q] = W_l[n(q)][k]
    const float* wn_hidden;  // [(L-1)][n=W][q=W]  reverse operand:  Wn[n][q] = W_l[n][k(q)]
    const float* b_hidden;   // [(L-1)][W]
    const float* w_out;      // [n_out][W]
    const float* b_out;      // [n_out]
    // streams
    float* out;
    int64_t ld_out;
    float* stash;  // [L][n_tiles][W][TP*J] or nullptr
    const float* seeds;
    int64_t ld_seeds;
    // backward partial gradients
    float* gpart;  // [n_ctas][gp_stride]
    int64_t gp_stride;
    int64_t gp_off_b_in, gp_off_hidden, gp_off_b_hidden, gp_off_w_out, gp_off_b_out;
    int kblk;    // k rows per dW block
    int rsplit;  // row (reduction) split of the dW phase
};

// column permutation of the packed weight matrices: thread group g (0..NG-1) owns neurons
// n = g + NG*j (j = 0..7) and reads them as two float4 at q = 4g + (j&3) + (j>>2)*(W/2).
__host__ __device__ inline int neuron_of_q(int q, int W) {
    const int NG = W / kNT;
    const int half = q >= W / 2;
    const int qq = q - half * (W / 2);
    const int g = qq >> 2;
    const int j = (qq & 3) + 4 * half;
    return g + NG * j;
}

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copy (TMA 1-D: SASS UBLKCP), named barriers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// many-warp wait (epilogue warps waiting for the MMAs): the suspend-time hint parks the warp in hardware until the phase
// completes instead of re-issuing try_wait / branch / yield -- those spin instructions were 41% of all issued instructions
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT_P:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE_P;\n"
        "bra LAB_WAIT_P;\n"
        "DONE_P:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// producer-side wait: long hardware suspend + nanosleep back-off (one thread; must not burn issue slots)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// barrier over the 256 compute threads only (the producer warp never joins it)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Activation jets.  h[0] value, h[1..NF] first order, h[1+NF..] second order (pairs pa/pb).
// sigma', sigma'', sigma''' from the stashed channel-0 quantity: tanh stores t = tanh(h0), sin stores h0.
// ---------------------------------------------------------------------------------------------
struct Sig {
    float s0, s1, s2, s3;
};
// activation codes: low nibble = outer function (TDE_ACT_TANH / TDE_ACT_SIN); kActInnerSin / kActInnerCos compose it with an
// inner sin / cos -- the Fourier embedding layer of torch_DE's Fourier_Net is act(cat[sin(Wx), cos(Wx)]) (Networks.py:53-75,104-120).
constexpr int kActInnerSin = 0x10;
constexpr int kActInnerCos = 0x20;

// sigma and its first three derivatives at v; `t_known`: v is already tanh(h) (the tanh stash) instead of h
__device__ __forceinline__ Sig sig_plain(int outer, float v, bool t_known) {
    Sig s;
    if (outer == TDE_ACT_TANH) {
        const float t = t_known ? v : tanhf(v);
        s.s0 = t;
        s.s1 = fmaf(-t, t, 1.0f);
        s.s2 = -2.0f * t * s.s1;
        s.s3 = -2.0f * s.s1 * fmaf(-3.0f * t, t, 1.0f);
    } else {
        float sn, cs;
        sincosf(v, &sn, &cs);
        s.s0 = sn;
        s.s1 = cs;
        s.s2 = -sn;
        s.s3 = -cs;
    }
    return s;
}
// g = outer(inner(h)), inner = sin or cos: chain rule up to third order
__device__ __forceinline__ Sig sig_composite(int act, float h0) {
    float sn, cs;
    sincosf(h0, &sn, &cs);
    float i0, i1, i2, i3;
    if (act & kActInnerSin) {
        i0 = sn; i1 = cs; i2 = -sn; i3 = -cs;
    } else {
        i0 = cs; i1 = -sn; i2 = -cs; i3 = sn;
    }
    const Sig o = sig_plain(act & 0xF, i0, false);
    Sig g;
    g.s0 = o.s0;
    g.s1 = o.s1 * i1;
    g.s2 = fmaf(o.s2 * i1, i1, o.s1 * i2);
    g.s3 = fmaf(o.s3 * i1 * i1, i1, fmaf(3.0f * o.s2 * i1, i2, o.s1 * i3));
    return g;
}
// what the stash keeps in channel 0: tanh keeps t = tanh(h0); sin and the composite activations keep h0 itself
__device__ __forceinline__ Sig sig_from_pre(int act, float h0, float& stash0) {
    if (act & (kActInnerSin | kActInnerCos)) {
        stash0 = h0;
        return sig_composite(act, h0);
    }
    const Sig s = sig_plain(act, h0, false);
    stash0 = act == TDE_ACT_TANH ? s.s0 : h0;
    return s;
}
__device__ __forceinline__ Sig sig_from_stash(int act, float st0) {
    if (act & (kActInnerSin | kActInnerCos)) return sig_composite(act, st0);
    return sig_plain(act, st0, true);
}
// activation of neuron n of activated layer l: layer 0 of a Fourier_Net is the embedding (first half sin, second half cos)
__device__ __forceinline__ int layer_act(int act, int fourier_half, int l, int n) {
    if (l != 0 || fourier_half == 0) return act;
    return act | (n < fourier_half ? kActInnerSin : kActInnerCos);
}

// z = act-jet(h); h[0] is replaced by the value to stash (t or h0); h[c>0] unchanged (they are stashed as is)
template <class S>
__device__ __forceinline__ void act_forward(int act, float (&h)[S::J], float (&z)[S::J]) {
    float st0;
    const Sig s = sig_from_pre(act, h[0], st0);
    z[0] = s.s0;
#pragma unroll
    for (int a = 0; a < S::NF; ++a) z[1 + a] = s.s1 * h[1 + a];
    for_each_pair<S>([&](const int q, const int a, const int b) {
        z[1 + S::NF + q] = fmaf(s.s2 * h[1 + a], h[1 + b], s.s1 * h[1 + S::NF + q]);
    });
    h[0] = st0;
}

// recompute z from the stashed (st0, h_c)
template <class S>
__device__ __forceinline__ void act_recompute(int act, const float (&st)[S::J], float (&z)[S::J]) {
    const Sig s = sig_from_stash(act, st[0]);
    z[0] = s.s0;
#pragma unroll
    for (int a = 0; a < S::NF; ++a) z[1 + a] = s.s1 * st[1 + a];
    for_each_pair<S>([&](const int q, const int a, const int b) {
        z[1 + S::NF + q] = fmaf(s.s2 * st[1 + a], st[1 + b], s.s1 * st[1 + S::NF + q]);
    });
}

// adjoint: given zb = dL/dz (jet) and the stash (st0, h_c) produce hb = dL/dh (jet), optionally z too
template <class S, bool WANT_Z>
__device__ __forceinline__ void act_backward(int act, const float (&st)[S::J], const float (&zb)[S::J],
                                             float (&hb)[S::J], float (&z)[S::J]) {
    const Sig s = sig_from_stash(act, st[0]);
    float h0b = s.s1 * zb[0];
#pragma unroll
    for (int a = 0; a < S::NF; ++a) {
        hb[1 + a] = s.s1 * zb[1 + a];
        h0b = fmaf(s.s2 * st[1 + a], zb[1 + a], h0b);
    }
    for_each_pair<S>([&](const int q, const int a, const int b) {
        const int ia = 1 + a, ib = 1 + b, iq = 1 + S::NF + q;
        const float zq = zb[iq];
        hb[iq] = s.s1 * zq;
        hb[ia] = fmaf(s.s2 * st[ib], zq, hb[ia]);
        hb[ib] = fmaf(s.s2 * st[ia], zq, hb[ib]);
        h0b = fmaf(fmaf(s.s3 * st[ia], st[ib], s.s2 * st[iq]), zq, h0b);
    });
    hb[0] = h0b;
    if (WANT_Z) {
        z[0] = s.s0;
#pragma unroll
        for (int a = 0; a < S::NF; ++a) z[1 + a] = s.s1 * st[1 + a];
        for_each_pair<S>([&](const int q, const int a, const int b) {
            z[1 + S::NF + q] = fmaf(s.s2 * st[1 + a], st[1 + b], s.s1 * st[1 + S::NF + q]);
        });
    }
}

// ---------------------------------------------------------------------------------------------
// The register-tile GEMM shared by the forward pass and the reverse dX pass:
//   acc[pc][j] += sum_k A[k][col0+pc] * Bq[k][q(g,j)]      (A: smem tile, K-major; Bq: streamed chunks)
// Consumers wait on full[stage], compute kKC reduction rows, release with one arrive per warp.
// ---------------------------------------------------------------------------------------------
// 3xTF32 split: hi = round-to-nearest TF32 of x, lo = round-to-nearest TF32 of (x - hi).  Rounding (not the
// hardware's truncation) keeps the split unbiased, so the dropped lo*lo / sub-TF32 bits do not accumulate over K.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// cvt.rna.tf32.f32 compiles to 4 instructions (it also handles NaN/Inf); for the finite values on this path the same
// rounding is an integer add of half a TF32 ulp and a mask.  The tensor core ignores the 13 low mantissa bits of a TF32
// operand, so the low part only needs the add (its mask is applied by the hardware).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}

// packed FP32 pairs (Blackwell fma.rn.f32x2: two FMAs per issue slot and per register-file read)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void fma2(uint64_t& d, uint64_t a, uint64_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// Accumulator tile of the register GEMM: PJ jet columns x 8 neurons, stored as column PAIRS
// (columns 2i, 2i+1 share one 64-bit register) so the inner product runs on FFMA2.
template <int PJ>
struct AccTile {
    uint64_t v[PJ / 2][kNT];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < PJ / 2; ++i)
#pragma unroll
            for (int j = 0; j < kNT; ++j) v[i][j] = 0ull;
    }
    __device__ __forceinline__ float get(int pc, int j) const {
        float lo, hi;
        unpack2(v[pc >> 1][j], lo, hi);
        return (pc & 1) ? hi : lo;
    }
};

template <int PJ>
__device__ __forceinline__ void gemm_chunk(AccTile<PJ>& acc, const float* __restrict__ Arow,
                                           const float* __restrict__ wchunk, int pitch, int W, int g) {
    constexpr int NV = PJ / 4;
#pragma unroll 4
    for (int kk = 0; kk < kKC; ++kk) {
        uint64_t a2[PJ / 2];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(Arow + kk * pitch + 4 * v);
            a2[2 * v] = t.x;
            a2[2 * v + 1] = t.y;
        }
        const float4 w0 = *reinterpret_cast<const float4*>(wchunk + kk * W + 4 * g);
        const float4 w1 = *reinterpret_cast<const float4*>(wchunk + kk * W + (W >> 1) + 4 * g);
        const uint64_t w2[kNT] = {pack2(w0.x, w0.x), pack2(w0.y, w0.y), pack2(w0.z, w0.z), pack2(w0.w, w0.w),
                                  pack2(w1.x, w1.x), pack2(w1.y, w1.y), pack2(w1.z, w1.z), pack2(w1.w, w1.w)};
#pragma unroll
        for (int i = 0; i < PJ / 2; ++i)
#pragma unroll
            for (int j = 0; j < kNT; ++j) fma2(acc.v[i][j], a2[i], w2[j]);
    }
}

}  // namespace tde
