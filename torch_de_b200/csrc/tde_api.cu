// C ABI of libtde_jet.so (see include/tde_jet.h): planning, weight packing, kernel dispatch,
// deterministic second-stage gradient reduction, fused residual/loss, Adam, FP32 peak probe.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <string>
#include <vector>

#include "jet_launch.h"

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define TDE_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e__));   \
    } while (0)

constexpr int kSmemLimit = 232448;  // 227 KB opt-in per CTA on sm_100

// ---- run-time switches (A/B measurements, debugging).  The environment is read ONCE, at the first call into the library;
// tde_set_option() overrides a switch afterwards.  Nothing on the per-launch path calls getenv.
struct Options {
    int no_tc = 0;        // TDE_NO_TC=1: FFMA2 kernels for every width
    int no_tc_bwd = 0;    // TDE_NO_TC_BWD=1: tensor-core forward, FFMA2 reverse sweep
    int tc_pair = 0;      // TDE_TC_PAIR=1: cta_group::2 kernels for W = 256
    int tc_slots = 0;     // TDE_TC_SLOTS=n: cap on resident clusters of the tensor-core kernels (0 = no cap)
    int tc_stages = 0;    // TDE_TC_STAGES=n: cap on weight stages (0 = as many as fit)
    int tc_dbg = 0;       // TDE_TC_DBG bit 2: tile-level cycle stamps (synchronises!)
};
Options& options() {
    static Options o = [] {
        Options v;
        auto rd = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; };
        v.no_tc = rd("TDE_NO_TC"); v.no_tc_bwd = rd("TDE_NO_TC_BWD"); v.tc_pair = rd("TDE_TC_PAIR");
        v.tc_slots = rd("TDE_TC_SLOTS"); v.tc_stages = rd("TDE_TC_STAGES"); v.tc_dbg = rd("TDE_TC_DBG");
        return v;
    }();
    return o;
}

// ---- instrumentation: launch counter and per-kernel CUDA-event timing (bench.py roofline) ----
long long g_launches = 0;
bool g_profiling = false;
struct TimedLaunch {
    cudaEvent_t e0, e1;
    int kind;  // 0 forward, 1 reverse sweep, 2 weight-gradient GEMM (tensor-core path only)
};
std::vector<TimedLaunch> g_timed;

struct ScopedTimer {
    cudaStream_t st;
    int kind;
    bool on;
    TimedLaunch t;
    ScopedTimer(cudaStream_t s, int k) : st(s), kind(k), on(g_profiling) {
        if (on) {
            cudaEventCreate(&t.e0);
            cudaEventCreate(&t.e1);
            t.kind = kind;
            cudaEventRecord(t.e0, st);
        }
    }
    ~ScopedTimer() {
        if (on) {
            cudaEventRecord(t.e1, st);
            g_timed.push_back(t);
        }
    }
};

struct Dispatch {
    int nf, smask;
    tde::JetLauncher fwd, bwd;
    tde::TcLauncher fwd_tc;
    tde::TcBwdLauncher bwd_tc;
    tde::TcDwLauncher dw_tc;
    int (*pair_max_clusters)(int smem);
};
#define TDE_ROW(nf, sm) {nf, sm, tde::launch_forward_##nf##_##sm, tde::launch_backward_##nf##_##sm, tde::launch_forward_tc_##nf##_##sm, tde::launch_backward_tc_##nf##_##sm, tde::launch_dw_tc_##nf##_##sm, tde::pair_max_clusters_##nf##_##sm},
const Dispatch kDispatch[] = {TDE_SPEC_LIST(TDE_ROW)};
#undef TDE_ROW

// resident-cluster capacity of the pair kernels, cached per (spec, smem)
int sm_count();
int current_device();

int pair_cluster_capacity(const Dispatch* d, int smem) {
    static std::mutex mu;
    static std::map<std::pair<const Dispatch*, std::pair<int, int>>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(d, std::make_pair(smem, current_device()));
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const int n = d->pair_max_clusters(smem);
    cache[key] = n;
    return n;
}

// clusters of the tensor-core kernels that can run at once: SMs / cluster size, for the pair kernels at most what the
// occupancy query reports (planning without a device assumes the full 37)
int tc_cluster_slots(const Dispatch* d, bool pair, int smem) {
    const int cl = pair ? tde::kPairCluster : tde::kTcCluster;
    int slots = sm_count() / cl;
    if (pair && tde::kPairCluster > 2) {
        const int cap = pair_cluster_capacity(d, smem);
        if (cap > 0) slots = std::min(slots, cap);
        else (void)cudaGetLastError();
    }
    if (options().tc_slots > 0) slots = std::min(slots, options().tc_slots);   // experiments only
    return std::max(1, slots);
}

const Dispatch* find_dispatch(int nf, int smask) {
    for (const Dispatch& d : kDispatch)
        if (d.nf == nf && d.smask == smask) return &d;
    return nullptr;
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Everything derived from (net, spec, n_points): geometry + workspace offsets (in floats).
struct Geometry {
    int nf, smask, J, PT, PJ;
    int width, W, L, d_in, n_out, act, fourier_half;
    int PG, TP, pitch, stages_fwd, stages_bwd, kblk, rsplit;
    int smem_fwd, smem_bwd;
    int n_ctas;
    int64_t n_tiles;
    // packed weight offsets
    int64_t o_w_in, o_b_in, o_wt, o_wn, o_bh, o_wo, o_bo, o_wc, packed_floats;
    // tcgen05 forward (wide layers)
    int use_tc, use_pair, tc_mb, tc_stages, tc_smem, tc_nacc;
    int tc_pts, tc_ksc;          // points per tensor-core tile (16; 8 for W = 512), k-steps per weight chunk
    int tc_dbuf;                 // reverse sweep: second (x | seeds) staging buffer fits
    int use_tc_bwd, dw_smem;
    int dw_w, dw_split;          // weight-gradient GEMM: block edge one CTA accumulates (<= 256), halves per dimension (W = 512: 2)
    int64_t o_wct, o_hbar, grid_b1, grid_b2;
    int64_t tc_tiles;
    // per-CTA partial gradient slab
    int64_t g_b_in, g_hidden, g_bh, g_wo, g_bo, gp_stride;
    int64_t o_gpart;
    int64_t workspace_floats;
    int64_t stash_floats, stash_total_floats, stash_layer_stride;
    int n_copies;
    int64_t g_dup_lo[4], g_dup_hi[4];
    int64_t n_params;
    const Dispatch* disp;
};

// SM count of the CURRENT device (the Python binding makes the tensors' device current around every call), cached per device
int sm_count() {
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return 148;  // planning without a device (unit tests on CPU): B200 has 148 SMs
    }
    if (cached[dev]) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
        cached[dev] = n;
        return n;
    }
    (void)cudaGetLastError();
    return 148;
}
int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    return dev;
}

int make_geometry(const TdeNet* net, const TdeJetSpec* spec, int64_t B, Geometry* G) {
    if (!net || !spec) return fail("null net/spec");
    if (net->n_layers < 2 || net->n_layers > TDE_MAX_LAYERS)
        return fail("need 2..%d Linear layers, got %d", TDE_MAX_LAYERS, net->n_layers);
    const int L = net->n_layers - 1;
    const int width = net->layers[0].out_features;
    const int fourier = net->layers[0].act & TDE_ACT_FOURIER;
    const int act = net->layers[0].act & ~TDE_ACT_FOURIER;
    if (act != TDE_ACT_TANH && act != TDE_ACT_SIN) return fail("unsupported activation code %d (tanh/sin only)", act);
    if (net->layers[0].in_features != spec->d_in) return fail("layer 0 in_features %d != d_in %d", net->layers[0].in_features, spec->d_in);
    if (spec->d_in < 1 || spec->d_in > TDE_MAX_IN) return fail("d_in %d out of range 1..%d", spec->d_in, TDE_MAX_IN);
    if (spec->n_out < 1 || spec->n_out > TDE_MAX_OUT) return fail("n_out %d out of range 1..%d", spec->n_out, TDE_MAX_OUT);
    for (int l = 0; l < net->n_layers; ++l) {
        const TdeLayer& ly = net->layers[l];
        if (!ly.weight || !ly.bias) return fail("layer %d: weight and bias are required", l);
        if (l < L) {
            if (ly.out_features != width) return fail("hidden widths must all be equal (layer %d has %d, expected %d)", l, ly.out_features, width);
            if ((ly.act & ~TDE_ACT_FOURIER) != act) return fail("all hidden activations must match");
            if (l > 0 && (ly.act & TDE_ACT_FOURIER)) return fail("only the first layer can be a Fourier embedding");
            if (l > 0 && ly.in_features != width) return fail("layer %d in_features %d != width %d", l, ly.in_features, width);
        } else {
            if (ly.act != TDE_ACT_NONE) return fail("last layer must be linear");
            if (ly.in_features != width || ly.out_features != spec->n_out) return fail("output layer shape mismatch");
        }
    }
    if (spec->n_first < 0 || spec->n_first > TDE_MAX_FIRST) return fail("n_first %d out of range", spec->n_first);
    if (spec->n_second < 0 || spec->n_second > TDE_MAX_SECOND) return fail("n_second %d out of range", spec->n_second);
    for (int a = 0; a < spec->n_first; ++a)
        if (spec->first_in[a] < 0 || spec->first_in[a] >= spec->d_in) return fail("first_in[%d]=%d is not an input column", a, spec->first_in[a]);
    int smask = 0, prev = -1;
    for (int q = 0; q < spec->n_second; ++q) {
        const int a = spec->second_a[q], b = spec->second_b[q];
        if (a < 0 || b < a || b >= spec->n_first) return fail("second-order pair (%d,%d) must satisfy 0<=a<=b<n_first", a, b);
        const int idx = tde::pair_index(spec->n_first, a, b);
        if (idx <= prev) return fail("second-order pairs must be sorted and unique");
        prev = idx;
        smask |= 1 << idx;
    }
    const Dispatch* d = find_dispatch(spec->n_first, smask);
    if (!d) return fail("jet spec (n_first=%d, pair mask=0x%x) is not compiled in; canonicalise on the host", spec->n_first, smask);

    Geometry g{};
    g.disp = d;
    g.nf = spec->n_first;
    g.smask = smask;
    g.J = 1 + spec->n_first + spec->n_second;
    g.PT = g.J <= 5 ? 4 : 2;
    g.PJ = g.PT * g.J;
    g.width = width;
    g.L = L;
    g.d_in = spec->d_in;
    g.n_out = spec->n_out;
    g.act = act;
    if (fourier && (width & 1)) return fail("a Fourier embedding needs an even hidden width");
    g.fourier_half = fourier ? width / 2 : 0;
    g.W = width <= 256 ? (int)align_up(width, 64) : (int)align_up(width, 256);
    if (g.W > 2048) return fail("hidden width %d too large (max 2048)", width);
    const int NG = g.W / tde::kNT;
    const int pg_max = std::max(1, tde::kComputeThreads / NG);
    const int NI = std::min(g.W / 8, 32);
    g.kblk = 64;
    const int KI = g.kblk / 8;
    g.rsplit = std::max(1, tde::kComputeThreads / (NI * KI));
    bool found = false;
    for (int pg = pg_max; pg >= 1 && !found; --pg) {
        if (pg % g.rsplit != 0) continue;
        const int TP = pg * g.PT;
        const int cols = TP * g.J;
        const int pitch = cols + ((cols % 8 == 4) ? 0 : 4);
        const int64_t common = (int64_t)g.n_out * g.W + (int64_t)TP * g.d_in + (int64_t)g.J * g.n_out * TP;
        for (int sb = tde::kMaxStages; sb >= 2 && !found; --sb) {
            const int64_t ub = std::max<int64_t>({(int64_t)sb * tde::kKC * g.W, (int64_t)g.kblk * pitch, (int64_t)pg * g.n_out * g.W});
            const int64_t smem_b = 128 + 4 * ((int64_t)g.W * pitch + ub + common);
            if (smem_b > kSmemLimit) continue;
            int sf = tde::kMaxStages;
            int64_t smem_f = 0;
            for (; sf >= 2; --sf) {
                smem_f = 128 + 4 * ((int64_t)g.W * pitch + (int64_t)sf * tde::kKC * g.W + common);
                if (smem_f <= kSmemLimit) break;
            }
            if (sf < 2) continue;
            g.PG = pg;
            g.TP = TP;
            g.pitch = pitch;
            g.stages_fwd = sf;
            g.stages_bwd = sb;
            g.smem_fwd = (int)smem_f;
            g.smem_bwd = (int)smem_b;
            found = true;
        }
    }
    if (!found) {
        // rsplit may not divide any feasible PG for exotic shapes: fall back to rsplit = 1
        return fail("no tile configuration fits shared memory for width %d, J %d", width, g.J);
    }
    g.n_tiles = B > 0 ? (B + g.TP - 1) / g.TP : 0;
    g.n_ctas = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), g.n_tiles));

    const int64_t WW = (int64_t)g.W * g.W;
    int64_t o = 0;
    g.o_w_in = o; o = align_up(o + (int64_t)g.W * g.d_in, 4);
    g.o_b_in = o; o = align_up(o + g.W, 4);
    g.o_wt = o;   o += (int64_t)(L - 1) * WW;
    g.o_wn = o;   o += (int64_t)(L - 1) * WW;
    g.o_bh = o;   o += (int64_t)(L - 1) * g.W;
    g.o_wo = o;   o = align_up(o + (int64_t)g.n_out * g.W, 4);
    g.o_bo = o;   o = align_up(o + g.n_out, 4);
    // tcgen05 forward: W in {128, 256}, accumulators of all neuron blocks fit 256 TMEM columns, backward tile is a
    // multiple of the 16-point tensor tile.  TDE_NO_TC=1 forces the FFMA path (A/B measurements, debugging).
    g.use_tc = 0;
    g.tc_pts = tde::tc_points(g.W, g.J);
    {
        const int Ntc = g.tc_pts * g.J;
        const int mb = g.W / 128;
        g.tc_ksc = tde::tc_ksc(g.W, Ntc);
        const bool no_tc = options().no_tc != 0;
        // W = 128: single-CTA kernels, N = 16 J columns per MMA.  W = 256: CTA-pair kernels (cta_group::2), N = 32 J <= 256.
        // TDE_TC_PAIR=1 (W = 256, N = 32 J <= 256): CTA-pair kernels (cta_group::2) instead of the single-CTA ones
        g.use_pair = g.W == 256 && 6 * Ntc <= 512 && options().tc_pair == 1;
        const bool shape_ok = (g.W == 128 || g.W == 256 || g.W == 512) && (g.use_pair || tde::tc_shape_ok(mb, Ntc));
        // (W <= 256 can pair a tensor-core forward with the FFMA reverse sweep: its tile must then be whole tensor tiles)
        if (shape_ok && L >= 2 && (g.W > 256 || g.TP % g.tc_pts == 0) && !no_tc) {
            const int64_t atomk = (Ntc / 8) * 1024;
            const int pts = g.use_pair ? 2 * tde::kTcPoints : g.tc_pts;     // points staged per CTA (x, seeds)
            // [planes][weight stages][w_out][2 x (x | seeds or y, 4 (padded) outputs per channel)][tail: barriers]; the reverse sweep
            // stages the next tile's x / seeds while the current one runs; the pair kernels keep a 1 KB header
            const int64_t staging = 4 * (pts * g.d_in + (int64_t)g.J * 4 * pts);
            // (the pair kernels also keep w_out in shared memory)
            int64_t fixed = (g.use_pair ? 1024 + 4 * ((int64_t)g.n_out * g.W) : tde::kTcTailBytes) + 2 * (g.W / 32) * atomk + staging;
            const int64_t chunk = g.use_pair ? tde::kPairChunkBytes : (int64_t)g.tc_ksc * 8192;
            // the second staging buffer is taken only when it does not cost a weight stage
            g.tc_dbuf = 0;
            if (!g.use_pair && (kSmemLimit - fixed - staging) / chunk == (kSmemLimit - fixed) / chunk && (kSmemLimit - fixed - staging) / chunk >= 2) {
                g.tc_dbuf = 1;
                fixed += staging;
            }
            for (int st = tde::kTcMaxStages; st >= 2; --st)
                if (fixed + st * chunk <= kSmemLimit) {
                    if (options().tc_stages > 0) st = std::max(2, std::min(st, options().tc_stages));   // experiments only
                    g.use_tc = 1; g.tc_mb = mb; g.tc_stages = st; g.tc_smem = (int)(fixed + st * chunk);
                    g.tc_nacc = tde::tc_acc_sets(g.use_pair ? 1 : mb, g.use_pair ? 2 * Ntc : Ntc);
                    break;
                }
        }
        // whole 16-point tiles of the weight-gradient GEMM: with 4- or 8-point sweep tiles the tail is padded with tiles of
        // x = 0 / seeds = 0 (they write finite stash rows and exact-zero Hbar rows, so the GEMM reads nothing uninitialised)
        const int64_t per16 = std::max(1, 16 / g.tc_pts);
        g.tc_tiles = B > 0 ? ((B + g.tc_pts - 1) / g.tc_pts + per16 - 1) / per16 * per16 : 0;
    }
    g.dw_split = g.W > 256 ? 2 : 1;
    g.dw_w = g.W / g.dw_split;
    g.dw_smem = 1024 + (tde::kDwStages * 4 + tde::dw_raw_stages(g.J) * 2) * g.dw_w * 32;
    // W = 512: the forward and the reverse sweep switch together (the FFMA sweep is not paired with the 8-point-tile forward)
    if (g.W > 256 && (g.J > 6 || g.dw_smem > kSmemLimit || options().no_tc_bwd)) g.use_tc = 0;
    o = align_up(o, 64);
    g.o_wc = o;   if (g.use_tc) o += (int64_t)(L - 1) * 2 * WW;
    g.use_tc_bwd = 0;
    {
        if (g.use_tc && g.J <= 6 && g.dw_smem <= kSmemLimit && !options().no_tc_bwd) {
            g.use_tc_bwd = 1;
            g.rsplit = 1;      // the FFMA dW phase (and its column split) is not used on this path
        }
    }
    g.o_wct = o;  if (g.use_tc_bwd) o += (int64_t)(L - 1) * 2 * WW;
    g.packed_floats = align_up(o, 64);

    int64_t q = 0;
    q = align_up((int64_t)g.W * g.d_in, 4);
    g.g_b_in = q;   q = align_up(q + g.W, 4);
    // hidden dW: [layer][n][k]; the W = 512 tensor path keeps one 256 x 256 quadrant per CTA
    g.g_hidden = q; q += (int64_t)(L - 1) * g.rsplit * (g.use_tc_bwd ? (int64_t)g.dw_w * g.dw_w : WW);
    g.g_bh = q;     q += (int64_t)(L - 1) * g.W;
    g.g_wo = q;     q = align_up(q + (int64_t)g.n_out * g.W, 4);
    g.g_bo = q;     q = align_up(q + g.n_out, 4);
    g.n_copies = 1;
    for (int k = 0; k < 4; ++k) g.g_dup_lo[k] = g.g_dup_hi[k] = 0;
    if (g.use_tc_bwd) {
        // the tcgen05 reverse sweep has several threads per neuron (one per group of 8 points: 2 for W = 128, 4 for the CTA
        // pair): each adds into its own copy of (w_in | b_in) and (b_hidden | w_out); the reduction sums the copies in a
        // fixed order
        g.n_copies = g.use_pair ? 4 : g.tc_pts / tde::tc_visit_points(g.W, g.tc_pts);
        const int64_t small_end = q;
        const int64_t dup_size = align_up(g.g_hidden + (small_end - g.g_bh), 4);
        for (int k = 1; k < g.n_copies; ++k) {
            g.g_dup_lo[k] = q;                                  // copy of [0, g_hidden)
            g.g_dup_hi[k] = q + g.g_hidden - g.g_bh;            // copy of [g_bh, small_end) starts at q + g_hidden
            q += dup_size;
        }
    }
    g.gp_stride = q;
    g.o_gpart = g.packed_floats;
    // The tcgen05 kernels walk 16-point tiles in cluster pairs: when the tile count is odd the last pair's second tile is
    // padding (x = 0, seeds = 0) and needs real stash rows behind it.
    g.stash_layer_stride = g.n_tiles * g.W * (int64_t)g.TP * g.J;
    if (g.use_tc) {
        const int cl = g.use_pair ? tde::kPairCluster : tde::kTcCluster;
        const int64_t padded = (g.tc_tiles + cl - 1) / cl * cl;
        g.stash_layer_stride = std::max<int64_t>(g.stash_layer_stride, padded * (g.use_pair ? tde::kTcPoints : g.tc_pts) * (int64_t)g.J * g.W);
    }
    g.stash_floats = (int64_t)L * g.stash_layer_stride;
    g.stash_total_floats = g.stash_floats;
    g.workspace_floats = align_up(g.packed_floats + (int64_t)std::max(g.n_ctas, sm_count()) * g.gp_stride, 64);
    g.o_hbar = 0;
    if (g.use_tc_bwd) {       // Hbar_l and Z_l streams for the tensor-core weight-gradient kernel
        // Hbar_l stream: written by the reverse sweep, read by the weight-gradient GEMM of the same call
        g.o_hbar = g.workspace_floats;
        g.workspace_floats = g.o_hbar + g.stash_floats;
        const int cl = g.use_pair ? tde::kPairCluster : tde::kTcCluster;
        const int64_t groups = (g.tc_tiles + cl - 1) / cl;
        g.grid_b1 = cl * std::max<int64_t>(1, std::min<int64_t>(tc_cluster_slots(g.disp, g.use_pair != 0, g.tc_smem), groups));
        // weight-gradient GEMM walks 16-point tiles (two 8-point tiles of the W = 512 sweep); one CTA per SM, a multiple of the
        // quadrant count, never more tile groups than tiles
        const int64_t dw_tiles = g.tc_tiles * g.tc_pts / 16 + ((g.tc_tiles * g.tc_pts) % 16 ? 1 : 0);
        const int nq = g.dw_split * g.dw_split;
        g.grid_b2 = nq * std::max<int64_t>(1, std::min<int64_t>(sm_count() / nq, dw_tiles));
    }

    int64_t np = 0;
    for (int l = 0; l < net->n_layers; ++l)
        np += (int64_t)net->layers[l].in_features * net->layers[l].out_features + net->layers[l].out_features;
    g.n_params = np;
    *G = g;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing: pad to W, transpose / permute columns for the register-tile GEMMs
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    TdeNet net;
    int W, L, d_in, n_out, width;
    int64_t o_w_in, o_b_in, o_wt, o_wn, o_bh, o_wo, o_bo, total;
    int want_wt, want_wn;
};

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a, float* __restrict__ ws) {
    const int W = a.W;
    const int64_t WW = (int64_t)W * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < a.o_b_in) {  // w_in [W][d_in]
            const int64_t r = i - a.o_w_in;
            if (r < (int64_t)W * a.d_in) {
                const int n = (int)(r / a.d_in), c = (int)(r % a.d_in);
                if (n < a.width) v = a.net.layers[0].weight[(int64_t)n * a.d_in + c];
            }
        } else if (i < a.o_wt) {  // b_in
            const int n = (int)(i - a.o_b_in);
            if (n < a.width) v = a.net.layers[0].bias[n];
        } else if (i < a.o_wn) {  // Wt[l][k][q] = W_l[n(q)][k]
            if (!a.want_wt) continue;
            const int64_t r = i - a.o_wt;
            const int l = (int)(r / WW) + 1;
            const int k = (int)((r % WW) / W), q = (int)(r % W);
            const int n = tde::neuron_of_q(q, W);
            if (n < a.width && k < a.width) v = a.net.layers[l].weight[(int64_t)n * a.width + k];
        } else if (i < a.o_bh) {  // Wn[l][n][q] = W_l[n][k(q)]
            if (!a.want_wn) continue;
            const int64_t r = i - a.o_wn;
            const int l = (int)(r / WW) + 1;
            const int n = (int)((r % WW) / W), q = (int)(r % W);
            const int k = tde::neuron_of_q(q, W);
            if (n < a.width && k < a.width) v = a.net.layers[l].weight[(int64_t)n * a.width + k];
        } else if (i < a.o_wo) {  // b_hidden
            const int64_t r = i - a.o_bh;
            const int l = (int)(r / W) + 1, n = (int)(r % W);
            if (n < a.width) v = a.net.layers[l].bias[n];
        } else if (i < a.o_bo) {  // w_out [n_out][W]
            const int64_t r = i - a.o_wo;
            if (r < (int64_t)a.n_out * W) {
                const int o = (int)(r / W), k = (int)(r % W);
                if (k < a.width) v = a.net.layers[a.L].weight[(int64_t)o * a.width + k];
            }
        } else {
            const int o = (int)(i - a.o_bo);
            if (o < a.n_out) v = a.net.layers[a.L].bias[o];
        }
        ws[i] = v;
    }
}

// tcgen05 operand image of the hidden weights: per layer, per 8-k chunk: [hi|lo][neuron block][k group of 4][16 row
// groups][8 rows][4 k] -- exactly the shared-memory layout of the K-major SWIZZLE_NONE UMMA A operand, split 3xTF32.
__global__ void pack_tc_weights_kernel(const __grid_constant__ PackArgs a, float* __restrict__ wc, int MB, int KSC, int transpose) {
    const int W = a.W;
    const int64_t per_layer = 2LL * W * W;
    const int64_t total = (int64_t)(a.L - 1) * per_layer;
    const int chunk = KSC * 2048;         // floats per chunk: one neuron block, KSC k-steps of [hi 1024 | lo 1024]
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(i / per_layer) + 1;
        int64_t r = i % per_layer;
        const int ci = (int)(r / chunk);
        int c = (int)(r % chunk);
        int hl, mb, ks, kg;
        if (MB == 1) {          // one chunk per k-step: [hi|lo][k group 2][16 row groups][8][4]
            ks = ci;
            mb = 0;
            hl = c / 1024;      c %= 1024;
        } else {                // MB >= 2: chunks follow the issue order of tc_issue_layer(_wide): units (k range of 128 inputs kq,
                                // neuron block mb), kq-major, each 16 k-steps = 16 / KSC chunks
            const int per_q = 16 / KSC;        // chunks per unit
            const int q = ci / per_q, cc = ci % per_q;
            mb = q % MB;
            const int s2 = c / 2048; c %= 2048;
            ks = (q / MB) * 16 + cc * KSC + s2;
            hl = c / 1024;      c %= 1024;
        }
        kg = c / 512;           c %= 512;
        const int rg = c / 32;  c %= 32;
        const int rr = c / 4, kk = c % 4;
        const int n = mb * 128 + rg * 8 + rr, k = ks * 8 + kg * 4 + kk;
        float v = 0.0f;
        if (n < a.width && k < a.width)
            v = transpose ? a.net.layers[l].weight[(int64_t)k * a.width + n] : a.net.layers[l].weight[(int64_t)n * a.width + k];
        float hi, lo;
        tde::split_tf32(v, hi, lo);
        wc[i] = hl ? lo : hi;
    }
}

int launch_pack(const TdeNet* net, const Geometry& g, float* ws, bool want_wt, bool want_wn, cudaStream_t st) {
    PackArgs a{};
    a.net = *net;
    a.W = g.W; a.L = g.L; a.d_in = g.d_in; a.n_out = g.n_out; a.width = g.width;
    a.o_w_in = g.o_w_in; a.o_b_in = g.o_b_in; a.o_wt = g.o_wt; a.o_wn = g.o_wn; a.o_bh = g.o_bh;
    a.o_wo = g.o_wo; a.o_bo = g.o_bo; a.total = g.o_bo + g.n_out;
    a.want_wt = want_wt; a.want_wn = want_wn;
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((g.packed_floats + threads - 1) / threads, 148 * 8);
    pack_weights_kernel<<<blocks, threads, 0, st>>>(a, ws);
    ++g_launches;
    TDE_CUDA(cudaGetLastError());
    if (want_wt && g.use_tc) {
        pack_tc_weights_kernel<<<148 * 4, threads, 0, st>>>(a, ws + g.o_wc, g.tc_mb, g.tc_ksc, 0);
        ++g_launches;
        TDE_CUDA(cudaGetLastError());
    }
    if (want_wn && g.use_tc_bwd) {
        pack_tc_weights_kernel<<<148 * 4, threads, 0, st>>>(a, ws + g.o_wct, g.tc_mb, g.tc_ksc, 1);
        ++g_launches;
        TDE_CUDA(cudaGetLastError());
    }
    return 0;
}

void fill_params(const Geometry& g, const TdeJetSpec* spec, float* ws, tde::JetParams* p) {
    memset(p, 0, sizeof *p);
    p->d_in = g.d_in; p->n_out = g.n_out; p->L = g.L; p->W = g.W; p->act = g.act; p->fourier_half = g.fourier_half;
    p->PG = g.PG; p->pitch = g.pitch; p->n_tiles = g.n_tiles; p->stash_layer_stride = g.stash_layer_stride;
    for (int a = 0; a < TDE_MAX_FIRST; ++a) p->first_in[a] = a < spec->n_first ? spec->first_in[a] : 0;
    p->w_in = ws + g.o_w_in; p->b_in = ws + g.o_b_in; p->wt_hidden = ws + g.o_wt; p->wn_hidden = ws + g.o_wn;
    p->b_hidden = ws + g.o_bh; p->w_out = ws + g.o_wo; p->b_out = ws + g.o_bo;
    p->gpart = ws + g.o_gpart; p->gp_stride = g.gp_stride;
    p->gp_off_b_in = g.g_b_in; p->gp_off_hidden = g.g_hidden; p->gp_off_b_hidden = g.g_bh;
    p->gp_off_w_out = g.g_wo; p->gp_off_b_out = g.g_bo;
    p->kblk = g.kblk; p->rsplit = g.rsplit;
}

// ---------------------------------------------------------------------------------------------
// second-stage reduction: flat_grad[i] += scale * sum_{cta (, split)} slab[cta][map(i)]   (fixed order)
// ---------------------------------------------------------------------------------------------
struct ReduceArgs {
    int n_layers, W, width, d_in, n_out, rsplit, n_active, n_active_hidden;
    int64_t n_params, gp_stride;
    int64_t g_b_in, g_hidden, g_bh, g_wo, g_bo;
    int n_copies;                // copies of the small sections (tcgen05 reverse sweep: one per thread of a neuron)
    int64_t dup_lo[4], dup_hi[4];
    int dw_w, dw_split;          // tensor-core weight-gradient GEMM: hidden dW kept as dw_split^2 quadrants of dw_w x dw_w, slab c holds
                                 // quadrant c % dw_split^2 (dw_split = 1: the whole matrix in every slab)
    float scale;
};

__global__ void reduce_grads_kernel(const __grid_constant__ ReduceArgs a, const float* __restrict__ gpart,
                                    float* __restrict__ flat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= a.n_params) return;
    // decode flat index -> (layer, weight|bias, row, col)
    int64_t r = i;
    int64_t src = -1;
    int splits = 1, slab0 = 0, slab_step = 1;
    const int L = a.n_layers - 1;
    const int64_t WW = (int64_t)a.W * a.W;
    // layer 0
    const int64_t n0 = (int64_t)a.width * a.d_in;
    if (r < n0) {
        src = (r / a.d_in) * a.d_in + (r % a.d_in);  // slab w_in is [W][d_in] with the same row pitch
    } else if ((r -= n0) < a.width) {
        src = a.g_b_in + r;
    } else {
        r -= a.width;
        const int64_t per = (int64_t)a.width * a.width + a.width;
        const int64_t hid = (int64_t)(L - 1) * per;
        if (r < hid) {
            const int l = (int)(r / per);
            const int64_t rr = r % per;
            if (rr < (int64_t)a.width * a.width) {
                const int n = (int)(rr / a.width), k = (int)(rr % a.width);
                if (a.dw_split > 1) {
                    slab0 = (n / a.dw_w) * a.dw_split + (k / a.dw_w);
                    slab_step = a.dw_split * a.dw_split;
                    src = a.g_hidden + (int64_t)l * a.dw_w * a.dw_w + (int64_t)(n % a.dw_w) * a.dw_w + (k % a.dw_w);
                } else {
                    src = a.g_hidden + (int64_t)l * a.rsplit * WW + (int64_t)n * a.W + k;
                }
                splits = a.rsplit;
            } else {
                src = a.g_bh + (int64_t)l * a.W + (rr - (int64_t)a.width * a.width);
            }
        } else {
            r -= hid;
            const int64_t nwo = (int64_t)a.n_out * a.width;
            if (r < nwo) {
                src = a.g_wo + (r / a.width) * a.W + (r % a.width);
            } else {
                src = a.g_bo + (r - nwo);
            }
        }
    }
    float s = 0.0f;
    const int n_slabs = splits > 1 || src >= a.g_hidden && src < a.g_bh ? a.n_active_hidden / slab_step : a.n_active;
    const int which = src < a.g_hidden ? 1 : (src >= a.g_bh && src < a.g_bo ? 2 : 0);
    if (splits == 1 && !(which && a.n_copies > 1)) {
        // common case (hidden weights on the tensor path): one value per slab.  Eight loads are in flight at a time; the
        // additions stay in slab order, so the result does not depend on the batching
        int c = 0;
        for (; c + 8 <= n_slabs; c += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(gpart + (int64_t)(slab0 + (c + u) * slab_step) * a.gp_stride + src);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; c < n_slabs; ++c) s += __ldcs(gpart + (int64_t)(slab0 + c * slab_step) * a.gp_stride + src);
    } else {
        for (int c = 0; c < n_slabs; ++c) {
            const float* slab = gpart + (int64_t)c * a.gp_stride + src;
            for (int q = 0; q < splits; ++q) s += slab[(int64_t)q * WW];
            if (which)
                for (int k = 1; k < a.n_copies; ++k) s += slab[which == 1 ? a.dup_lo[k] : a.dup_hi[k]];
        }
    }
    flat[i] += a.scale * s;
}

// bilinear grid lookup pieces (GridInterpolator.py:50-83), shared by tde_grid_interp2d and the loss kernel's GRID2D op
template <bool SMEM>
__device__ __forceinline__ float grid_at(const float* g, const int i) { return SMEM ? g[i] : __ldg(g + i); }
template <bool SMEM>
__device__ __forceinline__ void grid_bracket(const float* g, const int n, const float x, int& il, int& ir, float& dl, float& dr) {
    int lo = 0, hi = n;                       // first index with g[i] >= x  (torch.bucketize, right = false)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (grid_at<SMEM>(g, mid) < x) lo = mid + 1; else hi = mid;
    }
    ir = lo < n ? lo : n - 1;
    il = ir - 1 < 0 ? 0 : ir - 1;
    dl = x - grid_at<SMEM>(g, il);
    dr = grid_at<SMEM>(g, ir) - x;
    if (dl < 0.0f) dl = 0.0f;
    if (dr < 0.0f) dr = 0.0f;
    if (dl == 0.0f && dr == 0.0f) dl = dr = 1.0f;
}
// one lookup, corner order and weights as in the reference: value[left/right] * distance to the OTHER side
template <bool SMEM>
__device__ __forceinline__ float grid_lookup(const float* sx, const int nx, const float* sy, const int ny,
                                             const float* __restrict__ values, const float x, const float y) {
    int xl, xr, yl, yr;
    float dxl, dxr, dyl, dyr;
    grid_bracket<SMEM>(sx, nx, x, xl, xr, dxl, dxr);
    grid_bracket<SMEM>(sy, ny, y, yl, yr, dyl, dyr);
    float num = 0.0f;
    num += __ldg(values + (int64_t)xl * ny + yl) * (dxr * dyr);
    num += __ldg(values + (int64_t)xl * ny + yr) * (dxr * dyl);
    num += __ldg(values + (int64_t)xr * ny + yl) * (dxl * dyr);
    num += __ldg(values + (int64_t)xr * ny + yr) * (dxl * dyl);
    return num / ((dxl + dxr) * (dyl + dyr));
}

// ---------------------------------------------------------------------------------------------
// fused residual + loss: register-program interpreter, forward then reverse sweep per point
// ---------------------------------------------------------------------------------------------
constexpr int kLossBlocks = 592;  // 148 SMs x 4
constexpr int kLossThreads = 256;

// cross-block scratch of residual_loss_kernel: a ring of slots handed out round-robin per launch, so launches that are in
// flight at the same time on different streams never share partials or a ticket
constexpr int kLossSlots = 32;
__device__ float g_loss_partials[kLossSlots][kLossBlocks * TDE_MAX_ROWS];
__device__ unsigned int g_loss_ticket[kLossSlots];
unsigned long long g_loss_seq = 0;

__global__ void __launch_bounds__(kLossThreads) residual_loss_kernel(
    const __grid_constant__ TdeLossProgram prog, const float* __restrict__ streams, int64_t ld, int n_stream_rows,
    const float* __restrict__ aux, int64_t ld_aux, int64_t p0, int64_t n, float inv_count,
    const float* __restrict__ coef_dev, float* __restrict__ row_sums, float* __restrict__ seeds,
    float* __restrict__ point_res, float* __restrict__ point_w, const int slot) {
    __shared__ float red[TDE_MAX_ROWS][kLossThreads / 32];
    __shared__ bool is_last;
    float part[TDE_MAX_ROWS], coef[TDE_MAX_ROWS];
#pragma unroll
    for (int r = 0; r < TDE_MAX_ROWS; ++r) {
        part[r] = 0.0f;
        coef[r] = r < prog.n_rows ? prog.coef[r] * (coef_dev ? coef_dev[r] : 1.0f) : 0.0f;
    }

    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + i;
        float v[TDE_MAX_REGS], adj[TDE_MAX_REGS];
        for (int k = 0; k < prog.n_ops; ++k) {
            const TdeOp op = prog.ops[k];
            float r;
            switch (op.op) {
                case TDE_OP_CONST: r = op.imm; break;
                case TDE_OP_LOAD: r = streams[(int64_t)op.a * ld + p]; break;
                case TDE_OP_AUX: r = aux[(int64_t)op.a * ld_aux + i]; break;
                case TDE_OP_ADD: r = v[op.a] + v[op.b]; break;
                case TDE_OP_SUB: r = v[op.a] - v[op.b]; break;
                case TDE_OP_MUL: r = v[op.a] * v[op.b]; break;
                case TDE_OP_NEG: r = -v[op.a]; break;
                case TDE_OP_POWI: {
                    const int e = (int)op.imm;
                    float b = v[op.a];
                    r = 1.0f;
                    for (int t = 0; t < (e < 0 ? -e : e); ++t) r *= b;
                    if (e < 0) r = 1.0f / r;
                } break;
                case TDE_OP_ABS: r = fabsf(v[op.a]); break;
                case TDE_OP_SIN: r = sinf(v[op.a]); break;
                case TDE_OP_COS: r = cosf(v[op.a]); break;
                case TDE_OP_DIV: r = v[op.a] / v[op.b]; break;
                case TDE_OP_GRID2D: {
                    const TdeGrid2D& gd = prog.grids[(int)op.imm];
                    r = grid_lookup<false>(gd.gx, gd.nx, gd.gy, gd.ny, gd.values, v[op.a], v[op.b]);
                } break;
                default: r = 0.0f;
            }
            v[op.dst] = r;
            adj[op.dst] = 0.0f;
        }
        for (int r = 0; r < prog.n_rows; ++r) {
            const float res = v[prog.res_reg[r]];
            const float w = prog.w_reg[r] >= 0 ? v[prog.w_reg[r]]
                            : (prog.w_aux[r] >= 0 ? aux[(int64_t)prog.w_aux[r] * ld_aux + i] : prog.w_const[r]);
            part[r] += w * res * res;
            adj[prog.res_reg[r]] += coef[r] * inv_count * 2.0f * w * res;
            if (point_res) point_res[(int64_t)r * n + i] = res;
            if (point_w) point_w[(int64_t)r * n + i] = w;
        }
        if (seeds) {
            for (int s = 0; s < n_stream_rows; ++s) seeds[(int64_t)s * ld + p] = 0.0f;
            for (int k = prog.n_ops - 1; k >= 0; --k) {
                const TdeOp op = prog.ops[k];
                const float d = adj[op.dst];
                switch (op.op) {
                    case TDE_OP_LOAD: seeds[(int64_t)op.a * ld + p] += d; break;
                    case TDE_OP_ADD: adj[op.a] += d; adj[op.b] += d; break;
                    case TDE_OP_SUB: adj[op.a] += d; adj[op.b] -= d; break;
                    case TDE_OP_MUL: {
                        const float va = v[op.a], vb = v[op.b];
                        adj[op.a] += d * vb;
                        adj[op.b] += d * va;
                    } break;
                    case TDE_OP_NEG: adj[op.a] -= d; break;
                    case TDE_OP_POWI: {
                        const int e = (int)op.imm;
                        if (e == 0) break;               // d(b^0) = 0 (b^(e-1) would be 1/b: 0 * inf at b = 0)
                        const float b = v[op.a];
                        float pw = 1.0f;  // b^(e-1)
                        for (int t = 0; t < (e - 1 < 0 ? 1 - e : e - 1); ++t) pw *= b;
                        if (e - 1 < 0) pw = 1.0f / pw;
                        adj[op.a] += d * (float)e * pw;
                    } break;
                    case TDE_OP_ABS: adj[op.a] += v[op.a] >= 0.0f ? d : -d; break;
                    case TDE_OP_SIN: adj[op.a] += d * cosf(v[op.a]); break;
                    case TDE_OP_COS: adj[op.a] -= d * sinf(v[op.a]); break;
                    case TDE_OP_DIV: {
                        const float inv = 1.0f / v[op.b];
                        adj[op.a] += d * inv;
                        adj[op.b] -= d * v[op.dst] * inv;
                    } break;
                    default: break;
                }
            }
        }
    }
    // block reduction (fixed order), then last-block-done cross-block reduction (fixed order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < TDE_MAX_ROWS; ++r) {
        float s = part[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[r][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < TDE_MAX_ROWS) {
        float s = 0.0f;
        for (int w = 0; w < kLossThreads / 32; ++w) s += red[threadIdx.x][w];
        g_loss_partials[slot][blockIdx.x * TDE_MAX_ROWS + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&g_loss_ticket[slot], 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x < prog.n_rows && row_sums) {
            float s = 0.0f;
            for (unsigned b = 0; b < gridDim.x; ++b) s += g_loss_partials[slot][b * TDE_MAX_ROWS + threadIdx.x];
            row_sums[threadIdx.x] += s * inv_count;
        }
        if (threadIdx.x == 0) g_loss_ticket[slot] = 0;
    }
}

// torch.optim.Adam's update, operation for operation (torch/optim/adam.py, _multi_tensor_adam with python-float scalars):
//   exp_avg.lerp_(grad, 1 - beta1)                       -> m + w1 * (g - m)          (lerp with weight < 0.5)
//   exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2) -> v * beta2 + (w2 * g) * g
//   denom = exp_avg_sq.sqrt() / sqrt(bias_correction2) + eps;  param.addcdiv_(exp_avg, denom, value = -lr / bias_correction1)
// The scalars are computed in double on the host and rounded to float once, as torch does.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float w1, float b2, float w2, float eps,
                            float neg_step_size, float bc2_sqrt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float m0 = m[i];
        const float mi = fmaf(w1, gi - m0, m0);
        const float vi = fmaf(w2 * gi, gi, v[i] * b2);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = fmaf(neg_step_size, mi / denom, p[i]);
    }
}

__global__ void __launch_bounds__(512) ffma_probe_kernel(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// bilinear grid lookup (GridInterpolator.py:50-83), one thread per query point
// ---------------------------------------------------------------------------------------------
namespace {
// SMEM: the two coordinate arrays are staged in shared memory (the binary searches are 2 x log2(n) dependent loads per point)
template <bool SMEM>
__global__ void grid_interp2d_kernel(const float* __restrict__ gx, const int nx, const float* __restrict__ gy, const int ny,
                                     const float* __restrict__ values, const float* __restrict__ x, const float* __restrict__ y,
                                     const int64_t n, float* __restrict__ out) {
    extern __shared__ float grid_s[];
    const float* sx = gx;
    const float* sy = gy;
    if (SMEM) {
        for (int i = threadIdx.x; i < nx; i += blockDim.x) grid_s[i] = gx[i];
        for (int i = threadIdx.x; i < ny; i += blockDim.x) grid_s[nx + i] = gy[i];
        __syncthreads();
        sx = grid_s;
        sy = grid_s + nx;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __stcs(out + i, grid_lookup<SMEM>(sx, nx, sy, ny, values, __ldcs(x + i), __ldcs(y + i)));
    }
}
}  // namespace

extern "C" {

const char* tde_last_error(void) { return g_err.c_str(); }
int tde_version(void) { return 101; }

int tde_abi_struct_bytes(int which) {
    switch (which) {
        case 0: return (int)sizeof(TdeNet);
        case 1: return (int)sizeof(TdeJetSpec);
        case 2: return (int)sizeof(TdePlan);
        case 3: return (int)sizeof(TdeLossProgram);
        default: return -1;
    }
}

int tde_plan(const TdeNet* net, const TdeJetSpec* spec, int64_t n_points, TdePlan* plan) {
    if (!plan) return fail("null plan");
    Geometry g;
    if (int rc = make_geometry(net, spec, n_points, &g)) return rc;
    plan->n_channels = g.J;
    plan->width_padded = g.W;
    plan->tile_points = g.TP;
    plan->n_ctas = g.n_ctas;
    plan->n_tiles = g.n_tiles;
    plan->stash_bytes = g.stash_total_floats * 4;
    plan->workspace_bytes = g.workspace_floats * 4;
    plan->workspace_forward_bytes = g.packed_floats * 4;
    plan->n_params = g.n_params;
    plan->smem_forward = g.smem_fwd;
    plan->smem_backward = g.smem_bwd;
    return 0;
}

int tde_plan_paths(const TdeNet* net, const TdeJetSpec* spec, int64_t n_points) {
    Geometry g;
    if (make_geometry(net, spec, n_points, &g)) return -1;
    return (g.use_tc ? 1 : 0) | (g.use_tc_bwd ? 2 : 0);
}

int tde_jet_forward(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points, float* out,
                    int64_t ld_out, float* stash, void* workspace, void* stream) {
    Geometry g;
    if (int rc = make_geometry(net, spec, n_points, &g)) return rc;
    if (n_points == 0) return 0;
    if (!x || !out || !workspace) return fail("null x/out/workspace");
    if (ld_out < n_points) return fail("ld_out %lld < n_points %lld", (long long)ld_out, (long long)n_points);
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace;
    if (int rc = launch_pack(net, g, ws, true, false, st)) return rc;
    if (g.use_tc) {
        tde::TcParams t{};
        t.x = x; t.B = n_points; t.n_tiles = g.tc_tiles;
        t.d_in = g.d_in; t.n_out = g.n_out; t.L = g.L; t.W = g.W; t.act = g.act; t.MB = g.tc_mb; t.n_stages = g.tc_stages; t.n_acc = g.tc_nacc; t.fourier_half = g.fourier_half; t.pair = g.use_pair;
        t.dbg = options().tc_dbg;
        static long long* dbg_buf = nullptr;
        if (t.dbg & 4) {
            if (!dbg_buf) { cudaMalloc(&dbg_buf, 64 * sizeof(long long)); cudaMemset(dbg_buf, 0, 64 * sizeof(long long)); }
            t.dbg_buf = dbg_buf;
        }
        for (int a = 0; a < TDE_MAX_FIRST; ++a) t.first_in[a] = a < spec->n_first ? spec->first_in[a] : 0;
        t.w_in = ws + g.o_w_in; t.b_in = ws + g.o_b_in; t.wc = ws + g.o_wc; t.b_hidden = ws + g.o_bh;
        t.w_out = ws + g.o_wo; t.b_out = ws + g.o_bo;
        t.out = out; t.ld_out = ld_out; t.stash = stash;
        t.stash_layer_stride = g.stash_layer_stride;
        t.tiles_per_bw_tile = g.TP / tde::kTcPoints;
        // clusters of kTcCluster CTAs: even grid, never more clusters than tile pairs
        const int cl = g.use_pair ? tde::kPairCluster : tde::kTcCluster;
        const int64_t groups = (g.tc_tiles + cl - 1) / cl;
        const int grid = (int)(cl * std::max<int64_t>(1, std::min<int64_t>(tc_cluster_slots(g.disp, g.use_pair != 0, g.tc_smem), groups)));
        {
            ScopedTimer timer(st, 0);
            TDE_CUDA(g.disp->fwd_tc(t, grid, g.tc_smem, st));
        }
        if (t.dbg & 4) {
            long long h[64];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, dbg_buf, sizeof h, cudaMemcpyDeviceToHost);
            // 4th tile of CTA 0, layers 2..4, neuron blocks 0/1: [body start, accumulators ready, body end, published], relative
            for (int l = 0; l < 3; ++l)
                for (int b = 0; b < 2; ++b) {
                    const long long* d = h + 16 + (l * 2 + b) * 4;
                    fprintf(stderr, "[tc dbg] fwd layer %d block %d: start %lld  acc_ready %lld  body_end %lld  published %lld\n", l + 2, b,
                            d[0] - h[16], d[1] - h[16], d[2] - h[16], d[3] - h[16]);
                }
            fprintf(stderr, "[tc dbg] fwd tile: inputs staged %lld  last epilogue done %lld  output layer %lld  tile done %lld | layer-2 start %lld\n",
                    h[41] - h[40], h[42] - h[40], h[43] - h[40], h[44] - h[40], h[16] - h[40]);
        }
    } else {
        tde::JetParams p;
        fill_params(g, spec, ws, &p);
        p.x = x; p.B = n_points; p.out = out; p.ld_out = ld_out; p.stash = stash; p.n_stages = g.stages_fwd;
        ScopedTimer timer(st, 0);
        TDE_CUDA(g.disp->fwd(p, g.n_ctas, g.smem_fwd, st));
    }
    ++g_launches;
    return 0;
}

// Reverse sweep + weight gradient + second-stage reduction over the points the buffers start at.  `g` is the geometry of the
// launch (tile counts, grids, workspace offsets); its stash_layer_stride is the stride of the stash the pointers index into
// (a range call points into the stash of a larger forward).
static int backward_core(const TdeNet* net, const TdeJetSpec* spec, const Geometry& g, const float* x, int64_t n_points,
                         const float* stash, const float* seeds, int64_t ld_seeds, float grad_scale, float* flat_grad, float* ws,
                         cudaStream_t st) {
    if (int rc = launch_pack(net, g, ws, false, true, st)) return rc;
    int n_active = g.n_ctas, n_active_hidden = g.n_ctas;
    if (g.use_tc_bwd) {
        tde::TcBwdParams t{};
        t.x = x; t.B = n_points; t.n_tiles = g.tc_tiles;
        t.d_in = g.d_in; t.n_out = g.n_out; t.L = g.L; t.W = g.W; t.act = g.act; t.MB = g.tc_mb;
        t.n_stages = g.tc_stages; t.n_acc = g.tc_nacc; t.fourier_half = g.fourier_half; t.pair = g.use_pair;
        for (int a = 0; a < TDE_MAX_FIRST; ++a) t.first_in[a] = a < spec->n_first ? spec->first_in[a] : 0;
        t.wct = ws + g.o_wct; t.w_out = ws + g.o_wo;
        t.seeds = seeds; t.ld_seeds = ld_seeds; t.stash = stash;
        t.stage_dbuf = g.tc_dbuf;
        t.dbg = options().tc_dbg;
        t.hbar = ws + g.o_hbar;
        t.stash_layer_stride = g.stash_layer_stride;
        t.tiles_per_bw_tile = g.TP / tde::kTcPoints;
        t.gpart = ws + g.o_gpart; t.gp_stride = g.gp_stride;
        t.gp_off_b_in = g.g_b_in; t.gp_off_b_hidden = g.g_bh; t.gp_off_w_out = g.g_wo; t.gp_off_b_out = g.g_bo;
        for (int k = 0; k < 4; ++k) { t.gp_dup_lo[k] = g.g_dup_lo[k]; t.gp_dup_hi[k] = g.g_dup_hi[k]; }
        tde::TcDwParams d{};
        d.n_tiles = (g.tc_tiles * g.tc_pts + 15) / 16;     // the weight-gradient GEMM walks 16-point tiles whatever the sweep's tile
        d.L = g.L; d.W = g.dw_w; d.MB = g.dw_w / 128; d.J = g.J; d.Wr = g.W; d.n_split = g.dw_split;
        d.hbar = t.hbar; d.stash = stash; d.act = g.act; d.fourier_half = g.fourier_half;
        d.stash_layer_stride = t.stash_layer_stride; d.tiles_per_bw_tile = t.tiles_per_bw_tile;
        d.gpart = t.gpart; d.gp_stride = g.gp_stride; d.gp_off_hidden = g.g_hidden;
        static long long* bdbg = nullptr;
        if (options().tc_dbg & 4) {
            if (!bdbg) { cudaMalloc(&bdbg, 80 * sizeof(long long)); cudaMemset(bdbg, 0, 80 * sizeof(long long)); }
            t.dbg_buf = bdbg;
        }
        {
            ScopedTimer timer(st, 1);
            TDE_CUDA(g.disp->bwd_tc(t, (int)g.grid_b1, g.tc_smem, st));
        }
        if (t.dbg_buf) {
            long long h[80];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, bdbg, sizeof h, cudaMemcpyDeviceToHost);
            for (int l = 0; l < 3; ++l)
                for (int b = 0; b < 2; ++b) {
                    const long long* d = h + 16 + (l * 2 + b) * 4;
                    fprintf(stderr, "[tc dbg] bwd layer %d block %d: start %lld  acc_ready %lld  body_end %lld  published %lld\n", 5 - l, b,
                            d[0] - h[16], d[1] - h[16], d[2] - h[16], d[3] - h[16]);
                }
            for (int k = 0; k < 2; ++k)
                for (int b = 0; b < 2; ++b) {
                    const long long* d = h + (k ? 58 : 50) + b * 4;
                    fprintf(stderr, "[tc dbg] bwd %s block %d: start %lld  acc_ready %lld  body_end %lld  published %lld (from tile start)\n",
                            k ? "input layer" : "last hidden layer", b, d[0] - h[40], d[1] - h[40], d[2] - h[40], d[3] - h[40]);
                }
            fprintf(stderr, "[tc dbg] bwd tile: inputs staged %lld  last-layer adjoint %lld  layer-5 start %lld | layer 2 done %lld  layer 1 done %lld  input layer done %lld  tile done %lld\n",
                    h[41] - h[40], h[45] - h[40], h[16] - h[40], h[47] - h[40], h[46] - h[40], h[48] - h[40], h[44] - h[40]);
        }
        {
            ScopedTimer timer(st, 2);
            if (g.L >= 2) {
                TDE_CUDA(g.disp->dw_tc(d, (int)g.grid_b2, g.dw_smem, st));
            }
        }
        g_launches += 3;
        n_active = g.use_pair ? (int)g.grid_b1 / 2 : (int)g.grid_b1;     // a CTA pair shares one slab
        n_active_hidden = (int)g.grid_b2;
    } else {
        tde::JetParams p;
        fill_params(g, spec, ws, &p);
        p.x = x; p.B = n_points; p.stash = const_cast<float*>(stash); p.seeds = seeds; p.ld_seeds = ld_seeds;
        p.n_stages = g.stages_bwd;
        {
            ScopedTimer timer(st, 1);
            TDE_CUDA(g.disp->bwd(p, g.n_ctas, g.smem_bwd, st));
        }
        g_launches += 2;  // reverse sweep + second-stage reduction
    }
    ReduceArgs r{};
    r.n_layers = net->n_layers; r.W = g.W; r.width = g.width; r.d_in = g.d_in; r.n_out = g.n_out;
    r.rsplit = g.rsplit; r.n_active = n_active; r.n_active_hidden = n_active_hidden; r.n_params = g.n_params; r.gp_stride = g.gp_stride;
    r.g_b_in = g.g_b_in; r.g_hidden = g.g_hidden; r.g_bh = g.g_bh; r.g_wo = g.g_wo; r.g_bo = g.g_bo;
    r.n_copies = g.n_copies;
    r.dw_w = g.dw_w; r.dw_split = g.use_tc_bwd ? g.dw_split : 1;
    for (int k = 0; k < 4; ++k) { r.dup_lo[k] = g.g_dup_lo[k]; r.dup_hi[k] = g.g_dup_hi[k]; }
    r.scale = grad_scale;
    const int threads = 256;
    reduce_grads_kernel<<<(int)((g.n_params + threads - 1) / threads), threads, 0, st>>>(r, ws + g.o_gpart, flat_grad);
    TDE_CUDA(cudaGetLastError());
    return 0;
}

int tde_jet_backward(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points, const float* stash,
                     const float* seeds, int64_t ld_seeds, float grad_scale, float* flat_grad, void* workspace,
                     void* stream) {
    Geometry g;
    if (int rc = make_geometry(net, spec, n_points, &g)) return rc;
    if (n_points == 0) return 0;
    if (!x || !stash || !seeds || !flat_grad || !workspace) return fail("null x/stash/seeds/flat_grad/workspace");
    if (ld_seeds < n_points) return fail("ld_seeds %lld < n_points %lld", (long long)ld_seeds, (long long)n_points);
    return backward_core(net, spec, g, x, n_points, stash, seeds, ld_seeds, grad_scale, flat_grad, (float*)workspace, (cudaStream_t)stream);
}

static int64_t gcd64(int64_t a, int64_t b) { return b == 0 ? a : gcd64(b, a % b); }

int tde_jet_backward_range(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points, const float* stash,
                           const float* seeds, int64_t ld_seeds, int64_t p_lo, int64_t p_hi, float grad_scale,
                           float* flat_grad, void* workspace, void* stream) {
    Geometry g;
    if (int rc = make_geometry(net, spec, n_points, &g)) return rc;
    if (p_lo < 0 || p_hi > n_points || p_lo > p_hi) return fail("range [%lld, %lld) outside [0, %lld)", (long long)p_lo, (long long)p_hi, (long long)n_points);
    if (p_lo == p_hi) return 0;
    if (!x || !stash || !seeds || !flat_grad || !workspace) return fail("null x/stash/seeds/flat_grad/workspace");
    if (ld_seeds < n_points) return fail("ld_seeds %lld < n_points %lld", (long long)ld_seeds, (long long)n_points);
    // whole tiles of every kernel on the path: the FFMA tile (g.TP points) and the cluster's pair of 16-point tiles of the
    // tensor-core kernels (the sub-range's own padding tile must stay inside the padded extent of the full stash / Hbar).
    // Every layer of the stash is [tile][W x (tile points x J)] in both layouts, so a tile-aligned point offset p is the same
    // float offset p * J * W inside each layer.
    const int64_t align = (int64_t)g.TP / gcd64(g.TP, 32) * 32;
    const int64_t lo = p_lo / align * align;
    Geometry gs;
    if (int rc = make_geometry(net, spec, p_hi - lo, &gs)) return rc;
    if (gs.use_tc != g.use_tc || gs.use_tc_bwd != g.use_tc_bwd || gs.TP != g.TP || gs.workspace_floats > g.workspace_floats)
        return fail("range call plans a different kernel path than the forward that wrote the stash");
    gs.stash_layer_stride = g.stash_layer_stride;
    return backward_core(net, spec, gs, x + lo * g.d_in, p_hi - lo, stash + lo * (int64_t)g.J * g.W, seeds + lo, ld_seeds, grad_scale,
                         flat_grad, (float*)workspace, (cudaStream_t)stream);
}

static int launch_residual_loss(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                                const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float inv_count,
                                const float* coef_dev, float* row_sums, float* seeds, float* point_res, float* point_w, void* stream);

int tde_residual_loss(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                      const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float inv_count,
                      const float* coef_dev, float* row_sums, float* seeds, float* point_res, void* stream) {
    if (!row_sums && !seeds && !point_res) return fail("nothing to compute: row_sums, seeds and point_res are all null");
    return launch_residual_loss(prog, streams, ld, n_stream_rows, aux, ld_aux, p0, n, inv_count, coef_dev, row_sums, seeds, point_res,
                                nullptr, stream);
}

int tde_residual_loss_points(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                             const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float* point_res, float* point_w,
                             void* stream) {
    if (!point_res && !point_w) return fail("nothing to compute: point_res and point_w are both null");
    return launch_residual_loss(prog, streams, ld, n_stream_rows, aux, ld_aux, p0, n, 1.0f, nullptr, nullptr, nullptr, point_res,
                                point_w, stream);
}

static int launch_residual_loss(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                                const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float inv_count,
                                const float* coef_dev, float* row_sums, float* seeds, float* point_res, float* point_w, void* stream) {
    if (!prog || !streams) return fail("null program/streams");
    if (prog->n_grids < 0 || prog->n_grids > TDE_MAX_GRIDS) return fail("n_grids %d out of range", prog->n_grids);
    for (int gi = 0; gi < prog->n_grids; ++gi) {
        const TdeGrid2D& gd = prog->grids[gi];
        if (!gd.gx || !gd.gy || !gd.values || gd.nx < 1 || gd.ny < 1) return fail("grid %d: null pointer or empty axis", gi);
    }
    for (int r = 0; r < prog->n_rows && r < TDE_MAX_ROWS; ++r)
        if (prog->w_reg[r] >= TDE_MAX_REGS) return fail("row %d: weight register %d out of range", r, prog->w_reg[r]);
    if (prog->n_ops < 1 || prog->n_ops > TDE_MAX_OPS) return fail("n_ops %d out of range", prog->n_ops);
    if (prog->n_rows < 1 || prog->n_rows > TDE_MAX_ROWS) return fail("n_rows %d out of range", prog->n_rows);
    for (int k = 0; k < prog->n_ops; ++k) {
        const TdeOp& op = prog->ops[k];
        if (op.dst < 0 || op.dst >= TDE_MAX_REGS) return fail("op %d: dst register %d out of range", k, op.dst);
        if (op.op == TDE_OP_LOAD && (op.a < 0 || op.a >= n_stream_rows)) return fail("op %d: stream row %d out of range", k, op.a);
        if (op.op == TDE_OP_AUX && !aux) return fail("op %d reads aux but aux is null", k);
        if (op.op == TDE_OP_GRID2D && ((int)op.imm < 0 || (int)op.imm >= prog->n_grids)) return fail("op %d: grid %d out of range", k, (int)op.imm);
        if (op.op == TDE_OP_GRID2D && (op.b < 0 || op.b >= TDE_MAX_REGS)) return fail("op %d: source register out of range", k);
        if (op.op >= TDE_OP_ADD && (op.a < 0 || op.a >= TDE_MAX_REGS)) return fail("op %d: source register out of range", k);
    }
    if (n == 0) return 0;
    const int blocks = (int)std::min<int64_t>((n + kLossThreads - 1) / kLossThreads, kLossBlocks);
    residual_loss_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(*prog, streams, ld, n_stream_rows, aux, ld_aux,
                                                                             p0, n, inv_count, coef_dev, row_sums, seeds, point_res, point_w,
                                                                             (int)(g_loss_seq++ % kLossSlots));
    ++g_launches;
    TDE_CUDA(cudaGetLastError());
    return 0;
}

int tde_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, int32_t step, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq) return fail("null adam buffer");
    if (n == 0) return 0;
    if (step < 1) return fail("adam step counter starts at 1, got %d", step);
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 8);
    adam_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                                              (float)(1.0 - beta2), (float)eps, (float)(-lr / bc1), (float)sqrt(bc2));
    ++g_launches;
    TDE_CUDA(cudaGetLastError());
    return 0;
}

long long tde_launch_count(void) { return g_launches; }

int tde_set_option(const char* name, int value) {
    if (!name) return fail("null option name");
    Options& o = options();
    const std::string n(name);
    if (n == "no_tc") o.no_tc = value;
    else if (n == "no_tc_bwd") o.no_tc_bwd = value;
    else if (n == "tc_pair") o.tc_pair = value;
    else if (n == "tc_slots") o.tc_slots = value;
    else if (n == "tc_stages") o.tc_stages = value;
    else if (n == "tc_dbg") o.tc_dbg = value;
    else return fail("unknown option '%s' (no_tc, no_tc_bwd, tc_pair, tc_slots, tc_stages, tc_dbg)", name);
    return 0;
}

int tde_profile_begin(void) {
    for (auto& t : g_timed) {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    g_timed.clear();
    g_profiling = true;
    return 0;
}

static double g_prof_ms[3] = {0.0, 0.0, 0.0};
static int g_prof_cnt[3] = {0, 0, 0};

int tde_profile_kind(int kind, double* ms_total, int32_t* n) {
    if (kind < 0 || kind > 2) return fail("profile kind %d out of range (0 forward, 1 reverse sweep, 2 weight-gradient GEMM)", kind);
    if (ms_total) *ms_total = g_prof_ms[kind];
    if (n) *n = g_prof_cnt[kind];
    return 0;
}

int tde_profile_end(double* fwd_ms, int32_t* n_fwd, double* bwd_ms, int32_t* n_bwd) {
    g_profiling = false;
    double ms[3] = {0.0, 0.0, 0.0};
    int cnt[3] = {0, 0, 0};
    for (auto& t : g_timed) {
        TDE_CUDA(cudaEventSynchronize(t.e1));
        float m = 0.f;
        TDE_CUDA(cudaEventElapsedTime(&m, t.e0, t.e1));
        ms[t.kind] += m;
        cnt[t.kind]++;
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    g_timed.clear();
    for (int k = 0; k < 3; ++k) {
        g_prof_ms[k] = ms[k];
        g_prof_cnt[k] = cnt[k];
    }
    if (fwd_ms) *fwd_ms = ms[0];
    if (n_fwd) *n_fwd = cnt[0];
    if (bwd_ms) *bwd_ms = ms[1] + ms[2];     // whole reverse pass (sweep + weight-gradient GEMM when it is a separate kernel)
    if (n_bwd) *n_bwd = cnt[1];
    return 0;
}

int tde_fp32_peak_probe(double* tflops, void* stream) {
    if (!tflops) return fail("null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    const int threads = 512, grid = sms * 2, iters = 20000;
    float* buf = nullptr;
    TDE_CUDA(cudaMalloc(&buf, sizeof(float) * threads * grid));
    cudaEvent_t e0, e1;
    TDE_CUDA(cudaEventCreate(&e0));
    TDE_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, st);
        ffma_probe_kernel<<<grid, threads, 0, st>>>(buf, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    TDE_CUDA(cudaGetLastError());
    *tflops = 2.0 * 16 * iters * (double)threads * grid / (best * 1e-3) * 1e-12;
    return 0;
}

int tde_grid_interp2d(const float* gx, int32_t nx, const float* gy, int32_t ny, const float* values, const float* x, const float* y,
                      int64_t n, float* out, void* stream) {
    if (n == 0) return 0;
    if (!gx || !gy || !values || !x || !y || !out) return fail("null pointer");
    if (nx < 1 || ny < 1 || n < 0) return fail("grid %d x %d, %lld points", nx, ny, (long long)n);
    const int threads = 512;
    const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, (int64_t)sm_count() * 4);
    const size_t smem = (size_t)(nx + ny) * sizeof(float);
    if (smem <= 40 * 1024)
        grid_interp2d_kernel<true><<<blocks, threads, smem, (cudaStream_t)stream>>>(gx, nx, gy, ny, values, x, y, n, out);
    else
        grid_interp2d_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(gx, nx, gy, ny, values, x, y, n, out);
    TDE_CUDA(cudaGetLastError());
    ++g_launches;
    return 0;
}

}  // extern "C"
