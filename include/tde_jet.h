/*
 * tde_jet.h -- C ABI of the B200-native Taylor-mode jet engine for torch_DE's PINN residual hot path.
 *
 * Drop-in boundary.  Each entry point replaces a piece of the reference's
 * pure-Python/torch.func path; the reference-side binding is the `engine` subclass shown in
 * INTEGRATION.md (ctypes, `tensor.data_ptr()`, `torch.cuda.current_stream().cuda_stream`).
 *
 *   tde_jet_forward    replaces  AD_engine.autodiff + assign_derivs
 *                                (src/torch_DE/continuous/Engines/AD.py:70-90, 115-131):
 *                                vmap(jacrev(jacrev(net)))(x) and the per-name slicing.
 *   tde_jet_backward   replaces  torch autograd's walk of that graph in `loss.backward()`
 *                                (call sites Tutorials/Spring.py:54, Lid_Driven_Cavity.py:95).
 *   tde_residual_loss  replaces  Loss_handler.calculate + Loss.{point_error,weighted_point_error,aggregated_loss}
 *                                (src/torch_DE/utils/loss.py:187-206, 33-59) for residuals expressible as a
 *                                register program, and emits d(loss)/d(stream) seeds for tde_jet_backward.
 *   tde_adam_step      replaces  torch.optim.Adam.step on the flat parameter/gradient buffers
 *                                (Tutorials/Lid_Driven_Cavity.py:71,97) -- same update rule, one launch.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (torch's caching allocator),
 * all tensors FP32, kernels are enqueued on `stream` and never synchronise the host.  Return value
 * 0 = ok, non-zero = error (message via tde_last_error()).  No CPU fallback exists.
 *
 * Jet channels: channel 0 is the value; channels 1..n_first are d/dx_{first_in[i]}; the remaining
 * n_second channels are d2/(dx_a dx_b) where a,b index FIRST-ORDER CHANNELS (0-based, a<=b) and the
 * pairs are sorted by (a,b).  J = 1 + n_first + n_second.
 */
#ifndef TDE_JET_H
#define TDE_JET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDE_ACT_NONE 0
#define TDE_ACT_TANH 1
#define TDE_ACT_SIN 2
/* OR-ed into layers[0].act: the first layer is torch_DE's Fourier embedding, out = act(cat[sin(Wx+b), cos(Wx+b)]) over
 * out_features/2 frequencies (Networks.py:53-75,104-120).  The caller passes the duplicated weight/bias (rows f and f+half equal). */
#define TDE_ACT_FOURIER 0x100

#define TDE_MAX_LAYERS 24
#define TDE_MAX_IN 8
#define TDE_MAX_OUT 4
#define TDE_MAX_FIRST 3
#define TDE_MAX_SECOND 6

/* One nn.Linear: weight[out][in] row-major, bias[out] (may be NULL), followed by activation `act`. */
typedef struct TdeLayer {
    const float* weight;
    const float* bias;
    int32_t in_features;
    int32_t out_features;
    int32_t act;
    int32_t reserved;
} TdeLayer;

/* MLP.forward (Networks.py:148-152): layers[0..n_layers-2] activated, layers[n_layers-1] is linear_out. */
typedef struct TdeNet {
    int32_t n_layers;
    int32_t reserved;
    TdeLayer layers[TDE_MAX_LAYERS];
} TdeNet;

typedef struct TdeJetSpec {
    int32_t d_in;
    int32_t n_out;
    int32_t n_first;
    int32_t n_second;
    int32_t first_in[TDE_MAX_FIRST];   /* input column of each first-order channel */
    int32_t second_a[TDE_MAX_SECOND];  /* first-order channel numbers, a <= b, sorted */
    int32_t second_b[TDE_MAX_SECOND];
} TdeJetSpec;

/* Geometry the host needs to size its buffers; filled by tde_plan(). */
typedef struct TdePlan {
    int32_t n_channels;       /* J */
    int32_t width_padded;     /* hidden width rounded up for the kernels */
    int32_t tile_points;      /* collocation points per CTA tile */
    int32_t n_ctas;           /* persistent grid size */
    int64_t n_tiles;
    int64_t stash_bytes;      /* forward -> backward activation stash (0 if B == 0) */
    int64_t workspace_bytes;  /* packed weights + per-CTA partial gradients (+ the Hbar stream of the tensor-core backward) */
    int64_t n_params;         /* floats in the flat gradient, net.parameters() order */
    int32_t smem_forward;
    int32_t smem_backward;
    int64_t workspace_forward_bytes;  /* what tde_jet_forward alone touches of the workspace (packed weights) */
} TdePlan;

const char* tde_last_error(void);
int tde_version(void);
/* sizeof of the ABI structs as compiled (0 TdeNet, 1 TdeJetSpec, 2 TdePlan, 3 TdeLossProgram): bindings verify at load. */
int tde_abi_struct_bytes(int which);

/* Validates (net, spec) and reports buffer sizes for a batch of n_points. Host-only, no launch. */
int tde_plan(const TdeNet* net, const TdeJetSpec* spec, int64_t n_points, TdePlan* plan);

/*
 * Forward jets.  x[n_points][d_in] row-major.  out[(c*n_out + o)*ld_out + p] = channel c of output o at
 * point p (SoA: one contiguous stream per (channel, output)).  stash may be NULL (inference / value-only
 * groups: nothing is kept for a backward).  workspace must hold plan.workspace_bytes; workspace, stash and the stream
 * blocks must be 16-byte aligned (vector and bulk-copy accesses; any cudaMalloc / torch allocation is).
 */
int tde_jet_forward(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points,
                    float* out, int64_t ld_out, float* stash, void* workspace, void* stream);

/*
 * Reverse sweep.  seeds has the layout of `out` (d loss / d out); flat_grad[n_params] is ACCUMULATED
 * into (caller zeroes), laid out W0,b0,W1,b1,...,Wout,bout exactly like net.parameters().
 * grad_scale multiplies the contribution (1/N_global for data-parallel means).
 * May be called repeatedly on one stash (GradNorm: utils/Loss_weighting.py:39, retain_graph=True).
 */
int tde_jet_backward(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points,
                     const float* stash, const float* seeds, int64_t ld_seeds, float grad_scale,
                     float* flat_grad, void* workspace, void* stream);

/*
 * Reverse sweep restricted to the points [p_lo, p_hi) of a forward over n_points (same x / stash / seeds buffers as
 * tde_jet_backward; workspace of plan(n_points)).  The sweep starts at the tile that holds p_lo: the seeds of the points
 * between that tile's first point (p_lo rounded down to a multiple of lcm(plan.tile_points, 32)) and p_lo MUST be zero.
 * This is the building block of the fused GradNorm (utils/Loss_weighting.py:30-50 calls autograd.grad once per loss row,
 * each a full reverse pass): a loss row only has seeds on the points of its own group, so its gradient costs a sweep over
 * that group, not over the batch.
 */
int tde_jet_backward_range(const TdeNet* net, const TdeJetSpec* spec, const float* x, int64_t n_points,
                           const float* stash, const float* seeds, int64_t ld_seeds, int64_t p_lo, int64_t p_hi,
                           float grad_scale, float* flat_grad, void* workspace, void* stream);

/* ---- fused residual assembly + loss reduction ------------------------------------------------ */
#define TDE_OP_CONST 0  /* dst = imm                      */
#define TDE_OP_LOAD 1   /* dst = stream[a][p]             (a = row of `streams`)                 */
#define TDE_OP_AUX 2    /* dst = aux[a][p]                (per-point constants: x, y, sdf, rhs)  */
#define TDE_OP_ADD 3
#define TDE_OP_SUB 4
#define TDE_OP_MUL 5
#define TDE_OP_NEG 6
#define TDE_OP_POWI 7   /* dst = src_a ** (int)imm        */
#define TDE_OP_ABS 8
#define TDE_OP_SIN 9
#define TDE_OP_COS 10
#define TDE_OP_DIV 11
#define TDE_OP_GRID2D 12 /* dst = bilinear lookup of grids[(int)imm] at (x = reg a, y = reg b): the signed-distance /
                            point-weight field of utils/GridInterpolator.py:50-83, evaluated inside the loss kernel.
                            Constant w.r.t. the streams (its arguments are input coordinates): no adjoint. */
#define TDE_MAX_REGS 64
#define TDE_MAX_OPS 128
#define TDE_MAX_ROWS 8
#define TDE_MAX_GRIDS 2

/* a regular 2-D grid on the device: gx[nx], gy[ny] ascending, values[nx][ny] */
typedef struct TdeGrid2D {
    const float* gx;
    const float* gy;
    const float* values;
    int32_t nx, ny;
} TdeGrid2D;

typedef struct TdeOp {
    int32_t op, dst, a, b;
    float imm;
} TdeOp;

/* One loss-table group: `n_rows` residuals over the same points; row r reads register res_reg[r]; its per-point weight is
 * register w_reg[r] (a weighting function traced into the program, e.g. the SDF lookup), else aux row w_aux[r], else the
 * constant w_const[r]; loss_r = coef[r]*mean(w*|res|^2). */
typedef struct TdeLossProgram {
    int32_t n_ops;
    int32_t n_rows;
    TdeOp ops[TDE_MAX_OPS];
    int32_t res_reg[TDE_MAX_ROWS];
    int32_t w_aux[TDE_MAX_ROWS];
    float w_const[TDE_MAX_ROWS];
    float coef[TDE_MAX_ROWS];
    int32_t w_reg[TDE_MAX_ROWS];      /* -1: no traced weight */
    int32_t n_grids;
    int32_t reserved;
    TdeGrid2D grids[TDE_MAX_GRIDS];
} TdeLossProgram;

/*
 * streams[(row)*ld + p] for p in [p0, p0+n) (row indexing as `out` above), aux[(row)*ld_aux + p - p0].
 * Any of the three outputs may be NULL:
 *   row_sums[n_rows] += inv_count * sum_p w*|res|^2          (un-weighted by coef; caller recombines)
 *   seeds[(row)*ld + p] = sum_r coef_r * inv_count * d(w*res_r^2)/d(stream)   (overwrites the [p0,p0+n) slice of
 *                         every stream row; rows never loaded get zeros); coef_r = prog.coef[r] * coef_dev[r]
 *                         (coef_dev: optional DEVICE array = d total / d row, so no host sync is needed)
 *   point_res[n_rows][n]: per-point residuals (what utils/sampling.py:R3_sampler consumes);
 *   point_w[n_rows][n]   (tde_residual_loss_points only): the per-point weights the rows were reduced with.
 */
int tde_residual_loss(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                      const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float inv_count,
                      const float* coef_dev, float* row_sums, float* seeds, float* point_res, void* stream);

/* Same kernel, forward only, emitting what Loss.point_error() / weighted_point_error() and R3_sampler need: the per-point
 * residuals and weights of every row (either may be NULL). */
int tde_residual_loss_points(const TdeLossProgram* prog, const float* streams, int64_t ld, int32_t n_stream_rows,
                             const float* aux, int64_t ld_aux, int64_t p0, int64_t n, float* point_res, float* point_w,
                             void* stream);

/* Adam with torch.optim.Adam's arithmetic (bias-corrected, eps outside sqrt of v_hat; hyper-parameters are doubles rounded to
 * float once, like torch's python scalars), one launch.  `step` counts from 1. */
int tde_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  double lr, double beta1, double beta2, double eps, int32_t step, void* stream);

/* Instrumentation for bench.py: kernels launched by this library since load, and CUDA-event timing of the
 * two main kernels (events are recorded on the launching stream right around the kernel, nothing else). */
long long tde_launch_count(void);
int tde_profile_begin(void);
int tde_profile_end(double* forward_ms_total, int32_t* n_forward, double* backward_ms_total, int32_t* n_backward);
/* after tde_profile_end: per-kernel totals. kind 0 forward, 1 reverse sweep, 2 weight-gradient GEMM (tensor-core path). */
int tde_profile_kind(int kind, double* ms_total, int32_t* n_launches);
/* which kernels a (net, spec, n_points) call uses: bit0 tcgen05 forward, bit1 tcgen05 reverse sweep + dW GEMM. */
int tde_plan_paths(const TdeNet* net, const TdeJetSpec* spec, int64_t n_points);

/* Run-time switches for A/B measurements and debugging (the matching TDE_* environment variables are read once, at the first
 * call into the library): "no_tc", "no_tc_bwd", "tc_pair", "tc_slots", "tc_stages", "tc_dbg". */
int tde_set_option(const char* name, int value);

/* Measures the sustained scalar-FFMA rate of this GPU (TFLOP/s); the FFMA kernels' roofline denominator. */
int tde_fp32_peak_probe(double* tflops, void* stream);

/* Bilinear lookup on a regular 2-D grid (signed-distance / point-weight fields), the device replacement of
 * torch_DE/utils/GridInterpolator.py:50-83 (RegularGridInterpolator.interpolate for two dimensions):
 *   gx[nx], gy[ny] ascending grid coordinates, values[nx][ny] row-major, x[n], y[n] query coordinates, out[n].
 * Same conventions as the reference: right index = first grid point >= x (clamped), left = right - 1 (clamped), negative
 * distances are 0, a query exactly on a grid line whose both distances are 0 gets weights 1/1, outside the grid the edge value
 * is returned.  All pointers are device pointers. */
int tde_grid_interp2d(const float* gx, int32_t nx, const float* gy, int32_t ny, const float* values, const float* x, const float* y,
                      int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDE_JET_H */
