#!/usr/bin/env python
"""Benchmark of the PINN residual hot path (BASELINE.json metric: collocation points/sec per training step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ldc]

One step = jets forward -> residual + boundary loss -> reverse to the weight gradient -> (all-reduce) -> Adam,
through the public API (DE_Getter.calculate -> Loss_handler -> backward -> optimizer.step).
Weak scaling: every rank holds `--points-per-gpu` collocation points (default 4e6/8, the LDC config of
BASELINE.json spread over 8 GPUs); `value` is the whole-job points/s.

`--impl reference` times the reference's own CPU implementation of the path -- the UNMODIFIED torch_DE sources that
tools/install_ref.py materialises under baseline/_ref (DE_Getter + AD_engine + Loss_handler + backward + Adam; the oracle port
only if that tree is absent) -- on this box's host cores, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collocation points/sec per PINN training step (residual+grad)"
UNIT = "points/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ldc")
    ap.add_argument("--points-per-gpu", type=int, default=0)
    ap.add_argument("--boundary-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="collocation points per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphed", action="store_true",
                    help="also time the step replayed as one CUDA graph (torch_de_b200.GraphedStep); reported next to the eager value")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)          # the recipe's 200 ms (nvidia-smi -lms 200)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def _reference_stepper(workload, n_col, n_bnd, device):
    """One training step of the reference on fixed synthetic groups: the UNMODIFIED reference (`baseline/_ref`, driven through
    DE_Getter + AD_engine + Loss_handler + backward + Adam by oracle/ref_live.py) when it is in the tree, else the oracle port
    of its torch.func / autograd calls.  Returns (step_fn, points_per_step, kind)."""
    from oracle import ref_autograd as O
    from oracle import ref_live as RL
    owl = O.WORKLOADS[workload]()
    n = owl.net
    params = O.init_mlp_params(n["d_in"], n["n_out"], n["width"], n["n_act"], seed=1234)
    groups = O.sample_groups(owl, n_col, n_bnd, seed=1234)
    n_pts = sum(x.shape[0] for x in groups.values())
    if RL.available():
        live = RL.LiveStep(owl, params, groups, torch.float32, device)
        return live.train_step, n_pts, "reference"
    params = [p.to(device).requires_grad_(True) for p in params]
    opt = torch.optim.Adam(params, lr=1e-3)
    groups = {g: x.to(device) for g, x in groups.items()}

    def one():
        res = O.step(owl, params, groups)
        for p, g in zip(params, res["grads"]):
            p.grad = g
        opt.step()
        opt.zero_grad()
    return one, n_pts, "port"


def cpu_reference_steps(workload, n_col, n_bnd, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one, n_pts, kind = _reference_stepper(workload, n_col, n_bnd, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_pts, times, cores, kind


def gpu_reference_steps(workload, n_col, n_bnd, steps, warmup, dev):
    """The reference's own step (see _reference_stepper) on the SAME GPU at one micro-batch size, CUDA-event timed: the
    denominator of the north star's >= 20x target."""
    one, n_pts, kind = _reference_stepper(workload, n_col, n_bnd, dev)
    for _ in range(warmup):
        one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    return n_pts * steps / (e0.elapsed_time(e1) * 1e-3), n_pts, kind


def gpu_reference_best(workload, dev, sizes=(4096, 16384, 65536, 131072, 262144)):
    """Sweep the micro-batch and keep the reference's best throughput (it is given its best shot); a size that does not fit
    (the reference keeps ~0.5 MB of autograd state per LDC point) ends the sweep."""
    sweep, best = [], None
    for n_col in sizes:
        try:
            v, npts, kind = gpu_reference_steps(workload, n_col, max(1, n_col // 64), 3, 2, dev)
        except (torch.OutOfMemoryError, RuntimeError) as exc:
            if best is None:
                raise
            sweep.append({"points": n_col, "value": None, "note": "out of memory" if isinstance(exc, torch.OutOfMemoryError) else repr(exc)[:80]})
            torch.cuda.empty_cache()
            break
        sweep.append({"points": npts, "value": v})
        if best is None or v > best["value"]:
            best = {"value": v, "points": npts, "kind": kind}
        torch.cuda.empty_cache()
    return best, sweep


def measure_tf32_gemm(dev, n=8192, iters=10):
    """cuBLAS TF32 GEMM throughput (TFLOP/s) on this GPU, measured here: the per-MMA rate the 3xTF32 kernels divide by three.
    A library call used as a yardstick only."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            (a @ b)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            (a @ b)
        e1.record()
        torch.cuda.synchronize(dev)
        torch.backends.cuda.matmul.allow_tf32 = old
        return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    except Exception:  # noqa: BLE001 - yardstick only
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    n_bnd = max(1, args.cpu_sample // 64)
    n_pts, times, cores, kind = cpu_reference_steps(args.workload, args.cpu_sample, n_bnd, args.steps, args.warmup)
    total = sum(times)
    value = n_pts * len(times) / total
    sample = f"{args.cpu_sample} collocation + {n_bnd} pts per boundary group per step, {len(times)} timed steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "device": "host CPU", "points_per_step": n_pts,
                       "implementation": REF_KIND_NOTE[kind]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


REF_KIND_NOTE = {"reference": "unmodified torch_DE from baseline/_ref: DE_Getter + AD_engine (vmap(jacrev(jacrev))) + Loss_handler + "
                              "backward + torch.optim.Adam",
                 "port": "oracle port of the reference's torch.func / autograd calls (baseline/_ref absent)"}


def workload_name(w):
    return {"ldc": "Tutorials/Lid_Driven_Cavity.py: steady 2D Navier-Stokes (u,v,p), MLP 8x256 tanh, 10 derivative names (J=5)",
            "burgers": "Tutorials/Burgers1D.py: MLP 8x64 tanh (J=4)", "allen_cahn": "Tutorials/AllenCahn.py: MLP 6x128 tanh (J=4)",
            "spring": "Tutorials/Spring.py: Sequential(1,200,1) tanh (J=3)",
            "cylinder": "Tutorials/Stokes_Cylinder.py: steady 2D Navier-Stokes, MLP 10x512 tanh (J=5)",
            "cylinder_unsteady": "Tutorials/Unsteady_Cylinder.py: 2D+t Navier-Stokes (x,y,t), MLP 10x512 tanh (J=6)"}[w]


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    import torch_de_b200 as T
    from torch_de_b200 import _lib
    from torch_de_b200 import workloads as WL
    from torch_de_b200.dp import FlatGrads

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.lib()

    wl = WL.WORKLOADS[args.workload]()
    n_col = args.points_per_gpu or max(1, wl.default_points // 8)
    n_bnd = args.boundary_per_gpu or max(1, wl.default_boundary // 8)
    torch.manual_seed(1234)                                  # AllenCahn.py:23, Stokes_Cylinder.py:83
    net = wl.make_net().to(dev)
    pinn = T.DE_Getter(net, wl.input_vars, wl.output_vars, wl.derivatives)
    runner = pinn.deriv_method._full
    state = {}

    def make_shard(n_col_rank, res_scale):
        state["host"] = {g: x.pin_memory() for g, x in wl.sample(n_col_rank, n_bnd, seed=1234 + rank).items()}
        state["dev"] = {g: x.to(dev) for g, x in state["host"].items()}
        state["batch"] = T.make_batch(wl.input_vars, state["dev"])
        state["res_scale"] = res_scale          # weight of this rank's residual rows (1 with equal shards)
        state["n_col"] = n_col_rank
    make_shard(n_col, 1.0)
    losses = wl.build_losses(list(state["dev"].keys()))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)        # Lid_Driven_Cavity.py:71
    flat = FlatGrads(net)

    ar_events = []          # (step start, before, after) CUDA events around the gradient all-reduce of every timed step

    def step(from_host=False, record_ar=False):
        if record_ar and world > 1:
            es = torch.cuda.Event(enable_timing=True)
            es.record()
        if from_host:
            for g, x in state["host"].items():
                state["dev"][g].copy_(x, non_blocking=True)
        out = pinn.calculate(state["batch"])
        loss = losses(state["batch"], out)
        by_type = loss.grouped_losses("loss_type")
        if state["res_scale"] != 1.0:            # unequal shards: the residual mean of this rank weighs n_rank / N, not 1 / world
            by_type["residual"] = by_type["residual"] * state["res_scale"]
        total = wl.combine(by_type)
        total.backward()
        if world > 1:
            if record_ar:
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                flat.all_reduce_mean()
                eb.record()
                ar_events.append((es, ea, eb))
            else:
                flat.all_reduce_mean()
        opt.step()
        flat.zero()
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(n_steps, from_host):
        import gc
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()          # no generation-2 collection of the interpreter (tens of ms) inside a 0.5-1 s timed region
        try:
            return _timed(n_steps, from_host, e0, e1)
        finally:
            gc.enable()

    def _timed(n_steps, from_host, e0, e1):
        barrier()
        e0.record()
        for _ in range(n_steps):
            t = step(from_host, record_ar=not from_host)
            if from_host:
                t.item()          # device -> host read of the step's loss
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    shards = [n_col] * world
    n_pts_rank = sum(x.shape[0] for x in state["dev"].values())
    step()              # settle: one more untimed step after the first full synchronisation (allocator / clocks steady)
    barrier()

    launches0 = lib.tde_launch_count()
    lib.tde_profile_begin()
    with ClockSampler(local_rank) as clk:
        ms_total = timed(args.steps, from_host=False)
    launches = lib.tde_launch_count() - launches0
    lib.tde_profile_end(None, None, None, None)

    for _ in range(2):
        step(True).item()
    ms_e2e = timed(args.steps, from_host=True)

    # per-rank record: summed kernel time of the three main kernels and the time the gradient all-reduce held the stream
    # (wait for the slowest rank + the collective itself), gathered so that rank 0 can report min / median / max
    ksum = 0.0
    for k in range(3):
        m, n = C.c_double(), C.c_int32()
        lib.tde_profile_kind(k, C.byref(m), C.byref(n))
        ksum += m.value
    ar_ms = sum(a.elapsed_time(b) for _, a, b in ar_events)
    mine = torch.tensor([ksum / args.steps, ar_ms / args.steps], device=dev, dtype=torch.float64)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    # per step: this rank's compute time (step start -> gradient ready).  The step ends for everybody when the slowest rank of
    # THAT step is ready, so mean over steps of (max over ranks - mean over ranks) is what synchronous DP pays for rank skew
    # and step-to-step jitter together
    compute = torch.tensor([s0.elapsed_time(a) for s0, a, _ in ar_events] or [0.0], device=dev, dtype=torch.float64)
    all_compute = [torch.zeros_like(compute) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
        dist.all_gather(all_compute, compute)
    else:
        per_rank, all_compute = [mine], [compute]
    per_rank = torch.stack(per_rank).cpu()
    all_compute = torch.stack(all_compute).cpu()            # [ranks, steps]

    graphed_ms = None
    if args.graphed and world == 1:
        import gc
        gc.collect()            # no autograd graph of an eager step may be alive when the capture starts (its AccumulateGrad
        torch.cuda.synchronize(dev)         # nodes belong to the default stream)
        gopt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)

        def graph_body():
            out = pinn.calculate(state["batch"])
            total = wl.combine(losses(state["batch"], out).grouped_losses("loss_type"))
            total.backward()
            gopt.step()
            flat.zero()
            return total
        gstep = T.GraphedStep(graph_body)
        for _ in range(3):
            gstep()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(args.steps):
            gstep()
        e1.record()
        torch.cuda.synchronize(dev)
        graphed_ms = e0.elapsed_time(e1) / args.steps

    if rank != 0:
        return
    n_pts_job = sum(shards) + world * (n_pts_rank - shards[rank])      # every rank holds boundary groups of the same size
    value = n_pts_job * args.steps / (ms_total * 1e-3)
    e2e_value = n_pts_job * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    # algorithmic FLOPs: forward 2*J*sum(W_in*W_out) per point, reverse sweep (dX) 2*..., weight gradient
    # (dW) 2*...; on the FFMA path dX and dW are one kernel.  achieved = FLOPs of all launches of that kernel in the timed
    # region / its summed CUDA-event time (events recorded by the library right around each launch).
    peak = C.c_double()
    _lib.check(lib.tde_fp32_peak_probe(C.byref(peak), None), "tde_fp32_peak_probe")
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    measured = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
    bf16_sustained = measured.get("bf16_tflops_sustained") or 1400.0
    tensor_peak = bf16_sustained / 2.0 / 3.0            # TF32 = bf16/2; 3xTF32 issues 3 MMAs per FP32-accurate product
    tf32_gemm = measure_tf32_gemm(dev)                  # cuBLAS TF32 GEMM timed here, after the timed region
    tensor_src = ("3xTF32 tensor roofline = MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 rate) / 3 (three MMAs per "
                  "FP32-accurate product)" if measured else "fallback 1.4 PFLOP/s bf16 sustained / 2 / 3")
    fp32_src = "scalar-FFMA probe run in this process (tde_fp32_peak_probe); MEASURED_PEAKS.json has no FP32 entry"
    J = runner.ch.n_channels
    f_pt = wl.flops_per_point(J)
    plan = runner.plan(n_pts_rank)
    paths = lib.tde_plan_paths(C.byref(runner._net), C.byref(runner._spec), n_pts_rank)
    recompute = bool(runner.chunking and any(runner.chunking.values()))
    pts_total = n_pts_rank * args.steps
    ms, cnt = [], []
    for k in range(3):
        m, n = C.c_double(), C.c_int32()
        lib.tde_profile_kind(k, C.byref(m), C.byref(n))
        ms.append(m.value)
        cnt.append(n.value)
    kernels = []
    # recompute mode: the backward re-runs the forward chunk by chunk, except for the last chunk whose stash the first pass kept
    fwd_passes = 1.0
    if recompute:
        chunk = max(runner.chunking.values())
        fwd_passes = 2.0 - 1.0 / max(1, -(-n_pts_rank // chunk))
    fwd_flops = (1.0 / 3.0) * f_pt * pts_total * fwd_passes
    kernels.append(dict(kernel="jet_forward_tc_kernel" if paths & 1 else "jet_forward_kernel", flops=fwd_flops, ms=ms[0],
                        launches=cnt[0], tensor=bool(paths & 1)))
    if paths & 2:
        kernels.append(dict(kernel="jet_backward_tc_kernel", flops=(1.0 / 3.0) * f_pt * pts_total, ms=ms[1], launches=cnt[1], tensor=True))
        kernels.append(dict(kernel="jet_dw_tc_kernel", flops=(1.0 / 3.0) * f_pt * pts_total, ms=ms[2], launches=cnt[2], tensor=True))
    else:
        kernels.append(dict(kernel="jet_backward_kernel", flops=(2.0 / 3.0) * f_pt * pts_total, ms=ms[1], launches=cnt[1], tensor=False))
    for k in kernels:
        k["achieved"] = k["flops"] / (k["ms"] * 1e-3) * 1e-12 if k["ms"] > 0 else 0.0
        k["bound"] = "tensor" if k["tensor"] else "fp32"
        k["peak"] = tensor_peak if k["tensor"] else peak.value
        k["frac"] = k["achieved"] / k["peak"]
        k["launch_ms"] = k["ms"] / max(1, k["launches"])
        k["share_of_step"] = k["ms"] / ms_total
        del k["flops"], k["tensor"], k["ms"]
    # DRAM bytes per point and kernel: looked up in the committed table of `ncu --set full` captures
    # (profiles/dram_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch / its points; the table names
    # the capture every entry comes from).  A workload / kernel without a capture reports null.
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    table = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
    traffic_per_point = table.get("bytes_per_point", {}).get(args.workload, {})
    for k in kernels:
        per_pt = traffic_per_point.get(k["kernel"])
        k["traffic"] = per_pt * pts_total / max(1, k["launches"]) if per_pt else None
        k["frac_vs_tf32_gemm"] = k["achieved"] / (tf32_gemm / 3.0) if k["bound"] == "tensor" and tf32_gemm else None
    dom = max(kernels, key=lambda k: k["share_of_step"])
    roofline = {"bound": dom["bound"], "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": "TFLOP/s",
                "frac": dom["frac"], "traffic": dom["traffic"],
                "peak_source": tensor_src if dom["bound"] == "tensor" else fp32_src,
                "launch_ms": dom["launch_ms"], "share_of_step": dom["share_of_step"], "kernels": kernels,
                "fp32_ffma_peak_probe": peak.value,
                "step_fp32_equiv_tflops": f_pt * pts_total / (ms_total * 1e-3) * 1e-12,
                "hbm_gbs_measured": measured.get("hbm_gbs"), "bf16_tflops_measured": measured.get("bf16_tflops"),
                "bf16_tflops_sustained_measured": measured.get("bf16_tflops_sustained"),
                "tf32_gemm_tflops_here": tf32_gemm,
                "frac_vs_tf32_gemm": dom.get("frac_vs_tf32_gemm"),
                "traffic_source": table.get("source")}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        nb = max(1, args.cpu_sample // 64)
        npts, times, cores, kind = cpu_reference_steps(args.workload, args.cpu_sample, nb, 3, 1)
        cpu_baseline = {"value": npts / statistics.median(times), "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{args.cpu_sample} collocation + {nb} pts per boundary group, median of 3 steps after 1 warm-up"}

    reference_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        state.pop("batch"), state.pop("dev")
        torch.cuda.empty_cache()
        try:
            best, sweep = gpu_reference_best(args.workload, dev)
            reference_gpu = {"value": best["value"], "unit": UNIT, "kind": best["kind"], "device": torch.cuda.get_device_name(dev),
                             "sample": f"best of a micro-batch sweep: {best['points']} points per step, 3 timed steps after 2 warm-ups "
                                       "(DE_Getter/AD_engine + Loss_handler + backward + Adam on the same GPU)",
                             "sweep": sweep}
        except Exception as exc:  # noqa: BLE001 - informational leg only
            reference_gpu = {"unavailable": repr(exc)[:200]}

    h2d = sum(x.numel() * 4 for x in state["host"].values())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "points_per_gpu": n_col, "boundary_points_per_group_per_gpu": n_bnd,
                       "points_per_step_job": n_pts_job, "parallelism": f"dp{world} over collocation points, 1 NCCL all-reduce of the flat gradient",
                       "optimizer": "torch.optim.Adam lr=1e-3",
                       "paths": {"forward": "tcgen05 3xTF32" if paths & 1 else "FFMA2", "backward": "tcgen05 3xTF32 (sweep + dW GEMM)" if paths & 2 else "FFMA2",
                                 "recompute": recompute},
                       "l2": "per-step working set (activation stash %.1f GB, workspace %.1f GB) is far larger than the 126 MB L2"
                             % (plan.stash_bytes / 1e9, plan.workspace_bytes / 1e9)},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "reference_gpu": reference_gpu,
            "ranks": {"kernel_sum_ms_per_step": {"min": float(per_rank[:, 0].min()), "median": float(per_rank[:, 0].median()),
                                                 "max": float(per_rank[:, 0].max())},
                      "allreduce_ms_per_step": {"min": float(per_rank[:, 1].min()), "median": float(per_rank[:, 1].median()),
                                                "max": float(per_rank[:, 1].max())},
                      "compute_ms_per_step": {"mean_over_ranks": float(all_compute.mean()),
                                              "mean_of_per_step_max": float(all_compute.max(dim=0).values.mean()),
                                              "slowest_rank_mean": float(all_compute.mean(dim=1).max())},
                      "note": "all-reduce time = wait for the slowest rank of the step + the collective (21-28 us alone, "
                              "tools/allreduce_probe.py), from CUDA events around it; compute = step start -> gradient ready"}}
    if graphed_ms is not None:
        line["graphed"] = {"ms_per_step": graphed_ms, "value": n_pts_job / (graphed_ms * 1e-3), "unit": UNIT,
                           "note": "same step replayed as one CUDA graph (torch_de_b200.GraphedStep); `value` above is the eager public API"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
