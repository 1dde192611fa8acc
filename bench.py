#!/usr/bin/env python
"""Headline benchmark: sampled trajectories/s of a full rolling-window MPV sweep (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (configs[1], SURVEY 8d "config 2"): NsDiff on resource-biomass dynamics, Barabasi-Albert 100-node
graph, K = 100 samples per window.  Synthetic input [Node=100, Ls=1000, F=1] (per-node AR(1) around 5.0,
seed 0), L = O = 100, step 5 -> W = 181 windows, B = 100 rows, S = 10, T = 20; architecture from the
reference's model_compare/NsDiff/biomass YAML (f(x) d512/h8/ff256/e4/d2, g(x) R=50, 3x128 denoiser) with
seeded random weights (the checkpoint is absent from the reference tree), scaler mean 0 / std 1.

A "step" is one pass of the hot path over the whole sweep: f(x) + g(x) once per window row, the fused
reverse-diffusion sampler (W*B*K = 1.81e6 trajectories, 3.62e9 denoiser row-steps), the Welford MPV reduction.
  value : whole-job trajectories/s with the windows already resident in HBM (device-timed, max over ranks)
  e2e   : the same through the public host API (uncertainty.sample_sweep) with HOST buffers: H2D of the
          windows from pinned memory, D2H of the trajectory cache + per-window MPV inside the timed region
Multi-GPU (weak scaling): every rank sweeps its own W windows (rank-seeded series), one all-gather of the
per-window statistics, no data-path collective.
`--impl reference` times the reference's CPU implementation of the same path: the reference is pure Python
that cannot travel to the GPU box, so this arm runs the CPU oracle port (oracle/, pinned to the reference by
the golden fixtures; f(x) part "parity unpinned") on all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sampled trajectories/sec (K x windows), NsDiff MPV sweep"
UNIT = "trajectories/s"
YAML = os.path.join(ROOT, "tests", "golden", "ews_results", "model_compare", "NsDiff", "biomass", "model_trained.yaml")
MACS_PER_ROW_STEP_F1 = 3 * 1 * 128 + 2 * 128 * 128 + 2 * 128 * 1          # 33 408 for F = 1


def workload_config():
    import yaml
    with open(YAML) as f:
        cfg = yaml.safe_load(f)
    return cfg


def make_series(rank=0, nodes=100, length=1000):
    """[Node, Ls, F=1]: per-node AR(1) around 5.0, seed = rank."""
    g = torch.Generator().manual_seed(rank)
    e = torch.randn(nodes, length, generator=g) * 0.1
    x = torch.zeros(nodes, length)
    for t in range(1, length):
        x[:, t] = 0.99 * x[:, t - 1] + e[:, t]
    return (x + 5.0).unsqueeze(-1)


def config_dict(cfg, W, B, extra=None):
    net = cfg["net"]
    d = {"workload": "configs[1]: NsDiff, biomass dynamics, BA-100 graph, K=100 samples/window",
         "windows_per_gpu": W, "rows_per_window": B, "n_z_samples": net["n_z_samples"],
         "parallel_sample": net["parallel_sample"], "diffusion_steps": net["diffusion_steps"],
         "window_len": net["windows"], "pred_len": net["pred_len"], "dataset_nf": net["dataset_nf"],
         "weights": "seeded random (seed 123), arch = model_compare/NsDiff/biomass YAML",
         "cache_flush": "inputs larger than L2: each step writes a 724 MB trajectory cache"}
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        mx = max(float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "bf16_burst": p.get("bf16_tflops"),
                "hbm": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """Per-launch DRAM bytes of the sampler from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "sampler_tc_ncu_full.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch_bench_workload")
    return None


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, rows, n_steps, n_warm, threads):
    """Trajectories/s of the CPU oracle port (reference algorithm) on `rows` window-rows per step."""
    from oracle import fx_oracle, nsdiff_oracle, sigma_oracle
    from types import SimpleNamespace
    import updgm_b200  # noqa: F401  (only to instantiate the same seeded random weights on the CPU)
    from updgm_b200.nsdiff import NsDiff_model

    torch.set_num_threads(threads)
    net = dict(cfg["net"], device="cpu")
    torch.manual_seed(123)
    model = NsDiff_model(net, "NsDiff_model")          # parameter container only; arithmetic below is the oracle's
    model.scaler_std.fill_(1.0)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    fx_sd = {k[len("cond_pred_model."):]: v for k, v in sd.items() if k.startswith("cond_pred_model.")}
    fx_cfg = dict(net, seq_len=net["windows"], label_len=net["windows"] // 2)
    sched = nsdiff_oracle.nsdiff_schedule(net["diffusion_schedule"], net["diffusion_steps"], net["beta_start"], net["beta_end"])
    series = make_series(0)
    L, O, K = net["windows"], net["pred_len"], net["n_z_samples"]
    times = []
    with torch.no_grad():
        for it in range(n_warm + n_steps):
            w = it % 181
            x = series[:rows, w * 5: w * 5 + L, :]
            t0 = time.perf_counter()
            y0 = fx_oracle.ns_transformer(fx_sd, fx_cfg, x)[:, -O:, :]
            gx = sigma_oracle.sigma_estimation(sd, x, net["rolling_length"], O)
            outs = nsdiff_oracle.evaluation_step(sd, net, x, sched=sched, y_0_hat=y0, gx=gx)
            _ = outs.var(dim=-1, unbiased=False).mean()
            dt = time.perf_counter() - t0
            if it >= n_warm:
                times.append(dt)
    total = sum(times)
    return rows * K * len(times) / total, total / len(times)


def run_reference_arm(args, rank, world):
    cfg = workload_config()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: size the rows per step so the whole run stays within ~3 minutes
    probe_rate, _ = cpu_reference_rate(cfg, rows=2, n_steps=1, n_warm=0, threads=threads)
    budget = 150.0 / max(1, args.steps + args.warmup)
    rows = int(max(1, min(100, probe_rate * budget / cfg["net"]["n_z_samples"])))
    rate, sec = cpu_reference_rate(cfg, rows=rows, n_steps=args.steps, n_warm=args.warmup, threads=threads)
    sample = "{} of 100 rows of one window per step (K=100, T=20), f(x)+g(x)+sampler+MPV".format(rows)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, 181, 100, {"reference_arm": "CPU oracle port of the reference path (torch CPU ops, "
                                                 "reference is Python and absent on the GPU box)"}),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_own_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    import updgm_b200  # noqa: F401
    from updgm_b200 import kernels, uncertainty as U
    from updgm_b200.nsdiff import NsDiff_model

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = workload_config()
    net = dict(cfg["net"], device=dev)
    torch.manual_seed(123)
    model = NsDiff_model(net, "NsDiff_model").eval()
    model.scaler_std.fill_(1.0)                        # scaler mean 0 / std 1
    L, O, F = net["windows"], net["pred_len"], net["dataset_nf"]
    S = int(net["parallel_sample"])
    K = (int(net["n_z_samples"]) // S) * S
    T = net["diffusion_steps"]
    series = make_series(rank)
    stacked = U.stacked_sliding_windows(series, L, 5).contiguous()             # [W,B,L,F] raw units, host
    W, B = stacked.shape[0], stacked.shape[1]
    host_windows = stacked.pin_memory()
    x_dev = model.scaler_transform(host_windows.to(dev)).contiguous()
    traj = torch.empty((W * B, K, O, F), dtype=torch.float32, device=dev)
    packed = model.packed_weights()
    n_traj = W * B * K
    row_steps = n_traj * O * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from updgm_b200 import _lib
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    sampler_ms = []

    def step_resident(i, timed):
        with torch.no_grad():
            y0, gx = model.condition(x_dev.view(W * B, L, F))
            a, b = ev(), ev()
            a.record()
            kernels.nsdiff_sample(packed, y0, gx, W, B, K, S, O, F, T, seed=1234, window_base=i * W, out=traj)
            b.record()
            red = kernels.mpv_reduce(traj, W, B)
        if timed:
            sampler_ms.append((a, b))
        return red

    for i in range(args.warmup):
        step_resident(i, False)
    barrier()
    launches0 = _lib.kernel_launches()                 # own kernels only: every C-ABI launch is counted in _lib.check
    with ClockSampler(local_rank) as clocks:
        t0, t1 = ev(), ev()
        t0.record()
        for i in range(args.steps):
            red = step_resident(args.warmup + i, True)
        t1.record()
        barrier()
        resident_ms = t0.elapsed_time(t1)
    launches = _lib.kernel_launches() - launches0      # per step: sampler, g(x), Welford + window means, and f(x)'s
                                                       # attention / LayerNorm-split / split kernels per 4096-row chunk
    if world > 1:                                      # the one collective of a sweep (outside no stage of compute)
        local = torch.cat([red["mpv"].view(-1, 1), red["pred_mean"].view(-1, 1), red["mpv_f"]], dim=1)
        stats = U.gather_window_stats(local, W * world)
        assert stats.shape[0] == W * world
    kern_ms = sum(a.elapsed_time(b) for a, b in sampler_ms) / len(sampler_ms)

    # ---- end to end through the public host API: pinned host windows in, trajectory cache + MPV out ----
    for i in range(min(args.warmup, 2)):
        U.sample_sweep(model, host_windows, device=dev)
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(args.steps):
        cache = U.sample_sweep(model, host_windows, device=dev)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = host_windows.numel() * 4
    d2h = cache.numel() * 4 + sum(v.numel() * 4 for part in cache.upd_stats.values() for v in part.values())

    # ---- N > 1: ONE sweep (rank 0's series) sharded over the ranks: contiguous window blocks, rank-local caches, the
    # single all-gather of per-window statistics -- the strong-scaling wall time of BASELINE's second metric
    dist_ms = 0.0
    if world > 1:
        shared = U.stacked_sliding_windows(make_series(0), L, 5).contiguous().pin_memory()
        U.distributed_sweep(model, shared, device=dev)                     # warm-up (NCCL communicator, allocator)
        barrier()
        w0 = time.perf_counter()
        _, (lo, hi), stats = U.distributed_sweep(model, shared, device=dev)
        torch.cuda.synchronize(dev)
        dist_ms = (time.perf_counter() - w0) * 1e3
        assert stats["mpv"].shape[0] == shared.shape[0]

    times = torch.tensor([resident_ms, e2e_ms, dist_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    resident_ms, e2e_ms, dist_ms = times.tolist()
    if rank != 0:
        return
    peaks = measured_peaks()
    value = n_traj * world * args.steps / (resident_ms * 1e-3)
    e2e = n_traj * world * args.steps / (e2e_ms * 1e-3)
    flops = 2.0 * MACS_PER_ROW_STEP_F1 * row_steps
    achieved = flops / (kern_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": resident_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(cfg, W, B),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "kernel": "sampler_tc_kernel<NsDiff,F=1>", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": ncu_traffic(), "peak_source": peaks["source"] + " (bf16 dense, sustained: kernel timed inside a long step)",
                     "algorithmic_flops_per_launch": flops, "kernel_ms": kern_ms,
                     "row_steps_per_s": row_steps / (kern_ms * 1e-3),
                     "note": "algorithmic FLOPs = 2*33408 MAC per denoiser row-step (SURVEY 8a7); the MLP is MUFU-bound "
                             "(514 softplus per row-step), see DESIGN.md"},
    }
    if world > 1:
        line["sweep_wall"] = {"one_sweep_sharded_ms": dist_ms, "windows": W, "ranks": world,
                              "note": "one 181-window sweep split over the ranks (contiguous blocks) + one all-gather of "
                                      "per-window MPV; max over ranks, wall clock"}
    if world == 1:
        # BASELINE's second metric: MPV sweep wall time = windows -> sample -> reduce -> per-window MPV list, cache file
        # written (reference: diffusion_model_uncertainy.py:323-339 + :286-303).  One sweep, outside the timed region.
        import shutil
        import tempfile
        tmp = tempfile.mkdtemp(prefix="upd_bench_")
        try:
            torch.cuda.synchronize(dev)
            w0 = time.perf_counter()
            wins, _ = U.build_sliding_windows(series, torch.arange(series.shape[1]).numpy(), L, 5)
            preds = U.run_evaluation_cache(model, wins, O, os.path.join(tmp, "sweep.pt"), dev, force_recompute=True)
            _, ews = U.summarize_pred_future_list(preds, model=model)
            w1 = time.perf_counter()
            U.flush_cache_writes()
            w2 = time.perf_counter()
            line["sweep_wall"] = {"mpv_list_ms": (w1 - w0) * 1e3, "cache_file_written_ms": (w2 - w0) * 1e3,
                                  "windows": len(ews), "cache_bytes": os.path.getsize(os.path.join(tmp, "sweep.pt")),
                                  "note": "cache written by the background writer; MPV list is available at mpv_list_ms"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if world == 1 and not os.environ.get("UPD_BENCH_SKIP_CPU"):   # (set only when the run is wrapped in ncu)
        threads = os.cpu_count() or 1
        rows = 100
        rate, sec = cpu_reference_rate(cfg, rows=rows, n_steps=2, n_warm=1, threads=threads)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "{} of 100 rows of one window x K=100 (x2 repeats after 1 warm-up), "
                                          "f(x)+g(x)+sampler+MPV, {:.1f} s per repeat".format(rows, sec)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_own_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
