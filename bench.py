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
Multi-GPU (torchrun, STRONG scaling -- BASELINE's "MPV sweep wall time at 1/2/4/8 B200"): the same ONE 181-window sweep
is sharded over the ranks in contiguous window blocks (uncertainty.partition_windows); every step ends with the single
all-gather of per-window statistics, inside the timed region; value = the sweep's 1.81e6 trajectories / max-over-ranks
time.  The weak-scaling figure (every rank sweeps its own 181 windows) is reported as the side key "weak".
Extra keys at N = 1: "parity" (a whole window through the CPU oracle with recorded noise against the GPU path, with the
north-star tolerances), "configs" (BASELINE configs 1, 3, 4-truncated, 5: e2e rate, dominant kernel, roofline fraction),
"sweep_wall", "cpu_baseline".
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
# tensor work actually issued per row-step: two split passes (hi*hi + hi*lo(W)) of the two 128x128 layers (fp16) + three of
# the 8-wide layer 1 (tf32, half rate: counted twice); the dense fp16 rate of an SM is 8192 FLOP/clk (2.25 PFLOP/s / 148
# SMs / 1.86 GHz)
ISSUED_F16_FLOPS_PER_ROW_STEP = 2 * 2 * (2 * 128 * 128) + 2 * (3 * 2 * 8 * 128)


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
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, 181, 100, {"reference_arm": "CPU oracle port of the reference path (torch CPU ops, "
                                                 "reference is Python and absent on the GPU box)"}),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def _ev():
    return torch.cuda.Event(enable_timing=True)


def _timed_ms(fn, n, dev):
    fn()
    torch.cuda.synchronize(dev)
    a, b = _ev(), _ev()
    a.record()
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / n, r


def other_configs(dev, peaks):
    """BASELINE configs 1, 3, 4 (truncated) and 5 through the public sweep API on one GPU: e2e rate with host windows in
    and the trajectory cache out, the dominant kernel and a roofline fraction each (a few seconds per config).  Seeded
    random weights of the YAML architectures except config 1, which is the reference's shipped checkpoint."""
    import yaml
    from updgm_b200 import kernels, uncertainty as U
    G = os.path.join(ROOT, "tests", "golden", "ews_results")
    out = {}
    bf16 = peaks["bf16_sustained"]

    def guarded(name, fn):
        try:
            t0 = time.perf_counter()
            out[name] = fn()
            out[name]["bench_seconds"] = time.perf_counter() - t0
        except Exception as exc:                      # a failing side config must not take the headline line with it
            out[name] = {"error": "{}: {}".format(type(exc).__name__, str(exc)[:200])}

    def config1():
        m, _ = U.load_model_from_dir(os.path.join(G, "NsDiff_machine", "wo_fx"), device=dev)
        g = torch.Generator().manual_seed(0)
        x = torch.zeros(10000, 2)
        e = torch.randn(10000, 2, generator=g) * 0.1
        for t in range(1, 10000):
            x[t] = 0.99 * x[t - 1] + e[t]
        series = x * m.scaler_std.cpu() + m.scaler_mean.cpu()
        wins = series.unfold(0, 200, 10).permute(0, 2, 1).unsqueeze(1).contiguous().pin_memory()      # [981,1,200,2]
        W, K, O, F, T = wins.shape[0], 100, 200, 2, 20
        e2e_ms, _ = _timed_ms(lambda: U.sample_sweep(m, wins, device=dev), 3, dev)
        xs = m.scaler_transform(wins.to(dev)).view(W, 200, 2)
        with torch.no_grad():
            y0, gx = m.condition(xs)
        traj = torch.empty((W, K, O, F), dtype=torch.float32, device=dev)
        k_ms, _ = _timed_ms(lambda: kernels.nsdiff_sample(m.packed_weights(), y0, gx, W, 1, K, 100, O, F, T, seed=1, out=traj,
                                                          impl=m.sampler_impl), 5, dev)
        rs = W * K * O * T
        tf = 2.0 * (3 * F * 128 + 2 * 128 * 128 + 2 * 128 * F) * rs / (k_ms * 1e-3) / 1e12
        return {"workload": "configs[0]: NsDiff SLBP, shipped wo_fx checkpoint, W=981 windows, K=100, F=2, O=200",
                "e2e": {"value": W * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_sweep": e2e_ms},
                "dominant_kernel": "sampler_ws_kernel<NsDiff,F=2>", "kernel_ms": k_ms, "kernel_share_of_e2e": k_ms / e2e_ms,
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16, "unit": "TFLOP/s", "frac": tf / bf16,
                             "row_steps_per_s": rs / (k_ms * 1e-3)}}

    def config3():
        from updgm_b200.tmdm import TMDM_model
        cfg = yaml.safe_load(open(os.path.join(G, "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
        torch.manual_seed(123)
        tm = TMDM_model(dict(cfg["net"], device=dev)).eval()
        g = torch.Generator().manual_seed(0)
        s = torch.sigmoid((torch.randn(100, 1000, 1, generator=g) * 0.1).cumsum(1))
        wins = s.unfold(1, 100, 5).permute(1, 0, 3, 2).contiguous().pin_memory()                        # [181,100,100,1]
        W, B, K, Lr, T = wins.shape[0], 100, 100, 150, 20
        e2e_ms, _ = _timed_ms(lambda: U.sample_sweep(tm, wins, device=dev), 2, dev)
        with torch.no_grad():
            y0 = tm.condition(wins.to(dev).view(W * B, 100, 1))
        traj = torch.empty((W * B, K, Lr, 1), dtype=torch.float32, device=dev)
        k_ms, _ = _timed_ms(lambda: kernels.tmdm_sample(tm.packed_weights(), y0, W, B, K, 10, Lr, 1, T, seed=1, out=traj,
                                                        impl=tm.sampler_impl), 2, dev)
        rs = W * B * K * Lr * T
        tf = 2.0 * (2 * 128 + 2 * 128 * 128 + 128) * rs / (k_ms * 1e-3) / 1e12
        return {"workload": "configs[2]: TMDM neuronal ER-100, W=181, B=100, K=100, 150 positions per trajectory",
                "e2e": {"value": W * B * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_sweep": e2e_ms},
                "dominant_kernel": "sampler_ws_kernel<TMDM,F=1>", "kernel_ms": k_ms, "kernel_share_of_e2e": k_ms / e2e_ms,
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16, "unit": "TFLOP/s", "frac": tf / bf16,
                             "row_steps_per_s": rs / (k_ms * 1e-3)}}

    def config4():
        from updgm_b200.diffusionts import DiffusionTS_model
        cfg = yaml.safe_load(open(os.path.join(G, "model_compare", "DiffusionTS", "SIS", "model_trained.yaml")))
        torch.manual_seed(123)
        d = DiffusionTS_model(dict(cfg["net"], device=dev, n_z_samples=10, parallel_sample=10)).eval()
        d.rows_per_launch = 1000
        g = torch.Generator().manual_seed(0)
        wins = torch.sigmoid((torch.randn(1, 100, 100, 1, generator=g) * 0.1).cumsum(2)).pin_memory()   # 1 window, 100 nodes
        U.sample_sweep(d, wins[:, :10].contiguous(), device=dev)                                        # warm-up (100 rows)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        cache = U.sample_sweep(d, wins, device=dev)
        torch.cuda.synchronize(dev)
        sec = time.perf_counter() - t0
        n = cache.shape[0] * cache.shape[1] * cache.shape[2]
        # 0.43 GFLOP per forward per sequence, 100 forwards + 128 forward/backward (3x) passes per trajectory
        flop = (100 + 3 * 128) * 0.43e9
        tf = n / sec * flop / 1e12
        return {"workload": "configs[3] truncated: DiffusionTS SIS, 1 window x 100 nodes x K=10 (of 199 windows x K=100), "
                            "seq 200, 100 sampling steps with Langevin infill",
                "e2e": {"value": n / sec, "unit": UNIT, "ms_per_sweep": sec * 1e3},
                "dominant_kernel": "dts_attn_tc_bwd_kernel / dts_attn_tc_fwd_kernel, 20 % + 12 % of a step (profiles/r02c_families_launches_summary.txt)",
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16, "unit": "TFLOP/s", "frac": tf / bf16,
                             "note": "whole sweep, algorithmic 208 GFLOP per trajectory (SURVEY 8a14)"}}

    def config5():
        from updgm_b200.diffstg import DiffSTG
        import networkx as nx
        from types import SimpleNamespace
        cfg = yaml.safe_load(open(os.path.join(G, "model_compare", "DiffSTG", "biomass", "model_trained.yaml")))
        torch.manual_seed(123)
        m = DiffSTG(dict(cfg["net"], device=dev)).eval()
        Gr = nx.barabasi_albert_graph(100, 12, seed=0)
        ei = torch.tensor(list(Gr.to_directed().edges)).t().contiguous()
        graph = SimpleNamespace(edge_index=ei.to(dev), num_nodes=100)
        g = torch.Generator().manual_seed(0)
        wins = (torch.randn(4, 100, 100, 1, generator=g) * 0.1).cumsum(2).contiguous().pin_memory()     # 4 windows
        U.sample_sweep(m, wins[:1], device=dev, graph_data=graph)
        e2e_ms, cache = _timed_ms(lambda: U.sample_sweep(m, wins, device=dev, graph_data=graph), 2, dev)
        n = cache.shape[0] * cache.shape[1] * cache.shape[2]                                            # window x node x sample
        tf = n / (e2e_ms * 1e-3) * 1.0e9 / 1e12                                                         # ~1 GFLOP per node-trajectory
        return {"workload": "configs[4] at (T_h, T_p) = (100, 100): DiffSTG biomass, BA-100 graph, 4 windows x 100 nodes x "
                            "100 samples (10 rounds x 10 replicas), 20 DDIM steps",
                "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "node-trajectories/s", "ms_per_sweep": e2e_ms},
                "dominant_kernel": "stg_tcn_mma_kernel, 32 % of a step (gemm3_* 29 %, gated aggregation 18 %; profiles/r02c_families_launches_summary.txt)",
                "roofline": {"bound": "tensor", "achieved": tf, "peak": bf16, "unit": "TFLOP/s", "frac": tf / bf16,
                             "note": "whole sweep, ~1 GFLOP per node-trajectory (SURVEY 8a15 estimate)"}}

    guarded("config1_nsdiff_slbp", config1)
    guarded("config3_tmdm_neuronal", config3)
    guarded("config4_diffusionts_sis", config4)
    guarded("config5_diffstg_biomass", config5)
    return out


def parity_and_cpu_baseline(cfg, model, dev, threads):
    """One whole window of the workload (1e6 denoiser rows x 20 steps) through the CPU oracle with recorded noise:
    (a) the north-star parity check of the GPU path on exactly that window and noise, (b) the CPU baseline rate."""
    from oracle import parity
    net = dict(cfg["net"])
    series = make_series(0)
    x = series[:, 400:500, :].contiguous()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = parity.nsdiff_window_reference(sd, net, x, seed=11, with_fx=True, threads=threads)      # also the warm-up
    with torch.no_grad():
        outs, _ = model.evaluation_step(x.to(dev), noise=ref["noise"][0])
    chk = parity.compare(outs, ref["ref"])
    chk["what"] = ("window 80 of the sweep, whole (100 rows x K=100 x O=100 x T=20), f(x)+g(x)+sampler on the GPU vs the CPU "
                   "oracle on the same recorded noise; f(x) itself is parity-unpinned (DESIGN.md)")
    times = []
    for rep in range(2):
        times.append(parity.nsdiff_window_reference(sd, net, x, seed=12 + rep, with_fx=True, threads=threads)["seconds"])
    sec = sum(times) / len(times)
    rate = x.shape[0] * net["n_z_samples"] / sec
    base = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "100 of 100 rows of one window x K=100 (x2 repeats after 1 warm-up), f(x)+g(x)+sampler, "
                      "{:.1f} s per repeat".format(sec)}
    return chk, base


def run_own_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    import updgm_b200  # noqa: F401
    from updgm_b200 import _lib, kernels, uncertainty as U
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
    # ONE sweep (the series of rank 0), sharded over the ranks in contiguous window blocks: strong scaling
    stacked = U.stacked_sliding_windows(make_series(0), L, 5).contiguous()      # [W,B,L,F] raw units, host
    W, B = stacked.shape[0], stacked.shape[1]
    lo, hi = U.partition_windows(W, world, rank)
    Wl = hi - lo
    host_windows = stacked.pin_memory()
    x_dev = model.scaler_transform(host_windows[lo:hi].to(dev)).contiguous()
    traj = torch.empty((max(Wl, 1) * B, K, O, F), dtype=torch.float32, device=dev)
    packed = model.packed_weights()
    n_traj = W * B * K                                  # of the whole sweep
    row_steps_local = Wl * B * K * O * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler_ms = []

    def step_resident(i, timed):
        with torch.no_grad():
            y0, gx = model.condition(x_dev.view(Wl * B, L, F))
            a, b = _ev(), _ev()
            a.record()
            kernels.nsdiff_sample(packed, y0, gx, Wl, B, K, S, O, F, T, seed=1234, window_base=i * W + lo, out=traj,
                                  impl=model.sampler_impl)
            b.record()
            red = kernels.mpv_reduce(traj, Wl, B)
            if world > 1:                              # the one collective of a sweep, inside the timed region
                local = torch.cat([red["mpv"].view(-1, 1), red["pred_mean"].view(-1, 1), red["mpv_f"]], dim=1)
                stats = U.gather_window_stats(local, W)
                assert stats.shape[0] == W
        if timed:
            sampler_ms.append((a, b))
        return red

    for i in range(args.warmup):
        step_resident(i, False)
    barrier()
    launches0 = _lib.kernel_launches()                 # own kernels only: every C-ABI launch is counted in _lib.check
    with ClockSampler(local_rank) as clocks:
        t0, t1 = _ev(), _ev()
        t0.record()
        for i in range(args.steps):
            step_resident(args.warmup + i, True)
        t1.record()
        barrier()
        resident_ms = t0.elapsed_time(t1)
    launches = _lib.kernel_launches() - launches0      # per step: sampler, g(x), Welford + window means, and f(x)'s
                                                       # attention / LayerNorm-split / split kernels per 4096-row chunk
    kern_ms = sum(a.elapsed_time(b) for a, b in sampler_ms) / len(sampler_ms)

    # ---- end to end through the public host API: pinned host windows in, trajectory cache + per-window MPV out ----
    def sweep_e2e():
        if world > 1:
            return U.distributed_sweep(model, host_windows, device=dev)[0]
        return U.sample_sweep(model, host_windows, device=dev)

    for i in range(min(args.warmup, 2)):
        sweep_e2e()
    barrier()
    e0, e1 = _ev(), _ev()
    e0.record()
    for i in range(args.steps):
        cache = sweep_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = host_windows[lo:hi].numel() * 4
    d2h = cache.numel() * 4 + (sum(v.numel() * 4 for part in cache.upd_stats.values() for v in part.values())
                               if hasattr(cache, "upd_stats") else 0)

    # ---- N > 1, side measurement: weak scaling (every rank sweeps its OWN 181 windows, rank-seeded series) ----
    weak_ms = 0.0
    if world > 1:
        own = model.scaler_transform(U.stacked_sliding_windows(make_series(rank), L, 5).contiguous().to(dev)).contiguous()
        big = torch.empty((W * B, K, O, F), dtype=torch.float32, device=dev)

        def weak_step():
            with torch.no_grad():
                y0, gx = model.condition(own.view(W * B, L, F))
                kernels.nsdiff_sample(packed, y0, gx, W, B, K, S, O, F, T, seed=99, window_base=0, out=big, impl=model.sampler_impl)
                kernels.mpv_reduce(big, W, B)
        weak_step()
        barrier()
        w0_, w1_ = _ev(), _ev()
        w0_.record()
        for _ in range(2):
            weak_step()
        w1_.record()
        barrier()
        weak_ms = w0_.elapsed_time(w1_) / 2

    vals = torch.tensor([resident_ms, e2e_ms, weak_ms, kern_ms, float(launches), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    sums = vals.clone()
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    resident_ms, e2e_ms, weak_ms, kern_ms_max = vals.tolist()[:4]
    launches_all, h2d_all, d2h_all = (int(v) for v in sums.tolist()[4:])
    if rank != 0:
        return
    peaks = measured_peaks()
    value = n_traj * args.steps / (resident_ms * 1e-3)
    e2e = n_traj * args.steps / (e2e_ms * 1e-3)
    flops = 2.0 * MACS_PER_ROW_STEP_F1 * row_steps_local            # rank 0's launch
    achieved = flops / (kern_ms * 1e-3) / 1e12
    rs_per_s = row_steps_local / (kern_ms * 1e-3)
    clk = clocks.summary()
    sm_hz = (clk.get("sm_mhz") or 1900.0) * 1e6
    # MUFU ops issued per row-step: 386 ex2 (384 hidden softplus + the sigma head's outer one) + the lg2 that stay on the
    # MUFU pipe: half of layers 1-2 take lg2(1+u) as a packed-FMA polynomial (csrc/sampler_ws.cu, UPD_WS_PMASK12_NS = 0x33)
    mufu_per_row_step = 386 + (128 + 128 + 2)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": resident_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, W, B, {"windows_per_gpu": Wl if world > 1 else W, "windows_total": W,
                                          "sharding": "one {}-window sweep in contiguous window blocks over {} rank(s); one "
                                                      "all-gather of [W, 2+F] statistics per sweep, inside the timed region"
                                                      .format(W, world)}),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches_all,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "sampler_ws_kernel<NsDiff,F=1>", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": ncu_traffic(), "peak_source": peaks["source"] + " (bf16 dense, sustained: kernel timed inside a long step)",
                     "algorithmic_flops_per_launch": flops, "kernel_ms": kern_ms, "row_steps_per_s": rs_per_s,
                     "mufu_frac": rs_per_s * mufu_per_row_step / (148 * 16 * sm_hz),
                     "tensor_active": rs_per_s * ISSUED_F16_FLOPS_PER_ROW_STEP / (148 * 8192 * sm_hz),
                     "note": "per GPU (rank 0's launch).  algorithmic FLOPs = 2*33408 MAC per denoiser row-step (SURVEY 8a7). "
                             "mufu_frac = 644 issued ex2/lg2 per row-step (386 ex2 + 258 lg2; the other 128 lg2 run as FMA-pipe "
                             "polynomials) against 16 MUFU lanes/clk/SM at the sampled SM clock -- the pipe that bounds this MLP; tensor_active = issued MMA work (2 split passes in layers 2-3, 3 in layer 1) against the dense "
                             "fp16 rate; both derived from the measured row-steps/s, the ncu captures are in profiles/"},
    }
    if world > 1:
        line["weak"] = {"value": n_traj * world / (weak_ms * 1e-3), "unit": UNIT, "ms_per_step": weak_ms,
                        "note": "side measurement: every rank sweeps its own 181 windows (no collective in the loop)"}
    if world == 1:
        # BASELINE's second metric: MPV sweep wall time = windows -> sample -> reduce -> per-window MPV list, cache file
        # written (reference: diffusion_model_uncertainy.py:323-339 + :286-303).  One sweep, outside the timed region.
        import shutil
        import tempfile
        series = make_series(0)
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
        del preds, cache
        if not os.environ.get("UPD_BENCH_SKIP_CONFIGS"):
            line["configs"] = other_configs(dev, peaks)
        if not os.environ.get("UPD_BENCH_SKIP_CPU"):                 # (set only when the run is wrapped in ncu)
            line["parity"], line["cpu_baseline"] = parity_and_cpu_baseline(cfg, model, dev, os.cpu_count() or 1)
    else:
        line["sweep_wall"] = {"one_sweep_sharded_ms": e2e_ms / args.steps, "windows": W, "ranks": world,
                              "note": "= e2e.ms_per_step: one 181-window sweep through uncertainty.distributed_sweep (host "
                                      "windows in, rank-local caches + gathered MPV list out), device-timed, max over ranks"}
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
