"""Inference / analysis entry points with the reference's names, arguments, cache files and result
dictionaries (evaluation_and_analysis/diffusion_model_uncertainy.py), re-organised for the GPU:

  * a sweep is sampled in a few batched launches (all rolling windows at once) instead of a Python
    loop over windows, and the K trajectories are reduced on the device (csrc/mpv_reduce.cu);
  * everything numeric runs through the C ABI; there is no CPU fallback, so the functions that compute
    (as opposed to the pure path / index helpers) need a CUDA device.

Line references in docstrings are to the reference file above unless another file is named.
"""
import atexit
import concurrent.futures
import ctypes
import os
import threading
import weakref
from pathlib import Path

import numpy as np
import torch
import yaml

from . import _lib, kernels

NETWORK_DYNAMICS = {"SIS", "neuronal", "biomass"}
DEFAULT_SAMPLE_WINDOW_STEP = {"SIS": 50, "neuronal": 5, "biomass": 5, "SLBP": 10}
DEFAULT_SAMPLING_T = {"SIS": 0.1, "neuronal": 10, "biomass": 10, "SLBP": 100}

# The reference anchors relative paths (ews_results/..., dataset/...) at its repository root (:11).
PROJECT_ROOT = Path(os.environ.get("UPD_PROJECT_ROOT", os.getcwd())).resolve()
# Upper bound on device memory for one batched launch's trajectories (bytes).
SWEEP_BATCH_BYTES = int(os.environ.get("UPD_SWEEP_BATCH_BYTES", 8 << 30))


def set_project_root(path):
    global PROJECT_ROOT
    PROJECT_ROOT = Path(path).resolve()


def _as_path(path):
    """:39-42."""
    return None if path is None else Path(path)


def _resolve_project_path(path):
    if path is None:
        return None
    path = Path(path)
    return path if path.is_absolute() else PROJECT_ROOT / path


def _default_device(device=None):
    return device or torch.device("cuda" if torch.cuda.is_available() else "cpu")


# ------------------------------------------------------------------------------------------------
# data loading, sub-sampling and rolling windows (integer bookkeeping: bit-exact with :52-182)
# ------------------------------------------------------------------------------------------------
_CANONICAL = {"sis": "SIS", "slbp": "SLBP", "neuronal": "neuronal", "biomass": "biomass"}


def _dynamic_name(dynamic_type):
    if dynamic_type is None:
        return None
    text = str(dynamic_type)
    return _CANONICAL.get(text.lower(), text)


def _infer_dynamic_type(data_file=None, loaded_data=None):
    """:67-85."""
    if loaded_data is not None:
        if "N_values" in loaded_data:
            return "SLBP"
        if "tp_values" in loaded_data:
            return None
    if data_file is None:
        return None
    text = str(data_file).replace("\\", "/").lower()
    for needle, name in (("slbp", "SLBP"), ("sis", "SIS"), ("neuronal", "neuronal"), ("biomass", "biomass")):
        if needle in text:
            return name
    return None


def normalize_time_series(torch_time_series, dynamic_type=None):
    """:103-116: -> [Node, T, F]; 2-D network data is [T, Node], 2-D single series is [T, F]."""
    dynamic_type = _dynamic_name(dynamic_type)
    data = torch.as_tensor(torch_time_series).float()
    if data.ndim == 3:
        return data
    if data.ndim != 2:
        raise ValueError("time series must have shape [Node, T, F], [T, F], or [T, Node].")
    return data.t().unsqueeze(-1) if dynamic_type in NETWORK_DYNAMICS else data.unsqueeze(0)


def load_dynamic_data(data_file, dynamic_type=None, map_location="cpu"):
    """:88-100."""
    loaded = torch.load(data_file, map_location=map_location, weights_only=False)
    dynamic_type = _dynamic_name(dynamic_type) or _infer_dynamic_type(data_file=data_file, loaded_data=loaded)
    if "ys_dynamic" not in loaded or "ts_dynamic" not in loaded:
        raise KeyError("data_file must contain 'ys_dynamic' and 'ts_dynamic'.")
    return {"torch_time_series": normalize_time_series(loaded["ys_dynamic"], dynamic_type=dynamic_type),
            "time_data": loaded["ts_dynamic"], "dynamic_type": dynamic_type, "loaded_data": loaded}


def sampling_interval_from_t(sampling_t):
    """:119-125.  The float division is the reference's (int(0.3/0.1) == 2): same expression, same quirk."""
    sampling_t_min = 0.1
    if sampling_t is None or sampling_t <= sampling_t_min:
        return 1
    return max(1, int(sampling_t / sampling_t_min))


def sample_time_series(torch_time_series, time_data, sampling_t, return_numpy_time=True):
    """:128-134."""
    interval = sampling_interval_from_t(sampling_t)
    series = torch_time_series[:, ::interval, :]
    times = torch.as_tensor(time_data)[::interval]
    if return_numpy_time:
        times = times.cpu().detach().numpy()
    return series, times


def sliding_window_count(sampled_length, windows, sample_window_step):
    """:157-160."""
    if sampled_length < windows:
        return 0
    return (sampled_length - windows) // sample_window_step + 1


def stacked_sliding_windows(torch_time_series, windows, sample_window_step):
    """All rolling windows as one tensor [W, Node, windows, F] (a strided view, no copy)."""
    if torch_time_series.ndim != 3:
        raise ValueError("torch_time_series must have shape [Node_num, T_obs_num, F].")
    if torch_time_series.shape[1] < windows:
        raise ValueError("T_obs_num ({}) is shorter than windows ({}).".format(torch_time_series.shape[1], windows))
    return torch_time_series.unfold(1, windows, sample_window_step).permute(1, 0, 3, 2)


def build_sliding_windows(torch_time_series, time_data, windows, sample_window_step):
    """:137-147 -> (tuple of W tensors [Node, windows, F], time_points = t[windows-1::step])."""
    stacked = stacked_sliding_windows(torch_time_series, windows, sample_window_step)
    time_points = np.asarray(time_data)[windows - 1:: sample_window_step]
    return stacked.unbind(0), time_points


def default_sample_window_step(dynamic_type, task_model=None, dataset_config=None):
    """:150-154."""
    dataset_config = dataset_config or {}
    if task_model == "DiffSTG" and dataset_config.get("interval_step") is not None:
        return dataset_config["interval_step"]
    return DEFAULT_SAMPLE_WINDOW_STEP.get(dynamic_type, 10)


def infer_sample_window_step_from_cache(sampled_length, windows, cache_len, fallback_step):
    """:163-182: the step that reproduces ``cache_len`` windows; ties go to the step closest to the
    fallback, then to the larger step."""
    if cache_len <= 0 or sampled_length < windows:
        return fallback_step
    if sliding_window_count(sampled_length, windows, fallback_step) == cache_len or cache_len == 1:
        return fallback_step
    span = sampled_length - windows
    low = int(np.floor(span / cache_len)) + 1
    high = int(np.floor(span / (cache_len - 1)))
    best = None
    for step in range(max(1, low), max(1, high) + 1):
        if sliding_window_count(sampled_length, windows, step) == cache_len:
            key = (abs(step - fallback_step), -step)
            if best is None or key < best[0]:
                best = (key, step)
    return fallback_step if best is None else best[1]


def torch_data_preprocessing_like_slbp(time_data, sampling_t, return_numpy=False):
    """:486-491."""
    sampled = torch.as_tensor(time_data)[:: sampling_interval_from_t(sampling_t)]
    return sampled.cpu().detach().numpy() if return_numpy else sampled


def build_slbp_sensitivity_windows(torch_time_series, time_data, windows, pred_len, sampling_t, sample_window_step):
    """:467-483: SLBP series [T_raw, F] -> (inputs [windows,F] ..., targets [pred_len,F] ..., time points)."""
    series = torch_data_preprocessing_like_slbp(torch_time_series, sampling_t=sampling_t)
    sampled_time = torch_data_preprocessing_like_slbp(time_data, sampling_t=sampling_t, return_numpy=True)
    time_points = sampled_time[windows - 1:: sample_window_step]
    input_datas = series.unfold(0, windows, sample_window_step).permute(0, 2, 1).unbind(0)
    tail = series[windows:, :]
    pred_datas = ()
    if tail.shape[0] >= pred_len:
        pred_datas = tail.unfold(0, pred_len, sample_window_step).permute(0, 2, 1).unbind(0)
    return input_datas, pred_datas, time_points


# ------------------------------------------------------------------------------------------------
# model loading and cache files
# ------------------------------------------------------------------------------------------------
def read_model_config(model_save_file):
    """:185-190."""
    config_path = Path(model_save_file) / "model_trained.yaml"
    if not config_path.exists():
        raise FileNotFoundError("model config not found: {}".format(config_path))
    with open(config_path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def _train_model_select(method_config):
    if method_config.get("train") is not None:
        return method_config["train"].get("train_model_select")
    return None


def _load_checkpoint(model_path, device, infer_params, train_model_select):
    from .loader import load_diffusion_model

    old_cwd = Path.cwd()
    try:
        # the reference changes directory because TMDM opens tmdm.yml by relative path (:203-213);
        # kept so relative pretrain paths inside net_param resolve the same way
        if PROJECT_ROOT.exists():
            os.chdir(PROJECT_ROOT)
        model, loaded_net_param = load_diffusion_model(str(model_path), device=device, infer_para=infer_params,
                                                       train_model_select=train_model_select)
    finally:
        os.chdir(old_cwd)
    model.eval()
    return model, loaded_net_param


def load_model_from_dir(model_save_file, device=None, infer_params=None, method_config=None):
    """:193-215: ``<dir>/model_trained`` + ``<dir>/model_trained.yaml`` -> (model.eval(), net_param)."""
    model_save_file = Path(model_save_file)
    method_config = method_config or read_model_config(model_save_file)
    return _load_checkpoint(model_save_file / "model_trained", _default_device(device), infer_params,
                            _train_model_select(method_config))


def load_sensitivity_model(model_root, model_name, device=None, infer_params=None):
    """:425-455: ``<root>/models/<name>`` + ``.yaml`` -> (model, method_config, net_param)."""
    model_root = _resolve_project_path(model_root)
    config_path = model_root / "models" / "{}.yaml".format(model_name)
    model_path = model_root / "models" / model_name
    if not config_path.exists():
        raise FileNotFoundError("model config not found: {}".format(config_path))
    if not model_path.exists():
        raise FileNotFoundError("model checkpoint not found: {}".format(model_path))
    with open(config_path, "r", encoding="utf-8") as f:
        method_config = yaml.safe_load(f)
    model, net_param = _load_checkpoint(model_path, _default_device(device), infer_params,
                                        _train_model_select(method_config))
    return model, method_config, net_param


def default_cache_dir(model_save_file, dynamic_type):
    """:218-223."""
    if model_save_file is not None:
        return Path(model_save_file)
    return PROJECT_ROOT / "ews_results" / "model_uncertainy_cache" / "model" / (_dynamic_name(dynamic_type) or "unknown")


def data_cache_name(data_file, suffix=""):
    """:226-234."""
    if data_file is None:
        return "data{}.pt".format(suffix)
    p = Path(data_file)
    return "{}{}{}".format(p.stem, suffix, p.suffix or ".pt")


def resolve_cache_path(cache_path, model_save_file, data_file, dynamic_type, suffix=""):
    """:237-245."""
    if cache_path is None:
        return default_cache_dir(model_save_file, dynamic_type) / data_cache_name(data_file, suffix=suffix)
    cache_path = _resolve_project_path(cache_path)
    if cache_path.suffix == ".pt":
        return cache_path
    return cache_path / data_cache_name(data_file, suffix=suffix)


def resolve_figure_path(cache_file_path):
    return Path(cache_file_path).with_suffix(".png")


# Cache files are written by one background thread (at B200 rates the 0.15-0.7 GB ``torch.save`` of a
# sweep costs as much as sampling it).  A file appears atomically (temp file + rename) and every reader in this module
# waits for a pending write of the path it is about to open; ``flush_cache_writes()`` (also run at interpreter exit)
# waits for all of them and re-raises a failed write.  UPD_SYNC_CACHE_WRITES=1 restores the reference's blocking write.
_WRITER = None
_PENDING = {}
_PENDING_LOCK = threading.Lock()


def _write_tensor_list(data_list, cache_path):
    tmp = cache_path.with_name(cache_path.name + ".tmp{}".format(os.getpid()))
    with open(tmp, "wb") as f:
        torch.save(data_list, f)
    os.replace(tmp, cache_path)


def _save_tensor_list(data_list, cache_path):
    """:252-256: the cache is ``torch.save(list[Tensor])``."""
    global _WRITER
    cache_path = Path(cache_path)
    cache_path.parent.mkdir(parents=True, exist_ok=True)
    _await_write(cache_path)
    if os.environ.get("UPD_SYNC_CACHE_WRITES"):
        _write_tensor_list(data_list, cache_path)
        return
    with _PENDING_LOCK:
        if _WRITER is None:
            _WRITER = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="upd-cache-writer")
            atexit.register(flush_cache_writes)
        _PENDING[str(cache_path.resolve())] = _WRITER.submit(_write_tensor_list, data_list, cache_path)


def _await_write(cache_path):
    with _PENDING_LOCK:
        fut = _PENDING.pop(str(Path(cache_path).resolve()), None)
    if fut is not None:
        fut.result()


def flush_cache_writes():
    """Wait for every cache file queued by this process; raises the first write error."""
    with _PENDING_LOCK:
        futs = list(_PENDING.values())
        _PENDING.clear()
    for fut in futs:
        fut.result()


def _cache_ready(cache_path):
    """``Path.exists()`` that first waits for a pending background write of that file."""
    _await_write(cache_path)
    return Path(cache_path).exists()


def _load_tensor_list(cache_path):
    """:259-264."""
    _await_write(cache_path)
    with open(cache_path, "rb") as f:
        data_list = torch.load(f, map_location="cpu", weights_only=False)
    if not isinstance(data_list, list):
        raise TypeError("cache file must contain a list of tensors: {}".format(cache_path))
    return data_list


def slbp_sensitivity_cache_path(model_root, model_name, data_trend, sample_window_step=10):
    """:458-464."""
    return _resolve_project_path(model_root) / "datas" / "{}_pred_future_{}_{}.pt".format(
        model_name, data_trend, sample_window_step)


def slbp_fig6_cache_path(model_root, model_name, data_trend, sample_window_step=10, cache_subdir=None,
                         kind="pred_future"):
    """:624-634."""
    cache_dir = _resolve_project_path(model_root) / "datas"
    if cache_subdir:
        cache_dir = cache_dir / cache_subdir
    return cache_dir / "{}_{}_{}_{}.pt".format(model_name, kind, data_trend, sample_window_step)


def slbp_fig6_pred_future_gx_cache_path(model_root, model_name, data_trend, sample_window_step=10, cache_subdir=None):
    """:637-646."""
    cache_dir = _resolve_project_path(model_root) / "datas"
    if cache_subdir:
        cache_dir = cache_dir / cache_subdir
    return cache_dir / "{}_pred_future_{}_{}_gx.pt".format(model_name, data_trend, sample_window_step)


def _legacy_single_underscore_model_name(model_name):
    return str(model_name).replace("dataset__", "dataset_", 1)


def _read_slbp_fig6_model_config(model_root, model_name):
    config_path = _resolve_project_path(model_root) / "models" / "{}.yaml".format(model_name)
    if not config_path.exists():
        return None
    with open(config_path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def _slbp_fig6_dataset_config(model_root, model_name, windows=None, pred_len=None, sampling_t=None):
    """:662-670."""
    method_config = _read_slbp_fig6_model_config(model_root, model_name)
    ds = method_config.get("dataset", {}) if method_config else {}
    return {"windows": windows if windows is not None else ds.get("windows", 200),
            "pred_len": pred_len if pred_len is not None else ds.get("pred_len", 200),
            "sampling_t": sampling_t if sampling_t is not None else ds.get("sampling_t", 100),
            "method_config": method_config}


def _slbp_cache_elements_have_ndim(data_list, ndim):
    return bool(data_list) and all(torch.as_tensor(item).ndim == ndim for item in data_list)


def _slbp_cache_elements_are_gx(data_list):
    """:679-683: [O,F] elements, or [1,O,F]."""
    if _slbp_cache_elements_have_ndim(data_list, 2):
        return True
    return bool(data_list) and all(torch.as_tensor(i).ndim == 3 and torch.as_tensor(i).shape[0] == 1 for i in data_list)


def normalize_diffstg_pred_future_list(pred_future_list):
    """:354-366: legacy DiffSTG elements [Node,O,K] gain the feature axis."""
    out = []
    for pf in pred_future_list:
        pf = torch.as_tensor(pf).detach().cpu()
        if pf.ndim == 3:
            pf = pf.unsqueeze(-2)
        if pf.ndim != 4:
            raise ValueError("DiffSTG pred_future must have shape [Node, pred_len, F, samples] "
                             "or legacy [Node, pred_len, samples], got {}".format(tuple(pf.shape)))
        out.append(pf)
    return out


# ------------------------------------------------------------------------------------------------
# the hot loop: sampling sweeps
# ------------------------------------------------------------------------------------------------
def _model_device(model):
    return next(model.parameters()).device


def _scale_windows(model, stacked, device):
    """``model.scaler_transform(window.to(device))`` for every window at once (:332-334).  As in the
    reference the transform is applied whenever ``model.scaler`` is not None -- and it is a string, so it
    always is."""
    x = stacked.to(device, torch.float32)
    if getattr(model, "scaler", None) is not None:
        x = model.scaler_transform(x)
    return x


# Statistics computed on the device while a freshly sampled sweep is still resident, keyed by the host
# cache's storage address, so summarize_*() right after run_*_cache() does not upload the cache again.  An entry is
# only trusted while it provably describes the tensors handed to summarize_*: it holds a weak reference to the cache
# (a freed block reused by another tensor misses), the cache's version counter (any in-place edit of an element
# misses: views share the counter) and the scaler table the "raw" statistics were baked with.
_FRESH_STATS = {}


def _scaler_snapshot(model):
    if model is None or not hasattr(model, "scaler_mean") or not hasattr(model, "scaler_std"):
        return None
    return (tuple(model.scaler_mean.detach().float().cpu().tolist()), tuple(model.scaler_std.detach().float().cpu().tolist()))


def partition_windows(n_windows, world_size, rank):
    """Contiguous block of windows owned by ``rank``: [ceil-balanced) so that concatenating the ranks'
    caches in rank order restores window order."""
    per = -(-n_windows // world_size)
    w0 = min(n_windows, rank * per)
    return w0, min(n_windows, w0 + per)


def sample_sweep(model, stacked_windows, device=None, pin=True, reduce=True, window_offset=0, graph_data=None):
    """Sample every window of a sweep.  ``stacked_windows`` [W, B, L, F] in raw units (CPU or device).
    Returns the prediction cache as one CPU tensor [W, B, K, O, F] (pinned); element w of the reference's
    list is ``cache[w].permute(0, 2, 3, 1)``.  Work is cut into launches of at most SWEEP_BATCH_BYTES.
    With ``reduce`` the per-window MPV / mean statistics are computed on the device in the same pass
    (both in scaled units and, when the model has a scaler, in raw units) and remembered for the
    summarize_* functions; they are also returned as ``cache.upd_stats`` (dict of CPU tensors).
    ``graph_data`` (DiffSTG, NsDiff_spatial): object with ``edge_index`` / ``num_nodes``; every window, sampling round
    and parallel replica is then a replica of that graph inside one launch (:369-391)."""
    device = device or _model_device(model)
    W, B = stacked_windows.shape[0], stacked_windows.shape[1]
    if graph_data is not None and hasattr(model, "parallel_sampling"):
        probe_k = int(model.parallel_sampling) * int(model.sequential_sampling)
        O, F = model.T_p, model.F
    else:                       # NsDiff / TMDM / DiffusionTS, and NsDiff_spatial when ``graph_data`` is given
        probe_k = _samples_per_row(model)
        O, F = model.pred_len, model.dataset_nf
    per_window = B * probe_k * O * F * 4
    step = max(1, min(W, SWEEP_BATCH_BYTES // max(per_window, 1)))
    # The pinned host allocation (~0.2-0.5 ms/MB) is made after the first launch has been enqueued, so it overlaps the GPU
    # work.  (Allocating it on a helper thread in parallel with the enqueue was measured slower: cudaHostAlloc holds
    # driver locks that stall the kernel launches of the main thread -- e2e 1.44 -> 1.27 M traj/s on the bench.)
    cache = None
    base = getattr(model, "_windows_drawn", 0) + window_offset
    scale = _scaler_table(model) if reduce else None
    parts = {"scaled": [], "raw": []}
    for w0 in range(0, W, step):
        w1 = min(W, w0 + step)
        x = _scale_windows(model, stacked_windows[w0:w1], device)
        if graph_data is not None:
            traj = model.sample_windows(x, graph_data.edge_index, int(graph_data.num_nodes), window_base=base + w0)
            traj = traj[:, :, -O:, :].contiguous()
        else:
            traj = model.sample_windows(x, window_base=base + w0)
        if cache is None:
            cache = torch.empty((W, B, probe_k, O, F), dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
        cache[w0:w1].copy_(traj.view(w1 - w0, B, probe_k, O, F), non_blocking=True)
        if reduce:
            parts["scaled"].append(kernels.mpv_reduce(traj, w1 - w0, B, want_mean=True))
            if scale is not None:
                parts["raw"].append(kernels.mpv_reduce(traj, w1 - w0, B, scale=scale))
    if cache is None:       # empty sweep
        cache = torch.empty((W, B, probe_k, O, F), dtype=torch.float32)
    if hasattr(model, "_windows_drawn"):
        model._windows_drawn = base - window_offset + W
    stats = {}
    for name, lst in parts.items():
        if lst:
            stats[name] = {k: torch.cat([r[k] for r in lst]).cpu() for k in lst[0]}
    torch.cuda.current_stream(device).synchronize()
    if reduce:
        if len(_FRESH_STATS) > 8:
            _FRESH_STATS.clear()
        _FRESH_STATS[cache.untyped_storage().data_ptr()] = (weakref.ref(cache), cache._version, tuple(cache.shape),
                                                            _scaler_snapshot(model), stats)
    cache.upd_stats = stats
    return cache


def _fresh_stats_for(pred_future_list, elem_shape, model=None, need_scaler=False):
    """Stats remembered by sample_sweep if ``pred_future_list`` is exactly that sweep's list of views, unmodified
    (and, for the raw-unit statistics, if ``model`` still carries the scaler they were computed with)."""
    if not pred_future_list or not isinstance(pred_future_list[0], torch.Tensor):
        return None
    key = pred_future_list[0].untyped_storage().data_ptr()
    hit = _FRESH_STATS.get(key)
    if hit is None:
        return None
    ref, version, shape, scaler, stats = hit
    owner = ref()
    if owner is None or owner.untyped_storage().data_ptr() != key or pred_future_list[0]._version != version:
        _FRESH_STATS.pop(key, None)
        return None
    if need_scaler and scaler != _scaler_snapshot(model):
        return None
    W, B, K, O, F = shape
    if len(pred_future_list) != W or tuple(elem_shape) != (B, O, F, K):
        return None
    for w, t in enumerate(pred_future_list):
        if t.storage_offset() != w * B * K * O * F or t.numel() != B * K * O * F:
            return None
    return stats


def _samples_per_row(model):
    """Samples a model's evaluation_step returns per row, with the reference's own failure modes: the NsDiff classes loop
    ``range(n_z_samples // parallel_sample)`` (NsDiff_model.py:227 / :461 / :742: the remainder is dropped, and an empty
    loop fails in torch.cat); TMDM and DiffusionTS clamp parallel_sample to n_z_samples and insist on divisibility
    (tmdm_adapter.py:125-127, DiffusionTS_model.py:83-85)."""
    cfg = model.configs
    S = int(getattr(cfg, "parallel_sample", 1))
    K = int(getattr(cfg, "n_z_samples", 1))
    if type(model).__name__.startswith("NsDiff"):
        if S <= 0 or K // S == 0:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors (n_z_samples={} // parallel_sample={} "
                               "chunks)".format(K, S))
        return (K // S) * S
    S = min(S, K)
    if S <= 0 or K % S != 0:
        raise ValueError("n_z_samples must be divisible by parallel_sample")
    return K


def _as_cache_list(cache, squeeze_rows=False):
    """[W,B,K,O,F] -> list of W views shaped like the reference's elements ([B,O,F,K] or SLBP [O,F,K])."""
    out = []
    for w in range(cache.shape[0]):
        el = cache[w].permute(0, 2, 3, 1)
        out.append(el.squeeze(0) if squeeze_rows else el)
    return out


def run_evaluation_cache(model, timeseries_datas, pred_len, cache_path, device, force_recompute=False,
                         max_windows=None):
    """:323-339: read the cache, or sample every window (batched on the GPU) and write it."""
    cache_path = Path(cache_path)
    if _cache_ready(cache_path) and not force_recompute:
        return _load_tensor_list(cache_path)
    iterable = timeseries_datas[:max_windows] if max_windows is not None else timeseries_datas
    stacked = torch.stack([torch.as_tensor(w) for w in iterable])
    cache = sample_sweep(model, stacked, device=device)
    pred_future_list = [el[:, -pred_len:, :, :] for el in _as_cache_list(cache)]
    _save_tensor_list(pred_future_list, cache_path)
    return pred_future_list


def run_slbp_sensitivity_cache(model, input_datas, cache_path, device, force_recompute=False, max_windows=None):
    """:502-526: SLBP windows [L,F] -> list of [O,F,K]; a corrupt cache is recomputed (:494-499)."""
    cache_path = Path(cache_path)
    if _cache_ready(cache_path) and not force_recompute:
        cached = read_sensitivity_pred_future_cache(cache_path)
        if cached is not None:
            return cached
    iterable = input_datas[:max_windows] if max_windows is not None else input_datas
    stacked = torch.stack([torch.as_tensor(w) for w in iterable]).unsqueeze(1)      # [W,1,L,F]
    cache = sample_sweep(model, stacked, device=device)
    pred_future_list = _as_cache_list(cache, squeeze_rows=True)
    _save_tensor_list(pred_future_list, cache_path)
    return pred_future_list


def read_sensitivity_pred_future_cache(cache_path):
    try:
        return _load_tensor_list(cache_path)
    except Exception as exc:  # noqa: BLE001 - the reference swallows any read error and recomputes
        print("warning: failed to read cache {}, recomputing ({})".format(cache_path, exc))
        return None


def _gx_sweep(model, stacked, device):
    """g(x) for every window row of [W,B,L,F] in one launch -> CPU [W,B,O,F]."""
    W, B = stacked.shape[0], stacked.shape[1]
    x = _scale_windows(model, stacked, device)
    with torch.no_grad():
        gx = model.cond_pred_model_g(x.reshape(W * B, x.shape[2], x.shape[3])[:, :model.windows, :].contiguous())
    return gx.view(W, B, gx.shape[1], gx.shape[2]).cpu()


def run_nsdiff_g_cache(model, timeseries_datas, cache_path, device, pred_dim=0, force_recompute=False,
                       max_windows=None):
    """:400-422 -> list of [Node,O,F] (None when the model has no g(x))."""
    cache_path = Path(cache_path)
    if _cache_ready(cache_path) and not force_recompute:
        return _load_tensor_list(cache_path)
    if not hasattr(model, "cond_pred_model_g") or model.cond_pred_model_g is None:
        return None
    iterable = timeseries_datas[:max_windows] if max_windows is not None else timeseries_datas
    stacked = torch.stack([torch.as_tensor(w) for w in iterable])
    if pred_dim >= model.dataset_nf:
        raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, model.dataset_nf))
    g_list = list(_gx_sweep(model, stacked, device).unbind(0))
    _save_tensor_list(g_list, cache_path)
    return g_list


def run_slbp_gx_cache_for_fig6(model, input_datas, cache_path, device, pred_dim=0, force_recompute=False,
                               max_windows=None):
    """:731-765 -> list of [O,F]."""
    cache_path = Path(cache_path)
    if _cache_ready(cache_path) and not force_recompute:
        gx_list = _load_tensor_list(cache_path)
        if _slbp_cache_elements_are_gx(gx_list):
            return gx_list
    if not hasattr(model, "cond_pred_model_g") or model.cond_pred_model_g is None:
        raise ValueError("model does not provide cond_pred_model_g for gx generation.")
    iterable = input_datas[:max_windows] if max_windows is not None else input_datas
    stacked = torch.stack([torch.as_tensor(w) for w in iterable]).unsqueeze(1)
    if pred_dim >= model.dataset_nf:
        raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, model.dataset_nf))
    gx_list = [g.squeeze(0) for g in _gx_sweep(model, stacked, device).unbind(0)]
    _save_tensor_list(gx_list, cache_path)
    return gx_list


def real_data_gx_uncertainty(model, torch_model_time_series, time_data, windows, sampling_t, sample_window_step,
                             pred_dim, cache_path, device=None):
    """The model part of ``run_model_uncertainty`` (evaluation_and_analysis/real_data_analysis.py:331-348; SURVEY 8f
    row 3): sub-sample the prepared series [Node, T, F] (torch_data_preprocessing, :200-205), cut rolling windows
    ``unfold(1, windows, step)``, evaluate g(x) on every window -- here ALL windows in one ``upd_sigma_estimation`` launch
    instead of a Python loop -- and reduce ``gx.squeeze(-1).mean(-1)[pred_dim]`` per window.  As in that function the
    scaler is applied only when ``model.scaler == "StandardScaler"`` (:338, unlike run_evaluation_cache).  Writes the
    cache ``list[Tensor [Node, O]]`` to ``cache_path`` and returns (model_times[:n], values float64 [n])."""
    device = device or _model_device(model)
    if not hasattr(model, "cond_pred_model_g") or model.cond_pred_model_g is None:
        raise ValueError("model does not provide cond_pred_model_g for gx generation.")
    sampling_interval = int(sampling_t / 0.1) if sampling_t > 0.1 else 1          # :201-202
    sampled_series = torch.as_tensor(torch_model_time_series)[:, ::sampling_interval, :]
    sampled_time = torch.as_tensor(time_data)[::sampling_interval].detach().cpu().numpy()
    stacked = sampled_series.unfold(1, windows, sample_window_step).permute(1, 0, 3, 2).contiguous()   # [W, Node, L, F]
    model_times = sampled_time[windows - 1:: sample_window_step]
    W, B = stacked.shape[0], stacked.shape[1]
    x = stacked.to(device, torch.float32)
    if getattr(model, "scaler", None) == "StandardScaler":
        x = model.scaler_transform(x)
    with torch.no_grad():
        gx = model.cond_pred_model_g(x.reshape(W * B, x.shape[2], x.shape[3])[:, :model.windows, :].contiguous())
    gx = gx.view(W, B, gx.shape[1], gx.shape[2]).squeeze(-1)                      # [W, Node, O] for F = 1
    if gx.dim() != 3:
        # the reference does float(gx.mean(-1)[pred_dim]) per window, which only works for a single feature
        raise TypeError("only length-1 arrays can be converted to Python scalars")
    values = gx.mean(dim=-1)[:, pred_dim].double().cpu().numpy()
    data_save_list = list(gx.cpu().unbind(0))
    _save_tensor_list(data_save_list, Path(cache_path))
    return model_times[: len(values)], np.asarray(values, dtype=float)


def load_diffstg_graph(graph_file):
    """:342-351: graphml -> graph object with ``edge_index`` [2,E] int64 and ``num_nodes``.  The reference goes through
    networkx + torch_geometric.utils.from_networkx; the same edge order is produced here without torch_geometric:
    nodes relabelled 0..V-1 in file order, an undirected graph contributes both directions, edges listed per source
    node in adjacency (insertion) order -- what ``list(G.to_directed().edges)`` yields."""
    import networkx as nx
    from .diffstg import GraphData
    if graph_file is None:
        raise ValueError("graph_file is required for DiffSTG.")
    graph_file = _resolve_project_path(graph_file)
    nx_g = nx.read_graphml(graph_file)
    nx_g = nx.convert_node_labels_to_integers(nx_g)
    directed = nx_g if nx.is_directed(nx_g) else nx_g.to_directed()
    edges = list(directed.edges)
    edge_index = torch.tensor(edges, dtype=torch.long).t().contiguous() if edges else torch.zeros((2, 0), dtype=torch.long)
    return GraphData(x=None, edge_index=edge_index.view(2, -1), num_nodes=nx_g.number_of_nodes())


def run_diffstg_evaluation_cache(model, timeseries_datas, pred_len, graph_data, cache_path, device,
                                 force_recompute=False, max_windows=None):
    """:369-397: read the cache, or sample every window on the graph (all windows batched as graph replicas)."""
    cache_path = Path(cache_path)
    if _cache_ready(cache_path) and not force_recompute:
        return normalize_diffstg_pred_future_list(_load_tensor_list(cache_path))
    iterable = timeseries_datas[:max_windows] if max_windows is not None else timeseries_datas
    stacked = torch.stack([torch.as_tensor(w) for w in iterable])
    cache = sample_sweep(model, stacked, device=device, graph_data=graph_data)
    pred_future_list = [el[:, -pred_len:, :, :] for el in _as_cache_list(cache)]
    _save_tensor_list(pred_future_list, cache_path)
    return pred_future_list


# ------------------------------------------------------------------------------------------------
# reductions (device Welford kernel)
# ------------------------------------------------------------------------------------------------
def _reduce_elements(pred_future_list, scale=None, want_mean=False):
    """List of [B,O,F,K] elements (any strides) -> per-window statistics on the GPU, one launch per run of
    equal-shaped elements.  Returns dict of CPU tensors mpv [W], pred_mean [W], mpv_f [W,F] (+ mean)."""
    dev = scale.device if scale is not None else torch.device("cuda", torch.cuda.current_device())   # the scaler table lives on the model's device
    outs = {"mpv": [], "pred_mean": [], "mpv_f": [], "mean": []}
    i, n = 0, len(pred_future_list)
    while i < n:
        shape = tuple(pred_future_list[i].shape)
        j = i
        while j < n and tuple(pred_future_list[j].shape) == shape:
            j += 1
        B, O, F, K = shape
        # [B,O,F,K] views of [B,K,O,F] storage permute back for free; anything else is copied once
        traj = torch.stack([t.permute(0, 3, 1, 2) for t in pred_future_list[i:j]]).to(dev, torch.float32)
        traj = traj.reshape((j - i) * B, K, O, F).contiguous()
        r = kernels.mpv_reduce(traj, j - i, B, scale=scale, want_mean=want_mean)
        outs["mpv"].append(r["mpv"].cpu())
        outs["pred_mean"].append(r["pred_mean"].cpu())
        outs["mpv_f"].append(r["mpv_f"].cpu())
        if want_mean:
            outs["mean"].extend(r["mean"].view(j - i, B, O, F).cpu().unbind(0))
        i = j
    res = {k: torch.cat(v) for k, v in outs.items() if k != "mean" and v}
    res["mean"] = outs["mean"]
    return res


def _stats_scaled(elems, want_mean=False):
    """Statistics in the cache's own (scaled) units: remembered from the sweep when fresh, else reduced now."""
    fresh = _fresh_stats_for(elems, elems[0].shape)
    if fresh is not None and "scaled" in fresh:
        r = dict(fresh["scaled"])
        B, O, F, _ = elems[0].shape
        r["mean"] = list(r["mean"].view(len(elems), B, O, F).unbind(0))
        return r
    return _reduce_elements(elems, want_mean=want_mean)


def _scaler_table(model):
    """(mean,std) rows for the inverse transform the reference applies when a model object with a scaler is
    present (_feature_inverse_transform, :267-283); None otherwise."""
    if model is None or getattr(model, "scaler", None) is None:
        return None
    if not hasattr(model, "scaler_mean") or not hasattr(model, "scaler_std"):
        return None
    dev = _model_device(model)                      # the model's device, which need not be the current one
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    return torch.stack([model.scaler_mean.detach().float().cpu(), model.scaler_std.detach().float().cpu()]).to(dev).contiguous()


def _feature_inverse_transform(pred_future, model=None):
    """:267-283, for callers that post-process a cache element themselves (the summaries above apply the same
    (mean, std) table inside the reduction kernel): back to raw units along the feature axis, when a model with a
    scaler is present."""
    if model is None or getattr(model, "scaler", None) is None:
        return pred_future
    if not hasattr(model, "scaler_mean") or not hasattr(model, "scaler_std"):
        if hasattr(model, "scaler_inverse_transform"):
            return model.scaler_inverse_transform(pred_future)
        return pred_future
    mean = model.scaler_mean.detach().to(pred_future.device, pred_future.dtype)
    std = model.scaler_std.detach().to(pred_future.device, pred_future.dtype)
    if pred_future.ndim >= 3 and pred_future.shape[-2] == mean.numel():
        shape = [1] * pred_future.ndim
        shape[-2] = mean.numel()
        return pred_future * std.view(*shape) + mean.view(*shape)
    if pred_future.shape[-1] == mean.numel():
        return pred_future * std + mean
    return pred_future


def _np_scalars(t):
    return [np.asarray(v, dtype=np.float32).reshape(()) for v in t.tolist()]


def summarize_pred_future_list(pred_future_list, model=None):
    """:286-303 -> (pred_mean_list, uncertainty_ews_list) of numpy 0-d float32."""
    elems = []
    for pf in pred_future_list:
        pf = torch.as_tensor(pf).detach()
        if pf.ndim == 3:
            pf = pf.unsqueeze(0)
        if pf.ndim != 4:
            raise ValueError("pred_future must have shape [Node, pred_len, F, n_z_samples], got {}".format(
                tuple(pf.shape)))
        elems.append(pf)
    if not elems:
        return [], []
    scale = _scaler_table(model)
    if scale is not None and scale.shape[1] != elems[0].shape[-2]:
        scale = None
    fresh = _fresh_stats_for(elems, elems[0].shape, model=model, need_scaler=scale is not None)
    key = "raw" if scale is not None else "scaled"
    if fresh is not None and key in fresh:
        r = fresh[key]
    else:
        r = _reduce_elements(elems, scale=scale)
    return _np_scalars(r["pred_mean"]), _np_scalars(r["mpv"])


def summarize_nsdiff_g_list(g_list, pred_dim=0):
    """:306-320: gx-EWS = mean over rows of the time-mean of feature pred_dim."""
    ews, pmean = [], []
    for gx in g_list:
        gx = torch.as_tensor(gx).detach().float()
        if gx.ndim == 2:
            gx = gx.unsqueeze(0)
        if gx.ndim != 3:
            raise ValueError("NsDiff-g cache elements must have shape [Node, pred_len, F].")
        if pred_dim >= gx.shape[-1]:
            raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, gx.shape[-1]))
        g = gx.cuda()
        ews.append(g.mean(dim=1)[:, pred_dim].mean().cpu().numpy())
        pmean.append(g.mean().cpu().numpy())
    return pmean, ews


def summarize_slbp_sensitivity(pred_future_list, pred_datas, model=None, device=None, pred_dim=0):
    """:529-550 -> (mpv_list, prediction_error_list)."""
    elems = []
    for pf in pred_future_list:
        pf = torch.as_tensor(pf).detach()
        if pf.ndim != 3:
            raise ValueError("SLBP sensitivity cache elements must have shape [pred_len, F, n_z_samples].")
        if pred_dim >= pf.shape[1]:
            raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, pf.shape[1]))
        elems.append(pf.unsqueeze(0))
    if not elems:
        return [], []
    r = _stats_scaled(elems, want_mean=True)
    mpv_list = _np_scalars(r["mpv_f"][:, pred_dim])
    # prediction error: all windows in one launch on the Welford means (upd_prediction_error), no per-window upload
    n = min(len(elems), len(pred_datas))
    if n == 0:
        return mpv_list, []
    dev = torch.device("cuda", torch.cuda.current_device()) if model is None else _model_device(model)
    with torch.cuda.device(dev):
        target = torch.stack([torch.as_tensor(t).detach().float() for t in pred_datas[:n]]).to(dev)          # [n, O, F]
        if getattr(model, "scaler", None) is not None and hasattr(model, "scaler_transform"):
            target = model.scaler_transform(target)
        mean = torch.stack([m.reshape(target.shape[1:]) for m in r["mean"][:n]]).to(dev)
        err = torch.empty((n, target.shape[2]), dtype=torch.float32, device=dev)
        rc = _lib.lib().upd_prediction_error(_lib.ptr(mean.contiguous()), _lib.ptr(target.contiguous()), n, target.shape[1],
                                             target.shape[2], _lib.ptr(err), _lib.stream_ptr(dev))
        _lib.check(rc, "upd_prediction_error")
    return mpv_list, _np_scalars(err[:, pred_dim].cpu())


def _slbp_intrinsic_dimension(trajectories):
    """:686-698: number of principal components of the K x (O*F) sample cloud reaching 80 % of the
    variance.  Uses the K x K Gram matrix (same non-zero spectrum as the (O*F)^2 covariance) in float64."""
    traj = torch.as_tensor(trajectories, dtype=torch.float32)
    if traj.ndim != 2 or traj.shape[0] < 2:
        return np.nan
    t = traj.cuda().double()
    c = t - t.mean(dim=0, keepdim=True)
    gram = c @ c.T / max(traj.shape[0] - 1, 1)
    ev = torch.linalg.eigvalsh(gram).flip(0).clamp_min(0)
    total = ev.sum()
    if float(total) <= 0:
        return np.nan
    cum = torch.cumsum(ev / total, dim=0)
    return int(torch.where(cum >= 0.8)[0][0].item() + 1)


def _slbp_intrinsic_dimensions(elems, energy=0.8):
    """:686-698 for a whole sweep: every window's centred K x K Gram matrix in one launch (upd_gram_centered, double),
    one batched symmetric eigenvalue solve, and the count of leading eigenvalues reaching ``energy`` of the trace.
    elems: list of [1, O, F, K] cache elements.  -> list of int (nan where the reference returns nan)."""
    W = len(elems)
    O, F, K = elems[0].shape[1:]
    if K < 2:
        return [np.nan] * W
    dev = torch.device("cuda", torch.cuda.current_device())
    # [W, K, O*F]: a fresh sweep's elements are permuted views of one contiguous [W,1,K,O,F] cache, so this is a reshape
    traj = torch.stack([e[0].permute(2, 0, 1) for e in elems]).reshape(W, K, O * F).to(dev, torch.float32).contiguous()
    gram = torch.empty((W, K, K), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().upd_gram_centered(_lib.ptr(traj), W, K, O * F, ctypes.c_void_p(gram.data_ptr()), _lib.stream_ptr(dev))
        _lib.check(rc, "upd_gram_centered")
        ev = torch.linalg.eigvalsh(gram).flip(-1).clamp_min(0)                  # descending
        total = ev.sum(dim=-1, keepdim=True)
        cum = torch.cumsum(ev / total.clamp_min(1e-300), dim=-1)
        count = (cum < energy).sum(dim=-1) + 1
    count, total = count.cpu().tolist(), total.squeeze(-1).cpu().tolist()
    return [int(c) if t > 0 else np.nan for c, t in zip(count, total)]


def summarize_slbp_sampling_for_fig6(pred_future_list, pred_dim=0):
    """:701-714 -> (mpv list of float, intrinsic-dimension list of int)."""
    elems = []
    for pf in pred_future_list:
        pf = torch.as_tensor(pf).detach()
        if pf.ndim != 3:
            raise ValueError("SLBP sampling cache elements must have shape [pred_len, F, n_z_samples].")
        if pred_dim >= pf.shape[1]:
            raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, pf.shape[1]))
        elems.append(pf.unsqueeze(0))
    if not elems:
        return [], []
    r = _stats_scaled(elems)
    mpv = [float(v) for v in r["mpv_f"][:, pred_dim].tolist()]
    if elems[0].shape[-1] <= 128:
        dims = _slbp_intrinsic_dimensions(elems)
    else:                                   # more samples than the Gram kernel is built for: one window at a time
        dims = [_slbp_intrinsic_dimension(pf.squeeze(0).permute(2, 0, 1).reshape(pf.shape[-1], -1)) for pf in elems]
    return mpv, dims


def summarize_slbp_gx_for_fig6(gx_list, pred_dim=0):
    """:717-728."""
    out = []
    for gx in gx_list:
        gx = torch.as_tensor(gx).detach().float()
        if gx.ndim == 3 and gx.shape[0] == 1:
            gx = gx.squeeze(0)
        if gx.ndim != 2:
            raise ValueError("SLBP gx cache elements must have shape [pred_len, F] or [1, pred_len, F].")
        if pred_dim >= gx.shape[-1]:
            raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, gx.shape[-1]))
        out.append(float(gx.cuda()[:, pred_dim].mean().cpu().numpy()))
    return out


def summarize_slbp_mpv_cache_for_fig5(data_list, pred_dim=0):
    """:1026-1036: the cache kind is sniffed from the element rank."""
    if _slbp_cache_elements_have_ndim(data_list, 3):
        mpv, _ = summarize_slbp_sampling_for_fig6(data_list, pred_dim=pred_dim)
        return mpv, "sampling"
    if _slbp_cache_elements_are_gx(data_list):
        return summarize_slbp_gx_for_fig6(data_list, pred_dim=pred_dim), "gx"
    shape = tuple(torch.as_tensor(data_list[0]).shape) if data_list else None
    raise ValueError("Unsupported SLBP MPV cache element shape: {}".format(shape))


# ------------------------------------------------------------------------------------------------
# SLBP family (single-node series; :553-1250)
# ------------------------------------------------------------------------------------------------
def slbp_sensitivity_ews(model_root, model_name, torch_time_series, time_data, data_trend="increase", pred_dim=0,
                         sample_window_step=10, infer_params=None, force_recompute=False, max_windows=None,
                         device=None):
    """:553-621."""
    model_root = _resolve_project_path(model_root)
    device = _default_device(device)
    model, method_config, loaded_net_param = load_sensitivity_model(model_root, model_name, device=device,
                                                                    infer_params=infer_params)
    ds = method_config.get("dataset", {})
    windows, pred_len, sampling_t = ds["windows"], ds["pred_len"], ds["sampling_t"]
    input_datas, pred_datas, time_points = build_slbp_sensitivity_windows(
        torch_time_series, time_data, windows, pred_len, sampling_t, sample_window_step)
    cache_path = slbp_sensitivity_cache_path(model_root, model_name, data_trend, sample_window_step)
    pred_future_list = run_slbp_sensitivity_cache(model, input_datas, cache_path, device, force_recompute, max_windows)
    mpv_list, err_list = summarize_slbp_sensitivity(pred_future_list, pred_datas, model=model, device=device,
                                                    pred_dim=pred_dim)
    return {"time_points": time_points, "mpv": mpv_list, "prediction_error": err_list,
            "pred_future_list": pred_future_list, "cache_path": str(cache_path), "windows": windows,
            "pred_len": pred_len, "sampling_t": sampling_t, "sample_window_step": sample_window_step,
            "model_root": str(model_root), "model_name": model_name, "loaded_net_param": loaded_net_param}


def slbp_sampling_analysis(model_root, model_name, torch_time_series, time_data, data_trend="increase", pred_dim=0,
                           sample_window_step=10, cache_subdir=None, windows=None, pred_len=None, sampling_t=None,
                           infer_params=None, force_recompute=False, max_windows=None, device=None,
                           allow_unavailable=True):
    """:768-885: MPV + intrinsic dimension; falls back from ``*_pred_future_*`` to ``*_sampling_pred_future_*``
    when the former holds a gx cache; with ``allow_unavailable`` any failure yields available=False + reason."""
    cfg = _slbp_fig6_dataset_config(model_root, model_name, windows=windows, pred_len=pred_len, sampling_t=sampling_t)
    input_datas, _, time_points = build_slbp_sensitivity_windows(
        torch_time_series, time_data, cfg["windows"], cfg["pred_len"], cfg["sampling_t"], sample_window_step)
    cache_path = slbp_fig6_cache_path(model_root, model_name, data_trend, sample_window_step, cache_subdir, "pred_future")
    sampling_cache_path = slbp_fig6_cache_path(model_root, model_name, data_trend, sample_window_step, cache_subdir,
                                               "sampling_pred_future")
    base = {"windows": cfg["windows"], "pred_len": cfg["pred_len"], "sampling_t": cfg["sampling_t"],
            "sample_window_step": sample_window_step}
    try:
        active, pred_future_list = cache_path, None
        if not force_recompute:
            if _cache_ready(cache_path):
                cand = _load_tensor_list(cache_path)
                if _slbp_cache_elements_have_ndim(cand, 3):
                    pred_future_list = cand
                else:
                    active = sampling_cache_path
            if pred_future_list is None and _cache_ready(sampling_cache_path):
                cand = _load_tensor_list(sampling_cache_path)
                if not _slbp_cache_elements_have_ndim(cand, 3):
                    raise ValueError("sampling cache exists but is not [pred_len, F, n_z_samples]: {}".format(
                        sampling_cache_path))
                pred_future_list, active = cand, sampling_cache_path
        if pred_future_list is None:
            if _cache_ready(cache_path) and active == cache_path:
                active = sampling_cache_path
            device = _default_device(device)
            model, _, _ = load_sensitivity_model(model_root, model_name, device=device, infer_params=infer_params)
            pred_future_list = run_slbp_sensitivity_cache(model, input_datas, active, device, force_recompute, max_windows)
            if not _slbp_cache_elements_have_ndim(pred_future_list, 3):
                raise ValueError("generated sampling cache is not [pred_len, F, n_z_samples]: {}".format(active))
        mpv, dims = summarize_slbp_sampling_for_fig6(pred_future_list, pred_dim=pred_dim)
        return dict(base, available=True, time_points=time_points[: len(mpv)], mpv=mpv, intrinsic_dimension=dims,
                    pred_future_list=pred_future_list, cache_path=str(active), reason="")
    except Exception as exc:  # noqa: BLE001 - reference behaviour
        if not allow_unavailable:
            raise
        return dict(base, available=False, time_points=time_points, mpv=[], intrinsic_dimension=[],
                    pred_future_list=None, cache_path=str(sampling_cache_path if _cache_ready(cache_path) else cache_path),
                    reason=str(exc))


def slbp_gx_analysis(model_root, model_name, torch_time_series, time_data, data_trend="increase", pred_dim=0,
                     sample_window_step=10, cache_subdir=None, windows=None, pred_len=None, sampling_t=None,
                     infer_params=None, force_recompute=False, max_windows=None, device=None):
    """:888-1023: gx cache lookup order ``*_pred_future_*_gx.pt`` -> ``*_gx_*`` -> ``*_pred_future_*`` (if it holds
    gx) -> the same with the legacy single-underscore model name -> compute."""
    cfg = _slbp_fig6_dataset_config(model_root, model_name, windows=windows, pred_len=pred_len, sampling_t=sampling_t)
    input_datas, _, time_points = build_slbp_sensitivity_windows(
        torch_time_series, time_data, cfg["windows"], cfg["pred_len"], cfg["sampling_t"], sample_window_step)
    gx_cache_path = slbp_fig6_pred_future_gx_cache_path(model_root, model_name, data_trend, sample_window_step, cache_subdir)
    old_gx = slbp_fig6_cache_path(model_root, model_name, data_trend, sample_window_step, cache_subdir, "gx")
    legacy = slbp_fig6_cache_path(model_root, model_name, data_trend, sample_window_step, cache_subdir, "pred_future")
    legacy_name = slbp_fig6_cache_path(model_root, _legacy_single_underscore_model_name(model_name), data_trend,
                                       sample_window_step, cache_subdir, "pred_future")

    def compute():
        dev = _default_device(device)
        model, _, _ = load_sensitivity_model(model_root, model_name, device=dev, infer_params=infer_params)
        return run_slbp_gx_cache_for_fig6(model, input_datas, gx_cache_path, dev, pred_dim, force_recompute, max_windows)

    gx_list = None
    if not force_recompute:
        if _cache_ready(gx_cache_path):
            gx_list = _load_tensor_list(gx_cache_path)
        elif _cache_ready(old_gx):
            gx_list = _load_tensor_list(old_gx)
            if not _slbp_cache_elements_are_gx(gx_list):
                raise ValueError("gx cache exists but is not a gx cache: {}".format(old_gx))
            gx_cache_path = old_gx
        else:
            for cand_path in (legacy, legacy_name):
                if _cache_ready(cand_path):
                    cand = _load_tensor_list(cand_path)
                    if _slbp_cache_elements_are_gx(cand):
                        gx_list, gx_cache_path = cand, cand_path
                    else:
                        gx_list = compute()
                    break
    if gx_list is None:
        gx_list = compute()
    gx_mpv = summarize_slbp_gx_for_fig6(gx_list, pred_dim=pred_dim)
    return {"time_points": time_points[: len(gx_mpv)], "gx_mpv": gx_mpv, "gx_list": gx_list,
            "cache_path": str(gx_cache_path), "windows": cfg["windows"], "pred_len": cfg["pred_len"],
            "sampling_t": cfg["sampling_t"], "sample_window_step": sample_window_step}


def slbp_mpv_analysis(model_root, model_name, torch_time_series, time_data, cache_path, pred_dim=0,
                      sample_window_step=10, windows=None, pred_len=None, sampling_t=None, infer_params=None,
                      force_recompute=False, max_windows=None, device=None):
    """:1039-1114."""
    cfg = _slbp_fig6_dataset_config(model_root, model_name, windows=windows, pred_len=pred_len, sampling_t=sampling_t)
    cache_path = _resolve_project_path(cache_path)
    sampled_time = torch_data_preprocessing_like_slbp(time_data, sampling_t=cfg["sampling_t"], return_numpy=True)
    base = {"cache_path": str(cache_path), "windows": cfg["windows"], "pred_len": cfg["pred_len"],
            "sampling_t": cfg["sampling_t"]}
    if _cache_ready(cache_path) and not force_recompute:
        data_list = _load_tensor_list(cache_path)
        step = infer_sample_window_step_from_cache(len(sampled_time), cfg["windows"], len(data_list), sample_window_step)
        mpv, source = summarize_slbp_mpv_cache_for_fig5(data_list, pred_dim=pred_dim)
        return dict(base, time_points=sampled_time[cfg["windows"] - 1:: step][: len(mpv)], mpv=mpv,
                    pred_future_list=data_list, sample_window_step=step, uncertainty_source=source)
    device = _default_device(device)
    model, _, _ = load_sensitivity_model(model_root, model_name, device=device, infer_params=infer_params)
    input_datas, _, time_points = build_slbp_sensitivity_windows(
        torch_time_series, time_data, cfg["windows"], cfg["pred_len"], cfg["sampling_t"], sample_window_step)
    pred_future_list = run_slbp_sensitivity_cache(model, input_datas, cache_path, device, force_recompute, max_windows)
    mpv, source = summarize_slbp_mpv_cache_for_fig5(pred_future_list, pred_dim=pred_dim)
    return dict(base, time_points=time_points[: len(mpv)], mpv=mpv, pred_future_list=pred_future_list,
                sample_window_step=sample_window_step, uncertainty_source=source)


def slbp_direct_model_cache_analysis(model_save_file, torch_time_series, time_data, cache_path, pred_dim=0,
                                     sample_window_step=10, cache_kind="auto", infer_params=None,
                                     force_recompute=False, max_windows=None, device=None,
                                     compute_prediction_error=False):
    """:1117-1229."""
    method_config = read_model_config(model_save_file)
    ds = method_config.get("dataset", {})
    net = method_config.get("net", {})
    windows = int(ds.get("windows", net.get("windows", 200)))
    pred_len = int(ds.get("pred_len", net.get("pred_len", 200)))
    sampling_t = int(ds.get("sampling_t", 100))
    cache_path = _resolve_project_path(cache_path)
    sampled_time = torch_data_preprocessing_like_slbp(time_data, sampling_t=sampling_t, return_numpy=True)
    model = None
    if _cache_ready(cache_path) and not force_recompute:
        data_list = _load_tensor_list(cache_path)
    else:
        if cache_kind not in {"gx", "sampling"}:
            raise ValueError("cache_kind must be 'gx' or 'sampling' when cache is missing.")
        device = _default_device(device)
        model, _ = load_model_from_dir(model_save_file, device=device, infer_params=infer_params,
                                       method_config=method_config)
        input_datas, _, _ = build_slbp_sensitivity_windows(torch_time_series, time_data, windows, pred_len, sampling_t,
                                                           sample_window_step)
        if cache_kind == "gx":
            data_list = run_slbp_gx_cache_for_fig6(model, input_datas, cache_path, device, pred_dim, force_recompute,
                                                   max_windows)
        else:
            data_list = run_slbp_sensitivity_cache(model, input_datas, cache_path, device, force_recompute, max_windows)
    mpv, source = summarize_slbp_mpv_cache_for_fig5(data_list, pred_dim=pred_dim)
    step = infer_sample_window_step_from_cache(len(sampled_time), windows, len(data_list), sample_window_step)
    result = {"time_points": sampled_time[windows - 1:: step][: len(mpv)], "mpv": mpv, "pred_future_list": data_list,
              "cache_path": str(cache_path), "windows": windows, "pred_len": pred_len, "sampling_t": sampling_t,
              "sample_window_step": step, "uncertainty_source": source}
    if compute_prediction_error:
        if source != "sampling":
            raise ValueError("prediction_error requires a sampling cache, got '{}'.".format(source))
        if model is None:
            device = _default_device(device)
            model, _ = load_model_from_dir(model_save_file, device=device, infer_params=infer_params,
                                           method_config=method_config)
        _, pred_datas, _ = build_slbp_sensitivity_windows(torch_time_series, time_data, windows, pred_len, sampling_t, step)
        _, err = summarize_slbp_sensitivity(data_list, pred_datas[: len(data_list)], model=model, device=device,
                                            pred_dim=pred_dim)
        result["prediction_error"] = err
    return result


def slbp_raw_window_variance(torch_time_series, time_data, windows=200, sampling_t=100, sample_window_step=10,
                             pred_dim=0):
    """:1232-1250: classic rolling-variance indicator on the raw series (no model)."""
    series = torch_data_preprocessing_like_slbp(torch_time_series, sampling_t=sampling_t)
    sampled_time = torch_data_preprocessing_like_slbp(time_data, sampling_t=sampling_t, return_numpy=True)
    if series.ndim != 2:
        raise ValueError("SLBP raw series must have shape [T, F].")
    if pred_dim >= series.shape[1]:
        raise IndexError("pred_dim {} out of bounds for F={}.".format(pred_dim, series.shape[1]))
    w = series[:, pred_dim].unfold(0, windows, sample_window_step)
    variances = w.var(dim=1, unbiased=False).cpu().detach().numpy()
    return {"time_points": sampled_time[windows - 1:: sample_window_step][: len(variances)], "variance": variances,
            "windows": windows, "sampling_t": sampling_t, "sample_window_step": sample_window_step}


# ------------------------------------------------------------------------------------------------
# uncertainty_ews: network data, the north-star sweep (:1253-1541)
# ------------------------------------------------------------------------------------------------
_METHOD_ALIASES = {"variance": "sampling", "sampling_variance": "sampling", "pred_future": "sampling",
                   "pred": "sampling", "g": "gx", "preg": "gx", "nsdiff_g": "gx"}


def uncertainty_ews(model_save_file=None, data_file=None, torch_time_series=None, time_data=None, dynamic_type=None,
                    task_model=None, graph_file=None, cache_path=None, sample_window_step=None, sampling_t=None,
                    infer_params=None, pred_dim=0, force_recompute=False, save_nsdiff_g=True, nsdiff_g_path=None,
                    uncertainty_method="sampling", max_windows=None, device=None, load_model_when_cached=False):
    dynamic_type = _dynamic_name(dynamic_type)
    uncertainty_method = str(uncertainty_method).lower()
    uncertainty_method = _METHOD_ALIASES.get(uncertainty_method, uncertainty_method)
    if uncertainty_method not in {"sampling", "gx", "both"}:
        raise ValueError("uncertainty_method must be one of: sampling, gx, both.")

    if data_file is not None:
        data_file = _resolve_project_path(data_file)
        loaded = load_dynamic_data(data_file, dynamic_type=dynamic_type)
        torch_time_series, time_data = loaded["torch_time_series"], loaded["time_data"]
        dynamic_type = _dynamic_name(dynamic_type) or loaded["dynamic_type"]
    elif torch_time_series is not None:
        torch_time_series = normalize_time_series(torch_time_series, dynamic_type=dynamic_type)
    else:
        raise ValueError("Provide data_file or torch_time_series.")
    if time_data is None:
        raise ValueError("time_data is required when data_file is not provided.")

    method_config, model, loaded_net_param = None, None, None
    if model_save_file is not None:
        model_save_file = _resolve_project_path(model_save_file)
        method_config = read_model_config(model_save_file)
    if task_model is None and method_config is not None:
        task_model = method_config.get("net", {}).get("task_model")
    dataset_config = method_config.get("dataset", {}) if method_config else {}
    windows, pred_len = dataset_config.get("windows"), dataset_config.get("pred_len")
    if windows is None or pred_len is None:
        raise ValueError("model_trained.yaml must provide dataset.windows and dataset.pred_len.")

    cache_path = resolve_cache_path(cache_path, model_save_file, data_file, dynamic_type)
    need_sampling = uncertainty_method in {"sampling", "both"}
    need_gx = uncertainty_method in {"gx", "both"} or (save_nsdiff_g and uncertainty_method == "sampling")
    nsdiff_path = None
    if need_gx:
        nsdiff_path = resolve_cache_path(nsdiff_g_path if nsdiff_g_path is not None else cache_path.parent,
                                         model_save_file, data_file, dynamic_type, suffix="_gx")

    cached_pred, cached_g = None, None
    if need_sampling and _cache_ready(cache_path) and not force_recompute:
        cached_pred = _load_tensor_list(cache_path)
        if task_model == "DiffSTG":
            cached_pred = normalize_diffstg_pred_future_list(cached_pred)
    if need_gx and nsdiff_path is not None and _cache_ready(nsdiff_path) and not force_recompute:
        cached_g = _load_tensor_list(nsdiff_path)

    if sampling_t is None:
        sampling_t = dataset_config.get("sampling_t", DEFAULT_SAMPLING_T.get(dynamic_type, 0.1))
    sampled_series, sampled_time = sample_time_series(torch_time_series, time_data, sampling_t=sampling_t)
    if sample_window_step is None:
        fallback = default_sample_window_step(dynamic_type, task_model=task_model, dataset_config=dataset_config)
        n_cached = len(cached_pred) if cached_pred is not None else (len(cached_g) if cached_g is not None else None)
        sample_window_step = fallback if n_cached is None else infer_sample_window_step_from_cache(
            sampled_series.shape[1], windows, n_cached, fallback)
    timeseries_datas, time_points = build_sliding_windows(sampled_series, sampled_time, windows, sample_window_step)

    def load():
        return load_model_from_dir(model_save_file, device=_default_device(device), infer_params=infer_params,
                                   method_config=method_config)

    pred_future_list, pred_mean_list, ews_list = None, [], []
    if need_sampling:
        if task_model == "DiffSTG":
            if dynamic_type not in NETWORK_DYNAMICS:
                raise ValueError("DiffSTG only supports network dynamics: SIS, neuronal, biomass.")
            if graph_file is None:
                raise ValueError("graph_file is required for DiffSTG.")
        if cached_pred is not None:
            pred_future_list = cached_pred
            if model_save_file is not None and load_model_when_cached:
                model, loaded_net_param = load()
        else:
            if model_save_file is None:
                raise ValueError("model_save_file is required when cache_path does not exist or force_recompute=True.")
            if task_model == "DiffSTG":
                if infer_params is None:
                    infer_params = {"parallel_sampling": 10, "sequential_sampling": 1, "n_z_samples": 10,
                                    "diffusion_steps": 20}
                model, loaded_net_param = load()
                pred_future_list = run_diffstg_evaluation_cache(model, timeseries_datas, pred_len,
                                                                load_diffstg_graph(graph_file), cache_path,
                                                                _default_device(device), force_recompute, max_windows)
            else:
                model, loaded_net_param = load()
                pred_future_list = run_evaluation_cache(model, timeseries_datas, pred_len, cache_path,
                                                        _default_device(device), force_recompute, max_windows)
        pred_mean_list, ews_list = summarize_pred_future_list(pred_future_list, model=model)

    valid_len = len(ews_list)
    result = {
        "pred_future_list": pred_future_list, "pred_mean": pred_mean_list, "ews": ews_list,
        "time_points": time_points[:valid_len], "cache_path": str(cache_path),
        "figure_path": str(resolve_figure_path(cache_path)), "torch_time_series": torch_time_series,
        "time_data": torch.as_tensor(time_data).cpu().detach().numpy(), "dynamic_type": dynamic_type,
        "sampling_t": sampling_t, "sample_window_step": sample_window_step, "windows": windows, "pred_len": pred_len,
        "task_model": task_model, "uncertainty_method": uncertainty_method,
        "uncertainty_source": "sampling" if need_sampling else None,
        "graph_file": str(_resolve_project_path(graph_file)) if graph_file is not None else None,
        "model_save_file": str(model_save_file) if model_save_file is not None else None,
        "loaded_net_param": loaded_net_param,
    }

    has_g_model = model is not None and getattr(model, "cond_pred_model_g", None) is not None
    g_list = None
    if need_gx and ("NsDiff" in str(task_model) or has_g_model):
        if cached_g is not None:
            g_list = cached_g
        else:
            if model is None and model_save_file is not None:
                model, loaded_net_param = load()
                result["loaded_net_param"] = loaded_net_param
            if model is not None and getattr(model, "cond_pred_model_g", None) is not None:
                g_list = run_nsdiff_g_cache(model, timeseries_datas, nsdiff_path, _default_device(device),
                                            pred_dim=pred_dim, force_recompute=force_recompute, max_windows=max_windows)
        if g_list is not None:
            g_pred_mean, g_ews = summarize_nsdiff_g_list(g_list, pred_dim=pred_dim)
            result["nsdiff_g"] = {"pred_future_list": g_list, "pred_mean": g_pred_mean, "ews": g_ews,
                                  "time_points": time_points[: len(g_ews)], "cache_path": str(nsdiff_path)}
            if uncertainty_method == "gx":
                result.update(pred_future_list=None, pred_mean=g_pred_mean, ews=g_ews,
                              time_points=time_points[: len(g_ews)], cache_path=str(nsdiff_path),
                              figure_path=str(resolve_figure_path(nsdiff_path)), uncertainty_source="gx")
    if uncertainty_method == "gx" and g_list is None:
        raise ValueError("uncertainty_method='gx' requires a task_model containing 'NsDiff' "
                         "and a loaded model with cond_pred_model_g, or an existing _gx cache.")
    return result


# ------------------------------------------------------------------------------------------------
# multi-GPU: one process per GPU, windows partitioned, ONE gather of the per-window statistics
# ------------------------------------------------------------------------------------------------
def gather_window_stats(local, n_windows, group=None):
    """All-gather per-window rows.  ``local`` [n_local, C] holds this rank's contiguous block
    (partition_windows); returns [n_windows, C] on every rank.  The only collective of a sweep: windows
    are independent, so trajectories stay rank-local.  NCCL needs device tensors, gloo (CPU
    tests) takes host tensors; the tensor is used where it lives."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = -(-n_windows // world)
    buf = local.new_zeros((per,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat(out)[:n_windows]


def distributed_sweep(model, stacked_windows, device=None, group=None, graph_data=None):
    """Sweep sharded over the ranks of ``group``: rank r samples its contiguous block of windows (Philox
    keys use the global window index, so the union equals the single-GPU sweep), reduces it on its GPU,
    and the ranks exchange only [W, 2+F] floats.  Returns (local_cache [w1-w0,B,K,O,F], (w0,w1), stats) with
    stats = dict(mpv [W], pred_mean [W], mpv_f [W,F]) identical on every rank.  ``graph_data``: DiffSTG / NsDiff_spatial (windows and
    sample replicas shard; nodes are coupled by the graph conv and do not)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    W = stacked_windows.shape[0]
    w0, w1 = partition_windows(W, world, rank)
    device = device or _model_device(model)
    F = model.F if hasattr(model, "parallel_sampling") else model.dataset_nf
    start = getattr(model, "_windows_drawn", 0)
    if w1 > w0:
        cache = sample_sweep(model, stacked_windows[w0:w1], device=device, window_offset=w0, graph_data=graph_data)
        st = cache.upd_stats.get("raw", cache.upd_stats["scaled"])
        local = torch.cat([st["mpv"].view(-1, 1), st["pred_mean"].view(-1, 1), st["mpv_f"].view(-1, F)], dim=1)
    else:
        cache = torch.empty((0,), dtype=torch.float32)
        local = torch.zeros((0, 2 + F), dtype=torch.float32)
    if hasattr(model, "_windows_drawn"):
        model._windows_drawn = start + W          # every rank advances by the whole sweep
    if dist.get_backend(group) == "nccl":
        local = local.to(device)
    full = gather_window_stats(local, W, group=group).cpu()
    return cache, (w0, w1), {"mpv": full[:, 0], "pred_mean": full[:, 1], "mpv_f": full[:, 2:]}


def main():
    """:1591-1625 without the figure (plotting is outside this package): the reference's default run."""
    run_config = {
        "model_save_file": "ews_results/model_compare/NsDiff/SIS",
        "data_file": "dataset/spdata_sde_SIS/barabasi_albert_30_0/SIS_dynamic_eta0.0001d0.5_increase.pt",
        "dynamic_type": "SIS", "task_model": None, "graph_file": "dataset/test_graph/barabasi_albert_30_0.graphml",
        "cache_path": None, "sample_window_step": None, "sampling_t": None, "pred_dim": 0, "force_recompute": False,
        "uncertainty_method": "gx", "device": None,
        "infer_params": {"parallel_sampling": 50, "sequential_sampling": 1, "n_z_samples": 100, "diffusion_steps": 20},
    }
    result = uncertainty_ews(**run_config)
    flush_cache_writes()
    print("cache_path:", result["cache_path"])
    print("figure_path:", result["figure_path"])
    print("num_windows:", len(result["ews"]))


if __name__ == "__main__":
    main()
