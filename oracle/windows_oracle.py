"""CPU oracle (test infrastructure, never shipped): rolling-window / step bookkeeping.

Integer work, must be bit-exact.  Written with explicit Python/numpy index arithmetic
(no unfold), following evaluation_and_analysis/diffusion_model_uncertainy.py:119-182,467-483.
"""
import math

import numpy as np


def sampling_interval_from_t(sampling_t):
    """:119-125.  Keeps the reference's float quirk: the stride is int(sampling_t / 0.1)."""
    if sampling_t is None or sampling_t <= 0.1:
        return 1
    return max(1, int(sampling_t / 0.1))


def sample_indices(length, sampling_t):
    """:128-134: indices kept by x[:, ::interval, :]."""
    return list(range(0, length, sampling_interval_from_t(sampling_t)))


def sliding_window_count(sampled_length, windows, step):
    """:157-160."""
    if sampled_length < windows:
        return 0
    return (sampled_length - windows) // step + 1


def window_starts(sampled_length, windows, step):
    """:137-147: window w covers sampled indices [w*step, w*step + windows)."""
    return [w * step for w in range(sliding_window_count(sampled_length, windows, step))]


def time_point_indices(sampled_length, windows, step):
    """:146: time_points = t[windows-1::step] (may be one longer than the window count never; equal)."""
    return list(range(windows - 1, sampled_length, step))


def build_sliding_windows(series, time_data, windows, step):
    """:137-147 on a numpy array [Node, T, F] -> (list of [Node, windows, F], time_points)."""
    series = np.asarray(series)
    if series.ndim != 3:
        raise ValueError("torch_time_series must have shape [Node_num, T_obs_num, F].")
    if series.shape[1] < windows:
        raise ValueError("T_obs_num ({}) is shorter than windows ({}).".format(series.shape[1], windows))
    out = [series[:, s:s + windows, :] for s in window_starts(series.shape[1], windows, step)]
    tp = np.asarray(time_data)[time_point_indices(len(time_data), windows, step)]
    return out, tp


def build_slbp_windows(series, time_data, windows, pred_len, sampling_t, step):
    """:467-483 on numpy [T_raw, F]: inputs [windows,F], targets [pred_len,F] from series[windows:]."""
    series = np.asarray(series)
    time_data = np.asarray(time_data)
    iv = sampling_interval_from_t(sampling_t)
    s = series[::iv]
    t = time_data[::iv]
    inputs = [s[a:a + windows] for a in window_starts(s.shape[0], windows, step)]
    tail = s[windows:]
    targets = []
    if tail.shape[0] >= pred_len:
        targets = [tail[a:a + pred_len] for a in window_starts(tail.shape[0], pred_len, step)]
    return inputs, targets, t[windows - 1::step]


def infer_sample_window_step_from_cache(sampled_length, windows, cache_len, fallback_step):
    """:163-182: recover the step from a cache length; ties -> closest to fallback, then larger."""
    if cache_len <= 0 or sampled_length < windows:
        return fallback_step
    if sliding_window_count(sampled_length, windows, fallback_step) == cache_len:
        return fallback_step
    if cache_len == 1:
        return fallback_step
    max_offset = sampled_length - windows
    low = int(math.floor(max_offset / cache_len)) + 1
    high = int(math.floor(max_offset / (cache_len - 1)))
    best = None
    for step in range(max(1, low), max(1, high) + 1):
        if sliding_window_count(sampled_length, windows, step) != cache_len:
            continue
        key = (abs(step - fallback_step), -step)
        if best is None or key < best[0]:
            best = (key, step)
    return fallback_step if best is None else best[1]
