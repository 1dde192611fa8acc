"""CPU oracle (test infrastructure, never shipped): the dispersion reductions (MPV, gx-EWS).

float64 numpy restatement of evaluation_and_analysis/diffusion_model_uncertainy.py:267-320,
529-550, 686-728.  The reference reduces in fp32 on the CPU; float64 here is the exact value
both the reference and the CUDA Welford kernel approximate (tolerances live in the tests).
"""
import numpy as np


def inverse_transform(pred_future, mean, std):
    """:267-283: features sit on axis -2 of [..., F, K]."""
    x = np.asarray(pred_future, dtype=np.float64)
    shape = [1] * x.ndim
    shape[-2] = -1
    return x * np.asarray(std, np.float64).reshape(shape) + np.asarray(mean, np.float64).reshape(shape)


def network_mpv(pred_future, mean=None, std=None):
    """:286-303: (pred_mean, MPV) for one window [Node,O,F,K]; biased variance over K, mean over the rest."""
    x = np.asarray(pred_future, dtype=np.float64)
    if x.ndim == 3:
        x = x[None]
    if x.ndim != 4:
        raise ValueError("pred_future must have shape [Node, pred_len, F, n_z_samples]")
    if mean is not None:
        x = inverse_transform(x, mean, std)
    return x.mean(), x.var(axis=-1).mean()


def slbp_mpv(pred_future, pred_dim=0):
    """:529-541 / :701-713: var over K, mean over pred_len, pick feature pred_dim; input [O,F,K]."""
    x = np.asarray(pred_future, dtype=np.float64)
    if x.ndim != 3:
        raise ValueError("SLBP cache elements must have shape [pred_len, F, n_z_samples].")
    if pred_dim >= x.shape[1]:
        raise IndexError("pred_dim out of bounds")
    return x.var(axis=-1).mean(axis=0)[pred_dim]


def slbp_prediction_error(pred_future, target_scaled, pred_dim=0):
    """:542-549: | mean_K - target | averaged over pred_len, feature pred_dim."""
    x = np.asarray(pred_future, dtype=np.float64)
    return np.abs(x.mean(axis=-1) - np.asarray(target_scaled, np.float64)).mean(axis=0)[pred_dim]


def intrinsic_dimension(pred_future, energy=0.8):
    """:686-698: #eigenvalues of the (O*F)^2 sample covariance of the K trajectories reaching 80 %."""
    x = np.asarray(pred_future, dtype=np.float64)
    traj = np.transpose(x, (2, 0, 1)).reshape(x.shape[-1], -1)
    if traj.shape[0] < 2:
        return float("nan")
    c = traj - traj.mean(axis=0, keepdims=True)
    cov = c.T @ c / max(traj.shape[0] - 1, 1)
    ev = np.clip(np.sort(np.linalg.eigvalsh(cov))[::-1], 0, None)
    tot = ev.sum()
    if tot <= 0:
        return float("nan")
    return int(np.where(np.cumsum(ev / tot) >= energy)[0][0] + 1)


def gx_ews(gx, pred_dim=0):
    """:306-320: (gx.mean(), gx.mean(dim=1)[:, pred_dim].mean()) for [Node,O,F] (or [O,F])."""
    g = np.asarray(gx, dtype=np.float64)
    if g.ndim == 2:
        g = g[None]
    if g.ndim != 3:
        raise ValueError("NsDiff-g cache elements must have shape [Node, pred_len, F].")
    if pred_dim >= g.shape[-1]:
        raise IndexError("pred_dim out of bounds")
    return g.mean(), g.mean(axis=1)[:, pred_dim].mean()
