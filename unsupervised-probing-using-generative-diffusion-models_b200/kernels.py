"""Python-side launchers over the C ABI (include/upd_b200.h).  Tensors in, tensors out; every call
is asynchronous on torch's current CUDA stream.  PyTorch is used for device memory and streams only."""
import ctypes

import torch

from . import _lib
from ._lib import IMPL_SIMT, IMPL_TCGEN05, IMPL_TCGEN05_X2, IMPL_TCGEN05_WS, KIND_NSDIFF, KIND_TMDM  # noqa: F401

DEN = "model.diffussion_model."   # the reference's spelling, part of the checkpoint key names


def pack_denoiser(state_dict, kind, F, T, sched_rows, device):
    """state dict (reference key names) + [n_sched,T] schedule table -> packed device blob (uint8)."""
    L = _lib.lib()
    n = L.upd_denoiser_pack_bytes(kind, F, T)
    if n == 0:
        raise ValueError("unsupported denoiser dims kind={} F={} T={} (F in 1..4, T in 2..64)".format(kind, F, T))
    keep = []

    def host(key):
        t = state_dict[key].detach().to("cpu", torch.float32).contiguous()
        keep.append(t)
        return ctypes.c_void_p(t.data_ptr())

    w = _lib.UpdDenoiserWeights()
    w.kind, w.F, w.T = kind, F, T
    for i in (1, 2, 3):
        setattr(w, "lin%d_w" % i, host(DEN + "lin%d.lin.weight" % i))
        setattr(w, "lin%d_b" % i, host(DEN + "lin%d.lin.bias" % i))
        setattr(w, "embed%d" % i, host(DEN + "lin%d.embed.weight" % i))
    w.lin4_w, w.lin4_b = host(DEN + "lin4.weight"), host(DEN + "lin4.bias")
    if kind == KIND_NSDIFF:
        w.sigma_w, w.sigma_b = host(DEN + "sigma_lin.weight"), host(DEN + "sigma_lin.bias")
    in_dim = (3 if kind == KIND_NSDIFF else 2) * F
    te = T if kind == KIND_NSDIFF else T + 1
    if tuple(state_dict[DEN + "lin1.lin.weight"].shape) != (128, in_dim):
        raise ValueError("lin1 weight shape {} does not match dataset_nf={}".format(
            tuple(state_dict[DEN + "lin1.lin.weight"].shape), F))
    if tuple(state_dict[DEN + "lin1.embed.weight"].shape) != (te, 128):
        raise ValueError("step-embedding table has {} rows, expected {}".format(
            state_dict[DEN + "lin1.embed.weight"].shape[0], te))
    sched = sched_rows.detach().to("cpu", torch.float32).contiguous()
    keep.append(sched)
    w.sched = ctypes.c_void_p(sched.data_ptr())
    blob = torch.zeros(n, dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else torch.zeros(n, dtype=torch.uint8)
    _lib.check(L.upd_denoiser_pack(ctypes.byref(w), ctypes.c_void_p(blob.data_ptr()), n), "upd_denoiser_pack")
    if device is None:
        return blob
    return blob.to(device)


def nsdiff_sample(packed, y0_hat, gx, n_win, B, K, S, O, F, T, seed=0, window_base=0, noise=None, impl=IMPL_TCGEN05,
                  out=None):
    """-> out [n_win*B, K, O, F] fp32 on gx.device (see upd_nsdiff_sample)."""
    dev = gx.device
    _lib.require_cuda(dev)
    if out is None:
        out = torch.empty((n_win * B, K, O, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().upd_nsdiff_sample(_lib.ptr(packed), _lib.ptr(y0_hat), _lib.ptr(gx), n_win, B, K, S, O, F, T,
                                          seed & (2 ** 64 - 1), window_base, _lib.ptr(noise), _lib.ptr(out), impl,
                                          _lib.stream_ptr(dev))
    _lib.check(rc, "upd_nsdiff_sample")
    return out


def tmdm_sample(packed, y0_hat, n_win, B, K, S, Lr, F, T, seed=0, window_base=0, noise=None, impl=IMPL_TCGEN05,
                out=None):
    dev = y0_hat.device
    _lib.require_cuda(dev)
    if out is None:
        out = torch.empty((n_win * B, K, Lr, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().upd_tmdm_sample(_lib.ptr(packed), _lib.ptr(y0_hat), n_win, B, K, S, Lr, F, T,
                                        seed & (2 ** 64 - 1), window_base, _lib.ptr(noise), _lib.ptr(out), impl,
                                        _lib.stream_ptr(dev))
    _lib.check(rc, "upd_tmdm_sample")
    return out


def mpv_reduce(traj, n_win, B, scale=None, want_var=False, want_mean=False):
    """traj [n_win*B, K, O, F] -> dict(mpv [n_win], pred_mean [n_win], mpv_f [n_win,F], var?, mean?)."""
    dev = traj.device
    _lib.require_cuda(dev)
    R0, K, O, F = traj.shape
    if R0 != n_win * B:
        raise ValueError("traj has {} rows, expected n_win*B = {}".format(R0, n_win * B))
    L = _lib.lib()
    var = torch.empty((R0, O, F), dtype=torch.float32, device=dev)
    mean = torch.empty((R0, O, F), dtype=torch.float32, device=dev)
    mpv = torch.empty(n_win, dtype=torch.float32, device=dev)
    pmean = torch.empty(n_win, dtype=torch.float32, device=dev)
    mpv_f = torch.empty((n_win, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.upd_mpv_reduce(_lib.ptr(traj), _lib.ptr(scale), n_win, B, K, O, F, _lib.ptr(var), _lib.ptr(mean),
                              _lib.ptr(mpv), _lib.ptr(pmean), _lib.ptr(mpv_f), None, _lib.stream_ptr(dev))
    _lib.check(rc, "upd_mpv_reduce")
    res = {"mpv": mpv, "pred_mean": pmean, "mpv_f": mpv_f}
    if want_var:
        res["var"] = var
    if want_mean:
        res["mean"] = mean
    return res


def sigma_estimation(weights, x, R, O, add_eps=0.0):
    """weights: 10 device fp32 tensors (w0,b0,ln1_w,ln1_b,w3,b3,ln2_w,ln2_b,w6,b6); x [rows,L,F] -> gx [rows,O,F]."""
    dev = x.device
    _lib.require_cuda(dev)
    rows, Lw, F = x.shape
    H = weights[0].shape[0]
    if weights[0].shape[1] != Lw - R:
        raise ValueError("g(x) first layer expects {} inputs, window {} - rolling_length {} = {}".format(
            weights[0].shape[1], Lw, R, Lw - R))
    w = _lib.UpdSigmaWeights(*[ctypes.c_void_p(t.data_ptr()) for t in weights])
    gx = torch.empty((rows, O, F), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().upd_sigma_estimation(ctypes.byref(w), _lib.ptr(x), rows, Lw, R, F, H, O, float(add_eps),
                                             _lib.ptr(gx), _lib.stream_ptr(dev))
    _lib.check(rc, "upd_sigma_estimation")
    return gx
