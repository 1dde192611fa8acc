"""ctypes binding of include/upd_b200.h.  There is no CPU fallback: if the library is missing or a
call fails this module raises."""
import ctypes
import os

import torch

from . import _build

_LIB = None

SYMBOLS = (
    "upd_error_string", "upd_last_cuda_error", "upd_abi_version", "upd_denoiser_pack_bytes", "upd_denoiser_pack",
    "upd_nsdiff_sample", "upd_tmdm_sample", "upd_mpv_scratch_bytes", "upd_mpv_reduce", "upd_gram_centered", "upd_prediction_error", "upd_sigma_estimation",
    "upd_dts_ddim_step", "upd_dts_adagrad_step", "upd_dts_infill", "upd_gauss_fill",
    "upd_dts_fourier_topk", "upd_dts_fourier_topk_bwd", "upd_dts_attention", "upd_dts_attention_bwd", "upd_dts_layernorm", "upd_dts_layernorm_bwd", "upd_stg_posterior", "upd_nsx_step", "upd_stg_gated_aggregate", "upd_stg_tcn_ln", "upd_stg_tcn_ln_cat", "upd_stg_conv1d", "upd_fx_split", "upd_gemm3", "upd_fx_add_ln_split", "upd_fx_attention", "upd_fx_attention_hs16", "upd_fx_embed_split",
)

ABI_VERSION = 8     # = UPD_ABI_VERSION of the csrc/ this binding was written against (argument lists below)
KIND_NSDIFF, KIND_TMDM = 0, 1
IMPL_TCGEN05, IMPL_SIMT, IMPL_TCGEN05_X2, IMPL_TCGEN05_WS = 0, 1, 2, 4
_fp = ctypes.POINTER(ctypes.c_float)


class UpdDenoiserWeights(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("F", ctypes.c_int), ("T", ctypes.c_int)] + [
        (n, ctypes.c_void_p) for n in ("lin1_w", "lin1_b", "embed1", "lin2_w", "lin2_b", "embed2", "lin3_w", "lin3_b",
                                       "embed3", "lin4_w", "lin4_b", "sigma_w", "sigma_b", "sched")]


class UpdSigmaWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w0", "b0", "ln1_w", "ln1_b", "w3", "b3", "ln2_w", "ln2_b", "w6", "b6")]


def lib():
    """Load (once) the in-tree shared library; raise loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("UPD_LIB_PATH", _build.LIB_PATH)     # override: debug / experimental builds of csrc/
    if not os.path.exists(path):
        raise RuntimeError(
            "CUDA library {} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU fallback.".format(path))
    L = ctypes.CDLL(path)
    L.upd_error_string.restype = ctypes.c_char_p
    L.upd_error_string.argtypes = [ctypes.c_int]
    L.upd_last_cuda_error.restype = ctypes.c_int
    L.upd_abi_version.restype = ctypes.c_int
    if L.upd_abi_version() != ABI_VERSION:
        # a stale build would be called with mismatched argument lists (silent memory corruption): refuse it
        raise RuntimeError("{} has ABI version {}, this package needs {}: rebuild it (python -c 'import "
                           "__graft_entry__ as g; g.build()')".format(path, L.upd_abi_version(), ABI_VERSION))
    L.upd_denoiser_pack_bytes.restype = ctypes.c_size_t
    L.upd_denoiser_pack_bytes.argtypes = [ctypes.c_int] * 3
    L.upd_denoiser_pack.restype = ctypes.c_int
    L.upd_denoiser_pack.argtypes = [ctypes.POINTER(UpdDenoiserWeights), ctypes.c_void_p, ctypes.c_size_t]
    vp, i, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    L.upd_nsdiff_sample.restype = ctypes.c_int
    L.upd_nsdiff_sample.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, u64, u64, vp, vp, i, vp]
    L.upd_tmdm_sample.restype = ctypes.c_int
    L.upd_tmdm_sample.argtypes = [vp, vp, i, i, i, i, i, i, i, u64, u64, vp, vp, i, vp]
    L.upd_mpv_scratch_bytes.restype = ctypes.c_size_t
    L.upd_mpv_scratch_bytes.argtypes = [i, i, i, i]
    L.upd_mpv_reduce.restype = ctypes.c_int
    L.upd_mpv_reduce.argtypes = [vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.upd_gram_centered.restype = ctypes.c_int
    L.upd_gram_centered.argtypes = [vp, i, i, i, vp, vp]
    L.upd_prediction_error.restype = ctypes.c_int
    L.upd_prediction_error.argtypes = [vp, vp, i, i, i, vp, vp]
    L.upd_sigma_estimation.restype = ctypes.c_int
    L.upd_sigma_estimation.argtypes = [ctypes.POINTER(UpdSigmaWeights), vp, i, i, i, i, i, i, ctypes.c_float, vp, vp]
    ll, f32, u32 = ctypes.c_longlong, ctypes.c_float, ctypes.c_uint32
    L.upd_dts_ddim_step.restype = ctypes.c_int
    L.upd_dts_ddim_step.argtypes = [vp, vp, ll, f32, f32, f32, f32, f32, vp, i, vp, vp, vp, vp]
    L.upd_dts_adagrad_step.restype = ctypes.c_int
    L.upd_dts_adagrad_step.argtypes = [vp, vp, ll, f32, vp]
    L.upd_dts_infill.restype = ctypes.c_int
    L.upd_dts_infill.argtypes = [vp, vp, vp, vp, ll, i, i, i, f32, f32, vp]
    L.upd_gauss_fill.restype = ctypes.c_int
    L.upd_gauss_fill.argtypes = [vp, ll, ll, u64, u64, u32, vp]
    L.upd_dts_fourier_topk.restype = ctypes.c_int
    L.upd_dts_fourier_topk.argtypes = [vp, ll, ll, i, i, i, i, i, i, vp, vp, vp]
    L.upd_dts_fourier_topk_bwd.restype = ctypes.c_int
    L.upd_dts_fourier_topk_bwd.argtypes = [vp, vp, ll, ll, i, i, i, i, i, vp, vp]
    L.upd_dts_layernorm.restype = ctypes.c_int
    L.upd_dts_layernorm.argtypes = [vp, vp, vp, ll, i, vp, vp, vp, vp]
    L.upd_dts_layernorm_bwd.restype = ctypes.c_int
    L.upd_dts_layernorm_bwd.argtypes = [vp, vp, vp, vp, ll, i, vp, vp]
    L.upd_dts_attention.restype = ctypes.c_int
    L.upd_dts_attention.argtypes = [vp, ll, vp, vp, ll, i, i, i, i, i, f32, vp, vp, vp, vp]
    L.upd_dts_attention_bwd.restype = ctypes.c_int
    L.upd_dts_attention_bwd.argtypes = [vp, ll, vp, vp, ll, i, i, i, i, i, f32, vp, vp, vp, vp, ll, vp, vp, ll, vp]
    L.upd_nsx_step.restype = ctypes.c_int
    L.upd_nsx_step.argtypes = [vp] * 10 + [ctypes.c_int, ctypes.c_int, ll] + [ctypes.c_int] * 3 + [vp] * 4
    L.upd_stg_posterior.restype = ctypes.c_int
    L.upd_stg_posterior.argtypes = [vp, vp, vp, ll, f32, f32, f32, vp, vp]
    L.upd_stg_gated_aggregate.restype = ctypes.c_int
    L.upd_stg_gated_aggregate.argtypes = [vp, vp, vp, vp, ll, i, i, i, vp, vp]
    L.upd_gemm3.restype = ctypes.c_int
    L.upd_gemm3.argtypes = [vp, vp, ll, i, i, i, vp, vp, vp]
    L.upd_fx_split.restype = ctypes.c_int
    L.upd_fx_split.argtypes = [vp, ll, i, i, i, i, vp, vp]
    L.upd_fx_add_ln_split.restype = ctypes.c_int
    L.upd_fx_add_ln_split.argtypes = [vp, vp, vp, vp, vp, vp, ll, i, vp, vp, vp]
    L.upd_fx_attention_hs16.restype = ctypes.c_int
    L.upd_fx_attention_hs16.argtypes = [vp, ll, vp, vp, ll, vp, vp, i, i, i, i, i, i, f32, vp, vp]
    L.upd_fx_embed_split.restype = ctypes.c_int
    L.upd_fx_embed_split.argtypes = [vp, vp, vp, ll, i, i, i, vp, vp, vp]
    L.upd_fx_attention.restype = ctypes.c_int
    L.upd_fx_attention.argtypes = [vp, ll, vp, vp, ll, vp, vp, i, i, i, i, i, i, i, f32, vp, vp]
    L.upd_stg_tcn_ln.restype = ctypes.c_int
    L.upd_stg_tcn_ln.argtypes = [vp, vp, vp, vp, vp, vp, vp, ll, i, i, i, vp, vp, vp, vp, vp]
    L.upd_stg_tcn_ln_cat.restype = ctypes.c_int
    L.upd_stg_tcn_ln_cat.argtypes = [vp, i, vp, i, vp, vp, vp, vp, vp, vp, ll, i, i, vp, vp, vp, vp, vp]
    L.upd_stg_conv1d.restype = ctypes.c_int
    L.upd_stg_conv1d.argtypes = [vp, vp, vp, ll, i, i, i, i, i, i, i, vp, vp]
    _LIB = L
    return L


# C-ABI calls made so far, by entry point (bench.py reports its own kernel launches from this; every launcher in the
# package funnels its return code through check()).
CALLS = {}
KERNELS_PER_CALL = {"upd_mpv_reduce": 2, "upd_denoiser_pack": 0}


def kernel_launches():
    """Number of this library's kernels launched so far (upd_mpv_reduce = Welford + window means)."""
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in CALLS.items())


def check(code, what):
    CALLS[what] = CALLS.get(what, 0) + 1
    if code != 0:
        L = lib()
        msg = L.upd_error_string(code).decode()
        if code == 3:
            msg += " (cudaError {})".format(L.upd_last_cuda_error())
        raise RuntimeError("{} failed: {}".format(what, msg))


def ptr(t):
    """Device/host pointer of a contiguous fp32 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def dev_f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(fn):
    """Decorator for model methods that launch kernels: run with the CUDA device of ``self``'s parameters current, so
    that a model living on cuda:1 while cuda:0 is current launches on cuda:1 (the C ABI validates and launches on the
    CURRENT device with the stream it is handed)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapper


def require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("this package runs on CUDA (sm_100a) only; got device {!r} (no CPU fallback)".format(str(device)))
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available (no CPU fallback)")
    return device
