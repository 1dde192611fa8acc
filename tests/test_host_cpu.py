"""CPU-side tests: host logic of the product package (no GPU compute), the C-ABI library's exports,
bit-exact window bookkeeping and schedule tables against the reference-made fixtures, and the N>1 gather
path under gloo with world_size 2."""
import ctypes
import json
import os
import re
import struct
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden, load_wo_fx_checkpoint

import updgm_b200
from updgm_b200 import _build, _lib, kernels, schedules, uncertainty as U


# ------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "upd_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(upd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_build.build_library())
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().upd_abi_version() == _lib.ABI_VERSION == 8
    assert _lib.lib().upd_error_string(2) == b"unsupported shape"


def test_library_is_sm100a_tcgen05_code():
    """The built library carries sm_100a SASS with the Blackwell mnemonics (no GPU needed to check)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_pack_sizes_and_limits():
    L = _lib.lib()
    assert L.upd_denoiser_pack_bytes(0, 2, 20) > 4 * 128 * 128 * 2
    assert L.upd_denoiser_pack_bytes(0, 5, 20) == 0 and L.upd_denoiser_pack_bytes(0, 2, 1) == 0
    assert L.upd_denoiser_pack_bytes(2, 2, 20) == 0
    assert L.upd_denoiser_pack_bytes(0, 2, 20) % 128 == 0


def test_pack_layout_roundtrip(wo_fx):
    """Unpack the blob by the documented layout: fp32 transposes exact, fp16 hi+lo = W*wscale to 2^-21."""
    _, sd = wo_fx
    tab = schedules.nsdiff_tables("linear", 20, 1e-4, 0.02)
    rows = schedules.stack_rows(tab, schedules.NSDIFF_ROWS)
    blob = kernels.pack_denoiser(sd, kernels.KIND_NSDIFF, 2, 20, rows, device=None).numpy().tobytes()
    u = 4 * 128 * 128 * 2
    K1 = 8
    off = u + 2 * 128 * K1 * 4
    b2 = np.frombuffer(blob, np.float32, 128, off)
    assert np.array_equal(b2, sd[kernels.DEN + "lin2.lin.bias"].numpy())
    W2 = sd[kernels.DEN + "lin2.lin.weight"].numpy()
    hi = np.frombuffer(blob, np.float16, 128 * 128, 0).reshape(16, 128, 8)          # [k/8][n][k%8]
    lo = np.frombuffer(blob, np.float16, 128 * 128, 128 * 128 * 2).reshape(16, 128, 8)
    rec = (hi.astype(np.float64) + lo.astype(np.float64)).transpose(1, 0, 2).reshape(128, 128)
    scale = rec[np.abs(W2) > 1e-3] / W2[np.abs(W2) > 1e-3]
    s = np.median(scale)
    assert s == 2.0 ** round(np.log2(s)), "wscale must be a power of two"
    assert np.abs(rec / s - W2).max() <= np.abs(W2).max() * 2.0 ** -21
    # tf32 lin1 block: column IN carries the bias
    W1 = sd[kernels.DEN + "lin1.lin.weight"].numpy()
    u1hi = np.frombuffer(blob, np.float32, 128 * K1, u).reshape(K1 // 4, 128, 4).transpose(1, 0, 2).reshape(128, K1)
    u1lo = np.frombuffer(blob, np.float32, 128 * K1, u + 128 * K1 * 4).reshape(K1 // 4, 128, 4).transpose(1, 0, 2).reshape(128, K1)
    assert np.allclose(u1hi[:, :6] + u1lo[:, :6], W1, rtol=0, atol=np.abs(W1).max() * 2.0 ** -20)
    assert np.allclose(u1hi[:, 6] + u1lo[:, 6], sd[kernels.DEN + "lin1.lin.bias"].numpy(), rtol=0, atol=1e-7)
    assert np.all(u1hi[:, 7] == 0)
    assert np.all((u1hi.view(np.uint32) & 0x1FFF) == 0), "hi parts are tf32-representable"
    # schedule table sits verbatim in the image
    raw = np.frombuffer(blob, np.float32)
    pos = [i for i in range(0, len(raw) - 200) if raw[i] == rows[0, 0].item() and raw[i + 1] == rows[0, 1].item()]
    assert any(np.array_equal(raw[p:p + 200], rows.numpy().ravel()) for p in pos)


def test_pack_rejects_mismatched_state_dict(wo_fx):
    _, sd = wo_fx
    rows = schedules.stack_rows(schedules.nsdiff_tables(), schedules.NSDIFF_ROWS)
    with pytest.raises(ValueError, match="dataset_nf"):
        kernels.pack_denoiser(sd, kernels.KIND_NSDIFF, 1, 20, rows, device=None)
    rows10 = schedules.stack_rows(schedules.nsdiff_tables("linear", 10), schedules.NSDIFF_ROWS)
    with pytest.raises(ValueError, match="rows"):
        kernels.pack_denoiser(sd, kernels.KIND_NSDIFF, 2, 10, rows10, device=None)


# ------------------------------------------------------------------------------------------ schedules
def test_product_schedule_tables_bit_exact_with_reference():
    g = load_golden("nsdiff_schedule_T20_linear.npz")
    t = schedules.nsdiff_tables("linear", 20, 1e-4, 0.02)
    for k, v in g.items():
        assert torch.equal(t[k], v), k
    g = load_golden("tmdm_loop_randF1.npz")
    t = schedules.tmdm_tables("linear", 20, 1e-4, 0.02)
    assert torch.equal(t["alphas"], g["alphas"]) and torch.equal(t["one_minus_alphas_bar_sqrt"], g["one_minus_alphas_bar_sqrt"])


@pytest.mark.parametrize("name", ["linear", "const", "quad", "jsd", "sigmoid", "cosine", "cosine_reverse", "cosine_anneal"])
def test_all_schedule_names_match_oracle(name):
    from oracle import nsdiff_oracle
    a = schedules.beta_schedule(name, 20, 1e-4, 0.02)
    b = nsdiff_oracle.make_beta_schedule(name, 20, 1e-4, 0.02).float()
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        schedules.beta_schedule("nope", 20, 1e-4, 0.02)


# ------------------------------------------------------------------------------------------ windows
def test_window_bookkeeping_bit_exact_with_reference_fixture():
    with open(os.path.join(GOLDEN, "windows.json")) as f:
        cases = json.load(f)
    for c in cases:
        if c["kind"] == "interval":
            assert U.sampling_interval_from_t(c["sampling_t"]) == c["interval"]
        elif c["kind"] == "infer_step":
            assert U.infer_sample_window_step_from_cache(*c["args"]) == c["step"]
        elif c["kind"] == "count":
            assert U.sliding_window_count(*c["args"]) == c["count"]
        elif c["kind"] == "network":
            n, t, f = c["shape"]
            series = torch.arange(n * t, dtype=torch.float32).reshape(n, t, f)
            ss, st = U.sample_time_series(series, np.arange(t) * 0.1, c["sampling_t"])
            assert ss.shape[1] == c["sampled_len"]
            wins, tps = U.build_sliding_windows(ss, st, c["windows"], c["step"])
            assert len(wins) == c["n_windows"] and len(tps) == c["n_time_points"]
            assert [float(w[0, 0, 0]) for w in wins[:4]] + [float(wins[-1][0, 0, 0])] == c["first_elems"]
            assert float(wins[-1][2, -1, 0]) == c["last_elem_of_last"]
            assert [float(x) for x in tps[:3]] == c["time_points_head"] and float(tps[-1]) == c["time_points_tail"]
            assert tuple(wins[0].shape) == (n, c["windows"], f)
            stacked = U.stacked_sliding_windows(ss, c["windows"], c["step"])
            assert torch.equal(stacked[5], wins[5])
        elif c["kind"] == "slbp":
            t, f = c["shape"]
            raw = torch.arange(t * f, dtype=torch.float32).reshape(t, f)
            ins, tgts, tps = U.build_slbp_sensitivity_windows(raw, np.arange(t), c["windows"], c["pred_len"],
                                                              c["sampling_t"], c["step"])
            assert len(ins) == c["n_inputs"] and len(tgts) == c["n_targets"] and len(tps) == c["n_time_points"]
            assert [float(w[0, 0]) for w in ins[:3]] == c["input_first"]
            assert [float(w[0, 0]) for w in tgts[:3]] == c["target_first"]
            assert tuple(ins[0].shape) == (c["windows"], f)


def test_window_edge_cases():
    with pytest.raises(ValueError, match="shorter than windows"):
        U.build_sliding_windows(torch.zeros(2, 50, 1), np.arange(50), 100, 5)
    with pytest.raises(ValueError, match=r"\[Node_num, T_obs_num, F\]"):
        U.build_sliding_windows(torch.zeros(50, 1), np.arange(50), 10, 5)
    wins, tps = U.build_sliding_windows(torch.zeros(2, 100, 1), np.arange(100), 100, 5)
    assert len(wins) == 1 and list(tps) == [99]
    ins, tgts, _ = U.build_slbp_sensitivity_windows(torch.zeros(30000, 2), np.arange(30000), 20, 20, 100, 10)
    assert len(ins) == 2 and tgts == ()                      # 30 sampled points: tail shorter than pred_len
    assert U.normalize_time_series(torch.zeros(7, 3), "SIS").shape == (3, 7, 1)
    assert U.normalize_time_series(torch.zeros(7, 3), "SLBP").shape == (1, 7, 3)
    with pytest.raises(ValueError):
        U.normalize_time_series(torch.zeros(7), "SLBP")


def test_cache_names_and_sniffing(tmp_path):
    U.set_project_root(tmp_path)
    try:
        assert U.resolve_cache_path(None, tmp_path / "m", "d/x_increase.pt", "SIS") == tmp_path / "m" / "x_increase.pt"
        assert U.resolve_cache_path(None, tmp_path / "m", "d/x.pt", "SIS", suffix="_gx").name == "x_gx.pt"
        assert U.resolve_cache_path("rel/dir", None, None, "SIS") == tmp_path / "rel/dir" / "data.pt"
        assert U.resolve_cache_path("rel/file.pt", None, "a.pt", "SIS") == tmp_path / "rel/file.pt"
        assert U.default_cache_dir(None, "sis") == tmp_path / "ews_results/model_uncertainy_cache/model/SIS"
        assert U.slbp_sensitivity_cache_path("r", "n", "increase", 10) == tmp_path / "r/datas/n_pred_future_increase_10.pt"
        assert U.slbp_fig6_cache_path("r", "n", "up", 5, "sub", "gx") == tmp_path / "r/datas/sub/n_gx_up_5.pt"
        assert U.slbp_fig6_pred_future_gx_cache_path("r", "n", "up", 5).name == "n_pred_future_up_5_gx.pt"
        assert U._legacy_single_underscore_model_name("dataset__w200") == "dataset_w200"
    finally:
        U.set_project_root(os.getcwd())
    p = tmp_path / "c.pt"
    torch.save({"not": "a list"}, p)
    with pytest.raises(TypeError, match="list of tensors"):
        U._load_tensor_list(p)
    assert U.read_sensitivity_pred_future_cache(tmp_path / "missing.pt") is None
    U._save_tensor_list([torch.zeros(4, 2)], tmp_path / "deep" / "g.pt")
    assert U._slbp_cache_elements_are_gx(U._load_tensor_list(tmp_path / "deep" / "g.pt"))
    assert U._slbp_cache_elements_are_gx([torch.zeros(1, 4, 2)]) and not U._slbp_cache_elements_are_gx([torch.zeros(3, 4, 2)])
    assert U._slbp_cache_elements_have_ndim([torch.zeros(4, 2, 8)], 3) and not U._slbp_cache_elements_have_ndim([], 3)
    assert U.normalize_diffstg_pred_future_list([torch.zeros(5, 7, 3)])[0].shape == (5, 7, 1, 3)
    with pytest.raises(ValueError, match="Unsupported SLBP MPV cache"):
        U.summarize_slbp_mpv_cache_for_fig5([torch.zeros(2, 3, 4, 5)])


def test_uncertainty_ews_argument_errors(tmp_path):
    with pytest.raises(ValueError, match="sampling, gx, both"):
        U.uncertainty_ews(uncertainty_method="bogus")
    with pytest.raises(ValueError, match="Provide data_file or torch_time_series"):
        U.uncertainty_ews()
    with pytest.raises(ValueError, match="time_data is required"):
        U.uncertainty_ews(torch_time_series=torch.zeros(2, 50, 1), dynamic_type="SIS")
    with pytest.raises(ValueError, match="dataset.windows"):
        U.uncertainty_ews(torch_time_series=torch.zeros(2, 50, 1), time_data=np.arange(50), dynamic_type="SIS")
    with pytest.raises(FileNotFoundError):
        U.read_model_config(tmp_path)


# ------------------------------------------------------------------------------------------ no CPU fallback
def test_product_fails_loudly_without_cuda(wo_fx):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    ckpt = os.path.join(GOLDEN, "ews_results", "NsDiff_machine", "wo_fx")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        U.load_model_from_dir(ckpt, device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.mpv_reduce(torch.zeros(1, 2, 3, 1), 1, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unsupervised-probing-using-generative-diffusion-models_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "/root/reference" not in text, fn


# ------------------------------------------------------------------------------------------ model objects (CPU construction)
def test_model_state_dict_keys_match_reference_checkpoint(wo_fx):
    """The parameter containers reproduce the checkpoint's key set and shapes exactly (strict load)."""
    from updgm_b200.nsdiff import NsDiff_model_variants
    net_param, sd = wo_fx
    p = dict(net_param, device="cpu")
    m = NsDiff_model_variants(p, "cond_var")
    own = m.state_dict()
    assert set(own) == set(sd)
    for k in sd:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    assert m.cond_pred_model is None and m.cond_pred_model_g is not None
    assert m.scaler == "StandardScaler" and m.label_len == 100
    x = torch.randn(3, 200, 2)
    assert torch.equal(m.scaler_inverse_transform(m.scaler_transform(x)), (x - m.scaler_mean) / m.scaler_std * m.scaler_std + m.scaler_mean)
    with pytest.raises(ValueError):
        NsDiff_model_variants(dict(net_param, device="cpu"), "nope")


def test_factory_names():
    from updgm_b200 import loader
    with pytest.raises(ValueError, match="don't exit"):
        loader.diffusion_models("Nope", {})
    with pytest.raises(KeyError):                     # every task_model of models/models.py:5-32 is built; this one
        loader.diffusion_models("NsDiff_spatial", {})   # needs train_model_select like the reference's factory


# ------------------------------------------------------------------------------------------ N > 1 host path (gloo)
def test_partition_windows_covers_everything():
    for W in (0, 1, 5, 181, 199, 981):
        for world in (1, 2, 3, 8):
            blocks = [U.partition_windows(W, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == W
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= -(-W // world)


def _gloo_worker(rank, world, port, W, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w0, w1 = U.partition_windows(W, world, rank)
        local = torch.stack([torch.arange(w0, w1).float(), torch.arange(w0, w1).float() * 10 + rank * 0], dim=1)
        full = U.gather_window_stats(local, W)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W", [7, 8, 1])
def test_gather_window_stats_gloo_world2(W):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [[float(i), float(i) * 10] for i in range(W)]
    assert res[0] == want and res[1] == want


# ------------------------------------------------------------------------------------------ cache writer
def test_background_cache_writer_roundtrip(tmp_path, monkeypatch):
    """Cache files are written by a background thread, appear atomically, and every reader waits for them."""
    data = [torch.arange(12.).view(3, 4), torch.ones(2, 2)]
    p = tmp_path / "a" / "cache.pt"
    U._save_tensor_list(data, p)
    assert U._cache_ready(p)                                  # waits for the pending write
    back = U._load_tensor_list(p)
    assert torch.equal(back[0], data[0]) and torch.equal(back[1], data[1])
    assert not [f for f in os.listdir(p.parent) if ".tmp" in f]
    U._save_tensor_list([torch.zeros(1)], p)                  # overwrite, then read straight away
    assert torch.equal(U._load_tensor_list(p)[0], torch.zeros(1))
    U._save_tensor_list({"not": "a list"}, p)
    with pytest.raises(TypeError, match="must contain a list"):
        U._load_tensor_list(p)
    monkeypatch.setenv("UPD_SYNC_CACHE_WRITES", "1")
    q = tmp_path / "sync.pt"
    U._save_tensor_list(data, q)
    assert q.exists()
    U.flush_cache_writes()
    # a failing write surfaces at the next synchronisation point
    monkeypatch.delenv("UPD_SYNC_CACHE_WRITES")
    bad = tmp_path / "dir_in_the_way.pt"
    U._save_tensor_list(data, bad)
    U.flush_cache_writes()
    os.remove(bad)
    os.mkdir(bad)
    U._save_tensor_list(data, bad)
    with pytest.raises(OSError):
        U.flush_cache_writes()


def test_feature_inverse_transform_helper():
    """:267-283: feature axis is -2 for cache elements [.., O, F, K], -1 otherwise; identity without a model / scaler."""
    class M:
        scaler = "StandardScaler"
        scaler_mean, scaler_std = torch.tensor([1.0, -2.0]), torch.tensor([2.0, 0.5])
    x = torch.arange(24.).view(3, 2, 4)                       # [O, F, K]
    y = U._feature_inverse_transform(x, M())
    assert torch.equal(y[:, 0], x[:, 0] * 2.0 + 1.0) and torch.equal(y[:, 1], x[:, 1] * 0.5 - 2.0)
    z = torch.arange(6.).view(3, 2)                           # [.., F]
    assert torch.equal(U._feature_inverse_transform(z, M()), z * M.scaler_std + M.scaler_mean)
    assert U._feature_inverse_transform(x, None) is x and U._as_path(None) is None and str(U._as_path("a/b")) == "a/b"


def test_public_header_is_plain_c(tmp_path):
    """include/upd_b200.h is the drop-in boundary: it must compile as C (no C++ / CUDA / torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use.c"
    src.write_text('#include "upd_b200.h"\nint main(void) { return upd_abi_version() < 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_tmdm_checkpoint_with_temporal_embedding_tables_loads_strictly():
    """The reference's TMDM embeddings are torch-timeseries DataEmbedding(..., 'fixed', 'h', ...) (TMDM.py:90,
    tmdm_ns_transformer.py:53-56): a real checkpoint carries fixed sinusoidal time-mark tables the hot path never reads
    (x_mark is None).  They -- and nothing else -- are dropped before the strict load."""
    import yaml
    from updgm_b200 import loader
    from updgm_b200.tmdm import TMDM_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
    m = TMDM_model(dict(cfg["net"], device="cpu"))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    d = m.configs.d_model
    extra = {}
    for emb in ("cond_pred_model.enc_embedding", "cond_pred_model.dec_embedding", "model.enc_embedding"):
        for name, n in (("hour_embed", 24), ("weekday_embed", 7), ("day_embed", 32), ("month_embed", 13)):
            extra["{}.temporal_embedding.{}.emb.weight".format(emb, name)] = torch.zeros(n, d)
    full = dict(sd, **extra)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(full, strict=True)
    kept = loader.drop_unused_temporal_tables(m, full)
    assert set(kept) == set(sd)
    m.load_state_dict(kept, strict=True)
    # anything else unexpected still fails the strict load
    bad = dict(full, **{"model.diffussion_model.lin9.weight": torch.zeros(1)})
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(loader.drop_unused_temporal_tables(m, bad), strict=True)


def test_checkpoint_loader_is_tensor_only_by_default(tmp_path, monkeypatch):
    """loader._load_checkpoint reads the shipped checkpoint with the tensor-only unpickler and refuses a checkpoint that
    needs arbitrary classes unless UPD_ALLOW_PICKLE=1 opts into the reference's full unpickle (utils/utils.py:670)."""
    from updgm_b200 import loader
    st = loader._load_checkpoint(os.path.join(GOLDEN, "ews_results", "NsDiff_machine", "wo_fx", "model_trained"))
    assert set(st) == {"net_param", "state_dict"} and len(st["state_dict"]) == 25

    import fractions
    p = tmp_path / "model_trained"
    torch.save({"net_param": {"x": fractions.Fraction(1, 3)}, "state_dict": {}}, p)     # a class outside the tensor-only allowlist
    monkeypatch.delenv("UPD_ALLOW_PICKLE", raising=False)
    with pytest.raises(RuntimeError, match="UPD_ALLOW_PICKLE"):
        loader._load_checkpoint(str(p))
    monkeypatch.setenv("UPD_ALLOW_PICKLE", "1")
    assert loader._load_checkpoint(str(p))["net_param"]["x"] == fractions.Fraction(1, 3)


def test_samples_per_row_reproduces_the_reference_failure_modes():
    """NsDiff: range(K // S) chunks, remainder dropped, an empty loop fails in torch.cat (NsDiff_model.py:227-247);
    TMDM / DiffusionTS: S clamped to K, K % S must be 0 (tmdm_adapter.py:125-127, DiffusionTS_model.py:83-85)."""
    from types import SimpleNamespace

    class NsDiff_model_variants:      # only the class name and .configs matter to _samples_per_row
        def __init__(self, K, S):
            self.configs = SimpleNamespace(n_z_samples=K, parallel_sample=S)

    class TMDM_model(NsDiff_model_variants):
        pass

    assert U._samples_per_row(NsDiff_model_variants(100, 10)) == 100
    assert U._samples_per_row(NsDiff_model_variants(25, 10)) == 20
    with pytest.raises(RuntimeError, match="non-empty list"):
        U._samples_per_row(NsDiff_model_variants(5, 10))
    assert U._samples_per_row(TMDM_model(5, 10)) == 5
    with pytest.raises(ValueError, match="divisible"):
        U._samples_per_row(TMDM_model(25, 10))


def test_fresh_stats_are_dropped_when_the_cache_or_the_scaler_changes():
    """The device-computed statistics remembered by sample_sweep are only reused for exactly that, unmodified, cache and
    (for raw-unit statistics) the scaler they were baked with."""
    from types import SimpleNamespace
    cache = torch.zeros(3, 2, 4, 5, 1)
    model = SimpleNamespace(scaler_mean=torch.tensor([1.0]), scaler_std=torch.tensor([2.0]))
    stats = {"scaled": {"mpv": torch.arange(3.0)}, "raw": {"mpv": torch.arange(3.0) * 4}}
    U._FRESH_STATS.clear()
    import weakref
    U._FRESH_STATS[cache.untyped_storage().data_ptr()] = (weakref.ref(cache), cache._version, tuple(cache.shape),
                                                          U._scaler_snapshot(model), stats)
    elems = U._as_cache_list(cache)
    assert U._fresh_stats_for(elems, elems[0].shape) is stats
    assert U._fresh_stats_for(elems, elems[0].shape, model=model, need_scaler=True) is stats
    model.scaler_std = torch.tensor([3.0])                       # scaler_fit between sweep and summary
    assert U._fresh_stats_for(elems, elems[0].shape, model=model, need_scaler=True) is None
    assert U._fresh_stats_for(elems, elems[0].shape) is stats
    elems[1].mul_(2.0)                                           # in-place edit of one element: views share the version
    assert U._fresh_stats_for(elems, elems[0].shape) is None
    assert not U._FRESH_STATS


def test_polynomial_lg2_coefficients_in_the_sampler_source():
    """The warp-specialised sampler evaluates lg2(1 + u), u = 2^-|z| in (0, 1], as a Horner polynomial on the FMA pipe for
    part of the hidden units (csrc/sampler_math.cuh, softplus2_poly_b; the reference is F.softplus, denoise.py:47-51).
    The coefficients are read out of the source and evaluated in numpy, in exact arithmetic and as the kernel's chain of
    fp32 FMAs (float64 product-sum rounded to fp32 once per step): |error| <= 3.2e-7 / 4.5e-7 for degree 7 (the default),
    <= 1e-7 / 2.5e-7 for degree 8, over the whole argument range."""
    import re
    src = open(os.path.join(ROOT, "unsupervised-probing-using-generative-diffusion-models_b200", "csrc", "sampler_math.cuh")).read()
    body = src[src.index("softplus2_poly_b(float2 u, float2 z)"):]
    deg8, deg7 = body[body.index("#if UPD_LG2_DEG == 8"):body.index("#else")], body[body.index("#else"):body.index("#endif")]
    assert re.search(r"#define UPD_LG2_DEG 7", src), "default degree changed: update DESIGN 4.1 and the parity numbers"
    u = np.linspace(0.0, 1.0, 200001, dtype=np.float32)
    want = np.log2(1.0 + u.astype(np.float64))
    u64 = u.astype(np.float64)
    for text, n_coef, bound_exact, bound_fma in ((deg8, 9, 1e-7, 2.5e-7), (deg7, 8, 3.2e-7, 4.5e-7)):
        coef = [np.float32(c) for c in re.findall(r"splat\((-?[0-9.]+e[-+][0-9]+)f\)", text)]
        assert len(coef) == n_coef, (len(coef), n_coef)
        exact = np.full(u.shape, float(coef[0]))
        fma = np.full_like(u, coef[0])
        for c in coef[1:]:
            exact = exact * u64 + float(c)
            fma = (fma.astype(np.float64) * u64 + float(c)).astype(np.float32)
        assert np.abs(exact - want).max() <= bound_exact, (n_coef - 1, np.abs(exact - want).max())
        assert np.abs(fma.astype(np.float64) - want).max() <= bound_fma, (n_coef - 1, np.abs(fma.astype(np.float64) - want).max())
