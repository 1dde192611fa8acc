"""GPU parity of the DiffSTG sampler against fixtures made by the unmodified reference (with the
published-definition stand-in for torch_geometric's ResGatedGraphConv: that one layer is "parity unpinned").
Tolerances: eps prediction 1e-4 of its rms; whole evaluation_step (20 DDIM steps + the DDPM tail, injected noise)
rel 1e-3 per value + 1e-4 x rms floor, as the north star states for trajectories."""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import diffusionts_oracle as dto, diffstg_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load(name):
    g = np.load("{}/{}".format(GOLDEN, name))
    return g, json.loads(str(g["cfg"])), json.loads(str(g["keys"])), int(g["seed"])


def _weights(shapes, seed):
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    return sd


def _model(cfg, shapes, seed, **over):
    from updgm_b200.diffstg import DiffSTG
    m = DiffSTG(dict(cfg, device=DEV, **over)).eval()
    own = {k: list(v.shape) for k, v in m.state_dict().items() if k.startswith("model.")}
    assert own == shapes
    sd = _weights(shapes, seed)
    sd["scaler_mean"], sd["scaler_std"] = torch.zeros(cfg["F"]), torch.ones(cfg["F"])
    m.load_state_dict(sd, strict=True)
    return m, sd


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt())


@pytest.mark.parametrize("name", ["stg_small_evalstep.npz", "stg_yaml_evalstep.npz"])
def test_eps_prediction_matches_reference(name):
    g, cfg, shapes, seed = _load(name)
    m, _ = _model(cfg, shapes, seed)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V = x.shape[0]
    xm = torch.cat([x, torch.zeros(V, cfg["T_p"], 1)], 1).to(DEV)
    for t in (1, cfg["diffusion_steps"]):
        e = m.predict_eps(torch.from_numpy(g["eps%d:xt" % t]).to(DEV), xm, t, ei, V)
        assert _rel(e, torch.from_numpy(g["eps%d:out" % t])) < 1e-4, (t, _rel(e, torch.from_numpy(g["eps%d:out" % t])))


@pytest.mark.parametrize("name", ["stg_small_evalstep.npz", "stg_yaml_evalstep.npz"])
def test_evaluation_step_matches_reference(name):
    from updgm_b200.diffstg import GraphData
    g, cfg, shapes, seed = _load(name)
    m, _ = _model(cfg, shapes, seed)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V = x.shape[0]
    draws = [torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))]
    per = m.draws_per_round()
    assert per * cfg["sequential_sampling"] == len(draws)
    noise = [draws[r * per:(r + 1) * per] for r in range(cfg["sequential_sampling"])]
    outs, truth = m.evaluation_step(GraphData(x=x.to(DEV), edge_index=ei, num_nodes=V), noise=noise)
    ref = torch.from_numpy(g["outs"])
    assert truth is None and tuple(outs.shape) == tuple(ref.shape) and outs.device.type == "cpu"
    rms = ref.double().pow(2).mean().sqrt()
    d = (outs.double() - ref.double()).abs()
    assert (d <= 1e-3 * ref.abs().double() + 1e-4 * rms).all(), float(d.max() / rms)


def test_batched_replicas_match_single_launches():
    """Windows / rounds / replicas are all replicas of the graph: results must not depend on how they are cut."""
    g, cfg, shapes, seed = _load("stg_small_evalstep.npz")
    m, _ = _model(cfg, shapes, seed)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V = x.shape[0]
    wins = torch.stack([x, x.flip(0), x * 0.5], 0).to(DEV)
    a = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    m.rows_per_launch = V            # one replica per launch
    b = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    c = m.sample_windows(wins[1:], ei, V, seed=5, window_base=8)
    K = cfg["parallel_sampling"] * cfg["sequential_sampling"]
    assert tuple(a.shape) == (3 * V, K, cfg["T_h"] + cfg["T_p"], 1)
    assert _rel(a, b) < 1e-5 and _rel(a[V:], c) < 1e-5
    assert float(a.var(dim=1).mean()) > 0


def test_gated_aggregate_kernel_against_oracle():
    from updgm_b200 import _lib
    torch.manual_seed(0)
    V, C, reps = 7, 48, 3
    ei = torch.tensor([[0, 1, 2, 3, 4, 5, 6, 0, 2, 6, 6], [1, 2, 3, 4, 5, 6, 0, 3, 0, 1, 6]])
    sd = {"g.lin_%s.weight" % n: torch.randn(C, C) * 0.2 for n in ("key", "query", "value", "skip")}
    sd.update({"g.lin_%s.bias" % n: torch.randn(C) * 0.1 for n in ("key", "query", "value")})
    sd["g.bias"] = torch.randn(C) * 0.1
    x = torch.randn(reps * V, C)
    ref = torch.relu(so.res_gated_graph_conv(sd, "g.", x, so.duplicate_edge_index(reps, ei, V)))
    from updgm_b200.diffstg import graph_csr
    rowptr, col = graph_csr(ei, V)
    w = torch.cat([sd["g.lin_key.weight"], sd["g.lin_query.weight"], sd["g.lin_value.weight"], sd["g.lin_skip.weight"]], 0)
    bb = torch.cat([sd["g.lin_key.bias"], sd["g.lin_query.bias"], sd["g.lin_value.bias"], torch.zeros(C)], 0)
    kqvs = torch.addmm(bb.to(DEV), x.to(DEV), w.to(DEV).t()).contiguous()
    out = torch.empty(reps * V, C, device=DEV)
    bias = sd["g.bias"].to(DEV)
    rp, cl = rowptr.to(DEV), col.to(DEV)
    _lib.check(_lib.lib().upd_stg_gated_aggregate(_lib.ptr(kqvs), _lib.ptr(rp), _lib.ptr(cl), _lib.ptr(bias), reps * V, V, C,
                                                  1, _lib.ptr(out), _lib.stream_ptr(torch.device(DEV))), "agg")
    assert _rel(out, ref) < 1e-5


@pytest.mark.parametrize("C,CI,T", [(8, 4, 400), (16, 32, 200), (4, 12, 40), (16, 16, 20), (8, 4, 2000), (16, 32, 1200),
                                    (16, 32, 50), (8, 16, 514), (4, 4, 6), (8, 8, 400), (8, 24, 100), (16, 8, 100), (16, 12, 36),
                                    (8, 9, 12), (16, 17, 510)])
def test_fused_tcn_layernorm_kernel_against_library_ops(C, CI, T):
    """upd_stg_tcn_ln == causal conv -> causal conv -> LayerNorm over channels (torch fp32, TF32 off)."""
    import torch.nn.functional as F
    from updgm_b200 import _lib
    torch.manual_seed(C * 1000 + T)
    N = 37
    x = torch.randn(N, CI, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    wsc, sc = torch.randn(C, CI, device=DEV) * 0.3, torch.empty(N, C, T, device=DEV)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h = F.conv1d(F.pad(x, (2, 0)), w1, b1)
        h = F.conv1d(F.pad(h, (2, 0)), w2, b2)
    ref = F.layer_norm(h.transpose(1, 2), (C,), g, be).transpose(1, 2)
    out = torch.empty(N, C, T, device=DEV)
    _lib.check(_lib.lib().upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g),
                                         _lib.ptr(be), N, CI, C, T, _lib.ptr(out), None, _lib.ptr(wsc), _lib.ptr(sc), _lib.stream_ptr(torch.device(DEV))), "tcn")
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    assert _rel(sc, torch.matmul(wsc, x)) < 1e-5
    # the same row emitted as the fp16 split operand [hi | lo | hi | 1 1 0..] of the GEMM that follows
    K = C * T
    a3 = torch.empty(N, 3 * K + 8, dtype=torch.float16, device=DEV)
    _lib.check(_lib.lib().upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g),
                                         _lib.ptr(be), N, CI, C, T, None, _lib.ptr(a3), None, None, _lib.stream_ptr(torch.device(DEV))), "tcn")
    val = a3[:, :K].float() + a3[:, K:2 * K].float()
    assert torch.equal(a3[:, :K], a3[:, 2 * K:3 * K]) and float(a3[:, 3 * K].min()) == 1.0 and float(a3[:, 3 * K + 2:].abs().max()) == 0.0
    assert _rel(val, ref.reshape(N, K)) < 2e-5
    assert _lib.lib().upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g),
                                     _lib.ptr(be), N, CI, 5, T, _lib.ptr(out), None, None, None, None) == 2      # UPD_ERR_UNSUPPORTED


@pytest.mark.parametrize("C,CI,T,N", [(16, 32, 50, 9), (8, 16, 514, 5), (4, 4, 6, 3), (8, 4, 2000, 4), (16, 16, 1028, 8200),
                                      (16, 16, 200, 700), (8, 24, 50, 9), (8, 8, 398, 11)])
def test_fused_tcn_kernel_writes_stay_inside_their_buffers(C, CI, T, N):
    """Own bounds check (no sanitizer on this pool): every output sits between sentinel guard bands that must survive, for the
    scalar (T % 4 != 0), segmented (T > 512) and 8-rows-per-CTA (N >= 8192) paths."""
    from updgm_b200 import _lib
    torch.manual_seed(T)
    G = 256
    x = torch.randn(N, CI, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be, wsc = torch.randn(C, device=DEV), torch.randn(C, device=DEV), torch.randn(C, CI, device=DEV)
    K = C * T

    def guarded(n, dtype):
        buf = torch.full((n + 2 * G,), 777.0, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]

    hn_buf, hn = guarded(N * K, torch.float32)
    sc_buf, sc = guarded(N * K, torch.float32)
    a3_buf, a3 = guarded(N * (3 * K + 8), torch.float16)
    st = _lib.stream_ptr(torch.device(DEV))
    args = [_lib.ptr(v) for v in (x, w1, b1, w2, b2, g, be)]
    _lib.check(_lib.lib().upd_stg_tcn_ln(*args, N, CI, C, T, _lib.ptr(hn), None, _lib.ptr(wsc), _lib.ptr(sc), st), "tcn")
    _lib.check(_lib.lib().upd_stg_tcn_ln(*args, N, CI, C, T, None, _lib.ptr(a3), None, None, st), "tcn")
    torch.cuda.synchronize()
    for buf in (hn_buf, sc_buf, a3_buf):
        assert bool((buf[:G] == 777.0).all()) and bool((buf[-G:] == 777.0).all())
    assert not bool((hn == 777.0).any()) and torch.isfinite(hn).all() and torch.isfinite(sc).all()
    row = a3.view(N, 3 * K + 8)
    assert _rel(row[:, :K].float() + row[:, K:2 * K].float(), hn.view(N, K)) < 2e-5
    assert float(row[:, 3 * K].min()) == 1.0 and float(row[:, 3 * K + 2:].abs().max()) == 0.0
    assert _rel(sc.view(N, C, T), torch.matmul(wsc, x)) < 1e-5


@pytest.mark.parametrize("C,CI1,CI2,T,N", [(16, 16, 16, 200, 33), (8, 16, 8, 50, 7), (4, 8, 4, 1200, 5)])
def test_fused_tcn_kernel_reads_a_skip_concatenation_in_place(C, CI1, CI2, T, N):
    """upd_stg_tcn_ln_cat(x1, x2) == upd_stg_tcn_ln(cat(x1, x2)) bit for bit (same arithmetic, two source tensors)."""
    from updgm_b200 import _lib
    torch.manual_seed(CI1 * 100 + T)
    CI = CI1 + CI2
    x1, x2 = torch.randn(N, CI1, T, device=DEV), torch.randn(N, CI2, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be, wsc = torch.randn(C, device=DEV), torch.randn(C, device=DEV), torch.randn(C, CI, device=DEV)
    st = _lib.stream_ptr(torch.device(DEV))
    tail = [_lib.ptr(v) for v in (w1, b1, w2, b2, g, be)]
    K = C * T
    ref_hn, ref_sc, hn, sc = (torch.empty(N, K, device=DEV) for _ in range(4))
    ref_a3, a3 = (torch.empty(N, 3 * K + 8, dtype=torch.float16, device=DEV) for _ in range(2))
    xc = torch.cat((x1, x2), dim=1).contiguous()
    L = _lib.lib()
    _lib.check(L.upd_stg_tcn_ln(_lib.ptr(xc), *tail, N, CI, C, T, _lib.ptr(ref_hn), None, _lib.ptr(wsc), _lib.ptr(ref_sc), st), "a")
    _lib.check(L.upd_stg_tcn_ln(_lib.ptr(xc), *tail, N, CI, C, T, None, _lib.ptr(ref_a3), None, None, st), "b")
    _lib.check(L.upd_stg_tcn_ln_cat(_lib.ptr(x1), CI1, _lib.ptr(x2), CI2, *tail, N, C, T, _lib.ptr(hn), None, _lib.ptr(wsc),
                                    _lib.ptr(sc), st), "c")
    _lib.check(L.upd_stg_tcn_ln_cat(_lib.ptr(x1), CI1, _lib.ptr(x2), CI2, *tail, N, C, T, None, _lib.ptr(a3), None, None, st), "d")
    assert torch.equal(hn, ref_hn) and torch.equal(sc, ref_sc) and torch.equal(a3, ref_a3)
    assert L.upd_stg_tcn_ln_cat(_lib.ptr(x1), CI1, None, CI2, *tail, N, C, T, _lib.ptr(hn), None, None, None, st) == 1   # bad arg


@pytest.mark.gpu
@pytest.mark.parametrize("CI,CO,T,k,stride,pad,transposed", [(1, 4, 400, 1, 1, 0, False), (4, 1, 400, 1, 1, 0, False), (8, 8, 400, 3, 2, 1, False),
                                                            (8, 8, 200, 4, 2, 1, True), (16, 16, 50, 4, 2, 1, True), (12, 20, 37, 3, 2, 1, False),
                                                            (3, 5, 9, 4, 2, 1, True)])
def test_narrow_convolution_kernel_against_library_ops(CI, CO, T, k, stride, pad, transposed):
    """upd_stg_conv1d == F.conv1d / F.conv_transpose1d (fp32, TF32 off): x_proj / out.0 / DownSample / UpSample shapes."""
    import torch.nn.functional as F
    from updgm_b200.diffstg import conv1d_time
    torch.manual_seed(CI * 100 + CO)
    N = 300
    x = torch.randn(N, CI, T, device=DEV)
    w = torch.randn(*((CI, CO, k) if transposed else (CO, CI, k)), device=DEV) * 0.3
    b = torch.randn(CO, device=DEV)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = (F.conv_transpose1d if transposed else F.conv1d)(x, w, b, stride=stride, padding=pad)
    out = conv1d_time(x, w, b, k, stride=stride, pad=pad, transposed=transposed)
    assert tuple(out.shape) == tuple(ref.shape)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)


@pytest.mark.gpu
def test_gated_aggregate_kernel_full_size_graph_against_oracle():
    """The shared-memory form of upd_stg_gated_aggregate at the BASELINE graph size (100 nodes, ~2300 directed edges, 160
    channels = five 32-channel slabs) against the oracle's scatter-add form."""
    from updgm_b200 import _lib
    from updgm_b200.diffstg import graph_csr
    import networkx as nx
    torch.manual_seed(1)
    V, C, reps = 100, 160, 5
    G = nx.barabasi_albert_graph(V, 12, seed=0)
    ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
    sd = {"g.lin_%s.weight" % n: torch.randn(C, C) * 0.1 for n in ("key", "query", "value", "skip")}
    sd.update({"g.lin_%s.bias" % n: torch.randn(C) * 0.1 for n in ("key", "query", "value")})
    sd["g.bias"] = torch.randn(C) * 0.1
    x = torch.randn(reps * V, C)
    ref = torch.relu(so.res_gated_graph_conv(sd, "g.", x, so.duplicate_edge_index(reps, ei, V)))
    rowptr, col = graph_csr(ei, V)
    w = torch.cat([sd["g.lin_key.weight"], sd["g.lin_query.weight"], sd["g.lin_value.weight"], sd["g.lin_skip.weight"]], 0)
    bb = torch.cat([sd["g.lin_key.bias"], sd["g.lin_query.bias"], sd["g.lin_value.bias"], torch.zeros(C)], 0)
    kqvs = torch.addmm(bb, x, w.t()).to(DEV).contiguous()         # the projections in fp32 on the host: only the aggregation is under test
    out = torch.empty(reps * V, C, device=DEV)
    bias, rp, cl = sd["g.bias"].to(DEV), rowptr.to(DEV), col.to(DEV)
    _lib.check(_lib.lib().upd_stg_gated_aggregate(_lib.ptr(kqvs), _lib.ptr(rp), _lib.ptr(cl), _lib.ptr(bias), reps * V, V, C,
                                                  1, _lib.ptr(out), _lib.stream_ptr(torch.device(DEV))), "agg")
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
