"""DiffSTG model object with the reference's surface, backed by the GPU sampler.

Mirrors models/Diffusion_model/DiffSTG/graph_diffusion_model.py:103-282 (``DiffSTG``: constructor keys, ``scaler_*``,
``evaluation_step(data)`` with ``data.x [Node, T_h(+T_p), F]``, ``data.edge_index [2,E]``, ``data.num_nodes``) and keeps
every parameter name of ``UGnet`` (models/Diffusion_model/DiffSTG/ugnet.py:173-250; including TcnBlock's twice-registered
conv ``conv`` / ``net.0``) so reference checkpoints load with ``strict=True``.  Parameter containers only; the arithmetic
is restated for the GPU:

  * activations are [rows, C, T] (the reference's dummy "V = 1" image height is dropped: of each (3,k) kernel only
    the middle row ever touches data);
  * a TcnBlock is ONE causal conv: its 1x1 shortcut (or the identity) is folded into the last tap, the step embedding
    ``t_conv`` into its bias (all rows of a launch share the step);
  * the (1, T+1) down / up convolutions around the graph conv are dense [T*c <-> Td_h*c] GEMMs built at load time,
    emitting / consuming the spatial block's layout directly;
  * key / query / value / skip of the gated graph conv are one GEMM; gate, aggregation over in-neighbours, skip, bias
    and ReLU are one hand-written kernel working on the base graph's CSR for every replica
    (``upd_stg_gated_aggregate``); the posterior step is ``upd_stg_posterior``; Philox noise is ``upd_gauss_fill``;
  * windows, sequential rounds and parallel replicas are all just replicas of the graph: one launch covers as many as
    ``rows_per_launch`` allows.

GEMMs / convolutions are library calls in plain fp32 (TF32 disabled).  There is no CPU path.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .diffusionts import ParamTree
from .fx_encoder import _W3Cache, a3_split, gemm3

ROWS_PER_LAUNCH = 32768


class GraphData:
    """Minimal stand-in for torch_geometric.data.Data (the attributes evaluation_step reads)."""

    def __init__(self, x=None, edge_index=None, num_nodes=None):
        self.x, self.edge_index, self.num_nodes = x, edge_index, num_nodes

    def clone(self):
        return GraphData(None if self.x is None else self.x.clone(), self.edge_index.clone(), self.num_nodes)


def block_plan(net_param, T_total=None):
    """Execution-ordered (key prefix, kind, c_in, c_out, T_in) of UGnet (ugnet.py:190-239).  ``T_total``: length of the
    time axis the U-Net works on (DiffSTG: 2*(T_h+T_p), the default; NsDiff_spatial's UGnet: pred_len)."""
    d_h, mults, n_blocks = net_param["d_h"], net_param["channel_multipliers"], net_param["n_blocks"]
    if T_total is None:
        T_total = 2 * (net_param["T_p"] + net_param["T_h"])
    T_in = T_total
    n_res = len(mults)
    down, up = [], []
    out_c = in_c = d_h
    idx = 0
    for i in range(n_res):
        out_c = in_c * mults[i]
        for _ in range(n_blocks):
            down.append(("down.%d.res." % idx, "res", in_c, out_c, T_in))
            idx, in_c = idx + 1, out_c
        if i < n_res - 1:
            down.append(("down.%d." % idx, "downsample", in_c, in_c, T_in))
            idx += 1
            T_in = math.floor((T_in - 1) / 2 + 1)
    middle = [("middle.res1.", "res", out_c, out_c, T_in), ("middle.res2.", "res", out_c, out_c, T_in)]
    in_c, idx = out_c, 0
    for i in reversed(range(n_res)):
        out_c = in_c
        for _ in range(n_blocks):
            up.append(("up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
            idx += 1
        out_c = in_c // mults[i]
        up.append(("up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
        idx, in_c = idx + 1, out_c
        if i > 0:
            up.append(("up.%d." % idx, "upsample", in_c, in_c, T_in))
            idx += 1
            T_in = T_in * 2
    assert T_in == T_total, "T_in should be equal to T"
    return down, middle, up


def block_shapes(net_param, T_total=None):
    """Key -> shape of every U-Net block (without the input / output projections)."""
    d_h, Td_h = net_param["d_h"], net_param["Td_h"]
    if net_param["gnn_name"] != "ResGatedGraphConv":
        raise NotImplementedError("only gnn_name='ResGatedGraphConv' (every shipped DiffSTG YAML) is built")
    gp = net_param.get("gnn_param") or {}
    sh = {}
    down, middle, up = block_plan(net_param, T_total)
    for pre, kind, c_in, c_out, T_in in down + middle + up:
        if kind == "res":
            for tcn, ci in (("tcn1.", c_in), ("tcn2.", c_out)):
                for name in ("conv.", "net.0."):
                    sh[pre + tcn + name + "weight"], sh[pre + tcn + name + "bias"] = (c_out, ci, 3, 3), (c_out,)
                if ci != c_out:
                    sh[pre + tcn + "shortcut.weight"], sh[pre + tcn + "shortcut.bias"] = (c_out, ci, 1, 1), (c_out,)
            if c_in != c_out:
                sh[pre + "shortcut.weight"], sh[pre + "shortcut.bias"] = (c_out, c_in, 1, 1), (c_out,)
            sh[pre + "t_conv.weight"], sh[pre + "t_conv.bias"] = (c_out, d_h, 1, 1), (c_out,)
            sh[pre + "downsampling.weight"], sh[pre + "downsampling.bias"] = (c_out, c_out, 1, T_in + 1), (c_out,)
            sh[pre + "upsampling.weight"], sh[pre + "upsampling.bias"] = (c_out, c_out, 1, T_in + 1), (c_out,)
            C = Td_h * c_out
            for lin in ("lin_key", "lin_query", "lin_value"):
                sh[pre + "spatial.gnn." + lin + ".weight"], sh[pre + "spatial.gnn." + lin + ".bias"] = (C, C), (C,)
            if gp.get("root_weight", True):
                sh[pre + "spatial.gnn.lin_skip.weight"] = (C, C)
            if gp.get("bias", True):
                sh[pre + "spatial.gnn.bias"] = (C,)
            sh[pre + "norm.weight"], sh[pre + "norm.bias"] = (1, c_out), (1, c_out)
        elif kind == "downsample":
            sh[pre + "conv.weight"], sh[pre + "conv.bias"] = (c_in, c_in, 1, 3), (c_in,)
        else:
            sh[pre + "conv.weight"], sh[pre + "conv.bias"] = (c_in, c_in, 1, 4), (c_in,)
    return sh


def ugnet_shapes(net_param):
    F_, d_h = net_param["F"], net_param["d_h"]
    T = net_param["T_p"] + net_param["T_h"]
    sh = block_shapes(net_param)
    sh["x_proj.weight"], sh["x_proj.bias"] = (d_h, F_, 1, 1), (d_h,)
    sh["out.0.weight"], sh["out.0.bias"] = (F_, d_h, 1, 1), (F_,)
    sh["out.1.weight"], sh["out.1.bias"] = (T, 2 * T), (T,)
    return sh


def populate_ugnet(tree, shapes, alias):
    """Register UGnet's parameters on ``tree`` (a ParamTree) under the reference's key names, default-style init;
    TcnBlock registers its conv twice (``conv`` and ``net.0``: the same Parameter) -- ``alias(key, param)`` adds those."""
    gen = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 63))
    for key, shp in shapes.items():
        if ".net.0." in key:
            continue
        if ".norm." in key:
            w = torch.ones(shp) if key.endswith("weight") else torch.zeros(shp)
        elif key.endswith("spatial.gnn.bias"):
            w = torch.zeros(shp)
        else:
            fan_in = shp[0] if len(shp) == 1 else int(np.prod(shp[1:]))
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            w = torch.empty(shp).uniform_(-bound, bound, generator=gen)
        tree.add(key, w)
    for key in shapes:
        if ".net.0." in key:
            node = tree
            for part in key.replace(".net.0.", ".conv.").split("."):
                node = getattr(node, part) if not part.isdigit() else node._modules[part]
            alias(key, node)


def alias_parameter(tree, key, param):
    node = tree
    parts = key.split(".")
    for part in parts[:-1]:
        if part not in node._modules:
            node.add_module(part, ParamTree())
        node = node._modules[part]
    node.register_parameter(parts[-1], param)


class GaussianDiffusion:
    """float64 numpy schedule exactly as diffusion_schedulers.py:39-67 builds it."""

    def __init__(self, T, schedule, loss_weight_schedule="constant"):
        self.T, self.loss_weight_schedule = T, loss_weight_schedule
        if schedule == "linear":
            self.beta = np.linspace(1e-4, 2e-2, T)
        elif schedule == "quad":
            self.beta = np.linspace(1e-4 ** 0.5, 2e-2 ** 5, T) ** 2
        elif schedule == "cosine":
            def cos_noise(t):
                return np.cos(math.pi * 0.5 * (t / T + 0.008) / (1 + 0.008)) ** 2
            ab = cos_noise(np.arange(0, T + 1, 1)) / cos_noise(0)
            self.beta = np.clip(1 - (ab[1:] / ab[:-1]), None, 0.999)
        self.betabar = np.cumprod(self.beta)
        self.alpha = np.concatenate((np.array([1.0]), 1 - self.beta))
        self.alphabar = np.cumprod(self.alpha)


class InferenceSchedule:
    """diffusion_schedulers.py:95-125."""

    def __init__(self, inference_schedule="linear", T=1000, inference_T=1000):
        self.inference_schedule, self.T, self.inference_T = inference_schedule, T, inference_T

    def __call__(self, i):
        assert 0 <= i < self.inference_T
        if self.inference_schedule == "linear":
            t1 = self.T - int((float(i) / self.inference_T) * self.T)
            t2 = self.T - int((float(i + 1) / self.inference_T) * self.T)
        elif self.inference_schedule == "cosine":
            t1 = self.T - int(np.sin((float(i) / self.inference_T) * np.pi / 2) * self.T)
            t2 = self.T - int(np.sin((float(i + 1) / self.inference_T) * np.pi / 2) * self.T)
        else:
            raise ValueError("Unknown inference schedule: {}".format(self.inference_schedule))
        return np.clip(t1, 1, self.T), np.clip(t2, 0, self.T - 1)


def graph_csr(edge_index, num_nodes):
    """edge_index [2,E] (row 0 = source j, row 1 = target i) -> (rowptr [V+1], col [E]) int32 over targets, keeping
    the edge order inside each target (the order a sequential scatter-add visits them).  Integer, bit-exact."""
    ei = torch.as_tensor(edge_index).detach().cpu().reshape(2, -1).to(torch.int64)
    src, dst = ei[0], ei[1]
    if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= num_nodes):
        raise IndexError("edge_index refers to a node outside [0, num_nodes)")
    order = torch.sort(dst, stable=True).indices
    counts = torch.bincount(dst, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), src[order].to(torch.int32)


def gated_aggregate(kqvs, rowptr, col, bias, V, C, relu=True):
    """relu(sum_j sigmoid(k_i + q_j) * v_j + skip_i + bias) for every replica of the V-node graph (upd_stg_gated_aggregate)."""
    N = kqvs.shape[0]
    out = torch.empty((N, C), dtype=torch.float32, device=kqvs.device)
    rc = _lib.lib().upd_stg_gated_aggregate(_lib.ptr(kqvs), _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(bias), N, V, C,
                                            1 if relu else 0, _lib.ptr(out), _lib.stream_ptr(kqvs.device))
    _lib.check(rc, "upd_stg_gated_aggregate")
    return out


def conv1d_time(x, w, b, k, stride=1, pad=0, transposed=False):
    """Narrow convolution along the last axis of x [N, CI, Tin] (upd_stg_conv1d): w [CO, CI, k] (or [CI, CO, k] when
    ``transposed``), b [CO] -> [N, CO, Tout].  The 1x1 projections, DownSample and UpSample of UGnet."""
    x = x.contiguous()
    N, CI, Tin = x.shape
    CO = w.shape[1] if transposed else w.shape[0]
    Tout = (Tin - 1) * stride - 2 * pad + k if transposed else (Tin + 2 * pad - k) // stride + 1
    y = torch.empty((N, CO, Tout), dtype=torch.float32, device=x.device)
    rc = _lib.lib().upd_stg_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), N, CI, CO, Tin, k, stride, pad,
                                   1 if transposed else 0, _lib.ptr(y), _lib.stream_ptr(x.device))
    _lib.check(rc, "upd_stg_conv1d")
    return y


class PreparedUGnet:
    """Inference-time weight layout of UGnet (fp32 on the device, built once per model load)."""

    def __init__(self, model, net_param, diffusion_T, T_total=None):
        sd = {k: v.detach().to(torch.float32) for k, v in model.state_dict().items()}
        dev = sd["x_proj.weight"].device
        self.device, self.Td_h, self.d_h = dev, net_param["Td_h"], net_param["d_h"]
        self.down, self.middle, self.up = block_plan(net_param, T_total)
        d_h, Td = self.d_h, self.Td_h
        # TimeEmbedding (ugnet.py:15-33) for every diffusion step 0..T
        half = d_h // 2
        e = torch.exp(torch.arange(half, dtype=torch.float32, device=dev) * -(math.log(10000) / (half - 1)))
        e = torch.arange(diffusion_T + 1, dtype=torch.float32, device=dev)[:, None] * e[None, :]
        te = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
        if d_h % 2 == 1:
            te = F.pad(te, (0, 1, 0, 0))
        self.blocks = {}
        for pre, kind, c_in, c_out, T_in in self.down + self.middle + self.up:
            b = {}
            if kind == "res":
                for tcn, ci in (("tcn1.", c_in), ("tcn2.", c_out)):
                    w = sd[pre + tcn + "conv.weight"][:, :, 1, :].clone()            # [co, ci, 3], causal taps t-2..t
                    bias = sd[pre + tcn + "conv.bias"].clone()
                    if ci != c_out:
                        w[:, :, 2] += sd[pre + tcn + "shortcut.weight"][:, :, 0, 0]
                        bias += sd[pre + tcn + "shortcut.bias"]
                    else:
                        w[:, :, 2] += torch.eye(c_out, device=dev)
                    b[tcn + "w"], b[tcn + "b"] = w.contiguous(), bias
                # per-step bias of tcn1: conv bias (+ shortcut bias) + t_conv(time embedding)
                tvec = F.linear(te, sd[pre + "t_conv.weight"][:, :, 0, 0], sd[pre + "t_conv.bias"])   # [T+1, c_out]
                b["tcn1.b_step"] = (b["tcn1.b"][None, :] + tvec).contiguous()
                b["norm_w"], b["norm_b"] = sd[pre + "norm.weight"].reshape(-1), sd[pre + "norm.bias"].reshape(-1)
                pad = Td // 2
                s = torch.arange(T_in, device=dev)[:, None]
                tau = torch.arange(Td, device=dev)[None, :]
                k = s - tau + pad                                                   # [T_in, Td]
                ok = ((k >= 0) & (k <= T_in)).to(torch.float32)
                kc = k.clamp(0, T_in)
                wd = sd[pre + "downsampling.weight"][:, :, 0, :]                     # [co, ci, T+1]
                md = wd[:, :, kc] * ok                                              # [co, ci, s, tau]
                b["down_w"] = md.permute(1, 2, 3, 0).reshape(c_out * T_in, Td * c_out).contiguous()   # (ci,s) -> (tau,co)
                b["down_b"] = sd[pre + "downsampling.bias"].repeat(Td)
                wu = sd[pre + "upsampling.weight"][:, :, 0, :]                       # [ci, co, T+1]
                mu = wu[:, :, kc] * ok                                              # [ci, co, s, tau]
                b["up_w"] = mu.permute(3, 0, 1, 2).reshape(Td * c_out, c_out * T_in).contiguous()     # (tau,ci) -> (co,s)
                up_b = sd[pre + "upsampling.bias"]
                if c_in != c_out:
                    up_b = up_b + sd[pre + "shortcut.bias"]
                    b["sc_w2"] = sd[pre + "shortcut.weight"][:, :, 0, 0].contiguous()
                b["up_b_full"] = up_b.repeat_interleave(T_in).contiguous()               # [c*T], (co, s) order
                g = pre + "spatial.gnn."
                C = Td * c_out
                skip = sd.get(g + "lin_skip.weight")
                ws = [sd[g + "lin_key.weight"], sd[g + "lin_query.weight"], sd[g + "lin_value.weight"],
                      skip if skip is not None else torch.zeros(C, C, device=dev)]
                bs = [sd[g + "lin_key.bias"], sd[g + "lin_query.bias"], sd[g + "lin_value.bias"],
                      torch.zeros(C, device=dev)]
                b["kqvs_w"], b["kqvs_b"] = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()
                b["gnn_bias"] = sd.get(g + "bias")
                # split-operand weights of the three dense maps for the fp16 tensor-core path (upd_gemm3 wants 3K + 8 to be a
                # multiple of 8, i.e. K a multiple of 8 -- e.g. Td_h = 5 with 4 channels keeps the fp32 library GEMMs)
                if (c_out * T_in) % 8 == 0 and C % 8 == 0:
                    b["down_w3"] = _W3Cache().get([(b["down_w"].t().contiguous(), b["down_b"])])
                    b["kqvs_w3"] = _W3Cache().get([(b["kqvs_w"], b["kqvs_b"])])
                    b["up_w3"] = _W3Cache().get([(b["up_w"].t().contiguous(), b["up_b_full"])])
                else:
                    b["down_w3"] = b["kqvs_w3"] = b["up_w3"] = None
            else:
                b["w"], b["b"] = sd[pre + "conv.weight"][:, :, 0, :].contiguous(), sd[pre + "conv.bias"]
            self.blocks[pre] = b
        self.xproj_w, self.xproj_b = sd["x_proj.weight"][:, :, 0, :].contiguous(), sd["x_proj.bias"]
        self.out0_w, self.out0_b = sd["out.0.weight"][:, :, 0, :].contiguous(), sd["out.0.bias"]
        self.out1_w, self.out1_b = sd["out.1.weight"].contiguous(), sd["out.1.bias"].contiguous()
        # out.1 (Linear over the time axis) as a split-operand tensor-core GEMM when its width allows it
        self.out1_w3 = _W3Cache().get([(self.out1_w, self.out1_b)]) if self.out1_w.shape[1] % 8 == 0 and dev.type == "cuda" else None

    def _front(self, b, x, t, c_in, c_out, T_in, as_operand):
        """tcn1 (+ step embedding) -> tcn2 -> LayerNorm over channels: x [N, c_in, T] -> hn [N, c_out*T] fp32, or -- when
        ``as_operand`` -- directly the fp16 split operand [N, 3*c_out*T+8] of the down-sampling GEMM.  ``x`` may be a pair
        (x, skip): the up path's channel concatenation, which the kernel reads from the two tensors in place."""
        pair = x if isinstance(x, tuple) else None
        if pair is not None:
            x = pair[0]
        N = x.shape[0]
        if c_out in (4, 8, 16) and T_in % 2 == 0:
            K = c_out * T_in
            hn = None if as_operand else torch.empty((N, K), dtype=torch.float32, device=x.device)
            a3 = torch.empty((N, 3 * K + 8), dtype=torch.float16, device=x.device) if as_operand else None
            sc = torch.empty((N, K), dtype=torch.float32, device=x.device) if c_in != c_out else None
            tail = (_lib.ptr(b["tcn1.w"]), _lib.ptr(b["tcn1.b_step"][t]), _lib.ptr(b["tcn2.w"]), _lib.ptr(b["tcn2.b"]),
                    _lib.ptr(b["norm_w"]), _lib.ptr(b["norm_b"]))
            outs = (_lib.ptr(hn), _lib.ptr(a3), None if sc is None else _lib.ptr(b["sc_w2"]), _lib.ptr(sc),
                    _lib.stream_ptr(x.device))
            if pair is None:
                rc = _lib.lib().upd_stg_tcn_ln(_lib.ptr(x.contiguous()), *tail, N, c_in, c_out, T_in, *outs)
                _lib.check(rc, "upd_stg_tcn_ln")
            else:
                x1, x2 = pair[0].contiguous(), pair[1].contiguous()
                rc = _lib.lib().upd_stg_tcn_ln_cat(_lib.ptr(x1), x1.shape[1], _lib.ptr(x2), x2.shape[1], *tail, N, c_out, T_in,
                                                   *outs)
                _lib.check(rc, "upd_stg_tcn_ln_cat")
            return (a3 if as_operand else hn), sc
        # shapes outside the fused kernel's limits: the same arithmetic as library ops
        if pair is not None:
            x = torch.cat(pair, dim=1)
        h = F.conv1d(F.pad(x, (2, 0)), b["tcn1.w"], b["tcn1.b_step"][t])
        h = F.conv1d(F.pad(h, (2, 0)), b["tcn2.w"], b["tcn2.b"])
        var, mu = torch.var_mean(h, dim=1, unbiased=False, keepdim=True)
        hn = ((h - mu) * torch.rsqrt(var + 1e-5) * b["norm_w"][None, :, None] + b["norm_b"][None, :, None]).reshape(N, -1)
        sc = None if c_in == c_out else torch.matmul(b["sc_w2"], x).reshape(N, -1)
        return (a3_split(hn.contiguous()) if as_operand else hn), sc

    def _res(self, pre, x, t, c_in, c_out, T_in, rowptr, col, V):
        """One ResidualBlock.  The three dense maps (down-sampling, K|Q|V|skip, up-sampling) run as error-compensated
        fp16 tensor-core GEMMs (fx_encoder.gemm3: [x_hi | x_lo | x_hi | 1 1 0..] x [W_hi | W_hi | W_lo | b..]^T, 3e-6
        accuracy) when their widths allow it, else as fp32 library GEMMs."""
        b, Td = self.blocks[pre], self.Td_h
        N = x[0].shape[0] if isinstance(x, tuple) else x.shape[0]
        C = Td * c_out
        tc_ok = b["down_w3"] is not None
        front, sc = self._front(b, x, t, c_in, c_out, T_in, tc_ok)
        sc_is_temp = sc is not None                                                             # else a view of the block's input
        if sc is None:                                                                          # identity shortcut
            sc = (torch.cat(x, dim=1) if isinstance(x, tuple) else x).reshape(N, c_out * T_in)
        if tc_ok:
            sp = gemm3(front, b["down_w3"], C)                                                  # [N, Td*c]
            kqvs = gemm3(a3_split(sp.contiguous()), b["kqvs_w3"], 4 * C)                        # [N, 4C]
        else:
            sp = torch.addmm(b["down_b"], front, b["down_w"])
            kqvs = torch.addmm(b["kqvs_b"], sp, b["kqvs_w"].t())
        agg = gated_aggregate(kqvs.contiguous(), rowptr, col, b["gnn_bias"], V, C)
        # up-sampling GEMM with the shortcut as its accumulator input: out = shortcut + agg W_up + bias
        if tc_ok:
            up = gemm3(a3_split(agg), b["up_w3"], c_out * T_in, addend=sc.contiguous(), inplace=sc_is_temp)   # bias inside the GEMM
        else:
            up = torch.addmm(sc, agg, b["up_w"]) + b["up_b_full"]
        return up.view(N, c_out, T_in)

    def forward(self, xt, x_masked, t, rowptr, col, V):
        """xt, x_masked [N, T, F]; t: int diffusion step shared by all rows -> eps prediction [N, T, F]."""
        x = torch.cat((xt.transpose(1, 2), x_masked.transpose(1, 2)), dim=-1)                # [N, F, 2T]
        return self.trunk(x, t, rowptr, col, V).transpose(1, 2).contiguous()

    def trunk(self, x, t, rowptr, col, V, projected=False):
        """x [N, C_in, T_total] (or, with ``projected``, already x_proj(x) [N, d_h, T_total]) -> the out block's result
        [N, C_out, T_out]; t: index into the step-embedding table."""
        # the narrow convolutions (1x1 projections, DownSample, UpSample) run on upd_stg_conv1d; out.1 on upd_gemm3
        if not projected:
            x = conv1d_time(x, self.xproj_w, self.xproj_b, 1)
        hs = [x]

        def run(blk, x):
            pre, kind, c_in, c_out, T_in = blk
            if kind == "res":
                return self._res(pre, x, t, c_in, c_out, T_in, rowptr, col, V)
            b = self.blocks[pre]
            if kind == "downsample":
                return conv1d_time(x, b["w"], b["b"], 3, stride=2, pad=1)
            return conv1d_time(x, b["w"], b["b"], 4, stride=2, pad=1, transposed=True)

        for blk in self.down:
            x = run(blk, x)
            hs.append(x)
        for blk in self.middle:
            x = run(blk, x)
        for blk in self.up:
            x = run(blk, x if blk[1] == "upsample" else (x, hs.pop()))      # (x, skip): concatenated inside the kernel
        e = conv1d_time(x, self.out0_w, self.out0_b, 1)                      # [N, F, T_total]
        if self.out1_w3 is not None:
            N, nf, Tt = e.shape
            return gemm3(a3_split(e.view(N * nf, Tt)), self.out1_w3, self.out1_w.shape[0]).view(N, nf, -1)
        return F.linear(e, self.out1_w, self.out1_b)


class GraphedForward:
    """CUDA-graph replay of ``PreparedUGnet.forward``.  One eps prediction is ~750 kernel launches of a few microseconds
    each (12 residual blocks x (fused TCN, 3 GEMMs, 2 operand splits, gated aggregation, glue)): issued one by one from
    Python the GPU waits for the host.  A forward depends on the step only through a row of the per-step bias tables, so
    it is captured once per (step t, row count N) into a graph over static input buffers and replayed for every group of
    replicas and every sweep after that; all graphs of a model share one memory pool (they never run concurrently)."""

    def __init__(self, prep):
        self.prep, self.graphs, self.static, self.pool = prep, {}, {}, None

    def __call__(self, xt, xm, t, rowptr, col, V):
        skey = (tuple(xt.shape), rowptr.data_ptr(), col.data_ptr(), int(V))
        st = self.static.get(skey)
        if st is None:
            st = self.static[skey] = {"xt": torch.empty_like(xt), "xm": torch.empty_like(xm)}
            self.prep.forward(xt, xm, t, rowptr, col, V)            # eager once per shape: lazy caches, cuDNN plans
        st["xt"].copy_(xt)
        st["xm"].copy_(xm)
        g = self.graphs.get((skey, int(t)))
        if g is None:
            torch.cuda.current_stream(xt.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self.pool):
                out = self.prep.forward(st["xt"], st["xm"], t, rowptr, col, V)
            if self.pool is None:
                self.pool = graph.pool()
            g = self.graphs[(skey, int(t))] = (graph, out)
        g[0].replay()
        return g[1]


class DiffSTG(nn.Module):
    def __init__(self, net_param):
        super().__init__()
        self.diffusion_steps = net_param["diffusion_steps"]
        self.inference_diffusion_steps = net_param["inference_diffusion_steps"]
        self.inference_trick = net_param["inference_trick"] if net_param.get("inference_trick") else "ddim"
        self.device = net_param["device"]
        self.diffusion_schedule = net_param["diffusion_schedule"]
        self.inference_schedule = net_param["inference_schedule"]
        self.loss_weight_schedule = net_param["loss_weight_schedule"]
        self.parallel_sampling = net_param["parallel_sampling"]
        self.sequential_sampling = net_param["sequential_sampling"]
        self.sparse = True
        self.diffusion = GaussianDiffusion(T=self.diffusion_steps, schedule=self.diffusion_schedule,
                                           loss_weight_schedule=self.loss_weight_schedule)
        self.T_p, self.T_h = net_param["T_p"], net_param["T_h"]
        self.T = self.T_p + self.T_h
        self.F = net_param["F"]
        self.mask_ratio = net_param["mask_ratio"]
        self.net_param = net_param
        self.model = ParamTree()
        populate_ugnet(self.model, ugnet_shapes(net_param), self._alias)
        self.scaler = net_param["scaler_type"]
        self.register_buffer("scaler_mean", torch.zeros(self.F))
        self.register_buffer("scaler_std", torch.zeros(self.F))
        self.rows_per_launch = ROWS_PER_LAUNCH
        self._prepared = None
        self._prepared_key = None
        self._csr_cache = None
        self._windows_drawn = 0
        self.to(self.device)

    def _alias(self, key, param):
        alias_parameter(self.model, key, param)

    def scaler_fit(self, data):
        data_std = data.std(axis=(0, 1))
        data_std[data_std == 0] = 1
        self.scaler_mean = data.mean(axis=(0, 1))
        self.scaler_std = data_std

    def scaler_transform(self, data):
        return (data - self.scaler_mean) / self.scaler_std

    def scaler_inverse_transform(self, data):
        return (data * self.scaler_std) + self.scaler_mean

    def forward(self, gdatalist):
        raise NotImplementedError("training is outside the accelerated hot path (SURVEY section 8: out of scope)")

    training_step = forward

    def _apply(self, fn, *a, **k):
        self._prepared = None
        return super()._apply(fn, *a, **k)

    def prepared(self):
        params = list(self.model.parameters())
        key = (params[0].device,) + tuple(p._version for p in params)
        if self._prepared is None or self._prepared_key != key:
            _lib.require_cuda(params[0].device)
            with torch.no_grad():
                self._prepared = PreparedUGnet(self.model, self.net_param, self.diffusion.T)
            self._prepared_key = key
        return self._prepared

    def posterior_coefficients(self, t, target_t):
        """gaussian_posterior's scalars (graph_diffusion_model.py:46-73), formed in float64 like the reference
        -> (a, b, c, uses_noise): x_target = a*(xt - b*pred) + c*(z if uses_noise else pred)."""
        d = self.diffusion
        if target_t is None:
            target_t = t - 1
        atbar, atbar_target = d.alphabar[t], d.alphabar[target_t]
        if self.inference_trick == "ddpm" or t <= 1:
            at = d.alpha[t]
            atbar_prev = d.alphabar[t - 1]
            beta_tilde = d.beta[t - 1] * (1 - atbar_prev) / (1 - atbar)
            return (1 / np.sqrt(at)).item(), ((1 - at) / np.sqrt(1 - atbar)).item(), np.sqrt(beta_tilde).item(), True
        if self.inference_trick == "ddim":
            return (np.sqrt(atbar_target / atbar).item(), np.sqrt(1 - atbar).item(), np.sqrt(1 - atbar_target).item(),
                    False)
        raise ValueError("Unknown inference trick {}".format(self.inference_trick))

    def step_plan(self):
        """[(t1, t2, a, b, c, uses_noise)] of one sampling round (graph_diffusion_model.py:255-267)."""
        sched = InferenceSchedule(inference_schedule=self.inference_schedule, T=self.diffusion.T,
                                  inference_T=self.inference_diffusion_steps)
        plan = []
        for i in range(self.inference_diffusion_steps):
            t1, t2 = sched(i)
            plan.append((int(t1), int(t2)) + tuple(self.posterior_coefficients(int(t1), int(t2))))
        return plan

    def draws_per_round(self):
        return 1 + sum(1 for p in self.step_plan() if p[5])

    use_cuda_graphs = os.environ.get("UPD_STG_GRAPHS", "1") != "0"

    def _graphed_forward(self, prep):
        if getattr(self, "_graphed", None) is None or self._graphed.prep is not prep:
            self._graphed = GraphedForward(prep)
        return self._graphed

    def predict_eps(self, xt, x_masked, t, edge_index, num_nodes):
        """UGnet.forward for rows that share step t; rows are replicas of the num_nodes-node graph."""
        dev = _lib.require_cuda(xt.device)
        rowptr, col = self._csr(edge_index, num_nodes, dev)
        with torch.no_grad(), torch.cuda.device(dev):
            return self.prepared().forward(xt.contiguous(), x_masked.contiguous(), int(t), rowptr, col, num_nodes)

    def _csr(self, edge_index, num_nodes, dev):
        key = (id(edge_index), tuple(edge_index.shape), num_nodes, str(dev))
        if self._csr_cache is None or self._csr_cache[0] != key:
            rowptr, col = graph_csr(edge_index, num_nodes)
            if col.numel() == 0:
                col = torch.zeros(1, dtype=torch.int32)
            self._csr_cache = (key, rowptr.to(dev), col.to(dev), edge_index)
        return self._csr_cache[1], self._csr_cache[2]

    def sample_windows(self, windows, edge_index, num_nodes, noise=None, seed=None, window_base=None):
        """windows [W, Node, T_h(+T_p), F] scaled -> trajectories [W*Node, K, T, F] on the device, K =
        sequential_sampling * parallel_sampling with sample index sq*P + p (graph_diffusion_model.py:269-280).
        noise: validation mode, noise[w][sq] = list of that round's draws ([P*Node, T, F] each) in reference order."""
        dev = _lib.require_cuda(self.scaler_mean.device)
        W, V = windows.shape[0], windows.shape[1]
        if V != num_nodes:
            raise ValueError("windows hold {} nodes, graph has {}".format(V, num_nodes))
        P_, Sq = int(self.parallel_sampling), int(self.sequential_sampling)
        K = P_ * Sq
        T, nf = self.T, self.F
        hist = windows[:, :, :self.T_h, :].to(dev, torch.float32)
        x_masked_w = torch.cat([hist, torch.zeros(W, V, self.T_p, nf, device=dev)], dim=2)          # [W, V, T, F]
        if seed is None:
            seed = torch.initial_seed()
        if window_base is None:
            window_base = self._windows_drawn
            self._windows_drawn += W
        rowptr, col = self._csr(edge_index, num_nodes, dev)
        prep, lib, plan = self.prepared(), _lib.lib(), self.step_plan()
        fwd = self._graphed_forward(prep) if self.use_cuda_graphs and noise is None and dev.type == "cuda" else prep.forward
        reps = [(w, k) for w in range(W) for k in range(K)]             # replica = (window, sample)
        per = max(1, self.rows_per_launch // V)
        if noise is not None:
            per = max(P_, per // P_ * P_)                               # keep whole rounds together
        out = torch.empty((W, K, V, T, nf), dtype=torch.float32, device=dev)
        with torch.no_grad(), torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            for r0 in range(0, len(reps), per):
                group = reps[r0:r0 + per]
                N = len(group) * V
                xm = torch.stack([x_masked_w[w] for (w, k) in group], 0).reshape(N, T, nf).contiguous()

                def draw(i):
                    if noise is not None:
                        parts = [noise[w][k // P_][i][(k % P_) * V:(k % P_ + 1) * V] for (w, k) in group]
                        return torch.cat(parts, 0).to(dev, torch.float32).contiguous()
                    # replicas of a group are consecutive (window, sample) pairs, so their Philox row keys
                    # ((window_base + w)*K + k)*V + v are one contiguous range: one launch fills the whole group
                    z = torch.empty((N, T, nf), dtype=torch.float32, device=dev)
                    w_first, k_first = group[0]
                    base = ((window_base + w_first) * K + k_first) * V
                    _lib.check(lib.upd_gauss_fill(_lib.ptr(z), N, T * nf, seed & (2 ** 64 - 1), base, i, st), "upd_gauss_fill")
                    return z

                i_draw = 0
                xt = draw(i_draw)
                i_draw += 1
                nxt = torch.empty_like(xt)
                for (t1, t2, a, b, c, noisy) in plan:
                    pred = fwd(xt, xm, t1, rowptr, col, V)
                    z = None
                    if noisy:
                        z = draw(i_draw)
                        i_draw += 1
                    _lib.check(lib.upd_stg_posterior(_lib.ptr(xt), _lib.ptr(pred), _lib.ptr(z), N * T * nf, a, b, c,
                                                     _lib.ptr(nxt), st), "upd_stg_posterior")
                    xt, nxt = nxt, xt
                res = xt.view(len(group), V, T, nf)
                for gi, (w, k) in enumerate(group):
                    out[w, k] = res[gi]
        return out.permute(0, 2, 1, 3, 4).reshape(W * V, K, T, nf).contiguous()

    def evaluation_step(self, data, noise=None):
        """graph_diffusion_model.py:204-282 for a single graph -> (splitted_predict_x0 [Node, T, 1, K] cpu, x0_truth)."""
        if hasattr(data, "ptr") and getattr(data, "ptr") is not None and len(data.ptr) > 2:
            raise NotImplementedError("batched graphs (torch_geometric Batch of several graphs) are not on the inference path")
        x = data.x
        history = x[:, :self.T_h, :].to(self.device)
        if x.shape[1] - self.T_h >= self.T_p:
            future = x[:, self.T_h:, :].to(self.device)
            assert future.size(1) == self.T_p, "pred_len is not equal to the length of the prediction"
            x0_truth = torch.cat([history, future], dim=1)
        else:
            x0_truth = None
        traj = self.sample_windows(history.unsqueeze(0), data.edge_index, int(data.num_nodes),
                                   noise=None if noise is None else [noise])
        return traj.cpu().permute(0, 2, 3, 1), x0_truth
