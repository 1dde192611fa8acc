"""Stand-in for torch_geometric.nn.ResGatedGraphConv (torch-geometric==2.5.3, requirements.txt:5), the one
third-party layer on the DiffSTG path (models/layer/gnn_conv.py:18-19).  TEST INFRASTRUCTURE ONLY.

torch_geometric is not installed here and not vendored by the reference, so this restates the layer's published
definition (Bresson & Laurent, "Residual Gated Graph ConvNets"; PyG docs):

    x_i' = W_skip x_i + sum_{j in N(i)} sigmoid(W_key x_i + W_query x_j) * W_value x_j + b

with PyG's parameter names (lin_key / lin_query / lin_value have biases, lin_skip has none, `bias` is a separate
vector), flow source_to_target (edge_index[0] = j, edge_index[1] = i) and sum aggregation in edge order.
Nothing in the reference pins this layer and no DiffSTG checkpoint ships, so everything that flows through it is
"parity unpinned" (DESIGN.md); the rest of the DiffSTG fixtures is the reference's own torch code."""
import torch
import torch.nn as nn


class ResGatedGraphConv(nn.Module):
    def __init__(self, in_channels, out_channels, act=None, edge_dim=None, root_weight=True, bias=True, **kwargs):
        super().__init__()
        assert edge_dim is None
        self.in_channels, self.out_channels, self.root_weight = in_channels, out_channels, root_weight
        self.lin_key = nn.Linear(in_channels, out_channels)
        self.lin_query = nn.Linear(in_channels, out_channels)
        self.lin_value = nn.Linear(in_channels, out_channels)
        if root_weight:
            self.lin_skip = nn.Linear(in_channels, out_channels, bias=False)
        else:
            self.register_parameter("lin_skip", None)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index):
        k, q, v = self.lin_key(x), self.lin_query(x), self.lin_value(x)
        src, dst = edge_index[0], edge_index[1]
        msg = torch.sigmoid(k[dst] + q[src]) * v[src]
        out = torch.zeros_like(k).index_add_(0, dst, msg)
        if self.root_weight:
            out = out + self.lin_skip(x)
        if self.bias is not None:
            out = out + self.bias
        return out
