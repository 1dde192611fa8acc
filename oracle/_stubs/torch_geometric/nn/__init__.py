import torch.nn as _nn
class MessagePassing(_nn.Module):  # import-only stub
    def __init__(self, *a, **k):
        super().__init__()
def __getattr__(name):
    raise AttributeError("torch_geometric.nn stub has no arithmetic: " + name)
