def k_hop_subgraph(*a, **k):
    raise NotImplementedError("stub")
def from_networkx(*a, **k):
    raise NotImplementedError("stub")
