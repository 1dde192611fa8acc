class _S:
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_timeseries stub has no arithmetic")
class DSAttention(_S): pass
class AttentionLayer(_S): pass
