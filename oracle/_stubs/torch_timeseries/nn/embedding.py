class DataEmbedding:  # import-only stub
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_timeseries stub has no arithmetic")
