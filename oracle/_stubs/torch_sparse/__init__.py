class SparseTensor:  # import-only stub
    pass
def sum(*a, **k): raise NotImplementedError
def mean(*a, **k): raise NotImplementedError
def max(*a, **k): raise NotImplementedError
