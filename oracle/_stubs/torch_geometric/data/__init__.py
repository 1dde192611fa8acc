class Data:  # import-only stub
    pass
class Batch(Data):
    pass
class HeteroData(Data):
    pass
