class Data:  # attribute bag only (no arithmetic): what DiffSTG.evaluation_step reads is x, edge_index, num_nodes
    def __init__(self, x=None, edge_index=None, num_nodes=None, **kw):
        self.x, self.edge_index, self.num_nodes = x, edge_index, num_nodes
        for k, v in kw.items():
            setattr(self, k, v)
    def clone(self):
        import copy
        return copy.deepcopy(self)
class Batch(Data):
    pass
class HeteroData(Data):
    pass
