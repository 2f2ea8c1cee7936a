"""torch_geometric.data.Data — attribute bag (reference use: models/mpnn_2d.py:247-249)."""


class Data:
    def __init__(self, x=None, edge_index=None, pos=None, batch=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.pos = pos
        self.batch = batch
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if hasattr(v, "to"):
                setattr(self, k, v.to(device))
        return self
