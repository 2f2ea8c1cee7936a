"""Restated subset of torch_geometric.nn (2.0.3) + torch_scatter / torch_cluster pieces it
re-exports.  ORACLE ONLY — test infrastructure; never imported by the product.

Reference call sites:
  MessagePassing   models/mpnn_2d.py:27,46,69   models/magnet_gnn.py:44,54,76,92,103
  InstanceNorm     models/mpnn_2d.py:63,70
  radius_graph     models/mpnn_2d.py:245  models/mpnn.py:245  models/magnet_gnn.py:293
  knn              models/magnet_gnn.py:247
Semantics restated from SURVEY.md §8c (PyG 2.0.3 `propagate`/`__collect__`,
torch_scatter 2.0.x `scatter`, PyG `InstanceNorm`).  PARITY UNPINNED.
"""
import inspect

import torch
from torch import nn

from oracle.graph import radius_graph, knn, radius  # noqa: F401  (torch_cluster restatement)


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = 0, dim_size=None, reduce: str = "add"):
    """torch_scatter.scatter for dim = node dim of a 2-D tensor (all the path needs)."""
    if dim < 0:
        dim += src.dim()
    assert dim == 0 and src.dim() == 2
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    if reduce in ("add", "sum", "mean"):
        out = torch.zeros(n, src.shape[1], dtype=src.dtype, device=src.device)
        out = out.index_add(0, index, src)
        if reduce == "mean":
            # torch_scatter: count = scatter_sum(ones); count.clamp_(min=1); out.true_divide_(count)
            count = torch.zeros(n, dtype=src.dtype, device=src.device)
            count = count.index_add(0, index, torch.ones_like(index, dtype=src.dtype))
            out = out / count.clamp(min=1).unsqueeze(-1)
        return out
    if reduce == "max":
        out = torch.full((n, src.shape[1]), float("-inf"), dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(0, index[:, None].expand_as(src), src, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    raise ValueError(reduce)


class MessagePassing(nn.Module):
    """`propagate` of PyG 2.0.3, flow source_to_target: `*_j` = index_select(node_dim, edge_index[0]),
    `*_i` = index_select(node_dim, edge_index[1]); un-suffixed kwargs pass through untouched to
    `message` and `update` alike (this is what makes InteractionNetwork return its INPUT edge
    features); aggregation index is edge_index[1]."""

    special_args = {"edge_index", "adj_t", "edge_index_i", "edge_index_j", "size", "size_i", "size_j",
                    "ptr", "index", "dim_size"}

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert flow == "source_to_target"

    def _params(self, fn, pop_first):
        ps = list(inspect.signature(fn).parameters.items())
        return ps[1:] if pop_first else ps

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = 1, 0
        coll = {"edge_index": edge_index, "edge_index_i": edge_index[i], "edge_index_j": edge_index[j],
                "index": edge_index[i], "ptr": None, "size": size}
        user_args = [n for fn, pf in ((self.message, False), (self.update, True))
                     for n, _ in self._params(fn, pf) if n not in self.special_args]
        n_nodes = None
        for arg in user_args:
            if arg[-2:] not in ("_i", "_j"):
                coll[arg] = kwargs.get(arg, inspect.Parameter.empty)
            else:
                data = kwargs.get(arg[:-2], inspect.Parameter.empty)
                if torch.is_tensor(data):
                    n_nodes = data.size(self.node_dim)
                    data = data.index_select(self.node_dim, edge_index[j if arg[-2:] == "_j" else i])
                coll[arg] = data
        coll["dim_size"] = n_nodes
        coll["size_i"] = coll["size_j"] = n_nodes

        def distribute(fn, pop_first):
            out = {}
            for name, p in self._params(fn, pop_first):
                v = coll.get(name, inspect.Parameter.empty)
                if v is inspect.Parameter.empty:
                    if p.default is inspect.Parameter.empty:
                        raise TypeError(f"Required parameter {name} is empty.")
                    v = p.default
                out[name] = v
            return out

        msg = self.message(**distribute(self.message, False))
        out = scatter(msg, coll["index"], dim=self.node_dim, dim_size=n_nodes, reduce=self.aggr)
        return self.update(out, **distribute(self.update, True))

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


class InstanceNorm(nn.Module):
    """PyG InstanceNorm(C, eps=1e-5, affine=False, track_running_stats=False): per graph and
    channel over that graph's nodes, biased variance of the centred values; batch statistics in
    train AND eval.  No parameters, no buffers (state_dict stays empty)."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=False, track_running_stats=False):
        super().__init__()
        assert not affine and not track_running_stats
        self.in_channels = in_channels
        self.eps = eps

    def forward(self, x, batch=None):
        if batch is None:
            batch = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
        batch_size = int(batch.max()) + 1
        norm = torch.zeros(batch_size, dtype=x.dtype).index_add(0, batch, torch.ones_like(batch, dtype=x.dtype))
        norm = norm.clamp_(min=1).view(-1, 1)
        mean = scatter(x, batch, dim=0, dim_size=batch_size, reduce="add") / norm
        x = x - mean.index_select(0, batch)
        var = scatter(x * x, batch, dim=0, dim_size=batch_size, reduce="add") / norm
        return x / (var + self.eps).sqrt().index_select(0, batch)
