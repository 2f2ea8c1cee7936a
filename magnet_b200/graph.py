"""Graph construction and aggregation plans on the GPU (host side of csrc/graph.cu).

Public functions keep the names, argument meaning and result layout of the third-party
functions the reference calls:

* ``radius_graph(x, r, batch, loop, max_num_neighbors, flow)`` — torch_geometric.nn.radius_graph
  as used at models/mpnn_2d.py:245, models/mpnn.py:245, models/magnet_gnn.py:293
* ``knn(x, y, k, batch_x, batch_y)`` — torch_geometric.nn.knn as used at models/magnet_gnn.py:247

plus ``plan_for(edge_index, n_nodes)``: the private int32 CSR plans (sorted by the aggregation
endpoint ``edge_index[1]`` and its transpose) that the fused layer kernels consume.  Plans are
cached on the identity/version of the ``edge_index`` tensor, so the reference's habit of rebuilding
an identical graph every rollout step costs nothing after the first step.
"""
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


def ptr_from_batch(batch: Optional[torch.Tensor], n: int, device) -> torch.Tensor:
    """int64 [B+1] sample offsets from a SORTED batch vector (torch_geometric's convention, models/mpnn_2d.py:237,249): one
    binary-search launch; the sample count is the last entry + 1 (one 8-byte read instead of a max-reduction + histogram)."""
    if batch is None:
        return torch.tensor([0, n], dtype=torch.int64, device=device)
    if batch.numel() == 0:
        return torch.zeros(2, dtype=torch.int64, device=device)
    nb = int(batch[-1].item()) + 1
    return torch.searchsorted(batch.contiguous(), torch.arange(nb + 1, device=batch.device, dtype=batch.dtype)).to(torch.int64)


def uniform_ptr(n_samples: int, n_per_sample: int, device) -> torch.Tensor:
    """Offsets for B samples of N nodes each — no device work, no sync (the reference builds the
    same information from Python lists of length B*N, models/magnet_gnn.py:292)."""
    return torch.arange(0, (n_samples + 1) * n_per_sample, n_per_sample, dtype=torch.int64, device=device)


def radius_graph(x: torch.Tensor, r, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target", *,
                 ptr: Optional[torch.Tensor] = None, swap_rows: bool = False) -> torch.Tensor:
    """edge_index int64 [2, E], bit-identical to torch_cluster's CUDA result: for every centre the
    first `max_num_neighbors` in-radius nodes in node-index order, rows [neighbour; centre] for
    flow='source_to_target', sorted by centre then neighbour.  ``swap_rows`` returns
    [centre; neighbour] directly (what MAgNetGNN._build_graph builds, models/magnet_gnn.py:294-296).
    """
    assert flow in ("source_to_target", "target_to_source")
    _lib.require_cuda(x)
    L = _lib.lib()
    if x.dim() == 1:
        x = x[:, None]
    pos = _lib.f32c(x)
    n, d = pos.shape
    dev = pos.device
    if ptr is None:
        ptr = ptr_from_batch(batch, n, dev)
    ptr = ptr.to(device=dev, dtype=torch.int64).contiguous()
    n_samples = ptr.numel() - 1
    r = float(r)
    cap = max_num_neighbors + (0 if loop else 1)
    nbr = torch.empty((n, cap), dtype=torch.int32, device=dev)
    deg = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    rowptr = torch.empty((n + 1,), dtype=torch.int32, device=dev)
    ws = _lib.workspace(L.mgb_radius_graph_workspace(n, n_samples), dev)
    _lib.check(L.mgb_radius_graph_search(_lib.ptr(pos), n, d, _lib.ptr(ptr), n_samples, r, max_num_neighbors,
                                         int(loop), _lib.ptr(nbr), _lib.ptr(deg), _lib.ptr(rowptr), _lib.ptr(ws),
                                         ws.numel(), _lib.stream()), "radius_graph_search")
    n_edges = int(rowptr[n].item())           # the one host sync: the result shape depends on it
    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
    centre_row = 1 if flow == "source_to_target" else 0
    if swap_rows:
        centre_row = 1 - centre_row
    _lib.check(L.mgb_radius_graph_emit(_lib.ptr(nbr), _lib.ptr(rowptr), n, cap, centre_row, n_edges,
                                       _lib.ptr(edge_index), None, _lib.stream()), "radius_graph_emit")
    return edge_index


def knn_indices(x: torch.Tensor, y: torch.Tensor, k: int, ptr_x: torch.Tensor, ptr_y: torch.Tensor,
                return_dist: bool = False):
    """[ny, k] int64 nearest-x indices per query (ascending distance, ties -> lower index, -1 padded)."""
    _lib.require_cuda(x, y)
    L = _lib.lib()
    if x.dim() == 1:
        x = x[:, None]
    if y.dim() == 1:
        y = y[:, None]
    xs, ys = _lib.f32c(x), _lib.f32c(y)
    nx, d = xs.shape
    ny = ys.shape[0]
    dev = xs.device
    ptr_x = ptr_x.to(device=dev, dtype=torch.int64).contiguous()
    ptr_y = ptr_y.to(device=dev, dtype=torch.int64).contiguous()
    n_samples = ptr_x.numel() - 1
    out = torch.empty((ny, k), dtype=torch.int64, device=dev)
    dist = torch.empty((ny, k), dtype=torch.float32, device=dev) if return_dist else None
    ws = _lib.workspace(L.mgb_knn_workspace(nx, n_samples), dev)
    _lib.check(L.mgb_knn(_lib.ptr(xs), nx, _lib.ptr(ys), ny, d, _lib.ptr(ptr_x), _lib.ptr(ptr_y), n_samples, k,
                         _lib.ptr(out), _lib.ptr(dist), _lib.ptr(ws), ws.numel(), _lib.stream()), "knn")
    return (out, dist) if return_dist else out


def knn(x: torch.Tensor, y: torch.Tensor, k: int, batch_x: Optional[torch.Tensor] = None,
        batch_y: Optional[torch.Tensor] = None, *, ptr_x=None, ptr_y=None) -> torch.Tensor:
    """assign_index int64 [2, M*k] = [query index; x index] as torch_cluster.knn returns it."""
    nx = x.shape[0]
    ny = y.shape[0]
    if ptr_x is None:
        ptr_x = ptr_from_batch(batch_x, nx, x.device)
    if ptr_y is None:
        ptr_y = ptr_from_batch(batch_y, ny, y.device)
    idx = knn_indices(x, y, k, ptr_x, ptr_y)
    row = torch.arange(ny, device=idx.device, dtype=torch.int64)[:, None].expand(ny, k)
    full = bool((torch.diff(ptr_x.to(idx.device)) >= k).all().item()) if ptr_x.numel() > 1 else True
    if full:
        return torch.stack([row.reshape(-1), idx.reshape(-1)])
    mask = idx >= 0
    return torch.stack([row[mask], idx[mask]])


# --------------------------------------------------------------------------------------------
# aggregation plans
# --------------------------------------------------------------------------------------------
@dataclass
class AggregationPlan:
    n_nodes: int
    n_edges: int
    rowptr: torch.Tensor     # int32 [n_nodes+1]   segments of the aggregation endpoint edge_index[1]
    perm: torch.Tensor       # int32 [E]           COO edge id at each aggregation-order position
    dst: torch.Tensor        # int32 [E]           edge_index[1] in aggregation order
    src: torch.Tensor        # int32 [E]           edge_index[0] in aggregation order
    rowptr_t: torch.Tensor   # int32 [n_nodes+1]   segments of edge_index[0]
    pos_t: torch.Tensor      # int32 [E]           aggregation-order positions grouped by edge_index[0]
    _perm_src: Optional[torch.Tensor] = None

    def perm_src(self) -> torch.Tensor:
        """int32 [E]: COO edge ids grouped by edge_index[0] (segments rowptr_t) = perm[pos_t]."""
        if self._perm_src is None:
            self._perm_src = self.perm[self.pos_t.long()].contiguous()
        return self._perm_src


_PLAN_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_PLAN_CACHE_SIZE = 16


def build_plan(edge_index: torch.Tensor, n_nodes: int) -> AggregationPlan:
    _lib.require_cuda(edge_index)
    L = _lib.lib()
    assert edge_index.dim() == 2 and edge_index.shape[0] == 2 and edge_index.dtype == torch.int64
    ei = edge_index.contiguous()
    E = ei.shape[1]
    dev = ei.device
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr = torch.empty(n_nodes + 1, **i32)
    rowptr_t = torch.empty(n_nodes + 1, **i32)
    perm = torch.empty(max(E, 1), **i32)
    dst = torch.empty(max(E, 1), **i32)
    src = torch.empty(max(E, 1), **i32)
    pos_t = torch.empty(max(E, 1), **i32)
    bad = torch.zeros(1, **i32)
    ws = _lib.workspace(L.mgb_csr_plan_workspace(E), dev)
    _lib.check(L.mgb_csr_plan(_lib.ptr(ei[1]), _lib.ptr(ei[0]), E, n_nodes, _lib.ptr(rowptr), _lib.ptr(perm),
                              _lib.ptr(dst), _lib.ptr(src), _lib.ptr(rowptr_t), _lib.ptr(pos_t), _lib.ptr(bad),
                              _lib.ptr(ws), ws.numel(), _lib.stream()), "csr_plan")
    if int(bad.item()) != 0:
        raise RuntimeError(f"edge_index holds node ids outside [0, {n_nodes})")
    return AggregationPlan(n_nodes, E, rowptr, perm, dst, src, rowptr_t, pos_t)


def plan_for(edge_index: torch.Tensor, n_nodes: int) -> AggregationPlan:
    key = (edge_index.data_ptr(), _lib.ver(edge_index), tuple(edge_index.shape), int(n_nodes), edge_index.device.index)
    hit = _PLAN_CACHE.get(key)
    if hit is not None and hit[0]() is edge_index:
        _PLAN_CACHE.move_to_end(key)
        return hit[1]
    import weakref
    plan = build_plan(edge_index, n_nodes)
    _PLAN_CACHE[key] = (weakref.ref(edge_index), plan)
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)
    return plan


@dataclass
class GraphSegments:
    """`batch` vector as offsets: gptr int64 [G+1] on the device, plus the host-side sizes."""
    gptr: torch.Tensor
    n_graphs: int
    max_nodes: int


_SEG_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def segments_for(batch: torch.Tensor, n_nodes: int) -> GraphSegments:
    """Cached conversion of a sorted `batch` vector (models/mpnn_2d.py:237,249) to offsets."""
    import weakref
    key = (batch.data_ptr(), _lib.ver(batch), batch.numel(), batch.device.index)
    hit = _SEG_CACHE.get(key)
    if hit is not None and hit[0]() is batch:
        _SEG_CACHE.move_to_end(key)
        return hit[1]
    gptr = ptr_from_batch(batch, n_nodes, batch.device)
    sizes = torch.diff(gptr)
    seg = GraphSegments(gptr, gptr.numel() - 1, int(sizes.max().item()) if sizes.numel() else 0)
    _SEG_CACHE[key] = (weakref.ref(batch), seg)
    while len(_SEG_CACHE) > _PLAN_CACHE_SIZE:
        _SEG_CACHE.popitem(last=False)
    return seg


def uniform_segments(n_graphs: int, n_per_graph: int, device) -> GraphSegments:
    return GraphSegments(uniform_ptr(n_graphs, n_per_graph, device), n_graphs, n_per_graph)
