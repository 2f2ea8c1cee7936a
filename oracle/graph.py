"""ORACLE (test infrastructure only) — torch_cluster's radius / radius_graph / knn
restated on the CPU, on top of ``oracle/csrc/graph_oracle.c``.

Reference call sites: ``radius_graph`` models/mpnn_2d.py:245, models/mpnn.py:245,
models/magnet_gnn.py:293; ``knn`` models/magnet_gnn.py:247.  Semantics: SURVEY.md §8c
(CUDA ordering of torch_cluster is the canonical one).  PARITY UNPINNED.
"""
import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import build, LIB_PATH

_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB_PATH)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.oracle_radius.argtypes = [f32p, f32p, ctypes.c_int, i64p, i64p, ctypes.c_int64,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int, i64p, i32p, ctypes.c_int]
        lib.oracle_radius.restype = None
        lib.oracle_knn.argtypes = [f32p, f32p, ctypes.c_int, i64p, i64p, ctypes.c_int64,
                                   ctypes.c_int, ctypes.c_int, i64p, f32p, ctypes.c_int]
        lib.oracle_knn.restype = None
        lib.oracle_radius_margin_ulps.argtypes = [f32p, ctypes.c_int, i64p, ctypes.c_int64, ctypes.c_double]
        lib.oracle_radius_margin_ulps.restype = ctypes.c_double
        _lib = lib
    return _lib


def _f32(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu()
    if t.dim() == 1:
        t = t[:, None]
    return np.ascontiguousarray(t.to(torch.float32).numpy())


def _ptr_from_batch(batch: Optional[torch.Tensor], n: int) -> np.ndarray:
    # torch_cluster: batch_size = int(batch.max()) + 1; deg = bincount; ptr = cumsum
    if batch is None:
        return np.array([0, n], dtype=np.int64)
    b = batch.detach().cpu().to(torch.int64).numpy()
    assert np.all(np.diff(b) >= 0), "batch vector must be sorted"
    nb = int(b.max()) + 1 if b.size else 1
    deg = np.bincount(b, minlength=nb)
    return np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors: int = 32,
           fma: bool = True, threads: int = 0) -> torch.Tensor:
    """torch_cluster.radius: returns [2, E] = [row (query index into y); col (index into x)]."""
    lib = _load()
    xa, ya = _f32(x), _f32(y)
    d = xa.shape[1]
    px, py = _ptr_from_batch(batch_x, xa.shape[0]), _ptr_from_batch(batch_y, ya.shape[0])
    nb = min(len(px), len(py)) - 1
    ny = ya.shape[0]
    col = np.empty((ny, max_num_neighbors), dtype=np.int64)
    cnt = np.empty((ny,), dtype=np.int32)
    lib.oracle_radius(_p(xa, ctypes.c_float), _p(ya, ctypes.c_float), d, _p(px, ctypes.c_int64),
                      _p(py, ctypes.c_int64), nb, float(r), int(max_num_neighbors), int(fma),
                      _p(col, ctypes.c_int64), _p(cnt, ctypes.c_int32), threads)
    mask = col >= 0
    row = np.broadcast_to(np.arange(ny, dtype=np.int64)[:, None], col.shape)
    return torch.from_numpy(np.stack([row[mask], col[mask]]))


def radius_graph(x, r, batch=None, loop: bool = False, max_num_neighbors: int = 32,
                 flow: str = "source_to_target", fma: bool = True, threads: int = 0) -> torch.Tensor:
    """torch_cluster.radius_graph (re-exported by torch_geometric.nn)."""
    assert flow in ("source_to_target", "target_to_source")
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1,
                fma=fma, threads=threads)
    if flow == "source_to_target":
        row, col = ei[1], ei[0]
    else:
        row, col = ei[0], ei[1]
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col], dim=0)


def knn(x, y, k: int, batch_x=None, batch_y=None, fma: bool = True, threads: int = 0,
        return_dist: bool = False):
    """torch_cluster.knn: returns [2, M*k] = [query index into y; index into x], ascending distance."""
    lib = _load()
    assert k <= 100
    xa, ya = _f32(x), _f32(y)
    d = xa.shape[1]
    px, py = _ptr_from_batch(batch_x, xa.shape[0]), _ptr_from_batch(batch_y, ya.shape[0])
    nb = min(len(px), len(py)) - 1
    ny = ya.shape[0]
    idx = np.empty((ny, k), dtype=np.int64)
    dist = np.empty((ny, k), dtype=np.float32)
    lib.oracle_knn(_p(xa, ctypes.c_float), _p(ya, ctypes.c_float), d, _p(px, ctypes.c_int64),
                   _p(py, ctypes.c_int64), nb, int(k), int(fma), _p(idx, ctypes.c_int64),
                   _p(dist, ctypes.c_float), threads)
    row = np.broadcast_to(np.arange(ny, dtype=np.int64)[:, None], idx.shape)
    mask = idx >= 0
    out = torch.from_numpy(np.stack([row[mask], idx[mask]]))
    if return_dist:
        return out, torch.from_numpy(dist[mask])
    return out


def radius_margin_ulps(x, r, batch=None) -> float:
    """Smallest | |xi-xj|^2 - r^2 | in ulps of r^2 (generators reject meshes with <= 4)."""
    lib = _load()
    xa = _f32(x)
    ptr = _ptr_from_batch(batch, xa.shape[0])
    return float(lib.oracle_radius_margin_ulps(_p(xa, ctypes.c_float), xa.shape[1],
                                               _p(ptr, ctypes.c_int64), len(ptr) - 1, float(r)))
