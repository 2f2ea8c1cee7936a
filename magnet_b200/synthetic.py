"""Seeded synthetic inputs with the reference's batch-dict layout.

There is no network and no HDF5 data here, so meshes/fields are generated:
  * ``graph_batch``    -> {'u': [B,N,nt], 'x': [B,N,d], 't': [B,nt]}          (datamodule/dataset_2d.py:54-58)
  * ``implicit_batch`` -> {'t','lr_frames':[B,nt,1,L],'hr_points':[B,nt,Nq,1],
                           'coords_hr':[B,Nq,d],'coords_lr':[B,L,d]}          (datamodule/dataset_2d.py:101-137)
Fields are sums of four travelling sine modes (Burgers-like O(1) magnitude); the time grid
is shared by all samples, as in the real data (this also makes quirk F7 harmless).
"""
import math
from typing import Optional

import torch

BASE_SEED = 20221005


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(BASE_SEED + int(seed))
    return g


def mesh(kind: str, n: int, d: int, g: torch.Generator) -> torch.Tensor:
    """[n, d] fp32 coordinates in [0,1)^d.  kinds: 'uniform' (i.i.d.), 'concentrated' (50 % uniform +
    50 % N(centre, 0.1^2) clipped), 'regular' (meshgrid, row-major), 'sorted1d' (sorted U[0,16))."""
    if kind == "uniform":
        return torch.rand(n, d, generator=g)
    if kind == "concentrated":
        nu = n // 2
        a = torch.rand(nu, d, generator=g)
        b = (0.5 + 0.1 * torch.randn(n - nu, d, generator=g)).clamp(0.0, 1.0 - 1e-6)
        pts = torch.cat([a, b], 0)
        return pts[torch.randperm(n, generator=g)]
    if kind == "regular":
        w = int(round(n ** (1.0 / d)))
        assert w ** d == n, "regular mesh needs n = w^d"
        ax = torch.linspace(0, 1, w + 1)[:-1]
        if d == 1:
            return ax[:, None].clone()
        return torch.stack(torch.meshgrid(ax, ax, indexing="ij"), dim=-1).reshape(-1, 2).contiguous()
    if kind == "sorted1d":
        assert d == 1
        return (16.0 * torch.rand(n, 1, generator=g)).sort(0).values
    raise ValueError(kind)


def field(coords: torch.Tensor, t: torch.Tensor, g: torch.Generator) -> torch.Tensor:
    """u[n, nt] = sum_m a_m sin(2 pi k_m.x + phi_m + w_m t)."""
    n, d = coords.shape
    a = 2 * torch.rand(4, generator=g) - 1
    k = torch.randint(1, 4, (4, d), generator=g).float()
    phi = 2 * math.pi * torch.rand(4, generator=g)
    w = 2 * math.pi * torch.rand(4, generator=g)
    phase = 2 * math.pi * (coords @ k.T) + phi                     # [n, 4]
    return (a * torch.sin(phase[:, None, :] + w * t[:, None])).sum(-1).float()   # [n, nt]


def graph_batch(B: int, N: int, nt: int, d: int = 2, kind: str = "uniform", seed: int = 0,
                t_end: float = 1.0, shared_mesh: bool = True) -> dict:
    """MPNN / MPNN_2d batch.  ``shared_mesh``: the reference reuses sample 0's coordinates for
    every sample (models/mpnn_2d.py:235), so by default all samples carry the same mesh."""
    g = _gen(seed)
    t = torch.linspace(0, t_end, nt)
    xs, us = [], []
    x0 = mesh(kind, N, d, g)
    for b in range(B):
        x = x0 if shared_mesh else mesh(kind, N, d, g)
        xs.append(x)
        us.append(field(x, t, g))
    return {"u": torch.stack(us), "x": torch.stack(xs), "t": t[None].repeat(B, 1)}


def implicit_batch(B: int, L: int, Nq: int, nt: int, d: int = 2, kind: str = "concentrated",
                   seed: int = 0, t_end: float = 1.0) -> dict:
    """MAgNet[GNN] batch: a mesh of L+Nq points per sample, min-max normalised to [-1,1]^d
    (datamodule/dataset_2d.py:101); low-res = the first L points, queries = the other Nq."""
    g = _gen(seed)
    t = torch.linspace(0, t_end, nt)
    out = {k: [] for k in ("lr_frames", "hr_points", "coords_hr", "coords_lr")}
    for b in range(B):
        c = mesh(kind, L + Nq, d, g)
        c = 2 * (c - c.min(0).values) / (c.max(0).values - c.min(0).values) - 1
        u = field(c, t, g)                                          # [L+Nq, nt]
        out["coords_lr"].append(c[:L])
        out["coords_hr"].append(c[L:])
        out["lr_frames"].append(u[:L].T[:, None, :])               # [nt, 1, L]
        out["hr_points"].append(u[L:].T[:, :, None])               # [nt, Nq, 1]
    res = {k: torch.stack(v).contiguous() for k, v in out.items()}
    res["t"] = t[None].repeat(B, 1)
    return res


def seeded_state_dict(shapes: dict, seed: int = 0, dtype=torch.float32) -> dict:
    """Deterministic weights that do not depend on module construction order: every tensor is
    drawn from its own generator keyed on (seed, crc32(name)).  Linear/Conv weights and biases ~
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default bound); LayerNorm (`.1.weight`/`.1.bias`
    of the reference's nn.Sequential(MLP, LayerNorm)) ~ 1 + 0.1 U(-1,1) / 0.1 U(-1,1)."""
    import zlib
    out = {}
    fan = {}
    for name, shape in shapes.items():
        if name.endswith(".weight") and len(shape) >= 2:
            f = 1
            for s in shape[1:]:
                f *= s
            fan[name[:-len(".weight")]] = f
    for name, shape in shapes.items():
        g = torch.Generator(device="cpu")
        g.manual_seed((BASE_SEED + 7919 * int(seed) + zlib.crc32(name.encode())) % (2 ** 31))
        r = 2 * torch.rand(tuple(shape), generator=g) - 1
        base = name.rsplit(".", 1)[0]
        if base in fan:
            v = r / math.sqrt(fan[base])
        elif name.endswith(".weight"):      # LayerNorm gain
            v = 1 + 0.1 * r
        else:                               # LayerNorm bias
            v = 0.1 * r
        out[name] = v.to(dtype)
    return out
