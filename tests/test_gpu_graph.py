"""GPU parity: integer primitives, radius graph, kNN and aggregation plans — bit-exact vs the oracle."""
import pytest
import torch

from oracle import graph as OG
from magnet_b200 import _lib, graph as MG, synthetic as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n,bits", [(1, 8), (257, 8), (5000, 13), (100_000, 20), (1_000_003, 24), (70_000, 32)])
def test_radix_sort_is_stable(n, bits):
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(n)
    hi = (1 << min(bits, 31)) - 1
    keys = torch.randint(0, max(hi // 7, 2), (n,), generator=g, dtype=torch.int64)   # many duplicates
    if bits == 32:
        keys = keys | (torch.randint(0, 2, (n,), generator=g, dtype=torch.int64) << 31)
    k32 = keys.to(torch.uint32 if hasattr(torch, "uint32") else torch.int32).to(DEV)
    k_i = keys.to(DEV)
    vals = torch.arange(n, dtype=torch.int32, device=DEV)
    ko, vo = torch.empty(n, dtype=torch.int32, device=DEV), torch.empty(n, dtype=torch.int32, device=DEV)
    kin = (keys & 0xffffffff).to(torch.int64)
    kin32 = torch.where(kin >= 2 ** 31, kin - 2 ** 32, kin).to(torch.int32).to(DEV)
    ws = _lib.workspace(L.mgb_sort_workspace(n), DEV)
    _lib.check(L.mgb_sort_pairs_u32(_lib.ptr(kin32), _lib.ptr(vals), _lib.ptr(ko), _lib.ptr(vo), n, bits, _lib.ptr(ws),
                                    ws.numel(), _lib.stream()))
    want = torch.sort(k_i, stable=True)
    assert torch.equal(vo.long(), want.indices)


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 300_000, 5_000_000])
def test_exclusive_scan(n):
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(n + 1)
    x = torch.randint(0, 40, (max(n, 1),), generator=g, dtype=torch.int32)[:n].to(DEV)
    out = torch.empty(n + 1, dtype=torch.int32, device=DEV)
    ws = _lib.workspace(L.mgb_scan_workspace(n), DEV)
    _lib.check(L.mgb_exclusive_scan_i32(_lib.ptr(x), _lib.ptr(out), n, _lib.ptr(ws), ws.numel(), _lib.stream()))
    want = torch.cat([torch.zeros(1, dtype=torch.int64, device=DEV), x.long().cumsum(0)])
    assert torch.equal(out.long(), want)


def test_radius_graph_golden(golden):
    for name, c in golden("radius_graph.pt").items():
        ei = MG.radius_graph(c["x"].to(DEV), c["r"], c["batch"].to(DEV), loop=c["loop"])
        assert ei.dtype == torch.int64
        assert torch.equal(ei.cpu(), c["edge_index"]), name


CASES = [  # kind, B, N, d, r, loop, max_nbrs
    ("uniform", 4, 4096, 2, 0.03, False, 32),          # config-2 mesh, ~11 neighbours, no truncation
    ("uniform", 2, 4096, 2, 0.09, False, 32),          # ~100 candidates: truncation active (F4)
    ("regular", 2, 4096, 2, 4 * (2 ** 0.5) / 64 + 1e-4, False, 32),   # reference radius formula on a 64x64 grid
    ("concentrated", 3, 512, 2, 0.04, True, 32),
    ("concentrated", 1, 20000, 2, 0.04, True, 32),     # dense core: thousands of points per cell
    ("sorted1d", 16, 50, 1, 1.0, False, 32),
    ("uniform", 1, 1, 2, 0.5, True, 32),               # single node
    ("uniform", 3, 200, 2, 10.0, True, 4),             # radius >> extent, tiny cap
    ("uniform", 3, 200, 2, 1e-6, False, 32),           # no edges at all
]


@pytest.mark.parametrize("kind,B,N,d,r,loop,cap", CASES)
def test_radius_graph_vs_oracle(kind, B, N, d, r, loop, cap):
    g = S._gen(hash((kind, B, N)) % 1000)
    x = torch.cat([S.mesh(kind, N, d, g) for _ in range(B)], 0)
    batch = torch.arange(B).repeat_interleave(N)
    want = OG.radius_graph(x, r, batch, loop=loop, max_num_neighbors=cap, threads=8)
    got = MG.radius_graph(x.to(DEV), r, batch.to(DEV), loop=loop, max_num_neighbors=cap)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.equal(got.cpu(), want)
    swapped = MG.radius_graph(x.to(DEV), r, batch.to(DEV), loop=loop, max_num_neighbors=cap, swap_rows=True)
    assert torch.equal(swapped.cpu(), want.flip(0))


def test_radius_graph_ragged_and_empty_samples():
    g = S._gen(77)
    sizes = [0, 300, 1, 0, 57, 1000]
    x = torch.cat([S.mesh("uniform", n, 2, g) for n in sizes if n > 0], 0)
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    batch = torch.cat([torch.full((n,), b) for b, n in enumerate(sizes)]).long()
    # oracle takes ptr through the batch vector (empty samples in the middle are fine, trailing ones are not needed)
    want = OG.radius_graph(x, 0.1, batch, loop=False)
    got = MG.radius_graph(x.to(DEV), 0.1, ptr=ptr.to(DEV), loop=False)
    assert torch.equal(got.cpu(), want)


def test_radius_graph_deterministic():
    g = S._gen(5)
    x = S.mesh("uniform", 30000, 2, g).to(DEV)
    a = MG.radius_graph(x, 0.02, None, loop=True)
    b = MG.radius_graph(x, 0.02, None, loop=True)
    assert torch.equal(a, b)


def test_knn_golden(golden):
    for name, c in golden("knn.pt").items():
        ai = MG.knn(c["x"].to(DEV), c["y"].to(DEV), c["k"], c["batch_x"].to(DEV), c["batch_y"].to(DEV))
        assert torch.equal(ai.cpu(), c["assign_index"]), name


@pytest.mark.parametrize("d,B,L,Nq,k", [(2, 2, 5000, 3000, 4), (2, 1, 20000, 10000, 16), (2, 3, 700, 500, 32),
                                        (1, 4, 25, 16, 4), (2, 2, 3, 50, 4), (2, 1, 40000, 20000, 8)])
def test_knn_vs_oracle(d, B, L, Nq, k):
    g = S._gen(L + Nq + k)
    kind = "concentrated" if L >= 20000 else "uniform"
    xl = torch.cat([2 * S.mesh(kind if d == 2 else "uniform", L, d, g) - 1 for _ in range(B)], 0)
    xq = torch.cat([2.4 * S.mesh("uniform", Nq, d, g) - 1.2 for _ in range(B)], 0)    # some queries outside the hull
    bl, bq = torch.arange(B).repeat_interleave(L), torch.arange(B).repeat_interleave(Nq)
    want = OG.knn(xl, xq, k, bl, bq, threads=8)
    got = MG.knn(xl.to(DEV), xq.to(DEV), k, bl.to(DEV), bq.to(DEV))
    assert got.shape == want.shape
    assert torch.equal(got.cpu(), want)


def test_csr_plan_matches_stable_sort():
    g = torch.Generator().manual_seed(3)
    n, E = 5000, 120_000
    ei = torch.randint(0, n, (2, E), generator=g).to(DEV)
    ei[1, :1000] = 7          # one long segment spanning many tiles
    plan = MG.build_plan(ei, n)
    order = torch.sort(ei[1], stable=True).indices
    assert torch.equal(plan.perm.long(), order)
    assert torch.equal(plan.dst.long(), ei[1][order]) and torch.equal(plan.src.long(), ei[0][order])
    counts = torch.bincount(ei[1], minlength=n)
    assert torch.equal(plan.rowptr.long(), torch.cat([counts.new_zeros(1), counts.cumsum(0)]))
    order_t = torch.sort(ei[0], stable=True).indices
    inv = torch.empty_like(order)
    inv[order] = torch.arange(E, device=DEV)
    assert torch.equal(plan.pos_t.long(), inv[order_t])
    counts_t = torch.bincount(ei[0], minlength=n)
    assert torch.equal(plan.rowptr_t.long(), torch.cat([counts_t.new_zeros(1), counts_t.cumsum(0)]))
    with pytest.raises(RuntimeError):
        bad = ei.clone()
        bad[0, 5] = n
        MG.build_plan(bad, n)
