"""GPU: the tcgen05 / TMEM building blocks (csrc/umma.cuh) — one 128x128x128 bf16 tile through every
K-major / MN-major descriptor combination over the same swizzled tile image, vs an fp64 product of the
bf16-rounded inputs (bit-level agreement is not expected: accumulation order differs; 1e-6 is)."""
import pytest
import torch

from conftest import rel_err
from magnet_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_umma_tile_all_majors(a_mn, b_mn):
    L = _lib.lib()
    g = torch.Generator().manual_seed(7 + 2 * a_mn + b_mn)
    A = torch.randn(128, 128, generator=g)     # logical A[m][k]
    B = torch.randn(128, 128, generator=g)     # logical B[n][k]
    a_store = (A.t() if a_mn else A).contiguous().cuda()
    b_store = (B.t() if b_mn else B).contiguous().cuda()
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(L.mgb_umma_selftest(_lib.ptr(a_store), _lib.ptr(b_store), a_mn, b_mn, 0, 0, _lib.ptr(D), _lib.stream()),
               "umma_selftest")
    torch.cuda.synchronize()
    want = A.bfloat16().double() @ B.bfloat16().double().t()
    assert rel_err(D, want) < 1e-6, (a_mn, b_mn)


@pytest.mark.parametrize("b_mn", [0, 1])
def test_umma_a_operand_in_tensor_memory(b_mn):
    """TS form of tcgen05.mma: A read from tensor memory (lane = row m, packed bf16 pairs along K, written with tcgen05.st),
    B from shared memory through either descriptor kind."""
    g = torch.Generator().manual_seed(7 + b_mn)
    A = torch.randn(128, 128, generator=g)
    B = torch.randn(128, 128, generator=g)
    L = _lib.lib()
    a_store = A.contiguous().cuda()
    b_store = (B.t() if b_mn else B).contiguous().cuda()
    D = torch.full((128, 128), float("nan"), device="cuda")
    _lib.check(L.mgb_umma_selftest(_lib.ptr(a_store), _lib.ptr(b_store), 2, b_mn, 0, 0, _lib.ptr(D), _lib.stream()), "umma_selftest")
    torch.cuda.synchronize()
    want = A.bfloat16().double() @ B.bfloat16().double().t()
    assert rel_err(D, want) < 1e-6, b_mn
