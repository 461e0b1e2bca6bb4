"""GPU: the tcgen05 / TMEM / bulk-copy layer (csrc/umma.cuh) proves itself on the
device against a torch fp32 matmul of the same bf16 operands (tolerance: fp32
accumulation-order noise only, 1e-3 abs on O(10) outputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shift,bulk,n,k", [(0, 0, 64, 64), (3, 0, 192, 192), (2, 1, 384, 192), (4, 1, 160, 192), (1, 1, 384, 80)])
def test_umma_slab_layout_matches_matmul(shift, bulk, n, k):
    from glow_tts_b200 import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(shift * 7 + n)
    rows_a = 132
    a = (torch.randn(rows_a, k, device=dev) * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device=dev) * 0.5).to(torch.bfloat16)
    bp = b.view(n, k // 8, 8).permute(1, 0, 2).contiguous()
    d = torch.zeros(128, n, device=dev)
    rc = _lib.lib().glow_selftest_umma(_lib.ptr(a), _lib.ptr(bp), _lib.ptr(d), rows_a, k, n, shift,
                                       rows_a * 16, 128, n * 16, 128, bulk, _lib.stream_ptr())
    _lib.check(rc, "glow_selftest_umma")
    torch.cuda.synchronize()
    want = a[shift:shift + 128].float() @ b.float().t()
    assert torch.allclose(d, want, atol=1e-3, rtol=1e-4), (d - want).abs().max().item()
