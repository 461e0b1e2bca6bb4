"""GPU: the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cuh, glow_conv_wgrad) against a torch fp32 contraction of
the same bf16-rounded operands: dw[tap][ci][co] = sum_r x[r + tap - c][ci] * g[r][co] over packed rows.  Both
operands are MN-major for the MMA (the reduction runs over rows), every tap re-uses one staged x tile through a
descriptor offset, the row axis is split over CTAs with atomics -- each of which a wrong descriptor or offset would
turn into garbage, not into a small error.  Tolerance: fp32 accumulation-order noise (1e-3 of the largest entry)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(x, g, taps):
    rows = x.shape[0]
    c = (taps - 1) // 2
    xf, gf = x.float(), g.float()
    out = []
    for t in range(taps):
        sh = t - c
        xs = torch.zeros_like(xf)
        lo, hi = max(0, -sh), min(rows, rows - sh)
        xs[lo:hi] = xf[lo + sh:hi + sh]
        out.append(xs.t() @ gf)
    return torch.stack(out)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("taps,cin,cout,rows,split", [
    (1, 192, 192, 256, 1), (1, 192, 384, 1280, 0), (1, 80, 192, 640, 2), (1, 192, 160, 384, 1),
    (5, 192, 384, 1024, 1), (5, 192, 384, 2560, 0), (5, 192, 192, 640, 3), (3, 192, 768, 512, 2), (3, 768, 192, 512, 1),
    (1, 160, 192, 256, 1)])
def test_conv_wgrad_matches_torch(dtype, taps, cin, cout, rows, split):
    from glow_tts_b200 import _lib
    if dtype == "bf16" and (taps == 3 or cin == 160):
        pytest.skip("bf16 operands: only the decoder's shapes are built")
    if dtype == "f32" and cin == 80:
        pytest.skip("fp32 operands: only the encoder's shapes are built")
    dev = torch.device("cuda:0")
    torch.manual_seed(taps * 1000 + cin + cout + rows)
    x = torch.randn(rows, cin, device=dev) * 0.5
    g = torch.randn(rows, cout, device=dev) * 0.5
    x[:2] = 0; x[-2:] = 0; g[:2] = 0; g[-2:] = 0              # packed rows: guard rows lead and trail the axis
    if dtype == "bf16":
        x, g, tag = x.to(torch.bfloat16), g.to(torch.bfloat16), _lib.GLOW_BF16
    else:
        tag = _lib.GLOW_F32
    dw = torch.full((taps, cin, cout), 7.0, device=dev)         # must be overwritten (or zeroed first when split > 1)
    rc = _lib.lib().glow_conv_wgrad(_lib.ptr(x), tag, cin, cin, _lib.ptr(g), cout, cout, None, rows, taps, _lib.ptr(dw),
                                    cout, cin * cout, 0, split, _lib.stream_ptr())
    _lib.check(rc, "glow_conv_wgrad")
    torch.cuda.synchronize()
    want = _reference(x.to(torch.bfloat16), g.to(torch.bfloat16), taps)
    err = (dw - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("taps,cin,cout,rows", [(1, 192, 192, 256), (5, 192, 384, 640), (1, 80, 192, 384), (1, 192, 160, 256)])
def test_conv_wgrad_split_is_fp32_accurate(taps, cin, cout, rows):
    """GLOW_F32_TC operands: hi/lo bf16 split, three MMAs per product -> against the fp64 contraction of the fp32
    operands to 2e-5 of the largest entry (a single bf16 pass is ~4e-3)."""
    from glow_tts_b200 import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(taps + cin + cout)
    x = torch.randn(rows, cin, device=dev)
    g = torch.randn(rows, cout, device=dev)
    x[:2] = 0; x[-2:] = 0; g[:2] = 0; g[-2:] = 0
    dw = torch.empty((taps, cin, cout), device=dev)
    rc = _lib.lib().glow_conv_wgrad(_lib.ptr(x), _lib.GLOW_F32_TC, cin, cin, _lib.ptr(g), cout, cout, None, rows, taps, _lib.ptr(dw),
                                    cout, cin * cout, 0, 1, _lib.stream_ptr())
    _lib.check(rc, "glow_conv_wgrad")
    torch.cuda.synchronize()
    c = (taps - 1) // 2
    want = []
    for t in range(taps):
        sh = t - c
        xs = torch.zeros_like(x, dtype=torch.float64)
        lo, hi = max(0, -sh), min(rows, rows - sh)
        xs[lo:hi] = x.double()[lo + sh:hi + sh]
        want.append(xs.t() @ g.double())
    want = torch.stack(want)
    err = (dw.double() - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err


def test_conv_wgrad_accumulates_and_masks():
    """accumulate = 1 adds to what is there; fp32 operands with a row map read rows with row_utt < 0 as zeros."""
    from glow_tts_b200 import _lib, flow
    dev = torch.device("cuda:0")
    rm = flow.row_map([100, 37, 61], dev)
    rows, cin, cout = rm.rows_pad, 192, 192
    torch.manual_seed(1)
    x = torch.randn(rows, cin, device=dev)                       # junk on guard rows: must not contribute
    g = torch.randn(rows, cout, device=dev)
    valid = (rm.row_utt >= 0).float().unsqueeze(1)
    base = torch.randn(1, cin, cout, device=dev)
    dw = base.clone()
    rc = _lib.lib().glow_conv_wgrad(_lib.ptr(x), _lib.GLOW_F32, cin, cin, _lib.ptr(g), cout, cout, rm.row_utt.data_ptr(), rows, 1,
                                    _lib.ptr(dw), cout, cin * cout, 1, 2, _lib.stream_ptr())
    _lib.check(rc, "glow_conv_wgrad")
    torch.cuda.synchronize()
    want = base + _reference((x * valid).to(torch.bfloat16), (g * valid).to(torch.bfloat16), 1)
    assert (dw - want).abs().max().item() / want.abs().max().item() < 1e-3
