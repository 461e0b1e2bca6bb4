"""GPU: glow_rpr_attention_forward/_backward behind the drop-in RPR_Multihead_Attention
vs fixtures from the reference module (RPR_MHA.py:69-165).  Tolerance 1e-3 relative."""
import pytest
import torch

from tests._model_util import CASES, load_case
from tests._util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("mask_kind", ["mask", "lengths"])
def test_attention_matches_reference(name, mask_kind):
    model, sd, g, (tokens, tl, mels, ml, spk), mode = load_case(name, "fp32")
    model.eval()
    att = model.layer_Dict["Encoder"].layer_Dict["Transformer"].layer_Dict["ANCRDCN_0"].layer_Dict["Attention"]
    x = torch.from_numpy(g["att_x"]).cuda().requires_grad_(True)
    tm = (torch.arange(max(tl))[None, :] < tl[:, None]).unsqueeze(1).float().cuda()
    if mask_kind == "mask":
        out, align = att(queries=x, masks=(tm * tm.transpose(2, 1)).unsqueeze(1))
    else:
        out, align = att(queries=x, lengths=tl.cuda())
    assert rel_err(out.cpu(), g["att_out"]) < 1e-3
    assert rel_err(align.cpu(), g["att_align"]) < 1e-3
    gen = torch.Generator().manual_seed(5)
    torch.randn(x.shape, generator=gen)
    model.zero_grad(set_to_none=True)
    (out * torch.randn(out.shape, generator=gen).cuda()).sum().backward()
    assert rel_err(x.grad.cpu(), g["att_dx"]) < 1e-3
    assert rel_err(att.weight_K.grad.cpu(), g["att_dwk"]) < 1e-3
    assert rel_err(att.weight_V.grad.cpu(), g["att_dwv"]) < 1e-3
    assert rel_err(att.layer_Dict["Query"].weight.grad.cpu(), g["att_dqw"]) < 1e-3


def test_attention_dropout_is_statistical_and_consistent():
    """Training-mode dropout cannot match torch's RNG stream; check rate, scaling and that
    backward uses the same mask (finite-difference on a linear functional)."""
    model, sd, g, (tokens, tl, mels, ml, spk), mode = load_case("vanilla_small", "fp32")
    att = model.layer_Dict["Encoder"].layer_Dict["Transformer"].layer_Dict["ANCRDCN_0"].layer_Dict["Attention"]
    x = torch.from_numpy(g["att_x"]).cuda()
    att.eval()
    _, a_eval = att(queries=x, lengths=tl.cuda())
    att.train()
    # pool the valid cells of every utterance over several calls (each call draws a new mask) so the
    # keep-rate bound is a 5-sigma one for the pooled sample instead of a fixed band on a few hundred cells
    kept, cells = 0.0, 0
    for _ in range(8):
        _, a_train = att(queries=x, lengths=tl.cuda())
        for b in range(x.shape[0]):
            n = int(tl[b])
            nz = a_train[b, :, :n, :n] != 0
            kept += nz.float().sum().item()
            cells += nz.numel()
            assert rel_err(a_train[b, :, :n, :n][nz], (a_eval[b, :, :n, :n] / 0.9)[nz]) < 1e-5
    sigma = (0.9 * 0.1 / cells) ** 0.5
    assert abs(kept / cells - 0.9) < 5 * sigma + 1e-3, (kept / cells, cells)


def test_rows_tensor_core_forward_matches_cuda_core_kernel():
    """Packed-row layout: Q.K^T and Pd.V on mma.sync (bf16 operands) against the fp32 CUDA-core kernels
    on the [B,C,T] layout, same inputs, eval mode.  bf16 operand rounding: 2e-2 of the largest output."""
    import torch
    from glow_tts_b200 import rows
    from glow_tts_b200.rpr_mha import _AttnCoreFn, _AttnRowsFn
    torch.manual_seed(11)
    dev = torch.device("cuda:0")
    lens, t_max, heads, d, window = [150, 64, 7, 201, 33], 201, 2, 96, 4
    tr = rows.token_rows(lens, t_max, dev)
    lengths = torch.tensor(lens, dtype=torch.int32, device=dev)
    q, k, v = (torch.randn(len(lens), heads * d, t_max, device=dev) * tr.tmask.unsqueeze(1) for _ in range(3))
    wk = torch.randn(1, 2 * window + 1, d, device=dev) * d ** -0.5
    wv = torch.randn(1, 2 * window + 1, d, device=dev) * d ** -0.5
    for x in (q, k, v, wk, wv):
        x.requires_grad_(True)
    want, _ = _AttnCoreFn.apply(q, k, v, wk, wv, lengths, None, heads, window, 0.0, 0, False)
    qr, kr, vr = (tr.pack(x.detach().transpose(1, 2).contiguous()).requires_grad_(True) for x in (q, k, v))
    wk2, wv2 = wk.detach().clone().requires_grad_(True), wv.detach().clone().requires_grad_(True)
    got_rows = _AttnRowsFn.apply(qr, kr, vr, wk2, wv2, tr, lengths, heads, window, 0.0, 0)
    got = tr.unpack(got_rows).transpose(1, 2)
    want = want * tr.tmask.unsqueeze(1)                            # the [B,C,T] kernel also fills padded queries
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-2, err
    assert float((got_rows * (1 - tr.valid)).abs().max()) == 0.0   # guard rows untouched (zeros)
    # backward: same upstream gradient through both implementations
    gsrc = torch.randn(len(lens), heads * d, t_max, device=dev) * tr.tmask.unsqueeze(1)
    (want * gsrc).sum().backward()
    got_rows.backward(tr.pack(gsrc.transpose(1, 2).contiguous()))
    def unp(x):
        return tr.unpack(x).transpose(1, 2)
    for name, a, b in (("dq", unp(qr.grad), q.grad * tr.tmask.unsqueeze(1)), ("dk", unp(kr.grad), k.grad * tr.tmask.unsqueeze(1)),
                       ("dv", unp(vr.grad), v.grad * tr.tmask.unsqueeze(1)), ("dwk", wk2.grad, wk.grad), ("dwv", wv2.grad, wv.grad)):
        e = float((a - b).abs().max() / b.abs().max())
        assert e < 3e-2, (name, e)
