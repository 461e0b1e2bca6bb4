"""Packed-row text encoder (csrc/rows_conv.cu, glow_tts_b200/rows.py) against the torch fp32 path.

The tcgen05 convs round their operands to bf16 (fp32 accumulate), so the tolerance is the bf16 one
the decoder's tensor-core mode states: 2e-2 relative on outputs and gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def rel_fro(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("cin,cout,taps", [(192, 192, 5), (192, 192, 1), (192, 768, 3), (768, 192, 3), (192, 160, 1)])
def test_rows_conv_matches_conv1d(cin, cout, taps):
    from glow_tts_b200 import rows
    torch.manual_seed(cin + cout + taps)
    dev = torch.device("cuda:0")
    lens, t_max = [37, 5, 64, 1, 23], 64
    tr = rows.token_rows(lens, t_max, dev)
    conv = torch.nn.Conv1d(cin, cout, taps, padding=(taps - 1) // 2).to(dev)
    x = torch.randn(len(lens), t_max, cin, device=dev)                       # junk beyond the lengths on purpose
    mask = tr.tmask.unsqueeze(1)                                              # [B,1,T]
    # reference: Conv(x * mask) * mask on [B,C,T]
    xr = x.clone().requires_grad_(True)
    want = conv(xr.transpose(1, 2) * mask) * mask
    g = torch.randn_like(want)
    want.backward(g)
    want_dx, want_dw, want_db = xr.grad * tr.tmask.unsqueeze(2), conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.weight.grad = conv.bias.grad = None
    # packed rows, with junk on the guard rows of the input (the loader must mask it)
    xp = (tr.pack(x) + (1.0 - tr.valid) * 7.0).detach().requires_grad_(True)
    got_rows = rows.rows_conv(xp, conv, tr)
    assert float((got_rows * (1 - tr.valid)).abs().max()) == 0.0              # guard rows are written as zeros
    got = tr.unpack(got_rows).transpose(1, 2)
    assert rel(got, want.detach()) < 2e-2
    grows = tr.pack(g.transpose(1, 2)) + (1.0 - tr.valid) * 3.0               # junk gradient on guard rows
    got_rows.backward(grows)
    assert rel(tr.unpack(xp.grad), want_dx) < 2e-2
    assert rel(conv.weight.grad, want_dw) < 2e-2
    assert rel(conv.bias.grad, want_db) < 2e-2


@pytest.mark.parametrize("relu,with_b", [(False, True), (True, False)])
def test_rows_norm_matches_layernorm(relu, with_b):
    """Eval mode (no dropout): mask * ReLU?(LayerNorm(a + b)) and its gradients, fp32 both ways: 1e-4."""
    from glow_tts_b200 import rows
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    tr = rows.token_rows([37, 5, 64, 1, 23], 64, dev)
    ln = torch.nn.LayerNorm(192, eps=1e-4).to(dev)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(192, device=dev)); ln.bias.copy_(torch.randn(192, device=dev))
    a0 = torch.randn(tr.rows_pad, 192, device=dev)
    b0 = torch.randn(tr.rows_pad, 192, device=dev) if with_b else None
    g = torch.randn(tr.rows_pad, 192, device=dev)
    res = []
    for path in ("torch", "native"):
        ln.weight.grad = ln.bias.grad = None
        a = a0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True) if with_b else None
        if path == "torch":
            y = ln(a + b if with_b else a)
            y = (F.relu(y) if relu else y) * tr.valid
        else:
            y = rows.rows_norm(a, b, ln, tr, relu=relu)
        y.backward(g)
        res.append((y.detach(), a.grad * tr.valid, (b.grad * tr.valid) if with_b else None, ln.weight.grad.clone(),
                    ln.bias.grad.clone()))
    for got, want in zip(res[1], res[0]):
        if want is not None:
            assert rel(got, want) < 1e-4


def test_rows_dropout_forward_backward_agree():
    """Training mode: the masks the forward kernels draw are the ones their backward replays: with an
    all-ones upstream gradient, d(sum y)/dx from autograd must equal a central finite difference."""
    from glow_tts_b200 import rows
    torch.manual_seed(9)
    dev = torch.device("cuda:0")
    tr = rows.token_rows([20, 9], 20, dev)
    ln = torch.nn.LayerNorm(192, eps=1e-4).to(dev)
    a = (torch.randn(tr.rows_pad, 192, device=dev) * tr.valid).requires_grad_(True)
    b = (torch.randn(tr.rows_pad, 192, device=dev) * tr.valid)
    y = rows.rows_norm(a, b, ln, tr, p_in=0.3, seed_in=1234, relu=True, p_out=0.5, seed_out=99)
    frac = float((y[tr.valid[:, 0] > 0] == 0).float().mean())
    assert 0.55 < frac < 0.95                                   # ReLU (about half) and dropout 0.5 both zero outputs
    y.sum().backward()
    r, c, eps = int(tr.rm.utt_off[0]) + 3, 17, 1e-2
    with torch.no_grad():
        ap, am = a.detach().clone(), a.detach().clone()
        ap[r, c] += eps; am[r, c] -= eps
        fd = (rows.rows_norm(ap, b, ln, tr, p_in=0.3, seed_in=1234, relu=True, p_out=0.5, seed_out=99).sum()
              - rows.rows_norm(am, b, ln, tr, p_in=0.3, seed_in=1234, relu=True, p_out=0.5, seed_out=99).sum()) / (2 * eps)
    assert abs(float(fd) - float(a.grad[r, c])) < 5e-2 * max(1.0, abs(float(fd)))
    # conv with fused ReLU + dropout: same masks in backward
    conv = torch.nn.Conv1d(192, 768, 3, padding=1).to(dev)
    x = (torch.randn(tr.rows_pad, 192, device=dev) * tr.valid).requires_grad_(True)
    f = rows.rows_conv(x, conv, tr, relu=True, p=0.4, seed=77, x_masked=True)
    f.sum().backward()
    with torch.no_grad():
        f2 = rows.rows_conv(x.detach(), conv, tr, relu=True, p=0.4, seed=77)
        assert torch.equal(f.detach(), f2)                       # the mask is a pure function of (seed, row, column)
        keep = (f2 != 0).float()                                 # d(sum f)/d(conv out) = keep / (1 - p)
        want_db = (keep / 0.6).sum(0)
    assert rel(conv.bias.grad, want_db) < 1e-3


@pytest.mark.parametrize("name", ["vanilla_small", "se_small"])
def test_rows_encoder_matches_torch_encoder(name):
    from tests._model_util import load_case
    model, sd, g, batch, mode = load_case(name, "bf16")
    model.eval()
    enc = model.layer_Dict["Encoder"]
    assert enc.rows_supported
    tokens, tl, mels, ml, spk = batch
    dev = torch.device("cuda:0")
    tokens = tokens.to(dev)[:, :int(tl.max())]
    lens = [int(v) for v in tl.tolist()]
    mask = model._masks_from_host(lens, dev)
    t_len = torch.tensor(lens, dtype=torch.int32, device=dev)
    spk_e = model.layer_Dict["LUT"](spk.to(dev)) if "LUT" in model.layer_Dict else None
    outs = {}
    grads = {}
    probe = [p for n, p in enc.named_parameters() if n.endswith(("Prenet.layer_Dict.CLRD_0.layer_Dict.Conv.weight",
                                                                 "ANCRDCN_5.layer_Dict.Conv_1.weight",
                                                                 "ANCRDCN_0.layer_Dict.Attention.weight_K",
                                                                 "Embedding.weight", "Project.bias"))]
    assert len(probe) == 5
    for path in ("torch", "rows"):
        for p in enc.parameters():
            p.grad = None
        if path == "torch":
            mean, log_std, log_dur, _ = enc._forward(tokens, mask, spk_e, t_len)
        else:
            mean, log_std, log_dur, _ = enc._forward_rows(tokens, mask, spk_e, t_len, lens)
        outs[path] = (mean.detach(), log_std.detach(), log_dur.detach())
        torch.manual_seed(3)
        # explicit shapes: randn_like fills in memory order, and the two paths return differently strided views
        r1, r2 = torch.randn(mean.shape, device=dev), torch.randn(log_std.shape, device=dev)
        loss = (mean * r1).sum() + (log_std * r2).sum() + log_dur.sum()
        loss.backward()
        grads[path] = [p.grad.detach().clone() for p in probe]
    for a, b in zip(outs["rows"], outs["torch"]):
        assert rel(a, b) < 2e-2
    # Parameters with no ReLU between them and the loss (Project, the last block's Conv_1) see only
    # bf16 operand rounding: 2e-2.  Further upstream the two paths differentiate slightly different
    # functions: a pre-activation within bf16 rounding of zero takes the other branch of the ReLU
    # (Modules.py:478,569), which flips that element's gradient -- measured 5-10 % Frobenius, the
    # same effect torch.autocast(bf16) has on this encoder.  Bound: 1.5e-1.
    names = [n for n, _ in enc.named_parameters()]
    for p, a, b in zip(probe, grads["rows"], grads["torch"]):
        n = [k for k, q in enc.named_parameters() if q is p][0]
        tight = n.endswith(("Project.bias", "ANCRDCN_5.layer_Dict.Conv_1.weight"))
        assert rel_fro(a, b) < (2e-2 if tight else 1.5e-1), n


def test_side_stream_weight_gradients_match_autograd_path():
    """TrainStep lets the encoder convs accumulate their weight gradients into the flat gradient buffer
    on the side stream; the result must equal the autograd path (gradients returned and accumulated by
    torch) on the same step."""
    from glow_tts_b200 import rows
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep
    from tests._model_util import load_case
    grads = []
    for accumulate in (True, False):
        model, sd, g, batch, mode = load_case("vanilla_small", "bf16")
        model.eval()                                               # no dropout: the two runs are the same function
        step = TrainStep(model, load_hparams(Mode=mode, Precision="bf16"), torch.device("cuda:0"))
        step.opt.lr0, step.opt.wd, step.opt.max_norm = 0.0, 0.0, 0.0      # keep weights and gradients as they are
        rows.ACCUMULATE = accumulate
        step.run(step.to_device(batch))
        torch.cuda.synchronize()
        grads.append(step.flat.grad.detach().clone())
    rows.ACCUMULATE = False
    assert float(grads[0].abs().max()) > 0
    assert rel_fro(grads[0], grads[1]) < 1e-5
