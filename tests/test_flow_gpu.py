"""GPU: the flow decoder kernels (glow_flow_* through the C ABI, behind the drop-in
Decoder module) vs fixtures made by running the reference's Decoder (Modules.py:286-309).
Tolerance: 1e-3 relative (max|a-b|/max|b|, BASELINE north_star) in fp32 mode."""
import numpy as np
import pytest
import torch

from tests._model_util import CASES, digest, load_case, mel_mask
from tests._util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


# fp32: CUDA-core GEMMs.  fp32-tc: the same fp32 storage on tcgen05 with the bf16 hi/lo operand split (three MMAs per
# product, fp32 accumulate) -- the tensor-core mode that has to meet the same 1e-3.
@pytest.fixture(scope="module", params=[(n, p) for n in CASES for p in ("fp32", "fp32-tc")], ids=lambda v: "%s-%s" % v)
def case(request):
    return load_case(request.param[0], request.param[1])


def _spk(model, mode, spk):
    return model.layer_Dict["LUT"](spk.cuda()).detach() if mode == "SE" else None


def test_forward_matches_reference(case):
    model, sd, g, (tokens, tl, mels, ml, spk), mode = case
    model.eval()
    dec = model.layer_Dict["Decoder"]
    with torch.no_grad():
        z, ld, mask = dec(mels.cuda(), mel_mask(ml, "cuda"), _spk(model, mode, spk))
    assert rel_err(z.cpu(), g["dec_z"]) < TOL
    assert rel_err(ld.cpu(), g["dec_logdet"]) < TOL
    pad = torch.from_numpy(g["dec_z"]) == 0
    assert float(z.cpu()[pad].abs().max()) == 0.0            # padded frames are exactly zero


def test_reverse_matches_reference_and_inverts(case):
    model, sd, g, (tokens, tl, mels, ml, spk), mode = case
    model.eval()
    dec = model.layer_Dict["Decoder"]
    m = mel_mask(ml, "cuda")
    with torch.no_grad():
        zin = torch.from_numpy(g["dec_z"]).cuda()
        back, ld, _ = dec(zin, m, _spk(model, mode, spk), reverse=True)
    assert ld is None
    # The inverse of this synthetic fixture is ill-conditioned: it amplifies any forward-side rounding ~500x (two fp32
    # evaluations with different op order already differ by 5e-4: tests/test_oracle_model.py).  fp32: 2e-3.  fp32-tc
    # carries 2^-16 per product (measured forward error 2e-5 -> reverse 1.5e-2): 3e-2.
    tol = 2e-3 if dec.precision == "fp32" else 3e-2
    assert rel_err(back.cpu(), g["dec_reverse_of_z"]) < tol
    assert rel_err(back.cpu(), (mels * m.cpu())) < tol           # Decoder(reverse) o Decoder == id


def test_backward_matches_reference(case):
    model, sd, g, (tokens, tl, mels, ml, spk), mode = case
    model.eval()
    dec = model.layer_Dict["Decoder"]
    model.zero_grad(set_to_none=True)
    z, ld, _ = dec(mels.cuda(), mel_mask(ml, "cuda"), _spk(model, mode, spk))
    gen = torch.Generator().manual_seed(99)
    rz, rl = torch.randn(z.shape, generator=gen), torch.randn(ld.shape, generator=gen)
    ((z * rz.cuda()).sum() + (ld * rl.cuda()).sum()).backward()
    params = dict(dec.named_parameters())
    scale = float(g["dec_grad_digest"][:, 0].max())
    for key, want in zip(g["dec_grad_keys"], g["dec_grad_digest"]):
        got = digest(params[str(key)].grad, 7)
        assert abs(got[0] - want[0]) <= 2e-3 * want[0] + 1e-6 * scale, key
        assert abs(got[1] - want[1]) <= 2e-3 * want[0] + 1e-6 * scale, key
    flows = dec.layer_Dict["Flows"]
    assert rel_err(flows[0].layers[0].logs.grad.cpu(), g["dec_grad_b0_logs"]) < 2e-3
    assert rel_err(flows[0].layers[1].weight.grad.cpu(), g["dec_grad_b0_w"]) < 2e-3
    assert rel_err(flows[0].layers[2].layer_Dict["Start"].weight_g.grad.cpu(), g["dec_grad_b0_start_g"]) < 2e-3
    assert rel_err(flows[-1].layers[2].layer_Dict["End"].weight.grad.cpu(), g["dec_grad_b11_end_w"]) < 2e-3


def test_padding_is_dead_work(case):
    """An utterance decoded alone equals the same utterance inside a padded batch."""
    model, sd, g, (tokens, tl, mels, ml, spk), mode = case
    model.eval()
    dec = model.layer_Dict["Decoder"]
    e = _spk(model, mode, spk)
    with torch.no_grad():
        z, ld, _ = dec(mels.cuda(), mel_mask(ml, "cuda"), e)
        i = len(ml) - 1
        n = int(ml[i])
        z1, ld1, _ = dec(mels[i:i + 1, :, :n].cuda(), mel_mask(ml[i:i + 1], "cuda"), None if e is None else e[i:i + 1])
    assert rel_err(z1, z[i:i + 1, :, :n]) < 1e-5 and rel_err(ld1, ld[i:i + 1]) < 1e-5


def test_bf16_mode_close_to_reference():
    """bf16 operands cannot meet 1e-3 (eps_bf16 = 3.9e-3); stated tolerance 5e-2 on z, 2e-2 on logdet."""
    model, sd, g, (tokens, tl, mels, ml, spk), mode = load_case("vanilla_small", "bf16")
    model.eval()
    with torch.no_grad():
        z, ld, _ = model.layer_Dict["Decoder"](mels.cuda(), mel_mask(ml, "cuda"), None)
    assert rel_err(z.cpu(), g["dec_z"]) < 5e-2
    assert rel_err(ld.cpu(), g["dec_logdet"]) < 2e-2


def _fwd_bwd(precision, mode, seed, tls, mls, bseed):
    from tests._model_util import build_model
    from tests._util import synth_batch
    model, _ = build_model(mode, seed, precision)
    model.eval()
    tokens, tl, mels, ml, spk = synth_batch(bseed, tls, mls)
    dec = model.layer_Dict["Decoder"]
    model.zero_grad(set_to_none=True)
    e = _spk(model, mode, spk)
    z, ld, _ = dec(mels.cuda(), mel_mask(ml, "cuda"), e)
    gen = torch.Generator().manual_seed(5)
    rz, rl = torch.randn(z.shape, generator=gen), torch.randn(ld.shape, generator=gen)
    ((z * rz.cuda()).sum() + (ld * rl.cuda()).sum()).backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in dec.named_parameters()}
    with torch.no_grad():
        back, _, _ = dec(z.detach(), mel_mask(ml, "cuda"), e, reverse=True)
    return z.detach().cpu(), ld.detach().cpu(), grads, back.cpu()


@pytest.mark.parametrize("mode,tls,mls", [("Vanilla", [23, 17, 9], [140, 96, 50]),
                                          ("SE", [40, 31, 25, 12, 50], [612, 400, 258, 64, 1000])])
def test_tensor_core_path_matches_cuda_core_path(mode, tls, mls):
    """GLOW_BF16 (tcgen05, fp32 accumulate in TMEM) vs GLOW_BF16_SIMT (same bf16 storage, fp32 CUDA-core
    GEMM): only the accumulation order differs, so outputs agree to a few bf16 ulps of the activations
    (stated tolerance 2e-2 of the max magnitude; measured ~3e-3) and gradients to 6e-2 per tensor (measured worst 3.8e-2, a weight_g)."""
    z1, ld1, g1, back1 = _fwd_bwd("bf16", mode, 77, tls, mls, 8)
    z2, ld2, g2, back2 = _fwd_bwd("bf16-simt", mode, 77, tls, mls, 8)
    assert torch.isfinite(z1).all() and torch.isfinite(ld1).all()
    assert rel_err(z1, z2) < 2e-2
    assert rel_err(ld1, ld2) < 5e-3
    # the inverse amplifies bf16 rounding of the 12 chained couplings (max-norm outliers), so the two
    # reverse passes -- each fed its own z -- are compared in the mean
    assert float((back1 - back2).abs().mean() / back2.abs().mean()) < 3e-2
    worst = max((rel_err(g1[k], g2[k]), k) for k in g1 if float(g2[k].abs().max()) > 0)
    assert worst[0] < 6e-2, worst


def test_training_dropout_masks_agree_between_forward_and_backward():
    """Dropout on the gate pre-activation (Modules.py:862) is counter-based in the kernels: backward
    recomputes the forward's keep mask.  With the dropout stream pinned, the analytic gradient must
    equal a central finite difference of the (deterministic) forward -- fp32 mode, p = 0.3 so a
    mask mismatch would be a >10% error; tolerance 2e-2."""
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    from tests._util import synth_batch, synth_state_dict
    modules.set_hparams(load_hparams(Mode="Vanilla", Precision="fp32",
                                     **{"Decoder.Affine_Coupling.WaveNet.Dropout_Rate": 0.3, "Decoder.Stack": 4}))
    model = modules.GlowTTS()
    model.load_state_dict(synth_state_dict(model.state_dict(), 5), strict=True)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = True
    model = model.cuda()
    model.train()
    dec = model.layer_Dict["Decoder"]
    tokens, tl, mels, ml, spk = synth_batch(3, [20, 12], [120, 64])
    x, m = mels.cuda(), mel_mask(ml, "cuda")
    gen = torch.Generator().manual_seed(1)
    rz = torch.randn(x.shape, generator=gen).cuda()
    rl = torch.randn(len(ml), generator=gen).cuda()

    def loss():
        dec._step = 41                                   # same dropout stream on every call
        z, ld, _ = dec(x, m, None)
        return (z * rz).sum() + (ld * rl).sum()

    model.zero_grad(set_to_none=True)
    base = loss()
    base.backward()
    params = [p for p in dec.parameters() if p.grad is not None]
    dirs = [torch.randn(p.shape, generator=gen).cuda() * p.detach().abs().mean() for p in params]
    dot = float(sum((p.grad.double() * d.double()).sum() for p, d in zip(params, dirs)))
    with torch.no_grad():
        again = loss()
        assert float((again - base).abs()) <= 1e-5 * float(base.abs())       # deterministic given the seed
        eps = 2e-3
        for p, d in zip(params, dirs):
            p.add_(eps * d)
        lp = float(loss().double())
        for p, d in zip(params, dirs):
            p.sub_(2 * eps * d)
        lm = float(loss().double())
    fd = (lp - lm) / (2 * eps)
    assert abs(fd - dot) <= 2e-2 * abs(dot), (fd, dot)


@pytest.mark.parametrize("mode,tls,mls", [("Vanilla", [23, 17, 9], [140, 96, 50]),
                                          ("SE", [40, 31, 25, 12, 50], [612, 400, 258, 64, 1000])])
def test_split_tensor_core_path_matches_cuda_core_fp32(mode, tls, mls):
    """GLOW_F32_TC (tcgen05, bf16 hi/lo split, own weight-gradient kernel) vs GLOW_F32 (CUDA-core GEMMs, library weight
    gradients): two independent fp32-class evaluations of the same arithmetic at a second geometry (one that crosses
    several 128-row tiles); forward / logdet to 1e-4, every parameter gradient to 2e-3, the reverse pass to 3e-2."""
    z1, ld1, g1, back1 = _fwd_bwd("fp32-tc", mode, 77, tls, mls, 8)
    z2, ld2, g2, back2 = _fwd_bwd("fp32", mode, 77, tls, mls, 8)
    assert rel_err(z1, z2) < 1e-4, rel_err(z1, z2)
    assert rel_err(ld1, ld2) < 1e-4, rel_err(ld1, ld2)
    assert rel_err(back1, back2) < 3e-2, rel_err(back1, back2)      # ill-conditioned inverse, see above
    worst = max((rel_err(g1[k], g2[k]), k) for k in g1 if float(g2[k].abs().max()) > 0)
    assert worst[0] < 2e-3, worst


@pytest.mark.parametrize("mode,tls,mls", [("Vanilla", [23, 17, 9], [140, 96, 50]),
                                          ("SE", [40, 31, 25, 12, 50], [612, 400, 258, 64, 1000])])
def test_fused_layer_kernel_equals_the_two_launches(mode, tls, mls, monkeypatch):
    """flow_tc_layer.cuh (gate GEMM -> tanh * sigmoid kept in shared memory -> res/skip GEMM, one launch per WaveNet
    layer) against the two stand-alone GEMM launches (GLOW_FUSED_LAYER=0): the same MMAs in the same order on the
    same bf16 operands, so forward, reverse and every gradient agree to accumulation noise (1e-5 of the largest
    entry; the saved activations the backward reads are written by the same epilogue functor)."""
    monkeypatch.setenv("GLOW_FUSED_LAYER", "0")
    z2, ld2, g2, back2 = _fwd_bwd("bf16", mode, 77, tls, mls, 8)
    monkeypatch.setenv("GLOW_FUSED_LAYER", "1")
    z1, ld1, g1, back1 = _fwd_bwd("bf16", mode, 77, tls, mls, 8)
    assert torch.isfinite(z1).all()
    assert rel_err(z1, z2) < 1e-5, rel_err(z1, z2)
    assert rel_err(ld1, ld2) < 1e-5, rel_err(ld1, ld2)
    assert rel_err(back1, back2) < 1e-5, rel_err(back1, back2)
    worst = max((rel_err(g1[k], g2[k]), k) for k in g1 if float(g2[k].abs().max()) > 0)
    assert worst[0] < 1e-4, worst
