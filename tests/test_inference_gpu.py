"""GPU tests of the sync-free serving path: GlowTTS.inference_device / infer.GraphedInference
against GlowTTS.inference (itself pinned to the reference's fixtures in test_model_gpu.py).  Same arithmetic on a
different row geometry: identical up to fp32 rounding (1e-5 of the largest mel value; bf16 mode 2e-2)."""
import numpy as np
import pytest
import torch

from tests._model_util import load_case

pytestmark = pytest.mark.gpu


def _tokens(case_batch):
    tokens, tl, mels, ml, spk = case_batch
    return tokens.cuda(), tl, (spk.cuda() if spk is not None else None)


@pytest.mark.parametrize("name,precision,tol", [("vanilla_small", "fp32", 1e-5), ("se_small", "fp32", 1e-5),
                                                ("vanilla_small", "bf16", 2e-2)])
def test_inference_device_matches_inference(name, precision, tol):
    model, sd, g, batch, mode = load_case(name, precision)
    model.eval()
    tokens, tl, spk = _tokens(batch)
    t_mel = 400
    noises = torch.randn(tokens.shape[0], 80, t_mel, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    want, want_len, want_att = model.inference(tokens=tokens, token_lengths=tl, speakers=spk, noise_scale=0.6,
                                               length_scale=1.0, noises=noises)
    got, got_len, got_att = model.inference_device(tokens, tl.cuda(), speakers=spk, noise_scale=0.6, length_scale=1.0,
                                                   max_mel_length=t_mel, noises=noises)
    assert torch.equal(got_len.cpu(), want_len.cpu())
    n = want.shape[2] // 2 * 2                                  # the decoder drops an odd tail frame
    assert got.shape[2] == t_mel and n <= t_mel
    scale = want.abs().max().item()
    assert (got[:, :, :n] - want[:, :, :n]).abs().max().item() < tol * scale
    assert torch.equal(got_att[:, :, :want_att.shape[2]], want_att)
    # beyond every utterance's length: the fill value (Modules.py:202)
    for b in range(got.shape[0]):
        m = int(want_len[b]) // 2 * 2
        assert torch.all(got[b, :, m:] == -4.0)


def test_graphed_inference_replays_follow_their_inputs():
    from glow_tts_b200.infer import GraphedInference
    model, sd, g, batch, mode = load_case("vanilla_small", "bf16")
    model.eval()
    tokens, tl, _ = _tokens(batch)
    gi = GraphedInference(model, batch=4, max_text_length=32, max_mel_length=400, noise_scale=0.0)
    assert gi.launches_per_replay > 0
    mels, lens = gi.run(tokens.cpu(), tl)
    torch.cuda.synchronize()
    want, want_len, _ = model.inference(tokens=tokens, token_lengths=tl, noise_scale=0.0, length_scale=1.0)
    b = tokens.shape[0]
    assert torch.equal(lens[:b].cpu(), want_len.cpu())
    n = want.shape[2] // 2 * 2
    scale = want.abs().max().item()
    # bf16 decoder after a torch encoder whose library kernels may differ with the batch size (4 slots vs 3): bf16 tolerance
    assert (mels[:b, :, :n] - want[:, :, :n]).abs().max().item() < 2e-2 * scale
    first = mels[:b].clone()
    # a different request through the same graph: shorter sentences, reversed order
    tokens2 = torch.flip(tokens, dims=[0]).cpu()[:, :12].contiguous()
    tl2 = torch.clamp(torch.flip(tl, dims=[0]), max=12)
    tokens2[torch.arange(12)[None, :] >= tl2[:, None]] = 1
    mels2, lens2 = gi.run(tokens2, tl2)
    torch.cuda.synchronize()
    want2, want_len2, _ = model.inference(tokens=tokens2.cuda(), token_lengths=tl2, noise_scale=0.0, length_scale=1.0)
    assert torch.equal(lens2[:b].cpu(), want_len2.cpu())
    n2 = want2.shape[2] // 2 * 2
    assert (mels2[:b, :, :n2] - want2[:, :, :n2]).abs().max().item() < 2e-2 * scale
    assert not torch.equal(mels2[:b], first)
