"""CPU: pin oracle/glow_oracle.py (the torch-fp32 restatement of the reference's flow
decoder / attention encoder / glue / loss / train step) against the fixtures produced
by running the real reference modules (tests/golden/model_*.npz, tools/make_golden_model.py),
and check the drop-in modules expose the reference's exact state_dict layout."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import glow_oracle as O
from tests._util import GOLD, checksum, rel_err, synth_batch, synth_state_dict

CASES = {"vanilla_small": ("Vanilla", [23, 17, 9], [140, 96, 50], 21, 1234),
         "se_small": ("SE", [19, 12], [110, 64], 22, 4321)}
TOL = 2e-4      # fp32 vs fp32, different op order (direct banded sums vs pad/view skewing)


def _keys(mode):
    return json.load(open(os.path.join(GOLD, "state_dict_keys_%s.json" % mode.lower())))


def _state_dict(mode, seed):
    shapes = {k: torch.empty(s) for k, s in _keys(mode)}
    return synth_state_dict(shapes, seed)


def _digest(t, seed):
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(t.shape, generator=g)
    return [float(t.double().norm()), float((t.double() * r.double()).sum())]


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    mode, tls, mls, bseed, wseed = CASES[request.param]
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % request.param))
    sd = _state_dict(mode, wseed)
    assert checksum(torch.cat([sd[k].flatten() for k in sorted(sd)]).numpy()) == str(gold["weights_sha"])
    return dict(mode=mode, hp=O.OracleHP(mode=mode), gold=gold, sd=sd, batch=synth_batch(bseed, tls, mls))


@pytest.mark.parametrize("mode", ["Vanilla", "SE"])
def test_dropin_state_dict_layout_matches_reference(mode):
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    modules.set_hparams(load_hparams(Mode=mode))
    model = modules.GlowTTS()
    mine = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    assert mine == _keys(mode)                   # same keys, same shapes, same order
    names = [k for k, _ in model.named_parameters()]
    assert names == [k for k, _ in _keys(mode)]  # reference has no buffers: parameter order == key order


def test_decoder_forward_reverse(case):
    hp, sd, g = case["hp"], case["sd"], case["gold"]
    tokens, tl, mels, ml, spk = case["batch"]
    emb = sd["layer_Dict.LUT.weight"][spk] if hp.se else None
    with torch.no_grad():
        z, ld, _ = O.decoder(sd, mels, O.length_mask(ml), hp, emb)
        back, _, _ = O.decoder(sd, z, O.length_mask(ml), hp, emb, reverse=True)
    assert rel_err(z, g["dec_z"]) < TOL and rel_err(ld, g["dec_logdet"]) < TOL
    # the inverse amplifies fp32 rounding (exp(-logs), W^-1 over 12 blocks): two fp32 evaluations
    # with different op order agree to ~5e-4 here, so the reverse direction gets 2e-3
    assert rel_err(back, g["dec_reverse_of_z"]) < 2e-3


def test_decoder_gradients(case):
    hp, g = case["hp"], case["gold"]
    sd = O.state_dict_to_leaves(case["sd"])
    tokens, tl, mels, ml, spk = case["batch"]
    emb = sd["layer_Dict.LUT.weight"][spk].detach() if hp.se else None
    z, ld, _ = O.decoder(sd, mels, O.length_mask(ml), hp, emb)
    gen = torch.Generator().manual_seed(99)
    rz, rl = torch.randn(z.shape, generator=gen), torch.randn(ld.shape, generator=gen)
    ((z * rz).sum() + (ld * rl).sum()).backward()
    pre = "layer_Dict.Decoder."
    for key, want in zip(g["dec_grad_keys"], g["dec_grad_digest"]):
        got = _digest(sd[pre + str(key)].grad, 7)
        assert abs(got[0] - want[0]) <= 2e-3 * max(want[0], 1e-6), key
    assert rel_err(sd[pre + "layer_Dict.Flows.0.layers.1.weight"].grad, g["dec_grad_b0_w"]) < 2e-3
    assert rel_err(sd[pre + "layer_Dict.Flows.0.layers.0.logs"].grad, g["dec_grad_b0_logs"]) < 2e-3


def test_attention_forward_backward(case):
    hp, g = case["hp"], case["gold"]
    sd = O.state_dict_to_leaves(case["sd"])
    tokens, tl, mels, ml, spk = case["batch"]
    p = "layer_Dict.Encoder.layer_Dict.Transformer.layer_Dict.ANCRDCN_0.layer_Dict.Attention"
    x = torch.from_numpy(g["att_x"]).requires_grad_(True)
    tm = O.length_mask(tl)
    out, align = O.rpr_attention(sd, p, x, (tm * tm.transpose(1, 2)).unsqueeze(1), hp)
    assert rel_err(out, g["att_out"]) < TOL and rel_err(align, g["att_align"]) < TOL
    gen = torch.Generator().manual_seed(5)
    torch.randn(x.shape, generator=gen)
    (out * torch.randn(out.shape, generator=gen)).sum().backward()
    assert rel_err(x.grad, g["att_dx"]) < 1e-3
    assert rel_err(sd[p + ".weight_K"].grad, g["att_dwk"]) < 1e-3
    assert rel_err(sd[p + ".weight_V"].grad, g["att_dwv"]) < 1e-3


def test_encoder_and_full_forward(case):
    hp, sd, g = case["hp"], case["sd"], case["gold"]
    tokens, tl, mels, ml, spk = case["batch"]
    emb = sd["layer_Dict.LUT.weight"][spk] if hp.se else None
    with torch.no_grad():
        mean, log_std, logw, _ = O.encoder(sd, tokens, O.length_mask(tl), hp, emb)
    assert rel_err(mean, g["enc_mean"]) < TOL and rel_err(log_std, g["enc_log_std"]) < TOL
    assert rel_err(logw, g["enc_logw"]) < TOL
    leaves = O.state_dict_to_leaves(sd)
    out = O.glow_forward(leaves, hp, tokens, tl, mels, ml, spk)
    assert np.array_equal(out[6].argmax(1).numpy().astype(np.int16), g["fw_attn_pos"])
    for got, key in zip(out[:6], ["fw_z", "fw_mel_mean", "fw_mel_log_std", "fw_logdet", "fw_logw", "fw_logw_target"]):
        assert rel_err(got, g[key]) < TOL, key
    total, mle, mse = O.losses(out, ml, hp)
    assert abs(float(mle) - g["fw_losses"][0]) < 1e-3 * abs(g["fw_losses"][0])
    assert abs(float(mse) - g["fw_losses"][1]) < 1e-3 * abs(g["fw_losses"][1])
    total.backward()
    floor = 1e-6 * float(g["fw_grad_digest"][:, 0].max())     # Key.bias grads are analytically 0 (softmax shift)
    for key, want in zip(g["fw_grad_keys"], g["fw_grad_digest"]):
        got = _digest(leaves[str(key)].grad, 11)
        assert abs(got[0] - want[0]) <= 3e-3 * want[0] + floor, key


def test_train_steps_radam_noam(case):
    hp, g = case["hp"], case["gold"]
    leaves = O.state_dict_to_leaves(case["sd"])
    keys = [k for k, _ in _keys(case["mode"])]
    opt = O.RAdamOracle([leaves[k] for k in keys])
    for step in range(3):
        total, mle, mse = O.train_step(leaves, hp, opt, case["batch"], training=False)
        assert abs(mle - g["train_losses"][step][0]) < 2e-3 * abs(g["train_losses"][step][0]), step
        assert abs(mse - g["train_losses"][step][1]) < 2e-3 * abs(g["train_losses"][step][1]), step
    flat = torch.cat([leaves[k].detach().flatten() for k in keys])
    want = g["train_param_digest"]
    assert abs(float(flat.double().norm()) - want[0]) < 1e-4 * want[0]


# --------------------------------------------------------------------------- larger cases / data-dependent init
LARGE = {"vanilla_large": ("Vanilla", [202, 160, 131, 99, 85, 66, 40, 19], [1000, 946, 812, 640, 518, 402, 256, 104], 31, 1234),
         "se_large": ("SE", [130, 100, 64, 55, 47, 33, 21, 12], [800, 620, 404, 350, 280, 222, 140, 50], 32, 4321)}
DDI = {"ddi_vanilla": ("Vanilla", [40, 31, 18, 25], [300, 222, 96, 164], 41, 1234),
       "ddi_se": ("SE", [33, 21], [200, 128], 42, 4321)}


@pytest.mark.parametrize("name", list(LARGE))
def test_full_forward_at_baseline_sizes(name):
    """B = 8, T_mel up to 1000, T_text up to 202 (BASELINE.json configs[1] / [2] shapes): the oracle's forward,
    alignment and losses against the reference-made fixture (big tensors are stored every 8th frame + a digest)."""
    mode, tls, mls, bseed, wseed = LARGE[name]
    g = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    sd = _state_dict(mode, wseed)
    assert checksum(torch.cat([sd[k].flatten() for k in sorted(sd)]).numpy()) == str(g["weights_sha"])
    hp = O.OracleHP(mode=mode)
    tokens, tl, mels, ml, spk = synth_batch(bseed, tls, mls)
    stride = int(g["stride"])
    with torch.no_grad():
        out = O.glow_forward(sd, hp, tokens, tl, mels, ml, spk)
    assert np.array_equal(out[6].argmax(1).numpy().astype(np.int16), g["fw_attn_pos"])
    for got, key in zip(out[:3], ["fw_z", "fw_mel_mean", "fw_mel_log_std"]):
        assert rel_err(got[..., ::stride], g[key]) < TOL, key
        d = _digest(got, 13)
        assert abs(d[0] - g[key + "_digest"][0]) < TOL * g[key + "_digest"][0], key
    for got, key in zip(out[3:6], ["fw_logdet", "fw_logw", "fw_logw_target"]):
        assert rel_err(got, g[key]) < TOL, key
    total, mle, mse = O.losses(out, ml, hp)
    assert abs(float(mle) - g["fw_losses"][0]) < 1e-4 * abs(g["fw_losses"][0])
    assert abs(float(mse) - g["fw_losses"][1]) < 1e-4 * abs(g["fw_losses"][1])


@pytest.mark.parametrize("name", list(DDI))
def test_actnorm_data_dependent_init(name):
    """Activation_Norm.initialize (Modules.py:698-711) through all 12 blocks: logs / bias of every block and the
    resulting z / logdet against the reference run on an uninitialised model."""
    mode, tls, mls, bseed, wseed = DDI[name]
    g = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    sd = _state_dict(mode, wseed)
    hp = O.OracleHP(mode=mode)
    tokens, tl, mels, ml, spk = synth_batch(bseed, tls, mls)
    emb = sd["layer_Dict.LUT.weight"][spk] if hp.se else None
    z, ld, logs, bias = O.decoder_ddi(sd, mels, O.length_mask(ml), hp, emb)
    assert rel_err(logs, g["logs"]) < 1e-3 and rel_err(bias, g["bias"]) < 1e-3
    assert rel_err(z, g["z"]) < 1e-3 and rel_err(ld, g["logdet"]) < 1e-3
