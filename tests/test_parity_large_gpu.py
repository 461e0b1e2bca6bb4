"""GPU: parity of the BENCHMARKED configuration against reference-made fixtures at BASELINE sizes.

tests/golden/model_{vanilla,se}_large.npz come from running the reference's own modules
(tools/make_golden_model.py: B = 8, T_mel up to 1000, T_text up to 202; Vanilla = BASELINE configs[1]
shapes, SE = configs[2] shapes).  Every precision mode the product ships is run through

    full forward -> MAS -> MLE / MSE losses -> backward (gradient digests) -> 3 optimizer steps

and compared with STATED tolerances (the TOL table below; metric = max|a-b| / max|b| unless noted):

* ``fp32``     CUDA-core fp32 GEMMs + torch encoder: the north_star's 1e-3.
* ``fp32-tc``  the 1e-3 tensor-core mode: tcgen05 with every operand split into bf16 hi + lo parts and three
               MMAs per product (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), fp32 activations, exact
               tanhf / expf epilogues.
* ``bf16``     the headline bench mode (bf16 activations / weights, tcgen05, fast tanh / exp epilogues): bf16
               has eps = 3.9e-3 per rounding, so it cannot meet 1e-3 against an fp32 reference; its tolerances are
               what was measured on B200 with ~2x margin, and the fraction of mel frames whose MAS token differs
               from the reference path is bounded separately.

The measured errors of the last run are written to gpurun_out/parity_large.json (DESIGN.md section 2 quotes them).
"""
import json
import os

import numpy as np
import pytest
import torch

from tests._model_util import build_model, digest
from tests._util import GOLD, REPO, checksum, rel_err, synth_batch

pytestmark = pytest.mark.gpu

LARGE = {"vanilla_large": ("Vanilla", [202, 160, 131, 99, 85, 66, 40, 19], [1000, 946, 812, 640, 518, 402, 256, 104], 31, 1234),
         "se_large": ("SE", [130, 100, 64, 55, 47, 33, 21, 12], [800, 620, 404, 350, 280, 222, 140, 50], 32, 4321)}

# per precision: tolerance of each compared quantity
TOL = {
    "fp32": dict(z=1e-3, mel_mean=1e-3, mel_log_std=1e-3, logdet=1e-3, logw=1e-3, logw_target=1e-3, mle=1e-3, mse=1e-3,
                 mas_diff=0.0, grad_norm=3e-3, grad_total=3e-3, train_mle=2e-3, train_mse=2e-3, train_gn=5e-3, train_param_norm=1e-4),
    "fp32-tc": dict(z=1e-3, mel_mean=1e-3, mel_log_std=1e-3, logdet=1e-3, logw=1e-3, logw_target=1e-3, mle=1e-3, mse=1e-3,
                    mas_diff=0.0, grad_norm=3e-3, grad_total=3e-3, train_mle=2e-3, train_mse=2e-3, train_gn=5e-3, train_param_norm=1e-4),
    # bf16: z / logdet carry 12 blocks x 10 GEMMs of bf16 rounding; mel_mean / mel_log_std / logw_target differ where
    # the alignment differs (mas_diff = fraction of real mel frames whose token differs from the reference path) and
    # are only recorded.  Measured on B200 (profiles/parity_r02a_large.json; vanilla / SE): z 6.9e-3 / 1.5e-2, logdet
    # 2.2e-4 / 6.9e-5, MLE 3.7e-3 / 6.8e-3, mas_diff 2.1e-4 / 1.2e-2, whole-gradient norm 5.3e-3 / 2.8e-2, worst single
    # parameter's gradient digest 6.4e-2 / 1.8e-1 (small tensors: ActNorm bias, biases behind a ReLU).
    "bf16": dict(z=3e-2, mel_mean=None, mel_log_std=None, logdet=2e-3, logw=2e-2, logw_target=None, mle=1.5e-2, mse=5e-2,
                 mas_diff=0.03, grad_norm=3e-1, grad_total=6e-2, train_mle=2e-2, train_mse=6e-2, train_gn=6e-2, train_param_norm=1e-6),
}
MEASURED = {}


def _modes():
    from glow_tts_b200 import flow
    modes = ["fp32", "bf16"]
    try:
        flow.precision_tag("fp32-tc")
        modes.insert(1, "fp32-tc")
    except ValueError:
        pass
    return modes


def _load(name, precision):
    mode, tls, mls, bseed, wseed = LARGE[name]
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    model, sd = build_model(mode, wseed, precision)
    assert checksum(torch.cat([sd[k].flatten() for k in sorted(sd)]).numpy()) == str(gold["weights_sha"])
    return model, gold, synth_batch(bseed, tls, mls), mode


def _record(name, precision, key, value):
    MEASURED.setdefault("%s/%s" % (name, precision), {})[key] = float(value)
    out = os.path.join(REPO, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_large.json"), "w") as f:
        json.dump(MEASURED, f, indent=1, sort_keys=True)


FAILS = []


def _check(name, precision, key, err):
    """Record first, judge at the end of the test (_verdict): one run reports every quantity."""
    _record(name, precision, key, err)
    tol = TOL[precision][key]
    if tol is not None and not err <= tol:
        FAILS.append("%s %s %s: %.3e > %.1e" % (name, precision, key, err, tol))


def _verdict():
    msgs = list(FAILS)
    del FAILS[:]
    assert not msgs, "; ".join(msgs)


@pytest.mark.parametrize("precision", _modes())
@pytest.mark.parametrize("name", list(LARGE))
def test_forward_losses_gradients_at_baseline_sizes(name, precision):
    from glow_tts_b200 import modules
    model, g, batch, mode = _load(name, precision)
    model.eval()                                  # dropout off, as in the fixture
    tokens, tl, mels, ml, spk = (t.cuda() for t in batch)
    stride = int(g["stride"])
    model.zero_grad(set_to_none=True)
    out = model(tokens=tokens, token_lengths=tl, mels=mels, mel_lengths=ml, speakers=spk, mels_for_ge2e=None, pitches=None)
    z, mm, mls, ld, lw, lwt, attn = out[:7]
    # alignment: fraction of real mel frames whose token differs from the reference's path
    pos = attn.argmax(1).cpu().numpy().astype(np.int16)
    want_pos = g["fw_attn_pos"]
    valid = np.arange(pos.shape[1])[None, :] < batch[3].numpy()[:, None]
    diff = float(((pos != want_pos) & valid).sum()) / float(valid.sum())
    _check(name, precision, "mas_diff", diff)
    same = diff == 0.0
    for got, key in ((z, "z"), (mm, "mel_mean"), (mls, "mel_log_std")):
        err = rel_err(got.detach().cpu()[..., ::stride], g["fw_" + key])
        if key == "z" or same or TOL[precision][key] is not None:
            _check(name, precision, key, err)
        else:
            _record(name, precision, key, err)
    for got, key in ((ld, "logdet"), (lw, "logw"), (lwt, "logw_target")):
        err = rel_err(got.detach().cpu(), g["fw_" + key])
        if key != "logw_target" or same or TOL[precision][key] is not None:
            _check(name, precision, key, err)
        else:
            _record(name, precision, key, err)
    mle = modules.MLE_Loss()(z=z, mean=mm, std=mls, log_dets=ld, lengths=ml)
    mse = torch.nn.MSELoss()(lw, lwt)
    _check(name, precision, "mle", abs(float(mle) - g["fw_losses"][0]) / abs(g["fw_losses"][0]))
    _check(name, precision, "mse", abs(float(mse) - g["fw_losses"][1]) / abs(g["fw_losses"][1]))
    (mle + mse).backward()
    params = dict(model.named_parameters())
    scale = float(g["fw_grad_digest"][:, 0].max())
    worst = 0.0
    for key, want in zip(g["fw_grad_keys"], g["fw_grad_digest"]):
        got = digest(params[str(key)].grad, 11)
        # relative to the gradient's own norm, with a floor for analytically-zero gradients (Key.bias: softmax shift)
        worst = max(worst, abs(got[0] - want[0]) / (want[0] + 1e-4 * scale), abs(got[1] - want[1]) / (want[0] + 1e-4 * scale))
    _check(name, precision, "grad_norm", worst)
    total = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).double().norm().item()
    _check(name, precision, "grad_total", abs(total - float(g["fw_grad_total_norm"])) / float(g["fw_grad_total_norm"]))
    _verdict()


@pytest.mark.parametrize("precision", [m for m in _modes() if m != "fp32"])
@pytest.mark.parametrize("name", list(LARGE))
def test_three_train_steps_at_baseline_sizes(name, precision):
    """Loss / gradient-norm trajectory of three optimizer steps (clip 5.0, RAdam, Noam) against the reference's."""
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep
    model, g, batch, mode = _load(name, precision)
    model.eval()
    step = TrainStep(model, load_hparams(Mode=mode, Precision=precision), torch.device("cuda:0"))
    dev_batch = step.to_device(batch)
    worst = dict(train_mle=0.0, train_mse=0.0, train_gn=0.0)
    for i in range(3):
        step.run(dev_batch)
        want = g["train_losses"][i]
        worst["train_mle"] = max(worst["train_mle"], abs(float(step.last["mle"]) - want[0]) / abs(want[0]))
        worst["train_mse"] = max(worst["train_mse"], abs(float(step.last["mse"]) - want[1]) / abs(want[1]))
        worst["train_gn"] = max(worst["train_gn"], abs(float(step.last["grad_norm"]) - want[2]) / abs(want[2]))
    for k, v in worst.items():
        _check(name, precision, k, v)
    flat = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
    _check(name, precision, "train_param_norm", abs(float(flat.double().norm()) - g["train_param_digest"][0]) / g["train_param_digest"][0])
    _verdict()
