"""Helpers shared by the GPU model tests."""
import json
import os

import numpy as np
import torch

from tests._util import GOLD, checksum, synth_batch, synth_state_dict

CASES = {"vanilla_small": ("Vanilla", [23, 17, 9], [140, 96, 50], 21, 1234),
         "se_small": ("SE", [19, 12], [110, 64], 22, 4321)}


def build_model(mode, seed, precision="fp32", device="cuda:0"):
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    modules.set_hparams(load_hparams(Mode=mode, Precision=precision))
    model = modules.GlowTTS()
    sd = synth_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd, strict=True)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = True
    return model.to(device), sd


def load_case(name, precision="fp32"):
    mode, tls, mls, bseed, wseed = CASES[name]
    gold = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    model, sd = build_model(mode, wseed, precision)
    assert checksum(torch.cat([sd[k].flatten() for k in sorted(sd)]).numpy()) == str(gold["weights_sha"])
    return model, sd, gold, synth_batch(bseed, tls, mls), mode


def digest(t, seed):
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(t.shape, generator=g)
    t = t.detach().cpu()
    return [float(t.double().norm()), float((t.double() * r.double()).sum())]


def mel_mask(ml, device):
    n = int(max(ml))
    return (torch.arange(n, device=device)[None, :] < torch.as_tensor(ml, device=device)[:, None]).unsqueeze(1).float()
