"""GPU: one captured CUDA graph per geometry BUCKET serves ragged batches (train.GraphedTrainStep, geometry.py).

The reference's loop never repeats a batch geometry (Train.py:582-584); the captured step therefore reads every
length-derived quantity from fixed device buffers.  Checked here: replays on batches the graph has never seen give
the losses / gradient norms / parameters of the plain eager step on the same data; cache evictions between replays
cannot hurt a graph (it keeps its workspaces alive); a second forward refuses to silently reuse a
workspace that a pending backward still needs."""
import math

import pytest
import torch

from tests._util import synth_batch, synth_state_dict

pytestmark = pytest.mark.gpu

SMALL = {"Decoder.Stack": 3, "Encoder.Transformer.Stacks": 2}


def _model(precision, mode="Vanilla", seed=5):
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    hp = load_hparams(Mode=mode, Precision=precision, **SMALL)
    modules.set_hparams(hp)
    model = modules.GlowTTS()
    model.load_state_dict(synth_state_dict(model.state_dict(), seed), strict=True)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = True
    return model.cuda().eval(), hp


def _pin(batch):
    return tuple(t.pin_memory() if torch.is_tensor(t) and t.dtype != torch.bool and i not in (1, 3) else t
                 for i, t in enumerate(batch))


# four batches of one bucket (B = 3; decoder rows <= 512, encoder rows <= 256) with different lengths / paddings
GEOS = [([23, 17, 9], [140, 96, 50]), ([30, 12, 21], [180, 70, 120]), ([11, 40, 25], [64, 250, 150]),
        ([23, 17, 9], [140, 96, 50])]


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 2e-2)])
def test_replays_on_unseen_geometries_match_the_eager_step(precision, tol):
    from glow_tts_b200.train import TrainStep, GraphedTrainStep
    batches = [synth_batch(50 + i, tls, mls) for i, (tls, mls) in enumerate(GEOS)]
    model_e, hp = _model(precision)
    eager = TrainStep(model_e, hp, torch.device("cuda:0"))
    want = []
    for b in batches:
        eager.run(eager.to_device(b))
        want.append({k: float(v) for k, v in eager.last.items()})
    model_g, hp = _model(precision)
    step = TrainStep(model_g, hp, torch.device("cuda:0"))
    graphed = GraphedTrainStep(step)
    keys = {graphed.bucket_key(b[1], b[3]) for b in batches}
    assert len(keys) == 1, keys                                  # one bucket: one capture, three replays
    got = []
    for b in batches:
        graphed.run(_pin(b))
        torch.cuda.synchronize()
        got.append({k: float(v) for k, v in graphed.last.items()})
    assert len(graphed.buckets) == 1 and graphed.current.replays == len(batches) - 1
    for i, (g, w) in enumerate(zip(got, want)):
        for k in ("mle", "mse", "grad_norm"):
            assert abs(g[k] - w[k]) <= tol * abs(w[k]), (i, k, g[k], w[k])
    pe = torch.cat([p.detach().flatten() for p in model_e.parameters()])
    pg = torch.cat([p.detach().flatten() for p in model_g.parameters()])
    assert float((pe - pg).abs().max()) <= tol * float(pe.abs().max())


def test_graph_survives_cache_evictions():
    """Eager steps on > 64 other geometries evict every size-bounded cache (row maps, token rows, device ints,
    workspaces); the captured graph holds strong references to what it touched and replays correctly afterwards."""
    from glow_tts_b200.train import TrainStep, GraphedTrainStep
    model, hp = _model("bf16")
    step = TrainStep(model, hp, torch.device("cuda:0"))
    step.opt.lr0 = 0.0
    step.opt.wd = 0.0                                             # frozen weights: the same batch gives the same loss
    graphed = GraphedTrainStep(step)
    batch = _pin(synth_batch(7, *GEOS[0]))
    graphed.run(batch)
    graphed.run(batch)
    torch.cuda.synchronize()
    ref = float(graphed.last["mle"])
    for i in range(70):                                           # distinct geometries through the EAGER path
        tls, mls = [10 + (i % 13), 12 + i // 7], [60 + 2 * i, 100 + 2 * (i % 11)]
        loss = step.run(step.to_device(synth_batch(200 + i, tls, mls)))
    assert math.isfinite(float(loss))
    junk = [torch.randn(1 << 22, device="cuda") for _ in range(16)]   # reuse whatever the evictions freed
    del junk
    graphed.run(batch)
    torch.cuda.synchronize()
    assert abs(float(graphed.last["mle"]) - ref) <= 1e-5 * abs(ref)


def test_second_forward_before_backward_is_refused():
    """One workspace per shape holds the saved activations: a second grad-enabled forward of the same
    shape would silently corrupt the first one's backward -- it raises instead."""
    from glow_tts_b200 import _lib
    from tests._model_util import mel_mask
    model, hp = _model("bf16")
    tokens, tl, mels, ml, spk = synth_batch(3, [20, 12], [100, 64])
    dec = model.layer_Dict["Decoder"]
    x = mels.cuda().requires_grad_(True)
    z1, ld1, _ = dec(x, mel_mask(ml, "cuda"), None)
    z2, ld2, _ = dec(x, mel_mask(ml, "cuda"), None)
    z2.sum().backward()                                            # the latest forward owns the workspace
    with pytest.raises(_lib.GlowCoreError):
        z1.sum().backward()


def test_optimizer_state_round_trip_resumes_the_trajectory():
    """FusedRAdam.state_dict -> a fresh TrainStep.load_state_dict: step 3 after a resume equals step 3 of the
    uninterrupted run (moments, rectification step count and Noam position restored; Train.py:514-519)."""
    from glow_tts_b200.train import TrainStep
    batch = synth_batch(9, [23, 17, 9], [140, 96, 50])
    model_a, hp = _model("fp32")
    a = TrainStep(model_a, hp, torch.device("cuda:0"))
    for _ in range(3):
        a.run(a.to_device(batch))
    want = {k: float(v) for k, v in a.last.items()}
    model_b, hp = _model("fp32")
    b = TrainStep(model_b, hp, torch.device("cuda:0"))
    for _ in range(2):
        b.run(b.to_device(batch))
    ckpt = {"Model": {k: v.clone() for k, v in model_b.state_dict().items()}, "Optimizer": b.opt.state_dict(),
            "Scheduler": b.opt.scheduler_state_dict()}
    model_c, hp = _model("fp32", seed=77)                          # different weights until the checkpoint is loaded
    model_c.load_state_dict(ckpt["Model"], strict=True)
    c = TrainStep(model_c, hp, torch.device("cuda:0"))
    c.opt.load_state_dict(ckpt["Optimizer"], ckpt["Scheduler"])
    c.run(c.to_device(batch))
    got = {k: float(v) for k, v in c.last.items()}
    for k in ("mle", "mse", "grad_norm"):
        assert abs(got[k] - want[k]) <= 1e-5 * abs(want[k]), (k, got[k], want[k])
    pa = torch.cat([p.detach().flatten() for p in model_a.parameters()])
    pc = torch.cat([p.detach().flatten() for p in model_c.parameters()])
    assert float((pa - pc).abs().max()) <= 1e-6 * float(pa.abs().max())
