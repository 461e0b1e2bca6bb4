"""GPU: the whole drop-in GlowTTS (encoder + flow decoder + MAS + glue) and the train
step (loss, backward, clip, fused RAdam, Noam) vs fixtures made by running the reference
(tools/make_golden_model.py).  fp32 mode, tolerance 1e-3 relative; alignments exact."""
import numpy as np
import pytest
import torch

from tests._model_util import CASES, digest, load_case
from tests._util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    return load_case(request.param, "fp32")


def _dev(batch):
    tokens, tl, mels, ml, spk = batch
    return tokens.cuda(), tl.cuda(), mels.cuda(), ml.cuda(), spk.cuda()


def test_encoder_matches_reference(case):
    model, sd, g, batch, mode = case
    model.eval()
    tokens, tl, mels, ml, spk = _dev(batch)
    emb = model.layer_Dict["LUT"](spk) if mode == "SE" else None
    tmask = model.Mask_Generate(tl)
    with torch.no_grad():
        mean, log_std, logw, _ = model.layer_Dict["Encoder"](tokens, tmask, emb, None)
    assert rel_err(mean.cpu(), g["enc_mean"]) < 1e-3
    assert rel_err(log_std.cpu(), g["enc_log_std"]) < 1e-3
    assert rel_err(logw.cpu(), g["enc_logw"]) < 1e-3


def test_full_forward_losses_and_gradients(case):
    from glow_tts_b200 import modules
    model, sd, g, batch, mode = case
    model.eval()
    tokens, tl, mels, ml, spk = _dev(batch)
    model.zero_grad(set_to_none=True)
    out = model(tokens=tokens, token_lengths=tl, mels=mels, mel_lengths=ml, speakers=spk,
                mels_for_ge2e=None, pitches=None)
    assert len(out) == 8 and out[7] is None
    z, mm, mls, ld, lw, lwt, attn = out[:7]
    assert np.array_equal(attn.argmax(1).cpu().numpy().astype(np.int16), g["fw_attn_pos"])      # MAS: exact
    for got, key in zip((z, mm, mls, ld, lw, lwt), ["fw_z", "fw_mel_mean", "fw_mel_log_std", "fw_logdet", "fw_logw", "fw_logw_target"]):
        assert rel_err(got.cpu(), g[key]) < 1e-3, key
    mle = modules.MLE_Loss()(z=z, mean=mm, std=mls, log_dets=ld, lengths=ml)
    mse = torch.nn.MSELoss()(lw, lwt)
    assert abs(float(mle) - g["fw_losses"][0]) < 1e-3 * abs(g["fw_losses"][0])
    assert abs(float(mse) - g["fw_losses"][1]) < 1e-3 * abs(g["fw_losses"][1])
    (mle + mse).backward()
    params = dict(model.named_parameters())
    floor = 1e-6 * float(g["fw_grad_digest"][:, 0].max())
    for key, want in zip(g["fw_grad_keys"], g["fw_grad_digest"]):
        got = digest(params[str(key)].grad, 11)
        assert abs(got[0] - want[0]) <= 3e-3 * want[0] + floor, key
        assert abs(got[1] - want[1]) <= 3e-3 * want[0] + floor, key


def test_inference_matches_reference(case):
    model, sd, g, batch, mode = case
    model.eval()
    tokens, tl, mels, ml, spk = _dev(batch)
    mels_out, lengths, attn = model.inference(tokens=tokens, token_lengths=tl, speakers=spk, noise_scale=0.0,
                                              length_scale=torch.ones(len(tl), device="cuda"))
    assert np.array_equal(lengths.cpu().numpy(), g["inf_lengths"])
    want = torch.from_numpy(g["inf_mels"])
    got = mels_out.cpu()
    # the reference decodes T_max (possibly odd) frames and drops the odd tail inside Squeeze
    n = min(got.shape[2], want.shape[2])
    assert rel_err(got[:, :, :n], want[:, :, :n]) < 3e-3


def test_three_train_steps_match_reference(case):
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep
    name = [k for k, v in CASES.items() if v[0] == case[4]][0]
    model, sd, g, batch, mode = load_case(name, "fp32")      # fresh weights
    model.eval()                                              # dropout off, as in the fixture
    hp = load_hparams(Mode=mode, Precision="fp32")
    step = TrainStep(model, hp, torch.device("cuda:0"))
    dev_batch = step.to_device(batch)
    for i in range(3):
        step.run(dev_batch)
        want = g["train_losses"][i]
        assert abs(float(step.last["mle"]) - want[0]) < 2e-3 * abs(want[0]), i
        assert abs(float(step.last["mse"]) - want[1]) < 2e-3 * abs(want[1]), i
        assert abs(float(step.last["grad_norm"]) - want[2]) < 5e-3 * abs(want[2]), i
    flat = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
    assert abs(float(flat.double().norm()) - g["train_param_digest"][0]) < 1e-4 * g["train_param_digest"][0]


def test_graphed_train_steps_match_reference(case):
    """The same three steps with steps 2 and 3 replayed from the captured CUDA graph."""
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep, GraphedTrainStep
    name = [k for k, v in CASES.items() if v[0] == case[4]][0]
    model, sd, g, batch, mode = load_case(name, "fp32")
    model.eval()
    hp = load_hparams(Mode=mode, Precision="fp32")
    step = TrainStep(model, hp, torch.device("cuda:0"))
    graphed = GraphedTrainStep(step, batch, warmup=1)         # step 1 runs eagerly as the warm-up
    lasts = [dict((k, float(v)) for k, v in graphed.warmup_last.items())]
    for i in (1, 2):
        graphed.run(batch if i == 1 else None)                 # once through load(), once on resident inputs
        torch.cuda.synchronize()
        lasts.append(dict((k, float(v)) for k, v in graphed.last.items()))
    for i, got in enumerate(lasts):
        want = g["train_losses"][i]
        assert abs(got["mle"] - want[0]) < 2e-3 * abs(want[0]), i
        assert abs(got["mse"] - want[1]) < 2e-3 * abs(want[1]), i
        assert abs(got["grad_norm"] - want[2]) < 5e-3 * abs(want[2]), i
    flat = torch.cat([p.detach().flatten() for p in model.parameters()]).cpu()
    assert abs(float(flat.double().norm()) - g["train_param_digest"][0]) < 1e-4 * g["train_param_digest"][0]


def test_prefetched_batches_reach_the_graph():
    """GraphedTrainStep.prefetch stages a later step's batch on a copy stream; run() with the same object must
    train on exactly that data: same losses as the plain load() path, and different data gives different losses."""
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep, GraphedTrainStep
    from tests._util import synth_batch
    mode, tls, mls, bseed, wseed = CASES["vanilla_small"]
    other = synth_batch(bseed + 7, tls, mls)                   # same geometry, different tokens and mels
    runs = {}
    for how in ("load", "prefetch"):
        model, sd, g, batch, _ = load_case("vanilla_small", "fp32")
        model.eval()
        step = TrainStep(model, load_hparams(Mode=mode, Precision="fp32"), torch.device("cuda:0"))
        graphed = GraphedTrainStep(step, batch, warmup=1)
        pinned = [tuple(t.pin_memory() if torch.is_tensor(t) and not t.is_cuda else t for t in b) for b in (other, batch)]
        losses = []
        if how == "prefetch":
            graphed.prefetch(pinned[0])
        for i in range(4):
            cur = pinned[i % 2]
            graphed.run(cur)
            if how == "prefetch":
                graphed.prefetch(pinned[(i + 1) % 2])
            losses.append(float(graphed.last["mle"]))
        runs[how] = losses
    for a, b in zip(runs["load"], runs["prefetch"]):
        assert abs(a - b) < 1e-5 * abs(a), runs          # same data path up to atomic-order rounding
    assert abs(runs["load"][0] - runs["load"][1]) > 1e-4, runs


def test_per_block_parameter_gradients_equal_the_single_pass(monkeypatch):
    """glow_flow_backward_params (parameter gradients per block inside the backward call) against the default
    glow_flow_backward + glow_flow_param_grads: same flat gradient buffer after one step's backward."""
    from glow_tts_b200 import flow
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep
    grads = {}
    for fused in (False, True):
        monkeypatch.setattr(flow, "FUSED_PARAM_GRADS", fused)
        model, sd, g, batch, mode = load_case("se_small", "fp32")
        model.eval()
        step = TrainStep(model, load_hparams(Mode=mode, Precision="fp32"), torch.device("cuda:0"))
        step.opt.lr0 = 0.0
        step.run(step.to_device(batch))
        torch.cuda.synchronize()
        grads[fused] = step.flat.grad.detach().clone()
    scale = grads[False].abs().max().item()
    assert scale > 0
    assert (grads[True] - grads[False]).abs().max().item() < 1e-5 * scale


def test_graph_replays_draw_fresh_dropout_masks():
    """Two replays of one captured training step must not reuse the dropout masks: the kernels mix the
    device step counter into their seeds (glow_flow_call.step_dev / glow_attn_call.step_dev)."""
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep, GraphedTrainStep
    model, sd, g, batch, mode = load_case("vanilla_small", "fp32")
    model.train()
    hp = load_hparams(Mode=mode, Precision="fp32")
    step = TrainStep(model, hp, torch.device("cuda:0"))
    step.opt.lr0 = 0.0                                         # freeze the weights: only the masks can change the loss
    step.opt.wd = 0.0
    graphed = GraphedTrainStep(step, batch, warmup=1)
    losses = []
    for _ in range(3):
        graphed.run()
        torch.cuda.synchronize()
        losses.append(float(graphed.last["mle"]))
    assert len(set(losses)) == 3, losses
    assert max(losses) - min(losses) < 0.5 * abs(losses[0]), losses          # same weights, different masks


def test_eager_training_with_changing_batch_geometry():
    """Real data never repeats a batch geometry: every cache keyed on lengths / addresses (row maps,
    workspaces, token rows, split-wgrad pointer arrays, device ints) has to cope with fresh keys each step,
    eagerly, in bf16 training mode (dropout on), and the loss must keep falling on a repeated batch."""
    import math
    from glow_tts_b200.hparams import load_hparams
    from glow_tts_b200.train import TrainStep
    from tests._util import synth_batch
    model, sd, g, batch, mode = load_case("vanilla_small", "bf16")
    model.train()
    step = TrainStep(model, load_hparams(Mode=mode, Precision="bf16"), torch.device("cuda:0"))
    geos = [([23, 17, 9], [140, 96, 50]), ([40, 31], [300, 222]), ([12, 12, 12, 12], [64, 64, 64, 64]),
            ([55], [410]), ([23, 17, 9], [140, 96, 50]), ([33, 8, 21, 14, 29], [200, 52, 130, 88, 176])]
    for i, (tls, mls) in enumerate(geos * 2):
        loss = float(step.run(step.to_device(synth_batch(100 + i, tls, mls))))
        assert math.isfinite(loss), (i, loss)
    first = None
    for _ in range(6):
        loss = float(step.run(step.to_device(batch)))
        first = loss if first is None else first
    assert math.isfinite(loss) and loss < first


def test_state_dict_roundtrip_and_cpu_is_refused():
    from glow_tts_b200 import _lib
    model, sd, g, batch, mode = load_case("vanilla_small", "fp32")
    got = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k].cpu(), v), k
    flat = model.flatten_parameters()
    got = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k].cpu(), v), k            # flattening keeps values and keys
    tokens, tl, mels, ml, spk = batch
    with pytest.raises(_lib.GlowCoreError):
        model(tokens=tokens, token_lengths=tl, mels=mels, mel_lengths=ml, speakers=spk)   # CPU tensors: no fallback
