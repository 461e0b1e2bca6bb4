"""GPU: ActNorm data-dependent initialisation (Activation_Norm.initialize, Modules.py:698-711) through the C ABI
(glow_flow_pack_rows / glow_actnorm_stats / glow_flow_block_forward) against fixtures made by running the
reference's Decoder on an UNINITIALISED model (tools/make_golden_model.py::run_ddi), and its data-parallel
variant: two ranks, each holding half of the batch, must end up with bit-identical logs / bias equal to the
single-process initialisation on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch

from tests._model_util import build_model, mel_mask
from tests._util import GOLD, rel_err, synth_batch

pytestmark = pytest.mark.gpu

DDI = {"ddi_vanilla": ("Vanilla", [40, 31, 18, 25], [300, 222, 96, 164], 41, 1234),
       "ddi_se": ("SE", [33, 21], [200, 128], 42, 4321)}


def _uninitialised(mode, wseed, precision):
    model, sd = build_model(mode, wseed, precision)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = False
    return model


def _logs_bias(model):
    flows = model.layer_Dict["Decoder"].layer_Dict["Flows"]
    return (torch.stack([f.layers[0].logs.detach().view(-1) for f in flows]).cpu(),
            torch.stack([f.layers[0].bias.detach().view(-1) for f in flows]).cpu())


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
@pytest.mark.parametrize("name", list(DDI))
def test_data_dependent_init_matches_reference(name, precision, tol):
    mode, tls, mls, bseed, wseed = DDI[name]
    g = np.load(os.path.join(GOLD, "model_%s.npz" % name))
    model = _uninitialised(mode, wseed, precision)
    model.eval()
    tokens, tl, mels, ml, spk = synth_batch(bseed, tls, mls)
    dec = model.layer_Dict["Decoder"]
    emb = model.layer_Dict["LUT"](spk.cuda()).detach() if mode == "SE" else None
    with torch.no_grad():
        z, ld, _ = dec(mels.cuda(), mel_mask(ml, "cuda"), emb)
    assert all(f.layers[0].initialized for f in dec.layer_Dict["Flows"])
    logs, bias = _logs_bias(model)
    assert rel_err(logs, g["logs"]) < tol, rel_err(logs, g["logs"])
    assert rel_err(bias, g["bias"]) < tol, rel_err(bias, g["bias"])
    assert rel_err(z.cpu(), g["z"]) < max(tol, 1e-3) * (1 if precision == "fp32" else 2)
    assert rel_err(ld.cpu(), g["logdet"]) < tol
    # a second forward must not re-initialise
    with torch.no_grad():
        dec(mels.cuda(), mel_mask(ml, "cuda"), emb)
    logs2, bias2 = _logs_bias(model)
    assert torch.equal(logs, logs2) and torch.equal(bias, bias2)


def test_actnorm_stats_entry_point():
    """glow_actnorm_stats on packed rows == masked sums (count, sum x, sum x^2) per channel."""
    from glow_tts_b200 import _lib, flow
    rm = flow.row_map([17, 40, 5], torch.device("cuda:0"))
    torch.manual_seed(3)
    x = torch.randn(rm.rows_pad, 160, device="cuda")
    out = torch.empty(3 * 160, device="cuda")
    _lib.check(_lib.lib().glow_actnorm_stats(_lib.ptr(x), rm.row_utt.data_ptr(), rm.rows_pad, 160, _lib.ptr(out),
                                             _lib.stream_ptr()), "glow_actnorm_stats")
    valid = (rm.row_utt >= 0).double().unsqueeze(1)
    want = torch.stack([valid.sum().expand(160), (x.double() * valid).sum(0), (x.double() ** 2 * valid).sum(0)])
    assert rel_err(out.view(3, 160).cpu(), want.cpu()) < 1e-6
    assert float(out[0]) == 62.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddi_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    # one process per rank; both ranks share cuda:0 on a one-GPU box, so the collective goes through gloo
    # (NCCL refuses two ranks on one device); on the multi-GPU bench the same code path runs over NCCL
    n_gpu = torch.cuda.device_count()
    dev = torch.device("cuda", rank % n_gpu)
    torch.cuda.set_device(dev)
    backend = "nccl" if n_gpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from glow_tts_b200.train import shard_slice
        mode, tls, mls, bseed, wseed = DDI["ddi_vanilla"]
        from glow_tts_b200 import modules
        from glow_tts_b200.hparams import load_hparams
        from tests._util import synth_state_dict
        modules.set_hparams(load_hparams(Mode=mode, Precision="fp32"))
        model = modules.GlowTTS()
        model.load_state_dict(synth_state_dict(model.state_dict(), wseed), strict=True)
        model = model.to(dev).eval()
        tokens, tl, mels, ml, spk = synth_batch(bseed, tls, mls)
        lo, hi = shard_slice(len(mls), rank, world)
        ml_s = ml[lo:hi]
        shard = mels[lo:hi, :, :int(ml_s.max())].contiguous().to(dev)
        with torch.no_grad():
            model.layer_Dict["Decoder"](shard, mel_mask(ml_s, dev), None)
        logs, bias = _logs_bias(model)
        q.put((rank, logs.numpy(), bias.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_init_is_identical_and_equals_the_global_batch_init():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddi_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, logs, bias = q.get(timeout=900)
        got[rank] = (logs, bias)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])      # replicas agree bit for bit
    g = np.load(os.path.join(GOLD, "model_ddi_vanilla.npz"))                                   # == whole-batch init
    assert rel_err(got[0][0], g["logs"]) < 1e-3 and rel_err(got[0][1], g["bias"]) < 1e-3
