"""CPU, world_size 2, gloo: the data-parallel arithmetic of glow_tts_b200.train (utterance
sharding, loss weights, ONE all-reduce of the flat gradient, 1/world scaling) reproduces the
single-process global-batch gradient.  The model arithmetic on the CPU is the oracle's (the
product path has no CPU fallback); what is under test is the host-side sharding logic."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _small_model():
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    from tests._util import synth_state_dict
    modules.set_hparams(load_hparams(Mode="Vanilla", **{"Decoder.Stack": 2, "Encoder.Transformer.Stacks": 1}))
    model = modules.GlowTTS()
    return synth_state_dict(model.state_dict(), 9)


def _grads(sd, hp, batch, w_mle=1.0, w_mse=1.0):
    from oracle import glow_oracle as G
    leaves = G.state_dict_to_leaves(sd)
    tokens, tl, mels, ml, spk = batch
    out = G.glow_forward(leaves, hp, tokens, tl, mels, ml, None, False, "port")
    _, mle, mse = G.losses(out, ml, hp)
    c = 0.5 * math.log(2 * math.pi)
    loss = (mle - c) * w_mle + c + mse * w_mse
    loss.backward()
    keys = [k for k in sorted(leaves) if leaves[k].requires_grad and leaves[k].grad is not None]
    return torch.cat([leaves[k].grad.flatten() for k in keys]), int(out[4].numel())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from glow_tts_b200.train import ddp_loss_weights, shard_slice
        from oracle import glow_oracle as G
        from tests._util import synth_batch
        sd = _small_model()
        hp = G.OracleHP(dec_stack=2, enc_stacks=1)
        tls, mls = [23, 9, 17, 12], [140, 50, 96, 64]                 # ragged on purpose
        tokens, tl, mels, ml, spk = synth_batch(4, tls, mls)
        lo, hi = shard_slice(len(tls), rank, world)
        tl_s, ml_s = tl[lo:hi], ml[lo:hi]
        shard = (tokens[lo:hi, :int(tl_s.max())].contiguous(), tl_s, mels[lo:hi, :, :int(ml_s.max())].contiguous(),
                 ml_s, spk[lo:hi])
        counts = torch.tensor([float(ml_s.sum()), float(hi - lo)])
        dist.all_reduce(counts)                                        # global frames, global batch
        tmax = torch.tensor([float(tl_s.max())])
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)                    # T_x,max of the global batch
        gpos = float(counts[1]) * float(tmax)                          # what MSELoss averages over (Train.py:210)
        w_mle, w_mse = ddp_loss_weights(float(ml_s.sum()), (hi - lo) * int(tl_s.max()), world,
                                        float(counts[0]), gpos)
        g, _ = _grads(sd, hp, shard, w_mle, w_mse)
        dist.all_reduce(g)                                             # the step's single collective
        g /= world                                                     # FusedRAdam grad_scale
        if rank == 0:
            full, _ = _grads(sd, hp, (tokens, tl, mels, ml, spk))      # single process, whole batch
            q.put((g.numpy(), full.numpy(), float(counts[0]), gpos))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_equals_global_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, full, gframes, gpos = q.get(timeout=600)
    got, full = torch.from_numpy(got), torch.from_numpy(full)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert gframes == 350.0 and gpos == 4 * 23.0
    err = float((got - full).abs().max() / full.abs().max())
    assert err < 1e-4, err


def test_shard_slice_covers_everything_once():
    from glow_tts_b200.train import shard_slice
    for n in (1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_slice(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
