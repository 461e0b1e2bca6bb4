"""GPU x 2 (NCCL): the bucketed gradient all-reduce that overlaps the decoder's backward (train.TrainStep,
glow_flow_wait_block_grads) gives the gradients of the single whole-buffer all-reduce, eagerly and replayed from a
captured graph; and the process group tears down cleanly after graphs that captured collectives.  Skipped on a
one-GPU box (NCCL refuses two ranks on one device); tests/test_ddp_gloo.py covers the sharding arithmetic on CPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    try:
        from glow_tts_b200 import modules
        from glow_tts_b200.hparams import load_hparams
        from glow_tts_b200.train import TrainStep, GraphedTrainStep
        from tests._util import synth_batch, synth_state_dict
        hp = load_hparams(Mode="Vanilla", Precision="bf16", **{"Decoder.Stack": 4, "Encoder.Transformer.Stacks": 2})
        modules.set_hparams(hp)
        model = modules.GlowTTS()
        model.load_state_dict(synth_state_dict(model.state_dict(), 5), strict=True)
        for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
            blk.layers[0].initialized = True
        model = model.to(dev).eval()
        step = TrainStep(model, hp, dev)
        step.opt.lr0 = 0.0
        step.opt.wd = 0.0                                        # parameters stay put: gradients are comparable
        assert not step.overlap_allreduce                       # opt-in (GLOW_ALLREDUCE_OVERLAP=1)
        geos = [([23, 17, 9], [140, 96, 50]), ([30, 12, 21], [180, 70, 120])]
        batch = synth_batch(10 + rank, *geos[rank])
        gf, gp = 140 + 96 + 50 + 180 + 70 + 120, 6 * 30
        grads = {}

        def across_ranks(g):
            """(number of elements that differ from rank 0's copy, first and last such index)"""
            peer = g.clone()
            dist.broadcast(peer, src=0)
            bad = (peer != g).nonzero().flatten()
            return (int(bad.numel()), int(bad[0]) if bad.numel() else -1, int(bad[-1]) if bad.numel() else -1)

        diag = {}
        for overlap in (True, False):
            step.overlap_allreduce = overlap
            step.run(step.to_device(batch), global_frames=gf, global_positions=gp)
            torch.cuda.synchronize()
            grads[overlap] = step.flat.grad.detach().clone()
            diag["eager_overlap" if overlap else "eager_single"] = across_ranks(grads[overlap])
        scale = float(grads[False].abs().max())
        err = float((grads[True] - grads[False]).abs().max()) / scale
        # replayed from the captured halves of the step with the eager all-reduce between them (no collective lives
        # inside a graph: ranks capture their per-bucket graphs at different steps)
        step.overlap_allreduce = False
        graphed = GraphedTrainStep(step, global_frames=gf, global_positions=gp)
        assert graphed.split
        pinned = (batch[0].pin_memory(), batch[1], batch[2].pin_memory(), batch[3], batch[4].pin_memory())
        graphed.run(pinned)
        graphed.run(pinned)
        torch.cuda.synchronize()
        g_graph = step.flat.grad.detach().clone()
        err_graph = float((g_graph - grads[False]).abs().max()) / scale
        # every rank must hold the same reduced gradient
        diag["graph_overlap"] = across_ranks(g_graph)
        diag["buckets"] = step._grad_buckets()
        same = diag
        graphed.buckets.clear()
        graphed.current = None
        torch.cuda.synchronize()
        q.put((rank, err, err_graph, same, scale))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL: one rank per device)")
def test_bucketed_allreduce_matches_the_single_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)                                      # destroy_process_group returned on both ranks
        assert p.exitcode == 0
    for rank, err, err_graph, same, scale in res:
        assert scale > 0
        assert err < 2e-3, (rank, err)                           # bf16 weight gradients: split order differs slightly
        assert err_graph < 2e-3, (rank, err_graph)
        for name in ("eager_single", "eager_overlap", "graph_overlap"):
            assert same[name][0] == 0, (rank, name, same)           # the reduced gradient is identical on every rank
