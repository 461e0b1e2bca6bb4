"""One training step around the sm_100a core: the reference's Train.py:182-233
(forward -> MLE + MSE -> backward -> clip_grad_norm_(5.0) -> RAdam -> Noam) restated
sync-free, with utterance-sharded data parallelism: ONE all-reduce of the flat gradient
buffer per step and ONE fused clip+RAdam kernel over the same buffer.
"""
import ctypes
import math
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, flow as _flow, modules, rows as _rows


class FusedRAdam:
    """Radam.py:25-90 + Modified_Noam_Scheduler (Noam_Scheduler.py:17-29) on flat buffers.

    The schedule (step count, N_sma, step_size, Noam lr) is host arithmetic (`advance`); the
    update itself is one kernel (`launch`).  With `use_device_schedule()` the nine scalars travel
    through a small device buffer instead of kernel arguments, so the launch can sit in a captured
    CUDA graph while the host keeps advancing the schedule."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-6, base=4000, max_norm=5.0):
        self.flat = flat
        self.lr0, self.betas, self.eps, self.wd, self.base, self.max_norm = lr, betas, eps, weight_decay, base, max_norm
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.steps = 0          # optimizer steps taken
        self.epoch = 0          # scheduler.last_epoch
        self.hyper_dev = None   # [lr, b1, b2, eps, wd, step_size, rectified, max_norm, grad_scale] on the device

    def lr(self):
        e = max(1, self.epoch)
        return self.lr0 * self.base ** 0.5 * (e + self.base) ** -0.5

    def advance(self, grad_scale=1.0):
        """Host side of one step: returns the nine scalars of this step and moves the schedule on."""
        self.steps += 1
        b1, b2 = self.betas
        b2t = b2 ** self.steps
        n_max = 2 / (1 - b2) - 1
        n_sma = n_max - 2 * self.steps * b2t / (1 - b2t)
        if n_sma >= 5:
            step_size = math.sqrt((1 - b2t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma
                                  * n_max / (n_max - 2)) / (1 - b1 ** self.steps)
        else:
            step_size = 1.0 / (1 - b1 ** self.steps)
        hyper = [self.lr(), b1, b2, self.eps, self.wd, step_size, float(n_sma >= 5), self.max_norm, grad_scale]
        self.epoch += 1         # scheduler.step() (Train.py:233)
        return hyper

    def use_device_schedule(self):
        if self.hyper_dev is None:
            self.hyper_dev = torch.zeros(16, dtype=torch.float32, device=self.flat.data.device)
        return self.hyper_dev

    def upload(self, hyper):
        """Stream-ordered H2D of this step's scalars (fresh pinned staging per call)."""
        host = torch.tensor(hyper + [0.0] * (16 - len(hyper)), dtype=torch.float32).pin_memory()
        self.hyper_dev.copy_(host, non_blocking=True)

    def launch(self, hyper=None):
        """clip + RAdam over the flat buffers; hyper = host scalars, or None to read hyper_dev."""
        flat = self.flat
        g = flat.attach_grads()
        L = _lib.lib()
        dev = flat.data.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.glow_sqnorm(_lib.ptr(g), g.numel(), _lib.ptr(self.sqnorm), st), "glow_sqnorm")
            if hyper is None:
                _lib.check(L.glow_radam_step_dev(
                    _lib.ptr(flat.data), _lib.ptr(g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), g.numel(),
                    _lib.ptr(self.hyper_dev), _lib.ptr(self.sqnorm), _lib.ptr(self.grad_norm), st),
                    "glow_radam_step_dev")
            else:
                lr, b1, b2, eps, wd, step_size, rect, max_norm, grad_scale = hyper
                _lib.check(L.glow_radam_step(
                    _lib.ptr(flat.data), _lib.ptr(g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), g.numel(),
                    lr, b1, b2, eps, wd, step_size, int(rect), max_norm, grad_scale,
                    _lib.ptr(self.sqnorm), _lib.ptr(self.grad_norm), st), "glow_radam_step")

    def step(self, grad_scale=1.0):
        self.launch(self.advance(grad_scale))


def ddp_loss_weights(local_frames, local_positions, world, global_frames, global_positions):
    """Weights that make the data-parallel step reproduce the single-process global-batch loss.

    MLE_Loss divides by the LOCAL batch's frame count (Modules.py:1026) and MSELoss averages over the
    LOCAL B*T_x,max positions (Train.py:210), so a plain mean of per-rank losses is not the loss of the
    concatenated batch when shards are ragged.  With gradients SUMMED by the all-reduce and scaled by
    1/world afterwards (FusedRAdam.step(grad_scale=1/world)), rank r's terms must carry
    local/global * world.  Returns (w_mle, w_mse)."""
    return (local_frames * world / float(global_frames), local_positions * world / float(global_positions))


def shard_slice(n_items, rank, world):
    """Contiguous utterance shard of rank `rank`: items [lo, hi)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class TrainStep:
    """Owns the model's flat buffers and the optimizer; `run(batch)` is one Train_Step."""

    def __init__(self, model, hp, device):
        self.model, self.hp, self.device = model, hp, device
        self.flat = model.flatten_parameters()
        self.flat.attach_grads()
        t = hp.Train
        self.opt = FusedRAdam(self.flat, lr=t.Learning_Rate.Initial, betas=(t.ADAM.Beta1, t.ADAM.Beta2),
                              eps=t.ADAM.Epsilon, weight_decay=t.Weight_Decay, base=t.Learning_Rate.Base,
                              max_norm=t.Gradient_Norm)
        self.mle = modules.MLE_Loss()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.last = {}
        # device step counter: mixed into the kernels' dropout seeds (see _lib.set_step_counter)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.set_step_counter(device, self.step_counter)
        _rows.ACCUMULATE = True          # run() joins the side streams before it touches the gradients
        model.layer_Dict["Decoder"].defer_param_grads = True

    def to_device(self, batch_host):
        """H2D of one collated batch (pinned -> device, async).  Lengths stay on the host too."""
        tokens, tl, mels, ml, spk = batch_host
        dev = self.device
        return (tokens.to(dev, non_blocking=True), tl, mels.to(dev, non_blocking=True), ml,
                spk.to(dev, non_blocking=True))

    def run(self, batch, global_frames=None, global_positions=None, device_schedule=False):
        """batch = (tokens, token_lengths(host), mels, mel_lengths(host), speakers) with tensors on
        the device.  Under data parallelism pass the GLOBAL frame count and B*T_x,max so each
        rank's loss is weighted to reproduce the single-process global-batch loss.
        device_schedule=True leaves the optimizer scalars to `opt.hyper_dev` (GraphedTrainStep)."""
        tokens, tl, mels, ml, spk = batch
        hp, model = self.hp, self.model
        tl_h = [int(v) for v in tl.tolist()]
        ml_h = [int(v) for v in ml.tolist()]
        self.step_counter.add_(1)
        self.flat.zero_grad()
        out = model(tokens=tokens, token_lengths=None, mels=mels, mel_lengths=None,
                    speakers=spk if hp.Mode.upper() == "SE" else None,
                    host_token_lengths=tl_h, host_mel_lengths=ml_h)
        z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_t = out[:6]
        ml_dev = _lib.device_ints(ml_h, torch.int64, z.device)
        mle = self.mle(z=z, mean=mel_mean, std=mel_log_std, log_dets=log_dets, lengths=ml_dev)
        mse = F.mse_loss(log_dur, log_dur_t)
        if self.world > 1:
            local_frames = sum(n // 2 * 2 for n in ml_h)
            w_mle, w_mse = ddp_loss_weights(local_frames, log_dur.numel(), self.world, global_frames, global_positions)
            # constant 0.5*log(2*pi) keeps its weight 1 (no gradient, reporting only)
            loss = (mle - 0.5 * math.log(2 * math.pi)) * w_mle + 0.5 * math.log(2 * math.pi) + mse * w_mse
        else:
            loss = mle + mse
        loss.backward()
        _rows.join(self.device)                      # encoder weight gradients forked to the side stream
        _flow.join(self.device)                      # decoder parameter gradients (weight_norm backward)
        g = self.flat.grad
        if self.world > 1:
            dist.all_reduce(g)                       # the step's single collective
        if device_schedule:
            self.opt.launch(None)
        else:
            self.opt.step(grad_scale=1.0 / self.world)
        self.last = {"loss": loss.detach(), "mle": mle.detach(), "mse": mse.detach(),
                     "grad_norm": self.opt.grad_norm}
        return self.last["loss"]


class GraphedTrainStep:
    """One TrainStep.run captured in a CUDA graph and replayed.

    The eager step is CPU-bound (~1100 launches, ~14 us of host time each = 15 ms, against ~6.5 ms of device
    work at B=32); replaying it as one graph removes the host from the loop.  What makes the capture
    legal and the replays *different steps*:
      * inputs live in static device buffers, refreshed by `load()` (pinned H2D) before a replay;
      * everything derived from the host lengths (masks, row map, length tensors) is cached on the
        device (_lib.device_ints, flow.row_map), so the captured region has no host copy;
      * dropout masks come from the device step counter, which the graph increments itself;
      * the RAdam / Noam scalars are read from a device buffer uploaded before every replay.
    A graph is tied to the batch geometry it was captured with (the per-utterance lengths):
    `run()` replays when the lengths match and otherwise falls back to the eager step."""

    def __init__(self, step, batch_host, warmup=3, global_frames=None, global_positions=None):
        self.step, self.device = step, step.device
        tokens, tl, mels, ml, spk = batch_host
        self.key = self._key(tl, ml)
        self.tl, self.ml = tl, ml
        dev = self.device
        self.tokens = torch.empty(tokens.shape, dtype=tokens.dtype, device=dev)
        self.mels = torch.empty(mels.shape, dtype=mels.dtype, device=dev)
        self.spk = torch.empty(spk.shape, dtype=spk.dtype, device=dev)
        self.gf, self.gp = global_frames, global_positions
        self._copy_stream = self._staging = self._staged = None
        step.opt.use_device_schedule()
        self.load(batch_host)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):             # eager warm-up: handles, autotune, caches, ActNorm init
            for _ in range(warmup):
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.warmup_last = {k: v.clone() for k, v in step.last.items()}    # results of the last eager step
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        # GLOW_GRAPH_PRIORITY=-1 captures on a high-priority stream (kernel nodes inherit the priority of the stream
        # they were captured on, so the main chain would get free SMs before the forked weight-gradient / encoder
        # work).  Measured 0.13 ms/step SLOWER at B = 32 -- the forked work is needed on time too -- so the default
        # is a plain stream.
        prio = int(os.environ.get("GLOW_GRAPH_PRIORITY", "0"))
        cap = torch.cuda.Stream(dev, priority=prio) if prio != 0 else torch.cuda.Stream(dev)
        with torch.cuda.graph(self.graph, stream=cap):
            self.loss = step.run(self._batch(), self.gf, self.gp, device_schedule=True)
        self.launches_per_replay = _lib.launch_count() - n0      # libglowcore kernels inside the graph
        self.last = dict(step.last)

    @staticmethod
    def _key(tl, ml):
        return (tuple(int(v) for v in tl.tolist()), tuple(int(v) for v in ml.tolist()))

    def _batch(self):
        return (self.tokens, self.tl, self.mels, self.ml, self.spk)

    def _eager(self):
        opt = self.step.opt
        opt.upload(opt.advance(1.0 / self.step.world))
        return self.step.run(self._batch(), self.gf, self.gp, device_schedule=True)

    def load(self, batch_host):
        """Refresh the static input buffers (async H2D when the host tensors are pinned)."""
        tokens, _, mels, _, spk = batch_host
        self.tokens.copy_(tokens, non_blocking=True)
        self.mels.copy_(mels, non_blocking=True)
        self.spk.copy_(spk, non_blocking=True)

    def prefetch(self, batch_host):
        """Start the H2D copy of a LATER step's batch now, on a copy stream, into staging buffers: it overlaps the
        step in flight (what a data loader's pinned, non_blocking prefetch does).  `run(batch_host)` with the same
        object then only moves staging -> static buffers on the device."""
        if self._key(batch_host[1], batch_host[3]) != self.key:
            self._staged = None
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staging = tuple(torch.empty_like(t) for t in (self.tokens, self.mels, self.spk))
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(self.device))
        tokens, _, mels, _, spk = batch_host
        self._copy_stream.wait_event(self._staging_free)       # the previous staging -> static move has finished
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._staging, (tokens, mels, spk)):
                dst.copy_(src, non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._staged = batch_host

    def run(self, batch_host=None):
        """One optimizer step.  batch_host=None re-uses the resident inputs."""
        if batch_host is not None:
            if self._key(batch_host[1], batch_host[3]) != self.key:
                return self.step.run(self.step.to_device(batch_host), self.gf, self.gp)
            if self._staged is batch_host:                      # prefetched: device-to-device, already resident
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(self._staged_ready)
                for dst, src in zip((self.tokens, self.mels, self.spk), self._staging):
                    dst.copy_(src, non_blocking=True)
                self._staging_free.record(cur)
                self._staged = None
            else:
                self.load(batch_host)
        opt = self.step.opt
        opt.upload(opt.advance(1.0 / self.step.world))
        self.graph.replay()
        return self.loss
