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

from . import _lib, flow as _flow, geometry as _geo, modules, rows as _rows


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

    # ---- checkpoints in the reference's format (Train.py:514-519, 535-553) -------------------------------------
    # The reference saves {'Optimizer': RAdam.state_dict(), 'Scheduler': Modified_Noam_Scheduler.state_dict()}:
    # torch.optim layout {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} with one entry per
    # parameter in model.parameters() order, and the scheduler's {'last_epoch', ...}.  The flat buffers here hold
    # the same numbers at FlatBuffer offsets; `step` is one counter for all parameters (RAdam steps them together).
    def state_dict(self):
        flat = self.flat
        state = {}
        for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
            n = p.numel()
            state[i] = {"step": self.steps,
                        "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr(), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "initial_lr": self.lr0, "params": list(range(len(flat.params)))}
        if self.steps == 0:
            state = {}
        return {"state": state, "param_groups": [group]}

    def scheduler_state_dict(self):
        return {"base": self.base, "base_lrs": [self.lr0], "last_epoch": self.epoch, "_step_count": self.epoch + 1,
                "_last_lr": [self.lr()]}

    def load_state_dict(self, sd, scheduler_sd=None):
        """Accepts RAdam.state_dict() of the reference (Radam.py:25-90) -- or state_dict() above -- and, optionally,
        its scheduler's state_dict(); restores moments, the rectification step count and the LR schedule position."""
        flat = self.flat
        state = sd.get("state", {})
        if state and len(state) != len(flat.params):
            raise ValueError("optimizer state has %d entries, the model %d parameters" % (len(state), len(flat.params)))
        steps = 0
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
            st = state.get(i, state.get(str(i)))
            if st is None:
                continue
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state %d has shape %s, parameter %s" % (i, tuple(st["exp_avg"].shape), tuple(p.shape)))
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps = max(steps, int(st["step"]))
        self.steps = steps
        groups = sd.get("param_groups") or [{}]
        g = groups[0]
        self.betas = tuple(g.get("betas", self.betas))
        self.eps = g.get("eps", self.eps)
        self.wd = g.get("weight_decay", self.wd)
        self.lr0 = g.get("initial_lr", self.lr0)
        if scheduler_sd is not None:
            self.epoch = int(scheduler_sd.get("last_epoch", self.epoch))
            self.base = scheduler_sd.get("base", self.base)
            if scheduler_sd.get("base_lrs"):
                self.lr0 = scheduler_sd["base_lrs"][0]
        else:
            self.epoch = self.steps


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
        # Data parallelism: the flat gradient buffer is all-reduced in BUCKETS -- one per decoder block, issued on a
        # communication stream as soon as that block's parameter gradients are final (glow_flow_wait_block_grads), so
        # block k's reduction runs while blocks k-1 .. 0 are still in their backward; the encoder's slice goes last.
        # OPT-IN (GLOW_ALLREDUCE_OVERLAP=1), eager steps only: verified on 2 GPUs (tests/test_ddp_nccl_gpu.py: same
        # gradients as the single all-reduce, bit-identical across ranks), but the one 8-GPU run this round could afford
        # deadlocked with the buckets captured inside lazily built per-bucket graphs (ranks capture at different steps;
        # profiles/README.md), so the default is ONE all-reduce of the whole buffer after the backward, issued eagerly
        # BETWEEN the two captured halves of the step (GraphedTrainStep) -- no collective lives inside a CUDA graph.
        self.overlap_allreduce = self.world > 1 and os.environ.get("GLOW_ALLREDUCE_OVERLAP", "0") == "1"
        self._comm = None
        self._buckets = None

    def _grad_buckets(self):
        """[(lo, hi)] element ranges of the flat buffer: decoder blocks (in index order), then (rest_lo, rest_hi)."""
        if self._buckets is None:
            flat = self.flat
            flows = self.model.layer_Dict["Decoder"].layer_Dict["Flows"]
            starts = [flat.offset_of(next(iter(blk.parameters()))) for blk in flows]
            dec_lo = starts[0]
            last = list(flows[-1].parameters())[-1]
            dec_hi = flat.offset_of(last) + (last.numel() + 3) // 4 * 4
            assert all(a < b for a, b in zip(starts, starts[1:])) and dec_hi <= flat.total
            blocks = [(lo, hi) for lo, hi in zip(starts, starts[1:] + [dec_hi])]
            rest = [(0, dec_lo)] + ([(dec_hi, flat.total)] if dec_hi < flat.total else [])
            self._buckets = (blocks, [r for r in rest if r[1] > r[0]])
        return self._buckets

    def _allreduce_overlapped(self, g):
        """Called right after loss.backward() has been issued: per-block all-reduces behind the per-block events."""
        dev = self.device
        if self._comm is None:
            self._comm = torch.cuda.Stream(dev)
        comm, cur = self._comm, torch.cuda.current_stream(dev)
        blocks, rest = self._grad_buckets()
        L = _lib.lib()
        # no comm.wait_stream(cur) here: the whole backward is already queued on `cur`, waiting for its tail would
        # serialise the buckets behind it.  Each bucket waits for its own block's event instead, which is ordered
        # after this step's zero_grad through the backward itself.
        with torch.cuda.device(dev):
            for k in reversed(range(len(blocks))):
                _lib.check(L.glow_flow_wait_block_grads(ctypes.c_void_p(comm.cuda_stream), k), "glow_flow_wait_block_grads")
                with torch.cuda.stream(comm):
                    dist.all_reduce(g[blocks[k][0]:blocks[k][1]])
        _rows.join(dev)                                        # encoder weight gradients (side-stream lanes)
        _flow.join(dev)
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            for lo, hi in rest:
                dist.all_reduce(g[lo:hi])
        cur.wait_stream(comm)

    def to_device(self, batch_host):
        """H2D of one collated batch (pinned -> device, async).  Lengths stay on the host too."""
        tokens, tl, mels, ml, spk = batch_host
        dev = self.device
        return (tokens.to(dev, non_blocking=True), tl, mels.to(dev, non_blocking=True), ml,
                spk.to(dev, non_blocking=True))

    def run(self, batch, global_frames=None, global_positions=None, device_schedule=False, geometry=None, phase="all"):
        """phase: "all" (default) the whole step; "backward": zero_grad .. backward + joins, NO collective, NO update
        (GraphedTrainStep under data parallelism captures this half and `update()` separately and issues the
        all-reduce eagerly between the two replays).
        batch = (tokens, token_lengths(host), mels, mel_lengths(host), speakers) with tensors on
        the device.  Under data parallelism pass the GLOBAL frame count and B*T_x,max so each
        rank's loss is weighted to reproduce the single-process global-batch loss.
        device_schedule=True leaves the optimizer scalars to `opt.hyper_dev` (GraphedTrainStep).
        geometry: a geometry.StepGeometry already updated for this batch (bucketed graph replay): lengths, row
        maps and the loss weights are then read from its device buffers, nothing from the host lists."""
        tokens, tl, mels, ml, spk = batch
        hp, model = self.hp, self.model
        # the decoder's weight preparation (weight_norm -> slab images) only depends on the parameters: start it on
        # the side stream before anything else of the step is queued (GlowTTS.forward finds it already running)
        model.layer_Dict["Decoder"].begin_prepare(self.device)
        self.step_counter.add_(1)
        self.flat.zero_grad()
        # encoder weight gradients accumulate straight into the flat gradient buffer on the library's side stream
        # for the duration of THIS step only (joined below, before anything reads the gradients)
        prev_acc, prev_defer = _rows.ACCUMULATE, model.layer_Dict["Decoder"].defer_param_grads
        prev_fused = _flow.FUSED_PARAM_GRADS
        _rows.ACCUMULATE = True
        model.layer_Dict["Decoder"].defer_param_grads = True
        if self.overlap_allreduce and phase == "all":
            _flow.FUSED_PARAM_GRADS = True               # every block's parameter gradients right behind its weight gradients
        try:
            c = 0.5 * math.log(2 * math.pi)
            if geometry is not None:
                out = model(tokens=tokens, token_lengths=None, mels=mels, mel_lengths=None,
                            speakers=spk if hp.Mode.upper() == "SE" else None, geometry=geometry)
                z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_t = out[:6]
                mle = self.mle(z=z, mean=mel_mean, std=mel_log_std, log_dets=log_dets, lengths=geometry.ml64)
                # MSELoss over the batch's own B * T_x,max positions (Train.py:210): both operands are zero on the
                # padding, so the padded sum times the device-resident 1 / (B * T_x,max) is the same number
                sc = geometry.scal
                mse = ((log_dur - log_dur_t) ** 2).sum() * sc[0]
                loss = (mle - c) * sc[1] + c + mse * sc[2]
            else:
                tl_h = [int(v) for v in tl.tolist()]
                ml_h = [int(v) for v in ml.tolist()]
                out = model(tokens=tokens, token_lengths=None, mels=mels, mel_lengths=None,
                            speakers=spk if hp.Mode.upper() == "SE" else None,
                            host_token_lengths=tl_h, host_mel_lengths=ml_h)
                z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_t = out[:6]
                ml_dev = _lib.device_ints(ml_h, torch.int64, z.device)
                mle = self.mle(z=z, mean=mel_mean, std=mel_log_std, log_dets=log_dets, lengths=ml_dev)
                mse = F.mse_loss(log_dur, log_dur_t)
                if self.world > 1:
                    local_frames = sum(n // 2 * 2 for n in ml_h)
                    w_mle, w_mse = ddp_loss_weights(local_frames, log_dur.numel(), self.world, global_frames,
                                                    global_positions)
                    # constant 0.5*log(2*pi) keeps its weight 1 (no gradient, reporting only)
                    loss = (mle - c) * w_mle + c + mse * w_mse
                else:
                    loss = mle + mse
            dec = model.layer_Dict["Decoder"]
            fused_before = getattr(dec, "fused_backward_calls", 0)
            loss.backward()
            g = self.flat.grad
            # the per-block events only exist if the decoder's backward really took the per-block path
            if (phase == "all" and self.overlap_allreduce
                    and getattr(dec, "fused_backward_calls", 0) == fused_before + 1):
                self._allreduce_overlapped(g)
            else:
                _rows.join(self.device)                  # encoder weight gradients forked to the side stream
                _flow.join(self.device)                  # decoder parameter gradients (weight_norm backward)
                if self.world > 1 and phase == "all":
                    dist.all_reduce(g)                   # one collective over the whole flat buffer
        finally:
            _rows.ACCUMULATE = prev_acc
            model.layer_Dict["Decoder"].defer_param_grads = prev_defer
            _flow.FUSED_PARAM_GRADS = prev_fused
        self.last = {"loss": loss.detach(), "mle": mle.detach(), "mse": mse.detach(),
                     "grad_norm": self.opt.grad_norm}
        if phase == "all":
            self.update(device_schedule)
        return self.last["loss"]

    def update(self, device_schedule=False):
        """clip + RAdam + Noam over the (all-reduced) flat gradient buffer."""
        if device_schedule:
            self.opt.launch(None)
        else:
            self.opt.step(grad_scale=1.0 / self.world)

    def loss_scalars(self, tl, ml, global_frames=None, global_positions=None):
        """Host side of the geometry's loss scalars for one batch: [1/(B*T_x,max), w_mle, w_mse]."""
        b, tx = len(tl), max(int(n) for n in tl)
        if self.world > 1 and global_frames is not None:
            local_frames = sum(int(n) // 2 * 2 for n in ml)
            w_mle, w_mse = ddp_loss_weights(local_frames, b * tx, self.world, global_frames, global_positions)
        else:
            w_mle = w_mse = 1.0
        return [1.0 / float(b * tx), w_mle, w_mse]


class _Bucket:
    """One captured graph: its geometry buffers, the graph, the tensors it leaves its results in, and strong
    references to every cache-owned object the capture touched (the graph holds raw pointers)."""

    def __init__(self, geo):
        self.geo, self.graph, self.loss, self.last, self.keep, self.launches = geo, None, None, {}, [], 0
        self.replays = 0


class GraphedTrainStep:
    """TrainStep.run captured in CUDA graphs and replayed -- one graph per geometry BUCKET, not per batch.

    The eager step is CPU-bound (~1100 launches, ~14 us of host time each = 15 ms, against ~6 ms of device
    work at B=32); replaying it as one graph removes the host from the loop.  The reference's loop feeds a
    different batch geometry every step (Train.py:582-584, Datasets.py:225-250), so a graph must not depend on
    the per-utterance lengths:
      * inputs live in static device buffers padded to (T_text_pad, T_mel_pad), refreshed per step (pinned H2D);
      * everything derived from the lengths -- packed-row maps of decoder and encoder, length tensors, masks,
        loss normalisers, data-parallel loss weights -- lives in a geometry.StepGeometry blob refreshed by ONE
        small H2D copy per step; the captured launches only depend on the bucket (batch, decoder rows_pad,
        encoder rows_pad), which is the row count rounded up to 512 / 256 rows;
      * dropout masks come from the device step counter, which the graph increments itself;
      * the RAdam / Noam scalars are read from a device buffer uploaded before every replay.
    The first batch of a bucket runs eagerly (it IS that batch's training step) and the bucket's graph is
    captured right after it; every later batch of the bucket is a replay."""

    def __init__(self, step, batch_host=None, warmup=3, global_frames=None, global_positions=None,
                 t_text_pad=None, t_mel_pad=None):
        self.step, self.device = step, step.device
        pat = getattr(step.hp.Train, "Train_Pattern", None)
        t_text = int(t_text_pad) if t_text_pad else (int(pat.Text_Length.Max) + 2 if pat else 202)
        t_mel = int(t_mel_pad) if t_mel_pad else (int(pat.Mel_Length.Max) if pat else 1000)
        t_mel += t_mel % 2
        # padded sizes a batch may be given: the smallest menu entry that holds its longest item.  Padding is cheap for
        # the packed-row kernels (they never touch it) but not for what runs on [B, C, T] tensors (alignment search,
        # path expansion, losses, the torch-side glue), so short-utterance batches (VCTK-shaped) get smaller shapes.
        self.text_menu = [t_text] if t_text_pad else sorted({max(8, (t_text * q // 4 + 7) // 8 * 8) for q in (1, 2, 3)} | {t_text})
        self.mel_menu = [t_mel] if t_mel_pad else sorted({max(64, (t_mel * q // 4 + 63) // 64 * 64) for q in (1, 2, 3)} | {t_mel})
        self.t_text, self.t_mel = t_text, t_mel
        self.gf, self.gp = global_frames, global_positions
        self.buckets = {}
        self._inputs = {}                      # batch size -> static (tokens, mels, spk)
        self._staging = {}                     # batch size -> staging copies for prefetch()
        self._copy_stream = None
        self._staged = None
        self._pool = None
        self.current = None                    # bucket of the most recent step
        self.warmup_last = {}
        # Data parallelism: NO collective inside a captured graph.  Ranks build their per-bucket graphs at different
        # steps (their batches differ), and a graph that replays NCCL kernels on one rank while another rank issues
        # the same collective eagerly (or captures) is exactly the mix NCCL's ordering rules make fragile.  So a
        # bucket's graph ends before the all-reduce; the all-reduce is issued eagerly; clip + RAdam is a second,
        # bucket-independent graph.
        self.split = step.world > 1
        self._opt_graph = None
        step.opt.use_device_schedule()
        if batch_host is not None:
            # constructor contract of round 1: `warmup` eager steps on the example batch (handles, caches, ActNorm
            # init), then its bucket is captured; the results of the last eager step are in `warmup_last`
            for _ in range(max(1, warmup)):
                self._step_eager(batch_host, capture=False)
            self.warmup_last = {k: v.clone() for k, v in step.last.items()}
            self._capture(self.current)

    # ------------------------------------------------------------------ buckets
    def bucket_key(self, tl, ml):
        tl, ml = [int(v) for v in tl.tolist()], [int(v) for v in ml.tolist()]
        tt = next((t for t in self.text_menu if t >= max(tl)), None)
        tm = next((t for t in self.mel_menu if t >= max(ml)), None)
        if tt is None or tm is None:
            raise ValueError("batch exceeds the padded sizes (%d tokens, %d mel frames)" % (self.t_text, self.t_mel))
        return _geo.StepGeometry.bucket_of(tl, ml, tt, tm)

    @property
    def key(self):
        return self.current.geo.key if self.current is not None else None

    @property
    def graph(self):
        return self.current.graph

    @property
    def last(self):
        return self.current.last if self.current is not None else {}

    @property
    def loss(self):
        return self.current.loss

    @property
    def launches_per_replay(self):
        return self.current.launches if self.current is not None else 0

    def _bucket(self, key):
        bk = self.buckets.get(key)
        if bk is None:
            b, rd, re, tt, tm = key
            bk = self.buckets[key] = _Bucket(_geo.StepGeometry(b, tt, tm, rd, re, self.device))
        return bk

    def _shape_key(self, batch_host):
        key = self.bucket_key(batch_host[1], batch_host[3])
        return (key[0], key[3], key[4])                          # (batch, T_text_pad, T_mel_pad)

    def _static_inputs(self, batch_host):
        tokens, _, mels, _, spk = batch_host
        k = self._shape_key(batch_host)
        if k not in self._inputs:
            dev = self.device
            b, tt, tm = k
            self._inputs[k] = (torch.ones((b, tt), dtype=tokens.dtype, device=dev),
                               torch.full((b, mels.shape[1], tm), -4.0, dtype=mels.dtype, device=dev),
                               torch.zeros((b,), dtype=spk.dtype, device=dev))
        return self._inputs[k]

    def load(self, batch_host):
        """Refresh the static input buffers (async H2D when the host tensors are pinned).  Positions beyond the
        batch's own padded size keep stale values: every consumer masks by the lengths."""
        tokens, _, mels, _, spk = batch_host
        st, sm, ss = self._static_inputs(batch_host)
        st[:, :tokens.shape[1]].copy_(tokens, non_blocking=True)
        sm[:, :, :mels.shape[2]].copy_(mels, non_blocking=True)
        ss.copy_(spk, non_blocking=True)
        return st, sm, ss

    def prefetch(self, batch_host):
        """Start the H2D copy of a LATER step's batch now, on a copy stream, into staging buffers: it overlaps the
        step in flight (what a data loader's pinned, non_blocking prefetch does).  `run(batch_host)` with the same
        object then only moves staging -> static buffers on the device."""
        tokens, _, mels, _, spk = batch_host
        b = self._shape_key(batch_host)
        self._static_inputs(batch_host)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream(self.device))
        if b not in self._staging:
            self._staging[b] = tuple(torch.empty_like(t) for t in self._inputs[b])
        self._copy_stream.wait_event(self._staging_free)       # the previous staging -> static move has finished
        with torch.cuda.stream(self._copy_stream):
            gt, gm, gs = self._staging[b]
            gt[:, :tokens.shape[1]].copy_(tokens, non_blocking=True)
            gm[:, :, :mels.shape[2]].copy_(mels, non_blocking=True)
            gs.copy_(spk, non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._staged = batch_host

    def _inputs_for(self, batch_host):
        tokens, _, mels, _, _ = batch_host
        if self._staged is batch_host:                          # prefetched: device-to-device, already resident
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._staged_ready)
            st, sm, ss = self._static_inputs(batch_host)
            gt, gm, gs = self._staging[self._shape_key(batch_host)]
            st[:, :tokens.shape[1]].copy_(gt[:, :tokens.shape[1]], non_blocking=True)
            sm[:, :, :mels.shape[2]].copy_(gm[:, :, :mels.shape[2]], non_blocking=True)
            ss.copy_(gs, non_blocking=True)
            self._staging_free.record(cur)
            self._staged = None
            return st, sm, ss
        return self.load(batch_host)

    # ------------------------------------------------------------------ steps
    def _prepare(self, batch_host, global_frames, global_positions):
        tokens, tl, mels, ml, spk = batch_host
        bk = self._bucket(self.bucket_key(tl, ml))
        st, sm, ss = self._inputs_for(batch_host)
        gf = global_frames if global_frames is not None else self.gf
        gp = global_positions if global_positions is not None else self.gp
        bk.geo.update(tl.tolist(), ml.tolist(), self.step.loss_scalars(tl.tolist(), ml.tolist(), gf, gp))
        self.current = bk
        return bk, (st, tl, sm, ml, ss)

    def _step_eager(self, batch_host, capture=True, global_frames=None, global_positions=None):
        bk, dev_batch = self._prepare(batch_host, global_frames, global_positions)
        opt = self.step.opt
        opt.upload(opt.advance(1.0 / self.step.world))
        with _lib.capture_keepalive() as keep:
            loss = self.step.run(dev_batch, global_frames, global_positions, device_schedule=True, geometry=bk.geo)
        bk.keep.extend(keep)
        bk.last = dict(self.step.last)
        bk.loss = loss
        bk._dev_batch = dev_batch
        if capture:
            self._capture(bk)
        return loss

    def _capture(self, bk):
        """Capture one TrainStep.run on bucket `bk` (whose eager step has just run: caches are warm)."""
        dev = self.device
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()         # replays are serialised: all graphs share one pool
        # The step is captured on a HIGH-priority stream: kernel nodes inherit the priority of the stream they were
        # captured on, so the latency-bound main chain (decoder forward / data gradients, losses) gets free SMs before
        # the encoder's stream and the weight-gradient lanes (default priority).  Round 1 measured this 0.13 ms slower
        # (the library weight-gradient GEMMs it starved were needed on time); with the few-CTA background batches of
        # round 2 it is 1 % faster (profiles/bench_r02u_graph_priority_*.json).  GLOW_GRAPH_PRIORITY=0: plain stream.
        prio = int(os.environ.get("GLOW_GRAPH_PRIORITY", "-1"))
        cap = torch.cuda.Stream(dev, priority=prio) if prio != 0 else torch.cuda.Stream(dev)
        with _lib.capture_keepalive() as keep:
            with torch.cuda.graph(graph, pool=self._pool, stream=cap):
                loss = self.step.run(bk._dev_batch, device_schedule=True, geometry=bk.geo,
                                     phase="backward" if self.split else "all")
            if self.split and self._opt_graph is None:           # clip + RAdam: one graph for every bucket
                self._opt_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._opt_graph, pool=self._pool, stream=cap):
                    self.step.update(device_schedule=True)
        bk.keep.extend(keep)
        # the graph's result tensors are placeholders until its first replay: give them the eager step's results
        eager_loss, eager_last = bk.loss, bk.last
        bk.graph, bk.loss, bk.last = graph, loss, dict(self.step.last)
        with torch.no_grad():
            if eager_loss is not None:
                bk.loss.copy_(eager_loss)
            for k, v in eager_last.items():
                if k in bk.last and bk.last[k] is not v:
                    bk.last[k].copy_(v)
        bk.launches = _lib.launch_count() - n0                  # libglowcore kernels inside the graph

    def run(self, batch_host=None, global_frames=None, global_positions=None):
        """One optimizer step on `batch_host` (pinned host tensors + host length tensors).  batch_host=None
        re-uses the resident inputs and geometry of the previous step."""
        if batch_host is None:
            bk = self.current
            if bk is None or bk.graph is None:
                raise _lib.GlowCoreError("GraphedTrainStep.run(None): no batch has been loaded yet")
        else:
            key = self.bucket_key(batch_host[1], batch_host[3])
            bk = self.buckets.get(key)
            if bk is None or bk.graph is None:
                return self._step_eager(batch_host, True, global_frames, global_positions)   # first batch of the bucket
            bk, _ = self._prepare(batch_host, global_frames, global_positions)
        opt = self.step.opt
        opt.upload(opt.advance(1.0 / self.step.world))
        bk.graph.replay()
        if self.split:
            dist.all_reduce(self.step.flat.grad)                 # eager, between the two captured halves
            self._opt_graph.replay()
        bk.replays += 1
        return bk.loss
