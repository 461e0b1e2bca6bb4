"""One training step around the sm_100a core: the reference's Train.py:182-233
(forward -> MLE + MSE -> backward -> clip_grad_norm_(5.0) -> RAdam -> Noam) restated
sync-free, with utterance-sharded data parallelism: ONE all-reduce of the flat gradient
buffer per step and ONE fused clip+RAdam kernel over the same buffer.
"""
import ctypes
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, modules


class FusedRAdam:
    """Radam.py:25-90 + Modified_Noam_Scheduler (Noam_Scheduler.py:17-29) on flat buffers."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-6, base=4000, max_norm=5.0):
        self.flat = flat
        self.lr0, self.betas, self.eps, self.wd, self.base, self.max_norm = lr, betas, eps, weight_decay, base, max_norm
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.data.device)
        self.steps = 0          # optimizer steps taken
        self.epoch = 0          # scheduler.last_epoch

    def lr(self):
        e = max(1, self.epoch)
        return self.lr0 * self.base ** 0.5 * (e + self.base) ** -0.5

    def step(self, grad_scale=1.0):
        flat = self.flat
        g = flat.attach_grads()
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
        L = _lib.lib()
        dev = flat.data.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.glow_sqnorm(_lib.ptr(g), g.numel(), _lib.ptr(self.sqnorm), st), "glow_sqnorm")
            _lib.check(L.glow_radam_step(
                _lib.ptr(flat.data), _lib.ptr(g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), g.numel(),
                self.lr(), b1, b2, self.eps, self.wd, step_size, int(n_sma >= 5), self.max_norm, grad_scale,
                _lib.ptr(self.sqnorm), _lib.ptr(self.grad_norm), st), "glow_radam_step")
        self.epoch += 1         # scheduler.step() (Train.py:233)


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

    def to_device(self, batch_host):
        """H2D of one collated batch (pinned -> device, async).  Lengths stay on the host too."""
        tokens, tl, mels, ml, spk = batch_host
        dev = self.device
        return (tokens.to(dev, non_blocking=True), tl, mels.to(dev, non_blocking=True), ml,
                spk.to(dev, non_blocking=True))

    def run(self, batch, global_frames=None, global_positions=None):
        """batch = (tokens, token_lengths(host), mels, mel_lengths(host), speakers) with tensors on
        the device.  Under data parallelism pass the GLOBAL frame count and B*T_x,max so each
        rank's loss is weighted to reproduce the single-process global-batch loss."""
        tokens, tl, mels, ml, spk = batch
        hp, model = self.hp, self.model
        tl_h = [int(v) for v in tl.tolist()]
        ml_h = [int(v) for v in ml.tolist()]
        self.flat.zero_grad()
        out = model(tokens=tokens, token_lengths=None, mels=mels, mel_lengths=None,
                    speakers=spk if hp.Mode.upper() == "SE" else None,
                    host_token_lengths=tl_h, host_mel_lengths=ml_h)
        z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_t = out[:6]
        ml_dev = torch.as_tensor(ml_h, device=z.device)
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
        g = self.flat.grad
        if self.world > 1:
            dist.all_reduce(g)                       # the step's single collective
        self.opt.step(grad_scale=1.0 / self.world)
        self.last = {"loss": loss.detach(), "mle": mle.detach(), "mse": mse.detach(),
                     "grad_norm": self.opt.grad_norm}
        return self.last["loss"]
