"""The lines around the alignment search in ``GlowTTS.forward`` and the MLE loss, on libglowcore's fused
kernels (csrc/align.cu):

* ``log_p``            Modules.py:107-116, fp32, only the t_x[b] x t_y[b] corner the search reads
* ``expand_by_path``   Modules.py:120-122: ``mean @ attentions``, ``log_Std @ attentions`` as a gather by the
                       frame -> token index of the path, ``log_Duration_Targets`` from its per-token frame counts;
                       backward = per-token sums over contiguous frame runs (no atomics, deterministic)
* ``mle_loss``         Modules.py:1020-1029 with its backward, one pass over the data each

There is no CPU path: tensors must be CUDA tensors.
"""
import ctypes

import torch

from . import _lib


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def log_p(z, mean, log_std, t_len, m_len):
    """z [B,C,T_y], mean / log_std [B,C,T_x] -> log_P [B,T_x,T_y] (entries outside t_len[b] x m_len[b] are
    NOT written: the search only reads that corner)."""
    _lib.require_cuda(z, "z")
    z, mean, log_std = _f32c(z), _f32c(mean), _f32c(log_std)
    b, c, ty = z.shape
    tx = mean.shape[2]
    out = torch.empty((b, tx, ty), dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        rc = _lib.lib().glow_align_logp(_lib.ptr(z), _lib.ptr(mean), _lib.ptr(log_std), _lib.ptr(t_len), _lib.ptr(m_len),
                                        b, c, tx, ty, tx, ty, _lib.ptr(out), _lib.stream_ptr(z.device))
    _lib.check(rc, "glow_align_logp")
    return out


class _ExpandFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, log_std, frame_token, durations, t_len, m_len, t_y):
        mean, log_std = mean.contiguous(), log_std.contiguous()
        b, c, tx = mean.shape
        dev = mean.device
        mel_mean = torch.empty((b, c, t_y), dtype=torch.float32, device=dev)
        mel_std = torch.empty((b, c, t_y), dtype=torch.float32, device=dev)
        ldt = torch.empty((b, 1, durations.shape[1]), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().glow_align_expand_forward(
                _lib.ptr(mean), _lib.ptr(log_std), _lib.ptr(frame_token), _lib.ptr(durations), _lib.ptr(t_len),
                _lib.ptr(m_len), b, c, durations.shape[1], t_y, tx, _lib.ptr(mel_mean), _lib.ptr(mel_std), _lib.ptr(ldt),
                _lib.stream_ptr(dev))
        _lib.check(rc, "glow_align_expand_forward")
        ctx.save_for_backward(durations, t_len)
        ctx.shape = (b, c, tx, t_y)
        ctx.mark_non_differentiable(ldt)
        return mel_mean, mel_std, ldt

    @staticmethod
    def backward(ctx, g_mean, g_std, _g_ldt):
        durations, t_len = ctx.saved_tensors
        b, c, tx, t_y = ctx.shape
        dev = durations.device
        if g_mean is None:
            g_mean = torch.zeros((b, c, t_y), dtype=torch.float32, device=dev)
        if g_std is None:
            g_std = torch.zeros((b, c, t_y), dtype=torch.float32, device=dev)
        g_mean, g_std = g_mean.contiguous(), g_std.contiguous()
        d_mean = torch.empty((b, c, tx), dtype=torch.float32, device=dev)
        d_std = torch.empty((b, c, tx), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().glow_align_expand_backward(
                _lib.ptr(g_mean), _lib.ptr(g_std), _lib.ptr(durations), _lib.ptr(t_len), b, c, durations.shape[1], t_y, tx,
                _lib.ptr(d_mean), _lib.ptr(d_std), _lib.stream_ptr(dev))
        _lib.check(rc, "glow_align_expand_backward")
        return d_mean, d_std, None, None, None, None, None


def expand_by_path(mean, log_std, frame_token, durations, t_len, m_len, t_y):
    """-> (mel_mean [B,C,t_y], mel_log_std [B,C,t_y], log_dur_targets [B,1,T_x])."""
    _lib.require_cuda(mean, "mean")
    return _ExpandFn.apply(mean.float(), log_std.float(), frame_token, durations, t_len, m_len, int(t_y))


class _MLEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, mean, std, log_dets, lengths, squeeze, mel_dim):
        z, mean, std = z.contiguous(), mean.contiguous(), std.contiguous()
        dev = z.device
        L = _lib.lib()
        ws = torch.empty(int(L.glow_mle_loss_workspace_floats()), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        log_dets = log_dets.to(torch.float32).contiguous()
        lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            rc = L.glow_mle_loss_forward(_lib.ptr(z), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(log_dets), _lib.ptr(lengths),
                                         log_dets.numel(), ctypes.c_size_t(z.numel()), squeeze, mel_dim, _lib.ptr(ws),
                                         _lib.ptr(loss), _lib.stream_ptr(dev))
        _lib.check(rc, "glow_mle_loss_forward")
        ctx.save_for_backward(z, mean, std, ws)
        ctx.batch = log_dets.numel()
        ctx.log_dets_shape = log_dets.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        z, mean, std, ws = ctx.saved_tensors
        dev = z.device
        g = g.to(torch.float32).contiguous()
        dz, dm, ds = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
        dl = torch.empty(ctx.log_dets_shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().glow_mle_loss_backward(_lib.ptr(z), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(g), _lib.ptr(ws),
                                                   ctx.batch, ctypes.c_size_t(z.numel()), _lib.ptr(dz), _lib.ptr(dm),
                                                   _lib.ptr(ds), _lib.ptr(dl), _lib.stream_ptr(dev))
        _lib.check(rc, "glow_mle_loss_backward")
        return dz, dm, ds, dl, None, None, None


def mle_loss(z, mean, std, log_dets, lengths, squeeze, mel_dim):
    _lib.require_cuda(z, "z")
    if not (z.shape == mean.shape == std.shape):
        raise ValueError("mle_loss: z, mean and std must have one shape")
    return _MLEFn.apply(z.float(), mean.float(), std.float(), log_dets, lengths, int(squeeze), int(mel_dim))
