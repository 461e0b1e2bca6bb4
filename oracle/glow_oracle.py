"""oracle/glow_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch fp32) restatement of the reference's flow decoder, relative-position
attention encoder, GlowTTS glue, loss and train step, written functionally over
a reference-layout ``state_dict`` so that autograd on it is also the gradient
oracle for the CUDA backward kernels.  Every function cites the reference
file:line it follows.  The arithmetic underneath is PyTorch's (the reference
has no arithmetic of its own on this path: SURVEY.md 8c) -- so parity is pinned
by executing the real reference modules in this container and committing their
inputs/outputs as tests/golden/*.npz (tools/make_golden.py); this file is
checked against those fixtures in tests/test_oracle_model.py.

Formulations here are deliberately *direct* (banded relative-position sums,
explicit 4x4 group mixing, explicit squeeze index maps) rather than the
reference's pad/view/permute tricks, so the two agree only if the semantics
were understood, not because code was carried over.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from . import mas as _mas


@dataclass
class OracleHP:
    """The Hyper_Parameters.yaml keys the path reads (Hyper_Parameters.yaml:3-57)."""
    mel_dim: int = 80
    mode: str = "Vanilla"            # Vanilla | SE (LUT)
    enc_channels: int = 192
    tokens: int = 35
    prenet_kernel: int = 5
    prenet_dropout: float = 0.5
    prenet_stacks: int = 3
    heads: int = 2
    window: int = 4
    ffn_kernel: int = 3
    ffn_channels: int = 768
    enc_dropout: float = 0.1
    enc_stacks: int = 6
    dp_kernel: int = 3
    dp_channels: int = 256
    dp_stacks: int = 2
    dp_dropout: float = 0.1
    dec_stack: int = 12
    num_squeeze: int = 2
    num_split: int = 4
    wn_channels: int = 192
    wn_layers: int = 4
    wn_kernel: int = 5
    wn_dropout: float = 0.05
    num_speakers: int = 109
    spk_dim: int = 256
    use_cython_alignment: bool = True

    @property
    def se(self):
        return self.mode.upper() == "SE"


# --------------------------------------------------------------------------- #
# small pieces
# --------------------------------------------------------------------------- #
def length_mask(lengths, max_len=None):
    """Modules.py:206-211 Mask_Generate -> [B,1,T] float."""
    max_len = int(max_len if max_len is not None else int(lengths.max()))
    return (torch.arange(max_len, device=lengths.device)[None, :] < lengths[:, None]).unsqueeze(1).float()


def wn_weight(sd, prefix):
    """Old-style torch.nn.utils.weight_norm (Modules.py:766,818,825,833):
    w = g * v / ||v|| with the norm over every dim but 0."""
    g, v = sd[prefix + ".weight_g"], sd[prefix + ".weight_v"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))


def squeeze2(x, mask, n=2):
    """Modules.py:895-907.  out[b, j*C + c, t] = x[b, c, n*t + j]; mask'[t] = mask[n*t + n-1]."""
    b, c, t = x.shape
    t = (t // n) * n
    xs = torch.cat([x[:, :, j:t:n] for j in range(n)], dim=1)
    ms = mask[:, :, n - 1:t:n]
    return xs * ms, ms


def unsqueeze2(x, mask, n=2):
    """Modules.py:914-924, the inverse map; mask repeated n times."""
    b, c, t = x.shape
    c0 = c // n
    out = x.new_zeros(b, c0, t * n)
    for j in range(n):
        out[:, :, j::n] = x[:, j * c0:(j + 1) * c0]
    m = mask.repeat_interleave(n, dim=2)
    return out * m, m


# --------------------------------------------------------------------------- #
# flow decoder (Modules.py:653-924)
# --------------------------------------------------------------------------- #
def actnorm(x, mask, logs, bias, reverse=False):
    """Modules.py:689-694."""
    if reverse:
        return (x - bias) * torch.exp(-logs) * mask, None
    z = (bias + torch.exp(logs) * x) * mask
    return z, logs.sum() * mask.sum(dim=(1, 2))


def actnorm_ddi(x, mask):
    """Modules.py:698-711 data-dependent init -> (logs, bias) as [1,C,1]."""
    denom = mask.sum(dim=(0, 2))
    mean = (x * mask).sum(dim=(0, 2)) / denom
    var = (x * x * mask).sum(dim=(0, 2)) / denom - mean ** 2
    half_log_var = 0.5 * torch.log(torch.clamp_min(var, 1e-7))
    return (-half_log_var).view(1, -1, 1), (-mean * torch.exp(-half_log_var)).view(1, -1, 1)


def group_channel_index(channels, split=4):
    """Modules.py:738-740,754-756 read as an index map: the conv2d mixes, for each
    g in [0, channels/split), the `split` channels
        idx[k, g] = (k // (split/2)) * (channels/2) + g * (split/2) + (k % (split/2))."""
    half = split // 2
    k = torch.arange(split).view(-1, 1)
    g = torch.arange(channels // split).view(1, -1)
    return (k // half) * (channels // 2) + g * half + (k % half)


def inv1x1(x, mask, weight, reverse=False, split=4):
    """Modules.py:727-758 as an explicit per-group 4x4 mix."""
    b, c, t = x.shape
    idx = group_channel_index(c, split).to(x.device)         # [split, groups]
    w = torch.inverse(weight) if reverse else weight          # :743 / :746
    grouped = x[:, idx.reshape(-1)].view(b, split, c // split, t)
    mixed = torch.einsum("ok,bkgt->bogt", w, grouped)
    z = torch.zeros_like(x)
    z[:, idx.reshape(-1)] = mixed.reshape(b, c, t).to(z.dtype)
    z = z * mask
    if reverse:
        return z, None
    return z, torch.logdet(weight) * (c / split) * mask.sum(dim=(1, 2))      # :747


def wavenet(sd, p, x, mask, hp, spk=None, training=False):
    """Modules.py:858-887.  p = '...layers.2.layer_Dict.WaveNet.layer_Dict'."""
    out = torch.zeros_like(x)
    pad = (hp.wn_kernel - 1) // 2
    for i in range(hp.wn_layers):
        ins = F.conv1d(x, wn_weight(sd, f"{p}.In_{i}"), sd[f"{p}.In_{i}.bias"], padding=pad)   # :861
        ins = F.dropout(ins, hp.wn_dropout, training)                                          # :862
        if spk is not None:                                                                     # :863-864
            ins = ins + F.conv1d(spk.unsqueeze(2), wn_weight(sd, f"{p}.Speaker_{i}"), sd[f"{p}.Speaker_{i}.bias"])
        h = hp.wn_channels
        acts = torch.tanh(ins[:, :h]) * torch.sigmoid(ins[:, h:])                              # :885-887
        rs = F.conv1d(acts, wn_weight(sd, f"{p}.Res_Skip_{i}"), sd[f"{p}.Res_Skip_{i}.bias"])  # :871
        if i < hp.wn_layers - 1:
            x = (x + rs[:, :h]) * mask                                                          # :878
            out = out + rs[:, h:]                                                               # :879
        else:
            out = out + rs                                                                      # :881
    return out * mask                                                                           # :883


def coupling(sd, p, x, mask, hp, spk=None, reverse=False, training=False):
    """Modules.py:780-810.  p = '...layers.2.layer_Dict'."""
    half = x.shape[1] // 2
    xa, xb = x[:, :half], x[:, half:]
    h = F.conv1d(xa, wn_weight(sd, f"{p}.Start"), sd[f"{p}.Start.bias"]) * mask                # :791
    h = wavenet(sd, f"{p}.WaveNet.layer_Dict", h, mask, hp, spk, training)                     # :792
    outs = F.conv1d(h, sd[f"{p}.End.weight"], sd[f"{p}.End.bias"])                             # :793
    mean, logs = outs[:, :half], outs[:, half:]
    if reverse:
        xb = (xb - mean) * torch.exp(-logs) * mask                                              # :802
        ld = None
    else:
        xb = (mean + torch.exp(logs) * xb) * mask                                               # :805
        ld = (logs * mask).sum(dim=(1, 2))                                                      # :806
    return torch.cat([xa, xb], 1), ld


def flow_block(sd, p, x, mask, hp, spk=None, reverse=False, training=False):
    """Modules.py:662-668 AIA.  p = 'layer_Dict.Decoder.layer_Dict.Flows.{i}'."""
    if not reverse:
        x, ld0 = actnorm(x, mask, sd[f"{p}.layers.0.logs"], sd[f"{p}.layers.0.bias"])
        x, ld1 = inv1x1(x, mask, sd[f"{p}.layers.1.weight"], split=hp.num_split)
        x, ld2 = coupling(sd, f"{p}.layers.2.layer_Dict", x, mask, hp, spk, False, training)
        return x, ld0 + ld1 + ld2
    x, _ = coupling(sd, f"{p}.layers.2.layer_Dict", x, mask, hp, spk, True, training)
    x, _ = inv1x1(x, mask, sd[f"{p}.layers.1.weight"], reverse=True, split=hp.num_split)
    x, _ = actnorm(x, mask, sd[f"{p}.layers.0.logs"], sd[f"{p}.layers.0.bias"], reverse=True)
    return x, None


def decoder(sd, x, mask, hp, spk=None, reverse=False, training=False,
            prefix="layer_Dict.Decoder.layer_Dict"):
    """Modules.py:298-309.  x [B,80,T], mask [B,1,T] -> (x [B,80,2*(T//2)], logdet [B]|None, mask)."""
    x, m = squeeze2(x, mask, hp.num_squeeze)
    order = range(hp.dec_stack - 1, -1, -1) if reverse else range(hp.dec_stack)
    total = None
    for i in order:
        x, ld = flow_block(sd, f"{prefix}.Flows.{i}", x, m, hp, spk, reverse, training)
        if ld is not None:
            total = ld if total is None else total + ld
    x, mask = unsqueeze2(x, m, hp.num_squeeze)
    return x, total, mask


def decoder_ddi(sd, x, mask, hp, spk=None, prefix="layer_Dict.Decoder.layer_Dict"):
    """Decoder.forward (Modules.py:298-309) on a model whose Activation_Norm layers are all uninitialised: each
    block's first call sets logs / bias from the statistics of ITS input (Modules.py:685-711), i.e. of the previous
    block's output under the freshly initialised parameters.  Writes the new logs / bias into `sd` and returns
    (z, logdet, logs [blocks, C], bias [blocks, C])."""
    with torch.no_grad():
        x, m = squeeze2(x, mask, hp.num_squeeze)
        total, all_logs, all_bias = None, [], []
        for i in range(hp.dec_stack):
            p = f"{prefix}.Flows.{i}"
            logs, bias = actnorm_ddi(x, m)
            sd[f"{p}.layers.0.logs"] = logs
            sd[f"{p}.layers.0.bias"] = bias
            all_logs.append(logs.view(-1))
            all_bias.append(bias.view(-1))
            x, ld = flow_block(sd, p, x, m, hp, spk, False, False)
            total = ld if total is None else total + ld
        x, _ = unsqueeze2(x, m, hp.num_squeeze)
    return x, total, torch.stack(all_logs), torch.stack(all_bias)


# --------------------------------------------------------------------------- #
# relative-position attention (RPR_MHA.py:69-165) -- banded formulation
# --------------------------------------------------------------------------- #
def rel_band(t, window, device=None):
    """[T,T] long index j-i+window inside the band |j-i|<=window, and the band mask.
    RPR_MHA.py:131-165 realise this with zero-padded embeddings + skew views:
    positions outside the window contribute exactly 0 (padded, not clipped)."""
    i = torch.arange(t, device=device).view(-1, 1)
    j = torch.arange(t, device=device).view(1, -1)
    d = j - i
    return (d + window).clamp(0, 2 * window), (d.abs() <= window)


def rpr_attention(sd, p, x, attn_mask, hp, training=False):
    """RPR_MHA.py:69-128.  x [B,C,T]; attn_mask [B,1,T,T] 0/1.  p = '...layer_Dict.Attention'.
    Returns (out [B,C,T], alignments [B,H,T,T])."""
    b, c, t = x.shape
    hds, d = hp.heads, c // hp.heads
    q = F.conv1d(x, sd[f"{p}.layer_Dict.Query.weight"], sd[f"{p}.layer_Dict.Query.bias"])
    k = F.conv1d(x, sd[f"{p}.layer_Dict.Key.weight"], sd[f"{p}.layer_Dict.Key.bias"])
    v = F.conv1d(x, sd[f"{p}.layer_Dict.Value.weight"], sd[f"{p}.layer_Dict.Value.bias"])
    q = q.view(b, hds, d, t).transpose(2, 3)
    k = k.view(b, hds, d, t).transpose(2, 3)
    v = v.view(b, hds, d, t).transpose(2, 3)
    scale = 1.0 / math.sqrt(d)
    scores = (q @ k.transpose(2, 3)) * scale                                   # :103
    idx, band = rel_band(t, hp.window, x.device)
    wk, wv = sd[f"{p}.weight_K"][0], sd[f"{p}.weight_V"][0]                    # [2w+1, d]
    qr = q @ wk.t()                                                             # [B,H,T,2w+1]
    rel_k = torch.gather(qr, 3, idx.expand(b, hds, t, t)) * band               # :106-108
    scores = scores + rel_k * scale                                             # :109
    scores = scores.masked_fill(attn_mask == 0, -1e4)                           # :117
    align = F.softmax(scores, dim=-1)                                           # :119
    align = F.dropout(align, hp.enc_dropout, training)                          # :120
    out = align @ v                                                             # :121
    # :123-126  out[i] += sum_{|j-i|<=w} align[i,j] * wV[j-i+w]
    band_p = align * band
    rel_w = torch.zeros(b, hds, t, 2 * hp.window + 1, dtype=band_p.dtype, device=x.device)
    rel_w.scatter_add_(3, idx.expand(b, hds, t, t), band_p)
    out = out + rel_w @ wv
    out = out.transpose(2, 3).reshape(b, c, t)                                  # :128
    proj = F.conv1d(out, sd[f"{p}.layer_Dict.Projection.weight"], sd[f"{p}.layer_Dict.Projection.bias"])
    return proj, align


def _ln(x, sd, p):
    """LayerNorm over channels of [B,C,T] with eps 1e-4 (Modules.py:472-475,523-526,541-544)."""
    return F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[p + ".weight"], sd[p + ".bias"], 1e-4).transpose(1, 2)


def encoder(sd, tokens, mask, hp, spk=None, training=False, prefix="layer_Dict.Encoder.layer_Dict"):
    """Modules.py:262-284 (+ Prenet :453-459, CLRD :483-489, ANCRDCN :553-573,
    Duration_Predictor :602-618, CRND :642-648)."""
    c = hp.enc_channels
    x = F.embedding(tokens, sd[f"{prefix}.Embedding.weight"]).transpose(1, 2) * math.sqrt(c)   # :267
    # Prenet
    res = x
    for i in range(hp.prenet_stacks):
        p = f"{prefix}.Prenet.layer_Dict.CLRD_{i}.layer_Dict"
        x = F.conv1d(x * mask, sd[f"{p}.Conv.weight"], sd[f"{p}.Conv.bias"], padding=(hp.prenet_kernel - 1) // 2)
        x = F.dropout(F.relu(_ln(x, sd, f"{p}.LayerNorm")), hp.prenet_dropout, training)
    p = f"{prefix}.Prenet.layer_Dict.Conv1x1"
    x = (F.conv1d(x, sd[f"{p}.weight"], sd[f"{p}.bias"]) + res) * mask                          # :457-459
    # Transformer
    amask = (mask * mask.transpose(1, 2)).unsqueeze(1)                                          # :558
    pad = (hp.ffn_kernel - 1) // 2
    for i in range(hp.enc_stacks):
        p = f"{prefix}.Transformer.layer_Dict.ANCRDCN_{i}.layer_Dict"
        x = x * mask                                                                            # :554
        res = x
        a, _ = rpr_attention(sd, f"{p}.Attention", x, amask, hp, training)
        a = F.dropout(a, hp.enc_dropout, training)
        x = _ln(a + res, sd, f"{p}.LayerNorm_0")                                                # :562
        res = x
        y = F.conv1d(x * mask, sd[f"{p}.Conv_0.weight"], sd[f"{p}.Conv_0.bias"], padding=pad)   # :565
        y = F.dropout(F.relu(y), hp.enc_dropout, training)
        y = F.conv1d(y * mask, sd[f"{p}.Conv_1.weight"], sd[f"{p}.Conv_1.bias"], padding=pad)   # :568
        y = F.dropout(y, hp.enc_dropout, training)
        x = _ln(y * mask + res, sd, f"{p}.LayerNorm_1")                                         # :571
    x = x * mask                                                                                # :507
    proj = F.conv1d(x, sd[f"{prefix}.Project.weight"], sd[f"{prefix}.Project.bias"]) * mask     # :271-275
    mean, log_std = proj[:, :hp.mel_dim], proj[:, hp.mel_dim:]
    # Duration predictor on detached trunk (:282)
    d = x.detach()
    if spk is not None:                                                                         # :606-612
        d = torch.cat([d, spk.detach().unsqueeze(2).expand(-1, -1, d.shape[2])], 1)
    for i in range(hp.dp_stacks):
        p = f"{prefix}.Duration_Predictor.layer_Dict.CRND_{i}.layer_Dict.Conv"
        d = F.conv1d(d * mask, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=(hp.dp_kernel - 1) // 2)
        d = F.dropout(F.relu(d), hp.dp_dropout, training)
    p = f"{prefix}.Duration_Predictor.layer_Dict.Projection"
    logw = F.conv1d(d * mask, sd[f"{p}.weight"], sd[f"{p}.bias"]) * mask                        # :616-618
    return mean, log_std, logw, mask


# --------------------------------------------------------------------------- #
# GlowTTS glue, loss, train step
# --------------------------------------------------------------------------- #
def log_prior(z, mean, log_std):
    """Modules.py:108-114 -> log_P [B,T_x,T_y]."""
    r = torch.exp(-2 * log_std)
    return ((-0.5 * math.log(2 * math.pi) - log_std).sum(1).unsqueeze(-1)
            + r.transpose(2, 1) @ (-0.5 * z ** 2)
            + (mean * r).transpose(2, 1) @ z
            + (-0.5 * mean ** 2 * r).sum(1).unsqueeze(-1))


def glow_forward(sd, hp, tokens, token_lengths, mels, mel_lengths, speakers=None,
                 training=False, mas_core="port"):
    """Modules.py:50-126 (Vanilla / SE-LUT).  Returns the reference's 8-tuple."""
    assert bool((mel_lengths % hp.num_squeeze == 0).all())                                      # :71
    spk = F.embedding(speakers, sd["layer_Dict.LUT.weight"]) if hp.se else None                 # :73-74
    tmask = length_mask(token_lengths)
    mmask = length_mask(mel_lengths)
    mean, log_std, logw, tmask = encoder(sd, tokens, tmask, hp, spk, training)
    z, logdet, mmask = decoder(sd, mels, mmask, hp, spk, False, training)
    amask = (tmask.unsqueeze(-1) * mmask.unsqueeze(2)).squeeze(1)                               # :102-103
    with torch.no_grad():
        logp = log_prior(z, mean, log_std)
        # the reference's wrapper leaves the device here (monotonic_align/__init__.py:14-21): D2H, C loop, H2D
        path = _mas.maximum_path_numpy(logp.float().cpu().numpy(), amask.float().cpu().numpy(), core=mas_core)   # :116
        attn = torch.from_numpy(path).to(device=logp.device, dtype=mean.dtype)
    mel_mean = mean @ attn                                                                      # :120
    mel_log_std = log_std @ attn                                                                # :121
    logw_target = torch.log(attn.unsqueeze(1).sum(-1) + 1e-7) * tmask                           # :122
    return z, mel_mean, mel_log_std, logdet, logw, logw_target, attn, None


def mle_loss(z, mean, log_std, logdet, lengths, hp):
    """Modules.py:1020-1029."""
    loss = log_std.sum() + 0.5 * (torch.exp(-2 * log_std) * (z - mean) ** 2).sum() - logdet.sum()
    loss = loss / ((lengths // hp.num_squeeze).sum() * hp.num_squeeze * hp.mel_dim)
    return loss + 0.5 * math.log(2 * math.pi)


def losses(out, mel_lengths, hp):
    """Train.py:203-211 -> (total, mle, length)."""
    z, mel_mean, mel_log_std, logdet, logw, logw_t = out[:6]
    mle = mle_loss(z, mel_mean, mel_log_std, logdet, mel_lengths, hp)
    length = F.mse_loss(logw, logw_t)
    return mle + length, mle, length


class RAdamOracle:
    """Radam.py:25-90 restated on a list of tensors (lr, betas, eps, weight decay
    as Train.py:162-168) with the Modified Noam schedule of Noam_Scheduler.py:17-29."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-6, base=4000):
        self.params = list(params)
        self.lr0, self.betas, self.eps, self.wd, self.base = lr, betas, eps, weight_decay, base
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0          # optimizer step count
        self.epoch = 0      # scheduler last_epoch (starts at 0 => scale uses max(1, .))

    def lr(self):
        e = max(1, self.epoch)
        return self.lr0 * self.base ** 0.5 * (e + self.base) ** -0.5

    @torch.no_grad()
    def step(self):
        self.t += 1
        b1, b2 = self.betas
        lr = self.lr()
        b2t = b2 ** self.t
        n_max = 2 / (1 - b2) - 1
        n_sma = n_max - 2 * self.t * b2t / (1 - b2t)
        if n_sma >= 5:
            step_size = math.sqrt((1 - b2t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma
                                  * n_max / (n_max - 2)) / (1 - b1 ** self.t)
        else:
            step_size = 1.0 / (1 - b1 ** self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            m.mul_(b1).add_(g, alpha=1 - b1)
            if self.wd != 0:
                p.add_(p, alpha=-self.wd * lr)
            if n_sma >= 5:
                p.addcdiv_(m, v.sqrt().add_(self.eps), value=-step_size * lr)
            else:
                p.add_(m, alpha=-step_size * lr)
        self.epoch += 1     # scheduler.step() (Train.py:233)


def train_step(sd, hp, opt, batch, training=True, mas_core="port", clip=5.0):
    """Train.py:182-233 around the functional model.  sd values must be leaf tensors
    with requires_grad; opt is a RAdamOracle over them.  Returns (total, mle, length)."""
    tokens, token_lengths, mels, mel_lengths, speakers = batch
    out = glow_forward(sd, hp, tokens, token_lengths, mels, mel_lengths,
                       speakers if hp.se else None, training, mas_core)
    total, mle, length = losses(out, mel_lengths, hp)
    for p in opt.params:
        p.grad = None
    total.backward()
    torch.nn.utils.clip_grad_norm_([p for p in opt.params if p.grad is not None], clip)   # Train.py:227-231
    opt.step()
    return float(total.detach()), float(mle.detach()), float(length.detach())


def state_dict_to_leaves(sd):
    """Detach + clone every float entry into a leaf that requires grad."""
    return {k: (v.detach().clone().float().requires_grad_(True) if v.is_floating_point() else v.clone())
            for k, v in sd.items()}
