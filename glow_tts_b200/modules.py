"""Drop-in module surface of the reference's Modules.py, backed by libglowcore.

Same class names, zero-argument constructors reading ``hp`` (the parsed
Hyper_Parameters.yaml), identical ``state_dict`` key layout, so a
reference checkpoint loads with ``load_state_dict(strict=True)`` and
``Train.py`` / ``Inference.py`` can drive it unchanged.  What differs is where
the arithmetic happens:

* ``Decoder``              -> glow_flow_forward / _reverse / _backward (csrc/flow_*.cu)
* ``RPR_Multihead_Attention`` core -> glow_rpr_attention_* (csrc/attention.cu)
* ``Maximum_Path_Generater``  -> glow_mas_forward (csrc/mas.cu)

* ``Encoder`` convolutions / LayerNorm (bf16 mode) -> packed-row tcgen05 convs and fused LayerNorm kernels
  (rows.py, csrc/rows_conv.cu, csrc/rows_norm.cu); the fp32 parity mode keeps the torch ops.
* ``log_P`` / path expansion / ``MLE_Loss`` -> csrc/align.cu

There is no CPU fallback: tensors must be on a CUDA device.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib, align as _align, flow as _flow, rows as _rows
from .flat import FlatBuffer, offset_table
from .hparams import load_hparams
from .monotonic_align import maximum_path, maximum_path_align
from .rpr_mha import RPR_Multihead_Attention

hp = None


def set_hparams(new_hp):
    """Install the hyper-parameter namespace the zero-arg constructors read."""
    global hp
    hp = new_hp
    return hp


def _hp():
    global hp
    if hp is None:
        hp = load_hparams()          # ./Hyper_Parameters.yaml if present, like Modules.py:10-13
    return hp


def _se_mode():
    return _hp().Mode.upper() in ("SE", "GR")


# --------------------------------------------------------------------------- #
# parameter containers
# --------------------------------------------------------------------------- #
def _xavier_(w, gain):
    torch.nn.init.xavier_uniform_(w, gain=gain)


def _init_conv_weight(weight, gains):
    """Per-chunk init of the reference's Conv1d subclass (Modules.py:988-1003)."""
    if isinstance(gains, str):
        gains = [gains]
    for gain, chunk in zip(gains, torch.chunk(weight, len(gains), dim=0)):
        if gain == "zero":
            torch.nn.init.zeros_(chunk)
        elif gain in ("relu", "leaky_relu"):
            torch.nn.init.kaiming_uniform_(chunk, nonlinearity=gain)
        else:
            _xavier_(chunk, torch.nn.init.calculate_gain(gain))


class WNConv1d(torch.nn.Module):
    """Weight-normalised conv parameters with the reference's (old-style
    torch.nn.utils.weight_norm) names and registration order: bias, weight_g,
    weight_v (Modules.py:766,818,825,833)."""

    def __init__(self, in_channels, out_channels, kernel_size, gains):
        super().__init__()
        w = torch.empty(out_channels, in_channels, kernel_size)
        _init_conv_weight(w, gains)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        self.weight_g = torch.nn.Parameter(w.flatten(1).norm(dim=1).view(-1, 1, 1).clone())
        self.weight_v = torch.nn.Parameter(w)
        self.kernel_size, self.in_channels, self.out_channels = kernel_size, in_channels, out_channels

    def effective_weight(self):
        v = self.weight_v
        return v * (self.weight_g / v.flatten(1).norm(dim=1).view(-1, 1, 1))


class Activation_Norm(torch.nn.Module):
    """Parameters of Modules.py:670-711; arithmetic fused into the flow kernels."""

    def __init__(self):
        super().__init__()
        h = _hp()
        c = h.Sound.Mel_Dim * h.Decoder.Num_Squeeze
        self.initialized = False          # plain attribute, as in the reference (not in state_dict)
        self.logs = torch.nn.Parameter(torch.zeros(1, c, 1))
        self.bias = torch.nn.Parameter(torch.zeros(1, c, 1))


class Invertible_1x1_Conv(torch.nn.Module):
    """Parameters of Modules.py:713-758: a Num_Split x Num_Split orthogonal init with det > 0."""

    def __init__(self):
        super().__init__()
        n = _hp().Decoder.Num_Split
        assert n % 2 == 0
        q = torch.linalg.qr(torch.randn(n, n))[0]
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        self.weight = torch.nn.Parameter(q.contiguous())


class WaveNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        h = _hp()
        ch = h.Decoder.Affine_Coupling.Calc_Channels
        wn = h.Decoder.Affine_Coupling.WaveNet
        self.layer_Dict = torch.nn.ModuleDict()
        for i in range(wn.Num_Layers):
            self.layer_Dict["In_%d" % i] = WNConv1d(ch, ch * 2, wn.Kernel_Size, ["tanh", "sigmoid"])
            self.layer_Dict["Res_Skip_%d" % i] = WNConv1d(
                ch, ch * (2 if i < wn.Num_Layers - 1 else 1), 1, "linear")
            if _se_mode():
                self.layer_Dict["Speaker_%d" % i] = WNConv1d(
                    h.Speaker_Embedding.Embedding_Size, ch * 2, 1, ["tanh", "sigmoid"])
        self.layer_Dict["Dropout"] = torch.nn.Dropout(p=wn.Dropout_Rate)


class Affine_Coupling_Layer(torch.nn.Module):
    def __init__(self):
        super().__init__()
        h = _hp()
        c = h.Sound.Mel_Dim * h.Decoder.Num_Squeeze
        ch = h.Decoder.Affine_Coupling.Calc_Channels
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Start"] = WNConv1d(c // 2, ch, 1, "linear")
        self.layer_Dict["WaveNet"] = WaveNet()
        end = torch.nn.Conv1d(ch, c, 1)
        torch.nn.init.zeros_(end.weight)          # Modules.py:773-778 'zero'
        torch.nn.init.zeros_(end.bias)
        self.layer_Dict["End"] = end


class AIA(torch.nn.Module):
    """One Glow block: ActNorm -> invertible conv -> affine coupling (Modules.py:653-668)."""

    def __init__(self):
        super().__init__()
        self.layers = torch.nn.ModuleList([Activation_Norm(), Invertible_1x1_Conv(), Affine_Coupling_Layer()])

    def slot_params(self):
        """Parameters in the slot order of include/glowcore.h (glow_flow_* parameter table)."""
        an, inv, acl = self.layers
        out = [an.logs, an.bias, inv.weight]
        st = acl.layer_Dict["Start"]
        out += [st.bias, st.weight_g, st.weight_v]
        wn = acl.layer_Dict["WaveNet"].layer_Dict
        i = 0
        while "In_%d" % i in wn:
            for name in ("In_%d", "Res_Skip_%d", "Speaker_%d"):
                if name % i in wn:
                    m = wn[name % i]
                    out += [m.bias, m.weight_g, m.weight_v]
            i += 1
        end = acl.layer_Dict["End"]
        out += [end.weight, end.bias]
        return out


class Squeeze(torch.nn.Module):
    """Kept for the surface (Modules.py:890-907); the fused kernels do the index map."""

    def __init__(self, num_squeeze=2):
        super().__init__()
        self.num_Squeeze = num_squeeze


class Unsqueeze(torch.nn.Module):
    def __init__(self, num_squeeze=2):
        super().__init__()
        self.num_Squeeze = num_squeeze


def _host_lengths(mask=None, lengths=None):
    """Valid frame count per utterance as host ints.  `lengths` on the CPU costs nothing;
    a CUDA mask / lengths costs one D2H sync (the reference pays several per forward)."""
    if lengths is not None:
        return [int(v) for v in lengths.detach().cpu().tolist()]
    return [int(v) for v in mask.detach().sum(dim=(1, 2)).round().long().cpu().tolist()]


class Decoder(torch.nn.Module):
    """Modules.py:286-309.  forward(x [B,80,T], mask [B,1,T], speakers [B,256]|None, ...,
    reverse=False) -> (x [B,80,2*(T//2)], log_Dets [B] | None, mask)."""

    def __init__(self):
        super().__init__()
        h = _hp()
        if h.Decoder.Num_Squeeze != 2:
            raise _lib.GlowCoreError("the sm_100a flow kernels are built for Num_Squeeze=2")
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Squeeze"] = Squeeze(h.Decoder.Num_Squeeze)
        self.layer_Dict["Unsqueeze"] = Unsqueeze(h.Decoder.Num_Squeeze)
        self.layer_Dict["Flows"] = torch.nn.ModuleList([AIA() for _ in range(h.Decoder.Stack)])
        self.precision = getattr(h, "Precision", "bf16")
        self.mel_dim = h.Sound.Mel_Dim
        self.spk_dim = h.Speaker_Embedding.Embedding_Size if _se_mode() else 0
        self.dropout = float(h.Decoder.Affine_Coupling.WaveNet.Dropout_Rate)
        self._plans = {}
        self._flat = None
        self._offs = None
        self._step = 0
        self.host_lengths = None       # optional: set by the caller to skip the D2H sync
        self._prepared = None          # (wpack, wpack_tc) being prepared on the side stream (begin_prepare)
        self.device_lengths = None     # set around a call by GlowTTS.inference_device: mel lengths as a device tensor
        self.geometry = None           # set around a call by GlowTTS.forward: geometry.StepGeometry (static row map)
        self._dev_maps = {}
        self.defer_param_grads = False  # set by train.TrainStep, which joins the side stream before the optimizer

    # ---- flat parameter plumbing (see flat.py) ------------------------------------
    def slot_params(self):
        out = []
        for blk in self.layer_Dict["Flows"]:
            out += blk.slot_params()
        return out

    def adopt_flat(self, flat):
        self._flat, self._offs = flat, None

    def flat_params(self):
        if self._flat is None or not self._flat.valid():
            self._flat, self._offs = FlatBuffer(self.slot_params()), None
        if self._offs is None:
            self._offs = offset_table(self._flat, self.slot_params())
        return self._flat.data, self._offs

    def flat_grads(self):
        """(flat gradient buffer the kernels accumulate into, True if it IS the .grad storage)."""
        fb = self._flat
        params = self.slot_params()
        if fb.grads_attached(params):
            return fb.grad, True
        return torch.zeros_like(fb.data), False

    def split_grads(self, gflat):
        fb = self._flat
        return [gflat[fb.offset_of(p):fb.offset_of(p) + p.numel()].view(p.shape) for p in self.slot_params()]

    @property
    def plan(self):
        return self.plan_for(len(self.layer_Dict["Flows"]))

    def plan_for(self, blocks):
        if blocks not in self._plans:
            self._plans[blocks] = _flow.FlowPlan(blocks, self.spk_dim, self.dropout)
        return self._plans[blocks]

    def begin_prepare(self, device):
        """Start this step's weight preparation on the side stream (it only depends on the parameters);
        the next forward() picks the result up.  Skipped while ActNorm still needs its data-dependent init."""
        if self._prepared is not None or not all(blk.layers[0].initialized for blk in self.layer_Dict["Flows"]):
            return
        tag, _ = _flow.precision_tag(self.precision)
        flat, offs = self.flat_params()
        cur, side = torch.cuda.current_stream(device), _flow.side_stream(device)
        side.wait_stream(cur)
        with torch.cuda.device(device), torch.cuda.stream(side):
            self._prepared = self.plan.prepare(flat, offs, device, tag)
        _flow._SIDE_BUSY.add(str(torch.device(device)))

    # ---- ActNorm data-dependent init (Modules.py:685-711) -----------------------------
    @torch.no_grad()
    def _data_dependent_init(self, x, rm, speakers):
        """Block k's ActNorm statistics are those of block k-1's raw output, so the decoder is walked ONCE, block
        by block, on packed rows: glow_actnorm_stats -> logs / bias (Modules.py:698-711) -> glow_flow_prepare ->
        glow_flow_block_forward.  Under torch.distributed the three sums are all-reduced, so every rank gets the
        same parameters (the reference has no DDP; per-rank statistics would make the replicas diverge)."""
        import ctypes
        flows = self.layer_Dict["Flows"]
        dev = x.device
        plan = self.plan
        tag, act_dtype = _flow.precision_tag(self.precision)
        flat, offs = self.flat_params()
        L = _lib.lib()
        c = 2 * self.mel_dim
        with torch.cuda.device(dev):
            ws = plan.workspace(rm, dev, act_dtype, False)
            spk = speakers.contiguous().float() if speakers is not None else None
            xc = x.contiguous().float()
            cur = torch.zeros((rm.rows_pad, c), dtype=torch.float32, device=dev)
            nxt = torch.zeros_like(cur)
            sums = torch.empty(3 * c, dtype=torch.float32, device=dev)
            wp, wtc = plan.prepare(flat, offs, dev, tag)
            call = plan.call_struct(rm, xc.shape[2], tag, wp, wtc, spk, ws, False, 0, dev)
            _lib.check(L.glow_flow_pack_rows(ctypes.byref(call), _lib.ptr(xc), _lib.ptr(cur)), "glow_flow_pack_rows")
            for k, blk in enumerate(flows):
                an = blk.layers[0]
                if not an.initialized:
                    _lib.check(L.glow_actnorm_stats(_lib.ptr(cur), rm.row_utt.data_ptr(), rm.rows_pad, c, _lib.ptr(sums),
                                                    _lib.stream_ptr(dev)), "glow_actnorm_stats")
                    s3 = sums.view(3, c).double()
                    if torch.distributed.is_available() and torch.distributed.is_initialized():
                        torch.distributed.all_reduce(s3)
                    mean = s3[1] / s3[0]
                    var = s3[2] / s3[0] - mean ** 2
                    half_log_var = 0.5 * torch.log(torch.clamp_min(var, 1e-7))
                    an.logs.data.copy_((-half_log_var).view_as(an.logs))
                    an.bias.data.copy_((-mean * torch.exp(-half_log_var)).view_as(an.bias))
                    an.initialized = True
                    wp, wtc = plan.prepare(flat, offs, dev, tag)      # exp(logs), bias of block k -> packed weights
                if k + 1 < len(flows) and not all(b.layers[0].initialized for b in flows[k + 1:]):
                    call.stream = torch.cuda.current_stream(dev).cuda_stream
                    _lib.check(L.glow_flow_block_forward(ctypes.byref(call), k, _lib.ptr(cur), _lib.ptr(nxt)),
                               "glow_flow_block_forward")
                    cur, nxt = nxt, cur

    def _device_row_map(self, batch, sq_max, device):
        key = (batch, sq_max, str(device))
        if key not in self._dev_maps:
            if len(self._dev_maps) > 16:
                self._dev_maps.clear()
            self._dev_maps[key] = _flow.DeviceRowMap(batch, sq_max, device)
        return _lib.keepalive(self._dev_maps[key])

    # ---- forward -----------------------------------------------------------------------
    def forward(self, x, mask, speakers=None, prosodies=None, pitches=None, reverse=False):
        _lib.require_cuda(x, "Decoder input")
        if prosodies is not None or pitches is not None:
            raise _lib.GlowCoreError("PE/GR conditioning is outside the accelerated path (SURVEY 2: out of scope)")
        if (self.spk_dim > 0) != (speakers is not None):
            raise ValueError("speaker embeddings must be given exactly in SE mode")
        b, c, t = x.shape
        t2 = (t // 2) * 2
        xin = x[:, :, :t2]
        if self.device_lengths is not None:
            # lengths known on the device only (sync-free inference): fixed-geometry row map, validity per row
            # decided on the device
            if not reverse:
                raise _lib.GlowCoreError("device_lengths is for the reverse (inference) direction")
            sq_dev = torch.clamp(self.device_lengths.to(torch.int64), max=t2) // 2
            rm = self._device_row_map(b, t2 // 2, x.device).update(sq_dev)
            out_mask = (torch.arange(t2, device=x.device)[None, None, :] < (2 * sq_dev)[:, None, None]).to(x.dtype)
        elif self.geometry is not None:
            # bucketed training (geometry.py): the row map lives in fixed device buffers refreshed before every step
            geo = self.geometry
            if reverse or t2 != geo.t_mel:
                raise _lib.GlowCoreError("StepGeometry is for the forward direction on [B, 80, %d] inputs" % geo.t_mel)
            rm = geo.dec_rm
            out_mask = (torch.arange(t2, device=x.device)[None, None, :] < (2 * geo.sq64)[:, None, None]).to(x.dtype)
        else:
            lens = self.host_lengths if self.host_lengths is not None else _host_lengths(mask=mask)
            sq_len = [min(n, t2) // 2 for n in lens]
            # reference: squeezed mask = mask[:, :, 1::2] (Modules.py:903): frame pair kept iff its odd frame is valid
            rm = _flow.row_map(sq_len, x.device)
            out_mask = (torch.arange(t2, device=x.device)[None, None, :] <
                        (2 * _lib.device_ints(sq_len, torch.int64, x.device))[:, None, None]).to(x.dtype)
        self.flat_params()
        if reverse:
            with torch.no_grad():
                y = _flow.flow_reverse(self, rm, xin, speakers, 0.0)
            return y.to(x.dtype), None, out_mask
        if not all(blk.layers[0].initialized for blk in self.layer_Dict["Flows"]):
            self._data_dependent_init(xin.float(), rm, speakers)
        seed = 0
        if self.training and self.dropout > 0:
            self._step += 1
            seed = (int(torch.initial_seed()) * 0x9E3779B1 + self._step * 0x85EBCA77) & 0x7FFFFFFFFFFFFFFF or 1
        z, logdet = _flow.FlowDecoderFn.apply(self, rm, xin, speakers, seed, *self.slot_params())
        return z.to(x.dtype), logdet, out_mask


# --------------------------------------------------------------------------- #
# text encoder (Modules.py:232-284, 438-648)
# --------------------------------------------------------------------------- #
class CLRD(torch.nn.Module):
    """Conv -> LayerNorm -> ReLU -> Dropout (Modules.py:461-489)."""

    def __init__(self):
        super().__init__()
        h = _hp().Encoder
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Conv"] = torch.nn.Conv1d(h.Channels, h.Channels, h.Prenet.Kernel_Size,
                                                  padding=(h.Prenet.Kernel_Size - 1) // 2)
        self.layer_Dict["LayerNorm"] = torch.nn.LayerNorm(h.Channels, eps=1e-4)
        self.layer_Dict["ReLU"] = torch.nn.ReLU(inplace=True)
        self.layer_Dict["Dropout"] = torch.nn.Dropout(p=h.Prenet.Dropout_Rate)

    def forward(self, x, mask):
        y = self.layer_Dict["Conv"](x * mask)
        y = self.layer_Dict["LayerNorm"](y.transpose(1, 2)).transpose(1, 2)
        return self.layer_Dict["Dropout"](self.layer_Dict["ReLU"](y))


class Prenet(torch.nn.Module):
    def __init__(self, stacks):
        super().__init__()
        self.stacks = stacks
        ch = _hp().Encoder.Channels
        self.layer_Dict = torch.nn.ModuleDict()
        for i in range(stacks):
            self.layer_Dict["CLRD_%d" % i] = CLRD()
        self.layer_Dict["Conv1x1"] = torch.nn.Conv1d(ch, ch, 1)

    def forward(self, x, mask):
        y = x
        for i in range(self.stacks):
            y = self.layer_Dict["CLRD_%d" % i](y, mask)
        return (self.layer_Dict["Conv1x1"](y) + x) * mask


class ANCRDCN(torch.nn.Module):
    """Attention -> Norm -> Conv -> ReLU -> Dropout -> Conv -> Norm (Modules.py:509-573)."""

    def __init__(self):
        super().__init__()
        h = _hp().Encoder
        tr = h.Transformer
        pad = (tr.Conv.Kernel_Size - 1) // 2
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Attention"] = RPR_Multihead_Attention(
            query_channels=h.Channels, calc_channels=h.Channels, out_channels=h.Channels,
            num_heads=tr.Attention.Heads, relative_postion_clipping_distance=tr.Attention.Window_Size,
            dropout_rate=tr.Dropout_Rate)
        self.layer_Dict["LayerNorm_0"] = torch.nn.LayerNorm(h.Channels, eps=1e-4)
        self.layer_Dict["Conv_0"] = torch.nn.Conv1d(h.Channels, tr.Conv.Calc_Channels, tr.Conv.Kernel_Size, padding=pad)
        self.layer_Dict["Conv_1"] = torch.nn.Conv1d(tr.Conv.Calc_Channels, h.Channels, tr.Conv.Kernel_Size, padding=pad)
        self.layer_Dict["LayerNorm_1"] = torch.nn.LayerNorm(h.Channels, eps=1e-4)
        self.layer_Dict["ReLU"] = torch.nn.ReLU(inplace=True)
        self.layer_Dict["Dropout"] = torch.nn.Dropout(p=tr.Dropout_Rate)

    def forward(self, x, mask, lengths=None):
        d = self.layer_Dict
        x = x * mask                       # the reference does this in place (Modules.py:554)
        a, _ = d["Attention"](queries=x, masks=None if lengths is not None else
                              (mask * mask.transpose(2, 1)).unsqueeze(1), lengths=lengths,
                              need_alignments=False)
        y = d["LayerNorm_0"]((d["Dropout"](a) + x).transpose(1, 2)).transpose(1, 2)
        f = d["Dropout"](d["ReLU"](d["Conv_0"](y * mask)))
        f = d["Dropout"](d["Conv_1"](f * mask))
        return d["LayerNorm_1"]((f * mask + y).transpose(1, 2)).transpose(1, 2)


class Transformer(torch.nn.Module):
    def __init__(self, stacks):
        super().__init__()
        self.stacks = stacks
        self.layer_Dict = torch.nn.ModuleDict()
        for i in range(stacks):
            self.layer_Dict["ANCRDCN_%d" % i] = ANCRDCN()

    def forward(self, x, mask, lengths=None):
        for i in range(self.stacks):
            x = self.layer_Dict["ANCRDCN_%d" % i](x, mask, lengths)
        return x * mask


class CRND(torch.nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        h = _hp().Encoder.Duration_Predictor
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Conv"] = torch.nn.Conv1d(in_channels, h.Channels, h.Kernel_Size,
                                                  padding=(h.Kernel_Size - 1) // 2)
        self.layer_Dict["ReLU"] = torch.nn.ReLU(inplace=True)
        self.layer_Dict["Dropout"] = torch.nn.Dropout(p=h.Dropout_Rate)

    def forward(self, x, mask):
        return self.layer_Dict["Dropout"](self.layer_Dict["ReLU"](self.layer_Dict["Conv"](x * mask)))


class Duration_Predictor(torch.nn.Module):
    def __init__(self):
        super().__init__()
        h = _hp()
        dp = h.Encoder.Duration_Predictor
        ch = h.Encoder.Channels + (h.Speaker_Embedding.Embedding_Size if h.Mode.upper() == "SE" else 0)
        self.stacks = dp.Stacks
        self.layer_Dict = torch.nn.ModuleDict()
        for i in range(dp.Stacks):
            self.layer_Dict["CRND_%d" % i] = CRND(ch)
            ch = dp.Channels
        self.layer_Dict["Projection"] = torch.nn.Conv1d(ch, 1, 1)

    def forward(self, x, x_mask, speakers=None, prosodies=None):
        if speakers is not None:
            x = torch.cat([x, speakers.unsqueeze(2).expand(-1, -1, x.size(2))], dim=1)
        for i in range(self.stacks):
            x = self.layer_Dict["CRND_%d" % i](x, x_mask)
        return self.layer_Dict["Projection"](x * x_mask) * x_mask


class Encoder(torch.nn.Module):
    """Modules.py:232-284: (tokens [B,T], mask [B,1,T], speakers) ->
    (mean [B,80,T], log_Std [B,80,T], log_Durations [B,1,T], mask)."""

    def __init__(self):
        super().__init__()
        h = _hp()
        e = h.Encoder
        self.channels, self.mel_dim = e.Channels, h.Sound.Mel_Dim
        self.precision = getattr(h, "Precision", "bf16")
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Embedding"] = torch.nn.Embedding(e.Embedding_Tokens, e.Channels)
        torch.nn.init.normal_(self.layer_Dict["Embedding"].weight, mean=0.0, std=e.Channels ** -0.5)
        self.layer_Dict["Prenet"] = Prenet(e.Prenet.Stacks)
        self.layer_Dict["Transformer"] = Transformer(e.Transformer.Stacks)
        self.layer_Dict["Project"] = torch.nn.Conv1d(e.Channels, h.Sound.Mel_Dim * 2, 1)
        self.layer_Dict["Duration_Predictor"] = Duration_Predictor()

    def forward(self, x, mask, speakers=None, prosodies=None, lengths=None, host_lengths=None, token_rows=None):
        _lib.require_cuda(mask, "Encoder mask")
        # bf16 mode with the sentence lengths known on the host: packed-row encoder on the tcgen05 convs
        if self.precision == "bf16" and (host_lengths is not None or token_rows is not None) and self.rows_supported:
            return self._forward_rows(x, mask, speakers, lengths, host_lengths, token_rows)
        # fp32 mode is the parity mode: keep cuDNN from silently using TF32 for the convs
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=not self.precision.startswith("fp32")):
            return self._forward(x, mask, speakers, lengths)

    @property
    def rows_supported(self):
        """csrc/rows_conv.cu is built for the reference configuration (Hyper_Parameters.yaml:21-35)."""
        h = _hp().Encoder
        return (h.Channels == 192 and h.Prenet.Kernel_Size == 5 and h.Transformer.Conv.Kernel_Size == 3
                and h.Transformer.Conv.Calc_Channels == 768 and self.mel_dim == 80)

    def _packed_weights(self, dev):
        pk = getattr(self, "_pk", None)
        if pk is None or pk.buf.device != dev:
            convs = [m for m in self.layer_Dict["Prenet"].modules() if isinstance(m, torch.nn.Conv1d)]
            convs += [m for m in self.layer_Dict["Transformer"].modules() if isinstance(m, torch.nn.Conv1d)]
            convs.append(self.layer_Dict["Project"])
            pk = self._pk = _rows.PackedWeights(convs, dev)
        return pk

    def _site_seed(self, site):
        """Dropout stream of one call site of this step (0 in eval); csrc kernels mix in the device step counter."""
        if not self.training:
            return 0
        return ((int(torch.initial_seed()) * 0x9E3779B1 + self._calls * 0x85EBCA77 + site * 0xC2B2AE3D + 1)
                & 0x7FFFFFFFFFFFFFFF) or 1

    def _forward_rows(self, tokens, mask, speakers, lengths, host_lengths, token_rows=None):
        """Modules.py:262-284 on packed token rows [rows, C] (rows.py): same arithmetic at every real
        token; guard rows stand in for the padding the reference masks away.  Every op writes guard
        rows as zeros, so convs need no separate `* mask` pass."""
        d = self.layer_Dict
        dev = tokens.device
        h = _hp().Encoder
        self._calls = getattr(self, "_calls", 0) + 1
        tr = token_rows if token_rows is not None else _rows.token_rows(host_lengths, tokens.shape[1], dev)
        pk = self._packed_weights(dev)
        pk.pack()                                  # every conv weight of the encoder -> bf16 slab images, one launch
        tok = tokens.reshape(-1).index_select(0, tr.src_idx)
        x = d["Embedding"](tok) * (math.sqrt(self.channels) * tr.valid)
        pre = d["Prenet"]
        p_pre = float(h.Prenet.Dropout_Rate)
        y = x
        site = 0
        for i in range(pre.stacks):
            c = pre.layer_Dict["CLRD_%d" % i].layer_Dict
            y = _rows.rows_conv(y, c["Conv"], tr, x_masked=True, packed=pk)
            site += 1
            y = _rows.rows_norm(y, None, c["LayerNorm"], tr, relu=True, p_out=p_pre, seed_out=self._site_seed(site))
        x = _rows.rows_conv(y, pre.layer_Dict["Conv1x1"], tr, x_masked=True, packed=pk) + x
        tf = d["Transformer"]
        p_tf = float(h.Transformer.Dropout_Rate)
        for i in range(tf.stacks):
            b = tf.layer_Dict["ANCRDCN_%d" % i].layer_Dict
            a = b["Attention"].forward_rows(x, tr, lengths, pk)
            y = _rows.rows_norm(a, x, b["LayerNorm_0"], tr, p_in=p_tf, seed_in=self._site_seed(site + 1))
            f = _rows.rows_conv(y, b["Conv_0"], tr, relu=True, p=p_tf, seed=self._site_seed(site + 2), x_masked=True,
                                packed=pk)
            f = _rows.rows_conv(f, b["Conv_1"], tr, x_masked=True, packed=pk)
            x = _rows.rows_norm(f, y, b["LayerNorm_1"], tr, p_in=p_tf, seed_in=self._site_seed(site + 3))
            site += 3
        ms = tr.unpack(_rows.rows_conv(x, d["Project"], tr, x_masked=True, packed=pk)).transpose(1, 2)              # [B, 160, T]
        mean, log_std = torch.split(ms, [self.mel_dim, self.mel_dim], dim=1)
        xd = tr.unpack(x.detach()).transpose(1, 2)                                        # == (x * mask).detach()
        spk = speakers.detach() if speakers is not None else None
        log_dur = d["Duration_Predictor"](xd, mask, spk, None)
        return mean, log_std, log_dur, mask

    def _forward(self, x, mask, speakers, lengths):
        d = self.layer_Dict
        y = d["Embedding"](x).transpose(2, 1) * math.sqrt(self.channels)
        y = d["Prenet"](y, mask)
        y = d["Transformer"](y, mask, lengths)
        mean, log_std = torch.split(d["Project"](y) * mask, [self.mel_dim, self.mel_dim], dim=1)
        spk = speakers.detach() if speakers is not None else None
        log_dur = d["Duration_Predictor"](y.detach(), mask, spk, None)
        return mean, log_std, log_dur, mask


class Maximum_Path_Generater(torch.nn.Module):
    """Modules.py:927-980 / monotonic_align.maximum_path: same (log_p, mask) call."""

    def forward(self, log_p, mask=None, t_x=None, t_y=None):
        return maximum_path(log_p, mask, t_x, t_y)


class MLE_Loss(torch.nn.modules.loss._Loss):
    """Modules.py:1020-1029."""

    def forward(self, z, mean, std, log_dets, lengths):
        h = _hp()
        if FUSED_ALIGN and z.is_cuda and z.shape == mean.shape == std.shape:
            return _align.mle_loss(z, mean, std, log_dets, lengths, h.Decoder.Num_Squeeze, h.Sound.Mel_Dim)
        loss = torch.sum(std) + 0.5 * torch.sum(torch.exp(-2 * std) * (z - mean) ** 2) - torch.sum(log_dets)
        loss = loss / (torch.sum(lengths // h.Decoder.Num_Squeeze) * h.Decoder.Num_Squeeze * h.Sound.Mel_Dim)
        return loss + 0.5 * math.log(2 * math.pi)


def _expand_by_path(mean, log_std, attentions, mel_masks):
    """`mean @ attentions`, `log_Std @ attentions` (Modules.py:120-121).  The alignment is a 0/1 matrix with
    exactly one 1 per valid mel frame, so the two dense [80,T_x]x[T_x,T_y] products (99 % zeros, SURVEY 8f
    row 1) are a gather of token columns by frame -- and their backward a scatter-add -- with bit-identical
    values: the product's only non-zero term is the gathered element."""
    idx = attentions.argmax(dim=1)                                            # [B, T_y] token of each frame
    both = torch.cat([mean, log_std], dim=1)                                   # [B, 160, T_x]
    out = torch.gather(both, 2, idx.unsqueeze(1).expand(-1, both.shape[1], -1)) * mel_masks
    return out[:, :mean.shape[1]], out[:, mean.shape[1]:]


# Training runs the text encoder on its own stream next to the decoder (flow.enc_stream); GLOW_ENC_OVERLAP=0
# puts it back in front of the decoder on the caller's stream.
ENCODER_OVERLAP = os.environ.get("GLOW_ENC_OVERLAP", "1") != "0"
# log_P / path expansion / MLE loss on the fused kernels of csrc/align.cu; GLOW_FUSED_ALIGN=0 keeps the framework ops.
FUSED_ALIGN = os.environ.get("GLOW_FUSED_ALIGN", "1") != "0"


class GlowTTS(torch.nn.Module):
    """Modules.py:16-229 for Mode Vanilla and SE (LUT).  forward() returns the reference's
    8-tuple, inference() its 3-tuple."""

    def __init__(self):
        super().__init__()
        h = _hp()
        mode = h.Mode.upper()
        if mode not in ("VANILLA", "SE"):
            raise _lib.GlowCoreError("Mode %r is outside the accelerated path (Vanilla, SE-LUT)" % h.Mode)
        self.layer_Dict = torch.nn.ModuleDict()
        if mode == "SE":
            if h.Speaker_Embedding.Type.upper() != "LUT":
                raise _lib.GlowCoreError("only Speaker_Embedding.Type 'LUT' is supported (GE2E is out of scope)")
            self.layer_Dict["LUT"] = torch.nn.Embedding(h.Speaker_Embedding.Num_Speakers,
                                                        h.Speaker_Embedding.Embedding_Size)
            torch.nn.init.uniform_(self.layer_Dict["LUT"].weight, -1.0, 1.0)
        self.layer_Dict["Encoder"] = Encoder()
        self.layer_Dict["Decoder"] = Decoder()
        self.layer_Dict["Maximum_Path_Generater"] = Maximum_Path_Generater()
        self.num_squeeze = h.Decoder.Num_Squeeze
        self.max_abs_mel = h.Sound.Max_Abs_Mel
        self.flat = None

    # ---- flat storage for the whole model (one all-reduce, one fused optimizer step) ----
    def flatten_parameters(self):
        self.flat = FlatBuffer(list(self.parameters()))
        self.layer_Dict["Decoder"].adopt_flat(self.flat)
        return self.flat

    def Mask_Generate(self, lengths, max_lengths=None, dtype=torch.float):
        n = int(max_lengths) if max_lengths is not None else int(torch.max(lengths))
        return (torch.arange(n, device=lengths.device)[None, :] < lengths[:, None]).unsqueeze(1).to(dtype)

    @staticmethod
    def _masks_from_host(lens, device):
        n = max(lens)
        t = _lib.device_ints(lens, torch.int64, device)
        return (torch.arange(n, device=device)[None, :] < t[:, None]).unsqueeze(1).float()

    def forward(self, tokens, token_lengths, mels, mel_lengths, speakers=None, mels_for_ge2e=None, pitches=None,
                host_token_lengths=None, host_mel_lengths=None, geometry=None):
        """geometry: a geometry.StepGeometry already `update`d for this batch -- every length-derived tensor then
        comes from its fixed device buffers (tokens must be [B, geometry.t_text], mels [B, 80, geometry.t_mel]), so
        the call is the same stream of launches for every batch of the bucket (train.GraphedTrainStep)."""
        _lib.require_cuda(mels, "mels")
        d = self.layer_Dict
        dev = mels.device
        geo = geometry
        if geo is not None:
            geo.derive()
            tl, ml = geo.host_tl, geo.host_ml
            if tokens.shape[1] != geo.t_text or mels.shape[2] != geo.t_mel:
                raise ValueError("geometry expects tokens [B,%d] and mels [B,80,%d]" % (geo.t_text, geo.t_mel))
        else:
            tl = host_token_lengths if host_token_lengths is not None else _host_lengths(lengths=token_lengths)
            ml = host_mel_lengths if host_mel_lengths is not None else _host_lengths(lengths=mel_lengths)
        assert all(n % self.num_squeeze == 0 for n in ml), "Mel lengths must be diviable by Num_Squeeze."
        spk = d["LUT"](speakers) if "LUT" in d else None
        if geo is not None:
            token_masks, mel_masks, t_len, m_len = geo.token_masks, geo.mel_masks, geo.tl32, geo.ml32
        else:
            token_masks = self._masks_from_host(tl, dev)[:, :, :tokens.shape[1]]
            mel_masks = self._masks_from_host(ml, dev)
            t_len = _lib.device_ints(tl, torch.int32, dev)
            m_len = _lib.device_ints(ml, torch.int32, dev)

        dec = d["Decoder"]
        if torch.is_grad_enabled():
            dec.begin_prepare(dev)                 # weight_norm / slab images while the encoder runs
        overlap = torch.is_grad_enabled() and ENCODER_OVERLAP
        cur = torch.cuda.current_stream(dev)
        enc = _flow.enc_stream(dev) if overlap else cur
        if overlap:
            enc.wait_stream(cur)
        with torch.cuda.stream(enc):
            mean, log_std, log_dur, token_masks = d["Encoder"](tokens[:, :token_masks.shape[2]], token_masks, spk,
                                                               None, lengths=t_len, host_lengths=tl,
                                                               token_rows=geo.tok if geo is not None else None)
        dec.host_lengths, dec.geometry = ml, geo
        try:
            z, log_dets, mel_masks = dec(mels if geo is not None else mels[:, :, :max(ml)], mel_masks, spk, None, None)
        finally:
            dec.host_lengths = dec.geometry = None
        if overlap:
            cur.wait_stream(enc)
            for t in (mean, log_std, log_dur, token_masks):
                t.record_stream(cur)           # allocated on the encoder's stream, consumed from here on

        if FUSED_ALIGN:
            # log_P, the search and the expansion by its path on csrc/align.cu + mas.cu.
            # The fp32 parity mode keeps the reference's own expression for log_P (a differently ordered sum
            # could flip a near-tie of the search); the rest is exact either way.
            with torch.no_grad():
                if dec.precision.startswith("fp32"):
                    log_p = self._log_p(z, mean, log_std)
                else:
                    log_p = _align.log_p(z, mean, log_std, t_len, m_len)
                attentions, frame_token, durations = maximum_path_align(log_p, t_len, m_len)
            mel_mean, mel_log_std, log_dur_targets = _align.expand_by_path(mean, log_std, frame_token, durations,
                                                                           t_len, m_len, z.shape[2])
            return z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_targets, attentions, None

        with torch.no_grad():
            attentions = d["Maximum_Path_Generater"](self._log_p(z, mean, log_std), None, t_len, m_len)
        mel_mean, mel_log_std = _expand_by_path(mean, log_std, attentions, mel_masks)
        log_dur_targets = torch.log(attentions.sum(dim=-1).unsqueeze(1) + 1e-7) * token_masks
        return z, mel_mean, mel_log_std, log_dets, log_dur, log_dur_targets, attentions, None

    @staticmethod
    def _log_p(z, mean, log_std):
        """Modules.py:107-116 as the reference writes it (fp32)."""
        r = torch.exp(-2 * log_std)
        return ((-0.5 * math.log(2 * math.pi) - log_std).sum(dim=1).unsqueeze(-1)
                + r.transpose(2, 1) @ (-0.5 * z ** 2)
                + (mean * r).transpose(2, 1) @ z
                + (-0.5 * mean ** 2 * r).sum(dim=1).unsqueeze(-1))

    @torch.no_grad()
    def inference_device(self, tokens, token_lengths, speakers=None, noise_scale=1.0, length_scale=1.0,
                         max_mel_length=1000, noises=None):
        """`inference` without a single host round trip: the reference reads the predicted
        mel lengths back to size its tensors (Modules.py:174-176); here every tensor has the static length
        `max_mel_length` (longer predictions are cut there), the lengths stay on the device and the decoder runs
        on a fixed-geometry row map, so the whole call is one stream of launches -- capturable in a CUDA graph
        (infer.GraphedInference).  Same arithmetic as `inference`; returns (mels [B,80,max_mel_length],
        mel_Lengths [B] int64 (device), attentions [B,T_x,max_mel_length])."""
        d = self.layer_Dict
        dev = tokens.device
        _lib.require_cuda(tokens, "tokens")
        t_mel = int(max_mel_length) // self.num_squeeze * self.num_squeeze
        spk = d["LUT"](speakers) if "LUT" in d else None
        t_len = token_lengths.to(device=dev, dtype=torch.int32)
        token_masks = (torch.arange(tokens.shape[1], device=dev)[None, :] < t_len[:, None]).unsqueeze(1).float()
        mean, log_std, log_dur, mask = d["Encoder"](tokens, token_masks, spk, None, lengths=t_len)
        if not torch.is_tensor(length_scale):
            length_scale = torch.tensor([float(length_scale)], device=dev)
        length_scale = length_scale.to(dev).unsqueeze(-1).unsqueeze(-1)
        durations = torch.ceil(torch.exp(log_dur) * mask * length_scale).squeeze(1)          # :173
        mel_lengths = torch.clamp(durations.sum(dim=1), 1.0, float(t_mel)).long()             # :174, cut at the static size
        mel_masks = (torch.arange(t_mel, device=dev)[None, :] < mel_lengths[:, None]).unsqueeze(1).float()
        attention_masks = (token_masks.unsqueeze(-1) * mel_masks.unsqueeze(2)).squeeze(1)
        attentions = self.Path_Generate(durations, attention_masks)
        mel_mean = mean @ attentions
        mel_log_std = log_std @ attentions
        if noises is None:
            noises = torch.randn_like(mel_mean)
        z = (mel_mean + torch.exp(mel_log_std) * (noises[:, :, :t_mel] * noise_scale)) * mel_masks
        dec = d["Decoder"]
        dec.device_lengths = mel_lengths
        try:
            mels, _, mel_masks = dec(z, mel_masks, spk, None, None, reverse=True)
        finally:
            dec.device_lengths = None
        mels = mels.masked_fill(mel_masks == 0.0, -self.max_abs_mel)
        return mels, mel_lengths, attentions

    @torch.no_grad()
    def inference(self, tokens, token_lengths, mels_for_prosody=None, mel_lengths_for_prosody=None, speakers=None,
                  mels_for_ge2e=None, pitches=None, pitch_lengths=None, noise_scale=1.0, length_scale=1.0,
                  noises=None):
        d = self.layer_Dict
        dev = tokens.device
        _lib.require_cuda(tokens, "tokens")
        spk = d["LUT"](speakers) if "LUT" in d else None
        tl = _host_lengths(lengths=token_lengths)
        token_masks = self._masks_from_host(tl, dev)
        t_len = _lib.device_ints(tl, torch.int32, dev)
        mean, log_std, log_dur, mask = d["Encoder"](tokens[:, :token_masks.shape[2]], token_masks, spk, None,
                                                    lengths=t_len)
        if not torch.is_tensor(length_scale):
            length_scale = torch.tensor([float(length_scale)], device=dev)
        length_scale = length_scale.to(dev).unsqueeze(-1).unsqueeze(-1)
        durations = torch.ceil(torch.exp(log_dur) * mask * length_scale).squeeze(1)          # :173
        mel_lengths = torch.clamp_min(durations.sum(dim=1), 1.0).long()                       # :174
        mel_masks = self.Mask_Generate(mel_lengths)
        attention_masks = (token_masks.unsqueeze(-1) * mel_masks.unsqueeze(2)).squeeze(1)
        attentions = self.Path_Generate(durations, attention_masks)
        mel_mean = mean @ attentions
        mel_log_std = log_std @ attentions
        if noises is None:
            noises = torch.randn_like(mel_mean)
        noises = noises[:, :, :mel_mean.shape[2]] * noise_scale
        z = (mel_mean + torch.exp(mel_log_std) * noises) * mel_masks                          # :191
        mels, _, mel_masks = d["Decoder"](z, mel_masks, spk, None, None, reverse=True)
        mels = mels.masked_fill(mel_masks == 0.0, -self.max_abs_mel)                           # :202
        return mels, mel_lengths, attentions

    def Path_Generate(self, durations, masks):
        """Modules.py:213-229: cumulative durations -> 0/1 alignment."""
        b, tx, ty = masks.shape
        ends = torch.cumsum(durations, dim=1)
        starts = ends - durations
        frame = torch.arange(ty, device=masks.device, dtype=durations.dtype)[None, None, :]
        paths = ((frame < ends.unsqueeze(-1)) & (frame >= starts.unsqueeze(-1))).to(masks.dtype)
        return paths * masks
