"""Host side of the flow decoder: packed-row maps, workspaces, the per-step weight
preparation and the autograd bridge to libglowcore's glow_flow_* entry points
(include/glowcore.h).  Mirrors what Modules.py:286-309 (Decoder) drives in the
reference; no arithmetic happens in Python."""
import ctypes
import os

import numpy as np
import torch

from . import _lib

GUARD = 2
ROW_TILE = 128


class RowMap:
    """Packed row axis for a batch of squeezed lengths (host-built, device-resident)."""

    def __init__(self, sq_lengths, device):
        lens = np.asarray(sq_lengths, dtype=np.int64)
        if lens.ndim != 1 or len(lens) == 0 or (lens < 0).any():
            raise ValueError("squeezed lengths must be a non-empty 1-D array of non-negative ints")
        b = len(lens)
        off = np.zeros(b, np.int64)
        pos = GUARD
        for i in range(b):
            off[i] = pos
            pos += int(lens[i]) + GUARD
        rows_pad = max(ROW_TILE, (pos + ROW_TILE - 1) // ROW_TILE * ROW_TILE)
        row_utt = np.full(rows_pad, -1, np.int32)
        row_t = np.zeros(rows_pad, np.int32)
        for i in range(b):
            row_utt[off[i]:off[i] + lens[i]] = i
            row_t[off[i]:off[i] + lens[i]] = np.arange(lens[i], dtype=np.int32)
        self.batch, self.rows_pad, self.rows_real = b, int(rows_pad), int(lens.sum())
        self.lengths = lens
        packed = np.concatenate([row_utt, row_t, off.astype(np.int32), lens.astype(np.int32)])
        dev = torch.from_numpy(packed).to(device, non_blocking=True)
        self.row_utt = dev[:rows_pad]
        self.row_t = dev[rows_pad:2 * rows_pad]
        self.utt_off = dev[2 * rows_pad:2 * rows_pad + b]
        self.utt_len = dev[2 * rows_pad + b:]
        self._keep = dev


class DeviceRowMap:
    """The same packed row axis with a FIXED geometry -- every utterance owns `sq_max` rows plus guards -- and the
    validity of each row decided on the device from device-resident lengths (`update`).  No host copy of the
    lengths is needed, so a decoder call that uses it is sync-free and CUDA-graph capturable (GlowTTS.inference
    predicts the mel lengths on the device, Modules.py:173-174).  Padding rows are computed and masked."""

    def __init__(self, batch, sq_max, device):
        self.batch, self.sq_max = int(batch), int(sq_max)
        self.stride = self.sq_max + GUARD
        pos = GUARD + self.batch * self.stride
        self.rows_pad = max(ROW_TILE, (pos + ROW_TILE - 1) // ROW_TILE * ROW_TILE)
        self.rows_real = self.batch * self.sq_max
        dev = torch.device(device)
        r = torch.arange(self.rows_pad, device=dev, dtype=torch.int64) - GUARD
        self._b = torch.div(r, self.stride, rounding_mode="floor")
        self._t = r - self._b * self.stride
        self._in = (r >= 0) & (self._b < self.batch) & (self._t < self.sq_max)
        self._bc = self._b.clamp(0, self.batch - 1)
        self.row_utt = torch.full((self.rows_pad,), -1, dtype=torch.int32, device=dev)
        self.row_t = torch.zeros(self.rows_pad, dtype=torch.int32, device=dev)
        self.utt_off = (GUARD + torch.arange(self.batch, device=dev) * self.stride).to(torch.int32)
        self.utt_len = torch.zeros(self.batch, dtype=torch.int32, device=dev)

    def update(self, sq_lengths):
        """sq_lengths: device integer tensor [batch] (clamped to sq_max); stream-ordered, no sync."""
        n = sq_lengths.to(torch.int64).clamp(0, self.sq_max)
        valid = self._in & (self._t < n[self._bc])
        self.row_utt.copy_(torch.where(valid, self._b, torch.full_like(self._b, -1)))
        self.row_t.copy_(torch.where(valid, self._t, torch.zeros_like(self._t)))
        self.utt_len.copy_(n)
        return self


_ROWMAP_CACHE = {}


def row_map(sq_lengths, device):
    key = (tuple(int(x) for x in sq_lengths), str(device))
    rm = _ROWMAP_CACHE.get(key)
    if rm is None:
        if len(_ROWMAP_CACHE) > 64:
            _ROWMAP_CACHE.clear()
        rm = _ROWMAP_CACHE[key] = RowMap(sq_lengths, device)
    return _lib.keepalive(rm)


# ---- side stream for work that is off the step's critical path ------------------------------
# The per-step weight preparation (weight_norm -> packed / slab images) only depends on the parameters,
# so it can run while the text encoder does; the way back from effective-weight gradients to parameter
# gradients only feeds the optimizer, so it can run while the encoder's backward does.  Both go to one
# torch side stream per device; join() makes the current stream wait for it.  (Stream fork / join is
# capturable, so the overlap survives in the CUDA graph.)
_SIDE = {}
_SIDE_BUSY = set()
# GLOW_FUSED_PARAM_GRADS=1: glow_flow_backward_params -- every block's parameter gradients inside the backward call,
# right behind that block's weight gradients.  Off by default: measured 0.1 ms/step SLOWER at B = 32 (the extra
# memory traffic lands in the middle of the data-gradient chain instead of in the idle tail); default is one
# glow_flow_param_grads pass after the backward, on the torch side stream.
FUSED_PARAM_GRADS = os.environ.get("GLOW_FUSED_PARAM_GRADS", "0") == "1"


def side_stream(device):
    key = str(torch.device(device))
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


_ENC = {}


def enc_stream(device):
    """The text encoder's own stream: nothing in the encoder depends on the decoder (and vice versa) until
    log_P, so training runs the two side by side, forward and -- because autograd replays every node on the
    stream its forward ran on -- backward.  Fork / join with wait_stream, so it is capturable."""
    key = str(torch.device(device))
    if key not in _ENC:
        _ENC[key] = torch.cuda.Stream(device)
    return _ENC[key]


def join(device):
    key = str(torch.device(device))
    if key in _SIDE_BUSY:
        torch.cuda.current_stream(device).wait_stream(_SIDE[key])
        _SIDE_BUSY.discard(key)


def precision_tag(precision):
    if precision in ("fp32", "f32", torch.float32):
        return _lib.GLOW_F32, torch.float32
    if precision in ("bf16", torch.bfloat16):
        return _lib.GLOW_BF16, torch.bfloat16
    if precision == "bf16-simt":          # bf16 storage on the CUDA-core GEMM: device cross-check of the tcgen05 path
        return _lib.GLOW_BF16_SIMT, torch.bfloat16
    if precision == "fp32-tc":            # fp32 storage, tcgen05 with every operand split into bf16 hi + lo parts
        return _lib.GLOW_F32_TC, torch.float32
    raise ValueError("precision must be 'fp32', 'fp32-tc', 'bf16' or 'bf16-simt', got %r" % (precision,))


class FlowPlan:
    """Everything one decoder call needs besides the tensors: config, flat parameter
    buffer + offset table, packed weights, workspaces (cached per shape)."""

    def __init__(self, blocks, spk_dim, dropout, channels=160, hidden=192, layers=4, kernel=5, split=4):
        self.cfg = _lib.FlowConfig(blocks, channels, hidden, layers, kernel, split, spk_dim, dropout)
        L = _lib.lib()
        slots = L.glow_flow_param_slots(ctypes.byref(self.cfg))
        if slots <= 0:
            _lib.check(slots if slots < 0 else -2, "glow_flow_param_slots")
        self.slots = slots
        self.wpack_floats = L.glow_flow_wpack_floats(ctypes.byref(self.cfg))
        self.wpack_tc_elems = L.glow_flow_wpack_tc_elems(ctypes.byref(self.cfg))
        self.wpack_tc_elems_for = lambda tag: L.glow_flow_wpack_tc_elems_for(ctypes.byref(self.cfg), tag)
        self._ws = {}
        self._wpack = {}

    # -- packed effective weights -------------------------------------------------
    def wpack(self, device, tag):
        key = (str(device), tag)
        if key not in self._wpack:
            wp = torch.empty(self.wpack_floats, dtype=torch.float32, device=device)
            wtc = (torch.empty(self.wpack_tc_elems_for(tag), dtype=torch.bfloat16, device=device)
                   if tag != _lib.GLOW_F32 else None)
            dwp = torch.empty(self.wpack_floats, dtype=torch.float32, device=device)
            self._wpack[key] = (wp, wtc, dwp)
        return _lib.keepalive(self._wpack[key])

    def prepare(self, flat_params, offsets_host, device, tag):
        wp, wtc, _ = self.wpack(device, tag)
        rc = _lib.lib().glow_flow_prepare(ctypes.byref(self.cfg), _lib.ptr(flat_params),
                                          offsets_host.ctypes.data, tag, _lib.ptr(wp), _lib.ptr(wtc),
                                          _lib.stream_ptr(device))
        _lib.check(rc, "glow_flow_prepare")
        return wp, wtc

    # -- workspaces ------------------------------------------------------------------
    def workspace(self, rm, device, act_dtype, training):
        key = (rm.rows_pad, rm.batch, str(device), act_dtype, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) > 8:
                self._ws.clear()
            out = (ctypes.c_size_t * 4)()
            rc = _lib.lib().glow_flow_workspace_elems(ctypes.byref(self.cfg), rm.rows_pad, rm.batch,
                                                      int(training), out)
            _lib.check(rc, "glow_flow_workspace_elems")
            # zero-filled once: guard rows of every buffer must read as finite zeros
            ws = (torch.zeros(max(out[0], 1), dtype=torch.float32, device=device),
                  torch.zeros(max(out[1], 1), dtype=act_dtype, device=device),
                  torch.zeros(max(out[2], 1), dtype=torch.float32, device=device) if training else None,
                  torch.zeros(max(out[3], 1), dtype=act_dtype, device=device) if training else None)
            self._ws[key] = ws
        return _lib.keepalive(ws)

    # One workspace per shape holds the activations a backward needs, so only ONE grad-enabled forward of a
    # shape may be in flight: a second forward overwrites what the first one's backward would read.  Every
    # forward stamps the workspace; backward refuses to run on a workspace that has been re-stamped since.
    def stamp(self, ws):
        self._gen = getattr(self, "_gen", 0) + 1
        self._stamps = getattr(self, "_stamps", {})
        self._stamps[ws[0].data_ptr()] = self._gen
        return self._gen

    def stamp_of(self, ws):
        return getattr(self, "_stamps", {}).get(ws[0].data_ptr())

    def call_struct(self, rm, t_max, tag, wp, wtc, spk, ws, training, seed, device):
        c = _lib.FlowCall()
        c.cfg = self.cfg
        c.precision, c.batch, c.t_max, c.rows_pad = tag, rm.batch, int(t_max), rm.rows_pad
        c.training, c.seed = int(training), int(seed) & 0xFFFFFFFFFFFFFFFF
        c.step_dev = _lib.step_counter_ptr(device) if seed else None
        c.row_utt, c.row_t = rm.row_utt.data_ptr(), rm.row_t.data_ptr()
        c.utt_off, c.utt_len = rm.utt_off.data_ptr(), rm.utt_len.data_ptr()
        c.wpack = wp.data_ptr()
        c.wpack_tc = wtc.data_ptr() if wtc is not None else None
        c.spk = spk.data_ptr() if spk is not None else None
        c.ws_f32, c.ws_act = ws[0].data_ptr(), ws[1].data_ptr()
        c.bw_f32 = ws[2].data_ptr() if ws[2] is not None else None
        c.bw_act = ws[3].data_ptr() if ws[3] is not None else None
        c.stream = torch.cuda.current_stream(device).cuda_stream
        return c


class FlowDecoderFn(torch.autograd.Function):
    """mel [B,80,T] -> (z [B,80,T], logdet [B]).  `params` are passed only so autograd
    knows the dependency; the kernels read them through `owner`'s flat buffer."""

    @staticmethod
    def forward(ctx, owner, rm, mel, spk, seed, *params):
        plan, device = owner.plan, mel.device
        tag, act_dtype = precision_tag(owner.precision)
        training = any(ctx.needs_input_grad)
        flat, offs = owner.flat_params()
        mel = mel.contiguous().float()
        spk_c = spk.contiguous().float() if spk is not None else None
        b, _, t = mel.shape
        with torch.cuda.device(device):
            if owner._prepared is not None:                       # started early on the side stream (begin_prepare)
                wp, wtc = owner._prepared
                owner._prepared = None
                join(device)
            else:
                wp, wtc = plan.prepare(flat, offs, device, tag)
            ws = plan.workspace(rm, device, act_dtype, training)
            call = plan.call_struct(rm, t, tag, wp, wtc, spk_c, ws, training, seed, device)
            z = torch.empty_like(mel)
            logdet = torch.empty(b, dtype=torch.float32, device=device)
            rc = _lib.lib().glow_flow_forward(ctypes.byref(call), _lib.ptr(mel), _lib.ptr(z), _lib.ptr(logdet))
        _lib.check(rc, "glow_flow_forward")
        ctx.owner, ctx.rm, ctx.call, ctx.keep = owner, rm, call, (wp, wtc, ws, spk_c, flat)
        ctx.ws_stamp = plan.stamp(ws) if training else None
        ctx.need_dmel = mel.requires_grad
        ctx.has_spk = spk is not None
        ctx.training = training
        ctx.n_params = len(params)
        return z, logdet

    @staticmethod
    def backward(ctx, dz, dlogdet):
        if not ctx.training:
            raise _lib.GlowCoreError("flow decoder forward ran without grad: no saved activations")
        owner, rm, call = ctx.owner, ctx.rm, ctx.call
        plan = owner.plan
        wp, wtc, ws, spk_c, flat = ctx.keep
        if plan.stamp_of(ws) != ctx.ws_stamp:
            raise _lib.GlowCoreError(
                "flow decoder backward: the saved activations of this forward were overwritten by a later "
                "grad-enabled forward of the same shape (one workspace per shape: run backward before the next "
                "forward, e.g. accumulate gradients with one backward per micro-batch)")
        device = wp.device
        tag, _ = precision_tag(owner.precision)
        dz = dz.contiguous().float() if dz is not None else torch.zeros(
            (rm.batch, 80, call.t_max), dtype=torch.float32, device=device)
        dlogdet = (dlogdet.contiguous().float() if dlogdet is not None
                   else torch.zeros(rm.batch, dtype=torch.float32, device=device))
        _, _, dwp = plan.wpack(device, tag)
        dmel = torch.empty_like(dz) if ctx.need_dmel else None
        dspk = torch.empty_like(spk_c) if ctx.has_spk else None
        _, offs = owner.flat_params()
        gflat, direct = owner.flat_grads()
        with torch.cuda.device(device):
            call.stream = torch.cuda.current_stream(device).cuda_stream
            if direct and owner.defer_param_grads and FUSED_PARAM_GRADS:
                # gradients land in the attached flat buffer: every block's parameter gradients are produced on the
                # library's side stream right behind that block's weight gradients
                rc = _lib.lib().glow_flow_backward_params(ctypes.byref(call), _lib.ptr(dz), _lib.ptr(dlogdet),
                                                          _lib.ptr(dwp), _lib.ptr(dmel), _lib.ptr(dspk), _lib.ptr(flat),
                                                          offs.ctypes.data, _lib.ptr(gflat))
                _lib.check(rc, "glow_flow_backward_params")
                owner.fused_backward_calls = getattr(owner, "fused_backward_calls", 0) + 1
                return (None, None, dmel, dspk, None) + (None,) * ctx.n_params
            rc = _lib.lib().glow_flow_backward(ctypes.byref(call), _lib.ptr(dz), _lib.ptr(dlogdet), _lib.ptr(dwp),
                                               _lib.ptr(dmel), _lib.ptr(dspk))
            _lib.check(rc, "glow_flow_backward")
            if direct and owner.defer_param_grads:
                # gradients land in the attached flat buffer: nothing downstream in autograd needs them, so
                # the conversion runs on the side stream (the owner of the step joins before the optimizer)
                cur, side = torch.cuda.current_stream(device), side_stream(device)
                side.wait_stream(cur)
                dlogdet.record_stream(side)
                with torch.cuda.stream(side):
                    rc = _lib.lib().glow_flow_param_grads(ctypes.byref(plan.cfg), _lib.ptr(flat), offs.ctypes.data,
                                                          _lib.ptr(wp), _lib.ptr(dwp), _lib.ptr(dlogdet),
                                                          rm.utt_len.data_ptr(), rm.batch, _lib.ptr(gflat),
                                                          _lib.stream_ptr(device))
                _SIDE_BUSY.add(str(torch.device(device)))
            else:
                rc = _lib.lib().glow_flow_param_grads(ctypes.byref(plan.cfg), _lib.ptr(flat), offs.ctypes.data,
                                                      _lib.ptr(wp), _lib.ptr(dwp), _lib.ptr(dlogdet),
                                                      rm.utt_len.data_ptr(), rm.batch, _lib.ptr(gflat),
                                                      _lib.stream_ptr(device))
            _lib.check(rc, "glow_flow_param_grads")
        if direct:        # gradients were accumulated straight into the .grad views
            pgrads = (None,) * ctx.n_params
        else:
            pgrads = owner.split_grads(gflat)
        return (None, None, dmel, dspk, None) + tuple(pgrads)


def flow_reverse(owner, rm, z, spk, fill):
    """z [B,80,T] -> mel [B,80,T] (Decoder(reverse=True), Modules.py:303,664), no autograd."""
    plan, device = owner.plan, z.device
    tag, act_dtype = precision_tag(owner.precision)
    flat, offs = owner.flat_params()
    z = z.contiguous().float()
    spk_c = spk.contiguous().float() if spk is not None else None
    with torch.cuda.device(device):
        wp, wtc = plan.prepare(flat, offs, device, tag)
        ws = plan.workspace(rm, device, act_dtype, False)
        call = plan.call_struct(rm, z.shape[2], tag, wp, wtc, spk_c, ws, False, 0, device)
        mel = torch.empty_like(z)
        rc = _lib.lib().glow_flow_reverse(ctypes.byref(call), _lib.ptr(z), _lib.ptr(mel), ctypes.c_float(fill))
    _lib.check(rc, "glow_flow_reverse")
    return mel
