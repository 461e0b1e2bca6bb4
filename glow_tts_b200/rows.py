"""Packed token rows for the text encoder.

The reference runs the encoder on [B, C, T_x] tensors padded to the longest sentence and
re-multiplies by the mask before every conv (Modules.py:554,567,570).  Here the tokens of a batch
are packed along ONE row axis, channels-last -- the decoder's layout (csrc/flow_layout.cuh): two
guard rows between sentences are the convs' zero padding, LayerNorm normalises the contiguous last
dim (no transposes), and every conv is one tcgen05 GEMM launch (csrc/rows_conv.cu) that applies the
mask as it loads and as it stores.
"""
import ctypes

import torch

from . import _lib, flow as _flow


class TokenRows:
    """Row map of one batch of sentence lengths + the gather indices between [B, T] and rows."""

    def __init__(self, lengths, t_max, device):
        self.rm = _flow.row_map(lengths, device)
        rm = self.rm
        b = len(lengths)
        self.batch, self.t_max, self.rows_pad = b, int(t_max), rm.rows_pad
        row_utt = rm.row_utt.long()
        valid = row_utt >= 0
        self.valid = valid.to(torch.float32).unsqueeze(1)                         # [R, 1]
        # rows -> flat (b, t) source index (guards read element 0 and are zeroed by `valid`)
        self.src_idx = torch.where(valid, row_utt * self.t_max + rm.row_t.long(), torch.zeros_like(row_utt))
        # flat (b, t) -> row (positions beyond a sentence read row 0, a guard row, and are zeroed by `tmask`)
        t = torch.arange(self.t_max, device=device)[None, :]
        lens = rm.utt_len.long()[:, None]
        self.tmask = (t < lens).to(torch.float32)                                 # [B, T]
        self.dst_idx = torch.where(t < lens, rm.utt_off.long()[:, None] + t, torch.zeros_like(t)).reshape(-1)

    def pack(self, x_btc):
        """[B, T, C] -> [rows_pad, C] (guard rows zero)."""
        b, t, c = x_btc.shape
        return x_btc.reshape(b * t, c).index_select(0, self.src_idx) * self.valid

    def unpack(self, rows):
        """[rows_pad, C] -> [B, T, C] (positions beyond each sentence zero)."""
        c = rows.shape[1]
        return rows.index_select(0, self.dst_idx).view(self.batch, self.t_max, c) * self.tmask.unsqueeze(2)


_CACHE = {}
ACCUMULATE = False       # set by train.TrainStep: weight gradients go straight into the attached .grad storage
                         # on the side stream; whoever enables it must call join() before reading gradients
PENDING = []             # tensors the side stream still reads; cleared by join()


def join(device):
    """Wait (stream-ordered) for the side-stream weight gradients; call before reading any .grad."""
    _lib.side_join(device)
    PENDING.clear()


def token_rows(lengths, t_max, device):
    key = (tuple(int(n) for n in lengths), int(t_max), str(device))
    tr = _CACHE.get(key)
    if tr is None:
        if len(_CACHE) > 64:
            _CACHE.clear()
        tr = _CACHE[key] = TokenRows(key[0], t_max, device)
    return _lib.keepalive(tr)


def _call(tr, cin, cout, taps, device, relu=False, p=0.0, seed=0):
    c = _lib.RowsConvCall()
    c.cin, c.cout, c.taps, c.rows_pad = cin, cout, taps, tr.rows_pad
    c.row_utt = tr.rm.row_utt.data_ptr()
    c.relu, c.p_out, c.seed_out = int(bool(relu)), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF
    c.step_dev = _lib.step_counter_ptr(device) if seed else None
    c.stream = torch.cuda.current_stream(device).cuda_stream
    return c


class PackedWeights:
    """bf16 slab images of a set of Conv1d weights, refreshed by ONE kernel per step (pack())."""

    def __init__(self, convs, device):
        self.convs = list(convs)
        L = _lib.lib()
        sizes = []
        for c in self.convs:
            cout, cin, taps = c.weight.shape
            n = L.glow_rows_conv_slab_elems(cin, cout, taps)
            if n == 0:
                raise _lib.GlowCoreError("rows_conv: no kernel built for cin=%d cout=%d taps=%d" % (cin, cout, taps))
            sizes.append(n)
        self.buf = torch.empty(2 * sum(sizes) + 64 * len(sizes), dtype=torch.bfloat16, device=device)
        self.slabs, pos = {}, 0
        for c, n in zip(self.convs, sizes):
            w = self.buf[pos:pos + n]
            pos += (n + 63) // 64 * 64
            wt = self.buf[pos:pos + n]
            pos += (n + 63) // 64 * 64
            self.slabs[id(c)] = (w, wt)
        k = len(self.convs)
        self._shapes = (ctypes.c_int * (3 * k))(*[v for c in self.convs for v in (c.weight.shape[1], c.weight.shape[0],
                                                                                  c.weight.shape[2])])
        self._sw = (ctypes.c_void_p * k)(*[self.slabs[id(c)][0].data_ptr() for c in self.convs])
        self._swt = (ctypes.c_void_p * k)(*[self.slabs[id(c)][1].data_ptr() for c in self.convs])

    def pack(self):
        dev = self.buf.device
        k = len(self.convs)
        w = (ctypes.c_void_p * k)(*[c.weight.data_ptr() for c in self.convs])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().glow_rows_conv_pack_multi(k, self._shapes, w, self._sw, self._swt, _lib.stream_ptr(dev)),
                       "glow_rows_conv_pack_multi")

    def get(self, conv):
        return self.slabs.get(id(conv))


class RowsConvFn(torch.autograd.Function):
    """y = mask * Dropout(ReLU?(bias + conv1d(mask * x))) on packed rows; weight in torch's Conv1d layout
    [cout, cin, taps].  seed == 0 disables the dropout (eval)."""

    @staticmethod
    def forward(ctx, x, weight, bias, tr, relu, p, seed, slabs=None):
        cout, cin, taps = weight.shape
        dev = x.device
        x = x.contiguous().float()
        L = _lib.lib()
        y = torch.empty((tr.rows_pad, cout), dtype=torch.float32, device=dev)
        if not (p > 0 and seed):
            p, seed = 0.0, 0
        with torch.cuda.device(dev):
            call = _call(tr, cin, cout, taps, dev, relu, p, seed)
            if slabs is not None:                                  # packed once for the whole step (PackedWeights)
                slab_w, slab_wt = slabs
            else:
                n = L.glow_rows_conv_slab_elems(cin, cout, taps)
                if n == 0:
                    raise _lib.GlowCoreError("rows_conv: no kernel built for cin=%d cout=%d taps=%d" % (cin, cout, taps))
                slab_w = torch.empty(n, dtype=torch.bfloat16, device=dev)
                slab_wt = torch.empty(n, dtype=torch.bfloat16, device=dev)
                _lib.check(L.glow_rows_conv_pack(ctypes.byref(call), _lib.ptr(weight.detach().contiguous()),
                                                 _lib.ptr(slab_w), _lib.ptr(slab_wt)), "glow_rows_conv_pack")
            _lib.check(L.glow_rows_conv_forward(ctypes.byref(call), _lib.ptr(x), _lib.ptr(slab_w),
                                                _lib.ptr(bias.detach().contiguous()) if bias is not None else None,
                                                _lib.ptr(y)), "glow_rows_conv_forward")
        ctx.tr, ctx.shape, ctx.has_bias, ctx.act = tr, (cin, cout, taps), bias is not None, (bool(relu), p, seed)
        # accumulate-in-place mode: both gradients already live in (contiguous) flat storage
        wg = weight.grad if (ACCUMULATE and weight.grad is not None and weight.grad.is_contiguous()) else None
        bg = bias.grad if (wg is not None and bias is not None and bias.grad is not None) else None
        if wg is not None and bias is not None and bg is None:
            wg = None
        ctx.grad_views = (wg, bg)
        ctx.save_for_backward(x, slab_wt, y if (relu or seed) else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, slab_wt, y = ctx.saved_tensors
        tr = ctx.tr
        cin, cout, taps = ctx.shape
        relu, p, seed = ctx.act
        dev = dy.device
        dy = dy.contiguous().float()
        L = _lib.lib()
        dx = dw = db = None
        with torch.cuda.device(dev):
            call = _call(tr, cin, cout, taps, dev)
            st = _lib.stream_ptr(dev)
            # gradient w.r.t. the conv result: through the fused ReLU / dropout, zero on guard rows
            g = torch.empty_like(dy)
            _lib.check(L.glow_rows_act_backward(tr.rm.row_utt.data_ptr(), tr.rows_pad, cout, int(relu), float(p),
                                                int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                _lib.step_counter_ptr(dev) if seed else None,
                                                _lib.ptr(dy), _lib.ptr(y), _lib.ptr(g), st), "glow_rows_act_backward")
            if ctx.needs_input_grad[0]:
                dx = torch.empty((tr.rows_pad, cin), dtype=torch.float32, device=dev)
                _lib.check(L.glow_rows_conv_backward_data(ctypes.byref(call), _lib.ptr(g), _lib.ptr(slab_wt),
                                                          _lib.ptr(dx)), "glow_rows_conv_backward_data")
            wg, bg = ctx.grad_views
            if ctx.needs_input_grad[1] and wg is not None:
                # straight into the parameters' gradient storage, on the side stream (joined by the caller
                # through _lib.side_join before anything reads the gradients)
                xm = x if ctx.x_masked else x * tr.valid
                scratch = torch.empty((taps, cin, cout), dtype=torch.float32, device=dev)
                _lib.check(L.glow_rows_conv_backward_weight_accum(ctypes.byref(call), _lib.ptr(xm), _lib.ptr(g),
                                                                  _lib.ptr(wg), _lib.ptr(bg), _lib.ptr(scratch)),
                           "glow_rows_conv_backward_weight_accum")
                PENDING.append((xm, g, scratch))                  # alive until side_join
            elif ctx.needs_input_grad[1]:
                xm = x if ctx.x_masked else x * tr.valid          # the reduction runs over real rows only
                dwt = torch.empty((taps, cin, cout), dtype=torch.float32, device=dev)
                db = torch.empty(cout, dtype=torch.float32, device=dev) if ctx.has_bias else None
                _lib.check(L.glow_rows_conv_backward_weight(ctypes.byref(call), _lib.ptr(xm), _lib.ptr(g),
                                                            _lib.ptr(dwt), _lib.ptr(db)),
                           "glow_rows_conv_backward_weight")
                dw = dwt.permute(2, 1, 0)
        return dx, dw, db, None, None, None, None


def rows_conv(x, conv, tr, relu=False, p=0.0, seed=0, x_masked=False, packed=None):
    """Apply a torch.nn.Conv1d's parameters to packed rows.  x_masked: the caller guarantees x is zero on
    guard rows (every op of this module writes them as zeros), which saves a masking pass in backward.
    packed: a PackedWeights whose pack() already ran this step (else the weight is packed here)."""
    slabs = packed.get(conv) if packed is not None else None
    return _RowsConvApply.apply(x, conv.weight, conv.bias, tr, relu, p, seed, x_masked, slabs)


class _RowsConvApply(RowsConvFn):
    @staticmethod
    def forward(ctx, x, weight, bias, tr, relu, p, seed, x_masked, slabs):
        ctx.x_masked = bool(x_masked)
        return RowsConvFn.forward(ctx, x, weight, bias, tr, relu, p, seed, slabs)

    @staticmethod
    def backward(ctx, dy):
        return RowsConvFn.backward(ctx, dy) + (None, None)


class RowsNormFn(torch.autograd.Function):
    """y = mask * Dropout_out(ReLU?(LayerNorm(Dropout_in(a) + b))) on packed rows (csrc/rows_norm.cu)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, tr, eps, p_in, seed_in, relu, p_out, seed_out):
        dev = a.device
        a = a.contiguous().float()
        b = b.contiguous().float() if b is not None else None
        if not (p_in > 0 and seed_in):
            p_in, seed_in = 0.0, 0
        if not (p_out > 0 and seed_out):
            p_out, seed_out = 0.0, 0
        c = _lib.RowsNormCall()
        c.rows_pad, c.channels, c.row_utt, c.eps = tr.rows_pad, a.shape[1], tr.rm.row_utt.data_ptr(), float(eps)
        c.p_in, c.seed_in = float(p_in), int(seed_in) & 0xFFFFFFFFFFFFFFFF
        c.relu, c.p_out, c.seed_out = int(bool(relu)), float(p_out), int(seed_out) & 0xFFFFFFFFFFFFFFFF
        c.step_dev = _lib.step_counter_ptr(dev) if (seed_in or seed_out) else None
        s = torch.empty_like(a)
        y = torch.empty_like(a)
        stats = torch.empty((tr.rows_pad, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            c.stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.lib().glow_rows_norm_forward(ctypes.byref(c), _lib.ptr(a), _lib.ptr(b),
                                                         _lib.ptr(gamma.detach().contiguous()),
                                                         _lib.ptr(beta.detach().contiguous()),
                                                         _lib.ptr(s), _lib.ptr(stats), _lib.ptr(y)),
                       "glow_rows_norm_forward")
        ctx.call, ctx.has_b = c, b is not None
        ctx.save_for_backward(s, stats, y, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        s, stats, y, gamma = ctx.saved_tensors
        c, dev = ctx.call, dy.device
        dy = dy.contiguous().float()
        da = torch.empty_like(dy)
        db = torch.empty_like(dy) if (ctx.has_b and c.seed_in) else None     # without input dropout d b == d a
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        with torch.cuda.device(dev):
            c.stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.lib().glow_rows_norm_backward(ctypes.byref(c), _lib.ptr(dy), _lib.ptr(y), _lib.ptr(s),
                                                          _lib.ptr(stats), _lib.ptr(gamma.detach().contiguous()),
                                                          _lib.ptr(da), _lib.ptr(db), _lib.ptr(dgamma),
                                                          _lib.ptr(dbeta)), "glow_rows_norm_backward")
        if ctx.has_b and db is None:
            db = da
        return da, (db if ctx.has_b else None), dgamma, dbeta, None, None, None, None, None, None, None


def rows_norm(a, b, ln, tr, p_in=0.0, seed_in=0, relu=False, p_out=0.0, seed_out=0):
    """Apply a torch.nn.LayerNorm's parameters: mask * Dropout_out(ReLU?(LN(Dropout_in(a) + b)))."""
    return RowsNormFn.apply(a, b, ln.weight, ln.bias, tr, ln.eps, p_in, seed_in, relu, p_out, seed_out)
