"""Drop-in for the reference's RPR_MHA.py.

``RPR_Multihead_Attention`` keeps the constructor (misspelled keyword names are
part of the API, RPR_MHA.py:6-19), parameter names (``layer_Dict.{Query,Key,
Value,Projection}``, ``weight_K``, ``weight_V``) and the forward contract
``(queries, keys=None, values=None, masks=None) -> (out [B,C_out,T],
alignments [B,H,T,T])``.  The 1x1 projections stay torch convs; everything
between them (RPR_MHA.py:95-165) runs in libglowcore's
glow_rpr_attention_forward / _backward (csrc/attention.cu).
"""
import ctypes

import torch

from . import _lib


class _AttnCoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, wk, wv, lengths, mask, heads, window, dropout, seed, need_align):
        b, c, t = q.shape
        d = c // heads
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        wk2, wv2 = wk.reshape(-1, d).contiguous().float(), wv.reshape(-1, d).contiguous().float()
        dev = q.device
        call = _lib.AttnCall()
        call.q, call.k, call.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
        call.wk, call.wv = wk2.data_ptr(), wv2.data_ptr()
        call.lengths = lengths.data_ptr() if lengths is not None else None
        mask_c = mask.contiguous().float() if (mask is not None and lengths is None) else None
        call.mask = mask_c.data_ptr() if mask_c is not None else None
        call.batch, call.heads, call.t, call.head_dim, call.window = b, heads, t, d, window
        call.dropout, call.seed = float(dropout), int(seed)
        call.step_dev = _lib.step_counter_ptr(dev) if seed else None
        needs_grad = any(ctx.needs_input_grad[:5])
        out = torch.empty_like(q)
        probs = torch.empty((b, heads, t, t), dtype=torch.float32, device=dev) if needs_grad else None
        align = None
        if need_align:
            align = probs if (probs is not None and seed == 0) else torch.empty(
                (b, heads, t, t), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call.stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _lib.lib().glow_rpr_attention_forward(
                ctypes.byref(call), _lib.ptr(out), _lib.ptr(probs),
                _lib.ptr(align) if align is not probs else None)
        _lib.check(rc, "glow_rpr_attention_forward")
        ctx.call, ctx.keep = call, (q, k, v, wk2, wv2, lengths, mask_c, probs)
        ctx.wshape = wk.shape
        if align is None:
            align = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(align)
        return out, align

    @staticmethod
    def backward(ctx, dout, _dalign):
        q, k, v, wk2, wv2, lengths, mask_c, probs = ctx.keep
        if probs is None:
            raise _lib.GlowCoreError("attention forward ran without grad: nothing saved for backward")
        call, dev = ctx.call, q.device
        dout = dout.contiguous().float()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dwk, dwv = torch.empty_like(wk2), torch.empty_like(wv2)
        ds = torch.empty_like(probs)
        with torch.cuda.device(dev):
            call.stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _lib.lib().glow_rpr_attention_backward(
                ctypes.byref(call), _lib.ptr(dout), _lib.ptr(probs), _lib.ptr(ds),
                _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv), _lib.ptr(dwk), _lib.ptr(dwv))
        _lib.check(rc, "glow_rpr_attention_backward")
        return (dq, dk, dv, dwk.view(ctx.wshape), dwv.view(ctx.wshape),
                None, None, None, None, None, None, None)


class _AttnRowsFn(torch.autograd.Function):
    """The same core on packed token rows: q, k, v, out are [rows, heads*d] (glow_attn_call.utt_off / ld)."""

    @staticmethod
    def forward(ctx, q, k, v, wk, wv, tr, lengths, heads, window, dropout, seed):
        rows, c = q.shape
        d = c // heads
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        wk2, wv2 = wk.reshape(-1, d).contiguous().float(), wv.reshape(-1, d).contiguous().float()
        dev = q.device
        call = _lib.AttnCall()
        call.q, call.k, call.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
        call.wk, call.wv = wk2.data_ptr(), wv2.data_ptr()
        call.lengths, call.mask = lengths.data_ptr(), None
        call.batch, call.heads, call.t, call.head_dim, call.window = tr.batch, heads, tr.t_max, d, window
        call.dropout, call.seed = float(dropout), int(seed)
        call.step_dev = _lib.step_counter_ptr(dev) if seed else None
        call.utt_off, call.ld = tr.rm.utt_off.data_ptr(), c
        out = torch.zeros_like(q)                                   # guard rows stay zero
        probs = torch.empty((tr.batch, heads, tr.t_max, tr.t_max), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call.stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _lib.lib().glow_rpr_attention_forward(ctypes.byref(call), _lib.ptr(out), _lib.ptr(probs), None)
        _lib.check(rc, "glow_rpr_attention_forward")
        ctx.call, ctx.keep, ctx.wshape = call, (q, k, v, wk2, wv2, lengths, probs, tr), wk.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, wk2, wv2, lengths, probs, tr = ctx.keep
        call, dev = ctx.call, q.device
        dout = dout.contiguous().float()
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        dwk, dwv = torch.empty_like(wk2), torch.empty_like(wv2)
        ds = torch.empty_like(probs)
        with torch.cuda.device(dev):
            call.stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _lib.lib().glow_rpr_attention_backward(
                ctypes.byref(call), _lib.ptr(dout), _lib.ptr(probs), _lib.ptr(ds),
                _lib.ptr(dq), _lib.ptr(dk), _lib.ptr(dv), _lib.ptr(dwk), _lib.ptr(dwv))
        _lib.check(rc, "glow_rpr_attention_backward")
        return dq, dk, dv, dwk.view(ctx.wshape), dwv.view(ctx.wshape), None, None, None, None, None, None


class RPR_Multihead_Attention(torch.nn.Module):
    _built = 0

    def __init__(self, query_channels, calc_channels, out_channels, num_heads,
                 relative_postion_clipping_distance=None, share_relative_postion_weight=True,
                 proximal_bias=False, block_mask_length=None, dropout_rate=0.0,
                 key_channels=None, value_channels=None):
        assert calc_channels % num_heads == 0, "calc_channels must be dividable by num_heads."
        super().__init__()
        if relative_postion_clipping_distance is None or not share_relative_postion_weight \
                or proximal_bias or block_mask_length is not None:
            raise _lib.GlowCoreError(
                "the sm_100a attention core implements the configuration Glow-TTS uses: shared "
                "relative-position weights, no proximal bias, no block mask (Modules.py:514-521)")
        self.num_heads = num_heads
        self.calc_channels_per_head = calc_channels // num_heads
        self.relative_postion_clipping_distance = relative_postion_clipping_distance
        self.dropout_rate = float(dropout_rate)
        self._calls = 0
        # position of this module among the attention modules built so far: a deterministic stand-in for the module's
        # identity in the dropout seed (id(self) would differ from run to run under one torch.manual_seed)
        self._index = RPR_Multihead_Attention._built
        RPR_Multihead_Attention._built += 1
        self.layer_Dict = torch.nn.ModuleDict()
        self.layer_Dict["Query"] = torch.nn.Conv1d(query_channels, calc_channels, 1)
        self.layer_Dict["Key"] = torch.nn.Conv1d(key_channels or query_channels, calc_channels, 1)
        self.layer_Dict["Value"] = torch.nn.Conv1d(value_channels or key_channels or query_channels, calc_channels, 1)
        for name in ("Query", "Key", "Value"):
            torch.nn.init.xavier_uniform_(self.layer_Dict[name].weight)
        self.layer_Dict["Projection"] = torch.nn.Conv1d(calc_channels, out_channels, 1)
        self.layer_Dict["Dropout"] = torch.nn.Dropout(p=dropout_rate)
        std = self.calc_channels_per_head ** -0.5
        n_rel = relative_postion_clipping_distance * 2 + 1
        self.weight_K = torch.nn.Parameter(torch.randn(1, n_rel, self.calc_channels_per_head) * std)
        self.weight_V = torch.nn.Parameter(torch.randn(1, n_rel, self.calc_channels_per_head) * std)

    def forward_rows(self, x, tr, lengths, packed=None):
        """Self-attention on packed token rows [rows, C] (rows.py): the four 1x1 convs run as tcgen05
        GEMMs over the rows, the attention core on [B, C, T] views of their outputs."""
        from . import rows as _rows
        d = self.layer_Dict
        q, k, v = (_rows.rows_conv(x, d[n], tr, x_masked=True, packed=packed) for n in ("Query", "Key", "Value"))
        seed = self._next_seed()
        lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
        out = _AttnRowsFn.apply(q, k, v, self.weight_K, self.weight_V, tr, lengths, self.num_heads,
                                self.relative_postion_clipping_distance, self.dropout_rate, seed)
        return _rows.rows_conv(out, d["Projection"], tr, x_masked=True, packed=packed)

    def _next_seed(self):
        if not (self.training and self.dropout_rate > 0):
            return 0
        self._calls += 1
        return ((int(torch.initial_seed()) * 0x2545F491 + self._calls * 0x9E3779B1 + (self._index % 65521) * 0x632BE5AB)
                & 0x7FFFFFFFFFFFFFFF) or 1

    def forward(self, queries, keys=None, values=None, masks=None, lengths=None, need_alignments=True):
        assert keys is None and values is None, "Relative position is for self-attention."
        _lib.require_cuda(queries, "queries")
        d = self.layer_Dict
        q = d["Query"](queries)
        k = d["Key"](queries)
        v = d["Value"](queries)
        seed = self._next_seed()
        if lengths is not None:
            lengths = lengths.to(device=q.device, dtype=torch.int32).contiguous()
        out, align = _AttnCoreFn.apply(q, k, v, self.weight_K, self.weight_V, lengths, masks, self.num_heads,
                                       self.relative_postion_clipping_distance, self.dropout_rate, seed,
                                       need_alignments)
        return d["Projection"](out), (align if need_alignments else None)
