"""Drop-in for the reference's ``monotonic_align`` package.

``maximum_path(value, mask)`` keeps the contract of
/root/reference/monotonic_align/__init__.py:6-21 (path comes back on value's
device in value's dtype, values in {0,1}) but never leaves the GPU: the DP and
the backtrack run in libglowcore's ``glow_mas_forward`` (csrc/mas.cu).

``maximum_path_c`` keeps the contract of core.pyx:40 on host NumPy buffers
(int32 paths overwritten in place) through ``glow_mas_forward_host``.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def maximum_path(value, mask=None, t_x=None, t_y=None, max_neg_val=-1e9, out_dtype=None):
    """value [B,T_x,T_y]; mask [B,T_x,T_y] 0/1 (or int32 device lengths t_x/t_y [B]).

    The reference multiplies value by mask first (__init__.py:13); inside the
    band every mask entry is 1 and cells outside it are never read, so the
    product is skipped -- results are identical.
    """
    _lib.require_cuda(value, "value")
    if value.dim() != 3:
        raise ValueError("value must be [batch, t_x, t_y]")
    if (t_x is None) != (t_y is None):
        raise ValueError("pass both t_x and t_y or neither")
    if t_x is None and mask is None:
        raise ValueError("maximum_path needs mask or (t_x, t_y)")
    dtype = value.dtype
    v = value.detach()
    if v.dtype != torch.float32:          # __init__.py:14 .astype(np.float32)
        v = v.float()
    v = v.contiguous()
    b, tx, ty = v.shape
    m = None
    if t_x is None:
        if mask.shape != value.shape:
            raise ValueError("mask must have value's shape")
        m = mask.detach().to(device=v.device, dtype=torch.float32).contiguous()
    else:
        t_x = t_x.to(device=v.device, dtype=torch.int32).contiguous()
        t_y = t_y.to(device=v.device, dtype=torch.int32).contiguous()
    want = out_dtype or dtype
    if want == torch.int32:
        path, tag = torch.empty((b, tx, ty), dtype=torch.int32, device=v.device), _lib.GLOW_I32
    else:
        path, tag = torch.empty((b, tx, ty), dtype=torch.float32, device=v.device), _lib.GLOW_F32
    with torch.cuda.device(v.device):
        rc = _lib.lib().glow_mas_forward(
            _lib.ptr(v), _lib.ptr(m), _lib.ptr(t_x), _lib.ptr(t_y), b, tx, ty,
            _lib.ptr(path), tag, ctypes.c_float(max_neg_val), None, 0, _lib.stream_ptr(v.device))
    _lib.check(rc, "glow_mas_forward")
    return path if path.dtype == want else path.to(want)


def maximum_path_align(value, t_x, t_y, max_neg_val=-1e9):
    """maximum_path plus what its backtrack knows anyway: -> (path f32 [B,T_x,T_y], frame_token i32 [B,T_y]
    (row of the path in every column), durations i32 [B,T_x] (= path.sum(-1))).  t_x / t_y: device lengths."""
    _lib.require_cuda(value, "value")
    if value.dim() != 3:
        raise ValueError("value must be [batch, t_x, t_y]")
    v = value.detach()
    if v.dtype != torch.float32:
        v = v.float()
    v = v.contiguous()
    b, tx, ty = v.shape
    t_x = t_x.to(device=v.device, dtype=torch.int32).contiguous()
    t_y = t_y.to(device=v.device, dtype=torch.int32).contiguous()
    path = torch.empty((b, tx, ty), dtype=torch.float32, device=v.device)
    tok = torch.empty((b, ty), dtype=torch.int32, device=v.device)
    dur = torch.empty((b, tx), dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        rc = _lib.lib().glow_mas_align(_lib.ptr(v), _lib.ptr(t_x), _lib.ptr(t_y), b, tx, ty, _lib.ptr(path), _lib.GLOW_F32,
                                       ctypes.c_float(max_neg_val), _lib.ptr(tok), _lib.ptr(dur), _lib.stream_ptr(v.device))
    _lib.check(rc, "glow_mas_align")
    return path, tok, dur


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9, device=0):
    """core.pyx:40 contract on host buffers: paths int32 [b,t_x,t_y] overwritten.
    Unlike the Cython core, `values` is left untouched (its mutation there is a
    side effect of working in place on the wrapper's private copy)."""
    for name, a, dt in (("paths", paths, np.int32), ("values", values, np.float32),
                        ("t_xs", t_xs, np.int32), ("t_ys", t_ys, np.int32)):
        if not isinstance(a, np.ndarray) or a.dtype != dt or not a.flags.c_contiguous:
            # typed memoryviews in core.pyx raise ValueError on exactly these
            raise ValueError("%s must be a C-contiguous %s array" % (name, np.dtype(dt).name))
    if paths.shape != values.shape or values.ndim != 3:
        raise ValueError("paths/values must both be [b, t_x, t_y]")
    b, tx, ty = values.shape
    rc = _lib.lib().glow_mas_forward_host(
        paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data, t_ys.ctypes.data,
        b, tx, ty, ctypes.c_float(max_neg_val), int(device))
    _lib.check(rc, "glow_mas_forward_host")
