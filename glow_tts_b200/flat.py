"""Flat parameter / gradient storage.

libglowcore reads the decoder's parameters from ONE flat fp32 buffer through an
offset table, the data-parallel step all-reduces ONE flat gradient buffer
 and the fused RAdam kernel walks the same two buffers.  The
nn.Parameters the reference's surface exposes (state_dict keys, .parameters())
stay what they are -- they just become views into the flat buffer.
"""
import numpy as np
import torch


class FlatBuffer:
    def __init__(self, params):
        params = list(params)
        if not params:
            raise ValueError("FlatBuffer needs at least one parameter")
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("all parameters must be fp32 on one device")
        self.params = params
        sizes = [p.numel() for p in params]
        # 4-element (16 B) alignment of every tensor inside the buffer
        offs, pos = [], 0
        for n in sizes:
            offs.append(pos)
            pos += (n + 3) // 4 * 4
        self.offsets = offs
        self.total = pos
        self.data = torch.zeros(pos, dtype=torch.float32, device=dev)
        self.grad = None
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.data[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self._index = {id(p): i for i, p in enumerate(params)}

    # -- validity (module.to()/cuda() re-creates p.data and silently breaks the views)
    def valid(self):
        base = self.data.data_ptr()
        for i in (0, len(self.params) // 2, len(self.params) - 1):
            if self.params[i].data_ptr() != base + 4 * self.offsets[i]:
                return False
        return True

    def offset_of(self, p):
        return self.offsets[self._index[id(p)]]

    def contains(self, p):
        return id(p) in self._index

    # -- gradients
    def attach_grads(self):
        """Make every p.grad a view into one flat gradient buffer (allocated on first use)."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.data)
        base = self.grad.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        return self.grad

    def grads_attached(self, params):
        if self.grad is None:
            return False
        base = self.grad.data_ptr()
        for p in params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * self.offset_of(p):
                return False
        return True

    def zero_grad(self):
        self.attach_grads().zero_()


def offset_table(flat, slots):
    """int64 host table of element offsets of `slots` (list of Parameters) in flat.data."""
    return np.ascontiguousarray([flat.offset_of(p) for p in slots], dtype=np.int64)
