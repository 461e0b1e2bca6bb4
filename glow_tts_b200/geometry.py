"""Batch geometry that lives on the device in FIXED buffers: what lets ONE captured CUDA graph train on ragged data.

The reference's loop feeds a different batch geometry every step (Datasets.py:225-250 pads to the longest item of
each batch; Train.py:582-584).  Everything the step derives from the per-utterance lengths -- the packed-row maps
of the decoder and of the text encoder (csrc/flow_layout.cuh), the length tensors the attention / alignment / loss
kernels read, the loss normalisers -- is therefore kept in ONE int32 device blob per *bucket*

    bucket = (batch, decoder rows_pad, encoder rows_pad, T_text_pad, T_mel_pad)

and refreshed by ONE pinned host->device copy before a replay (`StepGeometry.update`).  Shapes, grids and
pointers inside the captured step only depend on the bucket, so every batch that falls into the bucket replays the
same graph; rows beyond the batch's real rows are tail rows (row_utt = -1), which every kernel already treats as
padding.  Rounding the row counts up to the bucket granularity costs a few percent of dead rows.
"""
import numpy as np
import torch

from . import _lib

GUARD = 2
ROW_TILE = 128
DEC_ROW_BUCKET = 512       # decoder rows_pad granularity of a bucket (squeezed mel frames + guards)
ENC_ROW_BUCKET = 256       # encoder rows_pad granularity (tokens + guards)


def rows_needed(lengths):
    """Packed rows of a batch: a leading guard pair, then every utterance followed by a guard pair."""
    return GUARD + int(sum(int(n) + GUARD for n in lengths))


def round_up(n, q):
    return max(q, (int(n) + q - 1) // q * q)


def fill_row_arrays(lengths, rows_pad, row_utt, row_t, off, length):
    """Host side of flow.RowMap: write the four arrays (numpy views, int32) for `lengths`."""
    if rows_needed(lengths) > rows_pad:
        raise ValueError("batch needs %d packed rows, the buffers hold %d" % (rows_needed(lengths), rows_pad))
    row_utt[:] = -1
    row_t[:] = 0
    pos = GUARD
    for i, n in enumerate(lengths):
        n = int(n)
        off[i] = pos
        length[i] = n
        row_utt[pos:pos + n] = i
        row_t[pos:pos + n] = np.arange(n, dtype=np.int32)
        pos += n + GUARD


class _StaticRowMap:
    """flow.RowMap's interface over views of the geometry blob."""

    def __init__(self, batch, rows_pad, row_utt, row_t, utt_off, utt_len):
        self.batch, self.rows_pad = batch, rows_pad
        self.row_utt, self.row_t, self.utt_off, self.utt_len = row_utt, row_t, utt_off, utt_len
        self.rows_real = None          # varies per step: ask the host lengths


class _StaticTokenRows:
    """rows.TokenRows' interface over the geometry blob.  The gather indices between [B, T] and rows are derived
    from the row map ON THE DEVICE by `derive()`, which the step calls inside the captured region: a replay
    recomputes them from the refreshed blob."""

    def __init__(self, rm, t_max):
        self.rm, self.batch, self.t_max, self.rows_pad = rm, rm.batch, int(t_max), rm.rows_pad
        self.valid = self.src_idx = self.tmask = self.dst_idx = None

    def derive(self):
        rm = self.rm
        row_utt = rm.row_utt.long()
        valid = row_utt >= 0
        self.valid = valid.to(torch.float32).unsqueeze(1)
        self.src_idx = torch.where(valid, row_utt * self.t_max + rm.row_t.long(), torch.zeros_like(row_utt))
        t = torch.arange(self.t_max, device=row_utt.device)[None, :]
        lens = rm.utt_len.long()[:, None]
        self.tmask = (t < lens).to(torch.float32)
        self.dst_idx = torch.where(t < lens, rm.utt_off.long()[:, None] + t, torch.zeros_like(t)).reshape(-1)
        return self

    def pack(self, x_btc):
        b, t, c = x_btc.shape
        return x_btc.reshape(b * t, c).index_select(0, self.src_idx) * self.valid

    def unpack(self, rows):
        c = rows.shape[1]
        return rows.index_select(0, self.dst_idx).view(self.batch, self.t_max, c) * self.tmask.unsqueeze(2)


class StepGeometry:
    """Static device buffers for one bucket + the host staging ring that refreshes them."""

    SCALARS = 8        # float32: [1 / (B * T_x,max of THIS batch), w_mle, w_mse, 0.5*log(2pi)*(1 - w_mle), ...]

    def __init__(self, batch, t_text_pad, t_mel_pad, dec_rows_pad, enc_rows_pad, device, ring=4):
        assert dec_rows_pad % ROW_TILE == 0 and enc_rows_pad % ROW_TILE == 0
        self.batch, self.t_text, self.t_mel = int(batch), int(t_text_pad), int(t_mel_pad)
        self.dec_rows_pad, self.enc_rows_pad = int(dec_rows_pad), int(enc_rows_pad)
        self.device = torch.device(device)
        b, rd, re = self.batch, self.dec_rows_pad, self.enc_rows_pad
        # int32 blob layout
        names = [("tl", b), ("ml", b), ("sq", b), ("d_utt", rd), ("d_t", rd), ("d_off", b), ("d_len", b),
                 ("e_utt", re), ("e_t", re), ("e_off", b), ("e_len", b), ("scal", self.SCALARS)]
        self._slices, pos = {}, 0
        for name, n in names:
            self._slices[name] = (pos, pos + n)
            pos += (n + 3) // 4 * 4                       # 16-byte aligned pieces
        self.blob = torch.zeros(pos, dtype=torch.int32, device=self.device)
        self._ring = [torch.zeros(pos, dtype=torch.int32).pin_memory() for _ in range(ring)]
        self._ring_ev = [None] * ring
        self._ring_pos = 0
        v = lambda name: self.blob[self._slices[name][0]:self._slices[name][1]]
        self.tl32, self.ml32, self.sq32 = v("tl"), v("ml"), v("sq")
        self.scal = v("scal").view(torch.float32)
        self.dec_rm = _StaticRowMap(b, rd, v("d_utt"), v("d_t"), v("d_off"), v("d_len"))
        self.enc_rm = _StaticRowMap(b, re, v("e_utt"), v("e_t"), v("e_off"), v("e_len"))
        self.tok = _StaticTokenRows(self.enc_rm, self.t_text)
        self.host_tl = self.host_ml = None
        self.tl64 = self.ml64 = self.sq64 = None
        self.token_masks = self.mel_masks = None

    @property
    def key(self):
        return (self.batch, self.dec_rows_pad, self.enc_rows_pad, self.t_text, self.t_mel)

    @staticmethod
    def bucket_of(tl, ml, t_text_pad, t_mel_pad):
        """Bucket key of a batch with token lengths `tl` and (even) mel lengths `ml`."""
        return (len(tl), round_up(rows_needed([int(n) // 2 for n in ml]), DEC_ROW_BUCKET),
                round_up(rows_needed(tl), ENC_ROW_BUCKET), int(t_text_pad), int(t_mel_pad))

    def update(self, tl, ml, scalars=None):
        """Refresh the blob for a batch (host lists of token / mel lengths) -- stream-ordered, no sync except when
        the staging ring wraps onto a copy that is still in flight."""
        tl = [int(n) for n in tl]
        ml = [int(n) for n in ml]
        if len(tl) != self.batch or len(ml) != self.batch:
            raise ValueError("StepGeometry: batch of %d, built for %d" % (len(tl), self.batch))
        if max(tl) > self.t_text or max(ml) > self.t_mel:
            raise ValueError("StepGeometry: lengths exceed the padded sizes (%d, %d)" % (self.t_text, self.t_mel))
        i = self._ring_pos
        self._ring_pos = (i + 1) % len(self._ring)
        if self._ring_ev[i] is not None:
            self._ring_ev[i].synchronize()
        host = self._ring[i]
        h = host.numpy()
        s = self._slices
        view = lambda name: h[s[name][0]:s[name][1]]
        view("tl")[:] = tl
        view("ml")[:] = ml
        sq = [n // 2 for n in ml]
        view("sq")[:] = sq
        fill_row_arrays(sq, self.dec_rows_pad, view("d_utt"), view("d_t"), view("d_off"), view("d_len"))
        fill_row_arrays(tl, self.enc_rows_pad, view("e_utt"), view("e_t"), view("e_off"), view("e_len"))
        sc = view("scal").view(np.float32)
        sc[:] = 0.0
        sc[0] = 1.0 / float(self.batch * max(tl))         # MSELoss averages over B * T_x,max (Train.py:210)
        sc[1] = sc[2] = 1.0
        if scalars is not None:
            for j, val in enumerate(scalars):
                sc[j] = float(val)
        self.blob.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._ring_ev[i] = ev
        self.host_tl, self.host_ml = tl, ml
        return self

    def derive(self):
        """Device-side derived tensors (int64 lengths, masks, token gather indices).  Called INSIDE the step, so a
        captured graph recomputes them from the refreshed blob on every replay."""
        dev = self.device
        self.tl64, self.ml64, self.sq64 = self.tl32.long(), self.ml32.long(), self.sq32.long()
        self.token_masks = (torch.arange(self.t_text, device=dev)[None, :] < self.tl64[:, None]).unsqueeze(1).float()
        self.mel_masks = (torch.arange(self.t_mel, device=dev)[None, :] < self.ml64[:, None]).unsqueeze(1).float()
        self.tok.derive()
        return self
