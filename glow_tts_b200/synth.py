"""Synthetic batches with the reference collater's output contract (Datasets.py:225-250):
(tokens [B,T_x] int64 padded with <E>=1, token_lengths [B], mels [B,80,T_y] f32 padded
with -Max_Abs_Mel, mel_lengths [B] even, speakers [B]).  Length distributions follow
SURVEY.md 8(d): LJSpeech-shaped (config 2) and the LJ+VCTK mixture (config 3).  There is
no network for datasets, so bench.py / tests use these and say so."""
import torch


def _lj_lengths(g, n):
    mel = torch.clamp(616 + 215 * torch.randn(n, generator=g), 104, 946)
    mel = (mel / 2).floor().long() * 2
    txt = torch.clamp(torch.round(mel / 6.2 + 6 * torch.randn(n, generator=g)), 10, 200).long() + 2
    return mel, txt


def _vctk_lengths(g, n):
    mel = torch.clamp(280 + 110 * torch.randn(n, generator=g), 50, 800)
    mel = (mel / 2).floor().long() * 2
    txt = torch.clamp(torch.round(mel / 6.5 + 4 * torch.randn(n, generator=g)), 10, 200).long() + 2
    return mel, txt


def make_batch(kind="lj", batch=32, seed=0, n_tokens=35, n_speakers=109, mel_dim=80, force_max=True):
    """kind: 'lj' (config 2), 'ljvctk' (config 3), 'plumbing' (config 1, B=2)."""
    g = torch.Generator().manual_seed(int(seed))
    if kind == "plumbing":
        mel_len, txt_len = torch.tensor([400, 318]), torch.tensor([60, 45])
        speakers = torch.zeros(2, dtype=torch.long)
    elif kind == "lj":
        mel_len, txt_len = _lj_lengths(g, batch)
        if force_max:
            mel_len[0], txt_len[0] = 1000, min(int(txt_len.max()), 202)
        speakers = torch.zeros(batch, dtype=torch.long)
    elif kind == "ljvctk":
        n_lj = max(1, round(0.23 * batch))
        m1, t1 = _lj_lengths(g, n_lj)
        m2, t2 = _vctk_lengths(g, batch - n_lj)
        mel_len, txt_len = torch.cat([m1, m2]), torch.cat([t1, t2])
        speakers = torch.cat([torch.zeros(n_lj, dtype=torch.long),
                              torch.randint(1, n_speakers, (batch - n_lj,), generator=g)])
        perm = torch.randperm(batch, generator=g)
        mel_len, txt_len, speakers = mel_len[perm], txt_len[perm], speakers[perm]
    else:
        raise ValueError(kind)
    # a text can never be longer than its (squeezed-friendly) mel
    txt_len = torch.minimum(txt_len, mel_len // 2)
    b = len(mel_len)
    tx, ty = int(txt_len.max()), int(mel_len.max())
    tokens = torch.ones(b, tx, dtype=torch.long)
    mels = torch.full((b, mel_dim, ty), -4.0)
    for i in range(b):
        tl, ml = int(txt_len[i]), int(mel_len[i])
        tokens[i, :tl] = torch.randint(2, n_tokens, (tl,), generator=g)
        tokens[i, 0], tokens[i, tl - 1] = 0, 1
        mels[i, :, :ml] = torch.clamp(1.5 * torch.randn(mel_dim, ml, generator=g), -4, 4)
    return tokens, txt_len, mels, mel_len, speakers
