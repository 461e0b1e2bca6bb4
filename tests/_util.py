"""Shared test helpers: seeded input generators used both by tools/make_golden.py
(when the fixtures are produced with the reference) and by the tests (when the
same inputs are regenerated on the GPU box, where /root/reference is absent)."""
import hashlib
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")


def checksum(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def mas_values(seed, b, tx, ty, quant=None):
    """log_P-like values, N(-113, 6^2); quant>0 snaps them
    to a coarse grid so that exact ties are frequent (strict-'<' rule)."""
    rng = np.random.default_rng(int(seed))
    v = rng.normal(-113.0, 6.0, size=(b, tx, ty)).astype(np.float32)
    if quant:
        v = (np.round(v / quant) * quant).astype(np.float32)
    return v


def rect_mask(tx, ty, t_xs, t_ys):
    t_xs = np.asarray(t_xs)
    t_ys = np.asarray(t_ys)
    return ((np.arange(tx)[None, :, None] < t_xs[:, None, None]) &
            (np.arange(ty)[None, None, :] < t_ys[:, None, None])).astype(np.float32)


def path_to_pos(path):
    """[B,Tx,Ty] 0/1 -> int16 [B,Ty]: row of the single 1 per column, -1 if none."""
    path = np.asarray(path)
    pos = np.full((path.shape[0], path.shape[2]), -1, np.int16)
    for b in range(path.shape[0]):
        cols = path[b].sum(0)
        assert set(np.unique(cols).tolist()) <= {0, 1}, "more than one 1 in a column"
        pos[b, cols == 1] = path[b].argmax(0)[cols == 1]
    return pos


def load_mas_case(name):
    g = np.load(os.path.join(GOLD, "mas_%s.npz" % name))
    b, tx, ty = (int(v) for v in g["shape"])
    quant = float(g["quant"]) or None
    value = mas_values(int(g["seed"]), b, tx, ty, quant)
    assert checksum(value) == str(g["value_sha"]), "regenerated MAS input differs from the fixture's"
    return dict(value=value, t_x=g["t_x"], t_y=g["t_y"], pos=g["pos"],
                mask=rect_mask(tx, ty, g["t_x"], g["t_y"]),
                path=(g["path"] if "path" in g.files else None))


MAS_CASE_NAMES = ["small_ragged", "ties", "mid_ragged", "lj_shaped", "ties_big"]


# --------------------------------------------------------------------------- #
# model fixtures
# --------------------------------------------------------------------------- #
def synth_state_dict(ref_sd, seed):
    """Deterministic, non-degenerate values for every entry of a reference-layout
    state_dict (same keys/shapes as `ref_sd`), independent of any constructor's init:
    End convs are non-zero (the coupling is not the identity), ActNorm has non-trivial
    logs/bias, 4x4 mixes are well conditioned with det > 0."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    out = {}
    for key in sorted(ref_sd.keys()):
        shape = tuple(ref_sd[key].shape)
        r = torch.randn(shape, generator=g)
        if key.endswith("layers.1.weight"):                       # 4x4 invertible conv
            w = torch.eye(shape[0]) + 0.25 * r
            if torch.det(w) < 0:
                w[:, 0] = -w[:, 0]
            v = w
        elif key.endswith("weight_g"):
            v = 0.6 + 0.4 * torch.rand(shape, generator=g)
        elif key.endswith("weight_v"):
            fan_in = shape[1] * shape[2]
            v = r / fan_in ** 0.5
        elif key.endswith("layers.0.logs"):
            v = 0.15 * r
        elif key.endswith("layers.0.bias"):
            v = 0.2 * r
        elif "LayerNorm" in key and key.endswith("weight"):
            v = 1.0 + 0.1 * r
        elif key.endswith("End.weight"):
            v = 0.4 * r / shape[1] ** 0.5
        elif key.endswith("bias"):
            v = 0.05 * r
        elif key.endswith("weight_K") or key.endswith("weight_V"):
            v = r * shape[-1] ** -0.5
        elif key.endswith("Embedding.weight"):
            v = r * shape[1] ** -0.5
        elif key.endswith("LUT.weight"):
            v = 2 * torch.rand(shape, generator=g) - 1
        elif len(shape) == 3:                                      # plain conv weights
            v = r / (shape[1] * shape[2]) ** 0.5
        else:
            v = 0.1 * r
        out[key] = v.float().contiguous()
    return out


def synth_batch(seed, token_lengths, mel_lengths, n_tokens=35, n_speakers=109, mel_dim=80):
    """Collater-shaped batch (Datasets.py:225-250): tokens padded with <E>=1, mels padded with -4."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    b = len(token_lengths)
    tx, ty = max(token_lengths), max(mel_lengths)
    tokens = torch.ones(b, tx, dtype=torch.long)
    mels = torch.full((b, mel_dim, ty), -4.0)
    for i, (tl, ml) in enumerate(zip(token_lengths, mel_lengths)):
        tokens[i, :tl] = torch.randint(2, n_tokens, (tl,), generator=g)
        tokens[i, 0], tokens[i, tl - 1] = 0, 1
        mels[i, :, :ml] = torch.clamp(1.5 * torch.randn(mel_dim, ml, generator=g), -4, 4)
    speakers = torch.randint(0, n_speakers, (b,), generator=g)
    return (tokens, torch.tensor(token_lengths), mels, torch.tensor(mel_lengths), speakers)


def rel_err(a, b):
    """max |a-b| / max |b| (the tolerance metric used throughout: 1e-3 per north_star)."""
    import torch
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
