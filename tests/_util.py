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
