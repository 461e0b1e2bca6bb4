"""GPU: glow_mas_forward (csrc/mas.cu) through the C ABI vs the oracle and the
reference-made fixtures.  Bit-exact: integer path indices must be identical."""
import numpy as np
import pytest
import torch

from tests._util import MAS_CASE_NAMES, load_mas_case, mas_values, path_to_pos, rect_mask

pytestmark = pytest.mark.gpu


def _gpu_path(value, mask=None, t_x=None, t_y=None, **kw):
    from glow_tts_b200.monotonic_align import maximum_path
    dev = torch.device("cuda:0")
    v = torch.from_numpy(value).to(dev)
    m = None if mask is None else torch.from_numpy(mask).to(dev)
    tx = None if t_x is None else torch.as_tensor(t_x, dtype=torch.int32, device=dev)
    ty = None if t_y is None else torch.as_tensor(t_y, dtype=torch.int32, device=dev)
    out = maximum_path(v, m, tx, ty, **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", MAS_CASE_NAMES)
@pytest.mark.parametrize("lengths_from", ["mask", "arrays"])
def test_matches_reference_fixture(name, lengths_from):
    c = load_mas_case(name)
    if lengths_from == "mask":
        out = _gpu_path(c["value"], mask=c["mask"])
    else:
        out = _gpu_path(c["value"], t_x=c["t_x"], t_y=c["t_y"])
    assert out.dtype == torch.float32 and out.is_cuda
    path = out.cpu().numpy()
    assert set(np.unique(path).tolist()) <= {0.0, 1.0}
    assert np.array_equal(path_to_pos(path), c["pos"])
    if c["path"] is not None:
        assert np.array_equal(path.astype(np.int8), c["path"])


def test_random_ragged_vs_oracle_all_lane_widths():
    from oracle import mas as omas
    rng = np.random.default_rng(7)
    for trial, tx in enumerate([1, 2, 31, 32, 33, 95, 96, 97, 159, 161, 202, 224, 225, 256]):
        b = 5
        ty = int(rng.integers(tx, tx + 300))
        t_x = rng.integers(1, tx + 1, size=b); t_x[0] = tx
        t_y = np.array([rng.integers(x, ty + 1) for x in t_x]); t_y[0] = ty
        v = mas_values(200 + trial, b, tx, ty, quant=(4.0 if trial % 2 else None))
        m = rect_mask(tx, ty, t_x, t_y)
        want = omas.maximum_path_numpy(v, m, "port")
        got = _gpu_path(v, t_x=t_x, t_y=t_y, out_dtype=torch.int32).cpu().numpy()
        assert np.array_equal(got, want), "tx=%d" % tx


def test_full_size_properties_config5():
    """BASELINE config 5 at full size (B=256, 200x1200): size-independent properties
    + exact agreement with the oracle on a sample of utterances."""
    from oracle import mas as omas
    b, tx, ty = 256, 200, 1200
    v = mas_values(1, b, tx, ty)
    out = _gpu_path(v, t_x=[tx] * b, t_y=[ty] * b, out_dtype=torch.int32)
    assert int(out.sum()) == b * ty                                   # one 1 per column
    assert torch.equal(out.sum(1), torch.ones(b, ty, dtype=torch.int64, device=out.device))
    pos = out.argmax(1)                                               # [B,Ty]
    step = pos[:, 1:] - pos[:, :-1]
    assert int(step.min()) >= 0 and int(step.max()) <= 1              # monotone, no skips
    assert bool((pos[:, 0] == 0).all()) and bool((pos[:, -1] == tx - 1).all())
    sample = [0, 17, 255]
    want = omas.maximum_path_numpy(v[sample], np.ones((3, tx, ty), np.float32), "port")
    assert np.array_equal(out[sample].cpu().numpy(), want)


def test_value_not_modified_and_padding_invariance():
    c = load_mas_case("mid_ragged")
    dev = torch.device("cuda:0")
    from glow_tts_b200.monotonic_align import maximum_path
    v = torch.from_numpy(c["value"]).to(dev)
    keep = v.clone()
    p1 = maximum_path(v, torch.from_numpy(c["mask"]).to(dev))
    assert torch.equal(v, keep)
    # garbage in the padded region must not change the path
    noisy = v + (1 - torch.from_numpy(c["mask"]).to(dev)) * 1e6
    p2 = maximum_path(noisy, torch.from_numpy(c["mask"]).to(dev))
    assert torch.equal(p1, p2)


def test_host_entry_point_matches_core_pyx_contract():
    from glow_tts_b200.monotonic_align import maximum_path_c
    c = load_mas_case("small_ragged")
    paths = np.full(c["value"].shape, 7, np.int32)            # overwritten, not accumulated
    maximum_path_c(paths, c["value"].copy(), c["t_x"].astype(np.int32), c["t_y"].astype(np.int32))
    assert np.array_equal(paths.astype(np.int8), c["path"])
    with pytest.raises(ValueError):
        maximum_path_c(paths.astype(np.int64), c["value"], c["t_x"], c["t_y"])


def test_degenerate_inputs():
    from glow_tts_b200.monotonic_align import maximum_path
    dev = torch.device("cuda:0")
    # t_x > t_y is outside the reference's defined behaviour -> all-zero plane, no crash
    v = torch.randn(2, 8, 6, device=dev)
    out = maximum_path(v, t_x=torch.tensor([8, 3], dtype=torch.int32), t_y=torch.tensor([6, 6], dtype=torch.int32))
    assert float(out[0].sum()) == 0 and float(out[1].sum()) == 6
    assert maximum_path(torch.zeros(0, 4, 4, device=dev), torch.zeros(0, 4, 4, device=dev)).shape == (0, 4, 4)
    with pytest.raises(Exception):
        maximum_path(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))    # CPU tensor: no fallback
