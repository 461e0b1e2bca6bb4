"""CPU: pin the MAS oracle (oracle/mas_oracle.c) against fixtures made by running
the reference (tests/golden/mas_*.npz, tools/make_golden.py), against the
reference's own Cython build when oracle/_ref is present, and against the
pure-Python twin (Modules.py:957-980)."""
import numpy as np
import pytest

from oracle import mas as omas
from tests._util import MAS_CASE_NAMES, GOLD, load_mas_case, path_to_pos, mas_values, rect_mask


@pytest.mark.parametrize("name", MAS_CASE_NAMES)
def test_port_matches_reference_fixture(name):
    c = load_mas_case(name)
    path = omas.maximum_path_numpy(c["value"], c["mask"], core="port")
    assert np.array_equal(path_to_pos(path), c["pos"])
    if c["path"] is not None:
        assert np.array_equal(path.astype(np.int8), c["path"])
    # every valid column holds exactly one 1, nothing outside the mask
    assert np.array_equal(path.sum(1), (np.arange(path.shape[2])[None, :] < c["t_y"][:, None]).astype(path.dtype))
    assert (path * (1 - c["mask"])).sum() == 0


@pytest.mark.parametrize("name", ["small_ragged", "ties"])
def test_python_twin_matches_fixture_and_port(name):
    c = load_mas_case(name)
    twin_ref = np.load("%s/mas_%s_pytwin.npz" % (GOLD, name))["path"]
    twin = omas.maximum_path_python(c["value"] * c["mask"], c["t_x"], c["t_y"])
    assert np.array_equal(twin.astype(np.int8), twin_ref)
    assert np.array_equal(twin.astype(np.int8), c["path"])      # -1e7 vs -1e9 sentinel: same paths


def test_port_equals_compiled_reference_random():
    if omas.ref_core() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(5)
    for trial in range(20):
        b = int(rng.integers(1, 6)); tx = int(rng.integers(1, 40)); ty = int(rng.integers(tx, 130))
        t_x = rng.integers(1, tx + 1, size=b); t_y = np.array([rng.integers(x, ty + 1) for x in t_x])
        v = mas_values(100 + trial, b, tx, ty, quant=(2.0 if trial % 3 == 0 else None))
        m = rect_mask(tx, ty, t_x, t_y)
        assert np.array_equal(omas.maximum_path_numpy(v, m, "port"), omas.maximum_path_numpy(v, m, "ref"))


def test_port_mutates_values_like_reference():
    """core.pyx works in place on `value`; the port keeps that (the DP table is the output of the forward pass)."""
    if omas.ref_core() is None:
        pytest.skip("oracle/_ref not built")
    v = mas_values(3, 2, 9, 30)
    a, b = v.copy(), v.copy()
    pa = np.zeros(v.shape, np.int32); pb = np.zeros(v.shape, np.int32)
    tx = np.array([9, 4], np.int32); ty = np.array([30, 11], np.int32)
    omas.maximum_path_c_port(pa, a, tx, ty)
    omas.ref_core().maximum_path_c(pb, b, tx, ty)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
