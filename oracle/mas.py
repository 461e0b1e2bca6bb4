"""oracle/mas.py -- TEST INFRASTRUCTURE.

Python face of the MAS oracle:

* ``maximum_path_c_port``  -- ctypes call into oracle/_build/libmas_oracle.so
  (the plain-C restatement in mas_oracle.c of
  /root/reference/monotonic_align/core.pyx:9-45).
* ``maximum_path_ref_core`` -- the reference's own Cython build in oracle/_ref
  (made by oracle/build_ref.sh from /root/reference/monotonic_align/core.pyx),
  when present.
* ``maximum_path_numpy``   -- restatement of the wrapper
  /root/reference/monotonic_align/__init__.py:6-21 around either core.
* ``maximum_path_python``  -- restatement of the pure-Python twin
  /root/reference/Modules.py:957-980 (sentinel -1e7, clamped index).

Parity is pinned by tests/test_oracle_mas.py against tests/golden/mas_*.npz,
which tools/make_golden.py produced by running the reference in this container.
"""
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile mas_oracle.c (gcc) and, if /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE, "_build/libmas_oracle.so"], check=True)
    subprocess.run(["bash", os.path.join(_HERE, "build_ref.sh")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libmas_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE, "_build/libmas_oracle.so"], check=True)
        lib = ctypes.CDLL(path)
        lib.mas_oracle_batch.restype = None
        lib.mas_oracle_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        _LIB = lib
    return _LIB


def maximum_path_c_port(paths, values, t_xs, t_ys, max_neg_val=-1e9):
    """Same contract as core.pyx:40 maximum_path_c: int32 paths (zeros in),
    float32 values (mutated in place), int32 lengths."""
    assert paths.dtype == np.int32 and values.dtype == np.float32
    assert paths.flags.c_contiguous and values.flags.c_contiguous
    t_xs = np.ascontiguousarray(t_xs, dtype=np.int32)
    t_ys = np.ascontiguousarray(t_ys, dtype=np.int32)
    b, tx, ty = values.shape
    _lib().mas_oracle_batch(paths.ctypes.data, values.ctypes.data,
                            t_xs.ctypes.data, t_ys.ctypes.data, b, tx, ty,
                            ctypes.c_float(max_neg_val))


def ref_core():
    """The reference's compiled Cython module (oracle/_ref/core*.so) or None."""
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(ref_dir):
        return None
    for name in sorted(os.listdir(ref_dir)):
        if name.startswith("core.") and name.endswith(".so"):
            spec = importlib.util.spec_from_file_location("core", os.path.join(ref_dir, name))
            try:
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod
            except ImportError:
                return None
    return None


def maximum_path_numpy(value, mask, core="port"):
    """monotonic_align/__init__.py:6-21 on numpy arrays ([b,t_x,t_y] each)."""
    value = (value * mask).astype(np.float32)                       # :13-16
    path = np.zeros_like(value).astype(np.int32)                    # :17
    t_x_max = mask.sum(1)[:, 0].astype(np.int32)                    # :20
    t_y_max = mask.sum(2)[:, 0].astype(np.int32)                    # :21
    value = np.ascontiguousarray(value)
    if core == "port":
        maximum_path_c_port(path, value, t_x_max, t_y_max)
    else:
        mod = ref_core()
        if mod is None:
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.sh)")
        mod.maximum_path_c(path, value, t_x_max, t_y_max)
    return path


def maximum_path_python(log_p, token_lengths, mel_lengths, neg=-1e7):
    """Modules.py:957-980, one utterance at a time, pure Python (small cases)."""
    out = []
    for x, tl, ml in zip(log_p, token_lengths, mel_lengths):
        x = np.array(x, dtype=np.float32, copy=True)
        tl, ml = int(tl), int(ml)
        path = np.zeros(x.shape, dtype=np.int32)
        for m in range(ml):
            for t in range(max(0, tl + m - ml), min(tl, m + 1)):
                cur = np.float32(neg) if m == t else x[t, m - 1]
                if t == 0:
                    prev = np.float32(0.0) if m == 0 else np.float32(neg)
                else:
                    prev = x[t - 1, m - 1]
                x[t, m] = max(cur, prev) + x[t, m]
        t = tl - 1
        for m in range(ml - 1, -1, -1):
            path[t, m] = 1
            if t == m or x[t, m - 1] < x[t - 1, m - 1]:
                t = max(0, t - 1)
        out.append(path)
    return np.stack(out, 0)


if __name__ == "__main__":
    build()
    print("oracle built:", os.listdir(os.path.join(_HERE, "_build")), file=sys.stderr)
