"""CPU: libglowcore.so builds for sm_100a without a GPU, loads, and exports every
symbol include/glowcore.h declares; the product never routes through oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

from glow_tts_b200 import _lib
from tests._util import REPO


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    handle = ctypes.CDLL(built_lib)
    declared = _lib.header_symbols()
    assert declared, "no glow_* declarations parsed from include/glowcore.h"
    for name in declared:
        assert hasattr(handle, name), "%s declared in glowcore.h but not exported" % name
        assert name in _lib.SIGNATURES, "%s has no ctypes signature in _lib.py" % name
    for name in _lib.SIGNATURES:
        assert name in declared, "%s bound in _lib.py but missing from glowcore.h" % name


def test_abi_version_and_error_text(built_lib):
    l = _lib.lib()
    assert l.glow_abi_version() >= 1
    # argument validation runs before any CUDA call, so it is testable without a GPU
    rc = l.glow_mas_forward(None, None, None, None, 1, 4, 4, None, 0, ctypes.c_float(-1e9), None, 0, None)
    assert rc == -1 and b"null" in l.glow_last_error()
    rc = l.glow_mas_forward(None, None, None, None, -1, 4, 4, None, 0, ctypes.c_float(-1e9), None, 0, None)
    assert rc == -1


def test_sass_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_gemm_kernels_issue_tcgen05_from_an_elected_lane(built_lib):
    """The decoder / encoder GEMM kernel and the layer kernel are Blackwell-native (tcgen05.mma, tcgen05.ld, bulk copies)
    and issue from an elected lane of a converged warp: issued from `if (lane == 0)` the compiler wraps every
    UTCHMMA / UBLKCP in an ELECT ... BRA.U.ANY loop (umma.cuh: elect_one, profiles/layer_timeline_r02y.md)."""
    obj = os.path.join(os.path.dirname(built_lib), "_obj", "flow_tc.o")
    if not os.path.exists(obj):
        pytest.skip("object files not kept")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "tc_layer_kernel" in sass and "tc_gemm3_kernel" in sass
    for mnemonic in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "ENL2.256" in sass                      # the layer kernel's one-row-per-thread 256-bit global accesses
    assert "BRA.U.ANY" not in sass


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "glow_tts_b200")
    bad = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|from\s+\.\.?oracle|oracle/", text, flags=re.M):
                    bad.append(os.path.join(root, f))
    assert not bad, "product files reference oracle/: %s" % bad
