#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
#
# Compiles the reference's ONLY native source, monotonic_align/core.pyx
# (/root/reference/monotonic_align/core.pyx:1-45), from where it lies under
# /root/reference into oracle/_ref/ (git-ignored, but it travels to the GPU box
# with the gpurun snapshot).  No reference source is copied into the repo: the
# Cython-generated C and the .so land in oracle/_ref/ only.
#
# Same flags as the reference's own setup.py (monotonic_align/setup.py:5-9):
# plain cythonize, numpy include dir, NO -fopenmp -> prange runs serially.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GLOW_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"

if [ ! -f "$REF/monotonic_align/core.pyx" ]; then
    echo "build_ref: $REF/monotonic_align/core.pyx not found (GPU box?) - keeping prebuilt files" >&2
    exit 0
fi
mkdir -p "$OUT"
"$PY" -m cython -3 "$REF/monotonic_align/core.pyx" -o "$OUT/core.c" >/dev/null
INC_PY="$("$PY" -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
INC_NP="$("$PY" -c 'import numpy; print(numpy.get_include())')"
EXT="$("$PY" -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
gcc -O2 -fPIC -shared -fwrapv -fno-strict-aliasing -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION \
    -I"$INC_PY" -I"$INC_NP" "$OUT/core.c" -o "$OUT/core$EXT"
echo "build_ref: built $OUT/core$EXT"
