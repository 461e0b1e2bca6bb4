"""tools/make_golden.py -- regenerate tests/golden/*.npz by RUNNING THE REFERENCE.

Only runs in the build container (needs /root/reference).  It copies nothing
from the reference into the repo: the reference tree is copied to a temp dir,
its Cython MAS is built there with its own setup.py, Hyper_Parameters.yaml is
patched (Mode / Device '-1' / Use_Cython_Alignment) and the reference modules
are imported and executed on seeded synthetic inputs.  Inputs that are cheap to
regenerate are stored as seeds + checksums; outputs are stored as arrays.

    python tools/make_golden.py [mas] [model]
"""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GLOW_REFERENCE_DIR", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def stage_reference(mode="Vanilla", cython=True):
    """Temp copy of the reference with a patched yaml and a built core.pyx."""
    tmp = tempfile.mkdtemp(prefix="glowref_")
    dst = os.path.join(tmp, "ref")
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("Figures", "Wav_for_Inference", "*.ipynb", ".git"))
    y = open(os.path.join(dst, "Hyper_Parameters.yaml"), encoding="utf-8").read()
    y = y.replace("Mode: 'SE'    #Vanilla, SE, PE, GR", "Mode: '%s'" % mode)
    y = y.replace("Device: '0'", "Device: '-1'")
    y = y.replace("Use_Cython_Alignment: true", "Use_Cython_Alignment: %s" % ("true" if cython else "false"))
    assert "Mode: '%s'" % mode in y and "Device: '-1'" in y
    open(os.path.join(dst, "Hyper_Parameters.yaml"), "w", encoding="utf-8").write(y)
    ma = os.path.join(dst, "monotonic_align")
    os.makedirs(os.path.join(ma, "monotonic_align"), exist_ok=True)
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=ma, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return tmp, dst


from tests._util import checksum, mas_values, path_to_pos, rect_mask  # noqa: E402


MAS_CASES = {
    # name: (seed, B, Tx, Ty, t_x list or None(full), t_y list or None(full), quant)
    "small_ragged": (11, 6, 24, 60, [24, 10, 1, 17, 5, 24], [60, 44, 9, 17, 60, 24], None),
    "ties": (12, 4, 20, 48, [20, 13, 7, 20], [48, 30, 48, 21], 4.0),
    "mid_ragged": (13, 8, 75, 330, [75, 40, 61, 33, 75, 12, 50, 70], [330, 200, 296, 150, 76, 330, 120, 322], None),
    "lj_shaped": (14, 4, 200, 1200, [200, 150, 111, 37], [1200, 946, 690, 250], None),
    "ties_big": (15, 3, 120, 700, [120, 90, 33], [700, 512, 700], 8.0),
}


def make_mas():
    tmp, dst = stage_reference("Vanilla", True)
    cwd = os.getcwd()
    try:
        os.chdir(dst)
        sys.path.insert(0, dst)
        import torch
        import monotonic_align as ref_ma          # the reference package (Cython core)
        import Modules as ref_modules              # for the pure-Python twin
        for name, (seed, b, tx, ty, txs, tys, quant) in MAS_CASES.items():
            v = mas_values(seed, b, tx, ty, quant)
            txs = np.array(txs or [tx] * b, np.int32)
            tys = np.array(tys or [ty] * b, np.int32)
            mask = rect_mask(tx, ty, txs, tys)
            path = ref_ma.maximum_path(torch.from_numpy(v), torch.from_numpy(mask)).numpy()
            out = dict(seed=seed, shape=np.array([b, tx, ty]), t_x=txs, t_y=tys, quant=quant or 0.0,
                       value_sha=checksum(v), pos=path_to_pos(path))
            if b * tx * ty <= 10000:
                out["value"] = v
                out["path"] = path.astype(np.int8)
            np.savez_compressed(os.path.join(GOLD, "mas_%s.npz" % name), **out)
            print("mas", name, path.shape, "ones", int(path.sum()))
        # Python twin (Modules.py:934-980) on the small cases, bypassing __init__'s Cython rebinding
        gen = torch.nn.Module.__new__(ref_modules.Maximum_Path_Generater)
        torch.nn.Module.__init__(gen)
        for name in ("small_ragged", "ties"):
            seed, b, tx, ty, txs, tys, quant = MAS_CASES[name]
            v = mas_values(seed, b, tx, ty, quant)
            txs = np.array(txs, np.int32); tys = np.array(tys, np.int32)
            mask = rect_mask(tx, ty, txs, tys)
            twin = ref_modules.Maximum_Path_Generater.forward(gen, torch.from_numpy(v.copy()), torch.from_numpy(mask)).numpy()
            np.savez_compressed(os.path.join(GOLD, "mas_%s_pytwin.npz" % name), path=twin.astype(np.int8))
            print("mas twin", name, int(twin.sum()))
    finally:
        os.chdir(cwd)
        sys.path.remove(dst)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["mas", "model"]
    os.makedirs(GOLD, exist_ok=True)
    if "mas" in what:
        make_mas()
    if "model" in what:
        from tools import make_golden_model
        make_golden_model.main()
