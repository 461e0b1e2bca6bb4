"""Run the tcgen05 self-test (glow_selftest_umma) over descriptor conventions.
Each variant runs in its own process under a timeout: a wrong convention shows
up as a numeric mismatch or a trapped launch, never as a hung box.

    python tools/probe_umma.py            # table of variants
    python tools/probe_umma.py one <lbo_is_slab> <shift> <bulk> <n> <k>
"""
import subprocess
import sys
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run_one(lbo_is_slab, shift, bulk, n, k):
    import torch
    from glow_tts_b200 import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    rows_a = 128 + 4
    a = (torch.randn(rows_a, k, device=dev) * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device=dev) * 0.5).to(torch.bfloat16)
    bp = b.view(n, k // 8, 8).permute(1, 0, 2).contiguous()          # [k/8][n][8] slab image
    d = torch.zeros(128, n, device=dev)
    slab_a, slab_b = rows_a * 16, n * 16
    if lbo_is_slab:
        la, sa, lb, sb = slab_a, 128, slab_b, 128
    else:
        la, sa, lb, sb = 128, slab_a, 128, slab_b
    rc = _lib.lib().glow_selftest_umma(_lib.ptr(a), _lib.ptr(bp), _lib.ptr(d), rows_a, k, n, shift,
                                       la, sa, lb, sb, int(bulk), _lib.stream_ptr())
    _lib.check(rc, "glow_selftest_umma")
    torch.cuda.synchronize()
    want = a[shift:shift + 128].float() @ b.float().t()
    err = (d - want).abs().max().item()
    print("RESULT lbo_is_slab=%d shift=%d bulk=%d n=%d k=%d max_abs_err=%.3e ref_max=%.3f" %
          (lbo_is_slab, shift, bulk, n, k, err, want.abs().max().item()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run_one(*[int(x) for x in sys.argv[2:7]])
        sys.exit(0)
    variants = [(1, 0, 0, 64, 64), (0, 0, 0, 64, 64), (1, 3, 0, 192, 192), (1, 0, 1, 192, 192),
                (1, 2, 1, 384, 192), (1, 4, 1, 160, 192), (1, 1, 1, 384, 80)]
    for v in variants:
        cmd = ["timeout", "120", sys.executable, os.path.abspath(__file__), "one"] + [str(x) for x in v]
        r = subprocess.run(cmd, capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(lines[0] if lines else "FAILED %s rc=%d %s" % (v, r.returncode, (r.stderr or "")[-300:].replace("\n", " | ")))
