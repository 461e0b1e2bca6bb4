"""MN-major tcgen05 descriptor probe (glow_selftest_umma_mn): which of (lbo, sbo) is the K-direction
(128 B) and which the MN-direction (plane pitch) stride.  Each variant in its own process."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run_one(variant, r, n):
    import torch
    from glow_tts_b200 import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    a = (torch.randn(r, 128, device=dev) * 0.5).to(torch.bfloat16)
    d = (torch.randn(r, n, device=dev) * 0.5).to(torch.bfloat16)
    ac = a.view(r, 16, 8).permute(1, 0, 2).contiguous()          # [128/8][r][8]
    dc = d.view(r, n // 8, 8).permute(1, 0, 2).contiguous()
    c = torch.zeros(128, n, device=dev)
    plane = r * 16
    lbo, sbo = (128, plane) if variant == 0 else (plane, 128)
    rc = _lib.lib().glow_selftest_umma_mn(_lib.ptr(ac), _lib.ptr(dc), _lib.ptr(c), r, n, lbo, sbo, _lib.stream_ptr())
    _lib.check(rc, "glow_selftest_umma_mn")
    torch.cuda.synchronize()
    want = a.float().t() @ d.float()
    err = (c - want).abs().max().item()
    print("RESULT variant=%d (lbo=%d sbo=%d) r=%d n=%d max_abs_err=%.3e ref_max=%.3f" % (variant, lbo, sbo, r, n, err, want.abs().max().item()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run_one(*[int(x) for x in sys.argv[2:5]])
        sys.exit(0)
    for v in [(0, 64, 64), (1, 64, 64), (0, 128, 192), (1, 128, 192)]:
        cmd = ["timeout", "120", sys.executable, os.path.abspath(__file__), "one"] + [str(x) for x in v]
        r = subprocess.run(cmd, capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(lines[0] if lines else "FAILED %s rc=%d %s" % (v, r.returncode, (r.stderr or "")[-300:].replace("\n", " | ")))
