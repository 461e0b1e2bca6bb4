"""Blackwell mnemonic counts per kernel of the built library (no GPU needed):
    python tools/sass_mnemonics.py > profiles/sass_r02_blackwell_mnemonics.md
UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, HMMA = mma.sync,
SYNCS = mbarrier ops, ENL2.256 = 256-bit global loads / stores, BRA.U.ANY = the per-instruction election loop the
compiler wraps around tcgen05 / bulk-copy instructions issued from a divergent `if (lane == 0)` region."""
import os, re, subprocess, sys, collections

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, "glow_tts_b200", "csrc", "libglowcore.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
archs = set(re.findall(r"arch = (sm_\w+)", out))
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "HMMA", "SYNCS", "ENL2.256", "BRA.U.ANY"]
per, cur, it = collections.OrderedDict(), None, iter(names)
tot = collections.Counter()
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        per[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for k in keys:
        if re.search(r"\b" + re.escape(k), line):
            per[cur][k] += 1
            tot[k] += 1
    if "UTCATOMSWS" in line:
        tot["UTCATOMSWS"] += 1
    if "UTMALDG" in line:
        tot["UTMALDG"] += 1
print("# SASS evidence, round 2 final build (`cuobjdump -sass glow_tts_b200/csrc/libglowcore.so`, architectures: %s)\n" % ", ".join(sorted(archs)))
print("Mnemonic counts over the whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in keys + ["UTCATOMSWS", "UTMALDG"]) + ".")
print("No UTMALDG (no tensor-map TMA: the SWIZZLE_NONE slab layout wants 16-byte rows, which the TMA engine delivers at ~1 row per")
print("cycle -- profiles/ubench_r01.md -- so activations go LDG.128 -> STS.128 and only the pre-packed weight stages ride the")
print("bulk-copy engine).  BRA.U.ANY: kernels that still issue tcgen05.mma from `if (lane == 0)` (weight gradients, probes);")
print("the decoder / encoder GEMM kernel and the layer kernel issue from an elected lane of a converged warp and have none.\n")
print("| kernel | " + " | ".join(keys) + " |\n|---|" + "---:|" * len(keys))
for name, c in per.items():
    if c["UTCHMMA"] or c["HMMA"] or c["UBLKCP"]:
        print("| `%s` | " % name[:170] + " | ".join(str(c[k]) for k in keys) + " |")
