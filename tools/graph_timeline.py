"""Kernel timeline of ONE replay of the captured train step (torch.profiler / CUPTI): span, per-stream
busy time, idle gaps on the main stream, and kernel families by summed duration.

    python tools/graph_timeline.py > profiles/timeline_rNN.md
"""
import os, sys, json, tempfile, collections
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from glow_tts_b200.train import TrainStep, GraphedTrainStep
from torch.profiler import profile, ProfilerActivity

def _arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


# --mode SE --kind ljvctk --batch 64: BASELINE configs[2] on one GPU
MODE, KIND, BATCH = _arg("--mode", "Vanilla"), _arg("--kind", "lj"), int(_arg("--batch", "32"))
model, hp = bench.build_cpu_model(MODE, _arg("--precision", "bf16"))
dev = torch.device("cuda:0")
model = model.to(dev).train()
step = TrainStep(model, hp, dev)
host = bench.workload_batch(KIND, BATCH, 0)
for _ in range(3):
    step.run(step.to_device(host))
g = GraphedTrainStep(step, host, warmup=1)
for _ in range(3):
    g.run()
torch.cuda.synchronize()
n_rep = 2 if "--two" in sys.argv else 1
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n_rep):
        g.run()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print("# %d replay(s): %d device activities, span %.3f ms" % (n_rep, len(ev), (t1 - t0) / 1e3))
if n_rep == 2:
    radam = [e for e in ev if "radam_kernel" in e["name"]]
    if len(radam) == 2:
        first_end = radam[0]["ts"] + radam[0]["dur"]
        nxt = min(e["ts"] for e in ev if e["ts"] >= first_end)
        print("replay period (radam to radam): %.3f ms; idle between the replays: %.3f ms" %
              ((radam[1]["ts"] - radam[0]["ts"]) / 1e3, (nxt - first_end) / 1e3))
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream", -1)].append(e)
print("\n| stream | activities | busy ms | first start ms | last end ms |\n|---|---:|---:|---:|---:|")
for sid, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print("| %s | %d | %.3f | %.3f | %.3f |" % (sid, len(es), sum(e["dur"] for e in es) / 1e3, (es[0]["ts"] - t0) / 1e3,
                                                (max(e["ts"] + e["dur"] for e in es) - t0) / 1e3))
# union busy over all streams -> idle time of the whole GPU
cur_end, idle = t0, 0.0
for e in ev:
    if e["ts"] > cur_end:
        idle += e["ts"] - cur_end
    cur_end = max(cur_end, e["ts"] + e["dur"])
print("\nGPU idle (no kernel on any stream): %.3f ms" % (idle / 1e3))
fam = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"]
    for key in ("tc_gemm3_kernel", "wgrad_tc_kernel", "nvjet", "colsum", "rpr_attn", "elementwise", "cutlass", "layer_norm", "LayerNorm",
                "GammaBeta", "index", "gather", "reduce_kernel", "wn_pack", "wn_grad", "rows_pack", "mas_kernel",
                "radam", "mix_bwd", "dropout", "Memset", "Memcpy"):
        if key in n:
            n = key
            break
    fam[n[:60]][0] += 1
    fam[n[:60]][1] += e["dur"]
print("\n| family | launches | sum ms |\n|---|---:|---:|")
for k, (n, d) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:28]:
    print("| `%s` | %d | %.3f |" % (k, n, d / 1e3))
# coarse phases on the busiest stream: time of first / last tc_gemm3 and of the attention kernels
main = max(streams.values(), key=lambda es: sum(e["dur"] for e in es))
marks = [(e["ts"] - t0, e["name"][:50]) for e in main if "mas_kernel" in e["name"] or "radam" in e["name"] or "wn_pack" in e["name"] or "wn_grad" in e["name"] or "sqnorm" in e["name"]]
print("\nmarkers on the main stream (ms): " + "; ".join("%s @ %.3f" % (n, t / 1e3) for t, n in marks))

if "--list" in sys.argv:
    n = int(sys.argv[sys.argv.index("--list") + 1])
    print("\nfirst %d activities (all streams): start us | dur us | gap to previous end us | stream | name" % n)
    prev_end = t0
    start_ms = float(sys.argv[sys.argv.index("--from-ms") + 1]) if "--from-ms" in sys.argv else 0.0
    sel = [e for e in ev if (e["ts"] - t0) / 1e3 >= start_ms]
    for e in sel[:n]:
        print("%9.1f %7.1f %7.1f  %s  %s" % (e["ts"] - t0, e["dur"], e["ts"] - prev_end, e["args"].get("stream", -1), e["name"][:70]))
        prev_end = max(prev_end, e["ts"] + e["dur"])
