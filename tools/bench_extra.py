"""Secondary workloads of bench.py (same JSON contract, one line per run):

  --workload mas      BASELINE.json configs[4]: monotonic_align.maximum_path, batch 1..256,
                      (T_text, T_mel) up to (200, 1200): GPU kernel vs the reference Cython core
                      (oracle/_ref) on the host.  Metric: alignments/s.  HBM-bound at large batch
                      (8 B/cell algorithmic: fp32 value read + fp32 path written).
  --workload decoder  configs[3]: inference, decoder only (z -> mel, Decoder(reverse=True)),
                      mel-length sweep 128..2048, B=16.  Metric: mel-frames/s.

Timing: CUDA events on the launch stream, >= 3 warm-up iterations, inputs rotated through a
set of buffers larger than L2 (126 MB) so no iteration finds its input cached.
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run_mas(args, bench):
    bench.emit_json(mas_line(args, bench))
    return 0


def mas_line(args, bench):
    import numpy as np
    import torch
    from glow_tts_b200 import _lib
    from glow_tts_b200.monotonic_align import maximum_path
    from oracle import mas as omas
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    pk = bench.peaks()
    grid = []
    shapes = [(50, 300), (100, 600), (200, 1200)]
    g = torch.Generator(device="cpu").manual_seed(1)
    for tx, ty in shapes:
        for b in (1, 4, 16, 64, 256):
            cells = b * tx * ty
            nbuf = max(2, int(160e6 // (cells * 4)) + 1)          # rotate > L2 worth of inputs
            nbuf = min(nbuf, 64)
            vals = [(-113.0 + 6.0 * torch.randn(b, tx, ty, generator=g)).to(dev) for _ in range(nbuf)]
            t_x = torch.full((b,), tx, dtype=torch.int32, device=dev)
            t_y = torch.full((b,), ty, dtype=torch.int32, device=dev)
            for i in range(3):
                maximum_path(vals[i % nbuf], None, t_x, t_y)
            torch.cuda.synchronize()
            iters = max(args.steps, 10)
            e0, e1 = _events(torch)
            e0.record()
            for i in range(iters):
                path = maximum_path(vals[i % nbuf], None, t_x, t_y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            grid.append({"batch": b, "t_text": tx, "t_mel": ty, "ms": ms, "aligns_per_s": b / (ms * 1e-3),
                         "gbs": cells * 8 / (ms * 1e-3) / 1e9})
    # headline: largest shape, batch 256, with per-launch timing from the library hook
    b, tx, ty = 256, 200, 1200
    vals = [(-113.0 + 6.0 * torch.randn(b, tx, ty, generator=g)).to(dev) for _ in range(2)]
    t_x = torch.full((b,), tx, dtype=torch.int32, device=dev)
    t_y = torch.full((b,), ty, dtype=torch.int32, device=dev)
    for i in range(max(args.warmup, 3)):
        maximum_path(vals[i % 2], None, t_x, t_y)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(dev.index or 0).start()
    n0 = _lib.launch_count()
    _lib.prof_enable(True)
    e0, e1 = _events(torch)
    e0.record()
    for i in range(args.steps):
        path = maximum_path(vals[i % 2], None, t_x, t_y)
    e1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_report()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: host values in (pinned), int32 path out -- the contract of maximum_path_c (core.pyx:40)
    hv = vals[0].cpu().pin_memory()
    hp = torch.empty((b, tx, ty), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(3, args.steps // 4)
    for _ in range(reps):
        dv = hv.to(dev, non_blocking=True)
        p = maximum_path(dv, None, t_x, t_y)
        hp.copy_(p, non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    clocks = sampler.stop()
    n, tot = prof.get("mas", (0, 0.0))
    kern_ms = tot / max(n, 1)
    cells = b * tx * ty
    achieved = cells * 8 / (kern_ms * 1e-3) / 1e9
    # CPU: reference Cython core on a bounded sample (serial as built: no -fopenmp in its setup.py)
    core = omas.ref_core()
    sb = 32
    v_np = vals[0][:sb].cpu().numpy().astype(np.float32)
    paths = np.zeros(v_np.shape, np.int32)
    tx_np = np.full(sb, tx, np.int32); ty_np = np.full(sb, ty, np.int32)
    times = []
    for _ in range(4):
        vv = v_np.copy()
        paths[:] = 0
        t0 = time.perf_counter()
        if core is not None:
            core.maximum_path_c(paths, vv, tx_np, ty_np)
        else:
            omas.maximum_path_c_port(paths, vv, tx_np, ty_np)
        times.append(time.perf_counter() - t0)
    cpu_s = min(times[1:])
    same = bool(np.array_equal(paths, maximum_path(vals[0][:sb], None, t_x[:sb], t_y[:sb], out_dtype=torch.int32).cpu().numpy()))
    line = {
        "metric": "mas_alignments_per_sec", "value": b / (ms * 1e-3), "unit": "alignments/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: MAS maximum_path, batch 256, (T_text,T_mel)=(200,1200), value ~ N(-113,6^2)",
                   "l2": "two 245 MB input batches alternate (> 126 MB L2)"},
        "e2e": {"value": b / (e2e_ms * 1e-3), "unit": "alignments/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": cells * 4, "d2h_bytes_per_step": cells * 4},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "mas_kernel", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                     "avg_launch_ms": kern_ms, "bytes_per_launch": cells * 8},
        "cpu_baseline": {"value": sb / cpu_s, "unit": "alignments/s", "cores": 1,
                         "kind": "reference" if core is not None else "port",
                         "sample": "%d alignments of the same shape, maximum_path_c core only (serial as built), best of 3" % sb,
                         "paths_equal_gpu": same},
        "grid": grid,
    }
    return line


def run_decoder(args, bench):
    bench.emit_json(decoder_line(args, bench))
    return 0


def decoder_line(args, bench):
    import torch
    from glow_tts_b200 import _lib
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    pk = bench.peaks()
    model, hp = bench.build_cpu_model("Vanilla", args.precision)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = True
    model = model.to(dev).eval()
    dec = model.layer_Dict["Decoder"]
    b = 16
    sweep = []
    g = torch.Generator().manual_seed(3)
    for t in (128, 256, 512, 1024, 2048):
        nbuf = 4
        zs = [torch.randn(b, 80, t, generator=g).to(dev) for _ in range(nbuf)]
        mask = torch.ones(b, 1, t, device=dev)
        dec.host_lengths = [t] * b
        res = {}
        for direction in ("reverse", "forward"):
            rev = direction == "reverse"
            with torch.no_grad():
                for i in range(3):
                    dec(zs[i % nbuf], mask, None, reverse=rev)
                torch.cuda.synchronize()
                iters = max(args.steps, 10)
                e0, e1 = _events(torch)
                e0.record()
                for i in range(iters):
                    dec(zs[i % nbuf], mask, None, reverse=rev)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            res[direction] = {"ms": ms, "mel_frames_per_s": b * t / (ms * 1e-3)}
        dec.host_lengths = None
        sweep.append({"t_mel": t, "batch": b, **res})
    top = sweep[-1]
    ms = top["reverse"]["ms"]
    frames = b * 2048
    s_bytes = 2 if args.precision == "bf16" else 4
    gbs = frames * 1920.0 * s_bytes / (ms * 1e-3) / 1e9
    tfl = frames * 21.35e6 / (ms * 1e-3) / 1e12
    line = {
        "metric": "decoder_inference_mel_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "mel-frames/s",
        "n_gpus": 1, "steps": max(args.steps, 10), "warmup": 3, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "configs[3]: decoder-only z->mel (reverse), B=16, T_mel=2048 (sweep 128..2048 in `sweep`)",
                   "l2": "4 rotating inputs; per call 12 blocks of activations >> L2 at T>=1024"},
        "roofline": {"bound": "tensor", "achieved": tfl, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": tfl / pk["bf16_tflops"], "traffic": None, "peak_source": pk["source"],
                     "note": "whole decoder call (12 blocks, ~122 launches), 21.35 MFLOP per mel frame"},
        "roofline_hbm_decoder": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": gbs / pk["hbm_gbs"], "note": "algorithmic 1920*s B per mel frame (SURVEY 8d)"},
        "sweep": sweep,
    }
    return line


def run_inference(args, bench):
    """GlowTTS.inference end to end (text -> mel): the CUDA-graph serving path (infer.GraphedInference) against
    the eager call with its host round trips.  Request = B sentences of ~100 tokens from pinned host memory;
    the timed region has the H2D copy of the tokens, the replay and the D2H read of the mel lengths."""
    import torch
    from glow_tts_b200.infer import GraphedInference
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    model, hp = bench.build_cpu_model("Vanilla", args.precision)
    for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
        blk.layers[0].initialized = True
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(5)
    rows = []
    for b in (1, 16):
        t_text, t_mel = 128, 1000
        lens = torch.randint(60, 121, (b,), generator=g)
        tokens = torch.randint(2, 35, (b, t_text), generator=g)
        tokens[:, 0] = 0
        for i in range(b):
            tokens[i, int(lens[i]) - 1:] = 1
        tokens, lens = tokens.pin_memory(), lens.to(torch.int32).pin_memory()
        # the duration predictor of a random-init model predicts ~1 frame per token; stretch to a speech-like 6
        gi = GraphedInference(model, b, t_text, t_mel, noise_scale=0.667, length_scale=6.0, warmup=3)
        iters = max(args.steps, 10)
        for _ in range(3):
            gi.run(tokens, lens)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            mels, ml = gi.run(tokens, lens)
            frames = int(ml.sum().item())                       # the D2H read a server needs to cut the mels
        ms_graph = (time.perf_counter() - t0) / iters * 1e3
        for _ in range(3):
            model.inference(tokens=tokens.to(dev), token_lengths=lens, noise_scale=0.667, length_scale=6.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            m2, ml2, _ = model.inference(tokens=tokens.to(dev, non_blocking=True), token_lengths=lens, noise_scale=0.667,
                                         length_scale=6.0)
            frames2 = int(ml2.sum().item())
        ms_eager = (time.perf_counter() - t0) / iters * 1e3
        rows.append({"batch": b, "t_text_max": t_text, "t_mel_max": t_mel, "mel_frames": frames,
                     "ms_graph": ms_graph, "ms_eager": ms_eager, "launches_per_replay": gi.launches_per_replay,
                     "mel_frames_per_s_graph": frames / (ms_graph * 1e-3), "mel_frames_per_s_eager": frames2 / (ms_eager * 1e-3)})
    top = rows[-1]
    line = {
        "metric": "inference_mel_frames_per_sec", "value": top["mel_frames_per_s_graph"], "unit": "mel-frames/s",
        "n_gpus": 1, "steps": max(args.steps, 10), "warmup": 3, "ms_per_step": top["ms_graph"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "GlowTTS.inference text->mel, B=16 sentences of 60-120 tokens, static T_mel=1000, "
                               "CUDA-graph replay (infer.GraphedInference); host-timed incl. H2D tokens + D2H lengths",
                   "l2": "not flushed: a serving request is latency-bound, ~60 MB of activations per call"},
        "latency_ms_batch1": rows[0]["ms_graph"], "rows": rows,
    }
    bench.emit_json(line)
    return 0


def run(args):
    import bench
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    if args.workload == "mas":
        return run_mas(args, bench)
    if args.workload == "inference":
        return run_inference(args, bench)
    return run_decoder(args, bench)
