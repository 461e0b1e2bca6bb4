#!/usr/bin/env python
"""bench.py -- mel-frames/sec of one full Glow-TTS train step on LJSpeech-shaped synthetic
batches (BASELINE.json metric, configs[1]: Vanilla, batch 32 per GPU, <= 1000 mel frames, bf16
flow kernels), on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--workload train|mas|decoder]
                    [--mode Vanilla|SE] [--kind lj|ljvctk] [--batch B] [--geometries G]

N > 1 is launched by torchrun (one rank per GPU, NCCL); every rank runs the same code, rank 0
prints ONE JSON line.  A "step" is forward + MLE/MSE losses + backward + ONE all-reduce of the
flat gradient buffer + clip + RAdam + Noam (Train.py:182-233 of the reference) on one batch.

* The timed region CYCLES through G (default 8) distinct batch geometries, as the reference's loop does
  (Train.py:582-584 never repeats a geometry): every step trains on a batch with other lengths; the captured
  CUDA graphs are per geometry bucket, not per batch (train.GraphedTrainStep, geometry.py).
* value  : real (unpadded) mel frames of all ranks / step time, batches resident in HBM (device-to-device
           refresh of the static input buffers + the ~0.2 MB row-map blob per step).
* e2e    : same steps through the public API with the batches in pinned HOST memory: H2D of every
           batch and a D2H read of the loss are inside the timed region, every step.
* roofline: the kernel family with the largest device time per step, timed per launch with CUDA events by the
           library's own hook (glow_prof_*) during extra steps run right after the timed region;
           `roofline.families` has every GEMM family's fraction of the measured bf16 peak.
* cpu_baseline / --impl reference: the CPU restatement of the reference's train step
           (oracle/glow_oracle.py, torch-CPU fp32 + the reference's Cython MAS from oracle/_ref)
           on all host threads, on the same batches.  /root/reference is never read here.
* gpu_reference: the same restatement run as eager torch ON THE B200 (fp32, and under autocast-bf16) -- the
           practical "GPU reference" of BASELINE.md section 3 step 5 (the reference itself is eager torch + a
           host-side Cython MAS).
* configs2_se_lut: BASELINE configs[2] in the same run: SE-LUT multispeaker, LJ+VCTK-shaped GLOBAL batch 64
           sharded by utterance over the N ranks (strong scaling), same step.
* extra_workloads (N = 1): headline numbers of configs[3] (decoder-only sweep) and configs[4] (MAS).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "mel_frames_per_sec_train_step"
UNIT = "mel-frames/s"


# ----------------------------------------------------------------------------- helpers
def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) or the profiling guide's fallback."""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            flat = {}

            def walk(x):
                if isinstance(x, dict):
                    for k, v in x.items():
                        if isinstance(v, (int, float)):
                            flat.setdefault(k, float(v))
                        else:
                            walk(v)
            walk(d)
            for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained"):
                if k in flat:
                    out[k] = flat[k]
            out["source"] = "measured"
        except Exception:
            pass
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; only samples read between mark_begin() and
    mark_end() -- the timed regions -- are summarised (the process is started early: nvidia-smi takes a few hundred
    ms to produce its first line, longer than a short timed region)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin, self.t_end = None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first(self, timeout):
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, reasons, power = [], [], set(), []
            for _, line in rows:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, power
        lo = self.t_begin if self.t_begin is not None else 0.0
        hi = (self.t_end if self.t_end is not None else time.time()) + 0.06      # a sample is read up to one period late
        inside = [r for r in self.lines if lo <= r[0] <= hi]
        window = "timed regions"
        sm, mx, reasons, power = summarise(inside)
        if not sm:                                                               # region shorter than one period
            window = "whole run (no sample fell inside the timed regions)"
            sm, mx, reasons, power = summarise(self.lines)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w": round(sum(power) / len(power), 1),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def build_cpu_model(mode, precision, seed=0):
    """Random-init model of the reference architecture (there is no network for checkpoints).
    End convs are zero-init in the reference (Modules.py:773-778), which makes every coupling the
    identity at step 0; they get N(0, 0.01) here so the arithmetic is generic."""
    import torch
    from glow_tts_b200 import modules
    from glow_tts_b200.hparams import load_hparams
    hp = load_hparams(Mode=mode, Precision=precision)
    modules.set_hparams(hp)
    torch.manual_seed(seed)
    model = modules.GlowTTS()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
            end = blk.layers[2].layer_Dict["End"]
            end.weight.copy_(0.01 * torch.randn(end.weight.shape, generator=g))
    return model, hp


def workload_batch(kind, batch, seed):
    from glow_tts_b200.synth import make_batch
    return make_batch(kind=kind, batch=batch, seed=seed)


# ----------------------------------------------------------------------------- CPU arm
def cpu_train_steps(model_sd, mode, batches, steps, warmup, threads=None, device="cpu", autocast=False):
    """The oracle's train step (Train.py:182-233 restated), step i on batches[i % len], torch fp32 on `device`
    (cpu: all host threads; cuda: eager torch on the GPU, optionally under autocast-bf16), the reference's Cython
    MAS from oracle/_ref when it is there.  Returns (frames per timed step list, seconds list, cores, mas core)."""
    import torch
    from oracle import glow_oracle as G
    from oracle import mas as omas
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    hp = G.OracleHP(mode=mode)
    sd = G.state_dict_to_leaves({k: v.to(device) for k, v in model_sd.items()})
    opt = G.RAdamOracle([v for v in sd.values() if v.requires_grad])
    core = "ref" if omas.ref_core() is not None else "port"
    subs = []
    for tokens, tl, mels, ml, spk in batches:
        subs.append((tokens.to(device), tl.to(device), mels.to(device), ml.to(device), spk.to(device)))
    times, frames = [], []
    for i in range(warmup + steps):
        sub = subs[i % len(subs)]
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                G.train_step(sd, hp, opt, sub, training=True, mas_core=core)
        else:
            G.train_step(sd, hp, opt, sub, training=True, mas_core=core)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            frames.append(int(sub[3].sum()))
    return frames, times, cores, core


def take_utterances(batch, n):
    tokens, tl, mels, ml, spk = batch
    tl, ml = tl[:n], ml[:n]
    return (tokens[:n, :int(tl.max())].contiguous(), tl, mels[:n, :, :int(ml.max())].contiguous(), ml, spk[:n])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the step, all host threads, on the SAME batches
    the GPU arm cycles through (rank 0's), torch fp32."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, hp = build_cpu_model(args.mode, "fp32")
    batches = [rank_batch(args, 0, 1, g) for g in range(args.geometries)]
    take = args.ref_sample or args.batch
    # size the per-step sample so that the whole run ends within a few minutes
    probe = [take_utterances(b, take) for b in batches[:1]]
    _, t_probe, cores, core = cpu_train_steps(model.state_dict(), args.mode, probe, 1, 0)
    budget = 240.0
    while take > 1 and t_probe[0] * (args.steps + args.warmup) > budget:
        take = max(1, take // 2)
        probe = [take_utterances(b, take) for b in batches[:1]]
        _, t_probe, cores, core = cpu_train_steps(model.state_dict(), args.mode, probe, 1, 0)
    subs = [take_utterances(b, take) for b in batches]
    frames, times, cores, core = cpu_train_steps(model.state_dict(), args.mode, subs, args.steps, args.warmup)
    sec = sum(times) / len(times)
    value = sum(frames) / sum(times)
    sample = ("%s %d utterances of each config batch (%d real mel frames per step on average), %d geometries cycled; "
              "oracle/glow_oracle.py train step (torch-CPU fp32, %d threads) + %s MAS" %
              ("all" if take == args.batch else "first", take, sum(frames) // len(frames), len(subs), cores,
               "the reference's Cython (oracle/_ref)" if core == "ref" else "C port"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": own_config(args),
        # kind "port": the train step is the oracle's restatement; only its MAS core is the reference's own compiled
        # Cython (mas_core says which)
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "mas_core": "reference (oracle/_ref, core.pyx compiled)" if core == "ref" else "port (oracle/mas_oracle.c)",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "utterances_per_step": take,
    }
    emit_json(line)
    return 0


def own_config(args, world=None):
    """The workload both arms run -- identical in the own and the reference line."""
    world = world or args.gpus
    name = {"Vanilla": "configs[1]: Vanilla single-speaker, LJSpeech-shaped batch=%d per GPU, <=1000 mel frames",
            "SE": "configs[2]: SE-LUT multispeaker, LJ+VCTK-shaped batch=%d per GPU"}[args.mode] % args.batch
    return {"workload": name + ", full train step (fwd+loss+bwd+allreduce+clip+RAdam), %d batch geometries cycled" %
                        args.geometries,
            "mode": args.mode, "kind": args.kind, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "geometries": args.geometries, "parallelism": "dp%d" % world}


def rank_batch(args, rank, world, g, kind=None, mode=None, batch=None, sharded=False):
    """Batch of geometry index g for `rank`.  Weak scaling: every rank its own batch.  sharded: one GLOBAL batch
    of batch * world utterances (same seed on all ranks), contiguous shard [rank*batch, (rank+1)*batch)."""
    kind = kind or args.kind
    batch = batch or args.batch
    if not sharded:
        return workload_batch(kind, batch, 1000 * rank + g)
    tokens, tl, mels, ml, spk = workload_batch(kind, batch * world, 7000 + g)
    lo, hi = rank * batch, (rank + 1) * batch
    tl_s, ml_s = tl[lo:hi], ml[lo:hi]
    return (tokens[lo:hi, :int(tl_s.max())].contiguous(), tl_s, mels[lo:hi, :, :int(ml_s.max())].contiguous(), ml_s,
            spk[lo:hi])


# ----------------------------------------------------------------------------- GPU arm
# algorithmic FLOPs per real squeezed frame (= packed row) of one launch of each GEMM family of the coupling net
# (Modules.py:780-887): 2 * K * N; res_skip / b_rs are 192 x 384 for layers 0-2 and 192 x 192 for the last
FAMILY_FLOPS_PER_ROW = {
    # one WaveNet layer in one launch (flow_tc_layer.cuh): gate GEMM + res/skip GEMM
    "layer": 2 * 960 * 384 + 2 * 192 * (3 * 384 + 192) / 4.0,
    "start": 2 * 80 * 192, "in_gate": 2 * 960 * 384, "res_skip": 2 * 192 * (3 * 384 + 192) / 4.0, "end": 2 * 192 * 160,
    "b_end": 2 * 160 * 192, "b_rs": 2 * (3 * 384 + 192) / 4.0 * 192, "b_in": 2 * 1920 * 192, "b_start": 2 * 192 * 80,
    # one block's weight gradients (wgrad_tc.cuh), one launch per shape class: the four k=5 gradients; end + 4 skip +
    # 3 res 192-wide 1x1 gradients; the start 1x1
    "wgrad_in": 4 * 2 * 960 * 384, "wgrad_1x1": 2 * (192 * 160 + 7 * 192 * 192), "wgrad_start": 2 * 80 * 192,
}
# CTAs a launch of the family occupies (one CTA per SM: ~200 KB of shared memory each): (column slices per row tile, or a
# fixed CTA count for the weight-gradient batches, which walk the whole row axis in 2 - 24 CTAs)
FAMILY_SLICES = {"layer": 1, "start": 1, "in_gate": 3, "res_skip": 2.5, "end": 1, "b_end": 1, "b_rs": 1, "b_in": 1, "b_start": 1}
FAMILY_CTAS = {"wgrad_in": 24, "wgrad_1x1": 16, "wgrad_start": 2}


class TrainBench:
    """One model + TrainStep + GraphedTrainStep on this rank and the timing loops over G cycled batches."""

    def __init__(self, args, mode, kind, batch, geometries, sharded, world, rank, dev):
        import torch
        import torch.distributed as dist
        from glow_tts_b200.train import TrainStep, GraphedTrainStep
        self.torch, self.dist, self.world, self.rank, self.dev, self.args = torch, dist, world, rank, dev, args
        self.mode = mode
        model, hp = build_cpu_model(mode, args.precision)
        self.cpu_sd = {k: v.detach().clone() for k, v in model.state_dict().items()} if rank == 0 else None
        self.model = model.to(dev)
        self.model.train()
        self.step = TrainStep(self.model, hp, dev)
        self.host = [rank_batch(args, rank, world, g, kind, mode, batch, sharded) for g in range(geometries)]
        self.pinned = [(t.pin_memory(), tl, m.pin_memory(), ml, s.pin_memory()) for t, tl, m, ml, s in self.host]
        self.resident = [(t.to(dev), tl, m.to(dev), ml, s.to(dev)) for t, tl, m, ml, s in self.pinned]
        # global (all ranks) real frames / B * T_x,max per geometry: what the data-parallel loss weights need
        counts = torch.tensor([[float(b[3].sum()), float(b[0].shape[0]), float(b[2].shape[0] * b[2].shape[2])]
                               for b in self.host], device=dev)
        tmax = torch.tensor([float(b[0].shape[1]) for b in self.host], device=dev)
        if world > 1:
            dist.all_reduce(counts)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        self.g_frames = [int(v) for v in counts[:, 0].tolist()]
        self.g_padded = [int(v) for v in counts[:, 2].tolist()]
        self.g_pos = [int(c) * int(t) for c, t in zip(counts[:, 1].tolist(), tmax.tolist())]
        self.graphed = None if args.no_graph else GraphedTrainStep(self.step)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def one(self, i, source):
        """Step i of the cycle on batches from `source` (self.resident: device tensors, self.pinned: host)."""
        g = i % len(source)
        if self.graphed is not None:
            return self.graphed.run(source[g], self.g_frames[g], self.g_pos[g])
        b = source[g]
        if not b[0].is_cuda:
            b = self.step.to_device(b)
        return self.step.run(b, global_frames=self.g_frames[g], global_positions=self.g_pos[g])

    def warm(self, n):
        """Every geometry once (first batch of a bucket: eager step + capture), then n more steps."""
        for i in range(len(self.resident)):
            self.one(i, self.resident)
        self.barrier()
        for i in range(n):
            self.one(i, self.resident)
        self.barrier()

    def timed(self, steps, source, read_loss=False, prefetch=False):
        torch = self.torch
        from glow_tts_b200 import _lib
        n0 = _lib.launch_count()
        replays0 = {k: b.replays for k, b in self.graphed.buckets.items()} if self.graphed is not None else {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        loss_host = None
        # D2H read of every step's loss through pinned host memory, one step behind: step i's loss is copied out
        # (stream-ordered) right after step i is issued and read on the host once step i + 1 has been issued, so the
        # host never idles the GPU (what a training loop that logs its loss does); all reads are inside the timed region
        host_loss = [torch.zeros(1).pin_memory() for _ in range(2)] if read_loss else None
        host_ev = [torch.cuda.Event(), torch.cuda.Event()] if read_loss else None
        if prefetch and self.graphed is not None:
            self.graphed.prefetch(source[0])
        for i in range(steps):
            loss = self.one(i, source)
            if prefetch and self.graphed is not None and i + 1 < steps:
                self.graphed.prefetch(source[(i + 1) % len(source)])
            if read_loss:
                host_loss[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
                host_ev[i & 1].record()
                if i > 0:
                    host_ev[(i - 1) & 1].synchronize()
                    loss_host = float(host_loss[(i - 1) & 1])
        if read_loss:
            host_ev[(steps - 1) & 1].synchronize()
            loss_host = float(host_loss[(steps - 1) & 1])             # the last step's loss, still before e1
        e1.record()
        self.barrier()
        launches = _lib.launch_count() - n0
        if self.graphed is not None:                        # replays launch the captured kernels without the host
            for k, b in self.graphed.buckets.items():
                launches += b.launches * (b.replays - replays0.get(k, 0))
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        frames = sum(self.g_frames[i % len(source)] for i in range(steps))
        padded = sum(self.g_padded[i % len(source)] for i in range(steps))
        return float(ms), frames, padded, launches, loss_host


def run_own(args):
    import torch
    import torch.distributed as dist
    from glow_tts_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a collective that cannot complete must fail in minutes, not after NCCL's 10-minute default
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    _lib.lib()

    tb = TrainBench(args, args.mode, args.kind, args.batch, args.geometries, False, world, rank, dev)
    step = tb.step

    if args.trace or args.profile_mode:
        return run_trace(args, tb)

    # warm-up: every geometry once (also runs the one-time ActNorm data-dependent init), then W >= 3 more steps
    sampler = ClockSampler(local).start() if rank == 0 else None
    tb.warm(max(args.warmup, 3))
    if sampler is not None:
        sampler.wait_first(3.0)
        sampler.mark_begin()
    ms_per_step, frames, padded, launches, _ = tb.timed(args.steps, tb.resident)
    # end to end through the public API: pinned host batch -> device -> step -> loss back on the host.  With the
    # graph, the NEXT step's batch starts its H2D copy (copy stream, staging buffers) right after this step's replay
    # is launched, so the copy overlaps the step in flight; every step's copy is inside the timed region.
    e2e_ms, e2e_frames, _, _, loss_host = tb.timed(args.steps, tb.pinned, read_loss=True, prefetch=True)
    if sampler is not None:
        sampler.mark_end()
    clocks = sampler.stop() if sampler else None
    h2d = sum(t.numel() * t.element_size() for b in tb.pinned for t in (b[0], b[2], b[4])) / float(len(tb.pinned))
    geo_bytes = 0
    if tb.graphed is not None and tb.graphed.current is not None:
        geo_bytes = tb.graphed.current.geo.blob.numel() * 4
    buckets = len(tb.graphed.buckets) if tb.graphed is not None else 0
    launches_per_replay = (sum(b.launches for b in tb.graphed.buckets.values()) / float(buckets)) if buckets else None

    # per-kernel-family device time (library hook), two eager steps of the same workload
    _lib.prof_enable(True)
    for i in range(2):
        b = tb.resident[i % len(tb.resident)]
        step.run(b, global_frames=tb.g_frames[i % len(tb.resident)], global_positions=tb.g_pos[i % len(tb.resident)])
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_report()
    prof_rows = sum(int(n) // 2 for i in range(2) for n in tb.host[i % len(tb.host)][3].tolist()) / 2.0
    tb.barrier()

    # BASELINE configs[2] in the same run: SE-LUT, LJ+VCTK-shaped GLOBAL batch 64 sharded over the ranks
    se = None
    if args.mode == "Vanilla" and not args.no_config3 and 64 % world == 0:
        se = run_config3(args, world, rank, dev)

    if rank != 0:
        _finish(world, tb)
        return 0

    pk = peaks()
    kernels = {k: {"launches_per_step": v[0] / 2.0, "ms_per_step": v[1] / 2.0} for k, v in sorted(prof.items())}
    fams = {}
    for fam, fpr in FAMILY_FLOPS_PER_ROW.items():
        if fam in prof and prof[fam][0] > 0:
            n, tot = prof[fam]
            avg_ms = tot / n
            flops = fpr * prof_rows
            ach = flops / (avg_ms * 1e-3) / 1e12
            tiles = math.ceil((prof_rows + 2 * args.batch + 2) / 128.0)
            ctas = FAMILY_CTAS.get(fam) or min(148, int(tiles * FAMILY_SLICES.get(fam, 1)))
            fams[fam] = {"launches_timed": n, "avg_launch_ms": avg_ms, "ms_per_step": tot / 2.0, "flops_per_launch": flops,
                         "achieved": ach, "frac": ach / pk["bf16_tflops_sustained"], "ctas": ctas,
                         # SMs x time the family holds per step, and its rate on the SMs it actually occupies
                         "sm_ms_per_step": tot / 2.0 * ctas / 148.0,
                         "frac_of_sms_used": ach / (pk["bf16_tflops_sustained"] * ctas / 148.0)}
    roof = None
    if fams:
        # the dominant family by SM-time: launch duration x SMs occupied (the weight-gradient batches run for ~150 us
        # on 2 - 24 SMs in the background of the data-gradient chain; by wall time alone they would always "dominate")
        fam = max(fams, key=lambda k: fams[k]["sm_ms_per_step"])
        traffic = None
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(fam)
        roof = {"bound": "tensor", "kernel": fam, "achieved": fams[fam]["achieved"], "peak": pk["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": fams[fam]["frac"], "traffic": traffic,
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS)", "launches_timed": fams[fam]["launches_timed"],
                "avg_launch_ms": fams[fam]["avg_launch_ms"], "flops_per_launch": fams[fam]["flops_per_launch"],
                "frac_of_sms_used": fams[fam]["frac_of_sms_used"], "ctas": fams[fam]["ctas"],
                "selection": "family with the largest SM-time per step (launch duration x SMs occupied); `families` lists all, "
                             "`ms_per_step` is plain device time", "families": fams}
    dec_ms = sum(v["ms_per_step"] for k, v in kernels.items() if not k.startswith(("rpr_", "mas", "enc_")))
    # the north_star's "fraction of the decoder's HBM roofline": algorithmic 4800 B x s per mel frame
    # (s = 2 B bf16 / 4 B fp32) over the decoder kernel time of one step
    s_bytes = 2 if args.precision == "bf16" else 4
    hbm = None
    if dec_ms > 0:
        gbs = 2.0 * prof_rows * 4800.0 * s_bytes / (dec_ms * 1e-3) / 1e9
        hbm = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
               "note": "decoder kernels of one train step; compute is the binding roof (SURVEY 8d)"}
    # whole-step tensor-roofline view: ~71 MFLOP per PADDED-free (real) mel frame over the step time
    step_tflops = (frames / float(args.steps)) * 71e6 / (ms_per_step * 1e-3) / 1e12 / world

    cpu = gpu_ref = None
    if world == 1 and not args.no_cpu:
        n_utt = args.cpu_sample or args.batch
        subs = [take_utterances(b, n_utt) for b in tb.host[:2]]
        fr, times, cores, core = cpu_train_steps(tb.cpu_sd, args.mode, subs, 2, 1)
        cpu = {"value": sum(fr) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "mas_core": "reference (oracle/_ref, core.pyx compiled)" if core == "ref" else "port (oracle/mas_oracle.c)",
               "sample": "%d of the %d utterances of two config batches (%d real mel frames per step), 1 warm-up + 2 timed "
                         "steps, oracle/glow_oracle.py (torch-CPU fp32, %d threads) + %s MAS" %
                         (n_utt, args.batch, sum(fr) // len(fr), cores,
                          "the reference's Cython (oracle/_ref)" if core == "ref" else "C port")}
    if world == 1 and not args.no_gpu_ref:
        gpu_ref = {}
        for name, ac in (("eager_torch_fp32", False), ("eager_torch_autocast_bf16", True)):
            try:
                fr, times, _, core = cpu_train_steps(tb.cpu_sd, args.mode, tb.host[:4], 4, 2, device="cuda", autocast=ac)
                gpu_ref[name] = {"value": sum(fr) / sum(times), "unit": UNIT, "ms_per_step": 1e3 * sum(times) / len(times)}
            except Exception as exc:                         # e.g. an op without a bf16 kernel under autocast
                gpu_ref[name] = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        gpu_ref["what"] = ("oracle/glow_oracle.py train step (the reference's forward / loss / backward / clip / RAdam "
                           "restated functionally) as eager torch on this B200, 4 of the cycled batches, 2 warm-up + 4 timed "
                           "steps, host-timed with synchronize; MAS on the host through %s as the reference does "
                           "(monotonic_align/__init__.py:14-21)" % ("oracle/_ref" if core == "ref" else "the C port"))

    extra = None
    if world == 1 and not args.no_extra:
        extra = run_extra(args)

    line = {
        "metric": METRIC, "value": frames / (ms_per_step * args.steps * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": own_config(args, world),
        "detail": {"real_mel_frames_per_step": frames / float(args.steps), "padded_mel_frames_per_step": padded / float(args.steps),
                   "launch": ("cuda-graph replay, one captured graph per geometry bucket (%d buckets for the %d cycled batches)"
                              % (buckets, len(tb.resident))) if tb.graphed is not None else "eager (one host launch per kernel)",
                   "l2": "no explicit flush: one step streams > 2 GB of saved activations and 0.46 GB of parameter / "
                         "optimizer state, >> 126 MB L2, and consecutive steps train on different batches",
                   "warmup_steps_run": max(args.warmup, 3) + len(tb.resident),
                   "warmup_note": "the W requested (at least 3) after one untimed step per cycled batch geometry (the first "
                                  "batch of a bucket runs eagerly and its graph is captured)",
                   "precision": args.precision, "geometry_blob_bytes_per_step": geo_bytes,
                   "step_tflops_per_gpu": step_tflops, "step_frac_of_bf16_peak": step_tflops / pk["bf16_tflops_sustained"]},
        "padded_frames_per_sec": padded / (ms_per_step * args.steps * 1e-3),
        "e2e": {"value": e2e_frames / (e2e_ms * args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d + geo_bytes), "d2h_bytes_per_step": 4, "loss": loss_host,
                "d2h": "every step's loss is copied to pinned host memory and read on the host one step later "
                       "(after the next step has been issued); all reads inside the timed region",
                "h2d": ("double-buffered: step i+1's batch is copied from pinned host memory while step i runs "
                        "(GraphedTrainStep.prefetch), every copy inside the timed region") if tb.graphed is not None
                       else "copied in front of every step"},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / float(args.steps),
        "gpu_launches_per_replay": launches_per_replay,
        "clocks": clocks, "roofline": roof, "roofline_hbm_decoder": hbm, "kernels": kernels,
        "cpu_baseline": cpu, "gpu_reference": gpu_ref, "configs2_se_lut": se, "extra_workloads": extra,
    }
    emit_json(line)
    _finish(world, tb)
    return 0


def run_config3(args, world, rank, dev):
    """BASELINE configs[2]: SE-LUT multispeaker, LJ+VCTK-shaped GLOBAL batch 64 sharded contiguously by utterance
    over the ranks, the same full train step; 4 geometries cycled."""
    import copy
    a = copy.copy(args)
    a.mode, a.kind, a.batch, a.geometries = "SE", "ljvctk", 64 // world, 4
    tb = TrainBench(a, "SE", "ljvctk", 64 // world, 4, True, world, rank, dev)
    tb.warm(3)
    steps = max(8, min(args.steps, 20))
    ms, frames, padded, launches, _ = tb.timed(steps, tb.resident)
    e2e_ms, e2e_frames, _, _, loss = tb.timed(steps, tb.pinned, read_loss=True, prefetch=True)
    out = {"metric": METRIC, "value": frames / (ms * steps * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "scaling": "strong", "config": own_config(a, world),
           "e2e": {"value": e2e_frames / (e2e_ms * steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "loss": loss},
           "real_mel_frames_per_step": frames / float(steps), "gpu_launches_per_step": launches / float(steps)}
    if tb.graphed is not None:
        del tb.graphed
    del tb
    return out


def run_extra(args):
    """Headline numbers of the secondary workloads (configs[3] decoder-only sweep, configs[4] MAS), condensed."""
    import copy
    from tools import bench_extra
    me = sys.modules[__name__]
    a = copy.copy(args)
    a.steps = max(10, min(args.steps, 20))
    out = {}
    try:
        m = bench_extra.mas_line(a, me)
        out["configs4_mas"] = {"value": m["value"], "unit": m["unit"], "ms_per_step": m["ms_per_step"], "roofline": m["roofline"],
                               "e2e": m["e2e"], "cpu_baseline": m["cpu_baseline"], "config": m["config"],
                               "grid": [{k: g[k] for k in ("batch", "t_text", "t_mel", "aligns_per_s")} for g in m["grid"]]}
    except Exception as exc:
        out["configs4_mas"] = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    try:
        d = bench_extra.decoder_line(a, me)
        out["configs3_decoder_sweep"] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                                         "roofline": d["roofline"], "roofline_hbm_decoder": d["roofline_hbm_decoder"],
                                         "config": d["config"], "sweep": d["sweep"]}
    except Exception as exc:
        out["configs3_decoder_sweep"] = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    return out


def run_trace(args, tb):
    """--trace: torch.profiler kernel table of the eager step; --profile-mode: bare steps for a run under ncu."""
    import torch
    step = tb.step

    def one_step(i):
        g = i % len(tb.resident)
        return step.run(tb.resident[g], global_frames=tb.g_frames[g], global_positions=tb.g_pos[g])

    if args.profile_mode:
        for i in range(args.warmup):
            one_step(i)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("timed")
        for i in range(args.steps):
            one_step(i)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return 0
    from torch.profiler import profile, ProfilerActivity
    for i in range(max(args.warmup, 3)):
        one_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        one_step(i)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            one_step(i)
        torch.cuda.synchronize()
    ev = prof.key_averages()
    from torch.autograd import DeviceType
    rows = [(e.key, e.count, e.self_device_time_total / 1e3) for e in ev
            if e.self_device_time_total > 0 and e.device_type == DeviceType.CUDA]
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print("# wall %.3f ms/step (unprofiled); device-busy %.3f ms/step over %d steps" %
          (wall * 1e3, tot / args.steps, args.steps))
    print("| kernel | launches/step | ms/step | share |\n|---|---:|---:|---:|")
    for k, n, ms in rows[:60]:
        print("| `%s` | %.1f | %.3f | %.1f%% |" % (k[:90], n / args.steps, ms / args.steps, 100 * ms / tot))
    print("| total | %.1f | %.3f | 100%% |" % (sum(r[1] for r in rows) / args.steps, tot / args.steps))
    return 0


def claim_stdout():
    """stdout carries exactly ONE line, the JSON: keep the real stdout for it and point fd 1 at stderr for
    everything else that may print there (NCCL's version banner, library warnings)."""
    if getattr(sys, "_glow_json_fd", None) is None:       # on `sys`: tools/bench_extra imports this file a second time
        sys.stdout.flush()
        sys._glow_json_fd = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    fd = getattr(sys, "_glow_json_fd", None)
    os.write(fd if fd is not None else 1, data)


def _finish(world, tb):
    """Orderly teardown: the captured graphs hold the NCCL all-reduce nodes, so they go first, then the process
    group.  destroy_process_group() with live graphs that captured a collective has been seen to hang for minutes
    (round 1); a watchdog ends the process if the teardown does not return."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()

    def _bail():
        time.sleep(20.0)
        os._exit(0)
    threading.Thread(target=_bail, daemon=True).start()
    try:
        if tb is not None and getattr(tb, "graphed", None) is not None:
            tb.graphed.buckets.clear()
            tb.graphed.current = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "mas", "decoder", "inference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32-tc"])
    ap.add_argument("--mode", default="Vanilla", choices=["Vanilla", "SE"], help="Vanilla: configs[1]; SE: configs[2] (SE-LUT)")
    ap.add_argument("--kind", default=None, choices=["lj", "ljvctk"], help="batch shapes (default: lj for Vanilla, ljvctk for SE)")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (default 32; 8 for --mode SE)")
    ap.add_argument("--geometries", type=int, default=8, help="distinct batch geometries cycled in the timed region")
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances per batch in the cpu_baseline sample (0: all)")
    ap.add_argument("--ref-sample", type=int, default=0, help="utterances per step of --impl reference (0: all)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the eager-torch-on-GPU reference point")
    ap.add_argument("--no-config3", action="store_true", help="skip the configs[2] (SE-LUT, global batch 64) leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] headline numbers")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--trace", action="store_true", help="torch.profiler kernel table of the step instead of the bench line")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: exactly --warmup warm-up steps, then --steps steps, nothing else")
    args = ap.parse_args()
    if args.kind is None:
        args.kind = "lj" if args.mode == "Vanilla" else "ljvctk"
    if args.batch is None:
        args.batch = 32 if args.mode == "Vanilla" else 8
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "train":
        from tools import bench_extra
        return bench_extra.run(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
