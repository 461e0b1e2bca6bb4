"""Run a few train steps of the bench workload (configs[1]) and nothing else: the short command
to put under `ncu` (launch list or --set full on one kernel).   python tools/one_step.py [steps]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from glow_tts_b200.train import TrainStep
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model, hp = bench.build_cpu_model("Vanilla", "bf16")
for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
    blk.layers[0].initialized = True
model = model.cuda().train()
step = TrainStep(model, hp, torch.device("cuda:0"))
b = step.to_device(bench.workload_batch("lj", 32, 0))
for _ in range(steps):
    step.run(b)
torch.cuda.synchronize()
print("ok", float(step.last["loss"]))
