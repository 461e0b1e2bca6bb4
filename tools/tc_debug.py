"""Timeline of the first launches of each tensor-core GEMM (GLOW_TC_DEBUG=1): cycles since kernel
start at which CTA 0's MMA warp saw its A panel and each weight stage, and its epilogue ran."""
import os, sys
os.environ["GLOW_TC_DEBUG"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import bench
from glow_tts_b200.train import TrainStep
model, hp = bench.build_cpu_model("Vanilla", "bf16")
for blk in model.layer_Dict["Decoder"].layer_Dict["Flows"]:
    blk.layers[0].initialized = True
model = model.cuda().train()
step = TrainStep(model, hp, torch.device("cuda:0"))
b = step.to_device(bench.workload_batch("lj", 32, 0))
for _ in range(int(os.environ.get("GLOW_TC_DEBUG_STEPS", "1"))):
    step.run(b)
torch.cuda.synchronize()
