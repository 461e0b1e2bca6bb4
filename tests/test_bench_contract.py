"""bench.py's output contract, checked on the CPU through the reference arm (`--impl reference`, the only arm that
needs no GPU): stdout is exactly ONE line, it is JSON, and it carries the keys the driver reads; non-zero ranks of a
torchrun launch print nothing and exit 0."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--ref-sample", "1"], cwd=REPO, env=env, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "mel_frames_per_sec_train_step"
    assert line["unit"] == "mel-frames/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    assert line["config"]["workload"].startswith("configs[1]")


def test_reference_arm_is_silent_on_other_ranks():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""
