"""CPU checks of the bench.py contract that do not need a GPU: the reference arm (--impl reference) runs
the oracle port on the host and prints ONE JSON line with the keys the driver reads, and both arms describe
the workload with the same `config` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s"
    assert d["n_gpus"] == 2 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["global_batch"] == 64 and d["config"]["batch_per_gpu"] == 32      # the GPU arm's global batch at N=2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["detect"]["value"] > 0 and d["gpu_launches"] == 0
    # same config object as the GPU arm builds for this N
    sys.path.insert(0, ROOT)
    import bench
    from multibox_b200 import synth
    d0 = synth.make_train_inputs(**synth.TRAIN_CONFIGS["cfg2"])
    assert d["config"] == json.loads(json.dumps(bench.train_config_dict(d0, 2)))


def test_rank_other_than_zero_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
