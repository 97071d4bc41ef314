"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the contract's
keys, on the same metric / unit / config as the CUDA arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "iS-DQN K=9 learner updates/sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0
    assert d["config"]["batch"] == 32 and d["config"]["K"] == 9 and d["config"]["features"] == [32, 64, 64, 512]
    # same workload as the CUDA arm, and the capacity the CPU arm REALLY held (a bounded sample), stated as such
    assert d["config"]["replay_capacity"] == 20_000 and "1000000" in d["config"]["replay_capacity_note"]
    rp = d["cpu_baseline"]["replay"]  # the replay half of the metric on the host cores
    assert rp["rb_sample32_samples_per_s"] > 0 and rp["sumtree_query32_samples_per_s"] > 0 and rp["sumtree_set32_leaves_per_s"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                          "--warmup", "3"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
