"""bench.py contract on the CPU: the reference arm (the unmodified reference from baseline/_ref or /root/reference on the
host cores; the oracle port only when neither is there) prints exactly one JSON line with the keys the driver reads.  The
B200 arm needs a GPU and is exercised by the driver / gpurun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--envs", "64", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/s" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    from oracle.ref_harness import reference_root
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_root() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(64), "both arms must name the workload identically"
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["gae"]["value"] > 0 and d["ppo"]["value"] > 0
    assert list(d)[-2:] == ["ppo", "e2e"], "PPO samples/s and e2e close the line"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--envs", "64",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
