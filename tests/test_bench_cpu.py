"""CPU test of bench.py's reference arm: without a GPU the rebuilt reference binaries cannot run, the arm falls back to
the host oracle port and must still print the contract's JSON line (one line, same metric / unit / config keys as the
engine's arm, `impl`, `cpu_baseline`, `e2e`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")           # what torchrun exports to its workers: the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "spgemm_gflops" and line["unit"] == "GFLOP/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("config1:") and line["config"]["product"] == "A^2"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["unit"] == "GFLOP/s" and cb["value"] == line["value"]
    if cb["kind"] == "port":
        assert cb["cores"] == (os.cpu_count() or 1)       # all host cores, whatever OMP_NUM_THREADS said
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
