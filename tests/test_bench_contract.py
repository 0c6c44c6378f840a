"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle on the host cores) prints ONE JSON line with the keys the
driver reads, rank != 0 prints nothing, and the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--quick", "--steps", "1", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mray/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["steps"] == 1 and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    p = _run(["--impl", "reference", "--quick", "--steps", "1", "--warmup", "1", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the CUDA arm would simply run")
    p = _run(["--quick", "--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert p.returncode != 0 and p.stdout.strip() == ""      # fails loudly, prints no bench line


def test_help_renders():
    """argparse expands % in help strings: an unescaped one only fails when somebody asks for --help."""
    p = _run(["--help"])
    assert p.returncode == 0 and "--gpus" in p.stdout, p.stderr[-1000:]


def test_clock_sampler_keeps_the_rows_of_the_timed_region(tmp_path, monkeypatch):
    """bench.ClockSampler against a stand-in nvidia-smi that needs 0.2 s to come up and then prints a row every 20 ms: rows are stamped on
    arrival, finish(t0, t1) keeps those of the timed region and parses clocks + throttle reasons."""
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.2\nwhile true; do echo '1965, 1980, 400.0, Not Active, Not Active, Not Active, Active'; sleep 0.02; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first()
    assert s.rows, "the sampler must be streaming before the first frame"
    time.sleep(0.05)
    t0 = time.perf_counter()
    time.sleep(0.1)
    t1 = time.perf_counter()
    time.sleep(0.05)
    c = s.finish(t0, t1)
    assert c["window"] == "timed region" and 2 <= c["samples"] <= 8, c
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1980.0 and c["reasons"] == ["sw_power_cap"]
    # a region too short for any row falls back to the rows taken under load around it, and says so
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first()
    time.sleep(0.03)
    now = time.perf_counter()
    c = s.finish(now - 0.0005, now - 0.0004)
    assert c["samples"] >= 1 and c["sm_mhz"] == 1965.0 and c["window"] in ("timed region", "under load around the timed region (warm-up / end-to-end frames)")
