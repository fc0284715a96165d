"""bench.py contract (GPU-less part): the reference arm -- the oracle port of the upstream eval forward on the host
cores -- prints ONE JSON line with the keys the driver reads, on BASELINE.json's metric and the native arm's workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun only rank 0 runs the host baseline; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_windows(tmp_path, monkeypatch):
    """bench.py samples nvidia-smi clocks DURING the timed region; the query process is started before the warm-up because
    its first row takes a few hundred ms.  Exercised here against a stand-in `nvidia-smi` on PATH."""
    import importlib.util
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.3\nwhile true; do echo '0, 1905, 1965, 700.1, Not Active, Not Active, Not Active, "
                    "Active'; sleep 0.05; done\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.start()
    time.sleep(0.7)                       # "warm-up"
    s.begin()
    time.sleep(0.3)                       # "timed region"
    s.end()
    time.sleep(0.2)                       # rows after the region must not count
    c = s.stop()
    assert c["window"] == "timed region" and 2 <= c["samples"] <= 8
    assert c["sm_mhz"] == 1905.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    s = bench.ClockSampler(0)             # a region too short for a single row falls back to the warm-up rows, and says so
    s.start()
    time.sleep(0.7)
    s.begin()
    s.end()
    c = s.stop()
    assert c["samples"] >= 1 and ("warm-up" in c["window"] or c["window"] == "timed region")
    monkeypatch.setenv("PATH", str(tmp_path / "missing"))
    s = bench.ClockSampler(0)             # no nvidia-smi at all: empty record, never an exception
    s.start(); s.begin(); s.end()
    assert s.stop()["samples"] == 0


def test_no_rank_conditional_collectives_in_bench():
    """Under torchrun every rank must make the same collective calls.  Static guard for the bench functions: nothing that
    contains a collective -- a trainer / sharded step, a barrier, an all-reduce / all-gather -- may sit under an
    `if rank == 0` (or `world == 1`-free rank test).  (A rank-0-only extra training step, with its gradient all-reduce
    inside, once hung a 2-GPU `--config 4` run.)"""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    collective = {"step_device", "step_e2e", "fence", "timed", "barrier", "all_reduce", "all_gather_into_tensor",
                  "sharded_forward", "run_steps", "step"}

    def calls(node):
        for n in ast.walk(node):
            if isinstance(n, ast.Call):
                f = n.func
                name = f.id if isinstance(f, ast.Name) else f.attr if isinstance(f, ast.Attribute) else None
                if name in collective:
                    yield name, n.lineno

    def mentions_rank_only(test):
        names = {n.id for n in ast.walk(test) if isinstance(n, ast.Name)}
        return "rank" in names and "world" not in names           # `rank == 0 and world == 1` legs run on a single process

    bad = []
    for fn in (n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name in ("run_train", "main", "run", "run_native")):
        for node in ast.walk(fn):
            if isinstance(node, ast.If) and mentions_rank_only(node.test):
                for stmt in node.body:
                    bad += ["%s:%d under `if %s`" % (name, line, ast.unparse(node.test)) for name, line in calls(stmt)]
    assert not bad, bad
