"""bench.py's reference arm runs on the host cores only (CPU oracle), so its JSON contract can be checked on the
GPU-less box: one line, the keys the driver reads, the cpu_baseline / e2e objects of the tier contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("PVConv fwd+bwd points/sec") and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "B=2" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["data"] == "synthetic"


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun only rank 0 measures and prints; the other ranks exit 0 without work."""
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_informational_subresults_never_raise_and_respect_their_time_budget():
    """configs / classic-op sub-results run in child processes under one deadline; without a GPU every child fails, which
    must surface as an 'unavailable' entry (never an exception, never past the budget): the driver-read line survives."""
    import time
    import types
    sys.path.insert(0, ROOT)
    import bench
    t0 = time.perf_counter()
    out = bench.configs_subresults(types.SimpleNamespace(precision="fp32"), t0 + 75)
    assert set(out) == {"columns", "s3dis_pvcnn", "shapenet_c0p25_train", "pvcnn2", "frustum_pvcnne"}
    assert all(isinstance(out[k], str) for k in out if k != "columns")          # 'unavailable: ...' / 'time budget spent'
    ops = bench.ops_subresults(time.perf_counter() + 5)                          # < 30 s left: not even started
    assert ops == {"unavailable": "sub-result time budget spent"}
    assert time.perf_counter() - t0 < 150


def test_bench_control_flow_under_a_fake_cuda_layer():
    """tests/tools/dry_run_bench_on_cpu.py: the default arm's timed loop, its double-buffered end-to-end loop (events, one
    side stream per direction), the mode sub-results, the JSON line and the CUDA-graph probe run to completion against
    no-op streams / events and a CPU stand-in block -- Python-level protection of the artefact the driver reads."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "dry_run_bench_on_cpu.py")],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 4
    line, probe, cfg_eval, cfg_train = lines
    for cfg in (cfg_eval, cfg_train):      # the children behind the `configs` sub-result (bench.configs_subresults reads these)
        assert cfg["ms_per_step"] > 0 and "comparison_arm" in cfg and "gpu_launches" in cfg and "metric" in cfg
    assert cfg_eval["cuda_graph"]["ms_per_step"] > 0 and cfg_train["cuda_graph"] is None
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "modes", "peaks"):
        assert k in line, k
    assert set(line["modes"]) == {"tf32", "fp32_dense", "tf32_dense"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert probe["matches_eager"] is True and probe["e2e_host_output_matches"] is True


def test_bench_control_flow_two_ranks_over_gloo():
    """the same dry run as two processes (gloo standing in for NCCL): what the driver's scaling run exercises -- barrier /
    max-over-ranks timing, the flat gradient bucket's all-reduce, the strong-scaling sub-result, rank 0 printing alone --
    plus `--config` under torchrun (data-parallel train config, replica inference config)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "dry_run_bench_on_cpu.py"), "--world", "2"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 3                                   # rank 0 only, one line per bench invocation
    line, cfg_eval, cfg_train = lines
    assert line["n_gpus"] == 2 and line["config"]["parallelism"] == "dp2" and line["scaling"] == "weak"
    assert line["modes"]["strong_scaling"]["per_gpu_batch"] * 2 == line["modes"]["strong_scaling"]["global_batch"]
    assert "roofline" not in line and "cpu_baseline" not in line          # rank 0 at N=1 only
    assert cfg_eval["n_gpus"] == 2 and "replicas" in cfg_eval["config"]["parallelism"]
    assert cfg_train["n_gpus"] == 2 and "all-reduce" in cfg_train["config"]["parallelism"]
