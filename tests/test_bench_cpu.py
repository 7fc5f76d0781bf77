"""CPU: the contract surface of bench.py that does not need a GPU -- the reference arm (`--impl reference`, the reference's own
functors compiled in place when oracle/_ref exists, else the restatement; the one place bench.py executes oracle/) prints one
JSON line with the agreed keys, non-zero ranks of a torchrun launch stay silent, and the product arm refuses to run without a
CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
ARGS = ["--impl", "reference", "--workload", "rh54_cubed5", "--steps", "2", "--warmup", "1", "--cpu-sample", "512"]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, BENCH, *ARGS], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fp64_pair_interactions_per_s" and d["unit"] == "interactions/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "rh54_cubed5" and d["config"]["stepper"] == "bve_rk4"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "vertex targets (BVEVertexVelocity)" in cb["sample"] and "face targets (BVEFaceVelocity" in cb["sample"]
    # the same config keys as the product arm prints (the driver compares the two objects)
    assert set(d["config"]) == {"workload", "description", "stepper", "evals_per_step", "n_verts", "n_faces", "n_leaf_sources",
                                "interactions_per_eval", "dt", "Omega", "parallelism", "l2"}
    assert "full_step_ms_extrapolated" in d and d["ms_per_step"] < d["full_step_ms_extrapolated"]
    assert d["e2e"] == {"value": d["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_all_host_threads_under_torchrun_and_is_bounded():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers (round 1: the 32-core arm ran on one thread and timed out at
    N = 2, 4, 8).  The arm sets the team size itself and sizes its sample by a calibration run (--ref-seconds per step)."""
    import time
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="4", OMP_NUM_THREADS="1")
    t0 = time.time()
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "rh54_cubed5", "--steps", "2", "--warmup", "1",
                        "--gpus", "4", "--ref-seconds", "0.5"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert d["cpu_baseline"]["cores"] == want and d["n_gpus"] == 4
    assert "over 4 GPU(s)" in d["config"]["parallelism"]
    assert time.time() - t0 < 120


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, BENCH, *ARGS, "--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    from conftest import HAVE_GPU
    if HAVE_GPU:
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, BENCH, "--workload", "rh54_cubed5", "--steps", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
